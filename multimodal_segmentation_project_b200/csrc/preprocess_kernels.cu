// preprocess_kernels.cu — device-side input pipeline: the arithmetic of the reference's
// utils/dataloader.py CombinedDataset.preprocess_ct (:111-117), preprocess_mri (:128-145) and the AMOS / CHAOS label
// remaps (:162-181) on tensors that are already in HBM.
//   * CT: clip to the abdominal window and scale, float32 like numpy (true division) — bit-exact;
//   * MRI: moments (fp64 accumulation, fixed order), exact order statistics by a 3-pass radix select on the
//     order-preserving integer image of the floats (no sort, no host synchronisation between passes), then one
//     elementwise pass that repeats numpy's dtype flow (z-score in float32, clip / min-max in float64, cast);
//   * labels: ordered range table (later entries override earlier ones, like the reference's loop of masked stores).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBlocks = B200_NUM_SMS * 4;

// Elementwise float32 map with 16-byte accesses, two vectors in flight per thread; the last n % 4 elements go one by one.
template <typename F>
__device__ __forceinline__ void map_f32(const float* __restrict__ x, float* __restrict__ y, int64_t n, F&& f) {
  const int64_t nv = n / 4, stride = (int64_t)gridDim.x * blockDim.x;
  const float4* xv = reinterpret_cast<const float4*>(x);
  float4* yv = reinterpret_cast<float4*>(y);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += 2 * stride) {
    const float4 a = xv[i];
    const bool two = i + stride < nv;
    const float4 b = two ? xv[i + stride] : a;
    yv[i] = make_float4(f(a.x), f(a.y), f(a.z), f(a.w));
    if (two) yv[i + stride] = make_float4(f(b.x), f(b.y), f(b.z), f(b.w));
  }
  for (int64_t i = nv * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = f(x[i]);
}

__global__ void __launch_bounds__(kThreads)
ct_window_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float lo, float hi) {
  const float range = hi - lo;
  // np.clip, then (image - window_min) / (window_max - window_min) with a true float32 division
  map_f32(x, y, n, [=](float v) { return __fdiv_rn(fminf(fmaxf(v, lo), hi) - lo, range); });
}

// partials[block][2] = (sum, sum of squared deviations from `center`) in fp64; center = 0 for the first pass
__global__ void __launch_bounds__(kThreads)
moments_kernel(const float* __restrict__ x, int64_t n, const double* __restrict__ center, double* __restrict__ partials) {
  const double c = center ? center[0] : 0.0;
  double s = 0.0, q = 0.0;
  auto acc = [&](float v) { const double d = (double)v - c; s += d; q = fma(d, d, q); };
  const int64_t nv = n / 4, stride = (int64_t)gridDim.x * blockDim.x;
  const float4* xv = reinterpret_cast<const float4*>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float4 a = xv[i];
    acc(a.x); acc(a.y); acc(a.z); acc(a.w);
  }
  for (int64_t i = nv * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc(x[i]);
  __shared__ double sm[2][kThreads / 32];
  s = warp_sum_d(s); q = warp_sum_d(q);
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = s; sm[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { a += sm[0][w]; b += sm[1][w]; }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
  }
}
// one block: folds the per-block partials in ascending order; stage 0 -> out[0] = mean, stage 1 -> out[1] = population variance
__global__ void moments_fold_kernel(const double* __restrict__ partials, int nblocks, int64_t n, int stage, double* __restrict__ out) {
  if (threadIdx.x != 0) return;
  double a = 0.0;
  for (int b = 0; b < nblocks; ++b) a += partials[2 * b + stage];
  out[stage] = a / (double)n;
}

// order-preserving map float -> uint32 (ascending floats = ascending integers; -0 < +0, NaN last)
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Radix select state per requested rank: {prefix, remaining rank}.  Pass p looks at digit p (11, 11, 10 bits from the top).
struct SelState { uint32_t prefix; uint32_t pad; unsigned long long rank; };
__host__ __device__ constexpr int digit_bits(int pass) { return pass == 2 ? 10 : 11; }
__host__ __device__ constexpr int digit_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
constexpr int kMaxRanks = 4;   // 4 x 2048 shared-memory bins: the two neighbours of the 1st and of the 99th percentile

__global__ void select_init_kernel(const int64_t* __restrict__ ranks, int nranks, SelState* __restrict__ st, unsigned int* __restrict__ hist) {
  for (int i = threadIdx.x; i < nranks * 2048; i += blockDim.x) hist[i] = 0;
  if ((int)threadIdx.x < nranks) { st[threadIdx.x].prefix = 0; st[threadIdx.x].rank = (unsigned long long)ranks[threadIdx.x]; }
}

template <int PASS>
__global__ void __launch_bounds__(kThreads)
select_hist_kernel(const float* __restrict__ x, int64_t n, int nranks, const SelState* __restrict__ st, unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[kMaxRanks][2048];
  constexpr int bins = 1 << digit_bits(PASS);
  for (int i = threadIdx.x; i < nranks * 2048; i += blockDim.x) (&sh[0][0])[i] = 0;
  uint32_t prefix[kMaxRanks];
#pragma unroll
  for (int r = 0; r < kMaxRanks; ++r) prefix[r] = r < nranks ? st[r].prefix : 0;
  __syncthreads();
  constexpr uint32_t himask = PASS == 0 ? 0u : (0xffffffffu << (digit_shift(PASS) + digit_bits(PASS)));
  auto count = [&](float v) {
    const uint32_t k = float_key(v);
    const uint32_t digit = (k >> digit_shift(PASS)) & (bins - 1);
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < nranks && (k & himask) == prefix[r]) atomicAdd(&sh[r][digit], 1u);
  };
  const int64_t nv = n / 4, stride = (int64_t)gridDim.x * blockDim.x;
  const float4* xv = reinterpret_cast<const float4*>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float4 a = xv[i];
    count(a.x); count(a.y); count(a.z); count(a.w);
  }
  for (int64_t i = nv * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) count(x[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < nranks * 2048; i += blockDim.x) {
    const unsigned int v = (&sh[0][0])[i];
    if (v) atomicAdd(&hist[i], v);
  }
}

// one block per rank: finds the digit whose cumulative count crosses the remaining rank, extends the prefix, clears the histogram
template <int PASS>
__global__ void select_pick_kernel(int nranks, SelState* __restrict__ st, unsigned int* __restrict__ hist, float* __restrict__ out) {
  const int r = blockIdx.x;
  if (threadIdx.x == 0) {
    constexpr int bins = 1 << digit_bits(PASS);
    unsigned long long rank = st[r].rank, cum = 0;
    int d = 0;
    for (; d < bins; ++d) {
      const unsigned long long c = hist[r * 2048 + d];
      if (cum + c > rank) break;
      cum += c;
    }
    if (d == bins) d = bins - 1;
    st[r].rank = rank - cum;
    st[r].prefix |= (uint32_t)d << digit_shift(PASS);
    if (PASS == 2) out[r] = key_float(st[r].prefix);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[r * 2048 + i] = 0;
}

// params (fp64): [0] mean (float32 value), [1] denom = float32(std + 1e-8), [2] low, [3] high, [4] high - low + 1e-8
__global__ void __launch_bounds__(kThreads)
mri_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, const double* __restrict__ params) {
  const float mean = (float)params[0], denom = (float)params[1];
  const double low = params[2], high = params[3], span = params[4];
  map_f32(x, y, n, [=](float v) {
    const float z = __fdiv_rn(v - mean, denom);                    // float32, like (image - mean) / (std + 1e-8)
    const double c = fmin(fmax((double)z, low), high);             // np.clip with float64 percentiles promotes to float64
    return (float)((c - low) / span);                              // ... / (high - low + 1e-8), then astype(float32)
  });
}

struct RangeTable { int n; long long lo[8], hi[8], val[8]; };
__device__ __forceinline__ void store2(int64_t* out, int64_t pair, int64_t a, int64_t b) { reinterpret_cast<longlong2*>(out)[pair] = make_longlong2(a, b); }
__device__ __forceinline__ void store2(uint8_t* out, int64_t pair, uint8_t a, uint8_t b) { reinterpret_cast<uchar2*>(out)[pair] = make_uchar2(a, b); }
template <typename OUT>
__global__ void __launch_bounds__(kThreads)
label_remap_kernel(const int64_t* __restrict__ in, OUT* __restrict__ out, int64_t n, const RangeTable t) {
  auto map = [&](long long v) {
    long long o = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < t.n && v >= t.lo[r] && v <= t.hi[r]) o = t.val[r];
    return (OUT)o;
  };
  const int64_t nv = n / 2, stride = (int64_t)gridDim.x * blockDim.x;
  const longlong2* iv = reinterpret_cast<const longlong2*>(in);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += 2 * stride) {
    const longlong2 a = iv[i];
    const bool two = i + stride < nv;
    const longlong2 b = two ? iv[i + stride] : a;
    store2(out, i, map(a.x), map(a.y));
    if (two) store2(out, i + stride, map(b.x), map(b.y));
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = map(in[n - 1]);
}

}  // namespace

extern "C" int b200_ct_window(const float* x, float* y, int64_t n, float lo, float hi, void* stream) {
  B200_REQUIRE(x && y && n > 0 && hi > lo, B200_ERR_SHAPE, "ct_window: bad arguments");
  B200_REQUIRE(b200_aligned(x, 16) && b200_aligned(y, 16), B200_ERR_ALIGN, "ct_window: pointers must be 16-byte aligned");
  ct_window_kernel<<<b200_grid_for(n, kThreads, kBlocks * 2), kThreads, 0, (cudaStream_t)stream>>>(x, y, n, lo, hi);
  B200_CHECK_LAUNCH("ct_window");
  return B200_OK;
}

extern "C" int64_t b200_preprocess_workspace_bytes(void) {
  return (int64_t)(2 * kBlocks * sizeof(double) + 8 * sizeof(double) + kMaxRanks * sizeof(SelState) + kMaxRanks * 2048 * sizeof(unsigned int));
}

/* out[0] = mean, out[1] = population variance (both fp64, device memory) */
extern "C" int b200_moments_f32(const float* x, int64_t n, void* workspace, double* out, void* stream) {
  B200_REQUIRE(x && workspace && out && n > 0, B200_ERR_SHAPE, "moments_f32: bad arguments");
  B200_REQUIRE(b200_aligned(x, 16), B200_ERR_ALIGN, "moments_f32: x must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  double* partials = (double*)workspace;
  const int g = b200_grid_for(n, kThreads, kBlocks);
  moments_kernel<<<g, kThreads, 0, st>>>(x, n, nullptr, partials);
  moments_fold_kernel<<<1, 32, 0, st>>>(partials, g, n, 0, out);
  moments_kernel<<<g, kThreads, 0, st>>>(x, n, out, partials);
  moments_fold_kernel<<<1, 32, 0, st>>>(partials, g, n, 1, out);
  B200_CHECK_LAUNCH("moments_f32");
  return B200_OK;
}

/* values[r] = the element of rank ranks[r] (0-based, ascending) of x[0..n); ranks and values in device memory */
extern "C" int b200_select_ranks_f32(const float* x, int64_t n, const int64_t* ranks, int nranks, float* values, void* workspace, void* stream) {
  B200_REQUIRE(x && ranks && values && workspace && n > 0, B200_ERR_SHAPE, "select_ranks_f32: bad arguments");
  B200_REQUIRE(nranks >= 1 && nranks <= kMaxRanks, B200_ERR_UNSUPPORTED, "select_ranks_f32: 1..%d ranks", kMaxRanks);
  B200_REQUIRE(b200_aligned(x, 16), B200_ERR_ALIGN, "select_ranks_f32: x must be 16-byte aligned");
  B200_REQUIRE(n < (int64_t)4294967296LL, B200_ERR_UNSUPPORTED, "select_ranks_f32: at most 2^32 - 1 elements");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = (uint8_t*)workspace + 2 * kBlocks * sizeof(double) + 8 * sizeof(double);
  SelState* state = (SelState*)w;
  unsigned int* hist = (unsigned int*)(w + kMaxRanks * sizeof(SelState));
  const int g = b200_grid_for(n, kThreads, kBlocks);
  select_init_kernel<<<1, 1024, 0, st>>>(ranks, nranks, state, hist);
  select_hist_kernel<0><<<g, kThreads, 0, st>>>(x, n, nranks, state, hist);
  select_pick_kernel<0><<<nranks, 256, 0, st>>>(nranks, state, hist, values);
  select_hist_kernel<1><<<g, kThreads, 0, st>>>(x, n, nranks, state, hist);
  select_pick_kernel<1><<<nranks, 256, 0, st>>>(nranks, state, hist, values);
  select_hist_kernel<2><<<g, kThreads, 0, st>>>(x, n, nranks, state, hist);
  select_pick_kernel<2><<<nranks, 256, 0, st>>>(nranks, state, hist, values);
  B200_CHECK_LAUNCH("select_ranks_f32");
  return B200_OK;
}

extern "C" int b200_mri_normalize(const float* x, float* y, int64_t n, const double* params, void* stream) {
  B200_REQUIRE(x && y && params && n > 0, B200_ERR_SHAPE, "mri_normalize: bad arguments");
  B200_REQUIRE(b200_aligned(x, 16) && b200_aligned(y, 16), B200_ERR_ALIGN, "mri_normalize: pointers must be 16-byte aligned");
  mri_normalize_kernel<<<b200_grid_for(n, kThreads, kBlocks * 2), kThreads, 0, (cudaStream_t)stream>>>(x, y, n, params);
  B200_CHECK_LAUNCH("mri_normalize");
  return B200_OK;
}

/* out = 0; for r in order: if lo[r] <= in <= hi[r]: out = val[r].  out_u8 != 0 writes uint8 labels (1 B/voxel for the loss kernels) */
extern "C" int b200_label_remap(const int64_t* in, void* out, int64_t n, const int64_t* lo, const int64_t* hi, const int64_t* val, int nranges,
                                int out_u8, void* stream) {
  B200_REQUIRE(in && out && n > 0 && lo && hi && val, B200_ERR_SHAPE, "label_remap: bad arguments");
  B200_REQUIRE(nranges >= 0 && nranges <= 8, B200_ERR_UNSUPPORTED, "label_remap: at most 8 ranges");
  B200_REQUIRE(b200_aligned(in, 16), B200_ERR_ALIGN, "label_remap: input must be 16-byte aligned");
  RangeTable t;
  t.n = nranges;
  for (int r = 0; r < 8; ++r) { t.lo[r] = r < nranges ? lo[r] : 1; t.hi[r] = r < nranges ? hi[r] : 0; t.val[r] = r < nranges ? val[r] : 0; }
  const int g = b200_grid_for(n, kThreads, kBlocks * 2);
  if (out_u8) label_remap_kernel<uint8_t><<<g, kThreads, 0, (cudaStream_t)stream>>>(in, (uint8_t*)out, n, t);
  else label_remap_kernel<int64_t><<<g, kThreads, 0, (cudaStream_t)stream>>>(in, (int64_t*)out, n, t);
  B200_CHECK_LAUNCH("label_remap");
  return B200_OK;
}
