// elementwise_kernels.cu — HBM-bound pieces of the U-Net block around the convolutions:
// layout conversion, BatchNorm3d statistics / finalize / apply (+ReLU +Dropout3d) forward and
// backward, MaxPool3d(2,2), nearest resize, global average pool, fused AdamW.
//
// Reference call sites: models/unet.py:12-14,16-18 (BN, ReLU, Dropout3d), :40,71 (MaxPool3d),
// :81-83 (F.interpolate), models/unet_dann.py:79 (GAP), train_unet.py:226,378 (AdamW).
// All activations are rows of C channels (NDHWC); every thread moves 8 channels (16 B bf16 /
// 32 B fp32) per access so that warps read and write whole 128 B lines.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxPartialBlocks = 592;  // statistics / channel sums: 4 resident CTAs per SM, one wave
constexpr int kBwdPartialBlocks = 296;  // BN backward reduction: 2 resident CTAs per SM (8 x 16 B loads in flight per thread)

// ------------------------------------------------------------------ layout conversion
template <typename T>
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int64_t N, int64_t C, int64_t S) {
  // tile transpose [C, S] -> [S, C] per sample through shared memory
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, s = s0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && s < S) ? x[(n * C + c) * S + s] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t s = s0 + i, c = c0 + threadIdx.x;
    if (s < S && c < C) y[(n * S + s) * C + c] = from_f32<T>(tile[threadIdx.x][i]);
  }
}
template <typename T>
__global__ void ndhwc_to_ncdhw_kernel(const T* __restrict__ x, float* __restrict__ y, int64_t N, int64_t C, int64_t S) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t s = s0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && s < S) ? to_f32<T>(x[(n * S + s) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, s = s0 + threadIdx.x;
    if (s < S && c < C) y[(n * C + c) * S + s] = tile[threadIdx.x][i];
  }
}
template <typename T>
__global__ void cast_from_f32_kernel(const float* __restrict__ x, T* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = from_f32<T>(x[i]);
}
template <typename T>
__global__ void cast_to_f32_kernel(const T* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = to_f32<T>(x[i]);
}

// ------------------------------------------------------------------ per-channel two-slot reductions
// A block owns `rows_per_iter = kThreads / CV` rows per iteration (CV = C/8 vector columns);
// thread (r, cv) accumulates 8 channels x 2 slots in registers, then the block reduces over r in
// shared memory and writes partials[block][2][C].  The finalize kernels sum the blocks in fixed
// order in fp64, so results are run-to-run deterministic.
struct RowMap {
  int CV, rows_per_iter, nblocks;
};
inline RowMap row_map(int64_t M, int C, int max_blocks = kMaxPartialBlocks) {
  RowMap m;
  m.CV = C / 8;
  m.rows_per_iter = kThreads / m.CV;
  if (m.rows_per_iter < 1) m.rows_per_iter = 1;
  int64_t iters = (M + m.rows_per_iter - 1) / m.rows_per_iter;
  // wide layers are small tensors: cap blocks x channels so that the finalize kernels fold at most 8192 partial sums per slot
  // (measured: 592 -> 512 blocks at C = 16, 256 at C = 32, ... is also slightly faster for the reduction passes themselves)
  const int cap = 8192 / C > 8 ? 8192 / C : 8;
  if (max_blocks > cap) max_blocks = cap;
  m.nblocks = (int)(iters < max_blocks ? (iters < 1 ? 1 : iters) : max_blocks);
  return m;
}

// U rows are in flight per thread: the loads of a batch are issued before any of them is consumed (the compiler does
// not hoist loads across the loop body by itself, see the SASS of a plain `#pragma unroll` loop).
template <int U, typename D, typename L, typename A>
__device__ __forceinline__ void block_channel_reduce(L&& load, A&& acc, int64_t M, int C, float* __restrict__ partials,
                                                     const float* __restrict__ post1 = nullptr) {
  extern __shared__ float sred[];  // [rows_per_iter][2][C]
  const int CV = C / 8;
  const int rows_per_iter = max(1, (int)blockDim.x / CV);
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  float a0[8], a1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
  if (r < rows_per_iter) {
    const int64_t stride = (int64_t)gridDim.x * rows_per_iter;
    for (int64_t row0 = (int64_t)blockIdx.x * rows_per_iter + r; row0 < M; row0 += U * stride) {
      D d[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (row0 + u * stride < M) load(row0 + u * stride, cv, d[u]);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (row0 + u * stride < M) acc(row0 + u * stride, d[u], a0, a1);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sred[(r * 2 + 0) * C + cv * 8 + i] = a0[i];
      sred[(r * 2 + 1) * C + cv * 8 + i] = a1[i];
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < 2 * C; q += blockDim.x) {
    float t = 0.f;
    for (int rr = 0; rr < rows_per_iter; ++rr) t += sred[rr * 2 * C + q];
    if (post1 && q >= C) t *= post1[q - C];
    partials[(int64_t)blockIdx.x * 2 * C + q] = t;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
bn_stats_kernel(const T* __restrict__ x, int64_t M, int C, float* __restrict__ partials) {
  // shifted sums: K[c] = x[row 0][c] is subtracted before accumulating, so that
  // var = E[(x-K)^2] - E[x-K]^2 does not cancel catastrophically when |mean| >> std
  const int cvk = threadIdx.x % (C / 8);
  float k[8];
  {
    Vec8<T> v0;
    v0.load(x + cvk * 8);
    v0.get(k);
  }
  block_channel_reduce<4, Vec8<T>>(
      [&](int64_t row, int cv, Vec8<T>& v) { v.load(x + row * C + cv * 8); },
      [&](int64_t, const Vec8<T>& v, float (&a0)[8], float (&a1)[8]) {
        float f[8];
        v.get(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = f[i] - k[i]; a0[i] += d; a1[i] = fmaf(d, d, a1[i]); }
      },
      M, C, partials);
}

// One BLOCK (kFinThreads) per channel: thread t folds blocks t, t + kFinThreads, ... (at most 3 loads per slot, all in flight
// together), then a fixed-shape warp butterfly and a fixed-order sum over the warps.  out[slot] valid in thread 0.
constexpr int kFinThreads = 256;
template <int SLOTS>
__device__ __forceinline__ void block_partial_sums(const float* __restrict__ partials, int nblocks, int C, int c, double (&out)[SLOTS]) {
  __shared__ double sm[SLOTS][kFinThreads / 32];
  double t[SLOTS];
#pragma unroll
  for (int sl = 0; sl < SLOTS; ++sl) t[sl] = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += kFinThreads)
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) t[sl] += (double)partials[(int64_t)b * 2 * C + sl * C + c];
#pragma unroll
  for (int sl = 0; sl < SLOTS; ++sl) {
    t[sl] = warp_sum_d(t[sl]);
    if ((threadIdx.x & 31) == 0) sm[sl][threadIdx.x >> 5] = t[sl];
  }
  __syncthreads();
#pragma unroll
  for (int sl = 0; sl < SLOTS; ++sl) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < kFinThreads / 32; ++w) a += sm[sl][w];
    out[sl] = a;
  }
}

template <typename T>
__global__ void bn_finalize_kernel(const float* __restrict__ partials, const T* __restrict__ x_row0, const float* __restrict__ shift_vec, int nblocks, int64_t M, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, int training, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int64_t* __restrict__ nbt,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  const int c = blockIdx.x;
  if (c == 0 && threadIdx.x == 0 && training && nbt) nbt[0] += 1;
  // the per-channel scalars are DRAM misses (touched once per step): issue their loads before the partial-sum reduction so
  // that the latencies overlap instead of forming a chain behind it (ncu: 8.4 us, almost all of it long-scoreboard stalls)
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float rm0 = running_mean ? running_mean[c] : 0.f, rv0 = running_var ? running_var[c] : 1.f;
  // the shift the partial sums were taken around: row 0 of the tensor (bn_stats) or an explicit vector (fused conv epilogue: the bias)
  const float x0 = !training ? 0.f : (shift_vec ? shift_vec[c] : (x_row0 ? to_f32<T>(x_row0[c]) : 0.f));
  float mean, var;
  if (training) {
    double sums2[2];
    block_partial_sums<2>(partials, nblocks, C, c, sums2);
    if (threadIdx.x != 0) return;
    const double s = sums2[0], ss = sums2[1];
    const double ms = s / (double)M;  // mean of the shifted values
    double v = ss / (double)M - ms * ms;
    if (v < 0.0) v = 0.0;
    const double m = ms + (double)x0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) running_mean[c] = (1.f - momentum) * rm0 + momentum * mean;
    if (running_var) {
      const float unbiased = (M > 1) ? (float)(v * (double)M / (double)(M - 1)) : var;
      running_var[c] = (1.f - momentum) * rv0 + momentum * unbiased;
    }
  } else {
    if (threadIdx.x != 0) return;
    mean = rm0;
    var = rv0;
  }
  const float invstd = 1.f / sqrtf(var + eps);
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = b;  // the affine offset; kernels evaluate (x - mean) * scale + shift, which does not cancel when |mean| >> std
  mean_out[c] = mean;
  invstd_out[c] = invstd;
}

// sample index of a row (Dropout3d masks are per (sample, channel)): 32-bit division whenever the row index allows it —
// a 64-bit division per row is ~40 instructions on kernels that otherwise run at the HBM roofline
__device__ __forceinline__ int64_t sample_of(int64_t row, int64_t S) {
  return ((row | S) >> 32) == 0 ? (int64_t)((uint32_t)row / (uint32_t)S) : row / S;
}

// Per-channel parameters of the 8 channels a thread owns, read once into registers.  In the apply / reduce kernels a
// thread keeps the same channel group cv = threadIdx.x % CV for its whole row walk (rows advance by whole blocks), so
// nothing per-channel is re-read inside the streaming loop: measured 63 -> 45 us on 2x128^3x16 bf16 (6.0 TB/s).
__device__ __forceinline__ void load8(const float* __restrict__ p, int cv, float (&o)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p + cv * 8), b = *reinterpret_cast<const float4*>(p + cv * 8 + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

template <typename T, bool DROP>
__global__ void __launch_bounds__(kThreads)
bn_act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ dropmask, int relu,
                  int64_t N, int64_t S, int C) {
  const int CV = C / 8, rpi = max(1, (int)blockDim.x / CV);
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  if (r >= rpi) return;
  float sc[8], sh[8], mu[8];
  load8(scale, cv, sc); load8(shift, cv, sh); load8(mean, cv, mu);
  const int64_t M = N * S, stride = (int64_t)gridDim.x * rpi;
  constexpr int U = 4;  // 4 x 16 B (bf16) loads in flight per thread before the first store
  for (int64_t row0 = (int64_t)blockIdx.x * rpi + r; row0 < M; row0 += U * stride) {
    Vec8<T> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (row0 + u * stride < M) v[u].load(x + (row0 + u * stride) * C + cv * 8);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + u * stride;
      if (row < M) {
        float f[8];
        v[u].get(f);
        const float* dm = DROP ? dropmask + sample_of(row, S) * C + cv * 8 : nullptr;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float t = to_f32<T>(from_f32<T>(fmaf(f[k] - mu[k], sc[k], sh[k])));  // BN output rounded to the activation dtype
          if (relu) t = fmaxf(t, 0.f);
          if (DROP) t *= dm[k];
          f[k] = t;
        }
        v[u].set(f);
        v[u].store(y + row * C + cv * 8);
      }
    }
  }
}

// g = gy * dropmask * [ReLU active]; xc = x - mean (the caller scales by invstd where it needs x-hat)
template <typename T> struct GradPair { Vec8<T> g, x; };
template <typename T, bool DROP>
__device__ __forceinline__ void bn_bwd_elem(const GradPair<T>& d, int64_t row, int cv, int C, int64_t S, const float (&sc)[8],
                                            const float (&sh)[8], const float (&mu)[8], const float* __restrict__ dropmask,
                                            int relu, float (&g)[8], float (&xc)[8]) {
  float fx[8];
  d.g.get(g);
  d.x.get(fx);
  const float* dm = DROP ? dropmask + sample_of(row, S) * C + cv * 8 : nullptr;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float gg = g[k];
    if (DROP) gg *= dm[k];
    const float dd = fx[k] - mu[k];
    if (relu) {
      const float pre = to_f32<T>(from_f32<T>(fmaf(dd, sc[k], sh[k])));
      if (!(pre > 0.f)) gg = 0.f;
    }
    g[k] = gg;
    xc[k] = dd;
  }
}

template <typename T, bool DROP>
__global__ void __launch_bounds__(kThreads, 2)
bn_act_bwd_reduce_kernel(const T* __restrict__ gy, const T* __restrict__ x, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ dropmask, int relu, int64_t N,
                         int64_t S, int C, float* __restrict__ partials) {
  const int cvk = threadIdx.x % (C / 8);
  float sc[8], sh[8], mu[8];
  load8(scale, cvk, sc); load8(shift, cvk, sh); load8(mean, cvk, mu);
  // slot 1 accumulates g * (x - mean); the per-channel factor invstd is applied once per block (post1)
  block_channel_reduce<4, GradPair<T>>(
      [&](int64_t row, int cv, GradPair<T>& d) { d.g.load(gy + row * C + cv * 8); d.x.load(x + row * C + cv * 8); },
      [&](int64_t row, const GradPair<T>& d, float (&a0)[8], float (&a1)[8]) {
        float g[8], xc[8];
        bn_bwd_elem<T, DROP>(d, row, cvk, C, S, sc, sh, mu, dropmask, relu, g, xc);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a0[i] += g[i]; a1[i] = fmaf(g[i], xc[i], a1[i]); }
      },
      N * S, C, partials, invstd);
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int nblocks, int64_t M, int C,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ sums) {
  const int c = blockIdx.x;
  double sums2[2];
  block_partial_sums<2>(partials, nblocks, C, c, sums2);
  if (threadIdx.x != 0) return;
  const double s = sums2[0], ss = sums2[1];
  if (dbeta) dbeta[c] = (float)s;
  if (dgamma) dgamma[c] = (float)ss;
  sums[c] = (float)(s / (double)M);
  sums[C + c] = (float)(ss / (double)M);
}

template <typename T, bool DROP>
__global__ void __launch_bounds__(kThreads, 3)
bn_act_bwd_apply_kernel(const T* __restrict__ gy, const T* __restrict__ x, T* __restrict__ dx,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ dropmask, int relu, const float* __restrict__ sums, int training,
                        int64_t N, int64_t S, int C) {
  const int CV = C / 8, rpi = max(1, (int)blockDim.x / CV);
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  if (r >= rpi) return;
  // dx = (g - mean(g) - xhat * mean(g*xhat)) * scale with xhat = (x - mean) * invstd:  k1 = invstd * mean(g*xhat)
  float sc[8], sh[8], mu[8], s0[8], k1[8];
  load8(scale, cv, sc); load8(shift, cv, sh); load8(mean, cv, mu);
  load8(sums, cv, s0); load8(sums + C, cv, k1);
  {
    float is[8];
    load8(invstd, cv, is);
#pragma unroll
    for (int k = 0; k < 8; ++k) k1[k] *= is[k];
  }
  const int64_t M = N * S, stride = (int64_t)gridDim.x * rpi;
  constexpr int U = 2;  // 2 rows x (gy, x) = 4 loads in flight per thread
  for (int64_t row0 = (int64_t)blockIdx.x * rpi + r; row0 < M; row0 += U * stride) {
    GradPair<T> d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + u * stride;
      if (row < M) { d[u].g.load(gy + row * C + cv * 8); d[u].x.load(x + row * C + cv * 8); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + u * stride;
      if (row < M) {
        float g[8], xc[8];
        bn_bwd_elem<T, DROP>(d[u], row, cv, C, S, sc, sh, mu, dropmask, relu, g, xc);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float t = g[k];
          if (training) t -= fmaf(xc[k], k1[k], s0[k]);
          g[k] = t * sc[k];
        }
        Vec8<T> o;
        o.set(g);
        o.store(dx + row * C + cv * 8);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
channel_sum_kernel(const T* __restrict__ x, int64_t M, int C, float* __restrict__ partials) {
  block_channel_reduce<4, Vec8<T>>(
      [&](int64_t row, int cv, Vec8<T>& v) { v.load(x + row * C + cv * 8); },
      [&](int64_t, const Vec8<T>& v, float (&a0)[8], float (&a1)[8]) {
        float f[8];
        v.get(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) a0[i] += f[i];
      },
      M, C, partials);
}
__global__ void channel_sum_finalize_kernel(const float* __restrict__ partials, int nblocks, int C, float* __restrict__ out) {
  const int c = blockIdx.x;
  double s1[1];
  block_partial_sums<1>(partials, nblocks, C, c, s1);
  if (threadIdx.x == 0) out[c] = (float)s1[0];
}

// ------------------------------------------------------------------ MaxPool3d(2,2)
__device__ __forceinline__ bool pool_better(float v, float m) { return (v > m) || (v != v); }

template <typename T>
__global__ void __launch_bounds__(kThreads)
maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int D, int H, int W, int C) {
  const int OD = D / 2, OH = H / 2, OW = W / 2, CV = C / 8;
  const int64_t total = (int64_t)N * OD * OH * OW * CV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int cv = (int)(t % CV); t /= CV;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH); t /= OH;
    const int od = (int)(t % OD);
    const int n = (int)(t / OD);
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int64_t row = (((int64_t)n * D + (2 * od + dz)) * H + (2 * oh + dy)) * W + (2 * ow + dx);
          Vec8<T> v;
          v.load(x + row * C + cv * 8);
          float f[8];
          v.get(f);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (pool_better(f[k], m[k])) m[k] = f[k];
        }
    Vec8<T> o;
    o.set(m);
    o.store(y + ((((int64_t)n * OD + od) * OH + oh) * OW + ow) * C + cv * 8);
  }
}

// BatchNorm + ReLU apply AND MaxPool3d(2,2) of the result in one pass (the encoder hand-off, models/unet.py:16-18 feeding :69-71):
// one thread owns a 2x2x2 window of an 8-channel group, reads the pre-BN tensor once, writes the activation and the pooled
// tensor — the separate pool pass re-read the 134 MB activation it had just written.  Element arithmetic = bn_act_fwd_kernel's
// (no Dropout3d here: the caller keeps the two-pass sequence when a mask is present) and maxpool2_fwd_kernel's.  Even D, H, W.
template <typename T>
__global__ void __launch_bounds__(kThreads)
bn_act_pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, T* __restrict__ pooled, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ mean, int N, int D, int H, int W, int C) {
  const int OD = D / 2, OH = H / 2, OW = W / 2, CV = C / 8;
  const int cv = threadIdx.x % CV;            // blockDim.x and the grid stride are multiples of CV: a thread keeps its channel group
  float sc[8], sh[8], mu[8];
  load8(scale, cv, sc); load8(shift, cv, sh); load8(mean, cv, mu);
  const int64_t total = (int64_t)N * OD * OH * OW * CV;
  const bool small = total < (1ll << 31);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ow, oh, od, n;
    if (small) {   // 32-bit index arithmetic whenever the tensor allows it
      uint32_t t = (uint32_t)i / (uint32_t)CV;
      ow = (int)(t % (uint32_t)OW); t /= (uint32_t)OW;
      oh = (int)(t % (uint32_t)OH); t /= (uint32_t)OH;
      od = (int)(t % (uint32_t)OD);
      n = (int)(t / (uint32_t)OD);
    } else {
      int64_t t = i / CV;
      ow = (int)(t % OW); t /= OW;
      oh = (int)(t % OH); t /= OH;
      od = (int)(t % OD);
      n = (int)(t / OD);
    }
    const int64_t row0 = (((int64_t)n * D + 2 * od) * H + 2 * oh) * W + 2 * ow;
    Vec8<T> v[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) v[p].load(x + (row0 + ((int64_t)(p >> 2) * H + ((p >> 1) & 1)) * W + (p & 1)) * C + cv * 8);
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      float f[8];
      v[p].get(f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t = fmaxf(to_f32<T>(from_f32<T>(fmaf(f[k] - mu[k], sc[k], sh[k]))), 0.f);
        f[k] = t;
        if (pool_better(t, m[k])) m[k] = t;
      }
      Vec8<T> o;
      o.set(f);
      o.store(y + (row0 + ((int64_t)(p >> 2) * H + ((p >> 1) & 1)) * W + (p & 1)) * C + cv * 8);
    }
    Vec8<T> po;
    po.set(m);
    po.store(pooled + ((((int64_t)n * OD + od) * OH + oh) * OW + ow) * C + cv * 8);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy, const T* __restrict__ gskip, T* __restrict__ gx, int N, int D, int H, int W,
                    int C) {
  const int OD = D / 2, OH = H / 2, OW = W / 2, CV = C / 8;
  const int64_t total = (int64_t)N * OD * OH * OW * CV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int cv = (int)(t % CV); t /= CV;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH); t /= OH;
    const int od = (int)(t % OD);
    const int n = (int)(t / OD);
    float m[8];
    int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { m[k] = -INFINITY; arg[k] = 0; }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int dz = p >> 2, dy = (p >> 1) & 1, dx = p & 1;
      const int64_t row = (((int64_t)n * D + (2 * od + dz)) * H + (2 * oh + dy)) * W + (2 * ow + dx);
      Vec8<T> v;
      v.load(x + row * C + cv * 8);
      float f[8];
      v.get(f);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (pool_better(f[k], m[k])) { m[k] = f[k]; arg[k] = p; }
    }
    Vec8<T> g;
    g.load(gy + ((((int64_t)n * OD + od) * OH + oh) * OW + ow) * C + cv * 8);
    float gf[8];
    g.get(gf);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int dz = p >> 2, dy = (p >> 1) & 1, dx = p & 1;
      const int64_t row = (((int64_t)n * D + (2 * od + dz)) * H + (2 * oh + dy)) * W + (2 * ow + dx);
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (arg[k] == p) ? gf[k] : 0.f;
      if (gskip) {  // fused gradient accumulation of the skip connection: gx = gskip + scatter(gy), rounded once like torch's add
        Vec8<T> sv;
        sv.load(gskip + row * C + cv * 8);
        float sf[8];
        sv.get(sf);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += sf[k];
      }
      Vec8<T> ov;
      ov.set(o);
      ov.store(gx + row * C + cv * 8);
    }
  }
}

// ------------------------------------------------------------------ nearest resize (F.interpolate default)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  const int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}
template <typename T>
__global__ void nearest_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int D, int H, int W, int OD, int OH,
                                   int OW, int C) {
  const int CV = C / 8;
  const float sd = (float)D / OD, sh = (float)H / OH, sw = (float)W / OW;
  const int64_t total = (int64_t)N * OD * OH * OW * CV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int cv = (int)(t % CV); t /= CV;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH); t /= OH;
    const int od = (int)(t % OD);
    const int n = (int)(t / OD);
    const int64_t src = (((int64_t)n * D + nearest_src(od, sd, D)) * H + nearest_src(oh, sh, H)) * W + nearest_src(ow, sw, W);
    Vec8<T> v;
    v.load(x + src * C + cv * 8);
    v.store(y + ((((int64_t)n * OD + od) * OH + oh) * OW + ow) * C + cv * 8);
  }
}
__device__ __forceinline__ void nearest_dst_range(int src, float scale, int in_size, int out_size, int& lo, int& hi) {
  // all dst with nearest_src(dst) == src form a contiguous range; scan a small window around src/scale
  int guess = (int)((float)src / scale);
  int a = guess - 2 < 0 ? 0 : guess - 2;
  lo = out_size; hi = 0;
  for (int d = a; d < out_size && d <= guess + 3; ++d)
    if (nearest_src(d, scale, in_size) == src) { if (d < lo) lo = d; hi = d + 1; }
  if (lo > hi) { lo = 0; hi = 0; }
}
template <typename T>
__global__ void nearest_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, int N, int D, int H, int W, int OD, int OH,
                                   int OW, int C) {
  const int CV = C / 8;
  const float sd = (float)D / OD, sh = (float)H / OH, sw = (float)W / OW;
  const int64_t total = (int64_t)N * D * H * W * CV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int cv = (int)(t % CV); t /= CV;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H); t /= H;
    const int d = (int)(t % D);
    const int n = (int)(t / D);
    int d0, d1, h0, h1, w0, w1;
    nearest_dst_range(d, sd, D, OD, d0, d1);
    nearest_dst_range(h, sh, H, OH, h0, h1);
    nearest_dst_range(w, sw, W, OW, w0, w1);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int a = d0; a < d1; ++a)
      for (int b = h0; b < h1; ++b)
        for (int c = w0; c < w1; ++c) {
          Vec8<T> v;
          v.load(gy + ((((int64_t)n * OD + a) * OH + b) * OW + c) * C + cv * 8);
          float f[8];
          v.get(f);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
    Vec8<T> o;
    o.set(acc);
    o.store(gx + ((((int64_t)n * D + d) * H + h) * W + w) * C + cv * 8);
  }
}

// ------------------------------------------------------------------ global average pool
template <typename T>
__global__ void gap_fwd_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t N, int64_t S, int C) {
  // one block per (n, channel-vector); threads stride over S
  const int CV = C / 8;
  const int64_t n = blockIdx.x / CV;
  const int cv = blockIdx.x % CV;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int64_t s = threadIdx.x; s < S; s += blockDim.x) {
    Vec8<T> v;
    v.load(x + (n * S + s) * C + cv * 8);
    float f[8];
    v.get(f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
  __shared__ float red[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    acc[k] = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
    out[n * C + cv * 8 + threadIdx.x] = t / (float)S;
  }
}
template <typename T>
__global__ void gap_bwd_kernel(const float* __restrict__ gout, T* __restrict__ gx, int accumulate, int64_t N, int64_t S, int C) {
  const int CV = C / 8;
  const int64_t total = N * S * CV;
  const float inv = 1.f / (float)S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const int64_t row = i / CV;
    const int64_t n = row / S;
    float f[8];
    if (accumulate) {
      Vec8<T> v;
      v.load(gx + row * C + cv * 8);
      v.get(f);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] += gout[n * C + cv * 8 + k] * inv;
    Vec8<T> o;
    o.set(f);
    o.store(gx + row * C + cv * 8);
  }
}

__global__ void scale_f32_kernel(const float* __restrict__ x, float* __restrict__ out, float alpha, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = alpha * x[i];
}

// ------------------------------------------------------------------ fused AdamW over a flat buffer
__global__ void adamw_prepare_kernel(int64_t* step, float beta1, float beta2, float* hyper) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int64_t t = step[0] + 1;
    step[0] = t;
    hyper[1] = 1.f - powf(beta1, (float)t);
    hyper[2] = 1.f - powf(beta2, (float)t);
  }
}
__global__ void __launch_bounds__(kThreads)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  int64_t n, const float* __restrict__ hyper, float beta1, float beta2, float eps, float wd,
                  float grad_scale) {
  const float lr = hyper[0], bc1 = hyper[1], bc2 = hyper[2];
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = 1.f - lr * wd;
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * grad_scale;
      ma[k] = beta1 * ma[k] + (1.f - beta1) * gk;
      va[k] = beta2 * va[k] + (1.f - beta2) * gk * gk;
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] = pa[k] * decay - step_size * ma[k] / denom;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gk = g[i] * grad_scale;
    m[i] = beta1 * m[i] + (1.f - beta1) * gk;
    v[i] = beta2 * v[i] + (1.f - beta2) * gk * gk;
    const float denom = sqrtf(v[i]) * inv_sqrt_bc2 + eps;
    p[i] = p[i] * decay - step_size * m[i] / denom;
  }
}

// BN apply kernels: blocks of kThreads / CV rows, 8 CTAs per SM at most
inline int apply_grid(int64_t M, int C, int ctas_per_sm) {
  const int rpi = (kThreads / (C / 8)) > 0 ? kThreads / (C / 8) : 1;
  return b200_grid_for(M, rpi, B200_NUM_SMS * ctas_per_sm);
}
inline int ew_grid(int64_t items) { return b200_grid_for(items, kThreads, B200_NUM_SMS * 16); }

int check_rows(const char* name, int64_t M, int C) {
  B200_REQUIRE(M > 0, B200_ERR_SHAPE, "%s: empty input (M=%lld)", name, (long long)M);
  B200_REQUIRE(C >= 8 && C % 8 == 0 && C <= 2048, B200_ERR_UNSUPPORTED, "%s: C=%d must be a multiple of 8 in [8,2048]", name, C);
  return B200_OK;
}

}  // namespace

// =========================================================================== exports
extern "C" int b200_ncdhw_to_ndhwc(int dtype, const float* x, void* y, int64_t N, int64_t C, int64_t S, void* stream) {
  B200_REQUIRE(x && y && N > 0 && C > 0 && S > 0, B200_ERR_SHAPE, "ncdhw_to_ndhwc: bad arguments");
  B200_REQUIRE(N <= 65535, B200_ERR_UNSUPPORTED, "ncdhw_to_ndhwc: N too large");
  dim3 grid((unsigned)((S + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N), block(32, 8);
  B200_DISPATCH_DTYPE(dtype, T, (ncdhw_to_ndhwc_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>(x, (T*)y, N, C, S)));
  B200_CHECK_LAUNCH("ncdhw_to_ndhwc");
  return B200_OK;
}
extern "C" int b200_ndhwc_to_ncdhw(int dtype, const void* x, float* y, int64_t N, int64_t C, int64_t S, void* stream) {
  B200_REQUIRE(x && y && N > 0 && C > 0 && S > 0, B200_ERR_SHAPE, "ndhwc_to_ncdhw: bad arguments");
  B200_REQUIRE(N <= 65535, B200_ERR_UNSUPPORTED, "ndhwc_to_ncdhw: N too large");
  dim3 grid((unsigned)((S + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N), block(32, 8);
  B200_DISPATCH_DTYPE(dtype, T, (ndhwc_to_ncdhw_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>((const T*)x, y, N, C, S)));
  B200_CHECK_LAUNCH("ndhwc_to_ncdhw");
  return B200_OK;
}
extern "C" int b200_cast_from_f32(int dtype, const float* x, void* y, int64_t n, void* stream) {
  B200_REQUIRE(x && y && n > 0, B200_ERR_SHAPE, "cast_from_f32: bad arguments");
  B200_DISPATCH_DTYPE(dtype, T, (cast_from_f32_kernel<T><<<ew_grid(n), kThreads, 0, (cudaStream_t)stream>>>(x, (T*)y, n)));
  B200_CHECK_LAUNCH("cast_from_f32");
  return B200_OK;
}
extern "C" int b200_cast_to_f32(int dtype, const void* x, float* y, int64_t n, void* stream) {
  B200_REQUIRE(x && y && n > 0, B200_ERR_SHAPE, "cast_to_f32: bad arguments");
  B200_DISPATCH_DTYPE(dtype, T, (cast_to_f32_kernel<T><<<ew_grid(n), kThreads, 0, (cudaStream_t)stream>>>((const T*)x, y, n)));
  B200_CHECK_LAUNCH("cast_to_f32");
  return B200_OK;
}

extern "C" int64_t b200_bn_partials_bytes(int C) { return (int64_t)kMaxPartialBlocks * 2 * C * sizeof(float); }

extern "C" int b200_bn_stats(int dtype, const void* x, int64_t M, int C, float* partials, void* stream) {
  int rc = check_rows("bn_stats", M, C);
  if (rc) return rc;
  B200_REQUIRE(x && partials, B200_ERR_SHAPE, "bn_stats: null pointer");
  const RowMap rm = row_map(M, C);
  const size_t smem = (size_t)rm.rows_per_iter * 2 * C * sizeof(float);
  B200_DISPATCH_DTYPE(dtype, T, (bn_stats_kernel<T><<<rm.nblocks, kThreads, smem, (cudaStream_t)stream>>>((const T*)x, M, C, partials)));
  B200_CHECK_LAUNCH("bn_stats");
  return B200_OK;
}

extern "C" int b200_bn_finalize(int dtype, const void* x, const float* partials, int64_t M, int C, const float* gamma, const float* beta,
                                float eps, float momentum, int training, float* running_mean, float* running_var,
                                int64_t* num_batches_tracked, float* scale, float* shift, float* mean, float* invstd,
                                void* stream) {
  int rc = check_rows("bn_finalize", M, C);
  if (rc) return rc;
  B200_REQUIRE(scale && shift && mean && invstd, B200_ERR_SHAPE, "bn_finalize: null output");
  B200_REQUIRE(training ? partials != nullptr : (running_mean && running_var), B200_ERR_SHAPE,
               "bn_finalize: %s", training ? "partials required in training mode" : "running stats required in eval mode");
  const RowMap rm = row_map(M, C);
  B200_REQUIRE(!training || x != nullptr, B200_ERR_SHAPE, "bn_finalize: x (the tensor bn_stats ran on) is required in training mode");
  B200_DISPATCH_DTYPE(dtype, T, (bn_finalize_kernel<T><<<C, kFinThreads, 0, (cudaStream_t)stream>>>(
                                    partials, (const T*)x, nullptr, rm.nblocks, M, C, gamma, beta, eps, momentum, training, running_mean, running_var,
                                    num_batches_tracked, scale, shift, mean, invstd)));
  B200_CHECK_LAUNCH("bn_finalize");
  return B200_OK;
}

// training-mode finalize over `nblocks` partial rows whose sums were taken around `shift_vec[c]` (NULL = 0): the partials of
// b200_conv3d_k3_bnstats (nblocks = its return value, shift_vec = the convolution bias)
extern "C" int b200_bn_finalize_ex(const float* partials, int nblocks, const float* shift_vec, int64_t M, int C, const float* gamma,
                                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                   int64_t* num_batches_tracked, float* scale, float* shift, float* mean, float* invstd, void* stream) {
  int rc = check_rows("bn_finalize_ex", M, C);
  if (rc) return rc;
  B200_REQUIRE(partials && nblocks > 0 && scale && shift && mean && invstd, B200_ERR_SHAPE, "bn_finalize_ex: null pointer / no partial rows");
  bn_finalize_kernel<float><<<C, kFinThreads, 0, (cudaStream_t)stream>>>(partials, nullptr, shift_vec, nblocks, M, C, gamma, beta, eps, momentum, 1,
                                                                        running_mean, running_var, num_batches_tracked, scale, shift, mean, invstd);
  B200_CHECK_LAUNCH("bn_finalize_ex");
  return B200_OK;
}

extern "C" int b200_bn_act_fwd(int dtype, const void* x, void* y, const float* scale, const float* shift, const float* mean,
                               const float* dropmask, int relu, int64_t N, int64_t S, int C, void* stream) {
  int rc = check_rows("bn_act_fwd", N * S, C);
  if (rc) return rc;
  B200_REQUIRE(x && y && scale && shift && mean, B200_ERR_SHAPE, "bn_act_fwd: null pointer");
  const int grid = apply_grid(N * S, C, 8);  // 4 resident CTAs per SM, two waves
  if (dropmask) {
    B200_DISPATCH_DTYPE(dtype, T, (bn_act_fwd_kernel<T, true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
                                      (const T*)x, (T*)y, scale, shift, mean, dropmask, relu, N, S, C)));
  } else {
    B200_DISPATCH_DTYPE(dtype, T, (bn_act_fwd_kernel<T, false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
                                      (const T*)x, (T*)y, scale, shift, mean, dropmask, relu, N, S, C)));
  }
  B200_CHECK_LAUNCH("bn_act_fwd");
  return B200_OK;
}

extern "C" int b200_bn_act_bwd_reduce(int dtype, const void* gy, const void* x, const float* scale, const float* shift,
                                      const float* mean, const float* invstd, const float* dropmask, int relu, int64_t N,
                                      int64_t S, int C, float* partials, void* stream) {
  int rc = check_rows("bn_act_bwd_reduce", N * S, C);
  if (rc) return rc;
  B200_REQUIRE(gy && x && scale && shift && mean && invstd && partials, B200_ERR_SHAPE, "bn_act_bwd_reduce: null pointer");
  const RowMap rm = row_map(N * S, C, kBwdPartialBlocks);
  const size_t smem = (size_t)rm.rows_per_iter * 2 * C * sizeof(float);
  if (dropmask) {
    B200_DISPATCH_DTYPE(dtype, T, (bn_act_bwd_reduce_kernel<T, true><<<rm.nblocks, kThreads, smem, (cudaStream_t)stream>>>(
                                      (const T*)gy, (const T*)x, scale, shift, mean, invstd, dropmask, relu, N, S, C, partials)));
  } else {
    B200_DISPATCH_DTYPE(dtype, T, (bn_act_bwd_reduce_kernel<T, false><<<rm.nblocks, kThreads, smem, (cudaStream_t)stream>>>(
                                      (const T*)gy, (const T*)x, scale, shift, mean, invstd, dropmask, relu, N, S, C, partials)));
  }
  B200_CHECK_LAUNCH("bn_act_bwd_reduce");
  return B200_OK;
}

extern "C" int b200_bn_bwd_finalize(const float* partials, int64_t M, int C, float* dgamma, float* dbeta, float* sums,
                                    void* stream) {
  int rc = check_rows("bn_bwd_finalize", M, C);
  if (rc) return rc;
  B200_REQUIRE(partials && sums, B200_ERR_SHAPE, "bn_bwd_finalize: null pointer");
  const RowMap rm = row_map(M, C, kBwdPartialBlocks);
  bn_bwd_finalize_kernel<<<C, kFinThreads, 0, (cudaStream_t)stream>>>(partials, rm.nblocks, M, C, dgamma, dbeta, sums);
  B200_CHECK_LAUNCH("bn_bwd_finalize");
  return B200_OK;
}

// same over an explicit number of partial rows (the fused head's backward writes b200_head_blocks() rows)
extern "C" int b200_bn_bwd_finalize_ex(const float* partials, int nblocks, int64_t M, int C, float* dgamma, float* dbeta, float* sums,
                                       void* stream) {
  int rc = check_rows("bn_bwd_finalize_ex", M, C);
  if (rc) return rc;
  B200_REQUIRE(partials && sums && nblocks > 0, B200_ERR_SHAPE, "bn_bwd_finalize_ex: null pointer / no partial rows");
  bn_bwd_finalize_kernel<<<C, kFinThreads, 0, (cudaStream_t)stream>>>(partials, nblocks, M, C, dgamma, dbeta, sums);
  B200_CHECK_LAUNCH("bn_bwd_finalize_ex");
  return B200_OK;
}

extern "C" int b200_bn_act_bwd_apply(int dtype, const void* gy, const void* x, void* dx, const float* scale,
                                     const float* shift, const float* mean, const float* invstd, const float* dropmask,
                                     int relu, const float* sums, int training, int64_t N, int64_t S, int C, void* stream) {
  int rc = check_rows("bn_act_bwd_apply", N * S, C);
  if (rc) return rc;
  B200_REQUIRE(gy && x && dx && scale && shift && mean && invstd && sums, B200_ERR_SHAPE, "bn_act_bwd_apply: null pointer");
  const int grid = apply_grid(N * S, C, 3);  // __launch_bounds__(kThreads, 3): every CTA resident, one wave
  if (dropmask) {
    B200_DISPATCH_DTYPE(dtype, T, (bn_act_bwd_apply_kernel<T, true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
                                      (const T*)gy, (const T*)x, (T*)dx, scale, shift, mean, invstd, dropmask, relu, sums,
                                      training, N, S, C)));
  } else {
    B200_DISPATCH_DTYPE(dtype, T, (bn_act_bwd_apply_kernel<T, false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
                                      (const T*)gy, (const T*)x, (T*)dx, scale, shift, mean, invstd, dropmask, relu, sums,
                                      training, N, S, C)));
  }
  B200_CHECK_LAUNCH("bn_act_bwd_apply");
  return B200_OK;
}

extern "C" int b200_channel_sum(int dtype, const void* x, int64_t M, int C, float* partials, float* out, void* stream) {
  int rc = check_rows("channel_sum", M, C);
  if (rc) return rc;
  B200_REQUIRE(x && partials && out, B200_ERR_SHAPE, "channel_sum: null pointer");
  const RowMap rm = row_map(M, C);
  const size_t smem = (size_t)rm.rows_per_iter * 2 * C * sizeof(float);
  B200_DISPATCH_DTYPE(dtype, T, (channel_sum_kernel<T><<<rm.nblocks, kThreads, smem, (cudaStream_t)stream>>>((const T*)x, M, C, partials)));
  B200_CHECK_LAUNCH("channel_sum");
  channel_sum_finalize_kernel<<<C, kFinThreads, 0, (cudaStream_t)stream>>>(partials, rm.nblocks, C, out);
  B200_CHECK_LAUNCH("channel_sum_finalize");
  return B200_OK;
}

extern "C" int b200_maxpool2_fwd(int dtype, const void* x, void* y, int N, int D, int H, int W, int C, void* stream) {
  int rc = check_rows("maxpool2_fwd", (int64_t)N * D * H * W, C);
  if (rc) return rc;
  B200_REQUIRE(x && y && D >= 2 && H >= 2 && W >= 2, B200_ERR_SHAPE, "maxpool2_fwd: spatial dims must be >= 2");
  const int64_t items = (int64_t)N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
  B200_DISPATCH_DTYPE(dtype, T, (maxpool2_fwd_kernel<T><<<ew_grid(items), kThreads, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, N, D, H, W, C)));
  B200_CHECK_LAUNCH("maxpool2_fwd");
  return B200_OK;
}

// y = relu(bn(x)) and pooled = MaxPool3d(2,2)(y) in one pass; D, H, W even and C/8 dividing the block size (else: bn_act_fwd + maxpool2_fwd)
extern "C" int b200_bn_act_pool_fwd_supported(int D, int H, int W, int C) {
  const int CV = C / 8;
  return C % 8 == 0 && CV >= 1 && kThreads % CV == 0 && D >= 2 && H >= 2 && W >= 2 && ((D | H | W) & 1) == 0;
}
extern "C" int b200_bn_act_pool_fwd(int dtype, const void* x, void* y, void* pooled, const float* scale, const float* shift, const float* mean,
                                    int N, int D, int H, int W, int C, void* stream) {
  int rc = check_rows("bn_act_pool_fwd", (int64_t)N * D * H * W, C);
  if (rc) return rc;
  B200_REQUIRE(b200_bn_act_pool_fwd_supported(D, H, W, C), B200_ERR_UNSUPPORTED, "bn_act_pool_fwd: needs even D/H/W and C/8 dividing %d (C=%d, %dx%dx%d)", kThreads, C, D, H, W);
  B200_REQUIRE(x && y && pooled && scale && shift && mean, B200_ERR_SHAPE, "bn_act_pool_fwd: null pointer");
  const int64_t items = (int64_t)N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
  B200_DISPATCH_DTYPE(dtype, T, (bn_act_pool_fwd_kernel<T><<<b200_grid_for(items, kThreads, B200_NUM_SMS * 4), kThreads, 0, (cudaStream_t)stream>>>(
                                    (const T*)x, (T*)y, (T*)pooled, scale, shift, mean, N, D, H, W, C)));
  B200_CHECK_LAUNCH("bn_act_pool_fwd");
  return B200_OK;
}

static int maxpool2_bwd_impl(const char* name, int dtype, const void* x, const void* gy, const void* gskip, void* gx, int N, int D, int H, int W,
                             int C, void* stream) {
  int rc = check_rows(name, (int64_t)N * D * H * W, C);
  if (rc) return rc;
  B200_REQUIRE(x && gy && gx && D >= 2 && H >= 2 && W >= 2, B200_ERR_SHAPE, "%s: spatial dims must be >= 2", name);
  cudaStream_t st = (cudaStream_t)stream;
  if ((D | H | W) & 1) {  // voxels outside every 2x2x2 window: gradient 0 (+ the skip gradient)
    const size_t bytes = (size_t)N * D * H * W * C * (dtype == B200_F32 ? 4 : 2);
    if (gskip) B200_CUDA(cudaMemcpyAsync(gx, gskip, bytes, cudaMemcpyDeviceToDevice, st));
    else B200_CUDA(cudaMemsetAsync(gx, 0, bytes, st));
  }
  const int64_t items = (int64_t)N * (D / 2) * (H / 2) * (W / 2) * (C / 8);
  B200_DISPATCH_DTYPE(dtype, T, (maxpool2_bwd_kernel<T><<<ew_grid(items), kThreads, 0, st>>>((const T*)x, (const T*)gy, (const T*)gskip, (T*)gx, N, D, H, W, C)));
  B200_CHECK_LAUNCH(name);
  return B200_OK;
}
extern "C" int b200_maxpool2_bwd(int dtype, const void* x, const void* gy, void* gx, int N, int D, int H, int W, int C, void* stream) {
  return maxpool2_bwd_impl("maxpool2_bwd", dtype, x, gy, nullptr, gx, N, D, H, W, C, stream);
}
extern "C" int b200_maxpool2_bwd_add(int dtype, const void* x, const void* gy, const void* gskip, void* gx, int N, int D, int H, int W, int C,
                                     void* stream) {
  B200_REQUIRE(gskip != nullptr, B200_ERR_SHAPE, "maxpool2_bwd_add: null skip gradient");
  return maxpool2_bwd_impl("maxpool2_bwd_add", dtype, x, gy, gskip, gx, N, D, H, W, C, stream);
}

extern "C" int b200_nearest_resize_fwd(int dtype, const void* x, void* y, int N, int D, int H, int W, int OD, int OH, int OW,
                                       int C, void* stream) {
  int rc = check_rows("nearest_resize_fwd", (int64_t)N * D * H * W, C);
  if (rc) return rc;
  B200_REQUIRE(x && y && OD > 0 && OH > 0 && OW > 0, B200_ERR_SHAPE, "nearest_resize_fwd: bad arguments");
  const int64_t items = (int64_t)N * OD * OH * OW * (C / 8);
  B200_DISPATCH_DTYPE(dtype, T, (nearest_fwd_kernel<T><<<ew_grid(items), kThreads, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, N, D, H, W, OD, OH, OW, C)));
  B200_CHECK_LAUNCH("nearest_resize_fwd");
  return B200_OK;
}
extern "C" int b200_nearest_resize_bwd(int dtype, const void* gy, void* gx, int N, int D, int H, int W, int OD, int OH, int OW,
                                       int C, void* stream) {
  int rc = check_rows("nearest_resize_bwd", (int64_t)N * D * H * W, C);
  if (rc) return rc;
  B200_REQUIRE(gy && gx && OD > 0 && OH > 0 && OW > 0, B200_ERR_SHAPE, "nearest_resize_bwd: bad arguments");
  const int64_t items = (int64_t)N * D * H * W * (C / 8);
  B200_DISPATCH_DTYPE(dtype, T, (nearest_bwd_kernel<T><<<ew_grid(items), kThreads, 0, (cudaStream_t)stream>>>((const T*)gy, (T*)gx, N, D, H, W, OD, OH, OW, C)));
  B200_CHECK_LAUNCH("nearest_resize_bwd");
  return B200_OK;
}

extern "C" int b200_gap_fwd(int dtype, const void* x, float* out, int64_t N, int64_t S, int C, void* stream) {
  int rc = check_rows("gap_fwd", N * S, C);
  if (rc) return rc;
  B200_REQUIRE(x && out, B200_ERR_SHAPE, "gap_fwd: null pointer");
  B200_DISPATCH_DTYPE(dtype, T, (gap_fwd_kernel<T><<<(unsigned)(N * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const T*)x, out, N, S, C)));
  B200_CHECK_LAUNCH("gap_fwd");
  return B200_OK;
}
extern "C" int b200_gap_bwd(int dtype, const float* gout, void* gx, int accumulate, int64_t N, int64_t S, int C, void* stream) {
  int rc = check_rows("gap_bwd", N * S, C);
  if (rc) return rc;
  B200_REQUIRE(gout && gx, B200_ERR_SHAPE, "gap_bwd: null pointer");
  B200_DISPATCH_DTYPE(dtype, T, (gap_bwd_kernel<T><<<ew_grid(N * S * (C / 8)), kThreads, 0, (cudaStream_t)stream>>>(gout, (T*)gx, accumulate, N, S, C)));
  B200_CHECK_LAUNCH("gap_bwd");
  return B200_OK;
}
extern "C" int b200_scale_f32(const float* x, float* out, float alpha, int64_t n, void* stream) {
  B200_REQUIRE(x && out && n > 0, B200_ERR_SHAPE, "scale_f32: bad arguments");
  scale_f32_kernel<<<ew_grid(n), kThreads, 0, (cudaStream_t)stream>>>(x, out, alpha, n);
  B200_CHECK_LAUNCH("scale_f32");
  return B200_OK;
}

extern "C" int b200_adamw_prepare(int64_t* step, float beta1, float beta2, float* hyper, void* stream) {
  B200_REQUIRE(step && hyper, B200_ERR_SHAPE, "adamw_prepare: null pointer");
  adamw_prepare_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step, beta1, beta2, hyper);
  B200_CHECK_LAUNCH("adamw_prepare");
  return B200_OK;
}
extern "C" int b200_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1,
                               float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  B200_REQUIRE(p && g && m && v && hyper && n > 0, B200_ERR_SHAPE, "adamw_flat: bad arguments");
  B200_REQUIRE(b200_aligned(p, 16) && b200_aligned(g, 16) && b200_aligned(m, 16) && b200_aligned(v, 16), B200_ERR_ALIGN,
               "adamw_flat: buffers must be 16-byte aligned");
  adamw_flat_kernel<<<ew_grid(n / 4 + 1), kThreads, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hyper, beta1, beta2, eps, weight_decay, grad_scale);
  B200_CHECK_LAUNCH("adamw_flat");
  return B200_OK;
}
