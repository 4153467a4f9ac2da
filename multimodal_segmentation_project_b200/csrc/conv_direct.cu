// conv_direct.cu — CUDA-core implicit-GEMM kernels (fp32 accumulate) for
//   * nn.Conv3d(k=3,p=1) forward / data gradient / weight gradient   models/unet.py:11,15
//   * nn.ConvTranspose3d(k=2,s=2) forward / data / weight gradients   models/unet.py:56-58,79
// in channels-last (NDHWC) layout, dtype fp32 or bf16.
//
// These are the exact-arithmetic path (fp32 parity anchor, BASELINE config #1), the
// HBM-bound layers (Cin = 1 first conv, the up-convolutions) and the checker the tcgen05
// kernels in conv_tc.cu are validated against on the GPU.  The virtual concat
// (two input tensors, models/unet.py:84) and the un-concat of the data gradient (two
// output tensors) are done in the operand gather / the epilogue: the concatenated tensor is
// never materialised.
//
// Two tile kernels:
//   rows_gemm : C[m, n] = sum_k A(m, k) B(k, n),  m = voxel rows (large), A gathered
//   cols_gemm : C[r, c] = sum_v A(v, r) B(v, c),  v = voxel rows (reduction, split over CTAs)
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int kThreads = 256;
constexpr int BM = 128;  // rows per tile
constexpr int BK = 16;   // reduction slice

struct Geom {
  int N, D, H, W;
  __host__ __device__ int64_t rows() const { return (int64_t)N * D * H * W; }
};

__device__ __forceinline__ void decode_row(const Geom& g, int64_t m, int& n, int& d, int& h, int& w) {
  w = (int)(m % g.W); m /= g.W;
  h = (int)(m % g.H); m /= g.H;
  d = (int)(m % g.D);
  n = (int)(m / g.D);
}

// ---------------------------------------------------------------- A gathers for rows_gemm
// Each returns 8 consecutive k values (k8 .. k8+7) of row m as floats.
template <typename T>
struct GatherK3 {  // 3x3x3 taps over [x0 | x1]
  const T* x0; const T* x1; int c0, c1; Geom g;
  __device__ __forceinline__ int K() const { return 27 * (c0 + c1); }
  template <bool VEC>
  __device__ __forceinline__ void load(int n, int d, int h, int w, bool row_ok, int k8, float (&f)[8]) const {
    const int Cin = c0 + c1;
    if (VEC) {
      const int tap = k8 / Cin, ci = k8 - tap * Cin;
      const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
      const int dd = d + kd - 1, hh = h + kh - 1, ww = w + kw - 1;
      const bool ok = row_ok && tap < 27 && (unsigned)dd < (unsigned)g.D && (unsigned)hh < (unsigned)g.H && (unsigned)ww < (unsigned)g.W;
      if (ok) {
        const int64_t row = (((int64_t)n * g.D + dd) * g.H + hh) * g.W + ww;
        Vec8<T> v;
        if (ci < c0) v.load(x0 + row * c0 + ci); else v.load(x1 + row * c1 + (ci - c0));
        v.get(f);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k8 + i;
        const int tap = k / Cin, ci = k - tap * Cin;
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        const int dd = d + kd - 1, hh = h + kh - 1, ww = w + kw - 1;
        const bool ok = row_ok && tap < 27 && (unsigned)dd < (unsigned)g.D && (unsigned)hh < (unsigned)g.H && (unsigned)ww < (unsigned)g.W;
        float v = 0.f;
        if (ok) {
          const int64_t row = (((int64_t)n * g.D + dd) * g.H + hh) * g.W + ww;
          v = ci < c0 ? to_f32<T>(x0[row * c0 + ci]) : to_f32<T>(x1[row * c1 + (ci - c0)]);
        }
        f[i] = v;
      }
    }
  }
};

template <typename T>
struct GatherPlain {  // A(m, k) = x[m][k]
  const T* x; int C; Geom g;
  __device__ __forceinline__ int K() const { return C; }
  template <bool VEC>
  __device__ __forceinline__ void load(int n, int d, int h, int w, bool row_ok, int k8, float (&f)[8]) const {
    const int64_t row = (((int64_t)n * g.D + d) * g.H + h) * g.W + w;
    if (VEC) {
      if (row_ok && k8 < C) { Vec8<T> v; v.load(x + row * C + k8); v.get(f); }
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = (row_ok && k8 + i < C) ? to_f32<T>(x[row * C + k8 + i]) : 0.f;
    }
  }
};

template <typename T>
struct GatherChild {  // A(m, k = child*C + c) = y[child_row(m, child)][c]; y is the 2x upsampled grid
  const T* y; int C; Geom g;  // g = coarse geometry
  __device__ __forceinline__ int K() const { return 8 * C; }
  template <bool VEC>
  __device__ __forceinline__ void load(int n, int d, int h, int w, bool row_ok, int k8, float (&f)[8]) const {
    const int child = k8 / C, c = k8 - child * C;  // VEC: C % 8 == 0 so the 8 k's share a child
    if (VEC) {
      if (row_ok && child < 8) {
        const int64_t row = (((int64_t)n * 2 * g.D + 2 * d + (child >> 2)) * 2 * g.H + 2 * h + ((child >> 1) & 1)) * 2 * g.W + 2 * w + (child & 1);
        Vec8<T> v; v.load(y + row * C + c); v.get(f);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k8 + i, ch = k / C, cc = k - ch * C;
        float v = 0.f;
        if (row_ok && ch < 8) {
          const int64_t row = (((int64_t)n * 2 * g.D + 2 * d + (ch >> 2)) * 2 * g.H + 2 * h + ((ch >> 1) & 1)) * 2 * g.W + 2 * w + (ch & 1);
          v = to_f32<T>(y[row * C + cc]);
        }
        f[i] = v;
      }
    }
  }
};

// ---------------------------------------------------------------- B operands for rows_gemm
template <typename T>
struct WeightDense {  // B[k][n] row-major, leading dimension ld
  const T* w; int ld;
  __device__ __forceinline__ float at(int k, int n) const { return to_f32<T>(w[(int64_t)k * ld + n]); }
};
template <typename T>
struct WeightConvTFwd {  // torch ConvTranspose3d weight [Cin][Cout][8] fp32; B[k=ci][n=child*Cout+co]
  const float* w; int Cin, Cout;
  __device__ __forceinline__ float at(int k, int n) const {
    const int child = n / Cout, co = n - child * Cout;
    return to_f32<T>(from_f32<T>(w[((int64_t)k * Cout + co) * 8 + child]));
  }
};
template <typename T>
struct WeightConvTBwd {  // B[k=child*Cout+co][n=ci]
  const float* w; int Cin, Cout;
  __device__ __forceinline__ float at(int k, int n) const {
    const int child = k / Cout, co = k - child * Cout;
    return to_f32<T>(from_f32<T>(w[((int64_t)n * Cout + co) * 8 + child]));
  }
};

// ---------------------------------------------------------------- epilogues for rows_gemm
template <typename T>
struct StoreRows {  // y0 gets channels [0, co0), y1 gets [co0, co0+co1)
  T* y0; T* y1; int co0, co1; const float* bias; Geom g;
  __device__ __forceinline__ int Ntot() const { return co0 + co1; }
  template <bool VEC>
  __device__ __forceinline__ void store(int64_t m, int n8, const float (&f)[8]) const {
    if (VEC) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = f[i] + (bias ? bias[n8 + i] : 0.f);
      Vec8<T> v; v.set(o);
      if (n8 < co0) v.store(y0 + m * co0 + n8); else v.store(y1 + m * co1 + (n8 - co0));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = n8 + i;
        if (n < co0 + co1) {
          const float o = f[i] + (bias ? bias[n] : 0.f);
          if (n < co0) y0[m * co0 + n] = from_f32<T>(o); else y1[m * co1 + (n - co0)] = from_f32<T>(o);
        }
      }
    }
  }
};
template <typename T>
struct StoreScatter8 {  // n = child*Cout + co -> y[child_row(m, child)][co] (+bias[co]); g = coarse geometry
  T* y; int Cout; const float* bias; Geom g;
  __device__ __forceinline__ int Ntot() const { return 8 * Cout; }
  template <bool VEC>
  __device__ __forceinline__ void store(int64_t m, int n8, const float (&f)[8]) const {
    int n, d, h, w;
    decode_row(g, m, n, d, h, w);
    if (VEC) {
      const int child = n8 / Cout, co = n8 - child * Cout;
      const int64_t row = (((int64_t)n * 2 * g.D + 2 * d + (child >> 2)) * 2 * g.H + 2 * h + ((child >> 1) & 1)) * 2 * g.W + 2 * w + (child & 1);
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = f[i] + (bias ? bias[co + i] : 0.f);
      Vec8<T> v; v.set(o); v.store(y + row * Cout + co);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int nn = n8 + i;
        if (nn < 8 * Cout) {
          const int child = nn / Cout, co = nn - child * Cout;
          const int64_t row = (((int64_t)n * 2 * g.D + 2 * d + (child >> 2)) * 2 * g.H + 2 * h + ((child >> 1) & 1)) * 2 * g.W + 2 * w + (child & 1);
          y[row * Cout + co] = from_f32<T>(f[i] + (bias ? bias[co] : 0.f));
        }
      }
    }
  }
};

// ---------------------------------------------------------------- rows_gemm
// grid = (ceil(M/128), ceil(Ntot/BN)); 256 threads as 16 (n) x 16 (m); thread tile 8 x BN/16.
template <int BN, bool AVEC, bool OVEC, class Gather, class Weight, class Store>
__global__ void __launch_bounds__(kThreads)
rows_gemm_kernel(Gather ga, Weight wt, Store st, int64_t M) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float smem[(BM * (BN + 1) > BK * BM + BK * BN) ? BM * (BN + 1) : (BK * BM + BK * BN)];
  float* As = smem;            // [BK][BM]
  float* Bs = smem + BK * BM;  // [BK][BN]
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = ga.K();
  const int Ntot = st.Ntot();

  // A staging: thread -> row (tid % 128), k half (tid / 128)
  const int am = tid % BM, akq = tid / BM;
  const int64_t arow = m0 + am;
  const bool arow_ok = arow < M;
  int an = 0, ad = 0, ah = 0, aw = 0;
  if (arow_ok) decode_row(ga.g, arow, an, ad, ah, aw);

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float areg[8];
  float breg[(BK * BN + kThreads - 1) / kThreads];
  auto fetch = [&](int k0) {
    ga.template load<AVEC>(an, ad, ah, aw, arow_ok, k0 + akq * 8, areg);
#pragma unroll
    for (int i = 0; i < (BK * BN + kThreads - 1) / kThreads; ++i) {
      const int idx = tid + i * kThreads;
      const int kk = idx / BN, nn = idx % BN;
      breg[i] = (idx < BK * BN && k0 + kk < K && n0 + nn < Ntot) ? wt.at(k0 + kk, n0 + nn) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[(akq * 8 + i) * BM + am] = areg[i];
#pragma unroll
    for (int i = 0; i < (BK * BN + kThreads - 1) / kThreads; ++i) {
      const int idx = tid + i * kThreads;
      if (idx < BK * BN) Bs[idx] = breg[i];
    }
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(As + kk * BM + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(As + kk * BM + ty * 8 + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk * BN + tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // stage C through shared memory so that global stores are whole channel vectors
  float* Cs = smem;  // [BM][BN+1]
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) Cs[(ty * 8 + i) * (BN + 1) + tx * TN + j] = acc[i][j];
  __syncthreads();
  for (int item = tid; item < BM * (BN / 8); item += kThreads) {
    const int r = item / (BN / 8), cv = item % (BN / 8);
    const int64_t m = m0 + r;
    if (m < M && n0 + cv * 8 < Ntot) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = Cs[r * (BN + 1) + cv * 8 + i];
      st.template store<OVEC>(m, n0 + cv * 8, f);
    }
  }
}

template <bool AVEC, bool OVEC, class Gather, class Weight, class Store>
int launch_rows_gemm(const char* name, Gather ga, Weight wt, Store st, int64_t M, int Ntot, cudaStream_t stream) {
  B200_REQUIRE(M > 0 && Ntot > 0, B200_ERR_SHAPE, "%s: empty problem", name);
  const int64_t gx = (M + BM - 1) / BM;
  B200_REQUIRE(gx < 2147483647LL, B200_ERR_UNSUPPORTED, "%s: too many rows", name);
  if (Ntot <= 16) {
    rows_gemm_kernel<16, AVEC, OVEC><<<dim3((unsigned)gx, 1), kThreads, 0, stream>>>(ga, wt, st, M);
  } else if (Ntot <= 32) {
    rows_gemm_kernel<32, AVEC, OVEC><<<dim3((unsigned)gx, 1), kThreads, 0, stream>>>(ga, wt, st, M);
  } else {
    rows_gemm_kernel<64, AVEC, OVEC><<<dim3((unsigned)gx, (unsigned)((Ntot + 63) / 64)), kThreads, 0, stream>>>(ga, wt, st, M);
  }
  B200_CHECK_LAUNCH(name);
  return B200_OK;
}

// ---------------------------------------------------------------- cols_gemm (weight gradients)
// C[r, c] = sum_v A(v, r) B(v, c) over this CTA's voxel range; partial[z][r][c].
template <typename T>
struct ColsK3 {  // A(v, r = tap*Cin + ci) = [x0|x1][v + tap][ci]
  const T* x0; const T* x1; int c0, c1; Geom g;
  __device__ __forceinline__ int R() const { return 27 * (c0 + c1); }
  __device__ __forceinline__ float at(int n, int d, int h, int w, int r) const {
    const int Cin = c0 + c1;
    const int tap = r / Cin, ci = r - tap * Cin;
    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    const int dd = d + kd - 1, hh = h + kh - 1, ww = w + kw - 1;
    if ((unsigned)dd < (unsigned)g.D && (unsigned)hh < (unsigned)g.H && (unsigned)ww < (unsigned)g.W) {
      const int64_t row = (((int64_t)n * g.D + dd) * g.H + hh) * g.W + ww;
      return ci < c0 ? to_f32<T>(x0[row * c0 + ci]) : to_f32<T>(x1[row * c1 + (ci - c0)]);
    }
    return 0.f;
  }
};
template <typename T>
struct ColsPlain {  // A(v, r) = x[v][r]
  const T* x; int C; Geom g;
  __device__ __forceinline__ int R() const { return C; }
  __device__ __forceinline__ float at(int n, int d, int h, int w, int r) const {
    const int64_t row = (((int64_t)n * g.D + d) * g.H + h) * g.W + w;
    return to_f32<T>(x[row * C + r]);
  }
};
template <typename T>
struct ColsDy {  // B(v, c) = dy[v][c]
  const T* dy; int C; Geom g;
  __device__ __forceinline__ int Cn() const { return C; }
  __device__ __forceinline__ float at(int n, int d, int h, int w, int c) const {
    const int64_t row = (((int64_t)n * g.D + d) * g.H + h) * g.W + w;
    return to_f32<T>(dy[row * C + c]);
  }
};
template <typename T>
struct ColsChild {  // B(v, c = child*Cout + co) = gy[child_row(v, child)][co]; g = coarse geometry
  const T* gy; int Cout; Geom g;
  __device__ __forceinline__ int Cn() const { return 8 * Cout; }
  __device__ __forceinline__ float at(int n, int d, int h, int w, int c) const {
    const int child = c / Cout, co = c - child * Cout;
    const int64_t row = (((int64_t)n * 2 * g.D + 2 * d + (child >> 2)) * 2 * g.H + 2 * h + ((child >> 1) & 1)) * 2 * g.W + 2 * w + (child & 1);
    return to_f32<T>(gy[row * Cout + co]);
  }
};

template <int BN, class ColsA, class ColsB>
__global__ void __launch_bounds__(kThreads)
cols_gemm_kernel(ColsA ca, ColsB cb, int64_t M, int64_t chunk, float* __restrict__ partial) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK * BM];
  __shared__ __align__(16) float Bs[BK * BN];
  __shared__ int coords[BK][4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int r0 = blockIdx.x * BM, c0 = blockIdx.y * BN;
  const int R = ca.R(), Cn = cb.Cn();
  const int64_t v_begin = (int64_t)blockIdx.z * chunk;
  const int64_t v_end = (v_begin + chunk < M) ? v_begin + chunk : M;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int ar = tid % BM, avq = tid / BM;  // A staging: row r0+ar, voxels avq*8..+8
  for (int64_t v0 = v_begin; v0 < v_end; v0 += BK) {
    if (tid < BK) {
      const int64_t v = v0 + tid;
      int n = -1, d = 0, h = 0, w = 0;
      if (v < v_end) decode_row(ca.g, v, n, d, h, w);
      coords[tid][0] = n; coords[tid][1] = d; coords[tid][2] = h; coords[tid][3] = w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int vv = avq * 8 + i;
      const int n = coords[vv][0];
      As[vv * BM + ar] = (n >= 0 && r0 + ar < R) ? ca.at(n, coords[vv][1], coords[vv][2], coords[vv][3], r0 + ar) : 0.f;
    }
    for (int idx = tid; idx < BK * BN; idx += kThreads) {
      const int vv = idx / BN, cc = idx % BN;
      const int n = coords[vv][0];
      Bs[idx] = (n >= 0 && c0 + cc < Cn) ? cb.at(n, coords[vv][1], coords[vv][2], coords[vv][3], c0 + cc) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(As + kk * BM + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(As + kk * BM + ty * 8 + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk * BN + tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + (int64_t)blockIdx.z * R * Cn;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty * 8 + i;
    if (r < R) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int c = c0 + tx * TN + j;
        if (c < Cn) out[(int64_t)r * Cn + c] = acc[i][j];
      }
    }
  }
}

struct SplitPlan { int nsplit; int64_t chunk; };
inline SplitPlan plan_split(int64_t M, int R, int Cn) {
  const int bn = Cn <= 16 ? 16 : (Cn <= 32 ? 32 : 64);
  const int64_t tiles = (int64_t)((R + BM - 1) / BM) * ((Cn + bn - 1) / bn);
  int64_t want = (4LL * B200_NUM_SMS + tiles - 1) / tiles;             // ~4 waves of CTAs
  const int64_t max_by_mem = (64LL << 20) / ((int64_t)R * Cn * 4) + 1;  // <= 64 MiB of partials
  if (want > max_by_mem) want = max_by_mem;
  int64_t chunk = (M + want - 1) / want;
  chunk = ((chunk + BK - 1) / BK) * BK;
  if (chunk < 64) chunk = 64;
  SplitPlan p;
  p.chunk = chunk;
  p.nsplit = (int)((M + chunk - 1) / chunk);
  return p;
}

template <class ColsA, class ColsB>
int launch_cols_gemm(const char* name, ColsA ca, ColsB cb, int64_t M, int R, int Cn, SplitPlan sp, float* partial, cudaStream_t stream) {
  const int gy16 = 1;
  (void)gy16;
  if (Cn <= 16) {
    cols_gemm_kernel<16><<<dim3((R + BM - 1) / BM, 1, sp.nsplit), kThreads, 0, stream>>>(ca, cb, M, sp.chunk, partial);
  } else if (Cn <= 32) {
    cols_gemm_kernel<32><<<dim3((R + BM - 1) / BM, 1, sp.nsplit), kThreads, 0, stream>>>(ca, cb, M, sp.chunk, partial);
  } else {
    cols_gemm_kernel<64><<<dim3((R + BM - 1) / BM, (Cn + 63) / 64, sp.nsplit), kThreads, 0, stream>>>(ca, cb, M, sp.chunk, partial);
  }
  B200_CHECK_LAUNCH(name);
  return B200_OK;
}

// partial[z][r = tap*Cin+ci][co] -> dw[co][ci][tap]   (torch Conv3d weight layout)
__global__ void reduce_k3_wgrad_kernel(const float* __restrict__ partial, int nsplit, int Cin, int Cout, float* __restrict__ dw) {
  const int64_t total = (int64_t)27 * Cin * Cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes the output so that writes are coalesced
    const int tap = (int)(i % 27);
    const int ci = (int)((i / 27) % Cin);
    const int co = (int)(i / (27 * (int64_t)Cin));
    const int64_t src = ((int64_t)tap * Cin + ci) * Cout + co;
    double s = 0.0;
    for (int z = 0; z < nsplit; ++z) s += (double)partial[(int64_t)z * total + src];
    dw[i] = (float)s;
  }
}
// partial[z][ci][child*Cout+co] -> dw[ci][co][child]  (torch ConvTranspose3d weight layout)
__global__ void reduce_convt_wgrad_kernel(const float* __restrict__ partial, int nsplit, int Cin, int Cout, float* __restrict__ dw) {
  const int64_t total = (int64_t)8 * Cin * Cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int child = (int)(i % 8);
    const int co = (int)((i / 8) % Cout);
    const int ci = (int)(i / (8 * (int64_t)Cout));
    const int64_t src = (int64_t)ci * 8 * Cout + (int64_t)child * Cout + co;
    double s = 0.0;
    for (int z = 0; z < nsplit; ++z) s += (double)partial[(int64_t)z * total + src];
    dw[i] = (float)s;
  }
}

// ---------------------------------------------------------------- weight packing (direct layouts)
template <typename T>
__global__ void pack_k3_kernel(const float* __restrict__ w, T* __restrict__ out, int Cout, int Cin, int dgrad) {
  const int64_t total = (int64_t)27 * Cin * Cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (!dgrad) {  // out[tap][ci][co] = w[co][ci][tap]
      const int co = (int)(i % Cout);
      const int ci = (int)((i / Cout) % Cin);
      const int tap = (int)(i / ((int64_t)Cout * Cin));
      out[i] = from_f32<T>(w[((int64_t)co * Cin + ci) * 27 + tap]);
    } else {  // out[tap'][co][ci] = w[co][ci][26 - tap']
      const int ci = (int)(i % Cin);
      const int co = (int)((i / Cin) % Cout);
      const int tapf = (int)(i / ((int64_t)Cout * Cin));
      out[i] = from_f32<T>(w[((int64_t)co * Cin + ci) * 27 + (26 - tapf)]);
    }
  }
}

}  // namespace

// tcgen05 path (conv_tc.cu)
int b200_conv3d_k3_tc(const void* x0, int c0, const void* x1, int c1, const void* wpack, const float* bias, void* y0,
                      int co0, void* y1, int co1, int N, int D, int H, int W, cudaStream_t stream);
int b200_pack_conv3_weights_tc(int mode, const float* w, void* out, int Cout, int Cin, cudaStream_t stream);
int64_t b200_pack_conv3_bytes_tc(int Cout, int Cin);
bool b200_conv3d_k3_tc_supported(int c0, int c1, int co0, int co1, int N, int D, int H, int W);
// second-generation tcgen05 path (conv_tc2.cu): TMA-fed stages, kd-fused MMAs
int b200_conv3d_k3_tc2(const void* x0, int c0, const void* x1, int c1, const void* wpack, const float* bias, void* y0,
                       int co0, void* y1, int co1, int N, int D, int H, int W, cudaStream_t stream);
int b200_pack_conv3_weights_tc2(int mode, const float* w, void* out, int Cout, int Cin, cudaStream_t stream);
// persistent variant for layers with many tiles (conv_tc3.cu); same packed weights as tc2
bool b200_conv3d_k3_tc3_wanted(int c0, int c1, int co0, int co1, int N, int D, int H, int W);
int b200_conv3d_k3_tc3(const void* x0, int c0, const void* x1, int c1, const void* wpack, const float* bias, void* y0,
                       int co0, void* y1, int co1, int N, int D, int H, int W, cudaStream_t stream, float* stats = nullptr);
int b200_conv3d_k3_tc3_stats_blocks(int c0, int c1, int co0, int co1, int N, int D, int H, int W);
// row-streaming variant for full-resolution layers with 16 / 32 output channels (conv_tc4.cu); same packed weights
bool b200_conv3d_k3_tc4_wanted(int c0, int c1, int co0, int co1, int N, int D, int H, int W);
int b200_conv3d_k3_tc4_stats_blocks(int c0, int c1, int co0, int co1, int N, int D, int H, int W);
int b200_conv3d_k3_tc4(const void* x0, int c0, const void* x1, int c1, const void* wpack, const float* bias, void* y0,
                       int co0, void* y1, int co1, int N, int D, int H, int W, cudaStream_t stream, float* stats = nullptr,
                       const void* xprev = nullptr, const float* bn_scale = nullptr, const float* bn_shift = nullptr,
                       const float* bn_mean = nullptr, const float* bn_invstd = nullptr);
void b200_conv3d_k3_tc4_enable(int on);
static bool conv_persistent_on(int c0, int c1, int co0, int co1);
static int g_conv_persistent = -1;  // -1 unset (env B200_CONV_PERSISTENT or 1), 0 never, 1 auto, 2 whenever the layer has enough tiles
static int tc_version() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200_CONV_TC_VERSION"); v = e ? atoi(e) : 2; }
  return v;
}
// tcgen05 weight gradient (wgrad_tc.cu)
bool b200_conv3d_wgrad_tc_supported(int c0, int c1, int Cout, int N, int D, int H, int W);
int64_t b200_conv3d_wgrad_tc_workspace(int c0, int c1, int Cout, int N, int D, int H, int W);
int b200_conv3d_wgrad_tc(const void* x0, int c0, const void* x1, int c1, const void* dy, int Cout, float* dw, void* workspace,
                         int N, int D, int H, int W, cudaStream_t stream);
// second-generation tcgen05 weight gradient (wgrad_tc2.cu): kw taps on the M side, kh taps on the N side
bool b200_conv3d_wgrad_tc2_supported(int c0, int c1, int Cout, int N, int D, int H, int W);
int64_t b200_conv3d_wgrad_tc2_workspace(int c0, int c1, int Cout, int N, int D, int H, int W);
int b200_conv3d_wgrad_tc2(const void* x0, int c0, const void* x1, int c1, const void* dy, int Cout, float* dw, void* workspace,
                          int N, int D, int H, int W, cudaStream_t stream);
static int wg_version() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200_WGRAD_TC_VERSION"); v = e ? atoi(e) : 2; }
  return v;
}
static bool wg_wide_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200_WGRAD_WIDE"); v = e ? atoi(e) : 1; }
  return v != 0;
}
// ConvTranspose3d forward / data gradient on the tensor cores (convt_tc.cu)
bool b200_convt2_tc_supported(int Cin, int Cout);
int b200_convt2_fwd_tc(const void* x, const float* w, const float* bias, void* y, int N, int D, int H, int W, int Cin, int Cout, cudaStream_t st);
int b200_convt2_bwd_data_tc(const void* gy, const float* w, void* gx, int N, int D, int H, int W, int Cin, int Cout, cudaStream_t st);
static int convt_tc_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200_CONVT_TC"); v = e ? atoi(e) : 1; }
  return v;
}
bool b200_convt2_wgrad_tc_supported(int Cin, int Cout, int N, int D, int H, int W);
int64_t b200_convt2_wgrad_tc_workspace(int Cin, int Cout, int N, int D, int H, int W);
int b200_convt2_wgrad_tc(const void* x, const void* gy, float* dw, float* dbias, void* workspace, int N, int D, int H, int W, int Cin, int Cout, cudaStream_t stream);
// wide-row variant (wgrad_tc3.cu): 32-channel operand rows for layers with >= 32 channels on both sides
bool b200_conv3d_wgrad_tc3_supported(int c0, int c1, int Cout, int N, int D, int H, int W);
int64_t b200_conv3d_wgrad_tc3_workspace(int c0, int c1, int Cout, int N, int D, int H, int W);
int b200_conv3d_wgrad_tc3(const void* x0, int c0, const void* x1, int c1, const void* dy, int Cout, float* dw, void* workspace,
                          int N, int D, int H, int W, cudaStream_t stream);
// voxel-pair variant (wgrad_tc4.cu): 16-channel tensors, operand rows = two w-adjacent voxels (B200_WGRAD_PAIR=0 disables it)
bool b200_conv3d_wgrad_tc4_supported(int c0, int c1, int Cout, int N, int D, int H, int W);
int64_t b200_conv3d_wgrad_tc4_workspace(int c0, int c1, int Cout, int N, int D, int H, int W);
int b200_conv3d_wgrad_tc4(const void* x0, int c0, const void* x1, int c1, const void* dy, int Cout, float* dw, void* workspace,
                          int N, int D, int H, int W, cudaStream_t stream);
static int g_wgrad_impl = 0;  // 0 auto, 1 CUDA-core, 2 tcgen05
static int g_inject_wgrad_failure = 0;  // tests: the next wide-row tcgen05 weight gradient reports failure
// in_channels == 1 first layer (conv_stem.cu)
bool b200_conv_stem_supported(int c0, int c1, int cout);
bool b200_conv_stem_wgrad_supported(int c0, int c1, int cout);
int b200_conv_stem_fwd(int dtype, const void* x, const void* wpack, const float* bias, void* y, int cout, int N, int D, int H, int W, cudaStream_t st,
                       float* stats = nullptr);
int b200_conv_stem_stats_blocks(int dtype, int cout, int N, int D, int H, int W);
int64_t b200_conv_stem_wgrad_workspace(int cout);
int b200_conv_stem_wgrad(int dtype, const void* x, const void* dy, int cout, float* dw, float* partials, int N, int D, int H, int W, cudaStream_t st);

// =========================================================================== exports
extern "C" int64_t b200_pack_conv3_bytes(int mode, int dtype, int Cout, int Cin) {
  if (mode == B200_PACK_FPROP_TC) return b200_pack_conv3_bytes_tc(Cout, Cin);
  if (mode == B200_PACK_DGRAD_TC) return b200_pack_conv3_bytes_tc(Cin, Cout);
  return (int64_t)27 * Cin * Cout * (dtype == B200_F32 ? 4 : 2);
}

extern "C" int b200_pack_conv3_weights(int mode, int dtype, const float* w, void* out, int Cout, int Cin, void* stream) {
  B200_REQUIRE(w && out && Cout > 0 && Cin > 0, B200_ERR_SHAPE, "pack_conv3_weights: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == B200_PACK_FPROP_TC || mode == B200_PACK_DGRAD_TC) {
    B200_REQUIRE(dtype == B200_BF16, B200_ERR_UNSUPPORTED, "pack_conv3_weights: tcgen05 layouts are bf16 only");
    if (tc_version() == 2) return b200_pack_conv3_weights_tc2(mode, w, out, Cout, Cin, st);
    return b200_pack_conv3_weights_tc(mode, w, out, Cout, Cin, st);
  }
  B200_REQUIRE(mode == B200_PACK_FPROP || mode == B200_PACK_DGRAD, B200_ERR_UNSUPPORTED, "pack_conv3_weights: mode %d", mode);
  const int64_t total = (int64_t)27 * Cin * Cout;
  const int grid = b200_grid_for(total, 256, B200_NUM_SMS * 8);
  B200_DISPATCH_DTYPE(dtype, T, (pack_k3_kernel<T><<<grid, 256, 0, st>>>(w, (T*)out, Cout, Cin, mode == B200_PACK_DGRAD)));
  B200_CHECK_LAUNCH("pack_conv3_weights");
  return B200_OK;
}

int b200_pack_conv3_batched_tc2(const void* jobs, int njobs, long long total_groups, cudaStream_t stream);

// All tcgen05 weight layouts of a model in one launch: `jobs` = device array of njobs + 1 records
// {const float* w; void* out; int32 Cout, Cin, dgrad, pad; int64 group_begin} (40 bytes each; group_begin counts 16-byte output
// groups = 27 * Cin * Cout / 8 per job, record njobs carries the total).  Destinations are what b200_pack_conv3_weights would write.
extern "C" int b200_pack_conv3_batched(const void* jobs, int njobs, int64_t total_groups, void* stream) {
  B200_REQUIRE(tc_version() == 2, B200_ERR_UNSUPPORTED, "pack_conv3_batched: only the current tcgen05 weight layout is supported");
  return b200_pack_conv3_batched_tc2(jobs, njobs, total_groups, (cudaStream_t)stream);
}

extern "C" int b200_conv3d_k3_select(int dtype, int impl, int c0, int c1, int co0, int co1, int N, int D, int H, int W) {
  if (impl == 1 || impl == 2) return impl;
  return (dtype == B200_BF16 && b200_conv3d_k3_tc_supported(c0, c1, co0, co1, N, D, H, W)) ? 2 : 1;
}

extern "C" int b200_conv3d_k3(int dtype, int impl, const void* x0, int c0, const void* x1, int c1, const void* wpack,
                              const float* bias, void* y0, int co0, void* y1, int co1, int N, int D, int H, int W,
                              void* stream) {
  B200_REQUIRE(x0 && wpack && y0, B200_ERR_SHAPE, "conv3d_k3: null pointer");
  B200_REQUIRE(c0 > 0 && c1 >= 0 && co0 > 0 && co1 >= 0, B200_ERR_SHAPE, "conv3d_k3: bad channel counts");
  B200_REQUIRE((c1 == 0) == (x1 == nullptr) && (co1 == 0) == (y1 == nullptr), B200_ERR_SHAPE, "conv3d_k3: second tensor / channel count mismatch");
  B200_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0, B200_ERR_SHAPE, "conv3d_k3: empty volume");
  cudaStream_t st = (cudaStream_t)stream;
  B200_REQUIRE(impl == 1 || impl == 2, B200_ERR_UNSUPPORTED,
               "conv3d_k3: impl must be 1 (CUDA-core) or 2 (tcgen05); resolve 0 with b200_conv3d_k3_select so that the weights are packed to match");
  if (impl == 2) {
    B200_REQUIRE(dtype == B200_BF16, B200_ERR_UNSUPPORTED, "conv3d_k3: tcgen05 path is bf16 only");
    B200_REQUIRE(b200_conv3d_k3_tc_supported(c0, c1, co0, co1, N, D, H, W), B200_ERR_UNSUPPORTED,
                 "conv3d_k3(tcgen05): channels (%d+%d)->(%d+%d) need multiples of 16 (Cout <= 128 or a multiple of 128)", c0, c1, co0, co1);
    if (tc_version() == 2) {
      // The persistent kernel (conv_tc3.cu: two MMA-issuing warps, double-buffered TMEM) serves every layer with >= 2 tiles per
      // SM; the two-CTA-per-SM kernel (conv_tc2.cu, split-K) the deep layers.  Mode 3 = the old 16->16-only policy.
      if (conv_persistent_on(c0, c1, co0, co1) && b200_conv3d_k3_tc4_wanted(c0, c1, co0, co1, N, D, H, W))
        return b200_conv3d_k3_tc4(x0, c0, x1, c1, wpack, bias, y0, co0, y1, co1, N, D, H, W, st);
      if (conv_persistent_on(c0, c1, co0, co1) && b200_conv3d_k3_tc3_wanted(c0, c1, co0, co1, N, D, H, W))
        return b200_conv3d_k3_tc3(x0, c0, x1, c1, wpack, bias, y0, co0, y1, co1, N, D, H, W, st);
      return b200_conv3d_k3_tc2(x0, c0, x1, c1, wpack, bias, y0, co0, y1, co1, N, D, H, W, st);
    }
    return b200_conv3d_k3_tc(x0, c0, x1, c1, wpack, bias, y0, co0, y1, co1, N, D, H, W, st);
  }
  if (co1 == 0 && b200_conv_stem_supported(c0, c1, co0) && (dtype == B200_F32 || dtype == B200_BF16))
    return b200_conv_stem_fwd(dtype, x0, wpack, bias, y0, co0, N, D, H, W, st);
  const Geom g{N, D, H, W};
  const int64_t M = g.rows();
  const int Cin = c0 + c1, Cout = co0 + co1;
  const bool avec = (c0 % 8 == 0) && (c1 % 8 == 0);
  const bool ovec = (co0 % 8 == 0) && (co1 % 8 == 0);
  (void)Cin;
#define RUN(T)                                                                                                   \
  do {                                                                                                           \
    GatherK3<T> ga{(const T*)x0, (const T*)x1, c0, c1, g};                                                       \
    WeightDense<T> wt{(const T*)wpack, Cout};                                                                    \
    StoreRows<T> so{(T*)y0, (T*)y1, co0, co1, bias, g};                                                          \
    if (avec && ovec) return launch_rows_gemm<true, true>("conv3d_k3", ga, wt, so, M, Cout, st);                 \
    if (avec) return launch_rows_gemm<true, false>("conv3d_k3", ga, wt, so, M, Cout, st);                        \
    if (ovec) return launch_rows_gemm<false, true>("conv3d_k3", ga, wt, so, M, Cout, st);                        \
    return launch_rows_gemm<false, false>("conv3d_k3", ga, wt, so, M, Cout, st);                                 \
  } while (0)
  if (dtype == B200_F32) RUN(float);
  if (dtype == B200_BF16) RUN(__nv_bfloat16);
#undef RUN
  B200_FAIL(B200_ERR_UNSUPPORTED, "conv3d_k3: unknown dtype %d", dtype);
}

static bool conv_persistent_on(int c0, int c1, int co0, int co1) {
  if (g_conv_persistent < 0) { const char* e = getenv("B200_CONV_PERSISTENT"); g_conv_persistent = e ? atoi(e) : 1; }
  return g_conv_persistent == 1 || g_conv_persistent == 2 || (g_conv_persistent == 3 && c0 + c1 == 16 && co0 + co1 == 16);
}

// Convolution + BatchNorm batch statistics in one kernel (models/unet.py:11-12 / :15-16): the persistent tcgen05 kernel's epilogue
// accumulates sum / sum of squares of (y - bias) per channel over the bf16 values it stores.  Returns the number of partial rows
// written to `partials` ([rows][2][Cout] fp32, the layout of b200_bn_stats) — pass it, with shift = bias, to b200_bn_finalize_ex.
extern "C" int b200_conv3d_k3_bnstats_blocks(int dtype, int impl, int c0, int c1, int co0, int co1, int N, int D, int H, int W) {
  if (co1 == 0 && b200_conv_stem_supported(c0, c1, co0)) return b200_conv_stem_stats_blocks(dtype, co0, N, D, H, W);   // first layer (Cin = 1)
  if (dtype != B200_BF16 || impl != 2 || tc_version() != 2 || !b200_conv3d_k3_tc_supported(c0, c1, co0, co1, N, D, H, W)) return 0;
  if (!conv_persistent_on(c0, c1, co0, co1)) return 0;
  const int rows = b200_conv3d_k3_tc4_stats_blocks(c0, c1, co0, co1, N, D, H, W);
  return rows > 0 ? rows : b200_conv3d_k3_tc3_stats_blocks(c0, c1, co0, co1, N, D, H, W);
}

extern "C" int b200_conv3d_k3_bnstats(int dtype, int impl, const void* x0, int c0, const void* x1, int c1, const void* wpack,
                                      const float* bias, void* y0, int co0, int N, int D, int H, int W, float* partials, void* stream) {
  B200_REQUIRE(x0 && wpack && y0 && partials, B200_ERR_SHAPE, "conv3d_k3_bnstats: null pointer");
  B200_REQUIRE((c1 == 0) == (x1 == nullptr), B200_ERR_SHAPE, "conv3d_k3_bnstats: second tensor / channel count mismatch");
  B200_REQUIRE(b200_conv3d_k3_bnstats_blocks(dtype, impl, c0, c1, co0, 0, N, D, H, W) > 0, B200_ERR_UNSUPPORTED,
               "conv3d_k3_bnstats: the fused-statistics kernel does not serve this problem (ask b200_conv3d_k3_bnstats_blocks first)");
  if (b200_conv_stem_supported(c0, c1, co0))
    return b200_conv_stem_fwd(dtype, x0, wpack, bias, y0, co0, N, D, H, W, (cudaStream_t)stream, partials);
  if (b200_conv3d_k3_tc4_stats_blocks(c0, c1, co0, 0, N, D, H, W) > 0)
    return b200_conv3d_k3_tc4(x0, c0, x1, c1, wpack, bias, y0, co0, nullptr, 0, N, D, H, W, (cudaStream_t)stream, partials);
  return b200_conv3d_k3_tc3(x0, c0, x1, c1, wpack, bias, y0, co0, nullptr, 0, N, D, H, W, (cudaStream_t)stream, partials);
}

// Data gradient + BatchNorm-backward reduction in one kernel: y0 = gy (the gradient w.r.t. the previous layer's relu(bn(xprev))) and
// partials[rows][2][co0] = per-CTA (sum g, invstd * sum g * (xprev - mean)), g = gy * [bn(xprev) > 0] — b200_bn_act_bwd_reduce's
// output for that layer; rows = b200_conv3d_k3_bnbwd_blocks() (0: not available, run the separate pass).
extern "C" int b200_conv3d_k3_bnbwd_blocks(int dtype, int impl, int c0, int co0, int N, int D, int H, int W) {
  if (dtype != B200_BF16 || impl != 2 || tc_version() != 2 || !b200_conv3d_k3_tc_supported(c0, 0, co0, 0, N, D, H, W)) return 0;
  if (!conv_persistent_on(c0, 0, co0, 0)) return 0;
  return b200_conv3d_k3_tc4_stats_blocks(c0, 0, co0, 0, N, D, H, W);
}

extern "C" int b200_conv3d_k3_bnbwd(int dtype, int impl, const void* x0, int c0, const void* wpack, void* y0, int co0, int N, int D, int H,
                                    int W, const void* xprev, const float* scale, const float* shift, const float* mean, const float* invstd,
                                    float* partials, void* stream) {
  B200_REQUIRE(x0 && wpack && y0 && xprev && partials, B200_ERR_SHAPE, "conv3d_k3_bnbwd: null pointer");
  B200_REQUIRE(b200_conv3d_k3_bnbwd_blocks(dtype, impl, c0, co0, N, D, H, W) > 0, B200_ERR_UNSUPPORTED,
               "conv3d_k3_bnbwd: no fused kernel serves this problem (ask b200_conv3d_k3_bnbwd_blocks first)");
  return b200_conv3d_k3_tc4(x0, c0, nullptr, 0, wpack, nullptr, y0, co0, nullptr, 0, N, D, H, W, (cudaStream_t)stream, partials, xprev, scale,
                            shift, mean, invstd);
}

/* 0 = never use the row-streaming kernel (conv_tc4.cu), 1 = wherever it applies (default) — tests and A/B timing */
extern "C" int b200_set_conv_rowstream(int on) {
  b200_conv3d_k3_tc4_enable(on);
  return B200_OK;
}

extern "C" int b200_set_conv_persistent(int mode) {
  B200_REQUIRE(mode >= 0 && mode <= 3, B200_ERR_UNSUPPORTED, "set_conv_persistent: mode must be 0 (never), 1 (auto), 2 (whenever possible) or 3 (16->16 only)");
  g_conv_persistent = mode;
  return B200_OK;
}

extern "C" int b200_debug_fail_next_wgrad(int on) {
  g_inject_wgrad_failure = on ? 1 : 0;
  return B200_OK;
}

extern "C" int b200_set_wgrad_impl(int impl) {
  B200_REQUIRE(impl >= 0 && impl <= 2, B200_ERR_UNSUPPORTED, "set_wgrad_impl: impl must be 0 (auto), 1 (CUDA-core) or 2 (tcgen05)");
  g_wgrad_impl = impl;
  return B200_OK;
}

// workspace layout: [bias partials][split-K partials of whichever kernel runs]
extern "C" int64_t b200_conv3d_wgrad_workspace(int c0, int c1, int Cout, int N, int D, int H, int W) {
  const int R = 27 * (c0 + c1);
  const SplitPlan sp = plan_split((int64_t)N * D * H * W, R, Cout);
  int64_t main_bytes = (int64_t)sp.nsplit * R * Cout * 4;
  if (b200_conv3d_wgrad_tc_supported(c0, c1, Cout, N, D, H, W)) {
    const int64_t t = b200_conv3d_wgrad_tc_workspace(c0, c1, Cout, N, D, H, W);
    if (t > main_bytes) main_bytes = t;
  }
  if (b200_conv3d_wgrad_tc2_supported(c0, c1, Cout, N, D, H, W)) {
    const int64_t t = b200_conv3d_wgrad_tc2_workspace(c0, c1, Cout, N, D, H, W);
    if (t > main_bytes) main_bytes = t;
  }
  if (b200_conv3d_wgrad_tc3_supported(c0, c1, Cout, N, D, H, W)) {
    const int64_t t = b200_conv3d_wgrad_tc3_workspace(c0, c1, Cout, N, D, H, W);
    if (t > main_bytes) main_bytes = t;
  }
  if (b200_conv3d_wgrad_tc4_supported(c0, c1, Cout, N, D, H, W)) {
    const int64_t t = b200_conv3d_wgrad_tc4_workspace(c0, c1, Cout, N, D, H, W);
    if (t > main_bytes) main_bytes = t;
  }
  if (b200_conv_stem_wgrad_supported(c0, c1, Cout) && b200_conv_stem_wgrad_workspace(Cout) > main_bytes)
    main_bytes = b200_conv_stem_wgrad_workspace(Cout);
  return b200_bn_partials_bytes(((Cout + 7) / 8) * 8) + main_bytes;
}

extern "C" int b200_conv3d_wgrad(int dtype, const void* x0, int c0, const void* x1, int c1, const void* dy, int Cout,
                                 float* dw, float* dbias, void* workspace, int64_t workspace_bytes, int N, int D, int H,
                                 int W, void* stream) {
  B200_REQUIRE(x0 && dy && dw && workspace, B200_ERR_SHAPE, "conv3d_wgrad: null pointer");
  B200_REQUIRE((c1 == 0) == (x1 == nullptr), B200_ERR_SHAPE, "conv3d_wgrad: second tensor / channel count mismatch");
  B200_REQUIRE(workspace_bytes >= b200_conv3d_wgrad_workspace(c0, c1, Cout, N, D, H, W), B200_ERR_SHAPE, "conv3d_wgrad: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const Geom g{N, D, H, W};
  const int64_t M = g.rows();
  const int Cin = c0 + c1, R = 27 * Cin;
  float* bpart = (float*)workspace;
  float* partial = (float*)((uint8_t*)workspace + b200_bn_partials_bytes(((Cout + 7) / 8) * 8));
  // v2 (kw on M, kh on N) wins while the channel counts are small; for wide layers v1 fills its 64 M rows with real channels
  const bool use_v3 = dtype == B200_BF16 && wg_version() >= 2 && wg_wide_enabled() && b200_conv3d_wgrad_tc3_supported(c0, c1, Cout, N, D, H, W);
  const bool use_v2 = wg_version() == 2 && (int64_t)(c0 + c1) * Cout <= 4096 && b200_conv3d_wgrad_tc2_supported(c0, c1, Cout, N, D, H, W);
  // 16-channel layers: the voxel-pair kernel (M = 128 x N = 64 instructions, 56 % of the blocks useful) replaces v2 (M = 64 x N = 48)
  const bool use_v4 = dtype == B200_BF16 && use_v2 && !use_v3 && b200_conv3d_wgrad_tc4_supported(c0, c1, Cout, N, D, H, W);
  const bool tc_ok = dtype == B200_BF16 && (use_v2 ? b200_conv3d_wgrad_tc2_supported(c0, c1, Cout, N, D, H, W)
                                                   : b200_conv3d_wgrad_tc_supported(c0, c1, Cout, N, D, H, W));
  B200_REQUIRE(g_wgrad_impl != 2 || tc_ok, B200_ERR_UNSUPPORTED, "conv3d_wgrad: tcgen05 path forced but unsupported for this problem");
  int rc;
  if (b200_conv_stem_wgrad_supported(c0, c1, Cout) && (dtype == B200_F32 || dtype == B200_BF16)) {
    rc = b200_conv_stem_wgrad(dtype, x0, dy, Cout, dw, partial, N, D, H, W, st);
    if (rc) return rc;
  } else if (use_v3 && g_wgrad_impl != 1) {
    rc = g_inject_wgrad_failure ? (g_inject_wgrad_failure = 0, b200_set_error("conv3d_wgrad: injected failure (b200_debug_fail_next_wgrad)"), B200_ERR_UNSUPPORTED)
                                : b200_conv3d_wgrad_tc3(x0, c0, x1, c1, dy, Cout, dw, partial, N, D, H, W, st);
    if (rc) return rc;
  } else if (use_v4 && g_wgrad_impl != 1) {
    rc = b200_conv3d_wgrad_tc4(x0, c0, x1, c1, dy, Cout, dw, partial, N, D, H, W, st);
    if (rc) return rc;
  } else if (tc_ok && g_wgrad_impl != 1) {
    rc = use_v2 ? b200_conv3d_wgrad_tc2(x0, c0, x1, c1, dy, Cout, dw, partial, N, D, H, W, st)
                : b200_conv3d_wgrad_tc(x0, c0, x1, c1, dy, Cout, dw, partial, N, D, H, W, st);
    if (rc) return rc;
  } else {
    const SplitPlan sp = plan_split(M, R, Cout);
    if (dtype == B200_F32) {
      rc = launch_cols_gemm("conv3d_wgrad", ColsK3<float>{(const float*)x0, (const float*)x1, c0, c1, g}, ColsDy<float>{(const float*)dy, Cout, g}, M, R, Cout, sp, partial, st);
    } else if (dtype == B200_BF16) {
      rc = launch_cols_gemm("conv3d_wgrad", ColsK3<__nv_bfloat16>{(const __nv_bfloat16*)x0, (const __nv_bfloat16*)x1, c0, c1, g},
                            ColsDy<__nv_bfloat16>{(const __nv_bfloat16*)dy, Cout, g}, M, R, Cout, sp, partial, st);
    } else {
      B200_FAIL(B200_ERR_UNSUPPORTED, "conv3d_wgrad: unknown dtype %d", dtype);
    }
    if (rc) return rc;
    const int64_t total = (int64_t)R * Cout;
    reduce_k3_wgrad_kernel<<<b200_grid_for(total, 256, B200_NUM_SMS * 8), 256, 0, st>>>(partial, sp.nsplit, Cin, Cout, dw);
    B200_CHECK_LAUNCH("conv3d_wgrad_reduce");
  }
  if (dbias) {
    B200_REQUIRE(Cout % 8 == 0, B200_ERR_UNSUPPORTED, "conv3d_wgrad: dbias needs Cout %% 8 == 0");
    return b200_channel_sum(dtype, dy, M, Cout, bpart, dbias, stream);
  }
  return B200_OK;
}

extern "C" int b200_convt2_fwd(int dtype, const void* x, const float* w, const float* bias, void* y, int N, int D, int H,
                               int W, int Cin, int Cout, void* stream) {
  B200_REQUIRE(x && w && y && N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, B200_ERR_SHAPE, "convt2_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200_BF16 && convt_tc_enabled() && b200_convt2_tc_supported(Cin, Cout))
    return b200_convt2_fwd_tc(x, w, bias, y, N, D, H, W, Cin, Cout, st);
  const Geom g{N, D, H, W};
  const bool avec = Cin % 8 == 0, ovec = Cout % 8 == 0;
#define RUN(T)                                                                                             \
  do {                                                                                                     \
    GatherPlain<T> ga{(const T*)x, Cin, g};                                                                \
    WeightConvTFwd<T> wt{w, Cin, Cout};                                                                    \
    StoreScatter8<T> so{(T*)y, Cout, bias, g};                                                             \
    if (avec && ovec) return launch_rows_gemm<true, true>("convt2_fwd", ga, wt, so, g.rows(), 8 * Cout, st);   \
    if (avec) return launch_rows_gemm<true, false>("convt2_fwd", ga, wt, so, g.rows(), 8 * Cout, st);          \
    if (ovec) return launch_rows_gemm<false, true>("convt2_fwd", ga, wt, so, g.rows(), 8 * Cout, st);          \
    return launch_rows_gemm<false, false>("convt2_fwd", ga, wt, so, g.rows(), 8 * Cout, st);                   \
  } while (0)
  if (dtype == B200_F32) RUN(float);
  if (dtype == B200_BF16) RUN(__nv_bfloat16);
#undef RUN
  B200_FAIL(B200_ERR_UNSUPPORTED, "convt2_fwd: unknown dtype %d", dtype);
}

extern "C" int b200_convt2_bwd_data(int dtype, const void* gy, const float* w, void* gx, int N, int D, int H, int W, int Cin,
                                    int Cout, void* stream) {
  B200_REQUIRE(gy && w && gx && N > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, B200_ERR_SHAPE, "convt2_bwd_data: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200_BF16 && convt_tc_enabled() && b200_convt2_tc_supported(Cin, Cout))
    return b200_convt2_bwd_data_tc(gy, w, gx, N, D, H, W, Cin, Cout, st);
  const Geom g{N, D, H, W};
  const bool avec = Cout % 8 == 0, ovec = Cin % 8 == 0;
#define RUN(T)                                                                                               \
  do {                                                                                                       \
    GatherChild<T> ga{(const T*)gy, Cout, g};                                                                \
    WeightConvTBwd<T> wt{w, Cin, Cout};                                                                      \
    StoreRows<T> so{(T*)gx, nullptr, Cin, 0, nullptr, g};                                                    \
    if (avec && ovec) return launch_rows_gemm<true, true>("convt2_bwd_data", ga, wt, so, g.rows(), Cin, st);     \
    if (avec) return launch_rows_gemm<true, false>("convt2_bwd_data", ga, wt, so, g.rows(), Cin, st);            \
    if (ovec) return launch_rows_gemm<false, true>("convt2_bwd_data", ga, wt, so, g.rows(), Cin, st);            \
    return launch_rows_gemm<false, false>("convt2_bwd_data", ga, wt, so, g.rows(), Cin, st);                     \
  } while (0)
  if (dtype == B200_F32) RUN(float);
  if (dtype == B200_BF16) RUN(__nv_bfloat16);
#undef RUN
  B200_FAIL(B200_ERR_UNSUPPORTED, "convt2_bwd_data: unknown dtype %d", dtype);
}

extern "C" int64_t b200_convt2_wgrad_workspace(int Cin, int Cout, int N, int D, int H, int W) {
  const SplitPlan sp = plan_split((int64_t)N * D * H * W, Cin, 8 * Cout);
  int64_t main_bytes = (int64_t)sp.nsplit * Cin * 8 * Cout * 4;
  if (b200_convt2_wgrad_tc_supported(Cin, Cout, N, D, H, W)) {
    const int64_t t = b200_convt2_wgrad_tc_workspace(Cin, Cout, N, D, H, W);
    if (t > main_bytes) main_bytes = t;
  }
  return b200_bn_partials_bytes(((Cout + 7) / 8) * 8) + main_bytes;
}

extern "C" int b200_convt2_bwd_weight(int dtype, const void* x, const void* gy, float* dw, float* dbias, void* workspace,
                                      int64_t workspace_bytes, int N, int D, int H, int W, int Cin, int Cout, void* stream) {
  B200_REQUIRE(x && gy && dw && workspace, B200_ERR_SHAPE, "convt2_bwd_weight: null pointer");
  B200_REQUIRE(workspace_bytes >= b200_convt2_wgrad_workspace(Cin, Cout, N, D, H, W), B200_ERR_SHAPE, "convt2_bwd_weight: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const Geom g{N, D, H, W};
  const int64_t M = g.rows();
  float* bpart = (float*)workspace;
  float* partial = (float*)((uint8_t*)workspace + b200_bn_partials_bytes(((Cout + 7) / 8) * 8));
  int rc;
  if (dtype == B200_BF16 && g_wgrad_impl != 1 && b200_convt2_wgrad_tc_supported(Cin, Cout, N, D, H, W)) {
    // the bias gradient rides on the kernel (column sums of the staged gy planes)
    return b200_convt2_wgrad_tc(x, gy, dw, dbias, partial, N, D, H, W, Cin, Cout, st);
  } else {
    const SplitPlan sp = plan_split(M, Cin, 8 * Cout);
    if (dtype == B200_F32) {
      rc = launch_cols_gemm("convt2_bwd_weight", ColsPlain<float>{(const float*)x, Cin, g}, ColsChild<float>{(const float*)gy, Cout, g}, M, Cin, 8 * Cout, sp, partial, st);
    } else if (dtype == B200_BF16) {
      rc = launch_cols_gemm("convt2_bwd_weight", ColsPlain<__nv_bfloat16>{(const __nv_bfloat16*)x, Cin, g},
                            ColsChild<__nv_bfloat16>{(const __nv_bfloat16*)gy, Cout, g}, M, Cin, 8 * Cout, sp, partial, st);
    } else {
      B200_FAIL(B200_ERR_UNSUPPORTED, "convt2_bwd_weight: unknown dtype %d", dtype);
    }
    if (rc) return rc;
    const int64_t total = (int64_t)8 * Cin * Cout;
    reduce_convt_wgrad_kernel<<<b200_grid_for(total, 256, B200_NUM_SMS * 8), 256, 0, st>>>(partial, sp.nsplit, Cin, Cout, dw);
    B200_CHECK_LAUNCH("convt2_wgrad_reduce");
  }
  if (dbias) {
    B200_REQUIRE(Cout % 8 == 0, B200_ERR_UNSUPPORTED, "convt2_bwd_weight: dbias needs Cout %% 8 == 0");
    return b200_channel_sum(dtype, gy, M * 8, Cout, bpart, dbias, stream);
  }
  return B200_OK;
}
