// conv_stem.cu — the network's first convolution: nn.Conv3d(1, 16, k=3, p=1) (models/unet.py:11 via
// encoder[0]), in_channels == 1.  K = 27 only, so this layer is HBM-bound on its 16-channel output
// (134 MB at 2x128^3) and runs on the CUDA cores:
//   * forward : one thread = 2 adjacent voxels x all Cout channels, input halo tile in shared memory,
//               weights broadcast from shared memory, 32-byte vector stores of the NDHWC output;
//   * wgrad   : dW[co][tap] = sum_v x[v+tap] dy[v][co] — persistent CTAs, each thread owns a
//               (kd, kh, 4-channel) slice with 3 kw x 4 co register accumulators and streams an
//               8x8x8 voxel tile out of shared memory; per-CTA partials are reduced in fixed order.
// (No data gradient: the network input does not require one.)
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int kMaxCout = 32;

// ------------------------------------------------------------------ forward
constexpr int FT_W = 64, FT_H = 8;               // output tile per CTA: 64 (w) x 8 (h) x 1 (d)
constexpr int FX_W = FT_W + 2, FX_H = FT_H + 2;  // halo tile
constexpr int kFwdThreads = 256;                 // 32 (w pairs) x 8 (h)

template <typename T, int COUT>
__global__ void __launch_bounds__(kFwdThreads)
stem_fwd_kernel(const T* __restrict__ x, const T* __restrict__ wpack /*[27][COUT]*/, const float* __restrict__ bias,
                T* __restrict__ y, int N, int D, int H, int W, int tiles_w, int tiles_h) {
  __shared__ float xs[3][FX_H][FX_W + 2];
  __shared__ __align__(16) float ws[27][COUT];
  const int tw = blockIdx.x % tiles_w, th = blockIdx.x / tiles_w % tiles_h;
  const int nd = blockIdx.x / (tiles_w * tiles_h);
  const int n = nd / D, d = nd % D;
  const int w0 = tw * FT_W, h0 = th * FT_H;
  for (int i = threadIdx.x; i < 27 * COUT; i += kFwdThreads) ws[i / COUT][i % COUT] = to_f32<T>(wpack[i]);
  for (int i = threadIdx.x; i < 3 * FX_H * FX_W; i += kFwdThreads) {
    const int ww = i % FX_W, hh = (i / FX_W) % FX_H, dd = i / (FX_W * FX_H);
    const int gd = d + dd - 1, gh = h0 + hh - 1, gw = w0 + ww - 1;
    float v = 0.f;
    if ((unsigned)gd < (unsigned)D && (unsigned)gh < (unsigned)H && (unsigned)gw < (unsigned)W)
      v = to_f32<T>(x[(((int64_t)n * D + gd) * H + gh) * W + gw]);
    xs[dd][hh][ww] = v;
  }
  __syncthreads();
  const int lw = (threadIdx.x & 31) * 2, lh = threadIdx.x >> 5;
  float acc0[COUT], acc1[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc0[c] = acc1[c] = bias ? bias[c] : 0.f;
#pragma unroll
  for (int kd = 0; kd < 3; ++kd)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const float* row = &xs[kd][lh + kh][lw];
      const float x0 = row[0], x1 = row[1], x2 = row[2], x3 = row[3];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float a = kw == 0 ? x0 : (kw == 1 ? x1 : x2);
        const float b = kw == 0 ? x1 : (kw == 1 ? x2 : x3);
        const int tap = kd * 9 + kh * 3 + kw;
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
          const float4 wv = *reinterpret_cast<const float4*>(&ws[tap][c4 * 4]);
          acc0[c4 * 4 + 0] = fmaf(a, wv.x, acc0[c4 * 4 + 0]); acc1[c4 * 4 + 0] = fmaf(b, wv.x, acc1[c4 * 4 + 0]);
          acc0[c4 * 4 + 1] = fmaf(a, wv.y, acc0[c4 * 4 + 1]); acc1[c4 * 4 + 1] = fmaf(b, wv.y, acc1[c4 * 4 + 1]);
          acc0[c4 * 4 + 2] = fmaf(a, wv.z, acc0[c4 * 4 + 2]); acc1[c4 * 4 + 2] = fmaf(b, wv.z, acc1[c4 * 4 + 2]);
          acc0[c4 * 4 + 3] = fmaf(a, wv.w, acc0[c4 * 4 + 3]); acc1[c4 * 4 + 3] = fmaf(b, wv.w, acc1[c4 * 4 + 3]);
        }
      }
    }
  const int gh = h0 + lh, gw = w0 + lw;
  if (gh < H) {
    const int64_t row = (((int64_t)n * D + d) * H + gh) * W + gw;
#pragma unroll
    for (int c8 = 0; c8 < COUT / 8; ++c8) {
      float f[8];
      if (gw < W) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = acc0[c8 * 8 + i];
        Vec8<T> v; v.set(f); v.store(y + row * COUT + c8 * 8);
      }
      if (gw + 1 < W) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = acc1[c8 * 8 + i];
        Vec8<T> v; v.set(f); v.store(y + (row + 1) * COUT + c8 * 8);
      }
    }
  }
}

// ------------------------------------------------------------------ weight gradient
constexpr int WT = 8;                              // 8 x 8 x 8 voxel tile
constexpr int kWgMaxBlocks = 592;

template <typename T, int COUT>
__global__ void __launch_bounds__(8 * 9 * (COUT / 4))
stem_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ partials /*[blocks][27*COUT]*/,
                  int N, int D, int H, int W, int tiles_w, int tiles_h, int tiles_d) {
  constexpr int CQ = COUT / 4;            // channel quads
  constexpr int TPG = 9 * CQ;             // threads per group: (kd, kh) x quad
  __shared__ float xs[WT + 2][WT + 2][WT + 4];
  __shared__ __align__(16) float ds[WT * WT * WT][COUT];
  const int grp = threadIdx.x / TPG, t = threadIdx.x % TPG;   // group = d-plane of the tile
  const int kd = t / (3 * CQ), kh = (t / CQ) % 3, cq = t % CQ;
  float acc[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int64_t ntiles = (int64_t)N * tiles_d * tiles_h * tiles_w;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int64_t r = tile;
    const int tw = (int)(r % tiles_w); r /= tiles_w;
    const int th = (int)(r % tiles_h); r /= tiles_h;
    const int td = (int)(r % tiles_d);
    const int n = (int)(r / tiles_d);
    const int w0 = tw * WT, h0 = th * WT, d0 = td * WT;
    __syncthreads();
    for (int i = threadIdx.x; i < (WT + 2) * (WT + 2) * (WT + 2); i += blockDim.x) {
      const int ww = i % (WT + 2), hh = (i / (WT + 2)) % (WT + 2), dd = i / ((WT + 2) * (WT + 2));
      const int gd = d0 + dd - 1, gh = h0 + hh - 1, gw = w0 + ww - 1;
      float v = 0.f;
      if ((unsigned)gd < (unsigned)D && (unsigned)gh < (unsigned)H && (unsigned)gw < (unsigned)W)
        v = to_f32<T>(x[(((int64_t)n * D + gd) * H + gh) * W + gw]);
      xs[dd][hh][ww] = v;
    }
    for (int i = threadIdx.x; i < WT * WT * WT * (COUT / 8); i += blockDim.x) {
      const int c8 = i % (COUT / 8), vox = i / (COUT / 8);
      const int ww = vox % WT, hh = (vox / WT) % WT, dd = vox / (WT * WT);
      const int gd = d0 + dd, gh = h0 + hh, gw = w0 + ww;
      float f[8];
      if (gd < D && gh < H && gw < W) {
        Vec8<T> v;
        v.load(dy + ((((int64_t)n * D + gd) * H + gh) * W + gw) * COUT + c8 * 8);
        v.get(f);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) ds[vox][c8 * 8 + k] = f[k];
    }
    __syncthreads();
    // group grp streams d-plane grp of the tile
#pragma unroll 2
    for (int hh = 0; hh < WT; ++hh) {
      const float* xr = &xs[grp + kd][hh + kh][0];
      float xa = xr[0], xb = xr[1];
#pragma unroll
      for (int ww = 0; ww < WT; ++ww) {
        const float xc = xr[ww + 2];
        const float4 g = *reinterpret_cast<const float4*>(&ds[(grp * WT + hh) * WT + ww][cq * 4]);
        acc[0][0] = fmaf(xa, g.x, acc[0][0]); acc[0][1] = fmaf(xa, g.y, acc[0][1]); acc[0][2] = fmaf(xa, g.z, acc[0][2]); acc[0][3] = fmaf(xa, g.w, acc[0][3]);
        acc[1][0] = fmaf(xb, g.x, acc[1][0]); acc[1][1] = fmaf(xb, g.y, acc[1][1]); acc[1][2] = fmaf(xb, g.z, acc[1][2]); acc[1][3] = fmaf(xb, g.w, acc[1][3]);
        acc[2][0] = fmaf(xc, g.x, acc[2][0]); acc[2][1] = fmaf(xc, g.y, acc[2][1]); acc[2][2] = fmaf(xc, g.z, acc[2][2]); acc[2][3] = fmaf(xc, g.w, acc[2][3]);
        xa = xb; xb = xc;
      }
    }
  }
  // reduce the 8 groups through shared memory (reuse ds), then one partial row per CTA
  __syncthreads();
  float* red = &ds[0][0];  // [8][27*COUT]
#pragma unroll
  for (int kw = 0; kw < 3; ++kw)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[grp * (27 * COUT) + (kd * 9 + kh * 3 + kw) * COUT + cq * 4 + j] = acc[kw][j];
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) s += red[gi * (27 * COUT) + i];
    partials[(int64_t)blockIdx.x * (27 * COUT) + i] = s;
  }
}

// partials[b][tap][co] -> dw[co][0][tap]
__global__ void stem_wgrad_finalize_kernel(const float* __restrict__ partials, int nblocks, int Cout, float* __restrict__ dw) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (o >= 27 * Cout) return;
  const int tap = o % 27, co = o / 27;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += (double)partials[(int64_t)b * 27 * Cout + tap * Cout + co];
  s = warp_sum_d(s);
  if (lane == 0) dw[o] = (float)s;
}


// ================================================================== tensor-core variants (bf16, Cout = 16)
// With one input channel the layer is a GEMM with K = 27 (padded to 32): far too small for tcgen05 tiles, but a good fit
// for warp-level mma.sync.m16n8k16, whose A fragment can be gathered straight out of the shared-memory halo tile
// (no im2col buffer): the FMA-bound CUDA-core kernels above spend 160 / 225 us on a layer whose output is 134 MB.
using bf16 = __nv_bfloat16;
constexpr int SW = 32, SH = 8, SD = 4;                     // CTA tile: 32 (w) x 8 (h) x 4 (d) voxels, warp = one h row
constexpr int XW = SW + 2, XH = SH + 2, XD = SD + 2;       // halo tile
constexpr int XP = 36;                                     // row pitch (elements)
constexpr int kMmaThreads = 256;

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(uint16_t lo, uint16_t hi) { return (uint32_t)lo | ((uint32_t)hi << 16); }
// halo-tile offset of filter tap k (kd, kh, kw) = (k / 9, k / 3 % 3, k % 3); padded taps 27..31 read element 0 (their weight is 0)
__device__ __forceinline__ int tap_off(int k) { return k < 27 ? ((k / 9) * XH + (k / 3) % 3) * XP + k % 3 : 0; }

__device__ __forceinline__ void load_halo(uint16_t* xs, const bf16* __restrict__ x, int n, int d0, int h0, int w0, int D, int H, int W) {
  const uint16_t* xr = reinterpret_cast<const uint16_t*>(x);
  for (int i = threadIdx.x; i < XD * XH * XW; i += kMmaThreads) {
    const int ww = i % XW, hh = (i / XW) % XH, dd = i / (XW * XH);
    const int gd = d0 + dd - 1, gh = h0 + hh - 1, gw = w0 + ww - 1;
    uint16_t v = 0;
    if ((unsigned)gd < (unsigned)D && (unsigned)gh < (unsigned)H && (unsigned)gw < (unsigned)W)
      v = __ldg(xr + (((int64_t)n * D + gd) * H + gh) * W + gw);
    xs[(dd * XH + hh) * XP + ww] = v;
  }
}

__global__ void __launch_bounds__(kMmaThreads)
stem_fwd_mma_kernel(const bf16* __restrict__ x, const bf16* __restrict__ wpack /*[27][16]*/, const float* __restrict__ bias,
                    bf16* __restrict__ y, int N, int D, int H, int W, int tiles_w, int tiles_h, int tiles_d, float* __restrict__ stats) {
  // stats (optional): per-CTA BatchNorm partial sums [block][2][16] of (stored output - bias) and its square — the contract of the
  // tcgen05 kernels' fused statistics (conv_tc3 / conv_tc4), so that the separate 134 MB bn_stats pass over y disappears
  __shared__ uint16_t xs[XD * XH * XP];
  __shared__ float sred[8][32];
  int r = blockIdx.x;
  const int tw = r % tiles_w; r /= tiles_w;
  const int th = r % tiles_h; r /= tiles_h;
  const int td = r % tiles_d;
  const int n = r / tiles_d;
  const int w0 = tw * SW, h0 = th * SH, d0 = td * SD;
  load_halo(xs, x, n, d0, h0, w0, D, H, W);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // B fragments: B[k = tap][n = co]; k-step s, n-tile nt
  const uint16_t* wr = reinterpret_cast<const uint16_t*>(wpack);
  uint32_t bf[2][2][2];
  int off[2][4];
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    const int kk[4] = {16 * s2 + 2 * t, 16 * s2 + 2 * t + 1, 16 * s2 + 2 * t + 8, 16 * s2 + 2 * t + 9};
#pragma unroll
    for (int j = 0; j < 4; ++j) off[s2][j] = tap_off(kk[j]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      uint16_t wv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = kk[j] < 27 ? __ldg(wr + kk[j] * 16 + nt * 8 + g) : (uint16_t)0;
      bf[s2][nt][0] = pack2(wv[0], wv[1]);
      bf[s2][nt][1] = pack2(wv[2], wv[3]);
    }
  }
  float bz[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) { bz[nt][0] = bias ? bias[nt * 8 + 2 * t] : 0.f; bz[nt][1] = bias ? bias[nt * 8 + 2 * t + 1] : 0.f; }
  __syncthreads();
  const int h = warp, gh = h0 + h;
  float ssum[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, ssq[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  if (gh >= H && !stats) return;
#pragma unroll 2
  for (int mt = 0; mt < SD * 2; ++mt) {
    const int d = mt >> 1, wl = (mt & 1) * 16, gd = d0 + d;
    if (gd >= D || gh >= H) break;
    if (w0 + wl >= W) continue;
    const uint16_t* base = xs + (d * XH + h) * XP + wl + g;
    uint32_t a[2][4];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      a[s2][0] = pack2(base[off[s2][0]], base[off[s2][1]]);
      a[s2][1] = pack2(base[8 + off[s2][0]], base[8 + off[s2][1]]);
      a[s2][2] = pack2(base[off[s2][2]], base[off[s2][3]]);
      a[s2][3] = pack2(base[8 + off[s2][2]], base[8 + off[s2][3]]);
    }
    float c[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      c[nt][0] = c[nt][2] = bz[nt][0];
      c[nt][1] = c[nt][3] = bz[nt][1];
      mma_16816(c[nt], a[0], bf[0][nt][0], bf[0][nt][1]);
      mma_16816(c[nt], a[1], bf[1][nt][0], bf[1][nt][1]);
    }
    const int gw = w0 + wl + g;
    bf16* yr = y + ((((int64_t)n * D + gd) * H + gh) * W + gw) * 16 + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(c[nt][0], c[nt][1]), hi = __floats2bfloat162_rn(c[nt][2], c[nt][3]);
      if (gw < W) *reinterpret_cast<__nv_bfloat162*>(yr + nt * 8) = lo;
      if (gw + 8 < W) *reinterpret_cast<__nv_bfloat162*>(yr + 8 * 16 + nt * 8) = hi;
      if (stats) {
        const float2 l = __bfloat1622float2(lo), u = __bfloat1622float2(hi);
        if (gw < W) {
          const float e0 = l.x - bz[nt][0], e1 = l.y - bz[nt][1];
          ssum[nt][0] += e0; ssq[nt][0] = fmaf(e0, e0, ssq[nt][0]);
          ssum[nt][1] += e1; ssq[nt][1] = fmaf(e1, e1, ssq[nt][1]);
        }
        if (gw + 8 < W) {
          const float e0 = u.x - bz[nt][0], e1 = u.y - bz[nt][1];
          ssum[nt][0] += e0; ssq[nt][0] = fmaf(e0, e0, ssq[nt][0]);
          ssum[nt][1] += e1; ssq[nt][1] = fmaf(e1, e1, ssq[nt][1]);
        }
      }
    }
  }
  if (stats) {
    // fixed-shape fold: the eight lanes that share t (xor 4, 8, 16), then the eight warps in order; channel = nt * 8 + 2 t + j
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float a = ssum[nt][j], b = ssq[nt][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (g == 0) { sred[warp][nt * 8 + 2 * t + j] = a; sred[warp][16 + nt * 8 + 2 * t + j] = b; }
      }
    __syncthreads();
    if (threadIdx.x < 32) {
      float tot = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) tot += sred[wq][threadIdx.x];
      stats[(size_t)blockIdx.x * 32 + threadIdx.x] = tot;
    }
  }
}

// dW[tap][co] = sum_v x[v + tap] * dy[v][co]:  M = taps (32 = two m-tiles), N = co (two n-tiles), K = 16 voxels along w
__global__ void __launch_bounds__(kMmaThreads)
stem_wgrad_mma_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ partials /*[blocks][27*16]*/,
                      int N, int D, int H, int W, int tiles_w, int tiles_h, int tiles_d) {
  __shared__ uint16_t xs[XD * XH * XP];
  __shared__ float red[8][32 * 16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint16_t* dyr = reinterpret_cast<const uint16_t*>(dy);
  int offm[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) { offm[mt][0] = tap_off(16 * mt + g); offm[mt][1] = tap_off(16 * mt + g + 8); }
  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
  const int64_t ntiles = (int64_t)N * tiles_d * tiles_h * tiles_w;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int64_t r = tile;
    const int tw = (int)(r % tiles_w); r /= tiles_w;
    const int th = (int)(r % tiles_h); r /= tiles_h;
    const int td = (int)(r % tiles_d);
    const int n = (int)(r / tiles_d);
    const int w0 = tw * SW, h0 = th * SH, d0 = td * SD;
    __syncthreads();
    load_halo(xs, x, n, d0, h0, w0, D, H, W);
    __syncthreads();
    const int h = warp, gh = h0 + h;
    if (gh >= H) continue;
#pragma unroll 2
    for (int ks = 0; ks < SD * 2; ++ks) {
      const int d = ks >> 1, wl = (ks & 1) * 16, gd = d0 + d;
      if (gd >= D || w0 + wl >= W) continue;
      // B fragments: voxels (k) 2t, 2t+1, 2t+8, 2t+9 of this 16-voxel run, channel g (+8)
      const int gw = w0 + wl + 2 * t;
      const uint16_t* dp = dyr + ((((int64_t)n * D + gd) * H + gh) * W + gw) * 16 + g;
      uint16_t q[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        q[nt][0] = gw < W ? __ldg(dp + nt * 8) : (uint16_t)0;
        q[nt][1] = gw + 1 < W ? __ldg(dp + 16 + nt * 8) : (uint16_t)0;
        q[nt][2] = gw + 8 < W ? __ldg(dp + 8 * 16 + nt * 8) : (uint16_t)0;
        q[nt][3] = gw + 9 < W ? __ldg(dp + 9 * 16 + nt * 8) : (uint16_t)0;
      }
      // A fragments: row m = tap, column k = voxel
      const uint16_t* base = xs + (d * XH + h) * XP + wl + 2 * t;
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        a[mt][0] = pack2(base[offm[mt][0]], base[offm[mt][0] + 1]);
        a[mt][1] = pack2(base[offm[mt][1]], base[offm[mt][1] + 1]);
        a[mt][2] = pack2(base[offm[mt][0] + 8], base[offm[mt][0] + 9]);
        a[mt][3] = pack2(base[offm[mt][1] + 8], base[offm[mt][1] + 9]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const uint32_t b0 = pack2(q[nt][0], q[nt][1]), b1 = pack2(q[nt][2], q[nt][3]);
        mma_16816(acc[0][nt], a[0], b0, b1);
        mma_16816(acc[1][nt], a[1], b0, b1);
      }
    }
  }
  // C[m = 16 mt + g (+8)][n = 8 nt + 2t (+1)] -> red[warp][tap][co]; then the eight warps are summed in fixed order
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float* rw = &red[warp][0];
      rw[(16 * mt + g) * 16 + 8 * nt + 2 * t] = acc[mt][nt][0];
      rw[(16 * mt + g) * 16 + 8 * nt + 2 * t + 1] = acc[mt][nt][1];
      rw[(16 * mt + g + 8) * 16 + 8 * nt + 2 * t] = acc[mt][nt][2];
      rw[(16 * mt + g + 8) * 16 + 8 * nt + 2 * t + 1] = acc[mt][nt][3];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * 16; i += kMmaThreads) {
    float s2 = 0.f;
#pragma unroll
    for (int wi = 0; wi < 8; ++wi) s2 += red[wi][i];
    partials[(int64_t)blockIdx.x * (27 * 16) + i] = s2;
  }
}

inline bool stem_mma_enabled() {
  static const int on = [] { const char* e = getenv("B200_STEM_MMA"); return e ? atoi(e) : 1; }();
  return on != 0;
}

inline int wg_blocks(int N, int D, int H, int W) {
  const int64_t tiles = (int64_t)N * ((D + WT - 1) / WT) * ((H + WT - 1) / WT) * ((W + WT - 1) / WT);
  return (int)(tiles < kWgMaxBlocks ? tiles : kWgMaxBlocks);
}

}  // namespace

bool b200_conv_stem_supported(int c0, int c1, int cout) { return c0 == 1 && c1 == 0 && (cout == 16 || cout == 32 || cout == 8); }
bool b200_conv_stem_wgrad_supported(int c0, int c1, int cout) { return c0 == 1 && c1 == 0 && (cout == 16 || cout == 8); }

// rows of fused BatchNorm partial sums the bf16 tensor-core stem kernel writes (0: that kernel does not serve the problem)
int b200_conv_stem_stats_blocks(int dtype, int cout, int N, int D, int H, int W) {
  if (dtype != B200_BF16 || cout != 16 || !stem_mma_enabled()) return 0;
  static int on = -1;             // env B200_STEM_STATS=0: separate bn_stats pass (A/B knob)
  if (on < 0) { const char* e = getenv("B200_STEM_STATS"); on = e ? atoi(e) : 1; }
  if (!on) return 0;
  const int64_t g = (int64_t)((W + SW - 1) / SW) * ((H + SH - 1) / SH) * ((D + SD - 1) / SD) * N;
  return g < (1 << 20) ? (int)g : 0;
}

int b200_conv_stem_fwd(int dtype, const void* x, const void* wpack, const float* bias, void* y, int cout, int N, int D, int H, int W,
                       cudaStream_t st, float* stats) {
  if (dtype == B200_BF16 && cout == 16 && stem_mma_enabled()) {
    const int tw = (W + SW - 1) / SW, th = (H + SH - 1) / SH, td = (D + SD - 1) / SD;
    const int64_t g = (int64_t)tw * th * td * N;
    B200_REQUIRE(g < 2147483647LL, B200_ERR_UNSUPPORTED, "conv_stem_fwd: volume too large");
    stem_fwd_mma_kernel<<<(unsigned)g, kMmaThreads, 0, st>>>((const bf16*)x, (const bf16*)wpack, bias, (bf16*)y, N, D, H, W, tw, th, td, stats);
    B200_CHECK_LAUNCH("conv_stem_fwd_mma");
    return B200_OK;
  }
  B200_REQUIRE(stats == nullptr, B200_ERR_UNSUPPORTED, "conv_stem_fwd: fused statistics need the bf16 16-channel tensor-core kernel");
  const int tiles_w = (W + FT_W - 1) / FT_W, tiles_h = (H + FT_H - 1) / FT_H;
  const int64_t grid = (int64_t)tiles_w * tiles_h * N * D;
  B200_REQUIRE(grid < 2147483647LL, B200_ERR_UNSUPPORTED, "conv_stem_fwd: volume too large");
#define RUN(T, C) stem_fwd_kernel<T, C><<<(unsigned)grid, kFwdThreads, 0, st>>>((const T*)x, (const T*)wpack, bias, (T*)y, N, D, H, W, tiles_w, tiles_h)
  if (dtype == B200_F32) { if (cout == 8) RUN(float, 8); else if (cout == 16) RUN(float, 16); else RUN(float, 32); }
  else { if (cout == 8) RUN(__nv_bfloat16, 8); else if (cout == 16) RUN(__nv_bfloat16, 16); else RUN(__nv_bfloat16, 32); }
#undef RUN
  B200_CHECK_LAUNCH("conv_stem_fwd");
  return B200_OK;
}

int64_t b200_conv_stem_wgrad_workspace(int cout) { return (int64_t)kWgMaxBlocks * 27 * cout * 4; }

int b200_conv_stem_wgrad(int dtype, const void* x, const void* dy, int cout, float* dw, float* partials, int N, int D, int H, int W,
                         cudaStream_t st) {
  if (dtype == B200_BF16 && cout == 16 && stem_mma_enabled()) {
    const int tw = (W + SW - 1) / SW, th = (H + SH - 1) / SH, td = (D + SD - 1) / SD;
    const int64_t tiles = (int64_t)tw * th * td * N;
    const int nb = (int)(tiles < kWgMaxBlocks ? tiles : kWgMaxBlocks);
    stem_wgrad_mma_kernel<<<nb, kMmaThreads, 0, st>>>((const bf16*)x, (const bf16*)dy, partials, N, D, H, W, tw, th, td);
    B200_CHECK_LAUNCH("conv_stem_wgrad_mma");
    stem_wgrad_finalize_kernel<<<(27 * cout * 32 + 127) / 128, 128, 0, st>>>(partials, nb, cout, dw);
    B200_CHECK_LAUNCH("conv_stem_wgrad_finalize");
    return B200_OK;
  }
  const int tiles_w = (W + WT - 1) / WT, tiles_h = (H + WT - 1) / WT, tiles_d = (D + WT - 1) / WT;
  const int nblocks = wg_blocks(N, D, H, W);
#define RUN(T, C) stem_wgrad_kernel<T, C><<<nblocks, 8 * 9 * (C / 4), 0, st>>>((const T*)x, (const T*)dy, partials, N, D, H, W, tiles_w, tiles_h, tiles_d)
  if (dtype == B200_F32) { if (cout == 8) RUN(float, 8); else RUN(float, 16); }
  else { if (cout == 8) RUN(__nv_bfloat16, 8); else RUN(__nv_bfloat16, 16); }
#undef RUN
  B200_CHECK_LAUNCH("conv_stem_wgrad");
  stem_wgrad_finalize_kernel<<<(27 * cout * 32 + 127) / 128, 128, 0, st>>>(partials, nblocks, cout, dw);
  B200_CHECK_LAUNCH("conv_stem_wgrad_finalize");
  return B200_OK;
}
