// head_fused.cu — the network's head as two single-pass, HBM-bound kernels (bf16 activations, 16 channels, <= 4 classes).
//
// Reference call sites fused here: models/unet.py:16-18 (last BatchNorm3d + ReLU of decoder[-1]), :62,87 (final 1x1x1 conv),
// utils/metrics.py:14-40 / :137-167 (softmax + CE + Dice / Tversky sums), :65-129 (argmax + per-class counts).
//
// Forward: one read of the PRE-BatchNorm activation (32 B / voxel) and of the labels (1 or 8 B), one write of the fp32 NCDHW
// logits (16 B): the normalised activation is never materialised.  Replaces bn_act_fwd (64 B/voxel) + conv1x1_fwd (48) +
// seg_loss_fwd (24) + confusion (24).  Backward: reads logits, labels and the pre-BN activation again, recomputes softmax and
// the normalised activation, and produces in one pass the loss gradient, the 1x1 conv's weight / bias gradient partials, the
// gradient w.r.t. the normalised activation (bf16, 32 B written) and the two BatchNorm-backward channel sums.  Replaces
// seg_loss_bwd (40) + conv1x1_bwd (80) + bn_act_bwd_reduce (64).
// Arithmetic is the unfused kernels' arithmetic (same rounding points: BN output and gx rounded to bf16, logits optionally
// rounded to bf16), so that logits and confusion counts are bit-identical to the unfused path.
#include "common.cuh"

namespace {

using bf16 = __nv_bfloat16;
constexpr int kThreads = 256;
constexpr int CIN = 16;
constexpr int CMAX = 4;
constexpr int kMaxBlocks = 2 * B200_NUM_SMS;     // backward: 2 resident CTAs per SM, one wave

__device__ __forceinline__ bool better(float cand, float best) { return (cand > best) || ((cand != cand) && (best == best)); }

struct Softmax4 {
  float p[CMAX];
  float lse;
  __device__ __forceinline__ void compute(const float (&z)[CMAX], int C) {
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) m = fmaxf(m, z[c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { p[c] = expf(z[c] - m); s += p[c]; }
    const float inv = 1.f / s;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) p[c] *= inv;
    lse = m + logf(s);
  }
};

// y = relu(bf16(fma(x - mean, scale, shift))) for the 16 channels of one voxel; returns the values as fp32
__device__ __forceinline__ void bn_relu16(const uint4 (&raw)[2], const float* __restrict__ sc, const float* __restrict__ sh,
                                          const float* __restrict__ mu, float (&xc)[CIN], float (&y)[CIN]) {
  const uint32_t w[8] = {raw[0].x, raw[0].y, raw[0].z, raw[0].w, raw[1].x, raw[1].y, raw[1].z, raw[1].w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xffff0000u);
    xc[2 * i] = lo - mu[2 * i];
    xc[2 * i + 1] = hi - mu[2 * i + 1];
  }
#pragma unroll
  for (int k = 0; k < CIN; ++k) {
    const float t = __bfloat162float(__float2bfloat16_rn(fmaf(xc[k], sc[k], sh[k])));
    y[k] = fmaxf(t, 0.f);
  }
}

// BatchNorm + ReLU of one voxel's 16 channels with the parameters read from shared memory as float4 (one read serves the U
// voxels a thread handles per iteration); the bf16 rounding and the ReLU run on packed pairs (one F2FP + one HMNMX2 per pair).
template <int U>
__device__ __forceinline__ void bn_relu16_multi(const uint4 (&raw)[U][2], const float4* __restrict__ sc4, const float4* __restrict__ sh4,
                                                const float4* __restrict__ mu4, float (&y)[U][CIN]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 sc = sc4[q], sh = sh4[q], mu = mu4[q];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t w0 = q < 2 ? (q == 0 ? raw[u][0].x : raw[u][0].z) : (q == 2 ? raw[u][1].x : raw[u][1].z);
      const uint32_t w1 = q < 2 ? (q == 0 ? raw[u][0].y : raw[u][0].w) : (q == 2 ? raw[u][1].y : raw[u][1].w);
      const float t0 = fmaf(__uint_as_float(w0 << 16) - mu.x, sc.x, sh.x), t1 = fmaf(__uint_as_float(w0 & 0xffff0000u) - mu.y, sc.y, sh.y);
      const float t2 = fmaf(__uint_as_float(w1 << 16) - mu.z, sc.z, sh.z), t3 = fmaf(__uint_as_float(w1 & 0xffff0000u) - mu.w, sc.w, sh.w);
      const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
      const __nv_bfloat162 p01 = __hmax2(__floats2bfloat162_rn(t0, t1), zero), p23 = __hmax2(__floats2bfloat162_rn(t2, t3), zero);
      const uint32_t b01 = *reinterpret_cast<const uint32_t*>(&p01), b23 = *reinterpret_cast<const uint32_t*>(&p23);
      y[u][4 * q] = __uint_as_float(b01 << 16);
      y[u][4 * q + 1] = __uint_as_float(b01 & 0xffff0000u);
      y[u][4 * q + 2] = __uint_as_float(b23 << 16);
      y[u][4 * q + 3] = __uint_as_float(b23 & 0xffff0000u);
    }
  }
}

// sums layout = seg_loss_fwd's (loss_kernels.cu): [0] CE sum, per class k: [4+4k] I, [5+4k] P, [6+4k] T
// ncu of the first version (one voxel per thread and iteration, scalar shared-memory parameter reads): 506 warp instructions per
// 32 voxels, 112 of them LDS, issue-bound at 109 us.  Now: U = 2 voxels per thread share float4 parameter reads.
template <typename LabelT>
__global__ void __launch_bounds__(kThreads)
head_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                const float* __restrict__ w, const float* __restrict__ bias, int round_bf16, const LabelT* __restrict__ target, int64_t N,
                int64_t S, int C, float* __restrict__ logits, double* __restrict__ sums, unsigned long long* __restrict__ conf) {
  constexpr int U = 2;
  __shared__ __align__(16) float ws[CMAX * CIN + 3 * CIN];
  __shared__ unsigned int hist[kThreads / 32][CMAX * CMAX];
  __shared__ float red[kThreads / 32][1 + 3 * CMAX];
  const float4* w4 = reinterpret_cast<const float4*>(ws);
  const float4* sc4 = w4 + CMAX * CIN / 4; const float4* sh4 = sc4 + CIN / 4; const float4* mu4 = sh4 + CIN / 4;
  for (int i = threadIdx.x; i < CMAX * CIN; i += blockDim.x) ws[i] = i < C * CIN ? w[i] : 0.f;
  if (threadIdx.x < CIN) {
    ws[CMAX * CIN + threadIdx.x] = scale[threadIdx.x]; ws[CMAX * CIN + CIN + threadIdx.x] = shift[threadIdx.x];
    ws[CMAX * CIN + 2 * CIN + threadIdx.x] = mean[threadIdx.x];
  }
  float br[CMAX];
#pragma unroll
  for (int co = 0; co < CMAX; ++co) br[co] = (bias && co < C) ? bias[co] : 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < CMAX * CMAX) hist[warp][lane] = 0;
  __syncthreads();

  float ce = 0.f, accI[CMAX], accP[CMAX], accT[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) accI[c] = accP[c] = accT[c] = 0.f;
  // every sample is walked separately (no 64-bit division per voxel); a warp covers 32 consecutive voxels, U such rows per
  // iteration, and all its lanes run the same number of iterations (match_any below needs the full mask)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t n = 0; n < N; ++n) {
    const bf16* xn = x + n * S * CIN;
    const LabelT* tn = target + n * S;
    float* ln = logits + n * C * S;
    for (int64_t vb = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); vb < S; vb += U * stride) {
      uint4 raw[U][2];
      long long yy[U];
      bool act[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = vb + u * stride + lane;
        act[u] = v < S;
        raw[u][0] = raw[u][1] = make_uint4(0, 0, 0, 0);
        yy[u] = -1;
        if (act[u]) {
          raw[u][0] = __ldcs(reinterpret_cast<const uint4*>(xn + v * CIN));
          raw[u][1] = __ldcs(reinterpret_cast<const uint4*>(xn + v * CIN) + 1);
          yy[u] = (long long)tn[v];
        }
      }
      float y[U][CIN], z[U][CMAX];
      bn_relu16_multi<U>(raw, sc4, sh4, mu4, y);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int co = 0; co < CMAX; ++co) z[u][co] = br[co];
      // the unfused 1x1 kernel's fma chain (channels in ascending order from the bias): logits stay bit-identical
#pragma unroll
      for (int co = 0; co < CMAX; ++co)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = w4[co * 4 + q];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            z[u][co] = fmaf(y[u][4 * q], wv.x, z[u][co]);
            z[u][co] = fmaf(y[u][4 * q + 1], wv.y, z[u][co]);
            z[u][co] = fmaf(y[u][4 * q + 2], wv.z, z[u][co]);
            z[u][co] = fmaf(y[u][4 * q + 3], wv.w, z[u][co]);
          }
        }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int key = -1;
        if (act[u]) {
          const int64_t v = vb + u * stride + lane;
#pragma unroll
          for (int co = 0; co < CMAX; ++co) {
            if (round_bf16) z[u][co] = __bfloat162float(__float2bfloat16_rn(z[u][co]));
            if (co < C) ln[co * S + v] = z[u][co];
          }
          Softmax4 sm;
          sm.compute(z[u], C);
          float zy = 0.f, best = z[u][0];
          int arg = 0;
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) {
              accP[c] += sm.p[c];
              if (c == yy[u]) { accI[c] += sm.p[c]; accT[c] += 1.f; zy = z[u][c]; }
              if (c > 0 && better(z[u][c], best)) { best = z[u][c]; arg = c; }
            }
          ce += sm.lse - zy;
          key = (yy[u] >= 0 && yy[u] < C) ? (int)yy[u] * C + arg : -1;
        }
        if (conf) {
          const unsigned peers = __match_any_sync(0xffffffffu, key);
          if (key >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(&hist[warp][key], (unsigned)__popc(peers));
        }
      }
    }
  }
  ce = warp_sum(ce);
#pragma unroll
  for (int c = 0; c < CMAX; ++c) { accI[c] = warp_sum(accI[c]); accP[c] = warp_sum(accP[c]); accT[c] = warp_sum(accT[c]); }
  if (lane == 0) {
    red[warp][0] = ce;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) { red[warp][1 + 3 * c] = accI[c]; red[warp][2 + 3 * c] = accP[c]; red[warp][3 + 3 * c] = accT[c]; }
  }
  __syncthreads();
  if (threadIdx.x < 1 + 3 * CMAX) {
    double t = 0.0;
    for (int wq = 0; wq < kThreads / 32; ++wq) t += (double)red[wq][threadIdx.x];
    const int q = threadIdx.x;
    if (q == 0) atomicAdd(&sums[0], t);
    else {
      const int c = (q - 1) / 3, f = (q - 1) % 3;
      if (c < C) atomicAdd(&sums[4 + 4 * c + f], t);
    }
  }
  if (conf && threadIdx.x < C * C) {
    unsigned long long t = 0;
    for (int wq = 0; wq < kThreads / 32; ++wq) t += hist[wq][threadIdx.x];
    if (t) atomicAdd(&conf[threadIdx.x], t);
  }
}

// coef layout = seg_loss_finalize's: [0] w_ce / Nvox, [1] kd (unused here), [2 + c] a_c, [2 + C + c] b_c
// wpart[block][CMAX * (CIN + 1)] = partial dW (co, ci) and db (co, CIN); bnpart[block][2][CIN] = partial (sum g, invstd * sum g * (x - mean))
//
// Two threads per voxel, eight channels each (lane parity = channel half): 32 dW + 16 BatchNorm accumulators per thread instead of
// 64 + 32, parameters and weights read from shared memory as float4 and shared by the U voxels of an iteration, every global
// load and store a full 16 bytes per lane on consecutive addresses.  (First version: one thread per voxel and all 16 channels —
// 128 registers with 320 bytes of spills, 112 scalar LDS per voxel, 2 x 256 threads per SM: 185-255 us for a 41 us traffic floor.)
// Both threads of a pair evaluate the softmax (cheaper than exchanging it).  gy is bit-identical to the first version (same fma
// chains); dW / db / BatchNorm sums are folded in a different, still fixed, order.
template <typename LabelT, int U>
__global__ void __launch_bounds__(kThreads, 2)
head_bwd_kernel(const float* __restrict__ logits, const LabelT* __restrict__ target, const float* __restrict__ coef, const float* __restrict__ gout,
                const bf16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                const float* __restrict__ invstd, const float* __restrict__ w, int64_t N, int64_t S, int C, bf16* __restrict__ gy,
                float* __restrict__ wpart, float* __restrict__ bnpart) {
  constexpr int P = CMAX * (CIN + 1);
  constexpr int HC = CIN / 2;                           // channels per thread
  __shared__ __align__(16) float ws[CMAX * CIN + 3 * CIN];
  __shared__ float red[kThreads / 32][P + 2 * CIN];
  for (int i = threadIdx.x; i < CMAX * CIN; i += blockDim.x) ws[i] = i < C * CIN ? w[i] : 0.f;
  if (threadIdx.x < CIN) {
    ws[CMAX * CIN + threadIdx.x] = scale[threadIdx.x]; ws[CMAX * CIN + CIN + threadIdx.x] = shift[threadIdx.x];
    ws[CMAX * CIN + 2 * CIN + threadIdx.x] = mean[threadIdx.x];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane & 1, vl = lane >> 1;
  const float4* w4 = reinterpret_cast<const float4*>(ws) + half * 2;                       // + co * 4 + q
  const float4* sc4 = reinterpret_cast<const float4*>(ws + CMAX * CIN) + half * 2;
  const float4* sh4 = sc4 + CIN / 4; const float4* mu4 = sh4 + CIN / 4;
  const float go = gout[0];
  const float w_ce = coef[0] * go;
  float ca[CMAX], cb[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) { ca[c] = c < C ? coef[2 + c] * go : 0.f; cb[c] = c < C ? coef[2 + C + c] * go : 0.f; }
  float dw[CMAX][HC], db[CMAX], a0[HC], a1[HC];
#pragma unroll
  for (int co = 0; co < CMAX; ++co) {
    db[co] = 0.f;
#pragma unroll
    for (int k = 0; k < HC; ++k) dw[co][k] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < HC; ++k) a0[k] = a1[k] = 0.f;

  // a warp covers 16 * U consecutive voxels per iteration; every sample is walked separately (no 64-bit division per voxel)
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * (16 * U);
  for (int64_t n = 0; n < N; ++n) {
    const bf16* xn = x + n * S * CIN;
    bf16* gn = gy + n * S * CIN;
    const LabelT* tn = target + n * S;
    const float* ln = logits + n * C * S;
    for (int64_t vb = ((int64_t)blockIdx.x * (kThreads / 32) + warp) * (16 * U); vb < S; vb += wstride) {
      uint4 raw[U];
      float z[U][CMAX];
      int yy[U];
      bool act[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = vb + u * 16 + vl;
        act[u] = v < S;
        raw[u] = make_uint4(0, 0, 0, 0);
        yy[u] = -1;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) z[u][c] = 0.f;
        if (act[u]) {
          raw[u] = __ldcs(reinterpret_cast<const uint4*>(xn + v * CIN) + half);
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) z[u][c] = __ldcs(ln + c * S + v);
          yy[u] = (int)tn[v];
        }
      }
      // the U voxels are finished one after the other (loads above stay in flight): live state per voxel ~ 40 registers
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float dz[CMAX], xc[HC], y[HC];
        {
          Softmax4 sm;
          sm.compute(z[u], C);
          float wv[CMAX], pw = 0.f;
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) { wv[c] = (c == yy[u] ? ca[c] : 0.f) + cb[c]; pw += sm.p[c] * wv[c]; }
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            dz[c] = (c < C && act[u]) ? w_ce * (sm.p[c] - (c == yy[u] ? 1.f : 0.f)) + sm.p[c] * (wv[c] - pw) : 0.f;
        }
        // BatchNorm + ReLU of this thread's eight channels (the forward kernels' arithmetic: fma in fp32, rounded to bf16, max with 0)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 sc = sc4[q], sh = sh4[q], mu = mu4[q];
          const uint32_t w0 = q == 0 ? raw[u].x : raw[u].z, w1 = q == 0 ? raw[u].y : raw[u].w;
          xc[4 * q] = __uint_as_float(w0 << 16) - mu.x;
          xc[4 * q + 1] = __uint_as_float(w0 & 0xffff0000u) - mu.y;
          xc[4 * q + 2] = __uint_as_float(w1 << 16) - mu.z;
          xc[4 * q + 3] = __uint_as_float(w1 & 0xffff0000u) - mu.w;
          const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
          const __nv_bfloat162 p01 = __hmax2(__floats2bfloat162_rn(fmaf(xc[4 * q], sc.x, sh.x), fmaf(xc[4 * q + 1], sc.y, sh.y)), zero);
          const __nv_bfloat162 p23 = __hmax2(__floats2bfloat162_rn(fmaf(xc[4 * q + 2], sc.z, sh.z), fmaf(xc[4 * q + 3], sc.w, sh.w)), zero);
          const uint32_t b01 = *reinterpret_cast<const uint32_t*>(&p01), b23 = *reinterpret_cast<const uint32_t*>(&p23);
          y[4 * q] = __uint_as_float(b01 << 16);
          y[4 * q + 1] = __uint_as_float(b01 & 0xffff0000u);
          y[4 * q + 2] = __uint_as_float(b23 << 16);
          y[4 * q + 3] = __uint_as_float(b23 & 0xffff0000u);
        }
        // gradient w.r.t. the normalised activation: a[k] = sum_co dz[co] * W[co][k] (classes in ascending order from 0)
        float a[HC];
#pragma unroll
        for (int k = 0; k < HC; ++k) a[k] = 0.f;
#pragma unroll
        for (int co = 0; co < CMAX; ++co)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 wv = w4[co * 4 + q];
            a[4 * q] = fmaf(dz[co], wv.x, a[4 * q]);
            a[4 * q + 1] = fmaf(dz[co], wv.y, a[4 * q + 1]);
            a[4 * q + 2] = fmaf(dz[co], wv.z, a[4 * q + 2]);
            a[4 * q + 3] = fmaf(dz[co], wv.w, a[4 * q + 3]);
          }
        uint32_t packed[HC / 2];
#pragma unroll
        for (int k = 0; k < HC; k += 2) {
          const __nv_bfloat162 r = __floats2bfloat162_rn(a[k], a[k + 1]);      // stored as bf16 ...
          const uint32_t rb = *reinterpret_cast<const uint32_t*>(&r);
          packed[k >> 1] = rb;
          // ... and through the ReLU (y > 0 <=> pre-activation > 0)
          const float g0 = y[k] > 0.f ? __uint_as_float(rb << 16) : 0.f, g1 = y[k + 1] > 0.f ? __uint_as_float(rb & 0xffff0000u) : 0.f;
          a0[k] += g0;
          a0[k + 1] += g1;
          a1[k] = fmaf(g0, xc[k], a1[k]);
          a1[k + 1] = fmaf(g1, xc[k + 1], a1[k + 1]);
        }
        if (act[u]) __stcs(reinterpret_cast<uint4*>(gn + (vb + u * 16 + vl) * CIN) + half, make_uint4(packed[0], packed[1], packed[2], packed[3]));
#pragma unroll
        for (int co = 0; co < CMAX; ++co) {
          if (half == 0) db[co] += dz[co];
#pragma unroll
          for (int k = 0; k < HC; ++k) dw[co][k] = fmaf(dz[co], y[k], dw[co][k]);
        }
      }
    }
  }
  // fixed-shape fold: butterflies over the 16 lanes of one parity, then the warps in order
  auto half_sum = [](float v) {
#pragma unroll
    for (int o = 2; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
#pragma unroll
  for (int co = 0; co < CMAX; ++co) {
#pragma unroll
    for (int k = 0; k < HC; ++k) {
      const float t = half_sum(dw[co][k]);
      if (lane < 2) red[warp][co * (CIN + 1) + half * HC + k] = t;
    }
    const float t = half_sum(db[co]);
    if (lane == 0) red[warp][co * (CIN + 1) + CIN] = t;
  }
#pragma unroll
  for (int k = 0; k < HC; ++k) {
    const float t0 = half_sum(a0[k]), t1 = half_sum(a1[k]);
    if (lane < 2) { red[warp][P + half * HC + k] = t0; red[warp][P + CIN + half * HC + k] = t1; }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < P + 2 * CIN; q += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int wq = 0; wq < kThreads / 32; ++wq) t += red[wq][q];
    if (q < P) wpart[(int64_t)blockIdx.x * P + q] = t;
    else if (q < P + CIN) bnpart[(int64_t)blockIdx.x * 2 * CIN + (q - P)] = t;
    else bnpart[(int64_t)blockIdx.x * 2 * CIN + CIN + (q - P - CIN)] = t * invstd[q - P - CIN];
  }
}

// dW / db: fold wpart over the blocks in fixed order (fp64)
__global__ void head_bwd_finalize_kernel(const float* __restrict__ wpart, int nblocks, int C, float* __restrict__ dw, float* __restrict__ db) {
  constexpr int P = CMAX * (CIN + 1);
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += (double)wpart[(int64_t)b * P + q];
  const int co = q / (CIN + 1), k = q % (CIN + 1);
  if (co >= C) return;
  if (k < CIN) dw[co * CIN + k] = (float)t;
  else if (db) db[co] = (float)t;
}

}  // namespace

extern "C" int b200_head_blocks(int64_t N, int64_t S) { return b200_grid_for(N * S, kThreads, kMaxBlocks); }

extern "C" int b200_head_fwd(const void* x, const float* scale, const float* shift, const float* mean, const float* w, const float* bias,
                             int round_bf16, const void* target, int label_bytes, int64_t N, int64_t S, int Cin, int C, float* logits,
                             double* sums, unsigned long long* conf, void* stream) {
  B200_REQUIRE(x && scale && shift && mean && w && target && logits && sums, B200_ERR_SHAPE, "head_fwd: null pointer");
  B200_REQUIRE(Cin == CIN && C >= 2 && C <= CMAX, B200_ERR_UNSUPPORTED, "head_fwd: serves %d input channels and 2..%d classes (got %d, %d)", CIN, CMAX, Cin, C);
  B200_REQUIRE(label_bytes == 1 || label_bytes == 8, B200_ERR_UNSUPPORTED, "head_fwd: labels must be uint8 or int64");
  B200_REQUIRE(b200_aligned(x, 16), B200_ERR_ALIGN, "head_fwd: activation must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (4 + 4 * C), st));
  if (conf) B200_CUDA(cudaMemsetAsync(conf, 0, sizeof(unsigned long long) * C * C, st));
  const int grid = b200_grid_for(S, 2 * kThreads, B200_NUM_SMS * 6);
  if (label_bytes == 1)
    head_fwd_kernel<uint8_t><<<grid, kThreads, 0, st>>>((const bf16*)x, scale, shift, mean, w, bias, round_bf16, (const uint8_t*)target, N, S, C, logits, sums, conf);
  else
    head_fwd_kernel<int64_t><<<grid, kThreads, 0, st>>>((const bf16*)x, scale, shift, mean, w, bias, round_bf16, (const int64_t*)target, N, S, C, logits, sums, conf);
  B200_CHECK_LAUNCH("head_fwd");
  return B200_OK;
}

// wpart: b200_head_blocks() * 4 * 17 floats, bnpart: b200_head_blocks() * 2 * 16 floats.  Afterwards dw[C][16], db[C] hold the 1x1 conv's
// gradients; bnpart goes to b200_bn_bwd_finalize_ex(bnpart, b200_head_blocks(), ...) and gy to b200_bn_act_bwd_apply.
extern "C" int b200_head_bwd(const float* logits, const void* target, int label_bytes, const float* coef, const float* gout, const void* x,
                             const float* scale, const float* shift, const float* mean, const float* invstd, const float* w, int64_t N,
                             int64_t S, int Cin, int C, void* gy, float* wpart, float* bnpart, float* dw, float* db, void* stream) {
  B200_REQUIRE(logits && target && coef && gout && x && scale && shift && mean && invstd && w && gy && wpart && bnpart && dw, B200_ERR_SHAPE,
               "head_bwd: null pointer");
  B200_REQUIRE(Cin == CIN && C >= 2 && C <= CMAX, B200_ERR_UNSUPPORTED, "head_bwd: serves %d input channels and 2..%d classes (got %d, %d)", CIN, CMAX, Cin, C);
  B200_REQUIRE(label_bytes == 1 || label_bytes == 8, B200_ERR_UNSUPPORTED, "head_bwd: labels must be uint8 or int64");
  B200_REQUIRE(b200_aligned(x, 16) && b200_aligned(gy, 16), B200_ERR_ALIGN, "head_bwd: activations must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = b200_head_blocks(N, S);
  if (label_bytes == 1)
    head_bwd_kernel<uint8_t, 2><<<grid, kThreads, 0, st>>>(logits, (const uint8_t*)target, coef, gout, (const bf16*)x, scale, shift, mean, invstd, w, N, S, C,
                                                        (bf16*)gy, wpart, bnpart);
  else
    head_bwd_kernel<int64_t, 2><<<grid, kThreads, 0, st>>>(logits, (const int64_t*)target, coef, gout, (const bf16*)x, scale, shift, mean, invstd, w, N, S, C,
                                                        (bf16*)gy, wpart, bnpart);
  B200_CHECK_LAUNCH("head_bwd");
  head_bwd_finalize_kernel<<<1, 128, 0, st>>>(wpart, grid, C, dw, db);
  B200_CHECK_LAUNCH("head_bwd_finalize");
  return B200_OK;
}
