// metric_kernels.cu — fused argmax + C x C confusion counts.
//
// Replaces the 3 x (argmax + per-class eq / & / sum + host sync) sequences of
// utils/metrics.py:65-129 (calculate_iou / calculate_dice / calculate_accuracy) with one
// HBM-bound pass: 4C + 8 B/voxel.  Counts are exact integers (int64), so the fp32 recipe of
// SURVEY.md Appendix E applied on the host reproduces the reference bit for bit.
// argmax semantics = torch.argmax: first maximum wins, NaN counts as the maximum.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxC = 16;

__device__ __forceinline__ bool better(float cand, float best) {
  // torch: (cand > best) || (isnan(cand) && !isnan(best))
  return (cand > best) || ((cand != cand) && (best == best));
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
confusion_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int64_t N, int C, int64_t S,
                 unsigned long long* __restrict__ conf) {
  // one histogram per warp: warp-aggregated (match_any) shared-memory atomics, then one
  // global atomic per non-zero cell per block.
  __shared__ unsigned int hist[kThreads / 32][kMaxC * kMaxC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cells = C * C;
  for (int i = lane; i < cells; i += 32) hist[warp][i] = 0;
  __syncwarp();

  const int64_t groups_per_sample = S / VEC;
  const int64_t total_groups = N * groups_per_sample;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // all lanes of a warp run the same number of iterations so that match_any sees a full mask
  const int64_t base0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  for (int64_t gb = base0; gb < total_groups; gb += stride) {
    const int64_t g = gb + lane;
    const bool active = g < total_groups;
    int key[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) key[j] = -1;
    if (active) {
      const int64_t n = g / groups_per_sample;
      const int64_t s0 = (g - n * groups_per_sample) * VEC;
      const float* zp = logits + (n * C) * S + s0;
      float best[VEC];
      int arg[VEC];
      long long y[VEC];
      if (VEC == 4) {
        float4 v = __ldcs(reinterpret_cast<const float4*>(zp));
        best[0] = v.x; best[1 % VEC] = v.y; best[2 % VEC] = v.z; best[3 % VEC] = v.w;
#pragma unroll
        for (int j = 0; j < VEC; ++j) arg[j] = 0;
        for (int c = 1; c < C; ++c) {
          v = __ldcs(reinterpret_cast<const float4*>(zp + c * S));
          const float cand[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            if (better(cand[j], best[j])) { best[j] = cand[j]; arg[j] = c; }
        }
        const longlong2 y01 = __ldcs(reinterpret_cast<const longlong2*>(target + n * S + s0));
        const longlong2 y23 = __ldcs(reinterpret_cast<const longlong2*>(target + n * S + s0 + 2));
        y[0] = y01.x; y[1 % VEC] = y01.y; y[2 % VEC] = y23.x; y[3 % VEC] = y23.y;
      } else {
        best[0] = zp[0];
        arg[0] = 0;
        for (int c = 1; c < C; ++c) {
          const float cand = zp[c * S];
          if (better(cand, best[0])) { best[0] = cand; arg[0] = c; }
        }
        y[0] = target[n * S + s0];
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        key[j] = (y[j] >= 0 && y[j] < C) ? (int)y[j] * C + arg[j] : -1;
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const unsigned peers = __match_any_sync(0xffffffffu, key[j]);
      if (key[j] >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(&hist[warp][key[j]], (unsigned)__popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cells; i += blockDim.x) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += hist[w][i];
    if (t) atomicAdd(&conf[i], t);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
argmax_kernel(const float* __restrict__ logits, int64_t N, int C, int64_t S, uint8_t* __restrict__ out) {
  const int64_t groups_per_sample = S / VEC;
  const int64_t total_groups = N * groups_per_sample;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_groups;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = g / groups_per_sample;
    const int64_t s0 = (g - n * groups_per_sample) * VEC;
    const float* zp = logits + (n * C) * S + s0;
    float best[VEC];
    int arg[VEC];
    if (VEC == 4) {
      float4 v = __ldcs(reinterpret_cast<const float4*>(zp));
      best[0] = v.x; best[1 % VEC] = v.y; best[2 % VEC] = v.z; best[3 % VEC] = v.w;
#pragma unroll
      for (int j = 0; j < VEC; ++j) arg[j] = 0;
      for (int c = 1; c < C; ++c) {
        v = __ldcs(reinterpret_cast<const float4*>(zp + c * S));
        const float cand[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          if (better(cand[j], best[j])) { best[j] = cand[j]; arg[j] = c; }
      }
      uchar4 o = make_uchar4((unsigned char)arg[0], (unsigned char)arg[1 % VEC], (unsigned char)arg[2 % VEC],
                             (unsigned char)arg[3 % VEC]);
      *reinterpret_cast<uchar4*>(out + n * S + s0) = o;
    } else {
      best[0] = zp[0];
      arg[0] = 0;
      for (int c = 1; c < C; ++c) {
        const float cand = zp[c * S];
        if (better(cand, best[0])) { best[0] = cand; arg[0] = c; }
      }
      out[n * S + s0] = (uint8_t)arg[0];
    }
  }
}

__global__ void window_accumulate_kernel(float* __restrict__ acc, float* __restrict__ cnt, const float* __restrict__ logits, int C,
                                         int D, int H, int W, int d0, int h0, int w0, int wd, int wh, int ww) {
  const int64_t wvox = (int64_t)wd * wh * ww;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < wvox; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % ww), y = (int)((i / ww) % wh), z = (int)(i / ((int64_t)ww * wh));
    const int64_t dst = ((int64_t)(d0 + z) * H + (h0 + y)) * W + (w0 + x);
    for (int c = 0; c < C; ++c) acc[(int64_t)c * D * H * W + dst] += logits[(int64_t)c * wvox + i];
    cnt[dst] += 1.f;
  }
}
__global__ void window_finalize_kernel(float* __restrict__ acc, const float* __restrict__ cnt, int C, int64_t S) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (int64_t)gridDim.x * blockDim.x) {
    const float n = cnt[i];
    if (n > 1.f) {
      const float inv = 1.f / n;
      for (int c = 0; c < C; ++c) acc[(int64_t)c * S + i] *= inv;
    }
  }
}

}  // namespace

extern "C" int b200_window_accumulate(float* acc, float* cnt, const float* logits, int C, int D, int H, int W, int d0, int h0, int w0,
                                      int wd, int wh, int ww, void* stream) {
  B200_REQUIRE(acc && cnt && logits, B200_ERR_SHAPE, "window_accumulate: null pointer");
  B200_REQUIRE(C > 0 && wd > 0 && wh > 0 && ww > 0 && d0 >= 0 && h0 >= 0 && w0 >= 0 && d0 + wd <= D && h0 + wh <= H && w0 + ww <= W,
               B200_ERR_SHAPE, "window_accumulate: window [%d,%d,%d]+[%d,%d,%d] outside volume [%d,%d,%d]", d0, h0, w0, wd, wh, ww, D, H, W);
  const int64_t wvox = (int64_t)wd * wh * ww;
  window_accumulate_kernel<<<b200_grid_for(wvox, kThreads, B200_NUM_SMS * 8), kThreads, 0, (cudaStream_t)stream>>>(acc, cnt, logits, C, D, H, W,
                                                                                                                 d0, h0, w0, wd, wh, ww);
  B200_CHECK_LAUNCH("window_accumulate");
  return B200_OK;
}
extern "C" int b200_window_finalize(float* acc, const float* cnt, int C, int64_t S, void* stream) {
  B200_REQUIRE(acc && cnt && C > 0 && S > 0, B200_ERR_SHAPE, "window_finalize: bad arguments");
  window_finalize_kernel<<<b200_grid_for(S, kThreads, B200_NUM_SMS * 8), kThreads, 0, (cudaStream_t)stream>>>(acc, cnt, C, S);
  B200_CHECK_LAUNCH("window_finalize");
  return B200_OK;
}

extern "C" int b200_confusion(const float* logits, const int64_t* target, int64_t N, int C, int64_t S, int64_t* conf,
                              void* stream) {
  B200_REQUIRE(logits && target && conf, B200_ERR_SHAPE, "confusion: null pointer");
  B200_REQUIRE(C >= 1 && C <= kMaxC, B200_ERR_UNSUPPORTED, "confusion: C=%d outside [1,%d]", C, kMaxC);
  B200_REQUIRE(N >= 0 && S >= 0, B200_ERR_SHAPE, "confusion: negative size");
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA(cudaMemsetAsync(conf, 0, sizeof(int64_t) * C * C, st));
  if (N == 0 || S == 0) return B200_OK;
  const bool vec = (S % 4 == 0) && b200_aligned(logits, 16) && b200_aligned(target, 16);
  const int64_t groups = N * (vec ? S / 4 : S);
  const int grid = b200_grid_for(groups, kThreads, B200_NUM_SMS * 8);
  if (vec)
    confusion_kernel<4><<<grid, kThreads, 0, st>>>(logits, target, N, C, S, reinterpret_cast<unsigned long long*>(conf));
  else
    confusion_kernel<1><<<grid, kThreads, 0, st>>>(logits, target, N, C, S, reinterpret_cast<unsigned long long*>(conf));
  B200_CHECK_LAUNCH("confusion");
  return B200_OK;
}

extern "C" int b200_argmax(const float* logits, int64_t N, int C, int64_t S, uint8_t* out, void* stream) {
  B200_REQUIRE(logits && out, B200_ERR_SHAPE, "argmax: null pointer");
  B200_REQUIRE(C >= 1 && C <= 255, B200_ERR_UNSUPPORTED, "argmax: C=%d outside [1,255]", C);
  if (N <= 0 || S <= 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (S % 4 == 0) && b200_aligned(logits, 16) && b200_aligned(out, 4);
  const int64_t groups = N * (vec ? S / 4 : S);
  const int grid = b200_grid_for(groups, kThreads, B200_NUM_SMS * 8);
  if (vec) argmax_kernel<4><<<grid, kThreads, 0, st>>>(logits, N, C, S, out);
  else argmax_kernel<1><<<grid, kThreads, 0, st>>>(logits, N, C, S, out);
  B200_CHECK_LAUNCH("argmax");
  return B200_OK;
}
