// partial_reduce.cuh — fixed-order (deterministic) reduction of per-CTA split-K partials.
//
// partial[group][cta < spatial][e < stride] -> out[map(group, e)].  A block covers EPB = 256 / LANES consecutive source elements
// times LANES partial-lanes: lane l sums partials l, l + LANES, ... in fp64 (four loads in flight), then thread lane 0 adds the
// LANES lane sums in ascending order.  With many partials LANES = 8 and EPB = 32: every warp-load is one whole 128-byte line of
// one partial row and a launch has stride / 32 fat blocks.  (Round 1: LANES = 32, EPB = 8 -> 32-byte pieces and 4x the blocks,
// 15 us per layer for a 14 MB read; before that one warp per output element with 8x read amplification.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

template <typename MapF>
__global__ void __launch_bounds__(256)
partial_reduce_kernel(const float* __restrict__ partial, int spatial, int64_t stride, int lanes, MapF map, float* __restrict__ out) {
  __shared__ double sm[256];
  const int epb = 256 / lanes;
  const int el = threadIdx.x % epb, l = threadIdx.x / epb;
  const int64_t e = (int64_t)blockIdx.x * epb + el;
  const int group = blockIdx.y;
  double s = 0.0;
  if (e < stride) {
    const float* src = partial + (int64_t)group * spatial * stride + e;
    int c = l;
    for (; c + 3 * lanes < spatial; c += 4 * lanes) {      // four independent loads in flight, summed in ascending order
      const float v0 = __ldcs(src + (int64_t)c * stride), v1 = __ldcs(src + (int64_t)(c + lanes) * stride);
      const float v2 = __ldcs(src + (int64_t)(c + 2 * lanes) * stride), v3 = __ldcs(src + (int64_t)(c + 3 * lanes) * stride);
      s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
    }
    for (; c < spatial; c += lanes) s += (double)__ldcs(src + (int64_t)c * stride);
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  if (l == 0 && e < stride) {
    for (int k = 1; k < lanes; ++k) s += sm[k * epb + el];
    const int64_t i = map(group, e);
    if (i >= 0) out[i] = (float)s;
  }
}

template <typename MapF>
inline cudaError_t launch_partial_reduce(const float* partial, int spatial, int64_t stride, int groups, MapF map, float* out, cudaStream_t stream) {
  // few partials: one thread per output, consecutive threads read consecutive floats
  const int lanes = spatial >= 16 ? 8 : 1;
  const int epb = 256 / lanes;
  dim3 grid((unsigned)((stride + epb - 1) / epb), (unsigned)groups);
  partial_reduce_kernel<<<grid, 256, 0, stream>>>(partial, spatial, stride, lanes, map, out);
  return cudaGetLastError();
}
