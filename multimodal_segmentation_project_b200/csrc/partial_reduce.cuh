// partial_reduce.cuh — fixed-order (deterministic) reduction of per-CTA split-K partials.
//
// partial[group][cta < spatial][e < stride] -> out[map(group, e)].  A block covers 256 / LANES consecutive source elements
// times LANES partial-lanes: for a fixed lane the block's threads read consecutive floats (whole 32-byte sectors), lane l
// sums partials l, l + LANES, ... in fp64, and thread lane 0 adds the LANES lane sums in ascending order.  The previous
// one-warp-per-output kernels read one float per 32-byte sector (8x read amplification, 13 us per layer).
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
#pragma unroll 4
    for (int c = l; c < spatial; c += lanes) s += (double)src[(int64_t)c * stride];
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
  const int lanes = spatial >= 64 ? 32 : 1;  // few partials: one thread per output, consecutive threads read consecutive floats
  const int epb = 256 / lanes;
  dim3 grid((unsigned)((stride + epb - 1) / epb), (unsigned)groups);
  partial_reduce_kernel<<<grid, 256, 0, stream>>>(partial, spatial, stride, lanes, map, out);
  return cudaGetLastError();
}
