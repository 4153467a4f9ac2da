// pdl_probe.cu — what does a kernel boundary cost inside a CUDA graph on B200, and what does programmatic dependent launch
// (griddepcontrol) buy?  A chain of N dependent small kernels (each CTA touches 4 KB) is captured and replayed.
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <bool PDL>
__global__ void __launch_bounds__(256) step_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
  if (PDL) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i] + 1.f;
}

template <bool PDL>
static cudaError_t launch(cudaStream_t st, int grid, const float* in, float* out, int n, bool attr) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute a[1];
  a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = a; cfg.numAttrs = attr ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, step_kernel<PDL>, in, out, n);
}

int main() {
  const int N = 1000;
  cudaStream_t st; CK(cudaStreamCreate(&st));
  for (int grid : {1, 148, 592, 2368}) {
    const int n = grid * 256 * 4;
    float *a, *b; CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMemset(a, 0, n * 4));
    for (int mode = 0; mode < 2; ++mode) {
      cudaGraph_t g; cudaGraphExec_t ge;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      for (int i = 0; i < N; ++i) {
        const float* src = (i & 1) ? b : a; float* dst = (i & 1) ? a : b;
        CK(mode ? launch<true>(st, grid, src, dst, n, true) : launch<false>(st, grid, src, dst, n, false));
      }
      CK(cudaStreamEndCapture(st, &g));
      CK(cudaGraphInstantiate(&ge, g, 0));
      CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      CK(cudaEventRecord(e0, st)); CK(cudaGraphLaunch(ge, st)); CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("grid %5d  %s: %.2f us per dependent kernel\n", grid, mode ? "PDL (wait + launch_dependents)" : "plain stream order           ", ms * 1e3 / N);
      cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    }
    float h; CK(cudaMemcpy(&h, a, 4, cudaMemcpyDeviceToHost));
    printf("  check value %.0f (expect %d)\n", h, 2 * N + 0 * (int)h);
    cudaFree(a); cudaFree(b);
  }
  return 0;
}
