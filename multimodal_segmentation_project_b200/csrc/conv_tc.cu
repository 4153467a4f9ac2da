// conv_tc.cu — tcgen05/TMEM implicit-GEMM 3x3x3 convolution (bf16, fp32 accumulate). Placeholder
// until the kernel lands: reports "unsupported" so that impl=auto uses the CUDA-core path.
#include "common.cuh"

bool b200_conv3d_k3_tc_supported(int, int, int, int, int, int, int, int) { return false; }
int64_t b200_pack_conv3_bytes_tc(int Cout, int Cin) { return (int64_t)27 * Cin * Cout * 2; }
int b200_pack_conv3_weights_tc(int, const float*, void*, int, int, cudaStream_t) {
  B200_FAIL(B200_ERR_UNSUPPORTED, "tcgen05 weight packing not built");
}
int b200_conv3d_k3_tc(const void*, int, const void*, int, const void*, const float*, void*, int, void*, int, int, int, int, int,
                      cudaStream_t) {
  B200_FAIL(B200_ERR_UNSUPPORTED, "tcgen05 conv not built");
}
