// tma_inner32_probe.cu — where does a TMA box whose inner extent (32 B) is narrower than its swizzle span (SWIZZLE_64B) land in
// shared memory?  Tensor [H][W][C = 32] bf16 read as {16 ch, slab, W, H}; box {16, 1, 8, 4}.  Every 16-byte chunk of the source is
// tagged, the whole shared buffer (pre-filled with a sentinel) is dumped and decoded on the host.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/tma_inner32_probe tools/tma_inner32_probe.cu
#include "../multimodal_segmentation_project_b200/csrc/tc_ptx.cuh"
#include "../multimodal_segmentation_project_b200/csrc/tma_maps.cuh"
#include <stdio.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int H = 4, W = 8, C = 32, kSmem = 4096;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int slab, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* buf = reinterpret_cast<uint32_t*>(smem + 1024);
  for (int i = threadIdx.x; i < kSmem / 4; i += blockDim.x) buf[i] = 0xFFFFFFFFu;
  if (threadIdx.x == 0) { tc::mbar_init(tc::smem_u32(bar), 1); tc::fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::mbar_arrive_expect_tx(tc::smem_u32(bar), 32 * W * H);
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(tc::smem_u32(buf)),
        "l"(reinterpret_cast<uint64_t>(&tm)), "r"(0), "r"(slab), "r"(0), "r"(0), "r"(tc::smem_u32(bar))
        : "memory");
  }
  tc::mbar_wait(tc::smem_u32(bar), 0);
  __syncthreads();
  for (int i = threadIdx.x; i < kSmem / 4; i += blockDim.x) out[i] = buf[i];
}

int main() {
  std::vector<uint16_t> h(H * W * C);
  // tag: every 16-byte chunk (8 elements) carries (h, w, chunk-of-voxel 0..3) in all its elements
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c) h[(y * W + x) * C + c] = (uint16_t)((y << 8) | (x << 4) | (c / 8));
  uint16_t* d; uint32_t* o;
  CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMalloc(&o, kSmem));
  CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  tma::EncodeTiledFn enc = tma::get_encode();
  if (!enc) { printf("no encode\n"); return 1; }
  for (int slab = 0; slab < 2; ++slab) {
    CUtensorMap tm;
    cuuint64_t gdim[4] = {16, C / 16, W, H};
    cuuint64_t gstr[3] = {32, C * 2, W * C * 2};
    cuuint32_t box[4] = {16, 1, W, H};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("slab %d: encode rc %d\n", slab, (int)r);
    if (r != CUDA_SUCCESS) continue;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem + 2048));
    probe<<<1, 128, kSmem + 2048>>>(tm, slab, o);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> res(kSmem / 4);
    CK(cudaMemcpy(res.data(), o, kSmem, cudaMemcpyDeviceToHost));
    int last = -1;
    for (int ch = 0; ch < kSmem / 16; ++ch) if (res[ch * 4] != 0xFFFFFFFFu) last = ch;
    printf("last written 16-byte chunk: %d (dense layout would end at %d)\n", last, 32 * W * H / 16 - 1);
    // expectation: dense [h][w][2 chunks] with chunk bits [4:5] ^= address bits [7:8]
    int bad = 0;
    for (int ch = 0; ch <= last && ch < kSmem / 16; ++ch) {
      const uint32_t v = res[ch * 4] & 0xFFFF;
      const int lin = ch ^ ((ch >> 3) & 3);                 // undo SWIZZLE_64B: 16-byte chunk index bits [0:1] ^= bits [3:4]
      const int eh = lin / (2 * W), ew = (lin / 2) % W, ec = slab * 2 + (lin & 1);
      const uint32_t expect = (uint32_t)((eh << 8) | (ew << 4) | ec);
      if (v != expect) ++bad;
      if (ch < 24) printf("chunk %2d: h %d w %d c %d   (dense+swizzle expects h %d w %d c %d)%s\n", ch, (v >> 8) & 0xF, (v >> 4) & 0xF, v & 0xF, eh, ew, ec,
                          res[ch * 4] == 0xFFFFFFFFu ? "  UNWRITTEN" : "");
    }
    printf("slab %d: %d chunks differ from the dense + address-swizzle expectation\n", slab, bad);
  }
  return 0;
}
