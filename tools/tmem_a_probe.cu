// tmem_a_probe.cu — does feeding the UMMA A operand from TMEM (tcgen05.cp smem -> TMEM, then tcgen05.mma with A in TMEM)
// beat the shared-memory A operand for the small-N instructions of the 16-channel convolutions (M=128, N=48, K=16)?
// One issuing thread per CTA, one CTA per SM, R back-to-back instructions; reports cycles per MMA and checks that both
// paths produce the same accumulator.
#include "../multimodal_segmentation_project_b200/csrc/tc_ptx.cuh"
#include <cuda_bf16.h>
#include <stdio.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t desc_kmajor_sw32(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem), "r"(a_tmem),
               "l"(b_desc), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t dst_tmem, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(dst_tmem), "l"(sdesc) : "memory");
}

// mode 0: SS;  1: cp + TS every instruction;  2: TS only (A copied once: lower bound for the MMA itself);  3: cp only
__global__ void __launch_bounds__(128, 1) probe(int mode, int N, int R, long long* cycles, float* dump) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* A = smem + 1024;            // 4 A tiles of 128 rows x 32 B (SW32 K-major, 8-row groups 256 B apart)
  uint8_t* B = A + 4 * 4096;           // [2 k-chunks][N rows][16 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operands: small integers (exact in bf16); physical placement follows the layouts the conv kernels use
  for (int i = threadIdx.x; i < 4 * 128 * 16; i += blockDim.x) {
    const int tile = i / (128 * 16), r = (i / 16) % 128, k = i % 16;
    const float v = (float)(((r * 7 + k * 3 + tile) % 5) - 2);
    // SW32: 16-byte chunk index (k / 8) is XORed with bit 7 of the byte address (row parity of 4-row pairs)
    const uint32_t row_off = tile * 4096 + r * 32;
    const uint32_t chunk = (k / 8) ^ ((row_off >> 7) & 1);
    reinterpret_cast<__nv_bfloat16*>(A + row_off + chunk * 16)[k % 8] = __float2bfloat16(v);
  }
  for (int i = threadIdx.x; i < 2 * N * 8; i += blockDim.x) {
    const int kc = i / (N * 8), n = (i / 8) % N, j = i % 8;
    reinterpret_cast<__nv_bfloat16*>(B + (kc * N + n) * 16)[j] = __float2bfloat16((float)(((n * 5 + (kc * 8 + j)) % 7) - 3));
  }
  if (threadIdx.x == 0) { tc::mbar_init(tc::smem_u32(bar), 1); tc::fence_barrier_init(); }
  if (warp == 0) { tc::tmem_alloc(tc::smem_u32(tmem_slot), 256); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t idesc = tc::idesc_bf16_f32(128, N);
  const uint32_t a_tm[2] = {tmem + 192, tmem + 208};   // A staging: 8 columns each
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const uint64_t bdesc = tc::smem_desc_kmajor_noswz(tc::smem_u32(B), (uint32_t)N * 16, 128);
    if (mode == 2 && tc::elect_one()) tmem_cp_128x256b(a_tm[0], desc_kmajor_sw32(tc::smem_u32(A), 256));
    uint64_t ad[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ad[i] = desc_kmajor_sw32(tc::smem_u32(A + i * 4096), 256);
    const uint32_t d0 = tmem, d1 = tmem + 96;
    // prime both accumulators so that every timed instruction accumulates (no predicate games inside the loop)
    if (tc::elect_one()) {
      tc::umma_bf16_ss(d0, ad[0], bdesc, idesc, 0);
      tc::umma_bf16_ss(d1, ad[1], bdesc, idesc, 0);
    }
    t0 = clock64();
    for (int r = 0; r < R; r += 8) {
      if (tc::elect_one()) {   // one election per batch of 8 instructions: operands stay in uniform registers
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t d = (u & 1) ? d1 : d0;
          if (mode == 0) tc::umma_bf16_ss(d, ad[u & 3], bdesc, idesc, 1);
          else if (mode == 1) { tmem_cp_128x256b(a_tm[u & 1], ad[u & 3]); umma_bf16_ts(d, a_tm[u & 1], bdesc, idesc, 1); }
          else if (mode == 2) umma_bf16_ts(d, a_tm[0], bdesc, idesc, 1);
          else tmem_cp_128x256b(a_tm[u & 1], ad[u & 3]);
        }
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::umma_commit(tc::smem_u32(bar));
    tc::mbar_wait(tc::smem_u32(bar), 0);
    t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc::tc_fence_after();
  if (dump && blockIdx.x == 0) {   // accumulator 0, first 16 columns of every lane
    uint32_t v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    tc::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) dump[(warp * 32 + lane) * 16 + i] = __uint_as_float(v[i]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

int main() {
  long long* cyc; float* dump;
  CK(cudaMalloc(&cyc, 148 * 8)); CK(cudaMalloc(&dump, 128 * 16 * 4));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  std::vector<float> ref(128 * 16), got(128 * 16);
  for (int N : {16, 48, 96, 192}) {
    // correctness: one instruction, SS vs cp + TS
    probe<<<1, 128, 48 * 1024>>>(0, N, 1, cyc, dump); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ref.data(), dump, ref.size() * 4, cudaMemcpyDeviceToHost));
    probe<<<1, 128, 48 * 1024>>>(1, N, 1, cyc, dump); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(got.data(), dump, got.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; double s = 0;
    for (size_t i = 0; i < ref.size(); ++i) { bad += ref[i] != got[i]; s += fabs(ref[i]); }
    printf("N=%d: cp+TS vs SS accumulator: %d of %zu values differ (sum |ref| = %.0f)\n", N, bad, ref.size(), s);
    const char* names[4] = {"SS (A, B from smem)", "cp + TS per MMA    ", "TS only (A resident)", "cp only             "};
    for (int mode = 0; mode < 4; ++mode) {
      const int R = 4096;
      probe<<<148, 128, 48 * 1024>>>(mode, N, R, cyc, nullptr); CK(cudaDeviceSynchronize());
      long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
      double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
      printf("  N=%3d  %s: %6.1f cycles per instruction\n", N, names[mode], avg / R);
    }
  }
  return 0;
}
