// ts_pipeline_probe.cu — round 2: can the A operand of the small-N convolution MMAs (M=128, K=16, N=48/96) be fed from
// TENSOR MEMORY by producer warps (ld.shared -> tcgen05.st) fast enough to take the shared-memory A fetch off the
// tensor core's critical path?  Round 1 measured tcgen05.cp smem->TMEM as no gain (64 B/clk, same wavefronts as the SS fetch).
// Here the A blocks are written with tcgen05.st (register -> TMEM) and each block is consumed by U MMAs (in the conv
// kernel a block (row h', kw) serves the three kh taps, so U = 3).
//   mode 0: SS   (A by descriptor from shared memory, start address shifted per window: today's kernels)
//   mode 1: TS, A resident (lower bound of the MMA itself)
//   mode 2: TS fed by 4 producer warps through a ring of A blocks (the candidate pipeline)
//   mode 3: producers only (ld.shared + tcgen05.st, no MMA)
// Prints cycles per MMA and checks mode 2 against mode 0 on the accumulator.
#include "../multimodal_segmentation_project_b200/csrc/tc_ptx.cuh"
#include <cuda_bf16.h>
#include <math.h>
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
// collector usage for the A operand: 0 default, 1 fill (load A and keep it), 2 use (reuse the kept A and keep it), 3 lastuse
#define B200_MMA_COLL(NAME, QUAL)                                                                                                 \
  __device__ __forceinline__ void NAME##_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {                                  \
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16" QUAL " [%0], [%1], %2, %3, p;\n}" ::"r"(d), \
                 "r"(a), "l"(b), "r"(idesc) : "memory");                                                                           \
  }                                                                                                                                \
  __device__ __forceinline__ void NAME##_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {                                  \
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16" QUAL " [%0], %1, %2, %3, p;\n}" ::"r"(d),   \
                 "l"(a), "l"(b), "r"(idesc) : "memory");                                                                           \
  }
B200_MMA_COLL(mma_fill, ".collector::a::fill")
B200_MMA_COLL(mma_use, ".collector::a::use")
B200_MMA_COLL(mma_last, ".collector::a::lastuse")

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kRing = 5;       // ring of A units; unit = 3 blocks (kw) of 8 TMEM columns
constexpr int kRows = 4;       // source rows in shared memory
constexpr int kRowVox = 136;   // 130 used, padded to a multiple of 8 voxels (256-byte swizzle period)

// unit i: source row i % kRows; its three blocks are the voxel shifts kw = 0, 1, 2; each block feeds 3 MMAs (kh = 0, 1, 2:
// three accumulators, three weight taps) -> 9 MMAs per unit, as in the candidate conv kernel.
// PW = producer warps (4 or 8: with 8, two warps share a TMEM lane quadrant and alternate units)
template <int MODE_, bool COLL>
__global__ void __launch_bounds__(288, 1) probe(int N, int PW, int R, long long* cycles, float* dump, int ND = 3) {
  constexpr int mode = MODE_;
  constexpr bool coll = COLL;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);            // [0,8) a_full, [8,16) a_empty, 16: done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
  uint8_t* A = smem + 1024;                       // kRows x kRowVox x 32 B, SWIZZLE_32B
  uint8_t* B = A + kRows * kRowVox * 32;          // 9 taps x [2 k-chunks][N rows][16 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = tc::smem_u32(bars);
  for (int i = threadIdx.x; i < kRows * kRowVox * 16; i += blockDim.x) {
    const int row = i / (kRowVox * 16), v = (i / 16) % kRowVox, k = i % 16;
    const float val = (float)(((v * 7 + k * 3 + row * 5) % 5) - 2);
    const uint32_t off = (uint32_t)(row * kRowVox + v) * 32;
    const uint32_t chunk = (k / 8) ^ ((off >> 7) & 1);
    reinterpret_cast<__nv_bfloat16*>(A + off + chunk * 16)[k % 8] = __float2bfloat16(val);
  }
  for (int i = threadIdx.x; i < 9 * 2 * N * 8; i += blockDim.x) {
    const int tap = i / (2 * N * 8), kc = (i / (N * 8)) % 2, n = (i / 8) % N, j = i % 8;
    reinterpret_cast<__nv_bfloat16*>(B + ((tap * 2 + kc) * N + n) * 16)[j] = __float2bfloat16((float)(((n * 5 + (kc * 8 + j) + tap) % 7) - 3));
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { tc::mbar_init(bar0 + 8 * i, 4); tc::mbar_init(bar0 + 8 * (8 + i), 1); }
    tc::mbar_init(bar0 + 8 * 16, 1);
    tc::fence_barrier_init();
  }
  if (warp == 8) { tc::tmem_alloc(tc::smem_u32(tmem_slot), 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t idesc = tc::idesc_bf16_f32(128, N);
  const uint32_t a_ring = tmem + 384;             // kRing x 24 columns
  const int nunits = R;
  long long t0 = clock64(), t1 = t0;

  const int spin_warps = PW / 100, spin_kind = (PW / 10) % 10;
  PW %= 10;
  if (warp < spin_warps && mode <= 1) {
    // idle roles of a warp-specialised kernel waiting for their barrier: 0 = try_wait spin (what tc2/tc3 do),
    // 1 = try_wait with a suspend-time hint, 2 = nanosleep back-off between polls
    const uint32_t bar = bar0 + 8 * 16;
    if (spin_kind == 0) {
      tc::mbar_wait(bar, 0);
    } else if (spin_kind == 1) {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(0u), "r"(1000000u) : "memory");
      }
    } else {
      while (!tc::mbar_try_wait(bar, 0)) __nanosleep(200);
    }
  } else if (warp < PW && (mode == 2 || mode == 3)) {
    // ---- producers: thread = TMEM lane = voxel of the M tile
    const int q = warp & 3, sub = warp >> 2, nsub = PW >> 2;
    const int m = q * 32 + lane;
    for (int i = sub; i < nunits; i += nsub) {
      const int slot = i % kRing;
      if (mode == 2) tc::mbar_wait(bar0 + 8 * (8 + slot), ((i / kRing) & 1) ^ 1);
      tc::tc_fence_after();
      uint32_t r[3][8];
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const uint32_t off = (uint32_t)((i % kRows) * kRowVox + m + kw) * 32;
        const uint32_t sw = ((off >> 7) & 1) * 16;                    // logical chunk 0 lives at off + sw, chunk 1 at off + (16 - sw)
        const uint4 lo = *reinterpret_cast<const uint4*>(A + off + sw);
        const uint4 hi = *reinterpret_cast<const uint4*>(A + off + (16 - sw));
        r[kw][0] = lo.x; r[kw][1] = lo.y; r[kw][2] = lo.z; r[kw][3] = lo.w; r[kw][4] = hi.x; r[kw][5] = hi.y; r[kw][6] = hi.z; r[kw][7] = hi.w;
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) tmem_st8(a_ring + ((uint32_t)(q * 32) << 16) + slot * 24 + kw * 8, r[kw]);
      tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (mode == 2 && lane == 0) tc::mbar_arrive(bar0 + 8 * slot);
    }
    t1 = clock64();
  } else if (warp == 8) {
    const uint64_t b_proto = tc::smem_desc_kmajor_noswz(tc::smem_u32(B), (uint32_t)N * 16, 128);
    const uint32_t b_hi = (uint32_t)(b_proto >> 32), b_lo = (uint32_t)b_proto, b_tap = (uint32_t)(2 * N * 16) >> 4;
    const uint64_t a_proto = desc_kmajor_sw32(tc::smem_u32(A), 256);
    const uint32_t a_hi = (uint32_t)(a_proto >> 32), a_lo = (uint32_t)a_proto;
    const uint32_t dstep = (uint32_t)N * 6 <= 384 ? (uint32_t)N : (N <= 96 ? (uint32_t)N : 40u);   // accumulators; overlapping for large N (timing only)
    uint32_t dsel[9];
    for (int t = 0; t < 9; ++t) dsel[t] = tmem + (uint32_t)(t % (ND % 100)) * dstep;
    if (tc::elect_one()) {
      for (int kh = 0; kh < 3; ++kh) tc::umma_bf16_ss(tmem + kh * dstep, a_proto, b_proto, idesc, 0);   // prime
    }
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < nunits; ++i) {
      const int slot = i % kRing;
      if (mode == 2) { tc::mbar_wait(bar0 + 8 * slot, (i / kRing) & 1); tc::tc_fence_after(); }
      const bool halo = ND >= 100;      // tc2/tc3 operand layout: M = 16 rows x 8 voxels of an 18 x 18 halo plane (8-row groups 576 B apart)
      const uint32_t a_row = halo ? (uint32_t)(((tc::smem_u32(A) + (i & 1) * 256) >> 4) & 0x3FFF) | (1u << 16)
                                  : a_lo + (uint32_t)(((i % kRows) * kRowVox * 32) >> 4);
      const uint32_t a_t = a_ring + slot * 24;
      if (tc::elect_one()) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)(kh * 3 + kw) * b_tap);
            const uint64_t adesc = halo ? ((((uint64_t)((576u >> 4) & 0x3FFF) | (1ull << 14) | (6ull << 29)) << 32) | (a_row + (uint32_t)((kh * 18 + kw) * 2)))
                                        : (((uint64_t)a_hi << 32) | (a_row + 2u * kw));
            if (!coll) {
              if (mode == 0) tc::umma_bf16_ss(dsel[kw * 3 + kh], adesc, bd, idesc, 1);
              else umma_bf16_ts(dsel[kw * 3 + kh], a_t + kw * 8, bd, idesc, 1);
            } else if (mode == 0) {
              if (kh == 0) mma_fill_ss(dsel[kw * 3 + kh], adesc, bd, idesc);
              else if (kh == 1) mma_use_ss(dsel[kw * 3 + kh], adesc, bd, idesc);
              else mma_last_ss(dsel[kw * 3 + kh], adesc, bd, idesc);
            } else {
              if (kh == 0) mma_fill_ts(dsel[kw * 3 + kh], a_t + kw * 8, bd, idesc);
              else if (kh == 1) mma_use_ts(dsel[kw * 3 + kh], a_t + kw * 8, bd, idesc);
              else mma_last_ts(dsel[kw * 3 + kh], a_t + kw * 8, bd, idesc);
            }
          }
        }
        if (mode == 2) tc::umma_commit(bar0 + 8 * (8 + slot));
      }
      __syncwarp();
    }
    if (mode != 3) {
      if (tc::elect_one()) tc::umma_commit(bar0 + 8 * 16);
      __syncwarp();
      tc::mbar_wait(bar0 + 8 * 16, 0);
    }
    t1 = clock64();
  }
  if (lane == 0 && ((mode == 3 && warp == 0) || (mode != 3 && warp == 8))) cycles[blockIdx.x] = t1 - t0;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (dump && blockIdx.x == 0 && warp < 4) {   // accumulator 0, first 16 columns of every lane
    uint32_t v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    tc::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) dump[(warp * 32 + lane) * 16 + i] = __uint_as_float(v[i]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

static void launch(int m, int grid, size_t smem, int N, int PW, int R, long long* cyc, float* dump, int ND = 3) {
  switch (m) {
    case 0: probe<0, false><<<grid, 288, smem>>>(N, PW, R, cyc, dump, ND); break;
    case 1: probe<1, false><<<grid, 288, smem>>>(N, PW, R, cyc, dump, ND); break;
    case 2: probe<2, false><<<grid, 288, smem>>>(N, PW, R, cyc, dump, ND); break;
    case 3: probe<3, false><<<grid, 288, smem>>>(N, PW, R, cyc, dump, ND); break;
    case 4: probe<0, true><<<grid, 288, smem>>>(N, PW, R, cyc, dump, ND); break;
    case 5: probe<1, true><<<grid, 288, smem>>>(N, PW, R, cyc, dump, ND); break;
    case 6: probe<2, true><<<grid, 288, smem>>>(N, PW, R, cyc, dump, ND); break;
  }
}

int main() {
  long long* cyc; float* dump;
  CK(cudaMalloc(&cyc, 148 * 8)); CK(cudaMalloc(&dump, 128 * 16 * 4));
  CK(cudaFuncSetAttribute(probe<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(probe<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(probe<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(probe<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(probe<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(probe<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(probe<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  std::vector<float> ref(128 * 16), got(128 * 16);
  const size_t smem = 96 * 1024;
  for (int N : {16, 48, 96, 144, 256}) {
    const int Rc = 24;   // correctness run
    launch(0, 1, smem, N, 4, Rc, cyc, dump); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ref.data(), dump, ref.size() * 4, cudaMemcpyDeviceToHost));
    for (int m : {2, 4, 6}) {
      launch(m, 1, smem, N, 4, Rc, cyc, dump); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(got.data(), dump, got.size() * 4, cudaMemcpyDeviceToHost));
      int bad = 0; double s = 0;
      for (size_t i = 0; i < ref.size(); ++i) { bad += ref[i] != got[i]; s += fabs(ref[i]); }
      printf("N=%d mode %d vs plain SS accumulator: %d of %zu values differ (sum |ref| = %.0f)\n", N, m, bad, ref.size(), s);
    }
    const char* names[7] = {"SS (A, B from smem)           ", "TS, A resident                ", "TS fed by ld.shared+tcgen05.st", "producers only                ",
                            "SS + collector::a reuse x3    ", "TS resident + collector reuse ", "TS fed + collector reuse      "};
    for (int mode = 0; mode < 7; ++mode) {
      for (int PW : {4, 8}) {
        if (mode != 2 && mode != 6 && PW == 8) continue;
        const int R = 4000;
        launch(mode, 148, smem, N, PW, R, cyc, nullptr); CK(cudaDeviceSynchronize());
        long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
        printf("  N=%3d %s PW=%d: %7.1f cycles per MMA (floor N/2 = %d), %7.1f per unit of 9\n", N, names[mode], PW, avg / R / 9, N / 2, avg / R);
      }
    }
  }
  printf("idle warps waiting on an mbarrier while one warp issues SS MMAs (N = 48), cycles per MMA\n");
  for (int kind = 0; kind < 3; ++kind) {
    for (int sw : {0, 1, 4, 8}) {
      const int R = 4000;
      launch(0, 148, smem, 48, sw * 100 + kind * 10 + 4, R, cyc, nullptr, 2); CK(cudaDeviceSynchronize());
      long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
      double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
      printf("  wait kind %d (0 try_wait spin, 1 try_wait + time hint, 2 nanosleep back-off), %d waiting warps: %6.1f\n", kind, sw, avg / R / 9);
    }
  }
  printf("operand layout: dense rows (SBO 256) vs 18x18 halo-plane windows (SBO 576, tc2/tc3), SS mode, cycles per MMA\n");
  for (int N : {16, 48, 96}) {
    double r[2];
    for (int k = 0; k < 2; ++k) {
      const int R = 4000;
      launch(0, 148, smem, N, 4, R, cyc, nullptr, k == 0 ? 2 : 102); CK(cudaDeviceSynchronize());
      long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
      double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
      r[k] = avg / R / 9;
    }
    printf("  N=%3d: dense %6.1f   halo-plane windows %6.1f\n", N, r[0], r[1]);
  }
  printf("accumulator-count sweep (SS mode / SS + collector reuse), cycles per MMA\n");
  for (int N : {16, 48, 64}) {
    for (int ND : {1, 2, 3, 4, 6}) {
      double r[2];
      for (int k = 0; k < 2; ++k) {
        const int R = 4000;
        launch(k == 0 ? 0 : 4, 148, smem, N, 4, R, cyc, nullptr, ND); CK(cudaDeviceSynchronize());
        long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
        r[k] = avg / R / 9;
      }
      printf("  N=%3d distinct accumulators %d: SS %6.1f   SS+collector %6.1f\n", N, ND, r[0], r[1]);
    }
  }
  return 0;
}
