// wgrad_tc3.cu — weight gradient of the 3x3x3 convolution for layers with >= 32 channels on both sides:
// wgrad_tc2.cu's arrangement (kw taps on the M side, kh taps on the N side, kd = accumulator index) with operand rows
// of 32 channels (64 bytes, SWIZZLE_64B) instead of 16:
//
//   * A = one halo row of a 32-channel slab of X, kw = leading-dimension stride of one voxel (64 B): rows (kw, ci) = 96
//     useful of M = 128;
//   * B = the (up to) three tile rows of a 32-channel slab of dY that pair with it through kh (stride = one tile row of
//     16 voxels = 1024 B): columns (kh', co) = 96;
//   * D[kd] = 128 x 96 fp32 in TMEM (288 columns).
//
// One instruction covers 96 x 96 x 16 useful MACs (4x wgrad_tc2's 48 x 48 x 16) for about twice the cycles of the
// shared-memory operand fetch + math (T ~ A + B + N/2 wavefronts: 32 + 24 + 48 against 16 + 12 + 24).
// Per-CTA partials [ci 32][tap 27][co 32] are reduced in fixed order (partial_reduce.cuh).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_maps.cuh"
#include "partial_reduce.cuh"
#include <stdlib.h>

namespace {

using bf16 = __nv_bfloat16;

constexpr int kCW = 32;                              // channels per operand row
constexpr int kRowB = kCW * 2;                       // 64 bytes per voxel
constexpr int kQPitch = 24;                          // voxels per staged halo row (18 used)
constexpr int kQRowBytes = kQPitch * kRowB;          // 1536 = 3 swizzle periods of 512 B
constexpr int kQBytes = 18 * kQRowBytes;             // 27648
constexpr int kPRowBytes = 16 * kRowB;               // one tile row of dY: 1024
constexpr int kPBytes = 16 * kPRowBytes;             // 16384
constexpr int kQStages = 3, kPStages = 4;            // a dY plane stays resident for three X planes (kd = 0,1,2)
constexpr int kThreads = 256;                        // w0: TMA, w1: MMA, w2: TMEM alloc, w4-7: epilogue
constexpr int kHeader = 256;
constexpr int kMaxDseg = 128;
constexpr int kTmemCols = 512;                       // 3 kd x 96 columns used
constexpr int kPartialFloats = kCW * 27 * kCW;       // 27648 per CTA

struct Wg3Params {
  float* partial;
  int cq0;                      // channels of the first X tensor (virtual concat); multiples of 32
  int N, D, H, W;
  int dseg, dblocks, tiles_w, tiles_h;
};

__device__ __forceinline__ uint64_t desc_mn_sw64(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // version
  d |= (uint64_t)4 << 61;                       // SWIZZLE_64B (absolute-address swizzle; stage bases are 1024-byte aligned)
  return d;
}

#define B200_WG3_MMA(NAME, QUAL)                                                                                             \
  __device__ __forceinline__ void NAME(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {                                  \
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16" QUAL " [%0], %1, %2, %3, p;\n}" \
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");                                                           \
  }
B200_WG3_MMA(mma_fill, ".collector::a::fill")
B200_WG3_MMA(mma_use, ".collector::a::use")
B200_WG3_MMA(mma_last, ".collector::a::lastuse")

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc3_kernel(const Wg3Params g, const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_q0,
                 const __grid_constant__ CUtensorMap tm_q1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = tc::smem_u32(bars);
  auto q_full = [&](int i) { return bar0 + 8u * i; };
  auto q_empty = [&](int i) { return bar0 + 8u * (3 + i); };
  auto p_full = [&](int i) { return bar0 + 8u * (6 + i); };
  auto p_empty = [&](int i) { return bar0 + 8u * (10 + i); };
  const uint32_t acc_done = bar0 + 8u * 14, acc_zero = bar0 + 8u * 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 192);
  uint8_t* pbuf = smem + 1024;                   // stage bases stay 1024-byte aligned (kPBytes, kQBytes are multiples of 512)
  uint8_t* qbuf = pbuf + kPStages * kPBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = blockIdx.x % g.tiles_w, th = blockIdx.x / g.tiles_w % g.tiles_h;
  const int rest = blockIdx.x / (g.tiles_w * g.tiles_h);
  const int n = rest / g.dblocks, db = rest % g.dblocks;
  const int pslab = blockIdx.y;       // 32-channel slab of dY
  const int qslab = blockIdx.z;       // 32-channel slab of X
  const int w0 = tw * 16, h0 = th * 16, d0 = db * g.dseg;
  const int planes = min(g.dseg, g.D - d0);

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kQStages; ++i) { tc::mbar_init(q_full(i), 1); tc::mbar_init(q_empty(i), 1); }
    for (int i = 0; i < kPStages; ++i) { tc::mbar_init(p_full(i), 1); tc::mbar_init(p_empty(i), 1); }
    tc::mbar_init(acc_done, 1);
    tc::mbar_init(acc_zero, 4);
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tc::smem_u32(tmem_slot), kTmemCols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma::prefetch(&tm_p);
      tma::prefetch(&tm_q0);
      const int qc = qslab * kCW;
      const CUtensorMap* tq = qc < g.cq0 ? &tm_q0 : &tm_q1;
      const int qoff = qc < g.cq0 ? qc : qc - g.cq0;
      int pcount = 0, qcount = 0;
      for (int i = 0; i <= planes + 1; ++i) {
        if (i < planes) {  // dY plane i
          const int st = pcount % kPStages;
          tc::mbar_wait(p_empty(st), ((pcount / kPStages) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(p_full(st), kPBytes);
          tma::load_5d(tc::smem_u32(pbuf + st * kPBytes), &tm_p, pslab * kCW, w0, h0, d0 + i, n, p_full(st));
          ++pcount;
        }
        {  // X halo plane i - 1
          const int st = qcount % kQStages;
          tc::mbar_wait(q_empty(st), ((qcount / kQStages) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(q_full(st), kQBytes);
          tma::load_5d(tc::smem_u32(qbuf + st * kQBytes), tq, qoff, w0 - 1, h0 - 1, d0 + i - 1, n, q_full(st));
          ++qcount;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issue =====================
    if (lane == 0) {
      // M = 128 rows (kw, ci) of which 96 are real; N = 32 * (number of valid kh); A and B MN-major
      uint32_t idesc_n[4];
#pragma unroll
      for (int i = 1; i <= 3; ++i)
        idesc_n[i] = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((kCW * i) >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t a_proto = desc_mn_sw64(0, kRowB, 8 * kRowB);        // M groups = kw taps, one voxel apart; K groups of 8 voxels
      const uint64_t b_proto = desc_mn_sw64(0, kPRowBytes, 8 * kRowB);   // N groups = tile rows, 16 voxels apart
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), a_lo0 = (uint32_t)a_proto;
      const uint32_t b_hi = (uint32_t)(b_proto >> 32), b_lo0 = (uint32_t)b_proto;
      constexpr uint32_t kARow16 = kQRowBytes >> 4, kBRow16 = kPRowBytes >> 4;
      tc::mbar_wait(acc_zero, 0);      // the epilogue warps have zeroed the three accumulators: every instruction accumulates
      tc::tc_fence_after();
      int p_ready = 0;
      for (int qi = 0; qi <= planes + 1; ++qi) {
        const int q = qi - 1;
        const int qst = qi % kQStages;
        tc::mbar_wait(q_full(qst), (qi / kQStages) & 1);
        const int need = min(q + 2, planes);
        while (p_ready < need) {
          tc::mbar_wait(p_full(p_ready % kPStages), (p_ready / kPStages) & 1);
          ++p_ready;
        }
        tc::tc_fence_after();
        const uint32_t q_lo = a_lo0 + (tc::smem_u32(qbuf + qst * kQBytes) >> 4);
        // dY plane q - kd + 1 meets this X plane through kd; kd_lo .. kd_hi are the planes this CTA owns
        const int kd_lo = max(0, q + 2 - planes), kd_hi = min(2, q + 1);
        uint32_t p_lo[3];
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) p_lo[kd] = b_lo0 + (tc::smem_u32(pbuf + ((q - kd + 1 + kPStages) % kPStages) * kPBytes) >> 4);
        // halo row rho pairs with tile rows rho-2+j (j = kh' = 2-kh): valid j in [max(0, 2-rho), min(2, 17-rho)]
        if (kd_lo == 0 && kd_hi == 2) {
          // interior plane: the three kd instructions of a halo row share A through the collector (48 instead of 56 cycles at N = 96)
#pragma unroll
          for (int rho = 0; rho < 18; ++rho) {
            const int j_lo = rho < 2 ? 2 - rho : 0, j_hi = rho > 15 ? 17 - rho : 2;
            const uint64_t ad = ((uint64_t)a_hi << 32) | (q_lo + (uint32_t)rho * kARow16);
            const uint32_t boff = (uint32_t)(rho - 2 + j_lo) * kBRow16, col = tmem_base + (uint32_t)(j_lo * kCW);
            const uint32_t idesc = idesc_n[j_hi - j_lo + 1];
            mma_fill(col, ad, ((uint64_t)b_hi << 32) | (p_lo[0] + boff), idesc);
            mma_use(col + 3 * kCW, ad, ((uint64_t)b_hi << 32) | (p_lo[1] + boff), idesc);
            mma_last(col + 6 * kCW, ad, ((uint64_t)b_hi << 32) | (p_lo[2] + boff), idesc);
          }
        } else {
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            if (kd < kd_lo || kd > kd_hi) continue;
            const uint32_t d_tmem = tmem_base + (uint32_t)(kd * 3 * kCW);
#pragma unroll
            for (int rho = 0; rho < 18; ++rho) {
              const int j_lo = rho < 2 ? 2 - rho : 0, j_hi = rho > 15 ? 17 - rho : 2;
              const uint64_t ad = ((uint64_t)a_hi << 32) | (q_lo + (uint32_t)rho * kARow16);
              tc::umma_bf16_ss(d_tmem + j_lo * kCW, ad, ((uint64_t)b_hi << 32) | (p_lo[kd] + (uint32_t)(rho - 2 + j_lo) * kBRow16), idesc_n[j_hi - j_lo + 1], 1);
            }
          }
        }
        tc::umma_commit(q_empty(qst));
        if (q - 1 >= 0 && q - 1 < planes) tc::umma_commit(p_empty((q - 1) % kPStages));
      }
      tc::umma_commit(acc_done);
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> partial dW =====================
    // accumulator row m = kw*32 + ci is TMEM lane m: warp ew holds kw = ew, lane = ci
    const int ew = warp - 4;
#pragma unroll
    for (int c = 0; c < 9 * kCW; c += 16) tmem_st16_zero(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c);
    tmem_st_wait();
    tc::tc_fence_before();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(acc_zero);
    tc::mbar_wait(acc_done, 0);
    tc::tc_fence_after();
    const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    float* out = g.partial + cta * (int64_t)kPartialFloats;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(kd * 3 * kCW + j * kCW + half * 16), r);
          tc::tmem_ld_wait();
          if (ew < 3) {
            const int tap = kd * 9 + (2 - j) * 3 + ew;  // kh = 2 - j, kw = ew
            float4* dst = reinterpret_cast<float4*>(out + ((int64_t)lane * 27 + tap) * kCW + half * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, kTmemCols);
}

// partial[(qslab, pslab)][spatial][ci 32][tap 27][co 32] -> dw[co][ci][27]
struct Wg3Map {
  int pslabs, Cout, Cin;
  __device__ int64_t operator()(int group, int64_t e) const {
    const int qslab = group / pslabs, pslab = group % pslabs;
    const int col = (int)(e % kCW), tap = (int)((e / kCW) % 27), cl = (int)(e / (kCW * 27));
    const int co = pslab * kCW + col, ci = qslab * kCW + cl;
    if (co >= Cout || ci >= Cin) return -1;
    return ((int64_t)co * Cin + ci) * 27 + tap;
  }
};

struct Wg3Plan { int pslabs, qslabs, dseg, dblocks, tiles_w, tiles_h, spatial; size_t smem; };

Wg3Plan make_plan(int c0, int c1, int Cout, int N, int D, int H, int W) {
  Wg3Plan pl;
  pl.pslabs = Cout / kCW;
  pl.qslabs = (c0 + c1) / kCW;
  pl.tiles_w = (W + 15) / 16;
  pl.tiles_h = (H + 15) / 16;
  const int64_t base = (int64_t)pl.tiles_w * pl.tiles_h * N * pl.pslabs * pl.qslabs;
  // one CTA per SM (TMEM 512 columns): pick the d-run that minimises waves x (planes streamed + fixed per-CTA cost);
  // a run of dseg planes streams dseg + 2 halo planes, the epilogue (110 KB of partials) costs about 3 planes' worth
  int dseg = 1;
  int64_t best = -1;
  for (int cand = 1; cand <= D && cand <= kMaxDseg; ++cand) {
    const int64_t ctas = base * ((D + cand - 1) / cand);
    const int64_t waves = (ctas + B200_NUM_SMS - 1) / B200_NUM_SMS;
    const int64_t cost = waves * (cand + 2 + 3);
    if (best < 0 || cost < best || (cost == best && cand > dseg)) { best = cost; dseg = cand; }
  }
  pl.dseg = dseg;
  pl.dblocks = (D + dseg - 1) / dseg;
  pl.spatial = pl.tiles_w * pl.tiles_h * N * pl.dblocks;
  pl.smem = 1024 + (size_t)kPStages * kPBytes + (size_t)kQStages * kQBytes + 1024;
  return pl;
}

}  // namespace

bool b200_conv3d_wgrad_tc3_supported(int c0, int c1, int Cout, int N, int D, int H, int W) {
  if (c0 <= 0 || c0 % kCW || c1 % kCW || Cout % kCW) return false;
  return N > 0 && D > 0 && H > 0 && W > 0;
}

int64_t b200_conv3d_wgrad_tc3_workspace(int c0, int c1, int Cout, int N, int D, int H, int W) {
  const Wg3Plan pl = make_plan(c0, c1, Cout, N, D, H, W);
  return (int64_t)pl.spatial * pl.pslabs * pl.qslabs * kPartialFloats * 4;
}

int b200_conv3d_wgrad_tc3(const void* x0, int c0, const void* x1, int c1, const void* dy, int Cout, float* dw, void* workspace, int N, int D,
                          int H, int W, cudaStream_t stream) {
  B200_REQUIRE(b200_conv3d_wgrad_tc3_supported(c0, c1, Cout, N, D, H, W), B200_ERR_UNSUPPORTED, "conv3d_wgrad(tcgen05, wide rows): unsupported channel counts");
  const Wg3Plan pl = make_plan(c0, c1, Cout, N, D, H, W);
  Wg3Params g;
  g.partial = (float*)workspace;
  g.cq0 = c0;
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.dseg = pl.dseg; g.dblocks = pl.dblocks; g.tiles_w = pl.tiles_w; g.tiles_h = pl.tiles_h;
  CUtensorMap tm_p, tm_q0, tm_q1;
  int rc = tma::make_ndhwc_map_wide(&tm_p, dy, Cout, N, D, H, W, kCW, 16, 16);
  if (rc) return rc;
  rc = tma::make_ndhwc_map_wide(&tm_q0, x0, c0, N, D, H, W, kCW, kQPitch, 18);
  if (rc) return rc;
  if (c1) { rc = tma::make_ndhwc_map_wide(&tm_q1, x1, c1, N, D, H, W, kCW, kQPitch, 18); if (rc) return rc; } else tm_q1 = tm_q0;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(wgrad_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  B200_REQUIRE(pl.pslabs <= 65535 && pl.qslabs <= 65535, B200_ERR_UNSUPPORTED, "conv3d_wgrad(tcgen05, wide rows): grid too large");
  dim3 grid((unsigned)pl.spatial, (unsigned)pl.pslabs, (unsigned)pl.qslabs);
  wgrad_tc3_kernel<<<grid, kThreads, pl.smem, stream>>>(g, tm_p, tm_q0, tm_q1);
  B200_CHECK_LAUNCH("conv3d_wgrad_tc3");
  B200_CUDA(launch_partial_reduce((const float*)workspace, pl.spatial, (int64_t)kPartialFloats, pl.qslabs * pl.pslabs,
                                  Wg3Map{pl.pslabs, Cout, c0 + c1}, dw, stream));
  return B200_OK;
}
