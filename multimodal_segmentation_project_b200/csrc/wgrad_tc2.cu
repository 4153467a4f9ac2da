// wgrad_tc2.cu — second generation of the tcgen05 weight gradient of the 3x3x3 convolution.
//
//     dW[co][ci][kd][kh][kw] = sum_v  X[v + (kd-1, kh-1, kw-1)][ci] * dY[v][co]
//
// v1 (wgrad_tc.cu) issued one M=64 x N=48 MMA per (tile row r, kh) with only 16 of the 64 accumulator
// rows useful.  Here the roles are arranged so that BOTH fused tap dimensions carry useful work:
//
//   * A (M side) = one halo row rho of a 16-channel slab of X, the three kw taps being the descriptor's
//     leading-dimension stride (one voxel = 32 B): rows (kw, ci) = 48 useful of 64;
//   * B (N side) = the (up to) three tile rows rho-2 .. rho of a 16-channel slab of dY that pair with
//     that halo row through kh = 2, 1, 0 — again a leading-dimension stride (one tile row = 512 B):
//     columns (kh', co) = 48, all useful;
//   * D[kd] = 64 x 48 fp32 in TMEM per (dY slab, kd): 144 columns per dY slab.
//
// One MMA now covers what nine v1-MMA-rows' worth of useful MACs... precisely: 48x48x16 useful MACs per
// instruction instead of 16x48x16, and the instruction count drops from 144 to 54 per plane pair.
// Operands are MN-major (= NDHWC), SWIZZLE_32B rows of 32 bytes, staged by TMA tensor maps (halo and
// ragged tiles zero-filled by the TMA unit).  Per-CTA partials are reduced in fixed order (deterministic).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_maps.cuh"
#include "partial_reduce.cuh"
#include <stdlib.h>

namespace {

using bf16 = __nv_bfloat16;

constexpr int kQPitch = 24;                          // voxels per staged halo row (18 used)
constexpr int kQBytes = 18 * kQPitch * 32;           // 13824 = 54 swizzle periods
constexpr int kPSlabBytes = 256 * 32;                // 16x16 voxels x 16 ch
constexpr int kQStages = 3, kPStages = 4;            // a dY plane stays resident for three X planes (kd = 0,1,2)
constexpr int kThreads = 256;                        // w0: TMA, w1: MMA, w2: TMEM alloc, w4-7: epilogue
constexpr int kHeader = 256;
constexpr int kMaxDseg = 128;  // long d-runs: fewer split-K partials to reduce; ~2 CTA slots per SM are enough

struct Wg2Params {
  float* partial;               // [cta][pslab (<=2)][ci 16][27][co 16]
  int cp;                       // channels of dY
  int cq0, cq1;                 // channels of the two X tensors (virtual concat)
  int N, D, H, W;
  int pslabs;                   // dY slabs handled per CTA (1 or 2)
  int dseg, dblocks, tiles_w, tiles_h, tmem_cols;
};

__device__ __forceinline__ uint64_t desc_mn_sw32(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // version
  d |= (uint64_t)6 << 61;                       // SWIZZLE_32B (absolute-address swizzle; bases are 256-byte aligned)
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc2_kernel(const Wg2Params g, const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_q0,
                 const __grid_constant__ CUtensorMap tm_q1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = tc::smem_u32(bars);
  auto q_full = [&](int i) { return bar0 + 8u * i; };
  auto q_empty = [&](int i) { return bar0 + 8u * (3 + i); };
  auto p_full = [&](int i) { return bar0 + 8u * (6 + i); };
  auto p_empty = [&](int i) { return bar0 + 8u * (10 + i); };
  const uint32_t acc_done = bar0 + 8u * 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 192);
  const uint32_t p_stage_bytes = (uint32_t)g.pslabs * kPSlabBytes;
  uint8_t* pbuf = smem + kHeader;
  uint8_t* qbuf = pbuf + kPStages * p_stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = blockIdx.x % g.tiles_w, th = blockIdx.x / g.tiles_w % g.tiles_h;
  const int rest = blockIdx.x / (g.tiles_w * g.tiles_h);
  const int n = rest / g.dblocks, db = rest % g.dblocks;
  const int pgroup = blockIdx.y;      // group of `pslabs` dY slabs
  const int qslab = blockIdx.z;       // 16-channel slab of X
  const int w0 = tw * 16, h0 = th * 16, d0 = db * g.dseg;
  const int planes = min(g.dseg, g.D - d0);
  const int pslabs_here = min(g.pslabs, g.cp / 16 - pgroup * g.pslabs);

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kQStages; ++i) { tc::mbar_init(q_full(i), 1); tc::mbar_init(q_empty(i), 1); }
    for (int i = 0; i < kPStages; ++i) { tc::mbar_init(p_full(i), 1); tc::mbar_init(p_empty(i), 1); }
    tc::mbar_init(acc_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tc::smem_u32(tmem_slot), g.tmem_cols);
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
      const int qc = qslab * 16;
      const CUtensorMap* tq = qc < g.cq0 ? &tm_q0 : &tm_q1;
      const int qoff = qc < g.cq0 ? qc : qc - g.cq0;
      int pcount = 0, qcount = 0;
      for (int i = 0; i <= planes + 1; ++i) {
        if (i < planes) {  // dY plane i
          const int st = pcount % kPStages;
          tc::mbar_wait(p_empty(st), ((pcount / kPStages) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(p_full(st), (uint32_t)pslabs_here * kPSlabBytes);
          for (int ps = 0; ps < pslabs_here; ++ps)
            tma::load_5d(tc::smem_u32(pbuf + st * p_stage_bytes + ps * kPSlabBytes), &tm_p, (pgroup * g.pslabs + ps) * 16, w0, h0, d0 + i, n,
                         p_full(st));
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
      // M = 64 rows (kw, ci) of which 48 are real; N = 16 * (number of valid kh); A and B MN-major
      uint32_t idesc_n[4];
#pragma unroll
      for (int i = 1; i <= 3; ++i)
        idesc_n[i] = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((16 * i) >> 3) << 17) | ((64u >> 4) << 24);
      const uint64_t a_proto = desc_mn_sw32(0, 32, 256);      // M groups = kw taps, one voxel apart
      const uint64_t b_proto = desc_mn_sw32(0, 512, 256);     // N groups = tile rows, 16 voxels apart
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), a_lo0 = (uint32_t)a_proto;
      const uint32_t b_hi = (uint32_t)(b_proto >> 32), b_lo0 = (uint32_t)b_proto;
      uint32_t touched = 0;  // bit (ps*3 + kd): accumulator already holds data
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
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          const int pl = q - kd + 1;
          if (pl < 0 || pl >= planes) continue;
          for (int ps = 0; ps < pslabs_here; ++ps) {
            const uint32_t p_lo = b_lo0 + (tc::smem_u32(pbuf + (pl % kPStages) * p_stage_bytes + ps * kPSlabBytes) >> 4);
            const uint32_t d_tmem = tmem_base + (uint32_t)((ps * 3 + kd) * 48);
            const bool fresh = ((touched >> (ps * 3 + kd)) & 1u) == 0;
            // halo row rho pairs with tile rows rho-2+j (j = kh' = 2-kh): valid j in [max(0, 2-rho), min(2, 17-rho)]
            if (fresh) {
              // the first three halo rows open the three kh' column blocks one by one (accumulate = 0 overwrites)
              // rho = 0: j = 2 (row 0)
              tc::umma_bf16_ss(d_tmem + 32, ((uint64_t)a_hi << 32) | q_lo, ((uint64_t)b_hi << 32) | p_lo, idesc_n[1], 0);
              // rho = 1: j = 1 (row 0, fresh) and j = 2 (row 1)
              tc::umma_bf16_ss(d_tmem + 16, ((uint64_t)a_hi << 32) | (q_lo + 48), ((uint64_t)b_hi << 32) | p_lo, idesc_n[1], 0);
              tc::umma_bf16_ss(d_tmem + 32, ((uint64_t)a_hi << 32) | (q_lo + 48), ((uint64_t)b_hi << 32) | (p_lo + 32), idesc_n[1], 1);
              // rho = 2: j = 0 (row 0, fresh) and j = 1, 2 (rows 1, 2)
              tc::umma_bf16_ss(d_tmem, ((uint64_t)a_hi << 32) | (q_lo + 96), ((uint64_t)b_hi << 32) | p_lo, idesc_n[1], 0);
              tc::umma_bf16_ss(d_tmem + 16, ((uint64_t)a_hi << 32) | (q_lo + 96), ((uint64_t)b_hi << 32) | (p_lo + 32), idesc_n[2], 1);
            }
#pragma unroll
            for (int rho = 0; rho < 18; ++rho) {
              if (fresh && rho < 3) continue;
              const int j_lo = rho < 2 ? 2 - rho : 0, j_hi = rho > 15 ? 17 - rho : 2;
              const uint64_t adesc = ((uint64_t)a_hi << 32) | (q_lo + (uint32_t)(rho * (kQPitch * 32) >> 4));
              const uint64_t bdesc = ((uint64_t)b_hi << 32) | (p_lo + (uint32_t)((rho - 2 + j_lo) * 512 >> 4));
              tc::umma_bf16_ss(d_tmem + j_lo * 16, adesc, bdesc, idesc_n[j_hi - j_lo + 1], 1);
            }
          }
        }
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          const int pl = q - kd + 1;
          if (pl >= 0 && pl < planes) touched |= (1u << kd) | (1u << (3 + kd));
        }
        tc::umma_commit(q_empty(qst));
        if (q - 1 >= 0 && q - 1 < planes) tc::umma_commit(p_empty((q - 1) % kPStages));
      }
      tc::umma_commit(acc_done);
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> partial dW =====================
    // M = 64 accumulators: row m = kw*16 + ci lives in TMEM lane (m % 16) + 32 * (m / 16) => warp ew holds kw = ew
    const int ew = warp - 4;
    tc::mbar_wait(acc_done, 0);
    tc::tc_fence_after();
    const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    float* out = g.partial + cta * (int64_t)(g.pslabs * 16 * 27 * 16);
    for (int ps = 0; ps < pslabs_here; ++ps) {
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t r[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((ps * 3 + kd) * 48 + j * 16), r);
          tc::tmem_ld_wait();
          if (ew < 3 && lane < 16) {
            const int tap = kd * 9 + (2 - j) * 3 + ew;  // kh = 2 - j, kw = ew
            float4* dst = reinterpret_cast<float4*>(out + (((int64_t)ps * 16 + lane) * 27 + tap) * 16);
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
  if (warp == 2) tc::tmem_dealloc(tmem_base, g.tmem_cols);
}

// partial[(qslab, pgroup)][spatial][ps][ci 16][tap 27][co 16] -> dw[co][ci][27]
struct Wg2Map {
  int pgroups, pslabs, Cout, Cin;
  __device__ int64_t operator()(int group, int64_t e) const {
    const int qslab = group / pgroups, pgroup = group % pgroups;
    const int col = (int)(e % 16), tap = (int)((e / 16) % 27), cl = (int)((e / (16 * 27)) % 16), ps = (int)(e / (16 * 27 * 16));
    const int co = (pgroup * pslabs + ps) * 16 + col, ci = qslab * 16 + cl;
    if (co >= Cout || ci >= Cin) return -1;
    return ((int64_t)co * Cin + ci) * 27 + tap;
  }
};

struct Wg2Plan { int pslabs, pgroups, qslabs, dseg, dblocks, tiles_w, tiles_h, spatial, tmem_cols; size_t smem; };

Wg2Plan make_plan(int c0, int c1, int Cout, int N, int D, int H, int W) {
  Wg2Plan pl;
  const int Cin = c0 + c1;
  const int cps = Cout / 16;
  pl.pslabs = cps >= 2 ? 2 : 1;
  pl.pgroups = (cps + pl.pslabs - 1) / pl.pslabs;
  pl.qslabs = Cin / 16;
  pl.tmem_cols = pl.pslabs == 1 ? 256 : 512;  // 144 columns per dY slab
  pl.tiles_w = (W + 15) / 16;
  pl.tiles_h = (H + 15) / 16;
  const int64_t base = (int64_t)pl.tiles_w * pl.tiles_h * N * pl.pgroups * pl.qslabs;
  int dseg = kMaxDseg;
  while (dseg > 2 && base * ((D + dseg - 1) / dseg) < (int64_t)(1.7 * B200_NUM_SMS)) dseg >>= 1;
  if (dseg > D) dseg = D;
  pl.dseg = dseg;
  pl.dblocks = (D + dseg - 1) / dseg;
  pl.spatial = pl.tiles_w * pl.tiles_h * N * pl.dblocks;
  pl.smem = kHeader + (size_t)kPStages * pl.pslabs * kPSlabBytes + (size_t)kQStages * kQBytes + 1024;
  return pl;
}

}  // namespace

bool b200_conv3d_wgrad_tc2_supported(int c0, int c1, int Cout, int N, int D, int H, int W) {
  if (c0 <= 0 || c0 % 16 || c1 % 16 || Cout % 16) return false;
  return N > 0 && D > 0 && H > 0 && W > 0;
}

int64_t b200_conv3d_wgrad_tc2_workspace(int c0, int c1, int Cout, int N, int D, int H, int W) {
  const Wg2Plan pl = make_plan(c0, c1, Cout, N, D, H, W);
  return (int64_t)pl.spatial * pl.pgroups * pl.qslabs * pl.pslabs * 16 * 27 * 16 * 4;
}

int b200_conv3d_wgrad_tc2(const void* x0, int c0, const void* x1, int c1, const void* dy, int Cout, float* dw, void* workspace, int N, int D,
                          int H, int W, cudaStream_t stream) {
  B200_REQUIRE(b200_conv3d_wgrad_tc2_supported(c0, c1, Cout, N, D, H, W), B200_ERR_UNSUPPORTED, "conv3d_wgrad(tcgen05): unsupported channel counts");
  const Wg2Plan pl = make_plan(c0, c1, Cout, N, D, H, W);
  Wg2Params g;
  g.partial = (float*)workspace;
  g.cp = Cout; g.cq0 = c0; g.cq1 = c1;
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.pslabs = pl.pslabs; g.dseg = pl.dseg; g.dblocks = pl.dblocks; g.tiles_w = pl.tiles_w; g.tiles_h = pl.tiles_h; g.tmem_cols = pl.tmem_cols;
  CUtensorMap tm_p, tm_q0, tm_q1;
  int rc = tma::make_ndhwc_map(&tm_p, dy, Cout, N, D, H, W, 16, 16);
  if (rc) return rc;
  rc = tma::make_ndhwc_map(&tm_q0, x0, c0, N, D, H, W, kQPitch, 18);
  if (rc) return rc;
  if (c1) { rc = tma::make_ndhwc_map(&tm_q1, x1, c1, N, D, H, W, kQPitch, 18); if (rc) return rc; } else tm_q1 = tm_q0;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  B200_REQUIRE(pl.pgroups <= 65535 && pl.qslabs <= 65535, B200_ERR_UNSUPPORTED, "conv3d_wgrad(tcgen05): grid too large");
  dim3 grid((unsigned)pl.spatial, (unsigned)pl.pgroups, (unsigned)pl.qslabs);
  wgrad_tc2_kernel<<<grid, kThreads, pl.smem, stream>>>(g, tm_p, tm_q0, tm_q1);
  B200_CHECK_LAUNCH("conv3d_wgrad_tc2");
  launch_partial_reduce((const float*)workspace, pl.spatial, (int64_t)pl.pslabs * 16 * 27 * 16, pl.qslabs * pl.pgroups,
                        Wg2Map{pl.pgroups, pl.pslabs, Cout, c0 + c1}, dw, stream);
  B200_CHECK_LAUNCH("conv3d_wgrad_tc2_reduce");
  return B200_OK;
}
