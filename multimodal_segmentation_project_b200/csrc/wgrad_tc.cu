// wgrad_tc.cu — tcgen05 / TMEM weight gradient of the 3x3x3 convolution (bf16 in, fp32 out).
//
// Replaces cuDNN's bwd-filter kernel behind nn.Conv3d(k=3,p=1) (models/unet.py:11,15):
//     dW[co][ci][kd][kh][kw] = sum_v  X[v + (kd-1, kh-1, kw-1)][ci] * dY[v][co]
// The reduction runs over voxels, so voxels are the UMMA K dimension and BOTH operands are
// "MN-major" (channels contiguous per voxel), which is exactly the NDHWC layout:
//
//   * P  = the tensor on the M side (<= 64 channels per CTA, no halo), staged per d-plane as
//          [16-ch slab][16x16 voxels][16 ch]  (32-byte rows, SWIZZLE_32B)
//   * Q  = the tensor on the N side (one 16-channel slab per CTA, 18x18 halo plane, row pitch 24
//          voxels so every row starts on the 256-byte swizzle period), [voxel][16 ch], SWIZZLE_32B
//   * one tcgen05.mma (M=64, N=48, K=16) per (row r of 16 voxels, kh): A = P[row r], B = Q[row r+kh]
//     with the three kw taps expressed as the descriptor's leading-dimension stride (LBO = one
//     voxel = 32 B), D[(kd,kh)] = 64 x (kw, 16 ch) fp32 in TMEM: 9 accumulators x 48 columns.
//   * whichever of X / dY has more channels is put on the M side (it is the padded dimension);
//     if that is X the result is the gradient of the mirrored tap, undone in the reduction.
//   * every CTA owns 16x16xDSEG voxels x one Q slab x one 64-channel P chunk and writes its partial
//     dW to a workspace; a second kernel reduces the partials in fixed order (deterministic).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_maps.cuh"
#include "partial_reduce.cuh"
#include <stdlib.h>

namespace {

using bf16 = __nv_bfloat16;

constexpr int kQPitch = 24;                          // voxels per halo row (18 used), 768 B = 3 swizzle periods
constexpr int kQBytes = 18 * kQPitch * 32;           // 13824
constexpr int kPSlabBytes = 256 * 32;                // one 16-channel slab of a 16x16 plane
constexpr int kQStages = 3, kPStages = 4;   // P planes stay resident for three Q planes (kd = 0,1,2)
constexpr int kThreads = 288;                        // 4 producer + 4 epilogue + 1 MMA warps
constexpr int kHeader = 256;
constexpr int kAccCols = 9 * 48;                     // 432 -> 512 allocated
constexpr int kMaxDseg = 128;  // long d-runs: fewer split-K partials to reduce; ~2 CTA slots per SM are enough

struct WgParams {
  const bf16* p; int cp;        // M-side tensor [N,D,H,W,cp]
  const bf16* q0; const bf16* q1; int cq0, cq1;  // N-side tensor = virtual concat [q0 | q1]
  float* partial;               // [cta][mrows][27][16]
  int mrows;                    // min(64, cp): rows stored per CTA
  int N, D, H, W;
  int mslabs;                   // 16-channel slabs of P handled per CTA (1..4)
  int dseg, dblocks, tiles_w, tiles_h;
  int m128;                     // issue M=128 MMAs (rows beyond the real channels are ignored)
  int repeat;                   // diagnostics only (B200_WG_REPEAT)
};

__device__ __forceinline__ uint32_t swz32(uint32_t off) { return off ^ (((off >> 7) & 1u) << 4); }

__device__ __forceinline__ uint64_t desc_mn_sw32(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // version
  d |= (uint64_t)((addr >> 7) & 0x7) << 49;     // base offset (0 for 256-byte aligned starts)
  d |= (uint64_t)6 << 61;                       // SWIZZLE_32B
  return d;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const WgParams g, const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_q0,
                const __grid_constant__ CUtensorMap tm_q1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_32B patterns repeat every 256 bytes of ABSOLUTE shared-memory address: align the carve-up ourselves
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = tc::smem_u32(bars);
  auto q_full = [&](int i) { return bar0 + 8u * i; };
  auto q_empty = [&](int i) { return bar0 + 8u * (3 + i); };
  auto p_full = [&](int i) { return bar0 + 8u * (6 + i); };
  auto p_empty = [&](int i) { return bar0 + 8u * (10 + i); };
  const uint32_t acc_done = bar0 + 8u * 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 192);
  uint8_t* pbuf = smem + kHeader;                                   // kPStages x mslabs x 8192
  const uint32_t p_stage_bytes = (uint32_t)g.mslabs * kPSlabBytes;
  uint8_t* qbuf = pbuf + kPStages * p_stage_bytes;                  // kQStages x 13824

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = blockIdx.x % g.tiles_w, th = blockIdx.x / g.tiles_w % g.tiles_h;
  const int rest = blockIdx.x / (g.tiles_w * g.tiles_h);
  const int n = rest / g.dblocks, db = rest % g.dblocks;
  const int mchunk = blockIdx.y;      // 64-channel chunk of P
  const int qslab = blockIdx.z;       // 16-channel slab of Q
  const int w0 = tw * 16, h0 = th * 16, d0 = db * g.dseg;
  const int planes = min(g.dseg, g.D - d0);

  if (warp == 8 && lane == 0) {
    for (int i = 0; i < kQStages; ++i) { tc::mbar_init(q_full(i), 1); tc::mbar_init(q_empty(i), 1); }
    for (int i = 0; i < kPStages; ++i) { tc::mbar_init(p_full(i), 1); tc::mbar_init(p_empty(i), 1); }
    tc::mbar_init(acc_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 4) {
    tc::tmem_alloc(tc::smem_u32(tmem_slot), 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer: one thread drives the TMA unit =====================
    // P plane i  : mslabs boxes of 16 ch x 16 (w) x 16 (h)  -> [slab][voxel][16 ch]
    // Q plane i-1: one box of 16 ch x 24 (w) x 18 (h) at (w0-1, h0-1) -> halo rows of 24 voxels (18 used)
    // out-of-volume voxels / channels are zero-filled by the TMA unit (conv padding, ragged tiles)
    if (lane == 0) {
      tma::prefetch(&tm_p);
      tma::prefetch(&tm_q0);
      const int cpo = mchunk * 64;
      const int qc = qslab * 16;
      const CUtensorMap* tq = qc < g.cq0 ? &tm_q0 : &tm_q1;
      const int qoff = qc < g.cq0 ? qc : qc - g.cq0;
      int pcount = 0, qcount = 0;
      for (int i = 0; i <= planes + 1; ++i) {
        if (i < planes) {
          const int st = pcount % kPStages;
          tc::mbar_wait(p_empty(st), ((pcount / kPStages) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(p_full(st), p_stage_bytes);
          const uint32_t dst = tc::smem_u32(pbuf + st * p_stage_bytes);
          for (int ms = 0; ms < g.mslabs; ++ms)
            tma::load_5d(dst + ms * kPSlabBytes, &tm_p, cpo + ms * 16, w0, h0, d0 + i, n, p_full(st));
          ++pcount;
        }
        {
          const int st = qcount % kQStages;
          tc::mbar_wait(q_empty(st), ((qcount / kQStages) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(q_full(st), kQBytes);
          tma::load_5d(tc::smem_u32(qbuf + st * kQBytes), tq, qoff, w0 - 1, h0 - 1, d0 + i - 1, n, q_full(st));
          ++qcount;
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issue =====================
    if (lane == 0) {
      // M = 64, N = 48, A and B MN-major
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((48u >> 3) << 17) | ((g.m128 ? (128u >> 4) : (64u >> 4)) << 24);
      const uint32_t a_lbo = g.mslabs > 1 ? kPSlabBytes : 0;   // M-group (16 ch) stride; rows beyond the real channels are ignored
      // all stage bases are 256-byte aligned, so the descriptor base-offset field stays 0
      const uint64_t a_proto = desc_mn_sw32(0, a_lbo, 256), b_proto = desc_mn_sw32(0, 32, 256);
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), a_lo0 = (uint32_t)a_proto;
      const uint32_t b_hi = (uint32_t)(b_proto >> 32), b_lo0 = (uint32_t)b_proto;
      uint32_t touched = 0;
      int p_ready = 0;  // number of P planes whose full-barrier has been observed
      for (int qi = 0; qi <= planes + 1; ++qi) {
        const int q = qi - 1;
        const int qst = qi % kQStages;
        tc::mbar_wait(q_full(qst), (qi / kQStages) & 1);
        const int need = min(q + 2, planes);  // planes 0 .. q+1 must have landed
        while (p_ready < need) {
          tc::mbar_wait(p_full(p_ready % kPStages), (p_ready / kPStages) & 1);
          ++p_ready;
        }
        tc::tc_fence_after();
        const uint32_t q_base = tc::smem_u32(qbuf + qst * kQBytes);
        // single in-order issuing lane: descriptors = constant high word + low word advanced by constants
        const uint32_t q_lo = b_lo0 + (q_base >> 4);
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          const int pl = q - kd + 1;
          if (pl < 0 || pl >= planes) continue;
          const uint32_t p_lo = a_lo0 + (tc::smem_u32(pbuf + (pl % kPStages) * p_stage_bytes) >> 4);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((kd * 3 + kh) * 48);
            const uint32_t acc0 = (touched >> (kd * 3 + kh)) & 1u;
            for (int rep = 0; rep < g.repeat; ++rep)
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              const uint64_t adesc = ((uint64_t)a_hi << 32) | (p_lo + (uint32_t)(r * 512 >> 4));
              const uint64_t bdesc = ((uint64_t)b_hi << 32) | (q_lo + (uint32_t)((r + kh) * (kQPitch * 32) >> 4));
              tc::umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (r == 0 && rep == 0) ? acc0 : 1u);
            }
            touched |= 1u << (kd * 3 + kh);
          }
        }
        tc::umma_commit(q_empty(qst));
        if (q - 1 >= 0 && q - 1 < planes) tc::umma_commit(p_empty((q - 1) % kPStages));  // plane q-1 was last used here (kd = 2)
      }
      tc::umma_commit(acc_done);
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> partial dW =====================
    const int ew = warp - 4;
    tc::mbar_wait(acc_done, 0);
    tc::tc_fence_after();
    const int m_real = min(64, g.cp - mchunk * 64);
    // M=64 accumulators: row m lives in TMEM lane (m % 16) + 32 * (m / 16); M=128: row m = lane m
    const int co = g.m128 ? ew * 32 + lane : ew * 16 + lane;
    const bool row_ok = (g.m128 || lane < 16) && co < m_real;
    const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    float* out = g.partial + cta * (int64_t)(g.mrows * 27 * 16);
    for (int t9 = 0; t9 < 9; ++t9) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        uint32_t r[16];
        tc::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(t9 * 48 + kw * 16), r);
        tc::tmem_ld_wait();
        if (row_ok) {
          float4* dst = reinterpret_cast<float4*>(out + ((int64_t)co * 27 + t9 * 3 + kw) * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem_base, 512);
}

// partial[(qslab, mchunk)][spatial cta][m 64][tap 27][n 16] -> dw[co][ci][27]
struct WgMap {
  int mchunks, Cout, Cin, swapped;
  __device__ int64_t operator()(int group, int64_t e) const {
    const int qslab = group / mchunks, mchunk = group % mchunks;
    const int nl = (int)(e % 16), t = (int)((e / 16) % 27), mm = (int)(e / (16 * 27));
    const int m = mchunk * 64 + mm, nn = qslab * 16 + nl;
    // M side = dY (co) and N side = X (ci) unless swapped; swapped results carry the mirrored tap
    const int co = swapped ? nn : m, ci = swapped ? m : nn, tap = swapped ? 26 - t : t;
    if (co >= Cout || ci >= Cin) return -1;
    return ((int64_t)co * Cin + ci) * 27 + tap;
  }
};

struct WgPlan { int swapped, cp, cq, mslabs, mchunks, qslabs, dseg, dblocks, tiles_w, tiles_h, spatial; size_t smem; };

WgPlan make_plan(int c0, int c1, int Cout, int N, int D, int H, int W) {
  WgPlan pl;
  const int Cin = c0 + c1;
  // the M side is padded to 64 rows: give it the wider tensor.  A concatenated X stays on the N side.
  pl.swapped = (c1 == 0 && Cin > Cout) ? 1 : 0;
  pl.cp = pl.swapped ? Cin : Cout;
  pl.cq = pl.swapped ? Cout : Cin;
  pl.mchunks = (pl.cp + 63) / 64;
  pl.mslabs = pl.cp >= 64 ? 4 : pl.cp / 16;
  pl.qslabs = pl.cq / 16;
  pl.tiles_w = (W + 15) / 16;
  pl.tiles_h = (H + 15) / 16;
  // enough CTAs to fill the machine a few times, but long d-runs to amortise the 432-column epilogue
  const int64_t base = (int64_t)pl.tiles_w * pl.tiles_h * N * pl.mchunks * pl.qslabs;
  int dseg = kMaxDseg;
  while (dseg > 2 && base * ((D + dseg - 1) / dseg) < (int64_t)(1.7 * B200_NUM_SMS)) dseg >>= 1;
  if (dseg > D) dseg = D;
  pl.dseg = dseg;
  pl.dblocks = (D + dseg - 1) / dseg;
  pl.spatial = pl.tiles_w * pl.tiles_h * N * pl.dblocks;
  pl.smem = kHeader + (size_t)kPStages * pl.mslabs * kPSlabBytes + (size_t)kQStages * kQBytes + 1024;
  return pl;
}


// =====================================================================================================
// ConvTranspose3d(k=2,s=2) weight gradient on the tensor cores (models/unet.py:56-58):
//     dW[ci][co][child] = sum_v x[v][ci] * gy[child(v)][co],   child = (dz,dy,dx) of the 2x finer grid
// Same MN-major / SWIZZLE_32B machinery as above: M = 64 input channels of x (TMA boxes), K = 16 coarse
// voxels of one tile row, N = 8 children x 16 output channels = 128: the eight child planes of gy arrive
// as [child][voxel][16 ch] (element-stride-2 TMA boxes) so that the children are the B descriptor's
// leading-dimension stride.  One 64 x 128 fp32 accumulator in TMEM per CTA, partials reduced in fixed order.
// =====================================================================================================
constexpr int kCtThreads = 320;          // w4-7: epilogue, w8: MMA, w9: TMA (w0-3 idle: the epilogue warps must be 4..7 for their TMEM lanes)
constexpr int kCtStages = 2;
constexpr int kCtQBytes = 8 * kPSlabBytes;  // 8 child planes of 16x16 voxels x 16 ch

struct CtParams {
  const bf16* gy; int cout;   // fine grid [N,2D,2H,2W,cout]
  int cin;
  float* partial;             // [cta][mrows][128]
  float* bias_partial;        // [qslab][spatial cta][16] per-CTA sums of gy over voxels and children (NULL: no bias gradient)
  int mrows, mslabs;
  int N, D, H, W;             // coarse geometry
  int dseg, dblocks, tiles_w, tiles_h;
};

__global__ void __launch_bounds__(kCtThreads, 1)
convt_wgrad_tc_kernel(const CtParams g, const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_gy) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = tc::smem_u32(bars);
  auto q_full = [&](int i) { return bar0 + 8u * i; };
  auto q_empty = [&](int i) { return bar0 + 8u * (2 + i); };
  auto p_full = [&](int i) { return bar0 + 8u * (4 + i); };
  auto p_empty = [&](int i) { return bar0 + 8u * (6 + i); };
  const uint32_t acc_done = bar0 + 8u * 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 192);
  const uint32_t p_stage_bytes = (uint32_t)g.mslabs * kPSlabBytes;
  uint8_t* pbuf = smem + kHeader;
  uint8_t* qbuf = pbuf + kCtStages * p_stage_bytes;
  float* bsum = reinterpret_cast<float*>(qbuf + kCtStages * kCtQBytes);   // [4 warps][16] column sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = blockIdx.x % g.tiles_w, th = blockIdx.x / g.tiles_w % g.tiles_h;
  const int rest = blockIdx.x / (g.tiles_w * g.tiles_h);
  const int n = rest / g.dblocks, db = rest % g.dblocks;
  const int mchunk = blockIdx.y, qslab = blockIdx.z;
  const int w0 = tw * 16, h0 = th * 16, d0 = db * g.dseg;
  const int planes = min(g.dseg, g.D - d0);

  if (warp == 8 && lane == 0) {
    for (int i = 0; i < kCtStages; ++i) {
      tc::mbar_init(q_full(i), 1); tc::mbar_init(q_empty(i), 1 + 4);   // released by the MMA commit AND the four column-sum warps
      tc::mbar_init(p_full(i), 1); tc::mbar_init(p_empty(i), 1);
    }
    tc::mbar_init(acc_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 4) {
    tc::tmem_alloc(tc::smem_u32(tmem_slot), 128);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // ---- one thread feeds both operands by TMA: the x tile, and the eight child planes of gy through an element-stride-2
    // tensor map (every other fine voxel in h and w: box = 16 x 16 coarse positions), landing as [child][voxel][16 ch]
    // SWIZZLE_32B slabs.  (First version: four warps gathered gy with 16-byte __ldg into the swizzled layout — 16 KB in
    // flight per SM, 148 us for the 134 MB top-level tensor; the TMA version keeps two 64 KB stages in flight.)
    if (lane == 0) {
      tma::prefetch(&tm_x);
      tma::prefetch(&tm_gy);
      for (int i = 0; i < planes; ++i) {
        const int st = i % kCtStages;
        const uint32_t ph = ((i / kCtStages) & 1) ^ 1;
        tc::mbar_wait(q_empty(st), ph);
        tc::mbar_arrive_expect_tx(q_full(st), kCtQBytes);
        const int d = d0 + i;
#pragma unroll
        for (int child = 0; child < 8; ++child)
          tma::load_5d(tc::smem_u32(qbuf + st * kCtQBytes + child * kPSlabBytes), &tm_gy, qslab * 16, 2 * w0 + (child & 1), 2 * h0 + ((child >> 1) & 1),
                       2 * d + (child >> 2), n, q_full(st));
        tc::mbar_wait(p_empty(st), ph);
        tc::mbar_arrive_expect_tx(p_full(st), p_stage_bytes);
        for (int ms = 0; ms < g.mslabs; ++ms)
          tma::load_5d(tc::smem_u32(pbuf + st * p_stage_bytes + ms * kPSlabBytes), &tm_x, mchunk * 64 + ms * 16, w0, h0, d, n, p_full(st));
      }
    }
  } else if (warp < 4) {
    // ---- bias gradient of the up-convolution = column sums of gy (models/unet.py:56-58: ConvTranspose3d has a bias): the staged
    // child planes are in shared memory anyway, so warps 0-3 add them up instead of a separate 134 MB pass over gy
    // (channel_sum_kernel: 32 us at the top level).  Thread t owns 16-byte chunk c = t & 1 (8 channels) of voxels (t >> 1) + 64 j.
    const int tid = threadIdx.x, c = tid & 1;
    const bool sum_here = g.bias_partial != nullptr && mchunk == 0;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int i = 0; i < planes; ++i) {
      const int st = i % kCtStages;
      tc::mbar_wait(q_full(st), (i / kCtStages) & 1);
      if (sum_here) {
        const uint8_t* src = qbuf + st * kCtQBytes;
#pragma unroll
        for (int child = 0; child < 8; ++child)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int v = (tid >> 1) + 64 * j;
            const uint4 q = *reinterpret_cast<const uint4*>(src + child * kPSlabBytes + swz32((uint32_t)v * 32 + c * 16));
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              acc[2 * k] += __uint_as_float(w4[k] << 16);
              acc[2 * k + 1] += __uint_as_float(w4[k] & 0xffff0000u);
            }
          }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(q_empty(st));
    }
    if (sum_here) {
      // fixed-shape fold: lanes of equal chunk parity inside the warp, then the four warps in order
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int o = 2; o < 32; o <<= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      }
      if (lane < 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) bsum[warp * 16 + c * 8 + k] = acc[k];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");     // warps 0-3 only
      if (tid < 16) {
        const float t = ((bsum[tid] + bsum[16 + tid]) + bsum[32 + tid]) + bsum[48 + tid];
        g.bias_partial[((int64_t)qslab * gridDim.x + blockIdx.x) * 16 + tid] = t;
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      // M = 64 (x channels), N = 128 (child, 16 co), both MN-major
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((64u >> 4) << 24);
      const uint64_t a_proto = desc_mn_sw32(0, g.mslabs > 1 ? kPSlabBytes : 0, 256), b_proto = desc_mn_sw32(0, kPSlabBytes, 256);
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), a_lo0 = (uint32_t)a_proto;
      const uint32_t b_hi = (uint32_t)(b_proto >> 32), b_lo0 = (uint32_t)b_proto;
      for (int i = 0; i < planes; ++i) {
        const int st = i % kCtStages;
        const uint32_t ph = (i / kCtStages) & 1;
        tc::mbar_wait(p_full(st), ph);
        tc::mbar_wait(q_full(st), ph);
        tc::tc_fence_after();
        const uint32_t p_lo = a_lo0 + (tc::smem_u32(pbuf + st * p_stage_bytes) >> 4);
        const uint32_t q_lo = b_lo0 + (tc::smem_u32(qbuf + st * kCtQBytes) >> 4);
#pragma unroll
        for (int r = 0; r < 16; ++r)
          tc::umma_bf16_ss(tmem_base, ((uint64_t)a_hi << 32) | (p_lo + r * 32), ((uint64_t)b_hi << 32) | (q_lo + r * 32), idesc, (i | r) != 0);
        tc::umma_commit(p_empty(st));
        tc::umma_commit(q_empty(st));
      }
      tc::umma_commit(acc_done);
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    tc::mbar_wait(acc_done, 0);
    tc::tc_fence_after();
    const int m_real = min(64, g.cin - mchunk * 64);
    const int ci = ew * 16 + lane;  // M = 64: row m in TMEM lane (m % 16) + 32 * (m / 16)
    const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    float* out = g.partial + cta * (int64_t)(g.mrows * 128);
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      uint32_t r[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(cc * 16), r);
      tc::tmem_ld_wait();
      if (lane < 16 && ci < m_real) {
        float4* dst = reinterpret_cast<float4*>(out + (int64_t)ci * 128 + cc * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          dst[k] = make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem_base, 128);
}

// partial[(qslab, mchunk)][spatial][m][child*16 + col] -> dw[ci][co][child]
struct CtMap {
  int mchunks, Cin, Cout;
  __device__ int64_t operator()(int group, int64_t e) const {
    const int qslab = group / mchunks, mchunk = group % mchunks;
    const int cl = (int)(e % 16), child = (int)((e / 16) % 8), mm = (int)(e / 128);
    const int ci = mchunk * 64 + mm, co = qslab * 16 + cl;
    if (ci >= Cin || co >= Cout) return -1;
    return ((int64_t)ci * Cout + co) * 8 + child;
  }
};

// bias_partial[qslab][spatial][16] -> db[co]
struct CtBiasMap {
  int Cout;
  __device__ int64_t operator()(int group, int64_t e) const {
    const int co = group * 16 + (int)e;
    return co < Cout ? co : -1;
  }
};

struct CtPlan { int mslabs, mchunks, qslabs, mrows, dseg, dblocks, tiles_w, tiles_h, spatial; size_t smem; };
CtPlan make_ct_plan(int Cin, int Cout, int N, int D, int H, int W) {
  CtPlan pl;
  pl.mchunks = (Cin + 63) / 64;
  pl.mslabs = Cin >= 64 ? 4 : Cin / 16;
  pl.mrows = Cin < 64 ? Cin : 64;
  pl.qslabs = Cout / 16;
  pl.tiles_w = (W + 15) / 16;
  pl.tiles_h = (H + 15) / 16;
  const int64_t base = (int64_t)pl.tiles_w * pl.tiles_h * N * pl.mchunks * pl.qslabs;
  // one CTA per SM (two 64-96 KB stages): the d-run minimising CTAs-per-SM x (planes streamed + fixed cost of a CTA: pipeline
  // fill, 128-column epilogue); longer runs win ties (fewer partials to fold).  (First version: runs of at most 8 planes ->
  // 256 CTAs = 1.73 waves at the top level.)
  int dseg = 1;
  int64_t best = -1;
  for (int cand = 1; cand <= D && cand <= 64; ++cand) {
    const int64_t ctas = base * ((D + cand - 1) / cand);
    const int64_t per_sm = (ctas + B200_NUM_SMS - 1) / B200_NUM_SMS;
    const int64_t cost = per_sm * (cand + 2);
    if (best < 0 || cost <= best) { best = cost; dseg = cand; }
  }
  pl.dseg = dseg;
  pl.dblocks = (D + dseg - 1) / dseg;
  pl.spatial = pl.tiles_w * pl.tiles_h * N * pl.dblocks;
  pl.smem = kHeader + (size_t)kCtStages * (pl.mslabs * kPSlabBytes + kCtQBytes) + 256 + 1024;
  return pl;
}

}  // namespace

bool b200_convt2_wgrad_tc_supported(int Cin, int Cout, int N, int D, int H, int W) {
  if (Cin % 16 || Cout % 16 || (Cin > 64 && Cin % 64)) return false;
  return N > 0 && D > 0 && H > 0 && W > 0;
}
int64_t b200_convt2_wgrad_tc_workspace(int Cin, int Cout, int N, int D, int H, int W) {
  const CtPlan pl = make_ct_plan(Cin, Cout, N, D, H, W);
  return (int64_t)pl.spatial * pl.mchunks * pl.qslabs * pl.mrows * 128 * 4 + (int64_t)pl.spatial * pl.qslabs * 16 * 4;
}
// dbias (optional): the bias gradient sum_v gy[v][co], taken from the child planes the kernel stages anyway
int b200_convt2_wgrad_tc(const void* x, const void* gy, float* dw, float* dbias, void* workspace, int N, int D, int H, int W, int Cin, int Cout,
                         cudaStream_t stream) {
  B200_REQUIRE(b200_convt2_wgrad_tc_supported(Cin, Cout, N, D, H, W), B200_ERR_UNSUPPORTED, "convt2_wgrad(tcgen05): unsupported channel counts");
  const CtPlan pl = make_ct_plan(Cin, Cout, N, D, H, W);
  CtParams g;
  g.gy = (const bf16*)gy; g.cout = Cout; g.cin = Cin;
  g.partial = (float*)workspace; g.mrows = pl.mrows; g.mslabs = pl.mslabs;
  g.bias_partial = dbias ? (float*)workspace + (int64_t)pl.spatial * pl.mchunks * pl.qslabs * pl.mrows * 128 : nullptr;
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.dseg = pl.dseg; g.dblocks = pl.dblocks; g.tiles_w = pl.tiles_w; g.tiles_h = pl.tiles_h;
  CUtensorMap tm_x, tm_gy;
  int rc = tma::make_ndhwc_map(&tm_x, x, Cin, N, D, H, W, 16, 16);
  if (rc) return rc;
  rc = tma::make_ndhwc_map_stride2(&tm_gy, gy, Cout, N, 2 * D, 2 * H, 2 * W, 16, 16);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(convt_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)pl.spatial, (unsigned)pl.mchunks, (unsigned)pl.qslabs);
  convt_wgrad_tc_kernel<<<grid, kCtThreads, pl.smem, stream>>>(g, tm_x, tm_gy);
  B200_CHECK_LAUNCH("convt2_wgrad_tc");
  launch_partial_reduce((const float*)workspace, pl.spatial, (int64_t)pl.mrows * 128, pl.qslabs * pl.mchunks, CtMap{pl.mchunks, Cin, Cout}, dw, stream);
  B200_CHECK_LAUNCH("convt2_wgrad_tc_reduce");
  if (dbias) {
    launch_partial_reduce((const float*)g.bias_partial, pl.spatial, (int64_t)16, pl.qslabs, CtBiasMap{Cout}, dbias, stream);
    B200_CHECK_LAUNCH("convt2_wgrad_tc_bias_reduce");
  }
  return B200_OK;
}

bool b200_conv3d_wgrad_tc_supported(int c0, int c1, int Cout, int N, int D, int H, int W) {
  if (c0 <= 0 || c0 % 16 || c1 % 16 || Cout % 16) return false;
  const int Cin = c0 + c1;
  if ((Cin > 64 && Cin % 64) || (Cout > 64 && Cout % 64)) return false;
  return N > 0 && D > 0 && H > 0 && W > 0;
}

int64_t b200_conv3d_wgrad_tc_workspace(int c0, int c1, int Cout, int N, int D, int H, int W) {
  const WgPlan pl = make_plan(c0, c1, Cout, N, D, H, W);
  return (int64_t)pl.spatial * pl.mchunks * pl.qslabs * (pl.cp < 64 ? pl.cp : 64) * 27 * 16 * 4;
}

int b200_conv3d_wgrad_tc(const void* x0, int c0, const void* x1, int c1, const void* dy, int Cout, float* dw, void* workspace,
                         int N, int D, int H, int W, cudaStream_t stream) {
  B200_REQUIRE(b200_conv3d_wgrad_tc_supported(c0, c1, Cout, N, D, H, W), B200_ERR_UNSUPPORTED, "conv3d_wgrad(tcgen05): unsupported channel counts");
  const WgPlan pl = make_plan(c0, c1, Cout, N, D, H, W);
  WgParams g;
  if (pl.swapped) {
    g.p = (const bf16*)x0; g.cp = c0;
    g.q0 = (const bf16*)dy; g.q1 = nullptr; g.cq0 = Cout; g.cq1 = 0;
  } else {
    g.p = (const bf16*)dy; g.cp = Cout;
    g.q0 = (const bf16*)x0; g.q1 = (const bf16*)x1; g.cq0 = c0; g.cq1 = c1;
  }
  g.partial = (float*)workspace;
  g.mrows = pl.cp < 64 ? pl.cp : 64;
  g.N = N; g.D = D; g.H = H; g.W = W;
  {
    const char* e = getenv("B200_WGRAD_M128");
    g.m128 = e ? atoi(e) : 0;
    const char* e2 = getenv("B200_WG_REPEAT");
    g.repeat = e2 ? atoi(e2) : 1;
    if (g.repeat < 1) g.repeat = 1;
  }
  g.mslabs = pl.mslabs; g.dseg = pl.dseg; g.dblocks = pl.dblocks; g.tiles_w = pl.tiles_w; g.tiles_h = pl.tiles_h;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  B200_REQUIRE(pl.mchunks <= 65535 && pl.qslabs <= 65535, B200_ERR_UNSUPPORTED, "conv3d_wgrad(tcgen05): grid too large");
  CUtensorMap tm_p, tm_q0, tm_q1;
  int rc = tma::make_ndhwc_map(&tm_p, g.p, g.cp, N, D, H, W, 16, 16);
  if (rc) return rc;
  rc = tma::make_ndhwc_map(&tm_q0, g.q0, g.cq0, N, D, H, W, kQPitch, 18);
  if (rc) return rc;
  if (g.cq1) { rc = tma::make_ndhwc_map(&tm_q1, g.q1, g.cq1, N, D, H, W, kQPitch, 18); if (rc) return rc; } else tm_q1 = tm_q0;
  dim3 grid((unsigned)pl.spatial, (unsigned)pl.mchunks, (unsigned)pl.qslabs);
  wgrad_tc_kernel<<<grid, kThreads, pl.smem, stream>>>(g, tm_p, tm_q0, tm_q1);
  B200_CHECK_LAUNCH("conv3d_wgrad_tc");
  launch_partial_reduce((const float*)workspace, pl.spatial, (int64_t)g.mrows * 27 * 16, pl.qslabs * pl.mchunks,
                        WgMap{pl.mchunks, Cout, c0 + c1, pl.swapped}, dw, stream);
  B200_CHECK_LAUNCH("conv3d_wgrad_tc_reduce");
  return B200_OK;
}
