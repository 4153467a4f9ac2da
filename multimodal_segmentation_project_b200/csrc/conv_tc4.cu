// conv_tc4.cu — row-streaming tcgen05 / TMEM implicit-GEMM 3x3x3 convolution for the full-resolution layers with few
// output channels (Cout = 16 or 32; reference models/unet.py:11,15 at the top level of the U-Net).
//
// Why another formulation.  For M = 128, K = 16, N = 48 the tensor core needs 24 cycles of math but 32 + 12 shared-memory
// wavefronts of operand fetch: conv_tc2 / conv_tc3 (M tile = 8 x 16 window of a halo plane, a different A window per filter
// tap) cannot go below 44 cycles per instruction (tools/ts_pipeline_probe.cu, profiles/r02_tcgen05_probe.md).  Here the M
// tile is ONE ROW of 128 consecutive voxels along W.  The kw tap is a 32-byte shift of the A descriptor's start address as
// before, but the three kh taps now read the SAME A matrix (input row h' feeds output rows h'+1, h', h'-1: three different
// accumulators, same TMEM lanes), so the second and third instruction reuse the A operand held in the tensor core's
// collector (tcgen05.mma .collector::a::fill / ::use / ::lastuse): 31.7 instead of 44 cycles per instruction measured.
// The three kd taps stay fused along N (adjacent TMEM column blocks = adjacent output planes) as in conv_tc2.
//
//   * work unit = output row (n, d-block of DSEG planes, h, 128-wide w tile); the 148 CTAs take equal contiguous runs of rows;
//   * stage = one input row for all DSEG + 2 planes of a 16-channel slab: ONE 5-D TMA box [16 ch][130 w][1 h][DSEG+2 d]
//     (SWIZZLE_32B, out-of-range w / d zero-filled = the padding), ~90 MMAs per stage, so the issuing warp's per-stage
//     bookkeeping (the thing that paced conv_tc3, profiles/r02_tc3_bottleneck.md) is amortised;
//   * TMEM = ring of four row accumulators (DSEG planes x Cout columns = 128 columns each): rows h'-1, h', h'+1 accumulate
//     while the epilogue drains and re-zeroes a fourth (tcgen05.ld -> +bias -> bf16 -> global, then tcgen05.st zeros: every
//     MMA accumulates, no first-touch special cases);
//   * weights of all slabs stay resident in shared memory (same packed layout as conv_tc2: b200_pack_conv3_weights modes
//     FPROP_TC / DGRAD_TC), so the same kernel serves the data gradient and the virtual concat / split of the decoder;
//   * optional BatchNorm statistics from the epilogue (STATS_CH), same contract as conv_tc3.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_maps.cuh"
#include <stdlib.h>

// `make EXTRA=-DB200_TC_DEBUG` + env B200_TC4_SKIP: bit 0 no MMAs, bit 2 no epilogue work (no TMEM loads / stores / zeroing),
// bit 3 no TMA activation loads after the first ring fill — the bottleneck knobs of conv_tc3.cu.  Absent from the default build.
#ifdef B200_TC_DEBUG
#define TC4_DBG(...) __VA_ARGS__
#else
#define TC4_DBG(...)
#endif

namespace {

using bf16 = __nv_bfloat16;

constexpr int kRowVox = 130;                         // 128 voxels + one halo voxel on each side
constexpr int kPlaneBytes = kRowVox * 32;            // one plane of a stage: [130 voxels][16 ch]
constexpr int kThreads = 384;                        // w0: act TMA, w1: weights + TMEM alloc, w2: MMA issue, w4-11: epilogue
constexpr int kSets = 4;                             // row accumulators in TMEM
constexpr int kSetCols = 128;
constexpr int kMaxStages = 5;
constexpr int kSmemHeader = 1024;

struct Tc4Params {
  const uint8_t* wpack; const float* bias;
  bf16* y0; bf16* y1; int co0, co1;
  int c0, c1;
  int N, D, H, W;
  int n_t, dseg, dblocks, wtiles, slabs, stages;
  int stage_bytes;
  long long total_rows;      // N * dblocks * wtiles * H
  float* stats;
  // BWD kernels: the output IS the gradient w.r.t. a BatchNorm+ReLU activation; xprev = that layer's pre-BN conv output
  const bf16* xprev; const float* bn_scale; const float* bn_shift; const float* bn_mean; const float* bn_invstd;
  int skip;                  // B200_TC_DEBUG builds only
};

// a run of consecutive output rows inside one (n, d-block, w-tile)
struct Seg { int n, db, wt, h_lo, h_hi; };
__device__ __forceinline__ Seg seg_at(const Tc4Params& p, long long r, long long r_end) {
  Seg s;
  const int h = (int)(r % p.H);
  long long t = r / p.H;
  s.wt = (int)(t % p.wtiles); t /= p.wtiles;
  s.db = (int)(t % p.dblocks);
  s.n = (int)(t / p.dblocks);
  s.h_lo = h;
  const long long left = r_end - r;
  s.h_hi = (int)((long long)(p.H - h) < left ? p.H : h + left);
  return s;
}

__device__ __forceinline__ uint64_t desc_kmajor_sw32(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}

// D[tmem] += A[smem] * B[smem] with an explicit collector policy for A: 0 plain, 1 fill (fetch and keep), 2 use (reuse and keep),
// 3 lastuse (reuse, then release)
#define B200_MMA_VARIANT(NAME, QUAL)                                                                                         \
  __device__ __forceinline__ void NAME(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {                                  \
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16" QUAL " [%0], %1, %2, %3, p;\n}" \
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");                                                           \
  }
B200_MMA_VARIANT(mma_plain, "")
B200_MMA_VARIANT(mma_fill, ".collector::a::fill")
B200_MMA_VARIANT(mma_use, ".collector::a::use")
B200_MMA_VARIANT(mma_last, ".collector::a::lastuse")

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// STATS_CH > 0, BWD = 0: BatchNorm batch statistics of the stored output (forward, as conv_tc3).
// STATS_CH > 0, BWD = 1: the stored output gy is the gradient w.r.t. the activation relu(bn(xprev)) of the PREVIOUS layer (this launch
// is a data gradient): the epilogue also reads xprev at the voxels it stores and accumulates the two channel sums of that
// layer's BatchNorm backward, sum(g) and invstd * sum(g * (xprev - mean)) with g = gy * [bn(xprev) > 0] — the work of
// bn_act_bwd_reduce (elementwise_kernels.cu) without its own pass over gy and xprev.
template <int STATS_CH, int BWD>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_tc4_kernel(const Tc4Params p, const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // barrier slots: [0,5) a_full, [5,10) a_empty, 10 w_full, [11,15) acc_full[set], [15,19) acc_empty[set]
  const uint32_t bar0 = tc::smem_u32(bars);
  auto a_full = [&](int i) { return bar0 + 8u * i; };
  auto a_empty = [&](int i) { return bar0 + 8u * (5 + i); };
  const uint32_t w_full = bar0 + 8u * 10;
  auto acc_full = [&](int s) { return bar0 + 8u * (11 + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (15 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  float* bias_s = reinterpret_cast<float*>(smem + 512);   // [n_t]
  uint8_t* act = smem + kSmemHeader;
  uint8_t* wts = act + (size_t)p.stages * p.stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wbytes = 864u * p.n_t;
  const long long r_begin = p.total_rows * (long long)blockIdx.x / gridDim.x;
  const long long r_end = p.total_rows * (long long)(blockIdx.x + 1) / gridDim.x;

  if (warp == 2 && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) { tc::mbar_init(a_full(i), 1); tc::mbar_init(a_empty(i), 1); }
    tc::mbar_init(w_full, 1);
    for (int s = 0; s < kSets; ++s) { tc::mbar_init(acc_full(s), 1); tc::mbar_init(acc_empty(s), 8); }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tc::smem_u32(tmem_slot), 512);
    tc::tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma::prefetch(&tm0);
    if (p.c1) tma::prefetch(&tm1);
  }
  if (threadIdx.x >= 128 && threadIdx.x - 128 < p.n_t) bias_s[threadIdx.x - 128] = p.bias ? p.bias[threadIdx.x - 128] : 0.f;
  float* bn_s = reinterpret_cast<float*>(smem + 640);     // BWD: [3][32] scale, shift, mean of the previous layer's BatchNorm
  if (BWD && threadIdx.x >= 128 && threadIdx.x - 128 < p.n_t) {
    const int c = threadIdx.x - 128;
    bn_s[c] = p.bn_scale[c]; bn_s[32 + c] = p.bn_shift[c]; bn_s[64 + c] = p.bn_mean[c];
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================== activation stages by TMA: (input row, slab), continuous across segments =====================
    int st = 0;
    uint32_t ph = 1;
    TC4_DBG(bool first_lap = true;)
    for (long long r = r_begin; r < r_end;) {
      const Seg sg = seg_at(p, r, r_end);
      const int d0 = sg.db * p.dseg, w0 = sg.wt * 128;
      const int hp_lo = max(sg.h_lo - 1, 0), hp_hi = min(sg.h_hi, p.H - 1);
      for (int hp = hp_lo; hp <= hp_hi; ++hp) {
        for (int s = 0; s < p.slabs; ++s) {
          const int c = s * 16;
          const CUtensorMap* tm = c < p.c0 ? &tm0 : &tm1;
          const int cc = c < p.c0 ? c : c - p.c0;
          tc::mbar_wait(a_empty(st), ph);
          if (tc::elect_one()) {
#ifdef B200_TC_DEBUG
            if ((p.skip & 8) && !first_lap) tc::mbar_arrive(a_full(st));
            else
#endif
            {
              tc::mbar_arrive_expect_tx(a_full(st), (uint32_t)((p.dseg + 2) * kPlaneBytes));
              tma::load_5d(tc::smem_u32(act + (size_t)st * p.stage_bytes), tm, cc, w0 - 1, hp, d0 - 1, sg.n, a_full(st));
            }
          }
          __syncwarp();
          if (++st == p.stages) { st = 0; ph ^= 1u; TC4_DBG(first_lap = false;) }
        }
      }
      r += sg.h_hi - sg.h_lo;
    }
  } else if (warp == 1) {
    // ===================== weights: every slab resident for the whole kernel =====================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(w_full, wbytes * p.slabs);
      for (int s = 0; s < p.slabs; ++s) tc::bulk_g2s(tc::smem_u32(wts + (size_t)s * wbytes), p.wpack + (size_t)s * wbytes, wbytes, w_full);
    }
  } else if (warp == 2) {
    // ===================== MMA issue (whole warp runs the loop, one elected lane issues) =====================
    const uint32_t n_t = p.n_t;
    const uint64_t a_proto = desc_kmajor_sw32(0, 256);                          // 8-voxel groups back to back
    const uint64_t b_proto = tc::smem_desc_kmajor_noswz(0, 48u * n_t, 128);     // [k-chunk][kd * n_t rows][16 B]
    const uint32_t a_hi = (uint32_t)(a_proto >> 32), a_lo0 = (uint32_t)a_proto + (tc::smem_u32(act) >> 4);
    const uint32_t b_hi = (uint32_t)(b_proto >> 32), b_lo0 = (uint32_t)b_proto + (tc::smem_u32(wts) >> 4);
    const uint32_t b_tap16 = 6u * n_t;                                          // (bytes per (kh,kw) = 96 * n_t) >> 4
    const uint32_t idesc0 = tc::idesc_bf16_f32(128, 0), idesc_step = (n_t >> 3) << 17;
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, wslab16 = wbytes >> 4;
    uint32_t btap[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) btap[t] = (uint32_t)t * b_tap16;      // B offset of filter tap (kh, kw) = kh * 3 + kw
    tc::mbar_wait(w_full, 0);
    tc::tc_fence_after();
    int st = 0;
    uint32_t ph = 0;
    long long ord0 = 0;       // ordinal (within this CTA) of the first output row of the current segment
    for (long long r = r_begin; r < r_end;) {
      const Seg sg = seg_at(p, r, r_end);
      const int d0 = sg.db * p.dseg;
      const int planes = min(p.dseg, p.D - d0);
      const int hp_lo = max(sg.h_lo - 1, 0), hp_hi = min(sg.h_hi, p.H - 1);
      for (int hp = hp_lo; hp <= hp_hi; ++hp) {
        // output rows fed by this input row: h0 = hp + 1 - kh; TMEM set of an output row = its ordinal & 3
        uint32_t dset[3];
        bool kh_ok[3];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int h0 = hp + 1 - kh;
          kh_ok[kh] = h0 >= sg.h_lo && h0 < sg.h_hi;
          dset[kh] = tmem_base + (uint32_t)((ord0 + (h0 - sg.h_lo)) & (kSets - 1)) * kSetCols;
        }
        // first touch of an output row: its accumulator must have been drained and re-zeroed (use u of a set waits for the
        // u-th completion of acc_empty; completion 0 is the initial zeroing).  Row hp + 1 is new with every input row; row hp
        // only when the segment starts at the top of the volume (no input row above it).
        if (hp == hp_lo && kh_ok[1]) {
          const long long o = ord0 + (hp - sg.h_lo);
          tc::mbar_wait(acc_empty((int)(o & (kSets - 1))), (uint32_t)((o >> 2) & 1));
        }
        if (kh_ok[0]) {
          const long long o = ord0 + (hp + 1 - sg.h_lo);
          tc::mbar_wait(acc_empty((int)(o & (kSets - 1))), (uint32_t)((o >> 2) & 1));
        }
        tc::tc_fence_after();
        for (int s = 0; s < p.slabs; ++s) {
          tc::mbar_wait(a_full(st), ph);
          tc::tc_fence_after();
          const uint32_t a_st = a_lo0 + (uint32_t)st * stage16;
          const uint32_t w_lo = b_lo0 + (uint32_t)s * wslab16;
          if (tc::elect_one()) {
            // the issuing thread's instruction stream IS the pace of the kernel: the interior case (all three output rows inside
            // the segment) is a branch-free block of nine MMAs per input plane; border rows take the generic path
            const int q_lo = d0 > 0 ? -1 : 0, q_hi = d0 + planes < p.D ? planes : planes - 1;   // planes that exist in the volume
            if (false TC4_DBG(|| (p.skip & 1))) {
            } else if (kh_ok[0] && kh_ok[2]) {
              for (int q = q_lo; q <= q_hi; ++q) {
                const int kd_lo = max(0, q + 2 - planes), kd_hi = min(2, q + 1);
                const uint32_t col = (uint32_t)(p.dseg - 2 - q + kd_lo) * n_t;
                const uint32_t idesc = idesc0 + idesc_step * (uint32_t)(kd_hi - kd_lo + 1);
                const uint32_t b_q = w_lo + (uint32_t)kd_lo * n_t;
                const uint32_t a_q = a_st + (uint32_t)(q + 1) * (kPlaneBytes >> 4);
                const uint32_t d0c = dset[0] + col, d1c = dset[1] + col, d2c = dset[2] + col;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                  const uint64_t ad = ((uint64_t)a_hi << 32) | (a_q + 2u * kw);
                  mma_fill(d0c, ad, ((uint64_t)b_hi << 32) | (b_q + btap[kw]), idesc);
                  mma_use(d1c, ad, ((uint64_t)b_hi << 32) | (b_q + btap[3 + kw]), idesc);
                  mma_last(d2c, ad, ((uint64_t)b_hi << 32) | (b_q + btap[6 + kw]), idesc);
                }
              }
            } else {
              for (int q = q_lo; q <= q_hi; ++q) {
                const int kd_lo = max(0, q + 2 - planes), kd_hi = min(2, q + 1);
                const uint32_t col = (uint32_t)(p.dseg - 2 - q + kd_lo) * n_t;
                const uint32_t idesc = idesc0 + idesc_step * (uint32_t)(kd_hi - kd_lo + 1);
                const uint32_t b_q = w_lo + (uint32_t)kd_lo * n_t;
                const uint32_t a_q = a_st + (uint32_t)(q + 1) * (kPlaneBytes >> 4);
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                  const uint64_t ad = ((uint64_t)a_hi << 32) | (a_q + 2u * kw);
#pragma unroll
                  for (int kh = 0; kh < 3; ++kh)
                    if (kh_ok[kh]) mma_plain(dset[kh] + col, ad, ((uint64_t)b_hi << 32) | (b_q + btap[kh * 3 + kw]), idesc);
                }
              }
            }
            tc::umma_commit(a_empty(st));
            if (s == p.slabs - 1) {
              // output row hp - 1 has now seen its three input rows; at the bottom of the volume row hp is complete as well
              if (kh_ok[2]) tc::umma_commit(acc_full((int)((ord0 + (hp - 1 - sg.h_lo)) & (kSets - 1))));
              if (hp == hp_hi && kh_ok[1] && hp == sg.h_hi - 1) tc::umma_commit(acc_full((int)((ord0 + (hp - sg.h_lo)) & (kSets - 1))));
            }
          }
          __syncwarp();
          if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
      }
      ord0 += sg.h_hi - sg.h_lo;
      r += sg.h_hi - sg.h_lo;
    }
  } else if (warp >= 4) {
    // ===================== epilogue: warps 4-7 drain the first half of the planes, warps 8-11 the second half =====================
    const int grp = (warp - 4) >> 2;
    const int ew = warp & 3;                 // TMEM lane quadrant
    const int pl_per = p.dseg >> 1;          // dseg is 8 (Cout 16) or 4 (Cout 32)
    const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
    float st_sum[STATS_CH > 0 ? STATS_CH : 1], st_sq[STATS_CH > 0 ? STATS_CH : 1];
#pragma unroll
    for (int i = 0; i < (STATS_CH > 0 ? STATS_CH : 1); ++i) st_sum[i] = st_sq[i] = 0.f;
    // this warp owns (reads, then re-zeroes) the columns of its planes: planes sit in DESCENDING column order, so the first
    // half of the planes is the upper half of the columns.  All accumulators start at zero.
    const uint32_t own_cols = (uint32_t)((1 - grp) * (kSetCols / 2));
    for (int s = 0; s < kSets; ++s)
      for (int c = 0; c < kSetCols / 2; c += 16) tmem_st16_zero(tmem_base + lane_base + (uint32_t)(s * kSetCols) + own_cols + (uint32_t)c);
    tmem_st_wait();
    tc::tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int s = 0; s < kSets; ++s) tc::mbar_arrive(acc_empty(s));
    long long ord = 0;
    for (long long r = r_begin; r < r_end;) {
      const Seg sg = seg_at(p, r, r_end);
      const int d0 = sg.db * p.dseg;
      const int planes = min(p.dseg, p.D - d0);
      const int w = sg.wt * 128 + ew * 32 + lane;
      const bool wok = w < p.W;
      for (int h0 = sg.h_lo; h0 < sg.h_hi; ++h0, ++ord) {
        const int set = (int)(ord & (kSets - 1));
        // BWD: the previous layer's pre-BN values of this row are fetched BEFORE waiting for the accumulator (their addresses do
        // not depend on it): four 32-byte voxels per thread in flight behind the wait instead of one exposed round trip per plane
        constexpr int kPre = (STATS_CH > 0 && BWD) ? 4 : 1;
        uint4 xpre[kPre][2];
        if (STATS_CH > 0 && BWD) {
#pragma unroll
          for (int j = 0; j < kPre; ++j) {
            const int pi = STATS_CH == 16 ? j : (j >> 1), cc = STATS_CH == 16 ? 0 : (j & 1);
            const int pl = grp * pl_per + pi;
            const int64_t row = (((int64_t)sg.n * p.D + d0 + pl) * p.H + h0) * p.W + w;
            xpre[j][0] = xpre[j][1] = make_uint4(0, 0, 0, 0);
            if (wok && pl < planes) {
              const uint4* xp = reinterpret_cast<const uint4*>(p.xprev + row * p.co0 + cc * 16);
              xpre[j][0] = __ldcs(xp);
              xpre[j][1] = __ldcs(xp + 1);
            }
          }
        }
        tc::mbar_wait(acc_full(set), (uint32_t)((ord >> 2) & 1));
        tc::tc_fence_after();
        const uint32_t tset = tmem_base + lane_base + (uint32_t)set * kSetCols;
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) {
          if (pi >= (true TC4_DBG(&& !(p.skip & 4)) ? pl_per : 0)) break;
          const int pl = grp * pl_per + pi;
          const uint32_t col0 = (uint32_t)(p.dseg - 1 - pl) * p.n_t;
          const bool valid = wok && pl < planes;
          const int64_t row = (((int64_t)sg.n * p.D + d0 + pl) * p.H + h0) * p.W + w;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            if (cc * 16 >= p.n_t) break;
            uint32_t v[16];
            tc::tmem_ld16(tset + col0 + cc * 16, v);
            tc::tmem_ld_wait();
            uint32_t packed[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 bb = *reinterpret_cast<const float2*>(bias_s + cc * 16 + 2 * i);
              __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(v[2 * i]) + bb.x, __uint_as_float(v[2 * i + 1]) + bb.y);
              packed[i] = *reinterpret_cast<uint32_t*>(&hb);
              if (STATS_CH > 0 && !BWD && valid) {
                const float2 yv = __bfloat1622float2(hb);
                const float e0 = yv.x - bb.x, e1 = yv.y - bb.y;
                const int c = STATS_CH > 0 ? (cc * 16 + 2 * i) % STATS_CH : 0;
                st_sum[c] += e0;
                st_sq[c] = fmaf(e0, e0, st_sq[c]);
                st_sum[STATS_CH > 0 ? c + 1 : 0] += e1;
                st_sq[STATS_CH > 0 ? c + 1 : 0] = fmaf(e1, e1, st_sq[STATS_CH > 0 ? c + 1 : 0]);
              }
            }
            if (STATS_CH > 0 && BWD && valid) {
              const int j = (STATS_CH > 0 && BWD) ? (STATS_CH == 16 ? pi : pi * 2 + cc) : 0;
              const uint4 xa = xpre[j][0], xb = xpre[j][1];
              const uint32_t xw[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  const int c = STATS_CH > 0 ? (cc * 16 + 2 * i + hh) % STATS_CH : 0;
                  const float xv = hh ? __uint_as_float(xw[i] & 0xffff0000u) : __uint_as_float(xw[i] << 16);
                  const float gv = hh ? __uint_as_float(packed[i] & 0xffff0000u) : __uint_as_float(packed[i] << 16);
                  const float xc = xv - bn_s[64 + c];
                  const float pre = __bfloat162float(__float2bfloat16_rn(fmaf(xc, bn_s[c], bn_s[32 + c])));
                  const float g = pre > 0.f ? gv : 0.f;
                  st_sum[c] += g;
                  st_sq[c] = fmaf(g, xc, st_sq[c]);
                }
              }
            }
            if (valid) {
              const int ch = cc * 16;
              bf16* dst = ch < p.co0 ? p.y0 + row * p.co0 + ch : p.y1 + row * p.co1 + (ch - p.co0);
              reinterpret_cast<uint4*>(dst)[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
              reinterpret_cast<uint4*>(dst)[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            }
          }
        }
        // re-zero this warp's part of the accumulator and hand it back
        if (true TC4_DBG(&& !(p.skip & 4)))
          for (int c = 0; c < kSetCols / 2; c += 16) tmem_st16_zero(tset + own_cols + (uint32_t)c);
        tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(acc_empty(set));
      }
      r += sg.h_hi - sg.h_lo;
    }
    if (STATS_CH > 0) {
      // fixed-shape fold (warp butterflies, then the eight warps in order): run-to-run deterministic.  The stage ring is free:
      // every accumulator this CTA waited for is complete, hence every stage consumed.
      float* red = reinterpret_cast<float*>(act);            // [8 warps][2][STATS_CH]
#pragma unroll
      for (int c = 0; c < (STATS_CH > 0 ? STATS_CH : 1); ++c) {
        const float a = warp_sum(st_sum[c]), b = warp_sum(st_sq[c]);
        if (lane == 0) { red[(warp - 4) * 2 * STATS_CH + c] = a; red[(warp - 4) * 2 * STATS_CH + STATS_CH + c] = b; }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int e = threadIdx.x - 128;
      if (e < 2 * STATS_CH) {
        float tot = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) tot += red[wq * 2 * STATS_CH + e];
        if (BWD && e >= STATS_CH) tot *= p.bn_invstd[e - STATS_CH];
        p.stats[(size_t)blockIdx.x * 2 * STATS_CH + e] = tot;
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// [N, D, H, W, C] bf16, box = 16 channels x 130 w x 1 h x (dseg + 2) d x 1 n, SWIZZLE_32B, zero fill outside
int make_row_map(CUtensorMap* tm, const void* base, int C, int N, int D, int H, int W, int box_d) {
  tma::EncodeTiledFn enc = tma::get_encode();
  B200_REQUIRE(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {16, (cuuint32_t)kRowVox, 1, (cuuint32_t)box_d, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(row box) failed (%d) for C=%d N=%d D=%d H=%d W=%d", (int)r, C, N, D, H, W);
  return B200_OK;
}

struct Plan { int n_t, dseg, dblocks, wtiles, slabs, stages, stage_bytes; size_t smem; long long rows; };
bool make_plan(int c0, int c1, int co0, int co1, int N, int D, int H, int W, Plan* pl) {
  const int cout = co0 + co1, cin = c0 + c1;
  if (cout != 16 && cout != 32) return false;
  if (cin % 16 || c0 % 16 || (co1 && (co0 != 16 || co1 != 16))) return false;
  pl->n_t = cout;
  pl->dseg = kSetCols / cout;                      // 8 or 4 planes per row accumulator
  pl->dblocks = (D + pl->dseg - 1) / pl->dseg;
  pl->wtiles = (W + 127) / 128;
  pl->slabs = cin / 16;
  pl->stage_bytes = ((pl->dseg + 2) * kPlaneBytes + 255) / 256 * 256;
  const size_t wtotal = (size_t)864 * cout * pl->slabs;
  const size_t budget = 226 * 1024 - kSmemHeader - 1024;
  if (wtotal + 2 * (size_t)pl->stage_bytes > budget) return false;
  int stages = (int)((budget - wtotal) / pl->stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  pl->stages = stages;
  pl->smem = kSmemHeader + (size_t)stages * pl->stage_bytes + wtotal + 1024;
  pl->rows = (long long)N * pl->dblocks * pl->wtiles * H;
  return true;
}

int g_rowstream = -1;
bool rowstream_on() {
  if (g_rowstream < 0) { const char* e = getenv("B200_CONV_ROWSTREAM"); g_rowstream = e ? atoi(e) : 1; }
  return g_rowstream != 0;
}

}  // namespace

void b200_conv3d_k3_tc4_enable(int on) { g_rowstream = on ? 1 : 0; }

// the row-streaming kernel wants full 128-voxel rows (W a multiple of 128 keeps every TMEM lane busy) and enough rows for 148 CTAs
bool b200_conv3d_k3_tc4_wanted(int c0, int c1, int co0, int co1, int N, int D, int H, int W) {
  Plan pl;
  if (!rowstream_on() || !make_plan(c0, c1, co0, co1, N, D, H, W, &pl)) return false;
  if (W % 128 > 0 && W % 128 < 96) return false;
  return pl.rows >= 8LL * B200_NUM_SMS;
}

int b200_conv3d_k3_tc4_stats_blocks(int c0, int c1, int co0, int co1, int N, int D, int H, int W) {
  if (co1 != 0 || !b200_conv3d_k3_tc4_wanted(c0, c1, co0, co1, N, D, H, W)) return 0;
  return B200_NUM_SMS;
}

int b200_conv3d_k3_tc4(const void* x0, int c0, const void* x1, int c1, const void* wpack, const float* bias, void* y0, int co0, void* y1,
                       int co1, int N, int D, int H, int W, cudaStream_t stream, float* stats, const void* xprev, const float* bn_scale,
                       const float* bn_shift, const float* bn_mean, const float* bn_invstd) {
  B200_REQUIRE(b200_aligned(x0, 16) && b200_aligned(x1, 16) && b200_aligned(y0, 16) && b200_aligned(y1, 16) && b200_aligned(wpack, 16),
               B200_ERR_ALIGN, "conv3d_k3(row-streaming): pointers must be 16-byte aligned");
  Plan pl;
  B200_REQUIRE(make_plan(c0, c1, co0, co1, N, D, H, W, &pl), B200_ERR_UNSUPPORTED, "conv3d_k3(row-streaming): unsupported problem (%d+%d)->(%d+%d)", c0, c1, co0, co1);
  B200_REQUIRE(!stats || co1 == 0, B200_ERR_UNSUPPORTED, "conv3d_k3(row-streaming): fused BatchNorm statistics need one output tensor");
  Tc4Params p;
  p.wpack = (const uint8_t*)wpack; p.bias = bias;
  p.y0 = (bf16*)y0; p.y1 = (bf16*)y1; p.co0 = co0; p.co1 = co1;
  p.c0 = c0; p.c1 = c1;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.n_t = pl.n_t; p.dseg = pl.dseg; p.dblocks = pl.dblocks; p.wtiles = pl.wtiles; p.slabs = pl.slabs; p.stages = pl.stages;
  p.stage_bytes = pl.stage_bytes;
  p.total_rows = pl.rows;
  p.stats = stats;
  p.xprev = (const bf16*)xprev; p.bn_scale = bn_scale; p.bn_shift = bn_shift; p.bn_mean = bn_mean; p.bn_invstd = bn_invstd;
  B200_REQUIRE(!xprev || (stats && bn_scale && bn_shift && bn_mean && bn_invstd && b200_aligned(xprev, 16)), B200_ERR_SHAPE,
               "conv3d_k3(row-streaming): the BatchNorm-backward epilogue needs partials, scale, shift, mean, invstd and an aligned xprev");
  p.skip = 0;
#ifdef B200_TC_DEBUG
  { const char* e = getenv("B200_TC4_SKIP"); p.skip = e ? atoi(e) : 0; }
#endif
  CUtensorMap tm0, tm1;
  int rc = make_row_map(&tm0, x0, c0, N, D, H, W, pl.dseg + 2);
  if (rc) return rc;
  if (c1) { rc = make_row_map(&tm1, x1, c1, N, D, H, W, pl.dseg + 2); if (rc) return rc; } else tm1 = tm0;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(conv3d_tc4_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CUDA(cudaFuncSetAttribute(conv3d_tc4_kernel<16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CUDA(cudaFuncSetAttribute(conv3d_tc4_kernel<32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CUDA(cudaFuncSetAttribute(conv3d_tc4_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CUDA(cudaFuncSetAttribute(conv3d_tc4_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = (int)(pl.rows < B200_NUM_SMS ? pl.rows : B200_NUM_SMS);
  B200_REQUIRE(!stats || grid == B200_NUM_SMS, B200_ERR_UNSUPPORTED, "conv3d_k3(row-streaming): fused statistics expect a full grid");
  if (stats && xprev && pl.n_t == 16) conv3d_tc4_kernel<16, 1><<<grid, kThreads, pl.smem, stream>>>(p, tm0, tm1);
  else if (stats && xprev) conv3d_tc4_kernel<32, 1><<<grid, kThreads, pl.smem, stream>>>(p, tm0, tm1);
  else if (stats && pl.n_t == 16) conv3d_tc4_kernel<16, 0><<<grid, kThreads, pl.smem, stream>>>(p, tm0, tm1);
  else if (stats) conv3d_tc4_kernel<32, 0><<<grid, kThreads, pl.smem, stream>>>(p, tm0, tm1);
  else conv3d_tc4_kernel<0, 0><<<grid, kThreads, pl.smem, stream>>>(p, tm0, tm1);
  B200_CHECK_LAUNCH("conv3d_k3_tc4");
  return B200_OK;
}
