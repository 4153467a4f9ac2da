// wgrad_tc4.cu — weight gradient of the 3x3x3 convolution for the full-resolution 16-channel layers of the U-Net
// (models/unet.py:11,15 with 16 or 16+16 input channels and 16 output channels), voxel-PAIR formulation.
//
//     dW[co][ci][kd][kh][kw] = sum_v  X[v + (kd-1, kh-1, kw-1)][ci] * dY[v][co]
//
// wgrad_tc2.cu reaches M = 64 (48 useful) x N = 48 per instruction — 37 % of the tensor core at best — because a
// 16-channel operand row only offers (kw, ci) = 48 rows and (kh, co) = 48 columns.  Here a 16-channel NDHWC tensor is read
// as [N*D][H][W/2][32]: an operand row is a PAIR of w-adjacent voxels (64 bytes, SWIZZLE_64B, plain dense TMA boxes), the
// K = 16 dimension of one instruction walks 16 pairs = 32 voxels, and
//
//   * A = four consecutive halo rows of one X plane: rows (g, p, ci), g = halo row, p = voxel parity       -> M = 128, all real;
//   * B = two consecutive tile rows of one dY plane:  columns (g', p', co)                                    -> N = 64;
//   * kh = g - g' is a real tap for 6 of the 8 (g, g') blocks;
//   * kw: the A descriptor starts one voxel (32 B) before / after the pair grid of dY — "type 0" (X voxel = dY voxel - 1 + p - p')
//     holds kw = 0 for p = p' and kw = 1 for (p, p') = (1, 0); "type 1" (start two voxels later) holds kw = 2 for p = p' and
//     kw = 1 for (p, p') = (0, 1): 6 of the 8 (type, p, p') blocks are real taps and every tap is complete (even and odd
//     output voxels arrive through different blocks);
//   * kd = X plane - dY plane + 1 selects one of three accumulators; the three instructions of an (X rows, type) share A
//     through the collector (fill / use / lastuse), which puts an N = 64 instruction on its 32-cycle math floor
//     (profiles/r02_tcgen05_probe.md).
//
// 36 of 64 blocks of every instruction are useful (56 %, against 28 % of the array for wgrad_tc2's M = 64 x N = 48), the
// operands are fetched from shared memory a third as often per useful MAC, and the issuing thread retires 48 instructions
// per 32 x 16-voxel plane.  Six accumulators D[kd][type] = 128 x 64 fp32 (384 TMEM columns) live for the whole d-run; the
// epilogue folds the (p, p') blocks of a tap (lanes 16 apart: one shuffle) and writes two partial slices per CTA (g' = 0, 1
// carry the same taps from different halo rows); partials are reduced in fixed order (partial_reduce.cuh, deterministic).
// Halo and ragged tiles are zero-filled by the TMA unit; planes outside the volume are skipped.
// Not extensible to 16-channel slabs of wider tensors (16 -> 32, encoder.1.c0): a TMA box whose inner extent (32 B) is narrower
// than its swizzle span lands in 64-byte slots, half of each unwritten (measured: tools/tma_inner32_probe.cu), so the voxel-pair
// rows cannot be formed from a 32-channel tensor; that layer stays on wgrad_tc2.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_maps.cuh"
#include "partial_reduce.cuh"
#include <stdlib.h>

namespace {

using bf16 = __nv_bfloat16;

constexpr int kRowB = 64;                            // bytes per voxel pair: [2 voxels][16 ch] bf16
constexpr int kTileW = 32, kTileH = 16;              // output voxels per CTA plane
constexpr int kQPairs = 20;                          // voxel pairs per staged halo row: voxels w0-2 .. w0+37 (w0-1 .. w0+32 used)
constexpr int kQRowBytes = kQPairs * kRowB;          // 1280
constexpr int kQBytes = 18 * kQRowBytes;             // 23040
constexpr int kQStageBytes = 23 * 1024;              // stage pitch (bases stay 1024-byte aligned)
constexpr int kPRowBytes = (kTileW / 2) * kRowB;     // one tile row of dY: 1024
constexpr int kPBytes = kTileH * kPRowBytes;         // 16384
constexpr int kQStages = 3, kPStages = 4;            // a dY plane stays resident for three X planes (kd = 0,1,2)
constexpr int kThreads = 256;                        // w0: TMA, w1: MMA, w2: TMEM alloc, w4-7: accumulator zeroing + epilogue
constexpr int kTmemCols = 512;                       // 3 kd x 2 types x 64 columns used
constexpr int kSliceFloats = 16 * 27 * 16;           // one partial slice [ci][tap][co]
constexpr int kPartialFloats = 2 * kSliceFloats;     // two slices (g' = 0, 1) per CTA
constexpr int kMaxDseg = 128;
constexpr size_t kSmemBytes = 1024 + (size_t)kPStages * kPBytes + (size_t)kQStages * kQStageBytes + 1024;

struct Wg4Params {
  float* partial;
  int N, D, H, W;
  int tiles_w, tiles_h;
  long long units, upc;         // plane-tiles per X slab (N * tiles * D) and per CTA: a CTA takes the contiguous run [x * upc, (x + 1) * upc)
  int skip;                     // bottleneck knobs (env B200_WG4_SKIP, tools/bench_wgrad_pair.py): 1 = no MMAs, 2 = no TMA after the ring fill
};

__device__ __forceinline__ uint64_t desc_mn_sw64(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // version
  d |= (uint64_t)4 << 61;                       // SWIZZLE_64B (absolute-address swizzle)
  return d;
}

#define B200_WG4_MMA(NAME, QUAL)                                                                                             \
  __device__ __forceinline__ void NAME(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {                                  \
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16" QUAL " [%0], %1, %2, %3, p;\n}" \
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");                                                           \
  }
B200_WG4_MMA(mma_plain, "")
B200_WG4_MMA(mma_fill, ".collector::a::fill")
B200_WG4_MMA(mma_use, ".collector::a::use")
B200_WG4_MMA(mma_last, ".collector::a::lastuse")

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void load_4d(uint32_t dst, const CUtensorMap* tm, int c, int wp, int h, int nd, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c), "r"(wp), "r"(h), "r"(nd), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc4_kernel(const Wg4Params g, const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_q0,
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
  uint8_t* pbuf = smem + 1024;                   // stage bases stay 1024-byte aligned
  uint8_t* qbuf = pbuf + kPStages * kPBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qslab = blockIdx.z;       // which of the two 16-channel X tensors (virtual concat)
  // Work = plane-tiles u = (tile column (n, th, tw), plane d), d fastest.  Every plane-tile of an X slab adds into the SAME dW, so a
  // CTA takes a contiguous run of u whatever tile columns it crosses — 148 (or 74 + 74) equal shares instead of 128 CTAs of whole
  // d-blocks — and walks it as segments (one per tile column touched), each with its own two halo planes of X.
  const long long u_begin = (long long)blockIdx.x * g.upc, u_end = min(u_begin + g.upc, g.units);

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
    // ===================== TMA producer: dY plane i, then X plane i - 1 =====================
    const CUtensorMap* tq = qslab == 0 ? &tm_q0 : &tm_q1;
    if (tc::elect_one()) {
      tma::prefetch(&tm_p);
      tma::prefetch(tq);
    }
    int pst = 0, qst = 0;
    uint32_t pph = 0, qph = 0;
    long long fed = 0;                // planes of either operand issued so far (knob B200_WG4_SKIP & 2)
    for (long long u = u_begin; u < u_end;) {
      const long long col = u / g.D;
      const int d0 = (int)(u - col * g.D), planes = (int)min((long long)(g.D - d0), u_end - u);
      const int tw = (int)(col % g.tiles_w), th = (int)(col / g.tiles_w % g.tiles_h), n = (int)(col / ((long long)g.tiles_w * g.tiles_h));
      const int w0 = tw * kTileW, h0 = th * kTileH, nd0 = n * g.D + d0;
      const int q_lo = d0 > 0 ? -1 : 0;                               // X planes (relative to d0) that exist in the volume
      const int q_hi = d0 + planes < g.D ? planes : planes - 1;
      for (int i = 0; i <= planes + 1; ++i) {
        if (i < planes) {
          tc::mbar_wait(p_empty(pst), pph ^ 1u);
          if (tc::elect_one()) {
            if ((g.skip & 2) && fed >= kPStages) {
              tc::mbar_arrive(p_full(pst));
            } else {
              tc::mbar_arrive_expect_tx(p_full(pst), kPBytes);
              load_4d(tc::smem_u32(pbuf + pst * kPBytes), &tm_p, 0, w0 >> 1, h0, nd0 + i, p_full(pst));
            }
          }
          __syncwarp();
          if (++pst == kPStages) { pst = 0; pph ^= 1u; }
        }
        const int q = i - 1;
        if (q >= q_lo && q <= q_hi) {
          tc::mbar_wait(q_empty(qst), qph ^ 1u);
          if (tc::elect_one()) {
            if ((g.skip & 2) && fed >= kPStages) {
              tc::mbar_arrive(q_full(qst));
            } else {
              tc::mbar_arrive_expect_tx(q_full(qst), kQBytes);
              load_4d(tc::smem_u32(qbuf + qst * kQStageBytes), tq, 0, (w0 >> 1) - 1, h0 - 1, nd0 + q, q_full(qst));
            }
          }
          __syncwarp();
          if (++qst == kQStages) { qst = 0; qph ^= 1u; }
        }
        ++fed;
      }
      u += planes;
    }
  } else if (warp == 1) {
    // ===================== MMA issue (whole warp walks the loop, one elected lane issues) =====================
    // M = 128 rows (g, p, ci); N = 64 columns (g', p', co); both operands MN-major, K = 16 voxel pairs
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a_proto = desc_mn_sw64(0, kQRowBytes, 8 * kRowB);   // M groups = halo rows; K groups of 8 pairs
    const uint64_t b_proto = desc_mn_sw64(0, kPRowBytes, 8 * kRowB);   // N groups = tile rows
    const uint32_t a_hi = (uint32_t)(a_proto >> 32), a_lo0 = (uint32_t)a_proto + (tc::smem_u32(qbuf) >> 4);
    const uint32_t b_hi = (uint32_t)(b_proto >> 32), b_lo0 = (uint32_t)b_proto + (tc::smem_u32(pbuf) >> 4);
    constexpr uint32_t kAStep = (2 * kQRowBytes) >> 4, kBStep = (2 * kPRowBytes) >> 4;   // one row pair
    tc::mbar_wait(acc_zero, 0);       // the epilogue warps have zeroed the accumulators: every instruction accumulates
    tc::tc_fence_after();
    int qst = 0, p_ready = 0;         // p_ready: dY planes waited for so far, over all segments (ring position = p_ready & 3)
    uint32_t qph = 0;
    for (long long u = u_begin; u < u_end;) {
      const long long col = u / g.D;
      const int d0 = (int)(u - col * g.D), planes = (int)min((long long)(g.D - d0), u_end - u);
      const int q_lo = d0 > 0 ? -1 : 0, q_hi = d0 + planes < g.D ? planes : planes - 1;
      const int p_base = p_ready;     // ring position of this segment's dY plane 0
      for (int q = q_lo; q <= q_hi; ++q) {
        tc::mbar_wait(q_full(qst), qph);
        const int need = p_base + min(q + 2, planes);
        while (p_ready < need) {
          tc::mbar_wait(p_full(p_ready & (kPStages - 1)), (uint32_t)(p_ready >> 2) & 1u);
          ++p_ready;
        }
        tc::tc_fence_after();
        if (tc::elect_one()) {
          // dY plane q - kd + 1 meets this X plane through kd; kd_lo .. kd_hi are the planes this segment owns
          const int kd_lo = max(0, q + 2 - planes), kd_hi = min(2, q + 1);
          const uint32_t a_q = a_lo0 + (uint32_t)qst * (kQStageBytes >> 4);
          uint32_t b_kd[3];
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) b_kd[kd] = b_lo0 + (uint32_t)((p_base + q + 1 - kd) & (kPStages - 1)) * (kPBytes >> 4);
          if (g.skip & 1) {
          } else if (kd_lo == 0 && kd_hi == 2) {
#pragma unroll 2
            for (int i = 0; i < kTileH / 2; ++i) {
#pragma unroll
              for (int type = 0; type < 2; ++type) {
                // type 0: X voxel w0 - 1 + 2k + p (one voxel = 32 B into the staged row); type 1: two voxels later
                const uint64_t ad = ((uint64_t)a_hi << 32) | (a_q + (uint32_t)i * kAStep + (type ? 6u : 2u));
                const uint32_t boff = (uint32_t)i * kBStep, dcol = tmem_base + (uint32_t)(type * 64);
                mma_fill(dcol, ad, ((uint64_t)b_hi << 32) | (b_kd[0] + boff), idesc);
                mma_use(dcol + 128u, ad, ((uint64_t)b_hi << 32) | (b_kd[1] + boff), idesc);
                mma_last(dcol + 256u, ad, ((uint64_t)b_hi << 32) | (b_kd[2] + boff), idesc);
              }
            }
          } else {
            for (int i = 0; i < kTileH / 2; ++i) {
#pragma unroll
              for (int type = 0; type < 2; ++type) {
                const uint64_t ad = ((uint64_t)a_hi << 32) | (a_q + (uint32_t)i * kAStep + (type ? 6u : 2u));
                const uint32_t boff = (uint32_t)i * kBStep, dcol = tmem_base + (uint32_t)(type * 64);
#pragma unroll
                for (int kd = 0; kd < 3; ++kd)
                  if (kd >= kd_lo && kd <= kd_hi) mma_plain(dcol + (uint32_t)kd * 128u, ad, ((uint64_t)b_hi << 32) | (b_kd[kd] + boff), idesc);
              }
            }
          }
          tc::umma_commit(q_empty(qst));
          if (q - 1 >= 0 && q - 1 < planes) tc::umma_commit(p_empty((p_base + q - 1) & (kPStages - 1)));
          // a segment that ends at the last plane of the volume has no X plane beyond it: release its last dY plane here
          if (q == q_hi && q_hi < planes) tc::umma_commit(p_empty((p_base + planes - 1) & (kPStages - 1)));
        }
        __syncwarp();
        if (++qst == kQStages) { qst = 0; qph ^= 1u; }
      }
      u += planes;
    }
    if (tc::elect_one()) tc::umma_commit(acc_done);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: zero the accumulators, then TMEM -> partial dW =====================
    // accumulator row m = g*32 + p*16 + ci is TMEM lane m: warp ew holds halo-row offset g = ew, lane = p*16 + ci
    const int ew = warp - 4;
    const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
#pragma unroll
    for (int c = 0; c < 384; c += 16) tmem_st16_zero(lane_base + (uint32_t)c);
    tmem_st_wait();
    tc::tc_fence_before();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(acc_zero);
    tc::mbar_wait(acc_done, 0);
    tc::tc_fence_after();
    const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    float* out = g.partial + cta * (int64_t)kPartialFloats;
    const int p = lane >> 4, ci = lane & 15;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
      for (int gp = 0; gp < 2; ++gp) {          // g' = tile row of the pair; kh = g - g'
        uint32_t t0p0[16], t0p1[16], t1p0[16], t1p1[16];
        const uint32_t col = (uint32_t)(kd * 128 + gp * 32);
        tc::tmem_ld16(lane_base + col, t0p0);
        tc::tmem_ld16(lane_base + col + 16, t0p1);
        tc::tmem_ld16(lane_base + col + 64, t1p0);
        tc::tmem_ld16(lane_base + col + 80, t1p1);
        tc::tmem_ld_wait();
        // kw = p - p' (type 0), 2 + p - p' (type 1):  kw 0 <- T0(p,p);  kw 1 <- p ? T0(1,0) : T1(0,1);  kw 2 <- T1(p,p)
        float s[3][16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float v0 = __uint_as_float(p ? t0p1[c] : t0p0[c]);
          const float v1 = __uint_as_float(p ? t0p0[c] : t1p1[c]);
          const float v2 = __uint_as_float(p ? t1p1[c] : t1p0[c]);
          const float o0 = __shfl_xor_sync(0xffffffffu, v0, 16), o1 = __shfl_xor_sync(0xffffffffu, v1, 16), o2 = __shfl_xor_sync(0xffffffffu, v2, 16);
          // the same order (parity 0 first) on both lanes of a pair: identical, deterministic sums
          s[0][c] = p ? o0 + v0 : v0 + o0;
          s[1][c] = p ? o1 + v1 : v1 + o1;
          s[2][c] = p ? o2 + v2 : v2 + o2;
        }
        const int kh = ew - gp;
        if (kh >= 0 && kh <= 2) {
          float* slice = out + gp * kSliceFloats;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int tap = kd * 9 + kh * 3 + kw;
            float4* dst = reinterpret_cast<float4*>(slice + ((int64_t)ci * 27 + tap) * 16 + p * 8);   // lane p stores co [8p, 8p + 8)
            dst[0] = p ? make_float4(s[kw][8], s[kw][9], s[kw][10], s[kw][11]) : make_float4(s[kw][0], s[kw][1], s[kw][2], s[kw][3]);
            dst[1] = p ? make_float4(s[kw][12], s[kw][13], s[kw][14], s[kw][15]) : make_float4(s[kw][4], s[kw][5], s[kw][6], s[kw][7]);
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, kTmemCols);
}

// partial[qslab][spatial][slice 2][ci 16][tap 27][co 16] -> dw[co][ci][27]; the two slices of a CTA are consecutive partials
struct Wg4Map {
  int Cin;
  __device__ int64_t operator()(int group, int64_t e) const {
    const int col = (int)(e % 16), tap = (int)((e / 16) % 27), cl = (int)(e / (16 * 27));
    const int ci = group * 16 + cl;
    return ((int64_t)col * Cin + ci) * 27 + tap;
  }
};

// [N, D, H, W, 16] bf16 read as [N*D][H][W/2][32]: box = 32 elements (one voxel pair) x box_pairs x box_h x 1, SWIZZLE_64B
int make_pair_map(CUtensorMap* tm, const void* base, int N, int D, int H, int W, int box_pairs, int box_h) {
  tma::EncodeTiledFn enc = tma::get_encode();
  B200_REQUIRE(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[4] = {32, (cuuint64_t)(W / 2), (cuuint64_t)H, (cuuint64_t)N * D};
  cuuint64_t gstr[3] = {64, (cuuint64_t)W * 32, (cuuint64_t)H * W * 32};
  cuuint32_t box[4] = {32, (cuuint32_t)box_pairs, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(voxel pairs) failed (%d) for N=%d D=%d H=%d W=%d", (int)r, N, D, H, W);
  return B200_OK;
}

int g_wg4_on = -1;      // -1: read B200_WGRAD_PAIR on first use
int g_wg4_dseg = 0;     // forced d-run (tests), 0 = planned
struct Wg4Plan { int qslabs, tiles_w, tiles_h, spatial; long long units, upc; };

Wg4Plan make_plan(int c0, int c1, int N, int D, int H, int W) {
  Wg4Plan pl;
  pl.qslabs = (c0 + c1) / 16;
  pl.tiles_w = (W + kTileW - 1) / kTileW;
  pl.tiles_h = (H + kTileH - 1) / kTileH;
  pl.units = (long long)pl.tiles_w * pl.tiles_h * N * D;
  // one CTA per SM (512 TMEM columns): the SM budget is shared out among the X slabs and every CTA of a slab takes an equal
  // contiguous run of plane-tiles (first version: whole d-blocks of whole tile columns, whatever CTA count that gave)
  // SMs this kernel fills (env B200_WG4_SMS).  It runs on the side stream next to the main chain's data-gradient kernels, a CTA owns
  // all 512 TMEM columns of its SM, and nothing waits for a weight gradient before the bucket all-reduce / the optimiser: SMs left
  // to the main chain shorten the critical path.  Measured ms/step of the 2 x 128^3 train step (two boxes, +-0.02 between boxes):
  // 148 SMs 3.784, 136: 3.787, 128: 3.771 / 3.763, 120: 3.767, 112: 3.727, 104: 3.716 | 96: 3.657, 80: 3.800, 64: 3.744 | 100: 3.679,
  // 96: 3.685, 88: 3.762 (tools/time_step.py) — the kernel alone is fastest on 148 (167 vs 180 us at 128, 230 at 96)
  static int sm_budget = -1;
  if (sm_budget < 0) { const char* e = getenv("B200_WG4_SMS"); sm_budget = e ? atoi(e) : 96; if (sm_budget < 1 || sm_budget > B200_NUM_SMS) sm_budget = B200_NUM_SMS; }
  long long ctas = sm_budget / pl.qslabs;
  if (ctas < 1) ctas = 1;
  if (ctas > pl.units) ctas = pl.units;
  pl.upc = (pl.units + ctas - 1) / ctas;
  if (g_wg4_dseg > 0) pl.upc = g_wg4_dseg;      // tests: short runs (many CTAs, segments that cross tile columns)
  pl.spatial = (int)((pl.units + pl.upc - 1) / pl.upc);
  return pl;
}

int wg4_enabled() {
  if (g_wg4_on < 0) { const char* e = getenv("B200_WGRAD_PAIR"); g_wg4_on = e ? atoi(e) : 1; }
  return g_wg4_on;
}

}  // namespace

extern "C" int b200_set_wgrad_pair(int on, int dseg) {
  B200_REQUIRE(dseg >= 0, B200_ERR_UNSUPPORTED, "set_wgrad_pair: dseg must be 0 (planned) or positive");
  g_wg4_on = on ? 1 : 0;
  g_wg4_dseg = dseg;
  return B200_OK;
}

// every tensor must be exactly 16 channels wide (a voxel pair is then 64 contiguous bytes) and W even
bool b200_conv3d_wgrad_tc4_supported(int c0, int c1, int Cout, int N, int D, int H, int W) {
  if (!wg4_enabled()) return false;
  if (c0 != 16 || (c1 != 0 && c1 != 16) || Cout != 16) return false;
  return N > 0 && D > 0 && H > 0 && W > 0 && W % 2 == 0 && (int64_t)N * D < (1ll << 31);
}

int64_t b200_conv3d_wgrad_tc4_workspace(int c0, int c1, int Cout, int N, int D, int H, int W) {
  (void)Cout;
  const Wg4Plan pl = make_plan(c0, c1, N, D, H, W);
  return (int64_t)pl.spatial * pl.qslabs * kPartialFloats * 4;
}

int b200_conv3d_wgrad_tc4(const void* x0, int c0, const void* x1, int c1, const void* dy, int Cout, float* dw, void* workspace, int N, int D,
                          int H, int W, cudaStream_t stream) {
  B200_REQUIRE(b200_conv3d_wgrad_tc4_supported(c0, c1, Cout, N, D, H, W), B200_ERR_UNSUPPORTED, "conv3d_wgrad(tcgen05, voxel pairs): unsupported problem");
  const Wg4Plan pl = make_plan(c0, c1, N, D, H, W);
  Wg4Params g;
  g.partial = (float*)workspace;
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.tiles_w = pl.tiles_w; g.tiles_h = pl.tiles_h; g.units = pl.units; g.upc = pl.upc;
  { const char* e = getenv("B200_WG4_SKIP"); g.skip = e ? atoi(e) : 0; }
  CUtensorMap tm_p, tm_q0, tm_q1;
  int rc = make_pair_map(&tm_p, dy, N, D, H, W, kTileW / 2, kTileH);
  if (rc) return rc;
  rc = make_pair_map(&tm_q0, x0, N, D, H, W, kQPairs, 18);
  if (rc) return rc;
  if (c1) { rc = make_pair_map(&tm_q1, x1, N, D, H, W, kQPairs, 18); if (rc) return rc; } else tm_q1 = tm_q0;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(wgrad_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  dim3 grid((unsigned)pl.spatial, 1, (unsigned)pl.qslabs);
  wgrad_tc4_kernel<<<grid, kThreads, kSmemBytes, stream>>>(g, tm_p, tm_q0, tm_q1);
  B200_CHECK_LAUNCH("conv3d_wgrad_tc4");
  B200_CUDA(launch_partial_reduce((const float*)workspace, pl.spatial * 2, (int64_t)kSliceFloats, pl.qslabs, Wg4Map{c0 + c1}, dw, stream));
  b200_count_launch();
  return B200_OK;
}
