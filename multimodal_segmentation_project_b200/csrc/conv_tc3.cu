// conv_tc3.cu — persistent version of the tcgen05 / TMEM implicit-GEMM 3x3x3 convolution of
// conv_tc2.cu (same math, same packed weights, same TMA stages and kd-fused MMAs) for layers with many tiles:
//   * one CTA per SM loops over output tiles (16 x 16 x DSEG voxels x n_tile channels); the TMA producer
//     streams (plane, slab) stages across tile boundaries, so the pipeline never drains between tiles;
//   * TMEM holds TWO accumulator sets (2 x 256 columns): the epilogue of tile k (TMEM -> +bias -> bf16 ->
//     global) overlaps the MMAs of tile k+1 (acc_full / acc_empty mbarriers with phase bits);
//   * when all weight slabs of the layer fit in shared memory they are loaded once per CTA
//     (weight-stationary) instead of once per tile.
// ncu on the non-persistent kernel showed ~73 us of un-overlapped fill / drain in a 126 us launch for the
// 16->16 layer at 2x128^3 (MMA-repeat experiment, profiles/r01_ncu_conv_wgrad_summary.md).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_maps.cuh"
#include <stdlib.h>

// Build with `make EXTRA=-DB200_TC_DEBUG` for per-role cycle counters (env B200_TC_DEBUG=1 prints them) and for the
// bottleneck knobs of env B200_TC3_SKIP (bit 0: no MMAs, 1: no epilogue global stores, 2: no epilogue TMEM loads,
// 3: no TMA activation loads after the first ring fill).  The default build contains none of this.
// What these knobs found (profiles/r02_tc3_bottleneck.md): the kernel was paced by the MMA-issuing warp's own
// per-stage bookkeeping, not by the tensor core.
#ifdef B200_TC_DEBUG
#define TC3_DBG(...) __VA_ARGS__
#define TC3_CLOCK() clock64()
#else
#define TC3_DBG(...)
#define TC3_CLOCK() 0ll
#endif

namespace {

using bf16 = __nv_bfloat16;

constexpr int kHalo = 18;
constexpr int kPlaneVox = kHalo * kHalo;
constexpr int kStageBytes = 10496;                  // 324 voxels x 32 B rounded up to the 256-byte swizzle period
constexpr int kStageTx = kPlaneVox * 32;
constexpr int kMaxStages = 12;
constexpr int kThreads = 384;                       // w0: act TMA, w1: weight TMA + TMEM alloc, w2/w3: MMA issue (one per w-tile), w4-11: epilogue
constexpr int kMaxDseg = 8;
constexpr int kSmemHeader = 1024;
constexpr int kSetCols = 256;                       // TMEM columns per accumulator set

struct Tc3Params {
  const uint8_t* wpack; const float* bias;
  bf16* y0; bf16* y1; int co0, co1;
  int c0, c1;
  int N, D, H, W;
  int n_tile, nchunks, dseg, dblocks, slabs, tiles_w, tiles_h, stages;
  int kd_per_mma;
  int wstationary;   // all slabs resident in shared memory (requires nchunks == 1)
  int wstages;       // streaming mode: ring depth (1 or 2)
  int total_tiles;
  float* stats;             // STATS_CH kernels: per-CTA BatchNorm partial sums [grid][2][Cout] (sum, sum of squares of y - bias)
  unsigned long long* dbg;  // optional per-CTA cycle counters (B200_TC_DEBUG builds): [cta][8]
  int skip;                 // B200_TC_DEBUG builds: bottleneck knobs (see top of file)
};

struct Tile { int tw, th, n, db, nchunk; };
__device__ __forceinline__ Tile decode_tile(const Tc3Params& p, int t) {
  Tile r;
  r.tw = t % p.tiles_w; t /= p.tiles_w;
  r.th = t % p.tiles_h; t /= p.tiles_h;
  r.db = t % p.dblocks; t /= p.dblocks;
  r.n = t % p.N;
  r.nchunk = t / p.N;
  return r;
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

// STATS_CH = 0: plain convolution.  STATS_CH = 16 / 32 (= Cout): the epilogue also accumulates, per output channel, the sum
// and the sum of squares of (y - bias) over the voxels it stores — y being the bf16-ROUNDED value the next kernel will read,
// which is what the reference's BatchNorm sees (models/unet.py:12,16 under autocast) — and the CTA writes one partial row in
// the layout of bn_stats (elementwise_kernels.cu).  This removes one full read of the activation per layer.
template <int STATS_CH>
__global__ void __launch_bounds__(kThreads, 1)
conv3d_tc3_kernel(const Tc3Params p, const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // barrier slots: [0,12) a_full, [12,24) a_empty, [24,26) w_full, [26,28) w_empty, [28,44) acc_full[set][plane], [44,46) acc_empty[set]
  const uint32_t bar0 = tc::smem_u32(bars);
  auto a_full = [&](int i) { return bar0 + 8u * i; };
  auto a_empty = [&](int i) { return bar0 + 8u * (12 + i); };
  auto w_full = [&](int i) { return bar0 + 8u * (24 + i); };
  auto w_empty = [&](int i) { return bar0 + 8u * (26 + i); };
  auto acc_full = [&](int set, int pl) { return bar0 + 8u * (28 + set * 8 + pl); };
  auto acc_empty = [&](int set) { return bar0 + 8u * (44 + set); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 384);
  float* bias_s = reinterpret_cast<float*>(smem + 512);   // [n_tile] (re-filled per tile when nchunks > 1)
  uint8_t* act = smem + kSmemHeader;
  uint8_t* wts = act + p.stages * kStageBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wbytes = 864u * p.n_tile;
  const int my_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 2 && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) { tc::mbar_init(a_full(i), 1); tc::mbar_init(a_empty(i), 2); }   // a_empty: both issuing warps
    for (int i = 0; i < 2; ++i) { tc::mbar_init(w_full(i), 1); tc::mbar_init(w_empty(i), 2); }
    for (int s = 0; s < 2; ++s) {
      for (int i = 0; i < kMaxDseg; ++i) tc::mbar_init(acc_full(s, i), 2);   // both issuing warps commit
      tc::mbar_init(acc_empty(s), 8);  // one arrival per epilogue warp
    }
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
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // shuffle from a constant lane: provably warp-uniform (uniform registers)

  if (warp == 0) {
    // ===================== activation stages by TMA, continuous across tiles =====================
    {
      // ring position / phase kept as counters: a division by the runtime ring depth is a ~25-instruction dependent chain,
      // and this warp has no other warp to hide it behind
      int st = 0;
      uint32_t ph = 1;          // producer waits for the "previous" phase of a_empty: passes immediately on the first lap
      TC3_DBG(int it = 0; long long dbg_prod_wait = 0; const long long tstart = clock64();)
      for (int k = 0; k < my_tiles; ++k) {
        const Tile t = decode_tile(p, blockIdx.x + k * gridDim.x);
        const int w0 = t.tw * 16, h0 = t.th * 16, d0 = t.db * p.dseg;
        const int planes = min(p.dseg, p.D - d0);
        for (int s = 0; s < p.slabs; ++s) {
          const int c = s * 16;
          const CUtensorMap* tm = c < p.c0 ? &tm0 : &tm1;
          const int cc = c < p.c0 ? c : c - p.c0;
          for (int q = -1; q <= planes; ++q) {
            TC3_DBG(const long long t0 = clock64();)
            tc::mbar_wait(a_empty(st), ph);
            TC3_DBG(dbg_prod_wait += clock64() - t0;)
            if (tc::elect_one()) {
#ifdef B200_TC_DEBUG
              if ((p.skip & 8) && it >= p.stages) tc::mbar_arrive(a_full(st));
              else
#endif
              {
                tc::mbar_arrive_expect_tx(a_full(st), kStageTx);
                tma::load_5d(tc::smem_u32(act + st * kStageBytes), tm, cc, w0 - 1, h0 - 1, d0 + q, t.n, a_full(st));
              }
            }
            __syncwarp();
            TC3_DBG(++it;)
            if (++st == p.stages) { st = 0; ph ^= 1u; }
          }
        }
      }
      TC3_DBG(if (p.dbg && lane == 0) { p.dbg[blockIdx.x * 8 + 0] = dbg_prod_wait; p.dbg[blockIdx.x * 8 + 1] = clock64() - tstart; })
    }
  } else if (warp == 1) {
    // ===================== weights =====================
    if (lane == 0) {
      if (p.wstationary) {
        tc::mbar_arrive_expect_tx(w_full(0), wbytes * p.slabs);
        for (int s = 0; s < p.slabs; ++s)
          tc::bulk_g2s(tc::smem_u32(wts + (size_t)s * wbytes), p.wpack + (size_t)s * wbytes, wbytes, w_full(0));
      } else {
        int wi = 0;
        for (int k = 0; k < my_tiles; ++k) {
          const Tile t = decode_tile(p, blockIdx.x + k * gridDim.x);
          for (int s = 0; s < p.slabs; ++s, ++wi) {
            const int ws = p.wstages == 2 ? (wi & 1) : 0;
            tc::mbar_wait(w_empty(ws), (((p.wstages == 2 ? (wi >> 1) : wi) & 1) ^ 1));
            tc::mbar_arrive_expect_tx(w_full(ws), wbytes);
            tc::bulk_g2s(tc::smem_u32(wts + (size_t)ws * wbytes), p.wpack + ((size_t)t.nchunk * p.slabs + s) * wbytes, wbytes, w_full(ws));
          }
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ===================== MMA issue: TWO issuing warps, one per w-tile (accumulator columns [wt * wt_cols, +wt_cols)) ===========
    // tcgen05.mma issue blocks once a few instructions are queued, so whatever the issuing warp does between two stages
    // (barrier wait, fence, descriptor arithmetic: ~300 exposed-latency cycles) leaves the tensor core idle.  With the two
    // w-tiles issued by two warps, one warp's bookkeeping overlaps the other warp's MMAs (measured on the 16 -> 16 layer at
    // 2 x 128^3 with the MMAs as the only work: 1 issuer 106 us, 2 issuers 90 us; forcing the two warps half a stage apart
    // with a handshake was slower again: 100 us).  Both warps wait on the same stage barriers and both commit:
    // a_empty / w_empty / acc_full count two arrivals.  The whole warp runs the loop, one elected lane issues.
    const int wt = warp - 2;
    const uint32_t n_t = p.n_tile;
    const uint32_t b_lbo = 48u * n_t, b_tap16 = 6u * n_t;
    const uint64_t a_proto = desc_kmajor_sw32(0, kHalo * 32);
    const uint64_t b_proto = tc::smem_desc_kmajor_noswz(0, b_lbo, 128);
    const uint32_t a_hi = (uint32_t)(a_proto >> 32);
    const uint32_t a_lo0 = (uint32_t)a_proto + (tc::smem_u32(act) >> 4) + (uint32_t)wt * 16u;   // + 8 voxels for the second w-tile
    const uint32_t b_hi = (uint32_t)(b_proto >> 32), b_lo0 = (uint32_t)b_proto;
    const uint32_t idesc0 = tc::idesc_bf16_f32(128, 0);       // N field added per MMA: (N >> 3) << 17
    const uint32_t idesc_step = (n_t >> 3) << 17;             // one more kd block
    const uint32_t wt_cols = (uint32_t)p.dseg * n_t;
    if (p.wstationary) { tc::mbar_wait(w_full(0), 0); tc::tc_fence_after(); }
    int st = 0, wi = 0;
    uint32_t ph = 0;
    TC3_DBG(long long dbg_wait_full = 0, dbg_wait_acc = 0; const long long tstart = clock64();)
    for (int k = 0; k < my_tiles; ++k) {
      const Tile t = decode_tile(p, blockIdx.x + k * gridDim.x);
      const int w0 = t.tw * 16, d0 = t.db * p.dseg;
      const int planes = min(p.dseg, p.D - d0);
      const bool active = w0 + wt * 8 < p.W TC3_DBG(&& !(p.skip & 1));   // a w-tile entirely outside the volume only keeps the barriers moving
      const int set = k & 1;
      const uint32_t tset = tmem_base + (uint32_t)set * kSetCols + (uint32_t)wt * wt_cols;
      TC3_DBG(const long long ta = clock64();)
      tc::mbar_wait(acc_empty(set), ((k >> 1) & 1) ^ 1);   // epilogue of tile k-2 has drained this set
      TC3_DBG(dbg_wait_acc += clock64() - ta;)
      tc::tc_fence_after();
      for (int s = 0; s < p.slabs; ++s) {
        uint32_t w_lo;
        int ws = 0;
        if (p.wstationary) {
          w_lo = b_lo0 + (tc::smem_u32(wts + (size_t)s * wbytes) >> 4);
        } else {
          ws = p.wstages == 2 ? (wi & 1) : 0;
          tc::mbar_wait(w_full(ws), (p.wstages == 2 ? (wi >> 1) : wi) & 1);
          tc::tc_fence_after();
          w_lo = b_lo0 + (tc::smem_u32(wts + (size_t)ws * wbytes) >> 4);
          ++wi;
        }
        const bool last_slab = s == p.slabs - 1;
        for (int q = -1; q <= planes; ++q) {
          TC3_DBG(const long long tf = clock64();)
          tc::mbar_wait(a_full(st), ph);
          TC3_DBG(dbg_wait_full += clock64() - tf;)
          tc::tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)st * (kStageBytes >> 4);
          const int kd_lo = max(0, q + 2 - planes), kd_hi = min(2, q + 1);
          const bool first = (s == 0 && kd_lo == 0);
          if (tc::elect_one()) {
            if (active) {
              if (first) {
                // tap (0,0), kd = 0: the plane's very first contribution overwrites its accumulator
                const uint32_t col = tset + (uint32_t)(p.dseg - 2 - q) * n_t;
                tc::umma_bf16_ss(col, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | w_lo, idesc0 + idesc_step, 0);
                if (kd_hi >= 1)
                  tc::umma_bf16_ss(col + n_t, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | (w_lo + n_t),
                                   idesc0 + idesc_step * (uint32_t)kd_hi, 1);
              }
              for (int a = kd_lo; a <= kd_hi; a += p.kd_per_mma) {
                const int cnt = min(kd_hi, a + p.kd_per_mma - 1) - a + 1;
                const uint32_t col = tset + (uint32_t)(p.dseg - 2 - q + a) * n_t;
                const uint32_t idesc = idesc0 + idesc_step * (uint32_t)cnt;
                const uint32_t b_lo_g = w_lo + (uint32_t)a * n_t;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                  if (first && tap == 0) continue;
                  const uint32_t a_off = (uint32_t)(((tap / 3) * kHalo + (tap % 3)) * 32) >> 4;
                  tc::umma_bf16_ss(col, ((uint64_t)a_hi << 32) | (a_lo + a_off), ((uint64_t)b_hi << 32) | (b_lo_g + (uint32_t)tap * b_tap16), idesc, 1);
                }
              }
            }
            tc::umma_commit(a_empty(st));
            if (last_slab && q >= 1) tc::umma_commit(acc_full(set, q - 1));
          }
          __syncwarp();
          if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
        if (!p.wstationary) {
          if (tc::elect_one()) tc::umma_commit(w_empty(ws));
          __syncwarp();
        }
      }
    }
    TC3_DBG(if (p.dbg && lane == 0 && wt == 0) { p.dbg[blockIdx.x * 8 + 2] = dbg_wait_full; p.dbg[blockIdx.x * 8 + 3] = dbg_wait_acc; p.dbg[blockIdx.x * 8 + 4] = clock64() - tstart; })
  } else if (warp >= 4) {
    // ===================== epilogue: warps 4-7 drain w-tile 0, warps 8-11 w-tile 1 =====================
    const int wt = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int m = ew * 32 + lane;
    int bias_chunk = -1;
    float st_sum[STATS_CH > 0 ? STATS_CH : 1], st_sq[STATS_CH > 0 ? STATS_CH : 1];
#pragma unroll
    for (int i = 0; i < (STATS_CH > 0 ? STATS_CH : 1); ++i) st_sum[i] = st_sq[i] = 0.f;
    TC3_DBG(long long dbg_epi_wait = 0; const long long tstart = clock64();)
    uint32_t full_phase = 0;  // bit (set*8 + plane): parity of the next completion of that acc_full barrier (tiles may have < dseg planes)
    for (int k = 0; k < my_tiles; ++k) {
      const Tile t = decode_tile(p, blockIdx.x + k * gridDim.x);
      const int w0 = t.tw * 16, h0 = t.th * 16, d0 = t.db * p.dseg;
      const int planes = min(p.dseg, p.D - d0);
      const int set = k & 1;
      if (t.nchunk != bias_chunk) {
        // all 8 epilogue warps refresh the bias slice together (named barrier 1, 256 threads)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int e = threadIdx.x - 128;
        if (e < p.n_tile) bias_s[e] = p.bias ? p.bias[t.nchunk * p.n_tile + e] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        bias_chunk = t.nchunk;
      }
      const int h = h0 + (m >> 3);
      const int w = w0 + wt * 8 + (m & 7);
      const bool wt_ok = w0 + wt * 8 < p.W;
      const bool valid = h < p.H && w < p.W;
      for (int pl = 0; pl < planes; ++pl) {
        TC3_DBG(const long long te = clock64();)
        tc::mbar_wait(acc_full(set, pl), (full_phase >> (set * 8 + pl)) & 1u);
        TC3_DBG(dbg_epi_wait += clock64() - te;)
        full_phase ^= 1u << (set * 8 + pl);
        if (!wt_ok) continue;
        TC3_DBG(if (p.skip & 4) continue;)
        tc::tc_fence_after();
        const int64_t row = (((int64_t)t.n * p.D + d0 + pl) * p.H + h) * p.W + w;
        const uint32_t col0 = (uint32_t)(set * kSetCols + (wt * p.dseg + (p.dseg - 1 - pl)) * p.n_tile);
#pragma unroll
        for (int cc = 0; cc < (STATS_CH > 0 ? STATS_CH / 16 : 8); ++cc) {
          if (STATS_CH == 0 && cc >= p.n_tile / 16) break;
          uint32_t r[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + col0 + cc * 16, r);
          tc::tmem_ld_wait();
          const int ch = t.nchunk * p.n_tile + cc * 16;
          uint32_t packed[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 bb = *reinterpret_cast<const float2*>(bias_s + cc * 16 + 2 * i);
            __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[2 * i]) + bb.x, __uint_as_float(r[2 * i + 1]) + bb.y);
            packed[i] = *reinterpret_cast<uint32_t*>(&hb);
            if (STATS_CH > 0 && valid) {
              const float2 yv = __bfloat1622float2(hb);
              const float d0 = yv.x - bb.x, d1 = yv.y - bb.y;
              const int c = (STATS_CH > 0 ? cc * 16 + 2 * i : 0);
              st_sum[c] += d0;
              st_sq[c] = fmaf(d0, d0, st_sq[c]);
              st_sum[(STATS_CH > 0 ? c + 1 : 0)] += d1;
              st_sq[(STATS_CH > 0 ? c + 1 : 0)] = fmaf(d1, d1, st_sq[(STATS_CH > 0 ? c + 1 : 0)]);
            }
          }
          if (valid TC3_DBG(&& !(p.skip & 2))) {
            bf16* dst = ch < p.co0 ? p.y0 + row * p.co0 + ch : p.y1 + row * p.co1 + (ch - p.co0);
            reinterpret_cast<uint4*>(dst)[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            reinterpret_cast<uint4*>(dst)[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
        }
      }
      // this warp's TMEM reads of the set are complete (tcgen05.wait::ld above): hand the set back to the MMA warp
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(acc_empty(set));
    }
    TC3_DBG(if (p.dbg && warp == 4 && lane == 0) { p.dbg[blockIdx.x * 8 + 5] = dbg_epi_wait; p.dbg[blockIdx.x * 8 + 6] = clock64() - tstart; })
    if (STATS_CH > 0) {
      // fixed-shape fold: warp butterflies, then the eight epilogue warps in order -> run-to-run deterministic.  The stage ring
      // is free by now (every accumulator this CTA waited for is complete, so every stage has been consumed): reuse stage 0.
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
        p.stats[(size_t)blockIdx.x * 2 * STATS_CH + e] = tot;
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

inline int n_tile_for(int cout) { return cout <= 128 ? cout : 128; }

}  // namespace

// decides whether the persistent kernel is the right tool; fills nothing
bool b200_conv3d_k3_tc3_wanted(int c0, int c1, int co0, int co1, int N, int D, int H, int W) {
  const int cout = co0 + co1, n_tile = n_tile_for(cout);
  if (n_tile > 64) return false;                        // deep / wide layers: few tiles, huge weight slabs
  int dseg = kSetCols / (2 * n_tile);
  if (dseg > kMaxDseg) dseg = kMaxDseg;
  if (dseg > D) dseg = D;
  const int64_t tiles = (int64_t)((W + 15) / 16) * ((H + 15) / 16) * ((D + dseg - 1) / dseg) * N * (cout / n_tile);
  return tiles >= 2 * B200_NUM_SMS;
}

// number of per-CTA BatchNorm partial rows the fused-statistics variant writes for this problem (= its grid), 0 when the
// fused variant does not apply (the caller then runs bn_stats on the output as before)
int b200_conv3d_k3_tc3_stats_blocks(int c0, int c1, int co0, int co1, int N, int D, int H, int W) {
  const int cout = co0 + co1;
  if (co1 != 0 || (cout != 16 && cout != 32)) return 0;
  if (!b200_conv3d_k3_tc3_wanted(c0, c1, co0, co1, N, D, H, W)) return 0;
  int dseg = kSetCols / (2 * cout);
  if (dseg > kMaxDseg) dseg = kMaxDseg;
  if (dseg > D) dseg = D;
  const int64_t tiles = (int64_t)((W + 15) / 16) * ((H + 15) / 16) * ((D + dseg - 1) / dseg) * N;
  return (int)(tiles < B200_NUM_SMS ? tiles : B200_NUM_SMS);
}

int b200_conv3d_k3_tc3(const void* x0, int c0, const void* x1, int c1, const void* wpack, const float* bias, void* y0, int co0, void* y1,
                       int co1, int N, int D, int H, int W, cudaStream_t stream, float* stats) {
  B200_REQUIRE(b200_aligned(x0, 16) && b200_aligned(x1, 16) && b200_aligned(y0, 16) && b200_aligned(y1, 16) && b200_aligned(wpack, 16),
               B200_ERR_ALIGN, "conv3d_k3(tcgen05): pointers must be 16-byte aligned");
  Tc3Params p;
  p.wpack = (const uint8_t*)wpack; p.bias = bias;
  p.y0 = (bf16*)y0; p.y1 = (bf16*)y1; p.co0 = co0; p.co1 = co1;
  p.c0 = c0; p.c1 = c1;
  p.N = N; p.D = D; p.H = H; p.W = W;
  const int cout = co0 + co1;
  p.n_tile = n_tile_for(cout);
  p.nchunks = cout / p.n_tile;
  p.slabs = (c0 + c1) / 16;
  p.kd_per_mma = 3 * p.n_tile <= 256 ? 3 : 2;
  int dseg = kSetCols / (2 * p.n_tile);
  if (dseg > kMaxDseg) dseg = kMaxDseg;
  if (dseg < 1) dseg = 1;
  if (dseg > D) dseg = D;
  p.dseg = dseg;
  p.dblocks = (D + dseg - 1) / dseg;
  p.tiles_w = (W + 15) / 16;
  p.tiles_h = (H + 15) / 16;
  p.total_tiles = p.tiles_w * p.tiles_h * p.dblocks * N * p.nchunks;
  const size_t wbytes = (size_t)864 * p.n_tile;
  const size_t budget = 200 * 1024 - kSmemHeader - 1024;
  p.wstationary = (p.nchunks == 1 && wbytes * p.slabs <= 112 * 1024) ? 1 : 0;
  p.wstages = 1;
  size_t wtotal;
  if (p.wstationary) wtotal = wbytes * p.slabs;
  else { p.wstages = (2 * wbytes + 6 * kStageBytes <= budget) ? 2 : 1; wtotal = wbytes * p.wstages; }
  int stages = (int)((budget - wtotal) / kStageBytes);
  if (stages > kMaxStages) stages = kMaxStages;
  B200_REQUIRE(stages >= 3, B200_ERR_UNSUPPORTED, "conv3d_k3(tcgen05 persistent): not enough shared memory");
  p.stages = stages;
  const size_t smem = kSmemHeader + (size_t)stages * kStageBytes + wtotal + 1024;
  CUtensorMap tm0, tm1;
  int rc = tma::make_ndhwc_map(&tm0, x0, c0, N, D, H, W, kHalo, kHalo);
  if (rc) return rc;
  if (c1) { rc = tma::make_ndhwc_map(&tm1, x1, c1, N, D, H, W, kHalo, kHalo); if (rc) return rc; } else tm1 = tm0;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(conv3d_tc3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
    B200_CUDA(cudaFuncSetAttribute(conv3d_tc3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
    B200_CUDA(cudaFuncSetAttribute(conv3d_tc3_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
    attr_set = true;
  }
  p.stats = stats;
  B200_REQUIRE(!stats || (p.nchunks == 1 && co1 == 0 && (cout == 16 || cout == 32)), B200_ERR_UNSUPPORTED,
               "conv3d_k3(tcgen05 persistent): fused BatchNorm statistics need Cout = 16 or 32 in one tensor");
  const int grid = p.total_tiles < B200_NUM_SMS ? p.total_tiles : B200_NUM_SMS;
  p.dbg = nullptr;
  p.skip = 0;
#ifdef B200_TC_DEBUG
  static unsigned long long* dbg_buf = nullptr;
  if (getenv("B200_TC_DEBUG")) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, sizeof(unsigned long long) * 8 * B200_NUM_SMS);
    p.dbg = dbg_buf;
  }
  { const char* e = getenv("B200_TC3_SKIP"); p.skip = e ? atoi(e) : 0; }
#endif
  if (stats && cout == 16) conv3d_tc3_kernel<16><<<grid, kThreads, smem, stream>>>(p, tm0, tm1);
  else if (stats) conv3d_tc3_kernel<32><<<grid, kThreads, smem, stream>>>(p, tm0, tm1);
  else conv3d_tc3_kernel<0><<<grid, kThreads, smem, stream>>>(p, tm0, tm1);
  B200_CHECK_LAUNCH("conv3d_k3_tc3");
#ifdef B200_TC_DEBUG
  if (p.dbg) {
    cudaStreamSynchronize(stream);
    static unsigned long long h[8 * B200_NUM_SMS];
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    double a[8] = {0};
    for (int i = 0; i < grid; ++i) for (int j = 0; j < 8; ++j) a[j] += (double)h[i * 8 + j] / grid;
    fprintf(stderr, "[tc3 dbg] per-CTA avg cycles: producer wait a_empty %.0f of %.0f | mma wait a_full %.0f, wait acc_empty %.0f of %.0f | epilogue(w4) wait acc_full %.0f of %.0f | tiles/CTA %.1f stages %d\n",
            a[0], a[1], a[2], a[3], a[4], a[5], a[6], (double)p.total_tiles / grid, p.stages);
  }
#endif
  return B200_OK;
}
