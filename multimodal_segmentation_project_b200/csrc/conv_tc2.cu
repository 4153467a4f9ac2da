// conv_tc2.cu — second generation of the tcgen05 / TMEM implicit-GEMM 3x3x3 convolution
// (bf16 in, fp32 accumulate in TMEM, bf16 out).  Same plane-streaming formulation as conv_tc.cu, plus:
//
//   * TMA-fed stages: every (d-plane, 16-channel slab) halo plane is fetched by ONE
//     cp.async.bulk.tensor.5d copy (box 16 ch x 18 x 18) from a SWIZZLE_32B tensor map over the NDHWC
//     activation into [voxel][16 ch] rows of 32 bytes — the K-major SWIZZLE_32B UMMA layout, in which a
//     filter tap (kh,kw) is still just a byte offset of the descriptor start (the swizzle is a function
//     of the absolute shared-memory address, so unaligned 32-byte shifts stay consistent with what TMA
//     wrote).  Out-of-volume coordinates are zero-filled by the TMA unit (that IS the conv padding);
//     completion is signalled on the stage's mbarrier — no producer warps, no registers.
//   * kd-fused MMAs: for a staged input plane and a tap (kh,kw) the A operand (the shifted 128-voxel
//     view) is the same for kd = 0,1,2 — only the output plane differs.  Accumulators of consecutive
//     planes sit in consecutive TMEM columns (descending plane order) and the packed weights keep the
//     three kd blocks adjacent, so ONE tcgen05.mma with N = 3*Cout covers three taps: the 4 KB A tile
//     is read from shared memory once instead of three times (the measured limiter of v1 for
//     Cout <= 32, see profiles/r01_ncu_conv_wgrad_summary.md).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace {

using bf16 = __nv_bfloat16;

constexpr int kHalo = 18;
constexpr int kPlaneVox = kHalo * kHalo;            // 324
constexpr int kStageBytes = 10496;                  // 324 voxels x 32 B = 10368, rounded up to the 256-byte SWIZZLE_32B period
constexpr int kStageTx = kPlaneVox * 32;            // bytes written per stage by one TMA box
constexpr int kMaxStages = 6;
constexpr int kThreads = 384;                       // w0: act TMA, w1: weight TMA + TMEM alloc, w2: MMA, w4-7 / w8-11: epilogue of w-tile 0 / 1
constexpr int kMaxDseg = 8;
constexpr int kSmemHeader = 1024;                   // barriers, TMEM slot, bias[<=128]

struct Tc2Params {
  const uint8_t* wpack; const float* bias;
  bf16* y0; bf16* y1; int co0, co1;
  int c0, c1;
  int N, D, H, W;
  int n_tile, dseg, dblocks, slabs, wstages, tmem_cols, tiles_w, stages;
  int kd_per_mma;   // 3 when 3*n_tile <= 256, else 2 (n_tile = 128)
  int mma_repeat;   // diagnostics only (B200_TC_REPEAT): issue the fused MMAs this many times
  int ksplit;       // > 1: this CTA reduces only slabs [zs*slabs_per, ...) and writes fp32 partials (deep layers: few tiles, many slabs)
  int slabs_per;
  float* partial;   // [ksplit][rows][cout]
};

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, int c, int w, int h, int d, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "r"(bar)
      : "memory");
}
// K-major SWIZZLE_32B operand: rows of 32 bytes (16 bf16 = one UMMA K step), 8-row groups SBO bytes apart
__device__ __forceinline__ uint64_t desc_kmajor_sw32(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                               // LBO: unused for a single swizzled K block
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                               // version
  d |= (uint64_t)6 << 61;                               // SWIZZLE_32B
  return d;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
conv3d_tc2_kernel(const Tc2Params p, const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // barrier slots: [0,6) a_full, [6,12) a_empty, [12,14) w_full, [14,16) w_empty, [16,24) acc_full
  const uint32_t bar0 = tc::smem_u32(bars);
  auto a_full = [&](int i) { return bar0 + 8u * i; };
  auto a_empty = [&](int i) { return bar0 + 8u * (6 + i); };
  auto w_full = [&](int i) { return bar0 + 8u * (12 + i); };
  auto w_empty = [&](int i) { return bar0 + 8u * (14 + i); };
  auto acc_full = [&](int i) { return bar0 + 8u * (16 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  float* bias_s = reinterpret_cast<float*>(smem + 512);
  uint8_t* act = smem + kSmemHeader;
  uint8_t* wts = act + p.stages * kStageBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = blockIdx.x % p.tiles_w, th = blockIdx.x / p.tiles_w;
  const int n = blockIdx.y / p.dblocks, db = blockIdx.y % p.dblocks;
  const int nchunk = blockIdx.z / p.ksplit, zs = blockIdx.z % p.ksplit;
  const int slab0 = zs * p.slabs_per;
  const int nslabs = min(p.slabs_per, p.slabs - slab0);
  const int w0 = tw * 16, h0 = th * 16, d0 = db * p.dseg;
  const int planes = min(p.dseg, p.D - d0);
  const int nq = planes + 2;
  const int total_stages = nslabs * nq;
  const uint32_t wbytes = 864u * p.n_tile;

  if (warp == 2 && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) { tc::mbar_init(a_full(i), 1); tc::mbar_init(a_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(w_full(i), 1); tc::mbar_init(w_empty(i), 1); }
    for (int i = 0; i < kMaxDseg; ++i) tc::mbar_init(acc_full(i), 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tc::smem_u32(tmem_slot), p.tmem_cols);
    tc::tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm0);
    if (p.c1) prefetch_tmap(&tm1);
  }
  if (warp >= 4 && threadIdx.x - 128 < p.n_tile) bias_s[threadIdx.x - 128] = p.bias ? p.bias[(blockIdx.z / p.ksplit) * p.n_tile + (threadIdx.x - 128)] : 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== activation stages by TMA =====================
    if (lane == 0) {
      // (slab, plane, ring slot, phase) as counters: divisions by run-time values are long dependent chains in a lone thread
      int s = 0, q = -1, st = 0;
      uint32_t ph = 1;
      for (int it = 0; it < total_stages; ++it) {
        tc::mbar_wait(a_empty(st), ph);
        tc::mbar_arrive_expect_tx(a_full(st), kStageTx);
        const int c = (slab0 + s) * 16;
        const CUtensorMap* tm = c < p.c0 ? &tm0 : &tm1;
        const int cc = c < p.c0 ? c : c - p.c0;
        const uint32_t dst = tc::smem_u32(act + st * kStageBytes);
        tma_load_5d(dst, tm, cc, w0 - 1, h0 - 1, d0 + q, n, a_full(st));
        if (++q > planes) { q = -1; ++s; }
        if (++st == p.stages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== weight slabs by TMA bulk copy =====================
    if (lane == 0) {
      for (int s = 0; s < nslabs; ++s) {
        const int ws = p.wstages == 2 ? (s & 1) : 0;
        tc::mbar_wait(w_empty(ws), ((p.wstages == 2 ? (s >> 1) : s) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(w_full(ws), wbytes);
        tc::bulk_g2s(tc::smem_u32(wts + (size_t)ws * wbytes), p.wpack + ((size_t)nchunk * p.slabs + slab0 + s) * wbytes, wbytes, w_full(ws));
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issue =====================
    if (lane == 0) {
      // The issuing thread is a single in-order lane: keep the per-MMA work to two 32-bit adds.  Descriptors are
      // split into a constant high word and a low word whose address field (addr >> 4) is advanced by constants.
      const uint32_t n_t = p.n_tile;
      const uint32_t b_lbo = 48u * n_t;        // K-chunk stride of one (kh,kw) B matrix: 3*n rows x 16 B
      const uint32_t b_tap16 = 6u * n_t;       // (bytes per (kh,kw) = 96*n) >> 4
      const uint64_t a_proto = desc_kmajor_sw32(0, kHalo * 32);
      const uint64_t b_proto = tc::smem_desc_kmajor_noswz(0, b_lbo, 128);
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), a_lo0 = (uint32_t)a_proto;
      const uint32_t b_hi = (uint32_t)(b_proto >> 32), b_lo0 = (uint32_t)b_proto;
      uint32_t idesc_n[4];
#pragma unroll
      for (int i = 1; i <= 3; ++i) idesc_n[i] = tc::idesc_bf16_f32(128, (int)(i * n_t));
      const bool wt1 = w0 + 8 < p.W;
      const uint32_t wt_cols = (uint32_t)p.dseg * n_t;
      int st = 0;
      uint32_t ph = 0;
      const uint32_t act16 = tc::smem_u32(act) >> 4;
      for (int s = 0; s < nslabs; ++s) {
        const int ws = p.wstages == 2 ? (s & 1) : 0;
        tc::mbar_wait(w_full(ws), (p.wstages == 2 ? (s >> 1) : s) & 1);
        tc::tc_fence_after();
        const uint32_t w_lo = b_lo0 + (tc::smem_u32(wts + (size_t)ws * wbytes) >> 4);
        for (int qi = 0; qi < nq; ++qi) {
          tc::mbar_wait(a_full(st), ph);
          tc::tc_fence_after();
          const int q = qi - 1;
          const uint32_t a_lo = a_lo0 + act16 + (uint32_t)st * (kStageBytes >> 4);
          // valid kd for this input plane: output plane pl = q + 1 - kd in [0, planes)
          const int kd_lo = max(0, q + 2 - planes), kd_hi = min(2, q + 1);
          // kd groups of at most kd_per_mma taps; planes q+1-a .. q+1-b occupy ascending TMEM column blocks
          int ga[2], gn[2], ngroups = 0;
          for (int a = kd_lo; a <= kd_hi; a += p.kd_per_mma) { ga[ngroups] = a; gn[ngroups] = min(kd_hi, a + p.kd_per_mma - 1) - a + 1; ++ngroups; }
          const bool first = (s == 0 && kd_lo == 0);  // plane q+1 receives its very first contribution in this stage
          if (first) {
            // tap (0,0): kd = 0 overwrites (accumulate = 0) and therefore cannot be fused with planes holding partial sums
            const uint32_t col = (uint32_t)(p.dseg - 2 - q) * n_t;
            tc::umma_bf16_ss(tmem_base + col, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | w_lo, idesc_n[1], 0);
            if (wt1) tc::umma_bf16_ss(tmem_base + wt_cols + col, ((uint64_t)a_hi << 32) | (a_lo + 16), ((uint64_t)b_hi << 32) | w_lo, idesc_n[1], 0);
            if (kd_hi >= 1) {
              const uint32_t nrem = (uint32_t)kd_hi;  // kd = 1..kd_hi (<= 2 taps: fits one MMA for every n_tile)
              const uint64_t bd = ((uint64_t)b_hi << 32) | (w_lo + n_t);  // skip kd = 0 rows: n rows x 16 B >> 4
              tc::umma_bf16_ss(tmem_base + col + n_t, ((uint64_t)a_hi << 32) | a_lo, bd, idesc_n[nrem], 1);
              if (wt1) tc::umma_bf16_ss(tmem_base + wt_cols + col + n_t, ((uint64_t)a_hi << 32) | (a_lo + 16), bd, idesc_n[nrem], 1);
            }
          }
          for (int rep = 0; rep < p.mma_repeat; ++rep)
          for (int gi = 0; gi < ngroups; ++gi) {
            const uint32_t col = tmem_base + (uint32_t)(p.dseg - 2 - q + ga[gi]) * n_t;
            const uint32_t idesc = idesc_n[gn[gi]];
            const uint32_t b_lo_g = w_lo + (uint32_t)ga[gi] * n_t;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              if (first && tap == 0) continue;
              const uint32_t a_off = (uint32_t)(((tap / 3) * kHalo + (tap % 3)) * 32) >> 4;
              const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo_g + (uint32_t)tap * b_tap16);
              tc::umma_bf16_ss(col, ((uint64_t)a_hi << 32) | (a_lo + a_off), bd, idesc, 1);
              if (wt1) tc::umma_bf16_ss(col + wt_cols, ((uint64_t)a_hi << 32) | (a_lo + a_off + 16), bd, idesc, 1);
            }
          }
          tc::umma_commit(a_empty(st));
          if (s == nslabs - 1 && q >= 1) tc::umma_commit(acc_full(q - 1));
          if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
        tc::umma_commit(w_empty(ws));
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: warps 4-7 drain w-tile 0, warps 8-11 w-tile 1 =====================
    const int wt = (warp - 4) >> 2;
    const int ew = warp & 3;          // TMEM lane quadrant this warp may touch
    const int m = ew * 32 + lane;
    const int h = h0 + (m >> 3);
    const int w = w0 + wt * 8 + (m & 7);
    if (w0 + wt * 8 < p.W) {
      const bool valid = h < p.H && w < p.W;
      for (int pl = 0; pl < planes; ++pl) {
        tc::mbar_wait(acc_full(pl), 0);
        tc::tc_fence_after();
        const int64_t row = (((int64_t)n * p.D + d0 + pl) * p.H + h) * p.W + w;
        const uint32_t col0 = (uint32_t)((wt * p.dseg + (p.dseg - 1 - pl)) * p.n_tile);
        for (int cc = 0; cc < p.n_tile / 16; ++cc) {
          uint32_t r[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + col0 + cc * 16, r);
          tc::tmem_ld_wait();
          const int ch = nchunk * p.n_tile + cc * 16;
          if (p.ksplit > 1) {
            if (valid) {
              float4* dst = reinterpret_cast<float4*>(p.partial + ((int64_t)zs * p.N * p.D * p.H * p.W + row) * (p.co0 + p.co1) + ch);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
            }
            continue;
          }
          uint32_t packed[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 bb = *reinterpret_cast<const float2*>(bias_s + cc * 16 + 2 * i);
            __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[2 * i]) + bb.x, __uint_as_float(r[2 * i + 1]) + bb.y);
            packed[i] = *reinterpret_cast<uint32_t*>(&hb);
          }
          if (valid) {
            bf16* dst = ch < p.co0 ? p.y0 + row * p.co0 + ch : p.y1 + row * p.co1 + (ch - p.co0);
            reinterpret_cast<uint4*>(dst)[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            reinterpret_cast<uint4*>(dst)[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

// out[nchunk][slab][kh][kw][kc(2)][kd(3)][n(n_tile)][8]: one thread builds one 16-byte group (8 consecutive k).  Threads are
// numbered with the filter tap fastest, so that neighbouring threads read neighbouring floats of w (the 27 taps of one
// (o, k) pair are contiguous); the index is decomposed once per group in 32-bit arithmetic (the first version did six
// 64-bit div/mod per ELEMENT and cost 0.14 ms per step).
__global__ void __launch_bounds__(256) pack_k3_tc2_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout_f, int Cin_f, int dgrad,
                                                          int n_tile_i, int slabs_i, int64_t total) {
  const uint32_t groups = (uint32_t)(total / 8), n_tile = (uint32_t)n_tile_i, slabs = (uint32_t)slabs_i;
  for (uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x; gid < groups; gid += gridDim.x * blockDim.x) {
    uint32_t t = gid;
    const uint32_t tap = t % 27; t /= 27;
    const uint32_t kc = t % 2; t /= 2;
    const uint32_t nn = t % n_tile; t /= n_tile;
    const uint32_t slab = t % slabs;
    const uint32_t nchunk = t / slabs;
    const uint32_t kd = tap / 9, khw = tap % 9, k0 = slab * 16 + kc * 8, o = nchunk * n_tile + nn;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      f[j] = !dgrad ? w[((int64_t)o * Cin_f + (k0 + j)) * 27 + tap] : w[((int64_t)(k0 + j) * Cin_f + o) * 27 + (26 - tap)];
    Vec8<bf16> v;
    v.set(f);
    const uint32_t g = ((((nchunk * slabs + slab) * 9 + khw) * 2 + kc) * 3 + kd) * n_tile + nn;
    v.store(out + (int64_t)g * 8);
  }
}

__host__ __device__ inline int n_tile_for(int cout) { return cout <= 128 ? cout : 128; }

// Every packed layout of every layer in ONE launch (the trainer calls this once per optimiser step instead of one pack kernel per
// layer and direction: 34 dependent ~3 us kernels on the step's main chain).  Job table in device memory, built once by the host.
struct PackJob {
  const float* w;          // torch weight [Cout][Cin][3][3][3] fp32
  bf16* out;               // packed destination
  int Cout, Cin, dgrad, pad;
  long long group_begin;   // first 16-byte group of this job in the global numbering; job njobs holds the total
};
static_assert(sizeof(PackJob) == 40, "PackJob layout is part of the C ABI (b200_pack_conv3_batched)");
constexpr int kMaxPackJobs = 128;

__global__ void __launch_bounds__(256) pack_k3_tc2_batched_kernel(const PackJob* __restrict__ jobs, int njobs) {
  __shared__ PackJob sj[kMaxPackJobs + 1];
  for (int i = threadIdx.x; i <= njobs; i += blockDim.x) sj[i] = jobs[i];
  __syncthreads();
  const long long groups = sj[njobs].group_begin;
  int j = 0;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < groups; gid += (long long)gridDim.x * blockDim.x) {
    while (gid >= sj[j + 1].group_begin) ++j;              // gid only grows: the scan resumes where it stopped
    const PackJob& q = sj[j];
    const int conv_in = q.dgrad ? q.Cout : q.Cin, conv_out = q.dgrad ? q.Cin : q.Cout;
    const uint32_t n_tile = (uint32_t)n_tile_for(conv_out), slabs = (uint32_t)(conv_in / 16);
    uint32_t t = (uint32_t)(gid - q.group_begin);
    const uint32_t tap = t % 27; t /= 27;
    const uint32_t kc = t % 2; t /= 2;
    const uint32_t nn = t % n_tile; t /= n_tile;
    const uint32_t slab = t % slabs;
    const uint32_t nchunk = t / slabs;
    const uint32_t kd = tap / 9, khw = tap % 9, k0 = slab * 16 + kc * 8, o = nchunk * n_tile + nn;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e)
      f[e] = !q.dgrad ? q.w[((int64_t)o * q.Cin + (k0 + e)) * 27 + tap] : q.w[((int64_t)(k0 + e) * q.Cin + o) * 27 + (26 - tap)];
    Vec8<bf16> v;
    v.set(f);
    const uint32_t g = ((((nchunk * slabs + slab) * 9 + khw) * 2 + kc) * 3 + kd) * n_tile + nn;
    v.store(q.out + (int64_t)g * 8);
  }
}

// split-K epilogue: y = bf16(sum_z partial[z] + bias), channels [0,co0) -> y0, [co0, co0+co1) -> y1
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int ksplit, int64_t rows, int cout, const float* __restrict__ bias,
                                     bf16* __restrict__ y0, bf16* __restrict__ y1, int co0, int co1) {
  const int CV = cout / 8;
  const int64_t total = rows * CV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const int64_t row = i / CV;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = bias ? bias[cv * 8 + k] : 0.f;
    for (int z = 0; z < ksplit; ++z) {
      const float4* src = reinterpret_cast<const float4*>(partial + ((int64_t)z * rows + row) * cout + cv * 8);
      const float4 a = src[0], b = src[1];
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    Vec8<bf16> v;
    v.set(acc);
    const int ch = cv * 8;
    if (ch < co0) v.store(y0 + row * co0 + ch); else v.store(y1 + row * co1 + (ch - co0));
  }
}

float* splitk_scratch(size_t bytes) {
  static float* buf = nullptr;
  static size_t cap = 0;
  if (cap < bytes) {
    if (buf) cudaFree(buf);
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { buf = nullptr; cap = 0; return nullptr; }
    cap = bytes;
  }
  return buf;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// tensor map over an NDHWC bf16 tensor, box = 16 channels x 18 (w) x 18 (h) x 1 x 1, 32-byte swizzle
int make_act_map(CUtensorMap* tm, const void* base, int C, int N, int D, int H, int W) {
  EncodeTiledFn enc = get_encode();
  B200_REQUIRE(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {16, kHalo, kHalo, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for C=%d N=%d D=%d H=%d W=%d", (int)r, C, N, D, H, W);
  return B200_OK;
}

}  // namespace

int64_t b200_pack_conv3_bytes_tc2(int Cout, int Cin) { return (int64_t)27 * Cin * Cout * 2; }

int b200_pack_conv3_weights_tc2(int mode, const float* w, void* out, int Cout, int Cin, cudaStream_t stream) {
  const int dgrad = mode == B200_PACK_DGRAD_TC;
  const int conv_in = dgrad ? Cout : Cin, conv_out = dgrad ? Cin : Cout;
  B200_REQUIRE(conv_in % 16 == 0 && conv_out % 16 == 0 && (conv_out <= 128 || conv_out % 128 == 0), B200_ERR_UNSUPPORTED,
               "pack_conv3_weights(tc2): channel counts %d -> %d not supported by the tcgen05 path", conv_in, conv_out);
  const int n_tile = n_tile_for(conv_out), slabs = conv_in / 16;
  const int64_t total = (int64_t)27 * Cin * Cout;
  pack_k3_tc2_kernel<<<b200_grid_for(total / 8, 256, B200_NUM_SMS * 4), 256, 0, stream>>>(w, (bf16*)out, Cout, Cin, dgrad, n_tile, slabs, total);
  B200_CHECK_LAUNCH("pack_conv3_weights_tc2");
  return B200_OK;
}

int b200_pack_conv3_batched_tc2(const void* jobs, int njobs, long long total_groups, cudaStream_t stream) {
  B200_REQUIRE(jobs && njobs > 0 && njobs <= kMaxPackJobs, B200_ERR_SHAPE, "pack_conv3_batched: 1..%d jobs", kMaxPackJobs);
  pack_k3_tc2_batched_kernel<<<b200_grid_for(total_groups, 256, B200_NUM_SMS * 4), 256, 0, stream>>>((const PackJob*)jobs, njobs);
  B200_CHECK_LAUNCH("pack_conv3_batched");
  return B200_OK;
}

int b200_conv3d_k3_tc2(const void* x0, int c0, const void* x1, int c1, const void* wpack, const float* bias, void* y0, int co0, void* y1,
                       int co1, int N, int D, int H, int W, cudaStream_t stream) {
  B200_REQUIRE(b200_aligned(x0, 16) && b200_aligned(x1, 16) && b200_aligned(y0, 16) && b200_aligned(y1, 16) && b200_aligned(wpack, 16),
               B200_ERR_ALIGN, "conv3d_k3(tcgen05): pointers must be 16-byte aligned");
  Tc2Params p;
  p.wpack = (const uint8_t*)wpack; p.bias = bias;
  p.y0 = (bf16*)y0; p.y1 = (bf16*)y1; p.co0 = co0; p.co1 = co1;
  p.c0 = c0; p.c1 = c1;
  p.N = N; p.D = D; p.H = H; p.W = W;
  const int cout = co0 + co1;
  p.n_tile = n_tile_for(cout);
  p.slabs = (c0 + c1) / 16;
  p.kd_per_mma = 3 * p.n_tile <= 256 ? 3 : 2;
  { const char* e = getenv("B200_TC_REPEAT"); p.mma_repeat = e ? atoi(e) : 1; if (p.mma_repeat < 0) p.mma_repeat = 0; }
  int dseg = 256 / (2 * p.n_tile);
  if (dseg > kMaxDseg) dseg = kMaxDseg;
  if (dseg < 1) dseg = 1;
  if (dseg > D) dseg = D;
  p.dseg = dseg;
  p.dblocks = (D + dseg - 1) / dseg;
  int cols = dseg * 2 * p.n_tile, pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  p.tmem_cols = pow2;
  p.tiles_w = (W + 15) / 16;
  const int tiles_h = (H + 15) / 16;
  const size_t wbytes = (size_t)864 * p.n_tile;
  // shared-memory plan: as many activation stages as fit next to the weight slab(s); keep two CTAs per SM when possible
  const size_t budget2 = 110 * 1024;  // per CTA for 2 CTAs / SM
  p.wstages = (p.slabs > 1 && 2 * wbytes + 4 * kStageBytes + kSmemHeader + 1024 <= budget2) ? 2 : 1;
  size_t avail = (p.wstages * wbytes + 3 * kStageBytes + kSmemHeader + 1024 <= budget2 ? budget2 : (size_t)200 * 1024) - p.wstages * wbytes - kSmemHeader - 1024;
  int stages = (int)(avail / kStageBytes);
  if (stages > kMaxStages) stages = kMaxStages;
  B200_REQUIRE(stages >= 2, B200_ERR_UNSUPPORTED, "conv3d_k3(tcgen05): not enough shared memory for n_tile=%d", p.n_tile);
  p.stages = stages;
  const size_t smem = kSmemHeader + (size_t)stages * kStageBytes + p.wstages * wbytes + 1024;
  B200_REQUIRE((int64_t)N * p.dblocks <= 65535 && cout / p.n_tile <= 65535, B200_ERR_UNSUPPORTED, "conv3d_k3(tcgen05): grid too large");
  CUtensorMap tm0, tm1;
  int rc = make_act_map(&tm0, x0, c0, N, D, H, W);
  if (rc) return rc;
  if (c1) { rc = make_act_map(&tm1, x1, c1, N, D, H, W); if (rc) return rc; } else tm1 = tm0;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(conv3d_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  // deep layers: few output tiles but many channel slabs -> split the reduction over CTAs (fp32 partials + one reduce pass)
  p.ksplit = 1; p.slabs_per = p.slabs; p.partial = nullptr;
  const int64_t ctas = (int64_t)p.tiles_w * tiles_h * N * p.dblocks * (cout / p.n_tile);
  static int splitk_enabled = -1;
  if (splitk_enabled < 0) { const char* e = getenv("B200_CONV_SPLITK"); splitk_enabled = e ? atoi(e) : 1; }
  if (splitk_enabled && ctas * 2 <= B200_NUM_SMS && p.slabs >= 4) {
    // aim for ONE wave of resident CTAs: a second wave repeats the fixed per-CTA cost (TMEM allocation, first weight slab,
    // fp32 partial drain), which dominates these small problems
    const int per_sm = 2 * smem <= (size_t)220 * 1024 ? 2 : 1;
    int want = (int)((per_sm * B200_NUM_SMS) / ctas);
    if (want < 1) want = 1;
    if (want > p.slabs) want = p.slabs;
    p.slabs_per = (p.slabs + want - 1) / want;
    p.ksplit = (p.slabs + p.slabs_per - 1) / p.slabs_per;
    if (p.ksplit > 1) {
      const size_t bytes = (size_t)p.ksplit * N * D * H * W * cout * sizeof(float);
      p.partial = splitk_scratch(bytes);
      B200_REQUIRE(p.partial != nullptr, B200_ERR_CUDA, "conv3d_k3(tcgen05): could not allocate the split-K scratch buffer (%zu bytes)", bytes);
    } else { p.slabs_per = p.slabs; }
  }
  dim3 grid((unsigned)(p.tiles_w * tiles_h), (unsigned)(N * p.dblocks), (unsigned)(cout / p.n_tile * p.ksplit));
  conv3d_tc2_kernel<<<grid, kThreads, smem, stream>>>(p, tm0, tm1);
  B200_CHECK_LAUNCH("conv3d_k3_tc2");
  if (p.ksplit > 1) {
    const int64_t rows = (int64_t)N * D * H * W;
    splitk_reduce_kernel<<<b200_grid_for(rows * (cout / 8), 256, B200_NUM_SMS * 8), 256, 0, stream>>>(p.partial, p.ksplit, rows, cout, bias, p.y0, p.y1, co0, co1);
    B200_CHECK_LAUNCH("conv3d_k3_tc2_splitk_reduce");
  }
  return B200_OK;
}
