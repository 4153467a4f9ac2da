// convt_tc.cu — ConvTranspose3d(k=2, s=2) forward and data gradient on the tensor cores
// (models/unet.py:56-58, 79).  With kernel = stride = 2 the up-convolution is a plain GEMM per coarse voxel:
//   forward : y[child(v)][co] = b[co] + sum_ci x[v][ci] W[ci][co][child]      M = voxels, K = Cin,     N = 8*Cout
//   backward: gx[v][ci]       = sum_{child,co} gy[child(v)][co] W[ci][co][child]  M = voxels, K = 8*Cout, N = Cin
// Both are HBM-bound (the fine tensor is 8x the coarse one); the tensor core replaces the CUDA-core
// implicit GEMM that ran at ~10 TFLOP/s.  A = 16x16 coarse voxels x 16 channels per stage, fetched by TMA
// (SWIZZLE_32B rows); for the backward the eight child planes of the fine gradient are fetched with
// element-stride-2 tensor maps, so the "gather" costs no instructions.  B = bf16 weights re-packed per call
// into the K-major UMMA layout.  The forward epilogue scatters each accumulator row to its eight children.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tma_maps.cuh"

namespace {

using bf16 = __nv_bfloat16;

constexpr int kABytes = 256 * 32;          // 16 x 16 voxels x 16 ch
constexpr int kStages = 4;
constexpr int kThreads = 384;              // w0: TMA, w1: TMEM alloc, w2: MMA, w4-11: epilogue
constexpr int kHeader = 1024;

struct CtcParams {
  const uint8_t* wpack;        // [nchunk][kslab][kc 2][n_tile][8] bf16
  const float* bias;           // forward only
  bf16* out;                   // forward: fine tensor [N,2D,2H,2W,Cout]; backward: coarse tensor [N,D,H,W,Cin]
  int N, D, H, W;              // coarse geometry
  int cin, cout;
  int n_tile, nchunks, kslabs, tiles_w;
  int cpc;                     // forward: children per N chunk
  int bwd;
};

__device__ __forceinline__ uint64_t desc_kmajor_sw32(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}

__global__ void __launch_bounds__(kThreads, 3)
convt_tc_kernel(const CtcParams p, const __grid_constant__ CUtensorMap tm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = tc::smem_u32(bars);
  auto full = [&](int i) { return bar0 + 8u * i; };
  auto empty = [&](int i) { return bar0 + 8u * (4 + i); };
  const uint32_t acc_done = bar0 + 8u * 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
  float* bias_s = reinterpret_cast<float*>(smem + 256);  // [<=128]
  const uint32_t wbytes = (uint32_t)p.n_tile * 32;
  const uint32_t stage_bytes = kABytes + ((wbytes + 255u) & ~255u);
  uint8_t* ring = smem + kHeader;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = blockIdx.x % p.tiles_w, th = blockIdx.x / p.tiles_w;
  const int n = blockIdx.y / p.D, d = blockIdx.y % p.D;
  const int nchunk = blockIdx.z;
  const int w0 = tw * 16, h0 = th * 16;
  const int tmem_cols = 2 * p.n_tile <= 32 ? 32 : (2 * p.n_tile <= 64 ? 64 : (2 * p.n_tile <= 128 ? 128 : (2 * p.n_tile <= 256 ? 256 : 512)));

  if (warp == 2 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { tc::mbar_init(full(i), 1); tc::mbar_init(empty(i), 1); }
    tc::mbar_init(acc_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tc::smem_u32(tmem_slot), tmem_cols);
    tc::tmem_relinquish();
  }
  if (!p.bwd && warp >= 4 && threadIdx.x - 128 < p.cout) bias_s[threadIdx.x - 128] = p.bias ? p.bias[threadIdx.x - 128] : 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    if (lane == 0) tma::prefetch(&tm);
    for (int s = 0; s < p.kslabs; ++s) {
      const int st = s % kStages;
      tc::mbar_wait(empty(st), ((s / kStages) & 1) ^ 1);
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(full(st), kABytes + wbytes);
        const uint32_t dst = tc::smem_u32(ring + st * stage_bytes);
        if (!p.bwd) {
          tma::load_5d(dst, &tm, s * 16, w0, h0, d, n, full(st));
        } else {
          const int child = s / (p.cout / 16), cs = s % (p.cout / 16);
          tma::load_5d(dst, &tm, cs * 16, 2 * w0 + (child & 1), 2 * h0 + ((child >> 1) & 1), 2 * d + (child >> 2), n, full(st));
        }
        tc::bulk_g2s(dst + kABytes, p.wpack + ((size_t)nchunk * p.kslabs + s) * wbytes, wbytes, full(st));
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    const uint64_t a_proto = desc_kmajor_sw32(0, 16 * 32);                         // 8-row groups = tile rows, 16 voxels apart
    const uint64_t b_proto = tc::smem_desc_kmajor_noswz(0, 16u * p.n_tile, 128);   // [kc][n][8]: K chunks n_tile*16 B apart
    const uint32_t a_hi = (uint32_t)(a_proto >> 32), a_lo0 = (uint32_t)a_proto;
    const uint32_t b_hi = (uint32_t)(b_proto >> 32), b_lo0 = (uint32_t)b_proto;
    const uint32_t idesc = tc::idesc_bf16_f32(128, p.n_tile);
    const bool wt1 = w0 + 8 < p.W;
    for (int s = 0; s < p.kslabs; ++s) {
      const int st = s % kStages;
      tc::mbar_wait(full(st), (s / kStages) & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t base = tc::smem_u32(ring + st * stage_bytes);
        const uint32_t a_lo = a_lo0 + (base >> 4), b_lo = b_lo0 + ((base + kABytes) >> 4);
        tc::umma_bf16_ss(tmem_base, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, s > 0);
        if (wt1) tc::umma_bf16_ss(tmem_base + p.n_tile, ((uint64_t)a_hi << 32) | (a_lo + 16), ((uint64_t)b_hi << 32) | b_lo, idesc, s > 0);
        tc::umma_commit(empty(st));
        if (s == p.kslabs - 1) tc::umma_commit(acc_done);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int wt = (warp - 4) >> 2, ew = warp & 3;
    const int m = ew * 32 + lane;
    const int h = h0 + (m >> 3), w = w0 + wt * 8 + (m & 7);
    tc::mbar_wait(acc_done, 0);
    tc::tc_fence_after();
    if (w0 + wt * 8 < p.W) {
      const bool valid = h < p.H && w < p.W;
      const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(wt * p.n_tile);
      if (!p.bwd) {
        // columns = (child_local, co): scatter each child to its fine voxel
        const int FD = 2 * p.D, FH = 2 * p.H, FW = 2 * p.W;
        for (int cl = 0; cl < p.cpc; ++cl) {
          const int child = nchunk * p.cpc + cl;
          const int64_t row = (((int64_t)n * FD + 2 * d + (child >> 2)) * FH + 2 * h + ((child >> 1) & 1)) * FW + 2 * w + (child & 1);
          for (int cc = 0; cc < p.cout / 16; ++cc) {
            uint32_t r[16];
            tc::tmem_ld16(tbase + (uint32_t)(cl * p.cout + cc * 16), r);
            tc::tmem_ld_wait();
            uint32_t packed[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 bb = *reinterpret_cast<const float2*>(bias_s + cc * 16 + 2 * i);
              __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[2 * i]) + bb.x, __uint_as_float(r[2 * i + 1]) + bb.y);
              packed[i] = *reinterpret_cast<uint32_t*>(&hb);
            }
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(p.out + row * p.cout + cc * 16);
              dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
              dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            }
          }
        }
      } else {
        const int64_t row = (((int64_t)n * p.D + d) * p.H + h) * p.W + w;
        for (int cc = 0; cc < p.n_tile / 16; ++cc) {
          uint32_t r[16];
          tc::tmem_ld16(tbase + (uint32_t)(cc * 16), r);
          tc::tmem_ld_wait();
          uint32_t packed[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            packed[i] = *reinterpret_cast<uint32_t*>(&hb);
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + row * p.cin + nchunk * p.n_tile + cc * 16);
            dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// torch ConvTranspose3d weight w[ci][co][child] fp32 -> [nchunk][kslab][kc][n_tile][8] bf16
__global__ void pack_convt_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cin, int Cout, int bwd, int n_tile, int kslabs, int cpc,
                                  int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int j = (int)(t % 8); t /= 8;
    const int nn = (int)(t % n_tile); t /= n_tile;
    const int kc = (int)(t % 2); t /= 2;
    const int ks = (int)(t % kslabs);
    const int nchunk = (int)(t / kslabs);
    int ci, co, child;
    if (!bwd) {  // K = ci, N = (child, co)
      ci = ks * 16 + kc * 8 + j;
      child = nchunk * cpc + nn / Cout;
      co = nn % Cout;
    } else {     // K = (child, co), N = ci
      child = ks / (Cout / 16);
      co = (ks % (Cout / 16)) * 16 + kc * 8 + j;
      ci = nchunk * n_tile + nn;
    }
    out[i] = __float2bfloat16_rn(w[((int64_t)ci * Cout + co) * 8 + child]);
  }
}

bf16* scratch_weights(size_t bytes) {
  // packed weights are re-derived every call (the fp32 master weights change every optimiser step); a process-wide
  // scratch buffer keeps the C ABI of b200_convt2_* free of a workspace argument.  One stream per process (torch's).
  static bf16* buf[2] = {nullptr, nullptr};
  static size_t cap[2] = {0, 0};
  static int flip = 0;
  flip ^= 1;  // alternate two buffers so that a forward and a backward call in flight do not share one
  if (cap[flip] < bytes) {
    if (buf[flip]) cudaFree(buf[flip]);
    if (cudaMalloc(&buf[flip], bytes) != cudaSuccess) return nullptr;
    cap[flip] = bytes;
  }
  return buf[flip];
}

int run(int bwd, const void* act, const float* w, const float* bias, void* out, int N, int D, int H, int W, int Cin, int Cout, cudaStream_t stream) {
  CtcParams p;
  p.bias = bias; p.out = (bf16*)out;
  p.N = N; p.D = D; p.H = H; p.W = W; p.cin = Cin; p.cout = Cout; p.bwd = bwd;
  if (!bwd) {
    p.n_tile = 8 * Cout <= 256 ? 8 * Cout : 256;
    p.nchunks = 8 * Cout / p.n_tile;
    p.kslabs = Cin / 16;
    p.cpc = p.n_tile / Cout;
  } else {
    p.n_tile = Cin <= 128 ? Cin : 128;
    p.nchunks = Cin / p.n_tile;
    p.kslabs = 8 * Cout / 16;
    p.cpc = 0;
  }
  p.tiles_w = (W + 15) / 16;
  const int tiles_h = (H + 15) / 16;
  const size_t wtotal = (size_t)8 * Cin * Cout * 2;
  bf16* wp = scratch_weights(wtotal < (1u << 20) ? (1u << 20) : wtotal);
  B200_REQUIRE(wp != nullptr, B200_ERR_CUDA, "convt2(tcgen05): could not allocate the packed-weight scratch buffer");
  const int64_t total = (int64_t)8 * Cin * Cout;
  pack_convt_kernel<<<b200_grid_for(total, 256, B200_NUM_SMS * 4), 256, 0, stream>>>(w, wp, Cin, Cout, bwd, p.n_tile, p.kslabs, p.cpc, total);
  B200_CHECK_LAUNCH("convt2_pack");
  p.wpack = (const uint8_t*)wp;
  CUtensorMap tm;
  int rc = bwd ? tma::make_ndhwc_map_stride2(&tm, act, Cout, N, 2 * D, 2 * H, 2 * W, 16, 16) : tma::make_ndhwc_map(&tm, act, Cin, N, D, H, W, 16, 16);
  if (rc) return rc;
  const uint32_t wbytes = (uint32_t)p.n_tile * 32;
  const size_t smem = kHeader + (size_t)kStages * (kABytes + ((wbytes + 255u) & ~255u)) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(convt_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set = true;
  }
  B200_REQUIRE((int64_t)N * D <= 65535, B200_ERR_UNSUPPORTED, "convt2(tcgen05): grid too large");
  dim3 grid((unsigned)(p.tiles_w * tiles_h), (unsigned)(N * D), (unsigned)p.nchunks);
  convt_tc_kernel<<<grid, kThreads, smem, stream>>>(p, tm);
  B200_CHECK_LAUNCH(bwd ? "convt2_bwd_data_tc" : "convt2_fwd_tc");
  return B200_OK;
}

}  // namespace

bool b200_convt2_tc_supported(int Cin, int Cout) {
  if (Cin % 16 || Cout % 16 || Cin < 16 || Cout < 16) return false;
  if (Cout > 128) return false;                       // bias slice in shared memory, n_tile bookkeeping
  if (8 * Cout > 256 && (8 * Cout) % 256) return false;
  if (Cin > 128 && Cin % 128) return false;
  return true;
}
int b200_convt2_fwd_tc(const void* x, const float* w, const float* bias, void* y, int N, int D, int H, int W, int Cin, int Cout, cudaStream_t st) {
  return run(0, x, w, bias, y, N, D, H, W, Cin, Cout, st);
}
int b200_convt2_bwd_data_tc(const void* gy, const float* w, void* gx, int N, int D, int H, int W, int Cin, int Cout, cudaStream_t st) {
  return run(1, gy, w, nullptr, gx, N, D, H, W, Cin, Cout, st);
}
