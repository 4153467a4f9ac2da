// conv_tc.cu — tcgen05 / TMEM implicit-GEMM 3x3x3 convolution, bf16 in, fp32 accumulate, bf16 out.
//
// Replaces cuDNN's fprop / bwd-data kernels behind nn.Conv3d(k=3,p=1) (models/unet.py:11,15) for
// channel counts that are multiples of 16.  "Plane-streaming" formulation:
//
//   * a CTA owns a 16(w) x 16(h) x DSEG(d) box of output voxels of one sample and all (<=128)
//     output channels of one N-chunk.  One output plane = two UMMA tiles of M = 128 rows
//     (8 consecutive w x 16 h), accumulators live in TMEM: (plane, w-tile) -> N fp32 columns.
//   * the input is streamed one (d-plane, 16-channel slab) at a time: producer warps gather an
//     18 x 18 halo plane (zero padded) with coalesced 32-byte reads and lay it out in shared memory
//     as [k-chunk(2)][h(18)][w(18)] x 16 B — the canonical K-major, no-swizzle UMMA layout in which
//     a filter tap (kh,kw) is just a byte offset of the descriptor start address (SBO = one halo
//     row, LBO = one chunk plane).  Each staged plane feeds the three output planes it touches
//     (kd = 0,1,2): 27 taps x 2 w-tiles = up to 54 tcgen05.mma per stage, so shared-memory halo
//     data is read by the tensor core only, never re-staged.
//   * the 27 x 16 x N weight slab arrives by cp.async.bulk (TMA bulk copy) in UMMA B layout,
//     pre-packed by b200_pack_conv3_weights(B200_PACK_*_TC).
//   * warp roles: 4 producer warps, 4 epilogue warps (TMEM -> registers -> +bias -> bf16 -> global),
//     1 MMA-issuing warp (one elected thread), 1 weight-TMA / TMEM-allocator warp; all hand-offs are
//     mbarriers (full/empty rings, tcgen05.commit arrivals).
//   * the virtual concat of two inputs (models/unet.py:84) selects the source tensor per channel
//     slab; the data-gradient un-concat selects the destination per 16-channel column chunk.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using bf16 = __nv_bfloat16;

constexpr int kHalo = 18;                         // 16 + 2
constexpr int kPlaneVox = kHalo * kHalo;          // 324 voxels per halo plane
constexpr int kChunkBytes = kPlaneVox * 16;       // one 8-channel chunk plane
constexpr int kStageBytes = 2 * kChunkBytes;      // one 16-channel slab plane = 10368 B
constexpr int kStages = 4;
constexpr int kProducerWarps = 4;
constexpr int kThreads = 320;                     // 4 producer + 4 epilogue + MMA + weights
constexpr int kMaxDseg = 8;
constexpr int kSmemHeader = 256;                  // barriers + tmem pointer

struct TcParams {
  const bf16* x0; const bf16* x1; int c0, c1;
  const uint8_t* wpack; const float* bias;
  bf16* y0; bf16* y1; int co0, co1;
  int N, D, H, W;
  int n_tile, dseg, dblocks, slabs, wstages, tmem_cols, tiles_w;
};

__global__ void __launch_bounds__(kThreads, 1)
conv3d_tc_kernel(const TcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // barrier slots: [0,4) a_full, [4,8) a_empty, [8,10) w_full, [10,12) w_empty, [12,20) acc_full
  const uint32_t bar0 = tc::smem_u32(bars);
  auto a_full = [&](int i) { return bar0 + 8u * i; };
  auto a_empty = [&](int i) { return bar0 + 8u * (4 + i); };
  auto w_full = [&](int i) { return bar0 + 8u * (8 + i); };
  auto w_empty = [&](int i) { return bar0 + 8u * (10 + i); };
  auto acc_full = [&](int i) { return bar0 + 8u * (12 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 192);
  uint8_t* act = smem + kSmemHeader;
  uint8_t* wts = act + kStages * kStageBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = blockIdx.x % p.tiles_w, th = blockIdx.x / p.tiles_w;
  const int n = blockIdx.y / p.dblocks, db = blockIdx.y % p.dblocks;
  const int nchunk = blockIdx.z;
  const int w0 = tw * 16, h0 = th * 16, d0 = db * p.dseg;
  const int planes = min(p.dseg, p.D - d0);
  const int nq = planes + 2;  // input planes q = -1 .. planes
  const int total_stages = p.slabs * nq;
  const uint32_t wbytes = 864u * p.n_tile;  // 27 taps x 16 ci x n_tile x 2 B

  if (warp == 8 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { tc::mbar_init(a_full(i), kProducerWarps); tc::mbar_init(a_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(w_full(i), 1); tc::mbar_init(w_empty(i), 1); }
    for (int i = 0; i < kMaxDseg; ++i) tc::mbar_init(acc_full(i), 1);
    tc::fence_barrier_init();
  }
  if (warp == 9) {
    tc::tmem_alloc(tc::smem_u32(tmem_slot), p.tmem_cols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kProducerWarps) {
    // ===================== producers: global halo plane -> shared memory (UMMA A layout) =====================
    const int tid = threadIdx.x;  // 0..127
    for (int it = 0; it < total_stages; ++it) {
      const int s = it / nq, q = it % nq - 1;
      const int st = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      tc::mbar_wait(a_empty(st), ph ^ 1);
      const int d = d0 + q;
      const bool dvalid = (unsigned)d < (unsigned)p.D;
      const int c = s * 16;
      const bf16* base; int cs, coff;
      if (c < p.c0) { base = p.x0; cs = p.c0; coff = c; } else { base = p.x1; cs = p.c1; coff = c - p.c0; }
      uint4 v[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int id = tid + 128 * j;
        v[j] = make_uint4(0, 0, 0, 0);
        if (id < 2 * kPlaneVox) {
          const int vox = id >> 1, ch = id & 1;
          const int hh = vox / kHalo, ww = vox - hh * kHalo;
          const int h = h0 + hh - 1, w = w0 + ww - 1;
          if (dvalid && (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W) {
            const int64_t row = (((int64_t)n * p.D + d) * p.H + h) * p.W + w;
            v[j] = __ldg(reinterpret_cast<const uint4*>(base + row * cs + coff + ch * 8));
          }
        }
      }
      uint8_t* stage = act + st * kStageBytes;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int id = tid + 128 * j;
        if (id < 2 * kPlaneVox) {
          const int vox = id >> 1, ch = id & 1;
          *reinterpret_cast<uint4*>(stage + ch * kChunkBytes + vox * 16) = v[j];
        }
      }
      tc::fence_proxy_async_smem();  // make generic-proxy writes visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(a_full(st));
    }
  } else if (warp == 9) {
    // ===================== weight slabs: one TMA bulk copy each =====================
    if (lane == 0) {
      for (int s = 0; s < p.slabs; ++s) {
        const int ws = s % p.wstages;
        const uint32_t ph = (s / p.wstages) & 1;
        tc::mbar_wait(w_empty(ws), ph ^ 1);
        tc::mbar_arrive_expect_tx(w_full(ws), wbytes);
        tc::bulk_g2s(tc::smem_u32(wts + (size_t)ws * wbytes), p.wpack + ((size_t)nchunk * p.slabs + s) * wbytes, wbytes, w_full(ws));
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issue: one thread drives the tensor core =====================
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_bf16_f32(128, p.n_tile);
      const uint32_t b_lbo = 16u * p.n_tile, b_tap = 32u * p.n_tile;
      for (int s = 0; s < p.slabs; ++s) {
        const int ws = s % p.wstages;
        tc::mbar_wait(w_full(ws), (s / p.wstages) & 1);
        tc::tc_fence_after();
        const uint32_t w_base = tc::smem_u32(wts + (size_t)ws * wbytes);
        for (int qi = 0; qi < nq; ++qi) {
          const int it = s * nq + qi;
          const int st = it % kStages;
          tc::mbar_wait(a_full(st), (it / kStages) & 1);
          tc::tc_fence_after();
          const int q = qi - 1;
          const uint32_t a_base = tc::smem_u32(act + st * kStageBytes);
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const int pl = q - kd + 1;  // output plane fed by this input plane through tap row kd
            if (pl < 0 || pl >= planes) continue;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const int tap = kd * 9 + kh * 3 + kw;
                const uint64_t bdesc = tc::smem_desc_kmajor_noswz(w_base + tap * b_tap, b_lbo, 128);
                const uint32_t accumulate = (s | kd | kh | kw) != 0;
#pragma unroll
                for (int wt = 0; wt < 2; ++wt) {
                  if (w0 + wt * 8 >= p.W) continue;
                  const uint64_t adesc = tc::smem_desc_kmajor_noswz(a_base + (kh * kHalo + kw + wt * 8) * 16, kChunkBytes, kHalo * 16);
                  tc::umma_bf16_ss(tmem_base + (uint32_t)((pl * 2 + wt) * p.n_tile), adesc, bdesc, idesc, accumulate);
                }
              }
            }
          }
          tc::umma_commit(a_empty(st));                                   // stage consumed once these MMAs retire
          if (s == p.slabs - 1 && q >= 1) tc::umma_commit(acc_full(q - 1));  // plane q-1 has all 27*slabs contributions
        }
        tc::umma_commit(w_empty(ws));
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> (+bias, bf16) -> global =====================
    const int ew = warp - 4;          // == warp % 4: this warp may touch TMEM lanes [32*ew, 32*ew+32)
    const int m = ew * 32 + lane;     // accumulator row = (h offset, w offset within the 8-wide tile)
    const int h = h0 + (m >> 3);
    for (int pl = 0; pl < planes; ++pl) {
      tc::mbar_wait(acc_full(pl), 0);
      tc::tc_fence_after();
#pragma unroll
      for (int wt = 0; wt < 2; ++wt) {
        if (w0 + wt * 8 >= p.W) continue;
        const int w = w0 + wt * 8 + (m & 7);
        const bool valid = h < p.H && w < p.W;
        const int64_t row = (((int64_t)n * p.D + d0 + pl) * p.H + h) * p.W + w;
        for (int cc = 0; cc < p.n_tile / 16; ++cc) {
          uint32_t r[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((pl * 2 + wt) * p.n_tile + cc * 16), r);
          tc::tmem_ld_wait();
          const int ch = nchunk * p.n_tile + cc * 16;
          uint32_t packed[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float a = __uint_as_float(r[2 * i]), b = __uint_as_float(r[2 * i + 1]);
            if (p.bias) { a += __ldg(p.bias + ch + 2 * i); b += __ldg(p.bias + ch + 2 * i + 1); }
            __nv_bfloat162 hb = __floats2bfloat162_rn(a, b);
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
  if (warp == 9) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

// out[nchunk][slab][tap][kc(2)][n(n_tile)][8]: the UMMA B operand (N x 16, K-major, no swizzle) of each tap
__global__ void pack_k3_tc_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout_f, int Cin_f, int dgrad,
                                  int n_tile, int slabs, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int j = (int)(t % 8); t /= 8;
    const int nn = (int)(t % n_tile); t /= n_tile;
    const int kc = (int)(t % 2); t /= 2;
    const int tap = (int)(t % 27); t /= 27;
    const int slab = (int)(t % slabs);
    const int nchunk = (int)(t / slabs);
    const int k = slab * 16 + kc * 8 + j;       // input channel of this convolution
    const int o = nchunk * n_tile + nn;         // output channel of this convolution
    float v;
    if (!dgrad) v = w[((int64_t)o * Cin_f + k) * 27 + tap];            // fprop: w[co=o][ci=k][tap]
    else v = w[((int64_t)k * Cin_f + o) * 27 + (26 - tap)];            // dgrad: w[co=k][ci=o][flipped tap]
    out[i] = __float2bfloat16_rn(v);
  }
}

inline int n_tile_for(int cout) { return cout <= 128 ? cout : 128; }

}  // namespace

bool b200_conv3d_k3_tc_supported(int c0, int c1, int co0, int co1, int N, int D, int H, int W) {
  const int cout = co0 + co1;
  if (c0 <= 0 || c0 % 16 || c1 % 16 || co0 % 16 || co1 % 16) return false;
  if (cout % 16 || (cout > 128 && cout % 128)) return false;
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) return false;
  return true;
}

int64_t b200_pack_conv3_bytes_tc(int Cout, int Cin) { return (int64_t)27 * Cin * Cout * 2; }

int b200_pack_conv3_weights_tc(int mode, const float* w, void* out, int Cout, int Cin, cudaStream_t stream) {
  // Cout/Cin are the dimensions of the torch weight [Cout, Cin, 3,3,3]
  const int dgrad = mode == B200_PACK_DGRAD_TC;
  const int conv_in = dgrad ? Cout : Cin, conv_out = dgrad ? Cin : Cout;
  B200_REQUIRE(conv_in % 16 == 0 && conv_out % 16 == 0 && (conv_out <= 128 || conv_out % 128 == 0), B200_ERR_UNSUPPORTED,
               "pack_conv3_weights(tc): channel counts %d -> %d not supported by the tcgen05 path", conv_in, conv_out);
  const int n_tile = n_tile_for(conv_out), slabs = conv_in / 16;
  const int64_t total = (int64_t)27 * Cin * Cout;
  pack_k3_tc_kernel<<<b200_grid_for(total, 256, B200_NUM_SMS * 8), 256, 0, stream>>>(w, (bf16*)out, Cout, Cin, dgrad, n_tile, slabs, total);
  B200_CHECK_LAUNCH("pack_conv3_weights_tc");
  return B200_OK;
}

int b200_conv3d_k3_tc(const void* x0, int c0, const void* x1, int c1, const void* wpack, const float* bias, void* y0, int co0,
                      void* y1, int co1, int N, int D, int H, int W, cudaStream_t stream) {
  B200_REQUIRE(b200_conv3d_k3_tc_supported(c0, c1, co0, co1, N, D, H, W), B200_ERR_UNSUPPORTED,
               "conv3d_k3(tcgen05): channels (%d+%d)->(%d+%d) need multiples of 16 (Cout <= 128 or a multiple of 128)", c0, c1, co0, co1);
  B200_REQUIRE(b200_aligned(x0, 16) && b200_aligned(x1, 16) && b200_aligned(y0, 16) && b200_aligned(y1, 16) && b200_aligned(wpack, 16),
               B200_ERR_ALIGN, "conv3d_k3(tcgen05): pointers must be 16-byte aligned");
  TcParams p;
  p.x0 = (const bf16*)x0; p.x1 = (const bf16*)x1; p.c0 = c0; p.c1 = c1;
  p.wpack = (const uint8_t*)wpack; p.bias = bias;
  p.y0 = (bf16*)y0; p.y1 = (bf16*)y1; p.co0 = co0; p.co1 = co1;
  p.N = N; p.D = D; p.H = H; p.W = W;
  const int cout = co0 + co1;
  p.n_tile = n_tile_for(cout);
  p.slabs = (c0 + c1) / 16;
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
  p.wstages = (p.slabs > 1 && 2 * wbytes + kStages * kStageBytes + kSmemHeader <= 100 * 1024) ? 2 : 1;
  const size_t smem = kSmemHeader + (size_t)kStages * kStageBytes + p.wstages * wbytes;
  B200_REQUIRE((int64_t)N * p.dblocks <= 65535 && cout / p.n_tile <= 65535, B200_ERR_UNSUPPORTED, "conv3d_k3(tcgen05): grid too large");
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)(p.tiles_w * tiles_h), (unsigned)(N * p.dblocks), (unsigned)(cout / p.n_tile));
  conv3d_tc_kernel<<<grid, kThreads, smem, stream>>>(p);
  B200_CHECK_LAUNCH("conv3d_k3_tc");
  return B200_OK;
}
