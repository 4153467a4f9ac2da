// head_kernels.cu — the small heads of the network:
//   * final 1x1x1 convolution (models/unet.py:62,87): NDHWC activations -> NCDHW fp32 logits
//     (HBM-bound: ~3 FLOP/B), forward and backward;
//   * DomainDiscriminator linear layers + 2-class CE (train_dann.py:34-49, 254-258).
#include "common.cuh"

namespace {

// voxel -> sample index: 32-bit division whenever the operands allow it (a 64-bit division per voxel is ~40 instructions)
__device__ __forceinline__ int64_t sample_index(int64_t v, int64_t S) {
  return ((v | S) >> 32) == 0 ? (int64_t)((uint32_t)v / (uint32_t)S) : v / S;
}


constexpr int kThreads = 256;
constexpr int kMaxCo = 16;
constexpr int kTileV = 128;  // voxels per tile in the backward kernel
constexpr int kMaxPartialBlocks = 592;

// ---------------------------------------------------------------- 1x1x1 forward
template <typename T>
__global__ void __launch_bounds__(kThreads)
conv1x1_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ y, int64_t N, int64_t S, int Cin, int Cout, int round_bf16) {
  extern __shared__ float ws[];  // [Cout][Cin] + [Cout]
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) ws[Cout * Cin + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int CV = Cin / 8;
  const int64_t total = N * S;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    float acc[kMaxCo];
#pragma unroll
    for (int co = 0; co < kMaxCo; ++co) acc[co] = (co < Cout) ? ws[Cout * Cin + co] : 0.f;
    for (int cv = 0; cv < CV; ++cv) {
      Vec8<T> xv;
      xv.load(x + v * Cin + cv * 8);
      float f[8];
      xv.get(f);
#pragma unroll
      for (int co = 0; co < kMaxCo; ++co)
        if (co < Cout) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[co] = fmaf(f[k], ws[co * Cin + cv * 8 + k], acc[co]);
        }
    }
    const int64_t n = v / S, s = v - n * S;
#pragma unroll
    for (int co = 0; co < kMaxCo; ++co)
      if (co < Cout) {
        float o = acc[co];
        if (round_bf16) o = __bfloat162float(__float2bfloat16_rn(o));
        __stcs(y + (n * Cout + co) * S + s, o);
      }
  }
}

// Fast path for the network's actual head (Cin = 16, Cout <= 4): weights and bias live in registers, two voxels per thread
// and iteration (their four 16-byte loads are issued before the math), streaming stores of the NCDHW fp32 logits.
// HBM-bound: 2*Cin B read + 4*Cout B written per voxel (201 MB at 2x128^3).
template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(kThreads)
conv1x1_fwd_small_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                         int64_t N, int64_t S, int round_bf16) {
  float wr[COUT][CIN], br[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    br[co] = bias ? bias[co] : 0.f;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) wr[co][ci] = w[co * CIN + ci];
  }
  const int64_t total = N * S, stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 2;
  for (int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < total; v0 += U * stride) {
    Vec8<T> xv[U][CIN / 8];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (v0 + u * stride < total) {
#pragma unroll
        for (int c8 = 0; c8 < CIN / 8; ++c8) xv[u][c8].load(x + (v0 + u * stride) * CIN + c8 * 8);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v >= total) continue;
      float acc[COUT];
#pragma unroll
      for (int co = 0; co < COUT; ++co) acc[co] = br[co];
#pragma unroll
      for (int c8 = 0; c8 < CIN / 8; ++c8) {
        float f[8];
        xv[u][c8].get(f);
#pragma unroll
        for (int co = 0; co < COUT; ++co)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[co] = fmaf(f[k], wr[co][c8 * 8 + k], acc[co]);
      }
      const int64_t n = sample_index(v, S), sp = v - n * S;
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        float o = acc[co];
        if (round_bf16) o = __bfloat162float(__float2bfloat16_rn(o));
        __stcs(y + (n * COUT + co) * S + sp, o);
      }
    }
  }
}

// ---------------------------------------------------------------- 1x1x1 backward
// One pass: gx = gy^T W (rows), and per-block partial dW/db.  P = Cout*(Cin+1) accumulators
// (the extra column is db).
template <typename T>
__global__ void __launch_bounds__(kThreads)
conv1x1_bwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ gy,
                   T* __restrict__ gx, float* __restrict__ partials, int64_t N, int64_t S, int Cin, int Cout) {
  extern __shared__ float sm[];
  const int Cin1 = Cin + 1;
  float* ws = sm;                        // [Cout][Cin]
  float* xs = ws + Cout * Cin;           // [kTileV][Cin1]  (last column = 1)
  float* gs = xs + kTileV * Cin1;        // [Cout][kTileV]
  float* red = gs + Cout * kTileV;       // [groups][P]
  const int P = Cout * Cin1;
  const int groups = kThreads / P > 0 ? kThreads / P : 1;
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) ws[i] = w[i];
  const int64_t total = N * S;
  const int64_t ntiles = (total + kTileV - 1) / kTileV;
  // accumulators: thread t < groups*P handles pair p = t % P on voxel slice group = t / P;
  // if P > kThreads each thread loops over pairs (acc kept in shared memory instead)
  const bool reg_mode = P <= kThreads;
  float acc = 0.f;
  if (!reg_mode)
    for (int i = threadIdx.x; i < P; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t v0 = tile * kTileV;
    for (int idx = threadIdx.x; idx < kTileV * Cin; idx += blockDim.x) {
      const int vv = idx / Cin, ci = idx % Cin;
      const int64_t v = v0 + vv;
      xs[vv * Cin1 + ci] = v < total ? to_f32<T>(x[v * Cin + ci]) : 0.f;
    }
    for (int vv = threadIdx.x; vv < kTileV; vv += blockDim.x) xs[vv * Cin1 + Cin] = (v0 + vv < total) ? 1.f : 0.f;
    for (int idx = threadIdx.x; idx < Cout * kTileV; idx += blockDim.x) {
      const int co = idx / kTileV, vv = idx % kTileV;
      const int64_t v = v0 + vv;
      float g = 0.f;
      if (v < total) {
        const int64_t n = v / S, s = v - n * S;
        g = __ldcs(gy + (n * Cout + co) * S + s);
      }
      gs[idx] = g;
    }
    __syncthreads();
    if (gx) {
      for (int idx = threadIdx.x; idx < kTileV * Cin; idx += blockDim.x) {
        const int vv = idx / Cin, ci = idx % Cin;
        const int64_t v = v0 + vv;
        if (v < total) {
          float s = 0.f;
          for (int co = 0; co < Cout; ++co) s = fmaf(gs[co * kTileV + vv], ws[co * Cin + ci], s);
          gx[v * Cin + ci] = from_f32<T>(s);
        }
      }
    }
    if (reg_mode) {
      if (threadIdx.x < groups * P) {
        const int p = threadIdx.x % P, grp = threadIdx.x / P;
        const int co = p / Cin1, ci = p % Cin1;
        for (int vv = grp; vv < kTileV; vv += groups) acc = fmaf(gs[co * kTileV + vv], xs[vv * Cin1 + ci], acc);
      }
    } else {
      for (int p = threadIdx.x; p < P; p += blockDim.x) {
        const int co = p / Cin1, ci = p % Cin1;
        float a = 0.f;
        for (int vv = 0; vv < kTileV; ++vv) a = fmaf(gs[co * kTileV + vv], xs[vv * Cin1 + ci], a);
        red[p] += a;
      }
    }
    __syncthreads();
  }
  if (reg_mode) {
    if (threadIdx.x < groups * P) red[threadIdx.x] = acc;
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      float t = 0.f;
      for (int gidx = 0; gidx < groups; ++gidx) t += red[gidx * P + p];
      partials[(int64_t)blockIdx.x * P + p] = t;
    }
  } else {
    for (int p = threadIdx.x; p < P; p += blockDim.x) partials[(int64_t)blockIdx.x * P + p] = red[p];
  }
}

// Fast path for the network's actual head (Cin = 16, Cout <= 4): one thread per voxel keeps the whole
// dW/db tile (Cout*(Cin+1) accumulators) in registers over a grid-stride loop; HBM-bound:
// 4*Cout B (gy) + 2*Cin B (x) read, 2*Cin B (gx) written per voxel.
template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 2)
conv1x1_bwd_small_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ gy, T* __restrict__ gx,
                         float* __restrict__ partials, int64_t N, int64_t S) {
  __shared__ float ws[COUT * CIN];
  __shared__ float red[kThreads / 32][COUT * (CIN + 1)];
  for (int i = threadIdx.x; i < COUT * CIN; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  float dw[COUT][CIN], db[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    db[co] = 0.f;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) dw[co][ci] = 0.f;
  }
  const int64_t total = N * S, stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 2;  // two voxels per thread in flight: 2 x (COUT gy loads + CIN/8 x loads) issued before the math
  for (int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < total; v0 += U * stride) {
    float g[U][COUT];
    Vec8<T> xv[U][CIN / 8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < total) {
        const int64_t n = sample_index(v, S), sp = v - n * S;
#pragma unroll
        for (int co = 0; co < COUT; ++co) g[u][co] = __ldcs(gy + (n * COUT + co) * S + sp);
#pragma unroll
        for (int c8 = 0; c8 < CIN / 8; ++c8) xv[u][c8].load(x + v * CIN + c8 * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v >= total) continue;
      float xf[CIN];
#pragma unroll
      for (int c8 = 0; c8 < CIN / 8; ++c8) {
        float f[8];
        xv[u][c8].get(f);
#pragma unroll
        for (int k = 0; k < 8; ++k) xf[c8 * 8 + k] = f[k];
      }
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        db[co] += g[u][co];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) dw[co][ci] = fmaf(g[u][co], xf[ci], dw[co][ci]);
      }
      if (gx) {
#pragma unroll
        for (int c8 = 0; c8 < CIN / 8; ++c8) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float a = 0.f;
#pragma unroll
            for (int co = 0; co < COUT; ++co) a = fmaf(g[u][co], ws[co * CIN + c8 * 8 + k], a);
            o[k] = a;
          }
          Vec8<T> ov;
          ov.set(o);
          ov.store(gx + v * CIN + c8 * 8);
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float t = warp_sum(dw[co][ci]);
      if (lane == 0) red[warp][co * (CIN + 1) + ci] = t;
    }
    const float t = warp_sum(db[co]);
    if (lane == 0) red[warp][co * (CIN + 1) + CIN] = t;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < COUT * (CIN + 1); p += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int wq = 0; wq < kThreads / 32; ++wq) t += red[wq][p];
    partials[(int64_t)blockIdx.x * (COUT * (CIN + 1)) + p] = t;
  }
}

__global__ void conv1x1_bwd_finalize_kernel(const float* __restrict__ partials, int nblocks, int Cin, int Cout,
                                            float* __restrict__ dw, float* __restrict__ db) {
  const int Cin1 = Cin + 1, P = Cout * Cin1;
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= P) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += (double)partials[(int64_t)b * P + p];
  s = warp_sum_d(s);
  if (lane != 0) return;
  const int co = p / Cin1, ci = p % Cin1;
  if (ci < Cin) { if (dw) dw[co * Cin + ci] = (float)s; }
  else if (db) db[co] = (float)s;
}

inline int bwd_blocks(int64_t total) {
  const int64_t tiles = (total + kTileV - 1) / kTileV;
  return (int)(tiles < kMaxPartialBlocks ? tiles : kMaxPartialBlocks);
}

// ---------------------------------------------------------------- linear layers (tiny)
// one warp per (b, o): lanes stride over I
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                  const float* __restrict__ dropmask, int relu, float* __restrict__ y, int B, int I, int O) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= B * O) return;
  const int bi = gw / O, o = gw % O;
  float s = 0.f;
  for (int i = lane; i < I; i += 32) s = fmaf(x[bi * I + i], w[o * I + i], s);
  s = warp_sum(s);
  if (lane == 0) {
    s += b ? b[o] : 0.f;
    if (relu) s = fmaxf(s, 0.f);
    if (dropmask) s *= dropmask[bi * O + o];
    y[bi * O + o] = s;
  }
}
__device__ __forceinline__ float linear_g(const float* y, const float* gy, const float* dropmask, int relu, int idx) {
  float g = gy[idx];
  if (dropmask) g *= dropmask[idx];
  // y already includes the dropout factor; y > 0 iff the pre-dropout activation was > 0 and kept
  if (relu && !(y[idx] > 0.f)) {
    // dropped-but-active units have mask 0 -> g is already 0; inactive units -> 0
    g = 0.f;
  }
  return g;
}
__global__ void linear_bwd_x_kernel(const float* __restrict__ w, const float* __restrict__ y, const float* __restrict__ gy,
                                    const float* __restrict__ dropmask, int relu, float* __restrict__ gx, int B, int I, int O) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * I) return;
  const int bi = idx / I, i = idx % I;
  float s = 0.f;
  for (int o = 0; o < O; ++o) s = fmaf(linear_g(y, gy, dropmask, relu, bi * O + o), w[o * I + i], s);
  gx[idx] = s;
}
__global__ void linear_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                                    const float* __restrict__ dropmask, int relu, float* __restrict__ dw,
                                    float* __restrict__ db, int B, int I, int O) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= O * (I + 1)) return;
  const int o = idx / (I + 1), i = idx % (I + 1);
  float s = 0.f;
  for (int bi = 0; bi < B; ++bi) {
    const float g = linear_g(y, gy, dropmask, relu, bi * O + o);
    s = fmaf(g, i < I ? x[bi * I + i] : 1.f, s);
  }
  if (i < I) dw[o * I + i] = s; else if (db) db[o] = s;
}

__global__ void ce_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int B, int C,
                               float* __restrict__ loss, float* __restrict__ dlogits) {
  // single block; B, C are tiny
  __shared__ float part[256];
  float local = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, logits[b * C + c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(logits[b * C + c] - m);
    const float lse = m + logf(s);
    const int y = (int)labels[b];
    local += lse - logits[b * C + y];
    if (dlogits)
      for (int c = 0; c < C; ++c)
        dlogits[b * C + c] = (expf(logits[b * C + c] - lse) - (c == y ? 1.f : 0.f)) / (float)B;
  }
  part[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)blockDim.x; ++i) t += part[i];
    loss[0] = t / (float)B;
  }
}

}  // namespace

// =========================================================================== exports
extern "C" int b200_conv1x1_fwd(int dtype, const void* x, const float* w, const float* bias, float* y, int64_t N, int64_t S,
                                int Cin, int Cout, int round_bf16, void* stream) {
  B200_REQUIRE(x && w && y && N > 0 && S > 0, B200_ERR_SHAPE, "conv1x1_fwd: bad arguments");
  B200_REQUIRE(Cin >= 8 && Cin % 8 == 0 && Cin <= 512, B200_ERR_UNSUPPORTED, "conv1x1_fwd: Cin=%d must be a multiple of 8 in [8,512]", Cin);
  B200_REQUIRE(Cout >= 1 && Cout <= kMaxCo, B200_ERR_UNSUPPORTED, "conv1x1_fwd: Cout=%d outside [1,%d]", Cout, kMaxCo);
  if (Cin == 16 && Cout >= 2 && Cout <= 4 && (dtype == B200_F32 || dtype == B200_BF16)) {
    const int g = b200_grid_for(N * S, kThreads * 2, B200_NUM_SMS * 2);  // 94 registers: two resident CTAs per SM, one wave
#define RUN_FS(T, CO) conv1x1_fwd_small_kernel<T, 16, CO><<<g, kThreads, 0, (cudaStream_t)stream>>>((const T*)x, w, bias, y, N, S, round_bf16)
    if (dtype == B200_F32) { if (Cout == 4) RUN_FS(float, 4); else if (Cout == 3) RUN_FS(float, 3); else RUN_FS(float, 2); }
    else { if (Cout == 4) RUN_FS(__nv_bfloat16, 4); else if (Cout == 3) RUN_FS(__nv_bfloat16, 3); else RUN_FS(__nv_bfloat16, 2); }
#undef RUN_FS
    B200_CHECK_LAUNCH("conv1x1_fwd_small");
    return B200_OK;
  }
  const size_t smem = (size_t)(Cout * Cin + Cout) * sizeof(float);
  const int grid = b200_grid_for(N * S, kThreads, B200_NUM_SMS * 8);
  B200_DISPATCH_DTYPE(dtype, T, (conv1x1_fwd_kernel<T><<<grid, kThreads, smem, (cudaStream_t)stream>>>((const T*)x, w, bias, y, N, S, Cin, Cout, round_bf16)));
  B200_CHECK_LAUNCH("conv1x1_fwd");
  return B200_OK;
}

extern "C" int64_t b200_conv1x1_partials_bytes(int Cin, int Cout) {
  return (int64_t)kMaxPartialBlocks * Cout * (Cin + 1) * sizeof(float);
}

extern "C" int b200_conv1x1_bwd(int dtype, const void* x, const float* w, const float* gy, void* gx, float* dw, float* db,
                                float* partials, int64_t N, int64_t S, int Cin, int Cout, void* stream) {
  B200_REQUIRE(x && w && gy && partials && N > 0 && S > 0, B200_ERR_SHAPE, "conv1x1_bwd: bad arguments");
  B200_REQUIRE(Cin >= 1 && Cin <= 512 && Cout >= 1 && Cout <= kMaxCo, B200_ERR_UNSUPPORTED, "conv1x1_bwd: unsupported channel counts %d -> %d", Cin, Cout);
  cudaStream_t st = (cudaStream_t)stream;
  const int P = Cout * (Cin + 1);
  const int groups = kThreads / P > 0 ? kThreads / P : 1;
  const size_t smem = (size_t)(Cout * Cin + kTileV * (Cin + 1) + Cout * kTileV + (size_t)(groups > 1 ? groups : 1) * P) * sizeof(float);
  B200_REQUIRE(smem <= 200 * 1024, B200_ERR_UNSUPPORTED, "conv1x1_bwd: channel counts too large for shared memory");
  int nblocks = bwd_blocks(N * S);
  if (Cin == 16 && (Cout == 4 || Cout == 2 || Cout == 3)) {
    nblocks = b200_grid_for(N * S, kThreads, B200_NUM_SMS * 2);  // 124 registers: two resident CTAs per SM, one wave
#define RUN_SMALL(T, CO) conv1x1_bwd_small_kernel<T, 16, CO><<<nblocks, kThreads, 0, st>>>((const T*)x, w, gy, (T*)gx, partials, N, S)
    if (dtype == B200_F32) { if (Cout == 4) RUN_SMALL(float, 4); else if (Cout == 3) RUN_SMALL(float, 3); else RUN_SMALL(float, 2); }
    else if (dtype == B200_BF16) { if (Cout == 4) RUN_SMALL(__nv_bfloat16, 4); else if (Cout == 3) RUN_SMALL(__nv_bfloat16, 3); else RUN_SMALL(__nv_bfloat16, 2); }
    else B200_FAIL(B200_ERR_UNSUPPORTED, "conv1x1_bwd: unknown dtype %d", dtype);
#undef RUN_SMALL
  } else if (dtype == B200_F32) {
    if (smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(conv1x1_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1x1_bwd_kernel<float><<<nblocks, kThreads, smem, st>>>((const float*)x, w, gy, (float*)gx, partials, N, S, Cin, Cout);
  } else if (dtype == B200_BF16) {
    if (smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(conv1x1_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1x1_bwd_kernel<__nv_bfloat16><<<nblocks, kThreads, smem, st>>>((const __nv_bfloat16*)x, w, gy, (__nv_bfloat16*)gx, partials, N, S, Cin, Cout);
  } else {
    B200_FAIL(B200_ERR_UNSUPPORTED, "conv1x1_bwd: unknown dtype %d", dtype);
  }
  B200_CHECK_LAUNCH("conv1x1_bwd");
  if (dw || db) {
    conv1x1_bwd_finalize_kernel<<<(P * 32 + 127) / 128, 128, 0, st>>>(partials, nblocks, Cin, Cout, dw, db);
    B200_CHECK_LAUNCH("conv1x1_bwd_finalize");
  }
  return B200_OK;
}

extern "C" int b200_linear_fwd(const float* x, const float* w, const float* b, const float* dropmask, int relu, float* y, int B,
                               int I, int O, void* stream) {
  B200_REQUIRE(x && w && y && B > 0 && I > 0 && O > 0, B200_ERR_SHAPE, "linear_fwd: bad arguments");
  const int warps = B * O;
  linear_fwd_kernel<<<(warps * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, w, b, dropmask, relu, y, B, I, O);
  B200_CHECK_LAUNCH("linear_fwd");
  return B200_OK;
}

extern "C" int b200_linear_bwd(const float* x, const float* w, const float* y, const float* gy, const float* dropmask, int relu,
                               float* gx, float* dw, float* db, int B, int I, int O, void* stream) {
  B200_REQUIRE(x && w && y && gy && B > 0 && I > 0 && O > 0, B200_ERR_SHAPE, "linear_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (gx) {
    linear_bwd_x_kernel<<<(B * I + 255) / 256, 256, 0, st>>>(w, y, gy, dropmask, relu, gx, B, I, O);
    B200_CHECK_LAUNCH("linear_bwd_x");
  }
  if (dw) {
    linear_bwd_w_kernel<<<(O * (I + 1) + 255) / 256, 256, 0, st>>>(x, y, gy, dropmask, relu, dw, db, B, I, O);
    B200_CHECK_LAUNCH("linear_bwd_w");
  }
  return B200_OK;
}

extern "C" int b200_ce_rows(const float* logits, const int64_t* labels, int B, int C, float* loss, float* dlogits, void* stream) {
  B200_REQUIRE(logits && labels && loss && B > 0 && C > 0, B200_ERR_SHAPE, "ce_rows: bad arguments");
  ce_rows_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, labels, B, C, loss, dlogits);
  B200_CHECK_LAUNCH("ce_rows");
  return B200_OK;
}
