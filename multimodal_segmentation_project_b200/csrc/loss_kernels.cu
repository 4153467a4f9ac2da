// loss_kernels.cu — single-pass softmax + CE + Dice/Tversky (+ KD-KL) forward and backward.
//
// Replaces the ~20 ATen kernels behind utils/metrics.py:14-40 (combined_loss), :137-156
// (tversky_loss), :158-167 (combined_ce_tversky_loss), :169-190 (distillation_loss).
// Formulas: SURVEY.md Appendix D.  HBM-bound: fwd reads 4C+8 B/voxel (+4C with a teacher),
// bwd reads the same and writes 4C B/voxel.  Each thread handles 4 consecutive voxels with
// 128-bit loads per class plane; partial sums go thread -> warp shuffle -> block -> one fp64
// atomic per quantity per block.
#include "common.cuh"

namespace {

constexpr int kLossThreads = 256;

template <int CMAX>
struct VoxelSoftmax {
  float p[CMAX];
  float lse;
  __device__ __forceinline__ void compute(const float (&z)[CMAX], int C, float inv_t) {
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) m = fmaxf(m, z[c] * inv_t);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        p[c] = expf(z[c] * inv_t - m);
        s += p[c];
      }
    const float inv = 1.f / s;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) p[c] *= inv;
    lse = m + logf(s);
  }
};

// sums layout (double): [0]=CE_sum [1]=KL_sum [2],[3] unused; then per class k: [4+4k+0]=I [4+4k+1]=P [4+4k+2]=T
template <int CMAX, int VEC, bool HAS_KD>
__global__ void __launch_bounds__(kLossThreads)
seg_loss_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ teacher,
                    const int64_t* __restrict__ target, int64_t N, int C, int64_t S, float inv_temp,
                    double* __restrict__ sums) {
  float ce = 0.f, kl = 0.f;
  float accI[CMAX], accP[CMAX], accT[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) accI[c] = accP[c] = accT[c] = 0.f;

  const int64_t groups_per_sample = S / VEC;  // VEC divides S (checked on host)
  const int64_t total_groups = N * groups_per_sample;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_groups;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = g / groups_per_sample;
    const int64_t s0 = (g - n * groups_per_sample) * VEC;
    const float* zp = logits + (n * C) * S + s0;
    float z[VEC][CMAX];
    float zt[VEC][CMAX];
    long long y[VEC];
    if (VEC == 4) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          const float4 v = __ldcs(reinterpret_cast<const float4*>(zp + c * S));
          z[0][c] = v.x; z[1 % VEC][c] = v.y; z[2 % VEC][c] = v.z; z[3 % VEC][c] = v.w;
          if (HAS_KD) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(teacher + (n * C + c) * S + s0));
            zt[0][c] = t.x; zt[1 % VEC][c] = t.y; zt[2 % VEC][c] = t.z; zt[3 % VEC][c] = t.w;
          }
        }
      const longlong2 y01 = __ldcs(reinterpret_cast<const longlong2*>(target + n * S + s0));
      const longlong2 y23 = __ldcs(reinterpret_cast<const longlong2*>(target + n * S + s0 + 2));
      y[0] = y01.x; y[1 % VEC] = y01.y; y[2 % VEC] = y23.x; y[3 % VEC] = y23.y;
    } else {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          z[0][c] = zp[c * S];
          if (HAS_KD) zt[0][c] = teacher[(n * C + c) * S + s0];
        }
      y[0] = target[n * S + s0];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      VoxelSoftmax<CMAX> sm;
      sm.compute(z[j], C, 1.f);
      const int yy = (int)y[j];
      float zy = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          const bool hit = (c == yy);
          accP[c] += sm.p[c];
          if (hit) { accI[c] += sm.p[c]; accT[c] += 1.f; zy = z[j][c]; }
        }
      ce += sm.lse - zy;
      if (HAS_KD) {
        VoxelSoftmax<CMAX> ss, st;
        ss.compute(z[j], C, inv_temp);
        st.compute(zt[j], C, inv_temp);
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) {
            const float logq = zt[j][c] * inv_temp - st.lse;
            const float logp = z[j][c] * inv_temp - ss.lse;
            const float q = st.p[c];
            kl += (q > 0.f) ? q * (logq - logp) : 0.f;
          }
      }
    }
  }

  // block reduction: 2 + 3C quantities
  __shared__ float red[kLossThreads / 32][2 + 3 * CMAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ce = warp_sum(ce);
  kl = warp_sum(kl);
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    accI[c] = warp_sum(accI[c]);
    accP[c] = warp_sum(accP[c]);
    accT[c] = warp_sum(accT[c]);
  }
  if (lane == 0) {
    red[warp][0] = ce;
    red[warp][1] = kl;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      red[warp][2 + 3 * c] = accI[c];
      red[warp][3 + 3 * c] = accP[c];
      red[warp][4 + 3 * c] = accT[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 + 3 * CMAX) {
    double t = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) t += (double)red[w][threadIdx.x];
    const int q = threadIdx.x;
    if (q < 2) {
      if (q == 0 || HAS_KD) atomicAdd(&sums[q], t);
    } else {
      const int c = (q - 2) / 3, f = (q - 2) % 3;
      if (c < C) atomicAdd(&sums[4 + 4 * c + f], t);
    }
  }
}

// coef layout: [0] = w_ce / Nvox, [1] = kd factor (1-a) * T / (Nvox*C), [2 + c] = a_c, [2 + C + c] = b_c
__global__ void seg_loss_finalize_kernel(const double* __restrict__ sums, int mode, float alpha, float beta,
                                         float kd_alpha, float temperature, int has_kd, int64_t nvox, int C,
                                         float* __restrict__ loss, float* __restrict__ coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double w_ce = 0.0, w_reg = 0.0;
  bool tversky = false;
  switch (mode) {
    case B200_LOSS_DICE_CE: w_ce = 1.0; w_reg = 1.0; break;
    case B200_LOSS_TVERSKY: w_reg = 1.0; tversky = true; break;
    case B200_LOSS_CE_TVERSKY: w_ce = 0.3; w_reg = 0.7; tversky = true; break;
    case B200_LOSS_DICE: w_reg = 1.0; break;
    case B200_LOSS_CE: w_ce = 1.0; break;
  }
  double seg_w = has_kd ? (double)kd_alpha : 1.0;
  w_ce *= seg_w;
  w_reg *= seg_w;
  const double ce_mean = sums[0] / (double)nvox;
  double region = 0.0;
  const double cm1 = (double)(C - 1);
  for (int k = 0; k < C; ++k) {
    double a = 0.0, b = 0.0;
    if (k >= 1 && w_reg != 0.0) {
      const double I = sums[4 + 4 * k + 0], P = sums[4 + 4 * k + 1], T = sums[4 + 4 * k + 2];
      if (!tversky) {
        const double eps = 1e-5;
        const double num = 2.0 * I + eps, den = P + T + eps;
        region += 1.0 - num / den;
        a = -2.0 / (cm1 * den);
        b = num / (cm1 * den * den);
      } else {
        const double eps = 1e-6;
        const double al = (double)alpha, be = (double)beta;
        const double num = I + eps;
        const double den = I + al * (P - I) + be * (T - I) + eps;
        region += 1.0 - num / den;
        a = (-den + num * (1.0 - al - be)) / (cm1 * den * den);
        b = num * al / (cm1 * den * den);
      }
    }
    coef[2 + k] = (float)(w_reg * a);
    coef[2 + C + k] = (float)(w_reg * b);
  }
  region /= cm1;
  double total = w_ce * ce_mean + w_reg * region;
  double kdf = 0.0;
  if (has_kd) {
    const double T = (double)temperature;
    const double kl_mean = sums[1] / ((double)nvox * (double)C);
    total += (1.0 - (double)kd_alpha) * T * T * kl_mean;
    kdf = (1.0 - (double)kd_alpha) * T / ((double)nvox * (double)C);
  }
  coef[0] = (float)(w_ce / (double)nvox);
  coef[1] = (float)kdf;
  loss[0] = (float)total;
}

template <int CMAX, int VEC, bool HAS_KD>
__global__ void __launch_bounds__(kLossThreads)
seg_loss_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ teacher,
                    const int64_t* __restrict__ target, const float* __restrict__ coef,
                    const float* __restrict__ gout, float inv_temp, int64_t N, int C, int64_t S,
                    float* __restrict__ dlogits) {
  const float go = gout[0];
  const float w_ce = coef[0] * go, w_kd = coef[1] * go;
  float ca[CMAX], cb[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    ca[c] = (c < C) ? coef[2 + c] * go : 0.f;
    cb[c] = (c < C) ? coef[2 + C + c] * go : 0.f;
  }
  const int64_t groups_per_sample = S / VEC;
  const int64_t total_groups = N * groups_per_sample;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_groups;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = g / groups_per_sample;
    const int64_t s0 = (g - n * groups_per_sample) * VEC;
    const float* zp = logits + (n * C) * S + s0;
    float z[VEC][CMAX], zt[VEC][CMAX], out[VEC][CMAX];
    long long y[VEC];
    if (VEC == 4) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          const float4 v = __ldcs(reinterpret_cast<const float4*>(zp + c * S));
          z[0][c] = v.x; z[1 % VEC][c] = v.y; z[2 % VEC][c] = v.z; z[3 % VEC][c] = v.w;
          if (HAS_KD) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(teacher + (n * C + c) * S + s0));
            zt[0][c] = t.x; zt[1 % VEC][c] = t.y; zt[2 % VEC][c] = t.z; zt[3 % VEC][c] = t.w;
          }
        }
      const longlong2 y01 = __ldcs(reinterpret_cast<const longlong2*>(target + n * S + s0));
      const longlong2 y23 = __ldcs(reinterpret_cast<const longlong2*>(target + n * S + s0 + 2));
      y[0] = y01.x; y[1 % VEC] = y01.y; y[2 % VEC] = y23.x; y[3 % VEC] = y23.y;
    } else {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          z[0][c] = zp[c * S];
          if (HAS_KD) zt[0][c] = teacher[(n * C + c) * S + s0];
        }
      y[0] = target[n * S + s0];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      VoxelSoftmax<CMAX> sm;
      sm.compute(z[j], C, 1.f);
      const int yy = (int)y[j];
      float w[CMAX];
      float pw = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          w[c] = (c == yy ? ca[c] : 0.f) + cb[c];
          pw += sm.p[c] * w[c];
        }
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          const float t = (c == yy) ? 1.f : 0.f;
          out[j][c] = w_ce * (sm.p[c] - t) + sm.p[c] * (w[c] - pw);
        }
      if (HAS_KD) {
        VoxelSoftmax<CMAX> ss, st;
        ss.compute(z[j], C, inv_temp);
        st.compute(zt[j], C, inv_temp);
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) out[j][c] += w_kd * (ss.p[c] - st.p[c]);
      }
    }
    float* dp = dlogits + (n * C) * S + s0;
    if (VEC == 4) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C)
          __stcs(reinterpret_cast<float4*>(dp + c * S),
                 make_float4(out[0][c], out[1 % VEC][c], out[2 % VEC][c], out[3 % VEC][c]));
    } else {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) dp[c * S] = out[0][c];
    }
  }
}

template <bool HAS_KD>
int launch_fwd(const float* logits, const float* teacher, const int64_t* target, float temperature, int64_t N, int C,
               int64_t S, double* sums, cudaStream_t st) {
  const bool vec = (S % 4 == 0) && b200_aligned(logits, 16) && b200_aligned(target, 16) &&
                   (!HAS_KD || b200_aligned(teacher, 16));
  const int64_t groups = N * (vec ? S / 4 : S);
  const int grid = b200_grid_for(groups, kLossThreads, B200_NUM_SMS * 8);
  const float inv_t = 1.f / temperature;
#define LAUNCH(CMAX)                                                                                         \
  if (vec) seg_loss_fwd_kernel<CMAX, 4, HAS_KD><<<grid, kLossThreads, 0, st>>>(logits, teacher, target, N, C, S, inv_t, sums); \
  else seg_loss_fwd_kernel<CMAX, 1, HAS_KD><<<grid, kLossThreads, 0, st>>>(logits, teacher, target, N, C, S, inv_t, sums)
  if (C <= 4) { LAUNCH(4); } else if (C <= 8) { LAUNCH(8); } else { LAUNCH(16); }
#undef LAUNCH
  B200_CHECK_LAUNCH("seg_loss_fwd");
  return B200_OK;
}

template <bool HAS_KD>
int launch_bwd(const float* logits, const float* teacher, const int64_t* target, const float* coef, const float* gout,
               float temperature, int64_t N, int C, int64_t S, float* dlogits, cudaStream_t st) {
  const bool vec = (S % 4 == 0) && b200_aligned(logits, 16) && b200_aligned(target, 16) && b200_aligned(dlogits, 16) &&
                   (!HAS_KD || b200_aligned(teacher, 16));
  const int64_t groups = N * (vec ? S / 4 : S);
  const int grid = b200_grid_for(groups, kLossThreads, B200_NUM_SMS * 8);
  const float inv_t = 1.f / temperature;
#define LAUNCH(CMAX)                                                                                          \
  if (vec) seg_loss_bwd_kernel<CMAX, 4, HAS_KD><<<grid, kLossThreads, 0, st>>>(logits, teacher, target, coef, gout, inv_t, N, C, S, dlogits); \
  else seg_loss_bwd_kernel<CMAX, 1, HAS_KD><<<grid, kLossThreads, 0, st>>>(logits, teacher, target, coef, gout, inv_t, N, C, S, dlogits)
  if (C <= 4) { LAUNCH(4); } else if (C <= 8) { LAUNCH(8); } else { LAUNCH(16); }
#undef LAUNCH
  B200_CHECK_LAUNCH("seg_loss_bwd");
  return B200_OK;
}

int check_loss_args(const void* logits, const void* target, int64_t N, int C, int64_t S) {
  B200_REQUIRE(logits && target, B200_ERR_SHAPE, "seg_loss: null pointer");
  B200_REQUIRE(N > 0 && S > 0, B200_ERR_SHAPE, "seg_loss: empty input N=%lld S=%lld", (long long)N, (long long)S);
  B200_REQUIRE(C >= 2 && C <= 16, B200_ERR_UNSUPPORTED, "seg_loss: C=%d outside [2,16]", C);
  return B200_OK;
}

}  // namespace

extern "C" int b200_kd_loss_fwd(const float* student, const float* teacher, const int64_t* target, float temperature,
                                int64_t N, int C, int64_t S, double* sums, void* stream) {
  int rc = check_loss_args(student, target, N, C, S);
  if (rc) return rc;
  B200_REQUIRE(sums, B200_ERR_SHAPE, "seg_loss: null sums");
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (4 + 4 * C), st));
  if (teacher) {
    B200_REQUIRE(temperature > 0.f, B200_ERR_SHAPE, "kd_loss: temperature must be > 0");
    return launch_fwd<true>(student, teacher, target, temperature, N, C, S, sums, st);
  }
  return launch_fwd<false>(student, nullptr, target, 1.f, N, C, S, sums, st);
}

extern "C" int b200_seg_loss_fwd(const float* logits, const int64_t* target, int64_t N, int C, int64_t S, double* sums,
                                 void* stream) {
  return b200_kd_loss_fwd(logits, nullptr, target, 1.f, N, C, S, sums, stream);
}

extern "C" int b200_seg_loss_finalize(const double* sums, int mode, float alpha, float beta, float kd_alpha,
                                      float temperature, int has_kd, int64_t N, int C, int64_t S, float* loss,
                                      float* coef, void* stream) {
  B200_REQUIRE(sums && loss && coef, B200_ERR_SHAPE, "seg_loss_finalize: null pointer");
  B200_REQUIRE(mode >= B200_LOSS_DICE_CE && mode <= B200_LOSS_CE, B200_ERR_UNSUPPORTED, "seg_loss_finalize: mode %d", mode);
  B200_REQUIRE(C >= 2 && C <= 16, B200_ERR_UNSUPPORTED, "seg_loss_finalize: C=%d outside [2,16]", C);
  seg_loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, mode, alpha, beta, kd_alpha, temperature, has_kd,
                                                               N * S, C, loss, coef);
  B200_CHECK_LAUNCH("seg_loss_finalize");
  return B200_OK;
}

extern "C" int b200_seg_loss_bwd(const float* logits, const float* teacher, const int64_t* target, const float* coef,
                                 const float* gout, float temperature, int64_t N, int C, int64_t S, float* dlogits,
                                 void* stream) {
  int rc = check_loss_args(logits, target, N, C, S);
  if (rc) return rc;
  B200_REQUIRE(coef && gout && dlogits, B200_ERR_SHAPE, "seg_loss_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (teacher) return launch_bwd<true>(logits, teacher, target, coef, gout, temperature, N, C, S, dlogits, st);
  return launch_bwd<false>(logits, nullptr, target, coef, gout, 1.f, N, C, S, dlogits, st);
}
