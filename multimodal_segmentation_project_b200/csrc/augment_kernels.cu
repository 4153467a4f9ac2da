// augment_kernels.cu — device versions of the five intensity augmentations the reference's combined_transform() applies
// (utils/dataloader.py:252-260: MONAI RandBiasFieldd, RandGaussianNoised, RandAdjustContrastd, RandHistogramShiftd,
// RandCoarseDropoutd).  MONAI (requirements.txt: monai>=1.2.0) is a third-party dependency whose source is not in the
// reference tree; the arithmetic below restates its published array transforms (monai/transforms/intensity/array.py):
//   RandBiasField     out = img * exp( sum_{i+j+k<=degree} c_ijk P_i(x) P_j(y) P_k(z) ), Legendre P, coordinates linspace(-1, 1, dim)
//                     per spatial axis, coefficients in MONAI's (i, j, k) loop order; evaluated in float64, rounded to float32
//   RandGaussianNoise out = img + N(mean, std'),  std' ~ U(0, std) drawn by the caller
//   AdjustContrast    out = ((img - min) / (max - min + 1e-7)) ** gamma * (max - min) + min           (float32)
//   HistogramShift    out = interp(img, ref * (max - min) + min, floating * (max - min) + min)        (float64 like np.interp)
//   CoarseDropout     img[hole] = fill, label[hole] = fill for every drawn box
// The random draws (which transform fires, gamma, control points, hole positions, the normal field) stay on the host /
// torch generator side: the kernels take them as arguments, so a given draw gives the same result as the numpy original.
// All passes are HBM-bound elementwise kernels on a 128^3 patch (8 MB): not on the step's critical path.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void legendre4(double x, double (&p)[4]) {
  p[0] = 1.0; p[1] = x; p[2] = 0.5 * (3.0 * x * x - 1.0); p[3] = 0.5 * (5.0 * x * x * x - 3.0 * x);
}

// coeff in MONAI's order: for i in 0..deg: for j in 0..deg-i: for k in 0..deg-i-j (degree <= 3: at most 20 terms)
__global__ void bias_field_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int D, int H, int W, int degree,
                                  const double* __restrict__ coeff) {
  __shared__ double cs[20];
  if (threadIdx.x < 20) cs[threadIdx.x] = coeff[threadIdx.x];
  __syncthreads();
  const int64_t S = (int64_t)D * H * W, total = (int64_t)C * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i % S;
    const int w = (int)(s % W), h = (int)((s / W) % H), d = (int)(s / ((int64_t)W * H));
    // np.linspace(-1, 1, n)[t] = -1 + t * (2 / (n - 1))  (n == 1: the single point -1)
    double pd[4], ph[4], pw[4];
    legendre4(D > 1 ? -1.0 + d * (2.0 / (D - 1)) : -1.0, pd);
    legendre4(H > 1 ? -1.0 + h * (2.0 / (H - 1)) : -1.0, ph);
    legendre4(W > 1 ? -1.0 + w * (2.0 / (W - 1)) : -1.0, pw);
    double f = 0.0;
    int t = 0;
    for (int a = 0; a <= degree; ++a)
      for (int b = 0; b <= degree - a; ++b)
        for (int c = 0; c <= degree - a - b; ++c) f += cs[t++] * pd[a] * ph[b] * pw[c];
    y[i] = (float)((double)x[i] * exp(f));
  }
}

__global__ void gaussian_noise_kernel(const float* __restrict__ x, const float* __restrict__ z, float* __restrict__ y, int64_t n, float mean,
                                      float std) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = x[i] + (mean + std * z[i]);
}

// order-preserving float <-> uint mapping for atomicMin / atomicMax
__device__ __forceinline__ unsigned int f2o(float f) { const unsigned int u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float o2f(unsigned int o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

__global__ void minmax_init_kernel(unsigned int* mm) { mm[0] = 0xffffffffu; mm[1] = 0u; }
__global__ void minmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ mm) {
  float lo = INFINITY, hi = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], f2o(lo)); atomicMax(&mm[1], f2o(hi)); }
}
__global__ void minmax_finish_kernel(const unsigned int* __restrict__ mm, float* __restrict__ out) { out[0] = o2f(mm[0]); out[1] = o2f(mm[1]); }

__global__ void adjust_contrast_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, const float* __restrict__ minmax, float gamma) {
  const float lo = minmax[0], range = minmax[1] - minmax[0];
  const float denom = (float)((double)range + 1e-7);      // float(img_range + epsilon) then float32 array / python float
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = powf((x[i] - lo) / denom, gamma) * range + lo;
}

// np.interp(x, xp, fp): left / right clamp, linear inside, float64
__global__ void histogram_shift_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, const float* __restrict__ minmax,
                                       const double* __restrict__ ref, const double* __restrict__ flt, int ncp) {
  __shared__ double xp[32], fp[32];
  const double lo = (double)minmax[0], range = (double)minmax[1] - (double)minmax[0];
  if (threadIdx.x < ncp) { xp[threadIdx.x] = ref[threadIdx.x] * range + lo; fp[threadIdx.x] = flt[threadIdx.x] * range + lo; }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    double r;
    if (range == 0.0) r = v;                       // MONAI: "image min and max are equal": returned unchanged
    else if (v <= xp[0]) r = fp[0];
    else if (v >= xp[ncp - 1]) r = fp[ncp - 1];
    else {
      int j = 0;
      while (j + 2 < ncp && v >= xp[j + 1]) ++j;
      const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
      r = slope * (v - xp[j]) + fp[j];
    }
    y[i] = (float)r;
  }
}

// holes[h] = {d0, d1, h0, h1, w0, w1}; every channel of img and label
__global__ void coarse_dropout_kernel(float* __restrict__ img, int64_t* __restrict__ label, int C, int D, int H, int W, const int* __restrict__ holes,
                                      int nholes, float fill) {
  const int64_t S = (int64_t)D * H * W, total = (int64_t)C * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i % S;
    const int w = (int)(s % W), h = (int)((s / W) % H), d = (int)(s / ((int64_t)W * H));
    bool hit = false;
    for (int q = 0; q < nholes; ++q) {
      const int* b = holes + 6 * q;
      hit |= d >= b[0] && d < b[1] && h >= b[2] && h < b[3] && w >= b[4] && w < b[5];
    }
    if (hit) {
      if (img) img[i] = fill;
      if (label) label[i] = (int64_t)fill;
    }
  }
}

inline int grid_for(int64_t n) { return b200_grid_for(n, kThreads, B200_NUM_SMS * 8); }

}  // namespace

extern "C" int b200_aug_bias_field(const float* x, float* y, int C, int D, int H, int W, int degree, const double* coeff, void* stream) {
  B200_REQUIRE(x && y && coeff && C > 0 && D > 0 && H > 0 && W > 0, B200_ERR_SHAPE, "aug_bias_field: bad arguments");
  B200_REQUIRE(degree >= 0 && degree <= 3, B200_ERR_UNSUPPORTED, "aug_bias_field: degree 0..3 (MONAI default 3)");
  bias_field_kernel<<<grid_for((int64_t)C * D * H * W), kThreads, 0, (cudaStream_t)stream>>>(x, y, C, D, H, W, degree, coeff);
  B200_CHECK_LAUNCH("aug_bias_field");
  return B200_OK;
}
extern "C" int b200_aug_gaussian_noise(const float* x, const float* z, float* y, int64_t n, float mean, float std, void* stream) {
  B200_REQUIRE(x && z && y && n > 0, B200_ERR_SHAPE, "aug_gaussian_noise: bad arguments");
  gaussian_noise_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(x, z, y, n, mean, std);
  B200_CHECK_LAUNCH("aug_gaussian_noise");
  return B200_OK;
}
/* minmax[2] float on the device; workspace = 2 x uint32 */
extern "C" int b200_minmax_f32(const float* x, int64_t n, float* minmax, void* workspace, void* stream) {
  B200_REQUIRE(x && minmax && workspace && n > 0, B200_ERR_SHAPE, "minmax_f32: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  minmax_init_kernel<<<1, 1, 0, st>>>((unsigned int*)workspace);
  minmax_kernel<<<grid_for(n), kThreads, 0, st>>>(x, n, (unsigned int*)workspace);
  minmax_finish_kernel<<<1, 1, 0, st>>>((const unsigned int*)workspace, minmax);
  B200_CHECK_LAUNCH("minmax_f32");
  return B200_OK;
}
extern "C" int b200_aug_adjust_contrast(const float* x, float* y, int64_t n, const float* minmax, float gamma, void* stream) {
  B200_REQUIRE(x && y && minmax && n > 0, B200_ERR_SHAPE, "aug_adjust_contrast: bad arguments");
  adjust_contrast_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(x, y, n, minmax, gamma);
  B200_CHECK_LAUNCH("aug_adjust_contrast");
  return B200_OK;
}
extern "C" int b200_aug_histogram_shift(const float* x, float* y, int64_t n, const float* minmax, const double* ref, const double* floating, int ncp,
                                        void* stream) {
  B200_REQUIRE(x && y && minmax && ref && floating && n > 0, B200_ERR_SHAPE, "aug_histogram_shift: bad arguments");
  B200_REQUIRE(ncp >= 2 && ncp <= 32, B200_ERR_UNSUPPORTED, "aug_histogram_shift: 2..32 control points");
  histogram_shift_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(x, y, n, minmax, ref, floating, ncp);
  B200_CHECK_LAUNCH("aug_histogram_shift");
  return B200_OK;
}
extern "C" int b200_aug_coarse_dropout(float* img, int64_t* label, int C, int D, int H, int W, const int* holes, int nholes, float fill, void* stream) {
  B200_REQUIRE((img || label) && holes && nholes > 0 && C > 0 && D > 0 && H > 0 && W > 0, B200_ERR_SHAPE, "aug_coarse_dropout: bad arguments");
  coarse_dropout_kernel<<<grid_for((int64_t)C * D * H * W), kThreads, 0, (cudaStream_t)stream>>>(img, label, C, D, H, W, holes, nholes, fill);
  B200_CHECK_LAUNCH("aug_coarse_dropout");
  return B200_OK;
}
