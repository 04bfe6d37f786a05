// fp16x3 validation mode: data-movement kernels for (hi | lo) fp16 pairs. value = hi + lo carries ~22 mantissa bits,
// so the whole engine can be checked against the fp32 reference at rtol 1e-3 through all 53/174 layers (the fast
// path's single fp16 rounding per operand is amplified to O(1e-2) by the batch-statistics BatchNorm stack).
// The contraction itself is gemm_tc.cu with K-concatenated operands (mauv_gemm_x3_f16 / mauv_conv2d_im2col_x3_f16).
// Throughput is irrelevant here: one thread per element.
#include "common.cuh"

namespace {

__device__ __forceinline__ void split16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// weights: out row = [hi | hi | lo]; flat layout (plain GEMM / stem): part-major over the padded K;
// per-tap layout (im2col convs): for every tap (r, s): [hi(cin) | hi(cin) | lo(cin)].
__global__ void __launch_bounds__(256)
sample_x3_kernel(const float* __restrict__ mu, const float* __restrict__ rho, const float* __restrict__ eps,
                 uint64_t seed, uint32_t layer_id, uint32_t sample0, int cout, int cin, int kh, int kw, int kp,
                 int per_tap, float scale, long long n, __half* __restrict__ out, const unsigned int* __restrict__ sample_base) {
  const int g = blockIdx.y;
  if (sample_base) sample0 += *sample_base;
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int khw = kh * kw;
  const long long per_out = static_cast<long long>(cin) * khw;
  const long long co = e / per_out;
  const int rem = static_cast<int>(e - co * per_out);
  const int c = rem / khw, rs = rem - c * khw;
  const float z = eps ? eps[static_cast<long long>(g) * n + e]
                      : philox_normal(seed, layer_id, sample0 + g, static_cast<uint64_t>(e));
  const float v = scale * fmaf(softplus_ref(rho[e]), z, mu[e]);
  __half hi, lo;
  split16(v, hi, lo);
  if (per_tap) {
    const long long row = static_cast<long long>(khw) * 3 * cin;
    __half* o = out + (static_cast<long long>(g) * cout + co) * row + static_cast<long long>(rs) * 3 * cin + c;
    o[0] = hi; o[cin] = hi; o[2 * cin] = lo;
  } else {
    const long long row = 3LL * kp;
    __half* o = out + (static_cast<long long>(g) * cout + co) * row + static_cast<long long>(rs) * cin + c;
    o[0] = hi; o[kp] = hi; o[2 * kp] = lo;
  }
}

__global__ void __launch_bounds__(256)
stem_im2col_x3_kernel(const float* __restrict__ x, int B, int C, int H, int W, int kh, int kw, int stride, int pad,
                      int Ho, int Wo, int kp, __half* __restrict__ out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long rows = static_cast<long long>(B) * Ho * Wo;
  if (idx >= rows * kp) return;
  const long long row = idx / kp;
  const int k = static_cast<int>(idx - row * kp);
  const int q = static_cast<int>(row % Wo);
  const int p = static_cast<int>((row / Wo) % Ho);
  const int b = static_cast<int>(row / (static_cast<long long>(Wo) * Ho));
  float val = 0.f;
  if (k < kh * kw * C) {
    const int c = k % C, rs = k / C, s = rs % kw, r = rs / kw;
    const int h = p * stride - pad + r, w = q * stride - pad + s;
    if (h >= 0 && h < H && w >= 0 && w < W) val = x[((static_cast<long long>(b) * C + c) * H + h) * W + w];
  }
  __half hi, lo;
  split16(val, hi, lo);
  out[row * 2 * kp + k] = hi;
  out[row * 2 * kp + kp + k] = lo;
}

__device__ __forceinline__ float ld2(const __half* p, int C) { return __half2float(p[0]) + __half2float(p[C]); }

// out = relu?( y*ss [+ res] [+ y2*ss2] ) on (hi | lo) pairs: tensors are [G][M][2C]
__global__ void __launch_bounds__(256)
bn_act_x3_kernel(const __half* __restrict__ y, const float2* __restrict__ ss, const __half* __restrict__ res,
                 const __half* __restrict__ y2, const float2* __restrict__ ss2, int relu, long long M, int C,
                 long long total /*G*M*C*/, __half* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const long long gm = i / C;
  const int g = static_cast<int>(gm / M);
  const long long base = gm * 2 * C + c;
  const float2 s = ss[static_cast<long long>(g) * C + c];
  float v = fmaf(ld2(y + base, C), s.x, s.y);
  if (y2) {
    const float2 t = ss2[static_cast<long long>(g) * C + c];
    v += fmaf(ld2(y2 + base, C), t.x, t.y);
  }
  if (res) v += ld2(res + base, C);
  if (relu) v = fmaxf(v, 0.f);
  __half hi, lo;
  split16(v, hi, lo);
  out[base] = hi;
  out[base + C] = lo;
}

__global__ void __launch_bounds__(256)
bn_relu_maxpool_x3_kernel(const __half* __restrict__ y, const float2* __restrict__ ss, int imgs_per_sample, int H,
                          int W, int C, int Ho, int Wo, long long total, __half* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  long long t = i / C;
  const int q = static_cast<int>(t % Wo); t /= Wo;
  const int p = static_cast<int>(t % Ho); t /= Ho;
  const long long n = t;
  const float2 s = ss[static_cast<long long>(n / imgs_per_sample) * C + c];
  float m = -INFINITY;
  for (int dr = 0; dr < 3; ++dr) {
    const int h = p * 2 - 1 + dr;
    if (h < 0 || h >= H) continue;
    for (int ds = 0; ds < 3; ++ds) {
      const int w = q * 2 - 1 + ds;
      if (w < 0 || w >= W) continue;
      m = fmaxf(m, fmaf(ld2(y + ((n * H + h) * W + w) * 2 * C + c, C), s.x, s.y));
    }
  }
  m = fmaxf(m, 0.f);
  __half hi, lo;
  split16(m, hi, lo);
  const long long o = ((n * Ho + p) * Wo + q) * 2 * C + c;
  out[o] = hi;
  out[o + C] = lo;
}

__global__ void __launch_bounds__(256)
avgpool_x3_kernel(const __half* __restrict__ x, int HW, int C, long long total, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / C;
  const int c = static_cast<int>(i - n * C);
  float acc = 0.f;
  for (int k = 0; k < HW; ++k) acc += ld2(x + (n * HW + k) * 2 * C + c, C);
  out[i] = acc / static_cast<float>(HW);
}

}  // namespace

extern "C" {

int mauv_sample_weights_x3_f16(const float* mu, const float* rho, const float* eps, uint64_t seed, uint32_t layer_id,
                               uint32_t sample0, int G, int cout, int cin, int kh, int kw, int kp, int per_tap,
                               float scale, void* w_out, void* stream) {
  MAUV_CHECK_ARG(mu && rho && w_out && G >= 1, "mauv_sample_weights_x3_f16: bad argument");
  MAUV_CHECK_ARG(per_tap ? (cin % 64 == 0) : (kp % 64 == 0 && kp >= cin * kh * kw), "mauv_sample_weights_x3_f16: bad layout");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = static_cast<long long>(cout) * cin * kh * kw;
  const long long row = per_tap ? static_cast<long long>(kh) * kw * 3 * cin : 3LL * kp;
  MAUV_CUDA(cudaMemsetAsync(w_out, 0, static_cast<size_t>(G) * cout * row * sizeof(__half), st));
  dim3 grid(static_cast<unsigned>(ceil_div_i64(n, 256)), G);
  sample_x3_kernel<<<grid, 256, 0, st>>>(mu, rho, eps, seed, layer_id, sample0, cout, cin, kh, kw, kp, per_tap, scale, n,
                                         static_cast<__half*>(w_out), mauv_sample_base());
  MAUV_LAUNCH_CHECK("sample_x3_kernel");
  return MAUV_OK;
}

int mauv_stem_im2col_x3_f16(const float* x_nchw, int B, int C, int H, int W, int kh, int kw, int stride, int pad, int kp,
                            void* out, void* stream) {
  MAUV_CHECK_ARG(x_nchw && out && kp % 64 == 0 && kp >= kh * kw * C, "mauv_stem_im2col_x3_f16: bad argument");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  const long long work = static_cast<long long>(B) * Ho * Wo * kp;
  stem_im2col_x3_kernel<<<static_cast<unsigned>(ceil_div_i64(work, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nchw, B, C, H, W, kh, kw, stride, pad, Ho, Wo, kp, static_cast<__half*>(out));
  MAUV_LAUNCH_CHECK("stem_im2col_x3_kernel");
  return MAUV_OK;
}

int mauv_bn_act_x3_f16(const void* y2, const float* scale_shift, const void* residual2, const void* y2b,
                       const float* scale_shift2, int relu, int G, long long M, int C, void* out2, void* stream) {
  MAUV_CHECK_ARG(y2 && scale_shift && out2, "mauv_bn_act_x3_f16: null pointer");
  const long long total = static_cast<long long>(G) * M * C;
  bn_act_x3_kernel<<<static_cast<unsigned>(ceil_div_i64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(y2), reinterpret_cast<const float2*>(scale_shift), static_cast<const __half*>(residual2),
      static_cast<const __half*>(y2b), reinterpret_cast<const float2*>(scale_shift2), relu, M, C, total,
      static_cast<__half*>(out2));
  MAUV_LAUNCH_CHECK("bn_act_x3_kernel");
  return MAUV_OK;
}

int mauv_bn_relu_maxpool_x3_f16(const void* y2, const float* scale_shift, int G, int imgs_per_sample, int H, int W, int C,
                                void* out2, void* stream) {
  MAUV_CHECK_ARG(y2 && scale_shift && out2, "mauv_bn_relu_maxpool_x3_f16: null pointer");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(G) * imgs_per_sample * Ho * Wo * C;
  bn_relu_maxpool_x3_kernel<<<static_cast<unsigned>(ceil_div_i64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(y2), reinterpret_cast<const float2*>(scale_shift), imgs_per_sample, H, W, C, Ho, Wo, total,
      static_cast<__half*>(out2));
  MAUV_LAUNCH_CHECK("bn_relu_maxpool_x3_kernel");
  return MAUV_OK;
}

int mauv_avgpool_x3_f16(const void* x2, long long N, int HW, int C, float* out, void* stream) {
  MAUV_CHECK_ARG(x2 && out, "mauv_avgpool_x3_f16: null pointer");
  const long long total = N * C;
  avgpool_x3_kernel<<<static_cast<unsigned>(ceil_div_i64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x2), HW, C, total, out);
  MAUV_LAUNCH_CHECK("avgpool_x3_kernel");
  return MAUV_OK;
}

}  // extern "C"
