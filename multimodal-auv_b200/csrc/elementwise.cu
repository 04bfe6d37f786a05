// K3: the HBM-bound passes around the tensor-core convs: stem im2col, BatchNorm
// (train-mode batch statistics, as the reference keeps the model in .train() for every
// MC pass: inference/predictors.py:27, train/multimodal.py:60,232), ReLU, residual add,
// 3x3/2 max-pool and global average pool (torchvision resnet.py:143-165, 266-282).
// All activations are NHWC fp16, 16-byte vector accesses, statistics in fp32/fp64.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------
// Stem: NCHW fp32 image -> explicit im2col matrix [B*Ho*Wo][k_pad] fp16 for the 7x7/2
// conv (Cin = 1 or 3 cannot feed TMA: a pixel is < 16 bytes). K order (r, s, c). The
// matrix depends only on the input batch, so one build serves all S MC samples.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ x, int B, int C, int H, int W, int kh, int kw,
                   int stride, int pad, int Ho, int Wo, int k_pad, __half* __restrict__ out) {
  // one thread per (row, 8-wide k chunk)
  const int chunks = k_pad / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long rows = static_cast<long long>(B) * Ho * Wo;
  if (idx >= rows * chunks) return;
  const long long row = idx / chunks;
  const int ch = static_cast<int>(idx - row * chunks);
  const int q = static_cast<int>(row % Wo);
  const int p = static_cast<int>((row / Wo) % Ho);
  const int b = static_cast<int>(row / (static_cast<long long>(Wo) * Ho));
  const int K = kh * kw * C;
  __align__(16) __half v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = ch * 8 + j;
    float val = 0.f;
    if (k < K) {
      const int c = k % C;
      const int rs = k / C;
      const int s = rs % kw;
      const int r = rs / kw;
      const int h = p * stride - pad + r;
      const int w = q * stride - pad + s;
      if (h >= 0 && h < H && w >= 0 && w < W)
        val = __ldg(x + ((static_cast<long long>(b) * C + c) * H + h) * W + w);
    }
    v[j] = __float2half_rn(val);
  }
  *reinterpret_cast<uint4*>(out + row * k_pad + ch * 8) = *reinterpret_cast<const uint4*>(v);
}

// ---------------------------------------------------------------------------
// BN finalize: reduce the per-tile (sum, sumsq) partials written by the conv epilogue,
// produce per-(sample, channel) scale/shift, and replay the running-stat updates of
// the S sequential reference passes (momentum 0.1, unbiased variance).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
bn_finalize_kernel(const float2* __restrict__ partial, int G, int m_tiles, int C, long long count,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                   float2* __restrict__ scale_shift, float2* __restrict__ batch_stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float rm = running_mean ? running_mean[c] : 0.f;
  float rv = running_var ? running_var[c] : 1.f;
  const float ga = gamma ? gamma[c] : 1.f;
  const float be = beta ? beta[c] : 0.f;
  for (int g = 0; g < G; ++g) {
    double s1 = 0.0, s2 = 0.0;
    const float2* pp = partial + static_cast<long long>(g) * m_tiles * C + c;
    for (int t = 0; t < m_tiles; ++t) {
      const float2 v = pp[static_cast<long long>(t) * C];
      s1 += v.x;
      s2 += v.y;
    }
    const double mean = s1 / static_cast<double>(count);
    double var = s2 / static_cast<double>(count) - mean * mean;  // biased (normalisation)
    if (var < 0.0) var = 0.0;
    const float inv_std = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float sc = ga * inv_std;
    scale_shift[static_cast<long long>(g) * C + c] = make_float2(sc, be - static_cast<float>(mean) * sc);
    if (batch_stats) batch_stats[static_cast<long long>(g) * C + c] = make_float2(static_cast<float>(mean), static_cast<float>(var));
    const double unbiased = count > 1 ? var * static_cast<double>(count) / static_cast<double>(count - 1) : var;
    rm = (1.f - momentum) * rm + momentum * static_cast<float>(mean);
    rv = (1.f - momentum) * rv + momentum * static_cast<float>(unbiased);
  }
  if (running_mean) running_mean[c] = rm;
  if (running_var) running_var[c] = rv;
}

// The partial buffer is [G][m_tiles][C]; one thread per channel reads with stride C ->
// coalesced across the warp. For the big early layers (m_tiles up to 32768) split the
// tile range over blockIdx.y and combine in a second tiny pass.
__global__ void __launch_bounds__(128)
bn_partial_reduce_kernel(const float2* __restrict__ partial, int G, int m_tiles, int C, int splits,
                         float2* __restrict__ out /*[G][splits][C]*/) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int sp = blockIdx.y;
  const int g = blockIdx.z;
  if (c >= C) return;
  const int per = (m_tiles + splits - 1) / splits;
  const int t0 = sp * per;
  const int t1 = min(m_tiles, t0 + per);
  float s1 = 0.f, s2 = 0.f;
  float c1 = 0.f, c2 = 0.f;  // Kahan compensation
  const float2* pp = partial + static_cast<long long>(g) * m_tiles * C + c;
  for (int t = t0; t < t1; ++t) {
    const float2 v = pp[static_cast<long long>(t) * C];
    float y = v.x - c1; float tt = s1 + y; c1 = (tt - s1) - y; s1 = tt;
    y = v.y - c2; tt = s2 + y; c2 = (tt - s2) - y; s2 = tt;
  }
  out[(static_cast<long long>(g) * splits + sp) * C + c] = make_float2(s1, s2);
}

// ---------------------------------------------------------------------------
// BN apply (+ optional residual, + optional second BN for the downsample branch) + ReLU.
//   out = relu( y*sc + sh  [+ res]  [+ y2*sc2 + sh2] )
// y, res, y2, out: [G][M][C] fp16 ; scale_shift: [G][C] float2.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__global__ void __launch_bounds__(256)
bn_act_kernel(const uint4* __restrict__ y, const float2* __restrict__ ss, const uint4* __restrict__ res,
              const uint4* __restrict__ y2, const float2* __restrict__ ss2, int relu,
              long long per_sample_vec /* M*C/8 */, int C, long long total_vec, uint4* __restrict__ out) {
  const int cvec = C / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i / per_sample_vec);
    const int c0 = static_cast<int>(i % cvec) * 8;
    float f[8];
    unpack8(__ldg(y + i), f);
    const float2* s = ss + static_cast<long long>(g) * C + c0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 t = __ldg(s + j);
      f[j] = fmaf(f[j], t.x, t.y);
    }
    if (y2) {
      float f2[8];
      unpack8(__ldg(y2 + i), f2);
      const float2* s2 = ss2 + static_cast<long long>(g) * C + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 t = __ldg(s2 + j);
        f[j] += fmaf(f2[j], t.x, t.y);
      }
    }
    if (res) {
      float fr[8];
      unpack8(__ldg(res + i), fr);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += fr[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    out[i] = pack8(f);
  }
}

// Stem tail: BN + ReLU + 3x3/2 max-pool (pad 1) in one pass.
// y: [G*B][H][W][C] fp16 raw conv output -> out: [G*B][Ho][Wo][C].
__global__ void __launch_bounds__(256)
bn_relu_maxpool_kernel(const uint4* __restrict__ y, const float2* __restrict__ ss, int imgs_per_sample,
                       int H, int W, int C, int Ho, int Wo, long long total_vec, uint4* __restrict__ out) {
  const int cvec = C / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % cvec);
    long long t = i / cvec;
    const int q = static_cast<int>(t % Wo); t /= Wo;
    const int p = static_cast<int>(t % Ho); t /= Ho;
    const long long n = t;  // image index in [0, G*B)
    const int g = static_cast<int>(n / imgs_per_sample);
    float sc[8], sh[8], m[8];
    const float2* s = ss + static_cast<long long>(g) * C + cv * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 tt = __ldg(s + j);
      sc[j] = tt.x; sh[j] = tt.y; m[j] = -INFINITY;
    }
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
      const int h = p * 2 - 1 + dr;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int ds = 0; ds < 3; ++ds) {
        const int w = q * 2 - 1 + ds;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(__ldg(y + ((n * H + h) * W + w) * cvec + cv), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], fmaf(f[j], sc[j], sh[j]));
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], 0.f);  // relu(max) == max(relu)
    out[i] = pack8(m);
  }
}

// Global average pool: x [N][HW][C] fp16 -> out [N][C] fp32 (head input).
__global__ void __launch_bounds__(256)
avgpool_kernel(const __half* __restrict__ x, int HW, int C, long long total /*N*C*/, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / C;
  const int c = static_cast<int>(i - n * C);
  const __half* p = x + n * HW * C + c;
  float acc = 0.f;
  for (int k = 0; k < HW; ++k) acc += __half2float(p[static_cast<long long>(k) * C]);
  out[i] = acc / static_cast<float>(HW);
}

// NCHW fp32 -> NHWC fp16 (generic layer-level entry: Conv2dReparameterization.forward)
__global__ void __launch_bounds__(256)
nchw_to_nhwc_f16_kernel(const float* __restrict__ x, int C, int HW, int c_pad, long long total /*N*HW*c_pad*/,
                        __half* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % c_pad);
  const long long t = i / c_pad;
  const int hw = static_cast<int>(t % HW);
  const long long n = t / HW;
  out[i] = (c < C) ? __float2half_rn(x[(n * C + c) * HW + hw]) : __float2half_rn(0.f);
}
// NHWC fp16 -> NCHW fp32
__global__ void __launch_bounds__(256)
nhwc_to_nchw_f32_kernel(const __half* __restrict__ x, int C, int HW, long long total /*N*C*HW*/,
                        float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int hw = static_cast<int>(i % HW);
  const long long t = i / HW;
  const int c = static_cast<int>(t % C);
  const long long n = t / C;
  out[i] = __half2float(x[(n * HW + hw) * C + c]);
}

inline unsigned grid_for(long long work, int block, int waves = 8) {
  long long blocks = ceil_div_i64(work, block);
  const long long cap = static_cast<long long>(mauv_num_sms()) * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

}  // namespace

extern "C" {

int mauv_stem_im2col_f16(const float* x_nchw, int B, int C, int H, int W, int kh, int kw, int stride,
                         int pad, int k_pad, void* out, void* stream) {
  MAUV_CHECK_ARG(x_nchw && out, "mauv_stem_im2col_f16: null pointer");
  MAUV_CHECK_ARG(k_pad % 8 == 0 && k_pad >= kh * kw * C, "mauv_stem_im2col_f16: bad k_pad=%d", k_pad);
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  const long long work = static_cast<long long>(B) * Ho * Wo * (k_pad / 8);
  stem_im2col_kernel<<<static_cast<unsigned>(ceil_div_i64(work, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nchw, B, C, H, W, kh, kw, stride, pad, Ho, Wo, k_pad, static_cast<__half*>(out));
  MAUV_LAUNCH_CHECK("stem_im2col_kernel");
  return MAUV_OK;
}

// workspace: float2[G * splits * C] when m_tiles > 64 (see mauv_bn_finalize_ws_bytes)
long long mauv_bn_finalize_ws_bytes(int G, int m_tiles, int C) {
  if (m_tiles <= 64) return 0;
  return static_cast<long long>(G) * 64 * C * sizeof(float2);
}

int mauv_bn_finalize(const float* stats_partial, int G, int m_tiles, int C, long long count,
                     const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var, float* scale_shift, float* batch_stats,
                     void* ws, void* stream) {
  MAUV_CHECK_ARG(stats_partial && scale_shift && G >= 1 && m_tiles >= 1 && C >= 1 && count >= 1,
                 "mauv_bn_finalize: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float2* part = reinterpret_cast<const float2*>(stats_partial);
  int tiles = m_tiles;
  if (m_tiles > 64) {
    MAUV_CHECK_ARG(ws != nullptr, "mauv_bn_finalize: workspace required for m_tiles=%d", m_tiles);
    const int splits = 64;
    dim3 grid((C + 127) / 128, splits, G);
    bn_partial_reduce_kernel<<<grid, 128, 0, st>>>(part, G, m_tiles, C, splits, static_cast<float2*>(ws));
    MAUV_LAUNCH_CHECK("bn_partial_reduce_kernel");
    part = static_cast<const float2*>(ws);
    tiles = splits;
  }
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(part, G, tiles, C, count, gamma, beta, eps, momentum,
                                                      running_mean, running_var,
                                                      reinterpret_cast<float2*>(scale_shift),
                                                      reinterpret_cast<float2*>(batch_stats));
  MAUV_LAUNCH_CHECK("bn_finalize_kernel");
  return MAUV_OK;
}

int mauv_bn_act_f16(const void* y, const float* scale_shift, const void* residual, const void* y2,
                    const float* scale_shift2, int relu, int G, long long M, int C, void* out, void* stream) {
  MAUV_CHECK_ARG(y && scale_shift && out, "mauv_bn_act_f16: null pointer");
  MAUV_CHECK_ARG(C % 8 == 0, "mauv_bn_act_f16: C must be a multiple of 8");
  MAUV_CHECK_ARG((y2 == nullptr) == (scale_shift2 == nullptr), "mauv_bn_act_f16: y2 and scale_shift2 go together");
  const long long per = M * C / 8, total = per * G;
  bn_act_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(y), reinterpret_cast<const float2*>(scale_shift),
      static_cast<const uint4*>(residual), static_cast<const uint4*>(y2),
      reinterpret_cast<const float2*>(scale_shift2), relu, per, C, total, static_cast<uint4*>(out));
  MAUV_LAUNCH_CHECK("bn_act_kernel");
  return MAUV_OK;
}

int mauv_bn_relu_maxpool_f16(const void* y, const float* scale_shift, int G, int imgs_per_sample, int H,
                             int W, int C, void* out, void* stream) {
  MAUV_CHECK_ARG(y && scale_shift && out && C % 8 == 0, "mauv_bn_relu_maxpool_f16: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(G) * imgs_per_sample * Ho * Wo * (C / 8);
  bn_relu_maxpool_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(y), reinterpret_cast<const float2*>(scale_shift), imgs_per_sample, H, W, C,
      Ho, Wo, total, static_cast<uint4*>(out));
  MAUV_LAUNCH_CHECK("bn_relu_maxpool_kernel");
  return MAUV_OK;
}

int mauv_avgpool_f16(const void* x, long long N, int HW, int C, float* out, void* stream) {
  MAUV_CHECK_ARG(x && out && N >= 1 && HW >= 1 && C >= 1, "mauv_avgpool_f16: bad argument");
  const long long total = N * C;
  avgpool_kernel<<<static_cast<unsigned>(ceil_div_i64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), HW, C, total, out);
  MAUV_LAUNCH_CHECK("avgpool_kernel");
  return MAUV_OK;
}

int mauv_nchw_f32_to_nhwc_f16(const float* x, long long N, int C, int HW, int c_pad, void* out, void* stream) {
  MAUV_CHECK_ARG(x && out && c_pad >= C, "mauv_nchw_f32_to_nhwc_f16: bad argument");
  const long long total = N * HW * c_pad;
  nchw_to_nhwc_f16_kernel<<<static_cast<unsigned>(ceil_div_i64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, C, HW, c_pad, total, static_cast<__half*>(out));
  MAUV_LAUNCH_CHECK("nchw_to_nhwc_f16_kernel");
  return MAUV_OK;
}

int mauv_nhwc_f16_to_nchw_f32(const void* x, long long N, int C, int HW, float* out, void* stream) {
  MAUV_CHECK_ARG(x && out, "mauv_nhwc_f16_to_nchw_f32: bad argument");
  const long long total = N * HW * C;
  nhwc_to_nchw_f32_kernel<<<static_cast<unsigned>(ceil_div_i64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), C, HW, total, out);
  MAUV_LAUNCH_CHECK("nhwc_to_nchw_f32_kernel");
  return MAUV_OK;
}

}  // extern "C"
