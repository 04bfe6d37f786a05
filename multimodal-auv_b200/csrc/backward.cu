// K6/K7: backward of the sampled conv / linear layers = the autograd of bayesian-torch's forward
// (weight = mu + log1p(exp(rho))*eps; out = F.conv2d / F.linear) that loss.backward() runs in the reference
// (train/multimodal.py:138, train/unimodal.py:145):
//     dX   = conv_transpose(dY, W_s)                       -> tcgen05 kernel over the flipped/transposed sample
//     dW_s = X^T * dY                                       -> tcgen05 kernel, reduction over pixels (split-K as batch)
//     dmu += dW_s ;  drho += dW_s * eps_s * sigmoid(rho)    -> wgrad_finalize (eps replayed from Philox / injected)
// The contractions reuse gemm_tc.cu; this file holds the data-movement helpers and the fp32 head backward.
#include "common.cuh"

namespace {

// dY [N][Ho][Wo][C] fp16 -> zero-stuffed [N][Hd][Wd][C] with dY at (p*s, q*s): stride-s dgrad as a stride-1 conv.
__global__ void __launch_bounds__(256)
dilate_kernel(const uint4* __restrict__ x, int Ho, int Wo, int cvec, int Hd, int Wd, int s, long long total,
              uint4* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = static_cast<int>(i % cvec);
  long long t = i / cvec;
  const int w = static_cast<int>(t % Wd); t /= Wd;
  const int h = static_cast<int>(t % Hd);
  const long long n = t / Hd;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (h % s == 0 && w % s == 0 && h / s < Ho && w / s < Wo) v = x[((n * Ho + h / s) * Wo + w / s) * cvec + cv];
  out[i] = v;
}

// src [M][C] fp16 (row-major) -> dst [splits][C][Mc] (pixels contiguous): the K-major operand of the wgrad GEMM.
__global__ void __launch_bounds__(256)
transpose_chunks_kernel(const __half* __restrict__ src, long long M, int C, int Mc, float scale, __half* __restrict__ dst) {
  __shared__ __half tile[32][33];
  const long long m0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const long long m = m0 + j;
    const int c = c0 + tx;
    tile[j][tx] = (m < M && c < C) ? src[m * C + c] : __float2half(0.f);
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const long long m = m0 + tx;
    if (c < C && m < M) {
      const long long sp = m / Mc, ml = m - sp * Mc;
      dst[(sp * C + c) * Mc + ml] = __float2half_rn(__half2float(tile[tx][j]) * scale);
    }
  }
}

// x NHWC fp16 -> transposed im2col dst [splits][Kp][Mc], K order (r, s, c); rows K..Kp-1 are zero.
__global__ void __launch_bounds__(256)
im2col_t_kernel(const __half* __restrict__ x, int H, int W, int Cin, int kh, int kw, int stride, int pad, int Ho,
                int Wo, int K, int Kp, long long M, int Mc, __half* __restrict__ dst) {
  const long long m = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (m >= M) return;
  __half v = __float2half(0.f);
  if (k < K) {
    const int c = k % Cin;
    const int rs = k / Cin;
    const int s = rs % kw, r = rs / kw;
    const int q = static_cast<int>(m % Wo);
    const int p = static_cast<int>((m / Wo) % Ho);
    const long long n = m / (static_cast<long long>(Wo) * Ho);
    const int h = p * stride - pad + r, w = q * stride - pad + s;
    if (h >= 0 && h < H && w >= 0 && w < W) v = x[((n * H + h) * W + w) * Cin + c];
  }
  const long long sp = m / Mc, ml = m - sp * Mc;
  dst[(sp * Kp + k) * Mc + ml] = v;
}

// Tiled version for Cin % 8 == 0 (every layer but the stems): one block moves a [64 pixels] x [64 channels] tile of one
// filter tap through shared memory, so that both the NHWC read (channels contiguous) and the transposed write (pixels
// contiguous) are 16-byte vector accesses. kh = kw = 1, stride 1 is the plain [M][C] -> [splits][C][Mc] transpose.
__global__ void __launch_bounds__(256)
im2col_t_tiled_kernel(const __half* __restrict__ x, int H, int W, int Cin, int kw, int stride, int pad, int Ho, int Wo,
                      int Kp, long long M, int Mc, float scale, __half* __restrict__ dst) {
  __shared__ __half tile[64][66];
  const long long m0 = static_cast<long long>(blockIdx.x) * 64;
  const int cblocks = (Cin + 63) >> 6;
  const int rs = blockIdx.y / cblocks, c0 = (blockIdx.y - rs * cblocks) << 6;
  const int r = rs / kw, s = rs - r * kw;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int row = (threadIdx.x >> 3) + it * 32, cv = threadIdx.x & 7;
    const long long m = m0 + row;
    const int c = c0 + cv * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (m < M && c < Cin) {
      const int q = static_cast<int>(m % Wo);
      const int p = static_cast<int>((m / Wo) % Ho);
      const long long n = m / (static_cast<long long>(Wo) * Ho);
      const int h = p * stride - pad + r, w = q * stride - pad + s;
      if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(reinterpret_cast<const uint4*>(x + ((n * H + h) * W + w) * Cin + c));
    }
    uint32_t* t32 = reinterpret_cast<uint32_t*>(&tile[row][cv * 8]);
    t32[0] = v.x; t32[1] = v.y; t32[2] = v.z; t32[3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int pg = threadIdx.x & 7, cl = (threadIdx.x >> 3) + it * 32;
    const long long m = m0 + pg * 8;
    const int c = c0 + cl;
    if (m < M && c < Cin) {
      __align__(16) __half o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __half hv = tile[pg * 8 + j][cl];
        o[j] = scale == 1.f ? hv : __float2half_rn(__half2float(hv) * scale);
      }
      const long long sp = m / Mc, ml = m - sp * Mc;
      *reinterpret_cast<uint4*>(dst + (sp * Kp + static_cast<long long>(rs) * Cin + c) * Mc + ml) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

// dW partials fp16 [splits][cout][Kp] ((r,s,c) K order, scaled by `scale`) -> grad_mu += dW, grad_rho += dW*eps*sigmoid(rho)
// in the PyTorch layout [cout][cin][kh][kw]. eps: injected [n] or Philox(seed, layer, sample, e).
__global__ void __launch_bounds__(256)
wgrad_finalize_kernel(const __half* __restrict__ dw, int splits, int cout, int cin, int kh, int kw, int Kp,
                      float inv_scale, const float* __restrict__ rho, const float* __restrict__ eps, uint64_t seed,
                      uint32_t layer_id, uint32_t sample_id, float* __restrict__ grad_mu, float* __restrict__ grad_rho) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int khw = kh * kw;
  const long long per_out = static_cast<long long>(cin) * khw;
  const long long n = per_out * cout;
  if (e >= n) return;
  const long long co = e / per_out;
  const int rem = static_cast<int>(e - co * per_out);
  const int c = rem / khw, rs = rem - c * khw;
  const long long k = static_cast<long long>(rs) * cin + c;
  float g = 0.f;
  for (int sp = 0; sp < splits; ++sp) g += __half2float(dw[(static_cast<long long>(sp) * cout + co) * Kp + k]);
  g *= inv_scale;
  const float z = eps ? eps[e] : philox_normal(seed, layer_id, sample_id, static_cast<uint64_t>(e));
  const float ex = expf(rho[e]);
  const float sgm = isinf(ex) ? 1.f : ex / (1.f + ex);
  grad_mu[e] += g;
  grad_rho[e] += g * z * sgm;
}

// ---------------------------------------------------------------- head (fp32 SIMT, sampling fused)
constexpr int LT = 16;
struct LinBwd {
  const float* x; const float* gy;           // [B][in], [B][out]
  const float* mu_w; const float* rho_w; const float* eps_w; const float* rho_b; const float* eps_b;
  uint64_t seed; uint32_t layer_id, sample_id;
  int B, in, out;
  float* gx;                                   // [B][in] or null
  float* gmu_w; float* grho_w; float* gmu_b; float* grho_b;   // accumulate (+=)
};

// gx[b][i] = sum_o gy[b][o] * w[o][i]
__global__ void __launch_bounds__(256)
linear_bwd_data_kernel(const LinBwd p) {
  __shared__ float gs[LT][LT + 1], ws[LT][LT + 1];
  const int tx = threadIdx.x % LT, ty = threadIdx.x / LT;
  const int i = blockIdx.x * LT + tx, b = blockIdx.y * LT + ty;
  float acc = 0.f;
  for (int o0 = 0; o0 < p.out; o0 += LT) {
    const int ob = o0 + tx;
    gs[ty][tx] = (b < p.B && ob < p.out) ? p.gy[static_cast<long long>(b) * p.out + ob] : 0.f;
    const int ow = o0 + ty;
    float w = 0.f;
    if (ow < p.out && i < p.in) {
      const long long e = static_cast<long long>(ow) * p.in + i;
      const float z = p.eps_w ? p.eps_w[e] : philox_normal(p.seed, p.layer_id, p.sample_id, static_cast<uint64_t>(e));
      w = fmaf(softplus_ref(p.rho_w[e]), z, p.mu_w[e]);
    }
    ws[ty][tx] = w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LT; ++k) acc = fmaf(gs[ty][k], ws[k][tx], acc);
    __syncthreads();
  }
  if (b < p.B && i < p.in) p.gx[static_cast<long long>(b) * p.in + i] = acc;
}

// dW[o][i] = sum_b gy[b][o] * x[b][i] ; grad_mu += dW ; grad_rho += dW * eps * sigmoid(rho) ; bias likewise
__global__ void __launch_bounds__(256)
linear_bwd_weight_kernel(const LinBwd p) {
  __shared__ float gs[LT][LT + 1], xs[LT][LT + 1];
  const int tx = threadIdx.x % LT, ty = threadIdx.x / LT;
  const int i = blockIdx.x * LT + tx, o = blockIdx.y * LT + ty;
  float acc = 0.f, accb = 0.f;
  for (int b0 = 0; b0 < p.B; b0 += LT) {
    const int bb = b0 + ty;
    const int oo = blockIdx.y * LT + tx;
    gs[ty][tx] = (bb < p.B && oo < p.out) ? p.gy[static_cast<long long>(bb) * p.out + oo] : 0.f;   // [b][o]
    xs[ty][tx] = (bb < p.B && i < p.in) ? p.x[static_cast<long long>(bb) * p.in + i] : 0.f;          // [b][i]
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LT; ++k) {
      acc = fmaf(gs[k][ty], xs[k][tx], acc);
      accb += gs[k][ty];
    }
    __syncthreads();
  }
  if (o < p.out && i < p.in) {
    const long long e = static_cast<long long>(o) * p.in + i;
    const float z = p.eps_w ? p.eps_w[e] : philox_normal(p.seed, p.layer_id, p.sample_id, static_cast<uint64_t>(e));
    const float ex = expf(p.rho_w[e]);
    p.gmu_w[e] += acc;
    p.grho_w[e] += acc * z * (isinf(ex) ? 1.f : ex / (1.f + ex));
  }
  if (p.gmu_b && blockIdx.x == 0 && tx == 0 && o < p.out) {
    const float z = p.eps_b ? p.eps_b[o]
                            : philox_normal(p.seed, p.layer_id | 0x80000000u, p.sample_id, static_cast<uint64_t>(o));
    const float ex = expf(p.rho_b[o]);
    p.gmu_b[o] += accb;
    p.grho_b[o] += accb * z * (isinf(ex) ? 1.f : ex / (1.f + ex));
  }
}

}  // namespace

extern "C" {

int mauv_dilate_f16(const void* x, long long N, int Ho, int Wo, int C, int Hd, int Wd, int stride, void* out, void* stream) {
  MAUV_CHECK_ARG(x && out && C % 8 == 0 && stride >= 1, "mauv_dilate_f16: bad argument");
  MAUV_CHECK_ARG((Ho - 1) * stride < Hd && (Wo - 1) * stride < Wd, "mauv_dilate_f16: target too small");
  const long long total = N * Hd * Wd * (C / 8);
  dilate_kernel<<<static_cast<unsigned>(ceil_div_i64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), Ho, Wo, C / 8, Hd, Wd, stride, total, static_cast<uint4*>(out));
  MAUV_LAUNCH_CHECK("dilate_kernel");
  return MAUV_OK;
}

int mauv_transpose_chunks_f16(const void* src, long long M, int C, int splits, float scale, void* dst, void* stream) {
  MAUV_CHECK_ARG(src && dst && M >= 1 && C >= 1 && splits >= 1 && M % splits == 0, "mauv_transpose_chunks_f16: bad argument");
  const long long Mc = M / splits;
  if (C % 8 == 0 && Mc % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    dim3 grid(static_cast<unsigned>(ceil_div_i64(M, 64)), (C + 63) / 64);
    MAUV_CHECK_ARG(grid.y <= 65535, "mauv_transpose_chunks_f16: too many channels");
    im2col_t_tiled_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(src), 1, 1, C, 1, 1, 0, 1, 1, C, M, static_cast<int>(Mc), scale, static_cast<__half*>(dst));
    MAUV_LAUNCH_CHECK("im2col_t_tiled_kernel");
    return MAUV_OK;
  }
  dim3 grid(static_cast<unsigned>(ceil_div_i64(M, 32)), (C + 31) / 32);
  transpose_chunks_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(src), M, C, static_cast<int>(Mc), scale, static_cast<__half*>(dst));
  MAUV_LAUNCH_CHECK("transpose_chunks_kernel");
  return MAUV_OK;
}

int mauv_im2col_t_f16(const void* x, long long N, int H, int W, int Cin, int kh, int kw, int stride, int pad,
                      int k_pad, int splits, void* dst, void* stream) {
  MAUV_CHECK_ARG(x && dst, "mauv_im2col_t_f16: null pointer");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  const long long M = N * Ho * Wo;
  const int K = kh * kw * Cin;
  MAUV_CHECK_ARG(k_pad >= K && splits >= 1 && M % splits == 0 && k_pad <= 65535, "mauv_im2col_t_f16: bad argument");
  const long long Mc = M / splits;
  if (Cin % 8 == 0 && k_pad == K && Mc % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    dim3 tg(static_cast<unsigned>(ceil_div_i64(M, 64)), static_cast<unsigned>(kh * kw * ((Cin + 63) / 64)));
    MAUV_CHECK_ARG(tg.y <= 65535, "mauv_im2col_t_f16: too many (tap, channel-block) pairs");
    im2col_t_tiled_kernel<<<tg, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(x), H, W, Cin, kw, stride, pad, Ho, Wo, k_pad, M, static_cast<int>(Mc), 1.f,
        static_cast<__half*>(dst));
    MAUV_LAUNCH_CHECK("im2col_t_tiled_kernel");
    return MAUV_OK;
  }
  dim3 grid(static_cast<unsigned>(ceil_div_i64(M, 256)), k_pad);
  im2col_t_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), H, W, Cin, kh, kw, stride, pad, Ho, Wo, K, k_pad, M, static_cast<int>(M / splits),
      static_cast<__half*>(dst));
  MAUV_LAUNCH_CHECK("im2col_t_kernel");
  return MAUV_OK;
}

int mauv_wgrad_finalize(const void* dw_partial, int splits, int cout, int cin, int kh, int kw, int k_pad, float inv_scale,
                        const float* rho, const float* eps, uint64_t seed, uint32_t layer_id, uint32_t sample_id,
                        float* grad_mu, float* grad_rho, void* stream) {
  MAUV_CHECK_ARG(dw_partial && rho && grad_mu && grad_rho && splits >= 1, "mauv_wgrad_finalize: bad argument");
  const long long n = static_cast<long long>(cout) * cin * kh * kw;
  wgrad_finalize_kernel<<<static_cast<unsigned>(ceil_div_i64(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(dw_partial), splits, cout, cin, kh, kw, k_pad, inv_scale, rho, eps, seed, layer_id,
      sample_id, grad_mu, grad_rho);
  MAUV_LAUNCH_CHECK("wgrad_finalize_kernel");
  return MAUV_OK;
}

int mauv_sampled_linear_bwd_f32(const float* x, const float* gy, const float* mu_w, const float* rho_w,
                                const float* eps_w, const float* rho_b, const float* eps_b, uint64_t seed,
                                uint32_t layer_id, uint32_t sample_id, int B, int in_features, int out_features,
                                float* gx, float* grad_mu_w, float* grad_rho_w, float* grad_mu_b, float* grad_rho_b,
                                void* stream) {
  MAUV_CHECK_ARG(x && gy && mu_w && rho_w && grad_mu_w && grad_rho_w, "mauv_sampled_linear_bwd_f32: null pointer");
  MAUV_CHECK_ARG((grad_mu_b == nullptr) == (grad_rho_b == nullptr) && (grad_mu_b == nullptr || rho_b != nullptr),
                 "mauv_sampled_linear_bwd_f32: bias gradient pointers go together with rho_b");
  LinBwd p;
  p.x = x; p.gy = gy; p.mu_w = mu_w; p.rho_w = rho_w; p.eps_w = eps_w; p.rho_b = rho_b; p.eps_b = eps_b;
  p.seed = seed; p.layer_id = layer_id; p.sample_id = sample_id; p.B = B; p.in = in_features; p.out = out_features;
  p.gx = gx; p.gmu_w = grad_mu_w; p.grho_w = grad_rho_w; p.gmu_b = grad_mu_b; p.grho_b = grad_rho_b;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gx) {
    dim3 g1((in_features + LT - 1) / LT, (B + LT - 1) / LT);
    linear_bwd_data_kernel<<<g1, 256, 0, st>>>(p);
    MAUV_LAUNCH_CHECK("linear_bwd_data_kernel");
  }
  dim3 g2((in_features + LT - 1) / LT, (out_features + LT - 1) / LT);
  linear_bwd_weight_kernel<<<g2, 256, 0, st>>>(p);
  MAUV_LAUNCH_CHECK("linear_bwd_weight_kernel");
  return MAUV_OK;
}

}  // extern "C"
