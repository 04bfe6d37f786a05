// K1 (operand staging): per-MC-sample weight sampling  w = mu + log1p(exp(rho)) * eps.
//
// Restates bayesian-torch 0.5.0 Conv2dReparameterization/LinearReparameterization.forward
// (sigma = log1p(exp(rho)); eps ~ N(0,1); weight = mu + sigma*eps), installed on every
// conv/linear of the reference by models/model_utils.py:26-35.
//
// One launch produces the fp16 operand tiles for G Monte-Carlo samples of ONE layer,
// already in the K order the implicit-GEMM kernel consumes ((r, s, c) = NHWC im2col
// order). Only this one layer's G copies ever exist; the buffer is a few MB, is
// consumed immediately by gemm_tc.cu through TMA and is reused by the next layer, so
// the S x 73M sampled network weights are never materialised.
// eps is either injected (validation: the oracle's captured eps, PyTorch layout) or
// generated in-kernel with Philox4x32-10 keyed by (seed; layer, sample, element).
#include "common.cuh"

namespace {

struct SampleParams {
  const float* mu;
  const float* rho;
  const float* eps;   // [G][n] or nullptr
  uint64_t seed;
  uint32_t layer_id, sample0;
  int cout, cin, kh, kw, k_pad;
  long long n;        // cout*cin*kh*kw
  __half* w;          // [G][cout][k_pad]   (dgrad: [G][cin][kh*kw*cout], taps flipped)
  int dgrad;          // 1: write the transposed + spatially flipped layout the data-gradient conv consumes
  const float2* row_scale;   // optional [G][cout] (scale, shift): w *= scale (a BatchNorm scale folded into the weights)
  long long g_stride; // elements between consecutive samples in w (0 = dense default)
  const unsigned int* sample_base;   // optional device word added to the sample ids at run time (mauv_set_sample_base)
};

__device__ __forceinline__ void normals4(uint64_t seed, uint32_t layer, uint32_t sample,
                                         uint64_t quad, float (&z)[4]) {
  philox_normals4(seed, layer, sample, quad, z);      // the one eps stream every kernel shares (common.cuh)
}

// Each thread handles 4 consecutive elements of the PyTorch-layout parameter tensor ([cout][cin][kh][kw]) for ALL
// G samples of the launch: mu/rho are read once and sigma = log1p(exp(rho)) is computed once (it does not depend on
// the sample); per sample only one Philox4x32-10 call + two Box-Muller pairs + 4 FMAs + the fp16 store remain.
__global__ void __launch_bounds__(256)
sample_weights_kernel(const SampleParams p, int G) {
  const long long quad = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long e0 = quad * 4;
  if (e0 >= p.n) return;
  float mu[4], sg[4];
  const bool full = (e0 + 3 < p.n);
  if (full) {
    const float4 m4 = *reinterpret_cast<const float4*>(p.mu + e0);
    const float4 r4 = *reinterpret_cast<const float4*>(p.rho + e0);
    mu[0] = m4.x; mu[1] = m4.y; mu[2] = m4.z; mu[3] = m4.w;
    sg[0] = softplus_ref(r4.x); sg[1] = softplus_ref(r4.y); sg[2] = softplus_ref(r4.z); sg[3] = softplus_ref(r4.w);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mu[i] = (e0 + i < p.n) ? p.mu[e0 + i] : 0.f;
      sg[i] = (e0 + i < p.n) ? softplus_ref(p.rho[e0 + i]) : 0.f;
    }
  }
  const int khw = p.kh * p.kw;
  const long long per_out = static_cast<long long>(p.cin) * khw;
  const bool direct = (!p.dgrad && khw == 1 && full && (p.cin % 4 == 0));   // K order == PyTorch order
  long long off[4];
  if (direct) {
    const long long co = e0 / per_out;
    off[0] = co * p.k_pad + (e0 - co * per_out);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long e = e0 + i;
      const long long co = e / per_out;
      const int rem = static_cast<int>(e - co * per_out);
      const int c = rem / khw;
      const int rs = rem - c * khw;  // r*kw + s
      off[i] = p.dgrad ? static_cast<long long>(c) * p.k_pad + static_cast<long long>(khw - 1 - rs) * p.cout + co
                       : co * p.k_pad + static_cast<long long>(rs) * p.cin + c;
    }
  }
  const uint32_t sample0 = p.sample0 + (p.sample_base ? *p.sample_base : 0u);
  const long long w_stride = p.g_stride ? p.g_stride
                             : (p.dgrad ? static_cast<long long>(p.cin) * p.k_pad : static_cast<long long>(p.cout) * p.k_pad);
  // two samples in flight per thread: a Philox block is a 30-deep dependent chain, and at 4 resident blocks per SM the
  // scheduler alone cannot hide it (ncu: issue slots 40 % busy, every pipe below 30 %)
#pragma unroll 2
  for (int g = 0; g < G; ++g) {
    float z[4];
    if (p.eps) {
      const float* ep = p.eps + static_cast<long long>(g) * p.n + e0;
#pragma unroll
      for (int i = 0; i < 4; ++i) z[i] = (e0 + i < p.n) ? ep[i] : 0.f;
    } else {
      normals4(p.seed, p.layer_id, sample0 + g, static_cast<uint64_t>(quad), z);
    }
    float w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = fmaf(sg[i], z[i], mu[i]);
    if (p.row_scale) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long co = (e0 + i) / per_out;
        if (co < p.cout) w[i] *= p.row_scale[static_cast<long long>(g) * p.cout + co].x;
      }
    }
    __half* wg = p.w + static_cast<long long>(g) * w_stride;
    if (direct) {
      __half2 h0 = __floats2half2_rn(w[0], w[1]);
      __half2 h1 = __floats2half2_rn(w[2], w[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(wg + off[0]) = pk;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (e0 + i < p.n) wg[off[i]] = __float2half_rn(w[i]);
    }
  }
}

// fp32 sampled vector (biases): out[g][i] = mu[i] + softplus(rho[i]) * eps
__global__ void __launch_bounds__(256)
sample_vector_f32_kernel(const float* __restrict__ mu, const float* __restrict__ rho,
                         const float* __restrict__ eps, uint64_t seed, uint32_t layer_id,
                         uint32_t sample0, int n, float* __restrict__ out) {
  const int g = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float z;
  if (eps) z = eps[static_cast<long long>(g) * n + i];
  else z = philox_normal(seed, layer_id, sample0 + g, static_cast<uint64_t>(i));
  out[static_cast<long long>(g) * n + i] = fmaf(softplus_ref(rho[i]), z, mu[i]);
}

// The eps stream itself (tests / oracle cross-check of the Philox spec).
__global__ void __launch_bounds__(256)
philox_normal_kernel(uint64_t seed, uint32_t layer_id, uint32_t sample_id, long long n,
                     float* __restrict__ out) {
  const long long quad = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (quad * 4 >= n) return;
  float z[4];
  normals4(seed, layer_id, sample_id, static_cast<uint64_t>(quad), z);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (quad * 4 + i < n) out[quad * 4 + i] = z[i];
}

}  // namespace

extern "C" {

int mauv_sample_weights_f16(const float* mu, const float* rho, const float* eps, uint64_t seed,
                            uint32_t layer_id, uint32_t sample0, int G, int cout, int cin, int kh,
                            int kw, int k_pad, void* w_out, void* stream) {
  MAUV_CHECK_ARG(mu && rho && w_out, "mauv_sample_weights_f16: null pointer");
  MAUV_CHECK_ARG(G >= 1 && cout >= 1 && cin >= 1 && kh >= 1 && kw >= 1, "mauv_sample_weights_f16: bad shape");
  const int K = cin * kh * kw;
  MAUV_CHECK_ARG(k_pad >= K && k_pad % 8 == 0, "mauv_sample_weights_f16: k_pad=%d must be >= K=%d and a multiple of 8", k_pad, K);
  MAUV_CHECK_ARG((reinterpret_cast<uintptr_t>(mu) & 15) == 0 && (reinterpret_cast<uintptr_t>(rho) & 15) == 0,
                 "mauv_sample_weights_f16: mu/rho must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k_pad != K)  // zero the K padding (stem convs: K = 147 / 49)
    MAUV_CUDA(cudaMemsetAsync(w_out, 0, static_cast<size_t>(G) * cout * k_pad * sizeof(__half), st));
  SampleParams p;
  p.mu = mu; p.rho = rho; p.eps = eps; p.seed = seed; p.layer_id = layer_id; p.sample0 = sample0;
  p.cout = cout; p.cin = cin; p.kh = kh; p.kw = kw; p.k_pad = k_pad;
  p.n = static_cast<long long>(cout) * K;
  p.w = static_cast<__half*>(w_out);
  p.dgrad = 0;
  p.row_scale = nullptr;
  p.g_stride = 0;
  p.sample_base = mauv_sample_base();
  const long long quads = ceil_div_i64(p.n, 4);
  sample_weights_kernel<<<static_cast<unsigned>(ceil_div_i64(quads, 256)), 256, 0, st>>>(p, G);
  MAUV_LAUNCH_CHECK("sample_weights_kernel");
  return MAUV_OK;
}

// 1x1 conv / linear weights sampled into a column block of a wider (K-concatenated) operand, each output row multiplied by
// a per-(sample, row) BatchNorm scale: w_out[g][co][col0 + c] = scale[g][co] * (mu + log1p(exp(rho)) * eps)[co][c].
// row_pitch = total columns of the concatenated operand; scale_shift [G][cout][2] as written by mauv_bn_finalize.
int mauv_sample_weights_scaled_f16(const float* mu, const float* rho, const float* eps, uint64_t seed, uint32_t layer_id,
                                   uint32_t sample0, int G, int cout, int cin, const float* scale_shift, int row_pitch,
                                   int col0, void* w_out, void* stream) {
  MAUV_CHECK_ARG(mu && rho && w_out && scale_shift, "mauv_sample_weights_scaled_f16: null pointer");
  MAUV_CHECK_ARG(G >= 1 && cout >= 1 && cin >= 1 && cin % 4 == 0 && row_pitch % 8 == 0 && col0 % 8 == 0 && col0 + cin <= row_pitch,
                 "mauv_sample_weights_scaled_f16: bad layout (cin=%d row_pitch=%d col0=%d)", cin, row_pitch, col0);
  SampleParams p;
  p.mu = mu; p.rho = rho; p.eps = eps; p.seed = seed; p.layer_id = layer_id; p.sample0 = sample0;
  p.cout = cout; p.cin = cin; p.kh = 1; p.kw = 1; p.k_pad = row_pitch;
  p.n = static_cast<long long>(cout) * cin;
  p.w = static_cast<__half*>(w_out) + col0;
  p.dgrad = 0;
  p.row_scale = reinterpret_cast<const float2*>(scale_shift);
  p.g_stride = static_cast<long long>(cout) * row_pitch;
  p.sample_base = mauv_sample_base();
  const long long quads = ceil_div_i64(p.n, 4);
  sample_weights_kernel<<<static_cast<unsigned>(ceil_div_i64(quads, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, G);
  MAUV_LAUNCH_CHECK("sample_weights_kernel(scaled)");
  return MAUV_OK;
}

// Same sample (same eps stream / same injected eps) in the layout of the data-gradient convolution:
// w_out[g][ci][(kh-1-r, kw-1-s, co)] = w[g][co][ci][r][s]   -> operand [G][N=cin][K=kh*kw*cout] of
// dX = conv(dY (zero-stuffed for stride 2), w_out, stride 1, pad k-1-p).
int mauv_sample_weights_dgrad_f16(const float* mu, const float* rho, const float* eps, uint64_t seed,
                                  uint32_t layer_id, uint32_t sample0, int G, int cout, int cin, int kh,
                                  int kw, void* w_out, void* stream) {
  MAUV_CHECK_ARG(mu && rho && w_out, "mauv_sample_weights_dgrad_f16: null pointer");
  MAUV_CHECK_ARG(G >= 1 && cout % 8 == 0 && cin >= 1 && kh >= 1 && kw >= 1, "mauv_sample_weights_dgrad_f16: bad shape");
  SampleParams p;
  p.mu = mu; p.rho = rho; p.eps = eps; p.seed = seed; p.layer_id = layer_id; p.sample0 = sample0;
  p.cout = cout; p.cin = cin; p.kh = kh; p.kw = kw; p.k_pad = kh * kw * cout;
  p.n = static_cast<long long>(cout) * cin * kh * kw;
  p.w = static_cast<__half*>(w_out);
  p.dgrad = 1;
  p.row_scale = nullptr;
  p.g_stride = 0;
  p.sample_base = mauv_sample_base();
  const long long quads = ceil_div_i64(p.n, 4);
  sample_weights_kernel<<<static_cast<unsigned>(ceil_div_i64(quads, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, G);
  MAUV_LAUNCH_CHECK("sample_weights_kernel(dgrad)");
  return MAUV_OK;
}

int mauv_sample_vector_f32(const float* mu, const float* rho, const float* eps, uint64_t seed,
                           uint32_t layer_id, uint32_t sample0, int G, int n, float* out, void* stream) {
  MAUV_CHECK_ARG(mu && rho && out && G >= 1 && n >= 1, "mauv_sample_vector_f32: bad argument");
  dim3 grid((n + 255) / 256, G);
  sample_vector_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mu, rho, eps, seed, layer_id, sample0, n, out);
  MAUV_LAUNCH_CHECK("sample_vector_f32_kernel");
  return MAUV_OK;
}

int mauv_philox_normal_f32(uint64_t seed, uint32_t layer_id, uint32_t sample_id, long long n, float* out, void* stream) {
  MAUV_CHECK_ARG(out && n >= 1, "mauv_philox_normal_f32: bad argument");
  const long long quads = ceil_div_i64(n, 4);
  philox_normal_kernel<<<static_cast<unsigned>(ceil_div_i64(quads, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, layer_id, sample_id, n, out);
  MAUV_LAUNCH_CHECK("philox_normal_kernel");
  return MAUV_OK;
}

}  // extern "C"
