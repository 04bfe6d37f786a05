// K2: the fusion head. 15 Bayesian Linear layers per MC pass (AdditiveAttention x3:
// models/base_models.py:35-52, then fc/fc1/fc2: models/base_models.py:60-65,86-89) with
// M = batch rows only, so here the weight sampling w = mu + log1p(exp(rho))*eps really is
// fused into operand staging: each CTA samples its [16 x 32] weight tile straight into
// shared memory (Philox or injected eps) and contracts it in fp32; no sampled weight
// ever reaches HBM. (bayesian-torch linear_variational.py forward; weight AND bias sampled.)
#include "common.cuh"

namespace {

constexpr int TO = 16;   // outputs per CTA
constexpr int TB = 32;   // batch rows per CTA: small tiles -> several CTAs per SM hide the global-load latency of the k loop
constexpr int TK = 32;   // k chunk

struct LinearParams {
  const float* x; long long x_gs; int ldx;      // [G][B][in], sample stride, row stride
  const float* mu_w; const float* rho_w; const float* eps_w;   // [out][in], eps [G][out][in] or null
  const float* mu_b; const float* rho_b; const float* eps_b;   // [out] or null (no bias)
  uint64_t seed; uint32_t layer_id, sample0;
  int G, B, in, out;
  float* y; long long y_gs; int ldy;            // [G][B][out]
};

__global__ void __launch_bounds__(256)
sampled_linear_kernel(const LinearParams p) {
  __shared__ float xs[TB][TK + 1];
  __shared__ float ws[TO][TK + 1];
  const int g = blockIdx.z;
  const int o0 = blockIdx.x * TO;
  const int b0 = blockIdx.y * TB;
  const int tx = threadIdx.x % TO;
  const int ty = threadIdx.x / TO;  // 0..15
  const float* xg = p.x + static_cast<long long>(g) * p.x_gs;
  float acc[TB / 16];
#pragma unroll
  for (int j = 0; j < TB / 16; ++j) acc[j] = 0.f;
  const bool quads = (p.in % 4 == 0);     // 4 consecutive k share one Philox4x32 block

  for (int k0 = 0; k0 < p.in; k0 += TK) {
    // activations: 64 x 32 tile, coalesced along k
    for (int i = threadIdx.x; i < TB * TK; i += 256) {
      const int r = i / TK, k = i % TK;
      const int b = b0 + r, kk = k0 + k;
      xs[r][k] = (b < p.B && kk < p.in) ? xg[static_cast<long long>(b) * p.ldx + kk] : 0.f;
    }
    // weights: sample the 16 x 32 tile in place
    if (quads && !p.eps_w) {
      if (threadIdx.x < TO * TK / 4) {
        const int r = threadIdx.x / (TK / 4), k = (threadIdx.x % (TK / 4)) * 4;
        const int o = o0 + r, kk = k0 + k;
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        if (o < p.out && kk < p.in) {            // in % 4 == 0: the whole quad is in range
          const long long e = static_cast<long long>(o) * p.in + kk;
          float z[4];
          philox_normals4(p.seed, p.layer_id, p.sample0 + g, static_cast<uint64_t>(e >> 2), z);
          const float4 m4 = *reinterpret_cast<const float4*>(p.mu_w + e);
          const float4 r4 = *reinterpret_cast<const float4*>(p.rho_w + e);
          w[0] = fmaf(softplus_ref(r4.x), z[0], m4.x);
          w[1] = fmaf(softplus_ref(r4.y), z[1], m4.y);
          w[2] = fmaf(softplus_ref(r4.z), z[2], m4.z);
          w[3] = fmaf(softplus_ref(r4.w), z[3], m4.w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) ws[r][k + j] = w[j];
      }
    } else {
      for (int i = threadIdx.x; i < TO * TK; i += 256) {
        const int r = i / TK, k = i % TK;
        const int o = o0 + r, kk = k0 + k;
        float w = 0.f;
        if (o < p.out && kk < p.in) {
          const long long e = static_cast<long long>(o) * p.in + kk;
          const float z = p.eps_w ? p.eps_w[static_cast<long long>(g) * p.out * p.in + e]
                                  : philox_normal(p.seed, p.layer_id, p.sample0 + g, static_cast<uint64_t>(e));
          w = fmaf(softplus_ref(p.rho_w[e]), z, p.mu_w[e]);
        }
        ws[r][k] = w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float w = ws[tx][k];
#pragma unroll
      for (int j = 0; j < TB / 16; ++j) acc[j] = fmaf(xs[ty + 16 * j][k], w, acc[j]);
    }
    __syncthreads();
  }
  const int o = o0 + tx;
  if (o >= p.out) return;
  float bias = 0.f;
  if (p.mu_b) {
    // bias eps uses layer_id | 0x80000000 so its Philox stream is disjoint from the weight's
    const float z = p.eps_b ? p.eps_b[static_cast<long long>(g) * p.out + o]
                            : philox_normal(p.seed, p.layer_id | 0x80000000u, p.sample0 + g, static_cast<uint64_t>(o));
    bias = fmaf(softplus_ref(p.rho_b[o]), z, p.mu_b[o]);
  }
  float* yg = p.y + static_cast<long long>(g) * p.y_gs;
#pragma unroll
  for (int j = 0; j < TB / 16; ++j) {
    const int b = b0 + ty + 16 * j;
    if (b < p.B) yg[static_cast<long long>(b) * p.ldy + o] = acc[j] + bias;
  }
}

__global__ void __launch_bounds__(256)
tanh_add_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tanhf(a[i] + b[i]);
}

// out[row][j] = v[row][j] * softmax_j(score[row][:])   (AdditiveAttention, base_models.py:48-51)
// one warp per row; out has its own row stride so the three modalities write straight into
// the [G][B][384] concat buffer of MultiModalModel.forward (base_models.py:86).
__global__ void __launch_bounds__(256)
softmax_gate_kernel(const float* __restrict__ score, const float* __restrict__ v, long long rows, int n,
                    float* __restrict__ out, int ld_out) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* s = score + row * n;
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, s[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float z = 0.f;
  for (int j = lane; j < n; j += 32) z += expf(s[j] - m);
  z = warp_sum(z);
  for (int j = lane; j < n; j += 32) out[row * ld_out + j] = v[row * n + j] * (expf(s[j] - m) / z);
}

}  // namespace

extern "C" {

int mauv_sampled_linear_f32(const float* x, long long x_sample_stride, int ldx, const float* mu_w,
                            const float* rho_w, const float* eps_w, const float* mu_b, const float* rho_b,
                            const float* eps_b, uint64_t seed, uint32_t layer_id, uint32_t sample0, int G,
                            int B, int in_features, int out_features, float* y, long long y_sample_stride,
                            int ldy, void* stream) {
  MAUV_CHECK_ARG(x && mu_w && rho_w && y, "mauv_sampled_linear_f32: null pointer");
  MAUV_CHECK_ARG((mu_b == nullptr) == (rho_b == nullptr), "mauv_sampled_linear_f32: mu_b and rho_b go together");
  MAUV_CHECK_ARG(G >= 1 && B >= 1 && in_features >= 1 && out_features >= 1, "mauv_sampled_linear_f32: bad shape");
  LinearParams p;
  p.x = x; p.x_gs = x_sample_stride; p.ldx = ldx;
  p.mu_w = mu_w; p.rho_w = rho_w; p.eps_w = eps_w;
  p.mu_b = mu_b; p.rho_b = rho_b; p.eps_b = eps_b;
  p.seed = seed; p.layer_id = layer_id; p.sample0 = sample0;
  p.G = G; p.B = B; p.in = in_features; p.out = out_features;
  p.y = y; p.y_gs = y_sample_stride; p.ldy = ldy;
  dim3 grid((out_features + TO - 1) / TO, (B + TB - 1) / TB, G);
  sampled_linear_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MAUV_LAUNCH_CHECK("sampled_linear_kernel");
  return MAUV_OK;
}

int mauv_tanh_add_f32(const float* a, const float* b, long long n, float* out, void* stream) {
  MAUV_CHECK_ARG(a && b && out && n >= 1, "mauv_tanh_add_f32: bad argument");
  tanh_add_kernel<<<static_cast<unsigned>(ceil_div_i64(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, out);
  MAUV_LAUNCH_CHECK("tanh_add_kernel");
  return MAUV_OK;
}

int mauv_softmax_gate_f32(const float* score, const float* v, long long rows, int n, float* out, int ld_out, void* stream) {
  MAUV_CHECK_ARG(score && v && out && rows >= 1 && n >= 1 && ld_out >= n, "mauv_softmax_gate_f32: bad argument");
  softmax_gate_kernel<<<static_cast<unsigned>(ceil_div_i64(rows * 32, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      score, v, rows, n, out, ld_out);
  MAUV_LAUNCH_CHECK("softmax_gate_kernel");
  return MAUV_OK;
}

}  // extern "C"
