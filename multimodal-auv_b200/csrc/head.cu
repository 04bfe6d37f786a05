// K2: the fusion head. 15 Bayesian Linear layers per MC pass (AdditiveAttention x3:
// models/base_models.py:35-52, then fc/fc1/fc2: models/base_models.py:60-65,86-89) with
// M = batch rows only, so here the weight sampling w = mu + log1p(exp(rho))*eps really is
// fused into operand staging: each CTA samples its [32 x 32] weight tile straight into
// shared memory (Philox or injected eps) and contracts it in fp32; no sampled weight
// ever reaches HBM. (bayesian-torch linear_variational.py forward; weight AND bias sampled.)
#include "common.cuh"

namespace {

constexpr int TO = 32;        // outputs per CTA
constexpr int TK = 32;        // k chunk
// batch rows per CTA (template parameter TBR): every row tile re-samples its weight tile - 4 exact softplus + one Philox block +
// two Box-Muller pairs per thread and k-step, ~80 % of the kernel's instructions at 32 rows - so large batches take 128-row
// tiles (4x fewer samplings per weight, 8 x 2 outputs per thread); small ones (training: 8 rows) keep 32.
constexpr int PITCH = TK + 4; // 16-byte aligned rows, float4 reads in the inner product

struct LinearParams {
  const float* x; long long x_gs; int ldx;      // [G][B][in], sample stride, row stride
  const float* mu_w; const float* rho_w; const float* eps_w;   // [out][in], eps [G][out][in] or null
  const float* mu_b; const float* rho_b; const float* eps_b;   // [out] or null (no bias)
  uint64_t seed; uint32_t layer_id, sample0;
  int G, B, in, out;
  float* y; long long y_gs; int ldy;            // [G][B][out]
  const unsigned int* sample_base;              // optional device word added to the sample ids (mauv_set_sample_base)
};

// 256 threads = 16 x 16; each thread owns a 2 x 2 block of the 32 x 32 output tile (rows ty, ty+16; outputs tx, tx+16) and
// walks k four at a time with 16-byte shared-memory reads (4 FMAs per LDS.128). Per k-step the CTA stages a 32 x 32
// activation tile (one float4 per thread) and SAMPLES its 32 x 32 weight tile in place (one Philox4x32 block per thread).
template <int TBR>
__global__ void __launch_bounds__(256)
sampled_linear_kernel(const LinearParams p) {
  constexpr int TB = TBR;
  constexpr int RT = TBR / 16;     // rows per thread
  const uint32_t sample0 = p.sample0 + (p.sample_base ? *p.sample_base : 0u);
  __shared__ __align__(16) float xs[TB][PITCH];
  __shared__ __align__(16) float ws[TO][PITCH];
  const int g = blockIdx.z;
  const int o0 = blockIdx.x * TO;
  const int b0 = blockIdx.y * TB;
  const int tx = threadIdx.x & 15;
  const int ty = threadIdx.x >> 4;  // 0..15
  const float* xg = p.x + static_cast<long long>(g) * p.x_gs;
  float acc[RT][2];
#pragma unroll
  for (int jb = 0; jb < RT; ++jb) acc[jb][0] = acc[jb][1] = 0.f;
  const bool quads = (p.in % 4 == 0);     // 4 consecutive k share one Philox4x32 block / one 16-byte load
  const bool x_vec = quads && ((reinterpret_cast<uintptr_t>(xg) & 15) == 0) && (p.ldx % 4 == 0);
  const int sr = threadIdx.x >> 3, sc = (threadIdx.x & 7) * 4;      // staging role: row sr, columns sc .. sc+3

  for (int k0 = 0; k0 < p.in; k0 += TK) {
#pragma unroll
    for (int rb = 0; rb < TB / 32; ++rb) {   // activations: 32 rows per pass
      const int b = b0 + 32 * rb + sr, kk = k0 + sc;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < p.B) {
        const float* src = xg + static_cast<long long>(b) * p.ldx + kk;
        if (x_vec && kk < p.in) v = *reinterpret_cast<const float4*>(src);
        else {
          if (kk < p.in) v.x = src[0];
          if (kk + 1 < p.in) v.y = src[1];
          if (kk + 2 < p.in) v.z = src[2];
          if (kk + 3 < p.in) v.w = src[3];
        }
      }
      *reinterpret_cast<float4*>(&xs[32 * rb + sr][sc]) = v;
    }
    {   // weights: sample the tile in place
      const int o = o0 + sr, kk = k0 + sc;
      float w[4] = {0.f, 0.f, 0.f, 0.f};
      if (o < p.out && kk < p.in) {
        const long long e = static_cast<long long>(o) * p.in + kk;
        if (quads && !p.eps_w) {           // in % 4 == 0: the whole quad is in range and 16-byte aligned
          float z[4];
          philox_normals4(p.seed, p.layer_id, sample0 + g, static_cast<uint64_t>(e >> 2), z);
          const float4 m4 = *reinterpret_cast<const float4*>(p.mu_w + e);
          const float4 r4 = *reinterpret_cast<const float4*>(p.rho_w + e);
          w[0] = fmaf(softplus_ref(r4.x), z[0], m4.x);
          w[1] = fmaf(softplus_ref(r4.y), z[1], m4.y);
          w[2] = fmaf(softplus_ref(r4.z), z[2], m4.z);
          w[3] = fmaf(softplus_ref(r4.w), z[3], m4.w);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (kk + j < p.in) {
              const float z = p.eps_w ? p.eps_w[static_cast<long long>(g) * p.out * p.in + e + j]
                                      : philox_normal(p.seed, p.layer_id, sample0 + g, static_cast<uint64_t>(e + j));
              w[j] = fmaf(softplus_ref(p.rho_w[e + j]), z, p.mu_w[e + j]);
            }
          }
        }
      }
      *reinterpret_cast<float4*>(&ws[sr][sc]) = make_float4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < TK; k4 += 4) {
      const float4 wa = *reinterpret_cast<const float4*>(&ws[tx][k4]);
      const float4 wb = *reinterpret_cast<const float4*>(&ws[tx + 16][k4]);
#pragma unroll
      for (int jb = 0; jb < RT; ++jb) {
        const float4 xa = *reinterpret_cast<const float4*>(&xs[ty + 16 * jb][k4]);
        // k order inside the chunk is fixed (x, y, z, w): deterministic accumulation
        acc[jb][0] = fmaf(xa.w, wa.w, fmaf(xa.z, wa.z, fmaf(xa.y, wa.y, fmaf(xa.x, wa.x, acc[jb][0]))));
        acc[jb][1] = fmaf(xa.w, wb.w, fmaf(xa.z, wb.z, fmaf(xa.y, wb.y, fmaf(xa.x, wb.x, acc[jb][1]))));
      }
    }
    __syncthreads();
  }
  float* yg = p.y + static_cast<long long>(g) * p.y_gs;
#pragma unroll
  for (int jo = 0; jo < 2; ++jo) {
    const int o = o0 + tx + 16 * jo;
    if (o >= p.out) continue;
    float bias = 0.f;
    if (p.mu_b) {
      // bias eps uses layer_id | 0x80000000 so its Philox stream is disjoint from the weight's
      const float z = p.eps_b ? p.eps_b[static_cast<long long>(g) * p.out + o]
                              : philox_normal(p.seed, p.layer_id | 0x80000000u, sample0 + g, static_cast<uint64_t>(o));
      bias = fmaf(softplus_ref(p.rho_b[o]), z, p.mu_b[o]);
    }
#pragma unroll
    for (int jb = 0; jb < RT; ++jb) {
      const int b = b0 + ty + 16 * jb;
      if (b < p.B) yg[static_cast<long long>(b) * p.ldy + o] = acc[jb][jo] + bias;
    }
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
  p.sample_base = mauv_sample_base();
  p.G = G; p.B = B; p.in = in_features; p.out = out_features;
  p.y = y; p.y_gs = y_sample_stride; p.ldy = ldy;
  // 128-row tiles only when they still fill the GPU (at 8 GPUs a rank walks 3-4 samples: 32 CTAs for a 128-wide projection)
  const long long ctas128 = static_cast<long long>((out_features + TO - 1) / TO) * ((B + 127) / 128) * G;
  if (B >= 128 && ctas128 >= mauv_num_sms()) {
    dim3 grid((out_features + TO - 1) / TO, (B + 127) / 128, G);
    sampled_linear_kernel<128><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  } else {
    dim3 grid((out_features + TO - 1) / TO, (B + 31) / 32, G);
    sampled_linear_kernel<32><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  }
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
