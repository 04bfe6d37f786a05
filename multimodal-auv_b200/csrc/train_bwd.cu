// S-batched ELBO backward (a5/a6): the kernels between the tensor-core contractions when the G Monte-Carlo passes of one
// training step (train/multimodal.py:104-145, train/unimodal.py:125-146) are walked backwards together:
//   train-mode BatchNorm backward with per-(sample, channel) statistics, fused with the ReLU mask and the residual fan-in,
//   max-pool / avg-pool backward, the grouped weight-gradient finalisation (dmu += sum_g dW_g, drho += sum_g dW_g*eps_g*sigmoid(rho)),
//   the fp32 fusion-head backward and cross_entropy(mean_s logits).
// Gradients travel as fp16 with a device-resident power-of-two scale per tensor (value = true gradient * scale). Every
// BatchNorm site re-normalises the scale from the running amax without a host round trip, so the whole backward is
// enqueued asynchronously; parameter gradients are unscaled on accumulation into fp32.
#include "common.cuh"

namespace {

constexpr float kHalfMax = 65504.f;

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    h[j] = __floats2half2_rn(fminf(fmaxf(f[2 * j], -kHalfMax), kHalfMax), fminf(fmaxf(f[2 * j + 1], -kHalfMax), kHalfMax));
  return v;
}

// upstream gradient of a site: d1 (+ d2 brought to d1's scale), masked by the ReLU output when given
struct Upstream {
  const uint4* d1; const uint4* d2; const float* s1; const float* s2; const uint4* mask;
};

__device__ __forceinline__ float upstream_ratio(const Upstream& u) { return u.d2 ? (*u.s1) / (*u.s2) : 0.f; }

__device__ __forceinline__ void load_dz(const Upstream& u, long long idx, float ratio, float* dz) {
  unpack8(__ldg(u.d1 + idx), dz);
  if (u.d2) {
    float t[8];
    unpack8(__ldg(u.d2 + idx), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] = fmaf(ratio, t[j], dz[j]);
  }
  if (u.mask) {
    float m[8];
    unpack8(__ldg(u.mask + idx), m);
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] = m[j] > 0.f ? dz[j] : 0.f;
  }
}

// Two-phase variant for software pipelining: issue every 16-byte load of a vector first, combine afterwards.
struct RawVec { uint4 d1, d2, m, y, y2; };

__device__ __forceinline__ void load_raw(const Upstream& u, const uint4* __restrict__ y, const uint4* __restrict__ y2,
                                         long long idx, bool valid, RawVec& r) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  r.d1 = valid ? __ldg(u.d1 + idx) : z;
  r.d2 = (valid && u.d2) ? __ldg(u.d2 + idx) : z;
  r.m = (valid && u.mask) ? __ldg(u.mask + idx) : z;
  r.y = valid ? __ldg(y + idx) : z;
  r.y2 = (valid && y2) ? __ldg(y2 + idx) : z;
}

__device__ __forceinline__ void raw_dz(const Upstream& u, const RawVec& r, float ratio, float* dz) {
  unpack8(r.d1, dz);
  if (u.d2) {
    float t[8];
    unpack8(r.d2, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] = fmaf(ratio, t[j], dz[j]);
  }
  if (u.mask) {
    float m[8];
    unpack8(r.m, m);
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] = m[j] > 0.f ? dz[j] : 0.f;
  }
}

// ---------------------------------------------------------------- BN backward, pass 1: per-(sample, channel) sums
// partial [G][nblk][3][C] = (sum dz, sum dz*y, sum dz*y2); amax = max |dz| (float bits, atomicMax)
__global__ void __launch_bounds__(256, 3)
bn_bwd_reduce_kernel(Upstream u, const uint4* __restrict__ y, const uint4* __restrict__ y2, long long M, int C, int nblk,
                     long long rows_per_block, float* __restrict__ partial, unsigned* __restrict__ amax) {
  __shared__ float red[3 * 2048];
  const int lanes = C >> 3, rl = 256 / lanes;
  const int cv = threadIdx.x % lanes, r = threadIdx.x / lanes;
  const int g = blockIdx.y, b = blockIdx.x;
  const long long row0 = b * rows_per_block;
  const long long row1 = min(M, row0 + rows_per_block);
  const float ratio = upstream_ratio(u);
  float s0[8], s1[8], s2[8], mx = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s0[j] = s1[j] = s2[j] = 0.f;
  // two rows in flight per thread (up to 10 independent 16-byte loads): the pass is latency-bound otherwise
  for (long long row = row0 + r; row < row1; row += 2 * rl) {
    const long long idx = (static_cast<long long>(g) * M + row) * lanes + cv;
    RawVec ra, rb;
    load_raw(u, y, y2, idx, true, ra);
    load_raw(u, y, y2, idx + static_cast<long long>(rl) * lanes, row + rl < row1, rb);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const RawVec& rv = h ? rb : ra;
      float dz[8], yv[8];
      raw_dz(u, rv, ratio, dz);
      unpack8(rv.y, yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s0[j] += dz[j];
        s1[j] = fmaf(dz[j], yv[j], s1[j]);
        mx = fmaxf(mx, fabsf(dz[j]));
      }
      if (y2) {
        unpack8(rv.y2, yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) s2[j] = fmaf(dz[j], yv[j], s2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[(r * 3 + 0) * C + cv * 8 + j] = s0[j];
    red[(r * 3 + 1) * C + cv * 8 + j] = s1[j];
    red[(r * 3 + 2) * C + cv * 8 + j] = s2[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float acc = 0.f;
      for (int q = 0; q < rl; ++q) acc += red[(q * 3 + k) * C + c];
      partial[((static_cast<long long>(g) * nblk + b) * 3 + k) * C + c] = acc;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f && isfinite(mx)) atomicMax(amax, __float_as_uint(mx));
}

// ---------------------------------------------------------------- BN backward, pass 2: coefficients + affine gradients
// dy = k0*dz + k1*y + k2 with k0 = gamma*invstd, k1 = -k0*invstd*dgamma/M, k2 = -k0*sum(dz)/M - k1*mean
// (dgamma = invstd*(sum dz*y - mean*sum dz)); grad_gamma += sum_g dgamma/s, grad_beta += sum_g sum(dz)/s.
// grid (C/32, G), block (32 channels x 8 partial-lanes): per-(sample, channel) coefficients + (dgamma, sum dz) into tmp
__global__ void __launch_bounds__(256)
bn_bwd_coeffs_kernel(const float* __restrict__ partial, int nblk, int C, int which, double inv_m,
                     const float2* __restrict__ stats, const float* __restrict__ gamma, float eps,
                     float4* __restrict__ coef, double2* __restrict__ tmp, unsigned* __restrict__ kmax) {
  __shared__ double red[2][8][33];
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl, g = blockIdx.y;
  double S0 = 0.0, S1 = 0.0;
  if (c < C) {
    for (int b = bl; b < nblk; b += 8) {
      const long long base = (static_cast<long long>(g) * nblk + b) * 3 * C;
      S0 += partial[base + c];
      S1 += partial[base + static_cast<long long>(which) * C + c];
    }
  }
  red[0][bl][cl] = S0;
  red[1][bl][cl] = S1;
  __syncthreads();
  if (bl != 0) return;
  float km = 0.f;
  if (c < C) {
#pragma unroll
    for (int q = 1; q < 8; ++q) { S0 += red[0][q][cl]; S1 += red[1][q][cl]; }
    const double gam = gamma ? static_cast<double>(gamma[c]) : 1.0;
    const float2 mv = stats[static_cast<long long>(g) * C + c];
    const double mean = mv.x, invstd = 1.0 / sqrt(static_cast<double>(mv.y) + eps);
    const double dgam = invstd * (S1 - mean * S0);
    const double k0 = gam * invstd;
    const double k1 = -k0 * invstd * dgam * inv_m;
    const double k2 = -k0 * S0 * inv_m - k1 * mean;
    coef[static_cast<long long>(g) * C + c] = make_float4(static_cast<float>(k0), static_cast<float>(k1), static_cast<float>(k2), 0.f);
    tmp[static_cast<long long>(g) * C + c] = make_double2(dgam, S0);
    km = fabsf(static_cast<float>(k0));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) km = fmaxf(km, __shfl_xor_sync(0xffffffffu, km, o));
  if (cl == 0 && km > 0.f && isfinite(km)) atomicMax(kmax, __float_as_uint(km));
}

// grad_gamma += sum_g dgamma / s ; grad_beta += sum_g sum(dz) / s (fixed summation order: deterministic)
__global__ void __launch_bounds__(128)
bn_bwd_affine_kernel(const double2* __restrict__ tmp, int G, int C, const float* __restrict__ s_in,
                     float* __restrict__ grad_gamma, float* __restrict__ grad_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double gs = 0.0, bs = 0.0;
  for (int g = 0; g < G; ++g) {
    const double2 t = tmp[static_cast<long long>(g) * C + c];
    gs += t.x;
    bs += t.y;
  }
  const double inv_s = 1.0 / static_cast<double>(*s_in);
  if (grad_gamma) grad_gamma[c] += static_cast<float>(gs * inv_s);
  if (grad_beta) grad_beta[c] += static_cast<float>(bs * inv_s);
}

__device__ __forceinline__ float pow2_rescale(const unsigned* amax, const unsigned* kmax, float target) {
  const float bound = 4.f * __uint_as_float(*amax) * __uint_as_float(*kmax);
  if (!(bound > 0.f) || !isfinite(bound)) return 1.f;
  int e;
  frexpf(target / bound, &e);
  e = max(-60, min(60, e - 1));
  return ldexpf(1.f, e);
}

// ---------------------------------------------------------------- BN backward, pass 3: apply
// dy = r*(k0*dz + k1*y + k2) (fp16, scale s1*r written to s_out); optional second BN sharing dz (downsample branch);
// optional dz output (identity branch of the residual), kept at scale s1.
struct ApplyArgs {
  Upstream u;
  const uint4* y; const uint4* y2;
  const float4* coef; const float4* coef2;
  const unsigned* amax; const unsigned* kmax; const unsigned* kmax2;
  float target;
  uint4* dy; uint4* dy2; uint4* dz;
  float* s_out; float* s_out2;
  long long M; int C;
};

template <bool HAS_Y2>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const ApplyArgs a) {
  const int lanes = a.C >> 3;
  const int g = blockIdx.y;
  const long long vecs = a.M * lanes;
  const long long first = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  const long long step = static_cast<long long>(gridDim.x) * 256;
  const int cv = static_cast<int>(first % lanes);           // invariant: 256 % lanes == 0
  const float ratio = upstream_ratio(a.u);
  const float r1 = pow2_rescale(a.amax, a.kmax, a.target);
  const float r2 = HAS_Y2 ? pow2_rescale(a.amax, a.kmax2, a.target) : 1.f;
  if (blockIdx.x == 0 && g == 0 && threadIdx.x == 0) {
    *a.s_out = (*a.u.s1) * r1;
    if (HAS_Y2) *a.s_out2 = (*a.u.s1) * r2;
  }
  float k0[8], k1[8], k2[8], q0[HAS_Y2 ? 8 : 1], q1[HAS_Y2 ? 8 : 1], q2[HAS_Y2 ? 8 : 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 k = a.coef[static_cast<long long>(g) * a.C + cv * 8 + j];
    k0[j] = k.x * r1; k1[j] = k.y * r1; k2[j] = k.z * r1;
    if (HAS_Y2) {
      const float4 q = a.coef2[static_cast<long long>(g) * a.C + cv * 8 + j];
      q0[j] = q.x * r2; q1[j] = q.y * r2; q2[j] = q.z * r2;
    }
  }
  const long long base = static_cast<long long>(g) * vecs;
  const uint4* y2p = HAS_Y2 ? a.y2 : nullptr;
  for (long long i = first; i < vecs; i += 2 * step) {      // two vectors in flight per thread
    RawVec ra, rb;
    const bool vb = i + step < vecs;
    load_raw(a.u, a.y, y2p, base + i, true, ra);
    load_raw(a.u, a.y, y2p, base + i + step, vb, rb);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h && !vb) break;
      const RawVec& rv = h ? rb : ra;
      const long long o_idx = base + i + (h ? step : 0);
      float dz[8], yv[8], o[8];
      raw_dz(a.u, rv, ratio, dz);
      unpack8(rv.y, yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(k0[j], dz[j], fmaf(k1[j], yv[j], k2[j]));
      a.dy[o_idx] = pack8(o);
      if (HAS_Y2) {
        unpack8(rv.y2, yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(q0[j], dz[j], fmaf(q1[j], yv[j], q2[j]));
        a.dy2[o_idx] = pack8(o);
      }
      if (a.dz) a.dz[o_idx] = pack8(dz);
    }
  }
}

// ---------------------------------------------------------------- stem: backward of maxpool3x3/2(relu(bn(y)))
// Pass A (pooled-output centric, the forward's access pattern): arg-max position 0..8 of every 3x3 window per channel
// (first maximum in row-major scan order wins, as ATen's max_pool2d), 15 when the window maximum is <= 0 (ReLU kills it).
__global__ void __launch_bounds__(128, 5)
maxpool_argmax_kernel(const uint4* __restrict__ y, const float2* __restrict__ ss, int imgs_per_sample, int H, int W, int C,
                      int Ho, int Wo, uint2* __restrict__ idx) {
  const int cvec = C >> 3;
  const int p = blockIdx.x;
  const long long n = blockIdx.y;
  const int g = static_cast<int>(n / imgs_per_sample);
  const int per_row = Wo * cvec;
  for (int t = threadIdx.x; t < per_row; t += blockDim.x) {
    const int cv = t % cvec;
    const int q = t / cvec;
    float sc[8], sh[8], m[8];
    uint32_t am[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 s = ss[static_cast<long long>(g) * C + cv * 8 + j];
      sc[j] = s.x; sh[j] = s.y; m[j] = 0.f; am[j] = 15u;     // only strictly positive values can win
    }
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
#pragma unroll
      for (int ds = 0; ds < 3; ++ds) {
        const int h = p * 2 - 1 + dr, w = q * 2 - 1 + ds;
        if (h >= 0 && h < H && w >= 0 && w < W) {
          float f[8];
          unpack8(__ldg(y + ((n * H + h) * W + w) * cvec + cv), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v = fmaf(f[j], sc[j], sh[j]);
            if (v > m[j]) { m[j] = v; am[j] = dr * 3 + ds; }
          }
        }
      }
    }
    uint2 o;
    o.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
    o.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
    idx[((n * Ho + p) * Wo + q) * cvec + cv] = o;
  }
}

// Pass B (stem-element centric gather): dz[n,h,w,c] = sum over the <= 4 windows containing (h,w) of d[n,p,q,c] * [idx == me]
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const uint2* __restrict__ idx, Upstream u, int H, int W, int C, int Ho, int Wo, long long total,
                   uint4* __restrict__ dz) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int lanes = C >> 3;
  const int cv = static_cast<int>(i % lanes);
  long long t = i / lanes;
  const int w = static_cast<int>(t % W); t /= W;
  const int h = static_cast<int>(t % H);
  const long long n = t / H;
  const float ratio = upstream_ratio(u);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int p_lo = h >> 1, p_hi = min((h + 1) >> 1, Ho - 1);
  const int q_lo = w >> 1, q_hi = min((w + 1) >> 1, Wo - 1);
  Upstream un = u;
  un.mask = nullptr;
  for (int p = p_lo; p <= p_hi; ++p) {
    for (int q = q_lo; q <= q_hi; ++q) {
      const uint32_t me = static_cast<uint32_t>((h - 2 * p + 1) * 3 + (w - 2 * q + 1));
      const long long o = ((n * Ho + p) * Wo + q) * lanes + cv;
      const uint2 id = __ldg(idx + o);
      float d[8];
      load_dz(un, o, ratio, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t a = ((j < 4 ? id.x : id.y) >> (8 * (j & 3))) & 0xffu;
        acc[j] += a == me ? d[j] : 0.f;
      }
    }
  }
  dz[i] = pack8(acc);
}

// ---------------------------------------------------------------- avgpool backward (+ the first loss scale)
__global__ void __launch_bounds__(256)
amax_f32_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ amax) {
  float mx = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    mx = fmaxf(mx, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f && isfinite(mx)) atomicMax(amax, __float_as_uint(mx));
}

__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const float* __restrict__ dfeat, const unsigned* __restrict__ amax, float target, int HW, int C,
                   long long total, __half* __restrict__ out, float* __restrict__ s_out) {
  const float am = __uint_as_float(*amax);
  float r = 1.f;
  if (am > 0.f && isfinite(am)) {
    int e;
    frexpf(target * static_cast<float>(HW) / am, &e);
    r = ldexpf(1.f, max(-60, min(60, e - 1)));
  }
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0) *s_out = r;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const long long n = i / (static_cast<long long>(C) * HW);
  out[i] = __float2half_rn(dfeat[n * C + c] * (r / static_cast<float>(HW)));
}

// ---------------------------------------------------------------- grouped weight-gradient finalisation
// dw fp16 [G*splits][cout][Kp] (K order (r,s,c)), value = true dW * (*s) / inv_alpha.
// Block = 32 element-quads x 8 sample lanes: a thread finalises 4 consecutive PyTorch-order elements (one Philox4x32 block
// per sample) for the samples g = lane, lane+8, ...; the 8 lanes are then summed in a fixed order (deterministic).
template <typename DW>   // __half (explicit-transpose path) or float (direct weight-gradient GEMM) partial sums
__global__ void __launch_bounds__(256)
wgrad_finalize_group_kernel(const DW* __restrict__ dw, int G, int splits, int cout, int cin, int kh, int kw, int Kp,
                            float inv_alpha, const float* __restrict__ s, const float* __restrict__ rho,
                            const float* __restrict__ eps, uint64_t seed, uint32_t layer_id, uint32_t sample0, int stale,
                            float* __restrict__ grad_mu, float* __restrict__ grad_rho,
                            const unsigned int* __restrict__ sample_base) {
  if (sample_base) sample0 += *sample_base;      // device-resident Philox sample-id base (mauv_set_sample_base)
  __shared__ float red[8][32][9];
  const int ql = threadIdx.x & 31, gl = threadIdx.x >> 5;
  const long long quad = static_cast<long long>(blockIdx.x) * 32 + ql;
  const long long e0 = quad * 4;
  const int khw = kh * kw;
  const unsigned per_out = static_cast<unsigned>(cin) * khw;
  const long long n = static_cast<long long>(per_out) * cout;       // < 2^31 (checked on the host): 32-bit index math
  const bool any = e0 < n;
  long long off[4];
  bool live[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long e = e0 + i;
    live[i] = e < n;
    const unsigned ee = live[i] ? static_cast<unsigned>(e) : 0u;
    const unsigned co = ee / per_out;
    const unsigned rem = ee - co * per_out;
    const unsigned c = rem / static_cast<unsigned>(khw), rs = rem - c * khw;
    off[i] = static_cast<long long>(co) * Kp + static_cast<long long>(rs) * cin + c;
  }
  const bool vec = khw == 1 && live[3] && (cin % 4 == 0) && (Kp % 4 == 0);     // 4 consecutive k: one 8-byte load
  const long long gstride = static_cast<long long>(cout) * Kp;
  float am[4] = {0.f, 0.f, 0.f, 0.f}, ar[4] = {0.f, 0.f, 0.f, 0.f};
  if (any) {
    for (int g = gl; g < G; g += 8) {
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      for (int sp = 0; sp < splits; ++sp) {
        const DW* base = dw + (static_cast<long long>(g) * splits + sp) * gstride;
        if constexpr (sizeof(DW) == 4) {
          if (vec) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(base + off[0]));
            d[0] += v.x; d[1] += v.y; d[2] += v.z; d[3] += v.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] += live[i] ? __ldg(base + off[i]) : 0.f;
          }
        } else {
          if (vec) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(base + off[0]));
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
            d[0] += a.x; d[1] += a.y; d[2] += b.x; d[3] += b.y;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] += live[i] ? __half2float(__ldg(base + off[i])) : 0.f;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) am[i] += d[i];
      if (!stale) {
        float z[4];
        if (eps) {
#pragma unroll
          for (int i = 0; i < 4; ++i) z[i] = live[i] ? eps[static_cast<long long>(g) * n + e0 + i] : 0.f;
        } else {
          philox_normals4(seed, layer_id, sample0 + g, static_cast<uint64_t>(quad), z);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) ar[i] = fmaf(d[i], z[i], ar[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { red[gl][ql][i] = am[i]; red[gl][ql][4 + i] = ar[i]; }
  __syncthreads();
  if (gl != 0 || !any) return;
#pragma unroll
  for (int q = 1; q < 8; ++q) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { am[i] += red[q][ql][i]; ar[i] += red[q][ql][4 + i]; }
  }
  if (stale) {   // reference quirk: the saved eps buffer holds the last pass's draw for every pass
    float z[4];
    if (eps) {
#pragma unroll
      for (int i = 0; i < 4; ++i) z[i] = live[i] ? eps[static_cast<long long>(G - 1) * n + e0 + i] : 0.f;
    } else {
      philox_normals4(seed, layer_id, sample0 + G - 1, static_cast<uint64_t>(quad), z);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) ar[i] = am[i] * z[i];
  }
  const float inv = inv_alpha / (*s);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (!live[i]) continue;
    const float ex = expf(rho[e0 + i]);
    const float sgm = isinf(ex) ? 1.f : ex / (1.f + ex);
    grad_mu[e0 + i] += am[i] * inv;
    grad_rho[e0 + i] += ar[i] * inv * sgm;
  }
}

// forward sample w [G][cout][K = (r,s,c)] -> data-gradient operand wd [G][cin][(kh*kw-1-rs)*cout + co]: a per-(sample, tap)
// [cout x cin] -> [cin x cout] transpose through shared memory (replaces re-sampling the weights for the backward pass)
__global__ void __launch_bounds__(256)
weights_to_dgrad_kernel(const __half* __restrict__ w, int cout, int cin, int khw, __half* __restrict__ wd) {
  __shared__ __half tile[64][66];
  const int g = blockIdx.z;
  const int cblocks = (cin + 63) >> 6;
  const int rs = blockIdx.y / cblocks, c0 = (blockIdx.y - rs * cblocks) << 6;
  const int co0 = blockIdx.x << 6;
  const long long K = static_cast<long long>(khw) * cin, Kd = static_cast<long long>(khw) * cout;
  const __half* src = w + static_cast<long long>(g) * cout * K;
  __half* dst = wd + static_cast<long long>(g) * cin * Kd;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int row = (threadIdx.x >> 3) + it * 32, cv = threadIdx.x & 7;
    const int co = co0 + row, c = c0 + cv * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (co < cout && c < cin) v = __ldg(reinterpret_cast<const uint4*>(src + co * K + static_cast<long long>(rs) * cin + c));
    uint32_t* t32 = reinterpret_cast<uint32_t*>(&tile[row][cv * 8]);
    t32[0] = v.x; t32[1] = v.y; t32[2] = v.z; t32[3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int og = threadIdx.x & 7, cl = (threadIdx.x >> 3) + it * 32;
    const int co = co0 + og * 8, c = c0 + cl;
    if (co < cout && c < cin) {
      __align__(16) __half o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = tile[og * 8 + j][cl];
      *reinterpret_cast<uint4*>(dst + c * Kd + static_cast<long long>(khw - 1 - rs) * cout + co) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

// ---------------------------------------------------------------- head backward, grouped over the G samples (fp32)
constexpr int LT = 16;
struct LinG {
  const float* x; long long x_sg; int x_sb;       // [G][B][in]
  const float* gy; long long gy_sg; int gy_sb;    // [G][B][out]
  const float* mu_w; const float* rho_w; const float* eps_w; const float* rho_b; const float* eps_b;
  uint64_t seed; uint32_t layer_id, sample0;
  const unsigned int* sample_base;     // optional device word added to the sample ids (mauv_set_sample_base)
  int G, B, in, out, stale, accumulate;
  float* gx; long long gx_sg; int gx_sb;          // [G][B][in] (= or +=)
  float* gmu_w; float* grho_w; float* gmu_b; float* grho_b;
};

__global__ void __launch_bounds__(256)
linear_bwd_data_group_kernel(const LinG p) {
  __shared__ float gs[LT][LT + 1], ws[LT][LT + 1];
  const int tx = threadIdx.x % LT, ty = threadIdx.x / LT;
  const int i = blockIdx.x * LT + tx, b = blockIdx.y * LT + ty, g = blockIdx.z;
  const float* gy = p.gy + g * p.gy_sg;
  const float* ew = p.eps_w ? p.eps_w + static_cast<long long>(g) * p.out * p.in : nullptr;
  float acc = 0.f;
  for (int o0 = 0; o0 < p.out; o0 += LT) {
    const int ob = o0 + tx;
    gs[ty][tx] = (b < p.B && ob < p.out) ? gy[static_cast<long long>(b) * p.gy_sb + ob] : 0.f;
    const int ow = o0 + ty;
    float w = 0.f;
    if (ow < p.out && i < p.in) {
      const long long e = static_cast<long long>(ow) * p.in + i;
      const float z = ew ? ew[e] : philox_normal(p.seed, p.layer_id, (p.sample0 + (p.sample_base ? *p.sample_base : 0u)) + g, static_cast<uint64_t>(e));
      w = fmaf(softplus_ref(p.rho_w[e]), z, p.mu_w[e]);
    }
    ws[ty][tx] = w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LT; ++k) acc = fmaf(gs[ty][k], ws[k][tx], acc);
    __syncthreads();
  }
  if (b < p.B && i < p.in) {
    float* o = p.gx + g * p.gx_sg + static_cast<long long>(b) * p.gx_sb + i;
    *o = p.accumulate ? *o + acc : acc;
  }
}

__global__ void __launch_bounds__(256)
linear_bwd_weight_group_kernel(const LinG p) {
  __shared__ float gs[LT][LT + 1], xs[LT][LT + 1];
  const int tx = threadIdx.x % LT, ty = threadIdx.x / LT;
  const int i = blockIdx.x * LT + tx, o = blockIdx.y * LT + ty;
  const long long e = static_cast<long long>(o) * p.in + i;
  const bool live = o < p.out && i < p.in;
  const bool bias_lane = p.gmu_b && blockIdx.x == 0 && tx == 0 && o < p.out;
  float am = 0.f, ar = 0.f, bm = 0.f, br = 0.f;
  for (int g = 0; g < p.G; ++g) {
    const float* gy = p.gy + g * p.gy_sg;
    const float* x = p.x + g * p.x_sg;
    float acc = 0.f, accb = 0.f;
    for (int b0 = 0; b0 < p.B; b0 += LT) {
      const int bb = b0 + ty;
      const int oo = blockIdx.y * LT + tx;
      gs[ty][tx] = (bb < p.B && oo < p.out) ? gy[static_cast<long long>(bb) * p.gy_sb + oo] : 0.f;
      xs[ty][tx] = (bb < p.B && i < p.in) ? x[static_cast<long long>(bb) * p.x_sb + i] : 0.f;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < LT; ++k) {
        acc = fmaf(gs[k][ty], xs[k][tx], acc);
        accb += gs[k][ty];
      }
      __syncthreads();
    }
    am += acc;
    bm += accb;
    if (!p.stale) {
      if (live) {
        const float z = p.eps_w ? p.eps_w[static_cast<long long>(g) * p.out * p.in + e]
                                : philox_normal(p.seed, p.layer_id, (p.sample0 + (p.sample_base ? *p.sample_base : 0u)) + g, static_cast<uint64_t>(e));
        ar = fmaf(acc, z, ar);
      }
      if (bias_lane) {
        const float z = p.eps_b ? p.eps_b[static_cast<long long>(g) * p.out + o]
                                : philox_normal(p.seed, p.layer_id | 0x80000000u, (p.sample0 + (p.sample_base ? *p.sample_base : 0u)) + g, static_cast<uint64_t>(o));
        br = fmaf(accb, z, br);
      }
    }
  }
  const int gl = p.G - 1;
  if (live) {
    if (p.stale) {
      const float z = p.eps_w ? p.eps_w[static_cast<long long>(gl) * p.out * p.in + e]
                              : philox_normal(p.seed, p.layer_id, (p.sample0 + (p.sample_base ? *p.sample_base : 0u)) + gl, static_cast<uint64_t>(e));
      ar = am * z;
    }
    const float ex = expf(p.rho_w[e]);
    p.gmu_w[e] += am;
    p.grho_w[e] += ar * (isinf(ex) ? 1.f : ex / (1.f + ex));
  }
  if (bias_lane) {
    if (p.stale) {
      const float z = p.eps_b ? p.eps_b[static_cast<long long>(gl) * p.out + o]
                              : philox_normal(p.seed, p.layer_id | 0x80000000u, (p.sample0 + (p.sample_base ? *p.sample_base : 0u)) + gl, static_cast<uint64_t>(o));
      br = bm * z;
    }
    const float ex = expf(p.rho_b[o]);
    p.gmu_b[o] += bm;
    p.grho_b[o] += br * (isinf(ex) ? 1.f : ex / (1.f + ex));
  }
}

// t = tanh(q + k): dq = dk = dt * (1 - t^2)
__global__ void __launch_bounds__(256)
tanh_bwd_kernel(const float* __restrict__ t, const float* __restrict__ dt, long long n, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dt[i] * (1.f - t[i] * t[i]);
}

// out = v * softmax(score): dv = dout*p ; dscore = p * (dout*v - sum_j p_j*dout_j*v_j). One warp per row.
__global__ void __launch_bounds__(256)
softmax_gate_bwd_kernel(const float* __restrict__ score, const float* __restrict__ v, const float* __restrict__ dout,
                        int ld_dout, long long rows, int n, float* __restrict__ dscore, float* __restrict__ dv) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* s = score + row * n;
  const float* vv = v + row * n;
  const float* d = dout + row * ld_dout;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, s[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float den = 0.f, dot = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float ex = expf(s[j] - mx);
    den += ex;
    dot = fmaf(ex, d[j] * vv[j], dot);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    den += __shfl_xor_sync(0xffffffffu, den, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  const float inv = 1.f / den;
  dot *= inv;
  for (int j = lane; j < n; j += 32) {
    const float pj = expf(s[j] - mx) * inv;
    dv[row * n + j] = d[j] * pj;
    dscore[row * n + j] = pj * (d[j] * vv[j] - dot);
  }
}

// loss = mean over the valid rows b of CE(mean_s logits[s,b,:], label_b); dlogits[s,b,c] = (softmax(mean)[c] - [c == label]) / (n_valid*S).
// Labels follow torch.nn.functional.cross_entropy: -100 (ignore_index) rows are skipped - no loss term, zero gradient, not
// counted in the mean; any other label outside [0, C) is an error that torch reports with a device assert: here the loss
// becomes NaN (the drivers' finite-loss guard skips the batch) and the row gets a zero gradient - never an out-of-bounds read.
__global__ void __launch_bounds__(256)
ce_mean_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int S, int B, int C,
               float* __restrict__ mean_logit, float* __restrict__ dlogits, float* __restrict__ loss) {
  __shared__ float warp_sum[8];
  __shared__ int warp_cnt[8], warp_bad[8];
  __shared__ int n_valid_s, n_bad_s;
  int cnt = 0, bad = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const long long lab = labels[b];
    if (lab >= 0 && lab < C) ++cnt;
    else if (lab != -100) ++bad;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) { warp_cnt[threadIdx.x >> 5] = cnt; warp_bad[threadIdx.x >> 5] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0, d = 0;
    for (int wi = 0; wi < (blockDim.x >> 5); ++wi) { c += warp_cnt[wi]; d += warp_bad[wi]; }
    n_valid_s = c; n_bad_s = d;
  }
  __syncthreads();
  const int n_valid = n_valid_s;
  const float w = n_valid > 0 ? 1.f / (static_cast<float>(n_valid) * static_cast<float>(S)) : 0.f;
  float local = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
      float m = 0.f;
      for (int s = 0; s < S; ++s) m += logits[(static_cast<long long>(s) * B + b) * C + c];
      m /= static_cast<float>(S);
      mean_logit[b * C + c] = m;
      mx = fmaxf(mx, m);
    }
    const long long lab64 = labels[b];
    const bool valid = lab64 >= 0 && lab64 < C;
    const int lab = valid ? static_cast<int>(lab64) : -1;
    float den = 0.f;
    for (int c = 0; c < C; ++c) den += expf(mean_logit[b * C + c] - mx);
    const float lse = mx + logf(den);
    if (valid) local += lse - mean_logit[b * C + lab];
    for (int c = 0; c < C; ++c) {
      const float gr = valid ? (expf(mean_logit[b * C + c] - lse) - (c == lab ? 1.f : 0.f)) * w : 0.f;
      for (int s = 0; s < S; ++s) dlogits[(static_cast<long long>(s) * B + b) * C + c] = gr;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int wi = 0; wi < (blockDim.x >> 5); ++wi) t += warp_sum[wi];
    // all rows ignored: torch returns NaN (0 / 0) as well
    *loss = (n_bad_s > 0 || n_valid == 0) ? __int_as_float(0x7fc00000) : t / static_cast<float>(n_valid);
  }
}

Upstream make_up(const void* d1, const void* d2, const float* s1, const float* s2, const void* mask) {
  Upstream u;
  u.d1 = static_cast<const uint4*>(d1); u.d2 = static_cast<const uint4*>(d2);
  u.s1 = s1; u.s2 = s2; u.mask = static_cast<const uint4*>(mask);
  return u;
}

bool pow2_channels(int C) { return C >= 64 && C <= 2048 && (C & (C - 1)) == 0; }

}  // namespace

extern "C" {

int mauv_bn_bwd_blocks(int G, long long M, int C) {
  // enough row blocks that G * blocks fills the GPU (~8 CTAs of 256 threads per SM), at least two row sweeps per block
  const int rl = 256 / (C >> 3 > 0 ? C >> 3 : 1);
  long long by_rows = (M + 2LL * rl - 1) / (2LL * rl);
  long long by_fill = (static_cast<long long>(mauv_num_sms()) * 8 + G - 1) / G;
  long long nb = by_rows < by_fill ? by_rows : by_fill;
  return static_cast<int>(nb < 1 ? 1 : (nb > 64 ? 64 : nb));
}

int mauv_weights_to_dgrad_f16(const void* w, int G, int cout, int cin, int kh, int kw, void* w_out, void* stream) {
  MAUV_CHECK_ARG(w && w_out && G >= 1 && G <= 65535 && cout % 8 == 0 && cin % 8 == 0, "mauv_weights_to_dgrad_f16: bad argument");
  const int khw = kh * kw;
  dim3 grid((cout + 63) / 64, khw * ((cin + 63) / 64), G);
  MAUV_CHECK_ARG(grid.y <= 65535, "mauv_weights_to_dgrad_f16: too many (tap, channel-block) pairs");
  weights_to_dgrad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(w), cout, cin, khw,
                                                                                static_cast<__half*>(w_out));
  MAUV_LAUNCH_CHECK("weights_to_dgrad_kernel");
  return MAUV_OK;
}

int mauv_bn_bwd_reduce(const void* d1, const void* d2, const float* s1, const float* s2, const void* relu_out, const void* y,
                       const void* y2, int G, long long M, int C, float* partial, unsigned int* amax, void* stream) {
  MAUV_CHECK_ARG(d1 && s1 && y && partial && amax && (!d2 || s2), "mauv_bn_bwd_reduce: null pointer");
  MAUV_CHECK_ARG(pow2_channels(C) && G >= 1 && M >= 1, "mauv_bn_bwd_reduce: C must be a power of two in [64, 2048] (got %d)", C);
  const int nblk = mauv_bn_bwd_blocks(G, M, C);
  const long long rpb = (M + nblk - 1) / nblk;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(nblk, G);
  bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(make_up(d1, d2, s1, s2, relu_out), static_cast<const uint4*>(y),
                                             static_cast<const uint4*>(y2), M, C, nblk, rpb, partial, amax);
  MAUV_LAUNCH_CHECK("bn_bwd_reduce_kernel");
  return MAUV_OK;
}

int mauv_bn_bwd_coeffs(const float* partial, int G, long long M, int C, int which, const float* batch_stats,
                       const float* gamma, float eps, const float* s_in, float* grad_gamma, float* grad_beta, float* coef,
                       unsigned int* kmax, void* ws, void* stream) {
  MAUV_CHECK_ARG(partial && batch_stats && s_in && coef && kmax && ws && (which == 1 || which == 2), "mauv_bn_bwd_coeffs: bad argument");
  MAUV_CHECK_ARG(G >= 1 && G <= 65535, "mauv_bn_bwd_coeffs: G out of range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((C + 31) / 32, G);
  bn_bwd_coeffs_kernel<<<grid, 256, 0, st>>>(partial, mauv_bn_bwd_blocks(G, M, C), C, which, 1.0 / static_cast<double>(M),
                                             reinterpret_cast<const float2*>(batch_stats), gamma, eps,
                                             reinterpret_cast<float4*>(coef), static_cast<double2*>(ws), kmax);
  MAUV_LAUNCH_CHECK("bn_bwd_coeffs_kernel");
  if (grad_gamma || grad_beta) {
    bn_bwd_affine_kernel<<<(C + 127) / 128, 128, 0, st>>>(static_cast<const double2*>(ws), G, C, s_in, grad_gamma, grad_beta);
    MAUV_LAUNCH_CHECK("bn_bwd_affine_kernel");
  }
  return MAUV_OK;
}

int mauv_bn_bwd_apply(const void* d1, const void* d2, const float* s1, const float* s2, const void* relu_out, const void* y,
                      const void* y2, const float* coef, const float* coef2, const unsigned int* amax,
                      const unsigned int* kmax, const unsigned int* kmax2, float target, int G, long long M, int C, void* dy,
                      void* dy2, void* dz, float* s_out, float* s_out2, void* stream) {
  MAUV_CHECK_ARG(d1 && s1 && y && coef && amax && kmax && dy && s_out && (!d2 || s2), "mauv_bn_bwd_apply: null pointer");
  MAUV_CHECK_ARG(!y2 || (coef2 && kmax2 && dy2 && s_out2), "mauv_bn_bwd_apply: the second BatchNorm needs coef2/kmax2/dy2/s_out2");
  MAUV_CHECK_ARG(pow2_channels(C) && G >= 1 && M >= 1, "mauv_bn_bwd_apply: C must be a power of two in [64, 2048] (got %d)", C);
  ApplyArgs a;
  a.u = make_up(d1, d2, s1, s2, relu_out);
  a.y = static_cast<const uint4*>(y); a.y2 = static_cast<const uint4*>(y2);
  a.coef = reinterpret_cast<const float4*>(coef); a.coef2 = reinterpret_cast<const float4*>(coef2);
  a.amax = amax; a.kmax = kmax; a.kmax2 = kmax2; a.target = target;
  a.dy = static_cast<uint4*>(dy); a.dy2 = static_cast<uint4*>(dy2); a.dz = static_cast<uint4*>(dz);
  a.s_out = s_out; a.s_out2 = s_out2; a.M = M; a.C = C;
  const long long vecs = M * (C / 8);
  long long bx = ceil_div_i64(vecs, 256 * 16);   // >= 16 vectors per thread: the 48-float coefficient prologue amortises
  const long long cap = static_cast<long long>(mauv_num_sms()) * 8;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<unsigned>(bx), G);
  if (y2) bn_bwd_apply_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  else bn_bwd_apply_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  MAUV_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return MAUV_OK;
}

int mauv_maxpool_bwd_f16(const void* y, const float* scale_shift, const void* d1, const void* d2, const float* s1,
                         const float* s2, int G, int imgs_per_sample, int H, int W, int C, void* idx_ws, void* dz, void* stream) {
  MAUV_CHECK_ARG(y && scale_shift && d1 && s1 && dz && idx_ws && (!d2 || s2) && C % 8 == 0, "mauv_maxpool_bwd_f16: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long imgs = static_cast<long long>(G) * imgs_per_sample;
  MAUV_CHECK_ARG(imgs <= 65535, "mauv_maxpool_bwd_f16: at most 65535 images per call (got %lld)", imgs);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int threads = Wo * (C / 8);
  threads = threads > 128 ? 128 : ((threads + 31) / 32) * 32;
  maxpool_argmax_kernel<<<dim3(Ho, static_cast<unsigned>(imgs)), threads, 0, st>>>(
      static_cast<const uint4*>(y), reinterpret_cast<const float2*>(scale_shift), imgs_per_sample, H, W, C, Ho, Wo,
      static_cast<uint2*>(idx_ws));
  MAUV_LAUNCH_CHECK("maxpool_argmax_kernel");
  const long long total = imgs * H * W * (C / 8);
  maxpool_bwd_kernel<<<static_cast<unsigned>(ceil_div_i64(total, 256)), 256, 0, st>>>(
      static_cast<const uint2*>(idx_ws), make_up(d1, d2, s1, s2, nullptr), H, W, C, Ho, Wo, total, static_cast<uint4*>(dz));
  MAUV_LAUNCH_CHECK("maxpool_bwd_kernel");
  return MAUV_OK;
}

int mauv_avgpool_bwd_f16(const float* dfeat, long long N, int HW, int C, float target, unsigned int* amax_ws, void* out,
                         float* s_out, void* stream) {
  MAUV_CHECK_ARG(dfeat && amax_ws && out && s_out && N >= 1 && HW >= 1 && C >= 1, "mauv_avgpool_bwd_f16: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nf = N * C;
  const int64_t ab = ceil_div_i64(nf, 256);
  amax_f32_kernel<<<static_cast<unsigned>(ab > 1024 ? 1024 : ab), 256, 0, st>>>(dfeat, nf, amax_ws);
  MAUV_LAUNCH_CHECK("amax_f32_kernel");
  const long long total = nf * HW;
  avgpool_bwd_kernel<<<static_cast<unsigned>(ceil_div_i64(total, 256)), 256, 0, st>>>(dfeat, amax_ws, target, HW, C, total,
                                                                                     static_cast<__half*>(out), s_out);
  MAUV_LAUNCH_CHECK("avgpool_bwd_kernel");
  return MAUV_OK;
}

int mauv_wgrad_finalize_group(const void* dw_partial, int partial_f32, int G, int splits, int cout, int cin, int kh, int kw, int k_pad,
                              float inv_alpha, const float* scale, const float* rho, const float* eps, uint64_t seed,
                              uint32_t layer_id, uint32_t sample0, int stale_eps, float* grad_mu, float* grad_rho,
                              void* stream) {
  MAUV_CHECK_ARG(dw_partial && scale && rho && grad_mu && grad_rho && G >= 1 && splits >= 1, "mauv_wgrad_finalize_group: bad argument");
  const long long n = static_cast<long long>(cout) * cin * kh * kw;
  MAUV_CHECK_ARG(n < (1LL << 31), "mauv_wgrad_finalize_group: tensor too large");
  const unsigned grid = static_cast<unsigned>(ceil_div_i64(ceil_div_i64(n, 4), 32));
  if (partial_f32)
    wgrad_finalize_group_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(dw_partial), G, splits, cout, cin, kh, kw, k_pad, inv_alpha, scale, rho, eps, seed, layer_id,
        sample0, stale_eps, grad_mu, grad_rho, mauv_sample_base());
  else
    wgrad_finalize_group_kernel<__half><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(dw_partial), G, splits, cout, cin, kh, kw, k_pad, inv_alpha, scale, rho, eps, seed, layer_id,
        sample0, stale_eps, grad_mu, grad_rho, mauv_sample_base());
  MAUV_LAUNCH_CHECK("wgrad_finalize_group_kernel");
  return MAUV_OK;
}

int mauv_sampled_linear_bwd_group_f32(const float* x, long long x_sg, int x_sb, const float* gy, long long gy_sg, int gy_sb,
                                      const float* mu_w, const float* rho_w, const float* eps_w, const float* rho_b,
                                      const float* eps_b, uint64_t seed, uint32_t layer_id, uint32_t sample0, int G, int B,
                                      int in_features, int out_features, int stale_eps, float* gx, long long gx_sg, int gx_sb,
                                      int accumulate_gx, float* grad_mu_w, float* grad_rho_w, float* grad_mu_b,
                                      float* grad_rho_b, void* stream) {
  MAUV_CHECK_ARG(x && gy && mu_w && rho_w && grad_mu_w && grad_rho_w && G >= 1 && G <= 65535, "mauv_sampled_linear_bwd_group_f32: bad argument");
  MAUV_CHECK_ARG((grad_mu_b == nullptr) == (grad_rho_b == nullptr) && (grad_mu_b == nullptr || rho_b != nullptr),
                 "mauv_sampled_linear_bwd_group_f32: bias gradient pointers go together with rho_b");
  LinG p;
  p.x = x; p.x_sg = x_sg; p.x_sb = x_sb; p.gy = gy; p.gy_sg = gy_sg; p.gy_sb = gy_sb;
  p.mu_w = mu_w; p.rho_w = rho_w; p.eps_w = eps_w; p.rho_b = rho_b; p.eps_b = eps_b;
  p.seed = seed; p.layer_id = layer_id; p.sample0 = sample0; p.G = G; p.B = B; p.in = in_features; p.out = out_features;
  p.sample_base = mauv_sample_base();
  p.stale = stale_eps; p.accumulate = accumulate_gx; p.gx = gx; p.gx_sg = gx_sg; p.gx_sb = gx_sb;
  p.gmu_w = grad_mu_w; p.grho_w = grad_rho_w; p.gmu_b = grad_mu_b; p.grho_b = grad_rho_b;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gx) {
    dim3 g1((in_features + LT - 1) / LT, (B + LT - 1) / LT, G);
    linear_bwd_data_group_kernel<<<g1, 256, 0, st>>>(p);
    MAUV_LAUNCH_CHECK("linear_bwd_data_group_kernel");
  }
  dim3 g2((in_features + LT - 1) / LT, (out_features + LT - 1) / LT);
  linear_bwd_weight_group_kernel<<<g2, 256, 0, st>>>(p);
  MAUV_LAUNCH_CHECK("linear_bwd_weight_group_kernel");
  return MAUV_OK;
}

int mauv_tanh_bwd_f32(const float* t, const float* dt, long long n, float* out, void* stream) {
  MAUV_CHECK_ARG(t && dt && out, "mauv_tanh_bwd_f32: null pointer");
  tanh_bwd_kernel<<<static_cast<unsigned>(ceil_div_i64(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(t, dt, n, out);
  MAUV_LAUNCH_CHECK("tanh_bwd_kernel");
  return MAUV_OK;
}

int mauv_softmax_gate_bwd_f32(const float* score, const float* v, const float* dout, int ld_dout, long long rows, int n,
                              float* dscore, float* dv, void* stream) {
  MAUV_CHECK_ARG(score && v && dout && dscore && dv && n >= 1 && ld_dout >= n, "mauv_softmax_gate_bwd_f32: bad argument");
  softmax_gate_bwd_kernel<<<static_cast<unsigned>(ceil_div_i64(rows, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      score, v, dout, ld_dout, rows, n, dscore, dv);
  MAUV_LAUNCH_CHECK("softmax_gate_bwd_kernel");
  return MAUV_OK;
}

int mauv_ce_mean_fwd_bwd_f32(const float* logits, const long long* labels, int S, int B, int C, float* mean_logit,
                             float* dlogits, float* loss, void* stream) {
  MAUV_CHECK_ARG(logits && labels && mean_logit && dlogits && loss && S >= 1 && B >= 1 && C >= 1, "mauv_ce_mean_fwd_bwd_f32: bad argument");
  ce_mean_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, labels, S, B, C, mean_logit, dlogits, loss);
  MAUV_LAUNCH_CHECK("ce_mean_kernel");
  return MAUV_OK;
}

}  // extern "C"
