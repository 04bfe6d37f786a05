// K5: Monte-Carlo predictive statistics, and K4: KL(q||p) forward + ELBO gradient.
//
// K5 restates, in one pass over logits[S][B][C]:
//   inference/predictors.py:65-84   softmax, var(dim=0).mean(dim=1), mean entropy, argmax mean prob
//   train/multimodal.py:287-310     mean logits, argmax mean logits, H[mean p], mean H[p], MI
//   train/unimodal.py:282-308       softmax(mean logits) argmax, var, mean entropy
// K4 restates bayesian-torch 0.5.0 BaseVariationalLayer_.kl_div (.mean() per tensor) summed
// over layers by get_kl_loss (train/multimodal.py:114,284; train/unimodal.py:130,262), plus
// its autograd gradient w.r.t. mu/rho.
#include "common.cuh"

namespace {

template <int MAXC>
__global__ void __launch_bounds__(128)
mc_reduce_kernel(const float* __restrict__ logits, int S, long long B, int C, float eps_entropy,
                 float* __restrict__ mean_prob, float* __restrict__ mean_logit,
                 long long* __restrict__ argmax_prob, long long* __restrict__ argmax_logit,
                 float* __restrict__ pred_entropy, float* __restrict__ aleatoric,
                 float* __restrict__ mutual_info, float* __restrict__ var_mean) {
  const long long b = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float sum_p[MAXC], sum_l[MAXC], wmean[MAXC], m2[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) { sum_p[c] = 0.f; sum_l[c] = 0.f; wmean[c] = 0.f; m2[c] = 0.f; }
  float sum_h = 0.f;
  for (int s = 0; s < S; ++s) {
    const float* row = logits + (static_cast<long long>(s) * B + b) * C;
    float x[MAXC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      x[c] = (c < C) ? __ldg(row + c) : -INFINITY;
      mx = fmaxf(mx, x[c]);
    }
    float z = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      x[c] = (c < C) ? (sum_l[c] += x[c], expf(x[c] - mx)) : 0.f;
      z += x[c];
    }
    const float inv_n = 1.f / static_cast<float>(s + 1);
    float h = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        const float pc = x[c] / z;
        sum_p[c] += pc;
        const float d = pc - wmean[c];
        wmean[c] += d * inv_n;
        m2[c] = fmaf(d, pc - wmean[c], m2[c]);
        h -= pc * logf(pc + eps_entropy);
      }
    }
    sum_h += h;
  }
  const float inv_s = 1.f / static_cast<float>(S);
  const float inv_sm1 = 1.f / static_cast<float>(S - 1);  // S == 1 -> inf; 0*inf = NaN like torch.var
  float hp = 0.f, vsum = 0.f;
  float best_p = -INFINITY, best_l = -INFINITY;
  int arg_p = 0, arg_l = 0;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < C) {
      const float mp = sum_p[c] * inv_s;
      const float ml = sum_l[c] * inv_s;
      if (mean_prob) mean_prob[b * C + c] = mp;
      if (mean_logit) mean_logit[b * C + c] = ml;
      hp -= mp * logf(mp + eps_entropy);
      vsum += m2[c] * inv_sm1;
      if (mp > best_p) { best_p = mp; arg_p = c; }   // first maximum wins, as torch.argmax
      if (ml > best_l) { best_l = ml; arg_l = c; }
    }
  }
  const float al = sum_h * inv_s;
  if (argmax_prob) argmax_prob[b] = arg_p;
  if (argmax_logit) argmax_logit[b] = arg_l;
  if (pred_entropy) pred_entropy[b] = hp;
  if (aleatoric) aleatoric[b] = al;
  if (mutual_info) mutual_info[b] = hp - al;
  if (var_mean) var_mean[b] = vsum / static_cast<float>(C);
}

// ---------------------------------------------------------------------------
// KL over a table of parameter tensors. table[t] = {mu, rho, grad_mu, grad_rho, n}
// (int64 each, device memory); chunk_prefix[t] = first chunk of tensor t.
// ---------------------------------------------------------------------------
constexpr int KL_CHUNK = 4096;   // elements per (block, iteration): 256 threads x 4 x float4

struct KlTensor { const float* mu; const float* rho; float* gmu; float* grho; long long n; };

__global__ void __launch_bounds__(256)
kl_kernel(const KlTensor* __restrict__ table, const long long* __restrict__ chunk_prefix, int n_tensors,
          long long total_chunks, float prior_mu, float prior_sigma, float grad_scale,
          double* __restrict__ block_partial) {
  __shared__ double red[8];
  const float log_sp = logf(prior_sigma);
  const float inv_2sp2 = 1.f / (2.f * prior_sigma * prior_sigma);
  const float inv_sp2 = 1.f / (prior_sigma * prior_sigma);
  double acc = 0.0;
  for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    // binary search: last t with chunk_prefix[t] <= chunk
    int lo = 0, hi = n_tensors - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (chunk_prefix[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    const KlTensor t = table[lo];
    const long long base = (chunk - chunk_prefix[lo]) * KL_CHUNK;
    const float inv_n = 1.f / static_cast<float>(t.n);
    const float gs = grad_scale * inv_n;
    float local = 0.f;
#pragma unroll
    for (int rep = 0; rep < KL_CHUNK / (256 * 4); ++rep) {
      const long long e0 = base + (rep * 256 + threadIdx.x) * 4;
      if (e0 >= t.n) continue;
      float mu[4], rho[4];
      const bool vec = (e0 + 3 < t.n) && ((reinterpret_cast<uintptr_t>(t.mu + e0) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(t.rho + e0) & 15) == 0);
      if (vec) {
        const float4 a = *reinterpret_cast<const float4*>(t.mu + e0);
        const float4 b = *reinterpret_cast<const float4*>(t.rho + e0);
        mu[0] = a.x; mu[1] = a.y; mu[2] = a.z; mu[3] = a.w;
        rho[0] = b.x; rho[1] = b.y; rho[2] = b.z; rho[3] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          mu[i] = (e0 + i < t.n) ? t.mu[e0 + i] : 0.f;
          rho[i] = (e0 + i < t.n) ? t.rho[e0 + i] : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (e0 + i < t.n) {
          const float ex = expf(rho[i]);
          const float sigma = log1pf(ex);
          const float d = mu[i] - prior_mu;
          local += log_sp - logf(sigma) + (sigma * sigma + d * d) * inv_2sp2 - 0.5f;
          if (t.gmu) {
            t.gmu[e0 + i] += gs * d * inv_sp2;
            const float dsig = sigma * inv_sp2 - 1.f / sigma;
            const float sgm = ex / (1.f + ex);       // d softplus / d rho
            t.grho[e0 + i] += gs * dsig * (isinf(ex) ? 1.f : sgm);
          }
        }
      }
    }
    acc += static_cast<double>(local) * static_cast<double>(inv_n);
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    block_partial[blockIdx.x] = s;
  }
}

__global__ void kl_final_kernel(const double* __restrict__ block_partial, int n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += block_partial[i];
    *out = static_cast<float>(s);
  }
}

}  // namespace

extern "C" {

int mauv_mc_reduce(const void* logits, int S, long long B, int C, int dtype, float eps_entropy,
                   float* mean_prob, float* mean_logit, long long* argmax_prob, long long* argmax_logit,
                   float* pred_entropy, float* aleatoric, float* mutual_info, float* var_mean, void* stream) {
  MAUV_CHECK_ARG(logits && S >= 1 && B >= 1 && C >= 1, "mauv_mc_reduce: bad argument");
  MAUV_CHECK_ARG(dtype == 0, "mauv_mc_reduce: only fp32 logits (dtype 0) are supported");
  MAUV_CHECK_ARG(C <= 32, "mauv_mc_reduce: C=%d > 32 classes not supported", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>(ceil_div_i64(B, 128));
  const float* lg = static_cast<const float*>(logits);
#define MAUV_MC(MAXC)                                                                                   \
  mc_reduce_kernel<MAXC><<<grid, 128, 0, st>>>(lg, S, B, C, eps_entropy, mean_prob, mean_logit,         \
                                               argmax_prob, argmax_logit, pred_entropy, aleatoric,      \
                                               mutual_info, var_mean)
  if (C <= 8) MAUV_MC(8); else if (C <= 16) MAUV_MC(16); else MAUV_MC(32);
#undef MAUV_MC
  MAUV_LAUNCH_CHECK("mc_reduce_kernel");
  return MAUV_OK;
}

int mauv_kl_chunk_elems(void) { return KL_CHUNK; }
// workspace: one double per block
long long mauv_kl_ws_bytes(void) { return static_cast<long long>(mauv_num_sms()) * 8 * sizeof(double); }

int mauv_kl_fwd_bwd(const void* table_dev, const long long* chunk_prefix_dev, int n_tensors,
                    long long total_chunks, float prior_mu, float prior_sigma, float grad_scale,
                    float* kl_out, void* ws, void* stream) {
  MAUV_CHECK_ARG(table_dev && chunk_prefix_dev && kl_out && ws, "mauv_kl_fwd_bwd: null pointer");
  MAUV_CHECK_ARG(n_tensors >= 1 && total_chunks >= 1 && prior_sigma > 0.f, "mauv_kl_fwd_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = static_cast<long long>(mauv_num_sms()) * 8;
  if (blocks > total_chunks) blocks = total_chunks;
  kl_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const KlTensor*>(table_dev), chunk_prefix_dev,
                                                          n_tensors, total_chunks, prior_mu, prior_sigma,
                                                          grad_scale, static_cast<double*>(ws));
  MAUV_LAUNCH_CHECK("kl_kernel");
  kl_final_kernel<<<1, 32, 0, st>>>(static_cast<const double*>(ws), static_cast<int>(blocks), kl_out);
  MAUV_LAUNCH_CHECK("kl_final_kernel");
  return MAUV_OK;
}

}  // extern "C"
