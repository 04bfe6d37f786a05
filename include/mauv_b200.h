/* mauv_b200 — C-ABI of the B200-native Monte-Carlo Bayesian hot path of Multimodal-AUV.
 *
 * The reference (sams-tom/Multimodal-AUV) is pure Python and has no FFI of its own; the
 * hot path lives behind two Python surfaces: the bayesian-torch layer
 * library (dnn_to_bnn / get_kl_loss / *Reparameterization.forward) and the five driver
 * functions in train/multimodal.py, train/unimodal.py, inference/predictors.py. Each entry
 * point below cites the reference code whose arithmetic it replaces. The Python package
 * `mauv` binds these symbols with ctypes (multimodal-auv_b200/mauv/_lib.py) and mirrors
 * the reference's Python API on top.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero (MAUV_ERR_*) on failure; the message is
 *    in mauv_last_error() (thread local). No C++ exception crosses this boundary.
 *  - all tensor pointers are caller-owned DEVICE pointers on the current CUDA device; the
 *    library never allocates or frees device memory and keeps no pointer past return.
 *  - `stream` is a cudaStream_t (CUstream). Work is only enqueued, never synchronised.
 *  - activations are NHWC fp16, parameters fp32 in the PyTorch layout ([Cout][Cin][kh][kw],
 *    [out][in]), accumulation and statistics fp32 (fp64 where noted).
 *  - requires an sm_100a device (mauv_device_check); there is no fallback of any kind.
 */
#ifndef MAUV_B200_H
#define MAUV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  MAUV_OK = 0,
  MAUV_ERR_BAD_ARG = 1,
  MAUV_ERR_CUDA = 2,
  MAUV_ERR_UNSUPPORTED_ARCH = 3,
  MAUV_ERR_DRIVER = 4
};

/* ---- runtime ---------------------------------------------------------------------- */
int mauv_version(void);
const char* mauv_last_error(void);
int mauv_device_check(void);
int mauv_num_sms_c(void);
/* Device-resident Philox sample-id base for the CALLING THREAD's subsequent launches of the forward sampling entry
 * points (mauv_sample_weights_f16, _scaled_f16, _dgrad_f16, _x3_f16, mauv_sampled_linear_f32) and of the S-batched backward's
 * eps replay (mauv_wgrad_finalize_group, mauv_sampled_linear_bwd_group_f32); NULL = none. Those kernels
 * add *sample_base to their sample ids when they RUN, so a CUDA graph captured with a base set draws fresh eps on every
 * replay once the caller bumps the word - the reference draws fresh eps on every pass of every batch
 * (inference/predictors.py:54-66; bayesian-torch `eps.data.normal_()`). The pointer must stay valid while such work
 * (or a graph holding it) can run. */
int mauv_set_sample_base(const unsigned int* sample_base);

/* ---- K1 operand staging: w = mu + log1p(exp(rho)) * eps ---------------------------------
 * Replaces bayesian-torch 0.5.0 Conv2dReparameterization.forward / LinearReparameterization
 * .forward lines "sigma_weight = log1p(exp(rho)); eps = eps.normal_(); weight = mu + sigma*eps"
 * (layers installed by reference models/model_utils.py:26-35).
 * Writes G sampled copies of ONE layer, fp16, [G][cout][k_pad] with K ordered (kh, kw, cin)
 * and zero padded to k_pad (multiple of 8). eps: injected [G][cout*cin*kh*kw] (PyTorch
 * element order) or NULL -> Philox4x32-10, counter (elem/4, sample0+g, layer_id), key seed. */
int mauv_sample_weights_f16(const float* mu, const float* rho, const float* eps, uint64_t seed,
                            uint32_t layer_id, uint32_t sample0, int G, int cout, int cin, int kh,
                            int kw, int k_pad, void* w_out, void* stream);
/* fp32 sampled vector (biases): out[g][i] = mu[i] + log1p(exp(rho[i])) * eps[g][i]. */
int mauv_sample_vector_f32(const float* mu, const float* rho, const float* eps, uint64_t seed,
                           uint32_t layer_id, uint32_t sample0, int G, int n, float* out, void* stream);
/* The N(0,1) stream itself (Philox + Box-Muller), for cross-checking oracle/philox.py. */
int mauv_philox_normal_f32(uint64_t seed, uint32_t layer_id, uint32_t sample_id, long long n,
                           float* out, void* stream);

/* ---- K1 contraction: tcgen05 implicit-GEMM conv / GEMM, grouped over MC samples ---------
 * Replaces F.conv2d(input, weight, None, stride, padding) / F.linear in the same forward
 * functions, called per MC pass from models/base_models.py:74-90 (torchvision resnet.py
 * Bottleneck.forward :143-165). y[g] = A[g] * W[g]^T, fp16 in, fp32 accumulate, fp16 out.
 * stats_partial (nullable): [G][mauv_gemm_m_tiles(M)][N][2] fp32 per-tile (sum, sum of squares)
 * of the fp32 accumulators = the BatchNorm batch statistics, reduced by mauv_bn_finalize. */
int mauv_gemm_m_tiles(long long M);
/* A: [G][M][K] row-major fp16 (a_sample_stride elements between samples; 0 = one A shared by
 * all samples, e.g. the stem's im2col matrix). W: [G][N][K]. bias (nullable): [G][N] fp32. */
int mauv_gemm_f16(const void* a, long long a_sample_stride, const void* w, const void* bias,
                  void* y, float* stats_partial, int G, long long M, int N, int K, void* stream);
/* Recompute scheme for a bottleneck's last 1x1 conv + BatchNorm + residual + ReLU (torchvision resnet.py:154-163),
 * which is HBM-WRITE bound: mode 1 = batch statistics only (nothing stored, y may be NULL); mode 2 = second pass,
 * out = relu?((A W^T) * scale + shift [+ residual]) written straight to y - the raw conv output never reaches HBM.
 * scale_shift [G][N][2] from mauv_bn_finalize, residual [G][M][N] fp16 or NULL; N in {64, 128, k*256}. */
int mauv_gemm_bn_stats_tiles(long long M);   /* leading dim of stats_partial in mode 1: [G][tiles][N][2] */
int mauv_gemm_bn_f16(const void* a, const void* w, void* y, float* stats_partial, const float* scale_shift,
                     const void* residual, int relu, int mode, int G, long long M, int N, int K, void* stream);
/* x: [G*imgs_per_sample][H][W][Cin] NHWC fp16, fetched with im2col-mode TMA (Cin % 64 == 0);
 * W: [G][Cout][kh*kw*Cin]; y: [G*imgs_per_sample][Ho][Wo][Cout]. */
int mauv_conv2d_im2col_f16(const void* x, const void* w, void* y, float* stats_partial, int G,
                           int imgs_per_sample, int H, int W, int Cin, int Cout, int kh, int kw,
                           int stride, int pad, void* stream);

/* ---- K3: BatchNorm(train) / ReLU / residual / pooling ------------------------------------
 * Replaces nn.BatchNorm2d in training mode (the reference keeps .train() on for every MC
 * pass: inference/predictors.py:27, train/multimodal.py:60,232), ReLU, the residual add,
 * MaxPool2d(3,2,1) and AdaptiveAvgPool2d(1) of torchvision resnet.py:266-282. */
/* x: NCHW fp32 [B][C][H][W] -> explicit im2col matrix [B*Ho*Wo][k_pad] fp16 for the 7x7/2 stem. */
int mauv_stem_im2col_f16(const float* x_nchw, int B, int C, int H, int W, int kh, int kw, int stride,
                         int pad, int k_pad, void* out, void* stream);
long long mauv_bn_finalize_ws_bytes(int G, int m_tiles, int C);
/* scale_shift: [G][C][2] fp32 (y*scale+shift == gamma*(y-mean)/sqrt(var+eps)+beta, biased var);
 * running_mean/var (nullable) receive the G sequential momentum updates (unbiased var) the
 * reference's G passes would apply and num_batches_tracked (nullable, int64) += G;
 * batch_stats (nullable): [G][C][2] = (mean, biased var). G <= 64 per call. */
int mauv_bn_finalize(const float* stats_partial, int G, int m_tiles, int C, long long count,
                     const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var, long long* num_batches_tracked,
                     float* scale_shift, float* batch_stats, void* ws, void* stream);
/* out = relu?( y*ss + [residual] + [y2*ss2] ), all [G][M][C] fp16. colsum_partial (nullable): [G][mauv_bn_act_blocks(G, M, C)][C]
 * fp32 per-block column sums of the fp16 OUTPUT (first moment for mauv_bn_stats_from_gram). */
int mauv_bn_act_blocks(int G, long long M, int C);
int mauv_bn_act_f16(const void* y, const float* scale_shift, const void* residual, const void* y2,
                    const float* scale_shift2, int relu, int G, long long M, int C, void* out, float* colsum_partial,
                    void* stream);
/* Column sums of x [G][M][C] fp16 in the same per-block layout (C: power of two in [8, 2048]). */
int mauv_colsum_f16(const void* x, int G, long long M, int C, float* colsum_partial, void* stream);
/* Closed-form BatchNorm(train) statistics of a 1x1 conv y = a W^T (nn.BatchNorm2d after the bottleneck's conv3 / downsample
 * conv, torchvision resnet.py:154-160) from the moments of its INPUT a [G][M][K]: sum_m y[m][n] = w_n . colsum(a),
 * sum_m y[m][n]^2 = w_n^T (a^T a) w_n. gram_partial: [G*splits][K][K] fp32 from mauv_wgrad_f16(dy = x = a) (pixel-chunk
 * partial sums of a^T a); colsum_partial: [G][nblk][K]; w: the sampled fp16 weights [G][N][K] the contraction uses. Outputs
 * and running-statistics updates exactly as mauv_bn_finalize. Replaces the N x K statistics pass of the recompute scheme
 * (mauv_gemm_bn_f16 mode 1) by a K x K contraction. K % 8 == 0, K <= 512, G <= 64. */
long long mauv_bn_stats_from_gram_ws_bytes(int G, int N, int K);
int mauv_bn_stats_from_gram(const float* gram_partial, int splits, const float* colsum_partial, int nblk, const void* w,
                            int G, int N, int K, long long count, const float* gamma, const float* beta, float eps,
                            float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                            float* scale_shift, float* batch_stats, void* ws, void* stream);
int mauv_bn_relu_maxpool_f16(const void* y, const float* scale_shift, int G, int imgs_per_sample, int H,
                             int W, int C, void* out, void* stream);
/* Inference stem in one kernel: conv1 (7x7/2, im2col matrix a0 [imgs*Ho*128][Kp] shared by all samples, w [G][64][Kp]) + the
 * bn1 batch statistics (stats_partial [G][imgs*Ho][64][2], same layout as mauv_gemm_f16's) + the 3x3/2 max-pool taken on the RAW
 * conv output in the epilogue (window max, or window min for channels with gamma < 0: relu(bn(.)) is monotone per channel, so
 * maxpool(relu(bn1(y))) == relu(bn1(pool(y)))); pooled [G][imgs][Ho/2][64][64] fp16. The full-resolution conv1 output (the
 * largest tensor of the network) never reaches HBM. Replaces conv1 / maxpool of models/base_models.py:74-76 (torchvision
 * resnet.py _forward_impl) for 256 x 256 inputs (Wo = 128); follow with mauv_bn_finalize + mauv_bn_act_f16 on `pooled`. */
int mauv_stem_conv_pool_f16(const void* a0, const void* w, void* pooled, float* stats_partial, const float* gamma, int G,
                            int imgs, int Ho, int Kp, void* stream);
int mauv_avgpool_f16(const void* x, long long N, int HW, int C, float* out, void* stream);
int mauv_nchw_f32_to_nhwc_f16(const float* x, long long N, int C, int HW, int c_pad, void* out, void* stream);
int mauv_nhwc_f16_to_nchw_f32(const void* x, long long N, int C, int HW, float* out, void* stream);

/* ---- K2: fusion head (sampling fused into operand staging, fp32) --------------------------
 * Replaces LinearReparameterization.forward for AdditiveAttention (models/base_models.py:35-52)
 * and fc/fc1/fc2 (:60-65,86-89): y[g] = x[g] * (mu_w + sp(rho_w)*eps_w[g])^T + (mu_b + sp(rho_b)*eps_b[g]).
 * Bias Philox stream uses layer_id | 0x80000000. */
int mauv_sampled_linear_f32(const float* x, long long x_sample_stride, int ldx, const float* mu_w,
                            const float* rho_w, const float* eps_w, const float* mu_b, const float* rho_b,
                            const float* eps_b, uint64_t seed, uint32_t layer_id, uint32_t sample0, int G,
                            int B, int in_features, int out_features, float* y, long long y_sample_stride,
                            int ldy, void* stream);
/* tanh(queries + keys)                              models/base_models.py:47 */
int mauv_tanh_add_f32(const float* a, const float* b, long long n, float* out, void* stream);
/* values * softmax(score, dim=1), row stride ld_out    models/base_models.py:48-51 */
int mauv_softmax_gate_f32(const float* score, const float* v, long long rows, int n, float* out,
                          int ld_out, void* stream);

/* ---- K5: MC predictive statistics over logits[S][B][C] -------------------------------------
 * Replaces inference/predictors.py:65-84, train/multimodal.py:287-310, train/unimodal.py:282-308.
 * Any output pointer may be NULL. var_mean = torch.var(p, dim=0).mean(dim=1) (unbiased; S=1 -> NaN),
 * entropies use log(p + eps_entropy) (1e-7 predictor/unimodal, 1e-8 multimodal eval),
 * mutual_info = pred_entropy - aleatoric, argmax = first maximum. dtype: 0 = fp32. */
int mauv_mc_reduce(const void* logits, int S, long long B, int C, int dtype, float eps_entropy,
                   float* mean_prob, float* mean_logit, long long* argmax_prob, long long* argmax_logit,
                   float* pred_entropy, float* aleatoric, float* mutual_info, float* var_mean, void* stream);

/* ---- K4: KL(q||p) forward + gradient --------------------------------------------------------
 * Replaces bayesian-torch BaseVariationalLayer_.kl_div (per-tensor .mean()) summed over layers by
 * get_kl_loss (train/multimodal.py:114,284; train/unimodal.py:130,262) and its autograd backward.
 * table_dev: device array of n_tensors records {mu*, rho*, grad_mu* (nullable), grad_rho*, n} as
 * five int64; chunk_prefix_dev[t] = sum_{u<t} ceil(n_u / mauv_kl_chunk_elems()).
 * kl_out = sum_t mean_i kl(mu_i, softplus(rho_i)); grads += grad_scale * dKL/d(mu, rho). */
int mauv_kl_chunk_elems(void);
long long mauv_kl_ws_bytes(void);
int mauv_kl_fwd_bwd(const void* table_dev, const long long* chunk_prefix_dev, int n_tensors,
                    long long total_chunks, float prior_mu, float prior_sigma, float grad_scale,
                    float* kl_out, void* ws, void* stream);

/* ---- K6/K7: backward of the sampled layers (autograd of the reference's loss.backward(),
 * train/multimodal.py:138, train/unimodal.py:145): dX = conv_transpose(dY, W_s); dW_s = X^T dY;
 * dmu += dW_s; drho += dW_s * eps_s * sigmoid(rho). Both contractions run on mauv_gemm_f16 /
 * mauv_conv2d_im2col_f16; these are the operand builders and the fused parameter-gradient epilogue. */
/* Same sample as mauv_sample_weights_f16 (same eps) in the data-gradient layout
 * w_out[g][ci][(kh-1-r, kw-1-s, co)]: dX = conv(dY zero-stuffed by the stride, w_out, stride 1, pad k-1-p). */
int mauv_sample_weights_dgrad_f16(const float* mu, const float* rho, const float* eps, uint64_t seed,
                                  uint32_t layer_id, uint32_t sample0, int G, int cout, int cin, int kh,
                                  int kw, void* w_out, void* stream);
/* [N][Ho][Wo][C] fp16 -> zero-stuffed [N][Hd][Wd][C] with the input at (p*stride, q*stride). */
int mauv_dilate_f16(const void* x, long long N, int Ho, int Wo, int C, int Hd, int Wd, int stride, void* out, void* stream);
/* [M][C] fp16 -> [splits][C][M/splits] (pixels contiguous; values multiplied by scale): wgrad GEMM operand. */
int mauv_transpose_chunks_f16(const void* src, long long M, int C, int splits, float scale, void* dst, void* stream);
/* NHWC fp16 -> transposed im2col [splits][k_pad][M/splits], K order (kh, kw, cin), rows >= K zero. */
int mauv_im2col_t_f16(const void* x, long long N, int H, int W, int Cin, int kh, int kw, int stride, int pad,
                      int k_pad, int splits, void* dst, void* stream);
/* dw_partial: fp16 [splits][cout][k_pad] (loss-scaled by 1/inv_scale) -> grad_mu += dW, grad_rho += dW*eps*sigmoid(rho)
 * in the PyTorch layout; eps injected [cout*cin*kh*kw] or NULL -> Philox(seed, layer_id, sample_id). */
int mauv_wgrad_finalize(const void* dw_partial, int splits, int cout, int cin, int kh, int kw, int k_pad, float inv_scale,
                        const float* rho, const float* eps, uint64_t seed, uint32_t layer_id, uint32_t sample_id,
                        float* grad_mu, float* grad_rho, void* stream);
/* Linear layer backward, fp32, sampling fused: gx (nullable) = gy * W_s; weight and bias (nullable) parameter grads += . */
int mauv_sampled_linear_bwd_f32(const float* x, const float* gy, const float* mu_w, const float* rho_w,
                                const float* eps_w, const float* rho_b, const float* eps_b, uint64_t seed,
                                uint32_t layer_id, uint32_t sample_id, int B, int in_features, int out_features,
                                float* gx, float* grad_mu_w, float* grad_rho_w, float* grad_mu_b, float* grad_rho_b,
                                void* stream);

/* ---- fp16x3 validation mode ------------------------------------------------------------------------
 * Every value travels as an fp16 (hi | lo) pair (hi + lo ~ 22 mantissa bits) and a*w is contracted as
 * a_hi*w_hi + a_lo*w_hi + a_hi*w_lo by the SAME tcgen05 kernel over K-concatenated operands
 * A' = [a_hi | a_lo | a_hi] (the third block re-reads the first), W' = [w_hi | w_hi | w_lo]; the epilogue writes the
 * fp32 accumulator as a (hi | lo) pair. ~3x the work of the fast path; exists so that the whole engine can be checked
 * against the fp32 reference at rtol 1e-3 through every layer. */
/* a2 [G][M][2K] (hi | lo), K % 64 == 0 (a_sample_stride 0 = shared); w3 [G][N][3K]; y2 [G][M][2N]. */
int mauv_gemm_x3_f16(const void* a2, long long a_sample_stride, const void* w3, void* y2, float* stats_partial, int G,
                     long long M, int N, int K, void* stream);
/* x2 [G*imgs][H][W][2Cin]; w3 [G][Cout][kh*kw*3*Cin] (per tap: hi | hi | lo); y2 [G*imgs][Ho][Wo][2Cout]. */
int mauv_conv2d_im2col_x3_f16(const void* x2, const void* w3, void* y2, float* stats_partial, int G, int imgs_per_sample,
                              int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, void* stream);
/* scale * (mu + log1p(exp(rho)) * eps) split into [hi | hi | lo]; per_tap = 0: flat over K padded to kp (% 64),
 * row length 3*kp; per_tap = 1: per filter tap, row length kh*kw*3*cin. */
int mauv_sample_weights_x3_f16(const float* mu, const float* rho, const float* eps, uint64_t seed, uint32_t layer_id,
                               uint32_t sample0, int G, int cout, int cin, int kh, int kw, int kp, int per_tap,
                               float scale, void* w_out, void* stream);
int mauv_stem_im2col_x3_f16(const float* x_nchw, int B, int C, int H, int W, int kh, int kw, int stride, int pad, int kp,
                            void* out, void* stream);
int mauv_bn_act_x3_f16(const void* y2, const float* scale_shift, const void* residual2, const void* y2b,
                       const float* scale_shift2, int relu, int G, long long M, int C, void* out2, void* stream);
int mauv_bn_relu_maxpool_x3_f16(const void* y2, const float* scale_shift, int G, int imgs_per_sample, int H, int W, int C,
                                void* out2, void* stream);
int mauv_avgpool_x3_f16(const void* x2, long long N, int HW, int C, float* out, void* stream);

/* ---- S-batched ELBO backward (a5/a6: train/multimodal.py:104-145, train/unimodal.py:125-146) ---------
 * The G Monte-Carlo passes of one training step are walked backwards together. Gradient tensors are fp16 NHWC with a
 * device-resident power-of-two scale each (value = true gradient * *scale): every BatchNorm site renormalises the scale
 * from the amax it measures, no host round trip. "Upstream" of a site = d1 (scale *s1) + optional d2 (scale *s2, e.g. the
 * identity branch of the residual), masked by relu_out > 0 when relu_out is given. C: power of two in [64, 2048]. */
/* number of row blocks pass 1 uses for G samples of M rows, C channels (sizes `partial`: [G][blocks][3][C] floats) */
int mauv_bn_bwd_blocks(int G, long long M, int C);
/* forward sample w [G][cout][kh*kw*cin] (mauv_sample_weights_f16) -> the data-gradient operand [G][cin][kh*kw*cout] with the
 * taps flipped (same layout as mauv_sample_weights_dgrad_f16, without replaying the noise). cout, cin % 8 == 0. */
int mauv_weights_to_dgrad_f16(const void* w, int G, int cout, int cin, int kh, int kw, void* w_out, void* stream);
/* pass 1: partial[g][blk][0..2][c] = sum dz, sum dz*y, sum dz*y2 (y2 nullable); *amax = max(*amax, max|dz|) as float
 * bits. The amax / kmax words of this family are atomicMax targets: the caller hands in ZEROED words (one memset of a
 * scratch buffer per backward walk instead of one per site). */
int mauv_bn_bwd_reduce(const void* d1, const void* d2, const float* s1, const float* s2, const void* relu_out, const void* y,
                       const void* y2, int G, long long M, int C, float* partial, unsigned int* amax, void* stream);
/* pass 2: per-(sample, channel) coefficients of dy = k0*dz + k1*y + k2 for train-mode BN (batch_stats = (mean, biased var)
 * [G][C][2] from mauv_bn_finalize); which = 1 (y) or 2 (y2); grad_gamma/grad_beta (nullable) += sum over the G samples,
 * unscaled by *s_in; coef [G][C][4]; *kmax = max|k0| as float bits. */
int mauv_bn_bwd_coeffs(const float* partial, int G, long long M, int C, int which, const float* batch_stats,
                       const float* gamma, float eps, const float* s_in, float* grad_gamma, float* grad_beta, float* coef,
                       unsigned int* kmax, void* ws /* G*C*16 bytes, 16-byte aligned */, void* stream);
/* pass 3: dy = r*(k0*dz + k1*y + k2), r = power of two bringing the bound 4*amax*kmax to `target`; *s_out = *s1 * r.
 * Optional second BN on the same dz (y2/coef2/kmax2 -> dy2, s_out2: the downsample branch) and optional dz output
 * (fp16 at scale *s1: the identity branch). */
int mauv_bn_bwd_apply(const void* d1, const void* d2, const float* s1, const float* s2, const void* relu_out, const void* y,
                      const void* y2, const float* coef, const float* coef2, const unsigned int* amax,
                      const unsigned int* kmax, const unsigned int* kmax2, float target, int G, long long M, int C, void* dy,
                      void* dy2, void* dz, float* s_out, float* s_out2, void* stream);
/* backward of maxpool3x3/2(relu(y*scale+shift)) (torchvision resnet.py stem): dz [G*imgs][H][W][C] at scale *s1.
 * idx_ws: G*imgs*Ho*Wo*C bytes (per-window arg-max positions, written by the first of the two launches). */
int mauv_maxpool_bwd_f16(const void* y, const float* scale_shift, const void* d1, const void* d2, const float* s1,
                         const float* s2, int G, int imgs_per_sample, int H, int W, int C, void* idx_ws, void* dz,
                         void* stream);
/* backward of the global average pool: dfeat [N][C] fp32 -> out [N][HW][C] fp16 = dfeat/HW * r, *s_out = r (first scale). */
int mauv_avgpool_bwd_f16(const float* dfeat, long long N, int HW, int C, float target, unsigned int* amax_ws, void* out,
                         float* s_out, void* stream);
/* Fused tail of a bottleneck with a downsample branch (torchvision resnet.py Bottleneck.forward: out = relu(bn3(conv3(.)) +
 * downsample(x))): ONE contraction over K-concatenated operands [a1 | a2] * [s3*W3 | sd*Wd]^T + (t3 + td). The BN scales are
 * folded into the sampled weights (mauv_sample_weights_scaled_f16), the shifts into the epilogue constants
 * (mauv_bn_shift_sum); the statistics come from two mauv_gemm_bn_f16 mode-1 passes. a1 [G][M][K1], a2 [G][M][K2] (a strided
 * downsample input is made dense with mauv_subsample_f16), w_cat [G][N][K1+K2], scale_shift [G][N][2] = (1, shift). */
int mauv_gemm_bn_cat_f16(const void* a1, int K1, const void* a2, int K2, const void* w_cat, void* y, const float* scale_shift,
                         int relu, int G, long long M, int N, void* stream);

/* ---- operand-transform variants: BatchNorm + ReLU of the PREVIOUS layer applied to the TMA-loaded operand tiles in shared memory
 * (4 transform warps between the TMA and the tcgen05 stage), so relu(bn2(conv2(.))) - the bottleneck's a2 - is never written to
 * or re-read from HBM (the reference materialises it: torchvision Bottleneck.forward `out = self.relu(self.bn2(out))`, reached
 * from models/base_models.py:74-76). Arithmetic of the transform = mauv_bn_act_f16's (fp32 fma, max, one fp16 rounding), so the
 * tensor core sees bit-identical operands.
 *   mauv_gram_bn_f16:        gram_partial [G*splits][K][K] = a^T a per pixel chunk, colsum_partial [G*splits][K] = column sums
 *                            of a per chunk, a = relu(y * scale + shift), y [G][M][K] raw, scale_shift [G][K][2]; K in {64,128,256},
 *                            (M / splits) % 64 == 0. Feeds mauv_bn_stats_from_gram (nblk = splits).
 *   mauv_gemm_bn_xf_f16:     mauv_gemm_bn_f16 mode 2 with A = relu(a_raw * s_a + t_a); K in {64,128,192,256}, N >= 128.
 *   mauv_gemm_bn_cat_xf_f16: mauv_gemm_bn_cat_f16 with the a1 k-blocks transformed the same way (a2 is read as is). */
int mauv_gram_bn_f16(const void* y, const float* scale_shift, float* gram_partial, float* colsum_partial, int G, int splits,
                     long long M, int K, void* stream);
int mauv_gemm_bn_xf_f16(const void* a_raw, const float* a_scale_shift, const void* w, void* y, const float* scale_shift,
                        const void* residual, int relu, int G, long long M, int N, int K, void* stream);
int mauv_gemm_bn_cat_xf_f16(const void* a1_raw, const float* a1_scale_shift, int K1, const void* a2, int K2, const void* w_cat,
                            void* y, const float* scale_shift, int relu, int G, long long M, int N, void* stream);
int mauv_sample_weights_scaled_f16(const float* mu, const float* rho, const float* eps, uint64_t seed, uint32_t layer_id,
                                   uint32_t sample0, int G, int cout, int cin, const float* scale_shift, int row_pitch,
                                   int col0, void* w_out, void* stream);
int mauv_bn_shift_sum(const float* scale_shift_a, const float* scale_shift_b, long long n, float* out, void* stream);
int mauv_subsample_f16(const void* x, long long N, int H, int W, int C, int stride, void* out, void* stream);
/* 3x3 / stride 1 / pad 1 conv with Cin = Cout = 64 (ResNet layer1 conv2: torchvision resnet.py Bottleneck.conv2) in
 * "padded stream" mode: tiles are 128 consecutive positions of the zero-padded pixel stream, one TMA box per filter row,
 * horizontal taps as shifted shared-memory descriptors, the sample's 9 weight blocks resident in shared memory. Same
 * results as mauv_conv2d_im2col_f16; statistics partials are [G][mauv_conv3x3_c64_tiles()][64][2]. W <= 254. */
int mauv_conv3x3_c64_tiles(int imgs_per_sample, int H, int W);
/* in_scale_shift (nullable) [G][64][2]: x is the RAW output of the previous conv and relu(x*scale + shift) - that layer's
 * BatchNorm + ReLU - is applied to each tile in shared memory before the MMAs (the activated tensor never exists in HBM). */
int mauv_conv3x3_c64_f16(const void* x, const void* w, void* y, float* stats_partial, const float* in_scale_shift, int G,
                         int imgs_per_sample, int H, int W, void* stream);
/* mauv_gemm_f16 whose [N][K] operand is shared by groups of batches: y[g] = a[g] * w[g % w_batches]^T (the stem's weight
 * gradient: per-sample dY^T chunks against the chunks of the one im2col matrix all samples share). */
int mauv_gemm_wmod_f16(const void* a, const void* w, int w_batches, void* y, int out_f32, int G, long long M, int N, int K,
                       void* stream);   /* out_f32 != 0: y is float [G][M][N], the fp32 accumulator stored unrounded */
/* Weight gradient of the grouped conv straight from the NHWC tensors (no transposed copies): dy [G*imgs][Ho][Wo][Cout],
 * x [G*imgs][H][W][Cin] -> dw [G*splits][Cout][kh*kw*Cin] FP32 partial sums over pixel chunks (K order (r, s, c)): the
 * accumulator is stored unrounded (a correlated sum over thousands of pixels can exceed fp16's range). Both
 * operands enter the tcgen05 MMA MN-major from [64 pixels][64 channels] TMA boxes (tiled for 1x1/stride 1, im2col mode
 * otherwise). Cin % 64 == 0; pixels per chunk (imgs*Ho*Wo / splits) % 64 == 0. */
int mauv_wgrad_f16(const void* dy, const void* x, void* dw, int G, int splits, int imgs_per_sample, int H, int W, int Cin,
                   int Cout, int kh, int kw, int stride, int pad, void* stream);
/* dw_partial fp16 or (partial_f32 != 0) fp32 [G*splits][cout][k_pad] (value = dW_g * *scale / inv_alpha) -> grad_mu += sum_g dW_g,
 * grad_rho += sum_g dW_g * eps_g * sigmoid(rho); eps injected [G][n] or Philox(seed, layer_id, sample0+g).
 * stale_eps != 0 reproduces the reference's saved-eps-buffer behaviour (every pass sees the last pass's eps). */
int mauv_wgrad_finalize_group(const void* dw_partial, int partial_f32, int G, int splits, int cout, int cin, int kh, int kw, int k_pad,
                              float inv_alpha, const float* scale, const float* rho, const float* eps, uint64_t seed,
                              uint32_t layer_id, uint32_t sample0, int stale_eps, float* grad_mu, float* grad_rho,
                              void* stream);
/* fusion-head linear backward over G samples (fp32, sampling fused): gx[g] (nullable; = or +=) = gy[g] * W_g;
 * parameter grads += sum over g. Strides in elements: *_sg per sample, *_sb per row. */
int mauv_sampled_linear_bwd_group_f32(const float* x, long long x_sg, int x_sb, const float* gy, long long gy_sg, int gy_sb,
                                      const float* mu_w, const float* rho_w, const float* eps_w, const float* rho_b,
                                      const float* eps_b, uint64_t seed, uint32_t layer_id, uint32_t sample0, int G, int B,
                                      int in_features, int out_features, int stale_eps, float* gx, long long gx_sg, int gx_sb,
                                      int accumulate_gx, float* grad_mu_w, float* grad_rho_w, float* grad_mu_b,
                                      float* grad_rho_b, void* stream);
/* AdditiveAttention backward pieces (models/base_models.py:43-52): d(q+k) = dt*(1-t^2); out = v*softmax(score). */
int mauv_tanh_bwd_f32(const float* t, const float* dt, long long n, float* out, void* stream);
int mauv_softmax_gate_bwd_f32(const float* score, const float* v, const float* dout, int ld_dout, long long rows, int n,
                              float* dscore, float* dv, void* stream);
/* loss = cross_entropy(mean_s logits, labels) (train/multimodal.py:118-121); dlogits [S][B][C] = d loss / d logits.
 * Labels as torch's cross_entropy: -100 rows are ignored (not in the mean, zero gradient); any other label outside
 * [0, C) makes the loss NaN (torch: device assert) and leaves that row's gradient zero - never read out of bounds. */
int mauv_ce_mean_fwd_bwd_f32(const float* logits, const long long* labels, int S, int B, int C, float* mean_logit,
                             float* dlogits, float* loss, void* stream);

/* ---- f1: fused Adam + finite-gradient guard over flat buffers ------------------------------------------------------
 * Replaces the reference's NaN/Inf guard loop + optimizer.step() (train/multimodal.py:141-145; the optimizers are
 * `optim.Adam(model.parameters(), lr=...)`, train/loop_utils.py:46-52). p, g, m, v: n contiguous fp32 each (parameters,
 * gradients, exp_avg, exp_avg_sq); state: mauv_adam_state_bytes() device bytes, zeroed once by the caller, layout
 * {int step, int applied, int scratch, int pad, float step_size, float inv_sqrt_bc2, pad[2]}. If any gradient is NaN/Inf
 * nothing is updated, `applied` = 0 and the step count does not advance - decided on the device, no host sync. Arithmetic
 * as torch.optim.Adam (amsgrad=False, maximize=False; weight_decay = L2 added to the gradient). */
int mauv_adam_state_bytes(void);
int mauv_adam_step_f32(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, void* state, void* stream);

/* ---- measurement aid (not on the product path): HBM write probes, see csrc/membench.cu ---------------------------
 * mode 0 st.global.v4, 1 st.global.cs.v4, 2 cp.async.bulk shared->global (the engine TMA stores use). */
int mauv_membench_fill(void* dst, long long bytes, unsigned int value, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAUV_B200_H */
