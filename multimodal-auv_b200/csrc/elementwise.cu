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
// CT / KWT > 0: channel count / filter width known at compile time (the reference's stems: 7x7 over 3 or 1 channels), so the
// k -> (r, s, c) decomposition is multiply-shift arithmetic instead of two integer divisions per element (the kernel is
// instruction-bound: ~200 instructions per 16 output bytes in the generic form).
template <int CT, int KWT>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ x, int B, int C_rt, int H, int W, int kh, int kw_rt,
                   int stride, int pad, int Ho, int Wo, int k_pad, __half* __restrict__ out) {
  const int C = CT > 0 ? CT : C_rt;
  const int kw = KWT > 0 ? KWT : kw_rt;
  // one thread per (row, 8-wide k chunk)
  const int chunks = k_pad / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long rows = static_cast<long long>(B) * Ho * Wo;
  if (idx >= rows * chunks) return;
  const long long row = idx / chunks;
  const int ch = static_cast<int>(idx - row * chunks);
  const unsigned row32 = static_cast<unsigned>(row);             // B*Ho*Wo < 2^31 (checked on the host)
  const unsigned pq = row32 / static_cast<unsigned>(Wo);
  const int q = static_cast<int>(row32 - pq * Wo);
  const int b = static_cast<int>(pq / static_cast<unsigned>(Ho));
  const int p = static_cast<int>(pq - static_cast<unsigned>(b) * Ho);
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

// Row-tiled variant for the reference's stems (7x7 / stride 2 / pad 3, C = 1 or 3): one block per output row (b, p). The 7
// input rows it needs are staged ONCE in shared memory as fp16 ([c][r][W + 6], zero padding materialised), then thread
// (pixel q, chunk ch) assembles its 16 output bytes from 8 two-byte shared-memory reads at offsets that depend only on ch
// (kept in registers) and writes them with one 16-byte store - the block's output, Wo rows of k_pad halves, is one contiguous
// range. ~25 instructions per 16 output bytes instead of ~200 (the gather version is instruction-bound at 1.1 TB/s; the matrix
// is rebuilt for every batch on EVERY rank, so at 8 GPUs it is 5 % of the step).
template <int C>
__global__ void __launch_bounds__(256)
stem_im2col_rows_kernel(const float* __restrict__ x, int H, int W, int Ho, int Wo, int k_pad, __half* __restrict__ out) {
  extern __shared__ __half srows[];                   // [C][7][W + 6]
  const int p = blockIdx.x, b = blockIdx.y;
  const int Wp = W + 6;
  for (int i = threadIdx.x; i < C * 7 * Wp; i += blockDim.x) {
    const int w = i % Wp - 3;
    const int cr = i / Wp;
    const int r = cr % 7, c = cr / 7;
    const int h = 2 * p - 3 + r;
    float v = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(x + ((static_cast<long long>(b) * C + c) * H + h) * W + w);
    srows[i] = __float2half_rn(v);
  }
  __syncthreads();
  const int chunks = k_pad / 8;
  const int per_pass = blockDim.x / chunks;           // pixels per pass; threads beyond per_pass * chunks idle
  const int ch = threadIdx.x % chunks, dq = threadIdx.x / chunks;
  if (dq >= per_pass) return;
  int off[8];                                         // element offsets of k = 8 ch + j in srows (for q = 0), -1 = padding column
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = ch * 8 + j;
    const int c = k % C, rs = k / C;
    const int s2 = rs % 7, r = rs / 7;
    off[j] = (k < 49 * C) ? (c * 7 + r) * Wp + s2 : -1;
  }
  __half* orow = out + (static_cast<long long>(b) * Ho + p) * Wo * k_pad;
  for (int q = dq; q < Wo; q += per_pass) {
    __align__(16) __half v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = off[j] >= 0 ? srows[off[j] + 2 * q] : __float2half_rn(0.f);
    *reinterpret_cast<uint4*>(orow + static_cast<long long>(q) * k_pad + ch * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// ---------------------------------------------------------------------------
// BN finalize: reduce the per-tile (sum, sumsq) partials written by the conv epilogue,
// produce per-(sample, channel) scale/shift, and replay the running-stat updates of
// the S sequential reference passes (momentum 0.1, unbiased variance).
// ---------------------------------------------------------------------------
// Stage 1: [G][m_tiles][C] float2 partials -> [G][splits][C] double2. Block = 32 channels x 8 tile lanes, so a
// warp reads 256 contiguous bytes per tile row; enough blocks (C/32 x splits x G) to pull the partials at HBM speed.
__global__ void __launch_bounds__(256)
bn_partial_reduce_kernel(const float2* __restrict__ partial, int m_tiles, int C, int splits,
                         double2* __restrict__ out /*[G][splits][C]*/) {
  __shared__ double2 red[8][32];
  const int cx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int sp = blockIdx.y, g = blockIdx.z;
  const int per = (m_tiles + splits - 1) / splits;
  const int t0 = sp * per;
  const int t1 = min(m_tiles, t0 + per);
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    const float2* pp = partial + static_cast<long long>(g) * m_tiles * C + c;
    int t = t0 + ty;
    for (; t + 24 < t1; t += 32) {   // 4 independent loads in flight
      const float2 a = __ldcs(pp + static_cast<long long>(t) * C);
      const float2 b = __ldcs(pp + static_cast<long long>(t + 8) * C);
      const float2 d = __ldcs(pp + static_cast<long long>(t + 16) * C);
      const float2 e = __ldcs(pp + static_cast<long long>(t + 24) * C);
      s1 += (static_cast<double>(a.x) + b.x) + (static_cast<double>(d.x) + e.x);
      s2 += (static_cast<double>(a.y) + b.y) + (static_cast<double>(d.y) + e.y);
    }
    for (; t < t1; t += 8) {
      const float2 a = __ldcs(pp + static_cast<long long>(t) * C);
      s1 += a.x; s2 += a.y;
    }
  }
  red[ty][cx] = make_double2(s1, s2);
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int j = 1; j < 8; ++j) { s1 += red[j][cx].x; s2 += red[j][cx].y; }
    out[(static_cast<long long>(g) * splits + sp) * C + c] = make_double2(s1, s2);
  }
}

// Stage 2: per-(sample, channel) scale/shift in parallel over (32 channels x samples), then the G sequential
// running-stat updates (momentum, unbiased variance) and the num_batches_tracked increment of G reference passes.
constexpr int BN_MAX_G = 64;
__global__ void __launch_bounds__(512)
bn_finalize_kernel(const double2* __restrict__ partial, int G, int splits, int C, long long count,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                   long long* __restrict__ num_batches_tracked, float2* __restrict__ scale_shift,
                   float2* __restrict__ batch_stats) {
  __shared__ float2 mv[BN_MAX_G][32];   // (mean, unbiased var) per sample for the running-stat replay
  const int cx = threadIdx.x, gy = threadIdx.y;
  const int c = blockIdx.x * 32 + cx;
  if (c < C) {
    const float ga = gamma ? gamma[c] : 1.f;
    const float be = beta ? beta[c] : 0.f;
    for (int g = gy; g < G; g += blockDim.y) {
      double s1 = 0.0, s2 = 0.0;
      const double2* pp = partial + static_cast<long long>(g) * splits * C + c;
      for (int t = 0; t < splits; ++t) {
        const double2 v = pp[static_cast<long long>(t) * C];
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
      mv[g][cx] = make_float2(static_cast<float>(mean), static_cast<float>(unbiased));
    }
  }
  __syncthreads();
  if (gy == 0 && c < C && running_mean && running_var) {
    float rm = running_mean[c], rv = running_var[c];
    for (int g = 0; g < G; ++g) {
      rm = (1.f - momentum) * rm + momentum * mv[g][cx].x;
      rv = (1.f - momentum) * rv + momentum * mv[g][cx].y;
    }
    running_mean[c] = rm;
    running_var[c] = rv;
  }
  if (num_batches_tracked && blockIdx.x == 0 && cx == 0 && gy == 0) *num_batches_tracked += G;
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

// One thread owns one 8-channel group (c0 is loop invariant because the thread stride is a multiple of C/8),
// keeps its 8 (scale, shift) pairs in registers and streams rows with 4 independent 16-byte loads in flight.
// colsum (nullable) [G][gridDim.x][C]: per-block column sums of the fp16 OUTPUT values (what the next conv reads) - the
// first moment of the closed-form BatchNorm statistics of that conv (mauv_bn_stats_from_gram).
template <bool HAS_Y2, bool HAS_RES>
__global__ void __launch_bounds__(256)
bn_act_kernel(const uint4* __restrict__ y, const float2* __restrict__ ss, const uint4* __restrict__ res,
              const uint4* __restrict__ y2, const float2* __restrict__ ss2, int relu,
              long long per_sample_vec /* M*C/8 */, int C, uint4* __restrict__ out, float* __restrict__ colsum) {
  const int g = blockIdx.y;
  const unsigned cvec = C >> 3;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;   // multiple of 256 >= cvec (power of 2)
  const int c0 = static_cast<int>(static_cast<unsigned>(tid) % cvec) * 8;
  float sc[8], sh[8], sc2[8], sh2[8];
  {
    const float4* s4 = reinterpret_cast<const float4*>(ss + static_cast<long long>(g) * C + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = __ldg(s4 + j);
      sc[2 * j] = t.x; sh[2 * j] = t.y; sc[2 * j + 1] = t.z; sh[2 * j + 1] = t.w;
    }
    if (HAS_Y2) {
      const float4* t4 = reinterpret_cast<const float4*>(ss2 + static_cast<long long>(g) * C + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = __ldg(t4 + j);
        sc2[2 * j] = t.x; sh2[2 * j] = t.y; sc2[2 * j + 1] = t.z; sh2[2 * j + 1] = t.w;
      }
    }
  }
  const long long base = static_cast<long long>(g) * per_sample_vec;
  y += base; out += base;
  if (HAS_Y2) y2 += base;
  if (HAS_RES) res += base;
  constexpr int U = 4;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i0 = tid; i0 < per_sample_vec; i0 += stride * U) {
    uint4 vy[U], v2[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < per_sample_vec) {
        vy[u] = __ldcs(y + i);
        if (HAS_Y2) v2[u] = __ldcs(y2 + i);
        if (HAS_RES) vr[u] = __ldcs(res + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < per_sample_vec) {
        float f[8];
        unpack8(vy[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
        if (HAS_Y2) {
          float f2[8];
          unpack8(v2[u], f2);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += fmaf(f2[j], sc2[j], sh2[j]);
        }
        if (HAS_RES) {
          float fr[8];
          unpack8(vr[u], fr);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += fr[j];
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        const uint4 o = pack8(f);
        out[i] = o;
        if (colsum) {
          float fo[8];
          unpack8(o, fo);
#pragma unroll
          for (int j = 0; j < 8; ++j) cs[j] += fo[j];
        }
      }
    }
  }
  if (colsum) {      // threads t, t + cvec, t + 2 cvec .. of the block own the same 8 channels (cvec <= 256 divides 256)
    __shared__ float red[256][9];
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = cs[j];
    __syncthreads();
    if (threadIdx.x < cvec) {
      for (unsigned t = threadIdx.x + cvec; t < 256; t += cvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[j] += red[t][j];
      }
      float* dst = colsum + (static_cast<long long>(g) * gridDim.x + blockIdx.x) * C + threadIdx.x * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(cs[0], cs[1], cs[2], cs[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(cs[4], cs[5], cs[6], cs[7]);
    }
  }
}

// Column sums of a [G][M][C] fp16 tensor: colsum [G][gridDim.x][C] per-block partials (same layout as bn_act's).
__global__ void __launch_bounds__(256)
colsum_kernel(const uint4* __restrict__ x, long long per_sample_vec, int C, float* __restrict__ colsum) {
  const int g = blockIdx.y;
  const unsigned cvec = C >> 3;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  x += static_cast<long long>(g) * per_sample_vec;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i0 = tid; i0 < per_sample_vec; i0 += stride * 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * stride < per_sample_vec) v[u] = __ldcs(x + i0 + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * stride < per_sample_vec) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[j] += f[j];
      }
  }
  __shared__ float red[256][9];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = cs[j];
  __syncthreads();
  if (threadIdx.x < cvec) {
    for (unsigned t = threadIdx.x + cvec; t < 256; t += cvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[j] += red[t][j];
    }
    float* dst = colsum + (static_cast<long long>(g) * gridDim.x + blockIdx.x) * C + threadIdx.x * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(cs[0], cs[1], cs[2], cs[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(cs[4], cs[5], cs[6], cs[7]);
  }
}

// ---------------------------------------------------------------------------
// Closed-form BatchNorm batch statistics of a 1x1 convolution y = a W^T from the moments of its INPUT:
//     sum_m y[m][n]   = w_n . s            s = column sums of a        (K values per sample)
//     sum_m y[m][n]^2 = w_n^T S w_n        S = a^T a (second moments)   (K x K per sample)
// so the statistics pass of the recompute scheme (a full N x K contraction over all pixels whose M x N result is only
// reduced) becomes one K x K contraction over the pixels (mauv_wgrad_f16 with dy = x = a: MN-major tcgen05 operands,
// fp32 partial sums per pixel chunk) plus this small evaluation. Stage A reduces the partials (double accumulation),
// stage B evaluates the two forms for 32 output channels per block with S staged through shared memory, and writes
// (sum, sum of squares) in the double2 layout bn_finalize_kernel consumes.
// ---------------------------------------------------------------------------
// Stage A. Element e < K*K is an entry of S (partials `splits` apart by K*K), element K*K + c is column sum c (partials
// `nblk` apart by K). block = 32 consecutive elements x LY partial lanes: every lane sums its partials p = ty, ty + LY, .. in
// double with 8 independent loads in flight, the lanes are combined in a fixed order (deterministic). LY = 1 for short
// partial lists (K = 256: 16 pixel chunks) - no shared memory, no barrier -, LY = 8 for long ones (K = 64: 256 chunks).
template <int LY>
__global__ void __launch_bounds__(32 * LY)
gram_reduce_kernel(const float* __restrict__ gram, int splits, long long kk /* K*K */, const float* __restrict__ colsum,
                   int nblk, int K, float* __restrict__ S /*[G][K*K]*/, float* __restrict__ s1 /*[G][K]*/) {
  __shared__ double red[LY][33];
  const int g = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long e = static_cast<long long>(blockIdx.x) * 32 + tx;
  const float* src = nullptr;
  long long pitch = 0;
  int n = 0;
  if (e < kk) {
    src = gram + static_cast<long long>(g) * splits * kk + e; pitch = kk; n = splits;
  } else if (e < kk + K) {
    src = colsum + static_cast<long long>(g) * nblk * K + (e - kk); pitch = K; n = nblk;
  }
  double acc = 0.0;
  int sp = ty;
  for (; sp + 7 * LY < n; sp += 8 * LY) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(src + static_cast<long long>(sp + u * LY) * pitch);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += static_cast<double>(v[u]);
  }
  for (; sp < n; sp += LY) acc += static_cast<double>(__ldcs(src + static_cast<long long>(sp) * pitch));
  if (LY > 1) {
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0) {
#pragma unroll
      for (int j = 1; j < LY; ++j) acc += red[j][tx];
    }
  }
  if (ty == 0 && e < kk + K) {
    if (e < kk) S[static_cast<long long>(g) * kk + e] = static_cast<float>(acc);
    else s1[static_cast<long long>(g) * K + (e - kk)] = static_cast<float>(acc);
  }
}

// Stage B. q_n = w_n^T S w_n and m_n = w_n . s1 for a block of GQ_N = 64 output channels and a tile of GQ_ROWS = 64 rows of S:
// T[row][n] = sum_c S[c][row] w[n][c] (S is symmetric: its rows are read where they lie, no transpose) as a register-tiled fp32
// contraction - thread (ty, tx) owns rows 4 ty .. 4 ty + 3 x channels 4 tx .. 4 tx + 3, two LDS.128 per 16 FMAs - over
// 64-column chunks staged in shared memory; then q_n += sum_rows w[n][row] T[row][n] in double, combined over the 16 row
// groups in a fixed order. The K/64 row tiles of one (sample, channel block) are separate blocks (blockIdx.z) and
// bn_finalize_kernel adds their partials. (The first version walked one row per lane with three shared-memory reads per
// 8 FMAs: 236 us for K = 256, N = 1024, G = 15 - more than the second-moment contraction it evaluates.)
constexpr int GQ_N = 64;
constexpr int GQ_ROWS = 64;
constexpr int GQ_WPITCH = 68;   // floats per staged weight row: 16-byte aligned float4 reads, 4-way conflicts on the transposing writes
__global__ void __launch_bounds__(256)
gram_quadform_kernel(const float* __restrict__ S, const float* __restrict__ s1, const __half* __restrict__ w /*[G][N][K]*/,
                     int N, int K, double2* __restrict__ out /*[G][gridDim.z][N] (sum, sum of squares) partials per row tile*/) {
  __shared__ __align__(16) float wt[64 * GQ_WPITCH];      // [c][channel]  this chunk's weights, transposed
  __shared__ __align__(16) float St[64 * GQ_ROWS];        // [c][row]      S[c0 + c][r0 + row]
  __shared__ double red[16][GQ_N];
  const int g = blockIdx.y, n0 = blockIdx.x * GQ_N, r0 = blockIdx.z * GQ_ROWS;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* Sg = S + static_cast<long long>(g) * K * K;
  const __half* wg = w + static_cast<long long>(g) * N * K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  double m1 = 0.0;                                         // partial first moment of channel n0 + (tid & 63) (row tile 0)
  for (int c0 = 0; c0 < K; c0 += 64) {
    __syncthreads();
    for (int i = tid; i < 64 * 64; i += 256) {
      const int nn = i >> 6, c = i & 63;                   // coalesced along c
      const int n = n0 + nn;
      wt[c * GQ_WPITCH + nn] = (n < N && c0 + c < K) ? __half2float(wg[static_cast<long long>(n) * K + c0 + c]) : 0.f;
    }
    for (int i = tid; i < 64 * GQ_ROWS; i += 256) {
      const int c = i >> 6, r = i & 63;                    // coalesced along the row of S
      St[i] = (c0 + c < K && r0 + r < K) ? Sg[static_cast<long long>(c0 + c) * K + r0 + r] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < 64; ++c) {
      const float4 sv = *reinterpret_cast<const float4*>(St + c * GQ_ROWS + 4 * ty);
      const float4 wv = *reinterpret_cast<const float4*>(wt + c * GQ_WPITCH + 4 * tx);
      const float sr[4] = {sv.x, sv.y, sv.z, sv.w}, wc[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(sr[i], wc[j], acc[i][j]);
    }
    if (blockIdx.z == 0) {       // first moment: every thread takes 16 of the chunk's 64 columns for channel tid & 63
      const int nn = tid & 63, part = tid >> 6;
#pragma unroll 4
      for (int c = part * 16; c < part * 16 + 16; ++c)
        if (c0 + c < K) m1 += static_cast<double>(s1[static_cast<long long>(g) * K + c0 + c]) * static_cast<double>(wt[c * GQ_WPITCH + nn]);
    }
  }
  // q_n partial of this thread: its 4 rows
  double q[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + 4 * tx + j;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + 4 * ty + i;
      if (n < N && r < K) q[j] += static_cast<double>(__half2float(wg[static_cast<long long>(n) * K + r])) * static_cast<double>(acc[i][j]);
    }
  }
  __shared__ double red1[4][GQ_N];
#pragma unroll
  for (int j = 0; j < 4; ++j) red[ty][4 * tx + j] = q[j];
  red1[tid >> 6][tid & 63] = m1;
  __syncthreads();
  if (tid < GQ_N) {
    double qs = 0.0;
#pragma unroll
    for (int r = 0; r < 16; ++r) qs += red[r][tid];
    const double ms = (red1[0][tid] + red1[1][tid]) + (red1[2][tid] + red1[3][tid]);
    const int n = n0 + tid;
    if (n < N) out[(static_cast<long long>(g) * gridDim.z + blockIdx.z) * N + n] = make_double2(ms, qs);
  }
}


// Stem tail: BN + ReLU + 3x3/2 max-pool (pad 1) in one pass.
// y: [G*B][H][W][C] fp16 raw conv output -> out: [G*B][Ho][Wo][C]. One block per (image, output row): no integer
// division in the hot loop, the three input rows are read with 9 independent 16-byte loads per thread.
__global__ void __launch_bounds__(128, 5)
bn_relu_maxpool_kernel(const uint4* __restrict__ y, const float2* __restrict__ ss, int imgs_per_sample,
                       int H, int W, int C, int Ho, int Wo, uint4* __restrict__ out) {
  const int cvec = C >> 3;
  const int p = blockIdx.x;                  // output row
  const long long n = blockIdx.y;            // image index in [0, G*B)
  const int g = static_cast<int>(n / imgs_per_sample);
  const int per_row = Wo * cvec;
  for (int t = threadIdx.x; t < per_row; t += blockDim.x) {
    const int cv = t % cvec;
    const int q = t / cvec;
    float sc[8], sh[8], m[8];
    const float4* s4 = reinterpret_cast<const float4*>(ss + static_cast<long long>(g) * C + cv * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 tt = __ldg(s4 + j);
      sc[2 * j] = tt.x; sh[2 * j] = tt.y; sc[2 * j + 1] = tt.z; sh[2 * j + 1] = tt.w;
    }
    uint4 v[9];
    bool ok[9];
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
#pragma unroll
      for (int ds = 0; ds < 3; ++ds) {
        const int h = p * 2 - 1 + dr, w = q * 2 - 1 + ds;
        ok[dr * 3 + ds] = (h >= 0 && h < H && w >= 0 && w < W);
        if (ok[dr * 3 + ds]) v[dr * 3 + ds] = __ldg(y + ((n * H + h) * W + w) * cvec + cv);
      }
    }
    // x -> x*scale + shift is monotone (increasing for scale >= 0, decreasing otherwise, rounding included), so the window
    // maximum of the affine values is the affine value of the window max (or min) of the RAW fp16 inputs: 9 x 8 packed
    // half2 min/max instead of 9 x 8 conversions + FMAs + fp32 max. The centre tap (index 4) is always inside the image.
    __half2 mx[4], mn[4];
    {
      const __half2* c2 = reinterpret_cast<const __half2*>(&v[4]);
#pragma unroll
      for (int j = 0; j < 4; ++j) { mx[j] = c2[j]; mn[j] = c2[j]; }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (k != 4 && ok[k]) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { mx[j] = __hmax2(mx[j], h2[j]); mn[j] = __hmin2(mn[j], h2[j]); }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(mx[j]), b = __half22float2(mn[j]);
      m[2 * j] = fmaxf(fmaf(sc[2 * j] >= 0.f ? a.x : b.x, sc[2 * j], sh[2 * j]), 0.f);              // relu(max) == max(relu)
      m[2 * j + 1] = fmaxf(fmaf(sc[2 * j + 1] >= 0.f ? a.y : b.y, sc[2 * j + 1], sh[2 * j + 1]), 0.f);
    }
    out[((n * Ho + p) * Wo + q) * cvec + cv] = pack8(m);
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


// (scale, shift) pairs of two BatchNorms that are summed after normalisation -> (1, shift_a + shift_b): the epilogue
// constants of the fused bottleneck tail whose scales were folded into the weights (mauv_gemm_bn_cat_f16).
__global__ void __launch_bounds__(256)
bn_shift_sum_kernel(const float2* __restrict__ a, const float2* __restrict__ b, long long n, float2* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float2(1.f, a[i].y + b[i].y);
}

// x [N][H][W][C] -> out [N][Ho][Wo][C], out(p, q) = x(p*stride, q*stride): the input of a strided 1x1 conv (downsample
// branch) as a dense matrix, so that the statistics pass and the fused pass can read it with tiled TMA.
__global__ void __launch_bounds__(256)
subsample_kernel(const uint4* __restrict__ x, int H, int W, int cvec, int Ho, int Wo, int stride, long long total,
                 uint4* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = static_cast<int>(i % cvec);
  long long t = i / cvec;
  const int q = static_cast<int>(t % Wo); t /= Wo;
  const int p = static_cast<int>(t % Ho);
  const long long n = t / Ho;
  out[i] = __ldg(x + ((n * H + static_cast<long long>(p) * stride) * W + static_cast<long long>(q) * stride) * cvec + cv);
}

}  // namespace

extern "C" {

int mauv_stem_im2col_f16(const float* x_nchw, int B, int C, int H, int W, int kh, int kw, int stride,
                         int pad, int k_pad, void* out, void* stream) {
  MAUV_CHECK_ARG(x_nchw && out, "mauv_stem_im2col_f16: null pointer");
  MAUV_CHECK_ARG(k_pad % 8 == 0 && k_pad >= kh * kw * C, "mauv_stem_im2col_f16: bad k_pad=%d", k_pad);
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  MAUV_CHECK_ARG(static_cast<long long>(B) * Ho * Wo < (1LL << 31), "mauv_stem_im2col_f16: too many output pixels");
  const long long work = static_cast<long long>(B) * Ho * Wo * (k_pad / 8);
  const unsigned grid = static_cast<unsigned>(ceil_div_i64(work, 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* o = static_cast<__half*>(out);
  if (kh == 7 && kw == 7 && stride == 2 && pad == 3 && (C == 3 || C == 1) && k_pad / 8 <= 256 && W <= 1024) {
    const int smem = C * 7 * (W + 6) * static_cast<int>(sizeof(__half));
    dim3 rgrid(Ho, B);
    if (C == 3) stem_im2col_rows_kernel<3><<<rgrid, 256, smem, st>>>(x_nchw, H, W, Ho, Wo, k_pad, o);
    else stem_im2col_rows_kernel<1><<<rgrid, 256, smem, st>>>(x_nchw, H, W, Ho, Wo, k_pad, o);
    MAUV_LAUNCH_CHECK("stem_im2col_rows_kernel");
    return MAUV_OK;
  }
  if (kw == 7 && C == 3) stem_im2col_kernel<3, 7><<<grid, 256, 0, st>>>(x_nchw, B, C, H, W, kh, kw, stride, pad, Ho, Wo, k_pad, o);
  else if (kw == 7 && C == 1) stem_im2col_kernel<1, 7><<<grid, 256, 0, st>>>(x_nchw, B, C, H, W, kh, kw, stride, pad, Ho, Wo, k_pad, o);
  else stem_im2col_kernel<0, 0><<<grid, 256, 0, st>>>(x_nchw, B, C, H, W, kh, kw, stride, pad, Ho, Wo, k_pad, o);
  MAUV_LAUNCH_CHECK("stem_im2col_kernel");
  return MAUV_OK;
}

static int bn_splits(int m_tiles) {
  int sp = (m_tiles + 31) / 32;
  return sp < 1 ? 1 : (sp > 64 ? 64 : sp);
}
// workspace: double2[G][splits][C]
long long mauv_bn_finalize_ws_bytes(int G, int m_tiles, int C) {
  return static_cast<long long>(G) * bn_splits(m_tiles) * C * sizeof(double2);
}

int mauv_bn_finalize(const float* stats_partial, int G, int m_tiles, int C, long long count,
                     const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var, long long* num_batches_tracked,
                     float* scale_shift, float* batch_stats, void* ws, void* stream) {
  MAUV_CHECK_ARG(stats_partial && scale_shift && ws && G >= 1 && m_tiles >= 1 && C >= 1 && count >= 1,
                 "mauv_bn_finalize: bad argument");
  MAUV_CHECK_ARG(G <= BN_MAX_G, "mauv_bn_finalize: at most %d samples per call (got %d)", BN_MAX_G, G);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int splits = bn_splits(m_tiles);
  dim3 grid((C + 31) / 32, splits, G);
  bn_partial_reduce_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(stats_partial), m_tiles, C, splits,
                                                 static_cast<double2*>(ws));
  MAUV_LAUNCH_CHECK("bn_partial_reduce_kernel");
  dim3 fblock(32, G < 16 ? G : 16);
  bn_finalize_kernel<<<(C + 31) / 32, fblock, 0, st>>>(static_cast<const double2*>(ws), G, splits, C, count, gamma, beta,
                                                       eps, momentum, running_mean, running_var, num_batches_tracked,
                                                       reinterpret_cast<float2*>(scale_shift),
                                                       reinterpret_cast<float2*>(batch_stats));
  MAUV_LAUNCH_CHECK("bn_finalize_kernel");
  return MAUV_OK;
}

int mauv_bn_act_blocks(int G, long long M, int C) {
  const long long per = M * C / 8;
  long long bx = ceil_div_i64(per, 256 * 4);
  const long long cap = static_cast<long long>(mauv_num_sms()) * 16 / (G < 16 ? G : 16) + 1;
  if (bx > cap) bx = cap;
  return static_cast<int>(bx < 1 ? 1 : bx);
}

int mauv_bn_act_f16(const void* y, const float* scale_shift, const void* residual, const void* y2,
                    const float* scale_shift2, int relu, int G, long long M, int C, void* out, float* colsum_partial,
                    void* stream) {
  MAUV_CHECK_ARG(y && scale_shift && out, "mauv_bn_act_f16: null pointer");
  MAUV_CHECK_ARG(C % 8 == 0, "mauv_bn_act_f16: C must be a multiple of 8");
  MAUV_CHECK_ARG((y2 == nullptr) == (scale_shift2 == nullptr), "mauv_bn_act_f16: y2 and scale_shift2 go together");
  MAUV_CHECK_ARG((C & (C - 1)) == 0 && C <= 2048, "mauv_bn_act_f16: C must be a power of two <= 2048 (got %d)", C);
  const long long per = M * C / 8;
  MAUV_CHECK_ARG(!colsum_partial || C >= 8, "mauv_bn_act_f16: column sums need C >= 8");
  dim3 grid(static_cast<unsigned>(mauv_bn_act_blocks(G, M, C)), G);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint4* py = static_cast<const uint4*>(y);
  const float2* pss = reinterpret_cast<const float2*>(scale_shift);
  const uint4* pres = static_cast<const uint4*>(residual);
  const uint4* py2 = static_cast<const uint4*>(y2);
  const float2* pss2 = reinterpret_cast<const float2*>(scale_shift2);
  uint4* po = static_cast<uint4*>(out);
  if (y2 && residual) bn_act_kernel<true, true><<<grid, 256, 0, st>>>(py, pss, pres, py2, pss2, relu, per, C, po, colsum_partial);
  else if (y2) bn_act_kernel<true, false><<<grid, 256, 0, st>>>(py, pss, pres, py2, pss2, relu, per, C, po, colsum_partial);
  else if (residual) bn_act_kernel<false, true><<<grid, 256, 0, st>>>(py, pss, pres, py2, pss2, relu, per, C, po, colsum_partial);
  else bn_act_kernel<false, false><<<grid, 256, 0, st>>>(py, pss, pres, py2, pss2, relu, per, C, po, colsum_partial);
  MAUV_LAUNCH_CHECK("bn_act_kernel");
  return MAUV_OK;
}

int mauv_bn_relu_maxpool_f16(const void* y, const float* scale_shift, int G, int imgs_per_sample, int H,
                             int W, int C, void* out, void* stream) {
  MAUV_CHECK_ARG(y && scale_shift && out && C % 8 == 0, "mauv_bn_relu_maxpool_f16: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long imgs = static_cast<long long>(G) * imgs_per_sample;
  MAUV_CHECK_ARG(imgs <= 65535, "mauv_bn_relu_maxpool_f16: at most 65535 images per call (got %lld)", imgs);
  // small blocks, several resident per SM: each block is one dependent load->max->store chain, so the memory
  // latency is hidden by block-level parallelism
  int threads = Wo * (C / 8);
  threads = threads > 128 ? 128 : ((threads + 31) / 32) * 32;
  dim3 grid(Ho, static_cast<unsigned>(imgs));
  bn_relu_maxpool_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(y), reinterpret_cast<const float2*>(scale_shift), imgs_per_sample, H, W, C,
      Ho, Wo, static_cast<uint4*>(out));
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

int mauv_bn_shift_sum(const float* scale_shift_a, const float* scale_shift_b, long long n, float* out, void* stream) {
  MAUV_CHECK_ARG(scale_shift_a && scale_shift_b && out && n >= 1, "mauv_bn_shift_sum: bad argument");
  bn_shift_sum_kernel<<<static_cast<unsigned>(ceil_div_i64(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(scale_shift_a), reinterpret_cast<const float2*>(scale_shift_b), n,
      reinterpret_cast<float2*>(out));
  MAUV_LAUNCH_CHECK("bn_shift_sum_kernel");
  return MAUV_OK;
}

int mauv_subsample_f16(const void* x, long long N, int H, int W, int C, int stride, void* out, void* stream) {
  MAUV_CHECK_ARG(x && out && C % 8 == 0 && stride >= 1, "mauv_subsample_f16: bad argument");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const long long total = N * Ho * Wo * (C / 8);
  subsample_kernel<<<static_cast<unsigned>(ceil_div_i64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), H, W, C / 8, Ho, Wo, stride, total, static_cast<uint4*>(out));
  MAUV_LAUNCH_CHECK("subsample_kernel");
  return MAUV_OK;
}

// Column sums of x [G][M][C] fp16 -> colsum_partial [G][mauv_bn_act_blocks(G, M, C)][C] fp32 (per-block partials).
int mauv_colsum_f16(const void* x, int G, long long M, int C, float* colsum_partial, void* stream) {
  MAUV_CHECK_ARG(x && colsum_partial && G >= 1 && M >= 1, "mauv_colsum_f16: bad argument");
  MAUV_CHECK_ARG(C >= 8 && (C & (C - 1)) == 0 && C <= 2048, "mauv_colsum_f16: C must be a power of two in [8, 2048] (got %d)", C);
  dim3 grid(static_cast<unsigned>(mauv_bn_act_blocks(G, M, C)), G);
  colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x), M * C / 8, C, colsum_partial);
  MAUV_LAUNCH_CHECK("colsum_kernel");
  return MAUV_OK;
}

// workspace of mauv_bn_stats_from_gram: S [G][K*K] + s1 [G][K] floats, then double2 [G][N]
long long mauv_bn_stats_from_gram_ws_bytes(int G, int N, int K) {
  const long long f = (static_cast<long long>(G) * K * K + static_cast<long long>(G) * K + 3) / 4 * 4;   // 16-byte aligned
  return f * 4 + static_cast<long long>(G) * ((K + GQ_ROWS - 1) / GQ_ROWS) * N * sizeof(double2);
}

int mauv_bn_stats_from_gram(const float* gram_partial, int splits, const float* colsum_partial, int nblk, const void* w,
                            int G, int N, int K, long long count, const float* gamma, const float* beta, float eps,
                            float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                            float* scale_shift, float* batch_stats, void* ws, void* stream) {
  MAUV_CHECK_ARG(gram_partial && colsum_partial && w && scale_shift && ws && splits >= 1 && nblk >= 1,
                 "mauv_bn_stats_from_gram: null pointer");
  MAUV_CHECK_ARG(G >= 1 && G <= BN_MAX_G && N >= 1 && K >= 8 && K <= 512 && K % 8 == 0 && count >= 1,
                 "mauv_bn_stats_from_gram: bad shape G=%d N=%d K=%d", G, N, K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* S = static_cast<float*>(ws);
  float* s1 = S + static_cast<long long>(G) * K * K;
  const long long f = (static_cast<long long>(G) * K * K + static_cast<long long>(G) * K + 3) / 4 * 4;
  double2* sums = reinterpret_cast<double2*>(static_cast<float*>(ws) + f);
  const long long kk = static_cast<long long>(K) * K;
  dim3 g1(static_cast<unsigned>(ceil_div_i64(kk + K, 32)), G);
  if (splits >= 64) gram_reduce_kernel<8><<<g1, 256, 0, st>>>(gram_partial, splits, kk, colsum_partial, nblk, K, S, s1);
  else gram_reduce_kernel<1><<<g1, 32, 0, st>>>(gram_partial, splits, kk, colsum_partial, nblk, K, S, s1);   // (256-thread blocks of 8 independent warps: 34 -> 83 us)
  MAUV_LAUNCH_CHECK("gram_reduce_kernel");
  const int n_rt = (K + GQ_ROWS - 1) / GQ_ROWS;
  dim3 g2((N + GQ_N - 1) / GQ_N, G, n_rt);
  gram_quadform_kernel<<<g2, 256, 0, st>>>(S, s1, static_cast<const __half*>(w), N, K, sums);
  MAUV_LAUNCH_CHECK("gram_quadform_kernel");
  dim3 fblock(32, G < 16 ? G : 16);
  bn_finalize_kernel<<<(N + 31) / 32, fblock, 0, st>>>(sums, G, n_rt, N, count, gamma, beta, eps, momentum, running_mean,
                                                       running_var, num_batches_tracked,
                                                       reinterpret_cast<float2*>(scale_shift),
                                                       reinterpret_cast<float2*>(batch_stats));
  MAUV_LAUNCH_CHECK("bn_finalize_kernel(gram)");
  return MAUV_OK;
}

}  // extern "C"
