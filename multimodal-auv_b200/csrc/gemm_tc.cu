// K1: grouped (per-MC-sample) implicit-GEMM convolution / plain GEMM on the
// Blackwell tensor cores.
//
//   Y[g][m][n] = sum_k A[g][m][k] * W[g][n][k]       fp16 in, fp32 accumulate in TMEM
//
// A is either a plain row-major [M,K] matrix (1x1/stride-1 convs, linear layers,
// the explicit im2col matrix of the 7x7 stem) fetched with tiled TMA, or an
// NHWC activation tensor fetched with *im2col-mode* TMA (3x3 and strided convs):
// the K loop then walks (filter tap r, s) x (64-channel block). W[g] is the g-th
// Monte-Carlo sample of the layer's weights (sample_weights.cu writes it in the
// same (r, s, c) K order). The reference does the same contraction with
// F.conv2d / F.linear per MC pass (bayesian-torch conv_variational.py forward;
// called from models/base_models.py:74-90 through torchvision resnet.py:143-165).
//
// Structure (one persistent CTA per SM, 12 warps):
//   warp 0      TMA producer   (one elected lane), kStages-deep smem ring
//   warp 1      MMA issuer     (one elected lane), tcgen05.mma 128 x BN x 16
//   warp 2      TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warps 4..11 epilogue (two sets of 4 TMEM lane quarters, alternating 64-column blocks):
//               tcgen05.ld -> fp16 -> swizzled smem -> TMA store, + per-channel sum / sum-of-squares
//               (the BatchNorm batch statistics of the reference's BN-train pass)
// Tiles are assigned round-robin (tile = blockIdx.x + i * gridDim.x) with the
// n-tile fastest so that CTAs running concurrently share their A tile in L2.
//
// Further modes of the same kernel (selected through GemmParams):
//   * EPI 1/3 statistics-only passes and EPI 2 fused BatchNorm(+residual)(+ReLU) epilogue: the recompute scheme of the
//     bottleneck tails; a2_kb > 0 K-concatenates A from two tensors (tail of a block with a downsample branch);
//   * mn = 1: weight gradient dW = dY^T X with BOTH operands MN-major, read where they lie ([64 px][64 ch] TMA boxes of
//     the row-major dY and of the NHWC activations; im2col-mode TMA on the B side for 3x3 / strided layers);
//   * split / a_wrap_kb / a_cwrap: the fp16x3 validation arithmetic (K-concatenated hi/lo operands).
// Compile-time instances (template parameters):
//   * PLAIN 1 / 2 / 3: every mode flag folded (hot inference shapes) / the stem's sample-stacked tiles / M-stacked tiles (two
//     128-row A tiles share one B tile: the N = 128 im2col convs);
//   * XF: BatchNorm + ReLU of the PREVIOUS layer applied to the TMA-loaded operand tiles in shared memory by transform warps
//     between the TMA and the MMA stage (fused tails and second moments read the raw conv2 output);
//   * NR / BR: fused-BN tails without a residual operand (third pipeline stage) / with a resident weight tile (one B load per
//     sample and CTA, only the A k-blocks stream).
// conv3x3_c64_stream_kernel (below) is the padded-stream sibling for layer1's 3x3 64->64 convs; stem_pool.cu holds the stem.
#include <cstdlib>
#include "common.cuh"
#include "tmap.cuh"

namespace {

struct GemmParams {
  int M;            // rows per sample
  int N;            // output channels
  int k_blocks;     // ceil(K / 64)
  int G;            // samples in this launch
  int m_tiles, n_tiles;
  int stat_tiles;   // leading dimension of the statistics buffer (128-row tiles per sample); == m_tiles except for M-stacked tiles
  long long total_tiles;
  int a_mode;       // 0 = tiled [K, M, G] ; 1 = im2col (C, W, H, N)
  int a_batch_mul;  // 0 when A is shared by all samples (stem), else 1
  int stack;        // > 1: the tile's BN columns hold `stack` consecutive samples of N channels each (shared A only):
                    //      one A tile feeds `stack` samples, B rows are the flattened [G*N][K] weights
  int g_blocks;     // ceil(G / stack)
  // im2col geometry
  int Wo, Ho, imgs_per_sample, stride, pad, kw, c_blocks;
  // outputs
  __half* y;             // [G][M][N]
  float* stats;          // [G][m_tiles][N][2] or nullptr
  const float* bias;     // [G][N] or nullptr (sampled bias, linear layers)
  // EPI == 2 (fused BatchNorm epilogue): out = relu?(acc * scale + shift [+ residual])
  const float2* ss;      // [G][N] (scale, shift)
  int has_res, relu;
  // fp16x3 validation mode (K-concatenated hi/lo operands, see mauv_gemm_x3_f16):
  int a_wrap_kb;         // tiled A: k-block count after which A's column coordinate wraps to 0 (0 = never)
  int a_cwrap;           // im2col A: channel-block period of A (0 = never)
  int split;             // epilogue writes (hi | lo) fp16 pairs: hi at column n, lo at column N + n
  int out_f32;           // EPI 0: the fp32 accumulator is written as is to y (float [G][M][N], plain vector stores, no TMA, no
                         // statistics): weight-gradient partial sums must not be rounded to / overflow fp16
  // weight-gradient mode (mauv_wgrad_f16): Y[batch][co][k] = sum_pixels dY[pixel][co] * Xcol[pixel][k]. Both operands are
  // MN-major: TMA boxes of [64 pixels][64 channels] straight from the row-major dY and the NHWC activations (tiled for
  // 1x1 / stride 1, im2col-mode for everything else); the batch index is (sample, pixel chunk), k_blocks = chunk / 64.
  int a2_kb;             // > 0: tiled A is K-concatenated from two tensors: k-blocks >= a2_kb come from the map in the tmR
                         //      slot (fused bottleneck tail with a downsample branch: [a2 | x] * [s3*W3 | sd*Wd]^T)
  int b_mod;             // > 0: B (the [N][K] operand) has only b_mod batch entries, entry = g % b_mod (an operand shared by
                         //      groups of batches: the stem's im2col matrix in the weight gradient)
  int mn;                // 1: weight-gradient mode
  int b_im2col;          // mn: B boxes come from im2col-mode TMA (else tiled rows of [pixels][Cin])
  int gram;              // mn: dY and X are the SAME tensor (second moments a^T a): only the B boxes are loaded and the A
                         //     descriptor points at boxes m0/64, m0/64+1 of the B tile (N = K <= BN: one n-tile)
  int cin;               // mn: input channels (column -> (tap, channel block))
  long long chunk;       // mn: pixels per batch entry
  // XF instances: the activation operand is the RAW output of the previous conv and that layer's BatchNorm + ReLU is applied to
  // the TMA-loaded tiles in shared memory (4 transform warps between the TMA and the MMA stage), so the activated tensor is
  // never written to / re-read from HBM. xf_ss [samples][xf_K] (scale, shift); K-major A: the first xf_kb k-blocks are
  // transformed (all of them, or the a2 part of a K-concatenated tail); gram mode: every box, sample = batch / xf_splits, and
  // the column sums of the transformed values go to xf_colsum [G][K] (one row per batch entry = (sample, pixel chunk)).
  const float2* xf_ss;
  int xf_K, xf_kb, xf_splits;
  float* xf_colsum;
};

// Epilogue flavours: 0 = raw fp16 store + BN statistics, 1 = BN statistics only (no output: first pass of the
// recompute scheme), 2 = fused BN-apply (+ residual) (+ ReLU) store (second pass; no raw conv output ever hits HBM).
enum { EPI_STORE_STATS = 0, EPI_STATS_ONLY = 1, EPI_FUSED_BN = 2, EPI_STATS_T = 3 };
// EPI_STATS_T: statistics pass with the operand roles swapped (MMA-M = output channels, MMA-N = pixels): a TMEM lane
// is a channel, so every epilogue thread sums its own channel over the tile's pixels in registers - no shuffles, no
// shared memory, no output. Pixels past M are zero-filled by TMA and contribute nothing.

// NR (fused-BN epilogue without a residual operand - the K-concatenated downsample tails): no residual prefetch, so two staging
// buffers per warp suffice and the 32 KB they free buy a third pipeline stage at BN = 256 (a tile is 2-6 k-blocks there and two
// stages gave the producer less than one tile of look-ahead: 7 000 clocks per tile against 4 200 of HBM time).
// BR (fused-BN tails whose weight tile is at most 64 KB: K <= 128 at BN = 256): with the n-tile fastest tile order and a grid
// that is a multiple of the n-tile count, a CTA meets ONE n-tile per sample, so its B tile ([BN][K] fp16) is loaded once per
// sample into a resident region and only the 16 KB A k-blocks stream: 3 A stages (5 without residual staging) = 1.5-5 tiles of
// look-ahead instead of 1-2, and no per-tile re-fetch of 32-64 KB of weights from L2.
template <int BN, int EPI = 0, bool GRAMX = false, bool NR = false, bool BR = false>
struct SmemLayout {
  // the fused epilogue needs 3 staging buffers per warp (residual prefetch / transform / store in flight) and is
  // only used for short-K (HBM-bound) layers, so it trades pipeline depth for staging space.
  // GRAMX (second moments with the operand transform): only the B boxes are loaded (the A descriptor aliases them), so a stage
  // is BN x 128 bytes and the ring is 2x deeper - every stage is also held for the ~700 clocks of its in-place transform, and
  // 6 x 8 KB in flight per SM did not cover the HBM latency
  static constexpr int kStages = BR ? (NR ? 5 : 3) : GRAMX ? ((BN == 256) ? 4 : (BN == 128 ? 8 : 12))
                                       : ((EPI == 2) ? ((BN == 256) ? (NR ? 3 : 2) : (NR ? 4 : 3)) : ((BN == 256) ? 3 : (BN == 128 ? 4 : 6)));
  static constexpr int kBResBytes = BR ? 65536 : 0;       // resident B tile (after the barrier block)
  static constexpr int kOutBufs = (EPI == 2 && !NR) ? 3 : 2;
  // epilogue warps: one set of 4 (TMEM lane quarters) per 64-column block in flight; two sets when BN >= 128
  static constexpr int kEpiWarps = (BN >= 128) ? 8 : 4;
  static constexpr int kABytes = GRAMX ? 0 : BM * BK * 2;
  static constexpr int kBBytes = BR ? 0 : BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // epilogue staging for the TMA store: per epilogue warp 2 buffers of 32 rows x 64 fp16 (128B-swizzled rows)
  static constexpr int kOutBufBytes = 32 * 64 * 2;
  static constexpr int kOutBytes = kEpiWarps * kOutBufs * kOutBufBytes;
  static constexpr int kStatBytes = 2 /*buffers*/ * 4 /*warps*/ * BN * 2 * 4;
  static constexpr int kBarBytes = 1024;  // pipeline + TMEM barriers, TMEM pointer, 3 residual barriers per epilogue warp, transform barriers
  static constexpr int kTotal = 1024 /*alignment slack*/ + kStages * kStageBytes + kOutBytes + kStatBytes + kBarBytes + kBResBytes;
};

// PLAIN = true: the common case (one sample per tile column block, fp16 output, no bias, K-major tiled / im2col A, no
// K-concatenation or wrap) with every mode flag folded at compile time - the generic epilogue carries ~590 instructions
// per 32x64 block, a third of them branches and flag loads for modes that are off (ncu: 14 % of the epilogue's stall
// samples are instruction-fetch stalls).
// PLAIN = 2: the stem's sample-stacked instance (A shared by all samples, 4 samples of 64 channels side by side in one
// 128 x 256 tile), again with every other flag folded: the stem writes the largest tensor of the network and ran the
// generic epilogue (2.1 TB/s of output against 4.1 TB/s for the same volume through the PLAIN = 1 instance).
// PLAIN = 3 (BN = 256, N <= 128, im2col A): TWO 128-row A tiles per CTA tile share one B tile - accumulator columns 0..127 hold
// rows m0 .. m0+127, columns 128..255 rows m0+128 .. m0+255. The N = 128 3x3 convs (K = 1152) were bound by the operand
// stream into shared memory (16 KB A + 16 KB B per 128 x 128 x 64 block = one 128-byte TMA row per MMA clock); sharing B
// moves 25 % fewer bytes and rows per flop.
// XF instances: the fused-BN epilogue needs its ~166 registers (a 448-thread block is capped at 128: it spilled), so there the
// transform runs on the two otherwise idle warps 2 and 3; the second-moment (gram) instance has a light epilogue and takes four
// transform warps (2, 3, 12, 13) in a 448-thread block.
template <int EPI> constexpr int xf_warps() { return EPI == 2 ? 2 : 4; }
// second-moment instance: k-blocks in transform at the same time (one group of 4 / split warps each), bounded by the ring depth
template <int BN> constexpr int xf_gram_split() { return BN == 256 ? 2 : 4; }
template <int BN, int EPI> constexpr int xf_arrivals() { return EPI == 2 ? 2 : 4 / xf_gram_split<BN>(); }
template <int EPI, bool XF> constexpr int gemm_threads() { return (XF && xf_warps<EPI>() == 4) ? 448 : 384; }

template <int BN, int EPI, int PLAIN, bool XF = false, bool NR = false, bool BR = false>
__global__ void __launch_bounds__(gemm_threads<EPI, XF>(), 1)
gemm_f16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                   const GemmParams p) {
  using L = SmemLayout<BN, EPI, XF && EPI == EPI_STORE_STATS, NR, BR>;
  static_assert(!NR || EPI == EPI_FUSED_BN, "NR: fused-BN epilogue without residual");
  static_assert(!BR || (EPI == EPI_FUSED_BN && BN == 256), "BR: resident weight tile of the fused-BN tails");
  constexpr int kStages = L::kStages;
  const int f_stack = PLAIN == 2 ? 4 : (PLAIN ? 1 : p.stack), f_split = PLAIN ? 0 : p.split, f_out_f32 = PLAIN ? 0 : p.out_f32;
  const int f_mn = PLAIN ? 0 : p.mn, f_gram = PLAIN ? 0 : p.gram, f_a2_kb = PLAIN ? 0 : p.a2_kb;
  const int f_a_wrap_kb = PLAIN ? 0 : p.a_wrap_kb, f_a_cwrap = PLAIN ? 0 : p.a_cwrap, f_b_mod = PLAIN ? 0 : p.b_mod;
  const int f_batch_mul = PLAIN == 2 ? 0 : (PLAIN ? 1 : p.a_batch_mul);
  const int f_N = PLAIN == 2 ? 64 : p.N;            // channels per sample (only read by the stacked-mode index math)
  constexpr bool MSTACK = (PLAIN == 3);             // two M tiles per CTA tile (p.m_tiles counts 256-row pairs in the tile index)
  static_assert(!MSTACK || (BN == 256 && EPI == EPI_STORE_STATS), "M-stacked tiles: BN = 256 accumulator columns, store + statistics");
  const float* const f_bias = PLAIN ? nullptr : p.bias;
  constexpr uint32_t kTmemCols = 2 * BN;  // 128 / 256 / 512: power of two >= 32

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t tiles_base = smem_base;
  const uint32_t out_base = smem_base + kStages * L::kStageBytes;   // 1024-aligned (stage bytes are multiples of 1024)
  float* stat_smem = reinterpret_cast<float*>(smem_gen + kStages * L::kStageBytes + L::kOutBytes);
  const uint32_t bar_base = smem_base + kStages * L::kStageBytes + L::kOutBytes + L::kStatBytes;
  // barrier layout (8 bytes each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kStages + 4);
  auto res_bar = [&](int w, int j) { return bar_base + 8u * (2 * kStages + 6 + w * 3 + j); };   // EPI == 2 only
  auto xf_ready = [&](int s) { return bar_base + 8u * (2 * kStages + 30 + s); };                // XF only: tile transformed
  const uint32_t b_full = bar_base + 8u * 120, b_empty = bar_base + 8u * 121;                   // BR only: resident B tile
  const uint32_t bres_base = bar_base + L::kBarBytes;                                           // BR only (1024-aligned)
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + kStages * L::kStageBytes + L::kOutBytes + L::kStatBytes + 8 * (2 * kStages + 4));

  const int warp = threadIdx.x >> 5;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
    if (EPI == EPI_FUSED_BN) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), EPI == EPI_STATS_T ? 4 : L::kEpiWarps);  // one arrive per epilogue warp
    }
    if (EPI == EPI_FUSED_BN)
      for (int w = 0; w < L::kEpiWarps; ++w)
        for (int j = 0; j < 3; ++j) mbar_init(res_bar(w, j), 1);
    if (XF)
      for (int s = 0; s < kStages; ++s) mbar_init(xf_ready(s), xf_arrivals<BN, EPI>());   // one arrive per warp that transforms the stage
    if (BR) {
      mbar_init(b_full, 1);
      mbar_init(b_empty, 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc<kTmemCols>(tmem_ptr_addr);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  const long long tiles_per_sample = static_cast<long long>(p.m_tiles) * p.n_tiles;
  // tile -> (g, m_tile, n_tile). n-tile fastest so concurrent CTAs share their A tile through L2; when A is shared
  // by all samples (stem) the sample index is the fastest instead, so the G samples of one m-tile run together.
  // (32-bit unsigned arithmetic: total_tiles < 2^31 is checked on the host; a 64-bit division costs ~100 instructions and
  // the epilogue decodes once per 64-column block)
  const uint32_t tps32 = static_cast<uint32_t>(tiles_per_sample);
  auto decode = [&](long long tile64, int& g, int& m_tile, int& n_tile) {
    const uint32_t tile = static_cast<uint32_t>(tile64);
    const uint32_t mt = static_cast<uint32_t>(p.m_tiles), nt = static_cast<uint32_t>(p.n_tiles);
    if (EPI == EPI_STATS_T) {       // channel tile fastest: the (big) pixel operand is fetched from HBM once
      const uint32_t rem = tile / mt;
      m_tile = static_cast<int>(tile - rem * mt);
      const uint32_t gg = rem / nt;
      n_tile = static_cast<int>(rem - gg * nt);
      g = static_cast<int>(gg);
    } else if (f_stack > 1) {       // g = first sample of the tile's sample block; a single n-tile
      const uint32_t gbk = static_cast<uint32_t>(p.g_blocks);
      const uint32_t mm = tile / gbk;
      g = static_cast<int>(tile - mm * gbk) * f_stack;
      m_tile = static_cast<int>(mm);
      n_tile = 0;
    } else if (f_batch_mul == 0) {
      const uint32_t G = static_cast<uint32_t>(p.G);
      const uint32_t rem = tile / G;
      g = static_cast<int>(tile - rem * G);
      const uint32_t mm = rem / nt;
      n_tile = static_cast<int>(rem - mm * nt);
      m_tile = static_cast<int>(mm);
    } else {
      const uint32_t gg = tile / tps32;
      const uint32_t rem = tile - gg * tps32;
      const uint32_t mm = (nt == 1u) ? rem : rem / nt;
      g = static_cast<int>(gg);
      m_tile = static_cast<int>(mm);
      n_tile = static_cast<int>(rem - mm * nt);
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int br_g = -1;
      uint32_t br_loads = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int g, m_tile, n_tile;
        decode(tile, g, m_tile, n_tile);
        if (BR && g != br_g) {      // new sample: its weight tile replaces the resident one once the MMAs reading it have retired
          if (br_loads > 0) mbar_wait(b_empty, (br_loads - 1) & 1u);
          mbar_expect_tx(b_full, static_cast<uint32_t>(p.k_blocks) * (BN * BK * 2));
          for (int kb = 0; kb < p.k_blocks; ++kb) tma_load_3d(bres_base + kb * (BN * BK * 2), &tmB, b_full, kb * BK, n_tile * BN, g);
          ++br_loads;
          br_g = g;
        }
        const int m0 = m_tile * (MSTACK ? 2 * BM : BM);
        // im2col start pixel of this tile (M-stacked: of its two 128-row halves)
        int iq = 0, ip = 0, in_ = 0, iq1 = 0, ip1 = 0, in1 = 0;
        if (p.a_mode == 1) {
          const int hw = p.Ho * p.Wo;
          const int b = m0 / hw;
          const int r2 = m0 - b * hw;
          ip = r2 / p.Wo;
          iq = r2 - ip * p.Wo;
          in_ = g * p.imgs_per_sample + b;
          if (MSTACK) {
            const int b1 = (m0 + BM) / hw;
            const int r3 = (m0 + BM) - b1 * hw;
            ip1 = r3 / p.Wo;
            iq1 = r3 - ip1 * p.Wo;
            in1 = g * p.imgs_per_sample + b1;
          }
        }
        if (f_mn) {
          // weight-gradient mode: 64-pixel k-blocks; A = 2 boxes of 64 output channels, B = BN/64 boxes of 64 columns
          const int n0 = n_tile * BN;
          int a_boxes = 0, b_boxes = 0;
          for (int j = 0; j < 2; ++j) a_boxes += (m0 + 64 * j < p.M);
          for (int j = 0; j < BN / 64; ++j) b_boxes += (n0 + 64 * j < p.N);
          const int hw = p.Ho * p.Wo;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t a_dst = tiles_base + stage * L::kStageBytes;
            const uint32_t b_dst = a_dst + L::kABytes;
            mbar_expect_tx(full_bar(stage), static_cast<uint32_t>((f_gram ? 0 : a_boxes) + b_boxes) * 8192u);
            const long long pix0 = static_cast<long long>(g) * p.chunk + static_cast<long long>(kb) * 64;
            if (!f_gram)
              for (int j = 0; j < a_boxes; ++j)
                tma_load_3d(a_dst + j * 8192, &tmA, full_bar(stage), m0 + 64 * j, static_cast<int>(pix0), 0);
            if (p.b_im2col) {
              const int img = static_cast<int>(pix0 / hw);
              const int r2 = static_cast<int>(pix0 - static_cast<long long>(img) * hw);
              const int bp = r2 / p.Wo, bq = r2 - bp * p.Wo;
              for (int j = 0; j < b_boxes; ++j) {
                const int col = n0 + 64 * j;
                const int tap = col / p.cin, c0 = col - tap * p.cin;
                const int r = tap / p.kw, s2 = tap - r * p.kw;
                tma_load_im2col_4d(b_dst + j * 8192, &tmB, full_bar(stage), c0, bq * p.stride - p.pad,
                                   bp * p.stride - p.pad, img, static_cast<uint16_t>(s2), static_cast<uint16_t>(r));
              }
            } else {
              for (int j = 0; j < b_boxes; ++j)
                tma_load_3d(b_dst + j * 8192, &tmB, full_bar(stage), n0 + 64 * j, static_cast<int>(pix0), 0);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          continue;
        }
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = tiles_base + stage * L::kStageBytes;
          const uint32_t b_dst = a_dst + L::kABytes;
          mbar_expect_tx(full_bar(stage), L::kStageBytes);
          if (p.a_mode == 0) {
            if (f_a2_kb && kb >= f_a2_kb) {
              tma_load_3d(a_dst, &tmR, full_bar(stage), (kb - f_a2_kb) * BK, m0, g);
            } else {
              const int akb = f_a_wrap_kb ? kb % (f_a_wrap_kb ? f_a_wrap_kb : 1) : kb;
              tma_load_3d(a_dst, &tmA, full_bar(stage), akb * BK, m0, g * f_batch_mul);
            }
          } else {
            const int tap = kb / p.c_blocks;
            int cb = kb - tap * p.c_blocks;
            if (f_a_cwrap) cb %= (f_a_cwrap ? f_a_cwrap : 1);     // (the ternary only silences the constant-folded '% 0' warning)
            const int r = tap / p.kw;
            const int s = tap - r * p.kw;
            tma_load_im2col_4d(a_dst, &tmA, full_bar(stage), cb * BK, iq * p.stride - p.pad,
                               ip * p.stride - p.pad, in_, static_cast<uint16_t>(s),
                               static_cast<uint16_t>(r));
            if (MSTACK)      // second 128-row half; stage layout: A0 16 KB | A1 16 KB | B 16 KB
              tma_load_im2col_4d(a_dst + L::kABytes, &tmA, full_bar(stage), cb * BK, iq1 * p.stride - p.pad,
                                 ip1 * p.stride - p.pad, in1, static_cast<uint16_t>(s), static_cast<uint16_t>(r));
          }
          if (BR) { /* B is resident */ }
          else if (MSTACK) tma_load_3d(a_dst + 2 * L::kABytes, &tmB, full_bar(stage), kb * BK, 0, g);
          else if (f_stack > 1) tma_load_3d(b_dst, &tmB, full_bar(stage), kb * BK, g * f_N, 0);   // flattened [G*N][K]
          else tma_load_3d(b_dst, &tmB, full_bar(stage), kb * BK, n_tile * BN, f_b_mod ? g % f_b_mod : g);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      int stage = 0, br_g = -1;
      uint32_t phase = 0, br_uses = 0;
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u;
        const uint32_t acc_phase = (it >> 1) & 1u;
        int tile_m0 = 0;
        if (f_gram) {
          int g_, m_tile_, n_tile_;
          decode(tile, g_, m_tile_, n_tile_);
          tile_m0 = m_tile_ * BM;
        }
        if (BR) {
          int g_, m_tile_, n_tile_;
          decode(tile, g_, m_tile_, n_tile_);
          if (g_ != br_g) {
            if (br_g >= 0) umma_commit(b_empty);      // arrives when every MMA issued so far (old weights) has completed
            mbar_wait(b_full, br_uses & 1u);
            tcgen05_fence_after();
            ++br_uses;
            br_g = g_;
          }
        }
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(XF ? xf_ready(stage) : full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t a_addr = tiles_base + stage * L::kStageBytes;
          if (f_mn) {
            // MN-major boxes [64 pixels][64 channels]: one MMA consumes 16 pixel rows = 2048 bytes (+128 in the >>4 field)
            constexpr uint32_t idesc_mn = umma_idesc_f16_mn(BM, BN);
            // gram mode: the A rows (channels m0 .. m0+127) are boxes m0/64, m0/64+1 of the B tile
            const uint32_t a_src = f_gram ? a_addr + L::kABytes + static_cast<uint32_t>((tile_m0 >> 6) * 8192) : a_addr;
            const uint64_t a_desc = umma_smem_desc_mn_sw128(a_src, 8192);
            const uint64_t b_desc = umma_smem_desc_mn_sw128(a_addr + L::kABytes, 8192);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_f16_ss(d_tmem, a_desc + 128u * k, b_desc + 128u * k, idesc_mn, (kb | k) != 0 ? 1u : 0u);
          } else if (MSTACK) {
            constexpr uint32_t idesc_half = umma_idesc_f16(BM, 128);
            const uint64_t a0_desc = umma_smem_desc_sw128(a_addr);
            const uint64_t a1_desc = umma_smem_desc_sw128(a_addr + L::kABytes);
            const uint64_t b_desc = umma_smem_desc_sw128(a_addr + 2 * L::kABytes);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              umma_f16_ss(d_tmem, a0_desc + 2u * k, b_desc + 2u * k, idesc_half, (kb | k) != 0 ? 1u : 0u);
              umma_f16_ss(d_tmem + 128u, a1_desc + 2u * k, b_desc + 2u * k, idesc_half, (kb | k) != 0 ? 1u : 0u);
            }
          } else {
            const uint64_t a_desc = umma_smem_desc_sw128(a_addr);
            const uint64_t b_desc = umma_smem_desc_sw128(BR ? bres_base + static_cast<uint32_t>(kb) * (BN * BK * 2) : a_addr + L::kABytes);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 fp16 = 32 bytes inside the 128-byte swizzle row: +2 in the (>>4) address field
              umma_f16_ss(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tmem_full_bar(acc));  // accumulator complete -> epilogue
      }
    }
  } else if (XF && (warp == 2 || warp == 3 || warp >= 12)) {
    constexpr int kXfThreads = 32 * xf_warps<EPI>();      // 64 or 128
    constexpr uint32_t kXfRowStep = kXfThreads / 8;       // rows between two chunks of one thread
    // ===================== operand transform (XF instances) =====================
    // The activation operand arrives RAW (the previous conv's output); BatchNorm + ReLU of that layer is applied here, in
    // place, to every TMA-loaded tile before the MMA warp may read it: a = relu(y * scale + shift), fp32 math and one fp16
    // rounding - exactly bn_act_kernel's arithmetic, so the MMA sees bit-identical operands. kXfThreads threads: thread t owns the
    // logical 16-byte chunk lc = t & 7 (8 channels) of rows (t >> 3) + kXfRowStep i; the physical chunk is lc ^ (row & 7) (128B
    // swizzle). Rows past the tensor (TMA zero fill) become relu(shift): K-major tiles never store those rows; gram-mode
    // chunks are whole multiples of 64 pixels.
    static_assert(!XF || EPI != EPI_FUSED_BN || BN >= 128, "the (scale, shift) table of the XF fused epilogue lives behind my_ss");
    const int tw = warp >= 12 ? warp - 10 : warp - 2;                   // 0..3
    const uint32_t t = static_cast<uint32_t>(tw) * 32u + lane_id();
    const uint32_t lc = t & 7u, r0 = t >> 3;
    float* xf_tab = stat_smem + (EPI == EPI_FUSED_BN ? 1024 : 0);     // scale[256] | shift[256], de-interleaved
    const uint32_t xf_tab_addr = smem_u32(xf_tab);
    float* xf_scr = reinterpret_cast<float*>(smem_gen + kStages * L::kStageBytes);               // gram: [16][K] (out staging is idle)
    // 8 consecutive channels' scales / shifts as packed fp32x2 operands (two broadcast LDS.128 each)
    auto load_ss = [&](uint32_t ch0, unsigned long long (&sc)[4], unsigned long long (&sh)[4]) {
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(sc[0]), "=l"(sc[1]) : "r"(xf_tab_addr + ch0 * 4u));
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(sc[2]), "=l"(sc[3]) : "r"(xf_tab_addr + ch0 * 4u + 16u));
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(sh[0]), "=l"(sh[1]) : "r"(xf_tab_addr + 1024u + ch0 * 4u));
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(sh[2]), "=l"(sh[3]) : "r"(xf_tab_addr + 1024u + ch0 * 4u + 16u));
    };
    auto lds128 = [&](uint32_t addr, uint32_t (&wv)[4]) {
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(wv[0]), "=r"(wv[1]), "=r"(wv[2]), "=r"(wv[3]) : "r"(addr));
    };
    auto sts128 = [&](uint32_t addr, const uint32_t (&wv)[4]) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]) : "memory");
    };
    // relu(y * scale + shift) on 8 fp16 values: fp32 fma per element (packed fma.rn.f32x2: the same rounding as fmaf), one fp16
    // rounding, ReLU on the rounded pair (max(round(x), 0) == round(max(x, 0))) - bit-identical to bn_act_kernel
    auto xform8 = [&](uint32_t (&wv)[4], const unsigned long long (&sc)[4], const unsigned long long (&sh)[4]) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __half22float2(*reinterpret_cast<__half2*>(&wv[q]));
        unsigned long long a2, r2;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "f"(f.x), "f"(f.y));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r2) : "l"(a2), "l"(sc[q]), "l"(sh[q]));
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r2));
        const __half2 o = __hmax2(__floats2half2_rn(lo, hi), __float2half2_rn(0.f));
        wv[q] = *reinterpret_cast<const uint32_t*>(&o);
      }
    };
    int stage = 0, cur_smp = -1;
    uint32_t phase = 0, kbc = 0;           // kbc: k-blocks seen so far (all tiles)
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int g, m_tile, n_tile;
      decode(tile, g, m_tile, n_tile);
      const int smp = f_mn ? g / p.xf_splits : g;
      if (smp != cur_smp) {        // this sample's (scale, shift) table; the 4 warps walk the same tile sequence
        asm volatile("bar.sync 3, %0;" ::"n"(kXfThreads) : "memory");
        for (int i = static_cast<int>(t); i < p.xf_K; i += kXfThreads) {
          const float2 v = __ldg(p.xf_ss + static_cast<long long>(smp) * p.xf_K + i);
          xf_tab[i] = v.x;
          xf_tab[256 + i] = v.y;
        }
        asm volatile("bar.sync 3, %0;" ::"n"(kXfThreads) : "memory");
        cur_smp = smp;
      }
      if (EPI == EPI_STORE_STATS && f_mn) {
        // gram mode: BN/64 boxes of [64 pixels][64 channels] per k-block, channels 64 j + 8 glc .. of box j. A k-block is only
        // 8 KB per box and every hand-over costs a fixed ~400 clocks (mbarrier wait, fence.proxy.async, arrive), so the four
        // warps do NOT share one k-block: they form kSplit groups that take every kSplit-th k-block (ring depth permitting),
        // i.e. kSplit k-blocks are being transformed at any time.
        constexpr int kBoxes = BN / 64;
        constexpr int kGw = 4 / xf_gram_split<BN>();              // warps per group
        constexpr uint32_t kGRowStep = 4u * kGw;                  // rows between two chunks of one thread
        constexpr uint32_t kGChunks = 64u / kGRowStep;            // chunks per box per thread: 16, 8 or 4
        const int grp = tw / kGw;
        const uint32_t tg = static_cast<uint32_t>(tw % kGw) * 32u + lane_id();
        const uint32_t glc = tg & 7u, gr0 = tg >> 3;
        unsigned long long cs[kBoxes][4];                         // packed column sums of channels 64 j + 8 glc + (2q, 2q+1)
#pragma unroll
        for (int j = 0; j < kBoxes; ++j)
#pragma unroll
          for (int q = 0; q < 4; ++q) cs[j][q] = 0ull;
        const int b_boxes = (p.N + 63) / 64;                      // one n-tile holds all K columns (gram)
        for (int kb = 0; kb < p.k_blocks; ++kb, ++kbc) {
          if (static_cast<int>(kbc % xf_gram_split<BN>()) == grp) {
            mbar_wait(full_bar(stage), phase);
            const uint32_t b_dst = tiles_base + stage * L::kStageBytes + L::kABytes;
#pragma unroll
            for (int j = 0; j < kBoxes; ++j) {
              if (j < b_boxes) {
                unsigned long long sc[4], sh[4];
                load_ss(64u * j + 8u * glc, sc, sh);
#pragma unroll
                for (uint32_t c4 = 0; c4 < kGChunks / 4u; ++c4) {
                  uint32_t wv[4][4];
#pragma unroll
                  for (uint32_t i = 0; i < 4; ++i) {                // four loads first: one exposed shared-memory latency
                    const uint32_t row = gr0 + kGRowStep * (4u * c4 + i);
                    lds128(b_dst + j * 8192u + row * 128u + ((glc ^ (row & 7u)) << 4), wv[i]);
                  }
#pragma unroll
                  for (uint32_t i = 0; i < 4; ++i) {
                    const uint32_t row = gr0 + kGRowStep * (4u * c4 + i);
                    xform8(wv[i], sc, sh);
                    sts128(b_dst + j * 8192u + row * 128u + ((glc ^ (row & 7u)) << 4), wv[i]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {        // column sums of the values the tensor core will read
                      const float2 fo = __half22float2(*reinterpret_cast<__half2*>(&wv[i][q]));
                      unsigned long long f2;
                      asm("mov.b64 %0, {%1, %2};" : "=l"(f2) : "f"(fo.x), "f"(fo.y));
                      asm("add.rn.f32x2 %0, %0, %1;" : "+l"(cs[j][q]) : "l"(f2));
                    }
                  }
                }
              }
            }
            fence_proxy_async_smem();        // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane_id() == 0) mbar_arrive(xf_ready(stage));
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (p.xf_colsum && m_tile == 0) {
          // column sums of the transformed chunk: 16 partial rows per channel (group x row lane), combined in a fixed order
          asm volatile("bar.sync 3, 128;" ::: "memory");        // the previous tile's readers are done with the scratch
#pragma unroll
          for (int j = 0; j < kBoxes; ++j)
            if (j < b_boxes) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float2*>(xf_scr + (grp * kGRowStep + gr0) * p.N + 64 * j + 8 * glc + 2 * q) =
                    *reinterpret_cast<float2*>(&cs[j][q]);
            }
          asm volatile("bar.sync 3, 128;" ::: "memory");
          for (int c = static_cast<int>(t); c < p.N; c += 128) {
            float a = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) a += xf_scr[r * p.N + c];
            p.xf_colsum[static_cast<long long>(g) * p.N + c] = a;
          }
        }
      } else if (EPI == EPI_FUSED_BN) {
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if (kb < p.xf_kb) {
            const uint32_t a_dst = tiles_base + stage * L::kStageBytes;
            unsigned long long sc[4], sh[4];
            load_ss(64u * kb + 8u * lc, sc, sh);
#pragma unroll
            for (uint32_t grp = 0; grp < 128u / kXfRowStep / 4u; ++grp) {      // 4 loads in flight, then 4 transforms + stores
              uint32_t wv[4][4];
#pragma unroll
              for (uint32_t i = 0; i < 4; ++i) {
                const uint32_t row = r0 + kXfRowStep * (4u * grp + i);
                lds128(a_dst + row * 128u + ((lc ^ (row & 7u)) << 4), wv[i]);
              }
#pragma unroll
              for (uint32_t i = 0; i < 4; ++i) {
                const uint32_t row = r0 + kXfRowStep * (4u * grp + i);
                xform8(wv[i], sc, sh);
                sts128(a_dst + row * 128u + ((lc ^ (row & 7u)) << 4), wv[i]);
              }
            }
            fence_proxy_async_smem();
          }
          __syncwarp();
          if (lane_id() == 0) mbar_arrive(xf_ready(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + L::kEpiWarps) {
    // ===================== epilogue =====================
    // TMEM -> registers -> fp16 -> 128B-swizzled smem staging -> TMA store (coalesced, asynchronous, clipped at
    // the tensor bounds), 64 output channels at a time per warp. The BatchNorm (sum, sum of squares) come from the
    // staged fp16 values - exactly the values the BN pass will normalise.
    constexpr int kEpiThreads = L::kEpiWarps * 32;
    constexpr int kColSets = L::kEpiWarps / 4;       // column-block sets working concurrently
    const int ew = warp & 3;                  // TMEM lane quarter accessible to this warp (warp_id % 4)
    const int cset = (warp - 4) >> 2;         // which 64-column blocks this warp takes (cb = cset, cset + kColSets, ..)
    const uint32_t lane = lane_id();
    const uint32_t my_out = out_base + (warp - 4) * (L::kOutBufs * L::kOutBufBytes);
    // swizzled 16-byte chunk offsets of this lane's channel pair for rows r = 0..7 (mod 8)
    uint32_t sw_off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sw_off[j] = (((lane >> 2) ^ static_cast<uint32_t>(j)) << 4) + ((lane & 3u) << 2);
    if constexpr (EPI == EPI_STATS_T) {
      if (warp < 8) {   // one warp per TMEM lane quarter (= 32 channels) is enough: pure register accumulation
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
          int g, m_tile, n_tile;
          decode(tile, g, m_tile, n_tile);          // m_tile: channel tile, n_tile: pixel tile
          const uint32_t acc = it & 1u;
          const uint32_t acc_phase = (it >> 1) & 1u;
          const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(ew * 32) << 16);
          mbar_wait(tmem_full_bar(acc), acc_phase);
          tcgen05_fence_after();
          unsigned long long s2 = 0ull, q2 = 0ull, s3 = 0ull, q3 = 0ull;
#pragma unroll 1
          for (int c = 0; c < BN / 64; ++c) {
            uint32_t r[32], t[32];
            tmem_ld_32x32b_x32(taddr + c * 64, r);        // two loads in flight per wait
            tmem_ld_32x32b_x32(taddr + c * 64 + 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              // consecutive accumulator registers are used directly as packed fp32x2 operands
              unsigned long long f2, g2;
              asm("mov.b64 %0, {%1, %2};" : "=l"(f2) : "r"(r[j]), "r"(r[j + 1]));
              asm("mov.b64 %0, {%1, %2};" : "=l"(g2) : "r"(t[j]), "r"(t[j + 1]));
              asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s2) : "l"(f2));
              asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q2) : "l"(f2));
              asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s3) : "l"(g2));
              asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q3) : "l"(g2));
            }
          }
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s2) : "l"(s3));
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(q2) : "l"(q3));
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
          const float2 sf = *reinterpret_cast<float2*>(&s2), qf = *reinterpret_cast<float2*>(&q2);
          const int ch = m_tile * BM + ew * 32 + static_cast<int>(lane);
          if (ch < p.M)   // p.M = channels, p.n_tiles = pixel tiles in this mode
            reinterpret_cast<float2*>(p.stats)[(static_cast<long long>(g) * p.n_tiles + n_tile) * p.M + ch] =
                make_float2(sf.x + sf.y, qf.x + qf.y);
        }
      }
    } else if constexpr (EPI == EPI_FUSED_BN) {
      // ---- fused BN-apply (+ residual) (+ ReLU): out = relu?(acc * scale[g][n] + shift[g][n] + res[g][m][n]).
      // Per warp a 3-deep ring of 32x64 fp16 staging buffers: while block b is transformed in place, the residual
      // tile of block b+1 is already in flight (TMA load, one mbarrier per buffer) and block b-1 is being stored.
      constexpr int kBlocksPerTile = BN / 64;
      const int wq = warp - 4;
      auto block_coords = [&](long long tile, int cb, int& gb, int& nb, int& row0, int& rmax) {
        int g, m_tile, n_tile;
        decode(tile, g, m_tile, n_tile);
        gb = g;
        nb = n_tile * BN + cb * 64;
        row0 = m_tile * BM + ew * 32;
        rmax = p.M - row0;
        rmax = rmax < 0 ? 0 : (rmax > 32 ? 32 : rmax);
      };
      auto issue_residual = [&](uint32_t b, long long tile, int cb) {   // lane 0 only
        int gb, nb, row0, rmax;
        block_coords(tile, cb, gb, nb, row0, rmax);
        if (!NR && p.has_res) {
          // issued for EVERY block so that buffer b%3's barrier completes exactly once per use (parity = (b/3)&1);
          // a slab entirely past M loads rows 0.. instead (never stored), N % BN == 0 keeps the columns in range
          const uint32_t bar = res_bar(wq, b % 3u);
          mbar_expect_tx(bar, L::kOutBufBytes);
          tma_load_3d(my_out + (b % static_cast<uint32_t>(L::kOutBufs)) * L::kOutBufBytes, &tmR, bar, nb, rmax > 0 ? row0 : 0, gb);
        }
      };
      // flat loop over this warp's work items (tile, 64-column block) with a one-item look-ahead on the TMEM side: the
      // tcgen05.ld of the next item's half is issued as soon as the current half has been transformed (see EPI 0 below)
      long long tile = blockIdx.x;
      int cb = cset;
      uint32_t it = 0, b = 0;
      uint32_t ra[32], rb[32];
      float* my_ss = stat_smem + wq * 128;          // the statistics staging area is unused by this epilogue flavour
      const uint32_t my_ss_addr = smem_u32(my_ss);
      int ss_g = -1, ss_n = -1;
      auto acc_addr = [&](uint32_t it_, int cb_) {
        return tmem_base + (it_ & 1u) * BN + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(cb_ * 64);
      };
      if (tile < p.total_tiles && cb < kBlocksPerTile) {
        if (lane == 0) issue_residual(0, tile, cb);
        mbar_wait(tmem_full_bar(0), 0);
        tcgen05_fence_after();
        tmem_ld_32x32b_x32(acc_addr(0, cb), ra);
        tmem_ld_32x32b_x32(acc_addr(0, cb) + 32, rb);
      }
      while (tile < p.total_tiles && cb < kBlocksPerTile) {
        int gb, nb, row0, rmax;
        block_coords(tile, cb, gb, nb, row0, rmax);
        long long ntile = tile;
        int ncb = cb + kColSets;
        uint32_t nit = it;
        if (ncb >= kBlocksPerTile) { ncb = cset; ntile += gridDim.x; ++nit; }
        const bool last_of_tile = ntile != tile;
        const bool has_next = ntile < p.total_tiles;
        if (lane == 0) {
          // buffer (b+1)%3 was last used by block b-2: its store must have finished reading smem
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          if (has_next) issue_residual(b + 1, ntile, ncb);
        }
        tmem_ld_wait();
        if (last_of_tile) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar(it & 1u));
        }
        auto prefetch = [&](uint32_t (&dst)[32], int half) {
          if (!has_next) return;
          if (last_of_tile && half == 0) {
            mbar_wait(tmem_full_bar(nit & 1u), (nit >> 1) & 1u);
            tcgen05_fence_after();
          }
          tmem_ld_32x32b_x32(acc_addr(nit, ncb) + 32 * half, dst);
        };
        const uint32_t buf = my_out + (b % static_cast<uint32_t>(L::kOutBufs)) * L::kOutBufBytes;
        if (!NR && p.has_res) mbar_wait(res_bar(wq, b % 3u), (b / 3u) & 1u);
        // (scale, shift) of this block's 64 channels, de-interleaved into this warp's private shared-memory slice (scale[64] |
        // shift[64]) and re-read with broadcast 16-byte loads: 32 LDS.128 per block instead of 64 LDG.64, and the pairs come
        // out as packed fp32x2 operands. Reloaded only when the (sample, channel block) changes - with N = BN once per sample.
        if (gb != ss_g || nb != ss_n) {
          __syncwarp();
          const float4 two = __ldg(reinterpret_cast<const float4*>(p.ss + static_cast<long long>(gb) * p.N + nb) + lane);
          my_ss[2 * lane] = two.x; my_ss[2 * lane + 1] = two.z;               // scales of channels 2*lane, 2*lane+1
          my_ss[64 + 2 * lane] = two.y; my_ss[64 + 2 * lane + 1] = two.w;     // shifts
          __syncwarp();
          ss_g = gb; ss_n = nb;
        }
        // One half = 4 chunks of 8 channels of this lane's row. The shared-memory reads are batched - all four residual chunks
        // first, then (scale, shift) for two chunks at a time - and the two chunks' arithmetic is interleaved: the ncu source view
        // of the chunk-at-a-time version had the epilogue warps (2 per scheduler) 32 % in `wait` and 19 % in `short_scoreboard`
        // stalls on eight serial LDS -> FFMA2 -> FADD2 -> F2FP -> HMNMX2 -> STS chains per block.
        auto transform_half = [&](const uint32_t (&r)[32], int half) {
          uint32_t rr[4][4];
          if (!NR && p.has_res) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint32_t addr = buf + lane * 128u + ((static_cast<uint32_t>(half * 4 + q4) ^ (lane & 7u)) << 4);
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[q4][0]), "=r"(rr[q4][1]), "=r"(rr[q4][2]), "=r"(rr[q4][3]) : "r"(addr));
            }
          }
#pragma unroll
          for (int pr = 0; pr < 2; ++pr) {
            unsigned long long sc[2][4], sh[2][4];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const uint32_t q = static_cast<uint32_t>(half * 4 + pr * 2 + c);
              asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(sc[c][0]), "=l"(sc[c][1]) : "r"(my_ss_addr + q * 32u));
              asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(sc[c][2]), "=l"(sc[c][3]) : "r"(my_ss_addr + q * 32u + 16u));
              asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(sh[c][0]), "=l"(sh[c][1]) : "r"(my_ss_addr + 256u + q * 32u));
              asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(sh[c][2]), "=l"(sh[c][3]) : "r"(my_ss_addr + 256u + q * 32u + 16u));
            }
            uint32_t h[2][4];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int q4 = pr * 2 + c;
              const uint32_t* src = &r[q4 * 8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                unsigned long long a2, v;
                asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "r"(src[2 * j]), "r"(src[2 * j + 1]));
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(a2), "l"(sc[c][j]), "l"(sh[c][j]));
                if (!NR && p.has_res) {
                  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rr[q4][j]));
                  unsigned long long f2;
                  asm("mov.b64 %0, {%1, %2};" : "=l"(f2) : "f"(f.x), "f"(f.y));
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(f2));
                }
                float lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
                __half2 hh = __floats2half2_rn(lo, hi);
                if (p.relu) hh = __hmax2(hh, __float2half2_rn(0.f));      // ReLU after rounding == rounding after ReLU
                h[c][j] = *reinterpret_cast<uint32_t*>(&hh);
              }
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const uint32_t addr = buf + lane * 128u + ((static_cast<uint32_t>(half * 4 + pr * 2 + c) ^ (lane & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h[c][0]), "r"(h[c][1]), "r"(h[c][2]), "r"(h[c][3]) : "memory");
            }
          }
        };
        transform_half(ra, 0);
        prefetch(ra, 0);
        transform_half(rb, 1);
        prefetch(rb, 1);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (rmax > 0)
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmY)),
                         "r"(buf), "r"(nb), "r"(row0), "r"(gb)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        tile = ntile; cb = ncb; it = nit; ++b;
      }
    } else {
    // Flat loop over this warp's work items (tile, 64-column block) with a one-item look-ahead on the TMEM side: as soon as
    // a half (32 columns) of the accumulator has been converted and staged, the tcgen05.ld of the NEXT item's half is
    // issued, so the TMEM read (64 B/clk per SM: 2 048 clk for a 128x256 fp32 tile) overlaps the staging, the TMA store
    // and the statistics pass instead of being waited for by all 8 warps at once. The two column sets (warps 4-7 / 8-11)
    // combine their statistics behind their OWN named barrier and may drift apart by up to one tile.
    constexpr int kBlocks = BN / 64;
    constexpr int kSetThreads = 128;
    const int es = threadIdx.x - 128 - cset * kSetThreads;        // 0..127 inside this column set
    long long tile = blockIdx.x;
    int cb = cset;
    uint32_t it = 0, blk = 0;
    uint32_t ra[32], rb[32];
    auto acc_addr = [&](uint32_t it_, int cb_) {
      return tmem_base + (it_ & 1u) * BN + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(cb_ * 64);
    };
    if (tile < p.total_tiles && cb < kBlocks) {
      mbar_wait(tmem_full_bar(0), 0);
      tcgen05_fence_after();
      tmem_ld_32x32b_x32(acc_addr(0, cb), ra);
      tmem_ld_32x32b_x32(acc_addr(0, cb) + 32, rb);
    }
    while (tile < p.total_tiles && cb < kBlocks) {
      int g, m_tile, n_tile;
      decode(tile, g, m_tile, n_tile);
      const uint32_t acc = it & 1u;
      // M-stacked tiles: column blocks 0,1 belong to rows m0 .., blocks 2,3 to rows m0 + 128 ..
      const int row0 = MSTACK ? m_tile * (2 * BM) + (cb >> 1) * BM + ew * 32 : m_tile * BM + ew * 32;
      int rmax = p.M - row0;            // valid rows of this warp's 32-row slab
      rmax = rmax < 0 ? 0 : (rmax > 32 ? 32 : rmax);
      const int n0 = MSTACK ? 0 : n_tile * BN;
      const float* bias = f_bias ? f_bias + static_cast<long long>(g) * p.N + n0 : nullptr;
      float* stat_buf = stat_smem + (it & 1u) * (4 * BN * 2);
      // next work item of this warp
      long long ntile = tile;
      int ncb = cb + kColSets;
      uint32_t nit = it;
      if (ncb >= kBlocks) { ncb = cset; ntile += gridDim.x; ++nit; }
      const bool last_of_tile = ntile != tile;
      const bool has_next = ntile < p.total_tiles;
      const int col0 = cb * 64;
      // stacked mode: this 64-column block belongs to sample gb at channel offset nb
      const int gb = (f_stack > 1) ? g + col0 / f_N : g;
      const int nb = MSTACK ? (cb & 1) * 64 : ((f_stack > 1) ? col0 % f_N : n0 + col0);
      const bool cols_ok = (f_stack > 1) ? (gb < p.G) : (nb < p.N);     // warp uniform

      tmem_ld_wait();                   // ra / rb of this item are in registers
      if (last_of_tile) {
        // every TMEM read of this accumulator stage by this warp is complete -> hand it back to the MMA warp early
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
      }
      auto prefetch = [&](uint32_t (&dst)[32], int half) {    // next item's half, once its registers are dead
        if (!has_next) return;
        if (last_of_tile && half == 0) {
          mbar_wait(tmem_full_bar(nit & 1u), (nit >> 1) & 1u);
          tcgen05_fence_after();
        }
        tmem_ld_32x32b_x32(acc_addr(nit, ncb) + 32 * half, dst);
      };
      if (f_out_f32) {
        // lane = output row (a Cout index in weight-gradient mode), 64 consecutive fp32 columns = 256 contiguous bytes
        float* dst = reinterpret_cast<float*>(p.y) + (static_cast<long long>(gb) * p.M + row0 + lane) * p.N + nb;
        const bool row_ok = cols_ok && static_cast<int>(lane) < rmax;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (row_ok && nb + j < p.N) *reinterpret_cast<uint4*>(dst + j) = make_uint4(ra[j], ra[j + 1], ra[j + 2], ra[j + 3]);
        prefetch(ra, 0);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (row_ok && nb + 32 + j < p.N)
            *reinterpret_cast<uint4*>(dst + 32 + j) = make_uint4(rb[j], rb[j + 1], rb[j + 2], rb[j + 3]);
        prefetch(rb, 1);
      } else if (!cols_ok) {
        prefetch(ra, 0);
        prefetch(rb, 1);
        if (p.stats) *reinterpret_cast<float4*>(stat_buf + (ew * BN + col0 + 2 * lane) * 2) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        if (bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n0 + col0 + j < p.N) ra[j] = __float_as_uint(__uint_as_float(ra[j]) + bias[col0 + j]);
            if (n0 + col0 + 32 + j < p.N) rb[j] = __float_as_uint(__uint_as_float(rb[j]) + bias[col0 + 32 + j]);
          }
        }
        // split (fp16x3 validation) mode uses both buffers per block: hi and lo halves of the fp32 accumulator
        const uint32_t buf = f_split ? my_out : my_out + (blk & 1u) * L::kOutBufBytes;
        const uint32_t buf_lo = my_out + L::kOutBufBytes;
        // the TMA store issued from this buffer two blocks ago (split: both previous stores) must have finished reading it
        if (lane == 0) {
          if (f_split) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
        auto stage_half = [&](const uint32_t (&r)[32], int half) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int q = half * 4 + q4;
            const uint32_t* src = &r[q4 * 8];
            __half2 h0 = __floats2half2_rn(__uint_as_float(src[0]), __uint_as_float(src[1]));
            __half2 h1 = __floats2half2_rn(__uint_as_float(src[2]), __uint_as_float(src[3]));
            __half2 h2 = __floats2half2_rn(__uint_as_float(src[4]), __uint_as_float(src[5]));
            __half2 h3 = __floats2half2_rn(__uint_as_float(src[6]), __uint_as_float(src[7]));
            const uint32_t off = lane * 128u + ((static_cast<uint32_t>(q) ^ (lane & 7u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + off), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                         "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                         "r"(*reinterpret_cast<uint32_t*>(&h3))
                         : "memory");
            if (f_split) {
              const __half2 hh[4] = {h0, h1, h2, h3};
              uint32_t lo[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(hh[j]);
                __half2 l = __floats2half2_rn(__uint_as_float(src[2 * j]) - f.x, __uint_as_float(src[2 * j + 1]) - f.y);
                lo[j] = *reinterpret_cast<uint32_t*>(&l);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf_lo + off), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]),
                           "r"(lo[3])
                           : "memory");
            }
          }
        };
        stage_half(ra, 0);
        prefetch(ra, 0);
        stage_half(rb, 1);
        prefetch(rb, 1);
        fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (EPI == EPI_STORE_STATS && lane == 0 && rmax > 0) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmY)),
                       "r"(buf), "r"(nb), "r"(row0), "r"(gb)
                       : "memory");
          if (f_split)
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmY)),
                         "r"(buf_lo), "r"(p.N + nb), "r"(row0), "r"(gb)
                         : "memory");
        }
        if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (p.stats) {
          // lane j owns output channels col0 + 2j, 2j+1: column sums over the warp's valid rows, packed fp32x2 math
          unsigned long long sa = 0ull, sb = 0ull, qa = 0ull, qb = 0ull;   // (s0,s1) / (q0,q1), two chains for ILP
          auto acc2 = [&](uint32_t w, unsigned long long& s2, unsigned long long& q2, uint32_t off = 0) {
            float2 f = __half22float2(*reinterpret_cast<__half2*>(&w));
            if (f_split) {      // statistics of the full-precision value hi + lo
              uint32_t wl;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wl) : "r"(buf_lo + off));
              const float2 fl = __half22float2(*reinterpret_cast<__half2*>(&wl));
              f.x += fl.x; f.y += fl.y;
            }
            const unsigned long long f2 = *reinterpret_cast<const unsigned long long*>(&f);
            asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s2) : "l"(f2));
            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q2) : "l"(f2));
          };
          if (rmax == 32) {
#pragma unroll
            for (int r = 0; r < 32; r += 2) {
              uint32_t w0, w1;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(buf + sw_off[r & 7] + r * 128u));
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(buf + sw_off[(r + 1) & 7] + (r + 1) * 128u));
              acc2(w0, sa, qa, sw_off[r & 7] + r * 128u);
              acc2(w1, sb, qb, sw_off[(r + 1) & 7] + (r + 1) * 128u);
            }
          } else {
            for (int r = 0; r < rmax; ++r) {
              uint32_t w0;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0)
                           : "r"(buf + ((((lane >> 2) ^ (r & 7u)) << 4) + ((lane & 3u) << 2)) + r * 128u));
              acc2(w0, sa, qa, ((((lane >> 2) ^ (r & 7u)) << 4) + ((lane & 3u) << 2)) + r * 128u);
            }
          }
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sa) : "l"(sb));
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(qa) : "l"(qb));
          const float2 sf = *reinterpret_cast<float2*>(&sa), qf = *reinterpret_cast<float2*>(&qa);
          const float s0 = sf.x, s1 = sf.y, q0 = qf.x, q1 = qf.y;
          *reinterpret_cast<float4*>(stat_buf + (ew * BN + col0 + 2 * lane) * 2) = make_float4(s0, q0, s1, q1);
        }
        ++blk;
      }

      if (p.stats && last_of_tile) {
        // combine the 4 lane-quarter warps of THIS column set and emit one deterministic partial per (tile, channel)
        asm volatile("bar.sync %0, %1;" ::"r"(1 + cset), "n"(kSetThreads) : "memory");
        for (int jj = es; jj < 64 * ((kBlocks - cset + kColSets - 1) / kColSets); jj += kSetThreads) {
          const int j = (cset + (jj >> 6) * kColSets) * 64 + (jj & 63);        // this set's column blocks
          const int gj = (f_stack > 1) ? g + j / f_N : g;
          const int nj = MSTACK ? (j & 127) : ((f_stack > 1) ? j % f_N : n0 + j);
          const int tj = MSTACK ? 2 * m_tile + (j >> 7) : m_tile;            // 128-row statistics tile
          if (nj < p.N && gj < p.G && tj < p.stat_tiles) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              a += stat_buf[(w * BN + j) * 2 + 0];
              b += stat_buf[(w * BN + j) * 2 + 1];
            }
            float2* dst = reinterpret_cast<float2*>(p.stats) +
                          (static_cast<long long>(gj) * p.stat_tiles + tj) * p.N + nj;
            *dst = make_float2(a, b);
          }
        }
      }
      tile = ntile; cb = ncb; it = nit;
    }
    }
    // smem must stay valid until the last bulk stores have read it; global visibility at kernel end
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int BN, int EPI, int PLAIN, bool XF = false, bool NR = false, bool BR = false>
int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                  const GemmParams& p, cudaStream_t stream) {
  using L = SmemLayout<BN, EPI, XF && EPI == EPI_STORE_STATS, NR, BR>;
  static bool attr_set = false;
  if (!attr_set) {
    MAUV_CUDA(cudaFuncSetAttribute(gemm_f16_tc_kernel<BN, EPI, PLAIN, XF, NR, BR>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  long long grid = p.total_tiles < mauv_num_sms() ? p.total_tiles : mauv_num_sms();
  if (BR) grid -= grid % p.n_tiles;       // a CTA must meet one n-tile only (tile index step = grid, n-tile = tile % n_tiles)
  gemm_f16_tc_kernel<BN, EPI, PLAIN, XF, NR, BR><<<static_cast<unsigned>(grid), gemm_threads<EPI, XF>(), L::kTotal, stream>>>(tmA, tmB, tmY, tmR, p);
  MAUV_LAUNCH_CHECK("gemm_f16_tc_kernel");
  return MAUV_OK;
}

template <int BN, int EPI>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                const GemmParams& p, cudaStream_t stream) {
  // XF instances (BatchNorm + ReLU of the previous layer applied to the operand tiles in shared memory): the fused-BN tails
  // (plain, or K-concatenated with a downsample branch) and the second-moment contraction
  // resident weight tile (BR): fused-BN tails at BN = 256 whose [BN][K] tile fits 64 KB and whose n-tile count divides the grid
  static const bool br_on = [] { const char* e = getenv("MAUV_BRES"); return !(e && e[0] == '0'); }();
  bool br = false;
  if constexpr (EPI == EPI_FUSED_BN && BN == 256) {
    br = br_on && p.a_mode == 0 && p.k_blocks * (BN * BK * 2) <= 65536 && p.stack <= 1 && !p.split && !p.out_f32 && !p.mn && !p.gram &&
         !p.a_wrap_kb && !p.a_cwrap && !p.b_mod && !p.bias && p.a_batch_mul == 1 && (p.n_tiles == 1 || p.n_tiles == 2 || p.n_tiles == 4) &&
         p.total_tiles >= mauv_num_sms();
  }
  if (p.xf_ss) {
    if constexpr (EPI == EPI_FUSED_BN && BN >= 128) {
      const bool plain = p.stack <= 1 && !p.split && !p.out_f32 && !p.mn && !p.gram && !p.a2_kb && !p.a_wrap_kb && !p.a_cwrap &&
                         !p.b_mod && !p.bias && p.a_batch_mul == 1;
      if constexpr (BN == 256) {
        if (br && plain) return launch_gemm_t<BN, EPI, 1, true, false, true>(tmA, tmB, tmY, tmR, p, stream);
        if (br && !p.has_res) return launch_gemm_t<BN, EPI, 0, true, true, true>(tmA, tmB, tmY, tmR, p, stream);
      }
      if (plain) return launch_gemm_t<BN, EPI, 1, true>(tmA, tmB, tmY, tmR, p, stream);
      if (!p.has_res) return launch_gemm_t<BN, EPI, 0, true, true>(tmA, tmB, tmY, tmR, p, stream);
      return launch_gemm_t<BN, EPI, 0, true>(tmA, tmB, tmY, tmR, p, stream);
    } else if constexpr (EPI == EPI_STORE_STATS) {
      if (p.mn && p.gram && p.out_f32) return launch_gemm_t<BN, EPI, 0, true>(tmA, tmB, tmY, tmR, p, stream);
    }
    return mauv_set_error(MAUV_ERR_BAD_ARG, "gemm_f16_tc: no operand-transform instance for this mode (BN=%d EPI=%d)", BN, EPI);
  }
  // the compile-time-specialised instance for the hot inference shapes (store + statistics, fused BatchNorm epilogue)
  if constexpr (EPI == EPI_STORE_STATS || EPI == EPI_FUSED_BN) {
    const bool plain = p.stack <= 1 && !p.split && !p.out_f32 && !p.mn && !p.gram && !p.a2_kb && !p.a_wrap_kb && !p.a_cwrap &&
                       !p.b_mod && !p.bias && p.a_batch_mul == 1;
    if constexpr (EPI == EPI_FUSED_BN && BN == 256) {
      if (br && plain) return launch_gemm_t<BN, EPI, 1, false, false, true>(tmA, tmB, tmY, tmR, p, stream);
      if (br && !p.has_res && p.a2_kb) return launch_gemm_t<BN, EPI, 0, false, true, true>(tmA, tmB, tmY, tmR, p, stream);
    }
    if (plain) return launch_gemm_t<BN, EPI, 1>(tmA, tmB, tmY, tmR, p, stream);
    if constexpr (EPI == EPI_FUSED_BN && BN >= 128) {
      if (!p.has_res && p.a2_kb) return launch_gemm_t<BN, EPI, 0, false, true>(tmA, tmB, tmY, tmR, p, stream);
    }
    if constexpr (BN == 256 && EPI == EPI_STORE_STATS) {
      const bool stacked = p.stack == 4 && p.N == 64 && p.a_batch_mul == 0 && p.a_mode == 0 && !p.split && !p.out_f32 && !p.mn &&
                           !p.gram && !p.a2_kb && !p.a_wrap_kb && !p.a_cwrap && !p.b_mod && !p.bias;
      if (stacked) return launch_gemm_t<BN, EPI, 2>(tmA, tmB, tmY, tmR, p, stream);
    }
  }
  return launch_gemm_t<BN, EPI, 0>(tmA, tmB, tmY, tmR, p, stream);
}

int pick_bn(int N) {
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return 256;
}

int dispatch(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, cudaStream_t stream, int epi = 0,
             const void* residual = nullptr, const CUtensorMap* tmA2 = nullptr) {
  const int bn = (p.stack > 1 || epi == EPI_STATS_T) ? 256 : pick_bn(p.N);
  // output [G][M][N] fp16 written by TMA: box = 64 channels x 32 rows (one epilogue warp's slab)
  CUtensorMap tmY;
  const int n_out = (p.split || p.out_f32) ? 2 * p.N : p.N;      // out_f32: the (unused) map spans the same bytes
  if (int rc = make_tiled_map(&tmY, p.y, n_out, p.M, p.G, static_cast<int64_t>(p.M) * n_out, 32)) return rc;
  p.m_tiles = static_cast<int>(ceil_div_i64(p.M, BM));
  p.stat_tiles = p.m_tiles;
  p.n_tiles = static_cast<int>(ceil_div_i64(p.N, bn));
  p.total_tiles = static_cast<long long>(p.m_tiles) * p.n_tiles * p.G;
  // M-stacked tiles (PLAIN = 3): the N = 128 im2col convs (3x3, K = 1152) are bound by the operand stream into shared memory
  static const bool mstack_on = [] { const char* e = getenv("MAUV_MSTACK"); return !(e && e[0] == '0'); }();
  const bool mstack = mstack_on && epi == EPI_STORE_STATS && bn == 128 && p.a_mode == 1 && p.M >= 2 * BM && p.stack <= 1 && !p.split &&
                      !p.out_f32 && !p.mn && !p.gram && !p.a2_kb && !p.a_wrap_kb && !p.a_cwrap && !p.b_mod && !p.bias &&
                      p.a_batch_mul == 1 && !p.xf_ss && p.k_blocks >= 8;
  if (mstack) {
    p.m_tiles = static_cast<int>(ceil_div_i64(p.M, 2 * BM));       // 256-row pairs
    p.n_tiles = 1;
    p.total_tiles = static_cast<long long>(p.m_tiles) * p.G;
  }
  if (p.stack > 1) {
    p.n_tiles = 1;
    p.g_blocks = static_cast<int>(ceil_div_i64(p.G, p.stack));
    p.total_tiles = static_cast<long long>(p.m_tiles) * p.g_blocks;
  }
  if (p.total_tiles == 0) return MAUV_OK;
  MAUV_CHECK_ARG(p.total_tiles < (1LL << 31), "gemm_f16_tc: too many tiles (%lld)", p.total_tiles);
  CUtensorMap tmR = tmY;
  if (epi == EPI_FUSED_BN && residual) {
    if (int rc = make_tiled_map(&tmR, residual, p.N, p.M, p.G, static_cast<int64_t>(p.M) * p.N, 32)) return rc;
  }
  if (tmA2) tmR = *tmA2;      // K-concatenated A (no residual in that mode): the spare map slot carries the second tensor
  if (epi == EPI_STATS_ONLY) {
    switch (bn) {
      case 64: return launch_gemm<64, 1>(tmA, tmB, tmY, tmR, p, stream);
      case 128: return launch_gemm<128, 1>(tmA, tmB, tmY, tmR, p, stream);
      default: return launch_gemm<256, 1>(tmA, tmB, tmY, tmR, p, stream);
    }
  }
  if (epi == EPI_STATS_T) return launch_gemm<256, 3>(tmA, tmB, tmY, tmR, p, stream);
  if (epi == EPI_FUSED_BN) {
    switch (bn) {
      case 64: return launch_gemm<64, 2>(tmA, tmB, tmY, tmR, p, stream);
      case 128: return launch_gemm<128, 2>(tmA, tmB, tmY, tmR, p, stream);
      default: return launch_gemm<256, 2>(tmA, tmB, tmY, tmR, p, stream);
    }
  }
  if (mstack) return launch_gemm_t<256, 0, 3>(tmA, tmB, tmY, tmR, p, stream);
  switch (bn) {
    case 64: return launch_gemm<64, 0>(tmA, tmB, tmY, tmR, p, stream);
    case 128: return launch_gemm<128, 0>(tmA, tmB, tmY, tmR, p, stream);
    default: return launch_gemm<256, 0>(tmA, tmB, tmY, tmR, p, stream);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 convolution with Cin = Cout = 64 (ResNet layer1 conv2, and its data gradient): "padded stream"
// mode. The generic im2col path fetches the input once per filter tap (9 boxes of [128 px][64 ch] per tile): for N = 64
// that is 9x the input through L2 per 128x64 outputs and the kernel is bound by L2->SM traffic (~7 TB/s), not by HBM or
// the tensor cores. Here the tile is 128 consecutive positions of the zero-padded pixel stream (W+2 positions per image
// row: an im2col-mode tensor map whose bounding box spans [-1, W] delivers exactly that stream, OOB zero-filled), so that
// EVERY tap is a constant row shift of the same smem rows: one TMA box of 130 positions per filter row (3 per tile instead
// of 9) and the three horizontal taps are the same box read through smem descriptors that start 0, 1, 2 rows (128 B)
// later. The 9 x [64][64] weight blocks of a sample (72 KB) stay resident in shared memory across all of that sample's
// tiles. 2 of every W+2 output positions are padding and are dropped by the epilogue (97 % useful at W = 64).
// Epilogue: two sets of 4 warps, one per TMEM accumulator stage (even / odd tiles), so two tiles drain concurrently.
struct StreamParams {
  int G, imgs_per_sample, H, W, Wp;       // Wp = W + 2
  long long Mv;                            // padded positions per sample = imgs * H * Wp
  int m_tiles;                             // ceil(Mv / 128)
  long long total_tiles;
  int base_offset_mode;                    // 1: descriptor base_offset = start row mod 8
  __half* y;                               // [G*imgs][H][W][64]
  float* stats;                            // [G][m_tiles][64][2] or nullptr
  // optional input transform: the kernel is handed the RAW output of the previous conv and applies that layer's
  // BatchNorm + ReLU to the tile in shared memory (in_ss [G][64] (scale, shift)); padding positions stay zero. The
  // activated tensor is then never written to / re-read from HBM.
  const float2* in_ss;
};

constexpr int kStreamBox = 130;                       // positions per TMA box (128 + 2 for the horizontal taps)
constexpr int kStreamRegion = 17408;                  // 130 * 128 B rounded up to 1024
constexpr int kStreamBBytes = 9 * 8192;
constexpr int kStreamABytes = 3 * kStreamRegion;      // per stage
constexpr int kStreamStages = 2;
constexpr int kStreamOutBytes = 8 * 4096;             // per epilogue warp: 32 rows x 128 B staging (for the statistics)
constexpr int kStreamStatBytes = 2 * 4 * 64 * 2 * 4 + 512;   // + 512: padding flags of the input transform
constexpr int kStreamSmem = 1024 + kStreamBBytes + kStreamStages * kStreamABytes + kStreamOutBytes + kStreamStatBytes + 256;

__global__ void __launch_bounds__(512, 1)
conv3x3_c64_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const StreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_base = smem_base;
  const uint32_t a_base = b_base + kStreamBBytes;
  const uint32_t out_base = a_base + kStreamStages * kStreamABytes;
  float* stat_smem = reinterpret_cast<float*>(smem_gen + kStreamBBytes + kStreamStages * kStreamABytes + kStreamOutBytes);
  const uint32_t bar_base = out_base + kStreamOutBytes + kStreamStatBytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (2 + s); };
  const uint32_t b_full = bar_base + 8u * 4, b_empty = bar_base + 8u * 5;
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (6 + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (8 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * 10;
  auto a_ready = [&](int s) { return bar_base + 8u * (12 + s); };     // input transform done (4 warps arrive)
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + kStreamBBytes + kStreamStages * kStreamABytes + kStreamOutBytes + kStreamStatBytes + 8 * 10);

  const int warp = threadIdx.x >> 5;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kStreamStages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 4);
      mbar_init(a_ready(a), 4);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<128>(tmem_ptr_addr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const long long hwp = static_cast<long long>(p.H) * p.Wp;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, cur_g = -1;
      uint32_t phase = 0, b_loads = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int g = static_cast<int>(tile / p.m_tiles);
        const int m_tile = static_cast<int>(tile - static_cast<long long>(g) * p.m_tiles);
        if (g != cur_g) {       // new sample: its 9 weight blocks replace the resident ones once the MMAs reading them retired
          if (b_loads > 0) mbar_wait(b_empty, (b_loads - 1) & 1u);
          mbar_expect_tx(b_full, kStreamBBytes);
          for (int tap = 0; tap < 9; ++tap) tma_load_3d(b_base + tap * 8192, &tmB, b_full, tap * 64, 0, g);
          ++b_loads;
          cur_g = g;
        }
        const long long v0 = static_cast<long long>(m_tile) * 128;
        const int img = static_cast<int>(v0 / hwp);
        const int rem = static_cast<int>(v0 - img * hwp);
        const int pr = rem / p.Wp, j0 = rem - pr * p.Wp;
        mbar_wait(a_empty(stage), phase ^ 1u);
        mbar_expect_tx(a_full(stage), 3u * kStreamBox * 128u);
        for (int r = 0; r < 3; ++r)
          tma_load_im2col_4d(a_base + stage * kStreamABytes + r * kStreamRegion, &tmA, a_full(stage), 0, j0 - 1, pr - 1,
                             g * p.imgs_per_sample + img, 0, static_cast<uint16_t>(r));
        if (++stage == kStreamStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(128, 64);
      int stage = 0, cur_g = -1;
      uint32_t phase = 0, it = 0, b_uses = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int g = static_cast<int>(tile / p.m_tiles);
        if (g != cur_g) {
          if (cur_g >= 0) umma_commit(b_empty);        // arrives when every MMA issued so far (old weights) has completed
          mbar_wait(b_full, b_uses & 1u);
          ++b_uses;
          cur_g = g;
        }
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
        mbar_wait(p.in_ss ? a_ready(stage) : a_full(stage), phase);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 64;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, sx = tap - r * 3;
          const uint32_t a_addr = a_base + stage * kStreamABytes + r * kStreamRegion + sx * 128;
          uint64_t a_desc = umma_smem_desc_sw128(a_addr);
          if (p.base_offset_mode) a_desc |= static_cast<uint64_t>(sx) << 49;    // start row within the 8-row swizzle atom
          const uint64_t b_desc = umma_smem_desc_sw128(b_base + tap * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (tap | k) != 0 ? 1u : 0u);
        }
        umma_commit(a_empty(stage));
        umma_commit(tmem_full_bar(acc));
        if (++stage == kStreamStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 12) {
    // ===================== input transform (optional): a = relu(y * scale + shift) in place, padding stays zero ==========
    // 128 threads. Thread t owns the LOGICAL 16-byte chunk lc = t & 7 (channels 8*lc .. 8*lc+7: its 8 (scale, shift) pairs
    // live in registers) of the rows (t >> 3) + 16*i of the 3 x 130-row boxes; the physical chunk is lc ^ (row & 7)
    // (128B swizzle). Which rows are conv padding is worked out once per tile (3 rows per thread) into shared memory.
    if (p.in_ss) {
      const int t = threadIdx.x - 384;          // 0..127
      const int lc = t & 7;
      uint8_t* inside_flag = reinterpret_cast<uint8_t*>(stat_smem) + kStreamStatBytes - 512;   // 3*130 flags (spare tail)
      float sc[8], sh[8];
      int stage = 0, cur_g = -1;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int g = static_cast<int>(tile / p.m_tiles);
        const int m_tile = static_cast<int>(tile - static_cast<long long>(g) * p.m_tiles);
        if (g != cur_g) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float2 v2 = p.in_ss[static_cast<long long>(g) * 64 + lc * 8 + q];
            sc[q] = v2.x; sh[q] = v2.y;
          }
          cur_g = g;
        }
        const int v0 = m_tile * 128;
        const int hwp_i = static_cast<int>(hwp);
        for (int i = t; i < 3 * kStreamBox; i += 128) {
          const int r = i / kStreamBox, row = i - r * kStreamBox;
          const int v = v0 + row;
          const int img = v / hwp_i;
          const int r2 = v - img * hwp_i;
          const int pr = r2 / p.Wp, j = r2 - pr * p.Wp;
          const int h = pr + r - 1;
          inside_flag[i] = (j >= 1 && j <= p.W && h >= 0 && h < p.H) ? 1 : 0;       // else: zero-filled conv padding
        }
        asm volatile("bar.sync 3, 128;" ::: "memory");
        mbar_wait(a_full(stage), phase);
        for (int i = t >> 3; i < 3 * kStreamBox; i += 16) {
          const int r = i / kStreamBox, row = i - r * kStreamBox;
          const uint32_t addr = a_base + stage * kStreamABytes + r * kStreamRegion + row * 128 + ((lc ^ (row & 7)) << 4);
          uint32_t wv[4];
          if (inside_flag[i]) {
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(wv[0]), "=r"(wv[1]), "=r"(wv[2]), "=r"(wv[3]) : "r"(addr));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __half22float2(*reinterpret_cast<__half2*>(&wv[q]));
              __half2 o = __floats2half2_rn(fmaxf(fmaf(f.x, sc[2 * q], sh[2 * q]), 0.f),
                                            fmaxf(fmaf(f.y, sc[2 * q + 1], sh[2 * q + 1]), 0.f));
              wv[q] = *reinterpret_cast<uint32_t*>(&o);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]) : "memory");
          }       // padding rows were zero-filled by TMA and stay zero
        }
        fence_proxy_async_smem();        // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(a_ready(stage));
        asm volatile("bar.sync 3, 128;" ::: "memory");      // flags are rewritten for the next tile
        if (++stage == kStreamStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    const int set = (warp - 4) >> 2;          // which accumulator stage (= tile parity) this warp drains
    const int ew = warp & 3;                  // TMEM lane quarter
    const uint32_t lane = lane_id();
    const uint32_t my_out = out_base + (warp - 4) * 4096;
    float* stat_buf = stat_smem + set * (4 * 64 * 2);
    const int et = threadIdx.x - 128 - set * 128;   // 0..127 inside the set
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if (static_cast<int>(it & 1u) != set) continue;
      const int g = static_cast<int>(tile / p.m_tiles);
      const int m_tile = static_cast<int>(tile - static_cast<long long>(g) * p.m_tiles);
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(tmem_full_bar(set), acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + set * 64 + (static_cast<uint32_t>(ew * 32) << 16);
      uint32_t ra[32], rb[32];
      tmem_ld_32x32b_x32(taddr, ra);
      tmem_ld_32x32b_x32(taddr + 32, rb);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(set));
      // this lane's position in the padded stream -> dense output pixel (or padding / past the end)
      const long long v = static_cast<long long>(m_tile) * 128 + ew * 32 + lane;
      bool valid = v < p.Mv;
      long long dense = 0;
      if (valid) {
        const int img = static_cast<int>(v / hwp);
        const int rem = static_cast<int>(v - img * hwp);
        const int pr = rem / p.Wp, u = rem - pr * p.Wp;
        valid = u < p.W;
        dense = ((static_cast<long long>(g) * p.imgs_per_sample + img) * p.H + pr) * p.W + u;
      }
      uint4* dst = reinterpret_cast<uint4*>(p.y + dense * 64);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t* src = (q < 4) ? &ra[q * 8] : &rb[(q - 4) * 8];
        uint4 o = make_uint4(0, 0, 0, 0);
        if (valid) {
          __half2 h0 = __floats2half2_rn(__uint_as_float(src[0]), __uint_as_float(src[1]));
          __half2 h1 = __floats2half2_rn(__uint_as_float(src[2]), __uint_as_float(src[3]));
          __half2 h2 = __floats2half2_rn(__uint_as_float(src[4]), __uint_as_float(src[5]));
          __half2 h3 = __floats2half2_rn(__uint_as_float(src[6]), __uint_as_float(src[7]));
          o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                         *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
          dst[q] = o;                                                    // 128 contiguous bytes per lane
        }
        if (p.stats) {
          const uint32_t off = lane * 128u + ((static_cast<uint32_t>(q) ^ (lane & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_out + off), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        }
      }
      if (p.stats) {
        __syncwarp();
        // lane j owns channels 2j, 2j+1: sums over the warp's 32 rows of the staged fp16 values (padding rows are zero)
        unsigned long long sa = 0ull, qa = 0ull;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          uint32_t w0;
          const uint32_t off = ((((lane >> 2) ^ (static_cast<uint32_t>(r) & 7u)) << 4) + ((lane & 3u) << 2)) + r * 128u;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(my_out + off));
          const float2 f = __half22float2(*reinterpret_cast<__half2*>(&w0));
          const unsigned long long f2 = *reinterpret_cast<const unsigned long long*>(&f);
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sa) : "l"(f2));
          asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(qa) : "l"(f2));
        }
        const float2 sf = *reinterpret_cast<float2*>(&sa), qf = *reinterpret_cast<float2*>(&qa);
        *reinterpret_cast<float4*>(stat_buf + (ew * 64 + 2 * lane) * 2) = make_float4(sf.x, qf.x, sf.y, qf.y);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
        if (et < 64) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            a += stat_buf[(w * 64 + et) * 2 + 0];
            b += stat_buf[(w * 64 + et) * 2 + 1];
          }
          reinterpret_cast<float2*>(p.stats)[(static_cast<long long>(g) * p.m_tiles + m_tile) * 64 + et] = make_float2(a, b);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

}  // namespace

extern "C" {

// Number of 128-row tiles per sample: the leading dimension of the BN partial-statistics buffer.
int mauv_gemm_m_tiles(long long M) { return static_cast<int>(ceil_div_i64(M, BM)); }
// Row tiles of the statistics buffer written by mauv_gemm_bn_f16 mode 1 (256 pixels per tile).
int mauv_gemm_bn_stats_tiles(long long M) { return static_cast<int>(ceil_div_i64(M, 256)); }

int mauv_gemm_f16(const void* a, long long a_sample_stride, const void* w, const void* bias,
                  void* y, float* stats_partial, int G, long long M, int N, int K, void* stream) {
  MAUV_CHECK_ARG(a && w && y, "mauv_gemm_f16: null pointer");
  MAUV_CHECK_ARG(G >= 1 && M >= 1 && N >= 8 && K >= 8, "mauv_gemm_f16: bad shape G=%d M=%lld N=%d K=%d", G, M, N, K);
  MAUV_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "mauv_gemm_f16: K and N must be multiples of 8 (K=%d N=%d)", K, N);
  MAUV_CHECK_ARG(a_sample_stride % 8 == 0, "mauv_gemm_f16: sample stride must be a multiple of 8 elements");
  MAUV_CHECK_ARG((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(y) & 15) == 0, "mauv_gemm_f16: pointers must be 16-byte aligned");
  if (int rc = load_driver_entry_points()) return rc;
  CUtensorMap tmA, tmB;
  const bool shared_a = (a_sample_stride == 0);
  if (int rc = make_tiled_map(&tmA, a, K, M, shared_a ? 1 : G, shared_a ? M * K : a_sample_stride, BM)) return rc;
  // shared A (stem) with narrow N: stack 256/N samples along the tile's N dimension so one A tile feeds them all
  const int stack = (shared_a && G > 1 && !bias && N <= 128 && 256 % N == 0) ? 256 / N : 1;
  if (stack > 1) {
    if (int rc = make_tiled_map(&tmB, w, K, static_cast<int64_t>(G) * N, 1, static_cast<int64_t>(G) * N * K, 256)) return rc;
  } else {
    if (int rc = make_tiled_map(&tmB, w, K, N, G, static_cast<int64_t>(N) * K, pick_bn(N))) return rc;
  }
  GemmParams p{};
  p.stack = stack;
  p.M = static_cast<int>(M);
  p.N = N;
  p.k_blocks = static_cast<int>(ceil_div_i64(K, BK));
  p.G = G;
  p.a_mode = 0;
  p.a_batch_mul = shared_a ? 0 : 1;
  p.y = static_cast<__half*>(y);
  p.stats = stats_partial;
  p.bias = static_cast<const float*>(bias);
  return dispatch(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

// 3x3 / stride 1 / pad 1, Cin = Cout = 64: padded-stream kernel (see conv3x3_c64_stream_kernel). stats_partial:
// [G][mauv_conv3x3_c64_tiles(...)][64][2].
int mauv_conv3x3_c64_tiles(int imgs_per_sample, int H, int W) {
  return static_cast<int>(ceil_div_i64(static_cast<long long>(imgs_per_sample) * H * (W + 2), 128));
}

int mauv_conv3x3_c64_f16(const void* x, const void* w, void* y, float* stats_partial, const float* in_scale_shift, int G,
                         int imgs_per_sample, int H, int W, void* stream) {
  MAUV_CHECK_ARG(x && w && y && G >= 1 && imgs_per_sample >= 1 && H >= 1 && W >= 1 && W <= 254, "mauv_conv3x3_c64_f16: bad argument");
  MAUV_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(y) & 15) == 0, "mauv_conv3x3_c64_f16: pointers must be 16-byte aligned");
  if (int rc = load_driver_entry_points()) return rc;
  const long long N = static_cast<long long>(G) * imgs_per_sample;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
    cuuint64_t strides[3] = {128, static_cast<cuuint64_t>(128) * W, static_cast<cuuint64_t>(128) * W * H};
    int lower[2] = {-1, -1};
    int upper[2] = {1, -1};          // W: positions -1 .. W (the zero-padded row); H: the usual 3-row filter window
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_im2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, lower, upper,
                                 64, kStreamBox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return mauv_set_error(MAUV_ERR_DRIVER, "cuTensorMapEncodeIm2col (padded stream) failed (%d) W=%d H=%d N=%lld", (int)r, W, H, N);
  }
  if (int rc = make_tiled_map(&tmB, w, 576, 64, G, 64 * 576, 64)) return rc;
  StreamParams p{};
  p.G = G; p.imgs_per_sample = imgs_per_sample; p.H = H; p.W = W; p.Wp = W + 2;
  p.Mv = static_cast<long long>(imgs_per_sample) * H * p.Wp;
  p.m_tiles = mauv_conv3x3_c64_tiles(imgs_per_sample, H, W);
  p.total_tiles = static_cast<long long>(p.m_tiles) * G;
  // measured on B200: the 128B swizzle is a function of the absolute shared-memory address, so a descriptor may start on any
  // 128-byte row of a TMA-written tile with base_offset = 0 (base_offset = start row mod 8 gives wrong results)
  p.base_offset_mode = 0;
  p.y = static_cast<__half*>(y);
  p.stats = stats_partial;
  p.in_ss = reinterpret_cast<const float2*>(in_scale_shift);
  static bool attr_set = false;
  if (!attr_set) {
    MAUV_CUDA(cudaFuncSetAttribute(conv3x3_c64_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
    attr_set = true;
  }
  const long long grid = p.total_tiles < mauv_num_sms() ? p.total_tiles : mauv_num_sms();
  conv3x3_c64_stream_kernel<<<static_cast<unsigned>(grid), 512, kStreamSmem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  MAUV_LAUNCH_CHECK("conv3x3_c64_stream_kernel");
  return MAUV_OK;
}

// mauv_gemm_f16 whose [N][K] operand is shared by groups of batches: y[g] = a[g] * w[g % w_batches]^T. Used by the stem's
// weight gradient (dY_s^T chunks against the chunks of the ONE im2col matrix all samples share): one launch for all samples.
int mauv_gemm_wmod_f16(const void* a, const void* w, int w_batches, void* y, int out_f32, int G, long long M, int N, int K,
                       void* stream) {
  MAUV_CHECK_ARG(a && w && y && w_batches >= 1 && G >= 1 && M >= 1, "mauv_gemm_wmod_f16: bad argument");
  MAUV_CHECK_ARG(N >= 8 && K >= 8 && K % 8 == 0 && N % 8 == 0, "mauv_gemm_wmod_f16: K and N must be multiples of 8 (K=%d N=%d)", K, N);
  MAUV_CHECK_ARG((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(y) & 15) == 0, "mauv_gemm_wmod_f16: pointers must be 16-byte aligned");
  if (int rc = load_driver_entry_points()) return rc;
  CUtensorMap tmA, tmB;
  if (int rc = make_tiled_map(&tmA, a, K, M, G, M * K, BM)) return rc;
  if (int rc = make_tiled_map(&tmB, w, K, N, w_batches, static_cast<int64_t>(N) * K, pick_bn(N))) return rc;
  GemmParams p{};
  p.stack = 1;
  p.M = static_cast<int>(M);
  p.N = N;
  p.k_blocks = static_cast<int>(ceil_div_i64(K, BK));
  p.G = G;
  p.a_mode = 0;
  p.a_batch_mul = 1;
  p.b_mod = w_batches;
  p.out_f32 = out_f32 ? 1 : 0;
  p.y = static_cast<__half*>(y);
  return dispatch(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

// ---- weight gradient of the grouped conv, straight from NHWC operands ---------------------------------------------
// dw[(g, chunk)][co][(r, s, c)] = sum over the chunk's output pixels of dy[pixel][co] * x[pixel shifted by tap (r,s)][c]
// dw is FP32 ([G*splits][Cout][kh*kw*Cin] floats): the accumulator is stored unrounded.
static int wgrad_impl(const void* dy, const void* x, void* dw, int G, int splits, int imgs_per_sample, int H, int W, int Cin,
                      int Cout, int kh, int kw, int stride, int pad, const float* xf_ss, float* xf_colsum, void* stream) {
  MAUV_CHECK_ARG(dy && x && dw && G >= 1 && splits >= 1, "mauv_wgrad_f16: bad argument");
  MAUV_CHECK_ARG(Cin % 64 == 0 && Cout % 8 == 0, "mauv_wgrad_f16: Cin must be a multiple of 64 and Cout of 8 (Cin=%d Cout=%d)", Cin, Cout);
  MAUV_CHECK_ARG((reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(dw) & 15) == 0, "mauv_wgrad_f16: pointers must be 16-byte aligned");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  const long long Mg = static_cast<long long>(imgs_per_sample) * Ho * Wo;
  MAUV_CHECK_ARG(Mg % splits == 0 && (Mg / splits) % 64 == 0,
                 "mauv_wgrad_f16: pixels per chunk (%lld / %d) must be a multiple of 64", Mg, splits);
  const long long rows = static_cast<long long>(G) * Mg;
  MAUV_CHECK_ARG(rows < (1LL << 31), "mauv_wgrad_f16: too many pixels");
  if (int rc = load_driver_entry_points()) return rc;
  const bool plain = (kh == 1 && kw == 1 && stride == 1 && pad == 0);
  const int K = kh * kw * Cin;
  CUtensorMap tmA, tmB;
  {   // row-major [pixels][channels] seen as (channels, pixels, 1), box 64 x 64
    auto rowmajor = [&](CUtensorMap* tm, const void* base, int64_t cols) -> int {
      cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), 1};
      cuuint64_t strides[2] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(cols) * 2 * static_cast<cuuint64_t>(rows)};
      cuuint32_t box[3] = {64, 64, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return mauv_set_error(MAUV_ERR_DRIVER, "cuTensorMapEncodeTiled (wgrad operand) failed (%d)", (int)r);
      return MAUV_OK;
    };
    if (int rc = rowmajor(&tmA, dy, Cout)) return rc;
    if (plain) {
      if (int rc = rowmajor(&tmB, x, Cin)) return rc;
    } else {
      if (int rc = make_im2col_map(&tmB, x, Cin, W, H, static_cast<int64_t>(G) * imgs_per_sample, kh, kw, stride, pad, 64)) return rc;
    }
  }
  GemmParams p{};
  p.stack = 1;
  p.M = Cout;
  p.N = K;
  p.G = G * splits;
  p.chunk = Mg / splits;
  p.k_blocks = static_cast<int>(p.chunk / 64);
  p.a_mode = 0;
  p.a_batch_mul = 1;
  p.mn = 1;
  // second moments (dy and x are one tensor): load the boxes once. Needs one n-tile holding all K columns (K <= 256) and the
  // M tile's 128 channels inside it
  p.gram = (dy == x && plain && K <= 256 && Cout == K && (K % 128 == 0 || K == 64)) ? 1 : 0;
  p.b_im2col = plain ? 0 : 1;
  p.cin = Cin;
  p.Wo = Wo; p.Ho = Ho; p.imgs_per_sample = imgs_per_sample; p.stride = stride; p.pad = pad; p.kw = kw;
  p.y = static_cast<__half*>(dw);
  p.out_f32 = 1;          // fp32 partial sums: a correlated dy*x sum over thousands of pixels can exceed fp16's 65504
  p.stats = nullptr;
  p.bias = nullptr;
  if (xf_ss) {
    MAUV_CHECK_ARG(p.gram, "mauv_gram_bn_f16: shape not eligible for the second-moment mode (K=%d must be 64, 128 or 256)", K);
    p.xf_ss = reinterpret_cast<const float2*>(xf_ss);
    p.xf_K = K;
    p.xf_splits = splits;
    p.xf_colsum = xf_colsum;
  }
  return dispatch(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

int mauv_wgrad_f16(const void* dy, const void* x, void* dw, int G, int splits, int imgs_per_sample, int H, int W, int Cin,
                   int Cout, int kh, int kw, int stride, int pad, void* stream) {
  return wgrad_impl(dy, x, dw, G, splits, imgs_per_sample, H, W, Cin, Cout, kh, kw, stride, pad, nullptr, nullptr, stream);
}

// Second moments of a = relu(y * scale + shift) WITHOUT materialising a: y [G][M][K] is the raw output of the previous conv,
// scale_shift [G][K][2] that layer's BatchNorm; the transform runs on the TMA-loaded tiles in shared memory.
// gram_partial [G*splits][K][K] fp32 (a^T a per pixel chunk), colsum_partial [G*splits][K] fp32 (column sums of a per chunk):
// the inputs of mauv_bn_stats_from_gram for the 1x1 conv that consumes a. K in {64, 128, 256}, (M / splits) % 64 == 0.
int mauv_gram_bn_f16(const void* y, const float* scale_shift, float* gram_partial, float* colsum_partial, int G, int splits,
                     long long M, int K, void* stream) {
  MAUV_CHECK_ARG(y && scale_shift && gram_partial && colsum_partial, "mauv_gram_bn_f16: null pointer");
  MAUV_CHECK_ARG(M >= 1 && M < (1LL << 31), "mauv_gram_bn_f16: bad M");
  return wgrad_impl(y, y, gram_partial, G, splits, 1, static_cast<int>(M), 1, K, K, 1, 1, 1, 0, scale_shift, colsum_partial, stream);
}

// ---- fp16x3 validation mode -------------------------------------------------------------------------------------
// a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with a = a_hi + a_lo, w = w_hi + w_lo (fp16 pairs, ~2^-22 relative): the
// three terms are ONE contraction over a K-concatenated operand pair A' = [a_hi | a_lo | a_hi], W' = [w_hi | w_hi | w_lo].
// Activations are stored as (hi | lo) pairs ([.., 2C]); the third block re-reads the first (A's coordinate wraps), the
// weights are sampled straight into the 3K layout, and the epilogue writes the fp32 accumulator as a (hi | lo) pair.
// Same tcgen05 kernel, ~3x the MMA work and bytes: this mode exists to check the engine end to end at ~fp32 accuracy.
int mauv_gemm_x3_f16(const void* a2, long long a_sample_stride, const void* w3, void* y2, float* stats_partial, int G,
                     long long M, int N, int K, void* stream) {
  // a2: [G][M][2K] (hi | lo), K % 64 == 0; w3: [G][N][3K]; y2: [G][M][2N] (hi | lo)
  MAUV_CHECK_ARG(a2 && w3 && y2, "mauv_gemm_x3_f16: null pointer");
  MAUV_CHECK_ARG(G >= 1 && M >= 1 && N >= 8 && N % 8 == 0 && K >= 64 && K % 64 == 0, "mauv_gemm_x3_f16: bad shape");
  if (int rc = load_driver_entry_points()) return rc;
  CUtensorMap tmA, tmB;
  const bool shared_a = (a_sample_stride == 0);
  if (int rc = make_tiled_map(&tmA, a2, 2 * K, M, shared_a ? 1 : G, shared_a ? M * 2 * K : a_sample_stride, BM)) return rc;
  if (int rc = make_tiled_map(&tmB, w3, 3 * K, N, G, static_cast<int64_t>(N) * 3 * K, pick_bn(N))) return rc;
  GemmParams p{};
  p.stack = 1;
  p.M = static_cast<int>(M);
  p.N = N;
  p.k_blocks = 3 * K / BK;
  p.a_wrap_kb = 2 * K / BK;
  p.split = 1;
  p.G = G;
  p.a_mode = 0;
  p.a_batch_mul = shared_a ? 0 : 1;
  p.y = static_cast<__half*>(y2);
  p.stats = stats_partial;
  return dispatch(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

int mauv_conv2d_im2col_x3_f16(const void* x2, const void* w3, void* y2, float* stats_partial, int G, int imgs_per_sample,
                              int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, void* stream) {
  // x2: [G*imgs][H][W][2*Cin] (hi | lo); w3: [G][Cout][kh*kw*3*Cin] per tap (hi | hi | lo); y2: [..][Ho][Wo][2*Cout]
  MAUV_CHECK_ARG(x2 && w3 && y2 && Cin % 64 == 0 && Cout % 8 == 0, "mauv_conv2d_im2col_x3_f16: bad argument");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  if (int rc = load_driver_entry_points()) return rc;
  CUtensorMap tmA, tmB;
  if (int rc = make_im2col_map(&tmA, x2, 2 * Cin, W, H, static_cast<int64_t>(G) * imgs_per_sample, kh, kw, stride, pad)) return rc;
  const int K3 = kh * kw * 3 * Cin;
  if (int rc = make_tiled_map(&tmB, w3, K3, Cout, G, static_cast<int64_t>(Cout) * K3, pick_bn(Cout))) return rc;
  GemmParams p{};
  p.stack = 1;
  p.M = imgs_per_sample * Ho * Wo;
  p.N = Cout;
  p.k_blocks = K3 / BK;
  p.G = G;
  p.a_mode = 1;
  p.a_batch_mul = 1;
  p.Wo = Wo; p.Ho = Ho; p.imgs_per_sample = imgs_per_sample;
  p.stride = stride; p.pad = pad; p.kw = kw;
  p.c_blocks = 3 * Cin / BK;
  p.a_cwrap = 2 * Cin / BK;
  p.split = 1;
  p.y = static_cast<__half*>(y2);
  p.stats = stats_partial;
  return dispatch(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

// Recompute scheme for the bottleneck's last 1x1 conv (HBM-write bound): mode 1 = statistics only (y may be NULL),
// mode 2 = out = relu?(A*W^T * scale + shift [+ residual]) written straight to `y`; the raw conv output never
// exists in HBM. scale_shift: [G][N][2] from mauv_bn_finalize; residual: [G][M][N] fp16 or NULL. N % 64 == 0.
static int gemm_bn_impl(const void* a, const void* w, void* y, float* stats_partial, const float* scale_shift,
                        const void* residual, int relu, int mode, int G, long long M, int N, int K, const float* a_scale_shift,
                        void* stream) {
  MAUV_CHECK_ARG(a && w, "mauv_gemm_bn_f16: null pointer");
  MAUV_CHECK_ARG(mode == EPI_STATS_ONLY || mode == EPI_FUSED_BN, "mauv_gemm_bn_f16: mode must be 1 (stats) or 2 (fused)");
  MAUV_CHECK_ARG(G >= 1 && M >= 1 && N >= 64 && N % 64 == 0 && K >= 8 && K % 8 == 0, "mauv_gemm_bn_f16: bad shape G=%d M=%lld N=%d K=%d", G, M, N, K);
  MAUV_CHECK_ARG(N % pick_bn(N) == 0, "mauv_gemm_bn_f16: N must be 64, 128 or a multiple of 256 (got %d)", N);
  MAUV_CHECK_ARG(mode != EPI_STATS_ONLY || stats_partial, "mauv_gemm_bn_f16: stats buffer required in mode 1");
  MAUV_CHECK_ARG(mode != EPI_FUSED_BN || (y && scale_shift), "mauv_gemm_bn_f16: y and scale_shift required in mode 2");
  if (int rc = load_driver_entry_points()) return rc;
  CUtensorMap tmA, tmB;
  if (mode == EPI_STATS_ONLY) {
    // statistics pass: swap the operand roles (MMA-M = channels, MMA-N = pixels), see EPI_STATS_T.
    // stats layout: [G][ceil(M/256)][N] float2 (mauv_gemm_bn_stats_tiles(M) pixel tiles)
    if (int rc = make_tiled_map(&tmA, w, K, N, G, static_cast<int64_t>(N) * K, BM)) return rc;
    if (int rc = make_tiled_map(&tmB, a, K, M, G, M * K, 256)) return rc;
    GemmParams q{};
    q.stack = 1;
    q.M = N;                               // channels
    q.N = static_cast<int>(M);             // pixels
    q.k_blocks = static_cast<int>(ceil_div_i64(K, BK));
    q.G = G;
    q.a_mode = 0;
    q.a_batch_mul = 1;
    q.y = static_cast<__half*>(const_cast<void*>(a));   // never written
    q.stats = stats_partial;
    // the (unused) output map must still be encodable: describe it over A's bytes with legal extents
    CUtensorMap tmY;
    if (int rc = make_tiled_map(&tmY, a, K, M, G, M * K, 32)) return rc;
    q.m_tiles = static_cast<int>(ceil_div_i64(q.M, BM));
    q.n_tiles = static_cast<int>(ceil_div_i64(q.N, 256));
    q.total_tiles = static_cast<long long>(q.m_tiles) * q.n_tiles * q.G;
    MAUV_CHECK_ARG(q.total_tiles < (1LL << 31), "mauv_gemm_bn_f16: too many tiles (%lld)", q.total_tiles);
    return launch_gemm<256, 3>(tmA, tmB, tmY, tmY, q, static_cast<cudaStream_t>(stream));
  }
  if (int rc = make_tiled_map(&tmA, a, K, M, G, M * K, BM)) return rc;
  if (int rc = make_tiled_map(&tmB, w, K, N, G, static_cast<int64_t>(N) * K, pick_bn(N))) return rc;
  GemmParams p{};
  p.stack = 1;
  p.M = static_cast<int>(M);
  p.N = N;
  p.k_blocks = static_cast<int>(ceil_div_i64(K, BK));
  p.G = G;
  p.a_mode = 0;
  p.a_batch_mul = 1;
  // in mode 1 nothing is stored; the output map still needs a valid base address -> reuse A's
  p.y = static_cast<__half*>(mode == EPI_FUSED_BN ? y : const_cast<void*>(a));
  p.stats = mode == EPI_STATS_ONLY ? stats_partial : nullptr;
  p.bias = nullptr;
  p.ss = reinterpret_cast<const float2*>(scale_shift);
  p.has_res = residual != nullptr;
  p.relu = relu;
  if (a_scale_shift) {
    MAUV_CHECK_ARG(mode == EPI_FUSED_BN && K % 64 == 0 && K <= 256 && N >= 128,
                   "mauv_gemm_bn_xf_f16: needs K in {64, 128, 192, 256} and N >= 128 (K=%d N=%d)", K, N);
    p.xf_ss = reinterpret_cast<const float2*>(a_scale_shift);
    p.xf_K = K;
    p.xf_kb = p.k_blocks;
  }
  return dispatch(tmA, tmB, p, static_cast<cudaStream_t>(stream), mode, residual);
}

int mauv_gemm_bn_f16(const void* a, const void* w, void* y, float* stats_partial, const float* scale_shift,
                     const void* residual, int relu, int mode, int G, long long M, int N, int K, void* stream) {
  return gemm_bn_impl(a, w, y, stats_partial, scale_shift, residual, relu, mode, G, M, N, K, nullptr, stream);
}

// mauv_gemm_bn_f16 mode 2 on the RAW output of the previous conv: out = relu?(relu(a_raw * s_a + t_a) * W^T * scale + shift
// [+ residual]); a_scale_shift [G][K][2] is the BatchNorm of the layer that produced a_raw, applied (with ReLU) to the A tiles
// in shared memory. The activated tensor is never written to / re-read from HBM. K in {64, 128, 192, 256}.
int mauv_gemm_bn_xf_f16(const void* a_raw, const float* a_scale_shift, const void* w, void* y, const float* scale_shift,
                        const void* residual, int relu, int G, long long M, int N, int K, void* stream) {
  MAUV_CHECK_ARG(a_scale_shift, "mauv_gemm_bn_xf_f16: null pointer");
  return gemm_bn_impl(a_raw, w, y, nullptr, scale_shift, residual, relu, EPI_FUSED_BN, G, M, N, K, a_scale_shift, stream);
}

// Fused tail of a bottleneck WITH a downsample branch: out = relu(bn3(a1 * W3^T) + bnd(a2 * Wd^T)) as ONE contraction over
// the K-concatenated operands [a1 | a2] * [s3*W3 | sd*Wd]^T + (t3 + td): the BatchNorm scales are folded into the sampled
// weights (mauv_sample_weights_scaled_f16), the shifts into the epilogue (scale_shift = (1, t3 + td)). Neither raw conv
// output ever reaches HBM. a1 [G][M][K1], a2 [G][M][K2], w_cat [G][N][K1+K2], K1 % 64 == 0.
static int gemm_bn_cat_impl(const void* a1, int K1, const void* a2, int K2, const void* w_cat, void* y, const float* scale_shift,
                            int relu, int G, long long M, int N, const float* a1_scale_shift, void* stream) {
  MAUV_CHECK_ARG(a1 && a2 && w_cat && y && scale_shift, "mauv_gemm_bn_cat_f16: null pointer");
  MAUV_CHECK_ARG(G >= 1 && M >= 1 && N >= 64 && N % 64 == 0 && N % pick_bn(N) == 0, "mauv_gemm_bn_cat_f16: N must be 64, 128 or a multiple of 256 (got %d)", N);
  MAUV_CHECK_ARG(K1 >= 64 && K1 % 64 == 0 && K2 >= 8 && K2 % 8 == 0, "mauv_gemm_bn_cat_f16: K1 must be a multiple of 64, K2 of 8 (K1=%d K2=%d)", K1, K2);
  if (int rc = load_driver_entry_points()) return rc;
  CUtensorMap tmA, tmA2, tmB;
  if (int rc = make_tiled_map(&tmA, a1, K1, M, G, M * K1, BM)) return rc;
  if (int rc = make_tiled_map(&tmA2, a2, K2, M, G, M * K2, BM)) return rc;
  const int K = K1 + K2;
  if (int rc = make_tiled_map(&tmB, w_cat, K, N, G, static_cast<int64_t>(N) * K, pick_bn(N))) return rc;
  GemmParams p{};
  p.stack = 1;
  p.M = static_cast<int>(M);
  p.N = N;
  p.a2_kb = K1 / BK;
  p.k_blocks = p.a2_kb + static_cast<int>(ceil_div_i64(K2, BK));
  p.G = G;
  p.a_mode = 0;
  p.a_batch_mul = 1;
  p.y = static_cast<__half*>(y);
  p.ss = reinterpret_cast<const float2*>(scale_shift);
  p.has_res = 0;
  p.relu = relu;
  if (a1_scale_shift) {
    MAUV_CHECK_ARG(K1 <= 256 && N >= 128, "mauv_gemm_bn_cat_xf_f16: needs K1 <= 256 and N >= 128 (K1=%d N=%d)", K1, N);
    p.xf_ss = reinterpret_cast<const float2*>(a1_scale_shift);
    p.xf_K = K1;
    p.xf_kb = p.a2_kb;
  }
  return dispatch(tmA, tmB, p, static_cast<cudaStream_t>(stream), EPI_FUSED_BN, nullptr, &tmA2);
}

int mauv_gemm_bn_cat_f16(const void* a1, int K1, const void* a2, int K2, const void* w_cat, void* y, const float* scale_shift,
                         int relu, int G, long long M, int N, void* stream) {
  return gemm_bn_cat_impl(a1, K1, a2, K2, w_cat, y, scale_shift, relu, G, M, N, nullptr, stream);
}

// mauv_gemm_bn_cat_f16 with a1 RAW: its BatchNorm + ReLU (a1_scale_shift [G][K1][2]) is applied to the a1 k-blocks in shared memory.
int mauv_gemm_bn_cat_xf_f16(const void* a1_raw, const float* a1_scale_shift, int K1, const void* a2, int K2, const void* w_cat,
                            void* y, const float* scale_shift, int relu, int G, long long M, int N, void* stream) {
  MAUV_CHECK_ARG(a1_scale_shift, "mauv_gemm_bn_cat_xf_f16: null pointer");
  return gemm_bn_cat_impl(a1_raw, K1, a2, K2, w_cat, y, scale_shift, relu, G, M, N, a1_scale_shift, stream);
}

int mauv_conv2d_im2col_f16(const void* x, const void* w, void* y, float* stats_partial, int G,
                           int imgs_per_sample, int H, int W, int Cin, int Cout, int kh, int kw,
                           int stride, int pad, void* stream) {
  MAUV_CHECK_ARG(x && w && y, "mauv_conv2d_im2col_f16: null pointer");
  MAUV_CHECK_ARG(Cin % 64 == 0, "mauv_conv2d_im2col_f16: Cin must be a multiple of 64 (got %d)", Cin);
  MAUV_CHECK_ARG(Cout % 8 == 0, "mauv_conv2d_im2col_f16: Cout must be a multiple of 8 (got %d)", Cout);
  MAUV_CHECK_ARG(stride >= 1 && stride <= 8 && pad >= 0 && kh >= 1 && kw >= 1, "mauv_conv2d_im2col_f16: bad geometry");
  const int Ho = (H + 2 * pad - kh) / stride + 1;
  const int Wo = (W + 2 * pad - kw) / stride + 1;
  MAUV_CHECK_ARG(Ho >= 1 && Wo >= 1, "mauv_conv2d_im2col_f16: empty output");
  if (int rc = load_driver_entry_points()) return rc;
  CUtensorMap tmA, tmB;
  if (int rc = make_im2col_map(&tmA, x, Cin, W, H, static_cast<int64_t>(G) * imgs_per_sample, kh, kw, stride, pad)) return rc;
  const int K = kh * kw * Cin;
  if (int rc = make_tiled_map(&tmB, w, K, Cout, G, static_cast<int64_t>(Cout) * K, pick_bn(Cout))) return rc;
  GemmParams p{};
  p.stack = 1;
  p.M = imgs_per_sample * Ho * Wo;
  p.N = Cout;
  p.k_blocks = K / BK;
  p.G = G;
  p.a_mode = 1;
  p.a_batch_mul = 1;
  p.Wo = Wo; p.Ho = Ho; p.imgs_per_sample = imgs_per_sample;
  p.stride = stride; p.pad = pad; p.kw = kw; p.c_blocks = Cin / BK;
  p.y = static_cast<__half*>(y);
  p.stats = stats_partial;
  p.bias = nullptr;
  return dispatch(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
