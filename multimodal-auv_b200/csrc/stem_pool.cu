// Stem of the three ResNet-50 trunks for inference: 7x7/2 Bayesian conv (explicit im2col matrix shared by every MC
// sample of the batch) + BatchNorm batch statistics + the 3x3/2 max-pool, with the max-pool taken on the RAW conv
// output inside the epilogue. The reference runs conv1 -> bn1 (train mode) -> relu -> maxpool once per MC pass
// (models/base_models.py:74-76 through torchvision resnet.py `_forward_impl`); its conv1 output [B,64,128,128] is the
// largest tensor of the network (5.4 GB per 10 samples at B = 256) and was written and re-read once per sample.
//
// Why the pool can run before the BatchNorm: y -> relu(s*y + t) is monotone per channel (non-decreasing for s >= 0,
// non-increasing otherwise, fp16 rounding included), so maxpool(relu(bn(y))) == relu(bn(pool(y))) with pool = window max
// (s >= 0) or window min (s < 0), and sign(s) = sign(gamma) is known before the statistics are (gamma is a parameter).
// The kernel therefore emits the statistics of the full-resolution output (for bn1) and only the POOLED raw tensor
// (1/4 of the bytes); a bn_act pass over the pooled tensor finishes the stem. Results are bit-identical to
// gemm_f16_tc_kernel (stacked mode) + bn_relu_maxpool_kernel: same fp16 rounding of the accumulator, same partial-sum order.
//
// Structure (one CTA per SM, 12 warps): tile = one conv output row (Wo = 128 pixels) x 4 samples x 64 channels
// (128 x 256 accumulator, 2 TMEM stages). A CTA owns ONE block of 4 samples for the whole launch - its sampled weights
// ([256][Kp] fp16, <= 96 KB) are loaded once and stay in shared memory - and walks units of (image, chunk of P pooled
// rows) row by row, so only the im2col rows stream (16 KB per k-block, shared through L2 by the CTAs that work on the
// same image for other sample blocks). Epilogue warps (2 column sets x 4 TMEM lane quarters): tcgen05.ld -> fp16 ->
// swizzled smem staging -> statistics -> horizontal 3-max from the staged rows -> vertical combine with the running
// row state kept in registers -> 16-byte global stores of the pooled rows.
#include "common.cuh"
#include "tmap.cuh"

namespace {

struct StemPoolParams {
  int G;               // MC samples
  int g_blocks;        // ceil(G / 4)
  int lanes;           // gridDim.x / g_blocks: CTAs per sample block
  int imgs;            // images per sample (B)
  int Ho;              // conv output rows per image (even); Wo == 128
  int k_blocks;        // ceil(Kp / 64), 1..3
  int P;               // pooled rows per unit
  int chunks;          // (Ho / 2) / P
  int rest;            // imgs * chunks: units per sample block
  __half* out;         // [G][imgs][Ho/2][64][64] pooled raw conv output
  float* stats;        // [G][imgs * Ho][64][2] partial (sum, sum of squares) per conv row
  const float* gamma;  // [64] BatchNorm weight (sign selects max / min) or nullptr (all non-negative)
};

constexpr int kSPBN = 256;                       // accumulator columns: 4 samples x 64 channels
constexpr int kSPAStages = 3;
constexpr int kSPABytes = BM * BK * 2;           // 16 KB per k-block of im2col rows
constexpr int kSPBBlockBytes = kSPBN * BK * 2;   // 32 KB per k-block of weights
constexpr int kSPBBytes = 3 * kSPBBlockBytes;    // resident weights of the CTA's sample block
constexpr int kSPOutBufBytes = 32 * 64 * 2;      // one warp's 32 rows x 64 channels, 128B-swizzled
constexpr int kSPOutBytes = 8 * 2 * kSPOutBufBytes;
constexpr int kSPStatBytes = 2 * 4 * kSPBN * 2 * 4;
constexpr int kSPSmem = 1024 + kSPBBytes + kSPAStages * kSPABytes + kSPOutBytes + kSPStatBytes + 256;

__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint4 hmax2_u4(const uint4& a, const uint4& b) {
  return make_uint4(hmax2_u32(a.x, b.x), hmax2_u32(a.y, b.y), hmax2_u32(a.z, b.z), hmax2_u32(a.w, b.w));
}
__device__ __forceinline__ uint4 xor_u4(const uint4& a, const uint4& m) {
  return make_uint4(a.x ^ m.x, a.y ^ m.y, a.z ^ m.z, a.w ^ m.w);
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(384, 1)
stem_conv_pool_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const StemPoolParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_base = smem_base;
  const uint32_t a_base = b_base + kSPBBytes;
  const uint32_t out_base = a_base + kSPAStages * kSPABytes;
  float* stat_smem = reinterpret_cast<float*>(smem_gen + kSPBBytes + kSPAStages * kSPABytes + kSPOutBytes);
  const uint32_t bar_base = out_base + kSPOutBytes + kSPStatBytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kSPAStages + s); };
  const uint32_t b_full = bar_base + 8u * (2 * kSPAStages);
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kSPAStages + 1 + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kSPAStages + 3 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * kSPAStages + 5);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + kSPBBytes + kSPAStages * kSPABytes + kSPOutBytes + kSPStatBytes + 8 * (2 * kSPAStages + 5));

  const int warp = threadIdx.x >> 5;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kSPAStages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    mbar_init(b_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 8);       // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr_addr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // this CTA's sample block and its share of the (image, row chunk) units; CTAs with the same `lane` work on the same
  // unit for the other sample blocks at the same time, so the im2col rows come out of L2 for all but one of them
  const int gb = static_cast<int>(blockIdx.x) % p.g_blocks;
  const int lane_cta = static_cast<int>(blockIdx.x) / p.g_blocks;
  // unit u -> image b, pooled rows [P*c, P*(c+1)): conv rows 2*P*c - 1 (clipped at 0) .. 2*P*(c+1) - 1
  auto unit_rows = [&](int u, int& b, int& r_begin, int& r_end) {
    b = u / p.chunks;
    const int c = u - b * p.chunks;
    r_begin = 2 * p.P * c - 1;
    if (r_begin < 0) r_begin = 0;
    r_end = 2 * p.P * (c + 1) - 1;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(b_full, static_cast<uint32_t>(p.k_blocks) * kSPBBlockBytes);
      for (int kb = 0; kb < p.k_blocks; ++kb)      // rows past G*64 (partial last sample block) are zero-filled
        tma_load_3d(b_base + kb * kSPBBlockBytes, &tmB, b_full, kb * BK, gb * kSPBN, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int u = lane_cta; u < p.rest; u += p.lanes) {
        int b, r_begin, r_end;
        unit_rows(u, b, r_begin, r_end);
        for (int r = r_begin; r <= r_end; ++r) {
          const int m0 = (b * p.Ho + r) * BM;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(a_empty(stage), phase ^ 1u);
            mbar_expect_tx(a_full(stage), kSPABytes);
            tma_load_3d(a_base + stage * kSPABytes, &tmA, a_full(stage), kb * BK, m0, 0);
            if (++stage == kSPAStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, kSPBN);
      int stage = 0;
      uint32_t phase = 0, it = 0;
      mbar_wait(b_full, 0);
      for (int u = lane_cta; u < p.rest; u += p.lanes) {
        int b, r_begin, r_end;
        unit_rows(u, b, r_begin, r_end);
        for (int r = r_begin; r <= r_end; ++r, ++it) {
          const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
          mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kSPBN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(a_full(stage), phase);
            tcgen05_fence_after();
            const uint64_t a_desc = umma_smem_desc_sw128(a_base + stage * kSPABytes);
            const uint64_t b_desc = umma_smem_desc_sw128(b_base + kb * kSPBBlockBytes);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_f16_ss(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(a_empty(stage));
            if (++stage == kSPAStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(tmem_full_bar(acc));
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp & 3;                  // TMEM lane quarter = pixels 32*ew .. 32*ew+31 of the conv row
    const int cset = (warp - 4) >> 2;         // column set: sample slots cset and cset + 2 of the block
    const uint32_t lane = lane_id();
    const int es = threadIdx.x - 128 - cset * 128;                      // 0..127 inside the column set
    const uint32_t my_out = out_base + (warp - 4) * (2 * kSPOutBufBytes);
    const uint32_t left_out = my_out - 2 * kSPOutBufBytes;              // lane quarter ew - 1 of the same set (ew > 0)
    // staged-row offsets of this lane's channel pair (statistics) for rows r mod 8
    uint32_t sw_off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sw_off[j] = (((lane >> 2) ^ static_cast<uint32_t>(j)) << 4) + ((lane & 3u) << 2);
    // pooling: lane owns the 16-byte channel chunk ch of the pooled pixels 16*ew + (lane >> 3) + 4*j, j = 0..3
    const uint32_t ch = lane & 7u;
    const uint32_t ql = lane >> 3;
    uint4 sgn = make_uint4(0u, 0u, 0u, 0u);   // 0x8000 per fp16 channel with gamma < 0: max(-y) == -min(y)
    if (p.gamma) {
      uint32_t m[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g0 = __ldg(p.gamma + ch * 8 + 2 * j), g1 = __ldg(p.gamma + ch * 8 + 2 * j + 1);
        m[j] = (g0 < 0.f ? 0x8000u : 0u) | (g1 < 0.f ? 0x80000000u : 0u);
      }
      sgn = make_uint4(m[0], m[1], m[2], m[3]);
    }
    const int n_valid = (p.G - gb * 4) < 4 ? (p.G - gb * 4) : 4;      // sample slots of this block that exist (1..4)
    const int Hp = p.Ho >> 1;
    uint4 state[2][4];                         // running vertical max of the open pooled row, per (slot, item)
    uint32_t it = 0, blk = 0;
    for (int u = lane_cta; u < p.rest; u += p.lanes) {
      int b, r_begin, r_end;
      unit_rows(u, b, r_begin, r_end);
      for (int r = r_begin; r <= r_end; ++r, ++it) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        // the first row of a unit below the image top belongs to the previous chunk (it only seeds the pool window)
        const bool own_row = !(r == r_begin && r > 0);
        const int m_tile = b * p.Ho + r;
        float* stat_buf = stat_smem + (it & 1u) * (4 * kSPBN * 2);
        mbar_wait(tmem_full_bar(acc), acc_phase);
        tcgen05_fence_after();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int slot = cset + 2 * i;
          const bool valid = slot < n_valid;               // uniform over the 128 threads of the set
          const bool last = (i == 1) || (slot + 2 >= n_valid);
          if (valid) {
            const int col0 = slot * 64;
            const uint32_t taddr = tmem_base + acc * kSPBN + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(col0);
            uint32_t ra[32], rb[32];
            tmem_ld_32x32b_x32(taddr, ra);
            tmem_ld_32x32b_x32(taddr + 32, rb);
            tmem_ld_wait();
            if (last) {      // this warp's last TMEM read of the tile: hand the accumulator stage back to the MMA warp
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
            }
            const uint32_t bsel = (blk & 1u) * kSPOutBufBytes;
            const uint32_t buf = my_out + bsel;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t* src = (q < 4) ? &ra[q * 8] : &rb[(q - 4) * 8];
              __half2 h0 = __floats2half2_rn(__uint_as_float(src[0]), __uint_as_float(src[1]));
              __half2 h1 = __floats2half2_rn(__uint_as_float(src[2]), __uint_as_float(src[3]));
              __half2 h2 = __floats2half2_rn(__uint_as_float(src[4]), __uint_as_float(src[5]));
              __half2 h3 = __floats2half2_rn(__uint_as_float(src[6]), __uint_as_float(src[7]));
              const uint32_t off = lane * 128u + ((static_cast<uint32_t>(q) ^ (lane & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + off), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                           "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                           "r"(*reinterpret_cast<uint32_t*>(&h3))
                           : "memory");
            }
            __syncwarp();
            if (own_row) {
              // lane j owns channels 2j, 2j+1: column sums over the warp's 32 rows of the staged fp16 values, in the same
              // order as gemm_f16_tc_kernel's store + statistics epilogue (two chains over even / odd rows)
              unsigned long long sa = 0ull, sb = 0ull, qa = 0ull, qb = 0ull;
#pragma unroll
              for (int r8 = 0; r8 < 32; r8 += 8) {          // 8 reads in flight, then their 4 x 2 accumulation steps (same order)
                uint32_t wv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv[u]) : "r"(buf + sw_off[(r8 + u) & 7] + (r8 + u) * 128u));
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                  const float2 f0 = __half22float2(*reinterpret_cast<__half2*>(&wv[u]));
                  const float2 f1 = __half22float2(*reinterpret_cast<__half2*>(&wv[u + 1]));
                  const unsigned long long p0 = *reinterpret_cast<const unsigned long long*>(&f0);
                  const unsigned long long p1 = *reinterpret_cast<const unsigned long long*>(&f1);
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sa) : "l"(p0));
                  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(qa) : "l"(p0));
                  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sb) : "l"(p1));
                  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(qb) : "l"(p1));
                }
              }
              asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sa) : "l"(sb));
              asm("add.rn.f32x2 %0, %0, %1;" : "+l"(qa) : "l"(qb));
              const float2 sf = *reinterpret_cast<float2*>(&sa), qf = *reinterpret_cast<float2*>(&qa);
              *reinterpret_cast<float4*>(stat_buf + (ew * kSPBN + col0 + 2 * lane) * 2) = make_float4(sf.x, qf.x, sf.y, qf.y);
            }
            // every lane quarter of this (tile, slot) is staged (and its statistics are in stat_buf)
            asm volatile("bar.sync %0, 128;" ::"r"(1 + cset) : "memory");
            if (own_row && last) {
              // one deterministic partial per (conv row, channel): the 4 lane quarters in a fixed order
              const int n_slots = (n_valid - cset + 1) >> 1;            // valid slots of this set: cset, cset + 2
              for (int jj = es; jj < 64 * n_slots; jj += 128) {
                const int sl = cset + 2 * (jj >> 6), c = jj & 63;
                const int j = sl * 64 + c;
                float a = 0.f, q2 = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  a += stat_buf[(w * kSPBN + j) * 2 + 0];
                  q2 += stat_buf[(w * kSPBN + j) * 2 + 1];
                }
                reinterpret_cast<float2*>(p.stats)[(static_cast<long long>(gb * 4 + sl) * (p.imgs * p.Ho) + m_tile) * 64 + c] =
                    make_float2(a, q2);
              }
            }
            // horizontal 3-max (stride 2, pad 1) from the staged rows, then the vertical combine
            const long long out_row = ((static_cast<long long>(gb * 4 + slot) * p.imgs + b) * Hp + (r >> 1)) * 64;
            // all 12 window reads first (one exposed shared-memory latency instead of twelve: the ncu source view had these
            // warps 20 % in short_scoreboard stalls on the LDS -> XOR chains)
            uint4 vc[4], vr[4], vl[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t q_local = ql + 4u * j;                    // pooled pixel 16*ew + q_local
              const uint32_t rc = 2u * q_local;                        // staged row of the window centre (0..30)
              const uint32_t a_c = buf + rc * 128u + ((ch ^ (rc & 7u)) << 4);
              // left neighbour: row rc - 1, or (rc == 0) pixel 32*ew - 1 = row 31 of the left lane quarter's slab, or (image
              // border) the centre again - max(x, x) = x
              const uint32_t a_l = rc > 0u ? buf + (rc - 1u) * 128u + ((ch ^ ((rc - 1u) & 7u)) << 4)
                                           : (ew > 0 ? left_out + bsel + 31u * 128u + ((ch ^ 7u) << 4) : a_c);
              vc[j] = lds_u4(a_c);
              vr[j] = lds_u4(buf + (rc + 1u) * 128u + ((ch ^ ((rc + 1u) & 7u)) << 4));
              vl[j] = lds_u4(a_l);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t q_local = ql + 4u * j;
              const uint4 h = hmax2_u4(hmax2_u4(xor_u4(vc[j], sgn), xor_u4(vr[j], sgn)), xor_u4(vl[j], sgn));
              if (r == r_begin) {
                state[i][j] = h;
              } else if ((r & 1) == 0) {
                state[i][j] = hmax2_u4(state[i][j], h);
              } else {
                const uint4 o = xor_u4(hmax2_u4(state[i][j], h), sgn);
                *reinterpret_cast<uint4*>(p.out + (out_row + 16 * ew + q_local) * 64 + ch * 8) = o;
                state[i][j] = h;
              }
            }
            ++blk;
          }
        }
        if (n_valid <= cset) {     // this column set has no sample in the block: still take part in the TMEM hand-shake
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

extern "C" {

// Conv rows (128-pixel tiles) of the statistics buffer written by mauv_stem_conv_pool_f16: imgs * Ho.
// a0: [imgs*Ho*128][Kp] fp16 im2col matrix of the batch (mauv_stem_im2col_f16), w: [G][64][Kp] sampled weights,
// pooled: [G][imgs][Ho/2][64][64] fp16 = 3x3/2 max-pool (pad 1) of the raw conv output - the window MIN for channels
// with gamma < 0 -, stats_partial: [G][imgs*Ho][64][2] fp32. Requires Wo = 128 (256 x 256 inputs), Ho even, Kp <= 192.
int mauv_stem_conv_pool_f16(const void* a0, const void* w, void* pooled, float* stats_partial, const float* gamma, int G,
                            int imgs, int Ho, int Kp, void* stream) {
  MAUV_CHECK_ARG(a0 && w && pooled && stats_partial, "mauv_stem_conv_pool_f16: null pointer");
  MAUV_CHECK_ARG(G >= 1 && imgs >= 1 && Ho >= 2 && Ho % 2 == 0, "mauv_stem_conv_pool_f16: bad shape G=%d imgs=%d Ho=%d", G, imgs, Ho);
  MAUV_CHECK_ARG(Kp >= 8 && Kp % 8 == 0 && Kp <= 3 * BK, "mauv_stem_conv_pool_f16: Kp must be a multiple of 8 in [8, 192] (got %d)", Kp);
  MAUV_CHECK_ARG((reinterpret_cast<uintptr_t>(a0) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(pooled) & 15) == 0, "mauv_stem_conv_pool_f16: pointers must be 16-byte aligned");
  const long long M = static_cast<long long>(imgs) * Ho * BM;
  MAUV_CHECK_ARG(M < (1LL << 31), "mauv_stem_conv_pool_f16: too many pixels");
  if (int rc = load_driver_entry_points()) return rc;
  CUtensorMap tmA, tmB;
  if (int rc = make_tiled_map(&tmA, a0, Kp, M, 1, M * Kp, BM)) return rc;
  if (int rc = make_tiled_map(&tmB, w, Kp, static_cast<int64_t>(G) * 64, 1, static_cast<int64_t>(G) * 64 * Kp, kSPBN)) return rc;
  StemPoolParams p{};
  p.G = G;
  p.g_blocks = (G + 3) / 4;
  const int sms = mauv_num_sms();
  MAUV_CHECK_ARG(p.g_blocks <= sms, "mauv_stem_conv_pool_f16: too many samples per launch (%d)", G);
  p.lanes = sms / p.g_blocks;
  p.imgs = imgs;
  p.Ho = Ho;
  p.k_blocks = static_cast<int>(ceil_div_i64(Kp, BK));
  // pooled rows per unit: as large as possible (one extra conv row is recomputed per unit) while every CTA still gets
  // a few units to balance on
  const int Hp = Ho / 2;
  int P = Hp;
  for (int cand = 16; cand >= 2; cand >>= 1)
    if (Hp % cand == 0) { P = cand; break; }
  while (P > 2 && P % 2 == 0 && static_cast<long long>(imgs) * (Hp / P) < 4LL * p.lanes) P >>= 1;
  p.P = P;
  p.chunks = Hp / P;
  p.rest = imgs * p.chunks;
  if (p.lanes > p.rest) p.lanes = p.rest;
  p.out = static_cast<__half*>(pooled);
  p.stats = stats_partial;
  p.gamma = gamma;
  static bool attr_set = false;
  if (!attr_set) {
    MAUV_CUDA(cudaFuncSetAttribute(stem_conv_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSPSmem));
    attr_set = true;
  }
  const unsigned grid = static_cast<unsigned>(p.lanes * p.g_blocks);
  stem_conv_pool_kernel<<<grid, 384, kSPSmem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  MAUV_LAUNCH_CHECK("stem_conv_pool_kernel");
  return MAUV_OK;
}

}  // extern "C"
