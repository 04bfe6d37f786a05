// Common device/host helpers for the mauv_b200 kernels (sm_100a only).
//
// Everything here is hand-written inline PTX for Blackwell: mbarrier, TMA
// (cp.async.bulk.tensor, tiled and im2col), tcgen05 (alloc / mma / commit / ld)
// plus the thread-local error string used by the C-ABI (include/mauv_b200.h).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// ----------------------------------------------------------------------------
// Error plumbing for the C-ABI: no exceptions cross the boundary, every entry
// point returns a status and leaves a message in a thread-local buffer.
// ----------------------------------------------------------------------------
enum {
  MAUV_OK = 0,
  MAUV_ERR_BAD_ARG = 1,
  MAUV_ERR_CUDA = 2,
  MAUV_ERR_UNSUPPORTED_ARCH = 3,
  MAUV_ERR_DRIVER = 4,
};

char* mauv_err_buf();  // 512 bytes, thread local
int mauv_set_error(int code, const char* fmt, ...);

#define MAUV_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) return mauv_set_error(MAUV_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define MAUV_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t _e = (call);                                                   \
    if (_e != cudaSuccess)                                                     \
      return mauv_set_error(MAUV_ERR_CUDA, "%s failed: %s (%s:%d)", #call,     \
                            cudaGetErrorString(_e), __FILE__, __LINE__);       \
  } while (0)

#define MAUV_LAUNCH_CHECK(name)                                                \
  do {                                                                         \
    cudaError_t _e = cudaGetLastError();                                       \
    if (_e != cudaSuccess)                                                     \
      return mauv_set_error(MAUV_ERR_CUDA, "launch of %s failed: %s", name,    \
                            cudaGetErrorString(_e));                           \
  } while (0)

// Device word added to the Philox sample ids by the sampling kernels (thread-local, see mauv_set_sample_base); may be null.
const unsigned int* mauv_sample_base();

// Number of SMs of the current device (cached).
int mauv_num_sms();

static inline int64_t ceil_div_i64(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__

// ----------------------------------------------------------------------------
// Small device utilities
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a pipeline bug becomes a trap (reported as a CUDA error by the
// next API call) rather than a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("mauv_b200: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------
// TMA loads (global -> shared, completion on an mbarrier)
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// im2col mode, NHWC tensor seen as (C, W, H, N); offsets are the filter tap (s, r).
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const void* tmap, uint32_t bar,
                                                   int c, int w, int h, int n, uint16_t off_w,
                                                   uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
        "h"(off_w), "h"(off_h)
      : "memory");
}

// ----------------------------------------------------------------------------
// tcgen05: tensor memory + 5th generation tensor core MMA
// ----------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_result_addr),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread retire.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor for a K-major tile stored with the 128-byte
// swizzle (rows at a 128 B pitch, 8-row groups 1024 B apart). Field layout as in
// the PTX ISA "tcgen05 shared memory descriptor": start address [0,14) (>>4),
// LBO [16,30) (unused for swizzled K-major, set to 1), SBO [32,46) (>>4),
// version [46,48) = 1, layout type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// MN-major operand tile: the M (or N) dimension is contiguous in memory and the reduction index walks the rows - what a
// TMA box [64 reduction rows][64 elements = 128 B] of a row-major [pixels][channels] tensor looks like in smem. Canonical
// layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: 64 contiguous MN elements per 128 B row, 8 K rows per 1024 B
// swizzle atom, SBO = 1024 B to the next 8 K rows, LBO = bytes to the next block of 64 MN elements (the next TMA box).
__device__ __forceinline__ uint64_t umma_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16: fp16 A/B (format 0), fp32 D (c_format 1),
// both operands K-major, dense, no negate. n_dim = N>>3 at bit 17, m_dim = M>>4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t m, uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
// Same with both operands MN-major (a_major bit 15, b_major bit 16).
__host__ __device__ constexpr uint32_t umma_idesc_f16_mn(uint32_t m, uint32_t n) {
  return umma_idesc_f16(m, n) | (1u << 15) | (1u << 16);
}

// ----------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller: the production eps ~ N(0,1) stream.
// counter = (elem_quad_lo, elem_quad_hi, sample_id, layer_id), key = seed.
// Element index is the index in the PyTorch parameter layout so that the stream
// is independent of our packed weight layout; oracle/philox.py restates it.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Box-Muller on the SFU: lg2.approx / sqrt.approx / sin.approx / cos.approx (relative error ~2^-22, absolute error of a
// normal ~1e-6) instead of libm's logf / sqrtf / sincospif: the sampler is ALU/SFU-bound (Philox + Box-Muller + softplus per
// weight per sample) and the exact routines were half of its instructions. The stream differs from oracle/philox.py (numpy,
// exact libm) by ~1e-6 absolute - three orders below the fp16 rounding of the sampled weight it feeds. EVERY kernel that
// draws or replays eps goes through these two functions, so forward and backward see bit-identical eps.
__device__ __forceinline__ void box_muller_pair(uint32_t ra_bits, uint32_t rb_bits, float& z0, float& z1) {
  const float k = 1.0f / 16777216.0f;
  const float u1 = (static_cast<float>(ra_bits >> 8) + 0.5f) * k;   // (0, 1)
  const float u2 = (static_cast<float>(rb_bits >> 8) + 0.5f) * k;
  float rad;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(rad) : "f"(-2.0f * __logf(u1)));
  float sn, cs;
  __sincosf(6.283185307179586f * (u2 - 0.5f), &sn, &cs);     // argument in (-pi, pi): the SFU's accurate range
  z0 = -rad * cs;                                             // cos(t + pi) = -cos t, sin(t + pi) = -sin t
  z1 = -rad * sn;
}

// Four N(0,1) values of one Philox counter block: elements 4*quad .. 4*quad+3 of (seed, layer, sample).
__device__ __forceinline__ void philox_normals4(uint64_t seed, uint32_t layer_id, uint32_t sample_id, uint64_t quad,
                                                float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(quad), static_cast<uint32_t>(quad >> 32), sample_id, layer_id,
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  box_muller_pair(r[0], r[1], z[0], z[1]);
  box_muller_pair(r[2], r[3], z[2], z[3]);
}

// One N(0,1) value for (seed, layer, sample, element index): pair (r0, r1) -> z0, z1 ; pair (r2, r3) -> z2, z3
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t layer_id, uint32_t sample_id,
                                               uint64_t elem) {
  uint32_t r[4];
  const uint64_t quad = elem >> 2;
  philox4x32_10(static_cast<uint32_t>(quad), static_cast<uint32_t>(quad >> 32), sample_id, layer_id,
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  const uint32_t which = static_cast<uint32_t>(elem & 3u);
  float z0, z1;
  box_muller_pair((which & 2u) ? r[2] : r[0], (which & 2u) ? r[3] : r[1], z0, z1);
  return (which & 1u) ? z1 : z0;
}

// sigma = log1p(exp(rho)) exactly as the reference computes it (no threshold).
__device__ __forceinline__ float softplus_ref(float rho) { return log1pf(expf(rho)); }

#endif  // __CUDACC__
