// Measurement aid (not on the product path): HBM write-bandwidth probes used to settle which resource bounds the
// write-heavy 1x1 convolutions ("What's weak" 6: is 3.9 TB/s really the write-only ceiling of this GPU, or is it
// the TMA-store epilogue?). Three ways to write `bytes` bytes of a constant:
//   mode 0  st.global.v4 (default cache policy), grid-stride, one CTA per SM x 8
//   mode 1  st.global.cs.v4 (streaming / evict-first hint)
//   mode 2  cp.async.bulk.global.shared::cta (the bulk-copy engine that TMA stores use): every CTA fills one 32 KB
//           shared-memory buffer once and streams it out in 32 KB pieces, 4 copies in flight
// tests/gpu_microbench.py hbmwrite times them next to cudaMemsetAsync and writes profiles/r2_hbm_write_probe.md.
#include "common.cuh"

namespace {

template <int MODE>
__global__ void __launch_bounds__(256)
fill_v4_kernel(uint4* __restrict__ dst, long long n16, uint32_t value) {
  const uint4 v = make_uint4(value, value, value, value);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
    if (MODE == 1) asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else dst[i] = v;
  }
}

constexpr int kBulkBytes = 32768;

__global__ void __launch_bounds__(128)
fill_bulk_kernel(uint8_t* __restrict__ dst, long long pieces, uint32_t value) {
  extern __shared__ __align__(128) uint8_t buf[];
  for (int i = threadIdx.x; i < kBulkBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(buf)[i] = make_uint4(value, value, value, value);
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t src = smem_u32(buf);
    for (long long pc = blockIdx.x; pc < pieces; pc += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + pc * kBulkBytes), "r"(src), "n"(kBulkBytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

}  // namespace

extern "C" {

int mauv_membench_fill(void* dst, long long bytes, unsigned int value, int mode, void* stream) {
  MAUV_CHECK_ARG(dst && bytes >= kBulkBytes && bytes % kBulkBytes == 0 && mode >= 0 && mode <= 2 &&
                 (reinterpret_cast<uintptr_t>(dst) & 127) == 0, "mauv_membench_fill: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == 2) {
    static bool attr = false;
    if (!attr) {
      MAUV_CUDA(cudaFuncSetAttribute(fill_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBulkBytes));
      attr = true;
    }
    fill_bulk_kernel<<<mauv_num_sms() * 4, 128, kBulkBytes, st>>>(static_cast<uint8_t*>(dst), bytes / kBulkBytes, value);
  } else if (mode == 1) {
    fill_v4_kernel<1><<<mauv_num_sms() * 8, 256, 0, st>>>(static_cast<uint4*>(dst), bytes / 16, value);
  } else {
    fill_v4_kernel<0><<<mauv_num_sms() * 8, 256, 0, st>>>(static_cast<uint4*>(dst), bytes / 16, value);
  }
  MAUV_LAUNCH_CHECK("mauv_membench_fill");
  return MAUV_OK;
}

}  // extern "C"
