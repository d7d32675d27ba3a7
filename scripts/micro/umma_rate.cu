// tcgen05.mma issue rate from shared-memory operands (SS mode) by tile shape: cycles per M x N x 16 bf16 instruction with
// everything resident (no loads, one accumulator, MMAs back to back) - is a narrow N bound by the MMA floor
// max(M,128) * N / 256 or by the operand reads?   Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I3d-latent-diffusion-model_b200/csrc \
//        -o /tmp/umma_rate scripts/micro/umma_rate.cu -lcuda && /tmp/umma_rate
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace b200ldm {
void set_error(const char*, ...) {}
int fail(int c, const char*, ...) { return c; }
std::atomic<uint64_t> g_launches{0};
}  // namespace b200ldm
using namespace b200ldm;

// A: 128 rows x 64 bf16 (16 KB, SWIZZLE_128B K-major), B: N rows x 64 bf16.  Contents are irrelevant for timing.
// ASHARE: all four k-steps of a group read the SAME A slice (k = 0) - no difference expected if A is re-read per MMA.
// AOFF / SBO: the A view starts AOFF bytes into the tile and steps SBO bytes between 8-row groups (conv_halo.cu reads its
// W-halo boxes through views with AOFF = kw * 128 and SBO = 1280)
// BMN: B is MN-major (the attention kernels read V as landed by TMA: rows = keys, 64 head-dim columns per 128-byte row,
// LBO = one 64-column block of 64 keys = 8 KB, 16 keys per k-step = 2048 bytes) - is that operand form slower?
template <int M, int N, int AOFF = 0, int SBO = 1024, bool BMN = false>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* cyc, int groups) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + 24576, bar = sb + 256 * 128, slot = bar + 8;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - ptx::smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (24576 + 256 * 128) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - ptx::smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (warp == 0 && lane == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (warp == 1) ptx::tmem_alloc<512>(slot);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = BMN ? ptx::umma_idesc_bf16_bmn(M, N) : ptx::umma_idesc_bf16(M, N);
    const uint32_t bmn_lo = ((sb & 0x3FFFFu) >> 4) | ((uint32_t)((8192 >> 4) & 0x3FFF) << 16);
    constexpr uint32_t khi = ptx::umma_desc_hi_sw128();
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (BMN)
          ptx::umma_bf16_lohi(tmem, ptx::umma_desc_lo_sw128(sa + k * 32), bmn_lo + (uint32_t)((k * 2048) >> 4), khi, khi, idesc,
                              (g | k) ? 1u : 0u);
        else
          ptx::umma_bf16(tmem, ptx::umma_desc_k_sw128(sa + AOFF + k * 32, SBO), ptx::umma_desc_k_sw128(sb + k * 32), idesc,
                         (g | k) ? 1u : 0u);
      }
    }
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, 0, 99);
    const long long t1 = clock64();
    if (blockIdx.x == 0) cyc[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem); }
}

template <int M, int N, int AOFF = 0, int SBO = 1024, bool BMN = false>
static void run(long long* d_cyc, int ctas) {
  const int groups = 2000, smem = 24576 + 256 * 128 + 2048;
  cudaFuncSetAttribute(rate_kernel<M, N, AOFF, SBO, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) rate_kernel<M, N, AOFF, SBO, BMN><<<ctas, 128, smem>>>(d_cyc, groups);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / (groups * 4.0), floor_c = (M > 128 ? M : 128) * (double)N / 256.0;
  printf("M=%3d N=%3d A+%3d SBO %4d %s %d CTAs: %7.1f cycles per MMA (floor %5.1f, %5.1f %% of the floor rate)%s\n", M, N, AOFF, SBO,
         BMN ? "B MN-major" : "", ctas, per, floor_c, 100.0 * floor_c / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_cyc;
  cudaMalloc(&d_cyc, 8);
  for (int ctas : {1, 148}) {
    run<128, 256>(d_cyc, ctas);
    run<128, 128>(d_cyc, ctas);
    run<128, 64>(d_cyc, ctas);
    run<128, 32>(d_cyc, ctas);
    run<64, 256>(d_cyc, ctas);
    run<64, 128>(d_cyc, ctas);
    run<128, 64, 128, 1280>(d_cyc, ctas);    // conv_halo.cu views
    run<128, 64, 256, 1280>(d_cyc, ctas);
    run<128, 64, 0, 1280>(d_cyc, ctas);
    run<128, 128, 128, 1280>(d_cyc, ctas);
    run<128, 64, 128, 1024>(d_cyc, ctas);
    run<128, 256, 0, 1024, true>(d_cyc, ctas);   // PV of the head-dim-256 attention: V MN-major
    run<128, 128, 0, 1024, true>(d_cyc, ctas);
    run<128, 64, 0, 1024, true>(d_cyc, ctas);    // PV of the head-dim-64 attention
  }
  return 0;
}
