// MUFU.EX2 throughput on this GPU: cycles per warp-instruction per SM sub-partition, alone and mixed with FFMA / F2FP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  float acc = 0.f;
  unsigned pk = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float x = a[i];
      if (MODE >= 1) x = fmaf(x, 0.999f, -0.001f);               // FFMA in front of every MUFU
      float y;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
      if (MODE >= 2) acc += y;                                    // FADD behind it
      a[i] = y * 1e-3f;
    }
    if (MODE >= 3) {                                              // one F2FP per two elements
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a[i], a[i + 1]);
        pk ^= *reinterpret_cast<unsigned*>(&h);
      }
    }
  }
  const long long t1 = clock64();
  float s = acc + __uint_as_float(pk);
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int mode = 0; mode < 4; ++mode)
    for (int warps_per_smsp : {1, 2, 4, 8}) {
      const int threads = warps_per_smsp * 4 * 32;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
        if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
        if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
        if (mode == 3) k<3><<<148, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      const double mufu_per_smsp = (double)iters * 8 * warps_per_smsp;
      printf("mode %d (%s) warps/SMSP %d: %.2f cycles per MUFU warp-instr per SMSP\n", mode,
             mode == 0 ? "MUFU only" : mode == 1 ? "FFMA+MUFU" : mode == 2 ? "FFMA+MUFU+FADD" : "FFMA+MUFU+FADD+0.5 F2FP",
             warps_per_smsp, avg / mufu_per_smsp);
    }
  return 0;
}
