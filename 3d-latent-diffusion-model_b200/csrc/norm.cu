// norm.cu - GroupNorm (+SiLU) on NDHWC activations: HBM-bound, vectorised (8 channels =
// one 16-byte bf16 load per thread), statistics in fp32 per thread -> fp32 shared atomics per
// block -> fp64 global atomics (so E[x^2]-mean^2 survives eps = 1e-6).
//   stats pass : 1 read;  apply pass : 1 read + 1 write  (algorithmic bytes)
// Thread mapping: a thread owns one 8-channel vector position `cv` for its whole life, so its
// group ids, gamma/beta and mean/rstd are loop invariant; it strides over voxels.
#include <cstdlib>

#include "common.cuh"

namespace b200ldm {

constexpr int kNormThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kNormThreads, 4)
gn_stats_kernel(const T* __restrict__ x, const T* __restrict__ x2, double* __restrict__ stats, long long S,
                int C1, int C2, int G, int vox_per_block) {
  extern __shared__ float sh[];  // [2*G]
  // blocks walk the tensor BACKWARDS: the producing conv wrote it front to back, so its tail is what still sits in the
  // 126 MB L2 (an activation at the top level is 181-363 MB); reading that first turns part of the pass into L2 hits
  const int n = gridDim.y - 1 - blockIdx.y;
  const int C = C1 + C2;
  const int vecs = C >> 3;
  const int cpg = C / G;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int vpp = blockDim.x / vecs;  // voxels per pass (>= 1)
  const int cv = threadIdx.x % vecs;
  const int vslot = threadIdx.x / vecs;
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
  if (vslot < vpp) {
    const long long v0 = (long long)(gridDim.x - 1 - blockIdx.x) * vox_per_block;
    const long long v1 = min(v0 + vox_per_block, S);
    const bool second = cv * 8 >= C1;
    const int Cs = second ? C2 : C1;
    const T* base = (second ? x2 : x) + ((size_t)n * S) * Cs + (size_t)(second ? cv * 8 - C1 : cv * 8);
    // 4 independent 16-byte loads in flight per thread (memory-level parallelism), then the tail
    long long v = v0 + vslot;
    for (; v + 3LL * vpp < v1; v += 4LL * vpp) {
      float f[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) Vec8<T>::load(base + (size_t)(v + (long long)u * vpp) * Cs, f[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) { s1[e] += f[u][e]; s2[e] += f[u][e] * f[u][e]; }
    }
    for (; v < v1; v += vpp) {
      float f[8];
      Vec8<T>::load(base + (size_t)v * Cs, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s1[e] += f[e]; s2[e] += f[e] * f[e]; }
    }
    // fold the 8 lanes into their groups (cpg >= 8: all one group; else 8/cpg groups)
    if (cpg % 8 == 0) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { a += s1[e]; b += s2[e]; }
      const int g = (cv * 8) / cpg;
      atomicAdd(&sh[2 * g], a);
      atomicAdd(&sh[2 * g + 1], b);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int g = (cv * 8 + e) / cpg;
        atomicAdd(&sh[2 * g], s1[e]);
        atomicAdd(&sh[2 * g + 1], s2[e]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x)
    atomicAdd(stats + (size_t)n * 2 * G + i, (double)sh[i]);
}

template <typename T, int UNROLL, int MINB>
__global__ void __launch_bounds__(kNormThreads, MINB)
gn_apply_kernel(const T* __restrict__ x, const T* __restrict__ x2, T* __restrict__ y,
                const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                long long S, int C1, int C2, int G, float eps, int silu, int vox_per_block, double inv_cnt_arg,
                const double* __restrict__ cst1, const double* __restrict__ cst2, int cw) {
  const int n = gridDim.y - 1 - blockIdx.y;   // back to front, see gn_stats_kernel
  const int C = C1 + C2;
  const int vecs = C >> 3;
  const int cpg = C / G;
  const int vpp = blockDim.x / vecs;
  const int cv = threadIdx.x % vecs;
  const int vslot = threadIdx.x / vecs;
  if (vslot >= vpp) return;
  float sc[8], sf[8];
  // mean / variance from the fp64 sums with two fp64 multiplies and one fp64 FMA (no fp64 divide / sqrt: the
  // per-block prologue used to rival the streaming time); the cancellation-prone subtraction stays in fp64.
  const double inv_cnt = inv_cnt_arg;
  int g_prev = -1;
  float mean_f = 0.f, rstd = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = cv * 8 + e;
    const int g = c / cpg;
    if (g != g_prev) {
      double sum1, sum2;
      if (cst1) {
        // assemble the group from per-cw-channel chunk sums (cw = 8, or 2 / 4 = the group width of narrow tensors);
        // chunks [0, C1/cw) live in cst1, the rest in cst2
        sum1 = 0.0; sum2 = 0.0;
        const int k0 = (g * cpg) / cw, k1 = ((g + 1) * cpg) / cw, kc1 = C1 / cw;
        for (int k = k0; k < k1; ++k) {
          const double* e2 = k < kc1 ? cst1 + ((size_t)n * kc1 + k) * 2 : cst2 + ((size_t)n * (C2 / cw) + (k - kc1)) * 2;
          sum1 += e2[0]; sum2 += e2[1];
        }
      } else {
        sum1 = stats[((size_t)n * G + g) * 2];
        sum2 = stats[((size_t)n * G + g) * 2 + 1];
      }
      const double m = sum1 * inv_cnt;
      double var = fma(-m, m, sum2 * inv_cnt);
      if (var < 0.0) var = 0.0;
      mean_f = (float)m;
      rstd = rsqrtf((float)var + eps);
      // one Newton step brings rsqrtf (2 ulp) to fp32 round-off for the 1e-4 validation mode
      rstd = rstd * (1.5f - 0.5f * ((float)var + eps) * rstd * rstd);
      g_prev = g;
    }
    const float ga = __ldg(gamma + c), be = __ldg(beta + c);
    sc[e] = rstd * ga;
    sf[e] = be - mean_f * rstd * ga;
  }
  const long long v0 = (long long)(gridDim.x - 1 - blockIdx.x) * vox_per_block;
  const long long v1 = min(v0 + vox_per_block, S);
  const size_t off = ((size_t)n * S) * C + (size_t)cv * 8;
  const bool second = cv * 8 >= C1;
  const int Cs = second ? C2 : C1;
  const T* src = (second ? x2 : x) + ((size_t)n * S) * Cs + (size_t)(second ? cv * 8 - C1 : cv * 8);
  long long v = v0 + vslot;
  for (; v + (long long)(UNROLL - 1) * vpp < v1; v += (long long)UNROLL * vpp) {
    float f[UNROLL][8];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) Vec8<T>::load(src + (size_t)(v + (long long)u * vpp) * Cs, f[u]);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float t = fmaf(f[u][e], sc[e], sf[e]);
        f[u][e] = silu ? silu_for<T>(t) : t;
      }
      Vec8<T>::store(y + off + (size_t)(v + (long long)u * vpp) * C, f[u]);
    }
  }
  for (; v < v1; v += vpp) {
    float f[8];
    Vec8<T>::load(src + (size_t)v * Cs, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float t = fmaf(f[e], sc[e], sf[e]);
      f[e] = silu ? silu_for<T>(t) : t;
    }
    Vec8<T>::store(y + off + (size_t)v * C, f);
  }
}

// Tiny tensors (the 4^3 / 5^3 levels: one (sample, group) slab is a few KB): statistics and normalisation in ONE launch,
// one CTA per (group, sample); the slab is read twice (the second read hits L1/L2).  ncu, 8 x 125 voxels x 1024
// channels: 7.0 us against 5.7 + 7.1 us (+ memset) for the stats + apply pair.  From ~700 voxels on the pair wins
// (12.4 us against 11.7 / 16.8 us at 693 / 1000 voxels) because a slab is a 32-byte sliver of every row; a
// sample-in-shared-memory cluster kernel (8 CTAs per sample, DSMEM reduction) was tried and was slower still
// (34 us at 693 voxels: too few threads in flight) - so this path is only taken up to 256 voxels.
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
gn_small_kernel(const T* __restrict__ x, const T* __restrict__ x2, T* __restrict__ y, const float* __restrict__ gamma,
                const float* __restrict__ beta, int S, int C1, int C2, int G, float eps, int silu) {
  __shared__ double red[2][kNormThreads / 32];
  __shared__ float s_mean, s_rstd;
  const int g = blockIdx.x, n = blockIdx.y;
  const int C = C1 + C2, cpg = C / G, vpg = cpg >> 3;   // 8-channel vectors per group
  const int items = S * vpg;
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int v = i / vpg, c = g * cpg + (i - v * vpg) * 8;
    const T* src = c < C1 ? x + ((size_t)n * S + v) * C1 + c : x2 + ((size_t)n * S + v) * C2 + (c - C1);
    float f[8];
    Vec8<T>::load(src, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { a += f[e]; b = fmaf(f[e], f[e], b); }
  }
  double da = (double)warp_sum(a), db = (double)warp_sum(b);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = da; red[1][threadIdx.x >> 5] = db; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s1 = 0.0, s2 = 0.0;
    for (int w = 0; w < kNormThreads / 32; ++w) { s1 += red[0][w]; s2 += red[1][w]; }
    const double inv = 1.0 / ((double)S * cpg);
    const double m = s1 * inv;
    double var = fma(-m, m, s2 * inv);
    if (var < 0.0) var = 0.0;
    float r = rsqrtf((float)var + eps);
    r = r * (1.5f - 0.5f * ((float)var + eps) * r * r);
    s_mean = (float)m;
    s_rstd = r;
  }
  __syncthreads();
  const float mean = s_mean, rstd = s_rstd;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int v = i / vpg, c = g * cpg + (i - v * vpg) * 8;
    const T* src = c < C1 ? x + ((size_t)n * S + v) * C1 + c : x2 + ((size_t)n * S + v) * C2 + (c - C1);
    float f[8];
    Vec8<T>::load(src, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float ga = __ldg(gamma + c + e) * rstd;
      const float t = fmaf(f[e] - mean, ga, __ldg(beta + c + e));
      f[e] = silu ? silu_for<T>(t) : t;
    }
    Vec8<T>::store(y + ((size_t)n * S + v) * C + c, f);
  }
}

// gn_small_kernel whose first source has not been finalised yet: it is still the fp32 split-K slabs of the conv that
// produced it.  The CTA of (group, sample) sums the slabs for its cpg channels x S voxels, adds bias / temb / residual,
// rounds to bf16 (the value every other consumer of the tensor sees - written to y_raw when asked), keeps those values in
// registers, reduces the statistics and normalises from the registers: conv -> finalize -> GroupNorm becomes
// conv -> GroupNorm at the 4^3 / 5^3 levels, where each launch costs about as much as the work in it.
// 512 threads so that a 125-voxel x 32-channel slab is ONE 8-channel vector per thread, and all of a thread's slab loads
// are in flight together (4 per batch): the kernel is a chain of L2 round trips, not bandwidth (first version, 256
// threads with a serial loop over the slabs: 10-12.7 us per launch against 5.6 us for gn_small alone).
constexpr int kSlabThreads = 512;
constexpr int kSlabItems = 4;   // 8-channel vectors per thread: S * (cpg / 8) <= 512 * 4
__global__ void __launch_bounds__(kSlabThreads)
gn_small_slabs_kernel(const float* __restrict__ slabs, int splits, long long slab_elems, const float* __restrict__ bias,
                      const float* __restrict__ temb, int temb_stride, const __nv_bfloat16* __restrict__ residual,
                      __nv_bfloat16* __restrict__ y_raw, const __nv_bfloat16* __restrict__ x2,
                      __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                      int S, int C1, int C2, int G, float eps, int silu) {
  __shared__ double red[2][kSlabThreads / 32];
  __shared__ float s_mean, s_rstd;
  const int g = blockIdx.x, n = blockIdx.y;
  const int C = C1 + C2, cpg = C / G, vpg = cpg >> 3;
  const int items = S * vpg;
  const bool from_slabs = g * cpg < C1;       // no group straddles the sources (checked on the host)
  float val[kSlabItems][8];
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int k = 0; k < kSlabItems; ++k) {
    const int i = threadIdx.x + k * kSlabThreads;
    if (i < items) {
      const int v = i / vpg, c = g * cpg + (i - v * vpg) * 8;
      float f[8];
      if (from_slabs) {
        const size_t off = ((size_t)n * S + v) * C1 + c;
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
        for (int z0 = 0; z0 < splits; z0 += 4) {     // 4 slab loads in flight, summed in slab order (as the finalize kernel)
          float t[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (z0 + u < splits) Vec8<float>::load(slabs + (size_t)(z0 + u) * slab_elems + off, t[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (z0 + u < splits) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += t[u][e];
            }
        }
        if (bias) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += __ldg(bias + c + e);
        }
        if (temb) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += __ldg(temb + (size_t)n * temb_stride + c + e);
        }
        if (residual) {
          float t[8];
          Vec8<__nv_bfloat16>::load(residual + off, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += t[e];
        }
        // same operation order and the same bf16 rounding as splitk_finalize_kernel: bit-identical tensor
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(__float2bfloat16_rn(f[e]));
        if (y_raw) Vec8<__nv_bfloat16>::store(y_raw + off, f);
      } else {
        Vec8<__nv_bfloat16>::load(x2 + ((size_t)n * S + v) * C2 + (c - C1), f);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { val[k][e] = f[e]; a += f[e]; b = fmaf(f[e], f[e], b); }
    }
  }
  double da = (double)warp_sum(a), db = (double)warp_sum(b);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = da; red[1][threadIdx.x >> 5] = db; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s1 = 0.0, s2 = 0.0;
    for (int w = 0; w < kSlabThreads / 32; ++w) { s1 += red[0][w]; s2 += red[1][w]; }
    const double inv = 1.0 / ((double)S * cpg);
    const double m = s1 * inv;
    double var = fma(-m, m, s2 * inv);
    if (var < 0.0) var = 0.0;
    float r = rsqrtf((float)var + eps);
    r = r * (1.5f - 0.5f * ((float)var + eps) * r * r);
    s_mean = (float)m;
    s_rstd = r;
  }
  __syncthreads();
  const float mean = s_mean, rstd = s_rstd;
#pragma unroll
  for (int k = 0; k < kSlabItems; ++k) {
    const int i = threadIdx.x + k * kSlabThreads;
    if (i < items) {
      const int v = i / vpg, c = g * cpg + (i - v * vpg) * 8;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float ga = __ldg(gamma + c + e) * rstd;
        const float t = fmaf(val[k][e] - mean, ga, __ldg(beta + c + e));
        f[e] = silu ? silu_for<__nv_bfloat16>(t) : t;
      }
      Vec8<__nv_bfloat16>::store(y + ((size_t)n * S + v) * C + c, f);
    }
  }
}

static int norm_geometry(int N, long long S, int C, int G, int* vox_per_block, dim3* grid) {
  B2_REQUIRE(C % 8 == 0 && C / 8 <= kNormThreads, B200LDM_BAD_SHAPE, "GroupNorm needs C %% 8 == 0 and C <= %d (got %d)",
             8 * kNormThreads, C);
  B2_REQUIRE(G > 0 && C % G == 0 && G <= 1024, B200LDM_BAD_SHAPE, "bad group count %d for C=%d", G, C);
  const int vpp = kNormThreads / (C / 8);
  // ~16 blocks per SM across the batch (8 resident x 2 waves); each block at least 8 passes
  long long want_blocks = (148LL * 16 + N - 1) / N;
  long long vpb = (S + want_blocks - 1) / want_blocks;
  if (vpb < 8LL * vpp) vpb = 8LL * vpp;
  vpb = (vpb + vpp - 1) / vpp * vpp;
  *vox_per_block = (int)vpb;
  *grid = dim3((unsigned)((S + vpb - 1) / vpb), (unsigned)N, 1);
  return B200LDM_OK;
}

}  // namespace b200ldm

extern "C" {

int b200ldm_groupnorm_stats(const void* x, const void* x2, double* stats, int32_t N, int64_t S, int32_t C1,
                            int32_t C2, int32_t G, int32_t dtype, b200ldm_stream stream) {
  using namespace b200ldm;
  B2_REQUIRE(x && stats && N > 0 && S > 0, B200LDM_BAD_ARG, "groupnorm_stats: bad arguments");
  B2_REQUIRE(aligned16(x) && aligned16(x2), B200LDM_BAD_ALIGN, "groupnorm_stats: x must be 16-byte aligned");
  B2_REQUIRE(C1 > 0 && C1 % 8 == 0 && C2 >= 0 && C2 % 8 == 0 && ((C2 > 0) == (x2 != nullptr)), B200LDM_BAD_SHAPE,
             "groupnorm_stats: C1/C2 must be multiples of 8 and match x2");
  int vpb; dim3 grid;
  int rc = norm_geometry(N, S, C1 + C2, G, &vpb, &grid);
  if (rc) return rc;
  const size_t sh = 2 * G * sizeof(float);
  if (dtype == B200LDM_BF16)
    gn_stats_kernel<__nv_bfloat16><<<grid, kNormThreads, sh, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)x2, stats, S, C1, C2, G, vpb);
  else if (dtype == B200LDM_F32)
    gn_stats_kernel<float><<<grid, kNormThreads, sh, (cudaStream_t)stream>>>((const float*)x, (const float*)x2, stats,
                                                                             S, C1, C2, G, vpb);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

static int gn_apply_impl(const void* x, const void* x2, void* y, const double* stats, const double* cst1,
                         const double* cst2, const float* gamma, const float* beta, int32_t N, int64_t S, int32_t C1,
                         int32_t C2, int32_t G, float eps, int32_t silu, int32_t dtype, b200ldm_stream stream,
                         int32_t cw = 8) {
  using namespace b200ldm;
  B2_REQUIRE(x && y && (stats || cst1) && gamma && beta && N > 0 && S > 0, B200LDM_BAD_ARG,
             "groupnorm_apply: bad arguments");
  if (cst1)
    B2_REQUIRE(G > 0 && (C1 + C2) % G == 0 && (cw == 2 || cw == 4 || cw == 8) && ((C1 + C2) / G) % cw == 0 &&
                   C1 % cw == 0 && ((C2 > 0) == (cst2 != nullptr)),
               B200LDM_BAD_SHAPE,
               "groupnorm_apply_chunked: channels/group must be a multiple of the chunk width (2, 4 or 8) and cstats2 "
               "match x2");
  B2_REQUIRE(aligned16(x) && aligned16(y) && aligned16(x2), B200LDM_BAD_ALIGN,
             "groupnorm_apply: x/y must be 16-byte aligned");
  B2_REQUIRE(C1 > 0 && C1 % 8 == 0 && C2 >= 0 && C2 % 8 == 0 && ((C2 > 0) == (x2 != nullptr)), B200LDM_BAD_SHAPE,
             "groupnorm_apply: C1/C2 must be multiples of 8 and match x2");
  int vpb; dim3 grid;
  int rc = norm_geometry(N, S, C1 + C2, G, &vpb, &grid);
  if (rc) return rc;
  const double inv_cnt = 1.0 / ((double)S * (double)((C1 + C2) / G));
  static const int variant = [] { const char* e = getenv("B200LDM_GN_VARIANT"); return e ? atoi(e) : 0; }();
#define GN_APPLY_BF16(U, B)                                                                                      \
  gn_apply_kernel<__nv_bfloat16, U, B><<<grid, kNormThreads, 0, (cudaStream_t)stream>>>(                         \
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)x2, (__nv_bfloat16*)y, stats, gamma, beta, S, C1, C2, G, eps, \
      silu, vpb, inv_cnt, cst1, cst2, cw)
  if (dtype == B200LDM_BF16) {
    switch (variant) {
      case 1: GN_APPLY_BF16(1, 8); break;
      case 2: GN_APPLY_BF16(2, 6); break;
      case 3: GN_APPLY_BF16(2, 4); break;
      case 4: GN_APPLY_BF16(4, 4); break;
      case 5: GN_APPLY_BF16(8, 2); break;
      case 6: GN_APPLY_BF16(4, 3); break;
      default: GN_APPLY_BF16(4, 4); break;  // best of the on-GPU sweep (scripts/gn_bench.py): 5.1-5.8 TB/s
    }
  } else if (dtype == B200LDM_F32)
    gn_apply_kernel<float, 4, 3><<<grid, kNormThreads, 0, (cudaStream_t)stream>>>(
        (const float*)x, (const float*)x2, (float*)y, stats, gamma, beta, S, C1, C2, G, eps, silu, vpb, inv_cnt, cst1,
        cst2, cw);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_groupnorm_apply(const void* x, const void* x2, void* y, const double* stats, const float* gamma,
                            const float* beta, int32_t N, int64_t S, int32_t C1, int32_t C2, int32_t G, float eps,
                            int32_t silu, int32_t dtype, b200ldm_stream stream) {
  B2_REQUIRE(stats, B200LDM_BAD_ARG, "groupnorm_apply: null stats");
  return gn_apply_impl(x, x2, y, stats, nullptr, nullptr, gamma, beta, N, S, C1, C2, G, eps, silu, dtype, stream);
}

int b200ldm_groupnorm_apply_chunked(const void* x, const void* x2, void* y, const double* cstats1,
                                    const double* cstats2, const float* gamma, const float* beta, int32_t N, int64_t S,
                                    int32_t C1, int32_t C2, int32_t G, float eps, int32_t silu, int32_t dtype,
                                    b200ldm_stream stream) {
  B2_REQUIRE(cstats1, B200LDM_BAD_ARG, "groupnorm_apply_chunked: null cstats1");
  return gn_apply_impl(x, x2, y, nullptr, cstats1, cstats2, gamma, beta, N, S, C1, C2, G, eps, silu, dtype, stream);
}

int b200ldm_groupnorm_apply_chunked_w(const void* x, const void* x2, void* y, const double* cstats1,
                                      const double* cstats2, const float* gamma, const float* beta, int32_t N,
                                      int64_t S, int32_t C1, int32_t C2, int32_t G, float eps, int32_t silu,
                                      int32_t dtype, int32_t chunk_width, b200ldm_stream stream) {
  B2_REQUIRE(cstats1, B200LDM_BAD_ARG, "groupnorm_apply_chunked: null cstats1");
  return gn_apply_impl(x, x2, y, nullptr, cstats1, cstats2, gamma, beta, N, S, C1, C2, G, eps, silu, dtype, stream,
                       chunk_width);
}

int b200ldm_groupnorm_small_fits(int64_t S, int32_t C, int32_t G, int32_t dtype) {
  (void)dtype;
  return (S > 0 && S <= 4096 && C > 0 && C % 8 == 0 && G > 0 && C % G == 0 && (C / G) % 8 == 0) ? 1 : 0;
}

int b200ldm_groupnorm_small(const void* x, const void* x2, void* y, const float* gamma, const float* beta, int32_t N,
                            int64_t S, int32_t C1, int32_t C2, int32_t G, float eps, int32_t silu, int32_t dtype,
                            b200ldm_stream stream) {
  using namespace b200ldm;
  B2_REQUIRE(x && y && gamma && beta && N > 0 && S > 0 && S <= 4096, B200LDM_BAD_ARG,
             "groupnorm_small: bad arguments (S must be <= 4096)");
  B2_REQUIRE(C1 > 0 && C1 % 8 == 0 && C2 >= 0 && C2 % 8 == 0 && ((C2 > 0) == (x2 != nullptr)) && G > 0 &&
                 (C1 + C2) % G == 0 && ((C1 + C2) / G) % 8 == 0,
             B200LDM_BAD_SHAPE, "groupnorm_small: C1/C2 multiples of 8 matching x2, channels/group a multiple of 8");
  B2_REQUIRE(aligned16(x) && aligned16(y) && aligned16(x2), B200LDM_BAD_ALIGN, "groupnorm_small: 16-byte alignment");
  const dim3 grid((unsigned)G, (unsigned)N, 1);
  if (dtype == B200LDM_BF16)
    gn_small_kernel<__nv_bfloat16><<<grid, kNormThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)x2, (__nv_bfloat16*)y, gamma, beta, (int)S, C1, C2, G, eps, silu);
  else if (dtype == B200LDM_F32)
    gn_small_kernel<float><<<grid, kNormThreads, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)x2, (float*)y,
                                                                            gamma, beta, (int)S, C1, C2, G, eps, silu);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_groupnorm_small_slabs(const float* slabs, int32_t splits, const float* bias, const float* temb,
                                  int32_t temb_stride, const void* residual, void* y_raw, const void* x2, void* y,
                                  const float* gamma, const float* beta, int32_t N, int64_t S, int32_t C1, int32_t C2,
                                  int32_t G, float eps, int32_t silu, b200ldm_stream stream) {
  using namespace b200ldm;
  B2_REQUIRE(slabs && y && gamma && beta && splits >= 1 && N > 0 && S > 0 && S <= 4096, B200LDM_BAD_ARG,
             "groupnorm_small_slabs: bad arguments");
  B2_REQUIRE(C1 > 0 && C1 % 8 == 0 && C2 >= 0 && C2 % 8 == 0 && ((C2 > 0) == (x2 != nullptr)) && G > 0 &&
                 (C1 + C2) % G == 0 && ((C1 + C2) / G) % 8 == 0 && C1 % ((C1 + C2) / G) == 0,
             B200LDM_BAD_SHAPE,
             "groupnorm_small_slabs: C1/C2 multiples of 8 matching x2, channels/group a multiple of 8 dividing C1");
  B2_REQUIRE(S * (((C1 + C2) / G) / 8) <= (int64_t)kSlabThreads * kSlabItems, B200LDM_BAD_SHAPE,
             "groupnorm_small_slabs: a (group, sample) slab must fit the CTA's registers (S * cpg / 8 <= %d)",
             kSlabThreads * kSlabItems);
  B2_REQUIRE(aligned16(slabs) && aligned16(y) && aligned16(x2) && aligned16(residual) && aligned16(y_raw),
             B200LDM_BAD_ALIGN, "groupnorm_small_slabs: 16-byte alignment");
  const dim3 grid((unsigned)G, (unsigned)N, 1);
  gn_small_slabs_kernel<<<grid, kSlabThreads, 0, (cudaStream_t)stream>>>(
      slabs, splits, (long long)N * S * C1, bias, temb, temb_stride, (const __nv_bfloat16*)residual,
      (__nv_bfloat16*)y_raw, (const __nv_bfloat16*)x2, (__nv_bfloat16*)y, gamma, beta, (int)S, C1, C2, G, eps, silu);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

}  // extern "C"
