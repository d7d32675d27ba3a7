// tokens.cu - token-major [M][C] helpers of the cross-attention path (MONAI SpatialTransformer /
// DiffusionUNetTransformerBlock): LayerNorm and the GEGLU gate.  Both are one read + one
// write, HBM-bound: 16-byte vectors, one warp per token row, no shared memory.
#include "common.cuh"

namespace b200ldm {

constexpr int kLnThreads = 256;  // 8 rows per CTA

// One warp per row; the row stays in registers (NCH chunks of 8 channels per lane) between the mean, the centred
// variance and the normalisation, so x is read exactly once.
template <typename T, int NCH>
__global__ void __launch_bounds__(kLnThreads)
layernorm_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                 const float* __restrict__ beta, long long M, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
  if (row >= M) return;
  const T* xr = x + row * C;
  float v[NCH][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < C) {
      Vec8<T>::load(xr + c, v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < C) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  T* yr = y + row * C;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < C) {
      float g[8], b[8], o[8];
      Vec8<float>::load(gamma + c, g);
      Vec8<float>::load(beta + c, b);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf((v[i][e] - mean) * rstd, g[e], b[e]);
      Vec8<T>::store(yr + c, o);
    }
  }
}

template <typename T, int NCH>
static void launch_ln(const void* x, void* y, const float* gamma, const float* beta, long long M, int C, float eps,
                      cudaStream_t st) {
  const unsigned blocks = (unsigned)((M + kLnThreads / 32 - 1) / (kLnThreads / 32));
  layernorm_kernel<T, NCH><<<blocks, kLnThreads, 0, st>>>((const T*)x, (T*)y, gamma, beta, M, C, eps);
}

template <typename T>
static int layernorm_dispatch(const void* x, void* y, const float* gamma, const float* beta, long long M, int C, float eps,
                              cudaStream_t st) {
  const int nch = ceil_div(C, 256);
  if (nch <= 1) launch_ln<T, 1>(x, y, gamma, beta, M, C, eps, st);
  else if (nch <= 2) launch_ln<T, 2>(x, y, gamma, beta, M, C, eps, st);
  else if (nch <= 4) launch_ln<T, 4>(x, y, gamma, beta, M, C, eps, st);
  else if (nch <= 8) launch_ln<T, 8>(x, y, gamma, beta, M, C, eps, st);
  else launch_ln<T, 16>(x, y, gamma, beta, M, C, eps, st);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// y[m][j] = h[m][j] * gelu(h[m][J + j]); one thread per 8 output channels
template <typename T>
__global__ void __launch_bounds__(256)
geglu_kernel(const T* __restrict__ h, T* __restrict__ y, long long M, int J) {
  const int jv = J >> 3;
  const long long total = M * jv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / jv;
    const int j = (int)(i - m * jv) * 8;
    float a[8], g[8], o[8];
    Vec8<T>::load(h + m * 2 * J + j, a);
    Vec8<T>::load(h + m * 2 * J + J + j, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = a[e] * gelu_erf(g[e]);
    Vec8<T>::store(y + m * J + j, o);
  }
}

// out[n][:] = a[n][:] + table[idx[n]][:]  (class embedding added to the time embedding, fp32)
__global__ void embedding_add_kernel(const float* __restrict__ a, const float* __restrict__ table,
                                     const long long* __restrict__ idx, float* __restrict__ out, int n_rows, int dim,
                                     int n_classes) {
  const int r = blockIdx.x;
  long long k = idx[r];
  k = k < 0 ? 0 : (k >= n_classes ? n_classes - 1 : k);
  for (int c = threadIdx.x; c < dim; c += blockDim.x) out[(size_t)r * dim + c] = a[(size_t)r * dim + c] + table[k * dim + c];
}

}  // namespace b200ldm

extern "C" {

int b200ldm_layernorm(const void* x, void* y, const float* gamma, const float* beta, int64_t M, int32_t C, float eps,
                      int32_t dtype, b200ldm_stream stream) {
  using namespace b200ldm;
  B2_REQUIRE(x && y && gamma && beta, B200LDM_BAD_ARG, "layernorm: null pointer");
  B2_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && C <= 4096, B200LDM_BAD_SHAPE,
             "layernorm: need 0 < C <= 4096, C %% 8 == 0 (got M=%lld C=%d)", (long long)M, C);
  B2_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), B200LDM_BAD_ALIGN,
             "layernorm: pointers must be 16-byte aligned");
  if (dtype == B200LDM_BF16) return layernorm_dispatch<__nv_bfloat16>(x, y, gamma, beta, M, C, eps, (cudaStream_t)stream);
  if (dtype == B200LDM_F32) return layernorm_dispatch<float>(x, y, gamma, beta, M, C, eps, (cudaStream_t)stream);
  return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
}

int b200ldm_geglu(const void* h, void* y, int64_t M, int32_t J, int32_t dtype, b200ldm_stream stream) {
  using namespace b200ldm;
  B2_REQUIRE(h && y, B200LDM_BAD_ARG, "geglu: null pointer");
  B2_REQUIRE(M > 0 && J > 0 && J % 8 == 0, B200LDM_BAD_SHAPE, "geglu: need J %% 8 == 0 (got M=%lld J=%d)", (long long)M, J);
  B2_REQUIRE(aligned16(h) && aligned16(y), B200LDM_BAD_ALIGN, "geglu: pointers must be 16-byte aligned");
  const long long total = (long long)M * (J / 8);
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 16);
  if (dtype == B200LDM_BF16)
    geglu_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)h, (__nv_bfloat16*)y, M, J);
  else if (dtype == B200LDM_F32)
    geglu_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)h, (float*)y, M, J);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_embedding_add(const float* a, const float* table, const int64_t* idx, float* out, int32_t n_rows,
                          int32_t dim, int32_t n_classes, b200ldm_stream stream) {
  using namespace b200ldm;
  B2_REQUIRE(a && table && idx && out && n_rows > 0 && dim > 0 && n_classes > 0, B200LDM_BAD_ARG,
             "embedding_add: bad arguments");
  embedding_add_kernel<<<n_rows, 256, 0, (cudaStream_t)stream>>>(a, table, (const long long*)idx, out, n_rows, dim,
                                                                n_classes);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

}  // extern "C"
