// attention_simt.cu - CUDA-core flash-style attention (online softmax, O(L) memory), fp32 math.
// The fp32 VALIDATION engine for SpatialAttentionBlock and the cross-check of the tcgen05
// attention kernel.  One query per `DSPLIT = d / DS` adjacent lanes, each lane owning a DS-wide
// slice of the head dimension; K/V tiles are staged in shared memory as fp32.
#include "common.cuh"

namespace b200ldm {

constexpr int kAttThreads = 128;

template <typename T, int DS>
__global__ void __launch_bounds__(kAttThreads)
attention_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
                      int H, int L, int Lk, int d, long long ld, long long ldkv, float scale, int KT) {
  extern __shared__ float smem[];
  float* Ks = smem;                 // [KT][d]
  float* Vs = smem + (size_t)KT * d;
  const int dsplit = d / DS;
  const int qpb = kAttThreads / dsplit;
  const int b = blockIdx.z, h = blockIdx.y;
  const int ql = threadIdx.x / dsplit, part = threadIdx.x % dsplit;
  const int qi = blockIdx.x * qpb + ql;
  const bool qvalid = qi < L;
  const size_t head_off = (size_t)h * d;
  const T* qrow = q + ((size_t)b * L + (qvalid ? qi : 0)) * ld + head_off + (size_t)part * DS;

  float qr[DS], acc[DS];
#pragma unroll
  for (int i = 0; i < DS; ++i) {
    qr[i] = to_f32<T>(qrow[i]) * scale;
    acc[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;

  for (int k0 = 0; k0 < Lk; k0 += KT) {
    const int kn = min(KT, Lk - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kn * d; i += kAttThreads) {
      const int r = i / d, c = i - r * d;
      const size_t g = ((size_t)b * Lk + k0 + r) * ldkv + head_off + c;
      Ks[i] = to_f32<T>(k[g]);
      Vs[i] = to_f32<T>(v[g]);
    }
    __syncthreads();
    for (int j = 0; j < kn; ++j) {
      const float* kr = Ks + (size_t)j * d + part * DS;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < DS; ++i) s = fmaf(qr[i], kr[i], s);
      for (int off = 1; off < dsplit; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn);
      const float p = __expf(s - mn);
      l = l * corr + p;
      const float* vr = Vs + (size_t)j * d + part * DS;
#pragma unroll
      for (int i = 0; i < DS; ++i) acc[i] = fmaf(acc[i], corr, p * vr[i]);
      m = mn;
    }
  }
  if (qvalid) {
    const float inv = 1.0f / l;
    T* orow = o + ((size_t)b * L + qi) * ((size_t)H * d) + head_off + (size_t)part * DS;
#pragma unroll
    for (int i = 0; i < DS; ++i) orow[i] = from_f32<T>(acc[i] * inv);
  }
}

template <typename T>
static int launch_attention_simt(const void* q, const void* k, const void* v, void* o, int B, int H, int L, int Lk,
                                 int d, long long ld, long long ldkv, float scale, cudaStream_t st) {
  const int DS = (d % 64 == 0) ? 64 : 32;
  const int dsplit = d / DS;
  B2_REQUIRE(d % 32 == 0 && dsplit <= 32 && (dsplit & (dsplit - 1)) == 0, B200LDM_BAD_SHAPE,
             "attention: head dim %d unsupported (need 32*2^k)", d);
  const int KT = d <= 64 ? 32 : (d <= 128 ? 32 : 16);
  const size_t smem = (size_t)2 * KT * d * sizeof(float);
  const int qpb = kAttThreads / dsplit;
  dim3 grid((unsigned)ceil_div(L, qpb), (unsigned)H, (unsigned)B);
  if (DS == 64) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(attention_simt_kernel<T, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    attention_simt_kernel<T, 64><<<grid, kAttThreads, smem, st>>>((const T*)q, (const T*)k, (const T*)v, (T*)o, H, L,
                                                                 Lk, d, ld, ldkv, scale, KT);
  } else {
    attention_simt_kernel<T, 32><<<grid, kAttThreads, smem, st>>>((const T*)q, (const T*)k, (const T*)v, (T*)o, H, L,
                                                                 Lk, d, ld, ldkv, scale, KT);
  }
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int attention_tcgen05(const void* q, const void* k, const void* v, void* o, int B, int H, int L, int Lk, int d,
                      long long ld, long long ldkv, float scale, void* ws, long long ws_bytes, cudaStream_t st);
long long attention_tcgen05_workspace_bytes(int B, int H, int L, int Lk, int d);
bool attention_tcgen05_supported(int d);

}  // namespace b200ldm

extern "C" int64_t b200ldm_attention_workspace_bytes(int32_t B, int32_t H, int32_t Lq, int32_t Lk, int32_t d, int32_t dtype,
                                                     int32_t engine) {
  if (dtype != B200LDM_BF16 || (engine != B200LDM_ENGINE_AUTO && engine != B200LDM_ENGINE_TCGEN05)) return 0;
  return b200ldm::attention_tcgen05_workspace_bytes(B, H, Lq, Lk, d);
}

extern "C" int b200ldm_cross_attention_ws(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t H,
                                          int32_t Lq, int32_t Lk, int32_t d, int64_t ldq, int64_t ldkv, float scale,
                                          int32_t dtype, int32_t engine, void* workspace, int64_t workspace_bytes,
                                          b200ldm_stream stream) {
  using namespace b200ldm;
  B2_REQUIRE(!workspace || aligned16(workspace), B200LDM_BAD_ALIGN, "attention: workspace must be 16-byte aligned");
  B2_REQUIRE(q && k && v && o, B200LDM_BAD_ARG, "attention: null pointer");
  B2_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0 && d > 0 && ldq >= (int64_t)H * d && ldkv >= (int64_t)H * d,
             B200LDM_BAD_SHAPE, "attention: bad shape");
  if (engine == B200LDM_ENGINE_AUTO) {
    // bf16 is the tensor-core product path: a head dim without a tcgen05 instance is an error, never a silent drop to
    // the CUDA-core validation engine (which stays reachable by asking for B200LDM_ENGINE_SIMT explicitly)
    B2_REQUIRE(dtype != B200LDM_BF16 || attention_tcgen05_supported(d), B200LDM_BAD_SHAPE,
               "attention: no tcgen05 kernel for head dim %d (32 / 64 / 128 / 256); the SIMT validation engine must be "
               "requested explicitly", d);
    engine = dtype == B200LDM_BF16 ? B200LDM_ENGINE_TCGEN05 : B200LDM_ENGINE_SIMT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (engine == B200LDM_ENGINE_TCGEN05) {
    B2_REQUIRE(dtype == B200LDM_BF16, B200LDM_BAD_ARG, "tcgen05 attention needs bf16");
    return attention_tcgen05(q, k, v, o, B, H, Lq, Lk, d, ldq, ldkv, scale, workspace, workspace_bytes, st);
  }
  if (dtype == B200LDM_BF16) return launch_attention_simt<__nv_bfloat16>(q, k, v, o, B, H, Lq, Lk, d, ldq, ldkv, scale, st);
  if (dtype == B200LDM_F32) return launch_attention_simt<float>(q, k, v, o, B, H, Lq, Lk, d, ldq, ldkv, scale, st);
  return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
}

extern "C" int b200ldm_cross_attention(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t H,
                                       int32_t Lq, int32_t Lk, int32_t d, int64_t ldq, int64_t ldkv, float scale,
                                       int32_t dtype, int32_t engine, b200ldm_stream stream) {
  return b200ldm_cross_attention_ws(q, k, v, o, B, H, Lq, Lk, d, ldq, ldkv, scale, dtype, engine, nullptr, 0, stream);
}

extern "C" int b200ldm_attention(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t H, int32_t L,
                                 int32_t d, int64_t ld, float scale, int32_t dtype, int32_t engine,
                                 b200ldm_stream stream) {
  return b200ldm_cross_attention_ws(q, k, v, o, B, H, L, L, d, ld, ld, scale, dtype, engine, nullptr, 0, stream);
}

extern "C" int b200ldm_attention_ws(const void* q, const void* k, const void* v, void* o, int32_t B, int32_t H, int32_t L,
                                    int32_t d, int64_t ld, float scale, int32_t dtype, int32_t engine, void* workspace,
                                    int64_t workspace_bytes, b200ldm_stream stream) {
  return b200ldm_cross_attention_ws(q, k, v, o, B, H, L, L, d, ld, ld, scale, dtype, engine, workspace, workspace_bytes,
                                    stream);
}
