// conv_simt.cu - CUDA-core Conv3d (fp32 accumulate, any channel count) used as the fp32
// VALIDATION engine (north-star: "1e-4 in an fp32 validation mode") and as an independent
// cross-check of the tcgen05 kernel.  Same operands, layouts and fused epilogue as conv_tc.cu.
// Classic 64x64x16 shared-memory tile, 4x4 outputs per thread; the A tile is gathered
// (implicit GEMM) with zero fill for out-of-range taps.
#include "common.cuh"

namespace b200ldm {

struct ConvSimtArgs {
  const void* x; const void* x2; const void* w;
  const float* bias; const float* temb; const void* residual; void* y;
  double* gn_stats;
  int N, D, H, W, C1, C2, Cout, Cout_rows;
  int Do, Ho, Wo;
  int ksize, stride, pad_lo, temb_stride, gn_groups, out_f32, up;
  long long M;  // N*Do*Ho*Wo
};

constexpr int SM_TM = 64, SM_TN = 64, SM_TK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
conv3d_simt_kernel(const ConvSimtArgs a) {
  __shared__ float As[SM_TK][SM_TM + 1];
  __shared__ float Bs[SM_TK][SM_TN + 1];
  __shared__ long long row_off[SM_TM];  // input voxel index for the current tap, or -1

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * SM_TM;
  const int co0 = blockIdx.y * SM_TN;
  const int Cin = a.C1 + a.C2;
  const int taps = a.ksize * a.ksize * a.ksize;
  const T* x = reinterpret_cast<const T*>(a.x);
  const T* x2 = reinterpret_cast<const T*>(a.x2);
  const T* wgt = reinterpret_cast<const T*>(a.w);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < taps; ++tap) {
    int kd = 0, kh = 0, kw = 0;
    if (a.ksize == 3) { kd = tap / 9; kh = (tap / 3) % 3; kw = tap % 3; }
    __syncthreads();
    if (tid < SM_TM) {
      long long m = m0 + tid;
      long long off = -1;
      if (m < a.M) {
        int wo = (int)(m % a.Wo); long long t = m / a.Wo;
        int ho = (int)(t % a.Ho); t /= a.Ho;
        int dd = (int)(t % a.Do); int n = (int)(t / a.Do);
        int iw = wo * a.stride + kw - a.pad_lo;
        int ih = ho * a.stride + kh - a.pad_lo;
        int id = dd * a.stride + kd - a.pad_lo;
        if (a.up) {  // the conv runs on the (virtual) nearest-upsampled tensor: bounds in 2x space, then >> 1
          const bool in = iw >= 0 && iw < 2 * a.W && ih >= 0 && ih < 2 * a.H && id >= 0 && id < 2 * a.D;
          iw = in ? iw >> 1 : -1; ih >>= 1; id >>= 1;
        }
        if (iw >= 0 && iw < a.W && ih >= 0 && ih < a.H && id >= 0 && id < a.D)
          off = (((long long)n * a.D + id) * a.H + ih) * a.W + iw;
      }
      row_off[tid] = off;
    }
    __syncthreads();
    for (int c0 = 0; c0 < Cin; c0 += SM_TK) {
      // A tile: 64 rows x 16 channels, 4 per thread
      {
        const int r = tid >> 2, cq = (tid & 3) * 4;
        const long long off = row_off[r];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c0 + cq + e;
          float v = 0.f;
          if (off >= 0 && c < Cin) {
            v = (c < a.C1) ? to_f32<T>(x[off * a.C1 + c]) : to_f32<T>(x2[off * a.C2 + (c - a.C1)]);
          }
          As[cq + e][r] = v;
        }
      }
      // B tile: 64 couts x 16 channels
      {
        const int nn = tid >> 2, kq = (tid & 3) * 4;
        const int co = co0 + nn;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c0 + kq + e;
          float v = 0.f;
          if (co < a.Cout && c < Cin) v = to_f32<T>(wgt[((size_t)tap * a.Cout_rows + co) * Cin + c]);
          Bs[kq + e][nn] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SM_TK; ++k) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
    const int n = (int)(m / ((long long)a.Do * a.Ho * a.Wo));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= a.Cout) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[co];
      if (a.temb) v += a.temb[(size_t)(a.temb_stride ? n : 0) * a.temb_stride + co];
      const size_t o = (size_t)m * a.Cout + co;
      if (a.residual)
        v += a.out_f32 ? reinterpret_cast<const float*>(a.residual)[o]
                       : to_f32<T>(reinterpret_cast<const T*>(a.residual)[o]);
      float stored;
      if (a.out_f32) {
        reinterpret_cast<float*>(a.y)[o] = v;
        stored = v;
      } else {
        T tv = from_f32<T>(v);
        reinterpret_cast<T*>(a.y)[o] = tv;
        stored = to_f32<T>(tv);
      }
      if (a.gn_stats) {
        const int g = co / (a.Cout / a.gn_groups);
        atomicAdd(a.gn_stats + ((size_t)n * a.gn_groups + g) * 2, (double)stored);
        atomicAdd(a.gn_stats + ((size_t)n * a.gn_groups + g) * 2 + 1, (double)stored * (double)stored);
      }
    }
  }
}

int conv3d_simt(const b200ldm_conv3d_params* p, cudaStream_t st) {
  B2_REQUIRE(p->ksize == 1 || p->ksize == 3, B200LDM_BAD_SHAPE, "ksize must be 1 or 3 (got %d)", p->ksize);
  B2_REQUIRE(p->stride == 1 || p->stride == 2, B200LDM_BAD_SHAPE, "stride must be 1 or 2 (got %d)", p->stride);
  B2_REQUIRE(p->C1 > 0 && p->C2 >= 0 && p->Cout > 0, B200LDM_BAD_SHAPE, "bad channel counts");
  B2_REQUIRE((p->C2 > 0) == (p->x2 != nullptr), B200LDM_BAD_ARG, "x2 / C2 mismatch");
  if (p->gn_stats)
    B2_REQUIRE(p->gn_groups > 0 && p->Cout % p->gn_groups == 0, B200LDM_BAD_SHAPE, "bad gn_groups");
  ConvSimtArgs a{};
  a.x = p->x; a.x2 = p->x2; a.w = p->w; a.bias = p->bias; a.temb = p->temb; a.residual = p->residual; a.y = p->y;
  a.gn_stats = p->gn_stats;
  a.N = p->N; a.D = p->D; a.H = p->H; a.W = p->W; a.C1 = p->C1; a.C2 = p->C2; a.Cout = p->Cout;
  a.Cout_rows = (p->Cout + 15) / 16 * 16;
  a.Do = p->Do; a.Ho = p->Ho; a.Wo = p->Wo;
  a.ksize = p->ksize; a.stride = p->stride; a.pad_lo = p->pad_lo; a.temb_stride = p->temb_stride;
  a.gn_groups = p->gn_groups; a.out_f32 = p->out_f32; a.up = p->upsample2x ? 1 : 0;
  a.M = (long long)p->N * p->Do * p->Ho * p->Wo;
  dim3 grid((unsigned)ceil_div(a.M, SM_TM), (unsigned)ceil_div(p->Cout, SM_TN), 1);
  if (p->dtype == B200LDM_F32)
    conv3d_simt_kernel<float><<<grid, 256, 0, st>>>(a);
  else
    conv3d_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int conv3d_tcgen05(const b200ldm_conv3d_params* p, cudaStream_t st);
int conv3d_tcgen05_tune(const b200ldm_conv3d_params* p, cudaStream_t st);
int conv3d_tcgen05_tune_pending(const b200ldm_conv3d_params* p);
int conv3d_tcgen05_splits(const b200ldm_conv3d_params* p);
int conv3d_splitk_finalize(const float* slabs, int splits, long long M, int Cout, long long vox_per_sample,
                           const float* bias, const float* temb, int temb_stride, const void* residual, void* y,
                           cudaStream_t st);
long long conv3d_tcgen05_workspace_bytes(const b200ldm_conv3d_params* p);
int conv3d_tcgen05_emits_chunk_stats(const b200ldm_conv3d_params* p);

}  // namespace b200ldm

extern "C" int b200ldm_conv3d_emits_chunk_stats(const b200ldm_conv3d_params* p) {
  if (!p) return 0;
  int engine = p->engine;
  if (engine == B200LDM_ENGINE_AUTO)
    engine = (p->dtype == B200LDM_BF16) ? B200LDM_ENGINE_TCGEN05 : B200LDM_ENGINE_SIMT;
  return engine == B200LDM_ENGINE_TCGEN05 ? b200ldm::conv3d_tcgen05_emits_chunk_stats(p) : 0;
}

extern "C" int64_t b200ldm_conv3d_workspace_bytes(const b200ldm_conv3d_params* p) {
  if (!p) return 0;
  int engine = p->engine;
  if (engine == B200LDM_ENGINE_AUTO)
    engine = (p->dtype == B200LDM_BF16) ? B200LDM_ENGINE_TCGEN05 : B200LDM_ENGINE_SIMT;
  return engine == B200LDM_ENGINE_TCGEN05 ? (int64_t)b200ldm::conv3d_tcgen05_workspace_bytes(p) : 0;
}

static int conv3d_engine_of(const b200ldm_conv3d_params* p) {
  if (p->engine != B200LDM_ENGINE_AUTO) return p->engine;
  return p->dtype == B200LDM_BF16 ? B200LDM_ENGINE_TCGEN05 : B200LDM_ENGINE_SIMT;
}

extern "C" int b200ldm_conv3d_splits(const b200ldm_conv3d_params* p) {
  if (!p || conv3d_engine_of(p) != B200LDM_ENGINE_TCGEN05) return 1;
  return b200ldm::conv3d_tcgen05_splits(p);
}

extern "C" int b200ldm_conv3d_finalize(const float* slabs, int32_t splits, int64_t M, int32_t Cout,
                                       int64_t vox_per_sample, const float* bias, const float* temb,
                                       int32_t temb_stride, const void* residual, void* y, b200ldm_stream stream) {
  return b200ldm::conv3d_splitk_finalize(slabs, splits, M, Cout, vox_per_sample, bias, temb, temb_stride, residual, y,
                                         (cudaStream_t)stream);
}

extern "C" int b200ldm_conv3d_tune_pending(const b200ldm_conv3d_params* p) {
  if (!p || conv3d_engine_of(p) != B200LDM_ENGINE_TCGEN05) return 0;
  return b200ldm::conv3d_tcgen05_tune_pending(p);
}

extern "C" int b200ldm_conv3d_tune(const b200ldm_conv3d_params* p, b200ldm_stream stream) {
  using namespace b200ldm;
  B2_REQUIRE(p != nullptr, B200LDM_BAD_ARG, "null params");
  B2_REQUIRE(p->x && p->w && p->y, B200LDM_BAD_ARG, "null x / w / y");
  if (conv3d_engine_of(p) != B200LDM_ENGINE_TCGEN05) return B200LDM_OK;   // the validation engine has one kernel
  return conv3d_tcgen05_tune(p, (cudaStream_t)stream);
}

extern "C" int b200ldm_conv3d(const b200ldm_conv3d_params* p, b200ldm_stream stream) {
  using namespace b200ldm;
  B2_REQUIRE(p != nullptr, B200LDM_BAD_ARG, "null params");
  B2_REQUIRE(p->x && p->w && p->y, B200LDM_BAD_ARG, "null x / w / y");
  B2_REQUIRE(p->N > 0 && p->D > 0 && p->H > 0 && p->W > 0 && p->Do > 0 && p->Ho > 0 && p->Wo > 0,
             B200LDM_BAD_SHAPE, "non-positive extent");
  B2_REQUIRE(p->dtype == B200LDM_F32 || p->dtype == B200LDM_BF16, B200LDM_BAD_ARG, "bad dtype %d", p->dtype);
  B2_REQUIRE(p->ksize == 1 || p->ksize == 3, B200LDM_BAD_SHAPE, "ksize must be 1 or 3 (got %d)", p->ksize);
  B2_REQUIRE(p->stride == 1 || p->stride == 2, B200LDM_BAD_SHAPE, "stride must be 1 or 2 (got %d)", p->stride);
  // output extents must agree with (k, stride, pad_lo) and a pad_hi in {0,1}
  for (int i = 0; i < 3; ++i) {
    const int in = (i == 0 ? p->D : (i == 1 ? p->H : p->W)) * (p->upsample2x ? 2 : 1);
    const int out = i == 0 ? p->Do : (i == 1 ? p->Ho : p->Wo);
    const int last = (out - 1) * p->stride + (p->ksize - 1) - p->pad_lo;  // last input index touched
    B2_REQUIRE(last <= in && last >= in - 1 - (p->stride - 1), B200LDM_BAD_SHAPE,
               "output extent %d inconsistent with input %d (k=%d s=%d pad_lo=%d)", out, in, p->ksize, p->stride,
               p->pad_lo);
  }
  int engine = p->engine;
  if (engine == B200LDM_ENGINE_AUTO)
    engine = (p->dtype == B200LDM_BF16) ? B200LDM_ENGINE_TCGEN05 : B200LDM_ENGINE_SIMT;
  if (engine == B200LDM_ENGINE_TCGEN05) return conv3d_tcgen05(p, (cudaStream_t)stream);
  if (engine == B200LDM_ENGINE_SIMT) {
    B2_REQUIRE(!p->chunk_stats && !p->skip_x, B200LDM_BAD_ARG,
               "chunk_stats / fused skip conv are tensor-core path features");
    return conv3d_simt(p, (cudaStream_t)stream);
  }
  return fail(B200LDM_BAD_ARG, "unknown engine %d", p->engine);
}
