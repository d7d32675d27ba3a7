// elementwise.cu - the memory-bound odds and ends of the hot path: weight repack, layout
// bridges (NCDHW fp32 <-> NDHWC), nearest x2 upsample, timestep embedding, per-step row select,
// the fused scheduler update (DDPM / DDIM, Philox noise in-kernel) and AE sampling / add_noise.
#include "common.cuh"

namespace b200ldm {

constexpr int kEwThreads = 256;
static inline unsigned ew_blocks(long long n, int per_block = kEwThreads) {
  long long b = (n + per_block - 1) / per_block;
  const long long cap = 148LL * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------ weight repack
template <typename T>
__global__ void repack_weight_kernel(const float* __restrict__ src, T* __restrict__ dst, int Cout, int Cin, int taps,
                                     int Cout_pad, int Cin_pad, int cin_split, int cin_split_pad) {
  const long long total = (long long)taps * Cout_pad * Cin_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci_d = (int)(i % Cin_pad);
    long long t = i / Cin_pad;
    const int co = (int)(t % Cout_pad);
    const int tap = (int)(t / Cout_pad);
    int ci = -1;
    if (ci_d < cin_split) ci = ci_d;
    else if (ci_d >= cin_split_pad && ci_d - cin_split_pad + cin_split < Cin) ci = ci_d - cin_split_pad + cin_split;
    float v = 0.f;
    if (co < Cout && ci >= 0 && ci < Cin) v = src[((size_t)co * Cin + ci) * taps + tap];
    dst[i] = from_f32<T>(v);
  }
}

// nearest-x2-upsample + conv3 == 8 phase convs of 2x2x2 taps on the low-res tensor; taps that read the same
// low-res voxel are summed in fp32 before the single rounding to the storage dtype
template <typename T>
__global__ void repack_upconv_weight_kernel(const float* __restrict__ src, T* __restrict__ dst, int Cout, int Cin,
                                            int Cout_pad, int Cin_pad) {
  const long long total = 64LL * Cout_pad * Cin_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_pad);
    long long t = i / Cin_pad;
    const int co = (int)(t % Cout_pad);
    const int pt = (int)(t / Cout_pad);  // phase*8 + tap
    float v = 0.f;
    if (co < Cout && ci < Cin) {
      const int ph = pt >> 3, tp = pt & 7;
      const int a[3] = {ph >> 2, (ph >> 1) & 1, ph & 1};
      const int b[3] = {tp >> 2, (tp >> 1) & 1, tp & 1};
      int lo[3], hi[3];
      for (int d = 0; d < 3; ++d) {
        if (a[d] == 0) { lo[d] = b[d] == 0 ? 0 : 1; hi[d] = b[d] == 0 ? 0 : 2; }
        else           { lo[d] = b[d] == 0 ? 0 : 2; hi[d] = b[d] == 0 ? 1 : 2; }
      }
      const float* w = src + ((size_t)co * Cin + ci) * 27;
      for (int kd = lo[0]; kd <= hi[0]; ++kd)
        for (int kh = lo[1]; kh <= hi[1]; ++kh)
          for (int kw = lo[2]; kw <= hi[2]; ++kw) v += w[(kd * 3 + kh) * 3 + kw];
    }
    dst[i] = from_f32<T>(v);
  }
}

// ------------------------------------------------------------------ layout bridges
template <typename T>
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int N, int C, long long S,
                                      int C_dst, int c_offset) {
  const long long total = (long long)N * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / S, s = i - n * S;
    for (int c = 0; c < C; ++c) dst[(size_t)i * C_dst + c_offset + c] = from_f32<T>(src[((size_t)n * C + c) * S + s]);
  }
}
template <typename T>
__global__ void ndhwc_to_ncdhw_kernel(const T* __restrict__ src, float* __restrict__ dst, int N, int C, long long S,
                                      int C_src) {
  const long long total = (long long)N * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / S, s = i - n * S;
    for (int c = 0; c < C; ++c) dst[((size_t)n * C + c) * S + s] = to_f32<T>(src[(size_t)i * C_src + c]);
  }
}

// ------------------------------------------------------------------ nearest x2 upsample
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int D, int H, int W, int C) {
  const int vecs = C >> 3;
  const long long total = (long long)N * 8 * D * H * W * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vecs);
    long long t = i / vecs;
    const int wo = (int)(t % (2 * W)); t /= 2 * W;
    const int ho = (int)(t % (2 * H)); t /= 2 * H;
    const int dd = (int)(t % (2 * D));
    const int n = (int)(t / (2 * D));
    const size_t in = ((((size_t)n * D + (dd >> 1)) * H + (ho >> 1)) * W + (wo >> 1)) * C + (size_t)cv * 8;
    if (sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(y + (size_t)i * 8) = *reinterpret_cast<const uint4*>(x + in);
    } else {
      const uint4* s = reinterpret_cast<const uint4*>(x + in);
      uint4* o = reinterpret_cast<uint4*>(y + (size_t)i * 8);
      o[0] = s[0]; o[1] = s[1];
    }
  }
}

// ------------------------------------------------------------------ time embedding helpers
__global__ void silu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = v / (1.0f + expf(-v));
  }
}
// cos first, then sin; freqs = exp(-ln(max_period) * j / half).  The
// frequency and the trig are evaluated in fp64 and rounded once, so the fp32 validation mode
// is not limited by device-vs-host libm differences on arguments as large as 999.
__global__ void timestep_embedding_kernel(const float* __restrict__ ts, float* __restrict__ out, int n_t, int dim,
                                          float max_period) {
  const int half = dim / 2;
  const int total = n_t * dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i / dim, j = i - t * dim;
    float v = 0.f;
    if (j < 2 * half) {
      const int jj = j < half ? j : j - half;
      const float freq = (float)exp(-log((double)max_period) * (double)jj / (double)half);
      const float arg = ts[t] * freq;
      v = (float)(j < half ? cos((double)arg) : sin((double)arg));
    }
    out[i] = v;
  }
}
__global__ void select_row_kernel(const float* __restrict__ table, const int* __restrict__ counter,
                                  float* __restrict__ dst, int width) {
  const size_t row = (size_t)counter[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < width; i += gridDim.x * blockDim.x)
    dst[i] = table[row * width + i];
}
__global__ void counter_add_kernel(int* counter, int delta) {
  if (threadIdx.x == 0 && blockIdx.x == 0) counter[0] += delta;
}

// ------------------------------------------------------------------ Philox4x32-10 + Box-Muller
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 4 standard normals for the 4-element group `quad` of stream (seed, step)
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t step, uint64_t quad, float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10((uint32_t)quad, (uint32_t)(quad >> 32), step, 0x5eedu, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const float k = 2.3283064365386963e-10f;  // 2^-32
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float u1 = ((float)r[2 * p] + 0.5f) * k;
    const float u2 = ((float)r[2 * p + 1] + 0.5f) * k;
    const float rad = sqrtf(-2.0f * logf(fminf(fmaxf(u1, 1e-12f), 1.0f)));
    float s, c;
    sincosf(6.283185307179586f * u2, &s, &c);
    z[2 * p] = rad * c;
    z[2 * p + 1] = rad * s;
  }
}
__global__ void philox_normal_kernel(float* __restrict__ out, long long n, uint64_t seed, int step) {
  const long long quads = (n + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads;
       q += (long long)gridDim.x * blockDim.x) {
    float z[4];
    philox_normal4(seed, (uint32_t)step, (uint64_t)q, z);
    for (int e = 0; e < 4; ++e)
      if (q * 4 + e < n) out[q * 4 + e] = z[e];
  }
}

// ------------------------------------------------------------------ scheduler update
template <typename T>
__global__ void scheduler_step_kernel(const b200ldm_sched_params p) {
  const int step = p.counter ? p.counter[0] : 0;
  const uint64_t seed = p.seed_dev ? *p.seed_dev : p.seed;
  const float* cf = p.coef + (size_t)step * 8;
  const float a_x = cf[0], r_e = cf[1], c_x0 = cf[2], c_xt = cf[3], c_eps = cf[4], sigma = cf[5], lo = cf[6], hi = cf[7];
  const int qpv = p.C >> 2;  // 4-channel groups per voxel
  const long long quads = p.n_vox * qpv;
  const T* eps = reinterpret_cast<const T*>(p.eps);
  const T* vp = reinterpret_cast<const T*>(p.var_pred);
  const float var_a = vp ? p.var_coef[(size_t)step * 2] : 0.f, var_b = vp ? p.var_coef[(size_t)step * 2 + 1] : 0.f;
  T* xm = reinterpret_cast<T*>(p.x_model);
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads;
       q += (long long)gridDim.x * blockDim.x) {
    const long long vox = q / qpv;
    const int c = (int)(q - vox * qpv) * 4;
    const size_t xi = (size_t)vox * p.C + c;
    const float4 xt = *reinterpret_cast<const float4*>(p.x_t + xi);
    float e[4], x[4] = {xt.x, xt.y, xt.z, xt.w}, z[4] = {0.f, 0.f, 0.f, 0.f}, o[4], x0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = to_f32<T>(eps[(size_t)vox * p.C_eps + c + j]);
    if (sigma != 0.f) {
      if (p.noise) {
        const float4 nz = *reinterpret_cast<const float4*>(p.noise + (size_t)step * p.noise_step_stride + xi);
        z[0] = nz.x; z[1] = nz.y; z[2] = nz.z; z[3] = nz.w;
      } else {
        philox_normal4(seed, (uint32_t)step, (uint64_t)(q + p.philox_offset), z);
      }
    }
    float sg[4] = {sigma, sigma, sigma, sigma};
    if (vp && sigma != 0.f) {  // learned variance: per-element noise scale from the model's second output half
#pragma unroll
      for (int j = 0; j < 4; ++j) sg[j] = sqrtf(fmaf(var_a, to_f32<T>(vp[(size_t)vox * p.C_eps + c + j]), var_b));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float u = r_e * e[j];
      x0[j] = fminf(fmaxf(a_x != 0.f ? (x[j] + u) * a_x : u, lo), hi);
      o[j] = c_x0 * x0[j] + c_xt * x[j] + c_eps * e[j] + sg[j] * z[j];
    }
    *reinterpret_cast<float4*>(p.x_t + xi) = make_float4(o[0], o[1], o[2], o[3]);
    if (p.x0_out) *reinterpret_cast<float4*>(p.x0_out + xi) = make_float4(x0[0], x0[1], x0[2], x0[3]);
    if (xm) {
#pragma unroll
      for (int j = 0; j < 4; ++j) xm[(size_t)vox * p.C_model + c + j] = from_f32<T>(o[j]);
    }
  }
}

// ------------------------------------------------------------------ narrow-output 3x3x3 conv, second half
// A 3x3x3 conv with very few output channels (UNet `out.2`: 256 -> 8|16, decoder `conv_out`: 64 -> 1) is bound by
// re-reading its input 27 times, not by math.  It runs as a 1x1x1 GEMM Z[v][tap*cp + co] = sum_ci W[tap][co][ci] x[v][ci]
// on the tensor cores (input read ONCE) followed by this gather:  y[v][co] = bias[co] + sum_tap Z[v + shift(tap)][tap*cp + co]
// (taps that fall outside the volume contribute zero = the conv's zero padding).  One thread per voxel; every
// 16*cp-byte slice of Z is read exactly once.
// SCHED: the gathered tensor is the UNet's noise prediction and the scheduler update (b200ldm_scheduler_step, same
// coefficient row / noise stream / element indexing) is applied right here, on the fp32 sums, before anything is
// stored: "the scheduler update fused into the last UNet epilogue".  y may then be NULL (eps is not materialised).
template <typename T, int CP, bool SCHED>
__global__ void tap_gather_kernel(const T* __restrict__ z, const float* __restrict__ bias, T* __restrict__ y, int N, int D,
                                  int H, int W, int ldz, int cy, const b200ldm_sched_params sp) {
  const long long vox = (long long)N * D * H * W;
  int step = 0;
  uint64_t seed = 0;
  float cf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (SCHED) {
    step = sp.counter ? sp.counter[0] : 0;
    seed = sp.seed_dev ? *sp.seed_dev : sp.seed;
#pragma unroll
    for (int i = 0; i < 8; ++i) cf[i] = sp.coef[(size_t)step * 8 + i];
  }
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < vox; v += (long long)gridDim.x * blockDim.x) {
    long long t = v;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H); t /= H;
    const int d = (int)(t % D);
    float acc[CP];
#pragma unroll
    for (int j = 0; j < CP; ++j) acc[j] = bias ? __ldg(bias + j) : 0.f;
#pragma unroll 1
    for (int kd = 0; kd < 3; ++kd) {
      const int dd = d + kd - 1;
      if (dd < 0 || dd >= D) continue;
#pragma unroll 1
      for (int kh = 0; kh < 3; ++kh) {
        const int hh = h + kh - 1;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int ww = w + kw - 1;
          if (ww < 0 || ww >= W) continue;
          const long long src = v + ((long long)(kd - 1) * H + (kh - 1)) * W + (kw - 1);
          const T* zr = z + src * ldz + ((kd * 3 + kh) * 3 + kw) * CP;
          if (CP % 8 == 0) {
#pragma unroll
            for (int j = 0; j < CP; j += 8) {
              float f[8];
              Vec8<T>::load(zr + j, f);
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[j + e] += f[e];
            }
          } else {
#pragma unroll
            for (int j = 0; j < CP; ++j) acc[j] += to_f32<T>(zr[j]);
          }
        }
      }
    }
    if (SCHED) {
      // identical arithmetic to scheduler_step_kernel: quad q = voxel * (C / 4) + quad-in-voxel
      const float a_x = cf[0], r_e = cf[1], c_x0 = cf[2], c_xt = cf[3], c_eps = cf[4], sigma = cf[5], lo = cf[6], hi = cf[7];
      const int qpv = sp.C >> 2;
      T* xm = reinterpret_cast<T*>(sp.x_model);
#pragma unroll
      for (int qi = 0; qi < CP / 4; ++qi) {
        if (qi < qpv) {
          const size_t xi = (size_t)v * sp.C + qi * 4;
          const float4 xt = *reinterpret_cast<const float4*>(sp.x_t + xi);
          const float x[4] = {xt.x, xt.y, xt.z, xt.w};
          float zz[4] = {0.f, 0.f, 0.f, 0.f}, o[4], x0[4];
          if (sigma != 0.f) {
            if (sp.noise) {
              const float4 nz = *reinterpret_cast<const float4*>(sp.noise + (size_t)step * sp.noise_step_stride + xi);
              zz[0] = nz.x; zz[1] = nz.y; zz[2] = nz.z; zz[3] = nz.w;
            } else {
              philox_normal4(seed, (uint32_t)step, (uint64_t)(v * qpv + qi + sp.philox_offset), zz);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float e = acc[qi * 4 + j];
            const float u = r_e * e;
            x0[j] = fminf(fmaxf(a_x != 0.f ? (x[j] + u) * a_x : u, lo), hi);
            o[j] = c_x0 * x0[j] + c_xt * x[j] + c_eps * e + sigma * zz[j];
          }
          *reinterpret_cast<float4*>(sp.x_t + xi) = make_float4(o[0], o[1], o[2], o[3]);
          if (sp.x0_out) *reinterpret_cast<float4*>(sp.x0_out + xi) = make_float4(x0[0], x0[1], x0[2], x0[3]);
          if (xm) {
#pragma unroll
            for (int j = 0; j < 4; ++j) xm[(size_t)v * sp.C_model + qi * 4 + j] = from_f32<T>(o[j]);
          }
        }
      }
      if (y == nullptr) continue;
    }
    T* yr = y + v * cy;
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += 8) {   // compile-time indices keep acc[] in registers
      if (j0 < cy) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = (j0 + e < CP) ? acc[(j0 + e < CP) ? j0 + e : 0] : 0.f;
        Vec8<T>::store(yr + j0, f);
      }
    }
    for (int j0 = 16; j0 < cy; j0 += 8) {
      const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      Vec8<T>::store(yr + j0, zero);
    }
  }
}

template <typename T>
__global__ void ae_sampling_kernel(const T* __restrict__ mu, const T* __restrict__ logvar,
                                   const float* __restrict__ noise, float* __restrict__ z, float* __restrict__ sig,
                                   long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float lv = fminf(fmaxf(to_f32<T>(logvar[i]), -30.f), 20.f);
    const float s = expf(0.5f * lv);
    if (sig) sig[i] = s;
    z[i] = to_f32<T>(mu[i]) + s * (noise ? noise[i] : 0.f);
  }
}

__global__ void mul_add_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                               float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = fmaf(b[i], c[i], a[i]);
}

__global__ void add_noise_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                 const float* __restrict__ sa, const float* __restrict__ sb, float* __restrict__ out,
                                 int N, long long per) {
  const long long total = (long long)N * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per);
    out[i] = sa[n] * x0[i] + sb[n] * noise[i];
  }
}

// ------------------------------------------------------------------ narrow-INPUT 3x3x3 conv, first half
// A 3x3x3 conv with 1 or 2 input channels (AutoencoderKL `conv_in`: 1 -> 64 at full image resolution) would run on the
// tensor cores with its channels zero-padded to one 64-wide k-iteration per TAP: 27 k-iterations for 1 / 64 of the
// work.  Instead the 27 * cin neighbourhood values of a voxel are packed into ONE 64-channel row (channel t * cin + c =
// input channel c at tap t, zero outside the volume = the conv's padding, zero above 27 * cin) and the conv runs as a
// 1x1x1 GEMM over that tensor (weights repacked to match): one k-iteration.
template <typename T>
__global__ void pack_taps_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int D, int H, int W, int Cp, int cin) {
  // one thread per voxel: 27 * cin gathered values (the neighbours of a warp's voxels share lines), 128 contiguous bytes
  // out.  Measured at 160^3 (cin 1, bf16): 0.30 ms; 8 threads per voxel with one coalesced 16-byte store each 0.76 ms (8x
  // the index arithmetic and loads), rows staged through shared memory for coalesced stores 0.35 ms - the gather, not the
  // store pattern, is the cost.
  const long long vox = (long long)N * D * H * W;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < vox; v += (long long)gridDim.x * blockDim.x) {
    long long t = v;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H); t /= H;
    const int d = (int)(t % D);
    T row[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) row[i] = from_f32<T>(0.f);
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
      const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
      const int dd = d + kd - 1, hh = h + kh - 1, ww = w + kw - 1;
      if (dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const T* src = x + (v + ((long long)(kd - 1) * H + (kh - 1)) * W + (kw - 1)) * Cp;
        if (cin == 1) row[tap] = src[0];
        else { row[2 * tap] = src[0]; row[2 * tap + 1] = src[1]; }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(y + v * 64);
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int i = 0; i < (int)(64 * sizeof(T) / 16); ++i) dst[i] = r4[i];
  }
}

}  // namespace b200ldm

extern "C" {
using namespace b200ldm;

int b200ldm_repack_conv_weight(const float* src, void* dst, int32_t Cout, int32_t Cin, int32_t ksize,
                               int32_t Cout_pad, int32_t Cin_pad, int32_t cin_split, int32_t cin_split_pad,
                               int32_t dtype, b200ldm_stream stream) {
  B2_REQUIRE(src && dst && Cout > 0 && Cin > 0 && Cout_pad >= Cout && Cin_pad >= Cin, B200LDM_BAD_ARG,
             "repack_conv_weight: bad arguments");
  B2_REQUIRE(ksize == 1 || ksize == 3, B200LDM_BAD_SHAPE, "ksize must be 1 or 3");
  B2_REQUIRE(cin_split >= 0 && cin_split <= Cin && cin_split_pad >= cin_split &&
                 cin_split_pad + (Cin - cin_split) <= Cin_pad,
             B200LDM_BAD_SHAPE, "repack_conv_weight: bad channel split");
  const int taps = ksize * ksize * ksize;
  const long long total = (long long)taps * Cout_pad * Cin_pad;
  if (dtype == B200LDM_BF16)
    repack_weight_kernel<__nv_bfloat16><<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>(
        src, (__nv_bfloat16*)dst, Cout, Cin, taps, Cout_pad, Cin_pad, cin_split, cin_split_pad);
  else if (dtype == B200LDM_F32)
    repack_weight_kernel<float><<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>(
        src, (float*)dst, Cout, Cin, taps, Cout_pad, Cin_pad, cin_split, cin_split_pad);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_repack_upconv_weight(const float* src, void* dst, int32_t Cout, int32_t Cin, int32_t Cout_pad,
                                 int32_t Cin_pad, int32_t dtype, b200ldm_stream stream) {
  B2_REQUIRE(src && dst && Cout > 0 && Cin > 0 && Cout_pad >= Cout && Cin_pad >= Cin, B200LDM_BAD_ARG,
             "repack_upconv_weight: bad arguments");
  const long long total = 64LL * Cout_pad * Cin_pad;
  if (dtype == B200LDM_BF16)
    repack_upconv_weight_kernel<__nv_bfloat16><<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>(
        src, (__nv_bfloat16*)dst, Cout, Cin, Cout_pad, Cin_pad);
  else if (dtype == B200LDM_F32)
    repack_upconv_weight_kernel<float><<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>(
        src, (float*)dst, Cout, Cin, Cout_pad, Cin_pad);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_ncdhw_to_ndhwc(const float* src, void* dst, int32_t N, int32_t C, int64_t S, int32_t C_dst,
                           int32_t c_offset, int32_t dtype, b200ldm_stream stream) {
  B2_REQUIRE(src && dst && N > 0 && C > 0 && S > 0 && c_offset >= 0 && c_offset + C <= C_dst, B200LDM_BAD_ARG,
             "ncdhw_to_ndhwc: bad arguments");
  const long long total = (long long)N * S;
  if (dtype == B200LDM_BF16)
    ncdhw_to_ndhwc_kernel<__nv_bfloat16><<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>(
        src, (__nv_bfloat16*)dst, N, C, S, C_dst, c_offset);
  else if (dtype == B200LDM_F32)
    ncdhw_to_ndhwc_kernel<float><<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>(src, (float*)dst, N, C, S,
                                                                                            C_dst, c_offset);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_ndhwc_to_ncdhw(const void* src, float* dst, int32_t N, int32_t C, int64_t S, int32_t C_src,
                           int32_t dtype, b200ldm_stream stream) {
  B2_REQUIRE(src && dst && N > 0 && C > 0 && S > 0 && C <= C_src, B200LDM_BAD_ARG, "ndhwc_to_ncdhw: bad arguments");
  const long long total = (long long)N * S;
  if (dtype == B200LDM_BF16)
    ndhwc_to_ncdhw_kernel<__nv_bfloat16><<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)src, dst, N, C, S, C_src);
  else if (dtype == B200LDM_F32)
    ndhwc_to_ncdhw_kernel<float><<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>((const float*)src, dst, N,
                                                                                            C, S, C_src);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_upsample_nearest2x(const void* x, void* y, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C,
                               int32_t dtype, b200ldm_stream stream) {
  B2_REQUIRE(x && y && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, B200LDM_BAD_SHAPE,
             "upsample_nearest2x: bad arguments (C must be a multiple of 8)");
  B2_REQUIRE(aligned16(x) && aligned16(y), B200LDM_BAD_ALIGN, "upsample_nearest2x: pointers must be 16-byte aligned");
  const long long total = (long long)N * 8 * D * H * W * (C / 8);
  if (dtype == B200LDM_BF16)
    upsample2x_kernel<__nv_bfloat16><<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, D, H, W, C);
  else if (dtype == B200LDM_F32)
    upsample2x_kernel<float><<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, N,
                                                                                        D, H, W, C);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_pack_taps(const void* x, void* y, int32_t N, int32_t D, int32_t H, int32_t W, int32_t Cp, int32_t cin,
                      int32_t dtype, b200ldm_stream stream) {
  B2_REQUIRE(x && y && N > 0 && D > 0 && H > 0 && W > 0 && (cin == 1 || cin == 2) && Cp >= cin, B200LDM_BAD_SHAPE,
             "pack_taps: 1 or 2 input channels (27 * cin <= 64), Cp >= cin");
  B2_REQUIRE(aligned16(y), B200LDM_BAD_ALIGN, "pack_taps: output must be 16-byte aligned");
  const long long vox = (long long)N * D * H * W;
  if (dtype == B200LDM_BF16)
    pack_taps_kernel<__nv_bfloat16><<<ew_blocks(vox), kEwThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, D, H, W, Cp, cin);
  else if (dtype == B200LDM_F32)
    pack_taps_kernel<float><<<ew_blocks(vox), kEwThreads, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, N, D, H, W,
                                                                                   Cp, cin);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_tap_gather(const void* z, const float* bias, void* y, int32_t N, int32_t D, int32_t H, int32_t W, int32_t ldz,
                       int32_t cp, int32_t cy, int32_t dtype, const b200ldm_sched_params* sched, b200ldm_stream stream) {
  B2_REQUIRE(z && (y || sched) && N > 0 && D > 0 && H > 0 && W > 0, B200LDM_BAD_ARG, "tap_gather: bad arguments");
  B2_REQUIRE(!sched || (!sched->var_pred && !sched->var_coef), B200LDM_BAD_ARG,
             "tap_gather: the fused scheduler update does not take a learned variance (use b200ldm_scheduler_step)");
  B2_REQUIRE((cp == 1 || cp == 2 || cp == 4 || cp == 8 || cp == 16) && cy >= cp && cy % 8 == 0 && ldz >= 27 * cp &&
                 ldz % 8 == 0,
             B200LDM_BAD_SHAPE, "tap_gather: channels per tap must be 1/2/4/8/16, cy a multiple of 8 >= cp, ldz >= 27*cp");
  B2_REQUIRE(aligned16(z) && aligned16(y), B200LDM_BAD_ALIGN, "tap_gather: pointers must be 16-byte aligned");
  B2_REQUIRE(dtype == B200LDM_BF16 || dtype == B200LDM_F32, B200LDM_BAD_ARG, "bad dtype %d", dtype);
  b200ldm_sched_params sp;
  memset(&sp, 0, sizeof(sp));
  if (sched) {
    sp = *sched;
    B2_REQUIRE(sp.x_t && sp.coef && sp.C > 0 && sp.C % 4 == 0 && sp.C <= cp && sp.n_vox == (int64_t)N * D * H * W &&
                   (!sp.x_model || sp.C_model >= sp.C) && sp.dtype == dtype,
               B200LDM_BAD_ARG, "tap_gather: fused scheduler update needs x_t, coef, C %% 4 == 0, C <= cp, n_vox = N*D*H*W");
    B2_REQUIRE(aligned16(sp.x_t) && aligned16(sp.noise) && aligned16(sp.x0_out) && sp.noise_step_stride % 4 == 0,
               B200LDM_BAD_ALIGN, "tap_gather: fp32 scheduler buffers must be 16-byte aligned");
  }
  const long long vox = (long long)N * D * H * W;
  const cudaStream_t st = (cudaStream_t)stream;
#define B2_TG(T, CP, S)                                                                                              \
  tap_gather_kernel<T, CP, S><<<ew_blocks(vox), kEwThreads, 0, st>>>((const T*)z, bias, (T*)y, N, D, H, W, ldz, cy, sp)
#define B2_TG_T(T)                                                       \
  if (sched) {                                                           \
    switch (cp) {                                                        \
      case 4: B2_TG(T, 4, true); break;                                  \
      case 8: B2_TG(T, 8, true); break;                                  \
      case 16: B2_TG(T, 16, true); break;                                \
      default: return fail(B200LDM_BAD_SHAPE, "tap_gather: fused scheduler update needs cp >= 4"); \
    }                                                                    \
  } else {                                                               \
    switch (cp) {                                                        \
      case 1: B2_TG(T, 1, false); break;                                 \
      case 2: B2_TG(T, 2, false); break;                                 \
      case 4: B2_TG(T, 4, false); break;                                 \
      case 8: B2_TG(T, 8, false); break;                                 \
      default: B2_TG(T, 16, false); break;                               \
    }                                                                    \
  }
  if (dtype == B200LDM_BF16) { B2_TG_T(__nv_bfloat16) } else { B2_TG_T(float) }
#undef B2_TG_T
#undef B2_TG
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_silu_f32(const float* x, float* y, int64_t n, b200ldm_stream stream) {
  B2_REQUIRE(x && y && n > 0, B200LDM_BAD_ARG, "silu_f32: bad arguments");
  silu_f32_kernel<<<ew_blocks(n), kEwThreads, 0, (cudaStream_t)stream>>>(x, y, n);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_timestep_embedding(const float* timesteps, float* out, int32_t n_t, int32_t dim, float max_period,
                               b200ldm_stream stream) {
  B2_REQUIRE(timesteps && out && n_t > 0 && dim > 1, B200LDM_BAD_ARG, "timestep_embedding: bad arguments");
  timestep_embedding_kernel<<<ew_blocks((long long)n_t * dim), kEwThreads, 0, (cudaStream_t)stream>>>(
      timesteps, out, n_t, dim, max_period);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_select_row(const float* table, const int32_t* counter, float* dst, int32_t width,
                       b200ldm_stream stream) {
  B2_REQUIRE(table && counter && dst && width > 0, B200LDM_BAD_ARG, "select_row: bad arguments");
  select_row_kernel<<<ew_blocks(width), kEwThreads, 0, (cudaStream_t)stream>>>(table, counter, dst, width);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_counter_add(int32_t* counter, int32_t delta, b200ldm_stream stream) {
  B2_REQUIRE(counter, B200LDM_BAD_ARG, "counter_add: null counter");
  counter_add_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counter, delta);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_scheduler_step(const b200ldm_sched_params* p, b200ldm_stream stream) {
  B2_REQUIRE(p && p->eps && p->x_t && p->coef, B200LDM_BAD_ARG, "scheduler_step: null eps / x_t / coef");
  B2_REQUIRE(p->n_vox > 0 && p->C > 0 && p->C % 4 == 0 && p->C_eps >= p->C, B200LDM_BAD_SHAPE,
             "scheduler_step: C must be a positive multiple of 4 and <= C_eps");
  B2_REQUIRE(!p->x_model || p->C_model >= p->C, B200LDM_BAD_SHAPE, "scheduler_step: C_model < C");
  B2_REQUIRE((p->var_pred == nullptr) == (p->var_coef == nullptr), B200LDM_BAD_ARG,
             "scheduler_step: var_pred and var_coef go together (learned variance)");
  B2_REQUIRE(aligned16(p->x_t) && aligned16(p->noise) && aligned16(p->x0_out) && p->noise_step_stride % 4 == 0,
             B200LDM_BAD_ALIGN, "scheduler_step: fp32 buffers must be 16-byte aligned");
  const long long quads = p->n_vox * (p->C / 4);
  if (p->dtype == B200LDM_BF16)
    scheduler_step_kernel<__nv_bfloat16><<<ew_blocks(quads), kEwThreads, 0, (cudaStream_t)stream>>>(*p);
  else if (p->dtype == B200LDM_F32)
    scheduler_step_kernel<float><<<ew_blocks(quads), kEwThreads, 0, (cudaStream_t)stream>>>(*p);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", p->dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_philox_normal(float* out, int64_t n, uint64_t seed, int32_t step, b200ldm_stream stream) {
  B2_REQUIRE(out && n > 0, B200LDM_BAD_ARG, "philox_normal: bad arguments");
  philox_normal_kernel<<<ew_blocks((n + 3) / 4), kEwThreads, 0, (cudaStream_t)stream>>>(out, n, seed, step);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_ae_sampling(const void* mu, const void* logvar, const float* noise, float* z, float* sigma_out, int64_t n,
                        int32_t dtype, b200ldm_stream stream) {
  B2_REQUIRE(mu && logvar && z && n > 0, B200LDM_BAD_ARG, "ae_sampling: bad arguments");
  if (dtype == B200LDM_BF16)
    ae_sampling_kernel<__nv_bfloat16><<<ew_blocks(n), kEwThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)mu, (const __nv_bfloat16*)logvar, noise, z, sigma_out, n);
  else if (dtype == B200LDM_F32)
    ae_sampling_kernel<float><<<ew_blocks(n), kEwThreads, 0, (cudaStream_t)stream>>>((const float*)mu,
                                                                                     (const float*)logvar, noise, z,
                                                                                     sigma_out, n);
  else
    return fail(B200LDM_BAD_ARG, "bad dtype %d", dtype);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_mul_add_f32(const float* a, const float* b, const float* c, float* out, int64_t n,
                        b200ldm_stream stream) {
  B2_REQUIRE(a && b && c && out && n > 0, B200LDM_BAD_ARG, "mul_add_f32: bad arguments");
  mul_add_kernel<<<ew_blocks(n), kEwThreads, 0, (cudaStream_t)stream>>>(a, b, c, out, n);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int b200ldm_add_noise(const float* x0, const float* noise, const float* sqrt_acp, const float* sqrt_1m_acp,
                      float* out, int32_t N, int64_t per_sample, b200ldm_stream stream) {
  B2_REQUIRE(x0 && noise && sqrt_acp && sqrt_1m_acp && out && N > 0 && per_sample > 0, B200LDM_BAD_ARG,
             "add_noise: bad arguments");
  add_noise_kernel<<<ew_blocks((long long)N * per_sample), kEwThreads, 0, (cudaStream_t)stream>>>(
      x0, noise, sqrt_acp, sqrt_1m_acp, out, N, per_sample);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

}  // extern "C"
