// conv_halo.cu - persistent tcgen05 implicit-GEMM Conv3d (3x3x3, stride 1, pad 1) for NARROW outputs (Cout tile 64 or
// 128: the AutoencoderKL's 64/128-channel convs at full image resolution).  Two things bound conv_tcp.cu there:
//   * bytes into the SM: with a 64-column tile every k-iteration stages 2 x 16 KB of activations for ~400 tensor
//     cycles, above the ~72 B/clk an SM ingests from L2;
//   * the single MMA-issuing thread: a barrier wait + tcgen05.commit pair costs it ~250 cycles and building a
//     descriptor with shifts and masks ~85 cycles per MMA - more than an N <= 128 instruction takes to execute
//     (scripts/conv_dry.py: with loads and epilogue switched off the kernel was no faster).
// Here one TMA load brings a W-HALO box - the 8-voxel lines of the tile plus one voxel on either side, 10 x 128 B per
// line - and the three kw taps read it through three shifted views: the UMMA descriptor starts kw rows (kw * 128 B) into
// the box and steps 1280 B between 8-row groups (one line each).  That is legal because both TMA and UMMA derive the
// SWIZZLE_128B XOR from absolute shared-memory address bits, so a row is found where the TMA put it whatever the
// view's origin, and it costs nothing (scripts/micro/umma_rate.cu: 48 cycles per M128 N64 K16 instruction either
// way).  A stage = the MT halo boxes of a (kd, kh, chunk) group + the weight tiles of its three taps under ONE barrier
// pair, i.e. 24 MMAs per wait / commit, and descriptors step by one 32-bit add (ptx::umma_bf16_lohi).
//
// MEASURED (profiles/r01e_conv_halo.log): 64 -> 64 at 8 x 80^3 0.82 ms = 1108 TFLOP/s against 1.25 ms for conv_tcp.cu -
// the N = 64 instruction rate (48 cycles instead of the 32-cycle floor, 66.6 %) is now the bound; 128 -> 128 at
// 8 x 40^3 0.30 ms = 1518 TFLOP/s against 0.38 ms.  (Its first version - two rings, a barrier pair per tap, descriptors
// rebuilt per MMA - was 12 % SLOWER than conv_tcp.cu although it moved 2.4x fewer bytes: the issuer, not L2, set the
// pace.)
//
// Everything else follows conv_tcp.cu: one CTA per SM walks the tile list, two TMEM accumulator sets (the epilogue of
// tile i overlaps the main loop of tile i+1), the same epilogue math (bias / temb / residual / GroupNorm chunk sums,
// staged through swizzled smem and stored by TMA), the fused 1x1x1 skip conv as extra K iterations.
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace b200ldm {

constexpr int kHaloTw = 8;                              // voxels per line of the output tile
constexpr int kHaloLines = 16;                          // td * th lines = 128 rows
constexpr int kHaloRowBytes = 128;                      // 64 bf16 channels
constexpr int kHaloLineBytes = (kHaloTw + 2) * kHaloRowBytes;   // 1280: one line with its two halo voxels
constexpr int kHaloABytes = kHaloLines * kHaloLineBytes;        // 20480 per sub-tile and stage

template <int BN, int MT>
struct ConvHaloCfg {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kAStage = MT * kHaloABytes;
  // ONE ring: a stage = the MT halo boxes of a (kd, kh, chunk) group + the weight tiles of its three kw taps, filled
  // under one barrier and released by one commit.  The single MMA-issuing thread pays ~250 cycles per barrier wait +
  // commit pair; with a pair per tap (8 instructions of 48 cycles for BN = 64) that overhead, not the tensor pipe,
  // set the pace.
  static constexpr int kStageBytes = kAStage + 3 * kBBytes;
  static constexpr int kOutBytes = MT * 16384;       // per sub-tile ONE [128 rows x 64 ch] bf16 staging buffer
  static constexpr int kBarBytes = 512;
  static constexpr int kFlagBytes = MT * 768;
  static constexpr int kTail = 1024 /*align slack*/ + kBarBytes + kFlagBytes;
  static constexpr int kBudget = 232448 - kOutBytes - kTail;
  static constexpr int kStages = kBudget / kStageBytes > 4 ? 4 : kBudget / kStageBytes;
  static constexpr int kAccCols = MT * BN;
  static constexpr int kTmemCols = 2 * kAccCols;
  static constexpr int kThreads = 64 + MT * 128;
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + kTail;
  static_assert(kStages >= 2, "pipeline depth");
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols >= 32, "TMEM columns");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  static_assert(kStageBytes % 1024 == 0 && kHaloABytes % 1024 == 0 && kBBytes % 1024 == 0, "1024-byte swizzle alignment");
  static_assert(2 * kStages * 8 + 32 + MT * 8 + 8 <= kBarBytes, "barrier area");
};

struct HaloTile {
  int n0, d0, h0, w0;
};
__device__ __forceinline__ HaloTile halo_box(const ConvTcArgs& a, int t) {
  HaloTile c;   // a box index past the end decodes to n0 >= N: TMA zero-fills the rows and the epilogue masks them
  c.w0 = (t % a.nb_w) * a.tw; t /= a.nb_w;
  c.h0 = (t % a.nb_h) * a.th; t /= a.nb_h;
  c.d0 = (t % a.nb_d) * a.td; t /= a.nb_d;
  c.n0 = t;
  return c;
}

template <int BN, int MT>
__global__ void __launch_bounds__(ConvHaloCfg<BN, MT>::kThreads, 1)
conv3d_halo_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs a) {
  using Cfg = ConvHaloCfg<BN, MT>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t a_base = smem_base;                              // stage s at + s * kStageBytes: MT halo boxes ...
  const uint32_t b_base = a_base + Cfg::kAStage;                  // ... then the three weight tiles
  const uint32_t out_base = smem_base + STAGES * Cfg::kStageBytes;
  const uint32_t bar_base = out_base + Cfg::kOutBytes;
  const uint32_t full_bar = bar_base;                    // [STAGES]
  const uint32_t empty_bar = full_bar + 8 * STAGES;      // [STAGES]
  const uint32_t tfull_bar = empty_bar + 8 * STAGES;     // [2] accumulator set complete
  const uint32_t tempty_bar = tfull_bar + 16;            // [2] accumulator set drained (MT*4 warp arrivals)
  const uint32_t res_bar = tempty_bar + 16;              // [MT] residual tile landed in the staging buffer
  const uint32_t tmem_slot = res_bar + 8 * MT;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));
  uint8_t* flag_gen = smem_gen + (bar_base - smem_base) + Cfg::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int work_x = a.work_x, work_y = a.work_y;
  const int total = work_x * work_y;
  const int nchunks = a.nchunk1 + a.nchunk2;
  const int groups = 9 * nchunks;                 // (kd, kh, chunk) groups of three kw taps
  const int nskip = a.nsk1 + a.nsk2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&maps.w);
    ptx::prefetch_tmap(&maps.a[0]);
    ptx::prefetch_tmap(&maps.y[0]);
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(full_bar + 8 * s, 1); ptx::mbar_init(empty_bar + 8 * s, 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull_bar + 8 * i, 1);
      ptx::mbar_init(tempty_bar + 8 * i, MT * 4);
    }
    for (int i = 0; i < MT; ++i) ptx::mbar_init(res_bar + 8 * i, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t itg = 0;
      for (int wk = blockIdx.x; wk < total; wk += gridDim.x) {
        const int x = wk % work_x, co0 = (wk / work_x) * BN;
        HaloTile c[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) c[mt] = halo_box(a, x * MT + mt);
        for (int g = 0; g < groups + nskip; ++g, ++itg) {
          const int s = (int)(itg % STAGES);
          ptx::mbar_wait(empty_bar + 8 * s, ((itg / STAGES) & 1u) ^ 1u, 0);
          const uint32_t full = full_bar + 8 * s;
          const uint32_t sa = a_base + s * Cfg::kStageBytes, sb = b_base + s * Cfg::kStageBytes;
          if (g >= groups) {   // fused 1x1x1 skip conv: plain 128-row boxes, one tap
            const int sc = g - groups;
            const bool second = sc >= a.nsk1;
            const CUtensorMap* m = second ? &maps.s[1] : &maps.s[0];
            ptx::mbar_expect_tx(full, (uint32_t)MT * 16384u + (uint32_t)Cfg::kBBytes);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              ptx::tma_load_5d(sa + mt * kHaloABytes, m, full, (second ? sc - a.nsk1 : sc) * 64, c[mt].w0, c[mt].h0,
                               c[mt].d0, c[mt].n0);
            ptx::tma_load_3d(sb, &maps.sw, full, sc * 64, co0, 0);
            continue;
          }
          const int kdh = g / nchunks, chunk = g - kdh * nchunks;
          const int kd = kdh / 3, kh = kdh - kd * 3;
          const bool second = chunk >= a.nchunk1;
          const CUtensorMap* m = second ? &maps.a[1] : &maps.a[0];
          const int c0 = (second ? chunk - a.nchunk1 : chunk) * 64;
          ptx::mbar_expect_tx(full, (uint32_t)MT * kHaloABytes + 3u * (uint32_t)Cfg::kBBytes);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            ptx::tma_load_5d(sa + mt * kHaloABytes, m, full, c0, c[mt].w0 - 1, c[mt].h0 + kh - 1, c[mt].d0 + kd - 1,
                             c[mt].n0);
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            ptx::tma_load_3d(sb + kw * Cfg::kBBytes, &maps.w, full, chunk * 64, co0, kdh * 3 + kw);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BN);
      // descriptor low words step by one add (ptx::umma_bf16_lohi); the halo views use SBO = one line = 1280 B
      constexpr uint32_t ahi = ptx::umma_desc_hi_sw128(kHaloLineBytes), bhi = ptx::umma_desc_hi_sw128();
      const uint32_t a_lo0 = ptx::umma_desc_lo_sw128(a_base), b_lo0 = ptx::umma_desc_lo_sw128(b_base);
      uint32_t itg = 0, tile = 0;
      for (int wk = blockIdx.x; wk < total; wk += gridDim.x, ++tile) {
        const uint32_t acc = tile & 1u, use = tile >> 1;
        ptx::mbar_wait(tempty_bar + 8 * acc, (use & 1u) ^ 1u, 5);   // the epilogue drained this accumulator set
        ptx::tc_fence_after();
        const uint32_t td = tmem_base + acc * Cfg::kAccCols;
        uint32_t s = itg % STAGES, ph = (itg / STAGES) & 1u;
        for (int g = 0; g < groups + nskip; ++g, ++itg) {
          ptx::mbar_wait(full_bar + 8 * s, ph, 1);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (uint32_t)(Cfg::kStageBytes >> 4);
          const uint32_t b_lo = b_lo0 + s * (uint32_t)(Cfg::kStageBytes >> 4);
          if (g < groups) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  ptx::umma_bf16_lohi(td + mt * BN,
                                      a_lo + (uint32_t)((mt * kHaloABytes + kw * kHaloRowBytes + k * 32) >> 4),
                                      b_lo + (uint32_t)((kw * Cfg::kBBytes + k * 32) >> 4), ahi, bhi, idesc,
                                      (g > 0 || kw > 0 || k > 0) ? 1u : 0u);
              }
            }
          } else {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_bf16_lohi(td + mt * BN, a_lo + (uint32_t)((mt * kHaloABytes + k * 32) >> 4),
                                    b_lo + (uint32_t)(k * 2), bhi, bhi, idesc, 1u);
            }
          }
          ptx::umma_commit(empty_bar + 8 * s);   // stage free once these MMAs retire
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        ptx::umma_commit(tfull_bar + 8 * acc);   // accumulator set complete: wake the epilogue warps
      }
    }
  } else {
    // =========================== epilogue (4 warps per sub-tile) ===========================
    constexpr int G = BN / 64;
    const int mt = (warp - 2) >> 2;  // which sub-tile / accumulator of the set
    const int q = warp & 3;          // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;   // accumulator row == voxel slot in the box
    const bool issuer = (((warp - 2) & 3) == 0) && lane == 0;
    const uint32_t stg = out_base + (uint32_t)mt * 16384;
    uint8_t* stg_gen = smem_gen + (out_base - smem_base) + (size_t)mt * 16384;
    uint8_t* rowp = stg_gen + (size_t)row * 128;
    const uint32_t my_res = res_bar + 8 * mt;
    uint8_t* vrow = flag_gen + mt * 768;                        // [128] row-valid flags
    float* csum = reinterpret_cast<float*>(vrow + 128);         // [G*8][2]
    const bool do_stats = a.chunk_stats != nullptr;
    const int ns = 8 / a.cw;   // sub-sums per 8-channel chunk (chunk width 8, 4 or 2)
    uint32_t tile = 0, res_ph = 0;
    int t = row;
    const int rw = t % a.tw; t /= a.tw;
    const int rh = t % a.th; t /= a.th;
    const int rd = t;
    double* s_dst = nullptr;   // GroupNorm chunk sums: cross-tile fp64 accumulator of this thread's entry
    double s_acc = 0.0;
    for (int wk = blockIdx.x; wk < total; wk += gridDim.x, ++tile) {
      const int x = wk % work_x, co0 = (wk / work_x) * BN;
      const HaloTile c = halo_box(a, x * MT + mt);
      const int n = c.n0, d = c.d0 + rd, h = c.h0 + rh, w = c.w0 + rw;
      const bool valid = n < a.N && d < a.Do && h < a.Ho && w < a.Wo;
      const float* temb_row = a.temb ? a.temb + (size_t)(a.temb_stride ? n : 0) * a.temb_stride : nullptr;
      const uint32_t acc = tile & 1u, use = tile >> 1;
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::kAccCols + (uint32_t)(mt * BN);
      const int gmax = min(G, (a.Cout - co0 + 63) / 64);   // 64-channel groups present in this Cout tile
      if (do_stats) {
        vrow[row] = valid ? 1 : 0;
        if (row < G * 16 * ns) csum[row] = 0.f;
      }
      ptx::mbar_wait_long(tfull_bar + 8 * acc, use & 1u, 2);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < gmax; ++g) {
        uint32_t rr[4][16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) ptx::tmem_ld16(tacc + (uint32_t)(g * 64 + cc * 16), rr[cc]);
        if (issuer) {
          ptx::tma_store_wait_read<0>();   // the previous store has read the staging buffer out
          if (a.residual) {
            ptx::mbar_expect_tx(my_res, 16384u);
            ptx::tma_load_5d(stg, &maps.r, my_res, co0 + g * 64, c.w0, c.h0, c.d0, c.n0);
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");   // the buffer is free, flags are visible
        ptx::tmem_ld_wait();
        if (g == gmax - 1) {
          // every column of this accumulator is in registers: hand it back to the MMA issuer
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(tempty_bar + 8 * acc);
        }
        if (a.residual) {
          ptx::mbar_wait(my_res, res_ph, 3);
          res_ph ^= 1u;
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int co = co0 + g * 64 + cc * 16;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[cc][j]);
          if (a.bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + co + j));
              v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
            }
          }
          if (temb_row && valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 tt = __ldg(reinterpret_cast<const float4*>(temb_row + co + j));
              v[j] += tt.x; v[j + 1] += tt.y; v[j + 2] += tt.z; v[j + 3] += tt.w;
            }
          }
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            const int chunk = cc * 2 + hlf;  // 16-byte chunk inside the 128-byte row
            uint4* slot = reinterpret_cast<uint4*>(rowp + ((chunk ^ (row & 7)) << 4));
            float f[8];
            if (a.residual) {
              const uint4 u = *slot;
              const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t2 = __bfloat1622float2(hp[e]);
                f[2 * e] = v[hlf * 8 + 2 * e] + t2.x;
                f[2 * e + 1] = v[hlf * 8 + 2 * e + 1] + t2.y;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = v[hlf * 8 + e];
            }
            uint4 o;
            __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) op[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
            *slot = o;
          }
        }
        ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA store
        asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");
        if (issuer) {
          ptx::tma_store_5d(&maps.y[0], stg, co0 + g * 64, c.w0, c.h0, c.d0, c.n0);
          ptx::tma_store_commit();
        }
        if (do_stats) {
          // thread (slice = row>>3, chunk = row&7) sums 8 rows x 8 channels of the staged tile
          const int cch = row & 7, slice = row >> 3;
          float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};   // per channel PAIR of the 8-channel chunk
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = slice * 8 + i;
            if (vrow[r]) {
              const uint4 u = *reinterpret_cast<const uint4*>(stg_gen + (size_t)r * 128 + ((cch ^ (r & 7)) << 4));
              const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t2 = __bfloat1622float2(hp[e]);
                s1[e] += t2.x + t2.y;
                s2[e] = fmaf(t2.x, t2.x, fmaf(t2.y, t2.y, s2[e]));
              }
            }
          }
          // fold the pair sums to the chunk width: ns = 8 / cw sub-sums per 8-channel chunk (1, 2 or 4)
          if (ns <= 2) { s1[0] += s1[1]; s2[0] += s2[1]; s1[1] = s1[2] + s1[3]; s2[1] = s2[2] + s2[3]; }
          if (ns == 1) { s1[0] += s1[1]; s2[0] += s2[1]; }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < ns) {   // warp-uniform
              float a1 = s1[i], a2 = s2[i];
              a1 += __shfl_xor_sync(0xffffffffu, a1, 8); a2 += __shfl_xor_sync(0xffffffffu, a2, 8);
              a1 += __shfl_xor_sync(0xffffffffu, a1, 16); a2 += __shfl_xor_sync(0xffffffffu, a2, 16);
              if (lane < 8) {
                atomicAdd(&csum[((g * 8 + lane) * ns + i) * 2], a1);
                atomicAdd(&csum[((g * 8 + lane) * ns + i) * 2 + 1], a2);
              }
            }
          }
          // the statistics read the staged tile: with ONE staging buffer the next group's writes must wait for them
          asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");
        }
      }
      if (do_stats) {
        if (row < gmax * 16 * ns && c.n0 < a.N) {
          // entry `row` of this tile's sums (chunk row >> 1 of the Cout tile, 0 = sum / 1 = sumsq) joins this thread's
          // fp64 accumulator; the global atomic is issued only when its target changes (another sample / Cout tile) or
          // after the last tile: a CTA of a batch-1 AutoencoderKL conv walks ~200 tiles of ONE sample, and 2 M atomics
          // on the same 64 addresses (four L2 lines) cost more than the statistics themselves
          double* dst = a.chunk_stats + ((size_t)c.n0 * (a.Cout / a.cw) + co0 / a.cw + (row >> 1)) * 2 + (row & 1);
          if (dst != s_dst) {
            if (s_dst) atomicAdd(s_dst, s_acc);
            s_dst = dst;
            s_acc = 0.0;
          }
          s_acc += (double)csum[row];
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");   // csum is re-zeroed at the top of the next tile
      }
    }
    if (s_dst) atomicAdd(s_dst, s_acc);
    if (issuer) ptx::tma_store_wait_all();   // smem must outlive the bulk reads
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ host
// Eligible problems: 3x3x3 / stride 1 / pad 1, bf16 TMA epilogue, Cout tile 64 or 128, no fused upsample, no split-K,
// W of at least one line.  The tile is 1 sample x td x th x 8 voxels with td * th = 16.
bool conv_halo_supported(int bn, const ConvTcArgs& a) {
  return (bn == 128 || bn == 64) && a.tma_epilogue && a.splits == 1 && !a.out_f32 && !a.gn_stats && !a.up &&
         a.ksize == 3 && a.stride == 1 && a.pad_lo == 1 && a.tw == kHaloTw && a.tn == 1 && a.td * a.th == kHaloLines;
}

void choose_halo_tile(int Do, int Ho, int* td, int* th) {
  long best = -1;
  *td = 1; *th = kHaloLines;
  for (int d = 1; d <= kHaloLines; d *= 2) {
    const int h = kHaloLines / d;
    const long tiles = (long)ceil_div(Do, d) * ceil_div(Ho, h);
    if (best < 0 || tiles < best || (tiles == best && h > *th)) { best = tiles; *td = d; *th = h; }
  }
}

template <int BN, int MT>
static int launch_halo(const ConvTcMaps& maps, ConvTcArgs a, int m_tiles, int n_tiles, cudaStream_t st) {
  using Cfg = ConvHaloCfg<BN, MT>;
  B2_SET_SMEM_ONCE(Cfg::kSmemBytes, conv3d_halo_kernel<BN, MT>);
  const int sms = sm_count();
  a.work_x = ceil_div(m_tiles, MT); a.work_y = n_tiles; a.work_z = 1;
  const long long total = (long long)a.work_x * a.work_y;
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  conv3d_halo_kernel<BN, MT><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(maps, a);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int launch_conv_halo(const ConvTcMaps& maps, const ConvTcArgs& a, int bn, int m_tiles, int n_tiles, cudaStream_t st) {
  switch (bn) {
    case 128: return launch_halo<128, 2>(maps, a, m_tiles, n_tiles, st);
    case 64: return launch_halo<64, 2>(maps, a, m_tiles, n_tiles, st);
  }
  return fail(B200LDM_BAD_SHAPE, "conv_halo: unsupported Cout tile %d", bn);
}

}  // namespace b200ldm
