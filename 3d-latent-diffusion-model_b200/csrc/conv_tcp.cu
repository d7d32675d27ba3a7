// conv_tcp.cu - persistent single-CTA variant of the tcgen05 implicit-GEMM Conv3d for problems whose tiles are SHORT
// (few k-iterations and / or narrow Cout), where conv_tc.cu's per-tile prologue + epilogue (~14 K cycles, tensor pipe
// idle) is as long as the main loop itself: the AutoencoderKL's 64/128-channel convs at full image resolution
// (27 k-iterations, load-bound at ~570 cycles each) and the 1x1x1 projections (4-16 k-iterations).
//
// One CTA per SM stays resident and walks the tile list.  The accumulator set of a tile (MT sub-tiles x BN columns)
// uses at most half of TMEM, so there are TWO sets: the MMA issuer starts tile i+1 while the MT*4 epilogue warps
// drain tile i through their own double-buffered 16 KB staging slots and TMA stores.  Same tensor maps, weights and
// epilogue math as conv_tc.cu; all barriers are CTA-local (the CTA-pair experiment, conv_tc2.cu, showed remote
// signalling to be the expensive part).
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace b200ldm {

// EW: epilogue warp SETS per sub-tile (4 warps each, one per TMEM lane quarter); set h drains the 64-column groups
// g = h, h + EW, ... of its sub-tile through its own staging buffers.  EW = 2 for the 128 x 256 tile doubles the drain
// rate where the main loop is short (1x1x1 projections: 4-16 k-iterations against ~10 K cycles of epilogue).
template <int BN, int MT, int EW = 1>
struct ConvTcpCfg {
  static constexpr int kSets = MT * EW;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = MT * kABytes + kBBytes;
  static constexpr int kOutBytes = kSets * 2 * 16384;   // per set: two [128 rows x 64 ch] bf16 staging buffers
  static constexpr int kBarBytes = 512;
  static constexpr int kFlagBytes = kSets * 768;        // per set: row-valid flags + per-chunk partial sums
  static constexpr int kTail = 1024 /*align slack*/ + kBarBytes + kFlagBytes;
  static constexpr int kBudget = 232448 - kOutBytes - kTail;
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  static constexpr int kAccCols = MT * BN;           // one accumulator set
  static constexpr int kTmemCols = 2 * kAccCols;
  static constexpr int kThreads = 64 + kSets * 128;  // warp0 TMA, warp1 MMA + TMEM owner, 4 epilogue warps per set
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + kTail;
  static_assert(kStages >= 3, "pipeline depth");
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols >= 32, "TMEM columns");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  static_assert(kStageBytes % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
  static_assert(BN / 64 * 8 * 2 * 4 * (BN <= 128 ? 4 : 1) + 128 <= 768, "flag area (chunk widths < 8 only for BN <= 128)");
};

struct TcpTile {
  int n0, d0, h0, w0;
};

__device__ __forceinline__ void tcp_decode_work(const ConvTcArgs& a, int wk, int bn, int* x, int* co0, int* phase) {
  *x = wk % a.work_x;
  const int r = wk / a.work_x;
  *co0 = (r % a.work_y) * bn;
  *phase = r / a.work_y;
}
__device__ __forceinline__ TcpTile tcp_box(const ConvTcArgs& a, int t) {
  TcpTile c;   // a box index past the end decodes to n0 >= N: TMA zero-fills the rows and the epilogue masks them
  c.w0 = (t % a.nb_w) * a.tw; t /= a.nb_w;
  c.h0 = (t % a.nb_h) * a.th; t /= a.nb_h;
  c.d0 = (t % a.nb_d) * a.td; t /= a.nb_d;
  c.n0 = t * a.tn;
  return c;
}

template <int BN, int MT, int EW>
__global__ void __launch_bounds__(ConvTcpCfg<BN, MT, EW>::kThreads, 1)
conv3d_tcp_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs a) {
  using Cfg = ConvTcpCfg<BN, MT, EW>;
  constexpr int SETS = Cfg::kSets;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t out_base = smem_base + STAGES * Cfg::kStageBytes;
  const uint32_t bar_base = out_base + Cfg::kOutBytes;
  const uint32_t full_bar = bar_base;                    // [STAGES]
  const uint32_t empty_bar = bar_base + 8 * STAGES;      // [STAGES]
  const uint32_t tfull_bar = bar_base + 16 * STAGES;     // [2] accumulator set complete
  const uint32_t tempty_bar = tfull_bar + 16;            // [2] accumulator set drained (SETS*4 warp arrivals)
  const uint32_t res_bar = tempty_bar + 16;              // [SETS][2] residual tile landed in staging buffer
  const uint32_t tmem_slot = res_bar + 8 * SETS * 2;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));
  uint8_t* flag_gen = smem_gen + (bar_base - smem_base) + Cfg::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = a.work_x * a.work_y * a.work_z;
  const int taps = a.up ? 8 : a.ksize * a.ksize * a.ksize;
  const int nchunks = a.nchunk1 + a.nchunk2;
  const int iters_main = taps * nchunks;
  const int iters_total = iters_main + a.nsk1 + a.nsk2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&maps.w);
    ptx::prefetch_tmap(&maps.a[0]);
    ptx::prefetch_tmap(&maps.y[0]);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar + 8 * s, 1);
      ptx::mbar_init(empty_bar + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull_bar + 8 * i, 1);
      ptx::mbar_init(tempty_bar + 8 * i, SETS * 4);
    }
    for (int i = 0; i < SETS * 2; ++i) ptx::mbar_init(res_bar + 8 * i, 1);
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
      // debug (B200LDM_CONV_ASHIFT, timing only - results are wrong): bit 0 = no activation loads, bit 1 = no weight
      // loads, bit 2 (epilogue warps) = hand the accumulator back without draining it
      const bool dry_a = (a.ashift & 1) != 0, dry_b = (a.ashift & 2) != 0;
      const uint32_t tx_bytes = (dry_a ? 0u : (uint32_t)MT * (uint32_t)a.rows_box * 128u) + (dry_b ? 0u : (uint32_t)Cfg::kBBytes);
      uint32_t itg = 0;
      for (int wk = blockIdx.x; wk < total; wk += gridDim.x) {
        int x, co0, phase;
        tcp_decode_work(a, wk, BN, &x, &co0, &phase);
        TcpTile c[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) c[mt] = tcp_box(a, x * MT + mt);
        for (int it = 0; it < iters_total; ++it, ++itg) {
          const int s = (int)(itg % STAGES);
          const uint32_t ph = (itg / STAGES) & 1u;
          ptx::mbar_wait(empty_bar + 8 * s, ph ^ 1u, 0);
          const uint32_t full = full_bar + 8 * s;
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint32_t sb = sa + MT * kABytes;
          if (dry_a && dry_b) { ptx::mbar_arrive(full); continue; }
          ptx::mbar_expect_tx(full, tx_bytes);
          if (it >= iters_main) {
            const int sc = it - iters_main;
            const bool second = sc >= a.nsk1;
            const CUtensorMap* m = second ? &maps.s[1] : &maps.s[0];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              ptx::tma_load_5d(sa + mt * kABytes, m, full, (second ? sc - a.nsk1 : sc) * 64, c[mt].w0, c[mt].h0,
                               c[mt].d0, c[mt].n0);
            ptx::tma_load_3d(sb, &maps.sw, full, sc * 64, co0, 0);
            continue;
          }
          const int tap = it / nchunks;
          const int chunk = it - tap * nchunks;
          int kd = 0, kh = 0, kw = 0;
          int pad_d = a.pad_lo, pad_h = a.pad_lo, pad_w = a.pad_lo, wtap = tap;
          if (a.up) {
            kd = tap >> 2; kh = (tap >> 1) & 1; kw = tap & 1;
            pad_d = 1 - (phase >> 2); pad_h = 1 - ((phase >> 1) & 1); pad_w = 1 - (phase & 1);
            wtap = phase * 8 + tap;
          } else if (a.ksize == 3) { kd = tap / 9; kh = (tap / 3) % 3; kw = tap % 3; }
          if (a.stride == 1) {
            const bool second = chunk >= a.nchunk1;
            const CUtensorMap* m = second ? &maps.a[1] : &maps.a[0];
            const int c0 = (second ? chunk - a.nchunk1 : chunk) * 64;
            if (!dry_a) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
                ptx::tma_load_5d(sa + mt * kABytes, m, full, c0, c[mt].w0 + kw - pad_w, c[mt].h0 + kh - pad_h,
                                 c[mt].d0 + kd - pad_d, c[mt].n0);
            }
          } else {
            const int od = kd - a.pad_lo, oh = kh - a.pad_lo, ow = kw - a.pad_lo;
            const int pd = od & 1, ph2 = oh & 1, pw = ow & 1;
            const CUtensorMap* m = &maps.a[pd * 4 + ph2 * 2 + pw];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              ptx::tma_load_5d(sa + mt * kABytes, m, full, chunk * 64, c[mt].w0 + ((ow - pw) >> 1),
                               c[mt].h0 + ((oh - ph2) >> 1), c[mt].d0 + ((od - pd) >> 1), c[mt].n0);
          }
          if (!dry_b) ptx::tma_load_3d(sb, &maps.w, full, chunk * 64, co0, wtap);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BN);
      constexpr uint32_t dhi = ptx::umma_desc_hi_sw128();
      // descriptor low words: base of stage 0, then one add per stage / sub-tile / k-slice (see umma_bf16_lohi)
      const uint32_t a_lo0 = ptx::umma_desc_lo_sw128(smem_base);
      const uint32_t b_lo0 = ptx::umma_desc_lo_sw128(smem_base + MT * kABytes);
      uint32_t itg = 0, tile = 0;
      for (int wk = blockIdx.x; wk < total; wk += gridDim.x, ++tile) {
        const uint32_t acc = tile & 1u, use = tile >> 1;
        ptx::mbar_wait(tempty_bar + 8 * acc, (use & 1u) ^ 1u, 5);   // the epilogue drained this accumulator set
        ptx::tc_fence_after();
        const uint32_t td = tmem_base + acc * Cfg::kAccCols;
        uint32_t s = itg % STAGES, ph = (itg / STAGES) & 1u;
        for (int it = 0; it < iters_total; ++it, ++itg) {
          ptx::mbar_wait(full_bar + 8 * s, ph, 1);
          ptx::tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (uint32_t)(Cfg::kStageBytes >> 4);
          const uint32_t b_lo = b_lo0 + s * (uint32_t)(Cfg::kStageBytes >> 4);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_lohi(td + mt * BN, a_lo + (uint32_t)((mt * kABytes + k * 32) >> 4), b_lo + (uint32_t)(k * 2),
                                  dhi, dhi, idesc, (it > 0 || k > 0) ? 1u : 0u);
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
    const int es = (warp - 2) >> 2;  // epilogue warp set
    const int mt = es % MT;          // which sub-tile / accumulator of the accumulator set
    const int half = es / MT;        // this set drains the 64-column groups g = half, half + EW, ...
    const int q = warp & 3;          // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;   // accumulator row == voxel slot in the box
    const bool issuer = (((warp - 2) & 3) == 0) && lane == 0;
    const uint32_t my_out = out_base + (uint32_t)es * 2 * 16384;
    uint8_t* my_out_gen = smem_gen + (out_base - smem_base) + (size_t)es * 2 * 16384;
    const uint32_t my_res = res_bar + 8 * es * 2;
    uint8_t* vrow = flag_gen + es * 768;                        // [128] row-valid flags
    float* csum = reinterpret_cast<float*>(vrow + 128);         // [G*8][2]
    const bool do_stats = a.chunk_stats != nullptr;
    const int ns = 8 / a.cw;   // sub-sums per 8-channel chunk (chunk width 8, 4 or 2)
    uint32_t tile = 0, gcount = 0, res_ph0 = 0, res_ph1 = 0;
    int t = row;
    const int rw = t % a.tw; t /= a.tw;
    const int rh = t % a.th; t /= a.th;
    const int rd = t % a.td; t /= a.td;
    const int rn = t;
    double* s_dst = nullptr;   // GroupNorm chunk sums: cross-tile fp64 accumulator of this thread's entry
    double s_acc = 0.0;
    for (int wk = blockIdx.x; wk < total; wk += gridDim.x, ++tile) {
      int x, co0, phase;
      tcp_decode_work(a, wk, BN, &x, &co0, &phase);
      const TcpTile c = tcp_box(a, x * MT + mt);
      const int n = c.n0 + rn, d = c.d0 + rd, h = c.h0 + rh, w = c.w0 + rw;
      const bool valid = row < a.rows_box && n < a.N && d < a.Do && h < a.Ho && w < a.Wo;
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
      if (a.ashift & 4) {   // debug: no epilogue work at all
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty_bar + 8 * acc);
        continue;
      }
      if (half >= gmax) {   // no group for this set in a narrow last Cout tile: hand the accumulator back untouched
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty_bar + 8 * acc);
      }
#pragma unroll 1
      for (int g = half; g < gmax; g += EW, ++gcount) {
        const uint32_t b = gcount & 1u;
        const uint32_t stg = my_out + b * 16384;
        uint8_t* stg_gen = my_out_gen + b * 16384;
        uint8_t* rowp = stg_gen + (size_t)row * 128;
        uint32_t rr[4][16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) ptx::tmem_ld16(tacc + (uint32_t)(g * 64 + cc * 16), rr[cc]);
        if (issuer) {
          ptx::tma_store_wait_read<1>();   // the store that last used buffer b (two groups ago) has read it out
          if (a.residual) {
            ptx::mbar_expect_tx(my_res + 8 * b, (uint32_t)a.rows_box * 128u);
            ptx::tma_load_5d(stg, &maps.r, my_res + 8 * b, co0 + g * 64, c.w0, c.h0, c.d0, c.n0);
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + es) : "memory");   // buffer b is free, flags are visible
        ptx::tmem_ld_wait();
        if (g + EW >= gmax) {
          // every column this set reads is in registers: hand the accumulator back to the MMA issuer
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(tempty_bar + 8 * acc);
        }
        if (a.residual) {
          const uint32_t rp = b ? res_ph1 : res_ph0;
          ptx::mbar_wait(my_res + 8 * b, rp, 3);
          if (b) res_ph1 ^= 1u; else res_ph0 ^= 1u;
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
        asm volatile("bar.sync %0, 128;" ::"r"(1 + es) : "memory");
        if (issuer) {
          ptx::tma_store_5d(&maps.y[phase], stg, co0 + g * 64, c.w0, c.h0, c.d0, c.n0);
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
        }
      }
      if (do_stats) {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + es) : "memory");
        if (row < gmax * 16 * ns && ((row / (16 * ns)) % EW) == half && c.n0 < a.N) {
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
bool conv_tcp_supported(int bn, const ConvTcArgs& a) {
  return (bn == 256 || bn == 128 || bn == 64) && a.tma_epilogue && a.splits == 1 && !a.out_f32 && !a.gn_stats;
}
int conv_tcp_mt(int bn) { return bn == 256 ? 1 : 2; }

template <int BN, int MT, int EW = 1>
static int launch_tcp(const ConvTcMaps& maps, ConvTcArgs a, int m_tiles, int n_tiles, int zdim, cudaStream_t st) {
  using Cfg = ConvTcpCfg<BN, MT, EW>;
  B2_SET_SMEM_ONCE(Cfg::kSmemBytes, conv3d_tcp_kernel<BN, MT, EW>);
  const int sms = sm_count();
  a.work_x = ceil_div(m_tiles, MT); a.work_y = n_tiles; a.work_z = zdim;
  const long long total = (long long)a.work_x * a.work_y * a.work_z;
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  conv3d_tcp_kernel<BN, MT, EW><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(maps, a);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int launch_conv_tcp(const ConvTcMaps& maps, const ConvTcArgs& a, int bn, int m_tiles, int n_tiles, int zdim,
                    cudaStream_t st, bool wide_epilogue) {
  switch (bn) {
    case 256: return wide_epilogue ? launch_tcp<256, 1, 2>(maps, a, m_tiles, n_tiles, zdim, st)
                                   : launch_tcp<256, 1>(maps, a, m_tiles, n_tiles, zdim, st);
    case 128: return launch_tcp<128, 2>(maps, a, m_tiles, n_tiles, zdim, st);
    case 64: return launch_tcp<64, 2>(maps, a, m_tiles, n_tiles, zdim, st);
  }
  return fail(B200LDM_BAD_SHAPE, "conv_tcp: unsupported Cout tile %d", bn);
}

}  // namespace b200ldm
