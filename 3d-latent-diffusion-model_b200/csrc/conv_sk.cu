// conv_sk.cu - stream-K variant of the tcgen05 implicit-GEMM Conv3d for the 256 x 256 tile (Cout tiles of 256, two
// 128-voxel boxes per tile): the launches of the 20^3 / 10^3 UNet levels whose tile count is not a multiple of the SM
// count.  conv_tc.cu runs those as whole tiles - 256 tiles on 148 SMs are two waves for 1.73 waves of work, and each
// tile pays its own prologue + epilogue with the tensor pipe idle (~14 K of ~125 K cycles); split-K by a fixed factor
// pays fp32 slabs for EVERY tile plus a finalize launch.
//
// Here the k-iterations of all tiles form ONE sequence (tile-major), cut into gridDim.x = #SMs equal spans; CTA c
// runs span c.  A span covers at most: the TAIL of one tile, whole tiles, the HEAD of one tile.  Per segment:
//   * whole tile                      -> the usual epilogue (bias / temb / residual / bf16 TMA store / GroupNorm chunk sums);
//   * tail or middle part (k0 > 0)    -> CONTRIBUTOR: the fp32 partial tile goes to this CTA's slot of the workspace
//                                        (coalesced: a warp stores 512 contiguous bytes), then its flag is released;
//   * head part (k0 == 0, k1 < iters) -> OWNER: adds the partial tiles of the following CTAs that share the tile
//                                        (acquire-polls their flags, resets them) and runs the usual epilogue.
// A contribution is always a CTA's FIRST segment and an owned head its LAST, so contributions are posted before
// anything is waited for, and an owner's contributors (the next 1-3 CTAs, which run the same number of k-iterations)
// finish when it does: the poll normally succeeds at once.  Forward progress needs the contributors to be scheduled,
// which holds with gridDim.x <= #SMs (one CTA per SM) and in-order CTA dispatch; every poll is bounded and traps.
// Flags are 64-bit words in the caller's workspace holding a per-CTA magic value when set; owners clear what they
// consume, so the workspace needs no initialisation (a stale match needs a 2^-64 collision with uninitialised memory).
//
// The ring of smem stages runs across segment boundaries (the producer prefetches the next segment while the epilogue
// drains TMEM); there is ONE accumulator set (2 x 256 columns = all of TMEM), so the MMA issuer waits for the epilogue
// warps to have read the accumulators into registers before the next segment's first MMA.
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace b200ldm {

constexpr int kSkBN = 256, kSkMT = 2;

struct ConvSkCfg {
  static constexpr int kBBytes = kSkBN * 128;
  static constexpr int kStageBytes = kSkMT * kABytes + kBBytes;     // 64 KB
  static constexpr int kOutBytes = kSkMT * 16384;                   // per sub-tile one [128 rows x 64 ch] bf16 staging buffer
  static constexpr int kBarBytes = 512;
  static constexpr int kFlagBytes = kSkMT * 768;
  static constexpr int kTail = 1024 /*align slack*/ + kBarBytes + kFlagBytes;
  static constexpr int kStages = (232448 - kOutBytes - kTail) / kStageBytes;
  static constexpr int kTmemCols = kSkMT * kSkBN;
  static constexpr int kThreads = 64 + kSkMT * 128;
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + kTail;
  static constexpr int kSlotFloats = kSkMT * 128 * kSkBN;           // one partial tile: 256 KB
  static_assert(kStages >= 3, "pipeline depth");
  static_assert(kTmemCols == 512, "one accumulator set = all of TMEM");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

__device__ __forceinline__ unsigned long long sk_magic(int cta) {
  return 0x9E3779B97F4A7C15ull ^ ((unsigned long long)(cta + 1) * 0xD1B54A32D192ED03ull);
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct SkBox {
  int n0, d0, h0, w0;
};
__device__ __forceinline__ SkBox sk_box(const ConvTcArgs& a, int t) {
  SkBox c;   // a box index past the end decodes to n0 >= N: TMA zero-fills the rows and the epilogue masks them
  c.w0 = (t % a.nb_w) * a.tw; t /= a.nb_w;
  c.h0 = (t % a.nb_h) * a.th; t /= a.nb_h;
  c.d0 = (t % a.nb_d) * a.td; t /= a.nb_d;
  c.n0 = t * a.tn;
  return c;
}

__global__ void __launch_bounds__(ConvSkCfg::kThreads, 1)
conv3d_sk_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs a) {
  using Cfg = ConvSkCfg;
  constexpr int BN = kSkBN, MT = kSkMT, STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t out_base = smem_base + STAGES * Cfg::kStageBytes;
  const uint32_t bar_base = out_base + Cfg::kOutBytes;
  const uint32_t full_bar = bar_base;                    // [STAGES]
  const uint32_t empty_bar = bar_base + 8 * STAGES;      // [STAGES]
  const uint32_t tfull_bar = bar_base + 16 * STAGES;     // accumulators of the current segment complete
  const uint32_t tempty_bar = tfull_bar + 8;             // accumulators read out (MT*4 warp arrivals)
  const uint32_t res_bar = tempty_bar + 8;               // [MT] residual tile landed in the staging buffer
  const uint32_t tmem_slot = res_bar + 8 * MT;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));
  uint8_t* flag_gen = smem_gen + (bar_base - smem_base) + Cfg::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = a.ksize * a.ksize * a.ksize;
  const int nchunks = a.nchunk1 + a.nchunk2;
  const int iters_main = taps * nchunks;
  const int ipt = iters_main + a.nsk1 + a.nsk2;                      // k-iterations per tile
  const int tiles = a.work_x * a.work_y;
  const long long total = (long long)tiles * ipt;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  const long long g_begin = total * cta / G, g_end = total * (cta + 1) / G;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&maps.w);
    ptx::prefetch_tmap(&maps.a[0]);
    ptx::prefetch_tmap(&maps.y[0]);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar + 8 * s, 1);
      ptx::mbar_init(empty_bar + 8 * s, 1);
    }
    ptx::mbar_init(tfull_bar, 1);
    ptx::mbar_init(tempty_bar, MT * 4);
    for (int i = 0; i < MT; ++i) ptx::mbar_init(res_bar + 8 * i, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // debug (b200ldm_debug_set_conv_trace): 16 clock64 stamps per CTA - [0] start, [1] first stage landed, then per segment
  // (up to 4) [2+3i] first MMA issued / [3+3i] last MMA issued / [4+3i] epilogue done, [15] CTA end
  unsigned long long* trc = a.trace ? a.trace + (size_t)blockIdx.x * 16 : nullptr;
  if (trc && threadIdx.x == 0) trc[0] = clock64();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)MT * (uint32_t)a.rows_box * 128u + (uint32_t)Cfg::kBBytes;
      uint32_t itg = 0;
      for (long long g = g_begin; g < g_end;) {
        const int tile = (int)(g / ipt);
        const int k0 = (int)(g - (long long)tile * ipt);
        const long long left = g_end - g;
        const int k1 = (int)((long long)(ipt - k0) < left ? ipt : k0 + left);
        const int x = tile % a.work_x, co0 = (tile / a.work_x) * BN;
        SkBox c[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) c[mt] = sk_box(a, x * MT + mt);
        for (int it = k0; it < k1; ++it, ++itg) {
          const int s = (int)(itg % STAGES);
          const uint32_t ph = (itg / STAGES) & 1u;
          ptx::mbar_wait(empty_bar + 8 * s, ph ^ 1u, 0);
          const uint32_t full = full_bar + 8 * s;
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint32_t sb = sa + MT * kABytes;
          ptx::mbar_expect_tx(full, tx_bytes);
          if (it >= iters_main) {   // fused 1x1x1 skip conv: same output box, no tap shift, its own sources and weights
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
          if (a.ksize == 3) { kd = tap / 9; kh = (tap / 3) % 3; kw = tap % 3; }
          if (a.stride == 1) {
            const bool second = chunk >= a.nchunk1;
            const CUtensorMap* m = second ? &maps.a[1] : &maps.a[0];
            const int c0 = (second ? chunk - a.nchunk1 : chunk) * 64;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              ptx::tma_load_5d(sa + mt * kABytes, m, full, c0, c[mt].w0 + kw - a.pad_lo, c[mt].h0 + kh - a.pad_lo,
                               c[mt].d0 + kd - a.pad_lo, c[mt].n0);
          } else {
            const int od = kd - a.pad_lo, oh = kh - a.pad_lo, ow = kw - a.pad_lo;
            const int pd = od & 1, ph2 = oh & 1, pw = ow & 1;
            const CUtensorMap* m = &maps.a[pd * 4 + ph2 * 2 + pw];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              ptx::tma_load_5d(sa + mt * kABytes, m, full, chunk * 64, c[mt].w0 + ((ow - pw) >> 1),
                               c[mt].h0 + ((oh - ph2) >> 1), c[mt].d0 + ((od - pd) >> 1), c[mt].n0);
          }
          ptx::tma_load_3d(sb, &maps.w, full, chunk * 64, co0, tap);
        }
        g += k1 - k0;
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BN);
      constexpr uint32_t dhi = ptx::umma_desc_hi_sw128();
      const uint32_t a_lo0 = ptx::umma_desc_lo_sw128(smem_base);
      const uint32_t b_lo0 = ptx::umma_desc_lo_sw128(smem_base + MT * kABytes);
      uint32_t s = 0, ph = 0, seg = 0;
      for (long long g = g_begin; g < g_end; ++seg) {
        const int tile = (int)(g / ipt);
        const int k0 = (int)(g - (long long)tile * ipt);
        const long long left = g_end - g;
        const int k1 = (int)((long long)(ipt - k0) < left ? ipt : k0 + left);
        ptx::mbar_wait(tempty_bar, (seg & 1u) ^ 1u, 5);   // the epilogue holds the previous segment's accumulators in registers
        ptx::tc_fence_after();
        for (int it = k0; it < k1; ++it) {
          ptx::mbar_wait(full_bar + 8 * s, ph, 1);
          ptx::tc_fence_after();
          if (trc && it == k0) {
            if (seg == 0) trc[1] = clock64();
            if (seg < 4) trc[2 + 3 * seg] = clock64();
          }
          const uint32_t a_lo = a_lo0 + s * (uint32_t)(Cfg::kStageBytes >> 4);
          const uint32_t b_lo = b_lo0 + s * (uint32_t)(Cfg::kStageBytes >> 4);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_lohi(tmem_base + mt * BN, a_lo + (uint32_t)((mt * kABytes + k * 32) >> 4), b_lo + (uint32_t)(k * 2),
                                  dhi, dhi, idesc, (it > k0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar + 8 * s);   // stage free once these MMAs retire
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        ptx::umma_commit(tfull_bar);             // segment complete: wake the epilogue warps
        if (trc && seg < 4) trc[3 + 3 * seg] = clock64();
        g += k1 - k0;
      }
    }
  } else {
    // =========================== epilogue (4 warps per sub-tile) ===========================
    constexpr int GR = BN / 64;
    const int mt = (warp - 2) >> 2;  // which sub-tile / accumulator
    const int q = warp & 3;          // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;   // accumulator row == voxel slot in the box
    const bool issuer = (((warp - 2) & 3) == 0) && lane == 0;
    const uint32_t stg = out_base + (uint32_t)mt * 16384;
    uint8_t* stg_gen = smem_gen + (out_base - smem_base) + (size_t)mt * 16384;
    uint8_t* rowp = stg_gen + (size_t)row * 128;
    const uint32_t my_res = res_bar + 8 * mt;
    uint8_t* vrow = flag_gen + mt * 768;                        // [128] row-valid flags
    float* csum = reinterpret_cast<float*>(vrow + 128);         // [GR*8][2]
    const bool do_stats = a.chunk_stats != nullptr;
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(a.sk_ws);
    float4* slots = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a.sk_ws) + a.sk_flag_bytes);
    constexpr int kSlotVec = Cfg::kSlotFloats / 4;              // float4 per partial tile
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BN);
    uint32_t seg = 0, res_ph = 0;
    int t = row;
    const int rw = t % a.tw; t /= a.tw;
    const int rh = t % a.th; t /= a.th;
    const int rd = t % a.td; t /= a.td;
    const int rn = t;
    for (long long g = g_begin; g < g_end; ++seg) {
      const int tile = (int)(g / ipt);
      const int k0 = (int)(g - (long long)tile * ipt);
      const long long left = g_end - g;
      const int k1 = (int)((long long)(ipt - k0) < left ? ipt : k0 + left);
      g += k1 - k0;
      ptx::mbar_wait_long(tfull_bar, seg & 1u, 2);
      ptx::tc_fence_after();
      if (k0 > 0) {
        // ---- contributor: fp32 partial tile -> slot `cta`, float4 index ((mt * 64 + column quad) * 128 + row)
        float4* slot = slots + (size_t)cta * kSlotVec + (size_t)mt * 64 * 128 + row;
#pragma unroll 1
        for (int gq = 0; gq < GR; ++gq) {
          uint32_t rr[4][16];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) ptx::tmem_ld16(tacc + (uint32_t)(gq * 64 + cc * 16), rr[cc]);
          ptx::tmem_ld_wait();
          if (gq == GR - 1) {   // every column is in registers: hand the accumulators back to the MMA issuer
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar);
          }
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              __stcg(slot + (size_t)(gq * 16 + cc * 4 + j) * 128,
                     make_float4(__uint_as_float(rr[cc][4 * j]), __uint_as_float(rr[cc][4 * j + 1]),
                                 __uint_as_float(rr[cc][4 * j + 2]), __uint_as_float(rr[cc][4 * j + 3])));
        }
        __threadfence();
        asm volatile("bar.sync 3, 256;" ::: "memory");   // all 8 epilogue warps have stored (and fenced) their rows
        if (warp == 2 && lane == 0) st_release_u64(flags + cta, sk_magic(cta));
        if (trc && warp == 2 && lane == 0 && seg < 4) trc[4 + 3 * seg] = clock64();
        continue;
      }
      // ---- owner of the tile (whole, or head part + the contributions of the CTAs that follow)
      int ncontrib = 0;
      if (k1 < ipt) {
        const long long tile_end = (long long)(tile + 1) * ipt;
        while (cta + 1 + ncontrib < G && total * (cta + 1 + ncontrib) / G < tile_end) ++ncontrib;
        if (lane == 0) {
          for (int i = 0; i < ncontrib; ++i) {
            const unsigned long long want = sk_magic(cta + 1 + i);
            long long spins = 0;
            while (ld_acquire_u64(flags + cta + 1 + i) != want) {
              __nanosleep(64);
              if (++spins > (1ll << 24)) {   // ~2 s: a contributor was never scheduled - trap instead of hanging the GPU
                printf("b200ldm conv_sk: contribution timeout cta=%d waits for %d\n", cta, cta + 1 + i);
                __trap();
              }
            }
          }
        }
        __syncwarp();
      }
      const int x = tile % a.work_x, co0 = (tile / a.work_x) * BN;
      const SkBox c = sk_box(a, x * MT + mt);
      const int n = c.n0 + rn, d = c.d0 + rd, h = c.h0 + rh, w = c.w0 + rw;
      const bool valid = row < a.rows_box && n < a.N && d < a.Do && h < a.Ho && w < a.Wo;
      const float* temb_row = a.temb ? a.temb + (size_t)(a.temb_stride ? n : 0) * a.temb_stride : nullptr;
      if (do_stats) {
        vrow[row] = valid ? 1 : 0;
        if (row < GR * 16) csum[row] = 0.f;
      }
#pragma unroll 1
      for (int gq = 0; gq < GR; ++gq) {
        uint32_t rr[4][16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) ptx::tmem_ld16(tacc + (uint32_t)(gq * 64 + cc * 16), rr[cc]);
        if (issuer) {
          ptx::tma_store_wait_read<0>();   // the previous store has read the staging buffer out
          if (a.residual) {
            ptx::mbar_expect_tx(my_res, (uint32_t)a.rows_box * 128u);
            ptx::tma_load_5d(stg, &maps.r, my_res, co0 + gq * 64, c.w0, c.h0, c.d0, c.n0);
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");   // the buffer is free, flags are visible
        ptx::tmem_ld_wait();
        if (gq == GR - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(tempty_bar);
        }
        float v[4][16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
          for (int j = 0; j < 16; ++j) v[cc][j] = __uint_as_float(rr[cc][j]);
        for (int i = 0; i < ncontrib; ++i) {   // partial tiles of the contributors, in CTA order (deterministic sum)
          const float4* slot = slots + (size_t)(cta + 1 + i) * kSlotVec + (size_t)mt * 64 * 128 + row;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 f = __ldcg(slot + (size_t)(gq * 16 + cc * 4 + j) * 128);
              v[cc][4 * j] += f.x; v[cc][4 * j + 1] += f.y; v[cc][4 * j + 2] += f.z; v[cc][4 * j + 3] += f.w;
            }
        }
        if (a.residual) {
          ptx::mbar_wait(my_res, res_ph, 3);
          res_ph ^= 1u;
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int co = co0 + gq * 64 + cc * 16;
          if (a.bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + co + j));
              v[cc][j] += bb.x; v[cc][j + 1] += bb.y; v[cc][j + 2] += bb.z; v[cc][j + 3] += bb.w;
            }
          }
          if (temb_row && valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 tt = __ldg(reinterpret_cast<const float4*>(temb_row + co + j));
              v[cc][j] += tt.x; v[cc][j + 1] += tt.y; v[cc][j + 2] += tt.z; v[cc][j + 3] += tt.w;
            }
          }
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            const int chunk = cc * 2 + hlf;  // 16-byte chunk inside the 128-byte row
            uint4* sl = reinterpret_cast<uint4*>(rowp + ((chunk ^ (row & 7)) << 4));
            float f[8];
            if (a.residual) {
              const uint4 u = *sl;
              const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t2 = __bfloat1622float2(hp[e]);
                f[2 * e] = v[cc][hlf * 8 + 2 * e] + t2.x;
                f[2 * e + 1] = v[cc][hlf * 8 + 2 * e + 1] + t2.y;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = v[cc][hlf * 8 + e];
            }
            uint4 o;
            __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) op[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
            *sl = o;
          }
        }
        ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA store
        asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");
        if (issuer) {
          ptx::tma_store_5d(&maps.y[0], stg, co0 + gq * 64, c.w0, c.h0, c.d0, c.n0);
          ptx::tma_store_commit();
        }
        if (do_stats) {
          // thread (slice = row>>3, chunk = row&7) sums 8 rows x 8 channels of the staged tile
          const int cch = row & 7, slice = row >> 3;
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = slice * 8 + i;
            if (vrow[r]) {
              const uint4 u = *reinterpret_cast<const uint4*>(stg_gen + (size_t)r * 128 + ((cch ^ (r & 7)) << 4));
              const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t2 = __bfloat1622float2(hp[e]);
                s1 += t2.x + t2.y;
                s2 = fmaf(t2.x, t2.x, fmaf(t2.y, t2.y, s2));
              }
            }
          }
          s1 += __shfl_xor_sync(0xffffffffu, s1, 8); s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 16); s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
          if (lane < 8) {
            atomicAdd(&csum[(gq * 8 + lane) * 2], s1);
            atomicAdd(&csum[(gq * 8 + lane) * 2 + 1], s2);
          }
          // the statistics read the staged tile: with ONE staging buffer the next group's writes must wait for them
          asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");
        }
      }
      if (do_stats) {
        if (row < GR * 16 && c.n0 < a.N) {
          const int k = row >> 1, which = row & 1;   // chunk k of this Cout tile, 0 = sum / 1 = sumsq
          atomicAdd(a.chunk_stats + ((size_t)c.n0 * (a.Cout >> 3) + (co0 >> 3) + k) * 2 + which, (double)csum[row]);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");   // csum is re-zeroed at the top of the next segment
      }
      if (ncontrib > 0) {
        // every epilogue warp is past its poll and its slot reads: clear the consumed flags for the next launch
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (warp == 2)
          for (int i = lane; i < ncontrib; i += 32) flags[cta + 1 + i] = 0ull;
      }
      if (trc && warp == 2 && lane == 0 && seg < 4) trc[4 + 3 * seg] = clock64();
    }
    if (issuer) ptx::tma_store_wait_all();   // smem must outlive the bulk reads
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
  if (trc && threadIdx.x == 0) trc[15] = clock64();
}

// ------------------------------------------------------------------------------------------ host
static unsigned long long* g_sk_trace = nullptr;
void set_conv_trace_sk(void* p) { g_sk_trace = reinterpret_cast<unsigned long long*>(p); }

bool conv_sk_supported(int bn, const ConvTcArgs& a) {
  return bn == kSkBN && a.tma_epilogue && a.splits == 1 && !a.out_f32 && !a.gn_stats && !a.up && a.cw == 8;
}

// CTAs of the launch: one per SM, fewer when that would leave a CTA less than 8 k-iterations
int conv_sk_grid(int m_tiles, int n_tiles, int iters) {
  const long long total = (long long)ceil_div(m_tiles, kSkMT) * n_tiles * iters;
  const long long cap = total / 8 > 0 ? total / 8 : 1;   // at least 8 k-iterations per CTA
  const int sms = sm_count();
  return (int)(cap < sms ? cap : sms);
}

// workspace: [grid flags of 8 bytes, rounded up to 256 B][grid partial-tile slots of 256 KB]
long long conv_sk_flag_bytes(int grid) { return ((long long)grid * 8 + 255) / 256 * 256; }
long long conv_sk_workspace_bytes(int grid) {
  return conv_sk_flag_bytes(grid) + (long long)grid * ConvSkCfg::kSlotFloats * 4;
}

int launch_conv_sk(const ConvTcMaps& maps, ConvTcArgs a, int m_tiles, int n_tiles, int iters, void* ws,
                   long long ws_bytes, cudaStream_t st) {
  using Cfg = ConvSkCfg;
  B2_SET_SMEM_ONCE(Cfg::kSmemBytes, conv3d_sk_kernel);
  const int grid = conv_sk_grid(m_tiles, n_tiles, iters);
  B2_REQUIRE(ws && ws_bytes >= conv_sk_workspace_bytes(grid) && (reinterpret_cast<uintptr_t>(ws) & 15) == 0, B200LDM_BAD_ARG,
             "conv3d: the stream-K plan needs %lld workspace bytes (b200ldm_conv3d_workspace_bytes)",
             conv_sk_workspace_bytes(grid));
  a.work_x = ceil_div(m_tiles, kSkMT); a.work_y = n_tiles; a.work_z = 1;
  a.sk_ws = ws;
  a.sk_flag_bytes = (int)conv_sk_flag_bytes(grid);
  a.trace = g_sk_trace;
  conv3d_sk_kernel<<<(unsigned)grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(maps, a);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

}  // namespace b200ldm
