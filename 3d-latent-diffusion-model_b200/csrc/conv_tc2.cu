// conv_tc2.cu - persistent CTA-pair variant of the tcgen05 implicit-GEMM Conv3d (same math and tensor maps as
// conv_tc.cu; see there for the GEMM view).
//
// Why a second kernel: conv_tc.cu runs one CTA per output tile, so every tile pays the pipeline fill, the TMEM
// allocation and - above all - a ~10 K-cycle epilogue during which the tensor pipe idles (clock64 traces,
// scripts/conv_trace.py), and launches with few tiles lose up to 40 % to wave quantisation.  Here
//   * 74 CTA pairs (one per TPC) stay resident and walk the tile list round-robin;
//   * a pair computes a 256-voxel x BN tile with ONE tcgen05.mma.cta_group::2 stream: each CTA stages its own
//     128-voxel A box and HALF of the weight tile (BN/2 rows), so a k-iteration moves 16 KB + BN*64 B into each SM for
//     128 x BN x 64 MACs - the same 64 B/clk/SM as the 256x256 single-CTA tile, but with only BN TMEM columns per
//     accumulator;
//   * that leaves room for TWO accumulators (2 x BN <= 512 columns): the MMA issuer starts tile i+1 while the epilogue
//     warps drain tile i through a dedicated, double-buffered 2 x 16 KB staging area and TMA stores.
// Barriers: full[s] lives in the leader (both producers credit it), empty[s] / tmem_full[a] are multicast to both
// CTAs by tcgen05.commit, tmem_empty[a] lives in the leader and collects the 8 epilogue warps of the pair.
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace b200ldm {

template <int BN>
struct ConvTc2Cfg {
  static constexpr int kBHalf = BN / 2;            // weight rows staged by each CTA of the pair
  static constexpr int kBBytes = kBHalf * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = 2 * 16384;      // two [128 rows x 64 ch] bf16 staging buffers
  static constexpr int kBarBytes = 512;
  static constexpr int kFlagBytes = 1024;          // row-valid flags + per-chunk partial sums
  static constexpr int kTail = 1024 /*align slack*/ + kBarBytes + kFlagBytes;
  static constexpr int kBudget = 232448 - kOutBytes - kTail;
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  static constexpr int kTmemCols = 2 * BN;         // two accumulators
  static constexpr int kThreads = 192;             // warp0 TMA, warp1 MMA + TMEM owner, warps 2..5 epilogue
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + kTail;
  static_assert(kStages >= 4, "pipeline depth");
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols >= 32, "TMEM columns");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  static_assert(kStageBytes % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
};

struct TileCoord {
  int n0, d0, h0, w0, co0, phase;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvTcArgs& a, int wk, int rank, int bn) {
  TileCoord c;
  const int x = wk % a.work_x;
  int r = wk / a.work_x;
  const int y = r % a.work_y;
  c.phase = r / a.work_y;
  int t = x * 2 + rank;   // this CTA's 128-voxel box (a box index past the end decodes to n0 >= N: all rows masked)
  c.w0 = (t % a.nb_w) * a.tw; t /= a.nb_w;
  c.h0 = (t % a.nb_h) * a.th; t /= a.nb_h;
  c.d0 = (t % a.nb_d) * a.td; t /= a.nb_d;
  c.n0 = t * a.tn;
  c.co0 = y * bn;
  return c;
}

template <int BN>
__global__ void __launch_bounds__(ConvTc2Cfg<BN>::kThreads, 1)
conv3d_tc2_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs a) {
  using Cfg = ConvTc2Cfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t out_base = smem_base + STAGES * Cfg::kStageBytes;
  const uint32_t bar_base = out_base + Cfg::kOutBytes;
  const uint32_t full_bar = bar_base;                       // [STAGES]  (leader's copy is the live one)
  const uint32_t empty_bar = bar_base + 8 * STAGES;         // [STAGES]
  const uint32_t pfull_bar = bar_base + 16 * STAGES;        // [STAGES] relay mode: the peer's stage is full (leader's copy)
  const uint32_t tfull_bar = bar_base + 24 * STAGES;        // [2] accumulator complete (multicast to both CTAs)
  const uint32_t tempty_bar = tfull_bar + 16;               // [2] accumulator drained (leader's copy, 8 warp arrivals)
  const uint32_t res_bar = tempty_bar + 16;                 // [2] residual tile landed in staging buffer b
  const uint32_t tmem_slot = res_bar + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));
  uint8_t* flag_gen = smem_gen + (bar_base - smem_base) + Cfg::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pairs = gridDim.x >> 1;
  const int pair = blockIdx.x >> 1;
  const int total = a.work_x * a.work_y * a.work_z;
  const int taps = a.up ? 8 : a.ksize * a.ksize * a.ksize;
  const int nchunks = a.nchunk1 + a.nchunk2;
  const int iters_main = taps * nchunks;
  const int iters_total = iters_main + a.nsk1 + a.nsk2;
  // experiment switches (B200LDM_CONV_TC2_MODE): bit0 = dry run (no loads: barrier handshake + MMA rate only),
  // bit1 = relay (each CTA's loads credit its OWN full barrier; the peer's idle MMA warp forwards one arrive per stage)
  const bool dry = (a.ashift & 1) != 0, relay = (a.ashift & 2) != 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&maps.w);
    ptx::prefetch_tmap(&maps.a[0]);
    ptx::prefetch_tmap(&maps.y[0]);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar + 8 * s, 1);
      ptx::mbar_init(empty_bar + 8 * s, 1);
      ptx::mbar_init(pfull_bar + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull_bar + 8 * i, 1);
      ptx::mbar_init(tempty_bar + 8 * i, 8);
      ptx::mbar_init(res_bar + 8 * i, 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();   // the peer's barriers exist before anyone credits / arrives on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // debug timeline (b200ldm_debug_set_conv_trace): [0] start, [1] MMA waits on full, [2] MMA waits on tmem_empty,
  // [3] producer waits on empty, [4] epilogue waits on tmem_full, [5] epilogue busy, [6] tiles, [7] end
  unsigned long long* trc = a.trace ? a.trace + (size_t)blockIdx.x * 8 : nullptr;
  if (trc && threadIdx.x == 0) trc[0] = clock64();

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    if (lane == 0) {
      const uint32_t tx_pair = 2u * ((uint32_t)a.rows_box * 128u + (uint32_t)Cfg::kBBytes);
      uint32_t itg = 0;
      long long w_empty = 0;
      for (int wk = pair; wk < total; wk += pairs) {
        const TileCoord c = decode_tile(a, wk, (int)rank, BN);
        const int brow0 = c.co0 + (int)rank * Cfg::kBHalf;
        for (int it = 0; it < iters_total; ++it, ++itg) {
          const int s = (int)(itg % STAGES);
          const uint32_t ph = (itg / STAGES) & 1u;
          const long long tw0 = trc ? clock64() : 0;
          ptx::mbar_wait(empty_bar + 8 * s, ph ^ 1u, 0);
          if (trc) w_empty += clock64() - tw0;
          const uint32_t full = full_bar + 8 * s;
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint32_t sb = sa + kABytes;
          if (dry) {
            if (rank == 0) ptx::mbar_arrive(full);
            continue;
          }
          if (relay) ptx::mbar_expect_tx(full, tx_pair >> 1);
          else if (rank == 0) ptx::mbar_expect_tx(full, tx_pair);
          if (it >= iters_main) {
            const int sc = it - iters_main;
            const bool second = sc >= a.nsk1;
            const CUtensorMap* m = second ? &maps.s[1] : &maps.s[0];
            if (relay) {
              ptx::tma_load_5d(sa, m, full, (second ? sc - a.nsk1 : sc) * 64, c.w0, c.h0, c.d0, c.n0);
              ptx::tma_load_3d(sb, &maps.sw, full, sc * 64, brow0, 0);
            } else {
              ptx::tma_load_5d_2sm(sa, m, full, (second ? sc - a.nsk1 : sc) * 64, c.w0, c.h0, c.d0, c.n0);
              ptx::tma_load_3d_2sm(sb, &maps.sw, full, sc * 64, brow0, 0);
            }
            continue;
          }
          const int tap = it / nchunks;
          const int chunk = it - tap * nchunks;
          int kd = 0, kh = 0, kw = 0;
          int pad_d = a.pad_lo, pad_h = a.pad_lo, pad_w = a.pad_lo, wtap = tap;
          if (a.up) {
            kd = tap >> 2; kh = (tap >> 1) & 1; kw = tap & 1;
            pad_d = 1 - (c.phase >> 2); pad_h = 1 - ((c.phase >> 1) & 1); pad_w = 1 - (c.phase & 1);
            wtap = c.phase * 8 + tap;
          } else if (a.ksize == 3) { kd = tap / 9; kh = (tap / 3) % 3; kw = tap % 3; }
          if (a.stride == 1) {
            const bool second = chunk >= a.nchunk1;
            const CUtensorMap* m = second ? &maps.a[1] : &maps.a[0];
            const int c0 = (second ? chunk - a.nchunk1 : chunk) * 64;
            if (relay)
              ptx::tma_load_5d(sa, m, full, c0, c.w0 + kw - pad_w, c.h0 + kh - pad_h, c.d0 + kd - pad_d, c.n0);
            else
              ptx::tma_load_5d_2sm(sa, m, full, c0, c.w0 + kw - pad_w, c.h0 + kh - pad_h, c.d0 + kd - pad_d, c.n0);
          } else {
            const int od = kd - a.pad_lo, oh = kh - a.pad_lo, ow = kw - a.pad_lo;
            const int pd = od & 1, ph2 = oh & 1, pw = ow & 1;
            const CUtensorMap* m = &maps.a[pd * 4 + ph2 * 2 + pw];
            if (relay)
              ptx::tma_load_5d(sa, m, full, chunk * 64, c.w0 + ((ow - pw) >> 1), c.h0 + ((oh - ph2) >> 1),
                               c.d0 + ((od - pd) >> 1), c.n0);
            else
              ptx::tma_load_5d_2sm(sa, m, full, chunk * 64, c.w0 + ((ow - pw) >> 1), c.h0 + ((oh - ph2) >> 1),
                                   c.d0 + ((od - pd) >> 1), c.n0);
          }
          if (relay) ptx::tma_load_3d(sb, &maps.w, full, chunk * 64, brow0, wtap);
          else ptx::tma_load_3d_2sm(sb, &maps.w, full, chunk * 64, brow0, wtap);
        }
      }
      if (trc) trc[3] = (unsigned long long)w_empty;
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA only) ===========================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(256, BN);
      uint32_t itg = 0, tile = 0;
      long long w_full = 0, w_tempty = 0;
      for (int wk = pair; wk < total; wk += pairs, ++tile) {
        const uint32_t acc = tile & 1u, use = tile >> 1;
        long long tw0 = trc ? clock64() : 0;
        ptx::mbar_wait_cluster(tempty_bar + 8 * acc, (use & 1u) ^ 1u, 5);   // both CTAs drained this accumulator
        if (trc) w_tempty += clock64() - tw0;
        ptx::tc_fence_after();
        const uint32_t td = tmem_base + acc * BN;
        for (int it = 0; it < iters_total; ++it, ++itg) {
          const int s = (int)(itg % STAGES);
          const uint32_t ph = (itg / STAGES) & 1u;
          tw0 = trc ? clock64() : 0;
          ptx::mbar_wait(full_bar + 8 * s, ph, 1);
          if (relay) ptx::mbar_wait_cluster(pfull_bar + 8 * s, ph, 6);   // the peer's half of the stage has landed too
          if (trc) w_full += clock64() - tw0;
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_2sm(td, ptx::umma_desc_k_sw128(sa + k * 32), ptx::umma_desc_k_sw128(sb + k * 32), idesc,
                               (it > 0 || k > 0) ? 1u : 0u);
          ptx::umma_commit_2sm(empty_bar + 8 * s, 3);   // stage free in both CTAs once these MMAs retire
        }
        ptx::umma_commit_2sm(tfull_bar + 8 * acc, 3);   // accumulator complete: wake both epilogues
      }
      if (trc) { trc[1] = (unsigned long long)w_full; trc[2] = (unsigned long long)w_tempty; trc[6] = tile; }
    } else if (rank == 1 && lane == 0 && relay && !dry) {
      // relay: one remote arrive per stage tells the leader that this CTA's operands are in place
      const uint32_t pfull_leader = ptx::mapa(pfull_bar, 0);
      uint32_t itg = 0;
      for (int wk = pair; wk < total; wk += pairs)
        for (int it = 0; it < iters_total; ++it, ++itg) {
          const int s = (int)(itg % STAGES);
          ptx::mbar_wait(full_bar + 8 * s, (itg / STAGES) & 1u, 7);
          ptx::mbar_arrive_cluster(pfull_leader + 8 * s);
        }
    }
  } else {
    // =========================== epilogue (4 warps, both CTAs) ===========================
    constexpr int G = BN / 64;
    const int q = warp & 3;          // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;   // accumulator row == voxel slot in this CTA's box
    const bool issuer = (warp == 2) && lane == 0;
    uint8_t* vrow = flag_gen;                                      // [128] row-valid flags
    float* csum = reinterpret_cast<float*>(flag_gen + 128);        // [G*8][2]
    const bool do_stats = a.chunk_stats != nullptr;
    const uint32_t tempty_leader = ptx::mapa(tempty_bar, 0);
    uint32_t tile = 0, gcount = 0, res_ph0 = 0, res_ph1 = 0;
    int t = row;
    const int rw = t % a.tw; t /= a.tw;
    const int rh = t % a.th; t /= a.th;
    const int rd = t % a.td; t /= a.td;
    const int rn = t;
    for (int wk = pair; wk < total; wk += pairs, ++tile) {
      const TileCoord c = decode_tile(a, wk, (int)rank, BN);
      const int n = c.n0 + rn, d = c.d0 + rd, h = c.h0 + rh, w = c.w0 + rw;
      const bool valid = row < a.rows_box && n < a.N && d < a.Do && h < a.Ho && w < a.Wo;
      const float* temb_row = a.temb ? a.temb + (size_t)(a.temb_stride ? n : 0) * a.temb_stride : nullptr;
      const uint32_t acc = tile & 1u, use = tile >> 1;
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const int gmax = min(G, (a.Cout - c.co0 + 63) / 64);   // 64-channel groups present in this Cout tile
      if (do_stats) {
        vrow[row] = valid ? 1 : 0;
        if (row < G * 16) csum[row] = 0.f;
      }
      const long long te0 = (trc && issuer) ? clock64() : 0;
      ptx::mbar_wait_long(tfull_bar + 8 * acc, use & 1u, 2);
      ptx::tc_fence_after();
      const long long te1 = (trc && issuer) ? clock64() : 0;
      if (a.ashift & 4) {   // experiment: no epilogue work at all, only the accumulator hand-back
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + 8 * acc);
        continue;
      }
#pragma unroll 1
      for (int g = 0; g < gmax; ++g, ++gcount) {
        const uint32_t b = gcount & 1u;
        const uint32_t stg = out_base + b * 16384;
        uint8_t* stg_gen = smem_gen + (out_base - smem_base) + b * 16384;
        uint8_t* rowp = stg_gen + (size_t)row * 128;
        uint32_t rr[4][16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) ptx::tmem_ld16(tacc + (uint32_t)(g * 64 + cc * 16), rr[cc]);
        if (issuer) {
          ptx::tma_store_wait_read<1>();   // the store that last used buffer b (two groups ago) has read it out
          if (a.residual) {
            ptx::mbar_expect_tx(res_bar + 8 * b, (uint32_t)a.rows_box * 128u);
            ptx::tma_load_5d(stg, &maps.r, res_bar + 8 * b, c.co0 + g * 64, c.w0, c.h0, c.d0, c.n0);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // buffer b is free (and the residual load is on its way)
        ptx::tmem_ld_wait();
        if (g == gmax - 1) {
          // every column of this accumulator is in registers: hand it back to the MMA issuer
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + 8 * acc);
        }
        if (a.residual) {
          const uint32_t rp = b ? res_ph1 : res_ph0;
          ptx::mbar_wait(res_bar + 8 * b, rp, 3);
          if (b) res_ph1 ^= 1u; else res_ph0 ^= 1u;
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int co = c.co0 + g * 64 + cc * 16;
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
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          ptx::tma_store_5d(&maps.y[c.phase], stg, c.co0 + g * 64, c.w0, c.h0, c.d0, c.n0);
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
            atomicAdd(&csum[(g * 8 + lane) * 2], s1);
            atomicAdd(&csum[(g * 8 + lane) * 2 + 1], s2);
          }
        }
      }
      if (do_stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (row < gmax * 16 && c.n0 < a.N) {
          const int k = row >> 1, which = row & 1;   // chunk k of this Cout tile, 0 = sum / 1 = sumsq
          atomicAdd(a.chunk_stats + ((size_t)c.n0 * (a.Cout >> 3) + (c.co0 >> 3) + k) * 2 + which, (double)csum[row]);
        }
      }
      if (trc && issuer) { trc[4] += (unsigned long long)(te1 - te0); trc[5] += (unsigned long long)(clock64() - te1); }
    }
    if (issuer) ptx::tma_store_wait_all();   // smem must outlive the bulk reads; writes complete before exit
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (trc && threadIdx.x == 0) trc[7] = clock64();
  ptx::cluster_sync();   // nobody leaves (or frees TMEM) while the peer can still touch this CTA's smem / barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ host
static unsigned long long* g_conv_trace2 = nullptr;
void set_conv_trace2(void* p) { g_conv_trace2 = reinterpret_cast<unsigned long long*>(p); }

bool conv_tc2_supported(int bn, const ConvTcArgs& a) {
  return (bn == 256 || bn == 128 || bn == 64) && a.tma_epilogue && a.splits == 1 && !a.out_f32 &&
         !a.gn_stats;
}

template <int BN>
static int launch_tc2(const ConvTcMaps& maps, ConvTcArgs a, int m_tiles, int n_tiles, int zdim, cudaStream_t st) {
  using Cfg = ConvTc2Cfg<BN>;
  B2_SET_SMEM_ONCE(Cfg::kSmemBytes, conv3d_tc2_kernel<BN>);
  static std::atomic<int> pairs_of[kMaxDevices];   // resident CTA pairs, per device
  const int dev = current_device();
  int max_pairs = pairs_of[dev].load(std::memory_order_relaxed);
  if (max_pairs <= 0) {
    const int sms = sm_count();
    cudaLaunchConfig_t q{};
    q.gridDim = dim3((unsigned)(sms & ~1), 1, 1);
    q.blockDim = dim3(Cfg::kThreads, 1, 1);
    q.dynamicSmemBytes = Cfg::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    q.attrs = at; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv3d_tc2_kernel<BN>, &q) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = sms / 2;
    }
    max_pairs = n < sms / 2 ? n : sms / 2;
    const char* e = getenv("B200LDM_CONV_TC2_PAIRS");
    if (e && atoi(e) > 0 && atoi(e) < max_pairs) max_pairs = atoi(e);
    pairs_of[dev].store(max_pairs, std::memory_order_relaxed);
  }
  a.work_x = ceil_div(m_tiles, 2); a.work_y = n_tiles; a.work_z = zdim;
  a.trace = g_conv_trace2;
  {
    static const int mode = [] { const char* e = getenv("B200LDM_CONV_TC2_MODE"); return e ? atoi(e) : 0; }();
    a.ashift = mode;
  }
  const long long total = (long long)a.work_x * a.work_y * a.work_z;
  const int pairs = (int)(total < max_pairs ? total : max_pairs);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
  cfg.blockDim = dim3(Cfg::kThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3d_tc2_kernel<BN>, maps, a));
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

int launch_conv_tc2(const ConvTcMaps& maps, const ConvTcArgs& a, int bn, int m_tiles, int n_tiles, int zdim,
                    cudaStream_t st) {
  switch (bn) {
    case 256: return launch_tc2<256>(maps, a, m_tiles, n_tiles, zdim, st);
    case 128: return launch_tc2<128>(maps, a, m_tiles, n_tiles, zdim, st);
    case 64: return launch_tc2<64>(maps, a, m_tiles, n_tiles, zdim, st);
  }
  return fail(B200LDM_BAD_SHAPE, "conv_tc2: unsupported Cout tile %d", bn);
}

}  // namespace b200ldm
