// conv_tc.cu - Conv3d (k=1|3, stride 1|2) as an implicit GEMM on tcgen05 tensor cores.
//
//   M = output voxels (a TMA box of tn x td x th x tw <= 128 voxels of the NDHWC tensor)
//   N = Cout tile (BN <= 256), K = taps x (C1 + C2) walked in 64-channel (128-byte) chunks.
//
// No im2col buffer exists: for every (tap, 64-channel chunk) the producer issues one 5-D TMA box
// load of the *shifted* activation window (out-of-bounds rows are zero-filled by the TMA unit =
// the conv zero padding) plus one 3-D box load of the matching [BN x 64] weight slab.  Both land
// in SWIZZLE_128B K-major layout, which is exactly the canonical UMMA operand layout, so one
// elected thread issues 4 x tcgen05.mma (K=16 each) per stage into a 128 x BN fp32 TMEM tile.
// Stride-2 convs use 8 parity-subsampled tensor maps (one per (d,h,w) parity of the tap offset);
// the UNet skip concat / concat conditioning is read in place as a second K range (x2).
// Epilogue warps read TMEM (one accumulator row per thread), add bias / time-embedding /
// residual in fp32, optionally accumulate GroupNorm partial sums, and store bf16 (or fp32).
//
// Warp roles (192 threads): warp0 = TMA producer, warp1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_idx % 4).
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace b200ldm {

// BN = Cout tile, MT = number of 128-row M sub-tiles per CTA that share one weight stage.
// MT = 2 halves the L2->SMEM bytes per FLOP (the v1 kernel was L2-bandwidth bound: profiles/r01_ncu_conv3d_tc_first.md).
// CL = CTAs per cluster sharing the weight stage: each loads BN/CL weight rows and multicasts them to all (CL = 2
// takes the 256x256 tile from 64 to 48 B/clk/SM of L2->SMEM traffic).
template <int BN, int MT>
struct ConvTcCfg {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = MT * kABytes + kBBytes;
  static constexpr int kStages = (196608 / kStageBytes) > 8 ? 8 : (196608 / kStageBytes);
  static constexpr int kTmemCols = (MT * BN) < 32 ? 32 : (MT * BN);
  static constexpr int kThreads = 64 + MT * 128;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*stats*/;
  static_assert(kStages >= 2, "pipeline needs at least two stages");
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns must be a power of two <= 512");
};

template <int BN, int MT, int CL>
__global__ void __launch_bounds__(ConvTcCfg<BN, MT>::kThreads, 1)
conv3d_tc_kernel(const __grid_constant__ ConvTcMaps maps, const ConvTcArgs a) {
  using Cfg = ConvTcCfg<BN, MT>;
  uint32_t cta_rank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::kStageBytes;
  // barriers: full[s] @ +8s, empty[s] @ +8(STAGES+s), tmem_full[mt] @ +16*STAGES + 8mt, tmem slot after them
  const uint32_t tmem_full_bar = bar_base + 16 * STAGES;
  const uint32_t tmem_slot = tmem_full_bar + 8 * MT;
  const uint32_t res_bar = tmem_slot + 8;  // [MT] residual-tile arrival (TMA epilogue)
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));  // generic pointer to the aligned base

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * BN;
  const int phase = a.up ? (int)blockIdx.z : 0;   // (a_d, a_h, a_w) = bits 2,1,0: output voxel (2d+a_d, 2h+a_h, 2w+a_w)
  const int zsplit = a.up ? 0 : (int)blockIdx.z;
  const int taps = a.up ? 8 : a.ksize * a.ksize * a.ksize;
  const int nchunks = a.nchunk1 + a.nchunk2;
  const int iters_main = taps * nchunks;
  const int iters_total = iters_main + a.nsk1 + a.nsk2;   // + the ResBlock's 1x1x1 skip conv accumulated in place
  const int it0 = (int)((long long)zsplit * iters_total / a.splits);
  const int it1 = (int)((long long)(zsplit + 1) * iters_total / a.splits);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&maps.w);
    ptx::prefetch_tmap(&maps.a[0]);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(bar_base + 8 * s, 1);
      ptx::mbar_init(bar_base + 8 * (STAGES + s), CL);  // every CTA of the cluster must release the stage
    }
    for (int mt = 0; mt < MT; ++mt) {
      ptx::mbar_init(tmem_full_bar + 8 * mt, 1);
      ptx::mbar_init(res_bar + 8 * mt, 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync();  // peers' barriers exist before anyone multicasts into / arrives on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  unsigned long long* trc = a.trace ? a.trace + (size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (trc && threadIdx.x == 0) trc[0] = clock64();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      // box origins of the MT sub-tiles (a tile index past the end decodes to n0 >= N: every row is
      // out of bounds, TMA zero-fills it and the epilogue masks it - no special case needed)
      int n0[MT], d0[MT], h0[MT], w0[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int t = blockIdx.x * MT + mt;
        w0[mt] = (t % a.nb_w) * a.tw; t /= a.nb_w;
        h0[mt] = (t % a.nb_h) * a.th; t /= a.nb_h;
        d0[mt] = (t % a.nb_d) * a.td; t /= a.nb_d;
        n0[mt] = t * a.tn;
      }
      const uint32_t tx_bytes = (uint32_t)MT * (uint32_t)a.rows_box * 128u + (uint32_t)Cfg::kBBytes;
      for (int it = it0; it < it1; ++it) {
        const int s = (it - it0) % STAGES;
        const uint32_t ph = (uint32_t)((it - it0) / STAGES) & 1u;
        ptx::mbar_wait(bar_base + 8 * (STAGES + s), ph ^ 1u, 0);
        const uint32_t full = bar_base + 8 * s;
        const uint32_t sa = smem_base + s * Cfg::kStageBytes;
        const uint32_t sb = sa + MT * kABytes;
        ptx::mbar_expect_tx(full, tx_bytes);
        if (it >= iters_main) {
          // fused skip conv: same output box, no tap shift, its own sources and weights
          const int sc = it - iters_main;
          const bool second = sc >= a.nsk1;
          const CUtensorMap* m = second ? &maps.s[1] : &maps.s[0];
          const int c0 = (second ? sc - a.nsk1 : sc) * 64;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            ptx::tma_load_5d(sa + mt * kABytes, m, full, c0, w0[mt], h0[mt], d0[mt], n0[mt]);
          ptx::tma_load_3d(sb, &maps.sw, full, sc * 64, co0, 0);
          continue;
        }
        const int tap = it / nchunks;
        const int chunk = it - tap * nchunks;
        int kd = 0, kh = 0, kw = 0;
        int pad_d = a.pad_lo, pad_h = a.pad_lo, pad_w = a.pad_lo, wtap = tap;
        if (a.up) {
          // phase conv: low-res offset = a - 1 + i per dim (i in {0,1}); weights pre-summed per phase
          kd = tap >> 2; kh = (tap >> 1) & 1; kw = tap & 1;
          pad_d = 1 - (phase >> 2); pad_h = 1 - ((phase >> 1) & 1); pad_w = 1 - (phase & 1);
          wtap = phase * 8 + tap;
        } else if (a.ksize == 3) { kd = tap / 9; kh = (tap / 3) % 3; kw = tap % 3; }
        if (a.stride == 1) {
          const bool second = chunk >= a.nchunk1;
          const CUtensorMap* m = second ? &maps.a[1] : &maps.a[0];
          const int c0 = (second ? chunk - a.nchunk1 : chunk) * 64;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            ptx::tma_load_5d(sa + mt * kABytes + a.ashift * 128, m, full, c0, w0[mt] + kw - pad_w,
                             h0[mt] + kh - pad_h, d0[mt] + kd - pad_d, n0[mt]);
        } else {
          // input coord = 2*o + k - pad_lo  ->  parity grid (k-pad_lo)&1 at index o + floor((k-pad_lo)/2)
          const int od = kd - a.pad_lo, oh = kh - a.pad_lo, ow = kw - a.pad_lo;
          const int pd = od & 1, ph2 = oh & 1, pw = ow & 1;
          const CUtensorMap* m = &maps.a[pd * 4 + ph2 * 2 + pw];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            ptx::tma_load_5d(sa + mt * kABytes, m, full, chunk * 64, w0[mt] + ((ow - pw) >> 1),
                             h0[mt] + ((oh - ph2) >> 1), d0[mt] + ((od - pd) >> 1), n0[mt]);
        }
        if (CL == 1) {
          ptx::tma_load_3d(sb, &maps.w, full, chunk * 64, co0, wtap);
        } else {
          // this CTA fetches its 1/CL slice of the weight tile and multicasts it to the same offset in every
          // CTA of the cluster; each CTA's full barrier collects all CL slices (tx_bytes counts the whole tile)
          constexpr int kSliceRows = BN / CL;
          ptx::tma_load_3d_mcast(sb + cta_rank * (kSliceRows * 128), &maps.w, full, chunk * 64,
                                 co0 + (int)cta_rank * kSliceRows, wtap, (uint16_t)((1u << CL) - 1));
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BN);
      constexpr uint32_t dhi = ptx::umma_desc_hi_sw128();
      // descriptor low words: stage 0, then one add per stage / sub-tile / k-slice (ptx::umma_bf16_lohi) - building each
      // descriptor with shifts and masks costs this single thread ~85 cycles per MMA (conv_halo.cu header)
      const uint32_t a_lo0 = ptx::umma_desc_lo_sw128(smem_base + a.ashift * 128);
      const uint32_t b_lo0 = ptx::umma_desc_lo_sw128(smem_base + MT * kABytes);
      uint32_t s = 0, ph = 0;
      for (int it = it0; it < it1; ++it) {
        ptx::mbar_wait(bar_base + 8 * s, ph, 1);
        ptx::tc_fence_after();
        if (trc && it == it0) trc[1] = clock64();
        const uint32_t a_lo = a_lo0 + s * (uint32_t)(Cfg::kStageBytes >> 4);
        const uint32_t b_lo = b_lo0 + s * (uint32_t)(Cfg::kStageBytes >> 4);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_lohi(tmem_base + mt * BN, a_lo + (uint32_t)((mt * kABytes + k * 32) >> 4),
                                b_lo + (uint32_t)(k * 2), dhi, dhi, idesc, (it > it0 || k > 0) ? 1u : 0u);
          if (it == it1 - 1) ptx::umma_commit(tmem_full_bar + 8 * mt);  // this accumulator is complete
        }
        // frees the smem stage (in every CTA of the cluster) when the MMAs retire
        if (CL == 1) ptx::umma_commit(bar_base + 8 * (STAGES + s));
        else ptx::umma_commit_mcast(bar_base + 8 * (STAGES + s), (uint16_t)((1u << CL) - 1));
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      if (trc) trc[2] = clock64();
    }
  } else {
    // =========================== epilogue (4 warps per M sub-tile) ===========================
    const int mt = (warp - 2) >> 2;    // which accumulator / sub-tile
    const int q = warp & 3;            // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;     // accumulator row == voxel slot in the box
    int t = blockIdx.x * MT + mt;
    const int w0 = (t % a.nb_w) * a.tw; t /= a.nb_w;
    const int h0 = (t % a.nb_h) * a.th; t /= a.nb_h;
    const int d0 = (t % a.nb_d) * a.td; t /= a.nb_d;
    const int n0 = t * a.tn;
    t = row;
    const int rw = t % a.tw; t /= a.tw;
    const int rh = t % a.th; t /= a.th;
    const int rd = t % a.td; t /= a.td;
    const int rn = t;
    const int n = n0 + rn, d = d0 + rd, h = h0 + rh, w = w0 + rw;
    const bool valid = row < a.rows_box && n < a.N && d < a.Do && h < a.Ho && w < a.Wo;
    const size_t vox = (((size_t)n * a.Do + d) * a.Ho + h) * a.Wo + w;
    const size_t obase = vox * (size_t)a.Cout + (size_t)zsplit * (size_t)a.slab_elems;  // split-K: fp32 slab z
    const float* temb_row = a.temb ? a.temb + (size_t)(a.temb_stride ? n : 0) * a.temb_stride : nullptr;
    const int cpg = a.gn_stats ? a.Cout / a.gn_groups : 1;
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BN);

    ptx::mbar_wait_long(tmem_full_bar + 8 * mt, 0, 2);
    ptx::tc_fence_after();
    if (trc && lane == 0 && q == 0) trc[3 + 2 * mt] = clock64();
    if (BN >= 64 && a.tma_epilogue) {
      // ---- coalesced epilogue: the main-loop stages are dead now, so the tile is staged there in 64-channel
      // SWIZZLE_128B groups ([group][128 rows][128 B], the layout of the y / residual tensor maps) and leaves
      // through TMA stores, which also clip rows outside the tensor.  A fused residual arrives the same way
      // (TMA load into the very slots the output will occupy; each thread only touches its own row).
      constexpr int G = BN / 64;
      // the LAST accumulator's MMAs still read the final stage: nobody may overwrite smem before they retire
      if (mt != MT - 1) ptx::mbar_wait(tmem_full_bar + 8 * (MT - 1), 0, 4);
      const uint32_t stg = smem_base + (uint32_t)mt * (G * 16384);
      uint8_t* stg_gen = smem_gen + (size_t)mt * (G * 16384);
      const bool issuer = (((warp - 2) & 3) == 0) && lane == 0;
      const int gmax = min(G, (a.Cout - co0 + 63) / 64);  // groups that exist in this Cout tile (warp-uniform)
      // fused GroupNorm statistics: (sum, sumsq) per 8-channel chunk of the tile as stored (bf16), reduced in smem
      uint8_t* vrow = smem_gen + (bar_base - smem_base) + 256 + mt * 512;   // [128] row-valid flags
      float* csum = reinterpret_cast<float*>(vrow + 128);                   // [G*8][2]
      const bool do_stats = a.chunk_stats != nullptr;
      if (do_stats) {
        vrow[row] = valid ? 1 : 0;
        if (row < G * 16) csum[row] = 0.f;
      }
      if (a.residual) {
        if (issuer) {
          ptx::mbar_expect_tx(res_bar + 8 * mt, (uint32_t)gmax * (uint32_t)a.rows_box * 128u);
          for (int g = 0; g < gmax; ++g)
            ptx::tma_load_5d(stg + g * 16384, &maps.r, res_bar + 8 * mt, co0 + g * 64, w0, h0, d0, n0);
        }
        ptx::mbar_wait(res_bar + 8 * mt, 0, 3);
      }
#pragma unroll 1
      for (int g = 0; g < gmax; ++g) {
        uint8_t* rowp = stg_gen + (size_t)g * 16384 + (size_t)row * 128;
        // all 64 columns of the group in flight before the single wait (the TMEM round trip is ~230 cycles)
        uint32_t rr[4][16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) ptx::tmem_ld16(tacc + (uint32_t)(g * 64 + cc * 16), rr[cc]);
        ptx::tmem_ld_wait();
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
        asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");  // the 4 warps of this accumulator
        if (issuer) {
          ptx::tma_store_5d(&maps.y[phase], stg + g * 16384, co0 + g * 64, w0, h0, d0, n0);
          ptx::tma_store_commit();
        }
        if (do_stats) {
          // thread (slice = row>>3, chunk = row&7) sums 8 rows x 8 channels of the staged tile; lanes that share a
          // chunk combine by shuffles, warps by shared-memory atomics
          const int cch = row & 7, slice = row >> 3;
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = slice * 8 + i;
            if (vrow[r]) {
              const uint4 u = *reinterpret_cast<const uint4*>(stg_gen + (size_t)g * 16384 + (size_t)r * 128 +
                                                              ((cch ^ (r & 7)) << 4));
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
        asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");
        if (row < gmax * 16 && n0 < a.N) {
          const int k = row >> 1, which = row & 1;   // chunk k of this Cout tile, 0 = sum / 1 = sumsq
          atomicAdd(a.chunk_stats + ((size_t)n0 * (a.Cout >> 3) + (co0 >> 3) + k) * 2 + which, (double)csum[row]);
        }
      }
      if (issuer) ptx::tma_store_wait_read0();  // smem must outlive the bulk reads
    } else if (BN >= 64 && a.slab_tma) {
      // ---- split-K partial sums (fp32, no bias / temb / residual): one thread owns one accumulator row, so direct
      // stores hand the LSU 32 different lines per instruction (~16 K cycles for a 256 x 256 tile).  Instead the tile is
      // staged in 32-channel SWIZZLE_128B groups ([128 rows][128 B]) in the dead main-loop stages and leaves through TMA
      // stores into slab `zsplit` (the fp32 map folds the slab index into the sample coordinate).
      constexpr int G32 = BN >= 64 ? BN / 32 : 2;   // (the BN = 16 instance never takes this branch)
      constexpr int kAvail = (STAGES * Cfg::kStageBytes / 16384) / MT;
      constexpr int SLOTS = kAvail < G32 ? kAvail : G32;
      static_assert(SLOTS >= 2, "slab staging slots");
      if (mt != MT - 1) ptx::mbar_wait(tmem_full_bar + 8 * (MT - 1), 0, 4);   // the last MMAs still read the final stage
      const uint32_t stg = smem_base + (uint32_t)mt * (SLOTS * 16384);
      uint8_t* stg_gen = smem_gen + (size_t)mt * (SLOTS * 16384);
      const bool issuer = (((warp - 2) & 3) == 0) && lane == 0;
      const int gmax = min(G32, (a.Cout - co0 + 31) / 32);
#pragma unroll 1
      for (int g = 0; g < gmax; ++g) {
        const int slot = g % SLOTS;
        uint32_t rr[2][16];
        ptx::tmem_ld16(tacc + (uint32_t)(g * 32), rr[0]);
        ptx::tmem_ld16(tacc + (uint32_t)(g * 32 + 16), rr[1]);
        if (g >= SLOTS) {   // slot reuse: the store of group g - SLOTS must have read the buffer out
          if (issuer) ptx::tma_store_wait_read<SLOTS - 1>();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");
        }
        ptx::tmem_ld_wait();
        uint8_t* rowp = stg_gen + (size_t)slot * 16384 + (size_t)row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) =
              make_uint4(rr[c >> 2][(c & 3) * 4], rr[c >> 2][(c & 3) * 4 + 1], rr[c >> 2][(c & 3) * 4 + 2],
                         rr[c >> 2][(c & 3) * 4 + 3]);
        ptx::fence_proxy_async();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + mt) : "memory");
        // (a sub-tile past the end - odd tile count with MT = 2 - decodes to n0 >= N: in the bf16 epilogue the TMA unit
        // clips it, but here the slab index is folded into the sample coordinate, so it would land in slab z + 1)
        if (issuer && n0 < a.N) {
          ptx::tma_store_5d(&maps.y[0], stg + slot * 16384, co0 + g * 32, w0, h0, d0, zsplit * a.N + n0);
          ptx::tma_store_commit();
        }
      }
      if (issuer) ptx::tma_store_wait_read0();
    } else
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t r[16];
      ptx::tmem_ld16(tacc + (uint32_t)c, r);
      ptx::tmem_ld_wait();
      const int co = co0 + c;
      if (co >= a.Cout) continue;  // warp-uniform
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      const int ncol = min(16, a.Cout - co);  // Cout % 8 == 0 -> 8 or 16
      if (a.bias) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < ncol) v[j] += __ldg(a.bias + co + j);
      }
      if (temb_row && valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < ncol) v[j] += __ldg(temb_row + co + j);
      }
      if (a.residual && valid) {
        if (a.out_f32) {
          const float* rp = reinterpret_cast<const float*>(a.residual) + obase + co;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (j < ncol) {
              float4 f = *reinterpret_cast<const float4*>(rp + j);
              v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
            }
        } else {
          const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(a.residual) + obase + co;
#pragma unroll
          for (int j = 0; j < 16; j += 8)
            if (j < ncol) {
              float f[8];
              Vec8<__nv_bfloat16>::load(rp + j, f);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[j + e] += f[e];
            }
        }
      }
      if (valid) {
        if (a.out_f32) {
          float* yp = reinterpret_cast<float*>(a.y) + obase + co;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (j < ncol) *reinterpret_cast<float4*>(yp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y) + obase + co;
#pragma unroll
          for (int j = 0; j < 16; j += 8)
            if (j < ncol) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = v[j + e];
              Vec8<__nv_bfloat16>::store(yp + j, f);
            }
        }
      }
      if (a.gn_stats) {
        // GroupNorm partials of the tensor just produced: (sum, sumsq) per (sample, group).
        // Statistics use the values as stored (bf16-rounded when the output is bf16).
        // cpg in {2,4,8,16,32,...}: a 16-column chunk covers 16/cpg groups or part of one.
        const int gmask = (cpg >= 16 ? 16 : cpg) - 1;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < ncol && valid) {
            const float x = a.out_f32 ? v[j] : __bfloat162float(__float2bfloat16_rn(v[j]));
            s1 += x; s2 += x * x;
          }
          if (((j + 1) & gmask) == 0 && j < ncol) {  // warp-uniform: end of a group span
            const int g = (co + j) / cpg;
            // rows of one warp belong to different samples only when tn > 1
            if (a.tn == 1) {
              const float t1 = warp_sum(s1), t2 = warp_sum(s2);
              if (lane == 0 && n0 < a.N) {
                atomicAdd(a.gn_stats + ((size_t)n0 * a.gn_groups + g) * 2, (double)t1);
                atomicAdd(a.gn_stats + ((size_t)n0 * a.gn_groups + g) * 2 + 1, (double)t2);
              }
            } else if (valid) {
              atomicAdd(a.gn_stats + ((size_t)n * a.gn_groups + g) * 2, (double)s1);
              atomicAdd(a.gn_stats + ((size_t)n * a.gn_groups + g) * 2 + 1, (double)s2);
            }
            s1 = 0.f; s2 = 0.f;
          }
        }
      }
    }
    if (trc && lane == 0 && q == 0) trc[4 + 2 * mt] = clock64();
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
  if (trc && threadIdx.x == 0) trc[7] = clock64();
  if (CL > 1) ptx::cluster_sync();  // do not exit while a peer can still arrive on our barriers
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// 5-D bf16 activation map: dims (C, W, H, D, N) innermost first
static int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int D, int N, uint64_t sW,
                        uint64_t sH, uint64_t sD, uint64_t sN, int tw, int th, int td, int tn) {
  auto enc = get_encode_fn();
  if (!enc) return fail(B200LDM_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {sW * 2, sH * 2, sD * 2, sN * 2};  // bytes
  cuuint32_t box[5] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)td, (cuuint32_t)tn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B200LDM_BAD_SHAPE,
                "cuTensorMapEncodeTiled(act) failed: %d dims=(%d,%d,%d,%d,%d) box=(64,%d,%d,%d,%d)", (int)r, C, W,
                H, D, N, tw, th, td, tn);
  return B200LDM_OK;
}

// fp32 split-K slabs [splits][N][Do][Ho][Wo][Cout] as ONE 5-D map (C, W, H, D, splits * N): 32-channel (128-byte) boxes
static int make_slab_map(CUtensorMap* m, const void* base, int C, int W, int H, int D, int NS, int tw, int th, int td,
                         int tn) {
  auto enc = get_encode_fn();
  if (!enc) return fail(B200LDM_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)NS};
  const uint64_t sW = (uint64_t)C * 4, sH = sW * W, sD = sH * H, sN = sD * D;
  cuuint64_t strides[4] = {sW, sH, sD, sN};  // bytes
  cuuint32_t box[5] = {32, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)td, (cuuint32_t)tn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B200LDM_BAD_SHAPE, "cuTensorMapEncodeTiled(slab) failed: %d dims=(%d,%d,%d,%d,%d) box=(32,%d,%d,%d,%d)",
                (int)r, C, W, H, D, NS, tw, th, td, tn);
  return B200LDM_OK;
}

static int make_weight_map(CUtensorMap* m, const void* base, int Cin, int Cout_rows, int taps, int bn) {
  auto enc = get_encode_fn();
  if (!enc) return fail(B200LDM_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout_rows, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * Cout_rows * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)bn, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B200LDM_BAD_SHAPE, "cuTensorMapEncodeTiled(weight) failed: %d Cin=%d Cout=%d taps=%d bn=%d", (int)r,
                Cin, Cout_rows, taps, bn);
  return B200LDM_OK;
}

// Pick the voxel box (tn,td,th,tw), product <= 128, that covers the output with the fewest tiles.
void choose_conv_tile(int N, int Do, int Ho, int Wo, int* tn, int* td, int* th, int* tw) {
  long best_tiles = -1;
  int b[4] = {1, 1, 1, 1};
  for (int w = 1; w <= Wo && w <= 128; ++w) {
    for (int h = 1; h <= Ho && w * h <= 128; ++h) {
      for (int d = 1; d <= Do && w * h * d <= 128; ++d) {
        int nmax = (w == Wo && h == Ho && d == Do) ? 128 / (w * h * d) : 1;
        if (nmax > N) nmax = N;
        for (int n = 1; n <= nmax; ++n) {
          long tiles = (long)ceil_div(N, n) * ceil_div(Do, d) * ceil_div(Ho, h) * ceil_div(Wo, w);
          bool better = best_tiles < 0 || tiles < best_tiles ||
                        (tiles == best_tiles && (w > b[3] || (w == b[3] && h > b[2])));
          if (better) { best_tiles = tiles; b[0] = n; b[1] = d; b[2] = h; b[3] = w; }
        }
      }
    }
  }
  *tn = b[0]; *td = b[1]; *th = b[2]; *tw = b[3];
}

template <int BN, int MT, int CL>
static int launch_conv_tc(const ConvTcMaps& maps, const ConvTcArgs& args, int m_tiles, int n_tiles, int splits,
                          cudaStream_t st) {
  using Cfg = ConvTcCfg<BN, MT>;
  B2_SET_SMEM_ONCE(Cfg::kSmemBytes, conv3d_tc_kernel<BN, MT, CL>);
  const unsigned gx = (unsigned)(ceil_div(ceil_div(m_tiles, MT), CL) * CL);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(gx, (unsigned)n_tiles, (unsigned)splits);
  cfg.blockDim = dim3(Cfg::kThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  B2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3d_tc_kernel<BN, MT, CL>, maps, args));
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

// split-K finalize: y = bf16( sum_z ws[z] + bias + temb[n] + residual ), 8 channels per thread
__global__ void splitk_finalize_kernel(const float* __restrict__ ws, int S, long long slab, long long M, int Cout,
                                       long long vox_per_sample, const float* __restrict__ bias,
                                       const float* __restrict__ temb, int temb_stride,
                                       const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ y) {
  const long long vecs = M * (Cout >> 3);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < vecs; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / (Cout >> 3);
    const int c = (int)(i - m * (Cout >> 3)) * 8;
    const size_t off = (size_t)m * Cout + c;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    for (int z = 0; z < S; ++z) {
      float f[8];
      Vec8<float>::load(ws + (size_t)z * slab + off, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += f[e];
    }
    if (bias) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += __ldg(bias + c + e);
    }
    if (temb) {
      const long long n = temb_stride ? m / vox_per_sample : 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += __ldg(temb + n * temb_stride + c + e);
    }
    if (residual) {
      float f[8];
      Vec8<__nv_bfloat16>::load(residual + off, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += f[e];
    }
    Vec8<__nv_bfloat16>::store(y + off, v);
  }
}

// The same finalize that also accumulates the GroupNorm chunk sums of the tensor it writes - (sum, sumsq) per 8 channels
// and sample of the bf16-rounded values, like the TMA epilogues of the unsplit kernels - so a split-K conv needs no
// statistics pass either.  grid = (blocks per sample, N): a block owns a contiguous row range of ONE sample, sums into
// shared-memory fp32 accumulators and issues one fp64 atomic pair per chunk.
__global__ void splitk_finalize_stats_kernel(const float* __restrict__ ws, int S, long long slab, int Cout,
                                             long long vox_per_sample, const float* __restrict__ bias,
                                             const float* __restrict__ temb, int temb_stride,
                                             const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ y,
                                             double* __restrict__ chunk_stats) {
  extern __shared__ float fs_csum[];   // [Cout/8][2]
  const int cpr = Cout >> 3;           // 8-channel chunks per row
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * cpr; i += blockDim.x) fs_csum[i] = 0.f;
  __syncthreads();
  const long long r0 = vox_per_sample * blockIdx.x / gridDim.x, r1 = vox_per_sample * (blockIdx.x + 1) / gridDim.x;
  const long long v0 = r0 * cpr, v1 = r1 * cpr;
  for (long long i = v0 + threadIdx.x; i < v1; i += blockDim.x) {
    const long long r = i / cpr;
    const int k = (int)(i - r * cpr), c = k * 8;
    const size_t off = (size_t)(n * vox_per_sample + r) * Cout + c;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    for (int z = 0; z < S; ++z) {
      float f[8];
      Vec8<float>::load(ws + (size_t)z * slab + off, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += f[e];
    }
    if (bias) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += __ldg(bias + c + e);
    }
    if (temb) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += __ldg(temb + (size_t)(temb_stride ? n : 0) * temb_stride + c + e);
    }
    if (residual) {
      float f[8];
      Vec8<__nv_bfloat16>::load(residual + off, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += f[e];
    }
    Vec8<__nv_bfloat16>::store(y + off, v);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x = __bfloat162float(__float2bfloat16_rn(v[e]));   // statistics of the tensor as stored
      s1 += x;
      s2 = fmaf(x, x, s2);
    }
    atomicAdd(&fs_csum[2 * k], s1);
    atomicAdd(&fs_csum[2 * k + 1], s2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * cpr; i += blockDim.x)
    atomicAdd(chunk_stats + (size_t)n * 2 * cpr + i, (double)fs_csum[i]);
}

static int launch_finalize(const b200ldm_conv3d_params* p, int splits, long long M_total, cudaStream_t st) {
  const long long vps = (long long)p->Do * p->Ho * p->Wo;
  if (p->chunk_stats) {
    int bps = ceil_div(2 * sm_count(), p->N);
    if (bps > vps / 4) bps = (int)(vps / 4);
    if (bps < 1) bps = 1;
    splitk_finalize_stats_kernel<<<dim3((unsigned)bps, (unsigned)p->N), 256, (size_t)(p->Cout / 8) * 2 * sizeof(float), st>>>(
        reinterpret_cast<const float*>(p->workspace), splits, M_total * p->Cout, p->Cout, vps, p->bias, p->temb,
        p->temb_stride, reinterpret_cast<const __nv_bfloat16*>(p->residual), reinterpret_cast<__nv_bfloat16*>(p->y),
        p->chunk_stats);
  } else {
    const long long vecs = M_total * (p->Cout / 8);
    const unsigned blocks = (unsigned)std::min<long long>((vecs + 255) / 256, (long long)sm_count() * 16);
    splitk_finalize_kernel<<<blocks, 256, 0, st>>>(
        reinterpret_cast<const float*>(p->workspace), splits, M_total * p->Cout, M_total, p->Cout, vps, p->bias, p->temb,
        p->temb_stride, reinterpret_cast<const __nv_bfloat16*>(p->residual), reinterpret_cast<__nv_bfloat16*>(p->y));
  }
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

static unsigned long long* g_conv_trace = nullptr;
void set_conv_trace(void* p) { g_conv_trace = reinterpret_cast<unsigned long long*>(p); }

// test / tuning hooks: B200LDM_CONV_MT=1|2 pins the M sub-tiles per CTA, B200LDM_CONV_CL=1|2 the cluster size
static const int g_force_mt = [] {
  const char* e = getenv("B200LDM_CONV_MT");
  return e ? atoi(e) : 0;
}();
static const int g_legacy_epilogue = [] {
  const char* e = getenv("B200LDM_CONV_LEGACY_EPILOGUE");
  return e ? atoi(e) : 0;
}();
// B200LDM_CONV_SLAB_TMA=0: split-K slabs by direct per-row stores (the round-1 epilogue) instead of smem + TMA stores
static const int g_slab_tma = [] {
  const char* e = getenv("B200LDM_CONV_SLAB_TMA");
  return e ? atoi(e) : 1;
}();
static const int g_force_cl = [] {
  const char* e = getenv("B200LDM_CONV_CL");
  return e ? atoi(e) : 0;
}();

// B200LDM_CONV_KERNEL pins the kernel wherever it supports the problem: 0 = one CTA per tile (conv3d_tc_kernel),
// 1 = persistent CTA pairs (conv_tc2.cu), 2 = persistent single CTAs (conv_tcp.cu), 3 = persistent W-halo kernel
// (conv_halo.cu); unset: measured / cost model
static int g_tc2 = [] {
  const char* e = getenv("B200LDM_CONV_KERNEL");
  return e ? atoi(e) : -1;
}();
// B200LDM_CONV_HALO=0 keeps the W-halo kernel (conv_halo.cu, kernel 3) out of the measured candidates.  It is the
// fastest kernel for 3x3x3 stride-1 convs with Cout tiles of 64 / 128 (64 -> 64 at 8 x 80^3: 0.82 ms against 1.25 ms
// for conv_tcp.cu, 128 -> 128 at 8 x 40^3: 0.30 against 0.38 ms).
static const int g_halo = [] {
  const char* e = getenv("B200LDM_CONV_HALO");
  return e ? atoi(e) : 1;
}();
// B200LDM_CONV_STREAMK=0 keeps the stream-K kernel (conv_sk.cu, kernel 5) out of the measured candidates
static const int g_stream_k = [] {
  const char* e = getenv("B200LDM_CONV_STREAMK");
  return e ? atoi(e) : 1;
}();
void set_conv_kernel(int k) { g_tc2 = k; }   // tests: exercise every kernel in one process (b200ldm_debug_set_conv_kernel)

static int g_force_splits = [] {
  const char* e = getenv("B200LDM_CONV_SPLITS");
  return e ? atoi(e) : 0;
}();
void set_conv_splits(int s) { g_force_splits = s; }
// tests (b200ldm_debug_set_conv_choice): run exactly one candidate of the tuner - a choice code (kernel | splits << 4 |
// mt << 12 | cl << 16 | Cout-tile code << 20) - for every problem; -1 = off
static int g_force_choice = -1;
void set_conv_choice(int c) { g_force_choice = c; }   // tests (b200ldm_debug_set_conv_splits): force a split-K factor

struct ConvPlan {
  int bn, n_tiles, tn, td, th, tw, nb_n, nb_d, nb_h, nb_w, m_tiles, mt, cl, splits;
  int bn0;     // the default Cout tile (the widest that divides Cout); bn < bn0 only by a measured choice
  int tc2;     // kernel: 0 one CTA per tile, 1 persistent CTA pairs (conv_tc2.cu), 2 persistent single CTAs (conv_tcp.cu),
               // 3 persistent W-halo kernel (conv_halo.cu), 4 conv_tcp.cu with two epilogue warp sets
  int tc2_ok;  // the problem is eligible for the persistent kernels
  int halo_ok; // ... and for the W-halo kernel
  int sk_ok;   // ... and for the stream-K kernel (conv_sk.cu, kernel 5)
  int split_ok, ctas1, iters;   // split-K allowed; CTAs of the one-CTA-per-tile kernel at splits = 1 (mt = 1); k-iterations
};

// Which of the two tensor-core kernels runs a problem is MEASURED once per problem signature (first un-captured call:
// both candidates are timed on the caller's own buffers with CUDA events) and remembered here; until then - and
// always with B200LDM_CONV_AUTOTUNE=0 or while the stream is being captured - the cost model below decides.
struct ConvKey {
  int v[20];
  bool operator<(const ConvKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
static ConvKey conv_key(const b200ldm_conv3d_params* p) {
  ConvKey k{{p->N, p->D, p->H, p->W, p->C1, p->C2, p->Cout, p->Do, p->Ho, p->Wo, p->ksize, p->stride, p->pad_lo,
             p->upsample2x, p->out_f32, p->gn_stats != nullptr, p->residual != nullptr, p->temb != nullptr,
             p->skip_x ? p->skip_C1 : 0, p->skip_x ? p->skip_C2 : 0}};
  return k;
}
static std::mutex g_choice_mu;
static std::map<ConvKey, int> g_choice;   // 0: one CTA per tile (conv3d_tc_kernel), 1: persistent CTA pairs (tc2)
static const int g_autotune = [] {
  const char* e = getenv("B200LDM_CONV_AUTOTUNE");
  return e ? atoi(e) : 1;
}();
// B200LDM_CONV_TUNE_FILE=<path>: measured choices are appended to / preloaded from this text file (21 integers per
// line), so a later process - a profiler run, a serving replica - starts with them and never times anything.
static const char* g_tune_file = [] {
  const char* path = getenv("B200LDM_CONV_TUNE_FILE");
  if (path && *path) {
    if (FILE* f = fopen(path, "r")) {
      ConvKey k;
      int choice;
      for (;;) {
        bool ok = true;
        for (int i = 0; i < 20 && ok; ++i) ok = fscanf(f, "%d", &k.v[i]) == 1;
        if (!ok || fscanf(f, "%d", &choice) != 1) break;
        g_choice[k] = choice;
      }
      fclose(f);
    }
  }
  return (path && *path) ? path : nullptr;
}();
static void remember_choice(const ConvKey& key, int choice) {
  std::lock_guard<std::mutex> lk(g_choice_mu);
  g_choice[key] = choice;
  if (g_tune_file) {
    if (FILE* f = fopen(g_tune_file, "a")) {
      for (int i = 0; i < 20; ++i) fprintf(f, "%d ", key.v[i]);
      fprintf(f, "%d\n", choice);
      fclose(f);
    }
  }
}

// Tile / cluster / split-K decisions (shared by the launch and by b200ldm_conv3d_workspace_bytes).
// A choice is kernel | splits << 4 | mt << 12 | cl << 16 (splits / mt / cluster size only for kernel 0; 0 = the cost
// model's value).
// force_tc2: -1 = remembered measurement, else cost model; >= 0 = that choice (tuning runs).
static int chunk_width_of(const b200ldm_conv3d_params* p) { return p->chunk_width ? p->chunk_width : 8; }
static ConvPlan plan_conv(const b200ldm_conv3d_params* p, int force_tc2 = -1) {
  ConvPlan pl{};
  const int SMS = sm_count();
  if (p->Cout % 256 == 0) pl.bn = 256;
  else if (p->Cout % 128 == 0) pl.bn = 128;
  else if (p->Cout % 64 == 0) pl.bn = 64;
  else if (p->Cout <= 16) pl.bn = 16;
  else if (p->Cout <= 64) pl.bn = 64;
  else if (p->Cout <= 128) pl.bn = 128;
  else pl.bn = 256;
  pl.bn0 = pl.bn;
  // the choice this plan follows: forced (tuning runs), remembered (measured), or none yet (cost model below)
  int choice_in = -1;
  if (force_tc2 < 0 && g_force_choice >= 0) force_tc2 = g_force_choice;
  if (force_tc2 >= 0) choice_in = force_tc2;
  else if (g_tc2 < 0) {
    std::lock_guard<std::mutex> lk(g_choice_mu);
    auto it = g_choice.find(conv_key(p));
    if (it != g_choice.end()) choice_in = it->second;
  }
  if (choice_in >= 0) {
    // a narrower Cout tile (bits 20..21: 1 = 64, 2 = 128) gives launches that cannot fill the chip more CTAs with
    // shorter epilogues: the 1x1x1 projections of the 5^3 / 10^3 levels
    const int code = (choice_in >> 20) & 3, alt = code == 1 ? 64 : code == 2 ? 128 : 0;
    if (alt && alt < pl.bn && p->Cout % alt == 0) pl.bn = alt;
  }
  pl.n_tiles = ceil_div(p->Cout, pl.bn);
  // M tiles enumerate output voxels; with the fused upsample they enumerate LOW-RES voxels (x 8 phases in grid.z)
  const int eD = p->upsample2x ? p->D : p->Do, eH = p->upsample2x ? p->H : p->Ho, eW = p->upsample2x ? p->W : p->Wo;
  choose_conv_tile(p->N, eD, eH, eW, &pl.tn, &pl.td, &pl.th, &pl.tw);
  pl.nb_n = ceil_div(p->N, pl.tn); pl.nb_d = ceil_div(eD, pl.td);
  pl.nb_h = ceil_div(eH, pl.th); pl.nb_w = ceil_div(eW, pl.tw);
  pl.m_tiles = pl.nb_n * pl.nb_d * pl.nb_h * pl.nb_w;
  // (mt, cl, splits) by a small cost model calibrated on the clock64 traces (scripts/conv_trace.py):
  //   time = rounds(ctas / 148) x (k-iterations x cycles/iteration + fixed CTA overhead)
  //   cycles/iteration = max(MMA issue time, bytes staged per iteration / ~55 B/clk/SM of L2->SMEM bandwidth)
  const int taps_eff = p->upsample2x ? 8 : p->ksize * p->ksize * p->ksize;
  const int iters = taps_eff * ((p->C1 + p->C2) / 64) + (p->skip_x ? (p->skip_C1 + p->skip_C2) / 64 : 0);
  const int zmul = p->upsample2x ? 8 : 1;
  const bool split_ok = !p->out_f32 && !p->gn_stats && !p->upsample2x && p->Cout % 8 == 0;
  double best = 1e30;
  pl.mt = 1; pl.cl = 1; pl.splits = 1;
  for (int mt = 1; mt <= 2; ++mt) {
    if ((g_force_mt == 1 || g_force_mt == 2) && mt != g_force_mt) continue;
    for (int cl = 1; cl <= 2; ++cl) {
      if (cl == 2 && (pl.bn < 128 || p->skip_x)) continue;
      if ((g_force_cl == 1 || g_force_cl == 2) && cl != g_force_cl && !(cl == 1 && pl.bn < 128)) continue;
      // An SM ingests at most ~64-72 B/clk from L2 (measured: 128x256 tiles run at 680 instead of 512 cycles per
      // k-iteration even with 96 CTAs on the chip; 256x256 tiles need exactly 64 B/clk and run at the MMA rate).
      // Multicast does not lower the bytes that enter each SM, so it does not change this bound.
      const double mma = mt * 4.0 * (128.0 * pl.bn / 256.0);
      const double load = (mt * 16384.0 + pl.bn * 128.0) / 72.0;
      const double cyc_iter = (mma > load ? mma : load) + 8.0;
      const double overhead = 4000.0 + 5000.0 * mt;   // prologue + epilogue (10K cycles for two accumulators)
      const long long ctas = (long long)ceil_div(ceil_div(pl.m_tiles, mt), cl) * cl * pl.n_tiles * zmul;
      for (int sp = 1; sp <= 12; ++sp) {
        if (sp > 1 && (!split_ok || cl != 1 || iters / sp < 16 || ctas >= SMS)) break;
        const double rounds = (double)ceil_div(ctas * sp, SMS);
        double t = rounds * (ceil_div(iters, sp) * cyc_iter + overhead);
        if (sp > 1) t += 6000.0;  // finalize launch + slab traffic
        if (cl == 2) t *= 0.99;   // multicast: slightly less L2 traffic chip-wide
        if (t < best * 0.98) { best = t; pl.mt = mt; pl.cl = cl; pl.splits = sp; }
      }
    }
  }
  {
    // persistent CTA-pair kernel (conv_tc2.cu): rounds of 74 pair-tiles, per-tile cost = max(main loop, epilogue) since
    // the two overlap.  Measured (scripts/conv_trace2.py): a cta_group::2 k-iteration takes ~600 cycles against 512
    // for two single-CTA ones, ~745 with the loads (44 B/clk/SM), and the 4-warp epilogue ~2450 cycles per 64 channels.
    const bool tma_epi = !p->out_f32 && p->Cout % 64 == 0 && !p->gn_stats && pl.bn >= 64 && g_legacy_epilogue == 0;
    pl.tc2_ok = tma_epi ? 1 : 0;
    pl.split_ok = split_ok ? 1 : 0;
    pl.ctas1 = pl.m_tiles * pl.n_tiles * zmul;
    pl.iters = iters;
    pl.tc2 = 0;
    int choice = 0;
    if (tma_epi && g_tc2 != 0) {
      // cost-model fallback (no measurement yet / tuning disabled): persistent single-CTA kernel, per-tile cost =
      // max(main loop, epilogue) since they overlap
      const int pmt = conv_tcp_mt(pl.bn);
      const double mma = pmt * 4.0 * (128.0 * pl.bn / 256.0);
      const double load = (pmt * 16384.0 + pl.bn * 128.0) / 72.0;
      const double main_loop = iters * ((mma > load ? mma : load) + 8.0);
      const double epi = 2000.0 * (pl.bn / 64);
      const long long work = (long long)ceil_div(pl.m_tiles, pmt) * pl.n_tiles * zmul;
      const double t2 = ceil_div(work, SMS) * (main_loop > epi ? main_loop : epi) + 3000.0 + epi;
      choice = (g_tc2 > 0) ? g_tc2 : (t2 < best ? 2 : 0);
    }
    if (choice_in >= 0) choice = choice_in;
    pl.tc2 = tma_epi ? (choice & 15) : 0;
    // kernel 3 (conv_halo.cu): its own tile geometry - one sample, 16 lines of 8 voxels
    pl.halo_ok = (tma_epi && pl.bn <= 128 && p->ksize == 3 && p->stride == 1 && p->pad_lo == 1 && !p->upsample2x &&
                  p->Wo >= 4) ? 1 : 0;
    pl.sk_ok = (tma_epi && pl.bn0 == 256 && !p->upsample2x && chunk_width_of(p) == 8 && g_stream_k) ? 1 : 0;
    if (pl.tc2 == 5 && !(pl.sk_ok && pl.bn == 256)) pl.tc2 = 0;
    if (pl.tc2 == 3 && !pl.halo_ok) pl.tc2 = 0;
    if (pl.tc2 == 4 && pl.bn != 256) pl.tc2 = 2;   // the wide-epilogue variant exists for the 128 x 256 tile only
    if (pl.tc2 == 3) {
      pl.tn = 1; pl.tw = 8;
      choose_halo_tile(p->Do, p->Ho, &pl.td, &pl.th);
      pl.nb_n = p->N; pl.nb_d = ceil_div(p->Do, pl.td); pl.nb_h = ceil_div(p->Ho, pl.th); pl.nb_w = ceil_div(p->Wo, pl.tw);
      pl.m_tiles = pl.nb_n * pl.nb_d * pl.nb_h * pl.nb_w;
    }
    if (pl.tc2) { pl.mt = 1; pl.cl = 1; pl.splits = 1; }
    else {
      const int sp = (choice >> 4) & 255, mt = (choice >> 12) & 15, cl = (choice >> 16) & 15;
      if (mt == 1 || mt == 2) pl.mt = mt;
      // (every slab needs at least one k-iteration: an empty slab would never complete its accumulator barrier)
      if (sp >= 1 && (sp == 1 || (split_ok && iters / sp >= 1))) { pl.splits = sp; pl.cl = sp > 1 ? 1 : pl.cl; }
      if ((cl == 1 || cl == 2) && pl.splits == 1 && (cl == 1 || (pl.bn >= 128 && !p->skip_x))) pl.cl = cl;
    }
  }
  if (!pl.tc2 && g_force_splits >= 1 && g_force_splits <= 12 && split_ok && iters / g_force_splits >= 1)
    pl.splits = g_force_splits;
  if (pl.splits > 1) pl.cl = 1;
  return pl;
}

// 1 when this exact problem runs the TMA epilogue with one sample per tile and no split-K, i.e. can emit chunk stats
// chunk sums finer than 8 channels exist in the persistent kernels (conv_tcp.cu, conv_halo.cu) with Cout tiles <= 128
static bool plan_emits_width(const ConvPlan& pl, int cw) {
  return cw == 8 || ((cw == 2 || cw == 4) && pl.bn <= 128 && (pl.tc2 == 2 || pl.tc2 == 3));
}
// a split-K plan emits the (8-channel) chunk sums from its finalize pass - unless that pass is left to the caller
static bool split_emits_stats(const b200ldm_conv3d_params* p) {
  return chunk_width_of(p) == 8 && !p->defer_finalize && !p->out_f32 && p->Cout % 64 == 0 && p->Cout <= 8192;
}
int conv3d_tcgen05_emits_chunk_stats(const b200ldm_conv3d_params* p) {
  if (p->dtype != B200LDM_BF16 || p->C1 % 64 || p->C2 % 64 || p->Cout % 64 || p->out_f32 || p->gn_stats) return 0;
  if (g_legacy_epilogue) return 0;
  const ConvPlan pl = plan_conv(p);
  if (pl.splits > 1) return split_emits_stats(p) ? 1 : 0;   // from the finalize kernel (splitk_finalize_stats_kernel)
  return (pl.tn == 1 && pl.bn >= 64 && plan_emits_width(pl, chunk_width_of(p))) ? 1 : 0;
}

// A measurement is pending for this problem: tuning is on, nothing is pinned, the problem has more than one candidate
// (persistent kernels and / or split-K factors) and no choice has been remembered yet.
static bool tune_pending(const b200ldm_conv3d_params* p, const ConvPlan& pl) {
  const bool splittable = pl.split_ok && pl.ctas1 < sm_count() && pl.iters >= 16;
  if (!g_autotune || g_tc2 >= 0 || g_force_choice >= 0 || !(pl.tc2_ok || splittable)) return false;
  std::lock_guard<std::mutex> lk(g_choice_mu);
  return g_choice.count(conv_key(p)) == 0;
}

static bool conv3d_tcgen05_eligible(const b200ldm_conv3d_params* p) {
  return p->dtype == B200LDM_BF16 && p->C1 > 0 && p->C1 % 64 == 0 && p->C2 >= 0 && p->C2 % 64 == 0 && p->Cout > 0 &&
         p->Cout % 8 == 0 && (p->ksize == 1 || p->ksize == 3) && (p->stride == 1 || p->stride == 2);
}

int conv3d_tcgen05_tune_pending(const b200ldm_conv3d_params* p) {
  if (!conv3d_tcgen05_eligible(p)) return 0;
  return tune_pending(p, plan_conv(p)) ? 1 : 0;
}

int conv3d_tcgen05_splits(const b200ldm_conv3d_params* p) {
  if (!conv3d_tcgen05_eligible(p)) return 1;
  return plan_conv(p).splits;
}

int conv3d_splitk_finalize(const float* slabs, int splits, long long M, int Cout, long long vox_per_sample,
                           const float* bias, const float* temb, int temb_stride, const void* residual, void* y,
                           cudaStream_t st) {
  B2_REQUIRE(slabs && y && splits >= 1 && M > 0 && Cout > 0 && Cout % 8 == 0 && vox_per_sample > 0, B200LDM_BAD_ARG,
             "conv3d_finalize: bad arguments");
  B2_REQUIRE(aligned16(slabs) && aligned16(y) && aligned16(residual), B200LDM_BAD_ALIGN,
             "conv3d_finalize: pointers must be 16-byte aligned");
  const long long vecs = M * (Cout / 8);
  const unsigned blocks = (unsigned)std::min<long long>((vecs + 255) / 256, (long long)sm_count() * 16);
  splitk_finalize_kernel<<<blocks, 256, 0, st>>>(slabs, splits, M * Cout, M, Cout, vox_per_sample, bias, temb, temb_stride,
                                                 reinterpret_cast<const __nv_bfloat16*>(residual),
                                                 reinterpret_cast<__nv_bfloat16*>(y));
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

long long conv3d_tcgen05_workspace_bytes(const b200ldm_conv3d_params* p) {
  if (p->dtype != B200LDM_BF16 || p->C1 % 64 || p->C2 % 64 || p->Cout % 8) return 0;
  const ConvPlan pl = plan_conv(p);
  if (pl.tc2 == 5) return conv_sk_workspace_bytes(conv_sk_grid(pl.m_tiles, pl.n_tiles, pl.iters));
  int splits = pl.splits;
  const int SMS = sm_count();
  long long sk = 0;
  const bool pending = tune_pending(p, pl);
  if (pending && pl.sk_ok) sk = conv_sk_workspace_bytes(conv_sk_grid(pl.m_tiles, pl.n_tiles, pl.iters));   // a candidate
  if (pl.split_ok && pl.ctas1 < SMS && pl.iters >= 16 && pending) {
    // b200ldm_conv3d_tune measures split factors up to 16 (those that keep >= 8 k-iterations per CTA)
    for (int sp = 16; sp > splits; --sp)
      if (pl.iters / sp >= 8 && (long long)pl.m_tiles * pl.n_tiles * sp <= 2 * SMS) { splits = sp; break; }
  }
  const long long slabs = splits <= 1 ? 0 : (long long)splits * p->N * p->Do * p->Ho * p->Wo * p->Cout * 4;
  return slabs > sk ? slabs : sk;
}

static int conv3d_run_plan(const b200ldm_conv3d_params* p, const ConvPlan& pl, cudaStream_t st);

static int conv3d_tcgen05_validate(const b200ldm_conv3d_params* p) {
  B2_REQUIRE(p->dtype == B200LDM_BF16, B200LDM_BAD_ARG, "tcgen05 conv needs bf16 activations");
  B2_REQUIRE(p->ksize == 1 || p->ksize == 3, B200LDM_BAD_SHAPE, "ksize must be 1 or 3 (got %d)", p->ksize);
  B2_REQUIRE(p->stride == 1 || p->stride == 2, B200LDM_BAD_SHAPE, "stride must be 1 or 2 (got %d)", p->stride);
  B2_REQUIRE(p->C1 > 0 && p->C1 % 64 == 0 && p->C2 >= 0 && p->C2 % 64 == 0, B200LDM_BAD_SHAPE,
             "tcgen05 conv needs C1, C2 multiples of 64 (got %d, %d)", p->C1, p->C2);
  B2_REQUIRE(p->Cout > 0 && p->Cout % 8 == 0, B200LDM_BAD_SHAPE, "Cout must be a multiple of 8 (got %d)", p->Cout);
  B2_REQUIRE(!(p->stride == 2 && p->C2 > 0), B200LDM_BAD_SHAPE, "stride-2 conv takes a single source");
  B2_REQUIRE(p->stride == 1 || (p->D >= 2 && p->H >= 2 && p->W >= 2), B200LDM_BAD_SHAPE,
             "stride-2 conv needs extents >= 2");
  B2_REQUIRE(aligned16(p->x) && aligned16(p->w) && aligned16(p->y) && (!p->x2 || aligned16(p->x2)) &&
                 (!p->residual || aligned16(p->residual)),
             B200LDM_BAD_ALIGN, "conv3d pointers must be 16-byte aligned");
  B2_REQUIRE((p->C2 > 0) == (p->x2 != nullptr), B200LDM_BAD_ARG, "x2 / C2 mismatch");
  if (p->gn_stats) {
    B2_REQUIRE(p->gn_groups > 0 && p->Cout % p->gn_groups == 0, B200LDM_BAD_SHAPE, "bad gn_groups");
    const int cpg = p->Cout / p->gn_groups;
    B2_REQUIRE((cpg & (cpg - 1)) == 0 && cpg >= 2, B200LDM_BAD_SHAPE,
               "fused GroupNorm partials need channels/group a power of two >= 2 (got %d)", cpg);
  }

  if (p->upsample2x)
    B2_REQUIRE(p->ksize == 3 && p->stride == 1 && p->pad_lo == 1 && p->C2 == 0 && !p->out_f32 && !p->gn_stats &&
                   !p->residual && !p->temb && p->Cout % 64 == 0 && p->Do == 2 * p->D && p->Ho == 2 * p->H &&
                   p->Wo == 2 * p->W,
               B200LDM_BAD_SHAPE, "fused upsample conv: needs k3 s1 p1, one source, bf16 out, Cout %% 64 == 0, out = 2x in");
  return B200LDM_OK;
}

// Always asynchronous, capture-safe, no host read of device data: runs the remembered (measured) choice of this problem
// signature if b200ldm_conv3d_tune has stored one, the cost model's plan otherwise.  It never measures.
int conv3d_tcgen05(const b200ldm_conv3d_params* p, cudaStream_t st) {
  if (int rc = conv3d_tcgen05_validate(p)) return rc;
  return conv3d_run_plan(p, plan_conv(p), st);
}

// The explicit measuring entry point (b200ldm_conv3d_tune): times the candidate kernels / split factors for this exact
// problem on the caller's own buffers with CUDA events, SYNCHRONISES on each, and remembers the winner for later
// b200ldm_conv3d calls of the same signature.  `y` holds a valid result afterwards; `chunk_stats` (if given) is used as
// scratch and left zeroed.  Not capture-safe by construction: refuses a capturing stream.
int conv3d_tcgen05_tune(const b200ldm_conv3d_params* p, cudaStream_t st) {
  if (int rc = conv3d_tcgen05_validate(p)) return rc;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  B2_CHECK_CUDA(cudaStreamIsCapturing(st, &cap));
  B2_REQUIRE(cap == cudaStreamCaptureStatusNone, B200LDM_BAD_ARG,
             "b200ldm_conv3d_tune synchronises: it cannot run on a capturing stream");
  const ConvPlan pl = plan_conv(p);
  if (!tune_pending(p, pl)) return B200LDM_OK;   // nothing to choose from, pinned, disabled, or already measured
  const int SMS = sm_count();
  const bool splittable = pl.split_ok && pl.ctas1 < SMS && pl.iters >= 16;
  const ConvKey key = conv_key(p);
  {
    {
      // time the candidates on the caller's own buffers (the result is the same tensor every time): the cost model's
      // plan, the persistent kernel, and - for launches that cannot fill the chip - other split-K factors / tile
      // heights of the one-CTA-per-tile kernel
      int choices[48], nc = 0;
      choices[nc++] = 0;
      if (pl.tc2_ok) choices[nc++] = 2;
      if (pl.halo_ok && g_halo) choices[nc++] = 3;
      if (pl.tc2_ok && pl.bn == 256) choices[nc++] = 4;   // persistent kernel with two epilogue warp sets
      if (pl.sk_ok && p->workspace &&
          p->workspace_bytes >= conv_sk_workspace_bytes(conv_sk_grid(pl.m_tiles, pl.n_tiles, pl.iters)))
        choices[nc++] = 5;                                // stream-K
      if (pl.tc2_ok && getenv("B200LDM_CONV_TUNE_PAIRS")) choices[nc++] = 1;   // CTA pairs: measured slower almost everywhere
      // (2-CTA weight-multicast variants were candidates for big launches for a while: they never won by the 2 %
      // margin, and the one intermittent launch failure of the round happened with them in the set - left out)
      if (splittable) {
        static const int sps[] = {1, 2, 3, 4, 6, 8, 12, 16};
        for (int mt = 1; mt <= 2; ++mt)
          for (int sp : sps) {
            const long long ctas = (long long)ceil_div(pl.m_tiles, mt) * pl.n_tiles * sp;
            if (pl.iters / sp < 8 || ctas > 2 * SMS || (sp > 1 && ctas < 32)) continue;
            choices[nc++] = (sp << 4) | (mt << 12);
          }
      }
      if (pl.tc2_ok && pl.ctas1 < SMS && !p->upsample2x) {
        // narrower Cout tiles for launches that leave SMs idle (more CTAs, a quarter / half of the epilogue each)
        for (int code = 1; code <= 2; ++code) {
          const int alt = code == 1 ? 64 : 128;
          if (alt >= pl.bn0 || p->Cout % alt) continue;
          choices[nc++] = (code << 20) | (1 << 12);
          choices[nc++] = (code << 20) | (2 << 12);
          choices[nc++] = (code << 20) | 2;
        }
      }
      const long long M_total = (long long)p->N * p->Do * p->Ho * p->Wo;
      float ms_best = 1e30f, ms0 = 0.f, ms_tcp = 0.f;
      int best = 0;
      cudaEvent_t e0, e1;
      B2_CHECK_CUDA(cudaEventCreate(&e0));
      B2_CHECK_CUDA(cudaEventCreate(&e1));
      int rc = B200LDM_OK;
      for (int c = 0; c < nc && rc == B200LDM_OK; ++c) {
        const ConvPlan cp = plan_conv(p, choices[c]);
        if (cp.splits > 1 && (!p->workspace || p->workspace_bytes < (long long)cp.splits * M_total * p->Cout * 4)) continue;
        b200ldm_conv3d_params q = *p;
        // statistics are accumulated by atomics: tuning launches must not add to the caller's buffer
        if (cp.splits > 1 ? !split_emits_stats(p)
                          : (cp.tn != 1 || cp.bn < 64 || !plan_emits_width(cp, chunk_width_of(p)) || p->Cout % 64 ||
                             p->out_f32 || p->gn_stats || g_legacy_epilogue))
          q.chunk_stats = nullptr;
        for (int rep = 0; rep < 3 && rc == B200LDM_OK; ++rep) {
          if (rep == 1) cudaEventRecord(e0, st);
          rc = conv3d_run_plan(&q, cp, st);
        }
        cudaEventRecord(e1, st);
        float ms = 1e30f;
        if (rc == B200LDM_OK && cudaEventSynchronize(e1) == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
        if (choices[c] == 0) ms0 = ms;
        if (choices[c] == 2) ms_tcp = ms;
        // a later candidate must win by 2 %; stream-K by 10 %: sustained sampling is POWER-bound (~980 W, SM clock 1.73
        // instead of 1.96 GHz; scripts/power_ab.py), where its partial-tile traffic costs what its better schedule gains
        // (32g step 5.88 -> 5.91-6.00 ms with the 5 % wins of the 20^3 level, unchanged with the 16-23 % wins of the
        // 16g model's 18x22x14 level), and it occupies every SM while a concurrent graph branch wants some
        if (ms < ms_best * ((choices[c] & 15) == 5 ? 0.90f : 0.98f)) { ms_best = ms; best = choices[c]; }
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      if (rc != B200LDM_OK) return rc;
      if (p->chunk_stats)   // undo what the tuning launches accumulated
        B2_CHECK_CUDA(cudaMemsetAsync(p->chunk_stats, 0,
                                      (size_t)p->N * (p->Cout / chunk_width_of(p)) * 2 * sizeof(double), st));
      remember_choice(key, best);
      if (getenv("B200LDM_CONV_AUTOTUNE_LOG"))
        fprintf(stderr, "b200ldm autotune: %d+%d->%d k%d s%d%s%s%s @%dx%dx%dx%d: model plan %.3f ms, persistent %.3f ms, "
                        "best %.3f ms = kernel %d splits %d mt %d cl %d (%d candidates)\n",
                p->C1, p->C2, p->Cout, p->ksize, p->stride, p->upsample2x ? " up" : "", p->residual ? " +res" : "",
                p->skip_x ? " +skip" : "", p->N, p->Do, p->Ho, p->Wo, ms0 / 2, ms_tcp / 2, ms_best / 2, best & 15,
                (best >> 4) & 255, (best >> 12) & 15, (best >> 16) & 15, nc);
      if (getenv("B200LDM_CONV_AUTOTUNE_LOG") && (best >> 20))
        fprintf(stderr, "b200ldm autotune:   ... with Cout tile %d\n", ((best >> 20) & 3) == 1 ? 64 : 128);
    }
  }
  return B200LDM_OK;
}

static int conv3d_run_plan(const b200ldm_conv3d_params* p, const ConvPlan& pl, cudaStream_t st) {
  const int taps = p->upsample2x ? 64 : p->ksize * p->ksize * p->ksize;  // rows of the packed weight tensor
  const int Cin = p->C1 + p->C2;
  const int bn = pl.bn, n_tiles = pl.n_tiles;
  const long long M_total = (long long)p->N * p->Do * p->Ho * p->Wo;
  if (pl.splits > 1)
    B2_REQUIRE(p->workspace && p->workspace_bytes >= (long long)pl.splits * M_total * p->Cout * 4 &&
                   aligned16(p->workspace),
               B200LDM_BAD_ARG, "conv3d: split-K needs %lld workspace bytes (b200ldm_conv3d_workspace_bytes)",
               (long long)pl.splits * M_total * p->Cout * 4);

  ConvTcArgs a{};
  a.tn = pl.tn; a.td = pl.td; a.th = pl.th; a.tw = pl.tw;
  a.nb_d = pl.nb_d; a.nb_h = pl.nb_h; a.nb_w = pl.nb_w;
  a.N = p->N; a.Do = p->Do; a.Ho = p->Ho; a.Wo = p->Wo; a.Cout = p->Cout;
  a.up = p->upsample2x ? 1 : 0;
  if (a.up) { a.Do = p->D; a.Ho = p->H; a.Wo = p->W; }  // rows are low-res voxels
  a.splits = pl.splits; a.slab_elems = pl.splits > 1 ? M_total * p->Cout : 0;
  a.nchunk1 = p->C1 / 64; a.nchunk2 = p->C2 / 64;
  a.nsk1 = p->skip_x ? p->skip_C1 / 64 : 0; a.nsk2 = p->skip_x ? p->skip_C2 / 64 : 0;
  a.ksize = p->ksize; a.stride = p->stride; a.pad_lo = p->pad_lo;
  a.rows_box = a.tn * a.td * a.th * a.tw;
  a.temb_stride = p->temb_stride; a.out_f32 = p->out_f32; a.gn_groups = p->gn_groups;
  a.bias = p->bias; a.temb = p->temb; a.residual = p->residual; a.y = p->y; a.gn_stats = p->gn_stats;
  if (pl.splits > 1) {  // partial sums go to fp32 slabs; bias / temb / residual are applied by the finalize kernel
    a.out_f32 = 1; a.bias = nullptr; a.temb = nullptr; a.residual = nullptr; a.y = p->workspace;
  }
  a.trace = g_conv_trace;
  {
    const char* e = getenv("B200LDM_CONV_ASHIFT");
    a.ashift = (e && p->stride == 1 && (a.rows_box + atoi(e) <= 128 || pl.tc2 == 2)) ? atoi(e) : 0;
  }

  ConvTcMaps maps;
  memset(&maps, 0, sizeof(maps));
  const uint64_t W = p->W, H = p->H, D = p->D;
  int rc;
  if (p->stride == 1) {
    const uint64_t C = p->C1;
    const int atw = pl.tc2 == 3 ? a.tw + 2 : a.tw;   // W-halo kernel: each line is loaded with one voxel on either side
    rc = make_act_map(&maps.a[0], p->x, p->C1, p->W, p->H, p->D, p->N, C, W * C, H * W * C, D * H * W * C, atw,
                      a.th, a.td, a.tn);
    if (rc) return rc;
    if (p->C2 > 0) {
      const uint64_t C2 = p->C2;
      rc = make_act_map(&maps.a[1], p->x2, p->C2, p->W, p->H, p->D, p->N, C2, W * C2, H * W * C2, D * H * W * C2,
                        atw, a.th, a.td, a.tn);
      if (rc) return rc;
    } else {
      maps.a[1] = maps.a[0];
    }
  } else {
    const uint64_t C = p->C1;
    for (int pd = 0; pd < 2; ++pd)
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const __nv_bfloat16* base =
              reinterpret_cast<const __nv_bfloat16*>(p->x) + ((uint64_t)(pd * H + ph) * W + pw) * C;
          rc = make_act_map(&maps.a[pd * 4 + ph * 2 + pw], base, p->C1, (p->W - pw + 1) / 2, (p->H - ph + 1) / 2,
                            (p->D - pd + 1) / 2, p->N, 2 * C, 2 * W * C, 2 * H * W * C, D * H * W * C, a.tw, a.th,
                            a.td, a.tn);
          if (rc) return rc;
        }
  }
  if (p->skip_x) {
    B2_REQUIRE(p->stride == 1 && !p->upsample2x && p->skip_w && p->skip_C1 > 0 && p->skip_C1 % 64 == 0 &&
                   p->skip_C2 >= 0 && p->skip_C2 % 64 == 0 && ((p->skip_C2 > 0) == (p->skip_x2 != nullptr)) &&
                   pl.cl == 1 && aligned16(p->skip_x) && aligned16(p->skip_w) && (!p->skip_x2 || aligned16(p->skip_x2)),
               B200LDM_BAD_SHAPE, "fused skip conv: stride-1 conv, 64-multiple skip channels, aligned pointers");
    const uint64_t Wo = p->Wo, Ho = p->Ho, Do = p->Do, S1 = p->skip_C1, S2 = p->skip_C2;
    rc = make_act_map(&maps.s[0], p->skip_x, p->skip_C1, p->Wo, p->Ho, p->Do, p->N, S1, Wo * S1, Ho * Wo * S1,
                      Do * Ho * Wo * S1, a.tw, a.th, a.td, a.tn);
    if (rc) return rc;
    maps.s[1] = maps.s[0];
    if (p->skip_C2 > 0) {
      rc = make_act_map(&maps.s[1], p->skip_x2, p->skip_C2, p->Wo, p->Ho, p->Do, p->N, S2, Wo * S2, Ho * Wo * S2,
                        Do * Ho * Wo * S2, a.tw, a.th, a.td, a.tn);
      if (rc) return rc;
    }
    rc = make_weight_map(&maps.sw, p->skip_w, p->skip_C1 + p->skip_C2, (p->Cout + 15) / 16 * 16, 1, bn);
    if (rc) return rc;
  }
  // weight rows are padded by the packer to a multiple of 16 so small-Cout tiles stay in bounds
  const int cout_rows = (p->Cout + 15) / 16 * 16;
  // coalesced TMA-store epilogue whenever the output is bf16 in whole 64-channel groups
  a.tma_epilogue = (!a.out_f32 && p->Cout % 64 == 0 && !p->gn_stats && bn >= 64 && g_legacy_epilogue == 0) ? 1 : 0;
  if (a.up) B2_REQUIRE(a.tma_epilogue, B200LDM_BAD_SHAPE, "fused upsample conv needs the TMA epilogue");
  a.chunk_stats = pl.splits > 1 ? nullptr : p->chunk_stats;   // split-K: the finalize pass accumulates them
  a.cw = chunk_width_of(p);
  if (p->chunk_stats && pl.splits > 1)
    B2_REQUIRE(split_emits_stats(p), B200LDM_BAD_ARG,
               "conv3d: chunk_stats not available for this split-K problem (ask b200ldm_conv3d_emits_chunk_stats first)");
  if (p->chunk_stats && pl.splits == 1)
    B2_REQUIRE(a.tma_epilogue && pl.tn == 1 && plan_emits_width(pl, a.cw), B200LDM_BAD_ARG,
               "conv3d: chunk_stats not available for this problem (ask b200ldm_conv3d_emits_chunk_stats first)");
  // split-K partial sums through smem + TMA stores (conv3d_tc_kernel only; a box never straddles two slabs)
  a.slab_tma = (pl.splits > 1 && !pl.tc2 && p->Cout % 64 == 0 && bn >= 64 && p->N % a.tn == 0 && g_slab_tma) ? 1 : 0;
  if (a.slab_tma) {
    rc = make_slab_map(&maps.y[0], p->workspace, p->Cout, p->Wo, p->Ho, p->Do, p->N * pl.splits, a.tw, a.th, a.td, a.tn);
    if (rc) return rc;
  }
  if (a.tma_epilogue) {
    const uint64_t Co = p->Cout, Wo = p->Wo, Ho = p->Ho, Do = p->Do;
    if (!a.up) {
      rc = make_act_map(&maps.y[0], p->y, p->Cout, p->Wo, p->Ho, p->Do, p->N, Co, Wo * Co, Ho * Wo * Co,
                        Do * Ho * Wo * Co, a.tw, a.th, a.td, a.tn);
      if (rc) return rc;
    } else {
      // phase (a_d,a_h,a_w) writes output voxels (2d+a_d, 2h+a_h, 2w+a_w): a strided view per phase
      for (int ph = 0; ph < 8; ++ph) {
        const int ad = ph >> 2, ah = (ph >> 1) & 1, aw = ph & 1;
        __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(p->y) + ((uint64_t)(ad * Ho + ah) * Wo + aw) * Co;
        rc = make_act_map(&maps.y[ph], base, p->Cout, p->W, p->H, p->D, p->N, 2 * Co, 2 * Wo * Co, 2 * Ho * Wo * Co,
                          Do * Ho * Wo * Co, a.tw, a.th, a.td, a.tn);
        if (rc) return rc;
      }
    }
    if (p->residual) {
      rc = make_act_map(&maps.r, p->residual, p->Cout, p->Wo, p->Ho, p->Do, p->N, Co, Wo * Co, Ho * Wo * Co,
                        Do * Ho * Wo * Co, a.tw, a.th, a.td, a.tn);
      if (rc) return rc;
    }
  }

  const int m_tiles = pl.m_tiles, mt = pl.mt, cl = pl.cl;
  if (pl.tc2 == 3) {
    B2_REQUIRE(conv_halo_supported(bn, a), B200LDM_BAD_SHAPE, "conv3d: problem not eligible for the W-halo kernel");
    rc = make_weight_map(&maps.w, p->w, Cin, cout_rows, taps, bn);
    if (rc) return rc;
    return launch_conv_halo(maps, a, bn, m_tiles, n_tiles, st);
  }
  if (pl.tc2 == 5 && conv_sk_supported(bn, a)) {
    rc = make_weight_map(&maps.w, p->w, Cin, cout_rows, taps, bn);
    if (rc) return rc;
    return launch_conv_sk(maps, a, m_tiles, n_tiles, pl.iters, p->workspace, p->workspace_bytes, st);
  }
  if ((pl.tc2 == 2 || pl.tc2 == 4) && conv_tcp_supported(bn, a)) {
    rc = make_weight_map(&maps.w, p->w, Cin, cout_rows, taps, bn);
    if (rc) return rc;
    return launch_conv_tcp(maps, a, bn, m_tiles, n_tiles, a.up ? 8 : 1, st, pl.tc2 == 4);
  }
  if (pl.tc2 == 1 && conv_tc2_supported(bn, a)) {
    // each CTA of a pair stages half of the weight tile
    rc = make_weight_map(&maps.w, p->w, Cin, cout_rows, taps, bn / 2);
    if (rc) return rc;
    if (p->skip_x) {
      rc = make_weight_map(&maps.sw, p->skip_w, p->skip_C1 + p->skip_C2, cout_rows, 1, bn / 2);
      if (rc) return rc;
    }
    return launch_conv_tc2(maps, a, bn, m_tiles, n_tiles, a.up ? 8 : 1, st);
  }
  // the weight map box holds the rows ONE CTA fetches
  rc = make_weight_map(&maps.w, p->w, Cin, cout_rows, taps, bn / cl);
  if (rc) return rc;
  const int sp = a.up ? 8 : pl.splits;  // grid.z: split-K slabs, or the 8 upsample phases
  switch (bn) {
    case 256: rc = mt == 2 ? (cl == 2 ? launch_conv_tc<256, 2, 2>(maps, a, m_tiles, n_tiles, sp, st)
                                      : launch_conv_tc<256, 2, 1>(maps, a, m_tiles, n_tiles, sp, st))
                           : (cl == 2 ? launch_conv_tc<256, 1, 2>(maps, a, m_tiles, n_tiles, sp, st)
                                      : launch_conv_tc<256, 1, 1>(maps, a, m_tiles, n_tiles, sp, st));
      break;
    case 128: rc = mt == 2 ? (cl == 2 ? launch_conv_tc<128, 2, 2>(maps, a, m_tiles, n_tiles, sp, st)
                                      : launch_conv_tc<128, 2, 1>(maps, a, m_tiles, n_tiles, sp, st))
                           : (cl == 2 ? launch_conv_tc<128, 1, 2>(maps, a, m_tiles, n_tiles, sp, st)
                                      : launch_conv_tc<128, 1, 1>(maps, a, m_tiles, n_tiles, sp, st));
      break;
    case 64: rc = mt == 2 ? launch_conv_tc<64, 2, 1>(maps, a, m_tiles, n_tiles, sp, st)
                          : launch_conv_tc<64, 1, 1>(maps, a, m_tiles, n_tiles, sp, st);
      break;
    default: rc = mt == 2 ? launch_conv_tc<16, 2, 1>(maps, a, m_tiles, n_tiles, sp, st)
                          : launch_conv_tc<16, 1, 1>(maps, a, m_tiles, n_tiles, sp, st);
  }
  if (rc || pl.splits == 1) return rc;
  if (p->defer_finalize) return B200LDM_OK;   // the caller sums the slabs (b200ldm_conv3d_finalize / groupnorm_small_slabs)
  return launch_finalize(p, pl.splits, M_total, st);
}

}  // namespace b200ldm
