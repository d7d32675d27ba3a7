// attention_tc.cu - fused flash attention on tcgen05 / TMEM for the SpatialAttentionBlock core
// (softmax(Q K^T * scale) V over D*H*W tokens; never materialises the L x L matrix).
//
// One CTA = 128 queries of one (batch, head).  Per KV tile of BN keys:
//   MMA warp   : S = Q K^T          (A = Q smem K-major, B = K smem K-major)  -> TMEM S[j%2]
//   softmax WG : S -> regs (1 row / thread), online max (lazy rescale), P = exp2(.) -> bf16 -> smem
//                (hand-swizzled into the SWIZZLE_128B K-major UMMA layout), O rescale in TMEM if needed
//   MMA warp   : O += P V           (A = P smem K-major, B = V smem MN-major as landed by TMA) -> TMEM O
// Q/K/V tiles arrive through 3-D TMA maps (cols, L, B) over the fused [B, L, 3C] projection buffer, so
// rows past L are zero-filled; S double-buffered so QK^T of tile j+1 overlaps the softmax of tile j.
// Warp roles (192 threads): warp0 TMA, warp1 MMA + TMEM owner, warps 2..5 softmax / correction / epilogue.
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace b200ldm {

struct AttnMaps {
  CUtensorMap q, k, v;  // dims (H*d, L, B), box (64, 128|BN, 1), SWIZZLE_128B
};

struct AttnArgs {
  int H, L, d;       // L = queries per (batch, head)
  int Lk;            // keys / values per (batch, head): == L for self attention, the context length for cross attention
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* o;  // [B][L][H*d]
  unsigned long long* trace;  // debug (b200ldm_debug_set_conv_trace): per-CTA cycle counters of the softmax warps
  int dbg;                    // debug: bit0 = exp2 replaced by a multiply (is the MUFU pipe the limiter?)
  int stagger;                // cycles the softmax warps of every second CTA of an SM wait before their first tile (0: off)
  // split-KV for the last, partial wave of query tiles (one-query-tile kernel; b200ldm_attention_ws): the 1-D grid holds
  // `split_t0` whole tiles followed by (tiles - split_t0) * nsplit parts.  A part runs a contiguous range of the key tiles
  // and leaves its un-normalised O (fp32), reference max and row sum in the workspace; attention_merge_kernel combines.
  int tiles_x;                // query tiles per (batch, head)
  int split_t0, nsplit;       // nsplit <= 1: no tile is split
  float* part_o;              // [part unit][D / 4][128] float4: O, column quad major (a warp stores 512 contiguous bytes)
  float* part_ml;             // [part unit][2][128]: reference max (log2 domain), row sum
};

// CTAs that have started on each SM so far (only its parity is used): two CTAs of the one-query-tile kernel share an
// SM, and their softmax warps share the MUFU pipe of a sub-partition.  Launched together they run in lock-step - both in
// the exp phase (MUFU at half rate each), then both in the MUFU-free part of a tile (clock64 trace, round 2: exp + P store
// 885 cycles per tile against 512 for the pipe alone, 758 cycles of other work).  Starting every second CTA of an SM
// half a tile late puts one CTA's exp phase next to the other's TMEM load / row maximum / barrier work.
__device__ unsigned int g_attn_sm_arrivals[1024];

constexpr int kAttnThreads = 192;
constexpr int kAttnPackedDefault = 1;   // B200LDM_ATTN_PACKED (see launch_attention_tc)
constexpr int kAttnPTmemDefault = 1;    // B200LDM_ATTN_PTMEM: P through TMEM for the 64-column kernels (+6 %, measured); the
                                        // wider heads (tensor-bound, one CTA per SM) show no difference and keep shared memory
constexpr int kAttnSBDefault = 0;       // B200LDM_ATTN_SB
constexpr int kAttnWideTmemDefault = 0;  // B200LDM_ATTN_WIDE_TMEM: head dim 128 / 256 one-tile kernel: 1 = P in TMEM, 2 = P and Q
constexpr int kAttnQTmemDefault = 1;    // B200LDM_ATTN_QTMEM: Q in TMEM too (+3-4 % on top of P in TMEM, measured)
constexpr int kAttnW2Default = 0;          // B200LDM_ATTN_W2 (see launch_attention_tc)
constexpr int kAttnTwoIssuersDefault = 1;  // B200LDM_ATTN_TWO_ISSUERS (see launch_attention_tc)
constexpr int kAttnStaggerDefault = 0;  // B200LDM_ATTN_STAGGER: start offset (cycles) of every second CTA of an SM

// SB: ONE S buffer (with P aliased onto it, i.e. PT) instead of two: BN + D = 128 TMEM columns and no P tiles in shared
// memory (~64 KB), so THREE CTAs share an SM - three softmax warps per sub-partition.  A CTA's tiles are then strictly
// serial (softmax_j -> PV_j -> S_{j+1} -> softmax_{j+1}); the other two CTAs fill the gaps.
template <int D, int BN, bool SB = false>
struct AttnCfg {
  static constexpr int kStages = (D == 64) ? 3 : 2;
  static constexpr int kQBytes = (D / 64) * 16384;
  static constexpr int kKVSub = BN * 128;             // one 64-column sub-tile of K or V
  static constexpr int kKVBytes = (D / 64) * kKVSub;  // K (or V) tile
  static constexpr int kPBytes = SB ? 0 : (BN / 64) * 16384;   // one P tile; there are TWO (tile parity) so exp(j+1) never waits for PV(j)
  static constexpr int kSmem = kQBytes + kStages * 2 * kKVBytes + 2 * kPBytes + 1024 + 256;
  static constexpr int kNB = SB ? 1 : 2;  // S (and P) buffers
  // 2*BN + D <= 256 columns and <= ~113 KB of smem (head dim 64 with 64-key tiles): TWO CTAs per SM, i.e. two
  // independent softmax warps per sub-partition that are never in lock-step
  static constexpr int kOcc = SB ? 3 : ((2 * BN + D <= 256 && kSmem <= 115712) ? 2 : 1);
  static constexpr int kTmemCols = SB ? 128 : (kOcc == 2 ? 256 : 512);
  static constexpr int kSCol0 = 0;          // S buffers at columns [0, kNB*BN)
  static constexpr int kOCol = kNB * BN;    // O at [kNB*BN, kNB*BN + D)
  static_assert(kNB * BN + D <= kTmemCols, "TMEM budget");
  static_assert(kSmem <= 232448 / kOcc, "smem budget");
};

// MN-major SWIZZLE_128B operand (V as landed by TMA: rows = keys (K dim), 64 head-dim columns per 128-byte row).
// LBO = byte distance between 64-column blocks, SBO = distance between 8-key groups (1024).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Blackwell's two-wide fp32 instructions (SASS FFMA2 / FADD2) and the three-input maximum (FMNMX3): the softmax warps'
// exp loop is bound by ISSUE slots before it is bound by the MUFU pipe (profiles/r01_attention_experiments.md: 11.4
// cycles per MUFU.EX2 against the pipe's 8.0, with FFMA + FADD + 0.5 F2FP issued next to every exponential), so one
// instruction per PAIR of scores for the scale-and-shift and for the row sum takes 2.5 -> 1.5 companion instructions
// per exponential, and the row maximum needs 32 instead of 64 instructions per 64-key tile.
__device__ __forceinline__ void ffma2_bcast(float& d0, float& d1, float a0, float a1, float s, float c) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(s), "f"(c));
}
__device__ __forceinline__ void fadd2_acc(float& d0, float& d1, float a0, float a1) {
  asm("{\n\t.reg .b64 ra, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// 2^x for a PAIR of scores on the FMA / ALU pipes instead of the MUFU (FlashAttention-4's trick, here with the two-wide
// instructions so that it costs 5.5 instead of 9 issue slots per element): x = n + f with n = round(x) taken from the
// mantissa of x + 1.5 * 2^23, 2^f on [-0.5, 0.5] by a cubic (max relative error 7.5e-5, two orders below the bf16
// rounding of P), and n added into the exponent field.  x is clamped at -125 so that the exponent cannot wrap (2^-125
// stands in for 0: masked keys and hopeless scores contribute ~1e-38).  The kernel's exponentials are MUFU-bound (16 per
// clock and SM against 4096 tensor FLOPs per clock and SM): a share of them moved here lowers that floor.
__device__ __forceinline__ void exp2_poly2(float& p0, float& p1, float x0, float x1) {
  constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  float r0, r1, n0, n1, f0, f1, q0, q1;
  fadd2(r0, r1, x0, x1, kMagic, kMagic);          // integer part lands in the low mantissa bits
  fadd2(n0, n1, r0, r1, -kMagic, -kMagic);
  fadd2(f0, f1, x0, x1, -n0, -n1);
  ffma2(q0, q1, f0, f1, 0.05517164617776871f, 0.05517164617776871f, 0.2426111251115799f, 0.2426111251115799f);
  ffma2(q0, q1, q0, q1, f0, f1, 0.6932609677314758f, 0.6932609677314758f);
  ffma2(q0, q1, q0, q1, f0, f1, 0.9999280571937561f, 0.9999280571937561f);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}
// which of every 8 consecutive score pairs take the polynomial path (spread out so MUFU and FMA work interleave)
__device__ __forceinline__ constexpr bool poly_pair(int poly8, int q) {
  return poly8 == 4 ? (q & 1) : poly8 == 3 ? (q == 2 || q == 5 || q == 7) : poly8 == 2 ? (q == 3 || q == 7)
       : poly8 == 1 ? (q == 7) : false;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// DR = the real head dim.  DR == D except for head dim 32 (config_train_stable.json: num_head_channels 32), which runs
// on the D = 64 tiles: the 64-column TMA box at (h / 2) * 64 holds heads 2h' and 2h'+1 side by side in one
// SWIZZLE_128B row, S = Q K^T contracts only this head's two 16-column K-slices (descriptor offset (h & 1) * 64 B), PV
// multiplies P_h by the whole 64-column V box (an N = 64 MMA costs the same ~46 cycles as N = 32,
// profiles/r01e_umma_rate_microbench.log) and the epilogue stores this head's 32 columns of O.
// PACKED: 0 = scalar FFMA / FADD softmax loop, 1 = two-wide fp32 instructions, 2 .. 5 = two-wide + (PACKED - 1) of every 8
// score pairs exponentiated by exp2_poly2 instead of MUFU.EX2
// PT: P goes back into TMEM (bf16 pairs over the first BN / 2 columns of the S buffer it came from) and is the A operand of
// the PV MMA (tcgen05.mma with A in TMEM).  The head-dim-64 kernel is bound by SHARED-MEMORY bandwidth, not by MUFU or
// issue slots (round 2: the exp phase can be made 15 % shorter, a third of the exponentials can leave the MUFU pipe, the
// co-resident CTAs can be staggered - the kernel time does not move): per 64-key tile a CTA moves 16 KB (Q) + 8 KB (K) +
// 16 KB (P) + 8 KB (V) of UMMA operand reads, 16 KB of P stores and 16 KB of TMA fills through a 128 B/clk port - 625
// cycles, 1250 for the two CTAs of an SM, of a ~1600-cycle tile period (an N = 64 MMA's 48-cycle "instruction floor" IS
// the 6 KB operand read).  P in TMEM removes 32 of those 80 KB.
// QT: the query tile lives in TMEM too (D / 2 columns of bf16 pairs next to O; the softmax warps copy it there from the TMA
// landing zone once per CTA) and S = Q K^T is issued with the A operand in TMEM: the other per-tile operand that never
// changes leaves the shared-memory port (16 of the remaining 48 KB per 64-key tile).
// TI: TWO MMA-issuing threads - warp 1 issues S = Q K^T, a seventh warp issues O += P V.  The head-dim-256 kernel was bound
// by its single issuing thread (clock64 trace, round 2: 1730 cycles per key tile - 866 to issue the 16 S instructions and
// their 2 commits, 343 for PV, ~125 for each of 4 barrier waits even when they are already complete - against 1280 cycles
// of tensor work).  With two threads neither exceeds the tensor time.  tcgen05.commit only covers the MMAs of the
// committing thread, so s_full(j) no longer implies that PV(j-2) has retired: the softmax warps wait for pv_done
// before they overwrite a P buffer.
// W2: TWO softmax warpgroups - warps 2..5 take the first half of a tile's keys, warps 6..9 the second half of the same rows
// (a TMEM lane quarter belongs to warp % 4, so both warps of a pair may read the same rows).  The row maxima meet in shared
// memory once per tile (a 64-thread named barrier per lane quarter); reference max and rescale decisions are then identical
// in both halves, the row sums are combined in the epilogue.  Four instead of two softmax warps per scheduler (two CTAs
// per SM) to hide the per-warp serial latency that bounds the head-dim-64 kernel (profiles/r02_attention_experiments.md).
template <int D, int BN, int DR = D, int PACKED = 1, bool PT = false, bool SB = false, bool QT = false, bool TI = false,
          bool W2 = false>
__global__ void __launch_bounds__(kAttnThreads + (TI ? 32 : 0) + (W2 ? 128 : 0), AttnCfg<D, BN, SB>::kOcc)
attention_tc_kernel(const __grid_constant__ AttnMaps maps, const AttnArgs a) {
  static_assert(!W2 || (BN % 64 == 0 && DR % 32 == 0 && !SB), "two softmax warpgroups: 32-column halves");
  constexpr int kTiWarp = W2 ? 10 : 6;          // the PV-issuing warp of TI
  constexpr int kSmxThreads = W2 ? 256 : 128;   // threads that arrive on s_empty / p_full
  static_assert(!TI || !SB, "two issuing threads need the two S buffers");
  static_assert(DR == D || (D == 64 && DR == 32), "head dim 32 rides on the 64-column tiles");
  static_assert(!SB || PT, "the single S buffer holds P as well");
  static_assert(!QT || (PT && AttnCfg<D, BN, SB>::kOCol + D + D / 2 <= AttnCfg<D, BN, SB>::kTmemCols),
                "Q in TMEM: D / 2 more columns (head dim 256 with 64-key tiles fills all 512)");
  constexpr int kQCol = AttnCfg<D, BN, SB>::kOCol + D;   // Q (QT) at [kOCol + D, kOCol + D + D / 2)
  using Cfg = AttnCfg<D, BN, SB>;
  // buffer index / mbarrier phase parity of tile j: two buffers alternate (a barrier completes every second tile), one
  // buffer completes a phase per tile
  auto bf = [](int j) { return SB ? 0 : (j & 1); };
  auto ph = [](int j) { return (uint32_t)(SB ? j : (j >> 1)) & 1u; };
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  // K and V are staged and released SEPARATELY: K_j is free as soon as S_j retires, a whole tile before PV_j frees
  // V_j, so the load of K_{j+STAGES} starts that much earlier (with a shared stage the d = 256 kernel waited ~1400
  // cycles per tile for the next S: the TMA round trip sat between PV_j and S_{j+2}, gpurun_out/attn_trace_d256.log)
  const uint32_t k_smem = q_smem + Cfg::kQBytes;                   // K stage s at + s*kKVBytes
  const uint32_t v_smem = k_smem + STAGES * Cfg::kKVBytes;         // V stage s at + s*kKVBytes
  const uint32_t p_smem = v_smem + STAGES * Cfg::kKVBytes;
  const uint32_t bar = p_smem + 2 * Cfg::kPBytes;
  // barrier map (8 bytes each)
  const uint32_t q_full = bar;
  const uint32_t k_full = bar + 8;                   // [STAGES]
  const uint32_t k_empty = k_full + 8 * STAGES;      // [STAGES]
  const uint32_t v_full = k_empty + 8 * STAGES;      // [STAGES]
  const uint32_t v_empty = v_full + 8 * STAGES;      // [STAGES]
  const uint32_t s_full = v_empty + 8 * STAGES;      // [2]
  const uint32_t s_empty = s_full + 16;              // [2]
  // p_full is per P buffer (tile parity), like pv_done: with ONE barrier for both buffers the softmax warps could
  // complete the phases of tiles j and j+1 (j+1 only needs S_{j+1} and PV_{j-1}) while the MMA thread still waited
  // for a late V_j; the thread then found the barrier two phases on, waited for the parity of tile j+2 - whose P
  // needs PV_j - and the CTA deadlocked (seen as a 2 s mbarrier timeout under concurrent streams, and most likely the
  // one unexplained launch failure of DESIGN.md section 6: V_0 is the third load of a CTA and can trail K_0 / K_1).
  const uint32_t p_full = s_empty + 16;              // [2]
  const uint32_t pv_done = p_full + 16;              // [2]: PV of even / odd tiles (one per P buffer)
  const uint32_t tmem_slot = pv_done + 16;
  const uint32_t q_tmem = tmem_slot + 8;             // QT: the softmax warps have copied Q into TMEM (128 arrivals)
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D grid: whole query tiles first (linear tile id = x + tiles_x * (h + H * b)), then the parts of the split tiles
  int tile_id = (int)blockIdx.x, part = 0, nparts = 1;
  if (a.nsplit > 1 && tile_id >= a.split_t0) {
    const int u = tile_id - a.split_t0;
    tile_id = a.split_t0 + u / a.nsplit;
    part = u - (u / a.nsplit) * a.nsplit;
    nparts = a.nsplit;
  }
  const int q0 = (tile_id % a.tiles_x) * 128;
  const int h = (tile_id / a.tiles_x) % a.H, b = tile_id / (a.tiles_x * a.H);
  const int ntiles_all = (a.Lk + BN - 1) / BN;
  const int jbase = (int)((long long)ntiles_all * part / nparts);               // first key tile of this part
  const int ntiles = (int)((long long)ntiles_all * (part + 1) / nparts) - jbase;  // key tiles of this part (>= 1)
  const int colh = h * DR;                    // this head's first column in the [.., H * DR] projections
  const int col0 = DR == D ? colh : (colh & ~63);   // first column of the TMA boxes
  const int kk0 = DR == D ? 0 : ((colh & 63) >> 4); // first 16-column K-slice of this head inside the box

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&maps.q); ptx::prefetch_tmap(&maps.k); ptx::prefetch_tmap(&maps.v);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(k_full + 8 * s, 1); ptx::mbar_init(k_empty + 8 * s, 1);
      ptx::mbar_init(v_full + 8 * s, 1); ptx::mbar_init(v_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(s_full + 8 * s, 1); ptx::mbar_init(s_empty + 8 * s, kSmxThreads); }
    ptx::mbar_init(p_full, kSmxThreads);
    ptx::mbar_init(p_full + 8, kSmxThreads);
    ptx::mbar_init(pv_done, 1);
    ptx::mbar_init(pv_done + 8, 1);
    ptx::mbar_init(q_tmem, 128);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  if (Cfg::kOcc >= 2 && a.stagger > 0 && threadIdx.x == 64) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tmem_slot_ptr[1] = atomicAdd(&g_attn_sm_arrivals[smid & 1023u], 1u);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const bool late = Cfg::kOcc >= 2 && a.stagger > 0 && (tmem_slot_ptr[1] & 1u);

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, Cfg::kQBytes);
#pragma unroll
      for (int sub = 0; sub < D / 64; ++sub)
        ptx::tma_load_3d(q_smem + sub * 16384, &maps.q, q_full, col0 + sub * 64, q0, b);
      auto load_k = [&](int j) {
        const int s = j % STAGES;
        ptx::mbar_wait_long(k_empty + 8 * s, ((uint32_t)(j / STAGES) & 1u) ^ 1u, 10);
        ptx::mbar_expect_tx(k_full + 8 * s, Cfg::kKVBytes);
#pragma unroll
        for (int sub = 0; sub < D / 64; ++sub)
          ptx::tma_load_3d(k_smem + s * Cfg::kKVBytes + sub * Cfg::kKVSub, &maps.k, k_full + 8 * s, col0 + sub * 64,
                           (jbase + j) * BN, b);
      };
      auto load_v = [&](int j) {
        const int s = j % STAGES;
        ptx::mbar_wait_long(v_empty + 8 * s, ((uint32_t)(j / STAGES) & 1u) ^ 1u, 10);
        ptx::mbar_expect_tx(v_full + 8 * s, Cfg::kKVBytes);
#pragma unroll
        for (int sub = 0; sub < D / 64; ++sub)
          ptx::tma_load_3d(v_smem + s * Cfg::kKVBytes + sub * Cfg::kKVSub, &maps.v, v_full + 8 * s, col0 + sub * 64,
                           (jbase + j) * BN, b);
      };
      load_k(0);   // K runs one tile ahead of V: its stage frees first and S needs it first
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 1 || (TI && warp == kTiWarp)) {
    // =========================== MMA issuer(s) ===========================
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, BN);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_bmn(128, D);
      // operand descriptors as (low word, constant high word): stepping to another stage / 64-column block / k-slice
      // is one 32-bit add (ptx::umma_bf16_lohi) - this thread is latency-bound on descriptor arithmetic otherwise
      constexpr uint32_t khi = ptx::umma_desc_hi_sw128();      // K-major SWIZZLE_128B (Q, K, P), SBO 1024
      const uint32_t q_lo = ptx::umma_desc_lo_sw128(q_smem);
      const uint32_t k_lo0 = ptx::umma_desc_lo_sw128(k_smem);
      const uint32_t p_lo0 = ptx::umma_desc_lo_sw128(p_smem);
      // V: MN-major SWIZZLE_128B, LBO = one 64-column block (kKVSub), SBO = 1024
      const uint32_t v_lo0 = ((v_smem & 0x3FFFFu) >> 4) | ((uint32_t)((Cfg::kKVSub >> 4) & 0x3FFF) << 16);
#ifdef B200LDM_ATTN_TRACE
      const bool mtr = a.trace != nullptr;
#else
      constexpr bool mtr = false;
#endif
      long long mc[6] = {0, 0, 0, 0, 0, 0};   // wait K | wait S buffer | issue S | wait V | wait P | issue PV
      long long mt0 = 0;
      auto tick = [&](int i) { if (mtr) { const long long t = clock64(); mc[i] += t - mt0; mt0 = t; } };
      auto issue_s = [&](int j) {
        const int s = j % STAGES;
        if (mtr) mt0 = clock64();
        ptx::mbar_wait(k_full + 8 * s, (uint32_t)(j / STAGES) & 1u, 11);
        tick(0);
        if (!SB && j >= 2) {
          // S_j overwrites the buffer of tile j - 2: free once the softmax warps hold S_{j-2} in registers - or, with P
          // aliased onto it (PT), once PV_{j-2} has consumed that P.  (SB: S_j is issued right behind PV_{j-1}, the
          // MMAs of one thread execute in issue order, so the WAR hazard on the shared buffer needs no barrier.)
          if constexpr (PT) ptx::mbar_wait(pv_done + 8 * (j & 1), (uint32_t)((j >> 1) - 1) & 1u, 12);
          else ptx::mbar_wait(s_empty + 8 * (j & 1), (uint32_t)((j >> 1) - 1) & 1u, 12);
        }
        ptx::tc_fence_after();
        tick(1);
        const uint32_t k_lo = k_lo0 + (uint32_t)s * (uint32_t)(Cfg::kKVBytes >> 4);
        const uint32_t tS = tmem + Cfg::kSCol0 + bf(j) * BN;
#pragma unroll
        for (int k = 0; k < DR / 16; ++k) {
          const int kk = k + kk0;
          if constexpr (QT) {
            ptx::umma_bf16_ts(tS, tmem + kQCol + kk * 8,
                              (static_cast<uint64_t>(khi) << 32) |
                                  static_cast<uint64_t>(k_lo + (uint32_t)(((kk >> 2) * Cfg::kKVSub + (kk & 3) * 32) >> 4)),
                              idesc_s, k > 0 ? 1u : 0u);
          } else {
            ptx::umma_bf16_lohi(tS, q_lo + (uint32_t)(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                                k_lo + (uint32_t)(((kk >> 2) * Cfg::kKVSub + (kk & 3) * 32) >> 4), khi, khi, idesc_s,
                                k > 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(s_full + 8 * bf(j));
        ptx::umma_commit(k_empty + 8 * s);   // K_j is free once S_j retires
        tick(2);
      };
      auto issue_pv = [&](int j) {
        const int s = j % STAGES;
        if (mtr) mt0 = clock64();
        ptx::mbar_wait(v_full + 8 * s, (uint32_t)(j / STAGES) & 1u, 16);
        tick(3);
        ptx::mbar_wait(p_full + 8 * bf(j), ph(j), 14);
        ptx::tc_fence_after();
        tick(4);
        const uint32_t v_lo = v_lo0 + (uint32_t)s * (uint32_t)(Cfg::kKVBytes >> 4);
        const uint32_t p_lo = p_lo0 + (uint32_t)(j & 1) * (uint32_t)(Cfg::kPBytes >> 4);
        const uint32_t tO = tmem + Cfg::kOCol;
        if constexpr (PT) {
          const uint32_t tP = tmem + Cfg::kSCol0 + bf(j) * BN;   // 16 keys = 8 columns of bf16 pairs per k-slice
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk)
            ptx::umma_bf16_ts(tO, tP + kk * 8,
                              (static_cast<uint64_t>(khi) << 32) | static_cast<uint64_t>(v_lo + (uint32_t)((kk * 2048) >> 4)),
                              idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        } else {
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk)
            ptx::umma_bf16_lohi(tO, p_lo + (uint32_t)(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                                v_lo + (uint32_t)((kk * 2048) >> 4), khi, khi, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(v_empty + 8 * s);   // V_j is free once PV_j retires
        ptx::umma_commit(pv_done + 8 * bf(j));
        tick(5);
      };
      if (!TI || warp == 1) {
        if constexpr (QT) ptx::mbar_wait(q_tmem, 0, 13);
        else ptx::mbar_wait(q_full, 0, 13);
      }
      if constexpr (TI) {
        if (warp == 1) { for (int j = 0; j < ntiles; ++j) issue_s(j); }   // throttled by K stages and s_empty
        else { for (int j = 0; j < ntiles; ++j) issue_pv(j); }
      } else {
        issue_s(0);
        for (int j = 0; j < ntiles; ++j) {
          if (!SB && j + 1 < ntiles) issue_s(j + 1);
          issue_pv(j);
          if (SB && j + 1 < ntiles) issue_s(j + 1);   // queued right behind PV_j: it overwrites the P that PV_j reads
        }
      }
      if (mtr && blockIdx.x < 1024) {
        unsigned long long* o = a.trace + (size_t)blockIdx.x * 2 * 8 + 8;
        for (int i = 0; i < 6; ++i)
          if (mc[i]) o[i] = (unsigned long long)mc[i];
        o[6] = (unsigned long long)ntiles;
      }
    }
  } else {
    // =========================== softmax / correction / epilogue ===========================
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const int half = W2 ? ((warp - 2) >> 2) : 0;   // which half of a tile's keys (and of the O columns) this thread owns
    constexpr int NW = W2 ? BN / 2 : BN;           // scores per thread and tile
    const int coff = half * NW;                    // first key column of this thread inside the tile
    float m_used = -INFINITY;  // reference max (log2 domain) the running sums are expressed against
    float l_sum = 0.f;
    uint8_t* p_base = smem_gen + (p_smem - smem_base);
    // W2 exchange area behind the barriers: row maxima [tile parity][half][128], row sums [half][128]
    // (head dim 256 has 1792 bytes left: ONE maxima buffer, protected by a second barrier per tile, which the row sums
    // reuse in the epilogue)
    constexpr bool kXD = Cfg::kSmem + 3072 <= 232448 / Cfg::kOcc;
    float* xmax = reinterpret_cast<float*>(smem_gen + (bar - smem_base) + 256);
    float* xsum = kXD ? xmax + 2 * 2 * 128 : xmax;
    if (QT && half == 0) {
      // this thread's query row: 8 x 16 bytes out of the SWIZZLE_128B landing zone (chunk c of row r sits at
      // r * 128 + ((c ^ (r & 7)) << 4)) -> 32 words of bf16 pairs in K order -> TMEM lane `row`
      ptx::mbar_wait(q_full, 0, 24);
#pragma unroll
      for (int sub = 0; sub < D / 64; ++sub) {     // one 64-column sub-tile of Q = 32 TMEM columns
        const uint8_t* q_gen = smem_gen + (q_smem - smem_base) + sub * 16384 + row * 128;
        uint32_t qw[2][16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(q_gen + ((c ^ (row & 7)) << 4));
          qw[c >> 2][(c & 3) * 4 + 0] = v.x; qw[c >> 2][(c & 3) * 4 + 1] = v.y;
          qw[c >> 2][(c & 3) * 4 + 2] = v.z; qw[c >> 2][(c & 3) * 4 + 3] = v.w;
        }
        ptx::tmem_st16(tmem + lane_off + kQCol + sub * 32, qw[0]);
        ptx::tmem_st16(tmem + lane_off + kQCol + sub * 32 + 16, qw[1]);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(q_tmem);
    }

#ifdef B200LDM_ATTN_TRACE
    const bool tr = a.trace != nullptr && lane == 0 && qd == 0;
#else
    constexpr bool tr = false;
#endif
    long long tc[6] = {0, 0, 0, 0, 0, 0};
    if (late) {   // every second CTA of this SM: start half a tile behind the co-resident CTA (see g_attn_sm_arrivals)
      const long long t_s = clock64();
      while (clock64() - t_s < a.stagger) {}
    }
    const long long t_begin = tr ? clock64() : 0;
    for (int j = 0; j < ntiles; ++j) {
      long long t0 = tr ? clock64() : 0, t1;
      uint8_t* p_gen = p_base + (size_t)(j & 1) * Cfg::kPBytes;
      ptx::mbar_wait(s_full + 8 * bf(j), ph(j), 20);
      ptx::tc_fence_after();
      if (tr) { t1 = clock64(); tc[0] += t1 - t0; t0 = t1; }
      const uint32_t tS = tmem + lane_off + Cfg::kSCol0 + bf(j) * BN;
      float sv[NW];  // RAW scores; the softmax scale is folded into the exp2 FFMA below
      {
        uint32_t r[NW / 32][32];   // every load in flight before the single wait
#pragma unroll
        for (int c = 0; c < NW / 32; ++c) ptx::tmem_ld32(tS + coff + c * 32, r[c]);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < NW / 32; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[c * 32 + i] = __uint_as_float(r[c][i]);
      }
      ptx::tc_fence_before();
      if constexpr (!PT) ptx::mbar_arrive(s_empty + 8 * (j & 1));  // S buffer may be overwritten by tile j+2
      if (tr) { t1 = clock64(); tc[1] += t1 - t0; t0 = t1; }

      const int kvalid = a.Lk - (jbase + j) * BN;  // keys in this tile that exist
      if (kvalid < BN) {                // last tile only (warp-uniform)
#pragma unroll
        for (int i = 0; i < NW; ++i)
          if (coff + i >= kvalid) sv[i] = -INFINITY;
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // 4 chains: no 128-deep dependency
      if constexpr (PACKED != 0) {
#pragma unroll
        for (int i = 0; i < NW; i += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mx4[u] = fmax3(mx4[u], sv[i + u], sv[i + 4 + u]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < NW; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], sv[i + u]);
        }
      }
      float mx_raw = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if constexpr (W2) {
        // the other half of this row: exchange through shared memory (buffer = tile parity: the next write to it comes
        // after the barrier of tile j + 1, which the partner only passes once it has read this one)
        float* xm = xmax + (kXD ? (j & 1) * 256 : 0);
        xm[half * 128 + row] = mx_raw;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
        mx_raw = fmaxf(mx_raw, xm[(half ^ 1) * 128 + row]);
        if constexpr (!kXD) asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");   // read before the next tile's write
      }
      const float mx = mx_raw * a.scale_log2;  // scale > 0
      // lazy rescale: keep the old reference unless the max grew by more than 2^8 (warp-uniform decision)
      const bool need = mx > m_used + 8.0f;
      const bool any_need = __any_sync(0xffffffffu, need);
      float alpha = 1.0f;
      if (tr) { t1 = clock64(); tc[2] += t1 - t0; t0 = t1; }
      // P is double-buffered by tile parity: writing P(j) only needs PV(j-2) (same buffer) to have retired, so this
      // tile's exp phase overlaps PV(j-1) - with a single buffer the chain softmax -> PV -> softmax serialised a
      // ~900-cycle MMA round trip behind every tile.  O itself is only touched when the lazy rescale fires.
      // (no pv_done wait for the P buffer: tcgen05.commit covers every earlier MMA of the issuing thread, S_j was issued
      // after PV_{j-2}, so s_full(j) - waited for above - already implies that PV_{j-2} has retired and P(j & 1) is free)
      if constexpr (TI && !PT) {
        // two issuing threads: s_full(j) says nothing about PV(j-2) any more - the shared-memory buffer P(j & 1) is free
        // once that PV has retired.  (P in TMEM: P(j) replaces S(j) in place, and S(j) was only issued after the S thread
        // itself waited for pv_done(j-2).)
        if (j >= 2) ptx::mbar_wait(pv_done + 8 * bf(j), ph(j - 2), 25);
      }
      if (tr) { t1 = clock64(); tc[3] += t1 - t0; t0 = t1; }
      if (any_need) {
        if (need) {
          alpha = exp2f(m_used - mx);  // 0 on the first tile (m_used = -inf)
          m_used = mx;
          l_sum *= alpha;
        }
        if (j > 0) {
          ptx::mbar_wait(pv_done + 8 * bf(j - 1), ph(j - 1), 23);  // PV_{j-1} retired: O is ours
          ptx::tc_fence_after();
          const uint32_t tO = tmem + lane_off + Cfg::kOCol + (uint32_t)(W2 ? half * (D / 2) : 0);   // W2: this half's columns
#pragma unroll
          for (int c = 0; c < (W2 ? D / 2 : D); c += 16) {
            uint32_t r[16];
            ptx::tmem_ld16(tO + c, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            ptx::tmem_st16(tO + c, r);
          }
          ptx::tmem_st_wait();
        }
      }
      if (tr) { t1 = clock64(); tc[4] += t1 - t0; t0 = t1; }
      // P = exp2(s*scale - m_used) -> bf16, written in the SWIZZLE_128B K-major layout (16-byte chunk c of row r
      // lives at r*128 + ((c ^ (r & 7)) << 4) inside each [128 x 64-key] sub-tile).  One FFMA + one MUFU.EX2 per
      // element; the row sum is taken in fp32 over 4 independent chains.
      const float neg_m = -m_used;
      float ps[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // 4 independent (two-wide) sum chains
      // pair q = scores 2q, 2q+1 of this row -> P as one packed bf16x2 word; q is a compile-time constant after unrolling
      auto exp_pair = [&](int q) -> uint32_t {
        float x0, x1, p0, p1;
        if constexpr (PACKED != 0) {
          ffma2_bcast(x0, x1, sv[2 * q], sv[2 * q + 1], a.scale_log2, neg_m);
        } else {
          x0 = fmaf(sv[2 * q], a.scale_log2, neg_m);
          x1 = fmaf(sv[2 * q + 1], a.scale_log2, neg_m);
        }
        if (PACKED >= 2 && poly_pair(PACKED - 1, q & 7)) {
          exp2_poly2(p0, p1, x0, x1);
        } else {
          p0 = ex2_approx(x0);
          p1 = ex2_approx(x1);
        }
        if constexpr (PACKED != 0) {
          fadd2_acc(ps[2 * (q & 3)], ps[2 * (q & 3) + 1], p0, p1);
        } else {
          ps[q & 3] += p0 + p1;
        }
        const __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
        return *reinterpret_cast<const uint32_t*>(&hb);
      };
      if constexpr (PT) {
        const uint32_t tP = tmem + lane_off + Cfg::kSCol0 + bf(j) * BN + (uint32_t)(coff / 2);   // over the columns S_j was read from
#pragma unroll
        for (int c32 = 0; c32 < NW / 32; ++c32) {
          uint32_t pw[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) pw[e] = exp_pair(c32 * 16 + e);
          ptx::tmem_st16(tP + c32 * 16, pw);
        }
        ptx::tmem_st_wait();
      } else {
#pragma unroll
        for (int c8l = 0; c8l < NW / 8; ++c8l) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) pk[e] = exp_pair(c8l * 4 + e);
          const int c8 = coff / 8 + c8l;   // 16-byte chunk of the whole tile row
          const int sub = c8 >> 3, cc = c8 & 7;
          uint4* dst = reinterpret_cast<uint4*>(p_gen + sub * 16384 + row * 128 + ((cc ^ (row & 7)) << 4));
          *dst = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      const float psum = ((ps[0] + ps[1]) + (ps[2] + ps[3])) + ((ps[4] + ps[5]) + (ps[6] + ps[7]));
      l_sum += psum;
      if constexpr (!PT) ptx::fence_proxy_async();   // generic-proxy smem writes -> visible to the UMMA (async proxy)
      ptx::tc_fence_before();     // orders the TMEM rescale stores before the MMA warp's next PV
      ptx::mbar_arrive(p_full + 8 * bf(j));
      if (tr) { t1 = clock64(); tc[5] += t1 - t0; }
    }
    if (tr) {
      const size_t cta = blockIdx.x;
      if (cta < 1024) {
        unsigned long long* o = a.trace + cta * 2 * 8;
        for (int i = 0; i < 6; ++i) o[i] = (unsigned long long)tc[i];
        o[6] = (unsigned long long)ntiles;
        o[7] = (unsigned long long)(clock64() - t_begin);
      }
    }
    // ---- epilogue: O / l -> bf16 -> global
    ptx::mbar_wait(pv_done + 8 * bf(ntiles - 1), ph(ntiles - 1), 22);
    ptx::tc_fence_after();
    const int qi = q0 + row;
    if constexpr (W2) {
      // the row sums of the two key halves meet here (the reference max is already common)
      xsum[half * 128 + row] = l_sum;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
      l_sum += xsum[(half ^ 1) * 128 + row];
    }
    constexpr int DRH = W2 ? DR / 2 : DR;          // O columns this thread writes out
    const int ocol = W2 ? half * DRH : 0;          // ... starting here inside the head
    const uint32_t tO = tmem + lane_off + Cfg::kOCol + (uint32_t)(colh - col0 + ocol);   // this head's columns of O
    if (nparts > 1) {
      // a part of a split tile: O stays un-normalised (fp32) next to its reference max and row sum
      const size_t unit = (size_t)blockIdx.x - (size_t)a.split_t0;
      float4* po = reinterpret_cast<float4*>(a.part_o) + unit * (size_t)(DR / 4) * 128 + (size_t)(ocol / 4) * 128 + row;
#pragma unroll
      for (int c = 0; c < DRH; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(tO + c, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          __stcg(po + (size_t)(c / 4 + j4) * 128,
                 make_float4(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1]), __uint_as_float(r[4 * j4 + 2]),
                             __uint_as_float(r[4 * j4 + 3])));
      }
      if (half == 0) {
        a.part_ml[unit * 256 + row] = m_used;
        a.part_ml[unit * 256 + 128 + row] = l_sum;
      }
      ptx::tc_fence_before();
    } else {
    const float inv = 1.0f / l_sum;
    __nv_bfloat16* orow = a.o + ((size_t)b * a.L + (qi < a.L ? qi : 0)) * ((size_t)a.H * DR) + colh + ocol;
#pragma unroll
    for (int c = 0; c < DRH; c += 16) {
      uint32_t r[16];
      ptx::tmem_ld16(tO + c, r);
      ptx::tmem_ld_wait();
      if (qi < a.L) {
        float f[8];
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[hlf * 8 + e]) * inv;
          Vec8<__nv_bfloat16>::store(orow + c + hlf * 8, f);
        }
      }
    }
    ptx::tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem);
  }
}

// ------------------------------------------------------------------------------------------
// v2: two 128-query tiles per CTA, one softmax warpgroup each ("ping-pong"): while warpgroup A runs the exp2 /
// rescale / P-write of its tile, the tensor core works for warpgroup B (PV_B, S_B of the next KV tile) and vice
// versa, so neither the MUFU pipe nor the MMA pipe waits on a single dependent chain; K/V tiles are fetched once
// for both query tiles.  S is single-buffered per warpgroup (S_w(j+1) is issued right after PV_w(j), which already
// requires P_w(j), i.e. that S_w(j) has been consumed).
// Warp roles (320 threads): warp0 TMA, warp1 MMA + TMEM owner, warps 2..5 softmax A, warps 6..9 softmax B.
static unsigned long long* g_attn_trace = nullptr;
void set_attn_trace(void* p) { g_attn_trace = reinterpret_cast<unsigned long long*>(p); }

// OCC = CTAs per SM: 2 (only D = 64, BN = 64: 256 TMEM columns, ~99 KB smem, <= 102 registers) puts FOUR softmax warps
// on every SM sub-partition - the exp loop is latency / issue bound with two (profiles/r01_attention_experiments.md).
template <int D, int BN, int OCC = 1>
struct Attn2Cfg {
  static constexpr int kStages = OCC == 2 ? 2 : 3;
  static constexpr int kQBytes = 2 * (D / 64) * 16384;       // two query tiles
  static constexpr int kKVSub = BN * 128;
  static constexpr int kKVBytes = (D / 64) * kKVSub;
  static constexpr int kPBytes = (BN / 64) * 16384;           // per warpgroup
  static constexpr int kSmem = kQBytes + kStages * 2 * kKVBytes + 2 * kPBytes + 1024 + 256;
  static constexpr int kTmemCols = OCC == 2 ? 256 : 512;
  static constexpr int kSCol = 0;             // S_w at w*BN
  static constexpr int kOCol = 2 * BN;        // O_w at kOCol + w*D
  static constexpr int kThreads = 320;
  static_assert(2 * BN + 2 * D <= kTmemCols, "TMEM budget");
  static_assert(kSmem <= (OCC == 2 ? 115712 : 232448), "smem budget");
};

template <int D, int BN, int OCC>
__global__ void __launch_bounds__(Attn2Cfg<D, BN, OCC>::kThreads, OCC)
attention_tc2_kernel(const __grid_constant__ AttnMaps maps, const AttnArgs a) {
  using Cfg = Attn2Cfg<D, BN, OCC>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;                                   // tile w at + w*(D/64)*16384
  const uint32_t kv_smem = q_smem + Cfg::kQBytes;
  const uint32_t p_smem = kv_smem + STAGES * 2 * Cfg::kKVBytes;        // P_w at + w*kPBytes
  const uint32_t bar = p_smem + 2 * Cfg::kPBytes;
  const uint32_t q_full = bar;
  const uint32_t kv_full = bar + 8;                  // [STAGES]
  const uint32_t kv_empty = kv_full + 8 * STAGES;    // [STAGES]
  const uint32_t s_full = kv_empty + 8 * STAGES;     // [2]  (per warpgroup)
  const uint32_t p_full = s_full + 16;               // [2]
  const uint32_t pv_done = p_full + 16;              // [2]
  const uint32_t s_free = pv_done + 16;              // [2]  S_w is in the softmax warps' registers (4 warp arrivals)
  const uint32_t tmem_slot = s_free + 16;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (a.Lk + BN - 1) / BN;
  const int col0 = h * D;
  const bool tileB_exists = q0 + 128 < a.L;   // CTA-uniform: the second query tile may be entirely past L

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&maps.q); ptx::prefetch_tmap(&maps.k); ptx::prefetch_tmap(&maps.v);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(kv_full + 8 * s, 1); ptx::mbar_init(kv_empty + 8 * s, 1); }
    for (int w = 0; w < 2; ++w) {
      ptx::mbar_init(s_full + 8 * w, 1);
      ptx::mbar_init(p_full + 8 * w, 128);
      ptx::mbar_init(pv_done + 8 * w, 1);
      ptx::mbar_init(s_free + 8 * w, 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const int nw = tileB_exists ? 2 : 1;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, Cfg::kQBytes);
#pragma unroll
      for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int sub = 0; sub < D / 64; ++sub)
          ptx::tma_load_3d(q_smem + (w * (D / 64) + sub) * 16384, &maps.q, q_full, col0 + sub * 64, q0 + w * 128, b);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (uint32_t)(j / STAGES) & 1u;
        ptx::mbar_wait_long(kv_empty + 8 * s, ph ^ 1u, 10);   // suspended wait: the producer has stages of slack and must not steal issue slots
        const uint32_t ks = kv_smem + s * 2 * Cfg::kKVBytes, vs = ks + Cfg::kKVBytes;
        ptx::mbar_expect_tx(kv_full + 8 * s, 2 * Cfg::kKVBytes);
#pragma unroll
        for (int sub = 0; sub < D / 64; ++sub) {
          ptx::tma_load_3d(ks + sub * Cfg::kKVSub, &maps.k, kv_full + 8 * s, col0 + sub * 64, j * BN, b);
          ptx::tma_load_3d(vs + sub * Cfg::kKVSub, &maps.v, kv_full + 8 * s, col0 + sub * 64, j * BN, b);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, BN);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_bmn(128, D);
      auto issue_s = [&](int w, int j) {   // S_w(j) = Q_w K_j^T (K_j must have landed)
        const uint32_t ks = kv_smem + (j % STAGES) * 2 * Cfg::kKVBytes;
        const uint32_t qs = q_smem + w * (D / 64) * 16384;
        const uint32_t tS = tmem + Cfg::kSCol + w * BN;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t da = ptx::umma_desc_k_sw128(qs + (kk >> 2) * 16384 + (kk & 3) * 32);
          const uint64_t db = ptx::umma_desc_k_sw128(ks + (kk >> 2) * Cfg::kKVSub + (kk & 3) * 32);
          ptx::umma_bf16(tS, da, db, idesc_s, kk > 0 ? 1u : 0u);
        }
        ptx::umma_commit(s_full + 8 * w);
      };
      ptx::mbar_wait(q_full, 0, 13);
      ptx::mbar_wait(kv_full, 0, 11);
      ptx::tc_fence_after();
      for (int w = 0; w < nw; ++w) issue_s(w, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % STAGES;
        const uint32_t vs = kv_smem + s * 2 * Cfg::kKVBytes + Cfg::kKVBytes;
        // S_w(j+1) goes out as soon as the softmax warps hold S_w(j) in registers - a whole exp phase before P_w(j)
        // exists - so the next tile's scores are waiting when the warpgroup comes back for them (the chain
        // softmax -> PV -> S -> softmax used to cost ~600 tensor-pipe cycles of idle MUFU per tile)
        if (j + 1 < ntiles) {
          ptx::mbar_wait(kv_full + 8 * ((j + 1) % STAGES), (uint32_t)((j + 1) / STAGES) & 1u, 11);
          for (int w = 0; w < nw; ++w) {
            ptx::mbar_wait(s_free + 8 * w, (uint32_t)j & 1u, 15);
            ptx::tc_fence_after();
            issue_s(w, j + 1);
          }
        }
        for (int w = 0; w < nw; ++w) {
          ptx::mbar_wait(p_full + 8 * w, (uint32_t)j & 1u, 14);   // P_w(j) staged, O_w rescaled
          ptx::tc_fence_after();
          const uint32_t ps = p_smem + w * Cfg::kPBytes;
          const uint32_t tO = tmem + Cfg::kOCol + w * D;
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            const uint64_t da = ptx::umma_desc_k_sw128(ps + (kk >> 2) * 16384 + (kk & 3) * 32);
            const uint64_t db = umma_desc_mn_sw128(vs + kk * 2048, Cfg::kKVSub, 1024);
            ptx::umma_bf16(tO, da, db, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          }
          ptx::umma_commit(pv_done + 8 * w);
          if (w == nw - 1) ptx::umma_commit(kv_empty + 8 * s);   // both PV(j) issued: K/V stage reusable when they retire
        }
      }
    }
  } else {
    // =========================== softmax warpgroups ===========================
    const int w = (warp - 2) >> 2;          // 0 = A, 1 = B
    if (w < nw) {
      const int qd = warp & 3;
      const int row = qd * 32 + lane;
      const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
      float m_used = -INFINITY, l_sum = 0.f;
      uint8_t* p_gen = smem_gen + (p_smem - smem_base) + (size_t)w * Cfg::kPBytes;
      const uint32_t tS = tmem + lane_off + Cfg::kSCol + w * BN;
      const uint32_t tO = tmem + lane_off + Cfg::kOCol + w * D;

      // trace slots (per CTA, per warpgroup): [0] wait s_full, [1] TMEM load, [2] max, [3] wait pv_done, [4] rescale,
      // [5] exp + P store, [6] tiles, [7] total
#ifdef B200LDM_ATTN_TRACE
      const bool tr = a.trace != nullptr && lane == 0 && qd == 0;
#else
      constexpr bool tr = false;   // the clock64 bookkeeping costs registers (spills): compiled in only for tracing
#endif
      long long tc[6] = {0, 0, 0, 0, 0, 0};
      const long long t_begin = tr ? clock64() : 0;
      for (int j = 0; j < ntiles; ++j) {
        long long t0 = tr ? clock64() : 0, t1;
        ptx::mbar_wait(s_full + 8 * w, (uint32_t)j & 1u, 20);
        ptx::tc_fence_after();
        if (tr) { t1 = clock64(); tc[0] += t1 - t0; t0 = t1; }
        float sv[BN];
        {
          // all of S_w(j) in flight before the single wait (four serialised ~230-cycle TMEM round trips before)
          uint32_t r[BN / 32][32];
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) ptx::tmem_ld32(tS + c * 32, r[c]);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < BN / 32; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) sv[c * 32 + i] = __uint_as_float(r[c][i]);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(s_free + 8 * w);   // the tensor core may overwrite S_w with the next tile's scores
        if (tr) { t1 = clock64(); tc[1] += t1 - t0; t0 = t1; }
        const int kvalid = a.Lk - j * BN;
        if (kvalid < BN) {
#pragma unroll
          for (int i = 0; i < BN; ++i)
            if (i >= kvalid) sv[i] = -INFINITY;
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < BN; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], sv[i + u]);
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * a.scale_log2;
        const bool need = mx > m_used + 8.0f;
        const bool any_need = __any_sync(0xffffffffu, need);
        float alpha = 1.0f;
        if (tr) { t1 = clock64(); tc[2] += t1 - t0; t0 = t1; }
        if (j > 0) ptx::mbar_wait(pv_done + 8 * w, (uint32_t)(j - 1) & 1u, 21);  // PV_w(j-1) retired
        if (tr) { t1 = clock64(); tc[3] += t1 - t0; t0 = t1; }
        if (any_need) {
          if (need) {
            alpha = exp2f(m_used - mx);
            m_used = mx;
            l_sum *= alpha;
          }
          if (j > 0) {
            ptx::tc_fence_after();
#pragma unroll
            for (int c = 0; c < D; c += 16) {
              uint32_t r[16];
              ptx::tmem_ld16(tO + c, r);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              ptx::tmem_st16(tO + c, r);
            }
            ptx::tmem_st_wait();
          }
        }
        if (tr) { t1 = clock64(); tc[4] += t1 - t0; t0 = t1; }
        const float neg_m = -m_used;
#ifdef B200LDM_ATTN_TRACE
        const bool no_mufu = (a.dbg & 1) != 0;
#else
        constexpr bool no_mufu = false;
#endif
        float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c8 = 0; c8 < BN / 8; ++c8) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x0 = fmaf(sv[c8 * 8 + 2 * e], a.scale_log2, neg_m);
            const float x1 = fmaf(sv[c8 * 8 + 2 * e + 1], a.scale_log2, neg_m);
            const float p0 = no_mufu ? x0 * 0.001f : ex2_approx(x0);
            const float p1 = no_mufu ? x1 * 0.001f : ex2_approx(x1);
            ps4[e] += p0 + p1;
            const __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
            pk[e] = *reinterpret_cast<const uint32_t*>(&hb);
          }
          const int sub = c8 >> 3, cc = c8 & 7;
          uint4* dst = reinterpret_cast<uint4*>(p_gen + sub * 16384 + row * 128 + ((cc ^ (row & 7)) << 4));
          *dst = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        l_sum += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        ptx::mbar_arrive(p_full + 8 * w);
        if (tr) { t1 = clock64(); tc[5] += t1 - t0; }
      }
      if (tr) {
        const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        if (cta < 1024) {
          unsigned long long* o = a.trace + (cta * 2 + w) * 8;
          for (int i = 0; i < 6; ++i) o[i] = (unsigned long long)tc[i];
          o[6] = (unsigned long long)ntiles;
          o[7] = (unsigned long long)(clock64() - t_begin);
        }
      }
      // ---- epilogue
      ptx::mbar_wait(pv_done + 8 * w, (uint32_t)(ntiles - 1) & 1u, 22);
      ptx::tc_fence_after();
      const int qi = q0 + w * 128 + row;
      const float inv = 1.0f / l_sum;
      __nv_bfloat16* orow = a.o + ((size_t)b * a.L + (qi < a.L ? qi : 0)) * ((size_t)a.H * D) + col0;
#pragma unroll
      for (int c = 0; c < D; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(tO + c, r);
        ptx::tmem_ld_wait();
        if (qi < a.L) {
          float f[8];
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[hlf * 8 + e]) * inv;
            Vec8<__nv_bfloat16>::store(orow + c + hlf * 8, f);
          }
        }
      }
      ptx::tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem);
  }
}

// ------------------------------------------------------------------------------------------ host
static PFN_cuTensorMapEncodeTiled_v12000 attn_encode_fn() {
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

static int make_qkv_map(CUtensorMap* m, const void* base, int cols, int L, int B, long long ld, int box_rows) {
  auto enc = attn_encode_fn();
  if (!enc) return fail(B200LDM_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)L};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B200LDM_BAD_SHAPE, "cuTensorMapEncodeTiled(qkv) failed: %d cols=%d L=%d B=%d ld=%lld", (int)r, cols, L, B,
                ld);
  return B200LDM_OK;
}

// ---- split-KV of the last, partial wave (one-query-tile kernels).  T query tiles on `slots` = SMs x CTAs/SM: the first
// floor(T / slots) * slots tiles run whole; the rest (rem < slots) would be one more wave with SMs idle, so each of them
// is cut into `nsplit` key ranges (rem * nsplit parts = a few short rounds) and merged by attention_merge_kernel.
struct AttnSplit {
  int t0, nsplit;   // nsplit <= 1: nothing is split
};
static int g_attn_split = [] {
  const char* e = getenv("B200LDM_ATTN_SPLIT");   // 0: never split; n >= 2: force that factor on the partial wave
  return e ? atoi(e) : -1;
}();
void set_attn_split(int s) { g_attn_split = s; }   // tests (b200ldm_debug_set_attn_split); -1 = planner
static AttnSplit plan_attn_split(long long T, int slots, int ntiles_kv) {
  AttnSplit sp{(int)T, 1};
  if (g_attn_split == 0 || T > (1ll << 30)) return sp;
  const long long full = T / slots, rem = T - full * slots;
  if (rem == 0) return sp;
  // time in key-tile units, calibrated on the AutoencoderKL attention (d = 256, L = 13824 / 32768 / 64000 with and without
  // the split, profiles/r02_attention/split_kv.log): a CTA costs its key tiles + ~32 more (prologue, Q load, pipeline
  // ramp, epilogue), the merge launch ~10
  const double kOvh = 32.0, kMerge = 10.0;
  const double now = (double)(full + 1) * (ntiles_kv + kOvh);
  double best = now;
  for (int s = 2; s <= 8; ++s) {
    if (ntiles_kv / s < 1) break;
    if (g_attn_split >= 2 && s != g_attn_split) continue;
    const double rounds = (double)((rem * s + slots - 1) / slots);
    const double t = (double)full * (ntiles_kv + kOvh) + rounds * ((double)ntiles_kv / s + kOvh) + kMerge;
    if (t < best || (g_attn_split >= 2 && sp.nsplit == 1)) { best = t; sp.nsplit = s; }
  }
  if (g_attn_split < 2 && best > now * 0.96) sp.nsplit = 1;   // not worth a second launch below 4 %
  if (sp.nsplit > 1) sp.t0 = (int)(full * slots);
  return sp;
}
static long long attn_split_workspace_bytes(long long T, const AttnSplit& sp, int dr) {
  if (sp.nsplit <= 1) return 0;
  return (T - sp.t0) * sp.nsplit * ((long long)dr * 128 * 4 + 2 * 128 * 4);
}

// out[q][c] = sum_p O_p[q][c] 2^(m_p - m) / sum_p l_p 2^(m_p - m),  m = max_p m_p.  grid = (split tiles, column slices):
// a CTA merges 128 rows x (dr / 4 / gridDim.y) column quads, thread = (column quad, row) like the partial layout.  (One
// CTA per tile ran at 0.9 TB/s - 61 us for the 56 MB of a 128^3 reconstruct, ncu r02af; the partial tiles are read once,
// so the launch just needs enough CTAs in flight.)
__global__ void attention_merge_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                       __nv_bfloat16* __restrict__ o, int t0, int nsplit, int tiles_x, int H, int L, int dr) {
  const int tile = t0 + (int)blockIdx.x;
  const int q0 = (tile % tiles_x) * 128, h = (tile / tiles_x) % H, b = tile / (tiles_x * H);
  const size_t unit0 = (size_t)blockIdx.x * nsplit;
  __shared__ float w[8][128];   // per part and row: 2^(m_p - m) / l
  for (int row = threadIdx.x; row < 128; row += blockDim.x) {
    float m = -INFINITY;
    for (int p = 0; p < nsplit; ++p) m = fmaxf(m, part_ml[(unit0 + p) * 256 + row]);
    float l = 0.f;
    for (int p = 0; p < nsplit; ++p) {
      const float e = exp2f(part_ml[(unit0 + p) * 256 + row] - m);
      w[p][row] = e;
      l += part_ml[(unit0 + p) * 256 + 128 + row] * e;
    }
    const float inv = 1.0f / l;
    for (int p = 0; p < nsplit; ++p) w[p][row] *= inv;
  }
  __syncthreads();
  const int quads = dr / 4;
  const int qper = (quads + (int)gridDim.y - 1) / (int)gridDim.y, q_lo = (int)blockIdx.y * qper;
  const int q_hi = q_lo + qper < quads ? q_lo + qper : quads;
  for (int i = q_lo * 128 + threadIdx.x; i < q_hi * 128; i += blockDim.x) {
    const int row = i & 127, cq = i >> 7;
    if (q0 + row >= L) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < nsplit; ++p) {
      const float4 f = __ldcg(reinterpret_cast<const float4*>(part_o) + ((unit0 + p) * (size_t)quads + cq) * 128 + row);
      const float ww = w[p][row];
      acc.x = fmaf(f.x, ww, acc.x); acc.y = fmaf(f.y, ww, acc.y); acc.z = fmaf(f.z, ww, acc.z); acc.w = fmaf(f.w, ww, acc.w);
    }
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(o + ((size_t)b * L + q0 + row) * ((size_t)H * dr) + (size_t)h * dr + cq * 4);
    dst[0] = __floats2bfloat162_rn(acc.x, acc.y);
    dst[1] = __floats2bfloat162_rn(acc.z, acc.w);
  }
}

template <int D, int BN, int DR = D>
static int launch_attention_tc(const void* q, const void* k, const void* v, void* o, int B, int H, int L, int Lk,
                               long long ld, long long ldkv, float scale, void* ws, long long ws_bytes, cudaStream_t st) {
  using Cfg = AttnCfg<D, BN>;
  AttnMaps maps;
  int rc;
  if ((rc = make_qkv_map(&maps.q, q, H * DR, L, B, ld, 128))) return rc;
  if ((rc = make_qkv_map(&maps.k, k, H * DR, Lk, B, ldkv, BN))) return rc;
  if ((rc = make_qkv_map(&maps.v, v, H * DR, Lk, B, ldkv, BN))) return rc;
  static const int dbg = [] { const char* e = getenv("B200LDM_ATTN_DBG"); return e ? atoi(e) : 0; }();
  static const int stagger = [] { const char* e = getenv("B200LDM_ATTN_STAGGER"); return e ? atoi(e) : kAttnStaggerDefault; }();
  AttnArgs a{H, L, DR, Lk, scale * 1.4426950408889634f, reinterpret_cast<__nv_bfloat16*>(o), g_attn_trace, dbg, stagger};
  a.tiles_x = ceil_div(L, 128);
  const long long T = (long long)a.tiles_x * H * B;
  B2_REQUIRE(T < (1ll << 31) / 8, B200LDM_BAD_SHAPE, "attention: too many query tiles (%lld)", T);
  AttnSplit sp = plan_attn_split(T, sm_count() * Cfg::kOcc, ceil_div(Lk, BN));
  if (!ws || ws_bytes < attn_split_workspace_bytes(T, sp, DR)) sp = AttnSplit{(int)T, 1};   // no scratch: whole tiles
  a.split_t0 = sp.t0; a.nsplit = sp.nsplit;
  if (sp.nsplit > 1) {
    a.part_o = reinterpret_cast<float*>(ws);
    a.part_ml = a.part_o + (size_t)(T - sp.t0) * sp.nsplit * DR * 128;
  }
  // B200LDM_ATTN_PACKED: 0 = scalar softmax loop, 1 = two-wide fp32 instructions, 2 .. 5 = + 1 .. 4 of every 8 score
  // pairs through the polynomial exp2 (head dim 32 / 64 only: the wider heads are tensor-bound); default: measured best
  static const int packed = [] { const char* e = getenv("B200LDM_ATTN_PACKED"); return e ? atoi(e) : kAttnPackedDefault; }();
  dim3 grid((unsigned)(sp.t0 + (T - sp.t0) * sp.nsplit), 1, 1);
  // the merge pass of the split tiles, queued behind the attention kernel by every return path below
  auto merge = [&]() -> int {
    if (sp.nsplit > 1) {
      attention_merge_kernel<<<dim3((unsigned)(T - sp.t0), (unsigned)(DR >= 64 ? 8 : 4)), 256, 0, st>>>(
          a.part_o, a.part_ml, a.o, sp.t0, sp.nsplit, a.tiles_x, H, L, DR);
      B2_LAUNCH_CHECK();
    }
    return B200LDM_OK;
  };
  // B200LDM_ATTN_PTMEM: 1 = P through TMEM (A operand of the PV MMA from TMEM), 0 = P through shared memory;
  // B200LDM_ATTN_SB: 1 = single S buffer + P through TMEM, three CTAs per SM (head dim 32 / 64 with 64-key tiles only)
  static const int ptmem = [] { const char* e = getenv("B200LDM_ATTN_PTMEM"); return e ? atoi(e) : kAttnPTmemDefault; }();
  static const int sbuf = [] { const char* e = getenv("B200LDM_ATTN_SB"); return e ? atoi(e) : kAttnSBDefault; }();
  // B200LDM_ATTN_QTMEM: 1 = Q in TMEM as well (S issued with the A operand in TMEM); needs PTMEM, 64-column kernels
  static const int qtmem = [] { const char* e = getenv("B200LDM_ATTN_QTMEM"); return e ? atoi(e) : kAttnQTmemDefault; }();
  // B200LDM_ATTN_SMEM_PAD=<bytes>: extra dynamic shared memory per CTA (experiments: forces fewer CTAs per SM)
  static const int smem_pad = [] { const char* e = getenv("B200LDM_ATTN_SMEM_PAD"); return e ? atoi(e) : 0; }();
#define B2_ATTN_LAUNCH(MODE, PTM, SBF)                                                                         \
  do {                                                                                                         \
    using LCfg = AttnCfg<D, BN, SBF>;                                                                          \
    B2_SET_SMEM_ONCE(LCfg::kSmem + smem_pad, attention_tc_kernel<D, BN, DR, MODE, PTM, SBF>);                  \
    attention_tc_kernel<D, BN, DR, MODE, PTM, SBF><<<grid, kAttnThreads, LCfg::kSmem + smem_pad, st>>>(maps, a); \
  } while (0)
  if constexpr (D == 64 && BN == 64) {
    if (sbuf) {
      switch (packed) {
        case 0: B2_ATTN_LAUNCH(0, true, true); break;
        case 3: B2_ATTN_LAUNCH(3, true, true); break;
        case 4: B2_ATTN_LAUNCH(4, true, true); break;
        default: B2_ATTN_LAUNCH(1, true, true); break;
      }
      B2_LAUNCH_CHECK();
      return merge();
    }
  }
  // B200LDM_ATTN_TWO_ISSUERS: 0 = one MMA-issuing thread everywhere, 1 = two for head dim 256 (measured +10 %), 2 = two for
  // the head-dim-32 / 64 kernel with P and Q in TMEM as well
  static const int two_d64 = [] { const char* e = getenv("B200LDM_ATTN_TWO_ISSUERS"); return e ? atoi(e) : kAttnTwoIssuersDefault; }();
  // B200LDM_ATTN_W2=1: two softmax warpgroups per CTA (each thread owns half of a tile's keys) for the head-dim-32 / 64
  // kernel with P and Q in TMEM
  static const int w2 = [] { const char* e = getenv("B200LDM_ATTN_W2"); return e ? atoi(e) : kAttnW2Default; }();
  if constexpr (D == 64 && BN == 64) {
    if (ptmem && qtmem && w2) {
      using QCfg = AttnCfg<D, BN, false>;
      constexpr int kSm = QCfg::kSmem + 3072;   // + the exchange area of the two halves
      static_assert(kSm <= 232448 / QCfg::kOcc, "smem budget with the exchange area");
      B2_SET_SMEM_ONCE(kSm + smem_pad, attention_tc_kernel<D, BN, DR, 1, true, false, true, false, true>);
      attention_tc_kernel<D, BN, DR, 1, true, false, true, false, true><<<grid, kAttnThreads + 128, kSm + smem_pad, st>>>(maps, a);
      B2_LAUNCH_CHECK();
      return merge();
    }
  }
  if constexpr (D == 64 && BN == 64) {
    if (ptmem && qtmem && two_d64 >= 2) {
      using QCfg = AttnCfg<D, BN, false>;
      B2_SET_SMEM_ONCE(QCfg::kSmem + smem_pad, attention_tc_kernel<D, BN, DR, 1, true, false, true, true>);
      attention_tc_kernel<D, BN, DR, 1, true, false, true, true><<<grid, kAttnThreads + 32, QCfg::kSmem + smem_pad, st>>>(maps, a);
      B2_LAUNCH_CHECK();
      return merge();
    }
  }
  if constexpr (D == 64 && BN == 64) {
    if (ptmem && qtmem) {
      using QCfg = AttnCfg<D, BN, false>;
      B2_SET_SMEM_ONCE(QCfg::kSmem + smem_pad, attention_tc_kernel<D, BN, DR, 1, true, false, true>);
      attention_tc_kernel<D, BN, DR, 1, true, false, true><<<grid, kAttnThreads, QCfg::kSmem + smem_pad, st>>>(maps, a);
      B2_LAUNCH_CHECK();
      return merge();
    }
  }
  if constexpr (D == 64) {
    if (ptmem) {
      switch (packed) {
        case 0: B2_ATTN_LAUNCH(0, true, false); break;
        case 3: B2_ATTN_LAUNCH(3, true, false); break;
        case 4: B2_ATTN_LAUNCH(4, true, false); break;
        case 5: B2_ATTN_LAUNCH(5, true, false); break;
        default: B2_ATTN_LAUNCH(1, true, false); break;
      }
    } else {
      switch (packed) {
        case 0: B2_ATTN_LAUNCH(0, false, false); break;
        case 3: B2_ATTN_LAUNCH(3, false, false); break;
        case 4: B2_ATTN_LAUNCH(4, false, false); break;
        case 5: B2_ATTN_LAUNCH(5, false, false); break;
        default: B2_ATTN_LAUNCH(1, false, false); break;
      }
    }
  } else {
    // wider heads: P (and Q) in TMEM when the columns are there - head dim 256 with 64-key tiles uses all 512
    // (S 2 x 64, O 256, Q 128); B200LDM_ATTN_WIDE_TMEM=0 keeps both operands in shared memory
    static const int wide = [] { const char* e = getenv("B200LDM_ATTN_WIDE_TMEM"); return e ? atoi(e) : kAttnWideTmemDefault; }();
    if constexpr (AttnCfg<D, BN, false>::kOCol + D + D / 2 <= AttnCfg<D, BN, false>::kTmemCols) {
      if (wide == 2) {
        using QCfg = AttnCfg<D, BN, false>;
        B2_SET_SMEM_ONCE(QCfg::kSmem + smem_pad, attention_tc_kernel<D, BN, DR, 1, true, false, true>);
        attention_tc_kernel<D, BN, DR, 1, true, false, true><<<grid, kAttnThreads, QCfg::kSmem + smem_pad, st>>>(maps, a);
        B2_LAUNCH_CHECK();
        return merge();
      }
    }
    // head dim 256: two MMA-issuing threads (B200LDM_ATTN_TWO_ISSUERS=0: one) - the single thread was the bound
    static const int two = [] { const char* e = getenv("B200LDM_ATTN_TWO_ISSUERS"); return e ? atoi(e) : kAttnTwoIssuersDefault; }();
    if constexpr (D == 256) {
      if (two && w2 && wide == 0 && packed != 0) {   // two issuing threads AND two softmax warpgroups
        using TCfg = AttnCfg<D, BN, false>;
        constexpr int kSm = TCfg::kSmem + (TCfg::kSmem + 3072 <= 232448 / TCfg::kOcc ? 3072 : 1024);
        static_assert(kSm <= 232448 / TCfg::kOcc, "smem budget with the exchange area");
        B2_SET_SMEM_ONCE(kSm + smem_pad, attention_tc_kernel<D, BN, DR, 1, false, false, false, true, true>);
        attention_tc_kernel<D, BN, DR, 1, false, false, false, true, true>
            <<<grid, kAttnThreads + 32 + 128, kSm + smem_pad, st>>>(maps, a);
        B2_LAUNCH_CHECK();
        return merge();
      }
    }
    if constexpr (D == 256) {
      if (two && wide == 0 && packed != 0) {
        using TCfg = AttnCfg<D, BN, false>;
        B2_SET_SMEM_ONCE(TCfg::kSmem + smem_pad, attention_tc_kernel<D, BN, DR, 1, false, false, false, true>);
        attention_tc_kernel<D, BN, DR, 1, false, false, false, true>
            <<<grid, kAttnThreads + 32, TCfg::kSmem + smem_pad, st>>>(maps, a);
        B2_LAUNCH_CHECK();
        return merge();
      }
    }
    if (wide >= 1) { if (packed == 0) B2_ATTN_LAUNCH(0, true, false); else B2_ATTN_LAUNCH(1, true, false); }
    else { if (packed == 0) B2_ATTN_LAUNCH(0, false, false); else B2_ATTN_LAUNCH(1, false, false); }
  }
#undef B2_ATTN_LAUNCH
  B2_LAUNCH_CHECK();
  return merge();
}

template <int D, int BN, int OCC>
static int launch_attention_tc2(const void* q, const void* k, const void* v, void* o, int B, int H, int L, int Lk,
                                long long ld, long long ldkv, float scale, cudaStream_t st) {
  using Cfg = Attn2Cfg<D, BN, OCC>;
  AttnMaps maps;
  int rc;
  if ((rc = make_qkv_map(&maps.q, q, H * D, L, B, ld, 128))) return rc;
  if ((rc = make_qkv_map(&maps.k, k, H * D, Lk, B, ldkv, BN))) return rc;
  if ((rc = make_qkv_map(&maps.v, v, H * D, Lk, B, ldkv, BN))) return rc;
  static const int dbg = [] { const char* e = getenv("B200LDM_ATTN_DBG"); return e ? atoi(e) : 0; }();
  AttnArgs a{H, L, D, Lk, scale * 1.4426950408889634f, reinterpret_cast<__nv_bfloat16*>(o), g_attn_trace, dbg, 0};
  B2_SET_SMEM_ONCE(Cfg::kSmem, attention_tc2_kernel<D, BN, OCC>);
  dim3 grid((unsigned)ceil_div(L, 256), (unsigned)H, (unsigned)B);
  attention_tc2_kernel<D, BN, OCC><<<grid, Cfg::kThreads, Cfg::kSmem, st>>>(maps, a);
  B2_LAUNCH_CHECK();
  return B200LDM_OK;
}

// B200LDM_ATTN_VERSION: 0 (default) = measured best per head dim; 1 = one-query-tile kernels everywhere; 2 = the
// two-query-tile ping-pong kernel wherever it exists (d = 64 / 128, L > 128)
static const int g_attn_v = [] {
  const char* e = getenv("B200LDM_ATTN_VERSION");
  return e ? atoi(e) : 0;
}();

bool attention_tcgen05_supported(int d) { return d == 32 || d == 64 || d == 128 || d == 256; }

// scratch that lets the launch split the key range of its last, partial wave of query tiles (0: nothing to gain)
long long attention_tcgen05_workspace_bytes(int B, int H, int L, int Lk, int d) {
  if (!attention_tcgen05_supported(d) || B <= 0 || H <= 0 || L <= 0 || Lk <= 0) return 0;
  if (d == 128 && L > 128 && Lk > 128 && g_attn_v != 1) return 0;   // ping-pong kernel: no split
  const long long T = (long long)ceil_div(L, 128) * H * B;
  const int bn = (d == 256 || Lk > 128) ? 64 : 128;
  if (d == 128) return 0;
  const int occ = (d <= 64 && bn == 64) ? 2 : 1;
  const AttnSplit sp = plan_attn_split(T, sm_count() * occ, ceil_div(Lk, bn));
  return attn_split_workspace_bytes(T, sp, d);
}

int attention_tcgen05(const void* q, const void* k, const void* v, void* o, int B, int H, int L, int Lk, int d,
                      long long ld, long long ldkv, float scale, void* ws, long long ws_bytes, cudaStream_t st) {
  B2_REQUIRE(attention_tcgen05_supported(d), B200LDM_BAD_SHAPE,
             "tcgen05 attention supports head dim 32/64/128/256 (got %d)", d);
  B2_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && ld % 8 == 0 && ldkv % 8 == 0,
             B200LDM_BAD_ALIGN, "attention: pointers / row strides must be 16-byte aligned");
  // head dim 64: the one-query-tile kernel with 64-key tiles runs TWO CTAs per SM (96 KB smem, 256 TMEM columns each)
  // and beats the two-tile ping-pong kernel by 7-19 % (0.406 vs 0.436 ms at B8 H4 L5544): its two softmax warps per
  // sub-partition belong to independent CTAs and never fall into lock-step.  Head dim 128 keeps the ping-pong kernel.
  // Cross attention (Lk = context length, usually one short tile) always takes the one-query-tile kernel.
  const bool pingpong = L > 128 && Lk > 128 && ((d == 128 && g_attn_v != 1) || (d == 64 && g_attn_v == 2));
  if (pingpong) {
    // d = 64: B200LDM_ATTN_OCC=2 runs 64-key tiles with two CTAs per SM (four softmax warps per sub-partition)
    static const int occ = [] { const char* e = getenv("B200LDM_ATTN_OCC"); return e ? atoi(e) : 1; }();
    if (d == 64 && occ == 2) return launch_attention_tc2<64, 64, 2>(q, k, v, o, B, H, L, Lk, ld, ldkv, scale, st);
    if (d == 64) return launch_attention_tc2<64, 128, 1>(q, k, v, o, B, H, L, Lk, ld, ldkv, scale, st);
    return launch_attention_tc2<128, 64, 1>(q, k, v, o, B, H, L, Lk, ld, ldkv, scale, st);
  }
  if (d == 32 && Lk > 128) return launch_attention_tc<64, 64, 32>(q, k, v, o, B, H, L, Lk, ld, ldkv, scale, ws, ws_bytes, st);
  if (d == 32) return launch_attention_tc<64, 128, 32>(q, k, v, o, B, H, L, Lk, ld, ldkv, scale, ws, ws_bytes, st);
  if (d == 64 && Lk > 128) return launch_attention_tc<64, 64>(q, k, v, o, B, H, L, Lk, ld, ldkv, scale, ws, ws_bytes, st);
  if (d == 64) return launch_attention_tc<64, 128>(q, k, v, o, B, H, L, Lk, ld, ldkv, scale, ws, ws_bytes, st);
  if (d == 128) return launch_attention_tc<128, 128>(q, k, v, o, B, H, L, Lk, ld, ldkv, scale, nullptr, 0, st);
  return launch_attention_tc<256, 64>(q, k, v, o, B, H, L, Lk, ld, ldkv, scale, ws, ws_bytes, st);
}

}  // namespace b200ldm
