// conv_tc.cuh - parameter blocks shared by the tcgen05 conv kernels (conv_tc.cu: one CTA per tile;
// conv_tc2.cu: persistent CTA pairs).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200ldm {

struct ConvTcMaps {
  CUtensorMap a[8];  // stride 1: a[0] = x, a[1] = x2;  stride 2: a[pd*4+ph*2+pw]
  CUtensorMap w;     // [taps][Cout_pad][Cin_total]
  CUtensorMap y[8];  // output, same voxel box as the A tiles (TMA-store epilogue); up mode: one strided map per phase
  CUtensorMap r;     // residual (same geometry as y) when fused
  CUtensorMap s[2];  // fused 1x1x1 skip conv: its one or two (concat) sources, output geometry
  CUtensorMap sw;    // skip weights [1][Cout_pad][Cs1+Cs2]
};

struct ConvTcArgs {
  int tn, td, th, tw;      // voxel box of one M tile
  int nb_d, nb_h, nb_w;    // tiles per spatial dim (tiles over N = gridDim.x / (nb_d*nb_h*nb_w))
  int N, Do, Ho, Wo, Cout;
  int nchunk1, nchunk2;    // 64-channel chunks in x / x2
  int nsk1, nsk2;          // 64-channel chunks of the fused 1x1x1 skip conv's sources (extra K iterations, centre tap)
  int ksize, stride, pad_lo;
  int rows_box;            // tn*td*th*tw
  int temb_stride, out_f32, gn_groups;
  const float* bias;
  const float* temb;
  const void* residual;
  void* y;
  double* gn_stats;
  double* chunk_stats;        // [N][Cout/cw][2] (sum, sumsq) per cw-channel chunk of y, accumulated by the TMA epilogue
  int cw;                     // chunk width: 8, or 2 / 4 (conv_tcp.cu / conv_halo.cu with Cout tiles <= 128 only)
  int up;                     // 1: nearest-x2 upsample fused as 8 phase convs of 2x2x2 taps (blockIdx.z = phase)
  int splits;                 // split-K: blockIdx.z handles k-iterations [z*I/S, (z+1)*I/S) and writes fp32 slab z
  long long slab_elems;       // elements per fp32 slab (N*Do*Ho*Wo*Cout)
  int ashift;                 // experiment: A sub-tile base shifted by ashift*128 bytes (swizzle-phase probe)
  int tma_epilogue;           // 1: stage the tile in smem and TMA-store it (bf16, Cout % 64 == 0, no GN partials)
  int slab_tma;               // 1: split-K partial sums leave as fp32 32-channel groups through smem + TMA stores (maps.y[0]
                              //    is an fp32 map over all slabs, sample coordinate = z * N + n)
  unsigned long long* trace;  // debug: per-CTA clock64 stamps (b200ldm_debug_set_conv_trace), else nullptr
  int work_x, work_y, work_z; // conv_tc2: M pair-tiles, Cout tiles, upsample phases of the persistent work list
  void* sk_ws;                // conv_sk.cu: workspace = [flags, sk_flag_bytes][one 256 KB partial-tile slot per CTA]
  int sk_flag_bytes;
};

constexpr int kABytes = 128 * 128;  // one M sub-tile: 128 rows x 64 bf16

// conv_tc2.cu: persistent 2-CTA kernel (cta_group::2 UMMA, double-buffered TMEM, overlapped TMA-store epilogue)
bool conv_tc2_supported(int bn, const ConvTcArgs& a);
int launch_conv_tc2(const ConvTcMaps& maps, const ConvTcArgs& a, int bn, int m_tiles, int n_tiles, int zdim,
                    cudaStream_t st);


// conv_tcp.cu: persistent single-CTA kernel with two accumulator sets (epilogue of tile i overlaps the main loop of
// tile i+1); MT sub-tiles per CTA = conv_tcp_mt(bn)
bool conv_tcp_supported(int bn, const ConvTcArgs& a);
int conv_tcp_mt(int bn);
// wide_epilogue (Cout tile 256 only): two epilogue warp sets drain alternate 64-column groups (kernel choice 4)
int launch_conv_tcp(const ConvTcMaps& maps, const ConvTcArgs& a, int bn, int m_tiles, int n_tiles, int zdim,
                    cudaStream_t st, bool wide_epilogue = false);

// conv_halo.cu: persistent kernel for 3x3x3 stride-1 convs with Cout tiles of 64 / 128: W-halo activation boxes
// (one TMA load serves the three kw taps through shifted UMMA views); tile = 1 x td x th x 8 voxels, td * th = 16
bool conv_halo_supported(int bn, const ConvTcArgs& a);
void choose_halo_tile(int Do, int Ho, int* td, int* th);
int launch_conv_halo(const ConvTcMaps& maps, const ConvTcArgs& a, int bn, int m_tiles, int n_tiles, cudaStream_t st);

// conv_sk.cu: stream-K kernel for the 256 x 256 tile (one CTA per SM runs an equal share of ALL k-iterations; partial
// tiles are exchanged through the caller's workspace inside the launch - no finalize pass)
bool conv_sk_supported(int bn, const ConvTcArgs& a);
int conv_sk_grid(int m_tiles, int n_tiles, int iters);
long long conv_sk_workspace_bytes(int grid);
int launch_conv_sk(const ConvTcMaps& maps, ConvTcArgs a, int m_tiles, int n_tiles, int iters, void* ws, long long ws_bytes,
                   cudaStream_t st);

}  // namespace b200ldm
