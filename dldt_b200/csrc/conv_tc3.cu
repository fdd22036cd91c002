// CTA-pair implicit-GEMM convolution / GEMM on tcgen05 (sm_100a), version 3: cta_group::2, 256 x BN tiles.
//
//   D[M, Cout] = im2col(X)[M, K] * Wp[Cout, K]^T  (+ bias, activation, + residual, activation)
//
// Same arithmetic contract as conv_tc2.cu (ngraph reference/convolution.hpp:89-186; post-op chain of
// nodes/mkldnn_conv_node.cpp:355-403).  What changes is the operand traffic per FLOP.  Every ResNet-50 layer at batch 256 is bound
// by the L2 -> SM fabric (about 6.3 KB/clk chip-wide, 42 B/clk per SM), not by the tensor pipe: a 128 x 128 tile needs 128 B/clk
// per SM to keep the MMA busy (profiles/README.md).  Here two CTAs of a cluster (the two SMs of a TPC) compute one 256 x BN tile
// with ONE tcgen05.mma.cta_group::2 per 16 K elements issued by the leader: each CTA stages its own 128 activation rows and only
// HALF of the weight tile, so a 256 x 256 tile moves 64 B/clk per SM.  Three more traffic cuts:
//   * weight-stationary mode: a layer with a single N tile keeps its half of the packed weights resident in shared memory
//   * the epilogue works one 64-channel slab at a time through small staging rings (2 output slabs, 4 residual slabs) instead of
//     whole-tile double buffers, which is what lets BN = 256 (512 TMEM columns, two accumulators) fit beside the operand rings
//   * "row" mode for 3x3 / stride-1 convolutions: the A operand of filter row r is ONE tiled TMA box {64 ch, P pixels, 128/P rows}
//     (P = power of two >= W + 2) and the three taps of that row are three MMAs whose descriptors start 0 / 128 / 256 bytes into the
//     box (one pixel = one 128-byte swizzled row), so the activations cross the fabric 3x instead of 9x.  Output rows with
//     w >= W are computed and dropped by a 3-D store map.
//
//   warp 0-7  epilogue (TMEM lane quarter = warp % 4, column half = warp / 4)   warp 8   MMA issuer (leader CTA, one lane) + TMEM alloc
//   warp 9    operand TMA issuer (A ring, B ring, resident weights)              warp 10  residual TMA issuer
//
// mbarriers: fullA/emptyA[s], fullB/emptyB[s] (operand rings; full* live in the leader and collect both CTAs' bytes),
// tfull[a] (multicast commit), tempty[a] (leader; both CTAs' epilogue warps arrive), rfull/rempty[slot] (local), wfull (leader).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "act.cuh"
#include "common.h"
#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tc_common.cuh"

namespace b200 {

// defined in conv_tc2.cu
CUtensorMap make_tmap_2d(const void* base, bool bf16, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer);
CUtensorMap make_tmap_im2col(const ConvParams& p, bool bf16);
PFN_cuTensorMapEncodeTiled_v12000 encode_tiled_fn();
tc::EpiAct make_epi(int kind, float a, float b);

namespace {

using namespace tc;

constexpr int BM = 128;                  // rows per CTA; the pair tile is 256 rows
constexpr int BK = 64;
constexpr int A_BYTES = BM * 128;
constexpr int SLAB_BYTES = BM * 128;     // one 64-channel slab of the output / residual tile of one CTA
constexpr int NUM_EPI_WARPS = 8;
constexpr int MMA_WARP = 8, TMA_WARP = 9, RES_WARP = 10;
constexpr int NUM_THREADS = 352;
constexpr int MAX_STAGES = 8;
constexpr int NOUT = 2, NRES = 4;        // staging rings, in slabs
constexpr int SMEM_LIMIT = 232448;
constexpr int WS_MAX_BYTES = 80 * 1024;  // resident half of the weights, per CTA

enum AMode : int { A_MATRIX = 1, A_IM2COL = 2, A_ROWS = 3, A_HALO = 4 };

struct Plan3 { int sa, sb; uint32_t a_stage, b_stage, a_off, b_off, out_off, res_off, w_off, bar_off, total; };

template <int BN>
__host__ __device__ inline Plan3 plan3(bool has_res, int ws_bytes, int a_mode, int rows_p) {
  Plan3 pl;
  pl.a_stage = (uint32_t)(A_BYTES + (a_mode == A_ROWS ? 1024 : 0));   // row mode reads up to 2 pixel rows past the box
  if (a_mode == A_HALO) pl.a_stage = (uint32_t)((BM + 2 * rows_p) * 128 + 1024);   // R + 2 image rows of P pixels, same overrun
  pl.b_stage = (uint32_t)((BN / 2) * 128);
  const int fixed = NOUT * SLAB_BYTES + (has_res ? NRES * SLAB_BYTES : 0) + ws_bytes + 1024 /*align slack*/ + 1024 /*barriers*/;
  const int budget = SMEM_LIMIT - fixed;
  if (ws_bytes) {
    pl.sb = 0;
    pl.sa = budget / (int)pl.a_stage;
  } else if (a_mode == A_ROWS || a_mode == A_HALO) {
    pl.sa = a_mode == A_ROWS ? 3 : 2;
    pl.sb = (budget - pl.sa * (int)pl.a_stage) / (int)pl.b_stage;
  } else {
    pl.sa = pl.sb = budget / (int)(pl.a_stage + pl.b_stage);
  }
  if (pl.sa > MAX_STAGES) pl.sa = MAX_STAGES;
  if (pl.sb > MAX_STAGES) pl.sb = MAX_STAGES;
  pl.a_off = 0;
  pl.b_off = pl.a_off + (uint32_t)pl.sa * pl.a_stage;
  pl.out_off = pl.b_off + (uint32_t)pl.sb * pl.b_stage;
  pl.res_off = pl.out_off + NOUT * SLAB_BYTES;
  pl.w_off = pl.res_off + (has_res ? NRES * SLAB_BYTES : 0);
  pl.bar_off = pl.w_off + (uint32_t)ws_bytes;
  pl.total = pl.bar_off + 1024u + 1024u;
  return pl;
}

struct Tc3Args {
  CUtensorMap tm_a;   // A_MATRIX: [M][Cin_p]; A_IM2COL: {C, W, H, N} im2col; A_ROWS: {C, W, H, N} tiled, box {64, P, 128/P, 1}
  CUtensorMap tm_w;   // packed weights [Cout_alloc][Ktot], box {64, BN/2}
  CUtensorMap tm_y;   // output: [M][Cout_p] box {64, 128}; A_ROWS: {C, W, N*H} box {64, P, 128/P}
  CUtensorMap tm_r;   // residual, same geometry as the output
  ConvParams p;
  EpiAct e1, e2;
  int a_mode;
  int num_tiles, n_tiles;   // pair tiles
  int ws;
  int reverse;
  int rows_p, rows_r;       // A_ROWS: pixel pitch P and image rows per CTA tile R = 128 / P
  int tiles_per_img;        // A_ROWS: CTA tiles per image = ceil(Ho / R)
  int desc_base_offset;     // probe knob (B200_TC3_BASEOFF=1): fill the descriptor's base-offset field for the shifted row-mode reads
};

template <typename T, bool GEN>
__device__ __forceinline__ void epilogue_chunk(uint32_t taddr, const float* __restrict__ bias, uint32_t out_row, uint32_t res_row, uint32_t rsw,
                                               int chunk0, bool has_res, const EpiAct& e1, const EpiAct& e2) {
  uint32_t v[32];
  ptx::tmem_ld_32x32(taddr, v);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + g * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + g * 8 + 4));
    float f[8];
    f[0] = __uint_as_float(v[g * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
    f[2] = __uint_as_float(v[g * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
    f[4] = __uint_as_float(v[g * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
    f[6] = __uint_as_float(v[g * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
    epi_act8<GEN>(f, e1);
    const uint32_t sw = (rsw ^ (uint32_t)(chunk0 + g)) << 4;      // 16-byte chunk index XOR (row & 7)
    if (has_res) {
      const uint4 rv = ptx::ld_shared_v4(res_row + sw);
      float r[8];
      unpack2<T>(rv.x, r[0], r[1]); unpack2<T>(rv.y, r[2], r[3]);
      unpack2<T>(rv.z, r[4], r[5]); unpack2<T>(rv.w, r[6], r[7]);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += r[e];
      epi_act8<GEN>(f, e2);
    }
    ptx::st_shared_v4(out_row + sw, make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7])));
  }
}

// first output row of CTA `rank` of pair tile `tile_m`, and (row mode) its image / first image row
struct TilePos { int m0; int img, h0; bool valid; };
__device__ __forceinline__ TilePos tile_pos(const Tc3Args& a, int tile_m, int rank) {
  TilePos t;
  const int ct = tile_m * 2 + rank;                 // CTA tile index along M
  if (a.a_mode == A_ROWS || a.a_mode == A_HALO) {
    t.img = ct / a.tiles_per_img;
    t.h0 = (ct - t.img * a.tiles_per_img) * a.rows_r;
    t.valid = t.img < a.p.N;
    t.m0 = 0;
  } else {
    t.m0 = ct * BM;
    t.valid = t.m0 < a.p.M;
    t.img = 0; t.h0 = 0;
  }
  return t;
}

template <typename T, int BN, bool GEN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) conv_tc3_kernel(const __grid_constant__ Tc3Args args) {
  constexpr int B_BYTES = (BN / 2) * 128;     // this CTA's half of a weight k-block
  constexpr int SLABS = BN / 64;
  const ConvParams& p = args.p;
  const bool has_res = p.res != nullptr;
  const int nkb = p.Ktot / BK;
  const bool ws = args.ws != 0;
  const int a_mode = args.a_mode;
  const Plan3 plan = plan3<BN>(has_res, ws ? nkb * B_BYTES : 0, a_mode, args.rows_p);
  const int SA = plan.sa, SB = plan.sb;
  const int cpt = p.Cin_p >> 6;                               // 64-channel chunks per filter tap
  const int reuse = a_mode == A_ROWS ? p.kw : (a_mode == A_HALO ? p.kh * p.kw : 1);   // k-blocks served by one A stage
  const int n_ablk = a_mode == A_ROWS ? p.kh * cpt : (a_mode == A_HALO ? cpt : nkb);
  const uint32_t a_tx = a_mode == A_HALO ? (uint32_t)((BM + 2 * args.rows_p) * 128) : (uint32_t)A_BYTES;   // bytes one CTA's A load delivers

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem_al + plan.bar_off);
  uint64_t* emptyA = fullA + MAX_STAGES;
  uint64_t* fullB = emptyA + MAX_STAGES;
  uint64_t* emptyB = fullB + MAX_STAGES;
  uint64_t* tfull = emptyB + MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* rfull = tempty + 2;
  uint64_t* rempty = rfull + NRES;
  uint64_t* wfull = rempty + NRES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = args.num_tiles, n_tiles = args.n_tiles;
  const int cluster = (int)ptx::cluster_id_x(), nclusters = (int)ptx::num_clusters_x();

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      ptx::mbar_init(&fullA[s], 2);      // leader's expect_tx arrive + the peer's remote arrive
      ptx::mbar_init(&emptyA[s], 1);
      ptx::mbar_init(&fullB[s], 2);
      ptx::mbar_init(&emptyB[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 2 * NUM_EPI_WARPS);
    }
    for (int r = 0; r < NRES; ++r) {
      ptx::mbar_init(&rfull[r], 1);
      ptx::mbar_init(&rempty[r], NUM_EPI_WARPS);
    }
    ptx::mbar_init(wfull, 2);
    ptx::fence_mbar_init();
  }
  if (warp == TMA_WARP && lane == 0) {
    ptx::tma_prefetch_desc(&args.tm_w);
    ptx::tma_prefetch_desc(&args.tm_a);
    ptx::tma_prefetch_desc(&args.tm_y);
    if (has_res) ptx::tma_prefetch_desc(&args.tm_r);
  }
  if (warp == MMA_WARP) {
    ptx::tmem_alloc_pair(tmem_slot, 2 * BN);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();                   // barriers of both CTAs are initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NUM_EPI_WARPS) {
    // ===================== epilogue: this CTA's 128 rows, one 64-channel slab at a time =====================
    const int q = warp & 3;
    const int half = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t rsw = (uint32_t)(row & 7);
    const bool store_thread = threadIdx.x == 0;
    const uint32_t tempty_leader0 = ptx::mapa(ptx::smem_u32(&tempty[0]), 0), tempty_leader1 = ptx::mapa(ptx::smem_u32(&tempty[1]), 0);
    uint32_t it = 0, g = 0;                // tiles / slabs processed so far
    for (int t_ = cluster; t_ < num_tiles; t_ += nclusters, ++it) {
      const int tile = args.reverse ? num_tiles - 1 - t_ : t_;
      const TilePos tp = tile_pos(args, tile / n_tiles, (int)rank);
      const int n0 = (tile % n_tiles) * BN;
      const int ncols = min(BN, p.Cout_p - n0);
      const int nslab = (ncols + 63) >> 6;
      const uint32_t acc = it & 1u;
      ptx::mbar_wait(&tfull[acc], (it >> 1) & 1u);
      ptx::tc_fence_after();
      for (int h = 0; h < nslab; ++h, ++g) {
        const uint32_t oslot = g % NOUT, rslot = g % NRES;
        const uint32_t out_slab = smem_base + plan.out_off + oslot * SLAB_BYTES;
        const uint32_t out_row = out_slab + (uint32_t)row * 128u;
        const uint32_t res_row = smem_base + plan.res_off + rslot * SLAB_BYTES + (uint32_t)row * 128u;
        if (has_res) ptx::mbar_wait(&rfull[rslot], (g / NRES) & 1u);
        if (store_thread) ptx::bulk_wait_read<NOUT - 1>();        // the store that last used this output slab has finished reading it
        ptx::named_bar_sync(1, NUM_EPI_WARPS * 32);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)(h * 64 + half * 32);
        epilogue_chunk<T, GEN>(taddr, p.bias + n0 + h * 64 + half * 32, out_row, res_row, rsw, half * 4, has_res, args.e1, args.e2);
        if (has_res) {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&rempty[rslot]);
        }
        ptx::fence_proxy_async();                                 // staging writes -> visible to the TMA engine
        ptx::named_bar_sync(1, NUM_EPI_WARPS * 32);
        if (store_thread) {
          if (tp.valid) {
            if (a_mode == A_ROWS || a_mode == A_HALO) ptx::tma_store_3d(&args.tm_y, out_slab, n0 + h * 64, 0, tp.img * p.Ho + tp.h0);
            else ptx::tma_store_2d(&args.tm_y, out_slab, n0 + h * 64, tp.m0);
          }
          ptx::bulk_commit();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);   // accumulator drained (both CTAs report to the leader)
    }
    if (store_thread) ptx::bulk_wait_all0();
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(TcType<T>::is_bf16, 2 * BM, BN);
      uint32_t it = 0;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      if (ws) ptx::mbar_wait(wfull, 0);
      for (int t_ = cluster; t_ < num_tiles; t_ += nclusters, ++it) {
        const uint32_t acc = it & 1u;
        ptx::mbar_wait(&tempty[acc], ((it >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t first = 0;
        for (int ab = 0; ab < n_ablk; ++ab) {
          ptx::mbar_wait(&fullA[sa], pha);
          const uint32_t a_addr = smem_base + plan.a_off + (uint32_t)sa * plan.a_stage;
          for (int j = 0; j < reuse; ++j) {
            // weight k-block that goes with (A block ab, tap j): K order of the packed weights is (filter row, filter col, channel)
            const int kb = a_mode == A_ROWS ? ((ab / cpt) * p.kw + j) * cpt + (ab % cpt) : (a_mode == A_HALO ? j * cpt + ab : ab);
            // start of the shifted view: one pixel = one 128-byte swizzled row; halo mode also steps whole image rows (P pixels)
            const uint32_t a_view = a_addr + (a_mode == A_HALO ? (uint32_t)((j / p.kw) * args.rows_p + (j % p.kw)) : (uint32_t)j) * 128u;
            uint32_t b_addr;
            if (ws) {
              b_addr = smem_base + plan.w_off + (uint32_t)kb * B_BYTES;
            } else {
              ptx::mbar_wait(&fullB[sb], phb);
              b_addr = smem_base + plan.b_off + (uint32_t)sb * plan.b_stage;
            }
            ptx::tc_fence_after();
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              uint64_t da = ptx::make_smem_desc_sw128(a_view + k * 32);
              if (args.desc_base_offset) da |= (uint64_t)((a_view >> 7) & 7u) << 49;
              const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * 32);
              ptx::umma_f16_pair(d_tmem, da, db, idesc, first | (uint32_t)k);
            }
            first = 1;
            if (!ws) {
              ptx::umma_commit_pair(&emptyB[sb], 3);
              if (++sb == SB) { sb = 0; phb ^= 1u; }
            }
          }
          ptx::umma_commit_pair(&emptyA[sa], 3);
          if (++sa == SA) { sa = 0; pha ^= 1u; }
        }
        ptx::umma_commit_pair(&tfull[acc], 3);
      }
    }
  } else if (warp == TMA_WARP) {
    // ===================== operand TMA issuer (one thread per CTA) =====================
    if (lane == 0) {
      const uint32_t fullA_leader = ptx::smem_u32(&fullA[0]) & ptx::PEER_BIT_MASK;
      const uint32_t fullB_leader = ptx::smem_u32(&fullB[0]) & ptx::PEER_BIT_MASK;
      const uint32_t wfull_leader = ptx::smem_u32(wfull) & ptx::PEER_BIT_MASK;
      const int HoWo = p.Ho * p.Wo;
      if (ws) {   // this CTA's half of the whole weight matrix, once
        if (leader) ptx::mbar_arrive_expect_tx(wfull, 2u * (uint32_t)nkb * B_BYTES);
        else ptx::mbar_arrive_cluster(wfull_leader);
        for (int kb = 0; kb < nkb; ++kb)
          ptx::tma_load_2d_pair(smem_base + plan.w_off + (uint32_t)kb * B_BYTES, &args.tm_w, wfull_leader, kb * BK, (int)rank * (BN / 2));
      }
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      for (int t_ = cluster; t_ < num_tiles; t_ += nclusters) {
        const int tile = args.reverse ? num_tiles - 1 - t_ : t_;
        const int tile_m = tile / n_tiles;
        TilePos tp = tile_pos(args, tile_m, (int)rank);
        if (!tp.valid) tp = tile_pos(args, tile_m, 0);            // tail: stage the leader's rows again (never stored)
        const int n0 = (tile % n_tiles) * BN + (int)rank * (BN / 2);
        int bn = 0, bh = 0, bw = 0;
        if (a_mode == A_IM2COL) {
          bn = tp.m0 / HoWo;
          const int rem = tp.m0 - bn * HoWo;
          const int ho = rem / p.Wo;
          bh = ho * p.sh - p.pt;
          bw = (rem - ho * p.Wo) * p.sw - p.pl;
        }
        int chunk = 0, fr = 0, fs = 0;
        for (int ab = 0; ab < n_ablk; ++ab) {
          ptx::mbar_wait(&emptyA[sa], pha ^ 1u);
          const uint32_t a_dst = smem_base + plan.a_off + (uint32_t)sa * plan.a_stage;
          const uint32_t bar = fullA_leader + (uint32_t)sa * 8u;
          if (leader) ptx::mbar_arrive_expect_tx(&fullA[sa], 2u * a_tx);
          else ptx::mbar_arrive_cluster(bar);
          if (a_mode == A_MATRIX) {
            ptx::tma_load_2d_pair(a_dst, &args.tm_a, bar, ab * BK, tp.m0);
          } else if (a_mode == A_IM2COL) {
            ptx::tma_load_im2col_4d_pair(a_dst, &args.tm_a, bar, chunk * BK, bw, bh, bn, (uint32_t)(fs * p.dw), (uint32_t)(fr * p.dh));
            if (++chunk == cpt) { chunk = 0; if (++fs == p.kw) { fs = 0; ++fr; } }
          } else if (a_mode == A_HALO) {
            // all filter rows at once: input rows h0 - pt .. h0 - pt + R + 1 of 64-channel chunk ab
            ptx::tma_load_4d_pair(a_dst, &args.tm_a, bar, ab * BK, -p.pl, tp.h0 - p.pt, tp.img);
          } else {
            // filter row fr: input rows h0 + fr - pt .., pixels -pl .. P - pl - 1 (out-of-image parts are zero-filled)
            ptx::tma_load_4d_pair(a_dst, &args.tm_a, bar, chunk * BK, -p.pl, tp.h0 + fr - p.pt, tp.img);
            if (++chunk == cpt) { chunk = 0; ++fr; }
          }
          if (!ws) {
            for (int j = 0; j < reuse; ++j) {
              const int kb = a_mode == A_ROWS ? ((ab / cpt) * p.kw + j) * cpt + (ab % cpt) : (a_mode == A_HALO ? j * cpt + ab : ab);
              ptx::mbar_wait(&emptyB[sb], phb ^ 1u);
              const uint32_t barb = fullB_leader + (uint32_t)sb * 8u;
              if (leader) ptx::mbar_arrive_expect_tx(&fullB[sb], 2u * B_BYTES);
              else ptx::mbar_arrive_cluster(barb);
              ptx::tma_load_2d_pair(smem_base + plan.b_off + (uint32_t)sb * plan.b_stage, &args.tm_w, barb, kb * BK, n0);
              if (++sb == SB) { sb = 0; phb ^= 1u; }
            }
          }
          if (++sa == SA) { sa = 0; pha ^= 1u; }
        }
      }
    }
  } else if (warp == RES_WARP) {
    // ===================== residual TMA issuer: one slab ahead of the epilogue, ring of NRES =====================
    if (has_res && lane == 0) {
      uint32_t g = 0;
      for (int t_ = cluster; t_ < num_tiles; t_ += nclusters) {
        const int tile = args.reverse ? num_tiles - 1 - t_ : t_;
        const TilePos tp = tile_pos(args, tile / n_tiles, (int)rank);
        const int n0 = (tile % n_tiles) * BN;
        const int ncols = min(BN, p.Cout_p - n0);
        const int nslab = (ncols + 63) >> 6;
        for (int h = 0; h < nslab; ++h, ++g) {
          const uint32_t rslot = g % NRES;
          ptx::mbar_wait(&rempty[rslot], ((g / NRES) & 1u) ^ 1u);
          const uint32_t dst = smem_base + plan.res_off + rslot * SLAB_BYTES;
          if (tp.valid) {
            ptx::mbar_arrive_expect_tx(&rfull[rslot], SLAB_BYTES);
            if (a_mode == A_ROWS || a_mode == A_HALO) ptx::tma_load_3d(dst, &args.tm_r, &rfull[rslot], n0 + h * 64, 0, tp.img * p.Ho + tp.h0);
            else ptx::tma_load_2d(dst, &args.tm_r, &rfull[rslot], n0 + h * 64, tp.m0);
          } else {
            ptx::mbar_arrive(&rfull[rslot]);                      // rows beyond M: the epilogue reads stale staging and stores nothing
          }
        }
      }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  ptx::cluster_sync();        // neither CTA may exit (or free TMEM) while the other can still touch its shared memory / barriers
  if (warp == MMA_WARP) ptx::tmem_dealloc_pair(tmem_base, 2 * BN);
}

// ---- host side ------------------------------------------------------------------------------------------
int g_sm_count3 = 0;

// rank-3/4 tiled maps over NHWC activations (row mode)
CUtensorMap make_tmap_tiled(const void* base, bool bf16, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  auto enc = encode_tiled_fn();
  B200_CHECK(enc, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap m;
  cuuint64_t d[5]; cuuint64_t st[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = enc(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled (rank " << rank << ") failed (" << (int)r << ")");
  return m;
}

int rows_pitch(const ConvParams& p) {
  int P = 8;
  while (P < p.W + p.kw - 1) P <<= 1;
  return P;
}

template <int BN>
bool ws_applies(const ConvParams& p, int a_mode) {
  const int ws_bytes = (p.Ktot / BK) * (BN / 2) * 128;
  const char* no_ws = getenv("B200_NO_WS");
  const int min_sa = a_mode == A_HALO ? 2 : 4;
  return p.Cout_alloc == BN && ws_bytes <= WS_MAX_BYTES && plan3<BN>(p.res != nullptr, ws_bytes, a_mode, rows_pitch(p)).sa >= min_sa &&
         !(no_ws && no_ws[0] == '1');
}

template <typename T, int BN>
void launch_tc3(const ConvParams& p, bool bf16, int a_mode, cudaStream_t s) {
  Tc3Args a;
  memset(&a, 0, sizeof(a));
  a.p = p;
  a.e1 = make_epi(p.act1, p.a1, p.b1);
  a.e2 = make_epi(p.act2, p.a2, p.b2);
  a.a_mode = a_mode;
  a.n_tiles = p.Cout_alloc / BN;
  a.reverse = p.reverse;
  int cta_tiles_m;
  if (a_mode == A_ROWS || a_mode == A_HALO) {
    a.rows_p = rows_pitch(p);
    a.rows_r = BM / a.rows_p;
    a.tiles_per_img = (p.Ho + a.rows_r - 1) / a.rows_r;
    cta_tiles_m = a.tiles_per_img * p.N;
  } else {
    cta_tiles_m = (p.M + BM - 1) / BM;
  }
  a.num_tiles = ((cta_tiles_m + 1) / 2) * a.n_tiles;
  a.ws = ws_applies<BN>(p, a_mode) ? 1 : 0;
  { const char* e = getenv("B200_TC3_BASEOFF"); a.desc_base_offset = (e && e[0] == '1') ? 1 : 0; }
  a.tm_w = make_tmap_2d(p.w, bf16, (uint64_t)p.Ktot, (uint64_t)p.Cout_alloc, (uint64_t)p.w_ld * 2, BK, BN / 2);
  if (a_mode == A_ROWS || a_mode == A_HALO) {
    const uint64_t C = (uint64_t)p.Cin_p, pitch = (uint64_t)p.ldy * 2;
    const uint64_t ad[4] = {C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    const uint64_t as[3] = {C * 2, (uint64_t)p.W * C * 2, (uint64_t)p.H * p.W * C * 2};
    const uint32_t ab[4] = {64, (uint32_t)a.rows_p, (uint32_t)(a.rows_r + (a_mode == A_HALO ? 2 : 0)), 1};
    a.tm_a = make_tmap_tiled(p.x, bf16, 4, ad, as, ab);
    // output / residual: {channels, w, n*h}; rows (w) beyond Wo and image rows beyond Ho of the LAST image are clipped.  A tile never
    // straddles two images because tiles_per_img * R >= Ho rows are walked per image and rows >= Ho ... see the store coordinate:
    // img * Ho + h0 with h0 + R possibly > Ho -> those rows belong to the next image, so the box is shortened by a per-tile map below
    const uint64_t yd[3] = {(uint64_t)p.Cout_p, (uint64_t)p.Wo, (uint64_t)p.N * p.Ho};
    const uint64_t ys[2] = {pitch, (uint64_t)p.Wo * pitch};
    const uint32_t yb[3] = {64, (uint32_t)a.rows_p, (uint32_t)a.rows_r};
    a.tm_y = make_tmap_tiled(p.y, bf16, 3, yd, ys, yb);
    a.tm_r = p.res ? make_tmap_tiled(p.res, bf16, 3, yd, ys, yb) : a.tm_y;
  } else {
    a.tm_y = make_tmap_2d(p.y, bf16, (uint64_t)p.Cout_p, (uint64_t)p.M, (uint64_t)p.ldy * 2, 64, BM);
    a.tm_r = p.res ? make_tmap_2d(p.res, bf16, (uint64_t)p.Cout_p, (uint64_t)p.M, (uint64_t)p.ldy * 2, 64, BM) : a.tm_y;
    a.tm_a = a_mode == A_MATRIX ? make_tmap_2d(p.x, bf16, (uint64_t)p.Cin_p, (uint64_t)p.M, (uint64_t)p.Cin_p * 2, BK, BM) : make_tmap_im2col(p, bf16);
  }
  const int sms = g_sm_count3 > 0 ? g_sm_count3 : 148;
  const int grid = 2 * std::min(a.num_tiles, sms / 2);
  const unsigned smem = plan3<BN>(p.res != nullptr, a.ws ? (p.Ktot / BK) * (BN / 2) * 128 : 0, a_mode, a.rows_p).total;
  if (conv_acts_are_clamps(p)) conv_tc3_kernel<T, BN, false><<<grid, NUM_THREADS, smem, s>>>(a);
  else conv_tc3_kernel<T, BN, true><<<grid, NUM_THREADS, smem, s>>>(a);
}

}  // namespace

// Row mode applies to dense 3x3 (any kh, kw <= 3) stride-1, dilation-1 "same"-padded convolutions whose rows fit a 64-pixel box and
// whose height is a multiple of the rows per tile (so that a tile never straddles two images).
bool conv_tc3_rows_ok(const ConvParams& p) {
  const char* e = getenv("B200_TC3_ROWS");
  if (e && e[0] == '0') return false;
  if (!(p.kw == 3 && p.kh >= 1 && p.kh <= 3 && p.sh == 1 && p.sw == 1 && p.dh == 1 && p.dw == 1 && p.pl == 1 && p.Wo == p.W && p.Ho == p.H &&
        p.pt == (p.kh - 1) / 2 && p.Cin_p % 64 == 0 && p.x_sw == 0))
    return false;
  int P = 8;
  while (P < p.W + 2) P <<= 1;
  if (P > 64) return false;
  const int R = BM / P;
  return p.Ho % R == 0 && p.ldy == p.Cout_p;
}

int conv_tc3_a_mode(const ConvParams& p) {
  if (conv_tc2_a_is_matrix(p) && p.Cin_p % 64 == 0) return A_MATRIX;
  if (conv_tc3_rows_ok(p)) {
    const char* e = getenv("B200_TC3_HALO");
    const bool halo = p.kh == 3 && p.pt == 1 && (e && e[0] == '1');
    return halo ? A_HALO : A_ROWS;
  }
  if (conv_tc2_a_is_im2col(p)) return A_IM2COL;
  return 0;
}

bool conv_tc3_supported(const ConvParams& p) {
  return conv_tc2_supported(p) && conv_tc3_a_mode(p) != 0 && p.Cout_alloc % 64 == 0 && p.Ktot % BK == 0;
}

// Estimated cycles of one launch with N tile `bn` on `sms` SMs: rounds x k-blocks x max(MMA, operand bytes / fabric rate) + epilogue traffic.
static double tc3_cost(const ConvParams& p, int bn, int a_mode, int sms) {
  const double fabric = 42.0;                          // B/clk per SM (6.3 KB/clk chip-wide)
  int cta_m;
  double a_scale = 1.0;
  if (a_mode == A_ROWS || a_mode == A_HALO) {
    int P = 8; while (P < p.W + 2) P <<= 1;
    cta_m = ((p.Ho + BM / P - 1) / (BM / P)) * p.N;
    a_scale = a_mode == A_ROWS ? 1.0 / p.kw : (BM / P + 2.0) / (BM / P) / (p.kh * p.kw);
  } else {
    cta_m = (p.M + BM - 1) / BM;
  }
  const int tiles = ((cta_m + 1) / 2) * (p.Cout_alloc / bn);
  const int rounds = (tiles + sms / 2 - 1) / (sms / 2);
  const int nkb = p.Ktot / BK;
  const bool ws = p.Cout_alloc == bn && nkb * (bn / 2) * 128 <= WS_MAX_BYTES;
  const double bytes_kb = A_BYTES * a_scale + (ws ? 0.0 : (bn / 2) * 128.0);
  const double mma_kb = bn * 2.0;                      // 128 x bn x 64 per CTA at 4096 MAC/clk
  const double epi = (p.res ? 2.0 : 1.0) * BM * bn * 2.0 / fabric;
  return rounds * (nkb * std::max(mma_kb, bytes_kb / fabric) + epi);
}

int conv_tc3_pick_bn(const ConvParams& p) {
  const int a_mode = conv_tc3_a_mode(p);
  const int sms = g_sm_count3 > 0 ? g_sm_count3 : 148;
  int best = 0;
  double best_c = 0;
  for (int bn : {256, 128, 64}) {
    if (p.Cout_alloc % bn) continue;
    const double c = tc3_cost(p, bn, a_mode, sms);
    if (!best || c < best_c) { best = bn; best_c = c; }
  }
  return best;
}

// auto policy: the pair kernel pays off once every SM pair has a tile to walk
bool conv_tc3_preferred(const ConvParams& p) {
  const char* e = getenv("B200_TC3");
  if (e && e[0] == '0') return false;
  if (!conv_tc3_supported(p)) return false;
  if (e && e[0] == '1') return true;
  const int bn = conv_tc3_pick_bn(p);
  const int sms = g_sm_count3 > 0 ? g_sm_count3 : 148;
  const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * (p.Cout_alloc / bn);
  return tiles >= sms / 2;
}

const char* conv_tc3_mode_name(const ConvParams& p) {
  switch (conv_tc3_a_mode(p)) {
    case A_MATRIX: return "tma";
    case A_ROWS: return "rows";
    case A_HALO: return "halo";
    case A_IM2COL: return "im2col";
  }
  return "?";
}

void conv_tc3_init_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sm_count3, cudaDevAttrMultiProcessorCount, dev);
  auto opt_in = [](auto kernel) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT); };
  opt_in(conv_tc3_kernel<__nv_bfloat16, 64, false>);  opt_in(conv_tc3_kernel<__nv_bfloat16, 64, true>);
  opt_in(conv_tc3_kernel<__nv_bfloat16, 128, false>); opt_in(conv_tc3_kernel<__nv_bfloat16, 128, true>);
  opt_in(conv_tc3_kernel<__nv_bfloat16, 256, false>); opt_in(conv_tc3_kernel<__nv_bfloat16, 256, true>);
  opt_in(conv_tc3_kernel<__half, 64, false>);  opt_in(conv_tc3_kernel<__half, 64, true>);
  opt_in(conv_tc3_kernel<__half, 128, false>); opt_in(conv_tc3_kernel<__half, 128, true>);
  opt_in(conv_tc3_kernel<__half, 256, false>); opt_in(conv_tc3_kernel<__half, 256, true>);
}

void launch_conv_tc3(const ConvParams& p, int dtype, int bn, cudaStream_t s) {
  B200_CHECK(conv_tc3_supported(p), B200_ERR_UNSUPPORTED, "conv_tc3: unsupported launch");
  B200_CHECK(p.Cout_alloc % bn == 0 && p.Cout_p % 8 == 0 && p.ldy % 8 == 0, B200_ERR_INVALID, "conv_tc3: bad geometry");
  const bool bf16 = dtype == B200_BF16;
  B200_CHECK(bf16 || dtype == B200_F16, B200_ERR_UNSUPPORTED, "conv_tc3: dtype must be bf16/f16");
  const int a_mode = conv_tc3_a_mode(p);
  if (bf16) {
    if (bn == 64) launch_tc3<__nv_bfloat16, 64>(p, true, a_mode, s);
    else if (bn == 128) launch_tc3<__nv_bfloat16, 128>(p, true, a_mode, s);
    else launch_tc3<__nv_bfloat16, 256>(p, true, a_mode, s);
  } else {
    if (bn == 64) launch_tc3<__half, 64>(p, false, a_mode, s);
    else if (bn == 128) launch_tc3<__half, 128>(p, false, a_mode, s);
    else launch_tc3<__half, 256>(p, false, a_mode, s);
  }
}

}  // namespace b200
