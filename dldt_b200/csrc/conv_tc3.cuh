// CTA-pair implicit-GEMM convolution / GEMM kernel on tcgen05 (sm_100a): the device code of conv_tc3 (see conv_tc3.cu for the
// host side and the design notes).  Included by conv_tc3.cu (bf16 instantiations) and conv_tc3_f16.cu (fp16 instantiations).
#pragma once
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
namespace tc3 {

using namespace tc;

constexpr int BM = 128;                  // rows per CTA; the pair tile is 256 rows
constexpr int BK = 64;
constexpr int A_BYTES = BM * 128;
constexpr int SLAB_BYTES = BM * 128;     // one 64-channel slab of the output / residual tile of one CTA
constexpr int NUM_EPI_WARPS = 8;
constexpr int MMA_WARP = 8, TMA_WARP = 9, RES_WARP = 10;
constexpr int NUM_THREADS = 352;
constexpr int MAX_STAGES = 8;
constexpr int MAX_ACC = 4;               // TMEM accumulator ring (512 columns / accumulator columns, at most this)
constexpr int NRES = 4;                  // residual ring, in slabs; output staging: 1 or 2 slabs per epilogue group (Plan3::nout = 2 or 4)
constexpr int SMEM_LIMIT = 232448;
constexpr int WS_MAX_BYTES = 144 * 1024; // resident half of the weights, per CTA (144 KB = the 3x3 x 128 -> 128 layer at N tile 128, beside two halo stages)
constexpr int BIAS_SMEM_FLOATS = 4096;   // the bias vector is staged in shared memory (launches with a larger padded Cout are not supported)

enum AMode : int { A_MATRIX = 1, A_IM2COL = 2, A_HALO = 4, A_S2D = 8 };
constexpr int S2D_R = 4, S2D_G = 32;     // space-to-depth stem tile: 4 output rows x 32 groups of 4 pixels (one group = 4 x 16 channels = one 128-byte row)

struct Plan3 { int sa, sb, nout; uint32_t a_stage, b_stage, a_off, b_off, out_off, res_off, w_off, bias_off, bar_off, total; };

// Shared-memory plan.  rows_p = pixel pitch P of the halo box (A_HALO only).
__host__ __device__ inline int bias_smem_bytes(int cout_alloc) { return (cout_alloc * 4 + 1023) & ~1023; }
__host__ __device__ inline Plan3 plan3(int bn, bool has_res, int ws_bytes, int a_mode, int rows_p, int cout_alloc, int nkb) {
  Plan3 pl;
  const int bias_bytes = bias_smem_bytes(cout_alloc);
  pl.a_stage = (uint32_t)A_BYTES;
  // R + 2 image rows of P pixels.  The last taps of MMA rows 126/127 read 2 box rows past the stage: whatever follows it in shared memory
  // (the next stage, the weight ring, the staging slabs) - those rows are pixel slots P - 2 / P - 1 >= W of the last image row, never stored
  if (a_mode == A_HALO) pl.a_stage = (uint32_t)((BM + 2 * rows_p) * 128);
  if (a_mode == A_S2D) pl.a_stage = (uint32_t)((S2D_R + 3) * S2D_G * 128 + 1024);   // R + 3 block rows of 32 pixel groups; the last taps read 1 row past
  pl.b_stage = (uint32_t)((bn / 2) * 128);
  // Two staging slabs per epilogue group (the next slab is converted while the TMA store of the previous one still reads its slab) for
  // launches with few k-blocks per tile, where the epilogue is what bounds the tile time, if that leaves three operand stages
  const int fixed2 = 2 * SLAB_BYTES + (has_res ? NRES * SLAB_BYTES : 0) + ws_bytes + bias_bytes + 1024 /*align slack*/ + 1024 /*barriers*/;
  const int rem4 = SMEM_LIMIT - fixed2 - 2 * SLAB_BYTES;
  pl.nout = (nkb <= 9 && rem4 >= 3 * (int)(pl.a_stage + (ws_bytes ? 0 : pl.b_stage))) ? 4 : 2;
  const int fixed = fixed2 + (pl.nout - 2) * SLAB_BYTES;
  const int budget = SMEM_LIMIT - fixed;
  if (ws_bytes) {
    pl.sb = 0;
    pl.sa = budget / (int)pl.a_stage;
    if (a_mode == A_S2D && pl.sa > 4) pl.sa = 4;
  } else if (a_mode == A_S2D) {
    pl.sb = 0;
    pl.sa = 4;
  } else if (a_mode == A_HALO) {
    // one A stage feeds nine k-blocks: the weights are the stream, the activations need depth only to hide one TMA latency
    pl.sb = bn == 256 ? 4 : 6;
    pl.sa = (budget - pl.sb * (int)pl.b_stage) / (int)pl.a_stage;
    if (pl.sa > 4) { pl.sa = 4; pl.sb = (budget - pl.sa * (int)pl.a_stage) / (int)pl.b_stage; }
  } else {
    pl.sa = pl.sb = budget / (int)(pl.a_stage + pl.b_stage);
  }
  if (pl.sa > MAX_STAGES) pl.sa = MAX_STAGES;
  if (pl.sb > MAX_STAGES) pl.sb = MAX_STAGES;
  pl.a_off = 0;
  pl.b_off = pl.a_off + (uint32_t)pl.sa * pl.a_stage;
  pl.out_off = pl.b_off + (uint32_t)pl.sb * pl.b_stage;
  pl.res_off = pl.out_off + (uint32_t)pl.nout * SLAB_BYTES;
  pl.w_off = pl.res_off + (has_res ? NRES * SLAB_BYTES : 0);
  pl.bias_off = pl.w_off + (uint32_t)ws_bytes;
  pl.bar_off = pl.bias_off + (uint32_t)bias_bytes;
  pl.total = pl.bar_off + 1024u + 1024u;
  return pl;
}

struct Tc3Args {
  CUtensorMap tm_a;   // A_MATRIX: [M][Cin_p]; A_IM2COL: {C, W, H, N} im2col; A_HALO: {C, W, H, N} tiled, box {64, P, 128/P + 2, 1}
  CUtensorMap tm_w;   // packed weights [Cout_alloc][Ktot], box {64, BN/2}
  CUtensorMap tm_y;   // output: [M][Cout_p] box {64, 128}; A_HALO: {C, W, H, N} box {64, P, 128/P, 1}
  CUtensorMap tm_r;   // residual, same geometry as the output
  CUtensorMap tm_yw;  // output, one warp's 32 rows per box (warp_store)
  int warp_store;     // every epilogue warp stores its own 32 rows (no block-level barrier per slab)
  CUtensorMap tm_a2;  // A_MATRIX with a second source (ConvParams::x2): im2col-mode map {K2, W2, H2, N}, 1x1 window, stride (sw2, sh2)
  int nkb1;           // k-blocks served by tm_a; the rest come from tm_a2 (== Ktot / 64 when there is no second source)
  ConvParams p;
  EpiAct e1, e2;
  int gen;                  // activations outside the clamp family: generic table in the epilogue (selects the GEN instantiation)
  int num_tiles, n_tiles;   // pair tiles
  int reverse;
  int rows_p, rows_r;       // A_HALO: pixel pitch P and image rows per CTA tile R = 128 / P
  int nseg;                 // A_S2D: 128-pixel segments per output row
  int tiles_per_img;        // A_S2D: ceil(Ho / 4) * nseg.  A_HALO: CTA tiles per image = ceil(Ho / R); the last one may hang over the image (clipped by the 4-D maps)
  int acc_log2;             // log2 of the TMEM accumulator ring depth (1 = two accumulators; up to 512 / accumulator columns, at most 4)
  int early_release;        // epilogue warps hand the accumulator back right after their last tcgen05.ld of the tile
  int pool;                 // A_S2D: fused 3x3 / stride-2 / pad-1 max pool, tm_y = pooled tensor {C, Wo / 2, Ho / 2, N}, box {64, 64, 1, 1}
  int debug;                // timing experiments only (B200_DEBUG_TC3; results are wrong): bit 0 = no output stores, bit 1 = no epilogue arithmetic, bit 2 = no programmatic-launch wait
  int prefetch_tiles;       // L2 prefetch distance (tiles of this cluster ahead) for the residual and the activation-matrix loads; 0 = off
  uint64_t a_policy, r_policy;   // L2 eviction priority of the activation-matrix loads (A_MATRIX) and of the residual loads
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// One 64-column slab of the clamp-family epilogue (none / ReLU / Clamp activations: every ResNet / MobileNet / SSD layer), as ONE straight
// line of code per (residual, clamp, trailing-ReLU) combination: + bias, [clamp], [+ residual, [clamp]], pack, swizzled staging store.
// The first version tested the activation kinds per 8 values (warp-uniform branches), which cut the slab into ~30 basic blocks: every
// shared-memory load sat directly in front of its first use and the epilogue warps spent 7 cycles per instruction (ncu, ResNet-50 stage-4
// expand layer: 31 % short-scoreboard, 21 % fixed-latency, 18 % instruction-fetch stalls, 2200 cycles per slab against ~350 instructions).
// Here the loads of chunk c + 1 (two bias vectors, one residual vector) are issued before the arithmetic of chunk c.
template <typename T, bool HAS_RES, bool CLAMP, bool RELU>
__device__ __forceinline__ void epilogue_slab(const uint32_t (&v0)[32], const uint32_t (&v1)[32], uint32_t bias /* shared-memory address */, uint32_t out_row,
                                              uint32_t res_row, uint32_t rsw, const EpiBounds& eb) {
  uint4 b0 = ptx::ld_shared_v4(bias), b1 = ptx::ld_shared_v4(bias + 16u);   // all lanes read the same address: broadcast
  uint4 rv = make_uint4(0u, 0u, 0u, 0u);
  if constexpr (HAS_RES) rv = ptx::ld_shared_v4(res_row + (rsw << 4));          // 16-byte chunk index XOR (row & 7)
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 nb0 = b0, nb1 = b1, nrv = rv;
    if (c < 7) {
      nb0 = ptx::ld_shared_v4(bias + (uint32_t)(c + 1) * 32u);
      nb1 = ptx::ld_shared_v4(bias + (uint32_t)(c + 1) * 32u + 16u);
      if constexpr (HAS_RES) nrv = ptx::ld_shared_v4(res_row + ((rsw ^ (uint32_t)(c + 1)) << 4));
    }
    const uint32_t* v = c < 4 ? &v0[c * 8] : &v1[(c - 4) * 8];
    float f[8];
    f[0] = __uint_as_float(v[0]) + __uint_as_float(b0.x); f[1] = __uint_as_float(v[1]) + __uint_as_float(b0.y);
    f[2] = __uint_as_float(v[2]) + __uint_as_float(b0.z); f[3] = __uint_as_float(v[3]) + __uint_as_float(b0.w);
    f[4] = __uint_as_float(v[4]) + __uint_as_float(b1.x); f[5] = __uint_as_float(v[5]) + __uint_as_float(b1.y);
    f[6] = __uint_as_float(v[6]) + __uint_as_float(b1.z); f[7] = __uint_as_float(v[7]) + __uint_as_float(b1.w);
    if constexpr (CLAMP) {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = fminf(fmaxf(f[e], eb.lo1), eb.hi1);
    }
    if constexpr (HAS_RES) {
      float r[8];
      unpack2<T>(rv.x, r[0], r[1]); unpack2<T>(rv.y, r[2], r[3]);
      unpack2<T>(rv.z, r[4], r[5]); unpack2<T>(rv.w, r[6], r[7]);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += r[e];
      if constexpr (CLAMP) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fminf(fmaxf(f[e], eb.lo2), eb.hi2);
      }
    }
    const uint32_t dst = out_row + ((rsw ^ (uint32_t)c) << 4);
    if constexpr (RELU) ptx::st_shared_v4(dst, make_uint4(pack2_relu<T>(f[0], f[1]), pack2_relu<T>(f[2], f[3]), pack2_relu<T>(f[4], f[5]), pack2_relu<T>(f[6], f[7])));
    else ptx::st_shared_v4(dst, make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7])));
    b0 = nb0; b1 = nb1; rv = nrv;
  }
}

// The generic-activation form for one whole 64-column slab: bias, ONE activation switch over the 64 accumulators,
// residual, one more switch, pack + swizzled staging store.
template <typename T>
__device__ __forceinline__ void epilogue_slab_generic(const uint32_t (&v0)[32], const uint32_t (&v1)[32], uint32_t bias, uint32_t out_row, uint32_t res_row,
                                                      uint32_t rsw, bool has_res, const EpiAct& e1, const EpiAct& e2) {
  float f[64];
#pragma unroll
  for (int g = 0; g < 16; ++g) {
    const uint4 b = ptx::ld_shared_v4(bias + (uint32_t)g * 16u);
    const uint32_t* v = g < 8 ? &v0[g * 4] : &v1[(g - 8) * 4];
    f[g * 4 + 0] = __uint_as_float(v[0]) + __uint_as_float(b.x); f[g * 4 + 1] = __uint_as_float(v[1]) + __uint_as_float(b.y);
    f[g * 4 + 2] = __uint_as_float(v[2]) + __uint_as_float(b.z); f[g * 4 + 3] = __uint_as_float(v[3]) + __uint_as_float(b.w);
  }
  if (e1.kind != 0 && !e1.relu_in_pack) epi_act_generic<64>(f, e1);
  if (has_res) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const uint4 rv = ptx::ld_shared_v4(res_row + ((rsw ^ (uint32_t)g) << 4));
      float r[8];
      unpack2<T>(rv.x, r[0], r[1]); unpack2<T>(rv.y, r[2], r[3]);
      unpack2<T>(rv.z, r[4], r[5]); unpack2<T>(rv.w, r[6], r[7]);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[g * 8 + e] += r[e];
    }
    if (e2.kind != 0 && !e2.relu_in_pack) epi_act_generic<64>(f, e2);
  }
  const bool relu_pack = (has_res ? e2 : e1).relu_in_pack != 0;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint32_t sw = (rsw ^ (uint32_t)g) << 4;
    const float* h = &f[g * 8];
    if (relu_pack) ptx::st_shared_v4(out_row + sw, make_uint4(pack2_relu<T>(h[0], h[1]), pack2_relu<T>(h[2], h[3]), pack2_relu<T>(h[4], h[5]), pack2_relu<T>(h[6], h[7])));
    else ptx::st_shared_v4(out_row + sw, make_uint4(pack2<T>(h[0], h[1]), pack2<T>(h[2], h[3]), pack2<T>(h[4], h[5]), pack2<T>(h[6], h[7])));
  }
}

// first output row of CTA `rank` of pair tile `tile_m`, and (halo mode) its image / first image row
struct TilePos { int m0; int img, h0, w0; bool valid; };
template <int AMODE>
__device__ __forceinline__ TilePos tile_pos(const Tc3Args& a, int tile_m, int rank) {
  TilePos t;
  const int ct = tile_m * 2 + rank;                 // CTA tile index along M
  t.w0 = 0;
  if (AMODE == A_HALO) {
    t.img = ct / a.tiles_per_img;
    t.h0 = (ct - t.img * a.tiles_per_img) * a.rows_r;
    t.valid = t.img < a.p.N;
    t.m0 = 0;
  } else if (AMODE == A_S2D) {
    t.img = ct / a.tiles_per_img;
    const int rem = ct - t.img * a.tiles_per_img;
    const int rt = rem / a.nseg;
    t.h0 = rt * S2D_R;                              // first output row
    t.w0 = (rem - rt * a.nseg) * S2D_G;             // first pixel GROUP of the segment (4 output pixels per group)
    t.valid = t.img < a.p.N;
    t.m0 = 0;
  } else {
    t.m0 = ct * BM;
    t.valid = t.m0 < a.p.M;
    t.img = 0; t.h0 = 0;
  }
  return t;
}

// The two single-issuer loops (MMA, operand TMA) run with the whole warp converged and only the instruction that must be issued once
// behind elect.sync: a loop executed by one diverged lane cannot use the uniform datapath, and the first version of this kernel spent
// ~1500 cycles of scalar bookkeeping per k-block in the MMA issuer against 512 cycles of tensor work (profiles/r1_call_e).
template <typename T, int BN, int AMODE, bool WS, bool GEN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) conv_tc3_kernel(const __grid_constant__ Tc3Args args) {
  constexpr int B_BYTES = (BN / 2) * 128;     // this CTA's half of a weight k-block
  constexpr int REUSE = AMODE == A_HALO ? 9 : 1;                // k-blocks served by one A stage (halo: the nine taps of a 3x3)
  constexpr bool BOX4D = AMODE == A_HALO;                       // output / residual addressed {channel, w, h, n}
  constexpr int ACC_COLS = AMODE == A_S2D ? 4 * BN : BN;        // s2d stem: four accumulators (output pixel phases 4g + 0..3) per tile
  const ConvParams& p = args.p;
  const bool has_res = p.res != nullptr;
  const int nkb = p.Ktot / BK;
  const Plan3 plan = plan3(BN, has_res, WS ? nkb * B_BYTES : 0, AMODE, args.rows_p, p.Cout_alloc, nkb);
  const int SA = plan.sa, SB = plan.sb;
  const int cpt = p.Cin_p >> 6;                               // 64-channel chunks per filter tap
  const int n_ablk = AMODE == A_HALO ? cpt : (AMODE == A_S2D ? 1 : nkb);
  const uint32_t a_tx = AMODE == A_HALO ? (uint32_t)((BM + 2 * args.rows_p) * 128) : (AMODE == A_S2D ? (uint32_t)((S2D_R + 3) * S2D_G * 128) : (uint32_t)A_BYTES);   // bytes one CTA's A load delivers

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem_al + plan.bar_off);
  uint64_t* emptyA = fullA + MAX_STAGES;
  uint64_t* fullB = emptyA + MAX_STAGES;
  uint64_t* emptyB = fullB + MAX_STAGES;
  uint64_t* tfull = emptyB + MAX_STAGES;
  uint64_t* tempty = tfull + MAX_ACC;
  uint64_t* rfull = tempty + MAX_ACC;
  uint64_t* rempty = rfull + NRES;
  uint64_t* wfull = rempty + NRES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = args.num_tiles, n_tiles = args.n_tiles;
  const int cluster = (int)ptx::cluster_id_x(), nclusters = (int)ptx::num_clusters_x();
  const uint32_t acc_sh = (uint32_t)args.acc_log2, acc_mask = (1u << acc_sh) - 1u;   // accumulator of tile `it`: it & acc_mask, phase (it >> acc_sh) & 1

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      // full*: ONE arrival, the leader's arrive.expect_tx for the bytes of BOTH CTAs.  The peer only issues its TMA (complete_tx on the
      // leader's barrier); a remote mbarrier.arrive per stage costs the peer's producer a ~1200-cycle cluster round trip (profiles/r1_call_b).
      ptx::mbar_init(&fullA[s], 1);
      ptx::mbar_init(&emptyA[s], 1);
      ptx::mbar_init(&fullB[s], 1);
      ptx::mbar_init(&emptyB[s], 1);
    }
    for (int a = 0; a < MAX_ACC; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      // accumulator drained: the eight local epilogue warps; in the leader also the peer's forwarder (the peer's otherwise idle MMA warp
      // waits for ITS eight warps on its local copy and does the one remote arrive, so no epilogue warp pays the cluster round trip)
      ptx::mbar_init(&tempty[a], leader ? NUM_EPI_WARPS + 1 : NUM_EPI_WARPS);
    }
    for (int r = 0; r < NRES; ++r) {
      ptx::mbar_init(&rfull[r], 1);
      ptx::mbar_init(&rempty[r], NUM_EPI_WARPS / 2);   // the four warps of the group that consumed the slab
    }
    ptx::mbar_init(wfull, 1);
    ptx::fence_mbar_init();
  }
  if (warp == TMA_WARP && lane == 0) {
    ptx::tma_prefetch_desc(&args.tm_w);
    ptx::tma_prefetch_desc(&args.tm_a);
    ptx::tma_prefetch_desc(&args.tm_y);
    if (has_res) ptx::tma_prefetch_desc(&args.tm_r);
  }
  if (warp == MMA_WARP) {
    ptx::tmem_alloc_pair(tmem_slot, (uint32_t)ACC_COLS << args.acc_log2);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();                   // barriers of both CTAs are initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) ran while the previous kernel of the stream was
  // still draining; its outputs are read only below this line.  Our own dependents may start their prologue as our CTAs retire.
  // Each role waits for the dependency itself, right before its first access to memory the previous kernel may have written: the constant
  // operands (bias vector, resident weights) are fetched BEFORE the wait, while the producer's last tiles drain.
  ptx::grid_launch_dependents();

  if (warp < NUM_EPI_WARPS) {
    // ===================== epilogue: this CTA's 128 rows, one 64-channel slab at a time =====================
    // Two groups of four warps (one warp per TMEM lane quarter) take alternate slabs, each with its own staging slab and its own
    // 128-thread named barrier: two slabs are in flight at any time, which is what hides the TMEM -> registers -> staging -> TMA store
    // latency chain of ~2300 cycles per slab that bounded every layer with few k-blocks per tile (profiles/r1_call_h).
    const int grp = warp >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t rsw = (uint32_t)(row & 7);
    const bool store_thread = (threadIdx.x & 127) == 0;
    const bool wstore = args.warp_store != 0;
    const uint32_t per_grp = (uint32_t)plan.nout >> 1;       // staging slabs of this group: 1 or 2
    uint32_t mine = 0;                                       // slabs this group has processed
    // bias vector -> shared memory (all 256 epilogue threads)
    {
      float* s_bias = reinterpret_cast<float*>(smem_al + plan.bias_off);
      for (int i = threadIdx.x; i < p.Cout_alloc; i += NUM_EPI_WARPS * 32) s_bias[i] = __ldg(p.bias + i);
      ptx::named_bar_sync(3, NUM_EPI_WARPS * 32);
    }
    if (!(args.debug & 4)) ptx::grid_dependency_wait();
    const uint32_t bias_base = smem_base + plan.bias_off;
    EpiBounds eb;
    const int variant = epi_variant(has_res, args.e1, args.e2, eb);
    if (AMODE == A_S2D && args.pool) {
      // ===== stem + 3x3 / stride-2 / pad-1 max pool (ConvParams::pool).  Warp (q, grp) owns output row h0 + q of the tile and channels
      // [32 grp, 32 grp + 32); lane = pixel group g (pixels 4g .. 4g + 3 = the four phase accumulators).  Per tile: the four phases -> bias,
      // ReLU, 16-bit pack (64 registers); pooled columns 2g = max(pixel 4g - 1 from lane g - 1, 4g, 4g + 1) and 2g + 1 = max(4g + 1, 4g + 2,
      // 4g + 3) in packed arithmetic; the two warps of a row fill one swizzled [64 pooled pixels][64 channels] slab, and one TMA max-REDUCE
      // per pooled row it belongs to (row r even: r / 2; r odd: (r - 1) / 2 and (r + 1) / 2) folds it into the zero-filled pooled tensor.
      // The 411 MB stem output and its re-read by a pooling kernel never exist (ResNet-50 b256: stem 137 us + pool 88 us before).
      const int half = grp;
      const bool issuer = half == 0 && lane == 0;
      const int Hp = p.Ho >> 1;
      uint32_t it = 0;
      for (int t_ = cluster; t_ < num_tiles; t_ += nclusters, ++it) {
        const int tile = args.reverse ? num_tiles - 1 - t_ : t_;
        const TilePos tp = tile_pos<AMODE>(args, tile / n_tiles, (int)rank);
        const uint32_t acc = it & acc_mask;
        ptx::mbar_wait(&tfull[acc], (it >> acc_sh) & 1u);
        ptx::tc_fence_after();
        uint32_t pk[4][16];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + (uint32_t)(f * BN + half * 32), v);
          uint4 b[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) b[j] = ptx::ld_shared_v4(bias_base + (uint32_t)(half * 32 + j * 4) * 4u);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pk[f][2 * j] = pack2_relu<T>(__uint_as_float(v[4 * j]) + __uint_as_float(b[j].x), __uint_as_float(v[4 * j + 1]) + __uint_as_float(b[j].y));
            pk[f][2 * j + 1] = pack2_relu<T>(__uint_as_float(v[4 * j + 2]) + __uint_as_float(b[j].z), __uint_as_float(v[4 * j + 3]) + __uint_as_float(b[j].w));
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
        // staging slab of (tile parity, row q): 8 KB.  The reduce that read it two tiles ago has finished reading
        const uint32_t slab = smem_base + plan.out_off + (it & 1u) * 32768u + (uint32_t)q * 8192u;
        if (issuer) ptx::bulk_wait_read<1>();
        ptx::named_bar_sync(4 + q, 64);
        const uint32_t px0 = 2u * (uint32_t)lane, px1 = px0 + 1u;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint32_t o0[4], o1[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = jj * 4 + e;
            uint32_t left = __shfl_up_sync(0xffffffffu, pk[3][j], 1);
            if (lane == 0) left = 0u;                                  // pixel -1 is padding; every value is >= 0
            o0[e] = max2<T>(max2<T>(left, pk[0][j]), pk[1][j]);
            o1[e] = max2<T>(max2<T>(pk[1][j], pk[2][j]), pk[3][j]);
          }
          const uint32_t ch = (uint32_t)(half * 4 + jj);              // 16-byte chunk of the 128-byte pixel row
          ptx::st_shared_v4(slab + px0 * 128u + ((ch ^ (px0 & 7u)) << 4), make_uint4(o0[0], o0[1], o0[2], o0[3]));
          ptx::st_shared_v4(slab + px1 * 128u + ((ch ^ (px1 & 7u)) << 4), make_uint4(o1[0], o1[1], o1[2], o1[3]));
        }
        ptx::fence_proxy_async();
        ptx::named_bar_sync(4 + q, 64);
        if (issuer) {
          const int r = tp.h0 + q;
          if (tp.valid && r < p.Ho && !(args.debug & 1)) {
            ptx::tma_reduce_max_4d(&args.tm_y, slab, 0, 0, r >> 1, tp.img);
            if ((r & 1) && ((r + 1) >> 1) < Hp) ptx::tma_reduce_max_4d(&args.tm_y, slab, 0, 0, (r + 1) >> 1, tp.img);
          }
          ptx::bulk_commit();
        }
      }
      if (issuer) ptx::bulk_wait_read<0>();
    } else {
    uint32_t it = 0, g = 0;                // tiles / slabs processed so far (all slabs, both groups)
    for (int t_ = cluster; t_ < num_tiles; t_ += nclusters, ++it) {
      const int tile = args.reverse ? num_tiles - 1 - t_ : t_;
      const TilePos tp = tile_pos<AMODE>(args, tile / n_tiles, (int)rank);
      const int n0 = (tile % n_tiles) * BN;
      const int ncols = min(BN, p.Cout_p - n0);
      const int nslab = AMODE == A_S2D ? 4 : (ncols + 63) >> 6;   // s2d stem: slab h = output pixel phase h (BN == 64 channels each)
      const uint32_t acc = it & acc_mask;
      ptx::mbar_wait(&tfull[acc], (it >> acc_sh) & 1u);
      ptx::tc_fence_after();
      bool released = false;
      for (int h = 0; h < nslab; ++h, ++g) {
        if ((int)(g & 1u) != grp) continue;
        const uint32_t rslot = g % NRES;
        const uint32_t res_row = smem_base + plan.res_off + rslot * SLAB_BYTES + (uint32_t)row * 128u;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + (uint32_t)(h * 64);
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32(taddr, v0);
        ptx::tmem_ld_32x32(taddr + 32, v1);
        const uint32_t out_slab = smem_base + plan.out_off + ((uint32_t)grp * per_grp + (mine & (per_grp - 1u))) * SLAB_BYTES;
        const uint32_t out_row = out_slab + (uint32_t)row * 128u;
        ++mine;
        if (has_res) ptx::mbar_wait(&rfull[rslot], (g / NRES) & 1u);
        if (wstore) {
          // per-warp stores: this warp's own previous store from this staging slab has finished reading its 32 rows
          if (lane == 0) { if (per_grp == 2) ptx::bulk_wait_read<1>(); else ptx::bulk_wait_read<0>(); }
          __syncwarp();
        } else {
          if (store_thread) {                                     // the store that last used this staging slab has finished reading it
            if (per_grp == 2) ptx::bulk_wait_read<1>(); else ptx::bulk_wait_read<0>();
          }
          ptx::named_bar_sync(1 + grp, 128);
        }
        ptx::tmem_ld_wait();
        if (args.early_release && h + 2 >= nslab) {
          // the last slab of the tile this group reads is in registers: hand the accumulator back now, a whole slab latency (conversion,
          // staging, store issue) earlier than at the end of the tile
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
          released = true;
        }
        const uint32_t bias = bias_base + (uint32_t)(n0 + (AMODE == A_S2D ? 0 : h * 64)) * 4u;
        if (args.debug & 2) {
        } else if constexpr (GEN) {
          epilogue_slab_generic<T>(v0, v1, bias, out_row, res_row, rsw, has_res, args.e1, args.e2);
        } else {
          if (has_res) {
            if (variant == 0) epilogue_slab<T, true, false, false>(v0, v1, bias, out_row, res_row, rsw, eb);
            else if (variant == 1) epilogue_slab<T, true, false, true>(v0, v1, bias, out_row, res_row, rsw, eb);
            else epilogue_slab<T, true, true, false>(v0, v1, bias, out_row, res_row, rsw, eb);
          } else {
            if (variant == 0) epilogue_slab<T, false, false, false>(v0, v1, bias, out_row, res_row, rsw, eb);
            else if (variant == 1) epilogue_slab<T, false, false, true>(v0, v1, bias, out_row, res_row, rsw, eb);
            else epilogue_slab<T, false, true, false>(v0, v1, bias, out_row, res_row, rsw, eb);
          }
        }
        if (has_res) {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&rempty[rslot]);
        }
        ptx::fence_proxy_async();                                 // staging writes -> visible to the TMA engine
        if (wstore) {
          __syncwarp();
          if (lane == 0) {
            if (tp.valid && !(args.debug & 1)) {
              const uint32_t src = out_slab + (uint32_t)q * 4096u;      // rows 32 q .. 32 q + 31 of the slab
              if (AMODE == A_S2D) ptx::tma_store_5d(&args.tm_yw, src, 0, h, tp.w0, tp.h0 + q, tp.img);
              else if (BOX4D) ptx::tma_store_4d(&args.tm_yw, src, n0 + h * 64, (q * 32) % args.rows_p, tp.h0 + (q * 32) / args.rows_p, tp.img);
              else if (args.p.y_heads_s) {                              // head-split output: {64, S, heads, N}
                const int row = tp.m0 + q * 32, nb = row / args.p.y_heads_s;
                ptx::tma_store_4d(&args.tm_yw, src, 0, row - nb * args.p.y_heads_s, (n0 + h * 64) >> 6, nb);
              } else ptx::tma_store_2d(&args.tm_yw, src, n0 + h * 64, tp.m0 + q * 32);
            }
            ptx::bulk_commit();
          }
          continue;
        }
        ptx::named_bar_sync(1 + grp, 128);
        if (store_thread) {
          if (tp.valid && !(args.debug & 1)) {
            if (AMODE == A_S2D) ptx::tma_store_5d(&args.tm_y, out_slab, 0, h, tp.w0, tp.h0, tp.img);   // {channel, phase, pixel group, row, image}
            else if (BOX4D) ptx::tma_store_4d(&args.tm_y, out_slab, n0 + h * 64, tp.w0, tp.h0, tp.img);   // rows >= Ho / pixels >= Wo are clipped
            else ptx::tma_store_2d(&args.tm_y, out_slab, n0 + h * 64, tp.m0);
          }
          ptx::bulk_commit();
        }
      }
      // this warp's TMEM reads of the tile are complete (tmem_ld_wait above): hand the accumulator back
      if (!released) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      }
    }
    // the staging slabs must outlive the stores' READS of them; the writes complete with the grid (every consumer waits on grid completion)
    if (wstore ? lane == 0 : store_thread) ptx::bulk_wait_read<0>();
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer (leader CTA; whole warp walks the loop, one elected lane issues) =====================
    if (!leader) {
      // peer CTA: forward "accumulator drained by my eight epilogue warps" to the leader's barrier, one remote arrive per tile
      const uint32_t tempty_leader0 = ptx::mapa(ptx::smem_u32(&tempty[0]), 0);
      uint32_t it = 0;
      for (int t_ = cluster; t_ < num_tiles; t_ += nclusters, ++it) {
        const uint32_t acc = it & acc_mask;
        ptx::mbar_wait(&tempty[acc], (it >> acc_sh) & 1u);
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader0 + acc * 8u);
        __syncwarp();
      }
    }
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(TcType<T>::is_bf16, 2 * BM, BN);
      // descriptor = DESC_HI | (address >> 4): version 1, SWIZZLE_128B, SBO 1024 B, LBO 16 B (ptx_sm100.cuh make_smem_desc_sw128)
      constexpr uint64_t DESC_HI = ((uint64_t)2 << 61) | ((uint64_t)1 << 46) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 16);
      const uint32_t a_lo0 = ((smem_base + plan.a_off) & 0x3FFFFu) >> 4, a_step = plan.a_stage >> 4;
      const uint32_t b_lo0 = ((smem_base + (WS ? plan.w_off : plan.b_off)) & 0x3FFFFu) >> 4;
      constexpr uint32_t b_step = B_BYTES >> 4;
      const uint32_t P8 = (uint32_t)args.rows_p * 8u;             // one image row of the halo box, in 16-byte units
      uint32_t it = 0;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      if (WS) ptx::mbar_wait(wfull, 0);
      for (int t_ = cluster; t_ < num_tiles; t_ += nclusters, ++it) {
        const uint32_t acc = it & acc_mask;
        ptx::mbar_wait(&tempty[acc], ((it >> acc_sh) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        uint32_t accum = 0;
        for (int ab = 0; ab < n_ablk; ++ab) {
          ptx::mbar_wait(&fullA[sa], pha);
          const uint32_t a_lo = a_lo0 + (uint32_t)sa * a_step;
          if (AMODE == A_S2D) {
            // space-to-depth stem.  A row of the box = one GROUP of 4 block pixels (4 x 16 channels = 128 swizzled bytes), 32 groups per
            // block row, R + 3 block rows.  MMA row m = (output row m / 32, group m % 32); phase f = output pixel 4 g + f; tap (r, s) of the
            // 4x4 window reads block pixel 4 g + f + s = group g + (f + s) / 4, K step (f + s) % 4, so every (phase, tap) is ONE MMA over a
            // shifted view of the same box: 64 MMAs per tile, four accumulators.  Row 32 * y + 31 + 1 wraps into the next block row: those
            // are pixel groups >= 29 of a 112-pixel row, dropped by the store map.
            ptx::tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int f = 0; f < 4; ++f) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int off = f + (j % 4);
                  const uint32_t a_view = a_lo + (uint32_t)(((j / 4) * S2D_G + off / 4) * 8 + (off % 4) * 2);
                  const uint32_t b_lo = b_lo0 + (uint32_t)(j / 4) * b_step + 2u * (uint32_t)(j % 4);
                  ptx::umma_f16_pair(d_tmem + (uint32_t)(f * BN), DESC_HI | (uint64_t)a_view, DESC_HI | (uint64_t)b_lo, idesc, (uint32_t)(j > 0));
                }
              }
            }
            __syncwarp();
            accum = 1;
          } else {
#pragma unroll
          for (int j = 0; j < REUSE; ++j) {
            // halo mode: tap (j / 3, j % 3) reads the box shifted by whole image rows and pixels (one pixel = one 128-byte swizzled row);
            // its weight k-block is (tap, chunk ab): K order of the packed weights is (filter row, filter col, channel)
            const uint32_t a_view = AMODE == A_HALO ? a_lo + (uint32_t)(j / 3) * P8 + (uint32_t)(j % 3) * 8u : a_lo;
            uint32_t b_lo;
            if (WS) {
              b_lo = b_lo0 + (uint32_t)(AMODE == A_HALO ? j * cpt + ab : ab) * b_step;
            } else {
              ptx::mbar_wait(&fullB[sb], phb);
              b_lo = b_lo0 + (uint32_t)sb * b_step;
            }
            ptx::tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                ptx::umma_f16_pair(d_tmem, DESC_HI | (uint64_t)(a_view + 2u * k), DESC_HI | (uint64_t)(b_lo + 2u * k), idesc, accum | (uint32_t)k);
              if (!WS) ptx::umma_commit_pair(&emptyB[sb], 3);
            }
            __syncwarp();
            accum = 1;
            if (!WS) { if (++sb == SB) { sb = 0; phb ^= 1u; } }
          }
          }
          if (elect_one()) ptx::umma_commit_pair(&emptyA[sa], 3);
          __syncwarp();
          if (++sa == SA) { sa = 0; pha ^= 1u; }
        }
        if (elect_one()) ptx::umma_commit_pair(&tfull[acc], 3);
        __syncwarp();
      }
    }
  } else if (warp == TMA_WARP) {
    // ===================== operand TMA issuer (one warp per CTA, one elected lane issues) =====================
    const uint32_t fullA_leader = ptx::smem_u32(&fullA[0]) & ptx::PEER_BIT_MASK;
    const uint32_t fullB_leader = ptx::smem_u32(&fullB[0]) & ptx::PEER_BIT_MASK;
    const uint32_t wfull_leader = ptx::smem_u32(wfull) & ptx::PEER_BIT_MASK;
    const int HoWo = p.Ho * p.Wo;
    if (p.w_dynamic && !(args.debug & 4)) ptx::grid_dependency_wait();
    if (WS) {   // this CTA's half of the weights of this cluster's (only) N tile, once: the host makes the cluster count a multiple of n_tiles
      if (lane == 0) {
        const int first = args.reverse ? num_tiles - 1 - cluster : cluster;
        const int n_fixed = first % n_tiles;
        if (leader) ptx::mbar_arrive_expect_tx(wfull, 2u * (uint32_t)nkb * B_BYTES);
        for (int kb = 0; kb < nkb; ++kb)
          ptx::tma_load_2d_pair(smem_base + plan.w_off + (uint32_t)kb * B_BYTES, &args.tm_w, wfull_leader, kb * BK, n_fixed * BN + (int)rank * (BN / 2));
      }
      __syncwarp();
    }
    if (!(args.debug & 4)) ptx::grid_dependency_wait();
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    for (int t_ = cluster; t_ < num_tiles; t_ += nclusters) {
      const int tile = args.reverse ? num_tiles - 1 - t_ : t_;
      const int tile_m = tile / n_tiles;
      TilePos tp = tile_pos<AMODE>(args, tile_m, (int)rank);
      if (!tp.valid) tp = tile_pos<AMODE>(args, tile_m, 0);            // tail: stage the leader's rows again (never stored)
      const int n0 = (tile - tile_m * n_tiles) * BN + (int)rank * (BN / 2);
      int bn = 0, bh = 0, bw = 0;
      if (AMODE == A_IM2COL) {
        bn = tp.m0 / HoWo;
        const int rem = tp.m0 - bn * HoWo;
        const int ho = rem / p.Wo;
        bh = ho * p.sh - p.pt;
        bw = (rem - ho * p.Wo) * p.sw - p.pl;
      }
      if (AMODE == A_MATRIX && args.prefetch_tiles > 0) {
        // activation matrix rows of a later tile -> L2 (once per M tile: by the cluster that walks its first N tile)
        const int tla = t_ + args.prefetch_tiles * nclusters;
        if (tla < num_tiles) {
          const int tl = args.reverse ? num_tiles - 1 - tla : tla;
          const int ml = tl / n_tiles;
          const TilePos tq = tile_pos<AMODE>(args, ml, (int)rank);
          if (tl - ml * n_tiles == 0 && tq.valid && lane < n_ablk) {
            for (int ab = lane; ab < n_ablk; ab += 32) ptx::tma_prefetch_2d(&args.tm_a, ab * BK, tq.m0);
          }
          __syncwarp();
        }
      }
      int bn2 = 0, bh2 = 0, bw2 = 0;
      if (AMODE == A_MATRIX && args.nkb1 < n_ablk) {      // base pixel of this tile in the second source
        bn2 = tp.m0 / HoWo;
        const int rem = tp.m0 - bn2 * HoWo;
        const int ho = rem / p.Wo;
        bh2 = ho * p.sh2;
        bw2 = (rem - ho * p.Wo) * p.sw2;
      }
      int chunk = 0, fr = 0, fs = 0;
      for (int ab = 0; ab < n_ablk; ++ab) {
        ptx::mbar_wait(&emptyA[sa], pha ^ 1u);
        if (elect_one()) {
          const uint32_t a_dst = smem_base + plan.a_off + (uint32_t)sa * plan.a_stage;
          const uint32_t bar = fullA_leader + (uint32_t)sa * 8u;
          if (leader) ptx::mbar_arrive_expect_tx(&fullA[sa], 2u * a_tx);
          if (AMODE == A_MATRIX) {
            if (ab < args.nkb1) ptx::tma_load_2d_pair_hint(a_dst, &args.tm_a, bar, ab * BK, tp.m0, args.a_policy);
            else ptx::tma_load_im2col_4d_pair(a_dst, &args.tm_a2, bar, (ab - args.nkb1) * BK, bw2, bh2, bn2, 0u, 0u);   // strided pixels of x2
          } else if (AMODE == A_IM2COL) {
            ptx::tma_load_im2col_4d_pair(a_dst, &args.tm_a, bar, chunk * BK, bw, bh, bn, (uint32_t)(fs * p.dw), (uint32_t)(fr * p.dh));
          } else if (AMODE == A_S2D) {
            // block rows h0 .. h0 + R + 2, pixel groups w0 .. w0 + 31; beyond the block grid -> zero fill
            ptx::tma_load_4d_pair(a_dst, &args.tm_a, bar, 0, tp.w0, tp.h0, tp.img);
          } else {
            // all filter rows at once: input rows h0 - pt .. h0 - pt + R + 1, pixels -pl .. P - pl - 1 of 64-channel chunk ab
            ptx::tma_load_4d_pair(a_dst, &args.tm_a, bar, ab * BK, -p.pl, tp.h0 - p.pt, tp.img);
          }
        }
        __syncwarp();
        if (AMODE == A_IM2COL) { if (++chunk == cpt) { chunk = 0; if (++fs == p.kw) { fs = 0; ++fr; } } }
        if (!WS) {
#pragma unroll 1
          for (int j = 0; j < REUSE; ++j) {
            const int kb = AMODE == A_HALO ? j * cpt + ab : (AMODE == A_S2D ? j : ab);
            ptx::mbar_wait(&emptyB[sb], phb ^ 1u);
            if (elect_one()) {
              const uint32_t barb = fullB_leader + (uint32_t)sb * 8u;
              if (leader) ptx::mbar_arrive_expect_tx(&fullB[sb], 2u * B_BYTES);
              ptx::tma_load_2d_pair(smem_base + plan.b_off + (uint32_t)sb * plan.b_stage, &args.tm_w, barb, kb * BK, n0);
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; phb ^= 1u; }
          }
        }
        if (++sa == SA) { sa = 0; pha ^= 1u; }
      }
    }
  } else if (warp == RES_WARP) {
    // ===================== residual TMA issuer: one slab ahead of the epilogue, ring of NRES =====================
    if (has_res && lane == 0) {
      if (!(args.debug & 4)) ptx::grid_dependency_wait();
      uint32_t g = 0;
      const int la = args.prefetch_tiles * nclusters;
      for (int t_ = cluster; t_ < num_tiles; t_ += nclusters) {
        const int tile = args.reverse ? num_tiles - 1 - t_ : t_;
        const TilePos tp = tile_pos<AMODE>(args, tile / n_tiles, (int)rank);
        const int n0 = (tile % n_tiles) * BN;
        const int ncols = min(BN, p.Cout_p - n0);
        const int nslab = (ncols + 63) >> 6;
        if (la > 0 && t_ + la < num_tiles) {
          // the ring holds NRES slabs (64 KB) per SM, one HBM latency's worth at most: pull the slabs of a later tile into L2 now so
          // that their ring loads are L2 hits
          const int tl = args.reverse ? num_tiles - 1 - (t_ + la) : t_ + la;
          const TilePos tq = tile_pos<AMODE>(args, tl / n_tiles, (int)rank);
          const int q0 = (tl % n_tiles) * BN;
          const int qslab = (min(BN, p.Cout_p - q0) + 63) >> 6;
          if (tq.valid)
            for (int h = 0; h < qslab; ++h) {
              if (BOX4D) ptx::tma_prefetch_4d(&args.tm_r, q0 + h * 64, tq.w0, tq.h0, tq.img);
              else ptx::tma_prefetch_2d(&args.tm_r, q0 + h * 64, tq.m0);
            }
        }
        for (int h = 0; h < nslab; ++h, ++g) {
          const uint32_t rslot = g % NRES;
          ptx::mbar_wait(&rempty[rslot], ((g / NRES) & 1u) ^ 1u);
          const uint32_t dst = smem_base + plan.res_off + rslot * SLAB_BYTES;
          if (tp.valid) {
            ptx::mbar_arrive_expect_tx(&rfull[rslot], SLAB_BYTES);
            if (BOX4D) ptx::tma_load_4d_hint(dst, &args.tm_r, &rfull[rslot], n0 + h * 64, tp.w0, tp.h0, tp.img, args.r_policy);
            else ptx::tma_load_2d_hint(dst, &args.tm_r, &rfull[rslot], n0 + h * 64, tp.m0, args.r_policy);
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
  if (warp == MMA_WARP) ptx::tmem_dealloc_pair(tmem_base, (uint32_t)ACC_COLS << args.acc_log2);
}

// cluster dims are static (__cluster_dims__); the one launch attribute is programmatic stream serialization (B200_PDL=0 turns it off)
template <typename K>
inline void launch_pdl(K kernel, int grid, unsigned smem, cudaStream_t s, const Tc3Args& a) {
  static const bool pdl = [] { const char* e = getenv("B200_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, a);
}

// ---- launch (one instantiation set per element type; defined in conv_tc3.cu / conv_tc3_f16.cu) ---------------------------------
template <typename T>
void launch_tc3_typed(const Tc3Args& a, int bn, int a_mode, bool ws, int grid, unsigned smem, cudaStream_t s);
template <typename T>
void opt_in_tc3_typed();

#define B200_TC3_FOR_EACH(T, X) \
  X(T, 64, A_MATRIX, false) X(T, 64, A_MATRIX, true) X(T, 64, A_IM2COL, false) X(T, 64, A_IM2COL, true) X(T, 64, A_HALO, false) X(T, 64, A_HALO, true) \
  X(T, 128, A_MATRIX, false) X(T, 128, A_MATRIX, true) X(T, 128, A_IM2COL, false) X(T, 128, A_IM2COL, true) X(T, 128, A_HALO, false) X(T, 128, A_HALO, true) \
  X(T, 256, A_MATRIX, false) X(T, 256, A_MATRIX, true) X(T, 256, A_IM2COL, false) X(T, 256, A_IM2COL, true) X(T, 256, A_HALO, false) X(T, 256, A_HALO, true) \
  X(T, 64, A_S2D, true)

#define B200_TC3_DEFINE_TYPED(T)                                                                                                          \
  template <> void launch_tc3_typed<T>(const Tc3Args& a, int bn, int a_mode, bool ws, int grid, unsigned smem, cudaStream_t s) {           \
    bool done = false;                                                                                                                     \
    B200_TC3_FOR_EACH(T, B200_TC3_LAUNCH_CASE)                                                                                             \
    B200_CHECK(done, B200_ERR_UNSUPPORTED, "conv_tc3: no kernel for bn " << bn << " mode " << a_mode);                                     \
  }                                                                                                                                        \
  template <> void opt_in_tc3_typed<T>() { B200_TC3_FOR_EACH(T, B200_TC3_OPTIN_CASE) }
#define B200_TC3_LAUNCH_CASE(T, BN_, MODE_, WS_)                                                          \
  if (!done && bn == BN_ && a_mode == MODE_ && ws == WS_) {                                               \
    if (a.gen) launch_pdl(conv_tc3_kernel<T, BN_, MODE_, WS_, true>, grid, smem, s, a);                   \
    else launch_pdl(conv_tc3_kernel<T, BN_, MODE_, WS_, false>, grid, smem, s, a);                        \
    done = true;                                                                                          \
  }
#define B200_TC3_OPTIN_CASE(T, BN_, MODE_, WS_)                                                                                 \
  cudaFuncSetAttribute(conv_tc3_kernel<T, BN_, MODE_, WS_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);     \
  cudaFuncSetAttribute(conv_tc3_kernel<T, BN_, MODE_, WS_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);

}  // namespace tc3
}  // namespace b200
