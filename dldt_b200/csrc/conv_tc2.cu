// Persistent, warp-specialised implicit-GEMM convolution / GEMM on tcgen05 (sm_100a), version 2.
//
//   D[M, Cout] = im2col(X)[M, K] * Wp[Cout, K]^T  (+ bias, activation, + residual, activation)
//
// Same arithmetic contract as conv_tc.cu (ngraph reference/convolution.hpp:89-186; post-op chain of
// nodes/mkldnn_conv_node.cpp:355-403); what changes is how the SM is kept busy:
//
//   * one CTA per SM, looping over output tiles (static round-robin; n-tiles of one m-tile run on
//     neighbouring SMs at the same time so the activation tile is shared through L2)
//   * TMEM holds TWO accumulators (2 x BN fp32 columns): the epilogue of tile i overlaps the main loop of tile i+1
//   * operand feed by TMA wherever the operand is a plain matrix: weights always ([Cout_alloc][Ktot], 128B-swizzled
//     boxes of 64 x BN); activations too when the convolution is a GEMM (1x1, stride 1, no padding, and FullyConnected).
//     Other geometries gather the A tile with 16-byte cp.async (zero-fill for padding) exactly as conv_tc.cu does.
//   * epilogue: TMEM -> registers -> bias/act/residual -> 128B-swizzled staging tile -> ONE TMA tensor store per 64
//     output channels (clipped at the M / Cout tails by the tensor map); the residual tile arrives through a TMA load
//     issued at the START of the tile (double buffered), so its HBM latency hides behind the main loop
//
//   warp 0-3  A producers (cp.async gather; idle in TMA-A mode)        warp 8  MMA issuer (one elected lane)
//   warp 4-7  epilogue (TMEM lane quarter = warp % 4)                  warp 9  TMA issuer (weights, residual, A)
//
// mbarriers: full[s]/empty[s] (operand stages), tfull[a]/tempty[a] (accumulators), rfull[b]/rempty[b] (residual tiles).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <mutex>

#include "act.cuh"
#include "common.h"
#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tc_common.cuh"

namespace b200 {

namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * 128;        // 128 rows x 64 elements x 2 B
constexpr int SLAB_BYTES = BM * 128;     // one 64-channel slab of the output / residual tile
constexpr int NUM_A_PRODUCERS = 128;     // warps 0-3
constexpr int NUM_EPI_WARPS = 8;         // warps 4-11: TMEM lane quarter = warp % 4, column half = (warp - 4) / 4
constexpr int MMA_WARP = 12, TMA_WARP = 13;
constexpr int NUM_THREADS = 448;
constexpr int LAG = 2;                   // cp.async groups in flight before the oldest is published
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 232448;       // 227 KB opt-in maximum per CTA on sm_100

// Shared-memory plan (all offsets multiples of 1024 from the aligned base):
//   [stages x (A tile 16 KB + B tile BN*128 B)] [output staging x2] [residual staging x2, only when a residual is fused] [barriers]
// The stage count takes whatever is left: 5 (BN=128) / 8 (BN=64) without a residual, 3 / 6 with one.
struct SmemPlan { int stages; uint32_t stage_bytes, out_off, res_off, w_off, bar_off, total; };
// ws_bytes != 0: weight-stationary mode - the whole packed weight matrix (single N tile, <= 96 KB) is loaded once per CTA and the
// ring stages carry activation tiles only (removes the per-tile re-fetch of the weights from L2, half of the operand traffic of the
// 3x3 64->64 layers and of the stem).
template <int BN>
__host__ __device__ inline SmemPlan smem_plan(bool has_res, int ws_bytes) {
  const int stage_bytes = ws_bytes ? A_BYTES : A_BYTES + BN * 128, tile_bytes = (BN / 64) * SLAB_BYTES;
  const int fixed = 2 * tile_bytes + (has_res ? 2 * tile_bytes : 0) + ws_bytes + 1024 /*align slack*/ + 1024 /*barriers*/;
  int stages = (SMEM_LIMIT - fixed) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  SmemPlan pl;
  pl.stages = stages;
  pl.stage_bytes = (uint32_t)stage_bytes;
  pl.out_off = (uint32_t)(stages * stage_bytes);
  pl.res_off = pl.out_off + 2u * tile_bytes;
  pl.w_off = pl.res_off + (has_res ? 2u * tile_bytes : 0u);
  pl.bar_off = pl.w_off + (uint32_t)ws_bytes;
  pl.total = pl.bar_off + 1024u + 1024u;
  return pl;
}
constexpr int WS_MAX_BYTES = 96 * 1024;

struct Tc2Args {
  CUtensorMap tm_a;   // activations: [M][Cin_p] matrix (a_tma 1) or {C, W, H, N} im2col tensor (a_tma 2)
  CUtensorMap tm_w;   // packed weights [Cout_alloc][Ktot]
  CUtensorMap tm_y;   // output [M][Cout_p], row pitch ldy
  CUtensorMap tm_r;   // residual, same geometry as the output
  ConvParams p;
  EpiAct e1, e2;      // activation after bias / after the residual sum
  int a_tma;          // 0: cp.async gather, 1: 2-D TMA (A is a matrix), 2: im2col-mode TMA
  int num_tiles, n_tiles;
  int ws;             // weight-stationary mode (see smem_plan)
  int debug;          // timing experiments only (B200_DEBUG_TC3 bit 2: no programmatic-launch wait; results are wrong)
  int reverse;        // walk the tiles from the last row block to the first (see ConvParams::reverse)
};

// Epilogue of one warp over NCH chunks of 32 accumulator columns of its 32 rows: TMEM -> registers -> + bias -> act -> (+ residual ->
// act) -> bf16/fp16 -> 128B-swizzled staging rows.  FULL: every column is a valid output channel (no per-group bounds checks).
template <typename T, int NCH, bool GEN, bool FULL>
__device__ __forceinline__ void epilogue_warp(uint32_t taddr, const float* __restrict__ bias, uint32_t out_row, uint32_t res_row, uint32_t xsw,
                                              int col0, int ncols, bool has_res, const EpiAct& e1, const EpiAct& e2) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (!FULL && col0 + c * 32 >= ncols) break;          // warp-uniform
    // one 32-column chunk at a time keeps the live accumulator registers at 32 (two chunks in flight spilled at 128 regs/thread);
    // the other epilogue warp on this scheduler covers the TMEM latency
    uint32_t v[1][32];
    ptx::tmem_ld_32x32(taddr + (uint32_t)(c * 32), v[0]);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = col0 + c * 32 + g * 8;            // column within the tile
      if (FULL || col < ncols) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
        float f[8];
        f[0] = __uint_as_float(v[0][g * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[0][g * 8 + 1]) + b0.y;
        f[2] = __uint_as_float(v[0][g * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[0][g * 8 + 3]) + b0.w;
        f[4] = __uint_as_float(v[0][g * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[0][g * 8 + 5]) + b1.y;
        f[6] = __uint_as_float(v[0][g * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[0][g * 8 + 7]) + b1.w;
        epi_act8<GEN>(f, e1);
        // 16-byte chunk (c*4+g) of this warp's column range, XOR-swizzled by the row (xsw already holds (first chunk & 7) ^ (row & 7))
        const uint32_t sw = (xsw ^ (uint32_t)(c * 4 + g)) << 4;
        if (has_res) {
          const uint4 rv = ptx::ld_shared_v4(res_row + sw);
          float r[8];
          unpack2<T>(rv.x, r[0], r[1]); unpack2<T>(rv.y, r[2], r[3]);
          unpack2<T>(rv.z, r[4], r[5]); unpack2<T>(rv.w, r[6], r[7]);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += r[e];
          epi_act8<GEN>(f, e2);
        }
        if ((has_res ? e2 : e1).relu_in_pack)
          ptx::st_shared_v4(out_row + sw, make_uint4(pack2_relu<T>(f[0], f[1]), pack2_relu<T>(f[2], f[3]), pack2_relu<T>(f[4], f[5]), pack2_relu<T>(f[6], f[7])));
        else
          ptx::st_shared_v4(out_row + sw, make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7])));
      }
    }
  }
}

// The clamp-family epilogue (none / ReLU / Clamp) of a FULL tile as straight-line code (see conv_tc3.cuh epilogue_slab for the why): all
// loads of a 32-column chunk - eight bias vectors, four residual vectors - are issued while the TMEM load is in flight.
template <typename T, int NCH, bool HAS_RES, bool CLAMP, bool RELU>
__device__ __forceinline__ void epilogue_warp_fast(uint32_t taddr, const float* __restrict__ bias /* + first column of this warp */, uint32_t out_row,
                                                   uint32_t res_row, uint32_t xsw, const EpiBounds& eb) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
    float4 bb[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) bb[g] = __ldg(reinterpret_cast<const float4*>(bias + c * 32 + g * 4));
    uint4 rv[4];
    if constexpr (HAS_RES) {
#pragma unroll
      for (int g = 0; g < 4; ++g) rv[g] = ptx::ld_shared_v4(res_row + ((xsw ^ (uint32_t)(c * 4 + g)) << 4));
    }
    ptx::tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 b0 = bb[2 * g], b1 = bb[2 * g + 1];
      float f[8];
      f[0] = __uint_as_float(v[g * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
      f[2] = __uint_as_float(v[g * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
      f[4] = __uint_as_float(v[g * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
      f[6] = __uint_as_float(v[g * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
      if constexpr (CLAMP) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fminf(fmaxf(f[e], eb.lo1), eb.hi1);
      }
      if constexpr (HAS_RES) {
        float r[8];
        unpack2<T>(rv[g].x, r[0], r[1]); unpack2<T>(rv[g].y, r[2], r[3]);
        unpack2<T>(rv[g].z, r[4], r[5]); unpack2<T>(rv[g].w, r[6], r[7]);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += r[e];
        if constexpr (CLAMP) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fminf(fmaxf(f[e], eb.lo2), eb.hi2);
        }
      }
      const uint32_t dst = out_row + ((xsw ^ (uint32_t)(c * 4 + g)) << 4);
      if constexpr (RELU) ptx::st_shared_v4(dst, make_uint4(pack2_relu<T>(f[0], f[1]), pack2_relu<T>(f[2], f[3]), pack2_relu<T>(f[4], f[5]), pack2_relu<T>(f[6], f[7])));
      else ptx::st_shared_v4(dst, make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7])));
    }
  }
}

template <typename T, int BN, bool GEN>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc2_kernel(const __grid_constant__ Tc2Args args) {
  constexpr int B_BYTES = BN * 128;
  constexpr int SLABS = BN / 64;
  constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  constexpr int CPW = BN / 2;            // accumulator columns per epilogue warp
  const ConvParams& p = args.p;
  const bool has_res = p.res != nullptr;
  const int nkb = args.p.Ktot / BK;
  const bool ws = args.ws != 0;
  const SmemPlan plan = smem_plan<BN>(has_res, ws ? nkb * B_BYTES : 0);
  const int STAGES = plan.stages;
  const uint32_t stage_bytes = plan.stage_bytes;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_al + plan.bar_off);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;
  uint64_t* rempty_bar = rfull_bar + 2;
  uint64_t* wfull_bar = rempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = args.num_tiles, n_tiles = args.n_tiles;
  const bool a_tma = args.a_tma != 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], a_tma ? 1u : (uint32_t)NUM_A_PRODUCERS + 1u);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], NUM_EPI_WARPS);
      ptx::mbar_init(&rfull_bar[a], 1);
      ptx::mbar_init(&rempty_bar[a], NUM_EPI_WARPS);
    }
    ptx::mbar_init(wfull_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == TMA_WARP && lane == 0) {
    ptx::tma_prefetch_desc(&args.tm_w);
    ptx::tma_prefetch_desc(&args.tm_y);
    if (a_tma) ptx::tma_prefetch_desc(&args.tm_a);
    if (has_res) ptx::tma_prefetch_desc(&args.tm_r);
  }
  if (warp == MMA_WARP) {
    ptx::tmem_alloc(tmem_slot, 2 * BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // programmatic dependent launch (see conv_tc3.cuh): the prologue above overlapped the previous kernel's drain; every role waits for the
  // dependency itself, the TMA issuer only after it has fetched the resident (constant) weights
  ptx::grid_launch_dependents();
  if (warp != TMA_WARP && !(args.debug & 4)) ptx::grid_dependency_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================== A producers: cp.async gather (skipped when TMA feeds A) =====================
    if (!a_tma) {
      const T* __restrict__ xg = reinterpret_cast<const T*>(p.x);
      const int t = threadIdx.x;
      const int c = t & 7;          // 16-byte chunk within the 128-byte K row
      const int r0 = t >> 3;        // first row of this thread (rows r0 + 16*i)
      const uint32_t sw_off = (uint32_t)((c ^ (r0 & 7)) << 4);
      const int HoWo = p.Ho * p.Wo;
      const int cpc = p.Cin_p >> 3;       // 16-byte chunks per filter tap
      const int taps = p.kh * p.kw;
      uint32_t kbg = 0;                   // k-blocks issued so far (all tiles)
      for (int t_ = blockIdx.x; t_ < num_tiles; t_ += gridDim.x) {
        const int tile = args.reverse ? num_tiles - 1 - t_ : t_;
        const int m0 = (tile / n_tiles) * BM;
        int hi0[8], wi0[8];
        int64_t xoff[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int m = m0 + r0 + 16 * i;
          bool rv = m < p.M;
          int mm = rv ? m : 0;
          int n = mm / HoWo;
          int rem = mm - n * HoWo;
          int ho = rem / p.Wo;
          int wo = rem - ho * p.Wo;
          hi0[i] = rv ? ho * p.sh - p.pt : -(1 << 28);   // invalid rows never pass the bounds test
          wi0[i] = wo * p.sw - p.pl;
          xoff[i] = (int64_t)n * p.H * p.W * p.Cin_p;
        }
        int cc = c, tap = 0, fr = 0, fs = 0;  // chunk within tap, tap index, filter row/col
        while (cc >= cpc) { cc -= cpc; ++tap; if (++fs == p.kw) { fs = 0; ++fr; } }
        for (int kb = 0; kb < nkb + LAG; ++kb) {
          if (kb < nkb) {
            const uint32_t g = kbg + (uint32_t)kb;
            const int s = (int)(g % (uint32_t)STAGES);
            ptx::mbar_wait(&empty_bar[s], ((g / (uint32_t)STAGES) & 1) ^ 1);
            const uint32_t a_dst = smem_base + s * stage_bytes + r0 * 128 + sw_off;
            const bool kvalid = tap < taps;
            const int dh_off = fr * p.dh, dw_off = fs * p.dw;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              int hi = hi0[i] + dh_off, wi = wi0[i] + dw_off;
              bool ok = kvalid && (unsigned)hi < (unsigned)p.H && (unsigned)wi < (unsigned)p.W;
              const T* src = ok ? xg + xoff[i] + ((int64_t)hi * p.W + wi) * p.Cin_p + cc * 8 : xg;
              ptx::cp_async16(a_dst + i * (16 * 128), src, ok ? 16u : 0u);
            }
            cc += 8;
            while (cc >= cpc) { cc -= cpc; ++tap; if (++fs == p.kw) { fs = 0; ++fr; } }
          }
          ptx::cp_async_commit();
          if (kb >= LAG) {
            ptx::cp_async_wait<LAG>();
            ptx::fence_proxy_async();
            ptx::mbar_arrive(&full_bar[(kbg + (uint32_t)(kb - LAG)) % (uint32_t)STAGES]);
          }
        }
        kbg += (uint32_t)nkb;
      }
    }
  } else if (warp < 4 + NUM_EPI_WARPS) {
    // ===================== epilogue (8 warps: 4 TMEM lane quarters x 2 column halves) =====================
    const int ew = warp - 4;
    const int q = ew & 3;                       // TMEM lane quarter (== warp % 4)
    const int half = ew >> 2;                   // which half of the tile's columns
    const int row = q * 32 + lane;
    const int col0 = half * CPW;                // first column of this warp within the tile
    const uint32_t slab_off = (uint32_t)(col0 >> 6) * SLAB_BYTES;
    const uint32_t xsw = (uint32_t)(((col0 >> 3) & 7) ^ (row & 7));
    const bool store_thread = threadIdx.x == 4 * 32;
    EpiBounds eb;
    const int variant = epi_variant(has_res, args.e1, args.e2, eb);
    uint32_t it = 0;
    for (int t_ = blockIdx.x; t_ < num_tiles; t_ += gridDim.x, ++it) {
      const int tile = args.reverse ? num_tiles - 1 - t_ : t_;
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      const int ncols = min(BN, p.Cout_p - n0);           // valid output channels of this tile (multiple of 8)
      const int nslab = (ncols + 63) >> 6;
      const uint32_t acc = it & 1u;
      const uint32_t par = (it >> 1) & 1u;
      const uint32_t out_tile = smem_base + plan.out_off + acc * TILE_BYTES;
      const uint32_t out_row = out_tile + slab_off + (uint32_t)row * 128u;
      const uint32_t res_row = smem_base + plan.res_off + acc * TILE_BYTES + slab_off + (uint32_t)row * 128u;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)col0;

      ptx::mbar_wait(&tfull_bar[acc], par);                // accumulator complete
      ptx::tc_fence_after();
      if (has_res) ptx::mbar_wait(&rfull_bar[acc], par);   // residual tile landed
      if (store_thread) ptx::bulk_wait_read<1>();          // the store issued two tiles ago has finished reading this staging buffer
      ptx::named_bar_sync(1, NUM_EPI_WARPS * 32);
      if (!GEN && ncols == BN) {
        const float* bw = p.bias + n0 + col0;
        if (has_res) {
          if (variant == 0) epilogue_warp_fast<T, CPW / 32, true, false, false>(taddr, bw, out_row, res_row, xsw, eb);
          else if (variant == 1) epilogue_warp_fast<T, CPW / 32, true, false, true>(taddr, bw, out_row, res_row, xsw, eb);
          else epilogue_warp_fast<T, CPW / 32, true, true, false>(taddr, bw, out_row, res_row, xsw, eb);
        } else {
          if (variant == 0) epilogue_warp_fast<T, CPW / 32, false, false, false>(taddr, bw, out_row, res_row, xsw, eb);
          else if (variant == 1) epilogue_warp_fast<T, CPW / 32, false, false, true>(taddr, bw, out_row, res_row, xsw, eb);
          else epilogue_warp_fast<T, CPW / 32, false, true, false>(taddr, bw, out_row, res_row, xsw, eb);
        }
      } else if (ncols == BN) epilogue_warp<T, CPW / 32, GEN, true>(taddr, p.bias + n0, out_row, res_row, xsw, col0, ncols, has_res, args.e1, args.e2);
      else epilogue_warp<T, CPW / 32, GEN, false>(taddr, p.bias + n0, out_row, res_row, xsw, col0, ncols, has_res, args.e1, args.e2);
      // accumulator + residual buffers are free again
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&tempty_bar[acc]);
        if (has_res) ptx::mbar_arrive(&rempty_bar[acc]);
      }
      ptx::fence_proxy_async();                            // staging writes -> visible to the TMA engine
      ptx::named_bar_sync(1, NUM_EPI_WARPS * 32);
      if (store_thread) {
        if (p.y_heads_inv) {                               // rows (n, h, s) -> y[n][s][h][:]; the tile's 128 rows share (n, h)
          const int g = m0 / p.y_heads_s, nb = g / p.y_heads_h;
          ptx::tma_store_4d(&args.tm_y, out_tile, 0, g - nb * p.y_heads_h, m0 - g * p.y_heads_s, nb);
        } else
        for (int h = 0; h < nslab; ++h) ptx::tma_store_2d(&args.tm_y, out_tile + (uint32_t)h * SLAB_BYTES, n0 + h * 64, m0);
        ptx::bulk_commit();
      }
    }
    if (store_thread) ptx::bulk_wait_read<0>();   // the staging tiles outlive the stores' reads; the writes complete with the grid
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // MN-major B (grouped [K][64] weights): bit 16 of the instruction descriptor; the tile is 64 k-rows of 128 swizzled bytes, the same
      // 8-row / 1024-byte group structure as the K-major tile, and one K = 16 step is two such groups = 2048 bytes further
      const uint32_t idesc = ptx::make_idesc_f16(TcType<T>::is_bf16, BM, BN) | (p.w_mn_major ? (1u << 16) : 0u);
      const uint32_t b_kstep = p.w_mn_major ? 2048u : 32u;
      uint32_t it = 0;
      int s = 0;
      uint32_t ph = 0;
      for (int t_ = blockIdx.x; t_ < num_tiles; t_ += gridDim.x, ++it) {
        const uint32_t acc = it & 1u;
        ptx::mbar_wait(&tempty_bar[acc], ((it >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        if (ws && it == 0) ptx::mbar_wait(wfull_bar, 0);   // resident weights have landed
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = smem_base + s * stage_bytes;
          const uint32_t b_addr = ws ? smem_base + plan.w_off + (uint32_t)kb * B_BYTES : a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(a_addr + k * 32);
            const uint64_t db = ptx::make_smem_desc_sw128(b_addr + (uint32_t)k * b_kstep);
            ptx::umma_f16(d_tmem, da, db, idesc, (kb | k) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[s]);     // frees this stage once the MMAs above have read it
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        ptx::umma_commit(&tfull_bar[acc]);     // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== TMA issuer (one thread): weights, A (matrix / im2col modes), residual =====================
    if (lane == 0) {
      uint32_t it = 0;
      int s = 0;
      uint32_t ph = 0;
      const int cpt = p.Cin_p >> 6;                 // 64-channel chunks per filter tap (im2col mode)
      const int HoWo = p.Ho * p.Wo;
      if ((p.w_dynamic || p.w_grp_m) && !(args.debug & 4)) ptx::grid_dependency_wait();
      if (ws) {                                     // one N tile: fetch the whole weight matrix once
        ptx::mbar_arrive_expect_tx(wfull_bar, (uint32_t)nkb * B_BYTES);
        for (int kb = 0; kb < nkb; ++kb) ptx::tma_load_2d(smem_base + plan.w_off + (uint32_t)kb * B_BYTES, &args.tm_w, wfull_bar, kb * BK, 0);
      }
      if (!(args.debug & 4)) ptx::grid_dependency_wait();
      const uint32_t stage_tx = (a_tma ? (uint32_t)A_BYTES : 0u) + (ws ? 0u : (uint32_t)B_BYTES);
      for (int t_ = blockIdx.x; t_ < num_tiles; t_ += gridDim.x, ++it) {
        const int tile = args.reverse ? num_tiles - 1 - t_ : t_;
        const int m0 = (tile / n_tiles) * BM;
        const int n0 = (tile % n_tiles) * BN;
        const int w_row0 = p.w_grp_m ? (m0 / p.w_grp_m) * (p.w_mn_major ? p.Ktot : p.w_grp_rows) : 0;   // grouped weights: this tile's weight matrix
        int bn = 0, bh = 0, bw = 0;
        if (args.a_tma == 2) {                       // first output pixel of the tile -> base pixel of the im2col bounding box
          bn = m0 / HoWo;
          const int rem = m0 - bn * HoWo;
          const int ho = rem / p.Wo;
          bh = ho * p.sh - p.pt;
          bw = (rem - ho * p.Wo) * p.sw - p.pl;
        }
        if (has_res) {
          const uint32_t rb = it & 1u;
          const int ncols = min(BN, p.Cout_p - n0);
          const int nslab = (ncols + 63) >> 6;
          ptx::mbar_wait(&rempty_bar[rb], ((it >> 1) & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(&rfull_bar[rb], (uint32_t)nslab * SLAB_BYTES);
          for (int h = 0; h < nslab; ++h)
            ptx::tma_load_2d(smem_base + plan.res_off + rb * TILE_BYTES + (uint32_t)h * SLAB_BYTES, &args.tm_r, &rfull_bar[rb], n0 + h * 64, m0);
        }
        int tap = 0, chunk = 0, fr = 0, fs = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
          const uint32_t a_dst = smem_base + s * stage_bytes;
          if (stage_tx) ptx::mbar_arrive_expect_tx(&full_bar[s], stage_tx);
          else ptx::mbar_arrive(&full_bar[s]);
          if (args.a_tma == 1) {
            ptx::tma_load_2d(a_dst, &args.tm_a, &full_bar[s], kb * BK, m0);
          } else if (args.a_tma == 2) {
            ptx::tma_load_im2col_4d(a_dst, &args.tm_a, &full_bar[s], chunk * BK, bw, bh, bn, (uint32_t)(fs * p.dw), (uint32_t)(fr * p.dh));
            if (++chunk == cpt) { chunk = 0; ++tap; if (++fs == p.kw) { fs = 0; ++fr; } }
          }
          if (p.w_mn_major) ptx::tma_load_2d(a_dst + A_BYTES, &args.tm_w, &full_bar[s], 0, w_row0 + kb * BK);   // 64 k-rows of 64 channels
          else if (!ws) ptx::tma_load_2d(a_dst + A_BYTES, &args.tm_w, &full_bar[s], kb * BK, n0 + w_row0);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) ptx::tmem_dealloc(tmem_base, 2 * BN);
}

}  // namespace

// ---- host side (tensor-map builders shared with conv_tc3.cu) -------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 encode_tiled_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) ptr = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }();
  return fn;
}

// [outer][inner] row-major matrix of 2-byte elements, 128B-swizzled boxes of box_inner x box_outer
CUtensorMap make_tmap_2d(const void* base, bool bf16, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner,
                         uint32_t box_outer) {
  auto enc = encode_tiled_fn();
  B200_CHECK(enc, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled failed (" << (int)r << ") for a [" << outer << "][" << inner << "] matrix, pitch "
                                                                                  << pitch_bytes << " B, base " << base);
  return m;
}

PFN_cuTensorMapEncodeIm2col_v12000 encode_im2col_fn() {
  static PFN_cuTensorMapEncodeIm2col_v12000 fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) ptr = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(ptr);
  }();
  return fn;
}

// NHWC activations as a rank-4 {C, W, H, N} tensor traversed in im2col mode: 128 output pixels x 64 channels per box.
// Bounding box: lower corner = -padding, upper corner chosen so that exactly Wo x Ho base pixels exist per image at the
// convolution's stride (traversal stride).
CUtensorMap make_tmap_im2col(const ConvParams& p, bool bf16) {
  auto enc = encode_im2col_fn();
  B200_CHECK(enc, B200_ERR_CUDA, "cuTensorMapEncodeIm2col is not available from this driver");
  const int64_t sw = p.x_sw ? p.x_sw : p.Cin_p, sh = p.x_sh ? p.x_sh : (int64_t)p.W * p.Cin_p, sn = p.x_sn ? p.x_sn : (int64_t)p.H * p.W * p.Cin_p;
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)p.Cin_p, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
  int lower[2] = {-p.pl, -p.pt};
  int upper[2] = {-p.pl + (p.Wo - 1) * p.sw - (p.W - 1), -p.pt + (p.Ho - 1) * p.sh - (p.H - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)p.sw, (cuuint32_t)p.sh, 1};
  CUresult r = enc(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(p.x), dims, strides, lower, upper,
                   (cuuint32_t)BK, (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeIm2col failed (" << (int)r << ") for a " << p.N << "x" << p.H << "x" << p.W << "x" << p.Cin_p
                                                                                   << " tensor, corners [" << lower[0] << "," << lower[1] << "]/[" << upper[0]
                                                                                   << "," << upper[1] << "]");
  return m;
}

namespace {

int g_sm_count = 0;
int g_im2col = -1;   // -1 unread; B200_IM2COL=0 disables the im2col-mode TMA path (cp.async gather instead)

}  // namespace

tc::EpiAct make_epi(int kind, float a, float b) {
  tc::EpiAct e;
  e.relu_in_pack = 0;
  e.kind = kind; e.a = a; e.b = b;
  e.lo = kind == ACT_RELU ? 0.f : a;
  e.hi = kind == ACT_RELU ? __builtin_huge_valf() : b;
  return e;
}

namespace {
using namespace tc;

template <typename T, int BN>
void launch_tc2(const ConvParams& p, bool bf16, cudaStream_t s) {
  Tc2Args a;
  memset(&a, 0, sizeof(a));
  a.p = p;
  a.e1 = make_epi(p.act1, p.a1, p.b1);
  a.e2 = make_epi(p.act2, p.a2, p.b2);
  {
    // a trailing ReLU rides on the float -> bf16/f16 conversion (B200_RELU_PACK=0: explicit max), as in the pair kernel
    const char* e = getenv("B200_RELU_PACK");
    const bool on = !(e && e[0] == '0');
    if (on && p.res != nullptr && p.act2 == ACT_RELU) a.e2.relu_in_pack = 1;
    if (on && p.res == nullptr && p.act1 == ACT_RELU) a.e1.relu_in_pack = 1;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  a.n_tiles = p.Cout_alloc / BN;
  a.num_tiles = m_tiles * a.n_tiles;
  a.reverse = p.reverse;
  { const char* e = getenv("B200_DEBUG_TC3"); a.debug = e ? atoi(e) : 0; }
  {
    const int ws_bytes = (p.Ktot / BK) * BN * 128;
    const char* no_ws = getenv("B200_NO_WS");
    a.ws = (a.n_tiles == 1 && p.w_grp_m == 0 && ws_bytes <= WS_MAX_BYTES && smem_plan<BN>(p.res != nullptr, ws_bytes).stages >= 4 && !(no_ws && no_ws[0] == '1')) ? 1 : 0;
  }
  a.a_tma = conv_tc2_a_is_matrix(p) ? 1 : (conv_tc2_a_is_im2col(p) ? 2 : 0);
  a.tm_w = make_tmap_2d(p.w, bf16, (uint64_t)p.Ktot, p.w_grp_m ? (uint64_t)(p.M / p.w_grp_m) * p.w_grp_rows : (uint64_t)p.Cout_alloc, (uint64_t)p.w_ld * 2, BK, BN);
  if (p.w_mn_major) {
    B200_CHECK(BN == 64 && p.Cout_p == 64 && p.Cout_alloc == 64 && p.w_grp_m > 0, B200_ERR_INVALID, "conv_tc2: MN-major weights need grouped 64-column launches");
    a.tm_w = make_tmap_2d(p.w, bf16, 64, (uint64_t)(p.M / p.w_grp_m) * p.Ktot, 128, 64, BK);
  }
  a.tm_y = make_tmap_2d(p.y, bf16, (uint64_t)p.Cout_p, (uint64_t)p.M, (uint64_t)p.ldy * 2, 64, BM);
  a.tm_r = p.res ? make_tmap_2d(p.res, bf16, (uint64_t)p.Cout_p, (uint64_t)p.M, (uint64_t)p.ldy * 2, 64, BM) : a.tm_y;
  if (p.y_heads_inv) {
    B200_CHECK(p.y_heads_s % BM == 0 && p.w_grp_m == p.y_heads_s && p.Cout_p == 64 && p.M % (p.y_heads_s * p.y_heads_h) == 0 && !p.res, B200_ERR_INVALID,
               "conv_tc2: bad merged-heads output geometry");
    auto enc = encode_tiled_fn();
    B200_CHECK(enc, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t S = (cuuint64_t)p.y_heads_s, H = (cuuint64_t)p.y_heads_h;
    cuuint64_t dims[4] = {64, H, S, (cuuint64_t)p.M / (S * H)};
    cuuint64_t strides[3] = {128, H * 128, S * H * 128};
    cuuint32_t box[4] = {64, 1, (cuuint32_t)BM, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&a.tm_y, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(p.y), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled (merged heads) failed (" << (int)r << ")");
  }
  a.tm_a = a.a_tma == 1 ? make_tmap_2d(p.x, bf16, (uint64_t)p.Cin_p, (uint64_t)p.M, (uint64_t)p.Cin_p * 2, BK, BM)
                        : (a.a_tma == 2 ? make_tmap_im2col(p, bf16) : a.tm_y);
  const int grid = std::min(a.num_tiles, g_sm_count > 0 ? g_sm_count : 148);
  const unsigned smem = smem_plan<BN>(p.res != nullptr, a.ws ? (p.Ktot / BK) * BN * 128 : 0).total;
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
  if (conv_acts_are_clamps(p)) cudaLaunchKernelEx(&cfg, conv_tc2_kernel<T, BN, false>, a);
  else cudaLaunchKernelEx(&cfg, conv_tc2_kernel<T, BN, true>, a);
}

}  // namespace

// A is a plain [M][Cin_p] matrix: FullyConnected, or a 1x1 / stride-1 / unpadded convolution over NHWC
bool conv_tc2_a_is_matrix(const ConvParams& p) {
  return p.kh == 1 && p.kw == 1 && p.sh == 1 && p.sw == 1 && p.pt == 0 && p.pl == 0 && p.Ho == p.H && p.Wo == p.W && p.Cin_p % 8 == 0 && p.x_sw == 0;
}

bool conv_tc2_im2col_enabled() {
  if (g_im2col < 0) {
    const char* e = getenv("B200_IM2COL");
    g_im2col = (e && e[0] == '0') ? 0 : 1;
  }
  return g_im2col != 0;
}

// The space-to-depth stem reads overlapping 64-element windows (pixel stride 32 B < 128 B of channels).  Ask the driver once whether
// it accepts such a tensor map (a dummy, 1 KiB aligned base address is enough: encoding never dereferences it).
bool conv_tc2_s2d_encode_ok() {
  static int ok = -1;
  if (ok >= 0) return ok != 0;
  auto enc = encode_im2col_fn();
  if (!enc) return (ok = 0) != 0;
  CUtensorMap m;
  cuuint64_t dims[4] = {64, 112, 115, 2};
  cuuint64_t strides[3] = {32, 115 * 32, 115 * 115 * 32};
  int lower[2] = {0, 0}, upper[2] = {0, -3};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, reinterpret_cast<void*>(uintptr_t(1) << 30), dims, strides, lower, upper, 64, 128, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ok = r == CUDA_SUCCESS ? 1 : 0;
  return ok != 0;
}

bool conv_tc2_a_is_im2col(const ConvParams& p) {
  if (!conv_tc2_im2col_enabled() || conv_tc2_a_is_matrix(p)) return false;
  const int uw = -p.pl + (p.Wo - 1) * p.sw - (p.W - 1), uh = -p.pt + (p.Ho - 1) * p.sh - (p.H - 1);
  return p.Cin_p % 64 == 0 && p.pl <= 128 && p.pt <= 128 && uw >= -128 && uw <= 127 && uh >= -128 && uh <= 127 && p.sw <= 8 && p.sh <= 8 &&
         (p.kw - 1) * p.dw < 65536 && (p.kh - 1) * p.dh < 65536;
}

bool conv_acts_are_clamps(const ConvParams& p) {
  auto ok = [](int k) { return k == ACT_NONE || k == ACT_RELU || k == ACT_CLAMP; };
  return ok(p.act1) && ok(p.act2);
}

bool conv_tc2_supported(const ConvParams& p) {
  // one weight matrix for the whole launch (batched activation x activation MatMuls stay on conv_tc.cu), weights fully materialised
  if (p.w_grp_m && (p.w_grp_m % 128 || p.M % p.w_grp_m || p.w_grp_rows != p.Cout_alloc)) return false;   // tiles must not straddle groups
  return p.batch == 1 && !p.out_f32 && p.w_rows == p.Cout_alloc && p.w_k == p.Ktot && p.w_ld == p.Ktot;
}

void conv_tc2_init_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  auto opt_in = [](auto kernel, int bytes) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); };
  opt_in(conv_tc2_kernel<__nv_bfloat16, 64, false>, SMEM_LIMIT);
  opt_in(conv_tc2_kernel<__nv_bfloat16, 64, true>, SMEM_LIMIT);
  opt_in(conv_tc2_kernel<__nv_bfloat16, 128, false>, SMEM_LIMIT);
  opt_in(conv_tc2_kernel<__nv_bfloat16, 128, true>, SMEM_LIMIT);
  opt_in(conv_tc2_kernel<__half, 64, false>, SMEM_LIMIT);
  opt_in(conv_tc2_kernel<__half, 64, true>, SMEM_LIMIT);
  opt_in(conv_tc2_kernel<__half, 128, false>, SMEM_LIMIT);
  opt_in(conv_tc2_kernel<__half, 128, true>, SMEM_LIMIT);
}

void launch_conv_tc2(const ConvParams& p, int dtype, int bn, cudaStream_t s) {
  B200_CHECK(conv_tc2_supported(p) && !(p.y_heads_s && !p.y_heads_inv), B200_ERR_UNSUPPORTED,
             "conv_tc2: unsupported launch (batched or partial weights, head-split stores)");
  B200_CHECK(p.Ktot % BK == 0 && p.Cin_p % 8 == 0 && p.Cout_alloc % bn == 0 && p.Cout_p % 8 == 0 && p.ldy % 8 == 0, B200_ERR_INVALID,
             "conv_tc2: bad geometry");
  const bool bf16 = dtype == B200_BF16;
  B200_CHECK(bf16 || dtype == B200_F16, B200_ERR_UNSUPPORTED, "conv_tc2: dtype must be bf16/f16");
  if (bf16) {
    if (bn == 64) launch_tc2<__nv_bfloat16, 64>(p, true, s); else launch_tc2<__nv_bfloat16, 128>(p, true, s);
  } else {
    if (bn == 64) launch_tc2<__half, 64>(p, false, s); else launch_tc2<__half, 128>(p, false, s);
  }
}

}  // namespace b200
