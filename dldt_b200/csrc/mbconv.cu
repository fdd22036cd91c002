// Inverted-residual block (MobileNetV2 style) as ONE persistent tcgen05 kernel (sm_100a):
//
//   y = project_1x1( clamp( depthwise_3x3( clamp( expand_1x1(x) + b1 ) ) + bd ) ) + b2 (+ x)
//
// The reference runs the three convolutions as three primitives with the BN / ReLU6 / Add folded in as post-ops
// (mkldnn_graph_optimizer.cpp FuseConvolutionAndDWConvolution / FuseConvolutionSumAndConvolutionSumActivation; arithmetic per
// ngraph reference/convolution.hpp:89-186 and group convolution); here the 6x-expanded tensor never leaves the SM.  Per output tile of
// TH x 16 pixels (TH = 8 at stride 1, 4 at stride 2) and per 64-channel chunk of the hidden dimension:
//   1. expand:  D_e[(TH-1)s+3 x (16-1)s+3 halo pixels, 64] = X_tile[pixels, Cin] * W1[chunk]^T   tcgen05.mma, X tile loaded ONCE per tile by TMA
//   2. the eight compute warps read D_e from TMEM, add b1, clamp, zero what lies outside the image (the depthwise's zero padding is
//      padding of the EXPANDED tensor) and write the 16-bit chunk to shared memory in the 128-byte-swizzled pixel-row layout
//   3. depthwise 3x3 on CUDA cores from that tile (fp32 accumulation, same tap order as dwconv3x3_roll_kernel), + bd, clamp, 16-bit pack,
//      written as the K-major A operand of
//   4. project: D_p[TH x 16 pixels, Cout] += A2[pixels, 64] * W2[:, chunk]^T                       tcgen05.mma, accumulated over the chunks
// and after the last chunk D_p + b2 (+ the centre pixels of the X tile: the residual) goes out through one TMA store.  Without an expand
// convolution (the first block, t = 1) step 1/2 is a TMA load of the chunk straight into the step-2 layout.
//
//   warp 0-3  tile epilogue (the compute warps go straight on to the next tile)      warp 12  TMA producer (X tile double-buffered, weights
//   warp 4-11 compute: expand accumulators -> hidden chunk -> depthwise -> A2                  resident, where they fit)
//   warp 13   MMA issuer + TMEM alloc
// (Variants measured on MobileNetV2 b64: dedicated warps for step 2 running ahead of the depthwise warps were SLOWER (0.91 vs 0.84 ms): one
// warp per scheduler is latency-bound on the TMEM -> clamp -> pack chain; ncu shows the kernel issue-bound at ~11k warp instructions per
// 64-channel chunk, 23 per depthwise output.)
//
// HBM bytes per block: x + y + weights (MobileNetV2 b64: 1.72 GB -> ~0.4 GB per inference batch), 17 x 3 kernels -> 17.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "common.h"
#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tc_common.cuh"

namespace b200 {

// defined in conv_tc2.cu
CUtensorMap make_tmap_2d(const void* base, bool bf16, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer);
PFN_cuTensorMapEncodeTiled_v12000 encode_tiled_fn();

namespace {
using namespace tc;

constexpr int MB_E_WARPS = 4;             // warps 0-3: expand accumulators -> hidden chunk (TMEM lane quarter = warp)
constexpr int MB_D_WARPS = 8;             // warps 4-11: depthwise stage + tile epilogue (TMEM lane quarter = warp % 4)
constexpr int MB_TMA_WARP = 12, MB_MMA_WARP = 13;
constexpr int MB_THREADS = 448;
constexpr int MB_D_THREADS = MB_D_WARPS * 32;
constexpr int MB_TW = 16;
constexpr int MB_SMEM_LIMIT = 232448;

struct MbGeom { int TH, IH, IW, rows_in, MT, kb1, nchunk; };
__host__ __device__ inline MbGeom mb_geom(int stride, int Cin_p, int Ch_p) {
  MbGeom g;
  g.TH = stride == 1 ? 8 : 4;
  g.IH = (g.TH - 1) * stride + 3;
  g.IW = (MB_TW - 1) * stride + 3;
  g.rows_in = g.IH * g.IW;
  g.MT = (g.rows_in + 127) / 128;
  g.kb1 = (Cin_p + 63) / 64;
  g.nchunk = Ch_p / 64;
  return g;
}

constexpr int MB_MAXW = 8;                // weight ring: at most this many 64-channel chunks of W1 / W2 in shared memory
struct MbPlan { uint32_t nxb, nw, nh, xa_kb, xa_bytes, xa_off, h_bytes, h_off, w1_bytes, w1_off, w2_bytes, w2_off, a2_off, out_off, out_bytes, wd_off, b1_off, bd_off, b2_off, bar_off, total; };
__host__ __device__ inline MbPlan mb_plan_n(const MbGeom& g, int has_expand, int Ch_p, int Cout_p, int Cout_alloc, uint32_t nxb, uint32_t nw, uint32_t nh = 0) {
  MbPlan p;
  uint32_t o = 0;
  p.nxb = nxb;
  p.nw = nw;
  p.xa_kb = (uint32_t)g.MT * 128u * 128u;                    // one 64-channel k-block of the X tile: MT x 128 rows of 128 bytes
  p.xa_bytes = (uint32_t)g.kb1 * p.xa_kb;
  p.xa_off = o; o += has_expand ? nxb * p.xa_bytes : 0u;
  p.h_bytes = ((uint32_t)g.rows_in * 128u + 1023u) & ~1023u;
  // hidden-chunk ring: two buffers behind the expand stage (the compute warps fill and read them themselves); four when the TMA fills
  // them (no expand convolution): with two, ONE load was in flight and every tile waited out a full TMA + DRAM latency (ncu / timing:
  // 3350 cycles per tile whatever the depthwise stage cost)
  p.nh = nh ? nh : (has_expand ? 2u : 4u);
  p.h_off = o; o += p.nh * p.h_bytes;
  p.w1_bytes = (uint32_t)g.kb1 * 64u * 128u;
  p.w1_off = o; o += has_expand ? nw * p.w1_bytes : 0u;
  p.w2_bytes = (uint32_t)Cout_alloc * 128u;
  p.w2_off = o; o += nw * p.w2_bytes;
  o = (o + 1023u) & ~1023u;
  p.a2_off = o; o += 128u * 128u;
  p.out_bytes = (128u * (uint32_t)Cout_p * 2u + 127u) & ~127u;
  p.out_off = o; o += p.out_bytes;
  p.wd_off = o; o += (uint32_t)(9 * Ch_p * 2);               // depthwise weights [9][Ch_p], 16-bit (fp32 copies: no faster, and the 576-wide blocks no longer fit)
  p.b1_off = o; o += (uint32_t)Ch_p * 4u;
  p.bd_off = o; o += (uint32_t)Ch_p * 4u;
  p.b2_off = o; o += (uint32_t)Cout_alloc * 4u;
  o = (o + 15u) & ~15u;
  p.bar_off = o; o += 512u;                                  // 40 mbarriers + the TMEM slot
  p.total = o + 1024u;                                       // alignment slack of the dynamic shared-memory base
  return p;
}
// Shared-memory budget, in order of what it buys: (1) ALL weight chunks resident (nw == chunks: loaded once per CTA instead of once per
// tile, and no chunk ever waits for the round trip "project MMAs of chunk c - nw done -> TMA -> expand MMAs of chunk c"), (2) two X tiles
// (the next tile's load overlaps this tile's chunks), (3) as deep a weight ring as fits.
__host__ __device__ inline MbPlan mb_plan(const MbGeom& g, int has_expand, int Ch_p, int Cout_p, int Cout_alloc) {
  const uint32_t nwmax = (uint32_t)(g.nchunk < MB_MAXW ? g.nchunk : MB_MAXW);
  const bool resident_possible = g.nchunk <= MB_MAXW;
  for (uint32_t nw = nwmax; nw >= 1u; --nw) {
    if (nw == 1u && nwmax > 1u) break;
    for (uint32_t nxb = 2u; nxb >= 1u; --nxb) {
      if (nxb == 2u && !has_expand) continue;
      const MbPlan p = mb_plan_n(g, has_expand, Ch_p, Cout_p, Cout_alloc, nxb, nw);
      if (p.total <= (uint32_t)MB_SMEM_LIMIT) return p;
      if (!has_expand) {                                       // wide TMA-fed chunks (576 channels at stride 2): two hidden buffers instead of four
        const MbPlan p2 = mb_plan_n(g, has_expand, Ch_p, Cout_p, Cout_alloc, nxb, nw, 2u);
        if (p2.total <= (uint32_t)MB_SMEM_LIMIT) return p2;
      }
    }
    (void)resident_possible;
  }
  return mb_plan_n(g, has_expand, Ch_p, Cout_p, Cout_alloc, 1u, nwmax > 1u ? 2u : 1u);      // does not fit: mbconv_supported() rejects it
}

struct MbArgs {
  CUtensorMap tm_x;    // {Cin_p | Ch (no expand), W, H, N}, box {64, IW, IH, 1}, SWIZZLE_128B, zero fill outside
  CUtensorMap tm_w1;   // [Ch_p][K1], box {64, 64}
  CUtensorMap tm_w2;   // [Cout_alloc][Ch_p], box {64, Cout_alloc}
  CUtensorMap tm_y;    // {Cout_p, Wo, Ho, N}, box {Cout_p, 16, TH, 1}, no swizzle
  MbParams p;
  int tiles_w, tiles_h, num_tiles;
};

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  unpack2<T>(v.x, f[0], f[1]); unpack2<T>(v.y, f[2], f[3]); unpack2<T>(v.z, f[4], f[5]); unpack2<T>(v.w, f[6], f[7]);
}

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

template <typename T, int STRIDE, bool EXPAND>
__global__ void __launch_bounds__(MB_THREADS, 1) mbconv_kernel(const __grid_constant__ MbArgs args) {
  const MbParams& p = args.p;
  constexpr int TH = STRIDE == 1 ? 8 : 4;
  constexpr int IH = (TH - 1) * STRIDE + 3, IW = (MB_TW - 1) * STRIDE + 3;
  constexpr int ROWS_IN = IH * IW;
  constexpr int MT = (ROWS_IN + 127) / 128;                   // 128-row M tiles of the halo tile (expand accumulators: MT x 64 TMEM columns)
  const MbGeom g = mb_geom(STRIDE, p.Cin_p, p.Ch_p);
  const MbPlan plan = mb_plan(g, EXPAND ? 1 : 0, p.Ch_p, p.Cout_p, p.Cout_alloc);
  const int kb1 = g.kb1, nchunk = g.nchunk;
  const bool two_x = plan.nxb == 2u;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + plan.bar_off);
  uint64_t* xfull = bars + 0;      // [2] X tile landed (TMA)
  uint64_t* xempty = bars + 2;     // [2] X tile no longer read (expand MMAs complete, residual read)
  uint64_t* efull = bars + 4;      // expand accumulators of the chunk complete
  uint64_t* eempty = bars + 5;     // ... drained by the four E warps
  uint64_t* hfull = bars + 32;     // [4] hidden chunk in shared memory (TMA, without an expand convolution)
  uint64_t* hempty = bars + 36;    // [4] hidden chunk consumed by the depthwise stage
  uint64_t* a2full = bars + 10;    // depthwise output chunk staged as the project A operand
  uint64_t* a2empty = bars + 11;   // ... consumed by the project MMAs
  uint64_t* pfull = bars + 12;     // project accumulator of the tile complete
  uint64_t* pempty = bars + 13;    // ... drained
  uint64_t* wfull = bars + 16;     // [MB_MAXW] weight chunk landed
  uint64_t* wempty = bars + 24;    // [MB_MAXW] weight chunk consumed (project MMAs of the chunk complete)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  const uint32_t nw = plan.nw;
  const bool w_resident = nw >= (uint32_t)nchunk;               // every chunk has its own stage: loaded during the first tile only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = args.num_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&xfull[i], 1); ptx::mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < MB_MAXW; ++i) { ptx::mbar_init(&wfull[i], 1); ptx::mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 4; ++i) { ptx::mbar_init(&hfull[i], 1); ptx::mbar_init(&hempty[i], MB_D_WARPS); }
    ptx::mbar_init(efull, 1); ptx::mbar_init(eempty, MB_D_WARPS);
    ptx::mbar_init(a2full, MB_D_WARPS); ptx::mbar_init(a2empty, 1);
    ptx::mbar_init(pfull, 1); ptx::mbar_init(pempty, MB_E_WARPS);
    ptx::fence_mbar_init();
  }
  if (warp == MB_TMA_WARP && lane == 0) {
    ptx::tma_prefetch_desc(&args.tm_x);
    ptx::tma_prefetch_desc(&args.tm_w2);
    ptx::tma_prefetch_desc(&args.tm_y);
    if (EXPAND) ptx::tma_prefetch_desc(&args.tm_w1);
  }
  if (warp == MB_MMA_WARP) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  if (warp < MB_E_WARPS + MB_D_WARPS) {
    // constants -> shared memory (not written by any kernel: fetched before the programmatic-launch dependency resolves)
    const int t = threadIdx.x;
    const uint32_t* wd_g = reinterpret_cast<const uint32_t*>(p.wd);
    uint32_t* wd_s = reinterpret_cast<uint32_t*>(smem_al + plan.wd_off);
    for (int i = t; i < 9 * p.Ch_p / 2; i += 384) wd_s[i] = __ldg(wd_g + i);
    float* b1_s = reinterpret_cast<float*>(smem_al + plan.b1_off);
    float* bd_s = reinterpret_cast<float*>(smem_al + plan.bd_off);
    float* b2_s = reinterpret_cast<float*>(smem_al + plan.b2_off);
    for (int i = t; i < p.Ch_p; i += 384) { b1_s[i] = EXPAND ? __ldg(p.b1 + i) : 0.f; bd_s[i] = __ldg(p.bd + i); }
    for (int i = t; i < p.Cout_alloc; i += 384) b2_s[i] = __ldg(p.b2 + i);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_e = tmem_base, acc_p = tmem_base + 256u;      // columns [0, 192) expand, [256, 512) project
  ptx::grid_launch_dependents();

  if (warp < MB_E_WARPS) {
    // ======================================= epilogue warps =======================================
    const int q = warp;
    const int te = threadIdx.x;                                 // 0 .. 127
    ptx::grid_dependency_wait();
    // ---- tile epilogue (these warps, so that the depthwise warps go straight on to the next tile): project accumulator + b2 (+ residual =
    // centre pixels of the X tile) -> staging -> one TMA store ----
    auto tile_epilogue = [&](int tile, uint32_t it) {
      const int n = tile / (args.tiles_h * args.tiles_w);
      const int rem = tile - n * args.tiles_h * args.tiles_w;
      const int th = rem / args.tiles_w;
      const int oh0 = th * TH, ow0 = (rem - th * args.tiles_w) * MB_TW;
      ptx::mbar_wait(pfull, it & 1u);
      ptx::tc_fence_after();
      if (te == 0) ptx::bulk_wait_read<0>();                     // the previous tile's store has finished reading the staging tile
      ptx::named_bar_sync(2, MB_E_WARPS * 32);
      const int prow = q * 32 + lane;                            // output pixel of the tile (row-major TH x 16)
      const uint32_t stage_row = smem_base + plan.out_off + (uint32_t)prow * (uint32_t)p.Cout_p * 2u;
      const int poh = prow / MB_TW, pow_ = prow - poh * MB_TW;
      const uint32_t xr = (uint32_t)((poh * STRIDE + p.pt) * IW + pow_ * STRIDE + p.pl);     // its centre pixel in the halo tile
      const uint32_t xbuf = smem_base + plan.xa_off + (two_x ? (it & 1u) : 0u) * plan.xa_bytes;
      for (int col = 0; col < p.Cout_alloc; col += 16) {
        uint32_t v[2][8];
        tmem_ld_32x8(acc_p + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v[0]);
        tmem_ld_32x8(acc_p + ((uint32_t)(q * 32) << 16) + (uint32_t)col + 8u, v[1]);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cl = col + hh * 8;
          const uint32_t b2a = smem_base + plan.b2_off + (uint32_t)cl * 4u;
          const uint4 ba = ptx::ld_shared_v4(b2a), bb = ptx::ld_shared_v4(b2a + 16u);
          float f[8];
          f[0] = __uint_as_float(v[hh][0]) + __uint_as_float(ba.x); f[1] = __uint_as_float(v[hh][1]) + __uint_as_float(ba.y);
          f[2] = __uint_as_float(v[hh][2]) + __uint_as_float(ba.z); f[3] = __uint_as_float(v[hh][3]) + __uint_as_float(ba.w);
          f[4] = __uint_as_float(v[hh][4]) + __uint_as_float(bb.x); f[5] = __uint_as_float(v[hh][5]) + __uint_as_float(bb.y);
          f[6] = __uint_as_float(v[hh][6]) + __uint_as_float(bb.z); f[7] = __uint_as_float(v[hh][7]) + __uint_as_float(bb.w);
          if (cl < p.Cout_p && prow < TH * MB_TW) {
            if (EXPAND && p.has_res) {
              float rr[8];
              const uint32_t kb = (uint32_t)cl >> 6, ch = ((uint32_t)cl & 63u) >> 3;
              unpack8<T>(ptx::ld_shared_v4(xbuf + kb * plan.xa_kb + xr * 128u + ((ch ^ (xr & 7u)) << 4)), rr);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += rr[e];
            }
            ptx::st_shared_v4(stage_row + (uint32_t)cl * 2u, make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7])));
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(pempty);
      ptx::fence_proxy_async();
      ptx::named_bar_sync(2, MB_E_WARPS * 32);
      if (te == 0) {
        if (EXPAND && p.has_res) ptx::mbar_arrive(&xempty[two_x ? (it & 1u) : 0u]);   // residual read: the producer may refill this X tile
        ptx::tma_store_4d(&args.tm_y, smem_base + plan.out_off, 0, ow0, oh0, n);
        ptx::bulk_commit();
      }
    };
    {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) tile_epilogue(tile, it);
    }
    if (te == 0) ptx::bulk_wait_read<0>();
  } else if (warp < MB_E_WARPS + MB_D_WARPS) {
    // ======================================= compute warps: expand accumulators -> hidden chunk -> depthwise -> project A operand =======================================
    const int t = threadIdx.x - MB_E_WARPS * 32;
    const int q = warp & 3, half = (warp - MB_E_WARPS) >> 2;    // TMEM lane quarter, column half of the expand accumulators
    // depthwise work split: 16-byte channel group cc (8 channels) x pixel group pg
    constexpr int PXT = STRIDE == 1 ? 4 : 2;                    // output pixels per thread (adjacent in a row)
    constexpr int NIN = (PXT - 1) * STRIDE + 3;                 // input columns they read
    const int cc = t & 7, pg = t >> 3;
    const int orow = STRIDE == 1 ? (pg >> 2) : (pg >> 3);
    const int ocol0 = STRIDE == 1 ? (pg & 3) * 4 : (pg & 7) * 2;
    uint32_t gc = 0;                                            // chunks processed so far (all tiles)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      uint32_t inside = 0;                                      // which of this thread's halo rows (one per M tile) lie inside the image
      if (EXPAND) {
        const int n = tile / (args.tiles_h * args.tiles_w);
        const int rem = tile - n * args.tiles_h * args.tiles_w;
        const int th = rem / args.tiles_w;
        const int ih0 = th * TH * STRIDE - p.pt, iw0 = (rem - th * args.tiles_w) * MB_TW * STRIDE - p.pl;   // image coordinates of halo pixel (0, 0)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const int r = m * 128 + q * 32 + lane;
          const int ih = r / IW, iw = r - ih * IW;
          if (r < ROWS_IN && (unsigned)(ih0 + ih) < (unsigned)p.H && (unsigned)(iw0 + iw) < (unsigned)p.W) inside |= 1u << m;
        }
      }
      for (int c = 0; c < nchunk; ++c, ++gc) {
        const uint32_t hb = gc & (plan.nh - 1u);
        const uint32_t hbuf = smem_base + plan.h_off + hb * plan.h_bytes;
        if (EXPAND) {
          // ---- expand accumulators -> + b1, clamp, zero outside the image (the depthwise's zero padding is padding of the EXPANDED
          // tensor) -> 16-bit hidden chunk in the 128-byte-swizzled pixel-row layout the depthwise stage reads ----
          ptx::mbar_wait(efull, gc & 1u);
          ptx::tc_fence_after();
          const uint32_t b1a = smem_base + plan.b1_off + (uint32_t)(c * 64 + half * 32) * 4u;
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const int r = m * 128 + q * 32 + lane;
            uint32_t v[32];
            ptx::tmem_ld_32x32(acc_e + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * 64 + half * 32), v);
            ptx::tmem_ld_wait();
            if (m == MT - 1) {                                  // this warp's part of the chunk is in registers
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(eempty);
            }
            // outside the image both bounds are 0: the clamp writes the depthwise's zero padding
            const float lo = ((inside >> m) & 1u) ? p.lo1 : 0.f, hi = ((inside >> m) & 1u) ? p.hi1 : 0.f;
            if (r < ROWS_IN) {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const uint4 ba = ptx::ld_shared_v4(b1a + (uint32_t)jj * 32u), bb = ptx::ld_shared_v4(b1a + (uint32_t)jj * 32u + 16u);
                float f[8];
                f[0] = __uint_as_float(v[jj * 8 + 0]) + __uint_as_float(ba.x); f[1] = __uint_as_float(v[jj * 8 + 1]) + __uint_as_float(ba.y);
                f[2] = __uint_as_float(v[jj * 8 + 2]) + __uint_as_float(ba.z); f[3] = __uint_as_float(v[jj * 8 + 3]) + __uint_as_float(ba.w);
                f[4] = __uint_as_float(v[jj * 8 + 4]) + __uint_as_float(bb.x); f[5] = __uint_as_float(v[jj * 8 + 5]) + __uint_as_float(bb.y);
                f[6] = __uint_as_float(v[jj * 8 + 6]) + __uint_as_float(bb.z); f[7] = __uint_as_float(v[jj * 8 + 7]) + __uint_as_float(bb.w);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fminf(fmaxf(f[e], lo), hi);
                ptx::st_shared_v4(hbuf + (uint32_t)r * 128u + ((((uint32_t)(half * 4 + jj)) ^ ((uint32_t)r & 7u)) << 4),
                                  make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7])));
              }
            }
          }
          ptx::named_bar_sync(1, MB_D_THREADS);                   // the whole chunk is in shared memory (two buffers: the chunk before last is read out)
        } else {
          ptx::mbar_wait(&hfull[hb], (gc >> (plan.nh == 4u ? 2 : 1)) & 1u);
        }
        // ---- depthwise 3x3: PXT adjacent outputs x 8 channels per thread, fp32 accumulation from the bias ----
        float acc[PXT][8];
        {
          const uint32_t bda = smem_base + plan.bd_off + (uint32_t)(c * 64 + cc * 8) * 4u;
          const uint4 ba = ptx::ld_shared_v4(bda), bb = ptx::ld_shared_v4(bda + 16u);
#pragma unroll
          for (int px = 0; px < PXT; ++px) {
            acc[px][0] = __uint_as_float(ba.x); acc[px][1] = __uint_as_float(ba.y); acc[px][2] = __uint_as_float(ba.z); acc[px][3] = __uint_as_float(ba.w);
            acc[px][4] = __uint_as_float(bb.x); acc[px][5] = __uint_as_float(bb.y); acc[px][6] = __uint_as_float(bb.z); acc[px][7] = __uint_as_float(bb.w);
          }
        }
        const uint32_t wda = smem_base + plan.wd_off + (uint32_t)(c * 64 + cc * 8) * 2u;
#pragma unroll
        for (int dr = 0; dr < 3; ++dr) {
          float xin[NIN][8];
#pragma unroll
          for (int i = 0; i < NIN; ++i) {
            const uint32_t r = (uint32_t)((orow * STRIDE + dr) * IW + ocol0 * STRIDE + i);
            unpack8<T>(ptx::ld_shared_v4(hbuf + r * 128u + ((((uint32_t)cc) ^ (r & 7u)) << 4)), xin[i]);
          }
#pragma unroll
          for (int dc = 0; dc < 3; ++dc) {
            float w8[8];
            unpack8<T>(ptx::ld_shared_v4(wda + (uint32_t)((dr * 3 + dc) * p.Ch_p) * 2u), w8);
#pragma unroll
            for (int px = 0; px < PXT; ++px)
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[px][e] = fmaf(xin[px * STRIDE + dc][e], w8[e], acc[px][e]);
          }
        }
        if (!EXPAND) {                                            // the TMA-loaded chunk has been read
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&hempty[hb]);
        }
        // the project MMAs of the previous chunk have finished reading the A2 tile
        ptx::mbar_wait(a2empty, (gc & 1u) ^ 1u);
        const uint32_t a2 = smem_base + plan.a2_off;
#pragma unroll
        for (int px = 0; px < PXT; ++px) {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[px][e] = fminf(fmaxf(acc[px][e], p.lo2), p.hi2);
          const uint32_t r = (uint32_t)(orow * MB_TW + ocol0 + px);
          ptx::st_shared_v4(a2 + r * 128u + ((((uint32_t)cc) ^ (r & 7u)) << 4),
                            make_uint4(pack2<T>(acc[px][0], acc[px][1]), pack2<T>(acc[px][2], acc[px][3]), pack2<T>(acc[px][4], acc[px][5]),
                                       pack2<T>(acc[px][6], acc[px][7])));
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a2full);
      }
    }
  } else if (warp == MB_TMA_WARP) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      ptx::grid_dependency_wait();
      uint32_t gc = 0, it = 0;
      uint32_t ws = 0, wph = 0;                                  // weight ring stage / phase
      auto load_x = [&](int tile, uint32_t k) {                  // X tile of the k-th tile of this CTA
        const int n = tile / (args.tiles_h * args.tiles_w);
        const int rem = tile - n * args.tiles_h * args.tiles_w;
        const int th = rem / args.tiles_w;
        const int ih0 = th * TH * STRIDE - p.pt, iw0 = (rem - th * args.tiles_w) * MB_TW * STRIDE - p.pl;
        const uint32_t xb = two_x ? (k & 1u) : 0u, use = two_x ? (k >> 1) : k;
        ptx::mbar_wait(&xempty[xb], (use & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(&xfull[xb], (uint32_t)(kb1 * ROWS_IN * 128));
        for (int kb = 0; kb < kb1; ++kb)
          ptx::tma_load_4d(smem_base + plan.xa_off + xb * plan.xa_bytes + (uint32_t)kb * plan.xa_kb, &args.tm_x, &xfull[xb], kb * 64, iw0, ih0, n);
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int n = tile / (args.tiles_h * args.tiles_w);
        const int rem = tile - n * args.tiles_h * args.tiles_w;
        const int th = rem / args.tiles_w;
        const int ih0 = th * TH * STRIDE - p.pt, iw0 = (rem - th * args.tiles_w) * MB_TW * STRIDE - p.pl;
        if (EXPAND && (it == 0 || !two_x)) load_x(tile, it);
        for (int c = 0; c < nchunk; ++c, ++gc) {
          if (!w_resident || it == 0) {
            if (!w_resident) ptx::mbar_wait(&wempty[ws], wph ^ 1u);
            ptx::mbar_arrive_expect_tx(&wfull[ws], (EXPAND ? plan.w1_bytes : 0u) + plan.w2_bytes);
            if (EXPAND)
              for (int kb = 0; kb < kb1; ++kb)
                ptx::tma_load_2d(smem_base + plan.w1_off + ws * plan.w1_bytes + (uint32_t)kb * 8192u, &args.tm_w1, &wfull[ws], kb * 64, c * 64);
            ptx::tma_load_2d(smem_base + plan.w2_off + ws * plan.w2_bytes, &args.tm_w2, &wfull[ws], c * 64, 0);
            if (++ws == nw) { ws = 0; wph ^= 1u; }
          }
          if (!EXPAND) {
            const uint32_t hb = gc & (plan.nh - 1u);
            ptx::mbar_wait(&hempty[hb], ((gc >> (plan.nh == 4u ? 2 : 1)) & 1u) ^ 1u);
            ptx::mbar_arrive_expect_tx(&hfull[hb], (uint32_t)(ROWS_IN * 128));
            ptx::tma_load_4d(smem_base + plan.h_off + hb * plan.h_bytes, &args.tm_x, &hfull[hb], c * 64, iw0, ih0, n);
          }
          // The next tile's X tile goes into the other buffer while this tile's last chunks are computed.  Only AFTER the tile's last weight
          // chunk: the buffer is released by the epilogue of the tile before this one, which the E warps take after producing this tile's
          // first two chunks - a producer waiting for it earlier would hold back the very weights those chunks need (seen as a hang with
          // >= 3 tiles per CTA and a residual)
          if (EXPAND && two_x && c + 1 == nchunk && tile + (int)gridDim.x < num_tiles) load_x(tile + (int)gridDim.x, it + 1u);
        }
      }
    }
  } else if (warp == MB_MMA_WARP) {
    // ======================================= MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc_e = ptx::make_idesc_f16(TcType<T>::is_bf16, 128, 64);
      const uint32_t idesc_p = ptx::make_idesc_f16(TcType<T>::is_bf16, 128, p.Cout_alloc);
      uint32_t gc = 0, it = 0;
      uint32_t xoff = 0;                                         // X tile of the current tile
      uint32_t es = 0, eph = 0, ps = 0, pph = 0;                 // weight ring position of the expand / project streams (resident: phase stays 0)
      // expand MMAs of global chunk k (the X tile is the current tile's, the weights the chunk's)
      auto expand = [&](uint32_t k) {
        const uint32_t wb = es;
        ptx::mbar_wait(&wfull[wb], eph);
        if (++es == nw) { es = 0; if (!w_resident) eph ^= 1u; }
        ptx::mbar_wait(eempty, (k & 1u) ^ 1u);
        ptx::tc_fence_after();
#pragma unroll
        for (int m = 0; m < MT; ++m)
          for (int kb = 0; kb < kb1; ++kb) {
            const uint64_t da = ptx::make_smem_desc_sw128(smem_base + plan.xa_off + xoff + (uint32_t)kb * plan.xa_kb + (uint32_t)m * 16384u);
            const uint64_t db = ptx::make_smem_desc_sw128(smem_base + plan.w1_off + wb * plan.w1_bytes + (uint32_t)kb * 8192u);
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) ptx::umma_f16(acc_e + (uint32_t)(m * 64), da + 2u * k16, db + 2u * k16, idesc_e, (uint32_t)(kb | k16));
          }
        ptx::umma_commit(efull);
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t xb = two_x ? (it & 1u) : 0u, xuse = two_x ? (it >> 1) : it;
        if (EXPAND) {
          xoff = xb * plan.xa_bytes;
          ptx::mbar_wait(&xfull[xb], xuse & 1u);
          expand(gc);
        }
        for (int c = 0; c < nchunk; ++c, ++gc) {
          const uint32_t wb = ps;
          // the next chunk's expand MMAs go out BEFORE this chunk's project MMAs are waited for: the E warps find the next expand
          // accumulator ready as soon as they have drained this one
          if (EXPAND && c + 1 < nchunk) expand(gc + 1);
          if (!EXPAND) ptx::mbar_wait(&wfull[wb], pph);
          if (++ps == nw) { ps = 0; if (!w_resident) pph ^= 1u; }
          if (c == 0) ptx::mbar_wait(pempty, (it & 1u) ^ 1u);     // the previous tile's epilogue has drained the project accumulator
          ptx::mbar_wait(a2full, gc & 1u);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_smem_desc_sw128(smem_base + plan.a2_off);
          const uint64_t db = ptx::make_smem_desc_sw128(smem_base + plan.w2_off + wb * plan.w2_bytes);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) ptx::umma_f16(acc_p, da + 2u * k16, db + 2u * k16, idesc_p, (uint32_t)(c | k16));
          ptx::umma_commit(a2empty);
          if (!w_resident) ptx::umma_commit(&wempty[wb]);
          if (c + 1 == nchunk) {
            ptx::umma_commit(pfull);
            if (EXPAND && !p.has_res) ptx::umma_commit(&xempty[xb]);   // (with a residual the epilogue releases the X tile)
          }
        }
      }
    }
  }
  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == MB_MMA_WARP) ptx::tmem_dealloc(tmem_base, 512);
}

CUtensorMap mb_tmap_4d(const void* base, bool bf16, const uint64_t* dims, const uint64_t* strides, const uint32_t* box, bool swizzle) {
  auto enc = encode_tiled_fn();
  B200_CHECK(enc, B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap m;
  cuuint64_t d[4]; cuuint64_t st[3]; cuuint32_t bx[4]; cuuint32_t es[4];
  for (int i = 0; i < 4; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < 3; ++i) st[i] = strides[i];
  CUresult r = enc(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), d, st, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, B200_ERR_CUDA, "mbconv: cuTensorMapEncodeTiled failed (" << (int)r << ")");
  return m;
}

int g_mb_sms = 0;

template <typename T>
void launch_mb(const MbParams& p, bool bf16, cudaStream_t s) {
  MbArgs a;
  memset(&a, 0, sizeof(a));
  a.p = p;
  const MbGeom g = mb_geom(p.stride, p.Cin_p, p.Ch_p);
  const MbPlan plan = mb_plan(g, p.has_expand, p.Ch_p, p.Cout_p, p.Cout_alloc);
  a.tiles_w = (p.Wo + MB_TW - 1) / MB_TW;
  a.tiles_h = (p.Ho + g.TH - 1) / g.TH;
  a.num_tiles = p.N * a.tiles_h * a.tiles_w;
  {
    const uint64_t C = (uint64_t)p.Cin_p;
    const uint64_t xd[4] = {C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    const uint64_t xs[3] = {C * 2, (uint64_t)p.W * C * 2, (uint64_t)p.H * p.W * C * 2};
    const uint32_t xb[4] = {64, (uint32_t)g.IW, (uint32_t)g.IH, 1};
    a.tm_x = mb_tmap_4d(p.x, bf16, xd, xs, xb, true);
    const uint64_t Co = (uint64_t)p.Cout_p;
    const uint64_t yd[4] = {Co, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.N};
    const uint64_t ys[3] = {Co * 2, (uint64_t)p.Wo * Co * 2, (uint64_t)p.Ho * p.Wo * Co * 2};
    const uint32_t yb[4] = {(uint32_t)p.Cout_p, MB_TW, (uint32_t)g.TH, 1};
    a.tm_y = mb_tmap_4d(p.y, bf16, yd, ys, yb, false);
  }
  if (p.has_expand) a.tm_w1 = make_tmap_2d(p.w1, bf16, (uint64_t)g.kb1 * 64, (uint64_t)p.Ch_p, (uint64_t)g.kb1 * 128, 64, 64);
  a.tm_w2 = make_tmap_2d(p.w2, bf16, (uint64_t)p.Ch_p, (uint64_t)p.Cout_alloc, (uint64_t)p.Ch_p * 2, 64, (uint32_t)p.Cout_alloc);
  {
    static const bool dbg = [] { const char* e = getenv("B200_DEBUG_MBCONV"); return e && e[0] == '1'; }();
    if (dbg)
      fprintf(stderr, "b200 mbconv: %dx%dx%d -> hidden %d -> %d, stride %d, expand %d res %d: tiles %d, chunks %d, weight stages %u, X tiles %u, smem %u\n", p.H, p.W,
              p.Cin_p, p.Ch_p, p.Cout_p, p.stride, p.has_expand, p.has_res, a.num_tiles, g.nchunk, plan.nw, plan.nxb, plan.total);
  }
  const int grid = std::min(a.num_tiles, g_mb_sms > 0 ? g_mb_sms : 148);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(MB_THREADS, 1, 1);
  cfg.dynamicSmemBytes = plan.total;
  cfg.stream = s;
  static const bool pdl = [] { const char* e = getenv("B200_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  if (p.stride == 1) {
    if (p.has_expand) cudaLaunchKernelEx(&cfg, mbconv_kernel<T, 1, true>, a);
    else cudaLaunchKernelEx(&cfg, mbconv_kernel<T, 1, false>, a);
  } else {
    if (p.has_expand) cudaLaunchKernelEx(&cfg, mbconv_kernel<T, 2, true>, a);
    else cudaLaunchKernelEx(&cfg, mbconv_kernel<T, 2, false>, a);
  }
}

template <typename T>
void mb_opt_in() {
  cudaFuncSetAttribute(mbconv_kernel<T, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM_LIMIT);
  cudaFuncSetAttribute(mbconv_kernel<T, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM_LIMIT);
  cudaFuncSetAttribute(mbconv_kernel<T, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM_LIMIT);
  cudaFuncSetAttribute(mbconv_kernel<T, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM_LIMIT);
}

}  // namespace

bool mbconv_supported(const MbParams& p, int dtype) {
  const char* e = getenv("B200_FUSE_MBCONV");
  if (e && e[0] == '0') return false;
  if (dtype != B200_BF16 && dtype != B200_F16) return false;
  if (p.stride != 1 && p.stride != 2) return false;
  if (p.Ch_p % 64 || p.Ch_p < 64 || p.Cout_alloc % 16 || p.Cout_alloc < 16 || p.Cout_alloc > 256 || p.Cout_p % 8 || p.Cout_p > p.Cout_alloc) return false;
  if (p.Cin_p % 8 || p.pt < 0 || p.pt > 1 || p.pl < 0 || p.pl > 1) return false;
  if (p.has_expand) {
    if (p.Cin_p > 192) return false;                         // X tile: at most three 64-channel k-blocks
    if (p.has_res && (p.stride != 1 || p.Cin_p != p.Cout_p || p.pt != 1 || p.pl != 1)) return false;
  } else {
    if (p.has_res || p.Cin_p > p.Ch_p || p.Ch_p - p.Cin_p >= 64) return false;   // the hidden tensor IS x: its padded width covers the chunks
  }
  if ((int64_t)p.N * ((p.Ho + 3) / 4) * ((p.Wo + MB_TW - 1) / MB_TW) >= ((int64_t)1 << 31)) return false;
  const MbGeom g = mb_geom(p.stride, p.Cin_p, p.Ch_p);
  return mb_plan(g, p.has_expand, p.Ch_p, p.Cout_p, p.Cout_alloc).total <= (uint32_t)MB_SMEM_LIMIT;
}

void mbconv_init_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_mb_sms, cudaDevAttrMultiProcessorCount, dev);
  mb_opt_in<__nv_bfloat16>();
  mb_opt_in<__half>();
}

void launch_mbconv(const MbParams& p, int dtype, cudaStream_t s) {
  B200_CHECK(mbconv_supported(p, dtype), B200_ERR_UNSUPPORTED, "mbconv: unsupported launch");
  if (dtype == B200_BF16) launch_mb<__nv_bfloat16>(p, true, s);
  else launch_mb<__half>(p, false, s);
}

}  // namespace b200
