// Two back-to-back 1x1 convolutions of a ResNet bottleneck boundary in ONE CTA-pair tcgen05 kernel (sm_100a):
//
//   Y[M, Cc] = act1( X[M, K] * Wc[Cc, K]^T + bias_c (+ R[M, Cc]) )          "expand" conv of block i (+ residual, ReLU): stored
//   Z[M, Ca] = act3( Y[M, Cc] * Wa[Ca, Cc]^T + bias_a )                     "reduce" conv of block i+1: stored
//
// Separately, the second GEMM re-reads Y (100-400 MB at batch 256) from HBM right after the first one wrote it.  Here every 64-channel
// slab of Y that the epilogue has converted to bf16 in its 128-byte-swizzled staging buffer (for the TMA store) IS the K-major A operand
// of a K = 64 step of the second GEMM: a second MMA issuer multiplies it with the matching 64 columns of Wa into a third accumulator.
// Same arithmetic contract as conv_tc3 (ngraph reference/convolution.hpp:89-186; nodes/mkldnn_conv_node.cpp:355-403); Z is computed from
// the bf16-rounded Y, exactly what the unfused pair of kernels does.
//
// One cluster walks whole M tiles (256 rows per pair): for n in Cc / 128 N tiles: main GEMM -> Y accumulator (two of them, 128 columns
// each) -> per 64-channel slab: epilogue -> staging slab -> {TMA store of Y, 4 x tcgen05.mma into the Z accumulator (Ca <= 256 columns)};
// after the last N tile the Z accumulator goes through the same epilogue (bias_a, act3) and is stored.
//
//   warp 0-7  epilogue: two groups of four warps on alternate slabs (Y slabs, then the Z slabs of the M tile)
//   warp 8    leader: main MMA issuer.  peer: forwards "Y accumulator drained" / "Z accumulator drained" to the leader
//   warp 9    operand TMA (activation matrix, optional second source = folded projection shortcut, weights)
//   warp 10   residual TMA
//   warp 11   leader: second MMA issuer (Y slab x Wa chunk).  peer: forwards "slab ready" of epilogue group 0 (warp 13: of group 1)
//   warp 12   Wa chunk loads (ring of 2-8 stages)
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
tc::EpiAct make_epi(int kind, float a, float b);

namespace tc3f {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int A_BYTES = BM * 128, B_BYTES = (BN / 2) * 128, SLAB_BYTES = BM * 128;
constexpr int NUM_THREADS = 448;
constexpr int MAX_SA = 4, NRES = 3, NSTG = 4, MAX_S2 = 8;
constexpr int SMEM_LIMIT = 232448;
constexpr int Z_COL = 2 * BN;            // first TMEM column of the Z accumulator

struct Plan { int sa, s2; uint32_t a_off, b_off, stg_off, res_off, w2_off, bias_off, bar_off, total, w2_stage; };
__host__ __device__ inline Plan make_plan(bool has_res, int ca, int cc) {
  Plan pl;
  pl.w2_stage = (uint32_t)((ca / 2) * 128);
  // Wa chunk ring: 32 KB (2 stages at Ca = 256, 4 at 128, 8 at 64).  One load takes ~1000+ cycles (L2 hit), a slab arrives every ~700
  pl.s2 = (32 * 1024) / (int)pl.w2_stage;
  if (pl.s2 > MAX_S2) pl.s2 = MAX_S2;
  const int bias_bytes = ((cc + ca) * 4 + 1023) & ~1023;
  const int fixed = NSTG * SLAB_BYTES + (has_res ? NRES * SLAB_BYTES : 0) + pl.s2 * (int)pl.w2_stage + bias_bytes + 1024 + 1024;
  pl.sa = (SMEM_LIMIT - fixed) / (A_BYTES + B_BYTES);
  if (pl.sa > MAX_SA) pl.sa = MAX_SA;
  pl.a_off = 0;
  pl.b_off = pl.a_off + (uint32_t)pl.sa * A_BYTES;
  pl.stg_off = pl.b_off + (uint32_t)pl.sa * B_BYTES;
  pl.res_off = pl.stg_off + NSTG * SLAB_BYTES;
  pl.w2_off = pl.res_off + (has_res ? NRES * SLAB_BYTES : 0);
  pl.bias_off = pl.w2_off + (uint32_t)pl.s2 * pl.w2_stage;
  pl.bar_off = pl.bias_off + (uint32_t)bias_bytes;
  pl.total = pl.bar_off + 1024u + 1024u;
  return pl;
}

struct Args {
  CUtensorMap tm_a, tm_a2, tm_w, tm_y, tm_r, tm_w2, tm_z;
  ConvParams p;
  EpiAct e1, e2, e3;
  int m_tiles, n_tiles;     // pair M tiles, N tiles of 128 columns
  int nkb1;                 // k-blocks served by tm_a (the rest: second source tm_a2)
  int n2slab;               // Ca / 64
  int reverse;
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// 32 accumulator columns of one row: + bias, act, (+ residual, act), bf16 pack, swizzled staging store
template <typename T, bool GEN>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], uint32_t bias, uint32_t out_row, uint32_t res_row, uint32_t rsw, int chunk0,
                                               bool has_res, const EpiAct& e1, const EpiAct& e2) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint4 b0 = ptx::ld_shared_v4(bias + (uint32_t)g * 32u);
    const uint4 b1 = ptx::ld_shared_v4(bias + (uint32_t)g * 32u + 16u);
    float f[8];
    f[0] = __uint_as_float(v[g * 8 + 0]) + __uint_as_float(b0.x); f[1] = __uint_as_float(v[g * 8 + 1]) + __uint_as_float(b0.y);
    f[2] = __uint_as_float(v[g * 8 + 2]) + __uint_as_float(b0.z); f[3] = __uint_as_float(v[g * 8 + 3]) + __uint_as_float(b0.w);
    f[4] = __uint_as_float(v[g * 8 + 4]) + __uint_as_float(b1.x); f[5] = __uint_as_float(v[g * 8 + 5]) + __uint_as_float(b1.y);
    f[6] = __uint_as_float(v[g * 8 + 6]) + __uint_as_float(b1.z); f[7] = __uint_as_float(v[g * 8 + 7]) + __uint_as_float(b1.w);
    epi_act8<GEN>(f, e1);
    const uint32_t sw = (rsw ^ (uint32_t)(chunk0 + g)) << 4;
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

template <typename T, bool GEN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) conv_tc3f_kernel(const __grid_constant__ Args args) {
  const ConvParams& p = args.p;
  const bool has_res = p.res != nullptr;
  const int nkb = p.Ktot / BK;
  const int Ca = p.N2;
  const Plan plan = make_plan(has_res, Ca, p.Cout_alloc);
  const int SA = plan.sa, S2 = plan.s2;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem_al + plan.bar_off);
  uint64_t* emptyA = fullA + MAX_SA;
  uint64_t* fullB = emptyA + MAX_SA;
  uint64_t* emptyB = fullB + MAX_SA;
  uint64_t* tfull = emptyB + MAX_SA;
  uint64_t* tempty = tfull + 2;
  uint64_t* rfull = tempty + 2;
  uint64_t* rempty = rfull + NRES;
  uint64_t* yready = rempty + NRES;
  uint64_t* yfree = yready + NSTG;
  uint64_t* w2full = yfree + NSTG;
  uint64_t* w2empty = w2full + MAX_S2;
  uint64_t* zfull = w2empty + MAX_S2;
  uint64_t* zempty = zfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(zempty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int m_tiles = args.m_tiles, n_tiles = args.n_tiles;
  const int cluster = (int)ptx::cluster_id_x(), nclusters = (int)ptx::num_clusters_x();
  const int my_m = cluster < m_tiles ? (m_tiles - cluster + nclusters - 1) / nclusters : 0;   // M tiles this cluster walks
  const int my_tiles = my_m * n_tiles;
  const int n2slab = args.n2slab;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_SA; ++s) {
      ptx::mbar_init(&fullA[s], 1); ptx::mbar_init(&emptyA[s], 1);
      ptx::mbar_init(&fullB[s], 1); ptx::mbar_init(&emptyB[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], leader ? 9 : 8);      // eight local epilogue warps (+ the peer's forwarder in the leader)
    }
    for (int r = 0; r < NRES; ++r) { ptx::mbar_init(&rfull[r], 1); ptx::mbar_init(&rempty[r], 4); }
    for (int s = 0; s < NSTG; ++s) {
      ptx::mbar_init(&yready[s], leader ? 2 : 1);      // this CTA's epilogue group (+ the peer's, forwarded)
      ptx::mbar_init(&yfree[s], 1);                    // multicast commit of the second issuer
    }
    for (int s = 0; s < MAX_S2; ++s) { ptx::mbar_init(&w2full[s], 1); ptx::mbar_init(&w2empty[s], 1); }
    ptx::mbar_init(zfull, 1);
    ptx::mbar_init(zempty, leader ? 9 : 8);
    ptx::fence_mbar_init();
  }
  if (warp == 9 && lane == 0) {
    ptx::tma_prefetch_desc(&args.tm_w); ptx::tma_prefetch_desc(&args.tm_a); ptx::tma_prefetch_desc(&args.tm_y);
    ptx::tma_prefetch_desc(&args.tm_w2); ptx::tma_prefetch_desc(&args.tm_z);
    if (has_res) ptx::tma_prefetch_desc(&args.tm_r);
  }
  if (warp == 8) {
    ptx::tmem_alloc_pair(tmem_slot, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();

  // tile `it` of this cluster -> (pair M tile, N tile)
  auto m_of = [&](int it) { const int mt = cluster + (it / n_tiles) * nclusters; return args.reverse ? m_tiles - 1 - mt : mt; };

  constexpr uint64_t DESC_HI = ((uint64_t)2 << 61) | ((uint64_t)1 << 46) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 16);

  if (warp < 8) {
    // ===================== epilogue =====================
    const int grp = warp >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t rsw = (uint32_t)(row & 7);
    const bool store_thread = (threadIdx.x & 127) == 0;
    const bool signal_thread = (threadIdx.x & 127) == 32;
    {
      float* s_bias = reinterpret_cast<float*>(smem_al + plan.bias_off);
      for (int i = threadIdx.x; i < p.Cout_alloc; i += 256) s_bias[i] = __ldg(p.bias + i);
      for (int i = threadIdx.x; i < Ca; i += 256) s_bias[p.Cout_alloc + i] = __ldg(p.bias2 + i);
      ptx::named_bar_sync(3, 256);
    }
    const uint32_t bias_base = smem_base + plan.bias_off;
    uint32_t g = 0, gy = 0, mine = 0, zph = 0, prev_ls = 2u;
    uint32_t yuse[2] = {0u, 0u};
    for (int it = 0; it < my_tiles; ++it) {
      const int n = it % n_tiles;
      const int m0 = (m_of(it) * 2 + (int)rank) * BM;
      const bool valid = m0 < p.M;
      const uint32_t acc = (uint32_t)it & 1u;
      ptx::mbar_wait(&tfull[acc], ((uint32_t)it >> 1) & 1u);
      ptx::tc_fence_after();
      for (int h = 0; h < BN / 64; ++h, ++g, ++gy) {
        if ((int)(g & 1u) != grp) continue;
        const uint32_t ls = mine & 1u, sidx = (uint32_t)grp * 2u + ls;   // Y slabs alternate this group's two staging slabs
        ++mine;
        const bool same_slot = ls == prev_ls;
        prev_ls = ls;
        const uint32_t rslot = gy % NRES;
        const uint32_t out_slab = smem_base + plan.stg_off + sidx * SLAB_BYTES;
        const uint32_t out_row = out_slab + (uint32_t)row * 128u;
        const uint32_t res_row = smem_base + plan.res_off + rslot * SLAB_BYTES + (uint32_t)row * 128u;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)(h * 64);
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32(taddr, v0);
        ptx::tmem_ld_32x32(taddr + 32, v1);
        if (has_res) ptx::mbar_wait(&rfull[rslot], (gy / NRES) & 1u);
        if (store_thread) {                                         // the TMA store that last used this staging slab has read it
          if (same_slot) ptx::bulk_wait_read<0>(); else ptx::bulk_wait_read<1>();
        }
        if (yuse[ls] > 0) ptx::mbar_wait(&yfree[sidx], (yuse[ls] - 1u) & 1u);   // ... and so has the second GEMM
        ptx::named_bar_sync(1 + grp, 128);
        ptx::tmem_ld_wait();
        const uint32_t bias = bias_base + (uint32_t)(n * BN + h * 64) * 4u;
        epilogue_chunk<T, GEN>(v0, bias, out_row, res_row, rsw, 0, has_res, args.e1, args.e2);
        epilogue_chunk<T, GEN>(v1, bias + 128u, out_row, res_row, rsw, 4, has_res, args.e1, args.e2);
        if (has_res) {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&rempty[rslot]);
        }
        ptx::fence_proxy_async();                                   // staging writes -> visible to the TMA engine and to tcgen05.mma
        ptx::named_bar_sync(1 + grp, 128);
        if (store_thread) {
          if (valid) ptx::tma_store_2d(&args.tm_y, out_slab, n * BN + h * 64, m0);
          ptx::bulk_commit();
        }
        if (signal_thread) ptx::mbar_arrive(&yready[sidx]);         // slab is an A operand now (rows beyond M: whatever the tail tile computed)
        ++yuse[ls];
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      // ---- Z accumulator of the PREVIOUS M tile, one tile late: by now its last Y slab has long been multiplied in (signal -> peer
      // forward -> MMA -> commit takes ~2000 cycles; waiting for it right after the last Y slab idled both groups: profiles/r1 call v) ----
      const bool last_tile = it == my_tiles - 1;
      if ((n == 0 && it > 0) || last_tile) {
        for (int rep = 0; rep < 2; ++rep) {
          // rep 0: the previous M tile (if this tile opened a new one); rep 1: the final M tile, after its own last tile
          if (rep == 0 && !(n == 0 && it > 0)) continue;
          if (rep == 1 && !(last_tile && n == n_tiles - 1)) continue;
          const int mz = rep == 0 ? m_of(it - 1) : m_of(it);
          const int mz0 = (mz * 2 + (int)rank) * BM;
          const bool zvalid = mz0 < p.M;
          ptx::mbar_wait(zfull, zph & 1u);
          ptx::tc_fence_after();
          for (int j = 0; j < n2slab; ++j, ++g) {
            if ((int)(g & 1u) != grp) continue;
            // Z slabs use the slab the NEXT Y slab of this group will use, never the one that holds the newest Y slab: that one is released
            // by an MMA that waits for this very epilogue to drain the Z accumulator (two Z slabs per group on alternating slabs deadlocked)
            const uint32_t ls = mine & 1u, sidx = (uint32_t)grp * 2u + ls;
            const bool same_slot = ls == prev_ls;
            prev_ls = ls;
            const uint32_t out_slab = smem_base + plan.stg_off + sidx * SLAB_BYTES;
            const uint32_t out_row = out_slab + (uint32_t)row * 128u;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(Z_COL + j * 64);
            uint32_t v0[32], v1[32];
            ptx::tmem_ld_32x32(taddr, v0);
            ptx::tmem_ld_32x32(taddr + 32, v1);
            if (store_thread) {
              if (same_slot) ptx::bulk_wait_read<0>(); else ptx::bulk_wait_read<1>();
            }
            if (yuse[ls] > 0) ptx::mbar_wait(&yfree[sidx], (yuse[ls] - 1u) & 1u);
            ptx::named_bar_sync(1 + grp, 128);
            ptx::tmem_ld_wait();
            const uint32_t bias = bias_base + (uint32_t)(p.Cout_alloc + j * 64) * 4u;
            epilogue_chunk<T, GEN>(v0, bias, out_row, 0u, rsw, 0, false, args.e3, args.e3);
            epilogue_chunk<T, GEN>(v1, bias + 128u, out_row, 0u, rsw, 4, false, args.e3, args.e3);
            ptx::fence_proxy_async();
            ptx::named_bar_sync(1 + grp, 128);
            if (store_thread) {
              if (zvalid) ptx::tma_store_2d(&args.tm_z, out_slab, j * 64, mz0);
              ptx::bulk_commit();
            }
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(zempty);
          ++zph;
        }
      }
    }
    if (store_thread) ptx::bulk_wait_all0();
  } else if (warp == 8) {
    if (!leader) {
      // ===================== peer: forward accumulator-drained signals =====================
      const uint32_t te0 = ptx::mapa(ptx::smem_u32(&tempty[0]), 0), te1 = ptx::mapa(ptx::smem_u32(&tempty[1]), 0);
      const uint32_t ze = ptx::mapa(ptx::smem_u32(zempty), 0);
      uint32_t zph = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t acc = (uint32_t)it & 1u;
        ptx::mbar_wait(&tempty[acc], ((uint32_t)it >> 1) & 1u);
        if (lane == 0) ptx::mbar_arrive_cluster(acc ? te1 : te0);
        __syncwarp();
        const int n = it % n_tiles;
        const int reps = ((n == 0 && it > 0) ? 1 : 0) + ((it == my_tiles - 1 && n == n_tiles - 1) ? 1 : 0);
        for (int r = 0; r < reps; ++r) {
          ptx::mbar_wait(zempty, zph & 1u);
          if (lane == 0) ptx::mbar_arrive_cluster(ze);
          __syncwarp();
          ++zph;
        }
      }
    } else {
      // ===================== leader: main MMA issuer =====================
      constexpr uint32_t idesc = ptx::make_idesc_f16(TcType<T>::is_bf16, 2 * BM, BN);
      const uint32_t a_lo0 = ((smem_base + plan.a_off) & 0x3FFFFu) >> 4, b_lo0 = ((smem_base + plan.b_off) & 0x3FFFFu) >> 4;
      int sa = 0;
      uint32_t pha = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t acc = (uint32_t)it & 1u;
        ptx::mbar_wait(&tempty[acc], (((uint32_t)it >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accum = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&fullA[sa], pha);
          ptx::mbar_wait(&fullB[sa], pha);
          ptx::tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + (uint32_t)sa * (A_BYTES >> 4), b_lo = b_lo0 + (uint32_t)sa * (B_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              ptx::umma_f16_pair(d_tmem, DESC_HI | (uint64_t)(a_lo + 2u * k), DESC_HI | (uint64_t)(b_lo + 2u * k), idesc, accum | (uint32_t)k);
            ptx::umma_commit_pair(&emptyA[sa], 3);      // one commit frees both operand stages (same index)
            ptx::umma_commit_pair(&emptyB[sa], 3);
          }
          __syncwarp();
          accum = 1;
          if (++sa == SA) { sa = 0; pha ^= 1u; }
        }
        if (elect_one()) ptx::umma_commit_pair(&tfull[acc], 3);
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ===================== operand TMA issuer =====================
    const uint32_t fullA_leader = ptx::smem_u32(&fullA[0]) & ptx::PEER_BIT_MASK;
    const uint32_t fullB_leader = ptx::smem_u32(&fullB[0]) & ptx::PEER_BIT_MASK;
    const int HoWo = p.Ho * p.Wo;
    int sa = 0;
    uint32_t pha = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int n = it % n_tiles;
      int m0 = (m_of(it) * 2 + (int)rank) * BM;
      if (m0 >= p.M) m0 -= BM;                                     // tail: stage the leader's rows again (never stored)
      int bn2 = 0, bh2 = 0, bw2 = 0;
      if (args.nkb1 < nkb) {
        bn2 = m0 / HoWo;
        const int rem = m0 - bn2 * HoWo;
        const int ho = rem / p.Wo;
        bh2 = ho * p.sh2;
        bw2 = (rem - ho * p.Wo) * p.sw2;
      }
      const int n0 = n * BN + (int)rank * (BN / 2);
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(&emptyA[sa], pha ^ 1u);
        ptx::mbar_wait(&emptyB[sa], pha ^ 1u);
        if (elect_one()) {
          const uint32_t a_dst = smem_base + plan.a_off + (uint32_t)sa * A_BYTES, b_dst = smem_base + plan.b_off + (uint32_t)sa * B_BYTES;
          const uint32_t bara = fullA_leader + (uint32_t)sa * 8u, barb = fullB_leader + (uint32_t)sa * 8u;
          if (leader) { ptx::mbar_arrive_expect_tx(&fullA[sa], 2u * A_BYTES); ptx::mbar_arrive_expect_tx(&fullB[sa], 2u * B_BYTES); }
          if (kb < args.nkb1) ptx::tma_load_2d_pair_hint(a_dst, &args.tm_a, bara, kb * BK, m0, ptx::L2_EVICT_FIRST);
          else ptx::tma_load_im2col_4d_pair(a_dst, &args.tm_a2, bara, (kb - args.nkb1) * BK, bw2, bh2, bn2, 0u, 0u);
          ptx::tma_load_2d_pair(b_dst, &args.tm_w, barb, kb * BK, n0);
        }
        __syncwarp();
        if (++sa == SA) { sa = 0; pha ^= 1u; }
      }
    }
  } else if (warp == 10) {
    // ===================== residual TMA issuer =====================
    if (has_res && lane == 0) {
      uint32_t gy = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int n = it % n_tiles;
        const int m0 = (m_of(it) * 2 + (int)rank) * BM;
        const bool valid = m0 < p.M;
        for (int h = 0; h < BN / 64; ++h, ++gy) {
          const uint32_t rslot = gy % NRES;
          ptx::mbar_wait(&rempty[rslot], ((gy / NRES) & 1u) ^ 1u);
          if (valid) {
            ptx::mbar_arrive_expect_tx(&rfull[rslot], SLAB_BYTES);
            ptx::tma_load_2d_hint(smem_base + plan.res_off + rslot * SLAB_BYTES, &args.tm_r, &rfull[rslot], n * BN + h * 64, m0, ptx::L2_EVICT_FIRST);
          } else {
            ptx::mbar_arrive(&rfull[rslot]);
          }
        }
      }
    }
  } else if (warp == 11 || warp == 13) {
    // ===================== second GEMM: Z += Yslab * Wa[:, slab]^T =====================
    const uint32_t yr_leader0 = ptx::mapa(ptx::smem_u32(&yready[0]), 0);
    const uint32_t idesc2 = ptx::make_idesc_f16(TcType<T>::is_bf16, 2 * BM, Ca);
    const uint32_t stg_lo0 = ((smem_base + plan.stg_off) & 0x3FFFFu) >> 4, w2_lo0 = ((smem_base + plan.w2_off) & 0x3FFFFu) >> 4;
    const int per_m = n_tiles * (BN / 64);                         // Y slabs (= 64-column chunks of Wa) per M tile
    // leader: warp 11 issues the second GEMM for every slab.  peer: one forwarder per epilogue group (warp 11: group 0, warp 13: group 1):
    // a remote arrive stalls its issuer for ~1200 cycles and a CTA finishes a slab every ~1000, so one forwarder could not keep up
    const uint32_t fwd_grp = warp == 11 ? 0u : 1u;
    uint32_t g = 0, zph = 0;
    uint32_t mineg[2] = {0u, 0u};
    uint32_t yuse[NSTG] = {0u, 0u, 0u, 0u};
    int i2 = 0;
    for (int it = 0; (leader ? warp == 11 : true) && it < my_tiles; ++it) {
      const int n = it % n_tiles;
      for (int h = 0; h < BN / 64; ++h, ++g, ++i2) {
        const int c = n * (BN / 64) + h;                            // 64-column chunk of Wa = Y slab index inside the M tile
        const uint32_t grp = g & 1u, sidx = grp * 2u + (mineg[grp] & 1u);
        ++mineg[grp];
        const int s2 = i2 % S2;
        const uint32_t ph2 = (uint32_t)(i2 / S2) & 1u;
        if (leader) {
          if (c == 0) { ptx::mbar_wait(zempty, (zph & 1u) ^ 1u); ++zph; }   // the epilogue has drained the previous M tile's Z accumulator
          ptx::mbar_wait(&w2full[s2], ph2);
          ptx::mbar_wait(&yready[sidx], yuse[sidx] & 1u);
          ptx::tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = stg_lo0 + sidx * (SLAB_BYTES >> 4), b_lo = w2_lo0 + (uint32_t)s2 * (plan.w2_stage >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16_pair(tmem_base + Z_COL, DESC_HI | (uint64_t)(a_lo + 2u * k), DESC_HI | (uint64_t)(b_lo + 2u * k), idesc2,
                                 (uint32_t)((c | k) != 0));
            ptx::umma_commit_pair(&w2empty[s2], 3);
            ptx::umma_commit_pair(&yfree[sidx], 3);
            if (c == per_m - 1) ptx::umma_commit_pair(zfull, 3);
          }
          __syncwarp();
        } else if (grp == fwd_grp) {
          ptx::mbar_wait(&yready[sidx], yuse[sidx] & 1u);           // my epilogue group has finished the slab: tell the leader
          if (lane == 0) ptx::mbar_arrive_cluster(yr_leader0 + sidx * 8u);
          __syncwarp();
        }
        ++yuse[sidx];
      }
      // the Z slabs of an M tile take their turn in the group alternation (one tile late / after the last tile) but do not move the
      // groups' Y slab alternation
      const int reps = ((n == 0 && it > 0) ? 1 : 0) + ((it == my_tiles - 1 && n == n_tiles - 1) ? 1 : 0);
      g += (uint32_t)(reps * n2slab);
    }
  } else if (warp == 12) {
    // ===================== Wa chunk loader: 64 columns of Wa per Y slab, this CTA's half of the rows, ring of S2 =====================
    const uint32_t w2full_leader = ptx::smem_u32(&w2full[0]) & ptx::PEER_BIT_MASK;
    const int per_m = n_tiles * (BN / 64);
    const int total_y = my_m * per_m;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < total_y; ++i) {
      ptx::mbar_wait(&w2empty[s], ph ^ 1u);
      if (elect_one()) {
        if (leader) ptx::mbar_arrive_expect_tx(&w2full[s], 2u * plan.w2_stage);
        ptx::tma_load_2d_pair(smem_base + plan.w2_off + (uint32_t)s * plan.w2_stage, &args.tm_w2, w2full_leader + (uint32_t)s * 8u, (i % per_m) * 64,
                              (int)rank * (Ca / 2));
      }
      __syncwarp();
      if (++s == S2) { s = 0; ph ^= 1u; }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 8) ptx::tmem_dealloc_pair(tmem_base, 512);
}

template <typename K>
inline void launch_pdl(K kernel, int grid, unsigned smem, cudaStream_t s, const Args& a) {
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

}  // namespace tc3f

namespace {
int g_pairs3f = 0;
}

bool conv_tc3f_supported(const ConvParams& p) {
  using namespace tc3f;
  if (p.N2 <= 0) return false;
  const bool geom = p.kh == 1 && p.kw == 1 && p.sh == 1 && p.sw == 1 && p.pt == 0 && p.pl == 0 && p.x_sw == 0 && p.Cin_p % 64 == 0 && p.K2 % 64 == 0 &&
                    p.Ktot % BK == 0 && p.Cout_p % BN == 0 && p.Cout_alloc == p.Cout_p && p.ldy == p.Cout_p && (p.N2 == 64 || p.N2 == 128 || p.N2 == 256) &&
                    p.ldz == p.N2 && p.batch == 1 && !p.out_f32 && p.w_ld == p.Ktot;
  return geom && make_plan(true, p.N2, p.Cout_alloc).sa >= 2;
}

void conv_tc3f_init_device() {
  using namespace tc3f;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  g_pairs3f = sms / 2;
  auto opt_in = [](auto k) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT); };
  opt_in(conv_tc3f_kernel<__nv_bfloat16, false>); opt_in(conv_tc3f_kernel<__nv_bfloat16, true>);
  opt_in(conv_tc3f_kernel<__half, false>); opt_in(conv_tc3f_kernel<__half, true>);
}

void launch_conv_tc3f(const ConvParams& p, int dtype, cudaStream_t s) {
  using namespace tc3f;
  B200_CHECK(conv_tc3f_supported(p), B200_ERR_UNSUPPORTED, "conv_tc3f: unsupported launch");
  const bool bf16 = dtype == B200_BF16;
  B200_CHECK(bf16 || dtype == B200_F16, B200_ERR_UNSUPPORTED, "conv_tc3f: dtype must be bf16/f16");
  Args a;
  memset(&a, 0, sizeof(a));
  a.p = p;
  a.e1 = make_epi(p.act1, p.a1, p.b1);
  a.e2 = make_epi(p.act2, p.a2, p.b2);
  a.e3 = make_epi(p.act3, p.a3, p.b3);
  if (p.res != nullptr && p.act2 == ACT_RELU) a.e2.relu_in_pack = 1;
  if (p.res == nullptr && p.act1 == ACT_RELU) a.e1.relu_in_pack = 1;
  if (p.act3 == ACT_RELU) a.e3.relu_in_pack = 1;
  a.reverse = p.reverse;
  a.n_tiles = p.Cout_alloc / BN;
  a.m_tiles = ((p.M + BM - 1) / BM + 1) / 2;
  a.n2slab = p.N2 / 64;
  a.nkb1 = (p.K2 > 0 ? p.Cin_p : p.Ktot) / BK;
  a.tm_w = make_tmap_2d(p.w, bf16, (uint64_t)p.Ktot, (uint64_t)p.Cout_alloc, (uint64_t)p.w_ld * 2, BK, BN / 2);
  a.tm_a = make_tmap_2d(p.x, bf16, (uint64_t)p.Cin_p, (uint64_t)p.M, (uint64_t)p.Cin_p * 2, BK, BM);
  if (p.K2 > 0) {
    ConvParams p2;
    p2.x = p.x2; p2.N = p.N; p2.H = p.H2; p2.W = p.W2; p2.Cin_p = p.K2; p2.Ho = p.Ho; p2.Wo = p.Wo;
    p2.kh = p2.kw = 1; p2.sh = p.sh2; p2.sw = p.sw2; p2.dh = p2.dw = 1; p2.pt = p2.pl = 0;
    a.tm_a2 = make_tmap_im2col(p2, bf16);
  } else {
    a.tm_a2 = a.tm_a;
  }
  a.tm_y = make_tmap_2d(p.y, bf16, (uint64_t)p.Cout_p, (uint64_t)p.M, (uint64_t)p.ldy * 2, 64, BM);
  a.tm_r = p.res ? make_tmap_2d(p.res, bf16, (uint64_t)p.Cout_p, (uint64_t)p.M, (uint64_t)p.ldy * 2, 64, BM) : a.tm_y;
  a.tm_w2 = make_tmap_2d(p.w2, bf16, (uint64_t)p.Cout_p, (uint64_t)p.N2, (uint64_t)p.Cout_p * 2, 64, (uint32_t)(p.N2 / 2));
  a.tm_z = make_tmap_2d(p.z, bf16, (uint64_t)p.N2, (uint64_t)p.M, (uint64_t)p.ldz * 2, 64, BM);
  const int grid = 2 * std::min(a.m_tiles, g_pairs3f > 0 ? g_pairs3f : 74);
  const unsigned smem = make_plan(p.res != nullptr, p.N2, p.Cout_alloc).total;
  if (conv_acts_are_clamps(p) && (p.act3 == ACT_NONE || p.act3 == ACT_RELU || p.act3 == ACT_CLAMP))
    bf16 ? launch_pdl(conv_tc3f_kernel<__nv_bfloat16, false>, grid, smem, s, a) : launch_pdl(conv_tc3f_kernel<__half, false>, grid, smem, s, a);
  else
    bf16 ? launch_pdl(conv_tc3f_kernel<__nv_bfloat16, true>, grid, smem, s, a) : launch_pdl(conv_tc3f_kernel<__half, true>, grid, smem, s, a);
}

}  // namespace b200
