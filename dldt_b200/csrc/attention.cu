// Fused scaled-dot-product attention for the BERT path (sm_100a, tcgen05 + TMEM + TMA): softmax(Q K^T * scale + mask) V, one persistent CTA
// per SM walking (image, head) items, 128-query tiles, the scores never leaving the SM.
//
// What it replaces in the schedule: MatMul (Q K^T) -> Multiply -> Add(mask) -> Softmax -> MatMul (P V) -> Transpose, i.e. the reference's
// nodes/mkldnn_matmul_node.cpp:225-262 (gemm looped over batch x heads), mkldnn_softmax_node.cpp and the eltwise nodes between them; as
// separate kernels the [N, heads, S, S] scores made three HBM round trips (Q K^T store, softmax read + write, P V load): 2.7 ms of a
// 6.9 ms BERT-base step at batch 64, S = 384.
//
// One CTA, 32 x (1 + S/32) threads:
//   warp 0 .. S/32-1   softmax: warp w owns TMEM lanes 32 (w % 4) .. +31 (the hardware's lane quarter rule) = 32 query rows, and the 128
//                      key columns [128 (w / 4), +128): one tcgen05.ld pass (the whole row segment stays in registers: 128 fp32), scale +
//                      mask in the log2 domain, row max / row sum exchanged between the S/128 column groups through shared memory,
//                      ex2.approx, bf16 pack, swizzled store of P as the K-major A operand of the second MMA.
//                      Warps 0-3 then read O (64 fp32 columns), divide by the row sum, pack and store their 128-byte rows.
//   last warp          one elected lane: TMA loads (Q 16 KB, K and V S x 128 B each), MMA 1: S[128 x S] = Q K^T as S/128 MMAs of
//                      128 x 128 x 64 into TMEM columns [0, S), MMA 2: O[128 x 64] = P V over K = S with V read MN-major
//                      (no transpose: V is [key][64], the descriptor's "B is N-contiguous" bit) into TMEM columns [S, S + 64).
// Shared memory: Q 16 KB | K S x 128 B | V S x 128 B | P 128 x S bf16 | mask, row max / sum, barriers: 215 KB at S = 384.
// TMEM: S + 64 <= 512 columns.  First version (one tile per CTA, nothing overlapped): 167 us per BERT-base layer at batch 64 - every tile
// paid CTA launch + TMEM allocation + TMA latency + the whole MMA1 -> softmax -> MMA2 -> store chain serially (10 us per tile).
// Normalisation is applied to O in fp32 (O / sum) rather than to P before its bf16 rounding; scores stay fp32 until P is formed, where
// the unfused path rounded them to bf16 twice (Q K^T output and softmax output).
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

CUtensorMap make_tmap_2d(const void* base, bool bf16, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer);

namespace {

using namespace tc;

constexpr int QT = 128;             // query rows per CTA
constexpr int DH = 64;              // head size
constexpr float LOG2E = 1.4426950408889634f;

struct AttnArgs {
  CUtensorMap tm_q, tm_k, tm_v;     // [N*H*S][64], boxes {64, 128}
  void* y;                          // [N*S][H*64], written with plain 16-byte stores (one 128-byte row per thread)
  const void* mask;                 // [mask_batch][S] in the compute type, or null
  int N, H, S, mask_batch;
  float scale_log2;                 // scale * log2(e)
};

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int S>
struct AttnSmem {
  static constexpr uint32_t q_off = 0;                               // 16 KB
  static constexpr uint32_t k_off = QT * 128;
  static constexpr uint32_t v_off = k_off + S * 128;
  static constexpr uint32_t p_off = v_off + S * 128;                 // S/64 K-blocks of 128 rows x 128 B
  static constexpr uint32_t mask_off = p_off + QT * S * 2;           // S floats
  static constexpr uint32_t red_off = mask_off + S * 4;              // row max [2][S/128][128], row sum [2][S/128][128] (double buffered by tile parity)
  static constexpr uint32_t bar_off = red_off + 4 * (S / 128) * QT * 4;
  static constexpr uint32_t total = bar_off + 128 + 1024;            // + alignment slack
};

// Persistent: CTA b walks the (image, head) items b, b + grid, ...; K and V of an item are loaded once and serve its S/128 query tiles.
// The single issuer thread runs one tile ahead on the first MMA: MMA1(t+1) is issued as soon as the softmax warps have pulled S(t) out of
// TMEM into registers, so it overlaps their exp / pack work; MMA2(t) follows when P(t) is in shared memory.  Q(t+1) is prefetched into
// the one Q buffer as soon as MMA1(t) has read it.
template <typename T, int S>
__global__ void __launch_bounds__(32 * (S / 32 + 1), 1) attention_kernel(const __grid_constant__ AttnArgs args) {
  constexpr int NG = S / 128;                 // column groups (softmax warps / 4) = query tiles per item
  constexpr int NSW = 4 * NG;                 // softmax warps
  constexpr uint32_t TMEM_COLS = (S + DH) <= 256 ? 256u : 512u;
  using L = AttnSmem<S>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bar_off);
  uint64_t* bar_qk = bars + 0;                // Q (+ K at the first tile of an item) landed
  uint64_t* bar_v = bars + 1;                 // V landed
  uint64_t* bar_s = bars + 2;                 // MMA 1 complete: S in TMEM
  uint64_t* bar_q1 = bars + 3;                // MMA 1 complete: the Q buffer (and K, at the item's last tile) may be overwritten
  uint64_t* bar_sfree = bars + 4;             // S pulled into registers by all softmax warps
  uint64_t* bar_p = bars + 5;                 // P written by all softmax warps
  uint64_t* bar_o = bars + 6;                 // MMA 2 complete: O in TMEM, P and (at the item's last tile) V free
  uint64_t* bar_ofree = bars + 7;             // O pulled into registers by the four output warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_mask = reinterpret_cast<float*>(smem + L::mask_off);
  float* s_max = reinterpret_cast<float*>(smem + L::red_off);
  float* s_sum = s_max + 2 * NG * QT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = args.N * args.H;
  const int my_items = ((int)blockIdx.x < items) ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int G = my_items * NG;                // tiles of this CTA

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_qk, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_q1, 1);
    ptx::mbar_init(bar_sfree, NSW);
    ptx::mbar_init(bar_p, NSW);
    ptx::mbar_init(bar_o, 1);
    ptx::mbar_init(bar_ofree, NG >= 2 ? 8 : 4);
    ptx::fence_mbar_init();
  }
  if (warp == NSW) {
    if (lane == 0) {
      ptx::tma_prefetch_desc(&args.tm_q);
      ptx::tma_prefetch_desc(&args.tm_k);
      ptx::tma_prefetch_desc(&args.tm_v);
    }
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::grid_dependency_wait();                // programmatic dependent launch: the Q / K / V producers have completed
  ptx::grid_launch_dependents();

  if (warp == NSW) {
    // ===================== TMA + MMA issuer (one elected lane) =====================
    if (elect_one_sync() && G > 0) {
      constexpr uint32_t idesc1 = ptx::make_idesc_f16(TcType<T>::is_bf16, QT, 128);
      constexpr uint32_t idesc2 = ptx::make_idesc_f16(TcType<T>::is_bf16, QT, DH) | (1u << 16);
      auto item_of = [&](int g) { return (int)blockIdx.x + (g / NG) * (int)gridDim.x; };
      auto load_qk = [&](int g) {             // Q of tile g, plus K when g opens an item
        const int nh = item_of(g), qt = g % NG;
        const bool first = qt == 0;
        ptx::mbar_arrive_expect_tx(bar_qk, (uint32_t)(QT * 128 + (first ? S * 128 : 0)));
        ptx::tma_load_2d(smem_base + L::q_off, &args.tm_q, bar_qk, 0, nh * S + qt * QT);
        if (first)
          for (int b = 0; b < NG; ++b) ptx::tma_load_2d(smem_base + L::k_off + (uint32_t)b * (128 * 128), &args.tm_k, bar_qk, 0, nh * S + b * 128);
      };
      auto load_v = [&](int g) {
        const int nh = item_of(g);
        ptx::mbar_arrive_expect_tx(bar_v, (uint32_t)(S * 128));
        for (int b = 0; b < NG; ++b) ptx::tma_load_2d(smem_base + L::v_off + (uint32_t)b * (128 * 128), &args.tm_v, bar_v, 0, nh * S + b * 128);
      };
      auto mma1 = [&]() {                     // S = Q K^T: K-major A (Q) and B (K), 128 key rows per MMA
#pragma unroll
        for (int b = 0; b < NG; ++b) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(smem_base + L::q_off + (uint32_t)k * 32u);
            const uint64_t db = ptx::make_smem_desc_sw128(smem_base + L::k_off + (uint32_t)b * (128 * 128) + (uint32_t)k * 32u);
            ptx::umma_f16(tmem_base + (uint32_t)(b * 128), da, db, idesc1, k ? 1u : 0u);
          }
        }
        ptx::umma_commit(bar_s);
        ptx::umma_commit(bar_q1);
      };
      load_qk(0);
      load_v(0);
      ptx::mbar_wait(bar_qk, 0);
      ptx::tc_fence_after();
      mma1();
      if (G > 1) {                            // Q(1) goes out as soon as MMA1(0) has read Q(0): a TMA round trip ahead of its use
        ptx::mbar_wait(bar_q1, 0);
        load_qk(1);
      }
      uint32_t n_v = 0;                       // V loads consumed so far (phase of bar_v)
      for (int g = 0; g < G; ++g) {
        const uint32_t ph = (uint32_t)g & 1u;
        if (g + 1 < G) {
          ptx::mbar_wait(bar_sfree, ph);      // S(g) is in registers: the accumulator may be overwritten
          ptx::mbar_wait(bar_qk, ph ^ 1u);
          ptx::tc_fence_after();
          mma1();
          if (g + 2 < G) {                    // prefetch the tile after next the moment MMA1(g+1) has read Q (and K)
            ptx::mbar_wait(bar_q1, ph ^ 1u);
            load_qk(g + 2);
          }
        }
        // ---- O = P V: A = P (K-major, S/64 blocks of 128 x 64), B = V as [key][64] (MN-major: bit 16, 2048 B per K = 16 step) ----
        if (g % NG == 0) { ptx::mbar_wait(bar_v, n_v & 1u); ++n_v; }
        ptx::mbar_wait(bar_p, ph);
        if (g > 0) ptx::mbar_wait(bar_ofree, ph ^ 1u);
        ptx::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < S / 64; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(smem_base + L::p_off + (uint32_t)kb * (QT * 128) + (uint32_t)k * 32u);
            const uint64_t db = ptx::make_smem_desc_sw128(smem_base + L::v_off + (uint32_t)kb * (64 * 128) + (uint32_t)k * 2048u);
            ptx::umma_f16(tmem_base + (uint32_t)S, da, db, idesc2, (kb | k) ? 1u : 0u);
          }
        }
        ptx::umma_commit(bar_o);
        if (g + 1 < G && (g + 1) % NG == 0) {
          ptx::mbar_wait(bar_o, ph);          // MMA2 of the item's last tile has read V: load the next item's
          load_v(g + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax warps =====================
    const int q = warp & 3, cg = warp >> 2;
    const int row = q * 32 + lane;
    const float sc = args.scale_log2;
    const uint32_t rsw = (uint32_t)(row & 7);
    int cur_n = -1;
    // ===================== output of tile ge: O / sum -> bf16 -> y[n][s][h][..]; column groups 0 (and 1) take 64 / EG columns each ============
    constexpr int EG = NG >= 2 ? 2 : 1;       // warps that share the output work per lane quarter
    constexpr int ECOLS = DH / EG;
    auto epilogue = [&](int ge) {
      if (cg >= EG) return;
      const uint32_t phe = (uint32_t)ge & 1u;
      const int nhe = (int)blockIdx.x + (ge / NG) * (int)gridDim.x, qte = ge % NG;
      const int ne = nhe / args.H, he = nhe - ne * args.H;
      ptx::mbar_wait(bar_o, phe);
      ptx::tc_fence_after();
      uint32_t o0[32], o1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)S + (uint32_t)(cg * ECOLS);
      ptx::tmem_ld_32x32(taddr, o0);
      if (EG == 1) ptx::tmem_ld_32x32(taddr + 32, o1);
      const float* e_sum = s_sum + phe * (NG * QT);
      float total = 0.f;
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) total += e_sum[gg * QT + row];
      const float inv = 1.f / total;
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_ofree);   // MMA2 of the next tile may overwrite O
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<T*>(args.y) + ((size_t)(ne * S + qte * QT + row) * args.H + he) * DH + cg * ECOLS);
#pragma unroll
      for (int ch = 0; ch < ECOLS / 8; ++ch) {
        const uint32_t* v = ch < 4 ? &o0[ch * 8] : &o1[(ch - 4) * 8];
        dst[ch] = make_uint4(pack2<T>(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv), pack2<T>(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv),
                             pack2<T>(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv), pack2<T>(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv));
      }
    };
    for (int g = 0; g < G; ++g) {
      const uint32_t ph = (uint32_t)g & 1u;
      const int nh = (int)blockIdx.x + (g / NG) * (int)gridDim.x, qt = g % NG;
      const int n = nh / args.H;
      (void)qt;
      if (n != cur_n) {
        // mask row of this image -> shared memory, fp32, in the log2 domain (every warp is past the previous tile's mask reads: they
        // precede that tile's row-max barrier)
        cur_n = n;
        const T* mrow = args.mask ? reinterpret_cast<const T*>(args.mask) + (size_t)(args.mask_batch > 1 ? n : 0) * S : nullptr;
        for (int i = threadIdx.x; i < S; i += NSW * 32) s_mask[i] = mrow ? to_f32<T>(mrow[i]) * LOG2E : 0.f;
        ptx::named_bar_sync(1, NSW * 32);
      }
      ptx::mbar_wait(bar_s, ph);
      ptx::tc_fence_after();
      float x[128];
      {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 128);
        uint32_t(&xu)[128] = reinterpret_cast<uint32_t(&)[128]>(x);
        ptx::tmem_ld_32x32(taddr, reinterpret_cast<uint32_t(&)[32]>(xu[0]));
        ptx::tmem_ld_32x32(taddr + 32, reinterpret_cast<uint32_t(&)[32]>(xu[32]));
        ptx::tmem_ld_32x32(taddr + 64, reinterpret_cast<uint32_t(&)[32]>(xu[64]));
        ptx::tmem_ld_32x32(taddr + 96, reinterpret_cast<uint32_t(&)[32]>(xu[96]));
        ptx::tmem_ld_wait();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_sfree);     // the issuer may start MMA1 of the next tile
      float mx = -3.0e38f;
      const uint32_t mask_addr = smem_base + L::mask_off + (uint32_t)(cg * 128) * 4u;
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
        const uint4 mv = ptx::ld_shared_v4(mask_addr + (uint32_t)c * 4u);     // same address in every lane: broadcast
        x[c + 0] = fmaf(x[c + 0], sc, __uint_as_float(mv.x)); x[c + 1] = fmaf(x[c + 1], sc, __uint_as_float(mv.y));
        x[c + 2] = fmaf(x[c + 2], sc, __uint_as_float(mv.z)); x[c + 3] = fmaf(x[c + 3], sc, __uint_as_float(mv.w));
        mx = fmaxf(fmaxf(mx, fmaxf(x[c + 0], x[c + 1])), fmaxf(x[c + 2], x[c + 3]));
      }
      float* t_max = s_max + ph * (NG * QT);
      float* t_sum = s_sum + ph * (NG * QT);
      if (NG > 1) {
        t_max[cg * QT + row] = mx;
        ptx::named_bar_sync(1, NSW * 32);
#pragma unroll
        for (int gg = 0; gg < NG; ++gg) mx = fmaxf(mx, t_max[gg * QT + row]);
      }
      // the P buffer is free once MMA2 of the previous tile has completed
      if (g > 0) ptx::mbar_wait(bar_o, ph ^ 1u);
      // P = 2^(x - max) as bf16/fp16, written as the K-major SWIZZLE_128B A operand: key column c lives in K-block c / 64 (128 rows x
      // 128 B), 16-byte chunk (c % 64) / 8 XOR (row & 7) of this row
      float sum = 0.f;
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { e[i] = ex2_approx(x[ch * 8 + i] - mx); }
        const uint32_t p0 = pack2<T>(e[0], e[1]), p1 = pack2<T>(e[2], e[3]), p2 = pack2<T>(e[4], e[5]), p3 = pack2<T>(e[6], e[7]);
        // row sum over the unrounded values: the rounding errors of the 16-bit P are zero-mean, so the sum of the rounded values differs
        // by ~2^-9 / sqrt(S) relative (1e-4 at S = 384) - and summing them would cost an unpack + add per element in an issue-bound loop
        sum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
        const int col = cg * 128 + ch * 8;
        const uint32_t dst = smem_base + L::p_off + (uint32_t)(col >> 6) * (QT * 128) + (uint32_t)row * 128u + ((((uint32_t)(col & 63) >> 3) ^ rsw) << 4);
        ptx::st_shared_v4(dst, make_uint4(p0, p1, p2, p3));
      }
      t_sum[cg * QT + row] = sum;
      ptx::fence_proxy_async();               // P (generic-proxy stores) -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_p);
      // the output of the PREVIOUS tile: by now its MMA2 has long completed, so nobody waits for the tensor core here, and the next
      // tile's LDTM / scale / max phase of every warp overlaps this tile's MMA2
      if (g > 0) epilogue(g - 1);
    }
    if (G > 0) epilogue(G - 1);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == NSW) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <typename K>
void launch_attn_pdl(K kernel, int grid, int threads, unsigned smem, cudaStream_t s, const AttnArgs& a) {
  static const bool pdl = [] { const char* e = getenv("B200_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, a);
}

template <typename T, int S>
void launch_typed(const AttnArgs& a, cudaStream_t s) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  launch_attn_pdl(attention_kernel<T, S>, std::min(a.N * a.H, sms), 32 * (S / 32 + 1), AttnSmem<S>::total, s, a);
}
template <typename T>
void opt_in() {
  cudaFuncSetAttribute(attention_kernel<T, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AttnSmem<128>::total);
  cudaFuncSetAttribute(attention_kernel<T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AttnSmem<256>::total);
  cudaFuncSetAttribute(attention_kernel<T, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AttnSmem<384>::total);
}

}  // namespace

void attention_init_device() {
  opt_in<__nv_bfloat16>();
  opt_in<__half>();
}

bool attention_supported(const AttnParams& p, int dtype) {
  const char* off = getenv("B200_FUSED_ATTENTION");
  if (off && off[0] == '0') return false;
  return (dtype == B200_BF16 || dtype == B200_F16) && p.dh == DH && (p.S == 128 || p.S == 256 || p.S == 384) && p.N >= 1 && p.H >= 1 &&
         (int64_t)p.N * p.H * p.S < ((int64_t)1 << 31) && (p.mask_batch == 1 || p.mask_batch == p.N || !p.mask);
}

void launch_attention(const AttnParams& p, int dtype, cudaStream_t s) {
  B200_CHECK(attention_supported(p, dtype), B200_ERR_UNSUPPORTED, "fused attention: unsupported geometry");
  const bool bf16 = dtype == B200_BF16;
  AttnArgs a;
  const uint64_t rows = (uint64_t)p.N * p.H * p.S;
  a.tm_q = make_tmap_2d(p.q, bf16, DH, rows, DH * 2, DH, 128);
  a.tm_k = make_tmap_2d(p.k, bf16, DH, rows, DH * 2, DH, 128);
  a.tm_v = make_tmap_2d(p.v, bf16, DH, rows, DH * 2, DH, 128);
  a.y = p.y;
  a.mask = p.mask;
  a.N = p.N; a.H = p.H; a.S = p.S; a.mask_batch = p.mask_batch;
  a.scale_log2 = p.scale * LOG2E;
#define B200_ATTN(T)                                      \
  switch (p.S) {                                          \
    case 128: launch_typed<T, 128>(a, s); break;          \
    case 256: launch_typed<T, 256>(a, s); break;          \
    default: launch_typed<T, 384>(a, s); break;           \
  }
  if (bf16) { B200_ATTN(__nv_bfloat16) } else { B200_ATTN(__half) }
#undef B200_ATTN
}

}  // namespace b200
