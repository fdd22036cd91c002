// Implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   D[M, Cout] = im2col(X)[M, K] * Wp[Cout, K]^T   (+ bias, activation, residual, activation)
//
// replaces dnnl::convolution_forward / inner_product_forward with post-ops as configured by
// MKLDNNConvolutionNode::setPostOps (inference-engine/src/mkldnn_plugin/nodes/mkldnn_conv_node.cpp:355-403)
// and MKLDNNFullyConnectedNode (nodes/mkldnn_fullyconnected_node.cpp:123-132).  The arithmetic
// contract is ngraph/core/reference/include/ngraph/runtime/reference/convolution.hpp:89-186
// (cross-correlation, zero padding, fp32 accumulation).
//
// Structure of one CTA (one 128 x BN output tile, 160 threads, 2 CTAs resident per SM so one CTA's
// epilogue overlaps the other's main loop):
//   warps 0-3  producers: gather the A tile (128 output pixels x 64 K-elements) straight from the
//              NHWC activation tensor with 16-byte cp.async (zero-fill for padding/tails) into the
//              128B-swizzled K-major layout the UMMA descriptor expects, same for the weight tile;
//              afterwards the same warps run the epilogue (TMEM -> registers -> bias/act/residual -> HBM)
//   warp 4     TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit releases smem stages
// Synchronisation is mbarrier only: full[s] (producers -> MMA), empty[s] (MMA commit -> producers),
// tmem_full (MMA commit -> epilogue).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "act.cuh"
#include "common.h"
#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tc_common.cuh"

namespace b200 {

namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 x 2 B = 128 B = one swizzle row
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int NUM_PRODUCERS = 128;

template <int BN> struct TcCfg;
template <> struct TcCfg<64>  { static constexpr int STAGES = 4, LAG = 2; };
template <> struct TcCfg<128> { static constexpr int STAGES = 3, LAG = 1; };

template <int BN>
constexpr int tc_smem_bytes() {
  return TcCfg<BN>::STAGES * (A_STAGE_BYTES + BN * 128) + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <typename T, int BN, bool GEN>
__global__ void __launch_bounds__(160, 2) conv_tc_kernel(const ConvParams p, const EpiAct e1, const EpiAct e2) {
  constexpr int STAGES = TcCfg<BN>::STAGES;
  constexpr int LAG = TcCfg<BN>::LAG;
  constexpr int B_STAGE_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_al + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* res_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = p.Cout_alloc / BN;
  const int n_tile = blockIdx.x % n_tiles;
  const int m_tile = blockIdx.x / n_tiles;
  const int m0 = m_tile * BM;
  const int n0 = n_tile * BN;
  const int nkb = p.Ktot / BK;
  const int bz = blockIdx.z;

  const T* __restrict__ xg = reinterpret_cast<const T*>(p.x) + (int64_t)bz * p.x_bs;
  const T* __restrict__ wg = reinterpret_cast<const T*>(p.w) + (int64_t)bz * p.w_bs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], NUM_PRODUCERS);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::mbar_init(res_bar, NUM_PRODUCERS);
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(tmem_slot, BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================== producers =====================
    const int t = threadIdx.x;
    const int c = t & 7;          // 16-byte chunk within the 128-byte K row
    const int r0 = t >> 3;        // first row handled by this thread (rows r0 + 16*i)
    const uint32_t sw_off = (uint32_t)((c ^ (r0 & 7)) << 4);
    const int HoWo = p.Ho * p.Wo;
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
    const int cpc = p.Cin_p >> 3;       // 16-byte chunks per filter tap
    const int taps = p.kh * p.kw;
    int cc = c, tap = 0, fr = 0, fs = 0;  // chunk within tap, tap index, filter row/col
    while (cc >= cpc) { cc -= cpc; ++tap; if (++fs == p.kw) { fs = 0; ++fr; } }

    for (int kb = 0; kb < nkb + LAG; ++kb) {
      if (kb < nkb) {
        const int s = kb % STAGES;
        ptx::mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
        const uint32_t a_dst = smem_base + s * STAGE_BYTES + r0 * 128 + sw_off;
        const uint32_t b_dst = a_dst + A_STAGE_BYTES;
        const bool kvalid = tap < taps;
        const int dh_off = fr * p.dh, dw_off = fs * p.dw;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int hi = hi0[i] + dh_off, wi = wi0[i] + dw_off;
          bool ok = kvalid && (unsigned)hi < (unsigned)p.H && (unsigned)wi < (unsigned)p.W;
          const T* src = ok ? xg + xoff[i] + ((int64_t)hi * p.W + wi) * p.Cin_p + cc * 8 : xg;
          ptx::cp_async16(a_dst + i * (16 * 128), src, ok ? 16u : 0u);
        }
        const int wk = kb * BK + c * 8;
        const bool wk_ok = wk < p.w_k;
        const T* wsrc = wg + (int64_t)(n0 + r0) * p.w_ld + wk;
#pragma unroll
        for (int j = 0; j < BN / 16; ++j) {
          const bool ok = wk_ok && (n0 + r0 + 16 * j) < p.w_rows;
          ptx::cp_async16(b_dst + j * (16 * 128), ok ? wsrc + (int64_t)j * 16 * p.w_ld : wg, ok ? 16u : 0u);
        }
        // advance (tap, chunk) by one K block = 8 chunks
        cc += 8;
        while (cc >= cpc) { cc -= cpc; ++tap; if (++fs == p.kw) { fs = 0; ++fr; } }
      }
      ptx::cp_async_commit();
      if (kb >= LAG) {
        ptx::cp_async_wait<LAG>();
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&full_bar[(kb - LAG) % STAGES]);
      }
    }

    // ===================== epilogue =====================
    // TMEM -> registers -> (bias, act, + residual, act) -> padded staging tile in the (now idle) pipeline
    // smem -> one bulk async store per output row.  The residual tile comes in the same way (bulk async
    // loads, one per row), so every HBM access of the epilogue is a full, coalesced row segment.
    const int row = warp * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    const int ncols = min(BN, p.Cout_p - n0);                 // valid output channels of this tile (multiple of 8)
    const uint32_t esz = p.out_f32 ? 4u : 2u;
    const uint32_t out_stride = BN * esz + 16u;               // +16 B: conflict-free 16-byte row-strided smem access
    const uint32_t res_stride = BN * 2u + 16u;
    const uint32_t srow = smem_base + row * out_stride;
    const uint32_t rrow = smem_base + BM * out_stride + row * res_stride;
    const int64_t yoff = (int64_t)bz * p.y_bs + (int64_t)m * p.ldy + n0;
    const bool has_res = p.res != nullptr;
    if (has_res && row_ok) ptx::bulk_prefetch_l2(reinterpret_cast<const T*>(p.res) + yoff, (uint32_t)ncols * 2u);

    ptx::mbar_wait(tmem_full_bar, 0);   // all MMAs retired: accumulator ready, smem stages free
    ptx::tc_fence_after();
    if (has_res) {
      if (row_ok) {
        ptx::mbar_arrive_expect_tx(res_bar, (uint32_t)ncols * 2u);
        ptx::bulk_load(rrow, reinterpret_cast<const T*>(p.res) + yoff, (uint32_t)ncols * 2u, res_bar);
      } else {
        ptx::mbar_arrive(res_bar);
      }
    }
    const float* __restrict__ bias = p.bias + n0;
    bool res_ready = !has_res;
#pragma unroll 1
    for (int j = 0; j < BN / 32; ++j) {
      if (j * 32 >= ncols) break;       // warp-uniform
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), v);
      ptx::tmem_ld_wait();
      if (!res_ready) { ptx::mbar_wait(res_bar, 0); res_ready = true; }
#pragma unroll
      for (int q = 0; q < 4; ++q) {     // groups of 8 output channels
        const int c0 = j * 32 + q * 8;
        if (c0 < ncols) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
          float f[8];
          f[0] = __uint_as_float(v[q * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[q * 8 + 1]) + b0.y;
          f[2] = __uint_as_float(v[q * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[q * 8 + 3]) + b0.w;
          f[4] = __uint_as_float(v[q * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[q * 8 + 5]) + b1.y;
          f[6] = __uint_as_float(v[q * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[q * 8 + 7]) + b1.w;
          epi_act8<GEN>(f, e1);
          if (has_res) {
            const uint4 rv = ptx::ld_shared_v4(rrow + (uint32_t)c0 * 2u);
            float r[8];
            unpack2<T>(rv.x, r[0], r[1]); unpack2<T>(rv.y, r[2], r[3]);
            unpack2<T>(rv.z, r[4], r[5]); unpack2<T>(rv.w, r[6], r[7]);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += r[e];
            epi_act8<GEN>(f, e2);
          }
          if (p.out_f32) {
            ptx::st_shared_v4(srow + (uint32_t)c0 * 4u, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
            ptx::st_shared_v4(srow + (uint32_t)c0 * 4u + 16u, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
          } else {
            ptx::st_shared_v4(srow + (uint32_t)c0 * 2u, make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7])));
          }
        }
      }
    }
    ptx::fence_proxy_async();           // this thread's staging row -> visible to the bulk-copy engine
    if (row_ok) {
      char* ydst = reinterpret_cast<char*>(p.y) + yoff * (int64_t)esz;
      ptx::bulk_store(ydst, srow, (uint32_t)ncols * esz);
      ptx::bulk_commit();
      ptx::bulk_wait_read0();           // smem must stay intact until the copy engine has read it
    }
  } else if (lane == 0) {
    // ===================== MMA issuer (one thread) =====================
    constexpr uint32_t idesc = ptx::make_idesc_f16(TcType<T>::is_bf16, BM, BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      ptx::mbar_wait(&full_bar[s], (kb / STAGES) & 1);
      ptx::tc_fence_after();
      const uint32_t a_addr = smem_base + s * STAGE_BYTES;
      const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) {
        const uint64_t da = ptx::make_smem_desc_sw128(a_addr + k * 32);
        const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * 32);
        ptx::umma_f16(tmem_base, da, db, idesc, (kb | k) ? 1u : 0u);
      }
      ptx::umma_commit(&empty_bar[s]);   // frees this smem stage once the MMAs above have read it
    }
    ptx::umma_commit(tmem_full_bar);     // accumulator complete -> epilogue
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem_base, BN);
}

template <typename T, int BN>
void launch_tc(const ConvParams& p, cudaStream_t s) {
  constexpr int smem = tc_smem_bytes<BN>();
  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = p.Cout_alloc / BN;
  dim3 grid((unsigned)(m_tiles * n_tiles), 1, (unsigned)p.batch);
  auto epi = [](int kind, float a, float b) {
    EpiAct e;
    e.kind = kind; e.a = a; e.b = b;
    e.lo = kind == ACT_RELU ? 0.f : a;
    e.hi = kind == ACT_RELU ? __builtin_huge_valf() : b;
    return e;
  };
  if (conv_acts_are_clamps(p)) conv_tc_kernel<T, BN, false><<<grid, 160, smem, s>>>(p, epi(p.act1, p.a1, p.b1), epi(p.act2, p.a2, p.b2));
  else conv_tc_kernel<T, BN, true><<<grid, 160, smem, s>>>(p, epi(p.act1, p.a1, p.b1), epi(p.act2, p.a2, p.b2));
}

}  // namespace

int conv_tc_pick_bn(int cout_p) { return cout_p <= 64 ? 64 : 128; }

// Opt in to >48 KB dynamic shared memory on the current device (called once per device at compile).
void conv_tc_init_device() {
  auto opt_in = [](auto kernel, int bytes) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); };
  opt_in(conv_tc_kernel<__nv_bfloat16, 64, false>, tc_smem_bytes<64>());
  opt_in(conv_tc_kernel<__nv_bfloat16, 64, true>, tc_smem_bytes<64>());
  opt_in(conv_tc_kernel<__nv_bfloat16, 128, false>, tc_smem_bytes<128>());
  opt_in(conv_tc_kernel<__nv_bfloat16, 128, true>, tc_smem_bytes<128>());
  opt_in(conv_tc_kernel<__half, 64, false>, tc_smem_bytes<64>());
  opt_in(conv_tc_kernel<__half, 64, true>, tc_smem_bytes<64>());
  opt_in(conv_tc_kernel<__half, 128, false>, tc_smem_bytes<128>());
  opt_in(conv_tc_kernel<__half, 128, true>, tc_smem_bytes<128>());
}

void launch_conv_tc(const ConvParams& p, int dtype, int bn, cudaStream_t s) {
  B200_CHECK(!(p.out_f32 && p.res), B200_ERR_UNSUPPORTED, "conv_tc: fp32 output with a residual input");
  B200_CHECK(p.Ktot % BK == 0 && p.Cin_p % 8 == 0 && p.Cout_alloc % bn == 0 && p.Cout_p % 8 == 0 && p.ldy % 8 == 0 &&
                 p.w_ld % 8 == 0 && p.w_k % 8 == 0,
             B200_ERR_INVALID, "conv_tc: bad geometry");
  if (dtype == B200_BF16) {
    if (bn == 64) launch_tc<__nv_bfloat16, 64>(p, s); else launch_tc<__nv_bfloat16, 128>(p, s);
  } else if (dtype == B200_F16) {
    if (bn == 64) launch_tc<__half, 64>(p, s); else launch_tc<__half, 128>(p, s);
  } else {
    B200_THROW(B200_ERR_UNSUPPORTED, "conv_tc: dtype must be bf16/f16");
  }
}

}  // namespace b200
