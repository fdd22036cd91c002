// CUDA-core fp32 implicit-GEMM convolution / GEMM (FFMA, fp32 accumulate) used by the plugin's
// FP32 inference mode, where outputs must match the reference CPU plugin to 1e-4 relative L2
// (TF32 tensor-core math is not accurate enough).  Same geometry / epilogue contract as
// conv_tc.cu; reference semantics: ngraph/core/reference/include/ngraph/runtime/reference/convolution.hpp:89-186.
#include "act.cuh"
#include "common.h"
#include "kernels.h"

namespace b200 {
namespace {

constexpr int FBM = 128, FBN = 64, FBK = 16;

__global__ void __launch_bounds__(256, 2) conv_f32_kernel(const ConvParams p) {
  __shared__ __align__(16) float As[2][FBK][FBM + 4];
  __shared__ __align__(16) float Bs[2][FBK][FBN + 4];

  const int tid = threadIdx.x;
  const int n_tiles = p.Cout_alloc / FBN;
  const int n_tile = blockIdx.x % n_tiles;
  const int m_tile = blockIdx.x / n_tiles;
  const int m0 = m_tile * FBM, n0 = n_tile * FBN;
  const int bz = blockIdx.z;
  const float* __restrict__ xg = reinterpret_cast<const float*>(p.x) + (int64_t)bz * p.x_bs;
  const float* __restrict__ wg = reinterpret_cast<const float*>(p.w) + (int64_t)bz * p.w_bs;

  // loader mapping: A rows (tid>>2) and (tid>>2)+64, 4 consecutive k at (tid&3)*4; B row tid>>2
  const int lk = (tid & 3) * 4;
  const int lr = tid >> 2;
  const int HoWo = p.Ho * p.Wo;
  int hi0[2], wi0[2];
  int64_t xoff[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int m = m0 + lr + 64 * i;
    bool rv = m < p.M;
    int mm = rv ? m : 0;
    int n = mm / HoWo;
    int rem = mm - n * HoWo;
    int ho = rem / p.Wo, wo = rem - (rem / p.Wo) * p.Wo;
    hi0[i] = rv ? ho * p.sh - p.pt : -(1 << 28);
    wi0[i] = wo * p.sw - p.pl;
    xoff[i] = (int64_t)n * p.H * p.W * p.Cin_p;
  }
  const int taps = p.kh * p.kw;
  const int nkb = p.Ktot / FBK;

  float4 ra[2], rb;
  auto load_tile = [&](int kb) {
    int k = kb * FBK + lk;            // 4 consecutive k stay inside one tap (Cin_p % 8 == 0)
    int tap = k / p.Cin_p;
    int ci = k - tap * p.Cin_p;
    int fr = tap / p.kw, fs = tap - fr * p.kw;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int hi = hi0[i] + fr * p.dh, wi = wi0[i] + fs * p.dw;
      bool ok = tap < taps && (unsigned)hi < (unsigned)p.H && (unsigned)wi < (unsigned)p.W;
      ra[i] = ok ? __ldg(reinterpret_cast<const float4*>(xg + xoff[i] + ((int64_t)hi * p.W + wi) * p.Cin_p + ci))
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    rb = (k < p.w_k && n0 + lr < p.w_rows) ? __ldg(reinterpret_cast<const float4*>(wg + (int64_t)(n0 + lr) * p.w_ld + k))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][lk + 0][lr + 64 * i] = ra[i].x;
      As[buf][lk + 1][lr + 64 * i] = ra[i].y;
      As[buf][lk + 2][lr + 64 * i] = ra[i].z;
      As[buf][lk + 3][lr + 64 * i] = ra[i].w;
    }
    Bs[buf][lk + 0][lr] = rb.x;
    Bs[buf][lk + 1][lr] = rb.y;
    Bs[buf][lk + 2][lr] = rb.z;
    Bs[buf][lk + 3][lr] = rb.w;
  };

  const int tx = tid & 15;   // 4 output channels each
  const int ty = tid >> 4;   // 8 output rows each
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkb) load_tile(kb + 1);
#pragma unroll
    for (int k = 0; k < FBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kb + 1 < nkb) store_tile(buf ^ 1);
    __syncthreads();
  }

  const int nc = n0 + tx * 4;
  if (nc >= p.Cout_p) return;   // Cout_p % 8 == 0 so a group of 4 is all-in or all-out
  const float4 bias = __ldg(reinterpret_cast<const float4*>(p.bias + nc));
  const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    int64_t off = (int64_t)bz * p.y_bs + (int64_t)m * p.ldy + nc;
    float f[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) f[j] = act_apply(acc[i][j] + bb[j], p.act1, p.a1, p.b1);
    if (p.res != nullptr) {
      float4 r = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + off));
      f[0] = act_apply(f[0] + r.x, p.act2, p.a2, p.b2);
      f[1] = act_apply(f[1] + r.y, p.act2, p.a2, p.b2);
      f[2] = act_apply(f[2] + r.z, p.act2, p.a2, p.b2);
      f[3] = act_apply(f[3] + r.w, p.act2, p.a2, p.b2);
    }
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + off) = make_float4(f[0], f[1], f[2], f[3]);
  }
}

}  // namespace

void launch_conv_f32(const ConvParams& p, cudaStream_t s) {
  B200_CHECK(p.Ktot % FBK == 0 && p.Cin_p % 8 == 0 && p.Cout_alloc % FBN == 0 && p.Cout_p % 8 == 0 && p.ldy % 4 == 0 &&
                 p.w_ld % 4 == 0 && p.w_k % 4 == 0,
             B200_ERR_INVALID, "conv_f32: bad geometry");
  const int m_tiles = (p.M + FBM - 1) / FBM;
  const int n_tiles = p.Cout_alloc / FBN;
  dim3 grid((unsigned)(m_tiles * n_tiles), 1, (unsigned)p.batch);
  conv_f32_kernel<<<grid, 256, 0, s>>>(p);
}

}  // namespace b200
