// Memory-bound convolution kernels of the MobileNet path (CUDA cores, 16-bit NHWC8 data, fp32 accumulate):
//   * dwconv3x3_strip_kernel: depthwise 3x3, stride 1 / 2.  One thread = 4 channels x 2 output columns x a strip of R output rows: every
//     input vector of the (R-1)*S+3 rows x (2S+2 | 2S+3) columns is loaded ONCE (all loads issued before the first use), the nine weight
//     vectors once, and the window slides down the strip in registers - 2 (stride 1) to 6 (stride 2) vector loads per output instead
//     of 6 to 7.5 with one thread per output pair, a quarter of the conversions.
//   * conv_small_cin_kernel: dense convolution over <= 8 input channels (the C = 3 first layer of MobileNetV2 / EfficientNet style
//     networks): one 16-byte vector per input pixel, weights staged once per block in shared memory as fp32.  As an implicit GEMM the
//     layer has K = 27 padded to 64+ and ran in the gather mode of the tensor-core kernel at 10 % of the step.
// Reference semantics: nodes/mkldnn_conv_node.cpp:78-100 (depthwise = GroupConvolution with G == C), bias + activation fused as
// mkldnn_graph_optimizer.cpp:150-182,587-629 fuses them.
#include "act.cuh"
#include "common.h"
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace b200 {
namespace {

template <typename T> __device__ __forceinline__ void unpack4(const uint2& u, float (&f)[4]);
template <> __device__ __forceinline__ void unpack4<__half>(const uint2& u, float (&f)[4]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
  f[0] = __low2float(h[0]); f[1] = __high2float(h[0]); f[2] = __low2float(h[1]); f[3] = __high2float(h[1]);
}
template <> __device__ __forceinline__ void unpack4<__nv_bfloat16>(const uint2& u, float (&f)[4]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
}
template <typename T> __device__ __forceinline__ uint2 pack4(const float (&f)[4]);
template <> __device__ __forceinline__ uint2 pack4<__half>(const float (&f)[4]) {
  uint2 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
  h[0] = __floats2half2_rn(f[0], f[1]); h[1] = __floats2half2_rn(f[2], f[3]);
  return u;
}
template <> __device__ __forceinline__ uint2 pack4<__nv_bfloat16>(const float (&f)[4]) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(f[0], f[1]); h[1] = __floats2bfloat162_rn(f[2], f[3]);
  return u;
}

template <typename T, int S, int R>
__global__ void __launch_bounds__(256) dwconv3x3_strip_kernel(const uint2* __restrict__ x, const uint2* __restrict__ w, const float* __restrict__ bias,
                                                              uint2* __restrict__ y, int H, int W, int cq, int Ho, int Wo, int strips, int pt, int pl,
                                                              int act, float aa, float ab, uint32_t total) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Wo2 = (Wo + 1) >> 1;
  const uint32_t c4 = idx % (uint32_t)cq;
  uint32_t t = idx / (uint32_t)cq;
  const int j = (int)(t % (uint32_t)Wo2);
  t /= (uint32_t)Wo2;
  const int strip = (int)(t % (uint32_t)strips);
  const int n = (int)(t / (uint32_t)strips);
  constexpr int NC = 2 * S + 2;               // input columns feeding two adjacent outputs: 4 (stride 1), 5 used of 6 (stride 2)
  constexpr int NCU = S == 1 ? 4 : 5;
  constexpr int NR = (R - 1) * S + 3;         // input rows feeding R output rows
  (void)NC;
  const uint2* xi = x + (size_t)n * H * W * cq + c4;
  const int ho0 = strip * R;
  const int c0 = 2 * j * S - pl, r0 = ho0 * S - pt;
  uint2 raw[NR][NCU];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int hi = r0 + r;
#pragma unroll
    for (int c = 0; c < NCU; ++c) {
      const int wi = c0 + c;
      const bool in = (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
      raw[r][c] = in ? __ldg(xi + ((size_t)hi * W + wi) * cq) : make_uint2(0u, 0u);
    }
  }
  float wv[9][4];
#pragma unroll
  for (int k = 0; k < 9; ++k) unpack4<T>(__ldg(w + (size_t)k * cq + c4), wv[k]);
  const float4 bq = __ldg(reinterpret_cast<const float4*>(bias) + c4);
  uint2* yo = y + (((size_t)n * Ho + ho0) * Wo + 2 * j) * cq + c4;
  const bool second = 2 * j + 1 < Wo;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    if (ho0 + k >= Ho) break;
    float acc0[4] = {bq.x, bq.y, bq.z, bq.w}, acc1[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float v[NCU][4];
#pragma unroll
      for (int c = 0; c < NCU; ++c) unpack4<T>(raw[k * S + r][c], v[c]);
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc0[e] = fmaf(v[s][e], wv[r * 3 + s][e], acc0[e]);
          acc1[e] = fmaf(v[s + S][e], wv[r * 3 + s][e], acc1[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc0[e] = act_apply(acc0[e], act, aa, ab); acc1[e] = act_apply(acc1[e], act, aa, ab); }
    yo[(size_t)k * Wo * cq] = pack4<T>(acc0);
    if (second) yo[(size_t)k * Wo * cq + cq] = pack4<T>(acc1);
  }
}

// Stride-1 layers (13 of MobileNetV2's 17): the three input rows of the window stay UNPACKED in registers and roll down a strip of R
// output rows - every input value is converted once (the strip kernel above re-converts a row for each of the three output rows that
// use it), the next two input rows are in flight while a row is computed, and the activation is a branch-free clamp.  ncu on the strip
// kernel: 49 % issue-slot utilisation at 24 % occupancy, 40 instructions per output value; this form needs ~16.
template <typename T, int R>
__global__ void __launch_bounds__(256, 2) dwconv3x3_roll_kernel(const uint2* __restrict__ x, const uint2* __restrict__ w, const float* __restrict__ bias,
                                                             uint2* __restrict__ y, int H, int W, int cq, int Ho, int Wo, int strips, int pt, int pl, float lo,
                                                             float hi, uint32_t total) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Wo2 = (Wo + 1) >> 1;
  const uint32_t c4 = idx % (uint32_t)cq;
  uint32_t t = idx / (uint32_t)cq;
  const int j = (int)(t % (uint32_t)Wo2);
  t /= (uint32_t)Wo2;
  const int strip = (int)(t % (uint32_t)strips);
  const int n = (int)(t / (uint32_t)strips);
  constexpr int NR = R + 2;
  const uint2* xi = x + (size_t)n * H * W * cq + c4;
  const int ho0 = strip * R;
  const int c0 = 2 * j - pl, r0 = ho0 - pt;
  bool cin[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) cin[c] = (unsigned)(c0 + c) < (unsigned)W;
  auto load_row = [&](int i, uint2 (&dst)[4]) {      // input row r0 + i, columns c0 .. c0 + 3; outside the image: zeros
    const int hi_ = r0 + i;
    const bool rin = (unsigned)hi_ < (unsigned)H;
    const uint2* row = xi + (size_t)(rin ? hi_ : 0) * W * cq;
#pragma unroll
    for (int c = 0; c < 4; ++c) dst[c] = (rin && cin[c]) ? __ldg(row + (size_t)(c0 + c) * cq) : make_uint2(0u, 0u);
  };
  uint2 first[3][4], pf[2][4];
#pragma unroll
  for (int i = 0; i < 3; ++i) load_row(i, first[i]);
  load_row(3, pf[0]);
  if (NR > 4) load_row(4, pf[1]);
  float wv[9][4];
#pragma unroll
  for (int k = 0; k < 9; ++k) unpack4<T>(__ldg(w + (size_t)k * cq + c4), wv[k]);
  const float4 bq = __ldg(reinterpret_cast<const float4*>(bias) + c4);
  float win[3][4][4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) unpack4<T>(first[i][c], win[i][c]);
  uint2* yo = y + (((size_t)n * Ho + ho0) * Wo + 2 * j) * cq + c4;
  const bool second = 2 * j + 1 < Wo;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    if (ho0 + k >= Ho) break;
    float acc0[4] = {bq.x, bq.y, bq.z, bq.w}, acc1[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int slot = (k + r) % 3;
#pragma unroll
      for (int s_ = 0; s_ < 3; ++s_)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc0[e] = fmaf(win[slot][s_][e], wv[r * 3 + s_][e], acc0[e]);
          acc1[e] = fmaf(win[slot][s_ + 1][e], wv[r * 3 + s_][e], acc1[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc0[e] = fminf(fmaxf(acc0[e], lo), hi); acc1[e] = fminf(fmaxf(acc1[e], lo), hi); }
    yo[(size_t)k * Wo * cq] = pack4<T>(acc0);
    if (second) yo[(size_t)k * Wo * cq + cq] = pack4<T>(acc1);
    if (k + 1 < R) {
      // row k + 3 arrives in the slot of row k; the prefetch buffer it came from takes row k + 5
#pragma unroll
      for (int c = 0; c < 4; ++c) unpack4<T>(pf[k & 1][c], win[k % 3][c]);
      if (k + 5 < NR) load_row(k + 5, pf[k & 1]);
    }
  }
}

// Dense convolution, Cin_p == 8 (one uint4 per input pixel), Cout_p % 8 == 0.  One thread = PX horizontally adjacent output pixels x 8
// output channels: a weight vector read from shared memory serves PX pixels.  (One pixel per thread saturated the shared-memory
// pipe - ncu: L1/TEX 91 % busy, 97 M warp instructions, 145 us for MobileNetV2's first layer at batch 64.)
// wsm: fp32 weights [tap][ci][Cout_p] for the CIN real input channels.
template <typename T, int CIN, int PX>
__global__ void __launch_bounds__(256) conv_small_cin_kernel(const uint4* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                                                             uint4* __restrict__ y, int N, int H, int W, int Ho, int Wo, int Cout_p, int ldy8, int kh, int kw,
                                                             int sh, int sw, int dh, int dw, int pt, int pl, int w_ld, int act, float aa, float ab,
                                                             uint32_t total) {
  extern __shared__ float wsm[];
  const int taps = kh * kw, co8n = Cout_p >> 3;
  for (int i = threadIdx.x; i < taps * CIN * Cout_p; i += blockDim.x) {
    const int co = i % Cout_p, ci = (i / Cout_p) % CIN, tap = i / (Cout_p * CIN);
    wsm[i] = (float)w[(size_t)co * w_ld + tap * 8 + ci];
  }
  __syncthreads();
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Wg = (Wo + PX - 1) / PX;
  const int co8 = (int)(idx % (uint32_t)co8n);
  uint32_t t = idx / (uint32_t)co8n;
  const int wg = (int)(t % (uint32_t)Wg);
  t /= (uint32_t)Wg;
  const int ho = (int)(t % (uint32_t)Ho);
  const int n = (int)(t / (uint32_t)Ho);
  const int wo0 = wg * PX;
  float acc[PX][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + co8 * 2), b1 = __ldg(reinterpret_cast<const float4*>(bias) + co8 * 2 + 1);
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      acc[p][0] = b0.x; acc[p][1] = b0.y; acc[p][2] = b0.z; acc[p][3] = b0.w; acc[p][4] = b1.x; acc[p][5] = b1.y; acc[p][6] = b1.z; acc[p][7] = b1.w;
    }
  }
  const uint4* xi = x + (size_t)n * H * W;
  for (int r = 0; r < kh; ++r) {
    const int hi = ho * sh - pt + r * dh;
    if ((unsigned)hi >= (unsigned)H) continue;
    for (int s = 0; s < kw; ++s) {
      float v[PX][CIN];
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        const int wi = (wo0 + p) * sw - pl + s * dw;
        const bool in = (unsigned)wi < (unsigned)W;
        const uint4 q = in ? __ldg(xi + (size_t)hi * W + wi) : make_uint4(0u, 0u, 0u, 0u);
        float a[4];
        unpack4<T>(make_uint2(q.x, q.y), a);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) v[p][ci] = a[ci];
      }
      const float* wt = wsm + (size_t)(r * kw + s) * CIN * Cout_p + co8 * 8;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float4 w0 = *reinterpret_cast<const float4*>(wt + ci * Cout_p), w1 = *reinterpret_cast<const float4*>(wt + ci * Cout_p + 4);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          acc[p][0] = fmaf(v[p][ci], w0.x, acc[p][0]); acc[p][1] = fmaf(v[p][ci], w0.y, acc[p][1]); acc[p][2] = fmaf(v[p][ci], w0.z, acc[p][2]);
          acc[p][3] = fmaf(v[p][ci], w0.w, acc[p][3]); acc[p][4] = fmaf(v[p][ci], w1.x, acc[p][4]); acc[p][5] = fmaf(v[p][ci], w1.y, acc[p][5]);
          acc[p][6] = fmaf(v[p][ci], w1.z, acc[p][6]); acc[p][7] = fmaf(v[p][ci], w1.w, acc[p][7]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    if (wo0 + p >= Wo) break;
    float lo[4], hi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { lo[e] = act_apply(acc[p][e], act, aa, ab); hi[e] = act_apply(acc[p][4 + e], act, aa, ab); }
    const uint2 a = pack4<T>(lo), b = pack4<T>(hi);
    y[(((size_t)n * Ho + ho) * Wo + wo0 + p) * ldy8 + co8] = make_uint4(a.x, a.y, b.x, b.y);
  }
}

}  // namespace

bool dwconv_strip_supported(const DwParams& p, int dtype) {
  const char* off = getenv("B200_DW_STRIP");
  if (off && off[0] == '0') return false;
  return (dtype == B200_BF16 || dtype == B200_F16) && p.kh == 3 && p.kw == 3 && p.dh == 1 && p.dw == 1 && p.sh == p.sw && (p.sh == 1 || p.sh == 2) &&
         p.pt >= 0 && p.pt <= 1 && p.pl >= 0 && p.pl <= 1 && p.Cp % 8 == 0;
}

void launch_dwconv_strip(const DwParams& p, int dtype, cudaStream_t s) {
  const int cq = p.Cp / 4, Wo2 = (p.Wo + 1) / 2;
  {
    // stride 1 with an activation that is a clamp (none / ReLU / Clamp - every MobileNet layer): the rolling-window kernel
    const char* off = getenv("B200_DW_ROLL");
    const bool clampish = p.act == ACT_NONE || p.act == ACT_RELU || p.act == ACT_CLAMP;
    if (p.sh == 1 && clampish && !(off && off[0] == '0')) {
      const float lo = p.act == ACT_NONE ? -3.4e38f : (p.act == ACT_RELU ? 0.f : p.a), hi = p.act == ACT_CLAMP ? p.b : 3.4e38f;
      const int R = p.Ho >= 16 ? 8 : (p.Ho >= 8 ? 4 : 2);
      const int strips = (p.Ho + R - 1) / R;
      const int64_t total = (int64_t)p.N * strips * Wo2 * cq;
      B200_CHECK(total > 0 && total < ((int64_t)1 << 31) && (int64_t)p.N * p.H * p.W * cq < ((int64_t)1 << 40), B200_ERR_UNSUPPORTED, "depthwise convolution too large");
      const unsigned grid = (unsigned)((total + 255) / 256);
#define B200_DWR(T, R_)                                                                                                                        \
  dwconv3x3_roll_kernel<T, R_><<<grid, 256, 0, s>>>((const uint2*)p.x, (const uint2*)p.w, p.bias, (uint2*)p.y, p.H, p.W, cq, p.Ho, p.Wo, strips, p.pt, p.pl, \
                                                    lo, hi, (uint32_t)total)
      if (dtype == B200_BF16) { if (R == 8) B200_DWR(__nv_bfloat16, 8); else if (R == 4) B200_DWR(__nv_bfloat16, 4); else B200_DWR(__nv_bfloat16, 2); }
      else { if (R == 8) B200_DWR(__half, 8); else if (R == 4) B200_DWR(__half, 4); else B200_DWR(__half, 2); }
#undef B200_DWR
      return;
    }
  }
  // rows per strip: 4 at stride 1 (6 x 4 vectors in flight per thread), 2 at stride 2 (5 x 5); short maps keep the strip count >= 2
  const int R = p.sh == 1 ? (p.Ho >= 8 ? 4 : 2) : 2;
  const int strips = (p.Ho + R - 1) / R;
  const int64_t total = (int64_t)p.N * strips * Wo2 * cq;
  B200_CHECK(total > 0 && total < ((int64_t)1 << 31) && (int64_t)p.N * p.H * p.W * cq < ((int64_t)1 << 40), B200_ERR_UNSUPPORTED, "depthwise convolution too large");
  const unsigned grid = (unsigned)((total + 255) / 256);
#define B200_DWS(T, S_, R_)                                                                                                                          \
  dwconv3x3_strip_kernel<T, S_, R_><<<grid, 256, 0, s>>>((const uint2*)p.x, (const uint2*)p.w, p.bias, (uint2*)p.y, p.H, p.W, cq, p.Ho, p.Wo, strips, p.pt, \
                                                         p.pl, p.act, p.a, p.b, (uint32_t)total)
  if (dtype == B200_BF16) {
    if (p.sh == 1) { if (R == 4) B200_DWS(__nv_bfloat16, 1, 4); else B200_DWS(__nv_bfloat16, 1, 2); }
    else B200_DWS(__nv_bfloat16, 2, 2);
  } else {
    if (p.sh == 1) { if (R == 4) B200_DWS(__half, 1, 4); else B200_DWS(__half, 1, 2); }
    else B200_DWS(__half, 2, 2);
  }
#undef B200_DWS
}

bool conv_small_cin_supported(const ConvParams& p, int cin, int dtype) {
  const char* off = getenv("B200_SMALL_CIN");
  if (off && off[0] == '0') return false;
  const int64_t smem = (int64_t)p.kh * p.kw * cin * p.Cout_p * 4;
  return (dtype == B200_BF16 || dtype == B200_F16) && p.Cin_p == 8 && cin >= 1 && cin <= 4 && p.Cout_p % 8 == 0 && p.Cout_p <= 128 && smem <= 40 * 1024 &&
         p.K2 == 0 && p.N2 == 0 && p.batch == 1 && !p.res && p.act2 == 0 && p.x_sw == 0 && p.ldy % 8 == 0 &&
         (int64_t)p.M * (p.Cout_p / 8) < ((int64_t)1 << 31);
}

void launch_conv_small_cin(const ConvParams& p, int cin, int dtype, cudaStream_t s) {
  B200_CHECK(conv_small_cin_supported(p, cin, dtype), B200_ERR_UNSUPPORTED, "conv_small_cin: unsupported launch");
  constexpr int PX = 4;
  const int64_t total = (int64_t)p.N * p.Ho * ((p.Wo + PX - 1) / PX) * (p.Cout_p / 8);
  const unsigned grid = (unsigned)((total + 255) / 256);
  const size_t smem = (size_t)p.kh * p.kw * cin * p.Cout_p * 4;
#define B200_SC(T, C_)                                                                                                                             \
  conv_small_cin_kernel<T, C_, PX><<<grid, 256, smem, s>>>((const uint4*)p.x, (const T*)p.w, p.bias, (uint4*)p.y, p.N, p.H, p.W, p.Ho, p.Wo, p.Cout_p, p.ldy / 8, \
                                                       p.kh, p.kw, p.sh, p.sw, p.dh, p.dw, p.pt, p.pl, p.w_ld, p.act1, p.a1, p.b1, (uint32_t)total)
#define B200_SC_T(T)                                                                     \
  switch (cin) {                                                                         \
    case 1: B200_SC(T, 1); break;                                                        \
    case 2: B200_SC(T, 2); break;                                                        \
    case 3: B200_SC(T, 3); break;                                                        \
    default: B200_SC(T, 4); break;                                                       \
  }
  if (dtype == B200_BF16) { B200_SC_T(__nv_bfloat16) } else { B200_SC_T(__half) }
#undef B200_SC_T
#undef B200_SC
}

}  // namespace b200
