// Memory-bound CUDA-core kernels: depthwise convolution, pooling, elementwise, softmax, MVN,
// gather, reductions, layout/dtype conversion, permute, TopK.  All activation tensors of rank 4
// are NHWC with the channel count padded to a multiple of 8 ("NHWC8") so that every access is a
// 128-bit vector; everything else is dense row-major.  Reference semantics are cited per kernel.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cfloat>

#include "act.cuh"
#include "common.h"
#include "kernels.h"

namespace b200 {
namespace {

// ---- typed 8-wide vector access ---------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p, int64_t i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p, int64_t i) { return __half2float(p[i]); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
template <> __device__ __forceinline__ float ldf<int32_t>(const int32_t* p, int64_t i) { return (float)p[i]; }
template <> __device__ __forceinline__ float ldf<uint8_t>(const uint8_t* p, int64_t i) { return (float)p[i]; }
template <typename T> __device__ __forceinline__ void stf(T* p, int64_t i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, int64_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, int64_t i, float v) { p[i] = __float2half_rn(v); }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void stf<int32_t>(int32_t* p, int64_t i, float v) { p[i] = (int32_t)v; }
template <> __device__ __forceinline__ void stf<uint8_t>(uint8_t* p, int64_t i, float v) { p[i] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f); }

template <typename T> __device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__half>(const __half* p, float (&f)[8]) {
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
// eight 16-bit values already in registers -> floats
template <typename T> __device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]);
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&f)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void store8<__half>(__half* p, const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

#define B200_DISPATCH_FLOAT(dtype, T, ...)                                             \
  do {                                                                                 \
    if ((dtype) == B200_F32) { using T = float; __VA_ARGS__; }                         \
    else if ((dtype) == B200_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }           \
    else if ((dtype) == B200_F16) { using T = __half; __VA_ARGS__; }                   \
    else B200_THROW(B200_ERR_UNSUPPORTED, "kernel: unsupported dtype " << (dtype));    \
  } while (0)

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide reductions (blockDim.x <= 1024), result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = l < nw ? sh[l] : 0.f;
  return warp_sum(r);
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = l < nw ? sh[l] : -FLT_MAX;
  return warp_max(r);
}

// ---- depthwise convolution ---------------------------------------------------------------------
// GroupConvolution with G == C (nodes/mkldnn_conv_node.cpp:78-100); one thread = one output pixel
// x 8 channels; weights [tap][Cp]; bias + activation fused.
template <typename T>
__global__ void dwconv_kernel(const DwParams p) {
  const int cpc = p.Cp >> 3;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)p.N * p.Ho * p.Wo * cpc;
  if (idx >= total) return;
  int c8 = (int)(idx % cpc);
  int64_t pix = idx / cpc;
  int wo = (int)(pix % p.Wo);
  int64_t t = pix / p.Wo;
  int ho = (int)(t % p.Ho);
  int n = (int)(t / p.Ho);
  const T* x = reinterpret_cast<const T*>(p.x) + (int64_t)n * p.H * p.W * p.Cp + c8 * 8;
  const T* w = reinterpret_cast<const T*>(p.w) + c8 * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = __ldg(p.bias + c8 * 8 + e);
  for (int r = 0; r < p.kh; ++r) {
    int hi = ho * p.sh - p.pt + r * p.dh;
    if ((unsigned)hi >= (unsigned)p.H) continue;
    for (int s = 0; s < p.kw; ++s) {
      int wi = wo * p.sw - p.pl + s * p.dw;
      if ((unsigned)wi >= (unsigned)p.W) continue;
      float xv[8], wv[8];
      load8<T>(x + ((int64_t)hi * p.W + wi) * p.Cp, xv);
      load8<T>(w + (int64_t)(r * p.kw + s) * p.Cp, wv);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(xv[e], wv[e], acc[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = act_apply(acc[e], p.act, p.a, p.b);
  store8<T>(reinterpret_cast<T*>(p.y) + pix * p.Cp + c8 * 8, acc);
}


// Fast path of the depthwise kernel for MobileNet-style layers: 3x3, dilation 1, pads <= 1, stride 1 or 2, 16-bit NHWC8 data.
// One thread = two horizontally adjacent output pixels x 8 channels: the 3 x (2 + S) input vectors are loaded once (all loads issued
// before the first use, out-of-image taps read as zero), the nine weight vectors once, fp32 accumulate, 32-bit index math.
template <typename T, int S>
__global__ void __launch_bounds__(256) dwconv3x3_16b_kernel(const uint4* __restrict__ x, const uint4* __restrict__ w, const float* __restrict__ bias,
                                                            uint4* __restrict__ y, int N, int H, int W, int cpc, int Ho, int Wo, int pt, int pl, int act,
                                                            float aa, float ab, uint32_t total) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Wo2 = (Wo + 1) >> 1;
  const uint32_t c8 = idx % (uint32_t)cpc;
  uint32_t t = idx / (uint32_t)cpc;
  const int j = (int)(t % (uint32_t)Wo2);
  t /= (uint32_t)Wo2;
  const int ho = (int)(t % (uint32_t)Ho);
  const int n = (int)(t / (uint32_t)Ho);
  constexpr int NC = 2 * S + 1 + (S == 1 ? 1 : 0);      // input columns feeding two outputs: 4 (stride 1) or 5 (stride 2)
  const uint4* xi = x + (size_t)n * H * W * cpc + c8;
  const int c0 = 2 * j * S - pl, r0 = ho * S - pt;
  uint4 v[3][NC];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hi = r0 + r;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int wi = c0 + c;
      const bool in = (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
      v[r][c] = in ? __ldg(xi + ((size_t)hi * W + wi) * cpc) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  float acc0[8], acc1[8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + c8 * 2), b1 = __ldg(reinterpret_cast<const float4*>(bias) + c8 * 2 + 1);
    acc0[0] = acc1[0] = b0.x; acc0[1] = acc1[1] = b0.y; acc0[2] = acc1[2] = b0.z; acc0[3] = acc1[3] = b0.w;
    acc0[4] = acc1[4] = b1.x; acc0[5] = acc1[5] = b1.y; acc0[6] = acc1[6] = b1.z; acc0[7] = acc1[7] = b1.w;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      float wv[8], a[8], b[8];
      const uint4 wq = __ldg(w + (size_t)(r * 3 + s) * cpc + c8);
      unpack8<T>(wq, wv);
      unpack8<T>(v[r][s], a);
      unpack8<T>(v[r][s + S], b);
#pragma unroll
      for (int e = 0; e < 8; ++e) { acc0[e] = fmaf(a[e], wv[e], acc0[e]); acc1[e] = fmaf(b[e], wv[e], acc1[e]); }
    }
#pragma unroll
  for (int e = 0; e < 8; ++e) { acc0[e] = act_apply(acc0[e], act, aa, ab); acc1[e] = act_apply(acc1[e], act, aa, ab); }
  T* yo = reinterpret_cast<T*>(y + (((size_t)n * Ho + ho) * Wo + 2 * j) * cpc + c8);
  store8<T>(yo, acc0);
  if (2 * j + 1 < Wo) store8<T>(yo + (size_t)cpc * 8, acc1);
}

// ---- pooling ------------------------------------------------------------------------------------
// MaxPool ignores padding; AvgPool divisor = kh*kw iff !exclude_pad and some pad != 0 - the WHOLE window, also where ceil rounding
// pushed it past the padded extent - else the number of in-bounds elements (reference/max_pool.hpp:74-117,
// reference/avg_pool.hpp:215-262: "the size of window doesn't change even if the window was clipped"; MKLDNNPlugin's jit pooling agrees.
// Its single-channel path divides by the clipped window instead: found by the randomized differential test, seed 43).
template <typename T>
__global__ void pool_kernel(const PoolParams p) {
  const int cpc = p.Cp >> 3;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)p.N * p.Ho * p.Wo * cpc;
  if (idx >= total) return;
  int c8 = (int)(idx % cpc);
  int64_t pix = idx / cpc;
  int wo = (int)(pix % p.Wo);
  int64_t t = pix / p.Wo;
  int ho = (int)(t % p.Ho);
  int n = (int)(t / p.Ho);
  const T* x = reinterpret_cast<const T*>(p.x) + (int64_t)n * p.H * p.W * p.Cp + c8 * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = p.is_avg ? 0.f : -FLT_MAX;
  int h0 = ho * p.sh - p.pt, w0 = wo * p.sw - p.pl;
  int cnt_in = 0;
  for (int r = 0; r < p.kh; ++r) {
    int hi = h0 + r;
    bool h_in = (unsigned)hi < (unsigned)p.H;
    for (int s = 0; s < p.kw; ++s) {
      int wi = w0 + s;
      bool w_in = (unsigned)wi < (unsigned)p.W;
      if (!(h_in && w_in)) continue;
      ++cnt_in;
      float xv[8];
      load8<T>(x + ((int64_t)hi * p.W + wi) * p.Cp, xv);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = p.is_avg ? acc[e] + xv[e] : fmaxf(acc[e], xv[e]);
    }
  }
  if (p.is_avg) {
    bool any_pad = (p.pt | p.pl | p.pb | p.pr) != 0;
    int div = (!p.exclude_pad && any_pad) ? p.kh * p.kw : cnt_in;
    float inv = div > 0 ? 1.f / (float)div : 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= inv;
  } else if (cnt_in == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  }
  store8<T>(reinterpret_cast<T*>(p.y) + pix * p.Cp + c8 * 8, acc);
}


// Fast path for the ResNet / SSD stem pool: 3x3 / stride-2 MaxPool over 16-bit NHWC8 data.  One thread = two horizontally adjacent output
// pixels x 8 channels: 15 unconditional 16-byte loads (3 rows x 5 columns, all issued before the first use) instead of 18 guarded ones.
// Out-of-image taps are clamped to the nearest in-image coordinate, which is always inside the same window (pads <= 1), so the maximum is
// unchanged (MaxPool ignores padding: reference/max_pool.hpp:74-117); max over packed pairs is exact.
template <typename T2>
__device__ __forceinline__ uint4 max8_16b(uint4 a, uint4 b) {
  uint4 r;
  T2 x, y, z;
  x = *reinterpret_cast<T2*>(&a.x); y = *reinterpret_cast<T2*>(&b.x); z = __hmax2(x, y); r.x = *reinterpret_cast<uint32_t*>(&z);
  x = *reinterpret_cast<T2*>(&a.y); y = *reinterpret_cast<T2*>(&b.y); z = __hmax2(x, y); r.y = *reinterpret_cast<uint32_t*>(&z);
  x = *reinterpret_cast<T2*>(&a.z); y = *reinterpret_cast<T2*>(&b.z); z = __hmax2(x, y); r.z = *reinterpret_cast<uint32_t*>(&z);
  x = *reinterpret_cast<T2*>(&a.w); y = *reinterpret_cast<T2*>(&b.w); z = __hmax2(x, y); r.w = *reinterpret_cast<uint32_t*>(&z);
  return r;
}
template <typename T2>
__global__ void __launch_bounds__(256) maxpool3x3s2_16b_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int cpc, int Ho,
                                                               int Wo, int pt, int pl, uint32_t total) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Wo2 = (Wo + 1) >> 1;
  const uint32_t c8 = idx % (uint32_t)cpc;
  uint32_t t = idx / (uint32_t)cpc;
  const int j = (int)(t % (uint32_t)Wo2);
  t /= (uint32_t)Wo2;
  const int ho = (int)(t % (uint32_t)Ho);
  const int n = (int)(t / (uint32_t)Ho);
  const uint4* xi = x + (size_t)n * H * W * cpc + c8;
  const int c0 = 4 * j - pl, r0 = 2 * ho - pt;
  uint4 v[3][5];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hi = min(max(r0 + r, 0), H - 1);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const int wi = min(max(c0 + c, 0), W - 1);
      v[r][c] = __ldg(xi + ((size_t)hi * W + wi) * cpc);
    }
  }
  uint4 col[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) col[c] = max8_16b<T2>(max8_16b<T2>(v[0][c], v[1][c]), v[2][c]);
  uint4* yo = y + (((size_t)n * Ho + ho) * Wo + 2 * j) * cpc + c8;
  yo[0] = max8_16b<T2>(max8_16b<T2>(col[0], col[1]), col[2]);
  if (2 * j + 1 < Wo) yo[cpc] = max8_16b<T2>(max8_16b<T2>(col[2], col[3]), col[4]);
}

// Global average pool, one warp per (image, 8 channels): the lanes stride over the pixels (16-byte loads), warp-shuffle sums.  The
// thread-per-(image, 8 channels) form below walks the 49 pixels of ResNet's last stage serially: 14 us at batch 1.
template <typename T>
__global__ void __launch_bounds__(256) global_avgpool_warp_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int HW, int Cp) {
  const int cpc = Cp >> 3;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)N * cpc) return;
  const int lane = threadIdx.x & 31;
  const int c8 = (int)(w % cpc);
  const int n = (int)(w / cpc);
  const T* px = x + (int64_t)n * HW * Cp + c8 * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = lane; i < HW; i += 32) {
    float v[8];
    load8<T>(px + (int64_t)i * Cp, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += v[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = warp_sum(acc[e]);
  if (lane == 0) {
    const float inv = 1.f / (float)HW;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= inv;
    store8<T>(y + (int64_t)n * Cp + c8 * 8, acc);
  }
}

// Large batches (ResNet-50 b256: 256 x 49 x 2048): one thread per (image, 8 channels), 128 consecutive channel groups per block, so every
// pixel row is read as one contiguous 2 KB segment per block and each thread keeps U 16-byte loads in flight.  The warp-per-(image, 8 channels)
// kernel above strides its lanes over pixels Cp elements apart (one 16-byte piece per 4 KB): 32 us for 51 MB at b256 against ~10 us here.
template <typename T, int U>
__global__ void __launch_bounds__(128) global_avgpool_cols_kernel(const T* __restrict__ x, T* __restrict__ y, int HW, int Cp) {
  const int cpc = Cp >> 3;
  const int c8 = blockIdx.x * 128 + threadIdx.x;
  if (c8 >= cpc) return;
  const int n = blockIdx.y;
  const T* px = x + (int64_t)n * HW * Cp + c8 * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int i = 0;
  for (; i + U <= HW; i += U) {
    float v[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) load8<T>(px + (int64_t)(i + u) * Cp, v[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[u][e];
    }
  }
  for (; i < HW; ++i) {
    float v[8];
    load8<T>(px + (int64_t)i * Cp, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += v[e];
  }
  const float inv = 1.f / (float)HW;
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] *= inv;
  store8<T>(y + (int64_t)n * Cp + c8 * 8, acc);
}

template <typename T>
__global__ void global_avgpool_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int HW, int Cp) {
  const int cpc = Cp >> 3;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * cpc) return;
  int c8 = (int)(idx % cpc);
  int n = (int)(idx / cpc);
  const T* px = x + (int64_t)n * HW * Cp + c8 * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < HW; ++i) {
    float v[8];
    load8<T>(px + (int64_t)i * Cp, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += v[e];
  }
  float inv = 1.f / (float)HW;
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] *= inv;
  store8<T>(y + (int64_t)n * Cp + c8 * 8, acc);
}

// ---- elementwise ---------------------------------------------------------------------------------
__device__ __forceinline__ float binop(float a, float b, int op) {
  switch (op) {
    case B200_OP_ADD: return a + b;
    case B200_OP_SUBTRACT: return a - b;
    case B200_OP_MULTIPLY: return a * b;
    case B200_OP_DIVIDE: return a / b;
    case B200_OP_MAXIMUM: return fmaxf(a, b);
    case B200_OP_MINIMUM: return fminf(a, b);
    case B200_OP_POWER: return powf(a, b);
    case B200_OP_SQUARED_DIFFERENCE: return (a - b) * (a - b);
    case B200_OP_PRELU: return a >= 0.f ? a : a * b;
  }
  return a;
}

template <typename T>
__global__ void unary_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, int act, float a, float b) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float v[8];
    load8<T>(x + i, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = act_apply(v[e], act, a, b);
    store8<T>(y + i, v);
  } else {
    for (; i < n; ++i) stf<T>(y, i, act_apply(ldf<T>(x, i), act, a, b));
  }
}

// FakeQuantize, element by element as the reference's quantize() (fake_quantize.hpp:176-194): below / at the low bound -> out_low, above the
// high bound -> out_high, else nearbyint((x - il) / (ih - il) * (levels - 1)) / (levels - 1) * (oh - ol) + ol, in fp32.
template <typename T>
__global__ void fake_quantize_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, const float* __restrict__ ranges, int cdim, int64_t inner,
                                     float lm1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cdim == 1 ? 0 : (int)((i / inner) % cdim);
  const float il = ranges[c], ih = ranges[cdim + c], ol = ranges[2 * cdim + c], oh = ranges[3 * cdim + c];
  const float v = ldf<T>(x, i);
  float r;
  if (v <= fminf(il, ih)) r = ol;
  else if (v > fmaxf(il, ih)) r = oh;
  else r = rintf((v - il) / (ih - il) * lm1) / lm1 * (oh - ol) + ol;
  stf<T>(y, i, r);
}

template <typename T>
__global__ void binary_same_kernel(const T* __restrict__ pa, const T* __restrict__ pb, T* __restrict__ y, int64_t n, int op, int act,
                                   float aa, float ab) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float u[8], v[8];
    load8<T>(pa + i, u);
    load8<T>(pb + i, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) u[e] = act_apply(binop(u[e], v[e], op), act, aa, ab);
    store8<T>(y + i, u);
  } else {
    for (; i < n; ++i) stf<T>(y, i, act_apply(binop(ldf<T>(pa, i), ldf<T>(pb, i), op), act, aa, ab));
  }
}

template <typename T>
__global__ void binary_bcast_kernel(const BinaryParams p) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  int64_t rem = i, ao = 0, bo = 0;
#pragma unroll 1
  for (int d = p.rank - 1; d >= 0; --d) {
    int64_t c = rem % p.oshape[d];
    rem /= p.oshape[d];
    ao += c * p.astride[d];
    bo += c * p.bstride[d];
  }
  float r = binop(ldf<T>(reinterpret_cast<const T*>(p.a), ao), ldf<T>(reinterpret_cast<const T*>(p.b), bo), p.op);
  stf<T>(reinterpret_cast<T*>(p.y), i, act_apply(r, p.act, p.aa, p.ab));
}

template <typename T>
__global__ void scale_shift_kernel(const T* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                   T* __restrict__ y, int64_t rows, int Cp, int act, float aa, float ab) {
  const int cpc = Cp >> 3;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cpc) return;
  int c8 = (int)(idx % cpc);
  float v[8];
  load8<T>(x + idx * 8, v);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = act_apply(fmaf(v[e], __ldg(scale + c8 * 8 + e), __ldg(shift + c8 * 8 + e)), act, aa, ab);
  store8<T>(y + idx * 8, v);
}

// ---- softmax: exp(x - max) / sum (reference/softmax.hpp:19-47) ----------------------------------
// y[row, :] = softmax(x[row, :] * scale + mask[row / rows_per_mask, :]) over rows of `cols` (<= 2048, multiple of 8) elements: one warp per
// row, the row lives in registers between the passes, fp32 math (reference/softmax.hpp:19-47 after the Multiply / Add of the attention
// pattern).  mask == nullptr: plain (scaled) softmax.  Reads and writes each element once: the HBM-bound form of BERT's
// scale -> mask-add -> softmax chain (three kernels, five passes over the [N, heads, S, S] scores otherwise).
// 2^x on the MUFU pipe (2 ulp): the row kernels below work in the base-2 domain (inputs pre-multiplied by log2 e), one instruction per
// exponential instead of expf's range-reduced sequence -- the one-row kernel was ISSUE-bound at 35 instructions per element with expf
// (ncu: 82% issue slots busy at 3.4 TB/s, profiles/r1_softmax_ncu.txt).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;

template <typename T, int CHUNKS>
__global__ void __launch_bounds__(256) softmax_warp_rows_kernel(const T* __restrict__ x, const T* __restrict__ mask, T* __restrict__ y, int64_t rows,
                                                                int cols, float scale, int64_t rows_per_mask) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* px = x + row * cols;
  const T* pm = mask ? mask + (row / rows_per_mask) * cols : nullptr;
  float v[CHUNKS][8];
  float m = -FLT_MAX;
  const float sc2 = scale * kLog2e;                 // base-2 domain: exp(a - m) = 2^((a - m) log2 e)
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    const int c = (k * 32 + lane) * 8;
    if (c < cols) {
      load8<T>(px + c, v[k]);
      if (pm) {
        float mk[8];
        load8<T>(pm + c, mk);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[k][e] = fmaf(v[k][e], sc2, mk[e] * kLog2e);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[k][e] *= sc2;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) m = fmaxf(m, v[k][e]);
    }
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    const int c = (k * 32 + lane) * 8;
    if (c < cols) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { v[k][e] = ex2_approx(v[k][e] - m); sum += v[k][e]; }
    }
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  T* py = y + row * cols;
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    const int c = (k * 32 + lane) * 8;
    if (c < cols) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[k][e] *= inv;
      store8<T>(py + c, v[k]);
    }
  }
}

// Exact-width variant for 16-bit rows of CHUNKS x 32 x 2W elements (384 = 3 x 32 x 4, 128, 256, 512, 768, 1024): every lane owns the same
// number of elements (the 8-per-lane layout above leaves half the warp idle on the second chunk of a 384-wide row while still issuing its
// instructions), no column guards, 32-bit row arithmetic (the 64-bit row / rows_per_mask division alone was ~80 instructions per thread).
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
template <int W> struct WordVec;
template <> struct WordVec<2> { using type = uint2; };
template <> struct WordVec<4> { using type = uint4; };

template <typename T, int W, int CHUNKS>
__global__ void __launch_bounds__(256) softmax_rows_exact_kernel(const T* __restrict__ x, const T* __restrict__ mask, T* __restrict__ y, uint32_t rows,
                                                                 float scale, uint32_t rows_per_mask) {
  using Vec = typename WordVec<W>::type;
  constexpr int COLS = CHUNKS * 32 * 2 * W;
  const uint32_t row = blockIdx.x * 8u + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const Vec* px = reinterpret_cast<const Vec*>(x + (size_t)row * COLS) + lane;
  union U { Vec v; uint32_t w[W]; };
  U raw[CHUNKS], mk[CHUNKS];
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) raw[k].v = __ldg(px + k * 32);
  if (mask) {
    const Vec* pm = reinterpret_cast<const Vec*>(mask + (size_t)(row / rows_per_mask) * COLS) + lane;
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) mk[k].v = __ldg(pm + k * 32);
  }
  const float sc2 = scale * kLog2e;
  float v[CHUNKS][2 * W];
  float m = -FLT_MAX;
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const float2 a = unpack2<T>(raw[k].w[i]);
      float2 b = make_float2(0.f, 0.f);
      if (mask) b = unpack2<T>(mk[k].w[i]);
      v[k][2 * i] = fmaf(a.x, sc2, b.x * kLog2e);
      v[k][2 * i + 1] = fmaf(a.y, sc2, b.y * kLog2e);
      m = fmaxf(m, fmaxf(v[k][2 * i], v[k][2 * i + 1]));
    }
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
#pragma unroll
    for (int e = 0; e < 2 * W; ++e) { v[k][e] = ex2_approx(v[k][e] - m); sum += v[k][e]; }
  }
  const float inv = 1.f / warp_sum(sum);
  Vec* py = reinterpret_cast<Vec*>(y + (size_t)row * COLS) + lane;
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    U o;
#pragma unroll
    for (int i = 0; i < W; ++i) o.w[i] = pack2<T>(v[k][2 * i] * inv, v[k][2 * i + 1] * inv);
    py[k * 32] = o.v;
  }
}

template <typename T>
__global__ void softmax_rows_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int64_t cols) {
  __shared__ float sh[32];
  int64_t row = blockIdx.x;
  const T* px = x + row * cols;
  T* py = y + row * cols;
  float m = -FLT_MAX;
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, ldf<T>(px, c));
  m = block_max(m, sh);
  float s = 0.f;
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) s += expf(ldf<T>(px, c) - m);
  s = block_sum(s, sh);
  float inv = 1.f / s;
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) stf<T>(py, c, expf(ldf<T>(px, c) - m) * inv);
}
// Short rows (SSD's per-prior class scores: 81 columns, 279 k rows at batch 32): one WARP per row, the row in registers (<= 4 values per
// lane), one read and one write per element.  The block-per-row kernel above launched one 32-thread block per row and read it three times.
template <typename T>
__global__ void __launch_bounds__(256) softmax_short_rows_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int cols) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* px = x + row * cols;
  T* py = y + row * cols;
  float v[4];
  float m = -FLT_MAX;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = lane + 32 * k;
    v[k] = c < cols ? ldf<T>(px, c) : -FLT_MAX;
    m = fmaxf(m, v[k]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = (lane + 32 * k) < cols ? expf(v[k] - m) : 0.f;
    sum += v[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = lane + 32 * k;
    if (c < cols) stf<T>(py, c, v[k] * inv);
  }
}

template <typename T>
__global__ void softmax_strided_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t outer, int64_t axis, int64_t inner) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= outer * inner) return;
  int64_t o = i / inner, in = i % inner;
  const T* px = x + o * axis * inner + in;
  T* py = y + o * axis * inner + in;
  float m = -FLT_MAX;
  for (int64_t a = 0; a < axis; ++a) m = fmaxf(m, ldf<T>(px, a * inner));
  float s = 0.f;
  for (int64_t a = 0; a < axis; ++a) s += expf(ldf<T>(px, a * inner) - m);
  float inv = 1.f / s;
  for (int64_t a = 0; a < axis; ++a) stf<T>(py, a * inner, expf(ldf<T>(px, a * inner) - m) * inv);
}

// ---- MVN / LayerNorm over the last `cols` elements (reference/mvn.hpp:66-118) --------------------
// MVN / LayerNorm over rows of `cols` (<= 2048, multiple of 8) elements: one warp per row, the row stays in registers between the three
// passes (reference/mvn.hpp:66-118: biased variance of the centred values; eps inside or outside the square root).
template <typename T, int CHUNKS>
__global__ void __launch_bounds__(256) mvn_warp_rows_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int64_t rows, int cols, int normalize_variance, int eps_outside,
                                                            float eps) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* px = x + row * cols;
  float v[CHUNKS][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    const int c = (k * 32 + lane) * 8;
    if (c < cols) {
      load8<T>(px + c, v[k]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[k][e];
    }
  }
  const float mean = warp_sum(s) / (float)cols;
  float scale = 1.f;
  if (normalize_variance) {
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < cols) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[k][e] - mean; q += d * d; }
      }
    }
    const float var = warp_sum(q) / (float)cols;
    scale = eps_outside ? 1.f / (sqrtf(var) + eps) : 1.f / sqrtf(var + eps);
  }
  T* py = y + row * cols;
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    const int c = (k * 32 + lane) * 8;
    if (c < cols) {
      float gm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, bt[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (gamma) load8<float>(gamma + c, gm);          // two 16-byte loads (c is a multiple of 8: 32-byte aligned)
      if (beta) load8<float>(beta + c, bt);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[k][e] = fmaf((v[k][e] - mean) * scale, gm[e], bt[e]);
      store8<T>(py + c, v[k]);
    }
  }
}

// A copy of a packed word the compiler cannot connect to its source: keeps the passes below from caching the unpacked fp32 values
// (which would triple the register footprint of the row).
__device__ __forceinline__ uint32_t opaque(uint32_t u) {
  asm volatile("" : "+r"(u));
  return u;
}

// Exact-width variant for 16-bit rows of CHUNKS x 256 elements (256/512/768/1024: BERT's hidden size): the row stays PACKED in registers
// (4 registers per 8 elements, unpacked again in each of the three passes) so that the kernel fits 32 registers and all 64 warps of an SM
// are resident -- one warp-row is a load -> reduce -> reduce -> store latency chain, and the bytes in flight per SM are what bound it.
template <typename T, int CHUNKS>
__global__ void __launch_bounds__(256, 8) mvn_rows_exact_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, uint32_t rows, int normalize_variance, int eps_outside,
                                                                float eps) {
  constexpr int COLS = CHUNKS * 256;
  const uint32_t row = blockIdx.x * 8u + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const uint4* px = reinterpret_cast<const uint4*>(x + (size_t)row * COLS) + lane;
  uint4 raw[CHUNKS];
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) raw[k] = __ldg(px + k * 32);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    const float2 a = unpack2<T>(raw[k].x), b = unpack2<T>(raw[k].y), c = unpack2<T>(raw[k].z), d = unpack2<T>(raw[k].w);
    s += ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (d.x + d.y));
  }
  const float mean = warp_sum(s) * (1.f / (float)COLS);
  float scale = 1.f;
  if (normalize_variance) {
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) {
      const uint32_t w[4] = {opaque(raw[k].x), opaque(raw[k].y), opaque(raw[k].z), opaque(raw[k].w)};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 a = unpack2<T>(w[i]);
        const float d0 = a.x - mean, d1 = a.y - mean;
        q = fmaf(d0, d0, q);
        q = fmaf(d1, d1, q);
      }
    }
    const float var = warp_sum(q) * (1.f / (float)COLS);
    scale = eps_outside ? 1.f / (sqrtf(var) + eps) : rsqrtf(var + eps);
  }
  uint4* py = reinterpret_cast<uint4*>(y + (size_t)row * COLS) + lane;
#pragma unroll
  for (int k = 0; k < CHUNKS; ++k) {
    const int c = (k * 32 + lane) * 8;
    uint4 o;
    {
      float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gamma) g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      if (beta) b = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float2 v0 = unpack2<T>(opaque(raw[k].x)), v1 = unpack2<T>(opaque(raw[k].y));
      o.x = pack2<T>(fmaf((v0.x - mean) * scale, g.x, b.x), fmaf((v0.y - mean) * scale, g.y, b.y));
      o.y = pack2<T>(fmaf((v1.x - mean) * scale, g.z, b.z), fmaf((v1.y - mean) * scale, g.w, b.w));
    }
    {
      float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gamma) g = __ldg(reinterpret_cast<const float4*>(gamma + c) + 1);
      if (beta) b = __ldg(reinterpret_cast<const float4*>(beta + c) + 1);
      const float2 v2 = unpack2<T>(opaque(raw[k].z)), v3 = unpack2<T>(opaque(raw[k].w));
      o.z = pack2<T>(fmaf((v2.x - mean) * scale, g.x, b.x), fmaf((v2.y - mean) * scale, g.y, b.y));
      o.w = pack2<T>(fmaf((v3.x - mean) * scale, g.z, b.z), fmaf((v3.y - mean) * scale, g.w, b.w));
    }
    py[k * 32] = o;
  }
}

template <typename T>
__global__ void mvn_rows_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int64_t cols, int normalize_variance, int eps_outside, float eps) {
  __shared__ float sh[32];
  int64_t row = blockIdx.x;
  const T* px = x + row * cols;
  T* py = y + row * cols;
  float s = 0.f;
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) s += ldf<T>(px, c);
  float mean = block_sum(s, sh) / (float)cols;
  float scale = 1.f;
  if (normalize_variance) {
    float v = 0.f;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) { float d = ldf<T>(px, c) - mean; v += d * d; }
    float var = block_sum(v, sh) / (float)cols;
    scale = eps_outside ? 1.f / (sqrtf(var) + eps) : 1.f / sqrtf(var + eps);
  }
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) {
    float r = (ldf<T>(px, c) - mean) * scale;
    if (gamma) r *= __ldg(gamma + c);
    if (beta) r += __ldg(beta + c);
    stf<T>(py, c, r);
  }
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ table, const int32_t* __restrict__ idx, T* __restrict__ y, int64_t n_idx,
                                   int64_t row_elems, int64_t table_rows) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_idx * row_elems) return;
  int64_t r = i / row_elems, c = i % row_elems;
  int64_t src = idx[r];
  if (src < 0) src += table_rows;
  y[i] = (src >= 0 && src < table_rows) ? table[src * row_elems + c] : T(0);
}

template <typename T>
__global__ void reduce_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t outer, int64_t axis, int64_t inner, int kind) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= outer * inner) return;
  int64_t o = i / inner, in = i % inner;
  const T* px = x + o * axis * inner + in;
  float acc = kind == 2 ? -FLT_MAX : 0.f;
  for (int64_t a = 0; a < axis; ++a) {
    float v = ldf<T>(px, a * inner);
    acc = kind == 2 ? fmaxf(acc, v) : acc + v;
  }
  if (kind == 0) acc /= (float)axis;
  stf<T>(y, i, acc);
}

// ---- layout + dtype conversion -------------------------------------------------------------------
// logical [N,C,H,W]; physical: ROWMAJOR = NCHW dense, NHWC8 = [N,H,W,round_up(C,8)], NHWC_DENSE = [N,H,W,C]
__device__ __forceinline__ int64_t phys_index(int layout, int n, int c, int h, int w, int C, int H, int W) {
  switch (layout) {
    case PL_ROWMAJOR: return (((int64_t)n * C + c) * H + h) * W + w;
    case PL_NHWC8: return (((int64_t)n * H + h) * W + w) * ((C + 7) & ~7) + c;
    default: return (((int64_t)n * H + h) * W + w) * C + c;
  }
}
// One thread per destination element of the padded extent; channel-fastest destinations read
// the NCHW source with stride H*W (L2 absorbs it: each 32-byte sector is reused by 8 neighbours
// in w, which are adjacent threads when C is small, e.g. the 3-channel network input).
template <typename S, typename D>
__global__ void convert4_kernel(const S* __restrict__ src, int slayout, D* __restrict__ dst, int dlayout, int N, int C, int H, int W) {
  const int Cd = dlayout == PL_NHWC8 ? ((C + 7) & ~7) : C;
  int64_t total = (int64_t)N * Cd * H * W;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int n, c, h, w;
  if (dlayout == PL_ROWMAJOR) {
    w = (int)(i % W); int64_t t = i / W; h = (int)(t % H); t /= H; c = (int)(t % C); n = (int)(t / C);
  } else {
    c = (int)(i % Cd); int64_t t = i / Cd; w = (int)(t % W); t /= W; h = (int)(t % H); n = (int)(t / H);
  }
  float v = c < C ? ldf<S>(src, phys_index(slayout, n, c, h, w, C, H, W)) : 0.f;
  stf<D>(dst, i, v);
}
// One thread per 16-channel block: gathers the 2x2 pixel patch (zero outside the image), converts, writes 32 bytes.
template <typename S, typename D, int C>
__global__ void __launch_bounds__(256) ingest_s2d_kernel(const S* __restrict__ src, int slayout, D* __restrict__ dst, int N, int H, int W, int pt, int pl,
                                                         int Hb, int Wb) {
  // one thread = one block pixel (2x2 input pixels x C channels -> 16 output channels, zero padded); 32-bit index math, C known at compile
  // time so that the 16 outputs stay in registers (the runtime-C version kept them in local memory and ran at 1/4 of the HBM rate)
  const uint32_t total = (uint32_t)N * (uint32_t)Hb * (uint32_t)Wb;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int wb = (int)(i % (uint32_t)Wb);
  const uint32_t t = i / (uint32_t)Wb;
  const int hb = (int)(t % (uint32_t)Hb), n = (int)(t / (uint32_t)Hb);
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.f;
#pragma unroll
  for (int dr = 0; dr < 2; ++dr)
#pragma unroll
    for (int ds = 0; ds < 2; ++ds) {
      const int h = 2 * hb + dr - pt, w = 2 * wb + ds - pl;
      if ((unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W) {
#pragma unroll
        for (int c = 0; c < C; ++c) v[(dr * 2 + ds) * C + c] = ldf<S>(src, phys_index(slayout, n, c, h, w, C, H, W));
      }
    }
  __align__(16) D out[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) stf<D>(out, k, v[k]);
  uint4* o = reinterpret_cast<uint4*>(dst + (size_t)i * 16);
  o[0] = reinterpret_cast<const uint4*>(out)[0];
  o[1] = reinterpret_cast<const uint4*>(out)[1];
}
template <typename S, typename D>
__global__ void convert_flat_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf<D>(dst, i, ldf<S>(src, i));
}
__global__ void copy_i32_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) permute_vec16_kernel(const PermuteParams p) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint32_t)p.n) return;
  uint32_t rem = i;
  int64_t so = 0;
#pragma unroll 1
  for (int d = p.rank - 1; d >= 0; --d) {
    const uint32_t od = (uint32_t)p.oshape[d];
    const uint32_t c = rem % od;
    rem /= od;
    so += (int64_t)c * p.istride[d];
  }
  reinterpret_cast<uint4*>(p.y)[i] = __ldg(reinterpret_cast<const uint4*>(p.x) + so);
}

// Batched transpose of the last two dimensions of a dense 16-bit tensor, [batch][B][A] -> [batch][A][B], through a 64 x 64 shared-memory
// tile (4-byte global accesses on both sides).  BERT's V^T for the attention-weighted sum went through the element-wise permute at 0.6 TB/s.
__global__ void __launch_bounds__(256) transpose_last2_16b_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int A, int B) {
  __shared__ uint16_t tile[64][66];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int a0 = blockIdx.x * 64, b0 = blockIdx.y * 64;
  const size_t base = (size_t)blockIdx.z * A * B;
  // input rows = b (B of them), columns = a (contiguous)
#pragma unroll
  for (int r = ty; r < 64; r += 8) {
    const int b = b0 + r, a = a0 + 2 * tx;
    uint32_t v = 0;
    if (b < B && a + 1 < A) v = __ldg(reinterpret_cast<const uint32_t*>(x + base + (size_t)b * A + a));
    else if (b < B && a < A) v = x[base + (size_t)b * A + a];
    tile[r][2 * tx] = (uint16_t)(v & 0xFFFFu);
    tile[r][2 * tx + 1] = (uint16_t)(v >> 16);
  }
  __syncthreads();
  // output rows = a, columns = b (contiguous)
#pragma unroll
  for (int r = ty; r < 64; r += 8) {
    const int a = a0 + r, b = b0 + 2 * tx;
    if (a < A && b + 1 < B) {
      const uint32_t v = (uint32_t)tile[2 * tx][r] | ((uint32_t)tile[2 * tx + 1][r] << 16);
      *reinterpret_cast<uint32_t*>(y + base + (size_t)a * B + b) = v;
    } else if (a < A && b < B) {
      y[base + (size_t)a * B + b] = tile[2 * tx][r];
    }
  }
}

// four consecutive 16-byte vectors of the innermost dimension per thread: one index decomposition per 64 bytes
__global__ void __launch_bounds__(256) permute_vec16x4_kernel(const PermuteParams p) {
  const uint32_t i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= (uint32_t)(p.n >> 2)) return;
  uint32_t rem = i4;
  const uint32_t inner4 = (uint32_t)p.oshape[p.rank - 1] >> 2;
  const uint32_t c_in = (rem % inner4) << 2;
  rem /= inner4;
  int64_t so = c_in;
#pragma unroll 1
  for (int d = p.rank - 2; d >= 0; --d) {
    const uint32_t od = (uint32_t)p.oshape[d];
    const uint32_t c = rem % od;
    rem /= od;
    so += (int64_t)c * p.istride[d];
  }
  const uint4* src = reinterpret_cast<const uint4*>(p.x) + so;
  const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
  uint4* dst = reinterpret_cast<uint4*>(p.y) + (size_t)i4 * 4;
  dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
}

template <typename E>
__global__ void permute_kernel(const PermuteParams p) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  int64_t rem = i, so = 0;
#pragma unroll 1
  for (int d = p.rank - 1; d >= 0; --d) {
    int64_t c = rem % p.oshape[d];
    rem /= p.oshape[d];
    so += c * p.istride[d];
  }
  reinterpret_cast<E*>(p.y)[i] = reinterpret_cast<const E*>(p.x)[so];
}

template <typename V>
__global__ void copy2d_kernel(const char* __restrict__ src, int64_t spitch, char* __restrict__ dst, int64_t dpitch, int64_t row_vecs,
                              int64_t rows) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * row_vecs) return;
  int64_t r = i / row_vecs, c = i % row_vecs;
  reinterpret_cast<V*>(dst + r * dpitch)[c] = reinterpret_cast<const V*>(src + r * spitch)[c];
}

// ---- TopK -----------------------------------------------------------------------------------------
// Order: MAX -> (value desc, index asc); MIN -> (value asc, index asc)
// (reference/topk.hpp:22-40 compare_max/compare_min; nodes/mkldnn_topk_node.cpp:405-463 strict compare).
// One block per (outer, inner) slice; k rounds of a block-wide arg-best over the elements that come
// strictly after the previously selected one in that total order.
__device__ __forceinline__ bool topk_before(float va, int ia, float vb, int ib, int mode_min) {
  if (va != vb) return mode_min ? va < vb : va > vb;
  return ia < ib;
}
__global__ void topk_kernel(const float* __restrict__ x, float* __restrict__ vals, int32_t* __restrict__ idx, int64_t axis, int64_t inner,
                            int k, int mode_min, int sort_mode) {
  extern __shared__ unsigned char topk_smem[];
  float* sv = reinterpret_cast<float*>(topk_smem);      // [nwarps]
  int* si = reinterpret_cast<int*>(sv + 32);            // [nwarps]
  float* outv = reinterpret_cast<float*>(si + 32);      // [k]
  int* outi = reinterpret_cast<int*>(outv + k);         // [k]
  __shared__ float last_v;
  __shared__ int last_i;
  int64_t slice = blockIdx.x;
  int64_t o = slice / inner, in = slice % inner;
  const float* px = x + o * axis * inner + in;
  const int nw = (blockDim.x + 31) >> 5;
  for (int r = 0; r < k; ++r) {
    float bv = 0.f;
    int bi = -1;
    for (int64_t a = threadIdx.x; a < axis; a += blockDim.x) {
      float v = px[a * inner];
      if (r > 0 && !topk_before(last_v, last_i, v, (int)a, mode_min)) continue;  // already selected or earlier
      if (bi < 0 || topk_before(v, (int)a, bv, bi, mode_min)) { bv = v; bi = (int)a; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (oi >= 0 && (bi < 0 || topk_before(ov, oi, bv, bi, mode_min))) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float fv = sv[0];
      int fi = si[0];
      for (int w = 1; w < nw; ++w)
        if (si[w] >= 0 && (fi < 0 || topk_before(sv[w], si[w], fv, fi, mode_min))) { fv = sv[w]; fi = si[w]; }
      last_v = fv; last_i = fi;
      outv[r] = fv; outi[r] = fi;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && sort_mode == 2) {  // SORT_INDICES: ascending index
    for (int a = 1; a < k; ++a) {
      float v = outv[a]; int id = outi[a]; int b = a - 1;
      while (b >= 0 && outi[b] > id) { outv[b + 1] = outv[b]; outi[b + 1] = outi[b]; --b; }
      outv[b + 1] = v; outi[b + 1] = id;
    }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < k; r += blockDim.x) {
    int64_t off = o * k * inner + (int64_t)r * inner + in;
    vals[off] = outv[r];
    idx[off] = outi[r];
  }
}

}  // namespace

// ================================= launchers ==========================================================
void launch_dwconv(const DwParams& p, int dtype, cudaStream_t s) {
  if (dwconv_strip_supported(p, dtype)) { launch_dwconv_strip(p, dtype, s); return; }
  int64_t total = (int64_t)p.N * p.Ho * p.Wo * (p.Cp / 8);
  const bool fast = (dtype == B200_BF16 || dtype == B200_F16) && p.kh == 3 && p.kw == 3 && p.dh == 1 && p.dw == 1 && p.sh == p.sw && (p.sh == 1 || p.sh == 2) &&
                    p.pt >= 0 && p.pt <= 1 && p.pl >= 0 && p.pl <= 1;
  const int64_t total2 = (int64_t)p.N * p.Ho * ((p.Wo + 1) / 2) * (p.Cp / 8);
  if (fast && total2 > 0 && total2 < ((int64_t)1 << 31)) {
    const unsigned grid = grid_for(total2, 256);
#define B200_DW_LAUNCH(T, S_) dwconv3x3_16b_kernel<T, S_><<<grid, 256, 0, s>>>((const uint4*)p.x, (const uint4*)p.w, p.bias, (uint4*)p.y, p.N, p.H, p.W, p.Cp / 8, \
                                                                               p.Ho, p.Wo, p.pt, p.pl, p.act, p.a, p.b, (uint32_t)total2)
    if (dtype == B200_BF16) { if (p.sh == 1) B200_DW_LAUNCH(__nv_bfloat16, 1); else B200_DW_LAUNCH(__nv_bfloat16, 2); }
    else { if (p.sh == 1) B200_DW_LAUNCH(__half, 1); else B200_DW_LAUNCH(__half, 2); }
    return;
  }
  B200_DISPATCH_FLOAT(dtype, T, (dwconv_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(p)));
}
void launch_pool(const PoolParams& p, int dtype, cudaStream_t s) {
  int64_t total = (int64_t)p.N * p.Ho * p.Wo * (p.Cp / 8);
  const bool fast = !p.is_avg && (dtype == B200_BF16 || dtype == B200_F16) && p.kh == 3 && p.kw == 3 && p.sh == 2 && p.sw == 2 && p.pt <= 1 && p.pl <= 1 &&
                    p.pb <= 1 && p.pr <= 1 && (p.Ho - 1) * 2 - p.pt <= p.H - 1 && (p.Wo - 1) * 2 - p.pl <= p.W - 1;
  const int64_t total2 = (int64_t)p.N * p.Ho * ((p.Wo + 1) / 2) * (p.Cp / 8);
  if (fast && total2 < (int64_t)1 << 31 && (int64_t)p.N * p.H * p.W * (p.Cp / 8) < (int64_t)1 << 40) {
    const unsigned grid = grid_for(total2, 256);
    if (dtype == B200_BF16)
      maxpool3x3s2_16b_kernel<__nv_bfloat162><<<grid, 256, 0, s>>>((const uint4*)p.x, (uint4*)p.y, p.N, p.H, p.W, p.Cp / 8, p.Ho, p.Wo, p.pt, p.pl, (uint32_t)total2);
    else
      maxpool3x3s2_16b_kernel<__half2><<<grid, 256, 0, s>>>((const uint4*)p.x, (uint4*)p.y, p.N, p.H, p.W, p.Cp / 8, p.Ho, p.Wo, p.pt, p.pl, (uint32_t)total2);
    return;
  }
  B200_DISPATCH_FLOAT(dtype, T, (pool_kernel<T><<<grid_for(total, 256), 256, 0, s>>>(p)));
}
void launch_global_avgpool(const void* x, void* y, int N, int HW, int Cp, int dtype, cudaStream_t s) {
  int64_t total = (int64_t)N * (Cp / 8);
  if (total >= 32768 && Cp / 8 >= 64 && N <= 65535) {   // enough (image, channel group) threads to fill the GPU with one thread each
    const dim3 grid((unsigned)((Cp / 8 + 127) / 128), (unsigned)N);
    B200_DISPATCH_FLOAT(dtype, T, (global_avgpool_cols_kernel<T, 7><<<grid, 128, 0, s>>>((const T*)x, (T*)y, HW, Cp)));
    return;
  }
  if (HW >= 16 && total < ((int64_t)1 << 28)) {
    B200_DISPATCH_FLOAT(dtype, T, (global_avgpool_warp_kernel<T><<<grid_for(total, 8), 256, 0, s>>>((const T*)x, (T*)y, N, HW, Cp)));
    return;
  }
  B200_DISPATCH_FLOAT(dtype, T, (global_avgpool_kernel<T><<<grid_for(total, 128), 128, 0, s>>>((const T*)x, (T*)y, N, HW, Cp)));
}
void launch_unary(const void* x, void* y, int64_t n, int dtype, int act, float a, float b, cudaStream_t s) {
  if (n == 0) return;
  B200_DISPATCH_FLOAT(dtype, T, (unary_kernel<T><<<grid_for((n + 7) / 8, 256), 256, 0, s>>>((const T*)x, (T*)y, n, act, a, b)));
}
void launch_fake_quantize(const void* x, void* y, int64_t n, int dtype, const float* ranges, int cdim, int64_t inner, int levels, cudaStream_t s) {
  if (n == 0) return;
  B200_DISPATCH_FLOAT(dtype, T, (fake_quantize_kernel<T><<<grid_for(n, 256), 256, 0, s>>>((const T*)x, (T*)y, n, ranges, cdim, inner, (float)(levels - 1))));
}
void launch_binary_same(const void* a, const void* b, void* y, int64_t n, int dtype, int op, int act, float aa, float ab, cudaStream_t s) {
  if (n == 0) return;
  B200_DISPATCH_FLOAT(dtype, T,
                      (binary_same_kernel<T><<<grid_for((n + 7) / 8, 256), 256, 0, s>>>((const T*)a, (const T*)b, (T*)y, n, op, act, aa, ab)));
}
void launch_binary(const BinaryParams& p, int dtype, cudaStream_t s) {
  if (p.n == 0) return;
  B200_DISPATCH_FLOAT(dtype, T, (binary_bcast_kernel<T><<<grid_for(p.n, 256), 256, 0, s>>>(p)));
}
void launch_scale_shift(const void* x, const float* scale, const float* shift, void* y, int64_t rows, int Cp, int dtype, int act,
                        float aa, float ab, cudaStream_t s) {
  int64_t total = rows * (Cp / 8);
  if (total == 0) return;
  B200_DISPATCH_FLOAT(dtype, T,
                      (scale_shift_kernel<T><<<grid_for(total, 256), 256, 0, s>>>((const T*)x, scale, shift, (T*)y, rows, Cp, act, aa, ab)));
}
bool softmax_fused_supported(int64_t cols, int dtype) { return cols % 8 == 0 && cols <= 2048 && (dtype == B200_BF16 || dtype == B200_F16 || dtype == B200_F32); }
void launch_softmax_scale_mask(const void* x, const void* mask, void* y, int64_t rows, int64_t cols, float scale, int64_t rows_per_mask, int dtype,
                               cudaStream_t s) {
  if (rows * cols == 0) return;
  B200_CHECK(softmax_fused_supported(cols, dtype), B200_ERR_UNSUPPORTED, "softmax_scale_mask: unsupported row length " << cols);
  if (dtype != B200_F32 && rows < (int64_t(1) << 31) && rows_per_mask < (int64_t(1) << 31)) {
    // 16-bit rows whose width splits evenly over the warp: exact-width kernel (B200_SOFTMAX_EXACT=0 keeps the guarded one)
    static const bool exact = [] { const char* e = getenv("B200_SOFTMAX_EXACT"); return !(e && e[0] == '0'); }();
    const unsigned g8 = grid_for(rows, 8);
#define B200_SMX(T, W_, C_) softmax_rows_exact_kernel<T, W_, C_><<<g8, 256, 0, s>>>((const T*)x, (const T*)mask, (T*)y, (uint32_t)rows, scale, (uint32_t)rows_per_mask)
#define B200_SMX_COLS(T)                                  \
    switch (exact ? cols : 0) {                            \
      case 128: B200_SMX(T, 2, 1); return;                 \
      case 256: B200_SMX(T, 4, 1); return;                 \
      case 384: B200_SMX(T, 2, 3); return;                 \
      case 512: B200_SMX(T, 4, 2); return;                 \
      case 768: B200_SMX(T, 4, 3); return;                 \
      case 1024: B200_SMX(T, 4, 4); return;                \
      default: break;                                      \
    }
    if (dtype == B200_BF16) { B200_SMX_COLS(__nv_bfloat16) } else { B200_SMX_COLS(__half) }
#undef B200_SMX_COLS
#undef B200_SMX
  }
  const unsigned grid = grid_for(rows, 8);
  const int chunks = (int)((cols + 255) / 256);
#define B200_SM_LAUNCH(T, C_) softmax_warp_rows_kernel<T, C_><<<grid, 256, 0, s>>>((const T*)x, (const T*)mask, (T*)y, rows, (int)cols, scale, rows_per_mask)
#define B200_SM_CHUNKS(T) do { if (chunks <= 1) B200_SM_LAUNCH(T, 1); else if (chunks <= 2) B200_SM_LAUNCH(T, 2); else if (chunks <= 4) B200_SM_LAUNCH(T, 4); else B200_SM_LAUNCH(T, 8); } while (0)
  B200_DISPATCH_FLOAT(dtype, T, B200_SM_CHUNKS(T));
}
void launch_softmax(const void* x, void* y, int64_t outer, int64_t axis, int64_t inner, int dtype, cudaStream_t s) {
  if (outer * axis * inner == 0) return;
  if (inner == 1 && softmax_fused_supported(axis, dtype) && axis >= 64) {   // warp per row, one read + one write per element
    launch_softmax_scale_mask(x, nullptr, y, outer, axis, 1.f, 1, dtype, s);
    return;
  }
  if (inner == 1 && axis <= 128) {
    B200_DISPATCH_FLOAT(dtype, T, (softmax_short_rows_kernel<T><<<(unsigned)((outer + 7) / 8), 256, 0, s>>>((const T*)x, (T*)y, outer, (int)axis)));
    return;
  }
  if (inner == 1) {
    int block = axis >= 512 ? 256 : (axis >= 128 ? 128 : 32);
    B200_DISPATCH_FLOAT(dtype, T, (softmax_rows_kernel<T><<<(unsigned)outer, block, 0, s>>>((const T*)x, (T*)y, outer, axis)));
  } else {
    B200_DISPATCH_FLOAT(dtype, T,
                        (softmax_strided_kernel<T><<<grid_for(outer * inner, 128), 128, 0, s>>>((const T*)x, (T*)y, outer, axis, inner)));
  }
}
void launch_mvn(const void* x, void* y, const float* gamma, const float* beta, int64_t rows, int64_t cols, int normalize_variance,
                int eps_outside, float eps, int dtype, cudaStream_t s) {
  if (rows * cols == 0) return;
  if (dtype != B200_F32 && cols % 256 == 0 && cols <= 1024 && rows < (int64_t(1) << 31) && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 &&
      (uintptr_t)gamma % 16 == 0 && (uintptr_t)beta % 16 == 0) {
    static const bool exact = [] { const char* e = getenv("B200_MVN_EXACT"); return !(e && e[0] == '0'); }();
    if (exact) {
      const unsigned g8 = grid_for(rows, 8);
#define B200_MVNX(T, C_) mvn_rows_exact_kernel<T, C_><<<g8, 256, 0, s>>>((const T*)x, (T*)y, gamma, beta, (uint32_t)rows, normalize_variance, eps_outside, eps)
#define B200_MVNX_COLS(T) switch (cols / 256) { case 1: B200_MVNX(T, 1); break; case 2: B200_MVNX(T, 2); break; case 3: B200_MVNX(T, 3); break; default: B200_MVNX(T, 4); break; }
      if (dtype == B200_BF16) { B200_MVNX_COLS(__nv_bfloat16) } else { B200_MVNX_COLS(__half) }
#undef B200_MVNX_COLS
#undef B200_MVNX
      return;
    }
  }
  if (cols % 8 == 0 && cols <= 2048 && cols >= 64 && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0) {
    const unsigned grid = grid_for(rows, 8);
    const int chunks = (int)((cols + 255) / 256);
#define B200_MVN_LAUNCH(T, C_) mvn_warp_rows_kernel<T, C_><<<grid, 256, 0, s>>>((const T*)x, (T*)y, gamma, beta, rows, (int)cols, normalize_variance, eps_outside, eps)
#define B200_MVN_CHUNKS(T) do { if (chunks <= 1) B200_MVN_LAUNCH(T, 1); else if (chunks <= 2) B200_MVN_LAUNCH(T, 2); else if (chunks <= 4) B200_MVN_LAUNCH(T, 4); else B200_MVN_LAUNCH(T, 8); } while (0)
    B200_DISPATCH_FLOAT(dtype, T, B200_MVN_CHUNKS(T));
    return;
  }
  int block = cols >= 1024 ? 256 : (cols >= 256 ? 128 : 32);
  B200_DISPATCH_FLOAT(dtype, T,
                      (mvn_rows_kernel<T><<<(unsigned)rows, block, 0, s>>>((const T*)x, (T*)y, gamma, beta, cols, normalize_variance,
                                                                          eps_outside, eps)));
}
void launch_gather_rows(const void* table, const int32_t* idx, void* y, int64_t n_idx, int64_t row_elems, int64_t table_rows, int dtype,
                        cudaStream_t s) {
  int64_t total = n_idx * row_elems;
  if (total == 0) return;
  if (dtype == B200_I32) {
    gather_rows_kernel<int32_t><<<grid_for(total, 256), 256, 0, s>>>((const int32_t*)table, idx, (int32_t*)y, n_idx, row_elems, table_rows);
    return;
  }
  B200_DISPATCH_FLOAT(dtype, T,
                      (gather_rows_kernel<T><<<grid_for(total, 256), 256, 0, s>>>((const T*)table, idx, (T*)y, n_idx, row_elems, table_rows)));
}
void launch_reduce(const void* x, void* y, int64_t outer, int64_t axis, int64_t inner, int dtype, int kind, cudaStream_t s) {
  if (outer * inner == 0) return;
  B200_DISPATCH_FLOAT(dtype, T,
                      (reduce_kernel<T><<<grid_for(outer * inner, 128), 128, 0, s>>>((const T*)x, (T*)y, outer, axis, inner, kind)));
}

// Network input with <= 8 channels (images: C = 3), NCHW or dense NHWC in the user's precision -> NHWC8 in the compute type: one thread per
// pixel, C plane-coalesced loads, ONE 16-byte store.  (The element-per-thread kernel above read with a stride of H x W and stored 2 bytes
// at a time: 175 us for MobileNetV2's 64 x 3 x 224 x 224 fp32 input, 17 % of the step.)
template <typename S, typename D>
__global__ void __launch_bounds__(256) ingest_small_c_kernel(const S* __restrict__ src, int slayout, uint4* __restrict__ dst, int C, int64_t HW, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t n = i / HW, px = i - n * HW;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = c < C ? ldf<S>(src, slayout == PL_ROWMAJOR ? (n * C + c) * HW + px : i * C + c) : 0.f;
  uint4 o;
  D* od = reinterpret_cast<D*>(&o);
#pragma unroll
  for (int c = 0; c < 8; ++c) stf<D>(od, c, v[c]);
  dst[i] = o;
}

template <typename S>
static void convert4_dst(const void* src, int slayout, void* dst, int ddtype, int dlayout, int N, int C, int H, int W, cudaStream_t s) {
  int Cd = dlayout == PL_NHWC8 ? (int)round_up(C, 8) : C;
  int64_t total = (int64_t)N * Cd * H * W;
  if (total == 0) return;
  if (dlayout == PL_NHWC8 && C <= 8 && (slayout == PL_ROWMAJOR || slayout == PL_NHWC_DENSE) && (ddtype == B200_BF16 || ddtype == B200_F16)) {
    const int64_t HW = (int64_t)H * W, pixels = (int64_t)N * HW;
    const unsigned g = grid_for(pixels, 256);
    if (ddtype == B200_BF16) ingest_small_c_kernel<S, __nv_bfloat16><<<g, 256, 0, s>>>((const S*)src, slayout, (uint4*)dst, C, HW, pixels);
    else ingest_small_c_kernel<S, __half><<<g, 256, 0, s>>>((const S*)src, slayout, (uint4*)dst, C, HW, pixels);
    return;
  }
  unsigned g = grid_for(total, 256);
  switch (ddtype) {
    case B200_F32: convert4_kernel<S, float><<<g, 256, 0, s>>>((const S*)src, slayout, (float*)dst, dlayout, N, C, H, W); break;
    case B200_F16: convert4_kernel<S, __half><<<g, 256, 0, s>>>((const S*)src, slayout, (__half*)dst, dlayout, N, C, H, W); break;
    case B200_BF16: convert4_kernel<S, __nv_bfloat16><<<g, 256, 0, s>>>((const S*)src, slayout, (__nv_bfloat16*)dst, dlayout, N, C, H, W); break;
    case B200_I32: convert4_kernel<S, int32_t><<<g, 256, 0, s>>>((const S*)src, slayout, (int32_t*)dst, dlayout, N, C, H, W); break;
    case B200_U8: convert4_kernel<S, uint8_t><<<g, 256, 0, s>>>((const S*)src, slayout, (uint8_t*)dst, dlayout, N, C, H, W); break;
    default: B200_THROW(B200_ERR_UNSUPPORTED, "convert4: dst dtype " << ddtype);
  }
}
void launch_convert4(const void* src, int sdtype, int slayout, void* dst, int ddtype, int dlayout, int N, int C, int H, int W, cudaStream_t s) {
  switch (sdtype) {
    case B200_F32: convert4_dst<float>(src, slayout, dst, ddtype, dlayout, N, C, H, W, s); break;
    case B200_F16: convert4_dst<__half>(src, slayout, dst, ddtype, dlayout, N, C, H, W, s); break;
    case B200_BF16: convert4_dst<__nv_bfloat16>(src, slayout, dst, ddtype, dlayout, N, C, H, W, s); break;
    case B200_I32: convert4_dst<int32_t>(src, slayout, dst, ddtype, dlayout, N, C, H, W, s); break;
    case B200_U8: convert4_dst<uint8_t>(src, slayout, dst, ddtype, dlayout, N, C, H, W, s); break;
    default: B200_THROW(B200_ERR_UNSUPPORTED, "convert4: src dtype " << sdtype);
  }
}
template <typename S>
static void ingest_s2d_dst(const void* src, int slayout, void* dst, int ddtype, int N, int C, int H, int W, int pt, int pl, int Hb, int Wb, cudaStream_t s) {
  const int64_t total = (int64_t)N * Hb * Wb;
  if (total == 0) return;
  unsigned g = grid_for(total, 256);
  B200_CHECK(total < ((int64_t)1 << 31), B200_ERR_UNSUPPORTED, "ingest_s2d: tensor too large");
#define B200_INGEST(D, C_) ingest_s2d_kernel<S, D, C_><<<g, 256, 0, s>>>((const S*)src, slayout, (D*)dst, N, H, W, pt, pl, Hb, Wb)
#define B200_INGEST_C(D) do { if (C == 1) B200_INGEST(D, 1); else if (C == 2) B200_INGEST(D, 2); else if (C == 3) B200_INGEST(D, 3); else B200_INGEST(D, 4); } while (0)
  if (ddtype == B200_BF16) B200_INGEST_C(__nv_bfloat16);
  else if (ddtype == B200_F16) B200_INGEST_C(__half);
  else B200_THROW(B200_ERR_UNSUPPORTED, "ingest_s2d: destination must be bf16/f16");
}
void launch_ingest_s2d(const void* src, int sdtype, int slayout, void* dst, int ddtype, int N, int C, int H, int W, int pt, int pl, int Hb, int Wb,
                       cudaStream_t s) {
  B200_CHECK(C >= 1 && C <= 4, B200_ERR_INVALID, "ingest_s2d: at most 4 channels");
  switch (sdtype) {
    case B200_F32: ingest_s2d_dst<float>(src, slayout, dst, ddtype, N, C, H, W, pt, pl, Hb, Wb, s); break;
    case B200_F16: ingest_s2d_dst<__half>(src, slayout, dst, ddtype, N, C, H, W, pt, pl, Hb, Wb, s); break;
    case B200_BF16: ingest_s2d_dst<__nv_bfloat16>(src, slayout, dst, ddtype, N, C, H, W, pt, pl, Hb, Wb, s); break;
    case B200_U8: ingest_s2d_dst<uint8_t>(src, slayout, dst, ddtype, N, C, H, W, pt, pl, Hb, Wb, s); break;
    default: B200_THROW(B200_ERR_UNSUPPORTED, "ingest_s2d: src dtype " << sdtype);
  }
}
template <typename S>
static void convert_flat_dst(const void* src, void* dst, int ddtype, int64_t n, cudaStream_t s) {
  unsigned g = grid_for(n, 256);
  switch (ddtype) {
    case B200_F32: convert_flat_kernel<S, float><<<g, 256, 0, s>>>((const S*)src, (float*)dst, n); break;
    case B200_F16: convert_flat_kernel<S, __half><<<g, 256, 0, s>>>((const S*)src, (__half*)dst, n); break;
    case B200_BF16: convert_flat_kernel<S, __nv_bfloat16><<<g, 256, 0, s>>>((const S*)src, (__nv_bfloat16*)dst, n); break;
    case B200_I32: convert_flat_kernel<S, int32_t><<<g, 256, 0, s>>>((const S*)src, (int32_t*)dst, n); break;
    case B200_U8: convert_flat_kernel<S, uint8_t><<<g, 256, 0, s>>>((const S*)src, (uint8_t*)dst, n); break;
    default: B200_THROW(B200_ERR_UNSUPPORTED, "convert: dst dtype " << ddtype);
  }
}
void launch_convert_flat(const void* src, int sdtype, void* dst, int ddtype, int64_t n, cudaStream_t s) {
  if (n == 0) return;
  if (sdtype == B200_I32 && ddtype == B200_I32) {
    copy_i32_kernel<<<grid_for(n, 256), 256, 0, s>>>((const int32_t*)src, (int32_t*)dst, n);
    return;
  }
  switch (sdtype) {
    case B200_F32: convert_flat_dst<float>(src, dst, ddtype, n, s); break;
    case B200_F16: convert_flat_dst<__half>(src, dst, ddtype, n, s); break;
    case B200_BF16: convert_flat_dst<__nv_bfloat16>(src, dst, ddtype, n, s); break;
    case B200_I32: convert_flat_dst<int32_t>(src, dst, ddtype, n, s); break;
    case B200_U8: convert_flat_dst<uint8_t>(src, dst, ddtype, n, s); break;
    default: B200_THROW(B200_ERR_UNSUPPORTED, "convert: src dtype " << sdtype);
  }
}
void launch_permute(const PermuteParams& p, int elem_size, cudaStream_t s) {
  if (p.n == 0) return;
  // last two dimensions swapped, leading dimensions untouched and dense: batched 2-D transpose through shared memory
  if (elem_size == 2 && p.rank >= 2 && p.rank <= 8 && p.istride[p.rank - 2] == 1 && p.istride[p.rank - 1] == p.oshape[p.rank - 2]) {
    const int64_t A = p.oshape[p.rank - 2], B = p.oshape[p.rank - 1];
    bool dense = A % 2 == 0 && B % 2 == 0 && (uintptr_t)p.x % 4 == 0 && (uintptr_t)p.y % 4 == 0;
    int64_t expect = A * B, batch = 1;
    for (int d = p.rank - 3; d >= 0; --d) { dense = dense && p.istride[d] == expect; expect *= p.oshape[d]; batch *= p.oshape[d]; }
    if (dense && batch <= 65535 * 1 && (A + 63) / 64 < 65536 && (B + 63) / 64 < 65536) {
      dim3 grid((unsigned)((A + 63) / 64), (unsigned)((B + 63) / 64), (unsigned)batch);
      transpose_last2_16b_kernel<<<grid, 256, 0, s>>>((const uint16_t*)p.x, (uint16_t*)p.y, (int)A, (int)B);
      return;
    }
  }
  // innermost dimension contiguous in the source too (e.g. BERT's [N,S,heads,64] <-> [N,heads,S,64]) and everything 16-byte aligned:
  // move 16-byte vectors instead of elements (32-bit index math; the element-wise kernel ran at 0.6 TB/s)
  if (p.rank >= 2 && p.rank <= 8 && p.istride[p.rank - 1] == 1 && elem_size <= 4) {
    const int64_t vec = 16 / elem_size;
    bool ok = p.oshape[p.rank - 1] % vec == 0 && (uintptr_t)p.x % 16 == 0 && (uintptr_t)p.y % 16 == 0 && p.n / vec < ((int64_t)1 << 31);
    for (int d = 0; d + 1 < p.rank; ++d) ok = ok && p.istride[d] % vec == 0;
    if (ok) {
      PermuteParams q = p;
      q.oshape[p.rank - 1] = p.oshape[p.rank - 1] / vec;
      for (int d = 0; d + 1 < p.rank; ++d) q.istride[d] = p.istride[d] / vec;
      q.n = p.n / vec;
      if (q.oshape[p.rank - 1] % 4 == 0) permute_vec16x4_kernel<<<grid_for(q.n / 4, 256), 256, 0, s>>>(q);
      else permute_vec16_kernel<<<grid_for(q.n, 256), 256, 0, s>>>(q);
      return;
    }
  }
  unsigned g = grid_for(p.n, 256);
  if (elem_size == 4) permute_kernel<uint32_t><<<g, 256, 0, s>>>(p);
  else if (elem_size == 2) permute_kernel<uint16_t><<<g, 256, 0, s>>>(p);
  else if (elem_size == 1) permute_kernel<uint8_t><<<g, 256, 0, s>>>(p);
  else B200_THROW(B200_ERR_UNSUPPORTED, "permute: element size " << elem_size);
}
void launch_copy2d(const void* src, int64_t spitch, void* dst, int64_t dpitch, int64_t row_bytes, int64_t rows, cudaStream_t s) {
  if (row_bytes * rows == 0) return;
  bool v16 = row_bytes % 16 == 0 && spitch % 16 == 0 && dpitch % 16 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0;
  bool v4 = row_bytes % 4 == 0 && spitch % 4 == 0 && dpitch % 4 == 0 && (uintptr_t)src % 4 == 0 && (uintptr_t)dst % 4 == 0;
  bool v2 = row_bytes % 2 == 0 && spitch % 2 == 0 && dpitch % 2 == 0 && (uintptr_t)src % 2 == 0 && (uintptr_t)dst % 2 == 0;
  if (v16) copy2d_kernel<uint4><<<grid_for(rows * (row_bytes / 16), 256), 256, 0, s>>>((const char*)src, spitch, (char*)dst, dpitch, row_bytes / 16, rows);
  else if (v4) copy2d_kernel<uint32_t><<<grid_for(rows * (row_bytes / 4), 256), 256, 0, s>>>((const char*)src, spitch, (char*)dst, dpitch, row_bytes / 4, rows);
  else if (v2) copy2d_kernel<uint16_t><<<grid_for(rows * (row_bytes / 2), 256), 256, 0, s>>>((const char*)src, spitch, (char*)dst, dpitch, row_bytes / 2, rows);
  else copy2d_kernel<uint8_t><<<grid_for(rows * row_bytes, 256), 256, 0, s>>>((const char*)src, spitch, (char*)dst, dpitch, row_bytes, rows);
}
void launch_fill_zero(void* dst, int64_t bytes, cudaStream_t s) {
  if (bytes > 0) cudaMemsetAsync(dst, 0, (size_t)bytes, s);
}
void launch_topk(const float* x, float* vals, int32_t* idx, int64_t outer, int64_t axis, int64_t inner, int k, int mode_min, int sort_mode,
                 cudaStream_t s) {
  if (outer * inner == 0) return;
  int block = axis >= 2048 ? 256 : (axis >= 256 ? 128 : 32);
  size_t smem = (32 + 32) * 4 + (size_t)k * 8;
  B200_CHECK(smem <= 48 * 1024, B200_ERR_UNSUPPORTED, "TopK: k=" << k << " too large");
  topk_kernel<<<(unsigned)(outer * inner), block, smem, s>>>(x, vals, idx, axis, inner, k, mode_min, sort_mode);
}

}  // namespace b200
