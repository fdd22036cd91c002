// Small helpers shared by the tcgen05 kernels (conv_tc.cu, conv_tc2.cu): element type traits and bf16/fp16 packing.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

#include "act.cuh"

namespace b200 {
namespace tc {

template <typename T> struct TcType;
template <> struct TcType<__nv_bfloat16> { static constexpr int is_bf16 = 1; };
template <> struct TcType<__half> { static constexpr int is_bf16 = 0; };

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// pack with ReLU folded into the conversion (cvt.rn.relu): max(x, 0) then round, one instruction per pair instead of three
template <typename T>
__device__ __forceinline__ uint32_t pack2_relu(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2_relu<__nv_bfloat16>(float a, float b) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));   // first source -> upper half
  return d;
}
template <> __device__ __forceinline__ uint32_t pack2_relu<__half>(float a, float b) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
template <typename T>
__device__ __forceinline__ void unpack2(uint32_t u, float& a, float& b);
template <> __device__ __forceinline__ void unpack2<__nv_bfloat16>(uint32_t u, float& a, float& b) {
  // a bf16 is the upper half of the fp32 with the same value: one shift and one mask (the library's __low2float / __high2float compile
  // to PRMT + IMAD.U32 per element; ncu: the unpacks were 40 % of the depthwise stage's instructions in mbconv.cu)
  a = __uint_as_float(u << 16); b = __uint_as_float(u & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack2<__half>(uint32_t u, float& a, float& b) {
  __half2 v = *reinterpret_cast<__half2*>(&u);
  a = __low2float(v); b = __high2float(v);
}

// element-wise max of two packed pairs
template <typename T>
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t max2<__nv_bfloat16>(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t max2<__half>(uint32_t a, uint32_t b) {
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// Epilogue activation over 8 accumulators.  GENERIC == false covers the activations that are a clamp (none / ReLU / Clamp,
// i.e. every ResNet / MobileNet / SSD layer) with two instructions per element; GENERIC == true keeps the full table of
// act.cuh but as ONE rolled loop, so the switch is instantiated once per call site instead of once per element (the fully
// unrolled form made the kernel 26k instructions long and the epilogue instruction-fetch bound: profiles/r1_run2).
struct EpiAct {
  int relu_in_pack;   // set by the launcher on the LAST activation of the chain when it is a ReLU: the epilogue skips it and packs with cvt.relu
  int kind;       // Act code (generic path)
  float a, b;     // its parameters
  float lo, hi;   // clamp bounds (fast path)
};
// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, two MUFU + ~10 FMA): used only where the result is rounded to bf16/fp16 right
// after (the tensor-core epilogues); the fp32 path keeps erff (act.cuh).  GELU(erf) = 0.5 x (1 + erf(x / sqrt 2)), reference/gelu.hpp:19-38.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.f - p * t * __expf(-ax * ax);
  return copysignf(r, x);
}

// Generic (non-clamp) activation over N accumulators: ONE warp-uniform switch, fully unrolled bodies for the activations the model zoo uses
// (a rolled loop indexes f[] dynamically, which puts the accumulators in local memory).  The pair kernel calls it once per 64-column slab:
// with one switch per 8 values the GELU path hopped through sixteen copies of the whole switch per slab and the epilogue warps sat in
// instruction-fetch stalls (ncu: stall_no_inst 43% of their samples, BERT's FC + GELU at 0.4x the plain FC, profiles/r1_gelu_ncu.txt).
// GELU(erf) with the same polynomial, rearranged: erf(z) = sign(z) (1 - q), q = p(t) t exp(-z^2), z = x / sqrt 2, so
// 0.5 x (1 + erf z) = relu(x) - 0.5 |x| q.  14 FP32 instructions + 2 MUFU per value (the composed form above takes 19).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.f)));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);      // the 0.5 of "0.5 |x| q" folded into the coefficients
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * (ax * -0.72134752044448170368f)));   // exp(-x^2 / 2) = 2^(-x^2 log2(e) / 2)
  return fmaf(-(ax * t), p * e, fmaxf(x, 0.f));
}

template <int N>
__device__ __forceinline__ void epi_act_generic(float (&f)[N], const EpiAct& s) {
  switch (s.kind) {
    case ACT_GELU_ERF:
#pragma unroll
      for (int e = 0; e < N; ++e) f[e] = gelu_erf_fast(f[e]);
      break;
    case ACT_SWISH:
#pragma unroll
      for (int e = 0; e < N; ++e) f[e] = __fdividef(f[e], 1.f + __expf(-s.a * f[e]));
      break;
    case ACT_SIGMOID:
#pragma unroll
      for (int e = 0; e < N; ++e) f[e] = __fdividef(1.f, 1.f + __expf(-f[e]));
      break;
    case ACT_TANH:
#pragma unroll
      for (int e = 0; e < N; ++e) f[e] = tanhf(f[e]);
      break;
    case ACT_HSWISH:
#pragma unroll
      for (int e = 0; e < N; ++e) f[e] = f[e] * fminf(fmaxf(f[e] + 3.f, 0.f), 6.f) * (1.f / 6.f);
      break;
    case ACT_RELU:
#pragma unroll
      for (int e = 0; e < N; ++e) f[e] = fmaxf(f[e], 0.f);
      break;
    case ACT_CLAMP:
#pragma unroll
      for (int e = 0; e < N; ++e) f[e] = fminf(fmaxf(f[e], s.a), s.b);
      break;
    default: {
      float g[N];
#pragma unroll
      for (int e = 0; e < N; ++e) g[e] = f[e];
#pragma unroll 1
      for (int e = 0; e < N; ++e) g[e] = act_apply(g[e], s.kind, s.a, s.b);
#pragma unroll
      for (int e = 0; e < N; ++e) f[e] = g[e];
    }
  }
}

template <bool GENERIC>
__device__ __forceinline__ void epi_act8(float (&f)[8], const EpiAct& s) {
  if (s.kind == 0 || s.relu_in_pack) return;   // ACT_NONE, or a trailing ReLU done by the bf16 conversion (warp-uniform)
  if constexpr (!GENERIC) {
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], s.lo);
    if (s.hi < 3.0e38f) {   // Clamp proper; ReLU has hi = +inf and needs one instruction per element (warp-uniform branch)
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = fminf(f[e], s.hi);
    }
  } else {
    epi_act_generic<8>(f, s);
  }
}

// Clamp-family epilogues (none / ReLU / Clamp) run as straight-line code, one variant per launch (conv_tc3.cuh epilogue_slab, conv_tc2.cu
// epilogue_warp_fast): bounds of the two activation slots, (-inf, +inf) where a slot is empty.
struct EpiBounds { float lo1, hi1, lo2, hi2; };
// which of the straight-line variants a launch runs (warp-uniform, fixed for the launch): 0 = plain, 1 = trailing ReLU in the pack, 2 = clamps
__device__ __forceinline__ int epi_variant(bool has_res, const EpiAct& e1, const EpiAct& e2, EpiBounds& eb) {
  const float inf = __int_as_float(0x7f800000);
  const bool n1 = e1.kind == 0, n2 = e2.kind == 0;
  // an activation that is "none" clamps to (-inf, +inf); a trailing ReLU the launcher moved into the pack is the clamp (0, +inf) here
  eb.lo1 = n1 ? -inf : (e1.relu_in_pack ? 0.f : e1.lo); eb.hi1 = (n1 || e1.relu_in_pack) ? inf : e1.hi;
  eb.lo2 = n2 ? -inf : (e2.relu_in_pack ? 0.f : e2.lo); eb.hi2 = (n2 || e2.relu_in_pack) ? inf : e2.hi;
  if (!has_res) return n1 ? 0 : (e1.relu_in_pack ? 1 : 2);
  if (!n1) return 2;
  return n2 ? 0 : (e2.relu_in_pack ? 1 : 2);
}

}  // namespace tc
}  // namespace b200
