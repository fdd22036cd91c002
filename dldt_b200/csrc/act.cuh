// Activation functions shared by the fused epilogues and the eltwise kernels.  Formulas follow the
// reference eltwise table (inference-engine/src/mkldnn_plugin/nodes/mkldnn_eltwise_node.cpp:767-901)
// and ngraph/core/reference/include/ngraph/runtime/reference/{gelu,swish,hswish,mish,...}.hpp.
#pragma once
#include "common.h"

namespace b200 {

__device__ __forceinline__ float act_apply(float v, int kind, float a, float b) {
  switch (kind) {
    case ACT_NONE: return v;
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case ACT_TANH: return tanhf(v);
    case ACT_CLAMP: return fminf(fmaxf(v, a), b);
    case ACT_SWISH: return v / (1.f + expf(-a * v));
    case ACT_GELU_ERF: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case ACT_GELU_TANH: return 0.5f * v * (1.f + tanhf(0.79788456080286535588f * (v + 0.044715f * v * v * v)));
    case ACT_HSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f;
    case ACT_HSIGMOID: return fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f;
    case ACT_MISH: return v * tanhf(log1pf(expf(v)));
    case ACT_EXP: return expf(v);
    case ACT_SQRT: return sqrtf(v);
    case ACT_ABS: return fabsf(v);
    case ACT_NEG: return -v;
    case ACT_ERF: return erff(v);
    case ACT_ELU: return v >= 0.f ? v : a * (expf(v) - 1.f);
    case ACT_ROUND_EVEN: return rintf(v);
    case ACT_ROUND_AWAY: return roundf(v);
  }
  return v;
}

}  // namespace b200
