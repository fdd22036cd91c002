// Graph construction + shape/type inference (host only).  Semantics follow the nGraph opset
// definitions the reference validates with (ngraph/core/src/op/*.cpp validate_and_infer_types);
// the formulas are cited next to each case.
#include <algorithm>
#include <cmath>

#include "common.h"

namespace b200 {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }

const char* op_name(int op) {
  static const char* names[] = {
      "Parameter", "Constant", "Result", "Convolution", "GroupConvolution", "BatchNormInference", "Add",
      "Subtract", "Multiply", "Divide", "Maximum", "Minimum", "Relu", "Sigmoid", "Tanh", "Clamp", "Swish",
      "Gelu", "HSwish", "HSigmoid", "Mish", "Exp", "Sqrt", "Abs", "Negative", "Erf", "Elu", "MaxPool",
      "AvgPool", "ReduceMean", "MatMul", "Softmax", "MVN", "Reshape", "Transpose", "Concat", "Split",
      "Gather", "TopK", "NonMaxSuppression", "DetectionOutput", "Convert", "Power", "PRelu", "ReduceSum",
      "ReduceMax", "SquaredDifference", "Round", "FakeQuantize", "StridedSlice"};
  if (op < 0 || op >= B200_OP__COUNT) return "?";
  return names[op];
}

bool is_unary_act(int op) {
  switch (op) {
    case B200_OP_RELU: case B200_OP_SIGMOID: case B200_OP_TANH: case B200_OP_CLAMP: case B200_OP_SWISH:
    case B200_OP_GELU: case B200_OP_HSWISH: case B200_OP_HSIGMOID: case B200_OP_MISH: case B200_OP_EXP:
    case B200_OP_SQRT: case B200_OP_ABS: case B200_OP_NEGATIVE: case B200_OP_ERF: case B200_OP_ELU: case B200_OP_ROUND:
      return true;
  }
  return false;
}
bool is_binary_eltwise(int op) {
  switch (op) {
    case B200_OP_ADD: case B200_OP_SUBTRACT: case B200_OP_MULTIPLY: case B200_OP_DIVIDE:
    case B200_OP_MAXIMUM: case B200_OP_MINIMUM: case B200_OP_POWER: case B200_OP_SQUARED_DIFFERENCE:
    case B200_OP_PRELU:
      return true;
  }
  return false;
}

int act_from_op(const Node& n, ActSpec& s) {
  s = ActSpec();
  switch (n.op) {
    case B200_OP_RELU: s.kind = ACT_RELU; return 1;
    case B200_OP_SIGMOID: s.kind = ACT_SIGMOID; return 1;
    case B200_OP_TANH: s.kind = ACT_TANH; return 1;
    case B200_OP_CLAMP: s.kind = ACT_CLAMP; s.a = n.f[0]; s.b = n.f[1]; return 1;
    case B200_OP_SWISH: s.kind = ACT_SWISH; s.a = n.f[0]; return 1;
    case B200_OP_GELU: s.kind = n.i[0] ? ACT_GELU_TANH : ACT_GELU_ERF; return 1;
    case B200_OP_HSWISH: s.kind = ACT_HSWISH; return 1;
    case B200_OP_HSIGMOID: s.kind = ACT_HSIGMOID; return 1;
    case B200_OP_MISH: s.kind = ACT_MISH; return 1;
    case B200_OP_EXP: s.kind = ACT_EXP; return 1;
    case B200_OP_SQRT: s.kind = ACT_SQRT; return 1;
    case B200_OP_ABS: s.kind = ACT_ABS; return 1;
    case B200_OP_NEGATIVE: s.kind = ACT_NEG; return 1;
    case B200_OP_ERF: s.kind = ACT_ERF; return 1;
    case B200_OP_ELU: s.kind = ACT_ELU; s.a = n.f[0]; return 1;
    case B200_OP_ROUND: s.kind = n.i[0] ? ACT_ROUND_AWAY : ACT_ROUND_EVEN; return 1;
  }
  return 0;
}

static Shape broadcast_shapes(const Shape& a, const Shape& b, const std::string& who) {
  size_t r = std::max(a.size(), b.size());
  Shape out(r, 1);
  for (size_t k = 0; k < r; ++k) {
    int64_t da = k < r - a.size() ? 1 : a[k - (r - a.size())];
    int64_t db = k < r - b.size() ? 1 : b[k - (r - b.size())];
    B200_CHECK(da == db || da == 1 || db == 1, B200_ERR_INVALID,
               who << ": shapes " << shape_str(a) << " and " << shape_str(b) << " do not broadcast");
    out[k] = std::max(da, db);
  }
  return out;
}

// out = floor((in + pb + pe - (dil*(k-1)+1)) / stride) + 1  (reference/convolution.hpp:89-146)
static int64_t conv_out_dim(int64_t in, int64_t k, int64_t s, int64_t d, int64_t pb, int64_t pe) {
  int64_t eff = d * (k - 1) + 1;
  int64_t num = in + pb + pe - eff;
  B200_CHECK(num >= 0 && s > 0, B200_ERR_INVALID, "convolution window larger than padded input");
  return num / s + 1;
}

// Pooling output dim, floor or ceil rounding (ngraph/core/src/validation_util.cpp infer_batched_pooling_forward).
static int64_t pool_out_dim(int64_t in, int64_t k, int64_t s, int64_t pb, int64_t pe, bool ceil_mode) {
  int64_t num = in + pb + pe - k;
  B200_CHECK(num >= 0 && s > 0, B200_ERR_INVALID, "pooling window larger than padded input");
  return (ceil_mode ? (num + s - 1) / s : num / s) + 1;
}

void infer_node(const b200_graph& g, const Node& n, std::vector<Shape>& os, std::vector<int>& od) {
  auto in = [&](int k) -> const Value& {
    B200_CHECK(k < (int)n.inputs.size(), B200_ERR_INVALID, op_name(n.op) << " '" << n.name << "': missing input " << k);
    return g.values[n.inputs[k]];
  };
  auto need_inputs = [&](int k) {
    B200_CHECK((int)n.inputs.size() == k, B200_ERR_INVALID,
               op_name(n.op) << " '" << n.name << "' expects " << k << " inputs, got " << n.inputs.size());
  };
  os.clear();
  od.clear();
  switch (n.op) {
    case B200_OP_PARAMETER:
    case B200_OP_CONSTANT:
      os.push_back(n.shape);
      od.push_back(n.dtype);
      break;
    case B200_OP_RESULT:
      need_inputs(1);
      os.push_back(in(0).shape);
      od.push_back(in(0).dtype);
      break;
    case B200_OP_CONVOLUTION: {
      need_inputs(2);
      const Shape& x = in(0).shape; const Shape& w = in(1).shape;
      B200_CHECK(x.size() == 4 && w.size() == 4, B200_ERR_UNSUPPORTED, "Convolution '" << n.name << "': only 2D (rank-4) supported");
      B200_CHECK(x[1] == w[1], B200_ERR_INVALID, "Convolution '" << n.name << "': C_in mismatch " << shape_str(x) << " vs " << shape_str(w));
      os.push_back({x[0], w[0], conv_out_dim(x[2], w[2], n.i[0], n.i[2], n.i[4], n.i[6]),
                    conv_out_dim(x[3], w[3], n.i[1], n.i[3], n.i[5], n.i[7])});
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_GROUP_CONVOLUTION: {
      need_inputs(2);
      const Shape& x = in(0).shape; const Shape& w = in(1).shape;
      B200_CHECK(x.size() == 4 && w.size() == 5, B200_ERR_UNSUPPORTED, "GroupConvolution '" << n.name << "': only 2D supported");
      B200_CHECK(x[1] == w[0] * w[2], B200_ERR_INVALID, "GroupConvolution '" << n.name << "': C_in mismatch");
      os.push_back({x[0], w[0] * w[1], conv_out_dim(x[2], w[3], n.i[0], n.i[2], n.i[4], n.i[6]),
                    conv_out_dim(x[3], w[4], n.i[1], n.i[3], n.i[5], n.i[7])});
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_BATCH_NORM:
      need_inputs(5);
      os.push_back(in(0).shape);
      od.push_back(in(0).dtype);
      break;
    case B200_OP_ADD: case B200_OP_SUBTRACT: case B200_OP_MULTIPLY: case B200_OP_DIVIDE:
    case B200_OP_MAXIMUM: case B200_OP_MINIMUM: case B200_OP_POWER: case B200_OP_SQUARED_DIFFERENCE:
      need_inputs(2);
      os.push_back(broadcast_shapes(in(0).shape, in(1).shape, n.name));
      od.push_back(in(0).dtype);
      break;
    case B200_OP_PRELU:
      need_inputs(2);
      os.push_back(broadcast_shapes(in(0).shape, prelu_slope_shape(in(0).shape, in(1).shape), n.name));
      od.push_back(in(0).dtype);
      break;
    case B200_OP_RELU: case B200_OP_SIGMOID: case B200_OP_TANH: case B200_OP_CLAMP: case B200_OP_SWISH:
    case B200_OP_GELU: case B200_OP_HSWISH: case B200_OP_HSIGMOID: case B200_OP_MISH: case B200_OP_EXP:
    case B200_OP_SQRT: case B200_OP_ABS: case B200_OP_NEGATIVE: case B200_OP_ERF: case B200_OP_ELU: case B200_OP_ROUND:
      B200_CHECK(n.inputs.size() >= 1, B200_ERR_INVALID, "unary op without input");
      os.push_back(in(0).shape);
      od.push_back(in(0).dtype);
      break;
    case B200_OP_FAKE_QUANTIZE: {
      // ngraph/core/src/op/fake_quantize.cpp: output = shape / type of the data input; the four range inputs broadcast to it.  Supported
      // here: every range is one value or one value per channel (axis 1) - what Model Optimizer / POT emit
      need_inputs(5);
      B200_CHECK(n.i[0] >= 2, B200_ERR_INVALID, "FakeQuantize '" << n.name << "': levels must be >= 2");
      const int64_t C = in(0).shape.size() >= 2 ? in(0).shape[1] : 1;
      for (int k = 1; k < 5; ++k) {
        const int64_t cnt = numel(in(k).shape);
        B200_CHECK(cnt == 1 || cnt == C, B200_ERR_UNSUPPORTED, "FakeQuantize '" << n.name << "': range input " << k << " must hold 1 or C = " << C << " values");
        if (cnt == C && cnt != 1) {
          const Shape& rs = in(k).shape;
          const size_t off = in(0).shape.size() - rs.size();
          for (size_t d = 0; d < rs.size(); ++d)
            B200_CHECK(rs[d] == 1 || d + off == 1, B200_ERR_UNSUPPORTED, "FakeQuantize '" << n.name << "': per-channel ranges must lie on axis 1");
        }
      }
      os.push_back(in(0).shape);
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_MAX_POOL: case B200_OP_AVG_POOL: {
      need_inputs(1);
      const Shape& x = in(0).shape;
      B200_CHECK(x.size() == 4, B200_ERR_UNSUPPORTED, "Pooling '" << n.name << "': only 2D supported");
      bool ceil_mode = n.i[8] != 0;
      os.push_back({x[0], x[1], pool_out_dim(x[2], n.i[0], n.i[2], n.i[4], n.i[6], ceil_mode),
                    pool_out_dim(x[3], n.i[1], n.i[3], n.i[5], n.i[7], ceil_mode)});
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_REDUCE_MEAN: case B200_OP_REDUCE_SUM: case B200_OP_REDUCE_MAX: {
      B200_CHECK(n.inputs.size() >= 1, B200_ERR_INVALID, "reduce without input");
      const Shape& x = in(0).shape;
      bool keep = n.i[0] != 0;
      int na = n.i[1];
      B200_CHECK(na >= 0 && na <= B200_MAX_IATTR - 2, B200_ERR_INVALID, "reduce '" << n.name << "': bad axis count " << na);
      std::vector<bool> red(x.size(), false);
      for (int k = 0; k < na; ++k) {
        int a = n.i[2 + k];
        if (a < 0) a += (int)x.size();
        B200_CHECK(a >= 0 && a < (int)x.size(), B200_ERR_INVALID, "reduce axis out of range");
        red[a] = true;
      }
      Shape o;
      for (size_t k = 0; k < x.size(); ++k) {
        if (!red[k]) o.push_back(x[k]);
        else if (keep) o.push_back(1);
      }
      os.push_back(o);
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_MATMUL: {
      // numpy batch broadcast, transpose flags on the last two dims (ngraph/core/src/op/matmul.cpp)
      need_inputs(2);
      Shape a = in(0).shape, b = in(1).shape;
      B200_CHECK(!a.empty() && !b.empty(), B200_ERR_INVALID, "MatMul on scalar");
      bool ta = n.i[0] != 0, tb = n.i[1] != 0;
      bool a1 = a.size() == 1, b1 = b.size() == 1;
      if (a1) a = {1, a[0]};            // 1-D lhs is a row vector (transpose ignored)
      else if (ta) std::swap(a[a.size() - 1], a[a.size() - 2]);
      if (b1) b = {b[0], 1};            // 1-D rhs is a column vector
      else if (tb) std::swap(b[b.size() - 1], b[b.size() - 2]);
      B200_CHECK(a.back() == b[b.size() - 2], B200_ERR_INVALID,
                 "MatMul '" << n.name << "': K mismatch " << shape_str(in(0).shape) << " x " << shape_str(in(1).shape));
      Shape ba(a.begin(), a.end() - 2), bb(b.begin(), b.end() - 2);
      Shape o = broadcast_shapes(ba, bb, n.name);
      if (!a1) o.push_back(a[a.size() - 2]);
      if (!b1) o.push_back(b.back());
      os.push_back(o);
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_SOFTMAX:
      need_inputs(1);
      B200_CHECK(n.i[0] >= 0 && n.i[0] < (int)in(0).shape.size(), B200_ERR_INVALID, "Softmax axis out of range");
      os.push_back(in(0).shape);
      od.push_back(in(0).dtype);
      break;
    case B200_OP_MVN:
      B200_CHECK(n.inputs.size() >= 1, B200_ERR_INVALID, "MVN without input");
      B200_CHECK(n.i[2] >= 0 && n.i[2] <= B200_MAX_IATTR - 3, B200_ERR_INVALID, "MVN '" << n.name << "': bad axis count " << n.i[2]);
      os.push_back(in(0).shape);
      od.push_back(in(0).dtype);
      break;
    case B200_OP_RESHAPE: {
      B200_CHECK(n.inputs.size() >= 1, B200_ERR_INVALID, "Reshape without input");
      B200_CHECK(numel(n.shape) == numel(in(0).shape), B200_ERR_INVALID,
                 "Reshape '" << n.name << "': " << shape_str(in(0).shape) << " -> " << shape_str(n.shape));
      os.push_back(n.shape);
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_STRIDED_SLICE: {
      B200_CHECK(n.inputs.size() >= 1, B200_ERR_INVALID, "StridedSlice without input");
      const int r = n.i[0];
      B200_CHECK(r >= 1 && r <= B200_MAX_RANK && 1 + 4 * r <= B200_MAX_IATTR, B200_ERR_INVALID, "StridedSlice '" << n.name << "': bad rank " << r);
      int64_t vin = 1, vout = 1;
      for (int d = 0; d < r; ++d) {
        const int64_t v = n.i[1 + d], b = n.i[1 + r + d], st = n.i[1 + 2 * r + d], cnt = n.i[1 + 3 * r + d];
        B200_CHECK(v >= 1 && cnt >= 1 && st != 0, B200_ERR_INVALID, "StridedSlice '" << n.name << "': empty dimension or zero stride");
        const int64_t last = b + (cnt - 1) * st;
        B200_CHECK(b >= 0 && b < v && last >= 0 && last < v, B200_ERR_INVALID, "StridedSlice '" << n.name << "': dimension " << d << " out of range");
        vin *= v; vout *= cnt;
      }
      B200_CHECK(vin == numel(in(0).shape) && vout == numel(n.shape), B200_ERR_INVALID,
                 "StridedSlice '" << n.name << "': " << shape_str(in(0).shape) << " -> " << shape_str(n.shape) << " does not match the slice plan");
      os.push_back(n.shape);
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_TRANSPOSE: {
      B200_CHECK(n.inputs.size() >= 1, B200_ERR_INVALID, "Transpose without input");
      const Shape& x = in(0).shape;
      Shape o(x.size());
      std::vector<bool> seen(x.size(), false);
      for (size_t k = 0; k < x.size(); ++k) {
        int p = n.i[k];
        B200_CHECK(p >= 0 && p < (int)x.size() && !seen[p], B200_ERR_INVALID, "Transpose '" << n.name << "': bad permutation");
        seen[p] = true;
        o[k] = x[p];
      }
      os.push_back(o);
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_CONCAT: {
      B200_CHECK(n.inputs.size() >= 1, B200_ERR_INVALID, "Concat without input");
      Shape o = in(0).shape;
      int ax = n.i[0] < 0 ? n.i[0] + (int)o.size() : n.i[0];
      B200_CHECK(ax >= 0 && ax < (int)o.size(), B200_ERR_INVALID, "Concat axis out of range");
      for (size_t k = 1; k < n.inputs.size(); ++k) {
        const Shape& s = in((int)k).shape;
        B200_CHECK(s.size() == o.size(), B200_ERR_INVALID, "Concat rank mismatch");
        for (size_t d = 0; d < s.size(); ++d)
          B200_CHECK((int)d == ax || s[d] == o[d], B200_ERR_INVALID, "Concat '" << n.name << "': shape mismatch");
        o[ax] += s[ax];
      }
      os.push_back(o);
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_SPLIT: {
      B200_CHECK(n.inputs.size() >= 1, B200_ERR_INVALID, "Split without input");
      Shape o = in(0).shape;
      int ax = n.i[0] < 0 ? n.i[0] + (int)o.size() : n.i[0];
      int parts = n.i[1];
      B200_CHECK(ax >= 0 && ax < (int)o.size() && parts > 0 && o[ax] % parts == 0 && parts <= B200_MAX_INPUTS,
                 B200_ERR_INVALID, "Split '" << n.name << "': bad axis/num_splits");
      o[ax] /= parts;
      for (int k = 0; k < parts; ++k) { os.push_back(o); od.push_back(in(0).dtype); }
      break;
    }
    case B200_OP_GATHER: {
      B200_CHECK(n.inputs.size() >= 2, B200_ERR_INVALID, "Gather needs data and indices");
      const Shape& d = in(0).shape; const Shape& idx = in(1).shape;
      int ax = n.i[0] < 0 ? n.i[0] + (int)d.size() : n.i[0];
      B200_CHECK(ax >= 0 && ax < (int)d.size(), B200_ERR_INVALID, "Gather axis out of range");
      Shape o(d.begin(), d.begin() + ax);
      o.insert(o.end(), idx.begin(), idx.end());
      o.insert(o.end(), d.begin() + ax + 1, d.end());
      os.push_back(o);
      od.push_back(in(0).dtype);
      break;
    }
    case B200_OP_TOPK: {
      B200_CHECK(n.inputs.size() >= 1, B200_ERR_INVALID, "TopK without input");
      Shape o = in(0).shape;
      int ax = n.i[1] < 0 ? n.i[1] + (int)o.size() : n.i[1];
      B200_CHECK(ax >= 0 && ax < (int)o.size() && n.i[0] >= 1 && n.i[0] <= o[ax], B200_ERR_INVALID, "TopK '" << n.name << "': bad k/axis");
      o[ax] = n.i[0];
      os.push_back(o); od.push_back(in(0).dtype);
      os.push_back(o); od.push_back(B200_I32);  // I64 graph outputs are I32 at the API (cnn_network_ngraph_impl.cpp:147-150)
      break;
    }
    case B200_OP_NMS: {
      // NonMaxSuppression-5 static upper bound: B * C * min(N, max_out) rows (nms_ie_internal shape inference)
      B200_CHECK(n.inputs.size() >= 2, B200_ERR_INVALID, "NMS needs boxes and scores");
      const Shape& b = in(0).shape; const Shape& s = in(1).shape;
      B200_CHECK(b.size() == 3 && s.size() == 3 && b[2] == 4 && b[0] == s[0] && b[1] == s[2], B200_ERR_INVALID, "NMS '" << n.name << "': bad input shapes");
      int64_t per = std::min<int64_t>(b[1], n.i[0]);
      int64_t rows = s[0] * s[1] * per;
      os.push_back({rows, 3}); od.push_back(B200_I32);
      os.push_back({rows, 3}); od.push_back(B200_F32);
      os.push_back({1}); od.push_back(B200_I32);
      break;
    }
    case B200_OP_DETECTION_OUTPUT: {
      // [1,1,N*keep_top_k,7] (ngraph/core/src/op/detection_output.cpp); keep_top_k<0 -> top_k or all priors
      B200_CHECK(n.inputs.size() >= 3, B200_ERR_INVALID, "DetectionOutput needs loc, conf, priors");
      const Shape& loc = in(0).shape; const Shape& pri = in(2).shape;
      B200_CHECK(loc.size() == 2 && pri.size() == 3, B200_ERR_INVALID, "DetectionOutput '" << n.name << "': bad ranks");
      int64_t N = loc[0];
      int64_t prior_size = n.i[7] ? 4 : 5;
      int64_t num_priors = pri[2] / prior_size;
      int64_t keep = n.i[3];
      int64_t rows;
      if (keep > 0) rows = N * keep;
      else if (n.i[2] > 0) rows = N * n.i[2] * n.i[0];
      else rows = N * num_priors * n.i[0];
      os.push_back({1, 1, rows, 7});
      od.push_back(B200_F32);
      break;
    }
    case B200_OP_CONVERT:
      need_inputs(1);
      os.push_back(in(0).shape);
      od.push_back(n.dtype);
      break;
    default:
      B200_THROW(B200_ERR_UNSUPPORTED, "unsupported op id " << n.op);
  }
}

}  // namespace b200

using namespace b200;

static float half_to_float(uint16_t h) {
  uint32_t sign = (h >> 15) & 1, exp = (h >> 10) & 0x1f, man = h & 0x3ff;
  uint32_t f;
  if (exp == 0) {
    if (man == 0) f = sign << 31;
    else {
      exp = 127 - 15 + 1;
      while (!(man & 0x400)) { man <<= 1; --exp; }
      man &= 0x3ff;
      f = (sign << 31) | (exp << 23) | (man << 13);
    }
  } else if (exp == 31) f = (sign << 31) | 0x7f800000u | (man << 13);
  else f = (sign << 31) | ((exp - 15 + 127) << 23) | (man << 13);
  float out;
  memcpy(&out, &f, 4);
  return out;
}

extern "C" {

int b200_abi_version(void) { return B200_ABI_VERSION; }
const char* b200_last_error(void) { return b200::last_error().c_str(); }

int b200_op_supported(int32_t op) { return op >= 0 && op < B200_OP__COUNT ? 1 : 0; }

int b200_graph_create(const char* name, b200_graph** out) {
  if (!out) return B200_ERR_INVALID;
  *out = new b200_graph();
  (*out)->name = name ? name : "network";
  return B200_OK;
}
void b200_graph_destroy(b200_graph* g) { delete g; }

int b200_graph_add(b200_graph* g, const b200_node_desc* d, int32_t* out_ids, int32_t* n_outputs) {
  try {
    B200_CHECK(g && d && out_ids && n_outputs, B200_ERR_INVALID, "null argument");
    B200_CHECK(d->n_inputs >= 0 && d->n_inputs <= B200_MAX_INPUTS, B200_ERR_INVALID, "too many inputs");
    Node n;
    n.id = (int)g->nodes.size();
    n.op = d->op;
    n.name = d->name ? d->name : (std::string(op_name(d->op)) + "_" + std::to_string(n.id));
    for (int k = 0; k < d->n_inputs; ++k) {
      B200_CHECK(d->inputs[k] >= 0 && d->inputs[k] < (int)g->values.size(), B200_ERR_INVALID,
                 n.name << ": input value id " << d->inputs[k] << " does not exist");
      n.inputs.push_back(d->inputs[k]);
    }
    memcpy(n.i, d->i, sizeof(n.i));
    memcpy(n.f, d->f, sizeof(n.f));
    n.dtype = d->dtype;
    B200_CHECK(d->rank >= 0 && d->rank <= B200_MAX_RANK, B200_ERR_INVALID, "rank too large");
    n.shape.assign(d->shape, d->shape + d->rank);
    if (n.op == B200_OP_CONSTANT) {
      int64_t cnt = numel(n.shape);
      B200_CHECK(d->data != nullptr || cnt == 0, B200_ERR_INVALID, n.name << ": constant without data");
      B200_CHECK((uint64_t)cnt * dtype_size(n.dtype) == d->data_bytes, B200_ERR_INVALID,
                 n.name << ": constant byte size " << d->data_bytes << " != " << cnt * dtype_size(n.dtype));
      // normalise storage: floats -> f32, ints -> i32, u8/bool kept
      if (n.dtype == B200_F32 || n.dtype == B200_I32 || n.dtype == B200_U8 || n.dtype == B200_BOOL || n.dtype == B200_I8) {
        n.data.assign((const uint8_t*)d->data, (const uint8_t*)d->data + d->data_bytes);
      } else if (n.dtype == B200_F16 || n.dtype == B200_BF16) {
        n.data.resize(cnt * 4);
        const uint16_t* src = (const uint16_t*)d->data;
        float* dst = (float*)n.data.data();
        for (int64_t k = 0; k < cnt; ++k) {
          if (n.dtype == B200_F16) dst[k] = half_to_float(src[k]);
          else { uint32_t u = (uint32_t)src[k] << 16; memcpy(&dst[k], &u, 4); }
        }
        n.dtype = B200_F32;
      } else if (n.dtype == B200_I64) {
        n.data.resize(cnt * 4);
        const int64_t* src = (const int64_t*)d->data;
        int32_t* dst = (int32_t*)n.data.data();
        for (int64_t k = 0; k < cnt; ++k) dst[k] = (int32_t)std::max<int64_t>(INT32_MIN, std::min<int64_t>(INT32_MAX, src[k]));
        n.dtype = B200_I32;
      } else {
        B200_THROW(B200_ERR_UNSUPPORTED, "constant dtype " << n.dtype);
      }
    }
    if (n.op == B200_OP_PARAMETER && n.dtype == B200_I64) n.dtype = B200_I32;
    std::vector<Shape> os;
    std::vector<int> od;
    infer_node(*g, n, os, od);
    *n_outputs = (int)os.size();
    for (size_t k = 0; k < os.size(); ++k) {
      Value v;
      v.id = (int)g->values.size();
      v.dtype = od[k];
      v.shape = os[k];
      v.producer = n.id;
      v.out_index = (int)k;
      v.name = os.size() > 1 ? n.name + "." + std::to_string(k) : n.name;
      n.outputs.push_back(v.id);
      out_ids[k] = v.id;
      g->values.push_back(v);
    }
    for (int vi : n.inputs) g->values[vi].consumers.push_back(n.id);
    if (n.op == B200_OP_PARAMETER) g->params.push_back(n.id);
    if (n.op == B200_OP_RESULT) g->results.push_back(n.id);
    g->nodes.push_back(std::move(n));
    return B200_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return B200_ERR_INVALID;
  }
}

int b200_graph_value_info(const b200_graph* g, int32_t id, b200_tensor_info* info) {
  if (!g || !info || id < 0 || id >= (int)g->values.size()) return B200_ERR_INVALID;
  const Value& v = g->values[id];
  memset(info, 0, sizeof(*info));
  snprintf(info->name, sizeof(info->name), "%s", v.name.c_str());
  info->dtype = v.dtype;
  info->rank = (int)v.shape.size();
  for (size_t k = 0; k < v.shape.size(); ++k) info->shape[k] = v.shape[k];
  return B200_OK;
}

}  // extern "C"
