// Internal declarations shared by the graph compiler (host C++) and the executor/kernels (CUDA).
// Nothing in here is visible through the C ABI (include/b200.h).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200.h"

namespace b200 {

// ---- errors -------------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string& msg);
#define B200_THROW(code, ...)                                  \
  do {                                                         \
    std::ostringstream _os;                                    \
    _os << __VA_ARGS__;                                        \
    throw ::b200::Error((code), _os.str());                    \
  } while (0)
#define B200_CHECK(cond, code, ...) \
  do {                              \
    if (!(cond)) B200_THROW(code, __VA_ARGS__); \
  } while (0)

inline size_t dtype_size(int dt) {
  switch (dt) {
    case B200_F32: case B200_I32: return 4;
    case B200_F16: case B200_BF16: return 2;
    case B200_I64: return 8;
    default: return 1;
  }
}
inline const char* dtype_name(int dt) {
  switch (dt) {
    case B200_F32: return "FP32"; case B200_F16: return "FP16"; case B200_BF16: return "BF16";
    case B200_I32: return "I32"; case B200_U8: return "U8"; case B200_I64: return "I64";
    case B200_I8: return "I8"; case B200_BOOL: return "BOOL";
  }
  return "?";
}
inline bool dtype_is_float(int dt) { return dt == B200_F32 || dt == B200_F16 || dt == B200_BF16; }

using Shape = std::vector<int64_t>;
// PRelu: a rank-1 slope whose size equals the channel extent of x is per-channel, i.e. [1,C,1,..] (nGraph reference
// ngraph/core/reference/include/ngraph/runtime/reference/prelu.hpp:27-34; the CPU plugin's ReshapePRelu pass,
// mkldnn_plugin/ngraph_transformations/convert_to_cpu_specific_opset.hpp:36).  Model Optimizer emits slope [C].
inline std::vector<int64_t> prelu_slope_shape(const std::vector<int64_t>& x, const std::vector<int64_t>& slope) {
  const size_t ch = x.size() > 1 ? 1 : 0;
  if (slope.size() == 1 && !x.empty() && x[ch] == slope[0]) {
    std::vector<int64_t> r(x.size(), 1);
    r[ch] = slope[0];
    return r;
  }
  return slope;
}

inline int64_t numel(const Shape& s) {
  int64_t n = 1;
  for (auto d : s) n *= d;
  return n;
}
inline std::string shape_str(const Shape& s) {
  std::ostringstream os;
  os << "[";
  for (size_t i = 0; i < s.size(); ++i) os << (i ? "," : "") << s[i];
  os << "]";
  return os.str();
}
inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// ---- graph IR (what b200_graph_add builds) ---------------------------------------------------
struct Value {
  int id = -1;
  int dtype = B200_F32;
  Shape shape;
  int producer = -1;   // node id
  int out_index = 0;
  std::vector<int> consumers;  // node ids (with multiplicity)
  std::string name;
};

struct Node {
  int id = -1;
  int op = 0;
  std::string name;
  std::vector<int> inputs;   // value ids
  std::vector<int> outputs;  // value ids
  int32_t i[B200_MAX_IATTR] = {0};
  float f[B200_MAX_FATTR] = {0};
  int dtype = B200_F32;
  Shape shape;
  std::vector<uint8_t> data;  // constants (stored as f32 / i32 / u8 after normalisation)
};

}  // namespace b200

struct b200_graph {
  std::string name;
  std::vector<b200::Node> nodes;
  std::vector<b200::Value> values;
  std::vector<int> params;   // node ids
  std::vector<int> results;  // node ids
};

namespace b200 {

const char* op_name(int op);
bool is_unary_act(int op);
bool is_binary_eltwise(int op);
// Shape/type inference for one node (throws Error); fills out_shapes/out_dtypes.
void infer_node(const b200_graph& g, const Node& n, std::vector<Shape>& out_shapes, std::vector<int>& out_dtypes);

// ---- lowered program ------------------------------------------------------------------------
// L_S2D16: the network input of a stride-2 stem convolution, stored space-to-depth: padded pixel (hp, wp) = (h + pt, w + pl) lives in
// block (hp / 2, wp / 2), channel ((hp % 2) * 2 + wp % 2) * C + c of a [N, Hb, Wb, 16] tensor (C <= 4; unused channels are zero)
enum Layout : int { L_ROWMAJOR = 0, L_NHWC8 = 1, L_S2D16 = 2 };

enum Space : int { SP_NONE = 0, SP_ARENA = 1, SP_WEIGHTS = 2, SP_IO_IN = 3, SP_IO_OUT = 4 };
struct BufRef {
  int space = SP_NONE;
  int index = 0;        // io index for SP_IO_*
  int64_t offset = 0;   // byte offset within the space
  int box = -1;         // arena box id before planning
};

struct LTensor {
  Shape shape;          // logical (nGraph) shape
  int dtype = B200_F32; // storage dtype
  int layout = L_ROWMAJOR;
  int s2d_pt = 0, s2d_pl = 0, s2d_hb = 0, s2d_wb = 0;   // L_S2D16 geometry
  BufRef buf;
  int64_t bytes() const;
  int64_t phys_elems() const;
};

// activation codes used by fused epilogues and the eltwise kernels
enum Act : int {
  ACT_NONE = 0, ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_CLAMP, ACT_SWISH, ACT_GELU_ERF, ACT_GELU_TANH,
  ACT_HSWISH, ACT_HSIGMOID, ACT_MISH, ACT_EXP, ACT_SQRT, ACT_ABS, ACT_NEG, ACT_ERF, ACT_ELU,
  ACT_ROUND_EVEN, ACT_ROUND_AWAY   // ngraph reference/round.hpp: HALF_TO_EVEN (nearbyint) / HALF_AWAY_FROM_ZERO (round)
};
struct ActSpec {
  int kind = ACT_NONE;
  float a = 0.f, b = 0.f;
};
int act_from_op(const Node& n, ActSpec& spec);  // returns 1 if n is a fusable activation

struct ArenaBox;
struct RunCtx {
  char* arena = nullptr;
  const int64_t* box_offset = nullptr;  // planned offset of every arena box
  char* weights = nullptr;
  std::vector<char*> io_in;
  std::vector<char*> io_out;
  void* stream = nullptr;  // cudaStream_t
  inline char* ptr(const BufRef& r) const {
    switch (r.space) {
      case SP_ARENA: return arena + box_offset[r.box] + r.offset;
      case SP_WEIGHTS: return weights + r.offset;
      case SP_IO_IN: return io_in[r.index] + r.offset;
      case SP_IO_OUT: return io_out[r.index] + r.offset;
      default: return nullptr;
    }
  }
};

struct Step {
  std::string name;      // friendly name of the producing node
  std::string type;      // kernel class
  std::string impl;
  std::string fused;     // comma separated original names
  std::string out_precision, out_layout;
  double flops = 0, bytes = 0;
  int n_launches = 1;
  std::vector<int> reads, writes;  // arena box ids (for the planner's live ranges)
  LTensor out;           // primary output (for dump)
  std::function<void(const RunCtx&)> launch;
  // alternative implementations of the same step (name, launcher); the first request of a network times them on its own arena and
  // keeps the fastest (b200_network::autotune).  Empty = nothing to choose.
  std::vector<std::pair<std::string, std::function<void(const RunCtx&)>>> alts;
};

struct IOBinding {
  std::string name;
  Shape shape;
  int graph_dtype;      // dtype of the graph edge
  int value_id;
};

struct ArenaBox {
  int64_t bytes = 0;
  int first = 1 << 30, last = -1;  // step indices
  int64_t offset = -1;
  bool pinned = false;             // lives for the whole run (not reused)
};

}  // namespace b200
