// Graph compiler: lowers the op graph built through b200_graph_add into a fused schedule of kernel
// launches over a statically planned HBM arena.
//
// What it mirrors in the reference (structure, not code):
//   * fusion legality: MKLDNNGraphOptimizer::ApplyCommonGraphOptimizations
//     (inference-engine/src/mkldnn_plugin/mkldnn_graph_optimizer.cpp:53-148): conv + bias (:150-182),
//     conv/FC + simple activation (:587-629, :809-852), conv + sum (+ activation) (:958-1130),
//     MVN + scale/shift (:1132-1168); BatchNorm folding as BatchNormDecomposition + ConvolutionMultiplyFusion
//     (transformations/op_conversions/batch_norm_decomposition.cpp:33-80, common_optimizations/conv_mul_fusion.cpp:20-80)
//   * MatMul with constant weights -> FullyConnected (ngraph_transformations/convert_matmul_to_fc_or_gemm.cpp:32-130)
//   * layout selection + reorder insertion (mkldnn_graph.cpp:455-511): rank-4 activations are NHWC with channels
//     padded to 8 ("NHWC8") wherever a layout-aware kernel exists, dense row-major elsewhere, conversion kernels
//     are inserted on demand
//   * weights are repacked once at compile time (ExecuteConstantNodesOnly, mkldnn_graph.cpp:322-351)
#include <algorithm>
#include <cmath>
#include <set>

#include "engine.h"

namespace b200 {

int64_t LTensor::phys_elems() const {
  if (layout == L_S2D16) return shape[0] * (int64_t)s2d_hb * s2d_wb * 16;
  if (layout == L_NHWC8 && shape.size() == 4) return shape[0] * shape[2] * shape[3] * round_up(shape[1], 8);
  return numel(shape);
}
int64_t LTensor::bytes() const { return phys_elems() * (int64_t)dtype_size(dtype); }

namespace {

// ---- host-side scalar conversions ----------------------------------------------------------------
uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)((u >> 16) | ((u & 0xffff) ? 0x40 : 0));
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
uint16_t f32_to_f16(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
  uint32_t man = x & 0x7fffffu;
  if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
  if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
  if (exp <= 0) {
    if (exp < -10) return (uint16_t)sign;
    man |= 0x800000u;
    int shift = 14 - exp;
    uint32_t r = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (r & 1))) ++r;
    return (uint16_t)(sign | r);
  }
  uint32_t r = (uint32_t)(exp << 10) | (man >> 13);
  uint32_t rem = man & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) ++r;
  return (uint16_t)(sign | r);
}
void store_float(uint8_t* dst, int64_t i, float v, int dtype) {
  if (dtype == B200_F32) reinterpret_cast<float*>(dst)[i] = v;
  else if (dtype == B200_BF16) reinterpret_cast<uint16_t*>(dst)[i] = f32_to_bf16(v);
  else reinterpret_cast<uint16_t*>(dst)[i] = f32_to_f16(v);
}

struct FusedConv {
  int root = -1;                 // conv / group conv / matmul node
  std::vector<int> absorbed;     // node ids folded in (in order)
  std::vector<float> scale, shift;  // per output channel
  ActSpec act1, act2;
  int res_value = -1;
  int out_value = -1;            // value produced by the last absorbed node
  int emit_at = -1;              // node id at whose position the kernel is emitted
  int heads_s = 0, heads_h = 0;  // FullyConnected whose [N,S,H*64] -> [N,H,S,64] head split is written by the kernel's stores
  // K-concatenated projection shortcut (fold_shortcuts): a second 1x1 convolution whose output was this one's residual
  int dual_root = -1;
  std::vector<float> dual_scale, dual_shift;
  std::string dual_fused;
  // the next block's 1x1 reduce convolution, computed from this one's output inside the same kernel (fold_next_reduce, conv_tc3f.cu)
  int next_root = -1, next_out_value = -1;
  std::vector<float> next_scale, next_shift;
  ActSpec next_act;
  std::string next_fused;
};
// Softmax over the last axis with the attention prologue folded in: softmax(x * scale + mask), mask broadcast over rows
struct FusedSoftmax {
  int x_value = -1, mask_value = -1;
  float scale = 1.f;
  std::string fused;
};
// softmax(Q K^T * scale + mask) V (+ the merged-heads transpose) as ONE kernel (attention.cu); emitted at the context MatMul's position
struct FusedAttn {
  int q_value = -1, k_value = -1, v_value = -1, mask_value = -1;
  float scale = 1.f;
  int out_value = -1;
  std::string fused;
};
struct FusedMvn {
  int root = -1;
  std::vector<int> absorbed;
  int gamma_value = -1, beta_value = -1;
  int out_value = -1, emit_at = -1;
};
struct FusedElt {  // binary / unary followed by one activation
  int root = -1;
  ActSpec act;
  int out_value = -1, emit_at = -1;
  int absorbed = -1;
};

class Lowerer {
 public:
  Lowerer(const b200_graph& g, const b200_compile_options& o, b200_network& n) : g_(g), opts_(o), net_(n) {
    cd_ = o.precision == B200_PREC_BF16 ? B200_BF16 : (o.precision == B200_PREC_FP16 ? B200_F16 : B200_F32);
    net_.cdtype = cd_;
    net_.precision = o.precision;
    use_tc_ = cd_ != B200_F32 && o.conv_impl == 0;
  }
  void run();

 private:
  const b200_graph& g_;
  const b200_compile_options& opts_;
  b200_network& net_;
  int cd_;
  bool use_tc_;
  std::map<int, LTensor> vt_;                       // value -> primary lowered tensor
  std::map<std::pair<int, int>, LTensor> alt_;      // (value, layout) -> converted copy
  std::vector<int> absorbed_by_;                    // node -> fused record owner (-1 none)
  std::map<int, FusedConv> fconv_;                  // emit_at -> record
  std::map<int, FusedSoftmax> fsm_;                 // Softmax node id -> record
  std::map<int, FusedAttn> fattn_;                  // emit_at -> record
  std::map<int, FusedMvn> fmvn_;
  std::map<int, FusedElt> felt_;
  std::map<int, int> box_dir_;                      // arena box -> 1 if its producer wrote it last-row-first

  // ---- small helpers ----
  const Node& node(int id) const { return g_.nodes[id]; }
  const Value& val(int id) const { return g_.values[id]; }
  const Node* const_node(int value) const {
    const Node& p = node(val(value).producer);
    return p.op == B200_OP_CONSTANT ? &p : nullptr;
  }
  int storage_dtype(int graph_dtype) const { return dtype_is_float(graph_dtype) ? cd_ : (graph_dtype == B200_I64 ? B200_I32 : graph_dtype); }
  std::vector<float> const_floats(int value) const {
    const Node* c = const_node(value);
    B200_CHECK(c, B200_ERR_UNSUPPORTED, "expected a constant input for value " << val(value).name);
    int64_t n = numel(c->shape);
    std::vector<float> out(n);
    if (c->dtype == B200_F32) memcpy(out.data(), c->data.data(), n * 4);
    else if (c->dtype == B200_I32) for (int64_t i = 0; i < n; ++i) out[i] = (float)reinterpret_cast<const int32_t*>(c->data.data())[i];
    else for (int64_t i = 0; i < n; ++i) out[i] = (float)c->data[i];
    return out;
  }
  int new_box(int64_t bytes, bool pinned = false) {
    ArenaBox b;
    b.bytes = round_up(std::max<int64_t>(bytes, 16), 1024);
    b.pinned = pinned;
    net_.boxes.push_back(b);
    return (int)net_.boxes.size() - 1;
  }
  LTensor alloc(const Shape& shape, int dtype, int layout, bool pinned = false) {
    LTensor t;
    t.shape = shape; t.dtype = dtype; t.layout = layout;
    t.buf.space = SP_ARENA;
    t.buf.box = new_box(t.bytes(), pinned);
    return t;
  }
  int64_t add_weights(const void* data, int64_t bytes) {
    int64_t off = round_up((int64_t)net_.weights_host.size(), 1024);
    net_.weights_host.resize(off + round_up(bytes, 16));
    if (data) memcpy(net_.weights_host.data() + off, data, bytes);
    return off;
  }
  BufRef weights_ref(int64_t off) { BufRef r; r.space = SP_WEIGHTS; r.offset = off; return r; }
  // fp32 per-channel vector in the weights blob
  BufRef add_f32_vector(const std::vector<float>& v) { return weights_ref(add_weights(v.data(), (int64_t)v.size() * 4)); }

  void emit(Step&& s, std::initializer_list<const LTensor*> reads, std::initializer_list<const LTensor*> writes) {
    for (auto* t : reads) if (t && t->buf.space == SP_ARENA) s.reads.push_back(t->buf.box);
    for (auto* t : writes) if (t && t->buf.space == SP_ARENA) s.writes.push_back(t->buf.box);
    if (writes.size()) {
      s.out = **writes.begin();
      s.out_precision = dtype_name(s.out.dtype);
      s.out_layout = s.out.layout == L_NHWC8 ? "nhwc8" : "rowmajor";
    }
    net_.flops += s.flops;
    net_.bytes += s.bytes;
    net_.steps.push_back(std::move(s));
  }

  // ---- tensors ----
  LTensor materialize_const(int value, int layout);
  LTensor get(int value, int layout /* -1 any */);
  LTensor to_dtype(const LTensor& t, int dtype, const std::string& why);
  void set(int value, const LTensor& t) { vt_[value] = t; }
  bool prefers_nhwc8(int value) const;
  bool stem_s2d_applies(int value) const;
  void lower_stem_s2d(const FusedConv& f, const LTensor& x);

  // ---- passes ----
  void fuse();
  void fold_shortcuts();
  void fold_next_reduce();
  void fold_attention();
  void lower_attention(const FusedAttn& f);
  bool per_channel_const(int value, int64_t C, bool matmul_style, std::vector<float>& out) const;
  void lower_node(const Node& n);
  void lower_conv(const FusedConv& f);
  bool try_lower_mbconv(const FusedConv& f);
  std::set<int> mb_done_;                           // emit positions of convolutions already computed by an mbconv kernel
  void lower_matmul_general(const Node& n, const FusedConv* f);
  void lower_binary(const Node& n, const ActSpec& act, int out_value);
  void lower_unary(const Node& n, const ActSpec& act);
  void lower_pool(const Node& n);
  void lower_fake_quantize(const Node& n);
  void lower_reduce(const Node& n);
  void lower_mvn(const Node& n, const FusedMvn* f);
  void lower_concat(const Node& n);
  void lower_split(const Node& n);
  void lower_transpose(const Node& n);
  void lower_strided_slice(const Node& n);
  void lower_detection(const Node& n);
  void conv_launch_step(const std::string& name, const std::string& type, const std::string& fused, ConvParams p, BufRef x, BufRef w,
                        BufRef bias, BufRef res, BufRef y, const LTensor* tx, const LTensor* tres, const LTensor* ty, double flops,
                        double bytes, bool tc, BufRef x2 = BufRef(), const LTensor* tx2 = nullptr, BufRef w2 = BufRef(), BufRef bias2 = BufRef(),
                        BufRef z = BufRef(), const LTensor* tz = nullptr);
};

bool Lowerer::prefers_nhwc8(int value) const {
  const Value& v = val(value);
  if (v.shape.size() != 4 || !dtype_is_float(v.dtype)) return false;
  for (int c : v.consumers) {
    int op = node(c).op;
    if (op == B200_OP_CONVOLUTION || op == B200_OP_GROUP_CONVOLUTION || op == B200_OP_MAX_POOL || op == B200_OP_AVG_POOL ||
        op == B200_OP_BATCH_NORM)
      return true;
  }
  return false;
}

// A network input whose single consumer is a dense 7x7 (or 8x8) stride-2 convolution over <= 4 channels (the ResNet / SSD stem)
// is ingested space-to-depth so that the stem becomes a 4x4 stride-1 convolution over 16 channels: one filter row = 4 taps x 16
// channels = 64 contiguous elements per output pixel = exactly one K block, fetched by one im2col-mode TMA box (conv_tc2.cu).
bool Lowerer::stem_s2d_applies(int value) const {
  if (!use_tc_ || !conv_tc2_im2col_enabled()) return false;
  if (opts_.device >= 0 && !conv_tc2_s2d_encode_ok()) return false;
  const char* off = getenv("B200_STEM_S2D");
  if (off && off[0] == '0') return false;
  const Value& v = val(value);
  if (v.shape.size() != 4 || !dtype_is_float(v.dtype) || v.shape[1] > 4 || v.consumers.size() != 1) return false;
  const Node& c = node(v.consumers[0]);
  if (c.op != B200_OP_CONVOLUTION || c.inputs[0] != value || !const_node(c.inputs[1])) return false;
  const Shape& w = val(c.inputs[1]).shape;
  const Shape& o = val(c.outputs[0]).shape;
  // 7x7 / 8x8 (ResNet, SSD) and 3x3 / 5x5 (MobileNet's first convolution: 3x3 stride 2 over 3 channels) windows all sit inside the 4x4 window
  // of 2x2-pixel blocks; the unused taps of the packed weights stay zero
  const bool k_ok = w[2] >= 3 && w[2] <= 8 && w[3] >= 3 && w[3] <= 8;
  // every tap row/column must stay inside the Hb x Wb block grid: padded extent <= 2 * (Ho + 3)
  const bool fits = v.shape[2] + c.i[4] <= 2 * (o[2] + 3) && v.shape[3] + c.i[5] <= 2 * (o[3] + 3);
  // windows smaller than 7x7 waste most of the 4x4-block MMAs: worth it only where the pair kernel's stem mode runs (every SM pair has a tile);
  // small launches keep the direct CUDA-core kernel for few input channels
  if (w[2] < 7 || w[3] < 7) {
    const char* f3 = getenv("B200_TC3");
    const bool forced = f3 && f3[0] == '1';
    if (!forced && o[0] * ((o[2] + 3) / 4) < 148) return false;
    if (f3 && f3[0] == '0') return false;
  }
  return k_ok && c.i[0] == 2 && c.i[1] == 2 && c.i[2] == 1 && c.i[3] == 1 && c.i[4] <= 8 && c.i[5] <= 8 && fits && o[2] >= 1 && o[3] >= 1;
}

void Lowerer::lower_stem_s2d(const FusedConv& f, const LTensor& x) {
  const Node& n = node(f.root);
  std::string fused = n.name;
  for (int a : f.absorbed) fused += "," + node(a).name;
  const Shape& ws = val(n.inputs[1]).shape;
  const Shape os = val(n.outputs[0]).shape;
  const int64_t Cin = x.shape[1], Cout = os[1], kh = ws[2], kw = ws[3], Cout_p = round_up(Cout, 8);
  const int bn = conv_tc_pick_bn((int)Cout_p);
  const int64_t Cout_alloc = round_up(Cout_p, bn), Ktot = 256;
  std::vector<float> wsrc = const_floats(n.inputs[1]);
  std::vector<uint8_t> wp((size_t)(Cout_alloc * Ktot) * dtype_size(cd_), 0);
  // K index = block row br (4) x block column bs (4) x channel ((dr * 2 + ds) * Cin + c); taps beyond the 7x7 window stay zero
  for (int64_t o = 0; o < Cout; ++o)
    for (int64_t c = 0; c < Cin; ++c)
      for (int64_t r = 0; r < kh; ++r)
        for (int64_t s2 = 0; s2 < kw; ++s2) {
          const int64_t br = r / 2, dr = r % 2, bs = s2 / 2, ds = s2 % 2;
          store_float(wp.data(), o * Ktot + br * 64 + bs * 16 + (dr * 2 + ds) * Cin + c, wsrc[((o * Cin + c) * kh + r) * kw + s2] * f.scale[o], cd_);
        }
  std::vector<float> bias(Cout_alloc, 0.f);
  for (int64_t o = 0; o < Cout; ++o) bias[o] = f.shift[o];
  BufRef wref = weights_ref(add_weights(wp.data(), (int64_t)wp.size()));
  BufRef bref = add_f32_vector(bias);
  LTensor res;
  BufRef rref;
  if (f.res_value >= 0) { res = get(f.res_value, L_NHWC8); rref = res.buf; }
  ConvParams p;
  p.N = (int)x.shape[0]; p.H = x.s2d_hb; p.W = (int)os[3]; p.Cin_p = 64; p.Ho = (int)os[2]; p.Wo = (int)os[3]; p.Cout_p = (int)Cout_p;
  p.kh = 4; p.kw = 1; p.sh = p.sw = p.dh = p.dw = 1; p.pt = p.pl = 0;
  p.x_sw = 16; p.x_sh = (int64_t)x.s2d_wb * 16; p.x_sn = (int64_t)x.s2d_hb * x.s2d_wb * 16;
  p.M = (int)(os[0] * os[2] * os[3]); p.Ktot = (int)Ktot; p.Cout_alloc = (int)Cout_alloc; p.ldy = (int)Cout_p;
  p.w_ld = (int)Ktot; p.w_rows = (int)Cout_alloc; p.w_k = (int)Ktot;
  p.act1 = f.act1.kind; p.a1 = f.act1.a; p.b1 = f.act1.b; p.act2 = f.act2.kind; p.a2 = f.act2.a; p.b2 = f.act2.b;
  // ReLU -> MaxPool 3x3 / stride 2 / pad 1 behind the stem (ResNet's pool1): taken in the stem's epilogue, which max-reduces its pooled
  // rows into the zero-filled pooled tensor; the stem's own output (411 MB at batch 256) is never written.  B200_FUSE_POOL=0 turns it off.
  int pool_node = -1;
  {
    const Value& ov = val(f.out_value);
    const int u = ov.consumers.size() == 1 ? ov.consumers[0] : -1;
    if (u >= 0 && absorbed_by_[u] < 0 && node(u).op == B200_OP_MAX_POOL && f.res_value < 0) {
      const Node& pn = node(u);
      const Shape& ps = val(pn.outputs[0]).shape;
      const bool geom = pn.i[0] == 3 && pn.i[1] == 3 && pn.i[2] == 2 && pn.i[3] == 2 && pn.i[4] == 1 && pn.i[5] == 1 && pn.i[6] <= 1 && pn.i[7] <= 1 && ps[2] * 2 == os[2] && ps[3] * 2 == os[3];
      ConvParams q = p;
      q.pool = 1;
      if (geom && conv_tc3_preferred(p) && conv_tc3_s2d_pool_ok(q)) pool_node = u;
    }
  }
  const int out_value = pool_node >= 0 ? node(pool_node).outputs[0] : f.out_value;
  LTensor out = alloc(val(out_value).shape, cd_, L_NHWC8);
  const double flops = 2.0 * p.M * Cout * Cin * kh * kw;
  const double bytes = (double)x.bytes() + (double)out.bytes() + (double)Cout * Cin * kh * kw * dtype_size(cd_) + (f.res_value >= 0 ? (double)out.bytes() : 0);
  if (pool_node >= 0) {
    p.pool = 1;
    fused += "," + node(pool_node).name;
    absorbed_by_[pool_node] = n.id;
    BufRef yr = out.buf;
    const size_t nbytes = (size_t)out.bytes();
    Step z;
    z.name = val(out_value).name + "/zero"; z.type = "Fill"; z.impl = "memset"; z.fused = z.name;
    z.bytes = (double)nbytes;
    z.launch = [=](const RunCtx& c) { cudaMemsetAsync(c.ptr(yr), 0, nbytes, (cudaStream_t)c.stream); };
    emit(std::move(z), {}, {&out});
  }
  conv_launch_step(val(out_value).name, "Convolution", fused, p, x.buf, wref, bref, rref, out.buf, &x, f.res_value >= 0 ? &res : nullptr, &out, flops,
                   bytes, true);
  net_.steps.back().impl += pool_node >= 0 ? "_s2d_stem_maxpool" : "_s2d_stem";
  set(out_value, out);
}

LTensor Lowerer::materialize_const(int value, int layout) {
  const Node* c = const_node(value);
  LTensor t;
  t.shape = c->shape;
  t.dtype = storage_dtype(c->dtype);
  t.layout = (layout == L_NHWC8 && c->shape.size() == 4 && dtype_is_float(c->dtype)) ? L_NHWC8 : L_ROWMAJOR;
  int64_t n = numel(c->shape);
  std::vector<uint8_t> buf(std::max<int64_t>(t.bytes(), 16), 0);
  if (dtype_is_float(c->dtype)) {
    const float* src = reinterpret_cast<const float*>(c->data.data());
    if (t.layout == L_NHWC8) {
      int64_t N = t.shape[0], C = t.shape[1], H = t.shape[2], W = t.shape[3], Cp = round_up(C, 8);
      for (int64_t in = 0; in < N; ++in)
        for (int64_t ic = 0; ic < C; ++ic)
          for (int64_t h = 0; h < H; ++h)
            for (int64_t w = 0; w < W; ++w)
              store_float(buf.data(), ((in * H + h) * W + w) * Cp + ic, src[((in * C + ic) * H + h) * W + w], t.dtype);
    } else {
      for (int64_t i = 0; i < n; ++i) store_float(buf.data(), i, src[i], t.dtype);
    }
  } else {
    memcpy(buf.data(), c->data.data(), std::min<size_t>(buf.size(), c->data.size()));
  }
  t.buf = weights_ref(add_weights(buf.data(), (int64_t)buf.size()));
  return t;
}

LTensor Lowerer::get(int value, int layout) {
  auto it = vt_.find(value);
  if (it == vt_.end()) {
    B200_CHECK(const_node(value), B200_ERR_INVALID, "value '" << val(value).name << "' used before it is produced");
    auto key = std::make_pair(value, layout == L_NHWC8 ? L_NHWC8 : L_ROWMAJOR);
    auto a = alt_.find(key);
    if (a != alt_.end()) return a->second;
    LTensor t = materialize_const(value, layout);
    alt_[std::make_pair(value, t.layout)] = t;
    return t;
  }
  LTensor t = it->second;
  if (layout < 0 || t.layout == layout) return t;
  auto a = alt_.find(std::make_pair(value, layout));
  if (a != alt_.end()) return a->second;
  B200_CHECK(t.shape.size() == 4, B200_ERR_INVALID, "layout conversion needs rank 4 (" << val(value).name << ")");
  const int64_t N = t.shape[0], C = t.shape[1], H = t.shape[2], W = t.shape[3];
  LTensor o;
  if (layout == L_ROWMAJOR && H == 1 && W == 1 && C % 8 == 0) {  // [N,1,1,C] NHWC8 is bit-identical to [N,C,1,1] dense
    o = t;
    o.layout = L_ROWMAJOR;
  } else if (layout == L_NHWC8 && H == 1 && W == 1 && C % 8 == 0) {
    o = t;
    o.layout = L_NHWC8;
  } else {
    o = alloc(t.shape, t.dtype, layout);
    Step s;
    s.name = val(value).name + (layout == L_NHWC8 ? "_to_nhwc8" : "_to_nchw");
    s.type = "Reorder";
    s.impl = layout == L_NHWC8 ? "nchw_to_nhwc8" : "nhwc8_to_nchw";
    s.bytes = (double)t.bytes() + (double)o.bytes();
    BufRef src = t.buf, dst = o.buf;
    int dt = t.dtype, sl = t.layout == L_NHWC8 ? PL_NHWC8 : PL_ROWMAJOR, dl = layout == L_NHWC8 ? PL_NHWC8 : PL_ROWMAJOR;
    s.launch = [=](const RunCtx& c) {
      launch_convert4(c.ptr(src), dt, sl, c.ptr(dst), dt, dl, (int)N, (int)C, (int)H, (int)W, (cudaStream_t)c.stream);
    };
    emit(std::move(s), {&t}, {&o});
  }
  alt_[std::make_pair(value, layout)] = o;
  return o;
}

LTensor Lowerer::to_dtype(const LTensor& t, int dtype, const std::string& why) {
  if (t.dtype == dtype) return t;
  B200_CHECK(t.layout == L_ROWMAJOR, B200_ERR_INVALID, "dtype conversion expects row-major tensor");
  LTensor o = alloc(t.shape, dtype, L_ROWMAJOR);
  Step s;
  s.name = why + "_convert";
  s.type = "Convert";
  s.impl = std::string(dtype_name(t.dtype)) + "_to_" + dtype_name(dtype);
  s.bytes = (double)t.bytes() + (double)o.bytes();
  BufRef src = t.buf, dst = o.buf;
  int sd = t.dtype;
  int64_t n = numel(t.shape);
  s.launch = [=](const RunCtx& c) { launch_convert_flat(c.ptr(src), sd, c.ptr(dst), dtype, n, (cudaStream_t)c.stream); };
  emit(std::move(s), {&t}, {&o});
  return o;
}

// Is `value` a constant that broadcasts per output channel?  conv style: scalar, [C,1,1], [1,C,1,1];
// matmul style: scalar, [N], [1,N], [1,1,N] ...
bool Lowerer::per_channel_const(int value, int64_t C, bool matmul_style, std::vector<float>& out) const {
  const Node* c = const_node(value);
  if (!c || !(dtype_is_float(c->dtype) || c->dtype == B200_I32)) return false;
  int64_t n = numel(c->shape);
  std::vector<float> v = const_floats(value);
  if (n == 1) { out.assign(C, v[0]); return true; }
  if (n != C) return false;
  const Shape& s = c->shape;
  if (matmul_style) {
    if (s.back() != C) return false;
  } else {
    // the C extent must sit at the channel position counting from the right: [.., C, 1, 1]
    if (s.size() < 3 || s[s.size() - 3] != C) return false;
  }
  out = v;
  return true;
}

void Lowerer::fuse() {
  absorbed_by_.assign(g_.nodes.size(), -1);
  auto single_consumer = [&](int value) -> int {
    const Value& v = val(value);
    return v.consumers.size() == 1 ? v.consumers[0] : -1;
  };
  // Attention scores: Softmax(Add(Multiply|Divide(x, scalar const), mask)) over the last axis with a mask of shape [N,1,..,1,S] or
  // [1,..,1,S] -> one kernel, one pass (launch_softmax_scale_mask).  Either of the two prologue ops may be missing.
  for (const Node& n : g_.nodes) {
    if (n.op != B200_OP_SOFTMAX) continue;
    const Shape& xs = val(n.inputs[0]).shape;
    const int rank = (int)xs.size();
    if (rank < 2 || n.i[0] != rank - 1 || !dtype_is_float(val(n.inputs[0]).dtype)) continue;
    const int64_t cols = xs.back();
    if (cols % 8 != 0 || cols > 2048) continue;
    FusedSoftmax f;
    f.x_value = n.inputs[0];
    std::vector<int> absorbed;
    {   // mask add
      const Value& v = val(f.x_value);
      if (v.producer >= 0 && v.consumers.size() == 1 && node(v.producer).op == B200_OP_ADD && absorbed_by_[v.producer] < 0) {
        const Node& a = node(v.producer);
        for (int side = 0; side < 2; ++side) {
          const int mv = a.inputs[side], xv = a.inputs[1 - side];
          const Shape& ms = val(mv).shape;
          if (val(xv).shape != xs || (int)ms.size() != rank || ms.back() != cols || const_node(mv)) continue;
          bool ok = true;
          for (int d = 1; d + 1 < rank; ++d) ok = ok && ms[d] == 1;
          ok = ok && (ms[0] == 1 || ms[0] == xs[0]) && dtype_is_float(val(mv).dtype);
          if (!ok) continue;
          f.mask_value = mv;
          f.x_value = xv;
          absorbed.push_back(a.id);
          f.fused = a.name;
          break;
        }
      }
    }
    {   // scalar scale
      const Value& v = val(f.x_value);
      if (v.producer >= 0 && v.consumers.size() == 1 && absorbed_by_[v.producer] < 0 &&
          (node(v.producer).op == B200_OP_MULTIPLY || node(v.producer).op == B200_OP_DIVIDE)) {
        const Node& m = node(v.producer);
        for (int side = 0; side < 2; ++side) {
          const int cv = m.inputs[side], xv = m.inputs[1 - side];
          if (!const_node(cv) || numel(val(cv).shape) != 1 || val(xv).shape != xs) continue;
          if (m.op == B200_OP_DIVIDE && side == 0) continue;   // scalar / x is not a scale
          const float c = const_floats(cv)[0];
          f.scale = m.op == B200_OP_MULTIPLY ? c : 1.f / c;
          f.x_value = xv;
          absorbed.push_back(m.id);
          f.fused = m.name + (f.fused.empty() ? "" : "," + f.fused);
          break;
        }
      }
    }
    if (absorbed.empty()) continue;
    for (int a : absorbed) absorbed_by_[a] = n.id;
    fsm_[n.id] = f;
  }
  for (const Node& n : g_.nodes) {
    if (absorbed_by_[n.id] >= 0) continue;
    bool is_conv = n.op == B200_OP_CONVOLUTION || n.op == B200_OP_GROUP_CONVOLUTION;
    bool is_fc = n.op == B200_OP_MATMUL && const_node(n.inputs[1]) != nullptr && n.i[0] == 0;
    bool is_mm = n.op == B200_OP_MATMUL && !is_fc;
    if (is_conv && !const_node(n.inputs[1])) continue;  // dynamic weights: not supported by the fused path
    if (is_conv || is_fc || is_mm) {
      FusedConv f;
      f.root = n.id;
      const Shape& os = val(n.outputs[0]).shape;
      const int64_t C = is_conv ? os[1] : os.back();
      f.scale.assign(C, 1.f);
      f.shift.assign(C, 0.f);
      bool dw = false;
      if (n.op == B200_OP_GROUP_CONVOLUTION) {
        const Shape& w = val(n.inputs[1]).shape;
        dw = w[1] == 1 && w[2] == 1;
      }
      int cur = n.outputs[0];
      bool pre = true;  // still allowed to fold linear per-channel ops into weights/bias
      while (true) {
        int u = single_consumer(cur);
        if (u < 0 || absorbed_by_[u] >= 0) break;
        const Node& U = node(u);
        std::vector<float> cv;
        ActSpec as;
        if (pre && !is_mm && U.op == B200_OP_BATCH_NORM && U.inputs[0] == cur && const_node(U.inputs[1]) && const_node(U.inputs[2]) &&
            const_node(U.inputs[3]) && const_node(U.inputs[4]) && is_conv) {
          auto ga = const_floats(U.inputs[1]), be = const_floats(U.inputs[2]), mu = const_floats(U.inputs[3]), va = const_floats(U.inputs[4]);
          if ((int64_t)ga.size() != C) break;
          for (int64_t c = 0; c < C; ++c) {
            float gsc = ga[c] / std::sqrt(va[c] + U.f[0]);
            f.scale[c] *= gsc;
            f.shift[c] = f.shift[c] * gsc + (be[c] - mu[c] * gsc);
          }
        } else if (pre && !is_mm && (U.op == B200_OP_MULTIPLY || U.op == B200_OP_ADD || U.op == B200_OP_SUBTRACT) &&
                   per_channel_const(U.inputs[0] == cur ? U.inputs[1] : U.inputs[0], C, !is_conv, cv) &&
                   !(U.op == B200_OP_SUBTRACT && U.inputs[0] != cur) && val(U.outputs[0]).shape == val(cur).shape) {
          for (int64_t c = 0; c < C; ++c) {
            if (U.op == B200_OP_MULTIPLY) { f.scale[c] *= cv[c]; f.shift[c] *= cv[c]; }
            else if (U.op == B200_OP_ADD) f.shift[c] += cv[c];
            else f.shift[c] -= cv[c];
          }
        } else if (act_from_op(U, as) && f.res_value < 0 && f.act1.kind == ACT_NONE && !(U.op == B200_OP_SWISH && U.inputs.size() > 1)) {
          f.act1 = as;
          pre = false;
        } else if (act_from_op(U, as) && f.res_value >= 0 && f.act2.kind == ACT_NONE && !(U.op == B200_OP_SWISH && U.inputs.size() > 1)) {
          f.act2 = as;
          f.absorbed.push_back(u);
          absorbed_by_[u] = n.id;
          cur = U.outputs[0];
          break;
        } else if (U.op == B200_OP_ADD && f.res_value < 0 && !dw && U.inputs[0] != U.inputs[1]) {
          int other = U.inputs[0] == cur ? U.inputs[1] : U.inputs[0];
          if (const_node(other) || val(other).shape != val(cur).shape || val(U.outputs[0]).shape != val(cur).shape) break;
          f.res_value = other;
          pre = false;
        } else {
          break;
        }
        f.absorbed.push_back(u);
        absorbed_by_[u] = n.id;
        cur = U.outputs[0];
      }
      // attention head split: Reshape [.., S, H*64] -> [N, S, H, 64], Transpose (0,2,1,3).  The pair kernel's epilogue stores 32-row x
      // 64-channel boxes, i.e. 32 tokens of one head: a 4-D store map writes them at [n][h][s][0..63] and the permute kernel goes away.
      if (is_fc && use_tc_ && f.res_value < 0) {
        const char* fh = getenv("B200_FOLD_HEADS");
        const bool fold_heads = !(fh && fh[0] == '0');
        const int r = single_consumer(cur);
        const int t = (fold_heads && r >= 0 && absorbed_by_[r] < 0 && node(r).op == B200_OP_RESHAPE) ? single_consumer(node(r).outputs[0]) : -1;
        if (t >= 0 && absorbed_by_[t] < 0 && node(t).op == B200_OP_TRANSPOSE) {
          const Shape& rs = val(node(r).outputs[0]).shape;
          const Node& T = node(t);
          const Shape& ws = val(n.inputs[1]).shape;
          const int64_t K = n.i[1] ? ws[1] : ws[0];
          if (rs.size() == 4 && rs[3] == 64 && rs[2] * 64 == C && rs[1] % 32 == 0 && T.i[0] == 0 && T.i[1] == 2 && T.i[2] == 1 && T.i[3] == 3 &&
              K % 64 == 0 && C <= 2048 && val(cur).shape.back() == C) {
            f.heads_s = (int)rs[1]; f.heads_h = (int)rs[2];
            f.absorbed.push_back(r); f.absorbed.push_back(t);
            absorbed_by_[r] = absorbed_by_[t] = n.id;
            cur = T.outputs[0];
          }
        }
      }
      // the inverse after the context MatMul: [N,H,S,64] -> Transpose (0,2,1,3) -> [N,S,H,64], written by the grouped launch's stores
      if (is_mm && use_tc_ && f.res_value < 0 && n.i[0] == 0) {
        const char* fh = getenv("B200_FOLD_HEADS");
        const char* flat = getenv("B200_MATMUL_FLAT");
        const int t = single_consumer(cur);
        const Shape& cs = val(cur).shape;
        const Shape& as_ = val(n.inputs[0]).shape;
        const Shape& bs = val(n.inputs[1]).shape;
        if (!(fh && fh[0] == '0') && !(flat && flat[0] == '0') && t >= 0 && absorbed_by_[t] < 0 && node(t).op == B200_OP_TRANSPOSE && cs.size() == 4 &&
            as_.size() == 4 && bs.size() == 4 && cs[3] == 64 && cs[2] % 128 == 0 && as_[3] % 64 == 0 && as_[0] == bs[0] && as_[1] == bs[1] && cs[0] * cs[1] > 1 &&
            cs[0] * cs[1] * cs[2] < (int64_t(1) << 31)) {
          const Node& T = node(t);
          if (T.i[0] == 0 && T.i[1] == 2 && T.i[2] == 1 && T.i[3] == 3) {
            f.heads_s = (int)cs[2]; f.heads_h = (int)cs[1];
            f.absorbed.push_back(t);
            absorbed_by_[t] = n.id;
            cur = T.outputs[0];
          }
        }
      }
      f.out_value = cur;
      f.emit_at = f.absorbed.empty() ? n.id : f.absorbed.back();
      absorbed_by_[n.id] = n.id;
      fconv_[f.emit_at] = f;
      continue;
    }
    if (n.op == B200_OP_MVN) {
      FusedMvn f;
      f.root = n.id;
      const Shape& xs = val(n.inputs[0]).shape;
      int cur = n.outputs[0];
      int na = n.i[2];
      // only a normalisation over the last axis can take per-element gamma/beta of that axis
      bool last_axis_only = na == 1 && ((n.i[3] < 0 ? n.i[3] + (int)xs.size() : n.i[3]) == (int)xs.size() - 1);
      if (last_axis_only) {
        int u = single_consumer(cur);
        std::vector<float> cv;
        if (u >= 0 && absorbed_by_[u] < 0 && node(u).op == B200_OP_MULTIPLY) {
          const Node& U = node(u);
          int other = U.inputs[0] == cur ? U.inputs[1] : U.inputs[0];
          if (per_channel_const(other, xs.back(), true, cv) && val(U.outputs[0]).shape == xs) {
            f.gamma_value = other; f.absorbed.push_back(u); absorbed_by_[u] = n.id; cur = U.outputs[0];
            int u2 = single_consumer(cur);
            if (u2 >= 0 && absorbed_by_[u2] < 0 && node(u2).op == B200_OP_ADD) {
              const Node& U2 = node(u2);
              int o2 = U2.inputs[0] == cur ? U2.inputs[1] : U2.inputs[0];
              if (per_channel_const(o2, xs.back(), true, cv) && val(U2.outputs[0]).shape == xs) {
                f.beta_value = o2; f.absorbed.push_back(u2); absorbed_by_[u2] = n.id; cur = U2.outputs[0];
              }
            }
          }
        }
      }
      f.out_value = cur;
      f.emit_at = f.absorbed.empty() ? n.id : f.absorbed.back();
      absorbed_by_[n.id] = n.id;
      fmvn_[f.emit_at] = f;
      continue;
    }
    if (is_binary_eltwise(n.op) || n.op == B200_OP_BATCH_NORM) {
      FusedElt f;
      f.root = n.id;
      f.out_value = n.outputs[0];
      f.emit_at = n.id;
      int u = single_consumer(n.outputs[0]);
      ActSpec as;
      if (u >= 0 && absorbed_by_[u] < 0 && act_from_op(node(u), as) && !(node(u).op == B200_OP_SWISH && node(u).inputs.size() > 1)) {
        f.act = as; f.absorbed = u; absorbed_by_[u] = n.id; f.out_value = node(u).outputs[0]; f.emit_at = u;
      }
      absorbed_by_[n.id] = n.id;
      felt_[f.emit_at] = f;
    }
  }
  fold_shortcuts();
  fold_attention();
}

// Attention: MatMul(Q, K^T) -> [x scalar] -> [+ mask] -> Softmax(last axis) -> MatMul(P, V) -> Transpose (0,2,1,3), all on [N, H, S, 64]
// tensors.  fuse() has already folded the scale / mask into the softmax record (fsm_) and the transpose into the context MatMul's stores
// (FusedConv::heads_s); this pass replaces the three kernels by the fused one when attention.cu takes the geometry.  The [N,H,S,S]
// scores then never exist in HBM.  B200_FUSED_ATTENTION=0 keeps the three kernels.
void Lowerer::fold_attention() {
  if (!use_tc_) return;
  std::vector<int> erase;
  for (auto& kv : fsm_) {
    const Node& sm = node(kv.first);
    const FusedSoftmax& fs = kv.second;
    // producer of the scores: an activation x activation MatMul with transpose_b, nothing absorbed
    const Value& xv = val(fs.x_value);
    if (xv.producer < 0 || xv.consumers.size() != 1) continue;
    FusedConv* mm1 = nullptr;
    for (auto& c : fconv_) if (c.second.out_value == fs.x_value && c.second.root == xv.producer) mm1 = &c.second;
    if (!mm1 || !mm1->absorbed.empty() || mm1->res_value >= 0) continue;
    const Node& m1 = node(mm1->root);
    if (m1.op != B200_OP_MATMUL || m1.i[0] != 0 || m1.i[1] != 1 || const_node(m1.inputs[0]) || const_node(m1.inputs[1])) continue;
    // consumer of the probabilities: MatMul(P, V), P first, no transposes
    const Value& pv = val(sm.outputs[0]);
    if (pv.consumers.size() != 1) continue;
    const Node& m2 = node(pv.consumers[0]);
    if (m2.op != B200_OP_MATMUL || m2.inputs[0] != sm.outputs[0] || m2.i[0] != 0 || m2.i[1] != 0 || const_node(m2.inputs[1])) continue;
    FusedConv* mm2 = nullptr;
    for (auto& c : fconv_) if (c.second.root == m2.id) mm2 = &c.second;
    if (!mm2 || mm2->res_value >= 0 || mm2->act1.kind != ACT_NONE || mm2->act2.kind != ACT_NONE) continue;
    const Shape& qs = val(m1.inputs[0]).shape;
    const Shape& ks = val(m1.inputs[1]).shape;
    const Shape& vs = val(m2.inputs[1]).shape;
    if (qs.size() != 4 || ks != qs || vs != qs) continue;
    // the output must be the merged-heads layout [N, S, H, 64] (Transpose folded by fuse()): that is what the kernel stores
    if (mm2->heads_s != (int)qs[2] || mm2->heads_h != (int)qs[1]) continue;
    AttnParams ap{};
    ap.N = (int)qs[0]; ap.H = (int)qs[1]; ap.S = (int)qs[2]; ap.dh = (int)qs[3];
    ap.mask_batch = 1;
    if (fs.mask_value >= 0) {
      const Shape& ms = val(fs.mask_value).shape;
      ap.mask_batch = (int)ms[0];
      ap.mask = reinterpret_cast<const void*>(1);    // "present" for the support query
    }
    if (!attention_supported(ap, cd_)) continue;
    FusedAttn fa;
    fa.q_value = m1.inputs[0]; fa.k_value = m1.inputs[1]; fa.v_value = m2.inputs[1]; fa.mask_value = fs.mask_value;
    fa.scale = fs.scale;
    fa.out_value = mm2->out_value;
    fa.fused = m1.name + "," + (fs.fused.empty() ? "" : fs.fused + ",") + sm.name + "," + m2.name;
    for (int a : mm2->absorbed) fa.fused += "," + node(a).name;
    const int emit_at = mm2->emit_at;
    erase.push_back(mm1->emit_at);
    erase.push_back(emit_at);
    absorbed_by_[sm.id] = m2.id;      // the Softmax node itself is no longer lowered on its own
    fattn_[emit_at] = fa;
  }
  for (int k : erase) fconv_.erase(k);
}

void Lowerer::lower_attention(const FusedAttn& f) {
  LTensor q = get(f.q_value, L_ROWMAJOR), k = get(f.k_value, L_ROWMAJOR), v = get(f.v_value, L_ROWMAJOR);
  LTensor mk;
  if (f.mask_value >= 0) mk = get(f.mask_value, L_ROWMAJOR);
  B200_CHECK(f.mask_value < 0 || mk.dtype == q.dtype, B200_ERR_UNSUPPORTED, "attention mask dtype differs from the scores'");
  LTensor out = alloc(val(f.out_value).shape, cd_, L_ROWMAJOR);
  AttnParams ap{};
  ap.N = (int)q.shape[0]; ap.H = (int)q.shape[1]; ap.S = (int)q.shape[2]; ap.dh = (int)q.shape[3];
  ap.mask_batch = f.mask_value >= 0 ? (int)mk.shape[0] : 1;
  ap.scale = f.scale;
  const bool has_mask = f.mask_value >= 0;
  BufRef qr = q.buf, kr = k.buf, vr = v.buf, mr = mk.buf, yr = out.buf;
  const int dt = cd_;
  Step s;
  s.name = val(f.out_value).name; s.type = "Attention"; s.impl = "tcgen05_fused_attention_s" + std::to_string(ap.S); s.fused = f.fused;
  s.flops = 4.0 * ap.N * ap.H * (double)ap.S * ap.S * ap.dh;
  s.bytes = (double)q.bytes() * 3 + (double)out.bytes();
  s.launch = [=](const RunCtx& c) {
    AttnParams p = ap;
    p.q = c.ptr(qr); p.k = c.ptr(kr); p.v = c.ptr(vr); p.mask = has_mask ? c.ptr(mr) : nullptr; p.y = c.ptr(yr);
    launch_attention(p, dt, (cudaStream_t)c.stream);
  };
  emit(std::move(s), {&q, &k, &v}, {&out});
  if (has_mask && mk.buf.space == SP_ARENA) net_.steps.back().reads.push_back(mk.buf.box);   // emit() records three reads; the mask lives on too
  set(f.out_value, out);
}

// ResNet downsampling block: out = ReLU(conv_c(b) + conv_proj(x)), both 1x1 with folded BN.  fuse() made conv_proj its own kernel and
// its output conv_c's residual: a 100-400 MB tensor written once and read once.  As ONE GEMM over concatenated K,
//   [b | x at the shortcut's stride] * [Wc ; Wproj]^T + (shift_c + shift_proj),
// the shortcut kernel, its output tensor and the residual read disappear (conv_tc3's second activation source, ConvParams::x2).
// The sum is formed in the fp32 accumulator instead of after rounding the shortcut to bf16.  B200_FOLD_SHORTCUT=0 turns it off.
void Lowerer::fold_shortcuts() {
  const char* off = getenv("B200_FOLD_SHORTCUT");
  if (!use_tc_ || (off && off[0] == '0')) return;
  auto pointwise = [&](const Node& n, bool stride1) {
    if (n.op != B200_OP_CONVOLUTION) return false;
    const Shape& w = val(n.inputs[1]).shape;
    if (w.size() != 4 || w[2] != 1 || w[3] != 1 || w[1] % 64 != 0 || w[0] % 64 != 0) return false;
    if (n.i[2] != 1 || n.i[3] != 1 || n.i[4] != 0 || n.i[5] != 0 || n.i[6] != 0 || n.i[7] != 0) return false;
    return !stride1 || (n.i[0] == 1 && n.i[1] == 1);
  };
  std::vector<int> erase;
  for (auto& kv : fconv_) {
    FusedConv& f = kv.second;
    if (f.res_value < 0 || f.dual_root >= 0 || f.act1.kind != ACT_NONE) continue;
    const Node& n = node(f.root);
    FusedConv* g = nullptr;
    for (auto& kv2 : fconv_)
      if (kv2.second.out_value == f.res_value && &kv2.second != &f) g = &kv2.second;
    if (!g || g->res_value >= 0 || g->dual_root >= 0 || g->act1.kind != ACT_NONE || g->act2.kind != ACT_NONE) continue;
    const Node& m = node(g->root);
    if (val(g->out_value).consumers.size() != 1 || val(m.outputs[0]).shape != val(n.outputs[0]).shape) continue;
    // either of the two convolutions may be the one that absorbed the Add (whichever comes later in the graph); the stride-1 one becomes
    // the main source (plain activation matrix), the other one (any stride) the second source
    const bool f_main = pointwise(n, true) && pointwise(m, false);
    const bool g_main = !f_main && pointwise(m, true) && pointwise(n, false);
    if (!f_main && !g_main) continue;
    std::string other_names = (f_main ? m : n).name;
    if (f_main) {
      for (int a : g->absorbed) other_names += "," + node(a).name;
      f.dual_root = g->root;
      f.dual_scale = g->scale;
      f.dual_shift = g->shift;
    } else {
      // f's own per-channel folds stay with its convolution (now the second source); the Add / activation nodes it absorbed stay absorbed
      f.dual_root = f.root;
      f.dual_scale = f.scale;
      f.dual_shift = f.shift;
      f.root = g->root;
      f.scale = g->scale;
      f.shift = g->shift;
      std::vector<int> ab = g->absorbed;
      ab.insert(ab.end(), f.absorbed.begin(), f.absorbed.end());
      f.absorbed = ab;
    }
    f.dual_fused = other_names;
    f.act1 = f.act2;
    f.act2 = ActSpec();
    f.res_value = -1;
    erase.push_back(g->emit_at);
  }
  for (int k : erase) fconv_.erase(k);
  fold_next_reduce();
}


// Bottleneck boundary: the expand convolution of block i (+ residual / folded shortcut, ReLU) produces Y, which is read by the reduce
// convolution of block i+1 (1x1, stride 1) and, much later, as that block's residual.  conv_tc3f computes the reduce convolution from the
// Y slabs while they sit in shared memory on their way to the TMA store, so Y is written once and read once (as the residual) instead of
// twice.  B200_FUSE_NEXT=1 turns it on (see below).
void Lowerer::fold_next_reduce() {
  // Measured on B200 (profiles/README.md, calls u-aa): correct, but at parity with / slightly slower than the two separate kernels at
  // ResNet-50 b256 (3.40 vs 3.14 ms per step): the N tile drops to 128 (TMEM: two Y accumulators + Z), the operand ring to 3 stages, and
  // every slab adds a cross-CTA hand-shake.  Off by default until it wins; B200_FUSE_NEXT=1 turns it on.
  const char* on = getenv("B200_FUSE_NEXT");
  if (!use_tc_ || !(on && on[0] == '1')) return;
  auto pointwise1 = [&](const Node& n) {
    if (n.op != B200_OP_CONVOLUTION) return false;
    const Shape& w = val(n.inputs[1]).shape;
    return w.size() == 4 && w[2] == 1 && w[3] == 1 && w[1] % 64 == 0 && w[0] % 64 == 0 && n.i[0] == 1 && n.i[1] == 1 && n.i[2] == 1 && n.i[3] == 1 &&
           n.i[4] == 0 && n.i[5] == 0 && n.i[6] == 0 && n.i[7] == 0;
  };
  std::vector<int> erase;
  for (auto& kv : fconv_) {
    FusedConv& f = kv.second;
    if (f.next_root >= 0) continue;
    const Node& n = node(f.root);
    if (!pointwise1(n)) continue;
    const int64_t Cc = val(n.inputs[1]).shape[0];
    if (Cc % 128 != 0 || Cc > 3072) continue;
    FusedConv* a = nullptr;
    for (auto& kv2 : fconv_) {
      FusedConv& c = kv2.second;
      if (&c == &f || c.res_value >= 0 || c.dual_root >= 0 || c.next_root >= 0 || c.act2.kind != ACT_NONE) continue;
      const Node& cn = node(c.root);
      if (!pointwise1(cn) || cn.inputs[0] != f.out_value) continue;
      const int64_t Ca = val(cn.inputs[1]).shape[0];
      if (Ca != 64 && Ca != 128 && Ca != 256) continue;
      bool taken = false;
      for (int e : erase) taken = taken || e == c.emit_at;
      if (!taken) { a = &c; break; }
    }
    if (!a) continue;
    const Node& an = node(a->root);
    f.next_root = a->root;
    f.next_out_value = a->out_value;
    f.next_scale = a->scale;
    f.next_shift = a->shift;
    f.next_act = a->act1;
    f.next_fused = an.name;
    for (int x : a->absorbed) f.next_fused += "," + node(x).name;
    erase.push_back(a->emit_at);
  }
  for (int k : erase) fconv_.erase(k);
}

// --------------------------------------------------------------------------------------------------
void Lowerer::conv_launch_step(const std::string& name, const std::string& type, const std::string& fused, ConvParams p, BufRef x, BufRef w,
                               BufRef bias, BufRef res, BufRef y, const LTensor* tx, const LTensor* tres, const LTensor* ty, double flops,
                               double bytes, bool tc, BufRef x2, const LTensor* tx2, BufRef w2, BufRef bias2, BufRef z, const LTensor* tz) {
  Step s;
  s.name = name; s.type = type; s.fused = fused; s.flops = flops; s.bytes = bytes;
  const int dt = cd_;
  int bn = 0, bn3 = 0;
  bool v2 = false, v3 = false, both = false, small_tune = false;
  std::string impl2, impl3;
  if (tc) {
    bn = conv_tc_pick_bn(p.Cout_p);
    const char* force_v1 = getenv("B200_CONV_TC_V1");
    const char* tune = getenv("B200_AUTOTUNE");
    const char* force3 = getenv("B200_TC3");
    v2 = conv_tc2_supported(p) && !(force_v1 && force_v1[0] == '1');
    v3 = v2 && conv_tc3_preferred(p);
    const bool chained = p.N2 > 0;   // a second GEMM rides on this one's output slabs: conv_tc3f is the only implementation
    if (chained) B200_CHECK(conv_tc3f_supported(p), B200_ERR_UNSUPPORTED, "chained convolution '" << name << "' needs conv_tc3f");
    const bool kcat = p.K2 > 0 && !chained;      // second activation source: the pair kernel is the only implementation
    if (kcat) {
      B200_CHECK(conv_tc3_supported(p), B200_ERR_UNSUPPORTED, "K-concatenated convolution '" << name << "' needs the pair kernel");
      v2 = true; v3 = true;
    }
    const bool pooled = p.pool != 0;                        // stem + max-pool fusion: the pair kernel is the only implementation
    if (pooled) {
      B200_CHECK(conv_tc3_supported(p) && conv_tc3_s2d_pool_ok(p), B200_ERR_UNSUPPORTED, "pooled stem '" << name << "' needs the pair kernel");
      v2 = true; v3 = true;
    }
    const bool heads = p.y_heads_s > 0 && !p.y_heads_inv;   // head-split stores: the pair kernel is the only implementation
    if (heads) {
      B200_CHECK(conv_tc3_supported(p), B200_ERR_UNSUPPORTED, "head-split FullyConnected '" << name << "' needs the pair kernel");
      v2 = true; v3 = true;
    }
    // large launches: both persistent kernels apply, the first request measures which one is faster for this shape
    both = v3 && !chained && !pooled && !(tune && tune[0] == '0') && !(force3 && force3[0] == '1') && ((!kcat && !heads) || conv_tc3_preferred(p));
    impl2 = std::string("tcgen05_") + (dt == B200_BF16 ? "bf16" : "f16") + "_128x" + std::to_string(bn) + "x64" +
            (v2 ? (conv_tc2_a_is_matrix(p) ? "_persistent_tma" : (conv_tc2_a_is_im2col(p) ? "_persistent_im2col" : "_persistent_gather")) : "") +
            (p.w_mn_major ? "_bmn" : "") + (p.y_heads_inv ? "_heads" : "");
    // small launches (fewer tiles than SM pairs: batch 1): the static policy keeps the single-CTA kernel, but the time is one operand-latency
    // chain per k-block, so the N tile (stages that fit, tiles in flight) and the kernel are worth measuring there too
    small_tune = v2 && !v3 && !chained && !kcat && !heads && !pooled && conv_tc3_supported(p) && !(tune && tune[0] == '0') && !(force3 && (force3[0] == '0' || force3[0] == '1'));
    if (v3 || small_tune) {
      bn3 = conv_tc3_pick_bn(p);
      impl3 = std::string("tcgen05_") + (dt == B200_BF16 ? "bf16" : "f16") + "_pair256x" + std::to_string(bn3) + "x64_" + conv_tc3_mode_name(p);
    }
    s.impl = v3 ? impl3 : impl2;
    if (chained)
      s.impl = std::string("tcgen05_") + (dt == B200_BF16 ? "bf16" : "f16") + "_pair256x128x64_tma" + (p.K2 > 0 ? "_kcat" : "") + "+1x1x" + std::to_string(p.N2);
  } else {
    s.impl = "simt_fp32_128x64x16";
  }
  const bool has_res = res.space != SP_NONE;
  const bool has_x2 = p.K2 > 0;
  const char* no_alt = getenv("B200_NO_ALTERNATE");
  // opposite to the direction in which the producer of x wrote it (kernels other than conv_tc2 write first-to-last)
  {
    int in_dir = 0;
    if (tx && tx->buf.space == SP_ARENA) { auto d = box_dir_.find(tx->buf.box); if (d != box_dir_.end()) in_dir = d->second; }
    p.reverse = (v2 && !(no_alt && no_alt[0] == '1')) ? !in_dir : 0;
    if (ty && ty->buf.space == SP_ARENA) box_dir_[ty->buf.box] = p.reverse;
  }
  // which: 0 = per the static policy, 2 = conv_tc2, 3 = conv_tc3 at its modelled N tile, 64/128/256 = conv_tc3 at that N tile,
  // 1064 = conv_tc2 at N tile 64
  auto make_launch = [=](int which) {
    return [=](const RunCtx& c) {
      ConvParams q = p;
      q.x = c.ptr(x); q.w = c.ptr(w); q.bias = reinterpret_cast<const float*>(c.ptr(bias));
      q.res = has_res ? c.ptr(res) : nullptr;
      q.y = c.ptr(y);
      q.x2 = has_x2 ? c.ptr(x2) : nullptr;
      if (p.N2 > 0) {
        q.w2 = c.ptr(w2); q.bias2 = reinterpret_cast<const float*>(c.ptr(bias2)); q.z = c.ptr(z);
        launch_conv_tc3f(q, dt, (cudaStream_t)c.stream);
        return;
      }
      const bool use3 = (which >= 3 && which < 1000) || (which == 0 && v3);
      if (tc && use3) launch_conv_tc3(q, dt, which > 3 ? which : bn3, (cudaStream_t)c.stream);
      else if (tc && v2) launch_conv_tc2(q, dt, which > 1000 ? which - 1000 : bn, (cudaStream_t)c.stream);
      else if (tc) launch_conv_tc(q, dt, bn, (cudaStream_t)c.stream);
      else launch_conv_f32(q, (cudaStream_t)c.stream);
    };
  };
  s.launch = make_launch(0);
  if (both) {
    // candidates: the pair kernel at its modelled N tile and at the other N tiles that divide the padded Cout, and the single-CTA kernel
    s.alts.emplace_back(impl3, make_launch(3));
    for (int b : {256, 128, 64}) {
      if (b == bn3 || p.Cout_alloc % b) continue;
      std::string nm = std::string("tcgen05_") + (dt == B200_BF16 ? "bf16" : "f16") + "_pair256x" + std::to_string(b) + "x64_" + conv_tc3_mode_name(p);
      s.alts.emplace_back(nm, make_launch(b));     // which > 3: explicit N tile of the pair kernel
    }
    if (!has_x2 && p.y_heads_s == 0) s.alts.emplace_back(impl2, make_launch(2));
  } else if (small_tune) {
    const std::string pre = std::string("tcgen05_") + (dt == B200_BF16 ? "bf16" : "f16");
    s.alts.emplace_back(impl2, make_launch(2));                    // the static choice first
    if (bn == 128) {
      std::string nm = impl2;
      nm.replace(nm.find("_128x128x64"), 11, "_128x64x64");
      s.alts.emplace_back(nm, make_launch(1064));
    }
    for (int b : {256, 128, 64}) {
      if (p.Cout_alloc % b) continue;
      s.alts.emplace_back(pre + "_pair256x" + std::to_string(b) + "x64_" + conv_tc3_mode_name(p), make_launch(b));
    }
  }
  emit(std::move(s), {tx, tres, tx2}, {ty, tz});
}

// MobileNetV2-style inverted-residual block: expand 1x1 (+BN, clamp) -> depthwise 3x3 (+BN, clamp) -> project 1x1 (+BN) [+ block input] as ONE
// kernel (mbconv.cu): the expanded tensor (6x the block's input at every resolution) is never written.  Also the expand-less first block
// (depthwise -> project).  Called with the FIRST convolution of the chain; the later ones are marked done.  B200_FUSE_MBCONV=0 turns it off.
bool Lowerer::try_lower_mbconv(const FusedConv& f) {
  if (!use_tc_ || (cd_ != B200_BF16 && cd_ != B200_F16)) return false;
  { const char* e = getenv("B200_FUSE_MBCONV"); if (e && e[0] == '0') return false; }
  auto plain = [&](const FusedConv& c) { return c.dual_root < 0 && c.next_root < 0 && c.heads_s == 0; };
  auto clampish = [&](const ActSpec& a) { return a.kind == ACT_NONE || a.kind == ACT_RELU || a.kind == ACT_CLAMP; };
  auto pointwise = [&](const FusedConv& c) {
    const Node& m = node(c.root);
    if (m.op != B200_OP_CONVOLUTION || !plain(c) || !const_node(m.inputs[1])) return false;
    const Shape& w = val(m.inputs[1]).shape;
    return w.size() == 4 && w[2] == 1 && w[3] == 1 && m.i[0] == 1 && m.i[1] == 1 && m.i[2] == 1 && m.i[3] == 1 && !(m.i[4] | m.i[5] | m.i[6] | m.i[7]);
  };
  auto depthwise3 = [&](const FusedConv& c) {
    const Node& m = node(c.root);
    if (m.op != B200_OP_GROUP_CONVOLUTION || !plain(c) || !const_node(m.inputs[1]) || c.res_value >= 0 || c.act2.kind != ACT_NONE || !clampish(c.act1)) return false;
    const Shape& w = val(m.inputs[1]).shape;
    return w.size() == 5 && w[1] == 1 && w[2] == 1 && w[3] == 3 && w[4] == 3 && m.i[0] == m.i[1] && (m.i[0] == 1 || m.i[0] == 2) && m.i[2] == 1 && m.i[3] == 1 &&
           m.i[4] >= 0 && m.i[4] <= 1 && m.i[5] >= 0 && m.i[5] <= 1 && m.i[6] <= 1 && m.i[7] <= 1;
  };
  // the one fused record that consumes `value` (and nothing else does)
  auto sole_consumer = [&](int value) -> const FusedConv* {
    const Value& v = val(value);
    if (v.consumers.size() != 1) return nullptr;
    for (auto& kv : fconv_)
      if (kv.second.root == v.consumers[0] && node(kv.second.root).inputs[0] == value && !mb_done_.count(kv.first)) return &kv.second;
    return nullptr;
  };
  const FusedConv *fe = nullptr, *fd = nullptr;
  if (pointwise(f)) {
    if (f.res_value >= 0 || f.act2.kind != ACT_NONE || !clampish(f.act1)) return false;
    fe = &f;
    fd = sole_consumer(f.out_value);
    if (!fd || !depthwise3(*fd)) return false;
  } else if (depthwise3(f)) {
    fd = &f;
  } else {
    return false;
  }
  const FusedConv* fp = sole_consumer(fd->out_value);
  if (!fp || !pointwise(*fp) || fp->act1.kind != ACT_NONE || fp->act2.kind != ACT_NONE) return false;
  const int in_value = node((fe ? fe : fd)->root).inputs[0];
  if (fp->res_value >= 0 && !(fe && fp->res_value == in_value)) return false;
  auto xi = vt_.find(in_value);
  if (xi == vt_.end() || xi->second.layout != L_NHWC8 || xi->second.dtype != cd_) return false;
  const LTensor x = xi->second;
  const Node& nd = node(fd->root);
  const Node& np = node(fp->root);
  const Shape& hs = val(nd.inputs[0]).shape;          // hidden tensor [N, Ch, H, W]
  const Shape os = val(np.outputs[0]).shape;          // [N, Cout, Ho, Wo]
  if (hs[2] != x.shape[2] || hs[3] != x.shape[3]) return false;
  MbParams p{};
  p.N = (int)x.shape[0]; p.H = (int)x.shape[2]; p.W = (int)x.shape[3];
  const int64_t Cin = x.shape[1], Ch = hs[1], Cout = os[1];
  p.Cin_p = (int)round_up(Cin, 8); p.Ch_p = (int)round_up(Ch, 64); p.Cout_p = (int)round_up(Cout, 8); p.Cout_alloc = (int)round_up(p.Cout_p, 16);
  p.Ho = (int)os[2]; p.Wo = (int)os[3]; p.stride = nd.i[0]; p.pt = nd.i[4]; p.pl = nd.i[5];
  p.has_expand = fe ? 1 : 0; p.has_res = fp->res_value >= 0 ? 1 : 0;
  auto bounds = [&](const ActSpec& a, float& lo, float& hi) {
    lo = a.kind == ACT_NONE ? -INFINITY : (a.kind == ACT_RELU ? 0.f : a.a);
    hi = a.kind == ACT_CLAMP ? a.b : INFINITY;
  };
  bounds(fe ? fe->act1 : ActSpec{}, p.lo1, p.hi1);
  bounds(fd->act1, p.lo2, p.hi2);
  if (val(nd.outputs[0]).shape[2] != os[2] || val(nd.outputs[0]).shape[3] != os[3] || (int64_t)(p.H + p.pt + nd.i[6] - 3) / p.stride + 1 != p.Ho) return false;
  // zero-padded hidden channels must stay zero through the expand clamp
  if (fe && Ch != p.Ch_p && !(p.lo1 <= 0.f && p.hi1 >= 0.f)) return false;
  if (!mbconv_supported(p, cd_)) return false;

  const int es = dtype_size(cd_);
  const int64_t K1 = round_up(p.Cin_p, 64);
  BufRef w1ref, b1ref;
  std::string fused;
  double flops = 0, wbytes = 0;
  if (fe) {
    const Node& ne = node(fe->root);
    std::vector<float> w = const_floats(ne.inputs[1]);                     // [Ch][Cin][1][1]
    std::vector<uint8_t> wp((size_t)(p.Ch_p * K1) * es, 0);
    for (int64_t o = 0; o < Ch; ++o)
      for (int64_t c = 0; c < Cin; ++c) store_float(wp.data(), o * K1 + c, w[o * Cin + c] * fe->scale[o], cd_);
    std::vector<float> b(p.Ch_p, 0.f);
    for (int64_t o = 0; o < Ch; ++o) b[o] = fe->shift[o];
    w1ref = weights_ref(add_weights(wp.data(), (int64_t)wp.size()));
    b1ref = add_f32_vector(b);
    fused = ne.name;
    for (int a : fe->absorbed) fused += "," + node(a).name;
    fused += ",";
    flops += 2.0 * p.N * p.H * p.W * Ch * Cin;
    wbytes += (double)wp.size();
  }
  BufRef wdref, bdref, w2ref, b2ref;
  {
    std::vector<float> w = const_floats(nd.inputs[1]);                     // [Ch][1][1][3][3]
    std::vector<uint8_t> wp((size_t)(9 * p.Ch_p) * es, 0);
    for (int64_t c = 0; c < Ch; ++c)
      for (int t = 0; t < 9; ++t) store_float(wp.data(), (int64_t)t * p.Ch_p + c, w[c * 9 + t] * fd->scale[c], cd_);
    std::vector<float> b(p.Ch_p, 0.f);
    for (int64_t c = 0; c < Ch; ++c) b[c] = fd->shift[c];
    wdref = weights_ref(add_weights(wp.data(), (int64_t)wp.size()));
    bdref = add_f32_vector(b);
    fused += nd.name;
    for (int a : fd->absorbed) fused += "," + node(a).name;
    flops += 2.0 * p.N * p.Ho * p.Wo * Ch * 9;
    wbytes += (double)wp.size();
  }
  {
    std::vector<float> w = const_floats(np.inputs[1]);                     // [Cout][Ch][1][1]
    std::vector<uint8_t> wp((size_t)(p.Cout_alloc * p.Ch_p) * es, 0);
    for (int64_t o = 0; o < Cout; ++o)
      for (int64_t c = 0; c < Ch; ++c) store_float(wp.data(), o * p.Ch_p + c, w[o * Ch + c] * fp->scale[o], cd_);
    std::vector<float> b(p.Cout_alloc, 0.f);
    for (int64_t o = 0; o < Cout; ++o) b[o] = fp->shift[o];
    w2ref = weights_ref(add_weights(wp.data(), (int64_t)wp.size()));
    b2ref = add_f32_vector(b);
    fused += "," + np.name;
    for (int a : fp->absorbed) fused += "," + node(a).name;
    flops += 2.0 * p.N * p.Ho * p.Wo * Cout * Ch;
    wbytes += (double)wp.size();
  }
  LTensor out = alloc(val(fp->out_value).shape, cd_, L_NHWC8);
  Step s;
  s.name = val(fp->out_value).name; s.type = "Convolution"; s.fused = fused;
  s.impl = std::string("tcgen05_") + (cd_ == B200_BF16 ? "bf16" : "f16") + "_mbconv_s" + std::to_string(p.stride) + (fe ? "" : "_noexpand") + (p.has_res ? "_res" : "");
  s.flops = flops;
  s.bytes = (double)x.bytes() + (double)out.bytes() + wbytes;
  BufRef xr = x.buf, yr = out.buf;
  const int dt = cd_;
  const bool has_e = fe != nullptr;
  s.launch = [=](const RunCtx& c) {
    MbParams q = p;
    q.x = c.ptr(xr); q.y = c.ptr(yr);
    q.w1 = has_e ? c.ptr(w1ref) : nullptr; q.b1 = has_e ? reinterpret_cast<const float*>(c.ptr(b1ref)) : nullptr;
    q.wd = c.ptr(wdref); q.bd = reinterpret_cast<const float*>(c.ptr(bdref));
    q.w2 = c.ptr(w2ref); q.b2 = reinterpret_cast<const float*>(c.ptr(b2ref));
    launch_mbconv(q, dt, (cudaStream_t)c.stream);
  };
  emit(std::move(s), {&x}, {&out});
  set(fp->out_value, out);
  if (fe) mb_done_.insert(fd->emit_at);
  mb_done_.insert(fp->emit_at);
  return true;
}

void Lowerer::lower_conv(const FusedConv& f) {
  const Node& n = node(f.root);
  std::string fused = n.name;
  for (int a : f.absorbed) fused += "," + node(a).name;
  const std::string out_name = val(f.out_value).name;

  if (n.op == B200_OP_MATMUL) {
    // activation x activation, or a transposed first operand (A^T * W): the batched-GEMM path (which materialises A^T); the
    // FullyConnected path below reads A as [M][K] rows and never looks at transpose_a
    if (!const_node(n.inputs[1]) || n.i[0] != 0) { lower_matmul_general(n, &f); return; }
    // ---- FullyConnected: Y[M,N] = X[M,K] * W^T (+bias, act, residual) ----
    LTensor x = get(n.inputs[0], L_ROWMAJOR);
    const Shape& ws = val(n.inputs[1]).shape;
    B200_CHECK(ws.size() == 2, B200_ERR_UNSUPPORTED, "MatMul '" << n.name << "': constant weights must be 2-D");
    const bool tb = n.i[1] != 0;
    const int64_t K = tb ? ws[1] : ws[0], Nn = tb ? ws[0] : ws[1];
    B200_CHECK(x.shape.back() == K, B200_ERR_INVALID, "MatMul '" << n.name << "': K mismatch");
    const int64_t M = numel(x.shape) / K;
    const int64_t Kp = round_up(K, 8);
    if (Kp != K) {
      // the kernels read the activation matrix in 16-byte granules (TMA row pitch): rows whose length is not a multiple of 8 elements
      // (LeNet's 84 -> 10 layer) are copied into a zero-padded scratch first; the packed weights are zero in the padding columns too
      Shape ps = {M, Kp};
      LTensor xp = alloc(ps, x.dtype, L_ROWMAJOR);
      Step s;
      s.name = out_name + "_padk"; s.type = "Reorder"; s.impl = "copy2d"; s.fused = s.name;
      BufRef src = x.buf, dst = xp.buf;
      const int64_t es = dtype_size(x.dtype), pad_bytes = xp.bytes();
      s.bytes = (double)x.bytes() + 2.0 * (double)pad_bytes;
      s.launch = [=](const RunCtx& c) {
        launch_fill_zero(c.ptr(dst), pad_bytes, (cudaStream_t)c.stream);
        launch_copy2d(c.ptr(src), K * es, c.ptr(dst), Kp * es, K * es, M, (cudaStream_t)c.stream);
      };
      emit(std::move(s), {&x}, {&xp});
      x = xp;
    }
    const bool tc = use_tc_;
    const int64_t Np = round_up(Nn, 8);
    const int bn = tc ? conv_tc_pick_bn((int)Np) : 64;
    const int64_t Nalloc = round_up(Np, bn), Ktot = round_up(K, 64);
    std::vector<float> wsrc = const_floats(n.inputs[1]);
    std::vector<uint8_t> wp((size_t)(Nalloc * Ktot) * dtype_size(cd_), 0);
    for (int64_t o = 0; o < Nn; ++o)
      for (int64_t k = 0; k < K; ++k) store_float(wp.data(), o * Ktot + k, (tb ? wsrc[o * K + k] : wsrc[k * Nn + o]) * f.scale[o], cd_);
    std::vector<float> bias(Nalloc, 0.f);
    for (int64_t o = 0; o < Nn; ++o) bias[o] = f.shift[o];
    BufRef wref = weights_ref(add_weights(wp.data(), (int64_t)wp.size()));
    BufRef bref = add_f32_vector(bias);
    Shape oshape = val(f.out_value).shape;
    LTensor out_dense = alloc(oshape, cd_, L_ROWMAJOR);
    LTensor out_k = out_dense;
    if (Nn % 8 != 0) {  // kernel rows are 16-byte granular: write a padded scratch, then compact
      Shape ps = {M, Np};
      out_k = alloc(ps, cd_, L_ROWMAJOR);
    }
    LTensor res;
    BufRef rref;
    if (f.res_value >= 0) {
      B200_CHECK(Nn % 8 == 0, B200_ERR_UNSUPPORTED, "FC residual with N % 8 != 0");
      res = get(f.res_value, L_ROWMAJOR);
      rref = res.buf;
    }
    ConvParams p;
    p.N = (int)M; p.H = p.W = p.Ho = p.Wo = 1; p.Cin_p = (int)Kp; p.Cout_p = (int)Np;
    p.M = (int)M; p.Ktot = (int)Ktot; p.Cout_alloc = (int)Nalloc; p.ldy = (int)Np;
    p.w_ld = (int)Ktot; p.w_rows = (int)Nalloc; p.w_k = (int)Ktot;
    p.act1 = f.act1.kind; p.a1 = f.act1.a; p.b1 = f.act1.b; p.act2 = f.act2.kind; p.a2 = f.act2.a; p.b2 = f.act2.b;
    p.y_heads_s = f.heads_s; p.y_heads_h = f.heads_h;
    double flops = 2.0 * M * K * Nn;
    double bytes = (double)x.bytes() + (double)out_dense.bytes() + (double)Nn * K * dtype_size(cd_) + (f.res_value >= 0 ? (double)out_dense.bytes() : 0);
    conv_launch_step(out_name, "FullyConnected", fused, p, x.buf, wref, bref, rref, out_k.buf, &x, f.res_value >= 0 ? &res : nullptr, &out_k,
                     flops, bytes, tc);
    if (Nn % 8 != 0) {
      Step s;
      s.name = out_name + "_compact"; s.type = "Reorder"; s.impl = "copy2d";
      BufRef src = out_k.buf, dst = out_dense.buf;
      int64_t es = dtype_size(cd_);
      s.launch = [=](const RunCtx& c) { launch_copy2d(c.ptr(src), Np * es, c.ptr(dst), Nn * es, Nn * es, M, (cudaStream_t)c.stream); };
      emit(std::move(s), {&out_k}, {&out_dense});
    }
    set(f.out_value, out_dense);
    return;
  }

  // ---- Convolution / GroupConvolution ----
  {
    auto xi = vt_.find(n.inputs[0]);
    if (xi != vt_.end() && xi->second.layout == L_S2D16) { lower_stem_s2d(f, xi->second); return; }
  }
  if (try_lower_mbconv(f)) return;
  LTensor x = get(n.inputs[0], L_NHWC8);
  const Shape& xs = x.shape;
  const Shape& ws = val(n.inputs[1]).shape;
  const Shape os = val(n.outputs[0]).shape;
  const bool grouped = n.op == B200_OP_GROUP_CONVOLUTION;
  const int64_t G = grouped ? ws[0] : 1;
  const int64_t Og = grouped ? ws[1] : ws[0], Ig = grouped ? ws[2] : ws[1];
  const int64_t kh = ws[grouped ? 3 : 2], kw = ws[grouped ? 4 : 3];
  const int64_t Cin = xs[1], Cout = os[1], Cin_p = round_up(Cin, 8), Cout_p = round_up(Cout, 8);
  std::vector<float> wsrc = const_floats(n.inputs[1]);
  LTensor out = alloc(val(f.out_value).shape, cd_, L_NHWC8);
  const int64_t Mrows = os[0] * os[2] * os[3];

  if (grouped && Og == 1 && Ig == 1) {
    // depthwise: weights [tap][Cp]
    std::vector<uint8_t> wp((size_t)(kh * kw * Cin_p) * dtype_size(cd_), 0);
    for (int64_t c = 0; c < Cout; ++c)
      for (int64_t t = 0; t < kh * kw; ++t) store_float(wp.data(), t * Cin_p + c, wsrc[c * kh * kw + t] * f.scale[c], cd_);
    std::vector<float> bias(Cout_p, 0.f);
    for (int64_t c = 0; c < Cout; ++c) bias[c] = f.shift[c];
    BufRef wref = weights_ref(add_weights(wp.data(), (int64_t)wp.size()));
    BufRef bref = add_f32_vector(bias);
    DwParams p{};
    p.N = (int)xs[0]; p.H = (int)xs[2]; p.W = (int)xs[3]; p.Cp = (int)Cin_p; p.Ho = (int)os[2]; p.Wo = (int)os[3];
    p.kh = (int)kh; p.kw = (int)kw; p.sh = n.i[0]; p.sw = n.i[1]; p.dh = n.i[2]; p.dw = n.i[3]; p.pt = n.i[4]; p.pl = n.i[5];
    p.act = f.act1.kind; p.a = f.act1.a; p.b = f.act1.b;
    Step s;
    s.name = out_name; s.type = "Convolution"; s.impl = "dw_nhwc8_cuda_core"; s.fused = fused;
    s.flops = 2.0 * Mrows * Cout * kh * kw;
    s.bytes = (double)x.bytes() + (double)out.bytes() + (double)wp.size();
    BufRef xr = x.buf, yr = out.buf;
    int dt = cd_;
    s.launch = [=](const RunCtx& c) {
      DwParams q = p;
      q.x = c.ptr(xr); q.w = c.ptr(wref); q.bias = reinterpret_cast<const float*>(c.ptr(bref)); q.y = c.ptr(yr);
      launch_dwconv(q, dt, (cudaStream_t)c.stream);
    };
    emit(std::move(s), {&x}, {&out});
    set(f.out_value, out);
    return;
  }

  const bool tc = use_tc_;
  const int bn = tc ? conv_tc_pick_bn((int)Cout_p) : 64;
  const int64_t Cout_alloc = round_up(Cout_p, bn);
  // K-concatenated projection shortcut (fold_shortcuts): K = [this conv's channels | the shortcut's input channels]
  const Node* dn = f.dual_root >= 0 ? &node(f.dual_root) : nullptr;
  LTensor x2;
  int64_t K2 = 0;
  if (dn) {
    x2 = get(dn->inputs[0], L_NHWC8);
    K2 = round_up(x2.shape[1], 8);
    B200_CHECK(tc && K2 % 64 == 0 && Cin_p % 64 == 0 && kh == 1 && kw == 1, B200_ERR_UNSUPPORTED, "shortcut folding: unsupported geometry");
  }
  const int64_t K1 = round_up(kh * kw * Cin_p, 64);
  const int64_t Ktot = K1 + K2;
  std::vector<uint8_t> wp((size_t)(Cout_alloc * Ktot) * dtype_size(cd_), 0);
  // K order = (filter row, filter col, padded input channel); grouped convs become block diagonal
  for (int64_t o = 0; o < Cout; ++o) {
    int64_t grp = o / Og;
    for (int64_t ig = 0; ig < Ig; ++ig) {
      int64_t ci = grp * Ig + ig;
      for (int64_t r = 0; r < kh; ++r)
        for (int64_t s2 = 0; s2 < kw; ++s2) {
          float v = grouped ? wsrc[(((grp * Og + (o % Og)) * Ig + ig) * kh + r) * kw + s2] : wsrc[((o * Ig + ig) * kh + r) * kw + s2];
          store_float(wp.data(), o * Ktot + (r * kw + s2) * Cin_p + ci, v * f.scale[o], cd_);
        }
    }
  }
  std::vector<float> bias(Cout_alloc, 0.f);
  for (int64_t o = 0; o < Cout; ++o) bias[o] = f.shift[o];
  if (dn) {
    std::vector<float> w2 = const_floats(dn->inputs[1]);   // [Cout][C2][1][1]
    const int64_t C2 = x2.shape[1];
    for (int64_t o = 0; o < Cout; ++o) {
      for (int64_t c = 0; c < C2; ++c) store_float(wp.data(), o * Ktot + K1 + c, w2[o * C2 + c] * f.dual_scale[o], cd_);
      bias[o] += f.dual_shift[o];
    }
    fused += "," + f.dual_fused;
  }
  BufRef wref = weights_ref(add_weights(wp.data(), (int64_t)wp.size()));
  BufRef bref = add_f32_vector(bias);
  LTensor res;
  BufRef rref;
  if (f.res_value >= 0) { res = get(f.res_value, L_NHWC8); rref = res.buf; }
  ConvParams p;
  if (dn) { p.K2 = (int)K2; p.H2 = (int)x2.shape[2]; p.W2 = (int)x2.shape[3]; p.sh2 = dn->i[0]; p.sw2 = dn->i[1]; }
  p.N = (int)xs[0]; p.H = (int)xs[2]; p.W = (int)xs[3]; p.Cin_p = (int)Cin_p; p.Ho = (int)os[2]; p.Wo = (int)os[3]; p.Cout_p = (int)Cout_p;
  p.kh = (int)kh; p.kw = (int)kw; p.sh = n.i[0]; p.sw = n.i[1]; p.dh = n.i[2]; p.dw = n.i[3]; p.pt = n.i[4]; p.pl = n.i[5];
  p.M = (int)Mrows; p.Ktot = (int)Ktot; p.Cout_alloc = (int)Cout_alloc; p.ldy = (int)Cout_p;
  p.w_ld = (int)Ktot; p.w_rows = (int)Cout_alloc; p.w_k = (int)Ktot;
  p.act1 = f.act1.kind; p.a1 = f.act1.a; p.b1 = f.act1.b; p.act2 = f.act2.kind; p.a2 = f.act2.a; p.b2 = f.act2.b;
  B200_CHECK(Mrows < (1ll << 31) && Mrows * Cout_p < (1ll << 40), B200_ERR_UNSUPPORTED, "conv too large");
  double flops = 2.0 * Mrows * Cout * (Cin / G) * kh * kw;
  double bytes = (double)x.bytes() + (double)out.bytes() + (double)Cout * (Cin / G) * kh * kw * dtype_size(cd_) +
                 (f.res_value >= 0 ? (double)out.bytes() : 0);
  if (dn) {
    flops += 2.0 * Mrows * Cout * x2.shape[1];
    // the shortcut reads every sh2*sw2-th pixel of x2
    bytes += (double)x2.bytes() / (dn->i[0] * dn->i[1]) + (double)Cout * x2.shape[1] * dtype_size(cd_);
  }
  // <= 4 input channels (the C = 3 first layer of MobileNet-style networks; ResNet/SSD stems are space-to-depth, above): as an implicit GEMM
  // K = taps x 8 padded channels is mostly zeros and ran in the tensor-core kernel's gather mode; a direct CUDA-core kernel is HBM-bound
  if (tc && !grouped && !dn && f.next_root < 0 && f.res_value < 0 && f.act2.kind == ACT_NONE && conv_small_cin_supported(p, (int)Cin, cd_)) {
    Step s;
    s.name = out_name; s.type = "Convolution"; s.impl = "direct_small_cin_cuda_core"; s.fused = fused; s.flops = flops; s.bytes = bytes;
    BufRef xr = x.buf, yr = out.buf;
    const int dt = cd_, cin = (int)Cin;
    s.launch = [=](const RunCtx& c) {
      ConvParams q = p;
      q.x = c.ptr(xr); q.w = c.ptr(wref); q.bias = reinterpret_cast<const float*>(c.ptr(bref)); q.y = c.ptr(yr);
      launch_conv_small_cin(q, cin, dt, (cudaStream_t)c.stream);
    };
    emit(std::move(s), {&x}, {&out});
    set(f.out_value, out);
    return;
  }
  BufRef w2ref, b2ref;
  LTensor z;
  if (f.next_root >= 0) {
    const Node& an = node(f.next_root);
    const int64_t Ca = val(an.inputs[1]).shape[0];
    B200_CHECK(tc && Cout_alloc == Cout_p && Cout == Cout_p, B200_ERR_UNSUPPORTED, "chained convolution: unsupported geometry");
    std::vector<float> wa = const_floats(an.inputs[1]);      // [Ca][Cout][1][1]
    std::vector<uint8_t> wp2((size_t)(Ca * Cout_p) * dtype_size(cd_), 0);
    for (int64_t o = 0; o < Ca; ++o)
      for (int64_t c = 0; c < Cout; ++c) store_float(wp2.data(), o * Cout_p + c, wa[o * Cout + c] * f.next_scale[o], cd_);
    std::vector<float> bias2(Ca, 0.f);
    for (int64_t o = 0; o < Ca; ++o) bias2[o] = f.next_shift[o];
    w2ref = weights_ref(add_weights(wp2.data(), (int64_t)wp2.size()));
    b2ref = add_f32_vector(bias2);
    z = alloc(val(f.next_out_value).shape, cd_, L_NHWC8);
    p.N2 = (int)Ca; p.ldz = (int)Ca;
    p.act3 = f.next_act.kind; p.a3 = f.next_act.a; p.b3 = f.next_act.b;
    flops += 2.0 * Mrows * Ca * Cout;
    bytes += (double)z.bytes() + (double)Ca * Cout * dtype_size(cd_);
    fused += "," + f.next_fused;
  }
  conv_launch_step(out_name, "Convolution", fused, p, x.buf, wref, bref, rref, out.buf, &x, f.res_value >= 0 ? &res : nullptr, &out, flops,
                   bytes, tc, dn ? x2.buf : BufRef(), dn ? &x2 : nullptr, w2ref, b2ref, z.buf, f.next_root >= 0 ? &z : nullptr);
  set(f.out_value, out);
  if (f.next_root >= 0) set(f.next_out_value, z);
}

// MatMul of two activations (attention scores / context): batched GEMM through the same kernels.
// A must be [.., M, K] row-major and B [.., N, K] (K-major); transposes are materialised with the
// permute kernel where the flags say otherwise (nodes/mkldnn_matmul_node.cpp:225-262 loops a gemm
// over the batch dims the same way).
void Lowerer::lower_matmul_general(const Node& n, const FusedConv* f) {
  LTensor a = get(n.inputs[0], L_ROWMAJOR), b = get(n.inputs[1], L_ROWMAJOR);
  B200_CHECK(a.shape.size() >= 2 && b.shape.size() >= 2, B200_ERR_UNSUPPORTED, "MatMul '" << n.name << "': rank < 2");
  auto permute_last2 = [&](const LTensor& t, const std::string& nm) {
    Shape os = t.shape;
    std::swap(os[os.size() - 1], os[os.size() - 2]);
    LTensor o = alloc(os, t.dtype, L_ROWMAJOR);
    PermuteParams pp{};
    pp.rank = (int)os.size();
    pp.n = numel(os);
    std::vector<int64_t> istr(os.size());
    int64_t st = 1;
    for (int d = (int)t.shape.size() - 1; d >= 0; --d) { istr[d] = st; st *= t.shape[d]; }
    for (size_t d = 0; d < os.size(); ++d) pp.oshape[d] = os[d];
    for (size_t d = 0; d < os.size(); ++d) pp.istride[d] = istr[d];
    std::swap(pp.istride[os.size() - 1], pp.istride[os.size() - 2]);
    Step s;
    s.name = nm; s.type = "Transpose"; s.impl = "permute";
    s.bytes = 2.0 * (double)t.bytes();
    BufRef src = t.buf, dst = o.buf;
    int es = (int)dtype_size(t.dtype);
    s.launch = [=](const RunCtx& c) { PermuteParams q = pp; q.x = c.ptr(src); q.y = c.ptr(dst); launch_permute(q, es, (cudaStream_t)c.stream); };
    emit(std::move(s), {&t}, {&o});
    return o;
  };
  if (n.i[0]) a = permute_last2(a, n.name + "_At");
  // B as [.., K, 64] (attention's V): the grouped launch reads it through an MN-major descriptor, no transpose (B200_MN_MAJOR=0 disables)
  bool mn = false;
  if (!n.i[1] && use_tc_) {
    const char* e1 = getenv("B200_MN_MAJOR");
    const char* e2 = getenv("B200_MATMUL_FLAT");
    const int64_t M0 = a.shape[a.shape.size() - 2], K0 = a.shape.back();
    const int64_t ba = numel(a.shape) / (M0 * K0), bb = numel(b.shape) / (K0 * 64);
    mn = !(e1 && e1[0] == '0') && !(e2 && e2[0] == '0') && b.shape.back() == 64 && b.shape[b.shape.size() - 2] == K0 && K0 % 64 == 0 && M0 % 128 == 0 &&
         ba > 1 && ba == bb && ba * M0 < (int64_t(1) << 31);
  }
  if (!n.i[1] && !mn) b = permute_last2(b, n.name + "_Bt");   // need [.., N, K]
  const int64_t M = a.shape[a.shape.size() - 2], K = a.shape.back(), Nn = mn ? b.shape.back() : b.shape[b.shape.size() - 2];
  B200_CHECK((mn ? b.shape[b.shape.size() - 2] : b.shape.back()) == K, B200_ERR_INVALID, "MatMul '" << n.name << "': K mismatch");
  const int64_t batch_a = numel(a.shape) / (M * K), batch_b = numel(b.shape) / (Nn * K);
  B200_CHECK(batch_a == batch_b || batch_b == 1, B200_ERR_UNSUPPORTED, "MatMul '" << n.name << "': unsupported batch broadcast");
  if (batch_b != 1) {
    // equal products are not enough: A batch [2,1] x B batch [1,2] has 4 output batches.  Only identical batch dims are taken.
    Shape ba(a.shape.begin(), a.shape.end() - 2), bb(b.shape.begin(), b.shape.end() - 2);
    while (!ba.empty() && ba.front() == 1) ba.erase(ba.begin());
    while (!bb.empty() && bb.front() == 1) bb.erase(bb.begin());
    B200_CHECK(ba == bb, B200_ERR_UNSUPPORTED, "MatMul '" << n.name << "': broadcast between batch dimensions is not supported");
  }
  B200_CHECK(batch_a <= 65535, B200_ERR_UNSUPPORTED, "MatMul batch too large");
  // The kernels read operand rows and write output rows in 16-byte granules.  A reduction length or an output width that is not a multiple
  // of 8 (attention over a sequence of 33 tokens: QK^T has N = 33, PV has K = 33) goes through zero-padded copies of the operands and a
  // padded output that is compacted afterwards; found by the randomized sequence-model tests.
  const int64_t Kp = round_up(K, 8), Np = round_up(Nn, 8);
  B200_CHECK(!mn || (Kp == K && Np == Nn), B200_ERR_UNSUPPORTED, "MatMul '" << n.name << "': MN-major weights need aligned sizes");
  auto pad_rows = [&](const LTensor& t, int64_t groups, int64_t rows, int64_t new_rows, int64_t k, int64_t new_k, const std::string& nm) {
    // [groups][rows][k] -> [groups][new_rows][new_k], zero filled
    Shape ps = {groups, new_rows, new_k};
    LTensor o = alloc(ps, t.dtype, L_ROWMAJOR);
    Step s;
    s.name = nm; s.type = "Reorder"; s.impl = "copy2d"; s.fused = nm;
    BufRef src = t.buf, dst = o.buf;
    const int64_t es = dtype_size(t.dtype), total = o.bytes();
    s.bytes = (double)t.bytes() + 2.0 * (double)total;
    s.launch = [=](const RunCtx& c) {
      launch_fill_zero(c.ptr(dst), total, (cudaStream_t)c.stream);
      if (new_k == k) {   // whole groups are contiguous runs of rows * k elements
        launch_copy2d(c.ptr(src), rows * k * es, c.ptr(dst), new_rows * new_k * es, rows * k * es, groups, (cudaStream_t)c.stream);
      } else {
        for (int64_t g = 0; g < groups; ++g)
          launch_copy2d(c.ptr(src) + g * rows * k * es, k * es, c.ptr(dst) + g * new_rows * new_k * es, new_k * es, k * es, rows, (cudaStream_t)c.stream);
      }
    };
    emit(std::move(s), {&t}, {&o});
    return o;
  };
  if (Kp != K) a = pad_rows(a, batch_a, M, M, K, Kp, n.name + "_padA");
  if (Kp != K || Np != Nn) b = pad_rows(b, batch_b, Nn, Np, K, Kp, n.name + "_padB");
  const int out_value = f ? f->out_value : n.outputs[0];
  LTensor out_dense = alloc(val(out_value).shape, cd_, L_ROWMAJOR);
  LTensor out = out_dense;
  if (Np != Nn) {
    B200_CHECK(!(f && (f->res_value >= 0 || f->heads_s)), B200_ERR_UNSUPPORTED, "MatMul '" << n.name << "': fused residual / head merge with N % 8 != 0");
    Shape ps = {batch_a, M, Np};
    out = alloc(ps, cd_, L_ROWMAJOR);
  }
  const bool tc = use_tc_;
  const int bn = tc ? conv_tc_pick_bn((int)Np) : 64;
  std::vector<float> bias(round_up(Np, bn), 0.f);
  BufRef bref = add_f32_vector(bias);
  ConvParams p;
  p.N = (int)M; p.H = p.W = p.Ho = p.Wo = 1; p.Cin_p = (int)Kp; p.Cout_p = (int)Np;
  p.M = (int)M; p.Ktot = (int)round_up(Kp, 64); p.Cout_alloc = (int)round_up(Np, bn); p.ldy = (int)Np;
  p.w_ld = (int)Kp; p.w_rows = (int)Np; p.w_k = (int)Kp;
  p.batch = (int)batch_a; p.x_bs = M * Kp; p.w_bs = batch_b == 1 ? 0 : Np * Kp; p.y_bs = M * Np;
  p.w_dynamic = 1;
  {
    // per-batch [M][K] x [N][K]^T with tile-aligned M: the batches are one flat [batch*M][K] activation matrix whose 128-row tiles each
    // pick their batch's weight rows -- one persistent launch with TMA stores instead of batch x tiles single-tile CTAs
    const char* flat = getenv("B200_MATMUL_FLAT");
    if (tc && batch_a > 1 && batch_b == batch_a && M % 128 == 0 && K % 64 == 0 && Nn % 64 == 0 && Nn % bn == 0 && batch_a * M < (int64_t(1) << 31) &&
        !(flat && flat[0] == '0')) {
      p.N = p.M = (int)(batch_a * M);
      p.batch = 1; p.x_bs = p.w_bs = p.y_bs = 0;
      p.w_grp_m = (int)M; p.w_grp_rows = (int)Nn;
      p.w_mn_major = mn ? 1 : 0;
    }
    B200_CHECK(!mn || p.w_mn_major, B200_ERR_UNSUPPORTED, "MatMul '" << n.name << "': MN-major weights need the grouped launch");
    if (f && f->heads_s) {
      B200_CHECK(p.w_grp_m == f->heads_s && Nn == 64, B200_ERR_UNSUPPORTED, "MatMul '" << n.name << "': merged-heads store needs the grouped launch");
      p.y_heads_s = f->heads_s; p.y_heads_h = f->heads_h; p.y_heads_inv = 1;
    }
  }
  LTensor res;
  BufRef rref;
  std::string fused = n.name;
  if (f) {
    p.act1 = f->act1.kind; p.a1 = f->act1.a; p.b1 = f->act1.b; p.act2 = f->act2.kind; p.a2 = f->act2.a; p.b2 = f->act2.b;
    if (f->res_value >= 0) { res = get(f->res_value, L_ROWMAJOR); rref = res.buf; }
    for (int ab : f->absorbed) fused += "," + node(ab).name;
  }
  double flops = 2.0 * batch_a * M * K * Nn;
  double bytes = (double)a.bytes() + (double)b.bytes() + (double)out.bytes();
  conv_launch_step(val(out_value).name, "MatMul", fused, p, a.buf, b.buf, bref, rref, out.buf, &a, (f && f->res_value >= 0) ? &res : &b, &out,
                   flops, bytes, tc);
  // conv_launch_step only records one extra read; make sure b stays alive too
  net_.steps.back().reads.push_back(b.buf.box);
  if (Np != Nn) {
    Step s;
    s.name = val(out_value).name + "_compact"; s.type = "Reorder"; s.impl = "copy2d"; s.fused = s.name;
    BufRef src = out.buf, dst = out_dense.buf;
    const int64_t es = dtype_size(cd_), rows = batch_a * M;
    s.bytes = (double)out.bytes() + (double)out_dense.bytes();
    s.launch = [=](const RunCtx& c) { launch_copy2d(c.ptr(src), Np * es, c.ptr(dst), Nn * es, Nn * es, rows, (cudaStream_t)c.stream); };
    emit(std::move(s), {&out}, {&out_dense});
  }
  set(out_value, out_dense);
}

void Lowerer::lower_binary(const Node& n, const ActSpec& act, int out_value) {
  const Shape& os = val(n.outputs[0]).shape;
  if (n.op == B200_OP_BATCH_NORM) {
    // standalone BatchNormInference: per-channel scale/shift (batch_norm_decomposition.cpp:33-80)
    auto ga = const_floats(n.inputs[1]), be = const_floats(n.inputs[2]), mu = const_floats(n.inputs[3]), va = const_floats(n.inputs[4]);
    B200_CHECK(os.size() == 4, B200_ERR_UNSUPPORTED, "BatchNorm '" << n.name << "': rank 4 only");
    int64_t C = os[1], Cp = round_up(C, 8);
    std::vector<float> sc(Cp, 0.f), sh(Cp, 0.f);
    for (int64_t c = 0; c < C; ++c) { sc[c] = ga[c] / std::sqrt(va[c] + n.f[0]); sh[c] = be[c] - mu[c] * sc[c]; }
    LTensor x = get(n.inputs[0], L_NHWC8);
    LTensor out = alloc(os, cd_, L_NHWC8);
    BufRef s1 = add_f32_vector(sc), s2 = add_f32_vector(sh), xr = x.buf, yr = out.buf;
    int64_t rows = os[0] * os[2] * os[3];
    int dt = cd_;
    Step s;
    s.name = val(out_value).name; s.type = "Eltwise"; s.impl = "scale_shift_nhwc8"; s.fused = n.name;
    s.bytes = (double)x.bytes() + (double)out.bytes();
    s.launch = [=](const RunCtx& c) {
      launch_scale_shift(c.ptr(xr), (const float*)c.ptr(s1), (const float*)c.ptr(s2), c.ptr(yr), rows, (int)Cp, dt, act.kind, act.a, act.b,
                         (cudaStream_t)c.stream);
    };
    emit(std::move(s), {&x}, {&out});
    set(out_value, out);
    return;
  }
  const Value& va = val(n.inputs[0]);
  const Value& vb = val(n.inputs[1]);
  const bool ints = !dtype_is_float(va.dtype);
  B200_CHECK(!ints, B200_ERR_UNSUPPORTED, "integer eltwise '" << n.name << "' not supported");
  const bool ca = const_node(n.inputs[0]) != nullptr, cb = const_node(n.inputs[1]) != nullptr;
  // (1) NHWC8 tensor (op) per-channel constant -> scale/shift kernel
  if (os.size() == 4 && (ca != cb) && (n.op == B200_OP_ADD || n.op == B200_OP_MULTIPLY || (n.op == B200_OP_SUBTRACT && cb))) {
    std::vector<float> cv;
    int tv = ca ? n.inputs[1] : n.inputs[0], cvv = ca ? n.inputs[0] : n.inputs[1];
    auto it = vt_.find(tv);
    if (it != vt_.end() && it->second.layout == L_NHWC8 && val(tv).shape == os && per_channel_const(cvv, os[1], false, cv)) {
      int64_t C = os[1], Cp = round_up(C, 8);
      std::vector<float> sc(Cp, 0.f), sh(Cp, 0.f);
      for (int64_t c = 0; c < C; ++c) {
        if (n.op == B200_OP_MULTIPLY) { sc[c] = cv[c]; sh[c] = 0.f; }
        else if (n.op == B200_OP_ADD) { sc[c] = 1.f; sh[c] = cv[c]; }
        else { sc[c] = 1.f; sh[c] = -cv[c]; }
      }
      LTensor x = it->second;
      LTensor out = alloc(os, cd_, L_NHWC8);
      BufRef s1 = add_f32_vector(sc), s2 = add_f32_vector(sh), xr = x.buf, yr = out.buf;
      int64_t rows = os[0] * os[2] * os[3];
      int dt = cd_;
      Step s;
      s.name = val(out_value).name; s.type = "Eltwise"; s.impl = "scale_shift_nhwc8"; s.fused = n.name;
      s.bytes = (double)x.bytes() + (double)out.bytes();
      s.launch = [=](const RunCtx& c) {
        launch_scale_shift(c.ptr(xr), (const float*)c.ptr(s1), (const float*)c.ptr(s2), c.ptr(yr), rows, (int)Cp, dt, act.kind, act.a, act.b,
                           (cudaStream_t)c.stream);
      };
      emit(std::move(s), {&x}, {&out});
      set(out_value, out);
      return;
    }
  }
  // (2) same-shape activations: operate on the physical buffers in a common layout
  if (!ca && !cb && va.shape == vb.shape && va.shape == os) {
    int layout = L_ROWMAJOR;
    auto ia = vt_.find(n.inputs[0]), ib = vt_.find(n.inputs[1]);
    bool pad_unsafe = (n.op == B200_OP_DIVIDE || n.op == B200_OP_POWER) && os.size() == 4 && os[1] % 8 != 0;
    if (!pad_unsafe && ((ia != vt_.end() && ia->second.layout == L_NHWC8) || (ib != vt_.end() && ib->second.layout == L_NHWC8))) layout = L_NHWC8;
    LTensor a = get(n.inputs[0], layout), b = get(n.inputs[1], layout);
    LTensor out = alloc(os, cd_, layout);
    BufRef ar = a.buf, br = b.buf, yr = out.buf;
    int64_t cnt = out.phys_elems();
    int dt = cd_, op = n.op;
    Step s;
    s.name = val(out_value).name; s.type = "Eltwise"; s.impl = "binary_vec8"; s.fused = n.name;
    s.bytes = 3.0 * (double)out.bytes();
    s.launch = [=](const RunCtx& c) { launch_binary_same(c.ptr(ar), c.ptr(br), c.ptr(yr), cnt, dt, op, act.kind, act.a, act.b, (cudaStream_t)c.stream); };
    emit(std::move(s), {&a, &b}, {&out});
    set(out_value, out);
    return;
  }
  // (3) general numpy broadcast over dense row-major tensors
  LTensor a = get(n.inputs[0], L_ROWMAJOR), b = get(n.inputs[1], L_ROWMAJOR);
  LTensor out = alloc(os, cd_, L_ROWMAJOR);
  BinaryParams bp{};
  bp.rank = (int)os.size();
  B200_CHECK(bp.rank <= 8, B200_ERR_UNSUPPORTED, "eltwise rank > 8");
  bp.n = numel(os);
  auto fill = [&](const Shape& s, int64_t* strides) {
    int64_t st = 1;
    int off = (int)os.size() - (int)s.size();
    for (int d = (int)os.size() - 1; d >= 0; --d) {
      if (d - off >= 0) {
        int64_t dim = s[d - off];
        strides[d] = dim == 1 ? 0 : st;
        st *= dim;
      } else strides[d] = 0;
    }
  };
  for (int d = 0; d < bp.rank; ++d) bp.oshape[d] = os[d];
  fill(a.shape, bp.astride);
  fill(n.op == B200_OP_PRELU ? prelu_slope_shape(a.shape, b.shape) : b.shape, bp.bstride);   // slope [C] is per-channel
  bp.op = n.op; bp.act = act.kind; bp.aa = act.a; bp.ab = act.b;
  BufRef ar = a.buf, br = b.buf, yr = out.buf;
  int dt = cd_;
  Step s;
  s.name = val(out_value).name; s.type = "Eltwise"; s.impl = "binary_broadcast"; s.fused = n.name;
  s.bytes = (double)a.bytes() + (double)b.bytes() + (double)out.bytes();
  s.launch = [=](const RunCtx& c) { BinaryParams q = bp; q.a = c.ptr(ar); q.b = c.ptr(br); q.y = c.ptr(yr); launch_binary(q, dt, (cudaStream_t)c.stream); };
  emit(std::move(s), {&a, &b}, {&out});
  set(out_value, out);
}

void Lowerer::lower_unary(const Node& n, const ActSpec& act) {
  LTensor x = get(n.inputs[0], -1);
  B200_CHECK(dtype_is_float(x.dtype), B200_ERR_UNSUPPORTED, "activation on non-float tensor");
  ActSpec a = act;
  if (n.op == B200_OP_SWISH && n.inputs.size() > 1) {
    auto bv = const_floats(n.inputs[1]);  // beta as second input must be constant
    a.a = bv.empty() ? 1.f : bv[0];
  }
  LTensor out = alloc(x.shape, x.dtype, x.layout);
  BufRef xr = x.buf, yr = out.buf;
  int64_t cnt = out.phys_elems();
  int dt = x.dtype;
  Step s;
  s.name = n.name; s.type = "Eltwise"; s.impl = "unary_vec8"; s.fused = n.name;
  s.bytes = 2.0 * (double)out.bytes();
  s.launch = [=](const RunCtx& c) { launch_unary(c.ptr(xr), c.ptr(yr), cnt, dt, a.kind, a.a, a.b, (cudaStream_t)c.stream); };
  emit(std::move(s), {&x}, {&out});
  set(n.outputs[0], out);
}

// FakeQuantize as one element-wise kernel over whatever layout the input has (the reference CPU plugin's FakeQuantize node in its fp32 form,
// nodes/mkldnn_fake_quantize_node.cpp; the INT8 execution of LPT-transformed graphs is not built).  The four range constants become one
// [4][channels] fp32 table in the weight blob; NHWC8 tensors index it by the padded innermost channel, row-major ones by axis 1.
void Lowerer::lower_fake_quantize(const Node& n) {
  LTensor x = get(n.inputs[0], -1);
  B200_CHECK(dtype_is_float(x.dtype), B200_ERR_UNSUPPORTED, "FakeQuantize on a non-float tensor");
  const int64_t C = x.shape.size() >= 2 ? x.shape[1] : 1;
  std::vector<float> r[4];
  bool per_channel = false;
  for (int k = 0; k < 4; ++k) {
    B200_CHECK(const_node(n.inputs[1 + k]), B200_ERR_UNSUPPORTED, "FakeQuantize '" << n.name << "': ranges must be constants");
    r[k] = const_floats(n.inputs[1 + k]);
    per_channel = per_channel || r[k].size() > 1;
  }
  int cdim = 1;
  int64_t inner = 1;
  if (per_channel) {
    if (x.layout == L_NHWC8) { cdim = (int)round_up(C, 8); inner = 1; }
    else { cdim = (int)C; inner = 1; for (size_t d = 2; d < x.shape.size(); ++d) inner *= x.shape[d]; }
  }
  std::vector<float> table((size_t)4 * cdim, 0.f);
  for (int k = 0; k < 4; ++k)
    for (int c = 0; c < cdim; ++c) table[(size_t)k * cdim + c] = c < C ? (r[k].size() > 1 ? r[k][c] : r[k][0]) : r[k][0];
  // padded channels of an NHWC8 tensor must stay what they were for the next kernel: FakeQuantize(0) with these ranges (any value is fine:
  // every consumer ignores the padding or multiplies it by zero weights)
  BufRef tref = add_f32_vector(table);
  LTensor out = alloc(x.shape, x.dtype, x.layout);
  BufRef xr = x.buf, yr = out.buf;
  const int64_t cnt = out.phys_elems();
  const int dt = x.dtype, levels = n.i[0];
  Step s;
  s.name = n.name; s.type = "FakeQuantize"; s.impl = per_channel ? "fake_quantize_per_channel" : "fake_quantize_per_tensor"; s.fused = n.name;
  s.bytes = 2.0 * (double)out.bytes();
  s.launch = [=](const RunCtx& c) {
    launch_fake_quantize(c.ptr(xr), c.ptr(yr), cnt, dt, reinterpret_cast<const float*>(c.ptr(tref)), cdim, inner, levels, (cudaStream_t)c.stream);
  };
  emit(std::move(s), {&x}, {&out});
  set(n.outputs[0], out);
}

void Lowerer::lower_pool(const Node& n) {
  LTensor x = get(n.inputs[0], L_NHWC8);
  const Shape os = val(n.outputs[0]).shape;
  LTensor out = alloc(os, x.dtype, L_NHWC8);
  const bool avg = n.op == B200_OP_AVG_POOL;
  const int64_t Cp = round_up(os[1], 8);
  BufRef xr = x.buf, yr = out.buf;
  int dt = x.dtype;
  Step s;
  s.name = n.name; s.type = "Pooling"; s.fused = n.name;
  s.bytes = (double)x.bytes() + (double)out.bytes();
  const bool global = avg && n.i[0] == x.shape[2] && n.i[1] == x.shape[3] && !(n.i[4] | n.i[5] | n.i[6] | n.i[7]) && os[2] == 1 && os[3] == 1;
  if (global) {
    s.impl = "global_avg_nhwc8";
    int N = (int)x.shape[0], HW = (int)(x.shape[2] * x.shape[3]);
    s.launch = [=](const RunCtx& c) { launch_global_avgpool(c.ptr(xr), c.ptr(yr), N, HW, (int)Cp, dt, (cudaStream_t)c.stream); };
  } else {
    PoolParams p{};
    p.N = (int)x.shape[0]; p.H = (int)x.shape[2]; p.W = (int)x.shape[3]; p.Cp = (int)Cp; p.Ho = (int)os[2]; p.Wo = (int)os[3];
    p.kh = n.i[0]; p.kw = n.i[1]; p.sh = n.i[2]; p.sw = n.i[3]; p.pt = n.i[4]; p.pl = n.i[5]; p.pb = n.i[6]; p.pr = n.i[7];
    p.is_avg = avg; p.exclude_pad = avg ? n.i[9] : 0;
    s.impl = avg ? "avg_nhwc8" : "max_nhwc8";
    s.launch = [=](const RunCtx& c) { PoolParams q = p; q.x = c.ptr(xr); q.y = c.ptr(yr); launch_pool(q, dt, (cudaStream_t)c.stream); };
  }
  emit(std::move(s), {&x}, {&out});
  set(n.outputs[0], out);
}

void Lowerer::lower_reduce(const Node& n) {
  const Shape& xs = val(n.inputs[0]).shape;
  const Shape os = val(n.outputs[0]).shape;
  const bool keep = n.i[0] != 0;
  const int na = n.i[1];
  std::vector<int> axes;
  for (int k = 0; k < na; ++k) axes.push_back(n.i[2 + k] < 0 ? n.i[2 + k] + (int)xs.size() : n.i[2 + k]);
  std::sort(axes.begin(), axes.end());
  const int kind = n.op == B200_OP_REDUCE_MEAN ? 0 : (n.op == B200_OP_REDUCE_SUM ? 1 : 2);
  auto it = vt_.find(n.inputs[0]);
  // spatial mean of an NHWC8 tensor = global average pooling
  if (kind == 0 && xs.size() == 4 && axes == std::vector<int>({2, 3}) && it != vt_.end() && it->second.layout == L_NHWC8) {
    LTensor x = it->second;
    Shape s4 = {xs[0], xs[1], 1, 1};
    LTensor o4 = alloc(s4, x.dtype, L_NHWC8);
    BufRef xr = x.buf, yr = o4.buf;
    int N = (int)xs[0], HW = (int)(xs[2] * xs[3]), Cp = (int)round_up(xs[1], 8), dt = x.dtype;
    Step s;
    s.name = n.name; s.type = "Pooling"; s.impl = "global_avg_nhwc8"; s.fused = n.name;
    s.bytes = (double)x.bytes() + (double)o4.bytes();
    s.launch = [=](const RunCtx& c) { launch_global_avgpool(c.ptr(xr), c.ptr(yr), N, HW, Cp, dt, (cudaStream_t)c.stream); };
    emit(std::move(s), {&x}, {&o4});
    if (keep) { set(n.outputs[0], o4); return; }
    // drop the unit dims: go through the dense view
    set(n.outputs[0], o4);
    LTensor d = get(n.outputs[0], L_ROWMAJOR);
    alt_.erase(std::make_pair(n.outputs[0], (int)L_ROWMAJOR));
    d.shape = os;
    set(n.outputs[0], d);
    return;
  }
  // contiguous runs of axes over a dense tensor, the last run first (axes {1, 3} of NCHW: W, then C): one strided reduction per run;
  // sum / max / mean compose across runs (every group of a mean has the same size)
  B200_CHECK(!axes.empty(), B200_ERR_INVALID, "Reduce without axes");
  std::vector<std::pair<int, int>> runs;
  for (int a : axes) {
    if (!runs.empty() && runs.back().second + 1 == a) runs.back().second = a;
    else runs.emplace_back(a, a);
  }
  LTensor cur = get(n.inputs[0], L_ROWMAJOR);
  Shape cs = xs;
  for (int r = (int)runs.size() - 1; r >= 0; --r) {
    int64_t outer = 1, ax = 1, inner = 1;
    for (int d = 0; d < (int)cs.size(); ++d) {
      if (d < runs[r].first) outer *= cs[d];
      else if (d <= runs[r].second) ax *= cs[d];
      else inner *= cs[d];
    }
    for (int d = runs[r].first; d <= runs[r].second; ++d) cs[d] = 1;
    LTensor out = alloc(r == 0 ? os : cs, cur.dtype, L_ROWMAJOR);
    BufRef xr = cur.buf, yr = out.buf;
    int dt = cur.dtype;
    Step s;
    s.name = r == 0 ? n.name : n.name + "_run" + std::to_string(r); s.type = "Reduce"; s.impl = "reduce_strided"; s.fused = n.name;
    s.bytes = (double)cur.bytes() + (double)out.bytes();
    s.launch = [=](const RunCtx& c) { launch_reduce(c.ptr(xr), c.ptr(yr), outer, ax, inner, dt, kind, (cudaStream_t)c.stream); };
    emit(std::move(s), {&cur}, {&out});
    cur = out;
  }
  set(n.outputs[0], cur);
}

void Lowerer::lower_mvn(const Node& n, const FusedMvn* f) {
  const Shape& xs = val(n.inputs[0]).shape;
  const int na = n.i[2];
  std::vector<int> axes;
  for (int k = 0; k < na; ++k) axes.push_back(n.i[3 + k] < 0 ? n.i[3 + k] + (int)xs.size() : n.i[3 + k]);
  std::sort(axes.begin(), axes.end());
  B200_CHECK(!axes.empty() && axes.back() == (int)xs.size() - 1, B200_ERR_UNSUPPORTED, "MVN '" << n.name << "': axes must be trailing");
  for (size_t k = 1; k < axes.size(); ++k) B200_CHECK(axes[k] == axes[k - 1] + 1, B200_ERR_UNSUPPORTED, "MVN '" << n.name << "': axes must be contiguous");
  int64_t cols = 1;
  for (int a : axes) cols *= xs[a];
  const int64_t rows = numel(xs) / cols;
  LTensor x = get(n.inputs[0], L_ROWMAJOR);
  const int out_value = f ? f->out_value : n.outputs[0];
  LTensor out = alloc(xs, x.dtype, L_ROWMAJOR);
  BufRef gr, br;
  std::string fused = n.name;
  if (f && f->gamma_value >= 0) { std::vector<float> g; per_channel_const(f->gamma_value, cols, true, g); gr = add_f32_vector(g); }
  if (f && f->beta_value >= 0) { std::vector<float> b; per_channel_const(f->beta_value, cols, true, b); br = add_f32_vector(b); }
  if (f) for (int a : f->absorbed) fused += "," + node(a).name;
  BufRef xr = x.buf, yr = out.buf;
  int dt = x.dtype, nv = n.i[0], eo = n.i[1];
  float eps = n.f[0];
  Step s;
  s.name = val(out_value).name; s.type = "MVN"; s.impl = "mvn_rows"; s.fused = fused;
  s.bytes = 2.0 * (double)x.bytes();
  s.launch = [=](const RunCtx& c) {
    launch_mvn(c.ptr(xr), c.ptr(yr), gr.space ? (const float*)c.ptr(gr) : nullptr, br.space ? (const float*)c.ptr(br) : nullptr, rows, cols, nv, eo,
               eps, dt, (cudaStream_t)c.stream);
  };
  emit(std::move(s), {&x}, {&out});
  set(out_value, out);
}

void Lowerer::lower_concat(const Node& n) {
  const Shape os = val(n.outputs[0]).shape;
  int ax = n.i[0] < 0 ? n.i[0] + (int)os.size() : n.i[0];
  int sdt = storage_dtype(val(n.outputs[0]).dtype);
  const int64_t es = dtype_size(sdt);
  // channel concat of NHWC8 tensors whose channel counts are multiples of 8: pure strided copies
  bool nhwc = os.size() == 4 && ax == 1;
  if (nhwc)
    for (int v : n.inputs) {
      auto it = vt_.find(v);
      if (it == vt_.end() || it->second.layout != L_NHWC8 || val(v).shape[1] % 8 != 0) nhwc = false;
    }
  const int layout = nhwc ? L_NHWC8 : L_ROWMAJOR;
  LTensor out = alloc(os, sdt, layout);
  int64_t outer = 1, inner = 1;
  if (nhwc) { outer = os[0] * os[2] * os[3]; inner = 1; }
  else {
    for (int d = 0; d < ax; ++d) outer *= os[d];
    for (int d = ax + 1; d < (int)os.size(); ++d) inner *= os[d];
  }
  const int64_t dpitch = os[ax] * inner * es;
  int64_t off = 0;
  for (int v : n.inputs) {
    LTensor x = get(v, layout);
    const int64_t row_bytes = val(v).shape[ax] * inner * es;
    BufRef xr = x.buf, yr = out.buf;
    Step s;
    s.name = n.name; s.type = "Concat"; s.impl = "copy2d"; s.fused = n.name;
    s.bytes = 2.0 * (double)x.bytes();
    int64_t o = off;
    s.launch = [=](const RunCtx& c) { launch_copy2d(c.ptr(xr), row_bytes, c.ptr(yr) + o, dpitch, row_bytes, outer, (cudaStream_t)c.stream); };
    emit(std::move(s), {&x}, {&out});
    off += row_bytes;
  }
  set(n.outputs[0], out);
}

void Lowerer::lower_split(const Node& n) {
  LTensor x = get(n.inputs[0], L_ROWMAJOR);
  const Shape& xs = x.shape;
  int ax = n.i[0] < 0 ? n.i[0] + (int)xs.size() : n.i[0];
  const int64_t es = dtype_size(x.dtype);
  int64_t outer = 1, inner = 1;
  for (int d = 0; d < ax; ++d) outer *= xs[d];
  for (int d = ax + 1; d < (int)xs.size(); ++d) inner *= xs[d];
  const int64_t spitch = xs[ax] * inner * es;
  int64_t off = 0;
  for (size_t k = 0; k < n.outputs.size(); ++k) {
    const Shape os = val(n.outputs[k]).shape;
    LTensor out = alloc(os, x.dtype, L_ROWMAJOR);
    const int64_t row_bytes = os[ax] * inner * es;
    BufRef xr = x.buf, yr = out.buf;
    int64_t o = off;
    Step s;
    s.name = val(n.outputs[k]).name; s.type = "Split"; s.impl = "copy2d"; s.fused = n.name;
    s.bytes = 2.0 * (double)out.bytes();
    s.launch = [=](const RunCtx& c) { launch_copy2d(c.ptr(xr) + o, spitch, c.ptr(yr), row_bytes, row_bytes, outer, (cudaStream_t)c.stream); };
    emit(std::move(s), {&x}, {&out});
    set(n.outputs[k], out);
    off += row_bytes;
  }
}

// StridedSlice (mkldnn_strided_slice_node.cpp; ngraph reference/strided_slice.hpp via slice_plan): a strided gather of a dense tensor -
// the permute kernel with the source pointer moved to the first element and per-dimension source strides (negative = reversed)
void Lowerer::lower_strided_slice(const Node& n) {
  const Shape os = val(n.outputs[0]).shape;
  const int r = n.i[0];
  LTensor x = get(n.inputs[0], L_ROWMAJOR);
  std::vector<int64_t> xstr(r);
  int64_t st = 1;
  for (int d = r - 1; d >= 0; --d) { xstr[d] = st; st *= n.i[1 + d]; }
  int64_t first = 0;
  bool whole = true;
  for (int d = 0; d < r; ++d) {
    first += (int64_t)n.i[1 + r + d] * xstr[d];
    whole = whole && n.i[1 + r + d] == 0 && n.i[1 + 2 * r + d] == 1 && n.i[1 + 3 * r + d] == n.i[1 + d];
  }
  if (whole) { LTensor v = x; v.shape = os; set(n.outputs[0], v); return; }
  LTensor out = alloc(os, x.dtype, L_ROWMAJOR);
  PermuteParams pp{};
  pp.rank = r;
  pp.n = numel(os);
  for (int d = 0; d < r; ++d) { pp.oshape[d] = n.i[1 + 3 * r + d]; pp.istride[d] = (int64_t)n.i[1 + 2 * r + d] * xstr[d]; }
  BufRef xr = x.buf, yr = out.buf;
  const int es = (int)dtype_size(x.dtype);
  const int64_t off = first * es;
  Step s;
  s.name = n.name; s.type = "StridedSlice"; s.impl = "permute"; s.fused = n.name;
  s.bytes = 2.0 * (double)out.bytes();
  s.launch = [=](const RunCtx& c) { PermuteParams q = pp; q.x = c.ptr(xr) + off; q.y = c.ptr(yr); launch_permute(q, es, (cudaStream_t)c.stream); };
  emit(std::move(s), {&x}, {&out});
  set(n.outputs[0], out);
}

void Lowerer::lower_transpose(const Node& n) {
  const Shape& xs = val(n.inputs[0]).shape;
  const Shape os = val(n.outputs[0]).shape;
  const int r = (int)xs.size();
  bool identity = true;
  for (int d = 0; d < r; ++d) identity &= n.i[d] == d;
  auto it = vt_.find(n.inputs[0]);
  // NCHW -> NHWC of a tensor that already lives channel-last: a view, or a pad-dropping copy
  if (r == 4 && n.i[0] == 0 && n.i[1] == 2 && n.i[2] == 3 && n.i[3] == 1 && it != vt_.end() && it->second.layout == L_NHWC8) {
    LTensor x = it->second;
    const int64_t C = xs[1], Cp = round_up(C, 8);
    if (C == Cp) {
      LTensor v = x;
      v.shape = os; v.layout = L_ROWMAJOR;
      set(n.outputs[0], v);
      return;
    }
    LTensor out = alloc(os, x.dtype, L_ROWMAJOR);
    const int64_t es = dtype_size(x.dtype), rows = xs[0] * xs[2] * xs[3];
    BufRef xr = x.buf, yr = out.buf;
    Step s;
    s.name = n.name; s.type = "Transpose"; s.impl = "nhwc8_compact"; s.fused = n.name;
    s.bytes = (double)x.bytes() + (double)out.bytes();
    s.launch = [=](const RunCtx& c) { launch_copy2d(c.ptr(xr), Cp * es, c.ptr(yr), C * es, C * es, rows, (cudaStream_t)c.stream); };
    emit(std::move(s), {&x}, {&out});
    set(n.outputs[0], out);
    return;
  }
  LTensor x = get(n.inputs[0], L_ROWMAJOR);
  if (identity) { LTensor v = x; v.shape = os; set(n.outputs[0], v); return; }
  LTensor out = alloc(os, x.dtype, L_ROWMAJOR);
  PermuteParams pp{};
  pp.rank = r;
  pp.n = numel(os);
  std::vector<int64_t> istr(r);
  int64_t st = 1;
  for (int d = r - 1; d >= 0; --d) { istr[d] = st; st *= xs[d]; }
  for (int d = 0; d < r; ++d) { pp.oshape[d] = os[d]; pp.istride[d] = istr[n.i[d]]; }
  BufRef xr = x.buf, yr = out.buf;
  int es = (int)dtype_size(x.dtype);
  Step s;
  s.name = n.name; s.type = "Transpose"; s.impl = "permute"; s.fused = n.name;
  s.bytes = 2.0 * (double)x.bytes();
  s.launch = [=](const RunCtx& c) { PermuteParams q = pp; q.x = c.ptr(xr); q.y = c.ptr(yr); launch_permute(q, es, (cudaStream_t)c.stream); };
  emit(std::move(s), {&x}, {&out});
  set(n.outputs[0], out);
}

void Lowerer::lower_detection(const Node& n) {
  if (n.op == B200_OP_TOPK) {
    LTensor x = to_dtype(get(n.inputs[0], L_ROWMAJOR), B200_F32, n.name);
    const Shape& xs = x.shape;
    int ax = n.i[1] < 0 ? n.i[1] + (int)xs.size() : n.i[1];
    int64_t outer = 1, inner = 1;
    for (int d = 0; d < ax; ++d) outer *= xs[d];
    for (int d = ax + 1; d < (int)xs.size(); ++d) inner *= xs[d];
    const Shape os = val(n.outputs[0]).shape;
    LTensor ov = alloc(os, B200_F32, L_ROWMAJOR), oi = alloc(os, B200_I32, L_ROWMAJOR);
    BufRef xr = x.buf, vr = ov.buf, ir = oi.buf;
    int64_t axis = xs[ax];
    int k = n.i[0], mode = n.i[2], sort = n.i[3];
    Step s;
    s.name = n.name; s.type = "TopK"; s.impl = "block_select"; s.fused = n.name;
    s.bytes = (double)x.bytes();
    s.launch = [=](const RunCtx& c) {
      launch_topk((const float*)c.ptr(xr), (float*)c.ptr(vr), (int32_t*)c.ptr(ir), outer, axis, inner, k, mode, sort, (cudaStream_t)c.stream);
    };
    emit(std::move(s), {&x}, {&ov, &oi});
    set(n.outputs[0], ov);
    set(n.outputs[1], oi);
    return;
  }
  if (n.op == B200_OP_NMS) {
    LTensor boxes = to_dtype(get(n.inputs[0], L_ROWMAJOR), B200_F32, n.name + "_boxes");
    LTensor scores = to_dtype(get(n.inputs[1], L_ROWMAJOR), B200_F32, n.name + "_scores");
    NmsParams p{};
    p.B = (int)scores.shape[0]; p.C = (int)scores.shape[1]; p.Nb = (int)scores.shape[2];
    p.max_out = (int)std::min<int64_t>(n.i[0], p.Nb); p.center_format = n.i[1]; p.sort_desc = n.i[2];
    p.iou_thr = n.f[0]; p.score_thr = n.f[1]; p.soft_sigma = n.f[2];
    LTensor o0 = alloc(val(n.outputs[0]).shape, B200_I32, L_ROWMAJOR), o1 = alloc(val(n.outputs[1]).shape, B200_F32, L_ROWMAJOR),
            o2 = alloc(val(n.outputs[2]).shape, B200_I32, L_ROWMAJOR);
    nms_check(p);     // unsupported geometry fails LoadNetwork / QueryNetwork, not the first inference
    int64_t wsb = nms_workspace_bytes(p.B, p.C, p.Nb, p.max_out, p.soft_sigma > 0.f);
    Shape wshape = {wsb};
    LTensor ws = alloc(wshape, B200_U8, L_ROWMAJOR);
    p.workspace_bytes = wsb;
    BufRef br = boxes.buf, sr = scores.buf, r0 = o0.buf, r1 = o1.buf, r2 = o2.buf, wr = ws.buf;
    Step s;
    s.name = n.name; s.type = "NonMaxSuppression"; s.impl = p.soft_sigma > 0.f ? "nms5_soft_queue" : "nms5_sort_greedy"; s.fused = n.name;
    s.bytes = (double)boxes.bytes() + (double)scores.bytes();
    s.launch = [=](const RunCtx& c) {
      NmsParams q = p;
      q.boxes = (const float*)c.ptr(br); q.scores = (const float*)c.ptr(sr); q.sel_idx = (int32_t*)c.ptr(r0);
      q.sel_scores = (float*)c.ptr(r1); q.valid = (int32_t*)c.ptr(r2); q.workspace = c.ptr(wr);
      launch_nms(q, (cudaStream_t)c.stream);
    };
    emit(std::move(s), {&boxes, &scores}, {&o0, &o1, &o2, &ws});
    set(n.outputs[0], o0); set(n.outputs[1], o1); set(n.outputs[2], o2);
    return;
  }
  // DetectionOutput
  LTensor loc = to_dtype(get(n.inputs[0], L_ROWMAJOR), B200_F32, n.name + "_loc");
  LTensor conf = to_dtype(get(n.inputs[1], L_ROWMAJOR), B200_F32, n.name + "_conf");
  LTensor pri = to_dtype(get(n.inputs[2], L_ROWMAJOR), B200_F32, n.name + "_priors");
  DetOutParams p{};
  p.N = (int)loc.shape[0];
  p.num_classes = n.i[0]; p.background = n.i[1]; p.top_k = n.i[2]; p.keep_top_k = n.i[3]; p.code_center = n.i[4];
  p.share_location = n.i[5]; p.var_in_target = n.i[6]; p.normalized = n.i[7]; p.clip_before = n.i[8]; p.clip_after = n.i[9];
  p.input_h = n.i[10]; p.input_w = n.i[11]; p.decrease_label_id = n.i[12];
  p.nms_thr = n.f[0]; p.conf_thr = n.f[1];
  const int prior_size = p.normalized ? 4 : 5;
  p.num_priors = (int)(pri.shape[2] / prior_size);
  p.num_loc_classes = p.share_location ? 1 : p.num_classes;
  p.priors_batch = (int)pri.shape[0];
  B200_CHECK(loc.shape[1] == (int64_t)p.num_priors * p.num_loc_classes * 4, B200_ERR_INVALID, "DetectionOutput '" << n.name << "': loc size mismatch");
  B200_CHECK(conf.shape[1] == (int64_t)p.num_priors * p.num_classes, B200_ERR_INVALID, "DetectionOutput '" << n.name << "': conf size mismatch");
  const Shape os = val(n.outputs[0]).shape;
  p.out_rows = os[2];
  detection_output_check(p);     // e.g. SSD512's 24564 priors with top_k = -1: refused here, at compile time
  LTensor out = alloc(os, B200_F32, L_ROWMAJOR);
  int64_t wsb = detout_workspace_bytes(p);
  Shape wshape = {wsb};
  LTensor ws = alloc(wshape, B200_U8, L_ROWMAJOR);
  p.workspace_bytes = wsb;
  BufRef lr = loc.buf, cr = conf.buf, pr = pri.buf, orf = out.buf, wr = ws.buf;
  Step s;
  s.name = n.name; s.type = "DetectionOutput"; s.impl = "decode_sort_nms_topk"; s.fused = n.name;
  s.bytes = (double)loc.bytes() + (double)conf.bytes();
  s.launch = [=](const RunCtx& c) {
    DetOutParams q = p;
    q.loc = (const float*)c.ptr(lr); q.conf = (const float*)c.ptr(cr); q.priors = (const float*)c.ptr(pr);
    q.out = (float*)c.ptr(orf); q.workspace = c.ptr(wr);
    launch_detection_output(q, (cudaStream_t)c.stream);
  };
  emit(std::move(s), {&loc, &conf, &pri}, {&out, &ws});
  set(n.outputs[0], out);
}

void Lowerer::lower_node(const Node& n) {
  // fused records are emitted at the position of their last absorbed node
  auto fat = fattn_.find(n.id);
  if (fat != fattn_.end()) { lower_attention(fat->second); return; }
  auto fc = fconv_.find(n.id);
  if (fc != fconv_.end()) { if (!mb_done_.count(n.id)) lower_conv(fc->second); return; }
  auto fm = fmvn_.find(n.id);
  if (fm != fmvn_.end()) { lower_mvn(node(fm->second.root), &fm->second); return; }
  auto fe = felt_.find(n.id);
  if (fe != felt_.end()) { lower_binary(node(fe->second.root), fe->second.act, fe->second.out_value); return; }
  if (absorbed_by_[n.id] >= 0) return;  // folded into another kernel, emitted elsewhere

  switch (n.op) {
    case B200_OP_PARAMETER: {
      const Value& v = val(n.outputs[0]);
      LTensor t;
      if (stem_s2d_applies(v.id)) {
        // the only consumer is a 7x7/8x8 stride-2 convolution over <= 4 channels: ingest straight into the space-to-depth layout
        const Node& cv = node(v.consumers[0]);
        const Shape os = val(cv.outputs[0]).shape;
        t.shape = v.shape; t.dtype = cd_; t.layout = L_S2D16;
        t.s2d_pt = cv.i[4]; t.s2d_pl = cv.i[5]; t.s2d_hb = (int)os[2] + 3; t.s2d_wb = (int)os[3] + 3;
        t.buf.space = SP_ARENA;
        t.buf.box = new_box(t.bytes(), /*pinned=*/true);
      } else {
        t = alloc(v.shape, storage_dtype(v.dtype), prefers_nhwc8(v.id) ? L_NHWC8 : L_ROWMAJOR, /*pinned=*/true);
      }
      set(v.id, t);
      IOBinding b;
      b.name = n.name; b.shape = v.shape; b.graph_dtype = v.dtype; b.value_id = v.id;
      net_.inputs.push_back(b);
      net_.in_tensors.push_back(t);
      break;
    }
    case B200_OP_CONSTANT:
      break;  // materialised lazily by get()
    case B200_OP_RESULT: {
      LTensor t = get(n.inputs[0], -1);
      IOBinding b;
      // output name = producer friendly name (+ .port) - mkldnn_graph.cpp:230-241
      b.name = val(n.inputs[0]).name; b.shape = t.shape; b.graph_dtype = val(n.inputs[0]).dtype; b.value_id = n.inputs[0];
      net_.outputs.push_back(b);
      net_.out_tensors.push_back(t);
      break;
    }
    case B200_OP_MAX_POOL: case B200_OP_AVG_POOL: lower_pool(n); break;
    case B200_OP_FAKE_QUANTIZE: lower_fake_quantize(n); break;
    case B200_OP_REDUCE_MEAN: case B200_OP_REDUCE_SUM: case B200_OP_REDUCE_MAX: lower_reduce(n); break;
    case B200_OP_SOFTMAX: {
      auto fs = fsm_.find(n.id);
      if (fs != fsm_.end()) {
        const FusedSoftmax& f = fs->second;
        LTensor x = get(f.x_value, L_ROWMAJOR);
        LTensor mk;
        if (f.mask_value >= 0) mk = get(f.mask_value, L_ROWMAJOR);
        const Shape& xs = x.shape;
        const int64_t cols = xs.back(), rows = numel(xs) / cols;
        // mask [N,1,..,1,S]: one mask row per leading index; [1,..,1,S]: one row for all
        const int64_t rows_per_mask = (f.mask_value >= 0 && mk.shape[0] == xs[0] && xs[0] > 1) ? rows / xs[0] : rows;
        LTensor out = alloc(xs, x.dtype, L_ROWMAJOR);
        BufRef xr = x.buf, mr = mk.buf, yr = out.buf;
        const bool has_mask = f.mask_value >= 0;
        const float scale = f.scale;
        const int dt = x.dtype;
        B200_CHECK(!has_mask || mk.dtype == x.dtype, B200_ERR_UNSUPPORTED, "softmax mask dtype differs from the scores'");
        Step s;
        s.name = n.name; s.type = "Softmax"; s.impl = has_mask ? "softmax_scale_mask_warp_rows" : "softmax_scale_warp_rows";
        s.fused = f.fused + "," + n.name;
        s.bytes = 2.0 * (double)x.bytes();
        s.launch = [=](const RunCtx& c) {
          launch_softmax_scale_mask(c.ptr(xr), has_mask ? c.ptr(mr) : nullptr, c.ptr(yr), rows, cols, scale, rows_per_mask, dt, (cudaStream_t)c.stream);
        };
        emit(std::move(s), {&x, has_mask ? &mk : nullptr}, {&out});
        set(n.outputs[0], out);
        break;
      }
      LTensor x = get(n.inputs[0], L_ROWMAJOR);
      const Shape& xs = x.shape;
      int64_t outer = 1, inner = 1;
      for (int d = 0; d < n.i[0]; ++d) outer *= xs[d];
      for (int d = n.i[0] + 1; d < (int)xs.size(); ++d) inner *= xs[d];
      LTensor out = alloc(xs, x.dtype, L_ROWMAJOR);
      BufRef xr = x.buf, yr = out.buf;
      int64_t axis = xs[n.i[0]];
      int dt = x.dtype;
      Step s;
      s.name = n.name; s.type = "Softmax"; s.fused = n.name;
      s.impl = inner == 1 ? ((softmax_fused_supported(axis, dt) && axis >= 64) ? "softmax_warp_rows" : (axis <= 128 ? "softmax_short_rows" : "softmax_rows")) : "softmax_strided";
      s.bytes = 2.0 * (double)x.bytes();
      s.launch = [=](const RunCtx& c) { launch_softmax(c.ptr(xr), c.ptr(yr), outer, axis, inner, dt, (cudaStream_t)c.stream); };
      emit(std::move(s), {&x}, {&out});
      set(n.outputs[0], out);
      break;
    }
    case B200_OP_RESHAPE: {
      LTensor x = get(n.inputs[0], L_ROWMAJOR);
      LTensor v = x;
      v.shape = val(n.outputs[0]).shape;
      set(n.outputs[0], v);
      break;
    }
    case B200_OP_TRANSPOSE: lower_transpose(n); break;
    case B200_OP_STRIDED_SLICE: lower_strided_slice(n); break;
    case B200_OP_CONCAT: lower_concat(n); break;
    case B200_OP_SPLIT: lower_split(n); break;
    case B200_OP_GATHER: {
      int ax = n.i[0] < 0 ? n.i[0] + (int)val(n.inputs[0]).shape.size() : n.i[0];
      B200_CHECK(ax == 0, B200_ERR_UNSUPPORTED, "Gather '" << n.name << "': only axis 0 is supported");
      LTensor table = get(n.inputs[0], L_ROWMAJOR), idx = get(n.inputs[1], L_ROWMAJOR);
      B200_CHECK(idx.dtype == B200_I32, B200_ERR_UNSUPPORTED, "Gather '" << n.name << "': indices must be i32/i64");
      const Shape os = val(n.outputs[0]).shape;
      LTensor out = alloc(os, table.dtype, L_ROWMAJOR);
      int64_t rows = table.shape[0], row_elems = numel(table.shape) / std::max<int64_t>(rows, 1), n_idx = numel(idx.shape);
      BufRef tr = table.buf, ir = idx.buf, yr = out.buf;
      int dt = table.dtype;
      Step s;
      s.name = n.name; s.type = "Gather"; s.impl = "gather_rows"; s.fused = n.name;
      s.bytes = 2.0 * (double)out.bytes();
      s.launch = [=](const RunCtx& c) { launch_gather_rows(c.ptr(tr), (const int32_t*)c.ptr(ir), c.ptr(yr), n_idx, row_elems, rows, dt, (cudaStream_t)c.stream); };
      emit(std::move(s), {&table, &idx}, {&out});
      set(n.outputs[0], out);
      break;
    }
    case B200_OP_TOPK: case B200_OP_NMS: case B200_OP_DETECTION_OUTPUT: lower_detection(n); break;
    case B200_OP_CONVERT: {
      LTensor x = get(n.inputs[0], -1);
      int target = storage_dtype(n.dtype);
      if (x.dtype == target) { set(n.outputs[0], x); break; }
      LTensor xr = get(n.inputs[0], L_ROWMAJOR);
      set(n.outputs[0], to_dtype(xr, target, n.name));
      break;
    }
    case B200_OP_CONVOLUTION: case B200_OP_GROUP_CONVOLUTION:
      B200_THROW(B200_ERR_UNSUPPORTED, "Convolution '" << n.name << "' with non-constant weights is not supported");
    default: {
      ActSpec as;
      if (act_from_op(n, as)) { lower_unary(n, as); break; }
      B200_THROW(B200_ERR_UNSUPPORTED, "no kernel for op " << op_name(n.op) << " ('" << n.name << "')");
    }
  }
}

void Lowerer::run() {
  net_.name = g_.name;
  B200_CHECK(!g_.results.empty(), B200_ERR_INVALID, "graph has no Result");
  fuse();
  for (const Node& n : g_.nodes) lower_node(n);
  // live ranges: parameters live from before step 0, results until after the last step
  const int nsteps = (int)net_.steps.size();
  for (int si = 0; si < nsteps; ++si) {
    for (int b : net_.steps[si].reads) { auto& bx = net_.boxes[b]; bx.first = std::min(bx.first, si); bx.last = std::max(bx.last, si); }
    for (int b : net_.steps[si].writes) { auto& bx = net_.boxes[b]; bx.first = std::min(bx.first, si); bx.last = std::max(bx.last, si); }
  }
  for (auto& t : net_.in_tensors) if (t.buf.space == SP_ARENA) { net_.boxes[t.buf.box].first = -1; net_.boxes[t.buf.box].last = std::max(net_.boxes[t.buf.box].last, 0); }
  for (auto& t : net_.out_tensors) if (t.buf.space == SP_ARENA) { net_.boxes[t.buf.box].last = nsteps; net_.boxes[t.buf.box].first = std::min(net_.boxes[t.buf.box].first, nsteps); }
  plan_arena(net_.boxes, net_.arena_bytes);
  net_.arena_bytes = round_up(net_.arena_bytes + 1024, 1 << 20);
  net_.box_offsets.resize(net_.boxes.size());
  for (size_t b = 0; b < net_.boxes.size(); ++b) net_.box_offsets[b] = net_.boxes[b].offset;
}

}  // namespace

// Greedy best-fit over live ranges, largest boxes first: the formulation of MemorySolver
// (inference-engine/src/mkldnn_plugin/mkldnn_memory_solver.cpp:51-96) with 1 KiB alignment.
void plan_arena(std::vector<ArenaBox>& boxes, int64_t& total) {
  std::vector<int> order(boxes.size());
  for (size_t i = 0; i < boxes.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return boxes[a].bytes > boxes[b].bytes; });
  std::vector<int> placed;
  total = 0;
  for (int id : order) {
    ArenaBox& b = boxes[id];
    if (b.last < b.first) { b.offset = 0; continue; }  // never used
    std::vector<std::pair<int64_t, int64_t>> busy;
    for (int pid : placed) {
      const ArenaBox& q = boxes[pid];
      bool overlap_time = b.pinned || q.pinned || !(b.last < q.first || q.last < b.first);
      if (overlap_time) busy.emplace_back(q.offset, q.offset + q.bytes);
    }
    std::sort(busy.begin(), busy.end());
    int64_t off = 0;
    for (auto& iv : busy) {
      if (off + b.bytes <= iv.first) break;
      off = std::max(off, iv.second);
    }
    b.offset = off;
    total = std::max(total, off + b.bytes);
    placed.push_back(id);
  }
}

void lower_graph(const b200_graph& g, const b200_compile_options& opts, b200_network& net) {
  Lowerer l(g, opts, net);
  l.run();
}

}  // namespace b200

// The planner on its own (unit-testable without a device): boxes are (first, last, bytes) over execution
// indices, last < 0 = "lives to the end" (MemorySolver::Box, mkldnn_memory_solver.hpp:52-69).
extern "C" int b200_plan_arena(int32_t n_boxes, const int64_t* bytes, const int32_t* first, const int32_t* last, int64_t* offsets,
                               int64_t* total_bytes) {
  if (n_boxes < 0 || (n_boxes > 0 && (!bytes || !first || !last || !offsets)) || !total_bytes) {
    b200::set_last_error("b200_plan_arena: null argument");
    return B200_ERR_INVALID;
  }
  std::vector<b200::ArenaBox> boxes((size_t)n_boxes);
  int max_idx = 0;
  for (int i = 0; i < n_boxes; ++i) max_idx = std::max(max_idx, std::max(first[i], last[i]));
  for (int i = 0; i < n_boxes; ++i) {
    if (bytes[i] < 0 || first[i] < 0) {
      b200::set_last_error("b200_plan_arena: negative size or start");
      return B200_ERR_INVALID;
    }
    boxes[i].bytes = bytes[i];
    boxes[i].first = first[i];
    boxes[i].last = last[i] < 0 ? max_idx + 1 : last[i];
  }
  int64_t total = 0;
  b200::plan_arena(boxes, total);
  for (int i = 0; i < n_boxes; ++i) offsets[i] = boxes[i].offset;
  *total_bytes = total;
  return B200_OK;
}
