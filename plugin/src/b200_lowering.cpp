// ngraph::Function -> b200_graph.  See b200_lowering.hpp.
#include "b200_lowering.hpp"

#include <ie_common.h>

#include <map>
#include <ngraph/graph_util.hpp>
#include <ngraph/opsets/opset1.hpp>
#include <ngraph/opsets/opset3.hpp>
#include <ngraph/opsets/opset4.hpp>
#include <ngraph/opsets/opset5.hpp>
#include <ngraph/opsets/opset6.hpp>
#include <ngraph/opsets/opset7.hpp>
#include <ngraph/pass/constant_folding.hpp>
#include <ngraph/pass/manager.hpp>
#include <ngraph/pass/pass.hpp>
#include <ngraph/rt_info.hpp>
#include <ngraph/slice_plan.hpp>
#include <ngraph_ops/nms_ie_internal.hpp>
#include <transformations/common_optimizations/common_optimizations.hpp>
#include <transformations/op_conversions/convert_nms_to_nms_ie_internal.hpp>
#include <transformations/op_conversions/convert_previous_nms_to_nms_5.hpp>
#include <transformations/op_conversions/convert_divide.hpp>
#include <transformations/op_conversions/convert_gelu.hpp>
#include <transformations/op_conversions/convert_minimum_to_power_and_max.hpp>
#include <transformations/op_conversions/convert_negative.hpp>
#include <transformations/op_conversions/convert_reduce_to_pooling.hpp>
#include <transformations/op_conversions/convert_subtract.hpp>
#include <transformations/op_conversions/gelu7_downgrade.hpp>
#include <transformations/op_conversions/hsigmoid_decomposition.hpp>
#include <transformations/op_conversions/hswish_decomposition.hpp>
#include <transformations/op_conversions/mvn6_decomposition.hpp>
#include <utility>
#include <vector>

namespace B200Plugin {

using ngraph::as_type_ptr;
using NodePtr = std::shared_ptr<ngraph::Node>;

namespace {

// Convolutions over 1 or 3 spatial dimensions expressed through the 2-D kernels (the tcgen05 implicit GEMM / the fp32 kernel), with
// nothing but data movement around them:
//   [N,C,W]      -> Reshape to [N,C,1,W], 2-D convolution with a 1 x kw filter, Reshape back;
//   [N,C,D,H,W]  -> the depth taps become input channels.  The input is zero-padded along D, a Gather of whole depth slices with the constant index list
//                   {d' sd + t dd : tap t, output depth d'} stacks the kd depth-shifted copies, Reshape / Transpose make it
//                   [N D', kd C, H, W], ONE 2-D convolution with the filter re-packed to [O, kd C, kh, kw] follows, and a Reshape /
//                   Transpose pair returns [N, O, D', H', W'].  Sum order differs from the reference's NCDHW loop only inside the fp32
//                   accumulation of one output element (ngraph reference/convolution.hpp:89-146 semantics: cross-correlation, zero padding).
// Found by the reference's conformance binary (single_layer_tests/convolution.cpp of the template instances: 80 NCDHW cases).
class ConvolutionToTwoD : public ngraph::pass::FunctionPass {
public:
    bool run_on_function(std::shared_ptr<ngraph::Function> f) override {
        bool changed = false;
        for (const auto& node : f->get_ordered_ops()) {
            auto conv = as_type_ptr<ngraph::opset1::Convolution>(node);
            if (!conv || !conv->get_input_partial_shape(0).is_static() || !conv->get_output_partial_shape(0).is_static()) continue;
            auto weights = as_type_ptr<ngraph::opset1::Constant>(conv->input_value(1).get_node_shared_ptr());
            const ngraph::Shape xs = conv->get_input_shape(0), os = conv->get_output_shape(0);
            if (!weights || (xs.size() != 3 && xs.size() != 5)) continue;
            const ngraph::Shape ws = weights->get_shape();
            const auto& strides = conv->get_strides();
            const auto& dil = conv->get_dilations();
            const auto& pb = conv->get_pads_begin();     // resolved by validate_and_infer_types for SAME_* / VALID
            const auto& pe = conv->get_pads_end();
            ngraph::NodeVector made;
            auto keep = [&](const NodePtr& n) { made.push_back(n); return n; };
            auto i64 = [&](const std::vector<int64_t>& v) { return keep(ngraph::opset1::Constant::create(ngraph::element::i64, ngraph::Shape{v.size()}, v)); };
            auto dims = [](const ngraph::Shape& sh) { return std::vector<int64_t>(sh.begin(), sh.end()); };
            ngraph::Output<ngraph::Node> out;
            if (xs.size() == 3) {
                auto x4 = keep(std::make_shared<ngraph::opset1::Reshape>(conv->input_value(0), i64({(int64_t)xs[0], (int64_t)xs[1], 1, (int64_t)xs[2]}), false));
                auto w4 = keep(std::make_shared<ngraph::opset1::Constant>(weights->get_element_type(), ngraph::Shape{ws[0], ws[1], 1, ws[2]}, weights->get_data_ptr()));
                auto c2 = keep(std::make_shared<ngraph::opset1::Convolution>(x4, w4, ngraph::Strides{1, strides[0]}, ngraph::CoordinateDiff{0, pb[0]},
                                                                            ngraph::CoordinateDiff{0, pe[0]}, ngraph::Strides{1, dil[0]}));
                out = keep(std::make_shared<ngraph::opset1::Reshape>(c2, i64(dims(os)), false));
            } else {
                if (pb[0] < 0 || pe[0] < 0) continue;
                const int64_t N = xs[0], C = xs[1], D = xs[2], H = xs[3], W = xs[4], O = ws[0], kd = ws[2], kh = ws[3], kw = ws[4];
                const int64_t Dp = D + pb[0] + pe[0], Do = os[2];
                const auto et = conv->get_input_element_type(0);
                ngraph::Output<ngraph::Node> x = conv->input_value(0);
                if (Dp != D) {
                    ngraph::OutputVector parts;
                    auto zeros = [&](int64_t d) {
                        return keep(ngraph::opset1::Constant::create(et, ngraph::Shape{(size_t)N, (size_t)C, (size_t)d, (size_t)H, (size_t)W}, std::vector<float>{0.f}));
                    };
                    if (pb[0] > 0) parts.push_back(zeros(pb[0]));
                    parts.push_back(x);
                    if (pe[0] > 0) parts.push_back(zeros(pe[0]));
                    x = keep(std::make_shared<ngraph::opset1::Concat>(parts, 2));
                }
                std::vector<int64_t> index;
                for (int64_t t = 0; t < kd; ++t)
                    for (int64_t d = 0; d < Do; ++d) index.push_back(d * (int64_t)strides[0] + t * (int64_t)dil[0]);
                // depth first (the gather kernel takes whole slices along axis 0), gather, then [kd, D', N, C, H, W] -> [N, D', kd, C, H, W]
                auto dfirst = keep(std::make_shared<ngraph::opset1::Transpose>(x, i64({2, 0, 1, 3, 4})));
                auto stacked = keep(std::make_shared<ngraph::opset1::Gather>(dfirst, i64(index), keep(ngraph::opset1::Constant::create(ngraph::element::i64, ngraph::Shape{}, {0}))));
                auto s6 = keep(std::make_shared<ngraph::opset1::Reshape>(stacked, i64({kd, Do, N, C, H, W}), false));
                auto t6 = keep(std::make_shared<ngraph::opset1::Transpose>(s6, i64({2, 1, 0, 3, 4, 5})));
                auto x4 = keep(std::make_shared<ngraph::opset1::Reshape>(t6, i64({N * Do, kd * C, H, W}), false));
                // filter [O, C, kd, kh, kw] -> [O, kd, C, kh, kw] -> [O, kd C, kh, kw], on the host
                const std::vector<float> wsrc = weights->cast_vector<float>();
                std::vector<float> w2(wsrc.size());
                for (int64_t o = 0; o < O; ++o)
                    for (int64_t c = 0; c < C; ++c)
                        for (int64_t t = 0; t < kd; ++t)
                            std::copy_n(&wsrc[(size_t)(((o * C + c) * kd + t) * kh * kw)], (size_t)(kh * kw), &w2[(size_t)(((o * kd + t) * C + c) * kh * kw)]);
                auto w4 = keep(ngraph::opset1::Constant::create(weights->get_element_type(), ngraph::Shape{(size_t)O, (size_t)(kd * C), (size_t)kh, (size_t)kw}, w2));
                auto c2 = keep(std::make_shared<ngraph::opset1::Convolution>(x4, w4, ngraph::Strides{strides[1], strides[2]}, ngraph::CoordinateDiff{pb[1], pb[2]},
                                                                            ngraph::CoordinateDiff{pe[1], pe[2]}, ngraph::Strides{dil[1], dil[2]}));
                auto r5 = keep(std::make_shared<ngraph::opset1::Reshape>(c2, i64({N, Do, O, (int64_t)os[3], (int64_t)os[4]}), false));
                out = keep(std::make_shared<ngraph::opset1::Transpose>(r5, i64({0, 2, 1, 3, 4})));
            }
            out.get_node_shared_ptr()->set_friendly_name(conv->get_friendly_name());
            ngraph::copy_runtime_info(conv, made);
            ngraph::replace_node(conv, out.get_node_shared_ptr());
            changed = true;
        }
        return changed;
    }
};

// GRN (global response normalisation across channels, opset1): y = x / sqrt(sum_c x^2 + bias), spelled with the element-wise and
// reduction kernels the library has (ngraph reference/grn.hpp; the reference CPU plugin runs nodes/mkldnn_grn_node.cpp).
class GrnDecomposition : public ngraph::pass::FunctionPass {
public:
    bool run_on_function(std::shared_ptr<ngraph::Function> f) override {
        bool changed = false;
        for (const auto& node : f->get_ordered_ops()) {
            auto grn = as_type_ptr<ngraph::opset1::GRN>(node);
            if (!grn || !grn->get_input_partial_shape(0).is_static() || grn->get_input_shape(0).size() < 2) continue;
            const auto et = grn->get_input_element_type(0);
            auto x = grn->input_value(0);
            auto sq = std::make_shared<ngraph::opset1::Multiply>(x, x);
            auto sum = std::make_shared<ngraph::opset1::ReduceSum>(sq, ngraph::opset1::Constant::create(ngraph::element::i64, ngraph::Shape{1}, {1}), true);
            auto biased = std::make_shared<ngraph::opset1::Add>(sum, ngraph::opset1::Constant::create(et, ngraph::Shape{}, {grn->get_bias()}));
            auto norm = std::make_shared<ngraph::opset1::Sqrt>(biased);
            auto out = std::make_shared<ngraph::opset1::Divide>(x, norm);
            out->set_friendly_name(grn->get_friendly_name());
            ngraph::copy_runtime_info(grn, {sq, sum, biased, norm, out});
            ngraph::replace_node(grn, out);
            changed = true;
        }
        return changed;
    }
};

// VariadicSplit (axis and lengths constant, one length may be -1) as one StridedSlice per output (mkldnn_split_node.cpp covers both
// Split and VariadicSplit; here the equal split keeps its own op and the variadic one becomes slices).  Slice k is named "<name>.<k>":
// the output-naming rule (producer name + "." + port) then gives network outputs the names they had.
class VariadicSplitToSlices : public ngraph::pass::FunctionPass {
public:
    bool run_on_function(std::shared_ptr<ngraph::Function> f) override {
        bool changed = false;
        for (const auto& node : f->get_ordered_ops()) {
            auto vs = as_type_ptr<ngraph::opset1::VariadicSplit>(node);
            if (!vs || !vs->get_input_partial_shape(0).is_static()) continue;
            auto axis_c = as_type_ptr<ngraph::opset1::Constant>(vs->input_value(1).get_node_shared_ptr());
            auto len_c = as_type_ptr<ngraph::opset1::Constant>(vs->input_value(2).get_node_shared_ptr());
            if (!axis_c || !len_c) continue;
            const ngraph::Shape xs = vs->get_input_shape(0);
            int64_t axis = axis_c->cast_vector<int64_t>()[0];
            if (axis < 0) axis += static_cast<int64_t>(xs.size());
            if (axis < 0 || axis >= static_cast<int64_t>(xs.size())) continue;
            ngraph::OutputVector outs;
            ngraph::NodeVector made;
            int64_t start = 0;
            for (size_t k = 0; k < vs->get_output_size(); ++k) {
                const int64_t len = static_cast<int64_t>(vs->get_output_shape(k)[static_cast<size_t>(axis)]);
                std::vector<int64_t> b(xs.size(), 0), e(xs.size(), 0), mask(xs.size(), 1);
                b[axis] = start; e[axis] = start + len; mask[axis] = 0;
                auto sl = std::make_shared<ngraph::opset1::StridedSlice>(vs->input_value(0), ngraph::opset1::Constant::create(ngraph::element::i64, {xs.size()}, b),
                                                                        ngraph::opset1::Constant::create(ngraph::element::i64, {xs.size()}, e), mask, mask);
                sl->set_friendly_name(vs->get_friendly_name() + "." + std::to_string(k));
                made.push_back(sl);
                outs.push_back(sl->output(0));
                start += len;
            }
            ngraph::copy_runtime_info(vs, made);
            ngraph::replace_node(vs, outs);
            changed = true;
        }
        return changed;
    }
};

}  // namespace

std::shared_ptr<ngraph::Function> TransformNetwork(const std::shared_ptr<const ngraph::Function>& function) {
    auto f = ngraph::clone_function(*function);
    ngraph::pass::Manager manager;
    manager.register_pass<ConvolutionToTwoD>();
    manager.register_pass<GrnDecomposition>();
    manager.register_pass<ngraph::pass::CommonOptimizations>();
    // after CommonOptimizations: its StridedSliceOptimization MERGES neighbouring slices of one tensor into a VariadicSplit
    manager.register_pass<VariadicSplitToSlices>();
    // NMS-1/3/4 -> NMS-5 -> NonMaxSuppressionIEInternal (static, padded output shapes), as mkldnn_plugin.cpp:150-156 does
    manager.register_pass<ngraph::pass::ConvertNMS1ToNMS5>();
    manager.register_pass<ngraph::pass::ConvertNMS3ToNMS5>();
    manager.register_pass<ngraph::pass::ConvertNMS4ToNMS5>();
    manager.register_pass<ngraph::pass::ConvertNMSToNMSIEInternal>();
    manager.register_pass<ngraph::pass::ConstantFolding>();
    // ops the kernel library executes natively must not be decomposed (same stance as mkldnn_plugin.cpp:281-300)
    auto cfg = manager.get_pass_config();
    cfg->disable<ngraph::pass::ConvertGELU>();
    cfg->disable<ngraph::pass::Gelu7Downgrade>();
    cfg->disable<ngraph::pass::HSwishDecomposition>();
    cfg->disable<ngraph::pass::HSigmoidDecomposition>();
    cfg->disable<ngraph::pass::MVN6Decomposition>();
    cfg->disable<ngraph::pass::ConvertSubtract>();
    cfg->disable<ngraph::pass::ConvertDivide>();
    cfg->disable<ngraph::pass::ConvertNegative>();
    cfg->disable<ngraph::pass::ConvertMinimum>();
    cfg->disable<ngraph::pass::ConvertReduceMeanToPooling>();
    cfg->disable<ngraph::pass::ConvertReduceMaxToPooling>();
    cfg->disable<ngraph::pass::ConvertReduceSumToPooling>();
    manager.run_passes(f);
    return f;
}

void AddMeanScale(const std::shared_ptr<ngraph::Function>& function, const InferenceEngine::InputsDataMap& inputs) {
    for (const auto& param : function->get_parameters()) {
        auto it = inputs.find(param->get_friendly_name());
        if (it == inputs.end()) continue;
        const InferenceEngine::PreProcessInfo& info = it->second->getPreProcess();
        const size_t channels = info.getNumberOfChannels();
        if (channels == 0) continue;
        std::vector<float> mean(channels), scale(channels);
        bool has_mean = false, has_scale = false, has_image = false;
        for (size_t c = 0; c < channels; ++c) {
            has_image = has_image || info[c]->meanData != nullptr;
            mean[c] = info[c]->meanValue;
            scale[c] = info[c]->stdScale;
            has_mean = has_mean || mean[c] != 0.f;
            has_scale = has_scale || scale[c] != 1.f;
        }
        if (!has_mean && !has_scale && !has_image) continue;
        const auto& shape = param->get_shape();
        if (shape.size() < 2 || shape[1] != channels) IE_THROW(ParameterMismatch) << "B200 plugin: " << channels << " pre-processing channels for input '" << it->first << "'";
        const auto consumers = param->output(0).get_target_inputs();
        ngraph::Output<ngraph::Node> cur = param->output(0);
        if (has_image) {
            // MEAN_IMAGE (ie_preprocess.hpp:30-45): one FP32 blob of the input's spatial size per channel, subtracted element-wise on the
            // device (the reference does the same through a Subtract: docs/template_plugin/src/transformations/preprocessing/preprocessing.cpp:67-82,
            // mkldnn_plugin/normalize_preprocess.cpp)
            if (has_mean) IE_THROW(ParameterMismatch) << "B200 plugin: input '" << it->first << "' sets both a mean image and mean values";
            ngraph::Shape ishape(shape);
            ishape[0] = 1;
            const size_t plane = ngraph::shape_size(ishape) / channels;
            std::vector<float> image(channels * plane);
            for (size_t c = 0; c < channels; ++c) {
                const auto& blob = info[c]->meanData;
                if (!blob || blob->getTensorDesc().getPrecision() != InferenceEngine::Precision::FP32 || blob->size() != plane)
                    IE_THROW(ParameterMismatch) << "B200 plugin: mean image of input '" << it->first << "', channel " << c << ": an FP32 blob of " << plane
                                                << " elements is required for every channel";
                auto mem = InferenceEngine::as<InferenceEngine::MemoryBlob>(blob);
                if (!mem) IE_THROW(ParameterMismatch) << "B200 plugin: mean image of input '" << it->first << "' is not a memory blob";
                auto locked = mem->rmap();
                std::memcpy(&image[c * plane], locked.as<const float*>(), plane * sizeof(float));
            }
            auto sub = std::make_shared<ngraph::opset1::Subtract>(cur, ngraph::opset1::Constant::create(ngraph::element::f32, ishape, image));
            sub->set_friendly_name(param->get_friendly_name() + "/mean_image");
            cur = sub->output(0);
            std::fill(mean.begin(), mean.end(), 0.f);
        }
        if (has_mean || has_scale) {
            // (x - mean) / scale as ONE per-channel scale + shift: BatchNormInference with gamma 1, beta 0, variance scale^2, eps 0 - the
            // compiler runs it as a single channels-last kernel (a Subtract + Divide pair would be two kernels and a layout conversion)
            std::vector<float> ones(channels, 1.f), zeros(channels, 0.f), var(channels);
            for (size_t c = 0; c < channels; ++c) var[c] = scale[c] * scale[c];
            const ngraph::Shape vshape{channels};
            auto bn = std::make_shared<ngraph::opset5::BatchNormInference>(cur, ngraph::opset1::Constant::create(ngraph::element::f32, vshape, ones),
                                                                          ngraph::opset1::Constant::create(ngraph::element::f32, vshape, zeros),
                                                                          ngraph::opset1::Constant::create(ngraph::element::f32, vshape, mean),
                                                                          ngraph::opset1::Constant::create(ngraph::element::f32, vshape, var), 0.0);
            bn->set_friendly_name(param->get_friendly_name() + "/mean_scale");
            cur = bn->output(0);
        }
        for (auto consumer : consumers) consumer.replace_source_output(cur);
    }
}

std::shared_ptr<ngraph::Function> ShardBatch(const std::shared_ptr<const ngraph::Function>& function, size_t shards) {
    auto f = ngraph::clone_function(*function);
    if (shards <= 1) return f;
    const auto& params = f->get_parameters();
    if (params.empty() || params[0]->get_shape().empty()) IE_THROW(ParameterMismatch) << "B200 plugin: network has no batched input to shard";
    const size_t batch = params[0]->get_shape()[0];
    if (batch % shards != 0) IE_THROW(ParameterMismatch) << "B200 plugin: batch " << batch << " does not divide over " << shards << " devices";
    const size_t part = batch / shards;
    for (const auto& p : params) {
        auto shape = p->get_shape();
        if (!shape.empty() && shape[0] == batch) {
            shape[0] = part;
            p->set_partial_shape(ngraph::PartialShape(shape));
        }
    }
    for (const auto& node : f->get_ordered_ops()) {
        if (auto reshape = as_type_ptr<ngraph::opset1::Reshape>(node)) {
            auto target = as_type_ptr<ngraph::opset1::Constant>(reshape->input_value(1).get_node_shared_ptr());
            const auto& in = reshape->get_input_partial_shape(0);
            if (target && in.rank().is_static() && in.rank().get_length() > 0 && in[0].is_static() && static_cast<size_t>(in[0].get_length()) == part) {
                auto dims = target->cast_vector<int64_t>();
                if (!dims.empty() && dims[0] == static_cast<int64_t>(batch)) {
                    dims[0] = static_cast<int64_t>(part);
                    auto replacement = ngraph::opset1::Constant::create(target->get_element_type(), target->get_shape(), dims);
                    replacement->set_friendly_name(target->get_friendly_name());
                    reshape->input(1).replace_source_output(replacement->output(0));
                }
            }
        }
        node->revalidate_and_infer_types();
    }
    return f;
}

namespace {

int ToDtype(const ngraph::element::Type& t) {
    switch (t) {
    case ngraph::element::Type_t::f32: return B200_F32;
    case ngraph::element::Type_t::f16: return B200_F16;
    case ngraph::element::Type_t::bf16: return B200_BF16;
    case ngraph::element::Type_t::i32: return B200_I32;
    case ngraph::element::Type_t::u8: return B200_U8;
    case ngraph::element::Type_t::i64: return B200_I64;
    case ngraph::element::Type_t::i8: return B200_I8;
    case ngraph::element::Type_t::boolean: return B200_BOOL;
    default: return -1;
    }
}

struct OpDesc {
    int op = -1;
    std::vector<int> data_inputs;   // which ngraph inputs are graph inputs of the b200 node (the rest are attributes)
    std::vector<int32_t> i;
    std::vector<float> f;
    bool shape_from_output = false; // RESHAPE-like: target shape = static output shape
    int dtype = B200_F32;
    std::string reject;             // non-empty: not supported, with the reason
};

std::vector<int64_t> ConstInts(const NodePtr& n, size_t input, std::string& err) {
    auto c = as_type_ptr<ngraph::opset1::Constant>(n->get_input_node_shared_ptr(input));
    if (!c) {
        err = "input " + std::to_string(input) + " of " + n->get_friendly_name() + " must be a Constant";
        return {};
    }
    return c->cast_vector<int64_t>();
}

template <typename T>
void PushSpatial(OpDesc& d, const T& v) {
    for (auto x : v) d.i.push_back(static_cast<int32_t>(x));
}

OpDesc Describe(const NodePtr& n) {
    OpDesc d;
    auto all_inputs = [&](size_t count) { for (size_t k = 0; k < count; ++k) d.data_inputs.push_back(static_cast<int>(k)); };
    auto simple = [&](int op, size_t inputs) { d.op = op; all_inputs(inputs); return d; };
    const auto& ti = n->get_type_info();
    std::string err;

    if (as_type_ptr<ngraph::opset1::Parameter>(n)) { d.op = B200_OP_PARAMETER; d.dtype = ToDtype(n->get_output_element_type(0)); return d; }
    if (as_type_ptr<ngraph::opset1::Constant>(n)) { d.op = B200_OP_CONSTANT; d.dtype = ToDtype(n->get_output_element_type(0)); return d; }
    if (as_type_ptr<ngraph::opset1::Result>(n)) return simple(B200_OP_RESULT, 1);

    if (auto c = as_type_ptr<ngraph::opset1::Convolution>(n)) {
        if (c->get_input_shape(0).size() != 4) { d.reject = "only 2D convolutions"; return d; }
        d.op = B200_OP_CONVOLUTION; all_inputs(2);
        PushSpatial(d, c->get_strides()); PushSpatial(d, c->get_dilations());
        d.i.push_back(static_cast<int32_t>(c->get_pads_begin()[0])); d.i.push_back(static_cast<int32_t>(c->get_pads_begin()[1]));
        d.i.push_back(static_cast<int32_t>(c->get_pads_end()[0])); d.i.push_back(static_cast<int32_t>(c->get_pads_end()[1]));
        return d;
    }
    if (auto c = as_type_ptr<ngraph::opset1::GroupConvolution>(n)) {
        if (c->get_input_shape(0).size() != 4) { d.reject = "only 2D group convolutions"; return d; }
        d.op = B200_OP_GROUP_CONVOLUTION; all_inputs(2);
        PushSpatial(d, c->get_strides()); PushSpatial(d, c->get_dilations());
        d.i.push_back(static_cast<int32_t>(c->get_pads_begin()[0])); d.i.push_back(static_cast<int32_t>(c->get_pads_begin()[1]));
        d.i.push_back(static_cast<int32_t>(c->get_pads_end()[0])); d.i.push_back(static_cast<int32_t>(c->get_pads_end()[1]));
        return d;
    }
    if (auto bn = as_type_ptr<ngraph::opset5::BatchNormInference>(n)) {   // x, gamma, beta, mean, var
        d.op = B200_OP_BATCH_NORM; all_inputs(5); d.f.push_back(static_cast<float>(bn->get_eps_value())); return d;
    }
    if (auto bn = as_type_ptr<ngraph::opset1::BatchNormInference>(n)) {   // v0 order: gamma, beta, x, mean, var
        d.op = B200_OP_BATCH_NORM; d.data_inputs = {2, 0, 1, 3, 4}; d.f.push_back(static_cast<float>(bn->get_eps_value())); return d;
    }
    if (as_type_ptr<ngraph::opset1::Add>(n)) return simple(B200_OP_ADD, 2);
    if (as_type_ptr<ngraph::opset1::Subtract>(n)) return simple(B200_OP_SUBTRACT, 2);
    if (as_type_ptr<ngraph::opset1::Multiply>(n)) return simple(B200_OP_MULTIPLY, 2);
    if (as_type_ptr<ngraph::opset1::Divide>(n)) return simple(B200_OP_DIVIDE, 2);
    if (as_type_ptr<ngraph::opset1::Maximum>(n)) return simple(B200_OP_MAXIMUM, 2);
    if (as_type_ptr<ngraph::opset1::Minimum>(n)) return simple(B200_OP_MINIMUM, 2);
    if (as_type_ptr<ngraph::opset1::Power>(n)) return simple(B200_OP_POWER, 2);
    if (as_type_ptr<ngraph::opset1::SquaredDifference>(n)) return simple(B200_OP_SQUARED_DIFFERENCE, 2);
    if (as_type_ptr<ngraph::opset1::PRelu>(n)) return simple(B200_OP_PRELU, 2);
    if (as_type_ptr<ngraph::opset1::Relu>(n)) return simple(B200_OP_RELU, 1);
    if (as_type_ptr<ngraph::opset1::Sigmoid>(n)) return simple(B200_OP_SIGMOID, 1);
    if (as_type_ptr<ngraph::opset1::Tanh>(n)) return simple(B200_OP_TANH, 1);
    if (as_type_ptr<ngraph::opset1::Exp>(n)) return simple(B200_OP_EXP, 1);
    if (as_type_ptr<ngraph::opset1::Sqrt>(n)) return simple(B200_OP_SQRT, 1);
    if (as_type_ptr<ngraph::opset1::Abs>(n)) return simple(B200_OP_ABS, 1);
    if (as_type_ptr<ngraph::opset1::Negative>(n)) return simple(B200_OP_NEGATIVE, 1);
    if (as_type_ptr<ngraph::opset1::Erf>(n)) return simple(B200_OP_ERF, 1);
    if (as_type_ptr<ngraph::opset4::HSwish>(n)) return simple(B200_OP_HSWISH, 1);
    if (as_type_ptr<ngraph::opset5::HSigmoid>(n)) return simple(B200_OP_HSIGMOID, 1);
    if (as_type_ptr<ngraph::opset4::Mish>(n)) return simple(B200_OP_MISH, 1);
    if (auto c = as_type_ptr<ngraph::opset1::Clamp>(n)) {
        d.op = B200_OP_CLAMP; all_inputs(1); d.f = {static_cast<float>(c->get_min()), static_cast<float>(c->get_max())}; return d;
    }
    if (auto r = as_type_ptr<ngraph::opset5::Round>(n)) {
        d.op = B200_OP_ROUND; all_inputs(1); d.i = {r->get_mode() == ngraph::opset5::Round::RoundMode::HALF_TO_EVEN ? 0 : 1}; return d;
    }
    if (auto fq = as_type_ptr<ngraph::opset1::FakeQuantize>(n)) {
        // executed in its fp32 / 16-bit form (one element-wise kernel); LPT's INT8 rewrite of the graph is not applied by this plugin
        d.op = B200_OP_FAKE_QUANTIZE; all_inputs(5); d.i = {static_cast<int>(fq->get_levels())};
        for (size_t k = 1; k < 5; ++k)
            if (!as_type_ptr<ngraph::opset1::Constant>(n->get_input_node_shared_ptr(k))) { d.reject = "FakeQuantize ranges must be Constants"; return d; }
        return d;
    }
    if (auto e = as_type_ptr<ngraph::opset1::Elu>(n)) { d.op = B200_OP_ELU; all_inputs(1); d.f = {static_cast<float>(e->get_alpha())}; return d; }
    if (as_type_ptr<ngraph::opset4::Swish>(n)) {
        d.op = B200_OP_SWISH; all_inputs(1); d.f = {1.0f};
        if (n->get_input_size() > 1) {
            auto c = as_type_ptr<ngraph::opset1::Constant>(n->get_input_node_shared_ptr(1));
            if (!c) { d.reject = "Swish beta must be a Constant"; return d; }
            d.f[0] = c->cast_vector<float>()[0];
        }
        return d;
    }
    if (auto g = as_type_ptr<ngraph::opset7::Gelu>(n)) {
        d.op = B200_OP_GELU; all_inputs(1); d.i = {g->get_approximation_mode() == ngraph::op::GeluApproximationMode::TANH ? 1 : 0}; return d;
    }
    if (as_type_ptr<ngraph::op::v0::Gelu>(n)) { d.op = B200_OP_GELU; all_inputs(1); d.i = {0}; return d; }   // v0/v2: erf form
    if (auto p = as_type_ptr<ngraph::opset1::MaxPool>(n)) {
        if (p->get_input_shape(0).size() != 4) { d.reject = "only 2D pooling"; return d; }
        d.op = B200_OP_MAX_POOL; all_inputs(1);
        PushSpatial(d, p->get_kernel()); PushSpatial(d, p->get_strides()); PushSpatial(d, p->get_pads_begin()); PushSpatial(d, p->get_pads_end());
        d.i.push_back(p->get_rounding_type() == ngraph::op::RoundingType::CEIL ? 1 : 0);
        return d;
    }
    if (auto p = as_type_ptr<ngraph::opset1::AvgPool>(n)) {
        if (p->get_input_shape(0).size() != 4) { d.reject = "only 2D pooling"; return d; }
        d.op = B200_OP_AVG_POOL; all_inputs(1);
        PushSpatial(d, p->get_kernel()); PushSpatial(d, p->get_strides()); PushSpatial(d, p->get_pads_begin()); PushSpatial(d, p->get_pads_end());
        d.i.push_back(p->get_rounding_type() == ngraph::op::RoundingType::CEIL ? 1 : 0);
        d.i.push_back(p->get_exclude_pad() ? 1 : 0);
        return d;
    }
    {
        int rop = as_type_ptr<ngraph::opset1::ReduceMean>(n) ? B200_OP_REDUCE_MEAN
                  : as_type_ptr<ngraph::opset1::ReduceSum>(n) ? B200_OP_REDUCE_SUM
                  : as_type_ptr<ngraph::opset1::ReduceMax>(n) ? B200_OP_REDUCE_MAX : -1;
        if (rop >= 0) {
            auto axes = ConstInts(n, 1, err);
            if (!err.empty()) { d.reject = err; return d; }
            bool keep = false;
            if (auto r = std::dynamic_pointer_cast<ngraph::op::util::ArithmeticReductionKeepDims>(n)) keep = r->get_keep_dims();
            d.op = rop; all_inputs(1);
            d.i = {keep ? 1 : 0, static_cast<int32_t>(axes.size())};
            for (auto a : axes) d.i.push_back(static_cast<int32_t>(a));
            return d;
        }
    }
    if (auto m = as_type_ptr<ngraph::opset1::MatMul>(n)) {
        d.op = B200_OP_MATMUL; all_inputs(2); d.i = {m->get_transpose_a() ? 1 : 0, m->get_transpose_b() ? 1 : 0}; return d;
    }
    if (auto s = as_type_ptr<ngraph::opset1::Softmax>(n)) { d.op = B200_OP_SOFTMAX; all_inputs(1); d.i = {static_cast<int32_t>(s->get_axis())}; return d; }
    if (auto m = as_type_ptr<ngraph::opset6::MVN>(n)) {
        auto axes = ConstInts(n, 1, err);
        if (!err.empty()) { d.reject = err; return d; }
        d.op = B200_OP_MVN; all_inputs(1);
        d.i = {m->get_normalize_variance() ? 1 : 0, m->get_eps_mode() == ngraph::op::MVNEpsMode::INSIDE_SQRT ? 0 : 1, static_cast<int32_t>(axes.size())};
        for (auto a : axes) d.i.push_back(static_cast<int32_t>(a));
        d.f = {m->get_eps()};
        return d;
    }
    if (as_type_ptr<ngraph::opset1::Reshape>(n) || as_type_ptr<ngraph::opset1::Squeeze>(n) || as_type_ptr<ngraph::opset1::Unsqueeze>(n)) {
        d.op = B200_OP_RESHAPE; all_inputs(1); d.shape_from_output = true; return d;
    }
    if (auto ss = as_type_ptr<ngraph::opset1::StridedSlice>(n)) {
        // opset1 StridedSlice -> ngraph::make_slice_plan (what the nGraph reference itself evaluates): a positive-stride slice of the input,
        // a non-transposing reshape (new axes / shrunk axes) and a reverse of some axes (negative strides).  The C ABI op is the slice
        // with signed strides + the free reshape; a reverse is folded into a negative stride when the reshape keeps the axes in place
        auto begin = ConstInts(n, 1, err), end = ConstInts(n, 2, err);
        std::vector<int64_t> stride(begin.size(), 1);
        if (n->get_input_size() > 3) stride = ConstInts(n, 3, err);
        if (!err.empty()) { d.reject = err; return d; }
        if (!n->get_input_partial_shape(0).is_static() || !n->get_output_partial_shape(0).is_static()) { d.reject = "dynamic StridedSlice"; return d; }
        auto axes = [](const std::vector<int64_t>& mask) {
            ngraph::AxisSet set;
            for (size_t k = 0; k < mask.size(); ++k) if (mask[k] == 1) set.emplace(k);
            return set;
        };
        ngraph::SlicePlan plan;
        try {
            plan = ngraph::make_slice_plan(n->get_input_shape(0), begin, end, stride, axes(ss->get_begin_mask()), axes(ss->get_end_mask()),
                                           axes(ss->get_new_axis_mask()), axes(ss->get_shrink_axis_mask()), axes(ss->get_ellipsis_mask()));
        } catch (const std::exception& e) {
            d.reject = std::string("StridedSlice: ") + e.what();
            return d;
        }
        const ngraph::Shape in = n->get_input_shape(0);
        const size_t r = in.size();
        if (r < 1 || r > 8 || plan.begins.size() != r || plan.reshape_in_shape.size() != r) { d.reject = "StridedSlice: unsupported rank"; return d; }
        if (!plan.reverse_axes.empty() && plan.reshape_in_shape != plan.reshape_out_shape) { d.reject = "StridedSlice: negative stride together with new / shrunk axes"; return d; }
        std::vector<int64_t> b(plan.begins), st(plan.strides);
        for (size_t k = 0; k < r; ++k) {
            if (plan.reshape_in_shape[k] == 0) { d.reject = "StridedSlice: empty result"; return d; }
            if (plan.reverse_axes.count(k)) {
                b[k] = plan.begins[k] + static_cast<int64_t>(plan.reshape_in_shape[k] - 1) * plan.strides[k];
                st[k] = -plan.strides[k];
            }
        }
        d.op = B200_OP_STRIDED_SLICE; d.data_inputs = {0}; d.shape_from_output = true;
        d.i = {static_cast<int32_t>(r)};
        for (auto v : in) d.i.push_back(static_cast<int32_t>(v));
        for (auto v : b) d.i.push_back(static_cast<int32_t>(v));
        for (auto v : st) d.i.push_back(static_cast<int32_t>(v));
        for (auto v : plan.reshape_in_shape) d.i.push_back(static_cast<int32_t>(v));
        return d;
    }
    if (as_type_ptr<ngraph::opset1::Transpose>(n)) {
        auto order = ConstInts(n, 1, err);
        if (!err.empty()) { d.reject = err; return d; }
        d.op = B200_OP_TRANSPOSE; all_inputs(1);
        const size_t rank = n->get_input_shape(0).size();
        if (order.empty()) for (size_t k = 0; k < rank; ++k) order.push_back(static_cast<int64_t>(rank - 1 - k));
        for (auto a : order) d.i.push_back(static_cast<int32_t>(a));
        return d;
    }
    if (auto c = as_type_ptr<ngraph::opset1::Concat>(n)) {
        d.op = B200_OP_CONCAT; all_inputs(n->get_input_size()); d.i = {static_cast<int32_t>(c->get_axis())}; return d;
    }
    if (auto s = as_type_ptr<ngraph::opset1::Split>(n)) {
        auto axis = ConstInts(n, 1, err);
        if (!err.empty()) { d.reject = err; return d; }
        d.op = B200_OP_SPLIT; all_inputs(1); d.i = {static_cast<int32_t>(axis[0]), static_cast<int32_t>(s->get_num_splits())}; return d;
    }
    if (as_type_ptr<ngraph::opset1::Gather>(n) || as_type_ptr<ngraph::opset7::Gather>(n)) {
        auto axis = ConstInts(n, 2, err);
        if (!err.empty()) { d.reject = err; return d; }
        if (auto g7 = as_type_ptr<ngraph::opset7::Gather>(n))
            if (g7->get_batch_dims() != 0) { d.reject = "Gather batch_dims != 0"; return d; }
        d.op = B200_OP_GATHER; d.data_inputs = {0, 1}; d.i = {static_cast<int32_t>(axis[0])}; return d;
    }
    {
        auto t1 = as_type_ptr<ngraph::opset1::TopK>(n);
        auto t3 = as_type_ptr<ngraph::opset3::TopK>(n);
        if (t1 || t3) {
            auto k = ConstInts(n, 1, err);
            if (!err.empty()) { d.reject = err; return d; }
            const auto mode = t1 ? t1->get_mode() : t3->get_mode();
            const auto sort = t1 ? t1->get_sort_type() : t3->get_sort_type();
            const int64_t axis = static_cast<int64_t>(t1 ? t1->get_axis() : t3->get_axis());
            d.op = B200_OP_TOPK; all_inputs(1);
            d.i = {static_cast<int32_t>(k[0]), static_cast<int32_t>(axis), mode == ngraph::op::TopKMode::MAX ? 0 : 1,
                   sort == ngraph::op::TopKSortType::NONE ? 0 : (sort == ngraph::op::TopKSortType::SORT_VALUES ? 1 : 2)};
            return d;
        }
    }
    if (auto nms = as_type_ptr<ngraph::opset5::NonMaxSuppression>(n)) {
        for (size_t k = 2; k < n->get_input_size(); ++k)
            if (!as_type_ptr<ngraph::opset1::Constant>(n->get_input_node_shared_ptr(k))) { d.reject = "NonMaxSuppression thresholds must be Constants"; return d; }
        d.op = B200_OP_NMS; d.data_inputs = {0, 1};
        d.i = {static_cast<int32_t>(std::min<int64_t>(nms->max_boxes_output_from_input(), INT32_MAX)),
               nms->get_box_encoding() == ngraph::opset5::NonMaxSuppression::BoxEncodingType::CENTER ? 1 : 0, nms->get_sort_result_descending() ? 1 : 0};
        d.f = {nms->iou_threshold_from_input(), nms->score_threshold_from_input(), nms->soft_nms_sigma_from_input()};
        return d;
    }
    if (auto nms = as_type_ptr<ngraph::op::internal::NonMaxSuppressionIEInternal>(n)) {
        std::vector<float> thr;
        for (size_t k = 2; k < n->get_input_size(); ++k) {
            auto c = as_type_ptr<ngraph::opset1::Constant>(n->get_input_node_shared_ptr(k));
            if (!c) { d.reject = "NonMaxSuppression thresholds must be Constants"; return d; }
            thr.push_back(c->cast_vector<float>()[0]);
        }
        while (thr.size() < 4) thr.push_back(0.f);
        d.op = B200_OP_NMS; d.data_inputs = {0, 1};
        d.i = {static_cast<int32_t>(std::min<float>(thr[0], 2147483520.f)), nms->m_center_point_box ? 1 : 0, nms->m_sort_result_descending ? 1 : 0};
        d.f = {thr[1], thr[2], thr[3]};
        return d;
    }
    if (auto det = as_type_ptr<ngraph::opset1::DetectionOutput>(n)) {
        if (n->get_input_size() != 3) { d.reject = "DetectionOutput with auxiliary inputs"; return d; }
        const auto& a = det->get_attrs();
        d.op = B200_OP_DETECTION_OUTPUT; all_inputs(3);
        const bool center = a.code_type.size() >= 11 && a.code_type.compare(a.code_type.size() - 11, 11, "CENTER_SIZE") == 0;
        d.i = {a.num_classes, a.background_label_id, a.top_k, a.keep_top_k.empty() ? -1 : a.keep_top_k[0], center ? 1 : 0, a.share_location ? 1 : 0,
               a.variance_encoded_in_target ? 1 : 0, a.normalized ? 1 : 0, a.clip_before_nms ? 1 : 0, a.clip_after_nms ? 1 : 0,
               static_cast<int32_t>(a.input_height), static_cast<int32_t>(a.input_width), a.decrease_label_id ? 1 : 0};
        d.f = {a.nms_threshold, a.confidence_threshold};
        return d;
    }
    if (auto c = as_type_ptr<ngraph::opset1::Convert>(n)) {
        d.op = B200_OP_CONVERT; all_inputs(1); d.dtype = ToDtype(c->get_destination_type());
        if (d.dtype < 0) d.reject = "Convert to " + c->get_destination_type().get_type_name();
        return d;
    }
    d.reject = std::string("no B200 kernel for ") + ti.name + " (opset version " + std::to_string(ti.version) + ")";
    return d;
}

}  // namespace

bool IsNodeSupported(const NodePtr& node, std::string* why) {
    OpDesc d = Describe(node);
    bool ok = d.reject.empty() && d.op >= 0 && b200_op_supported(d.op) == 1;
    for (size_t k = 0; ok && k < node->get_output_size(); ++k) {
        if (node->get_output_partial_shape(k).is_dynamic()) { ok = false; d.reject = "dynamic shapes are not supported"; }
        else if (ToDtype(node->get_output_element_type(k)) < 0) { ok = false; d.reject = "element type " + node->get_output_element_type(k).get_type_name(); }
    }
    if (!ok && why) *why = d.reject.empty() ? "unsupported operation" : d.reject;
    return ok;
}

// QueryNetwork: the op id alone does not say whether the kernels take THIS geometry (DetectionOutput with too many priors, TopK with a
// huge k, a MatMul batch broadcast, ...).  The node is rebuilt on its own - Parameters for its non-constant inputs, its constants
// kept - and pushed through the real compiler without a device (b200_compile, device -1); what LoadNetwork would refuse is refused here,
// so HETERO can place the node elsewhere (the reference fails such layers at LoadNetwork too, never at the first Infer).
bool DryCompileNode(const NodePtr& node, std::string* why) {
    if (ngraph::op::is_parameter(node) || ngraph::op::is_constant(node) || ngraph::op::is_output(node)) return true;
    try {
        ngraph::OutputVector inputs;
        ngraph::ParameterVector params;
        for (const auto& in : node->input_values()) {
            if (ngraph::op::is_constant(in.get_node())) {
                inputs.push_back(in);
            } else {
                auto p = std::make_shared<ngraph::opset1::Parameter>(in.get_element_type(), in.get_partial_shape());
                params.push_back(p);
                inputs.push_back(p->output(0));
            }
        }
        auto clone = node->clone_with_new_inputs(inputs);
        ngraph::ResultVector results;
        for (const auto& out : clone->outputs()) results.push_back(std::make_shared<ngraph::opset1::Result>(out));
        auto f = std::make_shared<ngraph::Function>(results, params, "query_" + node->get_friendly_name());
        b200_graph* g = LowerFunction(f);
        b200_compile_options opts;
        std::memset(&opts, 0, sizeof(opts));
        opts.abi_version = B200_ABI_VERSION;
        opts.device = -1;
        opts.precision = B200_PREC_BF16;
        b200_network* net = nullptr;
        const int rc = b200_compile(g, &opts, &net);
        b200_graph_destroy(g);
        if (net) b200_network_destroy(net);
        if (rc < 0) {
            if (why) *why = b200_last_error();
            return false;
        }
        return true;
    } catch (const std::exception& e) {
        if (why) *why = e.what();
        return false;
    }
}

b200_graph* LowerFunction(const std::shared_ptr<ngraph::Function>& function) {
    b200_graph* g = nullptr;
    if (b200_graph_create(function->get_friendly_name().c_str(), &g) != B200_OK) IE_THROW() << "B200: " << b200_last_error();
    std::map<std::pair<ngraph::Node*, size_t>, int32_t> value_of;
    try {
        for (const auto& node : function->get_ordered_ops()) {
            std::string why;
            if (!IsNodeSupported(node, &why))
                IE_THROW(NotImplemented) << "B200 plugin: operation '" << node->get_friendly_name() << "' (" << node->get_type_name() << ") is not supported: " << why;
            OpDesc d = Describe(node);
            b200_node_desc nd;
            std::memset(&nd, 0, sizeof(nd));
            nd.op = d.op;
            const std::string name = node->get_friendly_name();
            nd.name = name.c_str();
            if (d.data_inputs.size() > B200_MAX_INPUTS) IE_THROW(NotImplemented) << "B200 plugin: '" << name << "' has more than " << B200_MAX_INPUTS << " inputs";
            nd.n_inputs = static_cast<int32_t>(d.data_inputs.size());
            for (size_t k = 0; k < d.data_inputs.size(); ++k) {
                auto src = node->input_value(d.data_inputs[k]);
                auto it = value_of.find({src.get_node(), src.get_index()});
                if (it == value_of.end()) IE_THROW() << "B200 plugin: input " << k << " of '" << name << "' was not lowered";
                nd.inputs[k] = it->second;
            }
            if (d.i.size() > B200_MAX_IATTR || d.f.size() > B200_MAX_FATTR) IE_THROW(NotImplemented) << "B200 plugin: too many attributes on '" << name << "'";
            for (size_t k = 0; k < d.i.size(); ++k) nd.i[k] = d.i[k];
            for (size_t k = 0; k < d.f.size(); ++k) nd.f[k] = d.f[k];
            nd.dtype = d.dtype;
            if (d.op == B200_OP_PARAMETER || d.op == B200_OP_CONSTANT || d.shape_from_output) {
                const auto& shape = node->get_output_shape(0);
                if (shape.size() > B200_MAX_RANK) IE_THROW(NotImplemented) << "B200 plugin: rank > " << B200_MAX_RANK << " on '" << name << "'";
                nd.rank = static_cast<int32_t>(shape.size());
                for (size_t k = 0; k < shape.size(); ++k) nd.shape[k] = static_cast<int64_t>(shape[k]);
            }
            if (d.op == B200_OP_CONSTANT) {
                auto c = as_type_ptr<ngraph::opset1::Constant>(node);
                nd.data = c->get_data_ptr();
                nd.data_bytes = ngraph::shape_size(c->get_shape()) * c->get_element_type().size();
            }
            int32_t outs[B200_MAX_INPUTS];
            int32_t n_out = 0;
            int rc = b200_graph_add(g, &nd, outs, &n_out);
            if (rc == B200_ERR_UNSUPPORTED) IE_THROW(NotImplemented) << "B200 plugin: '" << name << "': " << b200_last_error();
            if (rc != B200_OK) IE_THROW() << "B200 plugin: '" << name << "': " << b200_last_error();
            for (int32_t k = 0; k < n_out && static_cast<size_t>(k) < node->get_output_size(); ++k) value_of[{node.get(), static_cast<size_t>(k)}] = outs[k];
        }
    } catch (...) {
        b200_graph_destroy(g);
        throw;
    }
    return g;
}

}  // namespace B200Plugin
