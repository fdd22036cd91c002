// ngraph::Function -> b200_graph (C ABI of libb200kernels.so): the plugin-side half of the graph compiler.
// Plays the role of MKLDNNGraph::Replicate (inference-engine/src/mkldnn_plugin/mkldnn_graph.cpp:181-320): one
// b200_graph_add call per ngraph::Node; fusion, layout selection and memory planning happen below the C ABI.
#pragma once
#include <memory>
#include <string>

#include <ie_input_info.hpp>
#include <ngraph/function.hpp>

#include "b200.h"

namespace B200Plugin {

// Runs the reused nGraph pipeline (CommonOptimizations with the decompositions of natively supported ops switched off,
// the analogue of Transformation() in mkldnn_plugin.cpp:115-383) on a clone of `function`.
std::shared_ptr<ngraph::Function> TransformNetwork(const std::shared_ptr<const ngraph::Function>& function);

// Clone of `function` whose batch (dim 0 of the first Parameter) is divided by `shards`: every Parameter that carries the batch is
// re-shaped, Reshape target-shape constants that spell the batch out are rewritten, and the clone is re-validated node by node (the
// plugin-side analogue of CNNNetwork::setBatchSize / reshape).  Throws ParameterMismatch if the batch does not divide.
std::shared_ptr<ngraph::Function> ShardBatch(const std::shared_ptr<const ngraph::Function>& function, size_t shards);

// PreProcessInfo mean / scale (ie_preprocess.hpp:23-40: per-channel meanValue and stdScale) become (x - mean) / scale nodes behind the
// Parameter, so they run on the device inside the request's CUDA graph (the CPU plugin normalises in MKLDNNGraph::PushInputData,
// mkldnn_graph.cpp:708-760; the template plugin inserts the same nodes, docs/template_plugin/src/transformations/preprocessing).
// Resize and colour conversion stay with the reference's host pre-processing library.  Mean IMAGES are refused (NotImplemented).
void AddMeanScale(const std::shared_ptr<ngraph::Function>& function, const InferenceEngine::InputsDataMap& inputs);

// true if `node` (after TransformNetwork) maps onto a b200_op; `why` receives the reason otherwise.
bool IsNodeSupported(const std::shared_ptr<ngraph::Node>& node, std::string* why = nullptr);

// true if the kernels take the node's actual geometry: the node alone is compiled without a device (QueryNetwork).
bool DryCompileNode(const std::shared_ptr<ngraph::Node>& node, std::string* why = nullptr);

// Builds the b200_graph; throws InferenceEngine::NotImplemented / GeneralError with the library's message on failure.
// The caller owns the returned graph (b200_graph_destroy).
b200_graph* LowerFunction(const std::shared_ptr<ngraph::Function>& function);

}  // namespace B200Plugin
