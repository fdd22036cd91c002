// B200 device plugin: Plugin / ExecutableNetwork / InferRequest / AsyncInferRequest.  See b200_plugin.hpp.
#include "b200_plugin.hpp"

#include <blob_factory.hpp>
#include <exec_graph_info.hpp>
#include <ie_algorithm.hpp>
#include <ie_metric_helpers.hpp>
#include <ie_ngraph_utils.hpp>
#include <ie_plugin_config.hpp>
#include <threading/ie_executor_manager.hpp>

#include <ngraph/opsets/opset1.hpp>
#include <ngraph/variant.hpp>
#include <transformations/rt_info/fused_names_attribute.hpp>
#include <unordered_set>

#include "b200/b200_config.hpp"
#include "b200_lowering.hpp"

using namespace InferenceEngine;

namespace B200Plugin {

// ---- error mapping: C-ABI status -> IE exception types (ie_common.h:231-249) -------------------------------
void ThrowFromStatus(int status, const std::string& what) {
    const std::string msg = "B200 plugin: " + what + ": " + b200_last_error();
    switch (status) {
    case B200_ERR_UNSUPPORTED: IE_THROW(NotImplemented) << msg;
    case B200_ERR_NO_DEVICE: IE_THROW(NotFound) << msg;
    case B200_ERR_BUSY: IE_THROW(RequestBusy) << msg;
    case B200_ERR_INVALID: IE_THROW(ParameterMismatch) << msg;
    default: IE_THROW(GeneralError) << msg;
    }
}
#define B200_CALL(expr, what)                          \
    do {                                               \
        int _rc = (expr);                              \
        if (_rc < 0) ::B200Plugin::ThrowFromStatus(_rc, what); \
    } while (0)

// ---- configuration ----------------------------------------------------------------------------------------------
Configuration::Configuration(const std::map<std::string, std::string>& config, const Configuration& defaults, bool throwOnUnsupported) {
    *this = defaults;
    for (const auto& kv : config) {
        const auto& key = kv.first;
        const auto& value = kv.second;
        if (key == CONFIG_KEY(DEVICE_ID)) {
            try {
                deviceId = value.empty() ? 0 : std::stoi(value);
            } catch (...) {
                IE_THROW(ParameterMismatch) << "B200 plugin: bad DEVICE_ID '" << value << "'";
            }
            int count = 0;
            b200_device_count(&count);
            if (deviceId < 0 || (count > 0 && deviceId >= count)) IE_THROW(NotFound) << "B200 plugin: DEVICE_ID " << deviceId << " out of range (" << count << " devices)";
        } else if (key == CONFIG_KEY(PERF_COUNT)) {
            if (value != CONFIG_VALUE(YES) && value != CONFIG_VALUE(NO)) IE_THROW(ParameterMismatch) << "B200 plugin: PERF_COUNT must be YES or NO";
            perfCount = value == CONFIG_VALUE(YES);
        } else if (key == B200ConfigParams::KEY_B200_INFERENCE_PRECISION) {
            if (value == "FP32") precision = B200_PREC_FP32;
            else if (value == "BF16") precision = B200_PREC_BF16;
            else if (value == "FP16") precision = B200_PREC_FP16;
            else IE_THROW(ParameterMismatch) << "B200 plugin: B200_INFERENCE_PRECISION must be FP32, BF16 or FP16, got '" << value << "'";
        } else if (key == B200ConfigParams::KEY_B200_THROUGHPUT_STREAMS) {
            try {
                streams = std::stoi(value);
            } catch (...) {
                streams = 0;
            }
            if (streams < 1 || streams > 64) IE_THROW(ParameterMismatch) << "B200 plugin: B200_THROUGHPUT_STREAMS must be in [1, 64], got '" << value << "'";
        } else if (key == "B200_DRY_COMPILE") {
            // test hook (not advertised in SUPPORTED_CONFIG_KEYS): lower + plan without touching a device, so that the host side
            // of LoadNetwork / metrics / exec-graph can be exercised on a machine without a GPU; requests cannot be created
            dryCompile = value == CONFIG_VALUE(YES);
        } else if (throwOnUnsupported) {
            IE_THROW(NotFound) << "Unsupported property " << key << " by B200 plugin";   // config.cpp:112-114 of the CPU plugin
        }
    }
}

std::string Configuration::PrecisionName() const { return precision == B200_PREC_FP32 ? "FP32" : (precision == B200_PREC_BF16 ? "BF16" : "FP16"); }

std::vector<std::string> Configuration::SupportedKeys() {
    return {CONFIG_KEY(DEVICE_ID), CONFIG_KEY(PERF_COUNT), B200ConfigParams::KEY_B200_INFERENCE_PRECISION, B200ConfigParams::KEY_B200_THROUGHPUT_STREAMS};
}

Parameter Configuration::Get(const std::string& name) const {
    if (name == CONFIG_KEY(DEVICE_ID)) return {std::to_string(deviceId)};
    if (name == CONFIG_KEY(PERF_COUNT)) return {std::string(perfCount ? CONFIG_VALUE(YES) : CONFIG_VALUE(NO))};
    if (name == B200ConfigParams::KEY_B200_INFERENCE_PRECISION) return {PrecisionName()};
    if (name == B200ConfigParams::KEY_B200_THROUGHPUT_STREAMS) return {std::to_string(streams)};
    IE_THROW(NotFound) << "Unsupported property " << name << " by B200 plugin";
}

// ---- plugin -------------------------------------------------------------------------------------------------------
Plugin::Plugin() {
    _pluginName = "B200";
    // wait stages block a thread on a CUDA event: they get their own executor so that start stages are never starved
    _waitExecutor = ExecutorManager::getInstance()->getIdleCPUStreamsExecutor({"B200WaitExecutor", 4});
}

Plugin::~Plugin() {
    ExecutorManager::getInstance()->clear("B200StreamsExecutor");
    ExecutorManager::getInstance()->clear("B200WaitExecutor");
}

void Plugin::SetConfig(const std::map<std::string, std::string>& config) { _cfg = Configuration{config, _cfg}; }

Parameter Plugin::GetConfig(const std::string& name, const std::map<std::string, Parameter>&) const { return _cfg.Get(name); }

void Plugin::AddExtension(const std::shared_ptr<IExtension>&) { IE_THROW(NotImplemented) << "B200 plugin does not support extensions"; }

Parameter Plugin::GetMetric(const std::string& name, const std::map<std::string, Parameter>& options) const {
    if (name == METRIC_KEY(SUPPORTED_METRICS)) {
        IE_SET_METRIC_RETURN(SUPPORTED_METRICS, std::vector<std::string>{METRIC_KEY(AVAILABLE_DEVICES), METRIC_KEY(SUPPORTED_METRICS), METRIC_KEY(SUPPORTED_CONFIG_KEYS),
                                                                         METRIC_KEY(FULL_DEVICE_NAME), METRIC_KEY(DEVICE_ARCHITECTURE),
                                                                         METRIC_KEY(OPTIMIZATION_CAPABILITIES), METRIC_KEY(RANGE_FOR_ASYNC_INFER_REQUESTS),
                                                                         METRIC_KEY(RANGE_FOR_STREAMS)});
    } else if (name == METRIC_KEY(SUPPORTED_CONFIG_KEYS)) {
        IE_SET_METRIC_RETURN(SUPPORTED_CONFIG_KEYS, Configuration::SupportedKeys());
    } else if (name == METRIC_KEY(AVAILABLE_DEVICES)) {
        int count = 0;
        b200_device_count(&count);
        std::vector<std::string> ids;
        for (int k = 0; k < count; ++k) ids.push_back(std::to_string(k));
        IE_SET_METRIC_RETURN(AVAILABLE_DEVICES, ids);
    } else if (name == METRIC_KEY(FULL_DEVICE_NAME)) {
        int dev = _cfg.deviceId;
        auto it = options.find(CONFIG_KEY(DEVICE_ID));
        if (it != options.end()) dev = std::stoi(it->second.as<std::string>());
        char buf[256] = "NVIDIA B200 (no device visible)";
        b200_device_name(dev, buf, sizeof(buf));
        IE_SET_METRIC_RETURN(FULL_DEVICE_NAME, std::string(buf));
    } else if (name == METRIC_KEY(DEVICE_ARCHITECTURE)) {
        IE_SET_METRIC_RETURN(DEVICE_ARCHITECTURE, std::string("sm_100a"));
    } else if (name == METRIC_KEY(OPTIMIZATION_CAPABILITIES)) {
        IE_SET_METRIC_RETURN(OPTIMIZATION_CAPABILITIES, std::vector<std::string>{METRIC_VALUE(FP32), METRIC_VALUE(BF16), METRIC_VALUE(FP16)});
    } else if (name == METRIC_KEY(RANGE_FOR_ASYNC_INFER_REQUESTS)) {
        IE_SET_METRIC_RETURN(RANGE_FOR_ASYNC_INFER_REQUESTS, std::make_tuple(1u, 16u, 1u));
    } else if (name == METRIC_KEY(RANGE_FOR_STREAMS)) {
        IE_SET_METRIC_RETURN(RANGE_FOR_STREAMS, std::make_tuple(1u, 8u));
    }
    IE_THROW(NotFound) << "Unsupported device metric: " << name;
}

IExecutableNetworkInternal::Ptr Plugin::LoadExeNetworkImpl(const CNNNetwork& network, const std::map<std::string, std::string>& config) {
    auto function = network.getFunction();
    if (!function) IE_THROW(NotImplemented) << "B200 plugin supports only nGraph-based networks (same stance as mkldnn_exec_network.cpp:46-48)";
    // user-side precisions the request objects can convert on the device (mkldnn_plugin.cpp:389-405 keeps a similar whitelist)
    for (const auto& in : network.getInputsInfo()) {
        const auto p = in.second->getPrecision();
        if (p != Precision::FP32 && p != Precision::FP16 && p != Precision::BF16 && p != Precision::U8 && p != Precision::I32)
            IE_THROW(NotImplemented) << "B200 plugin: input '" << in.first << "' has unsupported precision " << p;
    }
    Configuration cfg{config, _cfg};
    return std::make_shared<ExecutableNetwork>(function, cfg, std::static_pointer_cast<Plugin>(shared_from_this()));
}

// docs/template_plugin/src/template_plugin.cpp:117-195 describes the recipe: transform a clone, mark an original layer supported
// iff every transformed node that carries its name in the fused-names rt_info is supported.
QueryNetworkResult Plugin::QueryNetwork(const CNNNetwork& network, const std::map<std::string, std::string>& config) const {
    Configuration cfg{config, _cfg, false};
    (void)cfg;
    auto function = network.getFunction();
    if (!function) IE_THROW(NotImplemented) << "B200 plugin supports only nGraph-based networks";
    std::unordered_set<std::string> original;
    for (const auto& node : function->get_ops()) original.insert(node->get_friendly_name());
    auto transformed = TransformNetwork(function);
    std::unordered_set<std::string> supported, unsupported;
    for (const auto& node : transformed->get_ops()) {
        const bool ok = IsNodeSupported(node);
        for (const auto& fused : ngraph::getFusedNamesVector(node))
            if (original.count(fused)) (ok ? supported : unsupported).insert(fused);
    }
    for (const auto& name : unsupported) supported.erase(name);
    for (const auto& node : function->get_ops()) {   // constants / parameters / results follow their consumers / producers
        if (supported.count(node->get_friendly_name())) {
            for (const auto& in : node->input_values())
                if (ngraph::op::is_constant(in.get_node()) || ngraph::op::is_parameter(in.get_node())) supported.insert(in.get_node()->get_friendly_name());
            for (const auto& out : node->outputs())
                for (const auto& target : out.get_target_inputs())
                    if (ngraph::op::is_output(target.get_node())) supported.insert(target.get_node()->get_friendly_name());
        }
    }
    for (const auto& node : function->get_ops()) {
        if (ngraph::op::is_constant(node) || ngraph::op::is_parameter(node)) {
            bool used = false;
            for (const auto& target : node->output(0).get_target_inputs()) used |= supported.count(target.get_node()->get_friendly_name()) != 0;
            if (!used) supported.erase(node->get_friendly_name());
        } else if (ngraph::op::is_output(node)) {
            if (!supported.count(node->input_value(0).get_node()->get_friendly_name())) supported.erase(node->get_friendly_name());
        }
    }
    QueryNetworkResult res;
    for (const auto& name : supported) res.supportedLayersMap.emplace(name, GetName());
    return res;
}

// ---- executable network -----------------------------------------------------------------------------------------------
ExecutableNetwork::ExecutableNetwork(const std::shared_ptr<const ngraph::Function>& function, const Configuration& cfg, const std::shared_ptr<Plugin>& plugin)
    : ExecutableNetworkThreadSafeDefault(nullptr, nullptr), _cfg(cfg), _pluginRef(plugin), _name(function->get_friendly_name()) {
    auto transformed = TransformNetwork(function);
    b200_graph* graph = LowerFunction(transformed);
    b200_compile_options opts;
    std::memset(&opts, 0, sizeof(opts));
    opts.abi_version = B200_ABI_VERSION;
    opts.device = cfg.dryCompile ? -1 : cfg.deviceId;
    opts.precision = cfg.precision;
    int rc = b200_compile(graph, &opts, &_net);
    b200_graph_destroy(graph);
    if (rc < 0) ThrowFromStatus(rc, "cannot compile network '" + _name + "'");
    b200_tensor_info info;
    for (int k = 0; k < b200_network_num_inputs(_net); ++k) {
        B200_CALL(b200_network_io_info(_net, 0, k, &info), "io info");
        _inputIndex[info.name] = k;
    }
    for (int k = 0; k < b200_network_num_outputs(_net); ++k) {
        B200_CALL(b200_network_io_info(_net, 1, k, &info), "io info");
        _outputIndex[info.name] = k;
    }
    // start stages are cheap host work (enqueue); one executor thread per stream keeps submission order per stream
    auto streamsConfig = IStreamsExecutor::Config{"B200StreamsExecutor", std::max(1, cfg.streams)};
    _taskExecutor = ExecutorManager::getInstance()->getIdleCPUStreamsExecutor(streamsConfig);
    _callbackExecutor = _taskExecutor;
}

ExecutableNetwork::~ExecutableNetwork() {
    if (_net) b200_network_destroy(_net);
}

int ExecutableNetwork::inputIndex(const std::string& name) const {
    auto it = _inputIndex.find(name);
    if (it == _inputIndex.end()) IE_THROW(NotFound) << "B200 plugin: network has no input '" << name << "'";
    return it->second;
}
int ExecutableNetwork::outputIndex(const std::string& name) const {
    auto it = _outputIndex.find(name);
    if (it == _outputIndex.end()) IE_THROW(NotFound) << "B200 plugin: network has no output '" << name << "'";
    return it->second;
}

IInferRequestInternal::Ptr ExecutableNetwork::CreateInferRequestImpl(InputsDataMap networkInputs, OutputsDataMap networkOutputs) {
    return std::make_shared<InferRequest>(networkInputs, networkOutputs, std::static_pointer_cast<ExecutableNetwork>(shared_from_this()));
}

IInferRequestInternal::Ptr ExecutableNetwork::CreateInferRequest() {
    auto sync = CreateInferRequestImpl(_networkInputs, _networkOutputs);
    sync->setPointerToExecutableNetworkInternal(shared_from_this());
    return std::make_shared<AsyncInferRequest>(std::static_pointer_cast<InferRequest>(sync), _taskExecutor, _pluginRef->waitExecutor(), _callbackExecutor);
}

Parameter ExecutableNetwork::GetMetric(const std::string& name) const {
    if (name == EXEC_NETWORK_METRIC_KEY(SUPPORTED_METRICS)) {
        IE_SET_METRIC_RETURN(SUPPORTED_METRICS, std::vector<std::string>{EXEC_NETWORK_METRIC_KEY(NETWORK_NAME), EXEC_NETWORK_METRIC_KEY(SUPPORTED_METRICS),
                                                                         EXEC_NETWORK_METRIC_KEY(SUPPORTED_CONFIG_KEYS),
                                                                         EXEC_NETWORK_METRIC_KEY(OPTIMAL_NUMBER_OF_INFER_REQUESTS)});
    } else if (name == EXEC_NETWORK_METRIC_KEY(SUPPORTED_CONFIG_KEYS)) {
        IE_SET_METRIC_RETURN(SUPPORTED_CONFIG_KEYS, Configuration::SupportedKeys());
    } else if (name == EXEC_NETWORK_METRIC_KEY(NETWORK_NAME)) {
        IE_SET_METRIC_RETURN(NETWORK_NAME, _name);
    } else if (name == EXEC_NETWORK_METRIC_KEY(OPTIMAL_NUMBER_OF_INFER_REQUESTS)) {
        // two requests per stream: the H2D copy of request k+1 overlaps the kernels of request k
        IE_SET_METRIC_RETURN(OPTIMAL_NUMBER_OF_INFER_REQUESTS, static_cast<unsigned int>(2 * std::max(1, _cfg.streams)));
    }
    IE_THROW(NotFound) << "Unsupported ExecutableNetwork metric: " << name;
}

Parameter ExecutableNetwork::GetConfig(const std::string& name) const { return _cfg.Get(name); }

// One ExecutionNode per kernel of the compiled schedule with the rt_info keys of plugin_api/exec_graph_info.hpp:34-76
// (the reference fills them in mkldnn_graph_dumper.cpp); inputs are chained in execution order.
CNNNetwork ExecutableNetwork::GetExecGraphInfo() {
    using namespace ExecGraphInfoSerialization;
    ngraph::ParameterVector params;
    ngraph::ResultVector results;
    auto src = std::make_shared<ngraph::opset1::Parameter>(ngraph::element::f32, ngraph::Shape{1});
    src->set_friendly_name(_name + "_input");
    src->get_rt_info()[LAYER_TYPE] = std::make_shared<ngraph::VariantWrapper<std::string>>("Input");
    src->get_rt_info()[EXECUTION_ORDER] = std::make_shared<ngraph::VariantWrapper<std::string>>("0");
    params.push_back(src);
    ngraph::Output<ngraph::Node> prev = src;
    const int n = b200_network_num_exec_nodes(_net);
    b200_exec_node_info info;
    for (int k = 0; k < n; ++k) {
        B200_CALL(b200_network_exec_node(_net, k, &info), "exec node info");
        auto node = std::make_shared<ExecutionNode>(ngraph::OutputVector{prev}, 1);
        node->set_friendly_name(info.name);
        auto set = [&](const char* key, const std::string& v) { node->get_rt_info()[key] = std::make_shared<ngraph::VariantWrapper<std::string>>(v); };
        set(ORIGINAL_NAMES, info.fused);
        set(IMPL_TYPE, info.impl);
        set(LAYER_TYPE, info.type);
        set(OUTPUT_PRECISIONS, info.out_precision);
        set(RUNTIME_PRECISION, info.out_precision);
        set(OUTPUT_LAYOUTS, info.out_layout);
        set(EXECUTION_ORDER, std::to_string(info.exec_order + 1));
        set(PERF_COUNTER, "not_executed");
        node->set_output_type(0, ngraph::element::f32, ngraph::PartialShape{1});
        prev = node;
    }
    auto res = std::make_shared<ngraph::opset1::Result>(prev);
    res->set_friendly_name(_name + "_output");
    res->get_rt_info()[LAYER_TYPE] = std::make_shared<ngraph::VariantWrapper<std::string>>("Output");
    res->get_rt_info()[EXECUTION_ORDER] = std::make_shared<ngraph::VariantWrapper<std::string>>(std::to_string(n + 1));
    results.push_back(res);
    return CNNNetwork(std::make_shared<ngraph::Function>(results, params, _name));
}

// ---- infer request ----------------------------------------------------------------------------------------------------
namespace {

int BlobDtype(const Precision& p, const std::string& name) {
    switch (p) {
    case Precision::FP32: return B200_F32;
    case Precision::FP16: return B200_F16;
    case Precision::BF16: return B200_BF16;
    case Precision::I32: return B200_I32;
    case Precision::U8: return B200_U8;
    default: IE_THROW(NotImplemented) << "B200 plugin: blob '" << name << "' has unsupported precision " << p;
    }
}

int BlobLayout(const TensorDesc& desc, const std::string& name) {
    if (desc.getDims().size() != 4) return B200_LAYOUT_ANY;     // row-major
    if (desc.getLayout() == Layout::NHWC) return B200_LAYOUT_NHWC;
    if (desc.getLayout() == Layout::NCHW) return B200_LAYOUT_NCHW;
    IE_THROW(NotImplemented) << "B200 plugin: blob '" << name << "' has unsupported layout " << desc.getLayout();
}

const void* BlobData(const Blob::Ptr& blob) {
    auto mem = as<MemoryBlob>(blob);
    if (!mem) IE_THROW(NotImplemented) << "B200 plugin: only MemoryBlob inputs/outputs are supported";
    return mem->rmap().as<const uint8_t*>();   // includes the ROI / padding offset of the descriptor
}

}  // namespace

InferRequest::InferRequest(const InputsDataMap& networkInputs, const OutputsDataMap& networkOutputs, const std::shared_ptr<ExecutableNetwork>& network)
    : IInferRequestInternal(networkInputs, networkOutputs), _network(network) {
    B200_CALL(b200_request_create(network->handle(), &_req), "cannot create infer request");
    _inputNames.resize(_networkInputs.size());
    _outputNames.resize(_networkOutputs.size());
    // default blobs (GetBlob before SetBlob must work: ie_iinfer_request_internal.cpp:111-143) live in pinned host memory so
    // the H2D/D2H copies of start() are asynchronous and run at PCIe speed
    for (const auto& in : _networkInputs) {
        const int idx = network->inputIndex(in.first);
        const TensorDesc& desc = in.second->getTensorDesc();
        B200_CALL(b200_request_set_io_format(_req, 0, idx, BlobDtype(desc.getPrecision(), in.first), BlobLayout(desc, in.first)), "input format");
        _inputNames[idx] = in.first;
        _inputs[in.first] = allocatePinned(desc);
    }
    for (const auto& out : _networkOutputs) {
        const int idx = network->outputIndex(out.first);
        const TensorDesc& desc = out.second->getTensorDesc();
        B200_CALL(b200_request_set_io_format(_req, 1, idx, BlobDtype(desc.getPrecision(), out.first), BlobLayout(desc, out.first)), "output format");
        _outputNames[idx] = out.first;
        _outputs[out.first] = allocatePinned(desc);
    }
}

InferRequest::~InferRequest() {
    if (_req) b200_request_destroy(_req);   // drains the stream first
    _inputs.clear();
    _outputs.clear();
    for (void* p : _pinned) b200_host_free(p);
}

Blob::Ptr InferRequest::allocatePinned(const TensorDesc& desc) {
    size_t elems = 1;
    for (auto d : desc.getDims()) elems *= d;
    void* ptr = nullptr;
    B200_CALL(b200_host_alloc(&ptr, std::max<size_t>(elems * desc.getPrecision().size(), 16)), "pinned allocation");
    _pinned.push_back(ptr);
    return make_blob_with_precision(desc, ptr);
}

void InferRequest::preprocessAndStart() {
    execDataPreprocessing(_inputs);     // resize / colour conversion requested through PreProcessInfo (host, reused from the reference)
    std::vector<const void*> ins(_inputNames.size(), nullptr);
    std::vector<void*> outs(_outputNames.size(), nullptr);
    for (size_t k = 0; k < _inputNames.size(); ++k) {
        const Blob::Ptr& blob = _inputs.at(_inputNames[k]);
        uint64_t expect = 0;
        B200_CALL(b200_request_io_bytes(_req, 0, static_cast<int>(k), &expect), "io size");
        if (blob->byteSize() != expect)
            IE_THROW(ParameterMismatch) << "B200 plugin: input blob '" << _inputNames[k] << "' has " << blob->byteSize() << " bytes, the network expects " << expect;
        ins[k] = BlobData(blob);
    }
    for (size_t k = 0; k < _outputNames.size(); ++k) {
        const Blob::Ptr& blob = _outputs.at(_outputNames[k]);
        uint64_t expect = 0;
        B200_CALL(b200_request_io_bytes(_req, 1, static_cast<int>(k), &expect), "io size");
        if (blob->byteSize() != expect)
            IE_THROW(ParameterMismatch) << "B200 plugin: output blob '" << _outputNames[k] << "' has " << blob->byteSize() << " bytes, the network produces " << expect;
        outs[k] = const_cast<void*>(BlobData(blob));
    }
    B200_CALL(b200_request_start(_req, ins.data(), outs.data()), "start");
}

void InferRequest::waitDevice() { B200_CALL(b200_request_wait(_req, -1), "wait"); }

void InferRequest::InferImpl() {
    preprocessAndStart();
    waitDevice();
}

std::map<std::string, InferenceEngineProfileInfo> InferRequest::GetPerformanceCounts() const {
    std::map<std::string, InferenceEngineProfileInfo> perf;
    if (!_network->config().perfCount) return perf;   // same as the reference: empty unless PERF_COUNT=YES
    const int n = b200_network_num_exec_nodes(_network->handle());
    std::vector<float> usec(static_cast<size_t>(std::max(n, 1)), 0.f);
    B200_CALL(b200_request_perf_counts(_req, usec.data(), n), "perf counters");
    b200_exec_node_info info;
    for (int k = 0; k < n; ++k) {
        B200_CALL(b200_network_exec_node(_network->handle(), k, &info), "exec node info");
        InferenceEngineProfileInfo p;
        std::memset(&p, 0, sizeof(p));
        p.status = InferenceEngineProfileInfo::EXECUTED;
        p.realTime_uSec = p.cpu_uSec = static_cast<long long>(usec[k] + 0.5f);
        p.execution_index = static_cast<unsigned>(k);
        std::snprintf(p.exec_type, sizeof(p.exec_type), "%s", info.impl);
        std::snprintf(p.layer_type, sizeof(p.layer_type), "%s", info.type);
        perf[info.name] = p;
    }
    return perf;
}

// ---- asynchronous wrapper: 3-stage pipeline for a remote device ---------------------------------------------------------
AsyncInferRequest::AsyncInferRequest(const std::shared_ptr<InferRequest>& request, const ITaskExecutor::Ptr& taskExecutor,
                                     const ITaskExecutor::Ptr& waitExecutor, const ITaskExecutor::Ptr& callbackExecutor)
    : AsyncInferRequestThreadSafeDefault(request, taskExecutor, callbackExecutor), _request(request), _waitExecutor(waitExecutor) {
    _pipeline = {{taskExecutor, [this] { _request->preprocessAndStart(); }},   // enqueue only, returns immediately
                 {_waitExecutor, [this] { _request->waitDevice(); }}};         // D2H lands in the user's blobs: no third stage needed
}

AsyncInferRequest::~AsyncInferRequest() { StopAndWait(); }

}  // namespace B200Plugin

static const InferenceEngine::Version version = {{2, 1}, "b200-r1", "b200Plugin"};
IE_DEFINE_PLUGIN_CREATE_FUNCTION(B200Plugin::Plugin, version)
