// B200 device plugin: Plugin / ExecutableNetwork / InferRequest / AsyncInferRequest.  See b200_plugin.hpp.
#include "b200_plugin.hpp"

#include <iostream>

#include <blob_factory.hpp>
#include <exec_graph_info.hpp>
#include <ie_algorithm.hpp>
#include <ie_icore.hpp>
#include <ie_metric_helpers.hpp>
#include <ie_ngraph_utils.hpp>
#include <ie_plugin_config.hpp>
#include <threading/ie_executor_manager.hpp>

#include <ngraph/opsets/opset1.hpp>
#include <ngraph/variant.hpp>
#include <transformations/rt_info/fused_names_attribute.hpp>
#include <transformations/serialize.hpp>
#include <cstdlib>
#include <sstream>
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
            deviceId = value.empty() ? "0" : value;
        } else if (key == B200ConfigParams::KEY_B200_DEVICES) {
            deviceList = value;
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
    // resolve the device set: B200_DEVICES ("0,1,2,3") wins over DEVICE_ID ("3" | "ALL"); a dry compile touches no device
    int count = 0;
    b200_device_count(&count);
    devices.clear();
    auto add = [&](const std::string& text) {
        int id = -1;
        try {
            size_t used = 0;
            id = std::stoi(text, &used);
            if (used != text.size()) id = -1;
        } catch (...) {
            id = -1;
        }
        if (id < 0) IE_THROW(ParameterMismatch) << "B200 plugin: bad device id '" << text << "'";
        if (!dryCompile && count > 0 && id >= count) IE_THROW(NotFound) << "B200 plugin: DEVICE_ID " << id << " out of range (" << count << " devices)";
        devices.push_back(id);    // a device may be listed twice (two replicas on one GPU): legal, and how a 1-GPU box tests the scatter/gather
    };
    if (!deviceList.empty()) {
        std::stringstream ss(deviceList);
        std::string item;
        while (std::getline(ss, item, ',')) add(item);
    } else if (deviceId == "ALL") {
        for (int k = 0; k < std::max(count, 1); ++k) devices.push_back(k);
    } else {
        add(deviceId);
    }
    if (devices.empty()) IE_THROW(ParameterMismatch) << "B200 plugin: empty device list";
}

std::string Configuration::PrecisionName() const { return precision == B200_PREC_FP32 ? "FP32" : (precision == B200_PREC_BF16 ? "BF16" : "FP16"); }

std::vector<std::string> Configuration::SupportedKeys() {
    return {CONFIG_KEY(DEVICE_ID), CONFIG_KEY(PERF_COUNT), B200ConfigParams::KEY_B200_INFERENCE_PRECISION, B200ConfigParams::KEY_B200_THROUGHPUT_STREAMS,
            B200ConfigParams::KEY_B200_DEVICES};
}

Parameter Configuration::Get(const std::string& name) const {
    if (name == CONFIG_KEY(DEVICE_ID)) return {deviceId};
    if (name == B200ConfigParams::KEY_B200_DEVICES) {
        std::string list;
        for (size_t k = 0; k < devices.size(); ++k) list += (k ? "," : "") + std::to_string(devices[k]);
        return {list};
    }
    if (name == CONFIG_KEY(PERF_COUNT)) return {std::string(perfCount ? CONFIG_VALUE(YES) : CONFIG_VALUE(NO))};
    if (name == B200ConfigParams::KEY_B200_INFERENCE_PRECISION) return {PrecisionName()};
    if (name == B200ConfigParams::KEY_B200_THROUGHPUT_STREAMS) return {std::to_string(streams)};
    IE_THROW(NotFound) << "Unsupported property " << name << " by B200 plugin";
}

// ---- plugin -------------------------------------------------------------------------------------------------------
Plugin::Plugin() {
    _pluginName = "B200";
    // Process-wide default of B200_INFERENCE_PRECISION (a LoadNetwork / SetConfig value still wins).  The reference's CPU plugin has the
    // same need: its functional-test instances switch the implicit bf16 mode off for the whole run
    // (inference-engine/tests/functional/plugin/cpu/shared_tests_instances/core_config.cpp), because the behaviour suites compare
    // fp32 values for equality.  tools/run_conformance.sh sets it to FP32 for the same reason.
    if (const char* env = std::getenv("B200_INFERENCE_PRECISION")) _cfg = Configuration{{{B200ConfigParams::KEY_B200_INFERENCE_PRECISION, env}}, _cfg};
    // wait stages block a thread on a CUDA event: they get their own executor so that start stages are never starved
    // (RANGE_FOR_ASYNC_INFER_REQUESTS advertises up to 16 requests in flight: one blocked thread each at worst)
    _waitExecutor = ExecutorManager::getInstance()->getIdleCPUStreamsExecutor({"B200WaitExecutor", 16});
}

Plugin::~Plugin() {
    ExecutorManager::getInstance()->clear("B200StreamsExecutor");
    ExecutorManager::getInstance()->clear("B200WaitExecutor");
}

void Plugin::SetConfig(const std::map<std::string, std::string>& config) { _cfg = Configuration{config, _cfg}; }

// PERF_COUNT has two readers in the reference with two expectations: the behaviour suites read the PLUGIN's value as a bool
// (core_integration.cpp of the template instances; docs/template_plugin/src/template_config.cpp:44-45 returns one), the MULTI plugin reads
// the EXECUTABLE NETWORK's value as the string YES / NO (multi_device_plugin.cpp:219-223, as the CPU plugin returns it).  Each gets its own.
Parameter Plugin::GetConfig(const std::string& name, const std::map<std::string, Parameter>&) const {
    if (name == CONFIG_KEY(PERF_COUNT)) return {_cfg.perfCount};
    return _cfg.Get(name);
}

void Plugin::AddExtension(const std::shared_ptr<IExtension>&) { IE_THROW(NotImplemented) << "B200 plugin does not support extensions"; }

Parameter Plugin::GetMetric(const std::string& name, const std::map<std::string, Parameter>& options) const {
    if (name == METRIC_KEY(SUPPORTED_METRICS)) {
        IE_SET_METRIC_RETURN(SUPPORTED_METRICS, std::vector<std::string>{METRIC_KEY(AVAILABLE_DEVICES), METRIC_KEY(SUPPORTED_METRICS), METRIC_KEY(SUPPORTED_CONFIG_KEYS),
                                                                         METRIC_KEY(FULL_DEVICE_NAME), METRIC_KEY(DEVICE_ARCHITECTURE),
                                                                         METRIC_KEY(OPTIMIZATION_CAPABILITIES), METRIC_KEY(RANGE_FOR_ASYNC_INFER_REQUESTS),
                                                                         METRIC_KEY(RANGE_FOR_STREAMS), METRIC_KEY(IMPORT_EXPORT_SUPPORT)});
    } else if (name == METRIC_KEY(IMPORT_EXPORT_SUPPORT)) {
        IE_SET_METRIC_RETURN(IMPORT_EXPORT_SUPPORT, true);
    } else if (name == METRIC_KEY(SUPPORTED_CONFIG_KEYS)) {
        IE_SET_METRIC_RETURN(SUPPORTED_CONFIG_KEYS, Configuration::SupportedKeys());
    } else if (name == METRIC_KEY(AVAILABLE_DEVICES)) {
        int count = 0;
        b200_device_count(&count);
        std::vector<std::string> ids;
        for (int k = 0; k < count; ++k) ids.push_back(std::to_string(k));
        IE_SET_METRIC_RETURN(AVAILABLE_DEVICES, ids);
    } else if (name == METRIC_KEY(FULL_DEVICE_NAME)) {
        int dev = _cfg.devices.front();
        auto it = options.find(CONFIG_KEY(DEVICE_ID));
        if (it != options.end()) {
            try {
                dev = std::stoi(it->second.as<std::string>());
            } catch (...) {
                dev = _cfg.devices.front();   // "ALL"
            }
        }
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
    return std::make_shared<ExecutableNetwork>(function, network.getInputsInfo(), cfg, std::static_pointer_cast<Plugin>(shared_from_this()));
}

// ---- remote context (ie_remote_context.hpp:58-178; the Core forwards Core::CreateContext / GetDefaultContext / LoadNetwork(net, ctx)) -------
std::shared_ptr<InferenceEngine::RemoteContext> Plugin::CreateContext(const ParamMap& params) {
    int device = _cfg.devices.front();
    auto it = params.find(B200_PARAM_DEVICE_ID);
    if (it != params.end()) {
        if (it->second.is<int>()) device = it->second.as<int>();
        else device = std::stoi(it->second.as<std::string>());
    }
    int count = 0;
    b200_device_count(&count);
    if (device < 0 || device >= count) IE_THROW(NotFound) << "B200 plugin: no device " << device << " for a remote context (" << count << " devices)";
    return std::make_shared<RemoteContext>(device);
}

std::shared_ptr<InferenceEngine::RemoteContext> Plugin::GetDefaultContext(const ParamMap& params) {
    int device = _cfg.devices.front();
    auto it = params.find(B200_PARAM_DEVICE_ID);
    if (it != params.end()) device = it->second.is<int>() ? it->second.as<int>() : std::stoi(it->second.as<std::string>());
    std::lock_guard<std::mutex> lk(_ctxMutex);
    auto& ctx = _defaultContexts[device];
    if (!ctx) ctx = std::dynamic_pointer_cast<RemoteContext>(CreateContext({{B200_PARAM_DEVICE_ID, device}}));
    return ctx;
}

IExecutableNetworkInternal::Ptr Plugin::LoadExeNetworkImpl(const CNNNetwork& network, const std::shared_ptr<InferenceEngine::RemoteContext>& context,
                                                           const std::map<std::string, std::string>& config) {
    auto ctx = std::dynamic_pointer_cast<RemoteContext>(context);
    if (!ctx) IE_THROW(ParameterMismatch) << "B200 plugin: LoadNetwork was given a remote context of another device";
    auto cfg = config;
    cfg[CONFIG_KEY(DEVICE_ID)] = std::to_string(ctx->device());     // the network is compiled for the context's device
    return LoadExeNetworkImpl(network, cfg);
}

std::shared_ptr<InferenceEngine::RemoteContext> ExecutableNetwork::GetContext() const {
    return _pluginRef->GetDefaultContext({{B200_PARAM_DEVICE_ID, _cfg.devices.front()}});
}

// ---- Export / Import (docs/template_plugin/src/template_executable_network.cpp:179-200, template_plugin.cpp:108-112; the Core drives
// them for CACHE_DIR: inference-engine/src/inference_engine/ie_core.cpp:265-340).  Stream layout:
//   "B200NET2" | precision,streams (i32) | device id text, device list text | inputs: name, precision, layout, per-channel mean / scale,
//   mean image (elements per channel, 0 = none; then the planes) | outputs: name, precision |
//   IR v10 xml | IR weights | kernel-selection table (b200_network_get_tuning).
// The IR is the network as loaded; what the export saves on import is the measured part of LoadNetwork: the per-layer kernel selection.
namespace {
const char kMagic[8] = {'B', '2', '0', '0', 'N', 'E', 'T', '2'};

void putString(std::ostream& os, const std::string& v) {
    const uint64_t n = v.size();
    os.write(reinterpret_cast<const char*>(&n), sizeof(n));
    os.write(v.data(), static_cast<std::streamsize>(n));
}
std::string getString(std::istream& is) {
    uint64_t n = 0;
    is.read(reinterpret_cast<char*>(&n), sizeof(n));
    if (!is || n > (uint64_t(1) << 34)) IE_THROW(NetworkNotRead) << "B200 plugin: truncated or corrupt exported network";
    std::string v(static_cast<size_t>(n), '\0');
    is.read(&v[0], static_cast<std::streamsize>(n));
    if (!is) IE_THROW(NetworkNotRead) << "B200 plugin: truncated exported network";
    return v;
}
void putInt(std::ostream& os, int32_t v) { os.write(reinterpret_cast<const char*>(&v), sizeof(v)); }
int32_t getInt(std::istream& is) {
    int32_t v = 0;
    is.read(reinterpret_cast<char*>(&v), sizeof(v));
    if (!is) IE_THROW(NetworkNotRead) << "B200 plugin: truncated exported network";
    return v;
}
}  // namespace

void ExecutableNetwork::Export(std::ostream& model) {
    model.write(kMagic, sizeof(kMagic));
    putInt(model, _cfg.precision);
    putInt(model, _cfg.streams);
    putString(model, _cfg.deviceId);
    putString(model, _cfg.deviceList);
    putInt(model, static_cast<int32_t>(_networkInputs.size()));
    for (const auto& in : _networkInputs) {
        putString(model, in.first);
        putInt(model, static_cast<int32_t>(in.second->getPrecision()));
        putInt(model, static_cast<int32_t>(in.second->getLayout()));
        // PreProcessInfo mean values / scales (they are part of what was compiled: AddMeanScale)
        const PreProcessInfo& pre = in.second->getPreProcess();
        const int32_t channels = static_cast<int32_t>(pre.getNumberOfChannels());
        putInt(model, channels);
        for (int32_t c = 0; c < channels; ++c) {
            const float mv[2] = {pre[c]->meanValue, pre[c]->stdScale};
            model.write(reinterpret_cast<const char*>(mv), sizeof(mv));
        }
        // MEAN_IMAGE: the per-channel planes (AddMeanScale checked that every channel has an FP32 blob of the same size)
        const int32_t plane = channels && pre[0]->meanData ? static_cast<int32_t>(pre[0]->meanData->size()) : 0;
        putInt(model, plane);
        for (int32_t c = 0; plane && c < channels; ++c) {
            auto mem = as<MemoryBlob>(pre[c]->meanData);
            auto locked = mem->rmap();
            model.write(locked.as<const char*>(), static_cast<std::streamsize>(plane * sizeof(float)));
        }
    }
    putInt(model, static_cast<int32_t>(_networkOutputs.size()));
    for (const auto& out : _networkOutputs) {
        putString(model, out.first);
        putInt(model, static_cast<int32_t>(out.second->getPrecision()));
    }
    std::map<std::string, ngraph::OpSet> custom_opsets;
    std::stringstream xml, bin;
    ngraph::pass::Serialize serializer(xml, bin, ngraph::pass::Serialize::Version::IR_V10, custom_opsets);
    serializer.run_on_function(_function);
    putString(model, xml.str());
    putString(model, bin.str());
    size_t need = 0;
    std::string tuning;
    if (!_cfg.dryCompile && b200_network_get_tuning(_nets[0], nullptr, 0, &need) == B200_OK && need > 1) {
        tuning.resize(need);
        b200_network_get_tuning(_nets[0], &tuning[0], need, &need);
        tuning.resize(need - 1);
    }
    putString(model, tuning);
}

IExecutableNetworkInternal::Ptr Plugin::ImportNetwork(std::istream& model, const std::map<std::string, std::string>& config) {
    char magic[sizeof(kMagic)] = {};
    model.read(magic, sizeof(magic));
    if (!model || std::memcmp(magic, kMagic, sizeof(kMagic)) != 0) IE_THROW(NetworkNotRead) << "B200 plugin: not an exported B200 network";
    Configuration stored = _cfg;
    stored.precision = getInt(model);
    stored.streams = getInt(model);
    std::map<std::string, std::string> storedDevices;
    storedDevices[CONFIG_KEY(DEVICE_ID)] = getString(model);
    const std::string list = getString(model);
    if (!list.empty()) storedDevices[B200ConfigParams::KEY_B200_DEVICES] = list;
    std::vector<std::tuple<std::string, int, int>> ins;
    std::vector<std::vector<float>> inMeanScale, inMeanImage;
    std::vector<std::pair<std::string, int>> outs;
    for (int k = 0, n = getInt(model); k < n; ++k) {
        std::string name = getString(model);
        const int prec = getInt(model), layout = getInt(model);
        ins.emplace_back(name, prec, layout);
        const int channels = getInt(model);
        if (channels < 0 || channels > 65536) IE_THROW(NetworkNotRead) << "B200 plugin: corrupt exported network";
        std::vector<float> ms(static_cast<size_t>(2 * channels));
        if (channels) model.read(reinterpret_cast<char*>(ms.data()), static_cast<std::streamsize>(ms.size() * sizeof(float)));
        if (!model) IE_THROW(NetworkNotRead) << "B200 plugin: truncated exported network";
        inMeanScale.push_back(std::move(ms));
        const int plane = getInt(model);
        if (plane < 0 || plane > (1 << 28)) IE_THROW(NetworkNotRead) << "B200 plugin: corrupt exported network";
        std::vector<float> image(static_cast<size_t>(plane) * static_cast<size_t>(channels));
        if (!image.empty()) model.read(reinterpret_cast<char*>(image.data()), static_cast<std::streamsize>(image.size() * sizeof(float)));
        if (!model) IE_THROW(NetworkNotRead) << "B200 plugin: truncated exported network";
        inMeanImage.push_back(std::move(image));
    }
    for (int k = 0, n = getInt(model); k < n; ++k) {
        std::string name = getString(model);
        outs.emplace_back(name, getInt(model));
    }
    const std::string xml = getString(model);
    const std::string bin = getString(model);
    const std::string tuning = getString(model);
    // what the caller passes now wins over what was stored (the Core hashes the config into the cache key, so for CACHE_DIR they agree)
    std::map<std::string, std::string> merged = storedDevices;
    for (const auto& kv : config) merged[kv.first] = kv.second;
    Configuration cfg{merged, stored};
    auto core = GetCore();
    if (!core) IE_THROW(GeneralError) << "B200 plugin: ImportNetwork needs the Core (IR reader)";
    auto weights = make_shared_blob<uint8_t>(TensorDesc(Precision::U8, {bin.size()}, Layout::C));
    weights->allocate();
    if (!bin.empty()) std::memcpy(weights->buffer().as<uint8_t*>(), bin.data(), bin.size());
    CNNNetwork network = core->ReadNetwork(xml, weights);
    for (const auto& in : ins) {
        auto it = network.getInputsInfo().find(std::get<0>(in));
        if (it == network.getInputsInfo().end()) IE_THROW(NetworkNotRead) << "B200 plugin: exported network names an unknown input '" << std::get<0>(in) << "'";
        it->second->setPrecision(Precision(static_cast<Precision::ePrecision>(std::get<1>(in))));
        it->second->setLayout(static_cast<Layout>(std::get<2>(in)));
        const auto& ms = inMeanScale[static_cast<size_t>(&in - &ins[0])];
        if (!ms.empty()) {
            PreProcessInfo& pre = it->second->getPreProcess();
            pre.init(ms.size() / 2);
            bool any = false;
            for (size_t c = 0; c < ms.size() / 2; ++c) {
                pre[c]->meanValue = ms[2 * c];
                pre[c]->stdScale = ms[2 * c + 1];
                any = any || ms[2 * c] != 0.f || ms[2 * c + 1] != 1.f;
            }
            if (any) pre.setVariant(MEAN_VALUE);
            const auto& image = inMeanImage[static_cast<size_t>(&in - &ins[0])];
            if (!image.empty()) {
                const size_t channels = ms.size() / 2, plane = image.size() / channels;
                for (size_t c = 0; c < channels; ++c) {
                    auto blob = make_shared_blob<float>(TensorDesc(Precision::FP32, {plane}, Layout::C));
                    blob->allocate();
                    std::memcpy(blob->buffer().as<float*>(), &image[c * plane], plane * sizeof(float));
                    pre[c]->meanData = blob;
                }
                pre.setVariant(MEAN_IMAGE);
            }
        }
    }
    for (const auto& out : outs) {
        auto it = network.getOutputsInfo().find(out.first);
        if (it == network.getOutputsInfo().end()) IE_THROW(NetworkNotRead) << "B200 plugin: exported network names an unknown output '" << out.first << "'";
        it->second->setPrecision(Precision(static_cast<Precision::ePrecision>(out.second)));
    }
    auto exe = std::make_shared<ExecutableNetwork>(network.getFunction(), network.getInputsInfo(), cfg, std::static_pointer_cast<Plugin>(shared_from_this()), tuning);
    exe->setNetworkInputs(network.getInputsInfo());
    exe->setNetworkOutputs(network.getOutputsInfo());
    exe->SetPointerToPlugin(shared_from_this());
    return exe;
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
        // op-level check, then the node's own geometry through the compiler (cheap ops and weight-heavy convolutions are trusted to the
        // op-level check: their lowering has no shape-dependent refusal)
        std::string why = "no lowering for this operation";
        bool ok = IsNodeSupported(node);
        if (ok && !ngraph::is_type<ngraph::opset1::Convolution>(node) && !ngraph::is_type<ngraph::opset1::GroupConvolution>(node)) ok = DryCompileNode(node, &why);
        if (!ok && std::getenv("B200_QUERY_VERBOSE"))   // diagnostics for HETERO splits: which node was refused and why
            std::cerr << "[B200 QueryNetwork] " << node->get_type_name() << " '" << node->get_friendly_name() << "' refused: " << why << std::endl;
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
ExecutableNetwork::ExecutableNetwork(const std::shared_ptr<const ngraph::Function>& function, const InputsDataMap& inputs, const Configuration& cfg,
                                     const std::shared_ptr<Plugin>& plugin, const std::string& tuning)
    : ExecutableNetworkThreadSafeDefault(nullptr, nullptr), _cfg(cfg), _pluginRef(plugin), _name(function->get_friendly_name()) {
    _function = ngraph::clone_function(*function);
    const int shards = static_cast<int>(cfg.devices.size());
    // batch sharding: the replica every device runs is the same function at batch N / G (ShardBatch re-validates the clone)
    auto replica = shards > 1 ? ShardBatch(function, static_cast<size_t>(shards)) : ngraph::clone_function(*function);
    AddMeanScale(replica, inputs);      // PreProcessInfo mean / scale run on the device (Export keeps the network as loaded)
    auto transformed = TransformNetwork(replica);
    b200_graph* graph = LowerFunction(transformed);
    std::string table = tuning;
    try {
        for (int g = 0; g < shards; ++g) {
            b200_compile_options opts;
            std::memset(&opts, 0, sizeof(opts));
            opts.abi_version = B200_ABI_VERSION;
            opts.device = cfg.dryCompile ? -1 : cfg.devices[g];
            opts.precision = cfg.precision;
            b200_network* net = nullptr;
            int rc = b200_compile(graph, &opts, &net);
            if (rc < 0) ThrowFromStatus(rc, "cannot compile network '" + _name + "'" + (shards > 1 ? " for device " + std::to_string(cfg.devices[g]) : ""));
            _nets.push_back(net);
            if (cfg.dryCompile) continue;
            // kernel selection is measured once (first device, or imported with the network) and installed on the other replicas:
            // the devices of one box are identical parts
            if (!table.empty()) B200_CALL(b200_network_set_tuning(net, table.c_str()), "kernel-selection table");
            B200_CALL(b200_network_tune(net), "kernel selection");
            if (table.empty()) {
                size_t need = 0;
                if (b200_network_get_tuning(net, nullptr, 0, &need) == B200_OK && need > 1) {
                    table.resize(need);
                    b200_network_get_tuning(net, &table[0], need, &need);
                    table.resize(need - 1);
                }
            }
        }
    } catch (...) {
        b200_graph_destroy(graph);
        for (auto* n : _nets) b200_network_destroy(n);
        _nets.clear();
        throw;
    }
    b200_graph_destroy(graph);
    b200_tensor_info info;
    for (int k = 0; k < b200_network_num_inputs(_nets[0]); ++k) {
        B200_CALL(b200_network_io_info(_nets[0], 0, k, &info), "io info");
        _inputIndex[info.name] = k;
    }
    for (int k = 0; k < b200_network_num_outputs(_nets[0]); ++k) {
        B200_CALL(b200_network_io_info(_nets[0], 1, k, &info), "io info");
        _outputIndex[info.name] = k;
    }
    // scatter / gather plan: compare every parameter / result of the network as loaded with the replica's
    _inPlan.assign(_inputIndex.size(), IoPlan());
    _outPlan.assign(_outputIndex.size(), IoPlan());
    if (shards > 1) {
        const auto& fullParams = _function->get_parameters();
        const auto& partParams = replica->get_parameters();
        for (size_t k = 0; k < fullParams.size(); ++k) {
            const auto fs = fullParams[k]->get_shape(), ps = partParams[k]->get_shape();
            auto it = _inputIndex.find(fullParams[k]->get_friendly_name());
            if (it == _inputIndex.end()) continue;
            IoPlan& plan = _inPlan[it->second];
            if (fs == ps) plan.kind = IoPlan::Replicated;
            else if (!fs.empty() && fs[0] == ps[0] * shards && std::equal(fs.begin() + 1, fs.end(), ps.begin() + 1)) plan.kind = IoPlan::Sliced;
            else IE_THROW(NotImplemented) << "B200 plugin: input '" << it->first << "' cannot be sharded over " << shards << " devices";
        }
        const size_t batch = fullParams.empty() || fullParams[0]->get_shape().empty() ? 0 : fullParams[0]->get_shape()[0];
        const auto& fullRes = _function->get_results();
        const auto& partRes = replica->get_results();
        for (size_t k = 0; k < fullRes.size(); ++k) {
            const auto fs = fullRes[k]->get_shape(), ps = partRes[k]->get_shape();
            auto src = fullRes[k]->get_input_source_output(0);
            std::string oname = src.get_node()->get_friendly_name();
            if (src.get_node()->get_output_size() > 1) oname += "." + std::to_string(src.get_index());
            auto it = _outputIndex.find(oname);
            if (it == _outputIndex.end()) continue;
            IoPlan& plan = _outPlan[it->second];
            if (fs == ps) plan.kind = IoPlan::Replicated;
            else if (!fs.empty() && fs[0] == ps[0] * shards && std::equal(fs.begin() + 1, fs.end(), ps.begin() + 1)) plan.kind = IoPlan::Sliced;
            else if (fs.size() == 4 && fs[0] == 1 && fs[1] == 1 && fs[3] == 7 && ps[3] == 7 && fs[2] == ps[2] * shards) {
                plan.kind = IoPlan::Detections;      // DetectionOutput [1,1,N*keep_top_k,7]: rows re-packed on the host after the wait
                plan.images_per_shard = static_cast<int>(batch / shards);
            } else {
                IE_THROW(NotImplemented) << "B200 plugin: output '" << it->first << "' cannot be gathered from " << shards << " devices";
            }
        }
    }
    // start stages are cheap host work (enqueue); one executor thread per stream keeps submission order per stream
    auto streamsConfig = IStreamsExecutor::Config{"B200StreamsExecutor", std::max(1, cfg.streams)};
    _taskExecutor = ExecutorManager::getInstance()->getIdleCPUStreamsExecutor(streamsConfig);
    _callbackExecutor = _taskExecutor;
}

ExecutableNetwork::~ExecutableNetwork() {
    for (auto* n : _nets) b200_network_destroy(n);
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
        // two requests per stream (the H2D copy of request k+1 overlaps the kernels of request k) plus one that hides
        // the host-side turn-around between a request's completion and its next StartAsync (measured: 2 -> 74.4k, 3 -> 76.6k img/s)
        IE_SET_METRIC_RETURN(OPTIMAL_NUMBER_OF_INFER_REQUESTS, static_cast<unsigned int>(2 * std::max(1, _cfg.streams) + 1));
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
    const int n = b200_network_num_exec_nodes(_nets[0]);
    b200_exec_node_info info;
    for (int k = 0; k < n; ++k) {
        B200_CALL(b200_network_exec_node(_nets[0], k, &info), "exec node info");
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
    if (auto remote = std::dynamic_pointer_cast<RemoteBlob>(blob)) {
        // a blob that lives in HBM (b200_remote.hpp): the C ABI takes the device address, the "copy" is device to device
        if (!remote->devicePtr()) IE_THROW(NotAllocated) << "B200 plugin: remote blob has no device memory";
        return remote->devicePtr();
    }
    auto mem = as<MemoryBlob>(blob);
    if (!mem) IE_THROW(NotImplemented) << "B200 plugin: only MemoryBlob inputs/outputs are supported";
    return mem->rmap().as<const uint8_t*>();   // includes the ROI / padding offset of the descriptor
}

// Default blobs own pinned host memory (cudaHostAlloc, portable across the devices of the box) through the IAllocator interface
// (inference-engine/include/ie_allocator.hpp:26-40): a Blob::Ptr obtained with GetBlob stays valid after the request is gone, and
// the H2D / D2H copies of start() are asynchronous and run at PCIe speed.
class PinnedAllocator : public IAllocator {
public:
    void* lock(void* handle, LockOp = LOCK_FOR_WRITE) noexcept override { return handle; }
    void unlock(void*) noexcept override {}
    void* alloc(size_t size) noexcept override {
        void* ptr = nullptr;
        return b200_host_alloc(&ptr, std::max<size_t>(size, 16)) == B200_OK ? ptr : nullptr;
    }
    bool free(void* handle) noexcept override { return b200_host_free(handle) == B200_OK; }
};

}  // namespace

InferRequest::InferRequest(const InputsDataMap& networkInputs, const OutputsDataMap& networkOutputs, const std::shared_ptr<ExecutableNetwork>& network)
    : IInferRequestInternal(networkInputs, networkOutputs), _network(network), _allocator(std::make_shared<PinnedAllocator>()) {
    const int shards = network->numShards();
    for (int g = 0; g < shards; ++g) {
        b200_request* r = nullptr;
        B200_CALL(b200_request_create(network->handle(g), &r), "cannot create infer request");
        _reqs.push_back(r);
    }
    _inputNames.resize(_networkInputs.size());
    _outputNames.resize(_networkOutputs.size());
    _inShardBytes.assign(_networkInputs.size(), 0);
    _outShardBytes.assign(_networkOutputs.size(), 0);
    _inFormat.assign(_networkInputs.size(), {0, 0});
    _outFormat.assign(_networkOutputs.size(), {0, 0});
    _outStage.assign(_networkOutputs.size(), std::vector<void*>(static_cast<size_t>(shards), nullptr));
    for (const auto& in : _networkInputs) {
        const int idx = network->inputIndex(in.first);
        const TensorDesc& desc = in.second->getTensorDesc();
        _inFormat[idx] = {BlobDtype(desc.getPrecision(), in.first), BlobLayout(desc, in.first)};
        for (auto* r : _reqs) B200_CALL(b200_request_set_io_format(r, 0, idx, _inFormat[idx].first, _inFormat[idx].second), "input format");
        B200_CALL(b200_request_io_bytes(_reqs[0], 0, idx, &_inShardBytes[idx]), "io size");
        _inputNames[idx] = in.first;
        _inputs[in.first] = allocatePinned(desc);
    }
    for (const auto& out : _networkOutputs) {
        const int idx = network->outputIndex(out.first);
        const TensorDesc& desc = out.second->getTensorDesc();
        _outFormat[idx] = {BlobDtype(desc.getPrecision(), out.first), BlobLayout(desc, out.first)};
        for (auto* r : _reqs) B200_CALL(b200_request_set_io_format(r, 1, idx, _outFormat[idx].first, _outFormat[idx].second), "output format");
        B200_CALL(b200_request_io_bytes(_reqs[0], 1, idx, &_outShardBytes[idx]), "io size");
        _outputNames[idx] = out.first;
        _outputs[out.first] = allocatePinned(desc);
        // outputs that are not plain dim-0 slices land in per-shard staging first
        const IoPlan& plan = network->outputPlan(idx);
        for (int g = 0; g < shards; ++g) {
            if (plan.kind == IoPlan::Sliced || (plan.kind == IoPlan::Replicated && g == 0)) continue;
            void* ptr = nullptr;
            B200_CALL(b200_host_alloc(&ptr, std::max<size_t>(_outShardBytes[idx], 16)), "pinned staging");
            _stageAllocs.push_back(ptr);
            _outStage[idx][g] = ptr;
        }
    }
}

InferRequest::~InferRequest() {
    for (auto* r : _reqs) b200_request_destroy(r);   // drains the stream first
    for (void* p : _stageAllocs) b200_host_free(p);
    for (const auto& kv : _registered) b200_host_unregister(const_cast<void*>(kv.first));
}

Blob::Ptr InferRequest::allocatePinned(const TensorDesc& desc) {
    Blob::Ptr blob = make_blob_with_precision(desc, _allocator);
    blob->allocate();
    if (!blob->buffer().as<void*>()) IE_THROW(GeneralError) << "B200 plugin: pinned allocation failed: " << b200_last_error();
    return blob;
}

// Memory the user brought through SetBlob is page-locked on first use (cudaHostRegister) so that its copies are asynchronous too;
// a failure (already registered, foreign mapping) only costs the synchronous fallback of cudaMemcpyAsync.
void InferRequest::pinUserMemory(const void* ptr, size_t bytes) {
    if (!ptr || bytes < (64u << 10)) return;
    auto it = _registered.find(ptr);
    if (it != _registered.end() && it->second >= bytes) return;
    if (it != _registered.end()) {
        b200_host_unregister(const_cast<void*>(ptr));
        _registered.erase(it);
    }
    if (_registered.size() >= 32) {   // a caller that cycles through many buffers: forget the oldest registrations
        for (const auto& kv : _registered) b200_host_unregister(const_cast<void*>(kv.first));
        _registered.clear();
    }
    if (b200_host_register(const_cast<void*>(ptr), bytes) == B200_OK) _registered[ptr] = bytes;
}

void InferRequest::SetBlob(const std::string& name, const Blob::Ptr& data) {
    // size / precision checks, pre-processing set-up (resize / colour conversion on the host).  The blob's own layout is honoured at
    // the next start (syncFormat): a 4-D blob may be NCHW or NHWC whatever the network was loaded with, as in the reference's requests
    // (template_infer_request.cpp blobCopy, mkldnn_infer_request.cpp:212-240 pushInput with the blob's descriptor)
    IInferRequestInternal::SetBlob(name, data);
}

// The device-side edge kernels (ingest / emit) are recorded for one (precision, layout) per input / output; a blob that arrives with
// another layout re-records them (b200_request_set_io_format drops the captured graph of an idle request).
void InferRequest::syncFormat(bool output, size_t index, const Blob::Ptr& blob) {
    const std::string& name = output ? _outputNames[index] : _inputNames[index];
    const TensorDesc& desc = blob->getTensorDesc();
    auto& have = output ? _outFormat[index] : _inFormat[index];
    // a 4-D descriptor that names neither NCHW nor NHWC (Layout::ANY, a BLOCKED descriptor with default strides) keeps the layout the
    // request is recorded for, as before; only an explicit NCHW <-> NHWC difference re-records
    const bool named = desc.getDims().size() != 4 || desc.getLayout() == Layout::NCHW || desc.getLayout() == Layout::NHWC;
    const std::pair<int, int> want{BlobDtype(desc.getPrecision(), name), named ? BlobLayout(desc, name) : have.second};
    if (want == have || (want.first == have.first && want.second == B200_LAYOUT_ANY)) return;
    for (auto* r : _reqs) B200_CALL(b200_request_set_io_format(r, output ? 1 : 0, static_cast<int>(index), want.first, want.second), "blob format");
    B200_CALL(b200_request_io_bytes(_reqs[0], output ? 1 : 0, static_cast<int>(index), output ? &_outShardBytes[index] : &_inShardBytes[index]), "io size");
    have = want;
}

void InferRequest::preprocessAndStart() {
    execDataPreprocessing(_inputs);     // resize / colour conversion requested through PreProcessInfo (host, reused from the reference)
    const int shards = static_cast<int>(_reqs.size());
    std::vector<const uint8_t*> inBase(_inputNames.size(), nullptr);
    std::vector<uint8_t*> outBase(_outputNames.size(), nullptr);
    for (size_t k = 0; k < _inputNames.size(); ++k) {
        const Blob::Ptr& blob = _inputs.at(_inputNames[k]);
        if (!blob->is<InferenceEngine::RemoteBlob>()) syncFormat(false, k, blob);
        const uint64_t expect = _inShardBytes[k] * (_network->inputPlan(static_cast<int>(k)).kind == IoPlan::Sliced ? shards : 1);
        if (blob->byteSize() != expect)
            IE_THROW(ParameterMismatch) << "B200 plugin: input blob '" << _inputNames[k] << "' has " << blob->byteSize() << " bytes, the network expects " << expect;
        inBase[k] = static_cast<const uint8_t*>(BlobData(blob));
        pinUserMemory(inBase[k], blob->byteSize());
    }
    _mergeDst.assign(_outputNames.size(), nullptr);
    for (size_t k = 0; k < _outputNames.size(); ++k) {
        const Blob::Ptr& blob = _outputs.at(_outputNames[k]);
        const IoPlan& plan = _network->outputPlan(static_cast<int>(k));
        if (!blob->is<InferenceEngine::RemoteBlob>()) syncFormat(true, k, blob);
        const uint64_t expect = _outShardBytes[k] * (plan.kind == IoPlan::Replicated ? 1 : shards);
        if (blob->byteSize() != expect)
            IE_THROW(ParameterMismatch) << "B200 plugin: output blob '" << _outputNames[k] << "' has " << blob->byteSize() << " bytes, the network produces " << expect;
        if (plan.kind == IoPlan::Detections && blob->is<InferenceEngine::RemoteBlob>())
            IE_THROW(NotImplemented) << "B200 plugin: a sharded DetectionOutput is re-packed on the host and cannot land in a remote blob";
        outBase[k] = const_cast<uint8_t*>(static_cast<const uint8_t*>(BlobData(blob)));
        pinUserMemory(outBase[k], blob->byteSize());
        _mergeDst[k] = outBase[k];
    }
    // scatter + launch on every device before anyone waits: the devices run concurrently
    std::vector<const void*> ins(_inputNames.size(), nullptr);
    std::vector<void*> outs(_outputNames.size(), nullptr);
    for (int g = 0; g < shards; ++g) {
        for (size_t k = 0; k < ins.size(); ++k)
            ins[k] = inBase[k] + (_network->inputPlan(static_cast<int>(k)).kind == IoPlan::Sliced ? _inShardBytes[k] * g : 0);
        for (size_t k = 0; k < outs.size(); ++k) {
            const IoPlan& plan = _network->outputPlan(static_cast<int>(k));
            if (plan.kind == IoPlan::Sliced) outs[k] = outBase[k] + _outShardBytes[k] * g;
            else if (plan.kind == IoPlan::Replicated && g == 0) outs[k] = outBase[k];
            else outs[k] = _outStage[k][g];
        }
        B200_CALL(b200_request_start(_reqs[g], ins.data(), outs.data()), "start");
    }
}

void InferRequest::waitDevice() {
    int failed = 0;
    for (auto* r : _reqs) {     // every device is drained even if one reports an error: the blobs must not be in use afterwards
        const int rc = b200_request_wait(r, -1);
        if (rc < 0 && !failed) failed = rc;
    }
    if (failed) ThrowFromStatus(failed, "wait");
    const int shards = static_cast<int>(_reqs.size());
    if (shards == 1) return;
    // gather step of DetectionOutput: every shard produced [1,1,R,7] rows [image_id, label, conf, box] for ITS images, terminated by a
    // row with image_id -1 (mkldnn_detection_output_node.cpp:505-536); the unsharded tensor has the valid rows of all shards in image
    // order, ids renumbered, one terminator, zeros after it
    for (size_t k = 0; k < _outputNames.size(); ++k) {
        const IoPlan& plan = _network->outputPlan(static_cast<int>(k));
        if (plan.kind != IoPlan::Detections || !_mergeDst[k]) continue;
        const Blob::Ptr& blob = _outputs.at(_outputNames[k]);
        if (blob->getTensorDesc().getPrecision() != Precision::FP32) IE_THROW(NotImplemented) << "B200 plugin: sharded DetectionOutput needs an FP32 output blob";
        float* dst = reinterpret_cast<float*>(_mergeDst[k]);
        const size_t rowsPerShard = static_cast<size_t>(_outShardBytes[k] / (7 * sizeof(float)));
        const size_t total = rowsPerShard * shards;
        size_t w = 0;
        for (int g = 0; g < shards; ++g) {
            const float* src = static_cast<const float*>(_outStage[k][g]);
            for (size_t r = 0; r < rowsPerShard && src[r * 7] >= 0.f; ++r, ++w) {
                std::memcpy(dst + w * 7, src + r * 7, 7 * sizeof(float));
                dst[w * 7] += static_cast<float>(g * plan.images_per_shard);
            }
        }
        if (w < total) {
            std::memset(dst + w * 7, 0, (total - w) * 7 * sizeof(float));
            dst[w * 7] = -1.f;
        }
    }
}

void InferRequest::InferImpl() {
    preprocessAndStart();
    waitDevice();
}

std::map<std::string, InferenceEngineProfileInfo> InferRequest::GetPerformanceCounts() const {
    std::map<std::string, InferenceEngineProfileInfo> perf;
    if (!_network->config().perfCount) return perf;   // same as the reference: empty unless PERF_COUNT=YES
    const int n = b200_network_num_exec_nodes(_network->handle());
    std::vector<float> usec(static_cast<size_t>(std::max(n, 1)), 0.f);
    B200_CALL(b200_request_perf_counts(_reqs[0], usec.data(), n), "perf counters");
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
    // Cancel() (ie_infer_async_request_thread_safe_default.hpp:262-267) marks a busy request; the reference checks the mark between
    // nodes (mkldnn_graph.cpp:814-816, mkldnn_infer_request.cpp:519-523).  Here a request is one CUDA graph: the mark is honoured before
    // the enqueue, and after the device has drained (the buffers must not be in flight when Wait() reports INFER_CANCELLED).
    _pipeline = {{taskExecutor, [this] { ThrowIfCanceled(); _request->preprocessAndStart(); }},   // enqueue only, returns immediately
                 {_waitExecutor, [this] { _request->waitDevice(); ThrowIfCanceled(); }}};        // D2H lands in the user's blobs
}

AsyncInferRequest::~AsyncInferRequest() { StopAndWait(); }

}  // namespace B200Plugin

static const InferenceEngine::Version version = {{2, 1}, "b200-r1", "b200Plugin"};
IE_DEFINE_PLUGIN_CREATE_FUNCTION(B200Plugin::Plugin, version)
