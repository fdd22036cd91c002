// The B200 Inference-Engine device plugin: class set modelled on the plugin API contract
// (inference-engine/src/plugin_api/cpp_interfaces/interface/ie_iplugin_internal.hpp:86-302,
//  impl/ie_executable_network_thread_safe_default.hpp:25-67, interface/ie_iinfer_request_internal.hpp:28-243,
//  impl/ie_infer_async_request_thread_safe_default.hpp:40-420).  All compute goes through the C ABI of
// libb200kernels.so (include/b200.h); there is no CPU fallback.
#pragma once
#include <cpp_interfaces/impl/ie_executable_network_thread_safe_default.hpp>
#include <cpp_interfaces/impl/ie_infer_async_request_thread_safe_default.hpp>
#include <cpp_interfaces/interface/ie_iinfer_request_internal.hpp>
#include <cpp_interfaces/interface/ie_iplugin_internal.hpp>

#include <atomic>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "b200.h"

namespace B200Plugin {

struct Configuration {
    int deviceId = 0;
    bool perfCount = false;
    int precision = B200_PREC_BF16;
    int streams = 1;
    bool dryCompile = false;

    Configuration() = default;
    // merges `config` over `defaults`; unknown keys throw NotFound (behaviour tests: behavior/config.hpp)
    Configuration(const std::map<std::string, std::string>& config, const Configuration& defaults, bool throwOnUnsupported = true);
    InferenceEngine::Parameter Get(const std::string& name) const;
    static std::vector<std::string> SupportedKeys();
    std::string PrecisionName() const;
};

[[noreturn]] void ThrowFromStatus(int status, const std::string& what);

class Plugin : public InferenceEngine::IInferencePlugin {
public:
    Plugin();
    ~Plugin();
    void SetConfig(const std::map<std::string, std::string>& config) override;
    InferenceEngine::Parameter GetConfig(const std::string& name, const std::map<std::string, InferenceEngine::Parameter>& options) const override;
    InferenceEngine::Parameter GetMetric(const std::string& name, const std::map<std::string, InferenceEngine::Parameter>& options) const override;
    InferenceEngine::QueryNetworkResult QueryNetwork(const InferenceEngine::CNNNetwork& network,
                                                     const std::map<std::string, std::string>& config) const override;
    InferenceEngine::IExecutableNetworkInternal::Ptr LoadExeNetworkImpl(const InferenceEngine::CNNNetwork& network,
                                                                        const std::map<std::string, std::string>& config) override;
    void AddExtension(const std::shared_ptr<InferenceEngine::IExtension>& extension) override;

    InferenceEngine::ITaskExecutor::Ptr waitExecutor() const { return _waitExecutor; }

private:
    Configuration _cfg;
    InferenceEngine::ITaskExecutor::Ptr _waitExecutor;
};

class ExecutableNetwork : public InferenceEngine::ExecutableNetworkThreadSafeDefault {
public:
    ExecutableNetwork(const std::shared_ptr<const ngraph::Function>& function, const Configuration& cfg, const std::shared_ptr<Plugin>& plugin);
    ~ExecutableNetwork();

    InferenceEngine::IInferRequestInternal::Ptr CreateInferRequestImpl(InferenceEngine::InputsDataMap networkInputs,
                                                                       InferenceEngine::OutputsDataMap networkOutputs) override;
    InferenceEngine::IInferRequestInternal::Ptr CreateInferRequest() override;
    InferenceEngine::Parameter GetMetric(const std::string& name) const override;
    InferenceEngine::Parameter GetConfig(const std::string& name) const override;
    InferenceEngine::CNNNetwork GetExecGraphInfo() override;

    b200_network* handle() const { return _net; }
    const Configuration& config() const { return _cfg; }
    int inputIndex(const std::string& name) const;
    int outputIndex(const std::string& name) const;
    const std::string& name() const { return _name; }

private:
    b200_network* _net = nullptr;
    Configuration _cfg;
    std::shared_ptr<Plugin> _pluginRef;
    std::string _name;
    std::map<std::string, int> _inputIndex, _outputIndex;
};

class InferRequest : public InferenceEngine::IInferRequestInternal {
public:
    InferRequest(const InferenceEngine::InputsDataMap& networkInputs, const InferenceEngine::OutputsDataMap& networkOutputs,
                 const std::shared_ptr<ExecutableNetwork>& network);
    ~InferRequest();

    void InferImpl() override;
    std::map<std::string, InferenceEngine::InferenceEngineProfileInfo> GetPerformanceCounts() const override;

    // stages of the asynchronous pipeline (docs/template_plugin/src/template_async_infer_request.cpp:23-39)
    void preprocessAndStart();   // host pre-processing, H2D + CUDA graph + D2H enqueued on the request's stream (non-blocking)
    void waitDevice();           // blocks on the request's completion event

private:
    InferenceEngine::Blob::Ptr allocatePinned(const InferenceEngine::TensorDesc& desc);

    std::shared_ptr<ExecutableNetwork> _network;
    b200_request* _req = nullptr;
    std::vector<void*> _pinned;
    std::vector<std::string> _inputNames, _outputNames;   // in device I/O index order
};

class AsyncInferRequest : public InferenceEngine::AsyncInferRequestThreadSafeDefault {
public:
    AsyncInferRequest(const std::shared_ptr<InferRequest>& request, const InferenceEngine::ITaskExecutor::Ptr& taskExecutor,
                      const InferenceEngine::ITaskExecutor::Ptr& waitExecutor, const InferenceEngine::ITaskExecutor::Ptr& callbackExecutor);
    ~AsyncInferRequest();

private:
    std::shared_ptr<InferRequest> _request;
    InferenceEngine::ITaskExecutor::Ptr _waitExecutor;
};

}  // namespace B200Plugin
