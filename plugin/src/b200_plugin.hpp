// The B200 Inference-Engine device plugin: class set modelled on the plugin API contract
// (inference-engine/src/plugin_api/cpp_interfaces/interface/ie_iplugin_internal.hpp:86-302,
//  impl/ie_executable_network_thread_safe_default.hpp:25-67, interface/ie_iinfer_request_internal.hpp:28-243,
//  impl/ie_infer_async_request_thread_safe_default.hpp:40-420).  All compute goes through the C ABI of
// libb200kernels.so (include/b200.h); there is no CPU fallback.
//
// One ExecutableNetwork may span several GPUs of the box (DEVICE_ID=ALL or B200_DEVICES=0,1,..): the network is compiled once per
// device for batch N/G and every infer request of batch N is scattered as contiguous dim-0 slices of the user's (pinned) blobs,
// launched on all devices before the first wait, and gathered back - the batch-sharding counterpart of the reference's MULTI
// scheduler (inference-engine/src/multi_device/multi_device_exec_network.cpp:49-136, multi_device_async_infer_request.cpp:35-77),
// which hands WHOLE requests to idle devices instead.  Batch rows are independent on this path, so there is no collective.
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
#include "b200_remote.hpp"

namespace B200Plugin {

struct Configuration {
    std::vector<int> devices{0};   // CUDA ordinals the network is replicated on (one entry unless batch sharding was asked for)
    std::string deviceId = "0";    // DEVICE_ID as given ("3", "ALL")
    std::string deviceList;        // B200_DEVICES as given ("0,1,2,3"); wins over DEVICE_ID
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
    InferenceEngine::IExecutableNetworkInternal::Ptr LoadExeNetworkImpl(const InferenceEngine::CNNNetwork& network,
                                                                        const std::shared_ptr<InferenceEngine::RemoteContext>& context,
                                                                        const std::map<std::string, std::string>& config) override;
    InferenceEngine::IExecutableNetworkInternal::Ptr ImportNetwork(std::istream& model, const std::map<std::string, std::string>& config) override;
    std::shared_ptr<InferenceEngine::RemoteContext> CreateContext(const InferenceEngine::ParamMap& params) override;
    std::shared_ptr<InferenceEngine::RemoteContext> GetDefaultContext(const InferenceEngine::ParamMap& params) override;
    void AddExtension(const std::shared_ptr<InferenceEngine::IExtension>& extension) override;

    InferenceEngine::ITaskExecutor::Ptr waitExecutor() const { return _waitExecutor; }

private:
    Configuration _cfg;
    InferenceEngine::ITaskExecutor::Ptr _waitExecutor;
    std::mutex _ctxMutex;
    std::map<int, std::shared_ptr<RemoteContext>> _defaultContexts;   // one per device
};

// How one network input / output maps onto the per-device replicas
struct IoPlan {
    enum Kind { Sliced, Replicated, Detections } kind = Sliced;   // dim-0 slices | same tensor everywhere | DetectionOutput [1,1,R,7] re-pack
    int images_per_shard = 0;                                     // Detections: image ids of shard g are offset by g * images_per_shard
};

class ExecutableNetwork : public InferenceEngine::ExecutableNetworkThreadSafeDefault {
public:
    ExecutableNetwork(const std::shared_ptr<const ngraph::Function>& function, const InferenceEngine::InputsDataMap& inputs, const Configuration& cfg,
                      const std::shared_ptr<Plugin>& plugin, const std::string& tuning = std::string());
    ~ExecutableNetwork();

    InferenceEngine::IInferRequestInternal::Ptr CreateInferRequestImpl(InferenceEngine::InputsDataMap networkInputs,
                                                                       InferenceEngine::OutputsDataMap networkOutputs) override;
    InferenceEngine::IInferRequestInternal::Ptr CreateInferRequest() override;
    InferenceEngine::Parameter GetMetric(const std::string& name) const override;
    InferenceEngine::Parameter GetConfig(const std::string& name) const override;
    InferenceEngine::CNNNetwork GetExecGraphInfo() override;
    void Export(std::ostream& model) override;
    std::shared_ptr<InferenceEngine::RemoteContext> GetContext() const override;

    int numShards() const { return static_cast<int>(_nets.size()); }
    b200_network* handle(int shard = 0) const { return _nets[shard]; }
    const Configuration& config() const { return _cfg; }
    int inputIndex(const std::string& name) const;
    int outputIndex(const std::string& name) const;
    const IoPlan& inputPlan(int index) const { return _inPlan[index]; }
    const IoPlan& outputPlan(int index) const { return _outPlan[index]; }
    const std::string& name() const { return _name; }

private:
    std::vector<b200_network*> _nets;     // one compiled replica per device of _cfg.devices
    Configuration _cfg;
    std::shared_ptr<Plugin> _pluginRef;
    std::string _name;
    std::shared_ptr<ngraph::Function> _function;    // clone of the network as loaded (what Export serialises)
    std::map<std::string, int> _inputIndex, _outputIndex;
    std::vector<IoPlan> _inPlan, _outPlan;
};

class InferRequest : public InferenceEngine::IInferRequestInternal {
public:
    InferRequest(const InferenceEngine::InputsDataMap& networkInputs, const InferenceEngine::OutputsDataMap& networkOutputs,
                 const std::shared_ptr<ExecutableNetwork>& network);
    ~InferRequest();

    void InferImpl() override;
    void SetBlob(const std::string& name, const InferenceEngine::Blob::Ptr& data) override;
    std::map<std::string, InferenceEngine::InferenceEngineProfileInfo> GetPerformanceCounts() const override;

    // stages of the asynchronous pipeline (docs/template_plugin/src/template_async_infer_request.cpp:23-39)
    void preprocessAndStart();   // host pre-processing, then per device: H2D + CUDA graph + D2H enqueued on the request's stream (non-blocking)
    void waitDevice();           // blocks on every device's completion event, then re-packs DetectionOutput shards

private:
    InferenceEngine::Blob::Ptr allocatePinned(const InferenceEngine::TensorDesc& desc);
    void pinUserMemory(const void* ptr, size_t bytes);

    std::shared_ptr<ExecutableNetwork> _network;
    std::vector<b200_request*> _reqs;                      // one per shard
    std::shared_ptr<InferenceEngine::IAllocator> _allocator;
    std::vector<std::string> _inputNames, _outputNames;   // in device I/O index order
    std::vector<uint64_t> _inShardBytes, _outShardBytes;
    std::vector<std::pair<int, int>> _inFormat, _outFormat;   // (dtype, layout) the device side is currently recorded for
    void syncFormat(bool output, size_t index, const InferenceEngine::Blob::Ptr& blob);
    std::vector<std::vector<void*>> _outStage;             // [output][shard] pinned staging (Detections / Replicated outputs of shards > 0)
    std::vector<void*> _stageAllocs;
    std::map<const void*, size_t> _registered;             // user memory page-locked by this request (cudaHostRegister)
    std::vector<void*> _mergeDst;                          // user blob pointers captured at start for the post-wait re-pack
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
