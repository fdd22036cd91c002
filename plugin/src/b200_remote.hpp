// RemoteContext / RemoteBlob of the B200 plugin (inference-engine/include/ie_remote_context.hpp:58-178): blobs that live in the GPU's
// HBM.  An infer request whose inputs / outputs are remote blobs moves no data over PCIe - the request's ingest / egress kernels read
// and write the blob through a device-to-device copy - which is the fix SURVEY.md 8f-2 names for the host-I/O-bound end-to-end path.
// The blob can wrap memory the caller already owns (ParamMap {"B200_DEVICE_PTR": void*}) or allocate its own; rmap()/wmap() stage through
// pinned host memory (copy back on unlock for writable mappings), like the reference GPU plugin's remote blobs do.
#pragma once
#include <ie_remote_context.hpp>

#include <memory>
#include <mutex>
#include <string>

#include "b200.h"

namespace B200Plugin {

static constexpr auto B200_PARAM_DEVICE_PTR = "B200_DEVICE_PTR";   // void*: device address wrapped by / backing the blob
static constexpr auto B200_PARAM_DEVICE_ID = "DEVICE_ID";          // int: CUDA ordinal

class RemoteContext : public InferenceEngine::RemoteContext, public std::enable_shared_from_this<RemoteContext> {
public:
    explicit RemoteContext(int device) : _device(device) {}
    std::string getDeviceName() const noexcept override { return "B200." + std::to_string(_device); }
    InferenceEngine::RemoteBlob::Ptr CreateBlob(const InferenceEngine::TensorDesc& desc, const InferenceEngine::ParamMap& params = {}) override;
    InferenceEngine::ParamMap getParams() const override { return {{B200_PARAM_DEVICE_ID, _device}}; }
    int device() const { return _device; }

private:
    int _device;
};

class RemoteBlob : public InferenceEngine::RemoteBlob {
public:
    RemoteBlob(const std::shared_ptr<RemoteContext>& ctx, const InferenceEngine::TensorDesc& desc, void* wrapped);
    ~RemoteBlob() override;

    InferenceEngine::ParamMap getParams() const override;
    std::string getDeviceName() const noexcept override { return _ctx->getDeviceName(); }
    std::shared_ptr<InferenceEngine::RemoteContext> getContext() const noexcept override { return _ctx; }

    void allocate() noexcept override;
    bool deallocate() noexcept override;
    InferenceEngine::LockedMemory<void> buffer() noexcept override { return rwmap(); }
    InferenceEngine::LockedMemory<const void> cbuffer() const noexcept override { return rmap(); }
    InferenceEngine::LockedMemory<void> rwmap() noexcept override;
    InferenceEngine::LockedMemory<const void> rmap() const noexcept override;
    InferenceEngine::LockedMemory<void> wmap() noexcept override;

    void* devicePtr() const { return _dptr; }   // what the infer request hands to the C ABI instead of a host pointer

protected:
    const std::shared_ptr<InferenceEngine::IAllocator>& getAllocator() const noexcept override { return _mapper; }
    void* getHandle() const noexcept override { return _dptr; }

private:
    std::shared_ptr<RemoteContext> _ctx;
    void* _dptr = nullptr;
    bool _owned = false;
    std::shared_ptr<InferenceEngine::IAllocator> _mapper;   // lock = map to pinned host memory, unlock = write back + release
};

}  // namespace B200Plugin
