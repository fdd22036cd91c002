// RemoteContext / RemoteBlob.  See b200_remote.hpp.
#include "b200_remote.hpp"

#include <ie_common.h>

#include <cstring>
#include <map>

using namespace InferenceEngine;

namespace B200Plugin {

namespace {

// IAllocator used as the blob's "mapper": lock(handle) stages the device memory in pinned host memory, unlock(handle) writes a
// writable mapping back and releases the staging buffer.  handle = the device pointer.
class Mapper : public IAllocator {
public:
    Mapper(int device, size_t bytes) : _device(device), _bytes(bytes) {}
    void* lock(void* handle, LockOp op = LOCK_FOR_WRITE) noexcept override {
        if (!handle) return nullptr;
        std::lock_guard<std::mutex> lk(_mu);
        Mapping& m = _maps[handle];
        if (m.host == nullptr) {
            if (b200_host_alloc(&m.host, _bytes ? _bytes : 16) != B200_OK) { _maps.erase(handle); return nullptr; }
            // a writable mapping also starts from the device contents: buffer() / rwmap() promise read access
            if (b200_device_copy(_device, m.host, handle, _bytes, 1) != B200_OK) { b200_host_free(m.host); _maps.erase(handle); return nullptr; }
        }
        m.writable = m.writable || op == LOCK_FOR_WRITE;
        ++m.users;
        return m.host;
    }
    void unlock(void* handle) noexcept override {
        std::lock_guard<std::mutex> lk(_mu);
        auto it = _maps.find(handle);
        if (it == _maps.end() || --it->second.users > 0) return;
        if (it->second.writable) b200_device_copy(_device, handle, it->second.host, _bytes, 0);
        b200_host_free(it->second.host);
        _maps.erase(it);
    }
    void* alloc(size_t) noexcept override { return nullptr; }     // device memory is owned by the blob, not by the mapper
    bool free(void*) noexcept override { return false; }

private:
    struct Mapping { void* host = nullptr; int users = 0; bool writable = false; };
    int _device;
    size_t _bytes;
    std::mutex _mu;
    std::map<void*, Mapping> _maps;
};

}  // namespace

InferenceEngine::RemoteBlob::Ptr RemoteContext::CreateBlob(const TensorDesc& desc, const ParamMap& params) {
    void* wrapped = nullptr;
    auto it = params.find(B200_PARAM_DEVICE_PTR);
    if (it != params.end()) wrapped = it->second.as<void*>();
    auto blob = std::make_shared<RemoteBlob>(shared_from_this(), desc, wrapped);
    if (!wrapped) blob->allocate();
    return blob;
}

RemoteBlob::RemoteBlob(const std::shared_ptr<RemoteContext>& ctx, const TensorDesc& desc, void* wrapped)
    : InferenceEngine::RemoteBlob(desc), _ctx(ctx), _dptr(wrapped), _owned(false), _mapper(std::make_shared<Mapper>(ctx->device(), byteSize())) {}

RemoteBlob::~RemoteBlob() { deallocate(); }

ParamMap RemoteBlob::getParams() const { return {{B200_PARAM_DEVICE_PTR, _dptr}, {B200_PARAM_DEVICE_ID, _ctx->device()}}; }

void RemoteBlob::allocate() noexcept {
    if (_dptr) return;
    void* p = nullptr;
    if (b200_device_alloc(_ctx->device(), &p, byteSize() ? byteSize() : 16) == B200_OK) {
        _dptr = p;
        _owned = true;
    }
}

bool RemoteBlob::deallocate() noexcept {
    if (_dptr && _owned) b200_device_free(_ctx->device(), _dptr);
    const bool had = _dptr != nullptr;
    _dptr = nullptr;
    _owned = false;
    return had;
}

LockedMemory<void> RemoteBlob::rwmap() noexcept { return LockedMemory<void>(_mapper.get(), _dptr, 0); }
LockedMemory<const void> RemoteBlob::rmap() const noexcept { return LockedMemory<const void>(_mapper.get(), _dptr, 0); }
LockedMemory<void> RemoteBlob::wmap() noexcept { return LockedMemory<void>(_mapper.get(), _dptr, 0); }

}  // namespace B200Plugin
