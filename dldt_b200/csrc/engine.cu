// Compiled-network and infer-request objects + the C ABI entry points that touch CUDA.
// One request = one stream + one arena slice + one captured CUDA graph:
// start() enqueues H2D copies, a single cudaGraphLaunch and D2H copies; wait() blocks on an event.
#include <fstream>
#include <map>
#include "engine.h"

#include <dlfcn.h>

#include <chrono>
#include <thread>

using namespace b200;

namespace b200 {
const std::string& last_error();
}

b200_network::~b200_network() {
  if (weights_dev) {
    cudaSetDevice(device);
    cudaFree(weights_dev);
  }
}

b200_request::~b200_request() {
  if (!net) return;
  cudaSetDevice(net->device);
  if (stream) cudaStreamSynchronize(stream);
  if (graph_exec) cudaGraphExecDestroy(graph_exec);
  if (graph) cudaGraphDestroy(graph);
  if (done) cudaEventDestroy(done);
  for (char* p : in_dev) if (p) cudaFree(p);
  for (char* p : out_dev) if (p) cudaFree(p);
  if (arena) cudaFree(arena);
  if (stream) cudaStreamDestroy(stream);
}

RunCtx b200_request::ctx() const {
  RunCtx c;
  c.arena = arena;
  c.box_offset = net->box_offsets.data();
  c.weights = net->weights_dev;
  c.io_in = in_dev;
  c.io_out = out_dev;
  c.stream = stream;
  return c;
}

static int default_io_dtype(int graph_dtype, bool is_output) {
  // user-side precisions (cnn_network_ngraph_impl.cpp:115-125,147-155): f16/bf16 inputs are presented as
  // FP32, every non-FP32/I32 output becomes FP32, I64 -> I32
  if (graph_dtype == B200_I64 || graph_dtype == B200_I32) return B200_I32;
  if (is_output) return B200_F32;
  if (graph_dtype == B200_F16 || graph_dtype == B200_BF16) return B200_F32;
  return graph_dtype;
}

static int phys_layout_of(const LTensor& t) { return t.layout == L_NHWC8 ? PL_NHWC8 : PL_ROWMAJOR; }

// ingest: host-format staging -> parameter tensor; egress: result tensor -> host-format staging
// (MKLDNNGraph::PushInputData / PullOutputData, mkldnn_graph.cpp:708-802)
static void enqueue_ingest(const b200_request& r, const RunCtx& c, int i) {
  const LTensor& t = r.net->in_tensors[i];
  const IOFormat& f = r.in_fmt[i];
  cudaStream_t s = (cudaStream_t)c.stream;
  if (t.layout == L_S2D16) {
    int sl = f.layout == B200_LAYOUT_NHWC ? PL_NHWC_DENSE : PL_ROWMAJOR;
    launch_ingest_s2d(r.in_dev[i], f.dtype, sl, c.ptr(t.buf), t.dtype, (int)t.shape[0], (int)t.shape[1], (int)t.shape[2], (int)t.shape[3], t.s2d_pt,
                      t.s2d_pl, t.s2d_hb, t.s2d_wb, s);
  } else if (t.shape.size() == 4) {
    int sl = f.layout == B200_LAYOUT_NHWC ? PL_NHWC_DENSE : PL_ROWMAJOR;
    launch_convert4(r.in_dev[i], f.dtype, sl, c.ptr(t.buf), t.dtype, phys_layout_of(t), (int)t.shape[0], (int)t.shape[1], (int)t.shape[2],
                    (int)t.shape[3], s);
  } else {
    launch_convert_flat(r.in_dev[i], f.dtype, c.ptr(t.buf), t.dtype, numel(t.shape), s);
  }
}
static void enqueue_egress(const b200_request& r, const RunCtx& c, int i) {
  const LTensor& t = r.net->out_tensors[i];
  const IOFormat& f = r.out_fmt[i];
  cudaStream_t s = (cudaStream_t)c.stream;
  if (t.shape.size() == 4) {
    int dl = f.layout == B200_LAYOUT_NHWC ? PL_NHWC_DENSE : PL_ROWMAJOR;
    launch_convert4(c.ptr(t.buf), t.dtype, phys_layout_of(t), r.out_dev[i], f.dtype, dl, (int)t.shape[0], (int)t.shape[1], (int)t.shape[2],
                    (int)t.shape[3], s);
  } else {
    launch_convert_flat(c.ptr(t.buf), t.dtype, r.out_dev[i], f.dtype, numel(t.shape), s);
  }
}

void b200_request::enqueue_all(bool eager_events, std::vector<cudaEvent_t>* evs) {
  RunCtx c = ctx();
  for (size_t i = 0; i < in_dev.size(); ++i) enqueue_ingest(*this, c, (int)i);
  for (size_t k = 0; k < net->steps.size(); ++k) {
    if (eager_events) B200_CUDA(cudaEventRecord((*evs)[k], stream));
    net->steps[k].launch(c);
  }
  {
    cudaError_t le = cudaPeekAtLastError();
    if (le != cudaSuccess) B200_THROW(B200_ERR_CUDA, "kernel launch failed: " << cudaGetErrorString(le));
  }
  if (eager_events) B200_CUDA(cudaEventRecord((*evs)[net->steps.size()], stream));
  for (size_t i = 0; i < out_dev.size(); ++i) enqueue_egress(*this, c, (int)i);
}

// Times every alternative of every step that has some (5 event-bracketed launches after one warm-up, best of 5) with the real shapes
// and buffers, and keeps the fastest.  The buffers hold whatever the arena holds (nothing has run yet): the tensor-core and copy
// kernels are data independent in time.  Runs once per network.
namespace {
struct EventPair {   // RAII: a throwing launch must not leak the two timing events
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  EventPair() {
    B200_CUDA(cudaEventCreate(&e0));
    if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); e0 = nullptr; B200_THROW(B200_ERR_CUDA, "cudaEventCreate failed"); }
  }
  ~EventPair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};
}  // namespace

void b200_network::autotune(const b200::RunCtx& c, cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(tune_mu);
  if (tuned) return;
  bool any = false;
  for (auto& s : steps) any = any || s.alts.size() > 1;
  if (!any) { tuned = true; return; }
  EventPair ev;
  const char* dbg = getenv("B200_AUTOTUNE_DEBUG");
  // Decisions of an earlier load are reused instead of measured again: the table imported with the network (Export/ImportNetwork,
  // CACHE_DIR) first, then B200_TUNE_FILE ("<network>\t<step>\t<impl>" lines; new decisions are appended to it)
  const char* tune_file = getenv("B200_TUNE_FILE");
  std::map<std::string, std::string> cached = tune_preset;
  const std::string key = name + "@" + std::to_string((long long)arena_bytes);   // same graph at another batch size is another network
  if (tune_file) {
    std::ifstream in(tune_file);
    std::string line;
    while (std::getline(in, line)) {
      size_t a = line.find('\t'), b = a == std::string::npos ? a : line.find('\t', a + 1);
      if (b == std::string::npos || line.substr(0, a) != key) continue;
      cached.emplace(line.substr(a + 1, b - a - 1), line.substr(b + 1));
    }
  }
  std::string fresh;
  for (auto& s : steps) {
    if (s.alts.size() < 2) continue;
    auto hit = cached.find(s.name);
    if (hit != cached.end()) {
      bool found = false;
      for (auto& alt : s.alts)
        if (alt.first == hit->second) { s.impl = alt.first; s.launch = alt.second; found = true; break; }
      if (found) continue;
    }
    int best = 0;
    float best_ms = 0.f;
    for (size_t a = 0; a < s.alts.size(); ++a) {
      float ms_min = 1e30f;
      for (int rep = 0; rep < 6; ++rep) {
        B200_CUDA(cudaEventRecord(ev.e0, stream));
        s.alts[a].second(c);
        B200_CUDA(cudaEventRecord(ev.e1, stream));
        B200_CUDA(cudaEventSynchronize(ev.e1));
        float ms = 0.f;
        B200_CUDA(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
        if (rep > 0) ms_min = std::min(ms_min, ms);
      }
      if (dbg && dbg[0] == '1') fprintf(stderr, "b200 autotune %-24s %-44s %8.1f us\n", s.name.c_str(), s.alts[a].first.c_str(), ms_min * 1e3f);
      // the first candidate is the modelled choice: another one has to beat it by more than the measurement noise (3 %)
      if (a == 0 || ms_min < best_ms * 0.97f) { best = (int)a; best_ms = ms_min; }
    }
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) B200_THROW(B200_ERR_CUDA, "autotune launch failed: " << cudaGetErrorString(le));
    s.impl = s.alts[best].first;
    s.launch = s.alts[best].second;
    fresh += key + "\t" + s.name + "\t" + s.impl + "\n";
  }
  if (tune_file && !fresh.empty()) {
    std::ofstream out(tune_file, std::ios::app);
    out << fresh;
  }
  tuned = true;   // only after every measurement succeeded: a failed attempt is retried by the next request
}

std::string b200_network::tuning_text() const {
  std::string t;
  for (const auto& s : steps)
    if (s.alts.size() > 1) t += s.name + "\t" + s.impl + "\n";
  return t;
}

void b200_request::finalize() {
  if (finalized) return;
  B200_CUDA(cudaSetDevice(net->device));
  net->autotune(ctx(), stream);
  // a failed attempt (allocation, capture) leaves `finalized` false: the next start() retries, so staging buffers that already exist are kept
  for (size_t i = 0; i < in_fmt.size(); ++i) {
    in_fmt[i].bytes = numel(net->in_tensors[i].shape) * (int64_t)dtype_size(in_fmt[i].dtype);
    if (in_dev[i]) continue;
    B200_CUDA(cudaMalloc((void**)&in_dev[i], std::max<int64_t>(in_fmt[i].bytes, 16)));
    B200_CUDA(cudaMemsetAsync(in_dev[i], 0, std::max<int64_t>(in_fmt[i].bytes, 16), stream));
  }
  for (size_t i = 0; i < out_fmt.size(); ++i) {
    out_fmt[i].bytes = numel(net->out_tensors[i].shape) * (int64_t)dtype_size(out_fmt[i].dtype);
    if (out_dev[i]) continue;
    B200_CUDA(cudaMalloc((void**)&out_dev[i], std::max<int64_t>(out_fmt[i].bytes, 16)));
  }
  launches_per_infer = (int)(net->steps.size() + in_dev.size() + out_dev.size());
  // capture the whole schedule once; replay with one cudaGraphLaunch per inference
  const char* no_graph = getenv("B200_NO_CUDA_GRAPH");
  if (!(no_graph && no_graph[0] == '1')) {
    B200_CUDA(cudaStreamSynchronize(stream));
    B200_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    try {
      enqueue_all(false, nullptr);
    } catch (...) {
      cudaGraph_t tmp = nullptr;
      cudaStreamEndCapture(stream, &tmp);
      if (tmp) cudaGraphDestroy(tmp);
      throw;
    }
    B200_CUDA(cudaStreamEndCapture(stream, &graph));
    B200_CUDA(cudaGraphInstantiate(&graph_exec, graph, 0));
  }
  finalized = true;
}

namespace {

template <typename F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return B200_ERR_INVALID;
  }
}

void fill_info(b200_tensor_info* info, const std::string& name, int dtype, const Shape& shape) {
  memset(info, 0, sizeof(*info));
  snprintf(info->name, sizeof(info->name), "%s", name.c_str());
  info->dtype = dtype;
  info->rank = (int)shape.size();
  for (size_t k = 0; k < shape.size() && k < B200_MAX_RANK; ++k) info->shape[k] = shape[k];
}

// The CUDA runtime is linked statically into this library: unloading the library (the Inference Engine unloads a plugin and what it
// links when the last Core that used it goes away) tears the runtime and the primary context down, and the next Core pays context
// creation again (0.3 - 3 s on a B200 box; the reference's behaviour suites create a Core per test).  Once a device has been
// touched the library pins itself in the process (RTLD_NODELETE): contexts, the module and its kernels stay loaded until exit.
void pin_library_once() {
  static std::once_flag once;
  std::call_once(once, [] {
    Dl_info info;
    if (dladdr(reinterpret_cast<const void*>(&pin_library_once), &info) && info.dli_fname) dlopen(info.dli_fname, RTLD_NOW | RTLD_NODELETE);
  });
}

void launch_graph_or_eager(b200_request* r) {
  if (r->graph_exec) B200_CUDA(cudaGraphLaunch(r->graph_exec, r->stream));
  else r->enqueue_all(false, nullptr);
}

}  // namespace

extern "C" {

int b200_device_count(int* count) {
  return guarded([&] {
    B200_CHECK(count, B200_ERR_INVALID, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (n > 0) pin_library_once();
    *count = n;
    return B200_OK;
  });
}
int b200_device_name(int device, char* buf, size_t len) {
  return guarded([&] {
    cudaDeviceProp p;
    B200_CUDA(cudaGetDeviceProperties(&p, device));
    snprintf(buf, len, "%s", p.name);
    return B200_OK;
  });
}
int b200_device_props(int device, int* sm_count, int* cc_major, int* cc_minor, uint64_t* mem_bytes) {
  return guarded([&] {
    cudaDeviceProp p;
    B200_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (mem_bytes) *mem_bytes = p.totalGlobalMem;
    return B200_OK;
  });
}

int b200_host_alloc(void** ptr, size_t bytes) {
  return guarded([&] {
    B200_CHECK(ptr, B200_ERR_INVALID, "null argument");
    B200_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocPortable));
    return B200_OK;
  });
}
int b200_host_free(void* ptr) {
  return guarded([&] { if (ptr) B200_CUDA(cudaFreeHost(ptr)); return B200_OK; });
}
int b200_host_register(void* ptr, size_t bytes) {
  return guarded([&] {
    B200_CHECK(ptr && bytes, B200_ERR_INVALID, "null argument");
    // memory that is already page-locked (b200_host_alloc, or registered by somebody else) is left alone
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered) return (int)B200_ALREADY_PINNED;
    cudaGetLastError();
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
      cudaGetLastError();   // not sticky: must not surface later as a "kernel launch failed"
      B200_THROW(B200_ERR_CUDA, "cudaHostRegister(" << bytes << " bytes) failed: " << cudaGetErrorString(e));
    }
    return (int)B200_OK;
  });
}
int b200_host_unregister(void* ptr) {
  return guarded([&] {
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); B200_THROW(B200_ERR_CUDA, "cudaHostUnregister failed: " << cudaGetErrorString(e)); }
    return (int)B200_OK;
  });
}

int b200_device_alloc(int device, void** dptr, size_t bytes) {
  return guarded([&] {
    B200_CHECK(dptr, B200_ERR_INVALID, "null argument");
    B200_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 16);
    if (e != cudaSuccess) { cudaGetLastError(); B200_THROW(B200_ERR_OOM, "cannot allocate " << bytes << " bytes of device memory"); }
    return (int)B200_OK;
  });
}
int b200_device_free(int device, void* dptr) {
  return guarded([&] {
    if (!dptr) return (int)B200_OK;
    B200_CUDA(cudaSetDevice(device));
    B200_CUDA(cudaFree(dptr));
    return (int)B200_OK;
  });
}
int b200_device_copy(int device, void* dst, const void* src, size_t bytes, int direction) {
  return guarded([&] {
    B200_CHECK(dst && src, B200_ERR_INVALID, "null argument");
    B200_CUDA(cudaSetDevice(device));
    B200_CUDA(cudaMemcpy(dst, src, bytes, direction == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
    return (int)B200_OK;
  });
}

int b200_compile(const b200_graph* g, const b200_compile_options* opts, b200_network** out) {
  return guarded([&] {
    B200_CHECK(g && opts && out, B200_ERR_INVALID, "null argument");
    B200_CHECK(opts->abi_version == B200_ABI_VERSION, B200_ERR_INVALID, "ABI version mismatch");
    std::unique_ptr<b200_network> net(new b200_network());
    net->device = opts->device;
    if (opts->device >= 0) {
      // bring the runtime + primary context up BEFORE lowering: the compiler asks the driver which tensor maps it can encode
      // (cudaGetDriverEntryPoint fails on an uninitialised runtime, which silently cost the space-to-depth stem: profiles/README.md)
      int n0 = 0;
      if (cudaGetDeviceCount(&n0) == cudaSuccess && opts->device < n0) { pin_library_once(); cudaSetDevice(opts->device); cudaFree(nullptr); }
      cudaGetLastError();
    }
    lower_graph(*g, *opts, *net);
    if (opts->device >= 0) {
      int n = 0;
      cudaError_t e = cudaGetDeviceCount(&n);
      if (e != cudaSuccess || n == 0) { cudaGetLastError(); B200_THROW(B200_ERR_NO_DEVICE, "no CUDA device available (the B200 plugin has no CPU fallback)"); }
      B200_CHECK(opts->device < n, B200_ERR_NO_DEVICE, "DEVICE_ID " << opts->device << " out of range (" << n << " devices)");
      cudaDeviceProp p;
      B200_CUDA(cudaGetDeviceProperties(&p, opts->device));
      B200_CHECK(p.major == 10, B200_ERR_NO_DEVICE, "device " << opts->device << " (" << p.name << ", sm_" << p.major << p.minor
                                                              << ") is not an sm_100 part; kernels are built for sm_100a only");
      B200_CUDA(cudaSetDevice(opts->device));
      conv_tc_init_device();
      conv_tc2_init_device();
      conv_tc3_init_device();
      mbconv_init_device();
      conv_tc3f_init_device();
      detect_init_device();
      attention_init_device();
      size_t wb = std::max<size_t>(net->weights_host.size(), 16);
      B200_CUDA(cudaMalloc((void**)&net->weights_dev, wb));
      if (!net->weights_host.empty())
        B200_CUDA(cudaMemcpy(net->weights_dev, net->weights_host.data(), net->weights_host.size(), cudaMemcpyHostToDevice));
    }
    *out = net.release();
    return B200_OK;
  });
}
void b200_network_destroy(b200_network* net) { delete net; }
int b200_network_num_inputs(const b200_network* net) { return net ? (int)net->inputs.size() : B200_ERR_INVALID; }
int b200_network_num_outputs(const b200_network* net) { return net ? (int)net->outputs.size() : B200_ERR_INVALID; }
int b200_network_io_info(const b200_network* net, int is_output, int index, b200_tensor_info* info) {
  return guarded([&] {
    B200_CHECK(net && info, B200_ERR_INVALID, "null argument");
    const auto& v = is_output ? net->outputs : net->inputs;
    B200_CHECK(index >= 0 && index < (int)v.size(), B200_ERR_INVALID, "io index out of range");
    fill_info(info, v[index].name, v[index].graph_dtype, v[index].shape);
    return B200_OK;
  });
}
int b200_network_num_exec_nodes(const b200_network* net) { return net ? (int)net->steps.size() : B200_ERR_INVALID; }
int b200_network_exec_node(const b200_network* net, int index, b200_exec_node_info* info) {
  return guarded([&] {
    B200_CHECK(net && info && index >= 0 && index < (int)net->steps.size(), B200_ERR_INVALID, "bad exec node index");
    const Step& s = net->steps[index];
    memset(info, 0, sizeof(*info));
    snprintf(info->name, sizeof(info->name), "%s", s.name.c_str());
    snprintf(info->type, sizeof(info->type), "%s", s.type.c_str());
    snprintf(info->impl, sizeof(info->impl), "%s", s.impl.c_str());
    snprintf(info->fused, sizeof(info->fused), "%s", s.fused.c_str());
    snprintf(info->out_precision, sizeof(info->out_precision), "%s", s.out_precision.c_str());
    snprintf(info->out_layout, sizeof(info->out_layout), "%s", s.out_layout.c_str());
    info->exec_order = index;
    info->flops = s.flops;
    info->bytes = s.bytes;
    return B200_OK;
  });
}
int b200_network_stat(const b200_network* net, const char* key, double* value) {
  return guarded([&] {
    B200_CHECK(net && key && value, B200_ERR_INVALID, "null argument");
    std::string k = key;
    if (k == "arena_bytes") *value = (double)net->arena_bytes;
    else if (k == "weight_bytes") *value = (double)net->weights_host.size();
    else if (k == "flops") *value = net->flops;
    else if (k == "bytes") *value = net->bytes;
    else if (k == "num_kernels") *value = (double)net->steps.size();
    else if (k == "device") *value = net->device;
    else if (k == "precision") *value = net->precision;
    else B200_THROW(B200_ERR_INVALID, "unknown stat key " << k);
    return B200_OK;
  });
}

// Kernel selection now instead of at the first request: a temporary arena + stream, the alternatives of every step timed on it.
int b200_network_tune(b200_network* net) {
  return guarded([&] {
    B200_CHECK(net, B200_ERR_INVALID, "null argument");
    B200_CHECK(net->device >= 0 && net->weights_dev, B200_ERR_NO_DEVICE, "network was compiled without a device (dry compile)");
    if (net->tuned) return (int)B200_OK;
    B200_CUDA(cudaSetDevice(net->device));
    struct Tmp {
      cudaStream_t stream = nullptr;
      char* arena = nullptr;
      ~Tmp() { if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); } if (arena) cudaFree(arena); }
    } tmp;
    B200_CUDA(cudaStreamCreateWithFlags(&tmp.stream, cudaStreamNonBlocking));
    cudaError_t e = cudaMalloc((void**)&tmp.arena, (size_t)net->arena_bytes);
    if (e != cudaSuccess) { cudaGetLastError(); B200_THROW(B200_ERR_OOM, "cannot allocate " << net->arena_bytes << " bytes of activation arena"); }
    B200_CUDA(cudaMemsetAsync(tmp.arena, 0, (size_t)net->arena_bytes, tmp.stream));
    RunCtx c;
    c.arena = tmp.arena;
    c.box_offset = net->box_offsets.data();
    c.weights = net->weights_dev;
    c.stream = tmp.stream;
    net->autotune(c, tmp.stream);
    return (int)B200_OK;
  });
}
int b200_network_get_tuning(const b200_network* net, char* buf, size_t capacity, size_t* needed) {
  return guarded([&] {
    B200_CHECK(net && needed, B200_ERR_INVALID, "null argument");
    const std::string t = net->tuning_text();
    *needed = t.size() + 1;
    if (buf && capacity >= t.size() + 1) memcpy(buf, t.c_str(), t.size() + 1);
    return (int)B200_OK;
  });
}
int b200_network_set_tuning(b200_network* net, const char* text) {
  return guarded([&] {
    B200_CHECK(net && text, B200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(net->tune_mu);
    B200_CHECK(!net->tuned, B200_ERR_BUSY, "kernel selection already done");
    std::string all = text;
    size_t pos = 0;
    while (pos < all.size()) {
      size_t nl = all.find('\n', pos);
      if (nl == std::string::npos) nl = all.size();
      const std::string line = all.substr(pos, nl - pos);
      pos = nl + 1;
      const size_t tab = line.find('\t');
      if (tab == std::string::npos) continue;
      net->tune_preset[line.substr(0, tab)] = line.substr(tab + 1);
    }
    return (int)B200_OK;
  });
}

int b200_request_create(b200_network* net, b200_request** out) {
  return guarded([&] {
    B200_CHECK(net && out, B200_ERR_INVALID, "null argument");
    B200_CHECK(net->device >= 0 && net->weights_dev, B200_ERR_NO_DEVICE, "network was compiled without a device (dry compile)");
    B200_CUDA(cudaSetDevice(net->device));
    std::unique_ptr<b200_request> r(new b200_request());
    r->net = net;
    B200_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    B200_CUDA(cudaEventCreateWithFlags(&r->done, cudaEventDisableTiming));
    cudaError_t e = cudaMalloc((void**)&r->arena, (size_t)net->arena_bytes);
    if (e != cudaSuccess) { cudaGetLastError(); B200_THROW(B200_ERR_OOM, "cannot allocate " << net->arena_bytes << " bytes of activation arena"); }
    r->in_fmt.resize(net->inputs.size());
    r->out_fmt.resize(net->outputs.size());
    r->in_dev.assign(net->inputs.size(), nullptr);
    r->out_dev.assign(net->outputs.size(), nullptr);
    for (size_t i = 0; i < net->inputs.size(); ++i) r->in_fmt[i].dtype = default_io_dtype(net->inputs[i].graph_dtype, false);
    for (size_t i = 0; i < net->outputs.size(); ++i) r->out_fmt[i].dtype = default_io_dtype(net->outputs[i].graph_dtype, true);
    *out = r.release();
    return B200_OK;
  });
}
void b200_request_destroy(b200_request* req) { delete req; }

int b200_request_set_io_format(b200_request* req, int is_output, int index, int32_t dtype, int32_t layout) {
  return guarded([&] {
    B200_CHECK(req, B200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(req->mu);
    auto& v = is_output ? req->out_fmt : req->in_fmt;
    B200_CHECK(index >= 0 && index < (int)v.size(), B200_ERR_INVALID, "io index out of range");
    B200_CHECK(dtype == B200_F32 || dtype == B200_F16 || dtype == B200_BF16 || dtype == B200_I32 || dtype == B200_U8, B200_ERR_UNSUPPORTED,
               "unsupported blob precision " << dtype);
    const int new_layout = layout == B200_LAYOUT_NHWC ? B200_LAYOUT_NHWC : B200_LAYOUT_NCHW;
    if (v[index].dtype == dtype && v[index].layout == new_layout) return B200_OK;
    if (req->finalized) {
      // a blob with another layout / precision arrived after the first inference (SetBlob in the reference may change both at any time,
      // ie_iinfer_request_internal.cpp:96-150): the edge-conversion kernels are part of the captured graph, so the capture is dropped and
      // the next start() records it again with the new format.  A request that is still running is drained first.
      B200_CUDA(cudaSetDevice(req->net->device));
      B200_CUDA(cudaStreamSynchronize(req->stream));
      if (req->graph_exec) { cudaGraphExecDestroy(req->graph_exec); req->graph_exec = nullptr; }
      if (req->graph) { cudaGraphDestroy(req->graph); req->graph = nullptr; }
      auto& dev = is_output ? req->out_dev : req->in_dev;
      if (dtype_size(dtype) != dtype_size(v[index].dtype) && dev[index]) { cudaFree(dev[index]); dev[index] = nullptr; }
      req->finalized = false;
    }
    v[index].dtype = dtype;
    v[index].layout = new_layout;
    return B200_OK;
  });
}
int b200_request_io_bytes(const b200_request* req, int is_output, int index, uint64_t* bytes) {
  return guarded([&] {
    B200_CHECK(req && bytes, B200_ERR_INVALID, "null argument");
    const auto& v = is_output ? req->out_fmt : req->in_fmt;
    const auto& t = is_output ? req->net->out_tensors : req->net->in_tensors;
    B200_CHECK(index >= 0 && index < (int)v.size(), B200_ERR_INVALID, "io index out of range");
    *bytes = (uint64_t)(numel(t[index].shape) * (int64_t)dtype_size(v[index].dtype));
    return B200_OK;
  });
}

int b200_request_start(b200_request* req, const void* const* host_inputs, void* const* host_outputs) {
  return guarded([&] {
    B200_CHECK(req, B200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(req->mu);
    B200_CUDA(cudaSetDevice(req->net->device));
    req->finalize();
    for (size_t i = 0; i < req->in_dev.size(); ++i) {
      B200_CHECK(host_inputs && host_inputs[i], B200_ERR_INVALID, "missing host input " << i);
      B200_CUDA(cudaMemcpyAsync(req->in_dev[i], host_inputs[i], (size_t)req->in_fmt[i].bytes, cudaMemcpyDefault, req->stream));
    }
    launch_graph_or_eager(req);
    for (size_t i = 0; i < req->out_dev.size(); ++i) {
      B200_CHECK(host_outputs && host_outputs[i], B200_ERR_INVALID, "missing host output " << i);
      B200_CUDA(cudaMemcpyAsync(host_outputs[i], req->out_dev[i], (size_t)req->out_fmt[i].bytes, cudaMemcpyDefault, req->stream));
    }
    B200_CUDA(cudaEventRecord(req->done, req->stream));
    req->started = true;
    return B200_OK;
  });
}

int b200_request_wait(b200_request* req, int64_t timeout_ms) {
  return guarded([&] {
    B200_CHECK(req, B200_ERR_INVALID, "null argument");
    B200_CUDA(cudaSetDevice(req->net->device));
    if (!req->started.load()) return (int)B200_OK;
    if (timeout_ms < 0) {
      B200_CUDA(cudaEventSynchronize(req->done));
      return (int)B200_OK;
    }
    auto t0 = std::chrono::steady_clock::now();
    while (true) {
      cudaError_t e = cudaEventQuery(req->done);
      if (e == cudaSuccess) return (int)B200_OK;
      if (e != cudaErrorNotReady) B200_CUDA(e);
      auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
      if (ms >= timeout_ms) return (int)B200_NOT_READY;
      std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
  });
}

int b200_request_device_io(b200_request* req, int is_output, int index, void** dptr, uint64_t* bytes) {
  return guarded([&] {
    B200_CHECK(req && dptr, B200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(req->mu);
    B200_CUDA(cudaSetDevice(req->net->device));
    req->finalize();
    auto& v = is_output ? req->out_dev : req->in_dev;
    auto& f = is_output ? req->out_fmt : req->in_fmt;
    B200_CHECK(index >= 0 && index < (int)v.size(), B200_ERR_INVALID, "io index out of range");
    *dptr = v[index];
    if (bytes) *bytes = (uint64_t)f[index].bytes;
    return B200_OK;
  });
}

int b200_request_launch(b200_request* req) {
  return guarded([&] {
    B200_CHECK(req, B200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(req->mu);
    B200_CUDA(cudaSetDevice(req->net->device));
    req->finalize();
    launch_graph_or_eager(req);
    B200_CUDA(cudaEventRecord(req->done, req->stream));
    req->started = true;
    return B200_OK;
  });
}

int b200_request_time(b200_request* req, int iters, int mode, const void* const* host_inputs, void* const* host_outputs, float* total_ms) {
  return guarded([&] {
    B200_CHECK(req && total_ms && iters > 0, B200_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(req->mu);
    B200_CUDA(cudaSetDevice(req->net->device));
    req->finalize();
    cudaEvent_t e0, e1;
    B200_CUDA(cudaEventCreate(&e0));
    B200_CUDA(cudaEventCreate(&e1));
    B200_CUDA(cudaStreamSynchronize(req->stream));
    B200_CUDA(cudaEventRecord(e0, req->stream));
    for (int it = 0; it < iters; ++it) {
      if (mode == 1) {
        for (size_t i = 0; i < req->in_dev.size(); ++i)
          B200_CUDA(cudaMemcpyAsync(req->in_dev[i], host_inputs[i], (size_t)req->in_fmt[i].bytes, cudaMemcpyDefault, req->stream));
      }
      launch_graph_or_eager(req);
      if (mode == 1) {
        for (size_t i = 0; i < req->out_dev.size(); ++i)
          B200_CUDA(cudaMemcpyAsync(host_outputs[i], req->out_dev[i], (size_t)req->out_fmt[i].bytes, cudaMemcpyDefault, req->stream));
      }
    }
    B200_CUDA(cudaEventRecord(e1, req->stream));
    B200_CUDA(cudaEventSynchronize(e1));
    B200_CUDA(cudaEventElapsedTime(total_ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return B200_OK;
  });
}

int b200_request_perf_counts(b200_request* req, float* usec, int capacity) {
  return guarded([&] {
    B200_CHECK(req && usec, B200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(req->mu);
    B200_CUDA(cudaSetDevice(req->net->device));
    req->finalize();
    const size_t n = req->net->steps.size();
    std::vector<cudaEvent_t> evs(n + 1);
    for (auto& e : evs) B200_CUDA(cudaEventCreate(&e));
    B200_CUDA(cudaStreamSynchronize(req->stream));
    req->enqueue_all(true, &evs);
    B200_CUDA(cudaStreamSynchronize(req->stream));
    for (size_t k = 0; k < n && (int)k < capacity; ++k) {
      float ms = 0;
      B200_CUDA(cudaEventElapsedTime(&ms, evs[k], evs[k + 1]));
      usec[k] = ms * 1000.f;
    }
    for (auto& e : evs) cudaEventDestroy(e);
    return B200_OK;
  });
}

int b200_request_num_launches(const b200_request* req, int* kernels_per_infer) {
  if (!req || !kernels_per_infer) return B200_ERR_INVALID;
  *kernels_per_infer = (int)(req->net->steps.size() + req->net->inputs.size() + req->net->outputs.size());
  return B200_OK;
}

int b200_request_stream(b200_request* req, void** stream) {
  if (!req || !stream) return B200_ERR_INVALID;
  *stream = req->stream;
  return B200_OK;
}

int b200_request_dump_node(b200_request* req, int index, float* out, uint64_t count) {
  return guarded([&] {
    B200_CHECK(req && out, B200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(req->mu);
    B200_CUDA(cudaSetDevice(req->net->device));
    req->finalize();
    B200_CHECK(index >= 0 && index < (int)req->net->steps.size(), B200_ERR_INVALID, "bad exec node index");
    // re-run eagerly up to and including `index`, then convert that node's output to dense f32
    RunCtx c = req->ctx();
    for (size_t i = 0; i < req->in_dev.size(); ++i) enqueue_ingest(*req, c, (int)i);
    for (int k = 0; k <= index; ++k) req->net->steps[k].launch(c);
    const LTensor& t = req->net->steps[index].out;
    int64_t n = numel(t.shape);
    B200_CHECK((uint64_t)n <= count, B200_ERR_INVALID, "dump buffer too small: need " << n);
    float* tmp = nullptr;
    B200_CUDA(cudaMalloc((void**)&tmp, std::max<int64_t>(n, 4) * 4));
    if (t.shape.size() == 4) launch_convert4(c.ptr(t.buf), t.dtype, phys_layout_of(t), tmp, B200_F32, PL_ROWMAJOR, (int)t.shape[0], (int)t.shape[1], (int)t.shape[2], (int)t.shape[3], req->stream);
    else launch_convert_flat(c.ptr(t.buf), t.dtype, tmp, B200_F32, n, req->stream);
    B200_CUDA(cudaMemcpyAsync(out, tmp, n * 4, cudaMemcpyDeviceToHost, req->stream));
    B200_CUDA(cudaStreamSynchronize(req->stream));
    cudaFree(tmp);
    return B200_OK;
  });
}

}  // extern "C"
