// Compiled network + infer request objects behind the opaque C handles.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <mutex>

#include "common.h"
#include "kernels.h"

#define B200_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      B200_THROW(B200_ERR_CUDA, "CUDA error " << cudaGetErrorName(_e) << " (" << cudaGetErrorString(_e) \
                                              << ") at " << __FILE__ << ":" << __LINE__ << ": " #expr); \
  } while (0)

struct b200_network {
  std::string name;
  int device = 0;
  int precision = B200_PREC_FP32;
  int cdtype = B200_F32;  // compute dtype of float activations
  std::vector<b200::Step> steps;
  std::vector<b200::ArenaBox> boxes;
  std::vector<int64_t> box_offsets;
  int64_t arena_bytes = 0;
  std::vector<uint8_t> weights_host;
  char* weights_dev = nullptr;
  std::vector<b200::IOBinding> inputs, outputs;
  std::vector<b200::LTensor> in_tensors, out_tensors;
  double flops = 0, bytes = 0;
  // kernel selection by measurement: steps that carry alternatives are timed once, on the first request's arena and stream
  std::mutex tune_mu;
  bool tuned = false;
  std::map<std::string, std::string> tune_preset;   // step name -> impl: decisions imported with the network (b200_network_set_tuning)
  void autotune(const b200::RunCtx& c, cudaStream_t stream);
  std::string tuning_text() const;                  // "step\timpl\n" for every step that had alternatives
  ~b200_network();
};

struct IOFormat {
  int dtype = B200_F32;
  int layout = B200_LAYOUT_NCHW;
  int64_t bytes = 0;
};

struct b200_request {
  b200_network* net = nullptr;
  cudaStream_t stream = nullptr;
  char* arena = nullptr;
  std::vector<IOFormat> in_fmt, out_fmt;
  std::vector<char*> in_dev, out_dev;
  bool finalized = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  cudaEvent_t done = nullptr;
  std::atomic<bool> started{false};
  int launches_per_infer = 0;
  std::mutex mu;
  ~b200_request();
  void finalize();
  void enqueue_all(bool eager_events, std::vector<cudaEvent_t>* evs);
  b200::RunCtx ctx() const;
};

namespace b200 {
// graph compiler entry (compile.cu)
void lower_graph(const b200_graph& g, const b200_compile_options& opts, b200_network& net);
void plan_arena(std::vector<ArenaBox>& boxes, int64_t& total_bytes);
}  // namespace b200
