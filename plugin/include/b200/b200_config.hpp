// Public config keys of the B200 device plugin (analogue of docs/template_plugin/include/template/template_config.hpp
// and of ENFORCE_BF16 in inference-engine/include/ie_plugin_config.hpp:373).
#pragma once
#include <string>

namespace InferenceEngine {
namespace B200ConfigParams {

// "FP32" | "BF16" | "FP16": arithmetic of the compiled network (default BF16, like the CPU plugin on a bf16-capable host)
static constexpr auto KEY_B200_INFERENCE_PRECISION = "B200_INFERENCE_PRECISION";
// number of device streams; OPTIMAL_NUMBER_OF_INFER_REQUESTS = 2 x streams (naming rule <DEV>_THROUGHPUT_STREAMS,
// inference-engine/samples/benchmark_app/main.cpp:228)
static constexpr auto KEY_B200_THROUGHPUT_STREAMS = "B200_THROUGHPUT_STREAMS";
// "0,1,2,3": replicate the compiled network on these GPUs of the box and shard the batch of every infer request across them
// (batch N -> N/G contiguous rows per device, no collective).  DEVICE_ID=ALL ("-d B200.ALL") is the same over every visible GPU.
// The reference's own multi-device mode hands whole requests to idle devices instead (MULTI:B200.0,B200.1,..,
// inference-engine/src/multi_device/multi_device_exec_network.cpp:49-136); both work with this plugin.
static constexpr auto KEY_B200_DEVICES = "B200_DEVICES";

}  // namespace B200ConfigParams
}  // namespace InferenceEngine
