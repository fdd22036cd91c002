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

}  // namespace B200ConfigParams
}  // namespace InferenceEngine
