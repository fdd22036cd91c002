// fp16 instantiations of the CTA-pair kernel (conv_tc3.cuh); split from conv_tc3.cu so the two element types compile in parallel.
#include "conv_tc3.cuh"

namespace b200 {
namespace tc3 {
B200_TC3_DEFINE_TYPED(__half)
}
void conv_tc3_f16_opt_in() { tc3::opt_in_tc3_typed<__half>(); }
}  // namespace b200
