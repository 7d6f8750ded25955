// BSVD clip program (placeholder until the streaming engine in bsvd.cu lands).
#include "program.h"

namespace ss4k {
std::string build_bsvd_clip(const PlanCfgLite&, Program*) {
  return "BSVD plans are created through the streaming engine (ss4k_bsvd_stream_*)";
}
}  // namespace ss4k
