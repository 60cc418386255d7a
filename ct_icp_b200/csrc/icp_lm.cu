// icp_lm.cu — solver CERES reproduced on the device (placeholder until the LM/IRLS kernels land).
#include "engine.h"
#include "icp.h"

namespace cticp {

void IcpSolver::EnqueueCeres(const DeviceMap &, const cticp_icp_options &, const cticp_strategy_options &,
                             const float4 *, const int *, size_t, IcpState *, int, int, void *) {
    throw UnsupportedError("solver CERES: device LM/IRLS path not built yet");
}

}  // namespace cticp
