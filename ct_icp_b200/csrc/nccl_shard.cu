// nccl_shard.cu — keypoint-sharded multi-GPU mode: one NCCL all-reduce of the 96-double accumulator per iteration.
#include "engine.h"
#include "icp.h"

namespace cticp {

void IcpSolver::AllReduceAccumulator(void *) { throw UnsupportedError("multi-GPU sharding not built yet"); }
void Engine::EnableSharding(const void *, int, int) { throw UnsupportedError("multi-GPU sharding not built yet"); }

}  // namespace cticp

extern "C" int cticp_nccl_unique_id(void *) { return CTICP_ERR_UNSUPPORTED; }
