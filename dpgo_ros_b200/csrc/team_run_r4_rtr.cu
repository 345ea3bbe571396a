// k_team_run<4, 0>: the persistent RBCD kernel for relaxation rank r = 4, RTR-tCG local solver
#include "team_run.cuh"

namespace dpgo {
template cudaError_t launch_run_t<4, 0, false>(const TeamDev &, RunArgs, int, cudaStream_t);
}  // namespace dpgo
