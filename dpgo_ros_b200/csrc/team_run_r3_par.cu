// k_team_run<3, 2>: the persistent RBCD kernel for relaxation rank r = 3, RGD under the parallel
// (asynchronous-mode) schedule
#include "team_run.cuh"

namespace dpgo {
template cudaError_t launch_run_t<3, 2, false>(const TeamDev &, RunArgs, int, cudaStream_t);
}  // namespace dpgo
