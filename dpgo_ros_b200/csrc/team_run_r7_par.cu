// k_team_run<7, 2>: the persistent RBCD kernel for relaxation rank r = 7, RGD under the parallel
// (asynchronous-mode) schedule
#include "team_run.cuh"

namespace dpgo {
template cudaError_t launch_run_t<7, 2, false>(const TeamDev &, RunArgs, int, cudaStream_t);
}  // namespace dpgo
