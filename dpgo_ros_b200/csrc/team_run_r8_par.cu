// k_team_run<8, 2>: the persistent RBCD kernel for relaxation rank r = 8, RGD under the parallel
// (asynchronous-mode) schedule
#include "team_run.cuh"

namespace dpgo {
template cudaError_t launch_run_t<8, 2, false>(const TeamDev &, RunArgs, int, cudaStream_t);
}  // namespace dpgo
