// k_team_run<7, 1>: the persistent RBCD kernel for relaxation rank r = 7, RGD local solver
#include "team_run.cuh"

namespace dpgo {
template cudaError_t launch_run_t<7, 1, false>(const TeamDev &, RunArgs, int, cudaStream_t);
}  // namespace dpgo
