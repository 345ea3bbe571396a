// k_team_run<8, 1 / 2, BIG>: RGD kernels (synchronous and parallel schedule) for teams with an agent whose
// preconditioner does not fit shared memory: the dense pass streams it from HBM (BASELINE config 5)
#include "team_run.cuh"

namespace dpgo {
template cudaError_t launch_run_t<8, 1, true>(const TeamDev &, RunArgs, int, cudaStream_t);
template cudaError_t launch_run_t<8, 2, true>(const TeamDev &, RunArgs, int, cudaStream_t);
}  // namespace dpgo
