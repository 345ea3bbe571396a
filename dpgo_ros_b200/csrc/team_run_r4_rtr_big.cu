// k_team_run<4, 0, BIG>: RTR-tCG kernel for teams with an agent whose share of the dense preconditioner does not fit
// one CTA's shared memory (n > ~300 poses): every application streams Pinv from L2 / HBM with the register-prefetch pass
#include "team_run.cuh"

namespace dpgo {
template cudaError_t launch_run_t<4, 0, true>(const TeamDev &, RunArgs, int, cudaStream_t);
}  // namespace dpgo
