/* DPGO/DPGO_solver.h -- included by src/PGOAgentROS.cpp:10.  Upstream it declares the local solver
 * (QuadraticProblem / QuadraticOptimizer over ROPTLIB); here the solver IS the CUDA library behind
 * dpgo_b200_iterate, so only the types the wrapper names remain (DPGO_types.h). */
#ifndef DPGO_SHIM_SOLVER_H
#define DPGO_SHIM_SOLVER_H
#include "DPGO/DPGO_types.h"
#endif
