/*
 * dpgo_b200 -- C ABI of the B200-native RBCD hot path.
 *
 * This is the drop-in boundary under the DPGO::PGOAgent API surface that
 * dpgo_ros's PGOAgentROS subclasses (include/dpgo_ros/PGOAgentROS.h:121).
 * Every entry point names the reference call site whose arithmetic it
 * replaces (paths relative to the reference repo).  Plain pointers and sizes
 * only; every function returns 0 on success or a negative dpgo_b200_error.
 * No exceptions cross this boundary.  There is NO CPU fallback: every compute
 * call fails with DPGO_B200_ERR_CUDA when no CUDA device is usable.
 *
 * Array conventions
 *   X, Y, V      r x 4n doubles, column-major; pose i = columns 4i..4i+3
 *                ([Y_i | p_i], Y_i in St(3, r)) -- the layout of DPGO::Matrix X.
 *   poses        count x (r*4) doubles, each pose r x 4 column-major.
 *   R            m x 9 doubles, row-major 3x3 per measurement; t: m x 3.
 *   T (SE(3))    n x 3 x 4 doubles, row-major per pose ([R | t]).
 *   lifting Y    r x 3 doubles, column-major.
 */
#ifndef DPGO_B200_H
#define DPGO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dpgo_b200_agent_s *dpgo_b200_agent_t;
typedef struct dpgo_b200_team_s *dpgo_b200_team_t;

typedef enum {
  DPGO_B200_OK = 0,
  DPGO_B200_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  DPGO_B200_ERR_STATE = -2,     /* call not valid in the agent's current state */
  DPGO_B200_ERR_CUDA = -3,      /* CUDA runtime failure (or no device) */
  DPGO_B200_ERR_MISSING = -4,   /* neighbour pose / measurement not found */
  DPGO_B200_ERR_NUMERIC = -5    /* factorisation failed (matrix not PD) */
} dpgo_b200_error;

/* PGOAgentParameters fields set in src/PGOAgentROSNode.cpp:72-232. */
typedef struct {
  int d, r, num_robots;                       /* :72 (d must be 3, :63; r >= d, :67) */
  int method;                                 /* 0 RTR, 1 RGD (:82-93) */
  double rgd_stepsize;                        /* :96 */
  int rgd_use_preconditioner;                 /* :97 */
  int rtr_iterations, rtr_tcg_iterations;     /* :98-99 */
  double rtr_initial_radius, gradnorm_tol;    /* :100 */
  int acceleration, restart_interval;         /* :126-130 */
  int cost_type;                              /* 0 L2 ... 5 GNC_TLS (:178-188) */
  double gnc_barc, gnc_mu_step, gnc_init_mu;  /* :202-211 */
  int robust_opt_num_weight_updates, robust_opt_num_resets, robust_opt_inner_iters; /* :212-217 */
  double robust_opt_min_convergence_ratio;    /* :214 -- readyToTerminate needs this share of loop-closure weights settled at 0 or 1 */
  int max_num_iters;                          /* :226-231 */
  double rel_change_tol;                      /* :145 */
  double precond_lambda;                      /* Q + lambda I regularisation of the preconditioner */
} dpgo_b200_params;

/* mLocalOptResult, read at src/PGOAgentROS.cpp:169-172 */
typedef struct {
  int success;
  double f_init, f_opt, gradnorm_init, gradnorm_opt, relative_change;
  int rtr_outer_iters, tcg_iters, rtr_rejections;
} dpgo_b200_opt_result;

/* PGOAgentStatus (src/utils.cpp:262-281); state: 0 WAIT_FOR_DATA, 1 WAIT_FOR_INITIALIZATION, 2 INITIALIZED */
typedef struct {
  int agent_id, state, instance_number, iteration_number, ready_to_terminate;
  double relative_change;
} dpgo_b200_status;

typedef struct {
  int iterations;      /* global iterations executed by this call */
  int terminated;      /* leader's shouldTerminate() fired */
  int weight_updates;  /* GNC weight updates performed inside this call */
  float device_ms;     /* CUDA-event time of the persistent kernel launches */
  int kernel_launches; /* number of kernels this call launched */
  int stop_reason;     /* 0 ran max_iters, 1 terminated, 2 GNC weight update due (multi-GPU: the caller does it) */
} dpgo_b200_run_result;

/* ---- library ------------------------------------------------------------- */
const char *dpgo_b200_version(void);
const char *dpgo_b200_last_error(void);
int dpgo_b200_device_count(void);
/* number of kernels launched by this library since load (bench.py gpu_launches) */
long long dpgo_b200_kernel_launch_count(void);

/* ---- agent lifecycle: PGOAgent(ID, params), src/PGOAgentROS.cpp:26 ---------- */
int dpgo_b200_agent_create(int id, const dpgo_b200_params *params, int device, dpgo_b200_agent_t *out);
int dpgo_b200_agent_destroy(dpgo_b200_agent_t a);
int dpgo_b200_reset(dpgo_b200_agent_t a);                       /* PGOAgent::reset, :223 */

/* addMeasurement, :277 and :1307 (duplicates are ignored like hasMeasurement, :276) */
int dpgo_b200_add_measurements(dpgo_b200_agent_t a, int m, const int *r1, const int *p1, const int *r2,
                               const int *p2, const double *R, const double *t, const double *kappa,
                               const double *tau, const double *weight, const unsigned char *fixed);
int dpgo_b200_num_poses(dpgo_b200_agent_t a);                   /* num_poses(), :285 */
int dpgo_b200_iteration_number(dpgo_b200_agent_t a);            /* iteration_number(), :139 */
/* the RECOVER handler rewinds the protected member mIterationNumber (:1196); the shim forwards the new value */
int dpgo_b200_set_iteration_number(dpgo_b200_agent_t a, int iteration);
int dpgo_b200_num_neighbors(dpgo_b200_agent_t a);               /* getNeighbors(), :663 */
int dpgo_b200_get_neighbors(dpgo_b200_agent_t a, int *ids, int cap);
/* PoseGraph counters read at :343-345 */
int dpgo_b200_measurement_counts(dpgo_b200_agent_t a, int *odometry, int *private_lc, int *shared_lc);

int dpgo_b200_set_lifting_matrix(dpgo_b200_agent_t a, const double *Y);   /* :928 */
int dpgo_b200_get_lifting_matrix(dpgo_b200_agent_t a, double *Y);         /* :404 */
int dpgo_b200_initialize(dpgo_b200_agent_t a, const double *T_local_or_null);   /* :348 */
/* local_initialization_method "Chordal" (src/PGOAgentROSNode.cpp:106-112; demo default, launch/dpgo_demo.launch:9):
 * chordal relaxation of the rotations + linear translations over the robot's own edges, pose 0 = identity, solved
 * with the dense-inverse machinery on the device.  Leaves the agent in WAIT_FOR_INITIALIZATION like initialize(). */
int dpgo_b200_initialize_chordal(dpgo_b200_agent_t a);
/* the local-frame trajectory produced by initialize / initialize_chordal: n x 3 x 4 doubles, row-major per pose */
int dpgo_b200_get_local_trajectory(dpgo_b200_agent_t a, double *T_local);
int dpgo_b200_initialize_in_global_frame(dpgo_b200_agent_t a, const double *T_world_robot); /* :353,358 */

/* ---- the hot call: iterate(bool), :160 (true) and :1185 (false) -------------- */
int dpgo_b200_iterate(dpgo_b200_agent_t a, int do_optimization);
int dpgo_b200_get_opt_result(dpgo_b200_agent_t a, dpgo_b200_opt_result *out);   /* :169-172 */
/* same, but f_opt / gradnorm_opt of a stand-alone RGD iterate stay NaN when they have not been
 * evaluated yet (they cost a second gradient pass; the wrapper only prints them when verbose) */
int dpgo_b200_get_opt_result_lazy(dpgo_b200_agent_t a, dpgo_b200_opt_result *out);
int dpgo_b200_get_status(dpgo_b200_agent_t a, dpgo_b200_status *out);           /* getStatus, :616 */
int dpgo_b200_set_neighbor_status(dpgo_b200_agent_t a, const dpgo_b200_status *s);  /* :965 */
/* setRobotActive(id, active), :382 ... :1582: robots the leader has deactivated (disconnected, left the cluster) are
 * left out of shouldTerminate / shouldUpdateMeasurementWeights, and the shared loop closures with a deactivated
 * neighbour leave Q, G and the preconditioner (upstream's default; the wrapper's alternative is commented out, :151-156).
 * All robots are active after create / reset. */
int dpgo_b200_set_robot_active(dpgo_b200_agent_t a, int robot, int active);
int dpgo_b200_should_terminate(dpgo_b200_agent_t a);                            /* :208  (1/0, <0 error) */
int dpgo_b200_should_update_measurement_weights(dpgo_b200_agent_t a);           /* :210 */

/* which: 0 X, 1 Y (auxiliary), 2 V -- getX of the north star; host buffer r x 4n */
int dpgo_b200_get_x(dpgo_b200_agent_t a, int which, double *out);
int dpgo_b200_set_x(dpgo_b200_agent_t a, const double *X);
/* one pose (r x 4, column-major) of X / Y / V: getSharedPose, src/PGOAgentROS.cpp:424 */
int dpgo_b200_get_pose(dpgo_b200_agent_t a, int which, int index, double *out);

/* ---- public-pose exchange with HOST buffers (a9) ----------------------------
 * getSharedPoseDictWithNeighbor / getAuxSharedPoseDictWithNeighbor (:666-668)
 * and updateNeighborPoses / updateAuxNeighborPoses (:1276-1278).              */
int dpgo_b200_num_shared_poses(dpgo_b200_agent_t a, int neighbor);
int dpgo_b200_get_shared_pose_dict(dpgo_b200_agent_t a, int neighbor, int aux, int *frame_ids, double *poses,
                                   int cap, int *count);
int dpgo_b200_update_neighbor_poses(dpgo_b200_agent_t a, int neighbor, int aux, const int *frame_ids,
                                    const double *poses, int count);
/* Same exchange with raw DEVICE buffers (NCCL / NVLink peer transport replacing
 * the ROS MatrixMsg path): a packed outbox per neighbour, frame-id order of
 * dpgo_b200_get_shared_pose_dict, and the matching inbox on the receiver.     */
int dpgo_b200_outbox_device_ptr(dpgo_b200_agent_t a, int neighbor, int aux, void **dev_ptr, size_t *bytes);
int dpgo_b200_inbox_device_ptr(dpgo_b200_agent_t a, int neighbor, int aux, void **dev_ptr, size_t *bytes);
/* tell the agent that the inbox of `neighbor` was filled by an external transport */
int dpgo_b200_mark_inbox_updated(dpgo_b200_agent_t a, int neighbor, int aux);

/* ---- GNC-TLS (a8) -------------------------------------------------------------- */
int dpgo_b200_update_measurement_weights(dpgo_b200_agent_t a);                  /* :1218 */
int dpgo_b200_set_measurement_weight(dpgo_b200_agent_t a, int r1, int p1, int r2, int p2, double w,
                                     int fixed);                               /* :1341 */
int dpgo_b200_compute_measurement_residual(dpgo_b200_agent_t a, int r1, int p1, int r2, int p2,
                                           double *residual);                  /* :1049 */
double dpgo_b200_robust_weight(dpgo_b200_agent_t a, double residual);           /* mRobustCost.weight, :1050 */
int dpgo_b200_clear_data_matrices(dpgo_b200_agent_t a);                         /* :1351 */
/* loop-closure weights in insertion order: private LCs, then shared LCs */
int dpgo_b200_get_lc_weights(dpgo_b200_agent_t a, double *out, int cap);
/* mPoseGraph->sharedLoopClosures() as publishMeasurementWeights reads it (:721-754): ids, weight and
 * fixedWeight of every shared loop closure; returns the total count (arrays may be NULL / shorter) */
int dpgo_b200_get_shared_loop_closures(dpgo_b200_agent_t a, int *r1, int *p1, int *r2, int *p2, double *weight,
                                       unsigned char *fixed, int cap);
int dpgo_b200_weight_update_count(dpgo_b200_agent_t a);                         /* mWeightUpdateCount, :193 */

/* ---- problem-level evaluation on the device (parity hooks for a3/a5/a6) ------- */
int dpgo_b200_eval(dpgo_b200_agent_t a, const double *X, double *f, double *egrad, double *rgrad);
int dpgo_b200_hess(dpgo_b200_agent_t a, const double *X, const double *V, double *out);
int dpgo_b200_precond(dpgo_b200_agent_t a, const double *X, const double *V, double *out);
int dpgo_b200_manifold_project(int device, int r, int n, const double *M, double *out);
int dpgo_b200_tangent_project(int device, int r, int n, const double *X, const double *Z, double *out);
int dpgo_b200_retract(int device, int r, int n, const double *X, const double *xi, double *out);

/* ---- team: co-located agents, device-side exchange and schedule ---------------
 * Replays the wrapper's synchronous protocol (UPDATE token RoundRobin,
 * :464-472; non-selected robots iterate(false), :1185; leader decides
 * termination / weight update, :207-217) inside ONE persistent kernel.         */
int dpgo_b200_team_create(int device, dpgo_b200_team_t *out);
int dpgo_b200_team_destroy(dpgo_b200_team_t t);
int dpgo_b200_team_add_agent(dpgo_b200_team_t t, dpgo_b200_agent_t a);
int dpgo_b200_team_exchange_all(dpgo_b200_team_t t);   /* publishPublicPoses for every pair, :662-690 */
int dpgo_b200_team_run(dpgo_b200_team_t t, int max_iters, int stop_on_terminate, dpgo_b200_run_result *out);
/* One global iteration for a team that holds only PART of the robots (one team per GPU).
 * mode 0: whole iterate of every local agent (selected_robot optimises if it is local);
 * mode 1: Nesterov half -- Y of every local agent, X = Y for the non-selected, outboxes filled;
 * mode 2: the selected robot's local solve, to be called after its neighbours' poses of this
 *         iteration were delivered (the gate of src/PGOAgentROS.cpp:136-149).               */
int dpgo_b200_team_step(dpgo_b200_team_t t, int selected_robot, int mode);
double dpgo_b200_team_global_cost(dpgo_b200_team_t t, int *status);
/* schedule of team_run / team_fabric_run.  0 (default): the synchronous protocol -- one UPDATE token, RoundRobin
 * (src/PGOAgentROS.cpp:443-479, 1161-1189).  1: the asynchronous mode (asynchronous=true selects RGD and lets
 * every robot optimise on its own clock, src/PGOAgentROSNode.cpp:80-93, src/PGOAgentROS.cpp:119-127) as its
 * deterministic equal-rate / unit-delay schedule: in every iteration ALL robots take an RGD step against the
 * neighbour poses of the previous iteration, then all publish; no acceleration, no termination test.        */
int dpgo_b200_team_set_schedule(dpgo_b200_team_t t, int schedule);
/* tuning knob: CTAs of the persistent kernel (0 = one per SM) */
int dpgo_b200_team_set_grid(dpgo_b200_team_t t, int num_ctas);

/* ---- multi-GPU fabric: one team (process) per GPU, neighbour PublicPoses as raw device stores ----
 * Replaces the PublicPoses / MatrixMsg topics (msg/PublicPoses.msg:1-8, src/PGOAgentROS.cpp:662-690,
 * 1255-1284) and the iteration gate (:136-149) for robots that live on different GPUs of one node:
 * the inboxes of a team's agents sit in one cudaMalloc'ed window that the other ranks map with CUDA
 * IPC; the persistent kernels publish by storing into the neighbour's inbox over NVLink and keep in
 * step through flag words in the same windows.  Host-side sequence (every rank):
 *   fabric_init -> fabric_window (export) -> [all-gather handles + inbox offsets] -> fabric_import
 *   per peer -> fabric_route per (local robot, remote neighbour) -> team_exchange_all ->
 *   [host barrier] -> mark_inbox_updated -> fabric_run ...                                          */
int dpgo_b200_team_fabric_init(dpgo_b200_team_t t, int world, int rank);
/* base / size of this rank's window and its 64-byte CUDA IPC handle (ipc_handle may be NULL) */
int dpgo_b200_team_fabric_window(dpgo_b200_team_t t, void **base, size_t *bytes, void *ipc_handle_64);
/* map a peer's window: by IPC handle (other process) or by pointer (team of this process, tests) */
int dpgo_b200_team_fabric_import(dpgo_b200_team_t t, int peer_rank, const void *ipc_handle_64,
                                 void *same_process_base);
/* poses of local `robot` shared with remote `neighbor` go to byte offsets off_reg / off_aux of the
 * peer's window (= dpgo_b200_inbox_device_ptr(neighbor's agent, robot, aux) - its window base)      */
int dpgo_b200_team_fabric_route(dpgo_b200_team_t t, int robot, int neighbor, int peer_rank, size_t off_reg,
                                size_t off_aux);
/* up to max_iters (<= 65536) global iterations in ONE persistent launch per rank; called by every
 * rank at the same point of the schedule.  stop_reason 2: every rank must run the GNC weight update
 * (gnc_compute_weights -> carry shared-edge weights to the higher-ID owner's rank with
 * get_lc_weights / set_measurement_weight -> gnc_finish_update) and call fabric_run again.          */
int dpgo_b200_team_fabric_run(dpgo_b200_team_t t, int max_iters, int stop_on_terminate, dpgo_b200_run_result *out);
int dpgo_b200_team_fabric_set_timeout(dpgo_b200_team_t t, double seconds);
int dpgo_b200_team_fabric_close(dpgo_b200_team_t t);
/* UPDATE_WEIGHT (src/PGOAgentROS.cpp:1211-1233) in two halves: residuals + GNC-TLS weights of the
 * edges this team owns; then mu / counters / Q, G, preconditioner rebuild / re-publication.         */
int dpgo_b200_team_gnc_compute_weights(dpgo_b200_team_t t);
int dpgo_b200_team_gnc_finish_update(dpgo_b200_team_t t);

/* ---- host harness ------------------------------------------------------------------
 * ROS-free replay of PGOAgentROS's synchronous per-iteration call sequence on N
 * standalone agents, using ONLY the per-robot entry points above with HOST
 * buffers (iterate, getSharedPoseDict, updateNeighborPoses, getStatus,
 * shouldTerminate; src/PGOAgentROS.cpp:102-220, 1161-1189, 1255-1284), one OS
 * thread per robot.  Runs `steps` global iterations; *terminated_at = first step
 * (1-based) at which the leader's shouldTerminate() fired, or -1.               */
int dpgo_b200_sync_driver_run(dpgo_b200_agent_t *agents, int num_agents, int steps, int accelerated,
                              double *seconds, long long *payload_bytes, int *terminated_at);

/* The same replay with the robots spread over several processes of one node (one per GPU).  `shm` points to a
 * zero-initialised shared-memory segment of dpgo_b200_sync_driver_shm_bytes(num_robots, max_shared_poses) bytes
 * mapped by every process (with term = -1 written by its creator, see dpgo_ros_b200/dist.py); it holds the
 * mailboxes, the barrier and the status board that stand in for the TCPROS transport between the reference's
 * per-robot processes.  Every process passes its own robots; all pass the same steps / start_iter.            */
size_t dpgo_b200_sync_driver_shm_bytes(int num_robots, int max_shared_poses);
int dpgo_b200_sync_driver_run_shm(dpgo_b200_agent_t *agents, const int *robot_ids, int num_local, int num_robots,
                                  void *shm, int max_shared_poses, int steps, int accelerated, int start_iter,
                                  double *seconds, int *terminated_at);

/* ---- diagnostics ---------------------------------------------------------------------
 * Q as the DEVICE assembled it (per-edge accumulation kernel, dpgo_ros_b200/csrc/assemble.cu; the reference
 * assembles it in PoseGraph::quadraticMatrix after addMeasurement :277 / clearDataMatrices :1351): dense 4n x 4n
 * column-major, once from the block-CSR copy and once from the ELL + overflow copy.  Either pointer may be NULL. */
int dpgo_b200_debug_dense_q(dpgo_b200_agent_t a, double *Q_from_csr, double *Q_from_ell);
/* the LARGE-agent Riemannian-gradient kernel on its own (dpgo_ros_b200/csrc/edge_grad.cu; QuadraticProblem::f /
 * EucGrad + tangent projection straight from 128-byte edge records): f and rgrad at X (NULL: the agent's X) against
 * the current neighbour poses, the kernel's device time (first CTA start -> last CTA end, globaltimer ns) and the
 * CUDA-event time around the launch.  flush_l2: rewrite 512 MB first so that the graph comes from HBM. */
int dpgo_b200_debug_edge_grad(dpgo_b200_agent_t a, const double *X, int flush_l2, double *f, double *rgrad,
                              double *kernel_ns, double *event_ns);
/* the dense SPD inverse behind the preconditioner and the Chordal initialisation on its own
 * (dpgo_ros_b200/csrc/dense_inverse.cu; stands where the reference factors Q + lambda I with CHOLMOD,
 * QuadraticProblem's preconditioner / src/PGOAgentROS.cpp:1351 after every weight update): A is N x N column-major on the
 * host, N a multiple of 32, only its lower triangle is read; P receives A^-1 (both triangles).  device_ms: CUDA-event
 * time of the factorisation alone.  DPGO_B200_ERR_NUMERIC when A is not positive definite. */
int dpgo_b200_debug_spd_inverse(int device, int N, const double *A, double *P, double *device_ms);
/* wall-clock seconds spent inside each entry point of this library between two instants of
 * std::chrono::steady_clock (seconds since its epoch), one "name seconds calls" line per entry point; returns the
 * bytes the full report needs.  Lets a caller (the reference's wrapper in oracle/_ref) split the run time of a round
 * into library time and its own host code.  reset != 0 forgets the recorded calls afterwards.  The clock is OFF until
 * one call passes reset = 2 (which also clears it): entry points pay nothing for it otherwise. */
int dpgo_b200_debug_api_profile(double t_begin, double t_end, char *buf, int cap, int reset);

#ifdef __cplusplus
}
#endif
#endif /* DPGO_B200_H */
