/*
 * ttmpc.h -- C ABI of the B200-native batched truck-trailer NMPC solver.
 *
 * This is the drop-in boundary for ONE path of Avan1ko/car-trailer-mpc: the
 * per-timestep `controller.solve(initial_state, reference_states,
 * reference_inputs) -> (states, inputs)` call.  The reference has no FFI; every
 * entry point below replaces a piece of Python that sits in front of
 * `ca.nlpsol('solver','ipopt',...)`:
 *
 *   ttmpc_create            <- MPCTrackingControl.__init__ / _build_solver
 *                              (python-files/mpc_control.py:6-56), TruckTrailerNMPC.__init__
 *                              (python-files/mpc_control_nmpc.py:11-58): model constants,
 *                              Q/R, box bounds, Ipopt option preset.
 *   ttmpc_solve_batch       <- MPCTrackingControl.solve (mpc_control.py:67-110) and
 *                              TruckTrailerNMPC.solve (mpc_control_nmpc.py:90-113),
 *                              B independent problems per call instead of one.
 *   ttmpc_solve_batch_weighted <- MPCTrackingControlFuzzy.solve (mpc_control_fuzzy.py:121-167): same NLP
 *                              with per-solve diagonal weight scalings passed as parameters.
 *   ttmpc_solve_batch_shared<- the window extraction of the closed-loop drivers
 *                              (simulation.py:485-499, simulation_nmpc.py:193-204) fused
 *                              with the solve: every problem tracks the same trajectory.
 *   ttmpc_solve_batch_multi <- the same for F rigidly transformed copies of one trajectory: the batch
 *                              driver fed by compare_sweep.py / test_cases.json (scenario families).
 *   ttmpc_shift_warm_start  <- TruckTrailerNMPC._shift_solution (mpc_control_nmpc.py:69-88).
 *   ttmpc_plant_step        <- update()/f_dyn of the drivers (simulation.py:34-48,167-199,
 *                              simulation_nmpc.py:94-105) for on-device closed loops.
 *   ttmpc_episode_batch     <- the whole closed loop of simulation.py:484-560 /
 *                              simulation_nmpc.py:192-255 for B vehicles, on the device.
 *
 * Layouts are the reference's own:
 *   decision vector  z = [x_0;u_0;x_1;u_1;...;x_{N-1};u_{N-1};x_N], 8N+6 doubles
 *                       (trajectory_planning.py:38-60), x=(x,y,theta,psi,phi,v), u=(a,omega);
 *   reference window ref_states[N+1][6], ref_inputs[N][2]: stage-major, i.e. exactly the
 *                       parameter vector p of mpc_control.py:48-50 / :71-72 without x_init.
 *
 * All arithmetic is IEEE double.  No torch types, no C++ types: plain pointers and sizes.
 * Unless TTMPC_FLAG_HOST_POINTERS is set, every array pointer is a CUDA device pointer
 * owned by the caller (e.g. torch.Tensor.data_ptr()); the library only owns its private
 * scratch.  Calls are asynchronous on the supplied stream unless TTMPC_FLAG_SYNC is set;
 * consecutive calls through one handle are ordered on the device even when they are given
 * different streams (they share the handle's scratch).  Host-pointer calls are synchronous
 * unless TTMPC_FLAG_ASYNC_HOST is set (then ttmpc_sync() waits for the outputs); they run on
 * the handle's own streams and ignore the stream argument.  Every entry point leaves the
 * caller's current CUDA device unchanged.  A handle must not be used concurrently from
 * several threads; distinct handles are independent.
 *
 * Return value: 0 on success, negative TTMPC_E_* otherwise; never throws, never exits.
 */
#ifndef TTMPC_H
#define TTMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTMPC_NX 6
#define TTMPC_NU 2
#define TTMPC_MAX_HORIZON 256
#define TTMPC_MAX_OBSTACLES 16

/* error codes */
#define TTMPC_OK 0
#define TTMPC_E_INVAL (-22)   /* bad argument / config                         */
#define TTMPC_E_NOMEM (-12)   /* device scratch allocation failed              */
#define TTMPC_E_CUDA (-5)     /* CUDA runtime error, see ttmpc_last_error()    */
#define TTMPC_E_NODEV (-19)   /* no usable CUDA device (there is NO CPU fallback) */

/* per-problem status codes (status_out) */
#define TTMPC_ST_CONVERGED 0      /* Ipopt "Solve_Succeeded": scaled KKT error <= tol            */
#define TTMPC_ST_ACCEPTABLE 1     /* Ipopt "Solved_To_Acceptable_Level"                          */
#define TTMPC_ST_MAX_ITER 2       /* iteration limit; z_out holds the last iterate               */
#define TTMPC_ST_LINESEARCH 3     /* step rejected repeatedly (Ipopt would enter restoration)    */
#define TTMPC_ST_NUMERIC 4        /* NaN/Inf encountered, z_out holds the last finite iterate    */
#define TTMPC_ST_INFEASIBLE_X0 5  /* x_init violates a state bound (reference NLP infeasible) and the
                                     solve with x_0 as data did not converge within 30 iterations   */

/* flags */
#define TTMPC_FLAG_HOST_POINTERS 0x1u /* array arguments are host pointers (B=1 shim path)      */
#define TTMPC_FLAG_SYNC 0x2u          /* cudaStreamSynchronize before returning                 */
#define TTMPC_FLAG_ASYNC_HOST 0x8u    /* with HOST_POINTERS: a solve returns once its copy-in, kernels and
                                         copy-out are queued on the handle's three internal streams; up to three
                                         solves are in flight (a call first waits for the third-last one), so the
                                         copy-in of batch t+1 and the copy-out of batch t-1 overlap the solve of
                                         batch t.  Host buffers must stay untouched until ttmpc_sync() and should
                                         be page-locked (pageable memory makes the copies synchronous).           */

typedef struct ttmpc_config {
  int32_t horizon;          /* N, params['horizon'] (simulation.py:390), 1..TTMPC_MAX_HORIZON   */
  int32_t max_iter;         /* Ipopt max_iter (mpc_control.py:35 = 5000, nmpc :38 = 2000)       */
  int32_t acceptable_iter;  /* Ipopt acceptable_iter (default 15, nmpc :42 = 5)                 */
  uint32_t flags;           /* TTMPC_FLAG_*                                                     */
  double dt;                /* params['dt']  (simulation.py:389)                                */
  double L1, L2, M;         /* params['L1'],['L2'],['M'] (simulation.py:391-393)                */
  double Q[36];             /* row-major 6x6 state weight (simulation.py:400-406)               */
  double R[4];              /* row-major 2x2 input weight (simulation.py:407-409)               */
  double x_lb[6], x_ub[6];  /* state_bound (simulation.py:411-412); +-INFINITY = unbounded       */
  double u_lb[2], u_ub[2];  /* input_bound (simulation.py:413-414)                              */
  double tol;               /* Ipopt tol (default 1e-8, nmpc 1e-3)                              */
  double acceptable_tol;    /* Ipopt acceptable_tol (default 1e-6, nmpc 1e-2)                   */
  double mu_init;           /* Ipopt mu_init (default 0.1)                                      */
} ttmpc_config;

typedef struct ttmpc_handle ttmpc_handle;

/* Fill cfg with the MPCTrackingControl preset of simulation.py:388-414 (N as given). */
void ttmpc_default_config(ttmpc_config* cfg, int32_t horizon);

/* Create a solver bound to CUDA device `device`.  Fails with TTMPC_E_NODEV when no CUDA
 * device is usable -- there is no CPU path in this library. */
int ttmpc_create(const ttmpc_config* cfg, int device, ttmpc_handle** out);
int ttmpc_destroy(ttmpc_handle* h);
const char* ttmpc_last_error(const ttmpc_handle* h);
const char* ttmpc_version(void);

/* Solve B independent NLPs.
 *   x_init     [B][6]
 *   ref_states [B][N+1][6]      ref_inputs [B][N][2]
 *   z_warm     [B][8N+6] or NULL (NULL = cold start at the reference window, mpc_control.py:58-65)
 *   z_out      [B][8N+6] or NULL
 *   u0_out     [B][2]    or NULL  first control (what the caller applies, simulation.py:525)
 *   obj_out    [B]       or NULL  objective J(z*)
 *   kkt_out    [B][3]    or NULL  unscaled (dual infeasibility, constraint violation, complementarity)
 *   iters_out  [B]       or NULL  interior-point iterations used
 *   status_out [B]       or NULL  TTMPC_ST_*
 *   cuda_stream: cudaStream_t (NULL = legacy default stream)
 */
int ttmpc_solve_batch(ttmpc_handle* h, int64_t B, const double* x_init, const double* ref_states,
                      const double* ref_inputs, const double* z_warm, double* z_out,
                      double* u0_out, double* obj_out, double* kkt_out, int32_t* iters_out,
                      int32_t* status_out, void* cuda_stream);

/* Same as ttmpc_solve_batch with per-problem cost-weight scalings q_weights [B][6], r_weights [B][2]:
 * Q_w = diag(q) Q diag(q), R_w = diag(r) R diag(r), terminal weight Q_w -- the parametric weights of
 * MPCTrackingControlFuzzy (mpc_control_fuzzy.py:21-31,51-60).  Needs diagonal Q and R in the config. */
int ttmpc_solve_batch_weighted(ttmpc_handle* h, int64_t B, const double* x_init, const double* ref_states,
                               const double* ref_inputs, const double* q_weights, const double* r_weights,
                               const double* z_warm, double* z_out, double* u0_out, double* obj_out,
                               double* kkt_out, int32_t* iters_out, int32_t* status_out, void* cuda_stream);

/* Same, but every problem tracks one shared trajectory traj_states[T+1][6], traj_inputs[T][2];
 * problem i uses the window starting at k_index[i] with the three padding regimes of
 * simulation.py:485-499 (full slice / pad with last state + last input / past the end =
 * last state + zero input). */
int ttmpc_solve_batch_shared(ttmpc_handle* h, int64_t B, const double* x_init,
                             const int32_t* k_index, const double* traj_states,
                             const double* traj_inputs, int32_t T, const double* z_warm,
                             double* z_out, double* u0_out, double* obj_out, double* kkt_out,
                             int32_t* iters_out, int32_t* status_out, void* cuda_stream);

/* Shared-trajectory solve over F trajectories -- the batch driver's contract (compare_sweep.py shape: the scenario
 * families of test_cases.json enter as rigid transforms of one reference trajectory): traj_states [F][T+1][6],
 * traj_inputs [F][T][2]; problem i tracks trajectory traj_index[i] from window start k_index[i] (padding rules of
 * simulation.py:485-499).  Per problem 56 B go in (x_init, k_index, traj_index) and, with z_out = obj_out = kkt_out =
 * NULL, 24 B come out (u0, iterations, status). */
int ttmpc_solve_batch_multi(ttmpc_handle* h, int64_t B, const double* x_init, const int32_t* k_index,
                            const int32_t* traj_index, const double* traj_states, const double* traj_inputs,
                            int32_t F, int32_t T, const double* z_warm, double* z_out, double* u0_out,
                            double* obj_out, double* kkt_out, int32_t* iters_out, int32_t* status_out,
                            void* cuda_stream);

/* Wait until every host-pointer solve queued through this handle (TTMPC_FLAG_ASYNC_HOST) has delivered its outputs. */
int ttmpc_sync(ttmpc_handle* h);

/* Device-side duration (CUDA events on the handle's copy streams) of the last burst of host-pointer solves that
 * ttmpc_sync() -- or a synchronous host-pointer call -- waited for: from the first copy-in to the last copy-out, ms. */
double ttmpc_host_pipeline_ms(const ttmpc_handle* h);

/* Obstacle set of the obstacle-aware controller MPCTrackingControlObs (mpc_control_obs.py:8-30): axis-aligned
 * rectangles {centre x, centre y, width, height} as produced by get_obstacles.py:5-33, body widths W1 / W2
 * (params['W1'], params['W2'], simulation.py:393; the body lengths are L1, L2 of the config) and the safety distance
 * d_min (mpc_control_obs.py:67: 0.2). */
/* A line search that runs out of backtracking steps is where Ipopt enters its feasibility-restoration phase
 * (mpc_control_obs.py:193-197 leaves Ipopt's defaults on).  Default here: the solve recovers with a fresh interior-point
 * start at the current primal iterate, up to 3 times per solve.  With this flag the solve takes the
 * shortest trial step instead and reports TTMPC_ST_LINESEARCH after three consecutive failures (round-1 behaviour). */
#define TTMPC_OBCA_NO_RECOVERY 1
/* Opt-in, NOT the reference's behaviour: start the OBCA duals at the multipliers of the distance problem between each
 * body and obstacle for the pose of the starting trajectory (the closed form the recovery uses)
 * instead of the reference's constants mu = 100, lam = (100,105,110,115) (mpc_control_obs.py:226-237), whose rows start
 * 5e6 away from feasibility.  Same NLP, same tolerances, same states and inputs at convergence; a third of the
 * iterations.  Off by default because the reference's starting point is part of what its iterate path looks like. */
#define TTMPC_OBCA_GEOMETRIC_START 2
typedef struct ttmpc_obstacles {
  int32_t count; /* 1..TTMPC_MAX_OBSTACLES */
  int32_t flags; /* 0, or TTMPC_OBCA_NO_RECOVERY | TTMPC_OBCA_GEOMETRIC_START */
  double rect[TTMPC_MAX_OBSTACLES][4];
  double W1, W2, d_min;
} ttmpc_obstacles;

/* Solve B NLPs of MPCTrackingControlObs.solve (mpc_control_obs.py:282-322): tracking cost and dynamics of
 * ttmpc_solve_batch plus, for every stage, obstacle and body, the OBCA dual variables mu, lam >= 0 and the three
 * collision row groups of _collision_constraints (mpc_control_obs.py:65-139).  Always a cold start at the reference
 * window with mu = 100, lam = (100,105,110,115) (_get_initial_guess, :216-239) unless TTMPC_OBCA_GEOMETRIC_START is set.
 * Three kernels, chosen per call (ttmpc_kernel_name): one thread-block cluster per problem while every problem gets
 * its own cluster (the single solve of the shim: 8.7 ms at horizon 50 with 11 obstacles), one CTA per problem otherwise;
 * one warp per problem only on request (environment).  Outputs as ttmpc_solve_batch; z_out
 * holds the states and inputs only ([B][8N+6]) -- what _split_decision_variables (:241-281) hands back to the
 * caller; the OBCA duals are internal.  The second form windows one shared trajectory (simulation.py:485-499). */
int ttmpc_obca_solve_batch(ttmpc_handle* h, const ttmpc_obstacles* obstacles, int64_t B, const double* x_init,
                           const double* ref_states, const double* ref_inputs, double* z_out, double* u0_out,
                           double* obj_out, double* kkt_out, int32_t* iters_out, int32_t* status_out,
                           void* cuda_stream);
int ttmpc_obca_solve_batch_shared(ttmpc_handle* h, const ttmpc_obstacles* obstacles, int64_t B,
                                  const double* x_init, const int32_t* k_index, const double* traj_states,
                                  const double* traj_inputs, int32_t T, double* z_out, double* u0_out,
                                  double* obj_out, double* kkt_out, int32_t* iters_out, int32_t* status_out,
                                  void* cuda_stream);

/* The offline planner TrajectoryOptimization.plan(initial_state, goal_state) (trajectory_optimization.py:311-331; driver:
 * trajectory_animation.py:41-110 with horizon 200, dt 0.1, the 11 rectangles) for B start states towards one goal: the
 * obstacle-aware NLP above with
 *   cost   sum_{k<N} u_k'R u_k + (x_k - goal)'Q(x_k - goal)  +  (x_N - goal)' (terminal_weight Q) (x_N - goal)   (:175-183),
 *   final-state constraint |x_N - goal| <= terminal_box                                                 (:168-173),
 *   starting point z_guess [B][8N+6] = the caller's initial trajectory (states and inputs in the z layout; :227-274
 *   builds it from the Hybrid-A* waypoints, :208-225 as a straight line) or NULL = every state at the goal, zero inputs;
 *   the OBCA duals start at the reference's constants either way (or, with TTMPC_OBCA_GEOMETRIC_START, at the distance
 *   problems' multipliers for the poses of that trajectory).
 * goal: host pointer to 6 doubles (configuration, like the obstacle set).  The final-state constraint is held as bounds
 * of the terminal stage (the reference's range row has a slack that equals x_N - goal).  Outputs as
 * ttmpc_obca_solve_batch.  One problem runs on one CTA (N = 200, 11 obstacles: 37 k variables). */
int ttmpc_plan_batch(ttmpc_handle* h, const ttmpc_obstacles* obstacles, int64_t B, const double* x_init, const double* goal,
                     double terminal_weight, double terminal_box, const double* z_guess, double* z_out, double* u0_out,
                     double* obj_out, double* kkt_out, int32_t* iters_out, int32_t* status_out, void* cuda_stream);

/* Warm-start shift of B decision vectors (device pointers unless HOST flag in cfg):
 * z_shift[i] = shift(z[i]) as TruckTrailerNMPC._shift_solution. `mode` 0 = intended shift
 * (last stage repeated), 1 = the reference's mis-sliced tail. */
int ttmpc_shift_warm_start(ttmpc_handle* h, int64_t B, const double* z, double* z_shift,
                           int32_t mode, void* cuda_stream);

/* One explicit-Euler plant step for B states, q_next = update(q, u) of simulation.py:167-199.
 * disturb = NULL: nominal plant.  Otherwise disturb = {friction_coeff, slippage_coeff,
 * lateral_slip_gain, slip_angle_max}; additive noise [B][6] (may be NULL) is added as
 * q_next += noise_scale*noise (simulation_nmpc.py:100 uses noise_scale = dt). */
int ttmpc_plant_step(ttmpc_handle* h, int64_t B, const double* q, const double* u,
                     const double* disturb, const double* noise, double noise_scale,
                     double* q_next, void* cuda_stream);

/* B closed-loop episodes entirely on the device: the `while t <= T_sim` loop of simulation.py:484-560 (variant 0:
 * measurement noise on the state handed to the controller, simulation.py:513-517) or simulation_nmpc.py:192-255
 * (variant 1: plant noise*dt, zero control on a failed solve, and -- simulation_nmpc.py:212-216 -- the run of a vehicle
 * stops after more than 20 consecutive failed solves: it stays where it is for the remaining steps; unlike
 * TruckTrailerNMPC every solve is a cold start at the window) for B independent vehicles tracking one trajectory.
 * Per control step s the window starts at k_seq[s] (the float-accumulated floor(t/dt) sequence, built on the host);
 * every solve is a cold start at the window (mpc_control.py:58-65).  disturb = HOST array {friction_coeff,
 * slippage_coeff, lateral_slip_gain, slip_angle_max, process_noise_std} (DISTURBANCE_PARAMS, simulation.py:26-32) or
 * NULL for the nominal plant; noise is counter-based on (seed, step, ids[i]) so results do not depend on sharding.
 * metrics_out [B][8] = {final distance error, |heading error|, |hitch error| (simulation.py:574-580), max|psi|,
 * jackknife flag (|psi| > pi/3 + 1e-6 at any step), failed solves, mean iterations, RMS position tracking error};
 * final_state_out [B][6] optional.  Device pointers only. */
int ttmpc_episode_batch(ttmpc_handle* h, int64_t B, const double* x0, const int64_t* ids,
                        const double* traj_states, const double* traj_inputs, int32_t T,
                        const int32_t* k_seq, int32_t steps, const double* disturb, int32_t variant,
                        uint64_t seed, double* metrics_out, double* final_state_out, void* cuda_stream);

/* Which kernel flavour ran the last plain solve of this handle: 8, 16 or 32 = ttmpc_team_kernel with that many lanes
 * of a warp per problem (iterate resident in shared memory); 0 = ttmpc_solve_kernel (one lane per problem, per-problem
 * weights).  The library picks by horizon and batch size; TTMPC_KERNEL=lane|team and TTMPC_TEAM_LANES=8|16|32 override. */
int32_t ttmpc_last_solve_lanes(const ttmpc_handle* h);

/* Number of kernel launches issued through this handle so far (bench accounting). */
int64_t ttmpc_launch_count(const ttmpc_handle* h);

/* Name + accumulated launch count of kernel i (0-based); returns NULL past the end. */
const char* ttmpc_kernel_name(const ttmpc_handle* h, int32_t i, int64_t* launches);

/* Measure the non-tensor FP64 FMA peak of the device (independent DFMA chains on every SM),
 * returns GFLOP/s (2 flop per FMA) or a negative error. Used as the roofline denominator. */
double ttmpc_measure_fp64_peak(ttmpc_handle* h, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* TTMPC_H */
