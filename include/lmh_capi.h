/* lmh_capi.h -- C ABI of the B200-native batched NAO controller-and-simulator (liblmh_b200.so).
 *
 * The reference (Ema158/linearMpcHumanoid) has no FFI layer: its boundary is the C++ class surface
 * that apps/offline/main.cpp consumes.  This header is the thin C layer underneath our C++ facade
 * classes (include/linearMpcHumanoid_b200/facade.hpp) and the Python binding
 * (linearmpchumanoid_b200/api.py).  Each entry point names the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes; every array is FP64 unless its name says otherwise;
 * per-environment arrays are environment-major ([env][field], one env's record contiguous: a warp
 * loads its environment with coalesced 128-byte lines); all pointers are HOST pointers unless the
 * function name ends in _dev; every call returns 0 on success or a negative LMH_E* code, and
 * lmh_last_error() holds the message.  A handle owns one CUDA device, one stream and all device
 * buffers; there are no hidden globals.  There is no CPU fallback: without a usable CUDA device
 * lmh_create fails with LMH_ENODEVICE.
 */
#ifndef LMH_CAPI_H
#define LMH_CAPI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LMH_OK 0
#define LMH_EINVAL (-1)
#define LMH_ENODEVICE (-2)
#define LMH_ECUDA (-3)
#define LMH_ESTATE (-4)

#define LMH_NQ 30      /* Robot.hpp:11  NUM_JOINTS        */
#define LMH_NJ 24      /* Robot.hpp:12  NUM_ACTUAL_JOINTS */
#define LMH_NX 60      /* state = [q; v], controller.hpp:20-31 */

typedef struct LmhBatch LmhBatch;

/* Scenario literals of apps/offline/main.cpp:13-56 plus launch geometry. */
typedef struct {
    double dt;             /* main.cpp:14 */
    double sim_time;       /* main.cpp:13 */
    double time_horizon;   /* main.cpp:38 */
    double com[3];         /* main.cpp:31 */
    double rf[6], lf[6];   /* main.cpp:24-27 */
    double mu;             /* controller.hpp:81 */
    double mpc_alpha, mpc_beta;   /* mpcLinearPendulum.hpp:47-48 */
    int32_t qp_max_iter;   /* controller.cpp:438 (nWSR) */
    int32_t threads_per_env;      /* 32, 64 or 128; 0 = default */
    int32_t device;        /* CUDA device ordinal */
    int32_t log_steps;     /* capacity of the per-step log (0 = no log) */
    int32_t envs_per_block;/* one-warp environments per thread block of the fused rollout (threads_per_env = 32 only):
                              0 = as many as the shared memory holds (the default: the throughput kernel, results
                              independent of batch size and sharding bit for bit), 1..that many = fewer;
                              -1 = LATENCY MODE for a handful of environments: one environment per block plus two helper
                              warps (RNEA, CRBA and the sole Jacobians + MPC / PD references side by side): 46 instead of 62 us per RK4 step
                              at one environment; agrees with the default to rounding, not bit for bit */
    int32_t placement;     /* 0 = lmh_set_pushes places environments that are pushed at the same step in the same
                              block (default; results do not depend on it), 1 = keep the index order */
    double mpc_zcom;       /* Mpc3dLip's zCoM argument (mpcLinearPendulum.hpp:13-16); 0 = the CoM height after the inverse
                              kinematics of the initial posture, as main.cpp:39 passes it */
} LmhConfig;

/* Everything one dynamics() evaluation (apps/offline/main.cpp:91-122) produces, row-major. */
typedef struct {
    double M[900], C[30], Cg[30], AG[180], AGpqp[6], Jpqp[12], JFeet[360];      /* Dynamics.hpp:13-20, invKinematics.hpp:31-33 */
    double CoM[3], comVel[3], comAngMom[3], xRef[3], yRef[3];                    /* Robot.hpp:39-43, mpcLinearPendulum.hpp:22-23 */
    double qppRef[30], hGpRef[6], footAccRef[12];                                /* controller.cpp:296-386 */
    double Ha[900], ga[30];                                                      /* controller.cpp:94-132, joint block */
    double a[30], f[12], c[32], lam[6];                                          /* QP solution (controller.cpp:477) */
    double qpp[30], tau[24];                                                     /* WBCOutput, controller.hpp:42-48 */
    double T7[12], T14[12];                                                      /* sole poses: R row-major (9), p (3) */
    double xdot[60];
    int32_t k_index;       /* (int)(t/dt), mpcLinearPendulum.cpp:92 */
    int32_t qp_status;     /* 0 ok, 1 iteration limit */
    int32_t qp_iters;
    int32_t phase;         /* walking extension: min((int)(t/phase_T), n_phases-1) */
} LmhStageDump;

/* One record per environment and step when logging is on. */
typedef struct {
    double com_k4[3];          /* robot.getCoM() as printed by main.cpp:86 (k4 stage state) */
    double rfoot[3], lfoot[3]; /* sole positions at the same stage */
    double tau_k1[24];         /* torques of the k1 stage */
    int32_t k_index[4];
    int32_t qp_iters, qp_fail;
    int32_t phase[4];          /* walking extension: phase index of each of the four stages (0 in the stand task) */
    int32_t contact[4];        /* contact mask of each stage: bit 0 right foot, bit 1 left foot (3 in the stand task) */
} LmhStepLog;

/* EXTENSION (the reference has no notion of falling): the health rule of a closed-loop step.  A step is COMMITTED
 * (state, clock, statistics) only if, with x+ the state the RK4 combination would produce,
 *   max|tau(k1)| < LMH_TAU_LIMIT,  |CoM_xy(k4) - zmp_ref(k_index(k4))| < LMH_COM_ERR_LIMIT,
 *   every |x+_i| < LMH_STATE_LIMIT (NaN fails),  x+[2] (torso height) > LMH_TORSO_MIN.
 * Otherwise the environment is FALLEN: state and clock keep the values of the last healthy step, its statistics are
 * closed (they never see the failing step), alive = 0, and nothing more is evaluated for it: one bad environment
 * never takes the batch down and never pollutes the per-environment statistics. */
#define LMH_TAU_LIMIT 100.0
#define LMH_COM_ERR_LIMIT 0.5
#define LMH_STATE_LIMIT 1e6
#define LMH_TORSO_MIN 0.15

/* Per-environment rollout statistics (lmh_get_stats): */
typedef struct {
    double max_abs_tau;        /* over the k1 stages of the committed steps */
    double max_com_err;        /* max |CoM_xy(k4) - zmp_ref_xy| over the committed steps */
    double max_abs_qpp;
    int32_t qp_iters, qp_fail; /* totals over all stages of the committed steps */
    int32_t steps, alive;      /* steps committed; alive = 0 once a step failed the health rule above */
    int32_t qp_fallback;       /* stages whose all-free QP solution was infeasible (contact-limited: the general
                                  active-set iteration ran) */
    int32_t pad;
} LmhStats;

const char* lmh_last_error(void);
void lmh_default_config(LmhConfig* cfg);

/* main.cpp:16-58: Robot, Kinematics::compute (IK), Mpc3dLip, ZMP, footCoeffTrajectory, Controller.
 * All n_env environments start at the IK posture with zero velocity at t = 0. */
int32_t lmh_create(const LmhConfig* cfg, int32_t n_env, LmhBatch** out);
void lmh_destroy(LmhBatch* h);
int32_t lmh_n_env(const LmhBatch* h);

/* Setup results: Robot::getJoints after IK, Mpc3dLip::getZCom, Robot::getMass. */
int32_t lmh_get_nominal(const LmhBatch* h, double* q0_30, double* zcom, double* mass, int32_t* ik_iterations);
/* u0 = -(rPx . [pos,vel]) + sum_i r[i] zmp[k+i]  (condensed form of mpcLinearPendulum.cpp:89-101) */
int32_t lmh_get_mpc_gain(const LmhBatch* h, double* r, int32_t r_cap, int32_t* n1, double* rPx2);
/* Robot::getLinks after the constructor: mass[28], com[28][3], spatial inertia[28][36]. */
int32_t lmh_get_link_tables(const LmhBatch* h, double* mass28, double* com28x3, double* isp28x36);

/* The same table with the 3 x 3 inertia of every link about its centre of mass, in the joint frame (linkInertia::inertia
 * after the Robot constructor, Robot.cpp:14-22): inertia [28][9] row-major. */
int32_t lmh_get_link_inertia(const LmhBatch* h, double* inertia28x9);
/* createNaoParameters (src/robotParameters.cpp): the raw table -- mass [28], com [28][3], inertia [28][9] in frames
 * aligned with the world at q = 0.  Host only, no handle. */
int32_t lmh_get_nao_parameters(double* mass28, double* com28x3, double* inertia28x9);
/* matTrans (src/Robot.cpp:176-223): the 25 local modified-DH transforms for given theta [25] -> T [25][16] (4 x 4 row-major).  Host only. */
int32_t lmh_local_transforms(const double* theta25, double* T25x16);
/* findPolyCoeff (src/generalizedFunctions.cpp:103-163): polynomial (ascending powers) through n_pos positions, n_vel
 * velocities and n_acc accelerations given as (time, value) rows [n][2]; coeff [n_pos + n_vel + n_acc].  Host only. */
int32_t lmh_poly_fit(int32_t n_pos, const double* pos, int32_t n_vel, const double* vel, int32_t n_acc, const double* acc, double* coeff);

/* Plans replace the ZMP object and the two coefficient vectors handed to the Controller
 * (controller.hpp:52-57).  zmp_x/zmp_y: [n_plans][zmp_len]; foot_coef: [n_plans][2][20] =
 * x(6) y(6) z(8) per foot, ascending powers.  plan_ids (or NULL = 0) selects one per env. */
int32_t lmh_set_plans(LmhBatch* h, int32_t n_plans, int32_t zmp_len, const double* zmp_x, const double* zmp_y,
                      const double* foot_coef, const int32_t* plan_ids);
/* EXTENSION (walking; upstream declares walkZMP but never defines it, zmpGeneration.hpp:15-22): plans with phases.
 * Phase p = min((int)(t/phase_T), n_phases-1) selects foot_coef [n_plans][n_phases][2][20] (polynomials in the
 * phase-local time t - p*phase_T) and contact [n_plans][n_phases] (bit 0 right foot, bit 1 left foot may carry
 * force; a swing foot's 16 contact coefficients are held at zero).  lmh_set_plans == one phase, both feet. */
int32_t lmh_set_walk_plans(LmhBatch* h, int32_t n_plans, int32_t zmp_len, const double* zmp_x, const double* zmp_y,
                           int32_t n_phases, double phase_T, const double* foot_coef, const int32_t* contact,
                           const int32_t* plan_ids);
int32_t lmh_get_plan(const LmhBatch* h, int32_t plan, double* zmp_x, double* zmp_y, int32_t cap, int32_t* zmp_len, double* foot_coef40);

/* State: x [n_env][60] = [q; v]; v_held [n_env][30] = velocity the Robot object still holds from the
 * previous stage evaluation (Robot::v_, refreshed late in Controller::standStep, controller.cpp:48-79);
 * t [n_env].  NULL leaves a field unchanged (set) or unread (get).  lmh_set_state only ENQUEUES the copies on the
 * handle's stream (so that pinned host buffers overlap with earlier work): the caller's buffers must stay valid and
 * unmodified until the next lmh_sync or any lmh_get_* call, all of which synchronise.  lmh_get_state returns after
 * the data has arrived. */
int32_t lmh_set_state(LmhBatch* h, const double* x, const double* v_held, const double* t);
int32_t lmh_get_state(LmhBatch* h, double* x, double* v_held, double* t);
int32_t lmh_reset(LmhBatch* h);
/* Velocity impulse added to v[0:6] of env e at the START of step push_step[e] (counted from the last
 * reset / lmh_set_step_counter); push_step < 0 = none.  Extension: the reference has no pushes.
 * Side effect on placement only: environments are assigned to thread blocks in the order of their push step, so that
 * the environments of a block need their extra QP iterations at the same time; results do not depend on it. */
int32_t lmh_set_pushes(LmhBatch* h, const int32_t* push_step, const double* push_dv6);
int32_t lmh_set_step_counter(LmhBatch* h, int32_t step);

/* One dynamics(x,t) evaluation of every env at its current state, advancing only v_held (as the
 * reference's call does).  dumps: [n_env] or NULL. */
int32_t lmh_stage(LmhBatch* h, LmhStageDump* dumps);
/* n RK4 steps (rk4.hpp:5-18 + Clock::step) of every env, fused on the device; asynchronous. */
int32_t lmh_step(LmhBatch* h, int32_t n_steps);
/* same, bracketed by CUDA events on the handle's stream; returns device milliseconds */
int32_t lmh_step_timed(LmhBatch* h, int32_t n_steps, double* elapsed_ms);
int32_t lmh_sync(LmhBatch* h);

int32_t lmh_get_stats(LmhBatch* h, LmhStats* stats);
int32_t lmh_get_log(LmhBatch* h, int32_t first_step, int32_t n_steps, LmhStepLog* out); /* [n_steps][n_env] */
/* CoM of every env at the last evaluated stage: [n_env][3] */
int32_t lmh_get_com(LmhBatch* h, double* com);

/* Device-resident access for callers that keep data in HBM (torch tensors): raw device pointers to
 * x [n_env][60], v_held [n_env][30], t [n_env], and the stream (cudaStream_t) the handle launches on. */
int32_t lmh_state_dev(LmhBatch* h, double** x_dev, double** v_held_dev, double** t_dev, void** stream);
int32_t lmh_stats_dev(LmhBatch* h, LmhStats** stats_dev);
/* Device-to-device restore of the state from caller-owned device buffers (NULL = leave), on the handle's stream. */
int32_t lmh_set_state_dev(LmhBatch* h, const double* x_dev, const double* v_held_dev, const double* t_dev);
/* Zero the statistics and the step counter (pushes fire relative to it) without touching the state. */
int32_t lmh_reset_stats(LmhBatch* h);
/* CUDA-event stopwatch on the handle's stream (torch.cuda.Event only sees torch's current stream). */
int32_t lmh_timer_start(LmhBatch* h);
int32_t lmh_timer_stop(LmhBatch* h, double* elapsed_ms);

/* Config 5 micro-benchmark on the CURRENT (frozen) states: mode 0 = `repeats` cold-started WBC QP solves per
 * environment (data + factorisation + active set, i.e. what Controller::solveQP + qpOASES do per call,
 * controller.cpp:388-479); mode 1 = `repeats` LIPM-MPC evaluations (mpcLinearPendulum.cpp:78-133).  a_f_out
 * (or NULL): [n_env][42] = joint accelerations a (30, base twist in the torso frame) and contact wrenches f (12) of
 * the last solve.  elapsed_ms: device time of the launch (includes one evaluation of the model phases). */
int32_t lmh_qp_bench(LmhBatch* h, int32_t mode, int32_t repeats, double* a_f_out, double* elapsed_ms);

/* Kinematics::desiredOperationalState + Kinematics::compute (src/invKinematics.cpp:11-52: Newton inverse kinematics
 * on [sole poses, arm/head joints, base attitude, CoM]) for n targets at once on the device, one warp per target.
 * com3 [n][3], rf6 / lf6 [n][6] = sole position + Euler angles, as main.cpp:24-35 passes them; the arm/head joints keep
 * the batch's nominal posture and the attitude target is zero, as upstream.  q_out [n][30]; iters / residual (max
 * |operational error|) / status (0 converged to 1e-10, 1 iteration cap of 200, 2 singular Jacobian) [n]; any output may
 * be NULL.  apply_to_state != 0 (needs n == n_env): every environment restarts from its own posture with zero
 * velocity at t = 0 without a host round trip of the state (statistics and step counter are cleared).  The MPC
 * keeps the batch's CoM height (Mpc3dLip::getZCom of the nominal scenario). */
int32_t lmh_ik_batch(LmhBatch* h, int32_t n, const double* com3, const double* rf6, const double* lf6, double* q_out,
                     int32_t* iters, double* residual, int32_t* status, int32_t apply_to_state);

/* General form of the same: the full operational target desop [n][30] = [right sole position + Euler angles (6), left
 * sole (6), arm and head joints (12), base attitude (3), CoM (3)] as Kinematics::desiredOperationalState returns it
 * (invKinematics.cpp:11-25), and the posture the Newton iteration starts from, q_start [n][30] (NULL = the batch's
 * nominal posture for every target; Kinematics::compute starts from the Robot's current joints, invKinematics.cpp:27-33). */
int32_t lmh_ik_solve(LmhBatch* h, int32_t n, const double* q_start, const double* desop, double* q_out, int32_t* iters,
                     double* residual, int32_t* status);

/* Robot::updateState (src/Robot.cpp:264-269: forwardKinematics, computeCoM, parentTransMatrix, allVelocityMatrices) for n
 * postures on the device.  q [n][30] -> T [n][28][16] world transforms (4 x 4, row-major), X [n][28][36] motion
 * transforms of every frame with respect to its parent (6 x 6, row-major; X[0] = velocityMatrix(T[0])), com [n][3].
 * Any output may be NULL.  Frame 27 (the massless extra head frame, Robot.cpp:84,153-158) is completed on the host. */
int32_t lmh_kinematics(LmhBatch* h, int32_t n, const double* q, double* T, double* X, double* com);

/* Mpc3dLip::compute + solve1DQP (src/mpcLinearPendulum.cpp:78-133) for n problems on the device as a DENSE solve per
 * problem: g = beta Pu'(Px x_k - z), (alpha I + beta Pu'Pu) u = -g through the Cholesky factor of the (constant) Hessian,
 * shared by the batch and by both axes; x+ = A x_k + B u0.  The horizon, dt, zCoM, alpha, beta are the batch's.
 * com_xy, vel_xy [n][2]; zmp_x, zmp_y [n][horizon + 1] = zmpRef[k : k + horizon + 1] of each problem; xref, yref [n][3] =
 * (position, velocity, acceleration) as getXRef / getYRef return them.  use_dmma = 1: the dense tile products run on the
 * FP64 tensor cores (mma.sync m8n8k4), 0: on the FP64 FMA pipe.  repeats >= 1 solves every problem that many times
 * (micro-benchmark); elapsed_ms (or NULL): device time of the launch. */
int32_t lmh_mpc_solve(LmhBatch* h, int32_t n, const double* com_xy, const double* vel_xy, const double* zmp_x,
                      const double* zmp_y, double* xref, double* yref, int32_t use_dmma, int32_t repeats, double* elapsed_ms);
/* Setup tables of the same: Px [n1][2], Pu [n1][n1] (row-major; mpcLinearPendulum.cpp:41-68), n1 = horizon + 1. */
int32_t lmh_get_mpc_tables(const LmhBatch* h, double* Px, double* Pu, int32_t cap, int32_t* n1);

/* ---- host-only scenario pieces (no device work, no handle) ----
 * footCoeffTrajectory (src/footRefTrajectory.cpp:4-47, findPolyCoeff generalizedFunctions.cpp:103-163): boundary-value
 * polynomials of one foot over [0, T], ascending powers: x (6), y (6), z (8, through step_height at T/2). */
int32_t lmh_foot_coeff(const double* cur3, const double* des3, double step_height, double T, double* x6, double* y6, double* z8);
/* Config 2: one velocity impulse per environment on the base linear x,y velocity, magnitude ~U(0, max_speed), direction
 * ~U(0, 2 pi), at step ~U{0..max_step}; counter-based generator keyed by (seed, GLOBAL environment id): any sharding
 * of the ids gives the same per-environment integers and doubles.  push_step [n], push_dv6 [n][6]. */
int32_t lmh_push_plan(const int64_t* env_ids, int32_t n, uint64_t seed, double max_speed, int32_t max_step, int32_t* push_step,
                      double* push_dv6);
/* EXTENSION -- ZMP::walkZMP (declared, never defined upstream: zmpGeneration.hpp:15-22) + per-phase swing polynomials +
 * contact masks.  A plan is n_footsteps steps (right foot first) of one length, a phase of double support before and after;
 * the ZMP reference is piecewise constant: under the stance foot in single support, between the feet in double support.
 * Sizes: n_phases = n_footsteps + 2, zmp_len = n_phases * round(time_per_step / dt) + tail_steps + horizon_steps + 2. */
typedef struct {
    int32_t n_footsteps;       /* 4 */
    int32_t horizon_steps;     /* 160: samples the MPC looks ahead (zmp_len covers the last window) */
    int32_t tail_steps;        /* 40 */
    int32_t pad;
    double time_per_step;      /* 0.4 s per phase (header default 0.5, zmpGeneration.hpp:38) */
    double dt;                 /* 0.01 */
    double step_height;        /* 0.02 */
    double lateral;            /* 0.05: |y| of the feet */
    double min_len, max_len;   /* step length ~U(min_len, max_len) per environment: 0.02, 0.05 */
    double fixed_len;          /* >= 0: every environment takes this step length instead (ZMP(Task::Walk,...)); < 0: random */
} LmhWalkSpec;
void lmh_default_walk_spec(LmhWalkSpec* spec);
int32_t lmh_walk_plan_dims(const LmhWalkSpec* spec, int32_t* n_phases, int32_t* zmp_len);
/* zmp_x, zmp_y [n][zmp_len]; foot_coef [n][n_phases][2][20]; contact [n][n_phases]; step_len [n] (or NULL) */
int32_t lmh_walk_plan(const LmhWalkSpec* spec, const int64_t* env_ids, int32_t n, uint64_t seed, double* zmp_x, double* zmp_y,
                      double* foot_coef, int32_t* contact, double* step_len);
/* generate the plans of environments first_env_id .. first_env_id + n_env - 1 and install them (lmh_set_walk_plans with
 * one plan per environment) without a trip through the caller */
int32_t lmh_make_walk_plans(LmhBatch* h, const LmhWalkSpec* spec, int64_t first_env_id, uint64_t seed);

/* Data-format glue for viewing rollouts in MuJoCo (host only, no device work): relabelMujocoMatrix
 * (apps/mujoco/main.cpp:182-200: q_mj = L q[6:30], blocks head / left leg / right leg / left arm / right arm,
 * L(14,17) = -1; upstream leaves the rest of L uninitialised, here it is zero) and the initial free-joint pose of
 * MujocoSim.cpp:77-85 (z + 0.035, identity quaternion).  q30 [n][30] -> qpos31 [n][31] = [x y z, quat wxyz, 24 joints]. */
int32_t lmh_export_mujoco(int32_t n, const double* q30, double* qpos31);

/* Launch geometry actually used. */
int32_t lmh_launch_info(const LmhBatch* h, int32_t* threads_per_env, int32_t* envs_per_block, int32_t* smem_bytes,
                        int32_t* blocks, int32_t* kernel_launches_total);

/* Measured FP64 FMA throughput of the device (TFLOP/s): the roofline denominator of this path. */
int32_t lmh_measure_fp64_peak(int32_t device, double* tflops, double* sm_clock_mhz, int32_t* sm_count);

#ifdef __cplusplus
}
#endif
#endif
