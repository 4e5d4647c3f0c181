/* ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
 * C entry points of the CPU restatement (liblmh_oracle.so), loaded with ctypes
 * from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
 * The same struct layouts are produced by the reference build (oracle/_ref).
 */
#ifndef ORC_CAPI_H
#define ORC_CAPI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Scenario of apps/offline/main.cpp:12-64; zero-initialise then orc_default_config(). */
typedef struct {
    double dt;            /* 0.01  (main.cpp:14) */
    double sim_time;      /* 5.0   (main.cpp:13) */
    double time_horizon;  /* 0.5   (main.cpp:38) */
    double com[3];        /* -0.02, 0, 0.26 (main.cpp:31) */
    double rf[6];         /* 0,-0.05,0,0,0,0 (main.cpp:24-25) */
    double lf[6];         /* 0,+0.05,0,0,0,0 (main.cpp:26-27) */
    int32_t second_wbc;   /* 1 = also run the app's duplicate WBC(t) call (main.cpp:105) */
    int32_t mpc_rebuild_hessian; /* 1 = rebuild alpha I + beta Pu'Pu per call (mpcLinearPendulum.cpp:89-90) */
    int32_t qp_long_double;      /* 1 = run the QP in 80-bit arithmetic (self-check of round-off) */
    int32_t reserved;
} OrcConfig;

/* Everything one dynamics() evaluation produces (row-major matrices). */
typedef struct {
    double M[900], C[30], Cg[30], AG[180], AGpqp[6], Jpqp[12], JFeet[360];
    double CoM[3], comVel[3], comAngMom[3], xRef[3], yRef[3];
    double qppRef[30], hGpRef[6], footAccRef[12];
    double H[5476], g[74], x[74], y[50];
    double qpp[30], tau[24], f[12];
    double T7[16], T14[16];
    double kkt[4];     /* stationarity, primal, complementarity, dual-sign residuals */
    double xdot[60];
    int32_t k_index;   /* (int)(t/dt), mpcLinearPendulum.cpp:92 */
    int32_t qp_status, qp_iters, pad;
} OrcStageDump;

/* Per-step log record of a rollout. */
typedef struct {
    double x[60];      /* state after the step */
    double com_k4[3];  /* robot.getCoM() as printed by main.cpp:86 (k4 stage state) */
    double rfoot[3], lfoot[3]; /* T[7], T[14] translations at the same stage */
    double tau_k1[24]; /* torques of the k1 stage (state at the step start) */
    int32_t k_index[4];
    int32_t qp_iters, qp_fail;
    int32_t phase[4];  /* EXTENSION (walking): phase index of each stage; 0 in the reference's stand task */
    int32_t contact[4];/* contact mask of each stage: bit 0 right foot, bit 1 left foot; 3 in the stand task */
} OrcStepLog;

/* EXTENSION (not in the reference, which has no notion of falling): the health rule of a closed-loop step.
 * A step is COMMITTED (state, clock, statistics) only if, with x+ the state the RK4 combination would produce,
 *   max|tau(k1)| < ORC_TAU_LIMIT,  |CoM_xy(k4) - zmp_ref(k_index(k4))| < ORC_COM_ERR_LIMIT,
 *   every |x+_i| < ORC_STATE_LIMIT (NaN fails),  x+[2] (torso height) > ORC_TORSO_MIN.
 * Otherwise the environment is FALLEN: state and clock keep the values of the last healthy step, its statistics
 * are closed (they never see the failing step), alive = 0, and no further stage is evaluated for it.  The GPU
 * path (lmh_model.cuh rk4_step) implements the same rule with the same FP64 comparisons. */
#define ORC_TAU_LIMIT 100.0
#define ORC_COM_ERR_LIMIT 0.5
#define ORC_STATE_LIMIT 1e6
#define ORC_TORSO_MIN 0.15

void orc_default_config(OrcConfig* cfg);
void* orc_sim_create(const OrcConfig* cfg);   /* Robot + IK + Mpc3dLip + ZMP + coeffs + Controller */
void orc_sim_destroy(void* sim);
int32_t orc_sim_ik_iterations(void* sim);
double orc_sim_mass(void* sim);
double orc_sim_zcom(void* sim);
void orc_sim_get_state(void* sim, double* x60, double* t, double* v_held30);
void orc_sim_set_state(void* sim, const double* x60, double t, const double* v_held30);
/* one dynamics(x,t) evaluation (main.cpp:91-122), mutating the sim like the reference does */
void orc_sim_stage(void* sim, const double* x60, double t, OrcStageDump* out);
/* n RK4 steps from the sim's state; clock advances t += dt (Clock.hpp:11) */
void orc_sim_step(void* sim, int32_t n_steps, OrcStepLog* log_or_null);
/* the stand() loop: while |t - T| > 0.01 (main.cpp:74); returns steps run; comx gets CoM x per step */
int32_t orc_sim_run_stand(void* sim, double* comx, int32_t max_steps);
/* setup-time tables */
void orc_sim_mpc_tables(void* sim, double* Px, double* Pu, int32_t* horizon);
void orc_sim_foot_coeffs(void* sim, double* r6x, double* r6y, double* r8z, double* l6x, double* l6y, double* l8z);
int32_t orc_sim_zmp(void* sim, double* zx, double* zy, int32_t cap);
void orc_sim_link_table(void* sim, double* mass28, double* com28x3, double* inertia28x9);
/* Robot::getT / getX after the last updateState: world transforms [28][16] and parent-relative motion transforms [28][36], row-major */
void orc_sim_frames(void* sim, double* T28x16, double* X28x36);

/* Batch of independent environments split statically over n_threads host threads.
 * x0: E x 60 (or NULL = IK posture, zero velocity); push_step[e] (or <0 = none) and
 * push_dv: E x 6 impulse added to v[0:6] at the START of that step.
 * Outputs (any may be NULL): x_out E x 60, com_out E x 3 (k4 stage of last step),
 * stats E x 4 = {max|tau|, max |CoM_xy - zmp| , qp iterations, qp failures}. */
int32_t orc_batch_rollout(const OrcConfig* cfg, int32_t n_env, const double* x0, const int32_t* push_step,
                          const double* push_dv, int32_t n_steps, int32_t n_threads, double* x_out,
                          double* com_out, double* stats);

/* Per-environment walking plans for orc_batch_rollout_ex: arrays [n_env][...] (same meaning as orc_sim_set_walk_plan). */
typedef struct {
    int32_t zmp_len, n_phases;
    double phase_T;
    const double* zmp_x;      /* [n_env][zmp_len] */
    const double* zmp_y;
    const double* foot_coef;  /* [n_env][n_phases][2][20] */
    const int32_t* contact;   /* [n_env][n_phases] */
} OrcWalkPlans;

/* orc_batch_rollout with optional walking plans, the health rule above and the full statistics block:
 * stats8 E x 8 = {max|tau(k1)|, max |CoM_xy - zmp_ref|, max|qpp(k1)|, qp iterations, qp failures, steps committed,
 * alive, t}; log (or NULL): [n_env][n_steps] step records.  Returns the total number of QP failures. */
int32_t orc_batch_rollout_ex(const OrcConfig* cfg, int32_t n_env, const double* x0, const int32_t* push_step,
                             const double* push_dv, const OrcWalkPlans* plans, int32_t n_steps, int32_t n_threads,
                             double* x_out, double* com_out, double* stats8, OrcStepLog* log);
/* alive flag and committed steps of a single sim (health rule) */
int32_t orc_sim_alive(void* sim);
int32_t orc_sim_steps(void* sim);

/* EXTENSION (walking, not in the reference): replaces the ZMP reference and installs per-phase foot polynomials
 * (foot_coef: [n_phases][2][20] = x(6) y(6) z(8), ascending powers of the phase-local time) and contact masks
 * (contact[n_phases]: bit 0 right foot, bit 1 left foot). */
void orc_sim_set_walk_plan(void* sim, int32_t zmp_len, const double* zx, const double* zy, int32_t n_phases,
                           double phase_T, const double* foot_coef, const int32_t* contact);
int32_t orc_sim_last_phase(void* sim);

/* Generic QP entry (config 5 microbenchmark, solver self-tests). */
int32_t orc_qp_solve(int32_t n, const double* H, const double* g, int32_t m, const double* A,
                     const double* lbA, const double* ubA, double* x, double* y, int32_t long_double,
                     int32_t* iters, double* kkt4);

#ifdef __cplusplus
}
#endif
#endif
