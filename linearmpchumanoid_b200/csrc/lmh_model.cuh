// lmh_model.cuh -- per-environment closed-loop stage pipeline of the batched NAO controller/simulator.
//
// One environment is evaluated by a GROUP of NT cooperating threads (NT = 32: one warp per env,
// __syncwarp between phases; NT = 64/128: named barriers) over a workspace that lives in shared
// memory.  Every phase is a parallel-for over independent work items followed by a group sync.
// The same source compiles for the host with NT = 1 (sequential loops): that build is used by the
// setup code (inverse kinematics of the initial posture) and by the CPU unit tests of the kernel
// logic; it is never a fallback for the GPU path.
//
// What is computed follows the reference (paths relative to the upstream repo):
//   Robot::updateState            src/Robot.cpp:45-160,176-238,264-298
//   Dynamics::computeAll          src/Dynamics.cpp:4-216   (RNEA bias, CRBA, centroidal, Jdot*qdot)
//   Kinematics::computeAll        src/invKinematics.cpp:72-149,271-276
//   Robot::updateVelocityState    src/Robot.cpp:271-274,300-310
//   Mpc3dLip::compute             src/mpcLinearPendulum.cpp:78-133
//   Controller::WBC / solveQP     src/controller.cpp:81-154,156-294,296-479
//   dynamics() closure, rk4Step   apps/offline/main.cpp:91-122, include/.../general/rk4.hpp:5-18
// How it is computed is not: spatial transforms are kept as (R,p) pairs, the gravity bias uses the
// linearity of RNEA in the base acceleration, the MPC is a precomputed gain vector, and the WBC QP is
// solved in a reduced space (see qp_* below) by a primal active-set method on the 32 contact
// coefficients.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include "lmh_capi.h"

#if defined(__CUDACC__)
#define LMH_HD __host__ __device__ __forceinline__
#define LMH_HDN __host__ __device__
#else
#define LMH_HD inline
#define LMH_HDN
#endif

namespace lmh {

constexpr int NQ = 30;      // generalized coordinates (6 base + 24 joints), Robot.hpp:11
constexpr int NJ = 24;      // actuated joints, Robot.hpp:12
constexpr int NF = 27;      // frames that matter (frame 27 of the reference is massless, unactuated, unused)
constexpr int NCH = 5;      // kinematic chains hanging off the torso
constexpr int MAXDEPTH = 7;
constexpr int NC = 32;      // contact coefficients (2 feet x 4 vertices x 4 pyramid edges)
constexpr int MPC_MAXN = 320;
constexpr int LDM = 31;     // odd leading dimension: conflict-free column walks in shared memory
constexpr int LDQ = 33;
#ifndef LMH_QP_TABLE
#define LMH_QP_TABLE 1          // 1: every pass of the general active-set iteration reads what depends on the working set alone from
                                // the 65536-row table and solves the 6 x 6 system of the shortcut; 0: pivoted Cholesky + 12 x 12 system per pass
#endif
#ifndef LMH_SHORTCUT_DMMA
#define LMH_SHORTCUT_DMMA 1     // 1: the 6 x 6 system of the all-free QP shortcut (D^-1 Jb, Sa + Jb'D^-1 Jb) on the FP64 tensor cores
#endif
#if defined(LMH_TIMELINE) && defined(__CUDA_ARCH__)
#define LMH_STAMP(i) do { if (cx.tid == 0) w.stamp[(i)] = (double)clock64(); } while (0)
#else
#define LMH_STAMP(i) do {} while (0)
#endif
#ifndef LMH_ALIGN_STAGES
#define LMH_ALIGN_STAGES 1      // 1: the warps of a block meet at a barrier once per stage (shared instruction fetches)
#endif
#ifndef LMH_RSQRT_CUBIC
#define LMH_RSQRT_CUBIC 1
#endif
#ifndef LMH_LOGMAP_SERIES
#define LMH_LOGMAP_SERIES 1
#endif
#ifndef LMH_QP_GENERAL_NOINLINE
#define LMH_QP_GENERAL_NOINLINE 0     // measured: as a real call it costs 5 % (spills around the call); inlined, the split
                                      // of qp_phase into hot path + general iteration removed the kernel's last spills
#endif
#ifndef LMH_STAGE_NOINLINE
#define LMH_STAGE_NOINLINE 0     // 1: the whole stage as one non-inlined function (measured: -14 % throughput)
#endif
#ifndef LMH_GRAM_DMMA
#define LMH_GRAM_DMMA 1     // 1: the 10 x 30 x 10 Gram product of the QP front end on the FP64 tensor cores (one-warp path)
#endif
#ifndef LMH_WARP_WRENCH
#define LMH_WARP_WRENCH 1
#endif
constexpr int ND1 = 18;     // base + both legs: the block of H_a that the sole Jacobians couple
constexpr int LD1 = 19;
constexpr int NYR = 10;     // right-hand-side rows carried through the factorisation

// chain c: frames chain_first[c] .. chain_first[c]+chain_len[c]-1, depth 1..chain_len[c]
// (five one-byte entries packed into a constant: a shift instead of a chain of compares and branches)
LMH_HD int chain_first(int c) { return (int)((0x19140F0801ull >> (8 * c)) & 0xFF); }    // 1, 8, 15, 20, 25
LMH_HD int chain_len(int c) { return (int)((0x0205050707ull >> (8 * c)) & 0xFF); }      // frames (legs include the sole)
LMH_HD int chain_nact(int c) { return (int)((0x0205050606ull >> (8 * c)) & 0xFF); }     // actuated frames
LMH_HD int chain_joint0(int c) { return (int)((0x16110C0600ull >> (8 * c)) & 0xFF); }   // 0, 6, 12, 17, 22
// actuated joint index (0..23) of a frame, or -1 (src/Robot.cpp:169-174)
LMH_HD int frame_joint(int i) {
    if (i == 0 || i == 7 || i == 14) return -1;
    return i < 7 ? i - 1 : i < 14 ? i - 2 : i - 3;
}
LMH_HD int frame_parent(int i) { return ((0x2108102u >> i) & 1u) ? 0 : i - 1; }     // frames 1, 8, 15, 20, 25 hang off the torso (bit mask: no branches)
LMH_HD int joint_chain(int a) { return a < 6 ? 0 : a < 12 ? 1 : a < 17 ? 2 : a < 22 ? 3 : 4; }

// ------------------------------------------------------------------------------------------------
// Constants shared by every environment of a batch (device global memory, read-only).
struct Consts {
    double Isp[28 * 36];      // spatial inertia of each frame, joint frame, row-major 6x6 (Dynamics.cpp:4-13)
    double mc[28 * 4];        // {com_x, com_y, com_z, mass} in the joint frame
    double th_off[24];        // theta offset of each joint's DH transform (Robot.cpp:59-87)
    double dh_ca[24], dh_sa[24], dh_r[24], dh_d[24];   // cos/sin(alpha), r, d (Robot.cpp:180-196)
    double pre[NCH * 12];     // fixed transform between the torso and each chain's first joint (R row-major, p)
    double sole_x;            // sole offset along x of the ankle-roll frame (Robot.cpp:106-117)
    double mass, gravity;
    double inv_mass;          // 1 / mass: divisions by constants are multiplications on the device
    double qdes[NQ];          // Robot::desiredPosture (Robot.cpp:253-262)
    double Rdes[9];           // desired sole orientation (Robot.cpp:28-31)
    double Gf[6 * 16];        // per-foot wrench generators: rows 0-2 moment, 3-5 force (controller.cpp:156-294)
    double GGt[36];           // Gf Gf' (6 x 6), the all-free wrench Gram matrix of one foot
    double Pinv[36], Dinv[36];                // GGt^-1 and (wForce I + epsC GGt^-1)^-1: the all-free shortcut of the QP
    double sq_wCoML, isq_wJoints;   // sqrt(wCoML), 1/sqrt(wJoints)
    double KpJ, KdJ, KpM, KdM, KpF, KdF;                                    // controller.hpp:100-111
    double wCoML, wCoMK, wBasePos, wBaseAng, wJoints, wForce, wFoot, epsC;   // controller.hpp:118-124, controller.cpp:109-123
    // LIPM MPC (mpcLinearPendulum.cpp:41-68,89-101): u0 = -(rPx . [pos,vel]) + sum_i r[i]*zmp[k+i]
    double mpc_dt, mpc_A[4], mpc_B[2], mpc_rPx[2], mpc_zcom;
    int32_t mpc_n1;           // horizon + 1
    int32_t qp_max_iter;
    double mpc_r[MPC_MAXN];
    double dt;                // integrator step (Clock.hpp)
    const double* qp_tab;     // working-set table of the WBC QP: [65536][72] = Kf (36), Yf (36) per subset of a foot's generators
                              // (lmh_setup.hpp: qp_working_set_table); device pointer in the device copy of this struct
};

// Per-environment plan data shared through plan ids (global memory).
struct Plan {
    const double* zmp_x;      // [n_plans][zmp_len]
    const double* zmp_y;
    const double* foot_coef;  // [n_plans][n_phases][2 feet][20] : x(6) y(6) z(8), ascending powers of the phase-local
                              // time (footRefTrajectory.cpp:4-47)
    const int32_t* contact;   // [n_plans][n_phases] : bit 0 right foot, bit 1 left foot may carry force
    double phase_T;           // phase p = min((int)(t/phase_T), n_phases-1).  EXTENSION: upstream has one phase (stand).
    int32_t zmp_len, n_phases;
};

// Running statistics of one environment (a member of its workspace; thread 0 of the group updates it once per step).
struct StatsAcc {
    double max_tau, max_err, max_qpp;
    int32_t qp_iters, qp_fail, steps, alive, qp_fallback, pad;
};

// ------------------------------------------------------------------------------------------------
// Workspace of one environment (shared memory on the device).  Regions are reused across the
// model / assembly / QP phases; see DESIGN.md for the lifetime table.
struct EnvWS {
    // ---- state of the stage being evaluated
    double q[NQ], v[NQ], vh[NQ];   // vh: velocity held by the Robot object from the PREVIOUS stage (controller.cpp:48-79)
    alignas(16) double vb6[6], vhb6[6];        // base twists of v and vh as [omega; v] in the torso frame (joint rates are unchanged)
    double t;
    // ---- persistent through a stage
    double R0[9], p0[3];           // torso pose
    double esc[4];                 // sin, cos of pitch and yaw (kept for the Euler-rate map of the state derivative)
    double Rs[2][9], ps[2][3];     // sole poses (frames 7, 14)
    double M6[6 * NQ];             // rows 0..5 of the mass matrix [Ic0 | F2]
    double Hc[NCH * 36];           // joint-space inertia, one 6x6 block per chain (chains do not couple)
    double C[NQ], Cg[NQ];
    double AG[6 * NQ], AGpqp[6];
    double JFc[2 * 6 * 12];        // sole Jacobians, compact: [foot][row][0..5 base twist | 6..11 that leg's joints]
    double Jpqp[12];
    double com[3], comVel[3], angMom[3], xRef[3], yRef[3];
    double qppRef[NQ], hGpRef[6], footAccRef[12];
    double ga[NQ], a[NQ], f12[12], tau[NJ], xdot[2 * NQ];
    double Sa[36], Ls[36], Lsi[6], rhs[6], sv[6], lam[6];     // Lsi: 1 / diag(Ls)
    double c[NC], cs[NC], mu[NC], piv[NC], ladiag[ND1], ladinv[ND1];     // ladinv: 1 / ladiag (one-warp factorisation)
    double red[32];                // reduction scratch
    int32_t free_mask[NC];
    int32_t pdone[12], ppiv[2];
    int32_t pord[12];              // pivot row of column k of the wrench-space basis of foot f at [6 f + k], or -1 (beyond the rank)
    int32_t k_index, qp_iters, qp_status, qp_fallback;
    int32_t phase, contact;        // walking extension: phase index and contact mask of this stage
    int32_t warm_contact, dead;    // contact mask the stored working set (free_mask, c) belongs to (-1 = none);
                                   // dead: the environment fell (state frozen, statistics closed)
    // ---- bookkeeping of the closed loop, kept here so that nothing of it is live in registers across a stage evaluation
    double t_env;                  // the environment's clock (rollout kernels)
    double acc_tau, acc_qpp;       // step in flight: max |tau|, max |qpp| of its first stage
    StatsAcc stats;                // running statistics (rollout kernels)
    int32_t acc_iters, acc_fails, acc_nfall, acc_pad;
    alignas(16) double dump[4];    // where a lane with nothing to store writes (branch-free conditional stores); never read
#if defined(LMH_TIMELINE)
    double stamp[16];              // debug build: clock64 at the phase boundaries of the last stage evaluated
#endif
    // ---- phase-overlaid scratch
    union alignas(16) {
        struct {                       // model phase
            double Xl[NF * 12];        // local transform first, then parent-relative (R_rel, p_rel)
            union {
                double Tw[NF * 12];    // world transforms (R row-major 9, p 3): forward kinematics only
                struct { double vel[NF * 6], acc[NF * 6], frc[NF * 6]; } rnea;
                struct {
                    double IcIn[NCH * 36];     // composite inertia arriving from the level below
                    double IcCur[NCH * 36];    // composite inertia of the level being processed
                    double Yt[NCH * 36];
                    double Fb[2][NJ * 6];
                } crba;
                double Xn[2][2 * 18];  // running sole_X_j as (E, B) for both feet, double buffered
                struct {               // exchange buffers of the one-warp-per-env versions (lmh_warp.cuh)
                    // six-vectors at stride 8 (chain c at 8c, frame i at 8i): 16-byte aligned for vector loads
                    double vbuf[2][40], abuf[2][40], hbuf[40], contrib[40], col2[40], accs[12];
                    double frc[NF * 8];
                    double Ybuf[NCH * 36], Ybuf2[NCH * 36];
                } wr;
            } p;
        } m;
        struct {                       // assembly + QP phase
            double L1[ND1 * LD1];      // D1 = Wj + wFoot J'J on base+legs (18 x 18), then its Cholesky factor (lower)
            double Yt[NYR * NQ];       // rows M6[0..5], ga, U[0..2] (U = sqrt(wCoML) AG[3..5]); cols 0..17: L1^-1 applied,
                                       // cols 18..29: scaled by 1/sqrt(wJoints)
            double Gm[NYR * NYR];      // Gram matrix Yt Yt'
            double Bm[12 * 6], Pm[2 * 36], Wb[6 * 12];   // wrench-space basis, its Gram matrix, Ls^-1 Jb' B
            double Ay[13 * 13];        // wrench-space system: lower triangle factorised in place, row 12 = rhs / zeta
            double vv[36];
        } s;
    } u;
};

static_assert(offsetof(EnvWS, p0) == offsetof(EnvWS, R0) + 9 * sizeof(double), "(R0, p0) must be one (R, p) pair");

// compact column of generalized coordinate i in the Jacobian of foot s, or -1
LMH_HD int jf_col(int s, int i) {
    if (i < 6) return i;
    int l = i - 6 - 6 * s;
    return (l >= 0 && l < 6) ? 6 + l : -1;
}
LMH_HD double jf_full(const EnvWS& w, int row, int col) {
    int s = row / 6, cc = jf_col(s, col);
    return cc < 0 ? 0.0 : w.JFc[72 * s + 12 * (row % 6) + cc];
}
LMH_HD double vb_at(const EnvWS& w, int j) { return j < 6 ? w.vb6[j] : w.v[j]; }

// ------------------------------------------------------------------------------------------------
template <int NT_, bool TAB_ = false>
struct Ctx {
    static constexpr int NT = NT_;
    static constexpr bool TAB = TAB_;   // isp and dh point into the block's shared memory (known at compile time, so that
                                        // the tree recursions read them with LDS instead of generic loads)
    int tid;
    int bar_id;          // several one-warp environments share a block and re-align at align() points: < 0 all warps of
                         // the block (plain __syncthreads), > 32 the number of threads taking part (32 x live warps)
    // block-shared copies (shared memory) of the constant tables the tree recursions read at every level, or null =
    // read them from the Consts in global memory: Isp [28][36]; dh = th_off, dh_ca, dh_sa, dh_r, dh_d [5][24]; mpc_r
    const double* isp = nullptr;
    const double* dh = nullptr;
    const double* mpc_r = nullptr;
    // Latency mode (rollout_kernel_helpers): the environment's warp is joined by two helper warps that take the CRBA
    // and the sole Jacobians while it runs the RNEA (the three only read the transforms); helpers = threads of the block
    // taking part in the hand-over barriers, 0 = no helpers
    int helpers = 0;
    LMH_HD void handover(int id) const {
#if defined(__CUDA_ARCH__)
        if (helpers) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(helpers) : "memory");
#else
        (void)id;
#endif
    }
    LMH_HD void align() const {
#if defined(__CUDA_ARCH__) && LMH_ALIGN_STAGES
        if (NT == 32) {
            if (bar_id < 0) __syncthreads();                 // every warp of the block is live
            else if (bar_id > 32) asm volatile("bar.sync 1, %0;" ::"r"(bar_id) : "memory");
        }
#endif
    }
    LMH_HD void sync() const {
#if defined(__CUDA_ARCH__)
        if (NT == 32) __syncwarp();
        else if (NT > 32) asm volatile("bar.sync 0, %0;" ::"n"(NT_) : "memory");
#endif
    }
};
#define LMH_FOR(i, n) for (int i = cx.tid; i < (n); i += Cx::NT)
#if defined(__CUDACC__)
#define LMH_UNROLL _Pragma("unroll")
#else
#define LMH_UNROLL
#endif

#if defined(__CUDACC__)
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif

// a + b*c with the product rounded before the sum, as the reference's (unfused) CPU arithmetic does
LMH_HD double mul_then_add(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, __dmul_rn(b, c));
#else
    volatile double p = b * c;
    return a + p;
#endif
}

LMH_HD double lmh_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    // hardware seed (2^-22) and two Newton steps; the arguments are pivots of positive definite matrices, so the
    // library's special-case path (a branch and a call) has nothing to do here
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if LMH_RSQRT_CUBIC
    // one third-order step, y (1 + e/2 + 3 e^2/8) with e = 1 - x y^2: from a 2^-22 seed the error term (5/16) e^3 is below
    // 2^-66, so this is as exact as two Newton steps at five dependent operations instead of eight (the factorisations
    // call this eighteen times in a row per stage, each waiting for the one before)
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double q = e * fma(0.375, e, 0.5);
    return fma(y, q, y);
#else
    double t = x * y;
    double e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    t = x * y;
    e = fma(-t, y, 1.0);
    return fma(0.5 * y, e, y);
#endif
#else
    return 1.0 / sqrt(x);
#endif
}

// in-place Cholesky of a 6x6 SPD matrix held in registers (lower triangle), fully unrolled
LMH_HD void chol6(double* A, double* dinv = nullptr) {      // dinv (or null): 1 / diag(L), a by-product of the rsqrt
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int k = 0; k < 6; ++k) {
        double s = A[7 * k];
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int m = 0; m < 6; ++m) if (m < k) s -= A[6 * k + m] * A[6 * k + m];
        const double inv = lmh_rsqrt(s);
        A[7 * k] = s * inv;
        if (dinv) dinv[k] = inv;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (int i = 0; i < 6; ++i) {
            if (i > k) {
                double t = A[6 * i + k];
#if defined(__CUDACC__)
#pragma unroll
#endif
                for (int m = 0; m < 6; ++m) if (m < k) t -= A[6 * i + m] * A[6 * k + m];
                A[6 * i + k] = t * inv;
            }
        }
    }
}

LMH_HD void lmh_sincos(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
    sincos(x, s, c);
#else
    *s = sin(x);
    *c = cos(x);
#endif
}

// y = R x (R row-major 3x3), one component
LMH_HD double rot_row(const double* R, int k, const double* x) { return R[3 * k] * x[0] + R[3 * k + 1] * x[1] + R[3 * k + 2] * x[2]; }
// y = R' x, one component
LMH_HD double rotT_row(const double* R, int k, const double* x) { return R[k] * x[0] + R[3 + k] * x[1] + R[6 + k] * x[2]; }
LMH_HD double cross_k(const double* a, const double* b, int k) {
    // all three components, then a select: the lazy form compiles to a branch around each arm's loads
    const double a0 = a[0], a1 = a[1], a2 = a[2], b0 = b[0], b1 = b[1], b2 = b[2];
    const double c0 = a1 * b2 - a2 * b1, c1 = a2 * b0 - a0 * b2, c2 = a0 * b1 - a1 * b0;
    return k == 0 ? c0 : k == 1 ? c1 : c2;
}
// Motion transform of the reference, X = [[E,0],[-E p^,E]] with E = R', applied to [w; u] (generalizedFunctions.cpp:11-19)
LMH_HD double xmotion_k(const double* Rp, const double* m, int k) {
    const double* R = Rp;
    const double* p = Rp + 9;
    if (k < 3) return rotT_row(R, k, m);
    double u[3] = {m[3] + cross_k(m, p, 0), m[4] + cross_k(m, p, 1), m[5] + cross_k(m, p, 2)};   // v + w x p
    return rotT_row(R, k - 3, u);
}
// Its transpose applied to a force [n; f]:  [R n + p x (R f); R f]
LMH_HD double xforceT_k(const double* Rp, const double* f, int k) {
    const double* R = Rp;
    const double* p = Rp + 9;
    if (k >= 3) return rot_row(R, k - 3, f + 3);
    double rf[3] = {rot_row(R, 0, f + 3), rot_row(R, 1, f + 3), rot_row(R, 2, f + 3)};
    return rot_row(R, k, f) + cross_k(p, rf, k);
}
// dense 6x6 entry of X (row a, col b)
LMH_HD double xdense(const double* Rp, int a, int b) {
    const double* R = Rp;
    const double* p = Rp + 9;
    int ar = a % 3, bc = b % 3;
    if (a < 3) return b < 3 ? R[3 * bc + ar] : 0.0;
    if (b >= 3) return R[3 * bc + ar];
    // -(E p^)[ar][bc] = -sum_k E[ar][k] * cm[k][bc], cm = crossMatrix(p)
    double e0 = R[ar], e1 = R[3 + ar], e2 = R[6 + ar];
    return bc == 0 ? -(e1 * p[2] - e2 * p[1]) : bc == 1 ? -(-e0 * p[2] + e2 * p[0]) : -(e0 * p[1] - e1 * p[0]);
}

#if defined(__CUDACC__)
struct EnvWS;
struct Consts;
__device__ void fk_chain_warp(EnvWS& w, const Consts& K, int lane);
__device__ void fk_local_warp(EnvWS& w, const Consts& K, const double* dh, int lane);
#endif

// state that survives from one stage evaluation to the next besides vh: the QP working set (warm start)
template <class Cx>
LMH_HDN void ws_reset(const Cx& cx, EnvWS& w) {
    LMH_FOR(i, NC) { w.free_mask[i] = 1; w.c[i] = 0.0; }
    LMH_FOR(i, 1) { w.warm_contact = -1; w.dead = 0; }
    cx.sync();
}

// ================================================================================================
// Phase 1: forward kinematics, parent-relative transforms, CoM.
template <class Cx>
LMH_HDN void fk_phase(const Cx& cx, EnvWS& w, const Consts& K) {
    double* Tw = w.u.m.p.Tw;
    double* Xl = w.u.m.Xl;
#if defined(__CUDA_ARCH__)
    if (Cx::NT == 32) {
        const double* dh = Cx::TAB ? cx.dh : K.th_off;
        if (Cx::TAB) __builtin_assume(__isShared(dh));
        fk_local_warp(w, K, dh, cx.tid); fk_chain_warp(w, K, cx.tid); return;
    }
#endif
    // torso pose (generalizedFunctions.cpp:52-72) and local DH transforms (Robot.cpp:176-223)
    LMH_FOR(i, NF) {
        if (i == 0) {
            double sr, cr, sp, cp, sy, cy;
            lmh_sincos(w.q[3], &sr, &cr);
            lmh_sincos(w.q[4], &sp, &cp);
            lmh_sincos(w.q[5], &sy, &cy);
            double* R = Tw;
            R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
            R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
            R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
            R[9] = w.q[0]; R[10] = w.q[1]; R[11] = w.q[2];
            w.esc[0] = sp; w.esc[1] = cp; w.esc[2] = sy; w.esc[3] = cy;
            for (int k = 0; k < 9; ++k) w.R0[k] = R[k];
            for (int k = 0; k < 3; ++k) w.p0[k] = R[9 + k];
        } else if (i == 7 || i == 14) {
            double* L = Xl + 12 * i;
            L[0] = 1; L[1] = 0; L[2] = 0; L[3] = 0; L[4] = 1; L[5] = 0; L[6] = 0; L[7] = 0; L[8] = 1;
            L[9] = K.sole_x; L[10] = 0; L[11] = 0;
        } else {
            int a = frame_joint(i);
            double st, ct;
            lmh_sincos(w.q[6 + a] + K.th_off[a], &st, &ct);
            double ca = K.dh_ca[a], sa = K.dh_sa[a], r = K.dh_r[a], d = K.dh_d[a];
            double L[12] = {ct, -st, 0.0, ca * st, ca * ct, -sa, sa * st, sa * ct, ca, d, -r * sa, r * ca};
            double* O = Xl + 12 * i;
            int c = (i == 1) ? 0 : (i == 8) ? 1 : (i == 15) ? 2 : (i == 20) ? 3 : (i == 25) ? 4 : -1;
            if (c < 0) {
                for (int k = 0; k < 12; ++k) O[k] = L[k];
            } else {   // fold the fixed torso->chain transform into the first joint of the chain
                const double* P = K.pre + 12 * c;
                for (int r3 = 0; r3 < 3; ++r3) {
                    for (int cc = 0; cc < 3; ++cc)
                        O[3 * r3 + cc] = P[3 * r3] * L[cc] + P[3 * r3 + 1] * L[3 + cc] + P[3 * r3 + 2] * L[6 + cc];
                    O[9 + r3] = P[3 * r3] * L[9] + P[3 * r3 + 1] * L[10] + P[3 * r3 + 2] * L[11] + P[9 + r3];
                }
            }
        }
    }
    cx.sync();
    // chained world transforms, one tree level per pass
    for (int d = 1; d <= MAXDEPTH; ++d) {
        LMH_FOR(it, NCH * 12) {
            int c = it / 12, e = it % 12;
            if (d <= chain_len(c)) {
                int i = chain_first(c) + d - 1;
                const double* P = Tw + 12 * frame_parent(i);
                const double* L = Xl + 12 * i;
                double* O = Tw + 12 * i;
                if (e < 9) {
                    int r3 = e / 3, cc = e % 3;
                    O[e] = P[3 * r3] * L[cc] + P[3 * r3 + 1] * L[3 + cc] + P[3 * r3 + 2] * L[6 + cc];
                } else {
                    int r3 = e - 9;
                    O[e] = P[3 * r3] * L[9] + P[3 * r3 + 1] * L[10] + P[3 * r3 + 2] * L[11] + P[9 + r3];
                }
            }
        }
        cx.sync();
    }
    // parent-relative transform "inverse-by-transpose(T_parent) * T_i" (Robot.cpp:276-298, generalizedFunctions.cpp:21-27)
    LMH_FOR(it, NF * 12) {
        int i = it / 12, e = it % 12;
        double* O = Xl + 12 * i;
        const double* Ti = Tw + 12 * i;
        if (i == 0) {
            O[e] = Ti[e];
        } else {
            const double* P = Tw + 12 * frame_parent(i);
            if (e < 9) {
                int r3 = e / 3, cc = e % 3;
                O[e] = P[r3] * Ti[cc] + P[3 + r3] * Ti[3 + cc] + P[6 + r3] * Ti[6 + cc];
            } else {
                int r3 = e - 9;
                double pi = P[r3] * Ti[9] + P[3 + r3] * Ti[10] + P[6 + r3] * Ti[11];
                double pp = P[r3] * P[9] + P[3 + r3] * P[10] + P[6 + r3] * P[11];
                O[e] = pi - pp;
            }
        }
    }
    // centre of mass partial sums (Robot.cpp:225-238); 3 components x 8 partials
    LMH_FOR(it, 24) {
        int k = it % 3, part = it / 3;
        double s = 0.0;
        for (int i = part; i < NF; i += 8) {
            const double* Ti = Tw + 12 * i;
            const double* mc = K.mc + 4 * i;
            s += mc[3] * (Ti[3 * k] * mc[0] + Ti[3 * k + 1] * mc[1] + Ti[3 * k + 2] * mc[2] + Ti[9 + k]);
        }
        w.red[it] = s;
    }
    LMH_FOR(it, 2 * 12) {   // keep the sole poses
        int s = it / 12, e = it % 12;
        const double* Ti = Tw + 12 * (s == 0 ? 7 : 14);
        if (e < 9) w.Rs[s][e] = Ti[e]; else w.ps[s][e - 9] = Ti[e];
    }
    cx.sync();
    LMH_FOR(k, 3) {
        double s = 0.0;
        for (int part = 0; part < 8; ++part) s += w.red[3 * part + k];
        w.com[k] = s * K.inv_mass;
    }
    cx.sync();
}

// base twist of a generalized velocity into the torso frame: [w; v] -> X0 [w; v]  (generalizedFunctions.cpp:195-206)
template <class Cx>
LMH_HDN void base_twist_phase(const Cx& cx, EnvWS& w) {
    LMH_FOR(it, 12) {
        int which = it / 6, k = it % 6;
        const double* src = which == 0 ? w.v : w.vh;
        double m[6] = {src[3], src[4], src[5], src[0], src[1], src[2]};
        (which == 0 ? w.vb6 : w.vhb6)[k] = xmotion_k(w.R0, m, k);     // (R0, p0) are adjacent: one (R, p) pair
    }
    cx.sync();
}

// ================================================================================================
// Phase 2: RNEA with zero base acceleration on the HELD velocity -> Cg, Jdot*qdot of both soles
// (Dynamics.cpp:29-60,124-200).
template <class Cx>
LMH_HDN void rnea_phase(const Cx& cx, EnvWS& w, const Consts& K) {
    double* vel = w.u.m.p.rnea.vel;
    double* acc = w.u.m.p.rnea.acc;
    double* frc = w.u.m.p.rnea.frc;
    const double* Xl = w.u.m.Xl;
    LMH_FOR(k, 6) { vel[k] = w.vhb6[k]; acc[k] = 0.0; }
    cx.sync();
    for (int d = 1; d <= MAXDEPTH; ++d) {
        LMH_FOR(it, NCH * 6) {
            int c = it / 6, k = it % 6;
            if (d <= chain_len(c)) {
                int i = chain_first(c) + d - 1;
                int a = frame_joint(i);
                double x = xmotion_k(Xl + 12 * i, vel + 6 * frame_parent(i), k);
                if (a >= 0 && k == 2) x += w.vh[6 + a];
                vel[6 * i + k] = x;
            }
        }
        cx.sync();
        LMH_FOR(it, NCH * 6) {
            int c = it / 6, k = it % 6;
            if (d <= chain_len(c)) {
                int i = chain_first(c) + d - 1;
                int a = frame_joint(i);
                double x = xmotion_k(Xl + 12 * i, acc + 6 * frame_parent(i), k);
                if (a >= 0) {   // + (v x S) qd with S = e_z
                    double qd = w.vh[6 + a];
                    const double* vi = vel + 6 * i;
                    double col = k == 0 ? vi[1] : k == 1 ? -vi[0] : k == 3 ? vi[4] : k == 4 ? -vi[3] : 0.0;
                    x += col * qd;
                }
                acc[6 * i + k] = x;
            }
        }
        cx.sync();
    }
    // h = I v  (into frc), then f = I a + v x* h
    LMH_FOR(it, NF * 6) {
        int i = it / 6, k = it % 6;
        const double* I = K.Isp + 36 * i + 6 * k;
        const double* vi = vel + 6 * i;
        frc[it] = I[0] * vi[0] + I[1] * vi[1] + I[2] * vi[2] + I[3] * vi[3] + I[4] * vi[4] + I[5] * vi[5];
    }
    cx.sync();
    double fnew[(NF * 6 + Cx::NT - 1) / Cx::NT];
    {
        int n = 0;
        LMH_FOR(it, NF * 6) {
            int i = it / 6, k = it % 6;
            const double* I = K.Isp + 36 * i + 6 * k;
            const double* ai = acc + 6 * i;
            const double* vi = vel + 6 * i;
            const double* h = frc + 6 * i;
            double x = I[0] * ai[0] + I[1] * ai[1] + I[2] * ai[2] + I[3] * ai[3] + I[4] * ai[4] + I[5] * ai[5];
            // crf(v) h = [w x n + u x f ; w x f]
            if (k < 3) x += cross_k(vi, h, k) + cross_k(vi + 3, h + 3, k);
            else x += cross_k(vi, h + 3, k - 3);
            fnew[n++] = x;
        }
    }
    cx.sync();
    {
        int n = 0;
        LMH_FOR(it, NF * 6) frc[it] = fnew[n++];
    }
    cx.sync();
    // backward pass, deepest level first; the soles carry no force
    for (int d = MAXDEPTH; d >= 2; --d) {
        LMH_FOR(it, NCH * 6) {
            int c = it / 6, k = it % 6;
            if (d <= chain_nact(c)) {
                int i = chain_first(c) + d - 1;
                frc[6 * (i - 1) + k] += xforceT_k(Xl + 12 * i, frc + 6 * i, k);
            }
        }
        cx.sync();
    }
    LMH_FOR(k, 6) {
        double s = frc[k];
        for (int c = NCH - 1; c >= 0; --c) s += xforceT_k(Xl + 12 * chain_first(c), frc + 6 * chain_first(c), k);
        w.Cg[k] = s;
    }
    LMH_FOR(a, NJ) {
        int c = joint_chain(a);
        int i = chain_first(c) + (a - chain_joint0(c));
        w.Cg[6 + a] = frc[6 * i + 2];
    }
    // Jdot*qdot = world-rotated spatial acceleration of the soles (Dynamics.cpp:165-200)
    LMH_FOR(it, 12) {
        int s = it / 6, k = it % 6;
        const double* ai = acc + 6 * (s == 0 ? 7 : 14) + (k < 3 ? 0 : 3);
        w.Jpqp[it] = rot_row(w.Rs[s], k % 3, ai);
    }
    cx.sync();
}

// ================================================================================================
// Phase 3: CRBA (Dynamics.cpp:62-101) -> M6 = [Ic0 | F2], Hc (per-chain joint blocks).
template <class Cx>
LMH_HDN void crba_phase(const Cx& cx, EnvWS& w, const Consts& K) {
    double* IcIn = w.u.m.p.crba.IcIn;
    double* IcCur = w.u.m.p.crba.IcCur;
    double* Yt = w.u.m.p.crba.Yt;
    const double* Xl = w.u.m.Xl;
    LMH_FOR(it, NCH * 36) { IcIn[it] = 0.0; w.Hc[it] = 0.0; }
    cx.sync();
    int cur = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int d = MAXDEPTH - 1; d >= 1; --d) {   // actuated depths 6..1; only two levels of composite inertia are alive
        // (a) composite inertia of this level's frames = own inertia + what arrived from below; dense X;
        //     new F vector of each frame = column 2 of its composite inertia
        LMH_FOR(it, NCH * 36) {
            int c = it / 36, e = it % 36;
            if (d <= chain_nact(c)) {
                int i = chain_first(c) + d - 1;
                double v = K.Isp[36 * i + e] + IcIn[it];
                IcCur[it] = v;
                if (e % 6 == 2) {
                    int k = e / 6, a = chain_joint0(c) + d - 1;
                    w.u.m.p.crba.Fb[cur][6 * a + k] = v;
                    if (k == 2) w.Hc[36 * c + 7 * (d - 1)] = v;
                }
            }
        }
        cx.sync();
        // (b) Yt = Ic X ; F vectors of every joint at depth >= d move one frame up: F <- X' F
        LMH_FOR(it, NCH * 36) {
            int c = it / 36, e = it % 36;
            if (d <= chain_nact(c)) {
                int r = e / 6, cc = e % 6;
                const double* I = IcCur + 36 * c + 6 * r;
                const double* Rp = Xl + 12 * (chain_first(c) + d - 1);
                double sy = 0.0;
                for (int k = 0; k < 6; ++k) sy += I[k] * xdense(Rp, k, cc);
                Yt[it] = sy;
            }
        }
        LMH_FOR(it, NJ * 6) {
            int a = it / 6, k = it % 6;
            int c = joint_chain(a);
            int da = a - chain_joint0(c) + 1;       // depth of joint a
            if (da >= d) {
                int i = chain_first(c) + d - 1;     // frame being crossed
                double x = xforceT_k(Xl + 12 * i, w.u.m.p.crba.Fb[cur] + 6 * a, k);
                w.u.m.p.crba.Fb[cur ^ 1][it] = x;
                if (d > 1 && k == 2) {              // H(parent joint, a), symmetric
                    int pj = d - 2, aj = da - 1;
                    w.Hc[36 * c + 6 * pj + aj] = x;
                    w.Hc[36 * c + 6 * aj + pj] = x;
                }
                if (d == 1) w.M6[NQ * k + 6 + a] = x;   // F2 column
            }
        }
        cx.sync();
        // (c) what the parent level receives: X' Yt
        if (d > 1) {
            LMH_FOR(it, NCH * 36) {
                int c = it / 36, e = it % 36;
                if (d <= chain_nact(c)) {
                    int r = e / 6, cc = e % 6;
                    const double* Rp = Xl + 12 * (chain_first(c) + d - 1);
                    const double* Y = Yt + 36 * c + cc;
                    double sx = 0.0;
                    for (int k = 0; k < 6; ++k) sx += xdense(Rp, k, r) * Y[6 * k];
                    IcIn[it] = sx;
                }
            }
        } else {
            LMH_FOR(e, 36) {
                int r = e / 6, cc = e % 6;
                double sum = K.Isp[e];
                for (int c = NCH - 1; c >= 0; --c) {
                    const double* Rp = Xl + 12 * chain_first(c);
                    const double* Y = Yt + 36 * c + cc;
                    for (int k = 0; k < 6; ++k) sum += xdense(Rp, k, r) * Y[6 * k];
                }
                w.M6[NQ * r + cc] = sum;   // Ic0
            }
        }
        cx.sync();
        cur ^= 1;
    }
}

// mass matrix entry M(6+a, 6+b)
LMH_HD double hjoint(const EnvWS& w, int a, int b) {
    int c = joint_chain(a);
    if (joint_chain(b) != c) return 0.0;
    return w.Hc[36 * c + 6 * (a - chain_joint0(c)) + (b - chain_joint0(c))];
}

#if defined(__CUDACC__)
struct EnvWS;
__device__ void jacobian_warp(EnvWS& w, int lane);
__device__ void wrench_solve_warp(EnvWS& w, int lane);
#endif

// ================================================================================================
// Phase 4: gravity bias by linearity, centroidal matrix (Dynamics.cpp:103-121), sole Jacobians
// (invKinematics.cpp:72-149), CoM velocity / angular momentum (Robot.cpp:300-310).
template <class Cx>
LMH_HDN void centroidal_jacobian_phase(const Cx& cx, EnvWS& w, const Consts& K) {
    const double* Xl = w.u.m.Xl;
    // C = Cg + M[:,3:6] * (R0' g e_z)   (RNEA is linear in the base acceleration; Dynamics.cpp:34-35,52)
    LMH_FOR(i, NQ) {
        double g0 = K.gravity * w.R0[6], g1 = K.gravity * w.R0[7], g2 = K.gravity * w.R0[8];
        double m0, m1, m2;
        if (i < 6) { m0 = w.M6[NQ * i + 3]; m1 = w.M6[NQ * i + 4]; m2 = w.M6[NQ * i + 5]; }
        else { m0 = w.M6[NQ * 3 + i]; m1 = w.M6[NQ * 4 + i]; m2 = w.M6[NQ * 5 + i]; }
        w.C[i] = w.Cg[i] + (m0 * g0 + m1 * g1 + m2 * g2);
    }
    // AG = X1G M6, AGpqp = X1G Cg[0:6], X1G = [[R0, -R0 p1G^],[0, R0]]
    LMH_FOR(j, NQ + 1) {
        double p1G[3] = {w.M6[NQ * 2 + 4] * K.inv_mass, w.M6[NQ * 0 + 5] * K.inv_mass, w.M6[NQ * 1 + 3] * K.inv_mass};
        double top[3], bot[3];
        if (j < NQ) for (int k = 0; k < 3; ++k) { top[k] = w.M6[NQ * k + j]; bot[k] = w.M6[NQ * (3 + k) + j]; }
        else for (int k = 0; k < 3; ++k) { top[k] = w.Cg[k]; bot[k] = w.Cg[3 + k]; }
        double u[3] = {top[0] - cross_k(p1G, bot, 0), top[1] - cross_k(p1G, bot, 1), top[2] - cross_k(p1G, bot, 2)};
        for (int k = 0; k < 3; ++k) {
            double a = rot_row(w.R0, k, u), b = rot_row(w.R0, k, bot);
            if (j < NQ) { w.AG[NQ * k + j] = a; w.AG[NQ * (3 + k) + j] = b; }
            else { w.AGpqp[k] = a; w.AGpqp[3 + k] = b; }
        }
    }
    bool jac_done = false;
#if defined(__CUDA_ARCH__)
    if (Cx::NT == 32) { if (!cx.helpers) jacobian_warp(w, cx.tid); jac_done = true; }
#endif
    if (!jac_done) {
    // running sole_X_j = (E, B) for both feet, starting from the sole frame
    LMH_FOR(it, 36) {
        int s = it / 18, e = it % 18;
        const double* Rp = Xl + 12 * (s == 0 ? 7 : 14);
        w.u.m.p.Xn[0][it] = e < 9 ? xdense(Rp, e / 3, e % 3) : xdense(Rp, 3 + (e - 9) / 3, (e - 9) % 3);
    }
    cx.sync();
    int cur = 0;
    for (int step = 0; step < 6; ++step) {   // j = ankle roll ... hip yaw-pitch
        LMH_FOR(it, 36) {
            int s = it / 18, e = it % 18;
            int j = (s == 0 ? 6 : 13) - step;
            const double* Xc = w.u.m.p.Xn[cur] + 18 * s;     // E (9), B (9)
            const double* Rp = Xl + 12 * j;
            int r = (e % 9) / 3, cc = e % 3;
            double x;
            if (e < 9) {     // E_new = E * Ej
                x = Xc[3 * r] * xdense(Rp, 0, cc) + Xc[3 * r + 1] * xdense(Rp, 1, cc) + Xc[3 * r + 2] * xdense(Rp, 2, cc);
            } else {         // B_new = B * Ej + E * Bj
                x = Xc[9 + 3 * r] * xdense(Rp, 0, cc) + Xc[9 + 3 * r + 1] * xdense(Rp, 1, cc) + Xc[9 + 3 * r + 2] * xdense(Rp, 2, cc) +
                    Xc[3 * r] * xdense(Rp, 3, cc) + Xc[3 * r + 1] * xdense(Rp, 4, cc) + Xc[3 * r + 2] * xdense(Rp, 5, cc);
            }
            w.u.m.p.Xn[cur ^ 1][it] = x;
        }
        // joint column of frame j = world-rotated column 2 of the current sole_X_(j)
        LMH_FOR(it, 12) {
            int s = it / 6, k = it % 6;
            const double* Xc = w.u.m.p.Xn[cur] + 18 * s;
            const double* col = k < 3 ? Xc : Xc + 9;
            double cv[3] = {col[2], col[5], col[8]};
            w.JFc[72 * s + 12 * k + 11 - step] = rot_row(w.Rs[s], k % 3, cv);
        }
        cx.sync();
        cur ^= 1;
    }
    // base block: rotated sole_X_base = [[E,0],[B,E]]
    LMH_FOR(it, 72) {
        int s = it / 36, k = (it % 36) / 6, b = it % 6;
        const double* Xc = w.u.m.p.Xn[cur] + 18 * s;
        double cv[3];
        if (k < 3) { if (b < 3) { cv[0] = Xc[b]; cv[1] = Xc[3 + b]; cv[2] = Xc[6 + b]; } else { cv[0] = cv[1] = cv[2] = 0.0; } }
        else { const double* S = b < 3 ? Xc + 9 : Xc; int bb = b % 3; cv[0] = S[bb]; cv[1] = S[3 + bb]; cv[2] = S[6 + bb]; }
        w.JFc[72 * s + 12 * k + b] = rot_row(w.Rs[s], k % 3, cv);
    }
    cx.sync();
    }
    // h = AG vb (current velocity)
#if defined(__CUDA_ARCH__)
    if (Cx::NT == 32) {
        // thirty lanes: lane 6 seg + k sums six terms of row k (a 30-term chain on six lanes was 3 % of the stage),
        // the five partial sums of a row meet through shared memory
        const int lane = cx.tid, k = lane % 6, seg = lane / 6;
        if (lane < 30) {
            double s = 0.0;
            LMH_UNROLL
            for (int j = 0; j < 6; ++j) s += w.AG[NQ * k + 6 * seg + j] * vb_at(w, 6 * seg + j);
            w.red[lane] = s;
        }
        cx.sync();
        if (lane < 6) {
            const double s = (((w.red[lane] + w.red[6 + lane]) + w.red[12 + lane]) + w.red[18 + lane]) + w.red[24 + lane];
            if (lane < 3) w.angMom[lane] = s; else w.comVel[lane - 3] = s * K.inv_mass;
        }
        cx.sync();
        return;
    }
#endif
    LMH_FOR(k, 6) {
        double s = 0.0;
        for (int j = 0; j < NQ; ++j) s += w.AG[NQ * k + j] * vb_at(w, j);
        if (k < 3) w.angMom[k] = s; else w.comVel[k - 3] = s * K.inv_mass;
    }
    cx.sync();
}

// ================================================================================================
// Phase 5: LIPM MPC as a precomputed gain (mpcLinearPendulum.cpp:78-109) + PD references
// (controller.cpp:296-386).
// The pieces of mpc_pd_phase.  In latency mode (rollout_kernel_helpers) the pieces that do not need the dynamics -- the ZMP
// window sums of the MPC, the joint-space PD and the foot references -- run on the helper warp that computed the sole
// Jacobians, while the environment's warp is still in the RNEA; the arithmetic is the same source either way.
LMH_HD int mpc_window_start(const Consts& K, const Plan& P, double t, int& kidx) {
    const int n1 = K.mpc_n1;
    kidx = (int)(t / K.mpc_dt);                      // mpcLinearPendulum.cpp:92 -- same FP64 expression
    int k0 = kidx;
    if (k0 < 0) k0 = 0;
    if (k0 + n1 > P.zmp_len) k0 = P.zmp_len - n1;    // the reference would read out of bounds here; we clamp
    return k0;
}
// sum_i r[i] zmp[k + i] for both axes as 2 x 16 partial sums in w.red; returns the window index k of this stage
template <class Cx>
LMH_HDN int mpc_window_sums(const Cx& cx, EnvWS& w, const Consts& K, const Plan& P, int plan_id) {
    const int n1 = K.mpc_n1;
    int kidx;
    const int k0 = mpc_window_start(K, P, w.t, kidx);
    LMH_FOR(it, 32) {
        int ax = it / 16, part = it % 16;
        const double* z = (ax == 0 ? P.zmp_x : P.zmp_y) + (size_t)plan_id * P.zmp_len + k0;
        double s = 0.0;
        const double* mr = cx.mpc_r ? cx.mpc_r : K.mpc_r;
        // the sixteen partial sums of an axis take the same number of turns (a short one adds 0 x z[0]): no divergent exit
        for (int i0 = 0; i0 < n1; i0 += 16) {
            const int i = i0 + part;
            const int ii = i < n1 ? i : n1 - 1;
            double r = mr[ii];
            r = i < n1 ? r : 0.0;
            s += r * z[ii];
        }
        w.red[it] = s;
    }
    return kidx;
}
// u0 and the reference CoM state of both axes (mpcLinearPendulum.cpp:89-101).  sums_done: the window sums were left
// in xRef[2] / yRef[2] (latency mode), else they are the partial sums in w.red
template <class Cx>
LMH_HDN void mpc_close(const Cx& cx, EnvWS& w, const Consts& K, int kidx, bool sums_done) {
    LMH_FOR(ax, 2) {
        double* ref = ax == 0 ? w.xRef : w.yRef;
        double s = 0.0;
        if (sums_done) s = ref[2];
        else for (int part = 0; part < 16; ++part) s += w.red[16 * ax + part];
        double p = w.com[ax], vv = w.comVel[ax];
        double u = s - (K.mpc_rPx[0] * p + K.mpc_rPx[1] * vv);
        ref[0] = K.mpc_A[0] * p + K.mpc_A[1] * vv + K.mpc_B[0] * u;
        ref[1] = K.mpc_A[2] * p + K.mpc_A[3] * vv + K.mpc_B[1] * u;
        ref[2] = u;
        if (ax == 0 && !sums_done) w.k_index = kidx;      // (latency mode: the helper warp has stored it)
    }
}
template <class Cx>
LMH_HDN void pd_joint_refs(const Cx& cx, EnvWS& w, const Consts& K) {
    // joint-space PD (controller.cpp:296-309): world-form errors, 3<->3 swap of the base slots
    LMH_FOR(i, NQ) {
        int src = i < 3 ? i + 3 : i < 6 ? i - 3 : i;
        w.qppRef[i] = K.KpJ * (K.qdes[src] - w.q[src]) + K.KdJ * (0.0 - w.v[src]);
    }
}
template <class Cx>
LMH_HDN void pd_momentum_refs(const Cx& cx, EnvWS& w, const Consts& K) {
    // centroidal momentum rate (controller.cpp:311-327)
    LMH_FOR(k, 6) {
        if (k < 3) w.hGpRef[k] = K.KdM * (0.0 - w.angMom[k]);
        else {
            int j = k - 3;
            double pr = j == 0 ? w.xRef[0] : j == 1 ? w.yRef[0] : K.mpc_zcom;
            double vr = j == 0 ? w.xRef[1] : j == 1 ? w.yRef[1] : 0.0;
            double ar = j == 0 ? w.xRef[2] : j == 1 ? w.yRef[2] : 0.0;
            w.hGpRef[k] = K.mass * (K.KpM * (pr - w.com[j]) + K.KdM * (vr - w.comVel[j]) + ar);
        }
    }
}
template <class Cx>
LMH_HDN void pd_foot_refs(const Cx& cx, EnvWS& w, const Consts& K, const Plan& P, int plan_id) {
    // walking extension: phase of this stage (integer work, same FP64 expression as the oracle)
    int ph = P.n_phases > 1 ? (int)(w.t / P.phase_T) : 0;      // (one phase: the clamp below gives 0 whatever t is; no division)
    ph = ph < 0 ? 0 : (ph >= P.n_phases ? P.n_phases - 1 : ph);
    const double tl = mul_then_add(w.t, -(double)ph, P.phase_T);
    LMH_FOR(i, 1) { w.phase = ph; w.contact = P.contact[(size_t)plan_id * P.n_phases + ph]; }
    // feet (controller.cpp:329-386)
    LMH_FOR(it, 12) {
        int s = it / 6, k = it % 6;
        double velk = 0.0;
        {
            const double* Jr = w.JFc + 72 * s + 12 * k;
            for (int j = 0; j < 6; ++j) velk += Jr[j] * w.vb6[j];
            for (int j = 0; j < 6; ++j) velk += Jr[6 + j] * w.v[6 + 6 * s + j];
        }
        double out;
        if (k < 3) {
            // e = -Rdes * log(Rdes' R)
            double E[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) E[3 * r + c] = K.Rdes[r] * w.Rs[s][c] + K.Rdes[3 + r] * w.Rs[s][3 + c] + K.Rdes[6 + r] * w.Rs[s][6 + c];
            double tr = E[0] + E[4] + E[8];
            double cth = fmax(-1.0, fmin(1.0, (tr - 1.0) / 2.0));
            double ax[3] = {E[7] - E[5], E[2] - E[6], E[3] - E[1]};
            double sc;
#if defined(__CUDA_ARCH__) && LMH_LOGMAP_SERIES
            // phi / (2 sin phi) as a function of x = 1 - cos phi, exactly: acos(1 - x) = sqrt(2x) P(x/2), sin phi = sqrt(2x)
            // sqrt(1 - x/2), P(y) = sum (2n)! / (4^n n!^2 (2n+1)) y^n (the arcsine series).  Below x = 1/64 (10 degrees: a
            // foot in or near contact) eight terms are exact to 3e-17 and replace acos, sin and a division -- library
            // routines with slow paths -- by a dozen FMAs; 1 - cth is exact there (Sterbenz).  Same function, same value to
            // the last bits; beyond that angle the library path is taken.
            const double xs = 1.0 - cth;
            if (xs < 1.0 / 64.0) {
                const double Ps = 1.0 + xs * (1.0 / 12 + xs * (3.0 / 160 + xs * (5.0 / 896 + xs * (35.0 / 18432 + xs * (63.0 / 90112
                                + xs * (231.0 / 851968 + xs * (143.0 / 1310720)))))));
                sc = xs < 5e-13 ? 0.5 : 0.5 * Ps * lmh_rsqrt(1.0 - 0.5 * xs);       // (phi < 1e-6: upstream's small-angle branch)
            } else
#endif
            {
                double phi = acos(cth);
                sc = phi < 1e-6 ? 0.5 : phi / (2.0 * sin(phi));
            }
            double l[3] = {sc * ax[0], sc * ax[1], sc * ax[2]};
            double eo = -rot_row(K.Rdes, k, l);
            out = K.KpF * eo + K.KdF * (0.0 - velk) + 0.0;
        } else {
            int j = k - 3;
            const double* cf = P.foot_coef + (((size_t)plan_id * P.n_phases + ph) * 2 + s) * 20 + (j == 0 ? 0 : j == 1 ? 6 : 12);
            int nc = j == 2 ? 8 : 6;
            // polyval of the polynomial, of its derivative and of its second derivative (generalizedFunctions.cpp:165-193) in
            // the same order of operations, with ONE chain of powers serving the three sums (upstream's three chains hold the
            // same products) and coefficients beyond the degree read as zero (adding +0 changes nothing): three independent
            // accumulations of fixed length instead of three dependent loops of variable length
            double pw[8], c[8];
            pw[0] = 1.0;
            LMH_UNROLL
            for (int i = 1; i < 8; ++i) pw[i] = pw[i - 1] * tl;
            LMH_UNROLL
            for (int i = 0; i < 8; ++i) c[i] = i < nc ? cf[i] : 0.0;
            double pos = 0.0, vr = 0.0, ar = 0.0;
            LMH_UNROLL
            for (int i = 0; i < 8; ++i) pos += c[i] * pw[i];
            LMH_UNROLL
            for (int i = 1; i < 8; ++i) vr += ((double)i * c[i]) * pw[i - 1];
            LMH_UNROLL
            for (int i = 2; i < 8; ++i) ar += ((double)(i - 1) * ((double)i * c[i])) * pw[i - 2];
            out = K.KpF * (pos - w.ps[s][j]) + K.KdF * (vr - velk) + ar;
        }
        w.footAccRef[it] = out;
    }
}
// what the helper warp of the latency mode computes after the sole Jacobians (see above)
template <class Cx>
LMH_HDN void mpc_pd_helper(const Cx& cx, EnvWS& w, const Consts& K, const Plan& P, int plan_id) {
    const int kidx = mpc_window_sums(cx, w, K, P, plan_id);
    cx.sync();
    LMH_FOR(ax, 2) {
        double s = 0.0;
        for (int part = 0; part < 16; ++part) s += w.red[16 * ax + part];
        (ax == 0 ? w.xRef : w.yRef)[2] = s;
        if (ax == 0) w.k_index = kidx;
    }
    pd_joint_refs(cx, w, K);
    pd_foot_refs(cx, w, K, P, plan_id);
    cx.sync();
}

template <class Cx>
LMH_HDN void mpc_pd_phase(const Cx& cx, EnvWS& w, const Consts& K, const Plan& P, int plan_id) {
    if (cx.helpers) {                  // latency mode: the helper warp has left the sums and the references (mpc_pd_helper)
        mpc_close(cx, w, K, 0, true);
        cx.sync();
        pd_momentum_refs(cx, w, K);
        cx.sync();
        return;
    }
    const int kidx = mpc_window_sums(cx, w, K, P, plan_id);
    cx.sync();
    mpc_close(cx, w, K, kidx, false);
    pd_joint_refs(cx, w, K);
    cx.sync();
    pd_momentum_refs(cx, w, K);
    pd_foot_refs(cx, w, K, P, plan_id);
    cx.sync();
}

// linear term of the QP (controller.cpp:125-132)
template <class Cx>
LMH_HDN void assemble_ga_phase(const Cx& cx, EnvWS& w, const Consts& K) {
    LMH_FOR(i, NQ) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += (k < 3 ? K.wCoMK : K.wCoML) * w.AG[NQ * k + i] * (w.AGpqp[k] - w.hGpRef[k]);
        double sj = 0.0;
        LMH_UNROLL
        for (int f = 0; f < 2; ++f) {          // branch-free: a coordinate a foot does not depend on contributes 0 x (...)
            const int ci = jf_col(f, i);
            const bool on = ci >= 0;
            LMH_UNROLL
            for (int k = 0; k < 6; ++k) {
                const double jv = on ? w.JFc[72 * f + 12 * k + (on ? ci : 0)] : 0.0;
                sj += jv * (w.Jpqp[6 * f + k] - w.footAccRef[6 * f + k]);
            }
        }
        const double wbp = K.wBasePos, wba = K.wBaseAng, wjn = K.wJoints;
        const double wj = i < 3 ? wbp : i < 6 ? wba : wjn;
        w.ga[i] = s - wj * w.qppRef[i] + K.wFoot * sj;
    }
    cx.sync();
}

// ================================================================================================
// Phase 6: QP data (controller.cpp:94-132).  H_a = Wj + JF' Wf JF + AG' Wc AG is never formed: with the angular
// momentum weight at zero (controller.hpp:119) the centroidal term is U'U with U = sqrt(wCoML) AG[3..5] (rank 3),
// and the rest, D = Wj + wFoot JF'JF, only couples the 18 base+leg coordinates (D1); on the 12 arm/head joints it
// is the constant wJoints.  Everything that needs H_a^-1 goes through
//     H_a^-1 = D^-1 - D^-1 U' (I3 + U D^-1 U')^-1 U D^-1                       (Woodbury)
// so only the 18 x 18 block is factorised.  Overwrites the model scratch (union).
LMH_HD double ha_entry(const EnvWS& w, const Consts& K, int i, int j) {      // H_a(i,j), for dumps and tests
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s += K.wCoMK * w.AG[NQ * k + i] * w.AG[NQ * k + j];
    for (int k = 3; k < 6; ++k) s += K.wCoML * w.AG[NQ * k + i] * w.AG[NQ * k + j];
    double sj = 0.0;
    for (int f = 0; f < 2; ++f) {
        int ci = jf_col(f, i), cj = jf_col(f, j);
        if (ci >= 0 && cj >= 0)
            for (int k = 0; k < 6; ++k) sj += w.JFc[72 * f + 12 * k + ci] * w.JFc[72 * f + 12 * k + cj];
    }
    s += K.wFoot * sj;
    if (i == j) s += i < 3 ? K.wBasePos : i < 6 ? K.wBaseAng : K.wJoints;
    return s;
}
// entry (r, j) of the right-hand-side block [M6; ga'; U] before the factorisation
LMH_HD double yt_entry(const EnvWS& w, const Consts& K, int r, int j) {
    double v = r < 6 ? w.M6[NQ * r + j] : r == 6 ? w.ga[j] : K.sq_wCoML * w.AG[NQ * (r - 4) + j];
    return j < ND1 ? v : v * K.isq_wJoints;
}
template <class Cx>
LMH_HDN void assemble_phase(const Cx& cx, EnvWS& w, const Consts& K) {
    assemble_ga_phase(cx, w, K);        // (ends with a sync; the union is free from here on)
    double* L1 = w.u.s.L1;
    LMH_FOR(it, ND1 * ND1) {
        int i = it / ND1, j = it % ND1;
        if (j <= i) {
            double sj = 0.0;
            for (int f = 0; f < 2; ++f) {
                int ci = jf_col(f, i), cj = jf_col(f, j);
                if (ci >= 0 && cj >= 0)
                    for (int k = 0; k < 6; ++k) sj += w.JFc[72 * f + 12 * k + ci] * w.JFc[72 * f + 12 * k + cj];
            }
            double v = K.wFoot * sj;
            if (i == j) v += i < 3 ? K.wBasePos : i < 6 ? K.wBaseAng : K.wJoints;
            L1[LD1 * i + j] = v;
        }
    }
    LMH_FOR(it, NYR * NQ) w.u.s.Yt[it] = yt_entry(w, K, it / NQ, it % NQ);
    cx.sync();
}

#if defined(__CUDACC__)
__device__ void qp_factor_warp(EnvWS& w, const Consts& K, int lane);
__device__ void jacobian_warp(EnvWS& w, int lane);
__device__ void qp_backsolve_warp(EnvWS& w, int lane);
#endif

// The general primal active-set iteration of the WBC QP (see qp_phase) and the multiplier of its solution, kept apart
// from the hot path of qp_phase (it runs in 2 % of the stages of a balancing robot).  In: the projected starting point / warm-started working set in w.c, w.free_mask;
// out: w.c, w.f12 = G c, w.lam = Sa^-1 (rhs - Jb'f12), iteration count and status.
#if defined(__CUDACC__) && LMH_QP_GENERAL_NOINLINE
#define LMH_NOINLINE __noinline__
#else
#define LMH_NOINLINE
#endif
template <class Cx>
LMH_NOINLINE LMH_HDN void qp_general_active_set(const Cx& cx, EnvWS& w, const Consts& K, int contact, bool have_point, int& iters, int& status) {
#if LMH_QP_TABLE
    // One pass = the equality-constrained problem on the current working set F.  Everything that depends on F alone comes
    // from the table (Consts::qp_tab, one row per foot); with g = Jb lam the target in wrench space,
    //     f = Kf g,   (Sa + Jb' Kf Jb) lam = rhs,   c_F = G_F' Yf g,   mu_i = g_i'(wF f - g) on the clipped coefficients:
    // the same 6 x 6 system as the all-free shortcut of qp_phase, whose Dinv and Pinv Dinv are the row of the full set.
    double* KY = w.u.s.Ay;            // [2][72] : Kf (36), Yf (36) of each foot's working set
    double* U = w.u.s.Wb;             // [2][6][6] : Kf Jb_f
    double* S6 = w.u.s.Pm;            // [6][6]    : lower triangle of Sa + sum_f Jb_f' Kf Jb_f
    double* vv = w.u.s.vv;            // [36]      : g (12), y = Yf g (12), r = wF f - g (12)
    bool done = false;
    for (;;) {
        ++iters;
        {
            double lam[6] = {0, 0, 0, 0, 0, 0};
            unsigned m0 = 0u, m1 = 0u;
#if defined(__CUDA_ARCH__)
            if (Cx::NT == 32) {
                const unsigned b = __ballot_sync(0xffffffffu, w.free_mask[cx.tid] != 0);
                m0 = b & 0xffffu; m1 = b >> 16;
            } else
#endif
            {
                for (int g = 0; g < 16; ++g) { m0 |= w.free_mask[g] ? 1u << g : 0u; m1 |= w.free_mask[16 + g] ? 1u << g : 0u; }
            }
            const double* r0 = K.qp_tab + (size_t)m0 * 72;
            const double* r1 = K.qp_tab + (size_t)m1 * 72;
            LMH_FOR(it, 144) KY[it] = it < 72 ? r0[it] : r1[it - 72];
            cx.sync();
            LMH_FOR(it, 72) {
                const int f = it / 36, i = (it % 36) / 6, sidx = it % 6;
                double sd = 0.0;
                for (int l = 0; l < 6; ++l) sd += KY[72 * f + 6 * i + l] * w.JFc[72 * f + 12 * l + sidx];
                U[it] = sd;
            }
            cx.sync();
            LMH_FOR(it, 36) {
                const int r = it / 6, sidx = it % 6;
                if (sidx <= r) {
                    double sd = w.Sa[it];
                    for (int k = 0; k < 12; ++k) sd += w.JFc[72 * (k / 6) + 12 * (k % 6) + r] * U[36 * (k / 6) + 6 * (k % 6) + sidx];
                    S6[it] = sd;
                }
            }
            cx.sync();
            {
                double A[36], Ainv[6];
                LMH_UNROLL
                for (int e = 0; e < 36; ++e) A[e] = (e % 6 <= e / 6) ? S6[e] : 0.0;
                chol6(A, Ainv);
                LMH_UNROLL
                for (int r = 0; r < 6; ++r) {
                    double sd = w.rhs[r];
                    LMH_UNROLL
                    for (int m = 0; m < 6; ++m) if (m < r) sd -= A[6 * r + m] * lam[m];
                    lam[r] = sd * Ainv[r];
                }
                LMH_UNROLL
                for (int r = 5; r >= 0; --r) {
                    double sd = lam[r];
                    LMH_UNROLL
                    for (int m = 0; m < 6; ++m) if (m > r) sd -= A[6 * m + r] * lam[m];
                    lam[r] = sd * Ainv[r];
                }
            }
            LMH_FOR(i, 1) {                       // every thread holds the same six values; the caller reads them from here
                LMH_UNROLL
                for (int r = 0; r < 6; ++r) w.lam[r] = lam[r];
            }
            LMH_FOR(it, 24) {
                if (it < 12) {            // f = Kf Jb lam
                    double sd = 0.0;
                    for (int q = 0; q < 6; ++q) sd += U[6 * it + q] * lam[q];
                    w.f12[it] = sd;
                } else {                  // g = Jb lam
                    const int k = it - 12;
                    double sd = 0.0;
                    for (int q = 0; q < 6; ++q) sd += w.JFc[72 * (k / 6) + 12 * (k % 6) + q] * lam[q];
                    vv[k] = sd;
                }
            }
            cx.sync();
            LMH_FOR(it, 24) {
                if (it < 12) {            // y = Yf g
                    const int f = it / 6, k = it % 6;
                    double sy = 0.0;
                    for (int l = 0; l < 6; ++l) sy += KY[72 * f + 36 + 6 * k + l] * vv[6 * f + l];
                    vv[12 + it] = sy;
                } else {                  // r = wF f - g
                    const int k = it - 12;
                    vv[24 + k] = K.wForce * w.f12[k] - vv[k];
                }
            }
            cx.sync();
        }
#else
    double* Bm = w.u.s.Bm;            // [12][6] : row 6f+i, basis column k of foot f
    double* Pm = w.u.s.Pm;            // [2][36] : residual of the pivoted Cholesky, then T = B'B
    double* Wb = w.u.s.Wb;            // [6][12] : Ls^-1 Jb' B
    double* Ay = w.u.s.Ay;            // 13 x 13 : lower triangle factorised in place, row 12 = rhs / zeta
    double* vv = w.u.s.vv;            // [36]   : (12 unused), y with B'y = zeta (12), r = wF f - Jb lam (12)
    // Ls = chol(Sa): 6x6, every thread redundantly in registers, then parked in shared memory (general iteration only)
    {
        double A[36];
        LMH_UNROLL
        for (int e = 0; e < 36; ++e) A[e] = w.Sa[e];
        double Ainv[6];
        chol6(A, Ainv);
        // every thread holds the same factor: all of them store it (same value to the same word), which costs 27 plain
        // stores instead of 27 compare-and-branch pairs; static indices keep A in registers
        LMH_UNROLL
        for (int e = 0; e < 36; ++e) if (e % 6 <= e / 6) w.Ls[e] = A[e];
        LMH_UNROLL
        for (int e = 0; e < 6; ++e) w.Lsi[e] = Ainv[e];
    }
    const double* Ls = w.Ls;
    const double* Lsi = w.Lsi;
    bool first_pass = true;            // of the general iteration in this stage
    for (;;) {
        ++iters;
        {
            LMH_FOR(it, 72) {         // P_F per foot
                const int f = it / 36, i = (it % 36) / 6, j = it % 6;
                const double* gi = K.Gf + 16 * i;
                const double* gj = K.Gf + 16 * j;
                double pij = 0.0;
                for (int g = 0; g < 16; ++g) if (w.free_mask[16 * f + g]) pij += gi[g] * gj[g];
                Pm[it] = pij;
                Bm[it] = 0.0;
            }
            LMH_FOR(i, 12) w.pdone[i] = 0;
            cx.sync();
            double dmax0[2] = {0.0, 0.0};
            for (int f = 0; f < 2; ++f) for (int i = 0; i < 6; ++i) dmax0[f] = fmax(dmax0[f], Pm[36 * f + 7 * i]);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int k = 0; k < 6; ++k) {      // one pivot per foot and step; rows 6f+i are the work items
                LMH_FOR(it, 12) {
                    const int f = it / 6, i = it % 6;
                    const double* S = Pm + 36 * f;
                    int p = -1;
                    double best = 1e-10 * dmax0[f];
                    for (int j = 0; j < 6; ++j) if (!w.pdone[6 * f + j] && S[7 * j] > best) { best = S[7 * j]; p = j; }
                    double colv = 0.0;
                    const double rs = lmh_rsqrt(p >= 0 ? best : 1.0);
                    if (p >= 0 && !w.pdone[it]) colv = p == i ? best * rs : S[6 * i + p] * rs;
                    Bm[6 * it + k] = colv;
                    if (i == 0) { w.ppiv[f] = p; w.pord[6 * f + k] = p; w.piv[16 + 6 * f + k] = rs; }   // piv[16..27]: 1 / B[p_k][k]
                }
                cx.sync();
                LMH_FOR(it, 12) {
                    const int f = it / 6, i = it % 6;
                    double* S = Pm + 36 * f;
                    const int p = w.ppiv[f];
                    if (p >= 0) {
                        const double colv = Bm[6 * it + k];
                        for (int j = 0; j < 6; ++j) S[6 * i + j] -= colv * Bm[6 * (6 * f + j) + k];
                        if (i == p) w.pdone[it] = 1;          // nobody reads the flags in this pass
                    }
                }
                cx.sync();
            }
            LMH_FOR(it, 72) {         // T = B'B per foot (into Pm)
                const int f = it / 36, k = (it % 36) / 6, l = it % 6;
                double t = 0.0;
                for (int i = 0; i < 6; ++i) t += Bm[6 * (6 * f + i) + k] * Bm[6 * (6 * f + i) + l];
                Pm[it] = t;
            }
            cx.sync();
        }
        LMH_FOR(col, 13) {            // W[:, col] = Ls^-1 Jb' B[:, col] ; column 12: sv = Ls^-1 rhs (first pass of a stage only)
            if (col < 12 || first_pass) {
                const int cc = col < 12 ? col : 0, f = cc / 6, k = cc % 6;
                double t[6];
                for (int m = 0; m < 6; ++m) {
                    double sd = 0.0;
                    for (int i = 0; i < 6; ++i) sd += w.JFc[72 * f + 12 * i + m] * Bm[6 * (6 * f + i) + k];
                    t[m] = col < 12 ? sd : w.rhs[m];
                }
                for (int m = 0; m < 6; ++m) {
                    double sd = t[m];
                    for (int n = 0; n < m; ++n) sd -= Ls[6 * m + n] * t[n];
                    t[m] = sd * Lsi[m];
                    if (col < 12) Wb[12 * m + col] = t[m]; else w.sv[m] = t[m];
                }
            }
        }
        cx.sync();
        LMH_FOR(it, 13 * 12) {        // A = wF T + W'W + eps I ; rhs row = W' sv
            const int i = it / 12, j = it % 12;
            double sd = 0.0;
            if (i < 12) {
                if (j <= i) {
                    for (int m = 0; m < 6; ++m) sd += Wb[12 * m + i] * Wb[12 * m + j];
                    if (i / 6 == j / 6) sd += K.wForce * Pm[36 * (i / 6) + 6 * (i % 6) + (j % 6)];
                    if (i == j) sd += K.epsC;
                    Ay[13 * i + j] = sd;
                }
            } else {
                for (int m = 0; m < 6; ++m) sd += Wb[12 * m + j] * w.sv[m];
                Ay[13 * 12 + j] = sd;
            }
        }
        cx.sync();
        bool solved = false;
#if defined(__CUDA_ARCH__)
        if (Cx::NT == 32 && LMH_WARP_WRENCH) { wrench_solve_warp(w, cx.tid); solved = true; }
#endif
        if (!solved) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int k = 0; k < 12; ++k) {
            LMH_FOR(i0, 13 - k) {
                int i = k + i0;
                double* ri = Ay + 13 * i;
                const double* rk = Ay + 13 * k;
                double s0 = ri[k];
                for (int m = 0; m < k; ++m) s0 -= ri[m] * rk[m];
                ri[k] = s0;
            }
            cx.sync();
            double pv = Ay[13 * k + k];
            double dk = sqrt(pv > 0.0 ? pv : 1e-300);
            LMH_FOR(i0, 13 - k) {
                int i = k + i0;
                if (i == k) w.piv[k] = dk;
                else Ay[13 * i + k] = Ay[13 * i + k] / dk;
            }
            cx.sync();
        }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int k = 11; k >= 0; --k) {        // L' zeta = z (row 12 of Ay), column oriented
            double xk = Ay[13 * 12 + k] / w.piv[k];
            cx.sync();
            LMH_FOR(i, k + 1) {
                if (i == k) Ay[13 * 12 + k] = xk;
                else Ay[13 * 12 + i] -= Ay[13 * k + i] * xk;
            }
            cx.sync();
        }
        }
        const double* zeta = Ay + 13 * 12;
        {
        LMH_FOR(it, 14) {
            if (it < 12) {            // f = B zeta
                const int f = it / 6;
                double sd = 0.0;
                for (int k = 0; k < 6; ++k) sd += Bm[6 * it + k] * zeta[6 * f + k];
                w.f12[it] = sd;
            } else {
                // y with B'y = zeta, for c_F = G_F'y.  Any solution does (they differ by null(B') = null(G_F')), and in
                // pivot order B' is upper triangular -- column k is zero on the rows pivoted before it -- so one is had
                // by back substitution with y = 0 on the rows that were never pivoted (rank-deficient contact sets).
                const int f = it - 12;
                double* y = vv + 12 + 6 * f;
                for (int i = 0; i < 6; ++i) y[i] = 0.0;
                for (int k = 5; k >= 0; --k) {
                    const int p = w.pord[6 * f + k];
                    if (p >= 0) {
                        double sd = zeta[6 * f + k];
                        for (int i = 0; i < 6; ++i) sd -= Bm[6 * (6 * f + i) + k] * y[i];      // y[p] is still zero here
                        y[p] = sd * w.piv[16 + 6 * f + k];
                    }
                }
            }
        }
        cx.sync();
        {
            // lam = Sa^-1 (rhs - Jb'f) in every thread; r = wF f - Jb lam
            double l[6];
            LMH_UNROLL
            for (int r = 0; r < 6; ++r) {
                double sd = w.rhs[r];
                LMH_UNROLL
                for (int k = 0; k < 12; ++k) sd -= w.JFc[72 * (k / 6) + 12 * (k % 6) + r] * w.f12[k];
                l[r] = sd;
            }
            LMH_UNROLL
            for (int r = 0; r < 6; ++r) {
                double sd = l[r];
                LMH_UNROLL
                for (int m = 0; m < r; ++m) sd -= Ls[6 * r + m] * l[m];
                l[r] = sd * Lsi[r];
            }
            LMH_UNROLL
            for (int r = 5; r >= 0; --r) {
                double sd = l[r];
                LMH_UNROLL
                for (int m = r + 1; m < 6; ++m) sd -= Ls[6 * m + r] * l[m];
                l[r] = sd * Lsi[r];
            }
            LMH_FOR(it, 12) {
                const int f = it / 6;
                double sj = 0.0;
                for (int r = 0; r < 6; ++r) sj += w.JFc[72 * f + 12 * (it % 6) + r] * l[r];
                vv[24 + it] = K.wForce * w.f12[it] - sj;
            }
        }
        cx.sync();
        }
#endif
        LMH_FOR(i, NC) {
            const int foot = i / 16, gi = i % 16;
            double sc = 0.0, sm = 0.0;
            for (int k = 0; k < 6; ++k) { sc += K.Gf[16 * k + gi] * vv[12 + 6 * foot + k]; sm += K.Gf[16 * k + gi] * vv[24 + 6 * foot + k]; }
            w.cs[i] = w.free_mask[i] ? sc : 0.0;
            w.mu[i] = sm;                              // multiplier of c_i >= 0 (meaningful on the working set)
        }
        cx.sync();
        // decisions: every thread reaches the same ones (warp reductions for one warp per env, else redundant loops)
#if LMH_QP_TABLE
        done = false;
#else
        bool done = false;
#endif
        int change = -1, to_free = 0;
        double alpha = 1.0;
        // multipliers come out of the cancellation wF f - Jb lam: anything below the round-off of those two terms is
        // zero (such coefficients only move c inside the null space of G, which no output depends on)
#if defined(__CUDA_ARCH__)
        if (Cx::NT == 32) {
            const int i = cx.tid;
            const bool fr = w.free_mask[i] != 0;
            const double csi = w.cs[i], mui = w.mu[i], ci = w.c[i];
            const double cmin = warp_min(fr && csi < 0.0 ? csi : 0.0);
            if (!have_point) {
                have_point = true;
                if (cmin >= 0.0) done = true; else alpha = -1.0;
            } else if (cmin >= 0.0) {
                const double mu_tol = -2e-10 * warp_max(i < 12 ? fmax(1.0, fabs(K.wForce * w.f12[i]) + fabs(vv[24 + i])) : 1.0);
                const bool cand = !fr && ((contact >> (i / 16)) & 1) && mui < mu_tol;
                const double vmin = warp_min(cand ? mui : 1e300);
                const unsigned hit = __ballot_sync(0xffffffffu, cand && mui == vmin);
                if (hit == 0u) done = true; else { change = __ffs(hit) - 1; to_free = 1; }
            } else {
                const bool cand = fr && csi < 0.0;
                const double dl = ci - csi;
                const double ri = cand ? (dl > 0.0 ? ci / dl : 0.0) : 2.0;
                const double best = warp_min(ri);
                const unsigned hit = __ballot_sync(0xffffffffu, cand && ri == best);
                if (hit != 0u) change = __ffs(hit) - 1;
                alpha = best < 1.0 ? best : 1.0;
            }
        } else
#endif
        {
        double rmax = 1.0;
        for (int k = 0; k < 12; ++k) rmax = fmax(rmax, fabs(K.wForce * w.f12[k]) + fabs(vv[24 + k]));
        const double mu_tol = -2e-10 * rmax;
        double cmin = 0.0;
        for (int i = 0; i < NC; ++i) if (w.free_mask[i] && w.cs[i] < cmin) cmin = w.cs[i];
        if (!have_point) {
            have_point = true;
            if (cmin >= 0.0) done = true;              // the all-free minimiser is feasible: optimal
            else alpha = -1.0;                         // project onto the orthant; working set = the clipped coefficients
        } else if (cmin >= 0.0) {
            double worst = mu_tol;
            for (int i = 0; i < NC; ++i) if (!w.free_mask[i] && ((contact >> (i / 16)) & 1) && w.mu[i] < worst) { worst = w.mu[i]; change = i; }
            if (change < 0) done = true; else to_free = 1;
        } else {
            double best = 2.0;
            for (int i = 0; i < NC; ++i)
                if (w.free_mask[i] && w.cs[i] < 0.0) {
                    double dl = w.c[i] - w.cs[i];
                    double r = dl > 0.0 ? w.c[i] / dl : 0.0;
                    if (r < best) { best = r; change = i; }
                }
            alpha = best < 1.0 ? best : 1.0;
        }
        }
#if defined(LMH_DEBUG_QP) && !defined(__CUDA_ARCH__)
        if (iters < 8) {
            int nf = 0; double cm = 0; for (int i = 0; i < NC; ++i) { nf += w.free_mask[i]; if (w.free_mask[i] && w.cs[i] < cm) cm = w.cs[i]; }
            printf("it %d nfree %d cmin %.3e alpha %.3e change %d to_free %d done %d mu[change] %.3e\n", iters, nf, cm, alpha, change, to_free, (int)done, change >= 0 ? w.mu[change] : 0.0);
        }
#endif
        cx.sync();
        LMH_FOR(i, NC) {
            if (alpha < 0.0) {
                double x = w.cs[i];
                w.c[i] = (x > 0.0 && ((contact >> (i / 16)) & 1)) ? x : 0.0;
                w.free_mask[i] = (x > 0.0 && ((contact >> (i / 16)) & 1)) ? 1 : 0;
            } else {
                double x = w.c[i] + alpha * (w.cs[i] - w.c[i]);
                if (!w.free_mask[i]) x = 0.0;
                if (i == change) {
                    if (to_free) w.free_mask[i] = 1;
                    else { w.free_mask[i] = 0; x = 0.0; }
                }
                w.c[i] = x > 0.0 ? x : 0.0;
            }
        }
        cx.sync();
#if !LMH_QP_TABLE
        first_pass = false;
#endif
        if (done) break;
        if (iters >= K.qp_max_iter) { status = 1; break; }
    }
#if LMH_QP_TABLE
    // the last pass took the full step: c = c*, and f12 = Kf g, lam are those of that pass (Sa lam = rhs - Jb' f12 holds by
    // construction).  Only an iteration that ran out of passes needs them rebuilt from the point it stopped at.
    if (done) return;
    {
        double A[36], Ainv[6];
        LMH_UNROLL
        for (int e = 0; e < 36; ++e) A[e] = w.Sa[e];
        chol6(A, Ainv);
        LMH_UNROLL
        for (int e = 0; e < 36; ++e) if (e % 6 <= e / 6) w.Ls[e] = A[e];
        LMH_UNROLL
        for (int e = 0; e < 6; ++e) w.Lsi[e] = Ainv[e];
        cx.sync();
    }
    const double* Ls = w.Ls;
    const double* Lsi = w.Lsi;
#endif
    // f12 = G c ; lam = Sa^-1 (rhs - Jb' f12)
    LMH_FOR(r, 12) {
        int foot = r / 6, k = r % 6;
        double s = 0.0;
        for (int gi = 0; gi < 16; ++gi) s += K.Gf[16 * k + gi] * w.c[16 * foot + gi];
        w.f12[r] = s;
    }
    cx.sync();
    double l[6] = {0, 0, 0, 0, 0, 0};
#if defined(__CUDA_ARCH__)
    if (Cx::NT == 32) {       // lane r < 6 forms its own entry, then every lane collects the six
        const int r = cx.tid < 6 ? cx.tid : 0;
        double s = w.rhs[r];
        LMH_UNROLL
        for (int k = 0; k < 12; ++k) s -= w.JFc[72 * (k / 6) + 12 * (k % 6) + r] * w.f12[k];
        LMH_UNROLL
        for (int q = 0; q < 6; ++q) l[q] = __shfl_sync(0xffffffffu, s, q);
    } else
#endif
    {
        LMH_UNROLL
        for (int r = 0; r < 6; ++r) {
            double s = w.rhs[r];
            LMH_UNROLL
            for (int k = 0; k < 12; ++k) s -= w.JFc[72 * (k / 6) + 12 * (k % 6) + r] * w.f12[k];
            l[r] = s;
        }
    }
    LMH_UNROLL
    for (int r = 0; r < 6; ++r) {
        double s = l[r];
        LMH_UNROLL
        for (int m = 0; m < r; ++m) s -= Ls[6 * r + m] * l[m];
        l[r] = s * Lsi[r];
    }
    LMH_UNROLL
    for (int r = 5; r >= 0; --r) {
        double s = l[r];
        LMH_UNROLL
        for (int m = r + 1; m < 6; ++m) s -= Ls[6 * m + r] * l[m];
        l[r] = s * Lsi[r];
    }
    LMH_FOR(i, 1) {
        LMH_UNROLL
        for (int r = 0; r < 6; ++r) w.lam[r] = l[r];
    }
    cx.sync();
}

// ================================================================================================
// Phase 7: the WBC QP (controller.cpp:388-479) in reduced form.
//   variables a (30), c (32);  f12 = G c eliminated (controller.cpp:156-294 rows 6..17)
//   min 1/2 a'Ha a + ga'a + 1/2 wF |G c|^2 + 1/2 eps |c|^2   s.t.  M6 a - Jb' G c = -C6,  c >= 0
// a is eliminated through the Cholesky factor of Ha and the 6x6 Schur complement Sa = M6 Ha^-1 M6',
// leaving   min 1/2 c'Q c - p'c, c >= 0,  Q = wF G'G + (Ls^-1 Jb'G)'(Ls^-1 Jb'G) + eps I,
// which a primal active-set method solves exactly (masked Cholesky of Q per working set).
template <class Cx>
LMH_HDN void qp_phase(const Cx& cx, EnvWS& w, const Consts& K, bool factor_done) {
    double* L1 = w.u.s.L1;
    double* Yt = w.u.s.Yt;
    double* Gm = w.u.s.Gm;
    if (!factor_done) {
        // left-looking Cholesky of D1 (18 columns); the 10 right-hand-side rows ride along and come out as L1^-1 [.]
        for (int k = 0; k < ND1; ++k) {
            LMH_FOR(i0, ND1 + NYR - k) {
                int i = k + i0;
                double* ri = i < ND1 ? L1 + LD1 * i : Yt + NQ * (i - ND1);
                const double* rk = L1 + LD1 * k;
                double s0 = ri[k], s1 = 0.0;
                int m = 0;
                for (; m + 1 < k; m += 2) { s0 -= ri[m] * rk[m]; s1 -= ri[m + 1] * rk[m + 1]; }
                if (m < k) s0 -= ri[m] * rk[m];
                ri[k] = s0 + s1;
            }
            cx.sync();
            double dk = sqrt(L1[LD1 * k + k]);     // nobody writes the pivot in the next loop
            LMH_FOR(i0, ND1 + NYR - k) {
                int i = k + i0;
                if (i == k) w.ladiag[k] = dk;
                else {
                    double* ri = i < ND1 ? L1 + LD1 * i : Yt + NQ * (i - ND1);
                    ri[k] = ri[k] / dk;
                }
            }
            cx.sync();
        }
    }
    // Gram matrix of the rows: Gm[r][c] = [.]_r D^-1 [.]_c'
#if defined(__CUDA_ARCH__) && LMH_GRAM_DMMA
    if (Cx::NT == 32) {
        // 10 x 30 times 30 x 10 on the FP64 tensor cores: three 8 x 8 tiles (the lower triangle of the 16 x 16 padded
        // product), eight k-steps of mma.sync m8n8k4 each.  Yt serves as A (row-major) and, read the same way, as B (= Yt').
        const int lane = cx.tid, fr = lane >> 2, fk = lane & 3;
        double t00a = 0, t00b = 0, t10a = 0, t10b = 0, t11a = 0, t11b = 0;
        LMH_UNROLL
        for (int kt = 0; kt < 8; ++kt) {
            const int m = 4 * kt + fk;
            const double y0 = m < NQ ? Yt[NQ * fr + m] : 0.0;
            const double y1 = (m < NQ && fr < NYR - 8) ? Yt[NQ * (8 + fr) + m] : 0.0;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(t00a), "+d"(t00b) : "d"(y0), "d"(y0));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(t10a), "+d"(t10b) : "d"(y1), "d"(y0));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(t11a), "+d"(t11b) : "d"(y1), "d"(y1));
        }
        const int c0 = 2 * fk;
        Gm[NYR * fr + c0] = t00a; Gm[NYR * fr + c0 + 1] = t00b;
        if (fr < NYR - 8) {
            Gm[NYR * (8 + fr) + c0] = t10a; Gm[NYR * (8 + fr) + c0 + 1] = t10b;
            Gm[NYR * c0 + 8 + fr] = t10a; Gm[NYR * (c0 + 1) + 8 + fr] = t10b;
            if (fk == 0) { Gm[NYR * (8 + fr) + 8] = t11a; Gm[NYR * (8 + fr) + 9] = t11b; }
        }
    } else
#endif
    LMH_FOR(it, NYR * (NYR + 1) / 2) {
        // unrank the lower-triangular index (55 entries): 8 it + 1 = (2r + 1)^2 at the start of row r, exact in float
        const int r = (int)((sqrtf(8.0f * (float)it + 1.0f) - 1.0f) * 0.5f);
        const int c = it - r * (r + 1) / 2;
        {
            double s0 = 0.0, s1 = 0.0;
            const double* yr = Yt + NQ * r;
            const double* yc = Yt + NQ * c;
            LMH_UNROLL
            for (int m = 0; m < NQ; m += 2) {
#if defined(__CUDA_ARCH__)
                const double2 a = *reinterpret_cast<const double2*>(yr + m), b = *reinterpret_cast<const double2*>(yc + m);
                s0 += a.x * b.x; s1 += a.y * b.y;
#else
                s0 += yr[m] * yc[m]; s1 += yr[m + 1] * yc[m + 1];
#endif
            }
            Gm[NYR * r + c] = s0 + s1;
            Gm[NYR * c + r] = s0 + s1;
        }
    }
    cx.sync();
    // Ci = (I3 + U D^-1 U')^-1 in every thread; Sa = M6 Ha^-1 M6', rhs = C6 - M6 Ha^-1 ga
    double Ci[9];
    {
        double c00 = 1.0 + Gm[NYR * 7 + 7], c01 = Gm[NYR * 7 + 8], c02 = Gm[NYR * 7 + 9];
        double c11 = 1.0 + Gm[NYR * 8 + 8], c12 = Gm[NYR * 8 + 9], c22 = 1.0 + Gm[NYR * 9 + 9];
        double a00 = c11 * c22 - c12 * c12, a01 = c02 * c12 - c01 * c22, a02 = c01 * c12 - c02 * c11;
        double a11 = c00 * c22 - c02 * c02, a12 = c01 * c02 - c00 * c12, a22 = c00 * c11 - c01 * c01;
        double det = c00 * a00 + c01 * a01 + c02 * a02;
        const double rdet = 1.0 / det;
        Ci[0] = a00 * rdet; Ci[1] = a01 * rdet; Ci[2] = a02 * rdet;
        Ci[3] = Ci[1];      Ci[4] = a11 * rdet; Ci[5] = a12 * rdet;
        Ci[6] = Ci[2];      Ci[7] = Ci[5];      Ci[8] = a22 * rdet;
    }
    LMH_FOR(it, 42) {
        int r = it / 6, c = it % 6;          // r == 6: the ga row
        const int rr = r < 6 ? r : c, cc = r < 6 ? c : 6;
        double g0 = Gm[NYR * rr + 7], g1 = Gm[NYR * rr + 8], g2 = Gm[NYR * rr + 9];
        double h0 = Gm[NYR * 7 + cc], h1 = Gm[NYR * 8 + cc], h2 = Gm[NYR * 9 + cc];
        double t0 = Ci[0] * h0 + Ci[1] * h1 + Ci[2] * h2, t1 = Ci[3] * h0 + Ci[4] * h1 + Ci[5] * h2, t2 = Ci[6] * h0 + Ci[7] * h1 + Ci[8] * h2;
        double v = Gm[NYR * rr + cc] - (g0 * t0 + g1 * t1 + g2 * t2);
        if (r < 6) w.Sa[it] = v; else w.rhs[c] = w.C[c] - v;
    }
    cx.sync();
    // ---- primal active set on c >= 0, every working set solved in the wrench space of the feet.
    // For a free set F let P_F = G_F G_F' = B B' (B: 12 x 12, block diagonal per foot, from a diagonally pivoted
    // Cholesky that stops at the numerical rank: columns beyond the rank are zero).  With f = B zeta the
    // equality-constrained problem on F is
    //     (B'K B + eps I) zeta = B'd,   K = wF I + Jb Sa^-1 Jb',  d = Jb Sa^-1 rhs,
    // c_F = G_F' B (B'B)^-1 zeta, and the multipliers of the clipped coefficients are mu_i = g_i'(wF f - Jb lam),
    // lam = Sa^-1 (rhs - Jb'f).  Well conditioned for ANY contact pattern (a foot on an edge, on one vertex, or
    // unloaded only shortens B), which a c-space or a full 12 x 12 wrench-space system is not.
    double* Bm = w.u.s.Bm;            // [12][6] : row 6f+i, basis column k of foot f
    double* Pm = w.u.s.Pm;            // [2][36] : residual of the pivoted Cholesky, then T = B'B
    double* Wb = w.u.s.Wb;            // [6][12] : Ls^-1 Jb' B
    double* Ay = w.u.s.Ay;            // 13 x 13 : lower triangle factorised in place, row 12 = rhs / zeta
    double* vv = w.u.s.vv;            // [36]   : (12 unused), y with B'y = zeta (12), r = wF f - Jb lam (12)
    const int contact = w.contact;     // bit f: foot f may carry force (both set in the reference's stand task)
    // Warm start: consecutive stage evaluations of an environment nearly always share their working set, and any
    // c >= 0 is a feasible starting point, so the previous solution (w.c, w.free_mask) seeds the iteration when it
    // has clipped coefficients and the contact phase is the same.  The minimiser is unique: only the path changes.
    bool have_point = false;
    if (w.warm_contact == contact) {
#if defined(__CUDA_ARCH__)
        if (Cx::NT == 32) have_point = __any_sync(0xffffffffu, ((contact >> (cx.tid / 16)) & 1) && !w.free_mask[cx.tid]);
        else
#endif
        {
        int nclip = 0;
        for (int i = 0; i < NC; ++i) nclip += (((contact >> (i / 16)) & 1) && !w.free_mask[i]) ? 1 : 0;
        have_point = nclip > 0;
        }
    }
    cx.sync();
    if (!have_point) {
        LMH_FOR(i, NC) { w.free_mask[i] = (contact >> (i / 16)) & 1; w.c[i] = 0.0; }
    }
    LMH_FOR(i, 1) w.warm_contact = contact;
    cx.sync();
    // ---- all-free shortcut (no stored working set, at least one foot in contact: 98 % of the stages of a balancing robot).
    // With every coefficient free, f ranges over all of R^12 and c is its minimum-norm preimage, so the problem in f is
    //     min 1/2 f'D f  s.t.  Sa lam = rhs - Jb'f,   D = wF I + eps (G G')^-1  (constant, block diagonal per foot),
    // whose multiplier solves the 6 x 6 system (Sa + Jb' D^-1 Jb) lam = rhs, f = D^-1 Jb lam, c = G'(G G')^-1 f.
    // One 6 x 6 factorisation instead of chol(Sa) plus the 12 x 12 wrench-space system; if some c comes out negative
    // the general iteration below starts from scratch.
    int iters = 0, status = 0;
    bool fast_done = false;
    double lfast[6] = {0, 0, 0, 0, 0, 0};
    if (!have_point && contact != 0) {     // a swing foot (walking extension) carries no force: its block of D^-1 is zero
        double* U = Wb;                // [2][6][6] : D^-1 Jb_f
        double* S6 = Ay;               // [6][6]    : lower triangle of Sa + Jb' D^-1 Jb
#if defined(__CUDA_ARCH__) && LMH_SHORTCUT_DMMA
        if (Cx::NT == 32) {
            // the two 6 x 6 x 6 products per foot on the FP64 tensor cores (m8n8k4, two k-steps each): U_f = D^-1 Jb_f with
            // Jb_f as B, then S6 = Sa + sum_f Jb_f' U_f with the SAME fragments of Jb_f as A (transposed on the fly) and
            // U_f read back from shared memory as B (it is needed there anyway for f = D^-1 Jb lam)
            const int lane = cx.tid, fr = lane >> 2, fk = lane & 3, c0 = 2 * fk;
            const bool inr = fr < 6, inc = c0 < 6;
            double jb[2][2];           // [foot][k-step]: Jb_f[k = 4 kt + fk][fr]
            double s6a = (inr && inc) ? w.Sa[6 * fr + c0] : 0.0, s6b = (inr && inc) ? w.Sa[6 * fr + c0 + 1] : 0.0;
            LMH_UNROLL
            for (int f = 0; f < 2; ++f) {
                double ua = 0.0, ub = 0.0;
                LMH_UNROLL
                for (int kt = 0; kt < 2; ++kt) {
                    const int k = 4 * kt + fk;
                    jb[f][kt] = (k < 6 && inr) ? w.JFc[72 * f + 12 * k + fr] : 0.0;
                    const double di = (k < 6 && inr) ? K.Dinv[6 * fr + k] : 0.0;
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(ua), "+d"(ub) : "d"(di), "d"(jb[f][kt]));
                }
                const bool on = (contact >> f) & 1;
                {
                    double* pu = (inr && inc) ? U + 36 * f + 6 * fr + c0 : w.red + 2 * (lane & 15);     // (dump: w.red is free here)
                    pu[0] = on ? ua : 0.0; pu[1] = on ? ub : 0.0;
                }
            }
            __syncwarp();
            LMH_UNROLL
            for (int f = 0; f < 2; ++f) {
                LMH_UNROLL
                for (int kt = 0; kt < 2; ++kt) {
                    const int k = 4 * kt + fk;
                    const double ub = (k < 6 && inr) ? U[36 * f + 6 * k + fr] : 0.0;      // B fragment: U_f[k][n = fr]
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(s6a), "+d"(s6b) : "d"(jb[f][kt]), "d"(ub));
                }
            }
            if (inr && inc) { S6[6 * fr + c0] = s6a; S6[6 * fr + c0 + 1] = s6b; }
            __syncwarp();
        } else
#endif
        {
        LMH_FOR(it, 72) {
            const int f = it / 36, k = (it % 36) / 6, sidx = it % 6;
            double sd = 0.0;
            for (int l = 0; l < 6; ++l) sd += K.Dinv[6 * k + l] * w.JFc[72 * f + 12 * l + sidx];
            U[it] = ((contact >> f) & 1) ? sd : 0.0;
        }
        cx.sync();
        LMH_FOR(it, 36) {
            const int r = it / 6, sidx = it % 6;
            if (sidx <= r) {
                double sd = w.Sa[it];
                for (int k = 0; k < 12; ++k) sd += w.JFc[72 * (k / 6) + 12 * (k % 6) + r] * U[36 * (k / 6) + 6 * (k % 6) + sidx];
                S6[it] = sd;
            }
        }
        cx.sync();
        }
        {
            double A[36], Ainv[6];
            LMH_UNROLL
            for (int e = 0; e < 36; ++e) A[e] = (e % 6 <= e / 6) ? S6[e] : 0.0;
            chol6(A, Ainv);
            LMH_UNROLL
            for (int r = 0; r < 6; ++r) {
                double sd = w.rhs[r];
                LMH_UNROLL
                for (int m = 0; m < 6; ++m) if (m < r) sd -= A[6 * r + m] * lfast[m];
                lfast[r] = sd * Ainv[r];
            }
            LMH_UNROLL
            for (int r = 5; r >= 0; --r) {
                double sd = lfast[r];
                LMH_UNROLL
                for (int m = 0; m < 6; ++m) if (m > r) sd -= A[6 * m + r] * lfast[m];
                lfast[r] = sd * Ainv[r];
            }
        }
        LMH_FOR(it, 12) {              // f = D^-1 Jb lam
            double sd = 0.0;
            for (int q = 0; q < 6; ++q) sd += U[6 * it + q] * lfast[q];
            w.f12[it] = sd;
        }
        cx.sync();
        LMH_FOR(it, 12) {              // y = (G G')^-1 f per foot
            const int f = it / 6, k = it % 6;
            double sy = 0.0;
            for (int l = 0; l < 6; ++l) sy += K.Pinv[6 * k + l] * w.f12[6 * f + l];
            vv[12 + it] = sy;
        }
        cx.sync();
        double cmin = 0.0;
#if defined(__CUDA_ARCH__)
        if (Cx::NT == 32) {
            const int i = cx.tid, foot = i / 16, gi = i % 16;
            double sc = 0.0;
            for (int k = 0; k < 6; ++k) sc += K.Gf[16 * k + gi] * vv[12 + 6 * foot + k];
            cmin = warp_min(sc);
            // feasible: optimal.  Otherwise project onto the orthant, as the first pass of the general iteration
            // would: the clipped coefficients are the working set it starts from
            w.c[i] = sc > 0.0 ? sc : 0.0;
            w.free_mask[i] = (((contact >> foot) & 1) && (cmin >= 0.0 || sc > 0.0)) ? 1 : 0;
            w.cs[i] = sc; w.mu[i] = 0.0;
        } else
#endif
        {
            LMH_FOR(i, NC) {
                const int foot = i / 16, gi = i % 16;
                double sc = 0.0;
                for (int k = 0; k < 6; ++k) sc += K.Gf[16 * k + gi] * vv[12 + 6 * foot + k];
                w.cs[i] = sc; w.mu[i] = 0.0;
            }
            cx.sync();
            for (int i = 0; i < NC; ++i) cmin = w.cs[i] < cmin ? w.cs[i] : cmin;
            LMH_FOR(i, NC) { w.c[i] = w.cs[i] > 0.0 ? w.cs[i] : 0.0; w.free_mask[i] = (((contact >> (i / 16)) & 1) && (cmin >= 0.0 || w.cs[i] > 0.0)) ? 1 : 0; }
        }
        cx.sync();
        iters = 1;
        have_point = true;
        if (cmin >= 0.0) fast_done = true;
    }
#if defined(LMH_GENERAL_UNLIKELY) && defined(__CUDA_ARCH__)
    if (__builtin_expect(!fast_done, 0)) qp_general_active_set(cx, w, K, contact, have_point, iters, status);
#else
    if (!fast_done) qp_general_active_set(cx, w, K, contact, have_point, iters, status);
#endif
    w.qp_fallback = iters > 1 ? 1 : 0;     // stages whose all-free solution was infeasible (contact-limited)
    w.qp_iters = iters;
    w.qp_status = status;
    // lam = Sa^-1 (rhs - Jb' f12), f12 = G c.  After the shortcut both are already there: f12 = D^-1 Jb lam is the wrench
    // whose minimum-norm preimage c is (G c = f12 up to rounding), and lam came out of the 6 x 6 system; the general
    // iteration leaves them in w.f12, w.lam.
    {
        double l[6];
        LMH_UNROLL
        for (int r = 0; r < 6; ++r) l[r] = fast_done ? lfast[r] : w.lam[r];

        // a = -Ha^-1 (ga + M6' lam) through the same Woodbury pieces
        double coef[7] = {l[0], l[1], l[2], l[3], l[4], l[5], 1.0};
        double z[3] = {0, 0, 0};
        LMH_UNROLL                     // static indices keep coef in registers (a runtime index would put it in local memory)
        for (int r = 0; r < 7; ++r) { z[0] += Gm[NYR * 7 + r] * coef[r]; z[1] += Gm[NYR * 8 + r] * coef[r]; z[2] += Gm[NYR * 9 + r] * coef[r]; }
        double cu[3] = {-(Ci[0] * z[0] + Ci[1] * z[1] + Ci[2] * z[2]), -(Ci[3] * z[0] + Ci[4] * z[1] + Ci[5] * z[2]),
                        -(Ci[6] * z[0] + Ci[7] * z[1] + Ci[8] * z[2])};
        LMH_FOR(i, NQ) {
            double sy = 0.0;
            LMH_UNROLL
            for (int r = 0; r < 7; ++r) sy += Yt[NQ * r + i] * coef[r];
            LMH_UNROLL
            for (int u = 0; u < 3; ++u) sy += Yt[NQ * (7 + u) + i] * cu[u];
            w.a[i] = i < ND1 ? -sy : -sy * K.isq_wJoints;
        }
        LMH_FOR(i, 1) {
            LMH_UNROLL
            for (int r = 0; r < 6; ++r) w.lam[r] = l[r];
        }
    }
    cx.sync();
#if defined(__CUDA_ARCH__)
    if (Cx::NT == 32) { LMH_STAMP(9); qp_backsolve_warp(w, cx.tid); return; }
#endif
    for (int k = ND1 - 1; k >= 0; --k) {     // L1' a[0:18] = rhs, column oriented
        double xk = w.a[k] / w.ladiag[k];
        cx.sync();
        LMH_FOR(i, k + 1) {
            if (i == k) w.a[k] = xk;
            else w.a[i] -= L1[LD1 * k + i] * xk;
        }
        cx.sync();
    }
}

// ================================================================================================
// Phase 8: inverse dynamics to torques, base acceleration to the world frame, state derivative
// (controller.cpp:137-147, apps/offline/main.cpp:107-121).
template <class Cx>
LMH_HDN void output_phase(const Cx& cx, EnvWS& w) {
    // Branch-free on purpose: the work items of a pass differ (leg / arm / head joints; linear, angular and joint
    // coordinates) and a divergent branch costs a warp more than the few FMAs it guards.  Terms that do not apply are
    // multiplied by a zero operand (x + 0 y = x exactly) and indices are clamped to valid ones.
    LMH_FOR(j, NJ) {
        double s = w.C[6 + j];
        LMH_UNROLL
        for (int k = 0; k < 6; ++k) s += w.M6[NQ * k + 6 + j] * w.a[k];
        const int c = joint_chain(j), j0 = chain_joint0(c), n = chain_nact(c);
        LMH_UNROLL
        for (int b = 0; b < 6; ++b) {
            const bool on = b < n;
            const double hv = on ? w.Hc[36 * c + 6 * (j - j0) + b] : 0.0;
            s += hv * w.a[6 + j0 + (on ? b : 0)];
        }
        const bool leg = j < 12;
        const int f = leg ? j / 6 : 0, col = leg ? 6 + (j - 6 * f) : 6;
        LMH_UNROLL
        for (int k = 0; k < 6; ++k) {
            const double jv = leg ? w.JFc[72 * f + 12 * k + col] : 0.0;
            s -= jv * w.f12[6 * f + k];
        }
        w.tau[j] = s;
    }
    LMH_FOR(i, NQ) {
        const int k3 = i < 3 ? i : i < 6 ? i - 3 : 0;
        // X0^-1 a : linear part R0 a_lin + p0 x (R0 a_ang), angular part R0 a_ang
        const double ra[3] = {rot_row(w.R0, 0, w.a), rot_row(w.R0, 1, w.a), rot_row(w.R0, 2, w.a)};
        const double xlin = rot_row(w.R0, k3, w.a + 3) + cross_k(w.p0, ra, k3);
        const double xang = k3 == 0 ? ra[0] : k3 == 1 ? ra[1] : ra[2];
        w.xdot[NQ + i] = i < 3 ? xlin : i < 6 ? xang : w.a[i];
        // state derivative of the configuration: v + w x p, Euler rates (generalizedFunctions.cpp:43-50), joint rates
        const double qlin = w.v[k3] + cross_k(w.v + 3, w.q, k3);
        const double s2 = w.esc[2], c2 = w.esc[3], ic1 = 1.0 / w.esc[1], t1 = w.esc[0] * ic1;   // tan = sin/cos
        const double* om = w.v + 3;
        const double qe = k3 == 0 ? (c2 * ic1) * om[0] + (s2 * ic1) * om[1]
                        : k3 == 1 ? -s2 * om[0] + c2 * om[1]
                                  : c2 * t1 * om[0] + s2 * t1 * om[1] + om[2];
        w.xdot[i] = i < 3 ? qlin : i < 6 ? qe : w.v[i];
    }
    cx.sync();
}

// Debug/parity dump of one stage (global memory on the device).
template <class Cx>
LMH_HDN void dump_assembly(const Cx& cx, const EnvWS& w, const Consts& K, LmhStageDump* d) {
    LMH_FOR(it, NQ * NQ) d->Ha[it] = ha_entry(w, K, it / NQ, it % NQ);
    LMH_FOR(i, NQ) d->ga[i] = w.ga[i];
    cx.sync();
}
template <class Cx>
LMH_HDN void dump_stage(const Cx& cx, const EnvWS& w, LmhStageDump* d) {
    LMH_FOR(it, NQ * NQ) {
        int i = it / NQ, j = it % NQ;
        d->M[it] = i < 6 ? w.M6[NQ * i + j] : j < 6 ? w.M6[NQ * j + i] : hjoint(w, i - 6, j - 6);
    }
    LMH_FOR(i, NQ) { d->C[i] = w.C[i]; d->Cg[i] = w.Cg[i]; d->qppRef[i] = w.qppRef[i]; d->a[i] = w.a[i]; d->qpp[i] = w.xdot[NQ + i]; }
    LMH_FOR(i, 6 * NQ) d->AG[i] = w.AG[i];
    LMH_FOR(i, 12 * NQ) d->JFeet[i] = jf_full(w, i / NQ, i % NQ);
    LMH_FOR(i, 2 * NQ) d->xdot[i] = w.xdot[i];
    LMH_FOR(i, NC) d->c[i] = w.c[i];
    LMH_FOR(i, NJ) d->tau[i] = w.tau[i];
    LMH_FOR(i, 12) {
        d->Jpqp[i] = w.Jpqp[i]; d->footAccRef[i] = w.footAccRef[i]; d->f[i] = w.f12[i];
        d->T7[i] = i < 9 ? w.Rs[0][i] : w.ps[0][i - 9];
        d->T14[i] = i < 9 ? w.Rs[1][i] : w.ps[1][i - 9];
    }
    LMH_FOR(i, 6) { d->AGpqp[i] = w.AGpqp[i]; d->hGpRef[i] = w.hGpRef[i]; d->lam[i] = w.lam[i]; }
    LMH_FOR(i, 3) {
        d->CoM[i] = w.com[i]; d->comVel[i] = w.comVel[i]; d->comAngMom[i] = w.angMom[i];
        d->xRef[i] = w.xRef[i]; d->yRef[i] = w.yRef[i];
    }
    LMH_FOR(i, 1) { d->k_index = w.k_index; d->qp_status = w.qp_status; d->qp_iters = w.qp_iters; d->phase = w.phase; }
    cx.sync();
}

}  // namespace lmh
#include "lmh_warp.cuh"
namespace lmh {

// One dynamics(x,t) evaluation.  On entry w.q, w.v, w.vh, w.t are set; on exit w.xdot, w.tau, w.f12,
// w.qpp (and every intermediate) are valid and w.vh has been advanced to this stage's velocity.
// model part of a stage: kinematics, dynamics, Jacobians (everything up to the controller)
template <class Cx>
LMH_HDN void model_phases(const Cx& cx, EnvWS& w, const Consts& K) {
    fk_phase(cx, w, K);
    base_twist_phase(cx, w);
#if defined(__CUDA_ARCH__)
    if (Cx::NT == 32) {
        const LaneRole L = lane_role(cx.tid);
        const double* Isp = Cx::TAB ? cx.isp : K.Isp;
        if (Cx::TAB) __builtin_assume(__isShared(Isp));
        LMH_STAMP(1);
        cx.handover(2);                             // transforms and base twists are ready: the helper warps start
        LMH_STAMP(2);
        rnea_warp(w, K, Isp, L, cx.tid);
        LMH_STAMP(3);
        if (!cx.helpers) crba_warp(w, K, Isp, L, cx.tid);
        cx.handover(3);                             // mass matrix and sole Jacobians are there
        LMH_STAMP(4);
    } else
#endif
    {
        rnea_phase(cx, w, K);
        crba_phase(cx, w, K);
    }
    centroidal_jacobian_phase(cx, w, K);
}

// QP data + solve of a stage (controller.cpp:94-132, 388-479)
template <class Cx>
LMH_HDN void qp_stage(const Cx& cx, EnvWS& w, const Consts& K, LmhStageDump* dump) {
    bool factor_done = false;
#if defined(__CUDA_ARCH__)
    if (Cx::NT == 32 && K.wCoMK == 0.0) {
        assemble_ga_phase(cx, w, K);
        LMH_STAMP(7);
        if (dump) dump_assembly(cx, w, K, dump);
        qp_factor_warp(w, K, cx.tid);
        LMH_STAMP(8);
        factor_done = true;
    } else
#endif
    {
        assemble_phase(cx, w, K);
        if (dump) dump_assembly(cx, w, K, dump);
    }
    qp_phase(cx, w, K, factor_done);
}

// Not inlined on the device: the throughput kernel and the latency kernel then execute ONE compiled body of the stage
// (the same instructions, the same FMA contractions), which is what makes their results identical bit for bit.  (Inlined
// separately, the compiler contracted a few sums of products of the general active-set iteration differently in the two
// kernels: 1-ulp differences a few steps after a push.)
#if defined(__CUDACC__) && LMH_STAGE_NOINLINE
#define LMH_STAGE_ATTR __noinline__
#else
#define LMH_STAGE_ATTR
#endif
template <class Cx>
LMH_STAGE_ATTR LMH_HDN void stage_eval(const Cx& cx, EnvWS& w, const Consts& K, const Plan& P, int plan_id, LmhStageDump* dump = nullptr) {
    LMH_STAMP(0);
    model_phases(cx, w, K);
    LMH_STAMP(5);
    mpc_pd_phase(cx, w, K, P, plan_id);
    LMH_STAMP(6);
    qp_stage(cx, w, K, dump);
    LMH_STAMP(10);
    output_phase(cx, w);
    LMH_STAMP(11);
    if (dump) dump_stage(cx, w, dump);
    LMH_FOR(i, NQ) w.vh[i] = w.v[i];     // Robot::updateVelocityState (Robot.cpp:271-274)
    cx.sync();
}

// EXTENSION (the reference has no notion of falling): the health rule of a closed-loop step (the checker restates it
// independently).  A step is COMMITTED -- state, clock, statistics -- only if, with x+ the state the RK4
// combination would produce,
//     max|tau(k1)| < LMH_TAU_LIMIT,   |CoM_xy(k4) - zmp_ref(k_index(k4))| < LMH_COM_ERR_LIMIT,
//     every |x+_i| < LMH_STATE_LIMIT (NaN fails),   x+[2] (torso height) > LMH_TORSO_MIN.
// Otherwise the environment is FALLEN: state and clock keep the values of the last healthy step, the statistics are
// closed (they never see the failing step), alive = 0, and no further stage is evaluated for it.

// One RK4 step (rk4.hpp:5-18) + Clock::step (Clock.hpp:11).  x: [60] state in shared memory, kacc:
// [60] scratch, t: the env clock (same value in every thread of the group).
template <class Cx>
LMH_HDN void rk4_step(const Cx& cx, EnvWS& w, double* x, double* kacc, double& t, const Consts& K, const Plan& P,
                      int plan_id, LmhStepLog* log, StatsAcc* st) {
    const double dt = K.dt;
    const double half = 0.5 * dt;
    LMH_STAMP(12);
    cx.handover(1);                                 // helper warps learn here whether the environment is still alive
    if (w.dead != 0) {
        // a fallen environment evaluates nothing; it only keeps its appointments with the warps it shares a block with,
        // and its log record repeats the last evaluated stage
#if defined(__CUDA_ARCH__)
        for (int stage = 0; stage < 4; ++stage) cx.align();
#endif
        if (log) {
            LMH_FOR(i, 3) { log->com_k4[i] = w.com[i]; log->rfoot[i] = w.ps[0][i]; log->lfoot[i] = w.ps[1][i]; }
            LMH_FOR(i, NJ) log->tau_k1[i] = 0.0;
            LMH_FOR(i, 4) { log->k_index[i] = w.k_index; log->phase[i] = w.phase; log->contact[i] = w.contact; }
            LMH_FOR(i, 1) { log->qp_iters = 0; log->qp_fail = 0; }
        }
        return;
    }
    LMH_FOR(i, 1) { w.acc_iters = 0; w.acc_fails = 0; w.acc_nfall = 0; w.acc_tau = 0.0; w.acc_qpp = 0.0; }
    for (int stage = 0; stage < 4; ++stage) {
        const double cstage = stage == 0 ? 0.0 : stage == 3 ? dt : half;
#if defined(__CUDA_ARCH__)
        // several one-warp environments per block: re-align their instruction streams once per stage so that the
        // instruction cache (74 % hit rate: the stage code is far larger than the cache) serves them together
        cx.align();
#endif
        LMH_FOR(i, 2 * NQ) {
            double xi = stage == 0 ? x[i] : mul_then_add(x[i], cstage, w.xdot[i]);
            if (i < NQ) w.q[i] = xi; else w.v[i - NQ] = xi;
        }
        LMH_FOR(i, 1) w.t = stage == 0 ? t : mul_then_add(t, stage == 3 ? 1.0 : 0.5, dt);
        cx.sync();
        stage_eval(cx, w, K, P, plan_id);
        LMH_FOR(i, 2 * NQ) {
            double k = w.xdot[i];
            kacc[i] = stage == 0 ? k : stage == 3 ? kacc[i] + k : mul_then_add(kacc[i], 2.0, k);
        }
        LMH_FOR(i, 1) { w.acc_iters += w.qp_iters; w.acc_fails += w.qp_status != 0; w.acc_nfall += w.qp_fallback; }
        if (log) {
            LMH_FOR(i, 1) { log->k_index[stage] = w.k_index; log->phase[stage] = w.phase; log->contact[stage] = w.contact; }
            if (stage == 0) { LMH_FOR(i, NJ) log->tau_k1[i] = w.tau[i]; }
        }
        if (stage == 0) {
#if defined(__CUDA_ARCH__)
            if (Cx::NT == 32) {
                const int i = cx.tid;
                // NaN-propagating maxima (fmax would drop a NaN torque): a non-finite torque must fail the health test
                const double tv = fabs(w.tau[i < NJ ? i : 0]), av = fabs(w.xdot[NQ + (i < NQ ? i : 0)]);
                const double ti = i < NJ ? tv : 0.0, ai = i < NQ ? av : 0.0;
                const double tm = warp_max(ti == ti ? ti : 1e300), am = warp_max(ai == ai ? ai : 1e300);
                if (i == 0) { w.acc_tau = tm; w.acc_qpp = am; }
            } else
#endif
            {
                LMH_FOR(i, 1) {
                    double tm = 0.0, am = 0.0;
                    for (int j = 0; j < NJ; ++j) { const double v = fabs(w.tau[j]); tm = fmax(tm, v == v ? v : 1e300); }
                    for (int j = 0; j < NQ; ++j) { const double v = fabs(w.xdot[NQ + j]); am = fmax(am, v == v ? v : 1e300); }
                    w.acc_tau = tm; w.acc_qpp = am;
                }
            }
        }
        cx.sync();
    }
    const double sixth = dt / 6.0;
    LMH_FOR(i, 2 * NQ) kacc[i] = mul_then_add(x[i], sixth, kacc[i]);       // the candidate x+
    cx.sync();
    // health of the step: every thread reaches the same verdict
    int k0 = w.k_index;
    if (k0 < 0) k0 = 0;
    if (k0 >= P.zmp_len) k0 = P.zmp_len - 1;
    const double ex = w.com[0] - P.zmp_x[(size_t)plan_id * P.zmp_len + k0];
    const double ey = w.com[1] - P.zmp_y[(size_t)plan_id * P.zmp_len + k0];
    const double err = sqrt(ex * ex + ey * ey);
    const double tau_inf = w.acc_tau;
    bool healthy = tau_inf < LMH_TAU_LIMIT && err < LMH_COM_ERR_LIMIT && kacc[2] > LMH_TORSO_MIN;
#if defined(__CUDA_ARCH__)
    if (Cx::NT == 32) {
        const int i = cx.tid;
        const bool bad = !(fabs(kacc[i]) < LMH_STATE_LIMIT) || (i < 28 && !(fabs(kacc[32 + i]) < LMH_STATE_LIMIT));
        healthy = healthy && !__any_sync(0xffffffffu, bad);
    } else
#endif
    {
        for (int j = 0; j < 2 * NQ; ++j) healthy = healthy && (fabs(kacc[j]) < LMH_STATE_LIMIT);
    }
    if (healthy) {
        LMH_FOR(i, 2 * NQ) x[i] = kacc[i];
        LMH_FOR(i, 1) {
            t = t + dt;
            if (st) {
                st->max_tau = fmax(st->max_tau, tau_inf); st->max_qpp = fmax(st->max_qpp, w.acc_qpp); st->max_err = fmax(st->max_err, err);
                st->qp_iters += w.acc_iters; st->qp_fail += w.acc_fails; st->steps += 1; st->qp_fallback += w.acc_nfall;
            }
        }
    } else {
        LMH_FOR(i, 1) { if (st) st->alive = 0; }
    }
    if (log) {
        LMH_FOR(i, 3) { log->com_k4[i] = w.com[i]; log->rfoot[i] = w.ps[0][i]; log->lfoot[i] = w.ps[1][i]; }
        LMH_FOR(i, 1) { log->qp_iters = w.acc_iters; log->qp_fail = w.acc_fails; }
    }
#if defined(LMH_TIMELINE) && defined(__CUDA_ARCH__)
    LMH_STAMP(13);
    cx.sync();
    if (log) { LMH_FOR(i, 16) log->tau_k1[i] = w.stamp[i]; }
#endif
    cx.sync();                        // nobody reads kacc / w.dead of this step any more
    if (!healthy) { LMH_FOR(i, 1) w.dead = 1; }
    cx.sync();
}

}  // namespace lmh
