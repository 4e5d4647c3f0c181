// lmh_setup.hpp -- host-side, run-once setup of a batch: the constants every environment shares
// (robot tables, gains, the MPC gain vector), the nominal plan (ZMP reference, foot polynomials) and
// the initial posture by inverse kinematics.  Mirrors what apps/offline/main.cpp:12-64 does before it
// enters the loop.  Setup only: nothing here runs per step.
//
//   Robot::Robot                      src/Robot.cpp:5-43          -> build_robot_tables
//   Kinematics::compute (Newton IK)   src/invKinematics.cpp:11-70,151-267 -> solve_initial_posture
//   Mpc3dLip::initialize              src/mpcLinearPendulum.cpp:41-76 -> build_mpc_gain
//   footCoeffTrajectory/findPolyCoeff src/footRefTrajectory.cpp:4-47, generalizedFunctions.cpp:103-163
//   ZMP::stanceZMP                    src/zmpGeneration.cpp:39-60
#pragma once
#include "lmh_model.cuh"
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <vector>

namespace lmh {

struct Scenario {                       // literals of apps/offline/main.cpp:13-56
    double dt = 0.01, sim_time = 5.0, time_horizon = 0.5;
    double com[3] = {-0.02, 0.0, 0.26};
    double rf[6] = {0, -0.05, 0, 0, 0, 0};
    double lf[6] = {0, 0.05, 0, 0, 0, 0};
    double mu = 0.7;                    // controller.hpp:81
    double mpc_alpha = 1e-3, mpc_beta = 1.0;   // mpcLinearPendulum.hpp:47-48
    int qp_max_iter = 200;
    double mpc_zcom = 0.0;              // 0 = CoM height after the inverse kinematics (main.cpp:39)
};

struct NaoLinkRow { double m; double c[3]; double I[9]; };
static const NaoLinkRow kNaoLinks[28] = {
#include "lmh_nao_table.inc"
};

// dense solve A x = b (n x n, row-major), Gaussian elimination with partial pivoting.  Stands in for
// the reference's colPivHouseholderQr().solve on well-posed square systems.
inline bool dense_solve(int n, std::vector<double> A, std::vector<double> b, double* x) {
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i) if (std::fabs(A[i * n + k]) > std::fabs(A[p * n + k])) p = i;
        if (A[p * n + k] == 0.0) return false;
        if (p != k) { for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[p * n + j]); std::swap(b[k], b[p]); }
        for (int i = k + 1; i < n; ++i) {
            double f = A[i * n + k] / A[k * n + k];
            if (f == 0.0) continue;
            for (int j = k; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
            b[i] -= f * b[k];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * x[j];
        x[i] = s / A[i * n + i];
    }
    return true;
}

// boundary-value polynomial, ascending powers (generalizedFunctions.cpp:103-163)
inline std::vector<double> poly_fit(const std::vector<double>& pt, const std::vector<double>& pv,
                                    const std::vector<double>& vt, const std::vector<double>& vv,
                                    const std::vector<double>& at, const std::vector<double>& av) {
    int n = (int)(pt.size() + vt.size() + at.size());
    std::vector<double> A((size_t)n * n, 0.0), b(n, 0.0), c(n, 0.0);
    int row = 0;
    for (size_t i = 0; i < pt.size(); ++i, ++row) { double tp = 1; for (int j = 0; j < n; ++j) { A[row * n + j] = tp; tp *= pt[i]; } b[row] = pv[i]; }
    for (size_t i = 0; i < vt.size(); ++i, ++row) { double tp = 1; for (int j = 1; j < n; ++j) { A[row * n + j] = j * tp; tp *= vt[i]; } b[row] = vv[i]; }
    for (size_t i = 0; i < at.size(); ++i, ++row) { double tp = 1; for (int j = 2; j < n; ++j) { A[row * n + j] = j * (j - 1) * tp; tp *= at[i]; } b[row] = av[i]; }
    if (!dense_solve(n, A, b, c.data())) throw std::runtime_error("lmh: singular polynomial fit");
    return c;
}

// 20 coefficients of one foot: x(6), y(6), z(8) (footRefTrajectory.cpp:4-47)
inline void foot_coefficients(const double cur[3], const double des[3], double step_height, double T, double* out20) {
    std::vector<double> t2 = {0, T}, z2 = {0, 0}, t3 = {0, T / 2, T}, z3 = {0, 0, 0};
    std::vector<double> x = poly_fit(t2, {cur[0], des[0]}, t2, z2, t2, z2);
    std::vector<double> y = poly_fit(t2, {cur[1], des[1]}, t2, z2, t2, z2);
    std::vector<double> z = poly_fit(t3, {cur[2], step_height, des[2]}, t3, z3, t2, z2);
    for (int i = 0; i < 6; ++i) { out20[i] = x[i]; out20[6 + i] = y[i]; }
    for (int i = 0; i < 8; ++i) out20[12 + i] = z[i];
}

// modified-DH rows of the 25 local transforms (Robot.cpp:176-196): 24 joints + the extra head frame
inline void nao_dh_row(int i, double* r, double* d, double* alpha) {
    const double h = 3.14159265358979323846 / 2;      // Robot.cpp:3
    static const double R[25] = {-0.07071, 0, 0, 0, 0, 0, 0.07071, 0, 0, 0, 0, 0, 0, 0, 0.105, 0, 0.05595, 0, 0, 0.105, 0, 0.05595, 0, 0, 0};
    static const double D[25] = {0, 0, 0, -0.1, -0.1029, 0, 0, 0, 0, -0.1, -0.1029, 0, 0, 0, -0.015, 0, 0, 0, 0, -0.015, 0, 0, 0, 0, 0.030};
    const double AL[25] = {0, h, h, 0, 0, -h, -h, -h, h, 0, 0, -h, -h, h, h, -h, h, h, h, h, -h, h, 0, -h, 0};
    *r = R[i]; *d = D[i]; *alpha = AL[i];
}
// one local transform (Khalil's modified DH, Robot.cpp:200-220), 4 x 4 row-major
inline void nao_local_transform(int i, double theta, double* T16) {
    double r, d, al;
    nao_dh_row(i, &r, &d, &al);
    const double ct = std::cos(theta), st = std::sin(theta), ca = std::cos(al), sa = std::sin(al);
    const double T[16] = {ct, -st, 0.0, d, ca * st, ca * ct, -sa, -r * sa, sa * st, sa * ct, ca, r * ca, 0.0, 0.0, 0.0, 1.0};
    std::memcpy(T16, T, sizeof(T));
}

// Working-set table of the WBC QP (qp_general_active_set, lmh_model.cuh).  A working set fixes which of a foot's 16 wrench
// generators may carry force; everything that depends on that subset F ALONE is a constant of the robot and is tabulated
// here for all 2^16 subsets (bit g of the index = generator g free), 72 doubles each:
//   Kf [6][6] = B (wF B'B + eps I)^-1 B'          the wrench f_F = Kf g that answers the wrench-space target g = Jb lam,
//   Yf [6][6]                                     y = Yf g with c_F = G_F' y, the coefficients of that wrench,
// where B is the pivoted (rank-revealing) Cholesky factor of G_F G_F'; columns beyond the numerical rank are dropped, so
// a foot on an edge, on one vertex or unloaded is as well conditioned as a flat one.  y is the solution of B'y = zeta with
// zeros on the rows that were never pivoted (any solution gives the same c_F: they differ by null(G_F')).
// For F = everything, Kf = Consts::Dinv and Yf = Pinv Dinv (the all-free shortcut).
inline void qp_working_set_entry(const double* Gf, double wF, double eps, unsigned mask, double* out72) {
    // extended precision throughout (run once per process): the rows come out correctly rounded to a few ulp, so the
    // table adds nothing to the error budget of a pass
    typedef long double real;
    real S[36], B[36], Z[36], D[36], Kf[36];
    int pord[6];
    bool pdone[6] = {false, false, false, false, false, false};
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            real s = 0;
            for (int g = 0; g < 16; ++g) if ((mask >> g) & 1u) s += (real)Gf[16 * i + g] * (real)Gf[16 * j + g];
            S[6 * i + j] = s; B[6 * i + j] = 0;
        }
    real dmax0 = 0;
    for (int i = 0; i < 6; ++i) dmax0 = S[7 * i] > dmax0 ? S[7 * i] : dmax0;
    int rank = 0;
    for (int k = 0; k < 6; ++k) {
        int p = -1;
        real best = 1e-10L * dmax0;
        for (int j = 0; j < 6; ++j) if (!pdone[j] && S[7 * j] > best) { best = S[7 * j]; p = j; }
        if (p < 0) break;
        const real d = std::sqrt(best);
        for (int i = 0; i < 6; ++i) B[6 * i + k] = pdone[i] ? (real)0 : (i == p ? d : S[6 * i + p] / d);
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) S[6 * i + j] -= B[6 * i + k] * B[6 * j + k];
        pdone[p] = true; pord[k] = p; rank = k + 1;
    }
    for (int e = 0; e < 72; ++e) out72[e] = 0.0;
    if (rank == 0) return;
    // D = wF B'B + eps I (rank x rank, SPD) = L L'; Z = D^-1 B'   (rank x 6)
    for (int k = 0; k < rank; ++k)
        for (int l = 0; l < rank; ++l) {
            real t = 0;
            for (int i = 0; i < 6; ++i) t += B[6 * i + k] * B[6 * i + l];
            D[6 * k + l] = (real)wF * t + (k == l ? (real)eps : (real)0);
        }
    for (int k = 0; k < rank; ++k) {
        for (int i = k; i < rank; ++i) {
            real t = D[6 * i + k];
            for (int m = 0; m < k; ++m) t -= D[6 * i + m] * D[6 * k + m];
            if (i == k) {
                if (!(t > 0)) throw std::runtime_error("lmh: singular working-set weight");
                D[6 * k + k] = std::sqrt(t);
            } else D[6 * i + k] = t / D[6 * k + k];
        }
    }
    for (int j = 0; j < 6; ++j) {
        real x[6];
        for (int k = 0; k < rank; ++k) {
            real t = B[6 * j + k];
            for (int m = 0; m < k; ++m) t -= D[6 * k + m] * x[m];
            x[k] = t / D[6 * k + k];
        }
        for (int k = rank - 1; k >= 0; --k) {
            real t = x[k];
            for (int m = k + 1; m < rank; ++m) t -= D[6 * m + k] * x[m];
            x[k] = t / D[6 * k + k];
        }
        for (int k = 0; k < rank; ++k) Z[6 * k + j] = x[k];
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            real s = 0;
            for (int k = 0; k < rank; ++k) s += B[6 * i + k] * Z[6 * k + j];
            Kf[6 * i + j] = s;
        }
    for (int i = 0; i < 6; ++i)        // Kf is symmetric: remove the rounding asymmetry
        for (int j = 0; j <= i; ++j) { const double v = (double)(0.5L * (Kf[6 * i + j] + Kf[6 * j + i])); out72[6 * i + j] = v; out72[6 * j + i] = v; }
    for (int j = 0; j < 6; ++j) {      // column j of Yf: back substitution in pivot order
        real y[6] = {0, 0, 0, 0, 0, 0};
        for (int k = rank - 1; k >= 0; --k) {
            const int p = pord[k];
            real t = Z[6 * k + j];
            for (int i = 0; i < 6; ++i) t -= B[6 * i + k] * y[i];      // y[p] is still zero here
            y[p] = t / B[6 * p + k];
        }
        for (int i = 0; i < 6; ++i) out72[36 + 6 * i + j] = (double)y[i];
    }
}

constexpr int kQpTabRow = 72;                      // doubles per working set
constexpr size_t kQpTabDoubles = (size_t)65536 * kQpTabRow;

// the whole table, built once per process and set of generators (friction coefficient)
inline std::shared_ptr<const std::vector<double>> qp_working_set_table(const Consts& K) {
    struct Entry { std::vector<double> key; std::shared_ptr<const std::vector<double>> tab; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    std::vector<double> key(K.Gf, K.Gf + 96);
    key.push_back(K.wForce); key.push_back(K.epsC);
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& en : cache) if (en.key == key) return en.tab;
    auto tab = std::make_shared<std::vector<double>>(kQpTabDoubles);
    for (unsigned m = 0; m < 65536u; ++m) qp_working_set_entry(K.Gf, K.wForce, K.epsC, m, tab->data() + (size_t)m * kQpTabRow);
    cache.push_back(Entry{key, tab});
    return tab;
}

class HostModel {
public:
    Consts K;
    std::shared_ptr<const std::vector<double>> qp_tab;     // working-set table of the WBC QP; K.qp_tab points at the HOST copy
    std::vector<double> zmp_x, zmp_y, foot_coef;   // nominal plan 0
    double q0[NQ];
    int ik_iterations = 0;
    Scenario sc;
    // dense form of the LIPM-MPC QP for the device-side dense solver (lmh_mpc.cuh), padded to np = 8 ceil(n / 8):
    // beta Pu' (upper triangular), the Cholesky factor L of H = alpha I + beta Pu'Pu and its transpose, the inverses of
    // its diagonal 8 x 8 blocks (and their transposes), 1 / diag(L), Px
    std::vector<double> mpc_PuT, mpc_L, mpc_LT, mpc_Dinv, mpc_DinvT, mpc_rdiag, mpc_Px;
    std::vector<double> mpc_Px_raw, mpc_Pu_raw;
    double link_inertia[28 * 9];                     // inertia of every link about its CoM in the joint frame (Robot.cpp:14-22)      // Px [n][2], Pu [n][n] as Mpc3dLip::initialize builds them
    int mpc_np = 0;

    explicit HostModel(const Scenario& s = Scenario()) : sc(s), ws_(new EnvWS) {
        std::memset(&K, 0, sizeof(K));
        std::memset(ws_.get(), 0, sizeof(EnvWS));
        build_robot_tables();
        build_controller_tables();
        qp_tab = qp_working_set_table(K);
        K.qp_tab = qp_tab->data();
        // stand task: constant ZMP in the middle of the support zone (zmpGeneration.cpp:39-60)
        int samples = static_cast<int>((sc.sim_time + 0.5) / sc.dt);
        zmp_x.assign(samples, 0.0);
        zmp_y.assign(samples, 0.0);
        foot_coef.assign(40, 0.0);
        foot_coefficients(sc.rf, sc.rf, 0.0, sc.sim_time, foot_coef.data());
        foot_coefficients(sc.lf, sc.lf, 0.0, sc.sim_time, foot_coef.data() + 20);
        solve_initial_posture();
        build_mpc_gain(sc.mpc_zcom != 0.0 ? sc.mpc_zcom : ws_->com[2]);     // zCom = CoM height after IK (main.cpp:39)
    }

    // model phases of one configuration on the host (used by the IK and by tools)
    EnvWS& evaluate_model(const double* q, const double* v) {
        EnvWS& w = *ws_;
        for (int i = 0; i < NQ; ++i) { w.q[i] = q[i]; w.v[i] = v ? v[i] : 0.0; w.vh[i] = w.v[i]; }
        Ctx<1> cx{0, 0};
        fk_phase(cx, w, K);
        base_twist_phase(cx, w);
        rnea_phase(cx, w, K);
        crba_phase(cx, w, K);
        centroidal_jacobian_phase(cx, w, K);
        return w;
    }
    EnvWS& workspace() { return *ws_; }

private:
    std::unique_ptr<EnvWS> ws_;

    void build_robot_tables() {
        const double kPi = 3.14159265358979323846;   // Robot.cpp:3
        const double h = kPi / 2;
        // modified-DH rows of the 24 joint transforms (Robot.cpp:180-196)
        (void)h;
        for (int i = 0; i < 24; ++i) {
            double al;
            nao_dh_row(i, &K.dh_r[i], &K.dh_d[i], &al);
            K.dh_ca[i] = std::cos(al); K.dh_sa[i] = std::sin(al);   // cos(pi/2) residue kept, as upstream
            K.th_off[i] = 0.0;
        }
        K.th_off[1] = (3.0 / 4) * kPi; K.th_off[6] = -(1.0 / 2) * kPi; K.th_off[7] = (1.0 / 4) * kPi;   // Robot.cpp:59-87
        K.th_off[13] = (1.0 / 2) * kPi; K.th_off[18] = (1.0 / 2) * kPi; K.th_off[23] = -(1.0 / 2) * kPi;
        const double a = 0.7071;   // sic (Robot.cpp:92-103): the hip pre-rotations are not exactly orthonormal
        const double pre[NCH][12] = {
            {0, -1, 0, a, 0, a, -a, 0, a, 0, 0, 0},
            {1, 0, 0, 0, a, a, 0, -a, a, 0, 0, 0},
            {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, -0.098, 0.13591},
            {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0.098, 0.13591},
            {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.1615}};
        for (int c = 0; c < NCH; ++c) for (int k = 0; k < 12; ++k) K.pre[12 * c + k] = pre[c][k];
        K.sole_x = -0.0452;
        K.gravity = 9.81;
        const double q0n[NQ] = {-0.0185, 0, 0.282, 0, 0, 0, 0, 0, -0.5, 0.8, -0.3, 0, 0, 0, -0.5,
                                0.8, -0.3, 0, 1.6, 0, 0, 0, 0, -1.6, 0, 0, 0, 0, 0, 0};   // Robot.cpp:242-262
        for (int i = 0; i < NQ; ++i) { K.qdes[i] = q0n[i]; q0[i] = q0n[i]; }
        const double rd[9] = {0, 0, 1, 0, -1, 0, 1, 0, 0};
        for (int k = 0; k < 9; ++k) K.Rdes[k] = rd[k];

        // inertial data are given in world-aligned frames at q = 0: re-express in the joint frames (Robot.cpp:14-22)
        EnvWS& w = *ws_;
        for (int i = 0; i < NQ; ++i) { w.q[i] = 0.0; w.v[i] = 0.0; w.vh[i] = 0.0; }
        for (int i = 0; i < 28; ++i) { K.mc[4 * i + 3] = kNaoLinks[i].m; }   // masses first: fk_phase sums the CoM
        K.mass = 1.0; K.inv_mass = 1.0;
        Ctx<1> cx{0, 0};
        fk_phase(cx, w, K);
        double mass = 0.0;
        for (int i = 0; i < 28; ++i) {
            const NaoLinkRow& L = kNaoLinks[i];
            double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            if (i < NF) for (int k = 0; k < 9; ++k) R[k] = w.u.m.p.Tw[12 * i + k];
            double cj[3], RtI[9], Ij[9];
            for (int r3 = 0; r3 < 3; ++r3) cj[r3] = R[r3] * L.c[0] + R[3 + r3] * L.c[1] + R[6 + r3] * L.c[2];
            for (int r3 = 0; r3 < 3; ++r3) for (int c = 0; c < 3; ++c) RtI[3 * r3 + c] = R[r3] * L.I[c] + R[3 + r3] * L.I[3 + c] + R[6 + r3] * L.I[6 + c];
            for (int r3 = 0; r3 < 3; ++r3) for (int c = 0; c < 3; ++c) Ij[3 * r3 + c] = RtI[3 * r3] * R[c] + RtI[3 * r3 + 1] * R[3 + c] + RtI[3 * r3 + 2] * R[6 + c];
            mass += L.m;
            for (int k = 0; k < 9; ++k) link_inertia[9 * i + k] = Ij[k];
            K.mc[4 * i + 0] = cj[0]; K.mc[4 * i + 1] = cj[1]; K.mc[4 * i + 2] = cj[2]; K.mc[4 * i + 3] = L.m;
            // spatial inertia (Dynamics.cpp:4-13) for the torso and the actuated frames only (Dynamics.cpp:15-27)
            double* S = K.Isp + 36 * i;
            for (int k = 0; k < 36; ++k) S[k] = 0.0;
            bool used = (i == 0) || (i < NF && frame_joint(i) >= 0);
            if (!used) continue;
            double cx3[9] = {0, -cj[2], cj[1], cj[2], 0, -cj[0], -cj[1], cj[0], 0};
            double mcx[9], mcc[9];
            for (int k = 0; k < 9; ++k) mcx[k] = L.m * cx3[k];
            for (int r3 = 0; r3 < 3; ++r3) for (int c = 0; c < 3; ++c) mcc[3 * r3 + c] = mcx[3 * r3] * cx3[c] + mcx[3 * r3 + 1] * cx3[3 + c] + mcx[3 * r3 + 2] * cx3[6 + c];
            for (int r3 = 0; r3 < 3; ++r3)
                for (int c = 0; c < 3; ++c) {
                    S[6 * r3 + c] = Ij[3 * r3 + c] - mcc[3 * r3 + c];
                    S[6 * r3 + 3 + c] = mcx[3 * r3 + c];
                    S[6 * (3 + r3) + c] = -mcx[3 * r3 + c];
                    S[6 * (3 + r3) + 3 + c] = r3 == c ? L.m : 0.0;
                }
        }
        K.mass = mass; K.inv_mass = 1.0 / mass;
    }

    void build_controller_tables() {
        K.KpJ = 300; K.KdJ = 34; K.KpM = 10; K.KdM = 6.32; K.KpF = 500; K.KdF = 44;        // controller.hpp:100-111
        K.wCoML = 4000; K.wCoMK = 0; K.wBasePos = 10; K.wBaseAng = 10; K.wJoints = 1;      // controller.hpp:118-124
        K.wForce = 1; K.wFoot = 100000; K.epsC = 1e-8;
        K.sq_wCoML = std::sqrt(K.wCoML); K.isq_wJoints = 1.0 / std::sqrt(K.wJoints);                                      // controller.cpp:109-123
        K.qp_max_iter = sc.qp_max_iter;
        K.dt = sc.dt;
        // friction pyramid edges and contact vertices (controller.cpp:33-36, Robot.cpp:38-42)
        const double mu = sc.mu;
        const double B[4][3] = {{mu, 0, 1}, {0, mu, 1}, {-mu, 0, 1}, {0, -mu, 1}};
        const double V[4][3] = {{0.1, 0.025, 0}, {0.1, -0.025, 0}, {-0.05, 0.025, 0}, {-0.05, -0.025, 0}};
        for (int v = 0; v < 4; ++v)
            for (int e = 0; e < 4; ++e) {
                const double* p = V[v];
                const double* b = B[e];
                double mom[3] = {p[1] * b[2] - p[2] * b[1], p[2] * b[0] - p[0] * b[2], p[0] * b[1] - p[1] * b[0]};
                for (int k = 0; k < 3; ++k) { K.Gf[16 * k + 4 * v + e] = mom[k]; K.Gf[16 * (3 + k) + 4 * v + e] = b[k]; }
            }
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                double s = 0;
                for (int k = 0; k < 16; ++k) s += K.Gf[16 * i + k] * K.Gf[16 * j + k];
                K.GGt[6 * i + j] = s;
            }
        {   // all-free shortcut of the QP: Pinv = (G G')^-1, Dinv = (wForce I + epsC Pinv)^-1 per foot
            std::vector<double> P(K.GGt, K.GGt + 36), e(6), x(6), D(36);
            for (int i = 0; i < 6; ++i) {
                for (int k = 0; k < 6; ++k) e[k] = k == i ? 1.0 : 0.0;
                if (!dense_solve(6, P, e, x.data())) throw std::runtime_error("lmh: singular wrench Gram matrix");
                for (int k = 0; k < 6; ++k) K.Pinv[6 * k + i] = x[k];
            }
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) D[6 * i + j] = K.epsC * 0.5 * (K.Pinv[6 * i + j] + K.Pinv[6 * j + i]) + (i == j ? K.wForce : 0.0);
            for (int i = 0; i < 6; ++i) {
                for (int k = 0; k < 6; ++k) e[k] = k == i ? 1.0 : 0.0;
                if (!dense_solve(6, D, e, x.data())) throw std::runtime_error("lmh: singular wrench weight");
                for (int k = 0; k < 6; ++k) K.Dinv[6 * k + i] = x[k];
            }
        }
    }

    // u0 = -e0' H^-1 beta Pu' (Px x - z)  with H = alpha I + beta Pu'Pu constant  =>  gain vector r = beta Pu H^-1 e0
    void build_mpc_gain(double zcom) {
        const double dt = sc.dt;
        int N = static_cast<int>(sc.time_horizon / dt);     // mpcLinearPendulum.hpp:49
        int n = N + 1;
        if (n > MPC_MAXN) throw std::runtime_error("lmh: MPC horizon too long");
        const double A[4] = {1, dt, 0, 1}, Bv[2] = {dt * dt / 2, dt};
        const double D = -zcom / 9.81;
        std::vector<double> Px((size_t)n * 2), Pu((size_t)n * n, 0.0);
        double Ap[4] = {1, 0, 0, 1};
        Px[0] = 1; Px[1] = 0; Pu[0] = D;
        for (int i = 1; i <= N; ++i) {
            double T[4] = {Ap[0] * A[0] + Ap[1] * A[2], Ap[0] * A[1] + Ap[1] * A[3], Ap[2] * A[0] + Ap[3] * A[2], Ap[2] * A[1] + Ap[3] * A[3]};
            std::memcpy(Ap, T, sizeof(T));
            Px[2 * i] = Ap[0]; Px[2 * i + 1] = Ap[1];
            Pu[(size_t)i * n + i] = D;
        }
        // Pu[i+j][i-1] = C A^j B
        double Aj[4] = {1, 0, 0, 1};
        for (int j = 0; j <= N - 1; ++j) {
            double cab = Aj[0] * Bv[0] + Aj[1] * Bv[1];
            for (int i = 1; i + j <= N; ++i) Pu[(size_t)(i + j) * n + (i - 1)] = cab;
            double T[4] = {Aj[0] * A[0] + Aj[1] * A[2], Aj[0] * A[1] + Aj[1] * A[3], Aj[2] * A[0] + Aj[3] * A[2], Aj[2] * A[1] + Aj[3] * A[3]};
            std::memcpy(Aj, T, sizeof(T));
        }
        std::vector<double> H((size_t)n * n), e0(n, 0.0), h(n);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double s = 0;
                for (int k = 0; k < n; ++k) s += Pu[(size_t)k * n + i] * Pu[(size_t)k * n + j];
                H[(size_t)i * n + j] = sc.mpc_beta * s + (i == j ? sc.mpc_alpha : 0.0);
            }
        e0[0] = 1.0;
        if (!dense_solve(n, H, e0, h.data())) throw std::runtime_error("lmh: singular MPC Hessian");
        K.mpc_rPx[0] = K.mpc_rPx[1] = 0.0;
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += Pu[(size_t)i * n + k] * h[k];
            K.mpc_r[i] = sc.mpc_beta * s;
            K.mpc_rPx[0] += K.mpc_r[i] * Px[2 * i];
            K.mpc_rPx[1] += K.mpc_r[i] * Px[2 * i + 1];
        }
        K.mpc_n1 = n; K.mpc_dt = dt; K.mpc_zcom = zcom;
        for (int k = 0; k < 4; ++k) K.mpc_A[k] = A[k];
        K.mpc_B[0] = Bv[0]; K.mpc_B[1] = Bv[1];
        build_mpc_dense(n, Px, Pu, H);
        mpc_Px_raw = Px; mpc_Pu_raw = Pu;
    }

    // H = L L' (the reference hands H to qpOASES at every call, mpcLinearPendulum.cpp:89-90,111-133; H is constant)
    void build_mpc_dense(int n, const std::vector<double>& Px, const std::vector<double>& Pu, const std::vector<double>& H) {
        const int np = (n + 7) / 8 * 8;
        mpc_np = np;
        auto at = [np](int i, int j) { return (size_t)i * np + j; };
        mpc_PuT.assign((size_t)np * np, 0.0); mpc_L.assign((size_t)np * np, 0.0); mpc_LT.assign((size_t)np * np, 0.0);
        mpc_Px.assign((size_t)np * 2, 0.0); mpc_rdiag.assign(np, 1.0);
        for (int i = 0; i < n; ++i) {
            mpc_Px[2 * i] = Px[2 * i]; mpc_Px[2 * i + 1] = Px[2 * i + 1];
            for (int k = 0; k < n; ++k) mpc_PuT[at(i, k)] = sc.mpc_beta * Pu[(size_t)k * n + i];
        }
        for (int i = 0; i < np; ++i) mpc_L[at(i, i)] = 1.0;
        for (int j = 0; j < n; ++j) {                      // Cholesky, column by column
            double d = H[(size_t)j * n + j];
            for (int k = 0; k < j; ++k) d -= mpc_L[at(j, k)] * mpc_L[at(j, k)];
            if (!(d > 0.0)) throw std::runtime_error("lmh: MPC Hessian is not positive definite");
            const double ljj = std::sqrt(d);
            mpc_L[at(j, j)] = ljj;
            for (int i = j + 1; i < n; ++i) {
                double s = H[(size_t)i * n + j];
                for (int k = 0; k < j; ++k) s -= mpc_L[at(i, k)] * mpc_L[at(j, k)];
                mpc_L[at(i, j)] = s / ljj;
            }
        }
        for (int i = 0; i < np; ++i) {
            mpc_rdiag[i] = 1.0 / mpc_L[at(i, i)];
            for (int j = 0; j < np; ++j) mpc_LT[at(j, i)] = mpc_L[at(i, j)];
        }
        const int nt = np / 8;
        mpc_Dinv.assign((size_t)nt * 64, 0.0); mpc_DinvT.assign((size_t)nt * 64, 0.0);
        for (int t = 0; t < nt; ++t)                       // inverse of a lower-triangular 8 x 8 block by forward substitution
            for (int c = 0; c < 8; ++c)
                for (int r = c; r < 8; ++r) {
                    double s = r == c ? 1.0 : 0.0;
                    for (int k = c; k < r; ++k) s -= mpc_L[at(8 * t + r, 8 * t + k)] * mpc_Dinv[(size_t)64 * t + 8 * k + c];
                    const double v = s / mpc_L[at(8 * t + r, 8 * t + r)];
                    mpc_Dinv[(size_t)64 * t + 8 * r + c] = v;
                    mpc_DinvT[(size_t)64 * t + 8 * c + r] = v;
                }
    }

    static void euler_from(const double* R, const double* ref, double* eta) {   // invKinematics.cpp:257-267
        double N[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) N[3 * r + c] = R[3 * r] * ref[c] + R[3 * r + 1] * ref[3 + c] + R[3 * r + 2] * ref[6 + c];
        eta[2] = std::atan2(N[3], N[0]);
        eta[1] = std::atan2(-N[6], std::cos(eta[2]) * N[0] + std::sin(eta[2]) * N[3]);
        eta[0] = std::atan2(std::sin(eta[2]) * N[2] - std::cos(eta[2]) * N[5], -std::sin(eta[2]) * N[1] + std::cos(eta[2]) * N[4]);
    }
    static void omega_map(const double* eta, double* O) {   // generalizedFunctions.cpp:43-50
        double c2 = std::cos(eta[2]), s2 = std::sin(eta[2]), c1 = std::cos(eta[1]), t1 = std::tan(eta[1]);
        O[0] = c2 / c1; O[1] = s2 / c1; O[2] = 0; O[3] = -s2; O[4] = c2; O[5] = 0; O[6] = c2 * t1; O[7] = s2 * t1; O[8] = 1;
    }
    static void inv3(const double* m, double* o) {
        double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[2] * m[7] - m[1] * m[8], c02 = m[1] * m[5] - m[2] * m[4];
        double c10 = m[5] * m[6] - m[3] * m[8], c11 = m[0] * m[8] - m[2] * m[6], c12 = m[2] * m[3] - m[0] * m[5];
        double c20 = m[3] * m[7] - m[4] * m[6], c21 = m[1] * m[6] - m[0] * m[7], c22 = m[0] * m[4] - m[1] * m[3];
        double det = m[0] * c00 + m[1] * c10 + m[2] * c20;
        double c[9] = {c00, c01, c02, c10, c11, c12, c20, c21, c22};
        for (int k = 0; k < 9; ++k) o[k] = c[k] / det;
    }
    void operational_state(const EnvWS& w, double* Q) const {   // invKinematics.cpp:54-70
        for (int s = 0; s < 2; ++s) {
            for (int k = 0; k < 3; ++k) Q[6 * s + k] = w.ps[s][k];
            euler_from(w.Rs[s], K.Rdes, Q + 6 * s + 3);
        }
        for (int k = 0; k < 12; ++k) Q[12 + k] = w.q[18 + k];
        for (int k = 0; k < 3; ++k) Q[24 + k] = w.q[3 + k];
        for (int k = 0; k < 3; ++k) Q[27 + k] = w.com[k];
    }
    // Newton iteration on [feet poses, arm+head joints, base attitude, CoM] (invKinematics.cpp:27-52,151-245)
    void solve_initial_posture() {
        double des[NQ], cur[NQ], q[NQ];
        for (int i = 0; i < NQ; ++i) q[i] = q0[i];
        for (int k = 0; k < 6; ++k) { des[k] = sc.rf[k]; des[6 + k] = sc.lf[k]; }
        for (int k = 0; k < 12; ++k) des[12 + k] = q[18 + k];
        for (int k = 0; k < 3; ++k) des[24 + k] = 0.0;       // desiredOperationalState leaves the attitude slots at zero
        for (int k = 0; k < 3; ++k) des[27 + k] = sc.com[k];
        ik_iterations = 0;
        for (;;) {
            EnvWS& w = evaluate_model(q, nullptr);
            operational_state(w, cur);
            std::vector<double> e(NQ);
            double emax = 0;
            for (int i = 0; i < NQ; ++i) { e[i] = des[i] - cur[i]; emax = std::fmax(emax, std::fabs(e[i])); }
            if (!(emax > 1e-10) || ik_iterations >= 200) break;
            std::vector<double> J((size_t)NQ * NQ, 0.0);
            double Om[9], Omi[9];
            omega_map(w.q + 3, Om);
            inv3(Om, Omi);
            // feet rows: [ang;lin]->[lin;ang], base cols [ang|lin]->[lin|ang], attitude cols through Omega^-1
            for (int s = 0; s < 2; ++s) {
                double etaF[3], OmF[9];
                euler_from(w.Rs[s], K.Rdes, etaF);
                omega_map(etaF, OmF);
                for (int half = 0; half < 2; ++half)          // half 0: linear rows (from JF rows 3..5), half 1: angular rows
                    for (int r = 0; r < 3; ++r) {
                        double row[NQ];
                        for (int c = 0; c < NQ; ++c) {
                            int src_c = c < 3 ? c + 3 : c < 6 ? c - 3 : c;
                            row[c] = jf_full(w, 6 * s + (half == 0 ? 3 + r : r), src_c);
                        }
                        double att[3] = {row[3], row[4], row[5]};
                        for (int c = 0; c < 3; ++c) row[3 + c] = att[0] * Omi[c] + att[1] * Omi[3 + c] + att[2] * Omi[6 + c];
                        for (int c = 0; c < NQ; ++c) J[(size_t)(6 * s + 3 * half + r) * NQ + c] = row[c];
                    }
                // angular rows of this foot premultiplied by the foot's own Omega, attitude columns only (invKinematics.cpp:151-206)
                double blk[9];
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) blk[3 * r + c] = J[(size_t)(6 * s + 3 + r) * NQ + 3 + c];
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
                    J[(size_t)(6 * s + 3 + r) * NQ + 3 + c] = OmF[3 * r] * blk[c] + OmF[3 * r + 1] * blk[3 + c] + OmF[3 * r + 2] * blk[6 + c];
            }
            for (int k = 0; k < 12; ++k) J[(size_t)(12 + k) * NQ + 18 + k] = 1.0;
            for (int k = 0; k < 3; ++k) J[(size_t)(24 + k) * NQ + 3 + k] = 1.0;
            // CoM Jacobian (invKinematics.cpp:208-245)
            {
                Ctx<1> cx1{0, 0};
                fk_phase(cx1, w, K);       // the world transforms were overlaid by the dynamics phases: redo the kinematics
                const double* Tw = w.u.m.p.Tw;
                double Jc[3 * NQ];
                for (int k = 0; k < 3 * NQ; ++k) Jc[k] = 0.0;
                for (int i = 0; i < NF; ++i) {
                    double m = K.mc[4 * i + 3];
                    if (m == 0) continue;
                    const double* Ti = Tw + 12 * i;
                    double pc[3];
                    for (int k = 0; k < 3; ++k) pc[k] = Ti[3 * k] * K.mc[4 * i] + Ti[3 * k + 1] * K.mc[4 * i + 1] + Ti[3 * k + 2] * K.mc[4 * i + 2] + Ti[9 + k];
                    double JX[3 * NQ];
                    for (int k = 0; k < 3 * NQ; ++k) JX[k] = 0.0;
                    for (int k = 0; k < 3; ++k) JX[NQ * k + k] = 1.0;
                    double dp[3] = {Tw[9] - pc[0], Tw[10] - pc[1], Tw[11] - pc[2]};
                    double cm[9] = {0, -dp[2], dp[1], dp[2], 0, -dp[0], -dp[1], dp[0], 0};
                    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) JX[NQ * r + 3 + c] = cm[3 * r + c];
                    for (int j = i; j != 0; j = frame_parent(j)) {
                        int a = frame_joint(j);
                        if (a < 0) continue;
                        const double* Tj = Tw + 12 * j;
                        double z[3] = {Tj[2], Tj[5], Tj[8]};
                        double d[3] = {pc[0] - Tj[9], pc[1] - Tj[10], pc[2] - Tj[11]};
                        JX[NQ * 0 + 6 + a] = z[1] * d[2] - z[2] * d[1];
                        JX[NQ * 1 + 6 + a] = z[2] * d[0] - z[0] * d[2];
                        JX[NQ * 2 + 6 + a] = z[0] * d[1] - z[1] * d[0];
                    }
                    for (int k = 0; k < 3 * NQ; ++k) Jc[k] += m * JX[k];
                }
                for (int k = 0; k < 3 * NQ; ++k) Jc[k] /= K.mass;
                for (int r = 0; r < 3; ++r) {
                    double att[3] = {Jc[NQ * r + 3], Jc[NQ * r + 4], Jc[NQ * r + 5]};
                    for (int c = 0; c < 3; ++c) Jc[NQ * r + 3 + c] = att[0] * Omi[c] + att[1] * Omi[3 + c] + att[2] * Omi[6 + c];
                    for (int c = 0; c < NQ; ++c) J[(size_t)(27 + r) * NQ + c] = Jc[NQ * r + c];
                }
            }
            double dq[NQ];
            if (!dense_solve(NQ, J, e, dq)) throw std::runtime_error("lmh: singular IK Jacobian");
            for (int i = 0; i < NQ; ++i) q[i] += dq[i];
            ++ik_iterations;
        }
        for (int i = 0; i < NQ; ++i) q0[i] = q[i];
        evaluate_model(q0, nullptr);
    }
};

}  // namespace lmh
