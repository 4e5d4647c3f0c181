// ORACLE — TEST INFRASTRUCTURE ONLY.
// C entry points (oracle/orc_capi.h) over the REFERENCE'S OWN classes, compiled from the
// sources where they lie under /root/reference (src/*.cpp + apps/offline/main.cpp), against the
// shim headers next to this file (a minimal Eigen subset, and the qpOASES API routed to
// oracle/orc_qp.hpp).  Nothing of the reference is copied: main.cpp is #included in place so that
// its dynamics() closure (main.cpp:91-122) and rk4Step (rk4.hpp:5-18) are the reference's text.
// What this pins: everything except the two third-party boundaries (Eigen arithmetic order and the
// qpOASES solve), which stay PARITY UNPINNED because those libraries are absent from this image.
#include <vector>
#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>
#include <thread>
#include <memory>
#include <algorithm>
#include <type_traits>
#include <cassert>
#include <cstddef>
#include <ostream>

// test-only access to the members the dumps need (Hqp_, g_, PD*Acc, T_, ...)
#define private public
#define protected public
#define main lmh_ref_offline_main_unused
#include REF_OFFLINE_MAIN
#undef main
#undef private
#undef protected

#include "orc_capi.h"

namespace {

struct NullBuf : std::streambuf {
    int overflow(int c) override { return c; }
};
NullBuf g_nullbuf;

struct RefSim {
    OrcConfig cfg;
    Robot nao;
    Kinematics ik;
    Clock clock;
    ZMP zmp;
    std::unique_ptr<Mpc3dLip> mpc;
    std::vector<Eigen::VectorXd> rF, lF;
    std::unique_ptr<Controller> ctrl;
    Eigen::VectorXd state, state0;
    double t = 0;
    WBCOutput last;

    explicit RefSim(const OrcConfig& c)
        : cfg(c), clock(c.dt, c.sim_time), zmp(Task::Stand, c.sim_time, c.dt, SupportFoot::Double) {
        // apps/offline/main.cpp:16-58
        Eigen::VectorXd Rf = Eigen::VectorXd::Zero(6), Lf = Eigen::VectorXd::Zero(6);
        for (int k = 0; k < 6; ++k) { Rf(k) = cfg.rf[k]; Lf(k) = cfg.lf[k]; }
        Eigen::Vector3d com(cfg.com[0], cfg.com[1], cfg.com[2]);
        Eigen::VectorXd desOp = ik.desiredOperationalState(nao, Rf, Lf, com);
        ik.compute(nao, desOp);
        mpc.reset(new Mpc3dLip(clock.getTimeStep(), cfg.time_horizon, nao.getCoM()(2)));
        Eigen::Vector3d pr(cfg.rf[0], cfg.rf[1], cfg.rf[2]), pl(cfg.lf[0], cfg.lf[1], cfg.lf[2]);
        rF = footCoeffTrajectory(pr, pr, 0.0, cfg.sim_time);
        lF = footCoeffTrajectory(pl, pl, 0.0, cfg.sim_time);
        ctrl.reset(new Controller(nao, *mpc, zmp, rF, lF));
        state = Eigen::VectorXd::Zero(60);
        state.segment(0, 30) = nao.getJoints();
        state.segment(30, 30) = nao.getJointsVelocity();
        state0 = state;
    }
    Eigen::VectorXd f(const Eigen::VectorXd& x, double tt) { return dynamics(x, tt, nao, *ctrl); }
    void step(OrcStepLog* log) {
        double dt = cfg.dt;
        int kk[4] = {(int)(t / dt), (int)((t + 0.5 * dt) / dt), (int)((t + 0.5 * dt) / dt), (int)((t + dt) / dt)};
        Eigen::VectorXd tau1;
        int call = 0;  // torques of the k1 stage are read inside rk4Step's first call of f
        state = rk4Step(
            [&](const Eigen::VectorXd& x, double tt) {
                Eigen::VectorXd xd = f(x, tt);
                if (call++ == 0) tau1 = ctrl->getTorques();
                return xd;
            },
            state, t, dt);
        t += dt;
        if (log) {
            for (int k = 0; k < 60; ++k) log->x[k] = state(k);
            for (int k = 0; k < 3; ++k) {
                log->com_k4[k] = nao.getCoM()(k);
                log->rfoot[k] = nao.getT()[7](k, 3);
                log->lfoot[k] = nao.getT()[14](k, 3);
            }
            for (int k = 0; k < 24; ++k) log->tau_k1[k] = tau1(k);
            for (int k = 0; k < 4; ++k) { log->k_index[k] = kk[k]; log->phase[k] = 0; log->contact[k] = 3; }   // stand task
            log->qp_iters = 0;
            log->qp_fail = 0;
        }
    }
};

void cp(double* dst, const Eigen::Dense& m, bool rowmajor_out = true) {
    if (m.rows() == 1 || m.cols() == 1) {
        for (Eigen::Index i = 0; i < m.size(); ++i) dst[i] = m(i);
        return;
    }
    for (Eigen::Index i = 0; i < m.rows(); ++i)
        for (Eigen::Index j = 0; j < m.cols(); ++j)
            dst[rowmajor_out ? i * m.cols() + j : j * m.rows() + i] = m(i, j);
}

}  // namespace

extern "C" {

void orc_default_config(OrcConfig* c) {
    std::memset(c, 0, sizeof(*c));
    c->dt = 0.01; c->sim_time = 5.0; c->time_horizon = 0.5;
    c->com[0] = -0.02; c->com[1] = 0.0; c->com[2] = 0.26;
    c->rf[1] = -0.05; c->lf[1] = 0.05;
    c->second_wbc = 1;            // the reference always runs the duplicate WBC (main.cpp:105)
    c->mpc_rebuild_hessian = 1;   // and always rebuilds the MPC Hessian (mpcLinearPendulum.cpp:89-90)
}

void* orc_sim_create(const OrcConfig* cfg) {
    std::streambuf* old = std::cout.rdbuf(&g_nullbuf);  // the Controller ctor prints banners
    RefSim* s = new RefSim(*cfg);
    std::cout.rdbuf(old);
    return s;
}
void orc_sim_destroy(void* s) { delete static_cast<RefSim*>(s); }
int32_t orc_sim_ik_iterations(void*) { return -1; }
double orc_sim_mass(void* s) { return static_cast<RefSim*>(s)->nao.getMass(); }
double orc_sim_zcom(void* s) { return static_cast<RefSim*>(s)->mpc->getZCom(); }

void orc_sim_get_state(void* sp, double* x, double* t, double* vh) {
    RefSim* s = static_cast<RefSim*>(sp);
    if (x) cp(x, s->state);
    if (t) *t = s->t;
    if (vh) cp(vh, s->nao.v_);
}
void orc_sim_set_state(void* sp, const double* x, double t, const double* vh) {
    RefSim* s = static_cast<RefSim*>(sp);
    if (x) for (int k = 0; k < 60; ++k) s->state(k) = x[k];
    s->t = t;
    if (vh) for (int k = 0; k < 30; ++k) s->nao.v_(k) = vh[k];
}

void orc_sim_stage(void* sp, const double* x60, double t, OrcStageDump* o) {
    RefSim* s = static_cast<RefSim*>(sp);
    std::memset(o, 0, sizeof(*o));
    Eigen::VectorXd x = Eigen::VectorXd::Zero(60);
    for (int k = 0; k < 60; ++k) x(k) = x60[k];
    Eigen::VectorXd xd = dynamics(x, t, s->nao, *s->ctrl);
    Controller& c = *s->ctrl;
    cp(o->M, c.dyn_.getM()); cp(o->C, c.dyn_.getC()); cp(o->Cg, c.dyn_.getCg());
    cp(o->AG, c.dyn_.getAG()); cp(o->AGpqp, c.dyn_.getAGpqp()); cp(o->Jpqp, c.dyn_.getJpqp());
    cp(o->JFeet, c.kin_.getFeetJacobian());
    cp(o->CoM, s->nao.getCoM()); cp(o->comVel, s->nao.getComVel()); cp(o->comAngMom, s->nao.getComAngMom());
    cp(o->xRef, s->mpc->getXRef()); cp(o->yRef, s->mpc->getYRef());
    cp(o->qppRef, c.PDJointsAcc()); cp(o->hGpRef, c.PDMomentumAcc()); cp(o->footAccRef, c.PDFeetAcc(t));
    cp(o->H, c.Hqp_); cp(o->g, c.g_);
    c.qp_.getPrimalSolution(o->x);
    for (int k = 0; k < 50; ++k) o->y[k] = c.qp_.y_[(size_t)k];
    WBCOutput w = c.WBC(t);  // third identical solve: the only public way to read qpp/tau/f back
    cp(o->qpp, w.qpp); cp(o->tau, w.tau); cp(o->f, w.f);
    cp(o->T7, s->nao.getT()[7]); cp(o->T14, s->nao.getT()[14]);
    cp(o->xdot, xd);
    orc::qp_kkt_residuals(74, o->H, o->g, 50, c.A_.data(), c.lbA_.data(), c.ubA_.data(), o->x, o->y, o->kkt);
    o->k_index = (int)(t / s->cfg.dt);
    o->qp_status = c.qp_.lastInfo.status;
    o->qp_iters = c.qp_.lastInfo.iterations;
}

void orc_sim_step(void* sp, int32_t n, OrcStepLog* log) {
    RefSim* s = static_cast<RefSim*>(sp);
    for (int i = 0; i < n; ++i) s->step(log ? log + i : nullptr);
}

int32_t orc_sim_run_stand(void* sp, double* comx, int32_t max_steps) {
    RefSim* s = static_cast<RefSim*>(sp);
    // stand() of main.cpp:66-89 run in place on the reference objects; its stdout (CoM x per step,
    // main.cpp:86) is captured and parsed back so that the stream itself is what gets compared.
    s->clock = Clock(s->cfg.dt, s->cfg.sim_time);
    s->clock.t_ = s->t;
    std::ostringstream cap;
    std::streambuf* old = std::cout.rdbuf(cap.rdbuf());
    stand(s->nao, *s->ctrl, s->clock);
    std::cout.rdbuf(old);
    s->t = s->clock.getTime();
    std::istringstream in(cap.str());
    int n = 0;
    double v;
    while (in >> v) {
        if (comx && n < max_steps) comx[n] = v;
        ++n;
    }
    return n < max_steps ? n : max_steps;
}

void orc_sim_mpc_tables(void* sp, double* Px, double* Pu, int32_t* horizon) {
    RefSim* s = static_cast<RefSim*>(sp);
    if (horizon) *horizon = s->mpc->horizon_;
    if (Px) cp(Px, s->mpc->Px_);
    if (Pu) cp(Pu, s->mpc->Pu_);
}

void orc_sim_foot_coeffs(void* sp, double* rx, double* ry, double* rz, double* lx, double* ly, double* lz) {
    RefSim* s = static_cast<RefSim*>(sp);
    cp(rx, s->rF[0]); cp(ry, s->rF[1]); cp(rz, s->rF[2]);
    cp(lx, s->lF[0]); cp(ly, s->lF[1]); cp(lz, s->lF[2]);
}

int32_t orc_sim_zmp(void* sp, double* zx, double* zy, int32_t cap) {
    RefSim* s = static_cast<RefSim*>(sp);
    int n = (int)s->zmp.getZmpXRef().size();
    for (int i = 0; i < n && i < cap; ++i) { zx[i] = s->zmp.getZmpXRef()(i); zy[i] = s->zmp.getZmpYRef()(i); }
    return n;
}

void orc_sim_link_table(void* sp, double* mass, double* com, double* inertia) {
    RefSim* s = static_cast<RefSim*>(sp);
    const std::vector<linkInertia>& L = s->nao.getLinks();
    for (int i = 0; i < 28; ++i) {
        mass[i] = L[(size_t)i].mass;
        for (int k = 0; k < 3; ++k) com[3 * i + k] = L[(size_t)i].com(k);
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc) inertia[9 * i + 3 * r + cc] = L[(size_t)i].inertia(r, cc);
    }
}

void orc_sim_frames(void* sp, double* T, double* X) {
    RefSim* s = static_cast<RefSim*>(sp);
    for (int i = 0; i < 28; ++i) {
        if (T) cp(T + 16 * i, s->nao.getT()[(size_t)i]);
        if (X) cp(X + 36 * i, s->nao.getX()[(size_t)i]);
    }
}

int32_t orc_batch_rollout(const OrcConfig* cfg, int32_t E, const double* x0, const int32_t* push_step,
                          const double* push_dv, int32_t n_steps, int32_t n_threads, double* x_out,
                          double* com_out, double* stats) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > E) n_threads = E > 0 ? E : 1;
    std::streambuf* old = std::cout.rdbuf(&g_nullbuf);
    std::vector<std::thread> pool;
    for (int th = 0; th < n_threads; ++th) {
        pool.emplace_back([&, th]() {
            int lo = (int)((long long)E * th / n_threads), hi = (int)((long long)E * (th + 1) / n_threads);
            if (lo >= hi) return;
            RefSim sim(*cfg);
            for (int e = lo; e < hi; ++e) {
                if (x0) for (int k = 0; k < 60; ++k) sim.state(k) = x0[60 * (size_t)e + k];
                else sim.state = sim.state0;
                sim.t = 0;
                sim.nao.v_ = Eigen::VectorXd::Zero(30);
                double max_tau = 0, max_err = 0;
                for (int s = 0; s < n_steps; ++s) {
                    if (push_step && push_step[e] == s && push_dv)
                        for (int k = 0; k < 6; ++k) sim.state(30 + k) += push_dv[6 * (size_t)e + k];
                    OrcStepLog lg;
                    sim.step(stats ? &lg : nullptr);
                    if (stats) {
                        for (int k = 0; k < 24; ++k) max_tau = std::fmax(max_tau, std::fabs(lg.tau_k1[k]));
                        max_err = std::fmax(max_err, std::sqrt(lg.com_k4[0] * lg.com_k4[0] + lg.com_k4[1] * lg.com_k4[1]));
                    }
                }
                if (x_out) for (int k = 0; k < 60; ++k) x_out[60 * (size_t)e + k] = sim.state(k);
                if (com_out) for (int k = 0; k < 3; ++k) com_out[3 * (size_t)e + k] = sim.nao.getCoM()(k);
                if (stats) {
                    stats[4 * (size_t)e + 0] = max_tau;
                    stats[4 * (size_t)e + 1] = max_err;
                    stats[4 * (size_t)e + 2] = 0;
                    stats[4 * (size_t)e + 3] = 0;
                }
            }
        });
    }
    for (auto& t : pool) t.join();
    std::cout.rdbuf(old);
    return 0;
}

}  // extern "C"
