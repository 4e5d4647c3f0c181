// ORACLE — TEST INFRASTRUCTURE ONLY. C entry points over orc_model (see orc_capi.h).
#include "orc_capi.h"
#include "orc_model.hpp"
#include <thread>
#include <memory>
#include <cstring>

using namespace orc;

namespace {

struct Sim {
    OrcConfig cfg;
    Robot nao;
    Kinematics ik;
    ZMP zmp;
    std::unique_ptr<Mpc3dLip> mpc;
    FootCoeff rF, lF;
    std::unique_ptr<Controller> ctrl;
    Vec<60> state;
    double t = 0;
    int ik_iters = 0;
    Vec<60> state0;
    // health rule (EXTENSION, see orc_capi.h): closed statistics of the committed steps
    bool alive = true;
    int steps_done = 0;
    double max_tau = 0, max_err = 0, max_qpp = 0;
    long it_total = 0, fail_total = 0;

    explicit Sim(const OrcConfig& c) : cfg(c) {
        // apps/offline/main.cpp:16-58, same order
        zmp = ZMP(cfg.sim_time, cfg.dt, SupportFoot::Double);
        Vec<6> Rf, Lf;
        for (int k = 0; k < 6; ++k) { Rf[k] = cfg.rf[k]; Lf[k] = cfg.lf[k]; }
        Vec<30> desOp = ik.desiredOperationalState(nao, Rf, Lf, v3(cfg.com[0], cfg.com[1], cfg.com[2]));
        ik_iters = ik.compute(nao, desOp);
        mpc.reset(new Mpc3dLip(cfg.dt, cfg.time_horizon, nao.CoM[2]));
        mpc->rebuild_hessian_each_call = cfg.mpc_rebuild_hessian != 0;
        rF = footCoeffTrajectory(v3(cfg.rf[0], cfg.rf[1], cfg.rf[2]), v3(cfg.rf[0], cfg.rf[1], cfg.rf[2]), 0.0, cfg.sim_time);
        lF = footCoeffTrajectory(v3(cfg.lf[0], cfg.lf[1], cfg.lf[2]), v3(cfg.lf[0], cfg.lf[1], cfg.lf[2]), 0.0, cfg.sim_time);
        ctrl.reset(new Controller(nao, *mpc, zmp, rF, lF));
        ctrl->qp_long_double = cfg.qp_long_double != 0;
        state.set(0, 0, nao.q);   // stand(): main.cpp:68-72
        state.set(30, 0, nao.v);
        state0 = state;
    }
    void reset(const double* x0) {
        if (x0) std::memcpy(state.d, x0, sizeof(double) * 60);
        else state = state0;
        t = 0;
        nao.v = Vec<30>();  // a freshly constructed Robot holds zero velocity
        ctrl->qp_solves = ctrl->qp_iterations = ctrl->qp_failures = 0;
        alive = true; steps_done = 0; max_tau = max_err = max_qpp = 0; it_total = fail_total = 0;
    }
    void fill_dump(const Vec<60>& xdot, OrcStageDump* o) {
        Controller& c = *ctrl;
        std::memcpy(o->M, c.dyn.M.d, sizeof(o->M));
        std::memcpy(o->C, c.dyn.C.d, sizeof(o->C));
        std::memcpy(o->Cg, c.dyn.Cg.d, sizeof(o->Cg));
        std::memcpy(o->AG, c.dyn.AG.d, sizeof(o->AG));
        std::memcpy(o->AGpqp, c.dyn.AGpqp.d, sizeof(o->AGpqp));
        std::memcpy(o->Jpqp, c.dyn.Jpqp.d, sizeof(o->Jpqp));
        std::memcpy(o->JFeet, c.kin.JFeet.d, sizeof(o->JFeet));
        std::memcpy(o->CoM, nao.CoM.d, sizeof(o->CoM));
        std::memcpy(o->comVel, nao.comVel.d, sizeof(o->comVel));
        std::memcpy(o->comAngMom, nao.comAngMom.d, sizeof(o->comAngMom));
        std::memcpy(o->xRef, mpc->xRef.d, sizeof(o->xRef));
        std::memcpy(o->yRef, mpc->yRef.d, sizeof(o->yRef));
        std::memcpy(o->qppRef, c.qppRef.d, sizeof(o->qppRef));
        std::memcpy(o->hGpRef, c.hGpRef.d, sizeof(o->hGpRef));
        std::memcpy(o->footAccRef, c.footAccRef.d, sizeof(o->footAccRef));
        std::memcpy(o->H, c.Hqp.d, sizeof(o->H));
        std::memcpy(o->g, c.gqp.d, sizeof(o->g));
        std::memcpy(o->x, c.xqp.d, sizeof(o->x));
        std::memcpy(o->y, c.yqp.d, sizeof(o->y));
        std::memcpy(o->T7, nao.T[7].d, sizeof(o->T7));
        std::memcpy(o->T14, nao.T[14].d, sizeof(o->T14));
        std::memcpy(o->xdot, xdot.d, sizeof(o->xdot));
        qp_kkt_residuals(74, c.Hqp.d, c.gqp.d, 50, c.Aqp.d, c.lbA.d, c.ubA.d, c.xqp.d, c.yqp.d, o->kkt);
        o->k_index = mpc->last_k;
        o->qp_status = c.qpInfo.status;
        o->qp_iters = c.qpInfo.iterations;
        o->pad = 0;
    }
    // one RK4 step with the bookkeeping the logs need (rk4.hpp:12-17 inlined) and the health rule of orc_capi.h
    void step(OrcStepLog* log) {
        if (!alive) {   // fallen: nothing is evaluated any more; the record repeats the last evaluated stage
            if (log) fill_log(log, nullptr, nullptr, nullptr, nullptr, 0, 0);
            return;
        }
        const double dt = cfg.dt;
        const bool two = cfg.second_wbc != 0;
        long it0 = ctrl->qp_iterations, f0 = ctrl->qp_failures;
        WBCOutput o1;
        int kk[4], ph[4], ct[4];
        auto note = [&](int i) {
            kk[i] = mpc->last_k; ph[i] = ctrl->last_phase;
            ct[i] = ctrl->contact.empty() ? 3 : ctrl->contact[ctrl->last_phase];
        };
        Vec<60> k1 = dynamics(state, t, nao, *ctrl, two, &o1); note(0);
        Vec<60> k2 = dynamics(state + (0.5 * dt) * k1, t + 0.5 * dt, nao, *ctrl, two); note(1);
        Vec<60> k3 = dynamics(state + (0.5 * dt) * k2, t + 0.5 * dt, nao, *ctrl, two); note(2);
        Vec<60> k4 = dynamics(state + dt * k3, t + dt, nao, *ctrl, two); note(3);
        Vec<60> xn = state + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
        double tau_inf = 0, qpp_inf = 0;
        for (int k = 0; k < 24; ++k) tau_inf = std::fmax(tau_inf, std::fabs(o1.tau[k]));
        for (int k = 0; k < 30; ++k) qpp_inf = std::fmax(qpp_inf, std::fabs(o1.qpp[k]));
        int kz = kk[3] < 0 ? 0 : kk[3];
        const int zl = (int)ctrl->zmp.zmpXRef.size();
        if (kz >= zl) kz = zl - 1;
        const double ex = nao.CoM[0] - ctrl->zmp.zmpXRef[kz], ey = nao.CoM[1] - ctrl->zmp.zmpYRef[kz];
        const double err = std::sqrt(ex * ex + ey * ey);
        bool healthy = tau_inf < ORC_TAU_LIMIT && err < ORC_COM_ERR_LIMIT && xn[2] > ORC_TORSO_MIN;
        for (int k = 0; k < 60; ++k) healthy = healthy && (std::fabs(xn[k]) < ORC_STATE_LIMIT);
        const int dit = (int)(ctrl->qp_iterations - it0), dfl = (int)(ctrl->qp_failures - f0);
        if (healthy) {
            state = xn;
            t += dt;  // Clock::step
            max_tau = std::fmax(max_tau, tau_inf); max_qpp = std::fmax(max_qpp, qpp_inf); max_err = std::fmax(max_err, err);
            it_total += dit; fail_total += dfl;
            ++steps_done;
        } else {
            alive = false;
        }
        if (log) fill_log(log, &o1, kk, ph, ct, dit, dfl);
    }
    void fill_log(OrcStepLog* log, const WBCOutput* o1, const int* kk, const int* ph, const int* ct, int dit, int dfl) {
        std::memcpy(log->x, state.d, sizeof(log->x));
        for (int k = 0; k < 3; ++k) {
            log->com_k4[k] = nao.CoM[k];
            log->rfoot[k] = nao.T[7](k, 3);
            log->lfoot[k] = nao.T[14](k, 3);
        }
        if (o1) std::memcpy(log->tau_k1, o1->tau.d, sizeof(log->tau_k1));
        else std::memset(log->tau_k1, 0, sizeof(log->tau_k1));
        for (int i = 0; i < 4; ++i) {
            log->k_index[i] = kk ? kk[i] : mpc->last_k;
            log->phase[i] = ph ? ph[i] : ctrl->last_phase;
            log->contact[i] = ct ? ct[i] : (ctrl->contact.empty() ? 3 : ctrl->contact[ctrl->last_phase]);
        }
        log->qp_iters = dit;
        log->qp_fail = dfl;
    }
};

}  // namespace

extern "C" {

void orc_default_config(OrcConfig* c) {
    std::memset(c, 0, sizeof(*c));
    c->dt = 0.01; c->sim_time = 5.0; c->time_horizon = 0.5;
    c->com[0] = -0.02; c->com[1] = 0.0; c->com[2] = 0.26;
    c->rf[1] = -0.05; c->lf[1] = 0.05;
    c->second_wbc = 0;
}

void* orc_sim_create(const OrcConfig* cfg) { return new Sim(*cfg); }
void orc_sim_destroy(void* s) { delete static_cast<Sim*>(s); }
int32_t orc_sim_ik_iterations(void* s) { return static_cast<Sim*>(s)->ik_iters; }
double orc_sim_mass(void* s) { return static_cast<Sim*>(s)->nao.mass; }
double orc_sim_zcom(void* s) { return static_cast<Sim*>(s)->mpc->zCom; }

void orc_sim_get_state(void* sp, double* x, double* t, double* vh) {
    Sim* s = static_cast<Sim*>(sp);
    if (x) std::memcpy(x, s->state.d, sizeof(double) * 60);
    if (t) *t = s->t;
    if (vh) std::memcpy(vh, s->nao.v.d, sizeof(double) * 30);
}
void orc_sim_set_state(void* sp, const double* x, double t, const double* vh) {
    Sim* s = static_cast<Sim*>(sp);
    if (x) std::memcpy(s->state.d, x, sizeof(double) * 60);
    s->t = t;
    if (vh) std::memcpy(s->nao.v.d, vh, sizeof(double) * 30);
}

void orc_sim_stage(void* sp, const double* x60, double t, OrcStageDump* out) {
    Sim* s = static_cast<Sim*>(sp);
    Vec<60> x;
    std::memcpy(x.d, x60, sizeof(double) * 60);
    WBCOutput o;
    Vec<60> xd = dynamics(x, t, s->nao, *s->ctrl, s->cfg.second_wbc != 0, &o);
    s->fill_dump(xd, out);
    std::memcpy(out->qpp, o.qpp.d, sizeof(out->qpp));
    std::memcpy(out->tau, o.tau.d, sizeof(out->tau));
    std::memcpy(out->f, o.f.d, sizeof(out->f));
}

void orc_sim_step(void* sp, int32_t n, OrcStepLog* log) {
    Sim* s = static_cast<Sim*>(sp);
    for (int i = 0; i < n; ++i) s->step(log ? log + i : nullptr);
}

int32_t orc_sim_run_stand(void* sp, double* comx, int32_t max_steps) {
    Sim* s = static_cast<Sim*>(sp);
    int n = 0;
    while (std::fabs(s->t - s->cfg.sim_time) > 0.01 && n < max_steps) {
        s->step(nullptr);
        if (comx) comx[n] = s->nao.CoM[0];
        ++n;
    }
    return n;
}

void orc_sim_mpc_tables(void* sp, double* Px, double* Pu, int32_t* horizon) {
    Sim* s = static_cast<Sim*>(sp);
    int n = s->mpc->horizon + 1;
    if (horizon) *horizon = s->mpc->horizon;
    if (Px) std::memcpy(Px, s->mpc->Px.data(), sizeof(double) * n * 2);
    if (Pu) std::memcpy(Pu, s->mpc->Pu.data(), sizeof(double) * n * n);
}

void orc_sim_foot_coeffs(void* sp, double* rx, double* ry, double* rz, double* lx, double* ly, double* lz) {
    Sim* s = static_cast<Sim*>(sp);
    std::memcpy(rx, s->rF.x.data(), 6 * sizeof(double));
    std::memcpy(ry, s->rF.y.data(), 6 * sizeof(double));
    std::memcpy(rz, s->rF.z.data(), 8 * sizeof(double));
    std::memcpy(lx, s->lF.x.data(), 6 * sizeof(double));
    std::memcpy(ly, s->lF.y.data(), 6 * sizeof(double));
    std::memcpy(lz, s->lF.z.data(), 8 * sizeof(double));
}

int32_t orc_sim_zmp(void* sp, double* zx, double* zy, int32_t cap) {
    Sim* s = static_cast<Sim*>(sp);
    int n = (int)s->zmp.zmpXRef.size();
    for (int i = 0; i < n && i < cap; ++i) { zx[i] = s->zmp.zmpXRef[i]; zy[i] = s->zmp.zmpYRef[i]; }
    return n;
}

void orc_sim_link_table(void* sp, double* mass, double* com, double* inertia) {
    Sim* s = static_cast<Sim*>(sp);
    for (int i = 0; i < NUM_FRAMES; ++i) {
        mass[i] = s->nao.links[i].mass;
        for (int k = 0; k < 3; ++k) com[3 * i + k] = s->nao.links[i].com[k];
        for (int k = 0; k < 9; ++k) inertia[9 * i + k] = s->nao.links[i].inertia.d[k];
    }
}

void orc_sim_frames(void* sp, double* T, double* X) {
    Sim* s = static_cast<Sim*>(sp);
    for (int i = 0; i < NUM_FRAMES; ++i) {
        if (T) std::memcpy(T + 16 * i, s->nao.T[i].d, sizeof(double) * 16);
        if (X) std::memcpy(X + 36 * i, s->nao.X[i].d, sizeof(double) * 36);
    }
}

void orc_sim_set_walk_plan(void* sp, int32_t zmp_len, const double* zx, const double* zy, int32_t n_phases, double phase_T,
                           const double* fc, const int32_t* contact) {
    Sim* s = static_cast<Sim*>(sp);
    Controller& c = *s->ctrl;
    c.zmp.zmpXRef.assign(zx, zx + zmp_len);
    c.zmp.zmpYRef.assign(zy, zy + zmp_len);
    c.phase_T = phase_T;
    c.contact.assign(contact, contact + n_phases);
    c.rFp.resize(n_phases); c.lFp.resize(n_phases);
    for (int p = 0; p < n_phases; ++p)
        for (int f = 0; f < 2; ++f) {
            const double* k = fc + ((size_t)p * 2 + f) * 20;
            FootCoeff& d = f == 0 ? c.rFp[p] : c.lFp[p];
            d.x.assign(k, k + 6); d.y.assign(k + 6, k + 12); d.z.assign(k + 12, k + 20);
        }
}
int32_t orc_sim_last_phase(void* sp) { return static_cast<Sim*>(sp)->ctrl->last_phase; }
int32_t orc_sim_alive(void* sp) { return static_cast<Sim*>(sp)->alive ? 1 : 0; }
int32_t orc_sim_steps(void* sp) { return static_cast<Sim*>(sp)->steps_done; }

int32_t orc_batch_rollout_ex(const OrcConfig* cfg, int32_t E, const double* x0, const int32_t* push_step,
                             const double* push_dv, const OrcWalkPlans* plans, int32_t n_steps, int32_t n_threads,
                             double* x_out, double* com_out, double* stats8, OrcStepLog* log) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > E) n_threads = E > 0 ? E : 1;
    std::vector<std::thread> pool;
    std::vector<int> fails(n_threads, 0);
    for (int th = 0; th < n_threads; ++th) {
        pool.emplace_back([&, th]() {
            int lo = (int)((long long)E * th / n_threads), hi = (int)((long long)E * (th + 1) / n_threads);
            if (lo >= hi) return;
            Sim sim(*cfg);
            sim.ctrl->abort_on_nonfinite = false;
            for (int e = lo; e < hi; ++e) {
                sim.reset(x0 ? x0 + 60 * (size_t)e : nullptr);
                if (plans)
                    orc_sim_set_walk_plan(&sim, plans->zmp_len, plans->zmp_x + (size_t)e * plans->zmp_len,
                                          plans->zmp_y + (size_t)e * plans->zmp_len, plans->n_phases, plans->phase_T,
                                          plans->foot_coef + (size_t)e * plans->n_phases * 40,
                                          plans->contact + (size_t)e * plans->n_phases);
                for (int s = 0; s < n_steps; ++s) {
                    if (sim.alive && push_step && push_step[e] == s && push_dv)
                        for (int k = 0; k < 6; ++k) sim.state[30 + k] += push_dv[6 * (size_t)e + k];
                    sim.step(log ? log + (size_t)e * n_steps + s : nullptr);
                }
                if (x_out) std::memcpy(x_out + 60 * (size_t)e, sim.state.d, sizeof(double) * 60);
                if (com_out)
                    for (int k = 0; k < 3; ++k) com_out[3 * (size_t)e + k] = sim.nao.CoM[k];
                if (stats8) {
                    double* o = stats8 + 8 * (size_t)e;
                    o[0] = sim.max_tau; o[1] = sim.max_err; o[2] = sim.max_qpp; o[3] = (double)sim.it_total;
                    o[4] = (double)sim.fail_total; o[5] = (double)sim.steps_done; o[6] = sim.alive ? 1.0 : 0.0; o[7] = sim.t;
                }
                fails[th] += (int)sim.fail_total;
            }
        });
    }
    for (auto& t : pool) t.join();
    int total = 0;
    for (int f : fails) total += f;
    return total;
}

int32_t orc_batch_rollout(const OrcConfig* cfg, int32_t E, const double* x0, const int32_t* push_step,
                          const double* push_dv, int32_t n_steps, int32_t n_threads, double* x_out,
                          double* com_out, double* stats) {
    std::vector<double> s8(stats ? (size_t)E * 8 : 0);
    int total = orc_batch_rollout_ex(cfg, E, x0, push_step, push_dv, nullptr, n_steps, n_threads, x_out, com_out,
                                     stats ? s8.data() : nullptr, nullptr);
    if (stats)
        for (int e = 0; e < E; ++e) {
            stats[4 * (size_t)e + 0] = s8[8 * (size_t)e + 0];
            stats[4 * (size_t)e + 1] = s8[8 * (size_t)e + 1];
            stats[4 * (size_t)e + 2] = s8[8 * (size_t)e + 3];
            stats[4 * (size_t)e + 3] = s8[8 * (size_t)e + 4];
        }
    return total;
}

int32_t orc_qp_solve(int32_t n, const double* H, const double* g, int32_t m, const double* A, const double* lbA,
                     const double* ubA, double* x, double* y, int32_t ld, int32_t* iters, double* kkt4) {
    std::vector<double> yy(m > 0 ? m : 1);
    QpInfo info = ld ? qp_solve_gi<long double>(n, H, g, m, A, lbA, ubA, x, yy.data(), 2000)
                     : qp_solve_gi<double>(n, H, g, m, A, lbA, ubA, x, yy.data(), 2000);
    if (y) std::memcpy(y, yy.data(), sizeof(double) * m);
    if (iters) *iters = info.iterations;
    if (kkt4) qp_kkt_residuals(n, H, g, m, A, lbA, ubA, x, yy.data(), kkt4);
    return info.status;
}

}  // extern "C"
