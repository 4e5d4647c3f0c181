// TEST TOOL -- host build (NT = 1, sequential loops) of the device source lmh_model.cuh, so that the
// kernel logic can be unit-tested against the oracle on machines without a GPU.  Not linked into, loaded
// by, or reachable from the product library.
#include "lmh_setup.hpp"
#include <cstring>

using namespace lmh;

struct Emu {
    HostModel hm;
    Plan plan;
    int32_t contact_both = 3;
    std::vector<double> wzx, wzy, wfc;
    std::vector<int32_t> wct;
    explicit Emu(const Scenario& s) : hm(s) {
        plan.zmp_x = hm.zmp_x.data();
        plan.zmp_y = hm.zmp_y.data();
        plan.foot_coef = hm.foot_coef.data();
        plan.zmp_len = (int32_t)hm.zmp_x.size();
        plan.contact = &contact_both; plan.n_phases = 1; plan.phase_T = 1e30;
    }
};

extern "C" {
void* emu_create(double dt, double sim_time, double horizon, const double* com3) {
    Scenario s;
    s.dt = dt; s.sim_time = sim_time; s.time_horizon = horizon;
    if (com3) for (int k = 0; k < 3; ++k) s.com[k] = com3[k];
    return new Emu(s);
}
void emu_destroy(void* p) { delete static_cast<Emu*>(p); }
void emu_set_walk_plan(void* p, int32_t zmp_len, const double* zx, const double* zy, int32_t n_phases, double phase_T,
                       const double* fc, const int32_t* contact) {
    Emu* e = static_cast<Emu*>(p);
    e->wzx.assign(zx, zx + zmp_len); e->wzy.assign(zy, zy + zmp_len);
    e->wfc.assign(fc, fc + (size_t)n_phases * 40); e->wct.assign(contact, contact + n_phases);
    e->plan.zmp_x = e->wzx.data(); e->plan.zmp_y = e->wzy.data(); e->plan.foot_coef = e->wfc.data();
    e->plan.contact = e->wct.data(); e->plan.zmp_len = zmp_len; e->plan.n_phases = n_phases; e->plan.phase_T = phase_T;
}
void emu_nominal(void* p, double* q0, double* zcom, double* mass, int32_t* ik_iters) {
    Emu* e = static_cast<Emu*>(p);
    for (int i = 0; i < NQ; ++i) q0[i] = e->hm.q0[i];
    *zcom = e->hm.K.mpc_zcom; *mass = e->hm.K.mass; *ik_iters = e->hm.ik_iterations;
}
void emu_tables(void* p, double* mass28, double* com28x3, double* isp28x36, double* mpc_r, int32_t* n1, double* rPx, double* foot40) {
    Emu* e = static_cast<Emu*>(p);
    for (int i = 0; i < 28; ++i) { mass28[i] = e->hm.K.mc[4 * i + 3]; for (int k = 0; k < 3; ++k) com28x3[3 * i + k] = e->hm.K.mc[4 * i + k]; }
    std::memcpy(isp28x36, e->hm.K.Isp, sizeof(double) * 28 * 36);
    *n1 = e->hm.K.mpc_n1;
    for (int i = 0; i < e->hm.K.mpc_n1; ++i) mpc_r[i] = e->hm.K.mpc_r[i];
    rPx[0] = e->hm.K.mpc_rPx[0]; rPx[1] = e->hm.K.mpc_rPx[1];
    std::memcpy(foot40, e->hm.foot_coef.data(), sizeof(double) * 40);
}
// row `mask` of the QP working-set table (lmh_setup.hpp: Kf 36, Yf 36) and what it was built from
void emu_qp_table_row(void* p, int32_t mask, double* out72, double* Gf96, double* wf_eps2) {
    Emu* e = static_cast<Emu*>(p);
    std::memcpy(out72, e->hm.K.qp_tab + (size_t)mask * 72, sizeof(double) * 72);
    std::memcpy(Gf96, e->hm.K.Gf, sizeof(double) * 96);
    wf_eps2[0] = e->hm.K.wForce; wf_eps2[1] = e->hm.K.epsC;
}
// one dynamics(x,t) evaluation; vheld is read and advanced
void emu_stage(void* p, const double* x60, double* vheld30, double t, LmhStageDump* dump) {
    Emu* e = static_cast<Emu*>(p);
    EnvWS& w = e->hm.workspace();
    for (int i = 0; i < NQ; ++i) { w.q[i] = x60[i]; w.v[i] = x60[NQ + i]; w.vh[i] = vheld30[i]; }
    w.t = t;
    Ctx<1> cx{0, 0};
    ws_reset(cx, w);
    stage_eval(cx, w, e->hm.K, e->plan, 0, dump);
    for (int i = 0; i < NQ; ++i) vheld30[i] = w.vh[i];
}
void emu_rollout(void* p, double* x60, double* vheld30, double* t, int32_t n_steps, int32_t push_step, const double* push_dv6,
                 LmhStepLog* logs, LmhStats* stats);
void emu_step(void* p, double* x60, double* vheld30, double* t, int32_t n_steps, LmhStepLog* logs, LmhStats* stats) {
    emu_rollout(p, x60, vheld30, t, n_steps, -1, nullptr, logs, stats);
}
// n RK4 steps with an optional velocity impulse at the start of step push_step (what the rollout kernels do per env)
void emu_rollout(void* p, double* x60, double* vheld30, double* t, int32_t n_steps, int32_t push_step, const double* push_dv6,
                 LmhStepLog* logs, LmhStats* stats) {
    Emu* e = static_cast<Emu*>(p);
    EnvWS& w = e->hm.workspace();
    for (int i = 0; i < NQ; ++i) w.vh[i] = vheld30[i];
    double kacc[60];
    StatsAcc st{0, 0, 0, 0, 0, 0, 1, 0};
    Ctx<1> cx{0, 0};
    ws_reset(cx, w);
    for (int s = 0; s < n_steps; ++s) {
        if (s == push_step && push_dv6 && w.dead == 0) for (int k = 0; k < 6; ++k) x60[30 + k] += push_dv6[k];
        rk4_step(cx, w, x60, kacc, *t, e->hm.K, e->plan, 0, logs ? logs + s : nullptr, &st);
    }
    for (int i = 0; i < NQ; ++i) vheld30[i] = w.vh[i];
    if (stats) {
        stats->max_abs_tau = st.max_tau; stats->max_com_err = st.max_err; stats->max_abs_qpp = st.max_qpp;
        stats->qp_iters = st.qp_iters; stats->qp_fail = st.qp_fail; stats->steps = st.steps; stats->alive = st.alive; stats->qp_fallback = st.qp_fallback; stats->pad = 0;
    }
}
}
