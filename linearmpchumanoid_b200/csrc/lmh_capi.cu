// lmh_capi.cu -- kernels and C ABI of liblmh_b200.so (see include/lmh_capi.h).
//
// Execution model: one thread block per environment, NT = 32 / 64 / 128 threads cooperating on the
// environment's workspace in shared memory (lmh_model.cuh).  A launch advances every environment by
// n_steps RK4 steps with the state, the stage intermediates and the QP working set resident on chip;
// HBM is touched once per launch for the 60+30+1 state doubles (and once per step for the optional
// log record).  Environments never communicate.
#include "lmh_setup.hpp"
#include "lmh_ik.cuh"
#include "lmh_mpc.cuh"
#include "lmh_scenarios.hpp"
#include <cuda_runtime.h>
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <string>

using namespace lmh;

namespace {

thread_local std::string g_err;
int32_t fail(int32_t code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? LMH_ENODEVICE : LMH_ECUDA, \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                        \
    } while (0)

struct Params {
    const Consts* K;
    Plan plan;
    const int32_t* plan_id;
    double* x;          // [E][60]
    double* vh;         // [E][30]
    double* t;          // [E]
    const int32_t* push_step;
    const double* push_dv;   // [E][6]
    const int32_t* order;    // [E] slot -> environment (or null = identity): environments that are pushed at the same
                             // step share a block, so that a block's warps need their extra QP iterations together
    LmhStats* stats;
    LmhStepLog* log;    // [log_cap][E] or null
    LmhStageDump* dump; // [E] or null
    double* com;        // [E][3]
    int32_t n_env, n_steps, step0, log_cap;
};

constexpr int kExtraDoubles = 2 * 60;   // x and the RK4 accumulator
constexpr int kTabDh = 28 * 36, kTabMpc = kTabDh + 120;     // block-shared constant tables: Isp | DH parameters | mpc_r
constexpr int kMaxDynSmem = 227 * 1024;
constexpr int kMaxEpb = 10;              // one-warp environments per block of rollout_kernel_multi (its launch bound)
static_assert(offsetof(Consts, dh_d) == offsetof(Consts, th_off) + 4 * 24 * sizeof(double), "th_off..dh_d must be contiguous");

template <int NT>
__global__ void __launch_bounds__(NT, NT == 32 ? 10 : 1) rollout_kernel(Params p) {
    extern __shared__ double smem[];
    EnvWS& w = *reinterpret_cast<EnvWS*>(smem);
    double* x = smem + (sizeof(EnvWS) + 7) / 8;
    double* kacc = x + 60;
    const int env = blockIdx.x;
    if (env >= p.n_env) return;
    Ctx<NT> cx{(int)threadIdx.x, 0};
    const Consts& K = *p.K;
    ws_reset(cx, w);
    for (int i = threadIdx.x; i < 60; i += NT) x[i] = p.x[(size_t)env * 60 + i];
    for (int i = threadIdx.x; i < 30; i += NT) w.vh[i] = p.vh[(size_t)env * 30 + i];
    double& t = w.t_env;                                    // clock and statistics live in the workspace (shared memory):
    StatsAcc& st = w.stats;                                 // nothing of the bookkeeping is held in registers across a stage
    const int plan_id = p.plan_id ? p.plan_id[env] : 0;
    const int push_at = p.push_step ? p.push_step[env] : -1;
    if (cx.tid == 0) {
        t = p.t[env];
        const LmhStats& s0 = p.stats[env];
        st.max_tau = s0.max_abs_tau; st.max_err = s0.max_com_err; st.max_qpp = s0.max_abs_qpp;
        st.qp_iters = s0.qp_iters; st.qp_fail = s0.qp_fail; st.steps = s0.steps; st.alive = s0.alive; st.qp_fallback = s0.qp_fallback;
        w.dead = s0.alive ? 0 : 1;
    }
    cx.sync();
    for (int s = 0; s < p.n_steps; ++s) {
        const int gstep = p.step0 + s;
        if (gstep == push_at && w.dead == 0) {
            for (int k = threadIdx.x; k < 6; k += NT) x[30 + k] += p.push_dv[(size_t)env * 6 + k];
            cx.sync();
        }
        LmhStepLog* lg = (p.log && gstep < p.log_cap) ? p.log + (size_t)gstep * p.n_env + env : nullptr;
        rk4_step(cx, w, x, kacc, t, K, p.plan, plan_id, lg, &st);
    }
    for (int i = threadIdx.x; i < 60; i += NT) p.x[(size_t)env * 60 + i] = x[i];
    for (int i = threadIdx.x; i < 30; i += NT) p.vh[(size_t)env * 30 + i] = w.vh[i];
    for (int i = threadIdx.x; i < 3; i += NT) p.com[(size_t)env * 3 + i] = w.com[i];
    if (threadIdx.x == 0) {
        p.t[env] = t;
        LmhStats& s1 = p.stats[env];
        s1.max_abs_tau = st.max_tau; s1.max_com_err = st.max_err; s1.max_abs_qpp = st.max_qpp;
        s1.qp_iters = st.qp_iters; s1.qp_fail = st.qp_fail; s1.steps = st.steps; s1.alive = st.alive; s1.qp_fallback = st.qp_fallback; s1.pad = 0;
    }
}

// One warp per environment, EPB environments per block (warp w of block b owns environment b*EPB + w and its own
// slice of shared memory); the warps of a block are re-aligned at every stage so that they share instruction fetches.
__global__ void __launch_bounds__(32 * kMaxEpb, 1) rollout_kernel_multi(Params p, int epb, int env_bytes, int tab_doubles, int n_full) {
    extern __shared__ double smem[];
    // the warp index through a broadcast: the compiler then knows it is warp-uniform and keeps the workspace base in
    // a uniform register
    const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    char* base = reinterpret_cast<char*>(smem) + (size_t)wid * env_bytes;
    EnvWS& w = *reinterpret_cast<EnvWS*>(base);
    double* x = reinterpret_cast<double*>(base) + (sizeof(EnvWS) + 7) / 8;
    double* kacc = x + 60;
    // Blocks 0 .. n_full-1 carry epb environments each (whole waves of one block per SM); what is left is dealt out
    // evenly over the remaining blocks, so that the last wave runs everywhere with fewer warps per SM (a wave of 8 is
    // 15 % shorter than a wave of 10) instead of full on some SMs and not at all on the others.
    int slot, count;
    if ((int)blockIdx.x < n_full) { slot = blockIdx.x * epb + wid; count = epb; }
    else {
        const int tb = blockIdx.x - n_full, nt = gridDim.x - n_full, rest = p.n_env - n_full * epb;
        const int base_n = rest / nt, rem = rest % nt;
        count = base_n + (tb < rem ? 1 : 0);
        slot = n_full * epb + tb * base_n + (tb < rem ? tb : rem) + wid;
    }
    const bool live = wid < count;
    if (!live) slot = p.n_env - 1;
    const int env = p.order ? p.order[slot] : slot;
    Ctx<32, true> cx{lane, 32 * count == (int)blockDim.x ? -1 : 32 * count};
    const Consts& K = *p.K;
    {
        // behind the environments' slices: one copy per block of the tables every level of the tree recursions reads
        // (the 29 KB of L1 left beside 227 KB of shared memory do not keep them: they came from L2 half of the time)
        double* tab = reinterpret_cast<double*>(reinterpret_cast<char*>(smem) + (size_t)epb * env_bytes);
        for (int i = threadIdx.x; i < 28 * 36; i += blockDim.x) tab[i] = K.Isp[i];
        for (int i = threadIdx.x; i < 120; i += blockDim.x) tab[kTabDh + i] = K.th_off[i];      // th_off, dh_ca, dh_sa, dh_r, dh_d
        for (int i = threadIdx.x; i < tab_doubles - kTabMpc; i += blockDim.x) tab[kTabMpc + i] = K.mpc_r[i];
        __syncthreads();
        cx.isp = tab; cx.dh = tab + kTabDh;
        if (tab_doubles > kTabMpc) cx.mpc_r = tab + kTabMpc;
    }
    if (!live) return;                                      // the alignment barrier counts the live warps only
    ws_reset(cx, w);
    for (int i = lane; i < 60; i += 32) x[i] = p.x[(size_t)env * 60 + i];
    for (int i = lane; i < 30; i += 32) w.vh[i] = p.vh[(size_t)env * 30 + i];
    double& t = w.t_env;                                    // clock and statistics live in the workspace (shared memory):
    StatsAcc& st = w.stats;                                 // nothing of the bookkeeping is held in registers across a stage
    const int plan_id = p.plan_id ? p.plan_id[env] : 0;
    const int push_at = p.push_step ? p.push_step[env] : -1;
    if (cx.tid == 0) {
        t = p.t[env];
        const LmhStats& s0 = p.stats[env];
        st.max_tau = s0.max_abs_tau; st.max_err = s0.max_com_err; st.max_qpp = s0.max_abs_qpp;
        st.qp_iters = s0.qp_iters; st.qp_fail = s0.qp_fail; st.steps = s0.steps; st.alive = s0.alive; st.qp_fallback = s0.qp_fallback;
        w.dead = s0.alive ? 0 : 1;
    }
    cx.sync();
    for (int s = 0; s < p.n_steps; ++s) {
        const int gstep = p.step0 + s;
        if (gstep == push_at && w.dead == 0) {
            for (int k = lane; k < 6; k += 32) x[30 + k] += p.push_dv[(size_t)env * 6 + k];
            cx.sync();
        }
        LmhStepLog* lg = (p.log && gstep < p.log_cap) ? p.log + (size_t)gstep * p.n_env + env : nullptr;
        rk4_step(cx, w, x, kacc, t, K, p.plan, plan_id, lg, &st);
    }
    for (int i = lane; i < 60; i += 32) p.x[(size_t)env * 60 + i] = x[i];
    for (int i = lane; i < 30; i += 32) p.vh[(size_t)env * 30 + i] = w.vh[i];
    for (int i = lane; i < 3; i += 32) p.com[(size_t)env * 3 + i] = w.com[i];
    if (lane == 0) {
        p.t[env] = t;
        LmhStats& s1 = p.stats[env];
        s1.max_abs_tau = st.max_tau; s1.max_com_err = st.max_err; s1.max_abs_qpp = st.max_qpp;
        s1.qp_iters = st.qp_iters; s1.qp_fail = st.qp_fail; s1.steps = st.steps; s1.alive = st.alive; s1.qp_fallback = st.qp_fallback; s1.pad = 0;
    }
}

// Latency mode (LmhConfig.envs_per_block = -1, meant for a handful of environments): one environment per block of three
// warps.  Warp 0 is the environment's warp and runs the whole step; while it runs the RNEA, warp 1 computes the CRBA and
// warp 2 the sole Jacobians and then what of the MPC / PD phase needs no dynamics: ZMP window sums, joint PD, foot references
// (the three only read the transforms; their outputs and scratch are disjoint): 46 us instead of 62 us per RK4 step of a
// single environment.  Every phase is the same source on the same data, but this is another
// kernel: the compiler contracts a few sums of products differently, so results agree with the throughput kernel to
// rounding (1e-12 after 50 pushed steps), not bit for bit -- which is why it is opt-in and the default path is one kernel
// for every batch size (sharding invariance is bit exact).
__global__ void __launch_bounds__(96, 1) rollout_kernel_helpers(Params p, int env_bytes, int tab_doubles) {
    extern __shared__ double smem[];
    EnvWS& w = *reinterpret_cast<EnvWS*>(smem);
    double* x = smem + (sizeof(EnvWS) + 7) / 8;
    double* kacc = x + 60;
    const int env = blockIdx.x;
    const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (env >= p.n_env) return;
    const Consts& K = *p.K;
    // the same block-shared tables as the throughput kernel, so that both run the same instantiation of every phase
    // (Ctx<32, true>): identical arithmetic, bit for bit
    double* tab = reinterpret_cast<double*>(reinterpret_cast<char*>(smem) + env_bytes);
    for (int i = threadIdx.x; i < 28 * 36; i += blockDim.x) tab[i] = K.Isp[i];
    for (int i = threadIdx.x; i < 120; i += blockDim.x) tab[kTabDh + i] = K.th_off[i];
    for (int i = threadIdx.x; i < tab_doubles - kTabMpc; i += blockDim.x) tab[kTabMpc + i] = K.mpc_r[i];
    __syncthreads();
    const int plan_id = p.plan_id ? p.plan_id[env] : 0;
    Ctx<32, true> cx{lane, 0};
    cx.isp = tab; cx.dh = tab + kTabDh;
    if (tab_doubles > kTabMpc) cx.mpc_r = tab + kTabMpc;
    if (wid != 0) {
        const LaneRole L = lane_role(lane);
        for (int s = 0; s < p.n_steps; ++s) {
            asm volatile("bar.sync 1, 96;" ::: "memory");
            if (w.dead != 0) continue;
            for (int stage = 0; stage < 4; ++stage) {
                asm volatile("bar.sync 2, 96;" ::: "memory");
                if (wid == 1) crba_warp(w, K, tab, L, lane);
                else {
                    // the sole Jacobians, then everything of the MPC / PD phase that does not need the dynamics
                    jacobian_warp(w, lane);
                    __syncwarp();
                    mpc_pd_helper(cx, w, K, p.plan, plan_id);
                }
                asm volatile("bar.sync 3, 96;" ::: "memory");
            }
        }
        return;
    }
    ws_reset(cx, w);
    for (int i = lane; i < 60; i += 32) x[i] = p.x[(size_t)env * 60 + i];
    for (int i = lane; i < 30; i += 32) w.vh[i] = p.vh[(size_t)env * 30 + i];
    double& t = w.t_env;                                    // clock and statistics live in the workspace (shared memory):
    StatsAcc& st = w.stats;                                 // nothing of the bookkeeping is held in registers across a stage
    const int push_at = p.push_step ? p.push_step[env] : -1;
    if (cx.tid == 0) {
        t = p.t[env];
        const LmhStats& s0 = p.stats[env];
        st.max_tau = s0.max_abs_tau; st.max_err = s0.max_com_err; st.max_qpp = s0.max_abs_qpp;
        st.qp_iters = s0.qp_iters; st.qp_fail = s0.qp_fail; st.steps = s0.steps; st.alive = s0.alive; st.qp_fallback = s0.qp_fallback;
        w.dead = s0.alive ? 0 : 1;
    }
    cx.sync();
    cx.helpers = 96;
    for (int s = 0; s < p.n_steps; ++s) {
        const int gstep = p.step0 + s;
        if (gstep == push_at && w.dead == 0) {
            for (int k = lane; k < 6; k += 32) x[30 + k] += p.push_dv[(size_t)env * 6 + k];
            cx.sync();
        }
        LmhStepLog* lg = (p.log && gstep < p.log_cap) ? p.log + (size_t)gstep * p.n_env + env : nullptr;
        rk4_step(cx, w, x, kacc, t, K, p.plan, plan_id, lg, &st);
    }
    for (int i = lane; i < 60; i += 32) p.x[(size_t)env * 60 + i] = x[i];
    for (int i = lane; i < 30; i += 32) p.vh[(size_t)env * 30 + i] = w.vh[i];
    for (int i = lane; i < 3; i += 32) p.com[(size_t)env * 3 + i] = w.com[i];
    if (lane == 0) {
        p.t[env] = t;
        LmhStats& s1 = p.stats[env];
        s1.max_abs_tau = st.max_tau; s1.max_com_err = st.max_err; s1.max_abs_qpp = st.max_qpp;
        s1.qp_iters = st.qp_iters; s1.qp_fail = st.qp_fail; s1.steps = st.steps; s1.alive = st.alive; s1.qp_fallback = st.qp_fallback; s1.pad = 0;
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) stage_kernel(Params p) {
    extern __shared__ double smem[];
    EnvWS& w = *reinterpret_cast<EnvWS*>(smem);
    const int env = blockIdx.x;
    if (env >= p.n_env) return;
    Ctx<NT> cx{(int)threadIdx.x, 0};
    for (int i = threadIdx.x; i < 30; i += NT) {
        w.q[i] = p.x[(size_t)env * 60 + i];
        w.v[i] = p.x[(size_t)env * 60 + 30 + i];
        w.vh[i] = p.vh[(size_t)env * 30 + i];
    }
    if (threadIdx.x == 0) w.t = p.t[env];
    ws_reset(cx, w);
    stage_eval(cx, w, *p.K, p.plan, p.plan_id ? p.plan_id[env] : 0, p.dump ? p.dump + env : nullptr);
    for (int i = threadIdx.x; i < 30; i += NT) p.vh[(size_t)env * 30 + i] = w.vh[i];
    for (int i = threadIdx.x; i < 3; i += NT) p.com[(size_t)env * 3 + i] = w.com[i];
}

// Config 5 micro-benchmark: the model phases once at the frozen state, then `repeats` cold-started solves of the
// WBC QP (mode 0; H/g data, factorisation, active set, back-substitution: everything controller.cpp:388-479 +
// qpOASES do per call) or `repeats` evaluations of the LIPM-MPC (mode 1).  out: [E][42] = a (30), f (12).
template <int NT>
__global__ void __launch_bounds__(NT) qp_bench_kernel(Params p, int repeats, int mode, double* out) {
    extern __shared__ double smem[];
    EnvWS& w = *reinterpret_cast<EnvWS*>(smem);
    const int env = blockIdx.x;
    if (env >= p.n_env) return;
    Ctx<NT> cx{(int)threadIdx.x, 0};
    const Consts& K = *p.K;
    for (int i = threadIdx.x; i < 30; i += NT) {
        w.q[i] = p.x[(size_t)env * 60 + i];
        w.v[i] = p.x[(size_t)env * 60 + 30 + i];
        w.vh[i] = p.vh[(size_t)env * 30 + i];
    }
    if (threadIdx.x == 0) w.t = p.t[env];
    ws_reset(cx, w);
    const int plan_id = p.plan_id ? p.plan_id[env] : 0;
    model_phases(cx, w, K);
    mpc_pd_phase(cx, w, K, p.plan, plan_id);
    for (int r = 0; r < repeats; ++r) {
        if (mode == 1) { mpc_pd_phase(cx, w, K, p.plan, plan_id); continue; }
        ws_reset(cx, w);                  // cold start, like the reference's SQProblem::init every call
        qp_stage(cx, w, K, nullptr);
    }
    for (int i = threadIdx.x; i < 42; i += NT) out[(size_t)env * 42 + i] = i < 30 ? w.a[i] : w.f12[i - 30];
}

// Batched inverse kinematics of the initial posture (lmh_ik.cuh): one warp per target, four targets per block.
struct IkParams {
    const Consts* K;
    const double* q_start;     // [30] posture the Newton iteration starts from (and source of the arm/head targets)
    const double* q_start_each;// [n][30] or null: a starting posture per target
    const double* desop;       // [n][30] or null: the full operational target (then com / rf / lf are unused)
    const double* com;         // [n][3]
    const double* rf;          // [n][6]
    const double* lf;          // [n][6]
    double* q_out;             // [n][30]
    int32_t* iters;            // [n]
    double* residual;          // [n]
    int32_t* status;           // [n]
    double* x;                 // state to initialise ([n][60], with vh and t) or null
    double* vh;
    double* t;
    int32_t n, max_iter, env_bytes;
};

__global__ void __launch_bounds__(128) ik_kernel(IkParams p) {
    extern __shared__ double smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * 4 + wid;
    if (env >= p.n) return;
    EnvWS& w = *reinterpret_cast<EnvWS*>(reinterpret_cast<char*>(smem) + (size_t)wid * p.env_bytes);
    double* des = w.qppRef;
    if (lane < NQ) {
        w.q[lane] = p.q_start_each ? p.q_start_each[(size_t)env * NQ + lane] : p.q_start[lane];
        // Kinematics::desiredOperationalState (invKinematics.cpp:11-25): feet poses, the current arm/head joints,
        // zero attitude, CoM
        if (p.desop) des[lane] = p.desop[(size_t)env * NQ + lane];
        else des[lane] = lane < 6 ? p.rf[(size_t)env * 6 + lane] : lane < 12 ? p.lf[(size_t)env * 6 + lane - 6]
                       : lane < 24 ? p.q_start[18 + lane - 12] : lane < 27 ? 0.0 : p.com[(size_t)env * 3 + lane - 27];
    }
    __syncwarp();
    const IkResult r = ik_solve_warp(w, *p.K, des, p.max_iter, lane);
    if (lane < NQ) {
        p.q_out[(size_t)env * NQ + lane] = w.q[lane];
        if (p.x) {
            p.x[(size_t)env * 60 + lane] = w.q[lane];
            p.x[(size_t)env * 60 + NQ + lane] = 0.0;
            p.vh[(size_t)env * NQ + lane] = 0.0;
        }
    }
    if (lane == 0) {
        p.iters[env] = r.iters; p.residual[env] = r.residual; p.status[env] = r.status;
        if (p.x) p.t[env] = 0.0;
    }
}

// Robot::updateState for a batch of postures: one warp per posture, four per block; dumps the world transforms and the
// parent-relative transforms of the 27 frames as (R row-major, p) records plus the centre of mass.
__global__ void __launch_bounds__(128) kin_kernel(const Consts* K, const double* q, double* Tw_out, double* Xl_out, double* com_out,
                                                  int n, int env_bytes) {
    extern __shared__ double smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int env = blockIdx.x * 4 + wid;
    if (env >= n) return;
    EnvWS& w = *reinterpret_cast<EnvWS*>(reinterpret_cast<char*>(smem) + (size_t)wid * env_bytes);
    if (lane < NQ) w.q[lane] = q[(size_t)env * NQ + lane];
    __syncwarp();
    Ctx<32> cx{lane, 0};
    fk_phase(cx, w, *K);
    for (int i = lane; i < NF * 12; i += 32) {
        if (Tw_out) Tw_out[(size_t)env * NF * 12 + i] = w.u.m.p.Tw[i];
        if (Xl_out) Xl_out[(size_t)env * NF * 12 + i] = w.u.m.Xl[i];
    }
    if (com_out && lane < 3) com_out[(size_t)env * 3 + lane] = w.com[lane];
}

__global__ void fp64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace

// Device copies of the QP working-set table (lmh_setup.hpp: 65536 x 72 doubles = 37.7 MB): one per (device, host table),
// reference counted, shared by every handle that uses it.
namespace {
struct QpTabDev { int device; const void* key; double* d_tab; int refs; };
std::mutex g_qp_tab_mu;
std::vector<QpTabDev> g_qp_tabs;
int32_t qp_table_acquire(int device, const std::shared_ptr<const std::vector<double>>& host, const double** d_out) {
    std::lock_guard<std::mutex> lock(g_qp_tab_mu);
    for (QpTabDev& t : g_qp_tabs)
        if (t.device == device && t.key == host.get()) { ++t.refs; *d_out = t.d_tab; return LMH_OK; }
    double* d = nullptr;
    CU(cudaMalloc(&d, host->size() * sizeof(double)));
    cudaError_t e = cudaMemcpy(d, host->data(), host->size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return fail(LMH_ECUDA, std::string("working-set table upload: ") + cudaGetErrorString(e)); }
    g_qp_tabs.push_back(QpTabDev{device, host.get(), d, 1});
    *d_out = d;
    return LMH_OK;
}
void qp_table_release(int device, const void* key) {
    std::lock_guard<std::mutex> lock(g_qp_tab_mu);
    for (size_t i = 0; i < g_qp_tabs.size(); ++i)
        if (g_qp_tabs[i].device == device && g_qp_tabs[i].key == key) {
            if (--g_qp_tabs[i].refs == 0) { cudaFree(g_qp_tabs[i].d_tab); g_qp_tabs.erase(g_qp_tabs.begin() + i); }
            return;
        }
}
}  // namespace

struct LmhBatch {
    LmhConfig cfg;
    HostModel* hm = nullptr;
    int32_t n_env = 0, nt = 64, smem = 0, epb = 10, tab_doubles = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    Consts* dK = nullptr;
    const void* qp_tab_key = nullptr;      // which shared device table this handle holds a reference on
    double *d_x = nullptr, *d_vh = nullptr, *d_t = nullptr, *d_push_dv = nullptr, *d_com = nullptr;
    double *d_zmp_x = nullptr, *d_zmp_y = nullptr, *d_foot = nullptr;
    int32_t *d_plan_id = nullptr, *d_push_step = nullptr, *d_order = nullptr;
    LmhStats* d_stats = nullptr;
    LmhStats* d_stats0 = nullptr;
    std::vector<LmhStats> h_stats0;
    LmhStepLog* d_log = nullptr;
    LmhStageDump* d_dump = nullptr;
    int32_t n_plans = 1, zmp_len = 0, step_counter = 0, launches = 0, n_phases = 1;
    double phase_T = 1e30;
    int32_t* d_contact = nullptr;
    std::vector<double> h_zmp_x, h_zmp_y, h_foot;
    std::vector<int32_t> h_contact;
    bool has_pushes = false, has_plan_ids = false, has_order = false;
    bool small_batch = false;           // latency mode: one block of three warps per environment
    double* d_mpc = nullptr;            // dense MPC tables (PuT | L | LT | Dinv | DinvT | Px | rdiag), uploaded on first use
    // scratch of lmh_ik_batch / lmh_qp_bench / lmh_kinematics / lmh_mpc_solve, kept between calls (grown on demand)
    double* d_scratch = nullptr;
    size_t scratch_bytes = 0;
};

namespace {

Params make_params(LmhBatch* h) {
    Params p;
    p.K = h->dK;
    p.plan.zmp_x = h->d_zmp_x; p.plan.zmp_y = h->d_zmp_y; p.plan.foot_coef = h->d_foot; p.plan.zmp_len = h->zmp_len;
    p.plan.contact = h->d_contact; p.plan.n_phases = h->n_phases; p.plan.phase_T = h->phase_T;
    p.plan_id = h->has_plan_ids ? h->d_plan_id : nullptr;
    p.x = h->d_x; p.vh = h->d_vh; p.t = h->d_t;
    p.push_step = h->has_pushes ? h->d_push_step : nullptr;
    p.push_dv = h->d_push_dv;
    p.order = (h->has_pushes && h->has_order) ? h->d_order : nullptr;
    p.stats = h->d_stats; p.log = h->d_log; p.dump = nullptr; p.com = h->d_com;
    p.n_env = h->n_env; p.n_steps = 0; p.step0 = h->step_counter; p.log_cap = h->cfg.log_steps;
    return p;
}

template <int NT>
int32_t configure_kernels(int smem) {
    CU(cudaFuncSetAttribute(rollout_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(stage_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return LMH_OK;
}

int32_t upload_plans(LmhBatch* h) {
    cudaFree(h->d_zmp_x); cudaFree(h->d_zmp_y); cudaFree(h->d_foot); cudaFree(h->d_contact);
    h->d_zmp_x = h->d_zmp_y = h->d_foot = nullptr; h->d_contact = nullptr;
    size_t nz = (size_t)h->n_plans * h->zmp_len * sizeof(double), nf = (size_t)h->n_plans * h->n_phases * 40 * sizeof(double);
    size_t nc = (size_t)h->n_plans * h->n_phases * sizeof(int32_t);
    CU(cudaMalloc(&h->d_zmp_x, nz)); CU(cudaMalloc(&h->d_zmp_y, nz)); CU(cudaMalloc(&h->d_foot, nf)); CU(cudaMalloc(&h->d_contact, nc));
    CU(cudaMemcpyAsync(h->d_contact, h->h_contact.data(), nc, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_zmp_x, h->h_zmp_x.data(), nz, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_zmp_y, h->h_zmp_y.data(), nz, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_foot, h->h_foot.data(), nf, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LMH_OK;
}

int32_t ensure_scratch(LmhBatch* h, size_t bytes) {
    if (bytes <= h->scratch_bytes) return LMH_OK;
    if (h->d_scratch) { cudaFree(h->d_scratch); h->d_scratch = nullptr; h->scratch_bytes = 0; }
    CU(cudaMalloc(&h->d_scratch, bytes));
    h->scratch_bytes = bytes;
    return LMH_OK;
}

// Grid of the one-block-per-SM kernel: whole waves of full blocks, then the rest spread evenly over up to sm_count blocks.
int multi_grid(const LmhBatch* h, int* n_full_out) {
    const int cap = h->sm_count * h->epb;
    int n_full = ((h->n_env + cap - 1) / cap - 1) * h->sm_count;
    const int rest = h->n_env - n_full * h->epb;
    const int tail = std::min(rest, h->sm_count);
    if (n_full_out) *n_full_out = n_full;
    return n_full + tail;
}

int32_t launch_rollout(LmhBatch* h, int32_t n_steps) {
    Params p = make_params(h);
    p.n_steps = n_steps;
    dim3 grid(h->n_env), block(h->nt);
    if (h->nt == 32 && h->small_batch) {
        rollout_kernel_helpers<<<grid, 96, (size_t)h->smem + (size_t)h->tab_doubles * sizeof(double), h->stream>>>(p, h->smem, h->tab_doubles);
    } else
    if (h->nt == 32) {
        int n_full = 0;
        const int nblk = multi_grid(h, &n_full);
        rollout_kernel_multi<<<nblk, 32 * h->epb, (size_t)h->smem * h->epb + (size_t)h->tab_doubles * sizeof(double), h->stream>>>(p, h->epb, h->smem, h->tab_doubles, n_full);
    } else
    if (h->nt == 64) rollout_kernel<64><<<grid, block, h->smem, h->stream>>>(p);
    else rollout_kernel<128><<<grid, block, h->smem, h->stream>>>(p);
    CU(cudaGetLastError());
    h->step_counter += n_steps;
    h->launches += 1;
    return LMH_OK;
}

}  // namespace

extern "C" {

const char* lmh_last_error(void) { return g_err.c_str(); }

void lmh_default_config(LmhConfig* c) {
    Scenario s;
    std::memset(c, 0, sizeof(*c));
    c->dt = s.dt; c->sim_time = s.sim_time; c->time_horizon = s.time_horizon;
    for (int k = 0; k < 3; ++k) c->com[k] = s.com[k];
    for (int k = 0; k < 6; ++k) { c->rf[k] = s.rf[k]; c->lf[k] = s.lf[k]; }
    c->mu = s.mu; c->mpc_alpha = s.mpc_alpha; c->mpc_beta = s.mpc_beta;
    c->qp_max_iter = s.qp_max_iter;
    c->threads_per_env = 0; c->device = 0; c->log_steps = 0; c->envs_per_block = 0; c->placement = 0; c->mpc_zcom = 0.0;
}

// everything of lmh_create that can fail after the handle exists: the caller destroys the handle on any error
static int32_t create_on_device(LmhBatch* h, const LmhConfig* cfg, int32_t n_env) {
    h->smem = (int32_t)(((sizeof(EnvWS) + 7) / 8 + kExtraDoubles) * sizeof(double));
    h->smem = (h->smem + 15) / 16 * 16;
    CU(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));
    const int epb_max = std::min(kMaxEpb, (kMaxDynSmem - kTabMpc * 8) / h->smem);
    h->epb = cfg->envs_per_block > 0 ? cfg->envs_per_block : epb_max;
    if (h->nt != 32) h->epb = 1;
    // latency mode (envs_per_block = -1): one environment per block, two helper warps
    h->small_batch = h->nt == 32 && cfg->envs_per_block == -1;
    if (h->small_batch) CU(cudaFuncSetAttribute(rollout_kernel_helpers, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    if (h->epb > epb_max) return fail(LMH_EINVAL, "lmh_create: envs_per_block exceeds what the shared memory holds");
    if (h->nt == 32) {
        // block-shared constant tables behind the slices, as far as the 227 KB allow: Isp + DH (9 KB), then the MPC gains
        const int64_t slices = (int64_t)h->smem * h->epb;
        const int n1 = h->hm->K.mpc_n1;
        h->tab_doubles = slices + (int64_t)(kTabMpc + n1) * 8 <= kMaxDynSmem ? kTabMpc + n1 : kTabMpc;
        // the attribute belongs to the (process-wide) kernel, not to this handle: always ask for the maximum, so that
        // handles with different horizons (= different table sizes) can be live together
        CU(cudaFuncSetAttribute(rollout_kernel_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    }
    int32_t rc = h->nt == 32 ? configure_kernels<32>(h->smem) : h->nt == 64 ? configure_kernels<64>(h->smem) : configure_kernels<128>(h->smem);
    if (rc != LMH_OK) return rc;
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&h->ev0)); CU(cudaEventCreate(&h->ev1));
    size_t E = (size_t)n_env;
    CU(cudaMalloc(&h->dK, sizeof(Consts)));
    {
        // the device copy of the constants points at the device copy of the working-set table (shared by every handle
        // of this device that uses the same generators)
        const double* d_tab = nullptr;
        int32_t rt = qp_table_acquire(h->cfg.device, h->hm->qp_tab, &d_tab);
        if (rt != LMH_OK) return rt;
        h->qp_tab_key = h->hm->qp_tab.get();
        Consts kc = h->hm->K;
        kc.qp_tab = d_tab;
        CU(cudaMemcpy(h->dK, &kc, sizeof(Consts), cudaMemcpyHostToDevice));
    }
    CU(cudaMalloc(&h->d_x, E * 60 * sizeof(double))); CU(cudaMalloc(&h->d_vh, E * 30 * sizeof(double)));
    CU(cudaMalloc(&h->d_t, E * sizeof(double))); CU(cudaMalloc(&h->d_push_dv, E * 6 * sizeof(double)));
    CU(cudaMalloc(&h->d_com, E * 3 * sizeof(double)));
    CU(cudaMalloc(&h->d_plan_id, E * sizeof(int32_t))); CU(cudaMalloc(&h->d_push_step, E * sizeof(int32_t)));
    CU(cudaMalloc(&h->d_stats, E * sizeof(LmhStats)));
    CU(cudaMalloc(&h->d_dump, E * sizeof(LmhStageDump)));
    if (cfg->log_steps > 0) {
        CU(cudaMalloc(&h->d_log, (size_t)cfg->log_steps * E * sizeof(LmhStepLog)));
        CU(cudaMemset(h->d_log, 0, (size_t)cfg->log_steps * E * sizeof(LmhStepLog)));
    }
    h->n_plans = 1;
    h->zmp_len = (int32_t)h->hm->zmp_x.size();
    h->h_zmp_x = h->hm->zmp_x; h->h_zmp_y = h->hm->zmp_y; h->h_foot = h->hm->foot_coef;
    h->h_contact.assign(1, 3);
    rc = upload_plans(h);
    if (rc != LMH_OK) return rc;
    return lmh_reset(h);
}

int32_t lmh_create(const LmhConfig* cfg, int32_t n_env, LmhBatch** out) {
    if (!cfg || !out || n_env <= 0) return fail(LMH_EINVAL, "lmh_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(LMH_ENODEVICE, std::string("lmh_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(LMH_EINVAL, "lmh_create: device ordinal out of range");
    const int nt = cfg->threads_per_env == 0 ? 32 : cfg->threads_per_env;
    if (nt != 32 && nt != 64 && nt != 128) return fail(LMH_EINVAL, "lmh_create: threads_per_env must be 32, 64 or 128");
    if (cfg->envs_per_block < -1 || cfg->placement < 0 || cfg->placement > 1 || cfg->log_steps < 0)
        return fail(LMH_EINVAL, "lmh_create: envs_per_block / placement / log_steps out of range");
    CU(cudaSetDevice(cfg->device));
    LmhBatch* h = new LmhBatch;
    h->cfg = *cfg;
    h->n_env = n_env;
    h->nt = nt;
    try {
        Scenario s;
        s.dt = cfg->dt; s.sim_time = cfg->sim_time; s.time_horizon = cfg->time_horizon;
        for (int k = 0; k < 3; ++k) s.com[k] = cfg->com[k];
        for (int k = 0; k < 6; ++k) { s.rf[k] = cfg->rf[k]; s.lf[k] = cfg->lf[k]; }
        s.mu = cfg->mu; s.mpc_alpha = cfg->mpc_alpha; s.mpc_beta = cfg->mpc_beta;
        s.qp_max_iter = cfg->qp_max_iter > 0 ? cfg->qp_max_iter : 200;
        s.mpc_zcom = cfg->mpc_zcom;
        h->hm = new HostModel(s);
    } catch (const std::exception& ex) {
        delete h;
        return fail(LMH_EINVAL, std::string("lmh_create: setup failed: ") + ex.what());
    }
    const int32_t rc = create_on_device(h, cfg, n_env);
    if (rc != LMH_OK) {
        const std::string msg = g_err;      // lmh_destroy must not lose the message
        lmh_destroy(h);
        g_err = msg;
        return rc;
    }
    *out = h;
    return LMH_OK;
}

void lmh_destroy(LmhBatch* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->qp_tab_key) qp_table_release(h->cfg.device, h->qp_tab_key);
    cudaFree(h->dK); cudaFree(h->d_x); cudaFree(h->d_vh); cudaFree(h->d_t); cudaFree(h->d_push_dv); cudaFree(h->d_com);
    cudaFree(h->d_plan_id); cudaFree(h->d_push_step); cudaFree(h->d_order); cudaFree(h->d_stats); cudaFree(h->d_stats0); cudaFree(h->d_dump); cudaFree(h->d_log);
    cudaFree(h->d_zmp_x); cudaFree(h->d_zmp_y); cudaFree(h->d_foot); cudaFree(h->d_contact); cudaFree(h->d_scratch); cudaFree(h->d_mpc);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h->hm;
    delete h;
}

int32_t lmh_n_env(const LmhBatch* h) { return h ? h->n_env : 0; }

int32_t lmh_get_nominal(const LmhBatch* h, double* q0, double* zcom, double* mass, int32_t* ik) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    if (q0) for (int i = 0; i < NQ; ++i) q0[i] = h->hm->q0[i];
    if (zcom) *zcom = h->hm->K.mpc_zcom;
    if (mass) *mass = h->hm->K.mass;
    if (ik) *ik = h->hm->ik_iterations;
    return LMH_OK;
}

int32_t lmh_get_mpc_gain(const LmhBatch* h, double* r, int32_t cap, int32_t* n1, double* rPx2) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    const Consts& K = h->hm->K;
    if (n1) *n1 = K.mpc_n1;
    if (r) for (int i = 0; i < K.mpc_n1 && i < cap; ++i) r[i] = K.mpc_r[i];
    if (rPx2) { rPx2[0] = K.mpc_rPx[0]; rPx2[1] = K.mpc_rPx[1]; }
    return LMH_OK;
}

int32_t lmh_get_link_tables(const LmhBatch* h, double* mass, double* com, double* isp) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    const Consts& K = h->hm->K;
    for (int i = 0; i < 28; ++i) {
        if (mass) mass[i] = K.mc[4 * i + 3];
        if (com) for (int k = 0; k < 3; ++k) com[3 * i + k] = K.mc[4 * i + k];
    }
    if (isp) std::memcpy(isp, K.Isp, sizeof(double) * 28 * 36);
    return LMH_OK;
}

int32_t lmh_get_link_inertia(const LmhBatch* h, double* inertia) {
    if (!h || !inertia) return fail(LMH_EINVAL, "lmh_get_link_inertia: bad arguments");
    std::memcpy(inertia, h->hm->link_inertia, sizeof(double) * 28 * 9);
    return LMH_OK;
}

int32_t lmh_get_nao_parameters(double* mass, double* com, double* inertia) {
    for (int i = 0; i < 28; ++i) {
        if (mass) mass[i] = kNaoLinks[i].m;
        if (com) for (int k = 0; k < 3; ++k) com[3 * i + k] = kNaoLinks[i].c[k];
        if (inertia) for (int k = 0; k < 9; ++k) inertia[9 * i + k] = kNaoLinks[i].I[k];
    }
    return LMH_OK;
}

int32_t lmh_local_transforms(const double* theta25, double* T25x16) {
    if (!theta25 || !T25x16) return fail(LMH_EINVAL, "lmh_local_transforms: bad arguments");
    for (int i = 0; i < 25; ++i) nao_local_transform(i, theta25[i], T25x16 + 16 * i);
    return LMH_OK;
}

int32_t lmh_poly_fit(int32_t n_pos, const double* pos, int32_t n_vel, const double* vel, int32_t n_acc, const double* acc, double* coeff) {
    if (n_pos < 0 || n_vel < 0 || n_acc < 0 || n_pos + n_vel + n_acc <= 0 || !coeff || (n_pos && !pos) || (n_vel && !vel) || (n_acc && !acc))
        return fail(LMH_EINVAL, "lmh_poly_fit: bad arguments");
    std::vector<double> pt, pv, vt, vv, at, av;
    for (int i = 0; i < n_pos; ++i) { pt.push_back(pos[2 * i]); pv.push_back(pos[2 * i + 1]); }
    for (int i = 0; i < n_vel; ++i) { vt.push_back(vel[2 * i]); vv.push_back(vel[2 * i + 1]); }
    for (int i = 0; i < n_acc; ++i) { at.push_back(acc[2 * i]); av.push_back(acc[2 * i + 1]); }
    try {
        const std::vector<double> c = poly_fit(pt, pv, vt, vv, at, av);
        std::memcpy(coeff, c.data(), c.size() * sizeof(double));
    } catch (const std::exception& ex) { return fail(LMH_EINVAL, std::string("lmh_poly_fit: ") + ex.what()); }
    return LMH_OK;
}

int32_t lmh_set_walk_plans(LmhBatch* h, int32_t n_plans, int32_t zmp_len, const double* zx, const double* zy,
                           int32_t n_phases, double phase_T, const double* foot, const int32_t* contact,
                           const int32_t* plan_ids) {
    if (!h || n_plans <= 0 || zmp_len < h->hm->K.mpc_n1 || !zx || !zy || !foot || n_phases <= 0 || !(phase_T > 0) || !contact)
        return fail(LMH_EINVAL, "lmh_set_walk_plans: bad arguments (zmp_len must cover horizon+1 samples)");
    CU(cudaSetDevice(h->cfg.device));
    if (plan_ids)
        for (int e = 0; e < h->n_env; ++e)
            if (plan_ids[e] < 0 || plan_ids[e] >= n_plans) return fail(LMH_EINVAL, "lmh_set_walk_plans: plan id out of range");
    for (int i = 0; i < n_plans * n_phases; ++i)
        if (contact[i] < 0 || contact[i] > 3) return fail(LMH_EINVAL, "lmh_set_walk_plans: contact masks are 0..3");
    CU(cudaStreamSynchronize(h->stream));
    h->n_plans = n_plans; h->zmp_len = zmp_len; h->n_phases = n_phases; h->phase_T = phase_T;
    h->h_zmp_x.assign(zx, zx + (size_t)n_plans * zmp_len);
    h->h_zmp_y.assign(zy, zy + (size_t)n_plans * zmp_len);
    h->h_foot.assign(foot, foot + (size_t)n_plans * n_phases * 40);
    h->h_contact.assign(contact, contact + (size_t)n_plans * n_phases);
    int32_t rc = upload_plans(h);
    if (rc != LMH_OK) return rc;
    h->has_plan_ids = plan_ids != nullptr;
    if (plan_ids) {
        CU(cudaMemcpyAsync(h->d_plan_id, plan_ids, sizeof(int32_t) * h->n_env, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return LMH_OK;
}

int32_t lmh_set_plans(LmhBatch* h, int32_t n_plans, int32_t zmp_len, const double* zx, const double* zy,
                      const double* foot, const int32_t* plan_ids) {
    if (n_plans <= 0) return fail(LMH_EINVAL, "lmh_set_plans: bad arguments");
    std::vector<int32_t> both((size_t)n_plans, 3);
    return lmh_set_walk_plans(h, n_plans, zmp_len, zx, zy, 1, 1e30, foot, both.data(), plan_ids);
}

int32_t lmh_get_plan(const LmhBatch* h, int32_t plan, double* zx, double* zy, int32_t cap, int32_t* zmp_len, double* foot40) {
    if (!h || plan < 0 || plan >= h->n_plans) return fail(LMH_EINVAL, "lmh_get_plan: bad arguments");
    if (zmp_len) *zmp_len = h->zmp_len;
    for (int i = 0; i < h->zmp_len && i < cap; ++i) {
        if (zx) zx[i] = h->h_zmp_x[(size_t)plan * h->zmp_len + i];
        if (zy) zy[i] = h->h_zmp_y[(size_t)plan * h->zmp_len + i];
    }
    if (foot40) std::memcpy(foot40, h->h_foot.data() + (size_t)plan * h->n_phases * 40, 40 * sizeof(double));   // phase 0
    return LMH_OK;
}

int32_t lmh_set_state(LmhBatch* h, const double* x, const double* vh, const double* t) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    size_t E = (size_t)h->n_env;
    if (x) CU(cudaMemcpyAsync(h->d_x, x, E * 60 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (vh) CU(cudaMemcpyAsync(h->d_vh, vh, E * 30 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (t) CU(cudaMemcpyAsync(h->d_t, t, E * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return LMH_OK;
}

int32_t lmh_get_state(LmhBatch* h, double* x, double* vh, double* t) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    size_t E = (size_t)h->n_env;
    if (x) CU(cudaMemcpyAsync(x, h->d_x, E * 60 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (vh) CU(cudaMemcpyAsync(vh, h->d_vh, E * 30 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (t) CU(cudaMemcpyAsync(t, h->d_t, E * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LMH_OK;
}

int32_t lmh_reset(LmhBatch* h) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    size_t E = (size_t)h->n_env;
    std::vector<double> x(E * 60, 0.0);
    for (size_t e = 0; e < E; ++e) for (int i = 0; i < NQ; ++i) x[e * 60 + i] = h->hm->q0[i];
    CU(cudaMemcpyAsync(h->d_x, x.data(), E * 60 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(h->d_vh, 0, E * 30 * sizeof(double), h->stream));
    CU(cudaMemsetAsync(h->d_t, 0, E * sizeof(double), h->stream));
    CU(cudaMemsetAsync(h->d_com, 0, E * 3 * sizeof(double), h->stream));
    std::vector<LmhStats> st(E);
    std::memset(st.data(), 0, E * sizeof(LmhStats));
    for (size_t e = 0; e < E; ++e) st[e].alive = 1;
    CU(cudaMemcpyAsync(h->d_stats, st.data(), E * sizeof(LmhStats), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->step_counter = 0;
    return LMH_OK;
}

int32_t lmh_set_pushes(LmhBatch* h, const int32_t* push_step, const double* push_dv6) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    h->has_pushes = push_step != nullptr && push_dv6 != nullptr;
    h->has_order = false;
    if (h->has_pushes && h->cfg.placement == 0) {
        // placement only (every environment computes exactly what it would anywhere else): slots sorted by push step,
        // larger impulses first within a step
        std::vector<int32_t> order((size_t)h->n_env);
        for (int e = 0; e < h->n_env; ++e) order[e] = e;
        auto mag = [&](int e) { const double* d = push_dv6 + (size_t)e * 6; double m = 0; for (int k = 0; k < 6; ++k) m += d[k] * d[k]; return m; };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            if (push_step[a] != push_step[b]) return push_step[a] < push_step[b];
            return mag(a) > mag(b);
        });
        if (!h->d_order) CU(cudaMalloc(&h->d_order, sizeof(int32_t) * h->n_env));
        CU(cudaMemcpyAsync(h->d_order, order.data(), sizeof(int32_t) * h->n_env, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->has_order = true;
    }
    if (h->has_pushes) {
        CU(cudaMemcpyAsync(h->d_push_step, push_step, sizeof(int32_t) * h->n_env, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->d_push_dv, push_dv6, sizeof(double) * 6 * h->n_env, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return LMH_OK;
}

int32_t lmh_set_step_counter(LmhBatch* h, int32_t step) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    h->step_counter = step;
    return LMH_OK;
}

int32_t lmh_stage(LmhBatch* h, LmhStageDump* dumps) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    Params p = make_params(h);
    p.dump = dumps ? h->d_dump : nullptr;
    dim3 grid(h->n_env), block(h->nt);
    if (h->nt == 32) stage_kernel<32><<<grid, block, h->smem, h->stream>>>(p);
    else if (h->nt == 64) stage_kernel<64><<<grid, block, h->smem, h->stream>>>(p);
    else stage_kernel<128><<<grid, block, h->smem, h->stream>>>(p);
    CU(cudaGetLastError());
    h->launches += 1;
    if (dumps) CU(cudaMemcpyAsync(dumps, h->d_dump, (size_t)h->n_env * sizeof(LmhStageDump), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LMH_OK;
}

int32_t lmh_step(LmhBatch* h, int32_t n_steps) {
    if (!h || n_steps < 0) return fail(LMH_EINVAL, "lmh_step: bad arguments");
    if (n_steps == 0) return LMH_OK;
    CU(cudaSetDevice(h->cfg.device));
    return launch_rollout(h, n_steps);
}

int32_t lmh_step_timed(LmhBatch* h, int32_t n_steps, double* ms) {
    if (!h || n_steps <= 0) return fail(LMH_EINVAL, "lmh_step_timed: bad arguments");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaEventRecord(h->ev0, h->stream));
    int32_t rc = launch_rollout(h, n_steps);
    if (rc != LMH_OK) return rc;
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    float f = 0;
    CU(cudaEventElapsedTime(&f, h->ev0, h->ev1));
    if (ms) *ms = f;
    return LMH_OK;
}

int32_t lmh_sync(LmhBatch* h) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    return LMH_OK;
}

int32_t lmh_get_stats(LmhBatch* h, LmhStats* stats) {
    if (!h || !stats) return fail(LMH_EINVAL, "lmh_get_stats: bad arguments");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(stats, h->d_stats, (size_t)h->n_env * sizeof(LmhStats), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LMH_OK;
}

int32_t lmh_get_log(LmhBatch* h, int32_t first, int32_t n, LmhStepLog* out) {
    if (!h || !out || first < 0 || n < 0 || first + n > h->cfg.log_steps) return fail(LMH_EINVAL, "lmh_get_log: range outside the log");
    if (!h->d_log) return fail(LMH_ESTATE, "lmh_get_log: logging is off (log_steps = 0)");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(out, h->d_log + (size_t)first * h->n_env, (size_t)n * h->n_env * sizeof(LmhStepLog), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LMH_OK;
}

int32_t lmh_get_com(LmhBatch* h, double* com) {
    if (!h || !com) return fail(LMH_EINVAL, "lmh_get_com: bad arguments");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(com, h->d_com, (size_t)h->n_env * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LMH_OK;
}

int32_t lmh_state_dev(LmhBatch* h, double** x, double** vh, double** t, void** stream) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    if (x) *x = h->d_x;
    if (vh) *vh = h->d_vh;
    if (t) *t = h->d_t;
    if (stream) *stream = (void*)h->stream;
    return LMH_OK;
}

int32_t lmh_stats_dev(LmhBatch* h, LmhStats** s) {
    if (!h || !s) return fail(LMH_EINVAL, "bad arguments");
    *s = h->d_stats;
    return LMH_OK;
}

int32_t lmh_set_state_dev(LmhBatch* h, const double* x, const double* vh, const double* t) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    size_t E = (size_t)h->n_env;
    if (x) CU(cudaMemcpyAsync(h->d_x, x, E * 60 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (vh) CU(cudaMemcpyAsync(h->d_vh, vh, E * 30 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (t) CU(cudaMemcpyAsync(h->d_t, t, E * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return LMH_OK;
}

int32_t lmh_reset_stats(LmhBatch* h) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    if (h->h_stats0.size() != (size_t)h->n_env) {
        h->h_stats0.resize(h->n_env);
        std::memset(h->h_stats0.data(), 0, h->n_env * sizeof(LmhStats));
        for (auto& s : h->h_stats0) s.alive = 1;
        CU(cudaMalloc(&h->d_stats0, h->n_env * sizeof(LmhStats)));
        CU(cudaMemcpy(h->d_stats0, h->h_stats0.data(), h->n_env * sizeof(LmhStats), cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpyAsync(h->d_stats, h->d_stats0, (size_t)h->n_env * sizeof(LmhStats), cudaMemcpyDeviceToDevice, h->stream));
    h->step_counter = 0;
    return LMH_OK;
}

int32_t lmh_timer_start(LmhBatch* h) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaEventRecord(h->ev0, h->stream));
    return LMH_OK;
}

int32_t lmh_timer_stop(LmhBatch* h, double* ms) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    float f = 0;
    CU(cudaEventElapsedTime(&f, h->ev0, h->ev1));
    if (ms) *ms = f;
    return LMH_OK;
}

int32_t lmh_qp_bench(LmhBatch* h, int32_t mode, int32_t repeats, double* a_f_out, double* elapsed_ms) {
    if (!h || repeats <= 0 || mode < 0 || mode > 1) return fail(LMH_EINVAL, "lmh_qp_bench: bad arguments");
    CU(cudaSetDevice(h->cfg.device));
    int32_t rc = ensure_scratch(h, (size_t)h->n_env * 42 * sizeof(double));
    if (rc != LMH_OK) return rc;
    double* d_out = h->d_scratch;
    Params p = make_params(h);
    dim3 grid(h->n_env), block(h->nt);
    CU(cudaFuncSetAttribute(qp_bench_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem));
    CU(cudaFuncSetAttribute(qp_bench_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem));
    CU(cudaFuncSetAttribute(qp_bench_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem));
    CU(cudaEventRecord(h->ev0, h->stream));
    if (h->nt == 32) qp_bench_kernel<32><<<grid, block, h->smem, h->stream>>>(p, repeats, mode, d_out);
    else if (h->nt == 64) qp_bench_kernel<64><<<grid, block, h->smem, h->stream>>>(p, repeats, mode, d_out);
    else qp_bench_kernel<128><<<grid, block, h->smem, h->stream>>>(p, repeats, mode, d_out);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    float f = 0;
    CU(cudaEventElapsedTime(&f, h->ev0, h->ev1));
    if (elapsed_ms) *elapsed_ms = f;
    if (a_f_out) CU(cudaMemcpy(a_f_out, d_out, (size_t)h->n_env * 42 * sizeof(double), cudaMemcpyDeviceToHost));
    h->launches += 1;
    return LMH_OK;
}

int32_t lmh_ik_batch(LmhBatch* h, int32_t n, const double* com3, const double* rf6, const double* lf6, double* q_out,
                     int32_t* iters, double* residual, int32_t* status, int32_t apply_to_state) {
    if (!h || n <= 0 || !com3 || !rf6 || !lf6) return fail(LMH_EINVAL, "lmh_ik_batch: bad arguments");
    if (apply_to_state && n != h->n_env) return fail(LMH_EINVAL, "lmh_ik_batch: apply_to_state needs one target per environment");
    CU(cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)n;
    // one scratch block, kept in the handle: inputs (15 N), postures (30 N), residuals (N), start posture (30), integers (2 N)
    int32_t rc0 = ensure_scratch(h, (N * 47 + NQ) * sizeof(double) + N * 2 * sizeof(int32_t));
    if (rc0 != LMH_OK) return rc0;
    double *d_in = h->d_scratch, *d_q = d_in + N * 15, *d_res = d_q + N * NQ, *d_q0 = d_res + N;
    int32_t* d_int = reinterpret_cast<int32_t*>(d_q0 + NQ);
    CU(cudaMemcpyAsync(d_in, com3, N * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_in + N * 3, rf6, N * 6 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_in + N * 9, lf6, N * 6 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_q0, h->hm->q0, NQ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    IkParams p;
    p.K = h->dK; p.q_start = d_q0; p.q_start_each = nullptr; p.desop = nullptr; p.com = d_in; p.rf = d_in + N * 3; p.lf = d_in + N * 9;
    p.q_out = d_q; p.iters = d_int; p.status = d_int + N; p.residual = d_res;
    p.x = apply_to_state ? h->d_x : nullptr; p.vh = h->d_vh; p.t = h->d_t;
    p.n = n; p.max_iter = 200;
    p.env_bytes = (int32_t)((sizeof(EnvWS) + 15) / 16 * 16);
    CU(cudaFuncSetAttribute(ik_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p.env_bytes * 4));
    ik_kernel<<<(n + 3) / 4, 128, (size_t)p.env_bytes * 4, h->stream>>>(p);
    CU(cudaGetLastError());
    h->launches += 1;
    if (q_out) CU(cudaMemcpyAsync(q_out, d_q, N * NQ * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (iters) CU(cudaMemcpyAsync(iters, d_int, N * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (status) CU(cudaMemcpyAsync(status, d_int + N, N * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (residual) CU(cudaMemcpyAsync(residual, d_res, N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (apply_to_state) {
        std::vector<LmhStats> st(N);
        std::memset(st.data(), 0, N * sizeof(LmhStats));
        for (size_t e = 0; e < N; ++e) st[e].alive = 1;
        CU(cudaMemcpyAsync(h->d_stats, st.data(), N * sizeof(LmhStats), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->step_counter = 0;
    }
    return LMH_OK;
}


int32_t lmh_ik_solve(LmhBatch* h, int32_t n, const double* q_start, const double* desop, double* q_out, int32_t* iters,
                     double* residual, int32_t* status) {
    if (!h || n <= 0 || !desop) return fail(LMH_EINVAL, "lmh_ik_solve: bad arguments");
    CU(cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)n;
    // scratch: targets (30 N), starting postures (30 N), postures (30 N), residuals (N), nominal posture (30), integers (2 N)
    int32_t rc0 = ensure_scratch(h, (N * 91 + NQ) * sizeof(double) + N * 2 * sizeof(int32_t));
    if (rc0 != LMH_OK) return rc0;
    double *d_des = h->d_scratch, *d_qs = d_des + N * NQ, *d_q = d_qs + N * NQ, *d_res = d_q + N * NQ, *d_q0 = d_res + N;
    int32_t* d_int = reinterpret_cast<int32_t*>(d_q0 + NQ);
    CU(cudaMemcpyAsync(d_des, desop, N * NQ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (q_start) CU(cudaMemcpyAsync(d_qs, q_start, N * NQ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_q0, h->hm->q0, NQ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    IkParams p;
    p.K = h->dK; p.q_start = d_q0; p.q_start_each = q_start ? d_qs : nullptr; p.desop = d_des; p.com = nullptr; p.rf = nullptr; p.lf = nullptr;
    p.q_out = d_q; p.iters = d_int; p.status = d_int + N; p.residual = d_res;
    p.x = nullptr; p.vh = nullptr; p.t = nullptr;
    p.n = n; p.max_iter = 200;
    p.env_bytes = (int32_t)((sizeof(EnvWS) + 15) / 16 * 16);
    CU(cudaFuncSetAttribute(ik_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p.env_bytes * 4));
    ik_kernel<<<(n + 3) / 4, 128, (size_t)p.env_bytes * 4, h->stream>>>(p);
    CU(cudaGetLastError());
    h->launches += 1;
    if (q_out) CU(cudaMemcpyAsync(q_out, d_q, N * NQ * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (iters) CU(cudaMemcpyAsync(iters, d_int, N * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (status) CU(cudaMemcpyAsync(status, d_int + N, N * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (residual) CU(cudaMemcpyAsync(residual, d_res, N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LMH_OK;
}

int32_t lmh_kinematics(LmhBatch* h, int32_t n, const double* q, double* T, double* X, double* com) {
    if (!h || n <= 0 || !q) return fail(LMH_EINVAL, "lmh_kinematics: bad arguments");
    CU(cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)n, rec = (size_t)NF * 12;
    int32_t rc0 = ensure_scratch(h, N * (NQ + 2 * rec + 3) * sizeof(double));
    if (rc0 != LMH_OK) return rc0;
    double *d_q = h->d_scratch, *d_T = d_q + N * NQ, *d_X = d_T + N * rec, *d_c = d_X + N * rec;
    CU(cudaMemcpyAsync(d_q, q, N * NQ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const int env_bytes = (int)((sizeof(EnvWS) + 15) / 16 * 16);
    CU(cudaFuncSetAttribute(kin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, env_bytes * 4));
    kin_kernel<<<(n + 3) / 4, 128, (size_t)env_bytes * 4, h->stream>>>(h->dK, d_q, d_T, d_X, d_c, n, env_bytes);
    CU(cudaGetLastError());
    h->launches += 1;
    std::vector<double> hT(N * rec), hX(N * rec);
    CU(cudaMemcpyAsync(hT.data(), d_T, N * rec * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(hX.data(), d_X, N * rec * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (com) CU(cudaMemcpyAsync(com, d_c, N * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    // expand the (R, p) records: 4 x 4 homogeneous transforms and dense motion transforms [[E,0],[-E p^,E]], E = R'
    // (velocityMatrix, generalizedFunctions.cpp:11-19); frame 27 = frame 26 times the constant DH transform of the extra
    // head frame (theta = -pi/2, alpha = 0, r = 0, d = 0.030: Robot.cpp:84, 176-222)
    const double kPi = 3.14159265358979323846;
    const double c24 = std::cos(-kPi / 2), s24 = std::sin(-kPi / 2);
    const double Rp27[12] = {c24, -s24, 0.0, s24, c24, 0.0, 0.0, 0.0, 1.0, 0.030, 0.0, 0.0};     // local transform of frame 27
    for (size_t e = 0; e < N; ++e) {
        for (int i = 0; i < 28; ++i) {
            double Tw[12], Xr[12];
            if (i < NF) {
                std::memcpy(Tw, hT.data() + e * rec + 12 * i, sizeof(Tw));
                std::memcpy(Xr, hX.data() + e * rec + 12 * i, sizeof(Xr));
            } else {
                const double* P = hT.data() + e * rec + 12 * 26;
                for (int r = 0; r < 3; ++r) {
                    for (int c = 0; c < 3; ++c) Tw[3 * r + c] = P[3 * r] * Rp27[c] + P[3 * r + 1] * Rp27[3 + c] + P[3 * r + 2] * Rp27[6 + c];
                    Tw[9 + r] = P[3 * r] * Rp27[9] + P[3 * r + 1] * Rp27[10] + P[3 * r + 2] * Rp27[11] + P[9 + r];
                }
                // inverseTransformationMatrix(T26) * T27 (Robot.cpp:276-287), as the device does for the other frames
                for (int r = 0; r < 3; ++r) {
                    for (int c = 0; c < 3; ++c) Xr[3 * r + c] = P[r] * Tw[c] + P[3 + r] * Tw[3 + c] + P[6 + r] * Tw[6 + c];
                    Xr[9 + r] = (P[r] * Tw[9] + P[3 + r] * Tw[10] + P[6 + r] * Tw[11]) - (P[r] * P[9] + P[3 + r] * P[10] + P[6 + r] * P[11]);
                }
            }
            if (T) {
                double* o = T + (e * 28 + i) * 16;
                for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) o[4 * r + c] = Tw[3 * r + c]; o[4 * r + 3] = Tw[9 + r]; }
                o[12] = o[13] = o[14] = 0.0; o[15] = 1.0;
            }
            if (X) {
                double* o = X + (e * 28 + i) * 36;
                for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) o[6 * a + b] = xdense(Xr, a, b);
            }
        }
    }
    return LMH_OK;
}

static int32_t upload_mpc_tables(LmhBatch* h) {
    if (h->d_mpc) return LMH_OK;
    const HostModel& m = *h->hm;
    const size_t np = m.mpc_np, nt = np / 8;
    const size_t total = 3 * np * np + 2 * nt * 64 + 2 * np + np;
    CU(cudaMalloc(&h->d_mpc, total * sizeof(double)));
    double* d = h->d_mpc;
    const std::vector<double>* parts[7] = {&m.mpc_PuT, &m.mpc_L, &m.mpc_LT, &m.mpc_Dinv, &m.mpc_DinvT, &m.mpc_Px, &m.mpc_rdiag};
    for (const auto* v : parts) {
        CU(cudaMemcpy(d, v->data(), v->size() * sizeof(double), cudaMemcpyHostToDevice));
        d += v->size();
    }
    return LMH_OK;
}

int32_t lmh_mpc_solve(LmhBatch* h, int32_t n, const double* com_xy, const double* vel_xy, const double* zmp_x,
                      const double* zmp_y, double* xref, double* yref, int32_t use_dmma, int32_t repeats, double* elapsed_ms) {
    if (!h || n <= 0 || !com_xy || !vel_xy || !zmp_x || !zmp_y || repeats < 1 || use_dmma < 0 || use_dmma > 1)
        return fail(LMH_EINVAL, "lmh_mpc_solve: bad arguments");
    CU(cudaSetDevice(h->cfg.device));
    int32_t rc = upload_mpc_tables(h);
    if (rc != LMH_OK) return rc;
    const HostModel& m = *h->hm;
    const size_t N = (size_t)n, n1 = (size_t)m.K.mpc_n1, np = (size_t)m.mpc_np, nt = np / 8;
    rc = ensure_scratch(h, N * (4 + 2 * n1 + 6) * sizeof(double));
    if (rc != LMH_OK) return rc;
    double *d_com = h->d_scratch, *d_vel = d_com + 2 * N, *d_zx = d_vel + 2 * N, *d_zy = d_zx + N * n1, *d_xr = d_zy + N * n1, *d_yr = d_xr + 3 * N;
    CU(cudaMemcpyAsync(d_com, com_xy, 2 * N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_vel, vel_xy, 2 * N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_zx, zmp_x, N * n1 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_zy, zmp_y, N * n1 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    MpcDev M;
    M.PuT = h->d_mpc; M.L = M.PuT + np * np; M.LT = M.L + np * np; M.Dinv = M.LT + np * np; M.DinvT = M.Dinv + nt * 64;
    M.Px = M.DinvT + nt * 64; M.rdiag = M.Px + 2 * np;
    for (int k = 0; k < 4; ++k) M.A[k] = m.K.mpc_A[k];
    M.B[0] = m.K.mpc_B[0]; M.B[1] = m.K.mpc_B[1];
    M.n = (int32_t)n1; M.np = (int32_t)np;
    MpcProblemSet P{d_com, d_vel, d_zx, d_zy, d_xr, d_yr, n};
    CU(cudaEventRecord(h->ev0, h->stream));
    if (use_dmma) {
        const size_t smem = 4 * (np * 8 + 64) * sizeof(double);
        CU(cudaFuncSetAttribute(mpc_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mpc_dmma_kernel<<<(n + 15) / 16, 128, smem, h->stream>>>(M, P, repeats);
    } else {
        const size_t smem = 2 * np * 32 * sizeof(double);
        CU(cudaFuncSetAttribute(mpc_dfma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mpc_dfma_kernel<<<(n + 31) / 32, 64, smem, h->stream>>>(M, P, repeats);
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev1, h->stream));
    h->launches += 1;
    if (xref) CU(cudaMemcpyAsync(xref, d_xr, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (yref) CU(cudaMemcpyAsync(yref, d_yr, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float f = 0;
    CU(cudaEventElapsedTime(&f, h->ev0, h->ev1));
    if (elapsed_ms) *elapsed_ms = f;
    return LMH_OK;
}

int32_t lmh_get_mpc_tables(const LmhBatch* h, double* Px, double* Pu, int32_t cap, int32_t* n1) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    const int n = h->hm->K.mpc_n1;
    if (n1) *n1 = n;
    if ((Px || Pu) && cap < n) return fail(LMH_EINVAL, "lmh_get_mpc_tables: capacity below horizon + 1");
    if (Px) std::memcpy(Px, h->hm->mpc_Px_raw.data(), sizeof(double) * 2 * n);
    if (Pu) std::memcpy(Pu, h->hm->mpc_Pu_raw.data(), sizeof(double) * (size_t)n * n);
    return LMH_OK;
}

int32_t lmh_foot_coeff(const double* cur3, const double* des3, double step_height, double T, double* x6, double* y6, double* z8) {
    if (!cur3 || !des3 || !(T > 0.0)) return fail(LMH_EINVAL, "lmh_foot_coeff: bad arguments");
    double c[20];
    try { foot_coefficients(cur3, des3, step_height, T, c); }
    catch (const std::exception& ex) { return fail(LMH_EINVAL, std::string("lmh_foot_coeff: ") + ex.what()); }
    if (x6) std::memcpy(x6, c, 6 * sizeof(double));
    if (y6) std::memcpy(y6, c + 6, 6 * sizeof(double));
    if (z8) std::memcpy(z8, c + 12, 8 * sizeof(double));
    return LMH_OK;
}

int32_t lmh_push_plan(const int64_t* env_ids, int32_t n, uint64_t seed, double max_speed, int32_t max_step, int32_t* push_step,
                      double* push_dv6) {
    if (n < 0 || (n > 0 && (!env_ids || !push_step || !push_dv6)) || max_step < 0) return fail(LMH_EINVAL, "lmh_push_plan: bad arguments");
    push_plan(env_ids, n, seed, max_speed, max_step, push_step, push_dv6);
    return LMH_OK;
}

void lmh_default_walk_spec(LmhWalkSpec* s) {
    WalkSpec d;
    std::memset(s, 0, sizeof(*s));
    s->n_footsteps = d.n_footsteps; s->horizon_steps = d.horizon_steps; s->tail_steps = d.tail_steps;
    s->time_per_step = d.time_per_step; s->dt = d.dt; s->step_height = d.step_height; s->lateral = d.lateral;
    s->min_len = d.min_len; s->max_len = d.max_len; s->fixed_len = -1.0;
}

static bool walk_spec_from(const LmhWalkSpec* s, WalkSpec& w) {
    if (!s || s->n_footsteps < 0 || s->horizon_steps < 0 || s->tail_steps < 0 || !(s->time_per_step > 0.0) || !(s->dt > 0.0)) return false;
    w.n_footsteps = s->n_footsteps; w.horizon_steps = s->horizon_steps; w.tail_steps = s->tail_steps;
    w.time_per_step = s->time_per_step; w.dt = s->dt; w.step_height = s->step_height; w.lateral = s->lateral;
    w.min_len = s->min_len; w.max_len = s->max_len;
    return w.per() >= 1;
}

int32_t lmh_walk_plan_dims(const LmhWalkSpec* spec, int32_t* n_phases, int32_t* zmp_len) {
    WalkSpec w;
    if (!walk_spec_from(spec, w)) return fail(LMH_EINVAL, "lmh_walk_plan_dims: bad walk spec");
    if (n_phases) *n_phases = w.n_phases();
    if (zmp_len) *zmp_len = w.zmp_len();
    return LMH_OK;
}

int32_t lmh_walk_plan(const LmhWalkSpec* spec, const int64_t* env_ids, int32_t n, uint64_t seed, double* zmp_x, double* zmp_y,
                      double* foot_coef, int32_t* contact, double* step_len) {
    WalkSpec w;
    if (!walk_spec_from(spec, w) || n < 0 || (n > 0 && (!env_ids || !zmp_x || !zmp_y || !foot_coef || !contact)))
        return fail(LMH_EINVAL, "lmh_walk_plan: bad arguments");
    try { walk_plans(w, env_ids, n, seed, spec->fixed_len, zmp_x, zmp_y, foot_coef, contact, step_len); }
    catch (const std::exception& ex) { return fail(LMH_EINVAL, std::string("lmh_walk_plan: ") + ex.what()); }
    return LMH_OK;
}

int32_t lmh_make_walk_plans(LmhBatch* h, const LmhWalkSpec* spec, int64_t first_env_id, uint64_t seed) {
    WalkSpec w;
    if (!h || !walk_spec_from(spec, w)) return fail(LMH_EINVAL, "lmh_make_walk_plans: bad arguments");
    const size_t E = (size_t)h->n_env, P = (size_t)w.n_phases(), Z = (size_t)w.zmp_len();
    std::vector<int64_t> ids(E);
    std::vector<int32_t> ct(E * P), pid(E);
    std::vector<double> zx(E * Z), zy(E * Z), fc(E * P * 40);
    for (size_t e = 0; e < E; ++e) { ids[e] = first_env_id + (int64_t)e; pid[e] = (int32_t)e; }
    int32_t rc = lmh_walk_plan(spec, ids.data(), h->n_env, seed, zx.data(), zy.data(), fc.data(), ct.data(), nullptr);
    if (rc != LMH_OK) return rc;
    return lmh_set_walk_plans(h, h->n_env, (int32_t)Z, zx.data(), zy.data(), (int32_t)P, w.time_per_step, fc.data(), ct.data(), pid.data());
}

int32_t lmh_export_mujoco(int32_t n, const double* q30, double* qpos31) {
    if (n < 0 || !q30 || !qpos31) return fail(LMH_EINVAL, "lmh_export_mujoco: bad arguments");
    // row block -> first source joint, length (apps/mujoco/main.cpp:182-200): head, left leg, right leg, left arm, right arm
    static const int row0[5] = {0, 2, 8, 14, 19}, src0[5] = {22, 6, 0, 17, 12}, len[5] = {2, 6, 6, 5, 5};
    for (int e = 0; e < n; ++e) {
        const double* q = q30 + (size_t)e * 30;
        double* o = qpos31 + (size_t)e * 31;
        o[0] = q[0]; o[1] = q[1]; o[2] = q[2] + 0.035;            // MujocoSim.cpp:77-79: offset between the base frames
        o[3] = 1.0; o[4] = o[5] = o[6] = 0.0;                      // MujocoSim.cpp:82-85: identity quaternion, as upstream
        for (int b = 0; b < 5; ++b)
            for (int k = 0; k < len[b]; ++k) o[7 + row0[b] + k] = q[6 + src0[b] + k];
        o[7 + 14] = -q[6 + 17];                                    // L(14,17) = -1: LShoulderPitch turns the other way
    }
    return LMH_OK;
}

int32_t lmh_launch_info(const LmhBatch* h, int32_t* nt, int32_t* epb, int32_t* smem, int32_t* blocks, int32_t* launches) {
    if (!h) return fail(LMH_EINVAL, "null handle");
    if (nt) *nt = h->nt;
    if (epb) *epb = h->epb;
    if (smem) *smem = h->smem;
    if (blocks) *blocks = (h->nt == 32 && !h->small_batch) ? multi_grid(h, nullptr) : h->n_env;
    if (launches) *launches = h->launches;
    return LMH_OK;
}

int32_t lmh_measure_fp64_peak(int32_t device, double* tflops, double* sm_mhz, int32_t* sm_count) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) return fail(LMH_ENODEVICE, "lmh_measure_fp64_peak: no CUDA device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 4096;
    double* d = nullptr;
    CU(cudaMalloc(&d, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(a));
        fp64_peak_kernel<<<blocks, threads>>>(d, iters);
        CU(cudaEventRecord(b));
        CU(cudaEventSynchronize(b));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, a, b));
        double flops = 2.0 * 64.0 * iters * (double)threads * blocks;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    CU(cudaGetLastError());
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    if (tflops) *tflops = best;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, device);
    if (sm_mhz) *sm_mhz = clk / 1000.0;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return LMH_OK;
}

}  // extern "C"
