// lmh_warp.cuh -- one-warp-per-environment (NT = 32) device versions of the tree recursions (RNEA bias
// forces, CRBA).  Same inputs and outputs in the workspace as rnea_phase / crba_phase of lmh_model.cuh (which
// stay the reference implementation: host-testable, used for every other group size, and the thing these are
// checked against on the GPU), but organised around fixed lane roles instead of a parallel-for over items:
//
//   lane = 6*c + r :  chain c (0 right leg, 1 left leg, 2 right arm, 3 left arm, 4 head), spatial component /
//                     matrix row r.  Lanes 30, 31 idle in the chain phases.
//
// A lane keeps "its" row of the composite inertia and "its" joint's force vector in registers across tree
// levels; what the other five lanes of a chain need travels through a few shared-memory words with one
// __syncwarp per exchange.  No index arithmetic survives in the inner loops.
#pragma once
#if defined(__CUDACC__)

namespace lmh {

struct LaneRole {
    int c, r;          // chain, component
    int first, len, nact, j0;
    bool on;           // lane < 30
};

__device__ __forceinline__ LaneRole lane_role(int lane) {
    LaneRole L;
    L.on = lane < 30;
    L.c = L.on ? lane / 6 : 4;
    L.r = L.on ? lane % 6 : 0;
    L.first = chain_first(L.c);
    L.len = chain_len(L.c);
    L.nact = chain_nact(L.c);
    L.j0 = chain_joint0(L.c);
    return L;
}

// p ? a : b as a select instruction: lane-dependent choices between values that every lane has anyway must not
// become divergent branches (a lone warp pays ~15 cycles per resolved branch; the FP64 and ALU pipes are mostly idle)
__device__ __forceinline__ double dsel(bool p, double a, double b) {
    double r;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\tselp.f64 %0, %1, %2, q;\n\t}" : "=d"(r) : "d"(a), "d"(b), "r"((int)p));
    return r;
}
__device__ __forceinline__ double dsel3(int k, double x0, double x1, double x2) { return dsel(k == 0, x0, dsel(k == 1, x1, x2)); }
// component k (lane-dependent) of a x b: all three components, then a select
__device__ __forceinline__ double cross3_sel(const double* a, const double* b, int k) {
    return dsel3(k, a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
// Experiment switches (A/B builds: scripts/ab_bench.py); the defaults are what measured best on B200 (config 2, E = 4096):
//   LMH_BCAST64 = 0, LMH_XSTRIDE = 8 : 10.92 M env-steps/s     LMH_BCAST64 = 1, LMH_XSTRIDE = 8 : 10.44 M
//   LMH_BCAST64 = 0, LMH_XSTRIDE = 6 : 11.01 M (default)       LMH_BCAST64 = 1, LMH_XSTRIDE = 6 : 10.67 M
// i.e. trading shared-memory wavefronts for issue slots loses: at ten warps per SM the kernel is bound by the latency of
// its dependent instruction chains, not by the shared-memory pipe alone (gpurun_out/r2_ab2.log).
#ifndef LMH_XSTRIDE
#define LMH_XSTRIDE 6      // doubles between the chains' slots in the exchange buffers: 6 = lane order (lane 6c + r writes word
                           // 6c + r: no bank conflicts), 8 = one 64-byte line per chain
#endif
#ifndef LMH_FK_UNROLL
#define LMH_FK_UNROLL 1
#endif
#ifndef LMH_FACTOR_DMMA
#define LMH_FACTOR_DMMA 1  // 1: Schur complement of the leg blocks in the 18 x 18 factorisation on the FP64 tensor cores
#endif
#ifndef LMH_JTJ_DMMA
#define LMH_JTJ_DMMA 1     // 1: J'J of the sole Jacobians (the J'W_F J term of the QP Hessian) on the FP64 tensor cores
#endif
#ifndef LMH_BCAST64
#define LMH_BCAST64 0      // 1: chain-broadcast six-vectors / transforms through LDS.64 (one wavefront each) instead of LDS.128
#endif
// A six-vector / an (R, p) record that EVERY lane of a chain reads from the same address: five distinct addresses per
// warp.  An LDS.64 serves that in one wavefront; an LDS.128 is processed a quarter-warp at a time (3-4 wavefronts for 16
// bytes per lane), so the scalar form moves the same data through the shared-memory pipe in 6 instead of ~10 wavefronts.
// (explicit ld.shared.f64: plain C++ loads would be fused back into LDS.128, and C++ volatile loads are strongly ordered
// -- one in flight at a time, which halved the throughput when it was tried)
__device__ __forceinline__ void ld6b(const double* p, double* o) {
#if LMH_BCAST64
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
#pragma unroll
    for (int k = 0; k < 6; ++k) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(o[k]) : "r"(a + 8u * k) : "memory");
#else
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int k = 0; k < 3; ++k) { const double2 t = q[k]; o[2 * k] = t.x; o[2 * k + 1] = t.y; }
#endif
}
__device__ __forceinline__ void ld12b(const double* p, double* o) {
#if LMH_BCAST64
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
#pragma unroll
    for (int k = 0; k < 12; ++k) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(o[k]) : "r"(a + 8u * k) : "memory");
#else
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int k = 0; k < 6; ++k) { const double2 t = q[k]; o[2 * k] = t.x; o[2 * k + 1] = t.y; }
#endif
}
// 16-byte aligned sources, a different address per lane: LDS.128, half the shared-memory wavefronts of the scalar loads
__device__ __forceinline__ void ld6v(const double* p, double* o) {
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int k = 0; k < 3; ++k) { const double2 t = q[k]; o[2 * k] = t.x; o[2 * k + 1] = t.y; }
}
__device__ __forceinline__ void ld12v(const double* p, double* o) {
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int k = 0; k < 6; ++k) { const double2 t = q[k]; o[2 * k] = t.x; o[2 * k + 1] = t.y; }
}
// all six components of X' f : [R n + p x (R f); R f]
__device__ __forceinline__ void xf_all(const double* Rp, const double* f, double* o) {
    const double r0 = Rp[0] * f[3] + Rp[1] * f[4] + Rp[2] * f[5];
    const double r1 = Rp[3] * f[3] + Rp[4] * f[4] + Rp[5] * f[5];
    const double r2 = Rp[6] * f[3] + Rp[7] * f[4] + Rp[8] * f[5];
    o[0] = Rp[0] * f[0] + Rp[1] * f[1] + Rp[2] * f[2] + (Rp[10] * r2 - Rp[11] * r1);
    o[1] = Rp[3] * f[0] + Rp[4] * f[1] + Rp[5] * f[2] + (Rp[11] * r0 - Rp[9] * r2);
    o[2] = Rp[6] * f[0] + Rp[7] * f[1] + Rp[8] * f[2] + (Rp[9] * r1 - Rp[10] * r0);
    o[3] = r0; o[4] = r1; o[5] = r2;
}
__device__ __forceinline__ double cross3(const double* a, const double* b, int k) {
    return k == 0 ? a[1] * b[2] - a[2] * b[1] : k == 1 ? a[2] * b[0] - a[0] * b[2] : a[0] * b[1] - a[1] * b[0];
}


// ---- torso pose (generalizedFunctions.cpp:52-72) and local DH transforms (Robot.cpp:176-223): ONE sincos per lane --
// lanes 1..26 their frame's joint angle, lanes 27..29 roll, pitch, yaw, which reach lane 0 by shuffle (three more
// sincos calls on lane 0 alone would cost the whole warp three times the call).  The fixed torso->chain transforms
// are applied by fk_chain_warp, so every joint frame stores its plain DH matrix.
__device__ void fk_local_warp(EnvWS& w, const Consts& K, const double* dh, int lane) {
    const bool jointf = lane >= 1 && lane < NF && lane != 7 && lane != 14;
    const int a = jointf ? frame_joint(lane) : 0;
    double ang = 0.0;
    if (jointf) ang = w.q[6 + a] + dh[a];
    else if (lane >= 27 && lane < 30) ang = w.q[3 + lane - 27];
    double st, ct;
    sincos(ang, &st, &ct);
    const double sr = __shfl_sync(0xffffffffu, st, 27), cr = __shfl_sync(0xffffffffu, ct, 27);
    const double sp = __shfl_sync(0xffffffffu, st, 28), cp = __shfl_sync(0xffffffffu, ct, 28);
    const double sy = __shfl_sync(0xffffffffu, st, 29), cy = __shfl_sync(0xffffffffu, ct, 29);
    if (lane == 0) {
        double R[12] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                        sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                        -sp,     cp * sr,                cp * cr,
                        w.q[0], w.q[1], w.q[2]};
        double* T0 = w.u.m.p.Tw;
#pragma unroll
        for (int k = 0; k < 12; ++k) T0[k] = R[k];
#pragma unroll
        for (int k = 0; k < 9; ++k) w.R0[k] = R[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) w.p0[k] = R[9 + k];
        w.esc[0] = sp; w.esc[1] = cp; w.esc[2] = sy; w.esc[3] = cy;
    } else if (lane < NF) {
        // soles: identity rotation, offset along x (Robot.cpp:106-117); joints: Khalil's modified DH (Robot.cpp:200-220)
        const double ca = jointf ? dh[24 + a] : 1.0, sa = jointf ? dh[48 + a] : 0.0;
        const double r = jointf ? dh[72 + a] : 0.0, d = jointf ? dh[96 + a] : K.sole_x;
        double* O = w.u.m.Xl + 12 * lane;
        O[0] = ct; O[1] = -st; O[2] = 0.0;
        O[3] = ca * st; O[4] = ca * ct; O[5] = -sa;
        O[6] = sa * st; O[7] = sa * ct; O[8] = ca;
        O[9] = d; O[10] = -r * sa; O[11] = r * ca;
    }
    __syncwarp();
}

// ---- forward kinematics after the local transforms are in w.u.m.Xl (Robot.cpp:45-160,225-238,276-298):
// lane 3c + r3 (< 15) carries row r3 of the world transform of chain c's current frame down the chain (a row-local
// recurrence: no exchange between levels); then lane i (< 27) turns frame i's local transform into the
// parent-relative one and adds its link to the centre of mass.
__device__ void fk_chain_warp(EnvWS& w, const Consts& K, int lane) {
    double* Tw = w.u.m.p.Tw;
    double* Xl = w.u.m.Xl;
    if (lane < 15) {
        const int c = lane / 3, r3 = lane % 3;
        const int first = chain_first(c), len = chain_len(c);
        double R0 = Tw[3 * r3], R1 = Tw[3 * r3 + 1], R2 = Tw[3 * r3 + 2], pp = Tw[9 + r3];   // torso row
        {       // times the fixed torso->chain transform (Robot.cpp:92-103,134-152): same for all fifteen lanes, so it
                // is done here and not as a divergent special case of the five first-joint frames in fk_phase
            const double* P = K.pre + 12 * c;
            const double n0 = R0 * P[0] + R1 * P[3] + R2 * P[6];
            const double n1 = R0 * P[1] + R1 * P[4] + R2 * P[7];
            const double n2 = R0 * P[2] + R1 * P[5] + R2 * P[8];
            pp = R0 * P[9] + R1 * P[10] + R2 * P[11] + pp;
            R0 = n0; R1 = n1; R2 = n2;
        }
        // (unrolled: the local transforms of all levels are there already, so their loads can be issued ahead of the recurrence)
#if LMH_FK_UNROLL
#pragma unroll
        for (int d = 0; d < MAXDEPTH; ++d) {
            if (d >= len) break;
            const double* L = Xl + 12 * (first + d);
            double* O3 = Tw + 12 * (first + d) + 3 * r3;
            double* O9 = Tw + 12 * (first + d) + 9 + r3;
#else
#pragma unroll 1
        for (int d = 0; d < len; ++d) {
            const double* L = Xl + 12 * (first + d);
            double* O3 = Tw + 12 * (first + d) + 3 * r3;
            double* O9 = Tw + 12 * (first + d) + 9 + r3;
#endif
            const double n0 = R0 * L[0] + R1 * L[3] + R2 * L[6];
            const double n1 = R0 * L[1] + R1 * L[4] + R2 * L[7];
            const double n2 = R0 * L[2] + R1 * L[5] + R2 * L[8];
            pp = R0 * L[9] + R1 * L[10] + R2 * L[11] + pp;
            R0 = n0; R1 = n1; R2 = n2;
            O3[0] = R0; O3[1] = R1; O3[2] = R2; *O9 = pp;
        }
    }
    __syncwarp();
    double cm0 = 0.0, cm1 = 0.0, cm2 = 0.0;
    if (lane < NF) {
        double Ti[12];
        ld12v(Tw + 12 * lane, Ti);
        double* O = Xl + 12 * lane;
        {
            // the torso (lane 0) keeps its world transform: an identity parent gives exactly that (1 x + 0 y + 0 z = x),
            // without a divergent special case
            double P[12];
            ld12v(Tw + 12 * (lane == 0 ? 0 : frame_parent(lane)), P);
#pragma unroll
            for (int e = 0; e < 12; ++e) P[e] = lane == 0 ? ((e == 0 || e == 4 || e == 8) ? 1.0 : 0.0) : P[e];
#pragma unroll
            for (int r3 = 0; r3 < 3; ++r3) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) O[3 * r3 + cc] = P[r3] * Ti[cc] + P[3 + r3] * Ti[3 + cc] + P[6 + r3] * Ti[6 + cc];
                const double pi = P[r3] * Ti[9] + P[3 + r3] * Ti[10] + P[6 + r3] * Ti[11];
                const double pq = P[r3] * P[9] + P[3 + r3] * P[10] + P[6 + r3] * P[11];
                O[9 + r3] = pi - pq;
            }
        }
        const double* mc = K.mc + 4 * lane;
        const double m = mc[3];
        cm0 = m * (Ti[0] * mc[0] + Ti[1] * mc[1] + Ti[2] * mc[2] + Ti[9]);
        cm1 = m * (Ti[3] * mc[0] + Ti[4] * mc[1] + Ti[5] * mc[2] + Ti[10]);
        cm2 = m * (Ti[6] * mc[0] + Ti[7] * mc[1] + Ti[8] * mc[2] + Ti[11]);
        if (lane == 7 || lane == 14) {
            const int s = lane == 7 ? 0 : 1;
#pragma unroll
            for (int e = 0; e < 9; ++e) w.Rs[s][e] = Ti[e];
#pragma unroll
            for (int e = 0; e < 3; ++e) w.ps[s][e] = Ti[9 + e];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cm0 += __shfl_xor_sync(0xffffffffu, cm0, o);
        cm1 += __shfl_xor_sync(0xffffffffu, cm1, o);
        cm2 += __shfl_xor_sync(0xffffffffu, cm2, o);
    }
    if (lane == 0) { w.com[0] = cm0 * K.inv_mass; w.com[1] = cm1 * K.inv_mass; w.com[2] = cm2 * K.inv_mass; }
    __syncwarp();
}

// ---- RNEA with zero base acceleration on the held velocity (Dynamics.cpp:29-60,124-200) -> Cg, Jpqp
// Every lane runs every level (idle lanes shadow the last frame of their chain and store nothing): no divergent
// regions, so the exchanges of the backward pass can be shuffles.  A lane reads only the column of R it needs
// straight from shared memory (a lane-dependent index into a register copy would put the copy in local memory).
__device__ void rnea_warp(EnvWS& w, const Consts& K, const double* Isp, const LaneRole& L, int lane) {
    auto& S = w.u.m.p.wr;
    const double* Xl = w.u.m.Xl;
    const int c8 = LMH_XSTRIDE * L.c, r = L.r;
    const bool lin = r >= 3;
    const int k3 = lin ? r - 3 : r;
    const double csgn = (k3 == 0) ? 1.0 : (k3 == 1) ? -1.0 : 0.0;
    // base: f0 = v0 x* (I0 v0), acc0 = 0
    if (lane < 6) {
        double v0[6];
        ld6b(w.vhb6, v0);
        const double* I = Isp + 6 * lane;
        S.hbuf[lane] = I[0] * v0[0] + I[1] * v0[1] + I[2] * v0[2] + I[3] * v0[3] + I[4] * v0[4] + I[5] * v0[5];
    }
    __syncwarp();
    if (lane < 6) {
        double v0[6], h[6];
        ld6b(w.vhb6, v0);
        ld6b(S.hbuf, h);
        S.frc[lane] = lane < 3 ? cross3(v0, h, lane) + cross3(v0 + 3, h + 3, lane) : cross3(v0, h + 3, lane - 3);
    }
    int cur = 0;
    // velocity and acceleration of the frame above, all six components, carried in registers from level to level
    // (every lane of a chain reads the new ones anyway): two exchange-buffer reads per level instead of four
    double pv[6], pa[6] = {0, 0, 0, 0, 0, 0};
    ld6b(w.vhb6, pv);
#pragma unroll 1
    for (int d = 1; d <= MAXDEPTH; ++d) {
        const bool act = L.on && d <= L.len;
        const int i = L.first + (d <= L.len ? d : L.len) - 1;
        const bool joint = act && d <= L.nact;
        const double* X = Xl + 12 * i;
        const double e0 = X[k3], e1 = X[3 + k3], e2 = X[6 + k3];     // column k3 of R = row k3 of E
        const double p0 = X[9], p1 = X[10], p2 = X[11];
        // row r of X m = a . m_ang + b . m_lin: angular rows (a, b) = (e, 0); linear rows e . (m_lin + m_ang x p) =
        // (p x e) . m_ang + e . m_lin.  The pair serves both the velocity and the acceleration of this level.
        const double a0 = dsel(lin, p1 * e2 - p2 * e1, e0), a1 = dsel(lin, p2 * e0 - p0 * e2, e1), a2 = dsel(lin, p0 * e1 - p1 * e0, e2);
        const double b0 = dsel(lin, e0, 0.0), b1 = dsel(lin, e1, 0.0), b2 = dsel(lin, e2, 0.0);
        double nv[6], Irow[6];
        double x = (a0 * pv[0] + a1 * pv[1] + a2 * pv[2]) + (b0 * pv[3] + b1 * pv[4] + b2 * pv[5]);
        const double qd = joint ? w.vh[6 + L.j0 + d - 1] : 0.0;
        if (r == 2) x += qd;
        if (act) S.vbuf[cur ^ 1][c8 + r] = x;
        ld6v(Isp + 36 * i + 6 * r, Irow);
        __syncwarp();
        ld6b(S.vbuf[cur ^ 1] + c8, nv);
        double xa = (a0 * pa[0] + a1 * pa[1] + a2 * pa[2]) + (b0 * pa[3] + b1 * pa[4] + b2 * pa[5]);   // zero at the first level
        // (v x S) qd with S = e_z: component r is sgn * v[partner], lane constants (rows 2 and 5: sgn = 0)
        const double colv = dsel(lin, dsel(r == 3, nv[4], nv[3]), dsel(r == 0, nv[1], nv[0]));
        xa += (csgn * colv) * qd;
        const double hb = Irow[0] * nv[0] + Irow[1] * nv[1] + Irow[2] * nv[2] + Irow[3] * nv[3] + Irow[4] * nv[4] + Irow[5] * nv[5];
        {   // branch-free conditional stores (a lane with nothing to store writes to the dump word)
            *(act ? S.hbuf + c8 + r : w.dump) = hb;
            *(act ? S.abuf[cur ^ 1] + c8 + r : w.dump) = xa;
            *((act && d == 7) ? S.accs + 6 * L.c + r : w.dump) = xa;    // sole frames (legs only): spatial acceleration for Jdot*qdot
        }
        __syncwarp();
        {
            double na[6], h[6];
            ld6b(S.abuf[cur ^ 1] + c8, na);
            ld6b(S.hbuf + c8, h);
            double f = Irow[0] * na[0] + Irow[1] * na[1] + Irow[2] * na[2] + Irow[3] * na[3] + Irow[4] * na[4] + Irow[5] * na[5];
#pragma unroll
            for (int k = 0; k < 6; ++k) { pv[k] = nv[k]; pa[k] = na[k]; }
            // v x* h = [w x n + u x f; w x f]: (w x s)[k3] with s = f (linear rows) or n, plus (u x f)[k3] on the angular rows
            const double sv[3] = {dsel(lin, h[3], h[0]), dsel(lin, h[4], h[1]), dsel(lin, h[5], h[2])};
            f += cross3_sel(nv, sv, k3) + dsel(lin, 0.0, cross3_sel(nv + 3, h + 3, k3));
            if (joint) S.frc[8 * i + r] = f;
        }
        cur ^= 1;
    }
    __syncwarp();
    // backward pass, deepest actuated level first: row r of X' f = [R n + p x (R f); R f].  A lane forms (R f)[k3];
    // the angular lanes fetch the other two components of R f from their chain by shuffle.
    const int k1 = k3 == 2 ? 0 : k3 + 1, k2 = k3 == 0 ? 2 : k3 - 1;
    const int src1 = 6 * L.c + k1, src2 = 6 * L.c + k2;
#pragma unroll 1
    for (int d = MAXDEPTH - 1; d >= 1; --d) {
        const bool on = L.on && d <= L.nact;
        const int i = L.first + (d <= L.nact ? d : L.nact) - 1;
        const double* X = Xl + 12 * i;
        double f[6];
        ld6b(S.frc + 8 * i, f);
        const double g0 = X[3 * k3], g1 = X[3 * k3 + 1], g2 = X[3 * k3 + 2];     // row k3 of R
        const double rf = g0 * f[3] + g1 * f[4] + g2 * f[5];
        const double rf1 = __shfl_sync(0xffffffffu, rf, src1), rf2 = __shfl_sync(0xffffffffu, rf, src2);
        const double o = dsel(lin, rf, g0 * f[0] + g1 * f[1] + g2 * f[2] + (X[9 + k1] * rf2 - X[9 + k2] * rf1));
        {
            double* pf = (on && d > 1) ? S.frc + 8 * (i - 1) + r : w.dump + 1;
            *pf = *pf + o;
            *((on && d == 1) ? S.contrib + c8 + r : w.dump) = o;
        }
        __syncwarp();
    }
    if (L.on && r < L.nact) w.Cg[6 + L.j0 + r] = S.frc[8 * (L.first + r) + 2];
    if (lane < 6) {
        double s = S.frc[lane];
        for (int c = NCH - 1; c >= 0; --c) s += S.contrib[LMH_XSTRIDE * c + lane];
        w.Cg[lane] = s;
    }
    if (lane >= 6 && lane < 18) {       // Jpqp = world-rotated sole accelerations
        const int it = lane - 6, s = it / 6, k = it % 6;
        const double* ai = S.accs + 6 * s + (k < 3 ? 0 : 3);
        w.Jpqp[it] = rot_row(w.Rs[s], k % 3, ai);
    }
    __syncwarp();
}

// ---- CRBA (Dynamics.cpp:62-101) -> M6 = [Ic0 | F2], Hc
// The composite inertia of a chain goes up one frame per level: Ic' = X'(I + Ic)X with X = [[E,0],[B,E]], B = -E p^.
// Six lanes hold it, and the levels ALTERNATE between two layouts so that the only exchange per level is a 6 x 6
// transpose through shared memory (six words out, six words in per lane; reading whole rows of the intermediate
// product for every output row cost six times the wavefronts):
//   even d: lane r holds ROW r     -> row r of Y = Ic X        -> transpose -> column r of X'Y  (a column of Ic')
//   odd  d: lane r holds COLUMN r  -> column r of Z = X'Ic     -> transpose -> row r of Z X     (a row of Ic')
// d = 1 is odd, so every chain arrives at the torso in rows.  Nothing is assumed symmetric (the inertia literals are not).
// v X for a row vector v = [v_ang, v_lin]: with u = v_lin E, v X = [v_ang E + v_lin B, u] and v_lin B = -(u p^) = p x u
// (B never has to be formed: eighteen registers less across the exchange)
__device__ __forceinline__ void crba_row_times_X(const double* Rp, const double* v, double* o) {
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {          // E[k][cc] = R[cc][k] = Rp[3 cc + k]
        const double e0 = Rp[3 * cc], e1 = Rp[3 * cc + 1], e2 = Rp[3 * cc + 2];
        o[3 + cc] = v[3] * e0 + v[4] * e1 + v[5] * e2;
        o[cc] = v[0] * e0 + v[1] * e1 + v[2] * e2;
    }
    o[0] += Rp[10] * o[5] - Rp[11] * o[4];
    o[1] += Rp[11] * o[3] - Rp[9] * o[5];
    o[2] += Rp[9] * o[4] - Rp[10] * o[3];
}
// X' v for a column v = [v_ang; v_lin]: with u = E' v_lin, X' v = [E' v_ang + B' v_lin; u] and B' v_lin = p x u
__device__ __forceinline__ void crba_XT_times_col(const double* Rp, const double* v, double* o) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {             // E[k][a] = Rp[3 a + k]
        const double e0 = Rp[3 * a], e1 = Rp[3 * a + 1], e2 = Rp[3 * a + 2];
        o[3 + a] = e0 * v[3] + e1 * v[4] + e2 * v[5];
        o[a] = e0 * v[0] + e1 * v[1] + e2 * v[2];
    }
    o[0] += Rp[10] * o[5] - Rp[11] * o[4];
    o[1] += Rp[11] * o[3] - Rp[9] * o[5];
    o[2] += Rp[9] * o[4] - Rp[10] * o[3];
}

template <bool ROWS>
__device__ __forceinline__ void crba_level(EnvWS& w, const double* Isp, const LaneRole& L, int d, double* icin, double* F) {
    auto& S = w.u.m.p.wr;
    const int c8 = LMH_XSTRIDE * L.c, r = L.r;
    const bool act = L.on && d <= L.nact;
    const int i = L.first + d - 1;
    double* T = S.Ybuf + 36 * L.c;
    double Rp[12];
    if (act) {
        ld12b(w.u.m.Xl + 12 * i, Rp);
        double ic[6], t[6];
        if (ROWS) {
            ld6v(Isp + 36 * i + 6 * r, ic);
#pragma unroll
            for (int k = 0; k < 6; ++k) ic[k] += icin[k];
            S.col2[c8 + r] = ic[2];                              // Ic S, S = e_z (angular): column 2, one entry per lane
            crba_row_times_X(Rp, ic, t);
#pragma unroll
            for (int k = 0; k < 6; ++k) T[6 * r + k] = t[k];
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) ic[k] = Isp[36 * i + 6 * k + r] + icin[k];
            if (r == 2) {
#pragma unroll
                for (int k = 0; k < 6; ++k) S.col2[c8 + k] = ic[k];   // the lane of column 2 has all of Ic S
            }
            crba_XT_times_col(Rp, ic, t);
#pragma unroll
            for (int k = 0; k < 6; ++k) T[6 * k + r] = t[k];
        }
        if (r == 2) w.Hc[36 * L.c + 7 * (d - 1)] = ic[2];
    }
    __syncwarp();
    if (act) {
        // this lane's joint: pick up the new vector at its own depth, then cross frame i towards the root
        if (r == d - 1) ld6b(S.col2 + c8, F);
        if (r >= d - 1 && r < L.nact) {
            double o[6];
            xf_all(Rp, F, o);
#pragma unroll
            for (int k = 0; k < 6; ++k) F[k] = o[k];
            if (d > 1) {
                w.Hc[36 * L.c + 6 * (d - 2) + r] = F[2];
                w.Hc[36 * L.c + 6 * r + (d - 2)] = F[2];
            } else {
#pragma unroll
                for (int k = 0; k < 6; ++k) w.M6[NQ * k + 6 + L.j0 + r] = F[k];
            }
        }
        double t[6];
        if (ROWS) {
#pragma unroll
            for (int k = 0; k < 6; ++k) t[k] = T[6 * k + r];      // column r of Y
            crba_XT_times_col(Rp, t, icin);                    // column r of Ic'
        } else {
            ld6v(T + 6 * r, t);                                   // row r of Z
            crba_row_times_X(Rp, t, icin);                     // row r of Ic'
            if (d == 1) {
#pragma unroll
                for (int k = 0; k < 6; ++k) S.Ybuf2[36 * L.c + 6 * r + k] = icin[k];
            }
        }
    }
    __syncwarp();
}

__device__ void crba_warp(EnvWS& w, const Consts& K, const double* Isp, const LaneRole& L, int lane) {
    auto& S = w.u.m.p.wr;
    double icin[6] = {0, 0, 0, 0, 0, 0};     // row / column r of the composite inertia arriving from the level below
    double F[6] = {0, 0, 0, 0, 0, 0};        // force vector of joint (c, r), expressed in the frame reached so far
    if (lane < 30) {
#pragma unroll
        for (int k = 0; k < 6; ++k) w.Hc[36 * L.c + 6 * L.r + k] = 0.0;
    }
#pragma unroll 1
    for (int d = MAXDEPTH - 1; d >= 2; d -= 2) {
        crba_level<true>(w, Isp, L, d, icin, F);
        crba_level<false>(w, Isp, L, d - 1, icin, F);
    }
    if (lane < 6) {       // Ic0 = I0 + sum over the five chains
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
            double s = Isp[6 * lane + cc];
            for (int c = NCH - 1; c >= 0; --c) s += S.Ybuf2[36 * c + 6 * lane + cc];
            w.M6[NQ * lane + cc] = s;
        }
    }
    __syncwarp();
}

// ---- sole Jacobians (invKinematics.cpp:72-149), one column per lane: lane = 12*foot + col, col 0..5 the base twist
// block (sole_X_base e_col), col 6+l the joint l of that leg (sole_X_(joint frame) S).  A column is a unit vector
// pushed through the motion transforms of the frames between it and the sole; no cross-lane traffic.
__device__ void jacobian_warp(EnvWS& w, int lane) {
    if (lane < 24) {
        const int s = lane / 12, col = lane % 12;
        const int f0 = s == 0 ? 1 : 8;
        const double* Xl = w.u.m.Xl;
        double v[6] = {0, 0, 0, 0, 0, 0};
        int start;
        if (col < 6) { v[0] = col == 0; v[1] = col == 1; v[2] = col == 2; v[3] = col == 3; v[4] = col == 4; v[5] = col == 5; start = 0; }
        else { v[2] = 1.0; start = col - 5; }
#pragma unroll 1
        for (int i = start; i <= 6; ++i) {
            double Rp[12];
            ld12b(Xl + 12 * (f0 + i), Rp);
            const double u0 = v[3] + (v[1] * Rp[11] - v[2] * Rp[10]);
            const double u1 = v[4] + (v[2] * Rp[9] - v[0] * Rp[11]);
            const double u2 = v[5] + (v[0] * Rp[10] - v[1] * Rp[9]);
            const double t0 = Rp[0] * v[0] + Rp[3] * v[1] + Rp[6] * v[2];
            const double t1 = Rp[1] * v[0] + Rp[4] * v[1] + Rp[7] * v[2];
            const double t2 = Rp[2] * v[0] + Rp[5] * v[1] + Rp[8] * v[2];
            v[3] = Rp[0] * u0 + Rp[3] * u1 + Rp[6] * u2;
            v[4] = Rp[1] * u0 + Rp[4] * u1 + Rp[7] * u2;
            v[5] = Rp[2] * u0 + Rp[5] * u1 + Rp[8] * u2;
            v[0] = t0; v[1] = t1; v[2] = t2;
        }
        const double* R = w.Rs[s];
        double* J = w.JFc + 72 * s + col;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            J[12 * k] = R[3 * k] * v[0] + R[3 * k + 1] * v[1] + R[3 * k + 2] * v[2];
            J[12 * (3 + k)] = R[3 * k] * v[3] + R[3 * k + 1] * v[4] + R[3 * k + 2] * v[5];
        }
    }
    __syncwarp();
}


// ---- 12 x 12 wrench-space system for one warp: lane i < 12 owns row i of A (w.u.s.Ay, lower triangle), lane 12 the
// right-hand-side row; left-looking Cholesky with the forward substitution riding along, then the back
// substitution with lane i holding zeta_i.  Result in row 12 of Ay, as the generic loops leave it.
__device__ void wrench_solve_warp(EnvWS& w, int lane) {
    double* Ay = w.u.s.Ay;
    double h[12];
    const bool row = lane < 13;
#pragma unroll
    for (int j = 0; j < 12; ++j) h[j] = (row && (j <= lane)) ? Ay[13 * lane + j] : 0.0;
    double myinv = 0.0;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        double s = h[k];
        const double* Lk = Ay + 13 * k;
#pragma unroll
        for (int m = 0; m < k; ++m) s -= h[m] * Lk[m];
        const double piv = __shfl_sync(0xffffffffu, s, k);
        const double rs = lmh_rsqrt(piv > 0.0 ? piv : 1e-300);
        const double lik = lane == k ? piv * rs : s * rs;
        h[k] = lik;
        if (lane == k) myinv = rs;
        if (lane > k && row) Ay[13 * lane + k] = lik;
        __syncwarp();
    }
    double zi = lane < 12 ? Ay[13 * 12 + lane] : 0.0;
#pragma unroll
    for (int k = 11; k >= 0; --k) {
        const double xk = __shfl_sync(0xffffffffu, zi * myinv, k);
        if (lane == k) zi = xk;
        else if (lane < k) zi -= Ay[13 * k + lane] * xk;
    }
    __syncwarp();
    if (lane < 12) Ay[13 * 12 + lane] = zi;
    __syncwarp();
}

// ---- QP front end for one warp (see assemble_phase / qp_phase of lmh_model.cuh for the algebra).  D1 = Wj + wFoot JF'JF
// couples the base with each leg but not the legs with each other, so the factorisation runs in the order
// [right leg, left leg, base]: position p <-> coordinate qp_orig(p).  The two 6 x 6 leg blocks are then independent
// (their columns k and 6+k are eliminated in the same step: 12 sequential pivots instead of 18, no fill between the
// legs) and only the six base rows are dense.  Lane p < 18 owns row p, lanes 18..27 own the ten right-hand-side rows
// [M6; ga'; U]; rows live in registers, the finished part of a pivot row is read back from shared memory as a
// broadcast.  Leaves L1, ladiag/ladinv and Yt (columns 0..17 in factor order) in shared memory; the Gram matrix of Yt
// does not depend on the order, and qp_backsolve_warp returns a in coordinate order.
__device__ __forceinline__ int qp_orig(int p) { return p < 12 ? p + 6 : p - 12; }

__device__ void qp_factor_warp(EnvWS& w, const Consts& K, int lane) {
    double* L1 = w.u.s.L1;
    double* Yt = w.u.s.Yt;
    double* Xbl = w.u.s.Wb;                           // [12][6] base-leg block, handed from the leg rows to the base rows
    double h[ND1];
    const bool mrow = lane < ND1;
    const int yr = lane - ND1;                       // right-hand-side row of this lane (0..9) if >= 0
    const bool yrow = yr >= 0 && yr < NYR;
    const bool isbase = lane >= 12 && lane < ND1;
    double* const dump = w.red + lane;
#pragma unroll
    for (int j = 0; j < ND1; ++j) h[j] = 0.0;
#if LMH_JTJ_DMMA
    {
        // J_f'J_f of both feet on the FP64 tensor cores (controller.cpp:94-123: the J'W_F J term of the Hessian).  J_f is
        // 6 x 12 in compact columns [base 0..5 | that leg's joints 6..11]; the 12 x 12 product is three 8 x 8 tiles of the
        // lower triangle (k = 6 -> two k-steps of m8n8k4), J_f serving as A (transposed on the fly) and as B.  The tiles
        // go to shared memory and every row owner then picks its row of D1 in factor order.
        double* Gs = w.u.s.Bm;                            // [2][12][12] (Bm | Pm | Wb | Ay are contiguous and free here)
        const int fr = lane >> 2, fk = lane & 3, c0 = 2 * fk;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const double* J = w.JFc + 72 * f;
            double t00a = 0, t00b = 0, t10a = 0, t10b = 0, t11a = 0, t11b = 0;
#pragma unroll
            for (int kt = 0; kt < 2; ++kt) {
                const int k = 4 * kt + fk;
                const double a0 = k < 6 ? J[12 * k + fr] : 0.0;
                const double a1 = (k < 6 && fr < 4) ? J[12 * k + 8 + fr] : 0.0;
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(t00a), "+d"(t00b) : "d"(a0), "d"(a0));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(t10a), "+d"(t10b) : "d"(a1), "d"(a0));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(t11a), "+d"(t11b) : "d"(a1), "d"(a1));
            }
            double* G = Gs + 144 * f;
            G[12 * fr + c0] = t00a; G[12 * fr + c0 + 1] = t00b;
            if (fr < 4) {
                G[12 * (8 + fr) + c0] = t10a; G[12 * (8 + fr) + c0 + 1] = t10b;
                G[12 * c0 + 8 + fr] = t10a; G[12 * (c0 + 1) + 8 + fr] = t10b;
                if (fk < 2) { G[12 * (8 + fr) + 8 + c0] = t11a; G[12 * (8 + fr) + 8 + c0 + 1] = t11b; }
            }
        }
        __syncwarp();
        if (mrow) {
            const double wF = K.wFoot;
            if (!isbase) {                                  // leg rows: the leg-leg block of that foot
                const int f = lane < 6 ? 0 : 1;
                const double* Gr = Gs + 144 * f + 12 * (6 + lane - 6 * f) + 6;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const double v = wF * Gr[c];
                    h[c] = dsel(lane < 6, v, 0.0);
                    h[6 + c] = v;                           // read by the left-leg rows only
                }
            } else {                                        // base rows: base-leg blocks of both feet, base-base block summed
                const int b = lane - 12;
                const double* G0 = Gs + 12 * b;
                const double* G1 = Gs + 144 + 12 * b;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    h[c] = wF * G0[6 + c];
                    h[6 + c] = wF * G1[6 + c];
                    h[12 + c] = wF * (G0[c] + G1[c]);
                }
            }
        }
    }
    if (!mrow && yrow) {
        // row yr of [M6; ga'; U]: a pointer and a scale per lane, no per-entry branching
        const double* src = yr < 6 ? w.M6 + NQ * yr : yr == 6 ? w.ga : w.AG + NQ * (yr - 4);
        const double sc = yr < 7 ? 1.0 : K.sq_wCoML;
        const double sc2 = sc * K.isq_wJoints;
#pragma unroll
        for (int j = 0; j < ND1; ++j) h[j] = sc * src[j < 12 ? j + 6 : j - 12];
#pragma unroll
        for (int j = ND1; j < NQ; ++j) Yt[NQ * yr + j] = sc2 * src[j];     // the arm/head part needs no solve
    }
    __syncwarp();
#else
    if (mrow) {
        // Only the lower triangle is ever read (row p: columns <= p).  Two passes of 6 columns x 6 terms per lane:
        // pass 0 against the base block of a foot (a leg row's result is its piece of the base-leg block, which the
        // base rows need: it travels through Xbl), pass 1 against the own leg block (leg rows) or the base block of
        // the other foot (base rows).
        const int fa = lane < 6 ? 0 : lane < 12 ? 1 : 0;
        const int ci = isbase ? lane - 12 : 6 + (lane - 6 * fa);     // compact column of this row in foot fa
        const double* Ja = w.JFc + 72 * fa;
        const double* Jb = isbase ? w.JFc + 72 : Ja + 6;
        const double* Jcb = isbase ? w.JFc + 72 : Ja;
        double jc[6], t0[6], t1[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) jc[k] = K.wFoot * Ja[12 * k + ci];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double sj = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) sj += jc[k] * Ja[12 * k + c];
            t0[c] = sj;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) jc[k] = K.wFoot * Jcb[12 * k + ci];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double sj = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) sj += jc[k] * Jb[12 * k + c];
            t1[c] = sj;
        }
        if (!isbase) {
#pragma unroll
            for (int c = 0; c < 6; ++c) Xbl[6 * lane + c] = t0[c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            h[c] = dsel(lane < 6, t1[c], 0.0);                  // right-leg rows; zero for the left-leg rows
            h[6 + c] = t1[c];                                   // read by the left-leg rows only (base rows: below)
            h[12 + c] = t0[c] + t1[c];                          // read by the base rows only
        }
    } else if (yrow) {
        // row yr of [M6; ga'; U]: a pointer and a scale per lane, no per-entry branching
        const double* src = yr < 6 ? w.M6 + NQ * yr : yr == 6 ? w.ga : w.AG + NQ * (yr - 4);
        const double sc = yr < 7 ? 1.0 : K.sq_wCoML;
        const double sc2 = sc * K.isq_wJoints;
#pragma unroll
        for (int j = 0; j < ND1; ++j) h[j] = sc * src[j < 12 ? j + 6 : j - 12];
#pragma unroll
        for (int j = ND1; j < NQ; ++j) Yt[NQ * yr + j] = sc2 * src[j];     // the arm/head part needs no solve
    }
    __syncwarp();
    if (isbase) {
        const int bcol = lane - 12;
#pragma unroll
        for (int c = 0; c < 12; ++c) h[c] = Xbl[6 * c + bcol];
    }
#endif
    if (mrow) {
        const int o = qp_orig(lane);
        const double wbp = K.wBasePos, wba = K.wBaseAng, wjn = K.wJoints;
        const double wj = o < 3 ? wbp : o < 6 ? wba : wjn;
#pragma unroll
        for (int j = 0; j < ND1; ++j) h[j] += dsel(j == lane, wj, 0.0);
    }
    // the two leg blocks, columns k and 6 + k together
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double sa = h[k], sb = h[6 + k];
        const double* La = L1 + LD1 * k;
        const double* Lb = L1 + LD1 * (6 + k) + 6;
#pragma unroll
        for (int m = 0; m < k; ++m) { sa -= h[m] * La[m]; sb -= h[6 + m] * Lb[m]; }
        const double piva = __shfl_sync(0xffffffffu, sa, k), pivb = __shfl_sync(0xffffffffu, sb, 6 + k);
        const double rsa = lmh_rsqrt(piva), rsb = lmh_rsqrt(pivb);
        const double lika = lane == k ? piva * rsa : sa * rsa;
        const double likb = lane == 6 + k ? pivb * rsb : sb * rsb;
        h[k] = lika; h[6 + k] = likb;
        {
            // branch-free stores: a lane with nothing to store writes to its own dump word (w.red is free here)
            double* pa = (mrow && lane > k) ? L1 + LD1 * lane + k : dump;
            double* pb = (mrow && lane > 6 + k) ? L1 + LD1 * lane + 6 + k : dump;
            *pa = lika; *pb = likb;
            const bool da = lane == k, db = lane == 6 + k;
            double* pd = da ? w.ladiag + k : db ? w.ladiag + 6 + k : dump;
            double* pi = da ? w.ladinv + k : db ? w.ladinv + 6 + k : dump;
            *pd = da ? lika : likb; *pi = da ? rsa : rsb;
        }
        __syncwarp();
    }
#if LMH_FACTOR_DMMA
    // Schur complement of the two leg blocks on the base block and on the right-hand-side rows, as one product on the FP64
    // tensor cores: rows 12..27 (six base rows, ten right-hand-side rows) times the base rows' leg part,
    // (16 x 12) (12 x 6): two 8 x 8 tiles, three k-steps each, instead of 72 broadcast loads and 72 FMAs per lane
    {
        if (yrow) {
#pragma unroll
            for (int j = 0; j < 12; ++j) Yt[NQ * yr + j] = h[j];          // (stored at the end otherwise)
        }
        __syncwarp();
        const int fr = lane >> 2, fk = lane & 3, c0 = 2 * fk;
        double d0a = 0, d0b = 0, d1a = 0, d1b = 0;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            const int m = 4 * kt + fk;
            const int p0 = 12 + fr, p1 = 20 + fr;                          // rows of the two tiles
            const double a0 = p0 < ND1 ? L1[LD1 * p0 + m] : Yt[NQ * (p0 - ND1) + m];
            const double a1 = Yt[NQ * (p1 - ND1) + m];
            const double b = fr < 6 ? L1[LD1 * (12 + fr) + m] : 0.0;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0a), "+d"(d0b) : "d"(a0), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d1a), "+d"(d1b) : "d"(a1), "d"(b));
        }
        double* Ts = w.u.s.Bm;                                             // [16][6] (free again: the rows have their J'J entries)
        if (c0 < 6) {
            Ts[6 * fr + c0] = d0a; Ts[6 * fr + c0 + 1] = d0b;
            Ts[6 * (8 + fr) + c0] = d1a; Ts[6 * (8 + fr) + c0 + 1] = d1b;
        }
        __syncwarp();
        if (lane >= 12 && lane < ND1 + NYR) {
#pragma unroll
            for (int c = 0; c < 6; ++c) h[12 + c] -= Ts[6 * (lane - 12) + c];
        }
    }
#endif
    // the base block
#pragma unroll
    for (int k = 12; k < ND1; ++k) {
        double s = h[k], t = 0.0;
        const double* Lk = L1 + LD1 * k;
#pragma unroll
        for (int m = LMH_FACTOR_DMMA ? 12 : 0; m < k; ++m) {
            const double l = Lk[m];
            if (m & 1) t -= h[m] * l; else s -= h[m] * l;
        }
        s += t;
        const double piv = __shfl_sync(0xffffffffu, s, k);
        const double rs = lmh_rsqrt(piv);
        const double lik = lane == k ? piv * rs : s * rs;
        h[k] = lik;
        {
            double* pl = (lane > k && mrow) ? L1 + LD1 * lane + k : dump;
            *pl = lik;
            const bool dk = lane == k;
            double* pd = dk ? w.ladiag + k : dump;
            double* pi = dk ? w.ladinv + k : dump;
            *pd = lik; *pi = rs;
        }
        __syncwarp();
    }
    if (yrow) {
#pragma unroll
        for (int j = LMH_FACTOR_DMMA ? 12 : 0; j < ND1; ++j) Yt[NQ * yr + j] = h[j];
    }
    __syncwarp();
}

// back substitution L1' a = b in factor order with lane p holding a_p (w.a[0:18] in: factor order; out: coordinate order)
__device__ void qp_backsolve_warp(EnvWS& w, int lane) {
    const double* L1 = w.u.s.L1;
    double ai = lane < ND1 ? w.a[lane] : 0.0;
    const double invd = lane < ND1 ? w.ladinv[lane] : 0.0;
#pragma unroll
    for (int k = ND1 - 1; k >= 12; --k) {            // base
        const double xk = __shfl_sync(0xffffffffu, ai * invd, k);
        if (lane == k) ai = xk;
        else if (lane < k) ai -= L1[LD1 * k + lane] * xk;
    }
    const bool legb = lane >= 6;
#pragma unroll
    for (int k = 5; k >= 0; --k) {                   // the legs do not see each other: pivots 6 + k and k together
        // branch-free (selects and an always-valid address): a divergent branch per step costs more than the step
        const double t = ai * invd;
        const double xb = __shfl_sync(0xffffffffu, t, 6 + k), xa = __shfl_sync(0xffffffffu, t, k);
        const int piv = legb ? 6 + k : k;
        const double x = legb ? xb : xa;
        const double l = L1[LD1 * piv + (lane < piv ? lane : piv)];
        const double upd = ai - l * x;
        ai = lane == piv ? x : (lane < piv ? upd : ai);
    }
    __syncwarp();
    if (lane < ND1) w.a[qp_orig(lane)] = ai;
    __syncwarp();
}

}  // namespace lmh
#endif
