// lmh_ik.cuh -- batched Newton inverse kinematics of the initial posture, one warp per environment.
//
// Follows Kinematics::compute / jacInvKinematics / comJacobian / operationalState / rotMatrixToEulerAngles
// (src/invKinematics.cpp:11-70,151-267) exactly as the host setup code does (lmh_setup.hpp,
// solve_initial_posture), with one algebraic shortcut: twelve rows of the 30 x 30 Newton system are identity rows
// on the arm/head joints and three are identity rows on the base attitude, so those increments are the residuals
// themselves and what is left is a 15 x 15 system in [base position, right-leg joints, left-leg joints], solved by
// Gauss-Jordan elimination with partial pivoting (one row per lane, the pivot row travels by shuffle).
// Operational state: [right sole position, right sole Euler angles, left sole ..., arm+head joints (12),
// base attitude (3), CoM (3)].
#pragma once
#if defined(__CUDACC__)

namespace lmh {

// rotMatrixToEulerAngles(R * ref) (invKinematics.cpp:257-267)
__device__ __forceinline__ void ik_euler_from(const double* R, const double* ref, double* eta) {
    double N[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) N[3 * r + c] = R[3 * r] * ref[c] + R[3 * r + 1] * ref[3 + c] + R[3 * r + 2] * ref[6 + c];
    eta[2] = atan2(N[3], N[0]);
    double s2, c2;
    sincos(eta[2], &s2, &c2);
    eta[1] = atan2(-N[6], c2 * N[0] + s2 * N[3]);
    eta[0] = atan2(s2 * N[2] - c2 * N[5], -s2 * N[1] + c2 * N[4]);
}
// matrixAngularVelToEulerDot (generalizedFunctions.cpp:43-50)
__device__ __forceinline__ void ik_omega_map(const double* eta, double* O) {
    double s2, c2, s1, c1;
    sincos(eta[2], &s2, &c2);
    sincos(eta[1], &s1, &c1);
    const double t1 = s1 / c1;
    O[0] = c2 / c1; O[1] = s2 / c1; O[2] = 0; O[3] = -s2; O[4] = c2; O[5] = 0; O[6] = c2 * t1; O[7] = s2 * t1; O[8] = 1;
}
__device__ __forceinline__ void ik_inv3(const double* m, double* o) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[2] * m[7] - m[1] * m[8], c02 = m[1] * m[5] - m[2] * m[4];
    const double c10 = m[5] * m[6] - m[3] * m[8], c11 = m[0] * m[8] - m[2] * m[6], c12 = m[2] * m[3] - m[0] * m[5];
    const double c20 = m[3] * m[7] - m[4] * m[6], c21 = m[1] * m[6] - m[0] * m[7], c22 = m[0] * m[4] - m[1] * m[3];
    const double det = m[0] * c00 + m[1] * c10 + m[2] * c20;
    o[0] = c00 / det; o[1] = c01 / det; o[2] = c02 / det; o[3] = c10 / det; o[4] = c11 / det; o[5] = c12 / det;
    o[6] = c20 / det; o[7] = c21 / det; o[8] = c22 / det;
}

// reduced unknown k (0..14) -> generalized coordinate: base position, right leg, left leg
__device__ __forceinline__ int ik_unknown_coord(int k) { return k < 3 ? k : k + 3; }

struct IkResult { int iters; double residual; int status; };     // status 0 converged, 1 iteration cap, 2 singular

// One environment, one warp.  w.q holds the starting posture on entry and the solution on exit; w.com, w.Rs, w.ps
// describe the solution as the model phases leave them.  des: operational target (30), in shared memory.
__device__ IkResult ik_solve_warp(EnvWS& w, const Consts& K, const double* des, int max_iter, int lane) {
    Ctx<32> cx{lane, 0};
    double* Jc = w.M6;          // [3][30] CoM Jacobian (raw attitude columns)            -- M6 is free during the IK
    double* pc = w.Hc;          // [27][4] world CoM and mass of every link
    double* sc = w.AG;          // scratch: etaF[2][3] at 0, OmF[2][9] at 8, Omi[9] at 32, e[30] at 48, x[15] at 80
    IkResult res{0, 0.0, 1};
    for (;;) {
        fk_phase(cx, w, K);                 // Tw (world), Xl (parent-relative), Rs, ps, com
        jacobian_warp(w, lane);             // JFc; leaves Tw alone
        if (lane < 2) {
            double eta[3], O[9];
            ik_euler_from(w.Rs[lane], K.Rdes, eta);
            ik_omega_map(eta, O);
#pragma unroll
            for (int k = 0; k < 3; ++k) sc[3 * lane + k] = eta[k];
#pragma unroll
            for (int k = 0; k < 9; ++k) sc[8 + 9 * lane + k] = O[k];
        } else if (lane == 2) {
            double O[9], Oi[9];
            ik_omega_map(w.q + 3, O);
            ik_inv3(O, Oi);
#pragma unroll
            for (int k = 0; k < 9; ++k) sc[32 + k] = Oi[k];
        }
        if (lane < NF) {                    // world CoM of link `lane`
            const double* Ti = w.u.m.p.Tw + 12 * lane;
            const double* mc = K.mc + 4 * lane;
#pragma unroll
            for (int k = 0; k < 3; ++k) pc[4 * lane + k] = Ti[3 * k] * mc[0] + Ti[3 * k + 1] * mc[1] + Ti[3 * k + 2] * mc[2] + Ti[9 + k];
            pc[4 * lane + 3] = mc[3];
        }
        __syncwarp();
        // operational state (invKinematics.cpp:54-70) and residual
        double ei = 0.0;
        if (lane < NQ) {
            double cur;
            if (lane < 12) { const int s = lane / 6, k = lane % 6; cur = k < 3 ? w.ps[s][k] : sc[3 * s + k - 3]; }
            else if (lane < 24) cur = w.q[18 + lane - 12];
            else if (lane < 27) cur = w.q[3 + lane - 24];
            else cur = w.com[lane - 27];
            ei = des[lane] - cur;
            sc[48 + lane] = ei;
        }
        const double emax = warp_max(fabs(ei));
        res.residual = emax;
        if (!(emax > 1e-10)) { res.status = 0; break; }
        if (res.iters >= max_iter) { res.status = 1; break; }
        // CoM Jacobian (invKinematics.cpp:208-245): lane a < 24 the column of joint a, lanes 24..26 the raw attitude columns
        {
            const double* Tw = w.u.m.p.Tw;
            double j0 = 0.0, j1 = 0.0, j2 = 0.0;
            if (lane < NJ) {
                const int c = joint_chain(lane);
                const int fj = chain_first(c) + (lane - chain_joint0(c));      // frame of this joint
                const int last = chain_first(c) + chain_len(c) - 1;
                const double* Tj = Tw + 12 * fj;
                const double z0 = Tj[2], z1 = Tj[5], z2 = Tj[8];
                for (int i = fj; i <= last; ++i) {                             // links carried by the joint
                    const double m = pc[4 * i + 3];
                    const double d0 = pc[4 * i] - Tj[9], d1 = pc[4 * i + 1] - Tj[10], d2 = pc[4 * i + 2] - Tj[11];
                    j0 += m * (z1 * d2 - z2 * d1); j1 += m * (z2 * d0 - z0 * d2); j2 += m * (z0 * d1 - z1 * d0);
                }
            } else if (lane < 27) {
                const int cc = lane - 24;                                      // column cc of sum_i m_i crossMatrix(p_base - pc_i)
                for (int i = 0; i < NF; ++i) {
                    const double m = pc[4 * i + 3];
                    const double d0 = Tw[9] - pc[4 * i], d1 = Tw[10] - pc[4 * i + 1], d2 = Tw[11] - pc[4 * i + 2];
                    const double v0 = cc == 0 ? 0.0 : cc == 1 ? -d2 : d1;
                    const double v1 = cc == 0 ? d2 : cc == 1 ? 0.0 : -d0;
                    const double v2 = cc == 0 ? -d1 : cc == 1 ? d0 : 0.0;
                    j0 += m * v0; j1 += m * v1; j2 += m * v2;
                }
            }
            if (lane < 27) {
                const int col = lane < NJ ? 6 + lane : 3 + (lane - 24);
                Jc[col] = j0 * K.inv_mass; Jc[NQ + col] = j1 * K.inv_mass; Jc[2 * NQ + col] = j2 * K.inv_mass;
            }
        }
        __syncwarp();
        // reduced system: lane rho < 15 owns row rho = [A(15) | b]
        double A[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) A[k] = 0.0;
        {
            const double* Omi = sc + 32;
            const double* e = sc + 48;
            const double da0 = e[24], da1 = e[25], da2 = e[26];                // attitude increment = its residual
            if (lane < 12) {
                const int s = lane / 6, half = (lane % 6) / 3, r = lane % 3;
                const int jr = half == 0 ? 3 + r : r;                          // row of the sole Jacobian ([ang; lin])
                const double* Jrow = w.JFc + 72 * s + 12 * jr;                 // compact: 0..5 base twist [ang | lin], 6..11 leg joints
#pragma unroll
                for (int c = 0; c < 3; ++c) A[c] = Jrow[3 + c];                // base position columns = linear twist columns
#pragma unroll
                for (int l = 0; l < 6; ++l) A[3 + 6 * s + l] = Jrow[6 + l];
                double att[3];
                if (half == 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) att[c] = Jrow[0] * Omi[c] + Jrow[1] * Omi[3 + c] + Jrow[2] * Omi[6 + c];
                } else {                    // angular rows: the attitude block is premultiplied by the foot's own Omega
                    const double* OmF = sc + 8 + 9 * s;
                    att[0] = att[1] = att[2] = 0.0;
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) {
                        const double* Jr2 = w.JFc + 72 * s + 12 * rr;
                        const double o = OmF[3 * r + rr];
#pragma unroll
                        for (int c = 0; c < 3; ++c) att[c] += o * (Jr2[0] * Omi[c] + Jr2[1] * Omi[3 + c] + Jr2[2] * Omi[6 + c]);
                    }
                }
                A[15] = e[lane] - (att[0] * da0 + att[1] * da1 + att[2] * da2);
            } else if (lane < 15) {
                const int r = lane - 12;
                const double* Jr = Jc + NQ * r;
                A[r] = 1.0;
#pragma unroll
                for (int l = 0; l < 12; ++l) A[3 + l] = Jr[6 + l];
                double att[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) att[c] = Jr[3] * Omi[c] + Jr[4] * Omi[3 + c] + Jr[5] * Omi[6 + c];
                double b = e[27 + r] - (att[0] * da0 + att[1] * da1 + att[2] * da2);
#pragma unroll
                for (int k = 0; k < 12; ++k) b -= Jr[18 + k] * e[12 + k];       // arm/head increments = their residuals
                A[15] = b;
            }
        }
        // Gauss-Jordan with partial pivoting over the rows not used yet
        bool used = lane >= 15;
        int mycol = -1;
        bool singular = false;
#pragma unroll
        for (int k = 0; k < 15; ++k) {
            const double mag = used ? -1.0 : fabs(A[k]);
            const double best = warp_max(mag);
            const unsigned hit = __ballot_sync(0xffffffffu, !used && mag == best);
            if (!(best > 1e-300) || hit == 0u) { singular = true; break; }
            const int p = __ffs(hit) - 1;
            const double piv = __shfl_sync(0xffffffffu, A[k], p);
            const double fac = lane == p ? 0.0 : A[k] / piv;
#pragma unroll
            for (int j = k + 1; j < 16; ++j) {
                const double pj = __shfl_sync(0xffffffffu, A[j], p);
                A[j] -= fac * pj;
            }
            if (lane != p) A[k] = 0.0;
            if (lane == p) { used = true; mycol = k; }
        }
        if (singular) { res.status = 2; break; }
        if (lane < 15) {
#pragma unroll
            for (int k = 0; k < 15; ++k) if (k == mycol) w.q[ik_unknown_coord(k)] += A[15] / A[k];
        } else if (lane < 27) {
            w.q[18 + lane - 15] += sc[48 + 12 + lane - 15];
        } else if (lane < 30) {
            w.q[3 + lane - 27] += sc[48 + 24 + lane - 27];
        }
        __syncwarp();
        ++res.iters;
    }
    __syncwarp();
    return res;
}

}  // namespace lmh
#endif
