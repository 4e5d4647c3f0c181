// lmh_mpc.cuh -- batched DENSE solve of the condensed LIPM-MPC QP on the device (Mpc3dLip::compute + solve1DQP,
// src/mpcLinearPendulum.cpp:78-133):
//     min 1/2 u'H u + g'u,   H = alpha I + beta Pu'Pu  (nV = N + 1, no constraints),   g = beta Pu'(Px x_k - z[k : k+N+1])
// for both axes of every problem.  H is the same for the whole batch, so its Cholesky factor L (H = L L') is computed
// once on the host (lmh_setup.hpp) and shared by every problem and by both axes; per problem the work is dense and
// GEMM-shaped once the right-hand sides of a tile of problems are put side by side:
//     E = Px x_k - z            (n x C, C = right-hand-side columns = 2 per problem)
//     G = (beta Pu') E          dense product with an upper-triangular n x n matrix
//     L Y = -G,  L'U = Y        block forward / back substitution, 8 rows at a time: the off-diagonal part of every block row is a
//                               product with 8 x 8 tiles of L, the diagonal block is applied as its precomputed inverse
// u0 = U[0], x+ = A x_k + B u0.  Two implementations of the same blocked algorithm, selectable per call:
//   * mpc_dmma_kernel: the tile products on the FP64 tensor cores (mma.sync.aligned.m8n8k4.f64), one warp per 8 columns;
//   * mpc_dfma_kernel: the FP64 FMA pipe, one lane per column, four rows in flight per lane.
// (The closed-loop kernel does not call this: for the unconstrained QP u0 is a fixed linear form of (x_k, z) -- the gain
// vector of lmh_setup.hpp -- and it uses that.  This is the general dense path: Mpc3dLip::compute of the class surface
// and the config-5 micro-benchmark.)
#pragma once
#if defined(__CUDACC__)
#include <stdint.h>

namespace lmh {

struct MpcDev {
    const double* PuT;     // [np][np] beta * Pu' (upper triangular), zero padded
    const double* L;       // [np][np] Cholesky factor of H (lower), padded rows/columns = identity
    const double* LT;      // [np][np] its transpose (upper)
    const double* Dinv;    // [np/8][64] inverse of the diagonal 8 x 8 blocks of L (lower triangular)
    const double* DinvT;   // [np/8][64] their transposes
    const double* Px;      // [np][2]
    const double* rdiag;   // [np] 1 / diag(L)
    double A[4], B[2];
    int32_t n, np;         // horizon + 1, padded to a multiple of 8
};

struct MpcProblemSet {
    const double* com_xy;  // [n_prob][2]
    const double* vel_xy;  // [n_prob][2]
    const double* zmp_x;   // [n_prob][n]
    const double* zmp_y;
    double* xref;          // [n_prob][3]
    double* yref;
    int32_t n_prob;
};

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// One warp per 8 right-hand-side columns (4 problems x 2 axes).  buf: [np][8] per warp in shared memory.
// Fragment layouts of m8n8k4: A[row = lane / 4][k = lane % 4], B[k = lane % 4][col = lane / 4], C/D[row = lane / 4][col = 2 (lane % 4) + {0, 1}].
__global__ void __launch_bounds__(128) mpc_dmma_kernel(MpcDev M, MpcProblemSet P, int repeats) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* buf = smem + (size_t)wid * (M.np * 8 + 64);
    double* tile = buf + M.np * 8;                       // [8][8] staging of one block row between the D and the B layout
    const int col0 = (blockIdx.x * 4 + wid) * 8;         // first of this warp's columns; column c = 2 * problem + axis
    const int np = M.np, n = M.n, nt = np / 8;
    const int ar = lane >> 2, ak = lane & 3;             // A fragment: row, k;  B fragment: k = ak, column = ar
    const int dc = 2 * ak;                                // D fragment: row ar, columns dc, dc + 1
    // this lane fills column lane % 8 of E, rows lane / 8 + 4 m
    const int ec = lane & 7, er = lane >> 3;
    const int eprob = (col0 + ec) >> 1, eax = (col0 + ec) & 1;
    const bool evalid = eprob < P.n_prob;
    const double pos = evalid ? P.com_xy[2 * eprob + eax] : 0.0, vel = evalid ? P.vel_xy[2 * eprob + eax] : 0.0;
    const double* z = (eax == 0 ? P.zmp_x : P.zmp_y) + (size_t)(evalid ? eprob : 0) * n;
    double u0 = 0.0;
    for (int rep = 0; rep < repeats; ++rep) {
        for (int i = er; i < np; i += 4)
            buf[8 * i + ec] = (i < n && evalid) ? (M.Px[2 * i] * pos + M.Px[2 * i + 1] * vel) - z[i] : 0.0;
        __syncwarp();
        // G = PuT E, in place (block row ti only needs rows >= 8 ti of E)
        for (int ti = 0; ti < nt; ++ti) {
            double d0 = 0.0, d1 = 0.0;
            for (int kt = 2 * ti; kt < np / 4; ++kt)
                dmma8x8x4(d0, d1, __ldg(M.PuT + (size_t)(8 * ti + ar) * np + 4 * kt + ak), buf[8 * (4 * kt + ak) + ar]);
            __syncwarp();                                  // every lane has read rows 8 ti .. before they are overwritten
            buf[8 * (8 * ti + ar) + dc] = d0; buf[8 * (8 * ti + ar) + dc + 1] = d1;
            __syncwarp();
        }
        // L Y = -G, block rows top down
        for (int ti = 0; ti < nt; ++ti) {
            double d0 = 0.0, d1 = 0.0;
            for (int kt = 0; kt < 2 * ti; ++kt)
                dmma8x8x4(d0, d1, __ldg(M.L + (size_t)(8 * ti + ar) * np + 4 * kt + ak), buf[8 * (4 * kt + ak) + ar]);
            tile[8 * ar + dc] = -buf[8 * (8 * ti + ar) + dc] - d0;
            tile[8 * ar + dc + 1] = -buf[8 * (8 * ti + ar) + dc + 1] - d1;
            __syncwarp();
            d0 = d1 = 0.0;
            for (int kt = 0; kt < 2; ++kt) dmma8x8x4(d0, d1, __ldg(M.Dinv + 64 * ti + 8 * ar + 4 * kt + ak), tile[8 * (4 * kt + ak) + ar]);
            buf[8 * (8 * ti + ar) + dc] = d0; buf[8 * (8 * ti + ar) + dc + 1] = d1;
            __syncwarp();
        }
        // L'U = Y, block rows bottom up
        for (int ti = nt - 1; ti >= 0; --ti) {
            double d0 = 0.0, d1 = 0.0;
            for (int kt = 2 * (ti + 1); kt < np / 4; ++kt)
                dmma8x8x4(d0, d1, __ldg(M.LT + (size_t)(8 * ti + ar) * np + 4 * kt + ak), buf[8 * (4 * kt + ak) + ar]);
            tile[8 * ar + dc] = buf[8 * (8 * ti + ar) + dc] - d0;
            tile[8 * ar + dc + 1] = buf[8 * (8 * ti + ar) + dc + 1] - d1;
            __syncwarp();
            d0 = d1 = 0.0;
            for (int kt = 0; kt < 2; ++kt) dmma8x8x4(d0, d1, __ldg(M.DinvT + 64 * ti + 8 * ar + 4 * kt + ak), tile[8 * (4 * kt + ak) + ar]);
            buf[8 * (8 * ti + ar) + dc] = d0; buf[8 * (8 * ti + ar) + dc + 1] = d1;
            __syncwarp();
        }
        u0 = buf[ec];                                      // row 0, column ec
        __syncwarp();
    }
    if (lane < 8 && evalid) {                              // x+ = A x_k + B u0 (mpcLinearPendulum.cpp:104-108)
        double* o = (eax == 0 ? P.xref : P.yref) + 3 * (size_t)eprob;
        o[0] = (M.A[0] * pos + M.A[1] * vel) + M.B[0] * u0;
        o[1] = (M.A[2] * pos + M.A[3] * vel) + M.B[1] * u0;
        o[2] = u0;
    }
}

// One lane per right-hand-side column, 32 columns (16 problems) per warp; buf: [np][32] per warp (conflict free).
// Four rows in flight per lane: every y[k] read from shared memory feeds four FMAs, the matrix entries are warp-uniform loads.
__global__ void __launch_bounds__(64) mpc_dfma_kernel(MpcDev M, MpcProblemSet P, int repeats) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* buf = smem + (size_t)wid * M.np * 32;
    const int col = (blockIdx.x * 2 + wid) * 32 + lane;
    const int np = M.np, n = M.n;
    const int prob = col >> 1, ax = col & 1;
    const bool valid = prob < P.n_prob;
    const double pos = valid ? P.com_xy[2 * prob + ax] : 0.0, vel = valid ? P.vel_xy[2 * prob + ax] : 0.0;
    const double* z = (ax == 0 ? P.zmp_x : P.zmp_y) + (size_t)(valid ? prob : 0) * n;
    double u0 = 0.0;
#define Y(i) buf[32 * (i) + lane]
    for (int rep = 0; rep < repeats; ++rep) {
        for (int i = 0; i < np; ++i) Y(i) = (i < n && valid) ? (M.Px[2 * i] * pos + M.Px[2 * i + 1] * vel) - z[i] : 0.0;
        // G = PuT E in place, rows ascending (row i needs E[k], k >= i)
        for (int i = 0; i < np; i += 4) {
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            const double* r0 = M.PuT + (size_t)i * np;
            for (int k = i; k < np; ++k) {
                const double e = Y(k);
                s0 = fma(__ldg(r0 + k), e, s0); s1 = fma(__ldg(r0 + np + k), e, s1);
                s2 = fma(__ldg(r0 + 2 * np + k), e, s2); s3 = fma(__ldg(r0 + 3 * np + k), e, s3);
            }
            Y(i) = s0; Y(i + 1) = s1; Y(i + 2) = s2; Y(i + 3) = s3;
        }
        // L Y = -G, four rows at a time: the part left of the 4 x 4 diagonal block, then the block itself
        for (int i = 0; i < np; i += 4) {
            double s0 = -Y(i), s1 = -Y(i + 1), s2 = -Y(i + 2), s3 = -Y(i + 3);
            const double* r0 = M.L + (size_t)i * np;
            for (int k = 0; k < i; ++k) {
                const double y = Y(k);
                s0 = fma(-__ldg(r0 + k), y, s0); s1 = fma(-__ldg(r0 + np + k), y, s1);
                s2 = fma(-__ldg(r0 + 2 * np + k), y, s2); s3 = fma(-__ldg(r0 + 3 * np + k), y, s3);
            }
            s0 = s0 * __ldg(M.rdiag + i);
            s1 = (s1 - __ldg(r0 + np + i) * s0) * __ldg(M.rdiag + i + 1);
            s2 = (s2 - __ldg(r0 + 2 * np + i) * s0 - __ldg(r0 + 2 * np + i + 1) * s1) * __ldg(M.rdiag + i + 2);
            s3 = (s3 - __ldg(r0 + 3 * np + i) * s0 - __ldg(r0 + 3 * np + i + 1) * s1 - __ldg(r0 + 3 * np + i + 2) * s2) * __ldg(M.rdiag + i + 3);
            Y(i) = s0; Y(i + 1) = s1; Y(i + 2) = s2; Y(i + 3) = s3;
        }
        // L'U = Y, bottom up (LT is upper triangular, row i needs U[k], k > i)
        for (int i = np - 4; i >= 0; i -= 4) {
            double s0 = Y(i), s1 = Y(i + 1), s2 = Y(i + 2), s3 = Y(i + 3);
            const double* r0 = M.LT + (size_t)i * np;
            for (int k = i + 4; k < np; ++k) {
                const double y = Y(k);
                s0 = fma(-__ldg(r0 + k), y, s0); s1 = fma(-__ldg(r0 + np + k), y, s1);
                s2 = fma(-__ldg(r0 + 2 * np + k), y, s2); s3 = fma(-__ldg(r0 + 3 * np + k), y, s3);
            }
            s3 = s3 * __ldg(M.rdiag + i + 3);
            s2 = (s2 - __ldg(r0 + 2 * np + i + 3) * s3) * __ldg(M.rdiag + i + 2);
            s1 = (s1 - __ldg(r0 + np + i + 2) * s2 - __ldg(r0 + np + i + 3) * s3) * __ldg(M.rdiag + i + 1);
            s0 = (s0 - __ldg(r0 + i + 1) * s1 - __ldg(r0 + i + 2) * s2 - __ldg(r0 + i + 3) * s3) * __ldg(M.rdiag + i);
            Y(i) = s0; Y(i + 1) = s1; Y(i + 2) = s2; Y(i + 3) = s3;
        }
        u0 = Y(0);
    }
#undef Y
    if (valid) {
        double* o = (ax == 0 ? P.xref : P.yref) + 3 * (size_t)prob;
        o[0] = (M.A[0] * pos + M.A[1] * vel) + M.B[0] * u0;
        o[1] = (M.A[2] * pos + M.A[3] * vel) + M.B[1] * u0;
        o[2] = u0;
    }
}

}  // namespace lmh
#endif
