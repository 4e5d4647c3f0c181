// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// Dense strictly-convex QP solver standing in for qpOASES (third-party, NOT
// vendored under /root/reference, version unpinned: CMakeLists.txt:39-41,76).
// Reference call sites: src/controller.cpp:467-469 (SQProblem::init, cold
// start, nWSR=500) and src/mpcLinearPendulum.cpp:120-127 (QProblem, nC=0).
//
//     min 1/2 x'Hx + g'x   s.t.  lbA <= A x <= ubA
//
// PARITY UNPINNED: the reference ships no test or golden vector at this
// boundary and qpOASES itself is absent, so this solver is certified by its
// own KKT residuals (orc::qp_kkt_residuals) instead.  Both reference QPs are
// strictly convex (unique minimiser), so any exact method must agree with
// qpOASES up to the latter's termination tolerance.
//
// Method: Goldfarb & Idnani (1983) dual active-set in its numerically stable
// form (qp_solve_gi): J = L^-T Q and the triangular factor R of the active
// normals are carried along and updated with Givens rotations when a constraint
// enters or leaves, so the conditioning that matters is that of L^-1 N (about
// 1e4..1e8 here because of the 1e-8 weight on the contact coefficients), not its
// square.  The earlier "no-update" variant that re-factorises the Gram matrix
// A H^-1 A' every iteration is kept as qp_solve_gi_noupdate: it loses up to 1e-3
// relative accuracy on WBC problems with many active bounds (measured against
// 60-digit solutions, tests/golden/make_qp_exact.py) and is only there for timing
// comparisons.  Both are templated on the scalar so tests can run them in long
// double to measure their own round-off.
#pragma once
#include <cmath>
#include <vector>
#include <limits>

namespace orc {

constexpr double QP_INFTY = 1.0e20;  // qpOASES::INFTY

struct QpInfo {
    int status = 0;      // 0 ok, 1 iteration limit, 2 infeasible, 3 not PD
    int iterations = 0;  // active-set changes (adds + drops)
    int n_active = 0;
};

template <typename T>
static bool chol_factor(int n, std::vector<T>& L) {  // in place, lower, row-major n x n
    for (int j = 0; j < n; ++j) {
        T s = L[j * n + j];
        for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
        if (!(s > T(0))) return false;
        T d = std::sqrt(s);
        L[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            T t = L[i * n + j];
            for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = t / d;
        }
    }
    return true;
}
template <typename T>
static void chol_solve(int n, const std::vector<T>& L, T* b) {
    for (int i = 0; i < n; ++i) {
        T s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
        b[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        T s = b[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
        b[i] = s / L[i * n + i];
    }
}

// H: n x n row-major (symmetric PD), A: m x n row-major.
// x: n primal out; y: m multipliers out (>=0 at a lower bound, <=0 at an upper
// bound, free sign on equalities).
template <typename T>
QpInfo qp_solve_gi_noupdate(int n, const double* H, const double* g, int m, const double* A,
                   const double* lbA, const double* ubA, double* x_out, double* y_out,
                   int max_iter = 500) {
    QpInfo info;
    std::vector<T> L(n * n);
    for (int i = 0; i < n * n; ++i) L[i] = T(H[i]);
    if (!chol_factor<T>(n, L)) {
        info.status = 3;
        return info;
    }
    std::vector<T> x(n);
    for (int i = 0; i < n; ++i) x[i] = -T(g[i]);
    chol_solve<T>(n, L, x.data());

    // W[:,i] = H^-1 a_i (stored row i = constraint i), Gram = A H^-1 A'
    std::vector<T> W(m * n), Gram(m * m);
    for (int i = 0; i < m; ++i) {
        for (int k = 0; k < n; ++k) W[i * n + k] = T(A[i * n + k]);
        chol_solve<T>(n, L, &W[i * n]);
    }
    for (int i = 0; i < m; ++i)
        for (int j = 0; j <= i; ++j) {
            T s = 0;
            for (int k = 0; k < n; ++k) s += T(A[i * n + k]) * W[j * n + k];
            Gram[i * m + j] = s;
            Gram[j * m + i] = s;
        }

    // sign[i] = +1: a_i x >= lb_i active form; -1: -a_i x >= -ub_i
    std::vector<int> act;       // active constraint ids
    std::vector<int> sgn;       // their signs
    std::vector<char> is_eq(m, 0), in_act(m, 0);
    std::vector<T> u;           // multipliers of the active set (signed form)
    for (int i = 0; i < m; ++i) is_eq[i] = (lbA[i] == ubA[i]);

    auto residual = [&](int i) {  // a_i x
        T s = 0;
        for (int k = 0; k < n; ++k) s += T(A[i * n + k]) * x[k];
        return s;
    };

    const T tiny = std::numeric_limits<T>::epsilon();
    std::vector<T> r, Gs, rhs, z(n);
    int n_eq_done = 0;
    while (true) {
        // ---- choose the constraint p to add (equalities first, in order) ----
        int p = -1, ps = 1;
        T viol = 0;
        bool eq_phase = false;
        for (int i = 0; i < m && p < 0; ++i)
            if (is_eq[i] && !in_act[i]) {
                T s = residual(i) - T(lbA[i]);
                p = i;
                ps = (s > 0) ? -1 : 1;
                viol = (s > 0) ? -s : s;
                eq_phase = true;
            }
        if (p < 0) {
            T worst = 0;
            for (int i = 0; i < m; ++i) {
                if (is_eq[i] || in_act[i]) continue;
                T ax = residual(i);
                T scale = T(1);
                if (lbA[i] > -QP_INFTY) {
                    T s = ax - T(lbA[i]);
                    if (s < worst * scale) { worst = s; p = i; ps = 1; }
                }
                if (ubA[i] < QP_INFTY) {
                    T s = T(ubA[i]) - ax;
                    if (s < worst * scale) { worst = s; p = i; ps = -1; }
                }
            }
            // feasibility tolerance relative to the data scale
            T tol = T(1e3) * tiny;
            if (p >= 0) {
                T mag = std::fabs(residual(p)) + T(1);
                if (-worst <= tol * mag) p = -1;
            }
            viol = worst;
        }
        if (p < 0) break;  // optimal
        (void)n_eq_done;

        T up = 0;  // multiplier of p while it is being added
        bool added = false;
        while (!added) {
            if (++info.iterations > max_iter) {
                info.status = 1;
                goto done;
            }
            int q = (int)act.size();
            // r = (N' H^-1 N)^-1 N' H^-1 n_p   (signed constraints)
            r.assign(q, 0);
            if (q > 0) {
                Gs.assign(q * q, 0);
                for (int a = 0; a < q; ++a) {
                    for (int b = 0; b < q; ++b)
                        Gs[a * q + b] = T(sgn[a] * sgn[b]) * Gram[act[a] * m + act[b]];
                    r[a] = T(sgn[a] * ps) * Gram[act[a] * m + p];
                }
                if (!chol_factor<T>(q, Gs)) {
                    info.status = 2;
                    goto done;
                }
                chol_solve<T>(q, Gs, r.data());
            }
            // z = H^-1 (n_p - N r)
            T znorm = 0, zn = 0;  // zn = z' n_p
            for (int k = 0; k < n; ++k) {
                T s = T(ps) * W[p * n + k];
                for (int a = 0; a < q; ++a) s -= T(sgn[a]) * W[act[a] * n + k] * r[a];
                z[k] = s;
            }
            for (int k = 0; k < n; ++k) {
                znorm += z[k] * z[k];
                zn += z[k] * T(ps) * T(A[p * n + k]);
            }
            // dual step length (only inequality multipliers are sign-restricted)
            T t1 = std::numeric_limits<T>::infinity();
            int drop = -1;
            for (int a = 0; a < q; ++a) {
                if (is_eq[act[a]]) continue;
                if (r[a] > 0) {
                    T t = u[a] / r[a];
                    if (t < t1) { t1 = t; drop = a; }
                }
            }
            // primal step length
            T sp = T(ps) * residual(p) - T(ps) * T(ps > 0 ? lbA[p] : ubA[p]);  // s_p(x) < 0
            T gram_pp = Gram[p * m + p];
            bool z_zero = !(zn > T(1e4) * tiny * gram_pp);
            T t2 = z_zero ? std::numeric_limits<T>::infinity() : -sp / zn;
            if (!(sp < 0)) t2 = 0;  // already satisfied by earlier partial steps
            T t = (t1 < t2) ? t1 : t2;
            if (t == std::numeric_limits<T>::infinity()) {
                info.status = 2;  // infeasible
                goto done;
            }
            if (!z_zero || t2 == 0) {
                for (int k = 0; k < n; ++k) x[k] += t * z[k];
            }
            for (int a = 0; a < q; ++a) u[a] -= t * r[a];
            up += t;
            if (t == t2) {  // full step: p becomes active
                act.push_back(p);
                sgn.push_back(ps);
                u.push_back(up);
                in_act[p] = 1;
                added = true;
            } else {  // partial step: drop the blocking constraint and retry
                in_act[act[drop]] = 0;
                act.erase(act.begin() + drop);
                sgn.erase(sgn.begin() + drop);
                u.erase(u.begin() + drop);
            }
        }
        (void)eq_phase;
        (void)viol;
    }
done:
    for (int k = 0; k < n; ++k) x_out[k] = (double)x[k];
    if (y_out) {
        for (int i = 0; i < m; ++i) y_out[i] = 0.0;
        for (size_t a = 0; a < act.size(); ++a) y_out[act[a]] = (double)(T(sgn[a]) * u[a]);
    }
    info.n_active = (int)act.size();
    return info;
}

// ---------------------------------------------------------------------------
// Goldfarb-Idnani with orthogonal updates.
//   J (n x n): J = L^-T Q;  R (upper triangular, q x q): Q1' L^-1 N = R
//   where N holds the signed normals of the active constraints as columns.
template <typename T>
QpInfo qp_solve_gi(int n, const double* H, const double* g, int m, const double* A,
                   const double* lbA, const double* ubA, double* x_out, double* y_out,
                   int max_iter = 500) {
    QpInfo info;
    const T tiny = std::numeric_limits<T>::epsilon();
    std::vector<T> L(n * n);
    for (int i = 0; i < n * n; ++i) L[i] = T(H[i]);
    if (!chol_factor<T>(n, L)) { info.status = 3; return info; }
    // J = L^-T  (column k of J solves L' J[:,k] = e_k)
    std::vector<T> J(n * n, T(0)), R(n * n, T(0));
    for (int k = 0; k < n; ++k) {
        for (int i = n - 1; i >= 0; --i) {
            T s = (i == k) ? T(1) : T(0);
            for (int r = i + 1; r < n; ++r) s -= L[r * n + i] * J[r * n + k];
            J[i * n + k] = s / L[i * n + i];
        }
    }
    T Rnorm = T(1);
    std::vector<T> x(n);
    for (int i = 0; i < n; ++i) x[i] = -T(g[i]);
    chol_solve<T>(n, L, x.data());

    std::vector<int> act, sgn;          // active ids and signs, in order of R's columns
    std::vector<T> u;                   // their multipliers (>= 0 for inequalities in signed form)
    std::vector<char> is_eq(m, 0), in_act(m, 0);
    for (int i = 0; i < m; ++i) is_eq[i] = (lbA[i] == ubA[i]);
    std::vector<T> d(n), z(n), r(n), np_(n);
    auto ax = [&](int i) { T s = 0; for (int k = 0; k < n; ++k) s += T(A[i * n + k]) * x[k]; return s; };

    auto add_constraint = [&](int& iq) -> bool {   // d = J' n+ is current; rotate J so that d[iq+1..] = 0
        for (int j = n - 1; j >= iq + 1; --j) {
            T cc = d[j - 1], ss = d[j];
            T h = std::sqrt(cc * cc + ss * ss);
            if (h == T(0)) continue;
            d[j] = T(0);
            ss = ss / h; cc = cc / h;
            if (cc < T(0)) { cc = -cc; ss = -ss; d[j - 1] = -h; } else d[j - 1] = h;
            T xny = ss / (T(1) + cc);
            for (int k = 0; k < n; ++k) {
                T t1 = J[k * n + j - 1], t2 = J[k * n + j];
                J[k * n + j - 1] = t1 * cc + t2 * ss;
                J[k * n + j] = xny * (t1 + J[k * n + j - 1]) - t2;
            }
        }
        ++iq;
        for (int i = 0; i < iq; ++i) R[i * n + iq - 1] = d[i];
        if (std::fabs(d[iq - 1]) <= tiny * Rnorm) return false;   // linearly dependent
        Rnorm = std::max(Rnorm, std::fabs(d[iq - 1]));
        return true;
    };
    auto delete_constraint = [&](int& iq, int pos) {
        for (int i = pos; i < iq - 1; ++i) {
            act[i] = act[i + 1]; sgn[i] = sgn[i + 1]; u[i] = u[i + 1];
            for (int j = 0; j < n; ++j) R[j * n + i] = R[j * n + i + 1];
        }
        act.pop_back(); sgn.pop_back(); u.pop_back();
        for (int j = 0; j < n; ++j) R[j * n + iq - 1] = T(0);
        --iq;
        for (int j = pos; j < iq; ++j) {
            T cc = R[j * n + j], ss = R[(j + 1) * n + j];
            T h = std::sqrt(cc * cc + ss * ss);
            if (h == T(0)) continue;
            cc = cc / h; ss = ss / h;
            R[(j + 1) * n + j] = T(0);
            if (cc < T(0)) { R[j * n + j] = -h; cc = -cc; ss = -ss; } else R[j * n + j] = h;
            T xny = ss / (T(1) + cc);
            for (int k = j + 1; k < iq; ++k) {
                T t1 = R[j * n + k], t2 = R[(j + 1) * n + k];
                R[j * n + k] = t1 * cc + t2 * ss;
                R[(j + 1) * n + k] = xny * (t1 + R[j * n + k]) - t2;
            }
            for (int k = 0; k < n; ++k) {
                T t1 = J[k * n + j], t2 = J[k * n + j + 1];
                J[k * n + j] = t1 * cc + t2 * ss;
                J[k * n + j + 1] = xny * (J[k * n + j] + t1) - t2;
            }
        }
    };

    int iq = 0;
    while (true) {
        // ---- pick the constraint to add: equalities first (in order), then the most violated bound
        int p = -1, ps = 1;
        for (int i = 0; i < m && p < 0; ++i)
            if (is_eq[i] && !in_act[i]) { T s = ax(i) - T(lbA[i]); p = i; ps = (s > 0) ? -1 : 1; }
        if (p < 0) {
            T worst = 0;
            for (int i = 0; i < m; ++i) {
                if (is_eq[i] || in_act[i]) continue;
                T v = ax(i);
                if (lbA[i] > -QP_INFTY) { T s = v - T(lbA[i]); if (s < worst) { worst = s; p = i; ps = 1; } }
                if (ubA[i] < QP_INFTY) { T s = T(ubA[i]) - v; if (s < worst) { worst = s; p = i; ps = -1; } }
            }
            if (p >= 0) {
                T mag = std::fabs(ax(p)) + T(1);
                if (-worst <= T(1e3) * tiny * mag) p = -1;
            }
        }
        if (p < 0) break;
        for (int k = 0; k < n; ++k) np_[k] = T(ps) * T(A[p * n + k]);
        T up = 0;
        bool added = false;
        while (!added) {
            if (++info.iterations > max_iter) { info.status = 1; goto done; }
            // d = J' n+ ; z = J2 d2 ; r = R^-1 d1
            for (int i = 0; i < n; ++i) { T s = 0; for (int k = 0; k < n; ++k) s += J[k * n + i] * np_[k]; d[i] = s; }
            for (int i = 0; i < n; ++i) { T s = 0; for (int k = iq; k < n; ++k) s += J[i * n + k] * d[k]; z[i] = s; }
            for (int i = iq - 1; i >= 0; --i) {
                T s = d[i];
                for (int k = i + 1; k < iq; ++k) s -= R[i * n + k] * r[k];
                r[i] = s / R[i * n + i];
            }
            T t1 = std::numeric_limits<T>::infinity();
            int drop = -1;
            for (int a = 0; a < iq; ++a) {
                if (is_eq[act[a]]) continue;
                if (r[a] > T(0)) { T t = u[a] / r[a]; if (t < t1) { t1 = t; drop = a; } }
            }
            T zn = 0, zz = 0, nn = 0;
            for (int k = 0; k < n; ++k) { zn += z[k] * np_[k]; zz += z[k] * z[k]; nn += np_[k] * np_[k]; }
            T sp = T(ps) * (ax(p) - T(ps > 0 ? lbA[p] : ubA[p]));     // < 0 while violated
            bool z_zero = !(zn > T(1e2) * tiny * std::sqrt(zz * nn) ) || iq >= n;
            T t2 = z_zero ? std::numeric_limits<T>::infinity() : -sp / zn;
            if (!(sp < T(0))) t2 = T(0);
            T t = (t1 < t2) ? t1 : t2;
            if (t == std::numeric_limits<T>::infinity()) { info.status = 2; goto done; }
            if (!z_zero && t2 != T(0)) for (int k = 0; k < n; ++k) x[k] += t * z[k];
            else if (!z_zero) {}
            for (int a = 0; a < iq; ++a) u[a] -= t * r[a];
            up += t;
            if (t == t2) {
                act.push_back(p); sgn.push_back(ps); u.push_back(up);
                if (!add_constraint(iq)) {   // dependent normal: cannot happen for the reference's QPs; flag it
                    info.status = 2;
                    goto done;
                }
                in_act[p] = 1;
                added = true;
            } else {
                in_act[act[drop]] = 0;
                delete_constraint(iq, drop);
            }
        }
    }
done:
    for (int k = 0; k < n; ++k) x_out[k] = (double)x[k];
    if (y_out) {
        for (int i = 0; i < m; ++i) y_out[i] = 0.0;
        for (size_t a = 0; a < act.size() && a < u.size(); ++a) y_out[act[a]] = (double)(T(sgn[a]) * u[a]);
    }
    info.n_active = (int)act.size();
    return info;
}

// KKT certificate: stationarity max|Hx + g - A'y|, primal infeasibility,
// complementarity max|y_i * slack_i|, dual sign violation.
inline void qp_kkt_residuals(int n, const double* H, const double* g, int m, const double* A,
                             const double* lbA, const double* ubA, const double* x,
                             const double* y, double out[4]) {
    double stat = 0, prim = 0, comp = 0, dual = 0;
    for (int i = 0; i < n; ++i) {
        long double s = g[i];
        for (int k = 0; k < n; ++k) s += (long double)H[i * n + k] * x[k];
        for (int j = 0; j < m; ++j) s -= (long double)A[j * n + i] * y[j];
        stat = std::fmax(stat, std::fabs((double)s));
    }
    for (int j = 0; j < m; ++j) {
        long double ax = 0;
        for (int k = 0; k < n; ++k) ax += (long double)A[j * n + k] * x[k];
        double lo = (double)(ax - lbA[j]), hi = (double)(ubA[j] - ax);
        if (lbA[j] > -QP_INFTY) prim = std::fmax(prim, -lo);
        if (ubA[j] < QP_INFTY) prim = std::fmax(prim, -hi);
        if (lbA[j] == ubA[j]) continue;
        if (y[j] > 0) comp = std::fmax(comp, std::fabs(y[j] * lo));
        if (y[j] < 0) comp = std::fmax(comp, std::fabs(y[j] * hi));
        if (ubA[j] >= QP_INFTY && y[j] < 0) dual = std::fmax(dual, -y[j]);
        if (lbA[j] <= -QP_INFTY && y[j] > 0) dual = std::fmax(dual, y[j]);
    }
    out[0] = stat; out[1] = prim; out[2] = comp; out[3] = dual;
}

}  // namespace orc
