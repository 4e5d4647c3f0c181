// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
// Tiny fixed-size dense linear algebra used by the CPU restatement of
// Ema158/linearMpcHumanoid. Replaces the Eigen3 subset the reference uses
// (Eigen is not installed in this image; version unpinned upstream,
// CMakeLists.txt:35).  Products accumulate left-to-right over the inner index.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>

namespace orc {

template <int R, int C>
struct Mat {
    double d[R * C];
    Mat() { std::memset(d, 0, sizeof(d)); }
    double& operator()(int i, int j) { return d[i * C + j]; }
    double operator()(int i, int j) const { return d[i * C + j]; }
    double& operator[](int i) { return d[i]; }
    double operator[](int i) const { return d[i]; }
    static Mat identity() {
        Mat m;
        for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0;
        return m;
    }
    template <int BR, int BC>
    Mat<BR, BC> block(int r0, int c0) const {
        Mat<BR, BC> b;
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) b(i, j) = (*this)(r0 + i, c0 + j);
        return b;
    }
    template <int BR, int BC>
    void set(int r0, int c0, const Mat<BR, BC>& b) {
        for (int i = 0; i < BR; ++i)
            for (int j = 0; j < BC; ++j) (*this)(r0 + i, c0 + j) = b(i, j);
    }
    Mat<C, R> t() const {
        Mat<C, R> m;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) m(j, i) = (*this)(i, j);
        return m;
    }
};
template <int N>
using Vec = Mat<N, 1>;

template <int R, int K, int C>
inline Mat<R, C> operator*(const Mat<R, K>& a, const Mat<K, C>& b) {
    Mat<R, C> m;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += a(i, k) * b(k, j);
            m(i, j) = s;
        }
    return m;
}
template <int R, int C>
inline Mat<R, C> operator+(const Mat<R, C>& a, const Mat<R, C>& b) {
    Mat<R, C> m;
    for (int i = 0; i < R * C; ++i) m.d[i] = a.d[i] + b.d[i];
    return m;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C>& a, const Mat<R, C>& b) {
    Mat<R, C> m;
    for (int i = 0; i < R * C; ++i) m.d[i] = a.d[i] - b.d[i];
    return m;
}
template <int R, int C>
inline Mat<R, C> operator-(const Mat<R, C>& a) {
    Mat<R, C> m;
    for (int i = 0; i < R * C; ++i) m.d[i] = -a.d[i];
    return m;
}
template <int R, int C>
inline Mat<R, C> operator*(double s, const Mat<R, C>& a) {
    Mat<R, C> m;
    for (int i = 0; i < R * C; ++i) m.d[i] = s * a.d[i];
    return m;
}
template <int R, int C>
inline Mat<R, C> operator*(const Mat<R, C>& a, double s) { return s * a; }
template <int R, int C>
inline Mat<R, C> operator/(const Mat<R, C>& a, double s) {
    Mat<R, C> m;
    for (int i = 0; i < R * C; ++i) m.d[i] = a.d[i] / s;
    return m;
}
template <int N>
inline double dot(const Vec<N>& a, const Vec<N>& b) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += a[i] * b[i];
    return s;
}
template <int R, int C>
inline double max_abs(const Mat<R, C>& a) {
    double m = 0.0;
    for (int i = 0; i < R * C; ++i) m = std::max(m, std::fabs(a.d[i]));
    return m;
}
template <int R, int C>
inline bool all_finite(const Mat<R, C>& a) {
    for (int i = 0; i < R * C; ++i)
        if (!std::isfinite(a.d[i])) return false;
    return true;
}
inline Vec<3> v3(double x, double y, double z) {
    Vec<3> v;
    v[0] = x; v[1] = y; v[2] = z;
    return v;
}

// 3x3 inverse by cofactors (what Eigen does for fixed 3x3 `.inverse()`).
inline Mat<3, 3> inverse3(const Mat<3, 3>& m) {
    Mat<3, 3> c;
    c(0, 0) = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1);
    c(0, 1) = m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2);
    c(0, 2) = m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1);
    c(1, 0) = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2);
    c(1, 1) = m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0);
    c(1, 2) = m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2);
    c(2, 0) = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
    c(2, 1) = m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1);
    c(2, 2) = m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
    double det = m(0, 0) * c(0, 0) + m(0, 1) * c(1, 0) + m(0, 2) * c(2, 0);
    return c / det;
}

// Solve A x = b by Householder QR with column pivoting (the decomposition the
// reference asks Eigen for: src/controller.cpp:143, src/invKinematics.cpp:43,
// src/generalizedFunctions.cpp:161).  A is n x n row-major, overwritten.
inline void qr_colpiv_solve(int n, double* A, const double* b, double* x) {
    std::vector<int> perm(n);
    std::vector<double> rhs(b, b + n), cn(n), w(n);
    for (int j = 0; j < n; ++j) {
        perm[j] = j;
        double s = 0;
        for (int i = 0; i < n; ++i) s += A[i * n + j] * A[i * n + j];
        cn[j] = s;
    }
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int j = k + 1; j < n; ++j)
            if (cn[j] > cn[p]) p = j;
        if (p != k) {
            for (int i = 0; i < n; ++i) std::swap(A[i * n + k], A[i * n + p]);
            std::swap(cn[k], cn[p]);
            std::swap(perm[k], perm[p]);
        }
        double nrm = 0;
        for (int i = k; i < n; ++i) nrm += A[i * n + k] * A[i * n + k];
        nrm = std::sqrt(nrm);
        if (nrm == 0.0) continue;
        double alpha = (A[k * n + k] > 0) ? -nrm : nrm;
        // v = a_k - alpha e_k
        for (int i = k; i < n; ++i) w[i] = A[i * n + k];
        w[k] -= alpha;
        double vv = 0;
        for (int i = k; i < n; ++i) vv += w[i] * w[i];
        if (vv > 0) {
            for (int j = k; j < n; ++j) {
                double s = 0;
                for (int i = k; i < n; ++i) s += w[i] * A[i * n + j];
                s = 2.0 * s / vv;
                for (int i = k; i < n; ++i) A[i * n + j] -= s * w[i];
            }
            double s = 0;
            for (int i = k; i < n; ++i) s += w[i] * rhs[i];
            s = 2.0 * s / vv;
            for (int i = k; i < n; ++i) rhs[i] -= s * w[i];
        }
        for (int j = k + 1; j < n; ++j) {
            double s = 0;
            for (int i = k + 1; i < n; ++i) s += A[i * n + j] * A[i * n + j];
            cn[j] = s;
        }
    }
    std::vector<double> y(n);
    for (int i = n - 1; i >= 0; --i) {
        double s = rhs[i];
        for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * y[j];
        y[i] = s / A[i * n + i];
    }
    for (int j = 0; j < n; ++j) x[perm[j]] = y[j];
}

template <int N>
inline Vec<N> qr_solve(Mat<N, N> A, const Vec<N>& b) {
    Vec<N> x;
    qr_colpiv_solve(N, A.d, b.d, x.d);
    return x;
}

}  // namespace orc
