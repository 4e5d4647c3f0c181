// lmhEigen.hpp -- the dense vector/matrix types of the class surface.
//
// The reference's public headers are Eigen-typed (Eigen::VectorXd, Vector3d, MatrixXd ... in every signature of
// include/linearMpcHumanoid/**).  With Eigen3 on the include path this header is just <Eigen/Dense>.  Without it (the
// build image of this repository has no Eigen) it supplies, in namespace Eigen, the small subset of that API which the
// class surface and apps/offline/main.cpp use: Zero / Identity / Constant, (i) and (i,j), size/rows/cols/data, the comma
// initialiser, segment / head / tail / block as assignable views, + - * (matrix, vector, scalar), += -=, transpose,
// dot / cross / norm / cwiseAbs / maxCoeff.  Everything is eager and column-major like Eigen's defaults.  Nothing here
// runs per environment step: the numbers come from the GPU library, these types only carry them.
#pragma once
#if defined(__has_include)
#if __has_include(<Eigen/Dense>) && !defined(LMH_FORCE_EIGEN_LITE)
#include <Eigen/Dense>
#define LMH_HAVE_EIGEN 1
#endif
#endif

#ifndef LMH_HAVE_EIGEN
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <ostream>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace Eigen {

using Index = std::ptrdiff_t;
constexpr int Dynamic = -1;
class Dense;

// assignable view of a rectangular part of a Dense object (segment / head / tail / block / row / col)
class Block {
public:
    Block(Dense& m, Index r0, Index c0, Index nr, Index nc) : m_(&m), r0_(r0), c0_(c0), nr_(nr), nc_(nc) {}
    Block(const Block&) = default;
    Index rows() const { return nr_; }
    Index cols() const { return nc_; }
    Index size() const { return nr_ * nc_; }
    inline double& operator()(Index i, Index j);
    inline double operator()(Index i, Index j) const;
    double& operator()(Index i) { return nc_ == 1 ? (*this)(i, 0) : (*this)(0, i); }
    double operator()(Index i) const { return nc_ == 1 ? (*this)(i, 0) : (*this)(0, i); }
    inline Block& operator=(const Dense& v);
    inline Block& operator=(const Block& v);
    inline Block& operator+=(const Dense& v);
    inline Block& operator-=(const Dense& v);
    class Comma;
    inline Comma operator<<(double v);
    inline Comma operator<<(const Dense& v);
private:
    Dense* m_;
    Index r0_, c0_, nr_, nc_;
};

class Dense {
public:
    Dense() {}
    Dense(Index r, Index c) : r_(r), c_(c), d_((size_t)(r * c), 0.0) {}
    Dense(const Block& b) : r_(b.rows()), c_(b.cols()), d_((size_t)b.size()) {
        for (Index j = 0; j < c_; ++j) for (Index i = 0; i < r_; ++i) (*this)(i, j) = b(i, j);
    }
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
    double& operator()(Index i, Index j) { return d_[(size_t)(j * r_ + i)]; }
    double operator()(Index i, Index j) const { return d_[(size_t)(j * r_ + i)]; }
    double& operator()(Index i) { return d_[(size_t)i]; }
    double operator()(Index i) const { return d_[(size_t)i]; }
    double& operator[](Index i) { return d_[(size_t)i]; }
    double operator[](Index i) const { return d_[(size_t)i]; }
    double& x() { return d_[0]; }
    double& y() { return d_[1]; }
    double& z() { return d_[2]; }
    double x() const { return d_[0]; }
    double y() const { return d_[1]; }
    double z() const { return d_[2]; }
    void resize(Index r, Index c) { r_ = r; c_ = c; d_.assign((size_t)(r * c), 0.0); }
    void resize(Index n) { if (c_ == 1 || r_ != 1) resize(n, 1); else resize(1, n); }
    void setZero() { for (double& v : d_) v = 0.0; }
    void setZero(Index n) { resize(n); }
    void setZero(Index r, Index c) { resize(r, c); }
    void setConstant(double v) { for (double& x : d_) x = v; }
    void setIdentity() { setZero(); for (Index i = 0; i < r_ && i < c_; ++i) (*this)(i, i) = 1.0; }
    // views
    Block block(Index r0, Index c0, Index nr, Index nc) { return Block(*this, r0, c0, nr, nc); }
    Dense block(Index r0, Index c0, Index nr, Index nc) const { return Dense(Block(const_cast<Dense&>(*this), r0, c0, nr, nc)); }
    Block segment(Index i, Index n) { return c_ == 1 ? Block(*this, i, 0, n, 1) : Block(*this, 0, i, 1, n); }
    Dense segment(Index i, Index n) const { return Dense(const_cast<Dense&>(*this).segment(i, n)); }
    Block head(Index n) { return segment(0, n); }
    Dense head(Index n) const { return segment(0, n); }
    Block tail(Index n) { return segment(size() - n, n); }
    Dense tail(Index n) const { return segment(size() - n, n); }
    Block row(Index i) { return Block(*this, i, 0, 1, c_); }
    Dense row(Index i) const { return block(i, 0, 1, c_); }
    Block col(Index j) { return Block(*this, 0, j, r_, 1); }
    Dense col(Index j) const { return block(0, j, r_, 1); }
    Block::Comma operator<<(double v);
    Block::Comma operator<<(const Dense& v);
    // algebra
    Dense transpose() const { Dense t(c_, r_); for (Index j = 0; j < c_; ++j) for (Index i = 0; i < r_; ++i) t(j, i) = (*this)(i, j); return t; }
    double dot(const Dense& o) const { check_same(o); double s = 0; for (size_t k = 0; k < d_.size(); ++k) s += d_[k] * o.d_[k]; return s; }
    double squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt(squaredNorm()); }
    double sum() const { double s = 0; for (double v : d_) s += v; return s; }
    double maxCoeff() const { double m = d_.at(0); for (double v : d_) m = v > m ? v : m; return m; }
    double minCoeff() const { double m = d_.at(0); for (double v : d_) m = v < m ? v : m; return m; }
    Dense cwiseAbs() const { Dense a(*this); for (double& v : a.d_) v = std::fabs(v); return a; }
    Dense cross(const Dense& o) const {
        if (size() != 3 || o.size() != 3) throw std::invalid_argument("Eigen(lite): cross needs 3-vectors");
        Dense c(3, 1);
        c(0) = d_[1] * o.d_[2] - d_[2] * o.d_[1]; c(1) = d_[2] * o.d_[0] - d_[0] * o.d_[2]; c(2) = d_[0] * o.d_[1] - d_[1] * o.d_[0];
        return c;
    }
    Dense& operator+=(const Dense& o) { check_same(o); for (size_t k = 0; k < d_.size(); ++k) d_[k] += o.d_[k]; return *this; }
    Dense& operator-=(const Dense& o) { check_same(o); for (size_t k = 0; k < d_.size(); ++k) d_[k] -= o.d_[k]; return *this; }
    Dense& operator*=(double s) { for (double& v : d_) v *= s; return *this; }
    Dense& operator/=(double s) { for (double& v : d_) v /= s; return *this; }
    void check_same(const Dense& o) const {
        if (o.size() != size()) throw std::invalid_argument("Eigen(lite): size mismatch");
    }
protected:
    Index r_ = 0, c_ = 0;
    std::vector<double> d_;
};

// a << 1, 2, 3;  fills in row-major order like Eigen's comma initialiser (vectors may be appended as blocks)
class Block::Comma {
public:
    Comma(Block b) : b_(b) {}
    Comma& operator,(double v) { put(v); return *this; }
    Comma& operator,(const Dense& v) { for (Index k = 0; k < v.size(); ++k) put(v(k)); return *this; }
    void put(double v) {
        if (k_ >= b_.size()) throw std::out_of_range("Eigen(lite): too many coefficients in a comma initialiser");
        b_(k_ / b_.cols(), k_ % b_.cols()) = v;
        ++k_;
    }
private:
    Block b_;
    Index k_ = 0;
};

inline double& Block::operator()(Index i, Index j) { return (*m_)(r0_ + i, c0_ + j); }
inline double Block::operator()(Index i, Index j) const { return (*m_)(r0_ + i, c0_ + j); }
inline Block& Block::operator=(const Dense& v) {
    if (v.size() != size()) throw std::invalid_argument("Eigen(lite): block assignment of the wrong size");
    const bool same = v.rows() == nr_;
    for (Index j = 0; j < nc_; ++j) for (Index i = 0; i < nr_; ++i) (*this)(i, j) = same ? v(i, j) : v(j * nr_ + i);
    return *this;
}
inline Block& Block::operator=(const Block& v) { return *this = Dense(v); }
inline Block& Block::operator+=(const Dense& v) { Dense t(*this); t += v; return *this = t; }
inline Block& Block::operator-=(const Dense& v) { Dense t(*this); t -= v; return *this = t; }
inline Block::Comma Block::operator<<(double v) { Comma c(*this); c.put(v); return c; }
inline Block::Comma Block::operator<<(const Dense& v) { Comma c(*this); c, v; return c; }
inline Block::Comma Dense::operator<<(double v) { return Block(*this, 0, 0, r_, c_) << v; }
inline Block::Comma Dense::operator<<(const Dense& v) { return Block(*this, 0, 0, r_, c_) << v; }

inline Dense operator+(const Dense& a, const Dense& b) { Dense r(a); r += b; return r; }
inline Dense operator-(const Dense& a, const Dense& b) { Dense r(a); r -= b; return r; }
inline Dense operator-(const Dense& a) { Dense r(a); r *= -1.0; return r; }
inline Dense operator*(double s, const Dense& a) { Dense r(a); r *= s; return r; }
inline Dense operator*(const Dense& a, double s) { Dense r(a); r *= s; return r; }
inline Dense operator/(const Dense& a, double s) { Dense r(a); r /= s; return r; }
inline Dense operator*(const Dense& a, const Dense& b) {
    if (a.cols() != b.rows()) throw std::invalid_argument("Eigen(lite): product of incompatible shapes");
    Dense r(a.rows(), b.cols());
    for (Index j = 0; j < b.cols(); ++j)
        for (Index k = 0; k < a.cols(); ++k) {
            const double bkj = b(k, j);
            for (Index i = 0; i < a.rows(); ++i) r(i, j) += a(i, k) * bkj;
        }
    return r;
}
inline std::ostream& operator<<(std::ostream& os, const Dense& m) {
    for (Index i = 0; i < m.rows(); ++i) {
        for (Index j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

// Matrix<double, R, C>: the shape-carrying front of Dense (R, C = Dynamic or a fixed size)
template <typename Scalar, int R, int C, int Options = 0>
class Matrix : public Dense {
public:
    Matrix() : Dense(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C) {}
    explicit Matrix(Index n) : Dense(C == 1 ? n : 1, C == 1 ? 1 : n) {}
    Matrix(Index r, Index c) : Dense(r, c) {}
    template <typename T, typename = typename std::enable_if<std::is_floating_point<T>::value>::type>
    Matrix(T a, T b) : Dense(C == 1 ? 2 : 1, C == 1 ? 1 : 2) { d_[0] = a; d_[1] = b; }
    Matrix(double a, double b, double c) : Dense(3, 1) { d_[0] = a; d_[1] = b; d_[2] = c; }
    Matrix(std::initializer_list<double> v) : Dense(C == 1 ? (Index)v.size() : 1, C == 1 ? 1 : (Index)v.size()) {
        size_t k = 0;
        for (double x : v) d_[k++] = x;
    }
    Matrix(const Dense& o) : Dense(o) { adopt(); }
    Matrix(const Block& b) : Dense(b) { adopt(); }
    Matrix& operator=(const Dense& o) { Dense::operator=(o); adopt(); return *this; }
    Matrix& operator=(const Block& b) { Dense::operator=(Dense(b)); adopt(); return *this; }
    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(Index n) { return Matrix(n); }
    static Matrix Zero(Index r, Index c) { return Matrix(r, c); }
    static Matrix Constant(Index n, double v) { Matrix m(n); m.setConstant(v); return m; }
    static Matrix Constant(Index r, Index c, double v) { Matrix m(r, c); m.setConstant(v); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix Identity(Index r, Index c) { Matrix m(r, c); m.setIdentity(); return m; }
private:
    // a vector type takes the other orientation's data as is (Eigen allows row/column vector assignment)
    void adopt() {
        if (C == 1 && c_ != 1 && r_ == 1) { r_ = c_; c_ = 1; }
        if (R != Dynamic && C != Dynamic && size() != (Index)R * C) throw std::invalid_argument("Eigen(lite): fixed-size assignment of the wrong size");
    }
};

using VectorXd = Matrix<double, Dynamic, 1>;
using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector4d = Matrix<double, 4, 1>;
using RowVector2d = Matrix<double, 1, 2>;
using Matrix2d = Matrix<double, 2, 2>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;
constexpr int RowMajor = 1;

}  // namespace Eigen
#endif  // !LMH_HAVE_EIGEN
