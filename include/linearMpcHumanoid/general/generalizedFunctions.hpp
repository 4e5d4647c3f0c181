// generalizedFunctions.hpp -- the free helper functions of the reference's class surface
// (include/linearMpcHumanoid/general/generalizedFunctions.hpp:6-34, src/generalizedFunctions.cpp), same names and
// signatures, header-only.  These are small closed-form data-format helpers that the offline app's dynamics() closure
// and user code call on the HOST (apps/offline/main.cpp:111,116 use crossMatrix and matrixAngularVelToEulerDot); the
// per-step model computations they support upstream run on the GPU here (include/lmh_capi.h).  findPolyCoeff is
// the library's polynomial fit (lmh_poly_fit).
#pragma once
#include "linearMpcHumanoid/general/lmhEigen.hpp"
#include "lmh_capi.h"
#include <cmath>
#include <iostream>
#include <stdexcept>
#include <string>

// [v]x : the matrix with [v]x w = v x w (generalizedFunctions.cpp:3-9)
inline Eigen::Matrix3d crossMatrix(Eigen::Vector3d v) {
    Eigen::Matrix3d S = Eigen::Matrix3d::Zero();
    S(0, 1) = -v(2); S(0, 2) = v(1);
    S(1, 0) = v(2);  S(1, 2) = -v(0);
    S(2, 0) = -v(1); S(2, 1) = v(0);
    return S;
}

// 6 x 6 motion transform [[E, 0], [-E [p]x, E]], E = R', of a homogeneous transform (R, p) (generalizedFunctions.cpp:11-19)
inline Eigen::MatrixXd velocityMatrix(Eigen::Matrix4d T) {
    Eigen::Matrix3d E;
    Eigen::Vector3d p;
    for (int i = 0; i < 3; ++i) {
        p(i) = T(i, 3);
        for (int j = 0; j < 3; ++j) E(i, j) = T(j, i);
    }
    const Eigen::Matrix3d B = -(E * crossMatrix(p));
    Eigen::MatrixXd X = Eigen::MatrixXd::Zero(6, 6);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { X(i, j) = E(i, j); X(3 + i, 3 + j) = E(i, j); X(3 + i, j) = B(i, j); }
    return X;
}

// inverse by transposition of the rotation block, as upstream (generalizedFunctions.cpp:21-27)
inline Eigen::Matrix4d inverseTransformationMatrix(Eigen::Matrix4d T) {
    Eigen::Matrix4d I = Eigen::Matrix4d::Zero();
    for (int i = 0; i < 3; ++i) {
        double s = 0.0;
        for (int j = 0; j < 3; ++j) { I(i, j) = T(j, i); s += -T(j, i) * T(j, 3); }
        I(i, 3) = s;
    }
    I(3, 3) = 1.0;
    return I;
}

// spatial cross product matrices of a twist [w; u] (generalizedFunctions.cpp:29-41)
inline Eigen::MatrixXd spatialCrossMatrix(Eigen::VectorXd v) {
    const Eigen::Matrix3d W = crossMatrix(Eigen::Vector3d(v(0), v(1), v(2))), U = crossMatrix(Eigen::Vector3d(v(3), v(4), v(5)));
    Eigen::MatrixXd m = Eigen::MatrixXd::Zero(6, 6);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { m(i, j) = W(i, j); m(3 + i, j) = U(i, j); m(3 + i, 3 + j) = W(i, j); }
    return m;
}
inline Eigen::MatrixXd spatialCrossMatrixForce(Eigen::VectorXd v) {
    const Eigen::MatrixXd m = spatialCrossMatrix(v);
    Eigen::MatrixXd f = Eigen::MatrixXd::Zero(6, 6);
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) f(i, j) = -m(j, i);
    return f;
}

// angular velocity -> roll/pitch/yaw rates (generalizedFunctions.cpp:43-50)
inline Eigen::Matrix3d matrixAngularVelToEulerDot(Eigen::Vector3d eta) {
    const double cy = std::cos(eta(2)), sy = std::sin(eta(2)), cp = std::cos(eta(1)), tp = std::tan(eta(1));
    Eigen::Matrix3d O = Eigen::Matrix3d::Zero();
    O(0, 0) = cy / cp;  O(0, 1) = sy / cp;
    O(1, 0) = -sy;      O(1, 1) = cy;
    O(2, 0) = cy * tp;  O(2, 1) = sy * tp;  O(2, 2) = 1.0;
    return O;
}

// R = Rz(yaw) Ry(pitch) Rx(roll) (generalizedFunctions.cpp:52-72)
inline Eigen::Matrix3d eulerAnglesToSO3(const Eigen::Vector3d& rpy) {
    const double cr = std::cos(rpy(0)), sr = std::sin(rpy(0)), cp = std::cos(rpy(1)), sp = std::sin(rpy(1));
    const double cy = std::cos(rpy(2)), sy = std::sin(rpy(2));
    Eigen::Matrix3d R;
    R(0, 0) = cy * cp; R(0, 1) = cy * sp * sr - sy * cr; R(0, 2) = cy * sp * cr + sy * sr;
    R(1, 0) = sy * cp; R(1, 1) = sy * sp * sr + cy * cr; R(1, 2) = sy * sp * cr - cy * sr;
    R(2, 0) = -sp;     R(2, 1) = cp * sr;                R(2, 2) = cp * cr;
    return R;
}

// rotation vector (axis times angle) of a rotation matrix, small-angle safe (generalizedFunctions.cpp:74-101)
inline Eigen::Vector3d rotMatToAxisAngle(const Eigen::Matrix3d& R) {
    double c = 0.5 * ((R(0, 0) + R(1, 1) + R(2, 2)) - 1.0);
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    const double phi = std::acos(c);
    const double scale = phi < 1e-6 ? 0.5 : phi / (2.0 * std::sin(phi));
    return Eigen::Vector3d(scale * (R(2, 1) - R(1, 2)), scale * (R(0, 2) - R(2, 0)), scale * (R(1, 0) - R(0, 1)));
}

// polynomial (ascending powers) through position / velocity / acceleration constraints given as (time, value) rows
// (generalizedFunctions.cpp:103-163); solved inside the library
inline Eigen::VectorXd findPolyCoeff(Eigen::MatrixXd Pos, Eigen::MatrixXd Vel, Eigen::MatrixXd Acc) {
    if ((Pos.rows() && Pos.cols() != 2) || (Vel.rows() && Vel.cols() != 2) || (Acc.rows() && Acc.cols() != 2))
        std::cout << "Number of columns should be 2 for Position Velocity and Acceleration" << std::endl;
    auto rows = [](const Eigen::MatrixXd& m) {
        std::vector<double> r;
        for (int i = 0; i < (int)m.rows(); ++i) { r.push_back(m(i, 0)); r.push_back(m(i, 1)); }
        return r;
    };
    const std::vector<double> p = rows(Pos), v = rows(Vel), a = rows(Acc);
    const int n = (int)(Pos.rows() + Vel.rows() + Acc.rows());
    Eigen::VectorXd c = Eigen::VectorXd::Zero(n);
    if (lmh_poly_fit((int)Pos.rows(), p.data(), (int)Vel.rows(), v.data(), (int)Acc.rows(), a.data(), c.data()) != LMH_OK)
        throw std::runtime_error(std::string("findPolyCoeff: ") + lmh_last_error());
    return c;
}

inline double polyval(const Eigen::VectorXd& poly, double x) {          // generalizedFunctions.cpp:165-177
    double value = 0.0, xp = 1.0;
    for (int i = 0; i < (int)poly.size(); ++i) { value += poly(i) * xp; xp *= x; }
    return value;
}

inline const Eigen::VectorXd polyder(const Eigen::VectorXd& poly) {      // generalizedFunctions.cpp:179-193
    const int n = (int)poly.size();
    if (n <= 1) return Eigen::VectorXd::Zero(1);
    Eigen::VectorXd d(n - 1);
    for (int i = 1; i < n; ++i) d(i - 1) = i * poly(i);
    return d;
}

// [v_lin; w] in the world frame -> [w; v_lin] in the base frame (generalizedFunctions.cpp:195-206)
inline void swapBaseVelocityAndRefToWorldFrame(const Eigen::MatrixXd& X01, Eigen::VectorXd& v) {
    double tw[6] = {v(3), v(4), v(5), v(0), v(1), v(2)};
    for (int i = 0; i < 6; ++i) {
        double s = 0.0;
        for (int j = 0; j < 6; ++j) s += X01(i, j) * tw[j];
        v(i) = s;
    }
}
