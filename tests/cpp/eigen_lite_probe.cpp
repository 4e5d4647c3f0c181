// TEST PROGRAM (CPU only): the Eigen-API subset of include/linearMpcHumanoid/general/lmhEigen.hpp and the host helpers of
// generalizedFunctions.hpp / rk4.hpp, exercised with the expressions apps/offline/main.cpp and the class surface use.
// Prints "ok" or the first failed check.  (With real Eigen on the include path the same program tests Eigen itself.)
#include "linearMpcHumanoid/general/generalizedFunctions.hpp"
#include "linearMpcHumanoid/general/rk4.hpp"
#include "linearMpcHumanoid/general/Clock.hpp"
#include <cstdio>
#include <cstdlib>

#define CHECK(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)
static bool near(double a, double b, double tol = 1e-14) { return std::fabs(a - b) <= tol; }

int main() {
    // construction, Zero, element access, comma initialiser (main.cpp:24-31, 47-55)
    Eigen::VectorXd Rf = Eigen::VectorXd::Zero(6);
    Rf(1) = -0.05;
    CHECK(Rf.size() == 6 && Rf(1) == -0.05 && Rf(0) == 0.0);
    Eigen::Vector3d com = Eigen::Vector3d::Zero();
    com << -0.02, 0.0, 0.26;
    CHECK(com(0) == -0.02 && com(2) == 0.26 && com.size() == 3);
    Eigen::Vector3d com2(1.0, 2.0, 3.0);
    com2 << 4, 5, 6;                                              // re-initialisation (main.cpp:52-53)
    CHECK(com2(0) == 4.0 && com2(2) == 6.0);
    // segment / head / tail as views (main.cpp:71-72, 95-96, 111-120)
    const int n = 30;
    Eigen::VectorXd state(2 * n);
    Eigen::VectorXd q0 = Eigen::VectorXd::Zero(n), v0 = Eigen::VectorXd::Zero(n);
    for (int i = 0; i < n; ++i) { q0(i) = 0.1 * i; v0(i) = -0.01 * i; }
    state.segment(0, n) = q0;
    state.segment(n, n) = v0;
    CHECK(state(3) == q0(3) && state(n + 7) == v0(7));
    const Eigen::VectorXd& cs = state;
    Eigen::VectorXd q = cs.segment(0, n), qD = cs.segment(n, n);
    CHECK(q.size() == n && qD(5) == v0(5));
    Eigen::Vector3d w3 = qD.segment(3, 3), p3 = q.segment(0, 3);
    Eigen::Vector3d expect = Eigen::Vector3d(qD(0), qD(1), qD(2)) + crossMatrix(w3) * p3;
    qD.segment(0, 3) += crossMatrix(qD.segment(3, 3)) * q.segment(0, 3);      // main.cpp:111 verbatim
    CHECK(near(qD(0), expect(0)) && near(qD(1), expect(1)) && near(qD(2), expect(2)));
    Eigen::Vector3d om(qD(3), qD(4), qD(5));
    Eigen::Vector3d rates = matrixAngularVelToEulerDot(Eigen::Vector3d(q(3), q(4), q(5))) * om;
    qD.segment(3, 3) = matrixAngularVelToEulerDot(q.segment(3, 3)) * qD.segment(3, 3);   // main.cpp:116: the right side reads what the left side overwrites
    CHECK(near(qD(3), rates(0)) && near(qD(4), rates(1)) && near(qD(5), rates(2)));
    Eigen::VectorXd xp(2 * n);
    xp.head(n) = qD;
    xp.tail(n) = q;
    CHECK(xp(0) == qD(0) && xp(2 * n - 1) == q(n - 1) && xp(n) == q(0));
    // arithmetic used by rk4Step: x + 0.5 * dt * k, sums, scalar products (rk4.hpp:12-17)
    const double dt = 0.01;
    Eigen::VectorXd k1 = Eigen::VectorXd::Constant(4, 2.0), x = Eigen::VectorXd::Constant(4, 1.0);
    Eigen::VectorXd y = x + 0.5 * dt * k1;
    CHECK(near(y(2), 1.0 + 0.5 * dt * 2.0));
    // rk4Step on dx/dt = -x against the closed form of one RK4 step
    Eigen::VectorXd z = rk4Step([](const Eigen::VectorXd& s, double) { Eigen::VectorXd d = -1.0 * s; return d; }, x, 0.0, dt);
    const double g = 1 - dt + dt * dt / 2 - dt * dt * dt / 6 + dt * dt * dt * dt / 24;
    CHECK(near(z(0), g, 1e-15) && near(z(3), g, 1e-15));
    // helper functions against their definitions
    Eigen::Matrix3d S = crossMatrix(Eigen::Vector3d(1, 2, 3));
    CHECK(S(0, 1) == -3 && S(0, 2) == 2 && S(1, 0) == 3 && S(1, 2) == -1 && S(2, 0) == -2 && S(2, 1) == 1 && S(0, 0) == 0);
    Eigen::Matrix3d R = eulerAnglesToSO3(Eigen::Vector3d(0.1, -0.2, 0.3));
    Eigen::Matrix3d RtR = R.transpose() * R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) CHECK(near(RtR(i, j), i == j ? 1.0 : 0.0, 1e-15));
    Eigen::Vector3d aa = rotMatToAxisAngle(eulerAnglesToSO3(Eigen::Vector3d(0.0, 0.0, 0.4)));      // yaw 0.4 -> 0.4 e_z
    CHECK(near(aa(0), 0) && near(aa(1), 0) && near(aa(2), 0.4, 1e-14));
    Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T(i, j) = R(i, j); T(i, 3) = 0.1 * (i + 1); }
    Eigen::Matrix4d TiT = inverseTransformationMatrix(T) * T;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) CHECK(near(TiT(i, j), i == j ? 1.0 : 0.0, 1e-15));
    Eigen::MatrixXd X = velocityMatrix(T);
    CHECK(X.rows() == 6 && X.cols() == 6 && near(X(0, 0), R(0, 0)) && near(X(3, 3), R(0, 0)) && X(0, 3) == 0.0);
    Eigen::Vector3d pxe = -1.0 * (R.transpose() * (crossMatrix(Eigen::Vector3d(0.1, 0.2, 0.3)) * Eigen::Vector3d(1, 0, 0)));
    CHECK(near(X(3, 0), pxe(0)) && near(X(4, 0), pxe(1)) && near(X(5, 0), pxe(2)));
    Eigen::VectorXd tw = Eigen::VectorXd::Zero(6);
    tw << 1, 2, 3, 4, 5, 6;
    Eigen::MatrixXd cm = spatialCrossMatrix(tw), cf = spatialCrossMatrixForce(tw);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) CHECK(cf(i, j) == -cm(j, i));
    Eigen::VectorXd poly(4);
    poly << 1, -2, 0.5, 3;
    CHECK(near(polyval(poly, 2.0), 1 - 4 + 2 + 24) && polyder(poly).size() == 3 && near(polyder(poly)(2), 9.0) && polyder(Eigen::VectorXd::Zero(1)).size() == 1);
    Eigen::VectorXd vv = Eigen::VectorXd::Zero(30);
    for (int i = 0; i < 6; ++i) vv(i) = i + 1;
    Eigen::MatrixXd X01 = Eigen::MatrixXd::Identity(6, 6);
    swapBaseVelocityAndRefToWorldFrame(X01, vv);
    CHECK(vv(0) == 4 && vv(2) == 6 && vv(3) == 1 && vv(5) == 3);
    // findPolyCoeff through the library (host only): quintic rest-to-rest from 0 to 1 over [0, 2]
    Eigen::MatrixXd P2(2, 2), Z2(2, 2);
    P2 << 0, 0, 2, 1;
    Z2 << 0, 0, 2, 0;
    Eigen::VectorXd c5 = findPolyCoeff(P2, Z2, Z2);
    CHECK(c5.size() == 6 && near(polyval(c5, 2.0), 1.0, 1e-13) && near(polyval(c5, 1.0), 0.5, 1e-13) && near(polyval(polyder(c5), 2.0), 0.0, 1e-13));
    // Clock accumulates
    Clock clk(0.01, 5);
    double acc = 0;
    for (int i = 0; i < 100; ++i) { clk.step(); acc += 0.01; }
    CHECK(clk.getTime() == acc && clk.getTime() != 1.0 && clk.getTimeStep() == 0.01 && clk.getSimulationTime() == 5);
    std::printf("ok\n");
    return 0;
}
