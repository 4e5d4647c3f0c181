// ORACLE — TEST INFRASTRUCTURE ONLY. See orc_model.hpp for scope and citations.
#include "orc_model.hpp"
#include <cassert>
#include <cstdio>
#include <cstdlib>

namespace orc {

static const double kPi = 3.14159265358979323846;  // src/Robot.cpp:3

// ------------------------------------------------------------------ helpers
Mat<3, 3> crossMatrix(const Vec<3>& v) {
    Mat<3, 3> A;
    A(0, 1) = -v[2]; A(0, 2) = v[1];
    A(1, 0) = v[2];  A(1, 2) = -v[0];
    A(2, 0) = -v[1]; A(2, 1) = v[0];
    return A;
}

Mat<6, 6> velocityMatrix(const Mat<4, 4>& T) {
    Mat<3, 3> R = T.block<3, 3>(0, 0).t();
    Vec<3> p = T.block<3, 1>(0, 3);
    Mat<6, 6> X;
    X.set(0, 0, R);
    X.set(3, 0, (-R) * crossMatrix(p));
    X.set(3, 3, R);
    return X;
}

Mat<4, 4> inverseTransformationMatrix(const Mat<4, 4>& T) {
    Mat<4, 4> Ti;
    Mat<3, 3> Rt = T.block<3, 3>(0, 0).t();
    Ti.set(0, 0, Rt);
    Ti.set(0, 3, (-Rt) * T.block<3, 1>(0, 3));
    Ti(3, 3) = 1.0;
    return Ti;
}

Mat<6, 6> spatialCrossMatrix(const Vec<6>& v) {
    Mat<6, 6> m;
    Mat<3, 3> w = crossMatrix(v.block<3, 1>(0, 0)), l = crossMatrix(v.block<3, 1>(3, 0));
    m.set(0, 0, w);
    m.set(3, 0, l);
    m.set(3, 3, w);
    return m;
}

Mat<6, 6> spatialCrossMatrixForce(const Vec<6>& v) { return -(spatialCrossMatrix(v).t()); }

Mat<3, 3> matrixAngularVelToEulerDot(const Vec<3>& eta) {
    Mat<3, 3> O;
    double c2 = std::cos(eta[2]), s2 = std::sin(eta[2]), c1 = std::cos(eta[1]), t1 = std::tan(eta[1]);
    O(0, 0) = c2 / c1;  O(0, 1) = s2 / c1;  O(0, 2) = 0;
    O(1, 0) = -s2;      O(1, 1) = c2;       O(1, 2) = 0;
    O(2, 0) = c2 * t1;  O(2, 1) = s2 * t1;  O(2, 2) = 1;
    return O;
}

Mat<3, 3> eulerAnglesToSO3(const Vec<3>& rpy) {
    double cr = std::cos(rpy[0]), sr = std::sin(rpy[0]);
    double cp = std::cos(rpy[1]), sp = std::sin(rpy[1]);
    double cy = std::cos(rpy[2]), sy = std::sin(rpy[2]);
    Mat<3, 3> R;
    R(0, 0) = cy * cp; R(0, 1) = cy * sp * sr - sy * cr; R(0, 2) = cy * sp * cr + sy * sr;
    R(1, 0) = sy * cp; R(1, 1) = sy * sp * sr + cy * cr; R(1, 2) = sy * sp * cr - cy * sr;
    R(2, 0) = -sp;     R(2, 1) = cp * sr;                R(2, 2) = cp * cr;
    return R;
}

Vec<3> rotMatToAxisAngle(const Mat<3, 3>& R) {
    double tr = R(0, 0) + R(1, 1) + R(2, 2);
    double c = std::max(-1.0, std::min(1.0, (tr - 1.0) / 2.0));
    double phi = std::acos(c);
    Vec<3> v = v3(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
    if (phi < 1e-6) return 0.5 * v;
    return (phi / (2.0 * std::sin(phi))) * v;
}

std::vector<double> findPolyCoeff(const std::vector<double>& pos_t, const std::vector<double>& pos_v,
                                  const std::vector<double>& vel_t, const std::vector<double>& vel_v,
                                  const std::vector<double>& acc_t, const std::vector<double>& acc_v) {
    int n = (int)(pos_t.size() + vel_t.size() + acc_t.size());
    std::vector<double> A(n * n, 0.0), b(n, 0.0), c(n, 0.0);
    int row = 0;
    for (size_t i = 0; i < pos_t.size(); ++i, ++row) {
        double tp = 1;
        for (int j = 0; j < n; ++j) { A[row * n + j] = tp; tp *= pos_t[i]; }
        b[row] = pos_v[i];
    }
    for (size_t i = 0; i < vel_t.size(); ++i, ++row) {
        double tp = 1;
        for (int j = 1; j < n; ++j) { A[row * n + j] = j * tp; tp *= vel_t[i]; }
        b[row] = vel_v[i];
    }
    for (size_t i = 0; i < acc_t.size(); ++i, ++row) {
        double tp = 1;
        for (int j = 2; j < n; ++j) { A[row * n + j] = j * (j - 1) * tp; tp *= acc_t[i]; }
        b[row] = acc_v[i];
    }
    qr_colpiv_solve(n, A.data(), b.data(), c.data());
    return c;
}

double polyval(const std::vector<double>& poly, double x) {
    double value = 0, xp = 1;
    for (double ci : poly) { value += ci * xp; xp *= x; }
    return value;
}

std::vector<double> polyder(const std::vector<double>& poly) {
    if (poly.size() <= 1) return std::vector<double>(1, 0.0);
    std::vector<double> d(poly.size() - 1);
    for (size_t i = 0; i + 1 < poly.size(); ++i) d[i] = (double)(i + 1) * poly[i + 1];
    return d;
}

void swapBaseVelocityAndRefToWorldFrame(const Mat<6, 6>& X01, Vec<30>& v) {
    Vec<6> b;
    for (int i = 0; i < 3; ++i) { b[i] = v[3 + i]; b[3 + i] = v[i]; }
    Vec<6> r = X01 * b;
    for (int i = 0; i < 6; ++i) v[i] = r[i];
}

FootCoeff footCoeffTrajectory(const Vec<3>& cur, const Vec<3>& des, double stepHeight, double T) {
    FootCoeff c;
    std::vector<double> t2 = {0, T}, z2 = {0, 0}, t3 = {0, T / 2, T}, z3 = {0, 0, 0};
    c.x = findPolyCoeff(t2, {cur[0], des[0]}, t2, z2, t2, z2);
    c.y = findPolyCoeff(t2, {cur[1], des[1]}, t2, z2, t2, z2);
    c.z = findPolyCoeff(t3, {cur[2], stepHeight, des[2]}, t3, z3, t2, z2);
    return c;
}

ZMP::ZMP(double simulationTime, double timeStep, SupportFoot foot) {
    int samples = static_cast<int>((simulationTime + 0.5) / timeStep);
    zmpXRef.assign(samples, 0.0);
    double y = (foot == SupportFoot::Right) ? -0.05 : (foot == SupportFoot::Left ? 0.05 : 0.0);
    zmpYRef.assign(samples, y);
}

// -------------------------------------------------------------------- Robot
const int kParent[NUM_FRAMES] = {-1, 0, 1, 2, 3, 4, 5, 6, 0, 8, 9, 10, 11, 12, 13,
                                 0, 15, 16, 17, 18, 0, 20, 21, 22, 23, 0, 25, 26};
const int kActuated[NUM_FRAMES] = {0, 1, 2, 3, 4, 5, 6, 0, 7, 8, 9, 10, 11, 12, 0,
                                   13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 0};

Vec<30> initialConfiguration() {
    static const double q0[30] = {-0.0185, 0, 0.282, 0, 0, 0, 0, 0, -0.5, 0.8, -0.3, 0, 0, 0, -0.5,
                                  0.8, -0.3, 0, 1.6, 0, 0, 0, 0, -1.6, 0, 0, 0, 0, 0, 0};
    Vec<30> q;
    for (int i = 0; i < 30; ++i) q[i] = q0[i];
    return q;
}

void matTrans(const double theta[25], Mat<4, 4> T[25]) {
    const double r1 = -0.07071, r7 = 0.07071, r15 = 0.105, r17 = 0.05595, r20 = 0.105, r22 = 0.05595;
    const double d4 = -0.1, d5 = -0.1029, d10 = -0.1, d11 = -0.1029, d15 = -0.015, d20 = -0.015, d25 = 0.030;
    const double h = kPi / 2;
    const double r[25] = {r1, 0, 0, 0, 0, 0, r7, 0, 0, 0, 0, 0, 0, 0, r15, 0, r17, 0, 0, r20, 0, r22, 0, 0, 0};
    const double d[25] = {0, 0, 0, d4, d5, 0, 0, 0, 0, d10, d11, 0, 0, 0, d15, 0, 0, 0, 0, d20, 0, 0, 0, 0, d25};
    const double al[25] = {0, h, h, 0, 0, -h, -h, -h, h, 0, 0, -h, -h, h, h, -h, h, h, h, h, -h, h, 0, -h, 0};
    for (int i = 0; i < 25; ++i) {
        double ct = std::cos(theta[i]), st = std::sin(theta[i]);
        double ca = std::cos(al[i]), sa = std::sin(al[i]);
        Mat<4, 4>& M = T[i];
        M(0, 0) = ct;      M(0, 1) = -st;     M(0, 2) = 0;   M(0, 3) = d[i];
        M(1, 0) = ca * st; M(1, 1) = ca * ct; M(1, 2) = -sa; M(1, 3) = -r[i] * sa;
        M(2, 0) = sa * st; M(2, 1) = sa * ct; M(2, 2) = ca;  M(2, 3) = r[i] * ca;
        M(3, 0) = 0;       M(3, 1) = 0;       M(3, 2) = 0;   M(3, 3) = 1;
    }
}

Robot::Robot() {
    nao_link_table(links);
    updateState(Vec<30>());
    mass = 0;
    for (int i = 0; i < NUM_FRAMES; ++i) {  // Aldebaran (world-aligned at q=0) -> joint frame
        Mat<3, 3> Rj = T[i].block<3, 3>(0, 0);
        links[i].com = Rj.t() * links[i].com;
        links[i].inertia = (Rj.t() * links[i].inertia) * Rj;
        mass += links[i].mass;
    }
    updateState(initialConfiguration());
    double rf[9] = {0, 0, 1, 0, -1, 0, 1, 0, 0};
    for (int k = 0; k < 9; ++k) Rf_q0.d[k] = rf[k];
    Lf_q0 = Rf_q0;
    v = Vec<30>();
    S[2] = 1.0;
    footVertices[0] = v3(0.1, 0.025, 0);
    footVertices[1] = v3(0.1, -0.025, 0);
    footVertices[2] = v3(-0.05, 0.025, 0);
    footVertices[3] = v3(-0.05, -0.025, 0);
}

static Mat<4, 4> fixedOffset(double x) {
    Mat<4, 4> M = Mat<4, 4>::identity();
    M(0, 3) = x;
    return M;
}

void Robot::forwardKinematics() {
    T[0] = Mat<4, 4>();
    T[0].set(0, 3, v3(q[0], q[1], q[2]));
    T[0].set(0, 0, eulerAnglesToSO3(v3(q[3], q[4], q[5])));
    T[0](3, 3) = 1.0;

    double theta[25];
    for (int t = 0; t < 24; ++t) theta[t] = q[t + BASEDOF];
    theta[1] += (3.0 / 4) * kPi;
    theta[6] -= (1.0 / 2) * kPi;
    theta[7] += (1.0 / 4) * kPi;
    theta[13] += (1.0 / 2) * kPi;
    theta[18] += (1.0 / 2) * kPi;
    theta[23] -= (1.0 / 2) * kPi;
    theta[24] = -kPi / 2;
    Mat<4, 4> L[25];
    matTrans(theta, L);

    const double a = 0.7071;  // sic: not sqrt(2)/2 (src/Robot.cpp:92-103)
    Mat<4, 4> aux01, aux09;
    double m01[16] = {0, -1, 0, 0, a, 0, a, 0, -a, 0, a, 0, 0, 0, 0, 1};
    double m09[16] = {1, 0, 0, 0, 0, a, a, 0, 0, -a, a, 0, 0, 0, 0, 1};
    for (int k = 0; k < 16; ++k) { aux01.d[k] = m01[k]; aux09.d[k] = m09[k]; }
    Mat<4, 4> sole = fixedOffset(-0.0452);

    T[1] = (T[0] * aux01) * L[0];
    for (int i = 1; i < 6; ++i) T[i + 1] = T[i] * L[i];
    T[7] = T[6] * sole;
    T[8] = (T[0] * aux09) * L[6];
    for (int i = 8; i < 13; ++i) T[i + 1] = T[i] * L[i - 1];
    T[14] = T[13] * sole;

    auto attach = [&](int frame, int t, const Vec<3>& off) {
        Mat<4, 4> M = L[t];
        for (int k = 0; k < 3; ++k) M(k, 3) = M(k, 3) + off[k];
        T[frame] = T[0] * M;
    };
    attach(15, 12, v3(0, -0.098, 0.13591));
    for (int i = 15; i < 19; ++i) T[i + 1] = T[i] * L[i - 2];
    attach(20, 17, v3(0, 0.098, 0.13591));
    for (int i = 20; i < 24; ++i) T[i + 1] = T[i] * L[i - 2];
    attach(25, 22, v3(0, 0, 0.1615));
    for (int i = 25; i < 27; ++i) T[i + 1] = T[i] * L[i - 2];
}

void Robot::computeCoM() {
    Vec<3> com;
    for (int i = 0; i < NUM_FRAMES; ++i) {
        Vec<4> h;
        for (int k = 0; k < 3; ++k) h[k] = links[i].com[k];
        h[3] = 1.0;
        Vec<3> pw = T[i].block<3, 4>(0, 0) * h;
        com = com + links[i].mass * pw;
    }
    CoM = com / mass;
}

Vec<30> Robot::desiredPosture() const { return initialConfiguration(); }  // same literals, :253-262

void Robot::updateState(const Vec<30>& q_new) {
    q = q_new;
    forwardKinematics();
    computeCoM();
    // parentTransMatrix + allVelocityMatrices, :276-298
    X[0] = velocityMatrix(T[0]);
    for (int i = 1; i < NUM_FRAMES; ++i)
        X[i] = velocityMatrix(inverseTransformationMatrix(T[kParent[i]]) * T[i]);
}

void Robot::updateVelocityState(const Vec<30>& v_new, const Mat<6, 30>& AG) {
    v = v_new;
    computeComMomentum(v, AG);
}

void Robot::computeComMomentum(Vec<30> vv, const Mat<6, 30>& AG) {
    swapBaseVelocityAndRefToWorldFrame(X[0], vv);
    Vec<6> h = AG * vv;
    for (int k = 0; k < 3; ++k) {
        comVel[k] = h[3 + k] / mass;
        comAngMom[k] = h[k];
    }
}

// ----------------------------------------------------------------- Dynamics
Mat<6, 6> Dynamics::spatialInertiaMatrix(const LinkInertia& link) {
    Mat<6, 6> I;
    Mat<3, 3> cx = crossMatrix(link.com);
    I.set(0, 0, link.inertia - (link.mass * cx) * cx);
    I.set(0, 3, link.mass * cx);
    I.set(3, 0, (-link.mass) * cx);
    I.set(3, 3, link.mass * Mat<3, 3>::identity());
    return I;
}

void Dynamics::allSpatialInertiaMatrices(const Robot& robot) {
    I[0] = spatialInertiaMatrix(robot.links[0]);
    for (int i = 1; i < NUM_FRAMES; ++i)
        if (kActuated[i] != 0) I[i] = spatialInertiaMatrix(robot.links[i]);
}

void Dynamics::forwardNewtonEuler(const Robot& robot, const Vec<30>& qD, Vec<6> vel[], Vec<6> acc[], Vec<6> f[]) {
    for (int i = 1; i < NUM_FRAMES; ++i) {
        int p = kParent[i];
        if (kActuated[i] != 0) {
            double qd = qD[kActuated[i] + BASEDOF - 1];
            vel[i] = robot.X[i] * vel[p] + robot.S * qd;
            acc[i] = robot.X[i] * acc[p] + (spatialCrossMatrix(vel[i]) * robot.S) * qd;
            f[i] = I[i] * acc[i] + (spatialCrossMatrixForce(vel[i]) * I[i]) * vel[i];
        } else {
            vel[i] = robot.X[i] * vel[p];
            acc[i] = robot.X[i] * acc[p];
        }
    }
}

void Dynamics::backwardNewtonEuler(const Robot& robot, Vec<6> f[], Vec<30>& tau) {
    for (int i = NUM_FRAMES - 1; i > 0; --i)
        if (kActuated[i] != 0) {
            tau[kActuated[i] + BASEDOF - 1] = f[i][2];
            f[kParent[i]] = f[kParent[i]] + robot.X[i].t() * f[i];
        }
}

Vec<30> Dynamics::computeC(const Robot& robot, bool isGravity) {
    Vec<30> Cv;
    Vec<6> g;
    g[5] = isGravity * 9.81;
    Vec<30> qD = robot.v;  // NB: whatever velocity the Robot holds *now* (see Controller::standStep)
    swapBaseVelocityAndRefToWorldFrame(robot.X[0], qD);
    Vec<6> vel[NUM_FRAMES], acc[NUM_FRAMES], f[NUM_FRAMES];
    vel[0] = qD.block<6, 1>(0, 0);
    acc[0] = robot.X[0] * g;
    f[0] = I[0] * acc[0] + (spatialCrossMatrixForce(vel[0]) * I[0]) * vel[0];
    forwardNewtonEuler(robot, qD, vel, acc, f);
    backwardNewtonEuler(robot, f, Cv);
    for (int k = 0; k < 6; ++k) Cv[k] = f[0][k];
    return Cv;
}

void Dynamics::computeM(const Robot& robot) {
    M = Mat<30, 30>();
    Mat<24, 24> H;
    Mat<6, 24> F2;
    Mat<6, 6> Ic[NUM_FRAMES];
    for (int i = 0; i < NUM_FRAMES; ++i) Ic[i] = I[i];
    for (int i = NUM_FRAMES - 1; i >= 0; --i) {
        if (kActuated[i] == 0) continue;
        int a = kActuated[i] - 1, p = kParent[i];
        Ic[p] = Ic[p] + (robot.X[i].t() * Ic[i]) * robot.X[i];
        Vec<6> f = Ic[i] * robot.S;
        H(a, a) = f[2];
        int j = i;
        while (kParent[j] != 0) {
            f = robot.X[j].t() * f;
            j = kParent[j];
            H(kActuated[j] - 1, a) = f[2];
            H(a, kActuated[j] - 1) = f[2];
        }
        F2.set(0, a, robot.X[j].t() * f);
    }
    M.set(0, 0, Ic[0]);
    M.set(BASEDOF, BASEDOF, H);
    M.set(0, BASEDOF, F2);
    M.set(BASEDOF, 0, F2.t());
}

void Dynamics::centroidalMatrixAndBias(const Robot& robot) {
    Mat<3, 3> R01 = robot.T[0].block<3, 3>(0, 0);
    Mat<6, 6> Ic1 = M.block<6, 6>(0, 0);
    Mat<6, 24> F = M.block<6, 24>(0, 6);
    Vec<3> p1G = v3(Ic1(2, 4) / robot.mass, Ic1(0, 5) / robot.mass, Ic1(1, 3) / robot.mass);
    Mat<6, 6> X1G;
    X1G.set(0, 0, R01);
    X1G.set(3, 3, R01);
    X1G.set(0, 3, (-R01) * crossMatrix(p1G));
    AG.set(0, 0, X1G * Ic1);
    AG.set(0, 6, X1G * F);
    AGpqp = X1G * Cg.block<6, 1>(0, 0);
}

Vec<6> Dynamics::computeJpqpFrame(const Robot& robot, int frame) {
    Vec<30> qD = robot.v;
    swapBaseVelocityAndRefToWorldFrame(robot.X[0], qD);
    Vec<6> vel[NUM_FRAMES], acc[NUM_FRAMES], f[NUM_FRAMES];
    vel[0] = qD.block<6, 1>(0, 0);
    acc[0] = Vec<6>();
    f[0] = I[0] * acc[0] + (spatialCrossMatrixForce(vel[0]) * I[0]) * vel[0];
    forwardNewtonEuler(robot, qD, vel, acc, f);
    Mat<6, 6> Rot;
    Mat<3, 3> R = robot.T[frame].block<3, 3>(0, 0);
    Rot.set(0, 0, R);
    Rot.set(3, 3, R);
    return Rot * acc[frame];
}

void Dynamics::computeAll(const Robot& robot) {
    allSpatialInertiaMatrices(robot);
    C = computeC(robot, true);
    Cg = computeC(robot, false);
    computeM(robot);
    centroidalMatrixAndBias(robot);
    Jpqp.set(0, 0, computeJpqpFrame(robot, 7));
    Jpqp.set(6, 0, computeJpqpFrame(robot, 14));
}

// --------------------------------------------------------------- Kinematics
Mat<6, 30> Kinematics::frameJacobian(const Robot& robot, int frame) const {
    // sole_X_f accumulated from the sole towards the hip, one joint column per
    // chain frame, base block = sole_X_base (frame index walk of :120-146).
    Mat<6, 30> J;
    Mat<6, 6> Xn = robot.X[frame];
    int j = frame - 1;
    while (true) {
        J.set(0, kActuated[j] + BASEDOF - 1, Xn * robot.S);
        Xn = Xn * robot.X[j];
        if (kParent[j] == 0) break;
        --j;
    }
    J.set(0, 0, Xn * Mat<6, 6>::identity());
    return J;
}

Mat<12, 30> Kinematics::feetJacobian(const Robot& robot) const {
    Mat<12, 30> JF;
    const int soles[2] = {7, 14};
    for (int s = 0; s < 2; ++s) {
        Mat<6, 30> Jl = frameJacobian(robot, soles[s]);
        Mat<6, 6> R6;
        Mat<3, 3> R = robot.T[soles[s]].block<3, 3>(0, 0);
        R6.set(0, 0, R);
        R6.set(3, 3, R);
        JF.set(6 * s, 0, R6 * Jl);
    }
    return JF;
}

void Kinematics::computeAll(const Robot& robot) {
    JFeet = feetJacobian(robot);
    JR = JFeet.block<6, 30>(0, 0);
    JL = JFeet.block<6, 30>(6, 0);
}

Vec<3> Kinematics::rotMatrixToEulerAngles(const Mat<3, 3>& R, const Mat<3, 3>& ref) {
    Mat<3, 3> N = R * ref;
    Vec<3> eta;
    eta[2] = std::atan2(N(1, 0), N(0, 0));
    eta[1] = std::atan2(-N(2, 0), std::cos(eta[2]) * N(0, 0) + std::sin(eta[2]) * N(1, 0));
    eta[0] = std::atan2(std::sin(eta[2]) * N(0, 2) - std::cos(eta[2]) * N(1, 2),
                        -std::sin(eta[2]) * N(0, 1) + std::cos(eta[2]) * N(1, 1));
    return eta;
}

Vec<30> Kinematics::desiredOperationalState(const Robot& robot, const Vec<6>& Rf, const Vec<6>& Lf,
                                            const Vec<3>& com) const {
    Vec<30> Qd;
    Qd.set(0, 0, Rf);
    Qd.set(6, 0, Lf);
    for (int k = 0; k < 12; ++k) Qd[12 + k] = robot.q[12 + BASEDOF + k];
    Qd.set(27, 0, com);
    return Qd;
}

Vec<30> Kinematics::operationalState(const Robot& robot) const {
    Vec<30> Q;
    Q.set(0, 0, robot.T[7].block<3, 1>(0, 3));
    Q.set(3, 0, rotMatrixToEulerAngles(robot.T[7].block<3, 3>(0, 0), robot.Rf_q0));
    Q.set(6, 0, robot.T[14].block<3, 1>(0, 3));
    Q.set(9, 0, rotMatrixToEulerAngles(robot.T[14].block<3, 3>(0, 0), robot.Lf_q0));
    for (int k = 0; k < 12; ++k) Q[12 + k] = robot.q[12 + BASEDOF + k];
    for (int k = 0; k < 3; ++k) Q[24 + k] = robot.q[3 + k];
    Q.set(27, 0, robot.CoM);
    return Q;
}

Mat<3, 30> Kinematics::comJacobian(const Robot& robot) const {
    Mat<3, 30> J;
    Mat<3, 3> cm[NUM_FRAMES];
    Vec<3> p0 = robot.T[0].block<3, 1>(0, 3);
    for (int i = 0; i < NUM_FRAMES - 1; ++i) {
        cm[i] = crossMatrix(robot.T[i].block<3, 1>(0, 2));
        if (robot.links[i].mass == 0) continue;
        Vec<4> h;
        for (int k = 0; k < 3; ++k) h[k] = robot.links[i].com[k];
        h[3] = 1.0;
        Vec<3> pc = robot.T[i].block<3, 4>(0, 0) * h;
        Mat<3, 30> JX;
        JX.set(0, 0, Mat<3, 3>::identity());
        JX.set(0, 3, crossMatrix(p0 - pc));
        for (int j = i; j != 0; j = kParent[j])
            if (kActuated[j] != 0)
                JX.set(0, kActuated[j] + BASEDOF - 1, cm[j] * (pc - robot.T[j].block<3, 1>(0, 3)));
        J = J + robot.links[i].mass * JX;
    }
    J = J / robot.mass;
    Mat<3, 3> Om = matrixAngularVelToEulerDot(v3(robot.q[3], robot.q[4], robot.q[5]));
    J.set(0, 3, J.block<3, 3>(0, 3) * inverse3(Om));
    return J;
}

Mat<30, 30> Kinematics::jacInvKinematics(const Robot& robot) const {
    Mat<12, 30> Jf = feetJacobian(robot);
    // base columns [ang|lin] -> [lin|ang]; foot rows [ang;lin] -> [lin;ang]
    Mat<12, 3> tmpc = Jf.block<12, 3>(0, 0);
    Jf.set(0, 0, Jf.block<12, 3>(0, 3));
    Jf.set(0, 3, tmpc);
    for (int s = 0; s < 2; ++s) {
        Mat<3, 30> top = Jf.block<3, 30>(6 * s, 0);
        Jf.set(6 * s, 0, Jf.block<3, 30>(6 * s + 3, 0));
        Jf.set(6 * s + 3, 0, top);
    }
    Mat<3, 3> Om = matrixAngularVelToEulerDot(v3(robot.q[3], robot.q[4], robot.q[5]));
    for (int r = 0; r < 12; r += 3) Jf.set(r, 3, Jf.block<3, 3>(r, 3) * inverse3(Om));
    const int soles[2] = {7, 14};
    for (int s = 0; s < 2; ++s) {
        Vec<3> etaF = rotMatrixToEulerAngles(robot.T[soles[s]].block<3, 3>(0, 0), robot.Rf_q0);
        Mat<3, 3> OmF = matrixAngularVelToEulerDot(etaF);
        Jf.set(6 * s + 3, 3, OmF * Jf.block<3, 3>(6 * s + 3, 3));
    }
    Mat<30, 30> J;
    J.set(0, 0, Jf);
    for (int k = 0; k < 12; ++k) J(12 + k, 12 + BASEDOF + k) = 1.0;
    for (int k = 0; k < 3; ++k) J(24 + k, 3 + k) = 1.0;
    J.set(27, 0, comJacobian(robot));
    return J;
}

int Kinematics::compute(Robot& robot, const Vec<30>& desOp, int maxIter) {
    Vec<30> q = robot.q;
    Vec<30> e = desOp - operationalState(robot);
    int iter = 0;
    while (max_abs(e) > 1e-10 && iter < maxIter) {
        q = q + qr_solve(jacInvKinematics(robot), e);
        robot.updateState(q);
        e = desOp - operationalState(robot);
        ++iter;  // the reference never increments (unbounded loop); we cap
    }
    return iter;
}

// ---------------------------------------------------------------------- MPC
Mpc3dLip::Mpc3dLip(double dt_, double timeHorizon_, double zCom_, double alpha_, double beta_)
    : dt(dt_), timeHorizon(timeHorizon_), zCom(zCom_), alpha(alpha_), beta(beta_) {
    initialize();
}

void Mpc3dLip::initialize() {
    horizon = static_cast<int>(timeHorizon / dt);
    int n = horizon + 1;
    A[0][0] = 1; A[0][1] = dt; A[1][0] = 0; A[1][1] = 1;
    B[0] = (dt * dt) / 2; B[1] = dt;
    Cc[0] = 1; Cc[1] = 0;
    D = -zCom / gravity;
    Px.assign(n * 2, 0.0);
    Pu.assign(n * n, 0.0);
    double Ap[2][2] = {{1, 0}, {0, 1}};
    auto mulA = [&](double P[2][2]) {
        double R[2][2];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) R[i][j] = P[i][0] * A[0][j] + P[i][1] * A[1][j];
        std::memcpy(P, R, sizeof(R));
    };
    Px[0] = Cc[0]; Px[1] = Cc[1];
    Pu[0] = D;
    for (int i = 1; i <= horizon; ++i) {
        mulA(Ap);
        Px[i * 2 + 0] = Cc[0] * Ap[0][0] + Cc[1] * Ap[1][0];
        Px[i * 2 + 1] = Cc[0] * Ap[0][1] + Cc[1] * Ap[1][1];
        Pu[i * n + (i - 1)] = Cc[0] * B[0] + Cc[1] * B[1];
        Pu[i * n + i] = D;
        double Aj[2][2] = {{1, 0}, {0, 1}};
        for (int j = 1; j <= horizon - i; ++j) {
            mulA(Aj);
            double ca0 = Cc[0] * Aj[0][0] + Cc[1] * Aj[1][0], ca1 = Cc[0] * Aj[0][1] + Cc[1] * Aj[1][1];
            Pu[(i + j) * n + (i - 1)] = ca0 * B[0] + ca1 * B[1];
        }
    }
    Hchol.clear();
}

double Mpc3dLip::solve1D(const std::vector<double>& gvec) {
    int n = horizon + 1;
    if (Hchol.empty()) {
        Hchol.assign(n * n, 0.0);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double s = 0;
                for (int k = 0; k < n; ++k) s += Pu[k * n + i] * Pu[k * n + j];
                Hchol[i * n + j] = beta * s + (i == j ? alpha : 0.0);
            }
        bool ok = chol_factor<double>(n, Hchol);
        assert(ok);
        (void)ok;
    }
    std::vector<double> u(n);
    for (int i = 0; i < n; ++i) u[i] = -gvec[i];
    chol_solve<double>(n, Hchol, u.data());
    return u[0];
}

void Mpc3dLip::compute(const double posCom[2], const double velCom[2], const std::vector<double>& zx,
                       const std::vector<double>& zy, double t) {
    int n = horizon + 1;
    int k = static_cast<int>(t / dt);
    last_k = k;
    if (rebuild_hessian_each_call) Hchol.clear();  // reference cost model: H rebuilt per call (:89-90)
    if (k < 0 || (size_t)(k + n) > zx.size() || (size_t)(k + n) > zy.size()) {
        std::fprintf(stderr, "orc::Mpc3dLip: ZMP window [%d,%d) outside reference of %zu samples\n", k,
                     k + n, zx.size());
        std::abort();
    }
    double st[2][2] = {{posCom[0], velCom[0]}, {posCom[1], velCom[1]}};
    double acc[2];
    for (int ax = 0; ax < 2; ++ax) {
        const std::vector<double>& z = ax == 0 ? zx : zy;
        std::vector<double> e(n), gv(n);
        for (int i = 0; i < n; ++i) e[i] = (Px[i * 2] * st[ax][0] + Px[i * 2 + 1] * st[ax][1]) - z[k + i];
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int r = 0; r < n; ++r) s += (beta * Pu[r * n + i]) * e[r];
            gv[i] = s;
        }
        acc[ax] = solve1D(gv);
    }
    for (int ax = 0; ax < 2; ++ax) {
        double p = A[0][0] * st[ax][0] + A[0][1] * st[ax][1] + B[0] * acc[ax];
        double v = A[1][0] * st[ax][0] + A[1][1] * st[ax][1] + B[1] * acc[ax];
        Vec<3>& ref = ax == 0 ? xRef : yRef;
        ref[0] = p; ref[1] = v; ref[2] = acc[ax];
    }
}

// --------------------------------------------------------------- Controller
Controller::Controller(Robot& r, Mpc3dLip& m, const ZMP& z, const FootCoeff& rFc, const FootCoeff& lFc)
    : robot(r), mpc(m), zmp(z), rF(rFc), lF(lFc) {
    const double col[4][3] = {{mu, 0, 1}, {0, mu, 1}, {-mu, 0, 1}, {0, -mu, 1}};
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 3; ++k) frictionMatrix(k, c) = col[c][k];
}

void Controller::standStep(const ControllerInput& in) {
    robot.updateState(in.q);
    dyn.computeAll(robot);  // still sees the velocity of the PREVIOUS call: v_ is refreshed two lines below
    kin.computeAll(robot);
    robot.updateVelocityState(in.dq, dyn.AG);
    double c2[2] = {robot.CoM[0], robot.CoM[1]}, v2[2] = {robot.comVel[0], robot.comVel[1]};
    mpc.compute(c2, v2, zmp.zmpXRef, zmp.zmpYRef, in.time);
    WBCOutput out = WBC(in.time);
    tau = out.tau;
}

Vec<30> Controller::PDJointsAcc() const {
    Vec<30> r = KpJoints * (robot.desiredPosture() - robot.q) + KdJoints * (Vec<30>() - robot.v);
    for (int k = 0; k < 3; ++k) std::swap(r[k], r[3 + k]);
    return r;
}

Vec<6> Controller::PDMomentumAcc() const {
    Vec<3> pr = v3(mpc.xRef[0], mpc.yRef[0], mpc.zCom);
    Vec<3> vr = v3(mpc.xRef[1], mpc.yRef[1], 0);
    Vec<3> ar = v3(mpc.xRef[2], mpc.yRef[2], 0);
    Vec<3> lin = robot.mass * (KpMom * (pr - robot.CoM) + KdMom * (vr - robot.comVel) + ar);
    Vec<3> ang = KdMom * (Vec<3>() - robot.comAngMom);
    Vec<6> h;
    h.set(0, 0, ang);
    h.set(3, 0, lin);
    return h;
}

int Controller::phase_of(double t) const {
    if (contact.empty()) return 0;
    int p = static_cast<int>(t / phase_T);
    int n = (int)contact.size();
    return p < 0 ? 0 : (p >= n ? n - 1 : p);
}

Vec<12> Controller::PDFeetAcc(double t_abs) const {
    const int ph = phase_of(t_abs);
    const double t = contact.empty() ? t_abs : t_abs - ph * phase_T;
    Vec<30> v = robot.v;
    swapBaseVelocityAndRefToWorldFrame(robot.X[0], v);
    Vec<12> out;
    const int soles[2] = {7, 14};
    for (int s = 0; s < 2; ++s) {
        const FootCoeff& c = contact.empty() ? (s == 0 ? rF : lF) : (s == 0 ? rFp[ph] : lFp[ph]);
        const Mat<6, 30>& J = s == 0 ? kin.JR : kin.JL;
        const Mat<3, 3>& Rdes = s == 0 ? robot.Rf_q0 : robot.Lf_q0;
        Vec<3> pos = robot.T[soles[s]].block<3, 1>(0, 3);
        Vec<6> vel = J * v;
        Mat<3, 3> err = Rdes.t() * robot.T[soles[s]].block<3, 3>(0, 0);
        Vec<3> eo = (-Rdes) * rotMatToAxisAngle(err);
        Vec<3> pr = v3(polyval(c.x, t), polyval(c.y, t), polyval(c.z, t));
        Vec<6> vr, ar, pe;
        vr[3] = polyval(polyder(c.x), t); vr[4] = polyval(polyder(c.y), t); vr[5] = polyval(polyder(c.z), t);
        ar[3] = polyval(polyder(polyder(c.x)), t);
        ar[4] = polyval(polyder(polyder(c.y)), t);
        ar[5] = polyval(polyder(polyder(c.z)), t);
        pe.set(0, 0, eo);
        pe.set(3, 0, pr - pos);
        out.set(6 * s, 0, KpFeet * pe + KdFeet * (vr - vel) + ar);
    }
    return out;
}

void Controller::frictionConstraints(Mat<12, 74>& Aeq, Vec<12>& beq, Mat<32, 74>& Aineq, Vec<32>& bineq) const {
    Aeq = Mat<12, 74>(); Aineq = Mat<32, 74>(); beq = Vec<12>(); bineq = Vec<32>();
    // variable offsets: n_R 30, f_R 33, n_L 36, f_L 39, c_R 42.., c_L 58..   (:170-175)
    // row offsets: f_R 0, n_R 3, f_L 6, n_L 9                                 (:177-180)
    for (int foot = 0; foot < 2; ++foot) {
        int idx_n = 30 + 6 * foot, idx_f = idx_n + 3, idx_c = 42 + 16 * foot;
        int row_f = 6 * foot, row_n = row_f + 3;
        for (int k = 0; k < 3; ++k) {
            Aeq(row_f + k, idx_f + k) = -1;
            Aeq(row_n + k, idx_n + k) = -1;
        }
        for (int vtx = 0; vtx < 4; ++vtx) {
            Mat<3, 4> mom = crossMatrix(robot.footVertices[vtx]) * frictionMatrix;  // same vertices for both feet
            for (int k = 0; k < 3; ++k)
                for (int c = 0; c < 4; ++c) {
                    Aeq(row_f + k, idx_c + 4 * vtx + c) = frictionMatrix(k, c);
                    Aeq(row_n + k, idx_c + 4 * vtx + c) = mom(k, c);
                }
        }
    }
    for (int i = 0; i < 32; ++i) Aineq(i, 42 + i) = 1.0;
}

Vec<74> Controller::solveQP(const Mat<74, 74>& H, const Vec<74>& g) {
    Hqp = 0.5 * (H + H.t());
    gqp = g;
    Aqp = Mat<50, 74>();
    Mat<30, 12> JT = kin.JFeet.t();
    for (int r = 0; r < 6; ++r) {
        for (int c = 0; c < 30; ++c) Aqp(r, c) = dyn.M(r, c);
        for (int c = 0; c < 12; ++c) Aqp(r, 30 + c) = -JT(r, c);
        lbA[r] = ubA[r] = -dyn.C[r];
    }
    Mat<12, 74> Aeq; Vec<12> beq; Mat<32, 74> Ain; Vec<32> bin;
    frictionConstraints(Aeq, beq, Ain, bin);
    Aqp.set(6, 0, Aeq);
    Aqp.set(18, 0, Ain);
    for (int r = 0; r < 12; ++r) lbA[6 + r] = ubA[6 + r] = beq[r];
    for (int r = 0; r < 32; ++r) { lbA[18 + r] = bin[r]; ubA[18 + r] = QP_INFTY; }
    if (!contact.empty()) {   // extension: a foot that is not in contact carries no force
        int mask = contact[last_phase];
        for (int foot = 0; foot < 2; ++foot)
            if (!((mask >> foot) & 1))
                for (int r = 0; r < 16; ++r) ubA[18 + 16 * foot + r] = 0.0;
    }
    if (!all_finite(Hqp) || !all_finite(gqp) || !all_finite(Aqp) || !all_finite(lbA)) {
        if (abort_on_nonfinite) {
            std::fprintf(stderr, "orc::Controller::solveQP: NaN or Inf in QP data\n");
            std::abort();  // :448-466
        }
        Vec<74> bad;
        for (int i = 0; i < 74; ++i) bad[i] = std::nan("");
        qpInfo = QpInfo();
        qpInfo.status = 4;
        ++qp_solves;
        ++qp_failures;
        return bad;
    }
    Vec<74> x;
    if (qp_long_double)
        qpInfo = qp_solve_gi<long double>(74, Hqp.d, gqp.d, 50, Aqp.d, lbA.d, ubA.d, x.d, yqp.d, 500);
    else
        qpInfo = qp_solve_gi<double>(74, Hqp.d, gqp.d, 50, Aqp.d, lbA.d, ubA.d, x.d, yqp.d, 500);
    ++qp_solves;
    qp_iterations += qpInfo.iterations;
    if (qpInfo.status != 0) ++qp_failures;  // reference logs and carries on, :472-478
    return x;
}

WBCOutput Controller::WBC(double t) {
    const int n = 30;
    last_phase = phase_of(t);
    qppRef = PDJointsAcc();
    hGpRef = PDMomentumAcc();
    footAccRef = PDFeetAcc(t);
    Mat<30, 30> WJ;
    for (int k = 0; k < 3; ++k) { WJ(k, k) = wBasePos; WJ(3 + k, 3 + k) = wBaseAng; }
    for (int k = 6; k < 30; ++k) WJ(k, k) = wJoints;
    Mat<6, 6> WC;
    for (int k = 0; k < 3; ++k) { WC(k, k) = wCoMK; WC(3 + k, 3 + k) = wCoML; }
    Mat<12, 12> WFeet = wFoot * Mat<12, 12>::identity();

    Mat<30, 6> AGt = dyn.AG.t();
    Mat<30, 12> Jt = kin.JFeet.t();
    Mat<74, 74> H;
    H.set(0, 0, (AGt * WC) * dyn.AG + WJ + (Jt * WFeet) * kin.JFeet);
    for (int k = 0; k < 12; ++k) H(n + k, n + k) = wForce;
    for (int k = 0; k < 32; ++k) H(42 + k, 42 + k) = 1e-8;
    Vec<74> g;
    g.set(0, 0, (AGt * WC) * dyn.AGpqp - (AGt * WC) * hGpRef - WJ * qppRef + (Jt * WFeet) * dyn.Jpqp -
                    (Jt * WFeet) * footAccRef);
    xqp = solveQP(H, g);

    WBCOutput out;
    out.f = xqp.block<12, 1>(n, 0);
    Vec<30> a = xqp.block<30, 1>(0, 0);
    Vec<30> gen = dyn.M * a + dyn.C - Jt * out.f;
    out.tau = gen.block<24, 1>(6, 0);
    Vec<6> aw = qr_solve(robot.X[0], a.block<6, 1>(0, 0));
    for (int k = 0; k < 3; ++k) { out.qpp[k] = aw[3 + k]; out.qpp[3 + k] = aw[k]; }
    for (int k = 6; k < 30; ++k) out.qpp[k] = a[k];
    return out;
}

// ------------------------------------------------------------ offline app
Vec<60> dynamics(const Vec<60>& state, double t, Robot& robot, Controller& controller, bool second_wbc,
                 WBCOutput* outp) {
    ControllerInput in;
    in.q = state.block<30, 1>(0, 0);
    in.dq = state.block<30, 1>(30, 0);
    in.time = t;
    controller.standStep(in);
    WBCOutput out;
    if (second_wbc) {
        out = controller.WBC(t);
    } else {  // the first solve inside standStep had identical inputs; rebuild the outputs from it
        Vec<30> a = controller.xqp.block<30, 1>(0, 0);
        out.f = controller.xqp.block<12, 1>(30, 0);
        out.tau = controller.tau;
        Vec<6> aw = qr_solve(robot.X[0], a.block<6, 1>(0, 0));
        for (int k = 0; k < 3; ++k) { out.qpp[k] = aw[3 + k]; out.qpp[3 + k] = aw[k]; }
        for (int k = 6; k < 30; ++k) out.qpp[k] = a[k];
    }
    if (outp) *outp = out;
    Vec<3> p = in.q.block<3, 1>(0, 0), eta = in.q.block<3, 1>(3, 0);
    Vec<3> vl = in.dq.block<3, 1>(0, 0), w = in.dq.block<3, 1>(3, 0);
    Vec<30> qD = in.dq;
    qD.set(0, 0, vl + crossMatrix(w) * p);
    qD.set(3, 0, matrixAngularVelToEulerDot(eta) * w);
    Vec<60> xp;
    xp.set(0, 0, qD);
    xp.set(30, 0, out.qpp);
    return xp;
}

Vec<60> rk4Step(const Vec<60>& x, double t, double dt, Robot& robot, Controller& controller, bool second_wbc) {
    Vec<60> k1 = dynamics(x, t, robot, controller, second_wbc);
    Vec<60> k2 = dynamics(x + (0.5 * dt) * k1, t + 0.5 * dt, robot, controller, second_wbc);
    Vec<60> k3 = dynamics(x + (0.5 * dt) * k2, t + 0.5 * dt, robot, controller, second_wbc);
    Vec<60> k4 = dynamics(x + dt * k3, t + dt, robot, controller, second_wbc);
    return x + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
}

}  // namespace orc
