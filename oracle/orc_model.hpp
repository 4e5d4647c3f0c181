// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// CPU restatement of the Ema158/linearMpcHumanoid closed-loop hot path
// (single environment, FP64), class-for-class with the reference so parity
// tests read like the reference's own code would.  Each member cites the
// reference file:line it follows (paths relative to /root/reference).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, load or call this.
//
// PARITY PINNING: the reference ships no tests/golden vectors.  Everything but
// the QP is pinned against the reference's own sources compiled here with
// shim headers (oracle/refbuild -> oracle/_ref, see oracle/README.md).  The
// QP solver boundary (qpOASES, absent) is PARITY UNPINNED, see orc_qp.hpp.
#pragma once
#include "orc_linalg.hpp"
#include "orc_qp.hpp"
#include <vector>

namespace orc {

constexpr int NUM_JOINTS = 30;         // include/.../robotInfo/Robot.hpp:11
constexpr int NUM_ACTUAL_JOINTS = 24;  // :12
constexpr int NUM_FRAMES = 28;         // :13
constexpr int BASEDOF = 6;

struct LinkInertia {  // include/.../robotInfo/linkInertia.hpp:4-9
    double mass;
    Vec<3> com;
    Mat<3, 3> inertia;
};
void nao_link_table(LinkInertia out[NUM_FRAMES]);  // src/robotParameters.cpp:8-229

// ---- src/generalizedFunctions.cpp ----
Mat<3, 3> crossMatrix(const Vec<3>& v);                           // :3-9
Mat<6, 6> velocityMatrix(const Mat<4, 4>& T);                     // :11-19
Mat<4, 4> inverseTransformationMatrix(const Mat<4, 4>& T);        // :21-27
Mat<6, 6> spatialCrossMatrix(const Vec<6>& v);                    // :29-35
Mat<6, 6> spatialCrossMatrixForce(const Vec<6>& v);               // :37-41
Mat<3, 3> matrixAngularVelToEulerDot(const Vec<3>& eta);          // :43-50
Mat<3, 3> eulerAnglesToSO3(const Vec<3>& rpy);                    // :52-72
Vec<3> rotMatToAxisAngle(const Mat<3, 3>& R);                     // :74-101
std::vector<double> findPolyCoeff(const std::vector<double>& pos_t, const std::vector<double>& pos_v,
                                  const std::vector<double>& vel_t, const std::vector<double>& vel_v,
                                  const std::vector<double>& acc_t, const std::vector<double>& acc_v);  // :103-163
double polyval(const std::vector<double>& poly, double x);        // :165-177
std::vector<double> polyder(const std::vector<double>& poly);     // :179-193
void swapBaseVelocityAndRefToWorldFrame(const Mat<6, 6>& X01, Vec<30>& v);  // :195-206

// ---- src/footRefTrajectory.cpp:4-47 ----
struct FootCoeff { std::vector<double> x, y, z; };
FootCoeff footCoeffTrajectory(const Vec<3>& currentPos, const Vec<3>& desPos, double stepHeight, double T);

// ---- src/zmpGeneration.cpp ----
enum class SupportFoot { Right, Left, Double };  // include/.../general/Task.hpp:9-13
struct ZMP {
    std::vector<double> zmpXRef, zmpYRef;
    ZMP() {}
    ZMP(double simulationTime, double timeStep, SupportFoot foot);  // :12-23 + stanceZMP :39-60
};

// ---- src/Robot.cpp ----
extern const int kParent[NUM_FRAMES];    // :162-167
extern const int kActuated[NUM_FRAMES];  // :169-174
Vec<30> initialConfiguration();          // :242-251
void matTrans(const double theta[25], Mat<4, 4> T[25]);  // :176-223

struct Robot {
    Vec<30> q, v;               // v_ is only refreshed by updateVelocityState
    Mat<4, 4> T[NUM_FRAMES];
    Mat<6, 6> X[NUM_FRAMES];
    LinkInertia links[NUM_FRAMES];
    Vec<3> CoM, comVel, comAngMom;
    double mass = 0;
    Mat<3, 3> Rf_q0, Lf_q0;
    Vec<6> S;
    Vec<3> footVertices[4];

    Robot();                                             // :5-43
    void forwardKinematics();                            // :45-160
    void computeCoM();                                   // :225-238
    Vec<30> desiredPosture() const;                      // :253-262
    void updateState(const Vec<30>& q_new);              // :264-269
    void updateVelocityState(const Vec<30>& v_new, const Mat<6, 30>& AG);  // :271-274
    void computeComMomentum(Vec<30> vv, const Mat<6, 30>& AG);             // :300-310
};

// ---- src/Dynamics.cpp ----
struct Dynamics {
    Vec<30> C, Cg;
    Mat<30, 30> M;
    Mat<6, 30> AG;
    Vec<6> AGpqp;
    Vec<12> Jpqp;
    Mat<6, 6> I[NUM_FRAMES];
    void computeAll(const Robot& robot);                                    // :202-216
    static Mat<6, 6> spatialInertiaMatrix(const LinkInertia& link);         // :4-13
    void allSpatialInertiaMatrices(const Robot& robot);                     // :15-27
    Vec<30> computeC(const Robot& robot, bool isGravity);                   // :29-60
    void computeM(const Robot& robot);                                      // :62-101
    void centroidalMatrixAndBias(const Robot& robot);                       // :103-121
    void forwardNewtonEuler(const Robot& robot, const Vec<30>& qD, Vec<6> vel[], Vec<6> acc[],
                            Vec<6> f[]);                                    // :124-146
    void backwardNewtonEuler(const Robot& robot, Vec<6> f[], Vec<30>& tau); // :148-163
    Vec<6> computeJpqpFrame(const Robot& robot, int frame);                 // :165-200
};

// ---- src/invKinematics.cpp ----
struct Kinematics {
    Mat<6, 30> JR, JL;
    Mat<12, 30> JFeet;
    void computeAll(const Robot& robot);                                    // :271-276
    Vec<30> desiredOperationalState(const Robot& robot, const Vec<6>& Rf, const Vec<6>& Lf,
                                    const Vec<3>& com) const;               // :11-25
    int compute(Robot& robot, const Vec<30>& desOp, int maxIter = 200);     // :27-52 (+ iteration cap)
    Vec<30> operationalState(const Robot& robot) const;                     // :54-70
    Mat<12, 30> feetJacobian(const Robot& robot) const;                     // :72-107
    Mat<6, 30> frameJacobian(const Robot& robot, int frame) const;          // :109-149
    Mat<30, 30> jacInvKinematics(const Robot& robot) const;                 // :151-206
    Mat<3, 30> comJacobian(const Robot& robot) const;                       // :208-245
    static Vec<3> rotMatrixToEulerAngles(const Mat<3, 3>& R, const Mat<3, 3>& ref);  // :257-267
};

// ---- src/mpcLinearPendulum.cpp ----
struct Mpc3dLip {
    Vec<3> xRef, yRef;
    std::vector<double> Px, Pu;  // (N+1)x2, (N+1)x(N+1) row-major
    std::vector<double> Hchol;   // factor of alpha I + beta Pu'Pu, cached (H is constant; the
                                 // reference rebuilds it every call, :89-90)
    double A[2][2], B[2], Cc[2], D;
    double dt = 0.01, timeHorizon = 0.5, zCom = 0.26, gravity = 9.81, alpha = 1e-3, beta = 1;
    int horizon = 50;
    int last_k = 0;
    bool rebuild_hessian_each_call = false;  // true = cost model of the reference as written
    Mpc3dLip() { initialize(); }
    Mpc3dLip(double dt_, double timeHorizon_, double zCom_, double alpha_ = 1e-3, double beta_ = 1);  // :10-39
    void initialize();                                                       // :41-76
    void compute(const double posCom[2], const double velCom[2], const std::vector<double>& zmpXRef,
                 const std::vector<double>& zmpYRef, double t);              // :78-109
    double solve1D(const std::vector<double>& gvec);                         // :111-133
};

// ---- src/controller.cpp ----
struct WBCOutput {  // include/.../controller/controller.hpp:42-48
    Vec<30> qpp;
    Vec<24> tau;
    Vec<12> f;
};
struct ControllerInput { Vec<30> q, dq; double time; };  // controller.hpp:33-37

struct Controller {
    Robot& robot;
    Mpc3dLip& mpc;
    ZMP zmp;
    FootCoeff rF, lF;
    Dynamics dyn;
    Kinematics kin;
    Vec<24> tau;
    Mat<3, 4> frictionMatrix;
    double mu = 0.7;
    // gains / weights, controller.hpp:100-124
    double KpJoints = 300, KdJoints = 34, KpMom = 10, KdMom = 6.32, KpFeet = 500, KdFeet = 44;
    double wCoML = 4000, wCoMK = 0, wBasePos = 10, wBaseAng = 10, wJoints = 1, wForce = 1, wFoot = 100000;
    // last-solve artefacts, exposed for tests
    Mat<74, 74> Hqp;
    Vec<74> gqp, xqp;
    Mat<50, 74> Aqp;
    Vec<50> lbA, ubA, yqp;
    Vec<30> qppRef;
    Vec<6> hGpRef;
    Vec<12> footAccRef;
    QpInfo qpInfo;
    bool qp_long_double = false;
    bool abort_on_nonfinite = true;   // upstream aborts (controller.cpp:448-466); batch rollouts with falling robots
                                      // turn the abort into a failed solve that returns NaN (the step is then unhealthy)
    // EXTENSION (not in the reference): walking plans.  Time is cut into n phases of length
    // phase_T; phase p = min((int)(t/phase_T), n-1) selects the foot polynomials (evaluated at t - p*phase_T) and
    // which feet may carry force (bit 0 right, bit 1 left; a swing foot's 16 coefficients get bounds [0,0]).
    // The stand scenario is one phase of infinite length with both feet in contact: identical to upstream.
    std::vector<FootCoeff> rFp, lFp;
    std::vector<int> contact;
    double phase_T = 1e30;
    int last_phase = 0;
    int phase_of(double t) const;
    long qp_solves = 0, qp_iterations = 0, qp_failures = 0;

    Controller(Robot& r, Mpc3dLip& m, const ZMP& z, const FootCoeff& rFc, const FootCoeff& lFc);  // :5-46
    void standStep(const ControllerInput& in);                // :48-79
    WBCOutput WBC(double t);                                  // :81-154
    void frictionConstraints(Mat<12, 74>& Aeq, Vec<12>& beq, Mat<32, 74>& Aineq, Vec<32>& bineq) const;  // :156-294
    Vec<30> PDJointsAcc() const;                              // :296-309
    Vec<6> PDMomentumAcc() const;                             // :311-327
    Vec<12> PDFeetAcc(double t) const;                        // :329-386
    Vec<74> solveQP(const Mat<74, 74>& H, const Vec<74>& g);  // :388-479
};

// ---- apps/offline/main.cpp ----
// dynamics closure :91-122.  `second_wbc`: the app calls WBC(t) again after
// standStep (identical inputs, F7); false skips the duplicate solve.
Vec<60> dynamics(const Vec<60>& state, double t, Robot& robot, Controller& controller,
                 bool second_wbc, WBCOutput* out = nullptr);
// include/.../general/rk4.hpp:5-18
Vec<60> rk4Step(const Vec<60>& x, double t, double dt, Robot& robot, Controller& controller,
                bool second_wbc);

}  // namespace orc
