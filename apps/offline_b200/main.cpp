// The offline app on the B200 library.  The reference's own apps/offline/main.cpp compiles UNCHANGED against
// include/linearMpcHumanoid/** (tests/test_class_surface.py does exactly that); this file is the same program plus
// what the reference does not have: "--fused N" replaces the host-driven stand loop (4 launches per RK4 step) by the
// fused batched rollout -- one launch for all 500 steps of N environments -- and "--walk" runs the walking extension
// (ZMP::walkZMP) through the same class surface.
//
//   g++ -O2 -std=c++17 -Iinclude apps/offline_b200/main.cpp -Llinearmpchumanoid_b200 -llmh_b200
//       -Wl,-rpath,$PWD/linearmpchumanoid_b200 -o offline_b200
#include "linearMpcHumanoid/robotInfo/Robot.hpp"
#include "linearMpcHumanoid/controller/controller.hpp"
#include "linearMpcHumanoid/controller/invKinematics.hpp"
#include "linearMpcHumanoid/controller/mpcLinearPendulum.hpp"
#include "linearMpcHumanoid/general/Clock.hpp"
#include "linearMpcHumanoid/general/Task.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>

// the dynamics() closure of apps/offline/main.cpp:91-122
static Eigen::VectorXd dynamics(const Eigen::VectorXd& state, double t, Robot& robot, Controller& controller) {
    const int n = robot.getNumJoints();
    Eigen::VectorXd q = state.segment(0, n), qD = state.segment(n, n);
    ControllerInput in;
    in.q = q; in.dq = qD; in.time = t;
    controller.standStep(in);
    WBCOutput out = controller.WBC(t);
    qD.segment(0, 3) += crossMatrix(qD.segment(3, 3)) * q.segment(0, 3);            // spatial -> classical linear velocity
    qD.segment(3, 3) = matrixAngularVelToEulerDot(q.segment(3, 3)) * qD.segment(3, 3);   // angular velocity -> Euler rates
    Eigen::VectorXd xp(2 * n);
    xp.head(n) = qD;
    xp.tail(n) = out.qpp;
    return xp;
}

int main(int argc, char** argv) {
    int fused = 0;
    bool walk = false;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--fused") && i + 1 < argc) fused = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--walk")) walk = true;
    }
    const double simulationTime = walk ? 2.4 : 5, timeStep = 0.01;
    Robot nao;
    Kinematics ik;
    Clock clock(timeStep, simulationTime);
    ZMP zmp = walk ? ZMP(Task::Walk, 4, 0.4, simulationTime) : ZMP(Task::Stand, simulationTime, timeStep, SupportFoot::Double);
    if (walk) zmp.walkZMP();
    Eigen::VectorXd Rf = Eigen::VectorXd::Zero(6), Lf = Eigen::VectorXd::Zero(6);
    Rf(1) = -0.05; Lf(1) = 0.05;
    Eigen::Vector3d com(-0.02, 0.0, 0.26);
    ik.compute(nao, ik.desiredOperationalState(nao, Rf, Lf, com));
    Mpc3dLip mpc(clock.getTimeStep(), walk ? 1.6 : 0.5, nao.getCoM()(2));
    std::vector<Eigen::VectorXd> rF = footCoeffTrajectory(Eigen::Vector3d(0, -0.05, 0), Eigen::Vector3d(0, -0.05, 0), 0, simulationTime);
    std::vector<Eigen::VectorXd> lF = footCoeffTrajectory(Eigen::Vector3d(0, 0.05, 0), Eigen::Vector3d(0, 0.05, 0), 0, simulationTime);
    if (fused > 0) {
        // the batched path: N environments, every one the scenario above, all steps in one launch
        LmhConfig c;
        lmh_default_config(&c);
        c.time_horizon = mpc.getTimeHorizon(); c.mpc_zcom = mpc.getZCom();
        LmhBatch* h = nullptr;
        lmhf::check(lmh_create(&c, fused, &h), "offline_b200");
        if (walk) {
            LmhWalkSpec ws;
            lmh_default_walk_spec(&ws);
            lmhf::check(lmh_make_walk_plans(h, &ws, 0, 0x5EED0003), "offline_b200");
        }
        const int steps = (int)(simulationTime / timeStep + 0.5);
        lmhf::check(lmh_step(h, steps), "offline_b200");
        std::vector<double> cm(3 * (size_t)fused);
        lmhf::check(lmh_get_com(h, cm.data()), "offline_b200");
        std::printf("%.6g\n", cm[0]);
        lmh_destroy(h);
        return 0;
    }
    Controller controller(nao, mpc, zmp, rF, lF);
    const int n = nao.getNumJoints();
    Eigen::VectorXd state(2 * n);
    state.segment(0, n) = nao.getJoints();
    state.segment(n, n) = nao.getJointsVelocity();
    while (std::abs(clock.getTime() - clock.getSimulationTime()) > 0.01) {            // main.cpp:74
        state = rk4Step([&](const Eigen::VectorXd& x, double t) { return dynamics(x, t, nao, controller); }, state,
                        clock.getTime(), clock.getTimeStep());
        std::printf("%.6g\n\n", nao.getCoM()(0));                                        // main.cpp:86
        clock.step();
    }
    return 0;
}
