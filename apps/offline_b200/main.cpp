// The reference's offline app (apps/offline/main.cpp:12-122) rebuilt against the facade classes of
// include/linearMpcHumanoid_b200/facade.hpp: same objects, same call order, same printed stream (CoM x per step,
// main.cpp:86), every number computed on the GPU.  With "--fused N" the stand loop is replaced by the fused
// batched rollout (one launch for all steps of N environments) -- what the library is for.
//
//   g++ -O2 -std=c++17 -Iinclude apps/offline_b200/main.cpp -Llinearmpchumanoid_b200 -llmh_b200 -Wl,-rpath,$PWD/linearmpchumanoid_b200 -o offline_b200
#include "linearMpcHumanoid_b200/facade.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
using namespace lmhb;

static Vec dynamics(const Vec& state, double t, Robot& robot, Controller& controller) {   // main.cpp:91-122
    const int n = robot.getNumJoints();
    Vec q = state.segment(0, n), qD = state.segment(n, n);
    ControllerInput in;
    in.q = q; in.dq = qD; in.time = t;
    controller.standStep(in);
    WBCOutput out = controller.WBC(t);
    Vec wxp = crossTimes(qD.segment(3, 3), q.segment(0, 3));
    for (int k = 0; k < 3; ++k) qD[k] += wxp[k];
    qD.setSegment(3, eulerRates(q.segment(3, 3), qD.segment(3, 3)));
    Vec xp(2 * n);
    xp.setSegment(0, qD);
    xp.setSegment(n, out.qpp);
    return xp;
}

int main(int argc, char** argv) {
    double simulationTime = 5, timeStep = 0.01;
    Robot nao;
    Kinematics ik;
    Clock clock(timeStep, simulationTime);
    ZMP zmp(Task::Stand, simulationTime, timeStep, SupportFoot::Double);
    Vec Rf = Vec::Zero(6), Lf = Vec::Zero(6), com{-0.02, 0.0, 0.26};
    Rf[1] = -0.05; Lf[1] = 0.05;
    Vec desOp = ik.desiredOperationalState(nao, Rf, Lf, com);
    ik.compute(nao, desOp);
    Mpc3dLip mpc(clock.getTimeStep(), 0.5, nao.getCoM()(2));
    std::vector<Vec> rF = footCoeffTrajectory(Vec{0, -0.05, 0}, Vec{0, -0.05, 0}, 0, simulationTime);
    std::vector<Vec> lF = footCoeffTrajectory(Vec{0, 0.05, 0}, Vec{0, 0.05, 0}, 0, simulationTime);
    int fused = 0;
    for (int i = 1; i + 1 < argc; ++i) if (!std::strcmp(argv[i], "--fused")) fused = std::atoi(argv[i + 1]);
    Controller controller(nao, mpc, zmp, rF, lF, fused > 0 ? fused : 1);
    if (fused > 0) {
        controller.rollout(500);
        std::vector<double> c(3 * fused);
        lmh_get_com(controller.batch(), c.data());
        std::printf("%.6g\n", c[0]);
        return 0;
    }
    const int n = nao.getNumJoints();
    Vec state(2 * n);
    state.setSegment(0, nao.getJoints());
    state.setSegment(n, nao.getJointsVelocity());
    while (std::fabs(clock.getTime() - clock.getSimulationTime()) > 0.01) {            // main.cpp:74
        state = rk4Step([&](const Vec& x, double t) { return dynamics(x, t, nao, controller); }, state, clock.getTime(),
                        clock.getTimeStep());
        std::printf("%.6g\n\n", nao.getCoM()(0));                                        // main.cpp:86
        clock.step();
    }
    return 0;
}
