// TEST PROGRAM: exercises the reference's class surface as shipped in include/linearMpcHumanoid/** (Robot, Dynamics,
// Kinematics, Mpc3dLip, ZMP, footCoeffTrajectory, Controller) and prints named arrays, one per line:  name v0 v1 ...
// tests/test_class_surface.py compares them with the reference vectors.  Needs a GPU to run.
#include "linearMpcHumanoid/controller/controller.hpp"
#include <cstdio>

static void put(const char* name, const Eigen::MatrixXd& m) {
    std::printf("%s", name);
    for (int i = 0; i < (int)m.rows(); ++i) for (int j = 0; j < (int)m.cols(); ++j) std::printf(" %.17g", m(i, j));
    std::printf("\n");
}
static void putv(const char* name, const Eigen::VectorXd& v) {
    std::printf("%s", name);
    for (int i = 0; i < (int)v.size(); ++i) std::printf(" %.17g", v(i));
    std::printf("\n");
}

int main() {
    Robot nao;                                           // main.cpp:16-35
    Kinematics ik;
    Eigen::VectorXd Rf = Eigen::VectorXd::Zero(6), Lf = Eigen::VectorXd::Zero(6);
    Rf(1) = -0.05; Lf(1) = 0.05;
    Eigen::Vector3d com = Eigen::Vector3d::Zero();
    com << -0.02, 0.0, 0.26;
    std::printf("mass %.17g\nframes %d %d %d %d\n", nao.getMass(), nao.getNumFrames(), nao.getNumJoints(), nao.getNumActualJoints(), nao.getNumBodies());
    putv("link_mass", [&] { Eigen::VectorXd m(28); for (int i = 0; i < 28; ++i) m(i) = nao.getLinks()[i].mass; return m; }());
    put("link_com3", nao.getLinks()[3].com);
    put("link_inertia3", nao.getLinks()[3].inertia);
    putv("q_init", nao.getJoints());
    ik.compute(nao, ik.desiredOperationalState(nao, Rf, Lf, com));
    std::printf("ik_iterations %d\n", ik.iterations());
    putv("q0", nao.getJoints());
    putv("com0", nao.getCoM());
    put("T7", nao.getT()[7]);
    put("T27", nao.getT()[27]);
    put("X5", nao.getX()[5]);
    std::printf("tree %d %d %zu\n", nao.parentFrame()[14], nao.actuatedFrames()[13], nao.getFootVertices().size());

    Dynamics dyn;                                        // Dynamics.hpp:13-20 at the IK posture, zero velocity
    dyn.computeAll(nao);
    put("M", dyn.getM()); putv("C", dyn.getC()); putv("Cg", dyn.getCg()); put("AG", dyn.getAG());
    putv("AGpqp", dyn.getAGpqp()); putv("Jpqp", dyn.getJpqp());
    ik.computeAll(nao);
    put("JFeet", ik.getFeetJacobian()); put("JR", ik.getRightFootJacobian()); put("JL", ik.getLeftFootJacobian());
    nao.updateVelocityState(Eigen::VectorXd::Constant(30, 0.01), dyn.getAG());
    putv("comVel", nao.getComVel()); putv("angMom", nao.getComAngMom());
    nao.updateVelocityState(Eigen::VectorXd::Zero(30), dyn.getAG());

    Mpc3dLip mpc(0.01, 0.5, nao.getCoM()(2));            // mpcLinearPendulum.hpp:8-31
    Mpc3dLip mpcDefault;
    Mpc3dLip mpc5(0.01, 0.5, nao.getCoM()(2), 1e-3, 1.0);
    ZMP zmp(Task::Stand, 5, 0.01, SupportFoot::Double), zmpShort(Task::Stand), zmpLeft(Task::Stand, 5, 0.01, SupportFoot::Left);
    std::printf("zmp_sizes %d %d %.17g\n", (int)zmp.getZmpXRef().size(), (int)zmpShort.getZmpXRef().size(), zmpLeft.getZmpYRef()(7));
    Eigen::Vector2d pc(nao.getCoM()(0), nao.getCoM()(1)), vc(0.0, 0.0);
    mpc.compute(pc, vc, zmp.getZmpXRef(), zmp.getZmpYRef(), 0.0);
    putv("xRef", mpc.getXRef()); putv("yRef", mpc.getYRef());
    mpc5.compute(pc, vc, zmp.getZmpXRef(), zmpLeft.getZmpYRef(), 0.123);
    putv("yRef_left", mpc5.getYRef());
    std::printf("zcom %.17g %.17g\n", mpc.getZCom(), mpcDefault.getZCom());

    std::vector<Eigen::VectorXd> rF = footCoeffTrajectory(Eigen::Vector3d(0, -0.05, 0), Eigen::Vector3d(0, -0.05, 0), 0, 5);
    std::vector<Eigen::VectorXd> lF = footCoeffTrajectory(Eigen::Vector3d(0, 0.05, 0), Eigen::Vector3d(0, 0.05, 0), 0, 5);
    putv("rF0", rF[0]); putv("rF1", rF[1]); putv("rF2", rF[2]); putv("lF1", lF[1]);
    Controller controller(nao, mpc, zmp, rF, lF);        // controller.hpp:52-63
    ControllerInput in;
    in.q = nao.getJoints(); in.dq = Eigen::VectorXd::Zero(30); in.time = 0.0;
    controller.standStep(in);
    WBCOutput out = controller.WBC(0.0);
    putv("qpp", out.qpp); putv("tau", out.tau); putv("f", out.f); putv("tau_get", controller.getTorques());
    putv("mpc_xRef_after", mpc.getXRef());
    // non-default foot coefficients must reach the QP: a right foot asked to move 3 cm forward and lift 2 cm in 0.4 s
    std::vector<Eigen::VectorXd> rMove = footCoeffTrajectory(Eigen::Vector3d(0, -0.05, 0), Eigen::Vector3d(0.03, -0.05, 0), 0.02, 0.4);
    Controller moving(nao, mpc, zmp, rMove, lF);
    in.time = 0.2;
    moving.standStep(in);
    controller.standStep(in);
    std::printf("footAccRef_still");
    for (int k = 0; k < 12; ++k) std::printf(" %.17g", controller.lastStage().footAccRef[k]);
    std::printf("\nfootAccRef_moving");
    for (int k = 0; k < 12; ++k) std::printf(" %.17g", moving.lastStage().footAccRef[k]);
    std::printf("\n");
    putv("rMove0", rMove[0]); putv("rMove2", rMove[2]);
    // the walking extension through the class surface
    ZMP walk(Task::Walk, 4, 0.4, 2.4);
    walk.walkZMP();
    std::printf("walk %d %d %.17g %.17g\n", (int)walk.getZmpXRef().size(), walk.walkPhases(), walk.getZmpXRef()(100), walk.getZmpYRef()(60));
    Mpc3dLip mpcWalk(0.01, 1.6, nao.getCoM()(2));
    Controller walker(nao, mpcWalk, walk, rF, lF);
    in.time = 0.6;
    walker.standStep(in);
    std::printf("walk_stage %d", walker.lastStage().phase);
    for (int k = 0; k < 12; ++k) std::printf(" %.17g", walker.lastStage().f[k]);
    std::printf("\n");
    // errors are exceptions, not silence
    int caught = 0;
    try { ZMP tiny(Task::Stand, 0.0, 0.01, SupportFoot::Double);   /* 50 samples < horizon + 1 */ Controller bad(nao, mpc, tiny, rF, lF); } catch (const std::exception&) { ++caught; }
    try { std::vector<Eigen::VectorXd> two(2); Controller bad(nao, mpc, zmp, two, lF); } catch (const std::exception&) { ++caught; }
    try { mpc.compute(pc, vc, zmp.getZmpXRef(), zmp.getZmpYRef(), 5.2); } catch (const std::exception&) { ++caught; }
    std::printf("caught %d\n", caught);
    return 0;
}
