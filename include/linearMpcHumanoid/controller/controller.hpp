// controller.hpp -- the reference's Controller (include/linearMpcHumanoid/controller/controller.hpp:33-143,
// src/controller.cpp): LIPM-MPC + PD references + whole-body QP (momentum-rate, foot-acceleration and joint-acceleration
// tasks under contact-force constraints) + inverse dynamics, with the reference's constructor and methods.
//
//   state = [q, v],  q = [0p1 (base position), 0eta1 (base roll/pitch/yaw), qJ (24 joints)],
//                    v = [0v1 (base linear, spatial), 0w1 (base angular), qDJ]          (controller.hpp:20-31)
//
// One Controller owns one single-environment batch of the GPU library, built from the objects it is handed: time step,
// horizon, CoM height and weights of the Mpc3dLip, the ZMP object's references, the two foot coefficient sets.
// standStep() = set the state + ONE fused stage launch (model update, MPC, WBC QP, inverse dynamics); it then leaves the
// Robot and the Mpc3dLip as upstream's standStep does (joints, held velocity, CoM, CoM velocity, angular momentum,
// xRef/yRef).  WBC(t) returns that stage's solution (the app calls it right after standStep with the same inputs,
// main.cpp:103-105; the solve is deterministic).  rollout() exposes what the library is for: the fused batched RK4.
#pragma once
#include "linearMpcHumanoid/general/lmhEigen.hpp"

#include "linearMpcHumanoid/controller/mpcLinearPendulum.hpp"
#include "linearMpcHumanoid/controller/Dynamics.hpp"
#include "linearMpcHumanoid/controller/invKinematics.hpp"

#include "linearMpcHumanoid/robotInfo/Robot.hpp"

#include "linearMpcHumanoid/trajectories/zmpGeneration.hpp"
#include "linearMpcHumanoid/trajectories/footRefTrajectory.hpp"

#include "linearMpcHumanoid/general/Task.hpp"
#include "linearMpcHumanoid/general/Clock.hpp"
#include "linearMpcHumanoid/general/generalizedFunctions.hpp"
#include "linearMpcHumanoid/general/rk4.hpp"

struct ControllerInput {
    Eigen::VectorXd q;
    Eigen::VectorXd dq;
    double time;
};

struct ControllerOutput {
    Eigen::VectorXd tau;
};

struct WBCOutput {
    Eigen::VectorXd qpp;   // generalized accelerations (30), base part in the world frame
    Eigen::VectorXd tau;   // joint torques (24)
    Eigen::VectorXd f;     // contact wrenches of the two soles (12)
};

class Controller {
public:
    Controller(Robot& robot, Mpc3dLip& mpc, ZMP& zmp, std::vector<Eigen::VectorXd>& rFCoeff, std::vector<Eigen::VectorXd>& lFCoeff)
        : robot_(robot), mpc_(mpc), zmp_(zmp), rFCoeff_(rFCoeff), lFCoeff_(lFCoeff) {
        LmhConfig c;
        lmh_default_config(&c);
        c.dt = mpc.getTimeStep(); c.time_horizon = mpc.getTimeHorizon(); c.mpc_zcom = mpc.getZCom();
        c.mpc_alpha = mpc.getAlpha(); c.mpc_beta = mpc.getBeta();
        LmhBatch* h = nullptr;
        lmhf::check(lmh_create(&c, 1, &h), "Controller");
        h_ = std::shared_ptr<LmhBatch>(h, lmh_destroy);
        // the plan: the ZMP object's references and the two coefficient sets (controller.hpp:52-57), exactly as handed in
        const Eigen::VectorXd zx = zmp.getZmpXRef(), zy = zmp.getZmpYRef();
        if (zx.size() != zy.size() || (int)zx.size() < mpc.getHorizon() + 1)
            throw std::invalid_argument("Controller: the ZMP references must have equal lengths and cover the MPC horizon");
        std::vector<double> vx((size_t)zx.size()), vy((size_t)zy.size());
        for (int i = 0; i < (int)zx.size(); ++i) { vx[i] = zx(i); vy[i] = zy(i); }
        if (zmp.walkPhases() > 0) {      // EXTENSION: a walking plan from ZMP::walkZMP carries its own per-phase foot polynomials
            const int32_t pid = 0;
            lmhf::check(lmh_set_walk_plans(h, 1, (int32_t)vx.size(), vx.data(), vy.data(), zmp.walkPhases(), zmp.walkPhaseTime(),
                                           zmp.walkFootCoefficients().data(), zmp.walkContactMasks().data(), &pid), "Controller");
        } else {
            double foot[40];
            pack(rFCoeff, foot);
            pack(lFCoeff, foot + 20);
            lmhf::check(lmh_set_plans(h, 1, (int32_t)vx.size(), vx.data(), vy.data(), foot, nullptr), "Controller");
        }
        tau_ = Eigen::VectorXd::Zero(24);
        state_ = Eigen::VectorXd::Zero(60);
    }

    // controller.cpp:48-79: Robot::updateState, Dynamics/Kinematics::computeAll, Robot::updateVelocityState,
    // Mpc3dLip::compute, WBC -- one fused launch
    void standStep(const ControllerInput& in) {
        double x[60], vh[30], t = in.time;
        for (int i = 0; i < 30; ++i) { x[i] = in.q(i); x[30 + i] = in.dq(i); vh[i] = robot_.v_(i); }
        lmhf::check(lmh_set_state(h_.get(), x, vh, &t), "Controller::standStep");
        lmhf::check(lmh_stage(h_.get(), &last_), "Controller::standStep");
        unload(in.time);
        // what upstream's standStep leaves behind in the objects it references
        robot_.q_ = in.q;
        robot_.v_ = in.dq;
        robot_.frames_valid_ = false;                  // T_, X_ follow on demand (Robot::getT / getX)
        for (int k = 0; k < 3; ++k) {
            robot_.CoM_(k) = last_.CoM[k]; robot_.comVel_(k) = last_.comVel[k]; robot_.comAngMom_(k) = last_.comAngMom[k];
            mpc_.xRef_(k) = last_.xRef[k]; mpc_.yRef_(k) = last_.yRef[k];
        }
        dyn_.load(last_);
        kin_.load(last_);
        for (int i = 0; i < 60; ++i) state_(i) = x[i];
    }

    // controller.cpp:81-154.  At the time of the last standStep this is that stage's solution; at another time the
    // references (foot polynomials, ZMP window) are re-evaluated there on the same state.
    WBCOutput WBC(double t) {
        if (!have_stage_) throw std::logic_error("Controller::WBC: call standStep first");
        if (t != stage_time_) {
            double x[60], vh[30];
            for (int i = 0; i < 60; ++i) x[i] = state_(i);
            for (int i = 0; i < 30; ++i) vh[i] = robot_.v_(i);
            lmhf::check(lmh_set_state(h_.get(), x, vh, &t), "Controller::WBC");
            lmhf::check(lmh_stage(h_.get(), &last_), "Controller::WBC");
            unload(t);
        }
        return out_;
    }

    const Eigen::VectorXd& getTorques() { return tau_; }

    // ---- not upstream ----
    const LmhStageDump& lastStage() const { return last_; }          // every intermediate of the last stage evaluation
    const Dynamics& dynamics() const { return dyn_; }
    const Kinematics& kinematics() const { return kin_; }
    LmhBatch* batch() { return h_.get(); }

private:
    static void pack(const std::vector<Eigen::VectorXd>& c, double* out20) {
        if (c.size() != 3 || c[0].size() != 6 || c[1].size() != 6 || c[2].size() != 8)
            throw std::invalid_argument("Controller: foot coefficients must be the three vectors footCoeffTrajectory returns (6, 6, 8)");
        for (int i = 0; i < 6; ++i) { out20[i] = c[0](i); out20[6 + i] = c[1](i); }
        for (int i = 0; i < 8; ++i) out20[12 + i] = c[2](i);
    }
    void unload(double t) {
        out_.qpp = Eigen::VectorXd::Zero(30); out_.tau = Eigen::VectorXd::Zero(24); out_.f = Eigen::VectorXd::Zero(12);
        for (int i = 0; i < 30; ++i) out_.qpp(i) = last_.qpp[i];
        for (int i = 0; i < 24; ++i) { out_.tau(i) = last_.tau[i]; tau_(i) = last_.tau[i]; }
        for (int i = 0; i < 12; ++i) out_.f(i) = last_.f[i];
        stage_time_ = t;
        have_stage_ = true;
    }

    Eigen::VectorXd tau_;     // 24 joint torques
    Eigen::VectorXd state_;   // 60: configuration and velocity of the last standStep

    Robot& robot_;
    Mpc3dLip& mpc_;
    ZMP zmp_;
    std::vector<Eigen::VectorXd> rFCoeff_;
    std::vector<Eigen::VectorXd> lFCoeff_;
    Dynamics dyn_;
    Kinematics kin_;

    std::shared_ptr<LmhBatch> h_;
    LmhStageDump last_;
    WBCOutput out_;
    double stage_time_ = 0;
    bool have_stage_ = false;
};
