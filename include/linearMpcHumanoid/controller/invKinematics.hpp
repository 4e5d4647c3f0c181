// invKinematics.hpp -- the reference's Kinematics class (include/linearMpcHumanoid/controller/invKinematics.hpp:13-65,
// src/invKinematics.cpp).  Operational space: pose of each sole (12), arm and head joints (12), base attitude (3),
// centre of mass (3).  compute() is the Newton inverse kinematics on the device (lmh_ik_solve: one warp per target),
// computeAll() the sole Jacobians of one device stage evaluation.
#pragma once
#include "linearMpcHumanoid/general/lmhEigen.hpp"
#include "linearMpcHumanoid/robotInfo/Robot.hpp"
#include "linearMpcHumanoid/general/generalizedFunctions.hpp"
#include <cmath>
#include <iostream>

class Kinematics {
public:
    Kinematics() {
        JR_ = Eigen::MatrixXd::Zero(6, 30);
        JL_ = Eigen::MatrixXd::Zero(6, 30);
        JFeet_ = Eigen::MatrixXd::Zero(12, 30);
    }

    // sole Jacobians at the robot's configuration (invKinematics.cpp:271-276, 72-149): rows [angular; linear] in world axes
    void computeAll(Robot& robot) {
        double q[30], v[30];
        for (int i = 0; i < 30; ++i) { q[i] = robot.getJoints()(i); v[i] = robot.getJointsVelocity()(i); }
        load(lmhf::Engine::get().stage(q, v, v, 0.0));
    }

    // invKinematics.cpp:11-25: soles as given, arm/head joints as they are, level base, centre of mass as given
    Eigen::VectorXd desiredOperationalState(const Robot& robot, const Eigen::VectorXd& Rf, const Eigen::VectorXd& Lf,
                                            const Eigen::Vector3d& com) {
        Eigen::VectorXd Qd = Eigen::VectorXd::Zero(robot.getNumJoints());
        for (int k = 0; k < 6; ++k) { Qd(k) = Rf(k); Qd(6 + k) = Lf(k); }
        for (int k = 0; k < 12; ++k) Qd(12 + k) = robot.getJoints()(18 + k);
        for (int k = 0; k < 3; ++k) Qd(27 + k) = com(k);
        return Qd;
    }

    // Newton's method on the operational error, from the robot's current joints (invKinematics.cpp:27-52); the robot is
    // left at the solution.  Unlike upstream's loop (whose iteration counter never advances) an unreachable target ends.
    void compute(Robot& robot, const Eigen::VectorXd& desOp) {
        double qs[30], des[30], q[30], res = 0;
        int32_t it = 0, st = 0;
        for (int i = 0; i < 30; ++i) { qs[i] = robot.getJoints()(i); des[i] = desOp(i); }
        lmhf::check(lmh_ik_solve(lmhf::Engine::get().handle(), 1, qs, des, q, &it, &res, &st), "Kinematics::compute");
        iterations_ = it;
        if (st != 0) std::cout << "Inv Kinematics no solution founded" << std::endl;
        Eigen::VectorXd qn = Eigen::VectorXd::Zero(30);
        for (int i = 0; i < 30; ++i) qn(i) = q[i];
        robot.updateState(qn);
    }

    const Eigen::MatrixXd& getRightFootJacobian() const { return JR_; }
    const Eigen::MatrixXd& getLeftFootJacobian() const { return JL_; }
    const Eigen::MatrixXd& getFeetJacobian() const { return JFeet_; }
    int iterations() const { return iterations_; }      // (not upstream) Newton iterations of the last compute()

private:
    friend class Controller;
    void load(const LmhStageDump& d) {
        for (int r = 0; r < 12; ++r)
            for (int c = 0; c < 30; ++c) {
                const double v = d.JFeet[30 * r + c];
                JFeet_(r, c) = v;
                if (r < 6) JR_(r, c) = v; else JL_(r - 6, c) = v;
            }
    }
    Eigen::MatrixXd JR_;
    Eigen::MatrixXd JL_;
    Eigen::MatrixXd JFeet_;
    int iterations_ = 0;
};
