// Dynamics.hpp -- the reference's Dynamics class (include/linearMpcHumanoid/controller/Dynamics.hpp:9-20,
// src/Dynamics.cpp): floating-base mass matrix (CRBA), bias forces with and without gravity (RNEA), centroidal momentum
// matrix and bias, Jdot*qdot of both soles.  computeAll(robot) is ONE device stage evaluation at the robot's
// configuration and HELD velocity (robot.getJointsVelocity(), as upstream: Dynamics.cpp:29-60 reads robot's v_); the six
// getters hand out what that evaluation produced.
#pragma once
#include "linearMpcHumanoid/robotInfo/Robot.hpp"
#include "linearMpcHumanoid/robotInfo/linkInertia.hpp"
#include "linearMpcHumanoid/general/generalizedFunctions.hpp"
#include "linearMpcHumanoid/controller/invKinematics.hpp"
#include "linearMpcHumanoid/general/lmhEigen.hpp"
#include <vector>

class Dynamics {
public:
    Dynamics() = default;

    void computeAll(const Robot& robot) {
        double q[30], v[30];
        for (int i = 0; i < 30; ++i) { q[i] = robot.getJoints()(i); v[i] = robot.getJointsVelocity()(i); }
        load(lmhf::Engine::get().stage(q, v, v, 0.0));
    }

    const Eigen::VectorXd& getC() const { return C_; }         // bias forces with gravity (Dynamics.cpp:29-60)
    const Eigen::VectorXd& getCg() const { return Cg_; }       // ... without gravity (needed by the centroidal dynamics)
    const Eigen::MatrixXd& getM() const { return M_; }         // Dynamics.cpp:62-101
    const Eigen::MatrixXd& getAG() const { return AG_; }       // Dynamics.cpp:103-121
    const Eigen::VectorXd& getAGpqp() const { return AGpqp_; }
    const Eigen::VectorXd& getJpqp() const { return Jpqp_; }   // Dynamics.cpp:165-200

private:
    friend class Controller;
    void load(const LmhStageDump& d) {
        C_ = Eigen::VectorXd::Zero(30); Cg_ = Eigen::VectorXd::Zero(30); M_ = Eigen::MatrixXd::Zero(30, 30);
        AG_ = Eigen::MatrixXd::Zero(6, 30); AGpqp_ = Eigen::VectorXd::Zero(6); Jpqp_ = Eigen::VectorXd::Zero(12);
        for (int i = 0; i < 30; ++i) {
            C_(i) = d.C[i]; Cg_(i) = d.Cg[i];
            for (int j = 0; j < 30; ++j) M_(i, j) = d.M[30 * i + j];
        }
        for (int k = 0; k < 6; ++k) {
            AGpqp_(k) = d.AGpqp[k];
            for (int j = 0; j < 30; ++j) AG_(k, j) = d.AG[30 * k + j];
        }
        for (int k = 0; k < 12; ++k) Jpqp_(k) = d.Jpqp[k];
    }
    Eigen::VectorXd C_;
    Eigen::VectorXd Cg_;
    Eigen::MatrixXd M_;
    Eigen::MatrixXd AG_;
    Eigen::VectorXd AGpqp_;
    Eigen::VectorXd Jpqp_;
};
