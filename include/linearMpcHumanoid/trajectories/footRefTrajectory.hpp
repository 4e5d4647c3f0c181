// footRefTrajectory.hpp -- footCoeffTrajectory (reference: include/linearMpcHumanoid/trajectories/footRefTrajectory.hpp,
// src/footRefTrajectory.cpp:4-47): polynomial coefficients (ascending powers) of a foot motion over [0, T]:
// [0] x: 5th order, rest to rest; [1] y: the same; [2] z: 7th order through stepHeight at T/2.  The fits are the library's
// (lmh_foot_coeff); the Controller uploads whatever coefficient vectors it is handed.
#pragma once
#include <vector>
#include "linearMpcHumanoid/general/lmhEigen.hpp"
#include "linearMpcHumanoid/general/generalizedFunctions.hpp"
#include "linearMpcHumanoid/lmhEngine.hpp"

inline std::vector<Eigen::VectorXd> footCoeffTrajectory(const Eigen::Vector3d& currentPos, const Eigen::Vector3d& desPos,
                                                        double stepHeight, double T) {
    const double cur[3] = {currentPos(0), currentPos(1), currentPos(2)}, des[3] = {desPos(0), desPos(1), desPos(2)};
    double x[6], y[6], z[8];
    lmhf::check(lmh_foot_coeff(cur, des, stepHeight, T, x, y, z), "footCoeffTrajectory");
    std::vector<Eigen::VectorXd> Coeff(3);
    Coeff[0] = Eigen::VectorXd::Zero(6); Coeff[1] = Eigen::VectorXd::Zero(6); Coeff[2] = Eigen::VectorXd::Zero(8);
    for (int i = 0; i < 6; ++i) { Coeff[0](i) = x[i]; Coeff[1](i) = y[i]; }
    for (int i = 0; i < 8; ++i) Coeff[2](i) = z[i];
    return Coeff;
}
