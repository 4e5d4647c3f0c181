// rk4.hpp -- classical fourth-order Runge-Kutta step with the reference's signature
// (include/linearMpcHumanoid/general/rk4.hpp:5-18): four evaluations of f, combined with weights 1, 2, 2, 1.  This is the
// host-driven form the offline app uses (every f is one GPU stage evaluation through Controller::standStep); the fused
// device-side integrator is lmh_step (include/lmh_capi.h).
#pragma once
#include "linearMpcHumanoid/general/lmhEigen.hpp"

template <typename DynamicsFunction>
Eigen::VectorXd rk4Step(DynamicsFunction f, const Eigen::VectorXd& x, double t, double dt) {
    const double half = 0.5 * dt;
    const Eigen::VectorXd s1 = f(x, t);
    const Eigen::VectorXd s2 = f(x + half * s1, t + half);
    const Eigen::VectorXd s3 = f(x + half * s2, t + half);
    const Eigen::VectorXd s4 = f(x + dt * s3, t + dt);
    return x + (dt / 6.0) * (s1 + 2.0 * s2 + 2.0 * s3 + s4);
}
