// mpcLinearPendulum.hpp -- the reference's Mpc3dLip (include/linearMpcHumanoid/controller/mpcLinearPendulum.hpp:5-50,
// src/mpcLinearPendulum.cpp): model predictive control of the 3-D linear inverted pendulum with the ZMP as output,
//   min  alpha/2 |u|^2 + beta/2 |Px x_k + Pu u - z_ref|^2      over the jerk-free horizon (N + 1 accelerations).
// Same three constructors, compute() and getters.  compute() solves the condensed QP as a dense system ON THE DEVICE
// (lmh_mpc_solve: gradient by a dense product, Cholesky factor of the constant Hessian, forward and back substitution)
// where upstream hands H and g to qpOASES; the closed-loop Controller reaches the same u0 through a precomputed gain.
#pragma once
#include "linearMpcHumanoid/general/lmhEigen.hpp"
#include "linearMpcHumanoid/lmhEngine.hpp"
#include <memory>
#include <vector>

class Mpc3dLip {
public:
    Mpc3dLip() { initialize(); }
    Mpc3dLip(const double dt, const double timeHorizon, const double zCoM) : dt_(dt), timeHorizon_(timeHorizon), zCom_(zCoM) { initialize(); }
    Mpc3dLip(const double dt, const double timeHorizon, const double zCoM, const double alpha, const double beta)
        : dt_(dt), timeHorizon_(timeHorizon), zCom_(zCoM), alpha_(alpha), beta_(beta) { initialize(); }

    Eigen::Vector3d getXRef() const { return xRef_; }       // position, velocity, acceleration of the CoM in x
    Eigen::Vector3d getYRef() const { return yRef_; }
    double getZCom() const { return zCom_; }

    // mpcLinearPendulum.cpp:78-109: window k = (int)(t / dt) of the ZMP references, one dense solve per axis, x+ = A x + B u0
    void compute(const Eigen::Vector2d& posCom, const Eigen::Vector2d& velCom, const Eigen::VectorXd& zmpXRef,
                 const Eigen::VectorXd& zmpYRef, double t) {
        const int k = static_cast<int>(t / dt_), n1 = horizon_ + 1;
        if (k < 0 || k + n1 > (int)zmpXRef.size() || k + n1 > (int)zmpYRef.size())
            throw std::out_of_range("Mpc3dLip::compute: the ZMP reference does not cover the horizon at this time");
        std::vector<double> zx(n1), zy(n1);
        for (int i = 0; i < n1; ++i) { zx[i] = zmpXRef(k + i); zy[i] = zmpYRef(k + i); }
        const double com[2] = {posCom(0), posCom(1)}, vel[2] = {velCom(0), velCom(1)};
        double xr[3], yr[3];
        lmhf::check(lmh_mpc_solve(batch(), 1, com, vel, zx.data(), zy.data(), xr, yr, 0, 1, nullptr), "Mpc3dLip::compute");
        for (int i = 0; i < 3; ++i) { xRef_(i) = xr[i]; yRef_(i) = yr[i]; }
    }

    // (not upstream) the scenario literals a Controller needs to build its batch
    double getTimeStep() const { return dt_; }
    double getTimeHorizon() const { return timeHorizon_; }
    double getAlpha() const { return alpha_; }
    double getBeta() const { return beta_; }
    int getHorizon() const { return horizon_; }

private:
    friend class Controller;
    void initialize() {
        horizon_ = static_cast<int>(timeHorizon_ / dt_);     // mpcLinearPendulum.cpp:43
        xRef_ = Eigen::Vector3d::Zero();
        yRef_ = Eigen::Vector3d::Zero();
    }
    // the prediction matrices, the Hessian factor and the device tables live in a one-environment batch of this MPC's
    // parameters, created at the first compute()
    LmhBatch* batch() {
        if (!h_) {
            LmhConfig c;
            lmh_default_config(&c);
            c.dt = dt_; c.time_horizon = timeHorizon_; c.mpc_zcom = zCom_; c.mpc_alpha = alpha_; c.mpc_beta = beta_;
            LmhBatch* h = nullptr;
            lmhf::check(lmh_create(&c, 1, &h), "Mpc3dLip");
            h_ = std::shared_ptr<LmhBatch>(h, lmh_destroy);
        }
        return h_.get();
    }
    Eigen::Vector3d xRef_;
    Eigen::Vector3d yRef_;
    double dt_ = 0.01;
    double timeHorizon_ = 0.5;
    double zCom_ = 0.26;
    double gravity_ = 9.81;
    double alpha_ = 1e-3;
    double beta_ = 1;
    int horizon_ = 50;
    std::shared_ptr<LmhBatch> h_;
};
