// zmpGeneration.hpp -- the reference's ZMP class (include/linearMpcHumanoid/trajectories/zmpGeneration.hpp:5-38,
// src/zmpGeneration.cpp): the three constructors, stanceZMP(), and walkZMP() -- which upstream declares but never
// defines.  Here walkZMP() is the library's walking-plan generator (lmh_walk_plan: numSteps footsteps of one nominal
// length, right foot first, a phase of double support before and after; ZMP under the stance foot in single support,
// between the feet in double support), and the per-phase foot polynomials and contact masks that go with it are kept
// for the Controller.
#pragma once
#include "linearMpcHumanoid/general/lmhEigen.hpp"
#include "linearMpcHumanoid/general/Task.hpp"
#include "linearMpcHumanoid/lmhEngine.hpp"
#include <vector>

class ZMP {
public:
    ZMP(const Task task) : task_(task) { stanceZMP(); }                                   // regulation, double support
    ZMP(const Task task, const double simulationTime, const double timeStep, const SupportFoot supportFoot)
        : simulationTime_(simulationTime), timeStep_(timeStep), supportFoot_(supportFoot), task_(task) { stanceZMP(); }
    ZMP(const Task task, const int numSteps, const double timePerStep, const double simulationTime)
        : simulationTime_(simulationTime), task_(task), numSteps_(numSteps), timePerStep_(timePerStep) {}   // walk: call walkZMP()

    // constant reference in the middle of the support zone for the whole run plus 0.5 s (zmpGeneration.cpp:39-60)
    void stanceZMP() {
        const int samples = static_cast<int>((simulationTime_ + 0.5) / timeStep_);
        const double y = supportFoot_ == SupportFoot::Right ? -0.05 : supportFoot_ == SupportFoot::Left ? 0.05 : 0.0;
        zmpXRef_ = Eigen::VectorXd::Zero(samples);
        zmpYRef_ = Eigen::VectorXd::Zero(samples);
        for (int i = 0; i < samples; ++i) zmpYRef_(i) = y;
        walk_phases_ = 0;
    }
    // EXTENSION: the walking reference.  stepLength is the nominal step; horizonSteps = samples the MPC looks ahead.
    void walkZMP(double stepLength = 0.035, int horizonSteps = 160) {
        LmhWalkSpec s;
        lmh_default_walk_spec(&s);
        s.n_footsteps = numSteps_; s.time_per_step = timePerStep_; s.dt = timeStep_; s.fixed_len = stepLength;
        s.horizon_steps = horizonSteps;
        const int covered = static_cast<int>((simulationTime_ + 0.5) / timeStep_);
        int32_t P = 0, Z = 0;
        lmhf::check(lmh_walk_plan_dims(&s, &P, &Z), "ZMP::walkZMP");
        if (Z < covered + horizonSteps + 2) { s.tail_steps += covered + horizonSteps + 2 - Z; lmhf::check(lmh_walk_plan_dims(&s, &P, &Z), "ZMP::walkZMP"); }
        std::vector<double> zx(Z), zy(Z);
        walk_foot_coef_.assign((size_t)P * 40, 0.0);
        walk_contact_.assign(P, 3);
        const int64_t id = 0;
        lmhf::check(lmh_walk_plan(&s, &id, 1, 0, zx.data(), zy.data(), walk_foot_coef_.data(), walk_contact_.data(), nullptr), "ZMP::walkZMP");
        zmpXRef_ = Eigen::VectorXd::Zero(Z); zmpYRef_ = Eigen::VectorXd::Zero(Z);
        for (int i = 0; i < Z; ++i) { zmpXRef_(i) = zx[i]; zmpYRef_(i) = zy[i]; }
        walk_phases_ = P;
    }

    const Eigen::VectorXd getZmpXRef() const { return zmpXRef_; }
    const Eigen::VectorXd getZmpYRef() const { return zmpYRef_; }

    // (not upstream) what walkZMP() planned besides the reference: phases, phase length, foot polynomials, contact masks
    int walkPhases() const { return walk_phases_; }
    double walkPhaseTime() const { return timePerStep_; }
    const std::vector<double>& walkFootCoefficients() const { return walk_foot_coef_; }     // [phase][foot][20]
    const std::vector<int32_t>& walkContactMasks() const { return walk_contact_; }         // bit 0 right, bit 1 left

private:
    Eigen::VectorXd zmpXRef_;
    Eigen::VectorXd zmpYRef_;
    double simulationTime_ = 1;
    double timeStep_ = 0.01;
    SupportFoot supportFoot_ = SupportFoot::Double;
    Task task_ = Task::Stand;
    int numSteps_ = 1;
    double timePerStep_ = 0.5;
    int walk_phases_ = 0;
    std::vector<double> walk_foot_coef_;
    std::vector<int32_t> walk_contact_;
};
