// Clock.hpp -- the simulation clock of the offline app (reference: include/linearMpcHumanoid/general/Clock.hpp:3-18,
// src/Clock.cpp); header-only here, same constructor and accessors.
//
// Why such a small class matters for parity: step() accumulates t += dt in FP64, exactly as upstream, and the MPC's ZMP
// window index k = (int)(t / dt) (mpcLinearPendulum.cpp:92) and the walking extension's phase index are computed from
// that ACCUMULATED value.  0.01 added 100 times is not 1.0, so a clock that multiplied (t = n * dt) would select other
// windows.  The fused device integrator (lmh_step) keeps one such clock per environment with the same additions
// (__dadd_rn, no FMA contraction); the tests compare the resulting integer sequences bit for bit.
#pragma once

class Clock {
public:
    Clock(double dt, double simulationTime) : dt_(dt), T_(simulationTime) {}

    double getTime() const { return t_; }                 // accumulated time
    double getTimeStep() const { return dt_; }
    double getSimulationTime() const { return T_; }      // the stand() loop runs while |t - T| > 0.01 (main.cpp:74)

    void step() { t_ += dt_; }

private:
    double t_ = 0;
    double dt_;
    double T_;
};
