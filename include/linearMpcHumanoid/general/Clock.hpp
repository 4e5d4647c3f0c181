// Clock.hpp -- the simulation clock of the offline app (reference: include/linearMpcHumanoid/general/Clock.hpp:3-18,
// src/Clock.cpp).  Header-only here.  step() accumulates t += dt exactly as upstream: the MPC window index
// k = (int)(t / dt) depends on that accumulated value bit for bit.
#pragma once

class Clock {
public:
    Clock(double dt, double simulationTime) : dt_(dt), T_(simulationTime) {}

    double getTime() const { return t_; }
    double getTimeStep() const { return dt_; }
    double getSimulationTime() const { return T_; }

    void step() { t_ += dt_; }

private:
    double t_ = 0;
    double dt_;
    double T_;
};
