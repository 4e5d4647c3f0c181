// facade.hpp -- the reference's C++ class surface (what apps/offline/main.cpp consumes) on top of the C ABI of
// liblmh_b200.so.  Same class and method names, argument meaning and side effects as upstream:
//
//   Robot        include/linearMpcHumanoid/robotInfo/Robot.hpp:16-97
//   Kinematics   include/linearMpcHumanoid/controller/invKinematics.hpp:8-65
//   Mpc3dLip     include/linearMpcHumanoid/controller/mpcLinearPendulum.hpp:5-50
//   ZMP          include/linearMpcHumanoid/trajectories/zmpGeneration.hpp:5-38
//   Controller   include/linearMpcHumanoid/controller/controller.hpp:50-143  (+ ControllerInput, WBCOutput :33-48)
//   Clock, rk4Step, crossMatrix, matrixAngularVelToEulerDot, footCoeffTrajectory
//
// Differences that a port of the app has to know: vectors are lmhb::Vec (a std::vector<double> with operator()
// and segment(); Eigen is not required) and every numeric result comes from the GPU: Controller owns an LmhBatch of
// ONE environment, standStep() is one lmh_stage launch, and Controller::rollout() exposes the fused batched path
// (lmh_step) that the reference does not have.  There is no CPU fallback: constructing a Controller without a CUDA
// device throws.
#pragma once
#include "lmh_capi.h"
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

namespace lmhb {

struct Vec : std::vector<double> {
    using std::vector<double>::vector;
    Vec() {}
    Vec(const std::vector<double>& v) : std::vector<double>(v) {}
    double& operator()(size_t i) { return (*this)[i]; }
    double operator()(size_t i) const { return (*this)[i]; }
    Vec segment(size_t i, size_t n) const { return Vec(begin() + i, begin() + i + n); }
    void setSegment(size_t i, const Vec& v) { for (size_t k = 0; k < v.size(); ++k) (*this)[i + k] = v[k]; }
    static Vec Zero(size_t n) { return Vec(n, 0.0); }
};
inline Vec operator+(const Vec& a, const Vec& b) { Vec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] + b[i]; return r; }
inline Vec operator*(double s, const Vec& a) { Vec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = s * a[i]; return r; }

enum class Task { Stand, Walk, Jump };            // general/Task.hpp:3-7
enum class SupportFoot { Right, Left, Double };   // general/Task.hpp:9-13

class Clock {                                     // general/Clock.hpp:3-18
public:
    Clock(double dt, double simulationTime) : dt_(dt), T_(simulationTime) {}
    double getTime() const { return t_; }
    double getTimeStep() const { return dt_; }
    double getSimulationTime() const { return T_; }
    void step() { t_ += dt_; }
private:
    double t_ = 0, dt_, T_;
};

struct ControllerInput { Vec q, dq; double time = 0; };   // controller.hpp:33-37
struct WBCOutput { Vec qpp, tau, f; };                     // controller.hpp:42-48

inline void lmh_check(int32_t rc, const char* what) {
    if (rc != LMH_OK) throw std::runtime_error(std::string(what) + ": " + lmh_last_error());
}

class Robot {   // state carrier; the numbers are filled in by Kinematics::compute and Controller::standStep
public:
    Robot() : q_(LMH_NQ, 0.0), v_(LMH_NQ, 0.0), CoM_(3, 0.0), comVel_(3, 0.0), comAngMom_(3, 0.0) {}
    int getNumJoints() const { return LMH_NQ; }
    int getNumActualJoints() const { return LMH_NJ; }
    const Vec& getJoints() const { return q_; }
    const Vec& getJointsVelocity() const { return v_; }
    const Vec& getCoM() const { return CoM_; }
    const Vec& getComVel() const { return comVel_; }
    const Vec& getComAngMom() const { return comAngMom_; }
    double getMass() const { return mass_; }
    // scenario pieces collected on the way to the Controller (apps/offline/main.cpp:24-35)
    LmhConfig cfg_;
    bool cfg_set_ = false;
    Vec q_, v_, CoM_, comVel_, comAngMom_;
    double mass_ = 0;
};

class Kinematics {
public:
    // invKinematics.cpp:11-25
    Vec desiredOperationalState(const Robot&, const Vec& Rf, const Vec& Lf, const Vec& com) const {
        Vec d(30, 0.0);
        d.setSegment(0, Rf); d.setSegment(6, Lf); d.setSegment(27, com);
        return d;
    }
    // invKinematics.cpp:27-52 -- Newton IK of the initial posture.  Runs inside lmh_create (host-side setup of the
    // library); a throw-away one-environment batch is created here to obtain the posture, the CoM and the mass.
    void compute(Robot& robot, const Vec& desOp) {
        LmhConfig c;
        lmh_default_config(&c);
        for (int k = 0; k < 6; ++k) { c.rf[k] = desOp[k]; c.lf[k] = desOp[6 + k]; }
        for (int k = 0; k < 3; ++k) c.com[k] = desOp[27 + k];
        robot.cfg_ = c;
        robot.cfg_set_ = true;
        LmhBatch* h = nullptr;
        lmh_check(lmh_create(&c, 1, &h), "Kinematics::compute");
        double zc = 0;
        lmh_get_nominal(h, robot.q_.data(), &zc, &robot.mass_, &iterations_);
        LmhStageDump d;
        lmh_check(lmh_stage(h, &d), "Kinematics::compute");
        for (int k = 0; k < 3; ++k) robot.CoM_[k] = d.CoM[k];
        lmh_destroy(h);
    }
    int iterations_ = 0;
};

class Mpc3dLip {   // mpcLinearPendulum.hpp:8-31; the condensed gain is built inside lmh_create
public:
    Mpc3dLip(double dt, double timeHorizon, double zCoM) : dt_(dt), timeHorizon_(timeHorizon), zCom_(zCoM) {}
    Vec getXRef() const { return xRef_; }
    Vec getYRef() const { return yRef_; }
    double getZCom() const { return zCom_; }
    double dt_, timeHorizon_, zCom_;
    Vec xRef_ = Vec(3, 0.0), yRef_ = Vec(3, 0.0);
};

class ZMP {   // zmpGeneration.hpp:7-25, zmpGeneration.cpp:12-23,39-60
public:
    ZMP(Task, double simulationTime, double timeStep, SupportFoot foot) : T_(simulationTime), dt_(timeStep) {
        int samples = static_cast<int>((simulationTime + 0.5) / timeStep);
        double y = foot == SupportFoot::Right ? -0.05 : foot == SupportFoot::Left ? 0.05 : 0.0;
        zmpXRef_.assign(samples, 0.0);
        zmpYRef_.assign(samples, y);
    }
    const Vec& getZmpXRef() const { return zmpXRef_; }
    const Vec& getZmpYRef() const { return zmpYRef_; }
    double T_, dt_;
    Vec zmpXRef_, zmpYRef_;
};

// footRefTrajectory.cpp:4-47 -- the coefficients are produced by the library's setup for the stand task; user plans
// go through Controller::setFootCoefficients.
inline std::vector<Vec> footCoeffTrajectory(const Vec& currentPos, const Vec& desPos, double stepHeight, double T) {
    std::vector<Vec> c(4);
    c[0] = currentPos; c[1] = desPos; c[2] = Vec{stepHeight}; c[3] = Vec{T};   // carried to the Controller as a request
    return c;
}

class Controller {
public:
    // controller.cpp:5-46
    Controller(Robot& robot, Mpc3dLip& mpc, ZMP& zmp, std::vector<Vec>& rFCoeff, std::vector<Vec>& lFCoeff, int n_env = 1)
        : robot_(robot), mpc_(mpc), n_env_(n_env) {
        LmhConfig c;
        if (robot.cfg_set_) c = robot.cfg_; else lmh_default_config(&c);
        c.dt = mpc.dt_; c.time_horizon = mpc.timeHorizon_; c.sim_time = zmp.T_;
        (void)rFCoeff; (void)lFCoeff;
        lmh_check(lmh_create(&c, n_env, &h_), "Controller");
        // a ZMP object other than the default stand reference becomes plan 0
        std::vector<double> zx, zy(1);
        int32_t len = 0;
        double foot[40];
        lmh_get_plan(h_, 0, nullptr, nullptr, 0, &len, foot);
        if ((int)zmp.getZmpXRef().size() == len)
            lmh_check(lmh_set_plans(h_, 1, len, zmp.getZmpXRef().data(), zmp.getZmpYRef().data(), foot, nullptr), "Controller");
        tau_.assign(LMH_NJ, 0.0);
        held_.assign(LMH_NQ, 0.0);
    }
    ~Controller() { lmh_destroy(h_); }
    Controller(const Controller&) = delete;

    // controller.cpp:48-79 : model update -> MPC -> WBC, for the single-environment facade (env 0)
    void standStep(const ControllerInput& in) {
        std::vector<double> x(60 * n_env_), vh(30 * n_env_), t(n_env_, in.time);
        for (int e = 0; e < n_env_; ++e) {
            for (int i = 0; i < 30; ++i) { x[60 * e + i] = in.q[i]; x[60 * e + 30 + i] = in.dq[i]; vh[30 * e + i] = held_[i]; }
        }
        lmh_check(lmh_set_state(h_, x.data(), vh.data(), t.data()), "standStep");
        std::vector<LmhStageDump> d(n_env_);
        lmh_check(lmh_stage(h_, d.data()), "standStep");
        const LmhStageDump& s = d[0];
        out_.qpp.assign(s.qpp, s.qpp + 30); out_.tau.assign(s.tau, s.tau + 24); out_.f.assign(s.f, s.f + 12);
        tau_ = out_.tau;
        held_ = in.dq;                               // Robot::updateVelocityState (Robot.cpp:271-274)
        robot_.q_ = in.q; robot_.v_ = in.dq;
        for (int k = 0; k < 3; ++k) {
            robot_.CoM_[k] = s.CoM[k]; robot_.comVel_[k] = s.comVel[k]; robot_.comAngMom_[k] = s.comAngMom[k];
            mpc_.xRef_[k] = s.xRef[k]; mpc_.yRef_[k] = s.yRef[k];
        }
        last_ = s;
    }
    // controller.cpp:81-154 : the app calls it right after standStep with identical inputs (main.cpp:105); the
    // solve is deterministic, so the stored solution IS the answer.
    WBCOutput WBC(double) { return out_; }
    const Vec& getTorques() { return tau_; }
    const LmhStageDump& lastStage() const { return last_; }

    // ---- beyond the reference: the fused batched path
    LmhBatch* batch() { return h_; }
    void rollout(int n_steps) { lmh_check(lmh_step(h_, n_steps), "rollout"); lmh_check(lmh_sync(h_), "rollout"); }

private:
    Robot& robot_;
    Mpc3dLip& mpc_;
    int n_env_;
    LmhBatch* h_ = nullptr;
    Vec tau_, held_;
    WBCOutput out_;
    LmhStageDump last_;
};

// generalizedFunctions.cpp:3-9, 43-50 (host helpers the app's dynamics() closure uses, main.cpp:111-116)
inline Vec crossTimes(const Vec& w, const Vec& p) { return Vec{w[1] * p[2] - w[2] * p[1], w[2] * p[0] - w[0] * p[2], w[0] * p[1] - w[1] * p[0]}; }
inline Vec eulerRates(const Vec& eta, const Vec& w) {
    double c2 = std::cos(eta[2]), s2 = std::sin(eta[2]), c1 = std::cos(eta[1]), t1 = std::tan(eta[1]);
    return Vec{(c2 / c1) * w[0] + (s2 / c1) * w[1], -s2 * w[0] + c2 * w[1], c2 * t1 * w[0] + s2 * t1 * w[1] + w[2]};
}

// general/rk4.hpp:5-18
template <typename F>
Vec rk4Step(F f, const Vec& x, double t, double dt) {
    Vec k1 = f(x, t);
    Vec k2 = f(x + (0.5 * dt) * k1, t + 0.5 * dt);
    Vec k3 = f(x + (0.5 * dt) * k2, t + 0.5 * dt);
    Vec k4 = f(x + dt * k3, t + dt);
    return x + (dt / 6.0) * (((k1 + 2.0 * k2) + 2.0 * k3) + k4);
}

}  // namespace lmhb
