// Robot.hpp -- the reference's Robot class (include/linearMpcHumanoid/robotInfo/Robot.hpp:16-97, src/Robot.cpp) with
// the same public members, computed by the GPU library: updateState() is one lmh_kinematics launch (forward kinematics
// of all 28 frames, parent-relative motion transforms, centre of mass), the link table is lmh_get_link_tables /
// lmh_get_link_inertia.  State carried exactly as upstream: q_ (generalized coordinates), v_ (the velocity the Robot
// HOLDS: refreshed only by updateVelocityState, which Controller::standStep calls after the dynamics -- see
// controller.cpp:48-79 and the "held velocity" note in DESIGN.md).
#pragma once
#include "linearMpcHumanoid/general/lmhEigen.hpp"
#include <cmath>
#include <vector>
#include "linearMpcHumanoid/robotInfo/linkInertia.hpp"
#include "linearMpcHumanoid/robotInfo/robotParameters.hpp"
#include "linearMpcHumanoid/general/generalizedFunctions.hpp"
#include "linearMpcHumanoid/lmhEngine.hpp"

#define NUM_JOINTS 30         // 24 rotational joints plus the 6 coordinates of the floating base
#define NUM_ACTUAL_JOINTS 24
#define NUM_FRAMES 28         // base, 24 joint frames, 2 soles, 1 extra head frame
#define NUM_BODIES 25

Eigen::VectorXd initialConfiguration();
std::vector<Eigen::Matrix4d> matTrans(std::vector<double> theta);

class Robot {
public:
    Robot() {
        LmhBatch* h = lmhf::Engine::get().handle();
        double m[28], c[28 * 3], I[28 * 9];
        lmhf::check(lmh_get_link_tables(h, m, c, nullptr), "Robot");
        lmhf::check(lmh_get_link_inertia(h, I), "Robot");
        links_.resize(NUM_FRAMES);
        mass_ = 0;
        for (int i = 0; i < NUM_FRAMES; ++i) {        // joint-frame inertial data, as the upstream constructor leaves it (Robot.cpp:14-22)
            links_[i].mass = m[i];
            links_[i].com = Eigen::Vector3d(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
            links_[i].inertia = Eigen::Matrix3d::Zero();
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < 3; ++k) links_[i].inertia(r, k) = I[9 * i + 3 * r + k];
            mass_ += m[i];
        }
        T_.resize(NUM_FRAMES);
        X_.resize(NUM_FRAMES);
        CoM_ = Eigen::Vector3d::Zero(); comVel_ = Eigen::Vector3d::Zero(); comAngMom_ = Eigen::Vector3d::Zero();
        updateState(initialConfiguration());
        Rf_q0_ = Eigen::Matrix3d::Zero();             // sole frame orientation at q = 0 (Robot.cpp:28-31)
        Rf_q0_(0, 2) = 1; Rf_q0_(1, 1) = -1; Rf_q0_(2, 0) = 1;
        Lf_q0_ = Rf_q0_;
        v_ = Eigen::VectorXd::Zero(NUM_JOINTS);
        S_ = Eigen::VectorXd::Zero(6);
        S_(2) = 1;
        const double fv[4][3] = {{0.1, 0.025, 0}, {0.1, -0.025, 0}, {-0.05, 0.025, 0}, {-0.05, -0.025, 0}};   // Robot.cpp:38-42
        for (int k = 0; k < 4; ++k) footVertices_.push_back(Eigen::Vector3d(fv[k][0], fv[k][1], fv[k][2]));
    }

    const std::vector<int> parentFrame() const {     // Robot.cpp:162-167
        return {-1, 0, 1, 2, 3, 4, 5, 6, 0, 8, 9, 10, 11, 12, 13, 0, 15, 16, 17, 18, 0, 20, 21, 22, 23, 0, 25, 26};
    }
    const std::vector<int> actuatedFrames() const {  // Robot.cpp:169-174
        return {0, 1, 2, 3, 4, 5, 6, 0, 7, 8, 9, 10, 11, 12, 0, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 0};
    }

    int getNumFrames() const { return numFrames; }
    int getNumJoints() const { return numJoints; }
    int getNumActualJoints() const { return numActualJoints; }
    int getNumBodies() const { return numBodies; }
    const Eigen::VectorXd& getJoints() const { return q_; }
    const Eigen::VectorXd& getJointsVelocity() const { return v_; }
    const std::vector<Eigen::Matrix4d>& getT() const { refreshFrames(); return T_; }
    const std::vector<linkInertia>& getLinks() const { return links_; }
    const Eigen::Vector3d& getCoM() const { return CoM_; }
    const Eigen::Vector3d& getComVel() const { return comVel_; }
    const Eigen::Vector3d& getComAngMom() const { return comAngMom_; }
    double getMass() const { return mass_; }
    Eigen::Matrix3d getRf_q0() const { return Rf_q0_; }
    Eigen::Matrix3d getLf_q0() const { return Lf_q0_; }
    const std::vector<Eigen::MatrixXd>& getX() const { refreshFrames(); return X_; }
    Eigen::VectorXd getS() const { return S_; }

    // Robot.cpp:225-238 / 45-160: both come out of the same device pass as updateState
    void computeCoM() { frames_valid_ = false; refreshFrames(); }
    void forwardKinematics() { frames_valid_ = false; refreshFrames(); }

    // h = AG [w; v]_base: CoM velocity and centroidal angular momentum (Robot.cpp:300-310)
    void computeComMomentum(Eigen::VectorXd v, const Eigen::MatrixXd& AG) {
        swapBaseVelocityAndRefToWorldFrame(getX()[0], v);
        for (int k = 0; k < 6; ++k) {
            double s = 0.0;
            for (int j = 0; j < NUM_JOINTS; ++j) s += AG(k, j) * v(j);
            if (k < 3) comAngMom_(k) = s; else comVel_(k - 3) = s / getMass();
        }
    }

    void updateState(const Eigen::VectorXd& q_new) {                              // Robot.cpp:264-269
        q_ = q_new;
        frames_valid_ = false;
        refreshFrames();
    }
    void updateVelocityState(const Eigen::VectorXd& v_new, const Eigen::MatrixXd& AG) {   // Robot.cpp:271-274
        v_ = v_new;
        computeComMomentum(v_, AG);
    }

    std::vector<Eigen::Matrix4d> parentTransMatrix(const std::vector<Eigen::Matrix4d>& T) {   // Robot.cpp:276-287
        std::vector<Eigen::Matrix4d> rel(T.size());
        const std::vector<int> parent = parentFrame();
        rel[0] = T[0];
        for (size_t i = 1; i < T.size(); ++i) rel[i] = inverseTransformationMatrix(T[parent[i]]) * T[i];
        return rel;
    }
    void allVelocityMatrices(const std::vector<Eigen::Matrix4d>& piTi) {          // Robot.cpp:289-298
        X_.resize(piTi.size());
        for (size_t i = 0; i < piTi.size(); ++i) X_[i] = velocityMatrix(piTi[i]);
    }

    const std::vector<Eigen::Vector3d>& getFootVertices() const { return footVertices_; }
    const Eigen::VectorXd desiredPosture() { return initialConfiguration(); }     // Robot.cpp:253-262 (same literals)

private:
    friend class Controller;
    // T_, X_, CoM_ of the current q_ (one lmh_kinematics launch); Controller::standStep marks them stale instead of paying
    // for a second launch per stage evaluation, and the getters refresh on demand
    void refreshFrames() const {
        if (frames_valid_) return;
        double q[NUM_JOINTS], T[NUM_FRAMES * 16], X[NUM_FRAMES * 36], com[3];
        for (int i = 0; i < NUM_JOINTS; ++i) q[i] = q_(i);
        lmhf::check(lmh_kinematics(lmhf::Engine::get().handle(), 1, q, T, X, com), "Robot::updateState");
        for (int i = 0; i < NUM_FRAMES; ++i) {
            T_[i] = Eigen::Matrix4d::Zero();
            X_[i] = Eigen::MatrixXd::Zero(6, 6);
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T_[i](r, c) = T[16 * i + 4 * r + c];
            for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) X_[i](r, c) = X[36 * i + 6 * r + c];
        }
        for (int k = 0; k < 3; ++k) CoM_(k) = com[k];
        frames_valid_ = true;
    }

    int numJoints = NUM_JOINTS;
    int numActualJoints = NUM_ACTUAL_JOINTS;
    int numFrames = NUM_FRAMES;
    int numBodies = NUM_BODIES;
    Eigen::VectorXd q_;   // generalized coordinates [p(3), roll/pitch/yaw(3), joints(24)]
    Eigen::VectorXd v_;   // generalized velocities held by the robot (v != dq/dt: spatial base twist)
    mutable std::vector<Eigen::Matrix4d> T_;
    std::vector<linkInertia> links_;
    mutable std::vector<Eigen::MatrixXd> X_;
    mutable Eigen::Vector3d CoM_;
    Eigen::Vector3d comVel_;
    Eigen::Vector3d comAngMom_;
    double mass_;
    Eigen::Matrix3d Rf_q0_;
    Eigen::Matrix3d Lf_q0_;
    Eigen::VectorXd S_;
    std::vector<Eigen::Vector3d> footVertices_;
    mutable bool frames_valid_ = false;
};

// Robot.cpp:242-251
inline Eigen::VectorXd initialConfiguration() {
    static const double q0[NUM_JOINTS] = {-0.0185, 0, 0.282, 0, 0, 0, 0, 0, -0.5, 0.8, -0.3, 0, 0, 0, -0.5,
                                          0.8, -0.3, 0, 1.6, 0, 0, 0, 0, -1.6, 0, 0, 0, 0, 0, 0};
    Eigen::VectorXd q = Eigen::VectorXd::Zero(NUM_JOINTS);
    for (int i = 0; i < NUM_JOINTS; ++i) q(i) = q0[i];
    return q;
}

// Robot.cpp:176-223: the 25 local modified-DH transforms
inline std::vector<Eigen::Matrix4d> matTrans(std::vector<double> theta) {
    if (theta.size() < 25) throw std::invalid_argument("matTrans: 25 joint angles expected");
    double T[25 * 16];
    lmhf::check(lmh_local_transforms(theta.data(), T), "matTrans");
    std::vector<Eigen::Matrix4d> out(25);
    for (int i = 0; i < 25; ++i) {
        out[i] = Eigen::Matrix4d::Zero();
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[i](r, c) = T[16 * i + 4 * r + c];
    }
    return out;
}
