// lmhEngine.hpp -- internals shared by the class surface (not part of the reference's API): error translation and
// one lazily created single-environment batch of the default scenario.  Robot / Dynamics / Kinematics answer their
// scenario-independent queries (forward kinematics of every frame, mass matrix, bias forces, Jacobians, inverse
// kinematics) through it: one GPU launch per query, results cached while the inputs do not change.  There is no CPU
// fallback behind these classes: without a CUDA device the first query throws.
#pragma once
#include "lmh_capi.h"
#include <cstring>
#include <stdexcept>
#include <string>

namespace lmhf {

inline void check(int32_t rc, const char* what) {
    if (rc != LMH_OK) throw std::runtime_error(std::string(what) + ": " + lmh_last_error());
}

class Engine {
public:
    static Engine& get() {
        static Engine e;
        return e;
    }
    LmhBatch* handle() {
        if (!h_) {
            LmhConfig c;
            lmh_default_config(&c);
            check(lmh_create(&c, 1, &h_), "linearMpcHumanoid (GPU engine)");
        }
        return h_;
    }
    // one dynamics(x, t) evaluation at (q, v) with the Robot's held velocity vh; identical inputs reuse the last result
    const LmhStageDump& stage(const double* q, const double* v, const double* vh, double t) {
        double key[91];
        std::memcpy(key, q, 30 * sizeof(double)); std::memcpy(key + 30, v, 30 * sizeof(double));
        std::memcpy(key + 60, vh, 30 * sizeof(double)); key[90] = t;
        if (!valid_ || std::memcmp(key, key_, sizeof(key)) != 0) {
            check(lmh_set_state(handle(), key, key + 60, key + 90), "linearMpcHumanoid (set_state)");
            check(lmh_stage(handle(), &dump_), "linearMpcHumanoid (stage)");
            std::memcpy(key_, key, sizeof(key));
            valid_ = true;
        }
        return dump_;
    }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    ~Engine() { if (h_) lmh_destroy(h_); }

private:
    Engine() {}
    LmhBatch* h_ = nullptr;
    LmhStageDump dump_;
    double key_[91];
    bool valid_ = false;
};

}  // namespace lmhf
