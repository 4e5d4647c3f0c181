// lmh_scenarios.hpp -- host-side scenario generators behind the C ABI (run once per batch, never per step):
// per-environment push impulses (BASELINE config 2) and walking plans (configs 3-4: ZMP::walkZMP, which upstream
// declares but never defines -- include/linearMpcHumanoid/trajectories/zmpGeneration.hpp:15-22 -- plus the per-phase
// swing polynomials of src/footRefTrajectory.cpp:4-47 and the contact masks hinted at in controller.hpp:83-98).
// Everything is keyed by the GLOBAL environment id through a counter-based generator, so a batch receives the same
// per-environment inputs however it is sharded across GPUs.  The Python package calls these same functions.
#pragma once
#include "lmh_setup.hpp"
#include <cstdint>

namespace lmh {

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// uniform [0,1) from (seed, environment id, stream id): the 53 high bits of a twice-mixed counter
inline double u01(uint64_t seed, int64_t env, uint64_t stream) {
    uint64_t h = splitmix64(seed * 0x100000001B3ull + (uint64_t)env * 0x9E3779B1ull + stream * 0x85EBCA77ull);
    h = splitmix64(h);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

// config 2: one velocity impulse on the base linear x,y velocity; magnitude ~U(0, max_speed), direction ~U(0, 2 pi),
// at step ~U{0..max_step}
inline void push_plan(const int64_t* env_ids, int n, uint64_t seed, double max_speed, int max_step, int32_t* step, double* dv6) {
    const double kPi = 3.14159265358979323846;
    for (int i = 0; i < n; ++i) {
        const int64_t e = env_ids[i];
        const double mag = max_speed * u01(seed, e, 0), ang = 2.0 * kPi * u01(seed, e, 1);
        step[i] = (int32_t)(u01(seed, e, 2) * (double)(max_step + 1));
        for (int k = 0; k < 6; ++k) dv6[6 * i + k] = 0.0;
        dv6[6 * i + 0] = mag * std::cos(ang);
        dv6[6 * i + 1] = mag * std::sin(ang);
    }
}

struct WalkSpec {
    int n_footsteps = 4;
    double time_per_step = 0.4, dt = 0.01;
    int horizon_steps = 160;
    double step_height = 0.02, lateral = 0.05, min_len = 0.02, max_len = 0.05;
    int tail_steps = 40;
    int n_phases() const { return n_footsteps + 2; }                                  // double support before and after
    int per() const { return (int)std::nearbyint(time_per_step / dt); }               // samples per phase
    int zmp_len() const { return n_phases() * per() + tail_steps + horizon_steps + 2; }
};

// One environment's walking plan: n_footsteps steps (right foot first) of length L, one phase of double support before
// and after.  zmp_x / zmp_y [zmp_len]: piecewise constant, under the stance foot in single support and between the
// feet in double support; foot_coef [P][2][20]; contact [P] (bit 0 right foot, bit 1 left foot).
inline void walk_plan_one(const WalkSpec& s, double L, double* zx, double* zy, double* fc, int32_t* ct) {
    const int P = s.n_phases(), per = s.per(), zlen = s.zmp_len();
    double foot[2][3] = {{0.0, -s.lateral, 0.0}, {0.0, s.lateral, 0.0}};              // right, left
    for (int p = 0; p < P; ++p) {
        const int swing = (p == 0 || p == P - 1) ? -1 : (p - 1) % 2;                    // the right foot swings first
        ct[p] = swing < 0 ? 3 : (swing == 0 ? 2 : 1);
        double target[2][3];
        for (int f = 0; f < 2; ++f) for (int k = 0; k < 3; ++k) target[f][k] = foot[f][k];
        if (swing >= 0) target[swing][0] = foot[1 - swing][0] + L;
        for (int f = 0; f < 2; ++f)
            foot_coefficients(foot[f], target[f], f == swing ? s.step_height : 0.0, s.time_per_step, fc + ((size_t)p * 2 + f) * 20);
        double z[2];
        for (int k = 0; k < 2; ++k) z[k] = swing < 0 ? 0.5 * (foot[0][k] + foot[1][k]) : foot[1 - swing][k];
        const int k0 = p * per, k1 = p < P - 1 ? (p + 1) * per : zlen;
        for (int k = k0; k < k1; ++k) { zx[k] = z[0]; zy[k] = z[1]; }
        for (int f = 0; f < 2; ++f) for (int k = 0; k < 3; ++k) foot[f][k] = target[f][k];
    }
}

// config 3: step length ~U(min_len, max_len) per environment, or `fixed_len` for every environment when it is >= 0
inline void walk_plans(const WalkSpec& s, const int64_t* env_ids, int n, uint64_t seed, double fixed_len, double* zx, double* zy,
                       double* fc, int32_t* ct, double* step_len) {
    const int P = s.n_phases(), zlen = s.zmp_len();
    for (int i = 0; i < n; ++i) {
        const double L = fixed_len >= 0.0 ? fixed_len : s.min_len + (s.max_len - s.min_len) * u01(seed, env_ids[i], 0);
        if (step_len) step_len[i] = L;
        walk_plan_one(s, L, zx + (size_t)i * zlen, zy + (size_t)i * zlen, fc + (size_t)i * P * 40, ct + (size_t)i * P);
    }
}

}  // namespace lmh
