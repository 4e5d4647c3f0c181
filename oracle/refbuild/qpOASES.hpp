// ORACLE — TEST INFRASTRUCTURE ONLY.
// API stand-in for qpOASES (third party, absent from this image and from
// /root/reference; version unpinned upstream, CMakeLists.txt:39-41,76), so the
// reference's own controller.cpp / mpcLinearPendulum.cpp compile unmodified.
// Only the calls those two files make exist here.  The solve itself is
// orc::qp_solve_gi (oracle/orc_qp.hpp): an exact dual active-set method, NOT
// qpOASES' parametric active-set homotopy -> the QP boundary stays PARITY
// UNPINNED; both reference QPs are strictly convex so the minimiser is unique.
#pragma once
#include "orc_qp.hpp"
#include <vector>
#include <cstring>

namespace qpOASES {

typedef double real_t;
typedef int int_t;
const real_t INFTY = orc::QP_INFTY;
enum returnValue { SUCCESSFUL_RETURN = 0, RET_MAX_NWSR_REACHED = 64, RET_INIT_FAILED = 33 };
enum PrintLevel { PL_NONE = 0, PL_LOW, PL_MEDIUM, PL_HIGH };

struct Options {
    PrintLevel printLevel = PL_MEDIUM;
    void setToMPC() {}
    void setToDefault() {}
};

class QProblem {
public:
    QProblem() {}
    QProblem(int_t nV, int_t nC) : nV_(nV), nC_(nC), x_(nV, 0.0), y_(nC > 0 ? nC : 1, 0.0) {}
    void setOptions(const Options&) {}
    returnValue init(const real_t* H, const real_t* g, const real_t* A, const real_t* lb, const real_t* ub,
                     const real_t* lbA, const real_t* ubA, int_t& nWSR, real_t* = nullptr) {
        (void)lb; (void)ub;  // the reference never passes variable bounds
        H_.assign(H, H + (size_t)nV_ * nV_);
        if (nC_ > 0) A_.assign(A, A + (size_t)nC_ * nV_);
        return solve(g, lbA, ubA, nWSR);
    }
    returnValue hotstart(const real_t* g, const real_t* lb, const real_t* ub, const real_t* lbA,
                         const real_t* ubA, int_t& nWSR, real_t* = nullptr) {
        (void)lb; (void)ub;
        return solve(g, lbA, ubA, nWSR);
    }
    returnValue getPrimalSolution(real_t* x) const {
        std::memcpy(x, x_.data(), sizeof(real_t) * (size_t)nV_);
        return SUCCESSFUL_RETURN;
    }
    orc::QpInfo lastInfo;
    std::vector<real_t> y_;

protected:
    returnValue solve(const real_t* g, const real_t* lbA, const real_t* ubA, int_t& nWSR) {
        lastInfo = orc::qp_solve_gi<double>(nV_, H_.data(), g, nC_, A_.data(), lbA, ubA, x_.data(), y_.data(), nWSR);
        nWSR = lastInfo.iterations;
        return lastInfo.status == 0 ? SUCCESSFUL_RETURN : (lastInfo.status == 1 ? RET_MAX_NWSR_REACHED : RET_INIT_FAILED);
    }
    int_t nV_ = 0, nC_ = 0;
    std::vector<real_t> H_, A_, x_;
};

class SQProblem : public QProblem {
public:
    SQProblem() {}
    SQProblem(int_t nV, int_t nC) : QProblem(nV, nC) {}
};

}  // namespace qpOASES
