// robotParameters.hpp -- createNaoParameters() (reference: include/linearMpcHumanoid/robotInfo/robotParameters.hpp,
// src/robotParameters.cpp): the raw NAO table, 28 links, frames aligned with the world at q = 0.  The literals live in
// the library (lmh_get_nao_parameters); this is the reference-typed view of them.
#pragma once
#include <vector>
#include "linearMpcHumanoid/robotInfo/linkInertia.hpp"
#include "linearMpcHumanoid/lmhEngine.hpp"

inline std::vector<linkInertia> createNaoParameters() {
    double m[28], c[28 * 3], I[28 * 9];
    lmhf::check(lmh_get_nao_parameters(m, c, I), "createNaoParameters");
    std::vector<linkInertia> links(28);
    for (int i = 0; i < 28; ++i) {
        links[i].mass = m[i];
        links[i].com = Eigen::Vector3d(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
        links[i].inertia = Eigen::Matrix3d::Zero();
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) links[i].inertia(r, k) = I[9 * i + 3 * r + k];
    }
    return links;
}
