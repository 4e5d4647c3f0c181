// linkInertia.hpp -- inertial data of one link (reference: include/linearMpcHumanoid/robotInfo/linkInertia.hpp:4-9).
#pragma once
#include "linearMpcHumanoid/general/lmhEigen.hpp"

struct linkInertia {
    double mass;
    Eigen::Vector3d com;
    Eigen::Matrix3d inertia;
};
