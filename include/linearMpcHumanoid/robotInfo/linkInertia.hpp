// linkInertia.hpp -- inertial data of one link as the class surface hands it out (reference:
// include/linearMpcHumanoid/robotInfo/linkInertia.hpp:4-9; field names are the API).
//
// Two flavours travel in this type, as upstream: createNaoParameters() returns the RAW table (centre of mass and
// inertia in frames aligned with the world at q = 0), Robot::getLinks() the table re-expressed in the joint frames
// (what the dynamics consume; lmh_get_link_tables / lmh_get_link_inertia).  Typos and asymmetries of the upstream
// literals are data and are kept.
#pragma once
#include "linearMpcHumanoid/general/lmhEigen.hpp"

struct linkInertia {
    double mass;               // kg
    Eigen::Vector3d com;       // centre of mass in the link's frame, m
    Eigen::Matrix3d inertia;   // rotational inertia about the centre of mass, kg m^2 (not necessarily symmetric: data)
};
