// Task.hpp -- the two enumerations the class surface is parameterised by (reference:
// include/linearMpcHumanoid/general/Task.hpp:3-13; enumerator names and order are the API and are kept).
//
//   Task         what the ZMP reference describes.  Stand: regulation about a constant ZMP (ZMP::stanceZMP).  Walk: the
//                walking extension of this library (ZMP::walkZMP -> lmh_walk_plan: footsteps, swing polynomials, contact
//                masks; upstream declares it but never defines it).  Jump: declared upstream, implemented nowhere.
//   SupportFoot  where a standing robot keeps its ZMP: under the right sole (y = -0.05), under the left sole
//                (y = +0.05) or between the two (Double, y = 0), see zmpGeneration.hpp.
#pragma once

enum class Task {
    Stand,   // regulation (apps/offline/main.cpp:21)
    Walk,    // EXTENSION here
    Jump     // not implemented (upstream neither)
};

enum class SupportFoot {
    Right,   // zmp_y = -0.05
    Left,    // zmp_y = +0.05
    Double   // zmp_y = 0
};
