// Task.hpp -- same enumerations as the reference (include/linearMpcHumanoid/general/Task.hpp:3-13).
#pragma once

enum class Task { Stand, Walk, Jump };

enum class SupportFoot { Right, Left, Double };
