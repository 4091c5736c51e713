// Mirrors the reference include path ik/problem.hpp (dazzmo/ik ik/ik/problem.hpp); everything lives in ik/ik.hpp.
#pragma once
#include "ik.hpp"
