// Mirrors the reference include path ik/posture.hpp (dazzmo/ik ik/ik/posture.hpp); everything lives in ik/ik.hpp.
#pragma once
#include "ik.hpp"
