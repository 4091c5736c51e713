// Mirrors the reference include path ik/pik.hpp (dazzmo/ik ik/ik/pik.hpp); everything lives in ik/ik.hpp.
#pragma once
#include "ik.hpp"
