// Mirrors the reference include path ik/common.hpp (dazzmo/ik ik/ik/common.hpp); everything lives in ik/ik.hpp.
#pragma once
#include "ik.hpp"
