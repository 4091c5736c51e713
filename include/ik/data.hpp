// Mirrors the reference include path ik/data.hpp (dazzmo/ik ik/ik/data.hpp); everything lives in ik/ik.hpp.
#pragma once
#include "ik.hpp"
