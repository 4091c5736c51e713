// Mirrors the reference include path ik/dls.hpp (dazzmo/ik ik/ik/dls.hpp); everything lives in ik/ik.hpp.
#pragma once
#include "ik.hpp"
