// Mirrors the reference include path ik/centre_of_mass.hpp (dazzmo/ik ik/ik/centre_of_mass.hpp); everything lives in ik/ik.hpp.
#pragma once
#include "ik.hpp"
