// Mirrors the reference include path ik/visitor.hpp (dazzmo/ik ik/ik/visitor.hpp); everything lives in ik/ik.hpp.
#pragma once
#include "ik.hpp"
