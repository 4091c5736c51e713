// Mirrors the reference include path ik/frame.hpp (dazzmo/ik ik/ik/frame.hpp); everything lives in ik/ik.hpp.
#pragma once
#include "ik.hpp"
