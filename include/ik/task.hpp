// Mirrors the reference include path ik/task.hpp (dazzmo/ik ik/ik/task.hpp); everything lives in ik/ik.hpp.
#pragma once
#include "ik.hpp"
