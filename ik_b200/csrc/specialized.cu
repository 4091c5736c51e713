// Registry of specialised kernels (see specialized.hpp).
#include "specialized.hpp"

#include <cstdlib>

namespace ikb {

namespace {
const SpecializedKernel *const kRegistry[] = {nullptr};
}

const SpecializedKernel *find_specialized(const HostProblem &hp) {
    // IKB_FORCE_GENERIC=1 pins the table-driven kernel (used by the parity tests to cover both paths)
    const char *force = std::getenv("IKB_FORCE_GENERIC");
    if (force && force[0] == '1') return nullptr;
    for (const SpecializedKernel *k : kRegistry)
        if (k && k->matches(hp)) return k;
    return nullptr;
}

}  // namespace ikb
