// Registry of specialised kernels (see specialized.hpp).  Each specialisation is one generated header
// (ik_b200/csrc/gen/<name>.cuh, emitted at build time by tools/gen_kernel.py from ik_b200/specs/<name>.json and the
// product's own URDF flattener) compiled in its own translation unit (spec_<name>.cu) so they build in parallel.
#include "specialized.hpp"

#include <cstdlib>

namespace ikb {

extern const SpecializedKernel kSpecCassieFeetPelvis;
extern const SpecializedKernel kSpecManipulatorTool;
extern const SpecializedKernel kSpecHumanoidLimbs;
extern const SpecializedKernel kSpecCassieDemo;
extern const SpecializedKernel kSpecCassieDemoPosture;

namespace {
const SpecializedKernel *const kRegistry[] = {&kSpecCassieFeetPelvis, &kSpecManipulatorTool, &kSpecHumanoidLimbs, &kSpecCassieDemo, &kSpecCassieDemoPosture, nullptr};
}

const SpecializedKernel *const *specialized_registry() { return kRegistry; }

const SpecializedKernel *find_specialized(const HostProblem &hp) {
    // IKB_FORCE_GENERIC=1 pins the table-driven kernel (used by the parity tests to cover both paths)
    const char *force = std::getenv("IKB_FORCE_GENERIC");
    if (force && force[0] == '1') return nullptr;
    for (const SpecializedKernel *const *k = kRegistry; *k; ++k)
        if ((*k)->matches(hp)) return *k;
    return nullptr;
}

const SpecializedKernel *find_near_miss(const HostProblem &hp, std::string *why) {
    for (const SpecializedKernel *const *k = kRegistry; *k; ++k)
        if ((*k)->near_miss && (*k)->near_miss(hp, why)) return *k;
    return nullptr;
}

}  // namespace ikb
