// Specialised solve kernels: 32-DoF free-flyer humanoid with torso pose + four end-effector poses, all Full frame
// tasks in `universe` (BASELINE.json config 4: 30 task rows).  Five warp roles: torso and one per limb; every role steps
// the coordinates its own evaluate reads (Spec::DSTEP).  Two solves are compiled (tools/gen_kernel.py):
//   arrow    (default, r2) the bordered-block-diagonal step: the free-flyer and the torso / chest joints are the only
//            columns several roles touch, so every role factorises its own 6 x 6 block D_a = C_a C_a^T + l^2 I and the
//            roles meet in ONE 8 x 8 system (gen_solve_arrow) -- no dense 30 x 30 factor, one group barrier;
//   uniform  (r1) the dense factorisation distributed over the roles with one body (gen_solve_uniform), 7 barriers.
// IKB_HUMANOID_SOLVE=uniform|arrow selects one for A/B runs.
// (Spec option "arrow_solver_skips_cap" -- the torso role leaves the 8 x 8 shared-column system to the limb roles, as the
// pelvis role does for Cassie -- is off here: 262 144 problems take 13.91 ms with it against 13.55 ms without,
// tools/humanoid_lone.py.)
#include <cstdlib>
#include <cstring>

#include "dls_spec.cuh"
#include "gen/humanoid_limbs.cuh"
#include "gen/humanoid_limbs_arrow.cuh"

namespace ikb {
namespace {
using SU = SpecHumanoidLimbs;
using SA = SpecHumanoidLimbsArrow;
// (A third build -- spec option "tmem_j": the Jacobian strip of every role in tensor memory, which frees enough shared
// memory for two groups per SM in FP64 -- was measured and dropped: ten warps get 168 registers each, the body spills 1.7 KB
// per thread and 262 144 problems take 21.1 ms against 13.5 ms.  Generator and kernel support remain: TStrip, TMEMJ.)
int variant_of() {   // IKB_HUMANOID_SOLVE=uniform|arrow
    const char *e = std::getenv("IKB_HUMANOID_SOLVE");
    return (e && std::strcmp(e, "uniform") == 0) ? 0 : 1;
}
template <class S, typename T> int launch_s(const SpecHostConsts &hc, const SolveArgs<T> &a, int variant, long long n, int sms, cudaStream_t s) {
    return variant == SPEC_TAIL ? launch_spec_tail<S, T>(hc, a, n, sms, s) : launch_spec_bulk<S, T>(hc, a, n, sms, s);
}
template <typename T> int launch(const SpecHostConsts &hc, const SolveArgs<T> &a, int variant, long long n, int sms, cudaStream_t s) {
    switch (variant_of()) {
        case 0: return launch_s<SU, T>(hc, a, variant, n, sms, s);
        default: return launch_s<SA, T>(hc, a, variant, n, sms, s);
    }
}
int l64(const SpecHostConsts &hc, const SolveArgs<double> &a, int v, long long n, int sms, cudaStream_t s) { return launch<double>(hc, a, v, n, sms, s); }
int l32(const SpecHostConsts &hc, const SolveArgs<float> &a, int v, long long n, int sms, cudaStream_t s) { return launch<float>(hc, a, v, n, sms, s); }
}  // namespace
extern const SpecializedKernel kSpecHumanoidLimbs = {SU::name(), spec_matches<SU>, l64, l32, spec_near_miss<SU>};
}  // namespace ikb
