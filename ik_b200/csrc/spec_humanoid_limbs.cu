// Specialised solve kernel: 32-DoF free-flyer humanoid with torso pose + four end-effector poses, all Full frame
// tasks in `universe` (BASELINE.json config 4: 30 task rows).  Five warp roles: torso (+ back substitution) and one per
// limb; the factorisation runs on all five with ONE body (rows r + 5m by run-time role index, tools/gen_kernel.py
// gen_solve_uniform) and every role steps the coordinates its own evaluate reads (Spec::DSTEP).  The per-problem strips
// (303 J non-zeros + 525 factor / rhs / Gram-diagonal + 60 target scalars) allow one 32-problem group per SM in FP64
// (two in FP32), so both launch variants are the one-group configuration.
#include "dls_spec.cuh"
#include "gen/humanoid_limbs.cuh"

namespace ikb {
namespace {
using S = SpecHumanoidLimbs;
template <typename T> int launch(const SpecHostConsts &hc, const SolveArgs<T> &a, int variant, long long n, int sms, cudaStream_t s) {
    return variant == SPEC_TAIL ? launch_spec_tail<S, T>(hc, a, n, sms, s) : launch_spec_bulk<S, T>(hc, a, n, sms, s);
}
int l64(const SpecHostConsts &hc, const SolveArgs<double> &a, int v, long long n, int sms, cudaStream_t s) { return launch<double>(hc, a, v, n, sms, s); }
int l32(const SpecHostConsts &hc, const SolveArgs<float> &a, int v, long long n, int sms, cudaStream_t s) { return launch<float>(hc, a, v, n, sms, s); }
}  // namespace
extern const SpecializedKernel kSpecHumanoidLimbs = {S::name(), spec_matches<S>, l64, l32, spec_near_miss<S>};
}  // namespace ikb
