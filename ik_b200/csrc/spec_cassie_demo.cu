// Specialised solve kernel: the task set of the reference's own demo (ik_ros/src/cassie.cpp:43-81) -- LeftFootFront
// Position relative to the moving `pelvis` frame, pelvis Full in `universe`, AlignAxisTask (foot y-axis) in `universe`:
// 10 task rows.  Two warp roles: pelvis pose + solve | the two left-foot tasks (one FK of the left leg).
#include "dls_spec.cuh"
#include "gen/cassie_demo.cuh"

namespace ikb {
namespace {
using S = SpecCassieDemo;
template <typename T> int launch(const SpecHostConsts &hc, const SolveArgs<T> &a, int variant, long long n, int sms, cudaStream_t s) {
    return variant == SPEC_TAIL ? launch_spec_tail<S, T>(hc, a, n, sms, s) : launch_spec_bulk<S, T>(hc, a, n, sms, s);
}
int l64(const SpecHostConsts &hc, const SolveArgs<double> &a, int v, long long n, int sms, cudaStream_t s) { return launch<double>(hc, a, v, n, sms, s); }
int l32(const SpecHostConsts &hc, const SolveArgs<float> &a, int v, long long n, int sms, cudaStream_t s) { return launch<float>(hc, a, v, n, sms, s); }
}  // namespace
extern const SpecializedKernel kSpecCassieDemo = {S::name(), spec_matches<S>, l64, l32, spec_near_miss<S>};
}  // namespace ikb
