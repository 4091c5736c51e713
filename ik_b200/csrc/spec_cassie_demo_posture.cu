// Specialised solve kernel: the FULL declared task set of the reference's demo (ik_ros/src/cassie.cpp:43-81 with the
// commented-out posture line enabled) -- cassie_demo plus a PostureTask (posture.hpp:17-86) on the 16 revolutes at priority
// level 1: 26 stacked rows, stop test on the 10 priority-0 rows.  Two warp roles: pelvis pose + solve | left-foot tasks + posture.
#include "dls_spec.cuh"
#include "gen/cassie_demo_posture.cuh"

namespace ikb {
namespace {
using S = SpecCassieDemoPosture;
template <typename T> int launch(const SpecHostConsts &hc, const SolveArgs<T> &a, int variant, long long n, int sms, cudaStream_t s) {
    return variant == SPEC_TAIL ? launch_spec_tail<S, T>(hc, a, n, sms, s) : launch_spec_bulk<S, T>(hc, a, n, sms, s);
}
int l64(const SpecHostConsts &hc, const SolveArgs<double> &a, int v, long long n, int sms, cudaStream_t s) { return launch<double>(hc, a, v, n, sms, s); }
int l32(const SpecHostConsts &hc, const SolveArgs<float> &a, int v, long long n, int sms, cudaStream_t s) { return launch<float>(hc, a, v, n, sms, s); }
}  // namespace
extern const SpecializedKernel kSpecCassieDemoPosture = {S::name(), spec_matches<S>, l64, l32, spec_near_miss<S>};
}  // namespace ikb
