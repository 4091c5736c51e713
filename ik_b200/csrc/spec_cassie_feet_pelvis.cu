// Specialised solve kernel: Cassie (free-flyer) with pelvis Full + LeftFootFront / RightFootFront Position tasks in
// `universe` -- the BASELINE.json headline problem (reference task set-up: ik_ros/src/cassie.cpp:43-81).
#include "dls_spec.cuh"
#include "gen/cassie_feet_pelvis.cuh"

namespace ikb {
namespace {
using S = SpecCassieFeetPelvis;
int l64(const SpecHostConsts &hc, const SolveArgs<double> &a, int sms, cudaStream_t s) { return launch_spec<S, double>(hc, a, sms, s); }
int l32(const SpecHostConsts &hc, const SolveArgs<float> &a, int sms, cudaStream_t s) { return launch_spec<S, float>(hc, a, sms, s); }
}  // namespace
extern const SpecializedKernel kSpecCassieFeetPelvis = {S::name(), spec_matches<S>, l64, l32};
}  // namespace ikb
