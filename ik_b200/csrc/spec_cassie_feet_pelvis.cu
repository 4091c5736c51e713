// Specialised solve kernel: Cassie (free-flyer) with pelvis Full + LeftFootFront / RightFootFront Position tasks in
// `universe` -- the BASELINE.json headline problem (reference task set-up: ik_ros/src/cassie.cpp:43-81).
// Three decompositions of the same generated arithmetic are compiled (ik_b200/specs/cassie_feet_pelvis*.json): one
// warp role (W1), two (pelvis+left | right+solve, W2) and three (pelvis+solve | left | right, W3).  The BULK variant
// is the measured-best throughput decomposition per scalar type, the TAIL variant is W3 with one group per CTA
// (shortest critical path per iteration).  IKB_CASSIE_ROLES=1|2|3 overrides the BULK choice (bench_variants.sh).
#include <cstdlib>

#include "dls_spec.cuh"
#include "gen/cassie_feet_pelvis.cuh"
#include "gen/cassie_feet_pelvis_w1.cuh"
#include "gen/cassie_feet_pelvis_w2.cuh"

namespace ikb {
namespace {
using S3 = SpecCassieFeetPelvis;
using S2 = SpecCassieFeetPelvisW2;
using S1 = SpecCassieFeetPelvisW1;
int roles(int dflt) {
    const char *e = std::getenv("IKB_CASSIE_ROLES");
    return (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : dflt;
}
template <typename T> int launch(const SpecHostConsts &hc, const SolveArgs<T> &a, int variant, long long n, int sms, cudaStream_t s,
                                 int bulk_roles) {
    if (variant == SPEC_TAIL) return launch_spec_tail<S3, T>(hc, a, n, sms, s);
    switch (roles(bulk_roles)) {
        case 3: return launch_spec_bulk<S3, T>(hc, a, n, sms, s);
        case 2: return launch_spec_bulk<S2, T>(hc, a, n, sms, s);
        default: return launch_spec_bulk<S1, T>(hc, a, n, sms, s);
    }
}
int l64(const SpecHostConsts &hc, const SolveArgs<double> &a, int v, long long n, int sms, cudaStream_t s) {
    return launch<double>(hc, a, v, n, sms, s, 3);
}
int l32(const SpecHostConsts &hc, const SolveArgs<float> &a, int v, long long n, int sms, cudaStream_t s) {
    return launch<float>(hc, a, v, n, sms, s, 1);
}
}  // namespace
extern const SpecializedKernel kSpecCassieFeetPelvis = {S3::name(), spec_matches<S3>, l64, l32};
}  // namespace ikb
