// Specialised solve kernel: fixed-base 7-DoF manipulator with one Full frame task on `tool` (BASELINE.json config 5).
#include "dls_spec.cuh"
#include "gen/manipulator_tool.cuh"

namespace ikb {
namespace {
using S = SpecManipulatorTool;
template <typename T> int launch(const SpecHostConsts &hc, const SolveArgs<T> &a, int variant, long long n, int sms, cudaStream_t s) {
    return variant == SPEC_TAIL ? launch_spec_tail<S, T>(hc, a, n, sms, s) : launch_spec_bulk<S, T>(hc, a, n, sms, s);
}
int l64(const SpecHostConsts &hc, const SolveArgs<double> &a, int v, long long n, int sms, cudaStream_t s) { return launch<double>(hc, a, v, n, sms, s); }
int l32(const SpecHostConsts &hc, const SolveArgs<float> &a, int v, long long n, int sms, cudaStream_t s) { return launch<float>(hc, a, v, n, sms, s); }
}  // namespace
extern const SpecializedKernel kSpecManipulatorTool = {S::name(), spec_matches<S>, l64, l32, spec_near_miss<S>};
}  // namespace ikb
