// Specialised solve kernel: fixed-base 7-DoF manipulator with one Full frame task on `tool` (BASELINE.json config 5).
#include "dls_spec.cuh"
#include "gen/manipulator_tool.cuh"

namespace ikb {
namespace {
using S = SpecManipulatorTool;
int l64(const SpecHostConsts &hc, const SolveArgs<double> &a, int sms, cudaStream_t s) { return launch_spec<S, double>(hc, a, sms, s); }
int l32(const SpecHostConsts &hc, const SolveArgs<float> &a, int sms, cudaStream_t s) { return launch_spec<S, float>(hc, a, sms, s); }
}  // namespace
extern const SpecializedKernel kSpecManipulatorTool = {S::name(), spec_matches<S>, l64, l32};
}  // namespace ikb
