// Launcher of the table-driven team-per-problem kernel (dls_coop.cuh): every problem without a compiled specialisation,
// every problem with FrameConstraints or a CentreOfMassTask, and every ik::pik solve.
#include "capi_internal.hpp"
#include "dls_coop.cuh"
#include "dls_spec.cuh"  // DynSmemOptIn

using namespace ikb;
using namespace ikb::capi;

namespace {

template <typename T, int CLS, bool SHFL, bool PIK, int EXTRA>
int launch_one(const DevProblem<T> *dP, const SolveArgs<T> &a, int sm_count, cudaStream_t s) {
    using Cfg = typename CoopClass<CLS>::Cfg;
    using L = CoopLaunch<T, Cfg, EXTRA>;
    auto fn = dls_coop_kernel<T, Cfg, SHFL, PIK, EXTRA>;
    static DynSmemOptIn opt_in;
    if (!opt_in.ensure(fn, (int)L::kSmem)) return 1;
    long long ctas = a.B < sm_count ? a.B : sm_count;   // persistent: one CTA per SM; a small batch spreads one team per SM
    if (ctas < 1) ctas = 1;
    fn<<<(unsigned)ctas, L::kThreads, L::kSmem, s>>>(dP, a);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

template <typename T, int CLS, bool SHFL>
int launch_cls(const DevProblem<T> *dP, const SolveArgs<T> &a, bool pik, int extra, int sm_count, cudaStream_t s) {
    if (pik) return launch_one<T, CLS, SHFL, true, 2>(dP, a, sm_count, s);
    if (extra >= 2) return launch_one<T, CLS, SHFL, false, 2>(dP, a, sm_count, s);
    if (extra == 1) return launch_one<T, CLS, SHFL, false, 1>(dP, a, sm_count, s);
    return launch_one<T, CLS, SHFL, false, 0>(dP, a, sm_count, s);
}

template <typename T, bool SHFL>
int launch_shfl(int cls, const DevProblem<T> *dP, const SolveArgs<T> &a, bool pik, int extra, int sm_count, cudaStream_t s) {
    switch (cls) {
        case 0: return launch_cls<T, 0, SHFL>(dP, a, pik, extra, sm_count, s);
        case 1: return launch_cls<T, 1, SHFL>(dP, a, pik, extra, sm_count, s);
        case 3: return launch_cls<T, 3, SHFL>(dP, a, pik, extra, sm_count, s);
        default: return launch_cls<T, 2, SHFL>(dP, a, pik, extra, sm_count, s);
    }
}
}  // namespace

namespace ikb {
namespace capi {

template <typename T>
int launch_coop(int cls, const DevProblem<T> *dP, const SolveArgs<T> &a, bool pik, int extra, bool shfl, int sm_count, cudaStream_t s) {
    return shfl ? launch_shfl<T, true>(cls, dP, a, pik, extra, sm_count, s)
                : launch_shfl<T, false>(cls, dP, a, pik, extra, sm_count, s);
}
template int launch_coop<double>(int, const DevProblem<double> *, const SolveArgs<double> &, bool, int, bool, int, cudaStream_t);
template int launch_coop<float>(int, const DevProblem<float> *, const SolveArgs<float> &, bool, int, bool, int, cudaStream_t);

}  // namespace capi
}  // namespace ikb
