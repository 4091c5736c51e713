// Registry of topology-specialised solve kernels (the fast path).  ikb_problem_finalize() asks
// find_specialized() whether the finalized problem matches one of the compiled specialisations; if not, the
// generic table-driven kernel (dls_generic.cuh) is used.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "dev_problem.hpp"
#include "model.hpp"

namespace ikb {

// Host copies of the run-time constants a specialised kernel takes as a by-value parameter (spec_common.hpp).
struct SpecHostConsts {
    const double *lower, *upper;  // [nq]
    const double *weight;         // [rows], stacked order
    const double *mask;           // [rows]: posture masks, 1 elsewhere
};

// BULK: throughput configuration (persistent, one CTA per SM).  TAIL: latency configuration (one 32-problem group per
// CTA, more warp roles) -- used for the stragglers a BULK launch suspends and for batches too small to fill the GPU.
enum SpecVariant { SPEC_BULK = 0, SPEC_TAIL = 1 };

struct SpecializedKernel {
    const char *name;
    bool (*matches)(const HostProblem &hp);
    // n = upper bound on the number of problems the launch will see (sizes the grid)
    int (*launch64)(const SpecHostConsts &hc, const SolveArgs<double> &a, int variant, long long n, int sm_count, cudaStream_t s);
    int (*launch32)(const SpecHostConsts &hc, const SolveArgs<float> &a, int variant, long long n, int sm_count, cudaStream_t s);
    // same topology / task list, other placement values (dls_spec.cuh spec_near_miss): `why` explains
    bool (*near_miss)(const HostProblem &hp, std::string *why);
};
// Load a specialisation compiled after the library was built (a shared object exporting
// `extern "C" const ikb::SpecializedKernel *ikb_spec_plugin()`, made by ik_b200/specialise.py).  0 on success.
int load_specialisation_plugin(const char *path, std::string *err);
// the first compiled specialisation that is a near miss for `hp` (nullptr: none)
const SpecializedKernel *find_near_miss(const HostProblem &hp, std::string *why);

const SpecializedKernel *find_specialized(const HostProblem &hp);
// all compiled specialisations (NULL-terminated), for introspection / tests
const SpecializedKernel *const *specialized_registry();

template <typename T>
inline int launch_specialized(const SpecializedKernel &k, const SpecHostConsts &hc, const SolveArgs<T> &a, int variant,
                              long long n, int sm_count, cudaStream_t s);
template <>
inline int launch_specialized<double>(const SpecializedKernel &k, const SpecHostConsts &hc, const SolveArgs<double> &a,
                                      int variant, long long n, int sm_count, cudaStream_t s) {
    return k.launch64(hc, a, variant, n, sm_count, s);
}
template <>
inline int launch_specialized<float>(const SpecializedKernel &k, const SpecHostConsts &hc, const SolveArgs<float> &a,
                                     int variant, long long n, int sm_count, cudaStream_t s) {
    return k.launch32(hc, a, variant, n, sm_count, s);
}

}  // namespace ikb
