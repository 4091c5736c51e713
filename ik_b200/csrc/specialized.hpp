// Registry of topology-specialised solve kernels (the fast path).  ikb_problem_finalize() asks
// find_specialized() whether the finalized problem matches one of the compiled specialisations; if not, the
// generic table-driven kernel (dls_generic.cuh) is used.
#pragma once
#include <cuda_runtime.h>

#include "dev_problem.hpp"
#include "model.hpp"

namespace ikb {

struct SpecializedKernel {
    const char *name;
    bool (*matches)(const HostProblem &hp);
    int (*launch64)(const DevProblem<double> *P, const SolveArgs<double> &a, int sm_count, cudaStream_t s);
    int (*launch32)(const DevProblem<float> *P, const SolveArgs<float> &a, int sm_count, cudaStream_t s);
};

const SpecializedKernel *find_specialized(const HostProblem &hp);

template <typename T>
inline int launch_specialized(const SpecializedKernel &k, const DevProblem<T> *P, const SolveArgs<T> &a, int sm_count,
                              cudaStream_t s);
template <>
inline int launch_specialized<double>(const SpecializedKernel &k, const DevProblem<double> *P, const SolveArgs<double> &a,
                                      int sm_count, cudaStream_t s) {
    return k.launch64(P, a, sm_count, s);
}
template <>
inline int launch_specialized<float>(const SpecializedKernel &k, const DevProblem<float> *P, const SolveArgs<float> &a,
                                     int sm_count, cudaStream_t s) {
    return k.launch32(P, a, sm_count, s);
}

}  // namespace ikb
