// Topology-specialised batched DLS IK kernel: the fast path of ikb_dls_solve_batch.
//
// One IK problem per THREAD, the whole ik::dls loop (reference dls.cpp:14-74) in-kernel, exactly like the generic
// kernel -- but the per-iteration body is straight-line code generated for one fixed (tree, task list) pair by
// tools/gen_kernel.py (see its header for the arithmetic), so:
//   * there are no tables, no dynamic indexing and no local memory: FK, task errors and the Gram / LDL^T block
//     columns live in registers;
//   * the non-zero entries of the weighted task Jacobian, the LDL^T factor and the problem's target poses live in
//     thread-private strips of shared memory (element k of thread t at base[k * BLOCK + t]: conflict-free);
//   * limits and task weights arrive as a __grid_constant__ parameter, i.e. as constant-bank operands.
// Per-thread state for the Cassie feet+pelvis problem in FP64: 105 (J) + 78 (factor) + 36 (targets) doubles of shared
// memory = 1752 B, so 128 threads fill the 227 KB an SM offers; the register file holds the rest (255 regs/thread).
//
// Scheduling: persistent CTAs (one per SM for FP64); every LANE pulls problem indices from a global ticket counter
// and refills itself the moment its problem converges or runs out of iterations, so all lanes of a warp execute the
// same evaluate -> solve -> integrate body on every trip although iteration counts differ wildly across problems
// (median 4, p95 13, max 100 on the Cassie workload).
#pragma once
#include <cuda_runtime.h>

#include "dev_problem.hpp"
#include "model.hpp"
#include "spec_common.hpp"
#include "specialized.hpp"

namespace ikb {

constexpr int kSmemPerSm = 228 * 1024;   // B200: 228 KB per SM, 1 KB reserved per resident CTA
constexpr int kSmemPerCtaMax = 227 * 1024;

template <class Spec, typename T> struct SpecLaunch {
    static constexpr int kStrip = Spec::NSLOT + Spec::NFACT + Spec::TSZ;  // scalars per thread
    static constexpr int kBytesPerThread = kStrip * (int)sizeof(T);
    static constexpr int kFit = kSmemPerCtaMax / kBytesPerThread;  // threads one CTA could hold
    static constexpr int BLOCK = kFit >= 128 ? 128 : (kFit / 32) * 32;
    static constexpr int kSmemBytes = BLOCK > 0 ? BLOCK * kBytesPerThread : 0;
    // resident CTAs per SM: shared-memory bound, capped at 2 (256 threads x 255 registers is the whole register file)
    static constexpr int kBySmem = BLOCK > 0 ? kSmemPerSm / (kSmemBytes + 1024) : 0;
    static constexpr int MINB = kBySmem >= 2 ? 2 : 1;
    static constexpr bool kFits = BLOCK >= 32;
};

template <class Spec, typename T, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
    dls_spec_kernel(const __grid_constant__ SpecConsts<T, Spec::NQ, Spec::M> c, const __grid_constant__ SolveArgs<T> a) {
    constexpr int NQ = Spec::NQ, NV = Spec::NV, M = Spec::M, M0 = Spec::M0, TSZ = Spec::TSZ;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw) + threadIdx.x;
    const Strip<T, BLOCK> sJ{sm};                                          // weighted task Jacobian, non-zeros only
    const Strip<T, BLOCK> sL{sm + Spec::NSLOT * BLOCK};                    // LDL^T factor
    const Strip<T, BLOCK> sT{sm + (Spec::NSLOT + Spec::NFACT) * BLOCK};    // this problem's target poses

    T q[NQ];
    long long b;
    int it = 0;
    bool have;

    auto fetch = [&]() {
        b = (long long)atomicAdd(a.ticket, 1ULL);
        have = b < a.B;
        it = 0;
        if (have) {
            const T *q0 = a.q0 + b * a.q0_bs;
#pragma unroll
            for (int k = 0; k < NQ; ++k) q[k] = __ldg(q0 + k * a.q0_es);
            const T *tg = a.targets + b * a.tg_bs;
#pragma unroll 4
            for (int k = 0; k < TSZ; ++k) sT.set(k, __ldg(tg + k * a.tg_es));
        }
    };
    fetch();

    while (__any_sync(0xffffffffu, have)) {
        if (have) {
            T e[M];
            Spec::evaluate(q, sT, c, sJ, e);                    // data.cpp:25-58
            T res = T(0);
#pragma unroll
            for (int i = 0; i < M0; ++i) res += e[i] * e[i];    // visitor.hpp:19 (priority-0 rows)
            const bool converged = res < a.tolerance;
            bool finished = converged;
            if (!converged) {
                // dq is not part of the batch result, so the solve of a converged problem (dls.cpp:52-53 runs before
                // the stop test) is skipped: the returned q is the un-stepped iterate either way (dls.cpp:62-63).
                T y[M], dq[NV];
                Spec::solve(sJ, sL, a.damping2, e, y);          // dls.cpp:39-41,53
                Spec::step_direction(sJ, y, dq);                // dls.cpp:52
                Spec::integrate(q, dq, a.step_length, c);       // dls.cpp:67-71
                ++it;
                finished = it >= a.max_iterations;              // dls.cpp:14,76-77
            }
            if (finished) {
                T *qo = a.q + b * a.q_bs;
#pragma unroll
                for (int k = 0; k < NQ; ++k) qo[k * a.q_es] = q[k];
                if (a.success) a.success[b] = converged ? 1 : 0;
                if (a.iters) a.iters[b] = it;
                if (a.resid) a.resid[b] = res;
                fetch();
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------

inline bool same_values(const double *a, const double *b, int n) {  // value equality (-0.0 == 0.0), not bit equality
    for (int i = 0; i < n; ++i)
        if (!(a[i] == b[i])) return false;
    return true;
}

// Does the finalized problem equal the (tree, task list) this specialisation was generated for?
template <class Spec> bool spec_matches(const HostProblem &hp) {
    const HostModel &m = hp.model;
    if (m.njoints() != Spec::NJOINTS || m.nq != Spec::NQ || m.nv != Spec::NV) return false;
    if ((int)hp.tasks.size() != Spec::NTASKS || hp.rows() != Spec::M || hp.e_size(0) != Spec::M0) return false;
    const int *par = Spec::sig_parent(), *typ = Spec::sig_type();
    const double *pl = Spec::sig_placement();
    for (int j = 0; j < Spec::NJOINTS; ++j) {
        if (m.parent[j] != par[j] || m.jtype[j] != typ[j]) return false;
        if (!same_values(m.placement[j].data(), pl + 15 * j, 12)) return false;
        if (typ[j] == IKB_J_REV_UNALIGNED || typ[j] == IKB_J_PRIS_UNALIGNED)
            if (!same_values(m.axis[j].data(), pl + 15 * j + 12, 3)) return false;
    }
    // tasks must already be in stacked order (the generated code has no notion of insertion order)
    const SE3d ident = se3_identity();
    for (int t = 0; t < Spec::NTASKS; ++t) {
        const HostTask &ht = hp.tasks[t];
        if (ht.kind != IKB_TASK_FRAME || ht.type != Spec::sig_task_type()[t] || ht.priority != Spec::sig_task_priority()[t])
            return false;
        if (t > 0 && ht.priority < hp.tasks[t - 1].priority) return false;
        // reference frame must coincide with `universe`
        if (m.frame_parent[ht.ref] != 0 || m.frame_placement[ht.ref] != ident) return false;
        if (m.frame_parent[ht.frame] != Spec::sig_task_joint()[t]) return false;
        if (!same_values(m.frame_placement[ht.frame].data(), Spec::sig_task_placement() + 12 * t, 12)) return false;
    }
    return true;
}

template <class Spec, typename T> int launch_spec(const SpecHostConsts &hc, const SolveArgs<T> &a, int sm_count, cudaStream_t s) {
    using L = SpecLaunch<Spec, T>;
    static_assert(L::kFits, "per-thread strips do not fit in shared memory for this scalar type");
    auto fn = dls_spec_kernel<Spec, T, L::BLOCK, L::MINB>;
    static bool attr_set = false;  // per (Spec, T) instantiation
    if (!attr_set) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmemBytes) != cudaSuccess) return 1;
        attr_set = true;
    }
    SpecConsts<T, Spec::NQ, Spec::M> c;
    for (int k = 0; k < Spec::NQ; ++k) {
        c.lower[k] = (T)hc.lower[k];
        c.upper[k] = (T)hc.upper[k];
    }
    for (int i = 0; i < Spec::M; ++i) c.weight[i] = (T)hc.weight[i];
    long long blocks = (a.B + L::BLOCK - 1) / L::BLOCK;
    const long long resident = (long long)L::MINB * sm_count;
    if (blocks > resident) blocks = resident;
    fn<<<(unsigned)blocks, L::BLOCK, L::kSmemBytes, s>>>(c, a);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace ikb
