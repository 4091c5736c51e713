// Device-side constant blob of one IK problem (model + tasks), one instance per scalar type.
//
// This is the "URDF constants" block of the design: ~10 KB, read-only, identical for every problem of a
// batch.  It is written once into HBM by ikb_problem_finalize() and every CTA stages it into shared memory
// with one TMA bulk copy (cp.async.bulk) at kernel start; lanes then read it through shared-memory
// broadcasts.  Tasks are stored in STACKED order (priority level, then insertion order; reference
// dls.cpp:18-24) so the kernel never sorts.
#pragma once
#include <cstdint>

namespace ikb {

constexpr int kMaxJoints = 48;
constexpr int kMaxNq = 64;
constexpr int kMaxTasks = 16;
constexpr int kMaxFrames = 32;  // frames referenced by tasks (task frames + reference frames)
constexpr int kMaxRows = 48;

template <typename T> struct alignas(16) DevProblem {
    int32_t njoints, nq, nv, nframes, ntasks, rows, rows_p0, tsz;
    int32_t parent[kMaxJoints], jtype[kMaxJoints], idx_q[kMaxJoints], idx_v[kMaxJoints];
    T placement[kMaxJoints][12];
    T axis[kMaxJoints][3];
    T lower[kMaxNq], upper[kMaxNq];
    int32_t f_parent[kMaxFrames];
    T f_placement[kMaxFrames][12];
    // tasks, stacked order
    int32_t t_kind[kMaxTasks], t_frame[kMaxTasks], t_ref[kMaxTasks], t_type[kMaxTasks];
    int32_t t_row[kMaxTasks], t_dim[kMaxTasks], t_toff[kMaxTasks], t_moff[kMaxTasks];
    T weight[kMaxRows];  // stacked row order
    T mask[kMaxNq];      // posture masks, concatenated in stacked order
    int32_t pad_[4];     // keeps sizeof a multiple of 16 for the bulk copy
};

// Per-launch arguments (passed by value as a kernel parameter).
template <typename T> struct SolveArgs {
    const T *q0; long long q0_es, q0_bs;
    const T *targets; long long tg_es, tg_bs;
    T *q; long long q_es, q_bs;
    unsigned char *success;
    int *iters;
    T *resid;
    unsigned long long *ticket;  // work-stealing counter, zeroed before the launch
    long long B;
    int max_iterations;
    T step_length, damping2, tolerance;
    // Two-phase scheduling of the specialised kernels (dls_spec.cuh): the BULK launch suspends a problem that has not
    // finished after `it_cap` steps (its iterate goes to `q`, its step count to `iters_ws`, its index to `list`); the
    // TAIL launch (`resume` = 1) takes its problems from `list` and continues them from `q` / `iters_ws`.
    int it_cap;                       // >= max_iterations: never suspend
    int resume;                       // 0: tickets are problem indices 0..B-1; 1: tickets index `list`
    unsigned int *list;               // suspended problem indices
    unsigned long long *list_count;   // number of entries in `list` (zeroed before the BULK launch)
    int *iters_ws;                    // step counts of suspended problems (the caller's `iters` or scratch)
};

}  // namespace ikb
