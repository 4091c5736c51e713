// C ABI of libikb200.so (see include/ikb200.h for the contract and the reference interfaces replaced).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <type_traits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ikb200.h"
#include "dev_problem.hpp"
#include "dls_generic.cuh"
#include "model.hpp"
#include "specialized.hpp"

using namespace ikb;

// ---------------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------------
struct ikb_model {
    HostModel m;
};

namespace {
constexpr int kTicketSlots = 64;

constexpr int kScratchSlots = 8;

// Device scratch of one in-flight two-phase solve: the suspended-problem list and (when the caller passes no `iters`)
// the step counts the tail launch resumes from.  Slots rotate; `ev` marks the end of the slot's last user.
struct SolveScratch {
    unsigned int *list = nullptr;
    int *iters = nullptr;
    size_t cap = 0;
    cudaEvent_t ev = nullptr;
};

template <typename T> struct Staging {
    T *q0 = nullptr, *targets = nullptr, *q = nullptr, *resid = nullptr;
    size_t q0_cap = 0, tg_cap = 0, q_cap = 0, b_cap = 0;
};
}  // namespace

struct ikb_problem {
    HostProblem hp;
    bool finalized = false;
    int device = -1;
    int size_class = -1;
    int sm_count = 0;
    DevProblem<double> *d64 = nullptr;
    DevProblem<float> *d32 = nullptr;
    int *d_frame_parent = nullptr;  // all model frames (for ikb_fk_batch)
    double *d_frame_pl64 = nullptr;
    float *d_frame_pl32 = nullptr;
    unsigned long long *d_tickets = nullptr;
    std::atomic<unsigned> ticket_next{0};
    SolveScratch scratch[kScratchSlots];
    std::mutex scratch_mu;
    const SpecializedKernel *spec = nullptr;
    std::vector<double> weight_stacked;  // Task::weighting() rows in stacked order (constants of the specialised kernels)
    std::string kernel_name[2];
    // host-path staging (ikb_dls_solve_batch_host): main stream + the pipelined path's copy-in and second compute stream
    cudaStream_t stream = nullptr, stream_in = nullptr, stream_aux = nullptr;
    cudaEvent_t ev_in[8] = {}, ev_aux = nullptr, ev_main = nullptr;
    Staging<double> st64;
    Staging<float> st32;
    unsigned char *st_success = nullptr;
    int *st_iters = nullptr;
    size_t st_flag_cap = 0;
};

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? IKB_ERR_NO_DEVICE : IKB_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}
#define IKB_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);   \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct SizeClass {
    int nj, nv, m;
};
const SizeClass kClasses[] = {{10, 8, 6}, {20, 24, 12}, {32, 36, 30}};

template <typename T>
void fill_dev_problem(const HostProblem &hp, const std::vector<int> &order, const std::vector<int> &used_frames,
                      DevProblem<T> &P) {
    const HostModel &m = hp.model;
    std::memset(&P, 0, sizeof(P));
    P.njoints = m.njoints();
    P.nq = m.nq;
    P.nv = m.nv;
    P.nframes = (int)used_frames.size();
    P.ntasks = (int)hp.tasks.size();
    P.rows = hp.rows();
    P.rows_p0 = hp.e_size(0);
    P.tsz = hp.target_size();
    for (int j = 0; j < m.njoints(); ++j) {
        P.parent[j] = m.parent[j];
        P.jtype[j] = m.jtype[j];
        P.idx_q[j] = m.idx_q[j];
        P.idx_v[j] = m.idx_v[j];
        for (int k = 0; k < 12; ++k) P.placement[j][k] = (T)m.placement[j][k];
        for (int k = 0; k < 3; ++k) P.axis[j][k] = (T)m.axis[j][k];
    }
    const double big = (double)std::numeric_limits<T>::max();
    for (int k = 0; k < m.nq; ++k) {
        P.lower[k] = (T)std::max(m.lower[k], -big);
        P.upper[k] = (T)std::min(m.upper[k], big);
    }
    for (size_t f = 0; f < used_frames.size(); ++f) {
        P.f_parent[f] = m.frame_parent[used_frames[f]];
        for (int k = 0; k < 12; ++k) P.f_placement[f][k] = (T)m.frame_placement[used_frames[f]][k];
    }
    auto local_frame = [&](int fid) {
        return (int)(std::find(used_frames.begin(), used_frames.end(), fid) - used_frames.begin());
    };
    int row = 0, moff = 0;
    for (size_t s = 0; s < order.size(); ++s) {
        const HostTask &t = hp.tasks[order[s]];
        P.t_kind[s] = t.kind;
        P.t_frame[s] = t.kind == IKB_TASK_POSTURE ? 0 : local_frame(t.frame);
        P.t_ref[s] = t.kind == IKB_TASK_POSTURE ? 0 : local_frame(t.ref);
        P.t_type[s] = t.type;
        P.t_row[s] = row;
        P.t_dim[s] = t.dim;
        P.t_toff[s] = hp.target_offset(order[s]);
        P.t_moff[s] = moff;
        for (int i = 0; i < t.dim; ++i) P.weight[row + i] = (T)t.weight[i];
        if (t.kind == IKB_TASK_POSTURE) {
            for (int i = 0; i < t.type; ++i) P.mask[moff + i] = (T)t.mask[i];
            moff += t.type;
        }
        row += t.dim;
    }
}

template <typename T> struct KernelTable {
    using Fn = void (*)(const DevProblem<T> *, SolveArgs<T>);
    static Fn dls(int cls) {
        switch (cls) {
            case 0: return dls_generic_kernel<T, 10, 8, 6>;
            case 1: return dls_generic_kernel<T, 20, 24, 12>;
            default: return dls_generic_kernel<T, 32, 36, 30>;
        }
    }
};

// max_iterations <= 0: the reference returns q0 untouched, success = false, nothing evaluated
template <typename T> __global__ void passthrough_kernel(SolveArgs<T> a, int nq) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    for (int k = 0; k < nq; ++k) a.q[k * a.q_es + b * a.q_bs] = a.q0[k * a.q0_es + b * a.q0_bs];
    if (a.success) a.success[b] = 0;
    if (a.iters) a.iters[b] = 0;
    if (a.resid) a.resid[b] = T(0);
}

template <typename T> DevProblem<T> *dev_blob(const ikb_problem *p);
template <> DevProblem<double> *dev_blob<double>(const ikb_problem *p) { return p->d64; }
template <> DevProblem<float> *dev_blob<float>(const ikb_problem *p) { return p->d32; }

// Pipelined host path (solve_host): the inputs of batch slice [begin[c], begin[c + 1]) are on the device once `ready[c]`
// has happened.  The BULK launch is issued per slice, alternating between the caller's stream and `aux`, so that it
// overlaps the host-to-device copy of the next slices; the TAIL launch continues the stragglers of all slices at once.
struct ChunkPlan {
    int n = 0;
    long long begin[9] = {};
    cudaEvent_t ready[8] = {};
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_aux = nullptr, ev_main = nullptr;
};
__global__ void set_ticket_kernel(unsigned long long *t, unsigned long long v) { *t = v; }

// Is this solve going to take the two-launch (BULK + TAIL) path?  (the only one that can be pipelined by slices)
bool two_phase(const ikb_problem *p, const ikb_dls_params *prm, int64_t B, int *cap_out = nullptr) {
    const char *cap_env = std::getenv("IKB_BULK_CAP");
    const int cap = cap_env ? std::atoi(cap_env) : 16;
    if (cap_out) *cap_out = cap;
    return p->spec && B > 2LL * 32 * p->sm_count && cap > 0 && prm->max_iterations > cap;
}

// Merged launch of the pipelined queue (ikb_queue_*): `nseg` batches described by a device-resident table.
template <typename T> struct Merged {
    const BatchSeg<T> *seg;  // host array, `nseg` entries sorted by begin
    int nseg;
};

template <typename T>
int launch_solve(const ikb_problem *p, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io, cudaStream_t s,
                 const ChunkPlan *plan = nullptr, const Merged<T> *merged = nullptr) {
    SolveArgs<T> a{};
    if (!merged) {
        a.q0 = (const T *)io->q0; a.q0_es = io->q0_elem_stride; a.q0_bs = io->q0_batch_stride;
        a.targets = (const T *)io->targets; a.tg_es = io->targets_elem_stride; a.tg_bs = io->targets_batch_stride;
        a.q = (T *)io->q; a.q_es = io->q_elem_stride; a.q_bs = io->q_batch_stride;
        a.success = io->success;
        a.iters = io->iters;
        a.resid = (T *)io->resid;
        a.nseg = 0;
    } else {
        if (!p->spec || prm->max_iterations <= 0) return fail(IKB_ERR_INVALID_ARG, "internal: merged launch on a problem without a specialised kernel");
        a.nseg = merged->nseg;
        for (int i = 0; i < kMaxSegments; ++i) {
            if (i < merged->nseg) a.seg[i] = merged->seg[i];
            else a.seg[i].begin = LLONG_MAX;
        }
    }
    a.B = B;
    a.max_iterations = prm->max_iterations;
    a.step_length = (T)prm->step_length;
    a.damping2 = (T)(prm->damping * prm->damping);
    a.tolerance = (T)prm->tolerance;
    const unsigned slot = const_cast<ikb_problem *>(p)->ticket_next.fetch_add(1) % kTicketSlots;
    a.ticket = p->d_tickets + slot * 16;  // 128 B apart
    IKB_CUDA(cudaMemsetAsync(a.ticket, 0, sizeof(unsigned long long), s));

    if (prm->max_iterations <= 0) {
        // dls.cpp:14 never enters the loop: q0 is returned with success = false (dls.cpp:76-77)
        const int threads = 128;
        passthrough_kernel<T><<<(unsigned)((B + threads - 1) / threads), threads, 0, s>>>(a, p->hp.model.nq);
        IKB_CUDA(cudaGetLastError());
        g_launches.fetch_add(1);
        return IKB_OK;
    }
    a.it_cap = INT_MAX;
    a.resume = 0;
    a.list = nullptr;
    a.list_count = nullptr;
    a.iters_ws = merged ? nullptr : io->iters;
    if (p->spec) {
        const SpecHostConsts hc{p->hp.model.lower.data(), p->hp.model.upper.data(), p->weight_stacked.data()};
        // Scheduling.  A batch that the latency configuration keeps resident in one wave runs there
        // directly.  A larger batch runs BULK (throughput configuration) with a step cap: the few problems still
        // unfinished after `cap` steps -- the reference lets them run to max_iterations, 100 by default -- are suspended
        // and a TAIL launch continues all of them at once, each group of 32 with an SM's schedulers to itself, instead
        // of letting them trickle out of the bulk kernel one 100-step straggler at a time.
        const long long wave = 2LL * 32 * p->sm_count;
        int cap;
        const bool two = two_phase(p, prm, B, &cap);
        int rc;
        if (plan && !two) return fail(IKB_ERR_INVALID_ARG, "internal: slice plan on a single-launch solve");
        if (B <= wave) {
            rc = launch_specialized<T>(*p->spec, hc, a, SPEC_TAIL, B, p->sm_count, s);
            if (rc == IKB_OK) g_launches.fetch_add(1);
        } else if (cap <= 0 || prm->max_iterations <= cap) {
            rc = launch_specialized<T>(*p->spec, hc, a, SPEC_BULK, B, p->sm_count, s);
            if (rc == IKB_OK) g_launches.fetch_add(1);
        } else {
            ikb_problem *mp = const_cast<ikb_problem *>(p);
            SolveScratch *sc;
            {
                std::lock_guard<std::mutex> lk(mp->scratch_mu);
                if (mp->scratch[0].cap < (size_t)B) {
                    // grow every slot at once (one synchronisation, on the first large batch only)
                    IKB_CUDA(cudaDeviceSynchronize());
                    for (auto &x : mp->scratch) {
                        if (x.list) cudaFree(x.list);
                        if (x.iters) cudaFree(x.iters);
                        x.list = nullptr; x.iters = nullptr; x.cap = 0;
                        IKB_CUDA(cudaMalloc(&x.list, (size_t)B * sizeof(unsigned int)));
                        IKB_CUDA(cudaMalloc(&x.iters, (size_t)B * sizeof(int)));
                        x.cap = (size_t)B;
                        if (!x.ev) IKB_CUDA(cudaEventCreateWithFlags(&x.ev, cudaEventDisableTiming));
                    }
                }
                sc = &mp->scratch[slot % kScratchSlots];
                IKB_CUDA(cudaStreamWaitEvent(s, sc->ev, 0));  // the slot's previous user (any stream) must be done
            }
            IKB_CUDA(cudaMemsetAsync(a.ticket, 0, 3 * sizeof(unsigned long long), s));  // bulk ticket, tail ticket, list count
            a.it_cap = cap;
            a.list = sc->list;
            a.list_count = a.ticket + 2;
            a.iters_ws = (!merged && io->iters) ? io->iters : sc->iters;
            if (!plan) {
                rc = launch_specialized<T>(*p->spec, hc, a, SPEC_BULK, B, p->sm_count, s);
                if (rc == IKB_OK) g_launches.fetch_add(1);
            } else {
                // one BULK launch per slice: its tickets run from begin[c] to begin[c + 1] (own counter, words 3.. of the slot)
                rc = IKB_OK;
                IKB_CUDA(cudaEventRecord(plan->ev_main, s));  // counters zeroed
                IKB_CUDA(cudaStreamWaitEvent(plan->aux, plan->ev_main, 0));
                for (int c = 0; c < plan->n && rc == IKB_OK; ++c) {
                    cudaStream_t cs = (c & 1) ? plan->aux : s;
                    IKB_CUDA(cudaStreamWaitEvent(cs, plan->ready[c], 0));
                    SolveArgs<T> ac = a;
                    ac.ticket = a.ticket + 3 + c;
                    ac.B = plan->begin[c + 1];
                    set_ticket_kernel<<<1, 1, 0, cs>>>(ac.ticket, (unsigned long long)plan->begin[c]);
                    rc = launch_specialized<T>(*p->spec, hc, ac, SPEC_BULK, plan->begin[c + 1] - plan->begin[c], p->sm_count, cs);
                    if (rc == IKB_OK) g_launches.fetch_add(2);
                }
                IKB_CUDA(cudaEventRecord(plan->ev_aux, plan->aux));
                IKB_CUDA(cudaStreamWaitEvent(s, plan->ev_aux, 0));
            }
            if (rc == IKB_OK) {
                SolveArgs<T> t = a;
                t.resume = 1;
                t.it_cap = INT_MAX;
                t.ticket = a.ticket + 1;
                rc = launch_specialized<T>(*p->spec, hc, t, SPEC_TAIL, B, p->sm_count, s);
                if (rc == IKB_OK) g_launches.fetch_add(1);
            }
            IKB_CUDA(cudaEventRecord(sc->ev, s));
        }
        if (rc != IKB_OK) return cuda_fail(cudaGetLastError(), "specialised kernel launch");
        return IKB_OK;
    }
    auto fn = KernelTable<T>::dls(p->size_class);
    const int threads = 128;
    const size_t smem = sizeof(DevProblem<T>) + 16;
    int per_sm = 0;
    IKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (B + threads - 1) / threads;
    blocks = std::min<long long>(blocks, (long long)per_sm * p->sm_count);
    fn<<<(unsigned)blocks, threads, smem, s>>>(dev_blob<T>(p), a);
    IKB_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return IKB_OK;
}

size_t view_extent(int64_t n_elem, int64_t es, int64_t bs, int64_t B) {
    return (size_t)((n_elem - 1) * es + (B - 1) * bs + 1);
}

template <typename T> int ensure(T *&ptr, size_t &cap, size_t need) {
    if (need <= cap) return IKB_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = std::max(need, (size_t)1024);
    IKB_CUDA(cudaMalloc(&ptr, want * sizeof(T)));
    cap = want;
    return IKB_OK;
}

template <typename T> Staging<T> &staging(ikb_problem *p);
template <> Staging<double> &staging<double>(ikb_problem *p) { return p->st64; }
template <> Staging<float> &staging<float>(ikb_problem *p) { return p->st32; }

// A strided [n_elem][B] view of a host array (include/ikb200.h: element k of problem b at base[k * es + b * bs]).
struct View {
    const void *base;
    long long es, bs;
    int n_elem;
    // can batch slices be copied on their own?  SoA rows (bs == 1), dense AoS (es == 1, bs == n_elem), broadcast (bs == 0)
    bool sliceable(long long B) const {
        if (n_elem <= 0 || bs == 0) return true;
        if (bs == 1) return es >= B;
        return es == 1 && bs == n_elem;
    }
};
// Host-to-device copy of batch slice [b0, b1) of `v` into the staging buffer `dst` (same strides as the view).
template <typename T> int copy_in_slice(T *dst, const View &v, long long B, long long b0, long long b1, bool first, cudaStream_t s) {
    if (v.n_elem <= 0) return IKB_OK;
    const T *src = (const T *)v.base;
    if (v.bs == 0) {
        if (first) IKB_CUDA(cudaMemcpyAsync(dst, src, view_extent(v.n_elem, v.es, 0, 1) * sizeof(T), cudaMemcpyHostToDevice, s));
    } else if (v.bs == 1) {
        IKB_CUDA(cudaMemcpy2DAsync(dst + b0, (size_t)v.es * sizeof(T), src + b0, (size_t)v.es * sizeof(T), (size_t)(b1 - b0) * sizeof(T),
                                   (size_t)v.n_elem, cudaMemcpyHostToDevice, s));
    } else {
        IKB_CUDA(cudaMemcpyAsync(dst + b0 * v.bs, src + b0 * v.bs, (size_t)(b1 - b0) * v.bs * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    return IKB_OK;
}

// IKB_HOST_TRACE=1: print the device-side timeline of one host-path solve (debug aid for the e2e numbers in DESIGN.md)
struct HostTrace {
    bool on = false;
    std::vector<std::pair<const char *, cudaEvent_t>> ev;
    HostTrace() { const char *e = std::getenv("IKB_HOST_TRACE"); on = e && e[0] == '1'; }
    void mark(const char *name, cudaStream_t s) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        ev.emplace_back(name, e);
    }
    void dump() {
        if (!on || ev.empty()) return;
        for (auto &x : ev) {
            float ms = 0;
            cudaEventSynchronize(x.second);
            cudaEventElapsedTime(&ms, ev[0].second, x.second);
            std::fprintf(stderr, "[ikb host trace] %-14s %8.3f ms\n", x.first, ms);
            }
        for (auto &x : ev) cudaEventDestroy(x.second);
        ev.clear();
    }
};

template <typename T>
int solve_host(ikb_problem *p, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io) {
    const int nq = p->hp.model.nq, tsz = p->hp.target_size();
    Staging<T> &st = staging<T>(p);
    const size_t n_q0 = view_extent(nq, io->q0_elem_stride, io->q0_batch_stride, B);
    const size_t n_tg = tsz > 0 ? view_extent(tsz, io->targets_elem_stride, io->targets_batch_stride, B) : 0;
    const size_t n_q = view_extent(nq, io->q_elem_stride, io->q_batch_stride, B);
    int rc;
    if ((rc = ensure(st.q0, st.q0_cap, n_q0)) || (rc = ensure(st.targets, st.tg_cap, std::max<size_t>(n_tg, 1))) ||
        (rc = ensure(st.q, st.q_cap, n_q)) || (rc = ensure(st.resid, st.b_cap, (size_t)B)))
        return rc;
    if ((size_t)B > p->st_flag_cap) {
        if (p->st_success) cudaFree(p->st_success);
        if (p->st_iters) cudaFree(p->st_iters);
        p->st_success = nullptr; p->st_iters = nullptr; p->st_flag_cap = 0;
        IKB_CUDA(cudaMalloc(&p->st_success, (size_t)B));
        IKB_CUDA(cudaMalloc(&p->st_iters, (size_t)B * sizeof(int)));
        p->st_flag_cap = (size_t)B;
    }
    cudaStream_t s = p->stream;
    HostTrace tr;
    tr.mark("start", s);
    ikb_batch_io dio = *io;
    dio.q0 = st.q0;
    dio.targets = st.targets;
    dio.q = st.q;
    dio.success = p->st_success;
    dio.iters = p->st_iters;
    dio.resid = st.resid;
    // A two-launch solve whose input views can be cut into batch slices is pipelined: slice c + 1 crosses PCIe while the
    // BULK launch of slice c runs (the staging buffers keep the caller's strides, so a slice is a 2-D or a dense copy).
    const View vq{io->q0, (long long)io->q0_elem_stride, (long long)io->q0_batch_stride, nq};
    const View vt{io->targets, (long long)io->targets_elem_stride, (long long)io->targets_batch_stride, tsz};
    const char *slices_env = std::getenv("IKB_HOST_SLICES");
    const int nslice = (int)std::min<int64_t>(slices_env ? std::max(1, std::min(8, std::atoi(slices_env))) : 4, B / 8192);
    const char *pipe_env = std::getenv("IKB_HOST_PIPELINE");
    if (two_phase(p, prm, B) && nslice >= 2 && vq.sliceable(B) && vt.sliceable(B) && !(pipe_env && pipe_env[0] == '0')) {
        ChunkPlan plan;
        plan.n = nslice;
        plan.aux = p->stream_aux;
        plan.ev_aux = p->ev_aux;
        plan.ev_main = p->ev_main;
        for (int c = 0; c <= nslice; ++c) plan.begin[c] = c == nslice ? B : (B / nslice * c) / 32 * 32;
        for (int c = 0; c < nslice; ++c) {
            if ((rc = copy_in_slice<T>(st.q0, vq, B, plan.begin[c], plan.begin[c + 1], c == 0, p->stream_in)) ||
                (rc = copy_in_slice<T>(st.targets, vt, B, plan.begin[c], plan.begin[c + 1], c == 0, p->stream_in)))
                return rc;
            plan.ready[c] = p->ev_in[c];
            IKB_CUDA(cudaEventRecord(plan.ready[c], p->stream_in));
            tr.mark("h2d slice", p->stream_in);
        }
        if ((rc = launch_solve<T>(p, prm, B, &dio, s, &plan))) return rc;
    } else {
        IKB_CUDA(cudaMemcpyAsync(st.q0, io->q0, n_q0 * sizeof(T), cudaMemcpyHostToDevice, s));
        if (n_tg) IKB_CUDA(cudaMemcpyAsync(st.targets, io->targets, n_tg * sizeof(T), cudaMemcpyHostToDevice, s));
        tr.mark("h2d", s);
        if ((rc = launch_solve<T>(p, prm, B, &dio, s))) return rc;
    }
    tr.mark("solve", s);
    IKB_CUDA(cudaMemcpyAsync(io->q, st.q, n_q * sizeof(T), cudaMemcpyDeviceToHost, s));
    if (io->success) IKB_CUDA(cudaMemcpyAsync(io->success, p->st_success, (size_t)B, cudaMemcpyDeviceToHost, s));
    if (io->iters) IKB_CUDA(cudaMemcpyAsync(io->iters, p->st_iters, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (io->resid) IKB_CUDA(cudaMemcpyAsync(io->resid, st.resid, (size_t)B * sizeof(T), cudaMemcpyDeviceToHost, s));
    tr.mark("d2h", s);
    IKB_CUDA(cudaStreamSynchronize(s));
    tr.dump();
    return IKB_OK;
}

// ---------------------------------------------------------------------------------------------------
// pipelined queue: batches in flight on three streams (copy-in, compute, copy-out); consecutive batches are MERGED
// into one BULK + TAIL launch pair (the straggler chain of ~0.7 ms is paid once per group instead of once per batch)
// ---------------------------------------------------------------------------------------------------
}  // namespace
struct ikb_queue {
    struct Slot {
        cudaEvent_t ev_in = nullptr, ev_done = nullptr;
        bool busy = false, pending = false, host = false;
        int64_t ticket = -1;  // the batch occupying the slot
        int64_t B = 0;
        ikb_batch_io dio{};   // device view of the batch
        ikb_batch_io hio{};   // host mode: the caller's buffers (copy-out targets)
        // host-mode staging (per scalar type, grown on demand)
        Staging<double> st64;
        Staging<float> st32;
        unsigned char *success = nullptr;
        int *iters = nullptr;
        size_t flag_cap = 0;
    };
    ikb_problem *p = nullptr;
    int depth = 0, merge = 1;
    std::vector<Slot> slots;
    std::vector<int> open;        // slots of the group that has not been launched yet
    ikb_dls_params open_prm{};
    int open_dtype = -1;
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    cudaEvent_t ev_user = nullptr, ev_comp = nullptr;
    int64_t next = 0;
};
namespace {
constexpr int kMaxMerge = kMaxSegments;

template <typename T> Staging<T> &slot_staging(ikb_queue::Slot &sl);
template <> Staging<double> &slot_staging<double>(ikb_queue::Slot &sl) { return sl.st64; }
template <> Staging<float> &slot_staging<float>(ikb_queue::Slot &sl) { return sl.st32; }

bool same_params(const ikb_dls_params &a, const ikb_dls_params &b) {
    return a.max_iterations == b.max_iterations && a.step_length == b.step_length && a.damping == b.damping && a.tolerance == b.tolerance;
}

template <typename T> int queue_copy_out(ikb_queue *q, ikb_queue::Slot &sl) {
    const int nq = q->p->hp.model.nq;
    const ikb_batch_io &io = sl.hio;
    Staging<T> &st = slot_staging<T>(sl);
    const size_t n_q = view_extent(nq, io.q_elem_stride, io.q_batch_stride, sl.B);
    IKB_CUDA(cudaMemcpyAsync(io.q, st.q, n_q * sizeof(T), cudaMemcpyDeviceToHost, q->s_out));
    if (io.success) IKB_CUDA(cudaMemcpyAsync(io.success, sl.success, (size_t)sl.B, cudaMemcpyDeviceToHost, q->s_out));
    if (io.iters) IKB_CUDA(cudaMemcpyAsync(io.iters, sl.iters, (size_t)sl.B * sizeof(int), cudaMemcpyDeviceToHost, q->s_out));
    if (io.resid) IKB_CUDA(cudaMemcpyAsync(io.resid, st.resid, (size_t)sl.B * sizeof(T), cudaMemcpyDeviceToHost, q->s_out));
    return IKB_OK;
}

// Launch the open group: one merged BULK + TAIL pair when the problem has a specialised kernel, else batch by batch.
template <typename T> int queue_flush_t(ikb_queue *q) {
    const int n = (int)q->open.size();
    int rc;
    for (int i : q->open)
        if (q->slots[i].host) IKB_CUDA(cudaStreamWaitEvent(q->s_comp, q->slots[i].ev_in, 0));
    if (n >= 2 && q->p->spec && q->open_prm.max_iterations > 0) {
        BatchSeg<T> tab[kMaxSegments];
        long long total = 0;
        for (int k = 0; k < n; ++k) {
            const ikb_queue::Slot &sl = q->slots[q->open[k]];
            const ikb_batch_io &d = sl.dio;
            tab[k] = BatchSeg<T>{(const T *)d.q0, d.q0_elem_stride, d.q0_batch_stride, (const T *)d.targets, d.targets_elem_stride,
                                   d.targets_batch_stride, (T *)d.q, d.q_elem_stride, d.q_batch_stride, d.success, d.iters,
                                   (T *)d.resid, total};
            total += sl.B;
        }
        const Merged<T> m{tab, n};
        if ((rc = launch_solve<T>(q->p, &q->open_prm, total, nullptr, q->s_comp, nullptr, &m))) return rc;
    } else {
        for (int i : q->open) {
            ikb_queue::Slot &sl = q->slots[i];
            if (sl.B > 0 && (rc = launch_solve<T>(q->p, &q->open_prm, sl.B, &sl.dio, q->s_comp))) return rc;
        }
    }
    bool any_host = false;
    for (int i : q->open) any_host |= q->slots[i].host;
    if (any_host) {
        IKB_CUDA(cudaEventRecord(q->ev_comp, q->s_comp));
        IKB_CUDA(cudaStreamWaitEvent(q->s_out, q->ev_comp, 0));
    }
    for (int i : q->open) {
        ikb_queue::Slot &sl = q->slots[i];
        if (sl.host) {
            if (sl.B > 0 && (rc = queue_copy_out<T>(q, sl))) return rc;
            IKB_CUDA(cudaEventRecord(sl.ev_done, q->s_out));
        } else {
            IKB_CUDA(cudaEventRecord(sl.ev_done, q->s_comp));
        }
        sl.pending = false;
    }
    q->open.clear();
    return IKB_OK;
}
int queue_flush(ikb_queue *q) {
    if (q->open.empty()) return IKB_OK;
    return q->open_dtype == IKB_F64 ? queue_flush_t<double>(q) : queue_flush_t<float>(q);
}

// The slot of the next batch, free of its previous occupant (back-pressure: blocks while that batch is in flight).
int queue_acquire(ikb_queue *q, int dtype, const ikb_dls_params *prm, ikb_queue::Slot **out) {
    int rc;
    ikb_queue::Slot *sl = &q->slots[q->next % q->depth];
    if (sl->pending && (rc = queue_flush(q))) return rc;
    if (sl->busy) {
        IKB_CUDA(cudaEventSynchronize(sl->ev_done));
        sl->busy = false;
    }
    // a group shares one launch: same scalar type, same solver parameters
    if (!q->open.empty() && (q->open_dtype != dtype || !same_params(q->open_prm, *prm)) && (rc = queue_flush(q))) return rc;
    q->open_dtype = dtype;
    q->open_prm = *prm;
    *out = sl;
    return IKB_OK;
}
int64_t queue_commit(ikb_queue *q, ikb_queue::Slot *sl) {
    sl->busy = true;
    sl->pending = true;
    sl->ticket = q->next;
    q->open.push_back((int)(q->next % q->depth));
    if ((int)q->open.size() >= q->merge) {
        int rc = queue_flush(q);
        if (rc) return -rc;
    }
    return q->next++;
}

template <typename T> int queue_stage_host(ikb_queue *q, ikb_queue::Slot &sl, int64_t B, const ikb_batch_io *io) {
    ikb_problem *p = q->p;
    const int nq = p->hp.model.nq, tsz = p->hp.target_size();
    Staging<T> &st = slot_staging<T>(sl);
    const size_t n_q0 = view_extent(nq, io->q0_elem_stride, io->q0_batch_stride, B);
    const size_t n_tg = tsz > 0 ? view_extent(tsz, io->targets_elem_stride, io->targets_batch_stride, B) : 0;
    const size_t n_q = view_extent(nq, io->q_elem_stride, io->q_batch_stride, B);
    int rc;
    if ((rc = ensure(st.q0, st.q0_cap, n_q0)) || (rc = ensure(st.targets, st.tg_cap, std::max<size_t>(n_tg, 1))) ||
        (rc = ensure(st.q, st.q_cap, n_q)) || (rc = ensure(st.resid, st.b_cap, (size_t)B)))
        return rc;
    if ((size_t)B > sl.flag_cap) {
        if (sl.success) cudaFree(sl.success);
        if (sl.iters) cudaFree(sl.iters);
        sl.success = nullptr; sl.iters = nullptr; sl.flag_cap = 0;
        IKB_CUDA(cudaMalloc(&sl.success, (size_t)B));
        IKB_CUDA(cudaMalloc(&sl.iters, (size_t)B * sizeof(int)));
        sl.flag_cap = (size_t)B;
    }
    IKB_CUDA(cudaMemcpyAsync(st.q0, io->q0, n_q0 * sizeof(T), cudaMemcpyHostToDevice, q->s_in));
    if (n_tg) IKB_CUDA(cudaMemcpyAsync(st.targets, io->targets, n_tg * sizeof(T), cudaMemcpyHostToDevice, q->s_in));
    IKB_CUDA(cudaEventRecord(sl.ev_in, q->s_in));
    sl.hio = *io;
    sl.dio = *io;
    sl.dio.q0 = st.q0; sl.dio.targets = st.targets; sl.dio.q = st.q;
    sl.dio.success = sl.success; sl.dio.iters = sl.iters; sl.dio.resid = st.resid;
    return IKB_OK;
}

struct FrameList {
    int n;
    int id[16];
};

template <typename T, int NJ>
__global__ void __launch_bounds__(128) fk_model_frames_kernel(const DevProblem<T> *__restrict__ gP, const T *__restrict__ q,
                                                              long long q_es, long long q_bs, long long B, FrameList fl,
                                                              const int *__restrict__ frame_parent,
                                                              const T *__restrict__ frame_placement, T *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevProblem<T> &P = *reinterpret_cast<DevProblem<T> *>(smem_raw);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(DevProblem<T>));
    stage_blob_tma(&P, gP, (unsigned)sizeof(DevProblem<T>), bar);
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T ql[NJ + 8];
    T oR[NJ][9], op[NJ][3];
    for (int k = 0; k < P.nq; ++k) ql[k] = q[k * q_es + b * q_bs];
    fk_all<T, NJ>(P, ql, oR, op);
    for (int f = 0; f < fl.n; ++f) {
        const int fid = fl.id[f], pj = frame_parent[fid];
        T R[9], p[3];
        se3_mul(oR[pj], op[pj], frame_placement + 12 * fid, frame_placement + 12 * fid + 9, R, p);
#pragma unroll
        for (int i = 0; i < 9; ++i) out[((long long)f * 12 + i) * B + b] = R[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) out[((long long)f * 12 + 9 + i) * B + b] = p[i];
    }
}

template <typename T>
int launch_fk(const ikb_problem *p, int64_t B, const void *q, int64_t es, int64_t bs, const FrameList &fl, void *out,
              cudaStream_t s) {
    const int threads = 128;
    const size_t smem = sizeof(DevProblem<T>) + 16;
    const unsigned blocks = (unsigned)((B + threads - 1) / threads);
    const T *pl = std::is_same<T, double>::value ? (const T *)p->d_frame_pl64 : (const T *)p->d_frame_pl32;
    if (p->size_class <= 1)
        fk_model_frames_kernel<T, 20><<<blocks, threads, smem, s>>>(dev_blob<T>(p), (const T *)q, es, bs, B, fl,
                                                                     p->d_frame_parent, pl, (T *)out);
    else
        fk_model_frames_kernel<T, 32><<<blocks, threads, smem, s>>>(dev_blob<T>(p), (const T *)q, es, bs, B, fl,
                                                                     p->d_frame_parent, pl, (T *)out);
    IKB_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return IKB_OK;
}

// The specialised bodies lay targets out in stacked order, the ABI in insertion order: only problems whose tasks
// were added in non-decreasing priority (stacked order == insertion order) can take the fast path.
const SpecializedKernel *select_specialized(const HostProblem &hp) {
    for (size_t i = 1; i < hp.tasks.size(); ++i)
        if (hp.tasks[i].priority < hp.tasks[i - 1].priority) return nullptr;
    return find_specialized(hp);
}

int check_weights(const double *w, int dim, std::vector<double> &out) {
    out.assign(dim, 1.0);
    if (w) std::copy(w, w + dim, out.begin());
    return IKB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// exported functions
// ---------------------------------------------------------------------------------------------------
extern "C" {

const char *ikb_last_error(void) { return g_err.c_str(); }
int ikb_version(void) { return IKB_VERSION; }
int64_t ikb_kernel_launch_count(void) { return g_launches.load(); }

int ikb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void *ikb_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        g_err = "cudaMallocHost failed";
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void ikb_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

void ikb_dls_params_default(ikb_dls_params *p) {
    if (!p) return;
    p->max_iterations = 100;  // common.hpp:61
    p->random_restart = 0;    // dls.hpp:27
    p->max_time = 1.0;        // common.hpp:63
    p->step_length = 1.0;     // common.hpp:65
    p->damping = 1e-2;        // dls.hpp:25
    p->tolerance = 1e-4;      // visitor.hpp:19
}

int ikb_model_from_urdf(const char *xml, size_t len, int free_flyer, ikb_model **out) {
    if (!xml || !out) return fail(IKB_ERR_INVALID_ARG, "ikb_model_from_urdf: null argument");
    try {
        auto *h = new ikb_model();
        h->m = model_from_urdf(std::string(xml, len), free_flyer != 0);
        *out = h;
        return IKB_OK;
    } catch (const std::exception &e) {
        return fail(IKB_ERR_PARSE, e.what());
    }
}

int ikb_model_from_desc(const ikb_model_desc *d, ikb_model **out) {
    if (!d || !out || d->njoints < 1 || !d->parent || !d->jtype || !d->placement)
        return fail(IKB_ERR_INVALID_ARG, "ikb_model_from_desc: null / empty description");
    if (d->jtype[0] != IKB_J_UNIVERSE) return fail(IKB_ERR_INVALID_ARG, "joint 0 must be the universe");
    auto *h = new ikb_model();
    HostModel &m = h->m;
    int iq = 0;
    for (int j = 0; j < d->njoints; ++j) {
        const int t = d->jtype[j];
        if (t < IKB_J_UNIVERSE || t > IKB_J_PRIS_UNALIGNED || (j > 0 && (d->parent[j] < 0 || d->parent[j] >= j))) {
            delete h;
            return fail(IKB_ERR_INVALID_ARG, "ikb_model_from_desc: bad joint type or parent index (parents must precede children)");
        }
        SE3d pl;
        std::copy(d->placement + 12 * j, d->placement + 12 * j + 12, pl.begin());
        std::array<double, 3> ax{0, 0, 0};
        if (d->axis) ax = {d->axis[3 * j], d->axis[3 * j + 1], d->axis[3 * j + 2]};
        const int n = HostModel::joint_nq(t);
        std::vector<double> lo(n), hi(n);
        for (int k = 0; k < n; ++k) {
            lo[k] = d->lower ? d->lower[iq + k] : -1e300;
            hi[k] = d->upper ? d->upper[iq + k] : 1e300;
        }
        iq += n;
        m.add_joint(d->joint_names && d->joint_names[j] ? d->joint_names[j] : ("joint" + std::to_string(j)), t,
                    j == 0 ? 0 : d->parent[j], pl, ax, lo, hi);
    }
    for (int f = 0; f < d->nframes; ++f) {
        if (d->frame_parent[f] < 0 || d->frame_parent[f] >= d->njoints) {
            delete h;
            return fail(IKB_ERR_INVALID_ARG, "ikb_model_from_desc: frame parent out of range");
        }
        SE3d pl;
        std::copy(d->frame_placement + 12 * f, d->frame_placement + 12 * f + 12, pl.begin());
        m.add_frame(d->frame_names && d->frame_names[f] ? d->frame_names[f] : ("frame" + std::to_string(f)),
                    d->frame_parent[f], pl, FRAME_OP);
    }
    if (m.nframes() == 0) m.add_frame("universe", 0, se3_identity(), FRAME_OP);
    *out = h;
    return IKB_OK;
}

void ikb_model_free(ikb_model *m) { delete m; }
int ikb_model_njoints(const ikb_model *m) { return m ? m->m.njoints() : -IKB_ERR_INVALID_ARG; }
int ikb_model_nq(const ikb_model *m) { return m ? m->m.nq : -IKB_ERR_INVALID_ARG; }
int ikb_model_nv(const ikb_model *m) { return m ? m->m.nv : -IKB_ERR_INVALID_ARG; }
int ikb_model_nframes(const ikb_model *m) { return m ? m->m.nframes() : -IKB_ERR_INVALID_ARG; }
int ikb_model_frame_id(const ikb_model *m, const char *name) {
    if (!m || !name) return -IKB_ERR_INVALID_ARG;
    return m->m.frame_id(name);
}
const char *ikb_model_joint_name(const ikb_model *m, int j) {
    return (m && j >= 0 && j < m->m.njoints()) ? m->m.joint_names[j].c_str() : nullptr;
}
const char *ikb_model_frame_name(const ikb_model *m, int f) {
    return (m && f >= 0 && f < m->m.nframes()) ? m->m.frame_names[f].c_str() : nullptr;
}
int ikb_model_get_topology(const ikb_model *m, int32_t *parent, int32_t *jtype, int32_t *idx_q, int32_t *idx_v) {
    if (!m) return fail(IKB_ERR_INVALID_ARG, "null model");
    const int n = m->m.njoints();
    if (parent) std::copy(m->m.parent.begin(), m->m.parent.end(), parent);
    if (jtype) std::copy(m->m.jtype.begin(), m->m.jtype.end(), jtype);
    if (idx_q) std::copy(m->m.idx_q.begin(), m->m.idx_q.end(), idx_q);
    if (idx_v) std::copy(m->m.idx_v.begin(), m->m.idx_v.end(), idx_v);
    (void)n;
    return IKB_OK;
}
int ikb_model_get_placements(const ikb_model *m, double *placement, double *axis) {
    if (!m) return fail(IKB_ERR_INVALID_ARG, "null model");
    for (int j = 0; j < m->m.njoints(); ++j) {
        if (placement) std::copy(m->m.placement[j].begin(), m->m.placement[j].end(), placement + 12 * j);
        if (axis) std::copy(m->m.axis[j].begin(), m->m.axis[j].end(), axis + 3 * j);
    }
    return IKB_OK;
}
int ikb_model_get_limits(const ikb_model *m, double *lower, double *upper) {
    if (!m) return fail(IKB_ERR_INVALID_ARG, "null model");
    if (lower) std::copy(m->m.lower.begin(), m->m.lower.end(), lower);
    if (upper) std::copy(m->m.upper.begin(), m->m.upper.end(), upper);
    return IKB_OK;
}
int ikb_model_set_limits(ikb_model *m, const double *lower, const double *upper) {
    if (!m) return fail(IKB_ERR_INVALID_ARG, "null model");
    if (lower) std::copy(lower, lower + m->m.nq, m->m.lower.begin());
    if (upper) std::copy(upper, upper + m->m.nq, m->m.upper.begin());
    return IKB_OK;
}
int ikb_model_get_frames(const ikb_model *m, int32_t *parent_joint, int32_t *type, double *placement) {
    if (!m) return fail(IKB_ERR_INVALID_ARG, "null model");
    for (int f = 0; f < m->m.nframes(); ++f) {
        if (parent_joint) parent_joint[f] = m->m.frame_parent[f];
        if (type) type[f] = m->m.frame_type[f];
        if (placement) std::copy(m->m.frame_placement[f].begin(), m->m.frame_placement[f].end(), placement + 12 * f);
    }
    return IKB_OK;
}
int ikb_model_neutral(const ikb_model *m, double *q) {
    if (!m || !q) return fail(IKB_ERR_INVALID_ARG, "null argument");
    std::fill(q, q + m->m.nq, 0.0);
    for (int j = 0; j < m->m.njoints(); ++j)
        if (m->m.jtype[j] == IKB_J_FREEFLYER) q[m->m.idx_q[j] + 6] = 1.0;
    return IKB_OK;
}

int ikb_problem_create(const ikb_model *m, int max_priority_level, ikb_problem **out) {
    if (!m || !out || max_priority_level < 0) return fail(IKB_ERR_INVALID_ARG, "ikb_problem_create: bad argument");
    auto *p = new ikb_problem();
    p->hp.model = m->m;  // copy, like InverseKinematicsProblem (problem.hpp:183)
    p->hp.max_priority_level = max_priority_level;
    *out = p;
    return IKB_OK;
}

void ikb_problem_free(ikb_problem *p) {
    if (!p) return;
    if (p->finalized) {
        DeviceGuard g(p->device);
        cudaFree(p->d64); cudaFree(p->d32); cudaFree(p->d_frame_parent); cudaFree(p->d_frame_pl64);
        cudaFree(p->d_frame_pl32); cudaFree(p->d_tickets);
        cudaFree(p->st64.q0); cudaFree(p->st64.targets); cudaFree(p->st64.q); cudaFree(p->st64.resid);
        cudaFree(p->st32.q0); cudaFree(p->st32.targets); cudaFree(p->st32.q); cudaFree(p->st32.resid);
        cudaFree(p->st_success); cudaFree(p->st_iters);
        for (auto &sc : p->scratch) {
            cudaFree(sc.list); cudaFree(sc.iters);
            if (sc.ev) cudaEventDestroy(sc.ev);
        }
        if (p->stream) cudaStreamDestroy(p->stream);
        if (p->stream_in) cudaStreamDestroy(p->stream_in);
        if (p->stream_aux) cudaStreamDestroy(p->stream_aux);
        for (auto e : p->ev_in) if (e) cudaEventDestroy(e);
        if (p->ev_aux) cudaEventDestroy(p->ev_aux);
        if (p->ev_main) cudaEventDestroy(p->ev_main);
    }
    delete p;
}

static int add_task_common(ikb_problem *p, HostTask &t, int priority, const double *weights) {
    if (p->finalized) return -fail(IKB_ERR_INVALID_ARG, "problem is finalized (immutable)");
    if (priority < 0 || priority > p->hp.max_priority_level)
        return -fail(IKB_ERR_INVALID_ARG, "priority exceeds max_priority_level (problem.hpp:160-164)");
    t.priority = priority;
    check_weights(weights, t.dim, t.weight);
    p->hp.tasks.push_back(t);
    return (int)p->hp.tasks.size() - 1;
}

int ikb_problem_add_frame_task(ikb_problem *p, int frame, int ktype, int ref, int priority, const double *weights) {
    if (!p) return -fail(IKB_ERR_INVALID_ARG, "null problem");
    const int nf = p->hp.model.nframes();
    if (frame < 0 || frame >= nf || ref < 0 || ref >= nf) return -fail(IKB_ERR_UNKNOWN_FRAME, "frame index out of range");
    if (ktype < IKB_POSITION || ktype > IKB_FULL) return -fail(IKB_ERR_INVALID_ARG, "bad kinematic type");
    HostTask t;
    t.kind = IKB_TASK_FRAME;
    t.frame = frame;
    t.ref = ref;
    t.type = ktype;
    t.dim = ktype == IKB_FULL ? 6 : 3;  // frame.hpp:100-107
    t.target_size = 12;
    return add_task_common(p, t, priority, weights);
}

int ikb_problem_add_align_axis_task(ikb_problem *p, int frame, int axis, int ref, int priority, const double *weights) {
    if (!p) return -fail(IKB_ERR_INVALID_ARG, "null problem");
    const int nf = p->hp.model.nframes();
    if (frame < 0 || frame >= nf || ref < 0 || ref >= nf) return -fail(IKB_ERR_UNKNOWN_FRAME, "frame index out of range");
    if (axis < 0 || axis > 2) return -fail(IKB_ERR_INVALID_ARG, "bad axis");
    HostTask t;
    t.kind = IKB_TASK_ALIGN_AXIS;
    t.frame = frame;
    t.ref = ref;
    t.type = axis;
    t.dim = 1;  // frame.hpp:226
    t.target_size = 3;
    return add_task_common(p, t, priority, weights);
}

int ikb_problem_add_posture_task(ikb_problem *p, int nj, int priority, const double *weights, const double *mask) {
    if (!p) return -fail(IKB_ERR_INVALID_ARG, "null problem");
    if (nj < 1 || nj > p->hp.model.nv || nj > p->hp.model.nq) return -fail(IKB_ERR_INVALID_ARG, "posture size out of range");
    HostTask t;
    t.kind = IKB_TASK_POSTURE;
    t.type = nj;
    t.dim = nj;  // posture.hpp:32
    t.target_size = nj;
    t.mask.assign(nj, 1.0);
    if (mask) std::copy(mask, mask + nj, t.mask.begin());
    return add_task_common(p, t, priority, weights);
}

int ikb_problem_num_tasks(const ikb_problem *p) { return p ? (int)p->hp.tasks.size() : -IKB_ERR_INVALID_ARG; }
int ikb_problem_task_dim(const ikb_problem *p, int t) {
    return (p && t >= 0 && t < (int)p->hp.tasks.size()) ? p->hp.tasks[t].dim : -IKB_ERR_INVALID_ARG;
}
int ikb_problem_e_size(const ikb_problem *p, int priority) { return p ? p->hp.e_size(priority) : -IKB_ERR_INVALID_ARG; }
int ikb_problem_rows(const ikb_problem *p) { return p ? p->hp.rows() : -IKB_ERR_INVALID_ARG; }
int ikb_problem_target_size(const ikb_problem *p) { return p ? p->hp.target_size() : -IKB_ERR_INVALID_ARG; }
int ikb_problem_task_target_offset(const ikb_problem *p, int t) {
    return (p && t >= 0 && t < (int)p->hp.tasks.size()) ? p->hp.target_offset(t) : -IKB_ERR_INVALID_ARG;
}

int ikb_problem_finalize(ikb_problem *p, int device) {
    if (!p) return fail(IKB_ERR_INVALID_ARG, "null problem");
    if (p->finalized) return fail(IKB_ERR_INVALID_ARG, "problem already finalized");
    const HostProblem &hp = p->hp;
    const HostModel &m = hp.model;
    if (hp.tasks.empty()) return fail(IKB_ERR_INVALID_ARG, "problem has no tasks");
    // stacked order: priority level, then insertion order (dls.cpp:18-24)
    std::vector<int> order(hp.tasks.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hp.tasks[a].priority < hp.tasks[b].priority; });
    std::vector<int> used;
    for (const auto &t : hp.tasks)
        if (t.kind != IKB_TASK_POSTURE)
            for (int f : {t.frame, t.ref})
                if (std::find(used.begin(), used.end(), f) == used.end()) used.push_back(f);
    if (used.empty()) used.push_back(0);
    if (m.njoints() > kMaxJoints || m.nq > kMaxNq || (int)hp.tasks.size() > kMaxTasks || (int)used.size() > kMaxFrames ||
        hp.rows() > kMaxRows)
        return fail(IKB_ERR_UNSUPPORTED, "problem exceeds the compiled capacities of the constant blob");
    int cls = -1;
    for (int c = 0; c < 3; ++c)
        if (m.njoints() <= kClasses[c].nj && m.nv <= kClasses[c].nv && hp.rows() <= kClasses[c].m) {
            cls = c;
            break;
        }
    if (cls < 0)
        return fail(IKB_ERR_UNSUPPORTED, "problem larger than the generic kernel's largest size class (32 joints, nv 36, 30 rows)");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(IKB_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(IKB_ERR_INVALID_ARG, "device index out of range");
    DeviceGuard g(device);
    if (!g.ok) return fail(IKB_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    IKB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(IKB_ERR_NO_DEVICE, std::string("device ") + prop.name + " is not Blackwell-class (sm_100a code only)");
    p->sm_count = prop.multiProcessorCount;

    auto *h64 = new DevProblem<double>();
    auto *h32 = new DevProblem<float>();
    fill_dev_problem(hp, order, used, *h64);
    fill_dev_problem(hp, order, used, *h32);
    IKB_CUDA(cudaMalloc(&p->d64, sizeof(*h64)));
    IKB_CUDA(cudaMalloc(&p->d32, sizeof(*h32)));
    IKB_CUDA(cudaMemcpy(p->d64, h64, sizeof(*h64), cudaMemcpyHostToDevice));
    IKB_CUDA(cudaMemcpy(p->d32, h32, sizeof(*h32), cudaMemcpyHostToDevice));
    delete h64;
    delete h32;

    const int nf = m.nframes();
    std::vector<double> pl64((size_t)nf * 12);
    std::vector<float> pl32((size_t)nf * 12);
    for (int f = 0; f < nf; ++f)
        for (int k = 0; k < 12; ++k) {
            pl64[12 * f + k] = m.frame_placement[f][k];
            pl32[12 * f + k] = (float)m.frame_placement[f][k];
        }
    IKB_CUDA(cudaMalloc(&p->d_frame_parent, nf * sizeof(int)));
    IKB_CUDA(cudaMalloc(&p->d_frame_pl64, pl64.size() * sizeof(double)));
    IKB_CUDA(cudaMalloc(&p->d_frame_pl32, pl32.size() * sizeof(float)));
    IKB_CUDA(cudaMemcpy(p->d_frame_parent, m.frame_parent.data(), nf * sizeof(int), cudaMemcpyHostToDevice));
    IKB_CUDA(cudaMemcpy(p->d_frame_pl64, pl64.data(), pl64.size() * sizeof(double), cudaMemcpyHostToDevice));
    IKB_CUDA(cudaMemcpy(p->d_frame_pl32, pl32.data(), pl32.size() * sizeof(float), cudaMemcpyHostToDevice));
    IKB_CUDA(cudaMalloc(&p->d_tickets, kTicketSlots * 16 * sizeof(unsigned long long)));
    IKB_CUDA(cudaMemset(p->d_tickets, 0, kTicketSlots * 16 * sizeof(unsigned long long)));
    IKB_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    IKB_CUDA(cudaStreamCreateWithFlags(&p->stream_in, cudaStreamNonBlocking));
    IKB_CUDA(cudaStreamCreateWithFlags(&p->stream_aux, cudaStreamNonBlocking));
    for (auto &e : p->ev_in) IKB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    IKB_CUDA(cudaEventCreateWithFlags(&p->ev_aux, cudaEventDisableTiming));
    IKB_CUDA(cudaEventCreateWithFlags(&p->ev_main, cudaEventDisableTiming));

    p->size_class = cls;
    p->device = device;
    p->weight_stacked.clear();
    for (int t : order)
        for (double w : hp.tasks[t].weight) p->weight_stacked.push_back(w);
    p->spec = select_specialized(hp);
    char buf[96];
    std::snprintf(buf, sizeof buf, "generic<NJ=%d,NV=%d,M=%d>", kClasses[cls].nj, kClasses[cls].nv, kClasses[cls].m);
    p->kernel_name[0] = p->spec ? p->spec->name : buf;
    p->kernel_name[1] = p->kernel_name[0];
    p->finalized = true;
    return IKB_OK;
}

const char *ikb_problem_specialisation(const ikb_problem *p) {
    if (!p || p->hp.tasks.empty()) return nullptr;
    const SpecializedKernel *k = select_specialized(p->hp);
    return k ? k->name : nullptr;
}

const char *ikb_problem_kernel_name(const ikb_problem *p, int dtype) {
    if (!p || !p->finalized || dtype < 0 || dtype > 1) return nullptr;
    return p->kernel_name[dtype].c_str();
}

static int check_solve_args(const ikb_problem *p, int dtype, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io) {
    if (!p || !prm || !io) return fail(IKB_ERR_INVALID_ARG, "null argument");
    if (!p->finalized) return fail(IKB_ERR_NOT_FINALIZED, "call ikb_problem_finalize first");
    if (dtype != IKB_F64 && dtype != IKB_F32) return fail(IKB_ERR_INVALID_ARG, "dtype must be IKB_F64 or IKB_F32");
    if (B < 0 || prm->max_iterations < 0) return fail(IKB_ERR_INVALID_ARG, "negative batch size or iteration count");
    if (B > 0 && (!io->q0 || !io->q || (!io->targets && p->hp.target_size() > 0)))
        return fail(IKB_ERR_INVALID_ARG, "q0, targets and q must be non-null");
    return IKB_OK;
}

int ikb_dls_solve_batch(const ikb_problem *p, int dtype, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io,
                        void *cuda_stream) {
    int rc = check_solve_args(p, dtype, prm, B, io);
    if (rc) return rc;
    if (B == 0) return IKB_OK;
    DeviceGuard g(p->device);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    return dtype == IKB_F64 ? launch_solve<double>(p, prm, B, io, s) : launch_solve<float>(p, prm, B, io, s);
}

int ikb_dls_solve_batch_host(ikb_problem *p, int dtype, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io) {
    int rc = check_solve_args(p, dtype, prm, B, io);
    if (rc) return rc;
    if (B == 0) return IKB_OK;
    DeviceGuard g(p->device);
    return dtype == IKB_F64 ? solve_host<double>(p, prm, B, io) : solve_host<float>(p, prm, B, io);
}

/* ---- pipelined queue ---- */
int ikb_queue_create(ikb_problem *p, int depth, int merge, ikb_queue **out) {
    if (!p || !out) return fail(IKB_ERR_INVALID_ARG, "null argument");
    if (!p->finalized) return fail(IKB_ERR_NOT_FINALIZED, "call ikb_problem_finalize first");
    if (depth < 1 || depth > 16) return fail(IKB_ERR_INVALID_ARG, "queue depth must be between 1 and 16");
    if (merge < 1 || merge > kMaxMerge || merge > depth) return fail(IKB_ERR_INVALID_ARG, "merge must be between 1 and min(depth, 8)");
    DeviceGuard g(p->device);
    ikb_queue *q = new ikb_queue;
    q->p = p;
    q->depth = depth;
    q->merge = merge;
    q->slots.resize(depth);
    *out = q;  // the caller frees it also when creation fails half-way
    for (cudaStream_t *s : {&q->s_in, &q->s_comp, &q->s_out}) IKB_CUDA(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking));
    IKB_CUDA(cudaEventCreateWithFlags(&q->ev_user, cudaEventDisableTiming));
    IKB_CUDA(cudaEventCreateWithFlags(&q->ev_comp, cudaEventDisableTiming));
    for (auto &sl : q->slots)
        for (cudaEvent_t *e : {&sl.ev_in, &sl.ev_done}) IKB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return IKB_OK;
}

void ikb_queue_free(ikb_queue *q) {
    if (!q) return;
    DeviceGuard g(q->p->device);
    queue_flush(q);
    for (cudaStream_t s : {q->s_in, q->s_comp, q->s_out})
        if (s) {
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        }
    if (q->ev_user) cudaEventDestroy(q->ev_user);
    if (q->ev_comp) cudaEventDestroy(q->ev_comp);
    for (auto &sl : q->slots) {
        for (cudaEvent_t e : {sl.ev_in, sl.ev_done})
            if (e) cudaEventDestroy(e);
        cudaFree(sl.st64.q0); cudaFree(sl.st64.targets); cudaFree(sl.st64.q); cudaFree(sl.st64.resid);
        cudaFree(sl.st32.q0); cudaFree(sl.st32.targets); cudaFree(sl.st32.q); cudaFree(sl.st32.resid);
        cudaFree(sl.success); cudaFree(sl.iters);
    }
    delete q;
}

int64_t ikb_queue_submit(ikb_queue *q, int dtype, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io, void *in_stream) {
    if (!q) return -fail(IKB_ERR_INVALID_ARG, "null queue");
    int rc = check_solve_args(q->p, dtype, prm, B, io);
    if (rc) return -rc;
    DeviceGuard g(q->p->device);
    ikb_queue::Slot *sl;
    if ((rc = queue_acquire(q, dtype, prm, &sl))) return -rc;
    // the inputs are ready in `in_stream` order (NULL = the legacy default stream) at this point
    if (cudaEventRecord(q->ev_user, (cudaStream_t)in_stream) != cudaSuccess || cudaStreamWaitEvent(q->s_comp, q->ev_user, 0) != cudaSuccess)
        return -cuda_fail(cudaGetLastError(), "queue input dependency");
    sl->host = false;
    sl->B = B;
    sl->dio = *io;
    return queue_commit(q, sl);
}

int64_t ikb_queue_submit_host(ikb_queue *q, int dtype, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io) {
    if (!q) return -fail(IKB_ERR_INVALID_ARG, "null queue");
    int rc = check_solve_args(q->p, dtype, prm, B, io);
    if (rc) return -rc;
    DeviceGuard g(q->p->device);
    ikb_queue::Slot *sl;
    if ((rc = queue_acquire(q, dtype, prm, &sl))) return -rc;
    sl->host = true;
    sl->B = B;
    if (B > 0) {
        rc = dtype == IKB_F64 ? queue_stage_host<double>(q, *sl, B, io) : queue_stage_host<float>(q, *sl, B, io);
        if (rc) return -rc;
    } else if (cudaEventRecord(sl->ev_in, q->s_in) != cudaSuccess) {
        return -cuda_fail(cudaGetLastError(), "cudaEventRecord");
    }
    return queue_commit(q, sl);
}

int ikb_queue_flush(ikb_queue *q) {
    if (!q) return fail(IKB_ERR_INVALID_ARG, "null queue");
    DeviceGuard g(q->p->device);
    return queue_flush(q);
}

int ikb_queue_wait(ikb_queue *q, int64_t ticket) {
    if (!q || ticket < 0 || ticket >= q->next) return fail(IKB_ERR_INVALID_ARG, "unknown queue ticket");
    ikb_queue::Slot &sl = q->slots[ticket % q->depth];
    if (sl.ticket != ticket) return IKB_OK;  // its slot has been reused: it left the pipeline long ago
    DeviceGuard g(q->p->device);
    int rc;
    if (sl.pending && (rc = queue_flush(q))) return rc;
    IKB_CUDA(cudaEventSynchronize(sl.ev_done));
    sl.busy = false;
    return IKB_OK;
}

int ikb_queue_wait_on_stream(ikb_queue *q, int64_t ticket, void *cuda_stream) {
    if (!q || ticket < 0 || ticket >= q->next) return fail(IKB_ERR_INVALID_ARG, "unknown queue ticket");
    ikb_queue::Slot &sl = q->slots[ticket % q->depth];
    if (sl.ticket != ticket) return IKB_OK;
    DeviceGuard g(q->p->device);
    int rc;
    if (sl.pending && (rc = queue_flush(q))) return rc;
    IKB_CUDA(cudaStreamWaitEvent((cudaStream_t)cuda_stream, sl.ev_done, 0));
    return IKB_OK;
}

int ikb_queue_drain(ikb_queue *q) {
    if (!q) return fail(IKB_ERR_INVALID_ARG, "null queue");
    DeviceGuard g(q->p->device);
    int rc = queue_flush(q);
    if (rc) return rc;
    for (cudaStream_t s : {q->s_in, q->s_comp, q->s_out}) IKB_CUDA(cudaStreamSynchronize(s));
    for (auto &sl : q->slots) sl.busy = false;
    return IKB_OK;
}

int ikb_dls_solve(ikb_problem *p, const ikb_dls_params *prm, const double *q0, const double *targets, double *q_out,
                  int *success, int *iters, double *resid) {
    if (!p) return fail(IKB_ERR_INVALID_ARG, "null problem");
    ikb_dls_params dflt;
    ikb_dls_params_default(&dflt);
    const int nq = p->hp.model.nq, tsz = p->hp.target_size();
    uint8_t ok = 0;
    int32_t it = 0;
    double r = 0;
    ikb_batch_io io;
    io.q0 = q0; io.q0_elem_stride = 1; io.q0_batch_stride = nq;
    io.targets = targets; io.targets_elem_stride = 1; io.targets_batch_stride = tsz;
    io.q = q_out; io.q_elem_stride = 1; io.q_batch_stride = nq;
    io.success = &ok; io.iters = &it; io.resid = &r;
    int rc = ikb_dls_solve_batch_host(p, IKB_F64, prm ? prm : &dflt, 1, &io);
    if (rc) return rc;
    if (success) *success = ok;
    if (iters) *iters = it;
    if (resid) *resid = r;
    return IKB_OK;
}

int ikb_fk_batch(const ikb_problem *p, int dtype, int64_t B, const void *q, int64_t es, int64_t bs, int nf,
                 const int32_t *frames, void *out, void *cuda_stream) {
    if (!p || !q || !out || !frames) return fail(IKB_ERR_INVALID_ARG, "null argument");
    if (!p->finalized) return fail(IKB_ERR_NOT_FINALIZED, "call ikb_problem_finalize first");
    if (nf < 1 || nf > 16) return fail(IKB_ERR_INVALID_ARG, "between 1 and 16 frames per call");
    FrameList fl;
    fl.n = nf;
    for (int i = 0; i < nf; ++i) {
        if (frames[i] < 0 || frames[i] >= p->hp.model.nframes()) return fail(IKB_ERR_UNKNOWN_FRAME, "frame index out of range");
        fl.id[i] = frames[i];
    }
    if (B <= 0) return IKB_OK;
    DeviceGuard g(p->device);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    return dtype == IKB_F64 ? launch_fk<double>(p, B, q, es, bs, fl, out, s) : launch_fk<float>(p, B, q, es, bs, fl, out, s);
}

}  // extern "C"
