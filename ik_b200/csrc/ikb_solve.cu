// The solve launcher and the host-buffer path of the C ABI (include/ikb200.h): ikb_dls_solve_batch,
// ikb_dls_solve_batch_host, ikb_dls_solve.  The pipelined queue (ikb_queue.cu) launches through launch_solve too.
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "capi_internal.hpp"
#include "dls_generic.cuh"

using namespace ikb;
using namespace ikb::capi;

namespace {

template <typename T> struct KernelTable {
    using Fn = void (*)(const DevProblem<T> *, SolveArgs<T>);
    static Fn dls(int cls) {
        switch (cls) {
            case 0: return dls_generic_kernel<T, 10, 8, 6>;
            case 1: return dls_generic_kernel<T, 20, 24, 12>;
            default: return dls_generic_kernel<T, 32, 36, 30>;
        }
    }
    static Fn pik(int cls) {  // ik::pik (pik.cpp:31-96)
        switch (cls) {
            case 0: return dls_generic_kernel<T, 10, 8, 6, true>;
            case 1: return dls_generic_kernel<T, 20, 24, 12, true>;
            default: return dls_generic_kernel<T, 32, 36, 30, true>;
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

__global__ void set_ticket_kernel(unsigned long long *t, unsigned long long v) { *t = v; }
}  // namespace

namespace ikb {
namespace capi {

int check_solve_args(const ikb_problem *p, int dtype, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io) {
    if (!p || !prm || !io) return fail(IKB_ERR_INVALID_ARG, "null argument");
    if (!p->finalized) return fail(IKB_ERR_NOT_FINALIZED, "call ikb_problem_finalize first");
    if (dtype != IKB_F64 && dtype != IKB_F32) return fail(IKB_ERR_INVALID_ARG, "dtype must be IKB_F64 or IKB_F32");
    if (B < 0 || prm->max_iterations < 0) return fail(IKB_ERR_INVALID_ARG, "negative batch size or iteration count");
    if (B > 0 && (!io->q0 || !io->q || (!io->targets && p->hp.target_size() > 0)))
        return fail(IKB_ERR_INVALID_ARG, "q0, targets and q must be non-null");
    return IKB_OK;
}

// Is this solve going to take the two-launch (BULK + TAIL) path?  (the only one that can be pipelined by slices)
bool two_phase(const ikb_problem *p, const ikb_dls_params *prm, int64_t B, int *cap_out) {
    const char *cap_env = std::getenv("IKB_BULK_CAP");
    const int cap = cap_env ? std::atoi(cap_env) : 16;
    if (cap_out) *cap_out = cap;
    return p->spec && B > 2LL * 32 * p->sm_count && cap > 0 && prm->max_iterations > cap;
}

// Scratch of a two-launch solve (suspended-problem list, step counts): slot `slot` of the problem's ring, grown to B
// entries; stream `s` waits for the slot's previous user.
static int acquire_scratch(const ikb_problem *p, unsigned slot, int64_t B, cudaStream_t s, SolveScratch **out) {
    ikb_problem *mp = const_cast<ikb_problem *>(p);
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
    *out = &mp->scratch[slot % kScratchSlots];
    IKB_CUDA(cudaStreamWaitEvent(s, (*out)->ev, 0));  // the slot's previous user (any stream) must be done
    return IKB_OK;
}

template <typename T>
int launch_solve(const ikb_problem *p, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io, cudaStream_t s,
                 const ChunkPlan *plan, const Merged<T> *merged, const double *pik_lambda) {
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
    for (int l = 0; l < 7; ++l) a.pik_lambda2[l] = pik_lambda ? (T)(pik_lambda[l] * pik_lambda[l]) : T(0);
    if (p->spec && !pik_lambda) {
        const SpecHostConsts hc{p->hp.model.lower.data(), p->hp.model.upper.data(), p->weight_stacked.data(), p->mask_stacked.data()};
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
            SolveScratch *sc;
            if ((rc = acquire_scratch(p, slot, B, s, &sc))) return rc;
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
    if (merged) return fail(IKB_ERR_INVALID_ARG, "internal: merged launch on the table-driven kernel");
    // Table-driven problems: the team-per-problem kernel (dls_coop.cuh; J in shared memory, Gram rows and the factorisation
    // in registers).  IKB_GENERIC_LEGACY=1 keeps the thread-per-problem local-memory kernel (dls_generic.cuh) for A/B runs;
    // it is also the fallback for problems beyond the cooperative kernel's table capacities.
    const char *legacy_env = std::getenv("IKB_GENERIC_LEGACY");
    if (p->coop_ok && !(legacy_env && legacy_env[0] == '1')) {
        const char *shfl_env = std::getenv("IKB_COOP_SHFL");
        const bool shfl = shfl_env ? shfl_env[0] == '1' : true;
        if (launch_coop<T>(p->size_class, dev_blob<T>(p), a, pik_lambda != nullptr, !p->hp.constraints.empty(), shfl, p->sm_count, s))
            return cuda_fail(cudaGetLastError(), "team-per-problem kernel launch");
        g_launches.fetch_add(1);
        return IKB_OK;
    }
    auto fn = pik_lambda ? KernelTable<T>::pik(p->size_class) : KernelTable<T>::dls(p->size_class);
    const char *thr_env = std::getenv("IKB_GENERIC_THREADS"), *bps_env = std::getenv("IKB_GENERIC_BLOCKS_PER_SM");
    const int threads = thr_env ? std::max(32, std::min(128, std::atoi(thr_env) / 32 * 32)) : 128;
    const size_t smem = sizeof(DevProblem<T>) + 16;
    int per_sm = 0;
    IKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1) per_sm = 1;
    if (bps_env && std::atoi(bps_env) > 0) per_sm = std::min(per_sm, std::atoi(bps_env));
    long long blocks = (B + threads - 1) / threads;
    blocks = std::min<long long>(blocks, (long long)per_sm * p->sm_count);
    // Two launches for a batch that more than fills the GPU: problems unfinished after `cap` steps are
    // parked and a second launch continues them, 32 to a warp and one warp per CTA, so that a straggler's local-memory
    // scratch shares its cache lines with 31 others and stays in L1 instead of thrashing it from a mostly idle warp.
    const char *cap_env = std::getenv("IKB_GENERIC_CAP");
    const int cap = cap_env ? std::atoi(cap_env) : 32;  // measured: profiles/r1_generic_two_launch.txt (16 suits quick problems, 32 never loses to one launch)
    if (cap > 0 && prm->max_iterations > cap && B > 2048) {
        SolveScratch *sc;
        int rc;
        if ((rc = acquire_scratch(p, slot, B, s, &sc))) return rc;
        IKB_CUDA(cudaMemsetAsync(a.ticket, 0, 3 * sizeof(unsigned long long), s));  // first ticket, second ticket, list count
        a.it_cap = cap;
        a.list = sc->list;
        a.list_count = a.ticket + 2;
        a.iters_ws = io->iters ? io->iters : sc->iters;
        fn<<<(unsigned)blocks, threads, smem, s>>>(dev_blob<T>(p), a);
        IKB_CUDA(cudaGetLastError());
        SolveArgs<T> t = a;
        t.resume = 1;
        t.it_cap = INT_MAX;
        t.ticket = a.ticket + 1;
        int per_sm1 = 0;
        IKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, fn, 32, smem));
        const long long blocks1 = std::min<long long>((B + 31) / 32, (long long)std::max(per_sm1, 1) * p->sm_count);
        fn<<<(unsigned)blocks1, 32, smem, s>>>(dev_blob<T>(p), t);
        IKB_CUDA(cudaGetLastError());
        IKB_CUDA(cudaEventRecord(sc->ev, s));
        g_launches.fetch_add(2);
        return IKB_OK;
    }
    fn<<<(unsigned)blocks, threads, smem, s>>>(dev_blob<T>(p), a);
    IKB_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return IKB_OK;
}

template int launch_solve<double>(const ikb_problem *, const ikb_dls_params *, int64_t, const ikb_batch_io *, cudaStream_t,
                                  const ChunkPlan *, const Merged<double> *, const double *);
template int launch_solve<float>(const ikb_problem *, const ikb_dls_params *, int64_t, const ikb_batch_io *, cudaStream_t,
                                 const ChunkPlan *, const Merged<float> *, const double *);

}  // namespace capi
}  // namespace ikb

namespace {
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
int solve_host(ikb_problem *p, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io, const double *pik_lambda = nullptr) {
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
    if (!pik_lambda && two_phase(p, prm, B) && nslice >= 2 && vq.sliceable(B) && vt.sliceable(B) && !(pipe_env && pipe_env[0] == '0')) {
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
        if ((rc = launch_solve<T>(p, prm, B, &dio, s, nullptr, nullptr, pik_lambda))) return rc;
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
}  // namespace

extern "C" {

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

/* ---- ik::pik ---- */
void ikb_pik_params_default(ikb_pik_params *p) {
    if (!p) return;
    p->max_iterations = 100;  // pik.hpp:14
    p->step_length = 1.0;     // pik.hpp:16
    p->tolerance = 1e-4;      // visitor.hpp:19
    for (double &l : p->lambda) l = 1.0;  // pik_data::lambda (pik.hpp:31)
}

static int pik_to_dls(const ikb_problem *p, const ikb_pik_params *prm, ikb_dls_params *out) {
    if (!p || !prm) return fail(IKB_ERR_INVALID_ARG, "null argument");
    if (p->hp.max_priority_level + 1 > 7) return fail(IKB_ERR_UNSUPPORTED, "ik::pik supports at most 7 priority levels");
    ikb_dls_params_default(out);
    out->max_iterations = prm->max_iterations;
    out->step_length = prm->step_length;
    out->tolerance = prm->tolerance;
    return IKB_OK;
}

int ikb_pik_solve_batch(const ikb_problem *p, int dtype, const ikb_pik_params *prm, int64_t B, const ikb_batch_io *io, void *cuda_stream) {
    ikb_dls_params d;
    int rc = pik_to_dls(p, prm, &d);
    if (rc || (rc = check_solve_args(p, dtype, &d, B, io))) return rc;
    if (B == 0) return IKB_OK;
    DeviceGuard g(p->device);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    return dtype == IKB_F64 ? launch_solve<double>(p, &d, B, io, s, nullptr, nullptr, prm->lambda)
                            : launch_solve<float>(p, &d, B, io, s, nullptr, nullptr, prm->lambda);
}

int ikb_pik_solve_batch_host(ikb_problem *p, int dtype, const ikb_pik_params *prm, int64_t B, const ikb_batch_io *io) {
    ikb_dls_params d;
    int rc = pik_to_dls(p, prm, &d);
    if (rc || (rc = check_solve_args(p, dtype, &d, B, io))) return rc;
    if (B == 0) return IKB_OK;
    DeviceGuard g(p->device);
    return dtype == IKB_F64 ? solve_host<double>(p, &d, B, io, prm->lambda) : solve_host<float>(p, &d, B, io, prm->lambda);
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

}  // extern "C"
