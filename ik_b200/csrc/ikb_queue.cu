// The pipelined queue of the C ABI (include/ikb200.h, ikb_queue_*).
#include <chrono>
#include <cstdio>
#include <utility>

#include "capi_internal.hpp"

using namespace ikb;
using namespace ikb::capi;


// ---------------------------------------------------------------------------------------------------
// pipelined queue: batches in flight on three streams (copy-in, compute, copy-out); consecutive batches are MERGED
// into one BULK + TAIL launch pair (the straggler chain of ~0.7 ms is paid once per group instead of once per batch)
// ---------------------------------------------------------------------------------------------------

struct ikb_queue {
    struct Slot {
        cudaEvent_t ev_in = nullptr, ev_done = nullptr;
        cudaEvent_t tr_in0 = nullptr, tr_c0 = nullptr, tr_c1 = nullptr;  // IKB_QUEUE_TRACE only
        double tr_submit_ms = 0;
        bool busy = false, pending = false, host = false;
        bool deferred = false;  // launched, but its group's stragglers are still carried (ev_done not recorded yet)
        int64_t ticket = -1;  // the batch occupying the slot
        int64_t B = 0;
        ikb_batch_io dio{};   // device view of the batch
        ikb_batch_io hio{};   // host mode: the caller's buffers (copy-out targets)
        HostViews hv;         // host mode: the caller's views, classified
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
    // Carried stragglers: the last merged device-buffer group was launched WITHOUT its TAIL; the next such group's BULK
    // launch continues its stragglers (capi_internal.hpp, CarryState), and whoever needs the group's results before that
    // (wait, flush, drain, slot reuse, a group that cannot carry) launches the TAIL (queue_finish_carry).
    bool carry_on = true;           // IKB_QUEUE_CARRY=0 disables it (A/B runs)
    bool carry_host = false;        // host batches carry too: depth >= 3 * merge, or IKB_QUEUE_CARRY_HOST=0|1 (ikb_queue_create)
    bool carry_valid = false;
    int carry_dtype = -1;
    ikb_dls_params carry_prm{};
    std::vector<int> carry_slots;
    CarryState<double> carry64;
    CarryState<float> carry32;
    CarryScratch cscratch[2];
    int cscratch_next = 0;
    // IKB_QUEUE_TRACE=1: device timeline of every host batch on stderr (printed by ikb_queue_wait)
    bool trace = false, tr_started = false;
    cudaEvent_t tr_ref = nullptr;
    std::chrono::steady_clock::time_point tr_host_ref;
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
    const ikb_batch_io &io = sl.hio;
    Staging<T> &st = slot_staging<T>(sl);
    int rc;
    if ((rc = copy_view_out<T>(sl.hv.q, st.q, sl.B, q->s_out))) return rc;   // only the payload crosses PCIe (2-D copy for sliced views)
    if (io.success) IKB_CUDA(cudaMemcpyAsync(io.success, sl.success, (size_t)sl.B, cudaMemcpyDeviceToHost, q->s_out));
    if (io.iters) IKB_CUDA(cudaMemcpyAsync(io.iters, sl.iters, (size_t)sl.B * sizeof(int), cudaMemcpyDeviceToHost, q->s_out));
    if (io.resid) IKB_CUDA(cudaMemcpyAsync(io.resid, st.resid, (size_t)sl.B * sizeof(T), cudaMemcpyDeviceToHost, q->s_out));
    return IKB_OK;
}

template <typename T> CarryState<T> &carry_state(ikb_queue *q);
template <> CarryState<double> &carry_state<double>(ikb_queue *q) { return q->carry64; }
template <> CarryState<float> &carry_state<float>(ikb_queue *q) { return q->carry32; }

// The carried group's results become final with what has been enqueued on the compute stream so far: device batches
// are done there, host batches get their copy-out behind it.
int queue_complete_carry(ikb_queue *q) {
    bool any_host = false;
    for (int i : q->carry_slots) any_host |= q->slots[i].host;
    if (any_host) {
        IKB_CUDA(cudaEventRecord(q->ev_comp, q->s_comp));
        IKB_CUDA(cudaStreamWaitEvent(q->s_out, q->ev_comp, 0));
    }
    for (int i : q->carry_slots) {
        ikb_queue::Slot &sl = q->slots[i];
        if (sl.host) {
            int rc = IKB_OK;
            if (sl.B > 0) rc = q->carry_dtype == IKB_F64 ? queue_copy_out<double>(q, sl) : queue_copy_out<float>(q, sl);
            if (rc) return rc;
            IKB_CUDA(cudaEventRecord(sl.ev_done, q->s_out));
        } else {
            IKB_CUDA(cudaEventRecord(sl.ev_done, q->s_comp));
        }
        sl.deferred = false;
    }
    q->carry_slots.clear();
    q->carry_valid = false;
    q->carry64.valid = false;
    q->carry32.valid = false;
    return IKB_OK;
}
// Nobody will continue the carried stragglers: give them their TAIL launch.
int queue_finish_carry(ikb_queue *q) {
    if (!q->carry_valid) return IKB_OK;
    const int rc = q->carry_dtype == IKB_F64 ? launch_carry_tail<double>(q->p, q->carry64, q->s_comp) : launch_carry_tail<float>(q->p, q->carry32, q->s_comp);
    const int rc2 = queue_complete_carry(q);
    return rc ? rc : rc2;
}
int carry_scratch_reserve(CarryScratch &c, size_t n) {
    if (!c.counters) {
        IKB_CUDA(cudaMalloc(&c.counters, 4 * sizeof(unsigned long long)));
        IKB_CUDA(cudaMemset(c.counters, 0, 4 * sizeof(unsigned long long)));
    }
    if (c.cap >= n) return IKB_OK;
    // (the old buffers may still be read by kernels in flight: the queue's streams are drained first)
    IKB_CUDA(cudaDeviceSynchronize());
    cudaFree(c.list); cudaFree(c.iters);
    c.list = nullptr; c.iters = nullptr; c.cap = 0;
    IKB_CUDA(cudaMalloc(&c.list, n * sizeof(unsigned int)));
    IKB_CUDA(cudaMalloc(&c.iters, n * sizeof(int)));
    c.cap = n;
    return IKB_OK;
}

// Launch the open group: one merged BULK + TAIL pair when the problem has a specialised kernel, else batch by batch.
template <typename T> int queue_flush_t(ikb_queue *q) {
    const int n = (int)q->open.size();
    int rc;
    for (int i : q->open)
        if (q->slots[i].host) {
            ikb_queue::Slot &sl = q->slots[i];
            IKB_CUDA(cudaStreamWaitEvent(q->s_comp, sl.ev_in, 0));
            if (sl.B > 0 && (rc = expand_staged<T>(q->p, slot_staging<T>(sl), sl.hv, sl.B, 0, sl.B, true, q->s_comp))) return rc;
        }
    if (q->trace)
        for (int i : q->open) IKB_CUDA(cudaEventRecord(q->slots[i].tr_c0, q->s_comp));
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
        // (host batches carry only in a deep queue, see ikb_queue_create.  An earlier measurement -- 114 M against 137 M solves / s
        // end to end -- blamed the later copy-out; the real cause was the SE3 expansion kernel in the copy stream, queue_stage_host)
        bool any_host_ = false;
        for (int i : q->open) any_host_ |= q->slots[i].host;
        const bool can_carry = q->carry_on && (!any_host_ || q->carry_host) && !q->trace && two_phase(q->p, &q->open_prm, total);
        // a carried group that this launch cannot continue gets its TAIL now
        if (q->carry_valid && !(can_carry && q->carry_dtype == q->open_dtype && same_params(q->carry_prm, q->open_prm)) && (rc = queue_finish_carry(q)))
            return rc;
        if (can_carry) {
            CarryScratch &own = q->cscratch[q->cscratch_next];
            if ((size_t)total > own.cap) {   // growing the scratch synchronises: finish what is carried first; both sets, with head room
                if ((rc = queue_finish_carry(q))) return rc;
                const size_t want = (size_t)total + (size_t)total / 2;
                for (auto &c : q->cscratch)
                    if ((rc = carry_scratch_reserve(c, want))) return rc;
            }
            CarryState<T> out;
            if ((rc = launch_merged_carry<T>(q->p, &q->open_prm, total, &m, q->s_comp, own, q->carry_valid ? &carry_state<T>(q) : nullptr, &out))) return rc;
            q->cscratch_next ^= 1;
            if (q->carry_valid && (rc = queue_complete_carry(q))) return rc;   // the previous group's stragglers ran in this launch
            carry_state<T>(q) = out;
            q->carry_valid = true;
            q->carry_dtype = q->open_dtype;
            q->carry_prm = q->open_prm;
            q->carry_slots = q->open;
            for (int i : q->open) {
                q->slots[i].pending = false;
                q->slots[i].deferred = true;
            }
            q->open.clear();
            return IKB_OK;
        }
        if ((rc = launch_solve<T>(q->p, &q->open_prm, total, nullptr, q->s_comp, nullptr, &m))) return rc;
    } else {
        if (q->carry_valid && (rc = queue_finish_carry(q))) return rc;
        for (int i : q->open) {
            ikb_queue::Slot &sl = q->slots[i];
            if (sl.B > 0 && (rc = launch_solve<T>(q->p, &q->open_prm, sl.B, &sl.dio, q->s_comp))) return rc;
        }
    }
    bool any_host = false;
    for (int i : q->open) any_host |= q->slots[i].host;
    if (q->trace)
        for (int i : q->open) IKB_CUDA(cudaEventRecord(q->slots[i].tr_c1, q->s_comp));
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
    const int rc = q->open_dtype == IKB_F64 ? queue_flush_t<double>(q) : queue_flush_t<float>(q);
    if (rc != IKB_OK) {
        // a failed launch leaves no batch in flight: release the group's slots so that a retry does not push them twice
        for (int i : q->open) {
            q->slots[i].pending = false;
            q->slots[i].busy = false;
        }
        q->open.clear();
    }
    return rc;
}

// The slot of the next batch, free of its previous occupant (back-pressure: blocks while that batch is in flight).
int queue_acquire(ikb_queue *q, int dtype, const ikb_dls_params *prm, ikb_queue::Slot **out) {
    int rc;
    ikb_queue::Slot *sl = &q->slots[q->next % q->depth];
    if (sl->pending && (rc = queue_flush(q))) return rc;
    if (sl->deferred && (rc = queue_finish_carry(q))) return rc;
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
    Staging<T> &st = slot_staging<T>(sl);
    int rc;
    if ((rc = classify_host_views(p, B, io, &sl.hv))) return rc;
    sl.hio = *io;
    sl.dio = *io;
    if ((rc = prepare_staging<T>(p, st, sl.hv, B, &sl.dio))) return rc;
    if ((size_t)B > sl.flag_cap) {
        if (sl.success) cudaFree(sl.success);
        if (sl.iters) cudaFree(sl.iters);
        sl.success = nullptr; sl.iters = nullptr; sl.flag_cap = 0;
        IKB_CUDA(cudaMalloc(&sl.success, (size_t)B));
        IKB_CUDA(cudaMalloc(&sl.iters, (size_t)B * sizeof(int)));
        sl.flag_cap = (size_t)B;
    }
    if (q->trace) {
        if (!q->tr_started) {
            IKB_CUDA(cudaEventRecord(q->tr_ref, q->s_in));
            q->tr_host_ref = std::chrono::steady_clock::now();
            q->tr_started = true;
        }
        sl.tr_submit_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - q->tr_host_ref).count();
        IKB_CUDA(cudaEventRecord(sl.tr_in0, q->s_in));
    }
    // copy-in on the copy stream: copies ONLY.  The SE3 expansion of compact targets is a kernel; in the copy stream it could not
    // start while a persistent solve kernel holds every SM's registers, and the copies of the following batches would queue up
    // behind it until that kernel ends (measured: 0.5-0.7 ms of idle GPU per group).  It runs on the compute stream, in front of
    // the group's launch (queue_flush_t).
    if ((rc = stage_inputs<T>(p, st, sl.hv, B, 0, B, true, q->s_in, false))) return rc;
    IKB_CUDA(cudaEventRecord(sl.ev_in, q->s_in));
    sl.dio.success = sl.success; sl.dio.iters = sl.iters;
    return IKB_OK;
}
}  // namespace

extern "C" {

int ikb_queue_create(ikb_problem *p, int depth, int merge, ikb_queue **out) {
    if (!p || !out) return fail(IKB_ERR_INVALID_ARG, "null argument");
    if (!p->finalized) return fail(IKB_ERR_NOT_FINALIZED, "call ikb_problem_finalize first");
    if (depth < 1 || depth > 32) return fail(IKB_ERR_INVALID_ARG, "queue depth must be between 1 and 32");
    if (merge < 1 || merge > kMaxMerge || merge > depth) return fail(IKB_ERR_INVALID_ARG, "merge must be between 1 and min(depth, 8)");
    DeviceGuard g(p->device);
    ikb_queue *q = new ikb_queue;
    q->p = p;
    q->depth = depth;
    q->merge = merge;
    q->slots.resize(depth);
    *out = q;  // the caller frees it also when creation fails half-way
    for (cudaStream_t *s : {&q->s_in, &q->s_comp, &q->s_out}) IKB_CUDA(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking));
    const char *tr = std::getenv("IKB_QUEUE_TRACE");
    q->trace = tr && tr[0] == '1';
    const char *ce = std::getenv("IKB_QUEUE_CARRY");
    q->carry_on = !(ce && ce[0] == '0');
    // Host batches carry too when the queue is deep enough to hide the later copy-out (a carried group's results leave the
    // device one launch later): three groups in flight.  Measured (tools/e2e_host_probe.py, compact wire format): depth 12,
    // merge 4 -- 0.38 ms per step against 0.47 with a TAIL launch per group; depth 8, merge 4 would stall on the host's waits.
    const char *ch = std::getenv("IKB_QUEUE_CARRY_HOST");
    q->carry_host = ch ? ch[0] == '1' : depth >= 3 * merge;
    const unsigned evf = q->trace ? cudaEventDefault : cudaEventDisableTiming;
    IKB_CUDA(cudaEventCreateWithFlags(&q->ev_user, cudaEventDisableTiming));
    IKB_CUDA(cudaEventCreateWithFlags(&q->ev_comp, cudaEventDisableTiming));
    for (auto &sl : q->slots)
        for (cudaEvent_t *e : {&sl.ev_in, &sl.ev_done}) IKB_CUDA(cudaEventCreateWithFlags(e, evf));
    if (q->trace) {
        IKB_CUDA(cudaEventCreate(&q->tr_ref));
        for (auto &sl : q->slots)
            for (cudaEvent_t *e : {&sl.tr_in0, &sl.tr_c0, &sl.tr_c1}) IKB_CUDA(cudaEventCreate(e));
    }
    return IKB_OK;
}

void ikb_queue_free(ikb_queue *q) {
    if (!q) return;
    DeviceGuard g(q->p->device);
    queue_flush(q);
    queue_finish_carry(q);
    for (cudaStream_t s : {q->s_in, q->s_comp, q->s_out})
        if (s) {
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        }
    if (q->ev_user) cudaEventDestroy(q->ev_user);
    if (q->ev_comp) cudaEventDestroy(q->ev_comp);
    if (q->tr_ref) cudaEventDestroy(q->tr_ref);
    for (auto &sl : q->slots) {
        for (cudaEvent_t e : {sl.ev_in, sl.ev_done, sl.tr_in0, sl.tr_c0, sl.tr_c1})
            if (e) cudaEventDestroy(e);
        cudaFree(sl.st64.q0); cudaFree(sl.st64.targets); cudaFree(sl.st64.q); cudaFree(sl.st64.resid); cudaFree(sl.st64.compact);
        cudaFree(sl.st32.q0); cudaFree(sl.st32.targets); cudaFree(sl.st32.q); cudaFree(sl.st32.resid); cudaFree(sl.st32.compact);
        cudaFree(sl.success); cudaFree(sl.iters);
    }
    for (auto &c : q->cscratch) {
        cudaFree(c.list); cudaFree(c.iters); cudaFree(c.counters);
    }
    delete q;
}

int64_t ikb_queue_submit(ikb_queue *q, int dtype, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io, void *in_stream) {
    if (!q) return -fail(IKB_ERR_INVALID_ARG, "null queue");
    int rc = check_solve_args(q->p, dtype, prm, B, io);
    if (rc) return -rc;
    DeviceGuard g(q->p->device);
    if (io->targets_format != IKB_TARGETS_SE3) return -fail(IKB_ERR_INVALID_ARG, "compact targets are a wire format of the HOST entry points");
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
    const int rc = queue_flush(q);
    return rc ? rc : queue_finish_carry(q);   // an explicit flush leaves nothing behind: carried stragglers get their TAIL
}

int ikb_queue_wait(ikb_queue *q, int64_t ticket) {
    if (!q || ticket < 0 || ticket >= q->next) return fail(IKB_ERR_INVALID_ARG, "unknown queue ticket");
    ikb_queue::Slot &sl = q->slots[ticket % q->depth];
    if (sl.ticket != ticket) return IKB_OK;  // its slot has been reused: it left the pipeline long ago
    DeviceGuard g(q->p->device);
    int rc;
    if (sl.pending && (rc = queue_flush(q))) return rc;
    if (sl.deferred && (rc = queue_finish_carry(q))) return rc;
    IKB_CUDA(cudaEventSynchronize(sl.ev_done));
    if (q->trace && sl.host && sl.busy) {
        float a = 0, b = 0, c = 0, d = 0, e = 0;
        cudaEventElapsedTime(&a, q->tr_ref, sl.tr_in0);
        cudaEventElapsedTime(&b, q->tr_ref, sl.ev_in);
        cudaEventElapsedTime(&c, q->tr_ref, sl.tr_c0);
        cudaEventElapsedTime(&d, q->tr_ref, sl.tr_c1);
        cudaEventElapsedTime(&e, q->tr_ref, sl.ev_done);
        const double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - q->tr_host_ref).count();
        std::fprintf(stderr, "[ikb queue] ticket %lld: submit %.3f | copy-in %.3f-%.3f | solve %.3f-%.3f | results on host %.3f | wait returns %.3f ms\n",
                     (long long)ticket, sl.tr_submit_ms, a, b, c, d, e, now);
    }
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
    if (sl.deferred && (rc = queue_finish_carry(q))) return rc;
    IKB_CUDA(cudaStreamWaitEvent((cudaStream_t)cuda_stream, sl.ev_done, 0));
    return IKB_OK;
}

int ikb_queue_drain(ikb_queue *q) {
    if (!q) return fail(IKB_ERR_INVALID_ARG, "null queue");
    DeviceGuard g(q->p->device);
    int rc = queue_flush(q);
    if (rc) return rc;
    if ((rc = queue_finish_carry(q))) return rc;
    for (cudaStream_t s : {q->s_in, q->s_comp, q->s_out}) IKB_CUDA(cudaStreamSynchronize(s));
    for (auto &sl : q->slots) sl.busy = false;
    return IKB_OK;
}

}  // extern "C"
