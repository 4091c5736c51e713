// Several GPUs behind one handle (include/ikb200.h, ikb_multi_*): one finalized copy of the problem and one pipelined
// queue per device; a host batch is cut into contiguous slices, every slice is staged, solved and read back
// on its own device's streams, all devices concurrently under one host thread.  No collective: the results land in the
// caller's host arrays; the optional device-side gather is a set of peer copies.
#include <vector>

#include "capi_internal.hpp"

using namespace ikb;
using namespace ikb::capi;

struct ikb_multi {
    std::vector<int> devices;
    std::vector<ikb_problem *> problems;
    std::vector<ikb_queue *> queues;
    int64_t next = 0;
};

namespace {
// the sub-batch [b0, b0 + n) of a host batch: same views, moved base pointers
ikb_batch_io slice_io(const ikb_batch_io &io, int dtype, int64_t b0) {
    const size_t sz = dtype == IKB_F64 ? sizeof(double) : sizeof(float);
    ikb_batch_io s = io;
    auto off = [&](const void *base, int64_t bs) { return base ? (const void *)((const char *)base + (size_t)(b0 * bs) * sz) : nullptr; };
    s.q0 = off(io.q0, io.q0_batch_stride);
    s.targets = off(io.targets, io.targets_batch_stride);
    s.q = (void *)off(io.q, io.q_batch_stride);
    s.success = io.success ? io.success + b0 : nullptr;
    s.iters = io.iters ? io.iters + b0 : nullptr;
    s.resid = io.resid ? (void *)((char *)io.resid + (size_t)b0 * sz) : nullptr;
    return s;
}
}  // namespace

extern "C" {

int ikb_multi_create(const ikb_problem *p, const int *devices, int ndevices, int depth, int merge, ikb_multi **out) {
    if (!p || !devices || !out || ndevices < 1 || ndevices > 64) return fail(IKB_ERR_INVALID_ARG, "bad device list");
    ikb_multi *m = new ikb_multi;
    *out = m;   // the caller frees it also when creation fails half-way
    for (int i = 0; i < ndevices; ++i) {
        ikb_problem *c = new ikb_problem;
        c->hp = p->hp;   // problem.hpp:183: the problem owns its model by value, so a copy is self-contained
        m->problems.push_back(c);
        m->devices.push_back(devices[i]);
        int rc = ikb_problem_finalize(c, devices[i]);
        if (rc) return rc;
        ikb_queue *q = nullptr;
        rc = ikb_queue_create(c, depth, merge, &q);
        m->queues.push_back(q);
        if (rc) return rc;
    }
    return IKB_OK;
}

void ikb_multi_free(ikb_multi *m) {
    if (!m) return;
    for (ikb_queue *q : m->queues) ikb_queue_free(q);
    for (ikb_problem *p : m->problems) ikb_problem_free(p);
    delete m;
}

int ikb_multi_device_count(const ikb_multi *m) { return m ? (int)m->devices.size() : -IKB_ERR_INVALID_ARG; }

ikb_problem *ikb_multi_problem(ikb_multi *m, int index) {
    return (m && index >= 0 && index < (int)m->problems.size()) ? m->problems[index] : nullptr;
}

int64_t ikb_multi_submit_host(ikb_multi *m, int dtype, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io) {
    if (!m || !io || !prm || B < 0) return -fail(IKB_ERR_INVALID_ARG, "null argument");
    if (m->queues.size() != m->devices.size()) return -fail(IKB_ERR_NOT_FINALIZED, "ikb_multi_create did not complete");
    if (dtype != IKB_F64 && dtype != IKB_F32) return -fail(IKB_ERR_INVALID_ARG, "dtype must be IKB_F64 or IKB_F32");
    const int64_t G = (int64_t)m->devices.size();
    for (int64_t r = 0; r < G; ++r) {
        const int64_t b0 = r * B / G, b1 = (r + 1) * B / G;   // contiguous slices [r B / G, (r + 1) B / G)
        ikb_batch_io sub = slice_io(*io, dtype, b0);
        // a broadcast view (batch_stride 0) is the same for every slice; an SoA view keeps its row pitch (elem_stride)
        const int64_t t = ikb_queue_submit_host(m->queues[r], dtype, prm, b1 - b0, &sub);
        if (t < 0) return t;
        if (t != m->next) return -fail(IKB_ERR_INVALID_ARG, "internal: device queues out of step");
    }
    return m->next++;
}

int ikb_multi_wait(ikb_multi *m, int64_t ticket) {
    if (!m) return fail(IKB_ERR_INVALID_ARG, "null handle");
    // launch every device's open group first, then block on them one after the other: they run concurrently
    for (ikb_queue *q : m->queues) {
        int rc = ikb_queue_flush(q);
        if (rc) return rc;
    }
    for (ikb_queue *q : m->queues) {
        int rc = ikb_queue_wait(q, ticket);
        if (rc) return rc;
    }
    return IKB_OK;
}

int ikb_multi_drain(ikb_multi *m) {
    if (!m) return fail(IKB_ERR_INVALID_ARG, "null handle");
    for (ikb_queue *q : m->queues) {
        int rc = ikb_queue_flush(q);
        if (rc) return rc;
    }
    for (ikb_queue *q : m->queues) {
        int rc = ikb_queue_drain(q);
        if (rc) return rc;
    }
    return IKB_OK;
}

int ikb_multi_dls_solve_batch_host(ikb_multi *m, int dtype, const ikb_dls_params *prm, int64_t B, const ikb_batch_io *io) {
    const int64_t t = ikb_multi_submit_host(m, dtype, prm, B, io);
    if (t < 0) return (int)-t;
    return ikb_multi_wait(m, t);
}

int ikb_multi_gather_device(ikb_multi *m, const void *const *src, int count, int64_t B, int elem_bytes, int dst_device, void *dst) {
    if (!m || !src || !dst || count < 1 || B < 0 || elem_bytes < 1) return fail(IKB_ERR_INVALID_ARG, "bad gather arguments");
    const int64_t G = (int64_t)m->devices.size();
    for (int64_t r = 0; r < G; ++r) {
        const int64_t b0 = r * B / G, b1 = (r + 1) * B / G, n = b1 - b0;
        if (n <= 0) continue;
        ikb_problem *p = m->problems[r];
        DeviceGuard g(p->device);
        for (int k = 0; k < count; ++k)   // row k of the slice -> columns [b0, b1) of row k of the destination
            IKB_CUDA(cudaMemcpyPeerAsync((char *)dst + ((size_t)k * B + b0) * elem_bytes, dst_device,
                                         (const char *)src[r] + (size_t)k * n * elem_bytes, p->device, (size_t)n * elem_bytes, p->stream));
    }
    for (int64_t r = 0; r < G; ++r) {
        DeviceGuard g(m->problems[r]->device);
        IKB_CUDA(cudaStreamSynchronize(m->problems[r]->stream));
    }
    return IKB_OK;
}

}  // extern "C"
