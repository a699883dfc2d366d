// Multi-GPU result exchange over NVLink peer memory (include/holo_spf.h, "hspf_xchg_*").
//
// One process per GPU.  Every rank owns one device allocation
//
//     data    [n_buffers][world][slot_bytes]     slot r of buffer k = rank r's planes of a step
//     arrived [n_buffers][world] u32             sequence number of the last push that landed
//     acked   [n_buffers][world] u32             sequence number the owner of slot r has seen
//                                                consumed by this rank's peer
//
// exported with cudaIpcGetMemHandle and opened by every peer.  The batch kernel of rank r
// writes its result planes straight into slot r of its own buffer k; hspf_xchg_push(k) then
// copies that slot (or its first push_bytes) into slot r of buffer k of every peer with the copy
// engines — one stream per peer, so the copies to different peers run on different engines —
// and, behind each copy on the same stream, a 4-byte sequence number into the peer's `arrived`
// flag (written into a local staging word with a 32-bit memset, then copied: any 32-bit value,
// no limit on the number of pushes).
// No SM is involved, so the persistent batch kernel of the next step runs at full width
// while the planes travel (an NCCL all-gather needs CTAs, which that kernel does not leave).
// Consumers wait with stream memory operations (cuStreamWaitValue32), which need no SM
// either, and hand a buffer back with hspf_xchg_release (acks, same mechanism), so a peer
// cannot overwrite a slot that is still being read.
//
// This is the exchange step SURVEY.md §8e / BASELINE north_star name ("all-gather of
// per-partition SPT results over NVLink"); the reference has no counterpart (single process).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/holo_spf.h"

namespace {

typedef CUresult (*wait_value_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*memset32_fn)(CUdeviceptr, unsigned int, size_t, CUstream);

}  // namespace

struct hspf_xchg {
    int device = 0;
    uint32_t rank = 0, world = 0, n_buffers = 0;
    size_t slot_bytes = 0, data_bytes = 0, total_bytes = 0;
    uint8_t *local = nullptr;                 // own allocation
    std::vector<uint8_t *> peer;              // [world] mapped base of every rank (own = local)
    std::vector<bool> peer_is_ipc;            // opened with cudaIpcOpenMemHandle (to be closed)
    uint32_t *stage = nullptr;                // device: [2][n_buffers][world] staging words of the flag copies
    std::vector<uint32_t> seq;                // [n_buffers] pushes done on buffer k
    std::vector<uint32_t> released;           // [n_buffers] releases done on buffer k
    size_t push_bytes = 0;                    // bytes of the own slot that travel (<= slot_bytes)
    cudaStream_t compute = nullptr;           // the engine's stream (hspf_stream)
    std::vector<cudaStream_t> push_streams;   // [world] one per peer (own: unused)
    cudaStream_t consume_stream = nullptr;
    std::vector<cudaEvent_t> kernel_done, consumed;              // [n_buffers]
    std::vector<cudaEvent_t> push_done;                           // [n_buffers * world]
    wait_value_fn wait_value = nullptr;
    memset32_fn memset32 = nullptr;
    unsigned int wait_flags = CU_STREAM_WAIT_VALUE_GEQ;
    std::string err;

    uint32_t *stage_arrived(uint32_t k, uint32_t r) const { return stage + (size_t)k * world + r; }
    uint32_t *stage_acked(uint32_t k, uint32_t r) const { return stage + (size_t)(n_buffers + k) * world + r; }

    uint32_t *arrived(uint8_t *base, uint32_t k, uint32_t r) const {
        return reinterpret_cast<uint32_t *>(base + data_bytes) + (size_t)k * world + r;
    }
    uint32_t *acked(uint8_t *base, uint32_t k, uint32_t r) const {
        return reinterpret_cast<uint32_t *>(base + data_bytes) + (size_t)(n_buffers + k) * world + r;
    }
    uint8_t *slot(uint8_t *base, uint32_t k, uint32_t r) const {
        return base + ((size_t)k * world + r) * slot_bytes;
    }
};

namespace {

int xfail(hspf_xchg *x, int code, const std::string &m) {
    if (x) x->err = m;
    return code;
}

#define XCK(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return xfail(x, HSPF_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

}  // namespace

extern "C" {

int hspf_xchg_create(hspf_ctx *ctx, int device, uint32_t rank, uint32_t world, size_t slot_bytes,
                     uint32_t n_buffers, hspf_xchg **out, uint8_t handle[HSPF_IPC_HANDLE_BYTES]) {
    if (!ctx || !out || !handle || world < 2 || rank >= world || !slot_bytes || !n_buffers) return HSPF_E_INVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == HSPF_IPC_HANDLE_BYTES, "IPC handle size");
    *out = nullptr;
    hspf_xchg *x = new (std::nothrow) hspf_xchg();
    if (!x) return HSPF_E_NOMEM;
    x->device = device; x->rank = rank; x->world = world; x->n_buffers = n_buffers;
    x->slot_bytes = (slot_bytes + 255) / 256 * 256;
    x->data_bytes = x->slot_bytes * world * n_buffers;
    x->total_bytes = x->data_bytes + (size_t)2 * n_buffers * world * sizeof(uint32_t);
    x->compute = static_cast<cudaStream_t>(hspf_stream(ctx));
    auto bail = [&](int rc) { hspf_xchg_destroy(x); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(HSPF_E_CUDA);
    if (cudaMalloc(&x->local, x->total_bytes) != cudaSuccess) return bail(HSPF_E_NOMEM);
    if (cudaMemset(x->local, 0, x->total_bytes) != cudaSuccess) return bail(HSPF_E_CUDA);
    x->push_bytes = x->slot_bytes;
    if (cudaMalloc(&x->stage, (size_t)2 * n_buffers * world * sizeof(uint32_t)) != cudaSuccess) return bail(HSPF_E_NOMEM);
    x->push_streams.assign(world, nullptr);
    for (uint32_t r = 0; r < world; ++r)
        if (r != rank && cudaStreamCreateWithFlags(&x->push_streams[r], cudaStreamNonBlocking) != cudaSuccess)
            return bail(HSPF_E_CUDA);
    if (cudaStreamCreateWithFlags(&x->consume_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(HSPF_E_CUDA);
    x->kernel_done.resize(n_buffers); x->consumed.resize(n_buffers);
    x->push_done.assign((size_t)n_buffers * world, nullptr);
    for (uint32_t k = 0; k < n_buffers; ++k) {
        if (cudaEventCreateWithFlags(&x->kernel_done[k], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&x->consumed[k], cudaEventDisableTiming) != cudaSuccess)
            return bail(HSPF_E_CUDA);
        for (uint32_t r = 0; r < world; ++r)
            if (r != rank && cudaEventCreateWithFlags(&x->push_done[(size_t)k * world + r], cudaEventDisableTiming) != cudaSuccess)
                return bail(HSPF_E_CUDA);
    }
    x->seq.assign(n_buffers, 0);
    x->released.assign(n_buffers, 0);
    x->peer.assign(world, nullptr);
    x->peer_is_ipc.assign(world, false);
    x->peer[rank] = x->local;
    // stream memory operations through the runtime's driver entry point (libcuda is not linked,
    // so the library still loads on a machine without a driver)
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess) {
        hspf_xchg_destroy(x);
        return HSPF_E_UNSUPPORTED;
    }
    x->wait_value = reinterpret_cast<wait_value_fn>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuMemsetD32Async", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess) {
        hspf_xchg_destroy(x);
        return HSPF_E_UNSUPPORTED;
    }
    x->memset32 = reinterpret_cast<memset32_fn>(fn);
    // A consumer sees a peer's data only through the polled flag.  The design relies on a peer
    // copy being visible at the destination before the next copy of the same stream (the flag)
    // lands; where the device can flush remote writes the waits ask for it as well.
    int can_flush = 0;
    if (cudaDeviceGetAttribute(&can_flush, cudaDevAttrCanFlushRemoteWrites, device) == cudaSuccess && can_flush)
        x->wait_flags |= CU_STREAM_WAIT_VALUE_FLUSH;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, x->local) != cudaSuccess) return bail(HSPF_E_CUDA);
    std::memcpy(handle, &h, sizeof(h));
    *out = x;
    return HSPF_OK;
}

int hspf_xchg_attach(hspf_xchg *x, uint32_t peer_rank, const uint8_t handle[HSPF_IPC_HANDLE_BYTES]) {
    if (!x || !handle || peer_rank >= x->world) return HSPF_E_INVAL;
    if (peer_rank == x->rank) return HSPF_OK;
    XCK(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    XCK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer[peer_rank] = static_cast<uint8_t *>(p);
    x->peer_is_ipc[peer_rank] = true;
    return HSPF_OK;
}

/* Attach a peer whose allocation is addressable in THIS process (hspf_xchg_base of an exchange
 * created by the same process, e.g. two contexts on one device or on two devices with peer
 * access enabled): tests and single-process deployments. */
int hspf_xchg_attach_ptr(hspf_xchg *x, uint32_t peer_rank, void *peer_base) {
    if (!x || !peer_base || peer_rank >= x->world) return HSPF_E_INVAL;
    if (peer_rank == x->rank) return HSPF_OK;
    x->peer[peer_rank] = static_cast<uint8_t *>(peer_base);
    x->peer_is_ipc[peer_rank] = false;
    return HSPF_OK;
}

void *hspf_xchg_base(hspf_xchg *x) { return x ? x->local : nullptr; }

/* Only the first `nbytes` of the own slot travel in hspf_xchg_push (e.g. the planes the caller's
 * Vertex keeps, laid out first).  0 or more than the slot = the whole slot. */
int hspf_xchg_set_push_bytes(hspf_xchg *x, size_t nbytes) {
    if (!x) return HSPF_E_INVAL;
    x->push_bytes = (nbytes == 0 || nbytes > x->slot_bytes) ? x->slot_bytes : nbytes;
    return HSPF_OK;
}

void *hspf_xchg_slot(hspf_xchg *x, uint32_t buffer, uint32_t slot) {
    if (!x || buffer >= x->n_buffers || slot >= x->world) return nullptr;
    return x->slot(x->local, buffer, slot);
}

size_t hspf_xchg_slot_bytes(const hspf_xchg *x) { return x ? x->slot_bytes : 0; }

const char *hspf_xchg_last_error(const hspf_xchg *x) { return x ? x->err.c_str() : "null exchange"; }

/* Before the batch that writes own slot of `buffer` is enqueued on the engine's stream: that
 * stream waits until the previous push of this buffer has left the device and the local
 * consumer has released the buffer. */
int hspf_xchg_acquire(hspf_xchg *x, uint32_t buffer) {
    if (!x || buffer >= x->n_buffers) return HSPF_E_INVAL;
    if (x->seq[buffer])
        for (uint32_t r = 0; r < x->world; ++r)
            if (r != x->rank) XCK(cudaStreamWaitEvent(x->compute, x->push_done[(size_t)buffer * x->world + r], 0));
    // ... and until the local consumer has released it (it reads the own slot too)
    if (x->released[buffer]) XCK(cudaStreamWaitEvent(x->compute, x->consumed[buffer], 0));
    return HSPF_OK;
}

/* After the batch: copy own slot of `buffer` to every peer (copy engines), then the flags. */
int hspf_xchg_push(hspf_xchg *x, uint32_t buffer) {
    if (!x || buffer >= x->n_buffers) return HSPF_E_INVAL;
    for (uint32_t r = 0; r < x->world; ++r)
        if (!x->peer[r]) return xfail(x, HSPF_E_INVAL, "hspf_xchg_push before every peer is attached");
    const uint32_t k = buffer, me = x->rank;
    const uint32_t s = x->seq[k] + 1;
    XCK(cudaEventRecord(x->kernel_done[k], x->compute));
    for (uint32_t d = 1; d < x->world; ++d) {
        const uint32_t r = (me + d) % x->world;      // every rank starts with a different peer
        cudaStream_t ps = x->push_streams[r];
        XCK(cudaStreamWaitEvent(ps, x->kernel_done[k], 0));
        if (s > 1) {   // the peer must have released what it read from the previous push
            CUresult cr = x->wait_value(reinterpret_cast<CUstream>(ps),
                                        reinterpret_cast<CUdeviceptr>(x->acked(x->local, k, r)), s - 1,
                                        CU_STREAM_WAIT_VALUE_GEQ);
            if (cr != CUDA_SUCCESS) return xfail(x, HSPF_E_CUDA, "cuStreamWaitValue32(acked) failed");
        }
        XCK(cudaMemcpyAsync(x->slot(x->peer[r], k, me), x->slot(x->local, k, me), x->push_bytes,
                            cudaMemcpyDeviceToDevice, ps));
        if (x->memset32(reinterpret_cast<CUdeviceptr>(x->stage_arrived(k, r)), s, 1, reinterpret_cast<CUstream>(ps)) != CUDA_SUCCESS)
            return xfail(x, HSPF_E_CUDA, "cuMemsetD32Async(flag) failed");
        XCK(cudaMemcpyAsync(x->arrived(x->peer[r], k, me), x->stage_arrived(k, r), sizeof(uint32_t),
                            cudaMemcpyDeviceToDevice, ps));
        XCK(cudaEventRecord(x->push_done[(size_t)k * x->world + r], ps));
    }
    x->seq[k] = s;
    return HSPF_OK;
}

/* ---- fused exchange: the batch kernel itself stores into the peers' slots -----------------
 * hspf_xchg_acquire_direct replaces acquire (the engine's stream also waits until every peer has
 * released what it read from this rank's slot of `buffer`: the kernel is about to overwrite the
 * peers' copies), hspf_xchg_peer_deltas gives the address differences for hspf_ctx_set_peer_slots,
 * hspf_xchg_publish replaces push (flags only, behind the kernel). */
int hspf_xchg_acquire_direct(hspf_xchg *x, uint32_t buffer) {
    if (!x || buffer >= x->n_buffers) return HSPF_E_INVAL;
    int rc = hspf_xchg_acquire(x, buffer);
    if (rc) return rc;
    const uint32_t k = buffer, s = x->seq[k] + 1;
    if (s > 1)
        for (uint32_t r = 0; r < x->world; ++r) {
            if (r == x->rank) continue;
            CUresult cr = x->wait_value(reinterpret_cast<CUstream>(x->compute),
                                        reinterpret_cast<CUdeviceptr>(x->acked(x->local, k, r)), s - 1,
                                        CU_STREAM_WAIT_VALUE_GEQ);
            if (cr != CUDA_SUCCESS) return xfail(x, HSPF_E_CUDA, "cuStreamWaitValue32(acked) failed");
        }
    return HSPF_OK;
}

int hspf_xchg_peer_deltas(hspf_xchg *x, uint32_t buffer, int64_t *deltas, uint32_t *n_peers) {
    if (!x || buffer >= x->n_buffers || !deltas || !n_peers) return HSPF_E_INVAL;
    uint32_t n = 0;
    for (uint32_t d = 1; d < x->world; ++d) {
        const uint32_t r = (x->rank + d) % x->world;
        if (!x->peer[r]) return xfail(x, HSPF_E_INVAL, "hspf_xchg_peer_deltas before every peer is attached");
        deltas[n++] = (int64_t)(x->slot(x->peer[r], buffer, x->rank) - x->slot(x->local, buffer, x->rank));
    }
    *n_peers = n;
    return HSPF_OK;
}

int hspf_xchg_publish(hspf_xchg *x, uint32_t buffer) {
    if (!x || buffer >= x->n_buffers) return HSPF_E_INVAL;
    for (uint32_t r = 0; r < x->world; ++r)
        if (!x->peer[r]) return xfail(x, HSPF_E_INVAL, "hspf_xchg_publish before every peer is attached");
    const uint32_t k = buffer, me = x->rank;
    const uint32_t s = x->seq[k] + 1;
    XCK(cudaEventRecord(x->kernel_done[k], x->compute));
    for (uint32_t d = 1; d < x->world; ++d) {
        const uint32_t r = (me + d) % x->world;
        cudaStream_t ps = x->push_streams[r];
        XCK(cudaStreamWaitEvent(ps, x->kernel_done[k], 0));      // the kernel's peer stores are complete
        if (x->memset32(reinterpret_cast<CUdeviceptr>(x->stage_arrived(k, r)), s, 1, reinterpret_cast<CUstream>(ps)) != CUDA_SUCCESS)
            return xfail(x, HSPF_E_CUDA, "cuMemsetD32Async(flag) failed");
        XCK(cudaMemcpyAsync(x->arrived(x->peer[r], k, me), x->stage_arrived(k, r), sizeof(uint32_t),
                            cudaMemcpyDeviceToDevice, ps));
        XCK(cudaEventRecord(x->push_done[(size_t)k * x->world + r], ps));
    }
    x->seq[k] = s;
    return HSPF_OK;
}

/* The consumer stream of the exchange waits until every peer's slot of `buffer` carries the
 * data of this rank's latest push number (ranks step in lockstep), without using an SM. */
int hspf_xchg_wait(hspf_xchg *x, uint32_t buffer) {
    if (!x || buffer >= x->n_buffers) return HSPF_E_INVAL;
    const uint32_t k = buffer, s = x->seq[k];
    if (!s) return xfail(x, HSPF_E_INVAL, "hspf_xchg_wait before the first push");
    for (uint32_t r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        CUresult cr = x->wait_value(reinterpret_cast<CUstream>(x->consume_stream),
                                    reinterpret_cast<CUdeviceptr>(x->arrived(x->local, k, r)), s, x->wait_flags);
        if (cr != CUDA_SUCCESS) return xfail(x, HSPF_E_CUDA, "cuStreamWaitValue32(arrived) failed");
    }
    // own slot: the batch that produced it
    XCK(cudaStreamWaitEvent(x->consume_stream, x->kernel_done[k], 0));
    return HSPF_OK;
}

/* Work enqueued on the consumer stream so far has read `buffer`: tell every peer. */
int hspf_xchg_release(hspf_xchg *x, uint32_t buffer) {
    if (!x || buffer >= x->n_buffers) return HSPF_E_INVAL;
    const uint32_t k = buffer, s = x->seq[k];
    if (!s || x->released[k] == s) return HSPF_OK;
    x->released[k] = s;
    for (uint32_t d = 1; d < x->world; ++d) {
        const uint32_t r = (x->rank + d) % x->world;
        if (x->memset32(reinterpret_cast<CUdeviceptr>(x->stage_acked(k, r)), s, 1,
                        reinterpret_cast<CUstream>(x->consume_stream)) != CUDA_SUCCESS)
            return xfail(x, HSPF_E_CUDA, "cuMemsetD32Async(ack) failed");
        XCK(cudaMemcpyAsync(x->acked(x->peer[r], k, x->rank), x->stage_acked(k, r), sizeof(uint32_t),
                            cudaMemcpyDeviceToDevice, x->consume_stream));
    }
    XCK(cudaEventRecord(x->consumed[k], x->consume_stream));
    return HSPF_OK;
}

void *hspf_xchg_consumer_stream(hspf_xchg *x) { return x ? x->consume_stream : nullptr; }

/* Block the host until every push and every consumer operation enqueued so far is done. */
int hspf_xchg_sync(hspf_xchg *x) {
    if (!x) return HSPF_E_INVAL;
    for (cudaStream_t ps : x->push_streams)
        if (ps) XCK(cudaStreamSynchronize(ps));
    XCK(cudaStreamSynchronize(x->consume_stream));
    return HSPF_OK;
}

int hspf_xchg_destroy(hspf_xchg *x) {
    if (!x) return HSPF_OK;
    cudaSetDevice(x->device);
    for (cudaStream_t ps : x->push_streams) if (ps) cudaStreamSynchronize(ps);
    if (x->consume_stream) cudaStreamSynchronize(x->consume_stream);
    for (uint32_t r = 0; r < x->peer.size(); ++r)
        if (r != x->rank && x->peer[r] && x->peer_is_ipc[r]) cudaIpcCloseMemHandle(x->peer[r]);
    for (auto e : x->kernel_done) if (e) cudaEventDestroy(e);
    for (auto e : x->push_done) if (e) cudaEventDestroy(e);
    for (auto e : x->consumed) if (e) cudaEventDestroy(e);
    for (cudaStream_t ps : x->push_streams) if (ps) cudaStreamDestroy(ps);
    if (x->consume_stream) cudaStreamDestroy(x->consume_stream);
    if (x->stage) cudaFree(x->stage);
    if (x->local) cudaFree(x->local);
    delete x;
    return HSPF_OK;
}

}  // extern "C"
