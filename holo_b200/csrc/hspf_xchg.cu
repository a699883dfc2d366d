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
// copies that slot into slot r of buffer k of every peer with the copy engines and, behind
// each copy on the same stream, a 4-byte sequence number into the peer's `arrived` flag.
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

constexpr uint32_t kSeqTable = 1u << 16;   // device table of the values 1..65536 (flag copy sources)

typedef CUresult (*wait_value_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

}  // namespace

struct hspf_xchg {
    int device = 0;
    uint32_t rank = 0, world = 0, n_buffers = 0;
    size_t slot_bytes = 0, data_bytes = 0, total_bytes = 0;
    uint8_t *local = nullptr;                 // own allocation
    std::vector<uint8_t *> peer;              // [world] mapped base of every rank (own = local)
    uint32_t *seq_table = nullptr;            // device: 1..kSeqTable
    std::vector<uint32_t> seq;                // [n_buffers] pushes done on buffer k
    std::vector<uint32_t> released;           // [n_buffers] releases done on buffer k
    cudaStream_t compute = nullptr;           // the engine's stream (hspf_stream)
    cudaStream_t push_stream = nullptr, consume_stream = nullptr;
    std::vector<cudaEvent_t> kernel_done, push_done, consumed;   // [n_buffers]
    wait_value_fn wait_value = nullptr;
    std::string err;

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
    if (cudaMalloc(&x->seq_table, kSeqTable * sizeof(uint32_t)) != cudaSuccess) return bail(HSPF_E_NOMEM);
    {
        std::vector<uint32_t> t(kSeqTable);
        for (uint32_t i = 0; i < kSeqTable; ++i) t[i] = i + 1;
        if (cudaMemcpy(x->seq_table, t.data(), t.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return bail(HSPF_E_CUDA);
    }
    if (cudaStreamCreateWithFlags(&x->push_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(HSPF_E_CUDA);
    if (cudaStreamCreateWithFlags(&x->consume_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(HSPF_E_CUDA);
    x->kernel_done.resize(n_buffers); x->push_done.resize(n_buffers); x->consumed.resize(n_buffers);
    for (uint32_t k = 0; k < n_buffers; ++k) {
        if (cudaEventCreateWithFlags(&x->kernel_done[k], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&x->push_done[k], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&x->consumed[k], cudaEventDisableTiming) != cudaSuccess)
            return bail(HSPF_E_CUDA);
    }
    x->seq.assign(n_buffers, 0);
    x->released.assign(n_buffers, 0);
    x->peer.assign(world, nullptr);
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
    if (x->seq[buffer]) XCK(cudaStreamWaitEvent(x->compute, x->push_done[buffer], 0));
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
    const uint32_t s = ++x->seq[k];
    if (s > kSeqTable) return xfail(x, HSPF_E_UNSUPPORTED, "more than 65536 pushes on one buffer");
    XCK(cudaEventRecord(x->kernel_done[k], x->compute));
    XCK(cudaStreamWaitEvent(x->push_stream, x->kernel_done[k], 0));
    for (uint32_t d = 1; d < x->world; ++d) {
        const uint32_t r = (me + d) % x->world;      // every rank starts with a different peer
        if (s > 1) {   // the peer must have released what it read from the previous push
            CUresult cr = x->wait_value(reinterpret_cast<CUstream>(x->push_stream),
                                        reinterpret_cast<CUdeviceptr>(x->acked(x->local, k, r)), s - 1,
                                        CU_STREAM_WAIT_VALUE_GEQ);
            if (cr != CUDA_SUCCESS) return xfail(x, HSPF_E_CUDA, "cuStreamWaitValue32(acked) failed");
        }
        XCK(cudaMemcpyAsync(x->slot(x->peer[r], k, me), x->slot(x->local, k, me), x->slot_bytes,
                            cudaMemcpyDeviceToDevice, x->push_stream));
        XCK(cudaMemcpyAsync(x->arrived(x->peer[r], k, me), x->seq_table + (s - 1), sizeof(uint32_t),
                            cudaMemcpyDeviceToDevice, x->push_stream));
    }
    XCK(cudaEventRecord(x->push_done[k], x->push_stream));
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
                                    reinterpret_cast<CUdeviceptr>(x->arrived(x->local, k, r)), s,
                                    CU_STREAM_WAIT_VALUE_GEQ);
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
        XCK(cudaMemcpyAsync(x->acked(x->peer[r], k, x->rank), x->seq_table + (s - 1), sizeof(uint32_t),
                            cudaMemcpyDeviceToDevice, x->consume_stream));
    }
    XCK(cudaEventRecord(x->consumed[k], x->consume_stream));
    return HSPF_OK;
}

void *hspf_xchg_consumer_stream(hspf_xchg *x) { return x ? x->consume_stream : nullptr; }

/* Block the host until every push and every consumer operation enqueued so far is done. */
int hspf_xchg_sync(hspf_xchg *x) {
    if (!x) return HSPF_E_INVAL;
    XCK(cudaStreamSynchronize(x->push_stream));
    XCK(cudaStreamSynchronize(x->consume_stream));
    return HSPF_OK;
}

int hspf_xchg_destroy(hspf_xchg *x) {
    if (!x) return HSPF_OK;
    cudaSetDevice(x->device);
    if (x->push_stream) cudaStreamSynchronize(x->push_stream);
    if (x->consume_stream) cudaStreamSynchronize(x->consume_stream);
    for (uint32_t r = 0; r < x->peer.size(); ++r)
        if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
    for (auto e : x->kernel_done) if (e) cudaEventDestroy(e);
    for (auto e : x->push_done) if (e) cudaEventDestroy(e);
    for (auto e : x->consumed) if (e) cudaEventDestroy(e);
    if (x->push_stream) cudaStreamDestroy(x->push_stream);
    if (x->consume_stream) cudaStreamDestroy(x->consume_stream);
    if (x->seq_table) cudaFree(x->seq_table);
    if (x->local) cudaFree(x->local);
    delete x;
    return HSPF_OK;
}

}  // extern "C"
