// Device route stage for batches of OSPFv2 SPTs (include/holo_spf_lsdb.h, "batched intra-area
// route stage"): update_rib_intra_area (holo-ospf/src/route.rs:343-446) for every job of a batch.
//
// The SPT planes of a batch stay in HBM; one thread per (job, prefix) walks the prefix's advertisers
// (route_cells.h: route_cell_eval) and writes one 24-byte cell.  Prefix is the fast index: the
// advertiser lists and the cells are read / written coalesced, the plane values are gathers inside
// the job's own rows (200 KB at 10k vertices: L2 hits).  HBM traffic per job is the cells
// (24 B x prefixes) plus one pass over the three planes; the walk itself is a handful of integer
// instructions per advertiser, so the stage is bounded by the cell writes.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <new>
#include <vector>

#include "../../include/holo_spf_lsdb.h"
#include "route_cells.h"

namespace {

using hspf::RouteContrib;

template <class Planes, class D, class N>
__global__ void __launch_bounds__(256)
route_cells_kernel(uint32_t n_jobs, uint32_t P, uint32_t V, const uint32_t *__restrict__ off,
                   const RouteContrib *__restrict__ contribs, const D *__restrict__ dist,
                   const uint16_t *__restrict__ hops, const N *__restrict__ nh, const uint32_t *__restrict__ job_status,
                   hl_route_cell *__restrict__ cells, bool aligned16, uint32_t n_gather,
                   const uint32_t *__restrict__ gather_job, const uint32_t *__restrict__ gather_v,
                   uint64_t *__restrict__ gather_nh) {
    // a warp owns 32 consecutive cells = one contiguous 768-byte span of the output: the cells are staged in
    // shared memory and leave as 48 16-byte stores (full sectors) instead of 96 scattered 8-byte ones
    __shared__ __align__(16) uint64_t stage[8][96];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t total = (uint64_t)n_jobs * P;
    const uint64_t n_tiles = (total + 31) / 32;
    const uint64_t wstride = (uint64_t)gridDim.x * 8;
    for (uint64_t tile = (uint64_t)blockIdx.x * 8 + wib; tile < n_tiles; tile += wstride) {
        const uint64_t idx = tile * 32 + lane;
        hl_route_cell c;
        c.nh_mask = 0; c.lasthop_mask = 0; c.winner = 0xFFFFFFFFu; c.metric = 0; c.flags = 0; c._pad = 0;
        if (idx < total) {
            const uint32_t job = (uint32_t)(idx / P), p = (uint32_t)(idx - (uint64_t)job * P);
            if (!(job_status && job_status[job] != 0)) {      // planes of a refused job are undefined: empty cells
                const size_t base = (size_t)job * V;
                const Planes pl{dist + base, hops + base, nh + base};
                c = hspf::route_cell_eval(pl, contribs, off[p], off[p + 1]);
            }
        }
        const uint64_t w2 = (uint64_t)c.winner | ((uint64_t)c.metric << 32) | ((uint64_t)c.flags << 48);
        if (aligned16 && tile * 32 + 32 <= total) {
            uint64_t *s = stage[wib];
            s[lane * 3 + 0] = c.nh_mask; s[lane * 3 + 1] = c.lasthop_mask; s[lane * 3 + 2] = w2;
            __syncwarp();
            const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
            uint4 *o4 = reinterpret_cast<uint4 *>(cells + tile * 32);
            o4[lane] = s4[lane];
            if (lane < 16) o4[32 + lane] = s4[32 + lane];
            __syncwarp();
        } else if (idx < total) {
            uint64_t *o = reinterpret_cast<uint64_t *>(cells + idx);
            o[0] = c.nh_mask; o[1] = c.lasthop_mask; o[2] = w2;
        }
    }
    // the few plane values the host decode needs
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_gather; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t job = gather_job[g], v = gather_v[g];
        gather_nh[g] = (job < n_jobs && v < V) ? (uint64_t)nh[(size_t)job * V + v] : 0;
    }
}
static_assert(sizeof(hl_route_cell) == 24, "hl_route_cell layout");

template <class Planes, class D, class N>
int launch_cells(hspf_ctx *ctx, const hspf_ospfv2_rtable *rt, uint32_t n_jobs, const D *dist, const uint16_t *hops,
                 const N *nh, const uint32_t *status, hl_route_cell *cells, uint32_t n_gather,
                 const uint32_t *gather_job, const uint32_t *gather_v, uint64_t *gather_nh) {
    if (!ctx || !rt || !rt->d_blob || !dist || !hops || !nh || !cells) return HSPF_E_INVAL;
    if (n_gather && (!gather_job || !gather_v || !gather_nh)) return HSPF_E_INVAL;
    const uint32_t P = (uint32_t)rt->t.prefix.size();
    const uint64_t total = (uint64_t)n_jobs * P;
    if (total + n_gather == 0) return HSPF_OK;
    const int dev = hspf_ctx_device(ctx);
    if (rt->device != dev) return HSPF_E_INVAL;               // the table was uploaded to another device
    int sms = 0;
    if (cudaSetDevice(dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return HSPF_E_CUDA;
    // one resident wave (8 blocks of 256 per SM), warp-tile-stride beyond that
    const uint64_t want = std::max<uint64_t>((total + 255) / 256, 1);
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(want, (uint64_t)sms * 8);
    const bool aligned16 = (reinterpret_cast<uintptr_t>(cells) & 15u) == 0;
    cudaStream_t st = static_cast<cudaStream_t>(hspf_stream(ctx));
    route_cells_kernel<Planes, D, N><<<blocks, 256, 0, st>>>(n_jobs, P, rt->t.n_vertices, rt->d_off, rt->d_contribs, dist, hops,
                                                             nh, status, cells, aligned16, n_gather, gather_job, gather_v, gather_nh);
    if (cudaGetLastError() != cudaSuccess) return HSPF_E_CUDA;
    hspf_note_launches(ctx, 1);
    return HSPF_OK;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)) == cudaSuccess ? 0 : HSPF_E_NOMEM; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

}  // namespace

void hspf_rtable_release_device(hspf_ospfv2_rtable *rt) {
    if (rt && rt->d_blob) {
        cudaFree(rt->d_blob);
        rt->d_blob = nullptr; rt->d_off = nullptr; rt->d_contribs = nullptr;
    }
}

extern "C" {

int hspf_ospfv2_rtable_upload(hspf_ctx *ctx, hspf_ospfv2_rtable *rt) {
    if (!ctx || !rt) return HSPF_E_INVAL;
    hspf_rtable_release_device(rt);
    if (cudaSetDevice(hspf_ctx_device(ctx)) != cudaSuccess) return HSPF_E_CUDA;
    const size_t off_bytes = (rt->t.off.size() * sizeof(uint32_t) + 15) & ~(size_t)15;
    const size_t con_bytes = rt->t.contribs.size() * sizeof(RouteContrib);
    void *blob = nullptr;
    if (cudaMalloc(&blob, off_bytes + std::max<size_t>(con_bytes, 16)) != cudaSuccess) return HSPF_E_NOMEM;
    cudaStream_t st = static_cast<cudaStream_t>(hspf_stream(ctx));
    cudaError_t e = cudaMemcpyAsync(blob, rt->t.off.data(), rt->t.off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && con_bytes)
        e = cudaMemcpyAsync(static_cast<char *>(blob) + off_bytes, rt->t.contribs.data(), con_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);      // the host vectors may go away after the call
    if (e != cudaSuccess) { cudaFree(blob); return HSPF_E_CUDA; }
    rt->d_blob = blob;
    rt->device = hspf_ctx_device(ctx);
    rt->d_off = static_cast<const uint32_t *>(blob);
    rt->d_contribs = reinterpret_cast<const RouteContrib *>(static_cast<char *>(blob) + off_bytes);
    return HSPF_OK;
}

int hspf_ospfv2_routes_batch(hspf_ctx *ctx, const hspf_ospfv2_rtable *rt, uint32_t n_jobs, const hspf_result *pl,
                             hl_route_cell *cells, uint32_t n_gather, const uint32_t *gather_job,
                             const uint32_t *gather_v, uint64_t *gather_nh) {
    if (!pl || pl->nh_words != 1) return HSPF_E_INVAL;
    return launch_cells<hspf::PlanesWide, uint32_t, uint64_t>(ctx, rt, n_jobs, pl->dist, pl->hops, pl->nh_mask, pl->job_status,
                                                              cells, n_gather, gather_job, gather_v, gather_nh);
}

int hspf_ospfv2_routes_batch16(hspf_ctx *ctx, const hspf_ospfv2_rtable *rt, uint32_t n_jobs, const hspf_result16 *pl,
                               hl_route_cell *cells, uint32_t n_gather, const uint32_t *gather_job,
                               const uint32_t *gather_v, uint64_t *gather_nh) {
    if (!pl) return HSPF_E_INVAL;
    return launch_cells<hspf::PlanesNarrow, uint16_t, uint16_t>(ctx, rt, n_jobs, pl->dist, pl->hops, pl->nh_mask, pl->job_status,
                                                                cells, n_gather, gather_job, gather_v, gather_nh);
}

int hspf_ospfv2_run_area_batch(hspf_ctx *ctx, const hl_ospfv2_area *area, const uint32_t *root_router_ids, uint32_t n_roots,
                               hl_route_cell *cells, uint64_t cells_cap, uint32_t *n_prefixes, uint32_t *job_status,
                               uint32_t *gather_off, uint32_t *gather_v, uint64_t *gather_nh, uint32_t gather_cap,
                               double *device_ms) {
    if (!ctx || !area || (n_roots && !root_router_ids) || !n_prefixes || !job_status || !gather_off) return HSPF_E_INVAL;
    hspf_ospfv2_flat *flat = nullptr;
    hspf_ospfv2_rtable *rt = nullptr;
    hspf_graph *g = nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    struct Cleanup {
        hspf_ctx *ctx; hspf_ospfv2_flat *&flat; hspf_ospfv2_rtable *&rt; hspf_graph *&g; cudaEvent_t *ev;
        ~Cleanup() {
            if (g) hspf_graph_free(ctx, g);
            if (rt) hspf_ospfv2_rtable_free(rt);
            if (flat) hspf_ospfv2_flat_free(flat);
            for (int i = 0; i < 3; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
        }
    } cleanup{ctx, flat, rt, g, ev};
    try {
        int rc = hspf_ospfv2_flatten(area, &flat);
        if (rc) return rc;
        rc = hspf_ospfv2_rtable_create(flat, &rt);
        if (rc) return rc;
        const uint32_t P = hspf_ospfv2_rtable_prefixes(rt);
        *n_prefixes = P;
        if ((uint64_t)n_roots * P > cells_cap || (n_roots && P && !cells)) return HSPF_E_NOMEM;
        hspf_csr csr;
        rc = hspf_ospfv2_flat_csr(flat, &csr);
        if (rc) return rc;
        const uint32_t V = csr.n_vertices;
        const uint8_t *is_router = nullptr;
        hspf_ospfv2_flat_vertices(flat, nullptr, &is_router, nullptr);
        // roots, and per root the transit networks next to it (the only plane values the decode needs)
        std::vector<uint32_t> roots(n_roots), gj, gv;
        gather_off[0] = 0;
        for (uint32_t j = 0; j < n_roots; ++j) {
            const uint32_t r = hspf_ospfv2_flat_router_vertex(flat, root_router_ids[j]);
            if (r == 0xFFFFFFFFu) return HSPF_E_INVAL;                       // SpfRootNotFound
            roots[j] = r;
            for (uint32_t e = csr.row_ptr[r]; e < csr.row_ptr[r + 1]; ++e) {
                const uint32_t n = csr.col[e];
                if (is_router[n]) continue;
                bool dup = false;
                for (size_t q = gather_off[j]; q < gv.size() && !dup; ++q) dup = gv[q] == n;
                if (!dup) { gj.push_back(j); gv.push_back(n); }
            }
            gather_off[j + 1] = (uint32_t)gv.size();
        }
        const uint32_t G = (uint32_t)gv.size();
        if (cudaSetDevice(hspf_ctx_device(ctx)) != cudaSuccess) return HSPF_E_CUDA;
        if (G > gather_cap || (G && (!gather_v || !gather_nh))) return HSPF_E_NOMEM;
        rc = hspf_graph_upload(ctx, &csr, &g);
        if (rc) return rc;
        rc = hspf_ospfv2_rtable_upload(ctx, rt);
        if (rc) return rc;
        cudaStream_t st = static_cast<cudaStream_t>(hspf_stream(ctx));
        const size_t NV = (size_t)n_roots * V;
        DevBuf d_roots, d_dist, d_hops, d_nh, d_status, d_cells, d_gj, d_gv, d_gnh;
        if (d_roots.alloc(n_roots * 4) || d_dist.alloc(NV * 4) || d_hops.alloc(NV * 2) || d_nh.alloc(NV * 8) ||
            d_status.alloc(n_roots * 4) || d_cells.alloc((size_t)n_roots * P * sizeof(hl_route_cell)) ||
            d_gj.alloc(G * 4) || d_gv.alloc(G * 4) || d_gnh.alloc(G * 8)) return HSPF_E_NOMEM;
        for (auto &e : ev) if (cudaEventCreate(&e) != cudaSuccess) return HSPF_E_CUDA;
        if (cudaMemcpyAsync(d_roots.p, roots.data(), n_roots * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return HSPF_E_CUDA;
        if (G) {
            if (cudaMemcpyAsync(d_gj.p, gj.data(), G * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                cudaMemcpyAsync(d_gv.p, gv.data(), G * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return HSPF_E_CUDA;
        }
        hspf_jobs jobs{};
        jobs.n_jobs = n_roots;
        jobs.roots = d_roots.as<uint32_t>();
        hspf_result res{};
        res.dist = d_dist.as<uint32_t>(); res.hops = d_hops.as<uint16_t>(); res.nh_mask = d_nh.as<uint64_t>();
        res.nh_words = 1; res.job_status = d_status.as<uint32_t>();
        cudaEventRecord(ev[0], st);
        rc = hspf_run_batch_async(ctx, g, &jobs, &res);
        if (rc) return rc;
        cudaEventRecord(ev[1], st);
        rc = hspf_ospfv2_routes_batch(ctx, rt, n_roots, &res, d_cells.as<hl_route_cell>(), G, d_gj.as<uint32_t>(),
                                      d_gv.as<uint32_t>(), d_gnh.as<uint64_t>());
        if (rc) return rc;
        cudaEventRecord(ev[2], st);
        if (cudaMemcpyAsync(job_status, d_status.p, n_roots * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return HSPF_E_CUDA;
        if ((size_t)n_roots * P &&
            cudaMemcpyAsync(cells, d_cells.p, (size_t)n_roots * P * sizeof(hl_route_cell), cudaMemcpyDeviceToHost, st) != cudaSuccess)
            return HSPF_E_CUDA;
        if (G && cudaMemcpyAsync(gather_nh, d_gnh.p, G * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return HSPF_E_CUDA;
        if (cudaStreamSynchronize(st) != cudaSuccess) return HSPF_E_CUDA;
        if (G) std::memcpy(gather_v, gv.data(), G * 4);
        if (device_ms) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, ev[0], ev[1]);
            cudaEventElapsedTime(&b, ev[1], ev[2]);
            device_ms[0] = a; device_ms[1] = b;
        }
        for (uint32_t j = 0; j < n_roots; ++j) if (job_status[j]) return HSPF_E_JOB_STATUS;
        return HSPF_OK;
    } catch (const std::bad_alloc &) {
        return HSPF_E_NOMEM;
    } catch (...) {
        return HSPF_E_INVAL;
    }
}

}  // extern "C"
