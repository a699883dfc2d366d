// spf_kernel.cuh — device side of the batched SPF engine (sm_100a).
//
// One persistent CTA per SM slot; each CTA loops over jobs (roots / perturbed
// topologies).  All per-job mutable state (tentative distances, frontier
// queues, ECMP in-degree counters, hop counts, DAG-edge bitmaps) lives in shared
// memory; the read-only CSR (forward + transposed) is shared by every CTA of the
// launch and is served from L1/L2 (a 10k-vertex LSDB is < 1 MB, the B200 L2 is
// 126 MB).
//
// Per job the kernel reproduces the result of the reference Dijkstra
//   holo-ospf/src/spf.rs:587-729 (run_area) / holo-isis/src/spf.rs:525-707
// under the static-order assumption of SURVEY.md §8a "semantics that bite" #1
// (no zero-cost edge out of a hop-counting vertex, no saturation), in three
// phases:
//   1. SSSP   : near/far bucketed label-correcting relaxation, atomicMin on the
//               shared-memory distance array (distances are order independent);
//   2. parents: pull pass over the transposed CSR: ECMP in-degree, first parent =
//               DAG parent with the smallest (distance, id), i.e. the vertex whose
//               relaxation created the final candidate entry in the reference
//               (spf.rs:700-703); marks every ECMP-DAG edge and every first-parent
//               edge in two E-bit shared-memory bitmaps;
//   3. hops and next-hop atom sets (hops follow the first-parent edge, spf.rs:675-678; the
//               sets are OR-ed over all ECMP parents, spf.rs:747-766 / holo-isis spf.rs:678-702):
//      3J jump : common batch shape (kFast, packed CSR twins, <= 16 first-hop atoms): pointer
//               doubling over the first-parent tree, ECMP vertices as jump terminals resolved
//               by monotone sweeps — O(log depth) regular rounds;
//      3K Kahn : every other case: topological push over the marked DAG edges only, one round
//               per DAG level.
//
// Shared-memory plan (byte offsets computed on the host, see make_layout):
//   SSSP   : dist[V] u32 | qa[V] | qb[V] | pend[V] u16 (fast path: row16[V], kept across jobs) | bm0 | bm1
//   parents: dist (read) | 3K: dagbit,fpbit -> qa region, pend = in-degrees | 3J: first_parent u16 -> qb region,
//            bm0 = hops-0 heads of root edges, bm1 = ECMP vertices
//   3J     : jump words (ancestor:16 | aggregate:16) -> dist region (dist is written back to HBM first),
//            ECMP vertex list -> qb region
//   3K     : kq0,kq1 -> dist region | bitmaps | hops[V] u16 -> qb region | pend
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hspf {

constexpr uint32_t kInf = 0xFFFFFFFFu;
constexpr int kThreads = 512;
constexpr int kMaxOv = 8;          // HSPF_MAX_OVERRIDES
constexpr int kMaxRootDeg = 256;   // non-HOP root neighbours tracked in smem
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kVfHop = 1u, kVfLeaf = 2u, kVfLeafUnlessRoot = 4u;
constexpr uint32_t kGfNoHopTargetNoNh = 1u, kGfHopCount = 2u;
constexpr uint32_t kJsSaturated = 1u, kJsTooManyAtoms = 2u, kJsOrder = 4u, kJsBadJob = 8u;   // HSPF_JS_*

struct DevGraph {
    uint32_t V, E;
    const uint32_t *row;    // [V+1]
    const uint2 *edge;      // [E] {col, cost}
    const uint32_t *irow;   // [V+1] transposed
    const uint4 *iedge;     // [E] {src, cost, forward edge index, 0}: one 16 B load per in-edge
    const uint8_t *vflags;  // [V]
    const uint16_t *row16;  // [V] row offsets as u16 (null unless V, E < 65536)
    const uint32_t *iedge16;// [E] in-edges as (src | cost << 16) (null unless ids and costs fit 16 bits)
    const uint32_t *edge16; // [E] forward edges as (head | cost << 16), same condition
    const uint32_t *iquad_row; // [V+1] first quad of each vertex in iquad
    const uint4 *iquad;     // in-edge records of iedge16, each vertex padded to whole quads with 0xFFFFFFFF
    uint32_t reject_above, saturate_at, flags, delta;
};

struct Layout {   // byte offsets into the per-CTA state block
    uint32_t dist, qa, qb, pend, bm0, bm1;     // SSSP
    uint32_t fl_hop, fl_leaf, fl_lur;          // vertex-flag bitmaps (whole kernel)
    uint32_t kq0, kq1, hops, dagbit, fpbit;    // parents + Kahn
    uint32_t total;
};

struct BatchArgs {
    DevGraph g;
    Layout lay;
    uint32_t n_jobs;
    const uint32_t *roots;
    const uint32_t *ov_off;   // may be null
    const uint32_t *ov_edge;
    const uint32_t *ov_cost;
    uint32_t *out_dist;       // [n_jobs][V]
    uint16_t *out_hops;
    uint32_t *out_fp;
    uint16_t *out_npar;
    uint64_t *out_nh;         // [n_jobs][V][nhw]
    uint32_t *out_status;     // [n_jobs]
    uint32_t nhw;
    uint8_t *ws;              // per-CTA global workspace when state does not fit smem
    size_t ws_stride;         // bytes per CTA (0: state in smem)
    uint32_t *job_counter;    // dynamic job fetch
    uint32_t jump_ok;         // layout allows the pointer-jumping next-hop phase (see phase 3J)
    unsigned long long *prof; // optional [gridDim][16] per-phase cycle counters (debug), may be null
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Per-job state layout for V vertices, E edges and queue entry size qsz (2 or 4).
inline Layout make_layout(uint32_t V, uint32_t E, int qsz) {
    Layout L{};
    const size_t Vp = align_up(V, 4);
    const size_t nbw = (Vp + 31) / 32, nbe = ((size_t)E + 31) / 32;
    const size_t sz_dist = Vp * 4, sz_q = align_up(Vp * qsz, 16), sz_pend = align_up(Vp * 2, 16);
    const size_t sz_bm = align_up(nbw * 4, 16), sz_eb = align_up(nbe * 4, 16), sz_hops = align_up(Vp * 2, 16);
    size_t o = 0;
    L.dist = (uint32_t)o; o += sz_dist;
    L.qa = (uint32_t)o; o += sz_q;
    L.qb = (uint32_t)o; o += sz_q;
    L.pend = (uint32_t)o; o += sz_pend;
    L.bm0 = (uint32_t)o; o += sz_bm;
    L.bm1 = (uint32_t)o; o += sz_bm;
    L.fl_hop = (uint32_t)o; o += sz_bm;
    L.fl_leaf = (uint32_t)o; o += sz_bm;
    L.fl_lur = (uint32_t)o; o += sz_bm;
    // parents/Kahn arrays: alias onto SSSP arrays that are dead by then, else append
    if (2 * sz_eb <= sz_q) { L.dagbit = L.qa; L.fpbit = L.qa + (uint32_t)sz_eb; }
    else { L.dagbit = (uint32_t)o; o += sz_eb; L.fpbit = (uint32_t)o; o += sz_eb; }
    L.hops = L.qb;                                   // sz_hops <= sz_q always (qsz >= 2)
    if (2 * sz_q <= sz_dist) { L.kq0 = L.dist; L.kq1 = L.dist + (uint32_t)sz_q; }
    else { L.kq0 = (uint32_t)o; o += sz_q; L.kq1 = (uint32_t)o; o += sz_q; }
    (void)sz_hops;
    L.total = (uint32_t)align_up(o, 16);
    return L;
}

struct Ov {   // overrides of the current job, in smem
    uint32_t n;
    uint32_t tail[kMaxOv], head[kMaxOv], edge[kMaxOv], cost[kMaxOv];
};

struct Small {  // small per-CTA control block in smem
    Ov ov;
    uint32_t cnt[2];      // alternating queue fill counters
    uint32_t scan_min;
    uint32_t status;
    uint32_t job;
    uint32_t n_roottab;
    uint32_t rt_target[kMaxRootDeg];  // non-HOP heads of root edges
    uint32_t rt_base[kMaxRootDeg];    // first-hop atom base of that head
    uint32_t rt_cost[kMaxRootDeg];    // cost of that root edge
    uint32_t root_rb;                 // row[root]
    uint32_t n_atoms;                 // first-hop atoms of this root (root edges + edges of its non-HOP heads)
};

__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    return s < a ? kInf : s;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// Warp-aggregated queue push (all currently converged lanes take part).
template <typename VT>
__device__ __forceinline__ void q_push(VT *q, uint32_t *counter, uint32_t v) {
    const unsigned active = __activemask();
    const int leader = __ffs(active) - 1;
    const unsigned rank = __popc(active & ((1u << lane_id()) - 1));
    uint32_t base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(counter, __popc(active));
    base = __shfl_sync(active, base, leader);
    q[base + rank] = (VT)v;
}

__device__ __forceinline__ bool expands(uint32_t fl, uint32_t u, uint32_t root) {
    return !((fl & kVfLeaf) || ((fl & kVfLeafUnlessRoot) && u != root));
}

// Phase timing (debug aid, enabled by passing BatchArgs.prof): thread 0 adds the
// cycles since the previous mark to slot `k` of its CTA's row.
#define HSPF_MARK(k)                                                             \
    do {                                                                         \
        if (a.prof && tid == 0) {                                                \
            const long long now_ = clock64();                                    \
            a.prof[(size_t)blockIdx.x * 16 + (k)] += (unsigned long long)(now_ - t_mark); \
            t_mark = now_;                                                       \
        }                                                                        \
    } while (0)

// Compact the set bits of a V-bit bitmap into queue `q` (vertex ids, any order),
// clearing the bitmap; `keep(v)` filters.  One word per thread, warp-aggregated
// reservation in *counter.  Call from all threads; the caller supplies the barriers.
template <bool kFilter, typename VT, typename Keep>
__device__ __forceinline__ void bitmap_to_queue(uint32_t *bm, uint32_t nbw, VT *q, uint32_t *counter, Keep keep) {
    for (uint32_t w0 = 0; w0 < nbw; w0 += kThreads) {
        const uint32_t w = w0 + threadIdx.x;
        uint32_t bits = 0;
        if (w < nbw) {
            bits = bm[w];
            if (bits) {
                bm[w] = 0;
                if (kFilter) {
                    uint32_t kept = 0;
                    for (uint32_t b = bits; b; b &= b - 1) {
                        const uint32_t bit = __ffs(b) - 1;
                        if (keep(w * 32 + bit)) kept |= 1u << bit;
                    }
                    bits = kept;
                }
            }
        }
        const uint32_t n = __popc(bits);
        uint32_t incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)(threadIdx.x & 31) >= o) incl += t;
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = 0;
        if (tot) {
            if ((threadIdx.x & 31) == 31) base = atomicAdd(counter, tot);
            base = __shfl_sync(0xffffffffu, base, 31);
        }
        uint32_t pos = base + incl - n;
        for (uint32_t b = bits; b; b &= b - 1) q[pos++] = (VT)(w * 32 + (__ffs(b) - 1));
    }
}

// kSmemState: the per-job state block lives in dynamic shared memory (the normal
// case).  It is a template parameter, not a runtime select, so that the compiler
// can prove the address space and emit LDS/STS/ATOMS instead of generic LD/ST/ATOM
// (generic atomics that land in the shared window are an order of magnitude slower).
// kFast: the common batch shape — no edge overrides, no hop-count mode, no LEAF
// vertex flags, one next-hop word — with those branches compiled out.
template <typename VT, bool kSmemState, bool kFast>
__global__ void __launch_bounds__(kThreads, 2) spf_batch_kernel(const BatchArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ Small S;

    const DevGraph &g = a.g;
    const Layout &L = a.lay;
    const uint32_t V = g.V;
    const uint32_t tid = threadIdx.x;
    const uint32_t Vp = (uint32_t)align_up(V, 4);
    const uint32_t nbw = (Vp + 31) / 32;
    const uint32_t nbe = (g.E + 31) / 32;

    uint8_t *base;
    if constexpr (kSmemState) base = smem_raw;
    else base = a.ws + (size_t)blockIdx.x * a.ws_stride;
    uint32_t *dist = reinterpret_cast<uint32_t *>(base + L.dist);
    VT *qa = reinterpret_cast<VT *>(base + L.qa);
    VT *qb = reinterpret_cast<VT *>(base + L.qb);
    uint32_t *pend32 = reinterpret_cast<uint32_t *>(base + L.pend);  // packed u16 pairs
    uint16_t *pend = reinterpret_cast<uint16_t *>(base + L.pend);
    uint32_t *bm0 = reinterpret_cast<uint32_t *>(base + L.bm0);
    uint32_t *bm1 = reinterpret_cast<uint32_t *>(base + L.bm1);
    uint32_t *fl_hop = reinterpret_cast<uint32_t *>(base + L.fl_hop);
    uint32_t *fl_leaf = reinterpret_cast<uint32_t *>(base + L.fl_leaf);
    uint32_t *fl_lur = reinterpret_cast<uint32_t *>(base + L.fl_lur);
    VT *kq0 = reinterpret_cast<VT *>(base + L.kq0);
    VT *kq1 = reinterpret_cast<VT *>(base + L.kq1);
    uint16_t *hops_s = reinterpret_cast<uint16_t *>(base + L.hops);
    uint32_t *dagbit = reinterpret_cast<uint32_t *>(base + L.dagbit);
    uint32_t *fpbit = reinterpret_cast<uint32_t *>(base + L.fpbit);

    const uint32_t nhw = kFast ? 1u : a.nhw;
    long long t_mark = clock64();
    const uint32_t lane = threadIdx.x & 31;

    // vertex flags as three V-bit bitmaps in shared memory (built once per CTA)
    for (uint32_t w = tid; w < nbw; w += kThreads) {
        uint32_t h = 0, l = 0, r = 0;
        for (uint32_t b = 0; b < 32; ++b) {
            const uint32_t v = w * 32 + b;
            if (v < V) {
                const uint32_t f = g.vflags[v];
                h |= ((f & kVfHop) ? 1u : 0u) << b;
                l |= ((f & kVfLeaf) ? 1u : 0u) << b;
                r |= ((f & kVfLeafUnlessRoot) ? 1u : 0u) << b;
            }
        }
        fl_hop[w] = h; fl_leaf[w] = l; fl_lur[w] = r;
    }
    auto is_hop = [&](uint32_t v) -> bool { return (fl_hop[v >> 5] >> (v & 31)) & 1u; };
    auto vexpands = [&](uint32_t u, uint32_t root_) -> bool {
        if (kFast) return true;
        const uint32_t b = 1u << (u & 31);
        return !((fl_leaf[u >> 5] & b) || ((fl_lur[u >> 5] & b) && u != root_));
    };

    // Jump-capable launch (see phase 3J): the `pend` block is not needed by the jump path,
    // so it keeps the CSR row offsets (u16) across jobs and the SSSP reads them from
    // shared memory instead of L2.
    const bool jumpable = kFast && kSmemState && sizeof(VT) == 2 && a.jump_ok;
    const bool use_row16 = jumpable && g.row16 != nullptr;
    uint16_t *row_s = pend;
    bool row_valid = false;

    for (;;) {
        // ---- fetch next job -------------------------------------------------
        __syncthreads();
        if (tid == 0) S.job = atomicAdd(a.job_counter, 1u);
        __syncthreads();
        const uint32_t job = S.job;
        if (job >= a.n_jobs) break;
        const uint32_t root = a.roots[job];
        {
            // device-pointer callers are not validated on the host: a malformed job is flagged and skipped
            uint32_t n_raw = 0;
            if (a.ov_off) n_raw = a.ov_off[job + 1] - a.ov_off[job];
            bool bad = root >= V || n_raw > (uint32_t)kMaxOv;
            for (uint32_t k = 0; !bad && k < n_raw; ++k) bad = a.ov_edge[a.ov_off[job] + k] >= g.E;
            if (bad) {
                if (tid == 0) a.out_status[job] = kJsBadJob;
                continue;
            }
        }
        const size_t jo = (size_t)job * V;
        uint16_t *o_hops = a.out_hops + jo;
        uint32_t *o_fp = a.out_fp + jo;
        uint16_t *o_npar = a.out_npar + jo;
        uint64_t *o_nh = a.out_nh + jo * nhw;
        uint32_t *o_dist = a.out_dist + jo;

        // ---- per-job init -----------------------------------------------------
        for (uint32_t v = tid; v < Vp; v += kThreads) dist[v] = kInf;
        for (uint32_t w = tid; w < nbw; w += kThreads) { bm0[w] = 0; bm1[w] = 0; }
        if (use_row16 && !row_valid) {
            for (uint32_t v = tid; v < V; v += kThreads) row_s[v] = g.row16[v];
            row_valid = true;
        }
        if (tid == 0) {
            S.status = 0;
            S.cnt[0] = 1;
            S.cnt[1] = 0;
            S.scan_min = kInf;
            uint32_t n = 0;
            if (a.ov_off) n = a.ov_off[job + 1] - a.ov_off[job];
            S.ov.n = n > kMaxOv ? kMaxOv : n;
        }
        __syncthreads();
        if (tid < S.ov.n) {
            const uint32_t e = a.ov_edge[a.ov_off[job] + tid];
            const uint32_t c = a.ov_cost[a.ov_off[job] + tid];
            uint32_t lo = 0, hi = V;  // invariant row[lo] <= e < row[hi]
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (g.row[mid] <= e) lo = mid; else hi = mid;
            }
            S.ov.tail[tid] = lo;
            S.ov.head[tid] = g.edge[e].x;
            S.ov.edge[tid] = e;
            S.ov.cost[tid] = c;
            if (c == 0 && (g.vflags[lo] & kVfHop)) atomicOr(&S.status, kJsOrder);
        }
        if (tid == 0) {
            dist[root] = 0;
            qa[0] = (VT)root;
        }
        __syncthreads();
        const uint32_t n_ov = kFast ? 0u : S.ov.n;
        // root edge table: first-hop atom bases behind the root's non-HOP neighbours
        if (tid == 32) {
            const uint32_t rb = g.row[root], re = g.row[root + 1];
            uint32_t nt = 0, nextbase = re - rb;
            for (uint32_t e = rb; e < re; ++e) {
                const uint32_t h = g.edge[e].x;
                if (!(g.vflags[h] & kVfHop)) {
                    if (nt < (uint32_t)kMaxRootDeg) {
                        S.rt_target[nt] = h;
                        S.rt_base[nt] = nextbase;
                        S.rt_cost[nt] = g.edge[e].y;
                        ++nt;
                    } else {
                        atomicOr(&S.status, kJsTooManyAtoms);
                    }
                    nextbase += g.row[h + 1] - g.row[h];
                }
            }
            S.n_roottab = nt;
            S.root_rb = rb;
            S.n_atoms = nextbase;
        }
        __syncthreads();

        HSPF_MARK(0);   // job fetch + init
        // ======================= phase 1: SSSP ==================================
        const uint32_t delta = g.delta ? g.delta : 1u;
        uint32_t hi_thr = delta;       // near bucket is [*, hi_thr)
        uint32_t *bm_next = bm0;
        VT *qcur = qa, *qnext = qb;
        uint32_t p = 0;                // S.cnt[p] counts qcur, S.cnt[p^1] counts qnext
        for (;;) {
            for (;;) {
                const uint32_t n_cur = S.cnt[p];
                if (n_cur == 0) break;
                long long t_sub = 0;
                if (a.prof && tid == 0) { t_sub = clock64(); a.prof[(size_t)blockIdx.x * 16 + 12] += n_cur; }
                // Team expansion: a warp pass covers 32 frontier vertices as 4 independent
                // streams of 8 vertices; each vertex is served by a team of 4 lanes that
                // strides over its edge list.  No prefix sums or owner searches, the four
                // streams give the scheduler independent work, and an edge fetch covers
                // 4 x 32 edges per iteration.
                for (uint32_t i0 = (tid >> 5) * 32; i0 < n_cur; i0 += kThreads) {
                    const uint32_t sub = lane & 3, tv = lane >> 2;
                    uint32_t du[4], eb[4], ee[4], uu[4];
#pragma unroll
                    for (int st = 0; st < 4; ++st) {
                        const uint32_t i = i0 + st * 8 + tv;
                        uu[st] = (i < n_cur) ? (uint32_t)qcur[i] : kInf;
                    }
#pragma unroll
                    for (int st = 0; st < 4; ++st) {
                        du[st] = 0; eb[st] = 0; ee[st] = 0;
                        if (uu[st] != kInf && vexpands(uu[st], root)) {
                            du[st] = dist[uu[st]];
                            if (use_row16) {
                                eb[st] = row_s[uu[st]];
                                ee[st] = (uu[st] + 1 < V) ? (uint32_t)row_s[uu[st] + 1] : g.E;
                            } else {
                                eb[st] = g.row[uu[st]];
                                ee[st] = g.row[uu[st] + 1];
                            }
                        }
                    }
                    uint32_t maxdeg = 0;
#pragma unroll
                    for (int st = 0; st < 4; ++st) maxdeg = max(maxdeg, ee[st] - eb[st]);
                    maxdeg = __reduce_max_sync(0xffffffffu, maxdeg);
                    if (jumpable) {
                        // packed edges (head | cost << 16): a team covers 8 edges of its vertex per
                        // iteration with all 8 loads of the lane in flight together, so degrees up
                        // to 8 cost one L2 round trip
                        // A padding lane relaxes the team's own vertex with cost 0xFFFF (never an
                        // improvement), so the body needs no validity branch and the eight
                        // distance reads go out together.  No overflow: a finite distance is at
                        // most 65534 * 65535.
                        uint32_t pad[4];
#pragma unroll
                        for (int st = 0; st < 4; ++st) pad[st] = ((uu[st] == kInf) ? root : uu[st]) | 0xFFFF0000u;
                        for (uint32_t k = 0; k < maxdeg; k += 8) {
                            uint32_t r[8], dh[8];
#pragma unroll
                            for (int st = 0; st < 4; ++st) {
                                const uint32_t e = eb[st] + k + sub;
                                r[st] = (e < ee[st]) ? g.edge16[e] : pad[st];
                                r[st + 4] = (e + 4 < ee[st]) ? g.edge16[e + 4] : pad[st];
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) dh[j] = dist[r[j] & 0xFFFFu];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const uint32_t v = r[j] & 0xFFFFu;
                                const uint32_t nd = du[j & 3] + (r[j] >> 16);
                                if (nd <= g.reject_above && nd < dh[j]) {
                                    atomicMin(&dist[v], nd);
                                    if (nd < hi_thr) atomicOr(&bm_next[v >> 5], 1u << (v & 31));
                                }
                            }
                        }
                        continue;
                    }
                    for (uint32_t k = 0; k < maxdeg; k += 4) {
                        uint2 ec[4];
#pragma unroll
                        for (int st = 0; st < 4; ++st) {
                            const uint32_t e = eb[st] + k + sub;
                            ec[st] = (e < ee[st]) ? g.edge[e] : make_uint2(0u, kInf);
                        }
#pragma unroll
                        for (int st = 0; st < 4; ++st) {
                            uint32_t c = ec[st].y;
                            if (!kFast) {
                                const uint32_t e = eb[st] + k + sub;
                                for (uint32_t q = 0; q < n_ov; ++q)
                                    if (S.ov.tail[q] == uu[st] && S.ov.edge[q] == e && c != kInf) c = S.ov.cost[q];
                            }
                            if (c == kInf) continue;     // padding lane or disabled edge
                            const uint32_t nd = sat_add(du[st], c);
                            const uint32_t v = ec[st].x;
                            if (nd <= g.reject_above && nd < dist[v]) {
                                // fire-and-forget: nothing below waits on an atomic's result;
                                // the next frontier is compacted from the bitmap at round end
                                atomicMin(&dist[v], nd);
                                if (nd < hi_thr) atomicOr(&bm_next[v >> 5], 1u << (v & 31));
                            }
                        }
                    }
                }
                if (a.prof && tid == 0) { const long long n_ = clock64(); a.prof[(size_t)blockIdx.x * 16 + 8] += n_ - t_sub; t_sub = n_; }
                __syncthreads();
                if (a.prof && tid == 0) { const long long n_ = clock64(); a.prof[(size_t)blockIdx.x * 16 + 9] += n_ - t_sub; t_sub = n_; }
                if (a.prof && tid == 0) a.prof[(size_t)blockIdx.x * 16 + 7] += 1;   // SSSP rounds
                // every thread has consumed S.cnt[p]; recycle it for the round after next
                if (tid == 0) S.cnt[p] = 0;
                // next frontier = vertices marked this round (each once, the bitmap dedups)
                bitmap_to_queue<false>(bm_next, nbw, qnext, &S.cnt[p ^ 1], [](uint32_t) { return true; });
                { VT *t = qcur; qcur = qnext; qnext = t; }
                p ^= 1;
                if (a.prof && tid == 0) { const long long n_ = clock64(); a.prof[(size_t)blockIdx.x * 16 + 10] += n_ - t_sub; t_sub = n_; }
                __syncthreads();
                if (a.prof && tid == 0) { const long long n_ = clock64(); a.prof[(size_t)blockIdx.x * 16 + 11] += n_ - t_sub; }
            }
            // near bucket exhausted (S.cnt[0] == S.cnt[1] == 0): everything below
            // hi_thr is settled.  Collect the next bucket [m, m + delta) into qcur,
            // where m = min unsettled distance.
            uint32_t lo_thr = hi_thr;
            bool done = false;
            for (;;) {
                if (tid == 0) { S.scan_min = kInf; S.cnt[p] = 0; }
                __syncthreads();
                const uint32_t hi2 = sat_add(lo_thr, delta);
                uint32_t lmin = kInf;
                // a warp looks at 32 consecutive vertices per step, so the vertices that fall
                // into the new bucket are one ballot = one word of the frontier bitmap
                for (uint32_t v0 = tid & ~31u; v0 < V; v0 += kThreads) {
                    const uint32_t v = v0 + lane;
                    const uint32_t d = (v < V) ? dist[v] : kInf;
                    const bool open = d >= lo_thr && d != kInf;
                    if (open) lmin = min(lmin, d);
                    const uint32_t b = __ballot_sync(0xffffffffu, open && d < hi2);
                    if (lane == 0) bm_next[v0 >> 5] = b;
                }
                for (int o = 16; o > 0; o >>= 1) lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
                if (lane_id() == 0 && lmin != kInf) atomicMin(&S.scan_min, lmin);
                __syncthreads();
                bitmap_to_queue<false>(bm_next, nbw, qcur, &S.cnt[p], [](uint32_t) { return true; });
                __syncthreads();
                const uint32_t found = S.cnt[p];
                const uint32_t m = S.scan_min;
                __syncthreads();   // all reads done before a possible reset above
                if (found > 0) { hi_thr = hi2; break; }
                if (m == kInf) { done = true; break; }
                lo_thr = m;   // jump over the empty gap and rescan
            }
            if (done) break;
        }
        HSPF_MARK(1);   // SSSP
        // SSSP done: qa/qb/bm0/bm1 are dead, dist is final.
        // Next hops are propagated either by pointer jumping over the first-parent tree
        // (phase 3J: O(log depth) regular rounds) or by a Kahn push over the ECMP DAG
        // (phase 3K: one round per DAG level; handles every case).
        const bool use_jump = jumpable && S.n_atoms <= 16u;
        uint16_t *fp16 = hops_s;          // jump path: first parent per vertex (0xFFFF: none)
        uint32_t *h0bm = bm0;             // jump path: non-HOP heads of root edges that sit at hops 0
        uint32_t *ecmpbm = bm1;           // jump path: vertices with two or more parents
        for (uint32_t w = tid; w < nbw; w += kThreads) { bm0[w] = 0; bm1[w] = 0; }
        if (!use_jump) {
            row_valid = false;   // the Kahn path overwrites the pend block
            for (uint32_t w = tid; w < nbe; w += kThreads) { dagbit[w] = 0; fpbit[w] = 0; }
            for (uint32_t v = tid; v < Vp; v += kThreads) hops_s[v] = 0;
            // zero the next-hop plane (the Kahn push accumulates it with atomics)
            const size_t n = (size_t)V * nhw;
            for (size_t i = tid; i < n; i += kThreads) o_nh[i] = 0ull;
        }
        __syncthreads();
        if (use_jump) {
            if (tid < S.n_roottab) {
                const uint32_t h = S.rt_target[tid];
                if (dist[h] == S.rt_cost[tid]) atomicOr(&h0bm[h >> 5], 1u << (h & 31));
            }
            __syncthreads();
        }
        // u sits at hops 0: the root, or a non-HOP vertex whose first parent is the root
        auto hops0 = [&](uint32_t u) -> bool {
            return u == root || ((h0bm[u >> 5] >> (u & 31)) & 1u);
        };

        // ======================= phase 2: ECMP parents (pull) ====================
        uint32_t sat_flag = 0;
        uint32_t seed_v = kInf;   // jump path: head of this thread's first-hop atom edge
        if (use_jump) {
            // Packed in-edges (source | cost << 16), four per 16 bytes; no bitmaps and no
            // atomics: the jump phase needs only the first parent and the ECMP flag.
            // The quad range of the next vertex and its first quad are fetched one iteration ahead.
            const uint4 kPad = make_uint4(0u, 0u, 0u, 0u);   // placeholder, never examined (empty quad range)
            uint32_t qb_n = 0, qe_n = 0;
            uint4 r_n = kPad;
            if (tid < V) {
                qb_n = g.iquad_row[tid]; qe_n = g.iquad_row[tid + 1];
                if (qb_n < qe_n) r_n = g.iquad[qb_n];
            }
            for (uint32_t v = tid; v < Vp; v += kThreads) {
                const uint32_t qbeg = qb_n, qend = qe_n;
                uint4 r4 = r_n;
                qb_n = qe_n = 0; r_n = kPad;
                if (v + kThreads < V) {
                    qb_n = g.iquad_row[v + kThreads]; qe_n = g.iquad_row[v + kThreads + 1];
                    if (qb_n < qe_n) r_n = g.iquad[qb_n];
                }
                uint32_t cnt = 0, bu = kInf, bd = kInf;
                if (v < V) {
                    const uint32_t dv = dist[v];
                    if (dv != kInf && g.saturate_at && dv >= g.saturate_at) sat_flag = 1;
                    if (v != root && dv != kInf) {
                        for (uint32_t q = qbeg; q < qend; ++q) {
                            if (q != qbeg) r4 = g.iquad[q];
                            // branch-free: a pad record is (v | 0xFFFF << 16), which can never
                            // satisfy dist[v] + 65535 == dist[v]
                            const uint32_t r[4] = {r4.x, r4.y, r4.z, r4.w};
                            uint32_t du[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) du[k] = dist[r[k] & 0xFFFFu];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t u = r[k] & 0xFFFFu;
                                const bool ok = du[k] != kInf && du[k] + (r[k] >> 16) == dv;
                                const bool better = ok && (du[k] < bd || (du[k] == bd && u < bu));
                                cnt += ok ? 1u : 0u;
                                bd = better ? du[k] : bd;
                                bu = better ? u : bu;
                            }
                        }
                    }
                    o_fp[v] = bu;
                    o_npar[v] = (uint16_t)min(cnt, 0xFFFFu);
                }
                fp16[v] = (uint16_t)(cnt ? bu : 0xFFFFu);
                if (cnt >= 2) atomicOr(&ecmpbm[v >> 5], 1u << (v & 31));
            }
            if (tid < S.n_atoms) {   // one thread per first-hop atom: is its edge in the DAG?
                const uint32_t atom = tid, rdeg = g.row[root + 1] - S.root_rb;
                uint32_t e = kInf, u = root;
                if (atom < rdeg) {
                    e = S.root_rb + atom;
                } else {
                    for (uint32_t k = 0; k < S.n_roottab; ++k) {
                        const uint32_t N = S.rt_target[k], nb = S.rt_base[k];
                        if (atom < nb || atom >= nb + (g.row[N + 1] - g.row[N])) continue;
                        bool first = true;   // parallel root edges: only the first one's range is used
                        for (uint32_t q = 0; q < k; ++q) first = first && S.rt_target[q] != N;
                        if (first && hops0(N)) { u = N; e = g.row[N] + (atom - nb); }
                        break;
                    }
                }
                if (e != kInf) {
                    const uint2 ec = g.edge[e];
                    const uint32_t du = dist[u], dh = dist[ec.x];
                    if (du != kInf && dh != kInf && ec.y != kInf && sat_add(du, ec.y) == dh &&
                        !((g.flags & kGfNoHopTargetNoNh) && !is_hop(ec.x)))
                        seed_v = ec.x;
                }
            }
        } else {
            // in-edge range of the next vertex is fetched one iteration ahead
            uint32_t ib_n = 0, ie_n = 0;
            if (tid < V) { ib_n = g.irow[tid]; ie_n = g.irow[tid + 1]; }
            for (uint32_t v = tid; v < Vp; v += kThreads) {
                const uint32_t ib = ib_n, ie = ie_n;
                if (v + kThreads < V) { ib_n = g.irow[v + kThreads]; ie_n = g.irow[v + kThreads + 1]; }
                uint32_t cnt = 0, bu = kInf, be = kInf;
                unsigned long long bkey = ~0ull;   // (distance, id) of the best parent so far
                if (v < V) {
                    const uint32_t dv = dist[v];
                    if (dv != kInf && g.saturate_at && dv >= g.saturate_at) sat_flag = 1;
                    if (v != root && dv != kInf) {
                        // Hop-count mode: a pseudonode is parented only by the lowest-numbered
                        // attached router of its level (see HSPF_GF_HOPCOUNT in holo_spf.h).
                        uint32_t only_u = kInf;
                        const bool hopcount = kFast ? false : (g.flags & kGfHopCount) != 0;
                        if (hopcount && !is_hop(v)) {
                            for (uint32_t i = ib; i < ie; ++i) {
                                const uint4 sc1 = g.iedge[i];
                                const uint32_t u = sc1.x, d1 = dist[u];
                                if (d1 == kInf || !vexpands(u, root)) continue;
                                uint32_t c = sc1.y;
                                for (uint32_t q = 0; q < n_ov; ++q)
                                    if (S.ov.edge[q] == sc1.z) c = S.ov.cost[q];
                                if (c == kInf) continue;
                                if (sat_add(d1, c) == dv) only_u = min(only_u, u);
                            }
                        }
                        for (uint32_t i0 = ib; i0 < ie; i0 += 4) {
                            uint4 sc[4];
                            uint32_t du[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) sc[k] = (i0 + k < ie) ? g.iedge[i0 + k] : make_uint4(v, kInf, 0u, 0u);
#pragma unroll
                            for (int k = 0; k < 4; ++k) du[k] = dist[sc[k].x];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (i0 + k >= ie) break;
                                const uint32_t u = sc[k].x;
                                if (du[k] == kInf || !vexpands(u, root)) continue;
                                if (only_u != kInf && u != only_u) continue;
                                uint32_t c = sc[k].y;
                                const uint32_t e = sc[k].z;
                                if (!kFast) {
                                    for (uint32_t q = 0; q < n_ov; ++q)
                                        if (S.ov.edge[q] == e) c = S.ov.cost[q];
                                    if (c == kInf) continue;
                                }
                                if (sat_add(du[k], c) != dv) continue;
                                ++cnt;
                                atomicOr(&dagbit[e >> 5], 1u << (e & 31));
                                unsigned long long key = ((unsigned long long)du[k] << 32) | u;
                                if (hopcount) {
                                    // pop order inside a hop-count level is R1, pseudonodes first
                                    // reached from R1, R2, ...: a pseudonode sorts right after its
                                    // owner router
                                    if (!is_hop(u)) {
                                        uint32_t owner = kInf;
                                        for (uint32_t j = g.irow[u]; j < g.irow[u + 1]; ++j) {
                                            const uint4 s2 = g.iedge[j];
                                            if (dist[s2.x] != du[k] || !vexpands(s2.x, root)) continue;
                                            bool dis = false;
                                            for (uint32_t q = 0; q < n_ov; ++q)
                                                if (S.ov.edge[q] == s2.z && S.ov.cost[q] == kInf) dis = true;
                                            if (!dis) owner = min(owner, s2.x);
                                        }
                                        key = ((((unsigned long long)du[k] << 32) | owner) << 1) | 1ull;
                                    } else {
                                        key <<= 1;
                                    }
                                }
                                if (key < bkey || (key == bkey && u < bu)) { bkey = key; bu = u; be = e; }
                            }
                        }
                        if (cnt) atomicOr(&fpbit[be >> 5], 1u << (be & 31));
                    }
                    o_fp[v] = bu;
                    o_npar[v] = (uint16_t)min(cnt, 0xFFFFu);
                }
                pend[v] = (uint16_t)min(cnt, 0xFFFFu);
            }
        }
        if (sat_flag) atomicOr(&S.status, kJsSaturated);
        __syncthreads();
        HSPF_MARK(2);   // parents pull
        // distances are final: stream them out, then the region is reused by the Kahn queues
        for (uint32_t v = tid; v < V; v += kThreads) o_dist[v] = dist[v];
        __syncthreads();
        HSPF_MARK(3);   // dist write-back
        if (use_jump) {
            // ======================= phase 3J: pointer jumping =======================
            // Hops and next hops are path aggregates over the first-parent tree (sum of the
            // HOP flags, OR of the first-hop atoms), so they are computed by pointer doubling
            // in ceil(log2(depth)) rounds of perfectly regular work instead of one round per
            // DAG level.  Each vertex is one 32-bit word (ancestor:16 | aggregate:16) that
            // only its owner thread stores, so a reader always sees a consistent pair and the
            // rounds can update in place without double buffering.
            uint32_t *word = dist;
            // -- hops: sum of HOP flags over (root, v]
            for (uint32_t v = tid; v < Vp; v += kThreads) {
                const uint32_t f = fp16[v];
                word[v] = (f == 0xFFFFu) ? (v << 16) : ((f << 16) | (is_hop(v) ? 1u : 0u));
            }
            __syncthreads();
            for (;;) {
                int ch = 0;
                for (uint32_t v = tid; v < Vp; v += kThreads) {
                    const uint32_t w = word[v], A = w >> 16;
                    if (A != root && A != v) {
                        const uint32_t w2 = word[A];
                        word[v] = (w2 & 0xFFFF0000u) | ((w + w2) & 0xFFFFu);
                        ch |= (w2 >> 16) != root;
                    }
                }
                if (a.prof && tid == 0) a.prof[(size_t)blockIdx.x * 16 + 13] += 1;
                if (!__syncthreads_or(ch)) break;
            }
            for (uint32_t v = tid; v < V; v += kThreads) {
                const uint32_t w = word[v];
                const uint32_t h = (w >> 16) == v ? 0u : (w & 0xFFFFu);
                o_hops[v] = (uint16_t)h;
                // a hops-0 vertex that is not a head of a root edge cannot own atoms
                if (h == 0 && v != root && fp16[v] != 0xFFFFu && !hops0(v) && g.row[v + 1] != g.row[v])
                    atomicOr(&S.status, kJsTooManyAtoms);
            }
            __syncthreads();
            HSPF_MARK(5);   // jump path: hops
            // -- next hops.  nh[v] = atoms entering v | U nh[p] over DAG parents p that are not
            // at hops 0.  Between two ECMP vertices every vertex has one parent, so the tree is
            // cut below hops-0 vertices and AT ECMP vertices (they are jump terminals):
            //   word[v] = (top[v], atoms on the segment (top[v], v])
            // then the few ECMP vertices are resolved among themselves (monotone sweeps to
            // the fixpoint over a list of ~1 % of the vertices) and every vertex adds the
            // final set of its top.
            auto is_ecmp = [&](uint32_t v) -> bool { return (ecmpbm[v >> 5] >> (v & 31)) & 1u; };
            for (uint32_t v = tid; v < Vp; v += kThreads) {
                const uint32_t f = fp16[v];
                word[v] = ((f == 0xFFFFu || hops0(f)) ? root : f) << 16;
            }
            __syncthreads();
            if (seed_v != kInf) atomicOr(&word[seed_v], 1u << tid);   // first-hop atoms enter here
            __syncthreads();
            for (;;) {
                int ch = 0;
                for (uint32_t v = tid; v < Vp; v += kThreads) {
                    const uint32_t w = word[v], A = w >> 16;
                    if (A != root && !is_ecmp(A)) {
                        const uint32_t w2 = word[A], A2 = w2 >> 16;
                        word[v] = (w2 & 0xFFFF0000u) | ((w | w2) & 0xFFFFu);
                        ch |= A2 != root && !is_ecmp(A2);
                    }
                }
                if (a.prof && tid == 0) a.prof[(size_t)blockIdx.x * 16 + 14] += 1;
                if (!__syncthreads_or(ch)) break;
            }
            // ECMP vertices: own segment | final set of own top | the same of every other parent
            uint16_t *elist = fp16;   // (fp16 is dead once the words are initialised)
            if (tid == 0) S.cnt[0] = 0;
            __syncthreads();
            bitmap_to_queue<false>(ecmpbm, nbw, elist, &S.cnt[0], [](uint32_t) { return true; });
            __syncthreads();
            const uint32_t n_e = S.cnt[0];
            if (n_e) {
                // DAG parents of an ECMP vertex are re-derived from the distance plane just
                // written (L2); with one vertex per thread they are kept in registers.
                auto parents = [&](uint32_t x, auto &&f) {
                    const uint32_t dx = __ldcg(&o_dist[x]);
                    for (uint32_t j = g.irow[x]; j < g.irow[x + 1]; ++j) {
                        const uint32_t r = g.iedge16[j], u = r & 0xFFFFu;
                        if (sat_add(__ldcg(&o_dist[u]), r >> 16) == dx && !hops0(u)) f(u);
                    }
                };
                uint32_t pc[4] = {kInf, kInf, kInf, kInf};
                bool cached = false;
                if (n_e <= (uint32_t)kThreads && tid < n_e) {
                    uint32_t n = 0;
                    parents(elist[tid], [&](uint32_t u) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (n == (uint32_t)k) pc[k] = u;
                        ++n;
                    });
                    cached = n <= 4;
                }
                for (;;) {
                    int ch = 0;
                    for (uint32_t i = tid; i < n_e; i += kThreads) {
                        const uint32_t x = elist[i];
                        const uint32_t w = word[x], T = w >> 16;
                        uint32_t need = (T != root) ? word[T] : 0u;
                        auto pull_parent = [&](uint32_t u) {
                            const uint32_t wp = word[u], Tp = wp >> 16;
                            need |= wp;
                            if (Tp != root) need |= word[Tp];
                        };
                        if (cached) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) if (pc[k] != kInf) pull_parent(pc[k]);
                        } else {
                            parents(x, pull_parent);
                        }
                        need &= 0xFFFFu & ~w;
                        if (need) { word[x] = w | need; ch = 1; }
                    }
                    if (a.prof && tid == 0) a.prof[(size_t)blockIdx.x * 16 + 15] += 1;
                    if (!__syncthreads_or(ch)) break;
                }
            }
            for (uint32_t v = tid; v < V; v += kThreads) {
                const uint32_t w = word[v], T = w >> 16;
                uint32_t m = w;
                if (T != root) m |= word[T];
                o_nh[v] = (uint64_t)(m & 0xFFFFu);
            }
            if (tid == 0) a.out_status[job] = S.status;
            HSPF_MARK(4);
            continue;
        }
        if (tid == 0) { S.cnt[0] = 1; S.cnt[1] = 0; kq0[0] = (VT)root; }
        __syncthreads();

        // ======================= phase 3K: Kahn push over the ECMP DAG ===========
        {
            VT *kcur = kq0, *knext = kq1;
            p = 0;
            for (;;) {
                const uint32_t n_cur = S.cnt[p];
                if (n_cur == 0) break;
                long long t_sub = 0;
                if (a.prof && tid == 0) t_sub = clock64();
                // One frontier vertex per lane.  Its DAG out-edges are found with word-level
                // scans of the shared DAG bitmap (no memory traffic), collected four at a
                // time, and their heads fetched back to back (one L2 latency for up to four
                // children, together with the parent's next-hop words).  The ECMP DAG is
                // tree-like (~1.1 children per vertex), so one group is the common case.
                for (uint32_t i = tid; i < n_cur; i += kThreads) {
                    const uint32_t u = kcur[i];
                    const uint32_t hu = hops_s[u];
                    const uint32_t eb = g.row[u], ee = g.row[u + 1];
                    if (ee == eb) continue;
                    uint32_t abase = 0;
                    bool atoms_ok = true;
                    if (hu == 0 && u != root) {
                        atoms_ok = false;   // non-HOP vertex directly attached to the root
                        for (uint32_t k = 0; k < S.n_roottab; ++k)
                            if (S.rt_target[k] == u) { abase = S.rt_base[k]; atoms_ok = true; break; }
                        if (!atoms_ok) atomicOr(&S.status, kJsTooManyAtoms);
                    }
                    uint64_t nhu[4] = {0, 0, 0, 0};
                    bool nh_loaded = false;
                    const uint32_t wlast = (ee - 1) >> 5;
                    uint32_t w = eb >> 5;
                    uint32_t m = dagbit[w] & (0xFFFFFFFFu << (eb & 31));
                    if (w == wlast) m &= 0xFFFFFFFFu >> (31 - ((ee - 1) & 31));
                    for (;;) {
                        // collect up to 4 DAG edges
                        uint32_t e_[4], v_[4];
                        int n = 0;
                        while (n < 4) {
                            if (m == 0) {
                                if (w == wlast) break;
                                ++w;
                                m = dagbit[w];
                                if (w == wlast) m &= 0xFFFFFFFFu >> (31 - ((ee - 1) & 31));
                                continue;
                            }
                            e_[n++] = w * 32 + (__ffs(m) - 1);
                            m &= m - 1;
                        }
                        if (n == 0) break;
#pragma unroll
                        for (int k = 0; k < 4; ++k) if (k < n) v_[k] = g.edge[e_[k]].x;
                        if (hu != 0 && !nh_loaded) {
                            for (uint32_t x = 0; x < nhw; ++x) nhu[x] = __ldcg(&o_nh[(size_t)u * nhw + x]);
                            nh_loaded = true;
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k >= n) break;
                            const uint32_t e = e_[k], v = v_[k];
                            const bool is_fp = (fpbit[e >> 5] >> (e & 31)) & 1u;
                            const uint32_t hv = is_hop(v) ? 1u : 0u;
                            if (hu == 0) {
                                if (!((g.flags & kGfNoHopTargetNoNh) && !hv) && atoms_ok) {
                                    const uint32_t atom = abase + (e - eb);
                                    if (atom < 64u * nhw)
                                        atomicOr(reinterpret_cast<unsigned long long *>(&o_nh[(size_t)v * nhw + (atom >> 6)]),
                                                 1ull << (atom & 63));
                                    else
                                        atomicOr(&S.status, kJsTooManyAtoms);
                                }
                                if (is_fp) hops_s[v] = (uint16_t)hv;
                            } else {
                                for (uint32_t x = 0; x < nhw; ++x)
                                    if (nhu[x]) atomicOr(reinterpret_cast<unsigned long long *>(&o_nh[(size_t)v * nhw + x]), nhu[x]);
                                if (is_fp) hops_s[v] = (uint16_t)min(hu + hv, 0xFFFFu);
                            }
                            // packed u16 in-degree decrement + "touched" mark, both fire-and-forget;
                            // vertices whose counter reached zero are collected at round end
                            atomicSub(&pend32[v >> 1], 1u << ((v & 1) * 16));
                            atomicOr(&bm0[v >> 5], 1u << (v & 31));
                        }
                        if (n < 4) break;
                    }
                }
                if (a.prof && tid == 0) { const long long n_ = clock64(); a.prof[(size_t)blockIdx.x * 16 + 13] += n_ - t_sub; t_sub = n_; }
                __syncthreads();
                if (a.prof && tid == 0) { const long long n_ = clock64(); a.prof[(size_t)blockIdx.x * 16 + 14] += n_ - t_sub; t_sub = n_; }
                if (a.prof && tid == 0) a.prof[(size_t)blockIdx.x * 16 + 6] += 1;   // Kahn rounds
                if (tid == 0) S.cnt[p] = 0;
                bitmap_to_queue<true>(bm0, nbw, knext, &S.cnt[p ^ 1], [&](uint32_t v) { return pend[v] == 0; });
                { VT *t = kcur; kcur = knext; knext = t; }
                p ^= 1;
                __syncthreads();
                if (a.prof && tid == 0) { const long long n_ = clock64(); a.prof[(size_t)blockIdx.x * 16 + 15] += n_ - t_sub; }
            }
        }
        __syncthreads();

        HSPF_MARK(4);   // Kahn
        // ======================= write-back ======================================
        for (uint32_t v = tid; v < V; v += kThreads) o_hops[v] = hops_s[v];
        if (tid == 0) a.out_status[job] = S.status;
        HSPF_MARK(5);   // hops write-back
    }
}

}  // namespace hspf
