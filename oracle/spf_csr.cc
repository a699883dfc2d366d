// oracle/spf_csr.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's SPF main loop over an abstract graph:
//   holo-ospf/src/spf.rs:613-724   run_area   (OSPF)
//   holo-isis/src/spf.rs:544-706   compute_spt (IS-IS)
// The graph is the flattened CSR of include/holo_spf.h (vertex ids are
// numbered in the reference's VertexId order, so `(distance, id)` below is the
// reference's candidate-list key).  The data structures deliberately mirror
// the reference: an ordered map keyed (distance, VertexId) as candidate list
// with `pop_first`, a LINEAR scan of that map to find an existing candidate
// (spf.rs:681-685 "TODO: optimize lookup"), a per-edge mutual-link check that
// re-walks the neighbour's links (spf.rs:654-664), remove-and-recreate on a
// strictly better distance (spf.rs:686-703).  Because it executes the pops in
// order it is also exact for the order-dependent cases (zero-cost links,
// u16 saturation) that the device path refuses (HSPF_E_NEEDS_ORACLE).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference leg may load this library.
//
// Parity pinning: this restatement is checked against the reference's own
// golden topologies through oracle/spf_ospfv2.cc / spf_isis.cc (which share
// this loop's structure) in tests/test_oracle_golden.py.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <utility>
#include <vector>

#include "../include/holo_spf.h"

namespace {

struct OVertex {
    uint32_t id;
    uint32_t distance;
    uint32_t hops;
    std::vector<uint32_t> parents;      // IS-IS Vertex.parents (pop order, duplicates kept)
    std::vector<uint32_t> nh_vec;       // IS-IS Vertex.nexthops as atoms (Vec, duplicates kept)
    std::set<uint32_t> nh_set;          // OSPF Nexthops keys as atoms (BTreeMap => set)
    uint32_t first_parent = HSPF_NO_PARENT;
};

struct Link { uint32_t id; uint32_t cost; uint32_t edge; };

struct View {
    const hspf_csr *g;
    uint32_t n_ov;
    const uint32_t *ov_edge, *ov_cost;
    // Links of vertex u after overrides (a disabled edge is simply absent, as if
    // the LSA did not carry the link).
    template <typename F> void links(uint32_t u, F f) const {
        for (uint32_t e = g->row_ptr[u]; e < g->row_ptr[u + 1]; ++e) {
            uint32_t c = g->cost[e];
            for (uint32_t k = 0; k < n_ov; ++k) if (ov_edge[k] == e) c = ov_cost[k];
            if (c == HSPF_COST_DISABLED) continue;
            if (!f(Link{g->col[e], c, e})) return;
        }
    }
    bool linked_back(uint32_t from, uint32_t to) const {  // vertex_lsa_links(&link.lsa).any(|l| l.id == vertex.id)
        bool any = false;
        links(from, [&](const Link &l) { if (l.id == to) { any = true; return false; } return true; });
        return any;
    }
};

uint32_t atom_base(const hspf_csr *g, uint32_t root, uint32_t n) {
    // base of the atom block of non-HOP vertex n attached to root (see holo_spf.h)
    uint32_t rb = g->row_ptr[root], re = g->row_ptr[root + 1];
    uint32_t base = re - rb;
    for (uint32_t e = rb; e < re; ++e) {
        uint32_t h = g->col[e];
        if (g->vflags[h] & HSPF_VF_HOP) continue;
        if (h == n) return base;
        base += g->row_ptr[h + 1] - g->row_ptr[h];
    }
    return 0xFFFFFFFFu;
}

}  // namespace

extern "C" {

// Limits for the optional Vec-style outputs.
#define ORACLE_E_OVERFLOW (-100)

/*
 * Runs one SPF exactly as the reference would.
 *   vec_mode = 0: OSPF (nexthops are a set of atoms)
 *   vec_mode = 1: IS-IS (parents/nexthops are Vecs; the set outputs are the
 *                 de-duplicated Vecs)
 * Outputs ([V] unless noted; any may be NULL): dist, hops, first_parent,
 * n_parents, nh_mask [V][nhw]; parents_off [V+1] + parents [parents_cap]
 * (IS-IS parent lists in push order); nhvec_off [V+1] + nhvec [nhvec_cap].
 * pop_order [V] receives the vertices in pop order (n_popped written).
 */
int oracle_csr_spf(const hspf_csr *g, uint32_t root, uint32_t n_ov, const uint32_t *ov_edge,
                   const uint32_t *ov_cost, int vec_mode, uint32_t *dist, uint16_t *hops,
                   uint32_t *first_parent, uint16_t *n_parents, uint64_t *nh_mask, uint32_t nhw,
                   uint32_t *parents_off, uint32_t *parents, uint32_t parents_cap,
                   uint32_t *nhvec_off, uint32_t *nhvec, uint32_t nhvec_cap,
                   uint32_t *pop_order, uint32_t *n_popped, uint32_t *status) {
    const uint32_t V = g->n_vertices;
    View view{g, n_ov, ov_edge, ov_cost};
    uint32_t st = 0;

    std::map<uint32_t, OVertex> spt;                               // BTreeMap<VertexId, Vertex>
    std::map<std::pair<uint32_t, uint32_t>, OVertex> cand_list;    // BTreeMap<(distance, id), Vertex>
    std::vector<uint32_t> order;
    {
        OVertex r{};
        r.id = root; r.distance = 0; r.hops = 0;
        cand_list.emplace(std::make_pair(0u, root), std::move(r));
    }

    while (!cand_list.empty()) {
        // pop_first + spt.insert
        auto first = cand_list.begin();
        OVertex vertex = std::move(first->second);
        cand_list.erase(first);
        const uint32_t vid = vertex.id;
        order.push_back(vid);
        auto ins = spt.emplace(vid, std::move(vertex));
        const OVertex &vx = ins.first->second;

        // IS-IS transit gates (spf.rs:556-602) / none for OSPF
        const uint8_t fl = g->vflags[vid];
        if (fl & HSPF_VF_LEAF) continue;
        if ((fl & HSPF_VF_LEAF_UNLESS_ROOT) && vx.hops != 0) continue;

        view.links(vid, [&](const Link &link) {
            // mutual-link check
            if (!view.linked_back(link.id, vid)) return true;
            // already on the SPT?
            if (spt.count(link.id)) return true;
            // distance
            uint32_t distance;
            if (g->saturate_at) {   // OSPF: u16 saturating_add (spf.rs:672)
                uint64_t s = (uint64_t)vx.distance + link.cost;
                distance = s > g->saturate_at ? g->saturate_at : (uint32_t)s;
            } else {                // IS-IS: u32 saturating_add + max path metric (spf.rs:633-645)
                uint64_t s = (uint64_t)vx.distance + link.cost;
                distance = s > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s;
            }
            if (distance > g->reject_above) return true;
            // hops
            uint32_t hop = vx.hops;
            if (g->vflags[link.id] & HSPF_VF_HOP) hop = std::min<uint32_t>(hop + 1, 0xFFFF);
            // candidate lookup: linear scan like the reference
            auto it = cand_list.begin();
            for (; it != cand_list.end(); ++it) if (it->second.id == link.id) break;
            if (it != cand_list.end()) {
                if (distance < it->second.distance) {
                    cand_list.erase(it);
                } else if (distance > it->second.distance) {
                    return true;
                }
            }
            auto key = std::make_pair(distance, link.id);
            auto ce = cand_list.find(key);
            if (ce == cand_list.end()) {
                OVertex nv{};
                nv.id = link.id; nv.distance = distance; nv.hops = hop;
                nv.first_parent = vid;
                ce = cand_list.emplace(key, std::move(nv)).first;
            }
            OVertex &cand = ce->second;
            cand.parents.push_back(vid);
            // next hops
            if (vx.hops == 0) {
                const bool target_hop = g->vflags[link.id] & HSPF_VF_HOP;
                if (!((g->flags & HSPF_GF_NOHOP_TARGET_NO_NEXTHOP) && !target_hop)) {
                    uint32_t atom;
                    if (vid == root) atom = link.edge - g->row_ptr[root];
                    else {
                        uint32_t b = atom_base(g, root, vid);
                        atom = b == 0xFFFFFFFFu ? b : b + (link.edge - g->row_ptr[vid]);
                    }
                    if (atom == 0xFFFFFFFFu || atom >= 64u * nhw) st |= HSPF_JS_TOO_MANY_ATOMS;
                    else { cand.nh_set.insert(atom); cand.nh_vec.push_back(atom); }
                }
            } else {
                cand.nh_set.insert(vx.nh_set.begin(), vx.nh_set.end());
                if (vec_mode) {
                    if (cand.nh_vec.size() + vx.nh_vec.size() > (1u << 22)) st |= 0x80000000u;  // Vec blow-up guard
                    else cand.nh_vec.insert(cand.nh_vec.end(), vx.nh_vec.begin(), vx.nh_vec.end());
                }
            }
            return true;
        });
    }

    // ---- export ---------------------------------------------------------------
    for (uint32_t v = 0; v < V; ++v) {
        if (dist) dist[v] = HSPF_DIST_INF;
        if (hops) hops[v] = 0;
        if (first_parent) first_parent[v] = HSPF_NO_PARENT;
        if (n_parents) n_parents[v] = 0;
        if (nh_mask) for (uint32_t w = 0; w < nhw; ++w) nh_mask[(size_t)v * nhw + w] = 0;
    }
    for (auto &kv : spt) {
        const OVertex &x = kv.second;
        const uint32_t v = x.id;
        if (dist) dist[v] = x.distance;
        if (hops) hops[v] = (uint16_t)x.hops;
        if (first_parent) first_parent[v] = x.first_parent;
        if (n_parents) n_parents[v] = (uint16_t)std::min<size_t>(x.parents.size(), 0xFFFF);
        if (nh_mask) for (uint32_t a : x.nh_set) nh_mask[(size_t)v * nhw + (a >> 6)] |= 1ull << (a & 63);
        if (g->saturate_at && x.distance >= g->saturate_at) st |= HSPF_JS_SATURATED;
    }
    int rc = 0;
    if (parents_off) {
        uint32_t n = 0;
        for (uint32_t v = 0; v < V; ++v) {
            parents_off[v] = n;
            auto it = spt.find(v);
            if (it == spt.end()) continue;
            for (uint32_t p : it->second.parents) { if (parents && n < parents_cap) parents[n] = p; ++n; }
        }
        parents_off[V] = n;
        if (parents && n > parents_cap) rc = ORACLE_E_OVERFLOW;
    }
    if (nhvec_off) {
        uint32_t n = 0;
        for (uint32_t v = 0; v < V; ++v) {
            nhvec_off[v] = n;
            auto it = spt.find(v);
            if (it == spt.end()) continue;
            for (uint32_t a : it->second.nh_vec) { if (nhvec && n < nhvec_cap) nhvec[n] = a; ++n; }
        }
        nhvec_off[V] = n;
        if (nhvec && n > nhvec_cap) rc = ORACLE_E_OVERFLOW;
    }
    if (pop_order) for (size_t i = 0; i < order.size(); ++i) pop_order[i] = order[i];
    if (n_popped) *n_popped = (uint32_t)order.size();
    if (status) *status = st;
    return rc;
}

/*
 * Optimised CPU baseline: binary-heap Dijkstra over the
 * same CSR producing the same planes under the static-order rules
 * (requires: no zero-cost edge out of a HOP vertex; saturation is flagged).
 * It is NOT the reference algorithm's cost profile; it exists so that GPU
 * speed-ups are not inflated by the reference's O(E*frontier) scan, and to
 * check full-size batches quickly.  Cross-checked against oracle_csr_spf in
 * tests/test_oracle_csr.py.
 */
int oracle_csr_spf_heap(const hspf_csr *g, uint32_t root, uint32_t n_ov, const uint32_t *ov_edge,
                        const uint32_t *ov_cost, uint32_t *dist, uint16_t *hops, uint32_t *first_parent,
                        uint16_t *n_parents, uint64_t *nh_mask, uint32_t nhw, uint32_t *status) {
    const uint32_t V = g->n_vertices;
    View view{g, n_ov, ov_edge, ov_cost};
    uint32_t st = 0;
    std::vector<uint32_t> d(V, HSPF_DIST_INF), hp(V, 0), fp(V, HSPF_NO_PARENT), np(V, 0);
    std::vector<uint64_t> nh((size_t)V * nhw, 0);
    std::vector<uint8_t> done(V, 0);
    using Key = std::pair<uint32_t, uint32_t>;
    std::vector<Key> heap;
    auto cmp = [](const Key &a, const Key &b) { return a > b; };
    d[root] = 0;
    heap.push_back({0, root});
    // root's non-HOP heads -> atom bases
    std::vector<std::pair<uint32_t, uint32_t>> bases;
    {
        uint32_t rb = g->row_ptr[root], re = g->row_ptr[root + 1], base = re - rb;
        for (uint32_t e = rb; e < re; ++e) {
            uint32_t h = g->col[e];
            if (g->vflags[h] & HSPF_VF_HOP) continue;
            bases.push_back({h, base});
            base += g->row_ptr[h + 1] - g->row_ptr[h];
        }
    }
    while (!heap.empty()) {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        Key k = heap.back(); heap.pop_back();
        const uint32_t u = k.second;
        if (done[u] || k.first != d[u]) continue;
        done[u] = 1;
        const uint8_t fl = g->vflags[u];
        if (fl & HSPF_VF_LEAF) continue;
        if ((fl & HSPF_VF_LEAF_UNLESS_ROOT) && u != root) continue;
        uint32_t abase = 0; bool aok = true;
        if (hp[u] == 0 && u != root) {
            aok = false;
            for (auto &b : bases) if (b.first == u) { abase = b.second; aok = true; break; }
            if (!aok) st |= HSPF_JS_TOO_MANY_ATOMS;
        }
        view.links(u, [&](const Link &l) {
            const uint32_t v = l.id;
            if (done[v]) return true;
            uint64_t s = (uint64_t)d[u] + l.cost;
            uint32_t nd = s > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s;
            if (nd > g->reject_above) return true;
            if (nd > d[v]) return true;
            if (nd < d[v]) {
                d[v] = nd; np[v] = 0; fp[v] = u;
                hp[v] = hp[u] + ((g->vflags[v] & HSPF_VF_HOP) ? 1 : 0);
                for (uint32_t w = 0; w < nhw; ++w) nh[(size_t)v * nhw + w] = 0;
                heap.push_back({nd, v});
                std::push_heap(heap.begin(), heap.end(), cmp);
            }
            np[v]++;
            if (hp[u] == 0) {
                const bool th = g->vflags[v] & HSPF_VF_HOP;
                if (!((g->flags & HSPF_GF_NOHOP_TARGET_NO_NEXTHOP) && !th) && aok) {
                    uint32_t atom = abase + (l.edge - g->row_ptr[u]);
                    if (atom >= 64u * nhw) st |= HSPF_JS_TOO_MANY_ATOMS;
                    else nh[(size_t)v * nhw + (atom >> 6)] |= 1ull << (atom & 63);
                }
            } else {
                for (uint32_t w = 0; w < nhw; ++w) nh[(size_t)v * nhw + w] |= nh[(size_t)u * nhw + w];
            }
            return true;
        });
    }
    for (uint32_t v = 0; v < V; ++v) {
        if (d[v] != HSPF_DIST_INF && g->saturate_at && d[v] >= g->saturate_at) st |= HSPF_JS_SATURATED;
        if (dist) dist[v] = d[v];
        if (hops) hops[v] = (uint16_t)std::min<uint32_t>(hp[v], 0xFFFF);
        if (first_parent) first_parent[v] = d[v] == HSPF_DIST_INF ? HSPF_NO_PARENT : fp[v];
        if (n_parents) n_parents[v] = (uint16_t)std::min<uint32_t>(np[v], 0xFFFF);
    }
    if (nh_mask) std::memcpy(nh_mask, nh.data(), nh.size() * 8);
    if (status) *status = st;
    return 0;
}

}  // extern "C"
