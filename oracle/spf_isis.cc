// oracle/spf_isis.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Line-faithful CPU restatement of holo-isis' shortest-path-tree computation
// over the flat LSDB image of include/holo_lsdb.h:
//   compute_spt        holo-isis/src/spf.rs:525-707
//   Spt::insert        holo-isis/src/spf.rs:222-238 (first_hops / second_hops)
//   VertexId ordering  holo-isis/src/spf.rs:94-98, 299-315 (pseudonodes pop first)
//   vertex_edges       holo-isis/src/spf.rs:1005-1120
//   vertex_edge_cost   holo-isis/src/spf.rs:1122-1138
//   zeroth_lsp         holo-isis/src/spf.rs:1283-1294
// for `local = false` runs (next hops are VertexNexthop.system_id only; the
// interface/address resolution of `local = true`, spf.rs:948-1002, needs the
// adjacency arena and is not restated yet).  Ordered std::map stands in for
// BTreeMap; the candidate list keeps the linear lookup and the per-edge
// mutual-link re-iteration of the reference.
//
// Parity pinning: tests/test_oracle_golden.py checks the distances and first-hop
// system-id sets produced here against the reference's golden IS-IS local-ribs
// (tests/golden/isis.json).  Vertex.parents order and Vertex.hops are not exported
// by the reference's tests: "parity unpinned" beyond this restatement.
#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <tuple>
#include <utility>
#include <vector>

#include "../include/holo_lsdb.h"
#include "../include/holo_spf.h"

namespace {

constexpr uint32_t MAX_PATH_METRIC_STANDARD = 1023;
constexpr uint32_t MAX_PATH_METRIC_WIDE = 0xFE000000u;

inline bool is_pseudonode(hl_lan_id id) { return (id & 0xFF) != 0; }

struct VertexId {   // derived Ord: (non_pseudonode, lan_id); false < true
    bool non_pseudonode;
    hl_lan_id lan_id;
    bool operator<(const VertexId &o) const { return std::tie(non_pseudonode, lan_id) < std::tie(o.non_pseudonode, o.lan_id); }
    bool operator==(const VertexId &o) const { return non_pseudonode == o.non_pseudonode && lan_id == o.lan_id; }
};
inline VertexId vid(hl_lan_id id) { return VertexId{!is_pseudonode(id), id}; }

struct Vertex {
    VertexId id; uint32_t distance; uint16_t hops;
    std::vector<uint32_t> parents;      // arena indices == SPT insertion order
    std::vector<uint64_t> nexthops;     // VertexNexthop.system_id (48 bit)
};

struct Edge { VertexId id; uint32_t cost; };

struct Lsdb {
    const hl_isis_level *l;
    // fragments of a lan_id in LspId order: the image is sorted by (lan_id, fragment)
    std::pair<uint32_t, uint32_t> range(hl_lan_id id) const {
        uint32_t lo = 0, hi = l->n_lsps;
        while (lo < hi) { uint32_t m = (lo + hi) / 2; if (l->lsps[m].lan_id < id) lo = m + 1; else hi = m; }
        uint32_t b = lo;
        while (lo < l->n_lsps && l->lsps[lo].lan_id == id) ++lo;
        return {b, lo};
    }
    const hl_isis_lsp *zeroth(hl_lan_id id) const {
        auto r = range(id);
        for (uint32_t i = r.first; i < r.second; ++i) {
            const auto &p = l->lsps[i];
            if (p.fragment != 0) continue;
            if (p.seqno == 0 || p.rem_lifetime == 0) return nullptr;
            return &p;
        }
        return nullptr;
    }
    uint32_t edge_cost(hl_lan_id nbr, uint32_t metric) const {
        if (l->metric_mode == HL_ISIS_MODE_NORMAL) return metric;
        return is_pseudonode(nbr) ? 0 : 1;
    }
    // vertex_edges; f returns false to stop
    void vertex_edges(const VertexId &v, const std::function<bool(const Edge &)> &f) const {
        const bool mt_none = l->mt_id == HL_ISIS_MT_NONE;
        const bool mt_std = l->mt_id == HL_ISIS_MT_STANDARD;
        const bool std_en = l->metric_type == HL_ISIS_METRIC_STANDARD || l->metric_type == HL_ISIS_METRIC_BOTH;
        const bool wide_en = l->metric_type == HL_ISIS_METRIC_WIDE || l->metric_type == HL_ISIS_METRIC_BOTH;
        auto r = range(v.lan_id);
        for (uint32_t i = r.first; i < r.second; ++i) {
            const auto &lsp = l->lsps[i];
            if (lsp.seqno == 0 || lsp.rem_lifetime == 0) continue;
            const hl_isis_reach *re = l->reaches + lsp.reach_off;
            // standard_iter
            if ((mt_none || mt_std) && std_en)
                for (uint32_t k = 0; k < lsp.n_reach; ++k)
                    if (re[k].kind == HL_ISIS_REACH_LEGACY)
                        if (!f(Edge{vid(re[k].neighbor), edge_cost(re[k].neighbor, re[k].metric)})) return;
            // wide_iter
            if (((mt_none || mt_std) || is_pseudonode(lsp.lan_id)) && wide_en)
                for (uint32_t k = 0; k < lsp.n_reach; ++k)
                    if (re[k].kind == HL_ISIS_REACH_EXT && re[k].metric < MAX_PATH_METRIC_WIDE)
                        if (!f(Edge{vid(re[k].neighbor), edge_cost(re[k].neighbor, re[k].metric)})) return;
            // mt_iter
            if (!mt_none && !mt_std)
                for (uint32_t k = 0; k < lsp.n_reach; ++k)
                    if (re[k].kind == HL_ISIS_REACH_MT && re[k].mt_id == l->mt_id && re[k].metric < MAX_PATH_METRIC_WIDE)
                        if (!f(Edge{vid(re[k].neighbor), edge_cost(re[k].neighbor, re[k].metric)})) return;
            // mt_all_iter
            if (mt_none)
                for (uint32_t k = 0; k < lsp.n_reach; ++k)
                    if (re[k].kind == HL_ISIS_REACH_MT && re[k].metric < MAX_PATH_METRIC_WIDE)
                        if (!f(Edge{vid(re[k].neighbor), edge_cost(re[k].neighbor, re[k].metric)})) return;
        }
    }
};

}  // namespace

extern "C" int oracle_isis_compute_spt(const hl_isis_level *l, uint64_t root_system_id, hl_isis_spt *out) {
    Lsdb lsdb{l};
    std::vector<Vertex> arena;                       // Spt.arena (insertion == pop order)
    std::map<VertexId, uint32_t> id_tree;            // Spt.id_tree
    std::vector<uint32_t> first_hops, second_hops;
    std::map<std::pair<uint32_t, VertexId>, Vertex> cand_list;

    VertexId root_vid = vid((hl_lan_id)(root_system_id << 8));
    cand_list.emplace(std::make_pair(0u, root_vid), Vertex{root_vid, 0, 0, {}, {}});
    const uint32_t max_path_metric =
        l->metric_type == HL_ISIS_METRIC_STANDARD ? MAX_PATH_METRIC_STANDARD : MAX_PATH_METRIC_WIDE;
    bool overflow = false;

    while (!cand_list.empty()) {
        auto first = cand_list.begin();
        Vertex cand = std::move(first->second);
        cand_list.erase(first);
        // spt.insert
        const uint32_t vertex_idx = (uint32_t)arena.size();
        arena.push_back(std::move(cand));
        id_tree[arena[vertex_idx].id] = vertex_idx;
        if (!is_pseudonode(arena[vertex_idx].id.lan_id)) {
            if (arena[vertex_idx].hops == 1) first_hops.push_back(vertex_idx);
            if (arena[vertex_idx].hops == 2) second_hops.push_back(vertex_idx);
        }
        const VertexId vertex_id = arena[vertex_idx].id;
        const uint32_t vertex_distance = arena[vertex_idx].distance;
        const uint16_t vertex_hops = arena[vertex_idx].hops;

        const hl_isis_lsp *z = lsdb.zeroth(vertex_id.lan_id);
        if (!z) continue;
        const bool mt_some = l->mt_id != HL_ISIS_MT_NONE;
        if (vertex_hops != 0 && !is_pseudonode(z->lan_id) && mt_some) {
            bool ol = l->mt_id == HL_ISIS_MT_STANDARD ? (z->flags & HL_LSPF_OL) : (z->flags & HL_LSPF_MT_IPV6_OL);
            if (ol) continue;
        }
        if (mt_some && l->mt_id == HL_ISIS_MT_STANDARD && !is_pseudonode(z->lan_id)) {
            if (!(z->flags & HL_LSPF_HAS_PROTOCOLS)) continue;
            if (l->ipv4_enabled && !(z->flags & HL_LSPF_NLPID_IPV4)) continue;
            if (l->ipv6_enabled && !(z->flags & HL_LSPF_NLPID_IPV6)) continue;
        }

        lsdb.vertex_edges(vertex_id, [&](const Edge &link) {
            bool back = false;
            lsdb.vertex_edges(link.id, [&](const Edge &l2) { if (l2.id == vertex_id) { back = true; return false; } return true; });
            if (!back) return true;
            if (id_tree.count(link.id)) return true;
            uint64_t s = (uint64_t)vertex_distance + link.cost;
            uint32_t distance = s > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s;
            if (distance > max_path_metric) return true;
            uint16_t hops = vertex_hops;
            if (!is_pseudonode(link.id.lan_id)) hops = hops == 0xFFFF ? 0xFFFF : hops + 1;
            auto it = cand_list.begin();
            for (; it != cand_list.end(); ++it) if (it->second.id == link.id) break;
            if (it != cand_list.end()) {
                if (distance < it->second.distance) cand_list.erase(it);
                else if (distance > it->second.distance) return true;
            }
            auto key = std::make_pair(distance, link.id);
            auto ce = cand_list.find(key);
            if (ce == cand_list.end()) ce = cand_list.emplace(key, Vertex{link.id, distance, hops, {}, {}}).first;
            Vertex &cand_v = ce->second;
            cand_v.parents.push_back(vertex_idx);
            if (vertex_hops == 0) {
                if (!is_pseudonode(link.id.lan_id)) cand_v.nexthops.push_back(link.id.lan_id >> 8);
            } else {
                const auto &src = arena[vertex_idx].nexthops;
                if (cand_v.nexthops.size() + src.size() > (1u << 22)) overflow = true;
                else cand_v.nexthops.insert(cand_v.nexthops.end(), src.begin(), src.end());
            }
            return true;
        });
    }
    if (overflow) return -100;

    // ---- export in id_tree order -------------------------------------------------------
    std::vector<uint32_t> pos(arena.size());       // arena index -> position in id_tree order
    { uint32_t i = 0; for (auto &kv : id_tree) pos[kv.second] = i++; }
    uint32_t need_p = 0, need_n = 0;
    for (auto &v : arena) { need_p += (uint32_t)v.parents.size(); need_n += (uint32_t)v.nexthops.size(); }
    out->n_vertices = (uint32_t)arena.size();
    out->n_parents = need_p; out->n_nexthops = need_n;
    out->n_first_hops = (uint32_t)first_hops.size(); out->n_second_hops = (uint32_t)second_hops.size();
    if (out->n_vertices > out->vertices_cap || need_p > out->parents_cap || need_n > out->nexthops_cap ||
        out->n_first_hops > out->first_hops_cap || out->n_second_hops > out->second_hops_cap)
        return HSPF_E_NOMEM;
    uint32_t i = 0, p = 0, n = 0;
    for (auto &kv : id_tree) {
        const Vertex &v = arena[kv.second];
        hl_isis_vertex o{};
        o.lan_id = v.id.lan_id; o.distance = v.distance; o.hops = v.hops;
        o.par_off = p; o.n_par = (uint32_t)v.parents.size();
        o.nh_off = n; o.n_nh = (uint32_t)v.nexthops.size();
        for (uint32_t x : v.parents) out->parents[p++] = pos[x];
        for (uint64_t x : v.nexthops) out->nexthops[n++] = x;
        out->vertices[i++] = o;
    }
    for (uint32_t k = 0; k < first_hops.size(); ++k) out->first_hops[k] = pos[first_hops[k]];
    for (uint32_t k = 0; k < second_hops.size(); ++k) out->second_hops[k] = pos[second_hops[k]];
    return 0;
}
