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
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <set>
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
    // iter_for_system_id (collections.rs:670-678): every LSP of a system, any pseudonode number
    std::pair<uint32_t, uint32_t> range_system(uint64_t system_id) const {
        const hl_lan_id first = (hl_lan_id)(system_id << 8), last = first | 0xFF;
        uint32_t lo = 0, hi = l->n_lsps;
        while (lo < hi) { uint32_t m = (lo + hi) / 2; if (l->lsps[m].lan_id < first) lo = m + 1; else hi = m; }
        uint32_t b = lo;
        while (lo < l->n_lsps && l->lsps[lo].lan_id <= last) ++lo;
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

// =====================================================================================
// compute_spf's route path: compute_spt(local = true) per enabled topology with
// resolve_nexthop (holo-isis/src/spf.rs:948-1002), then compute_routes
// (spf.rs:838-941) with vertex_networks (spf.rs:1141-1281), Route::new /
// merge_nexthops / build_nexthops (holo-isis/src/route.rs:79-139) and the max_paths cut.
// SR Prefix-SID labels: prefix_sid_update and its helpers (holo-isis/src/sr.rs:33-99, 151-300);
// PARITY UNPINNED — no golden of the reference carries an IS-IS Prefix-SID.
// =====================================================================================
namespace {

struct LocalNexthop { uint64_t system_id; bool has_iface; uint32_t iface; bool has4; uint32_t ipv4; bool has6; hl_ip_addr ipv6; };
struct LVertex { VertexId id; uint32_t distance; uint16_t hops; std::vector<LocalNexthop> nexthops; };

struct SnpaLess { bool operator()(const std::array<uint8_t, 6> &a, const std::array<uint8_t, 6> &b) const { return a < b; } };

struct IpKey {
    hl_ip_addr a;
    bool operator<(const IpKey &o) const {
        if (a.is_v6 != o.a.is_v6) return a.is_v6 < o.a.is_v6;
        return std::memcmp(a.bytes, o.a.bytes, 16) < 0;
    }
};
struct NetKey {
    hl_ip_addr a; uint8_t len;
    bool operator<(const NetKey &o) const {
        if (a.is_v6 != o.a.is_v6) return a.is_v6 < o.a.is_v6;
        int c = std::memcmp(a.bytes, o.a.bytes, 16);
        if (c) return c < 0;
        return len < o.len;
    }
};
struct Psid { uint8_t flags; bool is_label; uint32_t value; };   // PrefixSidStlv (packet/subtlvs/prefix.rs:69-73)
struct RNexthop { uint64_t system_id; uint32_t iface; hl_ip_addr addr; bool has_label = false; uint32_t label = 0; };
struct Route {
    uint8_t route_type; uint32_t metric; uint8_t flags; std::map<IpKey, RNexthop> nexthops;
    bool has_psid = false; Psid psid{};          // Route.prefix_sid
    bool has_label = false; uint32_t label = 0;  // Route.sr_label
};
struct VNet { hl_ip_addr prefix; uint8_t len; uint32_t metric; bool external; bool has_psid = false; Psid psid{}; };

// ---- holo-isis/src/sr.rs:33-99, 151-300 --------------------------------------------------
// index_to_label (sr.rs:268-300): walk the label blocks; Err = no label
bool index_to_label(const hl_isis_level &l, const hl_isis_lsp &cap, uint32_t index, uint32_t &label) {
    for (uint32_t i = 0; i < cap.n_srgb; ++i) {
        const hl_srgb &b = l.srgbs[cap.srgb_off + i];
        if (b.first_is_index) continue;               // "SID ranges are rather obscure"
        if (index >= b.range) { index -= b.range; continue; }
        label = b.first + index;
        return true;
    }
    return false;
}

// first valid LSP of the system that carries an SR-Capabilities sub-TLV (sr.rs:176-184, 213-221)
const hl_isis_lsp *sr_cap_of(const Lsdb &lsdb, uint64_t system_id) {
    auto r = lsdb.range_system(system_id);
    for (uint32_t i = r.first; i < r.second; ++i) {
        const auto &p = lsdb.l->lsps[i];
        if (p.rem_lifetime == 0 || p.seqno == 0) continue;
        if (p.sr_flags & HL_LSP_SR_HAS_CAP) return &p;
    }
    return nullptr;
}

// prefix_sid_update (sr.rs:33-99)
void prefix_sid_update(const hl_isis_instance *in, const Lsdb &lsdb, hl_lan_id adv_rtr, bool is_v6, Route &route,
                       bool local, bool last_hop) {
    if (!route.has_psid) return;
    const Psid &ps = route.psid;
    // the advertising node must list SPF in an SR-Algorithm sub-TLV (sr.rs:49-62)
    bool algo = false;
    auto r = lsdb.range(adv_rtr);
    for (uint32_t i = r.first; i < r.second && !algo; ++i) {
        const auto &p = lsdb.l->lsps[i];
        if (p.rem_lifetime == 0 || p.seqno == 0) continue;
        if (p.sr_flags & HL_LSP_SR_ALGO_SPF) algo = true;
    }
    if (!algo) return;
    // input label (prefix_sid_input_label, sr.rs:151-195); Err leaves route.sr_label as it is
    if (local && (!(ps.flags & HL_ISIS_PSID_P) || (ps.flags & HL_ISIS_PSID_E))) {
        route.has_label = false;                                     // Ok(None)
    } else if (!ps.is_label) {
        const hl_isis_lsp *cap = sr_cap_of(lsdb, in->system_id);
        uint32_t label = 0;
        if (cap && index_to_label(*lsdb.l, *cap, ps.value, label)) { route.has_label = true; route.label = label; }
    } else {
        route.has_label = true; route.label = ps.value;
    }
    // output labels (prefix_sid_output_label, sr.rs:198-265); Err leaves the nexthop's label as it is
    for (auto &kv : route.nexthops) {
        RNexthop &nh = kv.second;
        if (last_hop && !(ps.flags & HL_ISIS_PSID_P)) { nh.has_label = true; nh.label = 3; continue; }   // implicit null
        const hl_isis_lsp *cap = sr_cap_of(lsdb, nh.system_id);
        if (!cap) continue;                                                                           // SrCapNotFound
        if (!(cap->sr_flags & (is_v6 ? HL_LSP_SR_CAP_V : HL_LSP_SR_CAP_I))) continue;                // SrCapUnsupportedAf
        if (last_hop && (ps.flags & HL_ISIS_PSID_E)) { nh.has_label = true; nh.label = is_v6 ? 2 : 0; continue; }   // explicit null
        if (!ps.is_label) {
            uint32_t label = 0;
            if (index_to_label(*lsdb.l, *cap, ps.value, label)) { nh.has_label = true; nh.label = label; }
        } else {
            // V/L SIDs have local significance: only adjacent routers can use them
            nh.has_label = true; nh.label = last_hop ? ps.value : 3;
        }
    }
}

// one compute_spt(local = true) run; returns vertices in id_tree order
std::vector<LVertex> local_spt(const hl_isis_instance *in, uint8_t mt_id) {
    hl_isis_level l = in->lvl;
    l.mt_id = mt_id;
    l.metric_mode = HL_ISIS_MODE_NORMAL;
    Lsdb lsdb{&l};
    std::vector<LVertex> arena;
    std::map<VertexId, uint32_t> id_tree;
    std::map<std::pair<uint32_t, VertexId>, LVertex> cand_list;
    std::set<std::array<uint8_t, 6>> used_adjs;
    VertexId root_vid = vid((hl_lan_id)(in->system_id << 8));
    cand_list.emplace(std::make_pair(0u, root_vid), LVertex{root_vid, 0, 0, {}});
    const uint32_t max_path_metric = l.metric_type == HL_ISIS_METRIC_STANDARD ? MAX_PATH_METRIC_STANDARD : MAX_PATH_METRIC_WIDE;
    const uint8_t level_bit = in->level == 1 ? 1 : 2;

    auto resolve_nexthop = [&](LocalNexthop &nh, const LVertex &vertex, const Edge &link) {
        const bool want_bcast = is_pseudonode(vertex.id.lan_id);
        for (uint32_t i = 0; i < in->n_ifaces; ++i) {
            const auto &iface = in->ifaces[i];
            if ((bool)iface.is_broadcast != want_bcast) continue;
            const hl_isis_adj *adj = nullptr;
            if (iface.is_broadcast) {
                for (uint32_t k = 0; k < iface.n_adj; ++k) {
                    const auto &a = in->adjs[iface.adj_off + k];
                    if (a.system_id == (link.id.lan_id >> 8)) { adj = &a; break; }   // get_by_system_id
                }
                if (adj && !(mt_id == HL_ISIS_MT_STANDARD ? adj->topo_std : adj->topo_ipv6)) adj = nullptr;
                if (adj && !adj->up) adj = nullptr;
            } else {
                if (iface.metric != link.cost) continue;
                if (iface.n_adj) {
                    const auto &a = in->adjs[iface.adj_off];
                    if ((mt_id == HL_ISIS_MT_STANDARD ? a.topo_std : a.topo_ipv6) && (a.level_usage & level_bit) &&
                        a.system_id == (link.id.lan_id >> 8) && a.up)
                        adj = &a;
                }
            }
            if (!adj) continue;
            std::array<uint8_t, 6> snpa;
            std::memcpy(snpa.data(), adj->snpa, 6);
            if (!used_adjs.insert(snpa).second) continue;     // .find(|..| used_adjs.insert(adj.snpa))
            nh.has_iface = true; nh.iface = i;
            nh.has4 = adj->has_ipv4; nh.ipv4 = adj->ipv4;
            nh.has6 = adj->has_ipv6; nh.ipv6 = adj->ipv6;
            return;
        }
    };

    while (!cand_list.empty()) {
        auto first = cand_list.begin();
        LVertex cand = std::move(first->second);
        cand_list.erase(first);
        const uint32_t vertex_idx = (uint32_t)arena.size();
        arena.push_back(std::move(cand));
        id_tree[arena[vertex_idx].id] = vertex_idx;
        const VertexId vertex_id = arena[vertex_idx].id;
        const uint32_t vertex_distance = arena[vertex_idx].distance;
        const uint16_t vertex_hops = arena[vertex_idx].hops;
        const hl_isis_lsp *z = lsdb.zeroth(vertex_id.lan_id);
        if (!z) continue;
        if (vertex_hops != 0 && !is_pseudonode(z->lan_id)) {
            bool ol = mt_id == HL_ISIS_MT_STANDARD ? (z->flags & HL_LSPF_OL) : (z->flags & HL_LSPF_MT_IPV6_OL);
            if (ol) continue;
        }
        if (mt_id == HL_ISIS_MT_STANDARD && !is_pseudonode(z->lan_id)) {
            if (!(z->flags & HL_LSPF_HAS_PROTOCOLS)) continue;
            if (l.ipv4_enabled && !(z->flags & HL_LSPF_NLPID_IPV4)) continue;
            if (l.ipv6_enabled && !(z->flags & HL_LSPF_NLPID_IPV6)) continue;
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
            if (ce == cand_list.end()) ce = cand_list.emplace(key, LVertex{link.id, distance, hops, {}}).first;
            LVertex &cand_v = ce->second;
            if (vertex_hops == 0) {
                if (!is_pseudonode(link.id.lan_id)) {
                    LocalNexthop nh{link.id.lan_id >> 8, false, 0, false, 0, false, hl_ip_addr{}};
                    resolve_nexthop(nh, arena[vertex_idx], link);
                    cand_v.nexthops.push_back(nh);
                }
            } else {
                const auto &src = arena[vertex_idx].nexthops;
                cand_v.nexthops.insert(cand_v.nexthops.end(), src.begin(), src.end());
            }
            return true;
        });
    }
    std::vector<LVertex> out;
    for (auto &kv : id_tree) out.push_back(arena[kv.second]);
    return out;
}

}  // namespace

extern "C" int oracle_isis_compute_routes(const hl_isis_instance *in, hl_isis_rib *out) {
    const hl_isis_level &l = in->lvl;
    const bool std_en = l.metric_type == HL_ISIS_METRIC_STANDARD || l.metric_type == HL_ISIS_METRIC_BOTH;
    const bool wide_en = l.metric_type == HL_ISIS_METRIC_WIDE || l.metric_type == HL_ISIS_METRIC_BOTH;
    std::map<NetKey, Route> rib;
    const uint8_t mts[2] = {HL_ISIS_MT_STANDARD, HL_ISIS_MT_IPV6};
    for (uint8_t mt_id : mts) {
        if (mt_id == HL_ISIS_MT_IPV6 && !in->mt_ipv6_enabled) continue;
        std::vector<LVertex> spt = local_spt(in, mt_id);
        hl_isis_level lm = l;
        lm.mt_id = mt_id;
        Lsdb lsdb{&lm};
        // is_l2_attached_to_backbone (instance.rs:575-589)
        bool attached = false;
        for (uint32_t i = 0; i < in->n_adjs; ++i) {
            const auto &a = in->adjs[i];
            if ((mt_id == HL_ISIS_MT_STANDARD ? a.topo_std : a.topo_ipv6) && a.up && (a.level_usage & 2) && a.area_disjoint) attached = true;
        }
        const bool ipv4_enabled = l.ipv4_enabled && mt_id == HL_ISIS_MT_STANDARD;
        const bool ipv6_enabled = l.ipv6_enabled && (mt_id == HL_ISIS_MT_STANDARD ? !in->mt_ipv6_enabled : true);
        for (const LVertex &vertex : spt) {
            const hl_isis_lsp *z = lsdb.zeroth(vertex.id.lan_id);
            if (!z) continue;
            const bool att_bit = !in->att_ignore && (mt_id == HL_ISIS_MT_STANDARD ? (z->flags & HL_LSPF_ATT) : (z->flags & HL_LSPF_MT_IPV6_ATT));
            // vertex_networks: per valid fragment, in LspId order
            std::vector<VNet> nets;
            auto r = lsdb.range(vertex.id.lan_id);
            for (uint32_t i = r.first; i < r.second; ++i) {
                const auto &lsp = l.lsps[i];
                if (lsp.seqno == 0 || lsp.rem_lifetime == 0) continue;
                if (att_bit && in->level == 1 && (in->level_type == 1 || !attached)) {
                    if (ipv4_enabled) nets.push_back(VNet{hl_ip_addr{}, 0, 0, false});
                    if (ipv6_enabled) { hl_ip_addr z6{}; z6.is_v6 = 1; nets.push_back(VNet{z6, 0, 0, false}); }
                }
                const hl_isis_ipreach *ip = l.ipreaches + lsp.ipreach_off;
                if (mt_id == HL_ISIS_MT_STANDARD && ipv4_enabled) {
                    if (std_en) {
                        for (uint32_t k = 0; k < lsp.n_ipreach; ++k)
                            if (ip[k].kind == HL_ISIS_IP_V4_INTERNAL) nets.push_back(VNet{ip[k].prefix, ip[k].len, ip[k].metric, false});
                        for (uint32_t k = 0; k < lsp.n_ipreach; ++k)
                            if (ip[k].kind == HL_ISIS_IP_V4_EXTERNAL) nets.push_back(VNet{ip[k].prefix, ip[k].len, ip[k].metric, true});
                    }
                    if (wide_en)
                        for (uint32_t k = 0; k < lsp.n_ipreach; ++k)
                            if (ip[k].kind == HL_ISIS_IP_V4_EXT && ip[k].metric <= MAX_PATH_METRIC_WIDE)
                                nets.push_back(VNet{ip[k].prefix, ip[k].len, ip[k].metric, (bool)ip[k].external, (bool)ip[k].has_psid,
                                                    Psid{ip[k].psid_flags, (bool)ip[k].psid_is_label, ip[k].psid_value}});
                }
                if (ipv6_enabled) {
                    for (uint32_t k = 0; k < lsp.n_ipreach; ++k) {
                        const bool take = mt_id == HL_ISIS_MT_IPV6 ? (ip[k].kind == HL_ISIS_IP_MT_V6 && ip[k].mt_id == HL_ISIS_MT_IPV6)
                                                                   : (ip[k].kind == HL_ISIS_IP_V6);
                        if (take) nets.push_back(VNet{ip[k].prefix, ip[k].len, ip[k].metric, (bool)ip[k].external, (bool)ip[k].has_psid,
                                                      Psid{ip[k].psid_flags, (bool)ip[k].psid_is_label, ip[k].psid_value}});
                    }
                }
            }
            for (const VNet &network : nets) {
                auto build = [&]() {
                    std::map<IpKey, RNexthop> m;
                    for (const auto &nh : vertex.nexthops) {
                        hl_ip_addr addr{};
                        if (!network.prefix.is_v6) {
                            if (!nh.has4) continue;
                            addr.bytes[0] = nh.ipv4 >> 24; addr.bytes[1] = nh.ipv4 >> 16; addr.bytes[2] = nh.ipv4 >> 8; addr.bytes[3] = nh.ipv4;
                        } else {
                            if (!nh.has6) continue;
                            addr = nh.ipv6; addr.is_v6 = 1;
                        }
                        m[IpKey{addr}] = RNexthop{nh.system_id, nh.iface, addr};   // iface_idx.unwrap()
                    }
                    return m;
                };
                auto mk = [&]() {
                    Route rt;
                    rt.flags = vertex.hops == 0 ? HL_ROUTE_CONNECTED : 0;
                    rt.route_type = in->level == 1 ? (network.external ? HL_ISIS_RT_L1_EXT : HL_ISIS_RT_L1_INTRA)
                                                   : (network.external ? HL_ISIS_RT_L2_EXT : HL_ISIS_RT_L2_INTRA);
                    rt.metric = vertex.distance + network.metric;
                    rt.nexthops = build();
                    rt.has_psid = network.has_psid; rt.psid = network.psid;      // route.rs:100
                    return rt;
                };
                NetKey key{network.prefix, network.len};
                auto it = rib.find(key);
                Route *route;
                if (it == rib.end()) {
                    route = &rib.emplace(key, mk()).first->second;
                } else {
                    Route &cur = it->second;
                    const uint32_t route_metric = vertex.distance + network.metric;
                    if (route_metric < cur.metric) cur = mk();
                    else if (route_metric == cur.metric) { for (auto &kv : build()) cur.nexthops[kv.first] = kv.second; }
                    else continue;
                    route = &cur;
                }
                if (route->nexthops.size() > in->max_paths) {
                    std::map<IpKey, RNexthop> cut; uint32_t n = 0;
                    for (auto &kv : route->nexthops) { if (n++ >= in->max_paths) break; cut.insert(kv); }
                    route->nexthops = std::move(cut);
                }
                // Update route's Prefix-SID (spf.rs:923-939)
                if (in->sr_enabled && route->has_psid)
                    prefix_sid_update(in, lsdb, vertex.id.lan_id, network.prefix.is_v6, *route, vertex.hops == 0, vertex.hops == 1);
            }
        }
    }
    uint32_t need_h = 0;
    for (auto &kv : rib) need_h += (uint32_t)kv.second.nexthops.size();
    out->n_routes = (uint32_t)rib.size(); out->n_nexthops = need_h;
    if (out->n_routes > out->routes_cap || need_h > out->nexthops_cap) return HSPF_E_NOMEM;
    uint32_t i = 0, h = 0;
    for (auto &kv : rib) {
        hl_isis_route o{};
        o.prefix = kv.first.a; o.len = kv.first.len; o.metric = kv.second.metric; o.route_type = kv.second.route_type;
        o.flags = kv.second.flags; o.nh_off = h; o.n_nh = (uint32_t)kv.second.nexthops.size();
        o.has_sr_label = kv.second.has_label; o.sr_label = kv.second.has_label ? kv.second.label : 0;
        for (auto &nk : kv.second.nexthops) {
            hl_isis_nexthop x{};
            x.system_id = nk.second.system_id; x.iface = nk.second.iface; x.addr = nk.second.addr;
            x.has_label = nk.second.has_label; x.sr_label = nk.second.has_label ? nk.second.label : 0;
            out->nexthops[h++] = x;
        }
        out->routes[i++] = o;
    }
    return 0;
}

// Restates the SPF-type decision of lsp_install (holo-isis/src/lsdb.rs:1450-1465, 1525-1531): the run is Full as
// soon as one installed LSP has no previous instance, or differs from it in `is_expired()`, `flags`, or in the
// `is_reach()` / `ext_is_reach()` iterators (TLV 2 and TLV 22 entries in TLV order; TLV 222 is not looked at);
// otherwise RouteOnly.  PARITY UNPINNED: no reference fixture records the SPF type.
extern "C" int oracle_isis_spf_type(const hl_isis_level *old_lvl, const hl_isis_level *new_lvl,
                                    const hl_isis_lsp_trigger *tr, uint32_t n, uint32_t *spf_type) {
    auto lookup = [](const hl_isis_level *l, uint64_t lan_id, uint8_t fragment) {
        int at = -1;
        for (uint32_t i = 0; i < l->n_lsps && at < 0; ++i)
            if (l->lsps[i].lan_id == lan_id && l->lsps[i].fragment == fragment) at = (int)i;
        return at;
    };
    auto reach_of = [](const hl_isis_level *l, int at, uint8_t kind) {
        std::vector<std::pair<uint64_t, uint32_t>> v;
        const hl_isis_lsp &p = l->lsps[at];
        for (uint32_t k = 0; k < p.n_reach; ++k) {
            const hl_isis_reach &r = l->reaches[p.reach_off + k];
            if (r.kind == kind) v.emplace_back(r.neighbor, r.metric);
        }
        return v;
    };
    bool full = false;
    for (uint32_t k = 0; k < n; ++k) {
        const int o = lookup(old_lvl, tr[k].lan_id, tr[k].fragment), w = lookup(new_lvl, tr[k].lan_id, tr[k].fragment);
        if (w < 0) return -1;
        bool topology_change = true;
        if (o >= 0) {
            const hl_isis_lsp &a = old_lvl->lsps[o], &b = new_lvl->lsps[w];
            const bool a_expired = a.rem_lifetime == 0, b_expired = b.rem_lifetime == 0;
            // LspFlags bits present in the image: overload and attached
            const bool same_flags = ((a.flags ^ b.flags) & (HL_LSPF_OL | HL_LSPF_ATT)) == 0;
            if (a_expired == b_expired && same_flags &&
                reach_of(old_lvl, o, HL_ISIS_REACH_LEGACY) == reach_of(new_lvl, w, HL_ISIS_REACH_LEGACY) &&
                reach_of(old_lvl, o, HL_ISIS_REACH_EXT) == reach_of(new_lvl, w, HL_ISIS_REACH_EXT))
                topology_change = false;
        }
        full = full || topology_change;
    }
    *spf_type = full ? HL_ISIS_SPF_FULL : HL_ISIS_SPF_ROUTE_ONLY;
    return 0;
}
