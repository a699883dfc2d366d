// TEST INFRASTRUCTURE — CPU restatement of the end of holo-isis' update_rib: the merge of the
// per-level tables (holo-isis/src/route.rs:236-242: rib_l2.iter().chain(rib_l1.iter()).collect(),
// i.e. an L1 route replaces the L2 route of the same prefix) and update_global_rib
// (route.rs:255-314).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use
// anything under oracle/.
//
// Pinned by: the RouteIpAdd streams of the reference's 38 IS-IS conformance snapshots
// (output/ibus.jsonl, final state per prefix): computing both levels, merging and diffing against
// an empty table gives exactly those routes with metric, ifindex and next-hop address
// (tests/test_isis_rib.py); summary (blackhole) routes by the nb-config-summary step tests
// (tests/test_isis_l1l2.py).  Tags are not modelled.
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "../include/holo_lsdb.h"
#include "../include/holo_spf.h"

namespace {

struct Key {
    hl_ip_addr a; uint8_t len;
    bool operator<(const Key &o) const {
        if (a.is_v6 != o.a.is_v6) return a.is_v6 < o.a.is_v6;            // IpNetwork: V4 < V6
        const int c = std::memcmp(a.bytes, o.a.bytes, 16);
        return c ? c < 0 : len < o.len;
    }
};
struct Ref { const hl_isis_rib *rib; uint32_t idx; };

bool nh_eq(const hl_isis_nexthop &a, const hl_isis_nexthop &b) {
    return a.system_id == b.system_id && a.iface == b.iface && a.addr.is_v6 == b.addr.is_v6 &&
           std::memcmp(a.addr.bytes, b.addr.bytes, 16) == 0 && a.has_label == b.has_label &&
           (!a.has_label || a.sr_label == b.sr_label);
}

}  // namespace

extern "C" int oracle_isis_rib_merge(const hl_isis_rib *l2, const hl_isis_rib *l1, hl_isis_rib *out) {
    std::map<Key, Ref> merged;                       // collect(): later entries replace earlier ones
    for (const hl_isis_rib *r : {l2, l1}) {
        if (!r) continue;
        for (uint32_t i = 0; i < r->n_routes; ++i) merged[Key{r->routes[i].prefix, r->routes[i].len}] = Ref{r, i};
    }
    uint32_t n_h = 0;
    for (auto &kv : merged) n_h += kv.second.rib->routes[kv.second.idx].n_nh;
    out->n_routes = (uint32_t)merged.size();
    out->n_nexthops = n_h;
    if (out->n_routes > out->routes_cap || n_h > out->nexthops_cap) return HSPF_E_NOMEM;
    uint32_t i = 0, h = 0;
    for (auto &kv : merged) {
        hl_isis_route o = kv.second.rib->routes[kv.second.idx];
        const hl_isis_nexthop *from = kv.second.rib->nexthops + o.nh_off;
        o.nh_off = h;
        for (uint32_t k = 0; k < o.n_nh; ++k) out->nexthops[h++] = from[k];
        out->routes[i++] = o;
    }
    return HSPF_OK;
}

extern "C" int oracle_isis_rib_diff(const hl_isis_rib *old_rib, hl_isis_rib *rib, hl_rib_action *out, uint32_t cap,
                                    uint32_t *n_out) {
    std::map<Key, uint32_t> old;
    if (old_rib)
        for (uint32_t i = 0; i < old_rib->n_routes; ++i) old[Key{old_rib->routes[i].prefix, old_rib->routes[i].len}] = i;
    std::vector<hl_rib_action> acts;
    for (uint32_t i = 0; i < rib->n_routes; ++i) {
        hl_isis_route &route = rib->routes[i];
        hl_rib_action a{};
        a.route = i;
        auto it = old.find(Key{route.prefix, route.len});
        if (it != old.end()) {
            const hl_isis_route &o = old_rib->routes[it->second];
            old.erase(it);
            a.has_old_sr_label = o.has_sr_label; a.old_sr_label = o.has_sr_label ? o.sr_label : 0;
            bool same = o.metric == route.metric && o.n_nh == route.n_nh;     // && tag (None on both sides)
            for (uint32_t k = 0; same && k < route.n_nh; ++k)
                same = nh_eq(old_rib->nexthops[o.nh_off + k], rib->nexthops[route.nh_off + k]);
            if (same) {
                if (o.flags & HL_ROUTE_INSTALLED) route.flags |= HL_ROUTE_INSTALLED;
                continue;
            }
        }
        if (!(route.flags & HL_ROUTE_CONNECTED) && ((route.flags & HL_ROUTE_SUMMARY) || route.n_nh != 0)) {   // route.rs:284-288
            a.kind = HL_RIB_INSTALL;
            acts.push_back(a);
            route.flags |= HL_ROUTE_INSTALLED;
        } else if (route.flags & HL_ROUTE_INSTALLED) {
            a.kind = HL_RIB_UNINSTALL; a.has_old_sr_label = 0; a.old_sr_label = 0;
            acts.push_back(a);
            route.flags &= (uint8_t)~HL_ROUTE_INSTALLED;
        }
    }
    for (auto &kv : old) {
        if (!(old_rib->routes[kv.second].flags & HL_ROUTE_INSTALLED)) continue;
        hl_rib_action a{};
        a.kind = HL_RIB_UNINSTALL_OLD; a.route = kv.second;
        acts.push_back(a);
    }
    *n_out = (uint32_t)acts.size();
    if (acts.size() > cap) return HSPF_E_NOMEM;
    for (size_t i = 0; i < acts.size(); ++i) out[i] = acts[i];
    return HSPF_OK;
}

// ---- L1/L2 routers: summary routes and L1 -> L2 propagation ----------------------------------
// Restates holo-isis/src/route.rs:189-231 (update_rib: the summaries of the L1 table, their
// routes in the L2 table) and lsdb.rs:1149-1357 (lsp_propagate_l1_to_l2, propagate_ip_reach) with
// the reference's own containers (BTreeMap -> std::map).  JointPrefixMap::get_spm is the
// prefix-trie crate's shortest-prefix match (third-party; restated from its documentation:
// "get the shortest prefix in the map that contains the given prefix").
// Pinned by: the own L2 LSPs of the L1/L2 routers in the reference's IS-IS conformance snapshots
// (their propagated entries = this function's output) and the nb-config-summary1/2 step tests
// (tests/test_isis_l1l2.py).
#include <algorithm>

namespace {

bool pfx_contains(const hl_ip_addr &net, uint8_t nlen, const hl_ip_addr &a, uint8_t alen) {
    if (net.is_v6 != a.is_v6 || nlen > alen) return false;
    for (int i = 0; i < nlen; ++i)
        if (((net.bytes[i / 8] >> (7 - i % 8)) & 1) != ((a.bytes[i / 8] >> (7 - i % 8)) & 1)) return false;
    return true;
}

const hl_isis_summary *get_spm(const hl_isis_summary *cfg, uint32_t n, const hl_ip_addr &a, uint8_t len) {
    const hl_isis_summary *best = nullptr;
    for (uint32_t i = 0; i < n; ++i)
        if (pfx_contains(cfg[i].prefix, cfg[i].len, a, len) && (!best || cfg[i].len < best->len)) best = &cfg[i];
    return best;
}

}  // namespace

extern "C" int oracle_isis_summaries(const hl_isis_rib *l1, const hl_isis_summary *cfg, uint32_t n_cfg,
                                     hl_isis_summary *out, uint32_t *n_out) {
    std::map<Key, hl_isis_summary> summaries;               // instance.state.summaries
    for (uint32_t i = 0; l1 && i < l1->n_routes; ++i) {
        const hl_isis_route &route = l1->routes[i];
        const hl_isis_summary *c = get_spm(cfg, n_cfg, route.prefix, route.len);
        if (!c) continue;
        Key k{c->prefix, c->len};
        auto it = summaries.find(k);
        if (it != summaries.end()) it->second.metric = std::min(it->second.metric, route.metric);
        else { hl_isis_summary s = *c; s.metric = route.metric; summaries.emplace(k, s); }
    }
    uint32_t n = 0;
    for (auto &kv : summaries) out[n++] = kv.second;
    *n_out = n;
    return 0;
}

extern "C" int oracle_isis_rib_add_summaries(const hl_isis_rib *l2, const hl_isis_summary *active, uint32_t n_active,
                                             hl_isis_rib *out) {
    struct Ent { bool summary; uint32_t idx; };
    std::map<Key, Ent> rib;
    for (uint32_t i = 0; l2 && i < l2->n_routes; ++i) rib[Key{l2->routes[i].prefix, l2->routes[i].len}] = Ent{false, i};
    for (uint32_t j = 0; j < n_active; ++j) rib[Key{active[j].prefix, active[j].len}] = Ent{true, j};     // extend()
    uint32_t r = 0, h = 0;
    for (auto &kv : rib) {
        hl_isis_route o;
        if (kv.second.summary) {
            const hl_isis_summary &s = active[kv.second.idx];
            std::memset(&o, 0, sizeof(o));
            o.prefix = s.prefix; o.len = s.len;
            o.metric = s.has_cfg_metric ? s.cfg_metric : s.metric;     // SummaryRoute::metric
            o.route_type = HL_ISIS_RT_L2_INTRA; o.flags = HL_ROUTE_SUMMARY; o.nh_off = h;
        } else {
            o = l2->routes[kv.second.idx];
            const uint32_t from = o.nh_off;
            o.nh_off = h;
            if (h + o.n_nh > out->nexthops_cap) return HSPF_E_NOMEM;
            for (uint32_t k = 0; k < o.n_nh; ++k) out->nexthops[h++] = l2->nexthops[from + k];
        }
        if (r >= out->routes_cap) return HSPF_E_NOMEM;
        out->routes[r++] = o;
    }
    out->n_routes = r; out->n_nexthops = h;
    return 0;
}

extern "C" int oracle_isis_l1_to_l2(const hl_isis_level *l1, const uint8_t *up_down, uint64_t local_system_id,
                                    const hl_isis_spt *spt_std, const hl_isis_spt *spt_v6, uint8_t l1_metric_type,
                                    uint8_t l2_metric_type, const hl_isis_summary *cfg, uint32_t n_cfg,
                                    const hl_isis_summary *active, uint32_t n_active, hl_isis_ipreach *out, uint32_t cap,
                                    uint32_t *n_out) {
    auto is_std = [](uint8_t t) { return t != HL_ISIS_METRIC_WIDE; };
    auto is_wide = [](uint8_t t) { return t != HL_ISIS_METRIC_STANDARD; };
    // one map per L2 TLV (l2_ipv4_internal_reach, l2_ipv4_external_reach, l2_ext_ipv4_reach, l2_ipv6_reach)
    std::map<Key, hl_isis_ipreach> tlv[4];
    auto propagate = [&](uint32_t dist, const hl_isis_ipreach &src, bool up, int which, bool narrow) {
        if (up) return;                                                    // up/down bit set
        if (get_spm(cfg, n_cfg, src.prefix, src.len)) return;              // covered by a configured summary
        hl_isis_ipreach reach = src;
        if (narrow) reach.metric = std::min<uint32_t>(reach.metric + dist, 63);               // LegacyIpv4Reach::metric_add
        else reach.metric = (uint64_t)reach.metric + dist > 0xFFFFFFFFull ? 0xFFFFFFFFu : reach.metric + dist;
        if (reach.has_psid) {
            reach.psid_flags |= HL_ISIS_PSID_R;
            reach.psid_flags |= HL_ISIS_PSID_P;
            reach.psid_flags &= (uint8_t)~HL_ISIS_PSID_E;
        }
        Key k{reach.prefix, reach.len};
        auto it = tlv[which].find(k);
        if (it == tlv[which].end()) tlv[which].emplace(k, reach);
        else if (reach.metric < it->second.metric) it->second = reach;
    };
    // spt.get(&VertexId::from(system_id)).map(|vertex| vertex.distance)
    auto vertex_distance = [](const hl_isis_spt *spt, uint64_t id, uint32_t *d) {
        for (uint32_t i = 0; spt && i < spt->n_vertices; ++i)
            if (spt->vertices[i].lan_id == id) { *d = spt->vertices[i].distance; return true; }
        return false;
    };
    for (uint32_t i = 0; i < l1->n_lsps; ++i) {
        const hl_isis_lsp &lsp = l1->lsps[i];
        if (lsp.seqno == 0 || lsp.rem_lifetime == 0 || (lsp.lan_id & 0xFF) || (lsp.lan_id >> 8) == local_system_id) continue;
        uint32_t d = 0;
        if (vertex_distance(spt_std, lsp.lan_id, &d)) {
            for (uint32_t k = lsp.ipreach_off; k < lsp.ipreach_off + lsp.n_ipreach; ++k) {
                const hl_isis_ipreach &e = l1->ipreaches[k];
                const bool up = up_down && up_down[k];
                if (l1->ipv4_enabled) {
                    if (is_std(l1_metric_type) && is_std(l2_metric_type)) {
                        if (e.kind == HL_ISIS_IP_V4_INTERNAL) propagate(d, e, up, 0, true);
                        if (e.kind == HL_ISIS_IP_V4_EXTERNAL) propagate(d, e, up, 1, true);
                    }
                    if (is_wide(l1_metric_type) && is_wide(l2_metric_type) && e.kind == HL_ISIS_IP_V4_EXT) propagate(d, e, up, 2, false);
                }
                if (!spt_v6 && l1->ipv6_enabled && e.kind == HL_ISIS_IP_V6) propagate(d, e, up, 3, false);
            }
        }
        if (vertex_distance(spt_v6, lsp.lan_id, &d))
            for (uint32_t k = lsp.ipreach_off; k < lsp.ipreach_off + lsp.n_ipreach; ++k) {
                hl_isis_ipreach e = l1->ipreaches[k];
                if (e.kind != HL_ISIS_IP_MT_V6 || e.mt_id != HL_ISIS_MT_IPV6) continue;
                e.kind = HL_ISIS_IP_V6; e.mt_id = 0;
                propagate(d, e, up_down && up_down[k], 3, false);
            }
    }
    for (uint32_t j = 0; j < n_active; ++j) {          // "Add active summary routes"
        const hl_isis_summary &s = active[j];
        const uint32_t metric = s.has_cfg_metric ? s.cfg_metric : s.metric;
        hl_isis_ipreach e;
        std::memset(&e, 0, sizeof(e));
        e.prefix = s.prefix; e.len = s.len;
        Key k{s.prefix, s.len};
        if (!s.prefix.is_v6) {
            if (!l1->ipv4_enabled) continue;
            if (is_std(l2_metric_type)) { e.kind = HL_ISIS_IP_V4_INTERNAL; e.metric = std::min<uint32_t>(metric, 63); tlv[0][k] = e; }
            if (is_wide(l2_metric_type)) { e.kind = HL_ISIS_IP_V4_EXT; e.metric = metric; tlv[2][k] = e; }
        } else {
            if (!l1->ipv6_enabled) continue;
            e.kind = HL_ISIS_IP_V6; e.metric = metric; tlv[3][k] = e;
        }
    }
    uint32_t n = 0;
    for (int w = 0; w < 4; ++w)
        for (auto &kv : tlv[w]) { if (n < cap) out[n] = kv.second; ++n; }
    *n_out = n;
    return n > cap ? HSPF_E_NOMEM : 0;
}
