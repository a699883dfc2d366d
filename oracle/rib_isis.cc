// TEST INFRASTRUCTURE — CPU restatement of the end of holo-isis' update_rib: the merge of the
// per-level tables (holo-isis/src/route.rs:236-242: rib_l2.iter().chain(rib_l1.iter()).collect(),
// i.e. an L1 route replaces the L2 route of the same prefix) and update_global_rib
// (route.rs:255-314).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use
// anything under oracle/.
//
// Pinned by: the RouteIpAdd streams of the reference's 38 IS-IS conformance snapshots
// (output/ibus.jsonl, final state per prefix): computing both levels, merging and diffing against
// an empty table gives exactly those routes with metric, ifindex and next-hop address
// (tests/test_isis_rib.py).  Summary (blackhole) routes and tags are not modelled.
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
        if (!(route.flags & HL_ROUTE_CONNECTED) && route.n_nh != 0) {          // (SUMMARY routes: not modelled)
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
