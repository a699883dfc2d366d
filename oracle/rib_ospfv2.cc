// TEST INFRASTRUCTURE — CPU restatement of the OSPFv2 routing-table stages that follow
// the per-area SPFs in the reference: update_rib_full (holo-ospf/src/route.rs:146-193) with
// update_rib_inter_area_networks (:449-533), update_rib_inter_area_routers (:653-714),
// update_rib_transit_area (:535-650), update_rib_external (:717-827), route_update (:895-942)
// and route_compare (:944-971).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
// may use anything under oracle/.
//
// Pinned by: the 63 golden OSPFv2 snapshots of the reference's conformance topologies
// (tests/golden/ospfv2.json): with the per-area results of oracle_ospfv2_run_area as input,
// every route of `local-rib` — intra-area and the 269 inter-area ones, including the next
// hops that virtual-link end points obtain from their transit area — is reproduced
// (tests/test_oracle_golden.py).  The goldens contain no type-4 and no type-5 LSAs: the
// inter-area-router and AS-external stages are restated but PARITY UNPINNED.
//
// Deviation kept on purpose (documented in include/holo_spf_lsdb.h): the per-area intra-area
// routes arrive already merged per area, so the shared-table walk of update_rib_intra_area
// across areas (route.rs:156-160) is applied per route, not per stub link.
#include <cstdint>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "../include/holo_lsdb.h"
#include "../include/holo_spf.h"

namespace {

// NexthopKey (route.rs:92-98): (iface_idx, addr) with None < Some; iface order = the
// explicit sort key (SURVEY.md §8a semantics #6).
using NhKey = std::tuple<uint32_t, uint8_t, uint32_t>;
using Nexthops = std::map<NhKey, hl_nexthop>;

struct RouteNet {   // route.rs:32-46, the fields these stages touch
    uint8_t path_type = HL_PATH_INTRA_AREA;
    bool has_area = false;
    uint32_t area_id = 0;
    uint32_t metric = 0;
    bool has_type2 = false;
    uint32_t type2_metric = 0;
    uint32_t tag = 0;
    uint8_t flags = 0;
    Nexthops nexthops;
};

struct RouteRtr {   // route.rs:57-66
    uint32_t area_id = 0;
    uint8_t path_type = HL_PATH_INTRA_AREA;
    uint8_t flags = 0;
    uint32_t metric = 0;
    Nexthops nexthops;
};

// Ipv4Network order: address, then prefix length (ipnetwork derive(Ord)).  Summary / external
// prefixes are with_netmask(lsa_id, mask) WITHOUT apply_mask (ospfv2/spf.rs:552,602), unlike the
// intra-area stubs (:479,512).
using Prefix = std::pair<uint32_t, uint32_t>;   // (addr, mask) — mask order == length order
using Rib = std::map<Prefix, RouteNet>;

int route_compare(const RouteNet &a, const RouteNet &b) {   // route.rs:944-971
    if (a.path_type != b.path_type) return a.path_type < b.path_type ? -1 : 1;
    auto cmp = [](uint32_t x, uint32_t y) { return x < y ? -1 : (x > y ? 1 : 0); };
    if (a.path_type == HL_PATH_TYPE2_EXTERNAL) {
        // Option<u32> order: None < Some
        if (a.has_type2 != b.has_type2) return a.has_type2 ? 1 : -1;
        if (int c = cmp(a.type2_metric, b.type2_metric)) return c;
    }
    return cmp(a.metric, b.metric);
}

void truncate(RouteNet &r, uint32_t max_paths) {   // route.rs:934-941
    while (r.nexthops.size() > max_paths) r.nexthops.erase(std::prev(r.nexthops.end()));
}

void route_update(Rib &rib, Prefix p, RouteNet route, uint32_t max_paths) {   // route.rs:895-942
    auto it = rib.find(p);
    if (it == rib.end()) {
        it = rib.emplace(p, std::move(route)).first;
    } else {
        const int c = route_compare(route, it->second);
        if (c < 0) it->second = std::move(route);
        else if (c == 0)
            for (auto &kv : route.nexthops) it->second.nexthops[kv.first] = kv.second;   // BTreeMap::extend
    }
    truncate(it->second, max_paths);
}

Nexthops lift(const hl_ospfv2_rib_area &a, uint32_t off, uint32_t n) {
    Nexthops m;
    for (uint32_t i = 0; i < n; ++i) {
        hl_nexthop nh = a.spf->nexthops[off + i];
        const uint32_t key = nh.iface < a.n_ifaces ? a.ifaces[nh.iface].sort_key : 0xFFFFFFFFu;
        nh.iface = key;
        m[NhKey(key, nh.has_addr, nh.has_addr ? nh.addr : 0)] = nh;
    }
    return m;
}

}  // namespace

extern "C" int oracle_ospfv2_update_rib_full(uint32_t router_id, uint32_t max_paths, const hl_ospfv2_rib_area *areas,
                                             uint32_t n_areas, const hl_ospfv2_external_lsa *ext, uint32_t n_ext,
                                             hl_ospfv2_rib *out) {
    if ((!areas && n_areas) || !out) return HSPF_E_INVAL;
    Rib rib;
    std::vector<std::map<uint32_t, RouteRtr>> routers(n_areas);

    // ---- intra-area routes of every area into one table (route.rs:156-160) and the
    //      per-area router tables (area.state.routers, spf.rs:627-637)
    for (uint32_t ai = 0; ai < n_areas; ++ai) {
        const hl_ospfv2_rib_area &a = areas[ai];
        if (!a.spf) return HSPF_E_INVAL;
        for (uint32_t i = 0; i < a.spf->n_routers; ++i) {
            const hl_route_rtr &r = a.spf->routers[i];
            RouteRtr e;
            e.area_id = a.area_id; e.path_type = HL_PATH_INTRA_AREA; e.flags = r.flags; e.metric = r.metric;
            e.nexthops = lift(a, r.nh_off, r.n_nh);
            routers[ai][r.router_id] = std::move(e);
        }
        for (uint32_t i = 0; i < a.spf->n_routes; ++i) {
            const hl_route_net &r = a.spf->routes[i];
            const Prefix p(r.prefix, r.mask);
            auto it = rib.find(p);
            if (it != rib.end() && r.metric > it->second.metric) continue;   // route.rs:372-376
            RouteNet n;
            n.path_type = HL_PATH_INTRA_AREA; n.has_area = true; n.area_id = a.area_id; n.metric = r.metric;
            n.flags = r.flags; n.nexthops = lift(a, r.nh_off, r.n_nh);
            route_update(rib, p, std::move(n), max_paths);
        }
    }

    // ---- inter-area routes (route.rs:163-179)
    uint32_t active_areas = 0;
    for (uint32_t ai = 0; ai < n_areas; ++ai) active_areas += areas[ai].active ? 1 : 0;
    for (uint32_t ai = 0; ai < n_areas; ++ai) {
        const hl_ospfv2_rib_area &a = areas[ai];
        // several active areas: only backbone summary-LSAs are examined
        if (active_areas > 1 && a.area_id != 0) continue;
        for (int pass = 0; pass < 2; ++pass) {          // networks (type 3), then routers (type 4)
            for (uint32_t i = 0; i < a.n_summaries; ++i) {
                const hl_ospfv2_summary_lsa &l = a.summaries[i];
                if (l.lsa_type != (pass == 0 ? 3 : 4) || l.maxage) continue;
                if (!(l.metric < HL_LSA_INFINITY) || l.adv_rtr == router_id) continue;
                auto br = routers[ai].find(l.adv_rtr);
                if (br == routers[ai].end() || !(br->second.flags & HL_RTR_FLAG_B)) continue;   // no ABR entry
                const uint32_t metric = br->second.metric + l.metric;
                if (pass == 0) {
                    RouteNet n;
                    n.path_type = HL_PATH_INTER_AREA; n.has_area = true; n.area_id = a.area_id; n.metric = metric;
                    n.nexthops = br->second.nexthops;
                    route_update(rib, Prefix(l.lsa_id, l.mask), std::move(n), max_paths);
                } else {
                    RouteRtr e;     // routers.insert(): replaces whatever was there (route.rs:713)
                    e.area_id = a.area_id; e.path_type = HL_PATH_INTER_AREA; e.flags = HL_RTR_FLAG_E; e.metric = metric;
                    e.nexthops = br->second.nexthops;
                    routers[ai][l.lsa_id] = std::move(e);
                }
            }
        }
    }

    // ---- transit areas: shorter paths through them, and the next hops of virtual links
    //      (route.rs:181-187, 535-650)
    for (uint32_t ai = 0; ai < n_areas; ++ai) {
        const hl_ospfv2_rib_area &a = areas[ai];
        if (!a.spf->transit_capability) continue;
        for (uint32_t i = 0; i < a.n_summaries; ++i) {
            const hl_ospfv2_summary_lsa &l = a.summaries[i];
            if (l.lsa_type != 3 || l.maxage) continue;
            if (!(l.metric < HL_LSA_INFINITY) || l.adv_rtr == router_id) continue;
            auto cur = rib.find(Prefix(l.lsa_id, l.mask));
            if (cur == rib.end()) continue;
            RouteNet &c = cur->second;
            if (!(c.path_type == HL_PATH_INTRA_AREA || c.path_type == HL_PATH_INTER_AREA) ||
                !(c.has_area && c.area_id == 0))
                continue;
            auto br = routers[ai].find(l.adv_rtr);
            if (br == routers[ai].end() || !(br->second.flags & HL_RTR_FLAG_B)) continue;
            const uint32_t metric = br->second.metric + l.metric;
            if (metric < c.metric) {
                RouteNet n;
                n.path_type = HL_PATH_INTER_AREA; n.has_area = true; n.area_id = a.area_id; n.metric = metric;
                n.nexthops = br->second.nexthops;
                c = std::move(n);
            } else if (metric == c.metric) {
                for (auto &kv : br->second.nexthops) c.nexthops[kv.first] = kv.second;   // extend
            }
            truncate(c, max_paths);
        }
    }

    // ---- AS-external routes (route.rs:189-190, 717-827)
    // areas.iter(): area-id order (collections.rs:287-293)
    std::vector<uint32_t> by_id(n_areas);
    for (uint32_t i = 0; i < n_areas; ++i) by_id[i] = i;
    for (uint32_t i = 0; i < n_areas; ++i)
        for (uint32_t j = i + 1; j < n_areas; ++j)
            if (areas[by_id[j]].area_id < areas[by_id[i]].area_id) std::swap(by_id[i], by_id[j]);
    for (uint32_t i = 0; i < n_ext; ++i) {
        const hl_ospfv2_external_lsa &l = ext[i];
        if (l.maxage || !(l.metric < HL_LSA_INFINITY) || l.adv_rtr == router_id) continue;
        std::vector<const RouteRtr *> asbr;
        for (uint32_t ai : by_id) {
            auto it = routers[ai].find(l.adv_rtr);
            if (it != routers[ai].end() && (it->second.flags & HL_RTR_FLAG_E)) asbr.push_back(&it->second);
        }
        std::vector<const RouteRtr *> pruned;   // intra-area paths through non-backbone areas first
        for (auto *r : asbr)
            if (r->path_type == HL_PATH_INTRA_AREA && r->area_id != 0) pruned.push_back(r);
        if (!pruned.empty()) asbr = pruned;
        const RouteRtr *best = nullptr;         // least cost; ties: largest area id
        for (auto *r : asbr) {
            if (!best) { best = r; continue; }
            if (r->metric < best->metric || (r->metric == best->metric && r->area_id > best->area_id)) best = r;
        }
        if (!best) continue;
        RouteNet n;
        n.has_area = false; n.tag = l.tag; n.nexthops = best->nexthops;
        if (l.e_bit) { n.path_type = HL_PATH_TYPE2_EXTERNAL; n.metric = best->metric; n.has_type2 = true; n.type2_metric = l.metric; }
        else { n.path_type = HL_PATH_TYPE1_EXTERNAL; n.metric = best->metric + l.metric; }
        route_update(rib, Prefix(l.lsa_id, l.mask), std::move(n), max_paths);
    }

    // ---- emit
    uint32_t need_h = 0;
    for (auto &kv : rib) need_h += (uint32_t)kv.second.nexthops.size();
    out->n_routes = (uint32_t)rib.size();
    out->n_nexthops = need_h;
    if (out->n_routes > out->routes_cap || need_h > out->nexthops_cap) return HSPF_E_NOMEM;
    uint32_t ri = 0, h = 0;
    for (auto &kv : rib) {
        hl_rib_route o;
        std::memset(&o, 0, sizeof(o));
        o.prefix = kv.first.first; o.mask = kv.first.second; o.metric = kv.second.metric;
        o.type2_metric = kv.second.type2_metric; o.has_type2 = kv.second.has_type2; o.tag = kv.second.tag;
        o.area_id = kv.second.area_id; o.has_area = kv.second.has_area; o.path_type = kv.second.path_type;
        o.flags = kv.second.flags; o.nh_off = h; o.n_nh = (uint32_t)kv.second.nexthops.size();
        for (auto &nk : kv.second.nexthops) out->nexthops[h++] = nk.second;
        out->routes[ri++] = o;
    }
    return HSPF_OK;
}
