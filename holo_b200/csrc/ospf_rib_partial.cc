// Partial SPF runs of OSPFv2 (include/holo_spf_lsdb.h): what holo-ospf does when only summary or
// AS-external LSAs changed (SpfComputation::Partial, holo-ospf/src/spf.rs:123-140) —
//   update_rib_partial               holo-ospf/src/route.rs:196-340
// over the same stages as the full run, each restricted to the named destinations:
//   update_rib_inter_area_networks   route.rs:449-533
//   update_rib_inter_area_routers    route.rs:653-714
//   update_rib_transit_area          route.rs:535-650
//   update_rib_external              route.rs:717-827
//   update_global_rib                route.rs:833-893,   route_update / route_compare  route.rs:895-971
// The state between runs is explicit here: the previous table (hl_ospfv2_rib, INSTALLED flags included) and
// the per-area router tables (hl_ospfv2_rtr_tables).  Host only; ordered maps, because a partial run touches
// a handful of destinations.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <new>
#include <set>
#include <tuple>
#include <vector>

#include "../../include/holo_lsdb.h"
#include "../../include/holo_spf.h"
#include "../../include/holo_spf_lsdb.h"

namespace {

using Prefix = std::pair<uint32_t, uint8_t>;                         // address, length: Ipv4Network order
using HopKey = std::tuple<uint32_t, uint8_t, uint32_t>;              // interface sort key, has address, address
using Hops = std::map<HopKey, hl_nexthop>;

struct Net {
    uint32_t metric = 0, type2 = 0, tag = 0, area = 0;
    uint8_t path = HL_PATH_INTRA_AREA, flags = 0;
    bool has_area = false, has_type2 = false, has_label = false;
    uint32_t label = 0;
    Hops hops;
};
struct Rtr {
    uint32_t area = 0, metric = 0;
    uint8_t path = HL_PATH_INTRA_AREA, flags = 0;
    Hops hops;
};
using Rib = std::map<Prefix, Net>;
using RtrTable = std::map<uint32_t, Rtr>;

inline uint8_t len_of(uint32_t mask) { return (uint8_t)__builtin_popcount(mask); }
inline uint32_t mask_of(uint8_t len) { return len == 0 ? 0u : 0xFFFFFFFFu << (32 - len); }
inline HopKey key_of(const hl_nexthop &h) { return HopKey{h.iface, h.has_addr ? 1 : 0, h.has_addr ? h.addr : 0u}; }

void truncate(Hops &h, uint32_t max_paths) {
    while (h.size() > max_paths) h.erase(std::prev(h.end()));
}

// route_compare (route.rs:944-971): negative when `a` is preferred
int route_compare(const Net &a, const Net &b) {
    if (a.path != b.path) return a.path < b.path ? -1 : 1;
    if (a.path == HL_PATH_TYPE2_EXTERNAL) {
        if (a.has_type2 != b.has_type2) return a.has_type2 ? 1 : -1;
        if (a.type2 != b.type2) return a.type2 < b.type2 ? -1 : 1;
    }
    if (a.metric != b.metric) return a.metric < b.metric ? -1 : 1;
    return 0;
}

void route_update(Rib &rib, const Prefix &p, Net &&route, uint32_t max_paths) {       // route.rs:895-942
    auto it = rib.find(p);
    if (it == rib.end()) {
        it = rib.emplace(p, std::move(route)).first;
    } else {
        const int c = route_compare(route, it->second);
        if (c < 0) it->second = std::move(route);
        else if (c == 0) for (auto &kv : route.hops) it->second.hops[kv.first] = kv.second;
    }
    truncate(it->second.hops, max_paths);
}

bool same_hops(const Hops &a, const Hops &b) {
    if (a.size() != b.size()) return false;
    auto i = a.begin();
    auto j = b.begin();
    for (; i != a.end(); ++i, ++j) {
        const hl_nexthop &x = i->second, &y = j->second;
        if (i->first != j->first || x.has_nbr != y.has_nbr || (x.has_nbr && x.nbr_router_id != y.nbr_router_id) ||
            x.has_label != y.has_label || (x.has_label && x.sr_label != y.sr_label)) return false;
    }
    return true;
}

struct Ctx {
    uint32_t router_id, max_paths;
    const hl_ospfv2_rib_area *areas;
    uint32_t n_areas;
    std::vector<RtrTable> rtrs;
    uint32_t n_active = 0;

    bool usable(const hl_ospfv2_summary_lsa &l) const { return !l.maxage && l.metric < HL_LSA_INFINITY && l.adv_rtr != router_id; }
    bool examined(uint32_t ai) const { return !(n_active > 1 && areas[ai].area_id != 0); }   // route.rs:164-168
    const Rtr *abr(uint32_t ai, uint32_t adv) const {
        auto it = rtrs[ai].find(adv);
        return (it != rtrs[ai].end() && (it->second.flags & HL_RTR_FLAG_B)) ? &it->second : nullptr;
    }

    void inter_area_networks(Rib &rib, const std::set<Prefix> *filter, uint32_t ai) const {
        const auto &a = areas[ai];
        for (uint32_t i = 0; i < a.n_summaries; ++i) {
            const auto &l = a.summaries[i];
            if (l.lsa_type != 3 || !usable(l)) continue;
            const Prefix p{l.lsa_id, len_of(l.mask)};
            if (filter && !filter->count(p)) continue;
            const Rtr *br = abr(ai, l.adv_rtr);
            if (!br) continue;
            Net n;
            n.path = HL_PATH_INTER_AREA; n.has_area = true; n.area = a.area_id; n.metric = br->metric + l.metric; n.hops = br->hops;
            route_update(rib, p, std::move(n), max_paths);
        }
    }

    void inter_area_routers(const std::set<uint32_t> *filter, uint32_t ai) {
        const auto &a = areas[ai];
        for (uint32_t i = 0; i < a.n_summaries; ++i) {
            const auto &l = a.summaries[i];
            if (l.lsa_type != 4 || !usable(l)) continue;
            if (filter && !filter->count(l.lsa_id)) continue;
            const Rtr *br = abr(ai, l.adv_rtr);
            if (!br) continue;
            Rtr e;
            e.area = a.area_id; e.metric = br->metric + l.metric; e.path = HL_PATH_INTER_AREA; e.flags = HL_RTR_FLAG_E; e.hops = br->hops;
            rtrs[ai][l.lsa_id] = std::move(e);               // BTreeMap::insert: replaces
        }
    }

    void transit_area(Rib &rib, uint32_t ai) const {         // route.rs:535-650
        const auto &a = areas[ai];
        for (uint32_t i = 0; i < a.n_summaries; ++i) {
            const auto &l = a.summaries[i];
            if (l.lsa_type != 3 || !usable(l)) continue;
            auto it = rib.find(Prefix{l.lsa_id, len_of(l.mask)});
            if (it == rib.end()) continue;
            Net &cur = it->second;
            if (cur.path > HL_PATH_INTER_AREA || !cur.has_area || cur.area != 0) continue;
            const Rtr *br = abr(ai, l.adv_rtr);
            if (!br) continue;
            const uint32_t metric = br->metric + l.metric;
            if (metric < cur.metric) {
                const bool installed = cur.flags & HL_ROUTE_INSTALLED;
                Net n;
                n.path = HL_PATH_INTER_AREA; n.has_area = true; n.area = a.area_id; n.metric = metric; n.hops = br->hops;
                if (installed) n.flags |= HL_ROUTE_INSTALLED;
                cur = std::move(n);
            } else if (metric == cur.metric) {
                for (auto &kv : br->hops) cur.hops[kv.first] = kv.second;
            }
            truncate(cur.hops, max_paths);
        }
    }

    void external(Rib &rib, const std::set<Prefix> *filter, const hl_ospfv2_external_lsa *ext, uint32_t n_ext) const {
        std::vector<uint32_t> order(n_areas);                // Areas iterate in area-id order
        for (uint32_t i = 0; i < n_areas; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return areas[x].area_id < areas[y].area_id; });
        for (uint32_t i = 0; i < n_ext; ++i) {
            const auto &l = ext[i];
            if (l.maxage || !(l.metric < HL_LSA_INFINITY) || l.adv_rtr == router_id) continue;
            const Prefix p{l.lsa_id, len_of(l.mask)};
            if (filter && !filter->count(p)) continue;
            std::vector<const Rtr *> cand, pruned;
            for (uint32_t ai : order) {
                auto it = rtrs[ai].find(l.adv_rtr);
                if (it != rtrs[ai].end() && (it->second.flags & HL_RTR_FLAG_E)) cand.push_back(&it->second);
            }
            for (const Rtr *r : cand)
                if (r->path == HL_PATH_INTRA_AREA && r->area != 0) pruned.push_back(r);
            if (!pruned.empty()) cand.swap(pruned);
            const Rtr *best = nullptr;
            for (const Rtr *r : cand)
                if (!best || r->metric < best->metric || (r->metric == best->metric && r->area > best->area)) best = r;
            if (!best) continue;
            Net n;
            n.tag = l.tag; n.hops = best->hops;
            if (l.e_bit) { n.path = HL_PATH_TYPE2_EXTERNAL; n.metric = best->metric; n.has_type2 = true; n.type2 = l.metric; }
            else { n.path = HL_PATH_TYPE1_EXTERNAL; n.metric = best->metric + l.metric; }
            route_update(rib, p, std::move(n), max_paths);
        }
    }
};

Hops hops_from(const hl_nexthop *nh, uint32_t off, uint32_t n) {
    Hops h;
    for (uint32_t i = 0; i < n; ++i) h[key_of(nh[off + i])] = nh[off + i];
    return h;
}

int load_tables(const hl_ospfv2_rtr_tables *t, const hl_ospfv2_rib_area *areas, uint32_t n_areas, std::vector<RtrTable> &out) {
    out.assign(n_areas, {});
    if ((t->n_rtrs && !t->rtrs) || (t->n_nexthops && !t->nexthops)) return HSPF_E_INVAL;
    for (uint32_t i = 0; i < t->n_rtrs; ++i) {
        const hl_rib_rtr &r = t->rtrs[i];
        if ((uint64_t)r.nh_off + r.n_nh > t->n_nexthops) return HSPF_E_INVAL;
        uint32_t ai = 0;
        while (ai < n_areas && areas[ai].area_id != r.area_id) ++ai;
        if (ai == n_areas) continue;                         // an area the instance no longer has
        Rtr e;
        e.area = r.area_id; e.metric = r.metric; e.path = r.path_type; e.flags = r.flags; e.hops = hops_from(t->nexthops, r.nh_off, r.n_nh);
        out[ai][r.router_id] = std::move(e);
    }
    return HSPF_OK;
}

int store_tables(const std::vector<RtrTable> &tabs, hl_ospfv2_rtr_tables *out) {
    uint32_t n = 0, h = 0;
    for (auto &t : tabs) for (auto &kv : t) { ++n; h += (uint32_t)kv.second.hops.size(); }
    out->n_rtrs = n; out->n_nexthops = h;
    if (n > out->rtrs_cap || h > out->nexthops_cap) return HSPF_E_NOMEM;
    if ((n && !out->rtrs) || (h && !out->nexthops)) return HSPF_E_INVAL;
    n = h = 0;
    for (auto &t : tabs)
        for (auto &kv : t) {
            hl_rib_rtr r;
            std::memset(&r, 0, sizeof(r));
            r.area_id = kv.second.area; r.router_id = kv.first; r.metric = kv.second.metric; r.path_type = kv.second.path;
            r.flags = kv.second.flags; r.nh_off = h; r.n_nh = (uint32_t)kv.second.hops.size();
            for (auto &x : kv.second.hops) out->nexthops[h++] = x.second;
            out->rtrs[n++] = r;
        }
    return HSPF_OK;
}

int router_tables(uint32_t router_id, const hl_ospfv2_rib_area *areas, uint32_t n_areas, hl_ospfv2_rtr_tables *out) {
    if ((!areas && n_areas) || !out) return HSPF_E_INVAL;
    Ctx c{router_id, 0, areas, n_areas, std::vector<RtrTable>(n_areas), 0};
    for (uint32_t ai = 0; ai < n_areas; ++ai) {
        const auto &a = areas[ai];
        if (!a.spf || (a.n_ifaces && !a.ifaces) || (a.n_summaries && !a.summaries)) return HSPF_E_INVAL;
        c.n_active += a.active ? 1u : 0u;
        for (uint32_t i = 0; i < a.spf->n_routers; ++i) {    // area.state.routers after run_area (spf.rs:627-637)
            const hl_route_rtr &r = a.spf->routers[i];
            if ((uint64_t)r.nh_off + r.n_nh > a.spf->n_nexthops) return HSPF_E_INVAL;
            Rtr e;
            e.area = a.area_id; e.metric = r.metric; e.path = HL_PATH_INTRA_AREA; e.flags = r.flags;
            for (uint32_t k = 0; k < r.n_nh; ++k) {
                hl_nexthop x = a.spf->nexthops[r.nh_off + k];
                x.iface = x.iface < a.n_ifaces ? a.ifaces[x.iface].sort_key : 0xFFFFFFFFu;   // instance-wide names
                e.hops[key_of(x)] = x;
            }
            c.rtrs[ai][r.router_id] = std::move(e);
        }
    }
    for (uint32_t ai = 0; ai < n_areas; ++ai)
        if (c.examined(ai)) c.inter_area_routers(nullptr, ai);
    return store_tables(c.rtrs, out);
}

int partial_run(uint32_t router_id, uint32_t max_paths, const hl_ospfv2_rib_area *areas, const uint8_t *transit, uint32_t n_areas,
                const hl_ospfv2_external_lsa *ext, uint32_t n_ext, const hl_spf_computation *pc, const hl_ospfv2_rib *prev,
                const hl_ospfv2_rtr_tables *prev_rtrs, hl_ospfv2_rib *out, hl_ospfv2_rtr_tables *out_rtrs, hl_rib_action *actions,
                uint32_t cap, uint32_t *n_actions) {
    if ((!areas && n_areas) || (n_areas && !transit) || (!ext && n_ext) || !pc || !prev || !prev_rtrs || !out || !out_rtrs || !n_actions ||
        (cap && !actions)) return HSPF_E_INVAL;
    if (pc->kind != HL_SPF_PARTIAL) return HSPF_E_INVAL;
    if ((pc->n_inter_network && !pc->inter_network) || (pc->n_inter_router && !pc->inter_router) || (pc->n_external && !pc->external))
        return HSPF_E_INVAL;
    if ((prev->n_routes && !prev->routes) || (prev->n_nexthops && !prev->nexthops)) return HSPF_E_INVAL;
    Ctx c{router_id, max_paths, areas, n_areas, {}, 0};
    for (uint32_t ai = 0; ai < n_areas; ++ai) {
        if (areas[ai].n_summaries && !areas[ai].summaries) return HSPF_E_INVAL;
        c.n_active += areas[ai].active ? 1u : 0u;
    }
    int rc = load_tables(prev_rtrs, areas, n_areas, c.rtrs);
    if (rc) return rc;

    // the table of the previous run; prev_index: where each of its routes sits (for UNINSTALL_OLD)
    Rib rib, partial_rib, old_rib;
    std::map<Prefix, uint32_t> prev_index;
    for (uint32_t i = 0; i < prev->n_routes; ++i) {
        const hl_rib_route &r = prev->routes[i];
        if ((uint64_t)r.nh_off + r.n_nh > prev->n_nexthops) return HSPF_E_INVAL;
        Net n;
        n.metric = r.metric; n.type2 = r.type2_metric; n.tag = r.tag; n.area = r.area_id; n.path = r.path_type; n.flags = r.flags;
        n.has_area = r.has_area; n.has_type2 = r.has_type2; n.has_label = r.has_sr_label; n.label = r.sr_label;
        n.hops = hops_from(prev->nexthops, r.nh_off, r.n_nh);
        const Prefix p{r.prefix, len_of(r.mask)};
        rib[p] = std::move(n);
        prev_index[p] = i;
    }
    std::set<Prefix> inter_network, external;
    std::set<uint32_t> inter_router(pc->inter_router, pc->inter_router + pc->n_inter_router);
    for (uint32_t i = 0; i < pc->n_inter_network; ++i) inter_network.insert({pc->inter_network[i].addr, len_of(pc->inter_network[i].mask)});
    for (uint32_t i = 0; i < pc->n_external; ++i) external.insert({pc->external[i].addr, len_of(pc->external[i].mask)});
    auto take_out = [&](auto pred) {
        for (auto it = rib.begin(); it != rib.end();) {
            if (pred(it->first, it->second)) { old_rib[it->first] = std::move(it->second); it = rib.erase(it); }
            else ++it;
        }
    };

    // inter-area networks (route.rs:252-287)
    if (!inter_network.empty()) {
        take_out([&](const Prefix &p, const Net &n) { return inter_network.count(p) && n.path == HL_PATH_INTER_AREA; });
        for (uint32_t ai = 0; ai < n_areas; ++ai)
            if (c.examined(ai)) c.inter_area_networks(partial_rib, &inter_network, ai);
        for (auto &kv : old_rib) external.insert(kv.first);          // newly unreachable: look for external paths
    }
    // inter-area routers (route.rs:288-312)
    if (!inter_router.empty()) {
        for (uint32_t ai = 0; ai < n_areas; ++ai) {
            if (!c.examined(ai)) continue;
            for (auto it = c.rtrs[ai].begin(); it != c.rtrs[ai].end();) {
                if (inter_router.count(it->first) && it->second.path == HL_PATH_INTER_AREA) it = c.rtrs[ai].erase(it);
                else ++it;
            }
            c.inter_area_routers(&inter_router, ai);
        }
    }
    // transit areas, on the routes that stayed (route.rs:314-321)
    for (uint32_t ai = 0; ai < n_areas; ++ai)
        if (transit[ai]) c.transit_area(rib, ai);
    // external routes (route.rs:323-354)
    if (!inter_router.empty() || !external.empty()) {
        const bool all = !inter_router.empty();
        take_out([&](const Prefix &p, const Net &n) {
            return (all || external.count(p)) && (n.path == HL_PATH_TYPE1_EXTERNAL || n.path == HL_PATH_TYPE2_EXTERNAL);
        });
        c.external(partial_rib, all ? nullptr : &external, ext, n_ext);
    }

    // update_global_rib(partial_rib, old_rib) (route.rs:833-893)
    struct Act { uint8_t kind; Prefix p; bool has_old; uint32_t old_label; };
    std::vector<Act> acts;
    auto metric_of = [](const Net &n) { return n.path == HL_PATH_TYPE2_EXTERNAL ? n.type2 : n.metric; };
    for (auto &kv : partial_rib) {
        Net &r = kv.second;
        bool has_old = false; uint32_t old_label = 0;
        auto o = old_rib.find(kv.first);
        if (o != old_rib.end()) {
            const Net old = std::move(o->second);
            old_rib.erase(o);
            has_old = old.has_label; old_label = old.label;
            if (metric_of(old) == metric_of(r) && old.tag == r.tag && old.has_label == r.has_label &&
                (!old.has_label || old.label == r.label) && same_hops(old.hops, r.hops)) {
                if (old.flags & HL_ROUTE_INSTALLED) r.flags |= HL_ROUTE_INSTALLED;
                continue;
            }
        }
        if (!(r.flags & HL_ROUTE_CONNECTED) && !r.hops.empty()) {
            acts.push_back({HL_RIB_INSTALL, kv.first, has_old, old_label});
            r.flags |= HL_ROUTE_INSTALLED;
        } else if (r.flags & HL_ROUTE_INSTALLED) {
            acts.push_back({HL_RIB_UNINSTALL, kv.first, false, 0});
            r.flags &= (uint8_t)~HL_ROUTE_INSTALLED;
        }
    }
    for (auto &kv : old_rib)
        if (kv.second.flags & HL_ROUTE_INSTALLED) acts.push_back({HL_RIB_UNINSTALL_OLD, kv.first, false, 0});
    for (auto &kv : partial_rib) rib[kv.first] = std::move(kv.second);       // rib.extend(partial_rib)

    // ---- outputs --------------------------------------------------------------------------------
    uint32_t n_h = 0;
    for (auto &kv : rib) n_h += (uint32_t)kv.second.hops.size();
    out->n_routes = (uint32_t)rib.size();
    out->n_nexthops = n_h;
    *n_actions = (uint32_t)acts.size();
    rc = store_tables(c.rtrs, out_rtrs);
    if (out->n_routes > out->routes_cap || n_h > out->nexthops_cap || acts.size() > cap) return HSPF_E_NOMEM;
    if (rc) return rc;
    if ((out->n_routes && !out->routes) || (n_h && !out->nexthops)) return HSPF_E_INVAL;
    std::map<Prefix, uint32_t> new_index;
    uint32_t i = 0, h = 0;
    for (auto &kv : rib) {
        const Net &n = kv.second;
        hl_rib_route o;
        std::memset(&o, 0, sizeof(o));
        o.prefix = kv.first.first; o.mask = mask_of(kv.first.second);
        o.metric = n.metric; o.type2_metric = n.type2; o.tag = n.tag; o.area_id = n.area; o.path_type = n.path; o.flags = n.flags;
        o.has_area = n.has_area ? 1 : 0; o.has_type2 = n.has_type2 ? 1 : 0;
        o.has_sr_label = n.has_label ? 1 : 0; o.sr_label = n.has_label ? n.label : 0;
        o.nh_off = h; o.n_nh = (uint32_t)n.hops.size();
        for (auto &x : n.hops) out->nexthops[h++] = x.second;
        new_index[kv.first] = i;
        out->routes[i++] = o;
    }
    for (size_t k = 0; k < acts.size(); ++k) {
        hl_rib_action a;
        std::memset(&a, 0, sizeof(a));
        a.kind = acts[k].kind;
        a.route = acts[k].kind == HL_RIB_UNINSTALL_OLD ? prev_index.at(acts[k].p) : new_index.at(acts[k].p);
        a.has_old_sr_label = acts[k].has_old ? 1 : 0;
        a.old_sr_label = acts[k].has_old ? acts[k].old_label : 0;
        actions[k] = a;
    }
    return HSPF_OK;
}

}  // namespace

extern "C" int hspf_ospfv2_rib_router_tables(uint32_t router_id, const hl_ospfv2_rib_area *areas, uint32_t n_areas,
                                             hl_ospfv2_rtr_tables *out) {
    try { return router_tables(router_id, areas, n_areas, out); }
    catch (const std::bad_alloc &) { return HSPF_E_NOMEM; } catch (...) { return HSPF_E_INVAL; }
}

extern "C" int hspf_ospfv2_update_rib_partial(uint32_t router_id, uint32_t max_paths, const hl_ospfv2_rib_area *areas,
                                              const uint8_t *transit_capability, uint32_t n_areas,
                                              const hl_ospfv2_external_lsa *ext, uint32_t n_ext, const hl_spf_computation *partial,
                                              const hl_ospfv2_rib *prev_rib, const hl_ospfv2_rtr_tables *prev_rtrs,
                                              hl_ospfv2_rib *out_rib, hl_ospfv2_rtr_tables *out_rtrs, hl_rib_action *actions,
                                              uint32_t actions_cap, uint32_t *n_actions) {
    try {
        return partial_run(router_id, max_paths, areas, transit_capability, n_areas, ext, n_ext, partial, prev_rib, prev_rtrs, out_rib,
                           out_rtrs, actions, actions_cap, n_actions);
    } catch (const std::bad_alloc &) { return HSPF_E_NOMEM; } catch (...) { return HSPF_E_INVAL; }
}
