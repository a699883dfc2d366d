// TEST INFRASTRUCTURE — CPU restatement of holo-ospf's partial SPF run for OSPFv2:
//   update_rib_partial (holo-ospf/src/route.rs:196-340) over update_rib_inter_area_networks (:449-533),
//   update_rib_inter_area_routers (:653-714), update_rib_transit_area (:535-650), update_rib_external (:717-827),
//   update_global_rib (:833-893), route_update / route_compare (:895-971),
// and of area.state.routers as a full run leaves it (spf.rs:627-637 + route.rs:653-714).
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use anything under oracle/.
//
// PARITY UNPINNED: no reference fixture records a partial run separately (its step tests show the resulting
// ibus messages, which tests/test_ospf_rib.py reproduces through the full stages).  The restatement is
// deliberately plain — vectors in table order and linear searches, each block in the order of the reference
// text — and shares no code with the product (holo_b200/csrc/ospf_rib_partial.cc).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/holo_lsdb.h"
#include "../include/holo_spf.h"

namespace {

struct P { uint32_t addr; uint8_t len; };
inline bool operator==(const P &a, const P &b) { return a.addr == b.addr && a.len == b.len; }
inline bool operator<(const P &a, const P &b) { return a.addr != b.addr ? a.addr < b.addr : a.len < b.len; }
inline uint8_t plen(uint32_t mask) { uint8_t n = 0; while (mask) { n += mask & 1; mask >>= 1; } return n; }

struct NhList {                      // BTreeMap<NexthopKey, Nexthop>: sorted vector, unique keys
    std::vector<hl_nexthop> v;
    static bool less(const hl_nexthop &a, const hl_nexthop &b) {
        if (a.iface != b.iface) return a.iface < b.iface;
        if ((a.has_addr != 0) != (b.has_addr != 0)) return !a.has_addr;          // None < Some
        return a.has_addr && a.addr < b.addr;
    }
    void insert(const hl_nexthop &x) {
        size_t i = 0;
        while (i < v.size() && less(v[i], x)) ++i;
        if (i < v.size() && !less(x, v[i])) v[i] = x; else v.insert(v.begin() + i, x);
    }
    void extend(const NhList &o) { for (auto &x : o.v) insert(x); }
    void take(uint32_t n) { if (v.size() > n) v.resize(n); }
    bool same(const NhList &o) const {
        if (v.size() != o.v.size()) return false;
        for (size_t i = 0; i < v.size(); ++i) {
            const auto &a = v[i], &b = o.v[i];
            if (a.iface != b.iface || (a.has_addr != 0) != (b.has_addr != 0) || (a.has_addr && a.addr != b.addr)) return false;
            if ((a.has_nbr != 0) != (b.has_nbr != 0) || (a.has_nbr && a.nbr_router_id != b.nbr_router_id)) return false;
            if ((a.has_label != 0) != (b.has_label != 0) || (a.has_label && a.sr_label != b.sr_label)) return false;
        }
        return true;
    }
};

struct Route {
    P prefix;
    uint8_t path_type, flags;
    bool has_area, has_type2, has_label;
    uint32_t area_id, metric, type2_metric, tag, label;
    NhList nexthops;
    uint32_t report_metric() const { return path_type == HL_PATH_TYPE2_EXTERNAL ? type2_metric : metric; }
};
struct Router { uint32_t area_id, router_id, metric; uint8_t path_type, flags; NhList nexthops; };

struct Table {                       // BTreeMap<prefix, route>: vector kept in prefix order
    std::vector<Route> r;
    Route *get(const P &p) { for (auto &x : r) if (x.prefix == p) return &x; return nullptr; }
    void put(const Route &x) {
        size_t i = 0;
        while (i < r.size() && r[i].prefix < x.prefix) ++i;
        if (i < r.size() && r[i].prefix == x.prefix) r[i] = x; else r.insert(r.begin() + i, x);
    }
};

int compare(const Route &a, const Route &b) {
    if (a.path_type != b.path_type) return a.path_type < b.path_type ? -1 : 1;
    if (a.path_type == HL_PATH_TYPE2_EXTERNAL) {
        if (a.has_type2 != b.has_type2) return a.has_type2 ? 1 : -1;
        if (a.type2_metric != b.type2_metric) return a.type2_metric < b.type2_metric ? -1 : 1;
    }
    if (a.metric != b.metric) return a.metric < b.metric ? -1 : 1;
    return 0;
}
void route_update(Table &t, const Route &nr, uint32_t max_paths) {
    Route *cur = t.get(nr.prefix);
    if (!cur) { t.put(nr); cur = t.get(nr.prefix); }
    else {
        const int c = compare(nr, *cur);
        if (c < 0) *cur = nr; else if (c == 0) cur->nexthops.extend(nr.nexthops);
    }
    cur->nexthops.take(max_paths);
}

struct World {
    uint32_t router_id, max_paths, n_areas;
    const hl_ospfv2_rib_area *areas;
    std::vector<Router> routers;     // all areas' tables; (area_id, router_id) identifies an entry
    uint32_t active_areas() const { uint32_t n = 0; for (uint32_t i = 0; i < n_areas; ++i) n += areas[i].active ? 1 : 0; return n; }
    Router *router(uint32_t area_id, uint32_t rid) {
        for (auto &x : routers) if (x.area_id == area_id && x.router_id == rid) return &x;
        return nullptr;
    }
    void router_insert(const Router &x) {
        if (Router *cur = router(x.area_id, x.router_id)) *cur = x; else routers.push_back(x);
    }
    static bool lsa_ok(const hl_ospfv2_summary_lsa &l, uint32_t self) { return !l.maxage && l.metric < HL_LSA_INFINITY && l.adv_rtr != self; }

    void inter_area_networks(Table &rib, const std::vector<P> *filter, const hl_ospfv2_rib_area &a) {
        for (uint32_t i = 0; i < a.n_summaries; ++i) {
            const auto &l = a.summaries[i];
            if (l.lsa_type != 3 || !lsa_ok(l, router_id)) continue;
            const P p{l.lsa_id, plen(l.mask)};
            if (filter && std::find(filter->begin(), filter->end(), p) == filter->end()) continue;
            Router *br = router(a.area_id, l.adv_rtr);
            if (!br || !(br->flags & HL_RTR_FLAG_B)) continue;
            Route nr{};
            nr.prefix = p; nr.path_type = HL_PATH_INTER_AREA; nr.has_area = true; nr.area_id = a.area_id;
            nr.metric = br->metric + l.metric; nr.nexthops = br->nexthops;
            route_update(rib, nr, max_paths);
        }
    }
    void inter_area_routers(const std::vector<uint32_t> *filter, const hl_ospfv2_rib_area &a) {
        for (uint32_t i = 0; i < a.n_summaries; ++i) {
            const auto &l = a.summaries[i];
            if (l.lsa_type != 4 || !lsa_ok(l, router_id)) continue;
            if (filter && std::find(filter->begin(), filter->end(), l.lsa_id) == filter->end()) continue;
            Router *br = router(a.area_id, l.adv_rtr);
            if (!br || !(br->flags & HL_RTR_FLAG_B)) continue;
            Router nr{a.area_id, l.lsa_id, br->metric + l.metric, HL_PATH_INTER_AREA, HL_RTR_FLAG_E, br->nexthops};
            router_insert(nr);
        }
    }
    void transit_area(Table &rib, const hl_ospfv2_rib_area &a) {
        for (uint32_t i = 0; i < a.n_summaries; ++i) {
            const auto &l = a.summaries[i];
            if (l.lsa_type != 3 || !lsa_ok(l, router_id)) continue;
            Route *cur = rib.get(P{l.lsa_id, plen(l.mask)});
            if (!cur) continue;
            if (!(cur->path_type == HL_PATH_INTRA_AREA || cur->path_type == HL_PATH_INTER_AREA) || !cur->has_area || cur->area_id != 0) continue;
            Router *br = router(a.area_id, l.adv_rtr);
            if (!br || !(br->flags & HL_RTR_FLAG_B)) continue;
            const uint32_t metric = br->metric + l.metric;
            if (metric < cur->metric) {
                const bool installed = cur->flags & HL_ROUTE_INSTALLED;
                Route nr{};
                nr.prefix = cur->prefix; nr.path_type = HL_PATH_INTER_AREA; nr.has_area = true; nr.area_id = a.area_id; nr.metric = metric;
                nr.nexthops = br->nexthops; nr.flags = installed ? HL_ROUTE_INSTALLED : 0;
                *cur = nr;
            } else if (metric == cur->metric) {
                cur->nexthops.extend(br->nexthops);
            }
            cur->nexthops.take(max_paths);
        }
    }
    void external(Table &rib, const std::vector<P> *filter, const hl_ospfv2_external_lsa *ext, uint32_t n_ext) {
        std::vector<uint32_t> by_id(n_areas);
        for (uint32_t i = 0; i < n_areas; ++i) by_id[i] = i;
        std::stable_sort(by_id.begin(), by_id.end(), [&](uint32_t x, uint32_t y) { return areas[x].area_id < areas[y].area_id; });
        for (uint32_t i = 0; i < n_ext; ++i) {
            const auto &l = ext[i];
            if (l.maxage || l.metric >= HL_LSA_INFINITY || l.adv_rtr == router_id) continue;
            const P p{l.lsa_id, plen(l.mask)};
            if (filter && std::find(filter->begin(), filter->end(), p) == filter->end()) continue;
            std::vector<Router *> asbr;
            for (uint32_t ai : by_id) {
                Router *r = router(areas[ai].area_id, l.adv_rtr);
                if (r && (r->flags & HL_RTR_FLAG_E)) asbr.push_back(r);
            }
            std::vector<Router *> pruned;
            for (Router *r : asbr) if (r->path_type == HL_PATH_INTRA_AREA && r->area_id != 0) pruned.push_back(r);
            if (!pruned.empty()) asbr = pruned;
            if (asbr.empty()) continue;
            Router *best = asbr[0];
            for (size_t k = 1; k < asbr.size(); ++k) {
                Router *r = asbr[k];
                if (r->metric < best->metric) best = r;
                else if (r->metric == best->metric && r->area_id > best->area_id) best = r;
            }
            Route nr{};
            nr.prefix = p; nr.tag = l.tag; nr.nexthops = best->nexthops;
            if (l.e_bit) { nr.path_type = HL_PATH_TYPE2_EXTERNAL; nr.metric = best->metric; nr.has_type2 = true; nr.type2_metric = l.metric; }
            else { nr.path_type = HL_PATH_TYPE1_EXTERNAL; nr.metric = best->metric + l.metric; }
            route_update(rib, nr, max_paths);
        }
    }
};

int put_tables(const std::vector<Router> &rs, const hl_ospfv2_rib_area *areas, uint32_t n_areas, hl_ospfv2_rtr_tables *out) {
    // (area order of the call, router id)
    std::vector<const Router *> order;
    for (uint32_t ai = 0; ai < n_areas; ++ai) {
        std::vector<const Router *> of;
        for (auto &r : rs) if (r.area_id == areas[ai].area_id) of.push_back(&r);
        std::sort(of.begin(), of.end(), [](const Router *a, const Router *b) { return a->router_id < b->router_id; });
        order.insert(order.end(), of.begin(), of.end());
    }
    uint32_t h = 0;
    for (auto *r : order) h += (uint32_t)r->nexthops.v.size();
    out->n_rtrs = (uint32_t)order.size(); out->n_nexthops = h;
    if (out->n_rtrs > out->rtrs_cap || h > out->nexthops_cap) return HSPF_E_NOMEM;
    h = 0;
    for (size_t i = 0; i < order.size(); ++i) {
        hl_rib_rtr o;
        std::memset(&o, 0, sizeof(o));
        o.area_id = order[i]->area_id; o.router_id = order[i]->router_id; o.metric = order[i]->metric; o.path_type = order[i]->path_type;
        o.flags = order[i]->flags; o.nh_off = h; o.n_nh = (uint32_t)order[i]->nexthops.v.size();
        for (auto &x : order[i]->nexthops.v) out->nexthops[h++] = x;
        out->rtrs[i] = o;
    }
    return 0;
}

}  // namespace

extern "C" int oracle_ospfv2_rib_router_tables(uint32_t router_id, const hl_ospfv2_rib_area *areas, uint32_t n_areas,
                                               hl_ospfv2_rtr_tables *out) {
    World w{router_id, 0, n_areas, areas, {}};
    for (uint32_t ai = 0; ai < n_areas; ++ai) {
        const auto &a = areas[ai];
        for (uint32_t i = 0; i < a.spf->n_routers; ++i) {
            const hl_route_rtr &r = a.spf->routers[i];
            Router e{a.area_id, r.router_id, r.metric, HL_PATH_INTRA_AREA, r.flags, {}};
            for (uint32_t k = 0; k < r.n_nh; ++k) {
                hl_nexthop x = a.spf->nexthops[r.nh_off + k];
                x.iface = x.iface < a.n_ifaces ? a.ifaces[x.iface].sort_key : 0xFFFFFFFFu;
                e.nexthops.insert(x);
            }
            w.router_insert(e);
        }
    }
    const uint32_t active = w.active_areas();
    for (uint32_t ai = 0; ai < n_areas; ++ai) {
        if (active > 1 && areas[ai].area_id != 0) continue;
        w.inter_area_routers(nullptr, areas[ai]);
    }
    return put_tables(w.routers, areas, n_areas, out);
}

extern "C" int oracle_ospfv2_update_rib_partial(uint32_t router_id, uint32_t max_paths, const hl_ospfv2_rib_area *areas,
                                                const uint8_t *transit_capability, uint32_t n_areas,
                                                const hl_ospfv2_external_lsa *ext, uint32_t n_ext, const hl_spf_computation *pc,
                                                const hl_ospfv2_rib *prev, const hl_ospfv2_rtr_tables *prev_rtrs,
                                                hl_ospfv2_rib *out, hl_ospfv2_rtr_tables *out_rtrs, hl_rib_action *actions,
                                                uint32_t cap, uint32_t *n_actions) {
    World w{router_id, max_paths, n_areas, areas, {}};
    for (uint32_t i = 0; i < prev_rtrs->n_rtrs; ++i) {
        const hl_rib_rtr &r = prev_rtrs->rtrs[i];
        bool known = false;
        for (uint32_t ai = 0; ai < n_areas; ++ai) known = known || areas[ai].area_id == r.area_id;
        if (!known) continue;
        Router e{r.area_id, r.router_id, r.metric, r.path_type, r.flags, {}};
        for (uint32_t k = 0; k < r.n_nh; ++k) e.nexthops.insert(prev_rtrs->nexthops[r.nh_off + k]);
        w.router_insert(e);
    }
    // let mut partial_rib = BTreeMap::new(); let mut rib = take(instance.state.rib); let mut old_rib = BTreeMap::new();
    Table rib, partial_rib, old_rib;
    for (uint32_t i = 0; i < prev->n_routes; ++i) {
        const hl_rib_route &r = prev->routes[i];
        Route x{};
        x.prefix = P{r.prefix, plen(r.mask)}; x.path_type = r.path_type; x.flags = r.flags; x.has_area = r.has_area; x.has_type2 = r.has_type2;
        x.has_label = r.has_sr_label; x.area_id = r.area_id; x.metric = r.metric; x.type2_metric = r.type2_metric; x.tag = r.tag; x.label = r.sr_label;
        for (uint32_t k = 0; k < r.n_nh; ++k) x.nexthops.insert(prev->nexthops[r.nh_off + k]);
        rib.put(x);
    }
    std::vector<P> inter_network, external;
    std::vector<uint32_t> inter_router(pc->inter_router, pc->inter_router + pc->n_inter_router);
    for (uint32_t i = 0; i < pc->n_inter_network; ++i) inter_network.push_back(P{pc->inter_network[i].addr, plen(pc->inter_network[i].mask)});
    for (uint32_t i = 0; i < pc->n_external; ++i) external.push_back(P{pc->external[i].addr, plen(pc->external[i].mask)});
    auto has = [](const std::vector<P> &v, const P &p) { return std::find(v.begin(), v.end(), p) != v.end(); };
    const uint32_t active = w.active_areas();

    // (partial.intra is always empty for OSPFv2: ospfv2/spf.rs:123-125)
    if (!inter_network.empty()) {
        // old_rib.extend(rib.extract_if(prefix in inter_network && InterArea))
        for (size_t i = 0; i < rib.r.size();) {
            if (has(inter_network, rib.r[i].prefix) && rib.r[i].path_type == HL_PATH_INTER_AREA) { old_rib.put(rib.r[i]); rib.r.erase(rib.r.begin() + i); }
            else ++i;
        }
        for (uint32_t ai = 0; ai < n_areas; ++ai) {
            if (active > 1 && areas[ai].area_id != 0) continue;
            w.inter_area_networks(partial_rib, &inter_network, areas[ai]);
        }
        for (auto &x : old_rib.r) if (!has(external, x.prefix)) external.push_back(x.prefix);     // partial.external.extend(old_rib.keys())
    }
    if (!inter_router.empty()) {
        for (uint32_t ai = 0; ai < n_areas; ++ai) {
            if (active > 1 && areas[ai].area_id != 0) continue;
            // area.state.routers.retain(!inter_router.contains(id) || path_type != InterArea)
            for (size_t i = 0; i < w.routers.size();) {
                const Router &r = w.routers[i];
                const bool listed = std::find(inter_router.begin(), inter_router.end(), r.router_id) != inter_router.end();
                if (r.area_id == areas[ai].area_id && listed && r.path_type == HL_PATH_INTER_AREA) w.routers.erase(w.routers.begin() + i);
                else ++i;
            }
            w.inter_area_routers(&inter_router, areas[ai]);
        }
    }
    for (uint32_t ai = 0; ai < n_areas; ++ai)
        if (transit_capability[ai]) w.transit_area(rib, areas[ai]);
    if (!inter_router.empty() || !external.empty()) {
        const bool reevaluate_all = !inter_router.empty();
        for (size_t i = 0; i < rib.r.size();) {
            const Route &x = rib.r[i];
            if ((reevaluate_all || has(external, x.prefix)) && (x.path_type == HL_PATH_TYPE1_EXTERNAL || x.path_type == HL_PATH_TYPE2_EXTERNAL)) {
                old_rib.put(x); rib.r.erase(rib.r.begin() + i);
            } else ++i;
        }
        w.external(partial_rib, reevaluate_all ? nullptr : &external, ext, n_ext);
    }
    // update_global_rib(&mut partial_rib, old_rib, ..)
    struct Act { uint8_t kind; P p; bool has_old; uint32_t old; };
    std::vector<Act> acts;
    for (auto &route : partial_rib.r) {
        bool has_old = false; uint32_t old_label = 0;
        Route *o = old_rib.get(route.prefix);
        if (o) {
            const Route old = *o;
            old_rib.r.erase(old_rib.r.begin() + (o - old_rib.r.data()));
            has_old = old.has_label; old_label = old.label;
            if (old.report_metric() == route.report_metric() && old.tag == route.tag && old.has_label == route.has_label &&
                (!old.has_label || old.label == route.label) && old.nexthops.same(route.nexthops)) {
                if (old.flags & HL_ROUTE_INSTALLED) route.flags |= HL_ROUTE_INSTALLED;
                continue;
            }
        }
        if (!(route.flags & HL_ROUTE_CONNECTED) && !route.nexthops.v.empty()) {
            acts.push_back({HL_RIB_INSTALL, route.prefix, has_old, old_label});
            route.flags |= HL_ROUTE_INSTALLED;
        } else if (route.flags & HL_ROUTE_INSTALLED) {
            acts.push_back({HL_RIB_UNINSTALL, route.prefix, false, 0});
            route.flags &= (uint8_t)~HL_ROUTE_INSTALLED;
        }
    }
    for (auto &x : old_rib.r) if (x.flags & HL_ROUTE_INSTALLED) acts.push_back({HL_RIB_UNINSTALL_OLD, x.prefix, false, 0});
    for (auto &x : partial_rib.r) rib.put(x);                                   // rib.extend(partial_rib)

    uint32_t n_h = 0;
    for (auto &x : rib.r) n_h += (uint32_t)x.nexthops.v.size();
    out->n_routes = (uint32_t)rib.r.size(); out->n_nexthops = n_h; *n_actions = (uint32_t)acts.size();
    const int trc = put_tables(w.routers, areas, n_areas, out_rtrs);
    if (out->n_routes > out->routes_cap || n_h > out->nexthops_cap || acts.size() > cap || trc) return HSPF_E_NOMEM;
    uint32_t h = 0;
    for (size_t i = 0; i < rib.r.size(); ++i) {
        const Route &x = rib.r[i];
        hl_rib_route o;
        std::memset(&o, 0, sizeof(o));
        o.prefix = x.prefix.addr; o.mask = x.prefix.len ? 0xFFFFFFFFu << (32 - x.prefix.len) : 0u;
        o.metric = x.metric; o.type2_metric = x.type2_metric; o.tag = x.tag; o.area_id = x.area_id; o.path_type = x.path_type; o.flags = x.flags;
        o.has_area = x.has_area; o.has_type2 = x.has_type2; o.has_sr_label = x.has_label; o.sr_label = x.has_label ? x.label : 0;
        o.nh_off = h; o.n_nh = (uint32_t)x.nexthops.v.size();
        for (auto &n : x.nexthops.v) out->nexthops[h++] = n;
        out->routes[i] = o;
    }
    for (size_t k = 0; k < acts.size(); ++k) {
        hl_rib_action a;
        std::memset(&a, 0, sizeof(a));
        a.kind = acts[k].kind; a.has_old_sr_label = acts[k].has_old; a.old_sr_label = acts[k].has_old ? acts[k].old : 0;
        if (acts[k].kind == HL_RIB_UNINSTALL_OLD) {
            for (uint32_t i = 0; i < prev->n_routes; ++i)
                if (prev->routes[i].prefix == acts[k].p.addr && plen(prev->routes[i].mask) == acts[k].p.len) a.route = i;
        } else {
            for (size_t i = 0; i < rib.r.size(); ++i) if (rib.r[i].prefix == acts[k].p) a.route = (uint32_t)i;
        }
        actions[k] = a;
    }
    return 0;
}
