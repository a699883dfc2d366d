// TEST INFRASTRUCTURE — CPU restatement of the OSPF routing-table stages that follow
// the per-area SPFs in the reference (generic over the OSPF version there, a template here): update_rib_full (holo-ospf/src/route.rs:146-193) with
// update_rib_inter_area_networks (:449-533), update_rib_inter_area_routers (:653-714),
// update_rib_transit_area (:535-650), update_rib_external (:717-827), update_global_rib (:833-893),
// route_update (:895-942) and route_compare (:944-971).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
// may use anything under oracle/.
//
// Pinned by: the 63 golden OSPFv2 snapshots of the reference's conformance topologies
// (tests/golden/ospfv2.json): with the per-area results of oracle_ospfv2_run_area as input,
// every route of `local-rib` — intra-area and the 269 inter-area ones, including the next
// hops that virtual-link end points obtain from their transit area — is reproduced
// (tests/test_oracle_golden.py); and by the 44 golden OSPFv3 snapshots (tests/golden/ospfv3.json,
// 248 intra-area + 192 inter-area routes, Inter-Area-Prefix-LSAs).  The goldens contain no
// type-4 and no AS-external LSAs: the inter-area-router and AS-external stages are restated
// but PARITY UNPINNED.
//
// Deviation kept on purpose (documented in include/holo_spf_lsdb.h): the per-area intra-area
// routes arrive already merged per area, so the shared-table walk of update_rib_intra_area
// across areas (route.rs:156-160) is applied per route, not per stub link: an area's route whose
// LS origin is a transit network meets the table as that network vertex would (route.rs:371-384:
// kept out when its LSA id is lower than the entry's origin, otherwise it replaces the entry).
#include <cstdint>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "../include/holo_lsdb.h"
#include "../include/holo_spf.h"

namespace {

// IpNetwork order: address, then prefix length (ipnetwork derive(Ord)).  IPv4 addresses are
// stored in the first four bytes, network byte order, so one key type serves both versions.
struct Prefix {
    uint8_t addr[16];
    uint8_t len;
    bool operator<(const Prefix &o) const {
        const int c = std::memcmp(addr, o.addr, 16);
        return c != 0 ? c < 0 : len < o.len;
    }
};

// NexthopKey (route.rs:92-98): (iface_idx, addr) with None < Some; iface order = the
// explicit sort key (SURVEY.md §8a semantics #6).
struct NhKey {
    uint32_t iface;
    uint8_t has_addr;
    uint8_t addr[16];
    bool operator<(const NhKey &o) const {
        if (iface != o.iface) return iface < o.iface;
        if (has_addr != o.has_addr) return has_addr < o.has_addr;
        return std::memcmp(addr, o.addr, 16) < 0;
    }
};

struct V2 {   // Ospfv2
    using Area = hl_ospfv2_rib_area; using Sum = hl_ospfv2_summary_lsa; using Ext = hl_ospfv2_external_lsa;
    using Rib = hl_ospfv2_rib; using Nh = hl_nexthop; using Out = hl_rib_route;
    static Prefix v4(uint32_t a, uint32_t mask) {
        Prefix p{};
        p.addr[0] = a >> 24; p.addr[1] = a >> 16; p.addr[2] = a >> 8; p.addr[3] = a;
        p.len = (uint8_t)__builtin_popcount(mask);
        return p;
    }
    static Prefix of(const hl_route_net &r) { return v4(r.prefix, r.mask); }
    // Summary / external prefixes are with_netmask(lsa_id, mask) WITHOUT apply_mask
    // (ospfv2/spf.rs:552,602), unlike the intra-area stubs (:479,512).
    static Prefix of(const Sum &l) { return v4(l.lsa_id, l.mask); }
    static Prefix of(const Ext &l) { return v4(l.lsa_id, l.mask); }
    static bool nu(const Sum &) { return false; }
    static bool nu(const Ext &) { return false; }
    static uint32_t asbr(const Sum &l) { return l.lsa_id; }                 // ospfv2/spf.rs:583
    static uint8_t opts(const hl_route_net &) { return 0; }
    static uint8_t opts(const Sum &) { return 0; }
    static uint8_t opts(const Ext &) { return 0; }
    static void label(bool &has, uint32_t &lab, const hl_route_net &r) { has = r.has_sr_label; lab = r.sr_label; }
    static void put_label(Out &o, bool has, uint32_t lab) { o.has_sr_label = has; o.sr_label = has ? lab : 0; }
    static bool label_eq(const Out &a, const Out &b) {          // Option<Label> ==
        return a.has_sr_label == b.has_sr_label && (!a.has_sr_label || a.sr_label == b.sr_label);
    }
    static void old_label(hl_rib_action &x, const Out &o) { x.has_old_sr_label = o.has_sr_label; x.old_sr_label = o.has_sr_label ? o.sr_label : 0; }
    static Prefix of(const Out &o) { return v4(o.prefix, o.mask); }
    static bool nh_eq(const Nh &a, const Nh &b) {               // Nexthop ==, field by field
        return a.iface == b.iface && a.has_addr == b.has_addr && (!a.has_addr || a.addr == b.addr) &&
               a.has_nbr == b.has_nbr && (!a.has_nbr || a.nbr_router_id == b.nbr_router_id) &&
               a.has_label == b.has_label && (!a.has_label || a.sr_label == b.sr_label);
    }
    static NhKey key(uint32_t sk, const Nh &n) {
        NhKey k{};
        k.iface = sk; k.has_addr = n.has_addr;
        if (n.has_addr) { k.addr[0] = n.addr >> 24; k.addr[1] = n.addr >> 16; k.addr[2] = n.addr >> 8; k.addr[3] = n.addr; }
        return k;
    }
    static void emit(Out &o, const Prefix &p, uint8_t) {
        o.prefix = ((uint32_t)p.addr[0] << 24) | ((uint32_t)p.addr[1] << 16) | ((uint32_t)p.addr[2] << 8) | p.addr[3];
        o.mask = p.len ? 0xFFFFFFFFu << (32 - p.len) : 0;
    }
};

struct V3 {   // Ospfv3
    using Area = hl_ospfv3_rib_area; using Sum = hl_ospfv3_inter_area_lsa; using Ext = hl_ospfv3_external_lsa;
    using Rib = hl_ospfv3_rib; using Nh = hl_nexthop6; using Out = hl_rib_route6;
    static Prefix v6(const hl_ip_addr &a, uint8_t len) { Prefix p{}; std::memcpy(p.addr, a.bytes, 16); p.len = len; return p; }
    static Prefix of(const hl_route_net6 &r) { return v6(r.prefix, r.len); }
    static Prefix of(const Sum &l) { return v6(l.prefix, l.len); }
    static Prefix of(const Ext &l) { return v6(l.prefix, l.len); }
    static bool nu(const Sum &l) { return l.lsa_type == 3 && (l.prefix_options & HL_PFX_OPT_NU); }   // ospfv3/spf.rs:494
    static bool nu(const Ext &l) { return l.prefix_options & HL_PFX_OPT_NU; }                          // ospfv3/spf.rs:538
    static uint32_t asbr(const Sum &l) { return l.router_id; }              // ospfv3/spf.rs:520
    static uint8_t opts(const hl_route_net6 &r) { return r.prefix_options; }
    static uint8_t opts(const Sum &l) { return l.prefix_options; }
    static uint8_t opts(const Ext &l) { return l.prefix_options; }
    static void label(bool &, uint32_t &, const hl_route_net6 &) {}
    static void put_label(Out &, bool, uint32_t) {}
    static bool label_eq(const Out &, const Out &) { return true; }
    static void old_label(hl_rib_action &, const Out &) {}
    static Prefix of(const Out &o) { return v6(o.prefix, o.len); }
    static bool nh_eq(const Nh &a, const Nh &b) {
        return a.iface == b.iface && a.has_addr == b.has_addr && (!a.has_addr || std::memcmp(a.addr.bytes, b.addr.bytes, 16) == 0) &&
               a.has_nbr == b.has_nbr && (!a.has_nbr || a.nbr_router_id == b.nbr_router_id);
    }
    static NhKey key(uint32_t sk, const Nh &n) {
        NhKey k{};
        k.iface = sk; k.has_addr = n.has_addr;
        if (n.has_addr) std::memcpy(k.addr, n.addr.bytes, 16);
        return k;
    }
    static void emit(Out &o, const Prefix &p, uint8_t options) {
        std::memcpy(o.prefix.bytes, p.addr, 16);
        o.prefix.is_v6 = 1; o.len = p.len; o.prefix_options = options;
    }
};

template <class V>
struct Stage {
    using Nexthops = std::map<NhKey, typename V::Nh>;

    struct RouteNet {   // route.rs:32-46, the fields these stages touch
        uint8_t path_type = HL_PATH_INTRA_AREA;
        bool has_area = false;
        uint32_t area_id = 0;
        uint32_t metric = 0;
        bool has_type2 = false;
        uint32_t type2_metric = 0;
        uint32_t tag = 0;
        uint8_t flags = 0;
        uint8_t prefix_options = 0;
        bool has_label = false;      // sr_label of an intra-area route
        uint32_t label = 0;
        uint32_t origin_lsa_id = 0;  // origin.lsa_id of an intra-area route
        Nexthops nexthops;
    };
    struct RouteRtr {   // route.rs:57-66
        uint32_t area_id = 0;
        uint8_t path_type = HL_PATH_INTRA_AREA;
        uint8_t flags = 0;
        uint32_t metric = 0;
        Nexthops nexthops;
    };
    using Rib = std::map<Prefix, RouteNet>;

    static int route_compare(const RouteNet &a, const RouteNet &b) {   // route.rs:944-971
        if (a.path_type != b.path_type) return a.path_type < b.path_type ? -1 : 1;
        auto cmp = [](uint32_t x, uint32_t y) { return x < y ? -1 : (x > y ? 1 : 0); };
        if (a.path_type == HL_PATH_TYPE2_EXTERNAL) {
            // Option<u32> order: None < Some
            if (a.has_type2 != b.has_type2) return a.has_type2 ? 1 : -1;
            if (int c = cmp(a.type2_metric, b.type2_metric)) return c;
        }
        return cmp(a.metric, b.metric);
    }

    static void truncate(RouteNet &r, uint32_t max_paths) {   // route.rs:934-941
        while (r.nexthops.size() > max_paths) r.nexthops.erase(std::prev(r.nexthops.end()));
    }

    static void route_update(Rib &rib, const Prefix &p, RouteNet route, uint32_t max_paths) {   // route.rs:895-942
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

    static Nexthops lift(const typename V::Area &a, uint32_t off, uint32_t n) {
        Nexthops m;
        for (uint32_t i = 0; i < n; ++i) {
            typename V::Nh nh = a.spf->nexthops[off + i];
            const uint32_t key = nh.iface < a.n_ifaces ? a.ifaces[nh.iface].sort_key : 0xFFFFFFFFu;
            nh.iface = key;
            m[V::key(key, nh)] = nh;
        }
        return m;
    }

    static int run(uint32_t router_id, uint32_t max_paths, const typename V::Area *areas, uint32_t n_areas,
                   const typename V::Ext *ext, uint32_t n_ext, typename V::Rib *out) {
        if ((!areas && n_areas) || !out) return HSPF_E_INVAL;
        Rib rib;
        std::vector<std::map<uint32_t, RouteRtr>> routers(n_areas);

        // ---- intra-area routes of every area into one table (route.rs:156-160) and the
        //      per-area router tables (area.state.routers, spf.rs:627-637)
        for (uint32_t ai = 0; ai < n_areas; ++ai) {
            const auto &a = areas[ai];
            if (!a.spf) return HSPF_E_INVAL;
            for (uint32_t i = 0; i < a.spf->n_routers; ++i) {
                const hl_route_rtr &r = a.spf->routers[i];
                RouteRtr e;
                e.area_id = a.area_id; e.path_type = HL_PATH_INTRA_AREA; e.flags = r.flags; e.metric = r.metric;
                e.nexthops = lift(a, r.nh_off, r.n_nh);
                routers[ai][r.router_id] = std::move(e);
            }
            for (uint32_t i = 0; i < a.spf->n_routes; ++i) {
                const auto &r = a.spf->routes[i];
                const Prefix p = V::of(r);
                auto it = rib.find(p);
                if (it != rib.end() && r.metric > it->second.metric) continue;   // route.rs:372-376
                if (it != rib.end() && r.origin_type == 2) {                     // route.rs:387-397
                    if (r.origin_lsa_id < it->second.origin_lsa_id) continue;
                    rib.erase(it);
                }
                RouteNet n;
                n.path_type = HL_PATH_INTRA_AREA; n.has_area = true; n.area_id = a.area_id; n.metric = r.metric;
                n.origin_lsa_id = r.origin_lsa_id;
                n.flags = r.flags; n.prefix_options = V::opts(r); n.nexthops = lift(a, r.nh_off, r.n_nh);
                V::label(n.has_label, n.label, r);
                route_update(rib, p, std::move(n), max_paths);
            }
        }

        // ---- inter-area routes (route.rs:163-179)
        uint32_t active_areas = 0;
        for (uint32_t ai = 0; ai < n_areas; ++ai) active_areas += areas[ai].active ? 1 : 0;
        for (uint32_t ai = 0; ai < n_areas; ++ai) {
            const auto &a = areas[ai];
            // several active areas: only backbone summary-LSAs are examined
            if (active_areas > 1 && a.area_id != 0) continue;
            for (int pass = 0; pass < 2; ++pass) {          // networks (type 3), then routers (type 4)
                for (uint32_t i = 0; i < a.n_summaries; ++i) {
                    const auto &l = a.summaries[i];
                    if (l.lsa_type != (pass == 0 ? 3 : 4) || l.maxage || V::nu(l)) continue;
                    if (!(l.metric < HL_LSA_INFINITY) || l.adv_rtr == router_id) continue;
                    auto br = routers[ai].find(l.adv_rtr);
                    if (br == routers[ai].end() || !(br->second.flags & HL_RTR_FLAG_B)) continue;   // no ABR entry
                    const uint32_t metric = br->second.metric + l.metric;
                    if (pass == 0) {
                        RouteNet n;
                        n.path_type = HL_PATH_INTER_AREA; n.has_area = true; n.area_id = a.area_id; n.metric = metric;
                        n.prefix_options = V::opts(l); n.nexthops = br->second.nexthops;
                        route_update(rib, V::of(l), std::move(n), max_paths);
                    } else {
                        RouteRtr e;     // routers.insert(): replaces whatever was there (route.rs:713)
                        e.area_id = a.area_id; e.path_type = HL_PATH_INTER_AREA; e.flags = HL_RTR_FLAG_E; e.metric = metric;
                        e.nexthops = br->second.nexthops;
                        routers[ai][V::asbr(l)] = std::move(e);
                    }
                }
            }
        }

        // ---- transit areas: shorter paths through them, and the next hops of virtual links
        //      (route.rs:181-187, 535-650)
        for (uint32_t ai = 0; ai < n_areas; ++ai) {
            const auto &a = areas[ai];
            if (!a.spf->transit_capability) continue;
            for (uint32_t i = 0; i < a.n_summaries; ++i) {
                const auto &l = a.summaries[i];
                if (l.lsa_type != 3 || l.maxage || V::nu(l)) continue;
                if (!(l.metric < HL_LSA_INFINITY) || l.adv_rtr == router_id) continue;
                auto cur = rib.find(V::of(l));
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
                    n.prefix_options = V::opts(l); n.nexthops = br->second.nexthops;
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
            const auto &l = ext[i];
            if (l.maxage || V::nu(l) || !(l.metric < HL_LSA_INFINITY) || l.adv_rtr == router_id) continue;
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
            n.has_area = false; n.tag = l.tag; n.prefix_options = V::opts(l); n.nexthops = best->nexthops;
            if (l.e_bit) { n.path_type = HL_PATH_TYPE2_EXTERNAL; n.metric = best->metric; n.has_type2 = true; n.type2_metric = l.metric; }
            else { n.path_type = HL_PATH_TYPE1_EXTERNAL; n.metric = best->metric + l.metric; }
            route_update(rib, V::of(l), std::move(n), max_paths);
        }

        // ---- emit
        uint32_t need_h = 0;
        for (auto &kv : rib) need_h += (uint32_t)kv.second.nexthops.size();
        out->n_routes = (uint32_t)rib.size();
        out->n_nexthops = need_h;
        if (out->n_routes > out->routes_cap || need_h > out->nexthops_cap) return HSPF_E_NOMEM;
        uint32_t ri = 0, h = 0;
        for (auto &kv : rib) {
            typename V::Out o;
            std::memset(&o, 0, sizeof(o));
            V::emit(o, kv.first, kv.second.prefix_options);
            o.metric = kv.second.metric;
            o.type2_metric = kv.second.type2_metric; o.has_type2 = kv.second.has_type2; o.tag = kv.second.tag;
            o.area_id = kv.second.area_id; o.has_area = kv.second.has_area; o.path_type = kv.second.path_type;
            o.flags = kv.second.flags; o.nh_off = h; o.n_nh = (uint32_t)kv.second.nexthops.size();
            V::put_label(o, kv.second.has_label, kv.second.label);
            for (auto &nk : kv.second.nexthops) out->nexthops[h++] = nk.second;
            out->routes[ri++] = o;
        }
        return HSPF_OK;
    }
};

// update_global_rib (route.rs:833-893): which installs / uninstalls the RIB manager sees
template <class V>
int global_rib(const typename V::Rib *old_rib, typename V::Rib *rib, hl_rib_action *out, uint32_t cap, uint32_t *n_out) {
    std::map<Prefix, uint32_t> old;                       // mut old_rib
    if (old_rib)
        for (uint32_t i = 0; i < old_rib->n_routes; ++i) old[V::of(old_rib->routes[i])] = i;
    std::vector<hl_rib_action> acts;
    auto metric = [](const typename V::Out &r) { return r.path_type == HL_PATH_TYPE2_EXTERNAL ? r.type2_metric : r.metric; };
    for (uint32_t i = 0; i < rib->n_routes; ++i) {        // the new table is in prefix order
        typename V::Out &route = rib->routes[i];
        hl_rib_action a{};
        a.route = i;
        auto it = old.find(V::of(route));
        if (it != old.end()) {
            const typename V::Out &o = old_rib->routes[it->second];
            old.erase(it);                                 // old_rib.remove(prefix)
            V::old_label(a, o);
            bool same = metric(o) == metric(route) && o.tag == route.tag && V::label_eq(o, route) && o.n_nh == route.n_nh;
            for (uint32_t k = 0; same && k < route.n_nh; ++k)
                same = V::nh_eq(old_rib->nexthops[o.nh_off + k], rib->nexthops[route.nh_off + k]);
            if (same) {                                    // skip reinstalling
                if (o.flags & HL_ROUTE_INSTALLED) route.flags |= HL_ROUTE_INSTALLED;
                continue;
            }
        }
        if (!(route.flags & HL_ROUTE_CONNECTED) && route.n_nh != 0) {
            a.kind = HL_RIB_INSTALL;
            acts.push_back(a);
            route.flags |= HL_ROUTE_INSTALLED;
        } else if (route.flags & HL_ROUTE_INSTALLED) {
            a.kind = HL_RIB_UNINSTALL; a.has_old_sr_label = 0; a.old_sr_label = 0;
            acts.push_back(a);
            route.flags &= (uint8_t)~HL_ROUTE_INSTALLED;
        }
    }
    for (auto &kv : old) {                                 // routes that are no longer available
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

}  // namespace

extern "C" int oracle_ospfv2_rib_diff(const hl_ospfv2_rib *o, hl_ospfv2_rib *n, hl_rib_action *out, uint32_t cap, uint32_t *n_out) {
    return global_rib<V2>(o, n, out, cap, n_out);
}
extern "C" int oracle_ospfv3_rib_diff(const hl_ospfv3_rib *o, hl_ospfv3_rib *n, hl_rib_action *out, uint32_t cap, uint32_t *n_out) {
    return global_rib<V3>(o, n, out, cap, n_out);
}

extern "C" int oracle_ospfv2_update_rib_full(uint32_t router_id, uint32_t max_paths, const hl_ospfv2_rib_area *areas,
                                             uint32_t n_areas, const hl_ospfv2_external_lsa *ext, uint32_t n_ext,
                                             hl_ospfv2_rib *out) {
    return Stage<V2>::run(router_id, max_paths, areas, n_areas, ext, n_ext, out);
}

extern "C" int oracle_ospfv3_update_rib_full(uint32_t router_id, uint32_t max_paths, const hl_ospfv3_rib_area *areas,
                                             uint32_t n_areas, const hl_ospfv3_external_lsa *ext, uint32_t n_ext,
                                             hl_ospfv3_rib *out) {
    return Stage<V3>::run(router_id, max_paths, areas, n_areas, ext, n_ext, out);
}
