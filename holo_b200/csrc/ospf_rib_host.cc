// OSPFv2 routing table after the per-area SPFs: hspf_ospfv2_update_rib_full
// (include/holo_spf_lsdb.h).  Host-side table joins over the results of
// hspf_ospfv2_run_area; replaces, for the stages after the SPT,
//   update_rib_full                  holo-ospf/src/route.rs:146-193
//   update_rib_inter_area_networks   route.rs:449-533   (type-3 Summary-LSAs)
//   update_rib_inter_area_routers    route.rs:653-714   (type-4 Summary-LSAs)
//   update_rib_transit_area          route.rs:535-650   (RFC 2328 16.3, virtual links)
//   update_rib_external              route.rs:717-827   (type-5 LSAs)
//   route_update / route_compare     route.rs:895-971
// Routes live in one flat vector with a hash index on (address, mask); next-hop sets are
// small vectors kept sorted by NexthopKey (interface sort key, then address with None first),
// merged like BTreeMap::extend (a later entry with the same key replaces the earlier one).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <new>
#include <unordered_map>
#include <vector>

#include "../../include/holo_lsdb.h"
#include "../../include/holo_spf.h"
#include "../../include/holo_spf_lsdb.h"

namespace {

struct Hop {
    uint64_t ka;         // NexthopKey, major part: sort_key << 1 | has_addr   (None sorts first)
    uint32_t kb;         // minor part: the address
    hl_nexthop nh;
};
using Hops = std::vector<Hop>;   // ascending key, unique

inline bool key_less(const Hop &p, const Hop &q) { return p.ka != q.ka ? p.ka < q.ka : p.kb < q.kb; }
inline bool key_same(const Hop &p, const Hop &q) { return p.ka == q.ka && p.kb == q.kb; }

// dst <- dst extended with src (same key: src wins)
void merge_hops(Hops &dst, const Hops &src) {
    Hops out;
    out.reserve(dst.size() + src.size());
    size_t i = 0, j = 0;
    while (i < dst.size() || j < src.size()) {
        if (j == src.size() || (i < dst.size() && key_less(dst[i], src[j]))) out.push_back(dst[i++]);
        else if (i == dst.size() || key_less(src[j], dst[i])) out.push_back(src[j++]);
        else { out.push_back(src[j++]); ++i; }
    }
    dst.swap(out);
}

void clip(Hops &h, uint32_t max_paths) {
    if (h.size() > max_paths) h.resize(max_paths);
}

struct Net {
    uint32_t prefix, mask;
    uint32_t metric, type2, tag, area;
    uint8_t path, flags;
    bool has_area, has_type2;
    Hops hops;
};

struct Rtr {
    uint32_t area, metric;
    uint8_t path, flags;
    Hops hops;
};

// negative: a preferred over b
int prefer(const Net &a, const Net &b) {
    if (a.path != b.path) return a.path < b.path ? -1 : 1;
    if (a.path == HL_PATH_TYPE2_EXTERNAL) {
        if (a.has_type2 != b.has_type2) return a.has_type2 ? 1 : -1;     // None < Some
        if (a.type2 != b.type2) return a.type2 < b.type2 ? -1 : 1;
    }
    if (a.metric != b.metric) return a.metric < b.metric ? -1 : 1;
    return 0;
}

struct Table {
    std::vector<Net> nets;
    std::unordered_map<uint64_t, uint32_t> index;
    uint32_t max_paths;

    static uint64_t k(uint32_t prefix, uint32_t mask) { return ((uint64_t)prefix << 32) | mask; }
    Net *find(uint32_t prefix, uint32_t mask) {
        auto it = index.find(k(prefix, mask));
        return it == index.end() ? nullptr : &nets[it->second];
    }
    void offer(Net &&n) {          // route_update
        Net *cur = find(n.prefix, n.mask);
        if (!cur) {
            index.emplace(k(n.prefix, n.mask), (uint32_t)nets.size());
            nets.push_back(std::move(n));
            cur = &nets.back();
        } else {
            const int c = prefer(n, *cur);
            if (c < 0) *cur = std::move(n);
            else if (c == 0) merge_hops(cur->hops, n.hops);
        }
        clip(cur->hops, max_paths);
    }
};

Hops hops_of(const hl_ospfv2_rib_area &a, uint32_t off, uint32_t n) {
    Hops h;
    h.reserve(n);
    for (uint32_t i = 0; i < n; ++i) {
        Hop x;
        x.nh = a.spf->nexthops[off + i];
        const uint32_t sk = x.nh.iface < a.n_ifaces ? a.ifaces[x.nh.iface].sort_key : 0xFFFFFFFFu;
        x.nh.iface = sk;                       // the merged table names interfaces by sort key
        x.ka = ((uint64_t)sk << 1) | (x.nh.has_addr ? 1u : 0u);
        x.kb = x.nh.has_addr ? x.nh.addr : 0u;
        h.push_back(x);
    }
    std::stable_sort(h.begin(), h.end(), key_less);
    Hops u;                                    // same key twice: the later one stays
    for (const Hop &x : h) {
        if (!u.empty() && key_same(u.back(), x)) u.back() = x;
        else u.push_back(x);
    }
    return u;
}

bool usable(const hl_ospfv2_summary_lsa &l, uint32_t router_id) {
    return !l.maxage && l.metric < HL_LSA_INFINITY && l.adv_rtr != router_id;
}

}  // namespace

extern "C" int hspf_ospfv2_update_rib_full(uint32_t router_id, uint32_t max_paths, const hl_ospfv2_rib_area *areas,
                                           uint32_t n_areas, const hl_ospfv2_external_lsa *ext, uint32_t n_ext,
                                           hl_ospfv2_rib *out) {
    if ((!areas && n_areas) || (!ext && n_ext) || !out) return HSPF_E_INVAL;
    try {
        for (uint32_t ai = 0; ai < n_areas; ++ai) {
            const hl_ospfv2_rib_area &a = areas[ai];
            if (!a.spf || (a.n_ifaces && !a.ifaces) || (a.n_summaries && !a.summaries)) return HSPF_E_INVAL;
        }
        Table t;
        t.max_paths = max_paths;
        std::vector<std::unordered_map<uint32_t, Rtr>> rtrs(n_areas);

        // 1. per-area router tables; intra-area routes of all areas into one table
        for (uint32_t ai = 0; ai < n_areas; ++ai) {
            const hl_ospfv2_rib_area &a = areas[ai];
            for (uint32_t i = 0; i < a.spf->n_routers; ++i) {
                const hl_route_rtr &r = a.spf->routers[i];
                Rtr e{a.area_id, r.metric, HL_PATH_INTRA_AREA, r.flags, hops_of(a, r.nh_off, r.n_nh)};
                rtrs[ai][r.router_id] = std::move(e);
            }
            for (uint32_t i = 0; i < a.spf->n_routes; ++i) {
                const hl_route_net &r = a.spf->routes[i];
                if (const Net *cur = t.find(r.prefix, r.mask))
                    if (r.metric > cur->metric) continue;
                Net n{r.prefix, r.mask, r.metric, 0, 0, a.area_id, HL_PATH_INTRA_AREA, r.flags, true, false,
                      hops_of(a, r.nh_off, r.n_nh)};
                t.offer(std::move(n));
            }
        }

        // 2. Summary-LSAs: only the backbone's when more than one area is active
        uint32_t n_active = 0;
        for (uint32_t ai = 0; ai < n_areas; ++ai) n_active += areas[ai].active ? 1u : 0u;
        auto abr = [&](uint32_t ai, uint32_t adv) -> const Rtr * {
            auto it = rtrs[ai].find(adv);
            return (it != rtrs[ai].end() && (it->second.flags & HL_RTR_FLAG_B)) ? &it->second : nullptr;
        };
        for (uint32_t ai = 0; ai < n_areas; ++ai) {
            const hl_ospfv2_rib_area &a = areas[ai];
            if (n_active > 1 && a.area_id != 0) continue;
            for (uint32_t i = 0; i < a.n_summaries; ++i) {            // networks
                const hl_ospfv2_summary_lsa &l = a.summaries[i];
                if (l.lsa_type != 3 || !usable(l, router_id)) continue;
                const Rtr *br = abr(ai, l.adv_rtr);
                if (!br) continue;
                Net n{l.lsa_id, l.mask, br->metric + l.metric, 0, 0, a.area_id, HL_PATH_INTER_AREA, 0, true, false, br->hops};
                t.offer(std::move(n));
            }
            for (uint32_t i = 0; i < a.n_summaries; ++i) {            // ASBRs
                const hl_ospfv2_summary_lsa &l = a.summaries[i];
                if (l.lsa_type != 4 || !usable(l, router_id)) continue;
                const Rtr *br = abr(ai, l.adv_rtr);
                if (!br) continue;
                Rtr e{a.area_id, br->metric + l.metric, HL_PATH_INTER_AREA, HL_RTR_FLAG_E, br->hops};
                rtrs[ai][l.lsa_id] = std::move(e);                    // replaces any earlier entry
            }
        }

        // 3. transit areas
        for (uint32_t ai = 0; ai < n_areas; ++ai) {
            const hl_ospfv2_rib_area &a = areas[ai];
            if (!a.spf->transit_capability) continue;
            for (uint32_t i = 0; i < a.n_summaries; ++i) {
                const hl_ospfv2_summary_lsa &l = a.summaries[i];
                if (l.lsa_type != 3 || !usable(l, router_id)) continue;
                Net *cur = t.find(l.lsa_id, l.mask);
                if (!cur || cur->path > HL_PATH_INTER_AREA || !cur->has_area || cur->area != 0) continue;
                const Rtr *br = abr(ai, l.adv_rtr);
                if (!br) continue;
                const uint32_t metric = br->metric + l.metric;
                if (metric < cur->metric) {
                    const uint32_t p = cur->prefix, m = cur->mask;
                    *cur = Net{p, m, metric, 0, 0, a.area_id, HL_PATH_INTER_AREA, 0, true, false, br->hops};
                } else if (metric == cur->metric) {
                    merge_hops(cur->hops, br->hops);
                }
                clip(cur->hops, max_paths);
            }
        }

        // 4. AS-external LSAs through the best ASBR entry (areas in area-id order)
        std::vector<uint32_t> order(n_areas);
        for (uint32_t i = 0; i < n_areas; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return areas[x].area_id < areas[y].area_id; });
        for (uint32_t i = 0; i < n_ext; ++i) {
            const hl_ospfv2_external_lsa &l = ext[i];
            if (l.maxage || !(l.metric < HL_LSA_INFINITY) || l.adv_rtr == router_id) continue;
            const Rtr *best = nullptr;
            bool best_pref = false;          // intra-area through a non-backbone area
            for (uint32_t ai : order) {
                auto it = rtrs[ai].find(l.adv_rtr);
                if (it == rtrs[ai].end() || !(it->second.flags & HL_RTR_FLAG_E)) continue;
                const Rtr *r = &it->second;
                const bool pref = r->path == HL_PATH_INTRA_AREA && r->area != 0;
                if (!best || (pref && !best_pref)) { best = r; best_pref = pref; continue; }
                if (pref != best_pref) continue;
                if (r->metric < best->metric || (r->metric == best->metric && r->area > best->area)) best = r;
            }
            if (!best) continue;
            Net n{l.lsa_id, l.mask, 0, 0, l.tag, 0, 0, 0, false, false, best->hops};
            if (l.e_bit) { n.path = HL_PATH_TYPE2_EXTERNAL; n.metric = best->metric; n.has_type2 = true; n.type2 = l.metric; }
            else { n.path = HL_PATH_TYPE1_EXTERNAL; n.metric = best->metric + l.metric; }
            t.offer(std::move(n));
        }

        // emit in Ipv4Network order (address, then prefix length)
        std::vector<uint32_t> idx(t.nets.size());
        for (uint32_t i = 0; i < idx.size(); ++i) idx[i] = i;
        std::sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) {
            const Net &p = t.nets[x], &q = t.nets[y];
            return p.prefix != q.prefix ? p.prefix < q.prefix : p.mask < q.mask;
        });
        uint32_t n_h = 0;
        for (const Net &n : t.nets) n_h += (uint32_t)n.hops.size();
        out->n_routes = (uint32_t)t.nets.size();
        out->n_nexthops = n_h;
        if (out->n_routes > out->routes_cap || n_h > out->nexthops_cap) return HSPF_E_NOMEM;
        if ((out->n_routes && !out->routes) || (n_h && !out->nexthops)) return HSPF_E_INVAL;
        uint32_t h = 0;
        for (uint32_t r = 0; r < idx.size(); ++r) {
            const Net &n = t.nets[idx[r]];
            hl_rib_route o;
            std::memset(&o, 0, sizeof(o));
            o.prefix = n.prefix; o.mask = n.mask; o.metric = n.metric; o.type2_metric = n.type2; o.tag = n.tag;
            o.area_id = n.area; o.path_type = n.path; o.flags = n.flags; o.has_area = n.has_area ? 1 : 0;
            o.has_type2 = n.has_type2 ? 1 : 0; o.nh_off = h; o.n_nh = (uint32_t)n.hops.size();
            for (const Hop &x : n.hops) out->nexthops[h++] = x.nh;
            out->routes[r] = o;
        }
        return HSPF_OK;
    } catch (const std::bad_alloc &) {
        return HSPF_E_NOMEM;
    } catch (...) {
        return HSPF_E_INVAL;
    }
}
