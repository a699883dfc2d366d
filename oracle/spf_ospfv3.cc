// oracle/spf_ospfv3.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Line-faithful CPU restatement of holo's OSPFv3 SPF path over the flat LSDB image of
// include/holo_lsdb.h:
//   run_area / calc_nexthops     holo-ospf/src/spf.rs:587-767 (version generic)
//   VertexId ordering            holo-ospf/src/ospfv3/spf.rs:37-41
//   Ospfv3::calc_nexthops        holo-ospf/src/ospfv3/spf.rs:164-283
//   vertex_lsa_find              holo-ospf/src/ospfv3/spf.rs:285-342 (Router vertex = ALL fragments)
//   vertex_lsa_links             holo-ospf/src/ospfv3/spf.rs:344-418 (enumerate BEFORE the filter)
//   intra_area_networks          holo-ospf/src/ospfv3/spf.rs:420-477 (Intra-Area-Prefix-LSAs)
//   calc_nexthop_lladdr          holo-ospf/src/ospfv3/spf.rs:592-611 (neighbour's Link-LSA)
//   update_rib_intra_area / route_update   holo-ospf/src/route.rs:343-446, 895-942
// Ordered std::map stands in for BTreeMap; linear candidate lookup and the per-edge
// mutual-link re-walk are kept.  SR / BIER for OSPFv3 are not restated.
//
// Parity pinning: tests/test_oracle_golden.py reproduces the reference's golden OSPFv3
// local-ribs (tests/golden/ospfv3.json, from
// holo-ospf/tests/conformance/ospfv3/topologies/*/rt*/output/northbound-state.json).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <optional>
#include <tuple>
#include <utility>
#include <vector>

#include "../include/holo_lsdb.h"
#include "../include/holo_spf.h"

namespace {

struct VertexId {   // Network{router_id, iface_id} < Router{router_id}
    bool is_router; uint32_t router_id; uint32_t iface_id;
    bool operator<(const VertexId &o) const {
        return std::tie(is_router, router_id, iface_id) < std::tie(o.is_router, o.router_id, o.iface_id);
    }
    bool operator==(const VertexId &o) const { return is_router == o.is_router && router_id == o.router_id && iface_id == o.iface_id; }
};

struct VertexLsa { bool is_router; std::vector<uint32_t> idx; };   // router: all fragments; network: one

struct Addr {
    hl_ip_addr a;
    bool operator<(const Addr &o) const {
        if (a.is_v6 != o.a.is_v6) return a.is_v6 < o.a.is_v6;
        return std::memcmp(a.bytes, o.a.bytes, 16) < 0;
    }
};
struct NexthopKey {
    uint32_t iface_sort; bool has_addr; Addr addr;
    bool operator<(const NexthopKey &o) const {
        if (iface_sort != o.iface_sort) return iface_sort < o.iface_sort;
        if (has_addr != o.has_addr) return has_addr < o.has_addr;
        return has_addr && addr < o.addr;
    }
};
struct Nexthop { uint32_t iface; bool has_addr; hl_ip_addr addr; bool has_nbr; uint32_t nbr; };
using Nexthops = std::map<NexthopKey, Nexthop>;

struct Vertex { VertexId id; VertexLsa lsa; uint16_t distance; uint16_t hops; Nexthops nexthops; };

struct SpfLink { bool has_parent; const hl_ospfv3_link *parent_link; VertexId id; VertexLsa lsa; uint16_t cost; };

struct Area {
    const hl_ospfv3_area *a;

    std::optional<VertexLsa> vertex_lsa_find(const VertexId &id) const {
        if (!id.is_router) {
            for (uint32_t i = 0; i < a->n_network_lsas; ++i) {
                const auto &n = a->network_lsas[i];
                if (n.adv_rtr == id.router_id && n.lsa_id == id.iface_id) {
                    if (n.age == HL_LSA_MAX_AGE) return std::nullopt;
                    return VertexLsa{false, {i}};
                }
            }
            return std::nullopt;
        }
        VertexLsa v{true, {}};
        for (uint32_t i = 0; i < a->n_router_lsas; ++i) {   // iter_by_type_advrtr: LsaKey order
            const auto &r = a->router_lsas[i];
            if (r.adv_rtr != id.router_id || r.age == HL_LSA_MAX_AGE) continue;
            if (!(r.options & HL_V3_OPT_R)) continue;
            if (a->af_ipv6 && !(r.options & HL_V3_OPT_V6)) continue;
            v.idx.push_back(i);
        }
        if (v.idx.empty()) return std::nullopt;
        return v;
    }

    void vertex_lsa_links(const VertexLsa &vl, const std::function<bool(const SpfLink &)> &f) const {
        if (!vl.is_router) {
            const auto &n = a->network_lsas[vl.idx[0]];
            for (uint32_t k = 0; k < n.n_att; ++k) {
                VertexId vid{true, a->attached[n.att_off + k], 0};
                auto l = vertex_lsa_find(vid);
                if (!l) continue;
                if (!f(SpfLink{false, nullptr, vid, *l, 0})) return;
            }
            return;
        }
        for (uint32_t li : vl.idx) {
            const auto &r = a->router_lsas[li];
            for (uint32_t k = 0; k < r.n_links; ++k) {
                const hl_ospfv3_link *link = &a->links[r.link_off + k];
                VertexId vid;
                if (link->link_type == HL_LINK_TRANSIT) vid = VertexId{false, link->nbr_router_id, link->nbr_iface_id};
                else vid = VertexId{true, link->nbr_router_id, 0};
                auto l = vertex_lsa_find(vid);
                if (!l) continue;
                if (!f(SpfLink{true, link, vid, *l, link->metric})) return;
            }
        }
    }

    int iface_by_ifindex(uint32_t ifindex) const {
        for (uint32_t i = 0; i < a->n_ifaces; ++i) if (a->ifaces[i].ifindex == ifindex) return (int)i;
        return -1;
    }

    // calc_nexthop_lladdr: neighbour's Link-LSA in the interface's link-scope LSDB
    std::optional<hl_ip_addr> lladdr(uint32_t iface, uint32_t nbr_router_id, uint32_t nbr_iface_id) const {
        for (uint32_t i = 0; i < a->n_link_lsas; ++i) {
            const auto &l = a->link_lsas[i];
            if (l.iface == iface && l.adv_rtr == nbr_router_id && l.lsa_id == nbr_iface_id) {
                if (l.age == HL_LSA_MAX_AGE) return std::nullopt;
                return l.linklocal;
            }
        }
        return std::nullopt;
    }

    bool v3_calc_nexthops(const Vertex &parent, const SpfLink &plink, const VertexLsa &dest_lsa, Nexthops &out) const {
        if (parent.lsa.is_router) {
            const hl_ospfv3_link *pl = plink.parent_link;
            int ii = iface_by_ifindex(pl->iface_id);
            if (ii < 0) return false;
            const auto &iface = a->ifaces[ii];
            if (iface.if_type == HL_IF_VLINK) return true;
            if (dest_lsa.is_router) {
                auto addr = lladdr((uint32_t)ii, pl->nbr_router_id, pl->nbr_iface_id);
                if (!addr) return false;
                uint32_t nbr = a->router_lsas[dest_lsa.idx[0]].adv_rtr;
                out[NexthopKey{iface.sort_key, true, Addr{*addr}}] = Nexthop{(uint32_t)ii, true, *addr, true, nbr};
            } else {
                out[NexthopKey{iface.sort_key, false, Addr{}}] = Nexthop{(uint32_t)ii, false, hl_ip_addr{}, false, 0};
            }
            return true;
        }
        const auto &pn = a->network_lsas[parent.lsa.idx[0]];
        const hl_ospfv3_link *dest_link = nullptr;
        for (uint32_t li : dest_lsa.idx) {
            const auto &r = a->router_lsas[li];
            for (uint32_t k = 0; k < r.n_links && !dest_link; ++k) {
                const auto &l = a->links[r.link_off + k];
                if (l.nbr_router_id == pn.adv_rtr && l.nbr_iface_id == pn.lsa_id) dest_link = &l;
            }
            if (dest_link) break;
        }
        if (!dest_link) return false;
        if (parent.nexthops.empty()) return false;
        uint32_t ii = parent.nexthops.begin()->second.iface;
        uint32_t nbr = a->router_lsas[dest_lsa.idx[0]].adv_rtr;
        auto addr = lladdr(ii, nbr, dest_link->iface_id);
        if (!addr) return false;
        out[NexthopKey{a->ifaces[ii].sort_key, true, Addr{*addr}}] = Nexthop{ii, true, *addr, true, nbr};
        return true;
    }
};

struct Prefix {
    hl_ip_addr addr; uint8_t len;
    bool operator<(const Prefix &o) const {
        if (addr.is_v6 != o.addr.is_v6) return addr.is_v6 < o.addr.is_v6;
        int c = std::memcmp(addr.bytes, o.addr.bytes, 16);
        if (c) return c < 0;
        return len < o.len;
    }
};

struct RouteNet { uint32_t metric; uint8_t flags, origin_type, options; uint32_t origin_adv, origin_id; Nexthops nexthops; };

}  // namespace

extern "C" int oracle_ospfv3_run_area(const hl_ospfv3_area *a, hl_ospfv3_result *out) {
    Area area{a};
    out->n_vertices = out->n_routers = out->n_routes = out->n_nexthops = 0;
    out->transit_capability = 0;
    out->root_found = 0;
    VertexId root_vid{true, a->router_id, 0};
    auto root_vlsa = area.vertex_lsa_find(root_vid);
    if (!root_vlsa) return 0;
    out->root_found = 1;

    std::map<VertexId, Vertex> spt;
    std::map<std::pair<uint16_t, VertexId>, Vertex> cand_list;
    cand_list.emplace(std::make_pair((uint16_t)0, root_vid), Vertex{root_vid, *root_vlsa, 0, 0, {}});
    std::map<uint32_t, std::tuple<uint32_t, uint8_t, uint8_t, Nexthops>> routers;
    bool transit_capability = false;

    while (!cand_list.empty()) {
        auto first = cand_list.begin();
        Vertex popped = std::move(first->second);
        cand_list.erase(first);
        VertexId vertex_id = popped.id;
        spt[vertex_id] = std::move(popped);
        const Vertex &vertex = spt[vertex_id];
        if (vertex.lsa.is_router) {
            // router_options / router_flags come from the first fragment
            const auto &rl = a->router_lsas[vertex.lsa.idx[0]];
            routers[rl.adv_rtr] = std::make_tuple((uint32_t)vertex.distance, rl.flags, rl.options, vertex.nexthops);
            if (rl.flags & HL_RTR_FLAG_V) transit_capability = true;
        }
        area.vertex_lsa_links(vertex.lsa, [&](const SpfLink &link) {
            bool back = false;
            area.vertex_lsa_links(link.lsa, [&](const SpfLink &l2) { if (l2.id == vertex.id) { back = true; return false; } return true; });
            if (!back) return true;
            if (spt.count(link.id)) return true;
            uint32_t s = (uint32_t)vertex.distance + link.cost;
            uint16_t distance = s > 0xFFFF ? 0xFFFF : (uint16_t)s;
            uint16_t hops = vertex.hops;
            if (link.lsa.is_router) hops = hops == 0xFFFF ? 0xFFFF : hops + 1;
            auto it = cand_list.begin();
            for (; it != cand_list.end(); ++it) if (it->second.id == link.id) break;
            if (it != cand_list.end()) {
                if (distance < it->second.distance) cand_list.erase(it);
                else if (distance > it->second.distance) return true;
            }
            auto key = std::make_pair(distance, link.id);
            auto ce = cand_list.find(key);
            if (ce == cand_list.end()) ce = cand_list.emplace(key, Vertex{link.id, link.lsa, distance, hops, {}}).first;
            Vertex &cand_v = ce->second;
            if (vertex.hops == 0) {
                Nexthops nh;
                if (area.v3_calc_nexthops(vertex, link, cand_v.lsa, nh))
                    for (auto &kv : nh) cand_v.nexthops[kv.first] = kv.second;
            } else {
                for (auto &kv : vertex.nexthops) cand_v.nexthops[kv.first] = kv.second;
            }
            return true;
        });
    }

    // ---- update_rib_intra_area over intra_area_networks (Intra-Area-Prefix-LSAs, LSDB order)
    std::map<Prefix, RouteNet> rib;
    for (uint32_t i = 0; i < a->n_iap_lsas; ++i) {
        const auto &l = a->iap_lsas[i];
        if (l.age == HL_LSA_MAX_AGE) continue;
        const Vertex *vertex = nullptr;
        if (l.ref_type == HL_V3_REF_ROUTER) {
            if (l.ref_lsa_id != 0) continue;
            auto it = spt.find(VertexId{true, l.ref_adv_rtr, 0});
            if (it != spt.end()) vertex = &it->second;
        } else if (l.ref_type == HL_V3_REF_NETWORK) {
            auto it = spt.find(VertexId{false, l.ref_adv_rtr, l.ref_lsa_id});
            if (it != spt.end()) vertex = &it->second;
        }
        if (!vertex) continue;
        for (uint32_t k = 0; k < l.n_prefixes; ++k) {
            const auto &px = a->prefixes[l.prefix_off + k];
            if (px.options & HL_PFX_OPT_NU) continue;
            Prefix prefix{px.addr, px.len};
            uint32_t s16 = (uint32_t)vertex->distance + px.metric;
            uint32_t metric = s16 > 0xFFFF ? 0xFFFF : s16;
            auto cur = rib.find(prefix);
            if (cur != rib.end() && metric > cur->second.metric) continue;
            // LS origin of the vertex: first fragment for routers
            uint8_t otype; uint32_t oadv, oid;
            if (vertex->lsa.is_router) { const auto &r = a->router_lsas[vertex->lsa.idx[0]]; otype = 1; oadv = r.adv_rtr; oid = r.lsa_id; }
            else { const auto &n = a->network_lsas[vertex->lsa.idx[0]]; otype = 2; oadv = n.adv_rtr; oid = n.lsa_id; }
            if (!vertex->lsa.is_router && cur != rib.end()) {
                if (metric > cur->second.metric || oid < cur->second.origin_id) continue;
                rib.erase(cur);
            }
            RouteNet nr{metric, (uint8_t)(vertex->hops == 0 ? HL_ROUTE_CONNECTED : 0), otype, px.options, oadv, oid, vertex->nexthops};
            auto ex = rib.find(prefix);
            RouteNet *route;
            if (ex != rib.end()) {
                RouteNet &cr = ex->second;
                if (nr.metric < cr.metric) cr = nr;
                else if (nr.metric == cr.metric) for (auto &kv : nr.nexthops) cr.nexthops[kv.first] = kv.second;
                route = &cr;
            } else {
                route = &rib.emplace(prefix, nr).first->second;
            }
            if (route->nexthops.size() > a->max_paths) {
                Nexthops cut; uint32_t n = 0;
                for (auto &kv : route->nexthops) { if (n++ >= a->max_paths) break; cut.insert(kv); }
                route->nexthops = std::move(cut);
            }
        }
    }

    // ---- export
    uint32_t need_h = 0;
    for (auto &kv : spt) need_h += (uint32_t)kv.second.nexthops.size();
    for (auto &kv : routers) need_h += (uint32_t)std::get<3>(kv.second).size();
    for (auto &kv : rib) need_h += (uint32_t)kv.second.nexthops.size();
    out->n_vertices = (uint32_t)spt.size(); out->n_routers = (uint32_t)routers.size();
    out->n_routes = (uint32_t)rib.size(); out->n_nexthops = need_h;
    out->transit_capability = transit_capability;
    if (out->n_vertices > out->vertices_cap || out->n_routers > out->routers_cap || out->n_routes > out->routes_cap ||
        need_h > out->nexthops_cap)
        return HSPF_E_NOMEM;
    uint32_t h = 0;
    auto put = [&](const Nexthops &n) {
        for (auto &kv : n) {
            hl_nexthop6 o{};
            o.iface = kv.second.iface; o.nbr_router_id = kv.second.has_nbr ? kv.second.nbr : 0;
            if (kv.second.has_addr) o.addr = kv.second.addr;
            o.has_addr = kv.second.has_addr; o.has_nbr = kv.second.has_nbr;
            out->nexthops[h++] = o;
        }
    };
    uint32_t i = 0;
    for (auto &kv : spt) {
        hl_spt_vertex6 o{};
        o.router_id = kv.first.router_id; o.iface_id = kv.first.iface_id; o.distance = kv.second.distance;
        o.hops = kv.second.hops; o.is_router = kv.first.is_router; o.nh_off = h; o.n_nh = (uint32_t)kv.second.nexthops.size();
        put(kv.second.nexthops);
        out->vertices[i++] = o;
    }
    i = 0;
    for (auto &kv : routers) {
        hl_route_rtr o{};
        o.router_id = kv.first; o.metric = std::get<0>(kv.second); o.flags = std::get<1>(kv.second);
        o.options = std::get<2>(kv.second); o.nh_off = h; o.n_nh = (uint32_t)std::get<3>(kv.second).size();
        put(std::get<3>(kv.second));
        out->routers[i++] = o;
    }
    i = 0;
    for (auto &kv : rib) {
        hl_route_net6 o{};
        o.prefix = kv.first.addr; o.len = kv.first.len; o.flags = kv.second.flags; o.origin_type = kv.second.origin_type;
        o.prefix_options = kv.second.options; o.metric = kv.second.metric; o.origin_adv_rtr = kv.second.origin_adv;
        o.origin_lsa_id = kv.second.origin_id; o.nh_off = h; o.n_nh = (uint32_t)kv.second.nexthops.size();
        put(kv.second.nexthops);
        out->routes[i++] = o;
    }
    return 0;
}

// Restates Ospfv3::spf_computation_type (holo-ospf/src/ospfv3/spf.rs:96-162): a Router-, Network-, Link- or
// Router-Information LSA among the triggers (function code normalised: extended LSAs count as their legacy twins)
// asks for a full run; otherwise partial, with BTreeSets of the prefixes of the changed Intra-Area-Prefix LSAs (new
// and old instance), Inter-Area-Prefix LSAs and AS-external LSAs, and of the routers of the Inter-Area-Router LSAs.
// PARITY UNPINNED: no reference test records this value.
#include <algorithm>
extern "C" int oracle_ospfv3_spf_computation_type(const hl_lsa_trigger6 *tr, uint32_t n, const hl_ip_prefix *prefixes,
                                                  uint32_t /*n_prefixes*/, hl_spf_computation6 *out) {
    auto code = [](uint16_t c) {
        switch (c) { case 33: return 1; case 34: return 2; case 35: return 3; case 36: return 4; case 37: return 5; case 40: return 8;
                     case 41: return 9; default: return (int)c; }
    };
    auto less = [](const hl_ip_prefix &a, const hl_ip_prefix &b) {
        if (a.addr.is_v6 != b.addr.is_v6) return a.addr.is_v6 < b.addr.is_v6;
        const int c = std::memcmp(a.addr.bytes, b.addr.bytes, 16);
        return c != 0 ? c < 0 : a.len < b.len;
    };
    auto same = [&](const hl_ip_prefix &a, const hl_ip_prefix &b) { return !less(a, b) && !less(b, a); };
    out->n_intra = out->n_inter_network = out->n_inter_router = out->n_external = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const int c = code(tr[i].function_code);
        if (c == 1 || c == 2 || c == 8 || c == 12) { out->kind = HL_SPF_FULL; return 0; }
    }
    out->kind = HL_SPF_PARTIAL;
    std::vector<hl_ip_prefix> intra, inter, ext;
    std::vector<uint32_t> rtr;
    for (uint32_t i = 0; i < n; ++i) {
        const int c = code(tr[i].function_code);
        const hl_ip_prefix *p = prefixes + tr[i].prefix_off;
        if (c == 9) intra.insert(intra.end(), p, p + tr[i].n_prefixes);
        if (c == 3 && tr[i].n_prefixes) inter.push_back(p[0]);
        if (c == 4) rtr.push_back(tr[i].router_id);
        if (c == 5 && tr[i].n_prefixes) ext.push_back(p[0]);
    }
    auto set_of = [&](std::vector<hl_ip_prefix> &v) {
        std::sort(v.begin(), v.end(), less);
        v.erase(std::unique(v.begin(), v.end(), same), v.end());
    };
    set_of(intra); set_of(inter); set_of(ext);
    std::sort(rtr.begin(), rtr.end());
    rtr.erase(std::unique(rtr.begin(), rtr.end()), rtr.end());
    for (auto &p : intra) out->intra[out->n_intra++] = p;
    for (auto &p : inter) out->inter_network[out->n_inter_network++] = p;
    for (uint32_t r : rtr) out->inter_router[out->n_inter_router++] = r;
    for (auto &p : ext) out->external[out->n_external++] = p;
    return 0;
}
