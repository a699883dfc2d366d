// oracle/spf_ospfv2.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Line-faithful CPU restatement of holo's OSPFv2 SPF path over the flat LSDB
// image of include/holo_lsdb.h:
//   run_area               holo-ospf/src/spf.rs:587-729
//   calc_nexthops          holo-ospf/src/spf.rs:733-767
//   Ospfv2::calc_nexthops  holo-ospf/src/ospfv2/spf.rs:173-354
//   vertex_lsa_find        holo-ospf/src/ospfv2/spf.rs:356-388
//   vertex_lsa_links       holo-ospf/src/ospfv2/spf.rs:390-461
//   intra_area_networks    holo-ospf/src/ospfv2/spf.rs:463-537
//   area_router_information / area_opaque_data_compile / route_prefix_sids
//                          holo-ospf/src/ospfv2/spf.rs:617-717
//   update_rib_intra_area  holo-ospf/src/route.rs:343-446
//   route_update/compare   holo-ospf/src/route.rs:895-971
//   prefix_sid_update ...  holo-ospf/src/sr.rs:29-77,127-255
// Ordered std::map / std::set stand in for BTreeMap / BTreeSet; the candidate
// list keeps the reference's linear lookup and the per-edge mutual-link
// re-iteration, so this file also has the reference's cost profile (it is the
// `cpu_baseline` "port" of bench.py).
//
// Parity pinning: checked against the reference's own golden topologies
// (tests/golden/ospfv2_*.json, extracted from
// holo-ospf/tests/conformance/ospfv2/topologies/*/rt*/output/northbound-state.json)
// in tests/test_oracle_golden.py.  SR prefix-SID labels are NOT pinned by any
// reference golden: "parity unpinned" for hl_route_net.sr_label /
// hl_nexthop.sr_label.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <optional>
#include <set>
#include <tuple>
#include <utility>
#include <vector>

#include "../include/holo_lsdb.h"
#include "../include/holo_spf.h"

namespace {

// VertexId: derived Ord => every Network < every Router, then by address
// (ospfv2/spf.rs:40-44).
struct VertexId {
    bool is_router;
    uint32_t addr;   // dr_addr or router_id
    bool operator<(const VertexId &o) const { return std::tie(is_router, addr) < std::tie(o.is_router, o.addr); }
    bool operator==(const VertexId &o) const { return is_router == o.is_router && addr == o.addr; }
};

// VertexLsa: index into router_lsas[] or network_lsas[]
struct VertexLsa { bool is_router; uint32_t idx; };

// NexthopKey { iface_idx, addr: Option } ordered by (arena index, addr) with
// None < Some (route.rs:92-98).
struct NexthopKey {
    uint32_t iface_sort; bool has_addr; uint32_t addr;
    bool operator<(const NexthopKey &o) const {
        return std::tie(iface_sort, has_addr, addr) < std::tie(o.iface_sort, o.has_addr, o.addr);
    }
};
struct Nexthop {
    uint32_t iface; bool has_addr; uint32_t addr; bool has_nbr; uint32_t nbr; bool has_label = false; uint32_t label = 0;
};
using Nexthops = std::map<NexthopKey, Nexthop>;

struct Vertex {
    VertexId id; VertexLsa lsa; uint16_t distance; uint16_t hops; Nexthops nexthops;
};

struct SpfLink {
    bool has_parent; uint32_t link_pos; const hl_ospfv2_link *parent_link;
    VertexId id; VertexLsa lsa; uint16_t cost;
};

struct Area {
    const hl_ospfv2_area *a;

    bool router_maxage(uint32_t i) const { return a->router_lsas[i].age == HL_LSA_MAX_AGE; }

    // vertex_lsa_find (ospfv2/spf.rs:356-388)
    std::optional<VertexLsa> vertex_lsa_find(VertexId id) const {
        if (!id.is_router) {
            // linear scan of all Network-LSAs in LsaKey order; first lsa_id match,
            // THEN the MaxAge filter
            for (uint32_t i = 0; i < a->n_network_lsas; ++i) {
                if (a->network_lsas[i].lsa_id == id.addr) {
                    if (a->network_lsas[i].age == HL_LSA_MAX_AGE) return std::nullopt;
                    return VertexLsa{false, i};
                }
            }
            return std::nullopt;
        }
        // lsdb.get(key = (Router, adv_rtr = router_id, lsa_id = router_id)); arrays are in key order
        uint32_t lo = 0, hi = a->n_router_lsas;
        auto key = std::make_pair(id.addr, id.addr);
        while (lo < hi) {
            uint32_t mid = (lo + hi) / 2;
            auto k = std::make_pair(a->router_lsas[mid].adv_rtr, a->router_lsas[mid].lsa_id);
            if (k < key) lo = mid + 1; else hi = mid;
        }
        if (lo < a->n_router_lsas && a->router_lsas[lo].adv_rtr == id.addr && a->router_lsas[lo].lsa_id == id.addr &&
            !router_maxage(lo))
            return VertexLsa{true, lo};
        return std::nullopt;
    }

    // vertex_lsa_links (ospfv2/spf.rs:390-461); f returns false to stop (Iterator::any)
    void vertex_lsa_links(const VertexLsa &vl, const std::function<bool(const SpfLink &)> &f) const {
        if (!vl.is_router) {
            const auto &n = a->network_lsas[vl.idx];
            for (uint32_t k = 0; k < n.n_att; ++k) {
                VertexId vid{true, a->attached[n.att_off + k]};
                auto l = vertex_lsa_find(vid);
                if (!l) continue;
                if (!f(SpfLink{false, 0, nullptr, vid, *l, 0})) return;
            }
            return;
        }
        const auto &r = a->router_lsas[vl.idx];
        uint32_t link_pos = 0;   // enumerate() AFTER the stub filter
        for (uint32_t k = 0; k < r.n_links; ++k) {
            const hl_ospfv2_link *link = &a->links[r.link_off + k];
            VertexId vid;
            if (link->link_type == HL_LINK_P2P || link->link_type == HL_LINK_VLINK) vid = VertexId{true, link->link_id};
            else if (link->link_type == HL_LINK_TRANSIT) vid = VertexId{false, link->link_id};
            else continue;
            uint32_t pos = link_pos++;
            auto l = vertex_lsa_find(vid);
            if (!l) continue;
            if (!f(SpfLink{true, pos, link, vid, *l, link->metric})) return;
        }
    }

    // Ospfv2::calc_nexthops (ospfv2/spf.rs:173-354); returns false on Err
    bool v2_calc_nexthops(const Vertex &parent, const SpfLink &plink, const VertexLsa &dest_lsa, Nexthops &out) const {
        if (parent.lsa.is_router) {
            // parent is the root
            uint32_t want = plink.link_pos, seen = 0;
            int iface_idx = -1;
            for (uint32_t i = 0; i < a->n_ifaces; ++i) {
                if (a->ifaces[i].n_nbrs == 0) continue;
                if (seen++ == want) { iface_idx = (int)i; break; }
            }
            if (iface_idx < 0) return false;
            const hl_ospf_iface &iface = a->ifaces[iface_idx];
            if (iface.if_type == HL_IF_VLINK) return true;   // empty set, Ok
            if (dest_lsa.is_router) {
                const auto &dl = a->router_lsas[dest_lsa.idx];
                if (iface.if_type == HL_IF_P2P || iface.if_type == HL_IF_VLINK) {
                    uint32_t nbr_router_id = dl.adv_rtr;
                    const hl_ospf_nbr *nbr = nullptr;
                    for (uint32_t k = 0; k < iface.n_nbrs; ++k)
                        if (a->nbrs[iface.nbr_off + k].router_id == nbr_router_id) { nbr = &a->nbrs[iface.nbr_off + k]; break; }
                    if (!nbr) return false;
                    out[NexthopKey{iface.sort_key, true, nbr->src}] = Nexthop{(uint32_t)iface_idx, true, nbr->src, true, nbr_router_id};
                } else if (iface.if_type == HL_IF_P2MP) {
                    for (uint32_t k = 0; k < dl.n_links; ++k) {
                        const auto &link = a->links[dl.link_off + k];
                        bool contains = false;
                        for (uint32_t q = 0; q < iface.n_addrs; ++q) {
                            const auto &net = a->iface_addrs[iface.addr_off + q];
                            if ((link.link_data & net.mask) == (net.addr & net.mask)) { contains = true; break; }
                        }
                        if (!contains) continue;
                        out[NexthopKey{iface.sort_key, true, link.link_data}] =
                            Nexthop{(uint32_t)iface_idx, true, link.link_data, true, dl.adv_rtr};
                    }
                }
                if (out.empty()) return false;
            } else {
                out[NexthopKey{iface.sort_key, false, 0}] = Nexthop{(uint32_t)iface_idx, false, 0, false, 0};
            }
            return true;
        }
        // parent is a network directly connecting the root to the destination router
        const auto &pl = a->network_lsas[parent.lsa.idx];
        const auto &dl = a->router_lsas[dest_lsa.idx];
        const hl_ospfv2_link *dest_link = nullptr;
        for (uint32_t k = 0; k < dl.n_links; ++k) {
            const auto &link = a->links[dl.link_off + k];
            if ((link.link_data & pl.mask) == (pl.lsa_id & pl.mask)) { dest_link = &link; break; }
        }
        if (!dest_link) return false;
        if (parent.nexthops.empty()) return false;
        uint32_t iface_idx = parent.nexthops.begin()->second.iface;
        out[NexthopKey{a->ifaces[iface_idx].sort_key, true, dest_link->link_data}] =
            Nexthop{iface_idx, true, dest_link->link_data, true, dl.adv_rtr};
        return true;
    }
};

struct RouterInfo { bool has_sr_algo = false; std::vector<const hl_srgb *> srgb; };

// area_router_information (ospfv2/spf.rs:617-654)
RouterInfo area_router_information(const hl_ospfv2_area *a, uint32_t router_id) {
    RouterInfo ri;
    for (uint32_t i = 0; i < a->n_ri_lsas; ++i) {
        const auto &l = a->ri_lsas[i];
        if (l.adv_rtr != router_id || l.age == HL_LSA_MAX_AGE) continue;
        if (l.has_sr_algo) ri.has_sr_algo = true;   // get_or_insert: first occurrence kept (only presence matters)
        for (uint32_t k = 0; k < l.n_srgb; ++k) ri.srgb.push_back(&a->srgbs[l.srgb_off + k]);
    }
    return ri;
}

// index_to_label (sr.rs:221-255)
std::optional<uint32_t> index_to_label(uint32_t index, const std::vector<const hl_srgb *> &srgbs) {
    for (auto *s : srgbs) {
        if (s->first_is_index) continue;
        if (index >= s->range) { index -= s->range; continue; }
        return s->first + index;
    }
    return std::nullopt;
}

struct PrefixSid { uint8_t flags; bool is_label; uint32_t value; };

struct RouteNet {
    uint32_t metric; uint8_t flags; uint8_t origin_type; uint32_t origin_adv_rtr, origin_lsa_id;
    std::optional<PrefixSid> prefix_sid; std::optional<uint32_t> sr_label; Nexthops nexthops;
};

using Prefix = std::pair<uint32_t, uint32_t>;   // (address, prefix length): Ipv4Network Ord

uint32_t mask_len(uint32_t mask) { return (uint32_t)__builtin_popcount(mask); }

}  // namespace

extern "C" int oracle_ospfv2_run_area(const hl_ospfv2_area *a, hl_ospfv2_result *out) {
    Area area{a};
    out->n_vertices = out->n_routers = out->n_routes = out->n_nexthops = 0;
    out->transit_capability = 0;
    out->root_found = 0;

    // area_opaque_data_compile (ospfv2/spf.rs:656-689)
    std::map<std::pair<uint32_t, Prefix>, const hl_ospfv2_ext_prefix *> ext_prefix_db;
    for (uint32_t i = 0; i < a->n_ext_prefixes; ++i) {
        const auto &e = a->ext_prefixes[i];
        if (e.age == HL_LSA_MAX_AGE) continue;
        ext_prefix_db.emplace(std::make_pair(e.adv_rtr, Prefix{e.prefix, mask_len(e.mask)}), &e);
    }

    // root
    VertexId root_vid{true, a->router_id};
    auto root_vlsa = area.vertex_lsa_find(root_vid);
    if (!root_vlsa) return 0;   // SpfRootNotFound
    out->root_found = 1;

    std::map<VertexId, Vertex> spt;
    std::map<std::pair<uint16_t, VertexId>, Vertex> cand_list;
    cand_list.emplace(std::make_pair((uint16_t)0, root_vid), Vertex{root_vid, *root_vlsa, 0, 0, {}});
    std::map<uint32_t, std::tuple<uint32_t, uint8_t, uint8_t, Nexthops>> routers;   // area.state.routers
    bool transit_capability = false;

    while (!cand_list.empty()) {
        auto first = cand_list.begin();
        Vertex popped = std::move(first->second);
        cand_list.erase(first);
        VertexId vertex_id = popped.id;
        spt[vertex_id] = std::move(popped);
        const Vertex &vertex = spt[vertex_id];

        if (vertex.lsa.is_router) {
            const auto &rl = a->router_lsas[vertex.lsa.idx];
            routers[rl.adv_rtr] = std::make_tuple((uint32_t)vertex.distance, rl.flags, rl.options, vertex.nexthops);
            if (rl.flags & HL_RTR_FLAG_V) transit_capability = true;
        }

        area.vertex_lsa_links(vertex.lsa, [&](const SpfLink &link) {
            // mutual link check (spf.rs:654-664)
            bool back = false;
            area.vertex_lsa_links(link.lsa, [&](const SpfLink &l2) {
                if (l2.id == vertex.id) { back = true; return false; }
                return true;
            });
            if (!back) return true;
            if (spt.count(link.id)) return true;
            uint32_t s = (uint32_t)vertex.distance + link.cost;
            uint16_t distance = s > 0xFFFF ? 0xFFFF : (uint16_t)s;            // saturating_add
            uint16_t hops = vertex.hops;
            if (link.lsa.is_router) hops = hops == 0xFFFF ? 0xFFFF : hops + 1;
            auto it = cand_list.begin();
            for (; it != cand_list.end(); ++it) if (it->second.id == link.id) break;   // linear scan
            if (it != cand_list.end()) {
                if (distance < it->second.distance) cand_list.erase(it);
                else if (distance > it->second.distance) return true;
            }
            auto key = std::make_pair(distance, link.id);
            auto ce = cand_list.find(key);
            if (ce == cand_list.end()) ce = cand_list.emplace(key, Vertex{link.id, link.lsa, distance, hops, {}}).first;
            Vertex &cand_v = ce->second;
            // calc_nexthops (spf.rs:733-767)
            if (vertex.hops == 0) {
                Nexthops nh;
                if (area.v2_calc_nexthops(vertex, link, cand_v.lsa, nh))
                    for (auto &kv : nh) cand_v.nexthops[kv.first] = kv.second;   // extend
            } else {
                for (auto &kv : vertex.nexthops) cand_v.nexthops[kv.first] = kv.second;
            }
            return true;
        });
    }

    // ---- update_rib_intra_area (route.rs:343-446) ----------------------------
    std::map<Prefix, RouteNet> rib;
    struct Stub { const Vertex *vertex; Prefix prefix; uint16_t metric; uint32_t adv_rtr; };
    std::vector<Stub> stubs;
    for (auto &kv : spt) {            // intra_area_networks (ospfv2/spf.rs:463-537): SPT order
        const Vertex &v = kv.second;
        if (!v.lsa.is_router) {
            const auto &nl = a->network_lsas[v.lsa.idx];
            stubs.push_back({&v, Prefix{nl.lsa_id & nl.mask, mask_len(nl.mask)}, 0, nl.adv_rtr});
        } else {
            const auto &rl = a->router_lsas[v.lsa.idx];
            for (uint32_t k = 0; k < rl.n_links; ++k) {
                const auto &link = a->links[rl.link_off + k];
                if (link.link_type != HL_LINK_STUB) continue;
                stubs.push_back({&v, Prefix{link.link_id & link.link_data, mask_len(link.link_data)}, link.metric, rl.adv_rtr});
            }
        }
    }
    RouterInfo local_ri;
    bool local_ri_loaded = false;
    for (const Stub &stub : stubs) {
        uint32_t s16 = (uint32_t)stub.vertex->distance + stub.metric;
        uint32_t metric = s16 > 0xFFFF ? 0xFFFF : s16;       // u16 saturating_add, then as u32
        auto cur = rib.find(stub.prefix);
        if (cur != rib.end() && metric > cur->second.metric) continue;
        uint8_t origin_type; uint32_t origin_adv, origin_id;
        if (stub.vertex->lsa.is_router) {
            const auto &rl = a->router_lsas[stub.vertex->lsa.idx];
            origin_type = 1; origin_adv = rl.adv_rtr; origin_id = rl.lsa_id;
        } else {
            const auto &nl = a->network_lsas[stub.vertex->lsa.idx];
            origin_type = 2; origin_adv = nl.adv_rtr; origin_id = nl.lsa_id;
        }
        if (!stub.vertex->lsa.is_router && cur != rib.end()) {
            if (metric > cur->second.metric || origin_id < cur->second.origin_lsa_id) continue;
            rib.erase(cur);
        }
        RouteNet nr{};
        nr.metric = metric;
        nr.flags = stub.vertex->hops == 0 ? HL_ROUTE_CONNECTED : 0;
        nr.origin_type = origin_type; nr.origin_adv_rtr = origin_adv; nr.origin_lsa_id = origin_id;
        nr.nexthops = stub.vertex->nexthops;

        // route_prefix_sids (ospfv2/spf.rs:694-717) + sr::prefix_sid_update (sr.rs:29-77)
        if (a->sr_enabled) {
            auto ep = ext_prefix_db.find(std::make_pair(stub.adv_rtr, stub.prefix));
            if (ep != ext_prefix_db.end() && ep->second->route_type == 1 /* IntraArea */ && ep->second->has_sid) {
                PrefixSid ps{ep->second->sid_flags, (bool)ep->second->sid_is_label, ep->second->sid_value};
                bool local = stub.vertex->hops == 0, last_hop = stub.vertex->hops == 1;
                RouterInfo ri = area_router_information(a, origin_adv);
                if (ri.has_sr_algo) {
                    nr.prefix_sid = ps;
                    // prefix_sid_input_label (sr.rs:127-167)
                    if (!(local && (!(ps.flags & HL_PSID_NP) || (ps.flags & HL_PSID_E)))) {
                        if (!ps.is_label) {
                            if (!local_ri_loaded) { local_ri = area_router_information(a, a->router_id); local_ri_loaded = true; }
                            if (!local_ri.srgb.empty()) {
                                auto l = index_to_label(ps.value, local_ri.srgb);
                                if (l) nr.sr_label = *l;
                            }
                        } else {
                            nr.sr_label = ps.value;
                        }
                    }
                    // prefix_sid_output_label per nexthop (sr.rs:170-219)
                    for (auto &kv : nr.nexthops) {
                        Nexthop &nh = kv.second;
                        if (!nh.has_nbr) continue;   // reference unwrap()s here; never reached on valid LSDBs
                        std::optional<uint32_t> lab;
                        bool decided = false;
                        if (last_hop) {
                            if (!(ps.flags & HL_PSID_NP)) { lab = 3; decided = true; }          // implicit null
                            else if (ps.flags & HL_PSID_E) { lab = 0; decided = true; }         // IPv4 explicit null
                        }
                        if (!decided) {
                            if (!ps.is_label) {
                                RouterInfo nri = area_router_information(a, nh.nbr);
                                if (!nri.srgb.empty()) lab = index_to_label(ps.value, nri.srgb);
                            } else {
                                lab = last_hop ? ps.value : 3u;
                            }
                        }
                        if (lab) { nh.has_label = true; nh.label = *lab; }
                    }
                }
            }
        }

        // route_update (route.rs:895-942); path types are all IntraArea here
        auto ex = rib.find(stub.prefix);
        RouteNet *route;
        if (ex != rib.end()) {
            RouteNet &cr = ex->second;
            if (nr.metric < cr.metric) cr = nr;
            else if (nr.metric == cr.metric) for (auto &kv : nr.nexthops) cr.nexthops[kv.first] = kv.second;
            route = &cr;
        } else {
            route = &rib.emplace(stub.prefix, nr).first->second;
        }
        if (route->nexthops.size() > a->max_paths) {
            Nexthops cut;
            uint32_t n = 0;
            for (auto &kv : route->nexthops) { if (n++ >= a->max_paths) break; cut.insert(kv); }
            route->nexthops = std::move(cut);
        }
    }

    // ---- export --------------------------------------------------------------------
    uint32_t need_v = (uint32_t)spt.size(), need_r = (uint32_t)routers.size(), need_n = (uint32_t)rib.size();
    uint32_t need_h = 0;
    for (auto &kv : spt) need_h += (uint32_t)kv.second.nexthops.size();
    for (auto &kv : routers) need_h += (uint32_t)std::get<3>(kv.second).size();
    for (auto &kv : rib) need_h += (uint32_t)kv.second.nexthops.size();
    bool fits = need_v <= out->vertices_cap && need_r <= out->routers_cap && need_n <= out->routes_cap &&
                need_h <= out->nexthops_cap;
    out->n_vertices = need_v; out->n_routers = need_r; out->n_routes = need_n; out->n_nexthops = need_h;
    out->transit_capability = transit_capability;
    if (!fits) return HSPF_E_NOMEM;
    uint32_t h = 0;
    auto put_nh = [&](const Nexthops &n) {
        for (auto &kv : n) {
            const Nexthop &x = kv.second;
            hl_nexthop o{};
            o.iface = x.iface; o.addr = x.has_addr ? x.addr : 0; o.nbr_router_id = x.has_nbr ? x.nbr : 0;
            o.sr_label = x.has_label ? x.label : 0;
            o.has_addr = x.has_addr; o.has_nbr = x.has_nbr; o.has_label = x.has_label;
            out->nexthops[h++] = o;
        }
    };
    uint32_t i = 0;
    for (auto &kv : spt) {
        hl_spt_vertex o{};
        o.id = kv.second.id.addr; o.distance = kv.second.distance; o.hops = kv.second.hops;
        o.is_router = kv.second.id.is_router; o.nh_off = h; o.n_nh = (uint32_t)kv.second.nexthops.size();
        put_nh(kv.second.nexthops);
        out->vertices[i++] = o;
    }
    i = 0;
    for (auto &kv : routers) {
        hl_route_rtr o{};
        o.router_id = kv.first; o.metric = std::get<0>(kv.second); o.flags = std::get<1>(kv.second);
        o.options = std::get<2>(kv.second); o.nh_off = h; o.n_nh = (uint32_t)std::get<3>(kv.second).size();
        put_nh(std::get<3>(kv.second));
        out->routers[i++] = o;
    }
    i = 0;
    for (auto &kv : rib) {
        const RouteNet &r = kv.second;
        hl_route_net o{};
        o.prefix = kv.first.first;
        o.mask = kv.first.second == 0 ? 0 : 0xFFFFFFFFu << (32 - kv.first.second);
        o.metric = r.metric; o.flags = r.flags; o.origin_type = r.origin_type;
        o.origin_adv_rtr = r.origin_adv_rtr; o.origin_lsa_id = r.origin_lsa_id;
        o.has_prefix_sid = r.prefix_sid.has_value();
        if (r.prefix_sid) { o.prefix_sid_value = r.prefix_sid->value; o.prefix_sid_flags = r.prefix_sid->flags; o.prefix_sid_is_label = r.prefix_sid->is_label; }
        o.has_sr_label = r.sr_label.has_value();
        o.sr_label = r.sr_label.value_or(0);
        o.nh_off = h; o.n_nh = (uint32_t)r.nexthops.size();
        put_nh(r.nexthops);
        out->routes[i++] = o;
    }
    return 0;
}

// Restates Ospfv2::spf_computation_type (holo-ospf/src/ospfv2/spf.rs:98-171): any Router-LSA, Network-LSA,
// area-scope Router-Information / Extended-Prefix / Extended-Link opaque LSA or AS-scope Extended-Prefix
// opaque LSA among the triggers asks for a full run; otherwise the run is partial over the prefixes of the
// changed type-3 and type-5 LSAs (Ipv4Network::with_netmask(lsa_id, mask), host bits kept) and the ASBR ids
// of the changed type-4 LSAs, each a BTreeSet.  PARITY UNPINNED: no reference test records this value.
#include <set>
extern "C" int oracle_ospfv2_spf_computation_type(const hl_lsa_trigger *tr, uint32_t n, hl_spf_computation *out) {
    bool full = false;
    std::set<std::pair<uint32_t, uint8_t>> inter_network, external;   // Ipv4Network Ord: address, then length
    std::set<uint32_t> inter_router;
    for (uint32_t i = 0; i < n; ++i) {
        switch (tr[i].lsa_type) {
        case 1: case 2: full = true; break;
        case 10: if (tr[i].opaque_type == 4 || tr[i].opaque_type == 7 || tr[i].opaque_type == 8) full = true; break;
        case 11: if (tr[i].opaque_type == 7) full = true; break;
        case 3: inter_network.insert({tr[i].lsa_id, (uint8_t)__builtin_popcount(tr[i].mask)}); break;
        case 4: inter_router.insert(tr[i].lsa_id); break;
        case 5: external.insert({tr[i].lsa_id, (uint8_t)__builtin_popcount(tr[i].mask)}); break;
        default: break;
        }
    }
    out->n_inter_network = out->n_inter_router = out->n_external = 0;
    if (full) { out->kind = HL_SPF_FULL; return 0; }
    out->kind = HL_SPF_PARTIAL;
    auto mask = [](uint8_t len) { return len ? 0xFFFFFFFFu << (32 - len) : 0u; };
    for (auto &p : inter_network) out->inter_network[out->n_inter_network++] = hl_ipv4_net{p.first, mask(p.second)};
    for (uint32_t r : inter_router) out->inter_router[out->n_inter_router++] = r;
    for (auto &p : external) out->external[out->n_external++] = hl_ipv4_net{p.first, mask(p.second)};
    return 0;
}
