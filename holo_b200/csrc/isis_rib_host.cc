// The end of holo-isis' update_rib: hspf_isis_rib_merge / hspf_isis_rib_diff
// (include/holo_spf_lsdb.h) — merge of the per-level tables with L1 preferred
// (holo-isis/src/route.rs:236-242) and update_global_rib (route.rs:255-314).  Both inputs are in
// prefix order (IPv4 before IPv6, address, then length), so both are single merge walks.
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/holo_lsdb.h"
#include "../../include/holo_spf.h"
#include "../../include/holo_spf_lsdb.h"

namespace {

int cmp_prefix(const hl_isis_route &a, const hl_isis_route &b) {
    if (a.prefix.is_v6 != b.prefix.is_v6) return a.prefix.is_v6 < b.prefix.is_v6 ? -1 : 1;
    const int c = std::memcmp(a.prefix.bytes, b.prefix.bytes, 16);
    if (c) return c < 0 ? -1 : 1;
    return a.len == b.len ? 0 : (a.len < b.len ? -1 : 1);
}

bool rib_ok(const hl_isis_rib *r) {
    if (!r) return true;
    if ((r->n_routes && !r->routes) || (r->n_nexthops && !r->nexthops)) return false;
    for (uint32_t i = 0; i < r->n_routes; ++i)
        if ((uint64_t)r->routes[i].nh_off + r->routes[i].n_nh > r->n_nexthops) return false;
    return true;
}

bool same_hop(const hl_isis_nexthop &a, const hl_isis_nexthop &b) {   // Nexthop == (route.rs:50-61)
    return a.system_id == b.system_id && a.iface == b.iface && a.addr.is_v6 == b.addr.is_v6 &&
           std::memcmp(a.addr.bytes, b.addr.bytes, 16) == 0 && a.has_label == b.has_label &&
           (!a.has_label || a.sr_label == b.sr_label);
}

}  // namespace

extern "C" int hspf_isis_rib_merge(const hl_isis_rib *l2, const hl_isis_rib *l1, hl_isis_rib *out) {
    if (!out || !rib_ok(l2) || !rib_ok(l1)) return HSPF_E_INVAL;
    try {
        struct Pick { const hl_isis_rib *src; uint32_t idx; };
        std::vector<Pick> picks;
        const uint32_t n2 = l2 ? l2->n_routes : 0, n1 = l1 ? l1->n_routes : 0;
        uint32_t i = 0, j = 0, n_h = 0;
        while (i < n2 || j < n1) {
            int c = i == n2 ? 1 : (j == n1 ? -1 : cmp_prefix(l2->routes[i], l1->routes[j]));
            if (c < 0) picks.push_back(Pick{l2, i++});
            else { picks.push_back(Pick{l1, j++}); if (c == 0) ++i; }   // same prefix: the L1 route wins
        }
        for (const Pick &p : picks) n_h += p.src->routes[p.idx].n_nh;
        out->n_routes = (uint32_t)picks.size();
        out->n_nexthops = n_h;
        if (out->n_routes > out->routes_cap || n_h > out->nexthops_cap) return HSPF_E_NOMEM;
        if ((out->n_routes && !out->routes) || (n_h && !out->nexthops)) return HSPF_E_INVAL;
        uint32_t h = 0;
        for (uint32_t r = 0; r < picks.size(); ++r) {
            hl_isis_route o = picks[r].src->routes[picks[r].idx];
            const hl_isis_nexthop *from = picks[r].src->nexthops + o.nh_off;
            o.nh_off = h;
            for (uint32_t k = 0; k < o.n_nh; ++k) out->nexthops[h++] = from[k];
            out->routes[r] = o;
        }
        return HSPF_OK;
    } catch (const std::bad_alloc &) { return HSPF_E_NOMEM; } catch (...) { return HSPF_E_INVAL; }
}

extern "C" int hspf_isis_rib_diff(const hl_isis_rib *old_rib, hl_isis_rib *new_rib, hl_rib_action *out, uint32_t cap,
                                  uint32_t *n_out) {
    if (!new_rib || !n_out || (cap && !out) || !rib_ok(new_rib) || !rib_ok(old_rib)) return HSPF_E_INVAL;
    try {
        const uint32_t n_old = old_rib ? old_rib->n_routes : 0;
        std::vector<hl_rib_action> acts;
        auto push = [&](uint8_t kind, uint32_t route, const hl_isis_route *replaced) {
            hl_rib_action a;
            std::memset(&a, 0, sizeof(a));
            a.kind = kind; a.route = route;
            if (replaced) { a.has_old_sr_label = replaced->has_sr_label; a.old_sr_label = replaced->has_sr_label ? replaced->sr_label : 0; }
            acts.push_back(a);
        };
        std::vector<uint32_t> gone;
        uint32_t io = 0;
        for (uint32_t in = 0; in < new_rib->n_routes; ++in) {
            hl_isis_route &r = new_rib->routes[in];
            while (io < n_old && cmp_prefix(old_rib->routes[io], r) < 0) {
                if (old_rib->routes[io].flags & HL_ROUTE_INSTALLED) gone.push_back(io);
                ++io;
            }
            const hl_isis_route *o = (io < n_old && cmp_prefix(old_rib->routes[io], r) == 0) ? &old_rib->routes[io++] : nullptr;
            if (o) {
                bool same = o->metric == r.metric && o->n_nh == r.n_nh;      // tag: always None here
                for (uint32_t k = 0; same && k < r.n_nh; ++k)
                    same = same_hop(old_rib->nexthops[o->nh_off + k], new_rib->nexthops[r.nh_off + k]);
                if (same) {
                    if (o->flags & HL_ROUTE_INSTALLED) r.flags |= HL_ROUTE_INSTALLED;
                    continue;
                }
            }
            if (!(r.flags & HL_ROUTE_CONNECTED) && r.n_nh != 0) {
                push(HL_RIB_INSTALL, in, o);
                r.flags |= HL_ROUTE_INSTALLED;
            } else if (r.flags & HL_ROUTE_INSTALLED) {
                push(HL_RIB_UNINSTALL, in, nullptr);
                r.flags &= (uint8_t)~HL_ROUTE_INSTALLED;
            }
        }
        for (; io < n_old; ++io)
            if (old_rib->routes[io].flags & HL_ROUTE_INSTALLED) gone.push_back(io);
        for (uint32_t i : gone) push(HL_RIB_UNINSTALL_OLD, i, nullptr);
        *n_out = (uint32_t)acts.size();
        if (acts.size() > cap) return HSPF_E_NOMEM;
        for (size_t i = 0; i < acts.size(); ++i) out[i] = acts[i];
        return HSPF_OK;
    } catch (const std::bad_alloc &) { return HSPF_E_NOMEM; } catch (...) { return HSPF_E_INVAL; }
}
