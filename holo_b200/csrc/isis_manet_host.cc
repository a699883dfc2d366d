// IS-IS flooding-reduction queries over the hop-count SPTs of the neighbour batch:
// hspf_isis_flood_reduction_hash / hspf_isis_remote_neighbors / hspf_isis_reflood_list
// (include/holo_spf_lsdb.h).  They replace, for the caller of the a16 batch,
//   the Remote Neighbor List loop of manet::init_cache   holo-isis/src/flooding/manet.rs:72-88
//   reflood_list                                         manet.rs:99-173
//   Spt::is_on_path / first_hops / second_hops           holo-isis/src/spf.rs:257-294
//   flood_reduction_hash                                 manet.rs:189-193
// is_on_path(a, d) asks whether a is an ancestor-or-self of d over ALL parent links of the SPT;
// reflood_list asks it for many pairs with few distinct descendants, so the ancestor set of a
// descendant is computed once (one upward sweep) and kept as a bit vector.
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

struct Ancestry {
    const hl_isis_spt *spt;
    std::unordered_map<uint64_t, uint32_t> router_vertex;            // system id -> vertex (pseudonode 0)
    std::unordered_map<uint32_t, std::vector<uint64_t>> memo;        // descendant vertex -> ancestor bits
    explicit Ancestry(const hl_isis_spt *s) : spt(s) {
        router_vertex.reserve(s->n_vertices);
        for (uint32_t i = 0; i < s->n_vertices; ++i)
            if ((s->vertices[i].lan_id & 0xFF) == 0) router_vertex.emplace(s->vertices[i].lan_id >> 8, i);
    }
    const std::vector<uint64_t> &ancestors(uint32_t d) {
        auto it = memo.find(d);
        if (it != memo.end()) return it->second;
        std::vector<uint64_t> bits((spt->n_vertices + 63) / 64, 0);
        std::vector<uint32_t> todo{d};
        bits[d >> 6] |= 1ull << (d & 63);
        while (!todo.empty()) {
            const uint32_t cur = todo.back();
            todo.pop_back();
            const hl_isis_vertex &v = spt->vertices[cur];
            for (uint32_t k = 0; k < v.n_par; ++k) {
                const uint32_t p = spt->parents[v.par_off + k];
                if (p >= spt->n_vertices || (bits[p >> 6] >> (p & 63)) & 1) continue;
                bits[p >> 6] |= 1ull << (p & 63);
                todo.push_back(p);
            }
        }
        return memo.emplace(d, std::move(bits)).first->second;
    }
    bool on_path(uint64_t ancestor_sys, uint64_t descendant_sys) {
        auto a = router_vertex.find(ancestor_sys);
        auto d = router_vertex.find(descendant_sys);
        if (a == router_vertex.end() || d == router_vertex.end()) return false;
        const auto &bits = ancestors(d->second);
        return (bits[a->second >> 6] >> (a->second & 63)) & 1;
    }
};

bool spt_ok(const hl_isis_spt *s) {
    if (!s) return false;
    if ((s->n_vertices && !s->vertices) || (s->n_parents && !s->parents)) return false;
    if ((s->n_first_hops && !s->first_hops) || (s->n_second_hops && !s->second_hops)) return false;
    for (uint32_t i = 0; i < s->n_vertices; ++i)
        if ((uint64_t)s->vertices[i].par_off + s->vertices[i].n_par > s->n_parents) return false;
    for (uint32_t i = 0; i < s->n_first_hops; ++i) if (s->first_hops[i] >= s->n_vertices) return false;
    for (uint32_t i = 0; i < s->n_second_hops; ++i) if (s->second_hops[i] >= s->n_vertices) return false;
    return true;
}

}  // namespace

extern "C" uint16_t hspf_isis_flood_reduction_hash(uint64_t system_id, uint8_t pseudonode, uint8_t fragment) {
    // Fletcher-16 (modulus 255) of the 8 LSP-id bytes, fragment shifted right by 3
    uint32_t s1 = 0, s2 = 0;
    auto feed = [&](uint8_t b) { s1 += b; if (s1 >= 255) s1 -= 255; s2 += s1; if (s2 >= 255) s2 -= 255; };
    for (int shift = 40; shift >= 0; shift -= 8) feed((uint8_t)(system_id >> shift));
    feed(pseudonode);
    feed((uint8_t)(fragment >> 3));
    return (uint16_t)((s2 << 8) | s1);
}

extern "C" int hspf_isis_remote_neighbors(const hl_isis_level *lvl, const hl_isis_spt *spt, hl_isis_rnl_entry *out,
                                          uint32_t cap, uint32_t *n_out) {
    if (!lvl || !n_out || !spt_ok(spt) || (cap && !out) || (lvl->n_lsps && !lvl->lsps)) return HSPF_E_INVAL;
    try {
        // per system: the flooding algorithm of its first valid LSP (LspId order) that carries the sub-TLV
        struct Best { uint64_t key; uint8_t algo; };     // key = lan_id << 8 | fragment
        std::unordered_map<uint64_t, Best> adv;
        for (uint32_t i = 0; i < lvl->n_lsps; ++i) {
            const hl_isis_lsp &p = lvl->lsps[i];
            if (!p.seqno || !p.rem_lifetime || !p.flood_algo) continue;
            const uint64_t key = ((uint64_t)(p.lan_id & 0xFF) << 8) | p.fragment;     // order inside one system
            auto it = adv.find(p.lan_id >> 8);
            if (it == adv.end()) adv.emplace(p.lan_id >> 8, Best{key, p.flood_algo});
            else if (key < it->second.key) it->second = Best{key, p.flood_algo};
        }
        std::vector<hl_isis_rnl_entry> rnl;
        for (uint32_t k = 0; k < spt->n_first_hops; ++k) {
            hl_isis_rnl_entry e;
            std::memset(&e, 0, sizeof(e));
            e.system_id = spt->vertices[spt->first_hops[k]].lan_id >> 8;
            auto it = adv.find(e.system_id);
            const uint8_t a = it == adv.end() ? 0 : it->second.algo;
            e.algo = (a == HL_ISIS_FLOOD_MODIFIED_MANET) ? a : (uint8_t)HL_ISIS_FLOOD_ZERO_PRUNER;   // unknown -> default
            rnl.push_back(e);
        }
        std::sort(rnl.begin(), rnl.end(), [](const hl_isis_rnl_entry &x, const hl_isis_rnl_entry &y) { return x.system_id < y.system_id; });
        rnl.erase(std::unique(rnl.begin(), rnl.end(), [](const hl_isis_rnl_entry &x, const hl_isis_rnl_entry &y) { return x.system_id == y.system_id; }),
                  rnl.end());
        *n_out = (uint32_t)rnl.size();
        if (rnl.size() > cap) return HSPF_E_NOMEM;
        for (size_t i = 0; i < rnl.size(); ++i) out[i] = rnl[i];
        return HSPF_OK;
    } catch (const std::bad_alloc &) { return HSPF_E_NOMEM; } catch (...) { return HSPF_E_INVAL; }
}

extern "C" int hspf_isis_reflood_list(const hl_isis_spt *spt, const hl_isis_rnl_entry *rnl, uint32_t n_rnl,
                                      uint64_t local_system_id, uint64_t lsp_system_id, uint8_t lsp_pseudonode,
                                      uint8_t lsp_fragment, uint64_t *out, uint32_t cap, uint32_t *n_out) {
    if (!n_out || !spt_ok(spt) || (n_rnl && !rnl) || (cap && !out)) return HSPF_E_INVAL;
    try {
        *n_out = 0;
        if (n_rnl == 0) return HSPF_OK;
        Ancestry anc(spt);
        // two-hop list: second hops that are neither the originator nor on a shortest path to it
        std::vector<uint64_t> thl;
        for (uint32_t k = 0; k < spt->n_second_hops; ++k) {
            const uint64_t sys = spt->vertices[spt->second_hops[k]].lan_id >> 8;
            if (sys == lsp_system_id || anc.on_path(sys, lsp_system_id)) continue;
            thl.push_back(sys);
        }
        std::sort(thl.begin(), thl.end());
        thl.erase(std::unique(thl.begin(), thl.end()), thl.end());
        const uint32_t start = (uint32_t)hspf_isis_flood_reduction_hash(lsp_system_id, lsp_pseudonode, lsp_fragment) % n_rnl;
        std::vector<uint64_t> reflood;
        for (uint32_t step = 0; step < n_rnl && !thl.empty(); ++step) {
            const hl_isis_rnl_entry &e = rnl[(start + step) % n_rnl];
            if (e.system_id == local_system_id) {
                for (uint64_t t : thl)
                    if (anc.on_path(e.system_id, t)) reflood.push_back(t);
                break;
            }
            if (e.algo != HL_ISIS_FLOOD_MODIFIED_MANET) continue;
            thl.erase(std::remove_if(thl.begin(), thl.end(), [&](uint64_t t) { return anc.on_path(e.system_id, t); }), thl.end());
        }
        *n_out = (uint32_t)reflood.size();                 // already ascending: thl is sorted
        if (reflood.size() > cap) return HSPF_E_NOMEM;
        for (size_t i = 0; i < reflood.size(); ++i) out[i] = reflood[i];
        return HSPF_OK;
    } catch (const std::bad_alloc &) { return HSPF_E_NOMEM; } catch (...) { return HSPF_E_INVAL; }
}
