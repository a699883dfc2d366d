// TEST INFRASTRUCTURE — CPU restatement of the modified-MANET flooding-reduction queries of
// holo-isis (holo-isis/src/flooding/manet.rs:39-193) that consume the hop-count SPTs of the
// neighbour batch, with Spt::is_on_path / first_hops / second_hops (holo-isis/src/spf.rs:257-294).
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use anything under oracle/.
//
// Pinned by: the four known-answer vectors of the reference's own unit test for
// flood_reduction_hash (manet.rs:201-231, from draft-ietf-lsr-distoptflood-12 section 1.2.3),
// tests/test_isis_manet.py.  reflood_list itself has no golden in the reference: PARITY UNPINNED
// beyond this restatement.  Third-party: crate `fletcher` 1.0 (Cargo.toml:51), calc_fletcher16 =
// Fletcher-16 with modulus 255 over the bytes, result = sum2 << 8 | sum1.
#include <cstdint>
#include <map>
#include <set>
#include <vector>

#include "../include/holo_lsdb.h"
#include "../include/holo_spf.h"

namespace {

// Spt::is_on_path (spf.rs:259-284): stack-based DFS over all parent chains
bool is_on_path(const hl_isis_spt *spt, const std::map<hl_lan_id, uint32_t> &id_tree, uint64_t ancestor, uint64_t descendant) {
    auto a = id_tree.find((hl_lan_id)(ancestor << 8));     // VertexId::from(SystemId): pseudonode 0
    if (a == id_tree.end()) return false;
    auto s = id_tree.find((hl_lan_id)(descendant << 8));
    if (s == id_tree.end()) return false;
    std::vector<uint32_t> stack{s->second};
    std::vector<uint8_t> seen(spt->n_vertices, 0);         // same answer as the reference's unmarked DFS
    while (!stack.empty()) {
        const uint32_t cur = stack.back();
        stack.pop_back();
        if (cur == a->second) return true;
        if (seen[cur]) continue;
        seen[cur] = 1;
        const hl_isis_vertex &v = spt->vertices[cur];
        for (uint32_t k = 0; k < v.n_par; ++k) stack.push_back(spt->parents[v.par_off + k]);
    }
    return false;
}

}  // namespace

extern "C" uint16_t oracle_isis_flood_reduction_hash(uint64_t system_id, uint8_t pseudonode, uint8_t fragment) {
    // manet.rs:189-193: fragment >>= 3, then Fletcher-16 of LspId::to_bytes (packet/mod.rs:330-338)
    const uint8_t bytes[8] = {(uint8_t)(system_id >> 40), (uint8_t)(system_id >> 32), (uint8_t)(system_id >> 24),
                              (uint8_t)(system_id >> 16), (uint8_t)(system_id >> 8),  (uint8_t)system_id,
                              pseudonode, (uint8_t)(fragment >> 3)};
    uint32_t c0 = 0, c1 = 0;
    for (uint8_t b : bytes) { c0 = (c0 + b) % 255; c1 = (c1 + c0) % 255; }
    return (uint16_t)((c1 << 8) | c0);
}

// the RNL loop of init_cache (manet.rs:72-88)
extern "C" int oracle_isis_remote_neighbors(const hl_isis_level *l, const hl_isis_spt *spt, hl_isis_rnl_entry *out,
                                            uint32_t cap, uint32_t *n_out) {
    std::map<uint64_t, uint8_t> rnl;
    for (uint32_t k = 0; k < spt->n_first_hops; ++k) {
        const uint64_t sys = spt->vertices[spt->first_hops[k]].lan_id >> 8;
        // iter_for_system_id: every LSP of the system in LspId order; first valid one with the sub-TLV
        uint8_t algo = HL_ISIS_FLOOD_ZERO_PRUNER;
        std::map<std::pair<hl_lan_id, uint8_t>, const hl_isis_lsp *> ordered;
        for (uint32_t i = 0; i < l->n_lsps; ++i)
            if ((l->lsps[i].lan_id >> 8) == sys) ordered[{l->lsps[i].lan_id, l->lsps[i].fragment}] = &l->lsps[i];
        for (auto &kv : ordered) {
            const hl_isis_lsp &p = *kv.second;
            if (p.rem_lifetime == 0 || p.seqno == 0 || p.flood_algo == 0) continue;
            // FloodingAlgo::from_u8: unknown numbers give None -> unwrap_or(ZeroPruner)
            algo = (p.flood_algo == HL_ISIS_FLOOD_ZERO_PRUNER || p.flood_algo == HL_ISIS_FLOOD_MODIFIED_MANET)
                       ? p.flood_algo : (uint8_t)HL_ISIS_FLOOD_ZERO_PRUNER;
            break;
        }
        rnl[sys] = algo;
    }
    *n_out = (uint32_t)rnl.size();
    if (rnl.size() > cap) return HSPF_E_NOMEM;
    uint32_t i = 0;
    for (auto &kv : rnl) { out[i] = hl_isis_rnl_entry{}; out[i].system_id = kv.first; out[i].algo = kv.second; ++i; }
    return HSPF_OK;
}

// reflood_list (manet.rs:99-173)
extern "C" int oracle_isis_reflood_list(const hl_isis_spt *spt, const hl_isis_rnl_entry *rnl, uint32_t n_rnl,
                                        uint64_t local_system_id, uint64_t lsp_system_id, uint8_t lsp_pseudonode,
                                        uint8_t lsp_fragment, uint64_t *out, uint32_t cap, uint32_t *n_out) {
    *n_out = 0;
    if (n_rnl == 0) return HSPF_OK;                       // cache.remote_nbr_list.is_empty()
    std::map<hl_lan_id, uint32_t> id_tree;
    for (uint32_t i = 0; i < spt->n_vertices; ++i) id_tree[spt->vertices[i].lan_id] = i;
    // Two-Hop List
    std::set<uint64_t> thl;
    for (uint32_t k = 0; k < spt->n_second_hops; ++k) {
        const uint64_t sys = spt->vertices[spt->second_hops[k]].lan_id >> 8;
        if (sys == lsp_system_id) continue;                                   // skip LSP originator
        if (is_on_path(spt, id_tree, sys, lsp_system_id)) continue;           // on the path TN -> originator
        thl.insert(sys);
    }
    const uint16_t h = oracle_isis_flood_reduction_hash(lsp_system_id, lsp_pseudonode, lsp_fragment);
    const uint32_t n = (uint32_t)h % n_rnl;
    std::set<uint64_t> reflood;
    for (uint32_t step = 0; step < n_rnl; ++step) {                           // cycle().skip(n).take(rnum)
        const hl_isis_rnl_entry &e = rnl[(n + step) % n_rnl];
        if (thl.empty()) break;
        if (e.system_id == local_system_id) {
            for (uint64_t t : thl)
                if (is_on_path(spt, id_tree, e.system_id, t)) reflood.insert(t);
            break;
        }
        if (e.algo != HL_ISIS_FLOOD_MODIFIED_MANET) continue;
        for (auto it = thl.begin(); it != thl.end();) {
            if (is_on_path(spt, id_tree, e.system_id, *it)) it = thl.erase(it);
            else ++it;
        }
    }
    *n_out = (uint32_t)reflood.size();
    if (reflood.size() > cap) return HSPF_E_NOMEM;
    uint32_t i = 0;
    for (uint64_t t : reflood) out[i++] = t;
    return HSPF_OK;
}
