// quad_layout.h — host-side construction of the "quad space" graph image used by
// spf_quad_kernel (spf_quad.cuh).
//
// The adjacency of every vertex is cut into 16-byte QUADS of four packed edge records
// (slot | cost << 16); a vertex with more than four links owns a CHAIN of consecutive
// quads.  Every quad is one uniform work item of the device SSSP: a lane loads one quad
// with a single 128-bit load and relaxes its four records, so the frontier expansion
// has no per-vertex degree loop, no row-offset lookups and no idle team lanes.
//
//  * forward quads (fq): records point at the SLOT of the head vertex = index of the
//    head's first forward quad.  The tentative distances of the SSSP are indexed by
//    slot (dist[NQ]; the entries of continuation quads are unused): a queue entry of the
//    SSSP names a quad and the first quad of its chain, whose distance it relaxes from.
//  * in-quads (iq): the transposed adjacency in the same form (records point at the
//    slot of the source vertex); one thread per in-quad evaluates the ECMP-DAG
//    predicate, partial results of a chain are combined with warp shuffles.
//  * chains never straddle a 32-quad boundary (dummy quads pad the gap), so a chain
//    lives in one word of the frontier bitmaps / one warp of the parents pass.
//
// A pad record of a forward quad of vertex v is (slot(v) | 0xFFFF << 16): relaxing the
// owner's own slot with cost 65535 can never improve it, so the relaxation body needs no
// validity branch (dummy quads pad with their own index and are never expanded).  A pad record of an in-quad of vertex v is (slot(v) | 0xFFFF << 16):
// dist[v] + 65535 == dist[v] never holds, so it is never a DAG edge.
//
// Reference semantics are unchanged: vertex order (= slot order) is the VertexId order
// of the reference's candidate list, holo-ospf/src/spf.rs:681-685.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace hspf {

struct QuadHost {
    bool eligible = false;
    const char *why = "";         // reason when not eligible
    uint32_t NQ = 0, NIQ = 0;     // forward / in quads incl. dummies, multiples of 32
    uint32_t shift = 0;           // log2 of the SSSP bucket width
    uint32_t max_ichain = 1;      // longest in-quad chain (quads)
    uint32_t max_atoms = 0;       // largest first-hop atom count of any root
    std::vector<uint32_t> fq;     // [NQ*4]
    std::vector<uint32_t> fcont;  // [NQ/32] bit q%32: quad q continues the chain of quad q-1
    std::vector<uint16_t> slot_of;// [V]
    std::vector<uint16_t> vert_of;// [NQ] owner vertex (0xFFFF: dummy)
    std::vector<uint32_t> iq;     // [NIQ*4]
    std::vector<uint32_t> imeta;  // [NIQ*2] x = slot(owner) | owner << 16 (0xFFFFFFFF dummy), y = rem | pos << 8
    std::vector<uint32_t> fpos;   // [E] forward edge -> fq quad * 4 + record
    std::vector<uint32_t> ipos;   // [E] forward edge -> iq quad * 4 + record
};

// irow/isrc/icost/ifwd: transposed CSR (in-edges of v ordered by (source, forward edge index)).
inline QuadHost build_quads(uint32_t V, uint32_t E, const uint32_t *row, const uint32_t *col, const uint32_t *cost,
                            const uint8_t *vflags, const uint32_t *irow, const uint32_t *isrc, const uint32_t *icost,
                            const uint32_t *ifwd, uint32_t delta_hint) {
    QuadHost Q;
    if (V >= 0xFFFFu) { Q.why = "more than 65534 vertices"; return Q; }
    uint32_t max_cost = 0;
    uint64_t cost_sum = 0;
    for (uint32_t e = 0; e < E; ++e) { max_cost = std::max(max_cost, cost[e]); cost_sum += cost[e]; }
    if (max_cost > 0xFFFEu) { Q.why = "link cost above 65534"; return Q; }
    auto quads_of = [](uint32_t deg) { return deg ? (deg + 3) / 4 : 1u; };
    // ---- forward quads --------------------------------------------------------------
    std::vector<uint32_t> slot(V);
    uint32_t pos = 0;
    for (uint32_t v = 0; v < V; ++v) {
        const uint32_t nq = quads_of(row[v + 1] - row[v]);
        if (nq > 32) { Q.why = "out-degree above 128"; return Q; }
        if ((pos & 31u) + nq > 32) pos = (pos + 31u) & ~31u;
        slot[v] = pos;
        pos += nq;
    }
    const uint32_t NQ = (pos + 31u) & ~31u;
    if (NQ >= 0xFFFFu) { Q.why = "more than 65534 forward quads"; return Q; }
    Q.NQ = NQ;
    Q.fq.resize((size_t)NQ * 4);
    Q.fcont.assign(NQ / 32, 0);
    Q.vert_of.assign(NQ, 0xFFFFu);
    Q.slot_of.resize(V);
    Q.fpos.resize(E);
    for (uint32_t q = 0; q < NQ; ++q)
        for (int k = 0; k < 4; ++k) Q.fq[(size_t)q * 4 + k] = q | 0xFFFF0000u;
    for (uint32_t v = 0; v < V; ++v) {
        Q.slot_of[v] = (uint16_t)slot[v];
        const uint32_t deg = row[v + 1] - row[v], nq = quads_of(deg);
        for (uint32_t j = 0; j < nq; ++j) {
            Q.vert_of[slot[v] + j] = (uint16_t)v;
            if (j) Q.fcont[(slot[v] + j) >> 5] |= 1u << ((slot[v] + j) & 31);
            for (int k = 0; k < 4; ++k) Q.fq[(size_t)(slot[v] + j) * 4 + k] = slot[v] | 0xFFFF0000u;   // pad: owner's slot
        }
        for (uint32_t i = 0; i < deg; ++i) {
            const uint32_t e = row[v] + i;
            Q.fq[(size_t)slot[v] * 4 + i] = slot[col[e]] | (cost[e] << 16);
            Q.fpos[e] = slot[v] * 4 + i;
        }
    }
    // ---- in-quads ---------------------------------------------------------------------
    std::vector<uint32_t> islot(V);
    pos = 0;
    for (uint32_t v = 0; v < V; ++v) {
        const uint32_t nq = quads_of(irow[v + 1] - irow[v]);
        if (nq > 32) { Q.why = "in-degree above 128"; return Q; }
        if ((pos & 31u) + nq > 32) pos = (pos + 31u) & ~31u;
        islot[v] = pos;
        pos += nq;
        Q.max_ichain = std::max(Q.max_ichain, nq);
    }
    const uint32_t NIQ = (pos + 31u) & ~31u;
    Q.NIQ = NIQ;
    Q.iq.assign((size_t)NIQ * 4, 0u | 0xFFFF0000u);
    Q.imeta.assign((size_t)NIQ * 2, 0xFFFFFFFFu);
    for (uint32_t q = 0; q < NIQ; ++q) Q.imeta[(size_t)q * 2 + 1] = 0;     // dummy: no chain
    Q.ipos.resize(E);
    for (uint32_t v = 0; v < V; ++v) {
        const uint32_t deg = irow[v + 1] - irow[v], nq = quads_of(deg);
        for (uint32_t j = 0; j < nq; ++j) {
            const uint32_t q = islot[v] + j;
            Q.imeta[(size_t)q * 2] = slot[v] | (v << 16);
            Q.imeta[(size_t)q * 2 + 1] = (nq - 1 - j) | (j << 8);
            for (int k = 0; k < 4; ++k) Q.iq[(size_t)q * 4 + k] = slot[v] | 0xFFFF0000u;
        }
        for (uint32_t i = 0; i < deg; ++i) {
            const uint32_t k = irow[v] + i;
            Q.iq[(size_t)islot[v] * 4 + i] = slot[isrc[k]] | (icost[k] << 16);
            Q.ipos[ifwd[k]] = islot[v] * 4 + i;
        }
    }
    // ---- bucket width: a power of two near 2.5x the mean cost (measured on the BASELINE
    // shapes: ~1.4x re-expansion, a few buckets), and at least a third of the largest cost
    // so that a relaxation out of bucket b lands in b .. b+3, the span of the ring of four
    // frontier bitmaps ------------------------------------------------------------------
    uint64_t want = delta_hint ? delta_hint : (E ? (5 * (cost_sum / E) + 1) / 2 : 1);
    want = std::max<uint64_t>(want, (uint64_t)(max_cost + 2) / 3);
    uint32_t sh = 0;
    while ((1ull << sh) < want) ++sh;
    Q.shift = sh;
    // ---- largest first-hop atom count (root edges + edges of its non-HOP heads) --------
    for (uint32_t r = 0; r < V; ++r) {
        uint32_t n = row[r + 1] - row[r];
        for (uint32_t e = row[r]; e < row[r + 1]; ++e)
            if (!(vflags[col[e]] & 1u)) n += row[col[e] + 1] - row[col[e]];
        Q.max_atoms = std::max(Q.max_atoms, n);
    }
    Q.eligible = true;
    return Q;
}

}  // namespace hspf
