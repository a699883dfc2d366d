// TEST HARNESS (not part of libholo_spf.so): runs route_cell_eval — the body of the device route
// kernel, holo_b200/csrc/route_cells.h — on the CPU over planes the test supplies, so that the cell
// walk and the host decode can be checked against the oracle without a GPU.
#include <cstdint>

#include "../../holo_b200/csrc/route_cells.h"

extern "C" int harness_route_cells(const hspf_ospfv2_rtable *rt, uint32_t n_jobs, const uint32_t *dist,
                                   const uint16_t *hops, const uint64_t *nh, hl_route_cell *cells) {
    const auto &t = rt->t;
    const uint32_t P = (uint32_t)t.prefix.size(), V = t.n_vertices;
    for (uint32_t j = 0; j < n_jobs; ++j) {
        const hspf::PlanesWide pl{dist + (size_t)j * V, hops + (size_t)j * V, nh + (size_t)j * V};
        for (uint32_t p = 0; p < P; ++p)
            cells[(size_t)j * P + p] = hspf::route_cell_eval(pl, t.contribs.data(), t.off[p], t.off[p + 1]);
    }
    return 0;
}

extern "C" int harness_route_cells16(const hspf_ospfv2_rtable *rt, uint32_t n_jobs, const uint16_t *dist,
                                     const uint16_t *hops, const uint16_t *nh, hl_route_cell *cells) {
    const auto &t = rt->t;
    const uint32_t P = (uint32_t)t.prefix.size(), V = t.n_vertices;
    for (uint32_t j = 0; j < n_jobs; ++j) {
        const hspf::PlanesNarrow pl{dist + (size_t)j * V, hops + (size_t)j * V, nh + (size_t)j * V};
        for (uint32_t p = 0; p < P; ++p)
            cells[(size_t)j * P + p] = hspf::route_cell_eval(pl, t.contribs.data(), t.off[p], t.off[p + 1]);
    }
    return 0;
}
