// batch_pool.cc — TEST / BENCH INFRASTRUCTURE (oracle side): run many independent SPF jobs of
// the CPU restatement on a pool of native threads.  Used only by bench.py's cpu_baseline and
// `--impl reference` legs and by the tests that check whole batches against the oracle.
//
// One std::thread per worker, jobs handed out by an atomic counter, every worker with its own
// result buffers (no shared allocator traffic besides what the algorithm itself does), so the
// figure is the reference algorithm's own cost on N cores, not a Python thread pool's.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include <sched.h>

#include "holo_spf.h"

extern "C" {
int oracle_csr_spf(const hspf_csr *g, uint32_t root, uint32_t n_ov, const uint32_t *ov_edge, const uint32_t *ov_cost,
                   int vec_mode, uint32_t *dist, uint16_t *hops, uint32_t *first_parent, uint16_t *n_parents,
                   uint64_t *nh_mask, uint32_t nhw, uint32_t *poff, uint32_t *parents, uint32_t pcap, uint32_t *nvoff,
                   uint32_t *nhvec, uint32_t ncap, uint32_t *order, uint32_t *n_popped, uint32_t *status);
int oracle_csr_spf_heap(const hspf_csr *g, uint32_t root, uint32_t n_ov, const uint32_t *ov_edge,
                        const uint32_t *ov_cost, uint32_t *dist, uint16_t *hops, uint32_t *first_parent,
                        uint16_t *n_parents, uint64_t *nh_mask, uint32_t nhw, uint32_t *status);

/* Cores this process may really use: the affinity mask, capped by the cgroup CPU quota
 * (cgroup v2 cpu.max, then v1 cfs_quota/cfs_period). */
int oracle_usable_cores(void) {
    int n = 0;
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    if (n <= 0) n = 1;
    double quota = 0;
    {
        std::ifstream f("/sys/fs/cgroup/cpu.max");
        std::string a, b;
        if (f >> a >> b && a != "max") quota = std::stod(a) / std::stod(b);
    }
    if (quota <= 0) {
        std::ifstream fq("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"), fp("/sys/fs/cgroup/cpu/cpu.cfs_period_us");
        double q = 0, p = 0;
        if (fq >> q && fp >> p && q > 0 && p > 0) quota = q / p;
    }
    if (quota > 0 && quota < n) n = (int)(quota + 0.5) > 0 ? (int)(quota + 0.5) : 1;
    return n;
}

/* What the visible cores are worth: a fixed integer workload timed on one thread and then on
 * `threads` threads at once; returns threads * t1 / tN (a VM that time-slices 8 visible cores on
 * one physical core answers ~1, whatever the affinity mask or the cgroup files say). */
double oracle_effective_cores(int threads) {
    if (threads < 1) threads = 1;
    auto spin = [](uint64_t iters, uint64_t *out) {
        uint64_t x = 0x9E3779B97F4A7C15ull;
        for (uint64_t i = 0; i < iters; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; }
        *out = x;
    };
    auto timed = [&](int n, uint64_t iters) {
        std::vector<uint64_t> sink(n);
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for (int t = 1; t < n; ++t) pool.emplace_back(spin, iters, &sink[t]);
        spin(iters, &sink[0]);
        for (auto &th : pool) th.join();
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    uint64_t iters = 1u << 22;
    double t1 = timed(1, iters);
    while (t1 < 0.02 && iters < (1ull << 32)) { iters *= 2; t1 = timed(1, iters); }
    t1 = std::min(t1, timed(1, iters));
    const double tn = std::min(timed(threads, iters), timed(threads, iters));
    const double eff = threads * t1 / tn;
    return eff < 1 ? 1 : (eff > threads ? threads : eff);
}

/* Run jobs [0, n_jobs) (roots[j], overrides ov_edge/ov_cost[ov_off[j] .. ov_off[j+1]), ov_off may be
 * NULL) on `threads` workers.  mode 0: reference-faithful restatement (oracle_csr_spf), 1: binary-heap
 * Dijkstra (oracle_csr_spf_heap).  If stop_after_s > 0 the workers stop taking jobs once that much time
 * has passed (bounded sample); *jobs_done and *seconds report what ran.  Optional outputs
 * [n_jobs][V] (NULL to skip): dist, hops, first_parent, n_parents, nh_mask (nhw words), status[n_jobs].
 * checksum: xor-rotate over every job's distance plane (order independent), so a caller can compare
 * whole batches without keeping the planes. */
int oracle_csr_batch(const hspf_csr *g, uint32_t n_jobs, const uint32_t *roots, const uint32_t *ov_off,
                     const uint32_t *ov_edge, const uint32_t *ov_cost, int mode, int vec_mode, int threads,
                     double stop_after_s, uint32_t nhw, uint32_t *dist, uint16_t *hops, uint32_t *first_parent,
                     uint16_t *n_parents, uint64_t *nh_mask, uint32_t *status, uint32_t *jobs_done, double *seconds,
                     uint64_t *checksum) {
    if (!g || !roots || threads < 1 || nhw < 1) return -1;
    const uint32_t V = g->n_vertices;
    std::atomic<uint32_t> next{0}, done{0};
    std::atomic<uint64_t> sum{0};
    std::atomic<int> err{0};
    const auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    auto work = [&]() {
        std::vector<uint32_t> d(V), fp(V), order(mode == 0 ? V : 0);
        std::vector<uint16_t> hp(V), np(V);
        std::vector<uint64_t> nh((size_t)V * nhw);
        for (;;) {
            if (stop_after_s > 0 && elapsed() >= stop_after_s) break;
            const uint32_t j = next.fetch_add(1);
            if (j >= n_jobs) break;
            const uint32_t o0 = ov_off ? ov_off[j] : 0, o1 = ov_off ? ov_off[j + 1] : 0;
            uint32_t st = 0, popped = 0;
            int rc;
            if (mode == 0)
                rc = oracle_csr_spf(g, roots[j], o1 - o0, ov_edge ? ov_edge + o0 : nullptr, ov_cost ? ov_cost + o0 : nullptr,
                                    vec_mode, d.data(), hp.data(), fp.data(), np.data(), nh.data(), nhw, nullptr, nullptr, 0,
                                    nullptr, nullptr, 0, order.data(), &popped, &st);
            else
                rc = oracle_csr_spf_heap(g, roots[j], o1 - o0, ov_edge ? ov_edge + o0 : nullptr,
                                         ov_cost ? ov_cost + o0 : nullptr, d.data(), hp.data(), fp.data(), np.data(),
                                         nh.data(), nhw, &st);
            if (rc) err.store(rc);
            uint64_t h = 0x9E3779B97F4A7C15ull * (j + 1);
            for (uint32_t v = 0; v < V; ++v) h = ((h << 7) | (h >> 57)) ^ (d[v] + 0x632BE59BD9B4E019ull * (v + 1));
            sum.fetch_xor(h);
            const size_t o = (size_t)j * V;
            if (dist) std::memcpy(dist + o, d.data(), (size_t)V * 4);
            if (hops) std::memcpy(hops + o, hp.data(), (size_t)V * 2);
            if (first_parent) std::memcpy(first_parent + o, fp.data(), (size_t)V * 4);
            if (n_parents) std::memcpy(n_parents + o, np.data(), (size_t)V * 2);
            if (nh_mask) std::memcpy(nh_mask + o * nhw, nh.data(), (size_t)V * nhw * 8);
            if (status) status[j] = st;
            done.fetch_add(1);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    if (jobs_done) *jobs_done = done.load();
    if (seconds) *seconds = elapsed();
    if (checksum) *checksum = sum.load();
    return err.load();
}

}  // extern "C"
