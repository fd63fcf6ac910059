/* Host model of the CTA-level protocol of kernel MODE 2 (sim_kernels.cu: the SETDIRECT claim + setdirect_rendezvous):
 * W threads stand for the warps of one CTA, C groups of them for the CTAs of a launch.  Shared state per CTA: the batch
 * word (batch id << 24 | next unit), the table base, the table, arrivals, round, exhausted flag; global: the batch
 * cursor and the int64 tensor.  A unit of work adds `unit_leaves` leaves to key (set, unit % keys); a leaf goes to the
 * CTA's table when its set is the table's set, else to the tensor (as count_leaves_setdirect does).  A "warp" that
 * finds the batch used up parks in the rendezvous; the last one drains, fetches, re-bases, releases.  The program
 * checks that it terminates and that the tensor holds exactly the expected counts - i.e. that no add crosses a
 * re-base and no unit is lost or done twice.  Only the PROTOCOL is modelled (C++ atomics, sequentially consistent), not
 * the GPU memory model: the kernel adds __threadfence_block() where this model relies on seq_cst. */
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

struct Cta {
    std::atomic<uint64_t> batch{(0xFFFFFFFFFFull << 24) | 0x800000ull};
    std::atomic<uint32_t> base{0};
    std::atomic<int> arrived{0}, round{0}, exhausted{0};
    std::vector<std::atomic<uint32_t>> table;
    explicit Cta(size_t slots) : table(slots) { for (auto& t : table) t = 0; }
};

int main(int argc, char** argv)
{
    const int C = argc > 1 ? atoi(argv[1]) : 4, W = argc > 2 ? atoi(argv[2]) : 8;
    const uint32_t n_sets = argc > 3 ? atoi(argv[3]) : 37, units_per_set = argc > 4 ? atoi(argv[4]) : 23;
    const uint32_t batch_units = 7, slots = 16;
    const uint32_t batches_per_set = (units_per_set + batch_units - 1) / batch_units;
    const uint64_t total_batches = (uint64_t)batches_per_set * n_sets;
    std::atomic<uint64_t> cursor{0};
    std::vector<std::atomic<uint64_t>> tensor((size_t)n_sets * slots);
    for (auto& t : tensor) t = 0;
    std::vector<Cta*> ctas;
    for (int c = 0; c < C; ++c) ctas.push_back(new Cta(slots));
    std::atomic<long> units_done{0};
    std::vector<std::thread> th;
    for (int c = 0; c < C; ++c)
        for (int w = 0; w < W; ++w)
            th.emplace_back([&, c, w]() {
                Cta& s = *ctas[c];
                std::mt19937 rng(c * 1000 + w);
                for (;;) {
                    if (s.exhausted.load()) break;                        /* the kernel: idle_wait / termination */
                    const uint64_t old = s.batch.fetch_add(1);
                    const uint32_t off = (uint32_t)old & 0xFFFFFFu;
                    const uint64_t gb = old >> 24;
                    int status = 0;
                    uint32_t set = 0, j = 0;
                    if (gb < total_batches) {
                        set = (uint32_t)(gb / batches_per_set);
                        const uint32_t first_unit = (uint32_t)(gb - (uint64_t)set * batches_per_set) * batch_units;
                        const uint32_t left = units_per_set - first_unit;
                        if (off < (left < batch_units ? left : batch_units)) { j = first_unit + off; status = 1; }
                    }
                    if (status == 0) status = s.exhausted.load() ? 2 : 4;
                    if (status == 1) {                                     /* a unit: SEED + DIVIDE iterations */
                        const uint32_t key = set * slots + j % slots;
                        for (int leaf = 0; leaf < 5; ++leaf) {
                            const uint32_t rel = key - s.base.load();
                            if (rel < slots) s.table[rel].fetch_add(1); else tensor[key].fetch_add(1);
                            if ((rng() & 7) == 0) std::this_thread::yield();
                        }
                        units_done.fetch_add(1);
                        continue;
                    }
                    if (status == 2) break;
                    /* status 4: rendezvous (this model's warps never hold nodes across iterations) */
                    const int round = s.round.load();
                    if (s.arrived.fetch_add(1) + 1 == W) {
                        const uint32_t b = s.base.load();
                        for (uint32_t i = 0; i < slots; ++i) {
                            const uint32_t v = s.table[i].exchange(0);
                            if (v) tensor[b + i].fetch_add(v);
                        }
                        const uint64_t g = cursor.fetch_add(1);
                        if (g >= total_batches) s.exhausted.store(1);
                        else { s.base.store((uint32_t)(g / batches_per_set) * slots); s.batch.store(g << 24); }
                        s.arrived.store(0);
                        s.round.fetch_add(1);
                    } else {
                        while (s.round.load() == round) std::this_thread::yield();
                    }
                }
            });
    for (auto& t : th) t.join();
    /* kernel epilogue: every CTA flushes its table with its last base */
    for (int c = 0; c < C; ++c)
        for (uint32_t i = 0; i < slots; ++i) tensor[ctas[c]->base.load() + i].fetch_add(ctas[c]->table[i].load());
    long bad = 0;
    for (uint32_t set = 0; set < n_sets; ++set) {
        std::vector<uint64_t> want(slots, 0);
        for (uint32_t j = 0; j < units_per_set; ++j) want[j % slots] += 5;
        for (uint32_t i = 0; i < slots; ++i) bad += tensor[(size_t)set * slots + i].load() != want[i];
    }
    printf("units %ld of %u, wrong slots %ld\n", units_done.load(), n_sets * units_per_set, bad);
    return (bad == 0 && units_done.load() == (long)n_sets * units_per_set) ? 0 : 1;
}
