/* capi.cu - the thin host layer above the sm_100a kernels: resident engine + one-shot proliferate.
 *
 * Replaces the reference's host drivers simulation::create_cells_population
 * (src/simulation/cells_population.cu:18-54) and simulation::proliferate / run_iteration
 * (src/simulation/proliferation.cu:26-284): no per-level cudaMalloc, no host loop, no seed-cell round trip
 * through host memory - one upload of kilobytes of tables, one persistent kernel, one download of the
 * count tensor.  There is no CPU fallback: without a CUDA device every simulate call returns
 * PROCELL_ERR_CUDA.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "host_plan.h"
#include "procell_math_tables.inc"
#include "sim_kernels.h"
#include "fitness_device.h"

#include <chrono>
#include <cstdlib>
#include <dlfcn.h>
#include <thread>

using namespace procell_b200;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e == cudaSuccess) cap = bytes ? bytes : 16;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

/* the math table of procell_spec.h (bit patterns): log rows, sin/cos rows and ziggurat rows - the part the kernels copy
 * into shared memory - then the ziggurat's wedge rows, which stay in HBM */
struct MathTable {
    uint64_t log_rows[1 << PCM_LOG_N_BITS][2];
    uint64_t sincos_rows[1 << PCM_SC_N_BITS][2];
    uint64_t zig_rows[1 << PCM_ZIG_N_BITS][2];
    uint64_t zig_wedge_rows[1 << PCM_ZIG_N_BITS][2];
};
const MathTable kLogRows = { { PCM_LOG_TABLE_ROWS }, { PCM_SINCOS_TABLE_ROWS }, { PCM_ZIG_TABLE_ROWS }, { PCM_ZIG_WEDGE_ROWS } };
static_assert(offsetof(MathTable, zig_wedge_rows) == (size_t)kLogTabDoubles * 8, "shared-memory part of the math table");
static_assert(sizeof(MathTable) == (size_t)kMathTabDoubles * 8, "math table size");

void set_round_keys(SimParams& P, uint64_t seed)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        P.rk[2 * r] = k0;
        P.rk[2 * r + 1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

int cuda_fail(cudaError_t e, const char* what)
{
    return fail(PROCELL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(call, what)                                   \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__, what); \
    } while (0)

}  // namespace

struct procell_engine {
    int device = 0;
    int sm_count = 0;
    DevBuf tables, logtab;                  /* tables: plan + type tables, one packed allocation */
    /* pinned host staging for the packed tables: two buffers used alternately, so that the upload of load i+1 may be
     * written while the asynchronous copy of load i is still in flight */
    void* stage[2] = { nullptr, nullptr };
    size_t stage_cap[2] = { 0, 0 };
    cudaEvent_t stage_done[2] = { nullptr, nullptr };   /* recorded behind the copy that reads stage[k] */
    int stage_next = 0;
    cudaStream_t up_stream = nullptr;       /* the table upload runs here, ordered behind the previous run (ev1) */
    cudaEvent_t up_done = nullptr;          /* the next run waits for it on its own stream */
    bool up_pending = false;
    bool ran = false;                       /* a run has been queued since creation (ev1 is recorded) */
    long long* last_counts = nullptr;       /* where the last run wrote: the engine's tensor or the caller's */
    long long* last_divisions = nullptr;
    bool fit_in_launch = false;             /* the last run computed the sweep fitness itself (fit_out holds it) */
    bool fit_last_from_launch = false;      /* ... and the last procell_engine_fitness call returned that */
    DevBuf dbg, fit_key_channel, fit_target, fit_out;
    uint32_t fit_channels = 0;
    std::vector<double> plan_row_value;     /* copies of what fitness needs, so the plan may be destroyed after load */
    std::vector<uint32_t> plan_key_row;
    DevBuf counts, ctl, q_seq, q_data, spill;   /* counts = count tensor followed by the division counters */
    SimParams P{};
    bool loaded = false;
    int kernel = PROCELL_KERNEL_COOP;
    int warps = 16;
    int ring = 1;                           /* 2: 256-node rings, two nodes per lane (16 warps) */
    int grid = 0, block = 0;
    size_t smem = 0;
    size_t counts_len = 0;
    size_t n_sets = 0;
    size_t n_times = 1;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int launches_last = 0;
};

extern "C" {

int procell_engine_create(int device, procell_engine** out)
{
    if (!out) return fail(PROCELL_ERR_ARG, "procell_engine_create: null out");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(PROCELL_ERR_CUDA, std::string("no CUDA device available (this library has no CPU path): ") +
                                          cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(PROCELL_ERR_ARG, "device index out of range");
    CU(cudaSetDevice(device), "cudaSetDevice");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(PROCELL_ERR_CUDA, "device is not compute capability 10.x: the kernels are built for sm_100a only");
    procell_engine* en = new procell_engine();
    en->device = device;
    en->sm_count = prop.multiProcessorCount;
    if (en->logtab.reserve(sizeof(kLogRows)) != cudaSuccess ||
        cudaMemcpy(en->logtab.p, &kLogRows, sizeof(kLogRows), cudaMemcpyHostToDevice) != cudaSuccess ||
        en->ctl.reserve(sizeof(ControlBlock)) != cudaSuccess ||
        en->q_seq.reserve(sizeof(unsigned long long) * kQueueCap) != cudaSuccess ||
        en->q_data.reserve(sizeof(unsigned long long) * (size_t)kQueueCap * kChunkWords) != cudaSuccess ||
        cudaMemset(en->ctl.p, 0, sizeof(ControlBlock)) != cudaSuccess ||      /* the status word is sticky: k_queue_init never clears it */
        cudaStreamCreateWithFlags(&en->up_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&en->up_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&en->stage_done[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&en->stage_done[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&en->ev0) != cudaSuccess || cudaEventCreate(&en->ev1) != cudaSuccess) {
        cudaError_t le = cudaGetLastError();
        procell_engine_destroy(en);
        return cuda_fail(le, "engine allocation");
    }
    *out = en;
    return PROCELL_OK;
}

void procell_engine_destroy(procell_engine* en)
{
    if (!en) return;
    cudaSetDevice(en->device);
    DevBuf* bufs[] = { &en->tables, &en->logtab, &en->counts, &en->dbg, &en->fit_key_channel, &en->fit_target, &en->fit_out, &en->ctl, &en->q_seq, &en->q_data, &en->spill };
    for (DevBuf* b : bufs) b->release();
    if (en->up_stream) { cudaStreamSynchronize(en->up_stream); cudaStreamDestroy(en->up_stream); }
    for (int k = 0; k < 2; ++k) {
        if (en->stage[k]) cudaFreeHost(en->stage[k]);
        if (en->stage_done[k]) cudaEventDestroy(en->stage_done[k]);
    }
    if (en->up_done) cudaEventDestroy(en->up_done);
    if (en->ev0) cudaEventDestroy(en->ev0);
    if (en->ev1) cudaEventDestroy(en->ev1);
    delete en;
}

size_t procell_engine_counts_len(const procell_engine* en) { return en ? en->counts_len : 0; }

int procell_engine_load(procell_engine* en, const procell_plan* plan, const procell_sim_params* sp)
{
    if (!en || !plan || !sp || !sp->types) return fail(PROCELL_ERR_ARG, "procell_engine_load: null argument");
    const size_t T = sp->n_types, S = sp->n_sets, B = plan->bin_value.size(), K = plan->n_keys;
    if (T == 0 || T > 64) return fail(PROCELL_ERR_ARG, "n_types must be 1..64");
    if (S == 0 || S > 65536) return fail(PROCELL_ERR_ARG, "n_sets must be 1..65536");
    const size_t M = (sp->checkpoints && sp->n_checkpoints) ? sp->n_checkpoints : 1;
    if (M > 8) return fail(PROCELL_ERR_ARG, "at most 8 checkpoints");
    double times[8];
    if (M == 1 && !(sp->checkpoints && sp->n_checkpoints)) times[0] = sp->t_max;
    else for (size_t j = 0; j < M; ++j) times[j] = sp->checkpoints[j];
    for (size_t j = 0; j < M; ++j)
        if (!(times[j] >= 0.0) || (j && !(times[j] > times[j - 1])))
            return fail(PROCELL_ERR_ARG, "t_max / checkpoints must be >= 0 and strictly ascending");
    if (M > 1 && sp->kernel == PROCELL_KERNEL_SIMPLE) return fail(PROCELL_ERR_ARG, "checkpoints need the cooperative kernel");
    if ((double)M * (double)S * (double)K * (double)T >= 4294967296.0)
        return fail(PROCELL_ERR_ARG, "n_checkpoints * n_sets * n_keys * n_types must be below 2^32");
    if (sp->shard_world > 1 && sp->shard_rank >= sp->shard_world) return fail(PROCELL_ERR_ARG, "shard_rank >= shard_world");
    const bool subtree = sp->shard_level > 0 && sp->shard_world > 1;
    if (subtree && (sp->shard_level > 30 || S != 1 || M != 1 || sp->kernel != PROCELL_KERNEL_COOP))
        return fail(PROCELL_ERR_ARG, "subtree sharding needs shard_level <= 30, the cooperative kernel, one parameter set and one checkpoint");
    for (size_t s = 0; s < S; ++s) {
        int rc = procell_check_proportions(sp->types + s * T, T);
        if (rc != PROCELL_OK) return rc;
    }
    CU(cudaSetDevice(en->device), "cudaSetDevice");

    /* histogram plan + type tables -> HBM: packed into ONE pinned staging buffer and uploaded with one copy.
     * layout (16-byte aligned pieces): bin_start[B+1] u32 | bin_keybase[B] u32 | bin_kdiv[B+1] u8 |
     *                                  type_thr[S][T rounded up to 4] u32 | type_musd[S*T] double2 | type_sel[S*T] u8 */
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t off_start = 0;
    const size_t off_keybase = off_start + align16((B + 1) * 4);
    const size_t off_kdiv = off_keybase + align16((B + 1) * 4);
    const size_t off_cum = off_kdiv + align16(B + 1);
    const size_t Tpad = (T + 3) & ~(size_t)3;          /* thresholds are scanned four at a time (one 16-byte load) */
    const size_t off_musd = off_cum + align16(S * Tpad * 4);
    const size_t off_sel = off_musd + align16(S * T * 16);
    const size_t off_rank = off_sel + align16(S * T);          /* set 0's types by kind: proliferating ids, then quiescent ids */
    const size_t table_bytes = off_rank + align16(T);
    const int sk = en->stage_next;
    en->stage_next ^= 1;
    CU(cudaEventSynchronize(en->stage_done[sk]), "wait for the staging buffer");   /* its last copy (two loads ago) is over */
    if (table_bytes > en->stage_cap[sk]) {
        if (en->stage[sk]) cudaFreeHost(en->stage[sk]);
        en->stage[sk] = nullptr; en->stage_cap[sk] = 0;
        CU(cudaMallocHost(&en->stage[sk], table_bytes), "alloc pinned staging");
        en->stage_cap[sk] = table_bytes;
    }
    unsigned char* st = static_cast<unsigned char*>(en->stage[sk]);
    uint32_t* h_start = reinterpret_cast<uint32_t*>(st + off_start);
    uint32_t* h_keybase = reinterpret_cast<uint32_t*>(st + off_keybase);
    uint8_t* h_kdiv = st + off_kdiv;
    uint32_t* h_thr = reinterpret_cast<uint32_t*>(st + off_cum);
    double2* h_musd = reinterpret_cast<double2*>(st + off_musd);
    uint8_t* h_sel = st + off_sel;
    {
        uint8_t* h_rank = st + off_rank;
        size_t n = 0;
        for (size_t j = 0; j < T; ++j) if (!(sp->types[j].mean < 0.0)) h_rank[n++] = (uint8_t)j;
        for (size_t j = 0; j < T; ++j) if (sp->types[j].mean < 0.0) h_rank[n++] = (uint8_t)j;
    }
    for (size_t b = 0; b <= B; ++b) h_start[b] = (uint32_t)plan->bin_start[b];
    for (size_t b = 0; b < B; ++b) {
        h_keybase[b] = plan->bin_keybase[b];
        h_kdiv[b] = (uint8_t)(plan->bin_kdiv[b] | (plan->bin_count0[b] ? 0x80u : 0u));
    }
    h_kdiv[B] = 0;
    /* type tables: selection order = descending proportion, stable (parser.cu:184; thrust::sort is not
     * stable - ties keep file order here), cumulative sums accumulated in that order (cell.cu:88) */
    for (size_t s = 0; s < S; ++s) {
        const procell_cell_type* ty = sp->types + s * T;
        int order[64];
        for (size_t j = 0; j < T; ++j) order[j] = (int)j;
        std::stable_sort(order, order + T, [&](int a, int b) { return ty[a].proportion > ty[b].proportion; });
        double acc = 0.0;
        for (size_t j = T; j < Tpad; ++j) h_thr[s * Tpad + j] = 0xFFFFFFFFu;
        for (size_t j = 0; j < T; ++j) {
            acc += ty[order[j]].proportion;
            /* the scan compares the 32-bit type word x with "last x below cum[j]" = threshold - 1 (hostio.cpp:
             * procell_type_threshold).  The first running sum is the largest proportion, >= (1 - 1e-8) / 64, so no
             * threshold is 0; a running sum of 1 gives 2^32 - 1, which no x exceeds. */
            const uint64_t thr = procell_type_threshold(acc);
            if (thr == 0) return fail(PROCELL_ERR_PROPORTION, "cell-type proportions: a running sum below 2^-33");
            /* the scan counts the thresholds x lies above among the first T - 1; the last type's slot and the padding
             * up to a multiple of four hold 2^32 - 1, which no x exceeds */
            h_thr[s * Tpad + j] = j + 1 < T ? (uint32_t)(thr - 1) : 0xFFFFFFFFu;
            h_sel[s * T + j] = (uint8_t)order[j];
            h_musd[s * T + j] = make_double2(ty[j].mean, ty[j].stddev);
        }
    }
    /* asynchronous upload on the engine's own stream: ordered behind the previous run of this engine (which may still
     * be reading the tables) and ahead of the next one (procell_engine_run makes its stream wait for up_done); the host
     * does not wait, so with two engines in flight the copy hides behind the other engine's kernel */
    if (table_bytes > en->tables.cap || !en->tables.p) {
        if (en->ran) CU(cudaEventSynchronize(en->ev1), "previous run");      /* the old allocation may still be in use */
        CU(en->tables.reserve(table_bytes), "alloc tables");
    }
    if (en->ran) CU(cudaStreamWaitEvent(en->up_stream, en->ev1, 0), "order upload behind the previous run");
    CU(cudaMemcpyAsync(en->tables.p, st, table_bytes, cudaMemcpyHostToDevice, en->up_stream), "upload tables");
    CU(cudaEventRecord(en->stage_done[sk], en->up_stream), "event record");
    CU(cudaEventRecord(en->up_done, en->up_stream), "event record");
    en->up_pending = true;
    unsigned char* dt = static_cast<unsigned char*>(en->tables.p);

    en->counts_len = M * S * K * T;
    en->n_sets = S;
    en->n_times = M;
    en->plan_row_value = plan->row_value;
    en->plan_key_row = plan->key_row;
    en->fit_channels = 0;
    /* one allocation: the count tensor followed by the division counters, so that a multi-GPU run needs ONE reduce */
    if ((en->counts_len + S) * 8 > en->counts.cap && en->ran) CU(cudaEventSynchronize(en->ev1), "previous run");
    CU(en->counts.reserve((en->counts_len + S) * 8), "alloc counts");

    SimParams& P = en->P;
    P.bin_start = (const uint32_t*)(dt + off_start);
    P.bin_keybase = (const uint32_t*)(dt + off_keybase);
    P.bin_kdiv = (const uint8_t*)(dt + off_kdiv);
    P.type_thr = (const uint32_t*)(dt + off_cum);
    P.type_sel = (const uint8_t*)(dt + off_sel);
    P.type_musd = (const double2*)(dt + off_musd);
    P.logtab = (const double*)en->logtab.p;
    P.counts = (long long*)en->counts.p;
    P.divisions = (long long*)en->counts.p + en->counts_len;
    P.ctl = (ControlBlock*)en->ctl.p;
    P.q_seq = (unsigned long long*)en->q_seq.p;
    P.q_data = (unsigned long long*)en->q_data.p;
    P.n_bins = (uint32_t)B; P.n_types = (uint32_t)T; P.n_sets = (uint32_t)S; P.n_keys = (uint32_t)K;
    P.n_cells = (uint32_t)plan->n_cells;
    P.shard_world = (sp->shard_world > 1 && !subtree) ? sp->shard_world : 1;
    P.shard_rank = (sp->shard_world > 1 && !subtree) ? sp->shard_rank : 0;
    P.sub_world = subtree ? sp->shard_world : 1;      /* subtree sharding: every GPU claims every seed unit */
    P.sub_rank = subtree ? sp->shard_rank : 0;
    P.sub_limit = subtree ? 1ull << sp->shard_level : 0ull;
    P.refcompat = sp->seeding_mode == PROCELL_SEEDING_REFCOMPAT;
    P.t_max = times[M - 1];
    P.n_times = (uint32_t)M;
    P.time_stride = (uint32_t)(S * K * T);
    for (size_t j = 0; j < 8; ++j) P.times[j] = j < M ? times[j] : times[M - 1];
    set_round_keys(P, sp->seed);

    /* claim unit: 32 seed cells = one SEED iteration of a warp (small units keep the tail balanced: a warp claims
     * a new unit with one atomic whenever its stack runs low); smaller still when there are too few cells to
     * give every warp of every GPU several units, larger when there are plenty */
    uint32_t unit = sp->shard_unit;
    if (unit == 0) {
        const double per_warp = (double)plan->n_cells * (double)S / (148.0 * kCoopWarpsMax * 4.0 * P.shard_world);
        unit = 32;
        while (unit > 1 && (double)unit > per_warp) unit >>= 1;
#ifndef PROCELL_UNIT_MAX
#define PROCELL_UNIT_MAX 256u
#endif
        /* large inputs: up to 256 cells per claim (8 SEED iterations) while every warp still gets 64 units or more */
        while (unit < PROCELL_UNIT_MAX && (double)unit * 32.0 <= per_warp) unit <<= 1;
    }
    P.unit = unit;
    P.units_per_set = (uint32_t)((plan->n_cells + unit - 1) / unit);
    P.local_units_per_set = P.units_per_set > P.shard_rank
                                ? (P.units_per_set - P.shard_rank + P.shard_world - 1) / P.shard_world : 0;
    P.total_local_units = (unsigned long long)P.local_units_per_set * S;
    {   /* sweep batching: about four batches per set, so a CTA stays on one parameter set for a long stretch */
        uint32_t per_set = P.local_units_per_set < 4u ? (P.local_units_per_set ? P.local_units_per_set : 1u) : 4u;
        uint32_t bu = (P.local_units_per_set + per_set - 1) / per_set;
        if (bu == 0) bu = 1;
        if (bu > (1u << 22)) bu = 1u << 22;
        P.batch_units = bu;
        P.batches_per_set = P.local_units_per_set ? (P.local_units_per_set + bu - 1) / bu : 1;
        P.total_batches = P.local_units_per_set ? (unsigned long long)P.batches_per_set * S : 0;
    }

    const char* denv = getenv("PROCELL_NO_DONATE");
    P.donate = !(denv && atoi(denv) == 1);
    en->kernel = sp->kernel;
    P.hist_setdirect = 0;
    P.leaf_merge = 0;
    P.slot_mode = 0;
    P.kstride = (uint32_t)T;
    if (en->kernel == PROCELL_KERNEL_SIMPLE) {
        en->block = kSimpleThreads;
        en->grid = en->sm_count * 8;
        en->smem = kLogTabDoubles * 8;
        P.smem_hist_slots = 0;
        P.hist_hashed = 0;
        P.spill = nullptr;
    } else {
        int max_smem = 0;
        CU(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, en->device), "query smem");
        const char* wenv = getenv("PROCELL_COOP_WARPS");     /* tuning knob: 16, 24 or 32 warps per CTA */
        const int wreq = wenv ? atoi(wenv) : 0;
        en->warps = (wreq == 16 || wreq == 24) ? wreq : 32;
        const char* renv = getenv("PROCELL_COOP_NPL");       /* tuning knob: 2 = 16 warps, two nodes per lane */
        (void)renv;                                          /* the two-nodes-per-lane instance is gone (round 2: the cheap
                                                                ziggurat draw left it nothing to interleave) */
        en->ring = 1;
        if (subtree) { en->warps = 32; en->ring = 1; }       /* the subtree-sharding instances exist in the product shape only */
        const size_t fixed = coop_smem_bytes(en->warps, en->ring, 0, 0);
        const size_t room = (size_t)max_smem > fixed ? (size_t)max_smem - fixed : 0;
        /* sweeps whose whole key space does not fit but ONE set's does: a direct table of the CTA's current set
         * (kernel MODE kModeSetDirect).  The CTA's warps meet at every batch switch, which pays off only when a batch
         * keeps them busy for a while: measured on a B200, config 5 at full size (30 units per warp and batch) runs in
         * 60.7 ms instead of 70.4 ms, a tenth of it (3 units per warp and batch) in 8.0 ms instead of 6.8 ms.  Default:
         * on from 24 units per warp and batch; PROCELL_SWEEP_DIRECT=1 / 0 forces it on / off. */
        const char* sdenv = getenv("PROCELL_SWEEP_DIRECT");
        const bool sd_forced_on = sdenv && atoi(sdenv) == 1, sd_forced_off = sdenv && atoi(sdenv) == 0;
        const bool sd_wanted = sd_forced_on || (!sd_forced_off && (double)P.batch_units >= 24.0 * 32.0);
        P.hist_setdirect = 0;
        /* one parameter set, one checkpoint (the PLAIN instances): the table is laid out by SLOTS - proliferating types
         * per key, then one row per bin for the quiescent types, which are only ever counted at level 0 (sim_kernels.h) */
        const bool plain = S == 1 && M == 1;
        uint32_t n_prolif = 0, n_quiet = 0;
        P.quiet_mask = 0ull;
        for (size_t j = 0; j < T; ++j) {
            if (sp->types[j].mean < 0.0) { P.quiet_mask |= 1ull << j; ++n_quiet; } else ++n_prolif;
        }
        P.rank_type = (const uint8_t*)(dt + off_rank);
        const size_t slot_count = K * n_prolif + B * n_quiet;
        P.kstride = (uint32_t)T;
        P.slot_mode = 0;
        P.n_prolif = n_prolif; P.n_quiet = n_quiet;
        P.slot_prolif_end = (uint32_t)(K * n_prolif);
        if (plain && slot_count * 4 <= room) {
            P.hist_hashed = 0;
            P.slot_mode = 1;
            P.kstride = n_prolif;
            P.smem_hist_slots = (uint32_t)slot_count;
        } else if (!plain && en->counts_len * 4 <= room) {   /* the whole key space fits: direct u32 table indexed by key */
            P.hist_hashed = 0;
            P.smem_hist_slots = (uint32_t)en->counts_len;
        } else if (sd_wanted && S > 1 && M == 1 && !subtree && en->warps == 32 && en->ring == 1 && K * T * 4 <= room) {
            P.hist_hashed = 0;
            P.hist_setdirect = 1;
            P.smem_hist_slots = (uint32_t)(K * T);
        } else {                                     /* direct-mapped {key,count} cache, power-of-two slots */
            uint32_t slots = 1;
            while ((size_t)slots * 2 * 8 <= room) slots *= 2;
            P.hist_hashed = 1;
            P.smem_hist_slots = slots;
        }
        if (plain && !P.hist_hashed && !subtree && en->warps == 32 && en->ring == 1) {
            /* Deep lineage trees -> the instance that merges equal leaf keys before the shared-memory atomic (sim_kernels.cu,
             * kModeMerge): it pays where DIVIDE iterations are nearly all of the work and the shared-memory pipe is the bound,
             * and costs 3 % where seed cells are a third of the draws.  Expected depth of a lineage = the smaller of the
             * generations t_max leaves room for (fastest type) and the halvings phi allows (mean over the seed cells). */
            const double depth = procell_plan_lineage_depth(plan, sp->types, T, P.t_max);
            /* ... and only where there is work for every warp for a while: a run that is mostly its tail (1e5 cells of
             * config 2's shape: 0.2 ms) is bound by latency, and the merge adds to it (measured: -3.6 %) */
            P.leaf_merge = depth >= 6.0 && (double)plan->n_cells * std::exp2(depth < 30.0 ? depth : 30.0) >= 5e8;
            const char* lm = getenv("PROCELL_LEAF_MERGE");       /* 0 / 1 forces the choice (tests, A/B) */
            if (lm) P.leaf_merge = atoi(lm) != 0;
        }
        en->smem = coop_smem_bytes(en->warps, en->ring, P.smem_hist_slots, P.hist_hashed);
        int grid = 0;
        if (P.hist_setdirect) CU(coop_max_grid_setdirect(en->device, en->smem, &grid), "occupancy query");
        else if (P.leaf_merge) CU(coop_max_grid_merge(en->device, en->smem, &grid), "occupancy query");
        else if (subtree) CU(coop_max_grid_subtree(en->device, P.hist_hashed, en->smem, &grid), "occupancy query");
        else CU(coop_max_grid(en->device, en->warps, en->ring, P.hist_hashed, (P.n_sets == 1u && P.n_times == 1u) ? 1 : 0, en->smem, &grid), "occupancy query");
        if (grid <= 0) return fail(PROCELL_ERR_CUDA, "cooperative kernel does not fit on this device");
        en->grid = grid;
        en->block = en->warps * 32;
        CU(en->spill.reserve((size_t)grid * en->warps * kSpillCap * kChunkWords * 8), "alloc spill rings");
        P.spill = (unsigned long long*)en->spill.p;
        const size_t dbg_bytes = (size_t)grid * en->warps * kDbgWords * 8;
        if (dbg_bytes > en->dbg.cap || !en->dbg.p) {
            CU(en->dbg.reserve(dbg_bytes), "alloc debug records");
            CU(cudaMemset(en->dbg.p, 0, dbg_bytes), "clear debug records");
        }
        P.dbg = (unsigned long long*)en->dbg.p;
        const char* wd = getenv("PROCELL_WATCHDOG_S");       /* abort a launch whose warps run longer than this */
        const double wd_s = wd && atof(wd) > 0 ? atof(wd) : 3600.0;
        P.watchdog_ns = (unsigned long long)(wd_s * 1e9);
    }
    en->loaded = true;
    return PROCELL_OK;
}

int procell_engine_run(procell_engine* en, uint64_t seed, void* stream_v, int64_t* d_counts, int64_t* d_divisions)
{
    if (!en || !en->loaded) return fail(PROCELL_ERR_ARG, "procell_engine_run: engine not loaded");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CU(cudaSetDevice(en->device), "cudaSetDevice");
    SimParams P = en->P;
    set_round_keys(P, seed);
    if (d_counts) P.counts = (long long*)d_counts;
    if (d_divisions) P.divisions = (long long*)d_divisions;
    en->timed = true;
    en->last_counts = P.counts;
    en->last_divisions = P.divisions;
    /* sweep fitness in the same launch: a target is set, the run is an unsharded sweep on the cooperative kernel, and
     * the accumulators fit in the (by then flushed) shared-memory table.  PROCELL_FITNESS_FUSED=0 keeps the separate pass. */
    P.fit_channels = 0;
    en->fit_in_launch = false;
    {
        const char* fenv = getenv("PROCELL_FITNESS_FUSED");
        const size_t table_bytes = (size_t)P.smem_hist_slots * (P.hist_hashed ? 8 : 4);
        if (en->fit_channels != 0 && en->kernel == PROCELL_KERNEL_COOP && !(P.n_sets == 1u && P.n_times == 1u) &&
            P.shard_world == 1u && P.sub_world == 1u && !(fenv && atoi(fenv) == 0) &&
            fitness_smem_bytes(en->fit_channels) <= table_bytes) {
            P.fit_channels = en->fit_channels;
            P.fit_key_channel = (const uint32_t*)en->fit_key_channel.p;
            P.fit_target = (const double*)en->fit_target.p;
            P.fit_out = (double*)en->fit_out.p;
            en->fit_in_launch = true;
        }
    }
    if (en->up_pending) {       /* the tables of the last load are on their way on the upload stream */
        CU(cudaStreamWaitEvent(stream, en->up_done, 0), "order run behind the table upload");
        en->up_pending = false;
    }
    CU(cudaEventRecord(en->ev0, stream), "event record");
    /* one launch resets queue + control block and zeroes the count tensor and the division counters */
    CU(launch_queue_init(P.q_seq, P.ctl, P.counts, en->counts_len, P.divisions, en->n_sets, en->sm_count, stream), "launch k_queue_init");
    if (en->kernel == PROCELL_KERNEL_SIMPLE) CU(launch_simple(P, en->grid, stream), "launch k_proliferate_simple");
    else CU(launch_coop(P, en->warps, en->ring, en->grid, stream), "launch k_proliferate_coop");
    en->launches_last = 2;
    CU(cudaEventRecord(en->ev1, stream), "event record");
    en->ran = true;
    return PROCELL_OK;
}

int procell_engine_finish(procell_engine* en, void* stream_v, int64_t* counts, int64_t* divisions, procell_run_stats* stats)
{
    if (!en || !en->loaded) return fail(PROCELL_ERR_ARG, "procell_engine_finish: engine not loaded");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CU(cudaSetDevice(en->device), "cudaSetDevice");
    const char* hto = getenv("PROCELL_HOST_TIMEOUT_S");     /* debugging aid: give up on a launch that does not end */
    if (hto && atof(hto) > 0) {
        cudaEvent_t done;
        CU(cudaEventCreateWithFlags(&done, cudaEventDisableTiming), "event create");
        CU(cudaEventRecord(done, stream), "event record");
        const double limit = atof(hto);
        const auto t_begin = std::chrono::steady_clock::now();
        while (cudaEventQuery(done) == cudaErrorNotReady) {
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() > limit) {
                std::string msg = "launch still running after host timeout;";
                if (en->dbg.p) {
                    cudaStream_t side;
                    cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
                    const size_t nw = (size_t)en->grid * en->warps;
                    std::vector<unsigned long long> rec(nw * kDbgWords);
                    cudaMemcpyAsync(rec.data(), en->dbg.p, rec.size() * 8, cudaMemcpyDeviceToHost, side);
                    cudaStreamSynchronize(side);
                    std::vector<size_t> hist(100, 0);
                    for (size_t i = 0; i < nw; ++i) hist[rec[i * kDbgWords + 1] % 100] += 1;
                    msg += " warps by trace code:";
                    for (size_t c = 0; c < 100; ++c)
                        if (hist[c]) msg += " " + std::to_string(c) + ":" + std::to_string(hist[c]);
                    ControlBlock cbs;
                    cudaMemcpyAsync(&cbs, en->ctl.p, sizeof cbs, cudaMemcpyDeviceToHost, side);
                    cudaStreamSynchronize(side);
                    msg += " | active " + std::to_string(cbs.active) + " idle " + std::to_string(cbs.idle) + " avail " +
                           std::to_string(cbs.avail) + " cursor " + std::to_string(cbs.cursor) + " head " +
                           std::to_string(cbs.q_head) + " tail " + std::to_string(cbs.q_tail) + " status " + std::to_string(cbs.status);
                    cudaStreamDestroy(side);
                }
                cudaEventDestroy(done);
                return fail(PROCELL_ERR_OVERFLOW, msg);
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        cudaEventDestroy(done);
    }
    CU(cudaStreamSynchronize(stream), "kernel execution");
    ControlBlock cb;
    CU(cudaMemcpy(&cb, en->ctl.p, sizeof(ControlBlock), cudaMemcpyDeviceToHost), "read status");
    const int status = cb.status;
    if (status != kStatusOk) {
        /* the word is sticky on the device (any run since the last finish may have set it): clear it now that it is
         * being reported, so that the engine can be used again */
        cudaMemset(reinterpret_cast<unsigned char*>(en->ctl.p) + offsetof(ControlBlock, status), 0, sizeof(int));
        static const char* const kStatusText[] = { "ok", "spill ring overflow", "donation queue timeout", "idle-wait timeout",
                                                   "watchdog", "dynamic shared memory does not start where the kernel was compiled for (sm_100: 0x400)" };
        std::string msg = "device work pool failure, status " + std::to_string(status);
        if (status > 0 && status < (int)(sizeof(kStatusText) / sizeof(kStatusText[0]))) msg += std::string(" (") + kStatusText[status] + ")";
        if (status == kStatusWatchdog && en->dbg.p) {      /* where were the warps when the watchdog fired */
            const size_t nw = (size_t)en->grid * en->warps;
            std::vector<unsigned long long> rec(nw * kDbgWords);
            if (cudaMemcpy(rec.data(), en->dbg.p, rec.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
                int shown = 0;
                size_t fired = 0;
                for (size_t i = 0; i < nw; ++i) fired += rec[i * kDbgWords] != 0;
                msg += "; " + std::to_string(fired) + " of " + std::to_string(nw) + " warps reported:";
                for (size_t i = 0; i < nw && shown < 12; ++i) {
                    const unsigned long long* r = &rec[i * kDbgWords];
                    if (!r[0]) continue;
                    msg += " [w" + std::to_string(i) + " code " + std::to_string(r[0]);
                    for (int k = 1; k < 7; ++k) msg += " " + std::to_string(r[k]);
                    msg += "]";
                    ++shown;
                }
            }
        }
        return fail(PROCELL_ERR_OVERFLOW, msg);
    }
    /* results are read from where the LAST run wrote them: the engine's own tensor or the caller's device buffers */
    const long long* src_counts = en->last_counts ? en->last_counts : (const long long*)en->counts.p;
    const long long* src_div = en->last_divisions ? en->last_divisions : (const long long*)en->counts.p + en->counts_len;
    if (counts) CU(cudaMemcpy(counts, src_counts, en->counts_len * 8, cudaMemcpyDeviceToHost), "download counts");
    std::vector<int64_t> div(en->n_sets);
    CU(cudaMemcpy(div.data(), src_div, en->n_sets * 8, cudaMemcpyDeviceToHost), "download divisions");
    if (divisions) memcpy(divisions, div.data(), en->n_sets * 8);
    if (stats) {
        stats->divisions = 0;
        for (int64_t d : div) stats->divisions += d;
        float ms = 0.f;
        if (en->timed) cudaEventElapsedTime(&ms, en->ev0, en->ev1);
        stats->kernel_ms = ms;
        stats->n_launches = en->launches_last;
        stats->grid = en->grid; stats->block = en->block; stats->smem_bytes = (int)en->smem;
        stats->donations = (int64_t)cb.q_tail;
        stats->seed_phase_us = cb.t_exhausted == ~0ull ? -1.0 : (double)(cb.t_exhausted - cb.t_start) * 1e-3;
        stats->total_us = cb.t_end ? (double)(cb.t_end - cb.t_start) * 1e-3 : -1.0;
        stats->idle_warp_us = (double)cb.idle_ns * 1e-3;
        stats->idle_waits = (int64_t)cb.idle_waits;
    }
    return PROCELL_OK;
}

int procell_proliferate(const procell_plan* plan, const procell_sim_params* params, int device, int64_t* counts,
                        int64_t* divisions, procell_run_stats* stats)
{
    if (!plan || !params || !counts) return fail(PROCELL_ERR_ARG, "procell_proliferate: null argument");
    procell_engine* en = nullptr;
    int rc = procell_engine_create(device, &en);
    if (rc != PROCELL_OK) return rc;
    rc = procell_engine_load(en, plan, params);
    if (rc == PROCELL_OK) rc = procell_engine_run(en, params->seed, nullptr, nullptr, nullptr);
    if (rc == PROCELL_OK) rc = procell_engine_finish(en, nullptr, counts, divisions, stats);
    procell_engine_destroy(en);
    return rc;
}

/* ---- on-GPU fitness of a sweep -------------------------------------------------------------- */
int procell_engine_set_target(procell_engine* en, const double* value, const uint64_t* freq, size_t n_channels)
{
    if (!en || !en->loaded) return fail(PROCELL_ERR_ARG, "procell_engine_set_target: engine not loaded");
    if (!value || !freq || n_channels == 0 || n_channels > 16384)
        return fail(PROCELL_ERR_ARG, "target histogram must have 1..16384 channels");
    for (size_t c = 1; c < n_channels; ++c)
        if (!(value[c] > value[c - 1])) return fail(PROCELL_ERR_ARG, "target channel values must be strictly ascending");
    CU(cudaSetDevice(en->device), "cudaSetDevice");
    const std::vector<double>& row_value = en->plan_row_value;
    const std::vector<uint32_t>& key_row = en->plan_key_row;
    const size_t n_keys = key_row.size();
    /* utils::rebin (src/utils/util.cu:111-138): a value goes to the first channel whose value is >= it */
    std::vector<uint32_t> row_channel(row_value.size());
    size_t pos = 0;
    for (size_t r = 0; r < row_value.size(); ++r) {
        while (pos + 1 < n_channels && row_value[r] > value[pos]) ++pos;
        row_channel[r] = (uint32_t)pos;
    }
    std::vector<uint32_t> key_channel(n_keys + 1, 0xFFFFFFFFu);
    for (size_t k = 0; k < n_keys; ++k)
        if (key_row[k] != 0xFFFFFFFFu) key_channel[k] = row_channel[key_row[k]];
    double total = 0.0;
    for (size_t c = 0; c < n_channels; ++c) total += (double)freq[c];
    if (!(total > 0.0)) return fail(PROCELL_ERR_ARG, "target histogram is empty");
    std::vector<double> share(n_channels);
    for (size_t c = 0; c < n_channels; ++c) share[c] = (double)freq[c] / total;
    CU(en->fit_key_channel.reserve((n_keys + 1) * 4), "alloc key_channel");
    CU(en->fit_target.reserve(n_channels * 8), "alloc target");
    CU(en->fit_out.reserve(en->n_sets * 8), "alloc fitness");
    CU(cudaMemcpy(en->fit_key_channel.p, key_channel.data(), (n_keys + 1) * 4, cudaMemcpyHostToDevice), "upload key_channel");
    CU(cudaMemcpy(en->fit_target.p, share.data(), n_channels * 8, cudaMemcpyHostToDevice), "upload target");
    en->fit_channels = (uint32_t)n_channels;
    en->fit_in_launch = false;          /* whatever the last run computed was for another target */
    return PROCELL_OK;
}

int procell_engine_fitness(procell_engine* en, void* stream_v, const int64_t* d_counts, double* fitness)
{
    if (!en || !en->loaded || en->fit_channels == 0 || !fitness)
        return fail(PROCELL_ERR_ARG, "procell_engine_fitness: no target set");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CU(cudaSetDevice(en->device), "cudaSetDevice");
    const long long* counts = d_counts ? (const long long*)d_counts : (const long long*)en->counts.p;
    /* the run that produced this tensor computed the distances in its own launch: only 8 bytes per set are left to do */
    en->fit_last_from_launch = en->fit_in_launch && counts == en->last_counts;
    if (en->fit_last_from_launch) {
        CU(cudaMemcpyAsync(fitness, en->fit_out.p, en->n_sets * 8, cudaMemcpyDeviceToHost, stream), "download fitness");
        CU(cudaStreamSynchronize(stream), "fitness download");
        return PROCELL_OK;
    }
    counts += (en->n_times - 1) * (size_t)en->P.time_stride;       /* time series: fitness of the last checkpoint */
    CU(launch_sweep_fitness(counts, (const uint32_t*)en->fit_key_channel.p, (const double*)en->fit_target.p,
                            (uint32_t)en->n_sets, en->P.n_keys, en->P.n_types, en->fit_channels, (double*)en->fit_out.p, stream),
       "launch k_sweep_fitness");
    CU(cudaMemcpyAsync(fitness, en->fit_out.p, en->n_sets * 8, cudaMemcpyDeviceToHost, stream), "download fitness");
    CU(cudaStreamSynchronize(stream), "fitness kernel");
    return PROCELL_OK;
}

int procell_engine_fitness_in_launch(const procell_engine* en) { return en && en->fit_last_from_launch ? 1 : 0; }

/* which instance of the cooperative kernel the loaded simulation runs on: 0 base, 1 subtree sharding, 2 sweep with a
 * set-relative table, 3 deep trees with merged leaf counts; -1: not the cooperative kernel / nothing loaded */
int procell_engine_kernel_mode(const procell_engine* en)
{
    if (!en || en->kernel == PROCELL_KERNEL_SIMPLE) return -1;
    if (en->P.hist_setdirect) return 2;
    if (en->P.sub_world > 1u) return 1;
    return en->P.leaf_merge ? 3 : 0;
}

/* ---- single-process multi-GPU: seed-cell units sharded over the GPUs of one box, ONE ncclReduce(sum, int64) ---- */
namespace {
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
    void* handle = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::once_flag once;
    bool ok = false;
    bool load()     /* thread-safe: the library is looked up once per process */
    {
        std::call_once(once, [this] { ok = load_once(); });
        return ok;
    }
    bool load_once()
    {
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
        Reduce = (decltype(Reduce))dlsym(handle, "ncclReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Reduce && GetErrorString;
    }
};
NcclApi g_nccl;
constexpr int kNcclInt64 = 4;   /* ncclInt64 */
constexpr int kNcclSum = 0;     /* ncclSum */
}  // namespace

int procell_proliferate_multi(const procell_plan* plan, const procell_sim_params* params, int n_gpus, int64_t* counts,
                              int64_t* divisions, procell_run_stats* stats)
{
    if (!plan || !params || !counts) return fail(PROCELL_ERR_ARG, "procell_proliferate_multi: null argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(PROCELL_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (n_gpus <= 0) n_gpus = n_dev;
    if (n_gpus > n_dev) return fail(PROCELL_ERR_ARG, "more GPUs requested than present");
    if (n_gpus == 1) return procell_proliferate(plan, params, 0, counts, divisions, stats);
    if (!g_nccl.load()) return fail(PROCELL_ERR_CUDA, "libnccl.so.2 not found: multi-GPU runs need NCCL");

    std::vector<procell_engine*> eng(n_gpus, nullptr);
    std::vector<cudaStream_t> streams(n_gpus, nullptr);
    std::vector<ncclComm_t> comms(n_gpus, nullptr);
    std::vector<int> devs(n_gpus);
    for (int i = 0; i < n_gpus; ++i) devs[i] = i;
    int rc = PROCELL_OK;
    auto cleanup = [&]() {
        for (int i = 0; i < n_gpus; ++i) {
            if (comms[i]) g_nccl.CommDestroy(comms[i]);
            if (streams[i]) { cudaSetDevice(i); cudaStreamDestroy(streams[i]); }
            procell_engine_destroy(eng[i]);
        }
    };
    for (int i = 0; i < n_gpus && rc == PROCELL_OK; ++i) {
        rc = procell_engine_create(i, &eng[i]);
        if (rc != PROCELL_OK) break;
        procell_sim_params sp = *params;
        sp.shard_rank = (uint32_t)i;
        sp.shard_world = (uint32_t)n_gpus;
        /* shard_unit == 0: every engine derives the same claim unit from (cells, sets, world) in procell_engine_load,
         * so the GPUs agree on who owns which unit - and large inputs get 256-cell units (config 3: -4 %) */
        rc = procell_engine_load(eng[i], plan, &sp);
        if (rc == PROCELL_OK && cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking) != cudaSuccess)
            rc = fail(PROCELL_ERR_CUDA, "cudaStreamCreate failed");
    }
    if (rc == PROCELL_OK) {
        int nrc = g_nccl.CommInitAll(comms.data(), n_gpus, devs.data());
        if (nrc != 0) rc = fail(PROCELL_ERR_CUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(nrc));
    }
    const size_t n_packed = eng[0] ? eng[0]->counts_len + eng[0]->n_sets : 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (rc == PROCELL_OK) {
        /* one small reduce first: a fresh communicator builds its channels and peer connections on first use, which
         * would otherwise be timed (and paid) inside the run's single exchange step */
        g_nccl.GroupStart();
        for (int i = 0; i < n_gpus; ++i) {
            int nrc = g_nccl.Reduce(eng[i]->counts.p, eng[i]->counts.p, 1, kNcclInt64, kNcclSum, 0, comms[i], streams[i]);
            if (nrc != 0) rc = fail(PROCELL_ERR_CUDA, std::string("ncclReduce (warm-up): ") + g_nccl.GetErrorString(nrc));
        }
        int nrc = g_nccl.GroupEnd();
        if (nrc != 0 && rc == PROCELL_OK) rc = fail(PROCELL_ERR_CUDA, std::string("ncclGroupEnd (warm-up): ") + g_nccl.GetErrorString(nrc));
        for (int i = 0; i < n_gpus; ++i) { cudaSetDevice(i); cudaStreamSynchronize(streams[i]); }
    }
    if (rc == PROCELL_OK) {
        cudaSetDevice(0);
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, streams[0]);
        for (int i = 0; i < n_gpus && rc == PROCELL_OK; ++i)
            rc = procell_engine_run(eng[i], params->seed, streams[i], nullptr, nullptr);
    }
    if (rc == PROCELL_OK) {     /* the single exchange step of the path: int64 sum onto GPU 0 over NVLink */
        g_nccl.GroupStart();
        for (int i = 0; i < n_gpus; ++i) {
            int nrc = g_nccl.Reduce(eng[i]->counts.p, eng[i]->counts.p, n_packed, kNcclInt64, kNcclSum, 0, comms[i], streams[i]);
            if (nrc != 0) rc = fail(PROCELL_ERR_CUDA, std::string("ncclReduce: ") + g_nccl.GetErrorString(nrc));
        }
        int nrc = g_nccl.GroupEnd();
        if (nrc != 0 && rc == PROCELL_OK) rc = fail(PROCELL_ERR_CUDA, std::string("ncclGroupEnd: ") + g_nccl.GetErrorString(nrc));
        cudaSetDevice(0);
        cudaEventRecord(e1, streams[0]);
    }
    /* every GPU's status word is checked; GPU 0 then holds the reduced tensor */
    procell_run_stats st0;
    memset(&st0, 0, sizeof st0);
    for (int i = n_gpus - 1; i >= 0 && rc == PROCELL_OK; --i)
        rc = procell_engine_finish(eng[i], streams[i], i == 0 ? counts : nullptr, i == 0 ? divisions : nullptr, i == 0 ? &st0 : nullptr);
    if (rc == PROCELL_OK && stats) {
        *stats = st0;
        float ms = 0.f;
        cudaSetDevice(0);
        cudaEventElapsedTime(&ms, e0, e1);
        stats->kernel_ms = ms;
    }
    if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
    cleanup();
    return rc;
}

int procell_simulate(const procell_input* in, procell_output* out)
{
    if (!in || !out) return fail(PROCELL_ERR_ARG, "procell_simulate: null argument");
    memset(out, 0, sizeof(*out));
    if (!in->types || in->n_types == 0) return fail(PROCELL_ERR_ARG, "procell_simulate: no cell types");
    const size_t n_sets = in->n_param_sets ? in->n_param_sets : 1;
    for (size_t s = 0; s < n_sets; ++s) {
        const int rc = procell_check_proportions(in->types + s * in->n_types, in->n_types);
        if (rc != PROCELL_OK) return rc;
    }
    procell_plan* plan = nullptr;
    int rc = procell_plan_create(in->bin_value, in->bin_freq, in->n_bins, in->phi, &plan);
    if (rc != PROCELL_OK) return rc;
    const size_t n_keys = procell_plan_n_keys(plan), n_rows = procell_plan_n_rows(plan), T = in->n_types;
    procell_sim_params sp;
    memset(&sp, 0, sizeof(sp));
    sp.types = in->types;
    sp.n_types = T;
    sp.n_sets = n_sets;
    sp.t_max = in->t_max;
    sp.seed = in->seed;
    sp.seeding_mode = in->seeding_mode;
    std::vector<int64_t> counts(n_sets * n_keys * T + 1, 0), divisions(n_sets, 0);
    procell_run_stats st;
    memset(&st, 0, sizeof(st));
    if (in->n_gpus == 0 || in->n_gpus == 1) rc = procell_proliferate(plan, &sp, 0, counts.data(), divisions.data(), &st);
    else rc = procell_proliferate_multi(plan, &sp, in->n_gpus, counts.data(), divisions.data(), &st);
    if (rc == PROCELL_OK) {
        out->n_rows = n_rows;
        out->value = static_cast<double*>(malloc((n_rows + 1) * sizeof(double)));
        out->freq = static_cast<int64_t*>(malloc((n_sets * n_rows + 1) * sizeof(int64_t)));
        if (in->track_ratio) out->ratio = static_cast<int64_t*>(malloc((n_sets * n_rows * T + 1) * sizeof(int64_t)));
        if (!out->value || !out->freq || (in->track_ratio && !out->ratio)) {
            procell_output_free(out);
            rc = fail(PROCELL_ERR_ARG, "out of host memory");
        }
    }
    if (rc == PROCELL_OK) {
        procell_plan_export(plan, out->value, nullptr, nullptr, nullptr);
        for (size_t s = 0; s < n_sets && rc == PROCELL_OK; ++s)
            rc = procell_merge_rows(plan, counts.data() + s * n_keys * T, T, out->freq + s * n_rows,
                                    out->ratio ? out->ratio + s * n_rows * T : nullptr);
        for (size_t s = 0; s < n_sets; ++s) out->divisions += divisions[s];
        out->kernel_ms = st.kernel_ms;
    }
    procell_plan_destroy(plan);
    return rc;
}

void procell_output_free(procell_output* out)
{
    if (!out) return;
    free(out->value);
    free(out->freq);
    free(out->ratio);
    out->value = nullptr;
    out->freq = nullptr;
    out->ratio = nullptr;
    out->n_rows = 0;
}

/* all shapes of the RNG-only loop, each timed once after a warm-up: ms[v] and pairs[v] for v = 0..2 */
int procell_rng_ceiling_variants(int device, int iters, double* ms_out, double* pairs_out)
{
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(PROCELL_ERR_CUDA, "no CUDA device available");
    if (!ms_out || !pairs_out) return fail(PROCELL_ERR_ARG, "procell_rng_ceiling_variants: null argument");
    CU(cudaSetDevice(device), "cudaSetDevice");
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device), "query SMs");
    double* d_tab = nullptr;
    unsigned long long* d_sink = nullptr;
    CU(cudaMalloc(&d_tab, sizeof(kLogRows)), "alloc");
    CU(cudaMalloc(&d_sink, 16), "alloc");
    CU(cudaMemcpy(d_tab, &kLogRows, sizeof(kLogRows), cudaMemcpyHostToDevice), "upload");
    CU(cudaMemset(d_sink, 0, 16), "memset");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    SimParams K{};
    set_round_keys(K, 0x0000000200000001ull);
    int rc = PROCELL_OK;
    for (int v = 0; v < kRngCeilingVariants && rc == PROCELL_OK; ++v) {
        const int block = 256, grid = sms * rng_ceiling_ctas_per_sm(v);     /* exactly one full wave */
        const int it = iters / rng_ceiling_chains(v);
        cudaError_t e = launch_rng_ceiling(v, grid, block, 16, d_tab, 48.33, 21.6, 168.0, K.rk, d_sink, nullptr);   /* warm-up */
        if (e == cudaSuccess) {
            cudaEventRecord(e0);
            e = launch_rng_ceiling(v, grid, block, it, d_tab, 48.33, 21.6, 168.0, K.rk, d_sink, nullptr);
            cudaEventRecord(e1);
        }
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { rc = cuda_fail(e, "rng ceiling kernel"); break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        ms_out[v] = ms;
        pairs_out[v] = (double)grid * block * (double)it * rng_ceiling_chains(v);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_tab); cudaFree(d_sink);
    return rc;
}

/* the ceiling the roofline is quoted against: the fastest shape (most pairs per ms) */
int procell_rng_ceiling(int device, int iters, double* ms_out, double* pairs_out)
{
    double ms[kRngCeilingVariants], pairs[kRngCeilingVariants];
    const int rc = procell_rng_ceiling_variants(device, iters, ms, pairs);
    if (rc != PROCELL_OK) return rc;
    int best = 0;
    for (int v = 1; v < kRngCeilingVariants; ++v)
        if (pairs[v] / ms[v] > pairs[best] / ms[best]) best = v;
    if (ms_out) *ms_out = ms[best];
    if (pairs_out) *pairs_out = pairs[best];
    return PROCELL_OK;
}

}  // extern "C"
