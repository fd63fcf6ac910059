/* sim_kernels.h - device-side parameter block and launchers of the sm_100a simulation kernels. */
#ifndef PROCELL_SIM_KERNELS_H
#define PROCELL_SIM_KERNELS_H

#include <cuda_runtime.h>
#include <stdint.h>

namespace procell_b200 {

/* ---- geometry of the warp-cooperative kernel ---- */
constexpr int kCoopWarpsMax = 32;              /* warps per CTA (16, 24 or 32), one CTA per SM */
constexpr int kStackCap = 128;                 /* nodes per warp kept in shared memory (ring), times `ring` (1 or 2) */
constexpr int kChunkNodes = 32;                /* spill / donation granule: one node per lane */
constexpr int kChunkWords = 4 * kChunkNodes;   /* 4 x u64 fields per node, field-major */
constexpr int kSpillCap = 128;                 /* private spill ring, chunks per warp */
constexpr int kQueueCap = 8192;                /* shared donation queue, chunks (power of two) */
constexpr int kLogTabDoubles = 1792;          /* math table in shared memory: 128 log rows {invc, logc} + 256 sin/cos rows + 512 ziggurat rows (procell_spec.h) */
constexpr int kMathTabDoubles = 2816;         /* the table in HBM: the same, then 512 wedge rows that only the rare wedge test reads */
constexpr int kSimpleThreads = 128;
constexpr int kSimpleStack = 136;              /* >= 2*63 + slack entries per thread */

/* device status word */
constexpr int kStatusOk = 0;
constexpr int kStatusSpillOverflow = 1;
constexpr int kStatusQueueTimeout = 2;
constexpr int kStatusIdleTimeout = 3;
constexpr int kStatusWatchdog = 4;
constexpr int kStatusSmemWindow = 5;           /* dynamic shared memory does not start where the kernel was compiled for */
constexpr int kDbgWords = 8;                   /* per-warp debug record written when the watchdog fires */

/* control block in global memory; every hot word on its own 128-byte line */
struct alignas(128) ControlBlock {
    unsigned long long cursor;      unsigned long long pad0[15];
    unsigned long long q_head;      unsigned long long pad1[15];
    unsigned long long q_tail;      unsigned long long pad2[15];
    int active;                     int pad3[31];
    int idle;                       int pad4[31];
    int status;                     int pad5[31];
    int avail;                      int pad7[31];   /* published chunks not yet claimed (permits) */
    /* diagnostics (globaltimer ns): first warp start, seed cursor exhausted, last warp exit; warp-ns spent waiting
     * for donated work (summed over warps) and the number of such waits */
    unsigned long long t_start, t_exhausted, t_end, idle_ns, idle_waits;
    /* warps that hold or may still find work + published chunks nobody has claimed yet.  It is the ONE word the
     * termination test reads: a chunk is added to it before its permit is published and a warp leaves it only when it
     * goes idle, so it can reach 0 only when no work exists anywhere, and then stays 0 */
    long long pending;
    unsigned long long done_ctas;   /* CTAs whose tables are flushed: the grid-wide rendezvous in front of the fused sweep fitness */
    unsigned long long pad6[9];
};

struct SimParams {
    /* histogram plan, resident in HBM */
    const uint32_t* bin_start;    /* [n_bins+1] first seed-cell id of each bin */
    const uint32_t* bin_keybase;  /* [n_bins]   first key of the bin */
    const uint8_t* bin_kdiv;      /* [n_bins]   halvings allowed (<=63) | 0x80 if level-0 leaves are countable */
    /* type tables, [n_sets][n_types] */
    const uint32_t* type_thr;     /* [n_sets][n_types rounded up to 4], selection order (descending proportion): the last
                                     32-bit type word x that lies below the running proportion sum - type j is the first
                                     with x <= type_thr[j]; the last type's slot and the padding hold 2^32 - 1 */
    const uint8_t* type_sel;      /* file id of the j-th type in selection order */
    const double2* type_musd;     /* (mean, sd) by file id */
    const double* logtab;         /* the math table: 128 x {invc, logc}, 256 x {sin, cos}, 512 x {x_i, x_i+1}, 512 x wedge rows */
    /* outputs */
    long long* counts;            /* [n_sets][n_keys][n_types] */
    long long* divisions;         /* [n_sets] */
    /* work distribution */
    ControlBlock* ctl;
    unsigned long long* q_seq;    /* [kQueueCap] slot sequence numbers (Vyukov bounded queue) */
    unsigned long long* q_data;   /* [kQueueCap][kChunkWords] */
    unsigned long long* spill;    /* [grid*warps][kSpillCap][kChunkWords] */
    uint32_t n_bins, n_types, n_sets, n_keys;
    uint32_t n_cells;
    uint32_t unit;                /* seed cells per claim unit */
    uint32_t units_per_set;       /* ceil(n_cells / unit) */
    uint32_t local_units_per_set; /* units with u % world == rank */
    uint32_t shard_world, shard_rank;
    unsigned long long total_local_units;   /* n_sets * local_units_per_set */
    /* parameter sweeps (n_sets > 1): a CTA claims a BATCH of consecutive units of one set from the global cursor and
     * its warps take units from it through shared memory, so that a CTA's histogram cache sees one set at a time */
    uint32_t batch_units;         /* units per batch (< 2^24) */
    uint32_t batches_per_set;
    unsigned long long total_batches;
    uint32_t rk[20];              /* Philox round keys: rk[2r] = seed_lo + r*W0, rk[2r+1] = seed_hi + r*W1 */
    uint32_t smem_hist_slots;     /* direct mode: keys below this are privatised in shared memory (u32 each);
                                     hashed mode: number of {key,count} u64 slots, a power of two */
    int hist_hashed;              /* 1: key space larger than shared memory -> direct-mapped {key,count} cache */
    int refcompat;
    int donate;                   /* 0: never hand work to starving warps (diagnostic) */
    unsigned long long watchdog_ns;   /* a warp that runs longer than this aborts the launch (status 4) */
    unsigned long long* dbg;      /* [grid*warps][kDbgWords] */
    double t_max;                 /* = times[n_times - 1] */
    /* time series: histograms at several checkpoints from ONE tree expansion.  A cell is counted at
     * checkpoint j iff it exists and is out of time there, i.e. birth <= times[j] < division time - exactly what a
     * run with t_max = times[j] would count with the same random stream.  counts: [n_times][n_sets][n_keys][n_types] */
    uint32_t n_times;             /* 1..8 */
    uint32_t time_stride;         /* n_sets * n_keys * n_types */
    double times[8];              /* ascending */
    /* subtree sharding (deep trees): appended LAST so that the parameter-block offsets of everything above -
     * and with them the machine code of the other kernel instances - stay as they are.  sub_world > 1: this GPU builds
     * EVERY seed cell and expands every node whose heap index is below sub_limit = 2^L (tree level < L); what those
     * nodes count is credited to GPU root % sub_world; a daughter at level L that will divide is kept by GPU
     * (root + heap) % sub_world alone.  shard_world / shard_rank are 1 / 0 in this mode. */
    uint32_t sub_world, sub_rank;
    unsigned long long sub_limit;
    /* sweeps: 1 = the shared-memory table is a direct u32 table of ONE parameter set (smem_hist_slots = n_keys *
     * n_types), re-based at batch switches behind a CTA-wide rendezvous (kernel MODE kModeSetDirect) */
    int hist_setdirect;
    /* slot layout of the shared-memory count table of the PLAIN direct instances (one parameter set, one checkpoint,
     * u32 table).  The count tensor is [key][type], but a quiescent type is only ever counted at tree level 0 of a bin,
     * so the TABLE keeps the proliferating types only, [key][n_prolif], followed by one row per bin for the quiescent
     * ones, [bin][n_quiet]: n_keys * n_prolif + n_bins * n_quiet slots instead of n_keys * n_types.  That is what lets
     * deep-tree runs (config 4: 10 422 keys x 3 types = 31 266 counters, one type quiescent) keep the direct table
     * instead of the hashed cache.  A node carries its SLOT (stride kstride = n_prolif per tree level); slots are
     * translated to tensor indices only when the table is drained (slot_key in sim_kernels.cu). */
    uint32_t kstride;             /* stride of the node's key field per tree level: n_prolif in slot mode, else n_types */
    uint32_t slot_mode;           /* 1: the table is laid out by slots (PLAIN direct instances only) */
    uint32_t slot_prolif_end;     /* n_keys * n_prolif: first slot of the quiescent rows */
    uint32_t n_prolif, n_quiet;
    unsigned long long quiet_mask;    /* bit j: type j (file id) is quiescent; a type's rank among its kind is a popcount below j */
    const uint8_t* rank_type;     /* [n_prolif] file ids of the proliferating types, then [n_quiet] of the quiescent ones (drains only) */
    /* fitness of a sweep in the same launch: when fit_channels > 0 the CTAs meet once all tables are
     * flushed and share the parameter sets among them; each set's slab is re-binned onto the target's channels straight
     * from L2 and its Hellinger distance written to fit_out[set] (fitness_device.h) */
    uint32_t fit_channels;
    const uint32_t* fit_key_channel;   /* [n_keys] key -> target channel, 0xFFFFFFFF = not an output row */
    const double* fit_target;          /* [fit_channels] target shares */
    double* fit_out;                   /* [n_sets] */
    /* 1: deep lineage trees (capi.cu: expected depth from t_max and the halvings phi allows) - the PLAIN direct instance that
     * merges equal leaf keys of a DIVIDE iteration before the shared-memory atomic (kernel MODE kModeMerge).  Host-side
     * choice of the instance only; last in the block so that the other instances' parameter offsets stay where they were */
    int leaf_merge;
};

/* ring = 1: 128-node ring per warp, one node per lane and iteration (warps = 32, 24 or 16);
 * ring = 2: 256-node ring per warp, two nodes per lane in the full DIVIDE iteration (warps = 16 only) */
size_t coop_smem_bytes(int warps, int ring, uint32_t hist_slots, int hashed);
cudaError_t launch_coop(const SimParams& p, int warps, int ring, int grid, cudaStream_t stream);
cudaError_t coop_max_grid(int device, int warps, int ring, int hashed, int plain, size_t smem_bytes, int* grid_out);
/* the subtree-sharding instances (p.sub_world > 1): 32 warps, 128-node rings, one parameter set, one checkpoint;
 * launch_coop picks them by p.sub_world */
cudaError_t coop_max_grid_subtree(int device, int hashed, size_t smem_bytes, int* grid_out);
/* the sweep instance with a set-relative direct table (p.hist_setdirect): 32 warps, 128-node rings */
cudaError_t coop_max_grid_setdirect(int device, size_t smem_bytes, int* grid_out);
/* the PLAIN direct instance with merged leaf counts (p.leaf_merge): 32 warps, 128-node rings */
cudaError_t coop_max_grid_merge(int device, size_t smem_bytes, int* grid_out);
cudaError_t launch_simple(const SimParams& p, int grid, cudaStream_t stream);
/* resets the queue and the control block (the status word excepted: it is sticky until the host has read it) and
 * zeroes the count tensor and the division counters of the run - one launch instead of a kernel and two memsets */
cudaError_t launch_queue_init(unsigned long long* q_seq, ControlBlock* ctl, long long* counts, size_t n_counts,
                              long long* divisions, size_t n_divisions, int sm_count, cudaStream_t stream);
/* variant 0..2: see k_rng_ceiling in sim_kernels.cu */
constexpr int kRngCeilingVariants = 3;
int rng_ceiling_ctas_per_sm(int variant);
int rng_ceiling_chains(int variant);
cudaError_t launch_rng_ceiling(int variant, int grid, int block, int iters, const double* logtab, double mean, double sd,
                               double t_max, const uint32_t* rk, unsigned long long* sink,
                               cudaStream_t stream);

cudaError_t launch_sweep_fitness(const long long* counts, const uint32_t* key_channel, const double* target_share,
                                 uint32_t n_sets, uint32_t n_keys, uint32_t n_types, uint32_t n_channels, double* out,
                                 cudaStream_t stream);

}  // namespace procell_b200
#endif
