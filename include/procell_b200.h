/* procell_b200.h - C ABI of the B200-native ProCell proliferation simulator (libprocell_b200.so).
 *
 * The reference (ericniso/cuda-pro-cell) has no FFI: its hot path sits behind three C++ host functions
 * that simulation::Simulator calls, plus the text parser/writer.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference root).  Plain pointers and sizes only;
 * nothing throws or exit()s across this boundary; every function returns PROCELL_OK (0) or a negative
 * PROCELL_ERR_* code and procell_last_error() gives the message of the calling thread's last failure.
 * There is NO CPU fallback: anything that simulates needs a CUDA device of compute capability 10.x.
 */
#ifndef PROCELL_B200_H
#define PROCELL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PROCELL_OK 0
#define PROCELL_ERR_ARG (-1)         /* bad argument / limit of the key layout exceeded */
#define PROCELL_ERR_CUDA (-2)        /* CUDA runtime error (no device, launch failure, ...) */
#define PROCELL_ERR_PROPORTION (-3)  /* |1 - sum(proportion)| > 1e-8 (src/io/parser.cu:46-66) */
#define PROCELL_ERR_IO (-4)          /* file could not be opened / written */
#define PROCELL_ERR_OVERFLOW (-5)    /* device work pool overflow or watchdog abort; results invalid */

#define PROCELL_SEEDING_IDEAL 0      /* all draws independent */
#define PROCELL_SEEDING_REFCOMPAT 1  /* seed cell's type uniform doubles as the Box-Muller radius uniform of
                                        its first timer, as the reference's one-seed-many-draws does
                                        (src/simulation/cell.cu:38,55 + src/utils/util.cu:153-169) */

#define PROCELL_KERNEL_COOP 0        /* warp-cooperative depth-first kernel (default) */
#define PROCELL_KERNEL_SIMPLE 1      /* one thread per lineage, global atomics (bring-up / cross-check) */

/* one subpopulation: src/simulation/data_types.h:21-27 (cell_type) minus `name` (= array index) */
typedef struct procell_cell_type {
    double proportion;
    double mean;     /* < 0: quiescent (README "-1 -1") */
    double stddev;
} procell_cell_type;

typedef struct procell_plan procell_plan;     /* host: bins, key space, merged output rows */
typedef struct procell_engine procell_engine; /* one GPU: resident tables, work pool, count tensor */

const char* procell_last_error(void);
const char* procell_version(void);

/* ---- text I/O (streaming: one read() + in-place scan, one buffer per write(); csrc/textio.cpp) ---- */
/* replaces io::load_fluorescences' reading loop (src/io/parser.cu:103-106): "<double> <uint64>" pairs
 * until the first parse failure; lines with frequency 0 are KEPT here (the plan skips them).
 * Arrays are malloc'd; release with procell_free. */
int procell_read_histogram(const char* path, double** value, uint64_t** freq, size_t* n_lines);
/* the same two readers on text already in memory (len bytes, no terminator needed) */
int procell_parse_histogram(const char* text, size_t len, double** value, uint64_t** freq, size_t* n_lines);
int procell_parse_cell_types(const char* text, size_t len, procell_cell_type** types, size_t* n_types);
/* replaces io::load_cell_types (src/io/parser.cu:156-185): "<proportion> <mean> <stddev>" triples, file
 * order kept (type id = line index); checks the proportion sum. */
int procell_read_cell_types(const char* path, procell_cell_type** types, size_t* n_types);
/* replaces io::save_fluorescences (src/io/parser.cu:187-217): rows with frequency > 0, ascending value,
 * "%.10g" TAB frequency [TAB per-type counts in file order]; path == NULL writes to stdout. */
int procell_write_histogram(const char* path, int save_ratio, size_t n_types, size_t n_rows,
                            const double* row_value, const int64_t* row_freq, const int64_t* row_ratio);
void procell_free(void* p);
int procell_check_proportions(const procell_cell_type* types, size_t n_types);
/* number of 32-bit words x whose seed-cell type uniform (2x + 1) / 2^33 lies below the cumulative proportion `cum`
 * (cell.cu:81-104 compares the uniform with the running sums): the integer threshold the kernels scan instead of the
 * doubles; exact, clamped to [0, 2^32].  Exposed for tests. */
uint64_t procell_type_threshold(double cum);

/* ---- plan: the result-key precomputation of io::load_fluorescences (src/io/parser.cu:68-154) ---- */
/* phi == 0 selects the default (smallest value with frequency > 0, parser.cu:80-96). */
int procell_plan_create(const double* value, const uint64_t* freq, size_t n_lines, double phi,
                        procell_plan** out);
void procell_plan_destroy(procell_plan* plan);
size_t procell_plan_n_bins(const procell_plan* plan);   /* lines with frequency > 0 */
size_t procell_plan_n_keys(const procell_plan* plan);   /* (bin, k) pairs, k = 0..kdiv(bin) */
size_t procell_plan_n_rows(const procell_plan* plan);   /* distinct values value/2^k >= phi, ascending */
uint64_t procell_plan_n_cells(const procell_plan* plan);
double procell_plan_phi(const procell_plan* plan);
int procell_plan_depth_capped(const procell_plan* plan);
/* Expected depth of a lineage tree (diagnostic; no GPU needed): the smaller of the generations t_max leaves room for -
 * t_max / the smallest positive mean among `types` - and the halvings phi allows, averaged over the seed cells.  The
 * library uses it to choose the kernel instance for deep trees (procell_engine_kernel_mode). */
double procell_plan_lineage_depth(const procell_plan* plan, const procell_cell_type* types, size_t n_types, double t_max);
/* any pointer may be NULL; sizes: row_value[n_rows], key_row[n_keys], bin_keybase[n_bins], bin_kdiv[n_bins] */
int procell_plan_export(const procell_plan* plan, double* row_value, uint32_t* key_row,
                        uint32_t* bin_keybase, uint8_t* bin_kdiv);
/* counts of ONE parameter set [n_keys][n_types] -> rows: row_freq[n_rows], row_ratio[n_rows][n_types] */
int procell_merge_rows(const procell_plan* plan, const int64_t* counts, size_t n_types,
                       int64_t* row_freq, int64_t* row_ratio);

/* ---- simulation ------------------------------------------------------------------------------- */
typedef struct procell_sim_params {
    const procell_cell_type* types; /* [n_sets][n_types], file order */
    size_t n_types;                 /* 1..64 */
    size_t n_sets;                  /* 1..65536 parameter sets simulated in ONE launch on the same histogram */
    double t_max;
    uint64_t seed;                  /* Philox key */
    int seeding_mode;               /* PROCELL_SEEDING_* */
    int kernel;                     /* PROCELL_KERNEL_* */
    uint32_t shard_rank;            /* this GPU simulates the seed-cell units u with u % shard_world == rank */
    uint32_t shard_world;           /* 0 or 1: everything */
    uint32_t shard_unit;            /* seed cells per unit; 0: default */
    /* time series (extension): histograms at up to 8 ascending checkpoints from ONE tree expansion; the last one
     * replaces t_max.  NULL / 0: the single checkpoint t_max.  With m checkpoints every count tensor is
     * [m][n_sets][n_keys][n_types]; slice j equals a run with t_max = checkpoints[j] and the same seed, bit for bit. */
    const double* checkpoints;
    size_t n_checkpoints;
    /* subtree sharding (extension; multi-GPU runs of DEEP trees, where a handful of lineages own all the work and
     * sharding whole lineages leaves the GPUs unevenly loaded): 0 = off, the rule above.  L >= 1 with shard_world > 1:
     * every GPU builds every seed cell and expands every node of tree level < L (cheap: at most 2^L nodes per
     * lineage); what those nodes count is credited to GPU root % shard_world; a daughter at level L that will divide
     * is kept by GPU (root + heap index) % shard_world alone.  The GPUs' tensors still sum to the single-GPU result
     * bit for bit.  Needs the cooperative kernel, one parameter set, one checkpoint, L <= 30; shard_unit is ignored. */
    uint32_t shard_level;
} procell_sim_params;

typedef struct procell_run_stats {
    int64_t divisions;   /* total over sets, this shard */
    double kernel_ms;    /* CUDA-event time of memsets + kernel on the engine's stream (host-API runs) */
    int n_launches;      /* kernels launched by the run */
    int grid, block;     /* launch shape of the simulation kernel */
    int smem_bytes;
    int64_t donations;   /* 32-node chunks handed from busy to starving warps through the device queue */
    double seed_phase_us; /* device time from the first warp's start until the seed-unit cursor ran out (-1: n/a) */
    double total_us;      /* device time from the first warp's start to the last warp's exit (-1: n/a) */
    double idle_warp_us;  /* time warps spent waiting for donated work, summed over all warps of the launch */
    int64_t idle_waits;   /* how often a warp ran dry and waited */
} procell_run_stats;

/* One-shot, host buffers in and out: replaces simulation::create_cells_population
 * (src/simulation/cells_population.h:12-18) + simulation::proliferate (src/simulation/proliferation.h:12-19);
 * seed cells are never materialised.  counts: [n_sets][n_keys][n_types] int64; divisions: [n_sets] or NULL. */
int procell_proliferate(const procell_plan* plan, const procell_sim_params* params, int device,
                        int64_t* counts, int64_t* divisions, procell_run_stats* stats);

/* Same, on the first n_gpus GPUs of this box from ONE process (n_gpus <= 0: all): seed-cell units (shard_unit cells;
 * 0 = chosen from the input size, 1 to 256) are
 * sharded GPU-strided, every GPU runs the same kernel on its units, and one ncclReduce(sum, int64) over NVLink
 * combines the count tensors on GPU 0.  The result equals the single-GPU result bit for bit.  The reference is
 * single-GPU (device 0 hard-coded, src/simulation/proliferation.cu:38).  NCCL is loaded with dlopen. */
int procell_proliferate_multi(const procell_plan* plan, const procell_sim_params* params, int n_gpus,
                              int64_t* counts, int64_t* divisions, procell_run_stats* stats);

/* One call from histogram arrays to result rows (the entry point SURVEY section 8b proposes): plan + simulation on
 * n_gpus GPUs + merge.  Replaces Simulator::load_params .. save_results minus the file I/O
 * (src/simulation/simulator.cu:10-56).  The library allocates the output arrays; release them with
 * procell_output_free.  Never exits, no exceptions cross the ABI; no global state besides the thread-local error. */
typedef struct procell_input {
    const double* bin_value;        /* histogram lines, file order (zero frequencies allowed) */
    const uint64_t* bin_freq;
    size_t n_bins;
    const procell_cell_type* types; /* [n_param_sets][n_types] */
    size_t n_types;
    size_t n_param_sets;            /* 0 is read as 1 */
    double t_max;
    double phi;                     /* 0: default (smallest value with a non-zero frequency) */
    int track_ratio;                /* fill `ratio` */
    uint64_t seed;
    int n_gpus;                     /* 0 or 1: device 0; > 1: the first n_gpus GPUs, one ncclReduce; < 0: all */
    int seeding_mode;               /* PROCELL_SEEDING_* */
} procell_input;

typedef struct procell_output {
    double* value;                  /* [n_rows] ascending */
    int64_t* freq;                  /* [n_param_sets][n_rows] (rows with 0 are not printed by the writer) */
    int64_t* ratio;                 /* [n_param_sets][n_rows][n_types], NULL unless track_ratio */
    size_t n_rows;
    int64_t divisions;              /* total over the parameter sets */
    double kernel_ms;
} procell_output;

int procell_simulate(const procell_input* in, procell_output* out);
void procell_output_free(procell_output* out);

/* Resident engine: tables live in HBM across runs. */
int procell_engine_create(int device, procell_engine** out);
void procell_engine_destroy(procell_engine* engine);
/* host -> device upload of the plan and type tables (the H2D leg of an end-to-end step) */
int procell_engine_load(procell_engine* engine, const procell_plan* plan, const procell_sim_params* params);
/* Enqueue one simulation on `stream` (a cudaStream_t, NULL = default stream): one launch resets the work pool and
 * zeroes the count tensor and the division counters, one runs the simulation.  d_counts / d_divisions are DEVICE
 * pointers ([n_sets][n_keys][n_types] / [n_sets] int64) or NULL to use the engine's own tensors.  Asynchronous; the
 * tables of the last procell_engine_load are uploaded asynchronously too and the run is ordered behind them. */
int procell_engine_run(procell_engine* engine, uint64_t seed, void* stream, int64_t* d_counts,
                       int64_t* d_divisions);
/* wait for the stream, check the device status word, copy the results of the LAST run - from the engine's own tensors
 * or from the device buffers that run was given - to host buffers (counts / divisions may be NULL).  The status word is
 * sticky on the device: a failure of any run queued since the previous finish is reported here (PROCELL_ERR_OVERFLOW)
 * and then cleared. */
int procell_engine_finish(procell_engine* engine, void* stream, int64_t* counts, int64_t* divisions,
                          procell_run_stats* stats);
size_t procell_engine_counts_len(const procell_engine* engine); /* n_sets*n_keys*n_types */

/* On-GPU fitness of a sweep (what a ProCell fitting loop consumes): re-bin every set's simulated histogram onto the
 * target's channels - first channel whose value is >= the row value, the rule of the reference's unused
 * utils::rebin (src/utils/util.cu:111-138) - and return the Hellinger distance sqrt(1 - sum sqrt(p q)) per set.
 * set_target uploads the target (values strictly ascending); fitness runs after procell_engine_run on the same
 * stream, on the engine's own count tensor (d_counts NULL) or on the caller's device tensor. fitness: host [n_sets]. */
int procell_engine_set_target(procell_engine* engine, const double* value, const uint64_t* freq, size_t n_channels);
int procell_engine_fitness(procell_engine* engine, void* stream, const int64_t* d_counts, double* fitness);
/* When the target was set BEFORE procell_engine_run and the run was an unsharded sweep (n_sets > 1 or checkpoints) on
 * the cooperative kernel, the simulation kernel computes the distances itself at the end of the same launch - the CTAs
 * meet once their tables are flushed and re-bin the slabs straight from L2 - and procell_engine_fitness only downloads
 * 8 bytes per set.  Same arithmetic, same bits as the separate pass (csrc/fitness_device.h).  Returns 1 if the last
 * procell_engine_fitness call was served that way, 0 if it ran the separate kernel (PROCELL_FITNESS_FUSED=0 forces 0). */
int procell_engine_fitness_in_launch(const procell_engine* engine);
/* Which instance of the simulation kernel the loaded simulation runs on (diagnostic; results do not depend on it):
 * 0 base, 1 subtree sharding, 2 sweep with a set-relative count table, 3 deep lineage trees - equal leaf keys of an
 * iteration merged before the shared-memory atomic (chosen at load time from t_max / fastest mean and the halvings phi
 * allows; PROCELL_LEAF_MERGE=0/1 forces it) -, -1 the bring-up kernel or nothing loaded. */
int procell_engine_kernel_mode(const procell_engine* engine);

/* RNG-only micro-kernel: per thread `iters` Philox blocks + two fast ziggurat tests + timers into a register
 * accumulator (the instruction-issue ceiling the roofline fraction is quoted against).  Returns ms. */
int procell_rng_ceiling(int device, int iters, double* ms_out, double* pairs_out);
/* The loop exists in three shapes (48 warps per SM; 64 warps per SM at 32 registers; two independent chains per thread):
 * procell_rng_ceiling reports the FASTEST, this one all of them - ms_out[3], pairs_out[3]. */
int procell_rng_ceiling_variants(int device, int iters, double* ms_out, double* pairs_out);

/* ---- the `procell` command line (src/main.cu:18-34 + src/io/cmdargs.cpp:11-76) ---------------- */
/* Returns the process exit code; prints errors to stdout as the reference does. */
int procell_main(int argc, char** argv);

#ifdef __cplusplus
}
#endif
#endif /* PROCELL_B200_H */
