/* sim_kernels.cu - hand-written sm_100a kernels of the proliferation simulator.
 *
 * Replaces, in the reference (ericniso/cuda-pro-cell):
 *   device::create_cells_from_fluorescence / create_cells_population  (src/simulation/cells_population.cu:74-117)
 *   device::create_cell and helpers                                   (src/simulation/cell.cu:25-143)
 *   device::proliferate + the host level loop                         (src/simulation/proliferation.cu:26-402)
 * The reference materialises every tree level densely in HBM (N*2^d 40-byte cells) and recurses by one
 * dynamic-parallelism launch per block per level.  Here nothing but the count tensor ever reaches HBM:
 *
 * k_proliferate_coop - persistent, one CTA per SM, 32 warps (16 / 24 as tuning shapes).  Each warp owns a 128-node
 *   ring in shared memory.  An iteration is warp-uniform: either SEED (32 lanes build 32 seed cells: bin, type,
 *   first timer, initial age; living roots are pushed) or DIVIDE (32 lanes pop the 32 newest = deepest nodes, each
 *   draws ONE Philox block and makes the fast test of the ziggurat method for both daughters' normals - a table row, a
 *   multiplication and a compare each -, classifies the daughters as leaf / dropped / internal and pushes the internal
 *   ones with predicated 16-byte stores; a daughter whose draw needs more - the ziggurat's wedge or tail, a redraw after
 *   a non-positive timer: 2-3 % of the divisions - goes to the BOTTOM of the ring as a retry node, and 32 of those are
 *   expanded together by a general iteration).  While a warp holds fewer than
 *   32 nodes every node is expanded each iteration (the breadth-first warm-up); from 32 on, expansion is depth-first,
 *   and the ring overflows in 32-node chunks into a private spill ring in HBM.  Starving warps are fed through a
 *   bounded MPMC queue of chunks that busy warps fill from the BOTTOM of their stacks (the shallowest nodes = the
 *   largest subtrees).  Leaves are counted in a shared-memory histogram keyed (bin, k, type): a u32 table with one
 *   fire-and-forget atomic per lane when the key space fits (the slots cannot wrap: every warp drains the CTA's
 *   table into the int64 tensor in HBM every 2^20 of its own DIVIDE iterations, see kHistFlushIters), else a
 *   direct-mapped {key, count} cache updated after __match_any_sync merging; the rest is flushed at the end.
 *
 *   Three further compile-time modes of the same kernel: MODE 1 = subtree sharding for multi-GPU runs of deep trees,
 *   MODE 2 = sweeps with a direct table of the CTA's current parameter set and a CTA-wide rendezvous at batch switches
 *   (config 5: 70.4 -> 60.7 ms in round 1), MODE 3 = deep trees on one parameter set: equal leaf keys of an iteration are
 *   merged before the shared-memory atomic (config 4: 612 -> 597 ms).  Every instance is bit-exact against the oracle on a
 *   B200 (tests/test_gpu_parity.py).
 *
 * k_proliferate_simple - one thread per lineage with a local-memory stack and global atomics: the bring-up
 *   kernel, kept as an independent device-side cross-check of the cooperative one.
 */
#include "sim_kernels.h"

#include <cstddef>

#include "procell_spec.h"
#include "fitness_device.h"

namespace procell_b200 {

static_assert(kLogTabDoubles == PCS_TAB_DOUBLES && kMathTabDoubles == PCS_TAB_ALL_DOUBLES, "math table size");

namespace {

#ifdef PROCELL_TRACE
#define TRACE(P, gw, lane, code) do { if ((lane) == 0) __stcg((P).dbg + (size_t)(gw) * kDbgWords + 1, (unsigned long long)(code)); } while (0)
#else
#define TRACE(P, gw, lane, code) do { } while (0)
#endif

/* ring capacity of a warp: kStackCap nodes per node-per-lane of the instance (RING = 1: 128 nodes, the product shape
 * with 32 warps; RING = 2: 256 nodes for the 16-warp instance whose full DIVIDE iteration expands two nodes per lane) */
template <int RING> struct Ring {
    static constexpr uint32_t kCap = (uint32_t)kStackCap * RING;
    static constexpr uint32_t kMask = kCap - 1u;
};
constexpr int kSmemMusdEntries = 64;  /* (mean, sd) table cached in shared memory when n_sets * n_types fits */
/* control words (64 B) + three 16-byte snapshots of the control block, then the type tables of small runs:
 * (mean, sd) 16 B, type threshold 4 B (+ 4 B: the per-warp donation epochs live in that half) and selection order 1 B per entry */
constexpr int kSmemCtlBytes = 128 + kSmemMusdEntries * (16 + 8 + 1);
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kSlot = 16u;                  /* ring positions count bytes of the (A, B) half: 16 per node (WarpCtx) */
constexpr uint32_t kDynSmemWindow = 0x400u;      /* where dynamic shared memory starts in the shared-memory window on sm_100 */
/* kernel MODE: 0 = the kernel as measured in round 1; 1 = subtree sharding compiled in (multi-GPU runs of deep trees);
 * 2 = sweeps with a set-relative direct histogram table and a CTA-wide rendezvous at batch switches; 3 = deep trees on one
 * parameter set: equal leaf keys of a DIVIDE iteration are merged before the shared-memory atomic */
constexpr int kModeBase = 0, kModeSubtree = 1, kModeSetDirect = 2, kModeMerge = 3;

/* Shared memory of k_proliferate_coop: (mean, sd) table 1 KB | rings | math tables | control words + threshold / epoch /
 * selection tables | count table.  The *Addr members are addresses in the shared-memory WINDOW (what ld.shared takes):
 * compile-time constants, so the table accesses of the hot loop are `LDS [register + immediate]` - see the window check at
 * the top of the kernel. */
template <int WARPS, int RING> struct SmemLayout {
    static constexpr uint32_t kRingsOff = (uint32_t)kSmemMusdEntries * 16u;
    static constexpr uint32_t kRingsBytes = (uint32_t)WARPS * 4u * Ring<RING>::kCap * 8u;
    static constexpr uint32_t kTabOff = kRingsOff + kRingsBytes;
    static constexpr uint32_t kCtlOff = kTabOff + (uint32_t)kLogTabDoubles * 8u;
    static constexpr uint32_t kMusdAddr = kDynSmemWindow;
    static constexpr uint32_t kZigAddr = kDynSmemWindow + kTabOff + (uint32_t)PCS_TAB_ZIG * 8u;
};

/* pcs_zig_fast (procell_spec.h) with the table row read from the shared-memory window address ZIG_ADDR + 16 * layer (the
 * asm keeps the row offset a register and the table base an immediate of the load), and with the sign put on the
 * multiplicand instead of OR-ed onto the product: (-M) * xs = -(M * xs) bit for bit (round-to-nearest is symmetric, a signed
 * zero included), so *z_out is the same double as pcs_zig_fast's, the test is |x| < x_next, and the draw's high word is not
 * needed again after the multiplication (tests: every GPU parity test compares against the oracle's pcs_zig_fast). */
template <uint32_t ZIG_ADDR>
__device__ __forceinline__ bool zig_fast_smem(uint32_t lo, uint32_t hi, double* z_out)
{
    static_assert(31 - PCM_ZIG_N_BITS >= 4, "layer field sits high enough to become a byte offset by one shift");
    const uint32_t off = (hi >> (31 - PCM_ZIG_N_BITS - 4)) & (((1u << PCM_ZIG_N_BITS) - 1u) << 4);      /* 16 * PCS_ZIG_LAYER(hi) */
    double xs, xn;
    asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(xs), "=d"(xn) : "r"(off), "n"(ZIG_ADDR));
    const double m = pcs_bits2d(((uint64_t)(hi & 0x801FFFFFu) << 32) | (uint64_t)lo);     /* +-M * 2^-1074 */
    const double x = PCS_MUL(m, xs);
    *z_out = x;
    return fabs(x) < xn;
}

/* the same with the table behind an ordinary pointer (the ceiling kernels' own static shared-memory copy) */
__device__ __forceinline__ bool zig_fast_signed(uint32_t lo, uint32_t hi, const double* tab_zig, double* z_out)
{
    const double2 row = *reinterpret_cast<const double2*>(tab_zig + 2u * PCS_ZIG_LAYER(hi));
    const double x = PCS_MUL(pcs_bits2d(((uint64_t)(hi & 0x801FFFFFu) << 32) | (uint64_t)lo), row.x);
    *z_out = x;
    return fabs(x) < row.y;
}

/* (mean, sd) of type `type` of the ONE parameter set of a PLAIN instance, from the table at the start of shared memory;
 * dlo = the node's D word (type in bits 16..21) */
template <uint32_t MUSD_ADDR>
__device__ __forceinline__ double2 musd_smem(uint32_t dlo)
{
    const uint32_t off = (dlo >> 12) & (63u << 4);
    double2 ms;
    asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(ms.x), "=d"(ms.y) : "r"(off), "n"(MUSD_ADDR));
    return ms;
}
/* A node = a cell that WILL divide: 4 x u64, kept in the ring as two 16-byte pairs (A,B) and (C,D) so that a pop is
 * two LDS.128 and a push two STS.128; chunks in HBM (spill rings, donation queue) are field-major
 *   A  t_div   time of its division = birth time of its daughters (double bits)
 *   B  heap    tree path: root = 1, daughters 2h and 2h+1 (Philox counter words 2,3)
 *   C  root cell id | key << 32, key = count-tensor index of (set, bin, level of THIS node, type)
 *   D  lo: set[0:16) | type[16:22) | rem[22:28) | pending-daughter mask[28:30); hi: retry[0:8)
 *      rem = halvings still allowed below this node's daughters (daughters divide iff rem > 0) */
__device__ __forceinline__ uint32_t pack_dlo(uint32_t set, uint32_t type, uint32_t rem, uint32_t mask)
{
    return set | (type << 16) | (rem << 22) | (mask << 28);
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ int ld_acquire_s32(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ int ld_volatile_s32(const int* p)
{
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

/* 16-byte asynchronous copy global -> shared that bypasses L1 (the source is written by other SMs) */
__device__ __forceinline__ void cp_async16(volatile int* smem_dst, const void* gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(const_cast<int*>(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}

/* the two one-word flags of a CTA that several warps read while another may set them - s_ctl[1] "quiescent", s_ctl[3]
 * "the seed cursor has run out" (both only ever go 0 -> 1) - and the rendezvous round counter are read and written
 * with shared-memory atomics: that states the intent to the hardware, and to compute-sanitizer's racecheck, which
 * otherwise (rightly) reports every plain read of a word another warp writes.  Lane 0 does the reading (RULE below). */
__device__ __forceinline__ int sflag_get(volatile int* p) { return atomicOr(const_cast<int*>(p), 0); }
__device__ __forceinline__ void sflag_set(volatile int* p, int v) { atomicExch(const_cast<int*>(p), v); }

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

/* shared-memory privatised count.
 * direct mode (HASHED = false): the whole key space fits: one u32 per key, updated with a fire-and-forget shared
 *   atomic (no return value, so no scoreboard wait).  The slots cannot wrap because of hist_drain below.
 * hashed mode (key space larger than shared memory: parameter sweeps, very deep histograms): a direct-mapped
 *   cache of {key:32 | count:32} words indexed by the low key bits.  Keys of one parameter set are contiguous, so
 *   they never collide with each other; a colliding key evicts the resident one, whose count goes to HBM. */
template <bool HASHED>
__device__ __forceinline__ void hist_add(const SimParams& P, uint32_t* s_hist, uint32_t key, uint32_t v)
{
    if (HASHED) {
        unsigned long long* slot = reinterpret_cast<unsigned long long*>(s_hist) + (key & (P.smem_hist_slots - 1u));
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(slot);
        for (;;) {
            const uint32_t tag = (uint32_t)(cur >> 32), cnt = (uint32_t)cur;
            const bool same = tag == key && cnt < 0xF0000000u;
            const unsigned long long want = same ? cur + v : (((unsigned long long)key << 32) | v);
            const unsigned long long old = atomicCAS(slot, cur, want);
            if (old == cur) {
                if (!same && cnt) atomicAdd(reinterpret_cast<unsigned long long*>(P.counts) + tag, (unsigned long long)cnt);
                return;
            }
            cur = old;
        }
    } else {
        atomicAdd(&s_hist[key], v);
    }
}

/* Wrap protection of the direct-mode u32 slots.  A lane adds at most 2 leaves per key and DIVIDE iteration, and a key
 * of tree level 0 (the only kind SEED iterations touch) receives at most freq[bin] < 2^32 seed leaves in a whole run.
 * Every warp drains ALL slots of its CTA (atomicExch -> int64 atomicAdd in HBM) each kHistFlushIters of its own
 * DIVIDE iterations.  Between two consecutive drains of a slot no warp can have run more than kHistFlushIters
 * iterations (its own drain would have been one of them), so the slot received < 1024 lanes * 2 * 2^20 = 2^31. */
#ifndef PROCELL_HIST_FLUSH_ITERS
#define PROCELL_HIST_FLUSH_ITERS (1u << 20)
#endif
constexpr uint32_t kHistFlushIters = PROCELL_HIST_FLUSH_ITERS;     /* power of two, multiple of 256 */

/* work-donation policy knobs (A/B'd on the GPU) */
#ifndef PROCELL_IDLE_BACKOFF_MAX_NS
#define PROCELL_IDLE_BACKOFF_MAX_NS 512u
#endif
#ifndef PROCELL_DONATE_MIN_NODES
#define PROCELL_DONATE_MIN_NODES 96u
#endif
#ifndef PROCELL_DONATE_RESERVE
#define PROCELL_DONATE_RESERVE 64
#endif
constexpr int kDonateReserve = PROCELL_DONATE_RESERVE;   /* chunks kept waiting in the queue even when nobody starves yet */
/* PLAIN instances of the kernel are compiled for the common case "one parameter set, one checkpoint"; the run-time
 * checks of the general instance cost 3.7 % there (measured, config 2: 1.395 -> 1.342 ms) */
constexpr bool coop_is_plain(const SimParams& p) { return p.n_sets == 1u && p.n_times == 1u; }
#ifndef PROCELL_ENDGAME_IDLE
#define PROCELL_ENDGAME_IDLE 512
#endif
constexpr int kEndgameIdle = PROCELL_ENDGAME_IDLE;
#ifndef PROCELL_PROBE_MASK
#define PROCELL_PROBE_MASK 15u
#endif
constexpr uint32_t kProbeMask = PROCELL_PROBE_MASK;       /* busy warps look at the hunger snapshot every (mask + 1)-th iteration */
#ifndef PROCELL_SNAP_MASK
#define PROCELL_SNAP_MASK 63u
#endif
constexpr uint32_t kSnapMask = PROCELL_SNAP_MASK;         /* a warp refreshes its CTA's snapshot every (mask + 1)-th iteration (staggered by warp) */
/* end game (at least kEndgameIdle warps starve): 1 = a busy warp refreshes the snapshot at EVERY probe and may hand over
 * a chunk at every probe instead of once per snapshot epoch */
#ifndef PROCELL_ENDGAME_FAST
#define PROCELL_ENDGAME_FAST 0
#endif
constexpr bool kEndgameFast = PROCELL_ENDGAME_FAST != 0;
#ifndef PROCELL_PARK_BACKOFF_MAX_NS
#define PROCELL_PARK_BACKOFF_MAX_NS 1024u
#endif
constexpr unsigned kParkBackoffMaxNs = PROCELL_PARK_BACKOFF_MAX_NS;   /* warps parked at a set switch (MODE 2) poll with back-off up to this */
constexpr unsigned kIdleBackoffMaxNs = PROCELL_IDLE_BACKOFF_MAX_NS;   /* idle warps poll with exponential back-off up to this */
constexpr uint32_t kDonateMinNodes = PROCELL_DONATE_MIN_NODES;        /* a warp gives a chunk away only when its ring is about to spill anyway */
static_assert((kHistFlushIters & (kHistFlushIters - 1u)) == 0u && kHistFlushIters >= 256u && kHistFlushIters <= (1u << 20), "");

/* slot of the shared-memory table -> index of the count tensor (SimParams::slot_mode; identity otherwise).  Only the
 * drains use it: a few thousand integer divisions per CTA and drain. */
__device__ __forceinline__ uint32_t slot_key(const SimParams& P, uint32_t s)
{
    if (!P.slot_mode) return s;
    if (s < P.slot_prolif_end) {
        const uint32_t bk = s / P.n_prolif;
        return bk * P.n_types + __ldg(P.rank_type + (s - bk * P.n_prolif));
    }
    const uint32_t q = s - P.slot_prolif_end, bin = q / P.n_quiet;
    return __ldg(P.bin_keybase + bin) * P.n_types + __ldg(P.rank_type + P.n_prolif + (q - bin * P.n_quiet));
}

__device__ __noinline__ void hist_drain(const SimParams& P, uint32_t* s_hist, int lane)
{
    for (uint32_t i = (uint32_t)lane; i < P.smem_hist_slots; i += 32u) {
        if (*reinterpret_cast<volatile uint32_t*>(s_hist + i) == 0u) continue;
        const uint32_t v = atomicExch(s_hist + i, 0u);
        if (v) atomicAdd(reinterpret_cast<unsigned long long*>(P.counts) + slot_key(P, i), (unsigned long long)v);
    }
}

/* ---- sweeps with a set-relative direct table (kernel MODE kModeSetDirect) ----
 * A sweep's key space (n_sets x n_keys x n_types) does not fit in shared memory, but ONE parameter set's does, and a
 * CTA works on one set at a time (batches).  The CTA's u32 table then holds the counts of the set whose first key is
 * `base` (kept in shared memory, s_ctl[7]); a leaf of any other set - a donated chunk from another CTA at the very end
 * of the run - goes straight to the int64 tensor in HBM.  `base` changes only inside setdirect_rendezvous, while every
 * warp of the CTA is parked there without a node, so no add can be in flight across a change.  Against the hashed
 * cache this removes MATCH + two ballots + a 64-bit CAS loop per DIVIDE iteration and every eviction. */
__device__ __forceinline__ void count_leaves_setdirect(const SimParams& P, uint32_t* s_hist, uint32_t key, uint32_t inc, uint32_t base)
{
    if (inc > 0u) {
        const uint32_t rel = key - base;
        if (rel < P.smem_hist_slots) atomicAdd(&s_hist[rel], inc);
        else atomicAdd(reinterpret_cast<unsigned long long*>(P.counts) + key, (unsigned long long)inc);
    }
}

__device__ __noinline__ void hist_drain_at(const SimParams& P, uint32_t* s_hist, int lane, uint32_t base)
{
    for (uint32_t i = (uint32_t)lane; i < P.smem_hist_slots; i += 32u) {
        if (*reinterpret_cast<volatile uint32_t*>(s_hist + i) == 0u) continue;
        const uint32_t v = atomicExch(s_hist + i, 0u);
        if (v) atomicAdd(reinterpret_cast<unsigned long long*>(P.counts) + base + i, (unsigned long long)v);
    }
}

/* every lane of the warp calls this; each may carry `inc` (0, 1 or 2) leaves for `key`.
 * direct mode: every lane issues its own shared atomic; the shared-memory atomic unit serialises equal addresses
 *   faster than any merging in registers (measured: config 2 1.64 -> 1.46 ms).
 * hashed mode: a slot update is a CAS loop, so equal keys are merged first: MATCH.ANY groups the lanes by key and the
 *   group's total is two population counts over the ballots of "has 1" and "has 2" - no REDUX, which the hardware
 *   executes once per distinct group - and one lane per distinct key updates the cache. */
template <bool HASHED>
__device__ __forceinline__ void warp_count_leaves(const SimParams& P, uint32_t* s_hist, uint32_t key, uint32_t inc)
{
#ifdef PROCELL_DIRECT_MATCH
    const bool direct = false;
#else
    const bool direct = !HASHED;
#endif
    if (direct) {
        if (inc > 0) hist_add<HASHED>(P, s_hist, key, inc);
        return;
    }
    const unsigned b1 = __ballot_sync(kFull, inc == 1u);
    const unsigned b2 = __ballot_sync(kFull, inc >= 2u);
    const unsigned has = b1 | b2;
    if (has == 0) return;
    if (inc > 0) {
        const unsigned grp = __match_any_sync(has, key);
        const uint32_t total = (uint32_t)__popc(grp & b1) + 2u * (uint32_t)__popc(grp & b2);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(grp) - 1)) hist_add<HASHED>(P, s_hist, key, total);
    }
}

/* watchdog: record where this warp is and abort the launch */
/* inlined at its five (cold) call sites: as a real call its argument registers - P.ctl, P.dbg, a zero - were set up at the
 * head of the main loop, five instructions in EVERY iteration */
#ifdef PROCELL_WD_NOINLINE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
void watchdog_fire(const SimParams& P, uint32_t gwarp, int lane, unsigned long long code,
                                           unsigned long long a, unsigned long long b, unsigned long long c,
                                           unsigned long long d, unsigned long long e, unsigned long long f)
{
    if (lane == 0) {
        unsigned long long* r = P.dbg + (size_t)gwarp * kDbgWords;
        r[0] = code; r[1] = a; r[2] = b; r[3] = c; r[4] = d; r[5] = e; r[6] = f; r[7] = global_timer_ns();
        __threadfence();
        atomicCAS(&P.ctl->status, kStatusOk, kStatusWatchdog);
    }
}

/* MODE kModeSetDirect: a warp that holds no node and finds the CTA's batch used up comes here.  The LAST of the CTA's
 * warps to arrive - every other one is parked in the wait below, none holds a node, no add is in flight - drains the
 * table of the old set into the int64 tensor, takes the next batch from the global cursor, re-bases the table on that
 * batch's set and releases the others.  shared control words: s_ctl[7] table base (first key of the CTA's set),
 * s_ctl[12] arrivals of this round, s_ctl[13] round number.  Waiting is lane 0's alone (see idle_wait) and bounded by
 * the watchdog deadline; false = abort. */
template <int WARPS>
__device__ __forceinline__ bool setdirect_rendezvous(const SimParams& P, volatile int* s_ctl, uint32_t* s_hist,
                                                     unsigned long long* s_batch, int lane, uint32_t gwarp)
{
    __threadfence_block();      /* this warp's shared atomics are performed before its arrival is counted */
    __syncwarp();
    int role = 0;               /* 1 last to arrive: does the switch, 2 released, 3 abort */
    if (lane == 0) {
        const int round = sflag_get(s_ctl + 13);
        const int arrived = atomicAdd(const_cast<int*>(s_ctl) + 12, 1) + 1;
        if (sflag_get(s_ctl + 3)) role = 2;              /* another warp has just found the cursor exhausted: nothing to switch to */
        else if (arrived == WARPS) role = 1;
        else {
            const unsigned long long deadline = *reinterpret_cast<volatile unsigned long long*>(const_cast<int*>(s_ctl) + 10);
            unsigned backoff = 64;
            for (;;) {
                if (sflag_get(s_ctl + 13) != round || sflag_get(s_ctl + 3)) { role = 2; break; }     /* released, or no batch left anywhere */
                __nanosleep(backoff);
                if (backoff < kParkBackoffMaxNs) backoff <<= 1;
                if (global_timer_ns() > deadline || ld_volatile_s32(&P.ctl->status) != kStatusOk) { role = 3; break; }
            }
        }
    }
    role = __shfl_sync(kFull, role, 0);
    if (role == 2) return true;
    if (role == 3) {
        watchdog_fire(P, gwarp, lane, 4, (unsigned long long)s_ctl[12], (unsigned long long)s_ctl[13], 0, 0, 0, 0);
        return false;
    }
    hist_drain_at(P, s_hist, lane, (uint32_t)__shfl_sync(kFull, lane == 0 ? sflag_get(s_ctl + 7) : 0, 0));
    __syncwarp();
    if (lane == 0) {
        const unsigned long long g = atomicAdd(&P.ctl->cursor, 1ull);
        if (g >= P.total_batches) {
            sflag_set(s_ctl + 3, 1);                        /* no batch left anywhere: the base stays where it is */
            atomicMin(&P.ctl->t_exhausted, global_timer_ns());
        } else {
            sflag_set(s_ctl + 7, (int)((uint32_t)(g / P.batches_per_set) * P.smem_hist_slots));
            atomicExch(s_batch, g << 24);
        }
        s_ctl[12] = 0;
        __threadfence_block();
        atomicAdd(const_cast<int*>(s_ctl) + 13, 1);         /* release the parked warps */
    }
    __syncwarp();
    return true;
}

struct WarpCtx {
    uint32_t ab;                     /* ring: shared-memory address of kCap (t_div bits, heap) pairs followed by kCap
                                        (root|key<<32, D) pairs.  The ring starts on a 2 KB boundary of the shared-memory
                                        window (SmemLayout; static_assert in the kernel), so the address of the slot at byte
                                        position p is ab | (p & 0x7F0): ONE logic instruction, no shift, no add */
    uint32_t bottom, top;            /* ring positions IN BYTES of the (A, B) half - kSlot = 16 per node -, n = (top - bottom) / 16: a
                                        slot's address is ab | (position & 0x7F0) with no shift, and position + 16 * count is one IMAD */
    uint32_t slow;                   /* the bottom-most slow / 16 nodes of the ring are retry nodes awaiting a general iteration (bytes too) */
    uint32_t nl16;                   /* -16 * (lane + 1): the newest node this lane pops sits at top + nl16 */
    uint32_t sp;                     /* private spill ring in HBM, packed into one register: oldest chunk's position in the low
                                        half (mod 2^16; kSpillCap divides it), number of chunks in the high half.  The ring's
                                        address is recomputed on use (spill_ring), which keeps
                                        a 64-bit pointer out of the 64 registers of the steady state */
    int lane;
};

/* nodes in the ring, recomputed where it is needed (cold code) instead of kept from the top of the main loop: the
 * empty asm keeps the compiler from re-using - and therefore keeping alive, in a register or a spill slot - the hot
 * loop's copy of the same difference */
__device__ __forceinline__ uint32_t ring_nodes(const WarpCtx& w)
{
    uint32_t t = w.top;
    asm volatile("" : "+r"(t));
    return (t - w.bottom) >> 4;
}

template <int RING>
__device__ __forceinline__ void ring_load(const WarpCtx& w, uint32_t pos, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d)
{
    const unsigned sab = w.ab | (pos & (Ring<RING>::kMask << 4));
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%4];\n\tld.shared.v2.u64 {%2, %3}, [%4+%5];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "r"(sab), "n"(Ring<RING>::kCap * 16u) : "memory");
}

template <int RING>
__device__ __forceinline__ void ring_store(WarpCtx& w, uint32_t pos, uint64_t a, uint64_t b, uint64_t c, uint64_t d)
{
    const unsigned sab = w.ab | (pos & (Ring<RING>::kMask << 4));
    asm volatile("st.shared.v2.u64 [%0], {%1, %2};\n\tst.shared.v2.u64 [%0+%5], {%3, %4};"
                 :: "r"(sab), "l"(a), "l"(b), "l"(c), "l"(d), "n"(Ring<RING>::kCap * 16u) : "memory");
}

/* A pop as the DIVIDE iterations want it: the (C, D) pair as four 32-bit words - root, key, D's low word, retry number -
 * so that nothing downstream masks or shifts a 64-bit register pair (the compiler otherwise multiplies Philox's first
 * round as a 64 x 64-bit product of a masked pair and adds a 64-bit constant to C to step the key). */
template <int RING>
__device__ __forceinline__ void ring_load_node(const WarpCtx& w, uint32_t pos, uint64_t& a, uint64_t& heap, uint32_t& root,
                                               uint32_t& key, uint32_t& dlo, uint32_t& retry)
{
    const unsigned sab = w.ab | (pos & (Ring<RING>::kMask << 4));
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%6];\n\tld.shared.v4.u32 {%2, %3, %4, %5}, [%6+%7];"
                 : "=l"(a), "=l"(heap), "=r"(root), "=r"(key), "=r"(dlo), "=r"(retry) : "r"(sab), "n"(Ring<RING>::kCap * 16u) : "memory");
}

/* the two pushes of a DIVIDE iteration, each under its predicate, in ONE statement: daughter c goes to slot idx_c as
 * (t_div bits, heap) and (root, key, D's low word, D's high word); the second pair is the same four registers for both
 * daughters, so the compiler builds it once (as two statements it built it twice). */
template <int RING>
__device__ __forceinline__ void ring_store_daughters_if(WarpCtx& w, bool p0, uint32_t pos0, uint64_t a0, uint64_t heap0,
                                                        bool p1, uint32_t pos1, uint64_t a1, uint64_t heap1,
                                                        uint32_t root, uint32_t key, uint32_t dlo, uint32_t dhi)
{
    const unsigned s0 = w.ab | (pos0 & (Ring<RING>::kMask << 4));
    const unsigned s1 = w.ab | (pos1 & (Ring<RING>::kMask << 4));
    uint64_t c, d;             /* packed here, as register pairs: both stores then name the same two 64-bit registers */
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "r"(root), "r"(key));
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(dlo), "r"(dhi));
    asm volatile("{\n\t.reg .pred q0, q1;\n\tsetp.ne.u32 q0, %0, 0;\n\tsetp.ne.u32 q1, %1, 0;\n\t"
                 "@q0 st.shared.v2.u64 [%2], {%4, %5};\n\t"
                 "@q0 st.shared.v2.u64 [%2+%10], {%8, %9};\n\t"
                 "@q1 st.shared.v2.u64 [%3], {%6, %7};\n\t"
                 "@q1 st.shared.v2.u64 [%3+%10], {%8, %9};\n\t}"
                 :: "r"((unsigned)p0), "r"((unsigned)p1), "r"(s0), "r"(s1), "l"(a0), "l"(heap0), "l"(a1), "l"(heap1),
                    "l"(c), "l"(d), "n"(Ring<RING>::kCap * 16u) : "memory");
}

/* Philox block of (root, set, retry, tag, heap).  The empty asm statements pin the two 32-bit counter words that come out
 * of 64-bit values, so that the first round's products are 32 x 32 -> 64 (one IMAD.WIDE each) like the other nine: left
 * to itself the compiler sees zext(trunc(x)) = x & 0xFFFFFFFF and emits a 64-bit multiply (two more instructions each). */
__device__ __forceinline__ pcs_u32x4 draw_block(uint32_t root, uint32_t set, uint32_t retry, uint32_t tag, uint64_t heap, const uint32_t* rk)
{
    uint32_t h_lo = (uint32_t)heap, h_hi = (uint32_t)(heap >> 32);
    asm("" : "+r"(root), "+r"(h_lo));
    return pcs_philox4x32_10_rk(root, set | (retry << 16) | (tag << 24), h_lo, h_hi, rk);
}

/* the warp's private spill ring: [grid * warps][kSpillCap][kChunkWords] */
__device__ __forceinline__ unsigned long long* spill_ring(const SimParams& P)
{
    return P.spill + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kSpillCap * kChunkWords;
}

/* take the 32 oldest REGULAR nodes out of the ring (spill, donation): they sit above the retry nodes collected at the
 * bottom, which move up into the gap - a chunk that leaves the warp must not take retry nodes along, or whoever pops it
 * later pays a general DIVIDE iteration for a few of them.  With a whole warp's worth of retry nodes at the bottom the
 * chunk is 32 of those instead (it then is expanded by ONE general iteration, every lane busy).  The caller guarantees 32
 * regular nodes or 32 retry nodes. */
template <int RING>
__device__ __forceinline__ void pop_bottom_chunk(WarpCtx& w, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d)
{
    constexpr uint32_t kChunk = (uint32_t)kChunkNodes * kSlot;
    const uint32_t lane_pos = (uint32_t)w.lane * kSlot;
    if (w.slow >= kChunk) {
        ring_load<RING>(w, w.bottom + lane_pos, a, b, c, d);
        w.slow -= kChunk;
    } else {
        ring_load<RING>(w, w.bottom + w.slow + lane_pos, a, b, c, d);
        if (w.slow != 0u) {
            uint64_t sa = 0, sb = 0, sc = 0, sd = 0;
            if (lane_pos < w.slow) ring_load<RING>(w, w.bottom + lane_pos, sa, sb, sc, sd);
            __syncwarp();       /* every lane has read its chunk node and its retry node before any slot is overwritten */
            if (lane_pos < w.slow) ring_store<RING>(w, w.bottom + kChunk + lane_pos, sa, sb, sc, sd);
        }
    }
    w.bottom += kChunk;
    __syncwarp();
}

template <int RING>
__device__ __forceinline__ void spill_bottom_chunk(WarpCtx& w, const SimParams& P)
{
    unsigned long long* dst = spill_ring(P) + (size_t)(((w.sp & 0xFFFFu) + (w.sp >> 16)) % kSpillCap) * kChunkWords;
    uint64_t a, b, c, d;
    pop_bottom_chunk<RING>(w, a, b, c, d);
    __stcg(dst + w.lane, a);
    __stcg(dst + 32 + w.lane, b);
    __stcg(dst + 64 + w.lane, c);
    __stcg(dst + 96 + w.lane, d);
    w.sp += 0x10000u;
    if ((w.sp >> 16) > (uint32_t)kSpillCap && w.lane == 0) atomicExch(&P.ctl->status, kStatusSpillOverflow);
    __syncwarp();
}

/* a chunk of 32 nodes enters the ring (from the spill ring or the donation queue).  INVARIANT of the ring: above the
 * `slow` retry nodes at its bottom every node is a FRESH one (first draw, both daughters wanted) - that is what lets the
 * common DIVIDE iteration pop 32 nodes without looking at them.  Chunks are all-fresh or all-retry by construction
 * (pop_bottom_chunk); one with ANY retry node in it goes to the bottom as a whole, where the general iteration - which
 * takes any node - expands it; a fresh one goes on top: the caller holds fewer than 32 regular nodes, so it is expanded
 * next either way. */
template <int RING>
__device__ __forceinline__ void accept_chunk(WarpCtx& w, uint64_t a, uint64_t b, uint64_t c, uint64_t d)
{
    if (__any_sync(kFull, (d >> 28) != 3ull)) {          /* mask != 3, or counted, or retry > 0 */
        w.bottom -= (uint32_t)kChunkNodes * kSlot;
        ring_store<RING>(w, w.bottom + (uint32_t)w.lane * kSlot, a, b, c, d);
        w.slow += (uint32_t)kChunkNodes * kSlot;
    } else {
        ring_store<RING>(w, w.top + (uint32_t)w.lane * kSlot, a, b, c, d);
        w.top += (uint32_t)kChunkNodes * kSlot;
    }
    __syncwarp();
}

template <int RING>
__device__ __forceinline__ void unspill_newest_chunk(WarpCtx& w, const SimParams& P)
{
    w.sp -= 0x10000u;
    const unsigned long long* src = spill_ring(P) + (size_t)(((w.sp & 0xFFFFu) + (w.sp >> 16)) % kSpillCap) * kChunkWords;
    accept_chunk<RING>(w, __ldcg(src + w.lane), __ldcg(src + 32 + w.lane), __ldcg(src + 64 + w.lane), __ldcg(src + 96 + w.lane));
}

/* bounded MPMC queue of chunks (Vyukov): slot s is writable by ticket p when seq[s]==p, readable when seq[s]==p+1 */
__device__ __forceinline__ bool queue_push(const SimParams& P, int lane, uint64_t a, uint64_t b, uint64_t c, uint64_t d)
{
    unsigned long long pos = 0;
    if (lane == 0) pos = atomicAdd(&P.ctl->q_tail, 1ull);
    pos = __shfl_sync(kFull, pos, 0);
    uint32_t slot = (uint32_t)pos & (kQueueCap - 1);
    int ok = 1;
    if (lane == 0) {
        unsigned long long t0 = global_timer_ns();
        while (ld_acquire_u64(P.q_seq + slot) != pos) {
            __nanosleep(100);
            if (global_timer_ns() - t0 > 20000000000ull) { atomicExch(&P.ctl->status, kStatusQueueTimeout); ok = 0; break; }
        }
    }
    ok = __shfl_sync(kFull, ok, 0);
    if (!ok) return false;
    unsigned long long* dst = P.q_data + (size_t)slot * kChunkWords;
    __stcg(dst + lane, a);
    __stcg(dst + 32 + lane, b);
    __stcg(dst + 64 + lane, c);
    __stcg(dst + 96 + lane, d);
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        st_release_u64(P.q_seq + slot, pos + 1);
        atomicAdd(reinterpret_cast<unsigned long long*>(&P.ctl->pending), 1ull);   /* the chunk counts as pending work ... */
        __threadfence();
        atomicAdd(&P.ctl->avail, 1);          /* ... before its permit exists; the permit only after the chunk is published */
    }
    return true;
}

/* read the chunk of head ticket h (the caller holds a permit, so the chunk is or will shortly be published) */
__device__ __forceinline__ bool queue_read_ticket(const SimParams& P, int lane, unsigned long long h,
                                                  uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d)
{
    const uint32_t slot = (uint32_t)h & (kQueueCap - 1);
    int ok = 1;
    if (lane == 0) {
        unsigned long long t0 = global_timer_ns();
        while (ld_acquire_u64(P.q_seq + slot) != h + 1) {
            __nanosleep(100);
            if (global_timer_ns() - t0 > 20000000000ull) { atomicExch(&P.ctl->status, kStatusQueueTimeout); ok = 0; break; }
        }
    }
    ok = __shfl_sync(kFull, ok, 0);
    if (!ok) return false;
    __threadfence();
    const unsigned long long* src = P.q_data + (size_t)slot * kChunkWords;
    a = __ldcg(src + lane);
    b = __ldcg(src + 32 + lane);
    c = __ldcg(src + 64 + lane);
    d = __ldcg(src + 96 + lane);
    __syncwarp();
    if (lane == 0) st_release_u64(P.q_seq + slot, h + kQueueCap);
    return true;
}

/* hand the shallowest chunk (oldest spilled, else ring bottom) to the shared queue */
template <int RING>
__device__ __forceinline__ void donate_chunk(WarpCtx& w, const SimParams& P)
{
    uint64_t a, b, c, d;
    if ((w.sp >> 16) != 0u) {
        const unsigned long long* src = spill_ring(P) + (size_t)((w.sp & 0xFFFFu) % kSpillCap) * kChunkWords;
        a = __ldcg(src + w.lane);
        b = __ldcg(src + 32 + w.lane);
        c = __ldcg(src + 64 + w.lane);
        d = __ldcg(src + 96 + w.lane);
        w.sp = ((w.sp + 1u) & 0xFFFFu) | ((w.sp & 0xFFFF0000u) - 0x10000u);
    } else {
        pop_bottom_chunk<RING>(w, a, b, c, d);
    }
    __syncwarp();
    queue_push(P, w.lane, a, b, c, d);
}

/* A warp with nothing left: wait for donated work or for global quiescence.  true = a chunk was loaded.
 * Only ONE idle warp per CTA touches the global control block at a time (shared-memory lock): it polls and, if a
 * permit is available, claims a chunk with fetch-adds only (permit counter, then head ticket) - no CAS retry
 * storms, at most 148 concurrent pollers.  Quiescence (no active warp, no permit) is stable, so the warp that
 * observes it publishes it to its CTA through s_ctl[1]. */
template <int RING>
__device__ __forceinline__ bool idle_wait(WarpCtx& w, const SimParams& P, volatile int* s_ctl, unsigned long long deadline, uint32_t gwarp)
{
    ControlBlock* ctl = P.ctl;
#ifndef PROCELL_NO_FAST_TAKE
    {   /* fast path: a published chunk is already waiting - take it with two fetch-adds and stay "active" (no idle /
         * active bookkeeping, no CTA poll lock); the chunk is then read exactly as on the slow path */
        int fast = 0;
        unsigned long long ticket = 0;
        if (w.lane == 0 && ld_acquire_s32(&ctl->avail) > 0) {
            if (atomicSub(&ctl->avail, 1) > 0) {
                ticket = atomicAdd(&ctl->q_head, 1ull);
                fast = 1;
                /* warp and chunk were two units of pending work, now they are one */
                atomicAdd(reinterpret_cast<unsigned long long*>(&ctl->pending), ~0ull);
            }
            else atomicAdd(&ctl->avail, 1);                       /* lost the race for the last permit */
        }
        fast = __shfl_sync(kFull, fast, 0);
        if (fast) {
            ticket = __shfl_sync(kFull, ticket, 0);
            uint64_t a, b, c, d;
            if (!queue_read_ticket(P, w.lane, ticket, a, b, c, d)) return false;     /* watchdog abort */
            accept_chunk<RING>(w, a, b, c, d);
            return true;
        }
    }
#endif
    if (w.lane == 0) {
        atomicAdd(&ctl->idle, 1);
        __threadfence();
        atomicSub(&ctl->active, 1);
        atomicAdd(reinterpret_cast<unsigned long long*>(&ctl->pending), ~0ull);
    }
    /* The wait itself is lane 0's alone: the other lanes park at the shuffle below and issue nothing, so an idle warp
     * costs a dozen instructions per poll instead of a warp-wide loop with a collective in it (idle polling was 10 % of
     * all executed warp instructions of config 2 and competes with the last busy warps for issue slots). */
    int state = 0;   /* 1 ticket taken, 2 exit */
    unsigned long long ticket = 0;
    if (w.lane == 0) {
        const unsigned long long t0 = global_timer_ns();
        unsigned backoff = 128;
        for (;;) {
            if (sflag_get(s_ctl + 1)) state = 2;
            else if (s_ctl[0] == 0 && atomicCAS(const_cast<int*>(s_ctl), 0, 1) == 0) {
                /* Termination reads ONE word: `pending` = warps that hold or may find work + published chunks not yet
                 * claimed.  A chunk enters it before its permit exists, an idle warp that claims a chunk takes the
                 * chunk's place in it (no change), and a warp leaves it only on going idle - so it is 0 only when no
                 * work exists anywhere, and then for good.  (The earlier test, active == 0 and then avail == 0, could
                 * be fooled by a warp that re-activated and took the last permit between the two reads.) */
                const long long pend = (long long)ld_acquire_u64(reinterpret_cast<const unsigned long long*>(&ctl->pending));
                const int act = ld_acquire_s32(&ctl->active);
                const int av = ld_acquire_s32(&ctl->avail);
                const int st = ld_volatile_s32(&ctl->status);
                if (st != kStatusOk) state = 2;
                else if (av > 0) {
                    atomicAdd(&ctl->active, 1);                    /* re-activate BEFORE taking work */
                    if (atomicSub(&ctl->avail, 1) > 0) {
                        ticket = atomicAdd(&ctl->q_head, 1ull);    /* this warp takes the chunk's place in `pending` */
                        state = 1;
                    } else {                                       /* lost the race for the last permit */
                        atomicAdd(&ctl->avail, 1);
                        __threadfence();
                        atomicSub(&ctl->active, 1);
                    }
                }
                else if (pend == 0) state = 2;
                else if (global_timer_ns() > deadline) {
                    watchdog_fire(P, gwarp, 0, 3, (unsigned long long)act, (unsigned long long)(long long)av, (unsigned long long)ld_volatile_s32(&ctl->idle), global_timer_ns() - t0, 0, 0);
                    state = 2;
                }
                if (state == 2) sflag_set(s_ctl + 1, 1);
                __threadfence_block();
                atomicExch(const_cast<int*>(s_ctl), 0);
            }
            if (state != 0) break;
            __nanosleep(backoff);
            if (backoff < kIdleBackoffMaxNs) backoff <<= 1;
        }
        atomicSub(&ctl->idle, 1);
        atomicAdd(&ctl->idle_ns, global_timer_ns() - t0);
        atomicAdd(&ctl->idle_waits, 1ull);
    }
    state = __shfl_sync(kFull, state, 0);
    if (state == 2) return false;
    ticket = __shfl_sync(kFull, ticket, 0);
    uint64_t a, b, c, d;
    if (!queue_read_ticket(P, w.lane, ticket, a, b, c, d)) return false;     /* watchdog abort */
    accept_chunk<RING>(w, a, b, c, d);
    return true;
}

/* result of building one seed cell (cell.cu:25-79 with type == -1, t == 0) */
struct SeedOut {
    uint32_t key;       /* count-tensor index of (set, bin, level 0, type) - or, SLOT, the table slot (SimParams::slot_mode) */
    uint32_t type, kdiv;
    int kind;           /* 0 dropped, 1 leaf at level 0, 2 living root */
    int count0;         /* level-0 leaves of this bin are output rows (value >= phi) */
    int quiescent;
    double t_div;
};

/* bin of a seed cell: largest b with bin_start[b] <= root (parser.cu "bounds"), searched in [lo, n_bins) */
__device__ __forceinline__ uint32_t find_bin(const SimParams& P, uint32_t root, uint32_t lo = 0)
{
    uint32_t hi = P.n_bins;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(P.bin_start + mid) <= root) lo = mid; else hi = mid;
    }
    return lo;
}

/* the same for the 32 consecutive seed cells root0 .. root0+31 of one SEED iteration, called by the whole warp.
 * The warp first finds root0's bin with a 32-ary search (every lane probes one boundary, a ballot counts the ones at
 * or below root0: two dependent loads for 1024 bins instead of ten); the lanes whose root lies beyond that bin's end
 * (sparse histograms) finish with a binary search above it. */
__device__ __forceinline__ uint32_t find_bin_warp(const SimParams& P, uint32_t root0, int lane, uint32_t hint)
{
    /* hint: the bin of the previous SEED iteration's last root when this one continues the same claim unit (roots
     * ascend within a unit, so bin_start[hint] <= root0 already holds), else 0xFFFFFFFF */
    uint32_t lo = 0, end = P.n_bins;                /* answer for root0 is in [lo, end) */
    if (hint != 0xFFFFFFFFu && root0 < __ldg(P.bin_start + hint + 1u)) { lo = hint; end = hint + 1u; }
    while (end - lo > 1u) {
        const uint32_t step = (end - lo + 31u) >> 5;
        const uint32_t probe = lo + ((uint32_t)lane + 1u) * step;
        const bool le = probe < end && __ldg(P.bin_start + probe) <= root0;
        lo += (uint32_t)__popc(__ballot_sync(kFull, le)) * step;      /* bin_start ascends: the hits are a prefix */
        end = end - lo < step ? end : lo + step;
    }
    const uint32_t root = root0 + (uint32_t)lane;
    if (root >= __ldg(P.bin_start + lo + 1u) && lo + 1u < P.n_bins) return find_bin(P, root, lo + 1u);
    return lo;
}

/* cum / sel / musd: the type tables of parameter set `set` (shared-memory copies or the HBM tables).
 * SLOT (PLAIN direct instances of the cooperative kernel): key is the slot of the shared-memory table instead of the
 * tensor index - proliferating types per key, then one row per bin for the quiescent types; a type's rank among its
 * kind is the number of types of that kind with a smaller file id (a popcount under SimParams::quiet_mask). */
template <bool SLOT>
__device__ __forceinline__ SeedOut build_seed(const SimParams& P, const double* s_log, uint32_t root, uint32_t set,
                                              uint32_t bin, const uint32_t* thr, const uint8_t* sel, const double2* musd)
{
    SeedOut o;
    const uint32_t kd = __ldg(P.bin_kdiv + bin);
    const uint32_t T = P.n_types;
    pcs_u32x4 w = pcs_draw_rk(root, set, 0u, PCS_TAG_SEED, 0ull, P.rk);      /* the seed cell's ONE block: type, age, radius, angle */
    const double u_age = pcs_u32unit(w.y);
    /* cell.cu:81-104: the FIRST j with u < cum[j]; Q17: none -> the last type.  u = (2x + 1) / 2^33 is below cum[j] exactly
     * for x <= thr[j] (the host computes thr in integers, hostio.cpp: procell_type_threshold), and the sums ascend, so j
     * is the number of thresholds x lies above: an integer compare and an add per type, no divergence */
    uint32_t j = 0u;
    const uint4* thr4 = reinterpret_cast<const uint4*>(thr);   /* four thresholds per 16-byte load; padding = 2^32 - 1 */
#pragma unroll 1
    for (uint32_t i = 0u; i < T; i += 4u) {
        const uint4 t = thr4[i >> 2];
        j += (uint32_t)(w.x > t.x) + (uint32_t)(w.x > t.y) + (uint32_t)(w.x > t.z) + (uint32_t)(w.x > t.w);
    }
    const uint32_t type = sel[j];
    const double2 ms = musd[type];
    o.type = type;
    o.kdiv = kd & 63u;
    const uint32_t keybase = __ldg(P.bin_keybase + bin);
    if (SLOT) {
        const unsigned long long below = (1ull << type) - 1ull;
        o.key = ms.x < 0.0 ? P.slot_prolif_end + bin * P.n_quiet + (uint32_t)__popcll(P.quiet_mask & below)
                           : keybase * P.n_prolif + (uint32_t)__popcll(~P.quiet_mask & below);
    } else {
        o.key = (set * P.n_keys + keybase) * T + type;
    }
    const bool count0 = (kd & 0x80u) != 0u;
    o.count0 = count0;
    o.quiescent = ms.x < 0.0;
    if (ms.x < 0.0) {                                          /* quiescent: timer -1, t 0 -> out_of_time */
        o.kind = count0 ? 1 : 0;
        o.t_div = 0.0;
        return o;
    }
    double timer = ms.x;                                       /* the mean, should 255 draws in a row be rejected */
    for (uint32_t retry = 0;;) {                               /* first timer from words z, w of the same block (cell.cu:106-122) */
        /* ideal seeding: one ziggurat trial on the 64 bits (z, w), as for daughter 1 of a division at heap 0 (extra uniforms
         * from the blocks tagged 2.. of (root, heap 0, retry)); refcompat seeding: the Box-Muller draw whose radius uniform
         * is the type uniform - the coupling is defined on that transform */
        double zn;
        bool acc = true;
        if (P.refcompat) zn = pcs_seed_normal(w, s_log, retry == 0u ? pcs_u32unit(w.x) : 0.0);
        else acc = pcs_zig_trial(w, 1u, &zn, root, set, retry, 0ull, P.rk, s_log, P.logtab + PCS_TAB_WEDGE);
        const double cand = pcs_timer(ms.x, ms.y, zn);
        if (acc && cand > 0.0) { timer = cand; break; }
        if (++retry == PCS_MAX_RETRY) break;
        w = pcs_draw_rk(root, set, retry, PCS_TAG_SEED, 0ull, P.rk);   /* rejected: words z, w of the next round's block */
    }
    const double t0 = PCS_MUL(timer, u_age);                   /* cell.cu:124-143 */
    const double t_div = PCS_ADD(t0, timer);
    o.t_div = t_div;
    if (t_div > P.t_max) o.kind = count0 ? 1 : 0;              /* proliferation.cu:404-410 */
    else o.kind = (o.kdiv > 0u) ? 2 : 0;                       /* f/2 <= phi: dropped (Q6) */
    return o;
}

/* per-lane division counters.  Sweeps and time series (general instances): divisions of parameter set `set` counted by
 * this lane and not yet flushed, flushed when the set changes.  PLAIN instances (one set): every DIVIDE iteration counts
 * as 32 divisions through the warp's iteration counter and `cnt` is this lane's CORRECTION to that (-1 when the lane had
 * no node or its node was a redraw), so that the common iteration - 32 lanes, 32 first expansions - does not touch it. */
struct DivCount {
    uint32_t set, cnt;
};

/* What one lane's popped node came to in a DIVIDE iteration; push_and_count turns it into ring pushes and leaf counts */
struct DivOut {
    bool int0, int1;            /* daughter 0 / 1 lives on and will divide */
    bool got0, got1;            /* daughter 0 / 1 received its timer in THIS iteration (time series) */
    bool rej0, rej1;            /* daughter 0 / 1 is still without a timer: the node goes back as a retry node */
    uint32_t retry_next;        /* retry number of that node */
    uint32_t leaf_inc, leaf_key, dlo;
    uint32_t root, key;         /* the node's C word: root cell id, count-tensor index (or table slot) of the node itself */
    uint32_t child_dhi;         /* high word of a daughter's D word: 0 (retry 0).  The common iteration passes the popped node's
                                   own high word, which is 0 by the ring's invariant - a register that is there already */
    uint64_t heap;
    double t_div, tc0, tc1;
};

__device__ __forceinline__ void divout_clear(DivOut& o)
{
    o.int0 = false; o.int1 = false; o.got0 = false; o.got1 = false;
    o.rej0 = false; o.rej1 = false; o.retry_next = 0; o.leaf_inc = 0; o.leaf_key = 0; o.dlo = 0;
    o.heap = 0; o.root = 0; o.key = 0; o.child_dhi = 0; o.t_div = 0.0; o.tc0 = 0.0; o.tc1 = 0.0;
}

/* bit 30 of a node's D word: the division of this node has been counted already (set on every retry node) */
constexpr uint32_t kDloCounted = 1u << 30;

/* The node rule for both daughters of one division (proliferation.cu:321-380, :404-410) given which daughters received a
 * timer (ok) and the timers; fills o.int / o.leaf_inc / o.tc */
__device__ __forceinline__ void classify_daughters(const SimParams& P, DivOut& o, bool ok0, bool ok1, double tm0, double tm1)
{
    o.tc0 = PCS_ADD(o.t_div, tm0);
    o.tc1 = PCS_ADD(o.t_div, tm1);
    const bool late0 = o.tc0 > P.t_max, late1 = o.tc1 > P.t_max;  /* proliferation.cu:404-410 */
    const bool deeper = (o.dlo & (63u << 22)) != 0u;              /* f/2 > phi one level down (:323) */
    /* (ok0 && late0) + (ok1 && late1) as select + predicated add: two instructions where the compiler's own form took four */
    asm("{\n\t.reg .pred pa, pb;\n\tsetp.ne.u32 pa, %1, 0;\n\tsetp.ne.u32 pb, %2, 0;\n\tselp.u32 %0, 1, 0, pa;\n\t@pb add.u32 %0, %0, 1;\n\t}"
        : "=r"(o.leaf_inc) : "r"((uint32_t)(ok0 && late0)), "r"((uint32_t)(ok1 && late1)));
    o.int0 = ok0 && !late0 && deeper;
    o.int1 = ok1 && !late1 && deeper;
    o.got0 = ok0; o.got1 = ok1;
}

/* subtree sharding and the division counters: `first` = this iteration is the node's first expansion (a redraw is not
 * another division) */
template <int MODE, bool PLAIN>
__device__ __forceinline__ void credit_division(const SimParams& P, DivOut& o, uint32_t first, uint32_t set, bool multi_set, DivCount& dc)
{
    if (MODE == kModeSubtree && o.heap < P.sub_limit) {
        /* subtree sharding: a node below the shard level is expanded by EVERY GPU (same stream, same outcome);
         * its division and the leaves among its daughters are credited to GPU root % world only, and of its
         * daughters AT the shard level this GPU keeps the ones with (root + heap) % world == rank */
        const uint32_t root = o.root;
        if (root % P.sub_world != P.sub_rank) { o.leaf_inc = 0u; first = 0u; }
        if (o.heap >= (P.sub_limit >> 1)) {
            const uint32_t h0 = root + (uint32_t)o.heap * 2u;
            o.int0 = o.int0 && h0 % P.sub_world == P.sub_rank;
            o.int1 = o.int1 && (h0 + 1u) % P.sub_world == P.sub_rank;
        }
    }
    if (PLAIN) { dc.cnt += first - 1u; return; }
    if (multi_set && first && set != dc.set) {
        if (dc.cnt) atomicAdd(reinterpret_cast<unsigned long long*>(P.divisions) + dc.set, (unsigned long long)dc.cnt);
        dc.cnt = 0; dc.set = set;
    }
    dc.cnt += first;
}

/* second half of a DIVIDE iteration: `take` nodes have been popped (from the top, or - BOTTOM - from the bottom of the
 * ring); push the daughters that will divide on top, the retry nodes at the BOTTOM, count the leaves.
 * Retry nodes - a daughter whose ziggurat trial left the fast path, or whose timer came out <= 0 (cell.cu:114-118) - are
 * rare (2-3 % of the divisions) and expensive (a second Philox block, a logarithm), so they are not expanded where they
 * arise: they collect at the bottom of the ring (w.slow counts them) and are expanded 32 at a time by a general iteration
 * with every lane busy.  Results do not depend on when a node is expanded: the stream is keyed by (root, path, retry). */
template <bool FULL, bool HASHED, bool PLAIN, int RING, int MODE>
__device__ __forceinline__ void push_and_count(WarpCtx& w, const SimParams& P, uint32_t* s_hist, const DivOut& o, uint32_t take,
                                               bool from_bottom, unsigned lt_mask, uint32_t hist_base)
{
    constexpr bool SETDIRECT = MODE == kModeSetDirect;
    const uint32_t KS = (PLAIN && !HASHED) ? P.kstride : P.n_types;
    /* all popped nodes have been read: every lane's loads have returned before it votes below (the predicates depend
     * on the loaded values of the node it popped) and no lane stores before all have voted on everything, so the
     * slots may be overwritten.  The barrier states that ordering formally (racecheck reports the pops and pushes of one
     * iteration as a hazard without it); on a converged warp it is one WARPSYNC. */
    __syncwarp();
    if (from_bottom) { w.bottom += take * kSlot; w.slow -= take * kSlot; } else w.top -= take * kSlot;
    const unsigned b0 = __ballot_sync(kFull, o.int0);
    const unsigned b1 = __ballot_sync(kFull, o.int1);
    const unsigned br = __ballot_sync(kFull, o.rej0 || o.rej1);
    const uint32_t child_key = o.key + KS;                 /* one tree level down (= o.leaf_key) */
    const uint32_t child_dlo = ((o.dlo & ~kDloCounted) | (3u << 28)) - (1u << 22);
    const uint32_t i0 = w.top + kSlot * __popc(b0 & lt_mask);
    w.top += kSlot * __popc(b0);
    const uint32_t i1 = w.top + kSlot * __popc(b1 & lt_mask);
    w.top += kSlot * __popc(b1);
    ring_store_daughters_if<RING>(w, o.int0, i0, pcs_d2bits(o.tc0), o.heap * 2ull, o.int1, i1, pcs_d2bits(o.tc1), o.heap * 2ull + 1ull,
                                  o.root, child_key, child_dlo, o.child_dhi);
    if (br) {
        if (o.rej0 || o.rej1) {
            const uint32_t rej = (uint32_t)o.rej0 | ((uint32_t)o.rej1 << 1);
            const uint32_t ir = w.bottom - kSlot - kSlot * __popc(br & lt_mask);
            ring_store<RING>(w, ir, pcs_d2bits(o.t_div), o.heap, (uint64_t)o.root | ((uint64_t)o.key << 32),
                             (uint64_t)((o.dlo & ~(3u << 28)) | (rej << 28) | kDloCounted) | ((uint64_t)o.retry_next << 32));
        }
        w.bottom -= kSlot * __popc(br);
        w.slow += kSlot * __popc(br);
    }
    __syncwarp();
    if (SETDIRECT) {
        count_leaves_setdirect(P, s_hist, o.leaf_key, o.leaf_inc, hist_base);
    } else if (PLAIN || P.n_times == 1u) {
        if (MODE == kModeMerge) {        /* (the subtree-sharding instance does not gain from it: 77.9 -> 78.6 ms for one rank of eight of config 4) */
            /* The 32 newest nodes of a depth-first front mostly belong to one lineage and one or two tree levels, so most lanes
             * count into the SAME slot and the shared-memory atomic unit serialises them: 10.5 wavefronts per iteration for 17
             * lanes (ncu), a sixth of the shared-memory traffic of a kernel that is bound by it on deep trees.  MATCH.ANY groups
             * the lanes by key, a group's total is two population counts over the ballots "one leaf" / "two leaves", and its
             * lowest lane adds: 8 wavefronts less for 11 instructions more - config 4 -2.4 %, config 2 -1.3 %, but +3.5 % on
             * config 3, where a DIVIDE iteration is half the work and issue is the bound: hence an instance of its own. */
            const unsigned b1 = __ballot_sync(kFull, o.leaf_inc == 1u);
            const unsigned b2 = __ballot_sync(kFull, o.leaf_inc == 2u);
            const unsigned grp = __match_any_sync(kFull, o.leaf_key);
            const uint32_t total = (uint32_t)__popc(grp & b1) + 2u * (uint32_t)__popc(grp & b2);
            if ((grp & lt_mask) == 0u && total > 0u) atomicAdd(&s_hist[o.leaf_key], total);
        } else
        warp_count_leaves<HASHED>(P, s_hist, o.leaf_key, o.leaf_inc);
    } else {
        /* time series: a daughter born at t_div that divides (or would divide) at tc is out of time at every
         * checkpoint in [t_div, tc) */
        for (uint32_t j = 0; j < P.n_times; ++j) {
            const double tj = P.times[j];
            const bool born = o.t_div <= tj;
            const uint32_t inc = (uint32_t)(o.got0 && born && tj < o.tc0) + (uint32_t)(o.got1 && born && tj < o.tc1);
            warp_count_leaves<HASHED>(P, s_hist, o.leaf_key + j * P.time_stride, inc);
        }
    }
}

/* ---- DIVIDE iteration, the common one: the lanes below `take` pop one node each from the top of the ring (newest
 * first), draw ONE Philox block and make the FAST ziggurat test for both daughters (procell_spec.h: one table row, one
 * fma, one compare each).  A daughter that passes and whose timer is positive is classified as leaf / dropped / internal;
 * the others go back as a retry node for the general iteration below.
 * FULL = all 32 lanes have a node: straight-line code.  Otherwise the lanes without a node skip the arithmetic, and every
 * warp collective is still executed by all 32 lanes with the full mask.
 * The popped nodes are fresh by the ring's invariant (accept_chunk): nothing is checked here. */
template <int WARPS, bool FULL, bool HASHED, bool PLAIN, int RING, int MODE>
__device__ __forceinline__ void divide_fresh(WarpCtx& w, const SimParams& P, const double* s_tab, uint32_t* s_hist,
                                             const double2* musd, uint32_t take, unsigned lt_mask, bool multi_set,
                                             DivCount& dc, uint32_t hist_base)
{
    DivOut o;
    divout_clear(o);
    const bool mine = FULL || (uint32_t)w.lane < take;
    pcs_u32x4 blk;
    blk.x = 0; blk.y = 0; blk.z = 0; blk.w = 0;
    uint32_t set = 0;
    if (mine) {
        uint64_t a;
        ring_load_node<RING>(w, w.top + w.nl16, a, o.heap, o.root, o.key, o.dlo, o.child_dhi);      /* high word of D: 0, the node is fresh */
        o.t_div = pcs_bits2d(a);
        set = PLAIN ? 0u : (o.dlo & 0xFFFFu);
        blk = draw_block(o.root, set, 0u, PCS_TAG_DIVISION, o.heap, P.rk);
        using L = SmemLayout<WARPS, RING>;
        /* PLAIN: one set, the (mean, sd) table is the shared-memory copy; else a generic pointer - that copy or the HBM table */
        const double2 ms = PLAIN ? musd_smem<L::kMusdAddr>(o.dlo) : musd[set * P.n_types + ((o.dlo >> 16) & 63u)];
        double z0, z1;
        const bool f0 = zig_fast_smem<L::kZigAddr>(blk.x, blk.y, &z0);
        const bool f1 = zig_fast_smem<L::kZigAddr>(blk.z, blk.w, &z1);
        const double tm0 = pcs_timer(ms.x, ms.y, z0), tm1 = pcs_timer(ms.x, ms.y, z1);
        const bool ok0 = f0 && tm0 > 0.0, ok1 = f1 && tm1 > 0.0;
        classify_daughters(P, o, ok0, ok1, tm0, tm1);
        o.rej0 = !ok0; o.rej1 = !ok1;
        /* a trial that left the fast path is finished by the general iteration at THIS retry number; if every rejected
         * daughter passed the fast test (its timer was <= 0), the next thing to do is the redraw */
        o.retry_next = (f0 && f1) ? 1u : 0u;
        o.leaf_key = o.key + ((PLAIN && !HASHED) ? P.kstride : P.n_types);
        credit_division<MODE, PLAIN>(P, o, 1u, set, multi_set, dc);
    }
    else if (PLAIN) dc.cnt -= 1u;
    push_and_count<FULL, HASHED, PLAIN, RING, MODE>(w, P, s_hist, o, take, false, lt_mask, hist_base);
}

/* ---- DIVIDE iteration, the general one: any node - any retry number, either or both daughters wanted.  Each wanted
 * daughter gets one WHOLE ziggurat trial at the node's retry number (fast test, else wedge test or tail sampler with
 * further Philox blocks, procell_spec.h: pcs_zig_trial); accepted with a positive timer it is classified, else it stays
 * wanted and the node goes back with retry + 1; at retry 255 the timer is the mean.  The popped nodes are the `take`
 * bottom-most of the ring (the collected retry nodes) or - a mixed chunk on top - the newest ones. */
template <bool HASHED, bool PLAIN, int RING, int MODE>
__device__ __forceinline__ void divide_general(WarpCtx& w, const SimParams& P, const double* s_tab, uint32_t* s_hist,
                                               const double2* musd, uint32_t take, bool from_bottom, unsigned lt_mask,
                                               bool multi_set, DivCount& dc, uint32_t hist_base)
{
    DivOut o;
    divout_clear(o);
    const bool mine = (uint32_t)w.lane < take;
    uint32_t retry = 0u, set = 0u, want = 0u;
    double2 ms = make_double2(0.0, 0.0);
    pcs_u32x4 blk;
    blk.x = 0; blk.y = 0; blk.z = 0; blk.w = 0;
    if (mine) {
        const uint32_t pos = from_bottom ? w.bottom + (uint32_t)w.lane * kSlot : w.top + w.nl16;
        uint64_t a;
        ring_load_node<RING>(w, pos, a, o.heap, o.root, o.key, o.dlo, retry);
        o.t_div = pcs_bits2d(a);
        set = PLAIN ? 0u : (o.dlo & 0xFFFFu);
        ms = musd[set * P.n_types + ((o.dlo >> 16) & 63u)];
        want = (o.dlo >> 28) & 3u;
        blk = draw_block(o.root, set, retry, PCS_TAG_DIVISION, o.heap, P.rk);
    }
    const bool forced = retry >= PCS_MAX_RETRY;                /* 255 redraws failed: the timer is the mean */
    /* The trials run in two passes of warp-uniform shape: pass 0 gives every lane's FIRST wanted daughter its trial (most
     * retry nodes want one daughter only), pass 1 - skipped by the whole warp when no node wants both - the second one.
     * Within a pass the slow part (second Philox block, wedge test or tail sampler) is entered by the lanes whose fast test
     * failed, and skipped by the warp when there is none: a node that is here for a non-positive timer redraws with a
     * fresh block and passes the fast test like any other. */
    double zc[2] = { 0.0, 0.0 };
    bool accc[2] = { false, false };
#pragma unroll
    for (uint32_t pass = 0; pass < 2u; ++pass) {
        const bool act = mine && !forced && (pass == 0u ? want != 0u : want == 3u);
        if (!__any_sync(kFull, act)) continue;
        const uint32_t c = pass == 0u ? ((want & 1u) ? 0u : 1u) : 1u;            /* which daughter this lane tries now */
        const uint32_t lo = c ? blk.z : blk.x, hi = c ? blk.w : blk.y;
        double z;
        bool acc = pcs_zig_fast(lo, hi, s_tab + PCS_TAB_ZIG, &z);
        if (act && !acc) acc = pcs_zig_slow(hi, c, &z, o.root, set, retry, o.heap, P.rk, s_tab, P.logtab + PCS_TAB_WEDGE);
        __syncwarp();
        if (act) {
            if (c) { zc[1] = z; accc[1] = acc; } else { zc[0] = z; accc[0] = acc; }
        }
    }
    if (mine) {
        const bool want0 = (want & 1u) != 0u, want1 = (want & 2u) != 0u;
        const double tm0 = forced ? ms.x : pcs_timer(ms.x, ms.y, zc[0]);
        const double tm1 = forced ? ms.x : pcs_timer(ms.x, ms.y, zc[1]);
        const bool ok0 = want0 && (forced || (accc[0] && tm0 > 0.0)), ok1 = want1 && (forced || (accc[1] && tm1 > 0.0));
        classify_daughters(P, o, ok0, ok1, tm0, tm1);
        o.rej0 = want0 && !ok0; o.rej1 = want1 && !ok1;
        o.retry_next = retry + 1u;
        o.leaf_key = o.key + ((PLAIN && !HASHED) ? P.kstride : P.n_types);
        credit_division<MODE, PLAIN>(P, o, (o.dlo & kDloCounted) ? 0u : 1u, set, multi_set, dc);
    }
    else if (PLAIN) dc.cnt -= 1u;
    push_and_count<false, HASHED, PLAIN, RING, MODE>(w, P, s_hist, o, take, from_bottom, lt_mask, hist_base);
}

}  // namespace

/* MODE: kModeBase, kModeSubtree (the subtree-sharding rule of multi-GPU runs of deep trees, SimParams::sub_world > 1)
 * or kModeSetDirect (sweeps: set-relative direct histogram table, SimParams::hist_setdirect) */
template <int WARPS, bool HASHED, bool PLAIN, int RING, int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1) k_proliferate_coop(const __grid_constant__ SimParams P)
{
    constexpr bool SUBTREE = MODE == kModeSubtree, SETDIRECT = MODE == kModeSetDirect;
    static_assert(RING == 1, "128-node ring per warp, one node per lane and iteration");
    static_assert(!SUBTREE || (PLAIN && RING == 1), "subtree sharding: one parameter set, one checkpoint, one node per lane");
    static_assert(!SETDIRECT || (!PLAIN && !HASHED && RING == 1), "set-relative table: sweeps, u32 slots, one node per lane");
    static_assert(MODE != kModeMerge || (PLAIN && !HASHED && RING == 1), "merged leaf counts: one parameter set, direct u32 table");
    constexpr uint32_t kCap = Ring<RING>::kCap;
    constexpr bool SLOT = PLAIN && !HASHED;      /* the u32 table is laid out by slots (SimParams::slot_mode is set) */
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    /* Dynamic shared memory starts kDynSmemWindow bytes into the CTA's shared-memory window (sm_100: the first 1 KB is the
     * system's), so the rings start on a 2 KB boundary - a ring slot's address is one OR (WarpCtx::ab) - and, with the
     * window offset a compile-time constant, every table address of the hot loop is an immediate (no window base to
     * rematerialise: S2R + MOV + LEA, and an add per access).  A launch that finds another offset stops with
     * kStatusSmemWindow before it touches anything (capi.cu reports it as a failed run). */
    using L = SmemLayout<WARPS, RING>;
    unsigned char* const smem_raw = reinterpret_cast<unsigned char*>(__cvta_shared_to_generic(kDynSmemWindow));
    if ((uint32_t)__cvta_generic_to_shared(smem_dyn) != kDynSmemWindow) {
        if (threadIdx.x == 0) atomicExch(&P.ctl->status, kStatusSmemWindow);
        return;
    }
    constexpr size_t kRingsOff = L::kRingsOff, kTabOff = L::kTabOff, kCtlOff = L::kCtlOff;
    static_assert(kSmemMusdEntries * 16 == 1024 && ((kDynSmemWindow + L::kRingsOff) & (kCap * 16u - 1u)) == 0u,
                  "the (mean, sd) table is the 1 KB in front of the rings, which start on a 2 KB boundary of the window");
    double* s_log = reinterpret_cast<double*>(smem_raw + kTabOff);
    volatile int* s_ctl = reinterpret_cast<volatile int*>(smem_raw + kCtlOff);   /* [0] poll lock, [1] quiescent */
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw + kCtlOff + (kSmemCtlBytes - kSmemMusdEntries * 16));

    double2* s_musd_buf = reinterpret_cast<double2*>(smem_raw);
    const bool musd_cached = P.n_sets * P.n_types <= (uint32_t)kSmemMusdEntries;
    uint32_t* s_thr_buf = reinterpret_cast<uint32_t*>(smem_raw + kCtlOff + 128);
    uint8_t* s_sel_buf = reinterpret_cast<uint8_t*>(smem_raw + kCtlOff + 128 + kSmemMusdEntries * 8);
    const uint32_t t_pad = (P.n_types + 3u) & ~3u;       /* row length of the threshold table */
    const bool thr_cached = musd_cached && P.n_sets * t_pad <= (uint32_t)kSmemMusdEntries;
    if (musd_cached && threadIdx.x < P.n_sets * P.n_types) {
        s_musd_buf[threadIdx.x] = __ldg(P.type_musd + threadIdx.x);
        s_sel_buf[threadIdx.x] = __ldg(P.type_sel + threadIdx.x);
    }
    if (thr_cached && threadIdx.x < P.n_sets * t_pad) s_thr_buf[threadIdx.x] = __ldg(P.type_thr + threadIdx.x);
    const double2* s_musd = musd_cached ? s_musd_buf : P.type_musd;
    const uint32_t* s_thr = thr_cached ? s_thr_buf : P.type_thr;
    const uint8_t* s_sel = musd_cached ? s_sel_buf : P.type_sel;
    if (threadIdx.x < 2) s_ctl[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < kLogTabDoubles; i += blockDim.x) s_log[i] = __ldg(P.logtab + i);
    if (HASHED) {
        unsigned long long* tab = reinterpret_cast<unsigned long long*>(s_hist);
        for (uint32_t i = threadIdx.x; i < P.smem_hist_slots; i += blockDim.x) tab[i] = 0xFFFFFFFF00000000ull;
    } else {
        for (uint32_t i = threadIdx.x; i < P.smem_hist_slots; i += blockDim.x) s_hist[i] = 0u;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    /* lanes below this one, from the special register: the compiler keeps it in a register through the loop, where it
     * rebuilt (1 << lane) - 1 in every iteration (-4 instructions per DIVIDE iteration) */
#ifdef PROCELL_LTMASK_SHIFT
    const unsigned lt_mask = (1u << lane) - 1u;
#else
    unsigned lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
#endif
    ControlBlock* ctl = P.ctl;

    WarpCtx w;
    w.ab = kDynSmemWindow + (uint32_t)kRingsOff + (uint32_t)warp * (4u * kCap * 8u);     /* on a 2 KB boundary (static_assert above) */
    w.bottom = 0; w.top = 0; w.slow = 0;
    w.nl16 = ~(uint32_t)lane * kSlot;
    asm volatile("" : "+r"(w.nl16));       /* a value of its own: kept in a register, not rebuilt from the lane id in every iteration */
    w.sp = 0;
    w.lane = lane;

    uint32_t seed_cur = 0, seed_end = 0, seed_set = 0;
    uint32_t seed_hint = 0xFFFFFFFFu;    /* bin of the last root of the previous SEED iteration of the same claim unit */
    /* s_ctl[3]: 1 once the seed-unit cursor has run out; s_ctl[5]: snapshot epoch; s_ctl[6]: batch-refill lock;
     * s_snap: copies of the control block's idle count, permit count and seed cursor, refreshed by one lane of the
     * CTA every few iterations with cp.async and read from shared memory: busy warps never poll HBM. */
    unsigned long long* s_batch = reinterpret_cast<unsigned long long*>(const_cast<int*>(s_ctl) + 8);
    volatile int* s_snap = s_ctl + 16;      /* [0] idle  [4] avail  [8..9] seed cursor: cp.async targets, 16 B each */
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile unsigned long long*>(const_cast<int*>(s_ctl) + 10) = global_timer_ns() + P.watchdog_ns;
        s_ctl[2] = 0; s_ctl[3] = P.total_local_units == 0; s_ctl[4] = 0; s_ctl[5] = 0; s_ctl[6] = 0;
        for (int i = 0; i < 12; ++i) s_snap[i] = 0;
        *s_batch = (0xFFFFFFFFFFull << 24) | 0x800000ull;   /* no batch yet (invalid id; offset bits leave room for increments) */
        if (SETDIRECT) { s_ctl[7] = 0; s_ctl[12] = 0; s_ctl[13] = 0; }
    }
    __syncthreads();
    DivCount dc;
    dc.set = 0; dc.cnt = 0;
    /* MODE 2: this warp's copy of the table base s_ctl[7].  The base changes only inside setdirect_rendezvous while this
     * warp is parked there, so it is re-read (by lane 0, atomically, then broadcast) exactly when the warp comes back */
    uint32_t set_base = 0u;
    const bool multi_set = !PLAIN && P.n_sets > 1u;
    uint32_t iter = 0;
    /* snapshot epoch of this warp's last donation: one word per warp in shared memory (the upper half of the threshold
     * buffer, free since the thresholds are 32-bit) - read by lane 0 every probe, written at a donation; as a register it
     * was the one value ptxas kept spilling at the 64-register cap */
    volatile int* s_epoch = reinterpret_cast<volatile int*>(s_thr_buf + kSmemMusdEntries) + warp;
    if (lane == 0) *s_epoch = -1;
    __syncwarp();

    if (lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&ctl->pending), 1ull);
        atomicAdd(&ctl->active, 1);
        atomicMin(&ctl->t_start, global_timer_ns());
    }
#define GWARP (blockIdx.x * WARPS + warp)
    /* watchdog deadline of this CTA, kept in shared memory (registers are scarce at 32 warps per SM) */
    volatile unsigned long long* s_deadline = reinterpret_cast<volatile unsigned long long*>(const_cast<int*>(s_ctl) + 10);

    for (;;) {
        /* the ring holds n nodes: w.slow retry nodes at its bottom, waiting for a general iteration, and above them the
         * regular ones.  mode 0: full fresh iteration (the common case: 32 regular nodes or more, room for the pushes of
         * one iteration, fewer than 32 retry nodes), 1: partial fresh iteration, 2: general iteration */
        const uint32_t n = w.top - w.bottom;                 /* bytes, like every ring position: kSlot per node */
        const uint32_t nreg = n - w.slow;
        int mode = 0;
        uint32_t take = 32u;
        bool from_bottom = false;
        if (!(w.slow < 32u * kSlot && nreg - 32u * kSlot <= (kCap - 64u) * kSlot - w.slow)) {
        if (n > (kCap - 32u) * kSlot) {    /* an iteration pops 32 nodes at most and pushes twice as many: keep that much room in the ring */
            TRACE(P, GWARP, lane, 40); spill_bottom_chunk<RING>(w, P); continue;
        }
        if (w.slow >= 32u * kSlot) { mode = 2; from_bottom = true; }      /* a warp's worth of retry nodes has collected */
        else {
            bool divide_rest = false;      /* MODE 2: expand what is left of the old set before waiting for the switch */
            if ((w.sp >> 16) != 0u) { TRACE(P, GWARP, lane, 41); unspill_newest_chunk<RING>(w, P); continue; }
            /* RULE: every decision that depends on mutable shared/global state is taken by lane 0 and broadcast.
             * Lanes of a warp are not guaranteed to be converged when they read a volatile flag, so a per-lane read
             * can see two different values inside one warp and split it for good. */
            int exhausted = 0;
            if (seed_cur == seed_end) {
                if (lane == 0) exhausted = sflag_get(s_ctl + 3);
                exhausted = __shfl_sync(kFull, exhausted, 0);
            }
            if (seed_cur != seed_end || !exhausted) {
                if (seed_cur == seed_end) {
                    TRACE(P, GWARP, lane, 10);
                    uint32_t set = 0, j = 0;
                    bool got = false, wait_switch = false;
                    if (!multi_set) {
                        unsigned long long c = 0;
                        if (lane == 0) c = atomicAdd(&ctl->cursor, 1ull);
                        c = __shfl_sync(kFull, c, 0);
                        got = c < P.total_local_units;
                        j = (uint32_t)c;
                    } else {
                        /* s_batch = batch id << 24 | next unit offset; one 64-bit shared atomicAdd claims a unit.
                         * Lane 0 runs the whole protocol and broadcasts (status, set, unit). */
                        int status = 0;          /* 0 retry, 1 got a unit, 2 no seed units left, 3 watchdog, 4 wait for the switch */
                        for (int spin = 0; spin < 1000000 && !got; ++spin) {
                            status = 0;
                            if (lane == 0) {
                                const unsigned long long old = atomicAdd(s_batch, 1ull);
                                const uint32_t off = (uint32_t)old & 0xFFFFFFu;
                                const unsigned long long gb = old >> 24;
                                if (gb < P.total_batches) {
                                    set = (uint32_t)(gb / P.batches_per_set);
                                    const uint32_t first_unit = (uint32_t)(gb - (unsigned long long)set * P.batches_per_set) * P.batch_units;
                                    const uint32_t left = P.local_units_per_set - first_unit;      /* units in this batch */
                                    if (off < (left < P.batch_units ? left : P.batch_units)) { j = first_unit + off; status = 1; }
                                }
                                if (status == 0) {
                                    if (sflag_get(s_ctl + 3)) status = 2;
                                    else if (SETDIRECT) {
                                        /* batch used up.  If no batch is left ANYWHERE there will be no further switch:
                                         * say so at once (the table keeps its base), so that this CTA's busy warps start
                                         * handing work to the idle ones instead of everyone parking behind the slowest.
                                         * Otherwise the CTA switches sets together. */
                                        if (ld_acquire_u64(&ctl->cursor) >= P.total_batches) {
                                            sflag_set(s_ctl + 3, 1);
                                            atomicMin(&ctl->t_exhausted, global_timer_ns());
                                            status = 2;
                                        } else status = 4;
                                    }
                                    else if ((spin & 1023) == 1023 && global_timer_ns() > *s_deadline) status = 3;
                                    else if (atomicCAS(const_cast<int*>(s_ctl) + 6, 0, 1) == 0) {
                                        /* batch used up: one warp of the CTA fetches the next one */
                                        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(s_batch);
                                        if ((cur >> 24) == gb) {      /* nobody has replaced it yet */
                                            const unsigned long long g = atomicAdd(&ctl->cursor, 1ull);
                                            if (g >= P.total_batches) { sflag_set(s_ctl + 3, 1); status = 2; }
                                            else atomicExch(s_batch, g << 24);
                                        }
                                        __threadfence_block();
                                        atomicExch(const_cast<int*>(s_ctl) + 6, 0);
                                    } else {
                                        __nanosleep(200);
                                    }
                                }
                            }
                            status = __shfl_sync(kFull, status, 0);
                            if (status == 1) {
                                set = __shfl_sync(kFull, set, 0);
                                j = __shfl_sync(kFull, j, 0);
                                got = true;
                            } else if (status == 3) {
                                watchdog_fire(P, GWARP, lane, 2, spin, 0, 0, 0, 0, 0);
                                break;
                            } else if (status == 2) {
                                break;
                            } else if (SETDIRECT && status == 4) {
                                wait_switch = true;
                                break;
                            }
                        }
                        if (!got && status == 0) {
                            /* a million rounds without an answer (a refill-lock holder that never returns): that is a
                             * failure, not "no seed units left" - abort the launch instead of dropping the batch's units */
                            watchdog_fire(P, GWARP, lane, 5, 0, 0, 0, 0, 0, 0);
                            break;
                        }
                        if (!got && status == 3) break;
                    }
                    if (SETDIRECT && wait_switch) {
                        /* the last nodes of the old set are expanded first (partial DIVIDE iterations below); only a
                         * warp without any node may wait for the switch */
                        if (ring_nodes(w) != 0u) divide_rest = true;
                        else {
                        if (!setdirect_rendezvous<WARPS>(P, s_ctl, s_hist, s_batch, lane, GWARP)) break;
                        set_base = (uint32_t)__shfl_sync(kFull, lane == 0 ? sflag_get(s_ctl + 7) : 0, 0);
                        continue;
                        }
                    }
                    if (!divide_rest) {
                    if (!got) {
                        if (lane == 0) { sflag_set(s_ctl + 3, 1); atomicMin(&ctl->t_exhausted, global_timer_ns()); }
                        __syncwarp();
                        continue;
                    }
                    /* units are handed out from the LAST one down: histograms are written in ascending value, high
                     * values may halve more often before reaching phi, so the biggest lineage trees start first and
                     * the run ends on the small ones (longest-processing-time-first; shortens the tail) */
#ifndef PROCELL_SEED_ORDER_ASC
                    j = P.local_units_per_set - 1u - j;
#endif
                    unsigned long long first = ((unsigned long long)j * P.shard_world + P.shard_rank) * P.unit;
                    unsigned long long last = first + P.unit;
                    if (last > P.n_cells) last = P.n_cells;
                    seed_cur = (uint32_t)first; seed_end = (uint32_t)last; seed_set = set;
                    seed_hint = 0xFFFFFFFFu;
                    if (seed_cur >= seed_end) { seed_cur = seed_end; continue; }
                    }
                }
                if (!divide_rest) {
                /* ---- SEED iteration: one seed cell per lane ---- */
                TRACE(P, GWARP, lane, 12);
                const uint32_t root = seed_cur + lane;
                const bool have = root < seed_end;
                seed_cur = (seed_end - seed_cur > 32u) ? seed_cur + 32u : seed_end;
                SeedOut so; so.kind = 0; so.key = 0; so.type = 0; so.kdiv = 0; so.t_div = 0.0; so.count0 = 0; so.quiescent = 0;
#ifdef PROCELL_LANE_BINSEARCH
                const uint32_t bin = find_bin(P, root);
#else
                const uint32_t bin = find_bin_warp(P, root - (uint32_t)lane, lane, seed_hint);
                seed_hint = __shfl_sync(kFull, bin, 31);
#endif
                if (have) {     /* PLAIN: one set, the tables are always the shared-memory copies */
                    const size_t tab = PLAIN ? 0u : (size_t)seed_set * P.n_types;
                    so = build_seed<SLOT>(P, s_log, root, seed_set, bin, (PLAIN ? s_thr_buf : s_thr + (size_t)seed_set * t_pad), (PLAIN ? s_sel_buf : s_sel) + tab,
                                    (PLAIN ? s_musd_buf : s_musd) + tab);
                }
                const unsigned live = __ballot_sync(kFull, so.kind == 2);
                if (so.kind == 2) {
                    const uint32_t pos = w.top + kSlot * __popc(live & lt_mask);
                    ring_store<RING>(w, pos, pcs_d2bits(so.t_div), 1ull, (uint64_t)root | ((uint64_t)so.key << 32),
                               (uint64_t)pack_dlo(seed_set, so.type, so.kdiv - 1u, 3u));
                }
                w.top += kSlot * __popc(live);
                __syncwarp();
                if (SETDIRECT) {
                    count_leaves_setdirect(P, s_hist, so.key, so.kind == 1 ? 1u : 0u, set_base);
                } else if (PLAIN || P.n_times == 1u) {
                    /* subtree sharding: every GPU builds every seed cell, GPU root % world counts its level-0 leaf */
                    const bool credit = !SUBTREE || root % P.sub_world == P.sub_rank;
                    warp_count_leaves<HASHED>(P, s_hist, so.key, (so.kind == 1 && credit) ? 1u : 0u);
                } else {        /* a seed cell exists from the start: out of time at every checkpoint before t_div */
                    for (uint32_t j = 0; j < P.n_times; ++j) {
                        const uint32_t inc = (have && so.count0 && (so.quiescent || P.times[j] < so.t_div)) ? 1u : 0u;
                        warp_count_leaves<HASHED>(P, s_hist, so.key + j * P.time_stride, inc);
                    }
                }
                continue;
                }
            }
            const uint32_t left = ring_nodes(w);
            if (left == 0u) {
                TRACE(P, GWARP, lane, 30);
                if (!idle_wait<RING>(w, P, s_ctl, *s_deadline, GWARP)) break;
                continue;
            }
            /* nothing to add: expand what the warp holds - its regular nodes first, then its last retry nodes */
            if (left != w.slow / kSlot) { mode = 1; take = left - w.slow / kSlot; }
            else { mode = 2; from_bottom = true; take = left; }
        }
        }

        ++iter;
        TRACE(P, GWARP, lane, 20);
        const uint32_t hist_base = SETDIRECT ? set_base : 0u;     /* changes only while this warp is parked in the rendezvous */
        /* PLAIN: one set and at most 64 types, so the (mean, sd) table is always the shared-memory copy (plain LDS) */
        const double2* musd = PLAIN ? s_musd_buf : s_musd;
        if (mode == 0) divide_fresh<WARPS, true, HASHED, PLAIN, RING, MODE>(w, P, s_log, s_hist, musd, 32u, lt_mask, multi_set, dc, hist_base);
        else if (mode == 1) divide_fresh<WARPS, false, HASHED, PLAIN, RING, MODE>(w, P, s_log, s_hist, musd, take, lt_mask, multi_set, dc, hist_base);
        else divide_general<HASHED, PLAIN, RING, MODE>(w, P, s_log, s_hist, musd, take, from_bottom, lt_mask, multi_set, dc, hist_base);

        /* hunger probe, every 16th iteration (every 8th costs 1 %, every 4th 3 % on config 2).  The CTA keeps a snapshot of "how many warps are starving", "how many
         * donated chunks are waiting" and "where is the seed cursor" in shared memory.  Every 64th iteration
         * (staggered by warp) one lane refreshes it with three 16-byte cp.async.cg copies straight from the control
         * block in HBM/L2 into shared memory: no registers, no waiting, the values simply turn up a little later.
         * Lane 0 reads the snapshot and decides, the warp follows (see RULE above). */
        if ((iter & kProbeMask) == 0u) {
            if ((iter & 255u) == 0u) {       /* every 256 iterations: flush the 32-bit division counters, check the watchdog */
                if (!multi_set) {
                    const uint32_t tot = (PLAIN ? 32u * 256u : 0u) + __reduce_add_sync(kFull, dc.cnt);
                    if (lane == 0 && tot) atomicAdd(reinterpret_cast<unsigned long long*>(P.divisions), (unsigned long long)tot);
                    dc.cnt = 0;
                } else if (dc.cnt > (1u << 30)) {
                    atomicAdd(reinterpret_cast<unsigned long long*>(P.divisions) + dc.set, (unsigned long long)dc.cnt);
                    dc.cnt = 0;
                }
                if (!HASHED && (iter & (kHistFlushIters - 1u)) == 0u) {
                    if (SETDIRECT) hist_drain_at(P, s_hist, lane, hist_base);
                    else hist_drain(P, s_hist, lane);
                }
                int late = 0;
                if (lane == 0) late = global_timer_ns() > *s_deadline;
                if (__shfl_sync(kFull, late, 0)) {
                    watchdog_fire(P, GWARP, lane, 1, ring_nodes(w), w.sp >> 16, seed_cur, seed_end, iter, 0);
                    break;
                }
            }
            int packed = 0;
            if (lane == 0) {
                int exhausted = sflag_get(s_ctl + 3);              /* one atomic read per probe */
                const bool endgame_now = kEndgameFast && exhausted && s_snap[0] >= kEndgameIdle;
                if (endgame_now || ((iter + (uint32_t)warp * (kProbeMask + 1u)) & kSnapMask) == 0u) {
                    cp_async16(s_snap, &ctl->idle);
                    cp_async16(s_snap + 4, &ctl->avail);
                    if (!multi_set && !exhausted) cp_async16(s_snap + 8, &ctl->cursor);
                    atomicAdd(const_cast<int*>(s_ctl) + 5, 1);      /* snapshot epoch */
                }
                const int idle_snap = s_snap[0], avail_snap = s_snap[4];
                if (!multi_set && !exhausted) {
                    const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(s_snap + 8);
                    if (cur >= P.total_local_units) { sflag_set(s_ctl + 3, 1); exhausted = 1; }
                }
                /* donate at most once per snapshot epoch, and only while fewer chunks wait than warps starve */
                const int ep = s_ctl[5];
                const int hg = P.donate && exhausted && idle_snap + kDonateReserve > (avail_snap > 0 ? avail_snap : 0) &&
                               avail_snap < kQueueCap / 2 && (ep != *s_epoch || (kEndgameFast && idle_snap >= kEndgameIdle));
                /* end game: when this many warps starve, a warp parts with a chunk as soon as it keeps 32 nodes */
                packed = (ep << 2) | ((idle_snap >= kEndgameIdle) << 1) | hg;
            }
            packed = __shfl_sync(kFull, packed, 0);
            if ((packed & 1) && ((w.top - w.bottom) / kSlot + 32u * (w.sp >> 16)) >= ((packed & 2) ? 64u : kDonateMinNodes * RING)) {
                /* somebody starves and no seeds are left: give away the shallowest chunk */
                TRACE(P, GWARP, lane, 50);
                if (lane == 0) *s_epoch = packed >> 2;
                donate_chunk<RING>(w, P);
            }
        }
    }

    asm volatile("cp.async.wait_all;" ::: "memory");
    TRACE(P, GWARP, lane, 60);
    if (lane == 0) atomicMax(&ctl->t_end, global_timer_ns());
    if (multi_set) {
        if (dc.cnt) atomicAdd(reinterpret_cast<unsigned long long*>(P.divisions) + dc.set, (unsigned long long)dc.cnt);
    } else {
        const uint32_t tot = (PLAIN ? 32u * (iter & 255u) : 0u) + __reduce_add_sync(kFull, dc.cnt);     /* <= 32 * 256 since the last periodic flush */
        if (lane == 0 && tot) atomicAdd(reinterpret_cast<unsigned long long*>(P.divisions), (unsigned long long)tot);
    }
    __syncthreads();
    if (HASHED) {
        const unsigned long long* tab = reinterpret_cast<const unsigned long long*>(s_hist);
        for (uint32_t i = threadIdx.x; i < P.smem_hist_slots; i += blockDim.x) {
            const unsigned long long e = tab[i];
            if ((uint32_t)e) atomicAdd(reinterpret_cast<unsigned long long*>(P.counts) + (uint32_t)(e >> 32), (unsigned long long)(uint32_t)e);
        }
    } else {
        const uint32_t flush_base = SETDIRECT ? (uint32_t)s_ctl[7] : 0u;
        for (uint32_t i = threadIdx.x; i < P.smem_hist_slots; i += blockDim.x) {
            uint32_t v = s_hist[i];
            if (v) atomicAdd(reinterpret_cast<unsigned long long*>(P.counts) + flush_base + (SLOT ? slot_key(P, i) : i), (unsigned long long)v);
        }
    }
    /* ---- fitness of the sweep in the same launch.  Every CTA of the grid is resident (one per SM), so they can meet:
     * a CTA arrives once its own flush is performed (fence), the last arrival releases everybody, and the sets are
     * shared out CTA-strided.  The slabs were written by L2 atomics moments ago - 104 MB for config 5, which fits the
     * 126 MB L2 - and are read back with ld.global.cg; 8 bytes per set leave the GPU instead of the tensor. */
    if (!PLAIN && P.fit_channels != 0u) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(&ctl->done_ctas, 1ull);
            while (ld_acquire_u64(&ctl->done_ctas) < (unsigned long long)gridDim.x) {
                __nanosleep(256);
                if (global_timer_ns() > *s_deadline) { atomicCAS(&ctl->status, kStatusOk, kStatusWatchdog); break; }
            }
        }
        __syncthreads();
        unsigned long long* s_acc = reinterpret_cast<unsigned long long*>(s_hist);     /* the table has been flushed: reuse it */
        const long long* slab0 = P.counts + (size_t)(P.n_times - 1u) * P.time_stride;   /* time series: the last checkpoint */
        for (uint32_t set = blockIdx.x; set < P.n_sets; set += gridDim.x)
            fitness_of_set<true>(slab0 + (size_t)set * P.n_keys * P.n_types, P.fit_key_channel, P.fit_target, P.n_keys, P.n_types,
                                 P.fit_channels, s_acc, P.fit_out + set);
    }
}

/* ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(kSimpleThreads) k_proliferate_simple(const __grid_constant__ SimParams P)
{
    __shared__ double s_log[kLogTabDoubles];
    for (int i = threadIdx.x; i < kLogTabDoubles; i += blockDim.x) s_log[i] = __ldg(P.logtab + i);
    __syncthreads();

    uint64_t st_heap[kSimpleStack];
    double st_t[kSimpleStack];
    uint32_t st_m[kSimpleStack];     /* mask | retry<<8 */
    const uint32_t T = P.n_types;
    const unsigned long long total = (unsigned long long)P.n_sets * P.n_cells;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(P.counts);

    for (unsigned long long gi = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += stride) {
        const uint32_t set = (uint32_t)(gi / P.n_cells);
        const uint32_t root = (uint32_t)(gi - (unsigned long long)set * P.n_cells);
        if (P.shard_world > 1u && (root / P.unit) % P.shard_world != P.shard_rank) continue;
        const size_t tab = (size_t)set * P.n_types;
        SeedOut so = build_seed<false>(P, s_log, root, set, find_bin(P, root), P.type_thr + (size_t)set * ((P.n_types + 3u) & ~3u), P.type_sel + tab,
                                       P.type_musd + tab);
        if (so.kind == 1) atomicAdd(counts + so.key, 1ull);
        if (so.kind != 2) continue;
        const double2 ms = __ldg(P.type_musd + (size_t)set * T + so.type);
        unsigned long long ndiv = 0;
        int sp = 0;
        st_heap[0] = 1ull; st_t[0] = so.t_div; st_m[0] = 3u; sp = 1;
        while (sp > 0) {
            --sp;
            const uint64_t heap = st_heap[sp];
            const double t_div = st_t[sp];
            uint32_t mask = st_m[sp] & 3u;
            const uint32_t retry = st_m[sp] >> 8;
            const uint32_t level = 63u - (uint32_t)__clzll((long long)heap);
            const bool forced = retry >= PCS_MAX_RETRY;
            pcs_u32x4 blk;
            blk.x = 0; blk.y = 0; blk.z = 0; blk.w = 0;
            if (!forced) blk = pcs_draw_rk(root, set, retry, PCS_TAG_DIVISION, heap, P.rk);
            if (retry == 0u) ++ndiv;
#pragma unroll 1
            for (uint32_t c = 0; c < 2u; ++c) {
                if (!(mask & (1u << c))) continue;
                double z = 0.0;
                const bool acc = forced || pcs_zig_trial(blk, c, &z, root, set, retry, heap, P.rk, s_log, P.logtab + PCS_TAB_WEDGE);
                const double timer = forced ? ms.x : pcs_timer(ms.x, ms.y, z);
                if (acc && (timer > 0.0 || forced)) {
                    mask &= ~(1u << c);
                    const double tc = PCS_ADD(t_div, timer);
                    if (tc > P.t_max) atomicAdd(counts + so.key + (level + 1u) * T, 1ull);
                    else if (level + 1u < so.kdiv) { st_heap[sp] = heap * 2ull + c; st_t[sp] = tc; st_m[sp] = 3u; ++sp; }
                }
            }
            if (mask) { st_heap[sp] = heap; st_t[sp] = t_div; st_m[sp] = mask | ((retry + 1u) << 8); ++sp; }
        }
        atomicAdd(reinterpret_cast<unsigned long long*>(P.divisions) + set, ndiv);
    }
}

/* ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) k_queue_init(unsigned long long* q_seq, ControlBlock* ctl, unsigned long long* counts,
                                                    size_t n_counts, unsigned long long* divisions, size_t n_divisions)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)kQueueCap) q_seq[i] = (unsigned long long)i;
    /* the whole control block, padding included (it is copied in 16-byte pieces and read back by the host) - except
     * the status word, which stays as it is: a failure of ANY run since the host last looked must still be there when
     * it looks again (procell_engine_finish reads and then clears it) */
    constexpr size_t kWords = sizeof(ControlBlock) / 8;
    if (i < kWords && i != offsetof(ControlBlock, status) / 8) {
        const bool ones = i == offsetof(ControlBlock, t_start) / 8 || i == offsetof(ControlBlock, t_exhausted) / 8;
        reinterpret_cast<unsigned long long*>(ctl)[i] = ones ? ~0ull : 0ull;
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = i; k < n_counts; k += stride) counts[k] = 0ull;
    for (size_t k = i; k < n_divisions; k += stride) divisions[k] = 0ull;
}

/* RNG-only ceiling: the per-division arithmetic of the common case (one Philox block, two fast ziggurat tests, two timers,
 * two time updates, four compares) with no tree, no stack and no atomics.  Three shapes of the same loop are built, and the
 * roofline is quoted against the FASTEST of them measured live (a ceiling that could be raised by reshaping the loop
 * would flatter the product kernel):
 *   variant 0: one chain per thread, 256-thread CTAs, 6 CTAs per SM (48 warps; the shape measured in round 1)
 *   variant 1: one chain per thread, register budget forced to 32 -> 8 CTAs per SM (64 warps, full occupancy)
 *   variant 2: two independent chains per thread, 4 CTAs per SM (32 warps, twice the instruction-level parallelism) */
struct RoundKeys { uint32_t rk[20]; };

template <int CHAINS>
__device__ __forceinline__ void rng_ceiling_body(int iters, const double* logtab, double mean, double sd, double t_max,
                                                 const RoundKeys& K, unsigned long long* sink)
{
    __shared__ double s_log[kLogTabDoubles];
    for (int i = threadIdx.x; i < kLogTabDoubles; i += blockDim.x) s_log[i] = __ldg(logtab + i);
    __syncthreads();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    double t[CHAINS];
    uint64_t heap[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { t[c] = 0.0; heap[c] = 1; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            /* the product kernel's forms of the block and of the fast test (draw_block, zig_fast_signed): the ceiling is the
             * best arithmetic we know, so whatever makes the product's arithmetic cheaper goes in here too */
            pcs_u32x4 blk = draw_block(tid, (uint32_t)c, 0u, PCS_TAG_DIVISION, heap[c], K.rk);
            double z0, z1;
            const bool f0 = zig_fast_signed(blk.x, blk.y, s_log + PCS_TAB_ZIG, &z0);
            const bool f1 = zig_fast_signed(blk.z, blk.w, s_log + PCS_TAB_ZIG, &z1);
            const double a = pcs_timer(mean, sd, z0), b = pcs_timer(mean, sd, z1);
            const double ta = PCS_ADD(t[c], a), tb = PCS_ADD(t[c], b);
            acc += (f0 && a > 0.0) + (f1 && b > 0.0) + (ta > t_max) + (tb > t_max);
            t[c] = (ta > t_max) ? 0.0 : ta;
            heap[c] = heap[c] * 2ull + (blk.x & 1u);
            if (heap[c] >> 62) heap[c] = 1;
        }
    }
    if (acc == 0xFFFFFFFFFFFFFFFFull) sink[0] = acc;
    atomicAdd(sink + 1, acc & 1ull);
}

__global__ void __launch_bounds__(256) k_rng_ceiling(int iters, const double* logtab, double mean, double sd, double t_max,
                                                     const __grid_constant__ RoundKeys K, unsigned long long* sink)
{
    rng_ceiling_body<1>(iters, logtab, mean, sd, t_max, K, sink);
}

__global__ void __launch_bounds__(256, 8) k_rng_ceiling_occ(int iters, const double* logtab, double mean, double sd, double t_max,
                                                            const __grid_constant__ RoundKeys K, unsigned long long* sink)
{
    rng_ceiling_body<1>(iters, logtab, mean, sd, t_max, K, sink);
}

__global__ void __launch_bounds__(256, 4) k_rng_ceiling_ilp2(int iters, const double* logtab, double mean, double sd, double t_max,
                                                             const __grid_constant__ RoundKeys K, unsigned long long* sink)
{
    rng_ceiling_body<2>(iters, logtab, mean, sd, t_max, K, sink);
}

/* ------------------------------------------------------------------------------------------------ host */
size_t coop_smem_bytes(int warps, int ring, uint32_t hist_slots, int hashed)
{
    return (size_t)kLogTabDoubles * 8 + kSmemCtlBytes + (size_t)warps * 4 * kStackCap * ring * 8 + (size_t)hist_slots * (hashed ? 8 : 4);
}

template <int WARPS, bool HASHED, bool PLAIN, int RING>
static cudaError_t coop_max_grid_t(int device, size_t smem_bytes, int* grid_out)
{
    cudaError_t e = cudaFuncSetAttribute(k_proliferate_coop<WARPS, HASHED, PLAIN, RING, kModeBase>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_proliferate_coop<WARPS, HASHED, PLAIN, RING, kModeBase>, WARPS * 32, smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    *grid_out = per_sm * sms;
    return cudaSuccess;
}

/* the 12 instances without the subtree-sharding rule: CTA shape (32 / 24 / 16 warps) x histogram mode x PLAIN */
#define COOP_DISPATCH(warps, ring, hashed, plain, X)                                                  \
    do {                                                                                             \
        switch (((warps) == 32 ? 0 : (warps) == 24 ? 4 : 8) + ((hashed) ? 2 : 0) + ((plain) ? 1 : 0)) { \
        case 0: X(32, false, false, 1); break;  case 1: X(32, false, true, 1); break;                \
        case 2: X(32, true, false, 1); break;   case 3: X(32, true, true, 1); break;                 \
        case 4: X(24, false, false, 1); break;  case 5: X(24, false, true, 1); break;                \
        case 6: X(24, true, false, 1); break;   case 7: X(24, true, true, 1); break;                 \
        case 8: X(16, false, false, 1); break;  case 9: X(16, false, true, 1); break;                \
        case 10: X(16, true, false, 1); break;  default: X(16, true, true, 1); break;                \
        }                                                                                            \
    } while (0)

cudaError_t coop_max_grid(int device, int warps, int ring, int hashed, int plain, size_t smem_bytes, int* grid_out)
{
    if (ring != 1) return cudaErrorInvalidValue;
#define X(W, H, PL, R) return coop_max_grid_t<W, H, PL, R>(device, smem_bytes, grid_out)
    COOP_DISPATCH(warps, ring, hashed, plain, X);
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t coop_max_grid_subtree(int device, int hashed, size_t smem_bytes, int* grid_out)
{
    int per_sm = 0, sms = 0;
    cudaError_t e = hashed ? cudaFuncSetAttribute(k_proliferate_coop<32, true, true, 1, kModeSubtree>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes)
                           : cudaFuncSetAttribute(k_proliferate_coop<32, false, true, 1, kModeSubtree>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    e = hashed ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_proliferate_coop<32, true, true, 1, kModeSubtree>, 1024, smem_bytes)
               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_proliferate_coop<32, false, true, 1, kModeSubtree>, 1024, smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    *grid_out = per_sm * sms;
    return cudaSuccess;
}

cudaError_t coop_max_grid_setdirect(int device, size_t smem_bytes, int* grid_out)
{
    int per_sm = 0, sms = 0;
    cudaError_t e = cudaFuncSetAttribute(k_proliferate_coop<32, false, false, 1, kModeSetDirect>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_proliferate_coop<32, false, false, 1, kModeSetDirect>, 1024, smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    *grid_out = per_sm * sms;
    return cudaSuccess;
}

cudaError_t coop_max_grid_merge(int device, size_t smem_bytes, int* grid_out)
{
    int per_sm = 0, sms = 0;
    cudaError_t e = cudaFuncSetAttribute(k_proliferate_coop<32, false, true, 1, kModeMerge>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_proliferate_coop<32, false, true, 1, kModeMerge>, 1024, smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    *grid_out = per_sm * sms;
    return cudaSuccess;
}

cudaError_t launch_coop(const SimParams& p, int warps, int ring, int grid, cudaStream_t stream)
{
    if (ring != 1) return cudaErrorInvalidValue;
    const size_t smem = coop_smem_bytes(warps, ring, p.smem_hist_slots, p.hist_hashed);
    const bool plain = coop_is_plain(p);
    if (p.hist_setdirect) {
        if (warps != 32 || ring != 1 || plain || p.hist_hashed || p.n_times != 1u || p.sub_world > 1u) return cudaErrorInvalidValue;
        k_proliferate_coop<32, false, false, 1, kModeSetDirect><<<grid, 1024, smem, stream>>>(p);
        return cudaGetLastError();
    }
    if (p.sub_world > 1u) {
        if (warps != 32 || ring != 1 || !plain) return cudaErrorInvalidValue;
        if (p.hist_hashed) k_proliferate_coop<32, true, true, 1, kModeSubtree><<<grid, 1024, smem, stream>>>(p);
        else k_proliferate_coop<32, false, true, 1, kModeSubtree><<<grid, 1024, smem, stream>>>(p);
        return cudaGetLastError();
    }
    if (p.leaf_merge) {
        if (warps != 32 || ring != 1 || !plain || p.hist_hashed) return cudaErrorInvalidValue;
        k_proliferate_coop<32, false, true, 1, kModeMerge><<<grid, 1024, smem, stream>>>(p);
        return cudaGetLastError();
    }
#define X(W, H, PL, R) k_proliferate_coop<W, H, PL, R, kModeBase><<<grid, W * 32, smem, stream>>>(p)
    COOP_DISPATCH(warps, ring, p.hist_hashed, plain, X);
#undef X
    return cudaGetLastError();
}

cudaError_t launch_simple(const SimParams& p, int grid, cudaStream_t stream)
{
    k_proliferate_simple<<<grid, kSimpleThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_queue_init(unsigned long long* q_seq, ControlBlock* ctl, long long* counts, size_t n_counts,
                              long long* divisions, size_t n_divisions, int sm_count, cudaStream_t stream)
{
    static_assert(offsetof(ControlBlock, status) % 8 == 0 && sizeof(ControlBlock) / 8 <= (size_t)kQueueCap, "control block layout");
    /* enough threads for the queue slots; for large tensors a grid-stride loop from 8 CTAs per SM */
    size_t blocks = (kQueueCap + 255) / 256;
    const size_t want = (n_counts + 255) / 256, cap = (size_t)(sm_count > 0 ? sm_count : 148) * 8;
    if (want > blocks) blocks = want < cap ? want : cap;
    k_queue_init<<<(unsigned)blocks, 256, 0, stream>>>(q_seq, ctl, reinterpret_cast<unsigned long long*>(counts), n_counts,
                                                      reinterpret_cast<unsigned long long*>(divisions), n_divisions);
    return cudaGetLastError();
}

int rng_ceiling_ctas_per_sm(int variant) { return variant == 1 ? 8 : variant == 2 ? 4 : 6; }
int rng_ceiling_chains(int variant) { return variant == 2 ? 2 : 1; }

cudaError_t launch_rng_ceiling(int variant, int grid, int block, int iters, const double* logtab, double mean, double sd,
                               double t_max, const uint32_t* rk, unsigned long long* sink, cudaStream_t stream)
{
    RoundKeys K;
    for (int i = 0; i < 20; ++i) K.rk[i] = rk[i];
    if (variant == 1) k_rng_ceiling_occ<<<grid, block, 0, stream>>>(iters, logtab, mean, sd, t_max, K, sink);
    else if (variant == 2) k_rng_ceiling_ilp2<<<grid, block, 0, stream>>>(iters, logtab, mean, sd, t_max, K, sink);
    else k_rng_ceiling<<<grid, block, 0, stream>>>(iters, logtab, mean, sd, t_max, K, sink);
    return cudaGetLastError();
}

}  // namespace procell_b200
