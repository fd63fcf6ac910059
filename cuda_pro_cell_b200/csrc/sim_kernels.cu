/* sim_kernels.cu - hand-written sm_100a kernels of the proliferation simulator.
 *
 * Replaces, in the reference (ericniso/cuda-pro-cell):
 *   device::create_cells_from_fluorescence / create_cells_population  (src/simulation/cells_population.cu:74-117)
 *   device::create_cell and helpers                                   (src/simulation/cell.cu:25-143)
 *   device::proliferate + the host level loop                         (src/simulation/proliferation.cu:26-402)
 * The reference materialises every tree level densely in HBM (N*2^d 40-byte cells) and recurses by one
 * dynamic-parallelism launch per block per level.  Here nothing but the count tensor ever reaches HBM:
 *
 * k_proliferate_coop - persistent, one CTA per SM, 16 warps.  Each warp owns a 128-node ring in shared
 *   memory.  An iteration is warp-uniform: either SEED (32 lanes build 32 seed cells: type pick, first timer,
 *   initial age; living roots are pushed) or DIVIDE (32 lanes pop the 32 newest = deepest nodes, each draws ONE
 *   Philox block -> one Box-Muller pair -> both daughters' timers, classifies the daughters as leaf / dropped /
 *   internal and pushes the internal ones).  While a warp holds fewer than 32 nodes every node is expanded each
 *   iteration (the breadth-first warm-up); from 32 on, expansion is depth-first, and the ring overflows in
 *   32-node chunks into a private spill ring in HBM.  Starving warps are fed through a bounded MPMC queue of
 *   chunks that busy warps fill from the BOTTOM of their stacks (the shallowest nodes = the largest subtrees).
 *   Leaves are counted in a shared-memory u32 histogram keyed (bin, k, type) with __match_any_sync aggregation;
 *   a wrap of a u32 slot carries 2^32 straight into the int64 tensor in HBM, the rest is flushed at the end.
 *
 * k_proliferate_simple - one thread per lineage with a local-memory stack and global atomics: the bring-up
 *   kernel, kept as an independent device-side cross-check of the cooperative one.
 */
#include "sim_kernels.h"

#include "procell_spec.h"

namespace procell_b200 {

namespace {

constexpr uint32_t kRingMask = kStackCap - 1;
constexpr unsigned kFull = 0xFFFFFFFFu;

/* node word D: set[0:16) | type[16:22) | kdiv[22:28) | pending-children mask[28:30) | retry[32:40) */
__device__ __forceinline__ uint64_t pack_d(uint32_t set, uint32_t type, uint32_t kdiv, uint32_t mask, uint32_t retry)
{
    return (uint64_t)(set | (type << 16) | (kdiv << 22) | (mask << 28)) | ((uint64_t)retry << 32);
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

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ int ld_volatile_s32(const int* p)
{
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

/* shared-memory privatised count; a u32 wrap carries 2^32 into the int64 tensor */
__device__ __forceinline__ void hist_add(const SimParams& P, uint32_t* s_hist, uint32_t key, uint32_t v)
{
    if (key < P.smem_hist_slots) {
        uint32_t old = atomicAdd(&s_hist[key], v);
        if (old + v < old) atomicAdd(reinterpret_cast<unsigned long long*>(P.counts) + key, 1ull << 32);
    } else {
        atomicAdd(reinterpret_cast<unsigned long long*>(P.counts) + key, (unsigned long long)v);
    }
}

/* every lane may carry `inc` (0..2) leaves for `key`; equal keys are merged, one atomic per distinct key */
__device__ __forceinline__ void warp_count_leaves(const SimParams& P, uint32_t* s_hist, uint32_t key, uint32_t inc)
{
    unsigned has = __ballot_sync(kFull, inc > 0);
    if (has == 0) return;
    unsigned two = __ballot_sync(kFull, inc == 2);
    if (inc > 0) {
        unsigned grp = __match_any_sync(has, key);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(grp) - 1))
            hist_add(P, s_hist, key, (uint32_t)(__popc(grp) + __popc(grp & two)));
    }
}

struct WarpCtx {
    uint64_t *sa, *sb, *sc, *sd;     /* ring fields: t_div bits, heap, root|keybase<<32, D */
    uint32_t bottom, top;            /* ring positions, n = top - bottom */
    unsigned long long* spill;       /* private spill ring */
    uint32_t sp_bottom, sp_top;
    int lane;
};

__device__ __forceinline__ void spill_bottom_chunk(WarpCtx& w, const SimParams& P)
{
    uint32_t idx = (w.bottom + w.lane) & kRingMask;
    unsigned long long* dst = w.spill + (size_t)(w.sp_top % kSpillCap) * kChunkWords;
    __stcg(dst + w.lane, w.sa[idx]);
    __stcg(dst + 32 + w.lane, w.sb[idx]);
    __stcg(dst + 64 + w.lane, w.sc[idx]);
    __stcg(dst + 96 + w.lane, w.sd[idx]);
    w.bottom += kChunkNodes;
    w.sp_top += 1;
    if (w.sp_top - w.sp_bottom > (uint32_t)kSpillCap && w.lane == 0) atomicExch(&P.ctl->status, kStatusSpillOverflow);
    __syncwarp();
}

__device__ __forceinline__ void unspill_newest_chunk(WarpCtx& w)
{
    w.sp_top -= 1;
    const unsigned long long* src = w.spill + (size_t)(w.sp_top % kSpillCap) * kChunkWords;
    w.bottom -= kChunkNodes;
    uint32_t idx = (w.bottom + w.lane) & kRingMask;
    w.sa[idx] = __ldcg(src + w.lane);
    w.sb[idx] = __ldcg(src + 32 + w.lane);
    w.sc[idx] = __ldcg(src + 64 + w.lane);
    w.sd[idx] = __ldcg(src + 96 + w.lane);
    __syncwarp();
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
    if (lane == 0) st_release_u64(P.q_seq + slot, pos + 1);
    return true;
}

/* returns true and the lane's node in (a,b,c,d) if a chunk was taken */
__device__ __forceinline__ bool queue_try_pop(const SimParams& P, int lane, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d)
{
    unsigned long long h = 0;
    int got = 0;
    if (lane == 0) {
        for (int attempt = 0; attempt < 8; ++attempt) {
            h = ld_volatile_u64(&P.ctl->q_head);
            unsigned long long t = ld_volatile_u64(&P.ctl->q_tail);
            if (h >= t) break;
            if (atomicCAS(&P.ctl->q_head, h, h + 1) == h) { got = 1; break; }
        }
    }
    got = __shfl_sync(kFull, got, 0);
    if (!got) return false;
    h = __shfl_sync(kFull, h, 0);
    uint32_t slot = (uint32_t)h & (kQueueCap - 1);
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
__device__ __forceinline__ void donate_chunk(WarpCtx& w, const SimParams& P)
{
    uint64_t a, b, c, d;
    if (w.sp_top != w.sp_bottom) {
        const unsigned long long* src = w.spill + (size_t)(w.sp_bottom % kSpillCap) * kChunkWords;
        a = __ldcg(src + w.lane);
        b = __ldcg(src + 32 + w.lane);
        c = __ldcg(src + 64 + w.lane);
        d = __ldcg(src + 96 + w.lane);
        w.sp_bottom += 1;
    } else {
        uint32_t idx = (w.bottom + w.lane) & kRingMask;
        a = w.sa[idx]; b = w.sb[idx]; c = w.sc[idx]; d = w.sd[idx];
        w.bottom += kChunkNodes;
    }
    __syncwarp();
    queue_push(P, w.lane, a, b, c, d);
}

/* A warp with nothing left: wait for donated work or for global quiescence.  true = a chunk was loaded. */
__device__ __forceinline__ bool idle_wait(WarpCtx& w, const SimParams& P)
{
    ControlBlock* ctl = P.ctl;
    if (w.lane == 0) {
        atomicAdd(&ctl->idle, 1);
        __threadfence();
        atomicSub(&ctl->active, 1);
    }
    unsigned long long t0 = global_timer_ns();
    for (;;) {
        int state = 0;   /* 0 wait, 1 try, 2 exit */
        if (w.lane == 0) {
            int act = ld_volatile_s32(&ctl->active);
            __threadfence();
            unsigned long long h = ld_volatile_u64(&ctl->q_head);
            unsigned long long t = ld_volatile_u64(&ctl->q_tail);
            int st = ld_volatile_s32(&ctl->status);
            if (st != kStatusOk) state = 2;
            else if (h < t) state = 1;
            else if (act == 0) state = 2;
            else if (global_timer_ns() - t0 > 120000000000ull) { atomicExch(&ctl->status, kStatusIdleTimeout); state = 2; }
        }
        state = __shfl_sync(kFull, state, 0);
        if (state == 2) {
            if (w.lane == 0) atomicSub(&ctl->idle, 1);
            return false;
        }
        if (state == 1) {
            if (w.lane == 0) atomicAdd(&ctl->active, 1);
            uint64_t a, b, c, d;
            if (queue_try_pop(P, w.lane, a, b, c, d)) {
                if (w.lane == 0) atomicSub(&ctl->idle, 1);
                uint32_t idx = (w.top + w.lane) & kRingMask;
                w.sa[idx] = a; w.sb[idx] = b; w.sc[idx] = c; w.sd[idx] = d;
                w.top += kChunkNodes;
                __syncwarp();
                return true;
            }
            if (w.lane == 0) { __threadfence(); atomicSub(&ctl->active, 1); }
        } else {
            __nanosleep(400);
        }
    }
}

/* result of building one seed cell (cell.cu:25-79 with type == -1, t == 0) */
struct SeedOut {
    uint32_t keybase;   /* ((set*n_keys + bin_keybase) * n_types + type) */
    uint32_t type, kdiv;
    int kind;           /* 0 dropped, 1 leaf at level 0, 2 living root */
    double t_div;
};

__device__ __forceinline__ SeedOut build_seed(const SimParams& P, const double* s_log, uint32_t root, uint32_t set)
{
    SeedOut o;
    /* bin of this seed cell: largest b with bin_start[b] <= root (parser.cu "bounds") */
    uint32_t lo = 0, hi = P.n_bins;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(P.bin_start + mid) <= root) lo = mid; else hi = mid;
    }
    const uint32_t kd = __ldg(P.bin_kdiv + lo);
    const uint32_t T = P.n_types;
    pcs_u32x4 w = pcs_draw(root, set, 0u, PCS_TAG_SEED, 0ull, P.key0, P.key1);
    const double u_type = pcs_u53(w.x, w.y);
    const double u_age = pcs_u53(w.z, w.w);
    uint32_t j = 0;
    for (; j + 1 < T; ++j)                                     /* cell.cu:81-104; Q17: none -> last */
        if (u_type < __ldg(P.type_cum + (size_t)set * T + j)) break;
    const uint32_t type = __ldg(P.type_sel + (size_t)set * T + j);
    const double2 ms = __ldg(P.type_musd + (size_t)set * T + type);
    o.type = type;
    o.kdiv = kd & 63u;
    o.keybase = (set * P.n_keys + __ldg(P.bin_keybase + lo)) * T + type;
    const bool count0 = (kd & 0x80u) != 0u;
    if (ms.x < 0.0) {                                          /* quiescent: timer -1, t 0 -> out_of_time */
        o.kind = count0 ? 1 : 0;
        o.t_div = 0.0;
        return o;
    }
    double timer = ms.x;
    for (uint32_t retry = 0; retry < PCS_MAX_RETRY; ++retry) { /* root = child 1 of the virtual division at heap 0 */
        pcs_u32x4 b = pcs_draw(root, set, retry, PCS_TAG_DIVISION, 0ull, P.key0, P.key1);
        double z0, z1;
        pcs_normal_pair(b, s_log, (P.refcompat && retry == 0u) ? u_type : 0.0, &z0, &z1);
        double cand = pcs_timer(ms.x, ms.y, z1);
        if (cand > 0.0) { timer = cand; break; }
    }
    const double t0 = PCS_MUL(timer, u_age);                   /* cell.cu:124-143 */
    const double t_div = PCS_ADD(t0, timer);
    o.t_div = t_div;
    if (t_div > P.t_max) o.kind = count0 ? 1 : 0;              /* proliferation.cu:404-410 */
    else o.kind = (o.kdiv > 0u) ? 2 : 0;                       /* f/2 <= phi: dropped (Q6) */
    return o;
}

}  // namespace

__global__ void __launch_bounds__(kCoopThreads, 1) k_proliferate_coop(const SimParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_log = reinterpret_cast<double*>(smem_raw);
    uint64_t* s_stack = reinterpret_cast<uint64_t*>(smem_raw + kLogTabDoubles * 8);
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw + kLogTabDoubles * 8 + (size_t)kCoopWarps * 4 * kStackCap * 8);

    for (int i = threadIdx.x; i < kLogTabDoubles; i += blockDim.x) s_log[i] = __ldg(P.logtab + i);
    for (uint32_t i = threadIdx.x; i < P.smem_hist_slots; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t T = P.n_types;
    ControlBlock* ctl = P.ctl;

    WarpCtx w;
    w.sa = s_stack + (size_t)warp * 4 * kStackCap;
    w.sb = w.sa + kStackCap;
    w.sc = w.sb + kStackCap;
    w.sd = w.sc + kStackCap;
    w.bottom = 0; w.top = 0;
    w.spill = P.spill + (size_t)(blockIdx.x * kCoopWarps + warp) * kSpillCap * kChunkWords;
    w.sp_bottom = 0; w.sp_top = 0;
    w.lane = lane;

    uint32_t seed_cur = 0, seed_end = 0, seed_set = 0;
    bool seeds_left = P.total_local_units > 0;
    uint32_t div_set = 0;
    unsigned long long div_cnt = 0;
    uint32_t iter = 0;

    if (lane == 0) atomicAdd(&ctl->active, 1);

    for (;;) {
        uint32_t n = w.top - w.bottom;
        if (n < 32u) {
            if (w.sp_top != w.sp_bottom) { unspill_newest_chunk(w); continue; }
            if (seeds_left) {
                if (seed_cur == seed_end) {
                    unsigned long long c = 0;
                    if (lane == 0) c = atomicAdd(&ctl->cursor, 1ull);
                    c = __shfl_sync(kFull, c, 0);
                    if (c >= P.total_local_units) { seeds_left = false; continue; }
                    uint32_t set = (uint32_t)(c / P.local_units_per_set);
                    uint32_t j = (uint32_t)(c - (unsigned long long)set * P.local_units_per_set);
                    unsigned long long first = ((unsigned long long)j * P.shard_world + P.shard_rank) * P.unit;
                    unsigned long long last = first + P.unit;
                    if (last > P.n_cells) last = P.n_cells;
                    seed_cur = (uint32_t)first; seed_end = (uint32_t)last; seed_set = set;
                    if (seed_cur >= seed_end) { seed_cur = seed_end; continue; }
                }
                /* ---- SEED iteration: one seed cell per lane ---- */
                const uint32_t root = seed_cur + lane;
                const bool have = root < seed_end;
                seed_cur = (seed_end - seed_cur > 32u) ? seed_cur + 32u : seed_end;
                SeedOut so; so.kind = 0; so.keybase = 0; so.type = 0; so.kdiv = 0; so.t_div = 0.0;
                if (have) so = build_seed(P, s_log, root, seed_set);
                const unsigned live = __ballot_sync(kFull, so.kind == 2);
                if (so.kind == 2) {
                    uint32_t idx = (w.top + __popc(live & lt_mask)) & kRingMask;
                    w.sa[idx] = pcs_d2bits(so.t_div);
                    w.sb[idx] = 1ull;
                    w.sc[idx] = (uint64_t)root | ((uint64_t)so.keybase << 32);
                    w.sd[idx] = pack_d(seed_set, so.type, so.kdiv, 3u, 0u);
                }
                w.top += __popc(live);
                __syncwarp();
                warp_count_leaves(P, s_hist, so.keybase, so.kind == 1 ? 1u : 0u);
                continue;
            }
            if (n == 0u) {
                if (!idle_wait(w, P)) break;
                continue;
            }
        }
        if (n > (uint32_t)(kStackCap - 32)) { spill_bottom_chunk(w, P); continue; }

        /* hunger probe (load issued now, consumed after the math) */
        ++iter;
        int probe_idle = 0;
        const bool probe = !seeds_left && (iter & 3u) == 0u && (n + 32u * (w.sp_top - w.sp_bottom)) >= 96u;
        if (probe && lane == 0) probe_idle = ld_volatile_s32(&ctl->idle);

        /* ---- DIVIDE iteration: one node per lane, newest first ---- */
        const uint32_t take = n < 32u ? n : 32u;
        const bool act = (uint32_t)lane < take;
        uint32_t npush = 0, leaf_inc = 0, leaf_key = 0;
        uint64_t pa0 = 0, pb0 = 0, pd0 = 0, pa1 = 0, pb1 = 0, pd1 = 0, pc = 0;
        if (act) {
            const uint32_t idx = (w.top - 1u - (uint32_t)lane) & kRingMask;
            const double t_div = pcs_bits2d(w.sa[idx]);
            const uint64_t heap = w.sb[idx];
            pc = w.sc[idx];
            const uint64_t d = w.sd[idx];
            const uint32_t root = (uint32_t)pc;
            const uint32_t keybase = (uint32_t)(pc >> 32);
            const uint32_t dl = (uint32_t)d;
            const uint32_t set = dl & 0xFFFFu;
            const uint32_t type = (dl >> 16) & 63u;
            const uint32_t kdiv = (dl >> 22) & 63u;
            uint32_t mask = (dl >> 28) & 3u;
            const uint32_t retry = (uint32_t)(d >> 32) & 0xFFu;
            const double2 ms = __ldg(P.type_musd + (size_t)set * T + type);
            const uint32_t level = 63u - (uint32_t)__clzll((long long)heap);
            const bool forced = retry >= PCS_MAX_RETRY;
            double z0 = 0.0, z1 = 0.0;
            if (!forced) {
                pcs_u32x4 blk = pcs_draw(root, set, retry, PCS_TAG_DIVISION, heap, P.key0, P.key1);
                pcs_normal_pair(blk, s_log, 0.0, &z0, &z1);
            }
            if (retry == 0u) {
                if (set != div_set) {
                    if (div_cnt) atomicAdd(reinterpret_cast<unsigned long long*>(P.divisions) + div_set, div_cnt);
                    div_cnt = 0; div_set = set;
                }
                div_cnt += 1;
            }
            const bool deeper = level + 1u < kdiv;
            leaf_key = keybase + (level + 1u) * T;
            if (mask & 1u) {
                const double timer = pcs_timer(ms.x, ms.y, z0);
                if (timer > 0.0 || forced) {
                    mask &= ~1u;
                    const double tc = PCS_ADD(t_div, timer);
                    if (tc > P.t_max) leaf_inc += 1;
                    else if (deeper) { pa0 = pcs_d2bits(tc); pb0 = heap * 2ull; pd0 = pack_d(set, type, kdiv, 3u, 0u); npush = 1; }
                }
            }
            if (mask & 2u) {
                const double timer = pcs_timer(ms.x, ms.y, z1);
                if (timer > 0.0 || forced) {
                    mask &= ~2u;
                    const double tc = PCS_ADD(t_div, timer);
                    if (tc > P.t_max) leaf_inc += 1;
                    else if (deeper) {
                        const uint64_t na = pcs_d2bits(tc), nb = heap * 2ull + 1ull, nd = pack_d(set, type, kdiv, 3u, 0u);
                        if (npush == 0) { pa0 = na; pb0 = nb; pd0 = nd; } else { pa1 = na; pb1 = nb; pd1 = nd; }
                        npush += 1;
                    }
                }
            }
            if (mask) {   /* a daughter's timer was <= 0: redraw it in a later iteration (cell.cu:114-118) */
                const uint64_t na = pcs_d2bits(t_div), nd = pack_d(set, type, kdiv, mask, retry + 1u);
                if (npush == 0) { pa0 = na; pb0 = heap; pd0 = nd; } else { pa1 = na; pb1 = heap; pd1 = nd; }
                npush += 1;
            }
        }
        w.top -= take;
        const unsigned b0 = __ballot_sync(kFull, npush >= 1u);
        const unsigned b1 = __ballot_sync(kFull, npush == 2u);
        if (npush >= 1u) {
            uint32_t idx = (w.top + __popc(b0 & lt_mask)) & kRingMask;
            w.sa[idx] = pa0; w.sb[idx] = pb0; w.sc[idx] = pc; w.sd[idx] = pd0;
        }
        if (npush == 2u) {
            uint32_t idx = (w.top + __popc(b0) + __popc(b1 & lt_mask)) & kRingMask;
            w.sa[idx] = pa1; w.sb[idx] = pb1; w.sc[idx] = pc; w.sd[idx] = pd1;
        }
        w.top += __popc(b0) + __popc(b1);
        __syncwarp();
        warp_count_leaves(P, s_hist, leaf_key, leaf_inc);

        if (probe) {
            int want = 0;
            if (lane == 0 && probe_idle > 0) {
                unsigned long long h = ld_volatile_u64(&ctl->q_head);
                unsigned long long t = ld_volatile_u64(&ctl->q_tail);
                want = (t - h) < (unsigned long long)probe_idle && (t - h) < (unsigned long long)(kQueueCap / 2);
            }
            want = __shfl_sync(kFull, want, 0);
            if (want && (w.top - w.bottom + 32u * (w.sp_top - w.sp_bottom)) >= 64u) donate_chunk(w, P);
        }
    }

    if (div_cnt) atomicAdd(reinterpret_cast<unsigned long long*>(P.divisions) + div_set, div_cnt);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P.smem_hist_slots; i += blockDim.x) {
        uint32_t v = s_hist[i];
        if (v) atomicAdd(reinterpret_cast<unsigned long long*>(P.counts) + i, (unsigned long long)v);
    }
}

/* ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(kSimpleThreads) k_proliferate_simple(const SimParams P)
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
        SeedOut so = build_seed(P, s_log, root, set);
        if (so.kind == 1) atomicAdd(counts + so.keybase, 1ull);
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
            double z[2] = { 0.0, 0.0 };
            if (!forced) {
                pcs_u32x4 blk = pcs_draw(root, set, retry, PCS_TAG_DIVISION, heap, P.key0, P.key1);
                pcs_normal_pair(blk, s_log, 0.0, &z[0], &z[1]);
            }
            if (retry == 0u) ++ndiv;
#pragma unroll
            for (uint32_t c = 0; c < 2u; ++c) {
                if (!(mask & (1u << c))) continue;
                const double timer = pcs_timer(ms.x, ms.y, z[c]);
                if (timer > 0.0 || forced) {
                    mask &= ~(1u << c);
                    const double tc = PCS_ADD(t_div, timer);
                    if (tc > P.t_max) atomicAdd(counts + so.keybase + (level + 1u) * T, 1ull);
                    else if (level + 1u < so.kdiv) { st_heap[sp] = heap * 2ull + c; st_t[sp] = tc; st_m[sp] = 3u; ++sp; }
                }
            }
            if (mask) { st_heap[sp] = heap; st_t[sp] = t_div; st_m[sp] = mask | ((retry + 1u) << 8); ++sp; }
        }
        atomicAdd(reinterpret_cast<unsigned long long*>(P.divisions) + set, ndiv);
    }
}

/* ------------------------------------------------------------------------------------------------ */
__global__ void k_queue_init(unsigned long long* q_seq, ControlBlock* ctl)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kQueueCap) q_seq[i] = (unsigned long long)i;
    if (i == 0) {
        ctl->cursor = 0; ctl->q_head = 0; ctl->q_tail = 0; ctl->active = 0; ctl->idle = 0; ctl->status = 0;
    }
}

/* RNG-only ceiling: the per-division arithmetic (one Philox block, one Box-Muller pair, two timers, two time
 * updates, four compares) with no tree, no stack and no atomics */
__global__ void __launch_bounds__(256) k_rng_ceiling(int iters, const double* logtab, double mean, double sd, double t_max,
                                                     uint32_t key0, uint32_t key1, unsigned long long* sink)
{
    __shared__ double s_log[kLogTabDoubles];
    for (int i = threadIdx.x; i < kLogTabDoubles; i += blockDim.x) s_log[i] = __ldg(logtab + i);
    __syncthreads();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    double t = 0.0;
    uint64_t heap = 1;
    for (int i = 0; i < iters; ++i) {
        pcs_u32x4 blk = pcs_draw(tid, 0u, 0u, PCS_TAG_DIVISION, heap, key0, key1);
        double z0, z1;
        pcs_normal_pair(blk, s_log, 0.0, &z0, &z1);
        const double a = pcs_timer(mean, sd, z0), b = pcs_timer(mean, sd, z1);
        const double ta = PCS_ADD(t, a), tb = PCS_ADD(t, b);
        acc += (a > 0.0) + (b > 0.0) + (ta > t_max) + (tb > t_max);
        t = (ta > t_max) ? 0.0 : ta;
        heap = heap * 2ull + (blk.x & 1u);
        if (heap >> 62) heap = 1;
    }
    if (acc == 0xFFFFFFFFFFFFFFFFull) sink[0] = acc;
    atomicAdd(sink + 1, acc & 1ull);
}

/* ------------------------------------------------------------------------------------------------ host */
size_t coop_smem_bytes(uint32_t hist_slots)
{
    return (size_t)kLogTabDoubles * 8 + (size_t)kCoopWarps * 4 * kStackCap * 8 + (size_t)hist_slots * 4;
}

cudaError_t coop_max_grid(int device, size_t smem_bytes, int* grid_out)
{
    cudaError_t e = cudaFuncSetAttribute(k_proliferate_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_proliferate_coop, kCoopThreads, smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    *grid_out = per_sm * sms;
    return cudaSuccess;
}

cudaError_t launch_coop(const SimParams& p, int grid, cudaStream_t stream)
{
    size_t smem = coop_smem_bytes(p.smem_hist_slots);
    cudaError_t e = cudaFuncSetAttribute(k_proliferate_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_proliferate_coop<<<grid, kCoopThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_simple(const SimParams& p, int grid, cudaStream_t stream)
{
    k_proliferate_simple<<<grid, kSimpleThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_queue_init(unsigned long long* q_seq, ControlBlock* ctl, cudaStream_t stream)
{
    k_queue_init<<<(kQueueCap + 255) / 256, 256, 0, stream>>>(q_seq, ctl);
    return cudaGetLastError();
}

cudaError_t launch_rng_ceiling(int grid, int block, int iters, const double* logtab, double mean, double sd,
                               double t_max, uint32_t key0, uint32_t key1, unsigned long long* sink,
                               cudaStream_t stream)
{
    k_rng_ceiling<<<grid, block, 0, stream>>>(iters, logtab, mean, sd, t_max, key0, key1, sink);
    return cudaGetLastError();
}

}  // namespace procell_b200
