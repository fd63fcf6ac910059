/* procell_spec.h - the arithmetic specification of the B200 proliferation simulator.
 *
 * Everything that decides WHICH histogram comes out lives here: the Philox4x32-10 stream
 * layout, the bits->uniform maps, the ziggurat sampler of the division timers, the fixed-operation-sequence
 * FP64 log / sincos behind it and behind the seed cell's Box-Muller transform, and the node rule.  The sm_100a kernels (sim_kernels.cu) include this
 * header; the CPU oracle (oracle/procell_oracle.c) restates it independently in plain C.
 *
 * Replaces, in the reference (ericniso/cuda-pro-cell):
 *   src/utils/util.cu:143-169     init_random / uniform_random / normal_random (cuRAND XORWOW,
 *                                 a fresh curand_init per draw) -> counter-based Philox keyed by
 *                                 (root cell, tree path), one block per DIVISION (both daughters).
 *                                 The standard normal behind a division timer is drawn with the ziggurat
 *                                 method (Marsaglia & Tsang 2000; exact in law): 99.2 % of the draws are one
 *                                 table row, one fma and one compare; the seed cell's first timer keeps the
 *                                 Box-Muller transform (the refcompat coupling of SURVEY Q1 is defined on it).
 *   src/simulation/cell.cu:106-143 determine_cell_timer / determine_cell_initial_t
 *   src/simulation/proliferation.cu:321-350,404-410  node rule / out_of_time
 *
 * Bit-reproducibility: only IEEE-754 correctly-rounded primitives are used (add, mul, fma,
 * sqrt) in a fixed order, through intrinsics the compiler may not contract or reassociate.
 */
#ifndef PROCELL_SPEC_H
#define PROCELL_SPEC_H

#include <stdint.h>
#include <string.h>
#include "procell_math_tables.inc"

#if defined(__CUDACC__)
#define PCS_HD __host__ __device__ __forceinline__
#else
#define PCS_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define PCS_ADD(a, b) __dadd_rn((a), (b))
#define PCS_MUL(a, b) __dmul_rn((a), (b))
#define PCS_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PCS_SQRT(a) __dsqrt_rn((a))
#else
/* host build: compile with -ffp-contract=off (the Makefile does) so a*b+c is never fused */
#define PCS_ADD(a, b) ((a) + (b))
#define PCS_MUL(a, b) ((a) * (b))
#define PCS_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define PCS_SQRT(a) __builtin_sqrt((a))
#endif

PCS_HD double pcs_bits2d(uint64_t u);

/* scalar coefficients: on the device they are read from a __constant__ table, so FP64 instructions take them
 * as constant-bank operands instead of rebuilding 64-bit immediates in registers every iteration */
#if defined(__CUDACC__)
static __constant__ double pcs_coef[PCM_N_COEF] = { PCM_COEF_HEXFLOATS };
#endif
#if defined(__CUDA_ARCH__)
#define PCS_C(NAME) (pcs_coef[PCM_IDX_##NAME])
#else
#define PCS_C(NAME) (pcs_bits2d(PCM_BITS_##NAME))
#endif

PCS_HD double pcs_bits2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}

PCS_HD uint64_t pcs_d2bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}

/* ------------------------------------------------------------------ limits of the key layout */
#define PCS_MAX_LEVEL 63u        /* heap index of a node at level k lies in [2^k, 2^(k+1)) */
#define PCS_MAX_RETRY 255u       /* retry field is 8 bits; at 255 the timer falls back to the mean */
#define PCS_MAX_SETS 65536u      /* parameter-set id is 16 bits */
#define PCS_MAX_TYPES 64u
#define PCS_MAX_BINS 65535u
#define PCS_TAG_DIVISION 0u      /* one block per division: both daughters' timers */
#define PCS_TAG_ZIGX 2u          /* extra uniforms of a ziggurat trial that left the fast path: wedge test, or attempt k of
                                    the tail sampler with tag 2 + k */
#define PCS_ZIG_TAIL_TRIES 200u  /* tags 2 .. 201; the tail sampler accepts 94 % of its attempts */
#define PCS_TAG_SEED 1u          /* ONE block per seed cell and draw round: word x = type uniform, y = initial-age uniform
                                    (32-bit grade each - what cuRAND's float generators use), words z, w = the normal of
                                    its first timer: ideal seeding makes one ziggurat trial on the 64 bits, as daughter 1
                                    of a division at tree path 0 would (extra uniforms: blocks tagged 2.. of (root, path 0,
                                    retry)); refcompat seeding a Box-Muller draw with z the radius uniform and w the angle.
                                    A rejected first timer (trial rejected, or <= 0) is redrawn from words z, w of the
                                    block with retry + 1; x and y are taken from round 0 only. */

/* ------------------------------------------------------------------ Philox4x32-10 (Salmon et al., SC'11) */
#define PCS_PHILOX_M0 0xD2511F53u
#define PCS_PHILOX_M1 0xCD9E8D57u
#define PCS_PHILOX_W0 0x9E3779B9u
#define PCS_PHILOX_W1 0xBB67AE85u

struct pcs_u32x4 { uint32_t x, y, z, w; };

PCS_HD pcs_u32x4 pcs_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                   uint32_t k0, uint32_t k1)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PCS_PHILOX_M0 * (uint64_t)c0;
        uint64_t p1 = (uint64_t)PCS_PHILOX_M1 * (uint64_t)c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += PCS_PHILOX_W0;
        k1 += PCS_PHILOX_W1;
    }
    pcs_u32x4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

/* the same ten rounds with the 20 round keys precomputed (rk[2r] = k0 + r*W0, rk[2r+1] = k1 + r*W1): on the
 * device rk points into the kernel parameter block, so every key is a constant-bank operand of its LOP3 */
PCS_HD void pcs_round_keys(uint32_t k0, uint32_t k1, uint32_t* rk)
{
    for (int r = 0; r < 10; ++r) {
        rk[2 * r] = k0;
        rk[2 * r + 1] = k1;
        k0 += PCS_PHILOX_W0;
        k1 += PCS_PHILOX_W1;
    }
}

PCS_HD pcs_u32x4 pcs_philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* rk)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PCS_PHILOX_M0 * (uint64_t)c0;
        uint64_t p1 = (uint64_t)PCS_PHILOX_M1 * (uint64_t)c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk[2 * r];
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    pcs_u32x4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

PCS_HD pcs_u32x4 pcs_draw_rk(uint32_t root, uint32_t set, uint32_t retry, uint32_t tag, uint64_t heap,
                             const uint32_t* rk)
{
    return pcs_philox4x32_10_rk(root, set | (retry << 16) | (tag << 24), (uint32_t)heap, (uint32_t)(heap >> 32), rk);
}

/* counter layout: c0 = root cell id, c1 = set | retry<<16 | tag<<24, (c3:c2) = heap index */
PCS_HD pcs_u32x4 pcs_draw(uint32_t root, uint32_t set, uint32_t retry, uint32_t tag, uint64_t heap,
                          uint32_t k0, uint32_t k1)
{
    return pcs_philox4x32_10(root, set | (retry << 16) | (tag << 24), (uint32_t)heap,
                             (uint32_t)(heap >> 32), k0, k1);
}

/* ------------------------------------------------------------------ bits -> uniform in (0,1) */
/* 52 random mantissa bits m -> (2m+1) * 2^-53, exact: never 0, never 1 */
PCS_HD double pcs_unit_from_mant52(uint64_t mant52)
{
    double d = pcs_bits2d(0x3FF0000000000000ULL | mant52);      /* [1,2) */
    return PCS_ADD(d, -PCS_C(ONE_M));              /* exact */
}

PCS_HD double pcs_u53(uint32_t lo, uint32_t hi)
{
    return pcs_unit_from_mant52((((uint64_t)hi << 32) | (uint64_t)lo) >> 12);
}

/* 32 random bits m -> (2m+1) * 2^-33, exact: never 0, never 1 (the seed-cell draws) */
PCS_HD double pcs_u32unit(uint32_t m)
{
    return PCS_FMA((double)m, 0x1p-32, 0x1p-33);    /* both constants have an all-zero low word: cheap immediates */
}

/* the math table every consumer passes as `tab`: 128 log rows {invc, logc}, then 256 sin/cos rows {sin, cos}, then 512
 * ziggurat rows {x_i, x_i+1} (the kernels keep these three in shared memory), then 512 wedge rows {f(x_i), f(x_i+1) - f(x_i)}
 * that only the rare wedge test reads (global memory on the device) */
#define PCS_TAB_SINCOS (2 << PCM_LOG_N_BITS)
#define PCS_TAB_ZIG (PCS_TAB_SINCOS + (2 << PCM_SC_N_BITS))
#define PCS_TAB_DOUBLES (PCS_TAB_ZIG + (2 << PCM_ZIG_N_BITS))          /* the shared-memory part */
#define PCS_TAB_WEDGE PCS_TAB_DOUBLES
#define PCS_TAB_ALL_DOUBLES (PCS_TAB_WEDGE + (2 << PCM_ZIG_N_BITS))

/* ------------------------------------------------------------------ -2*ln(u), u in (0,1), normal double
 * tab: 128 rows {invc, logc} as 256 doubles (a shared-memory copy on the device).
 * log(u) = k*ln2 + logc_i + log1p(r),  r = z*invc_i - 1,  |r| <= 2^-8,  log1p by Taylor to r^6. */
PCS_HD double pcs_neg2log(double u, const double* tab)
{
    uint64_t ix = pcs_d2bits(u);
    uint64_t tmp = ix - PCM_LOG_OFF;
    uint32_t i = (uint32_t)(tmp >> (52 - PCM_LOG_N_BITS)) & ((1u << PCM_LOG_N_BITS) - 1u);
    int32_t k = (int32_t)((int64_t)tmp >> 52);
    double z = pcs_bits2d(ix - (tmp & 0xFFF0000000000000ULL));
    double invc = tab[2 * i];
    double logc = tab[2 * i + 1];
    double r = PCS_FMA(z, invc, -1.0);
    double q = PCS_C(LOG1P_B6);
    q = PCS_FMA(q, r, PCS_C(LOG1P_B5));
    q = PCS_FMA(q, r, PCS_C(LOG1P_B4));
    q = PCS_FMA(q, r, PCS_C(LOG1P_B3));
    q = PCS_FMA(q, r, PCS_C(LOG1P_B2));
    double r2 = PCS_MUL(r, r);
    double l = PCS_FMA(r2, q, r);
    double base = PCS_FMA((double)k, PCS_C(LN2), logc);
    double lg = PCS_ADD(base, l);
    return PCS_MUL(lg, -2.0);
}

/* ------------------------------------------------------------------ sin/cos(2*pi*V/2^64) from 64 random bits
 * sector j = top 8 bits (256 sectors of pi/128); the next 52 bits d in [1,2) give the offset from the sector's
 * CENTRE, delta = fma(d, pi/128, -(3/2 - 2^-53) * pi/128) in (-pi/256, pi/256), never 0.  With the table row
 * {S, C} = {sin c_j, cos c_j} of the centre angle:
 *   sin(c+delta) = S + (S*(cos delta - 1) + C*sin delta),  cos(c+delta) = C + (C*(cos delta - 1) - S*sin delta)
 * sin delta to delta^5 and cos delta - 1 to delta^6 (truncation 8e-18 and 1e-20): 12 FP64 operations and one
 * 16-byte table load instead of 21 operations and the octant selects of a pi/4 reduction.
 * tab_sc: 256 rows {sin, cos} as 512 doubles (a shared-memory copy on the device). */
PCS_HD void pcs_sincos2pi(uint64_t v, const double* tab_sc, double* s_out, double* c_out)
{
    uint32_t j = (uint32_t)(v >> (64 - PCM_SC_N_BITS));
    uint64_t m = (v >> (12 - PCM_SC_N_BITS)) & 0x000FFFFFFFFFFFFFULL;
    double d = pcs_bits2d(0x3FF0000000000000ULL | m);            /* [1,2) */
    double dl = PCS_FMA(d, PCS_C(SC_A), PCS_C(SC_B));
    double d2 = PCS_MUL(dl, dl);
    double ps = PCS_FMA(PCS_C(SD_S2), d2, PCS_C(SD_S1));
    double d3 = PCS_MUL(dl, d2);
    double sd = PCS_FMA(d3, ps, dl);                            /* sin delta */
    double pc = PCS_FMA(PCS_C(CM_C3), d2, PCS_C(CM_C2));
    pc = PCS_FMA(pc, d2, PCS_C(CM_C1));
    double cm = PCS_MUL(pc, d2);                                /* cos delta - 1 */
    double sj = tab_sc[2 * j];
    double cj = tab_sc[2 * j + 1];
    double ts = PCS_FMA(sj, cm, sj);
    double tc = PCS_FMA(cj, cm, cj);
    *s_out = PCS_FMA(cj, sd, ts);
    *c_out = PCS_FMA(-sj, sd, tc);
}

/* ------------------------------------------------------------------ one Box-Muller pair from one Philox block
 * z0 = rad*sin, z1 = rad*cos with rad = sqrt(-2 ln u); child c of a division uses z_c.
 * u_override (refcompat seeding) replaces the radius uniform when > 0.
 * The transform is given in two halves so that a kernel expanding several nodes per lane can run the polynomial
 * halves of all of them before the square roots (whose slow-path branch ends a basic block); pcs_normal_pair is
 * their composition, so the operation sequence per node is the same either way. */
PCS_HD void pcs_normal_pair_polys(pcs_u32x4 w, const double* tab, double u_override, double* rad2, double* s, double* c)
{
    double u = pcs_u53(w.x, w.y);
    if (u_override > 0.0) u = u_override;
    *rad2 = pcs_neg2log(u, tab);
    pcs_sincos2pi(((uint64_t)w.w << 32) | (uint64_t)w.z, tab + PCS_TAB_SINCOS, s, c);
}

PCS_HD void pcs_normal_pair_finish(double rad2, double s, double c, double* z0, double* z1)
{
    double rad = PCS_SQRT(rad2);
    *z0 = PCS_MUL(rad, s);
    *z1 = PCS_MUL(rad, c);
}

/* the seed cell's first-timer normal from words z (radius uniform, 32 bits) and w (angle: the top 32 bits of the 64-bit
 * angle word, the rest zero) of its SEED block; u_override (refcompat seeding) replaces the radius uniform
 * when > 0.  The seed cell is daughter 1 of the virtual division at heap 0: it takes the cosine component. */
PCS_HD double pcs_seed_normal(pcs_u32x4 w, const double* tab, double u_override)
{
    double u = pcs_u32unit(w.z);
    if (u_override > 0.0) u = u_override;
    const double rad2 = pcs_neg2log(u, tab);
    double s, c;
    pcs_sincos2pi((uint64_t)w.w << 32, tab + PCS_TAB_SINCOS, &s, &c);
    return PCS_MUL(PCS_SQRT(rad2), c);
}

PCS_HD void pcs_normal_pair(pcs_u32x4 w, const double* tab, double u_override, double* z0, double* z1)
{
    double rad2, s, c;
    pcs_normal_pair_polys(w, tab, u_override, &rad2, &s, &c);
    pcs_normal_pair_finish(rad2, s, c, z0, z1);
}

/* ------------------------------------------------------------------ one standard normal by the ziggurat method
 * (Marsaglia & Tsang 2000; 512 layers) from 64 random bits (lo, hi): sign = the top bit of hi, layer i = the 9 bits below
 * it, and the 53 bits M = (hi & 0x1FFFFF):lo the uniform M * 2^-53 in [0, 1).  x = M * 2^-53 * x_i, rounded once: the 64-bit
 * word M IS the double M * 2^-1074 (exponent fields 0 and 1 both read that way), and the table holds x_i * 2^1021, so
 * that is ONE multiplication and no integer-to-double conversion (FP64 subnormals run at full speed on the device).
 * FAST: x < x_{i+1}: the point lies under the curve whatever its height - accept (99.2 % of the draws).
 * Otherwise the trial goes on in pcs_zig_slow with further uniforms from the blocks tagged PCS_TAG_ZIGX:
 *   layer 0 (x >= r): Marsaglia's tail sampler, a = -ln(U1)/r until -2 ln(U2) > a^2, x = r + a  (always accepts);
 *   layer i >= 1: the wedge - height y = f(x_i) + U (f(x_{i+1}) - f(x_i)), accept iff y < f(x), tested as
 *   -2 ln(y) > x^2 with the same fixed-sequence logarithm as everywhere else; a rejected trial is a rejected DRAW:
 *   the caller redraws with retry + 1 exactly as for a timer <= 0.
 * tab_zig: 512 rows {x_i * 2^1021, x_{i+1}} (shared memory on the device). */
#define PCS_ZIG_LAYER(hi) (((hi) >> (31 - PCM_ZIG_N_BITS)) & ((1u << PCM_ZIG_N_BITS) - 1u))
PCS_HD bool pcs_zig_fast(uint32_t lo, uint32_t hi, const double* tab_zig, double* z_out)
{
    const uint32_t i = PCS_ZIG_LAYER(hi);
    const double xs = tab_zig[2 * i];
    const double xn = tab_zig[2 * i + 1];
    const double m = pcs_bits2d(((uint64_t)(hi & 0x001FFFFFu) << 32) | (uint64_t)lo);     /* M * 2^-1074 */
    const double x = PCS_MUL(m, xs);
    *z_out = pcs_bits2d(pcs_d2bits(x) | ((uint64_t)(hi & 0x80000000u) << 32));
    return x < xn;
}

/* the rest of a trial whose fast test failed.  *z_io = the signed x of pcs_zig_fast; c = which daughter (its half of the
 * extra blocks); tab = the whole math table; wedge = the wedge rows.  true = accepted, *z_io is the normal. */
PCS_HD bool pcs_zig_slow(uint32_t hi, uint32_t c, double* z_io, uint32_t root, uint32_t set, uint32_t retry, uint64_t heap,
                         const uint32_t* rk, const double* tab, const double* wedge)
{
    const uint32_t i = PCS_ZIG_LAYER(hi);
    const uint64_t sign = pcs_d2bits(*z_io) & 0x8000000000000000ULL;
    const double x = pcs_bits2d(pcs_d2bits(*z_io) & 0x7FFFFFFFFFFFFFFFULL);
    if (i != 0u) {
        const pcs_u32x4 e = pcs_draw_rk(root, set, retry, PCS_TAG_ZIGX, heap, rk);
        const double uu = c ? pcs_u53(e.z, e.w) : pcs_u53(e.x, e.y);
        const double y = PCS_FMA(uu, wedge[2 * i + 1], wedge[2 * i]);
        return pcs_neg2log(y, tab) > PCS_MUL(x, x);
    }
    double a = 0.0;                                 /* should every attempt fail (probability 1e-240): x = r */
    for (uint32_t k = 0; k < PCS_ZIG_TAIL_TRIES; ++k) {
        const pcs_u32x4 e = pcs_draw_rk(root, set, retry, PCS_TAG_ZIGX + k, heap, rk);
        const double u1 = pcs_u32unit(c ? e.z : e.x);
        const double u2 = pcs_u32unit(c ? e.w : e.y);
        const double t = PCS_MUL(pcs_neg2log(u1, tab), PCS_C(ZIG_INV2R));
        if (pcs_neg2log(u2, tab) > PCS_MUL(t, t)) { a = t; break; }
    }
    *z_io = pcs_bits2d(pcs_d2bits(PCS_ADD(PCS_C(ZIG_R), a)) | sign);
    return true;
}

/* one whole trial for daughter c of the division whose block (tag 0, this retry) is w */
PCS_HD bool pcs_zig_trial(pcs_u32x4 w, uint32_t c, double* z_out, uint32_t root, uint32_t set, uint32_t retry, uint64_t heap,
                          const uint32_t* rk, const double* tab, const double* wedge)
{
    const uint32_t lo = c ? w.z : w.x, hi = c ? w.w : w.y;
    if (pcs_zig_fast(lo, hi, tab + PCS_TAB_ZIG, z_out)) return true;
    return pcs_zig_slow(hi, c, z_out, root, set, retry, heap, rk, tab, wedge);
}

/* timer = mean + sd*z (one fma), accepted iff > 0 (cell.cu:106-122: redraw while rnd <= 0) */
PCS_HD double pcs_timer(double mean, double sd, double z) { return PCS_FMA(sd, z, mean); }

#endif /* PROCELL_SPEC_H */
