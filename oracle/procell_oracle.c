/* procell_oracle.c - CPU ORACLE for the B200 proliferation simulator.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call this file.  The product (cuda_pro_cell_b200/) never links or calls it and has no CPU path.
 *
 * What it restates (ericniso/cuda-pro-cell, paths relative to /root/reference):
 *   - result-key set and default phi ......... src/io/parser.cu:68-154   (oracle_plan_*)
 *   - type table, sort by proportion ......... src/io/parser.cu:156-185  (sort_types)
 *   - seed cell construction ................. src/simulation/cells_population.cu:104-115,
 *                                              src/simulation/cell.cu:25-143 (seed_cell)
 *   - node rule (divide / count / drop) ...... src/simulation/proliferation.cu:309-382,404-410 (expand_root)
 *   - output rows ............................ src/io/parser.cu:187-217  (merged rows, ascending, zero rows kept
 *                                              here and skipped by the writer)
 * What is NOT the reference's: the random stream.  The reference re-seeds a cuRAND XORWOW state from
 * wall-clock-derived integers for every draw (src/utils/util.cu:143-169), which is neither reproducible
 * nor shardable.  The north-star spec replaces it with Philox4x32-10 (Salmon et al., SC'11; the published
 * algorithm, restated below) keyed by (root cell, tree path), one block per division, whose two normals are drawn by
 * the ziggurat method (division timers) or a Box-Muller transform (the seed cell's first timer), both built from a fixed
 * sequence of correctly-rounded IEEE-754 operations so that GPU and CPU agree bit for bit.  oracle/xorwow_ref.c restates the reference's own stream for the distributional check.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures for this path, so the
 * pin is the reference itself: tests/golden/ref_*.json hold outputs of the UNMODIFIED reference binary run on a
 * B200 (tests/golden/make_ref_fixtures.py).  oracle/xorwow_ref.c - the restatement of the reference as written,
 * sharing this file's semantics but the reference's XORWOW stream - reproduces every fixture bit for bit, and
 * this Philox oracle agrees with it in law (chi-square / KS, tests/test_reference_fixtures.py).  Further pins:
 * Random123's published Philox4x32-10 known-answer vectors and the deterministic invariants derivable from the
 * reference source (t_max=0 identity, all-quiescent identity, fluorescence-mass conservation, ratio-row sums,
 * sigma=0 closed form).  See DESIGN.md "Oracle and pinning".
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -shared -fPIC (oracle/Makefile).  -ffp-contract=off matters:
 * no a*b+c may be fused except where fma() is written.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_math_tables.inc"

#define ORC_MAX_LEVEL 63
#define ORC_MAX_RETRY 255

typedef union { uint64_t u; double d; } orc_pun;
static double as_f64(uint64_t u) { orc_pun p; p.u = u; return p.d; }
static uint64_t as_u64(double d) { orc_pun p; p.d = d; return p.u; }

static const uint64_t log_rows[1 << PCM_LOG_N_BITS][2] = { PCM_LOG_TABLE_ROWS };
static const uint64_t sincos_rows[1 << PCM_SC_N_BITS][2] = { PCM_SINCOS_TABLE_ROWS };

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10.  ctr[4], key[2]; ten rounds of (mulhi, mullo, xor), key bumped by the Weyl constants.
 * ---------------------------------------------------------------------------------------------- */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t a = ctr[0], b = ctr[1], c = ctr[2], d = ctr[3];
    uint32_t ka = key[0], kb = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t prod_a = 0xD2511F53ull * a;
        uint64_t prod_c = 0xCD9E8D57ull * c;
        uint32_t na = (uint32_t)(prod_c >> 32) ^ b ^ ka;
        uint32_t nb = (uint32_t)prod_c;
        uint32_t nc = (uint32_t)(prod_a >> 32) ^ d ^ kb;
        uint32_t nd = (uint32_t)prod_a;
        a = na; b = nb; c = nc; d = nd;
        ka += 0x9E3779B9u;
        kb += 0xBB67AE85u;
    }
    out[0] = a; out[1] = b; out[2] = c; out[3] = d;
}

/* stream layout: ctr = { root, set | retry<<16 | tag<<24, heap_lo, heap_hi }, key = { seed_lo, seed_hi } */
static void draw_block(uint32_t root, uint32_t set, uint32_t retry, uint32_t tag, uint64_t heap,
                       uint64_t seed, uint32_t out[4])
{
    uint32_t ctr[4] = { root, set | (retry << 16) | (tag << 24), (uint32_t)heap, (uint32_t)(heap >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    oracle_philox4x32_10(ctr, key, out);
}

/* (2m+1)/2^53 for a 52-bit m: build 1.m and subtract (1 - 2^-53); exact */
static double unit_from_mantissa(uint64_t m52)
{
    return as_f64(0x3FF0000000000000ull | m52) - as_f64(PCM_BITS_ONE_M);
}

double oracle_uniform53(uint32_t lo, uint32_t hi)
{
    uint64_t w = ((uint64_t)hi << 32) | lo;
    return unit_from_mantissa(w >> 12);
}

/* (2m+1)/2^33 for 32 random bits m: the seed-cell draws */
double oracle_uniform32(uint32_t m)
{
    return fma((double)m, ldexp(1.0, -32), ldexp(1.0, -33));
}

/* -2 ln(u) for a normal double u in (0,1) */
double oracle_neg2log(double u)
{
    uint64_t bits = as_u64(u);
    uint64_t rel = bits - PCM_LOG_OFF;
    unsigned idx = (unsigned)((rel >> (52 - PCM_LOG_N_BITS)) & ((1u << PCM_LOG_N_BITS) - 1));
    int expo = (int)((int64_t)rel >> 52);
    double z = as_f64(bits - (rel & 0xFFF0000000000000ull));
    double invc = as_f64(log_rows[idx][0]);
    double logc = as_f64(log_rows[idx][1]);
    double r = fma(z, invc, -1.0);
    double q = as_f64(PCM_BITS_LOG1P_B6);
    q = fma(q, r, as_f64(PCM_BITS_LOG1P_B5));
    q = fma(q, r, as_f64(PCM_BITS_LOG1P_B4));
    q = fma(q, r, as_f64(PCM_BITS_LOG1P_B3));
    q = fma(q, r, as_f64(PCM_BITS_LOG1P_B2));
    double rr = r * r;
    double tail = fma(rr, q, r);
    double head = fma((double)expo, as_f64(PCM_BITS_LN2), logc);
    double lg = head + tail;
    return lg * -2.0;
}

/* sin and cos of 2*pi*v/2^64: the circle is cut into 256 sectors; the top 8 bits pick the sector, the next 52 the
 * offset from the sector's centre angle c, |offset| < pi/256; angle-addition with short series for the offset */
void oracle_sincos2pi(uint64_t v, double* sin_out, double* cos_out)
{
    unsigned sector = (unsigned)(v >> 56);
    uint64_t frac = (v >> 4) & 0xFFFFFFFFFFFFFull;
    double one_to_two = as_f64(0x3FF0000000000000ull | frac);
    double off = fma(one_to_two, as_f64(PCM_BITS_SC_A), as_f64(PCM_BITS_SC_B));
    double off2 = off * off;
    /* sin(off) = off + off^3 * (S1 + S2 off^2) */
    double sin_poly = fma(as_f64(PCM_BITS_SD_S2), off2, as_f64(PCM_BITS_SD_S1));
    double off3 = off * off2;
    double sin_off = fma(off3, sin_poly, off);
    /* cos(off) - 1 = off^2 * (C1 + off^2 * (C2 + C3 off^2)) */
    double cos_poly = fma(as_f64(PCM_BITS_CM_C3), off2, as_f64(PCM_BITS_CM_C2));
    cos_poly = fma(cos_poly, off2, as_f64(PCM_BITS_CM_C1));
    double cos_off_m1 = cos_poly * off2;
    double sin_c = as_f64(sincos_rows[sector][0]);
    double cos_c = as_f64(sincos_rows[sector][1]);
    double sin_part = fma(sin_c, cos_off_m1, sin_c);      /* sin c * cos off */
    double cos_part = fma(cos_c, cos_off_m1, cos_c);      /* cos c * cos off */
    *sin_out = fma(cos_c, sin_off, sin_part);
    *cos_out = fma(-sin_c, sin_off, cos_part);
}

/* one Philox block -> two independent standard normals; u_forced > 0 replaces the radius uniform */
void oracle_normal_pair(const uint32_t w[4], double u_forced, double z[2])
{
    double u = oracle_uniform53(w[0], w[1]);
    if (u_forced > 0.0) u = u_forced;
    double rad = sqrt(oracle_neg2log(u));
    double s, c;
    oracle_sincos2pi(((uint64_t)w[3] << 32) | w[2], &s, &c);
    z[0] = rad * s;
    z[1] = rad * c;
}

/* ------------------------------------------------------------------------------------------------
 * Plan: the (bin, k) key space and the merged output rows (parser.cu:68-154).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    size_t n_bins;          /* lines with frequency > 0, file order (parser.cu:105-123) */
    double* bin_value;
    uint64_t* bin_freq;
    uint64_t* bin_start;    /* running start index of the bin's first seed cell ("bounds") */
    uint8_t* bin_kdiv;      /* halvings allowed: #{k>=0 : value/2^(k+1) > phi}  (proliferation.cu:323) */
    uint8_t* bin_count0;    /* level-0 leaves are countable iff value >= phi (parser.cu:127) */
    uint32_t* bin_keybase;  /* first key of the bin; key = keybase + k */
    size_t n_keys;
    uint32_t* key_row;      /* key -> merged output row */
    size_t n_rows;
    double* row_value;      /* ascending, the std::map order of parser.cu:142-151 */
    uint64_t n_cells;
    double phi;
    int depth_capped;
} oracle_plan;

static int cmp_double(const void* a, const void* b)
{
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

void oracle_plan_free(oracle_plan* p)
{
    if (!p) return;
    free(p->bin_value); free(p->bin_freq); free(p->bin_start); free(p->bin_kdiv); free(p->bin_count0);
    free(p->bin_keybase); free(p->key_row); free(p->row_value); free(p);
}

oracle_plan* oracle_plan_create(const double* value, const uint64_t* freq, size_t n_lines, double phi)
{
    oracle_plan* p = (oracle_plan*)calloc(1, sizeof *p);
    /* default phi: smallest value with frequency > 0 (parser.cu:80-96) */
    if (phi == 0.0) {
        for (size_t i = 0; i < n_lines; ++i)
            if (freq[i] > 0 && (phi == 0.0 || value[i] < phi)) phi = value[i];
    }
    p->phi = phi;
    size_t nb = 0;
    for (size_t i = 0; i < n_lines; ++i) nb += freq[i] > 0;
    p->n_bins = nb;
    p->bin_value = (double*)malloc((nb + 1) * sizeof(double));
    p->bin_freq = (uint64_t*)malloc((nb + 1) * sizeof(uint64_t));
    p->bin_start = (uint64_t*)malloc((nb + 1) * sizeof(uint64_t));
    p->bin_kdiv = (uint8_t*)malloc(nb + 1);
    p->bin_count0 = (uint8_t*)malloc(nb + 1);
    p->bin_keybase = (uint32_t*)malloc((nb + 1) * sizeof(uint32_t));
    size_t b = 0, nkeys = 0;
    uint64_t total = 0;
    for (size_t i = 0; i < n_lines; ++i) {
        if (freq[i] == 0) continue;
        p->bin_value[b] = value[i];
        p->bin_freq[b] = freq[i];
        p->bin_start[b] = total;
        total += freq[i];
        unsigned k = 0;
        double f = value[i];
        while (k < ORC_MAX_LEVEL && f / 2 > phi) { f = f / 2; ++k; }
        if (k == ORC_MAX_LEVEL && f / 2 > phi) p->depth_capped = 1;
        p->bin_kdiv[b] = (uint8_t)k;
        p->bin_count0[b] = (uint8_t)(value[i] >= phi);
        p->bin_keybase[b] = (uint32_t)nkeys;
        nkeys += k + 1;
        ++b;
    }
    p->bin_start[nb] = total;
    p->n_cells = total;
    p->n_keys = nkeys;
    /* key values by repeated halving exactly as parser.cu:126-137 does, then sort + unique */
    double* kv = (double*)malloc((nkeys + 1) * sizeof(double));
    for (b = 0; b < nb; ++b) {
        double f = p->bin_value[b];
        for (unsigned k = 0; k <= p->bin_kdiv[b]; ++k) { kv[p->bin_keybase[b] + k] = f; f = f / 2; }
    }
    double* sorted = (double*)malloc((nkeys + 1) * sizeof(double));
    size_t ns = 0;
    for (b = 0; b < nb; ++b)
        for (unsigned k = 0; k <= p->bin_kdiv[b]; ++k)
            if (k > 0 || p->bin_count0[b]) sorted[ns++] = kv[p->bin_keybase[b] + k];
    qsort(sorted, ns, sizeof(double), cmp_double);
    size_t nr = 0;
    for (size_t i = 0; i < ns; ++i)
        if (nr == 0 || sorted[i] != sorted[nr - 1]) sorted[nr++] = sorted[i];
    p->n_rows = nr;
    p->row_value = sorted;
    p->key_row = (uint32_t*)malloc((nkeys + 1) * sizeof(uint32_t));
    for (size_t key = 0; key < nkeys; ++key) {
        size_t lo = 0, hi = nr;     /* first row with value >= kv[key] */
        while (lo < hi) { size_t mid = (lo + hi) / 2; if (sorted[mid] < kv[key]) lo = mid + 1; else hi = mid; }
        p->key_row[key] = (lo < nr && sorted[lo] == kv[key]) ? (uint32_t)lo : 0xFFFFFFFFu;
    }
    free(kv);
    return p;
}

size_t oracle_plan_n_bins(const oracle_plan* p) { return p->n_bins; }
size_t oracle_plan_n_keys(const oracle_plan* p) { return p->n_keys; }
size_t oracle_plan_n_rows(const oracle_plan* p) { return p->n_rows; }
uint64_t oracle_plan_n_cells(const oracle_plan* p) { return p->n_cells; }
double oracle_plan_phi(const oracle_plan* p) { return p->phi; }
int oracle_plan_depth_capped(const oracle_plan* p) { return p->depth_capped; }
void oracle_plan_export(const oracle_plan* p, double* row_value, uint32_t* key_row, uint32_t* bin_keybase,
                        uint8_t* bin_kdiv)
{
    if (row_value) memcpy(row_value, p->row_value, p->n_rows * sizeof(double));
    if (key_row) memcpy(key_row, p->key_row, p->n_keys * sizeof(uint32_t));
    if (bin_keybase) memcpy(bin_keybase, p->bin_keybase, p->n_bins * sizeof(uint32_t));
    if (bin_kdiv) memcpy(bin_kdiv, p->bin_kdiv, p->n_bins);
}

/* ------------------------------------------------------------------------------------------------
 * Types (parser.cu:156-185): id = file line, selection order = descending proportion (stable).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double prop, mean, sd; int id; } orc_type;

static void sort_types(const double* tri, size_t n, orc_type* out)
{
    for (size_t j = 0; j < n; ++j) {
        orc_type t = { tri[3 * j], tri[3 * j + 1], tri[3 * j + 2], (int)j };
        size_t pos = j;
        while (pos > 0 && out[pos - 1].prop < t.prop) { out[pos] = out[pos - 1]; --pos; }
        out[pos] = t;
    }
}

/* proportion check of parser.cu:46-66: |1 - sum| > 1e-8 is an error.  Returns 0 when acceptable. */
int oracle_check_proportions(const double* tri, size_t n)
{
    double sum = 0.0;
    for (size_t j = 0; j < n; ++j) sum = sum + tri[3 * j];
    return fabs(1.0 - sum) > 1.0 / pow(10.0, 8.0);
}

/* ------------------------------------------------------------------------------------------------
 * The simulation.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t heap; double t_div; } orc_node;

typedef struct {
    const oracle_plan* plan;
    const orc_type* types;  /* sorted, this set */
    size_t n_types;
    uint32_t set;
    double t_max;
    uint64_t seed;
    int refcompat;
    int64_t* counts;        /* [n_keys][n_types] for this set, type index = file order */
    int64_t divisions;
    /* subtree sharding (shard_level L >= 1, world > 1): see oracle_simulate */
    uint32_t sub_level, sub_world, sub_rank;
} orc_ctx;

/* ------------------------------------------------------------------------------------------------
 * Standard normal for a division timer: the ziggurat method (Marsaglia & Tsang, "The Ziggurat Method for Generating
 * Random Variables", J. Stat. Softw. 5(8), 2000 - the published algorithm, restated).  512 layers of equal area under
 * f(x) = exp(-x^2/2); zig_rows[i] = { x_i, x_(i+1) } with x_0 = V/f(r), x_1 = r, x_512 = 0.
 *   64 random bits (lo, hi): sign = bit 31 of hi, layer = the 9 bits below it, and the 53-bit integer M = (hi & 0x1FFFFF):lo
 *   gives the uniform M / 2^53.  x = (M / 2^53) * x_layer (one rounding).  x < x_(layer+1): under the curve for sure, accept.
 *   (The table row holds x_layer * 2^1021 - the kernels multiply it with M read as a subnormal double - and is unscaled here.)
 *   layer 0 otherwise: the tail beyond r, by a = -ln(U1)/r until -2 ln(U2) > a^2, x = r + a.
 *   layer >= 1 otherwise: the wedge; y uniform between f(x_layer) and f(x_(layer+1)), accept iff y < f(x), i.e.
 *   -2 ln y > x^2.  A rejected trial counts as a rejected draw (the division redraws with the next retry number).
 * The extra uniforms come from further Philox blocks of the same (root, heap, retry) with tag 2 (+ attempt number in
 * the tail); daughter c reads half c of a block, as for the first one.
 * ---------------------------------------------------------------------------------------------- */
static const uint64_t zig_rows[1 << PCM_ZIG_N_BITS][2] = { PCM_ZIG_TABLE_ROWS };
static const uint64_t zig_wedge[1 << PCM_ZIG_N_BITS][2] = { PCM_ZIG_WEDGE_ROWS };
#define ORC_ZIG_TAIL_TRIES 200

/* 1 = accepted (normal in *z), 0 = rejected.  words[4] = the division's block of this retry */
int oracle_zig_trial(const uint32_t words[4], unsigned c, uint32_t root, uint32_t set, uint32_t retry, uint64_t heap,
                     uint64_t seed, double* z)
{
    const uint32_t lo = words[2 * c], hi = words[2 * c + 1];
    const unsigned layer = (hi >> (31 - PCM_ZIG_N_BITS)) & ((1u << PCM_ZIG_N_BITS) - 1u);
    const int negative = (int)(hi >> 31);
    const uint64_t mant53 = ((uint64_t)(hi & 0x1FFFFFu) << 32) | lo;
    const double uniform = (double)mant53 * 0x1p-53;               /* exact: 53 bits, a power of two */
    const double edge = as_f64(zig_rows[layer][0]) * 0x1p-1021;     /* exact: undoes the table's scaling */
    double x = uniform * edge;
    if (!(x < as_f64(zig_rows[layer][1]))) {
        uint32_t e[4];
        if (layer == 0) {
            double beyond = 0.0;
            for (uint32_t attempt = 0; attempt < ORC_ZIG_TAIL_TRIES; ++attempt) {
                draw_block(root, set, retry, 2u + attempt, heap, seed, e);
                double a = oracle_neg2log(oracle_uniform32(e[2 * c])) * as_f64(PCM_BITS_ZIG_INV2R);
                if (oracle_neg2log(oracle_uniform32(e[2 * c + 1])) > a * a) { beyond = a; break; }
            }
            x = as_f64(PCM_BITS_ZIG_R) + beyond;
        } else {
            draw_block(root, set, retry, 2u, heap, seed, e);
            double height = fma(oracle_uniform53(e[2 * c], e[2 * c + 1]), as_f64(zig_wedge[layer][1]), as_f64(zig_wedge[layer][0]));
            if (!(oracle_neg2log(height) > x * x)) return 0;
        }
    }
    *z = negative ? -x : x;
    return 1;
}

/* n standard normals exactly as a division's daughter 0 / 1 would receive them (root = sample index / 2, heap 1, redraw on
 * rejection): for the law tests of the sampler.  trials (optional) receives the total number of trials made. */
void oracle_zig_fill(uint64_t seed, uint64_t n, double* out, uint64_t* trials)
{
    uint64_t made = 0;
    for (uint64_t k = 0; k < n; ++k) {
        out[k] = 0.0;
        for (uint32_t retry = 0; retry < ORC_MAX_RETRY; ++retry) {
            uint32_t w[4];
            draw_block((uint32_t)(k >> 1), 0u, retry, 0u, 1ull, seed, w);
            ++made;
            if (oracle_zig_trial(w, (unsigned)(k & 1), (uint32_t)(k >> 1), 0u, retry, 1ull, seed, &out[k])) break;
        }
    }
    if (trials) *trials = made;
}

/* truncated-normal timers for the children in `want` (bit c = child c) of the division at `heap`: child c is tried at
 * retry 0, 1, 2, ... until a trial is accepted AND the timer is positive (cell.cu:106-122: redraw while <= 0) */
static void division_timers(const orc_ctx* cx, uint32_t root, uint64_t heap, const orc_type* ty,
                            unsigned want, double timer[2])
{
    for (uint32_t retry = 0; want; ++retry) {
        if (retry == ORC_MAX_RETRY) {            /* 255 rejections in a row: fall back to the mean */
            if (want & 1) timer[0] = ty->mean;
            if (want & 2) timer[1] = ty->mean;
            break;
        }
        uint32_t w[4];
        draw_block(root, cx->set, retry, 0u, heap, cx->seed, w);
        for (unsigned c = 0; c < 2; ++c) {
            double z;
            if (!(want & (1u << c))) continue;
            if (!oracle_zig_trial(w, c, root, cx->set, retry, heap, cx->seed, &z)) continue;
            double cand = fma(ty->sd, z, ty->mean);
            if (cand > 0.0) { timer[c] = cand; want &= ~(1u << c); }
        }
    }
}

static void count_leaf(orc_ctx* cx, unsigned bin, unsigned level, int type_id)
{
    const oracle_plan* p = cx->plan;
    if (level == 0 && !p->bin_count0[bin]) return;       /* value < phi: the bsearch of :353-380 misses */
    cx->counts[(size_t)(p->bin_keybase[bin] + level) * cx->n_types + (size_t)type_id] += 1;
}

static void expand_root(orc_ctx* cx, uint32_t root, unsigned bin)
{
    const oracle_plan* p = cx->plan;
    /* seed cell: cells_population.cu:108-111 -> cell.cu:25-79 with type == -1, t == 0.
     * ONE Philox block per seed cell (tag 1, heap 0): word 0 -> type uniform, word 1 -> initial-age uniform (32 bits each),
     * words 2, 3 -> the first timer's normal: 64 bits for a ziggurat trial (ideal seeding), or radius uniform and angle of a
     * Box-Muller draw at 32 bits each (refcompat seeding) */
    uint32_t w[4];
    draw_block(root, cx->set, 0u, 1u, 0ull, cx->seed, w);
    double u_type = oracle_uniform32(w[0]);
    double u_age = oracle_uniform32(w[1]);
    const orc_type* ty = &cx->types[cx->n_types - 1];     /* Q17: nothing matched -> last in selection order */
    double acc = 0.0;
    for (size_t j = 0; j < cx->n_types; ++j) {            /* cell.cu:81-104 */
        acc += cx->types[j].prop;
        if (u_type < acc) { ty = &cx->types[j]; break; }
    }
    /* with subtree sharding every rank walks the first sub_level levels of EVERY lineage; what those levels count
     * is credited to one rank only, root % world */
    const int sub = cx->sub_level > 0;
    const int low_owner = !sub || root % cx->sub_world == cx->sub_rank;
    if (ty->mean < 0.0) {                                 /* quiescent: timer = -1, t = 0 -> out_of_time */
        if (low_owner) count_leaf(cx, bin, 0, ty->id);
        return;
    }
    double timer[2];
    /* first timer: truncated normal by redraw (cell.cu:106-122).  Round 0 takes words 2, 3 of the seed block itself;
     * a rejected draw (trial rejected, or timer <= 0) takes words 2, 3 of the block with the next retry number; after 255
     * rejections the mean. */
    timer[1] = ty->mean;
    for (uint32_t retry = 0; retry < ORC_MAX_RETRY; ++retry) {
        if (retry > 0) draw_block(root, cx->set, retry, 1u, 0ull, cx->seed, w);
        double z;
        if (cx->refcompat) {
            /* the reference's coupling is defined on a Box-Muller draw: round 0's radius uniform IS the type uniform */
            double u_rad = retry == 0 ? u_type : oracle_uniform32(w[2]);
            double sn, cs;
            oracle_sincos2pi((uint64_t)w[3] << 32, &sn, &cs);
            z = sqrt(oracle_neg2log(u_rad)) * cs;
        } else {
            /* ideal seeding: one ziggurat trial on words 2, 3, exactly as daughter 1 of a division at tree path 0 would make it */
            if (!oracle_zig_trial(w, 1u, root, cx->set, retry, 0ull, cx->seed, &z)) continue;
        }
        double cand = fma(ty->sd, z, ty->mean);
        if (cand > 0.0) { timer[1] = cand; break; }
    }
    double t0 = timer[1] * u_age;                         /* cell.cu:124-143: t = timer * U' */
    double t_div = t0 + timer[1];
    if (t_div > cx->t_max) { if (low_owner) count_leaf(cx, bin, 0, ty->id); return; }     /* proliferation.cu:404-410 */
    if (p->bin_kdiv[bin] == 0) return;                    /* f/2 <= phi: silently dropped (Q6) */

    orc_node stack[2 * ORC_MAX_LEVEL + 4];
    int sp = 0;
    stack[sp].heap = 1; stack[sp].t_div = t_div; ++sp;
    while (sp > 0) {
        orc_node nd = stack[--sp];
        unsigned level = 63u - (unsigned)__builtin_clzll(nd.heap);
        /* a node below the shard level is expanded by every rank and credited to one; from the shard level on a node
         * exists on one rank only */
        const int credit = !sub || level >= cx->sub_level || low_owner;
        if (credit) cx->divisions += 1;
        division_timers(cx, root, nd.heap, ty, 3u, timer);
        for (unsigned c = 0; c < 2; ++c) {
            double t_child = nd.t_div + timer[c];         /* child.t = parent.t + parent.timer; + own timer */
            if (t_child > cx->t_max) { if (credit) count_leaf(cx, bin, level + 1, ty->id); }
            else if (level + 1 < p->bin_kdiv[bin]) {
                const uint64_t child = 2 * nd.heap + c;
                /* the subtree under a node AT the shard level belongs to rank (root + heap) mod world */
                if (sub && level + 1 == cx->sub_level && (root + (uint32_t)child) % cx->sub_world != cx->sub_rank) continue;
                stack[sp].heap = child; stack[sp].t_div = t_child; ++sp;
            }
            /* else: alive, in time, f/2 <= phi -> vanishes uncounted */
        }
    }
}

/* counts: [n_sets][n_keys][n_types] int64 (zeroed here); divisions: [n_sets].
 * types: [n_sets][n_types][3] = proportion, mean, sd in FILE order.
 * shard_level == 0: roots in [root_begin, root_end) with (root / shard_unit) % shard_world == shard_rank are simulated
 *   (whole lineages per rank).
 * shard_level == L >= 1 (subtree sharding, SURVEY 8e: "for config 4 subtree granularity is required for balance"):
 *   every rank builds every seed cell and expands every node of tree level < L (root = level 0); level-0 leaves and
 *   the divisions and leaves of nodes below level L are credited to rank root % world; a daughter AT level L that will
 *   divide is kept by rank (root + heap) % world alone, and everything under it is that rank's.  The ranks' tensors
 *   sum to the unsharded result because the random stream is keyed by (root, tree path), not by who expands a node. */
int oracle_simulate(const oracle_plan* p, const double* types, size_t n_types, size_t n_sets, double t_max,
                    uint64_t seed, int refcompat, uint64_t root_begin, uint64_t root_end,
                    uint32_t shard_unit, uint32_t shard_world, uint32_t shard_rank, uint32_t shard_level, int n_threads,
                    int64_t* counts, int64_t* divisions)
{
    if (shard_level > 62) return -3;
    if (n_types == 0 || n_types > 64 || n_sets == 0 || n_sets > 65536) return -1;
    if (p->n_cells > 0xFFFFFFFFull) return -2;
    if (root_end > p->n_cells) root_end = p->n_cells;
    if (shard_world == 0) { shard_world = 1; shard_rank = 0; }
    if (shard_unit == 0) shard_unit = 1;
    size_t per_set = p->n_keys * n_types;
    memset(counts, 0, n_sets * per_set * sizeof(int64_t));
    memset(divisions, 0, n_sets * sizeof(int64_t));
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    for (size_t s = 0; s < n_sets; ++s) {
        orc_type sorted[64];
        sort_types(types + s * n_types * 3, n_types, sorted);
        int64_t div_total = 0;
#pragma omp parallel num_threads(n_threads) reduction(+ : div_total)
        {
            orc_ctx cx;
            cx.plan = p; cx.types = sorted; cx.n_types = n_types; cx.set = (uint32_t)s; cx.t_max = t_max;
            cx.seed = seed; cx.refcompat = refcompat; cx.divisions = 0;
            cx.sub_level = shard_world > 1 ? shard_level : 0; cx.sub_world = shard_world; cx.sub_rank = shard_rank;
            cx.counts = (int64_t*)calloc(per_set ? per_set : 1, sizeof(int64_t));
#pragma omp for schedule(dynamic, 1)
            for (size_t b = 0; b < p->n_bins; ++b) {
                uint64_t lo = p->bin_start[b], hi = p->bin_start[b + 1];
                if (lo < root_begin) lo = root_begin;
                if (hi > root_end) hi = root_end;
                for (uint64_t r = lo; r < hi; ++r) {
                    if (!cx.sub_level && (r / shard_unit) % shard_world != shard_rank) continue;
                    expand_root(&cx, (uint32_t)r, (unsigned)b);
                }
            }
            div_total += cx.divisions;
#pragma omp critical
            for (size_t i = 0; i < per_set; ++i) counts[s * per_set + i] += cx.counts[i];
            free(cx.counts);
        }
        divisions[s] = div_total;
    }
    return 0;
}

/* merged rows (parser.cu:142-151,187-217): row_freq[n_rows], row_ratio[n_rows][n_types] for one set */
void oracle_merge_rows(const oracle_plan* p, const int64_t* counts_one_set, size_t n_types,
                       int64_t* row_freq, int64_t* row_ratio)
{
    memset(row_freq, 0, p->n_rows * sizeof(int64_t));
    memset(row_ratio, 0, p->n_rows * n_types * sizeof(int64_t));
    for (size_t key = 0; key < p->n_keys; ++key) {
        uint32_t row = p->key_row[key];
        if (row == 0xFFFFFFFFu) continue;
        for (size_t j = 0; j < n_types; ++j) {
            int64_t c = counts_one_set[key * n_types + j];
            row_freq[row] += c;
            row_ratio[(size_t)row * n_types + j] += c;
        }
    }
}
