/* xorwow_ref.c - CPU restatement of the REFERENCE's own random stream and seed schedule.  TEST INFRASTRUCTURE ONLY.
 *
 * procell_oracle.c restates the reference's simulation semantics with the north-star Philox streams.  This file
 * restates the reference exactly as written - cuRAND XORWOW re-initialised from an integer seed for every draw,
 * the wall-clock-derived seed schedule, the level-synchronous dense ids - so that it can be checked against the
 * OUTPUT OF THE REFERENCE BINARY ITSELF run on a B200 (tests/golden/ref_*.json; see tests/golden/make_ref_fixtures.py).
 * That pins the restatement of everything except the random stream; the Philox oracle then shares those
 * semantics and is compared with this one distributionally.
 *
 * Followed, line by line (paths relative to /root/reference):
 *   src/utils/util.cu:143-169               init_random / uniform_random / normal_random
 *   CUDA 12.9 curand_kernel.h:772-798,863-874 XORWOW init (subsequence 0, offset 0) and step   [third party]
 *   CUDA 12.9 curand_uniform.h:101-106, curand_normal.h:110-132,581-596 hq uniform, Box-Muller  [third party]
 *   src/simulation/cells_population.cu:97-117  seed cell: seed = T0 + id_in_bin + f*10000
 *   src/simulation/cell.cu:25-143              create_cell, type pick, timer retry (seed *= sigma), initial age
 *   src/simulation/proliferation.cu:309-382    node rule; daughters' seeds T1 +- timer*10000 + id, ids 2*id, 2*id+1
 *   src/io/parser.cu:68-185                    result rows, type sort (descending proportion)
 * nvcc contracts a*b+c into fma by default, so the seed arithmetic and mean + sd*z are written with fma() here.
 * libdevice log/sincospi are within ~1 ulp of the long-double evaluations used here; a 1-ulp difference in a timer
 * changes a comparison or a truncated seed with probability ~1e-9 per node, i.e. never at fixture sizes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t d, v[5]; } xorwow;

static void xorwow_init(xorwow* s, uint64_t seed)           /* curand_init(seed, 0, 0, &state) */
{
    uint32_t s0 = (uint32_t)seed ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0;
    s->v[1] = 362436069u ^ t0;
    s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;
    s->v[4] = 5783321u + t0;
}

static uint32_t xorwow_next(xorwow* s)                      /* curand(&state) */
{
    uint32_t t = s->v[0] ^ (s->v[0] >> 2);
    s->v[0] = s->v[1]; s->v[1] = s->v[2]; s->v[2] = s->v[3]; s->v[3] = s->v[4];
    s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
    s->d += 362437u;
    return s->v[4] + s->d;
}

#define TWO_POW53_INV 1.1102230246251565e-16

double xorwow_uniform(uint64_t seed)                         /* utils::device::uniform_random */
{
    xorwow s;
    xorwow_init(&s, seed);
    uint32_t x = xorwow_next(&s), y = xorwow_next(&s);
    uint64_t z = (uint64_t)x ^ ((uint64_t)y << 21);
    return fma((double)z, TWO_POW53_INV, TWO_POW53_INV / 2.0);
}

double xorwow_normal(uint64_t seed, double mean, double sd)  /* utils::device::normal_random */
{
    xorwow s;
    xorwow_init(&s, seed);
    uint32_t x0 = xorwow_next(&s), x1 = xorwow_next(&s), y0 = xorwow_next(&s), y1 = xorwow_next(&s);
    uint64_t zx = (uint64_t)x0 ^ ((uint64_t)x1 << 21);
    double u = fma((double)zx, TWO_POW53_INV, TWO_POW53_INV / 2.0);
    uint64_t zy = (uint64_t)y0 ^ ((uint64_t)y1 << 21);
    double v = fma((double)zy, TWO_POW53_INV * 2.0, TWO_POW53_INV);
    double rad = sqrt(-2.0 * (double)logl((long double)u));
    double sn = (double)sinl(3.14159265358979323846264338327950288L * (long double)v);   /* sincospi(v).x */
    double z = sn * rad;
    return fma(z, sd, mean);
}

static uint64_t to_u64(double x)                             /* cvt.rzi.u64.f64: truncate, saturate */
{
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)x;
}

typedef struct { int name; double prop, mean, sd; } ref_type;
typedef struct { int type; double f, timer, t; uint64_t id; } ref_cell;

static int cmp_double(const void* a, const void* b)
{
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* cell.cu:25-79 */
static ref_cell create_cell(const ref_type* params, size_t n_types, uint64_t random_seed, int type, double f, double t)
{
    ref_cell c;
    c.type = type; c.f = f; c.timer = 0.0; c.t = 0.0; c.id = 0;
    size_t index = 0;
    if (type == -1) {                                        /* determine_cell_type, cell.cu:81-104 */
        double rnd = xorwow_uniform(random_seed);
        double acc = 0.0;
        for (size_t i = 0; i < n_types; ++i) {
            acc += params[i].prop;
            if (rnd < acc) { c.type = params[i].name; c.timer = params[i].mean; index = i; break; }
        }
    } else {
        for (size_t i = 0; i < n_types; ++i)
            if (type == params[i].name) { index = i; break; }
    }
    const ref_type* p = &params[index];
    if (p->mean >= 0.0) {                                    /* determine_cell_timer, cell.cu:106-122 */
        if (p->mean > 0.0) {
            double rnd = -1.0;
            uint64_t s = random_seed;
            while (rnd <= 0.0) {
                rnd = xorwow_normal(s, p->mean, p->sd);
                s = to_u64((double)s * p->sd);
            }
            c.timer = rnd;
        }
    } else {
        c.timer = -1.0;
    }
    if (t > 0) {
        c.t = t;
    } else if (p->mean >= 0.0) {                             /* determine_cell_initial_t, cell.cu:124-143 */
        if (p->mean > 0.0) {
            double rnd = -1.0;
            uint64_t s = random_seed;
            while (rnd <= 0.0) {
                rnd = xorwow_normal(s, p->mean, p->sd);
                s = to_u64((double)s * p->sd);
            }
            double factor = xorwow_uniform(to_u64((double)s * p->sd));
            c.t = rnd * factor;
        }
    } else {
        c.t = 0;
    }
    return c;
}

/* Output arrays are caller-allocated with capacity max_rows; returns the number of rows (ascending value, zero rows
 * included) or a negative error.  T0 = time(NULL) seen by create_cells_population (cells_population.cu:34), T1 = the one
 * seen by run_iteration (proliferation.cu:242).  max_depth = levels of one iteration; -2 is returned if a cell is
 * still alive there (the reference would start another iteration with a new wall-clock seed). */
long xorwow_ref_simulate(const double* value, const uint64_t* freq, size_t n_lines, const double* types, size_t n_types,
                         double t_max, double phi, uint64_t T0, uint64_t T1, int max_depth, size_t max_rows,
                         double* row_value, uint64_t* row_freq, int32_t* row_ratio, uint64_t* divisions_out)
{
    /* parser.cu:156-185: types in file order get name = index, then sorted by proportion, descending */
    ref_type* params = (ref_type*)malloc(n_types * sizeof(ref_type));
    for (size_t j = 0; j < n_types; ++j) {
        ref_type t = { (int)j, types[3 * j], types[3 * j + 1], types[3 * j + 2] };
        size_t pos = j;
        while (pos > 0 && params[pos - 1].prop < t.prop) { params[pos] = params[pos - 1]; --pos; }
        params[pos] = t;
    }
    /* parser.cu:68-154: result rows */
    size_t cap = 0, total = 0;
    for (size_t i = 0; i < n_lines; ++i)
        if (freq[i] > 0) {
            total += freq[i];
            for (double c = value[i]; c >= phi && cap < (1u << 26); c = c / 2) ++cap;
        }
    double* rows = (double*)malloc((cap + 1) * sizeof(double));
    size_t nr = 0;
    for (size_t i = 0; i < n_lines; ++i)
        if (freq[i] > 0)
            for (double c = value[i]; c >= phi && nr < cap; c = c / 2) rows[nr++] = c;
    qsort(rows, nr, sizeof(double), cmp_double);
    size_t nu = 0;
    for (size_t i = 0; i < nr; ++i)
        if (nu == 0 || rows[i] != rows[nu - 1]) rows[nu++] = rows[i];
    if (nu > max_rows) { free(rows); free(params); return -1; }
    memcpy(row_value, rows, nu * sizeof(double));
    memset(row_freq, 0, nu * sizeof(uint64_t));
    memset(row_ratio, 0, nu * n_types * sizeof(int32_t));
    free(rows);

    /* cells_population.cu:97-117: one seed cell per counted event, id = index within its bin */
    ref_cell* cur = (ref_cell*)malloc((total + 1) * sizeof(ref_cell));
    size_t n_cur = 0;
    uint64_t start = 0;
    for (size_t i = 0; i < n_lines; ++i) {
        if (freq[i] == 0) continue;
        for (uint64_t id = 0; id < freq[i]; ++id) {
            uint64_t seed = to_u64(fma(value[i], 10000.0, (double)(T0 + id)));
            ref_cell c = create_cell(params, n_types, seed, -1, value[i], 0.0);
            c.id = start + id;
            cur[n_cur++] = c;
        }
        start += freq[i];
    }
    /* proliferation.cu:309-382, level by level */
    uint64_t divisions = 0;
    long rc = (long)nu;
    for (int level = 0; n_cur > 0; ++level) {
        if (level >= max_depth) { rc = -2; break; }
        ref_cell* next = (ref_cell*)malloc((2 * n_cur + 1) * sizeof(ref_cell));
        size_t n_next = 0;
        for (size_t k = 0; k < n_cur; ++k) {
            ref_cell* c = &cur[k];
            int out_of_time = (c->timer < 0.0) || (c->t + c->timer > t_max);
            if (!out_of_time) {
                if (c->f / 2 > phi) {
                    double f2 = c->f / 2;
                    double t2 = c->t + c->timer;
                    uint64_t seed_c1 = to_u64(fma(c->timer, 10000.0, (double)T1) + (double)c->id);
                    uint64_t seed_c2 = to_u64(fma(-c->timer, 10000.0, (double)T1) + (double)c->id);
                    ref_cell a = create_cell(params, n_types, seed_c1, c->type, f2, t2);
                    ref_cell b = create_cell(params, n_types, seed_c2, c->type, f2, t2);
                    a.id = 2 * c->id; b.id = 2 * c->id + 1;
                    next[n_next++] = a; next[n_next++] = b;
                    ++divisions;
                }
            } else {
                size_t lo = 0, hi = nu;
                while (lo < hi) { size_t mid = (lo + hi) / 2; if (row_value[mid] < c->f) lo = mid + 1; else hi = mid; }
                if (lo < nu && row_value[lo] == c->f) {
                    row_freq[lo] += 1;
                    if (c->type >= 0) row_ratio[lo * n_types + (size_t)c->type] += 1;
                }
            }
        }
        free(cur);
        cur = next;
        n_cur = n_next;
    }
    free(cur);
    free(params);
    if (divisions_out) *divisions_out = divisions;
    return rc;
}
