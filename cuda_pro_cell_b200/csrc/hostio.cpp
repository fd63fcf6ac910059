/* hostio.cpp - the result-key plan and the proportion check (host only, no CUDA).
 *
 * Replaces, in the reference's src/io/parser.cu, the result-key precomputation of load_fluorescences (:68-154)
 * and assert_proportion_sum (:46-66).  The text readers and the writer live in textio.cpp.
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "host_plan.h"

namespace procell_b200 {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

}  // namespace procell_b200

namespace procell_b200 {
namespace {

/* rows and the key -> row map from (value, key) pairs in ascending value order: equal values merge (parser.cu:142-151) */
void rows_from_sorted(procell_plan* p, const std::vector<std::pair<double, uint32_t>>& kv)
{
    p->key_row.assign(p->n_keys, 0xFFFFFFFFu);
    p->row_value.clear();
    p->row_value.reserve(kv.size());
    for (size_t i = 0; i < kv.size(); ++i) {
        if (p->row_value.empty() || kv[i].first != p->row_value.back()) p->row_value.push_back(kv[i].first);
        p->key_row[kv[i].second] = (uint32_t)(p->row_value.size() - 1);
    }
}

/* the general way: one sort of (value, key) pairs */
void order_keys_by_sort(procell_plan* p)
{
    const size_t nb = p->bin_value.size();
    std::vector<std::pair<double, uint32_t>> kv;
    kv.reserve(p->n_keys);
    for (size_t b = 0; b < nb; ++b) {
        double f = p->bin_value[b];
        for (unsigned k = 0; k <= p->bin_kdiv[b]; ++k) {
            if (k > 0 || p->bin_count0[b]) kv.emplace_back(f, (uint32_t)(p->bin_keybase[b] + k));
            f = f / 2;
        }
    }
    std::sort(kv.begin(), kv.end(), [](const std::pair<double, uint32_t>& x, const std::pair<double, uint32_t>& y) {
        return x.first < y.first;
    });
    rows_from_sorted(p, kv);
}

/* The same order without sorting the keys.  Halving a normal double whose half is normal too only decrements the
 * exponent field, so key (b, k) has the bit pattern bits(value_b) - k * 2^52 and positive doubles order like their bit
 * patterns: sort the BINS by mantissa once, then bucket the keys by exponent field, filling every bucket in mantissa
 * order.  O(bins log bins + keys) instead of O(keys log keys) (config 2: 385 bins, 4 420 keys; it is the largest host
 * item of an end-to-end step).  Returns false - nothing touched - unless every key value is a positive normal double. */
bool order_keys_by_exponent(procell_plan* p)
{
    const size_t nb = p->bin_value.size();
    if (nb == 0) return false;
    std::vector<uint64_t> bits(nb);
    uint64_t e_min = ~0ull, e_max = 0;
    for (size_t b = 0; b < nb; ++b) {
        const double v = p->bin_value[b];
        memcpy(&bits[b], &v, 8);
        const uint64_t e = bits[b] >> 52;                     /* sign bit included: must be 0 */
        if (!(v > 0.0) || e == 0 || e >= 0x7FF || e <= p->bin_kdiv[b]) return false;
        e_min = std::min(e_min, e - p->bin_kdiv[b]);
        e_max = std::max(e_max, e);
    }
    std::vector<uint32_t> by_mant(nb);
    for (size_t b = 0; b < nb; ++b) by_mant[b] = (uint32_t)b;
    const uint64_t mant = (1ull << 52) - 1;
    std::sort(by_mant.begin(), by_mant.end(), [&](uint32_t x, uint32_t y) { return (bits[x] & mant) < (bits[y] & mant); });
    std::vector<uint32_t> start(e_max - e_min + 2, 0);        /* bucket e holds the keys with exponent field e_min + e */
    size_t n_kv = 0;
    for (size_t b = 0; b < nb; ++b) {
        const uint64_t e = bits[b] >> 52;
        for (unsigned k = p->bin_count0[b] ? 0 : 1; k <= p->bin_kdiv[b]; ++k) { ++start[e - k - e_min + 1]; ++n_kv; }
    }
    for (size_t i = 1; i < start.size(); ++i) start[i] += start[i - 1];
    std::vector<std::pair<double, uint32_t>> kv(n_kv);
    for (size_t i = 0; i < nb; ++i) {
        const uint32_t b = by_mant[i];
        const uint64_t e = bits[b] >> 52;
        for (unsigned k = p->bin_count0[b] ? 0 : 1; k <= p->bin_kdiv[b]; ++k) {
            const uint64_t kb = bits[b] - ((uint64_t)k << 52);
            double v;
            memcpy(&v, &kb, 8);
            kv[start[e - k - e_min]++] = std::make_pair(v, (uint32_t)(p->bin_keybase[b] + k));
        }
    }
    rows_from_sorted(p, kv);
    return true;
}

}  // namespace
}  // namespace procell_b200

using procell_b200::fail;
using procell_b200::order_keys_by_exponent;
using procell_b200::order_keys_by_sort;

extern "C" {

const char* procell_last_error(void) { return procell_b200::g_last_error.c_str(); }

const char* procell_version(void) { return "procell-b200 0.1 (sm_100a)"; }

void procell_free(void* p) { free(p); }

/* The seed cell's type uniform is u = (2x + 1) / 2^33 for a 32-bit random word x (procell_spec.h: pcs_u32unit), and its
 * type is the first j with u < cum[j] (cell.cu:81-104).  For a double c, "u < c" holds exactly for the x below
 * thr(c) = ceil(c * 2^33) >> 1, computed here in integers from c's mantissa and exponent (no rounding anywhere), so the
 * kernel compares 32-bit integers instead of doubles.  Returned clamped to [0, 2^32]. */
uint64_t procell_type_threshold(double cum)
{
    if (!(cum > 0.0)) return 0;
    int e = 0;
    const double f = frexp(cum, &e);                 /* cum = f * 2^e, f in [0.5, 1) */
    const uint64_t m = (uint64_t)ldexp(f, 53);       /* 53-bit integer mantissa: cum = m * 2^(e - 53) */
    const int s = e - 53 + 33;                       /* cum * 2^33 = m * 2^s */
    if (s >= 11) return 1ull << 32;                  /* cum * 2^33 >= 2^63: far beyond the last x */
    uint64_t r;                                      /* ceil(m * 2^s) */
    if (s >= 0) r = m << s;
    else if (s <= -64) r = 1;
    else r = (m >> -s) + ((m & ((1ull << -s) - 1ull)) ? 1ull : 0ull);
    const uint64_t thr = r >> 1;
    return thr > (1ull << 32) ? (1ull << 32) : thr;
}

int procell_check_proportions(const procell_cell_type* types, size_t n_types)
{
    double sum = 0.0;    /* thrust::reduce from 0.0, left to right (parser.cu:52-58) */
    for (size_t j = 0; j < n_types; ++j) sum = sum + types[j].proportion;
    double err = 1 / pow(10.0, 8.0);
    if (std::fabs(1.0 - sum) > err)
        return fail(PROCELL_ERR_PROPORTION, "ERROR: proportion distribution of cell types does not sum to 1, aborting.");
    return PROCELL_OK;
}

int procell_plan_create(const double* value, const uint64_t* freq, size_t n_lines, double phi, procell_plan** out)
{
    if (!out || (n_lines && (!value || !freq))) return fail(PROCELL_ERR_ARG, "procell_plan_create: null argument");
    if (!(phi >= 0.0)) return fail(PROCELL_ERR_ARG, "phi must be >= 0 (0 selects the default)");
    procell_plan* p = new procell_plan();
    if (phi == 0.0) {   /* parser.cu:80-96: smallest value whose frequency is > 0 */
        for (size_t i = 0; i < n_lines; ++i)
            if (freq[i] > 0 && (phi == 0.0 || value[i] < phi)) phi = value[i];
    }
    p->phi = phi;
    uint64_t total = 0;
    size_t n_keys = 0;
    p->bin_value.reserve(n_lines);
    p->bin_freq.reserve(n_lines);
    p->bin_start.reserve(n_lines + 1);
    p->bin_kdiv.reserve(n_lines);
    p->bin_count0.reserve(n_lines);
    p->bin_keybase.reserve(n_lines);
    for (size_t i = 0; i < n_lines; ++i) {
        if (freq[i] == 0) continue;   /* parser.cu:110-111 */
        p->bin_value.push_back(value[i]);
        p->bin_freq.push_back(freq[i]);
        p->bin_start.push_back(total);
        total += freq[i];
        /* how often may a cell of this bin halve: node rule f/2 > phi (proliferation.cu:323), evaluated with
         * the same repeated division by two the reference applies to the fluorescence itself */
        unsigned k = 0;
        double f = value[i];
        while (k < 63 && f / 2 > phi) { f = f / 2; ++k; }
        if (k == 63 && f / 2 > phi) p->depth_capped = true;
        p->bin_kdiv.push_back(static_cast<uint8_t>(k));
        p->bin_count0.push_back(static_cast<uint8_t>(value[i] >= phi));   /* parser.cu:127 */
        p->bin_keybase.push_back(static_cast<uint32_t>(n_keys));
        n_keys += k + 1;
        if (n_keys > 0x7FFFFFFFu) { delete p; return fail(PROCELL_ERR_ARG, "key space too large"); }
    }
    p->bin_start.push_back(total);
    p->n_cells = total;
    p->n_keys = n_keys;
    const size_t nb = p->bin_value.size();
    if (nb > 65535) { delete p; return fail(PROCELL_ERR_ARG, "more than 65535 non-empty histogram lines"); }
    if (total > 0xFFFFFFFFull) { delete p; return fail(PROCELL_ERR_ARG, "more than 2^32-1 seed cells"); }
    /* value of every key by repeated halving (parser.cu:126-137), then the ordered set of them (:142-151) */
    if (!order_keys_by_exponent(p)) order_keys_by_sort(p);
    *out = p;
    return PROCELL_OK;
}

void procell_plan_destroy(procell_plan* plan) { delete plan; }
size_t procell_plan_n_bins(const procell_plan* plan) { return plan ? plan->bin_value.size() : 0; }
size_t procell_plan_n_keys(const procell_plan* plan) { return plan ? plan->n_keys : 0; }

double procell_plan_lineage_depth(const procell_plan* plan, const procell_cell_type* types, size_t n_types, double t_max)
{
    if (!plan || !types || plan->n_cells == 0) return 0.0;
    double fastest = 0.0;
    for (size_t j = 0; j < n_types; ++j)
        if (types[j].mean > 0.0 && (fastest == 0.0 || types[j].mean < fastest)) fastest = types[j].mean;
    if (!(fastest > 0.0) || !(t_max > 0.0)) return 0.0;
    double halvings = 0.0;
    for (size_t b = 0; b + 1 < plan->bin_start.size(); ++b)
        halvings += (double)(plan->bin_kdiv[b] & 63u) * (double)(plan->bin_start[b + 1] - plan->bin_start[b]);
    halvings /= (double)plan->n_cells;
    const double generations = t_max / fastest;
    return generations < halvings ? generations : halvings;
}
size_t procell_plan_n_rows(const procell_plan* plan) { return plan ? plan->row_value.size() : 0; }
uint64_t procell_plan_n_cells(const procell_plan* plan) { return plan ? plan->n_cells : 0; }
double procell_plan_phi(const procell_plan* plan) { return plan ? plan->phi : 0.0; }
int procell_plan_depth_capped(const procell_plan* plan) { return plan && plan->depth_capped; }

int procell_plan_export(const procell_plan* plan, double* row_value, uint32_t* key_row, uint32_t* bin_keybase,
                        uint8_t* bin_kdiv)
{
    if (!plan) return fail(PROCELL_ERR_ARG, "procell_plan_export: null plan");
    if (row_value && !plan->row_value.empty())
        memcpy(row_value, plan->row_value.data(), plan->row_value.size() * sizeof(double));
    if (key_row && plan->n_keys) memcpy(key_row, plan->key_row.data(), plan->n_keys * sizeof(uint32_t));
    if (bin_keybase && !plan->bin_keybase.empty())
        memcpy(bin_keybase, plan->bin_keybase.data(), plan->bin_keybase.size() * sizeof(uint32_t));
    if (bin_kdiv && !plan->bin_kdiv.empty()) memcpy(bin_kdiv, plan->bin_kdiv.data(), plan->bin_kdiv.size());
    return PROCELL_OK;
}

int procell_merge_rows(const procell_plan* plan, const int64_t* counts, size_t n_types, int64_t* row_freq,
                       int64_t* row_ratio)
{
    if (!plan || !counts || !row_freq) return fail(PROCELL_ERR_ARG, "procell_merge_rows: null argument");
    const size_t n_rows = plan->row_value.size();
    std::fill(row_freq, row_freq + n_rows, int64_t(0));
    if (row_ratio) std::fill(row_ratio, row_ratio + n_rows * n_types, int64_t(0));
    for (size_t key = 0; key < plan->n_keys; ++key) {
        const uint32_t row = plan->key_row[key];
        if (row == 0xFFFFFFFFu) continue;
        for (size_t j = 0; j < n_types; ++j) {
            const int64_t c = counts[key * n_types + j];
            row_freq[row] += c;
            if (row_ratio) row_ratio[static_cast<size_t>(row) * n_types + j] += c;
        }
    }
    return PROCELL_OK;
}

}  // extern "C"
