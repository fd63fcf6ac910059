/* textio.cpp - streaming reader / writer of the reference's three text formats (host only, no CUDA).
 *
 * Replaces the reference's src/io/parser.cu reading loops (:103-106 histogram, :167-175 cell types) and its
 * writer save_fluorescences (:187-217), which go through std::istream::operator>> and std::ostream::operator<<
 * token by token.  Formats and semantics are kept byte for byte:
 *   histogram : whitespace-separated "<double value> <uint64 frequency>" pairs, read until the first parse
 *               failure; order preserved; duplicates allowed (frequency-0 lines are KEPT here; the plan skips them)
 *   types     : "<proportion> <mean> <stddev>" triples, type id = 0-based line index
 *   output    : rows with frequency > 0, value printed with precision 10 (== %.10g), TAB, frequency, then with -r
 *               one TAB-separated count per type in file order, NEWLINE
 *
 * Reader: the file is pulled in with one read() and scanned in place.  A token in CANONICAL form -
 *   double:  [+-]? ( digits [ . digits* ]? | . digits ) ( [eE] [+-]? digits )?      uint64: digits, no overflow
 * - is converted with std::from_chars (correctly rounded, like the strtod behind operator>>).  The first token that
 * is not canonical (a sign on a frequency, "inf", a bare "e", overflow ...) hands the REST of the buffer, from the
 * start of the record it belongs to, to the very operator>> loop the reference runs, so every corner of the
 * iostream semantics (where reading stops, what a half-read record does) is the reference's by construction.
 * Writer: std::to_chars (general, precision 10) is specified as printf("%.10g") in the C locale, which is what
 * ostream::operator<<(double) does at precision(10); rows are formatted into one 1 MiB buffer per write().
 * tests/test_textio.py checks both against an iostream restatement of the reference (tests/textio_ref.cpp). */
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "host_plan.h"

namespace procell_b200 {

namespace {

/* the classic "C" locale whitespace set that operator>> skips */
inline bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

enum Scan { kOk, kEnd, kFallback };

struct Cursor {
    const char* p;
    const char* end;
    void skip_space() { while (p < end && is_space(*p)) ++p; }
};

/* canonical double at c.p (after whitespace); advances past the token on kOk */
Scan scan_double(Cursor& c, double* out)
{
    c.skip_space();
    if (c.p == c.end) return kEnd;
    const char* s = c.p;
    const char* first = s;                    /* what from_chars gets: it takes '-' but, unlike operator>>, no '+' */
    if (*s == '+') { ++s; first = s; }
    else if (*s == '-') ++s;
    const char* m0 = s;
    while (s < c.end && is_digit(*s)) ++s;
    size_t n_digits = (size_t)(s - m0);
    if (s < c.end && *s == '.') {
        ++s;
        const char* f0 = s;
        while (s < c.end && is_digit(*s)) ++s;
        n_digits += (size_t)(s - f0);
    }
    if (n_digits == 0) return kFallback;
    if (s < c.end && (*s == 'e' || *s == 'E')) {
        const char* e = s + 1;
        if (e < c.end && (*e == '+' || *e == '-')) ++e;
        const char* d0 = e;
        while (e < c.end && is_digit(*e)) ++e;
        if (e == d0) return kFallback;        /* "1e", "1e+": the stream extracts the 'e' and then fails */
        s = e;
    }
    /* a token that runs into the end of the buffer is complete too: the whole file is in memory */
    double v = 0.0;
    const std::from_chars_result r = std::from_chars(first, s, v, std::chars_format::general);
    if (r.ec != std::errc() || r.ptr != s) return kFallback;     /* out of range, or a form the two disagree on */
    *out = v;
    c.p = s;
    return kOk;
}

/* canonical uint64 at c.p: plain digits that fit */
Scan scan_u64(Cursor& c, uint64_t* out)
{
    c.skip_space();
    if (c.p == c.end) return kEnd;
    const char* s = c.p;
    uint64_t v = 0;
    int n = 0;
    while (s < c.end && is_digit(*s) && n < 19) { v = v * 10u + (uint64_t)(*s - '0'); ++s; ++n; }
    if (n == 0) return kFallback;
    if (s < c.end && is_digit(*s)) {          /* 20 digits or more: let from_chars decide about overflow */
        const char* e = s;
        while (e < c.end && is_digit(*e)) ++e;
        const std::from_chars_result r = std::from_chars(c.p, e, v, 10);
        if (r.ec != std::errc() || r.ptr != e) return kFallback;
        s = e;
    }
    *out = v;
    c.p = s;
    return kOk;
}

int slurp(const char* path, const char* what, std::string* buf)
{
    FILE* f = fopen(path, "rb");
    if (!f) return fail(PROCELL_ERR_IO, std::string("cannot open ") + what + " file " + path);
    char chunk[1 << 16];
    size_t n;
    while ((n = fread(chunk, 1, sizeof chunk, f)) > 0) buf->append(chunk, n);
    const bool bad = ferror(f) != 0;
    fclose(f);
    if (bad) return fail(PROCELL_ERR_IO, std::string("read error on ") + what + " file " + path);
    return PROCELL_OK;
}

template <class T>
int to_malloc(const std::vector<T>& v, T** out)
{
    *out = static_cast<T*>(malloc((v.size() + 1) * sizeof(T)));
    if (!*out) return fail(PROCELL_ERR_ARG, "out of host memory");
    if (!v.empty()) memcpy(*out, v.data(), v.size() * sizeof(T));
    return PROCELL_OK;
}

}  // namespace

/* parser.cu:103-106 on a memory buffer */
void parse_histogram_text(const char* text, size_t len, std::vector<double>* value, std::vector<uint64_t>* freq)
{
    Cursor c{ text, text + len };
    for (;;) {
        const char* record = c.p;
        double x;
        uint64_t n;
        Scan s = scan_double(c, &x);
        if (s == kOk) s = scan_u64(c, &n);
        if (s == kOk) { value->push_back(x); freq->push_back(n); continue; }
        if (s == kFallback) {                 /* the reference's own loop on the rest, from this record on */
            std::istringstream in(std::string(record, (size_t)(c.end - record)));
            while (in >> x >> n) { value->push_back(x); freq->push_back(n); }
        }
        return;
    }
}

/* parser.cu:167-175 on a memory buffer */
void parse_cell_types_text(const char* text, size_t len, std::vector<procell_cell_type>* types)
{
    Cursor c{ text, text + len };
    for (;;) {
        const char* record = c.p;
        procell_cell_type t;
        Scan s = scan_double(c, &t.proportion);
        if (s == kOk) s = scan_double(c, &t.mean);
        if (s == kOk) s = scan_double(c, &t.stddev);
        if (s == kOk) { types->push_back(t); continue; }
        if (s == kFallback) {
            std::istringstream in(std::string(record, (size_t)(c.end - record)));
            while (in >> t.proportion >> t.mean >> t.stddev) types->push_back(t);
        }
        return;
    }
}

/* parser.cu:187-217: rows are formatted into one buffer that goes to the file in ~1 MiB pieces */
namespace {
struct RowWriter {
    FILE* f;
    std::vector<char> buf;
    size_t n = 0;
    bool ok = true;
    explicit RowWriter(FILE* file) : f(file), buf((1u << 20) + 4096) {}
    void flush()
    {
        if (n && fwrite(buf.data(), 1, n, f) != n) ok = false;
        n = 0;
    }
    /* room for one number (a double at precision 10 is at most 17 characters, an int64 20) and a separator */
    char* reserve() { if (n + 64 > buf.size()) flush(); return buf.data() + n; }
    void put_double(double v)
    {
        char* p = reserve();
        n += (size_t)(std::to_chars(p, p + 48, v, std::chars_format::general, 10).ptr - p);
    }
    void put_int(int64_t v)
    {
        char* p = reserve();
        n += (size_t)(std::to_chars(p, p + 48, v).ptr - p);
    }
    void put_char(char ch) { *reserve() = ch; ++n; }
};
}  // namespace

}  // namespace procell_b200

using procell_b200::fail;

extern "C" {

int procell_read_histogram(const char* path, double** value, uint64_t** freq, size_t* n_lines)
{
    if (!path || !value || !freq || !n_lines) return fail(PROCELL_ERR_ARG, "procell_read_histogram: null argument");
    std::string text;
    int rc = procell_b200::slurp(path, "histogram", &text);
    if (rc != PROCELL_OK) return rc;
    std::vector<double> v;
    std::vector<uint64_t> f;
    v.reserve(text.size() / 12 + 1);
    f.reserve(text.size() / 12 + 1);
    procell_b200::parse_histogram_text(text.data(), text.size(), &v, &f);
    *n_lines = v.size();
    rc = procell_b200::to_malloc(v, value);
    if (rc == PROCELL_OK) rc = procell_b200::to_malloc(f, freq);
    return rc;
}

int procell_parse_histogram(const char* text, size_t len, double** value, uint64_t** freq, size_t* n_lines)
{
    if ((len && !text) || !value || !freq || !n_lines) return fail(PROCELL_ERR_ARG, "procell_parse_histogram: null argument");
    std::vector<double> v;
    std::vector<uint64_t> f;
    procell_b200::parse_histogram_text(text, len, &v, &f);
    *n_lines = v.size();
    int rc = procell_b200::to_malloc(v, value);
    if (rc == PROCELL_OK) rc = procell_b200::to_malloc(f, freq);
    return rc;
}

int procell_parse_cell_types(const char* text, size_t len, procell_cell_type** types, size_t* n_types)
{
    if ((len && !text) || !types || !n_types) return fail(PROCELL_ERR_ARG, "procell_parse_cell_types: null argument");
    std::vector<procell_cell_type> t;
    procell_b200::parse_cell_types_text(text, len, &t);
    *n_types = t.size();
    int rc = procell_b200::to_malloc(t, types);
    if (rc != PROCELL_OK) return rc;
    return procell_check_proportions(*types, *n_types);
}

int procell_read_cell_types(const char* path, procell_cell_type** types, size_t* n_types)
{
    if (!path || !types || !n_types) return fail(PROCELL_ERR_ARG, "procell_read_cell_types: null argument");
    std::string text;
    int rc = procell_b200::slurp(path, "cell types", &text);
    if (rc != PROCELL_OK) return rc;
    std::vector<procell_cell_type> t;
    procell_b200::parse_cell_types_text(text.data(), text.size(), &t);
    *n_types = t.size();
    rc = procell_b200::to_malloc(t, types);
    if (rc != PROCELL_OK) return rc;
    return procell_check_proportions(*types, *n_types);
}

int procell_write_histogram(const char* path, int save_ratio, size_t n_types, size_t n_rows,
                            const double* row_value, const int64_t* row_freq, const int64_t* row_ratio)
{
    if (n_rows && (!row_value || !row_freq)) return fail(PROCELL_ERR_ARG, "procell_write_histogram: null rows");
    if (save_ratio && n_rows && !row_ratio) return fail(PROCELL_ERR_ARG, "procell_write_histogram: null ratios");
    FILE* f = stdout;
    if (path) {
        f = fopen(path, "wb");
        if (!f) return fail(PROCELL_ERR_IO, std::string("cannot open output file ") + path);
    }
    procell_b200::RowWriter w(f);
    for (size_t i = 0; i < n_rows; ++i) {
        if (row_freq[i] <= 0) continue;       /* parser.cu:199 */
        w.put_double(row_value[i]);
        w.put_char('\t');
        w.put_int(row_freq[i]);
        if (save_ratio)
            for (size_t j = 0; j < n_types; ++j) { w.put_char('\t'); w.put_int(row_ratio[i * n_types + j]); }
        w.put_char('\n');
    }
    w.flush();
    bool ok = w.ok && fflush(f) == 0;
    if (path && fclose(f) != 0) ok = false;
    if (!ok) return fail(PROCELL_ERR_IO, "write failed");
    return PROCELL_OK;
}

}  // extern "C"
