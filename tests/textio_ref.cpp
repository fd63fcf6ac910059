/* textio_ref.cpp - TEST INFRASTRUCTURE ONLY: an iostream restatement of the reference's text reading loops and
 * writer, the checker for cuda_pro_cell_b200/csrc/textio.cpp.  Built by tests/test_textio.py into tests/_build/.
 *
 *   ref_read_histogram   src/io/parser.cu:103-106   while (in >> value >> frequency)      (double, uint64_t)
 *   ref_read_cell_types  src/io/parser.cu:167-175   while (in >> proportion >> timer >> sigma)
 *   ref_write_histogram  src/io/parser.cu:187-217   precision(10); value TAB frequency [TAB ratio...] endl,
 *                                                   rows with frequency > 0 only
 * Nothing of the product links or calls this file. */
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

/* returns the number of records; fills up to `cap` of them */
static size_t read_histogram(std::istream& in, double* value, uint64_t* freq, size_t cap)
{
    double v = 0.0;
    uint64_t f = 0;
    size_t n = 0;
    while (in >> v >> f) {
        if (n < cap) { value[n] = v; freq[n] = f; }
        ++n;
    }
    return n;
}

static size_t read_cell_types(std::istream& in, double* triples, size_t cap)
{
    double p = 0.0, t = 0.0, s = 0.0;
    size_t n = 0;
    while (in >> p >> t >> s) {
        if (n < cap) { triples[3 * n] = p; triples[3 * n + 1] = t; triples[3 * n + 2] = s; }
        ++n;
    }
    return n;
}

extern "C" {

size_t ref_read_histogram(const char* path, double* value, uint64_t* freq, size_t cap)
{
    std::ifstream in(path);
    return read_histogram(in, value, freq, cap);
}

size_t ref_read_cell_types(const char* path, double* triples, size_t cap)
{
    std::ifstream in(path);
    return read_cell_types(in, triples, cap);
}

/* the same loops on text in memory (fuzzing without one file per case) */
size_t ref_parse_histogram(const char* text, size_t len, double* value, uint64_t* freq, size_t cap)
{
    std::istringstream in(std::string(text, len));
    return read_histogram(in, value, freq, cap);
}

size_t ref_parse_cell_types(const char* text, size_t len, double* triples, size_t cap)
{
    std::istringstream in(std::string(text, len));
    return read_cell_types(in, triples, cap);
}

int ref_write_histogram(const char* path, int save_ratio, int32_t ratio_size, size_t n_rows, const double* value,
                        const uint64_t* frequency, const int64_t* ratio)
{
    std::ofstream stream(path);
    if (!stream.is_open()) return -1;
    stream.precision(10);
    for (size_t i = 0; i < n_rows; ++i) {
        if (frequency[i] > 0) {
            stream << value[i] << "\t" << frequency[i];
            if (save_ratio)
                for (int32_t j = 0; j < ratio_size; ++j) stream << "\t" << ratio[i * (size_t)ratio_size + j];
            stream << std::endl;
        }
    }
    return 0;
}

}  // extern "C"
