/* ref_host_shim.cu - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A C entry into the reference's OWN host-side text I/O, so that tests can pin this repo's reader, plan and
 * writer to the reference's code itself rather than to a restatement of it.  `make -C oracle refhost` compiles the
 * reference's translation units where they lie under /root/reference (nothing is copied) together with this file
 * into oracle/_ref/libref_host.so.  The three functions called below are pure host code (std::ifstream,
 * std::map, thrust::host_vector); they run without a GPU:
 *   io::load_fluorescences   src/io/parser.cu:68-154   (reading loop, default phi, bounds, result-row set)
 *   io::load_cell_types      src/io/parser.cu:156-185  (reading loop, proportion check, sort by proportion;
 *                                                       its stray cudaMalloc just fails without a device)
 *   io::save_fluorescences   src/io/parser.cu:187-217  (output format)
 * load_cell_types exit()s on a bad proportion sum: callers must pass files that sum to 1 (the rejection itself is
 * pinned against the reference binary in tests/test_host_boundary.py).
 */
#include <cstdint>
#include <cstddef>
#include <fstream>

#include "io/parser.h"

using namespace procell;

extern "C" int ref_load_fluorescences(const char* path, double threshold_in, double* threshold_out, uint64_t* total,
                                      double* value, uint64_t* freq, uint64_t* bound, size_t cap, size_t* n_bins,
                                      double* row_value, size_t row_cap, size_t* n_rows)
{
    simulation::fluorescences data;
    simulation::initial_bounds bounds;
    simulation::fluorescences_result rows;
    double_t threshold = threshold_in;
    uint64_t size = 0;
    io::load_fluorescences(path, data, bounds, rows, threshold, &size);
    *threshold_out = threshold;
    *total = size;
    *n_bins = data.size();
    *n_rows = rows.size();
    if (data.size() > cap || rows.size() > row_cap) return -1;
    for (size_t i = 0; i < data.size(); ++i) { value[i] = data[i].value; freq[i] = data[i].frequency; bound[i] = bounds[i]; }
    for (size_t i = 0; i < rows.size(); ++i) row_value[i] = rows[i].value;
    return 0;
}

extern "C" int ref_load_cell_types(const char* path, int32_t* name, double* proportion, double* mean, double* sd,
                                   size_t cap, size_t* n_types)
{
    simulation::cell_types data;
    io::load_cell_types(path, data);
    *n_types = data.size();
    if (data.size() > cap) return -1;
    for (size_t i = 0; i < data.size(); ++i) {
        name[i] = data[i].name; proportion[i] = data[i].proportion; mean[i] = data[i].timer; sd[i] = data[i].sigma;
    }
    return 0;
}

extern "C" int ref_save_fluorescences(const char* path, int save_ratio, int32_t ratio_size, const double* value,
                                      const uint64_t* freq, const int32_t* ratio, size_t n_rows)
{
    simulation::fluorescences_result rows;
    for (size_t i = 0; i < n_rows; ++i) {
        simulation::fluorescence_with_ratio r;
        r.value = value[i];
        r.frequency = freq[i];
        r.ratio = const_cast<int32_t*>(ratio ? ratio + i * (size_t)ratio_size : nullptr);
        rows.push_back(r);
    }
    std::ofstream out(path, std::ios::binary);
    if (!out) return -1;
    io::save_fluorescences(out, save_ratio != 0, ratio_size, rows);
    return out.good() ? 0 : -2;
}
