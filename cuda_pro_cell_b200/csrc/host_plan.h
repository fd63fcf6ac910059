/* host_plan.h - host-side data model shared by the C-ABI translation units. */
#ifndef PROCELL_HOST_PLAN_H
#define PROCELL_HOST_PLAN_H

#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/procell_b200.h"

/* the result-key precomputation of io::load_fluorescences (reference src/io/parser.cu:68-154) */
struct procell_plan {
    std::vector<double> bin_value;       /* lines with frequency > 0, file order */
    std::vector<uint64_t> bin_freq;
    std::vector<uint64_t> bin_start;     /* n_bins + 1 running starts ("bounds") */
    std::vector<uint8_t> bin_kdiv;       /* #{k >= 0 : value / 2^(k+1) > phi}, capped at 63 */
    std::vector<uint8_t> bin_count0;     /* value >= phi */
    std::vector<uint32_t> bin_keybase;
    std::vector<uint32_t> key_row;       /* 0xFFFFFFFF: not an output row (level 0 of a bin below phi) */
    std::vector<double> row_value;       /* ascending */
    uint64_t n_cells = 0;
    size_t n_keys = 0;
    double phi = 0.0;
    bool depth_capped = false;
};

namespace procell_b200 {
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
}  // namespace procell_b200

#endif
