/* cli.cpp - the `procell` command line on top of the C ABI.
 *
 * Replaces the reference's src/main.cu:18-34 (driver), src/io/cmdargs.cpp:11-375 (option parser) and
 * src/simulation/simulator.cu:10-56 (four-step lifecycle).  Flags, their spellings, the messages and the exit
 * status follow the reference: every diagnostic goes to STDOUT and the process exits with status 1
 * (cmdargs.cpp:41-45,48-75).  Both the spellings the code accepts (--output-histogram, --phi-min) and the
 * ones the README documents (--output, --phi) are taken; -p is optional as the README says
 * (README.md:98-102; the reference code wrongly requires it, cmdargs.cpp:48-49), default = smallest value
 * with a non-zero frequency (parser.cu:80-96).  -d/--tree-depth (1..23) is accepted and ignored: it only
 * sized the reference's dense level arrays.  Extensions (not in the reference): --seed, --seeding,
 * --kernel, --device, --gpus, --shard-level, --checkpoints, --stats.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "host_plan.h"

namespace {

struct Args {
    bool h0_given = false, types_given = false, out_given = false, tmax_given = false, phi_given = false;
    bool depth_given = false, track_ratio = false, seed_given = false, stats = false;
    std::string h0, types, out;
    double t_max = 0.0, phi = 0.0;
    uint64_t seed = 0x5EED0000ull;
    int seeding = PROCELL_SEEDING_IDEAL;
    int kernel = PROCELL_KERNEL_COOP;
    int device = 0;
    int gpus = 1;
    int shard_level = 0;
    std::vector<double> checkpoints;
};

void usage()
{
    std::cout <<
        "procell - ProCell stochastic cell-proliferation simulator (B200 build)\n"
        "  -h, --histogram FILE         initial fluorescence histogram: <value> <frequency> per line (required)\n"
        "  -c, --cell-types FILE        subpopulations: <proportion> <mean> <stddev> per line, -1 -1 = quiescent (required)\n"
        "  -t, --time-max T             simulated time, >= 0 (required)\n"
        "  -o, --output[-histogram] FILE  result histogram (stdout when absent)\n"
        "  -p, --phi[-min] PHI          minimum fluorescence, > 0 (default: smallest value with frequency > 0)\n"
        "  -r, --track-ratio            add one count column per cell type\n"
        "  -d, --tree-depth N           accepted for compatibility (1..23), ignored\n"
        "      --seed N                 Philox key (default 0x5EED0000)\n"
        "      --seeding ideal|refcompat\n"
        "      --kernel coop|simple\n"
        "      --device N               CUDA device index\n"
        "      --gpus N                 shard the seed cells over N GPUs of this box (0 = all), one NCCL reduce\n"
        "      --shard-level L          with --gpus: shard SUBTREES at tree level L instead of whole lineages (deep trees)\n"
        "      --checkpoints T1,T2,...  also write the histogram at these earlier times (FILE.t<T>), one tree expansion\n"
        "      --stats                  print run statistics as JSON on stderr\n";
}

/* returns 0 = not this option, 1 = consumed, -1 = error already printed */
int take_value(int argc, char** argv, int& i, const char* long_name, const char* short_name, bool& given,
               const char* need, std::string& value)
{
    (void)short_name;
    if (given) {
        std::cout << "Option " << long_name << " already given" << std::endl;
        return -1;
    }
    if (i >= argc - 1) {
        std::cout << "Option " << long_name << " requires " << need << std::endl;
        return -1;
    }
    ++i;
    given = true;
    value = argv[i];
    return 1;
}

}  // namespace

extern "C" int procell_main(int argc, char** argv)
{
    Args a;
    for (int i = 1; i < argc; ++i) {
        const std::string s(argv[i]);
        std::string v;
        int r = 0;
        if (s == "--histogram" || s == "-h") {
            r = take_value(argc, argv, i, "--histogram (-h)", "-h", a.h0_given, "a filename", a.h0);
        } else if (s == "--cell-types" || s == "-c") {
            r = take_value(argc, argv, i, "--cell-types (-c)", "-c", a.types_given, "a filename", a.types);
        } else if (s == "--output-histogram" || s == "--output" || s == "-o") {
            r = take_value(argc, argv, i, "--output-histogram (-o)", "-o", a.out_given, "a filename", a.out);
        } else if (s == "--time-max" || s == "-t") {
            r = take_value(argc, argv, i, "--time-max (-t)", "-t", a.tmax_given, "an integer value >= 0", v);
            if (r == 1) {
                a.t_max = atof(v.c_str());   /* cmdargs.cpp:200 */
                if (!(a.t_max >= 0)) {
                    std::cout << "Option --time-max (-t) requires an integer value >= 0" << std::endl;
                    return 1;
                }
            }
        } else if (s == "--phi-min" || s == "--phi" || s == "-p") {
            r = take_value(argc, argv, i, "--phi-min (-p)", "-p", a.phi_given, "a double value > 0", v);
            if (r == 1) {
                a.phi = atof(v.c_str());     /* cmdargs.cpp:256 */
                if (!(a.phi > 0)) {
                    std::cout << "Option --phi-min (-p) requires a double value > 0" << std::endl;
                    return 1;
                }
            }
        } else if (s == "--tree-depth" || s == "-d") {
            r = take_value(argc, argv, i, "--tree-depth (-d)", "-d", a.depth_given, "an integer value >= 1", v);
            if (r == 1) {
                unsigned depth = 0;
                sscanf(v.c_str(), "%u", &depth);
                if (depth < 1 || depth > 23) {
                    std::cout << "Option --tree-depth (-d) requires an integer value >= 1 && <= 23" << std::endl;
                    return 1;
                }
            }
        } else if (s == "--track-ratio" || s == "-r") {
            if (a.track_ratio) {
                std::cout << "Option --track-ratio (-r) already given" << std::endl;
                return 1;
            }
            a.track_ratio = true;
            r = 1;
        } else if (s == "--seed") {
            r = take_value(argc, argv, i, "--seed", "", a.seed_given, "an unsigned integer", v);
            if (r == 1) a.seed = strtoull(v.c_str(), nullptr, 0);
        } else if (s == "--seeding" && i < argc - 1) {
            v = argv[++i];
            if (v == "ideal") a.seeding = PROCELL_SEEDING_IDEAL;
            else if (v == "refcompat") a.seeding = PROCELL_SEEDING_REFCOMPAT;
            else { std::cout << "Option --seeding requires ideal or refcompat" << std::endl; return 1; }
            r = 1;
        } else if (s == "--kernel" && i < argc - 1) {
            v = argv[++i];
            if (v == "coop") a.kernel = PROCELL_KERNEL_COOP;
            else if (v == "simple") a.kernel = PROCELL_KERNEL_SIMPLE;
            else { std::cout << "Option --kernel requires coop or simple" << std::endl; return 1; }
            r = 1;
        } else if (s == "--device" && i < argc - 1) {
            a.device = atoi(argv[++i]);
            r = 1;
        } else if (s == "--gpus" && i < argc - 1) {
            a.gpus = atoi(argv[++i]);
            r = 1;
        } else if (s == "--shard-level" && i < argc - 1) {
            a.shard_level = atoi(argv[++i]);
            if (a.shard_level < 0 || a.shard_level > 30) { std::cout << "Option --shard-level requires an integer value >= 0 && <= 30" << std::endl; return 1; }
            r = 1;
        } else if (s == "--checkpoints" && i < argc - 1) {
            /* extension: histograms at several times from one expansion; -o FILE gets the last one, FILE.t<time> the others */
            std::string list = argv[++i];
            size_t pos = 0;
            while (pos < list.size()) {
                size_t comma = list.find(',', pos);
                if (comma == std::string::npos) comma = list.size();
                a.checkpoints.push_back(atof(list.substr(pos, comma - pos).c_str()));
                pos = comma + 1;
            }
            r = 1;
        } else if (s == "--stats") {
            a.stats = true;
            r = 1;
        } else if (s == "--help") {
            usage();
            return 0;
        }
        if (r < 0) return 1;
        if (r == 0) {
            std::cout << "Invalid option " << s << std::endl;   /* cmdargs.cpp:43 */
            return 1;
        }
    }
    if (!a.h0_given || !a.types_given || !a.tmax_given) {       /* cmdargs.cpp:48-75, minus -p (README) */
        std::cout << "The following missing arguments are required:" << std::endl;
        if (!a.h0_given) std::cout << "--histogram (-h)" << std::endl;
        if (!a.types_given) std::cout << "--cell-types (-c)" << std::endl;
        if (!a.tmax_given) std::cout << "--t-max (-t)" << std::endl;
        return 1;
    }

    /* Simulator::load_params (simulator.cu:10-20) */
    double* value = nullptr;
    uint64_t* freq = nullptr;
    size_t n_lines = 0, n_types = 0;
    procell_cell_type* types = nullptr;
    procell_plan* plan = nullptr;
    int rc = procell_read_histogram(a.h0.c_str(), &value, &freq, &n_lines);
    if (rc == PROCELL_OK) rc = procell_plan_create(value, freq, n_lines, a.phi, &plan);
    if (rc == PROCELL_OK) rc = procell_read_cell_types(a.types.c_str(), &types, &n_types);
    if (rc == PROCELL_OK && n_types == 0) rc = procell_b200::fail(PROCELL_ERR_ARG, "no cell types given");
    std::vector<int64_t> counts, row_freq, row_ratio;
    procell_run_stats st;
    memset(&st, 0, sizeof st);
    if (rc == PROCELL_OK) {
        /* Simulator::create_cell_population + start_simulation (simulator.cu:22-38) */
        procell_sim_params sp;
        memset(&sp, 0, sizeof sp);
        sp.types = types; sp.n_types = n_types; sp.n_sets = 1; sp.t_max = a.t_max; sp.seed = a.seed;
        sp.seeding_mode = a.seeding; sp.kernel = a.kernel; sp.shard_level = (uint32_t)a.shard_level;
        if (!a.checkpoints.empty()) {
            if (a.checkpoints.back() != a.t_max) a.checkpoints.push_back(a.t_max);     /* -t is always the last one */
            sp.checkpoints = a.checkpoints.data();
            sp.n_checkpoints = a.checkpoints.size();
        }
        const size_t n_cp = a.checkpoints.empty() ? 1 : a.checkpoints.size();
        counts.assign(n_cp * procell_plan_n_keys(plan) * n_types + 1, 0);
        if (a.gpus == 1) rc = procell_proliferate(plan, &sp, a.device, counts.data(), nullptr, &st);
        else rc = procell_proliferate_multi(plan, &sp, a.gpus, counts.data(), nullptr, &st);
    }
    if (rc == PROCELL_OK) {
        /* Simulator::save_results (simulator.cu:40-56) */
        const size_t n_rows = procell_plan_n_rows(plan);
        std::vector<double> row_value(n_rows + 1);
        row_freq.assign(n_rows + 1, 0);
        row_ratio.assign(n_rows * n_types + 1, 0);
        procell_plan_export(plan, row_value.data(), nullptr, nullptr, nullptr);
        const size_t n_cp = a.checkpoints.empty() ? 1 : a.checkpoints.size();
        const size_t per_cp = procell_plan_n_keys(plan) * n_types;
        for (size_t j = 0; j < n_cp && rc == PROCELL_OK; ++j) {
            const bool last = j + 1 == n_cp;
            if (!last && !a.out_given) continue;          /* earlier checkpoints need files to go to */
            rc = procell_merge_rows(plan, counts.data() + j * per_cp, n_types, row_freq.data(), row_ratio.data());
            std::string path = a.out;
            if (!last) {
                char suffix[64];
                snprintf(suffix, sizeof suffix, ".t%.10g", a.checkpoints[j]);
                path += suffix;
            }
            if (rc == PROCELL_OK)
                rc = procell_write_histogram(a.out_given ? path.c_str() : nullptr, a.track_ratio ? 1 : 0, n_types, n_rows,
                                             row_value.data(), row_freq.data(), row_ratio.data());
        }
    }
    if (rc == PROCELL_OK && a.stats) {
        fprintf(stderr, "{\"divisions\": %lld, \"kernel_ms\": %.6f, \"grid\": %d, \"block\": %d, \"smem_bytes\": %d, "
                        "\"cells\": %llu, \"phi\": %.17g}\n",
                (long long)st.divisions, st.kernel_ms, st.grid, st.block, st.smem_bytes,
                (unsigned long long)procell_plan_n_cells(plan), procell_plan_phi(plan));
    }
    if (rc != PROCELL_OK) std::cout << procell_last_error() << std::endl;
    procell_plan_destroy(plan);
    procell_free(value);
    procell_free(freq);
    procell_free(types);
    return rc == PROCELL_OK ? 0 : 1;
}
