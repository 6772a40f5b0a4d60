// Host side of the C++ twin `rlsim_gpu` (same flags / FASTA / report as the Go reference, compute through the C ABI).
// Mirrors /root/reference/src/{args.go, target_mix.go, parse_raw.go, fasta.go, input.go, report.go, fragstats.go, logging.go, main.go}.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>
#include "../../../include/rlsim_gpu.h"

namespace host {

[[noreturn]] void fatal(const std::string &msg);          // logging.go:67-73: log + exit(1)
void log_line(const std::string &msg);                    // logging.go: timestamped stderr line
extern bool g_verbose;
inline void logv(const std::string &msg) { if (g_verbose) log_line(msg); }

struct MixComp {                                          // target_mix.go:37-44
    std::string type;
    double location = 0, scale = 0, shape = 0, weight = 1;
    uint64_t low = 0, high = 0;
    std::string str() const;                              // target_mix.go:46-53
};
std::vector<MixComp> parse_mixture(const std::string &s); // target_mix.go:90-214
std::string mixture_string(const std::vector<MixComp> &m);

struct RawParams {                                        // parse_raw.go:39-50
    int64_t req_frags = 0, nr_cycles = 0;
    std::vector<uint32_t> lengths;
    std::vector<double> probs, gc_effs;
};

struct CmdArgs {                                          // args.go:52-80
    int64_t req_frags = 0, nr_cycles = 11, max_procs = 4, init_seed = 0, pcr_seed = 0, sampling_seed = 0;
    double strand_bias = 0.5, priming_temp = 5.0, fixed_eff = 0.0, frag_loss_prob = 0.0, min_raw_gc_eff = 0.0, expr_mul = 1.0;
    bool have_gc_param = false, verbose = false, have_raw = false;
    double gc_shape = 0, gc_min = 0, gc_max = 0, len_shape = 0, len_min = 1, len_max = 1;
    std::string report_file = "rlsim_report.json", frag_method = "after_prim_double", gob_dir, prof_file, raw_params_file, rng = "philox";
    int frag_param = 0, kmer_length = 6, gc_freq = 100;
    std::vector<MixComp> target_mix, polya;
    std::vector<std::string> input_files;
    std::vector<double> raw_gc_effs;                      // empty => parametric (-eg)
    RawParams raw;
    std::vector<std::string> warnings;
    void parse(int argc, char **argv);                    // args.go:83-270 (exits for -h / -V / no fragments)
    rlsim_params to_params() const;
};

struct Transcripts {                                      // input.go:103-124 result
    std::vector<std::string> names;
    std::vector<uint8_t> bases;                           // concatenated, upper-cased, de-spaced
    std::vector<uint64_t> offsets, levels;
    std::map<uint32_t, uint64_t> tr_lengths, expr_levels; // fragstats.go TrLengths / ExprLevels
};
std::string read_input(const std::vector<std::string> &files);                          // fasta.go:98-104 (io.MultiReader) or stdin
Transcripts load_transcripts(const std::string &data, double expr_mul);                 // fasta.go:113-212 + input.go:82-124 on the host
bool load_transcripts_device(rlsim_handle *h, const std::string &data, double expr_mul, Transcripts *out);  // the same on the GPU (bases stay there)

class Report {                                            // report.go:46-148
  public:
    explicit Report(std::string path) : path_(std::move(path)) {}
    void add_u64(const std::map<uint32_t, uint64_t> &m, const char *xl, const char *yl, const char *title, const char *vis);
    void add_f64(const std::map<uint32_t, double> &m, const char *xl, const char *yl, const char *title, const char *vis);
    void add_xy(const std::vector<double> &x, const std::vector<double> &y, const char *xl, const char *yl, const char *title, const char *vis);
    void write_json();
  private:
    struct Panel { std::string xl, yl, vis; int pos; std::vector<std::pair<std::string, std::string>> data; };
    std::string path_;
    std::map<std::string, Panel> pool_;
    int cursor_ = 0;
};

double go_pow(double x, double y);                        // structure of Go's math.Pow (thermocycler.go:85-86,155,177)

}  // namespace host
