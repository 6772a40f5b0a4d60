// rlsim_gpu: the rlsim driver (main.go:41-162) over the GPU pool - same flags, same FASTA in / out, same JSON report.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <map>
#include <string>
#include <vector>
#include "host.hpp"

using namespace host;

static void check(rlsim_handle *h, int rc) {
    if (rc) fatal(rlsim_last_error(h));  // logging.go:67-73: every error is fatal
}
static std::map<uint32_t, uint64_t> hist(rlsim_handle *h, int kind, uint32_t n, bool keep_zero_of = false, const std::map<uint32_t, uint64_t> *keys = nullptr) {
    std::vector<uint64_t> buf(n, 0);
    check(h, rlsim_get_hist(h, kind, buf.data(), n));
    std::map<uint32_t, uint64_t> m;
    if (keep_zero_of && keys) { for (auto &kv : *keys) m[kv.first] = kv.first < n ? buf[kv.first] : 0; return m; }
    for (uint32_t i = 0; i < n; i++) if (buf[i]) m[i] = buf[i];
    return m;
}

int main(int argc, char **argv) {
    CmdArgs args;
    args.parse(argc, argv);
    g_verbose = args.verbose;
    for (auto &w : args.warnings) log_line(w);
    char line[512];
    snprintf(line, sizeof line, "Starting up rlsim using maximum %lld cores.", (long long)args.max_procs); logv(line);
    if (args.init_seed == 0) args.init_seed = (int64_t)time(nullptr);  // main.go:76-79
    snprintf(line, sizeof line, "Initial random seed: %lld", (long long)args.init_seed); logv(line);
    if (!args.raw_params_file.empty()) logv("Using raw parameter file: " + args.raw_params_file);
    Report report(args.report_file);

    rlsim_handle *h = nullptr;
    if (rlsim_create(0, &h)) fatal(rlsim_last_error(nullptr));
    rlsim_params p = args.to_params();
    check(h, rlsim_set_model(h, &p));
    if (args.have_raw) check(h, rlsim_set_raw_target(h, args.raw.lengths.data(), args.raw.probs.data(), (uint32_t)args.raw.lengths.size()));
    snprintf(line, sizeof line, "Number of requested fragments: %lld", (long long)args.req_frags); logv(line);
    if (!args.have_raw) { logv("Target fragment length distribution components:"); logv(mixture_string(args.target_mix)); }
    check(h, rlsim_sample_target(h, args.req_frags));   // NewTarget / NewRawTarget
    logv("Finished sampling target lengths.");
    rlsim_stats st;
    check(h, rlsim_get_stats(h, &st));
    uint32_t nl = (uint32_t)st.high + 2;
    auto target = hist(h, RLSIM_H_TARGET, nl);
    report.add_u64(target, "Length", "Count", "Target lenghts", "bar");
    snprintf(line, sizeof line, "Fragmentation method: \"%s\" with parameter %d.", args.frag_method.c_str(), args.frag_param); logv(line);
    snprintf(line, sizeof line, "Number of PCR cycles: %lld", (long long)args.nr_cycles); logv(line);
    if (!(args.fixed_eff > 0.0)) {  // thermocycler.go:195-221
        std::vector<double> xs, ys;
        double gc = 0.0;
        for (int i = 0; i <= 200; i++) {
            double g = gc / 100.0, y;
            if (p.gc_mode == 1) y = p.gc_raw[(int)(g * 100)];
            else y = p.gc_min + (p.gc_max - p.gc_min) * go_pow(1 - go_pow(g, p.gc_shape), p.gc_shape);
            xs.push_back(gc); ys.push_back(y);
            gc += 0.5;
        }
        report.add_xy(xs, ys, "GC content (%)", "Efficiency", "GC efficiency function", "line");
        std::map<uint32_t, double> le;
        std::vector<double> table;
        double A = 0, B = 0;
        if (p.len_shape != 0.0) {  // thermocycler.go:83-92
            double alpha = go_pow((double)st.low, -p.len_shape), beta = go_pow((double)st.high, -p.len_shape);
            A = (p.len_max - p.len_min) / (alpha - beta);
            B = p.len_min - A * beta;
        }
        for (uint64_t l = st.low; l <= st.high; l++) {
            double e = p.len_shape == 0.0 ? 1.0 : A * go_pow((double)l, -p.len_shape) + B;
            le[(uint32_t)l] = e;
            table.push_back(e);
        }
        report.add_f64(le, "Length", "Efficiency", "Length efficiency function", "line");
        // the library amplifies with exactly the values the report shows
        check(h, rlsim_set_len_eff(h, table.data(), st.low, st.high));
    }
    logv("Poly(A) tail distribution components:");
    logv(mixture_string(args.polya));
    const std::string data = read_input(args.input_files);
    logv("Fragmenting transcripts and amplifying fragments:");
    Transcripts tr;
    if (getenv("RLSIM_HOST_FASTA") || !load_transcripts_device(h, data, args.expr_mul, &tr)) {
        check(h, rlsim_clear_transcripts(h));
        tr = load_transcripts(data, args.expr_mul);
        if (!tr.levels.empty())
            check(h, rlsim_add_transcripts(h, tr.bases.data(), tr.offsets.data(), tr.levels.data(), (uint32_t)tr.levels.size()));
    }
    check(h, rlsim_build_library(h));                   // Pool.InitTranscripts
    size_t expressed = 0;
    for (auto l : tr.levels) expressed += l > 0;
    snprintf(line, sizeof line, "Initialized %zu transcripts.", expressed); logv(line);
    check(h, rlsim_sample(h, nullptr, nullptr, 0));     // Sampler.SampleFragments
    check(h, rlsim_get_stats(h, &st));
    // FASTA records formatted on the device (frag.go:123-126, sampler.go:206)
    std::vector<uint8_t> names;
    std::vector<uint64_t> name_off{0};
    for (auto &n : tr.names) { names.insert(names.end(), n.begin(), n.end()); name_off.push_back(names.size()); }
    if (names.empty()) names.push_back(0);
    uint64_t size = 0;
    check(h, rlsim_format_fasta(h, names.data(), name_off.data(), nullptr, 0, &size));
    std::vector<uint8_t> out(size ? size : 1);
    if (size) check(h, rlsim_format_fasta(h, names.data(), name_off.data(), out.data(), size, &size));
    fwrite(out.data(), 1, size, stdout);
    double total = (double)st.pool_total;
    snprintf(line, sizeof line, "Total number of fragments in the pool: %g", total); logv(line);
    snprintf(line, sizeof line, "Sampling ratio: %g", (double)st.n_sampled / total); logv(line);
    snprintf(line, sizeof line, "Sampled %llu fragments.", (unsigned long long)st.n_sampled); logv(line);
    snprintf(line, sizeof line, "Missing fragments: %llu", (unsigned long long)((uint64_t)args.req_frags - st.n_sampled)); logv(line);
    // fragstats.go:120-141
    auto after_pcr = hist(h, RLSIM_H_AFTER_PCR, nl);
    report.add_u64(hist(h, RLSIM_H_AFTER_FRAG, nl), "Length", "Count", "Fragdist after fragmentation", "bar");
    report.add_u64(after_pcr, "Length", "Count", "Fragdist after PCR", "bar");
    report.add_u64(hist(h, RLSIM_H_AFTER_SAMPLING, nl, true, &after_pcr), "Length", "Count", "Fragdist after sampling", "bar");
    report.add_u64(hist(h, RLSIM_H_MISSING, nl, true, &target), "Length", "Count", "Missing fragments", "bar");
    report.add_u64(hist(h, RLSIM_H_SAMPLED, nl, true, &target), "Length", "Count", "Sampled fragments", "bar");
    report.add_u64(hist(h, RLSIM_H_POLYA, st.polya_max + 1), "Length", "Count", "Poly-A tail lengths", "bar");
    report.add_u64(tr.tr_lengths, "Length", "Count", "Transcript lengths", "hist");
    report.add_u64(tr.expr_levels, "Expression level", "Count", "Expression levels", "hist");
    report.add_xy({0.0}, {(double)st.n_sampled / total}, "", "Sampling ratio", "Sampling ratio", "bar");
    report.write_json();
    rlsim_destroy(h);                                   // Pool.Cleanup
    return 0;
}
