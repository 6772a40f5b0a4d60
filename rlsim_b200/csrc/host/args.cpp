// CLI flags and parameter files of rlsim, unchanged (args.go, target_mix.go, parse_raw.go).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unistd.h>
#include "host.hpp"

namespace host {

static const char *VERSION = "1.3";  // version.go:34
// Raw GC efficiencies estimated from SRR521457: the reference's built-in default for -eg (args.go:194); part of the CLI contract.
static const double EFF_SRR521457[101] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.6214855142919264, 0.4159233164222049, 0.44285355190022124,
    0.5111795670132255, 0.5261218695690415, 0.5657937478213575, 0.6057227372065885, 0.652512493251505, 0.6551266520303076,
    0.6586493220034146, 0.6710292671335725, 0.6828989286775027, 0.7103061475969692, 0.7145967526946446, 0.7125612213162864,
    0.7158400735145201, 0.7361963080537015, 0.7507700417579295, 0.7936248041169689, 0.8745263930903109, 0.8971210947107862,
    0.8979381841721974, 0.8723089598770069, 0.8768516147049334, 0.8892267818464259, 0.9064925780415922, 0.9077163532915999,
    0.8951126800219693, 0.9176033189485007, 0.9116016387064556, 0.8890492343084937, 0.8803727480781831, 0.8898403139631976,
    0.9348273906591118, 0.9387867073874407, 0.9213150498959719, 0.8888279850470155, 0.8918981127741714, 0.8604755993712578,
    0.8382051613193982, 0.8515837826695183, 0.858730482918872, 0.8415828215289731, 0.8190334105144152, 0.8317518172797949,
    0.8079382135754152, 0.7900809748209163, 0.7800701153643972, 0.7698151649955114, 0.7549149228315972, 0.7393688207123867,
    0.7085441619225441, 0.7010892852241013, 0.6740704877921768, 0.6613179904098958, 0.6424878029890573, 0.6351387703085332,
    0.6218820987970053, 0.5928624580840685, 0.5646145379419862, 0.5097661577740067, 0.5432830324275337, 0.5100248798209441,
    0.47253575285403415, 0.3571200478171759, 0.31034679770879103, 0.23330181449341403, 0.07433564353203437, 0,
    0.27666333287124445, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

static const char *USAGE =
    "Simulate RNA-seq library preparation with priming biases, PCR biases and size selection (version: %s).\n\n"
    "Usage:\n        rlsim_gpu [arguments] [transcriptome fasta files (optional)]\n\n"
    "Optional arguments:\n"
    "        -n      requested fragments         int\n"
    "        -d      fragment size distribution  string  \"1.0:sn:(189, 24, -1.09975, 76, 294)\"\n"
    "        -f      fragmentation method        string  \"after_prim_double\"\n"
    "        -b      strand bias                 float   0.5\n"
    "        -c      PCR cycles                  int     11\n"
    "        -p      priming bias parameter      float   5.0\n"
    "        -k      primer length               int     6\n"
    "        -a      poly(A) tail size dist.     string  \"1.0:g:(150,1,0,220)\"\n"
    "        -flg    fragment loss probability   float   0.0\n"
    "        -m      expression level multiplier float   1.0\n"
    "        -e      fixed PCR efficiency        float   0.0\n"
    "        -eg     GC efficiency parameters as \"(shape, min, max)\": raw from SRR521457\n"
    "        -el     length efficiency parameters as \"(shape, min, max)\": (0.0, 1.0, 1.0)\n"
    "        -j      raw parameter file          string  (superseeds -d, -c, -eg)\n"
    "        -jm     minimum raw gc efficiency   float   0.0\n"
    "        -r      report file                 string  \"rlsim_report.json\"\n"
    "        -t      number of cores to use      int     4\n"
    "        -g      keep fragments in memory    bool    false\n"
    "        -si / -sp / -ss   initial / pcr / sampling random seed\n"
    "        -gobdir -gcfreq -prof -randt       accepted for compatibility\n"
    "        -v      toggle verbose mode         bool    false\n"
    "        -h      print usage and exit        -V      print version and exit\n"
    "        -rng    random stream               string  \"philox\" (GPU path)\n";

std::string MixComp::str() const {
    char buf[256];
    if (type == "sn") snprintf(buf, sizeof buf, "%s:(%g, %g, %g, %llu, %llu)", type.c_str(), location, scale, shape, (unsigned long long)low, (unsigned long long)high);
    else snprintf(buf, sizeof buf, "%s:(%g, %g, %llu, %llu)", type.c_str(), location, scale, (unsigned long long)low, (unsigned long long)high);
    return buf;
}

static std::string trim(const std::string &s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
static std::vector<std::string> split(const std::string &s, char c) {
    std::vector<std::string> out;
    std::string cur;
    for (char ch : s) { if (ch == c) { out.push_back(cur); cur.clear(); } else cur += ch; }
    out.push_back(cur);
    return out;
}
static bool scan_double(const std::string &s, double *v) {
    std::string t = trim(s);
    if (t.empty()) return false;
    char *end = nullptr;
    *v = strtod(t.c_str(), &end);
    return end && *end == 0;
}
static bool scan_u64(const std::string &s, uint64_t *v) {
    std::string t = trim(s);
    if (t.empty() || t[0] == '-') return false;
    char *end = nullptr;
    *v = strtoull(t.c_str(), &end, 10);
    return end && *end == 0;
}

static MixComp parse_component(const std::string &s) {  // target_mix.go:128-214
    auto spl1 = split(s, ':');
    if (spl1.size() != 3) fatal("Something is missing from mixture component: " + s);
    MixComp c;
    if (!scan_double(spl1[0], &c.weight)) fatal("Invalid mixture weight: " + spl1[0]);
    c.type = trim(spl1[1]);
    if (c.type != "n" && c.type != "sn" && c.type != "g") fatal("Invalid mixture type: " + c.type);
    std::string tmp = spl1[2];
    if (tmp.empty() || tmp.front() != '(' || tmp.back() != ')') fatal("Missing paranthesis in mixture component string: " + tmp);
    while (!tmp.empty() && (tmp.front() == '(' || tmp.front() == ')')) tmp.erase(tmp.begin());
    while (!tmp.empty() && (tmp.back() == '(' || tmp.back() == ')')) tmp.pop_back();
    auto spl2 = split(tmp, ',');
    if (c.type == "sn") { if (spl2.size() != 5) fatal("Not enough parameters in skew normal mixture string: \"" + tmp + "\""); }
    else if (spl2.size() != 4) fatal("Not enough parameters in mixture string: \"" + tmp + "\"");
    size_t i = 0;
    if (!scan_double(spl2[i], &c.location)) fatal("Invalid location \"" + spl2[i] + "\" in mixture: \"" + tmp + "\"");
    i++;
    if (!scan_double(spl2[i], &c.scale)) fatal("Invalid scale \"" + spl2[i] + "\" in mixture: \"" + tmp + "\"");
    i++;
    if (c.type == "sn") { if (!scan_double(spl2[i], &c.shape)) fatal("Invalid shape \"" + spl2[i] + "\" in mixture: \"" + tmp + "\""); i++; }
    if (!scan_u64(spl2[i], &c.low)) fatal("Invalid lower bound \"" + spl2[i] + "\" in mixture: \"" + tmp + "\"");
    i++;
    if (!scan_u64(spl2[i], &c.high)) fatal("Invalid upper bound \"" + spl2[i] + "\" in mixture: \"" + tmp + "\"");
    if (c.location < 0) fatal("Location parameter is negative in mixture string: " + s);
    if (c.scale < 0) fatal("Scale parameter is negative in mixture string: " + s);
    return c;
}

std::vector<MixComp> parse_mixture(const std::string &s) {  // target_mix.go:90-113
    std::vector<MixComp> out;
    for (auto &part : split(s, '+')) {
        std::string p = trim(part);
        if (p.empty()) fatal("Empty target mixture string!");
        MixComp c = parse_component(p);
        if (c.high < c.low) fatal("Maximum fragment size is smaller than minimum fragment size for component \"" + p + "\"!");
        out.push_back(c);
    }
    if (out.size() > RLSIM_MAX_COMPS) fatal("At most 8 mixture components are supported");
    return out;
}

std::string mixture_string(const std::vector<MixComp> &m) {  // target_mix.go:78-88
    std::string s;
    for (size_t i = 0; i < m.size(); i++) {
        char buf[64];
        snprintf(buf, sizeof buf, " \t%.2f:", m[i].weight);
        s += buf + m[i].str();
        if (i + 1 != m.size()) s += "\n";
    }
    return s;
}

static void parse_eff(const std::string &str, double *shape, double *mn, double *mx) {  // args.go:310-333
    std::string s = trim(str);
    if (s.empty() || s.front() != '(' || s.back() != ')') fatal("Missing paranthesis in efficiency parameter string: " + s);
    std::string body = s.substr(1, s.size() - 2);
    auto v = split(body, ',');
    if (v.size() != 3 || !scan_double(v[0], shape) || !scan_double(v[1], mn) || !scan_double(v[2], mx))
        fatal("Failed to parse efficieny parameter string: " + body);
    char buf[128];
    if (*shape < 0) { snprintf(buf, sizeof buf, "Efficiency shape parameter %g is negative!", *shape); fatal(buf); }
    if (*mn < 0 || *mn > 1.0) { snprintf(buf, sizeof buf, "The minimum efficiency %g is outside the range [0, 1]!", *mn); fatal(buf); }
    if (*mx < 0 || *mx > 1.0) { snprintf(buf, sizeof buf, "The maximum efficiency %g is outside the range [0, 1]!", *mx); fatal(buf); }
}

// ---- minimal JSON reader for raw_params.json (parse_raw.go:51-126): {"nr_frags":x,"nr_cycles":x,"frag_dist":{"l":c,...},"gc_eff":{"i":e,...}}
struct JsonCursor {
    const std::string &s;
    size_t i = 0;
    void ws() { while (i < s.size() && isspace((unsigned char)s[i])) i++; }
    bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { i++; return true; } return false; }
    std::string str() { ws(); std::string out; if (!eat('"')) return out; while (i < s.size() && s[i] != '"') out += s[i++]; i++; return out; }
    double num() { ws(); char *end; double v = strtod(s.c_str() + i, &end); i = end - s.c_str(); return v; }
};
static RawParams decode_raw_params(const std::string &fname) {
    std::ifstream f(fname);
    if (!f) fatal("Error when reading raw parameter file " + fname);
    std::stringstream ss; ss << f.rdbuf();
    std::string txt = ss.str();
    JsonCursor c{txt};
    RawParams rp;
    bool have[4] = {false, false, false, false};
    std::vector<std::pair<int, double>> gc;
    if (!c.eat('{')) fatal("Malformed raw parameter file " + fname);
    while (true) {
        std::string key = c.str();
        if (!c.eat(':')) break;
        if (key == "nr_frags") { rp.req_frags = (int64_t)c.num(); have[0] = true; }
        else if (key == "nr_cycles") { rp.nr_cycles = (int64_t)c.num(); have[1] = true; }
        else if (key == "frag_dist" || key == "gc_eff") {
            if (!c.eat('{')) fatal("Malformed raw parameter file " + fname);
            while (true) {
                std::string k = c.str();
                if (!c.eat(':')) break;
                double v = c.num();
                if (key == "frag_dist") { rp.lengths.push_back((uint32_t)atol(k.c_str())); rp.probs.push_back(v); }
                else gc.push_back({atoi(k.c_str()), v});
                if (!c.eat(',')) break;
            }
            c.eat('}');
            have[key == "frag_dist" ? 2 : 3] = true;
        } else c.num();
        if (!c.eat(',')) break;
    }
    const char *names[4] = {"nr_frags", "nr_cycles", "frag_dist", "gc_eff"};
    for (int k = 0; k < 4; k++) if (!have[k]) fatal(std::string("The ") + names[k] + " key is missing from the raw parameter file!");
    double total = 0;
    for (double v : rp.probs) total += v;
    for (double &v : rp.probs) v /= total;
    rp.gc_effs.assign(gc.size(), 0.0);
    for (auto &kv : gc) if (kv.first >= 0 && (size_t)kv.first < gc.size()) rp.gc_effs[kv.first] = kv.second;
    return rp;
}

void CmdArgs::parse(int argc, char **argv) {  // args.go:83-270; Go `flag` syntax: -x v | -x=v | --x, stops at first non-flag
    std::string targ_mix = "1.0:sn:(189, 24, -1.09975, 76, 294)", polya_s = "1.0:g:(150,1,0,220)", gc_eff_s, len_eff_s = "(0.0,1.0,1.0)";
    bool help = false, version = false, gob = false, randt = false;
    struct F { const char *name; int kind; void *dst; };  // kind: 0 int64, 1 double, 2 string, 3 bool, 4 int
    F flags[] = {{"n", 0, &req_frags}, {"d", 2, &targ_mix}, {"c", 0, &nr_cycles}, {"a", 2, &polya_s}, {"b", 1, &strand_bias},
                 {"p", 1, &priming_temp}, {"k", 4, &kmer_length}, {"e", 1, &fixed_eff}, {"eg", 2, &gc_eff_s}, {"el", 2, &len_eff_s},
                 {"j", 2, &raw_params_file}, {"jm", 1, &min_raw_gc_eff}, {"r", 2, &report_file}, {"t", 0, &max_procs},
                 {"f", 2, &frag_method}, {"flg", 1, &frag_loss_prob}, {"m", 1, &expr_mul}, {"g", 3, &gob}, {"si", 0, &init_seed},
                 {"sp", 0, &pcr_seed}, {"ss", 0, &sampling_seed}, {"gobdir", 2, &gob_dir}, {"gcfreq", 4, &gc_freq},
                 {"prof", 2, &prof_file}, {"h", 3, &help}, {"V", 3, &version}, {"v", 3, &verbose}, {"randt", 3, &randt}, {"rng", 2, &rng}};
    int i = 1;
    for (; i < argc; i++) {
        std::string tok = argv[i];
        if (tok.size() < 2 || tok[0] != '-') break;
        if (tok == "--") { i++; break; }
        std::string name = tok.substr(tok[1] == '-' ? 2 : 1), value;
        bool has_value = false;
        size_t eq = name.find('=');
        if (eq != std::string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); has_value = true; }
        F *fl = nullptr;
        for (auto &f : flags) if (name == f.name) fl = &f;
        if (!fl) fatal("flag provided but not defined: -" + name);
        if (fl->kind == 3) { *(bool *)fl->dst = !has_value || value == "1" || value == "t" || value == "true"; continue; }
        if (!has_value) { if (i + 1 >= argc) fatal("flag needs an argument: -" + name); value = argv[++i]; }
        char *end = nullptr;
        bool ok = true;
        switch (fl->kind) {
        case 0: *(int64_t *)fl->dst = strtoll(value.c_str(), &end, 10); ok = end && *end == 0 && !value.empty(); break;
        case 4: *(int *)fl->dst = (int)strtol(value.c_str(), &end, 10); ok = end && *end == 0 && !value.empty(); break;
        case 1: *(double *)fl->dst = strtod(value.c_str(), &end); ok = end && *end == 0 && !value.empty(); break;
        default: *(std::string *)fl->dst = value;
        }
        if (!ok) fatal("invalid value \"" + value + "\" for flag -" + name);
    }
    for (; i < argc; i++) input_files.push_back(argv[i]);
    if (help) { fprintf(stderr, USAGE, VERSION); exit(0); }
    if (version) { printf("%s\n", VERSION); exit(0); }
    bool default_raw_gc = false;
    if (!gc_eff_s.empty()) { parse_eff(gc_eff_s, &gc_shape, &gc_min, &gc_max); have_gc_param = true; }
    else { raw_gc_effs.assign(EFF_SRR521457, EFF_SRR521457 + 101); default_raw_gc = true; }
    parse_eff(len_eff_s, &len_shape, &len_min, &len_max);
    if (kmer_length < 1) fatal("Invalid primer length!");
    if (frag_loss_prob < 0.0 || frag_loss_prob > 1) fatal("The fragment loss probability must be in the interval [0, 1]!");
    polya = parse_mixture(polya_s);
    if (req_frags < 0) fatal("Illegal fragment number!");
    target_mix = parse_mixture(targ_mix);
    {  // args.go:277-307
        auto spl = split(frag_method, ':');
        static const char *names[] = {"after_prim", "after_prim_double", "after_noprim", "after_noprim_double", "pre_prim", "prim_jump"};
        bool known = false;
        for (auto n : names) known |= spl[0] == n;
        if (spl.size() > 2 || !known) fatal("Malformed fragmentation method string: " + frag_method);
        if (spl.size() == 2) { char *end; frag_param = (int)strtol(spl[1].c_str(), &end, 10); if (!end || *end || spl[1].empty()) fatal("Malformed fragmentation method string: " + frag_method); }
        frag_method = spl[0];
        if (frag_method == "pre_prim" && frag_param == 0) frag_param = 2000;
    }
    if (!gob && gob_dir.empty()) gob_dir = "rlsim_gob_" + std::to_string((long)getpid());
    if (!raw_params_file.empty()) {  // args.go:247-255
        raw = decode_raw_params(raw_params_file);
        have_raw = true;
        if (req_frags == 0) req_frags = raw.req_frags;
        nr_cycles = raw.nr_cycles;
        raw_gc_effs = raw.gc_effs;
    }
    if (rng != "philox") fatal("Unsupported -rng \"" + rng + "\": the GPU path implements \"philox\" only");
    if (default_raw_gc && req_frags > 0 && raw_params_file.empty()) {
        warnings.push_back("WARNING: Using raw efficiencies estimated from SRR521457 by default");
        warnings.push_back("(see http://bit.ly/rlsim-pl and http://bit.ly/rlsim-pa for the details).");
    }
    if (req_frags == 0) { fprintf(stderr, "No fragments requested! Exiting.\n"); fprintf(stderr, USAGE, VERSION); exit(0); }
}

rlsim_params CmdArgs::to_params() const {
    rlsim_params p;
    memset(&p, 0, sizeof p);
    static const char *names[] = {"after_prim", "after_prim_double", "after_noprim", "after_noprim_double", "pre_prim", "prim_jump"};
    for (int i = 0; i < 6; i++) if (frag_method == names[i]) p.frag_method = i;
    p.kmer_len = kmer_length; p.frag_param = frag_param; p.nr_cycles = (int32_t)nr_cycles;
    p.priming_temp = priming_temp; p.frag_loss_prob = frag_loss_prob; p.fixed_eff = fixed_eff; p.strand_bias = strand_bias;
    if (!raw_gc_effs.empty()) {  // thermocycler.go:62-66,94-104
        if (raw_gc_effs.size() != 101) fatal("Raw GC efficiency table must have 101 entries");
        p.gc_mode = 1;
        for (int i = 0; i < 101; i++) p.gc_raw[i] = std::max(std::min(raw_gc_effs[i], 1.0), min_raw_gc_eff);
    } else { p.gc_mode = 0; p.gc_shape = gc_shape; p.gc_min = gc_min; p.gc_max = gc_max; }
    p.len_shape = len_shape; p.len_min = len_min; p.len_max = len_max;
    auto fill = [](rlsim_mixcomp *dst, const std::vector<MixComp> &m) {
        for (size_t i = 0; i < m.size(); i++) {
            dst[i].type = m[i].type == "n" ? RLSIM_COMP_N : m[i].type == "sn" ? RLSIM_COMP_SN : RLSIM_COMP_G;
            dst[i].weight = m[i].weight; dst[i].location = m[i].location; dst[i].scale = m[i].scale; dst[i].shape = m[i].shape;
            dst[i].low = m[i].low; dst[i].high = m[i].high;
        }
    };
    p.n_target = have_raw ? 0 : (int32_t)target_mix.size();
    if (!have_raw) fill(p.target, target_mix);
    p.n_polya = (int32_t)polya.size();
    fill(p.polya, polya);
    p.seed_init = init_seed; p.seed_pcr = pcr_seed; p.seed_sampling = sampling_seed;
    return p;
}

double go_pow(double x, double y) {  // structure of Go's math.Pow
    if (y == 0 || x == 1) return 1.0;
    if (y == 1) return x;
    if (y == 0.5) return std::sqrt(x);
    if (y == -0.5) return 1.0 / std::sqrt(x);
    if (std::isnan(x) || std::isnan(y)) return NAN;
    if (x == 0) return y < 0 ? INFINITY : 0.0;
    double absy = std::fabs(y);
    bool flip = y < 0;
    double yi = std::floor(absy), yf = absy - yi;
    if (yf != 0 && x < 0) return NAN;
    if (yi >= 9.223372036854775808e18) return std::exp(y * std::log(x));
    double a1 = 1.0;
    int ae = 0;
    if (yf != 0) { if (yf > 0.5) { yf -= 1; yi += 1; } a1 = std::exp(yf * std::log(x)); }
    int xe;
    double x1 = std::frexp(x, &xe);
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1; xe <<= 1;
        if (x1 < 0.5) { x1 += x1; xe--; }
    }
    if (flip) { a1 = 1 / a1; ae = -ae; }
    return std::ldexp(a1, ae);
}

}  // namespace host
