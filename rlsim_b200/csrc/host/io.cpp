// FASTA input (fasta.go:113-212, input.go:82-124), JSON report (report.go:46-148), logging (logging.go:49-122).
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sstream>
#include "host.hpp"

namespace host {

bool g_verbose = false;

void log_line(const std::string &msg) {  // log flag 4 = microseconds (logging.go:61)
    using namespace std::chrono;
    auto now = system_clock::now();
    time_t tt = system_clock::to_time_t(now);
    struct tm lt;
    localtime_r(&tt, &lt);
    long us = (long)(duration_cast<microseconds>(now.time_since_epoch()).count() % 1000000);
    fprintf(stderr, "%02d:%02d:%02d.%06ld %s\n", lt.tm_hour, lt.tm_min, lt.tm_sec, us, msg.c_str());
}
void fatal(const std::string &msg) { log_line(msg); exit(1); }

static bool parse_seq_name(const std::string &s, std::string *name, uint64_t *level) {  // input.go:82-101
    size_t d = s.find('$');
    if (d == std::string::npos || d == 0 || s.find('$', d + 1) != std::string::npos) return false;
    *name = s.substr(0, d);
    size_t i = d + 1;
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\r' || s[i] == '\n' || s[i] == '\v' || s[i] == '\f')) i++;
    if (i >= s.size() || !isdigit((unsigned char)s[i])) return false;
    char *end;
    errno = 0;
    *level = strtoull(s.c_str() + i, &end, 10);
    return errno != ERANGE;  // strconv.ParseUint: value out of range
}

std::string read_input(const std::vector<std::string> &files) {
    std::string data;
    if (files.empty()) { std::stringstream ss; ss << std::cin.rdbuf(); data = ss.str(); }
    for (auto &f : files) {  // io.MultiReader: plain concatenation
        std::ifstream in(f, std::ios::binary);
        if (!in) fatal("Could not open input file " + f);
        std::stringstream ss; ss << in.rdbuf();
        data += ss.str();
    }
    return data;
}

// Device-side reader: rlsim_add_fasta splits, despaces, upper-cases and parses the records on the GPU and adds them.
// Returns false when the text needs the host reader (RLSIM_E_UNSUPPORTED); other errors are fatal as in the reference.
bool load_transcripts_device(rlsim_handle *h, const std::string &data, double expr_mul, Transcripts *out) {
    uint32_t n_rec = 0, n_add = 0;
    int rc = rlsim_add_fasta(h, (const uint8_t *)data.data(), data.size(), expr_mul, &n_rec, &n_add);
    if (rc == RLSIM_E_UNSUPPORTED) return false;
    if (rc) fatal(rlsim_last_error(h));
    std::vector<uint64_t> start(n_rec), level(n_rec), seq_len(n_rec);
    std::vector<uint32_t> line_len(n_rec), name_len(n_rec);
    std::vector<uint8_t> valid(n_rec);
    rlsim_get_fasta_index(h, start.data(), line_len.data(), name_len.data(), level.data(), seq_len.data(), valid.data());
    Transcripts tr;
    tr.offsets.push_back(0);
    for (uint32_t r = 0; r < n_rec; r++) {
        if (!valid[r]) { log_line("Skipping transcript with malformed name: " + data.substr(start[r], line_len[r])); continue; }
        tr.names.push_back(data.substr(start[r], name_len[r]));
        tr.levels.push_back(level[r]);
        tr.offsets.push_back(tr.offsets.back() + seq_len[r]);
        tr.tr_lengths[(uint32_t)seq_len[r]] += level[r];   // input.go:115
        tr.expr_levels[(uint32_t)level[r]] += 1;           // input.go:117
    }
    *out = std::move(tr);
    return true;
}

Transcripts load_transcripts(const std::string &data, double expr_mul) {
    Transcripts tr;
    tr.offsets.push_back(0);
    size_t n = data.size(), rec = 0;
    while (rec < n) {
        // fasta.go:138: any '>' opens a new record once the buffer holds more than one byte
        size_t next = rec + 1;
        while (next < n && !(data[next] == '>' && next - rec > 1)) next++;
        if (data[rec] != '>') fatal("Illegal fasta record:\n" + data.substr(rec, std::min<size_t>(80, next - rec)));
        size_t nl = data.find('\n', rec);
        if (nl == std::string::npos || nl > next) nl = next;
        std::string header = data.substr(rec + 1, nl - rec - 1), name;
        uint64_t level = 0;
        if (!parse_seq_name(header, &name, &level)) {
            log_line("Skipping transcript with malformed name: " + header);
        } else {
            level = (uint64_t)((double)level * expr_mul);  // input.go:113
            size_t before = tr.bases.size();
            for (size_t i = std::min(nl + 1, next); i < next; i++) {
                unsigned char c = (unsigned char)data[i];
                if (c == ' ' || c == '\t' || c == '\r' || c == '\n') continue;
                tr.bases.push_back((uint8_t)toupper(c));
            }
            uint32_t len = (uint32_t)(tr.bases.size() - before);
            tr.names.push_back(name);
            tr.offsets.push_back(tr.bases.size());
            tr.levels.push_back(level);
            tr.tr_lengths[len] += level;            // input.go:115
            tr.expr_levels[(uint32_t)level] += 1;   // input.go:117
        }
        rec = next;
    }
    return tr;
}

// ---- report.go
static std::string num(double v) {  // shortest round-trip representation
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    return std::string(buf, r.ptr);
}
static std::string gfmt(double v) { char buf[64]; snprintf(buf, sizeof buf, "%g", v); return buf; }

void Report::add_u64(const std::map<uint32_t, uint64_t> &m, const char *xl, const char *yl, const char *title, const char *vis) {
    Panel p{xl, yl, vis, cursor_++, {}};
    for (auto &kv : m) p.data.push_back({std::to_string(kv.first), std::to_string(kv.second)});
    pool_[title] = p;
}
void Report::add_f64(const std::map<uint32_t, double> &m, const char *xl, const char *yl, const char *title, const char *vis) {
    Panel p{xl, yl, vis, cursor_++, {}};
    for (auto &kv : m) p.data.push_back({std::to_string(kv.first), num(kv.second)});
    pool_[title] = p;
}
void Report::add_xy(const std::vector<double> &x, const std::vector<double> &y, const char *xl, const char *yl, const char *title, const char *vis) {
    Panel p{xl, yl, vis, cursor_++, {}};
    for (size_t i = 0; i < x.size(); i++) p.data.push_back({gfmt(x[i]), std::isfinite(y[i]) ? num(y[i]) : std::string("null")});
    pool_[title] = p;
}
void Report::write_json() {  // report.go:140-148; Go marshals map keys in sorted order
    std::ofstream f(path_);
    if (!f) fatal("Could not create report file " + path_);
    f << "{";
    bool first = true;
    for (auto &kv : pool_) {
        if (!first) f << ",";
        first = false;
        auto data = kv.second.data;
        std::sort(data.begin(), data.end());
        f << "\"" << kv.first << "\":{\"Xl\":\"" << kv.second.xl << "\",\"Yl\":\"" << kv.second.yl << "\",\"Vis\":\"" << kv.second.vis
          << "\",\"Pos\":" << kv.second.pos << ",\"Data\":{";
        for (size_t i = 0; i < data.size(); i++) f << (i ? "," : "") << "\"" << data[i].first << "\":" << data[i].second;
        f << "}}";
    }
    f << "}";
}

}  // namespace host
