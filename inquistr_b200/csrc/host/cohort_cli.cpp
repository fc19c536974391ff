// cohort_cli.cpp -- `inquistr-b200 combine` and `inquistr-b200 outlier`.
// Host side of the reference's src/combine.rs:27-59 and src/outlier.rs:33-72 / src/main.rs:75-99,202-229
// above the C ABI of include/inqcohort.h: text in, text out; the per-row arithmetic runs on the GPU.
// A Rust panic in the reference maps to exit code 101, clap usage errors to 2.
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../../include/inqcall.h"
#include "../../../include/inqcohort.h"

namespace inqhost {

namespace {

[[noreturn]] void cpanic(const std::string &msg)
{
    fprintf(stderr, "thread 'main' panicked: %s\n", msg.c_str());
    exit(101);
}
[[noreturn]] void cusage(const std::string &msg, const char *sub)
{
    fprintf(stderr, "error: %s\n\nUsage: inquistr-b200 %s\n\nFor more information, try '--help'.\n", msg.c_str(), sub);
    exit(2);
}
bool exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

// plain or gzip-compressed text, line by line (combine.rs:10-25 decides by the .gz extension,
// utils.rs:7-13 sniffs the content; gzread is transparent for both)
struct LineReader {
    gzFile f = nullptr;
    std::string buf;
    bool open(const std::string &path) { f = gzopen(path.c_str(), "rb"); if (f) gzbuffer(f, 1 << 20); return f != nullptr; }
    bool next(std::string &line)
    {
        line.clear();
        char tmp[1 << 16];
        bool any = false;
        while (gzgets(f, tmp, sizeof(tmp))) {
            any = true;
            const size_t n = strlen(tmp);
            if (n && tmp[n - 1] == '\n') {
                line.append(tmp, n - 1);
                if (!line.empty() && line.back() == '\r') line.pop_back();     // BufRead::lines strips \r\n too
                return true;
            }
            line.append(tmp, n);
        }
        return any;
    }
    ~LineReader() { if (f) gzclose(f); }
};

void split_tabs(const std::string &s, std::vector<std::string> &out)
{
    out.clear();
    size_t a = 0;
    for (;;) {
        const size_t b = s.find('\t', a);
        out.emplace_back(s, a, b == std::string::npos ? std::string::npos : b - a);
        if (b == std::string::npos) break;
        a = b + 1;
    }
}

// Rust's `str::parse::<f32>`: decimal digits / '.' / exponent with optional sign, or (case-insensitive)
// "nan", "inf", "infinity"; no surrounding whitespace
bool parse_f32(const std::string &s, float *out)
{
    if (s.empty()) return false;
    size_t i = (s[0] == '+' || s[0] == '-') ? 1 : 0;
    std::string low;
    for (size_t k = i; k < s.size(); ++k) low.push_back((char)tolower((unsigned char)s[k]));
    if (low == "nan") { *out = NAN; return true; }
    if (low == "inf" || low == "infinity") { *out = s[0] == '-' ? -INFINITY : INFINITY; return true; }
    bool digit = false;
    for (size_t k = i; k < s.size(); ++k) {
        const char c = s[k];
        if (c >= '0' && c <= '9') digit = true;
        else if (!(c == '.' || c == 'e' || c == 'E' || ((c == '+' || c == '-') && k > i && (s[k - 1] == 'e' || s[k - 1] == 'E')))) return false;
    }
    if (!digit) return false;
    char *end = nullptr;
    *out = strtof(s.c_str(), &end);
    return end && *end == '\0';
}

std::string strip_hap(std::string s)                           // outlier.rs:112 .replace("_H1", "").replace("_H2", "")
{
    for (const char *pat : {"_H1", "_H2"})
        for (size_t p; (p = s.find(pat)) != std::string::npos;) s.erase(p, 3);
    return s;
}

}  // namespace

// combine.rs:27-59
int combine_main(const std::vector<std::string> &files)
{
    if (files.empty()) cusage("the following required arguments were not provided:\n  <CALLS>...", "combine <CALLS>...");
    for (const auto &f : files)
        if (!exists(f)) cpanic("File " + f + " does not exist!");
    std::vector<LineReader> rd(files.size());
    for (size_t i = 0; i < files.size(); ++i)
        if (!rd[i].open(files[i])) cpanic("couldn't open " + files[i]);
    std::string line, other, out;
    std::vector<std::string> cols;
    while (rd[0].next(line)) {
        out = line;
        for (size_t i = 1; i < files.size(); ++i) {
            if (!rd[i].next(other)) cpanic("called `Option::unwrap()` on a `None` value (" + files[i] + " has fewer lines than " + files[0] + ")");
            split_tabs(other, cols);
            for (size_t c = 3; c < cols.size(); ++c) { out.push_back('\t'); out += cols[c]; }   // only the scores (combine.rs:50-54)
        }
        out.push_back('\n');
        fwrite(out.data(), 1, out.size(), stdout);
    }
    return 0;
}

// main.rs:75-99,202-229 + outlier.rs:33-72
int outlier_main(const std::vector<std::string> &v, int device)
{
    const char *usage = "outlier [OPTIONS] <COMBINED>";
    std::string combined, sample, subset_file, method = "zscore";
    bool has_sample = false, has_subset = false;
    uint64_t minsize = 10;
    float zscore = 3.0f;
    for (size_t i = 0; i < v.size(); ++i) {
        const std::string &a = v[i];
        auto value = [&](const char *name) -> std::string {
            if (i + 1 >= v.size()) cusage(std::string("a value is required for '") + name + "' but none was supplied", usage);
            return v[++i];
        };
        if (a == "--minsize") {
            const std::string s = value("--minsize <MINSIZE>");
            char *e = nullptr;
            minsize = strtoull(s.c_str(), &e, 10);
            if (s.empty() || *e || s[0] == '-' || minsize > 0xFFFFFFFFull) cusage("invalid value '" + s + "' for '--minsize <MINSIZE>'", usage);
        } else if (a == "-z" || a == "--zscore") {
            const std::string s = value("--zscore <ZSCORE>");
            if (!parse_f32(s, &zscore)) cusage("invalid value '" + s + "' for '--zscore <ZSCORE>'", usage);
        } else if (a == "--method") {
            method = value("--method <METHOD>");
            if (method != "zscore" && method != "dbscan") cusage("invalid value '" + method + "' for '--method <METHOD>'\n  [possible values: zscore, dbscan]", usage);
        } else if (a == "-s" || a == "--sample") { sample = value("--sample <SAMPLE>"); has_sample = true; }
        else if (a == "-S" || a == "--subset") { subset_file = value("--subset <SUBSET>"); has_subset = true; }
        else if (a == "-h" || a == "--help") {
            printf("Find outliers from TSV\n\nUsage: inquistr-b200 %s\n\nArguments:\n  <COMBINED>  combined file of calls\n\nOptions:\n"
                   "      --minsize <MINSIZE>  minimal length of expansion to be present in cohort [default: 10]\n"
                   "  -z, --zscore <ZSCORE>    zscore cutoff to decide if a value is an outlier [default: 3]\n"
                   "      --method <METHOD>    method to test for outliers [default: zscore] [possible values: zscore, dbscan]\n"
                   "  -s, --sample <SAMPLE>    sample to consider\n  -S, --subset <SUBSET>    file with subset of samples to consider\n"
                   "  -h, --help               Print help\n", usage);
            return 0;
        } else if (!a.empty() && a[0] == '-' && a.size() > 1) cusage("unexpected argument '" + a + "' found", usage);
        else if (combined.empty()) combined = a;
        else cusage("unexpected argument '" + a + "' found", usage);
    }
    if (combined.empty()) cusage("the following required arguments were not provided:\n  <COMBINED>", usage);
    if (!exists(combined)) cpanic("Combined file does not exist!");                       // main.rs:210-212
    if (has_sample && has_subset) cpanic("Cannot use both -s and -S arguments");          // main.rs:214-216
    std::unordered_set<std::string> subset;
    const bool use_subset = has_sample || has_subset;
    if (has_sample) subset.insert(sample);
    if (has_subset) {
        LineReader sr;
        if (!sr.open(subset_file)) cpanic("Problem opening file");                        // utils.rs:9
        std::string l;
        while (sr.next(l)) subset.insert(l);
    }

    LineReader rd;
    if (!rd.open(combined)) cpanic("Problem opening file");
    std::string line;
    if (!rd.next(line)) cpanic("called `Option::unwrap()` on a `None` value");            // outlier.rs:36
    fputs("chrom\tbegin\tend\toutliers\n", stdout);                                       // outlier.rs:37
    std::vector<std::string> cols;
    split_tabs(line, cols);
    std::vector<std::string> samples(cols.begin() + std::min<size_t>(3, cols.size()), cols.end());
    if (samples.empty()) cpanic("argument of integer logarithm must be positive");       // ilog2(0), outlier.rs:39
    const uint32_t n_cols = (uint32_t)samples.size();
    std::vector<std::string> names(n_cols);
    for (uint32_t c = 0; c < n_cols; ++c) names[c] = strip_hap(samples[c]);
    const int meth = method == "dbscan" ? INQ_OUTLIER_DBSCAN : INQ_OUTLIER_ZSCORE;

    // rows are processed in batches of the combined matrix
    const size_t batch_rows = std::max<size_t>(1, std::min<size_t>(1u << 16, (256u << 20) / ((size_t)n_cols * sizeof(float))));
    std::vector<float> values;
    std::vector<std::string> keys;                       // "chrom\tbegin\tend"
    std::vector<uint64_t> hits;
    auto flush = [&]() {
        uint64_t rows = keys.size();
        if (!rows) return;
        uint64_t n_hits = 0;
        if (hits.size() < 1024) hits.resize(1024);
        int rc;
        std::string no_mode;
        for (;;) {
            while ((rc = inq_outlier(device, meth, rows, n_cols, values.data(), (uint32_t)minsize, zscore, nullptr, &n_hits,
                                     hits.data(), hits.size(), nullptr)) == INQ_ERR_HITS_CAP)
                hits.resize(n_hits);
            if (rc != INQ_ERR_NO_MODE || !no_mode.empty()) break;
            // The reference streams row by row: the rows before the one without a mode have been printed when it panics
            // (outlier.rs:144). The message names a row without a mode; the batch is cut at the FIRST one.
            no_mode = inq_cohort_last_error();
            const size_t at = no_mode.find("row ");
            uint64_t bad = at == std::string::npos ? 0 : strtoull(no_mode.c_str() + at + 4, nullptr, 10);
            // rows are independent: find the first bad row by re-running ever shorter prefixes (rare path)
            while (bad > 0) {
                uint64_t nh = 0;
                const int r2 = inq_outlier(device, meth, bad, n_cols, values.data(), (uint32_t)minsize, zscore, nullptr, &nh, hits.data(), 0, nullptr);
                if (r2 != INQ_ERR_NO_MODE) break;
                const std::string m2 = inq_cohort_last_error();
                const size_t a2 = m2.find("row ");
                const uint64_t b2 = a2 == std::string::npos ? 0 : strtoull(m2.c_str() + a2 + 4, nullptr, 10);
                if (b2 >= bad) break;
                bad = b2;
            }
            rows = bad;
            if (!rows) break;
        }
        if (rc == INQ_ERR_NO_MODE) cpanic(std::string("No mode found for repeat (") + no_mode + ")");          // the first row of the batch
        if (rc != INQ_OK) { fprintf(stderr, "inquistr-b200: %s\n", inq_cohort_last_error()); exit(1); }
        if (rc != INQ_OK) { fprintf(stderr, "inquistr-b200: %s\n", inq_cohort_last_error()); exit(1); }
        std::string out;
        for (uint64_t i = 0; i < n_hits;) {
            const uint64_t row = hits[i] >> 32;
            uint64_t j = i;
            bool wanted = !use_subset;
            out = keys[row];
            out.push_back('\t');
            for (; j < n_hits && (hits[j] >> 32) == row; ++j) {
                const std::string &nm = names[hits[j] & 0xFFFFFFFFull];
                if (j > i) out.push_back(',');
                out += nm;
                if (use_subset && subset.count(nm)) wanted = true;                        // outlier.rs:59-64
            }
            out.push_back('\n');
            if (wanted) fwrite(out.data(), 1, out.size(), stdout);
            i = j;
        }
        if (!no_mode.empty()) { fflush(stdout); cpanic(std::string("No mode found for repeat (") + no_mode + ")"); }   // outlier.rs:144
        values.clear();
        keys.clear();
    };
    while (rd.next(line)) {
        split_tabs(line, cols);
        // the reference prints row by row, so everything before a row it panics on is already out: flush first
        auto die = [&](const std::string &msg) { values.resize(keys.size() * (size_t)n_cols); flush(); fflush(stdout); cpanic(msg); };
        if (cols.size() < 3) die("index out of bounds: the len is " + std::to_string(cols.size()) + " but the index is 2");
        if (cols.size() - 3 > n_cols) die("index out of bounds: more values than samples in the header");
        if (cols.size() - 3 < n_cols)
            die("row '" + cols[0] + "\t" + cols[1] + "\t" + cols[2] + "' has fewer values than the header has samples (this build needs a rectangular matrix)");
        for (uint32_t c = 0; c < n_cols; ++c) {
            float f;
            if (!parse_f32(cols[3 + c], &f)) die("Failed to parse number: ParseFloatError { kind: Invalid }");   // outlier.rs:79
            values.push_back(f);
        }
        keys.push_back(cols[0] + "\t" + cols[1] + "\t" + cols[2]);
        if (keys.size() == batch_rows) flush();
    }
    flush();
    return 0;
}

}  // namespace inqhost
