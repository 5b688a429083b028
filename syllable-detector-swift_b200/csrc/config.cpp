// Parser for the `key = value` network description written by convert_to_text.m, with the rules of
// SyllableDetectorConfig.init(fromTextFile:) (Common/SyllableDetectorConfig.swift:170-277):
//   * a line counts only if splitting at '=' (empty pieces dropped) gives exactly two pieces (:183-189), both trimmed
//     of whitespace/newlines (Common/Common.swift:16-24); later duplicates replace earlier ones;
//   * numbers use the whole trimmed string (Swift Double/Float/Int initialisers), Float is decimal->binary32 directly;
//   * lists split at ',' with empty pieces dropped (:81-113);
//   * fields are read in the upstream order so the first error names the same key.
#include "config.hpp"

#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <locale.h>
#include <sstream>
#include <string_view>
#include <unordered_map>

namespace syldet {

namespace {
thread_local std::string g_err_msg;
thread_local std::string g_err_key;

using Dict = std::unordered_map<std::string, std::string>;

bool is_space(char ch) { return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == '\v' || ch == '\f'; }

std::string_view trimmed(std::string_view s) {
    size_t a = 0, b = s.size();
    while (a < b && is_space(s[a])) ++a;
    while (b > a && is_space(s[b - 1])) --b;
    return s.substr(a, b - a);
}

// Swift split(omittingEmptySubsequences: true)
std::vector<std::string_view> split_nonempty(std::string_view s, char sep) {
    std::vector<std::string_view> parts;
    size_t pos = 0;
    while (pos <= s.size()) {
        size_t nxt = s.find(sep, pos);
        if (nxt == std::string_view::npos) nxt = s.size();
        if (nxt > pos) parts.push_back(s.substr(pos, nxt - pos));
        pos = nxt + 1;
    }
    return parts;
}

Dict read_pairs(std::string_view text) {
    Dict d;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t nl = text.find('\n', pos);
        if (nl == std::string_view::npos) nl = text.size();
        auto parts = split_nonempty(text.substr(pos, nl - pos), '=');
        if (parts.size() == 2) d[std::string(trimmed(parts[0]))] = std::string(trimmed(parts[1]));
        pos = nl + 1;
    }
    return d;
}

// Swift's Double(String) / Float(String) do not depend on the process locale: parse under the "C" locale whatever LC_NUMERIC says
// (a host application running with e.g. de_DE would otherwise read "0.5" as 0).
locale_t c_locale() {
    static locale_t loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);
    return loc;
}

template <typename T>
bool whole_number(const std::string &s, T &out);

template <>
bool whole_number<double>(const std::string &s, double &out) {
    if (s.empty() || is_space(s.front())) return false;
    char *end = nullptr;
    out = strtod_l(s.c_str(), &end, c_locale());
    return end != s.c_str() && *end == '\0';
}
template <>
bool whole_number<float>(const std::string &s, float &out) {
    if (s.empty() || is_space(s.front())) return false;
    char *end = nullptr;
    out = strtof_l(s.c_str(), &end, c_locale());
    return end != s.c_str() && *end == '\0';
}
template <>
bool whole_number<int>(const std::string &s, int &out) {
    size_t i = (!s.empty() && (s[0] == '+' || s[0] == '-')) ? 1 : 0;
    if (i >= s.size()) return false;
    for (size_t j = i; j < s.size(); ++j)
        if (!std::isdigit((unsigned char)s[j])) return false;
    errno = 0;
    long v = std::strtol(s.c_str(), nullptr, 10);
    if (errno != 0 || v > INT32_MAX || v < INT32_MIN) return false;
    out = (int)v;
    return true;
}

struct Reader {
    const Dict &d;
    syldet_status st = SYLDET_OK;
    std::string key;

    bool fail(syldet_status s, const std::string &k) {
        if (st == SYLDET_OK) { st = s; key = k; }
        return false;
    }
    const std::string *find(const std::string &k) const {
        auto it = d.find(k);
        return it == d.end() ? nullptr : &it->second;
    }
    bool text(const std::string &k, std::string &out) {
        auto v = find(k);
        if (!v) return fail(SYLDET_ERR_MISSING, k);
        out = *v;
        return true;
    }
    template <typename T>
    bool number(const std::string &k, T &out) {
        auto v = find(k);
        if (!v) return fail(SYLDET_ERR_MISSING, k);
        if (!whole_number<T>(*v, out)) return fail(SYLDET_ERR_INVALID, k);
        return true;
    }
    template <typename T>
    bool list(const std::string &k, long want, std::vector<T> &out) {
        auto v = find(k);
        if (!v) return fail(SYLDET_ERR_MISSING, k);
        auto pieces = split_nonempty(*v, ',');
        std::vector<T> vals;
        vals.reserve(pieces.size());
        bool all_ok = true;
        for (auto p : pieces) {
            T x;
            if (whole_number<T>(std::string(trimmed(p)), x)) vals.push_back(x);
            else all_ok = false;
        }
        if (!all_ok) return fail(SYLDET_ERR_INVALID, k);
        if (want >= 0 && (long)vals.size() != want) return fail(SYLDET_ERR_MISMATCH, k);
        out.swap(vals);
        return true;
    }
};

bool read_processing(Reader &r, const std::string &name, int count, bool is_input, Processing &p) {
    std::string fn;
    if (!r.text(name + ".function", fn)) return false;
    if (fn == "mapminmax" || fn == "mapstd") {
        p.function = fn == "mapminmax" ? SYLDET_PROC_MAPMINMAX : SYLDET_PROC_MAPSTD;
        return r.list<float>(name + ".xOffsets", count, p.x_offsets) && r.list<float>(name + ".gains", count, p.gains) &&
               r.number<float>(name + (fn == "mapminmax" ? ".yMin" : ".yMean"), p.y);
    }
    if (is_input) {
        if (fn == "l2normalize") { p.function = SYLDET_PROC_L2NORMALIZE; return true; }
        if (fn == "normalize") { p.function = SYLDET_PROC_NORMALIZE; return true; }
        if (fn == "normalizestd") { p.function = SYLDET_PROC_NORMALIZESTD; return true; }
    }
    return r.fail(SYLDET_ERR_INVALID, name + ".function");
}

bool read_all(Reader &r, Config &c) {
    if (!r.number("samplingRate", c.sampling_rate)) return false;
    if (!r.number("fourierLength", c.fourier_length)) return false;
    if (c.fourier_length == 0 || (c.fourier_length & (c.fourier_length - 1)) != 0)  // Int.isPowerOfTwo, Common.swift:27-29
        return r.fail(SYLDET_ERR_INVALID, "fourierLength");
    if (r.find("windowLength") == nullptr) c.window_length = c.fourier_length;      // :204-209
    else if (!r.number("windowLength", c.window_length)) return false;
    if (!r.number("windowOverlap", c.window_overlap)) return false;
    std::vector<double> fr;
    if (!r.list<double>("freqRange", 2, fr)) return false;
    c.freq_lo = fr[0];
    c.freq_hi = fr[1];
    if (!r.number("timeRange", c.time_range)) return false;
    {   // `thresholds`, falling back to the legacy key on ANY failure (:223-229)
        Reader attempt{r.d};
        if (!attempt.list<double>("thresholds", -1, c.thresholds) && !r.list<double>("threshold", -1, c.thresholds)) return false;
    }
    std::string sc;
    if (!r.text("scaling", sc)) return false;
    if (sc == "linear") c.scaling = SYLDET_SCALING_LINEAR;
    else if (sc == "log") c.scaling = SYLDET_SCALING_LOG;
    else if (sc == "db") c.scaling = SYLDET_SCALING_DB;
    else return r.fail(SYLDET_ERR_INVALID, "scaling");

    int n_layers = 0;
    if (!r.number("layers", n_layers)) return false;
    if (n_layers < 0) return r.fail(SYLDET_ERR_CONFIG, "layers");
    for (int i = 0; i < n_layers; ++i) {
        const std::string nm = "layer" + std::to_string(i);
        Layer l;
        if (!r.number(nm + ".inputs", l.inputs) || !r.number(nm + ".outputs", l.outputs)) return false;
        long cnt = (long)l.inputs * (long)l.outputs;
        if (cnt < 0) return r.fail(SYLDET_ERR_MISMATCH, nm + ".outputs");
        if (!r.list<float>(nm + ".weights", cnt, l.weights)) return false;
        if (!r.list<float>(nm + ".biases", l.outputs, l.biases)) return false;
        std::string tf;
        if (!r.text(nm + ".transferFunction", tf)) return false;
        if (tf == "TanSig") l.transfer = SYLDET_TF_TANSIG;
        else if (tf == "LogSig") l.transfer = SYLDET_TF_LOGSIG;
        else if (tf == "PureLin") l.transfer = SYLDET_TF_PURELIN;
        else if (tf == "SatLin") l.transfer = SYLDET_TF_SATLIN;
        else return r.fail(SYLDET_ERR_INVALID, nm + ".transferFunction");
        // NeuralNetLayer.init guard (NeuralNet.swift:340-342) is a fatalError raised while the file is being read
        if (l.inputs <= 0 || l.outputs <= 0) return r.fail(SYLDET_ERR_CONFIG, nm + ".transferFunction");
        c.layers.push_back(std::move(l));
    }
    int n_in = 0, n_out = 0;
    if (!r.number("processInputsCount", n_in)) return false;
    if (n_in > 0 && c.layers.empty()) return r.fail(SYLDET_ERR_CONFIG, "layers");
    for (int i = 0; i < n_in; ++i) {
        Processing p;
        if (!read_processing(r, "processInputs" + std::to_string(i), c.layers.front().inputs, true, p)) return false;
        c.input_processing.push_back(std::move(p));
    }
    if (!r.number("processOutputsCount", n_out)) return false;
    if (n_out > 0 && c.layers.empty()) return r.fail(SYLDET_ERR_CONFIG, "layers");
    for (int i = 0; i < n_out; ++i) {
        Processing p;
        if (!read_processing(r, "processOutputs" + std::to_string(i), c.layers.back().outputs, false, p)) return false;
        c.output_processing.push_back(std::move(p));
    }
    // NeuralNet.init (NeuralNet.swift:245-255): fatalError upstream
    if (c.layers.empty()) return r.fail(SYLDET_ERR_CONFIG, "layers");
    for (size_t i = 1; i < c.layers.size(); ++i)
        if (c.layers[i - 1].outputs != c.layers[i].inputs) return r.fail(SYLDET_ERR_CONFIG, "layer" + std::to_string(i) + ".inputs");
    c.inputs = c.layers.front().inputs;
    c.outputs = c.layers.back().outputs;
    return true;
}

const char *status_name(syldet_status st) {
    switch (st) {
        case SYLDET_ERR_OPEN: return "unableToOpenPath";
        case SYLDET_ERR_MISSING: return "missingValue";
        case SYLDET_ERR_INVALID: return "invalidValue";
        case SYLDET_ERR_MISMATCH: return "mismatchedLength";
        case SYLDET_ERR_CONFIG: return "invalid configuration";
        default: return "error";
    }
}
}  // namespace

syldet_status set_error(syldet_status st, const std::string &msg, const std::string &key) {
    g_err_msg = msg;
    g_err_key = key;
    return st;
}
const std::string &last_error_message() { return g_err_msg; }
const std::string &last_error_key() { return g_err_key; }

syldet_status parse_config_text(const char *text, size_t len, Config &out) {
    Dict d = read_pairs(std::string_view(text, len));
    Reader r{d};
    Config c;
    if (!read_all(r, c)) return set_error(r.st, std::string(status_name(r.st)) + "(\"" + r.key + "\")", r.key);
    out = std::move(c);
    return SYLDET_OK;
}

syldet_status load_config_file(const char *path, Config &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return set_error(SYLDET_ERR_OPEN, std::string("unableToOpenPath(\"") + path + "\")", path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    return parse_config_text(s.data(), s.size(), out);
}

bool frequency_index_range(int fft_len, double f_lo, double f_hi, double rate, int &start, int &end) {
    // CSTFT.swift:166-191
    if (!(f_lo >= 0.0 && f_hi > f_lo)) return false;
    const int half = fft_len / 2;
    const double per_hz = (double)fft_len / rate;
    const double s = std::ceil(per_hz * f_lo);
    if (!(s < (double)half)) return false;
    const double e = std::floor(per_hz * f_hi) + 1.0;
    if (e < s) return false;
    start = (int)s;
    end = e > (double)half ? half : (int)e;
    return true;
}

syldet_status validate_config(Config &c) {
    auto bad = [](const std::string &what, const std::string &key) { return set_error(SYLDET_ERR_CONFIG, what, key); };
    if (c.layers.empty()) return bad("Neural network must have 1 or more layers.", "layers");
    // CSTFT.swift:66-91
    c.gap = c.window_overlap < 0 ? -c.window_overlap : 0;
    c.overlap = c.window_overlap < 0 ? 0 : c.window_overlap;
    if (c.window_length <= 0) return bad("Invalid window length.", "windowLength");
    if (c.window_overlap >= c.window_length) return bad("Invalid overlap value.", "windowOverlap");
    if (c.fourier_length < 2 || (c.fourier_length & (c.fourier_length - 1)) != 0)
        return bad("The FFT size must be a power of 2.", "fourierLength");
    if (c.window_length > c.fourier_length)
        return bad("The FFT size must be greater than or equal to the window length.", "fourierLength");
    c.hop = c.gap + c.window_length - c.overlap;
    // SyllableDetector.swift:46-60
    if (!frequency_index_range(c.fourier_length, c.freq_lo, c.freq_hi, c.sampling_rate, c.k0, c.k1))
        return bad("The frequency range is invalid.", "freqRange");
    c.band = c.k1 - c.k0;
    if (c.time_range < 1) return bad("Invalid time range.", "timeRange");
    c.inputs = c.layers.front().inputs;
    c.outputs = c.layers.back().outputs;
    const long expected = (long)c.band * c.time_range;
    if (expected != c.inputs)
        return bad("The neural network has " + std::to_string(c.inputs) + " inputs, but the configuration settings suggest there should be " +
                       std::to_string(expected) + ".", "layer0.inputs");
    if ((int)c.thresholds.size() != c.outputs)
        return bad("The neural network has " + std::to_string(c.outputs) + " outputs, but the configuration settings suggest there should be " +
                       std::to_string(c.thresholds.size()) + ".", "thresholds");
    c.valid = true;
    return SYLDET_OK;
}

}  // namespace syldet
