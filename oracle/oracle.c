/*
 * oracle.c — CPU restatement of the syllable-detection hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (libsyldet_cuda.so, the
 * host mirror, the CLI) links, loads or calls this file.  It is used by
 * tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs as the checker / CPU baseline.
 *
 * PARITY UNPINNED.  The reference (gardner-lab/syllable-detector-swift @
 * fa1c200) cannot be built here (Swift + Apple Accelerate + AVFoundation) and
 * ships no golden vectors for this path; its arithmetic lives in the closed
 * Accelerate framework (vDSP/vForce, not version pinned).  This file restates
 * the algorithm in IEEE float32 in the order the Swift code calls vDSP, using
 * Apple's documented semantics for each call.  It is anchored by analytic
 * known-answer tests, a float64 numpy twin (oracle/twin64.py) and the
 * reference's only fixture, sample.txt.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off so results do not depend
 * on whether the host CPU has FMA).
 *
 * Reference citations are relative to /root/reference:
 *   CSTFT   = Common/CircularShortTimeFourierTransform.swift
 *   SD      = Common/SyllableDetector.swift
 *   NN      = Common/NeuralNet.swift
 *   CFG     = Common/SyllableDetectorConfig.swift
 *   RS      = Common/Resampler.swift
 *   TD      = SyllableDetectorCLI/TrackDetector.swift
 */
#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_ERR_OPEN 1     /* ParseError.unableToOpenPath   CFG:51 */
#define ORC_ERR_MISSING 2  /* ParseError.missingValue       CFG:52 */
#define ORC_ERR_INVALID 3  /* ParseError.invalidValue       CFG:53 */
#define ORC_ERR_MISMATCH 4 /* ParseError.mismatchedLength   CFG:54 */
#define ORC_ERR_CONFIG 5   /* fatalError() invariants in SD:46-60, CSTFT:76-91, NN:245-255,340-348 */

enum { PROC_MAPMINMAX = 0, PROC_MAPSTD = 1, PROC_L2NORMALIZE = 2, PROC_NORMALIZE = 3, PROC_NORMALIZESTD = 4 };
enum { TF_TANSIG = 0, TF_LOGSIG = 1, TF_PURELIN = 2, TF_SATLIN = 3 };
enum { SCALE_LINEAR = 0, SCALE_LOG = 1, SCALE_DB = 2 };

typedef struct {
    int kind;
    int n;
    float *xoff;
    float *gain;
    float y; /* yMin (mapminmax) or yMean (mapstd) */
} orc_proc;

typedef struct {
    int in, out, tf;
    float *w; /* row-major [out][in]  NN:368 (vDSP_mmul M=out,N=1,P=in), convert_to_text.m:202 */
    float *b;
} orc_layer;

typedef struct {
    double fs;
    int fft_len, win_len, win_overlap;
    double f_lo, f_hi;
    int time_range;
    int n_thr;
    double *thr;
    int scaling;
    int n_layers;
    orc_layer *layers;
    int n_ip;
    orc_proc *ip;
    int n_op;
    orc_proc *op;
    /* derived (SD:37-60, CSTFT:61-129) */
    int gap, overlap, stride, k0, k1, L, I, O;
    float *window;
    float *tw_re, *tw_im; /* e^{-2 pi i k / fft_len}, k < fft_len/2 */
    int max_width;
} orc_config;

/* ------------------------------------------------------------------ */
/* key = value store (CFG:176-189, Common.swift:16-24)                  */

typedef struct {
    char **keys;
    char **vals;
    int n, cap;
} kv_store;

static int is_ws(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

static char *trim_dup(const char *s, size_t len) {
    while (len > 0 && is_ws((unsigned char)s[0])) { s++; len--; }
    while (len > 0 && is_ws((unsigned char)s[len - 1])) len--;
    char *r = (char *)malloc(len + 1);
    memcpy(r, s, len);
    r[len] = 0;
    return r;
}

static void kv_set(kv_store *kv, char *k, char *v) {
    for (int i = 0; i < kv->n; i++)
        if (strcmp(kv->keys[i], k) == 0) { /* later lines overwrite (dictionary assignment CFG:187) */
            free(kv->vals[i]);
            kv->vals[i] = v;
            free(k);
            return;
        }
    if (kv->n == kv->cap) {
        kv->cap = kv->cap ? kv->cap * 2 : 64;
        kv->keys = (char **)realloc(kv->keys, sizeof(char *) * kv->cap);
        kv->vals = (char **)realloc(kv->vals, sizeof(char *) * kv->cap);
    }
    kv->keys[kv->n] = k;
    kv->vals[kv->n] = v;
    kv->n++;
}

static const char *kv_get(const kv_store *kv, const char *k) {
    for (int i = 0; i < kv->n; i++)
        if (strcmp(kv->keys[i], k) == 0) return kv->vals[i];
    return NULL;
}

static void kv_free(kv_store *kv) {
    for (int i = 0; i < kv->n; i++) { free(kv->keys[i]); free(kv->vals[i]); }
    free(kv->keys);
    free(kv->vals);
}

/* Swift's split(omittingEmptySubsequences: true) on '=' must give exactly two pieces (CFG:185-188). */
static void kv_parse_line(kv_store *kv, const char *line, size_t len) {
    size_t starts[3], lens[3];
    int np = 0;
    size_t i = 0;
    while (i < len) {
        while (i < len && line[i] == '=') i++;
        if (i >= len) break;
        size_t s = i;
        while (i < len && line[i] != '=') i++;
        if (np < 3) { starts[np] = s; lens[np] = i - s; }
        np++;
    }
    if (np != 2) return;
    kv_set(kv, trim_dup(line + starts[0], lens[0]), trim_dup(line + starts[1], lens[1]));
}

static void kv_parse_text(kv_store *kv, const char *buf, size_t len) {
    size_t s = 0;
    for (size_t i = 0; i <= len; i++) {
        if (i == len || buf[i] == '\n') { /* StreamReader delimiter "\n" (StreamReader.swift:19) */
            if (i > s || i < len) kv_parse_line(kv, buf + s, i - s);
            s = i + 1;
        }
    }
}

/* ------------------------------------------------------------------ */
/* strict scalar parsing: Swift Double(String)/Float(String)/Int(String) take the whole string. */

static int parse_double_strict(const char *s, double *out) {
    if (!*s || is_ws((unsigned char)*s)) return 0;
    char *end;
    errno = 0;
    double d = strtod(s, &end);
    if (end == s || *end) return 0;
    *out = d;
    return 1;
}

static int parse_float_strict(const char *s, float *out) {
    if (!*s || is_ws((unsigned char)*s)) return 0;
    char *end;
    errno = 0;
    float f = strtof(s, &end); /* decimal -> binary32 in ONE rounding, like Swift Float(String) */
    if (end == s || *end) return 0;
    *out = f;
    return 1;
}

static int parse_int_strict(const char *s, long *out) {
    const char *p = s;
    if (*p == '+' || *p == '-') p++;
    if (!*p) return 0;
    for (const char *q = p; *q; q++)
        if (!isdigit((unsigned char)*q)) return 0;
    errno = 0;
    char *end;
    long v = strtol(s, &end, 10);
    if (errno || *end) return 0;
    *out = v;
    return 1;
}

typedef struct {
    int code;
    char key[192];
} orc_err;

static int fail(orc_err *e, int code, const char *key) {
    e->code = code;
    snprintf(e->key, sizeof e->key, "%s", key);
    return code;
}

static int get_double(const kv_store *kv, const char *k, double *out, orc_err *e) {
    const char *v = kv_get(kv, k);
    if (!v) return fail(e, ORC_ERR_MISSING, k);
    if (!parse_double_strict(v, out)) return fail(e, ORC_ERR_INVALID, k);
    return 0;
}
static int get_float(const kv_store *kv, const char *k, float *out, orc_err *e) {
    const char *v = kv_get(kv, k);
    if (!v) return fail(e, ORC_ERR_MISSING, k);
    if (!parse_float_strict(v, out)) return fail(e, ORC_ERR_INVALID, k);
    return 0;
}
static int get_int(const kv_store *kv, const char *k, int *out, orc_err *e) {
    const char *v = kv_get(kv, k);
    long l;
    if (!v) return fail(e, ORC_ERR_MISSING, k);
    if (!parse_int_strict(v, &l)) return fail(e, ORC_ERR_INVALID, k);
    *out = (int)l;
    return 0;
}

/* CFG:81-113: split at ',', drop empty pieces, trim, every piece must parse, then count check. */
static int get_array(const kv_store *kv, const char *k, int want, int is_double, void **out, int *n_out, orc_err *e) {
    const char *v = kv_get(kv, k);
    if (!v) return fail(e, ORC_ERR_MISSING, k);
    size_t len = strlen(v), i = 0;
    int n = 0, cap = 16, bad = 0;
    double *d = (double *)malloc(sizeof(double) * cap);
    float *f = (float *)malloc(sizeof(float) * cap);
    while (i < len) {
        while (i < len && v[i] == ',') i++;
        if (i >= len) break;
        size_t s = i;
        while (i < len && v[i] != ',') i++;
        char *piece = trim_dup(v + s, i - s);
        if (n == cap) {
            cap *= 2;
            d = (double *)realloc(d, sizeof(double) * cap);
            f = (float *)realloc(f, sizeof(float) * cap);
        }
        int ok = is_double ? parse_double_strict(piece, &d[n]) : parse_float_strict(piece, &f[n]);
        free(piece);
        if (!ok) bad = 1;
        n++;
    }
    if (bad) { free(d); free(f); return fail(e, ORC_ERR_INVALID, k); }
    if (want >= 0 && n != want) { free(d); free(f); return fail(e, ORC_ERR_MISMATCH, k); }
    if (is_double) { *out = d; free(f); } else { *out = f; free(d); }
    *n_out = n;
    return 0;
}

/* ------------------------------------------------------------------ */

void orc_config_free(orc_config *c) {
    if (!c) return;
    free(c->thr);
    for (int i = 0; i < c->n_layers; i++) { free(c->layers[i].w); free(c->layers[i].b); }
    free(c->layers);
    for (int i = 0; i < c->n_ip; i++) { free(c->ip[i].xoff); free(c->ip[i].gain); }
    free(c->ip);
    for (int i = 0; i < c->n_op; i++) { free(c->op[i].xoff); free(c->op[i].gain); }
    free(c->op);
    free(c->window);
    free(c->tw_re);
    free(c->tw_im);
    free(c);
}

static int parse_proc(const kv_store *kv, const char *nm, int cnt, int is_input, orc_proc *p, orc_err *e) {
    char key[160];
    snprintf(key, sizeof key, "%s.function", nm);
    const char *fn = kv_get(kv, key);
    if (!fn) return fail(e, ORC_ERR_MISSING, key);
    memset(p, 0, sizeof *p);
    p->n = cnt;
    int is_mm = strcmp(fn, "mapminmax") == 0, is_ms = strcmp(fn, "mapstd") == 0;
    if (is_mm || is_ms) { /* CFG:115-127 */
        int n;
        char k2[160];
        p->kind = is_mm ? PROC_MAPMINMAX : PROC_MAPSTD;
        snprintf(k2, sizeof k2, "%s.xOffsets", nm);
        if (get_array(kv, k2, cnt, 0, (void **)&p->xoff, &n, e)) return e->code;
        snprintf(k2, sizeof k2, "%s.gains", nm);
        if (get_array(kv, k2, cnt, 0, (void **)&p->gain, &n, e)) return e->code;
        snprintf(k2, sizeof k2, is_mm ? "%s.yMin" : "%s.yMean", nm);
        if (get_float(kv, k2, &p->y, e)) return e->code;
        return 0;
    }
    if (is_input) { /* CFG:141-151 */
        if (strcmp(fn, "l2normalize") == 0) { p->kind = PROC_L2NORMALIZE; return 0; }
        if (strcmp(fn, "normalize") == 0) { p->kind = PROC_NORMALIZE; return 0; }
        if (strcmp(fn, "normalizestd") == 0) { p->kind = PROC_NORMALIZESTD; return 0; }
    }
    return fail(e, ORC_ERR_INVALID, key);
}

static int is_pow2(long v) { return v != 0 && (v & (v - 1)) == 0; } /* Common.swift:27-29 */

/* CSTFT:166-191 */
static int freq_index_range(int fft_len, double f0, double f1, double rate, int *k0, int *k1) {
    if (!(f0 >= 0.0 && f1 > f0)) return 0;
    int half = fft_len / 2;
    double from = (double)fft_len / rate;
    long s = (long)ceil(from * f0);
    if (s >= half) return 0;
    long en = (long)floor(from * f1) + 1;
    if (en < s) return 0;
    if (en > half) en = half;
    *k0 = (int)s;
    *k1 = (int)en;
    return 1;
}

/* window tables, vDSP N-denominator forms (CSTFT:19-28); detector forces Hamming (SD:43) */
void orc_make_window(int type, int n, float *w) {
    for (int i = 0; i < n; i++) {
        double a = 2.0 * M_PI * (double)i / (double)n, v;
        switch (type) {
            case 1: v = 0.54 - 0.46 * cos(a); break;                  /* vDSP_hamm_window  */
            case 2: v = 0.5 * (1.0 - cos(a)); break;                  /* vDSP_hann_window, DENORM */
            case 3: v = 0.42 - 0.5 * cos(a) + 0.08 * cos(2.0 * a); break; /* vDSP_blkman_window */
            default: v = 1.0;
        }
        w[i] = (float)v;
    }
}

static int derive(orc_config *c, orc_err *e) {
    /* NN:245-255 */
    if (c->n_layers < 1) return fail(e, ORC_ERR_CONFIG, "layers");
    for (int i = 0; i < c->n_layers; i++) {
        if (c->layers[i].in <= 0 || c->layers[i].out <= 0) return fail(e, ORC_ERR_CONFIG, "layer dims");
        if (i > 0 && c->layers[i - 1].out != c->layers[i].in) return fail(e, ORC_ERR_CONFIG, "layer chain");
    }
    c->I = c->layers[0].in;
    c->O = c->layers[c->n_layers - 1].out;
    /* CSTFT:66-91 */
    if (c->win_overlap < 0) { c->gap = -c->win_overlap; c->overlap = 0; } else { c->gap = 0; c->overlap = c->win_overlap; }
    if (c->win_len <= 0) return fail(e, ORC_ERR_CONFIG, "windowLength");
    if (c->win_overlap >= c->win_len) return fail(e, ORC_ERR_CONFIG, "windowOverlap");
    if (!is_pow2(c->fft_len) || c->win_len > c->fft_len || c->fft_len < 2) return fail(e, ORC_ERR_CONFIG, "fourierLength");
    c->stride = c->gap + c->win_len - c->overlap;
    /* SD:46-60 */
    if (!freq_index_range(c->fft_len, c->f_lo, c->f_hi, c->fs, &c->k0, &c->k1)) return fail(e, ORC_ERR_CONFIG, "freqRange");
    c->L = c->k1 - c->k0;
    if (c->time_range < 1) return fail(e, ORC_ERR_CONFIG, "timeRange");
    if ((long)c->L * c->time_range != c->I) return fail(e, ORC_ERR_CONFIG, "inputs");
    if (c->n_thr != c->O) return fail(e, ORC_ERR_CONFIG, "thresholds");
    c->window = (float *)malloc(sizeof(float) * c->win_len);
    orc_make_window(1, c->win_len, c->window);
    int half = c->fft_len / 2;
    c->tw_re = (float *)malloc(sizeof(float) * (half > 0 ? half : 1));
    c->tw_im = (float *)malloc(sizeof(float) * (half > 0 ? half : 1));
    for (int k = 0; k < half; k++) {
        double a = -2.0 * M_PI * (double)k / (double)c->fft_len;
        c->tw_re[k] = (float)cos(a);
        c->tw_im[k] = (float)sin(a);
    }
    c->max_width = c->I;
    for (int i = 0; i < c->n_layers; i++)
        if (c->layers[i].out > c->max_width) c->max_width = c->layers[i].out;
    return 0;
}

/* CFG:170-277, in the same order so the first failing key matches. */
int orc_config_parse(const char *buf, size_t len, orc_config **out, int *err_code, char *err_key, int err_key_len) {
    kv_store kv = {0};
    orc_err e = {0};
    orc_config *c = (orc_config *)calloc(1, sizeof *c);
    kv_parse_text(&kv, buf, len);
    int n;
    do {
        if (get_double(&kv, "samplingRate", &c->fs, &e)) break;
        if (get_int(&kv, "fourierLength", &c->fft_len, &e)) break;
        if (!is_pow2(c->fft_len)) { fail(&e, ORC_ERR_INVALID, "fourierLength"); break; }
        if (!kv_get(&kv, "windowLength")) c->win_len = c->fft_len;
        else if (get_int(&kv, "windowLength", &c->win_len, &e)) break;
        if (get_int(&kv, "windowOverlap", &c->win_overlap, &e)) break;
        double *fr = NULL;
        if (get_array(&kv, "freqRange", 2, 1, (void **)&fr, &n, &e)) break;
        c->f_lo = fr[0];
        c->f_hi = fr[1];
        free(fr);
        if (get_int(&kv, "timeRange", &c->time_range, &e)) break;
        if (get_array(&kv, "thresholds", -1, 1, (void **)&c->thr, &c->n_thr, &e)) { /* CFG:223-229 */
            if (get_array(&kv, "threshold", -1, 1, (void **)&c->thr, &c->n_thr, &e)) break;
            e.code = 0;
        }
        const char *sc = kv_get(&kv, "scaling");
        if (!sc) { fail(&e, ORC_ERR_MISSING, "scaling"); break; }
        if (strcmp(sc, "linear") == 0) c->scaling = SCALE_LINEAR;
        else if (strcmp(sc, "log") == 0) c->scaling = SCALE_LOG;
        else if (strcmp(sc, "db") == 0) c->scaling = SCALE_DB;
        else { fail(&e, ORC_ERR_INVALID, "scaling"); break; }
        int nl;
        if (get_int(&kv, "layers", &nl, &e)) break;
        if (nl < 0) { fail(&e, ORC_ERR_CONFIG, "layers"); break; } /* Swift: 0..<negative traps */
        c->layers = (orc_layer *)calloc(nl > 0 ? nl : 1, sizeof(orc_layer));
        for (int i = 0; i < nl && !e.code; i++) {
            char k[64];
            orc_layer *l = &c->layers[i];
            c->n_layers = i + 1;
            snprintf(k, sizeof k, "layer%d.inputs", i);
            if (get_int(&kv, k, &l->in, &e)) break;
            snprintf(k, sizeof k, "layer%d.outputs", i);
            if (get_int(&kv, k, &l->out, &e)) break;
            if ((long)l->in * l->out < 0) { fail(&e, ORC_ERR_MISMATCH, k); break; }
            snprintf(k, sizeof k, "layer%d.weights", i);
            if (get_array(&kv, k, l->in * l->out, 0, (void **)&l->w, &n, &e)) break;
            snprintf(k, sizeof k, "layer%d.biases", i);
            if (get_array(&kv, k, l->out, 0, (void **)&l->b, &n, &e)) break;
            snprintf(k, sizeof k, "layer%d.transferFunction", i);
            const char *tf = kv_get(&kv, k);
            if (!tf) { fail(&e, ORC_ERR_MISSING, k); break; }
            if (strcmp(tf, "TanSig") == 0) l->tf = TF_TANSIG;
            else if (strcmp(tf, "LogSig") == 0) l->tf = TF_LOGSIG;
            else if (strcmp(tf, "PureLin") == 0) l->tf = TF_PURELIN;
            else if (strcmp(tf, "SatLin") == 0) l->tf = TF_SATLIN;
            else { fail(&e, ORC_ERR_INVALID, k); break; }
            /* NN:340-348 (fatalError upstream, raised while parsing) */
            if (l->in <= 0 || l->out <= 0) { fail(&e, ORC_ERR_CONFIG, k); break; }
        }
        if (e.code) break;
        int nip, nop;
        if (get_int(&kv, "processInputsCount", &nip, &e)) break;
        if (nip > 0 && c->n_layers < 1) { fail(&e, ORC_ERR_CONFIG, "layers"); break; }
        c->ip = (orc_proc *)calloc(nip > 0 ? nip : 1, sizeof(orc_proc));
        for (int i = 0; i < nip && !e.code; i++) {
            char k[64];
            snprintf(k, sizeof k, "processInputs%d", i);
            c->n_ip = i + 1;
            parse_proc(&kv, k, c->layers[0].in, 1, &c->ip[i], &e);
        }
        if (e.code) break;
        if (get_int(&kv, "processOutputsCount", &nop, &e)) break;
        if (nop > 0 && c->n_layers < 1) { fail(&e, ORC_ERR_CONFIG, "layers"); break; }
        c->op = (orc_proc *)calloc(nop > 0 ? nop : 1, sizeof(orc_proc));
        for (int i = 0; i < nop && !e.code; i++) {
            char k[64];
            snprintf(k, sizeof k, "processOutputs%d", i);
            c->n_op = i + 1;
            parse_proc(&kv, k, c->layers[c->n_layers - 1].out, 0, &c->op[i], &e);
        }
        if (e.code) break;
        derive(c, &e);
    } while (0);
    kv_free(&kv);
    if (err_code) *err_code = e.code;
    if (err_key && err_key_len > 0) snprintf(err_key, err_key_len, "%s", e.key);
    if (e.code) { orc_config_free(c); *out = NULL; return e.code; }
    *out = c;
    return 0;
}

int orc_config_load(const char *path, orc_config **out, int *err_code, char *err_key, int err_key_len) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        if (err_code) *err_code = ORC_ERR_OPEN;
        if (err_key && err_key_len > 0) snprintf(err_key, err_key_len, "%s", path);
        *out = NULL;
        return ORC_ERR_OPEN;
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc(sz + 1);
    size_t rd = fread(buf, 1, sz, f);
    fclose(f);
    int r = orc_config_parse(buf, rd, out, err_code, err_key, err_key_len);
    free(buf);
    return r;
}

/* ---- getters for the tests --------------------------------------- */
double orc_cfg_scalar(const orc_config *c, const char *name) {
#define S(n, v) if (strcmp(name, n) == 0) return (double)(v)
    S("samplingRate", c->fs); S("fourierLength", c->fft_len); S("windowLength", c->win_len);
    S("windowOverlap", c->win_overlap); S("freqLo", c->f_lo); S("freqHi", c->f_hi);
    S("timeRange", c->time_range); S("scaling", c->scaling); S("layers", c->n_layers);
    S("processInputsCount", c->n_ip); S("processOutputsCount", c->n_op); S("nThresholds", c->n_thr);
    S("gap", c->gap); S("overlap", c->overlap); S("stride", c->stride); S("k0", c->k0); S("k1", c->k1);
    S("L", c->L); S("I", c->I); S("O", c->O);
#undef S
    int i;
    char f[32];
    if (sscanf(name, "layer%d.%31s", &i, f) == 2 && i >= 0 && i < c->n_layers) {
        if (!strcmp(f, "inputs")) return c->layers[i].in;
        if (!strcmp(f, "outputs")) return c->layers[i].out;
        if (!strcmp(f, "transferFunction")) return c->layers[i].tf;
    }
    if (sscanf(name, "processInputs%d.%31s", &i, f) == 2 && i >= 0 && i < c->n_ip) {
        if (!strcmp(f, "function")) return c->ip[i].kind;
        if (!strcmp(f, "y")) return c->ip[i].y;
    }
    if (sscanf(name, "processOutputs%d.%31s", &i, f) == 2 && i >= 0 && i < c->n_op) {
        if (!strcmp(f, "function")) return c->op[i].kind;
        if (!strcmp(f, "y")) return c->op[i].y;
    }
    return NAN;
}

long orc_cfg_array(const orc_config *c, const char *name, double *out, long cap) {
    const float *src = NULL;
    const double *dsrc = NULL;
    long n = 0;
    int i;
    char f[32];
    if (!strcmp(name, "thresholds")) { dsrc = c->thr; n = c->n_thr; }
    else if (!strcmp(name, "window")) { src = c->window; n = c->win_len; }
    else if (sscanf(name, "layer%d.%31s", &i, f) == 2 && i >= 0 && i < c->n_layers) {
        if (!strcmp(f, "weights")) { src = c->layers[i].w; n = (long)c->layers[i].in * c->layers[i].out; }
        else if (!strcmp(f, "biases")) { src = c->layers[i].b; n = c->layers[i].out; }
    } else if (sscanf(name, "processInputs%d.%31s", &i, f) == 2 && i >= 0 && i < c->n_ip) {
        if (!strcmp(f, "xOffsets") && c->ip[i].xoff) { src = c->ip[i].xoff; n = c->ip[i].n; }
        else if (!strcmp(f, "gains") && c->ip[i].gain) { src = c->ip[i].gain; n = c->ip[i].n; }
    } else if (sscanf(name, "processOutputs%d.%31s", &i, f) == 2 && i >= 0 && i < c->n_op) {
        if (!strcmp(f, "xOffsets") && c->op[i].xoff) { src = c->op[i].xoff; n = c->op[i].n; }
        else if (!strcmp(f, "gains") && c->op[i].gain) { src = c->op[i].gain; n = c->op[i].n; }
    } else return -1;
    for (long j = 0; j < n && j < cap; j++) out[j] = src ? (double)src[j] : dsrc[j];
    return n;
}

/* ------------------------------------------------------------------ */
/* STFT column: CSTFT.extractPower (CSTFT:280-337)                      */

/* Counts: columns C = floor((n - gap - W)/s) + 1, evaluations E = C - T + 1. */
long orc_num_columns(const orc_config *c, long n) {
    long need = (long)c->gap + c->win_len;
    if (n < need) return 0;
    return (n - need) / c->stride + 1;
}
long orc_num_evals(const orc_config *c, long n) {
    long C = orc_num_columns(c, n);
    long E = C - c->time_range + 1;
    return E > 0 ? E : 0;
}
/* TD:39-42: sample number of evaluation j */
long orc_eval_sample(const orc_config *c, long j) {
    return (long)c->gap + c->win_len + (long)c->stride * (c->time_range - 1) + (long)c->stride * j;
}

/*
 * One windowed frame -> |X[k]|, k < N/2, float32.
 *   vDSP_vmul (CSTFT:311) -> zero padded to N (CSTFT:110) -> vDSP_ctoz even/odd split (CSTFT:314-316)
 *   -> vDSP_fft_zript radix-2 packed real FFT, output = 2 x DFT (CSTFT:320) -> imag[0] = 0 (CSTFT:323)
 *   -> vDSP_zvabs sqrt(re^2+im^2) (CSTFT:329) -> vDSP_vsdiv by 2 (CSTFT:332-333).
 * scratch: 2*N floats.
 */
void orc_frame_spectrum(const orc_config *c, const float *frame, float *out, float *scratch) {
    const int N = c->fft_len, M = N / 2, W = c->win_len;
    float *zr = scratch, *zi = scratch + M, *xw = scratch + N;
    for (int m = 0; m < N; m++) xw[m] = m < W ? frame[m] * c->window[m] : 0.0f;
    if (M == 1) { /* N == 2 */
        out[0] = fabsf(2.0f * (xw[0] + xw[1])) / 2.0f;
        return;
    }
    /* ctoz + bit reversal over M complex points */
    int bits = 0;
    while ((1 << bits) < M) bits++;
    for (int n = 0; n < M; n++) {
        int r = 0;
        for (int b = 0; b < bits; b++)
            if (n & (1 << b)) r |= 1 << (bits - 1 - b);
        zr[r] = xw[2 * n];
        zi[r] = xw[2 * n + 1];
    }
    /* radix-2 decimation-in-time, twiddle e^{-2 pi i j / len} = tw[j * (N/len)] */
    for (int len = 2; len <= M; len <<= 1) {
        int half = len / 2, step = N / len;
        for (int base = 0; base < M; base += len)
            for (int j = 0; j < half; j++) {
                float wr = c->tw_re[j * step], wi = c->tw_im[j * step];
                int a = base + j, b = a + half;
                float tr = zr[b] * wr - zi[b] * wi;
                float ti = zr[b] * wi + zi[b] * wr;
                zr[b] = zr[a] - tr;
                zi[b] = zi[a] - ti;
                zr[a] = zr[a] + tr;
                zi[a] = zi[a] + ti;
            }
    }
    /* real-FFT untangle, vDSP scaling (2 x DFT): A[k] = (Z[k] + conj Z[M-k]) - i w^k (Z[k] - conj Z[M-k]) */
    {
        float re0 = 2.0f * (zr[0] + zi[0]); /* realp[0] = 2 X[0]; imagp[0] = 2 X[N/2] is zeroed (CSTFT:323) */
        out[0] = sqrtf(re0 * re0 + 0.0f * 0.0f) / 2.0f;
    }
    for (int k = 1; k < M; k++) {
        float ar = zr[k], ai = zi[k], br = zr[M - k], bi = -zi[M - k];
        float sr = ar + br, si = ai + bi; /* Z[k] + conj Z[M-k] */
        float dr = ar - br, di = ai - bi; /* Z[k] - conj Z[M-k] */
        float wr = c->tw_re[k], wi = c->tw_im[k];
        /* -i * w * d = -i (wr + i wi)(dr + i di) = (wr di + wi dr) - i (wr dr - wi di) */
        float re = sr + (wr * di + wi * dr);
        float im = si - (wr * dr - wi * di);
        out[k] = sqrtf(re * re + im * im) / 2.0f;
    }
}

/* Band-sliced columns of one channel: out[C][L] (SD:134-151). Returns C. */
long orc_stft_band(const orc_config *c, const float *x, long n, float *out) {
    long C = orc_num_columns(c, n);
    float *scratch = (float *)malloc(sizeof(float) * (2 * c->fft_len + c->fft_len / 2 + 1));
    float *spec = scratch + 2 * c->fft_len;
    for (long col = 0; col < C; col++) {
        orc_frame_spectrum(c, x + col * c->stride + c->gap, spec, scratch);
        memcpy(out + col * c->L, spec + c->k0, sizeof(float) * c->L);
    }
    free(scratch);
    return C;
}

/* ------------------------------------------------------------------ */
/* NeuralNet.apply (NN:294-326). v: I scaled features. buf: 2*max_width floats scratch. */

static void proc_forward(const orc_proc *p, float *x, int n) {
    switch (p->kind) {
        case PROC_MAPMINMAX: /* NN:127-131  vsbm then vsadd */
            for (int i = 0; i < n; i++) { float t = (x[i] - p->xoff[i]) * p->gain[i]; x[i] = t + p->y; }
            break;
        case PROC_MAPSTD: /* NN:162-169 */
            for (int i = 0; i < n; i++) x[i] = (x[i] - p->xoff[i]) * p->gain[i];
            if (0 != p->y) for (int i = 0; i < n; i++) x[i] = x[i] + p->y;
            break;
        case PROC_L2NORMALIZE: { /* NN:47-59 */
            float ss = 0.0f;
            for (int i = 0; i < n; i++) ss += x[i] * x[i];
            float d = sqrtf(ss);
            for (int i = 0; i < n; i++) x[i] = x[i] / d;
            break;
        }
        case PROC_NORMALIZE: { /* NN:69-96 */
            float mn = x[0], mx = x[0];
            for (int i = 1; i < n; i++) { if (x[i] < mn) mn = x[i]; if (x[i] > mx) mx = x[i]; }
            float range = mx - mn;
            if (0 == range) { for (int i = 0; i < n; i++) x[i] = -1.0f; break; }
            float slope = 2.0f / range, icpt = (0 - mn - mx) / range;
            for (int i = 0; i < n; i++) { float t = x[i] * slope; x[i] = t + icpt; }
            break;
        }
        case PROC_NORMALIZESTD: { /* NN:105-108 vDSP_normalize: population mean / std */
            float s = 0.0f;
            for (int i = 0; i < n; i++) s += x[i];
            float mean = s / (float)n, v = 0.0f;
            for (int i = 0; i < n; i++) { float d = x[i] - mean; v += d * d; }
            float sd = sqrtf(v / (float)n);
            for (int i = 0; i < n; i++) x[i] = (x[i] - mean) / sd;
            break;
        }
    }
}

static void proc_reverse(const orc_proc *p, float *y, int n) {
    /* NN:137-143 / 175-181: (y + (-yMin)) / gain + xOffset */
    float neg = 0 - p->y;
    for (int i = 0; i < n; i++) { float t = y[i] + neg; t = t / p->gain[i]; y[i] = t + p->xoff[i]; }
}

static void transfer(int tf, float *v, int n) {
    switch (tf) {
        case TF_TANSIG: for (int i = 0; i < n; i++) v[i] = tanhf(v[i]); break;           /* NN:189-194 */
        case TF_LOGSIG:                                                                  /* NN:196-215 */
            for (int i = 0; i < n; i++) { float t = v[i] * -1.0f; t = expf(t); t = t + 1.0f; v[i] = 1.0f / t; }
            break;
        case TF_SATLIN:                                                                  /* NN:223-228 */
            for (int i = 0; i < n; i++) { if (v[i] < 0.0f) v[i] = 0.0f; else if (v[i] > 1.0f) v[i] = 1.0f; }
            break;
        default: break;
    }
}

void orc_net_apply_scaled(const orc_config *c, const float *v, float *out, float *buf) {
    float *cur = buf, *nxt = buf + c->max_width;
    memcpy(cur, v, sizeof(float) * c->I);
    for (int k = 0; k < c->n_ip; k++) proc_forward(&c->ip[k], cur, c->I);
    for (int l = 0; l < c->n_layers; l++) { /* NN:366-377 */
        const orc_layer *ly = &c->layers[l];
        for (int o = 0; o < ly->out; o++) {
            float acc = 0.0f;
            const float *w = ly->w + (long)o * ly->in;
            for (int i = 0; i < ly->in; i++) acc += w[i] * cur[i];
            nxt[o] = acc + ly->b[o];
        }
        transfer(ly->tf, nxt, ly->out);
        float *t = cur; cur = nxt; nxt = t;
    }
    for (int o = 0; o < c->O; o++) out[o] = cur[o];
    for (int k = 0; k < c->n_op; k++) proc_reverse(&c->op[k], out, c->O);
}

/* SD:184-212. `log` upstream is broken (vvlogf args swapped, SD:207); intent ln(x) is implemented (DEVIATION). */
static void scale_features(const orc_config *c, const float *in, float *out, int n) {
    switch (c->scaling) {
        case SCALE_DB: for (int i = 0; i < n; i++) out[i] = 20.0f * log10f(in[i] / 1.0f); break;
        case SCALE_LOG: for (int i = 0; i < n; i++) out[i] = logf(in[i]); break;
        default: memcpy(out, in, sizeof(float) * n);
    }
}

/* raw (unscaled) feature vector -> outputs; used by unit tests */
void orc_net_apply(const orc_config *c, const float *v, float *out) {
    float *buf = (float *)malloc(sizeof(float) * (2 * c->max_width + c->I));
    scale_features(c, v, buf + 2 * c->max_width, c->I);
    orc_net_apply_scaled(c, buf + 2 * c->max_width, out, buf);
    free(buf);
}

/*
 * Whole path for one channel (SD.processNewValue in a loop, TD:65-77).
 * outputs[E][O] (may be NULL), det_any[E] = any(double(out[i]) >= thr[i]) (TD:71-77),
 * det_first[E] = double(out[0]) >= thr[0] (SD:27-31). Either flag array may be NULL. band (C*L floats) optional.
 */
long orc_run(const orc_config *c, const float *x, long n, float *outputs, unsigned char *det_any,
             unsigned char *det_first, float *band_out) {
    long C = orc_num_columns(c, n), E = orc_num_evals(c, n);
    if (C == 0) return 0;
    float *band = band_out ? band_out : (float *)malloc(sizeof(float) * C * c->L);
    orc_stft_band(c, x, n, band);
    float *buf = (float *)malloc(sizeof(float) * (2 * c->max_width + c->I + c->O));
    float *sv = buf + 2 * c->max_width, *o = sv + c->I;
    for (long j = 0; j < E; j++) {
        /* v[t*L+f] = P_{j+t}[k0+f]: contiguous in the column stream (SD:140-148,175-178) */
        scale_features(c, band + j * c->L, sv, c->I);
        orc_net_apply_scaled(c, sv, o, buf);
        int any = 0;
        for (int i = 0; i < c->O; i++) if ((double)o[i] >= c->thr[i]) { any = 1; break; }
        if (outputs) memcpy(outputs + j * c->O, o, sizeof(float) * c->O);
        if (det_any) det_any[j] = (unsigned char)any;
        if (det_first) det_first[j] = (unsigned char)((double)o[0] >= c->thr[0]);
    }
    free(buf);
    if (!band_out) free(band);
    return E;
}

/* Many channels, planar with `stride` samples between channels; OpenMP over channels (cpu_baseline all-cores leg).
 * outputs[ch][E][O]; det[ch][E]. */
long orc_run_multi(const orc_config *c, const float *x, int n_channels, long n, long stride, int n_threads,
                   float *outputs, unsigned char *det_any) {
    long E = orc_num_evals(c, n);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int ch = 0; ch < n_channels; ch++)
        orc_run(c, x + (long)ch * stride, n, outputs ? outputs + (long)ch * E * c->O : NULL,
                det_any ? det_any + (long)ch * E : NULL, NULL, NULL);
    return E;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* TrackDetector debounce (TD:18-30, 80, 99): emit iff det && debounceUntil < S_j; then until = S_j + D. */
long orc_debounce(const orc_config *c, const unsigned char *det, long E, long debounce_frames, long *ev_eval) {
    long until = -1, n = 0;
    for (long j = 0; j < E; j++) {
        long S = orc_eval_sample(c, j);
        if (det[j] && until < S) {
            ev_eval[n++] = j;
            until = S + debounce_frames;
        }
    }
    return n;
}
long orc_debounce_frames(const orc_config *c, double seconds) { return (long)(seconds * c->fs); } /* TD:23-25 */

/* ------------------------------------------------------------------ */
/* ResamplerLinear (RS:20-70), float32 including the index ramp.        */
typedef struct {
    float step, last, offset;
} orc_resampler;

orc_resampler *orc_resampler_new(double rate_in, double rate_out) {
    orc_resampler *r = (orc_resampler *)calloc(1, sizeof *r);
    r->step = (float)(rate_in / rate_out); /* RS:32 */
    return r;
}
void orc_resampler_free(orc_resampler *r) { free(r); }
void orc_resampler_state(const orc_resampler *r, float *step, float *last, float *offset) {
    *step = r->step; *last = r->last; *offset = r->offset;
}

/* Returns number of output samples (<= cap), or -1 if cap is too small. */
long orc_resampler_process(orc_resampler *r, const float *x, long n_in, float *out, long cap) {
    if (n_in <= 0) return 0;
    int across = r->offset < 0;                                   /* RS:37 */
    long n_out = (long)(((float)n_in - r->offset) / r->step);     /* RS:40 */
    if (n_out > cap) return -1;
    if (n_out <= 0) {
        /* upstream indexes indices[-1] here (undefined); we carry the phase forward instead */
        r->offset = r->offset - (float)n_in;
        r->last = x[n_in - 1];
        return 0;
    }
    float last_idx = 0.0f;
    for (long k = 0; k < n_out; k++) {
        float fk = (float)k;
        float idx = r->offset + fk * r->step;                     /* vDSP_vramp RS:52 */
        if (k == 0 && across) idx = 0.0f;                         /* RS:54-56 */
        long b = (long)idx;                                       /* vDSP_vlint: trunc */
        float a = idx - (float)b;
        float x0 = x[b];
        float x1 = (b + 1 < n_in) ? x[b + 1] : x[n_in - 1];       /* upstream reads past the end when step<1; we hold */
        out[k] = x0 + a * (x1 - x0);
        last_idx = idx;
    }
    if (across) out[0] = (r->last * (0 - r->offset)) + (x[0] * (1 + r->offset)); /* RS:61-63 */
    r->offset = last_idx + r->step - (float)(n_in - 1);           /* RS:65 */
    r->last = x[n_in - 1];                                        /* RS:66 */
    return n_out;
}
