/* oracle/lightctr_oracle.c -- TEST INFRASTRUCTURE ONLY (see lightctr_oracle.h).
 *
 * Plain-C restatement of the LightCTR FM/FFM/NFM hot path.  Arithmetic follows the reference
 * expression by expression: AVX lane order of avx_dotProduct, separate
 * mul/add roundings (compile with -ffp-contract=off), float/double promotions where the C++
 * source has them.  Each function cites the reference file:line it restates.
 *
 * "parity pinned": tests/test_oracle_vs_ref.py checks this file bit-for-bit against
 * oracle/_ref/libref.so (the unmodified reference compiled in place) wherever /root/reference is
 * available, and tests/test_oracle_golden.py checks it against tests/golden/ everywhere.
 */
#include "lightctr_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ============================ glibc rand() (random_r TYPE_3), restated ======================= */
static uint32_t g_r[34 + 310 + 4];
static int32_t g_ring[34];
static int g_ri = 0;
void orc_srand(unsigned seed) {
    int32_t r[344];
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    for (int i = 34; i < 344; i++) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
    for (int i = 0; i < 34; i++) g_ring[i] = r[310 + i];
    g_ri = 0;
    (void)g_r;
}
int orc_rand(void) {
    /* ring holds the last 34 values; new = v[-31] + v[-3] */
    int i31 = (g_ri + 34 - 31) % 34, i3 = (g_ri + 34 - 3) % 34;
    uint32_t v = (uint32_t)g_ring[i31] + (uint32_t)g_ring[i3];
    g_ring[g_ri] = (int32_t)v;
    g_ri = (g_ri + 1) % 34;
    return (int)(v >> 1);
}
#define ORC_RAND_MAX 2147483647
double orc_uniform(void) { return (double)orc_rand() / ((double)ORC_RAND_MAX + 1.0); }        /* random.h:21-23 */
double orc_uniform2(void) { return ((double)orc_rand() + 1.0) / ((double)ORC_RAND_MAX + 2.0); } /* random.h:25-27 */
static double gs_V1, gs_V2, gs_S;
static int gs_phase = 0;
void orc_gauss_reset(void) { gs_phase = 0; }
double orc_gauss(void) { /* random.h:42-58 */
    double X;
    if (gs_phase == 0) {
        do {
            gs_V1 = 2.0 * orc_uniform2() - 1.0;
            gs_V2 = 2.0 * orc_uniform2() - 1.0;
            gs_S = gs_V1 * gs_V1 + gs_V2 * gs_V2;
        } while (gs_S >= 1.0 || gs_S == 0.0);
        X = gs_V1 * sqrt(-2.0 * log(gs_S) / gs_S);
    } else {
        X = gs_V2 * sqrt(-2.0 * log(gs_S) / gs_S);
    }
    gs_phase = 1 - gs_phase;
    return X;
}
int orc_sample_binary(double p) { return orc_uniform() < p; } /* random.h:82-84 */
void orc_init_V(float* V, size_t n, size_t factor_cnt) {      /* fm_algo_abst.h:62-65 */
    const float scale = (float)(1.0 / sqrt((double)factor_cnt));
    for (size_t i = 0; i < n; i++) V[i] = (float)(orc_gauss() * (double)scale);
}

/* ============================ common/avx.h primitives ======================================== */
float orc_dot(const float* x, const float* y, size_t len) { /* avx.h:102-127 */
    float result = 0;
    if (len > 7) {
        float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (; len > 7; len -= 8) {
            for (int l = 0; l < 8; l++) {
                float p = x[l] * y[l];
                d[l] = d[l] + p;
            }
            x += 8; y += 8;
        }
        /* hsum256_ps_avx: (hi128 + lo128) -> movehl add -> shuffle add */
        float a0 = d[4] + d[0], a1 = d[5] + d[1], a2 = d[6] + d[2], a3 = d[7] + d[3];
        float b0 = a0 + a2, b1 = a1 + a3;
        float h = b0 + b1;
        result += h;
    }
    for (; len > 0; len--) {
        float p = *x * *y;
        result += p;
        x++; y++;
    }
    return result;
}
static void v_scale(const float* x, float* res, size_t len, float s) { /* avx.h:210-225 */
    for (size_t i = 0; i < len; i++) res[i] = x[i] * s;
}
static void v_scale_v(const float* x, float* res, size_t len, const float* s) { /* avx.h:227-243 */
    for (size_t i = 0; i < len; i++) res[i] = x[i] * s[i];
}
static void v_add(const float* x, const float* y, float* res, size_t len) { /* avx.h:21-38 */
    for (size_t i = 0; i < len; i++) res[i] = x[i] + y[i];
}
static void v_scaler_add(const float* x, const float* y, float* res, float s, size_t len) { /* avx.h:58-77 */
    for (size_t i = 0; i < len; i++) { float p = y[i] * s; res[i] = x[i] + p; }
}
static void v_scaler_add_v(const float* x, const float* y, float* res, const float* s, size_t len) { /* avx.h:79-100 */
    for (size_t i = 0; i < len; i++) { float p = y[i] * s[i]; res[i] = x[i] + p; }
}

float orc_sigmoid(float x) { /* activations.h:65-72 */
    if (x < -16) return (float)1e-7;
    else if (x > 16) return (float)(1.0 - 1e-7);
    return 1.0f / (1.0f + expf(-x));
}
static void sigmoid_vec(float* a, size_t n) { /* activations.h:73-84 */
    for (size_t i = 0; i < n; i++) {
        if (a[i] < -16) a[i] = (float)1e-7;
        else if (a[i] > 16) a[i] = (float)(1.0 - 1e-7);
        else a[i] = 1.0f / (1.0f + expf(-a[i]));
    }
}
static void tanh_vec(float* a, size_t n) { /* activations.h:132-138 */
    for (size_t i = 0; i < n; i++) {
        float t1 = expf(a[i]), t2 = expf(-a[i]);
        a[i] = (t1 - t2) / (t1 + t2);
    }
}

/* ============================ loader ========================================================== */
typedef struct { char* p; size_t n, cap; } linebuf;
static int read_line(FILE* f, linebuf* lb) { /* std::getline semantics: strips '\n' only */
    lb->n = 0;
    int c, got = 0;
    while ((c = fgetc(f)) != EOF) {
        got = 1;
        if (c == '\n') break;
        if (lb->n + 2 > lb->cap) { lb->cap = lb->cap ? lb->cap * 2 : 4096; lb->p = (char*)realloc(lb->p, lb->cap); }
        lb->p[lb->n++] = (char)c;
    }
    if (lb->cap == 0) { lb->cap = 16; lb->p = (char*)malloc(16); }
    lb->p[lb->n] = 0;
    return got || c != EOF;
}
typedef struct { int64_t n, cap; void* p; size_t es; } vec;
static void vpush(vec* v, const void* e) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->p = realloc(v->p, (size_t)v->cap * v->es); }
    memcpy((char*)v->p + (size_t)v->n * v->es, e, v->es);
    v->n++;
}
static orc_data* load_impl(const char* path, uint64_t field_cnt, uint64_t feature_cnt, int test_mode,
                           uint64_t train_feature_cnt) {
    FILE* f = fopen(path, "r");
    if (!f) return NULL;
    vec rp = {0, 0, 0, sizeof(int64_t)}, fi = {0, 0, 0, sizeof(uint64_t)}, fl = {0, 0, 0, sizeof(uint64_t)},
        va = {0, 0, 0, sizeof(float)}, la = {0, 0, 0, sizeof(int)};
    linebuf lb = {0, 0, 0};
    int64_t zero = 0;
    vpush(&rp, &zero);
    int nchar = 0, y = 0;
    size_t fid = 0, fieldid = 0;
    float val = 0;
    /* while(!fin.eof()){ getline(fin,line); ... }  fm_algo_abst.h:84-105: after the last '\n' one more
     * (empty) line is read; an empty line parses nothing, so it is harmless. */
    while (read_line(f, &lb)) {
        const char* line = lb.p;
        const char* end = line + (int)lb.n;
        const char* pline = line;
        int64_t row_start = fi.n;
        if (!test_mode) {
            if (sscanf(pline, "%d%n", &y, &nchar) >= 1) { /* :88 */
                pline += nchar + 1;
                vpush(&la, &y);
                while (pline < end && sscanf(pline, "%zu:%zu:%f%n", &fieldid, &fid, &val, &nchar) >= 2) { /* :91-92 */
                    pline += nchar + 1;
                    uint64_t a = fid, b = fieldid;
                    vpush(&fi, &a); vpush(&fl, &b); vpush(&va, &val);
                    if (fid + 1 > feature_cnt) feature_cnt = fid + 1;                 /* :95 */
                    if (field_cnt > 0 && fieldid + 1 > field_cnt) field_cnt = fieldid + 1; /* :96-98 */
                }
            }
        } else { /* predict/fm_predict.cpp:108-128 */
            if (sscanf(pline, "%d%n", &y, &nchar) >= 1) {
                vpush(&la, &y);
                pline += nchar + 1;
            }
            if (sscanf(pline, "%zu:%zu:%f%n", &fieldid, &fid, &val, &nchar) >= 2) { /* first feature parsed and DROPPED */
                pline += nchar + 1;
                while (pline < end && sscanf(pline, "%zu:%zu:%f%n", &fieldid, &fid, &val, &nchar) >= 2) {
                    pline += nchar + 1;
                    if (fid < train_feature_cnt) {
                        uint64_t a = fid, b = fieldid;
                        vpush(&fi, &a); vpush(&fl, &b); vpush(&va, &val);
                    }
                }
            }
        }
        if (fi.n == row_start) continue; /* tmp.empty() -> row skipped (label already pushed!) */
        vpush(&rp, &fi.n);
        if (feof(f)) { /* getline hit EOF without trailing newline: loop ends */ }
    }
    fclose(f);
    free(lb.p);
    orc_data* d = (orc_data*)calloc(1, sizeof(orc_data));
    d->rows = rp.n - 1; d->nnz = fi.n;
    d->row_ptr = (int64_t*)rp.p; d->fid = (uint64_t*)fi.p; d->field = (uint64_t*)fl.p;
    d->val = (float*)va.p; d->label = (int*)la.p; d->label_cnt = la.n;
    d->feature_cnt = feature_cnt; d->field_cnt = field_cnt;
    return d;
}
orc_data* orc_load(const char* path, uint64_t field_cnt_in, uint64_t feature_cnt_in) {
    return load_impl(path, field_cnt_in, feature_cnt_in, 0, 0);
}
orc_data* orc_load_test(const char* path, uint64_t train_feature_cnt) {
    return load_impl(path, 0, 0, 1, train_feature_cnt);
}
void orc_free_data(orc_data* d) {
    if (!d) return;
    free(d->row_ptr); free(d->fid); free(d->field); free(d->val); free(d->label); free(d);
}

/* ============================ optimizers ====================================================== */
void orc_adagrad(size_t len, float* w, float* g, float* accum, size_t minibatch, float lr) {
    /* gradientUpdater.h:141: avx_vecScale(grad, grad, len, 1.0 / minibatch) -- scalar narrowed to float */
    const float inv = (float)(1.0 / (double)minibatch);
    v_scale(g, g, len, inv);
    for (size_t i = 0; i < len; i++) { /* :142-148 */
        const float gi = g[i];
        if (gi != 0) {
            float sq = gi * gi;
            accum[i] = accum[i] + sq;
            float num = lr * gi;
            w[i] = (float)((double)w[i] - (double)num / sqrt((double)accum[i] + 1e-7));
        }
    }
    memset(g, 0, len * sizeof(float)); /* :149 */
}
void orc_rmsprop(size_t len, float* w, float* g, float* accum, size_t minibatch, float lr, float ema) {
    /* RMSpropUpdater_Num::update, gradientUpdater.h:216-229 */
    for (size_t i = 0; i < len; i++) {
        float gi = g[i] / (float)minibatch; /* :218 float / size_t -> float */
        if (gi != 0) {
            /* :221-223 float*float, then (1.0 - ema) * g * g in double, sum in double, stored to float */
            accum[i] = (float)((double)(accum[i] * ema) + (1.0 - (double)ema) * (double)gi * (double)gi);
            float tmp = (float)(1.0 / ((double)accum[i] + 1e-7)); /* :224 */
            gi = gi * sqrtf(tmp);                                  /* :225 sqrt(float) -> std::sqrt(float) */
            w[i] = w[i] - lr * gi;                                 /* :227 */
        }
        g[i] = 0.0f; /* :229 */
    }
}
void orc_adadelta(size_t len, float* w, float* g, float* eg, float* ed, size_t minibatch, float momentum) {
    /* AdadeltaUpdater_Num::update, momentumUpdater.h:91-106 */
    for (size_t i = 0; i < len; i++) {
        float gi = g[i] / (float)minibatch; /* :93 */
        if (gi != 0) {
            eg[i] = (float)((double)(eg[i] * momentum) + (1.0 - (double)momentum) * (double)gi * (double)gi); /* :95-96 */
            float tmp = (float)(((double)ed[i] + 1e-7) / ((double)eg[i] + 1e-7));                              /* :97-98 */
            gi = gi * sqrtf(tmp);                                                                               /* :99 */
            ed[i] = (float)((double)(ed[i] * momentum) + (1.0 - (double)momentum) * (double)gi * (double)gi); /* :101-103 */
            w[i] = w[i] - gi;                                                                                   /* :105 */
        }
        g[i] = 0.0f;
    }
}
void orc_ftrl(size_t len, float* w, float* g, float* z, float* n, int zero_grad) {
    const float alpha = 0.15f, lambda1 = 1.0f, beta = 1.0f, lambda2 = 1.0f; /* gradientUpdater.h:275 */
    for (size_t i = 0; i < len; i++) { /* :254-272 */
        if (g[i] == 0) continue;
        const float g2 = g[i] * g[i];
        /* sqrt(float) -> std::sqrt(float) overload = sqrtf */
        float sigma = (sqrtf(n[i] + g2) - sqrtf(n[i])) / alpha;
        float sw = sigma * w[i];
        float d = g[i] - sw;
        z[i] = z[i] + d;
        n[i] = n[i] + g2;
        if (fabsf(z[i]) <= lambda1) { /* fabs(float) -> float overload under <cmath> */
            w[i] = 0.0f;
        } else {
            float tmpr = z[i];
            if (tmpr >= 0) tmpr -= lambda1; else tmpr += lambda1;
            float den = (beta + sqrtf(n[i])) / alpha + lambda2;
            w[i] = -tmpr / den;
        }
    }
    if (zero_grad) memset(g, 0, len * sizeof(float));
}
void orc_adam(size_t len, float* w, float* g, float* m, float* v, size_t* iter, size_t minibatch, float lr,
              float beta1, float beta2) {
    (*iter)++; /* momentumUpdater.h:191 */
    /* :192-193  pow(float, size_t) -> std::pow promotes to double; sqrt(double); result narrowed to float */
    float correction = (float)(sqrt(1 - pow((double)beta2, (double)*iter)) / (1 - pow((double)beta1, (double)*iter)));
    for (size_t i = 0; i < len; i++) { /* :195-209 */
        float gi = g[i] / (float)minibatch; /* float / size_t -> float / (float)size_t */
        if (gi != 0) {
            /* m*beta1 (float) + (1.0-beta1)*g (double) -> double -> float */
            m[i] = (float)((double)(m[i] * beta1) + (1.0 - (double)beta1) * (double)gi);
            v[i] = (float)((double)(v[i] * beta1) + (1.0 - (double)beta1) * (double)gi * (double)gi);
            float tmp = (float)((double)m[i] / ((double)sqrtf(v[i]) + 1e-7));
            w[i] = w[i] - lr * correction * tmp;
        }
        g[i] = 0.0f;
    }
}

/* ============================ FM ============================================================== */
static void loss_acc(float pred, float target, float* loss, float* acc) {
    /* train_fm_algo.cpp:93-98: target==1 ? -log(pred)[float] : -log(1.0-pred)[double]; ?: is double */
    double term = (target == 1) ? (double)(-logf(pred)) : -log(1.0 - (double)pred);
    *loss = (float)((double)*loss + term);
    if (pred > 0.5 && target == 1) *acc = *acc + 1;
    else if (pred < 0.5 && target == 0) *acc = *acc + 1;
}
void orc_fm_pass(int64_t rows, const int64_t* row_ptr, const uint32_t* fid, const float* val,
                 const int* label, size_t F, size_t k, const float* W, const float* V, float* sumVX,
                 float* update_g, float l2, float* loss, float* acc_cnt, float* pred_out) {
    float* tmp = (float*)malloc(sizeof(float) * k);
    float* update_V = update_g + F; /* train_fm_algo.h:55-57 */
    for (int64_t rid = 0; rid < rows; rid++) {
        float fm_pred = 0.0f; /* train_fm_algo.cpp:69 */
        float* srow = sumVX + (size_t)rid * k;
        for (int64_t i = row_ptr[rid]; i < row_ptr[rid + 1]; i++) {
            const size_t f = fid[i];
            const float X = val[i];
            float wx = W[f] * X;
            fm_pred += wx;                                              /* :74 */
            v_scale(V + f * k, tmp, k, X);                              /* :76 */
            v_add(srow, tmp, srow, k);                                  /* :77 */
            fm_pred = (float)((double)fm_pred - 0.5 * (double)orc_dot(tmp, tmp, k)); /* :78 */
        }
        fm_pred = (float)((double)fm_pred + 0.5 * (double)orc_dot(srow, srow, k));   /* :82 */
        const float pred = orc_sigmoid(fm_pred);                                     /* :84 */
        if (pred_out) pred_out[rid] = pred;
        /* accumWVGrad :90-118 */
        const float target = (float)label[rid];
        loss_acc(pred, target, loss, acc_cnt);
        for (int64_t i = row_ptr[rid]; i < row_ptr[rid + 1]; i++) {
            const size_t f = fid[i];
            const float x = val[i];
            float a = (pred - target) * x, b = l2 * W[f];
            const float gradW = a + b;                                  /* :108 */
            update_g[f] = update_g[f] + gradW;                          /* :109 */
            float* ptr = update_V + f * k;
            v_scaler_add(srow, V + f * k, tmp, -x, k);                  /* :112-113 */
            v_scaler_add(ptr, tmp, ptr, gradW, k);                      /* :114 */
            v_scaler_add(ptr, V + f * k, ptr, l2, k);                   /* :115 */
        }
    }
    free(tmp);
}

/* ============================ FFM ============================================================= */
void orc_ffm_pass(int64_t rows, const int64_t* row_ptr, const uint32_t* fid, const uint32_t* field,
                  const float* val, const int* label, size_t F, size_t Fc, size_t k, const float* W,
                  const float* V, float* update_g, float l2, float* loss, float* acc_cnt, float* pred_out) {
    float* uV = update_g + F; /* train_ffm_algo.h:56-59 */
    const size_t rs = Fc * k;
    for (int64_t rid = 0; rid < rows; rid++) {
        const int64_t b = row_ptr[rid], e = row_ptr[rid + 1];
        float fm_pred = 0.0f;
        for (int64_t i = b; i < e; i++) { /* train_ffm_algo.cpp:55-71 */
            const size_t f1 = fid[i], fl1 = field[i];
            const float X = val[i];
            float wx = W[f1] * X;
            fm_pred += wx;
            for (int64_t j = i + 1; j < e; j++) {
                const size_t f2 = fid[j], fl2 = field[j];
                const float X2 = val[j];
                float field_w = orc_dot(V + f1 * rs + fl2 * k, V + f2 * rs + fl1 * k, k);
                float t = field_w * X; t = t * X2;
                fm_pred += t; /* :69 */
            }
        }
        const float pred = orc_sigmoid(fm_pred);
        if (pred_out) pred_out[rid] = pred;
        /* accumWVGrad :78-118 */
        const float target = (float)label[rid];
        const float lossv = pred - target;
        if (lossv == 0) continue; /* :81-83 */
        loss_acc(pred, target, loss, acc_cnt);
        for (int64_t i = b; i < e; i++) {
            const size_t f1 = fid[i], fl1 = field[i];
            const float x = val[i];
            float a = lossv * x, c = l2 * W[f1];
            float gw = a + c;
            update_g[f1] = update_g[f1] + gw; /* :98 */
            for (int64_t j = i + 1; j < e; j++) {
                const size_t f2 = fid[j], fl2 = field[j];
                const float x2 = val[j];
                float scaler = x * x2; scaler = scaler * lossv;       /* :105 */
                const float* v1 = V + f1 * rs + fl2 * k;
                const float* v2 = V + f2 * rs + fl1 * k;
                float* u1 = uV + f1 * rs + fl2 * k;
                float* u2 = uV + f2 * rs + fl1 * k;
                v_scaler_add(u1, v2, u1, scaler, k);                    /* :111 */
                v_scaler_add(u1, v1, u1, l2, k);                        /* :112 */
                v_scaler_add(u2, v1, u2, scaler, k);                    /* :114 */
                v_scaler_add(u2, v2, u2, l2, k);                        /* :115 */
            }
        }
    }
}

/* ============================ Fully_Conn_Layer chain ========================================== */
orc_mlp* orc_mlp_create(int n_layers, const size_t* dims, int act, float sparse_rate) {
    orc_mlp* m = (orc_mlp*)calloc(1, sizeof(orc_mlp));
    m->n_layers = n_layers; m->act = act;
    for (int l = 0; l <= n_layers; l++) m->dims[l] = dims[l];
    m->input = (float*)calloc(dims[0], sizeof(float));
    for (int l = 0; l < n_layers; l++) { /* fullyconnLayer.h:36-61, layers constructed input -> output */
        size_t in = dims[l], out = dims[l + 1];
        m->weight[l] = (float*)calloc(in * out, sizeof(float));
        m->bias[l] = (float*)calloc(out, sizeof(float));
        m->mask[l] = (float*)calloc(out, sizeof(float));
        m->dW[l] = (float*)calloc(in * out, sizeof(float)); /* "memset sizes fixed" semantics */
        m->db[l] = (float*)calloc(out, sizeof(float));
        m->accum[l] = (float*)calloc(out * (in + 1), sizeof(float));
        m->out_act[l] = (float*)calloc(out, sizeof(float));
        m->in_delta[l] = (float*)calloc(in, sizeof(float));
        for (size_t i = 0; i < out; i++) { /* :48-54 */
            m->bias[l][i] = 0.0f;
            m->mask[l][i] = orc_sample_binary((double)sparse_rate) ? 1.f : 0.f;
            for (size_t j = 0; j < in; j++) m->weight[l][i * in + j] = (float)(orc_uniform() - (double)0.5f);
        }
    }
    return m;
}
void orc_mlp_free(orc_mlp* m) {
    if (!m) return;
    for (int l = 0; l < m->n_layers; l++) {
        free(m->weight[l]); free(m->bias[l]); free(m->mask[l]); free(m->dW[l]); free(m->db[l]);
        free(m->accum[l]); free(m->out_act[l]); free(m->in_delta[l]);
    }
    free(m->input); free(m);
}
float orc_mlp_forward(orc_mlp* m, const float* x) { /* fullyconnLayer.h:80-118 */
    memcpy(m->input, x, sizeof(float) * m->dims[0]); /* :91-95 */
    const float* prev = m->input;
    for (int l = 0; l < m->n_layers; l++) {
        size_t in = m->dims[l], out = m->dims[l + 1];
        int has_next = l + 1 < m->n_layers;
        float* o = m->out_act[l];
        for (size_t i = 0; i < out; i++) {
            if (has_next && !m->mask[l][i]) { o[i] = 0.0f; continue; } /* :96-99 */
            float sum = orc_dot(prev, m->weight[l] + i * in, in);       /* :100 */
            sum += m->bias[l][i];                                       /* :101 */
            o[i] = sum;
        }
        if (has_next) { /* :110-113 activation applied to ALL outputs incl. masked ones */
            if (m->act == 0) sigmoid_vec(o, out); else tanh_vec(o, out);
        }
        prev = o;
    }
    return m->out_act[m->n_layers - 1][0];
}
void orc_mlp_backward(orc_mlp* m, float out_delta) { /* fullyconnLayer.h:120-180, output layer first */
    int L = m->n_layers;
    size_t max_out = 0;
    for (int l = 0; l < L; l++) if (m->dims[l + 1] > max_out) max_out = m->dims[l + 1];
    float* delta = (float*)malloc(sizeof(float) * (max_out > 1 ? max_out : 1));
    float* tmp = (float*)malloc(sizeof(float) * max_out);
    delta[0] = out_delta;
    for (int l = L - 1; l >= 0; l--) {
        size_t in = m->dims[l], out = m->dims[l + 1];
        int has_next = l + 1 < L;
        for (size_t j = 0; j < out; j++) { /* clipping(15) :129-131, matrix.h:152-162 */
            if (delta[j] < -15.f) delta[j] = -15.f; else if (delta[j] > 15.f) delta[j] = 15.f;
        }
        /* input layer has needInputDelta=true in both users (train_nfm_algo.cpp:25, distributed_algo_abst.h:116) */
        float* idl = m->in_delta[l];
        for (size_t i = 0; i < in; i++) { /* :139-147 */
            for (size_t j = 0; j < out; j++) tmp[j] = m->weight[l][j * in + i];
            if (has_next) v_scale_v(tmp, tmp, out, m->mask[l]);
            idl[i] = orc_dot(tmp, delta, out);
        }
        const float* x = (l == 0) ? m->input : m->out_act[l - 1];
        /* weight/bias grads :165-179 (independent of the recursion order) */
        for (size_t j = 0; j < out; j++) v_scaler_add(m->dW[l] + j * in, x, m->dW[l] + j * in, delta[j], in);
        v_add(m->db[l], delta, m->db[l], out);
        if (l > 0) { /* prev layer's activation backward :153-156 */
            const float* fo = m->out_act[l - 1];
            for (size_t i = 0; i < in; i++) {
                if (m->act == 0) { float t = idl[i] * fo[i]; delta[i] = t * (1.0f - fo[i]); }   /* activations.h:85-90 */
                else { float t = fo[i] * fo[i]; delta[i] = idl[i] * (1.0f - t); }               /* activations.h:139-143 */
            }
        }
    }
    free(delta); free(tmp);
}
void orc_mlp_apply(orc_mlp* m, size_t minibatch, float lr, float sparse_rate) { /* fullyconnLayer.h:194-206 */
    for (int l = 0; l < m->n_layers; l++) {
        size_t in = m->dims[l], out = m->dims[l + 1];
        orc_adagrad(out, m->bias[l], m->db[l], m->accum[l], minibatch, lr);
        orc_adagrad(out * in, m->weight[l], m->dW[l], m->accum[l] + out, minibatch, lr);
        for (size_t i = 0; i < out; i++) m->mask[l][i] = orc_sample_binary((double)sparse_rate) ? 1.f : 0.f;
    }
}

/* ============================ NFM ============================================================= */
void orc_nfm_epoch(int64_t rows, const int64_t* row_ptr, const uint32_t* fid, const float* val,
                   const int* label, size_t F, size_t k, float* W, float* V, float* sumVX, float* update_g,
                   float* accum, orc_mlp* mlp, size_t batch_size, size_t minibatch, float lr, float l2,
                   float sparse_rate, float* loss_out, size_t* acc_out) {
    float loss = 0; size_t accuracy = 0; /* train_nfm_algo.cpp:36-37 */
    memset(sumVX, 0, sizeof(float) * (size_t)rows * k); /* :38 */
    float* tmp = (float*)malloc(sizeof(float) * k);
    float* tmp2 = (float*)malloc(sizeof(float) * k);
    float* z = (float*)malloc(sizeof(float) * k);
    float* uV = update_g + F;
    size_t n_batches = ((size_t)rows + batch_size - 1) / batch_size;
    for (size_t p = 0; p < n_batches; p++) {
        memset(update_g, 0, sizeof(float) * F * (k + 1)); /* :43 */
        int64_t rb = (int64_t)(p * batch_size), re = rb + (int64_t)batch_size;
        if (re > rows) re = rows;
        for (int64_t rid = rb; rid < re; rid++) { /* batchGradCompute :56-124 */
            float* srow = sumVX + (size_t)rid * k;
            memset(z, 0, sizeof(float) * k);
            float fm_pred = 0.0f;
            for (int64_t i = row_ptr[rid]; i < row_ptr[rid + 1]; i++) {
                const size_t f = fid[i];
                const float X = val[i];
                float wx = W[f] * X;
                fm_pred += wx;                              /* :83 */
                v_scale(V + f * k, tmp, k, X);              /* :85 */
                v_add(srow, tmp, srow, k);                  /* :86 */
                v_scale(tmp, tmp2, k, (float)-0.5);         /* :87 */
                v_scaler_add_v(z, tmp, z, tmp2, k);         /* :88-91 */
            }
            v_scale(srow, tmp, k, (float)0.5);              /* :93 */
            v_scaler_add_v(z, srow, z, tmp, k);             /* :94 */
            float fc = orc_mlp_forward(mlp, z);             /* :98 */
            fm_pred += fc;                                  /* :101 */
            fm_pred = orc_sigmoid(fm_pred);                 /* :102 */
            {   /* :104-109 */
                double term = (label[rid] == 1) ? (double)(-logf(fm_pred)) : -log(1.0 - (double)fm_pred);
                loss = (float)((double)loss + term);
                if (fm_pred > 0.5 && label[rid] == 1) accuracy++;
                else if (fm_pred < 0.5 && label[rid] == 0) accuracy++;
            }
            const float target = (float)label[rid];
            for (int64_t i = row_ptr[rid]; i < row_ptr[rid + 1]; i++) { /* accumWideGrad :126-137 */
                const size_t f = fid[i];
                const float x = val[i];
                float a = (fm_pred - target) * x, b = l2 * W[f];
                float g = a + b;
                update_g[f] = update_g[f] + g;
            }
            orc_mlp_backward(mlp, fm_pred - (float)label[rid]); /* :115-117 */
            const float* delta = mlp->in_delta[0];
            for (int64_t i = row_ptr[rid]; i < row_ptr[rid + 1]; i++) { /* accumDeepGrad :139-159 */
                const size_t f = fid[i];
                const float X = val[i];
                v_scaler_add(srow, V + f * k, tmp, -X, k);                    /* :152-153 */
                v_scale(delta, tmp2, k, X);                                   /* :154 */
                v_scaler_add_v(uV + f * k, tmp, uV + f * k, tmp2, k);         /* :155-156 */
                v_scaler_add(uV + f * k, V + f * k, uV + f * k, l2, k);       /* :157 */
            }
        }
        /* ApplyGrad :161-169 */
        orc_adagrad(F, W, update_g, accum, minibatch, lr);
        orc_adagrad(F * k, V, uV, accum + F, minibatch, lr);
        orc_mlp_apply(mlp, minibatch, lr, sparse_rate);
    }
    free(tmp); free(tmp2); free(z);
    *loss_out = loss; *acc_out = accuracy;
}


/* ============================ Wide&Deep with per-field concat input ============================ */
/* Distributed_Algo_Abst::batchGradCompute (distributed_algo_abst.h:176-280) restated as ONE synchronous process:
 * what a worker computes per row, with the pull = "read the current parameters" and the push = "add into update_g",
 * applied once per minibatch by the trainer's updater (the reference applies pushes on asynchronous parameter
 * servers with fp16 transport; SURVEY.md 8e lists those deltas).  This SCHEDULE is the product's, not the reference's; every
 * function it is made of is PINNED through orc_wnd_epoch_ref below -- the same worker and server rules in the cluster's own
 * schedule, held digit for digit against a real Master + ParamServer + worker run (tests/test_oracle_wnd_pin_cpu.py).
 *   wide:  pred += w[fid] * X over the row's entries (:205-211)
 *   deep:  input[field*d .. +d) = tensor of the FIRST entry of each field (:213-216, :224-229), 0 for absent fields
 *   pCTR = sigmoid(pred + MLP(input)) (:236); loss / accuracy (:237-245, note >= 0.5)
 *   gradW = loss * X + L2 * w (:256), pushed per entry; MLP backward from `loss` (:270-273); inputDelta is the
 *   gradient of the first-entry tensors (:274-276); MLP applyBatchGradient after the batch (:283). */
void orc_ps_update(int kind, size_t len, float* w, const float* g, float* accum, float* shadow, size_t minibatch, float lr, int tensor);
static int orc_wnd_ps_rule = 0;
void orc_set_wnd_ps_rule(int on) { orc_wnd_ps_rule = on; }
void orc_wnd_epoch(int64_t rows, const int64_t* row_ptr, const uint32_t* fid, const uint32_t* field,
                   const float* val, const int* label, size_t F, size_t Fc, size_t d, float* W, float* E,
                   float* update_g, float* accum, orc_mlp* mlp, size_t batch_size, size_t minibatch, float lr,
                   float l2, float sparse_rate, float* loss_out, size_t* acc_out) {
    float loss = 0; size_t accuracy = 0;
    float* deep = (float*)malloc(sizeof(float) * Fc * d);
    int64_t* first = (int64_t*)malloc(sizeof(int64_t) * Fc);
    float* gE = update_g + F;
    size_t n_batches = ((size_t)rows + batch_size - 1) / batch_size;
    for (size_t p = 0; p < n_batches; p++) {
        memset(update_g, 0, sizeof(float) * F * (d + 1));
        int64_t rb = (int64_t)(p * batch_size), re = rb + (int64_t)batch_size;
        if (re > rows) re = rows;
        for (int64_t rid = rb; rid < re; rid++) {
            float pred = 0.0f;
            memset(deep, 0, sizeof(float) * Fc * d);
            for (size_t a = 0; a < Fc; a++) first[a] = -1;
            for (int64_t i = row_ptr[rid]; i < row_ptr[rid + 1]; i++) {
                const size_t f = fid[i], a = field[i];
                float wx = W[f] * val[i];
                pred += wx;
                if (first[a] < 0) { first[a] = (int64_t)f; memcpy(deep + a * d, E + f * d, sizeof(float) * d); }
            }
            const float fc = orc_mlp_forward(mlp, deep);
            const float pCTR = orc_sigmoid(pred + fc);
            {
                double term = (label[rid] == 1) ? (double)(-logf(pCTR)) : -log(1.0 - (double)pCTR);
                loss = (float)((double)loss + term);
                if (pCTR >= 0.5 && label[rid] == 1) accuracy++;
                else if (pCTR < 0.5 && label[rid] == 0) accuracy++;
            }
            const float lossv = pCTR - (float)label[rid];
            for (int64_t i = row_ptr[rid]; i < row_ptr[rid + 1]; i++) {
                const size_t f = fid[i];
                float a = lossv * val[i], c = l2 * W[f];
                update_g[f] = update_g[f] + (a + c);
            }
            orc_mlp_backward(mlp, lossv);
            const float* delta = mlp->in_delta[0];
            for (size_t a = 0; a < Fc; a++) {
                if (first[a] < 0) continue;
                for (size_t c = 0; c < d; c++) gE[(size_t)first[a] * d + c] = gE[(size_t)first[a] * d + c] + delta[a * d + c];
            }
        }
        if (orc_wnd_ps_rule) { /* the parameter server's default rules: scalar SGD for w (paramserver.h:295-300), tensor SGD for
                                  the tensors (:232-237); a zero gradient leaves a coordinate unchanged, so the dense sweep
                                  equals the PS touching the pushed keys only */
            orc_ps_update(0, F, W, update_g, accum, accum, minibatch, lr, 0);
            orc_ps_update(0, F * d, E, gE, accum, accum, minibatch, lr, 1);
        } else {
            orc_adagrad(F, W, update_g, accum, minibatch, lr);
            orc_adagrad(F * d, E, gE, accum + F, minibatch, lr);
        }
        orc_mlp_apply(mlp, minibatch, lr, sparse_rate);
    }
    free(deep); free(first);
    *loss_out = loss; *acc_out = accuracy;
}

/* ---- the worker + parameter server in the REFERENCE'S OWN schedule (pinned) ------------------------------------------------
 * orc_wnd_epoch above is the synchronous batch restatement the CUDA path implements.  What the unmodified cluster does
 * differs in three places, all restated here so that this function can be held against the loss curve of a real
 * Master + ParamServer + worker run over ZeroMQ (tests/golden/wnd_ref_curve.json, made by tests/golden/make_wnd_ref_curve.py
 * from oracle/ref_dist_driver.cpp):
 *   1. every number that crosses the wire is a binary16: pulled parameters (distribut/paramserver.h:160-162,177, read back by
 *      pull.h:111-112,150), pushed gradients (push.h, read by paramserver.h:217-220,241) -- common/float16.h:105-152 rounds to
 *      nearest even, flushes nothing, maps -0 to +0;
 *   2. the wide weights are pulled ONCE per minibatch (distributed_algo_abst.h:178-193) and their gradients pushed once
 *      (:275-278), but a sample's per-field tensors are pulled (:217) and their gradients pushed (:272) PER SAMPLE: the server
 *      applies plain SGD to a tensor right away (paramserver.h:229-237), so the next sample sees it;
 *   3. a tensor is created on the server the first time it is pulled, from the SERVER's rand() stream (paramserver.h:40-46):
 *      the caller passes E already initialised in that order (oracle/umap_order.cpp).
 * Wide parameters start at 0 (Value::initParam, distributed_algo_abst.h:70-72), L2Reg_ratio = 0 (:103). */
static uint16_t orc_f32_to_f16(float src) {            /* float16.h:105-152 */
    uint32_t s; memcpy(&s, &src, 4);
    const uint16_t sign = (uint16_t)((s >> 16) & 0x8000);
    int exp = (int)((s >> 23) & 0xff) - 127;
    int mant = (int)(s & 0x7fffff);
    if ((s & 0x7fffffff) == 0) return 0;               /* +-0 -> +0 */
    if (exp > 15) return (exp == 128 && mant) ? 0x7fff : (uint16_t)(sign | 0x7c00);
    uint16_t u = 0; int sticky = 0;
    if (exp >= -14) { u = (uint16_t)((((exp + 15) & 0x1f) << 10) | (mant >> 13)); }
    else {
        const int rshift = -(exp + 14);
        if (rshift < 32) { mant |= 1 << 23; sticky = (mant & ((1 << rshift) - 1)) != 0; mant >>= rshift; u = (uint16_t)((mant >> 13) & 0x3ff); }
        else { mant = 0; u = 0; }
    }
    const int round_bit = (mant >> 12) & 1;
    sticky |= (mant & ((1 << 12) - 1)) != 0;
    if ((round_bit && sticky) || (round_bit && (u & 1))) u = (uint16_t)(u + 1);
    return (uint16_t)(u | sign);
}
static float orc_f16_to_f32(uint16_t h) {              /* float16.h:65-100 */
    const uint32_t sign = (h >> 15) & 1; int exp = (h >> 10) & 0x1f; uint32_t mant = h & 0x3ff, f = 0;
    if (exp > 0 && exp < 31) f = (sign << 31) | ((uint32_t)(exp + 112) << 23) | (mant << 13);
    else if (exp == 0) {
        if (mant) { exp += 113; while ((mant & (1u << 10)) == 0) { mant <<= 1; exp--; } mant &= 0x3ff; f = (sign << 31) | ((uint32_t)exp << 23) | (mant << 13); }
        else f = sign << 31;
    } else f = mant ? 0x7fffffffu : ((0xffu << 23) | (sign << 31));
    float r; memcpy(&r, &f, 4); return r;
}
float orc_wire_f16(float x) { return orc_f16_to_f32(orc_f32_to_f16(x)); }

void orc_wnd_epoch_ref(int64_t rows, const int64_t* row_ptr, const uint32_t* fid, const uint32_t* field, const float* val,
                       const int* label, size_t F, size_t Fc, size_t d, float* W, float* E, float* push_g, float* pulled,
                       orc_mlp* mlp, size_t batch_size, size_t minibatch, float lr, float sparse_rate, int ps_kind, float* accum,
                       float* shadow, float* loss_out, size_t* acc_out) {
    float loss = 0; size_t accuracy = 0;
    float* deep = (float*)malloc(sizeof(float) * Fc * d);
    int64_t* first = (int64_t*)malloc(sizeof(int64_t) * Fc);
    const float scaler = (float)(-1.0 * (double)lr / (double)minibatch);       /* paramserver.h:232-233 */
    const float den = (float)minibatch / lr;                                     /* :297-299 */
    size_t n_batches = ((size_t)rows + batch_size - 1) / batch_size;
    for (size_t p = 0; p < n_batches; p++) {
        int64_t rb = (int64_t)(p * batch_size), re = rb + (int64_t)batch_size;
        if (re > rows) re = rows;
        /* pull_op.sync: one binary16 copy of every wide weight of the minibatch (distributed_algo_abst.h:178-193) */
        for (int64_t i = row_ptr[rb]; i < row_ptr[re]; i++) { pulled[fid[i]] = orc_wire_f16(W[fid[i]]); push_g[fid[i]] = 0.0f; }
        for (int64_t rid = rb; rid < re; rid++) {
            float pred = 0.0f;
            memset(deep, 0, sizeof(float) * Fc * d);
            for (size_t a = 0; a < Fc; a++) first[a] = -1;
            for (int64_t i = row_ptr[rid]; i < row_ptr[rid + 1]; i++) {         /* :201-212 */
                const size_t f = fid[i], a = field[i];
                float wx = pulled[f] * val[i];
                pred += wx;
                if (first[a] < 0) {                                              /* pull_tensor_op.sync (:217), binary16 */
                    first[a] = (int64_t)f;
                    for (size_t c = 0; c < d; c++) deep[a * d + c] = orc_wire_f16(E[f * d + c]);
                }
            }
            const float fc = orc_mlp_forward(mlp, deep);
            const float pCTR = orc_sigmoid(pred + fc);
            {
                double term = (label[rid] == 1) ? (double)(-logf(pCTR)) : -log(1.0 - (double)pCTR);
                loss = (float)((double)loss + term);
                if (pCTR >= 0.5 && label[rid] == 1) accuracy++;
                else if (pCTR < 0.5 && label[rid] == 0) accuracy++;
            }
            if (minibatch == 0) continue;                                        /* predicting */
            const float lossv = pCTR - (float)label[rid];
            for (int64_t i = row_ptr[rid]; i < row_ptr[rid + 1]; i++) {         /* :250-266, L2Reg_ratio == 0 */
                const size_t f = fid[i];
                float a = lossv * val[i], c = 0.0f * pulled[f];
                push_g[f] = push_g[f] + (a + c);
            }
            orc_mlp_backward(mlp, lossv);
            const float* delta = mlp->in_delta[0];
            for (size_t a = 0; a < Fc; a++) {                                    /* push_tensor_op.sync (:272) -> SGD on the server */
                if (first[a] < 0) continue;
                for (size_t c = 0; c < d; c++) {
                    float t = orc_wire_f16(delta[a * d + c]) * scaler;
                    E[(size_t)first[a] * d + c] = E[(size_t)first[a] * d + c] + t;
                }
            }
        }
        if (minibatch == 0) continue;
        /* push_op.sync (:276-278): binary16 gradient sums, scalar SGD on the server (paramserver.h:295-300).  A key occurs once
         * in push_map; walk the batch's entries and clear each sum after use. */
        for (int64_t i = row_ptr[rb]; i < row_ptr[re]; i++) {
            const size_t f = fid[i];
            if (pulled[f] != pulled[f]) continue;                                /* already applied (marked NaN below) */
            pulled[f] = NAN;
            /* push.h:62-65: a gradient outside 1e-7 < |g| < 15 (Value::checkPreferredValue, distributed_algo_abst.h:76-79) is
             * not sent at all -- a feature present in most rows of the minibatch easily sums to |g| >= 15 */
            if (!((double)fabsf(push_g[f]) > 1e-7 && fabsf(push_g[f]) < 15.0f)) continue;
            const float g16 = orc_wire_f16(push_g[f]);
            if (ps_kind == 0) { float t = g16 / den; W[f] = W[f] - t; }                /* SGD, the server's default */
            else orc_ps_update(ps_kind, 1, W + f, &g16, accum + f, shadow + f, minibatch, lr, 0);  /* Adagrad / DCASGD / DCASGDA */
        }
        orc_mlp_apply(mlp, minibatch, lr, sparse_rate);                          /* :280 */
    }
    free(deep); free(first);
    *loss_out = loss; *acc_out = accuracy;
}

/* ============================ predict + AUC =================================================== */
float orc_auc(const float* pctr, const int* label, size_t n) { /* evaluator.h:61-93 */
    const size_t kHashLen = (1u << 24) - 1;
    int* pos = (int*)calloc(kHashLen + 1, sizeof(int));
    int* neg = (int*)calloc(kHashLen + 1, sizeof(int));
    for (size_t i = 0; i < n; i++) {
        size_t index = (size_t)(pctr[i] * (float)kHashLen); /* float * size_t -> float */
        if (label[i] == 1) pos[index]++; else neg[index]++;
    }
    float totPos = 0, totNeg = 0, totPosPrev = 0, totNegPrev = 0, auc = 0;
    for (int64_t idx = (int64_t)kHashLen; idx >= 0; idx--) {
        totPosPrev = totPos; totNegPrev = totNeg;
        totPos += (float)pos[idx]; totNeg += (float)neg[idx];
        float dx = totNeg > totNegPrev ? (totNeg - totNegPrev) : (totNegPrev - totNeg);
        /* trapezoidArea: (..)*(Y1+Y2)/2.0 -> double, returned as float */
        float area = (float)((double)(dx * (totPos + totPosPrev)) / 2.0);
        auc += area;
    }
    free(pos); free(neg);
    if (totPos > 0.0 && totNeg > 0.0) return auc / totPos / totNeg;
    return 0.0f;
}
void orc_predict(int64_t rows, const int64_t* row_ptr, const uint32_t* fid, const uint32_t* field,
                 const float* val, const int* label, size_t Fc, size_t k, const float* W, const float* V,
                 const float* train_sumVX, int is_ffm, float* pctr_out, float* loss_out, int* correct_out,
                 float* auc_out) {
    float* tmp = (float*)malloc(sizeof(float) * k);
    const size_t rs = Fc * k;
    for (int64_t rid = 0; rid < rows; rid++) {
        float fm_pred = 0.0f;
        const int64_t b = row_ptr[rid], e = row_ptr[rid + 1];
        if (!is_ffm) { /* fm_predict.cpp:20-33 */
            for (int64_t i = b; i < e; i++) {
                const size_t f = fid[i];
                const float X = val[i];
                float wx = W[f] * X;
                fm_pred += wx;
                v_scale(V + f * k, tmp, k, X);
                fm_pred = (float)((double)fm_pred - 0.5 * (double)orc_dot(tmp, tmp, k));
            }
            /* quirk: uses the TRAINING sumVX of row `rid` (:31) */
            const float* s = train_sumVX + (size_t)rid * k;
            fm_pred = (float)((double)fm_pred + 0.5 * (double)orc_dot(s, s, k));
        } else { /* :34-53 */
            for (int64_t i = b; i < e; i++) {
                const size_t f1 = fid[i], fl1 = field[i];
                const float X = val[i];
                float wx = W[f1] * X;
                fm_pred += wx;
                for (int64_t j = i + 1; j < e; j++) {
                    const size_t f2 = fid[j], fl2 = field[j];
                    float fw = orc_dot(V + f1 * rs + fl2 * k, V + f2 * rs + fl1 * k, k);
                    float t = fw * X; t = t * val[j];
                    fm_pred += t;
                }
            }
        }
        pctr_out[rid] = orc_sigmoid(fm_pred);
    }
    free(tmp);
    float loss = 0; int correct = 0; /* :63-72 */
    for (int64_t i = 0; i < rows; i++) {
        double term = (label[i] == 1) ? (double)(-logf(pctr_out[i])) : -log(1.0 - (double)pctr_out[i]);
        loss = (float)((double)loss + term);
        if (pctr_out[i] > 0.5 && label[i] == 1) correct++;
        else if (pctr_out[i] < 0.5 && label[i] == 0) correct++;
    }
    *loss_out = loss; *correct_out = correct;
    *auc_out = orc_auc(pctr_out, label, (size_t)rows);
}

/* ParamServer push handler, per coordinate (distribut/paramserver.h:232-300), one push of worker 0 carrying the step's
 * summed gradient.  PINNED: all four rules reproduce the loss curves of the reference cluster run with the corresponding
 * UpdaterType (tests/golden/wnd_ref_curve.json, tests/test_oracle_wnd_pin_cpu.py; the pin caught that Adagrad steps by
 * g / minibatch).  The Value operators mutate their left operand (distributed_algo_abst.h:39-72), which the sequences below
 * follow literally.
 * kind: 0 SGD (:295-300; tensor != 0: tensor SGD :232-237), 1 Adagrad (:288-294), 2 DCASGD (:252-267), 3 DCASGDA (:268-286).
 * accum starts at 1e-7 (:323), shadow at 0 (:327). */
void orc_ps_update(int kind, size_t len, float* w, const float* g, float* accum, float* shadow, size_t minibatch, float lr,
                   int tensor) {
    const float mb = (float)minibatch;
    for (size_t i = 0; i < len; i++) {
        if (kind == 0) {
            if (tensor) { const float scaler = (float)(-1.0 * (double)lr / (double)minibatch); float t = g[i] * scaler; w[i] = w[i] + t; }
            else { float den = mb / lr; float t = g[i] / den; w[i] = w[i] - t; }
        } else if (kind == 1) {
            /* :288-294.  `TValue grad = data_pair.second / minibatch` divides the pushed value IN PLACE (Value::operator/ mutates
             * and returns *this), so the step is (g / minibatch) / (sqrt(accum) / lr) -- pinned: wnd_ref_curve.json "adagrad" */
            const float gm = g[i] / mb; float grad = gm * gm; accum[i] = accum[i] + grad;
            float sq = (float)sqrt((double)accum[i] + 1e-7); sq = sq / lr;
            float t = gm / sq; w[i] = w[i] - t;
        } else if (kind == 2) {
            float grad = g[i] / mb, reserve = grad;
            grad = grad * grad; float cur = w[i] - shadow[i]; grad = grad * cur; grad = grad * 0.1f;
            reserve = reserve + grad; reserve = reserve * lr; w[i] = w[i] - reserve; shadow[i] = w[i];
        } else {
            float grad = g[i] / mb;
            accum[i] = accum[i] * 0.95f; grad = grad * grad; grad = grad * (1 - 0.95f); accum[i] = accum[i] + grad;
            float reserve = grad; /* :277 copies grad after it was overwritten */
            float sq = (float)sqrt((double)accum[i] + 1e-7);
            grad = grad * grad; float cur = w[i] - shadow[i]; grad = grad * cur; grad = grad * 0.1f; grad = grad / sq;
            reserve = reserve + grad; reserve = reserve * lr; w[i] = w[i] - reserve; shadow[i] = w[i];
        }
    }
}

/* ============================ distributed semantics =========================================== */
uint32_t orc_murmur_u64(uint64_t k) { /* hash.h:51-58 */
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (uint32_t)k;
}
uint32_t orc_murmur_str(const char* key, int len) { /* hash.h:16-49 (MurmurHash2, seed 97) */
    const uint32_t m = 0x5bd1e995u; const int r = 24;
    uint32_t h = 97u ^ (uint32_t)len;
    const unsigned char* data = (const unsigned char*)key;
    while (len >= 4) {
        uint32_t k; memcpy(&k, data, 4);
        k *= m; k ^= k >> r; k *= m; h *= m; h ^= k;
        data += 4; len -= 4;
    }
    switch (len) {
        case 3: h ^= (uint32_t)data[2] << 16; /* fall through */
        case 2: h ^= (uint32_t)data[1] << 8;  /* fall through */
        case 1: h ^= data[0]; h *= m;
    }
    h ^= h >> 13; h *= m; h ^= h >> 15;
    return h;
}
/* ConsistentHash::getNode (distribut/consistent_hash.h:29-40,51-60): the ring holds 5 virtual nodes per server at
 * murMurHash("<server>-<replica>") (a later insert at the same position overwrites, std::map::operator[]); a key goes to the
 * first virtual node at or after murMurHash(key), wrapping to the ring's first.  Pinned: tests/golden/dht_nodes.json comes
 * from the reference class itself (oracle/dht_nodes.cpp). */
uint32_t orc_dht_node(uint64_t key, uint32_t node_cnt) {
    uint32_t pos[5 * 64], srv[5 * 64]; int n = 0;
    if (node_cnt > 64) node_cnt = 64;
    for (uint32_t i = 0; i < node_cnt; i++)
        for (uint32_t j = 0; j < 5; j++) {
            char buf[32];
            int len = snprintf(buf, sizeof(buf), "%u-%u", i, j);
            uint32_t p = orc_murmur_str(buf, len);
            int at = -1;
            for (int t = 0; t < n; t++) if (pos[t] == p) at = t;
            if (at >= 0) srv[at] = i; else { pos[n] = p; srv[n] = i; n++; }
        }
    const uint32_t part = orc_murmur_u64(key);
    int best = -1, first = 0;
    for (int t = 0; t < n; t++) {
        if (pos[t] < pos[first]) first = t;
        if (pos[t] >= part && (best < 0 || pos[t] < pos[best])) best = t;
    }
    return srv[best >= 0 ? best : first];
}
void orc_ring_segments(size_t P, size_t R, size_t* seg_size, size_t* seg_end) { /* ring_collect.h:86-109 */
    size_t s = P / R, res = P % R;
    for (size_t i = 0; i < R; i++) {
        seg_size[i] = s + (i < res ? 1 : 0);
        seg_end[i] = (i == 0 ? 0 : seg_end[i - 1]) + seg_size[i];
    }
}
void orc_ring_allreduce(float** bufs, size_t R, size_t P, int do_average) {
    /* reduce_step ring_collect.h:114-165: at step i rank r sends segment (r - i) mod R to rank r+1,
     * which adds it into its own copy (avx_vecAdd(buffer, begin, begin): received + local).
     * PINNED bit for bit against real runs of the reference's ring master + R worker processes
     * (oracle/ref_ring_driver.cpp, tests/golden/ring_allreduce.json, tests/test_oracle_ring_cpu.py). */
    size_t* ss = (size_t*)malloc(sizeof(size_t) * R), *se = (size_t*)malloc(sizeof(size_t) * R);
    orc_ring_segments(P, R, ss, se);
    float* wire = (float*)malloc(sizeof(float) * (P / R + 1) * R);
    for (size_t i = 0; i + 1 < R; i++) {
        for (size_t r = 0; r < R; r++) { /* snapshot what every rank sends this step */
            size_t seg = (r + R - i) % R;
            memcpy(wire + r * (P / R + 1), bufs[r] + (se[seg] - ss[seg]), sizeof(float) * ss[seg]);
        }
        for (size_t r = 0; r < R; r++) {
            size_t src = (r + R - 1) % R;
            size_t seg = (r + R - i - 1) % R;
            float* dst = bufs[r] + (se[seg] - ss[seg]);
            const float* w = wire + src * (P / R + 1);
            for (size_t t = 0; t < ss[seg]; t++) dst[t] = w[t] + dst[t];
        }
    }
    /* gather_step :167-218: segment (r + 1 - i) mod R travels r -> r+1 */
    for (size_t i = 0; i + 1 < R; i++) {
        for (size_t r = 0; r < R; r++) {
            size_t seg = (r + 1 + R - i) % R;
            memcpy(wire + r * (P / R + 1), bufs[r] + (se[seg] - ss[seg]), sizeof(float) * ss[seg]);
        }
        for (size_t r = 0; r < R; r++) {
            size_t src = (r + R - 1) % R;
            size_t seg = (r + R - i) % R;
            memcpy(bufs[r] + (se[seg] - ss[seg]), wire + src * (P / R + 1), sizeof(float) * ss[seg]);
        }
    }
    if (do_average) { /* :61-67 */
        const float scalar = (float)(1.0 / (double)R);
        for (size_t r = 0; r < R; r++) v_scale(bufs[r], bufs[r], P, scalar);
    }
    free(ss); free(se); free(wire);
}
