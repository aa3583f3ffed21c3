// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" driver around the UNMODIFIED reference trainers, compiled by
// oracle/Makefile together with the reference's own .cpp files *where they lie*
// under /root/reference (nothing is copied).  Output: oracle/_ref/libref.so.
// Used (a) to pin the C restatement in oracle/lightctr_oracle.c bit-for-bit,
// (b) to generate tests/golden/*, (c) as bench.py's `--impl reference` CPU arm.
//
// Deviations from running the reference's main.cpp, all non-arithmetic except (3):
//  (1) srand(seed) is fixed instead of srand(time(NULL))           (main.cpp:78)
//  (2) proc_cnt may be forced to 1 for a strictly sequential row order
//      (fm_algo_abst.h:42,142; SURVEY.md 8c "canonical mode")
//  (3) global operator new[] returns zeroed memory.  This makes two reference
//      bugs deterministic without editing its sources: Fully_Conn_Layer::init's
//      memset byte counts miss sizeof(float) (train/layer/fullyconnLayer.h:57-60)
//      and Train_FFM_Algo::init never zeroes update_g (train/train_ffm_algo.cpp:17).
//      Zeroed fresh arrays == "reference + memset sizes fixed" (SURVEY.md 8c).
//  (4) `private`/`protected` are made public for this TU so per-epoch loss
//      (fm_algo_abst.h:167) and the NFM's FC weights can be read back.
#include <algorithm>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <list>
#include <memory>
#include <mutex>
#include <new>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <unistd.h>
#include <fcntl.h>

void* operator new[](std::size_t n) {
    void* p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

#define private public
#define protected public
#include "LightCTR/fm_algo_abst.h"
#include "LightCTR/train/train_fm_algo.h"
#include "LightCTR/train/train_ffm_algo.h"
#include "LightCTR/train/train_nfm_algo.h"
#include "LightCTR/predict/fm_predict.h"
#include "LightCTR/util/gradientUpdater.h"
#include "LightCTR/util/momentumUpdater.h"
#undef private
#undef protected

// main.cpp:64-73 (same defaults)
size_t GradientUpdater::__global_minibatch_size(50);
float GradientUpdater::__global_learning_rate(0.05);
float GradientUpdater::__global_ema_rate(0.99);
float GradientUpdater::__global_sparse_rate(0.8);
float GradientUpdater::__global_lambdaL2(0.001f);
float GradientUpdater::__global_lambdaL1(1e-5);
float MomentumUpdater::__global_momentum(0.8);
float MomentumUpdater::__global_momentum_adam2(0.999);
bool GradientUpdater::__global_bTraining(true);

namespace {
enum Kind { K_FM = 1, K_FFM = 2, K_NFM = 3 };
struct Handle {
    Kind kind;
    FM_Algo_Abst* algo;
};
int g_quiet_fd = -1;
struct Quiet {  // silence the reference's printf/cout chatter
    int saved;
    explicit Quiet(bool on) : saved(-1) {
        if (!on) return;
        fflush(stdout); std::cout.flush();
        saved = dup(1);
        if (g_quiet_fd < 0) g_quiet_fd = open("/dev/null", O_WRONLY);
        dup2(g_quiet_fd, 1);
    }
    ~Quiet() {
        if (saved < 0) return;
        fflush(stdout); std::cout.flush();
        dup2(saved, 1); close(saved);
    }
};
size_t v_size(const FM_Algo_Abst* a) {
    size_t m = a->feature_cnt * a->factor_cnt;
    if (a->field_cnt > 0) m = a->feature_cnt * a->field_cnt * a->factor_cnt;
    return m;
}
void realign_gauss(const FM_Algo_Abst* a) {
    // GaussRand (util/random.h:42-58) caches one deviate in function-local statics;
    // consume the cached one so the next create() starts a fresh pair.
    if (v_size(a) & 1) (void)GaussRand();
}
}  // namespace

extern "C" {

void ref_set_hyper(size_t minibatch, float lr, float ema, float sparse_rate, float l2, float l1,
                   float momentum, float adam2) {
    GradientUpdater::__global_minibatch_size = minibatch;
    GradientUpdater::__global_learning_rate = lr;
    GradientUpdater::__global_ema_rate = ema;
    GradientUpdater::__global_sparse_rate = sparse_rate;
    GradientUpdater::__global_lambdaL2 = l2;
    GradientUpdater::__global_lambdaL1 = l1;
    MomentumUpdater::__global_momentum = momentum;
    MomentumUpdater::__global_momentum_adam2 = adam2;
}

// proc_cnt <= 0 keeps the reference default hardware_concurrency().
void* ref_fm_create(const char* path, int k, unsigned seed, int proc_cnt) {
    Quiet q(true);
    srand(seed);
    Train_FM_Algo* t = new Train_FM_Algo(path, /*epoch*/ 1, (size_t)k);
    if (proc_cnt > 0) t->proc_cnt = (size_t)proc_cnt;
    realign_gauss(t);
    return new Handle{K_FM, t};
}
void* ref_ffm_create(const char* path, int k, int field_cnt, unsigned seed, int proc_cnt) {
    Quiet q(true);
    srand(seed);
    Train_FFM_Algo* t = new Train_FFM_Algo(path, 1, (size_t)k, (size_t)field_cnt);
    if (proc_cnt > 0) t->proc_cnt = (size_t)proc_cnt;
    realign_gauss(t);
    return new Handle{K_FFM, t};
}
void* ref_nfm_create(const char* path, int k, int hidden, unsigned seed) {
    Quiet q(true);
    srand(seed);
    Train_NFM_Algo* t = new Train_NFM_Algo(path, 1, (size_t)k, (size_t)hidden);
    realign_gauss(t);
    return new Handle{K_NFM, t};
}
void ref_destroy(void* hp) {
    Handle* h = (Handle*)hp;
    Quiet q(true);
    delete h->algo;
    delete h;
}
void ref_dims(void* hp, size_t* rows, size_t* nnz, size_t* feature_cnt, size_t* field_cnt,
              size_t* factor_cnt) {
    Handle* h = (Handle*)hp;
    *rows = h->algo->dataRow_cnt;
    size_t n = 0;
    for (auto& r : h->algo->dataSet) n += r.size();
    *nnz = n;
    *feature_cnt = h->algo->feature_cnt;
    *field_cnt = h->algo->field_cnt;
    *factor_cnt = h->algo->factor_cnt;
}
// Dump of FM_Algo_Abst::dataSet / label (fm_algo_abst.h:156,170) as CSR.
void ref_get_data(void* hp, int64_t* row_ptr, uint64_t* fid, uint64_t* field, float* val, int* label) {
    Handle* h = (Handle*)hp;
    size_t p = 0;
    row_ptr[0] = 0;
    for (size_t r = 0; r < h->algo->dataSet.size(); r++) {
        for (auto& f : h->algo->dataSet[r]) {
            fid[p] = f.first; field[p] = f.field; val[p] = f.second; p++;
        }
        row_ptr[r + 1] = (int64_t)p;
        label[r] = h->algo->label[r];
    }
}
void ref_get_params(void* hp, float* W, float* V, float* sumVX) {
    Handle* h = (Handle*)hp;
    FM_Algo_Abst* a = h->algo;
    if (W) memcpy(W, a->W, sizeof(float) * a->feature_cnt);
    if (V) memcpy(V, a->V, sizeof(float) * v_size(a));
    if (sumVX && a->sumVX) memcpy(sumVX, a->sumVX, sizeof(float) * a->dataRow_cnt * a->factor_cnt);
}
void ref_set_params(void* hp, const float* W, const float* V) {
    Handle* h = (Handle*)hp;
    FM_Algo_Abst* a = h->algo;
    if (W) memcpy(a->W, W, sizeof(float) * a->feature_cnt);
    if (V) memcpy(a->V, V, sizeof(float) * v_size(a));
}
// One call == one reference Train() with epoch_cnt == 1 (main.cpp:228 calls Train() in a loop
// on the same object; optimizer accumulators persist).  Returns the summed logloss and accuracy
// exactly as the reference would print them.
void ref_train_epoch(void* hp, float* loss, float* acc) {
    Handle* h = (Handle*)hp;
    Quiet q(true);
    h->algo->Train();
    if (h->kind == K_NFM) {
        Train_NFM_Algo* t = (Train_NFM_Algo*)h->algo;
        *loss = t->loss;
        *acc = (float)(1.0 * t->accuracy / t->dataRow_cnt);
    } else {
        *loss = h->algo->__loss;
        *acc = h->algo->__accuracy / h->algo->dataRow_cnt;
    }
}
// NFM's two FC layers (train_nfm_algo.cpp:23-27): weights [out][in], bias[out], dropout_mask[out].
void ref_nfm_get_fc(void* hp, int layer, float* weight, float* bias, float* mask) {
    Handle* h = (Handle*)hp;
    Train_NFM_Algo* t = (Train_NFM_Algo*)h->algo;
    Fully_Conn_Layer<Sigmoid>* L = layer == 0 ? t->inputLayer : t->outputLayer;
    size_t in = L->input_dimension, out = L->output_dimension;
    if (weight) memcpy(weight, L->weight, sizeof(float) * in * out);
    if (bias) memcpy(bias, L->bias, sizeof(float) * out);
    if (mask) memcpy(mask, L->dropout_mask, sizeof(float) * out);
}
// Wall time of Train() alone for `epochs` epochs (CPU baseline; SURVEY.md 8d "CPU reference timing").
double ref_time_train(void* hp, int epochs) {
    Handle* h = (Handle*)hp;
    Quiet q(true);
    auto t0 = std::chrono::steady_clock::now();
    for (int e = 0; e < epochs; e++) h->algo->Train();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
// FM_Predict (predict/fm_predict.cpp:12-90).  Runs the reference predictor, saving the per-row pCTR
// through its own savePath (text, default ostream precision) and capturing its stdout line.
// out_text receives "total log likelihood = ... correct = ... auc = ..." .
int ref_predict(void* hp, const char* test_path, const char* save_path, char* out_text, int out_cap) {
    Handle* h = (Handle*)hp;
    char tmpl[] = "/tmp/lctr_ref_predXXXXXX";
    int fd = mkstemp(tmpl);
    if (fd < 0) return -1;
    {
        fflush(stdout); std::cout.flush();
        int saved = dup(1);
        dup2(fd, 1);
        {
            FM_Predict pred(h->algo, test_path, true);
            pred.Predict(save_path ? save_path : "");
        }
        fflush(stdout); std::cout.flush();
        dup2(saved, 1); close(saved);
    }
    lseek(fd, 0, SEEK_SET);
    int n = (int)read(fd, out_text, out_cap - 1);
    if (n < 0) n = 0;
    out_text[n] = 0;
    close(fd); unlink(tmpl);
    return 0;
}

// ---- unit-level access to the updaters that no in-tree trainer wires up ----
// AdagradUpdater_Num::update (util/gradientUpdater.h:139-150).
void ref_adagrad_update(size_t n, size_t minibatch, float lr, float* accum, float* w, float* g) {
    GradientUpdater::__global_minibatch_size = minibatch;
    GradientUpdater::__global_learning_rate = lr;
    AdagradUpdater_Num u;
    u.learnable_params_cnt(n);
    memcpy(u.__adagrad_accum.data(), accum, sizeof(float) * n);
    u.update(0, n, w, g);
    memcpy(accum, u.__adagrad_accum.data(), sizeof(float) * n);
}
// RMSpropUpdater_Num::update (util/gradientUpdater.h:200-233).
void ref_rmsprop_update(size_t n, size_t minibatch, float lr, float ema, float* accum, float* w, float* g) {
    GradientUpdater::__global_minibatch_size = minibatch;
    GradientUpdater::__global_learning_rate = lr;
    GradientUpdater::__global_ema_rate = ema;
    RMSpropUpdater_Num u;
    u.learnable_params_cnt(n);
    memcpy(u.__rms_accum.data(), accum, sizeof(float) * n);
    u.update(0, n, w, g);
    memcpy(accum, u.__rms_accum.data(), sizeof(float) * n);
}
// FTRLUpdater::update (util/gradientUpdater.h:252-273); state arrays z,n (sigma is scratch).
void ref_ftrl_update(size_t n, float* z, float* nn, float* w, float* g) {
    FTRLUpdater u;
    u.learnable_params_cnt(n);
    memcpy(u.ftrl_z, z, sizeof(float) * n);
    memcpy(u.ftrl_n, nn, sizeof(float) * n);
    u.update(0, n, w, g);
    memcpy(z, u.ftrl_z, sizeof(float) * n);
    memcpy(nn, u.ftrl_n, sizeof(float) * n);
}
// AdamUpdater_Num::update (util/momentumUpdater.h:187-210); iter_before = calls already made.
void ref_adam_update(size_t n, size_t minibatch, float lr, float beta1, float beta2, size_t iter_before,
                     float* m, float* v, float* w, float* g) {
    GradientUpdater::__global_minibatch_size = minibatch;
    GradientUpdater::__global_learning_rate = lr;
    MomentumUpdater::__global_momentum = beta1;
    MomentumUpdater::__global_momentum_adam2 = beta2;
    AdamUpdater_Num u;
    u.learnable_params_cnt(n);
    u.iter = iter_before;
    memcpy(u.__adam_accum.data(), m, sizeof(float) * n);
    memcpy(u.__adam_accum.data() + n, v, sizeof(float) * n);
    u.update(0, n, w, g);
    memcpy(m, u.__adam_accum.data(), sizeof(float) * n);
    memcpy(v, u.__adam_accum.data() + n, sizeof(float) * n);
}
// AdadeltaUpdater_Num::update (util/momentumUpdater.h:74-111); eg = E[g^2], ed = E[delta^2].
void ref_adadelta_update(size_t n, size_t minibatch, float momentum, float* eg, float* ed, float* w, float* g) {
    GradientUpdater::__global_minibatch_size = minibatch;
    MomentumUpdater::__global_momentum = momentum;
    AdadeltaUpdater_Num u;
    u.learnable_params_cnt(n);
    memcpy(u.__adadelta_accum.data(), eg, sizeof(float) * n);
    memcpy(u.__adadelta_accum.data() + n, ed, sizeof(float) * n);
    u.update(0, n, w, g);
    memcpy(eg, u.__adadelta_accum.data(), sizeof(float) * n);
    memcpy(ed, u.__adadelta_accum.data() + n, sizeof(float) * n);
}
// Sigmoid::forward (util/activations.h:65-72) and avx_dotProduct (common/avx.h:109-127).
float ref_sigmoid(float x) { Sigmoid s; return s.forward(x); }
float ref_dot(const float* x, const float* y, size_t n) { return avx_dotProduct(x, y, n); }
// GaussRand stream (util/random.h:42-58) after srand(seed): n deviates, scaled like fm_algo_abst.h:62-65.
void ref_gauss_fill(unsigned seed, size_t n, size_t factor_cnt, float* out) {
    srand(seed);
    const float scale = 1.0 / sqrt(factor_cnt);
    for (size_t i = 0; i < n; i++) out[i] = GaussRand() * scale;
    if (n & 1) (void)GaussRand();
}
unsigned ref_hw_threads() { return std::thread::hardware_concurrency(); }

}  // extern "C"
