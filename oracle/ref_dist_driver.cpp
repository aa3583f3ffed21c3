// oracle/ref_dist_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// One translation unit around the UNMODIFIED reference cluster roles (main.cpp:118-139): Master, ParamServer<Key, Value>
// and the Distributed_Algo_Abst worker, compiled where they lie under /root/reference by oracle/Makefile (target refdist)
// into three binaries -- the reference's networking switches behaviour on the MASTER / PS / WORKER macros at compile time
// (common/network.h:254-261), so one binary per role as in its own Makefile:28-35.  libzmq: the reference vendors a
// Mach-O libzmq.a that cannot link here; the pyzmq wheel's bundled libzmq.so.5 (4.3.5) provides the same C API
// (third/zeromq/include/zmq.h).
//
// Purpose: the loss curve of a real master + PS + worker run over ZeroMQ on this machine is the pin of the oracle's
// restatement of the worker (distributed_algo_abst.h:170-282) and of the parameter server's update rules
// (distribut/paramserver.h:130-300); tests/golden/make_wnd_ref_curve.py records it.
//
// Deviations from main.cpp, none arithmetic:
//  (1) srand(argv seed) instead of srand(time(NULL)) (main.cpp:78);
//  (2) the worker reads <prefix>_<rank>.csv and trains for argv epochs (main.cpp:135-138 hard-codes ./data/ad_data, 100);
//  (3) the worker process exits after Train() without the shutdown handshake (the destructor blocks on the cluster);
//  (4) the GradientUpdater statics carry main.cpp:64-74's values.
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
#include <map>
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

// (6) as in oracle/ref_driver.cpp: fresh arrays are zeroed.  Fully_Conn_Layer::init's memset byte counts miss sizeof(float)
//     (train/layer/fullyconnLayer.h:57-60), so the tail of weightDelta / biasDelta is heap garbage in the reference and the
//     loss curve depends on unrelated allocations (observed: 444.4 .. 445.3 for the same seeds); zeroed arrays make the run
//     what the reference means, and reproducible.
void* operator new[](std::size_t n) {
    void* p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

#define private public      // (5) the worker's dense layers are read back when LCTR_REF_DEBUG is set (observation only)
#define protected public
#include "LightCTR/distribut/master.h"
#include "LightCTR/distribut/paramserver.h"
#include "LightCTR/distribut/dist_machine_abst.h"
#include "LightCTR/distribut/worker.h"
#include "LightCTR/distributed_algo_abst.h"

size_t GradientUpdater::__global_minibatch_size(50);
float GradientUpdater::__global_learning_rate(0.05);
float GradientUpdater::__global_ema_rate(0.99);
float GradientUpdater::__global_sparse_rate(0.8);
float GradientUpdater::__global_lambdaL2(0.001f);
float GradientUpdater::__global_lambdaL1(1e-5);
float MomentumUpdater::__global_momentum(0.8);
float MomentumUpdater::__global_momentum_adam2(0.999);
bool GradientUpdater::__global_bTraining(true);

int main(int argc, char** argv) {
    setvbuf(stdout, NULL, _IOLBF, 0);
    const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
    srand(seed);
#ifdef MASTER
    { Master master(Run_Mode::PS_Mode); }
#elif defined PS
    {   // argv[2]: the server's updater (paramserver.h:22-27; main.cpp:130 uses the default, SGD)
        const int u = argc > 2 ? atoi(argv[2]) : 0;
        ParamServer<Key, Value>(u == 1 ? UpdaterType::Adagrad : (u == 2 ? UpdaterType::DCASGD : (u == 3 ? UpdaterType::DCASGDA : UpdaterType::SGD)));
    }
#elif defined WORKER
    {
        if (argc < 4) { fprintf(stderr, "usage: %s seed data_prefix epochs\n", argv[0]); return 2; }
        Distributed_Algo_Abst* train = new Distributed_Algo_Abst(argv[2], (size_t)atoi(argv[3]));
        if (getenv("LCTR_REF_DEBUG")) {
            double s0 = 0, s1 = 0;
            for (size_t i = 0; i < train->field_cnt * train->factor_dim * 50; i++) s0 += train->inputLayer->weight[i];
            for (size_t i = 0; i < 50; i++) s1 += train->outputLayer->weight[i];
            printf("[debug] fields %zu input weights %.9g %.9g %.9g sum %.9g | output %.9g %.9g sum %.9g | mask0 %d %d %d %d\n", train->field_cnt,
                   train->inputLayer->weight[0], train->inputLayer->weight[1], train->inputLayer->weight[2], s0,
                   train->outputLayer->weight[0], train->outputLayer->weight[1], s1, (int)train->inputLayer->dropout_mask[0],
                   (int)train->inputLayer->dropout_mask[1], (int)train->inputLayer->dropout_mask[2], (int)train->inputLayer->dropout_mask[3]);
        }
        if (getenv("LCTR_REF_DEBUG")) {  // sample 0's tensors as the worker receives them (the same pull Train() starts with)
            train->tensor_map.clear();
            std::set<size_t> fields;
            for (size_t i = 0; i < train->dataSet[0].size(); i++)
                if (fields.count(train->dataSet[0][i].field) == 0) {
                    train->tensor_map.insert(std::make_pair(train->dataSet[0][i].first, train->dataSet[0][i].field));
                    fields.insert(train->dataSet[0][i].field);
                }
            train->worker.pull_tensor_op.sync(train->tensor_map, 1);
            printf("[debug] sample0 tensors:");
            for (auto it = train->tensor_map.begin(); it != train->tensor_map.end(); it++) {
                auto mem = train->param_buf->getMemory(it->second);
                printf(" %zu:%zu:%.9g,%.9g,%.9g,%.9g", it->first, it->second, mem.first[0], mem.first[1], mem.first[2], mem.first[3]);
            }
            printf("\n");
            Matrix deep_input(1, train->field_cnt * train->factor_dim);
            deep_input.zeroInit();
            for (auto it = train->tensor_map.begin(); it != train->tensor_map.end(); it++) {
                auto mem = train->param_buf->getMemory(it->second);
                memcpy(deep_input.pointer()->data() + it->second * train->factor_dim, mem.first, train->factor_dim * sizeof(float));
            }
            vector<Matrix*> wrapper(1);
            wrapper[0] = &deep_input;
            auto ans = train->inputLayer->forward(wrapper);
            printf("[debug] sample0 deep forward %.9g\n", ans[0]);
            if (getenv("LCTR_REF_DEBUG_BATCHES")) {  // cumulative loss after each of the first minibatches, then stop
                const int nb = atoi(getenv("LCTR_REF_DEBUG_BATCHES"));
                GradientUpdater::__global_bTraining = true;
                train->train_loss = 0; train->accuracy = 0;
                for (int p = 0; p < nb; p++) {
                    train->batchGradCompute(p + 1, p * 50, (p + 1) * 50, false);
                    printf("[debug] after batch %d loss %.9g\n", p, train->train_loss);
                    {
                        double s0 = 0, s1 = 0, b0 = 0;
                        for (size_t i = 0; i < train->field_cnt * train->factor_dim * 50; i++) s0 += train->inputLayer->weight[i];
                        for (size_t i = 0; i < 50; i++) { s1 += train->outputLayer->weight[i]; b0 += train->inputLayer->bias[i]; }
                        printf("[debug]   dense: input w sum %.9g bias sum %.9g output w sum %.9g bias %.9g mask %d%d%d%d%d%d%d%d\n", s0, b0, s1,
                               train->outputLayer->bias[0], (int)train->inputLayer->dropout_mask[0], (int)train->inputLayer->dropout_mask[1],
                               (int)train->inputLayer->dropout_mask[2], (int)train->inputLayer->dropout_mask[3], (int)train->inputLayer->dropout_mask[4],
                               (int)train->inputLayer->dropout_mask[5], (int)train->inputLayer->dropout_mask[6], (int)train->inputLayer->dropout_mask[7]);
                        // the wide weights of sample 0's entries as a fresh pull returns them
                        train->pull_map.clear();
                        for (size_t i = 0; i < 6 && i < train->dataSet[0].size(); i++) train->pull_map.insert(make_pair(train->dataSet[0][i].first, Value()));
                        train->worker.pull_op.sync(train->pull_map, p + 1);
                        printf("[debug]   wide:");
                        for (size_t i = 0; i < 6 && i < train->dataSet[0].size(); i++) printf(" %zu=%.9g", train->dataSet[0][i].first, train->pull_map[train->dataSet[0][i].first].w);
                        printf("\n");
                    }
                }
                if (getenv("LCTR_REF_DEBUG_PREDICT")) {
                    const size_t r0 = (size_t)nb * 50;
                    train->tensor_map.clear();
                    std::set<size_t> fs;
                    for (size_t i = 0; i < train->dataSet[r0].size(); i++)
                        if (fs.count(train->dataSet[r0][i].field) == 0) {
                            train->tensor_map.insert(std::make_pair(train->dataSet[r0][i].first, train->dataSet[r0][i].field));
                            fs.insert(train->dataSet[r0][i].field);
                        }
                    train->worker.pull_tensor_op.sync(train->tensor_map, nb + 1);
                    printf("[debug] row %zu tensors:", r0);
                    for (auto it = train->tensor_map.begin(); it != train->tensor_map.end(); it++) {
                        auto mem = train->param_buf->getMemory(it->second);
                        printf(" %zu:%zu:%.9g,%.9g,%.9g,%.9g", it->first, it->second, mem.first[0], mem.first[1], mem.first[2], mem.first[3]);
                    }
                    printf("\n");
                    {
                        Matrix deep_input(1, train->field_cnt * train->factor_dim);
                        deep_input.zeroInit();
                        for (auto it = train->tensor_map.begin(); it != train->tensor_map.end(); it++) {
                            auto mem = train->param_buf->getMemory(it->second);
                            memcpy(deep_input.pointer()->data() + it->second * train->factor_dim, mem.first, train->factor_dim * sizeof(float));
                        }
                        vector<Matrix*> wrapper(1);
                        wrapper[0] = &deep_input;
                        auto ans = train->inputLayer->forward(wrapper);
                        train->pull_map.clear();
                        for (size_t i = 0; i < train->dataSet[r0].size(); i++) train->pull_map.insert(make_pair(train->dataSet[r0][i].first, Value()));
                        train->worker.pull_op.sync(train->pull_map, nb + 1);
                        float pred = 0;
                        for (size_t i = 0; i < train->dataSet[r0].size(); i++) pred += train->pull_map[train->dataSet[r0][i].first].w * train->dataSet[r0][i].second;
                        printf("[debug] row %zu wide %.9g deep %.9g\n", r0, pred, ans[0]);
                    }
                    // forward-only losses of the next rows under the state reached
                    const int n = atoi(getenv("LCTR_REF_DEBUG_PREDICT"));
                    for (int q = 0; q < n; q++) {
                        train->train_loss = 0;
                        train->batchGradCompute(nb + 1, nb * 50 + q, nb * 50 + q + 1, true);
                        printf("[debug] forward-only row %d loss %.9g\n", nb * 50 + q, train->train_loss);
                    }
                }
                fflush(stdout);
                _exit(0);
            }
        }
        train->Train();
        fflush(stdout);
        _exit(0);
    }
#else
#error "compile with -D MASTER, -D PS or -D WORKER"
#endif
    return 0;
}
