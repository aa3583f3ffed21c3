// lightctr_b200/host/main_example.cpp -- the reference's driver flow (main.cpp:140-162,226-253) on the CUDA path.
//
//   main_example fm|ffm|nfm <train.csv> <test.csv> <T> <factor_cnt> <field_cnt|hidden> <seed> [out.bin]
//   main_example nfmc|nfmc_bf16 <train.csv> <test.csv> <T> <factor_cnt> <h0,h1,...> <seed>     (config C4: layer chain;
//                                                                   _bf16 = dense layers on the tensor cores)
//
// Builds with plain g++ against include/lightctr_b200.h + liblightctr_b200.so (tests/test_host_shim_gpu.py):
//   g++ -O2 -std=c++11 main_example.cpp -L../lib -llightctr_b200 -Wl,-rpath,$PWD/../lib -o main_example
#include "lightctr_gpu.h"

LIGHTCTR_B200_DEFINE_GLOBALS

using namespace lightctr_b200;

int main(int argc, const char* argv[]) {
    if (argc < 8) {
        puts("usage: main_example fm|ffm|nfm train.csv test.csv T factor_cnt field_cnt|hidden seed [out.bin]");
        return 2;
    }
    const std::string algo = argv[1], train_path = argv[2], test_path = argv[3];
    int T = atoi(argv[4]);
    const size_t k = (size_t)atoi(argv[5]), extra = (size_t)atoi(argv[6]);
    srand((uint32_t)atoi(argv[7]));  // main.cpp:78 uses time(NULL)

    FM_Algo_Abst* train = NULL;
    if (algo == "fm") train = new Train_FM_Algo(train_path, /*epoch*/ 1, /*factor_cnt*/ k);
    else if (algo == "ffm") train = new Train_FFM_Algo(train_path, /*epoch*/ 1, /*factor_cnt*/ k, /*field*/ extra);
    else if (algo == "nfm") train = new Train_NFM_Algo(train_path, /*epoch*/ 1, /*factor_cnt*/ k, /*hidden*/ extra);
    else if (algo == "nfmc" || algo == "nfmc_bf16") {
        std::vector<size_t> hidden;
        for (const char* p = argv[6]; *p;) { hidden.push_back((size_t)strtoul(p, (char**)&p, 10)); if (*p == ',') p++; }
        Train_NFM_Algo* t = new Train_NFM_Algo(train_path, 1, k, hidden);
        if (algo == "nfmc_bf16") t->mlp_precision = LCTR_MLP_BF16;
        train = t;
    }
    else { puts("unknown algo"); return 2; }
    FM_Predict* pred = algo.compare(0, 3, "nfm") == 0 ? NULL : new FM_Predict(train, test_path, true);

    while (T--) {
        train->Train();
        if (pred && T == 0) pred->Predict("");  // main.cpp:230-233 predicts after every Train(); once is enough here
        std::cout << "------------" << std::endl;
    }
    if (argc > 8) {  // dump W then V for the parity test
        FILE* f = fopen(argv[8], "wb");
        size_t nv = train->feature_cnt * train->factor_cnt * (train->field_cnt > 0 ? train->field_cnt : 1);
        fwrite(train->W, sizeof(float), train->feature_cnt, f);
        fwrite(train->V, sizeof(float), nv, f);
        fclose(f);
    }
    delete pred;
    delete train;
    puts("Exit 0");
    return 0;
}
