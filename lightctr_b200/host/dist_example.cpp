// lightctr_b200/host/dist_example.cpp -- the reference's worker driver (main.cpp:253: `new Distributed_Algo_Abst(path, epoch)`,
// Train(), Predict()) on the CUDA path; one process per GPU.
//
//   LIGHTCTR_B200_RANK=r LIGHTCTR_B200_WORLD=R LIGHTCTR_B200_RDV=/tmp/rdv dist_example <data-prefix> <epoch> <seed> [out.bin]
#include "lightctr_gpu.h"

LIGHTCTR_B200_DEFINE_GLOBALS

using namespace lightctr_b200;

int main(int argc, const char* argv[]) {
    if (argc < 4) { puts("usage: dist_example data-prefix epoch seed [out.bin]"); return 2; }
    srand((uint32_t)atoi(argv[3]));
    Distributed_Algo_Abst* train = new Distributed_Algo_Abst(argv[1], (size_t)atoi(argv[2]));
    train->Train();
    train->Predict();
    if (argc > 4) {  // the shared parameters as this rank sees them (owned rows only when world > 1)
        std::vector<float> W(train->feature_cnt), E(train->feature_cnt * Distributed_Algo_Abst::factor_dim);
        LCTR_OK(lctr_download_params(train->ctx, W.data(), E.data()));
        FILE* f = fopen(argv[4], "wb");
        fwrite(W.data(), sizeof(float), W.size(), f);
        fwrite(E.data(), sizeof(float), E.size(), f);
        fclose(f);
    }
    delete train;
    puts("Exit 0");
    return 0;
}
