// lightctr_b200/host/dnn_example.cpp -- a DL_Algo_Abst subclass over a Fully_Conn_Layer chain, the way the reference's
// Train_CNN_Algo / Train_RNN_Algo subclass it (train/train_cnn_algo.h), on the CUDA path.
//
//   dnn_example <mnist-style.csv> <epoch> <feature_cnt> <hidden> <seed> [noshuffle]
#include <algorithm>

#include "lightctr_gpu.h"

LIGHTCTR_B200_DEFINE_GLOBALS

using namespace lightctr_b200;

class Train_DNN_Algo : public DL_Algo_Abst<Logistic, Sigmoid, Sigmoid> {
public:
    Train_DNN_Algo(std::string dataPath, size_t epoch, size_t feature_cnt, size_t hidden_size)
        : DL_Algo_Abst<Logistic, Sigmoid, Sigmoid>(dataPath, epoch, feature_cnt, hidden_size) {
        initNetwork(hidden_size);
    }
    void initNetwork(size_t hidden_size) {  // input -> hidden -> hidden / 2 -> 1
        this->inputLayer = new Fully_Conn_Layer<Sigmoid>(NULL, this->feature_cnt, hidden_size);
        this->appendNNLayer(this->inputLayer);
        Layer_Base* mid = new Fully_Conn_Layer<Sigmoid>(this->inputLayer, hidden_size, hidden_size / 2);
        this->appendNNLayer(mid);
        this->outputLayer = new Fully_Conn_Layer<Sigmoid>(mid, hidden_size / 2, 1);
        this->appendNNLayer(this->outputLayer);
    }
    const std::vector<float>& Predict(const std::vector<size_t>& rids, std::vector<std::vector<float> >& dataSet) {
        batch.resize(rids.size() * this->feature_cnt);
        for (size_t i = 0; i < rids.size(); i++)
            std::copy(dataSet[rids[i]].begin(), dataSet[rids[i]].end(), batch.begin() + i * this->feature_cnt);
        return this->inputLayer->forward(batch, rids.size());
    }
    void BP(const std::vector<size_t>& rids, const std::vector<float>& grad) { this->outputLayer->backward(grad, rids.size()); }
    void applyBP(size_t) const { this->inputLayer->applyBatchGradient(); }

private:
    std::vector<float> batch;
};

int main(int argc, const char* argv[]) {
    if (argc < 6) { puts("usage: dnn_example data.csv epoch feature_cnt hidden seed [noshuffle]"); return 2; }
    srand((uint32_t)atoi(argv[5]));
    Train_DNN_Algo t(argv[1], (size_t)atoi(argv[2]), (size_t)atoi(argv[3]), (size_t)atoi(argv[4]));
    if (argc > 6) t.shuffle = false;
    t.Train();
    puts("Exit 0");
    return 0;
}
