import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import api
from lightctr_b200 import capi, trainers as T
from golden_util import load_csr, write_libffm
full = load_csr("train_sparse_csr.npz")
for R in (50, 100, 150):
    nnz = full.row_ptr[R]
    ds = api.Dataset(full.row_ptr[:R+1], full.fid[:nnz], full.field[:nnz], full.val[:nnz], full.label[:R], full.feature_cnt, 0)
    p = "/tmp/nfm_%d.txt" % R
    write_libffm(ds, p)
    T.srand(1); T.GradientUpdater.minibatch_size = 50
    nfm = T.Train_NFM_Algo(p, 1, 10, 32)
    o = api.NFMOracle(api.load(p), 10, 32, seed=1)
    print(R, "F", nfm.feature_cnt, o.ds.feature_cnt, "init eq", np.array_equal(nfm.V, o.V), np.array_equal(nfm.layers[0].weight.ravel(), o.mlp.arrays("weight",0)))
    for e in range(2):
        nfm.Train(); lo, ao = o.epoch()
        print("  epoch", e, nfm.loss_curve[-1], lo, "mask eq", np.array_equal(nfm.layers[0].mask, o.mlp.arrays("mask", 0)),
              "V diff", np.abs(nfm.V - o.V).max(), "w0 diff", np.abs(nfm.layers[0].weight.ravel() - o.mlp.arrays("weight", 0)).max())
