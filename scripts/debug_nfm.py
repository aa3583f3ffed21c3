import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import api
from lightctr_b200 import capi
from golden_util import load_csr
full = load_csr("train_sparse_csr.npz")
R = 50
nnz = full.row_ptr[R]
ds = api.Dataset(full.row_ptr[:R+1], full.fid[:nnz], full.field[:nnz], full.val[:nnz], full.label[:R], full.feature_cnt, 0)
k, H = 10, 32
o = api.NFMOracle(ds, k, H, seed=1)
ctx = capi.Context(capi.MODEL_NFM, ds.feature_cnt, k, 0, hidden=[H], minibatch_size=50)
ctx.upload_params(o.W, o.V)
for l in range(2):
    i_, o_ = o.mlp.dims[l], o.mlp.dims[l+1]
    ctx.mlp_upload(l, o.mlp.arrays("weight", l).copy(), o.mlp.arrays("bias", l).copy())
    ctx.mlp_set_mask(l, o.mlp.arrays("mask", l).copy())
ctx.upload_batch(0, ds.row_ptr, ds.fid, None, None, ds.label)
V0 = o.V.copy(); W0 = o.W.copy(); w0 = [o.mlp.arrays("weight", l).copy() for l in range(2)]
lg, cg = ctx.train_step(0)
lo, ao = o.epoch()
print("loss", lg, lo)
Wg, Vg = ctx.download_params()
print("dW max", np.abs(Wg - o.W).max(), "ref change", np.abs(o.W - W0).max())
print("dV max", np.abs(Vg - o.V).max(), "ref change", np.abs(o.V - V0).max())
for l in range(2):
    i_, o_ = o.mlp.dims[l], o.mlp.dims[l+1]
    w, b = ctx.mlp_download(l, i_, o_)
    print("layer", l, "w diff", np.abs(w - o.mlp.arrays("weight", l)).max(), "ref change", np.abs(o.mlp.arrays("weight", l) - w0[l]).max(),
          "b diff", np.abs(b - o.mlp.arrays("bias", l)).max(), "ref b", np.abs(o.mlp.arrays("bias", l)).max())
sv = ctx.download_sumvx(0)
print("sumvx diff", np.abs(sv - o.sumVX).max())
