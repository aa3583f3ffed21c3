"""The oracle's Wide&Deep restatement (orc_wnd_epoch; Distributed_Algo_Abst::batchGradCompute, distributed_algo_abst.h:
176-280) cannot be pinned to a compiled reference (ZeroMQ).  This test re-derives one minibatch independently in numpy
float64 from the source text -- wide sum, first-entry-per-field concat, sigmoid chain, gradients -- and checks the C
restatement against it, so that a slip in the restatement does not silently become the GPU tests' ground truth."""
import numpy as np


def test_wnd_restatement_matches_independent_numpy(oracle_api):
    rng = np.random.default_rng(5)
    rows, F, Fc, d, H = 60, 500, 7, 4, 8
    cnt = rng.integers(3, 15, rows)
    rp = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    fid = np.concatenate([rng.choice(F, c, replace=False) for c in cnt]).astype(np.uint32)
    fld = rng.integers(0, Fc, len(fid)).astype(np.uint32)
    val = (rng.random(len(fid)) + 0.5).astype(np.float32)
    lab = (rng.random(rows) < 0.4).astype(np.int32)
    W0 = (rng.standard_normal(F) * 0.05).astype(np.float32)
    E0 = (rng.standard_normal(F * d) * 0.2).astype(np.float32)
    ds = oracle_api.Dataset(rp, fid, fld, val, lab, F, Fc)
    lr, l2 = 0.05, 0.001
    o = oracle_api.WNDOracle(ds, d, [H], W0, E0, lr=lr, l2=l2, batch_size=rows, minibatch=rows)
    for l in range(2):
        o.mlp.arrays("mask", l)[:] = 1.0
    w1 = o.mlp.arrays("weight", 0).reshape(H, Fc * d).astype(np.float64)
    b1 = o.mlp.arrays("bias", 0).astype(np.float64)
    w2 = o.mlp.arrays("weight", 1).reshape(1, H).astype(np.float64)
    b2 = o.mlp.arrays("bias", 1).astype(np.float64)
    # ---- independent float64 forward / backward of the whole minibatch
    W, E = W0.astype(np.float64), E0.reshape(F, d).astype(np.float64)
    gW, gE = np.zeros(F), np.zeros((F, d))
    gw1, gb1, gw2, gb2 = np.zeros_like(w1), np.zeros_like(b1), np.zeros_like(w2), np.zeros_like(b2)
    loss = 0.0
    for r in range(rows):
        ids, fl, x = fid[rp[r]:rp[r + 1]], fld[rp[r]:rp[r + 1]], val[rp[r]:rp[r + 1]].astype(np.float64)
        wide = float(np.sum(W[ids] * x))
        deep, first = np.zeros(Fc * d), {}
        for f, a in zip(ids, fl):
            if a not in first:
                first[a] = f
                deep[a * d:(a + 1) * d] = E[f]
        h = 1.0 / (1.0 + np.exp(-(w1 @ deep + b1)))
        out = float((w2 @ h + b2)[0])
        p = 1.0 / (1.0 + np.exp(-(wide + out)))
        loss += -np.log(p) if lab[r] == 1 else -np.log(1.0 - p)
        dl = p - lab[r]
        gW[ids] += dl * x + l2 * W[ids]
        gw2 += dl * h[None, :]
        gb2 += dl
        dh = (w2[0] * dl) * h * (1 - h)
        gw1 += dh[:, None] * deep[None, :]
        gb1 += dh
        dz = w1.T @ dh
        for a, f in first.items():
            gE[f] += dz[a * d:(a + 1) * d]
    lo, _ = o.epoch()
    assert abs(lo - loss) < 1e-4 * loss

    def adagrad(w, g):  # first step from zero accumulators: w - lr * g1 / sqrt(g1^2 + 1e-7)
        g1 = g / rows
        return np.where(g1 != 0, w - lr * g1 / np.sqrt(g1 * g1 + 1e-7), w)
    assert np.max(np.abs(o.W - adagrad(W, gW))) < 2e-5
    assert np.max(np.abs(o.E.reshape(F, d) - adagrad(E, gE))) < 2e-5
    # where the gradient is not tiny the first Adagrad step is +-lr: the comparison above would hide a wrong magnitude,
    # so also compare the dense layers' updated weights where |g| is small enough to be in the linear regime
    assert np.max(np.abs(o.mlp.arrays("weight", 0).reshape(H, Fc * d) - adagrad(w1, gw1))) < 2e-4
    assert np.max(np.abs(o.mlp.arrays("weight", 1).reshape(1, H) - adagrad(w2, gw2))) < 2e-4
