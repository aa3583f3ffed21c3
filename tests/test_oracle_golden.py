"""The C oracle against the committed golden vectors (generated from the compiled reference by
tests/golden/make_golden.py).  Runs anywhere (no /root/reference needed)."""
import ctypes as C
import hashlib

import numpy as np

from golden_util import golden, load_csr


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def f32(x):
    return float(np.float32(x))


def test_kat_sigmoid_dot_gauss(oracle_api):
    g = golden()["kat"]
    L = oracle_api.lib()
    for x, bits in zip(g["sigmoid_x"], g["sigmoid_bits"]):
        assert int(np.float32(L.orc_sigmoid(x)).view(np.uint32)) == bits
    for d in g["dot"]:
        x, y = np.array(d["x"], np.float32), np.array(d["y"], np.float32)
        assert int(np.float32(L.orc_dot(x, y, len(x))).view(np.uint32)) == d["bits"]
    L.orc_srand(1)
    L.orc_gauss_reset()
    v = np.zeros(64, np.float32)
    L.orc_init_V(v, 64, 8)
    assert [int(b) for b in v.view(np.uint32)] == g["gauss_seed1_k8_bits"]


def test_kat_optimizers(oracle_api):
    g = golden()["kat"]
    L = oracle_api.lib()
    i = g["opt_in"]
    w, gr, s1, s2 = (np.array(i[k], np.float32) for k in ("w", "g", "s1", "s2"))
    n = len(w)
    a = [x.copy() for x in (s1, w, gr)]
    L.orc_adagrad(n, a[1], a[2], a[0], 1000, 0.05)
    assert [int(v) for v in a[0].view(np.uint32)] == g["adagrad"]["acc"]
    assert [int(v) for v in a[1].view(np.uint32)] == g["adagrad"]["w"]
    a = [x.copy() for x in (s1, s2, w, gr)]
    L.orc_ftrl(n, a[2], a[3], a[0], a[1], 0)
    assert [int(v) for v in a[0].view(np.uint32)] == g["ftrl"]["z"]
    assert [int(v) for v in a[1].view(np.uint32)] == g["ftrl"]["n"]
    assert [int(v) for v in a[2].view(np.uint32)] == g["ftrl"]["w"]
    a = [x.copy() for x in (s1, s2, w, gr)]
    it = C.c_size_t(3)
    L.orc_adam(n, a[2], a[3], a[0], a[1], C.byref(it), 1000, 0.05, 0.8, 0.999)
    assert [int(v) for v in a[0].view(np.uint32)] == g["adam_iter3"]["m"]
    assert [int(v) for v in a[1].view(np.uint32)] == g["adam_iter3"]["v"]
    assert [int(v) for v in a[2].view(np.uint32)] == g["adam_iter3"]["w"]


def test_fm_k8_curve_and_params(oracle_api):
    g = golden()
    ds = load_csr("train_sparse_csr.npz")
    assert (ds.rows, ds.nnz, ds.feature_cnt) == (g["train"]["rows"], g["train"]["nnz"], g["train"]["feature_cnt"])
    assert sha(ds.fid.astype(np.uint32)) == g["train"]["sha_fid"]
    W, V = oracle_api.init_params(1, ds.feature_cnt, 8)
    assert sha(V) == g["fm_k8"]["sha_V0"]
    o = oracle_api.FMOracle(ds, 8, W, V)
    for e in range(20):
        loss, acc = o.epoch()
        assert f32(loss) == g["fm_k8"]["loss"][e], e
        assert f32(acc) == g["fm_k8"]["acc"][e], e
    assert sha(o.W) == g["fm_k8"]["sha_W"] and sha(o.V) == g["fm_k8"]["sha_V"] and sha(o.sumVX) == g["fm_k8"]["sha_sumVX"]
    # FM_Predict (with its quirks) against the reference's printed line and saved pCTR file
    test = load_csr("test_sparse_csr.npz")
    pctr, loss, correct, auc = oracle_api.predict(test, 0, 8, o.W, o.V, o.sumVX, False)
    text = g["fm_k8"]["predict_text"]
    assert float(text.split("likelihood = ")[1].split()[0]) in (float("%.6g" % loss), float("%.5g" % loss))
    assert float("%.4f" % auc) == float(text.split("auc = ")[1].split()[0])
    ref_p = np.load(__import__("os").path.join(__import__("golden_util").GOLDEN, "fm_k8_pctr_test.npy"))
    assert np.allclose(pctr, ref_p, rtol=6e-6, atol=0)  # the reference file holds 6 significant digits


def test_ffm_k4_curve_and_params(oracle_api):
    g = golden()["ffm_k4"]
    ds = load_csr("train_sparse_csr.npz", field_cnt=68)
    W, V = oracle_api.init_params(1, ds.feature_cnt, 4, 68)
    assert sha(V) == g["sha_V0"]
    o = oracle_api.FFMOracle(ds, 4, W, V)
    for e in range(4):
        loss, acc = o.epoch()
        assert f32(loss) == g["loss"][e], e
        assert f32(acc) == g["acc"][e], e
    assert sha(o.W) == g["sha_W"] and sha(o.V) == g["sha_V"]


def test_nfm_k10_h32_curve_and_params(oracle_api):
    g = golden()["nfm_k10_h32"]
    ds = load_csr("train_sparse_csr.npz")
    o = oracle_api.NFMOracle(ds, 10, 32, seed=1)
    for e in range(3):
        loss, acc = o.epoch()
        assert f32(loss) == g["loss"][e], e
        assert f32(np.float32(acc)) == g["acc"][e] or abs(acc - g["acc"][e]) < 1e-6
    assert sha(o.W) == g["sha_W"] and sha(o.V) == g["sha_V"]
    assert sha(o.mlp.arrays("weight", 0)) == g["sha_fc1_w"] and sha(o.mlp.arrays("bias", 0)) == g["sha_fc1_b"]
    assert sha(o.mlp.arrays("weight", 1)) == g["sha_fc2_w"] and sha(o.mlp.arrays("bias", 1)) == g["sha_fc2_b"]
