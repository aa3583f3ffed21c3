"""Pin the plain-C restatement (oracle/lightctr_oracle.c) bit-for-bit against the unmodified
reference compiled in place (oracle/_ref/libref.so).  CPU-only; needs /root/reference for the data."""
import numpy as np
import pytest

from conftest import REF_DATA, needs_reference

TRAIN = REF_DATA + "/train_sparse.csv"
TEST = REF_DATA + "/test_sparse.csv"
pytestmark = needs_reference


def test_rand_stream_matches_glibc(oracle_api):
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    L = oracle_api.lib()
    for seed in (1, 7, 12345, 0):
        libc.srand(seed)
        L.orc_srand(seed)
        a = [libc.rand() for _ in range(1000)]
        b = [L.orc_rand() for _ in range(1000)]
        assert a == b


def test_gauss_init_bit_exact(oracle_api):
    for seed, n, k in ((1, 4096, 8), (3, 1001, 16), (9, 10, 4)):
        ref = np.zeros(n, np.float32)
        oracle_api.ref().ref_gauss_fill(seed, n, k, ref)
        L = oracle_api.lib()
        L.orc_srand(seed)
        L.orc_gauss_reset()
        mine = np.zeros(n, np.float32)
        L.orc_init_V(mine, n, k)
        assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32))


def test_dot_and_sigmoid_bit_exact(oracle_api):
    rng = np.random.default_rng(0)
    R, L = oracle_api.ref(), oracle_api.lib()
    for n in (1, 3, 4, 7, 8, 9, 10, 15, 16, 17, 31, 32, 33, 64, 100, 255):
        for _ in range(20):
            x = rng.standard_normal(n).astype(np.float32)
            y = rng.standard_normal(n).astype(np.float32)
            a, b = R.ref_dot(x, y, n), L.orc_dot(x, y, n)
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)
    for x in list(np.linspace(-20, 20, 4001, dtype=np.float32)) + [16.0, -16.0, 16.000002, -16.000002]:
        a, b = R.ref_sigmoid(float(x)), L.orc_sigmoid(float(x))
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)


def test_loader_bit_exact(oracle_api):
    t = oracle_api.RefTrainer("ffm", TRAIN, 4, field_cnt=68)
    d_ref = t.data()
    d = oracle_api.load(TRAIN, field_cnt=68)
    assert (d.rows, d.nnz, d.feature_cnt, d.field_cnt) == (1000, 281975, 233789, 68)
    assert d_ref.feature_cnt == d.feature_cnt and d_ref.field_cnt == d.field_cnt
    for a, b in ((d.row_ptr, d_ref.row_ptr), (d.fid, d_ref.fid), (d.field, d_ref.field), (d.val, d_ref.val),
                 (d.label[:d.rows], d_ref.label)):
        assert np.array_equal(a, b)
    t.close()


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_fm_training_bit_exact(oracle_api):
    k = 8
    t = oracle_api.RefTrainer("fm", TRAIN, k, seed=1, proc_cnt=1)
    ds = oracle_api.load(TRAIN)
    W0, V0, _ = t.params()
    Wi, Vi = oracle_api.init_params(1, ds.feature_cnt, k)
    assert np.array_equal(_bits(V0), _bits(Vi)) and np.array_equal(_bits(W0), _bits(Wi))
    o = oracle_api.FMOracle(ds, k, Wi, Vi)
    for e in range(6):
        lr, ar = t.epoch()
        lo, ao = o.epoch()
        assert np.float32(lr).view(np.uint32) == np.float32(lo).view(np.uint32), (e, lr, lo)
        assert ar == ao
    W, V, S = t.params()
    assert np.array_equal(_bits(W), _bits(o.W))
    assert np.array_equal(_bits(V), _bits(o.V))
    assert np.array_equal(_bits(S), _bits(o.sumVX))
    # FM_Predict with its quirks (predict/fm_predict.cpp)
    text = t.predict(TEST)
    test = oracle_api.load_test(TEST, ds.feature_cnt)
    pctr, loss, correct, auc = oracle_api.predict(test, 0, k, o.W, o.V, o.sumVX, False)
    assert test.rows == 200
    ref_loss = float(text.split("likelihood = ")[1].split()[0])
    ref_acc = float(text.split("correct = ")[1].split()[0])
    ref_auc = float(text.split("auc = ")[1].split()[0])
    assert ref_loss in (float("%.6g" % loss), float("%.5g" % loss))
    assert float("%.5g" % (np.float32(correct) / np.float32(test.rows))) == pytest.approx(ref_acc, rel=1e-6)
    assert float("%.4f" % auc) == ref_auc
    t.close()


def test_ffm_training_bit_exact(oracle_api):
    k, Fc = 4, 68
    t = oracle_api.RefTrainer("ffm", TRAIN, k, seed=1, proc_cnt=1, field_cnt=Fc)
    ds = oracle_api.load(TRAIN, field_cnt=Fc)
    Wi, Vi = oracle_api.init_params(1, ds.feature_cnt, k, Fc)
    W0, V0, _ = t.params()
    assert np.array_equal(_bits(V0), _bits(Vi))
    o = oracle_api.FFMOracle(ds, k, Wi, Vi)
    for e in range(3):
        lr, ar = t.epoch()
        lo, ao = o.epoch()
        assert np.float32(lr).view(np.uint32) == np.float32(lo).view(np.uint32), (e, lr, lo)
        assert ar == ao
    W, V, _ = t.params()
    assert np.array_equal(_bits(W), _bits(o.W))
    assert np.array_equal(_bits(V), _bits(o.V))
    text = t.predict(TEST)
    test = oracle_api.load_test(TEST, ds.feature_cnt)
    pctr, loss, correct, auc = oracle_api.predict(test, Fc, k, o.W, o.V, None, True)
    # cout keeps setprecision(5) from an earlier Predict() in this process (fm_predict.cpp:73-74)
    ref_loss = float(text.split("likelihood = ")[1].split()[0])
    assert ref_loss in (float("%.6g" % loss), float("%.5g" % loss))
    assert float("%.4f" % auc) == float(text.split("auc = ")[1].split()[0])
    t.close()


def test_nfm_training_bit_exact(oracle_api):
    k, H = 10, 32
    t = oracle_api.RefTrainer("nfm", TRAIN, k, seed=1, hidden=H)
    ds = oracle_api.load(TRAIN)
    o = oracle_api.NFMOracle(ds, k, H, seed=1)
    w_ref, b_ref, m_ref = t.fc(0, k, H)
    assert np.array_equal(_bits(w_ref), _bits(o.mlp.arrays("weight", 0)))
    assert np.array_equal(m_ref, o.mlp.arrays("mask", 0))
    for e in range(3):
        lr, ar = t.epoch()
        lo, ao = o.epoch()
        assert np.float32(lr).view(np.uint32) == np.float32(lo).view(np.uint32), (e, lr, lo)
        assert ar == pytest.approx(ao, abs=1e-7)
    W, V, S = t.params()
    assert np.array_equal(_bits(W), _bits(o.W))
    assert np.array_equal(_bits(V), _bits(o.V))
    w_ref, b_ref, m_ref = t.fc(1, H, 1)
    assert np.array_equal(_bits(w_ref), _bits(o.mlp.arrays("weight", 1)))
    assert np.array_equal(_bits(b_ref), _bits(o.mlp.arrays("bias", 1)))
    t.close()


def test_optimizer_units_bit_exact(oracle_api):
    rng = np.random.default_rng(5)
    R, L = oracle_api.ref(), oracle_api.lib()
    n = 5000
    import ctypes as C
    for trial in range(3):
        w = rng.standard_normal(n).astype(np.float32)
        g = (rng.standard_normal(n) * (rng.random(n) < 0.7)).astype(np.float32)
        s1 = np.abs(rng.standard_normal(n)).astype(np.float32) * (trial > 0)
        s2 = np.abs(rng.standard_normal(n)).astype(np.float32) * (trial > 0)
        # adagrad
        a = [x.copy() for x in (s1, w, g)]
        b = [x.copy() for x in (s1, w, g)]
        R.ref_adagrad_update(n, 1000, 0.05, a[0], a[1], a[2])
        L.orc_adagrad(n, b[1], b[2], b[0], 1000, 0.05)
        for x, y in zip(a, b):
            assert np.array_equal(_bits(x), _bits(y))
        # rmsprop
        a = [x.copy() for x in (s1, w, g)]
        b = [x.copy() for x in (s1, w, g)]
        R.ref_rmsprop_update(n, 1000, 0.05, 0.99, a[0], a[1], a[2])
        L.orc_rmsprop(n, b[1], b[2], b[0], 1000, 0.05, 0.99)
        for x, y in zip(a, b):
            assert np.array_equal(_bits(x), _bits(y))
        # adadelta
        a = [x.copy() for x in (s1, s2, w, g)]
        b = [x.copy() for x in (s1, s2, w, g)]
        R.ref_adadelta_update(n, 1000, 0.8, a[0], a[1], a[2], a[3])
        L.orc_adadelta(n, b[2], b[3], b[0], b[1], 1000, 0.8)
        for x, y in zip(a, b):
            assert np.array_equal(_bits(x), _bits(y))
        # ftrl
        a = [x.copy() for x in (s1, s2, w, g)]
        b = [x.copy() for x in (s1, s2, w, g)]
        R.ref_ftrl_update(n, a[0], a[1], a[2], a[3])
        L.orc_ftrl(n, b[2], b[3], b[0], b[1], 0)
        for x, y in zip(a, b):
            assert np.array_equal(_bits(x), _bits(y))
        # adam
        a = [x.copy() for x in (s1, s2, w, g)]
        b = [x.copy() for x in (s1, s2, w, g)]
        R.ref_adam_update(n, 1000, 0.05, 0.8, 0.999, trial * 3, a[0], a[1], a[2], a[3])
        it = C.c_size_t(trial * 3)
        L.orc_adam(n, b[2], b[3], b[0], b[1], C.byref(it), 1000, 0.05, 0.8, 0.999)
        for x, y in zip(a, b):
            assert np.array_equal(_bits(x), _bits(y))
