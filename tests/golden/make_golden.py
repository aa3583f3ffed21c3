"""tests/golden/make_golden.py -- regenerates the committed golden fixtures.

Runs ONLY in the build container (needs /root/reference and oracle/_ref/libref.so, i.e. the
unmodified reference compiled in place).  Everything numeric below is produced by the REFERENCE
itself (libref.so), not by our restatement; tests/test_oracle_golden.py then checks the C oracle
against these files on any machine.

    python tests/golden/make_golden.py

Recorded provenance: reference commit 6204377 (SURVEY.md header), g++ 13.3.0,
flags -std=c++11 -O3 -D__AVX__ -mavx -mssse3 (reference Makefile:3), glibc rand(), srand(1), proc_cnt=1.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import api  # noqa: E402

REF = "/root/reference/data"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def dump_csr(name, ds, rows):
    np.savez_compressed(os.path.join(HERE, name), row_ptr=ds.row_ptr, fid=ds.fid, field=ds.field.astype(np.uint8),
                        val=ds.val if not np.all(ds.val == 1.0) else np.zeros(0, np.float32),
                        label=ds.label[:rows], feature_cnt=ds.feature_cnt)


def main():
    api.build(ref=True)
    out = {}
    # ---- inputs, exactly as the reference parser sees them --------------------------------------
    t = api.RefTrainer("ffm", REF + "/train_sparse.csv", 4, seed=1, proc_cnt=1, field_cnt=68)
    train = t.data()
    dump_csr("train_sparse_csr.npz", train, train.rows)
    out["train"] = dict(rows=train.rows, nnz=train.nnz, feature_cnt=train.feature_cnt, field_cnt=train.field_cnt,
                        sha_fid=sha(train.fid), sha_row_ptr=sha(train.row_ptr), sha_field=sha(train.field))
    t.close()
    test = api.load_test(REF + "/test_sparse.csv", train.feature_cnt)  # bit-checked against libref's Predict below
    dump_csr("test_sparse_csr.npz", test, test.rows)
    out["test"] = dict(rows=test.rows, nnz=test.nnz)

    # ---- FM k=8 (config C1) -----------------------------------------------------------------------
    t = api.RefTrainer("fm", REF + "/train_sparse.csv", 8, seed=1, proc_cnt=1)
    W0, V0, _ = t.params()
    curve = [t.epoch() for _ in range(20)]
    W, V, S = t.params()
    text = t.predict(REF + "/test_sparse.csv", "/tmp/_golden_fm_pred.txt")
    out["fm_k8"] = dict(loss=[float(np.float32(x[0])) for x in curve], acc=[float(np.float32(x[1])) for x in curve],
                        sha_V0=sha(V0), sha_W=sha(W), sha_V=sha(V), sha_sumVX=sha(S), predict_text=text.strip(),
                        sum_W=float(W.astype(np.float64).sum()), sum_V=float(V.astype(np.float64).sum()))
    np.save(os.path.join(HERE, "fm_k8_pctr_test.npy"), np.loadtxt("/tmp/_golden_fm_pred.txt", dtype=np.float64))
    t.close()

    # ---- FFM k=4, 68 fields ------------------------------------------------------------------------
    t = api.RefTrainer("ffm", REF + "/train_sparse.csv", 4, seed=1, proc_cnt=1, field_cnt=68)
    _, V0, _ = t.params()
    curve = [t.epoch() for _ in range(4)]
    W, V, _ = t.params()
    text = t.predict(REF + "/test_sparse.csv")
    out["ffm_k4"] = dict(loss=[float(np.float32(x[0])) for x in curve], acc=[float(np.float32(x[1])) for x in curve],
                         sha_V0=sha(V0), sha_W=sha(W), sha_V=sha(V), predict_text=text.strip(),
                         sum_W=float(W.astype(np.float64).sum()), sum_V=float(V.astype(np.float64).sum()))
    t.close()

    # ---- NFM k=10, H=32 (reference + zeroed fresh arrays == "memset sizes fixed", ref_driver.cpp) ----
    t = api.RefTrainer("nfm", REF + "/train_sparse.csv", 10, seed=1, hidden=32)
    curve = [t.epoch() for _ in range(3)]
    W, V, S = t.params()
    w1, b1, _ = t.fc(0, 10, 32)
    w2, b2, _ = t.fc(1, 32, 1)
    out["nfm_k10_h32"] = dict(loss=[float(np.float32(x[0])) for x in curve],
                              acc=[float(np.float32(x[1])) for x in curve], sha_W=sha(W), sha_V=sha(V),
                              sha_fc1_w=sha(w1), sha_fc1_b=sha(b1), sha_fc2_w=sha(w2), sha_fc2_b=sha(b2))
    t.close()

    # ---- small known-answer vectors -------------------------------------------------------------------
    R = api.ref()
    rng = np.random.default_rng(2024)
    kat = {}
    xs = np.concatenate([np.linspace(-18, 18, 73), [16.0, -16.0, 16.000002, -16.000002]]).astype(np.float32)
    kat["sigmoid_x"] = [float(x) for x in xs]
    kat["sigmoid_bits"] = [int(np.float32(R.ref_sigmoid(float(x))).view(np.uint32)) for x in xs]
    dots = []
    for n in (4, 8, 10, 16, 32, 50):
        x = rng.standard_normal(n).astype(np.float32)
        y = rng.standard_normal(n).astype(np.float32)
        dots.append(dict(x=[float(v) for v in x], y=[float(v) for v in y],
                         bits=int(np.float32(R.ref_dot(x, y, n)).view(np.uint32))))
    kat["dot"] = dots
    g = np.zeros(64, np.float32)
    R.ref_gauss_fill(1, 64, 8, g)
    kat["gauss_seed1_k8_bits"] = [int(v) for v in g.view(np.uint32)]
    n = 64
    w = rng.standard_normal(n).astype(np.float32)
    gr = (rng.standard_normal(n) * (rng.random(n) < 0.7)).astype(np.float32)
    s1 = np.abs(rng.standard_normal(n)).astype(np.float32)
    s2 = np.abs(rng.standard_normal(n)).astype(np.float32)
    kat["opt_in"] = dict(w=[float(v) for v in w], g=[float(v) for v in gr], s1=[float(v) for v in s1],
                         s2=[float(v) for v in s2])
    a = [x.copy() for x in (s1, w, gr)]
    R.ref_adagrad_update(n, 1000, 0.05, a[0], a[1], a[2])
    kat["adagrad"] = dict(acc=[int(v) for v in a[0].view(np.uint32)], w=[int(v) for v in a[1].view(np.uint32)])
    a = [x.copy() for x in (s1, s2, w, gr)]
    R.ref_ftrl_update(n, a[0], a[1], a[2], a[3])
    kat["ftrl"] = dict(z=[int(v) for v in a[0].view(np.uint32)], n=[int(v) for v in a[1].view(np.uint32)],
                       w=[int(v) for v in a[2].view(np.uint32)])
    a = [x.copy() for x in (s1, s2, w, gr)]
    R.ref_adam_update(n, 1000, 0.05, 0.8, 0.999, 3, a[0], a[1], a[2], a[3])
    kat["adam_iter3"] = dict(m=[int(v) for v in a[0].view(np.uint32)], v=[int(v) for v in a[1].view(np.uint32)],
                             w=[int(v) for v in a[2].view(np.uint32)])
    out["kat"] = kat
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
