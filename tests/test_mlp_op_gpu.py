"""The Fully_Conn_Layer chain as a stand-alone operator (lctr_mlp_forward / lctr_mlp_backward / lctr_mlp_apply,
include/lightctr_b200.h) and the C++ Layer_Base / Fully_Conn_Layer / DL_Algo_Abst shims over it
(lightctr_b200/host/lightctr_gpu.h), against the oracle's per-sample restatement of fullyconnLayer.h:80-206."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "lightctr_b200", "host")
LIBDIR = os.path.join(ROOT, "lightctr_b200", "lib")


def _oracle_chain(api, dims, seed, act=0, sparse_rate=0.8):
    api.lib().orc_srand(seed)
    return api.Mlp(dims, act, sparse_rate)


def test_mlp_operator_forward_backward_apply(oracle_api):
    from lightctr_b200 import capi
    api = oracle_api
    dims = [12, 20, 8, 1]
    rows, mb = 48, 48
    m = _oracle_chain(api, dims, 3)
    ctx = capi.Context(capi.MODEL_NFM, 1, dims[0], hidden=tuple(dims[1:-1]), mlp_precision=capi.MLP_FP32, minibatch_size=mb, lr=0.05)
    for l in range(len(dims) - 1):
        ctx.mlp_upload(l, m.arrays("weight", l).copy(), m.arrays("bias", l).copy())
        ctx.mlp_set_mask(l, m.arrays("mask", l).copy())
    rng = np.random.default_rng(0)
    for step in range(3):
        x = rng.standard_normal((rows, dims[0])).astype(np.float32)
        dout = (rng.standard_normal(rows) * np.where(rng.random(rows) < 0.1, 40.0, 1.0)).astype(np.float32)  # some beyond the +-15 clip
        out = ctx.mlp_forward(x)
        dx = ctx.mlp_backward(dout, dims[0])
        L = api.lib()
        want_out, want_dx = np.empty(rows, np.float32), np.empty((rows, dims[0]), np.float32)
        for r in range(rows):  # the reference's order: sample by sample, dW accumulated in sample order
            want_out[r] = L.orc_mlp_forward(m.p, x[r])
            L.orc_mlp_backward(m.p, float(dout[r]))
            want_dx[r] = m.arrays("in_delta", 0)
        assert np.array_equal(out.view(np.uint32), want_out.view(np.uint32)), step   # reference-order dots: bit-exact
        assert np.max(np.abs(dx - want_dx)) <= 1e-6 * max(1.0, float(np.max(np.abs(want_dx))))
        for l in range(len(dims) - 1):
            gw, gb = ctx.mlp_download_grad(l, dims[l], dims[l + 1])
            assert np.max(np.abs(gw - m.arrays("dW", l))) <= 1e-5 * max(1.0, float(np.max(np.abs(m.arrays("dW", l)))))
            assert np.max(np.abs(gb - m.arrays("db", l))) <= 1e-5 * max(1.0, float(np.max(np.abs(m.arrays("db", l)))))
        ctx.mlp_apply(mb)
        L.orc_mlp_apply(m.p, mb, np.float32(0.05), np.float32(0.8))  # also re-draws the masks from the rand() stream
        for l in range(len(dims) - 1):
            w, b = ctx.mlp_download(l, dims[l], dims[l + 1])
            assert np.max(np.abs(w - m.arrays("weight", l))) < 1e-6 and np.max(np.abs(b - m.arrays("bias", l))) < 1e-6, (step, l)
            ctx.mlp_set_mask(l, m.arrays("mask", l).copy())
    ctx.close()


def test_cxx_dl_algo_abst_subclass(oracle_api, tmp_path):
    """dnn_example.cpp: a DL_Algo_Abst<Logistic, Sigmoid, Sigmoid> subclass over a Fully_Conn_Layer chain, compiled with
    plain g++ against the C ABI, trained on an MNIST-format file; its validation losses against the same loop written
    with the oracle's per-sample layer functions and the same glibc rand() stream (shuffle off: std::random_shuffle's use
    of rand() is libstdc++-specific)."""
    from lightctr_b200 import build as lbuild
    lbuild.build()
    api = oracle_api
    exe = str(tmp_path / "dnn_example")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++11", os.path.join(HOST, "dnn_example.cpp"), "-L" + LIBDIR,
                           "-llightctr_b200", "-Wl,-rpath," + LIBDIR, "-L/usr/local/cuda/lib64",
                           "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe])
    rng = np.random.default_rng(5)
    rows, feat, hidden, seed, epochs, mb = 230, 16, 12, 7, 21, 50
    pix = rng.integers(0, 256, (rows, feat))
    digit = rng.integers(0, 10, rows)
    path = str(tmp_path / "mnist.csv")
    with open(path, "w") as f:
        for r in range(rows):
            f.write("%d %s\n" % (digit[r], " ".join(str(v) for v in pix[r])))
    text = subprocess.check_output([exe, path, str(epochs), str(feat), str(hidden), str(seed), "noshuffle"], text=True)
    got = [float(v) for v in re.findall(r"Epoch \d+ Loss = ([0-9.eE+-]+)", text)]
    # the same run on the oracle
    X = (pix / 255.0).astype(np.float32)
    X[pix == 0] = 0.0
    y = (digit >= 5).astype(np.int32)
    dims = [feat, hidden, hidden // 2, 1]
    L = api.lib()
    L.orc_srand(seed)
    m = api.Mlp(dims, 0, 0.8)

    def sig(v):
        return np.float32(L.orc_sigmoid(np.float32(v)))

    want, be = [], 0
    for _ in range(epochs):
        for b in range(0, rows, mb):
            for r in range(b, min(b + mb, rows)):
                p = sig(L.orc_mlp_forward(m.p, X[r]))
                L.orc_mlp_backward(m.p, float(np.float32(p - np.float32(y[r]))))
            L.orc_mlp_apply(m.p, mb, np.float32(0.05), np.float32(0.8))
            if be % 50 == 0:
                loss = np.float32(0)
                for r in range(rows):
                    p = sig(L.orc_mlp_forward(m.p, X[r]))
                    loss = np.float32(loss + (np.float32(-np.log(p)) if y[r] == 1 else np.float32(-np.log(np.float32(1.0) - p))))
                want.append(float(loss))
            be += 1
    assert len(got) == len(want) and len(got) >= 1, (got, want, text)
    for g, w in zip(got, want):
        assert abs(g - w) <= 1e-5 * abs(w), (got, want)
