"""GPU parity tests at the shapes of BASELINE.json's configs (the shapes every published number is measured on) and of the
remaining updater x model combinations in the exact-order mode.  One step from identical state against the CPU oracle;
the bar is north_star's 1e-5 relative on the summed logloss, stated again at each assert.

C2 (FM k=16, 1 M features, batch 4096) lives in tests/test_fm_fused_gpu.py::test_fused_c2_shape_one_step;
C5 (two ranks) in tests/test_dist.py."""
import numpy as np
import pytest

from golden_util import load_csr

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def test_c3_shape_ffm_ftrl_one_step(oracle_api):
    """BASELINE configs[2]: FFM k=4, 39 fields, 1 M features, batch 8192, FTRL -- the benched kernels (fused FFM step + sparse
    FTRL apply) on the batch bench.py --workload ffm_c3 times, against the oracle's pair-loop restatement."""
    from lightctr_b200 import capi
    from lightctr_b200.data import BASE_SEED, CriteoSynth
    F, k, Fc, B = 1_000_000, 4, 39, 8192
    rp, fid, fld, lab = CriteoSynth(F, seed=BASE_SEED).batch(B)
    rng = np.random.default_rng(3)
    W0 = (rng.standard_normal(F) * 0.01).astype(np.float32)
    V0 = (rng.standard_normal(F * Fc * k, dtype=np.float32) * np.float32(0.5))  # N(0,1)/sqrt(k)
    ds = oracle_api.Dataset(rp, fid, fld.astype(np.uint32), np.ones(len(fid), np.float32), lab, F, Fc)
    o = oracle_api.FFMOracle(ds, k, W0, V0, optimizer="ftrl")
    ctx = capi.Context(capi.MODEL_FFM, F, k, Fc, optimizer=capi.OPT_FTRL, deterministic=0)
    ctx.upload_params(W0, V0)
    ctx.upload_batch(0, rp, fid, fld, None, lab)
    lg, cg = ctx.train_step(0)
    lo, ao = o.epoch()
    assert _rel(lg, lo) < 1e-5, (lg, lo)
    assert abs(cg - round(ao * B)) <= 1
    Wg, Vg = ctx.download_params()
    # FTRL's hard threshold (|z| <= lambda1 -> w = 0, gradientUpdater.h:262-264): a coordinate whose z lies within fp32
    # re-association noise of lambda1 may land on the other side -- at most 1e-5 of the coordinates, everything else 1e-4
    for got, want in ((Wg, o.W), (Vg, o.V)):
        d = np.abs(got - want)
        assert float(np.mean(d > 1e-4)) <= 1e-5, float(np.mean(d > 1e-4))
    # second step: the updated state on both sides
    lg2, _ = ctx.train_step(0)
    lo2, _ = o.epoch()
    assert _rel(lg2, lo2) < 1e-5, (lg2, lo2)
    ctx.close()


def test_c4_shape_nfm_chain_fp32_one_step(oracle_api):
    """BASELINE configs[3] in its parity precision: NFM k=16 + Fully_Conn_Layer chain 16 -> 256 -> 128 -> 64 -> 1, fp32
    reference-order dense layers, one minibatch of 16 384 rows, against the oracle's chain (fullyconnLayer.h:80-206)."""
    from lightctr_b200 import capi
    from lightctr_b200.data import BASE_SEED, CriteoSynth
    F, k, B = 1_000_000, 16, 16384
    hidden = [256, 128, 64]
    rp, fid, fld, lab = CriteoSynth(F, seed=BASE_SEED).batch(B)
    rng = np.random.default_rng(4)
    W0 = (rng.standard_normal(F) * 0.01).astype(np.float32)
    V0 = (rng.standard_normal(F * k, dtype=np.float32) * np.float32(0.25))
    ds = oracle_api.Dataset(rp, fid, fld.astype(np.uint32), np.ones(len(fid), np.float32), lab, F, 0)
    o = oracle_api.NFMOracle(ds, k, hidden, W=W0, V=V0, batch_size=B, minibatch=B)
    dims = [k] + hidden + [1]
    layers = []
    for l in range(len(dims) - 1):
        w = ((rng.random(dims[l] * dims[l + 1], dtype=np.float32) - np.float32(0.5)) * np.float32(2.0 / np.sqrt(dims[l]))).astype(np.float32)
        o.mlp.arrays("weight", l)[:] = w
        o.mlp.arrays("bias", l)[:] = 0
        o.mlp.arrays("mask", l)[:] = 1.0
        layers.append(w)
    ctx = capi.Context(capi.MODEL_NFM, F, k, hidden=tuple(hidden), mlp_precision=capi.MLP_FP32, minibatch_size=B, deterministic=0)
    ctx.upload_params(W0, V0)
    for l, w in enumerate(layers):
        ctx.mlp_upload(l, w, np.zeros(dims[l + 1], np.float32))
    ctx.upload_batch(0, rp, fid, None, None, lab)
    lg, cg = ctx.train_step(0)
    lo, ao = o.epoch()
    assert _rel(lg, lo) < 1e-5, (lg, lo)
    for l in range(len(dims) - 1):
        w, b = ctx.mlp_download(l, dims[l], dims[l + 1])
        # dense Adagrad's first step is lr * g / sqrt(g^2 + 1e-7): sign-like where |g| is tiny
        d = np.abs(w - o.mlp.arrays("weight", l))
        assert float(np.mean(d > 1e-4)) <= 1e-3 and float(d.max()) < 0.11, (l, float(d.max()))
    Wg, Vg = ctx.download_params()
    assert np.max(np.abs(Wg - o.W)) < 1e-4
    d = np.abs(Vg - o.V)
    assert float(np.mean(d > 1e-4)) <= 1e-5
    ctx.close()


@pytest.mark.parametrize("opt", ["ftrl", "adam"])
def test_fm_exact_order_mode_with_ftrl_and_adam(oracle_api, opt):
    """FTRLUpdater / AdamUpdater_Num as the FM trainer's `updater` member in the exact-order mode (cfg.deterministic = 1:
    gradients summed in ascending row order like the reference's canonical single-thread run): 6 epochs on the reference's
    own train_sparse.csv, loss curve within 1e-5 of the oracle -- the bar the RED-based FFM tests could only hold at 5e-5."""
    from lightctr_b200 import capi
    ds = load_csr("train_sparse_csr.npz", field_cnt=68)
    k = 8
    W0, V0 = oracle_api.init_params(1, ds.feature_cnt, k)
    o = oracle_api.FMOracle(ds, k, W0, V0)
    o.opt = opt
    code = {"ftrl": capi.OPT_FTRL, "adam": capi.OPT_ADAM}[opt]
    ctx = capi.Context(capi.MODEL_FM, ds.feature_cnt, k, optimizer=code, deterministic=1)
    ctx.upload_params(W0, V0)
    ctx.upload_batch(0, ds.row_ptr, ds.fid, None, None, ds.label)
    for e in range(6):
        lg = ctx.train_step(0)[0]
        lo, _ = o.epoch()
        assert _rel(lg, lo) < 1e-5, (opt, e, lg, lo)
    W, V = ctx.download_params()
    assert np.max(np.abs(W - o.W)) < 1e-5 and np.max(np.abs(V - o.V)) < 1e-5
    s1, s2 = ctx.download_opt_state()
    assert np.max(np.abs(s1 - o.accum)) < 1e-5 * max(1.0, float(np.max(np.abs(o.accum))))
    assert np.max(np.abs(s2 - o.s2)) < 1e-5 * max(1.0, float(np.max(np.abs(o.s2))))
    ctx.close()


@pytest.mark.parametrize("opt", ["ps_sgd", "ps_adagrad", "ps_dcasgd", "ps_dcasgda"])
def test_fm_parameter_server_update_rules(oracle_api, opt):
    """The ParamServer's own per-coordinate rules (distribut/paramserver.h:232-300: SGD -- the default --, Adagrad, DCASGD,
    DCASGDA with their mutating Value arithmetic) as the trainer's updater, exact-order mode, against the oracle's
    restatement (pinned: tests/test_oracle_wnd_pin_cpu.py reproduces the reference cluster's curve for each rule).  The learning rate is raised so that
    the SGD-type rules (step = lr * g / minibatch) move the loss."""
    from lightctr_b200 import capi
    ds = load_csr("train_sparse_csr.npz", field_cnt=68)
    k = 8
    lr = {"ps_sgd": 5.0, "ps_adagrad": 0.05, "ps_dcasgd": 5.0, "ps_dcasgda": 0.01}[opt]
    W0, V0 = oracle_api.init_params(1, ds.feature_cnt, k)
    o = oracle_api.FMOracle(ds, k, W0, V0, lr=lr)
    o.opt = opt
    code = {"ps_sgd": capi.OPT_PS_SGD, "ps_adagrad": capi.OPT_PS_ADAGRAD, "ps_dcasgd": capi.OPT_PS_DCASGD,
            "ps_dcasgda": capi.OPT_PS_DCASGDA}[opt]
    ctx = capi.Context(capi.MODEL_FM, ds.feature_cnt, k, optimizer=code, deterministic=1, lr=lr)
    ctx.upload_params(W0, V0)
    ctx.upload_batch(0, ds.row_ptr, ds.fid, None, None, ds.label)
    losses = []
    for e in range(5):
        lg = ctx.train_step(0)[0]
        lo, _ = o.epoch()
        losses.append(lo)
        assert _rel(lg, lo) < 1e-5, (opt, e, lg, lo)
    assert np.isfinite(losses[-1]) and abs(losses[-1] - losses[0]) > 1e-6 * abs(losses[0])  # the rule moves the parameters
    W, V = ctx.download_params()
    assert np.max(np.abs(W - o.W)) < 1e-5 and np.max(np.abs(V - o.V)) < 1e-5
    s1, s2 = ctx.download_opt_state()
    if opt in ("ps_adagrad", "ps_dcasgda"):
        assert np.max(np.abs(s1 - o.accum)) <= 1e-5 * max(1.0, float(np.max(np.abs(o.accum))))
    if opt in ("ps_dcasgd", "ps_dcasgda"):
        assert np.max(np.abs(s2 - o.s2)) < 1e-5
    ctx.close()


@pytest.mark.parametrize("Fc,k,max_n,with_val", [(6, 4, 40, True),      # 6 slots: one pass, 26 shadow lanes
                                                  (10, 8, 300, True),    # rows longer than the 128-entry index window
                                                  (33, 12, 90, False),   # 99 slots: four passes, 3 parts per field
                                                  (64, 4, 70, False)])   # the largest field count the kernel takes
def test_ffm_warp_kernel_edge_shapes(oracle_api, Fc, k, max_n, with_val):
    """ffm_warp.cu beyond the benchmark shapes: every pass count, values != 1, rows of 0 / 1 / hundreds of entries (the
    whole-sample index window is 128 entries: longer rows reload it in the gradient phase), fields missing from a row,
    several entries of one field that are NOT adjacent (the per-field register run is flushed and T[a] re-opened), rows whose
    prediction equals the label exactly are left alone.  One step from identical state against the oracle's pair loop."""
    from lightctr_b200 import capi
    rng = np.random.default_rng(Fc * 100 + k)
    F, B = 3000, 257
    rp, fid, fld, val = [0], [], [], []
    for r in range(B):
        n = 0 if r % 37 == 5 else (1 if r % 41 == 7 else int(rng.integers(2, max_n + 1)))
        fs = rng.integers(0, Fc, n)                     # unordered: the same field recurs at distance
        if r % 3 == 0 and n > 4:
            fs[:] = np.sort(fs)                          # and libffm-style sorted rows
        ids = rng.choice(F, n, replace=False)
        fid += list(ids); fld += list(fs); val += list((0.5 + rng.random(n)) if with_val else np.ones(n)); rp.append(len(fid))
    rp = np.array(rp, np.int64); fid = np.array(fid, np.uint32); fld = np.array(fld, np.uint16)
    val = np.array(val, np.float32); lab = (rng.random(B) < 0.4).astype(np.int32)
    W0 = (rng.standard_normal(F) * 0.05).astype(np.float32)
    V0 = (rng.standard_normal(F * Fc * k) * 0.1).astype(np.float32)
    ds = oracle_api.Dataset(rp, fid, fld.astype(np.uint32), val, lab, F, Fc)
    o = oracle_api.FFMOracle(ds, k, W0, V0)
    ctx = capi.Context(capi.MODEL_FFM, F, k, Fc, deterministic=0)
    ctx.upload_params(W0, V0)
    ctx.upload_batch(0, rp, fid, fld, val if with_val else None, lab)
    for step in range(2):
        lg, cg = ctx.train_step(0)
        lo, ao = o.epoch()
        assert _rel(lg, lo) < 1e-5, (step, lg, lo)     # north-star bar on the summed logloss
        assert abs(cg - round(ao * B)) <= 1
        Wg, Vg = ctx.download_params()
        assert np.max(np.abs(Wg - o.W)) < 2e-5 and np.max(np.abs(Vg - o.V)) < 2e-5
        ctx.upload_params(o.W, o.V)                      # per-step parity from identical state (order-free path)
        ctx.upload_opt_state(o.s1)
    ctx.close()
