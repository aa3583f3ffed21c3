"""GPU parity tests of the order-free fused FM step (csrc/fm_fused.cu: slot map at upload, gather + RED scatter into the
batch-compact buffer with hot-slot replicas, compact updater) against the CPU oracle, through the C ABI.

The path sums gradients with REDs in arbitrary order (like the reference's own multi-threaded Hogwild mode),
so it is held to PER-STEP parity from identical state: summed logloss 1e-6 .. 2e-6 relative (tolerance stated at each
assert; north-star bar 1e-5), parameters to fp32 re-association noise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def _synth(F, B, seed, with_val=False):
    from lightctr_b200.data import CriteoSynth
    rp, fid, fld, lab = CriteoSynth(F, seed=seed).batch(B)
    val = None
    if with_val:
        val = (0.5 + np.random.default_rng(seed).random(len(fid))).astype(np.float32)
    return rp, fid, fld, lab, val


OPTS = {"adagrad": 0, "ftrl": 1, "adam": 2, "rmsprop": 3, "adadelta": 4}


def _params_close(got, want, tol, threshold_updater):
    """max |got - want| < tol; for updaters with a discontinuity (FTRL: |z| <= lambda1 -> w = 0; Adam: the first steps
    are m / sqrt(v) = +-sqrt(1 - beta), i.e. the SIGN of a gradient sum that may be within rounding of 0) a few
    coordinates may land on the other side of it: at most 1e-5 of them, each by at most one such jump (1e-2)."""
    d = np.abs(got - want)
    if not threshold_updater:
        return float(d.max()) < tol
    return float(np.mean(d > tol)) <= 1e-5 and float(d.max()) < 1e-2


@pytest.mark.parametrize("opt", ["adagrad", "ftrl", "adam", "rmsprop", "adadelta"])
@pytest.mark.parametrize("k,with_val", [(16, False), (8, True), (4, False), (32, True)])
def test_fused_step_vs_oracle(oracle_api, opt, k, with_val):
    """Three consecutive steps on one synthetic Criteo-shaped batch (hot ids of the small-vocabulary fields -> replica
    rows are exercised; k = 4, 8, 16, 32 = every lane layout of the kernel; with and without feature values)."""
    from lightctr_b200 import capi
    F, B = 20000, 700
    rp, fid, fld, lab, val = _synth(F, B, 5 + k, with_val)
    rng = np.random.default_rng(k)
    W0 = (rng.standard_normal(F) * 0.01).astype(np.float32)
    V0 = (rng.standard_normal(F * k) / np.sqrt(k)).astype(np.float32) * np.float32(0.5)
    lr = 0.002 if opt == "rmsprop" else 0.05
    ds = oracle_api.Dataset(rp, fid, fld.astype(np.uint32), val if with_val else np.ones(len(fid), np.float32), lab, F, 0)
    o = oracle_api.FMOracle(ds, k, W0, V0, lr=lr)
    o.opt = opt
    ctx = capi.Context(capi.MODEL_FM, F, k, optimizer=OPTS[opt], deterministic=0, lr=lr)
    ctx.upload_params(W0, V0)
    ctx.upload_batch(0, rp, fid, None, val, lab)
    F1 = F * (k + 1)
    for step in range(3):
        # every step starts from the ORACLE's exact state (parameters + updater state), so only the summation order of
        # this one step differs: observed <= 3e-7 on the loss.  (Left to itself the path drifts like the reference's own
        # Hogwild mode: the sign-like first updater steps amplify fp32 re-association noise to ~1e-5 within 3 steps.)
        if step > 0:
            ctx.upload_params(o.W, o.V)
            ctx.upload_opt_state(o.accum, getattr(o, "s2", np.zeros(F1, np.float32)))
        lg, cg = ctx.train_step(0)
        lo, ao = o.epoch()
        assert _rel(lg, lo) < 1e-6, (opt, k, step, lg, lo)
        assert abs(cg - round(ao * B)) <= 1
        Wg, Vg = ctx.download_params()
        thr = opt in ("ftrl", "adam")
        assert _params_close(Wg, o.W, 2e-5, thr) and _params_close(Vg, o.V, 2e-5, thr), (opt, k, step)
    assert ctx.launch_count() > 0
    ctx.close()


def test_fused_long_rows_and_window_overflow(oracle_api):
    """Rows longer than the kernel's register window (k = 16: 128 entries; k = 8: 256) take the re-gather loop."""
    from lightctr_b200 import capi
    rng = np.random.default_rng(2)
    F, rows = 5000, 64
    for k in (16, 8):
        nn = rng.integers(1, 400, rows)
        nn[0], nn[1], nn[2] = 129, 128, 1
        rp = np.concatenate([[0], np.cumsum(nn)]).astype(np.int64)
        fid = np.concatenate([rng.choice(F, n, replace=False) for n in nn]).astype(np.uint32)
        lab = (rng.random(rows) < 0.3).astype(np.int32)
        val = (0.5 + rng.random(len(fid))).astype(np.float32)
        W0 = (rng.standard_normal(F) * 0.01).astype(np.float32)
        V0 = (rng.standard_normal(F * k) * 0.05).astype(np.float32)
        ds = oracle_api.Dataset(rp, fid, np.zeros(len(fid), np.uint32), val, lab, F, 0)
        o = oracle_api.FMOracle(ds, k, W0, V0)
        ctx = capi.Context(capi.MODEL_FM, F, k, deterministic=0)
        ctx.upload_params(W0, V0)
        ctx.upload_batch(0, rp, fid, None, val, lab)
        lg, _ = ctx.train_step(0)
        lo, _ = o.epoch()
        assert _rel(lg, lo) < 1e-6, (k, lg, lo)
        Wg, Vg = ctx.download_params()
        assert np.max(np.abs(Wg - o.W)) < 2e-5 and np.max(np.abs(Vg - o.V)) < 2e-5
        assert np.max(np.abs(ctx.download_sumvx(0) - o.sumVX)) < 1e-4
        ctx.close()


def test_fused_every_row_shares_one_id(oracle_api):
    """One id present in ALL rows (multiplicity = batch size): the hot-replica path must fold exactly."""
    from lightctr_b200 import capi
    rng = np.random.default_rng(9)
    F, rows, k = 3000, 2048, 16
    per = 6
    fid = np.empty((rows, per), np.uint32)
    fid[:, 0] = 7
    fid[:, 1] = 11 + (np.arange(rows) % 3)
    for r in range(rows):
        fid[r, 2:] = rng.choice(np.arange(100, F), per - 2, replace=False)
    rp = (np.arange(rows + 1) * per).astype(np.int64)
    fid = fid.ravel()
    lab = (rng.random(rows) < 0.4).astype(np.int32)
    W0 = (rng.standard_normal(F) * 0.01).astype(np.float32)
    V0 = (rng.standard_normal(F * k) * 0.1).astype(np.float32)
    ds = oracle_api.Dataset(rp, fid, np.zeros(len(fid), np.uint32), np.ones(len(fid), np.float32), lab, F, 0)
    o = oracle_api.FMOracle(ds, k, W0, V0)
    ctx = capi.Context(capi.MODEL_FM, F, k, deterministic=0)
    ctx.upload_params(W0, V0)
    ctx.upload_batch(0, rp, fid, None, None, lab)
    for step in range(2):
        lg, _ = ctx.train_step(0)
        lo, _ = o.epoch()
        assert _rel(lg, lo) < 2e-6, (step, lg, lo)
    Wg, Vg = ctx.download_params()
    # id 7 sums 2048 terms in a different order than the CPU's row order: 1e-4 absolute on a coordinate of magnitude ~0.1
    assert np.max(np.abs(Wg - o.W)) < 1e-4 and np.max(np.abs(Vg - o.V)) < 1e-4
    ctx.close()


def test_fused_c2_shape_one_step(oracle_api):
    """BASELINE configs[1] at its exact shape -- FM k=16, 1 M features, 39 fields, batch 4096, Adagrad -- one step of the
    benched kernels against the oracle on the SAME batch bench.py times (same generator and seed)."""
    from lightctr_b200 import capi
    from lightctr_b200.data import BASE_SEED, CriteoSynth
    F, k, B = 1_000_000, 16, 4096
    rp, fid, fld, lab = CriteoSynth(F, seed=BASE_SEED).batch(B)
    rng = np.random.default_rng(1)
    W0 = (rng.standard_normal(F) * 0.01).astype(np.float32)
    V0 = (rng.standard_normal(F * k) / 4).astype(np.float32)
    ds = oracle_api.Dataset(rp, fid, fld.astype(np.uint32), np.ones(len(fid), np.float32), lab, F, 0)
    o = oracle_api.FMOracle(ds, k, W0, V0)
    ctx = capi.Context(capi.MODEL_FM, F, k, deterministic=0)
    ctx.upload_params(W0, V0)
    ctx.upload_batch(0, rp, fid, None, None, lab)
    lg, cg = ctx.train_step(0)
    lo, ao = o.epoch()
    assert _rel(lg, lo) < 1e-6, (lg, lo)
    assert abs(cg - round(ao * B)) <= 1
    Wg, Vg = ctx.download_params()
    assert np.max(np.abs(Wg - o.W)) < 2e-5 and np.max(np.abs(Vg - o.V)) < 2e-5
    s1, _ = ctx.download_opt_state()
    assert np.allclose(s1, o.accum, rtol=1e-4, atol=1e-9)
    # second step, again from the oracle's exact state (hot replicas and the compact buffer must have been re-zeroed
    # exactly by the first one; left to itself the path drifts by ~2e-5 here, like the reference's Hogwild mode)
    ctx.upload_params(o.W, o.V)
    ctx.upload_opt_state(o.accum)
    lg2, _ = ctx.train_step(0)
    lo2, _ = o.epoch()
    # After one sign-like Adagrad step from a random init many rows are saturated (mean loss 8.5 per row), and the
    # reference's Sigmoid::forward is DISCONTINUOUS at its clamps (activations.h:65-72: x < -16 -> 1e-7 but
    # sigmoid(-16) = 1.125e-7, a 0.118 jump of the row's loss): a row whose logit lies within fp32 re-association noise
    # of +-16 lands on either side.  So: every unsaturated row's prediction to 1e-5, and the summed loss to within a
    # handful of such jumps.
    pg, po = ctx.download_pred(0), o.pred
    mid = (po > 1e-5) & (po < 1 - 1e-5)
    assert mid.sum() > 100
    assert np.max(np.abs(pg[mid] - po[mid]) / po[mid]) < 1e-4
    assert abs(lg2 - lo2) < 8 * 0.118 + 1e-5 * lo2, (lg2, lo2)
    ctx.close()


def test_fused_sub_ranges_of_a_resident_slot(oracle_api):
    """lctr_train_step on row sub-ranges of a resident slot (the slot map covers the whole slot; a step touches a subset
    of its key set): two half steps == the oracle run on the two halves as separate datasets."""
    from lightctr_b200 import capi
    F, B, k = 8000, 600, 16
    rp, fid, fld, lab, _ = _synth(F, B, 31)
    rng = np.random.default_rng(6)
    W0 = np.zeros(F, np.float32)
    V0 = (rng.standard_normal(F * k) / 4).astype(np.float32)
    ctx = capi.Context(capi.MODEL_FM, F, k, deterministic=0)
    ctx.upload_params(W0, V0)
    ctx.upload_batch(0, rp, fid, None, None, lab)
    W, V, acc = W0.copy(), V0.copy(), np.zeros(F * (k + 1), np.float32)
    for (rb, re) in ((0, 250), (250, 600)):
        lg, _ = ctx.train_step(0, rb, re)
        sub_rp = (rp[rb:re + 1] - rp[rb]).astype(np.int64)
        sl = slice(rp[rb], rp[re])
        ds = oracle_api.Dataset(sub_rp, fid[sl], fld[sl].astype(np.uint32), np.ones(rp[re] - rp[rb], np.float32), lab[rb:re], F, 0)
        o = oracle_api.FMOracle(ds, k, W, V)
        o.accum[:] = acc
        lo, _ = o.epoch()
        W, V, acc = o.W.copy(), o.V.copy(), o.accum.copy()
        assert _rel(lg, lo) < 2e-6, (rb, lg, lo)
    Wg, Vg = ctx.download_params()
    assert np.max(np.abs(Wg - W)) < 2e-5 and np.max(np.abs(Vg - V)) < 2e-5
    ctx.close()


def test_dependent_launch_of_the_updater_changes_nothing(oracle_api, monkeypatch):
    """The compact updater is launched programmatically dependent on the gradient kernel (it starts early and synchronises on
    the gradient kernel's completion itself, fm_fused.cu: apply_go).  Many short steps back to back -- where an updater that
    read G too early, or a gradient kernel that read parameters mid-update, would show -- with the attribute on and off
    (LCTR_PDL), from identical state: the order-free path reproduces itself to re-association noise per step."""
    from lightctr_b200 import capi
    F, k, B = 20000, 16, 2048
    rp, fid, fld, lab, _ = _synth(F, B, 77)
    rng = np.random.default_rng(8)
    W0 = (rng.standard_normal(F) * 0.01).astype(np.float32)
    V0 = (rng.standard_normal(F * k) * 0.05).astype(np.float32)
    curves = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("LCTR_PDL", flag)
        ctx = capi.Context(capi.MODEL_FM, F, k, deterministic=0, lr=0.02)
        ctx.upload_params(W0, V0)
        ctx.upload_batch(0, rp, fid, None, None, lab)
        curves[flag] = [ctx.train_step(0)[0] for _ in range(150)]
        W, V = ctx.download_params()
        curves[flag + "p"] = (W, V)
        ctx.close()
    a, b = np.array(curves["0"]), np.array(curves["1"])
    assert np.all(np.isfinite(a)) and b[-1] < b[0]
    assert np.max(np.abs(a - b) / np.abs(a)) < 2e-4, np.max(np.abs(a - b) / np.abs(a))  # 150 chaotic steps apart, not a race
    assert np.max(np.abs(a[:10] - b[:10]) / np.abs(a[:10])) < 5e-6
    ds = oracle_api.Dataset(rp, fid, fld.astype(np.uint32), np.ones(len(fid), np.float32), lab, F, 0)
    o = oracle_api.FMOracle(ds, k, W0, V0, lr=0.02)
    lo = [o.epoch()[0] for _ in range(10)]
    assert np.max(np.abs(b[:10] - np.array(lo)) / np.array(lo)) < 1e-5


def test_streamed_pipeline_three_deep(oracle_api):
    """lctr_train_batch_async with LCTR_PIPE_DEPTH = 3 tickets in flight (step t computing, batch t+1 building its slot map,
    batch t+2 on the copy engine): every step's loss equals the synchronous lctr_train_batch sequence on a second context (the
    order-free path reproduces itself to re-association noise), a fourth outstanding ticket is refused."""
    from lightctr_b200 import capi
    F, k, B, NB, steps = 30000, 16, 1024, 5, 17
    batches = []
    for i in range(NB):
        rp, fid, fld, lab, _ = _synth(F, B, 100 + i)
        batches.append((rp, fid.astype(np.uint32), None, None, lab.astype(np.int32)))
    rng = np.random.default_rng(9)
    W0 = (rng.standard_normal(F) * 0.01).astype(np.float32)
    V0 = (rng.standard_normal(F * k) * 0.05).astype(np.float32)
    a = capi.Context(capi.MODEL_FM, F, k, deterministic=0)
    b = capi.Context(capi.MODEL_FM, F, k, deterministic=0)
    for c in (a, b):
        c.upload_params(W0, V0)
    want = [a.train_batch(*batches[i % NB])[0] for i in range(steps)]
    got, pending = [], []
    for i in range(steps):
        pending.append(b.train_batch_async(*batches[i % NB]))
        if len(pending) == capi.PIPE_DEPTH:
            if i == capi.PIPE_DEPTH - 1:  # three in flight: one more must be refused, not queued over a live slot
                with pytest.raises(capi.LctrError):
                    b.train_batch_async(*batches[0])
            got.append(b.wait(pending.pop(0))[0])
    while pending:
        got.append(b.wait(pending.pop(0))[0])
    want, got = np.array(want), np.array(got)
    assert np.max(np.abs(want - got) / np.abs(want)) < 2e-5, np.max(np.abs(want - got) / np.abs(want))
    Wa, Va = a.download_params()
    Wb, Vb = b.download_params()
    assert np.max(np.abs(Wa - Wb)) < 1e-4 and np.max(np.abs(Va - Vb)) < 1e-4
    a.close(); b.close()
