"""Host-side mirrors of the reference trainer classes, driving the CUDA path through the C ABI.

Same constructor arguments, method names and public members as the reference so parity tests read
like its own `main.cpp` (main.cpp:144-162,228-233,253):

    Train_FM_Algo(dataPath, epoch_cnt, factor_cnt)                       train/train_fm_algo.h:23
    Train_FFM_Algo(dataPath, epoch_cnt, factor_cnt, field_cnt)           train/train_ffm_algo.h:25
    Train_NFM_Algo(dataPath, epoch_cnt, factor_cnt, hidden_layer_size)   train/train_nfm_algo.h:21
    FM_Predict(fm, testDataPath, with_valid_label).Predict(savePath)     predict/fm_predict.h:19

The device boundary sits at Train() entry/exit: data and parameters are uploaded once,
host W / V / sumVX are refreshed when Train() returns because FM_Predict and saveModel read them
(fm_predict.cpp:25-32, fm_algo_abst.h:118-131).  (The C++ twin of this file is lightctr_b200/host/.)
"""
import os
import sys

import numpy as np

from . import capi
from .hostrng import GlibcRand


class GradientUpdater:
    """The reference's process-global statics (util/gradientUpdater.h:36-42, main.cpp:64-73)."""
    minibatch_size = 50
    learning_rate = 0.05
    ema_rate = 0.99
    sparse_rate = 0.8
    lambdaL2 = 0.001
    lambdaL1 = 1e-5
    bTraining = True


class MomentumUpdater:
    momentum = 0.8
    momentum_adam2 = 0.999


_rng = GlibcRand(1)


def srand(seed):
    """main.cpp:78 does srand(time(NULL)); tests fix the seed."""
    global _rng
    _rng = GlibcRand(seed)


class FM_Algo_Abst:
    """fm_algo_abst.h:37-172."""

    def __init__(self, dataPath, factor_cnt, field_cnt=0, feature_cnt=0, optimizer=capi.OPT_ADAGRAD, device=0,
                 deterministic=True):
        ds = capi.load_libffm(dataPath, field_cnt, feature_cnt)  # loadDataRow, :70-107
        self.data = ds
        self.feature_cnt, self.field_cnt, self.factor_cnt = ds.feature_cnt, ds.field_cnt, factor_cnt
        self.dataRow_cnt = ds.rows
        self.L2Reg_ratio = 0.001
        self.optimizer = optimizer
        self.device = device
        # deterministic: gradients summed in ascending row order (the reference's canonical proc_cnt=1 order) through
        # the feature-major view; False selects the RED-based scatter used for streamed batches
        self.deterministic = deterministic
        # init(), :53-68
        self.W = np.zeros(self.feature_cnt, np.float32)
        n = self.feature_cnt * factor_cnt * (self.field_cnt if self.field_cnt > 0 else 1)
        self.V = _rng.gauss_fill(n, factor_cnt)
        self.sumVX = None
        self._ctx = None
        self.loss_curve, self.acc_curve = [], []

    # -- device plumbing ---------------------------------------------------------------------------
    def _make_ctx(self, model, **kw):
        self._ctx = capi.Context(model, self.feature_cnt, self.factor_cnt, self.field_cnt if model == capi.MODEL_FFM else 0,
                                 optimizer=self.optimizer, lr=GradientUpdater.learning_rate, l2=self.L2Reg_ratio,
                                 momentum=MomentumUpdater.momentum, momentum_adam2=MomentumUpdater.momentum_adam2,
                                 device=self.device, deterministic=1 if self.deterministic else 0, **kw)
        self._ctx.upload_params(self.W, self.V)
        self._ctx.upload_dataset(0, self.data)

    def saveModel(self, epoch):
        """fm_algo_abst.h:109-135: ./output/model_epoch_<n>.txt, `fid:w ` for non-zero W then one V line per fid."""
        os.makedirs("./output", exist_ok=True)
        k = self.factor_cnt
        with open("./output/model_epoch_%d.txt" % epoch, "w") as md:
            md.write("".join("%d:%s " % (f, _cxx_float(self.W[f])) for f in np.nonzero(self.W)[0]) + "\n")
            V = self.V.reshape(self.feature_cnt, -1)
            for f in range(self.feature_cnt):
                md.write("%d:" % f + "".join(_cxx_float(v) + " " for v in V[f, :k]) + "\n")


def _cxx_float(v):
    """operator<<(ostream, float) with the default precision 6 (%g)."""
    return "%g" % float(v)


class Train_FM_Algo(FM_Algo_Abst):
    def __init__(self, dataPath, epoch_cnt, factor_cnt, **kw):
        super().__init__(dataPath, factor_cnt, **kw)
        assert self.feature_cnt != 0
        self.epoch_cnt = epoch_cnt
        self.sumVX = np.zeros(self.dataRow_cnt * factor_cnt, np.float32)  # init(), train_fm_algo.cpp:19-21

    def Train(self):
        GradientUpdater.bTraining = True
        GradientUpdater.minibatch_size = self.dataRow_cnt  # train_fm_algo.cpp:38 (left set afterwards, like the reference)
        if self._ctx is None:
            self._make_ctx(capi.MODEL_FM)
        for i in range(self.epoch_cnt):
            loss, correct = self._ctx.train_step(0)  # full batch == one epoch (:44-57)
            acc = float(np.float32(correct) / np.float32(self.dataRow_cnt))
            print("Epoch %d Train Loss = %f Accuracy = %f" % (i, loss, acc))
            self.loss_curve.append(loss)
            self.acc_curve.append(acc)
        self.W, self.V = self._ctx.download_params()
        self.sumVX = self._ctx.download_sumvx(0)
        GradientUpdater.bTraining = False


class Train_FFM_Algo(FM_Algo_Abst):
    def __init__(self, dataPath, epoch_cnt, factor_cnt, field_cnt, **kw):
        super().__init__(dataPath, factor_cnt, field_cnt, **kw)
        assert self.feature_cnt != 0
        self.epoch = epoch_cnt
        print("Training FFM")

    def Train(self):
        GradientUpdater.bTraining = True
        GradientUpdater.minibatch_size = self.dataRow_cnt
        if self._ctx is None:
            self._make_ctx(capi.MODEL_FFM)
        for i in range(self.epoch):
            loss, correct = self._ctx.train_step(0)
            acc = float(np.float32(correct) / np.float32(self.dataRow_cnt))
            print("Epoch %d Train Loss = %f Accuracy = %f" % (i, loss, acc))
            self.loss_curve.append(loss)
            self.acc_curve.append(acc)
        self.W, self.V = self._ctx.download_params()
        GradientUpdater.bTraining = False


class Fully_Conn_Layer:
    """Host shadow of one layer (train/layer/fullyconnLayer.h:36-61): holds init values + the dropout mask, both
    drawn from the reference's rand() stream in the reference's order."""

    def __init__(self, n_in, n_out):
        self.n_in, self.n_out = n_in, n_out
        self.bias = np.zeros(n_out, np.float32)
        self.mask = np.zeros(n_out, np.float32)
        self.weight = np.zeros((n_out, n_in), np.float32)
        p = float(np.float32(GradientUpdater.sparse_rate))
        for i in range(n_out):  # :48-54
            self.mask[i] = 1.0 if _rng.sample_binary(p) else 0.0
            draws = _rng.rand_array(n_in)
            self.weight[i] = (draws / (2147483647 + 1.0) - float(np.float32(0.5))).astype(np.float32)

    def resample_mask(self):  # applyBatchGradient, :200-202
        p = float(np.float32(GradientUpdater.sparse_rate))
        draws = _rng.rand_array(self.n_out)
        self.mask = ((draws / (2147483647 + 1.0)) < p).astype(np.float32)


class Train_NFM_Algo(FM_Algo_Abst):
    def __init__(self, dataPath, epoch_cnt, factor_cnt, hidden_layer_size, activation=capi.ACT_SIGMOID, **kw):
        super().__init__(dataPath, factor_cnt, **kw)
        assert self.feature_cnt != 0
        self.epoch = epoch_cnt
        self.hidden = list(hidden_layer_size) if isinstance(hidden_layer_size, (list, tuple)) else [hidden_layer_size]
        self.activation = activation
        self.batch_size = GradientUpdater.minibatch_size  # init(), train_nfm_algo.cpp:13
        self.sumVX = np.zeros(self.dataRow_cnt * factor_cnt, np.float32)
        dims = [factor_cnt] + self.hidden + [1]
        self.layers = [Fully_Conn_Layer(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]  # :21-27

    def Train(self):
        GradientUpdater.bTraining = True
        if self._ctx is None:
            # the updater divides by the GLOBAL minibatch size even for the short tail batch (:161-169)
            self._make_ctx(capi.MODEL_NFM, hidden=self.hidden, activation=self.activation,
                           minibatch_size=GradientUpdater.minibatch_size, csc_row_block=self.batch_size)
            for l, L in enumerate(self.layers):
                self._ctx.mlp_upload(l, L.weight, L.bias)
                self._ctx.mlp_set_mask(l, L.mask)
        n = self.dataRow_cnt
        for i in range(self.epoch):
            loss = np.float32(0)
            correct = 0
            for p in range((n + self.batch_size - 1) // self.batch_size):
                rb = p * self.batch_size
                l, c = self._ctx.train_step(0, rb, min(rb + self.batch_size, n))
                loss = np.float32(loss + np.float32(l))
                correct += int(c)
                for li, L in enumerate(self.layers):  # applyBatchGradient re-draws every layer's mask
                    L.resample_mask()
                    self._ctx.mlp_set_mask(li, L.mask)
            acc = 1.0 * correct / n
            print("Epoch %d loss = %f accuracy = %f" % (i, loss, acc))
            self.loss_curve.append(float(loss))
            self.acc_curve.append(acc)
        self.W, self.V = self._ctx.download_params()
        self.sumVX = self._ctx.download_sumvx(0)
        for l, L in enumerate(self.layers):
            w, b = self._ctx.mlp_download(l, L.n_in, L.n_out)
            L.weight, L.bias = w.reshape(L.n_out, L.n_in), b
        GradientUpdater.bTraining = False


class FM_Predict:
    """predict/fm_predict.{h,cpp}.  quirks=True reproduces the reference bit-for-bit in its indexing: the first
    feature of every test row is dropped, fids >= train feature_cnt are dropped (:117-126), and the FM branch
    adds 0.5*|sumVX_train[rid]|^2 of the TRAINING row with the same index (:27-32)."""

    def __init__(self, fm, testDataPath, with_valid_label=True, quirks=True):
        self.fm, self.quirks = fm, quirks
        ds = capi.load_libffm(testDataPath, 0, 0)
        keep_rows, fid, fld, val, rp, lab = [], [], [], [], [0], []
        F = fm.feature_cnt
        for r in range(ds.rows):
            b, e = ds.row_ptr[r], ds.row_ptr[r + 1]
            if quirks:
                b += 1
            sel = np.arange(b, e)
            sel = sel[ds.fid[sel] < F]
            if len(sel) == 0:
                continue
            fid.append(ds.fid[sel]); fld.append(ds.field[sel]); val.append(ds.val[sel])
            rp.append(rp[-1] + len(sel))
            keep_rows.append(r)
        self.test = capi.HostDataset(np.array(rp), np.concatenate(fid), np.concatenate(fld), np.concatenate(val),
                                     ds.label[:len(rp) - 1] if quirks else ds.label[keep_rows], F, fm.field_cnt)
        self.test_dataRow_cnt = self.test.rows
        self.pCTR = None

    def Predict(self, savePath=""):
        fm = self.fm
        ctx = fm._ctx
        ctx.upload_dataset(1, self.test)
        is_ffm = fm.sumVX is None
        ans = ctx.predict(1, quirk_sumvx_slot=(0 if (self.quirks and not is_ffm) else -1))
        self.pCTR = ans
        y = self.test.label
        loss = np.float32(0)
        for p, t in zip(ans, y):  # fm_predict.cpp:63-72 (float accumulator, double term)
            term = -float(np.log(np.float32(p))) if t == 1 else -np.log(1.0 - float(p))
            loss = np.float32(float(loss) + term)
        correct = int(np.sum(((ans > 0.5) & (y == 1)) | ((ans < 0.5) & (y == 0))))
        self.loss, self.correct = float(loss), correct
        self.auc = auc_evaluator(ans, y)
        print("total log likelihood = %s correct = %s auc = %.4f" % (_cxx_float(loss),
              "%.5g" % (np.float32(correct) / np.float32(self.test_dataRow_cnt)), self.auc))
        if savePath:
            with open(savePath, "w") as md:
                for v in ans:
                    md.write(_cxx_float(v) + "\n")
        return ans


def auc_evaluator(pctr, label):
    """AucEvaluator (util/evaluator.h:51-104): 2^24-1 buckets, trapezoid sweep from the top bucket down, f32."""
    k = np.float32((1 << 24) - 1)
    idx = (pctr.astype(np.float32) * k).astype(np.int64)
    pos = np.bincount(idx[label == 1], minlength=(1 << 24))
    neg = np.bincount(idx[label != 1], minlength=(1 << 24))
    nz = np.nonzero(pos + neg)[0][::-1]
    totPos = totNeg = np.float32(0)
    auc = np.float32(0)
    for i in nz:
        pp, nn = totPos, totNeg
        totPos = np.float32(totPos + np.float32(pos[i]))
        totNeg = np.float32(totNeg + np.float32(neg[i]))
        dx = totNeg - nn if totNeg > nn else nn - totNeg
        auc = np.float32(auc + np.float32(float(np.float32(dx * (totPos + pp))) / 2.0))
    if totPos > 0 and totNeg > 0:
        return float(np.float32(np.float32(auc / totPos) / totNeg))
    return 0.0
