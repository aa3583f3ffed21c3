"""ctypes binding of the C ABI in include/lightctr_b200.h (lightctr_b200/lib/liblightctr_b200.so).

This is plumbing for the Python tests and bench.py; the drop-in host side for the reference's C++
callers is lightctr_b200/host/*.h.  There is NO CPU fallback: if the shared library is missing or
no CUDA device is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liblightctr_b200.so")

MODEL_FM, MODEL_FFM, MODEL_NFM, MODEL_WND = 1, 2, 3, 4
OPT_ADAGRAD, OPT_FTRL, OPT_ADAM, OPT_RMSPROP, OPT_ADADELTA = 0, 1, 2, 3, 4
PIPE_DEPTH = 3  # LCTR_PIPE_DEPTH: tickets of train_batch_async that may be outstanding
OPT_PS_SGD, OPT_PS_ADAGRAD, OPT_PS_DCASGD, OPT_PS_DCASGDA = 5, 6, 7, 8
ACT_SIGMOID, ACT_TANH = 0, 1
MLP_FP32, MLP_BF16 = 0, 1
MAX_LAYERS = 8
ABI_VERSION = 1

# every symbol include/lightctr_b200.h declares (tests check the .so exports each of them)
SYMBOLS = [
    "lctr_last_error", "lctr_abi_version", "lctr_create", "lctr_destroy", "lctr_sync", "lctr_upload_params",
    "lctr_download_params", "lctr_fill_params", "lctr_download_opt_state", "lctr_upload_opt_state", "lctr_upload_batch",
    "lctr_train_step", "lctr_train_batch", "lctr_train_batch_async", "lctr_wait", "lctr_predict", "lctr_download_sumvx", "lctr_download_pred",
    "lctr_mlp_forward", "lctr_mlp_backward", "lctr_mlp_apply", "lctr_mlp_upload", "lctr_mlp_download", "lctr_mlp_set_mask", "lctr_mlp_download_grad", "lctr_set_dense_allreduce", "lctr_save_checkpoint", "lctr_load_checkpoint",
    "lctr_save_dataset_bin", "lctr_load_dataset_bin", "lctr_eval", "lctr_upload_pred", "lctr_ipc_export", "lctr_ipc_import",
    "lctr_dense_grad_buffer", "lctr_device_bytes", "lctr_load_libffm", "lctr_free_dataset", "lctr_launch_count", "lctr_stream", "lctr_profile", "lctr_profile_read",
]


class Cfg(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("model", C.c_int32), ("optimizer", C.c_int32), ("device", C.c_int32),
                ("feature_cnt", C.c_uint64), ("field_cnt", C.c_uint32), ("factor_cnt", C.c_uint32),
                ("learning_rate", C.c_float), ("l2_reg", C.c_float), ("minibatch_size", C.c_uint64),
                ("momentum", C.c_float), ("momentum_adam2", C.c_float), ("ftrl_alpha", C.c_float),
                ("ftrl_beta", C.c_float), ("ftrl_lambda1", C.c_float), ("ftrl_lambda2", C.c_float),
                ("n_hidden", C.c_int32), ("hidden", C.c_uint32 * MAX_LAYERS), ("activation", C.c_int32),
                ("mlp_precision", C.c_int32), ("max_rows", C.c_uint64), ("max_nnz", C.c_uint64), ("rank", C.c_int32),
                ("world", C.c_int32), ("deterministic", C.c_int32), ("reserved0", C.c_int32),
                ("csc_row_block", C.c_uint64), ("ema_rate", C.c_float), ("reserved", C.c_uint32 * 3)]


class DatasetC(C.Structure):
    _fields_ = [("rows", C.c_int64), ("nnz", C.c_int64), ("label_cnt", C.c_int64), ("feature_cnt", C.c_uint64),
                ("field_cnt", C.c_uint64), ("row_ptr", C.POINTER(C.c_int64)), ("fid", C.POINTER(C.c_uint32)),
                ("field", C.POINTER(C.c_uint16)), ("val", C.POINTER(C.c_float)), ("label", C.POINTER(C.c_int32))]


_lib = None


def load_library():
    """dlopen the CUDA extension; raises loudly when it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("lightctr_b200: CUDA extension %s is missing -- run `python -m lightctr_b200.build` "
                           "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, f32p = C.c_void_p, C.c_int64, C.c_void_p
    L.lctr_last_error.restype = C.c_char_p
    L.lctr_create.argtypes = [C.POINTER(Cfg), C.POINTER(vp)]
    L.lctr_destroy.argtypes = [vp]
    L.lctr_sync.argtypes = [vp]
    L.lctr_upload_params.argtypes = [vp, f32p, f32p]
    L.lctr_download_params.argtypes = [vp, f32p, f32p]
    L.lctr_fill_params.argtypes = [vp, C.c_uint64, C.c_float]
    L.lctr_download_opt_state.argtypes = [vp, f32p, f32p]
    L.lctr_upload_opt_state.argtypes = [vp, f32p, f32p]
    L.lctr_upload_batch.argtypes = [vp, C.c_int, i64, i64, vp, vp, vp, vp, vp]
    L.lctr_train_step.argtypes = [vp, C.c_int, i64, i64, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.lctr_train_batch.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.lctr_train_batch_async.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, C.POINTER(C.c_uint64)]
    L.lctr_wait.argtypes = [vp, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.lctr_predict.argtypes = [vp, C.c_int, C.c_int, f32p]
    L.lctr_download_sumvx.argtypes = [vp, C.c_int, f32p]
    L.lctr_download_pred.argtypes = [vp, C.c_int, f32p]
    L.lctr_mlp_upload.argtypes = [vp, C.c_int, f32p, f32p]
    L.lctr_mlp_download.argtypes = [vp, C.c_int, f32p, f32p]
    L.lctr_mlp_set_mask.argtypes = [vp, C.c_int, f32p]
    L.lctr_mlp_forward.argtypes = [vp, i64, vp, vp]
    L.lctr_mlp_backward.argtypes = [vp, i64, vp, vp]
    L.lctr_mlp_apply.argtypes = [vp, C.c_uint64]
    L.lctr_mlp_download_grad.argtypes = [vp, C.c_int, f32p, f32p]
    L.lctr_set_dense_allreduce.argtypes = [vp, ALLREDUCE_FN, vp]
    L.lctr_eval.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_float)]
    L.lctr_upload_pred.argtypes = [vp, C.c_int, f32p]
    L.lctr_save_checkpoint.argtypes = [vp, C.c_char_p]
    L.lctr_load_checkpoint.argtypes = [vp, C.c_char_p]
    L.lctr_save_dataset_bin.argtypes = [C.POINTER(DatasetC), C.c_char_p]
    L.lctr_load_dataset_bin.argtypes = [C.c_char_p, C.POINTER(C.POINTER(DatasetC))]
    L.lctr_ipc_export.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.lctr_ipc_import.argtypes = [vp, vp, C.c_size_t]
    L.lctr_dense_grad_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.lctr_load_libffm.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.POINTER(C.POINTER(DatasetC))]
    L.lctr_free_dataset.argtypes = [C.POINTER(DatasetC)]
    L.lctr_launch_count.argtypes = [vp]
    L.lctr_launch_count.restype = C.c_int64
    L.lctr_stream.argtypes = [vp]
    L.lctr_stream.restype = vp
    L.lctr_profile.argtypes = [vp, C.c_int]
    L.lctr_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int, C.c_int]
    _lib = L
    return L


class LctrError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise LctrError(load_library().lctr_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data


class HostDataset:
    """CSR arrays as produced by lctr_load_libffm (FM_Algo_Abst::dataSet / label)."""

    def __init__(self, row_ptr, fid, field, val, label, feature_cnt, field_cnt):
        self.row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        self.fid = np.ascontiguousarray(fid, np.uint32)
        self.field = None if field is None else np.ascontiguousarray(field, np.uint16)
        self.val = None if val is None else np.ascontiguousarray(val, np.float32)
        self.label = np.ascontiguousarray(label, np.int32)
        self.feature_cnt, self.field_cnt = int(feature_cnt), int(field_cnt)

    rows = property(lambda self: len(self.row_ptr) - 1)
    nnz = property(lambda self: len(self.fid))


def load_libffm(path, field_cnt=0, feature_cnt=0):
    """FM_Algo_Abst::loadDataRow (fm_algo_abst.h:70-107) through the library's own parser."""
    L = load_library()
    dp = C.POINTER(DatasetC)()
    _chk(L.lctr_load_libffm(path.encode(), field_cnt, feature_cnt, C.byref(dp)))
    d = dp.contents
    n, r, lc = d.nnz, d.rows, d.label_cnt
    out = HostDataset(np.ctypeslib.as_array(d.row_ptr, (r + 1,)).copy(),
                      np.ctypeslib.as_array(d.fid, (max(n, 1),))[:n].copy(),
                      np.ctypeslib.as_array(d.field, (max(n, 1),))[:n].copy(),
                      np.ctypeslib.as_array(d.val, (max(n, 1),))[:n].copy(),
                      np.ctypeslib.as_array(d.label, (max(lc, 1),))[:lc].copy(), d.feature_cnt, d.field_cnt)
    L.lctr_free_dataset(dp)
    return out


def _dataset_from_c(d):
    n, r, lc = d.nnz, d.rows, d.label_cnt
    return HostDataset(np.ctypeslib.as_array(d.row_ptr, (r + 1,)).copy(),
                       np.ctypeslib.as_array(d.fid, (max(n, 1),))[:n].copy(),
                       np.ctypeslib.as_array(d.field, (max(n, 1),))[:n].copy(),
                       np.ctypeslib.as_array(d.val, (max(n, 1),))[:n].copy(),
                       np.ctypeslib.as_array(d.label, (max(lc, 1),))[:lc].copy(), d.feature_cnt, d.field_cnt)


def libffm_to_bin(path, bin_path, field_cnt=0, feature_cnt=0):
    """Parse a libffm text file once and write the binary CSR cache next to it."""
    L = load_library()
    dp = C.POINTER(DatasetC)()
    _chk(L.lctr_load_libffm(path.encode(), field_cnt, feature_cnt, C.byref(dp)))
    try:
        _chk(L.lctr_save_dataset_bin(dp, bin_path.encode()))
    finally:
        L.lctr_free_dataset(dp)


def load_dataset_bin(bin_path):
    L = load_library()
    dp = C.POINTER(DatasetC)()
    _chk(L.lctr_load_dataset_bin(bin_path.encode(), C.byref(dp)))
    out = _dataset_from_c(dp.contents)
    L.lctr_free_dataset(dp)
    return out


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class Context:
    """Owning wrapper around lctr_ctx."""

    def __init__(self, model, feature_cnt, factor_cnt, field_cnt=0, optimizer=OPT_ADAGRAD, lr=0.05, l2=0.001,
                 minibatch_size=0, momentum=0.8, momentum_adam2=0.999, hidden=(), activation=ACT_SIGMOID,
                 mlp_precision=MLP_FP32, device=0, rank=0, world=1, deterministic=0, csc_row_block=0, max_rows=0,
                 max_nnz=0, ema_rate=0.99):
        L = load_library()
        cfg = Cfg()
        cfg.abi_version = ABI_VERSION
        cfg.model, cfg.optimizer, cfg.device = model, optimizer, device
        cfg.feature_cnt, cfg.field_cnt, cfg.factor_cnt = feature_cnt, field_cnt, factor_cnt
        cfg.learning_rate, cfg.l2_reg, cfg.minibatch_size = lr, l2, minibatch_size
        cfg.momentum, cfg.momentum_adam2 = momentum, momentum_adam2
        cfg.ema_rate = ema_rate
        cfg.n_hidden = len(hidden)
        for i, h in enumerate(hidden):
            cfg.hidden[i] = h
        cfg.activation, cfg.mlp_precision = activation, mlp_precision
        cfg.rank, cfg.world = rank, world
        cfg.deterministic, cfg.csc_row_block = deterministic, csc_row_block
        cfg.max_rows, cfg.max_nnz = max_rows, max_nnz
        self.cfg = cfg
        self.h = C.c_void_p()
        _chk(L.lctr_create(C.byref(cfg), C.byref(self.h)))
        self.L = L
        self.F, self.k, self.Fc = feature_cnt, factor_cnt, field_cnt
        self.rowlen = factor_cnt * (field_cnt if model == MODEL_FFM else 1)
        self.slot_rows = {}

    def close(self):
        if self.h:
            self.L.lctr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _chk(self.L.lctr_sync(self.h))

    def upload_params(self, W, V):
        W = None if W is None else np.ascontiguousarray(W, np.float32)
        V = None if V is None else np.ascontiguousarray(V, np.float32)
        _chk(self.L.lctr_upload_params(self.h, _p(W), _p(V)))

    def fill_params(self, seed, scale):
        _chk(self.L.lctr_fill_params(self.h, seed, scale))

    def download_params(self):
        W = np.empty(self.F, np.float32)
        V = np.empty(self.F * self.rowlen, np.float32)
        _chk(self.L.lctr_download_params(self.h, W.ctypes.data, V.ctypes.data))
        return W, V

    def download_opt_state(self):
        n = self.F * (self.rowlen + 1)
        s1, s2 = np.zeros(n, np.float32), np.zeros(n, np.float32)
        _chk(self.L.lctr_download_opt_state(self.h, s1.ctypes.data, s2.ctypes.data))
        return s1, s2

    def upload_opt_state(self, s1, s2=None):
        s1 = np.ascontiguousarray(s1, np.float32)
        s2 = None if s2 is None else np.ascontiguousarray(s2, np.float32)
        _chk(self.L.lctr_upload_opt_state(self.h, _p(s1), _p(s2)))

    def upload_batch(self, slot, row_ptr, fid, field, val, label):
        rows, nnz = len(row_ptr) - 1, len(fid)
        keep = [np.ascontiguousarray(row_ptr, np.int64), np.ascontiguousarray(fid, np.uint32),
                None if field is None else np.ascontiguousarray(field, np.uint16),
                None if val is None else np.ascontiguousarray(val, np.float32), np.ascontiguousarray(label, np.int32)]
        _chk(self.L.lctr_upload_batch(self.h, slot, rows, nnz, *[_p(a) for a in keep]))
        self.sync()  # host arrays may be pageable temporaries
        self.slot_rows[slot] = rows

    def upload_dataset(self, slot, ds, all_ones_as_null=True):
        val = ds.val
        if val is not None and all_ones_as_null and np.all(val == 1.0):
            val = None
        self.upload_batch(slot, ds.row_ptr, ds.fid, ds.field, val, ds.label[:len(ds.row_ptr) - 1])

    def train_step(self, slot=0, row_begin=0, row_end=None, want_stats=True):
        if row_end is None:
            row_end = self.slot_rows[slot]
        if want_stats:
            loss, acc = C.c_float(), C.c_float()
            _chk(self.L.lctr_train_step(self.h, slot, row_begin, row_end, C.byref(loss), C.byref(acc)))
            return loss.value, acc.value
        _chk(self.L.lctr_train_step(self.h, slot, row_begin, row_end, None, None))
        return None

    def train_batch(self, row_ptr, fid, field, val, label):
        """End-to-end call on host buffers (numpy arrays, ideally pinned)."""
        loss, acc = C.c_float(), C.c_float()
        _chk(self.L.lctr_train_batch(self.h, len(row_ptr) - 1, len(fid), _p(row_ptr), _p(fid), _p(field), _p(val),
                                     _p(label), C.byref(loss), C.byref(acc)))
        return loss.value, acc.value

    def train_batch_async(self, row_ptr, fid, field, val, label):
        t = C.c_uint64()
        _chk(self.L.lctr_train_batch_async(self.h, len(row_ptr) - 1, len(fid), _p(row_ptr), _p(fid), _p(field), _p(val),
                                           _p(label), C.byref(t)))
        return t.value

    def wait(self, ticket):
        loss, acc = C.c_float(), C.c_float()
        _chk(self.L.lctr_wait(self.h, ticket, C.byref(loss), C.byref(acc)))
        return loss.value, acc.value

    def predict(self, slot, quirk_sumvx_slot=-1):
        out = np.empty(self.slot_rows[slot], np.float32)
        _chk(self.L.lctr_predict(self.h, slot, quirk_sumvx_slot, out.ctypes.data))
        return out

    def predict_resident(self, slot, quirk_sumvx_slot=-1):
        """forward only, predictions stay on the device (no host copy, no synchronisation)"""
        _chk(self.L.lctr_predict(self.h, slot, quirk_sumvx_slot, None))

    def download_sumvx(self, slot):
        out = np.empty(self.slot_rows[slot] * self.k, np.float32)
        _chk(self.L.lctr_download_sumvx(self.h, slot, out.ctypes.data))
        return out

    def download_pred(self, slot):
        out = np.empty(self.slot_rows[slot], np.float32)
        _chk(self.L.lctr_download_pred(self.h, slot, out.ctypes.data))
        return out

    def mlp_upload(self, layer, weight, bias):
        w = np.ascontiguousarray(weight, np.float32)
        b = np.ascontiguousarray(bias, np.float32)
        _chk(self.L.lctr_mlp_upload(self.h, layer, w.ctypes.data, b.ctypes.data))

    def mlp_download(self, layer, n_in, n_out):
        w, b = np.empty(n_in * n_out, np.float32), np.empty(n_out, np.float32)
        _chk(self.L.lctr_mlp_download(self.h, layer, w.ctypes.data, b.ctypes.data))
        return w, b

    def set_dense_allreduce(self, fn):
        """fn(dev_ptr:int, n_floats:int, cuda_stream:int) -> None: in-place SUM all-reduce enqueued on that stream."""
        def tramp(user, buf, n, stream):
            try:
                fn(buf, n, stream or 0)
                return 0
            except Exception:  # exceptions must not cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        self._allreduce_cb = ALLREDUCE_FN(tramp)  # keep alive
        _chk(self.L.lctr_set_dense_allreduce(self.h, self._allreduce_cb, None))

    def eval_metrics(self, slot):
        """(summed logloss, correct count, AUC) of the slot's pCTR / labels, computed on the device."""
        loss, correct, auc = C.c_float(), C.c_int64(), C.c_float()
        _chk(self.L.lctr_eval(self.h, slot, C.byref(loss), C.byref(correct), C.byref(auc)))
        return loss.value, correct.value, auc.value

    def upload_pred(self, slot, pctr):
        p = np.ascontiguousarray(pctr, np.float32)
        _chk(self.L.lctr_upload_pred(self.h, slot, p.ctypes.data))

    def save_checkpoint(self, path):
        _chk(self.L.lctr_save_checkpoint(self.h, path.encode()))

    def load_checkpoint(self, path):
        _chk(self.L.lctr_load_checkpoint(self.h, path.encode()))

    def mlp_download_grad(self, layer, n_in, n_out):
        w, b = np.empty(n_in * n_out, np.float32), np.empty(n_out, np.float32)
        _chk(self.L.lctr_mlp_download_grad(self.h, layer, w.ctypes.data, b.ctypes.data))
        return w, b

    def mlp_forward(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(x.shape[0], np.float32)
        _chk(self.L.lctr_mlp_forward(self.h, x.shape[0], x.ctypes.data, out.ctypes.data))
        return out

    def mlp_backward(self, dout, in0):
        dout = np.ascontiguousarray(dout, np.float32)
        dx = np.empty((len(dout), in0), np.float32)
        _chk(self.L.lctr_mlp_backward(self.h, len(dout), dout.ctypes.data, dx.ctypes.data))
        return dx

    def mlp_apply(self, minibatch):
        _chk(self.L.lctr_mlp_apply(self.h, minibatch))

    def mlp_set_mask(self, layer, mask):
        m = np.ascontiguousarray(mask, np.float32)
        _chk(self.L.lctr_mlp_set_mask(self.h, layer, m.ctypes.data))

    PROF_NAMES = ["fm_forward", "fm_backward_red", "apply", "ffm_fused", "fm_backward_csc", "mlp", "dist_mark", "dist_compact",
                  "dist_pull", "dist_push", "dist_barrier0", "dist_merge", "dist_barrier1", "csc_build", "fm_fused", "apply_compact"]

    def profile(self, enable=True):
        _chk(self.L.lctr_profile(self.h, 1 if enable else 0))

    def profile_read(self, reset=True):
        ms = (C.c_double * 16)()
        cnt = (C.c_int64 * 16)()
        _chk(self.L.lctr_profile_read(self.h, ms, cnt, 16, 1 if reset else 0))
        return {self.PROF_NAMES[i]: (ms[i], cnt[i]) for i in range(16) if cnt[i] > 0}

    def ipc_export(self):
        n = C.c_size_t()
        _chk(self.L.lctr_ipc_export(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _chk(self.L.lctr_ipc_export(self.h, buf, n.value, C.byref(n)))
        return buf.raw

    def ipc_import(self, all_blobs, bytes_per_rank):
        buf = C.create_string_buffer(all_blobs, len(all_blobs))
        _chk(self.L.lctr_ipc_import(self.h, buf, bytes_per_rank))

    def device_bytes(self):
        a, b = C.c_uint64(), C.c_uint64()
        _chk(self.L.lctr_device_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self):
        return self.L.lctr_launch_count(self.h)

    def stream(self):
        return self.L.lctr_stream(self.h)
