"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/lightctr_b200.h declares; without a GPU the product fails loudly (no CPU fallback); the host-side
libffm parser is bit-exact against the oracle's restatement of the reference parser."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT

from lightctr_b200 import build as lbuild
from lightctr_b200 import capi


@pytest.fixture(scope="module")
def lib():
    lbuild.build()
    return capi.load_library()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "lightctr_b200.h")).read()
    declared = set(re.findall(r"\b(lctr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.lctr_abi_version() == capi.ABI_VERSION


def test_cfg_struct_layout_matches_header():
    # sizeof(lctr_cfg) computed by the C compiler must equal the ctypes mirror
    import subprocess, tempfile
    src = '#include "lightctr_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu\\n", sizeof(lctr_cfg), sizeof(lctr_dataset));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "s.c")
        open(p, "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), p, "-o", os.path.join(d, "s")])
        a, b = subprocess.check_output([os.path.join(d, "s")]).split()
    assert int(a) == C.sizeof(capi.Cfg) and int(b) == C.sizeof(capi.DatasetC)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.LctrError, match="no CUDA device|CUDA"):
        capi.Context(capi.MODEL_FM, 100, 8)


def test_bad_args_report_errors(lib):
    cfg = capi.Cfg()
    h = C.c_void_p()
    assert lib.lctr_create(C.byref(cfg), C.byref(h)) != 0
    assert b"abi_version" in lib.lctr_last_error()
    assert lib.lctr_sync(None) != 0


LINES = [
    "1\t0:3:1 1:7:0.5 2:11:2\n",
    "0 3:1:1 4:2:1\n",
    "\n",                                  # empty line
    "1\t\n",                               # label but no features: label kept, row skipped (fm_algo_abst.h:90,101)
    "0\t5:9:1.25 6:10 7:12:3\n",           # token with missing value -> keeps previous val, stale %n
    "-1\t0:0:1e-3 1:1:-2.5 2:2:+4\n",      # exponent / signs -> scanf fallback path
    "1\t  8:20:1   9:21:1\n",              # extra spaces
    "0\t1:5:1 junk 2:6:1\n",               # parse stops at junk
    "1\t10:100000:1\r\n",                  # CRLF
    "0\t11:4294967295:7",                  # max u32 fid, no trailing newline
]


def test_loader_bit_exact_vs_oracle(lib, oracle_api, tmp_path):
    p = str(tmp_path / "edge.txt")
    open(p, "w").write("".join(LINES))
    for fc in (0, 3):
        a = capi.load_libffm(p, field_cnt=fc)
        b = oracle_api.load(p, field_cnt=fc)
        assert (a.rows, a.nnz, a.feature_cnt, a.field_cnt) == (b.rows, b.nnz, b.feature_cnt, b.field_cnt)
        assert np.array_equal(a.row_ptr, b.row_ptr)
        assert np.array_equal(a.fid, b.fid)
        assert np.array_equal(a.field.astype(np.uint32), b.field)
        assert np.array_equal(a.val.view(np.uint32), b.val.view(np.uint32))
        assert np.array_equal(a.label, b.label)


def test_loader_golden_roundtrip(lib, oracle_api, tmp_path):
    from golden_util import load_csr, write_libffm
    ds = load_csr("train_sparse_csr.npz", field_cnt=68)
    p = str(tmp_path / "train.txt")
    write_libffm(ds, p)
    a = capi.load_libffm(p, field_cnt=68)
    assert (a.rows, a.nnz, a.feature_cnt, a.field_cnt) == (1000, 281975, 233789, 68)
    assert np.array_equal(a.row_ptr, ds.row_ptr) and np.array_equal(a.fid, ds.fid)
    assert np.array_equal(a.field.astype(np.uint32), ds.field) and np.array_equal(a.label[:1000], ds.label)
    assert np.all(a.val == 1.0)


def test_missing_file_error(lib):
    with pytest.raises(capi.LctrError, match="open file error"):
        capi.load_libffm("/nonexistent/file.csv")
