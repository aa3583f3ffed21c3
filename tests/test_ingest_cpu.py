"""Host-side ingest: the library's libffm parser against the oracle's restatement of
FM_Algo_Abst::loadDataRow (fm_algo_abst.h:70-107) on the reference's own train_sparse.csv (as committed CSR fixture,
regenerated to text), and the binary CSR cache round trip.  No GPU needed: these entry points are host code."""
import numpy as np

from golden_util import load_csr, write_libffm


def test_libffm_parser_matches_oracle_and_bin_cache_roundtrips(tmp_path, oracle_api):
    from lightctr_b200 import capi
    tr = load_csr("train_sparse_csr.npz", field_cnt=68)
    txt, binp = str(tmp_path / "train.csv"), str(tmp_path / "train.csr")
    write_libffm(tr, txt)
    a = capi.load_libffm(txt, 68, 0)
    o = oracle_api.load(txt, 68)
    assert a.rows == o.rows and a.feature_cnt == o.feature_cnt and a.field_cnt == o.field_cnt
    assert np.array_equal(a.row_ptr, o.row_ptr) and np.array_equal(a.fid, o.fid)      # indexing bit-exact
    assert np.array_equal(a.field.astype(np.uint32), o.field.astype(np.uint32))
    assert np.array_equal(a.val.view(np.uint32), o.val.view(np.uint32))
    assert np.array_equal(a.label[:a.rows], np.asarray(o.label[:o.rows], dtype=a.label.dtype))
    capi.libffm_to_bin(txt, binp, 68, 0)
    b = capi.load_dataset_bin(binp)
    for name in ("row_ptr", "fid", "field", "label"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert np.array_equal(a.val.view(np.uint32), b.val.view(np.uint32))
    assert (a.rows, a.feature_cnt, a.field_cnt) == (b.rows, b.feature_cnt, b.field_cnt)


def test_bin_cache_rejects_foreign_files(tmp_path):
    import pytest
    from lightctr_b200 import capi
    p = tmp_path / "junk.bin"
    p.write_bytes(b"not a cache" * 10)
    with pytest.raises(capi.LctrError):
        capi.load_dataset_bin(str(p))


def test_bin_cache_header_is_not_trusted(tmp_path):
    """a cache whose counts disagree with each other or with the file length is an error, not an allocation of whatever the
    header says (rows / nnz / label count are checked before any buffer is sized from them)"""
    import struct
    import pytest
    from golden_util import load_csr, write_libffm
    from lightctr_b200 import capi
    ds = load_csr("train_sparse_csr.npz", field_cnt=68)
    txt, binp = str(tmp_path / "t.csv"), str(tmp_path / "t.bin")
    write_libffm(ds, txt)
    capi.libffm_to_bin(txt, binp, 68, 0)
    good = open(binp, "rb").read()
    rows, nnz, labels = struct.unpack_from("<3Q", good, 8)
    for bad_hdr in ((rows, nnz + 1, labels), (rows + 5, nnz, labels), (rows, nnz, rows - 1), (1 << 62, nnz, labels)):
        p = tmp_path / "bad.bin"
        p.write_bytes(good[:8] + struct.pack("<3Q", *bad_hdr) + good[32:])
        with pytest.raises(capi.LctrError):
            capi.load_dataset_bin(str(p))
    p = tmp_path / "short.bin"
    p.write_bytes(good[:-100])
    with pytest.raises(capi.LctrError):
        capi.load_dataset_bin(str(p))
