"""Helpers to load the committed golden fixtures (tests/golden/, produced by make_golden.py)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


def load_csr(name, field_cnt=0):
    """-> oracle.api.Dataset built from the committed CSR of the reference's parsed dataSet."""
    from oracle import api
    z = np.load(os.path.join(GOLDEN, name))
    nnz = len(z["fid"])
    val = z["val"] if len(z["val"]) else np.ones(nnz, np.float32)
    fc = int(z["field"].max()) + 1 if field_cnt else 0
    return api.Dataset(z["row_ptr"], z["fid"], z["field"].astype(np.uint32), val, z["label"], int(z["feature_cnt"]),
                       max(fc, field_cnt) if field_cnt else 0)


def write_libffm(ds, path):
    """Write a Dataset back out in the reference's text format (fm_algo_abst.h:88-93): label<TAB>field:fid:val ..."""
    with open(path, "w") as f:
        for r in range(ds.rows):
            b, e = ds.row_ptr[r], ds.row_ptr[r + 1]
            toks = ["%d:%d:%s" % (ds.field[i], ds.fid[i], repr(float(ds.val[i])).rstrip("0").rstrip(".")
                                  if ds.val[i] == int(ds.val[i]) else repr(float(ds.val[i]))) for i in range(b, e)]
            f.write("%d\t%s\n" % (ds.label[r], " ".join(toks)))
