"""Writes synthetic Criteo-shaped batches (lightctr_b200/data.py) as flat binaries for scripts/lab/fm_lab.
    python scripts/lab/dump_batch.py OUT_PREFIX F B [B ...]
File: int64 header {rows, nnz, F, n_fields}, row_ptr int64[rows+1], fid uint32[nnz], field uint16[nnz] (padded to 4 B), label int32[rows]."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from lightctr_b200.data import BASE_SEED, CriteoSynth  # noqa: E402


def main():
    prefix, F = sys.argv[1], int(sys.argv[2])
    for B in map(int, sys.argv[3:]):
        gen = CriteoSynth(F, seed=BASE_SEED)
        rp, fid, fld, lab = gen.batch(B)
        with open("%s_F%d_B%d.bin" % (prefix, F, B), "wb") as f:
            np.array([B, len(fid), F, 39], np.int64).tofile(f)
            rp.astype(np.int64).tofile(f)
            fid.astype(np.uint32).tofile(f)
            fl = fld.astype(np.uint16)
            if len(fl) % 2:
                fl = np.concatenate([fl, np.zeros(1, np.uint16)])
            fl.tofile(f)
            lab.astype(np.int32).tofile(f)
        print("wrote B=%d nnz=%d uniq=%d" % (B, len(fid), len(np.unique(fid))))


if __name__ == "__main__":
    main()
