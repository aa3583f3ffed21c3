"""Synthetic Criteo-shaped sparse batches (SURVEY.md section 8d "Synthetic inputs").

39 fields (13 numeric-bucket + 26 categorical); the id space [0, F) is split into 39 contiguous per-field
vocabularies whose sizes follow a geometric ladder (smallest 16, largest ~F/4); per row field j emits
c_j = 1 + Poisson(1.05) ids (~80 ids/row in total), drawn Zipf(alpha=1.1) inside the field's vocabulary and
de-duplicated within the row; all values 1.0 (as in the reference's train_sparse.csv); labels are
Bernoulli(sigmoid(planted linear logit)) calibrated to ~9 % positives.
"""
import numpy as np

N_FIELDS = 39
BASE_SEED = 20240917


def field_vocab(F, n_fields=N_FIELDS):
    """Per-field vocabulary sizes (sum == F) and start offsets."""
    ladder = 16.0 * (max(F / 4.0, 32.0) / 16.0) ** (np.arange(n_fields) / (n_fields - 1))
    sizes = np.maximum(16, np.floor(ladder * (F / ladder.sum()))).astype(np.int64)
    sizes[-1] += F - sizes.sum()
    assert sizes.min() >= 1 and sizes.sum() == F
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    return sizes, starts


class CriteoSynth:
    def __init__(self, F, seed=BASE_SEED, n_fields=N_FIELDS, alpha=1.1, lam=1.05, pos_rate=0.09):
        self.F, self.n_fields = F, n_fields
        self.sizes, self.starts = field_vocab(F, n_fields)
        self.rng = np.random.default_rng(seed)
        self.lam = lam
        # Zipf CDF per field (rank r has weight r^-alpha), capped table size for the huge fields: ranks beyond the
        # table are drawn uniformly from the tail mass
        self.cdfs = []
        for n in self.sizes:
            m = int(min(n, 1 << 20))
            w = np.arange(1, m + 1, dtype=np.float64) ** (-alpha)
            tail = 0.0
            if n > m:  # integral approximation of the remaining mass
                tail = ((n + 0.5) ** (1 - alpha) - (m + 0.5) ** (1 - alpha)) / (1 - alpha)
            c = np.cumsum(w)
            self.cdfs.append((c / (c[-1] + tail), m, int(n)))
        wr = np.random.default_rng(seed ^ 0x5bd1e995)
        self.w_true = (wr.standard_normal(F) * 0.35).astype(np.float32)
        self.bias = None
        self.pos_rate = pos_rate

    def _draw_field(self, j, count):
        cdf, m, n = self.cdfs[j]
        u = self.rng.random(count)
        r = np.searchsorted(cdf, u)
        tail = r >= m
        if tail.any():
            r[tail] = self.rng.integers(m, n, tail.sum())
        return self.starts[j] + r

    def batch(self, rows):
        """-> (row_ptr int64[rows+1], fid uint32[nnz], field uint16[nnz], label int32[rows]); entries of a row are
        ordered by field, ids unique within the row."""
        rng = self.rng
        rid_parts, fid_parts, fld_parts = [], [], []
        for j in range(self.n_fields):
            c = 1 + rng.poisson(self.lam, rows)
            rid = np.repeat(np.arange(rows, dtype=np.int64), c)
            ids = self._draw_field(j, len(rid))
            key = np.unique(rid * self.F + ids)  # de-duplicate within (row, field)
            rid_parts.append(key // self.F)
            fid_parts.append(key % self.F)
            fld_parts.append(np.full(len(key), j, np.uint16))
        rid = np.concatenate(rid_parts)
        fid = np.concatenate(fid_parts)
        fld = np.concatenate(fld_parts)
        order = np.lexsort((fid, fld, rid))
        rid, fid, fld = rid[order], fid[order], fld[order]
        row_ptr = np.zeros(rows + 1, np.int64)
        np.add.at(row_ptr, rid + 1, 1)
        row_ptr = np.cumsum(row_ptr)
        logit = np.bincount(rid, weights=self.w_true[fid], minlength=rows)
        if self.bias is None:  # calibrate once so that mean(sigmoid) ~= pos_rate
            lo, hi = -20.0, 20.0
            for _ in range(60):
                mid = 0.5 * (lo + hi)
                if (1.0 / (1.0 + np.exp(-(logit + mid)))).mean() > self.pos_rate:
                    hi = mid
                else:
                    lo = mid
            self.bias = 0.5 * (lo + hi)
        p = 1.0 / (1.0 + np.exp(-(logit + self.bias)))
        label = (rng.random(rows) < p).astype(np.int32)
        return row_ptr, fid.astype(np.uint32), fld, label


def write_libffm(path, row_ptr, fid, field, label, val=None):
    """label<TAB>field:fid:val ... (the reference's on-disk format, fm_algo_abst.h:88-93)."""
    with open(path, "w") as f:
        for r in range(len(row_ptr) - 1):
            b, e = row_ptr[r], row_ptr[r + 1]
            if val is None:
                toks = " ".join("%d:%d:1" % (field[i], fid[i]) for i in range(b, e))
            else:
                toks = " ".join("%d:%d:%r" % (field[i], fid[i], float(val[i])) for i in range(b, e))
            f.write("%d\t%s\n" % (label[r], toks))
