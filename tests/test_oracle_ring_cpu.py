"""PIN of the oracle's ring all-reduce (oracle/lightctr_oracle.c:orc_ring_allreduce, the model the gloo / CUDA multi-rank tests
use for `Worker_RingReduce` + `BufferFusion`) against the UNMODIFIED reference: tests/golden/ring_allreduce.json
holds what real runs of the reference's ring master and R worker processes over ZeroMQ leave in every worker's buffer
(oracle/ref_ring_driver.cpp, tests/golden/make_ring_golden.py).  fp32 sums whose ORDER is fixed by the ring schedule
(ring_collect.h:112-200: at step i rank r adds what it receives into segment (r - i - 1) mod R, received + local): bit-exact."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from golden_util import GOLDEN

sys.path.insert(0, GOLDEN)
import make_ring_golden as gen  # noqa: E402


def _hex(a):
    return "".join("%08x" % v for v in a.view(np.uint32))


def test_ring_allreduce_is_bit_exact(oracle_api):
    g = json.load(open(os.path.join(GOLDEN, "ring_allreduce.json")))
    L = oracle_api.lib()
    assert len(g["cases"]) >= 5
    for case in g["cases"]:
        R, P, avg = case["workers"], case["floats"], case["do_average"]
        bufs = gen.initial(R, P)
        arr = (C.POINTER(C.c_float) * R)(*[b.ctypes.data_as(C.POINTER(C.c_float)) for b in bufs])
        L.orc_ring_allreduce(arr, R, P, avg)
        for r in range(R):  # every rank ends with the same buffer, and it is the reference's
            assert _hex(bufs[r]) == case["result_hex"], (R, P, avg, r)
    # the order matters: the same 4 x 5 prefix summed under the 4 x 1000 segmentation differs in the last place
    c1000 = next(c for c in g["cases"] if c["floats"] == 1000)["result_hex"]
    c5 = next(c for c in g["cases"] if c["floats"] == 5)["result_hex"]
    assert c1000[:40] != c5


@pytest.mark.skipif(not os.path.exists(os.path.join(gen.REF, "role_worker_ring")), reason="reference ring roles not built (make -C oracle refdist)")
def test_golden_is_what_the_reference_ring_computes():
    g = json.load(open(os.path.join(GOLDEN, "ring_allreduce.json")))
    case = next(c for c in g["cases"] if c["workers"] == 3 and c["floats"] == 37)
    try:
        got = gen.run(3, 37, case["do_average"], timeout=60)
    except (OSError, subprocess.SubprocessError, AssertionError) as e:
        pytest.skip("could not run the reference ring here: %r" % (e,))
    assert got == [case["result_hex"]] * 3
