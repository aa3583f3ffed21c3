"""Generate tests/golden/ring_allreduce.json: what the UNMODIFIED reference ring all-reduce (Worker_RingReduce<float>::
syncGradient, distribut/ring_collect.h:50-72,112-200) leaves in every worker's buffer, from real runs of a ring master and R
worker processes over ZeroMQ on 127.0.0.1 (oracle/ref_ring_driver.cpp; `make -C oracle refdist`).  Element i of rank r starts as
float(sin(0.37 i + 1.3 r) * (1 + r)).  tests/test_oracle_ring_cpu.py holds oracle.orc_ring_allreduce against it, bit for bit."""
import json
import math
import os
import re
import signal
import socket
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref")
CASES = [(2, 10, 1), (3, 10, 1), (3, 37, 0), (4, 1000, 1), (4, 5, 1)]  # (workers, floats, do_average); 37, 5: ragged segments (the reference hangs on an empty segment: floats >= workers)


def initial(R, P):
    return [np.array([np.float32(math.sin(0.37 * i + 1.3 * r) * (1.0 + r)) for i in range(P)], np.float32) for r in range(R)]


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run(R, P, avg, timeout=120):
    """-> list over ranks of the result as a hex string (8 digits per float)"""
    env = dict(os.environ, LightCTR_PS_NUM="0", LightCTR_WORKER_NUM=str(R), LightCTR_MASTER_ADDR="127.0.0.1:%d" % free_port())
    procs = [subprocess.Popen([os.path.join(REF, "role_master_ring")], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)]
    outs = []
    try:
        time.sleep(1.0)
        workers = []
        for r in range(R):
            workers.append(subprocess.Popen([os.path.join(REF, "role_worker_ring"), str(P), str(avg), str(5 + r)], env=env,
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            time.sleep(0.3)  # ranks are handed out in arrival order
        for w in workers:
            outs.append(w.communicate(timeout=timeout)[0])
    finally:
        for p in procs + (workers if "workers" in dir() else []):
            if p.poll() is None:
                p.send_signal(signal.SIGKILL)
    res = {}
    for o in outs:
        m = re.search(r"\[ring result\] rank (\d+)((?: [0-9a-f]{8})+)", o)
        assert m, o[-1500:]
        res[int(m.group(1))] = m.group(2).replace(" ", "")
    assert sorted(res) == list(range(R)), sorted(res)
    return [res[r] for r in range(R)]


if __name__ == "__main__":
    rec = {"what": "Worker_RingReduce<float>::syncGradient of the unmodified reference, one run per case", "generator": "tests/golden/make_ring_golden.py",
           "cases": []}
    for (R, P, avg) in CASES:
        got = run(R, P, avg)
        assert len(set(got)) == 1, "ranks disagree"
        rec["cases"].append({"workers": R, "floats": P, "do_average": avg, "result_hex": got[0]})
        print(R, P, avg, got[0][:64])
    json.dump(rec, open(os.path.join(HERE, "ring_allreduce.json"), "w"))
