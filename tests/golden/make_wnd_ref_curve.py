"""Generate tests/golden/wnd_ref_curve.json: the loss curve of the UNMODIFIED reference cluster -- Master, ParamServer and
one Distributed_Algo_Abst worker talking ZeroMQ on 127.0.0.1 -- on the reference's own data/train_sparse.csv.

    make -C oracle refdist            # oracle/_ref/role_{master,ps,worker}, oracle/_ref/umap_order (needs /root/reference)
    python tests/golden/make_wnd_ref_curve.py [epochs]

What is recorded besides the curve, and why:
  * `first_touch`: the order in which the parameter server first sees the per-field tensor keys (oracle/umap_order.cpp replays
    the worker's std::unordered_map): the server initialises a tensor from ITS rand() stream at that moment
    (distribut/paramserver.h:40-46), so the order decides which tensor gets which Gauss draws;
  * `ps_rand_skip` / `worker_rand_skip`: how many rand() calls each process spent on picking a listen port
    (common/network.h:366-383: one per bind attempt) before the numbers that matter (tensors; dense-layer weights and dropout
    masks).  Derived from the printed port and glibc's rand() sequence for the seed.
The CPU test tests/test_oracle_wnd_pin_cpu.py replays the run with oracle.api.WNDOracle(schedule="reference")."""
import ctypes
import json
import os
import re
import signal
import socket
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")


def rand_sequence(seed, n):
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(seed)
    return [libc.rand() for _ in range(n)]


def rand_skip(seed, port):
    """number of rand() calls until 1024 + rand() % 64512 == port (common/network.h:372)"""
    for i, r in enumerate(rand_sequence(seed, 64)):
        if 1024 + r % (65536 - 1024) == port:
            return i + 1
    raise RuntimeError("port %d not in the rand() sequence of seed %d" % (port, seed))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


UPDATERS = ["sgd", "adagrad", "dcasgd", "dcasgda"]  # paramserver.h:22-27; the reference's main.cpp runs the first


def run_cluster(prefix, tmp, epochs, seed_ps, seed_worker, updater, timeout=600):
    env = dict(os.environ, LightCTR_PS_NUM="1", LightCTR_WORKER_NUM="1", LightCTR_MASTER_ADDR="127.0.0.1:%d" % free_port())
    logs = {r: open(os.path.join(tmp, r + ".log"), "w") for r in ("master", "ps")}
    procs = []
    try:
        procs.append(subprocess.Popen([os.path.join(REF, "role_master"), "1"], env=env, stdout=logs["master"], stderr=subprocess.STDOUT))
        time.sleep(1.0)
        procs.append(subprocess.Popen([os.path.join(REF, "role_ps"), str(seed_ps), str(updater)], env=env, stdout=logs["ps"], stderr=subprocess.STDOUT))
        time.sleep(1.0)
        out = subprocess.run([os.path.join(REF, "role_worker"), str(seed_worker), prefix, str(epochs)], env=env, capture_output=True,
                             text=True, timeout=timeout).stdout
    finally:
        for p in procs:
            p.send_signal(signal.SIGKILL)
        for f in logs.values():
            f.close()
    got = re.findall(r"\[Worker Train\] epoch = (\d+) loss = ([0-9.eE+-]+) accuracy = ([0-9.eE+-]+)", out)
    assert len(got) == epochs, out[-2000:]
    port_w = int(re.search(r"\[Network\] Listening tcp://[0-9.]+:(\d+)", out).group(1))
    port_ps = int(re.search(r"\[Network\] Listening tcp://[0-9.]+:(\d+)", open(os.path.join(tmp, "ps.log")).read()).group(1))
    return ([float(g[1]) for g in got], [float(g[2]) for g in got], rand_skip(seed_ps, port_ps), rand_skip(seed_worker, port_w))


def main():
    epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    seed_ps, seed_worker = 11, 7
    from golden_util import load_csr, write_libffm
    ds = load_csr("train_sparse_csr.npz", field_cnt=68)
    tmp = tempfile.mkdtemp(prefix="wnd_ref_")
    prefix = os.path.join(tmp, "ad_data")
    write_libffm(ds, prefix + "_1.csv")  # the first worker's rank is 1 (distributed_algo_abst.h:97-100)
    order = [int(x) for x in subprocess.check_output([os.path.join(REF, "umap_order"), prefix + "_1.csv"], text=True).split()]
    rec = {"what": "reference Master + ParamServer + Distributed_Algo_Abst worker over ZeroMQ, data/train_sparse.csv (1000 rows), "
                   "one run per server updater (paramserver.h:22-27)",
           "generator": "tests/golden/make_wnd_ref_curve.py", "epochs": epochs, "minibatch": 50, "learning_rate": 0.05,
           "sparse_rate": 0.8, "factor_dim": 4, "hidden": 50, "seed_ps": seed_ps, "seed_worker": seed_worker, "curves": {},
           "first_touch": order}
    for u, name in enumerate(UPDATERS):
        loss, acc, skip_ps, skip_w = run_cluster(prefix, tmp, epochs, seed_ps, seed_worker, u)
        rec["curves"][name] = {"loss": loss, "accuracy": acc}
        rec["ps_rand_skip"], rec["worker_rand_skip"] = skip_ps, skip_w
        print(name, loss, acc, skip_ps, skip_w)
    with open(os.path.join(HERE, "wnd_ref_curve.json"), "w") as f:
        json.dump(rec, f)
    print(len(order), "tensors")


if __name__ == "__main__":
    main()
