"""Generate tests/golden/dht_nodes.json from the reference's own ConsistentHash (oracle/_ref/dht_nodes, built by
`make -C oracle refdist` from distribut/consistent_hash.h where it lies): server index of a fixed key set for several cluster
sizes.  tests/test_oracle_dht_cpu.py holds oracle.orc_dht_node against it."""
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "dht_nodes")


def keys():
    rng = np.random.default_rng(20)
    return list(range(4096)) + [int(x) for x in rng.integers(0, 2 ** 63, 512, dtype=np.uint64)] + [2 ** 64 - 1, 2 ** 32, 2 ** 32 - 1]


def run(ps_cnt, ks):
    out = subprocess.run([EXE, str(ps_cnt)], input="\n".join(str(k) for k in ks), capture_output=True, text=True, check=True).stdout
    return [int(x) for x in out.split()]


if __name__ == "__main__":
    ks = keys()
    nodes = {n: run(n, ks) for n in (1, 2, 3, 4, 5, 8, 16)}
    rec = {"what": "ConsistentHash::getNode of the unmodified reference (consistent_hash.h:29-40) for the keys of keys() in this script",
           "generator": "tests/golden/make_dht_golden.py", "n_keys": len(ks),
           "nodes_hex": {str(n): "".join("%x" % v for v in vs) for n, vs in nodes.items()}}  # one hex digit per key
    json.dump(rec, open(os.path.join(HERE, "dht_nodes.json"), "w"))
    print({n: np.bincount(v).tolist() for n, v in nodes.items()})
