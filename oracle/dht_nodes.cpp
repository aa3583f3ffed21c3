// oracle/dht_nodes.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's key -> parameter-server map, ConsistentHash::getNode (distribut/consistent_hash.h:29-40: murmur hash of the
// key, lower_bound on a ring of 5 virtual nodes per server whose positions are murmur hashes of "<server>-<replica>"),
// compiled UNMODIFIED where it lies under /root/reference (oracle/Makefile, target refdist; header-only, no ZeroMQ).
// tests/golden/make_dht_golden.py records its answers; oracle/lightctr_oracle.c:orc_dht_node restates it.
//   usage: dht_nodes <ps_cnt> < keys (one unsigned 64-bit key per line)  ->  one server index per line
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>

static uint32_t g_ps_cnt = 1;
#define __global_cluster_ps_cnt g_ps_cnt   // distribut/master.h:23 reads it from LightCTR_PS_NUM; here from argv
#include "LightCTR/distribut/consistent_hash.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    g_ps_cnt = (uint32_t)atoi(argv[1]);
    unsigned long long key;
    while (scanf("%llu", &key) == 1) printf("%u\n", ConsistentHash::Instance().getNode((size_t)key));
    return 0;
}
