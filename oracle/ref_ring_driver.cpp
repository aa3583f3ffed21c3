// oracle/ref_ring_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// The UNMODIFIED reference ring all-reduce (distribut/ring_collect.h: Worker_RingReduce<float>::syncGradient over a
// BufferFusion<float>, common/buffer_fusion.h) and its ring master (main.cpp:123-127), compiled where they lie under
// /root/reference by oracle/Makefile (target refdist) into role_ring_master / role_ring_worker, linked against the pyzmq wheel's
// libzmq like oracle/ref_dist_driver.cpp.  tests/golden/make_ring_golden.py runs a master and R workers on 127.0.0.1 and records
// every worker's buffer after one syncGradient: the pin of oracle/lightctr_oracle.c:orc_ring_allreduce (segment layout, the
// order in which partial sums meet, the final 1/R scaling).
//   worker: role_ring_worker <P> <do_average> ; buffer element i of rank r (0-based) is sin(0.37 i + 1.3 r) * (1 + r), as float
#include <algorithm>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <unistd.h>
#include <fcntl.h>

using namespace std;  // the reference headers below rely on it (main.cpp gets it from the headers it includes before them)
#include "LightCTR/distribut/master.h"
#include "LightCTR/distribut/dist_machine_abst.h"
#include "LightCTR/distribut/ring_collect.h"

int main(int argc, char** argv) {
    setvbuf(stdout, NULL, _IOLBF, 0);
    srand(argc > 3 ? (unsigned)atoi(argv[3]) : 5u);
#ifdef MASTER_RING
    { Master master(Run_Mode::Ring_Mode); }
#elif defined WORKER_RING
    {
        if (argc < 3) { fprintf(stderr, "usage: %s P do_average [seed]\n", argv[0]); return 2; }
        const size_t P = (size_t)atol(argv[1]);
        const bool avg = atoi(argv[2]) != 0;
        Worker_RingReduce<float>* ring = new Worker_RingReduce<float>(__global_cluster_worker_cnt);
        const size_t r = ring->Rank();
        std::vector<float> data(P);
        for (size_t i = 0; i < P; i++) data[i] = (float)(std::sin(0.37 * (double)i + 1.3 * (double)r) * (1.0 + (double)r));
        std::shared_ptr<BufferFusion<float> > buf = std::make_shared<BufferFusion<float> >(false, false);
        buf->registMemChunk(data.data(), P);
        ring->syncGradient(buf, 1, avg);
        printf("[ring result] rank %zu", r);
        for (size_t i = 0; i < P; i++) { unsigned u; memcpy(&u, &data[i], 4); printf(" %08x", u); }
        printf("\n");
        fflush(stdout);
        // stay reachable for a moment: a neighbour's last segment may still be waiting for this process's acknowledgement
        // (send_sync, ring_collect.h:205-215), and there is no shutdown handshake in this driver
        std::this_thread::sleep_for(std::chrono::milliseconds(2500));
        _exit(0);
    }
#else
#error "compile with -D MASTER_RING or -D WORKER_RING"
#endif
    return 0;
}
