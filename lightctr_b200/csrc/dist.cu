// lightctr_b200/csrc/dist.cu -- multi-GPU sparse exchange over NVLink peer memory (one process per GPU).
//
// Replaces the reference's Parameter-Server round trips: Pull::sync of the batch's unique keys
// (distribut/pull.h:43-68; unique-key build distributed_algo_abst.h:181-195) and Push::sync of the per-key
// gradients (distribut/push.h:36-51) with the owner applying the update (distribut/paramserver.h:181-310).
// Differences we state rather than reproduce: synchronous instead of SSP-async, fp32 instead of
// fp16 on the wire, no |g| thresholding of pushes, owner = fid mod R instead of the murmur DHT ring.
//
// Tables are owner-sharded: row f lives on rank f % R at shard-local index f / R.  Every rank maps every peer's
// W / V / update_g shards, touched map and ONE arena (flags, key inboxes, parameter cache, gradient inboxes) through
// CUDA IPC.  Per-rank memory is the shard plus O(keys of a batch): the compute kernels work on a BATCH-COMPACT cache
// (row = slot of the batch's key set, fm_fused.cu's slot map) instead of full-size copies of the tables.
//
//   upload (depends only on the batch; on the upload stream, overlaps the previous step)
//     slot map of my batch (mark / compact / assign) and, per owner o, the list of the shard-local rows o owns, written
//     straight into o's key inbox with posted stores + a generation flag (send_keys_kernel).  From then on a row of the
//     exchange is addressed by p = o * cap_pair + its position in that list, by requester and owner alike: the entries of
//     the batch are re-indexed to p (remap_entries_kernel), so the cache rows of one owner and the gradient rows for one
//     owner are CONTIGUOUS and both transfers are long coalesced streams instead of scattered 64 B packets (the first
//     version of this protocol wrote pulled rows to scattered cache slots: 48 us for the 3 MB of an FM C2 step).
//   step
//     1 serve   OWNER-driven pull: I read the key lists my peers posted and WRITE the rows they asked for into their
//               caches (posted peer stores instead of read round trips), then raise "rows delivered" on each peer
//                                                                                                (serve_pull_kernel)
//     2 compute the single-GPU kernels on the cache; the first one waits (in-kernel) for every owner's flag
//     3 push    owner by owner, my gradient rows [gV | gW] (hot replicas folded) stream into the owner's gradient inbox
//               at the positions of the list it received -- no slot reservation, no atomics on the wire; then "pushes
//               landed" flags                                                                    (push_rows_kernel)
//     4 owner   inbox rows are added into update_g with local REDs (waits for the flags)          (merge_kernel)
//               sparse updater on the shard                                                      (opt.cu)
// Launches per step: 6; the barriers of the r01 protocol are flags written at the tail of one kernel and polled at the
// head of the next -- no barrier launches, no host involvement.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "opt.cuh"

namespace lctr {

constexpr int kMaxWorld = 8;
constexpr int kHotRepD = kHotRep, kHotMaxD = kHotMax;
constexpr int kNumHandles = 6;  // W, V, gW, gV, touched, arena

struct Peer {
    float *W, *V, *gW, *gV;
    uint8_t* touched;
    unsigned char* arena;
};
struct PeerTable { Peer p[kMaxWorld]; };

// byte offsets inside the arena (identical on every rank)
struct ArenaLayout {
    size_t flags;        // u64 [3 + kNumSlots][kMaxWorld]: row 0 rows delivered, 1 pushes landed, 3 + s keys of slot s
    size_t key_inbox;    // [kNumSlots][2][world] regions (2: parity of the slot's upload generation -- a list may still be read
                         // by a slower owner's merge when its sender already uploads the slot's next batch):
                         // 64 B header {u32 count} + cap_pair x uint2 {row, requester slot}
    size_t key_region;   // bytes per region
    size_t cacheW;       // world * cap_pair floats (row p)
    size_t cacheV;       // world * cap_pair x rowlen floats
    size_t grad_inbox;   // [world] regions of cap_pair x recw floats
    size_t grad_region;  // bytes per region
    size_t total;
};
enum { FLAG_PULLED = 0, FLAG_PUSHED = 1, FLAG_KEYS = 3 };

struct DistState {
    int rank = 0, world = 1, shift = 0;
    unsigned char* arena = nullptr;
    ArenaLayout A;
    PeerTable peers;
    size_t cap_keys = 0, cap_pair = 0;
    int recw = 0;                     // floats per gradient-inbox record: rowlen + 4 ([gV | gW | pad])
    size_t rows_x = 0;                // world * cap_pair: rows of the exchange index space p
    unsigned int* send_cnt = nullptr; // [kMaxWorld] records appended per owner by the running send_keys
    unsigned int* seg_cnt = nullptr;  // [kNumSlots][kMaxWorld]: keys per owner of each slot's batch
    uint32_t* opos = nullptr;         // [kNumSlots][cap_keys]: exchange row p of each of my slots
    uint32_t* hot_p = nullptr;        // [kNumSlots][rows_x]: replica block of hot exchange rows (fused kernels), else ~0
    unsigned int* done_ctr = nullptr; // [4] last-block counters
    int* overflow = nullptr;          // device flag: a key list outgrew cap_pair
    float *cgV = nullptr, *cgW = nullptr;  // [rows_x][rowlen], [rows_x]: compact gradient rows of the non-fused kernels
    // owner side of the fused FM / NFM step: the UNION of the key lists the requesters sent for a slot's batch, built once per
    // upload, and for every union row its position in each requester's list (~0: not asked for) -- the updater walks it and
    // sums the gradient inboxes itself (no dense update_g, no touched map, no O(F / R) scan per step)
    uint32_t* own_uniq = nullptr;     // [kNumSlots][cap_own] shard-local rows
    unsigned int* n_own = nullptr;    // [kNumSlots]
    uint32_t* own_pos = nullptr;      // [kNumSlots][world][cap_own]
    uint32_t* posmap = nullptr;       // [world][Fl] scratch: list position of a row in requester q's current list (self-validating)
    uint8_t* own_mark = nullptr;      // permuted byte map over the shard (128 * own_T)
    size_t cap_own = 0, own_T = 0;
    unsigned long long own_gen[kNumSlots] = {0};
    void* opened[kMaxWorld][kNumHandles] = {{nullptr}};
    bool imported = false;
    unsigned long long epoch = 0;
    unsigned long long gen[kNumSlots] = {0};
    size_t bytes = 0;                 // device memory this module allocated
};

__device__ __forceinline__ unsigned long long* flag_ptr(unsigned char* arena, const ArenaLayout& A, int row, int col) {
    return reinterpret_cast<unsigned long long*>(arena + A.flags) + (size_t)row * kMaxWorld + col;
}
// System-scope fences are executed by the few threads that poll / raise flags, never by whole CTAs: membar.sys drains the
// issuing SM's outstanding peer traffic, and hundreds of CTAs x 256 threads doing it cost ~40 us per kernel (measured: the
// pull / push kernels of an FM C2 step, 3 MB each, took 46 us with per-thread fences).  The CTA barrier before / after
// makes the fence cumulative over the other threads' accesses (PTX memory model: causality order through bar.sync).
__device__ __forceinline__ void wait_flags(unsigned char* my_arena, const ArenaLayout& A, int row, int world,
                                           unsigned long long value) {
    if ((int)threadIdx.x < world) {
        const volatile unsigned long long* f = flag_ptr(my_arena, A, row, threadIdx.x);
        while (*f < value) __nanosleep(40);
        // acquire side: the flag lives in MY memory, whose point of coherence is my L2 -- the peer's row stores landed
        // there before its flag store became visible, and nothing of this kernel has read those rows yet
        __threadfence();
    }
    __syncthreads();
}
// every block: one thread fences the block's stores and counts in; the last block raises flag[row][me] = value on every peer
__device__ int g_dbg_mode = 0;  // LCTR_DIST_DEBUG bit0: serve_pull skips the row copies (timing experiment only)
// Blocks fence their stores at DEVICE scope and count in; only the last block (which has observed every other block's
// count) issues the system-scope fence before raising the flags -- cumulativity carries the other blocks' peer stores
// along.  One membar.sys per kernel instead of one per CTA: 41 -> 24 us for the FM C2 pull (LCTR_DIST_FENCE=sys restores
// the per-block system fences).
__device__ int g_block_fence_sys = 0;
__device__ __forceinline__ void raise_flags_last_block(const PeerTable& P, const ArenaLayout& A, int row, int me, int world,
                                                       unsigned long long value, unsigned int* ctr) {
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (g_block_fence_sys) __threadfence_system(); else __threadfence();
        last = atomicAdd(ctr, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && (int)threadIdx.x < world) {
        __threadfence_system();
        volatile unsigned long long* f = flag_ptr(P.p[threadIdx.x].arena, A, row, me);
        *f = value;
        if (threadIdx.x == 0) *ctr = 0;
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// upload: per owner, the list of shard-local rows -> the owner's key inbox; rows of the exchange are then addressed by
// p = owner * cap_pair + position in that list, on both sides
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
send_keys_kernel(const uint32_t* __restrict__ uniq, const unsigned int* __restrict__ n_uniq, PeerTable P, ArenaLayout A,
                 int me, int world, int slot, int shift, unsigned cap_pair, unsigned int* __restrict__ send_cnt,
                 uint32_t* __restrict__ opos, int* __restrict__ overflow) {
    __shared__ unsigned s_cnt[kMaxWorld], s_base[kMaxWorld];
    const unsigned n = *n_uniq;
    const unsigned mask = (unsigned)world - 1;
    for (unsigned c0 = blockIdx.x * blockDim.x; c0 < n; c0 += gridDim.x * blockDim.x) {
        if (threadIdx.x < kMaxWorld) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const unsigned i = c0 + threadIdx.x;
        uint32_t f = 0;
        unsigned o = 0, rk = 0;
        if (i < n) {
            f = uniq[i];
            o = f & mask;
            rk = atomicAdd(&s_cnt[o], 1u);
        }
        __syncthreads();
        if (threadIdx.x < world && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&send_cnt[threadIdx.x], s_cnt[threadIdx.x]);
        __syncthreads();
        if (i < n) {
            const unsigned j = s_base[o] + rk;
            if (j < cap_pair) {
                uint2* pairs = reinterpret_cast<uint2*>(P.p[o].arena + A.key_inbox + ((size_t)slot * world + me) * A.key_region + 64);
                pairs[j] = make_uint2(f >> shift, i);
                opos[i] = o * cap_pair + j;
            } else {
                *overflow = 1;  // reported at the next host synchronisation; the row index stays in bounds
                opos[i] = o * cap_pair;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) __threadfence_system();  // after the loop's last CTA barrier: cumulative over the CTA's stores
}
// counts into the owners' headers (and kept locally for the push), counters re-armed, then the generation flag of
// (slot, me) on every owner
__global__ void send_keys_finish_kernel(PeerTable P, ArenaLayout A, int me, int world, int slot, int flag_slot, unsigned cap_pair,
                                        unsigned int* send_cnt, unsigned int* seg_cnt, unsigned long long gen) {
    const int o = threadIdx.x;
    __threadfence_system();
    if (o < world) {
        unsigned int* hdr = reinterpret_cast<unsigned int*>(P.p[o].arena + A.key_inbox + ((size_t)slot * world + me) * A.key_region);
        const unsigned n = min(send_cnt[o], cap_pair);
        hdr[0] = n;
        seg_cnt[o] = n;
        send_cnt[o] = 0;
        __threadfence_system();
        volatile unsigned long long* f = flag_ptr(P.p[o].arena, A, FLAG_KEYS + flag_slot, me);
        *f = gen;
    }
    __threadfence_system();
}
// entries: slot -> exchange row p (plain index of the parameter cache; gradient index unless the slot is hot); and the
// replica block of every hot exchange row
__global__ void __launch_bounds__(256)
remap_entries_kernel(const int64_t* __restrict__ hdr, int64_t nnz_arg, const uint32_t* __restrict__ opos,
                     uint32_t* __restrict__ ent_pslot, uint32_t* __restrict__ ent_slot, const uint32_t* __restrict__ hot_of,
                     const unsigned int* __restrict__ n_uniq, uint32_t* __restrict__ hot_p) {
    const int64_t nnz = hdr ? hdr[1] : nnz_arg;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = t0; i < nnz; i += nt) {
        const uint32_t pp = opos[ent_pslot[i]];
        ent_pslot[i] = pp;
        if (!(ent_slot[i] & kHotBit)) ent_slot[i] = pp;
    }
    if (hot_p) {
        const unsigned n = *n_uniq;
        for (int64_t i = t0; i < n; i += nt) {
            const uint32_t h = hot_of[i];
            if (h != 0xffffffffu) hot_p[opos[i]] = h;
        }
    }
}
// undo the hot marks of the previous batch of the slot (hot_p is all ~0 between uploads)
__global__ void __launch_bounds__(256)
clear_hot_p_kernel(const uint32_t* __restrict__ opos_prev, const uint32_t* __restrict__ hot_of_prev, unsigned n_prev,
                   uint32_t* __restrict__ hot_p) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_prev; i += gridDim.x * blockDim.x)
        if (hot_of_prev[i] != 0xffffffffu) hot_p[opos_prev[i]] = 0xffffffffu;
}

// ---------------------------------------------------------------------------------------------------------------
// step 1: owner-driven pull.  Requester r's rows from me land at its cache rows [me * cap_pair, me * cap_pair + n):
// consecutive list positions are consecutive rows, so a warp's stores form long contiguous runs on the wire
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxSl = 4;
__global__ void __launch_bounds__(256)
serve_pull_kernel(PeerTable P, ArenaLayout A, int me, int world, int slot, int flag_slot, unsigned long long gen,
                  unsigned long long epoch, int rowlen, unsigned cap_pair, const float* __restrict__ W,
                  const float* __restrict__ V, unsigned int* done_ctr) {
    cudaTriggerProgrammaticLaunchCompletion();  // the compute kernel behind may be scheduled early: it polls the "rows delivered" flags
    wait_flags(P.p[me].arena, A, FLAG_KEYS + flag_slot, world, gen);  // every requester's key list of this upload has landed
    const int lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const int vec = (rowlen % 4 == 0) ? 4 : 1;
    const int slices = rowlen / vec;
    int lpr = 1;
    while (lpr < slices && lpr < 32) lpr <<= 1;
    const int G = 32 / lpr, q = lane % lpr, g = lane / lpr;
    for (int r = 0; r < world; r++) {
        if (g_dbg_mode & 1) break;
        const unsigned char* region = P.p[me].arena + A.key_inbox + ((size_t)slot * world + r) * A.key_region;
        const unsigned n = *reinterpret_cast<const volatile unsigned int*>(region);
        const uint2* pairs = reinterpret_cast<const uint2*>(region + 64);
        float* cW = reinterpret_cast<float*>(P.p[r].arena + A.cacheW) + (size_t)me * cap_pair;
        float* cV = reinterpret_cast<float*>(P.p[r].arena + A.cacheV) + (size_t)me * cap_pair * rowlen;
        if (vec == 4 && slices <= lpr * kMaxSl) {
            constexpr int U = 4;  // row groups in flight before the first (posted) store leaves
            for (unsigned b0 = warp * (G * U); b0 < n; b0 += nwarps * (G * U)) {
                unsigned l[U];
                float4 v[U][kMaxSl];
                float w[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const unsigned j = b0 + u * G + g;
                    l[u] = j < n ? pairs[j].x : 0xffffffffu;
                    if (l[u] == 0xffffffffu) continue;
                    const float* src = V + (size_t)l[u] * rowlen;
#pragma unroll
                    for (int i = 0; i < kMaxSl; i++) {
                        const int sl = q + i * lpr;
                        if (sl < slices) v[u][i] = *reinterpret_cast<const float4*>(src + 4 * sl);
                    }
                    if (q == 0) w[u] = W[l[u]];
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (l[u] == 0xffffffffu) continue;
                    const unsigned j = b0 + u * G + g;
                    float* dst = cV + (size_t)j * rowlen;
#pragma unroll
                    for (int i = 0; i < kMaxSl; i++) {
                        const int sl = q + i * lpr;
                        if (sl < slices) *reinterpret_cast<float4*>(dst + 4 * sl) = v[u][i];
                    }
                    if (q == 0) cW[j] = w[u];
                }
            }
        } else {
            for (unsigned j = warp * G + g; j < n; j += nwarps * G) {  // generic fallback (odd row lengths)
                const float* src = V + (size_t)pairs[j].x * rowlen;
                float* dst = cV + (size_t)j * rowlen;
                for (int sl = q * vec; sl < rowlen; sl += lpr * vec)
                    for (int c = 0; c < vec; c++) dst[sl + c] = src[sl + c];
                if (q == 0) cW[j] = W[pairs[j].x];
            }
        }
    }
    raise_flags_last_block(P, A, FLAG_PULLED, me, world, epoch, done_ctr);
}

// stand-alone wait for compute kernels without an in-kernel wait (FFM / non-fused FM and NFM)
__global__ void wait_flags_kernel(PeerTable P, ArenaLayout A, int me, int row, int world, unsigned long long value) {
    wait_flags(P.p[me].arena, A, row, world, value);
}

// ---------------------------------------------------------------------------------------------------------------
// step 3: push.  The gradient rows of owner o are rows [o * cap_pair, o * cap_pair + n_o) of my compact buffer, in the
// order of the list o received: one contiguous stream of records [gV (rowlen) | gW | pad] into o's gradient inbox.
// ---------------------------------------------------------------------------------------------------------------
// gv / gw: row p at gv + p * gvs (rowlen floats) and gw + p * gws.  hot_p (fused FM / NFM kernels): rows whose entries were
// accumulated in kHotRep replica rows of Ghot -- folded here.  Local rows (and replicas) are re-zeroed.
__global__ void __launch_bounds__(256)
push_rows_kernel(const unsigned int* __restrict__ seg_cnt, float* __restrict__ gv, int gvs, float* __restrict__ gw, int gws,
                 const uint32_t* __restrict__ hot_p, float* __restrict__ Ghot, int GS, int rowlen, int recw, unsigned cap_pair,
                 PeerTable P, ArenaLayout A, int me, int world, unsigned long long epoch, unsigned int* done_ctr) {
    const int lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const int slices = rowlen / 4;  // rowlen % 4 == 0 is required by dist_alloc
    int lpr = 1;
    while (lpr < slices && lpr < 32) lpr <<= 1;
    const int G = 32 / lpr, q = lane % lpr, g = lane / lpr;
    cudaTriggerProgrammaticLaunchCompletion();  // the owner-side kernel behind polls the "pushes landed" flags
    cudaGridDependencySynchronize();            // launched dependent on the gradient kernel: its G / Ghot are complete from here
    for (int o = 0; o < world; o++) {
        const unsigned n = seg_cnt[o];
        float* inbox = reinterpret_cast<float*>(P.p[o].arena + A.grad_inbox + (size_t)me * A.grad_region);
        for (unsigned j = warp * G + g; j < n; j += nwarps * G) {
            const size_t pr = (size_t)o * cap_pair + j;
            float* src = gv + pr * gvs;
            float* dst = inbox + (size_t)j * recw;
            const uint32_t h = hot_p ? __ldg(hot_p + pr) : 0xffffffffu;
            float gwv = q == 0 ? gw[pr * gws] : 0.f;
            for (int sl = q; sl < slices; sl += lpr) {
                float4 v = *reinterpret_cast<const float4*>(src + 4 * sl);
                if (h != 0xffffffffu) {  // fold (and re-zero) the replica rows of a hot slot: column block sl; 16 loads in flight
                    float* tile = Ghot + (size_t)h * kHotRep * GS;
                    for (int r0 = 0; r0 < kHotRep; r0 += 16) {
                        float4 t[16];
#pragma unroll
                        for (int i = 0; i < 16; i++) t[i] = __ldcg(reinterpret_cast<const float4*>(tile + (size_t)(r0 + i) * GS + 4 * sl));
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            v.x += t[i].x; v.y += t[i].y; v.z += t[i].z; v.w += t[i].w;
                            *reinterpret_cast<float4*>(tile + (size_t)(r0 + i) * GS + 4 * sl) = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
                *reinterpret_cast<float4*>(dst + 4 * sl) = v;
                *reinterpret_cast<float4*>(src + 4 * sl) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (q == 0) {
                if (h != 0xffffffffu) {
                    float* tile = Ghot + (size_t)h * kHotRep * GS;
                    float t[kHotRep];
#pragma unroll
                    for (int rp = 0; rp < kHotRep; rp++) t[rp] = __ldcg(tile + (size_t)rp * GS + rowlen);
#pragma unroll
                    for (int rp = 0; rp < kHotRep; rp++) { gwv += t[rp]; tile[(size_t)rp * GS + rowlen] = 0.f; }
                }
                *reinterpret_cast<float4*>(dst + rowlen) = make_float4(gwv, 0.f, 0.f, 0.f);  // whole 16 B: the record leaves as full slices
                gw[pr * gws] = 0.f;
            }
        }
    }
    raise_flags_last_block(P, A, FLAG_PUSHED, me, world, epoch, done_ctr);
}

// ---------------------------------------------------------------------------------------------------------------
// step 4: owner folds the R gradient inboxes into its shard's update_g (local REDs) and marks the rows
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
merge_kernel(PeerTable P, ArenaLayout A, int me, int world, int slot, unsigned long long epoch, int rowlen, int recw,
             float* __restrict__ gW, float* __restrict__ gV, uint8_t* __restrict__ touched) {
    wait_flags(P.p[me].arena, A, FLAG_PUSHED, world, epoch);  // every requester's records have landed
    const int lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const int slices = rowlen / 4;
    int lpr = 1;
    while (lpr < slices && lpr < 32) lpr <<= 1;
    const int G = 32 / lpr, q = lane % lpr, g = lane / lpr;
    for (int src = 0; src < world; src++) {
        const unsigned char* region = P.p[me].arena + A.key_inbox + ((size_t)slot * world + src) * A.key_region;
        const unsigned n = *reinterpret_cast<const volatile unsigned int*>(region);
        const uint2* pairs = reinterpret_cast<const uint2*>(region + 64);
        const float* recs = reinterpret_cast<const float*>(P.p[me].arena + A.grad_inbox + (size_t)src * A.grad_region);
        for (unsigned j = warp * G + g; j < n; j += nwarps * G) {
            const float* rec = recs + (size_t)j * recw;
            const size_t l = pairs[j].x;
            float* gdst = gV + l * (size_t)rowlen;
            bool any = false;
            for (int sl = q; sl < slices; sl += lpr) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(rec + 4 * sl));
                if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) { red_add_v4(gdst + 4 * sl, v); any = true; }
            }
            if (q == 0) {
                const float w = __ldcg(rec + rowlen);
                if (w != 0.f) { red_add_f32(gW + l, w); any = true; }
            }
            if (any) touched[l] = 1;  // (idempotent byte store; every lane that added something marks the row)
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fused FM / NFM path, owner side.  Once per upload: the union of the requesters' key lists (mark -> compact, the slot-map
// kernels of fm_fused.cuh on the shard) and each union row's position in every list.  Every step: ONE kernel that waits for
// the pushes, sums a row's records over the requesters in rank order (deterministic, no REDs) and applies the updater.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
own_mark_kernel(PeerTable P, ArenaLayout A, int me, int world, int slot2, int flag_slot, unsigned long long gen,
                uint32_t* __restrict__ posmap, size_t Fl, uint8_t* __restrict__ mark, size_t T) {
    wait_flags(P.p[me].arena, A, FLAG_KEYS + flag_slot, world, gen);  // every requester's key list of this upload has landed
    for (int q = 0; q < world; q++) {
        const unsigned char* region = P.p[me].arena + A.key_inbox + ((size_t)slot2 * world + q) * A.key_region;
        const unsigned n = *reinterpret_cast<const volatile unsigned int*>(region);
        const uint2* pairs = reinterpret_cast<const uint2*>(region + 64);
        for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
            const uint32_t l = pairs[j].x;
            posmap[(size_t)q * Fl + l] = j;  // never cleared: an entry counts only if list[entry] == row (own_pos_kernel)
            mark[(size_t)(l & 127u) * T + (l >> 7)] = 1;
        }
    }
}
__global__ void __launch_bounds__(256)
own_pos_kernel(PeerTable P, ArenaLayout A, int me, int world, int slot2, const uint32_t* __restrict__ own_uniq,
               const unsigned int* __restrict__ n_own, const uint32_t* __restrict__ posmap, size_t Fl,
               uint32_t* __restrict__ own_pos, unsigned cap_own) {
    const unsigned total = min(*n_own, cap_own);
    for (int q = 0; q < world; q++) {
        const unsigned char* region = P.p[me].arena + A.key_inbox + ((size_t)slot2 * world + q) * A.key_region;
        const unsigned n = *reinterpret_cast<const volatile unsigned int*>(region);
        const uint2* pairs = reinterpret_cast<const uint2*>(region + 64);
        for (unsigned u = blockIdx.x * blockDim.x + threadIdx.x; u < total; u += gridDim.x * blockDim.x) {
            const uint32_t l = own_uniq[u];
            const uint32_t j = posmap[(size_t)q * Fl + l];
            own_pos[(size_t)q * cap_own + u] = (j < n && pairs[j].x == l) ? j : 0xffffffffu;
        }
    }
}
template <int K, int OPT>
__global__ void __launch_bounds__(256)
merge_apply_kernel(PeerTable P, ArenaLayout A, int me, int world, unsigned long long epoch, int recw,
                   const uint32_t* __restrict__ own_uniq, const unsigned int* __restrict__ n_own,
                   const uint32_t* __restrict__ own_pos, unsigned cap_own, float* __restrict__ W, float* __restrict__ V,
                   float* __restrict__ s1W, float* __restrict__ s1V, float* __restrict__ s2W, float* __restrict__ s2V, OptParams Pp) {
    wait_flags(P.p[me].arena, A, FLAG_PUSHED, world, epoch);  // every requester's records of this step have landed
    cudaGridDependencySynchronize();  // (launched dependent on my own push kernel; its flag to myself is already in)
    constexpr int LPR = K / 4, GR = 32 / LPR;
    constexpr bool two = OPT == LCTR_OPT_FTRL || OPT == LCTR_OPT_ADAM || OPT == LCTR_OPT_ADADELTA || OPT == LCTR_OPT_PS_DCASGD ||
                         OPT == LCTR_OPT_PS_DCASGDA;
    Pp.opt = OPT;
    const int lane = threadIdx.x & 31, q = lane % LPR, g = lane / LPR;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const unsigned total = min(*n_own, cap_own);
    for (unsigned b0 = warp * GR; b0 < total; b0 += nwarps * GR) {
        const unsigned idx = b0 + g;
        if (idx >= total) continue;
        const uint32_t l = own_uniq[idx];
        const size_t o = (size_t)l * K + 4 * q;
        float4 v4 = *reinterpret_cast<const float4*>(V + o), a4 = *reinterpret_cast<const float4*>(s1V + o);
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (two) b4 = *reinterpret_cast<const float4*>(s2V + o);
        float w = 0.f, a = 0.f, bb = 0.f;
        if (q == 0) { w = W[l]; a = s1W[l]; if (two) bb = s2W[l]; }
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float gw = 0.f;
        for (int src = 0; src < world; src++) {  // rank order: the sum is reproducible
            const uint32_t j = own_pos[(size_t)src * cap_own + idx];
            if (j == 0xffffffffu) continue;
            const float* rec = reinterpret_cast<const float*>(P.p[me].arena + A.grad_inbox + (size_t)src * A.grad_region) + (size_t)j * recw;
            const float4 t = __ldcg(reinterpret_cast<const float4*>(rec + 4 * q));
            g4.x += t.x; g4.y += t.y; g4.z += t.z; g4.w += t.w;
            if (q == 0) gw += __ldcg(rec + K);
        }
        update_one(Pp, Pp.corrV, v4.x, g4.x, a4.x, b4.x);
        update_one(Pp, Pp.corrV, v4.y, g4.y, a4.y, b4.y);
        update_one(Pp, Pp.corrV, v4.z, g4.z, a4.z, b4.z);
        update_one(Pp, Pp.corrV, v4.w, g4.w, a4.w, b4.w);
        *reinterpret_cast<float4*>(V + o) = v4;
        *reinterpret_cast<float4*>(s1V + o) = a4;
        if (two) *reinterpret_cast<float4*>(s2V + o) = b4;
        if (q == 0) {
            update_one(Pp, Pp.corrW, w, gw, a, bb);
            W[l] = w; s1W[l] = a;
            if (two) s2W[l] = bb;
        }
    }
}

template <int K>
static void merge_apply_go(lctr_ctx* c, DistState* d, int slot, const OptParams& Pp, unsigned grid) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_on() ? 1 : 0;  // behind my push kernel
#define MA_GO(OPTC)                                                                                                               \
    cudaLaunchKernelEx(&cfg, merge_apply_kernel<K, OPTC>, d->peers, d->A, d->rank, d->world, d->epoch, d->recw,                    \
        (const uint32_t*)(d->own_uniq + (size_t)slot * d->cap_own), (const unsigned int*)(d->n_own + slot),                        \
        (const uint32_t*)(d->own_pos + (size_t)slot * d->world * d->cap_own), (unsigned)d->cap_own, c->W, c->V, c->s1W, c->s1V,   \
        c->s2W, c->s2V, Pp)
    switch (Pp.opt) {
        case LCTR_OPT_ADAGRAD: MA_GO(LCTR_OPT_ADAGRAD); break;
        case LCTR_OPT_FTRL: MA_GO(LCTR_OPT_FTRL); break;
        case LCTR_OPT_ADAM: MA_GO(LCTR_OPT_ADAM); break;
        case LCTR_OPT_RMSPROP: MA_GO(LCTR_OPT_RMSPROP); break;
        case LCTR_OPT_ADADELTA: MA_GO(LCTR_OPT_ADADELTA); break;
        case LCTR_OPT_PS_SGD: MA_GO(LCTR_OPT_PS_SGD); break;
        case LCTR_OPT_PS_ADAGRAD: MA_GO(LCTR_OPT_PS_ADAGRAD); break;
        case LCTR_OPT_PS_DCASGD: MA_GO(LCTR_OPT_PS_DCASGD); break;
        default: MA_GO(LCTR_OPT_PS_DCASGDA); break;
    }
#undef MA_GO
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int dist_alloc(lctr_ctx* c) {
    const int R = c->cfg.world;
    LCTR_CHECK(R <= kMaxWorld && (R & (R - 1)) == 0, "world=%d: need a power of two <= %d", R, kMaxWorld);
    DistState* d = new DistState();
    c->dist = d;
    d->rank = c->cfg.rank; d->world = R;
    while ((1 << d->shift) < R) d->shift++;
    if (const char* de = getenv("LCTR_DIST_DEBUG")) {
        const int m = atoi(de);
        LCTR_CUDA(cudaMemcpyToSymbol(g_dbg_mode, &m, sizeof(int)));
    }
    if (const char* fe = getenv("LCTR_DIST_FENCE")) {
        const int sys = strcmp(fe, "gpu") != 0;
        LCTR_CUDA(cudaMemcpyToSymbol(g_block_fence_sys, &sys, sizeof(int)));
    }
    LCTR_CHECK(c->rowlen % 4 == 0, "multi-GPU exchange moves rows in 16 B slices: rowlen %zu must be a multiple of 4", c->rowlen);
    // keys of one batch: at most its entry count (cfg.max_nnz), at most the id space
    d->cap_keys = c->cfg.max_nnz ? std::min<size_t>(c->cfg.max_nnz, c->F) : c->F;
    // keys one requester sends one owner: U / R on average (owner = fid mod R, binomially tight); 1.5x that plus slack
    d->cap_pair = std::min<size_t>(c->Fl, 3 * d->cap_keys / (2 * R) + 4096);
    d->rows_x = (size_t)R * d->cap_pair;
    d->recw = (int)((c->rowlen + 4 + 3) / 4 * 4);
    ArenaLayout& A = d->A;
    size_t off = 0;
    A.flags = off; off = align_up(off + (size_t)(3 + kNumSlots) * kMaxWorld * sizeof(unsigned long long), 256);
    A.key_region = align_up(64 + d->cap_pair * sizeof(uint2), 256);
    A.key_inbox = off; off += A.key_region * kNumSlots * 2 * R;
    A.cacheW = off; off = align_up(off + d->rows_x * sizeof(float), 256);
    A.cacheV = off; off = align_up(off + d->rows_x * c->rowlen * sizeof(float), 256);
    A.grad_region = align_up(d->cap_pair * (size_t)d->recw * sizeof(float), 256);
    A.grad_inbox = off; off += A.grad_region * R;
    A.total = off;
    LCTR_CUDA(cudaMalloc((void**)&d->arena, A.total));
    LCTR_CUDA(cudaMemsetAsync(d->arena, 0, A.total, c->stream));
    d->bytes += A.total;
    LCTR_CUDA(cudaMalloc((void**)&d->send_cnt, kMaxWorld * sizeof(unsigned int)));
    LCTR_CUDA(cudaMemsetAsync(d->send_cnt, 0, kMaxWorld * sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMalloc((void**)&d->seg_cnt, (size_t)kNumSlots * kMaxWorld * sizeof(unsigned int)));
    LCTR_CUDA(cudaMemsetAsync(d->seg_cnt, 0, (size_t)kNumSlots * kMaxWorld * sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMalloc((void**)&d->opos, (size_t)kNumSlots * d->cap_keys * sizeof(uint32_t)));
    d->bytes += (size_t)kNumSlots * d->cap_keys * sizeof(uint32_t);
    LCTR_CUDA(cudaMalloc((void**)&d->done_ctr, 4 * sizeof(unsigned int)));
    LCTR_CUDA(cudaMemsetAsync(d->done_ctr, 0, 4 * sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMalloc((void**)&d->overflow, sizeof(int)));
    LCTR_CUDA(cudaMemsetAsync(d->overflow, 0, sizeof(int), c->stream));
    if (fused_kernels_ok(c)) {  // the fused FM / NFM kernels keep their gradients in fm_fused.cu's G / Ghot (rows p)
        LCTR_CUDA(cudaMalloc((void**)&d->hot_p, (size_t)kNumSlots * d->rows_x * sizeof(uint32_t)));
        LCTR_CUDA(cudaMemsetAsync(d->hot_p, 0xff, (size_t)kNumSlots * d->rows_x * sizeof(uint32_t), c->stream));
        d->bytes += (size_t)kNumSlots * d->rows_x * sizeof(uint32_t);
        d->cap_own = std::min<size_t>(c->Fl, d->rows_x);
        d->own_T = (c->Fl + 127) / 128;  // rows of the permuted byte map (fm_fused.cuh: mark_rows)
        LCTR_CUDA(cudaMalloc((void**)&d->own_uniq, (size_t)kNumSlots * d->cap_own * sizeof(uint32_t)));
        LCTR_CUDA(cudaMalloc((void**)&d->own_pos, (size_t)kNumSlots * R * d->cap_own * sizeof(uint32_t)));
        LCTR_CUDA(cudaMalloc((void**)&d->n_own, kNumSlots * sizeof(unsigned int)));
        LCTR_CUDA(cudaMemsetAsync(d->n_own, 0, kNumSlots * sizeof(unsigned int), c->stream));
        LCTR_CUDA(cudaMalloc((void**)&d->posmap, (size_t)R * c->Fl * sizeof(uint32_t)));
        LCTR_CUDA(cudaMemsetAsync(d->posmap, 0xff, (size_t)R * c->Fl * sizeof(uint32_t), c->stream));
        LCTR_CUDA(cudaMalloc((void**)&d->own_mark, 128 * d->own_T));
        LCTR_CUDA(cudaMemsetAsync(d->own_mark, 0, 128 * d->own_T, c->stream));
        d->bytes += (size_t)kNumSlots * (R + 1) * d->cap_own * sizeof(uint32_t) + (size_t)R * c->Fl * sizeof(uint32_t) + 128 * d->own_T;
    } else {
        LCTR_CUDA(cudaMalloc((void**)&d->cgW, d->rows_x * sizeof(float)));
        LCTR_CUDA(cudaMalloc((void**)&d->cgV, d->rows_x * c->rowlen * sizeof(float)));
        LCTR_CUDA(cudaMemsetAsync(d->cgW, 0, d->rows_x * sizeof(float), c->stream));
        LCTR_CUDA(cudaMemsetAsync(d->cgV, 0, d->rows_x * c->rowlen * sizeof(float), c->stream));
        d->bytes += d->rows_x * (c->rowlen + 1) * sizeof(float);
    }
    c->dist_rows = d->rows_x;
    // compute view: the batch-compact cache and gradient rows (indexed by exchange row p)
    c->cW = reinterpret_cast<float*>(d->arena + A.cacheW);
    c->cV = reinterpret_cast<float*>(d->arena + A.cacheV);
    c->cgW = d->cgW;
    c->cgV = d->cgV;
    memset(&d->peers, 0, sizeof(d->peers));
    Peer& me = d->peers.p[d->rank];
    me.W = c->W; me.V = c->V; me.gW = c->gW; me.gV = c->gV; me.touched = c->touched; me.arena = d->arena;
    return 0;
}

int dist_free(lctr_ctx* c) {
    DistState* d = c->dist;
    if (!d) return 0;
    for (int r = 0; r < d->world; r++)
        for (int j = 0; j < kNumHandles; j++)
            if (d->opened[r][j]) cudaIpcCloseMemHandle(d->opened[r][j]);
    cudaFree(d->arena); cudaFree(d->send_cnt); cudaFree(d->seg_cnt); cudaFree(d->opos); cudaFree(d->done_ctr); cudaFree(d->overflow);
    if (d->hot_p) cudaFree(d->hot_p);
    if (d->own_uniq) { cudaFree(d->own_uniq); cudaFree(d->own_pos); cudaFree(d->n_own); cudaFree(d->posmap); cudaFree(d->own_mark); }
    if (d->cgW) cudaFree(d->cgW);
    if (d->cgV) cudaFree(d->cgV);
    c->cW = c->cV = c->cgW = c->cgV = nullptr;
    delete d;
    c->dist = nullptr;
    return 0;
}

size_t dist_bytes(const lctr_ctx* c) { return c->dist ? c->dist->bytes : 0; }

void dist_wait_info(lctr_ctx* c, const unsigned long long** flags, int* n, unsigned long long* epoch) {
    DistState* d = c->dist;
    *flags = reinterpret_cast<const unsigned long long*>(d->arena + d->A.flags) + (size_t)FLAG_PULLED * kMaxWorld;
    *n = d->world;
    *epoch = d->epoch;
}

// key lists of the slot's batch -> the owners' inboxes (after the slot map of fm_fused.cu has been built on `st`)
int dist_send_keys(lctr_ctx* c, Slot& s, int slot, cudaStream_t st) {
    DistState* d = c->dist;
    LCTR_CHECK(d->imported, "multi-GPU upload before lctr_ipc_import");
    LCTR_CHECK((size_t)std::min<int64_t>(s.nnz, (int64_t)c->F) <= d->cap_keys,
               "batch of %lld entries exceeds the key capacity %zu of the multi-GPU context (cfg.max_nnz)", (long long)s.nnz, d->cap_keys);
    d->gen[slot]++;
    const int slot2 = slot * 2 + (int)(d->gen[slot] & 1);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((std::min<int64_t>(s.nnz, (int64_t)c->F) + 255) / 256, (int64_t)c->sm_count * 4));
    uint32_t* opos = d->opos + (size_t)slot * d->cap_keys;
    uint32_t* hot_p = d->hot_p ? d->hot_p + (size_t)slot * d->rows_x : nullptr;
    if (hot_p) LCTR_CUDA(cudaMemsetAsync(hot_p, 0xff, d->rows_x * sizeof(uint32_t), st));  // the previous batch's hot rows
    send_keys_kernel<<<grid, 256, 0, st>>>(s.uniq, s.n_uniq, d->peers, d->A, d->rank, d->world, slot2, d->shift, (unsigned)d->cap_pair,
                                           d->send_cnt, opos, d->overflow);
    send_keys_finish_kernel<<<1, 32, 0, st>>>(d->peers, d->A, d->rank, d->world, slot2, slot, (unsigned)d->cap_pair, d->send_cnt,
                                              d->seg_cnt + (size_t)slot * kMaxWorld, d->gen[slot]);
    const unsigned rg = (unsigned)std::max<int64_t>(1, std::min<int64_t>((s.nnz + 255) / 256, (int64_t)c->sm_count * 8));
    remap_entries_kernel<<<rg, 256, 0, st>>>(nullptr, s.nnz, opos, s.ent_pslot, s.ent_slot, hot_p ? s.hot_of : nullptr, s.n_uniq, hot_p);
    c->launches += 3;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

// host-visible check (called where the host synchronises anyway): a key list that outgrew its inbox region is an error,
// never a silent drop
int dist_check_overflow(lctr_ctx* c) {
    DistState* d = c->dist;
    int h = 0;
    LCTR_CUDA(cudaMemcpyAsync(&h, d->overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    LCTR_CHECK(h == 0, "multi-GPU: a per-owner key list outgrew its inbox (%zu records); raise cfg.max_nnz", d->cap_pair);
    return 0;
}

int dist_pre_step(lctr_ctx* c, Slot& s, int slot, bool in_kernel_wait) {
    DistState* d = c->dist;
    LCTR_CHECK(d->imported, "multi-GPU step before lctr_ipc_import");
    LCTR_CHECK(s.fused_valid, "multi-GPU step on a slot without its key set");
    d->epoch++;
    if (d->own_uniq && d->own_gen[slot] != d->gen[slot]) {  // first step on this upload of the slot: the owner-side union
        const int slot2 = slot * 2 + (int)(d->gen[slot] & 1);
        const unsigned g1 = (unsigned)std::max<int64_t>(8, std::min<int64_t>((int64_t)c->sm_count * 4, ((int64_t)d->cap_pair + 255) / 256));
        LCTR_CUDA(cudaMemsetAsync(d->n_own + slot, 0, sizeof(unsigned int), c->stream));
        own_mark_kernel<<<g1, 256, 0, c->stream>>>(d->peers, d->A, d->rank, d->world, slot2, slot, d->gen[slot], d->posmap, c->Fl,
                                                   d->own_mark, d->own_T);
        launch_slotmap_compact(c, d->own_mark, d->own_T, d->own_uniq + (size_t)slot * d->cap_own, d->n_own + slot, c->stream);
        own_pos_kernel<<<g1, 256, 0, c->stream>>>(d->peers, d->A, d->rank, d->world, slot2, d->own_uniq + (size_t)slot * d->cap_own,
                                                  d->n_own + slot, d->posmap, c->Fl, d->own_pos + (size_t)slot * d->world * d->cap_own,
                                                  (unsigned)d->cap_own);
        c->launches += 3;
        d->own_gen[slot] = d->gen[slot];
    }
    { ProfScope prof(c, PROF_DIST_PULL);
    // rows this rank serves ~ the union of what R requesters ask of it ~ (keys of a batch): one warp iteration = 32 rows (FM)
    const unsigned pull_grid = (unsigned)std::max<int64_t>(8, std::min<int64_t>((int64_t)c->sm_count * 4,
        (std::min<int64_t>(s.nnz, (int64_t)c->F) * (int64_t)std::max<size_t>(1, c->rowlen / 16) + 255) / 256));
    serve_pull_kernel<<<pull_grid, 256, 0, c->stream>>>(d->peers, d->A, d->rank, d->world, slot * 2 + (int)(d->gen[slot] & 1), slot,
                                                             d->gen[slot], d->epoch, (int)c->rowlen, (unsigned)d->cap_pair, c->W, c->V,
                                                             d->done_ctr + 0); }
    c->launches++;
    if (!in_kernel_wait) {
        ProfScope prof(c, PROF_DIST_BAR1);
        wait_flags_kernel<<<1, 32, 0, c->stream>>>(d->peers, d->A, d->rank, FLAG_PULLED, d->world, d->epoch);
        c->launches++;
    }
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

int dist_post_step(lctr_ctx* c, Slot& s, int slot, int64_t rows_divisor) {
    DistState* d = c->dist;
    const unsigned xgrid = (unsigned)std::max<int64_t>(8, std::min<int64_t>((int64_t)c->sm_count * 4,
        (std::min<int64_t>(s.nnz, (int64_t)c->F) * (int64_t)std::max<size_t>(1, c->rowlen / 16) + 255) / 256));
    const unsigned int* seg = d->seg_cnt + (size_t)slot * kMaxWorld;
    { ProfScope prof(c, PROF_DIST_PUSH);
    if (fused_kernels_ok(c)) {
        FusedState* f = c->fused;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(xgrid); cfg.blockDim = dim3(256); cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_on() ? 1 : 0;  // behind the gradient kernel
        cudaLaunchKernelEx(&cfg, push_rows_kernel, seg, f->G, f->GS, f->G + c->rowlen, f->GS,
                           (const uint32_t*)(d->hot_p + (size_t)slot * d->rows_x), f->Ghot, f->GS, (int)c->rowlen, d->recw,
                           (unsigned)d->cap_pair, d->peers, d->A, d->rank, d->world, d->epoch, d->done_ctr + 1);
    } else {
        push_rows_kernel<<<xgrid, 256, 0, c->stream>>>(seg, d->cgV, (int)c->rowlen, d->cgW, 1, nullptr, nullptr, 0, (int)c->rowlen,
                                                                 d->recw, (unsigned)d->cap_pair, d->peers, d->A, d->rank, d->world, d->epoch,
                                                                 d->done_ctr + 1);
    } }
    if (d->own_uniq) {  // fused FM / NFM: merge + updater in one kernel over the owner-side union of the key lists
        ProfScope prof(c, PROF_DIST_MERGE);
        const OptParams Pp = make_opt_params(c, rows_divisor);
        const unsigned mgrid = (unsigned)c->sm_count * 4;
        switch ((int)c->cfg.factor_cnt) {
            case 4: merge_apply_go<4>(c, d, slot, Pp, mgrid); break;
            case 8: merge_apply_go<8>(c, d, slot, Pp, mgrid); break;
            case 16: merge_apply_go<16>(c, d, slot, Pp, mgrid); break;
            default: merge_apply_go<32>(c, d, slot, Pp, mgrid); break;
        }
        c->launches += 2;
        LCTR_CUDA(cudaGetLastError());
        return 0;
    }
    { ProfScope prof(c, PROF_DIST_MERGE);
    merge_kernel<<<xgrid, 256, 0, c->stream>>>(d->peers, d->A, d->rank, d->world, slot * 2 + (int)(d->gen[slot] & 1), d->epoch,
                                                         (int)c->rowlen, d->recw, c->gW, c->gV, c->touched); }
    c->launches += 2;
    LCTR_CUDA(cudaGetLastError());
    return launch_apply(c, rows_divisor);  // sparse updater on the shard (the merge waited for every requester's pushes)
}

}  // namespace lctr

using namespace lctr;

extern "C" {

// handles exported per rank, in this order: W, V shards, update_g W, V shards, touched map, arena
int lctr_ipc_export(lctr_ctx* c, void* handles_out, size_t cap, size_t* bytes) {
    LCTR_CHECK(c && bytes, "null argument");
    LCTR_CHECK(c->dist, "lctr_ipc_export: ctx was created with world == 1");
    const size_t need = kNumHandles * sizeof(cudaIpcMemHandle_t);
    *bytes = need;
    if (!handles_out) return 0;
    LCTR_CHECK(cap >= need, "lctr_ipc_export: need %zu bytes", need);
    cudaIpcMemHandle_t* h = reinterpret_cast<cudaIpcMemHandle_t*>(handles_out);
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[0], c->W));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[1], c->V));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[2], c->gW));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[3], c->gV));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[4], c->touched));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[5], c->dist->arena));
    return 0;
}

int lctr_ipc_import(lctr_ctx* c, const void* all_handles, size_t bytes_per_rank) {
    LCTR_CHECK(c && all_handles, "null argument");
    LCTR_CHECK(c->dist, "lctr_ipc_import: ctx was created with world == 1");
    LCTR_CHECK(bytes_per_rank == kNumHandles * sizeof(cudaIpcMemHandle_t), "lctr_ipc_import: bytes_per_rank %zu", bytes_per_rank);
    DistState* d = c->dist;
    const unsigned char* base = reinterpret_cast<const unsigned char*>(all_handles);
    for (int r = 0; r < d->world; r++) {
        if (r == d->rank) continue;
        const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(base + (size_t)r * bytes_per_rank);
        for (int j = 0; j < kNumHandles; j++) {
            cudaIpcMemHandle_t hh;
            memcpy(&hh, &h[j], sizeof(hh));
            LCTR_CUDA(cudaIpcOpenMemHandle(&d->opened[r][j], hh, cudaIpcMemLazyEnablePeerAccess));
        }
        Peer& p = d->peers.p[r];
        p.W = (float*)d->opened[r][0];
        p.V = (float*)d->opened[r][1];
        p.gW = (float*)d->opened[r][2];
        p.gV = (float*)d->opened[r][3];
        p.touched = (uint8_t*)d->opened[r][4];
        p.arena = (unsigned char*)d->opened[r][5];
    }
    d->imported = true;
    return 0;
}

/* device memory of this context in bytes: table shard + updater state + multi-GPU arena / caches */
int lctr_device_bytes(lctr_ctx* c, uint64_t* shard_bytes, uint64_t* exchange_bytes) {
    LCTR_CHECK(c, "null ctx");
    const bool two = c->s2W != nullptr;
    if (shard_bytes) *shard_bytes = (uint64_t)(c->Fl * (c->rowlen + 1) * sizeof(float) * (two ? 4 : 3) + c->Fl);
    if (exchange_bytes) *exchange_bytes = (uint64_t)dist_bytes(c);
    return 0;
}

}  // extern "C"
