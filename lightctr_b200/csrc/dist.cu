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
//     slot map of my batch (mark / compact / assign) and, per owner o, the list {shard-local row, my slot} of the keys
//     o owns, written straight into o's key inbox with posted stores + a generation flag        (send_keys_kernel)
//   step
//     1 serve   OWNER-driven pull: I read the key lists my peers posted and WRITE the rows they asked for into their
//               caches (posted peer stores instead of read round trips), then raise "rows delivered" on each peer
//                                                                                                (serve_pull_kernel)
//     2 compute the single-GPU kernels on the cache; the first one waits (in-kernel) for every owner's flag
//     3 push    FM (64 B rows): the slot's gradient row (hot replicas folded) is added into the owner's update_g with
//               peer vector REDs; FFM (1.2 KB rows): plain posted stores into the owner's gradient inbox at the
//               position of the key in the list it received -- no slot reservation; then "pushes landed" flags
//                                                                              (push_fused_kernel / push_rows_kernel)
//     4 owner   wide rows: inbox rows are added into update_g (local REDs)                        (merge_kernel)
//               sparse updater on the shard; its first kernel waits for every requester's flag   (opt.cu)
// Launches per step: 5 (FM) / 6 (FFM); the barriers of the r01 protocol are flags written at the tail of one kernel and
// polled at the head of the next -- no barrier launches, no host involvement.
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
    size_t cacheW;       // cap_keys floats
    size_t cacheV;       // cap_keys x rowlen floats
    size_t grad_inbox;   // [world] regions of cap_pair x recw floats (wide rows only)
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
    bool mailbox = false;             // wide rows: stores into the owner's inbox + owner-side merge
    unsigned int* send_cnt = nullptr; // [kMaxWorld] records appended per owner by the running send_keys
    uint32_t* opos = nullptr;         // [kNumSlots][cap_keys]: position of my slot's key in its owner's list
    unsigned int* done_ctr = nullptr; // [4] last-block counters
    int* overflow = nullptr;          // device flag: a key list outgrew cap_pair
    float *cgV = nullptr, *cgW = nullptr;  // [cap_keys][rowlen], [cap_keys]: compact gradient rows of the non-fused kernels
    void* opened[kMaxWorld][kNumHandles] = {{nullptr}};
    bool imported = false;
    unsigned long long epoch = 0;
    unsigned long long gen[kNumSlots] = {0};
    size_t bytes = 0;                 // device memory this module allocated
};

__device__ __forceinline__ unsigned long long* flag_ptr(unsigned char* arena, const ArenaLayout& A, int row, int col) {
    return reinterpret_cast<unsigned long long*>(arena + A.flags) + (size_t)row * kMaxWorld + col;
}
__device__ __forceinline__ void wait_flags(unsigned char* my_arena, const ArenaLayout& A, int row, int world,
                                           unsigned long long value) {
    if ((int)threadIdx.x < world) {
        const volatile unsigned long long* f = flag_ptr(my_arena, A, row, threadIdx.x);
        while (*f < value) __nanosleep(40);
    }
    __syncthreads();
    __threadfence_system();
}
// every block: fence its stores, count in; the last block raises flag[row][me] = value on every peer
__device__ __forceinline__ void raise_flags_last_block(const PeerTable& P, const ArenaLayout& A, int row, int me, int world,
                                                       unsigned long long value, unsigned int* ctr) {
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(ctr, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if ((int)threadIdx.x < world) {
            volatile unsigned long long* f = flag_ptr(P.p[threadIdx.x].arena, A, row, me);
            *f = value;
        }
        if (threadIdx.x == 0) *ctr = 0;
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// upload: per owner, the list of {shard-local row, my slot} -> the owner's key inbox
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
                opos[i] = j;
            } else {
                *overflow = 1;
                opos[i] = 0xffffffffu;
            }
        }
        __syncthreads();
    }
    __threadfence_system();
}
// counts into the owners' headers, counters re-armed, then the generation flag of (slot, me) on every owner
__global__ void send_keys_finish_kernel(PeerTable P, ArenaLayout A, int me, int world, int slot, int flag_slot, unsigned cap_pair,
                                        unsigned int* send_cnt, unsigned long long gen) {
    const int o = threadIdx.x;
    __threadfence_system();
    if (o < world) {
        unsigned int* hdr = reinterpret_cast<unsigned int*>(P.p[o].arena + A.key_inbox + ((size_t)slot * world + me) * A.key_region);
        hdr[0] = min(send_cnt[o], cap_pair);
        send_cnt[o] = 0;
        __threadfence_system();
        volatile unsigned long long* f = flag_ptr(P.p[o].arena, A, FLAG_KEYS + flag_slot, me);
        *f = gen;
    }
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------------------------
// step 1: owner-driven pull
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxSl = 4;
__global__ void __launch_bounds__(256)
serve_pull_kernel(PeerTable P, ArenaLayout A, int me, int world, int slot, int flag_slot, unsigned long long gen,
                  unsigned long long epoch, int rowlen, const float* __restrict__ W, const float* __restrict__ V,
                  unsigned int* done_ctr) {
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
        const unsigned char* region = P.p[me].arena + A.key_inbox + ((size_t)slot * world + r) * A.key_region;
        const unsigned n = *reinterpret_cast<const volatile unsigned int*>(region);
        const uint2* pairs = reinterpret_cast<const uint2*>(region + 64);
        float* cW = reinterpret_cast<float*>(P.p[r].arena + A.cacheW);
        float* cV = reinterpret_cast<float*>(P.p[r].arena + A.cacheV);
        if (vec == 4 && slices <= lpr * kMaxSl) {
            constexpr int U = 4;  // row groups in flight before the first (posted) store leaves
            for (unsigned b0 = warp * (G * U); b0 < n; b0 += nwarps * (G * U)) {
                uint2 pr[U];
                float4 v[U][kMaxSl];
                float w[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const unsigned j = b0 + u * G + g;
                    pr[u] = j < n ? pairs[j] : make_uint2(0xffffffffu, 0);
                    if (pr[u].x == 0xffffffffu) continue;
                    const float* src = V + (size_t)pr[u].x * rowlen;
#pragma unroll
                    for (int i = 0; i < kMaxSl; i++) {
                        const int sl = q + i * lpr;
                        if (sl < slices) v[u][i] = *reinterpret_cast<const float4*>(src + 4 * sl);
                    }
                    if (q == 0) w[u] = W[pr[u].x];
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (pr[u].x == 0xffffffffu) continue;
                    float* dst = cV + (size_t)pr[u].y * rowlen;
#pragma unroll
                    for (int i = 0; i < kMaxSl; i++) {
                        const int sl = q + i * lpr;
                        if (sl < slices) *reinterpret_cast<float4*>(dst + 4 * sl) = v[u][i];
                    }
                    if (q == 0) cW[pr[u].y] = w[u];
                }
            }
        } else {
            for (unsigned j = warp * G + g; j < n; j += nwarps * G) {  // generic fallback (odd row lengths)
                const uint2 pr = pairs[j];
                const float* src = V + (size_t)pr.x * rowlen;
                float* dst = cV + (size_t)pr.y * rowlen;
                for (int sl = q * vec; sl < rowlen; sl += lpr * vec)
                    for (int c = 0; c < vec; c++) dst[sl + c] = src[sl + c];
                if (q == 0) cW[pr.y] = W[pr.x];
            }
        }
    }
    raise_flags_last_block(P, A, FLAG_PULLED, me, world, epoch, done_ctr);
}

// stand-alone wait for compute kernels without an in-kernel wait (FFM / NFM / non-fused FM)
__global__ void wait_flags_kernel(PeerTable P, ArenaLayout A, int me, int row, int world, unsigned long long value) {
    wait_flags(P.p[me].arena, A, row, world, value);
}

// ---------------------------------------------------------------------------------------------------------------
// step 3: push
// ---------------------------------------------------------------------------------------------------------------
// FM fused path: row i of G (stride GS, [gV (K) | gW]) + the replica rows of hot slots -> the owner's update_g (peer REDs).
// Blocks < main_blocks walk the ordinary slots, LPR lanes per row; the others fold one hot slot per warp (lane = column).
template <int K>
__global__ void __launch_bounds__(256)
push_fused_kernel(const uint32_t* __restrict__ uniq, const unsigned int* __restrict__ n_uniq, float* __restrict__ G,
                  const uint32_t* __restrict__ hot_of, const uint32_t* __restrict__ hot_slot,
                  const unsigned int* __restrict__ n_hot, float* __restrict__ Ghot, int GS, int main_blocks, PeerTable P,
                  ArenaLayout A, int me, int world, int shift, unsigned long long epoch, unsigned int* done_ctr) {
    constexpr int LPR = K / 4, GR = 32 / LPR, U = 4;
    const unsigned mask = (unsigned)world - 1;
    const int lane = threadIdx.x & 31;
    if ((int)blockIdx.x >= main_blocks) {
        const unsigned hwarp = (blockIdx.x - main_blocks) * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const unsigned nhw = (gridDim.x - main_blocks) * (blockDim.x >> 5);
        const unsigned nh = min(*n_hot, (unsigned)kHotMaxD);
        for (unsigned h = hwarp; h < nh; h += nhw) {
            const uint32_t f = __ldg(uniq + __ldg(hot_slot + h));
            const unsigned o = f & mask;
            const size_t l = f >> shift;
            float* tile = Ghot + (size_t)h * kHotRepD * GS;
            for (int c0 = 0; c0 < K + 1; c0 += 32) {
                const int ncol = GS < 32 ? GS : 32;
                const int col = c0 + lane % ncol;
                const int grp = lane / ncol, ngrp = 32 / ncol;
                float sum = 0.f;
#pragma unroll 8
                for (int i = 0; i < kHotRepD; i++) {
                    if (i < kHotRepD / ngrp) {
                        float* p = tile + (size_t)(grp + i * ngrp) * GS + col;
                        sum += __ldcg(p);
                        *p = 0.f;
                    }
                }
                for (int o2 = ncol; o2 < 32; o2 <<= 1) sum += __shfl_xor_sync(kFull, sum, o2);
                if (grp == 0 && col <= K && sum != 0.f) {
                    red_add_f32(col < K ? P.p[o].gV + l * K + col : P.p[o].gW + l, sum);
                    P.p[o].touched[l] = 1;  // (idempotent byte store; every lane that sent something marks the row)
                }
            }
        }
    } else {
        const int q = lane % LPR, g = lane / LPR;
        const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const unsigned nwarps = (unsigned)main_blocks * (blockDim.x >> 5);
        const unsigned total = *n_uniq;
        for (unsigned b0 = warp * (GR * U); b0 < total; b0 += nwarps * (GR * U)) {
            uint32_t f[U];
            float4 v[U];
            float w[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const unsigned idx = b0 + u * GR + g;
                ok[u] = idx < total && __ldg(hot_of + idx) == 0xffffffffu;
                f[u] = ok[u] ? __ldg(uniq + idx) : 0u;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                w[u] = 0.f;
                if (ok[u]) {
                    v[u] = *reinterpret_cast<const float4*>(G + (size_t)idx * GS + 4 * q);
                    if (q == 0) w[u] = G[(size_t)idx * GS + K];
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (!ok[u]) continue;
                const unsigned idx = b0 + u * GR + g;
                const unsigned o = f[u] & mask;
                const size_t l = f[u] >> shift;
                if (v[u].x != 0.f || v[u].y != 0.f || v[u].z != 0.f || v[u].w != 0.f) {
                    red_add_v4(P.p[o].gV + l * K + 4 * q, v[u]);
                    *reinterpret_cast<float4*>(G + (size_t)idx * GS + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
                    P.p[o].touched[l] = 1;  // (idempotent byte store; every lane that sent something marks the row)
                }
                if (q == 0 && w[u] != 0.f) {
                    red_add_f32(P.p[o].gW + l, w[u]);
                    G[(size_t)idx * GS + K] = 0.f;
                    P.p[o].touched[l] = 1;
                }
            }
        }
    }
    raise_flags_last_block(P, A, FLAG_PUSHED, me, world, epoch, done_ctr);
}

// generic path (FFM / NFM / non-fused FM): compact rows cgV[slot][rowlen], cgW[slot].  MAILBOX: posted stores into the
// owner's gradient inbox at the position of the key in the list the owner received (opos); otherwise peer REDs into the
// owner's update_g.  One warp per row, 16 B slices, all loads of a row before its stores; the local rows are re-zeroed.
template <bool MAILBOX>
__global__ void __launch_bounds__(256)
push_rows_kernel(const uint32_t* __restrict__ uniq, const unsigned int* __restrict__ n_uniq, const uint32_t* __restrict__ opos,
                 float* __restrict__ cgW, float* __restrict__ cgV, int rowlen, int recw, PeerTable P, ArenaLayout A, int me,
                 int world, int shift, unsigned long long epoch, unsigned int* done_ctr) {
    const unsigned n = *n_uniq;
    const unsigned mask = (unsigned)world - 1;
    const int lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const bool vec4 = rowlen % 4 == 0 && rowlen <= 128 * kMaxSl;
    const int slices = rowlen / 4;
    for (unsigned i = warp; i < n; i += nwarps) {
        const uint32_t f = uniq[i];
        const unsigned o = f & mask;
        const size_t l = f >> shift;
        float* gsrc = cgV + (size_t)i * rowlen;
        const float gw = cgW[i];
        float* dstrow;
        if (MAILBOX) {
            const uint32_t j = opos[i];
            if (j == 0xffffffffu) continue;  // list overflow: reported through the overflow flag at upload
            dstrow = reinterpret_cast<float*>(P.p[o].arena + A.grad_inbox + (size_t)me * A.grad_region) + (size_t)j * recw;
        } else {
            dstrow = P.p[o].gV + l * rowlen;
        }
        if (vec4) {
            float4 v[kMaxSl];
#pragma unroll
            for (int s = 0; s < kMaxSl; s++) {
                const int sl = lane + 32 * s;
                if (sl < slices) v[s] = *reinterpret_cast<const float4*>(gsrc + 4 * sl);
            }
#pragma unroll
            for (int s = 0; s < kMaxSl; s++) {
                const int sl = lane + 32 * s;
                if (sl < slices) {
                    if (MAILBOX) *reinterpret_cast<float4*>(dstrow + 4 * sl) = v[s];
                    else if (v[s].x != 0.f || v[s].y != 0.f || v[s].z != 0.f || v[s].w != 0.f) red_add_v4(dstrow + 4 * sl, v[s]);
                    *reinterpret_cast<float4*>(gsrc + 4 * sl) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        } else {
            for (int c = lane; c < rowlen; c += 32) {
                const float v = gsrc[c];
                if (MAILBOX) dstrow[c] = v;
                else if (v != 0.f) red_add_f32(dstrow + c, v);
                gsrc[c] = 0.f;
            }
        }
        if (lane == 0) {
            if (MAILBOX) dstrow[rowlen] = gw;
            else {
                red_add_f32(P.p[o].gW + l, gw);
                P.p[o].touched[l] = 1;
            }
            cgW[i] = 0.f;
        }
    }
    raise_flags_last_block(P, A, FLAG_PUSHED, me, world, epoch, done_ctr);
}

// ---------------------------------------------------------------------------------------------------------------
// step 4 (wide rows): owner folds the R gradient inboxes into its shard's update_g (local REDs) and marks the rows
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
merge_kernel(PeerTable P, ArenaLayout A, int me, int world, int slot, unsigned long long epoch, int rowlen, int recw,
             float* __restrict__ gW, float* __restrict__ gV, uint8_t* __restrict__ touched) {
    wait_flags(P.p[me].arena, A, FLAG_PUSHED, world, epoch);  // every requester's records have landed
    const int lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const bool vec4 = rowlen % 4 == 0 && recw % 4 == 0 && rowlen <= 128 * kMaxSl;
    const int slices = rowlen / 4;
    for (int src = 0; src < world; src++) {
        const unsigned char* region = P.p[me].arena + A.key_inbox + ((size_t)slot * world + src) * A.key_region;
        const unsigned n = *reinterpret_cast<const volatile unsigned int*>(region);
        const uint2* pairs = reinterpret_cast<const uint2*>(region + 64);
        const float* recs = reinterpret_cast<const float*>(P.p[me].arena + A.grad_inbox + (size_t)src * A.grad_region);
        for (unsigned j = warp; j < n; j += nwarps) {
            const float* rec = recs + (size_t)j * recw;
            const size_t l = pairs[j].x;
            float* gdst = gV + l * (size_t)rowlen;
            if (vec4) {
                float4 v[kMaxSl];
#pragma unroll
                for (int s = 0; s < kMaxSl; s++) {
                    const int sl = lane + 32 * s;
                    if (sl < slices) v[s] = __ldcg(reinterpret_cast<const float4*>(rec + 4 * sl));
                }
#pragma unroll
                for (int s = 0; s < kMaxSl; s++) {
                    const int sl = lane + 32 * s;
                    if (sl < slices && (v[s].x != 0.f || v[s].y != 0.f || v[s].z != 0.f || v[s].w != 0.f)) red_add_v4(gdst + 4 * sl, v[s]);
                }
            } else {
                for (int c = lane; c < rowlen; c += 32) red_add_f32(gdst + c, __ldcg(rec + c));
            }
            if (lane == 0) {
                red_add_f32(gW + l, __ldcg(rec + rowlen));
                touched[l] = 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int dist_alloc(lctr_ctx* c) {
    const int R = c->cfg.world;
    LCTR_CHECK(R <= kMaxWorld && (R & (R - 1)) == 0, "world=%d: need a power of two <= %d", R, kMaxWorld);
    DistState* d = new DistState();
    c->dist = d;
    d->rank = c->cfg.rank; d->world = R;
    while ((1 << d->shift) < R) d->shift++;
    // keys of one batch: at most its entry count (cfg.max_nnz), at most the id space
    d->cap_keys = c->cfg.max_nnz ? std::min<size_t>(c->cfg.max_nnz, c->F) : c->F;
    // keys one requester sends one owner: U / R on average (owner = fid mod R); twice that plus slack, at most the shard
    d->cap_pair = std::min<size_t>(c->Fl, 2 * d->cap_keys / R + 4096);
    d->recw = (int)((c->rowlen + 4 + 3) / 4 * 4);
    const char* pm0 = getenv("LCTR_DIST_PUSH");
    d->mailbox = pm0 ? strcmp(pm0, "mailbox") == 0 : c->rowlen >= 64;
    ArenaLayout& A = d->A;
    size_t off = 0;
    A.flags = off; off = align_up(off + (size_t)(3 + kNumSlots) * kMaxWorld * sizeof(unsigned long long), 256);
    A.key_region = align_up(64 + d->cap_pair * sizeof(uint2), 256);
    A.key_inbox = off; off += A.key_region * kNumSlots * 2 * R;
    A.cacheW = off; off = align_up(off + d->cap_keys * sizeof(float), 256);
    A.cacheV = off; off = align_up(off + d->cap_keys * c->rowlen * sizeof(float), 256);
    A.grad_region = d->mailbox ? align_up(d->cap_pair * (size_t)d->recw * sizeof(float), 256) : 0;
    A.grad_inbox = off; off += A.grad_region * R;
    A.total = off;
    LCTR_CUDA(cudaMalloc((void**)&d->arena, A.total));
    LCTR_CUDA(cudaMemsetAsync(d->arena, 0, A.total, c->stream));
    d->bytes += A.total;
    LCTR_CUDA(cudaMalloc((void**)&d->send_cnt, kMaxWorld * sizeof(unsigned int)));
    LCTR_CUDA(cudaMemsetAsync(d->send_cnt, 0, kMaxWorld * sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMalloc((void**)&d->opos, (size_t)kNumSlots * d->cap_keys * sizeof(uint32_t)));
    d->bytes += (size_t)kNumSlots * d->cap_keys * sizeof(uint32_t);
    LCTR_CUDA(cudaMalloc((void**)&d->done_ctr, 4 * sizeof(unsigned int)));
    LCTR_CUDA(cudaMemsetAsync(d->done_ctr, 0, 4 * sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMalloc((void**)&d->overflow, sizeof(int)));
    LCTR_CUDA(cudaMemsetAsync(d->overflow, 0, sizeof(int), c->stream));
    if (!fused_kernels_ok(c)) {  // the fused FM kernels keep their gradients in fm_fused.cu's G / Ghot
        LCTR_CUDA(cudaMalloc((void**)&d->cgW, d->cap_keys * sizeof(float)));
        LCTR_CUDA(cudaMalloc((void**)&d->cgV, d->cap_keys * c->rowlen * sizeof(float)));
        LCTR_CUDA(cudaMemsetAsync(d->cgW, 0, d->cap_keys * sizeof(float), c->stream));
        LCTR_CUDA(cudaMemsetAsync(d->cgV, 0, d->cap_keys * c->rowlen * sizeof(float), c->stream));
        d->bytes += d->cap_keys * (c->rowlen + 1) * sizeof(float);
    }
    // compute view: the batch-compact cache and gradient rows (indexed by slot)
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
    cudaFree(d->arena); cudaFree(d->send_cnt); cudaFree(d->opos); cudaFree(d->done_ctr); cudaFree(d->overflow);
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
    send_keys_kernel<<<grid, 256, 0, st>>>(s.uniq, s.n_uniq, d->peers, d->A, d->rank, d->world, slot2, d->shift, (unsigned)d->cap_pair,
                                           d->send_cnt, d->opos + (size_t)slot * d->cap_keys, d->overflow);
    send_keys_finish_kernel<<<1, 32, 0, st>>>(d->peers, d->A, d->rank, d->world, slot2, slot, (unsigned)d->cap_pair, d->send_cnt, d->gen[slot]);
    c->launches += 2;
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
    { ProfScope prof(c, PROF_DIST_PULL);
    serve_pull_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d->peers, d->A, d->rank, d->world, slot * 2 + (int)(d->gen[slot] & 1), slot,
                                                             d->gen[slot], d->epoch, (int)c->rowlen, c->W, c->V, d->done_ctr + 0); }
    c->launches++;
    if (!in_kernel_wait) {
        ProfScope prof(c, PROF_DIST_BAR1);
        wait_flags_kernel<<<1, 32, 0, c->stream>>>(d->peers, d->A, d->rank, FLAG_PULLED, d->world, d->epoch);
        c->launches++;
    }
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

template <int K>
static void push_fused_go(lctr_ctx* c, Slot& s) {
    DistState* d = c->dist;
    FusedState* f = c->fused;
    const int main_blocks = c->sm_count * 2;
    push_fused_kernel<K><<<main_blocks + kHotMaxD / 8, 256, 0, c->stream>>>(s.uniq, s.n_uniq, f->G, s.hot_of, s.hot_slot, s.n_hot, f->Ghot,
                                                                          f->GS, main_blocks, d->peers, d->A, d->rank, d->world,
                                                                          d->shift, d->epoch, d->done_ctr + 1);
}

int dist_post_step(lctr_ctx* c, Slot& s, int slot, int64_t rows_divisor) {
    DistState* d = c->dist;
    if (fused_kernels_ok(c)) {
        ProfScope prof(c, PROF_DIST_PUSH);
        switch ((int)c->cfg.factor_cnt) {
            case 4: push_fused_go<4>(c, s); break;
            case 8: push_fused_go<8>(c, s); break;
            case 16: push_fused_go<16>(c, s); break;
            default: push_fused_go<32>(c, s); break;
        }
        c->launches++;
    } else {
        { ProfScope prof(c, PROF_DIST_PUSH);
        const uint32_t* opos = d->opos + (size_t)slot * d->cap_keys;
        if (d->mailbox)
            push_rows_kernel<true><<<c->sm_count * 4, 256, 0, c->stream>>>(s.uniq, s.n_uniq, opos, d->cgW, d->cgV, (int)c->rowlen, d->recw,
                                                                           d->peers, d->A, d->rank, d->world, d->shift, d->epoch, d->done_ctr + 1);
        else
            push_rows_kernel<false><<<c->sm_count * 4, 256, 0, c->stream>>>(s.uniq, s.n_uniq, opos, d->cgW, d->cgV, (int)c->rowlen, d->recw,
                                                                            d->peers, d->A, d->rank, d->world, d->shift, d->epoch, d->done_ctr + 1); }
        c->launches++;
        if (d->mailbox) {
            ProfScope prof(c, PROF_DIST_MERGE);
            merge_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d->peers, d->A, d->rank, d->world, slot * 2 + (int)(d->gen[slot] & 1), d->epoch, (int)c->rowlen, d->recw,
                                                                 c->gW, c->gV, c->touched);
            c->launches++;
        }
    }
    LCTR_CUDA(cudaGetLastError());
    // owner: sparse updater on the shard; its first kernel waits for "pushes landed" unless the merge already did
    const unsigned long long* pushed = reinterpret_cast<const unsigned long long*>(d->arena + d->A.flags) + (size_t)FLAG_PUSHED * kMaxWorld;
    c->apply_wait_flags = (fused_kernels_ok(c) || !d->mailbox) ? pushed : nullptr;
    c->apply_wait_n = d->world;
    c->apply_wait_epoch = d->epoch;
    const int rc = launch_apply(c, rows_divisor);
    c->apply_wait_flags = nullptr;
    return rc;
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
