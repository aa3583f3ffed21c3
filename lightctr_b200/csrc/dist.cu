// lightctr_b200/csrc/dist.cu -- multi-GPU sparse exchange over NVLink peer memory (one process per GPU).
//
// Replaces the reference's Parameter-Server round trips: Pull::sync of the batch's unique keys
// (distribut/pull.h:43-68; unique-key build distributed_algo_abst.h:181-195) and Push::sync of the per-key
// gradients (distribut/push.h:36-51) with the owner applying the update (distribut/paramserver.h:181-310).
// Differences we state rather than reproduce: synchronous instead of SSP-async, fp32 instead of
// fp16 on the wire, no |g| thresholding of pushes, owner = fid mod R instead of the murmur DHT ring.
//
// Tables are owner-sharded: row f lives on rank f % R at shard-local index f / R.  Every rank maps every peer's
// W/V shard, mailbox and barrier words through CUDA IPC.  One training step on rank r:
//   1 mark      byte-mark the fids of MY batch in a local map              (mark_kernel)
//   2 compact   -> list of my unique fids (the "pull_map" keys)            (compact_touched_kernel)
//   3 pull      copy each unique row W[f], V[f,:] from its owner's shard (peer loads, one read per unique id per
//               rank) into my full-size local cache                        (pull_kernel)
//   4 fwd/bwd   the single-GPU kernels on the cache + local update_g        (fm.cu / ffm.cu, unchanged)
//   5 push      one record {fid, gW, gV[rowlen]} per unique id into the OWNER's mailbox[src = r] with plain peer
//               stores; zero my local update_g row                         (push_kernel, push_counts_kernel)
//   6 barrier   all pushes landed                                          (rank_barrier_kernel: release/acquire
//                                                                           flags in peer memory, no host round trip)
//   7 merge     owner adds the R mailboxes into its shard's update_g with local REDs and marks touched
//   8 apply     the single-GPU sparse updater on the shard               (opt.cu, unchanged)
//   9 barrier   updated rows visible before anybody's next pull
// NVLink carries each unique row once per rank and direction (vs. once per occurrence for naive peer gathers).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "opt.cuh"

namespace lctr {

constexpr int kMaxWorld = 8;

struct PeerPtrs {
    float* W[kMaxWorld];
    float* V[kMaxWorld];
    unsigned char* mail[kMaxWorld];
    unsigned long long* bar[kMaxWorld];
};
struct PeerPtrs2 {  // second set (update_g shards + touched maps), kept in device memory
    float* gW[kMaxWorld];
    float* gV[kMaxWorld];
    uint8_t* touched[kMaxWorld];
    unsigned long long* bar[kMaxWorld];
};

struct DistState {
    int rank = 0, world = 1, shift = 0;
    uint8_t* mark = nullptr;        // F bytes (global fid), per-step fallback on the compute stream
    uint8_t* mark_up = nullptr;     // F bytes, used by uploads (their own stream)
    uint32_t* uniq = nullptr;       // unique fids of my batch
    unsigned int* n_uniq = nullptr;
    unsigned int* push_cnt = nullptr;  // [world] records written per destination this step
    unsigned int* scratch_done = nullptr;
    // exported buffers (owned by this rank)
    unsigned char* mailbox = nullptr;  // world regions: [src][header 64 B | cap records]
    unsigned long long* bar = nullptr; // [2][world] epoch words written by peers
    size_t rec_floats = 0, rec_cap = 0, region_bytes = 0;
    PeerPtrs peers;                 // device pointers valid in THIS process
    PeerPtrs2 peers2;
    PeerPtrs2* d_peers2 = nullptr;  // device copy
    void* opened[kMaxWorld][7] = {{nullptr}};
    bool imported = false;
    bool use_mailbox = false;       // LCTR_DIST_PUSH=mailbox selects the record/merge variant
    bool bar1_pending = false;      // barrier 1 has been signalled but not yet waited for
    const uint32_t* cur_uniq = nullptr;       // key set of the step in flight (slot-owned or the per-step list)
    unsigned int* cur_n = nullptr;
    bool cur_dynamic = false;
    unsigned long long epoch = 0;
};

// ---------------------------------------------------------------------------------------------------------------
__global__ void mark_kernel(const uint32_t* __restrict__ fid, int64_t b, int64_t e, uint8_t* __restrict__ mark) {
    for (int64_t i = b + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x)
        mark[fid[i]] = 1;
}

// each unique row is read ONCE from its owner (peer or local shard) into the local cache.  LPR lanes cover a row with
// 16 B loads (slices q, q+LPR, ...: up to kMaxSl per lane), G = 32/LPR rows per warp step, and kPullU steps are kept in
// flight before the first store so that one NVLink round trip covers G*kPullU rows.
constexpr int kMaxSl = 4;
constexpr int kPullU = 4;
__global__ void __launch_bounds__(256)
pull_kernel(const uint32_t* __restrict__ uniq, const unsigned int* __restrict__ n_uniq, PeerPtrs P, int shift,
            unsigned mask, int rowlen, float* __restrict__ cW, float* __restrict__ cV) {
    const unsigned n = *n_uniq;
    const int lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const int vec = (rowlen % 4 == 0) ? 4 : 1;
    const int slices = rowlen / vec;
    int lpr = 1;
    while (lpr < slices && lpr < 32) lpr <<= 1;
    const int G = 32 / lpr, q = lane % lpr, g = lane / lpr;
    if (vec == 4 && slices <= lpr * kMaxSl) {
        for (unsigned b0 = warp * (G * kPullU); b0 < n; b0 += nwarps * (G * kPullU)) {
            uint32_t f[kPullU];
            float4 v[kPullU][kMaxSl];
            float w[kPullU];
#pragma unroll
            for (int u = 0; u < kPullU; u++) {
                const unsigned idx = b0 + u * G + g;
                f[u] = idx < n ? uniq[idx] : 0xffffffffu;
                if (f[u] == 0xffffffffu) continue;
                const unsigned o = f[u] & mask;
                const size_t l = f[u] >> shift;
                const float* src = P.V[o] + l * (size_t)rowlen;
#pragma unroll
                for (int i = 0; i < kMaxSl; i++) {
                    const int sl = q + i * lpr;
                    if (sl < slices) v[u][i] = *reinterpret_cast<const float4*>(src + 4 * sl);
                }
                if (q == 0) w[u] = P.W[o][l];
            }
#pragma unroll
            for (int u = 0; u < kPullU; u++) {
                if (f[u] == 0xffffffffu) continue;
                float* dst = cV + (size_t)f[u] * rowlen;
#pragma unroll
                for (int i = 0; i < kMaxSl; i++) {
                    const int sl = q + i * lpr;
                    if (sl < slices) *reinterpret_cast<float4*>(dst + 4 * sl) = v[u][i];
                }
                if (q == 0) cW[f[u]] = w[u];
            }
        }
        return;
    }
    for (unsigned b0 = warp * G; b0 < n; b0 += nwarps * G) {  // generic fallback (odd row lengths)
        const unsigned idx = b0 + g;
        if (idx >= n) continue;
        const uint32_t f = uniq[idx];
        const unsigned o = f & mask;
        const size_t l = f >> shift;
        const float* src = P.V[o] + l * (size_t)rowlen;
        float* dst = cV + (size_t)f * rowlen;
        for (int sl = q * vec; sl < rowlen; sl += lpr * vec)
            for (int c = 0; c < vec; c++) dst[sl + c] = src[sl + c];
        if (q == 0) cW[f] = P.W[o][l];
    }
}

// record layout (floats): [0 .. rowlen) gV (16 B aligned), [rowlen] fid bits, [rowlen+1] gW ; padded to a multiple of 4.
// Slots are reserved per CTA chunk of kPushChunk records: a shared histogram over the <= 8 destinations, ONE global
// atomicAdd per destination per chunk (the per-record atomics of a naive version serialise on 8 counters), ranks by
// a short scan over the chunk; then each warp copies its records with the row stores fully coalesced.
constexpr int kPushChunk = 64;
__global__ void __launch_bounds__(256)
push_kernel(const uint32_t* __restrict__ uniq, const unsigned int* __restrict__ n_uniq, PeerPtrs P, int me, int shift,
            unsigned mask, int rowlen, int rec_floats, size_t region_bytes, unsigned rec_cap,
            float* __restrict__ cgW, float* __restrict__ cgV, unsigned int* __restrict__ push_cnt) {
    __shared__ uint32_t s_f[kPushChunk];
    __shared__ unsigned s_slot[kPushChunk];
    __shared__ unsigned s_hist[kMaxWorld], s_base[kMaxWorld];
    const unsigned n = *n_uniq;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    for (unsigned c0 = blockIdx.x * kPushChunk; c0 < n; c0 += gridDim.x * kPushChunk) {
        if (t < kMaxWorld) s_hist[t] = 0;
        __syncthreads();
        uint32_t f = 0xffffffffu;
        if (t < kPushChunk && c0 + t < n) { f = uniq[c0 + t]; atomicAdd(&s_hist[f & mask], 1u); }
        if (t < kPushChunk) s_f[t] = f;
        __syncthreads();
        if (t < kMaxWorld && s_hist[t]) s_base[t] = atomicAdd(&push_cnt[t], s_hist[t]);
        __syncthreads();
        if (t < kPushChunk && f != 0xffffffffu) {
            const unsigned o = f & mask;
            unsigned rank = 0;
            for (int j = 0; j < t; j++) rank += (s_f[j] != 0xffffffffu && (s_f[j] & mask) == o) ? 1u : 0u;
            s_slot[t] = s_base[o] + rank;
        }
        __syncthreads();
        for (int rI = wid; rI < kPushChunk; rI += (int)(blockDim.x >> 5)) {
            const uint32_t ff = s_f[rI];
            if (ff == 0xffffffffu) continue;
            const unsigned slot = s_slot[rI];
            if (slot >= rec_cap) continue;
            const unsigned o = ff & mask;
            float* rec = reinterpret_cast<float*>(P.mail[o] + (size_t)me * region_bytes + 64) + (size_t)slot * rec_floats;
            float* gsrc = cgV + (size_t)ff * rowlen;
            if (lane == 0) {
                rec[rowlen] = __uint_as_float(ff);
                rec[rowlen + 1] = cgW[ff];
                cgW[ff] = 0.f;
            }
            if (rowlen % 4 == 0 && rowlen <= 128 * kMaxSl) {
                // all 16 B loads of the row first, then the (posted) peer stores and the re-zeroing of update_g
                const int slices = rowlen / 4;
                float4 v[kMaxSl];
#pragma unroll
                for (int i = 0; i < kMaxSl; i++) {
                    const int sl = lane + 32 * i;
                    if (sl < slices) v[i] = *reinterpret_cast<const float4*>(gsrc + 4 * sl);
                }
#pragma unroll
                for (int i = 0; i < kMaxSl; i++) {
                    const int sl = lane + 32 * i;
                    if (sl < slices) {
                        *reinterpret_cast<float4*>(rec + 4 * sl) = v[i];
                        *reinterpret_cast<float4*>(gsrc + 4 * sl) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            } else {
                for (int i = lane; i < rowlen; i += 32) {
                    rec[i] = gsrc[i];
                    gsrc[i] = 0.f;
                }
            }
        }
        __syncthreads();
    }
    __threadfence_system();
}

// push variant without mailboxes: each unique row's gradient is added straight into the OWNER's update_g with vector
// REDs through the peer mapping (NVLink forwards the atomics) and the owner's touched byte is set with a peer store.
// G rows per warp step; the local update_g row is zeroed on the way.  One RED row per unique id per rank.
__global__ void __launch_bounds__(256)
push_red_kernel(const uint32_t* __restrict__ uniq, const unsigned int* __restrict__ n_uniq, float* const* __restrict__ pgW,
                float* const* __restrict__ pgV, uint8_t* const* __restrict__ ptouched, int shift, unsigned mask, int rowlen,
                float* __restrict__ cgW, float* __restrict__ cgV) {
    const unsigned n = *n_uniq;
    const int lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const int vec = (rowlen % 4 == 0) ? 4 : 1;
    const int slices = rowlen / vec;
    int lpr = 1;
    while (lpr < slices && lpr < 32) lpr <<= 1;
    const int G = 32 / lpr, q = lane % lpr, g = lane / lpr;
    if (vec == 4 && slices <= lpr) {
        // narrow rows (FM): one float4 per lane; kPullU row groups are loaded before the first RED leaves
        for (unsigned b0 = warp * (G * kPullU); b0 < n; b0 += nwarps * (G * kPullU)) {
            uint32_t f[kPullU];
            float4 v[kPullU];
            float w[kPullU];
#pragma unroll
            for (int u = 0; u < kPullU; u++) {
                const unsigned idx = b0 + u * G + g;
                f[u] = idx < n ? uniq[idx] : 0xffffffffu;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                w[u] = 0.f;
                if (f[u] == 0xffffffffu) continue;
                if (q < slices) v[u] = *reinterpret_cast<const float4*>(cgV + (size_t)f[u] * rowlen + 4 * q);
                if (q == 0) w[u] = cgW[f[u]];
            }
#pragma unroll
            for (int u = 0; u < kPullU; u++) {
                if (f[u] == 0xffffffffu) continue;
                const unsigned o = f[u] & mask;
                const size_t l = f[u] >> shift;
                if (q < slices) {
                    *reinterpret_cast<float4*>(cgV + (size_t)f[u] * rowlen + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (v[u].x != 0.f || v[u].y != 0.f || v[u].z != 0.f || v[u].w != 0.f)
                        red_add_v4(pgV[o] + l * (size_t)rowlen + 4 * q, v[u]);
                }
                if (q == 0) {
                    red_add_f32(pgW[o] + l, w[u]);
                    cgW[f[u]] = 0.f;
                    ptouched[o][l] = 1;
                }
            }
        }
        __threadfence_system();
        return;
    }
    for (unsigned b0 = warp * G; b0 < n; b0 += nwarps * G) {
        const unsigned idx = b0 + g;
        if (idx >= n) continue;
        const uint32_t f = uniq[idx];
        const unsigned o = f & mask;
        const size_t l = f >> shift;
        float* src = cgV + (size_t)f * rowlen;
        float* dst = pgV[o] + l * (size_t)rowlen;
        if (vec == 4) {
            for (int sl = q; sl < slices; sl += lpr) {
                const float4 v = *reinterpret_cast<const float4*>(src + 4 * sl);
                *reinterpret_cast<float4*>(src + 4 * sl) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add_v4(dst + 4 * sl, v);
            }
        } else {
            for (int sl = q; sl < slices; sl += lpr) {
                const float v = src[sl];
                src[sl] = 0.f;
                if (v != 0.f) red_add_f32(dst + sl, v);
            }
        }
        if (q == 0) {
            red_add_f32(pgW[o] + l, cgW[f]);
            cgW[f] = 0.f;
            ptouched[o][l] = 1;
        }
    }
    __threadfence_system();
}

__global__ void reset_counter_kernel(unsigned int* n) { *n = 0; }

// barrier split in two so that rank-local work can sit between signalling and waiting
__global__ void rank_arrive_kernel(unsigned long long* const* __restrict__ pbar, int me, int world, int which,
                                   unsigned long long epoch) {
    const int d = threadIdx.x;
    __threadfence_system();
    if (d < world) {
        volatile unsigned long long* theirs = pbar[d] + (size_t)which * kMaxWorld + me;
        *theirs = epoch;
    }
    __threadfence_system();
}
__global__ void rank_wait_kernel(unsigned long long* const* __restrict__ pbar, int me, int world, int which,
                                 unsigned long long epoch) {
    const int d = threadIdx.x;
    if (d < world) {
        volatile unsigned long long* mine = pbar[me] + (size_t)which * kMaxWorld + d;
        while (*mine < epoch) { __nanosleep(32); }
    }
    __syncthreads();
    __threadfence_system();
}

// publish my record counts into every owner's mailbox header and re-arm the local counters / unique list
__global__ void push_counts_kernel(PeerPtrs P, int me, int world, size_t region_bytes, unsigned rec_cap,
                                   unsigned int* push_cnt, unsigned int* n_uniq) {
    const int d = threadIdx.x;
    if (d < world) {
        unsigned int* hdr = reinterpret_cast<unsigned int*>(P.mail[d] + (size_t)me * region_bytes);
        hdr[0] = min(push_cnt[d], rec_cap);
        push_cnt[d] = 0;
    }
    if (d == 0) *n_uniq = 0;
    __threadfence_system();
}

// all-ranks barrier through peer memory: write my epoch into slot [me] of every rank's word array, then wait until
// every slot of MY array has reached the epoch.  One tiny kernel per rank; no host involvement.
__global__ void rank_barrier_kernel(PeerPtrs P, int me, int world, int which, unsigned long long epoch) {
    const int d = threadIdx.x;
    __threadfence_system();
    if (d < world) {
        volatile unsigned long long* theirs = P.bar[d] + (size_t)which * kMaxWorld + me;
        *theirs = epoch;
        __threadfence_system();
        volatile unsigned long long* mine = P.bar[me] + (size_t)which * kMaxWorld + d;
        while (*mine < epoch) { __nanosleep(64); }
    }
    __syncthreads();
    __threadfence_system();
}

// owner side: fold the R mailboxes into the shard's update_g (local REDs) and mark the shard-local rows
__global__ void __launch_bounds__(256)
merge_kernel(const unsigned char* __restrict__ mailbox, int world, size_t region_bytes, int rec_floats, int rowlen,
             int shift, float* __restrict__ gW, float* __restrict__ gV, uint8_t* __restrict__ touched) {
    const int lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    for (int src = 0; src < world; src++) {
        const unsigned char* region = mailbox + (size_t)src * region_bytes;
        const unsigned n = *reinterpret_cast<const volatile unsigned int*>(region);
        const float* recs = reinterpret_cast<const float*>(region + 64);
        for (unsigned idx = warp; idx < n; idx += nwarps) {
            const float* rec = recs + (size_t)idx * rec_floats;
            const uint32_t f = __float_as_uint(rec[rowlen]);
            const size_t l = f >> shift;
            float* gdst = gV + l * (size_t)rowlen;
            if (rowlen % 4 == 0 && (rec_floats % 4) == 0 && rowlen <= 128 * kMaxSl) {
                const int slices = rowlen / 4;
                float4 v[kMaxSl];
#pragma unroll
                for (int i = 0; i < kMaxSl; i++) {
                    const int sl = lane + 32 * i;
                    if (sl < slices) v[i] = *reinterpret_cast<const float4*>(rec + 4 * sl);
                }
#pragma unroll
                for (int i = 0; i < kMaxSl; i++) {
                    const int sl = lane + 32 * i;
                    if (sl < slices) red_add_v4(gdst + 4 * sl, v[i]);
                }
            } else {
                for (int i = lane; i < rowlen; i += 32) red_add_f32(gdst + i, rec[i]);
            }
            if (lane == 0) {
                red_add_f32(gW + l, rec[rowlen + 1]);
                touched[l] = 1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
int dist_alloc(lctr_ctx* c) {
    const int R = c->cfg.world;
    LCTR_CHECK(R <= kMaxWorld && (R & (R - 1)) == 0, "world=%d: need a power of two <= %d", R, kMaxWorld);
    DistState* d = new DistState();
    c->dist = d;
    d->rank = c->cfg.rank; d->world = R;
    while ((1 << d->shift) < R) d->shift++;
    const size_t nv = c->F * c->rowlen;
    LCTR_CUDA(cudaMalloc((void**)&c->cW, c->F * sizeof(float)));
    LCTR_CUDA(cudaMalloc((void**)&c->cV, nv * sizeof(float)));
    LCTR_CUDA(cudaMalloc((void**)&c->cgW, c->F * sizeof(float)));
    LCTR_CUDA(cudaMalloc((void**)&c->cgV, nv * sizeof(float)));
    LCTR_CUDA(cudaMemsetAsync(c->cW, 0, c->F * sizeof(float), c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->cV, 0, nv * sizeof(float), c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->cgW, 0, c->F * sizeof(float), c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->cgV, 0, nv * sizeof(float), c->stream));
    LCTR_CUDA(cudaMalloc((void**)&d->mark, c->F + 512));
    LCTR_CUDA(cudaMemsetAsync(d->mark, 0, c->F + 512, c->stream));
    LCTR_CUDA(cudaMalloc((void**)&d->mark_up, c->F + 512));
    LCTR_CUDA(cudaMemsetAsync(d->mark_up, 0, c->F + 512, c->stream));
    const size_t cap = c->cfg.max_nnz ? std::min<size_t>(c->cfg.max_nnz, c->F) : c->F;
    LCTR_CUDA(cudaMalloc((void**)&d->uniq, (c->F + 32) * sizeof(uint32_t)));
    LCTR_CUDA(cudaMalloc((void**)&d->n_uniq, sizeof(unsigned int)));
    LCTR_CUDA(cudaMalloc((void**)&d->push_cnt, kMaxWorld * sizeof(unsigned int)));
    LCTR_CUDA(cudaMalloc((void**)&d->scratch_done, sizeof(unsigned int)));
    LCTR_CUDA(cudaMemsetAsync(d->n_uniq, 0, sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMemsetAsync(d->push_cnt, 0, kMaxWorld * sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMemsetAsync(d->scratch_done, 0, sizeof(unsigned int), c->stream));
    // push variant: narrow rows (FM) go straight into the owner's update_g with peer REDs; wide rows (FFM) are cheaper
    // as plain peer stores into a mailbox + owner-side merge (measured on C5: 12.8 vs 14.2 ms per step)
    const char* pm0 = getenv("LCTR_DIST_PUSH");
    d->use_mailbox = pm0 ? strcmp(pm0, "mailbox") == 0 : c->rowlen >= 64;
    d->rec_floats = ((2 + c->rowlen) + 3) / 4 * 4;
    // records one source may send to one owner per step: at most the owner's shard size (distinct fids)
    d->rec_cap = d->use_mailbox ? std::min<size_t>(cap, c->Fl) : 0;
    d->region_bytes = (64 + d->rec_cap * d->rec_floats * sizeof(float) + 255) / 256 * 256;
    LCTR_CUDA(cudaMalloc((void**)&d->mailbox, d->region_bytes * R));
    LCTR_CUDA(cudaMemsetAsync(d->mailbox, 0, d->region_bytes * R, c->stream));
    LCTR_CUDA(cudaMalloc((void**)&d->bar, 2 * kMaxWorld * sizeof(unsigned long long)));
    LCTR_CUDA(cudaMemsetAsync(d->bar, 0, 2 * kMaxWorld * sizeof(unsigned long long), c->stream));
    memset(&d->peers, 0, sizeof(d->peers));
    d->peers.W[d->rank] = c->W; d->peers.V[d->rank] = c->V;
    d->peers.mail[d->rank] = d->mailbox; d->peers.bar[d->rank] = d->bar;
    memset(&d->peers2, 0, sizeof(d->peers2));
    d->peers2.gW[d->rank] = c->gW; d->peers2.gV[d->rank] = c->gV; d->peers2.touched[d->rank] = c->touched;
    d->peers2.bar[d->rank] = d->bar;
    LCTR_CUDA(cudaMalloc((void**)&d->d_peers2, sizeof(PeerPtrs2)));
    return 0;
}

int dist_free(lctr_ctx* c) {
    DistState* d = c->dist;
    if (!d) return 0;
    for (int r = 0; r < d->world; r++)
        for (int j = 0; j < 7; j++)
            if (d->opened[r][j]) cudaIpcCloseMemHandle(d->opened[r][j]);
    if (d->d_peers2) cudaFree(d->d_peers2);
    if (c->cW) cudaFree(c->cW); if (c->cV) cudaFree(c->cV); if (c->cgW) cudaFree(c->cgW); if (c->cgV) cudaFree(c->cgV);
    c->cW = c->cV = c->cgW = c->cgV = nullptr;
    cudaFree(d->mark); cudaFree(d->mark_up); cudaFree(d->uniq); cudaFree(d->n_uniq); cudaFree(d->push_cnt); cudaFree(d->scratch_done);
    cudaFree(d->mailbox); cudaFree(d->bar);
    delete d;
    c->dist = nullptr;
    return 0;
}

static int barrier(lctr_ctx* c, int which) {
    DistState* d = c->dist;
    rank_barrier_kernel<<<1, 32, 0, c->stream>>>(d->peers, d->rank, d->world, which, d->epoch);
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

// key set of a whole slot (its unique fids), computed once per upload on the upload stream: mark + compact depend
// only on the batch, not on the parameters, so for streamed batches they overlap the previous step's kernels
int dist_build_uniq(lctr_ctx* c, Slot& s, cudaStream_t st) {
    DistState* d = c->dist;
    s.uniq_valid = false;
    if (!d || s.nnz == 0) return 0;
    const int64_t need = std::min<int64_t>(s.nnz, (int64_t)c->F);
    if (need > s.cap_uniq) {
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        if (s.uniq) cudaFree(s.uniq);
        if (!s.n_uniq) LCTR_CUDA(cudaMalloc((void**)&s.n_uniq, sizeof(unsigned int)));
        const int64_t cap = std::max<int64_t>(need, s.cap_uniq + s.cap_uniq / 2);
        LCTR_CUDA(cudaMalloc((void**)&s.uniq, (size_t)(cap + 32) * sizeof(uint32_t)));
        s.cap_uniq = cap;
    }
    // a private mark map per upload stream would be needed if two uploads ran concurrently; uploads are serialised
    // on their stream, and the per-step fallback below uses the same map only on the compute stream when no slot
    // list exists, so one map suffices
    LCTR_CUDA(cudaMemsetAsync(s.n_uniq, 0, sizeof(unsigned int), st));
    const unsigned mg = (unsigned)std::min<int64_t>((s.nnz + 255) / 256, (int64_t)c->sm_count * 8);
    mark_kernel<<<std::max(mg, 1u), 256, 0, st>>>(s.fid, 0, s.nnz, d->mark_up);
    const size_t ntiles = (c->F + 511) / 512;
    const unsigned cg = (unsigned)std::min<size_t>((ntiles + 7) / 8, (size_t)c->sm_count * 8);
    compact_touched_kernel<<<std::max(cg, 1u), 256, 0, st>>>(d->mark_up, c->F, s.uniq, s.n_uniq);
    c->launches += 2;
    LCTR_CUDA(cudaGetLastError());
    s.uniq_valid = true;
    return 0;
}

int dist_pre_step(lctr_ctx* c, Slot& s, int64_t rb, int64_t re) {
    DistState* d = c->dist;
    LCTR_CHECK(d->imported, "multi-GPU step before lctr_ipc_import");
    if (re - rb <= 0) return 0;
    const unsigned mask = (unsigned)d->world - 1;
    if (s.uniq_valid && rb == 0 && re == s.rows) {
        d->cur_uniq = s.uniq; d->cur_n = s.n_uniq; d->cur_dynamic = false;
    } else {
        // a sub-range of the slot: build the key set of just these rows on the compute stream
        int64_t rp[2];
        LCTR_CUDA(cudaMemcpyAsync(&rp[0], s.row_ptr + rb, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        LCTR_CUDA(cudaMemcpyAsync(&rp[1], s.row_ptr + re, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        const unsigned mg = (unsigned)std::min<int64_t>((rp[1] - rp[0] + 255) / 256, (int64_t)c->sm_count * 8);
        { ProfScope prof(c, PROF_DIST_MARK);
        mark_kernel<<<std::max(mg, 1u), 256, 0, c->stream>>>(s.fid, rp[0], rp[1], d->mark); }
        const size_t ntiles = (c->F + 511) / 512;
        const unsigned cg = (unsigned)std::min<size_t>((ntiles + 7) / 8, (size_t)c->sm_count * 8);
        { ProfScope prof(c, PROF_DIST_COMPACT);
        compact_touched_kernel<<<std::max(cg, 1u), 256, 0, c->stream>>>(d->mark, c->F, d->uniq, d->n_uniq); }
        c->launches += 2;
        d->cur_uniq = d->uniq; d->cur_n = d->n_uniq; d->cur_dynamic = true;
    }
    if (d->bar1_pending) {  // every owner's update of the previous step must be visible before the pull
        ProfScope prof(c, PROF_DIST_BAR1);
        rank_wait_kernel<<<1, 32, 0, c->stream>>>(d->d_peers2->bar, d->rank, d->world, 1, d->epoch);
        c->launches++;
        d->bar1_pending = false;
    }
    { ProfScope prof(c, PROF_DIST_PULL);
    pull_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d->cur_uniq, d->cur_n, d->peers, d->shift, mask, (int)c->rowlen,
                                                        c->cW, c->cV); }
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

int dist_post_step(lctr_ctx* c, int64_t rows_divisor) {
    DistState* d = c->dist;
    const unsigned mask = (unsigned)d->world - 1;
    d->epoch++;
    if (d->use_mailbox) {
        { ProfScope prof(c, PROF_DIST_PUSH);
        push_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d->cur_uniq, d->cur_n, d->peers, d->rank, d->shift, mask,
                                                            (int)c->rowlen, (int)d->rec_floats, d->region_bytes,
                                                            (unsigned)d->rec_cap, c->cgW, c->cgV, d->push_cnt);
        push_counts_kernel<<<1, 32, 0, c->stream>>>(d->peers, d->rank, d->world, d->region_bytes, (unsigned)d->rec_cap,
                                                    d->push_cnt, d->cur_dynamic ? d->cur_n : d->scratch_done); }
        c->launches += 2;
        { ProfScope prof(c, PROF_DIST_BAR0);
        if (barrier(c, 0)) return 1; }
        { ProfScope prof(c, PROF_DIST_MERGE);
        merge_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d->mailbox, d->world, d->region_bytes, (int)d->rec_floats,
                                                             (int)c->rowlen, d->shift, c->gW, c->gV, c->touched); }
        c->launches++;
    } else {
        { ProfScope prof(c, PROF_DIST_PUSH);
        push_red_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d->cur_uniq, d->cur_n, d->d_peers2->gW, d->d_peers2->gV,
                                                                d->d_peers2->touched, d->shift, mask, (int)c->rowlen,
                                                                c->cgW, c->cgV);
        if (d->cur_dynamic) { reset_counter_kernel<<<1, 1, 0, c->stream>>>(d->cur_n); c->launches++; } }
        c->launches++;
        { ProfScope prof(c, PROF_DIST_BAR0);
        if (barrier(c, 0)) return 1; }
    }
    if (launch_apply(c, rows_divisor)) return 1;
    // barrier 1 is split: signal now, wait at the start of the next step after its rank-local mark + compact
    rank_arrive_kernel<<<1, 32, 0, c->stream>>>(d->d_peers2->bar, d->rank, d->world, 1, d->epoch);
    c->launches++;
    d->bar1_pending = true;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lctr

using namespace lctr;

extern "C" {

// handles exported per rank, in this order: W, V shards, mailbox, barrier words, update_g W, V shards, touched map
int lctr_ipc_export(lctr_ctx* c, void* handles_out, size_t cap, size_t* bytes) {
    LCTR_CHECK(c && bytes, "null argument");
    LCTR_CHECK(c->dist, "lctr_ipc_export: ctx was created with world == 1");
    const size_t need = 7 * sizeof(cudaIpcMemHandle_t);
    *bytes = need;
    if (!handles_out) return 0;
    LCTR_CHECK(cap >= need, "lctr_ipc_export: need %zu bytes", need);
    cudaIpcMemHandle_t* h = reinterpret_cast<cudaIpcMemHandle_t*>(handles_out);
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[0], c->W));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[1], c->V));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[2], c->dist->mailbox));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[3], c->dist->bar));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[4], c->gW));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[5], c->gV));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[6], c->touched));
    return 0;
}

int lctr_ipc_import(lctr_ctx* c, const void* all_handles, size_t bytes_per_rank) {
    LCTR_CHECK(c && all_handles, "null argument");
    LCTR_CHECK(c->dist, "lctr_ipc_import: ctx was created with world == 1");
    LCTR_CHECK(bytes_per_rank == 7 * sizeof(cudaIpcMemHandle_t), "lctr_ipc_import: bytes_per_rank %zu", bytes_per_rank);
    DistState* d = c->dist;
    const unsigned char* base = reinterpret_cast<const unsigned char*>(all_handles);
    for (int r = 0; r < d->world; r++) {
        if (r == d->rank) continue;
        const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(base + (size_t)r * bytes_per_rank);
        for (int j = 0; j < 7; j++) {
            cudaIpcMemHandle_t hh;
            memcpy(&hh, &h[j], sizeof(hh));
            LCTR_CUDA(cudaIpcOpenMemHandle(&d->opened[r][j], hh, cudaIpcMemLazyEnablePeerAccess));
        }
        d->peers.W[r] = (float*)d->opened[r][0];
        d->peers.V[r] = (float*)d->opened[r][1];
        d->peers.mail[r] = (unsigned char*)d->opened[r][2];
        d->peers.bar[r] = (unsigned long long*)d->opened[r][3];
        d->peers2.gW[r] = (float*)d->opened[r][4];
        d->peers2.gV[r] = (float*)d->opened[r][5];
        d->peers2.touched[r] = (uint8_t*)d->opened[r][6];
        d->peers2.bar[r] = d->peers.bar[r];
    }
    LCTR_CUDA(cudaMemcpy(d->d_peers2, &d->peers2, sizeof(PeerPtrs2), cudaMemcpyHostToDevice));
    d->imported = true;
    return 0;
}

}  // extern "C"
