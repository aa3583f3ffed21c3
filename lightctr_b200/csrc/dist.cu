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

struct DistState {
    int rank = 0, world = 1, shift = 0;
    uint8_t* mark = nullptr;        // F bytes (global fid)
    uint32_t* uniq = nullptr;       // unique fids of my batch
    unsigned int* n_uniq = nullptr;
    unsigned int* push_cnt = nullptr;  // [world] records written per destination this step
    unsigned int* scratch_done = nullptr;
    // exported buffers (owned by this rank)
    unsigned char* mailbox = nullptr;  // world regions: [src][header 64 B | cap records]
    unsigned long long* bar = nullptr; // [2][world] epoch words written by peers
    size_t rec_floats = 0, rec_cap = 0, region_bytes = 0;
    PeerPtrs peers;                 // device pointers valid in THIS process
    void* opened[kMaxWorld][4] = {{nullptr}};
    bool imported = false;
    unsigned long long epoch = 0;
};

// ---------------------------------------------------------------------------------------------------------------
__global__ void mark_kernel(const uint32_t* __restrict__ fid, int64_t b, int64_t e, uint8_t* __restrict__ mark) {
    for (int64_t i = b + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x)
        mark[fid[i]] = 1;
}

// each unique row is read ONCE from its owner (peer or local shard) into the local cache; G rows per warp step
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
    for (unsigned b0 = warp * G; b0 < n; b0 += nwarps * G) {
        const unsigned idx = b0 + g;
        if (idx >= n) continue;
        const uint32_t f = uniq[idx];
        const unsigned o = f & mask;
        const size_t l = f >> shift;
        const float* src = P.V[o] + l * (size_t)rowlen;
        float* dst = cV + (size_t)f * rowlen;
        if (vec == 4) {
            for (int sl = q; sl < slices; sl += lpr)
                *reinterpret_cast<float4*>(dst + 4 * sl) = *reinterpret_cast<const float4*>(src + 4 * sl);
        } else {
            for (int sl = q; sl < slices; sl += lpr) dst[sl] = src[sl];
        }
        if (q == 0) cW[f] = P.W[o][l];
    }
}

// record layout (floats): [0] fid bits, [1] gW, [2 .. 2+rowlen) gV ; records are padded to a multiple of 4 floats
__global__ void __launch_bounds__(256)
push_kernel(const uint32_t* __restrict__ uniq, const unsigned int* __restrict__ n_uniq, PeerPtrs P, int me, int shift,
            unsigned mask, int rowlen, int rec_floats, size_t region_bytes, unsigned rec_cap,
            float* __restrict__ cgW, float* __restrict__ cgV, unsigned int* __restrict__ push_cnt) {
    const unsigned n = *n_uniq;
    const int lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    for (unsigned idx = warp; idx < n; idx += nwarps) {  // one warp per record: row stores are contiguous
        const uint32_t f = uniq[idx];
        const unsigned o = f & mask;
        unsigned slot = 0;
        if (lane == 0) slot = atomicAdd(&push_cnt[o], 1u);
        slot = __shfl_sync(kFull, slot, 0);
        if (slot >= rec_cap) continue;  // capacity is sized from max_nnz; overflow is reported by the host
        float* rec = reinterpret_cast<float*>(P.mail[o] + (size_t)me * region_bytes + 64) + (size_t)slot * rec_floats;
        float* gsrc = cgV + (size_t)f * rowlen;
        if (lane == 0) {
            rec[0] = __uint_as_float(f);
            rec[1] = cgW[f];
            cgW[f] = 0.f;
        }
        for (int i = lane; i < rowlen; i += 32) {
            rec[2 + i] = gsrc[i];
            gsrc[i] = 0.f;
        }
    }
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
            const uint32_t f = __float_as_uint(rec[0]);
            const size_t l = f >> shift;
            float* gdst = gV + l * (size_t)rowlen;
            if (rowlen % 4 == 0 && (rec_floats % 4) == 0) {
                // gV row is 16 B aligned; the record payload starts at float 2 (8 B) -> scalar loads, vector REDs
                for (int i = lane * 4; i < rowlen; i += 128)
                    red_add_v4(gdst + i, make_float4(rec[2 + i], rec[3 + i], rec[4 + i], rec[5 + i]));
            } else {
                for (int i = lane; i < rowlen; i += 32) red_add_f32(gdst + i, rec[2 + i]);
            }
            if (lane == 0) {
                red_add_f32(gW + l, rec[1]);
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
    const size_t cap = c->cfg.max_nnz ? std::min<size_t>(c->cfg.max_nnz, c->F) : c->F;
    LCTR_CUDA(cudaMalloc((void**)&d->uniq, (c->F + 32) * sizeof(uint32_t)));
    LCTR_CUDA(cudaMalloc((void**)&d->n_uniq, sizeof(unsigned int)));
    LCTR_CUDA(cudaMalloc((void**)&d->push_cnt, kMaxWorld * sizeof(unsigned int)));
    LCTR_CUDA(cudaMalloc((void**)&d->scratch_done, sizeof(unsigned int)));
    LCTR_CUDA(cudaMemsetAsync(d->n_uniq, 0, sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMemsetAsync(d->push_cnt, 0, kMaxWorld * sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMemsetAsync(d->scratch_done, 0, sizeof(unsigned int), c->stream));
    d->rec_floats = ((2 + c->rowlen) + 3) / 4 * 4;
    d->rec_cap = cap;  // records one source may send to one owner per step (<= its unique ids)
    d->region_bytes = (64 + d->rec_cap * d->rec_floats * sizeof(float) + 255) / 256 * 256;
    LCTR_CUDA(cudaMalloc((void**)&d->mailbox, d->region_bytes * R));
    LCTR_CUDA(cudaMemsetAsync(d->mailbox, 0, d->region_bytes * R, c->stream));
    LCTR_CUDA(cudaMalloc((void**)&d->bar, 2 * kMaxWorld * sizeof(unsigned long long)));
    LCTR_CUDA(cudaMemsetAsync(d->bar, 0, 2 * kMaxWorld * sizeof(unsigned long long), c->stream));
    memset(&d->peers, 0, sizeof(d->peers));
    d->peers.W[d->rank] = c->W; d->peers.V[d->rank] = c->V;
    d->peers.mail[d->rank] = d->mailbox; d->peers.bar[d->rank] = d->bar;
    return 0;
}

int dist_free(lctr_ctx* c) {
    DistState* d = c->dist;
    if (!d) return 0;
    for (int r = 0; r < d->world; r++)
        for (int j = 0; j < 4; j++)
            if (d->opened[r][j]) cudaIpcCloseMemHandle(d->opened[r][j]);
    if (c->cW) cudaFree(c->cW); if (c->cV) cudaFree(c->cV); if (c->cgW) cudaFree(c->cgW); if (c->cgV) cudaFree(c->cgV);
    c->cW = c->cV = c->cgW = c->cgV = nullptr;
    cudaFree(d->mark); cudaFree(d->uniq); cudaFree(d->n_uniq); cudaFree(d->push_cnt); cudaFree(d->scratch_done);
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

int dist_pre_step(lctr_ctx* c, Slot& s, int64_t rb, int64_t re) {
    DistState* d = c->dist;
    LCTR_CHECK(d->imported, "multi-GPU step before lctr_ipc_import");
    const int64_t eb_rows = re - rb;
    if (eb_rows <= 0) return 0;
    // entry range of the row range lives on the device; mark over the whole slot when the step covers it, else a
    // conservative host copy of the two row_ptr values
    int64_t rp[2];
    if (rb == 0 && re == s.rows) { rp[0] = 0; rp[1] = s.nnz; }
    else {
        LCTR_CUDA(cudaMemcpyAsync(&rp[0], s.row_ptr + rb, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        LCTR_CUDA(cudaMemcpyAsync(&rp[1], s.row_ptr + re, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
    }
    const unsigned mg = (unsigned)std::min<int64_t>((rp[1] - rp[0] + 255) / 256, (int64_t)c->sm_count * 8);
    mark_kernel<<<std::max(mg, 1u), 256, 0, c->stream>>>(s.fid, rp[0], rp[1], d->mark);
    const size_t ntiles = (c->F + 511) / 512;
    const unsigned cg = (unsigned)std::min<size_t>((ntiles + 7) / 8, (size_t)c->sm_count * 8);
    compact_touched_kernel<<<std::max(cg, 1u), 256, 0, c->stream>>>(d->mark, c->F, d->uniq, d->n_uniq);
    const unsigned mask = (unsigned)d->world - 1;
    pull_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d->uniq, d->n_uniq, d->peers, d->shift, mask, (int)c->rowlen,
                                                        c->cW, c->cV);
    c->launches += 3;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

int dist_post_step(lctr_ctx* c, int64_t rows_divisor) {
    DistState* d = c->dist;
    const unsigned mask = (unsigned)d->world - 1;
    push_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d->uniq, d->n_uniq, d->peers, d->rank, d->shift, mask,
                                                        (int)c->rowlen, (int)d->rec_floats, d->region_bytes,
                                                        (unsigned)d->rec_cap, c->cgW, c->cgV, d->push_cnt);
    push_counts_kernel<<<1, 32, 0, c->stream>>>(d->peers, d->rank, d->world, d->region_bytes, (unsigned)d->rec_cap,
                                                d->push_cnt, d->n_uniq);
    c->launches += 2;
    d->epoch++;
    if (barrier(c, 0)) return 1;
    merge_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d->mailbox, d->world, d->region_bytes, (int)d->rec_floats,
                                                         (int)c->rowlen, d->shift, c->gW, c->gV, c->touched);
    c->launches++;
    if (launch_apply(c, rows_divisor)) return 1;
    if (barrier(c, 1)) return 1;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lctr

using namespace lctr;

extern "C" {

// handles exported per rank, in this order: W shard, V shard, mailbox, barrier words
int lctr_ipc_export(lctr_ctx* c, void* handles_out, size_t cap, size_t* bytes) {
    LCTR_CHECK(c && bytes, "null argument");
    LCTR_CHECK(c->dist, "lctr_ipc_export: ctx was created with world == 1");
    const size_t need = 4 * sizeof(cudaIpcMemHandle_t);
    *bytes = need;
    if (!handles_out) return 0;
    LCTR_CHECK(cap >= need, "lctr_ipc_export: need %zu bytes", need);
    cudaIpcMemHandle_t* h = reinterpret_cast<cudaIpcMemHandle_t*>(handles_out);
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[0], c->W));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[1], c->V));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[2], c->dist->mailbox));
    LCTR_CUDA(cudaIpcGetMemHandle(&h[3], c->dist->bar));
    return 0;
}

int lctr_ipc_import(lctr_ctx* c, const void* all_handles, size_t bytes_per_rank) {
    LCTR_CHECK(c && all_handles, "null argument");
    LCTR_CHECK(c->dist, "lctr_ipc_import: ctx was created with world == 1");
    LCTR_CHECK(bytes_per_rank == 4 * sizeof(cudaIpcMemHandle_t), "lctr_ipc_import: bytes_per_rank %zu", bytes_per_rank);
    DistState* d = c->dist;
    const unsigned char* base = reinterpret_cast<const unsigned char*>(all_handles);
    for (int r = 0; r < d->world; r++) {
        if (r == d->rank) continue;
        const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(base + (size_t)r * bytes_per_rank);
        for (int j = 0; j < 4; j++) {
            cudaIpcMemHandle_t hh;
            memcpy(&hh, &h[j], sizeof(hh));
            LCTR_CUDA(cudaIpcOpenMemHandle(&d->opened[r][j], hh, cudaIpcMemLazyEnablePeerAccess));
        }
        d->peers.W[r] = (float*)d->opened[r][0];
        d->peers.V[r] = (float*)d->opened[r][1];
        d->peers.mail[r] = (unsigned char*)d->opened[r][2];
        d->peers.bar[r] = (unsigned long long*)d->opened[r][3];
    }
    d->imported = true;
    return 0;
}

}  // extern "C"
