// lightctr_b200/csrc/fm_fused.cuh -- order-free FM step kernels (cfg.deterministic == 0), sm_100a.
//
// Reference semantics: Train_FM_Algo::batchGradCompute + accumWVGrad + ApplyGrad (train/train_fm_algo.cpp:63-126).
// The in-order kernels of fm.cu reproduce the reference's arithmetic SEQUENCE (needed for multi-epoch 1e-5 parity);
// these kernels compute the same expressions with shuffle-tree sums (the reference's own multi-threaded Hogwild mode
// is order-free too) and are the throughput path:
//
//   batch layout (built once per batch at upload, depends only on the batch):
//     uniq[U]        the distinct feature ids of the batch ("pull_map" keys, distributed_algo_abst.h:181-190)
//     ent_slot[nnz]  position of each entry's fid in uniq  -> gradients live in a BATCH-COMPACT buffer
//                    G[U][rowlen + 4]  (row = [gV (rowlen) | gW | pad]: one 16 B-aligned record per distinct feature)
//   fm_fused_kernel   warp = sample.  One gather of the sample's V rows (K/4 lanes x 16 B per row, all rows of the
//                     sample in flight, kept in REGISTERS), shuffle-tree sumVX / |Vx|^2 / Wx, sigmoid, loss; the
//                     gradient rows are formed from the register-resident rows and leave as 16 B vector REDs into G.
//                     One gather per step instead of two; no fid-indexed dense update_g; no touched map.
//   fm_bwd_sorted_kernel  atomic-light alternative backward over the slot-sorted entry list (see below).
//   apply_compact_kernel  the per-coordinate updater over G: coalesced gradient read, gather/scatter of the U parameter
//                     and state rows, G re-zeroed in the same pass (replaces compact_touched + apply of opt.cu).
// HBM/L2-bound integer + fp32 work: no tensor cores by design.
#pragma once
#include "opt.cuh"

namespace lctr {

// ---------------------------------------------------------------------------------------------------------------
// slot map of a batch: mark -> compact (+ fid -> slot table) -> assign.  Integer work; runs on the upload stream.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
slotmap_mark_kernel(const uint32_t* __restrict__ fid, const int64_t* __restrict__ hdr, int64_t nnz_arg,
                    uint8_t* __restrict__ mark) {
    const int64_t nnz = hdr ? hdr[1] : nnz_arg;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
        mark[fid[i]] = 1;
}

// scans the byte map 16 marks per lane, clears it, appends the set positions to `uniq` (one warp-aggregated atomicAdd per
// non-empty 512-id tile) and records slot_of[fid] = position
__global__ void __launch_bounds__(256)
slotmap_compact_kernel(uint8_t* __restrict__ mark, size_t F, uint32_t* __restrict__ uniq, unsigned int* __restrict__ n_uniq,
                       uint32_t* __restrict__ slot_of) {
    const int lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    const size_t ntiles = (F + 511) / 512;
    for (size_t tile = warp; tile < ntiles; tile += nwarps) {
        const size_t base = tile * 512 + (size_t)lane * 16;
        uint4 m = make_uint4(0, 0, 0, 0);
        if (base + 16 <= F) {
            m = *reinterpret_cast<const uint4*>(mark + base);
        } else if (base < F) {
            unsigned char tmp[16];
            for (int i = 0; i < 16; i++) tmp[i] = base + i < F ? mark[base + i] : 0;
            m = *reinterpret_cast<uint4*>(tmp);
        }
        const bool any = (m.x | m.y | m.z | m.w) != 0;
        if (!__any_sync(kFull, any)) continue;
        unsigned bits = 0;
        {
            const unsigned wv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if ((wv[i] >> (8 * b)) & 0xffu) bits |= 1u << (i * 4 + b);
        }
        if (any) {
            if (base + 16 <= F) *reinterpret_cast<uint4*>(mark + base) = make_uint4(0, 0, 0, 0);
            else for (int i = 0; i < 16 && base + i < F; i++) mark[base + i] = 0;
        }
        const int mycnt = __popc(bits);
        int incl = mycnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(kFull, incl, 31);
        unsigned int gbase = 0;
        if (lane == 31) gbase = atomicAdd(n_uniq, (unsigned int)total);
        gbase = __shfl_sync(kFull, gbase, 31);
        unsigned int pos = gbase + (unsigned int)(incl - mycnt);
        while (bits) {
            const int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            const uint32_t f = (uint32_t)(base + bit);
            uniq[pos] = f;
            slot_of[f] = pos;
            pos++;
        }
    }
}

__global__ void __launch_bounds__(256)
slotmap_assign_kernel(const uint32_t* __restrict__ fid, const int64_t* __restrict__ hdr, int64_t nnz_arg,
                      const uint32_t* __restrict__ slot_of, uint32_t* __restrict__ ent_slot) {
    const int64_t nnz = hdr ? hdr[1] : nnz_arg;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
        ent_slot[i] = slot_of[fid[i]];
}

// ---------------------------------------------------------------------------------------------------------------
// fused forward (+ RED backward)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldg_u32_pinned(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// MODE 0: forward only (pred, sumVX, d = pred - label, statistics).  MODE 1: forward + RED backward into G.
// pidx: per-entry index of the PARAMETER row (the fid; or the slot when the rows live in a batch-compact cache),
// gidx: per-entry index of the gradient row in G (the slot).  SAME_IDX: pidx == gidx (one index load).
template <int K, bool HAS_VAL, int MODE, bool SAME_IDX>
__global__ void __launch_bounds__(128, 4)
fm_fused_kernel(const int64_t* __restrict__ row_ptr, const uint32_t* __restrict__ pidx, const uint32_t* __restrict__ gidx,
                const float* __restrict__ val, const float* __restrict__ label, const float* __restrict__ W,
                const float* __restrict__ V, float* __restrict__ pred, float* __restrict__ sumvx, float* __restrict__ dvec,
                float* __restrict__ G, int GS, float l2, int64_t rb, int64_t re_arg, const int64_t* __restrict__ hdr,
                double* partial, unsigned int* done, double* out_slot, int do_stats) {
    static_assert(K % 4 == 0 && K <= 32 && (K / 4 & (K / 4 - 1)) == 0, "fused FM step: K in {4, 8, 16, 32}");
    constexpr int LPR = K / 4;                           // lanes per V row (one float4 each)
    constexpr int GR = 32 / LPR;                         // rows per gather instruction
    constexpr int NPASS = K <= 8 ? 8 : (K == 16 ? 4 : 2);  // 32-entry passes whose rows stay in registers
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int q = lane % LPR, g = lane / LPR;
    const int64_t re = hdr ? hdr[0] : re_arg;
    const int64_t r = rb + (int64_t)blockIdx.x * (blockDim.x >> 5) + wid;
    double loss = 0.0, correct = 0.0;
    if (r < re) {
        const int64_t b = row_ptr[r];
        const int n = (int)(row_ptr[r + 1] - b);
        // ---- indices of the whole sample (coalesced), then every row gather, back to back
        uint32_t pf[NPASS], gs[NPASS];
        float xs[NPASS], ws[NPASS];
#pragma unroll
        for (int p = 0; p < NPASS; p++) {
            const int i = p * 32 + lane;
            const bool ok = i < n;
            pf[p] = ok ? ldg_u32_pinned(pidx + b + i) : 0u;
            gs[p] = SAME_IDX ? pf[p] : ((ok && MODE == 1) ? ldg_u32_pinned(gidx + b + i) : 0u);
            xs[p] = ok ? (HAS_VAL ? ldg_f32_pinned(val + b + i) : 1.f) : 0.f;
        }
        float4 v[NPASS][LPR];
#pragma unroll
        for (int p = 0; p < NPASS; p++) {
            if (p * 32 < n) {  // warp-uniform
                ws[p] = ldg_f32_pinned(W + pf[p]);  // lane-own W (index 0 beyond n: harmless, x = 0)
#pragma unroll
                for (int it = 0; it < LPR; it++) {
                    const uint32_t fj = __shfl_sync(kFull, pf[p], it * GR + g);
                    v[p][it] = ldg_f4_pinned(V + (size_t)fj * K + 4 * q);
                }
            } else {
                ws[p] = 0.f;
#pragma unroll
                for (int it = 0; it < LPR; it++) v[p][it] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        // ---- interaction sums (order-free): s = sum x V (this lane's 4 factors over its rows), sq = sum |xV|^2
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        float sq = 0.f, wsum = 0.f;
#pragma unroll
        for (int p = 0; p < NPASS; p++) {
            if (p * 32 < n) {
                wsum = __fmaf_rn(ws[p], xs[p], wsum);                        // fm_pred += W[fid] * X   train_fm_algo.cpp:74
#pragma unroll
                for (int it = 0; it < LPR; it++) {
                    const float xj = __shfl_sync(kFull, xs[p], it * GR + g);  // 0 beyond n
                    const float4 t = make_float4(v[p][it].x * xj, v[p][it].y * xj, v[p][it].z * xj, v[p][it].w * xj);
                    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;          // sumVX += tmp            :77
                    sq = __fmaf_rn(t.x, t.x, sq); sq = __fmaf_rn(t.y, t.y, sq);
                    sq = __fmaf_rn(t.z, t.z, sq); sq = __fmaf_rn(t.w, t.w, sq);  // dot(tmp, tmp)      :78
                }
            }
        }
        for (int base = NPASS * 32; base < n; base += 32) {  // samples longer than the register window
            const int i = base + lane;
            const bool ok = i < n;
            const uint32_t f = ok ? __ldg(pidx + b + i) : 0u;
            const float x = ok ? (HAS_VAL ? __ldg(val + b + i) : 1.f) : 0.f;
            wsum = __fmaf_rn(__ldg(W + f), x, wsum);
            float4 vv[LPR];
#pragma unroll
            for (int it = 0; it < LPR; it++) {
                const uint32_t fj = __shfl_sync(kFull, f, it * GR + g);
                vv[it] = ldg_f4(V + (size_t)fj * K + 4 * q);
            }
#pragma unroll
            for (int it = 0; it < LPR; it++) {
                const float xj = __shfl_sync(kFull, x, it * GR + g);
                const float4 t = make_float4(vv[it].x * xj, vv[it].y * xj, vv[it].z * xj, vv[it].w * xj);
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
                sq = __fmaf_rn(t.x, t.x, sq); sq = __fmaf_rn(t.y, t.y, sq);
                sq = __fmaf_rn(t.z, t.z, sq); sq = __fmaf_rn(t.w, t.w, sq);
            }
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {  // over the row groups: every lane ends with the full sumVX of its 4 factors
            s.x += __shfl_xor_sync(kFull, s.x, o); s.y += __shfl_xor_sync(kFull, s.y, o);
            s.z += __shfl_xor_sync(kFull, s.z, o); s.w += __shfl_xor_sync(kFull, s.w, o);
        }
        sq = warp_sum(sq);
        wsum = warp_sum(wsum);
        float dot = s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;  // |sumVX|^2 over the LPR lanes of a row group
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) dot += __shfl_xor_sync(kFull, dot, o);
        const float fm = wsum - 0.5f * sq + 0.5f * dot;             // :78, :82
        const float pr = ref_sigmoid(fm);                           // :84
        const float y = __ldg(label + r);
        const float d = pr - y;                                     // LogisticGradW  fm_algo_abst.h:159-161
        if (lane < LPR) *reinterpret_cast<float4*>(sumvx + (size_t)r * K + 4 * q) = s;  // FM_Algo_Abst::sumVX (:145)
        if (lane == 0) {
            pred[r] = pr;
            if (dvec) dvec[r] = d;
            if (do_stats) loss_terms(pr, y, loss, correct);
        }
        if (MODE == 1) {
            // ---- backward from the register-resident rows (train_fm_algo.cpp:101-116), vector REDs into G
#pragma unroll
            for (int p = 0; p < NPASS; p++) {
                if (p * 32 < n) {
#pragma unroll
                    for (int it = 0; it < LPR; it++) {
                        const int j = it * GR + g;
                        const float xj = __shfl_sync(kFull, xs[p], j);
                        const float wj = __shfl_sync(kFull, ws[p], j);
                        const uint32_t sj = __shfl_sync(kFull, gs[p], j);
                        if (p * 32 + j < n) {
                            const float gw = __fmaf_rn(d, xj, l2 * wj);                       // :108
                            const float4 vv = v[p][it];
                            float4 gv;
                            gv.x = __fmaf_rn(__fmaf_rn(-xj, vv.x, s.x), gw, l2 * vv.x);       // :112-115
                            gv.y = __fmaf_rn(__fmaf_rn(-xj, vv.y, s.y), gw, l2 * vv.y);
                            gv.z = __fmaf_rn(__fmaf_rn(-xj, vv.z, s.z), gw, l2 * vv.z);
                            gv.w = __fmaf_rn(__fmaf_rn(-xj, vv.w, s.w), gw, l2 * vv.w);
                            float* dst = G + (size_t)sj * GS;
                            red_add_v4(dst + 4 * q, gv);
                            if (q == 0) red_add_f32(dst + K, gw);                             // :109
                        }
                    }
                }
            }
            for (int base = NPASS * 32; base < n; base += 32) {
                const int i = base + lane;
                const bool ok = i < n;
                const uint32_t f = ok ? __ldg(pidx + b + i) : 0u;
                const uint32_t sl = ok ? __ldg(gidx + b + i) : 0u;
                const float x = ok ? (HAS_VAL ? __ldg(val + b + i) : 1.f) : 0.f;
                const float w = __ldg(W + f);
#pragma unroll
                for (int it = 0; it < LPR; it++) {
                    const int j = it * GR + g;
                    const uint32_t fj = __shfl_sync(kFull, f, j);
                    const uint32_t sj = __shfl_sync(kFull, sl, j);
                    const float xj = __shfl_sync(kFull, x, j);
                    const float wj = __shfl_sync(kFull, w, j);
                    if (base + j < n) {
                        const float4 vv = ldg_f4(V + (size_t)fj * K + 4 * q);
                        const float gw = __fmaf_rn(d, xj, l2 * wj);
                        float4 gv;
                        gv.x = __fmaf_rn(__fmaf_rn(-xj, vv.x, s.x), gw, l2 * vv.x);
                        gv.y = __fmaf_rn(__fmaf_rn(-xj, vv.y, s.y), gw, l2 * vv.y);
                        gv.z = __fmaf_rn(__fmaf_rn(-xj, vv.z, s.z), gw, l2 * vv.z);
                        gv.w = __fmaf_rn(__fmaf_rn(-xj, vv.w, s.w), gw, l2 * vv.w);
                        float* dst = G + (size_t)sj * GS;
                        red_add_v4(dst + 4 * q, gv);
                        if (q == 0) red_add_f32(dst + K, gw);
                    }
                }
            }
        }
    }
    if (do_stats) publish_stats(loss, correct, partial, done, out_slot, false);
}

// ---------------------------------------------------------------------------------------------------------------
// slot-sorted entry list ("CSC" of the batch) built with warp-aggregated cursors, and the backward over it
// ---------------------------------------------------------------------------------------------------------------
// Lane = row (a warp walks 32 rows in lock step, entry j of each): the hot ids of the small-vocabulary fields show up
// in many rows at the same position, so __match_any_sync folds them into ONE atomic per distinct id per step instead
// of one per entry (same-address atomics on the hottest ids were 75 % of the r01 grouping time).
__global__ void __launch_bounds__(256)
sorted_count_kernel(const int64_t* __restrict__ row_ptr, const uint32_t* __restrict__ ent_slot, int64_t rows_arg,
                    const int64_t* __restrict__ hdr, unsigned int* __restrict__ cnt) {
    const int64_t rows = hdr ? hdr[0] : rows_arg;
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = r < rows ? row_ptr[r] : 0;
    const int n = r < rows ? (int)(row_ptr[r + 1] - b) : 0;
    int nmax = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(kFull, nmax, o));
    for (int j = 0; j < nmax; j++) {
        const bool has = j < n;
        const uint32_t s = has ? __ldg(ent_slot + b + j) : 0xffffffffu - (uint32_t)lane;  // distinct dummies
        const unsigned m = __match_any_sync(kFull, s);
        if (has && lane == __ffs(m) - 1) atomicAdd(&cnt[s], (unsigned)__popc(m));
    }
}

// exclusive scan of cnt[0..U) by ONE block -> seg_ptr[U+1]; clears cnt (becomes the fill cursor)
__global__ void __launch_bounds__(1024)
sorted_scan_kernel(unsigned int* __restrict__ cnt, const unsigned int* __restrict__ n_uniq, unsigned int* __restrict__ seg_ptr) {
    __shared__ unsigned wsum[32];
    const unsigned n = *n_uniq;
    const unsigned per = (n + 1023) / 1024;
    const unsigned b0 = threadIdx.x * per, e0 = min(n, b0 + per);
    unsigned loc = 0;
    for (unsigned i = b0; i < e0; i++) loc += cnt[i];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned incl = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(kFull, wi, o);
            if (lane >= o) wi += t;
        }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    unsigned run = wsum[wid] + incl - loc;
    for (unsigned i = b0; i < e0; i++) {
        const unsigned c = cnt[i];
        seg_ptr[i] = run;
        cnt[i] = 0;
        run += c;
    }
    if (threadIdx.x == 1023) seg_ptr[n] = run;  // the last thread's running sum is the total (its chunk may be empty)
}

__global__ void __launch_bounds__(256)
sorted_fill_kernel(const int64_t* __restrict__ row_ptr, const uint32_t* __restrict__ fid, const uint32_t* __restrict__ ent_slot,
                   const float* __restrict__ val, int64_t rows_arg, const int64_t* __restrict__ hdr,
                   const unsigned int* __restrict__ seg_ptr, unsigned int* __restrict__ cursor, uint32_t* __restrict__ srt_row,
                   uint32_t* __restrict__ srt_slot, uint32_t* __restrict__ srt_fid, float* __restrict__ srt_x) {
    const int64_t rows = hdr ? hdr[0] : rows_arg;
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = r < rows ? row_ptr[r] : 0;
    const int n = r < rows ? (int)(row_ptr[r + 1] - b) : 0;
    int nmax = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(kFull, nmax, o));
    for (int j = 0; j < nmax; j++) {
        const bool has = j < n;
        const uint32_t s = has ? __ldg(ent_slot + b + j) : 0xffffffffu - (uint32_t)lane;
        const unsigned m = __match_any_sync(kFull, s);
        const int leader = __ffs(m) - 1;
        unsigned base = 0;
        if (has && lane == leader) base = atomicAdd(&cursor[s], (unsigned)__popc(m));
        base = __shfl_sync(kFull, base, leader);
        if (has) {
            const unsigned pos = seg_ptr[s] + base + (unsigned)__popc(m & ((1u << lane) - 1u));
            srt_row[pos] = (uint32_t)r;
            srt_slot[pos] = s;
            srt_fid[pos] = __ldg(fid + b + j);
            if (val) srt_x[pos] = __ldg(val + b + j);
        }
    }
}

// Backward over the slot-sorted entry list: a warp takes GR * EPG consecutive entries, lane group g (LPR lanes, one
// float4 of the row each) walks EPG consecutive entries: gathers sumVX[row] and V[fid] of all of them up front, forms
// each entry's gradient row (train_fm_algo.cpp:108-115) and sums runs of equal slot in registers.  A run that is a
// whole segment leaves with ONE plain 16 B store per lane; only segments cut by a lane-group boundary use REDs (one per
// EPG entries instead of one per entry).
template <int K, bool HAS_VAL>
__global__ void __launch_bounds__(128, 4)
fm_bwd_sorted_kernel(const uint32_t* __restrict__ srt_row, const uint32_t* __restrict__ srt_slot,
                     const uint32_t* __restrict__ srt_fid, const float* __restrict__ srt_x, int64_t nnz_arg,
                     const int64_t* __restrict__ hdr, const float* __restrict__ sumvx, const float* __restrict__ dvec,
                     const float* __restrict__ W, const float* __restrict__ V, float* __restrict__ G, int GS, float l2) {
    constexpr int LPR = K / 4, GR = 32 / LPR, EPG = 8;
    const int64_t nnz = hdr ? hdr[1] : nnz_arg;
    const int lane = threadIdx.x & 31;
    const int q = lane % LPR, g = lane / LPR;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t e0 = (warp * GR + g) * EPG;  // first entry of this lane group
    if (warp * GR * EPG >= nnz) return;
    uint32_t row[EPG], slot[EPG], fidv[EPG];
    float x[EPG];
#pragma unroll
    for (int j = 0; j < EPG; j++) {
        const bool ok = e0 + j < nnz;
        row[j] = ok ? ldg_u32_pinned(srt_row + e0 + j) : 0u;
        slot[j] = ok ? ldg_u32_pinned(srt_slot + e0 + j) : 0xffffffffu;
        fidv[j] = ok ? ldg_u32_pinned(srt_fid + e0 + j) : 0u;
        x[j] = HAS_VAL ? (ok ? ldg_f32_pinned(srt_x + e0 + j) : 0.f) : 1.f;
    }
    const uint32_t prevs = (e0 > 0 && e0 <= nnz) ? __ldg(srt_slot + e0 - 1) : 0xfffffffeu;
    const uint32_t nexts = (e0 + EPG < nnz) ? __ldg(srt_slot + e0 + EPG) : 0xfffffffeu;
    float4 sv[EPG], vv[EPG];
    float dd[EPG], ww[EPG];
#pragma unroll
    for (int j = 0; j < EPG; j++) {
        sv[j] = ldg_f4_pinned(sumvx + (size_t)row[j] * K + 4 * q);
        vv[j] = ldg_f4_pinned(V + (size_t)fidv[j] * K + 4 * q);
        dd[j] = ldg_f32_pinned(dvec + row[j]);
        ww[j] = ldg_f32_pinned(W + fidv[j]);
    }
    uint32_t cur = slot[0];
    bool first = true;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float accw = 0.f;
    auto flush = [&](uint32_t sl, bool whole) {
        if (sl >= 0xfffffffeu) return;
        float* dst = G + (size_t)sl * GS;
        if (whole) {
            *reinterpret_cast<float4*>(dst + 4 * q) = acc;
            if (q == 0) dst[K] = accw;
        } else {
            red_add_v4(dst + 4 * q, acc);
            if (q == 0) red_add_f32(dst + K, accw);
        }
    };
#pragma unroll
    for (int j = 0; j < EPG; j++) {
        if (slot[j] != cur) {
            flush(cur, !(first && prevs == cur));
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
            accw = 0.f;
            cur = slot[j];
            first = false;
        }
        if (slot[j] != 0xffffffffu) {
            const float gw = __fmaf_rn(dd[j], x[j], l2 * ww[j]);
            acc.x += __fmaf_rn(__fmaf_rn(-x[j], vv[j].x, sv[j].x), gw, l2 * vv[j].x);
            acc.y += __fmaf_rn(__fmaf_rn(-x[j], vv[j].y, sv[j].y), gw, l2 * vv[j].y);
            acc.z += __fmaf_rn(__fmaf_rn(-x[j], vv[j].z, sv[j].z), gw, l2 * vv[j].z);
            acc.w += __fmaf_rn(__fmaf_rn(-x[j], vv[j].w, sv[j].w), gw, l2 * vv[j].w);
            accw += gw;
        }
    }
    flush(cur, !(first && prevs == cur) && nexts != cur);
}

// ---------------------------------------------------------------------------------------------------------------
// updater over the batch-compact gradient buffer
// ---------------------------------------------------------------------------------------------------------------
// AdagradUpdater_Num::update / FTRLUpdater::update / AdamUpdater_Num::update ... restricted to the batch's features
// (exactly equivalent to the reference's dense sweep: every updater skips g == 0).  Row i of G belongs to
// feature uniq[i]; LPR lanes per row, GR rows per warp step, U steps in flight.  Zeroes G on the way (the memset of
// gradientUpdater.h:149).
template <int K, int OPT>
__global__ void __launch_bounds__(256, 2)
apply_compact_kernel(const uint32_t* __restrict__ uniq, const unsigned int* __restrict__ n_uniq, float* __restrict__ G, int GS,
                     float* __restrict__ W, float* __restrict__ V, float* __restrict__ s1W, float* __restrict__ s1V,
                     float* __restrict__ s2W, float* __restrict__ s2V, OptParams P_in, const OptParams* __restrict__ P_dev) {
    constexpr int LPR = K / 4, GR = 32 / LPR, U = 2;
    OptParams P = P_dev ? *P_dev : P_in;
    P.opt = OPT;
    constexpr bool two = OPT == LCTR_OPT_FTRL || OPT == LCTR_OPT_ADAM || OPT == LCTR_OPT_ADADELTA;
    const int lane = threadIdx.x & 31;
    const int q = lane % LPR, g = lane / LPR;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const unsigned total = *n_uniq;
    for (unsigned b0 = warp * (GR * U); b0 < total; b0 += nwarps * (GR * U)) {
        uint32_t f[U];
        bool ok[U];
        float4 g4[U], v4[U], a4[U], b4[U];
        float gw[U], w[U], a[U], bb[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned idx = b0 + u * GR + g;
            ok[u] = idx < total;
            f[u] = ok[u] ? __ldg(uniq + idx) : 0u;
            g4[u] = v4[u] = a4[u] = b4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            gw[u] = w[u] = a[u] = bb[u] = 0.f;
            if (ok[u]) {
                const float* grow = G + (size_t)idx * GS;
                const size_t o = (size_t)f[u] * K + 4 * q;
                g4[u] = *reinterpret_cast<const float4*>(grow + 4 * q);
                v4[u] = *reinterpret_cast<const float4*>(V + o);
                a4[u] = *reinterpret_cast<const float4*>(s1V + o);
                if (two) b4[u] = *reinterpret_cast<const float4*>(s2V + o);
                if (q == 0) {
                    gw[u] = grow[K]; w[u] = W[f[u]]; a[u] = s1W[f[u]];
                    if (two) bb[u] = s2W[f[u]];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!ok[u]) continue;
            const unsigned idx = b0 + u * GR + g;
            float* grow = G + (size_t)idx * GS;
            const size_t o = (size_t)f[u] * K + 4 * q;
            update_one(P, P.corrV, v4[u].x, g4[u].x, a4[u].x, b4[u].x);
            update_one(P, P.corrV, v4[u].y, g4[u].y, a4[u].y, b4[u].y);
            update_one(P, P.corrV, v4[u].z, g4[u].z, a4[u].z, b4[u].z);
            update_one(P, P.corrV, v4[u].w, g4[u].w, a4[u].w, b4[u].w);
            *reinterpret_cast<float4*>(V + o) = v4[u];
            *reinterpret_cast<float4*>(s1V + o) = a4[u];
            if (two) *reinterpret_cast<float4*>(s2V + o) = b4[u];
            *reinterpret_cast<float4*>(grow + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q == 0) {
                update_one(P, P.corrW, w[u], gw[u], a[u], bb[u]);
                W[f[u]] = w[u]; s1W[f[u]] = a[u];
                if (two) s2W[f[u]] = bb[u];
                grow[K] = 0.f;
            }
        }
    }
}

}  // namespace lctr
