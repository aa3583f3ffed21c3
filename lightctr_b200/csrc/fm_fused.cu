// lightctr_b200/csrc/fm_fused.cu -- host side of the order-free FM step (kernels: fm_fused.cuh).
//
// Train_FM_Algo::Train() per batch (train/train_fm_algo.cpp:44-57) for cfg.deterministic == 0 on one GPU:
//   upload   slot map of the batch (5 integer kernels on the upload stream: mark, compact, sample, hot, assign)
//   step     fm_fused_kernel (gather + interaction + loss + RED scatter into the batch-compact buffer)
//            apply_compact_kernel (updater over the compact buffer; re-zeroes it)
#include <algorithm>

#include "fm_fused.cuh"

namespace lctr {

// the fused kernels exist for FM and NFM (embedding side) with K in {4, 8, 16, 32}, order-free mode; on one GPU they read the tables directly,
// on several (dist.cu) the batch-compact cache of the pulled rows
bool fused_kernels_ok(const lctr_ctx* c) {
    const int k = (int)c->cfg.factor_cnt;
    return (c->cfg.model == LCTR_MODEL_FM || c->cfg.model == LCTR_MODEL_NFM) && c->cfg.deterministic == 0 &&
           (k == 4 || k == 8 || k == 16 || k == 32);
}
bool fused_supported(const lctr_ctx* c) { return c->cfg.world == 1 && fused_kernels_ok(c); }

void fused_free(lctr_ctx* c) {
    FusedState* f = c->fused;
    if (!f) return;
    cudaFree(f->mark); cudaFree(f->slot_of); cudaFree(f->cnt); cudaFree(f->G); cudaFree(f->Ghot); cudaFree(f->d_opt);
    delete f;
    c->fused = nullptr;
}

static int fused_init(lctr_ctx* c) {
    if (c->fused) return 0;
    FusedState* f = new FusedState();
    c->fused = f;
    f->T = mark_rows(c->F);
    f->GS = fused_kernels_ok(c) ? grad_stride((int)c->cfg.factor_cnt) : 0;  // other models: slot map only
    LCTR_CUDA(cudaMalloc((void**)&f->mark, 128 * f->T + 512));
    LCTR_CUDA(cudaMemsetAsync(f->mark, 0, 128 * f->T + 512, c->stream));
    LCTR_CUDA(cudaMalloc((void**)&f->slot_of, c->F * sizeof(uint32_t)));
    if (f->GS) {
        LCTR_CUDA(cudaMalloc((void**)&f->Ghot, (size_t)kHotMax * kHotRep * f->GS * sizeof(float)));
        LCTR_CUDA(cudaMemsetAsync(f->Ghot, 0, (size_t)kHotMax * kHotRep * f->GS * sizeof(float), c->stream));
    }
    LCTR_CUDA(cudaMalloc((void**)&f->d_opt, sizeof(OptParams)));
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// capacity for a batch of `nnz` entries in slot s (per-slot key set + per-context gradient buffer)
int fused_reserve(lctr_ctx* c, Slot& s, int64_t nnz) {
    if (fused_init(c)) return 1;
    FusedState* f = c->fused;
    const int64_t need_u = std::min<int64_t>(std::max<int64_t>(nnz, 1), (int64_t)c->F);
    if (nnz > s.cap_ent_slot) {
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        if (s.ent_slot) cudaFree(s.ent_slot);
        if (s.ent_pslot) cudaFree(s.ent_pslot);
        s.ent_pslot = nullptr;
        const int64_t cap = std::max<int64_t>(nnz, s.cap_ent_slot + s.cap_ent_slot / 2);
        LCTR_CUDA(cudaMalloc((void**)&s.ent_slot, (size_t)(cap + 64) * sizeof(uint32_t)));
        if (c->cfg.world > 1) LCTR_CUDA(cudaMalloc((void**)&s.ent_pslot, (size_t)(cap + 64) * sizeof(uint32_t)));
        s.cap_ent_slot = cap;
    }
    if (need_u > s.cap_uniq || !s.hot_of) {
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        if (s.uniq) cudaFree(s.uniq);
        if (s.hot_of) cudaFree(s.hot_of);
        const int64_t cap = std::min<int64_t>(std::max<int64_t>(need_u, s.cap_uniq + s.cap_uniq / 2), (int64_t)c->F);
        LCTR_CUDA(cudaMalloc((void**)&s.uniq, (size_t)(cap + 64) * sizeof(uint32_t)));
        LCTR_CUDA(cudaMalloc((void**)&s.hot_of, (size_t)(cap + 64) * sizeof(uint32_t)));
        if (!s.n_uniq) LCTR_CUDA(cudaMalloc((void**)&s.n_uniq, sizeof(unsigned int)));
        if (!s.n_hot) LCTR_CUDA(cudaMalloc((void**)&s.n_hot, sizeof(unsigned int)));
        if (!s.hot_slot) LCTR_CUDA(cudaMalloc((void**)&s.hot_slot, (size_t)kHotMax * sizeof(uint32_t)));
        s.cap_uniq = cap;
    }
    if ((size_t)s.cap_uniq > f->cnt_cap) {
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        if (c->copy_stream) LCTR_CUDA(cudaStreamSynchronize(c->copy_stream));
        if (f->cnt) cudaFree(f->cnt);
        LCTR_CUDA(cudaMalloc((void**)&f->cnt, (size_t)(s.cap_uniq + 64) * sizeof(unsigned int)));
        LCTR_CUDA(cudaMemset(f->cnt, 0, (size_t)(s.cap_uniq + 64) * sizeof(unsigned int)));
        f->cnt_cap = (size_t)s.cap_uniq;
    }
    // gradient rows: one per slot; on several GPUs one per exchange row (dist.cu re-indexes the entries)
    const size_t g_rows = c->cfg.world > 1 ? c->dist_rows : (size_t)s.cap_uniq;
    if (f->GS && g_rows > f->G_rows) {
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        if (f->G) cudaFree(f->G);
        LCTR_CUDA(cudaMalloc((void**)&f->G, (g_rows + 64) * f->GS * sizeof(float)));
        LCTR_CUDA(cudaMemset(f->G, 0, (g_rows + 64) * f->GS * sizeof(float)));
        f->G_rows = g_rows;
    }
    return 0;
}

// slot map of the batch held by slot s, on stream st.  hdr != nullptr (graph capture): {rows, nnz} are read from device
// memory and the grids are sized for the slot's capacities.
int fused_build_slot(lctr_ctx* c, Slot& s, cudaStream_t st, const int64_t* hdr, int64_t rows_cap, int64_t nnz_cap) {
    FusedState* f = c->fused;
    s.fused_valid = false;
    if (rows_cap <= 0 || nnz_cap <= 0) return 0;
    const int SM = c->sm_count;
    LCTR_CUDA(cudaMemsetAsync(s.n_uniq, 0, sizeof(unsigned int), st));
    LCTR_CUDA(cudaMemsetAsync(s.n_hot, 0, sizeof(unsigned int), st));
    const unsigned mkg = (unsigned)std::max<int64_t>(1, std::min<int64_t>((nnz_cap + 2047) / 2048, (int64_t)SM * 4));
    slotmap_mark_kernel<<<mkg, 256, 0, st>>>(s.fid, hdr, nnz_cap, f->mark, f->T);
    const size_t ntiles = (128 * f->T + 511) / 512;
    const unsigned cg = (unsigned)std::max<size_t>(1, std::min<size_t>((ntiles + 7) / 8, (size_t)SM * 8));
    slotmap_compact_kernel<<<cg, 256, 0, st>>>(f->mark, f->T, s.uniq, s.n_uniq, f->slot_of);
    const bool hot = f->GS != 0;  // replica rows only exist for the fused FM kernels
    if (hot) {
        const unsigned sg = (unsigned)std::min<int64_t>(((int64_t)kHotSampleRows * 128 + 255) / 256, (int64_t)SM * 8);
        slotmap_sample_kernel<<<sg, 256, 0, st>>>(s.row_ptr, s.fid, hdr, rows_cap, f->slot_of, f->cnt);
        slotmap_hot_kernel<<<SM * 2, 256, 0, st>>>(f->cnt, s.n_uniq, hdr, rows_cap, s.hot_of, s.hot_slot, s.n_hot);
        c->launches += 2;
    }
    const unsigned ag = (unsigned)std::max<int64_t>(1, std::min<int64_t>((nnz_cap + 255) / 256, (int64_t)SM * 8));
    slotmap_assign_kernel<<<ag, 256, 0, st>>>(s.fid, hdr, nnz_cap, f->slot_of, hot ? s.hot_of : nullptr, s.ent_slot,
                                              c->cfg.world > 1 ? s.ent_pslot : nullptr);
    c->launches += 3;
    LCTR_CUDA(cudaGetLastError());
    s.fused_valid = true;
    return 0;
}

// mode 1: FM forward + backward;  2: NFM forward (z, wide part);  3: NFM backward from dz
void launch_slotmap_compact(lctr_ctx* c, uint8_t* mark, size_t T, uint32_t* uniq, unsigned int* n_uniq, cudaStream_t st) {
    const size_t ntiles = (128 * T + 511) / 512;
    const unsigned cg = (unsigned)std::max<size_t>(1, std::min<size_t>((ntiles + 7) / 8, (size_t)c->sm_count * 8));
    slotmap_compact_kernel<<<cg, 256, 0, st>>>(mark, T, uniq, n_uniq, nullptr);
}

// programmatic dependent launches (updater behind the gradient kernel; dense kernels and the NFM backward behind their
// predecessors): LCTR_PDL=0 turns them off; read per launch, the tests toggle it
bool pdl_on() {
    const char* e = getenv("LCTR_PDL");
    return !(e && atoi(e) == 0);
}

template <int K>
static void fused_go(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, int stats, double* out_slot, const int64_t* hdr, int mode) {
    FusedState* f = c->fused;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((re - rb + 3) / 4, (int64_t)c->sm_count * 4));
    // one GPU: parameters straight from the tables (index = fid); several: from the batch-compact cache the owners filled
    // (index = plain slot), after the owners' "rows delivered" flags of this step (dist.cu)
    const bool multi = c->cfg.world > 1;
    const unsigned long long* wf = nullptr;
    int nw = 0;
    unsigned long long ep = 0;
    if (multi && mode != 3) dist_wait_info(c, &wf, &nw, &ep);
#define FUSED_ARGS s.row_ptr, multi ? s.ent_pslot : s.fid, s.ent_slot, s.val, s.label, c->cW, c->cV, s.pred, s.sumvx, nullptr, f->G, \
                   f->Ghot, f->GS, c->cfg.l2_reg, rb, re, hdr, c->stat_partial, c->stat_done, out_slot, stats, wf, nw, ep
#define FUSED_LAUNCH(HV)                                                                                                         \
    do {                                                                                                                         \
        if (mode == 1 && multi && pdl_on()) {  /* several GPUs: dependent on my own serve kernel; polls the owners' flags at its head */ \
            cudaLaunchConfig_t cfg = {};                                                                                         \
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = c->stream;                                          \
            cudaLaunchAttribute at[1];                                                                                           \
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                       \
            at[0].val.programmaticStreamSerializationAllowed = 1;                                                                \
            cfg.attrs = at; cfg.numAttrs = 1;                                                                                    \
            cudaLaunchKernelEx(&cfg, fm_fused_kernel<K, HV, 1, false>, (const int64_t*)s.row_ptr, (const uint32_t*)s.ent_pslot,  \
                               (const uint32_t*)s.ent_slot, (const float*)s.val, (const float*)s.label, (const float*)c->cW,   \
                               (const float*)c->cV, s.pred, s.sumvx, (float*)nullptr, f->G, f->Ghot, f->GS, c->cfg.l2_reg, rb, re, hdr, \
                               c->stat_partial, c->stat_done, out_slot, stats, wf, nw, ep, (float*)nullptr, (float*)nullptr);    \
        } else if (mode == 1) fm_fused_kernel<K, HV, 1, false><<<grid, 128, 0, c->stream>>>(FUSED_ARGS);                         \
        else if (mode == 2) fm_fused_kernel<K, HV, 2, false><<<grid, 128, 0, c->stream>>>(FUSED_ARGS, c->z, s.wide);             \
        else {  /* NFM backward: dependent on the dense kernels in front of it (it requests its batch data before their end) */ \
            cudaLaunchConfig_t cfg = {};                                                                                         \
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = c->stream;                                          \
            cudaLaunchAttribute at[1];                                                                                           \
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                       \
            at[0].val.programmaticStreamSerializationAllowed = 1;                                                                \
            cfg.attrs = at; cfg.numAttrs = (!multi && pdl_on()) ? 1 : 0;                                                         \
            cudaLaunchKernelEx(&cfg, fm_fused_kernel<K, HV, 3, false>, (const int64_t*)s.row_ptr, (const uint32_t*)(multi ? s.ent_pslot : s.fid), \
                               (const uint32_t*)s.ent_slot, (const float*)s.val, (const float*)s.label, (const float*)c->cW,   \
                               (const float*)c->cV, s.pred, s.sumvx, (float*)nullptr, f->G, f->Ghot, f->GS, c->cfg.l2_reg, rb, re, hdr, \
                               c->stat_partial, c->stat_done, out_slot, stats, wf, nw, ep, c->dz, (float*)nullptr);              \
        }                                                                                                                        \
    } while (0)
    if (s.has_val) FUSED_LAUNCH(true); else FUSED_LAUNCH(false);
#undef FUSED_LAUNCH
#undef FUSED_ARGS
}

static int launch_fused_mode(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool stats, const int64_t* hdr, double* out_slot_override,
                             int mode, int prof_id) {
    if (re - rb <= 0) return 0;
    LCTR_CHECK(s.fused_valid, "fused FM step on a slot without its slot map (uploaded before the context supported it?)");
    double* out_slot = out_slot_override ? out_slot_override : c->stats + 2 * (c->step % kStatRing);
    ProfScope prof(c, prof_id);
    switch ((int)c->cfg.factor_cnt) {
        case 4: fused_go<4>(c, s, rb, re, stats ? 1 : 0, out_slot, hdr, mode); break;
        case 8: fused_go<8>(c, s, rb, re, stats ? 1 : 0, out_slot, hdr, mode); break;
        case 16: fused_go<16>(c, s, rb, re, stats ? 1 : 0, out_slot, hdr, mode); break;
        default: fused_go<32>(c, s, rb, re, stats ? 1 : 0, out_slot, hdr, mode); break;
    }
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

// forward + RED backward of rows [rb, re) of the slot.  hdr != nullptr: `re` only sizes the grid, the row count comes
// from hdr[0].
int launch_fm_fused(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool stats, const int64_t* hdr, double* out_slot_override) {
    return launch_fused_mode(c, s, rb, re, stats, hdr, out_slot_override, 1, PROF_FM_FUSED);
}
// NFM embedding side around the dense layers: forward fills c->z (bi-interaction) and the slot's wide part; backward reads
// c->dz (the first dense layer's input delta) and the predictions the loss kernel left in the slot
int launch_nfm_forward_fused(lctr_ctx* c, Slot& s, int64_t rb, int64_t re) {
    return launch_fused_mode(c, s, rb, re, false, nullptr, nullptr, 2, PROF_FM_FWD);
}
int launch_nfm_backward_fused(lctr_ctx* c, Slot& s, int64_t rb, int64_t re) {
    return launch_fused_mode(c, s, rb, re, false, nullptr, nullptr, 3, PROF_FM_FUSED);
}

// order-free forward alone (predictions, sumVX, statistics): the throughput predictor of cfg.deterministic == 0 contexts and
// the kernel bench.py times for the gather roofline.  Needs no slot map.
template <int K>
static void fwd_tree_go(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, int stats, double* out_slot) {
    const int GS = grad_stride(K);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((re - rb + 3) / 4, (int64_t)c->sm_count * 4));
#define FWD_ARGS s.row_ptr, s.fid, s.fid, s.val, s.label, c->W, c->V, s.pred, s.sumvx, nullptr, nullptr, nullptr, GS, c->cfg.l2_reg, \
                 rb, re, nullptr, c->stat_partial, c->stat_done, out_slot, stats
    if (s.has_val) fm_fused_kernel<K, true, 0, true><<<grid, 128, 0, c->stream>>>(FWD_ARGS);
    else fm_fused_kernel<K, false, 0, true><<<grid, 128, 0, c->stream>>>(FWD_ARGS);
#undef FWD_ARGS
}
int launch_fm_forward_tree(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool stats) {
    if (re - rb <= 0) return 0;
    double* out_slot = c->stats + 2 * (c->step % kStatRing);
    ProfScope prof(c, PROF_FM_FWD);
    switch ((int)c->cfg.factor_cnt) {
        case 4: fwd_tree_go<4>(c, s, rb, re, stats ? 1 : 0, out_slot); break;
        case 8: fwd_tree_go<8>(c, s, rb, re, stats ? 1 : 0, out_slot); break;
        case 16: fwd_tree_go<16>(c, s, rb, re, stats ? 1 : 0, out_slot); break;
        default: fwd_tree_go<32>(c, s, rb, re, stats ? 1 : 0, out_slot); break;
    }
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

template <int K>
static void apply_go(lctr_ctx* c, Slot& s, const OptParams& P, const OptParams* P_dev) {
    FusedState* f = c->fused;
    const int main_blocks = c->sm_count * 3;
    const unsigned grid = (unsigned)(main_blocks + kHotMax / 8);  // + one warp per possible hot slot
    // Programmatic dependent launch behind the gradient kernel (one GPU; LCTR_PDL=0 turns it off): the updater's CTAs start as
    // the gradient kernel's retire, request their ids, parameter and state rows, and only then wait for its completion.
    const bool pdl_off = !pdl_on();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (c->cfg.world == 1 && !pdl_off) ? 1 : 0;
#define AC_GO(OPTC)                                                                                                          \
    cudaLaunchKernelEx(&cfg, apply_compact_kernel<K, OPTC>, (const uint32_t*)s.uniq, (const unsigned int*)s.n_uniq, f->G,          \
                       (const uint32_t*)s.hot_of, (const uint32_t*)s.hot_slot, (const unsigned int*)s.n_hot, f->Ghot, f->GS,     \
                       main_blocks, c->W, c->V, c->s1W, c->s1V, c->s2W, c->s2V, P, P_dev)
    switch (P.opt) {
        case LCTR_OPT_ADAGRAD: AC_GO(LCTR_OPT_ADAGRAD); break;
        case LCTR_OPT_FTRL: AC_GO(LCTR_OPT_FTRL); break;
        case LCTR_OPT_ADAM: AC_GO(LCTR_OPT_ADAM); break;
        case LCTR_OPT_RMSPROP: AC_GO(LCTR_OPT_RMSPROP); break;
        case LCTR_OPT_ADADELTA: AC_GO(LCTR_OPT_ADADELTA); break;
        case LCTR_OPT_PS_SGD: AC_GO(LCTR_OPT_PS_SGD); break;
        case LCTR_OPT_PS_ADAGRAD: AC_GO(LCTR_OPT_PS_ADAGRAD); break;
        case LCTR_OPT_PS_DCASGD: AC_GO(LCTR_OPT_PS_DCASGD); break;
        default: AC_GO(LCTR_OPT_PS_DCASGDA); break;
    }
#undef AC_GO
}

// updater over the slot's key set.  P_host (optional) / dP: parameters already staged in device memory (graph launches).
int launch_apply_compact(lctr_ctx* c, Slot& s, int64_t rows_in_step, const OptParams* P_host, const OptParams* dP) {
    const OptParams P = P_host ? *P_host : make_opt_params(c, rows_in_step);
    ProfScope prof(c, PROF_APPLY_COMPACT);
    switch ((int)c->cfg.factor_cnt) {
        case 4: apply_go<4>(c, s, P, dP); break;
        case 8: apply_go<8>(c, s, P, dP); break;
        case 16: apply_go<16>(c, s, P, dP); break;
        default: apply_go<32>(c, s, P, dP); break;
    }
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

void fused_opt_params(lctr_ctx* c, int64_t rows, void* out) { *reinterpret_cast<OptParams*>(out) = make_opt_params(c, rows); }
void* fused_dev_opt(lctr_ctx* c) { return c->fused ? (void*)c->fused->d_opt : nullptr; }

}  // namespace lctr
