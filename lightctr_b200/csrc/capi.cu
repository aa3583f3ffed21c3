// lightctr_b200/csrc/capi.cu -- the C ABI declared in include/lightctr_b200.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "opt.cuh"

namespace lctr {
static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

template <typename T>
static int dalloc(T** p, size_t n) {
    *p = nullptr;
    if (n == 0) return 0;
    LCTR_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
    return 0;
}
template <typename T>
static void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

static int slot_reserve(lctr_ctx* c, Slot& s, int64_t rows, int64_t nnz) {
    const size_t k = c->cfg.factor_cnt;
    if (rows > s.cap_rows) {
        int64_t cap = std::max<int64_t>(rows, s.cap_rows + s.cap_rows / 2);
        dfree(s.row_ptr); dfree(s.label); dfree(s.pred); dfree(s.sumvx); dfree(s.wide);
        if (dalloc(&s.row_ptr, (size_t)cap + 1)) return 1;
        if (dalloc(&s.label, (size_t)cap)) return 1;
        if (dalloc(&s.pred, (size_t)cap)) return 1;
        if (dalloc(&s.wide, (size_t)cap)) return 1;
        if (c->cfg.model != LCTR_MODEL_FFM) {
            if (dalloc(&s.sumvx, (size_t)cap * k)) return 1;
            LCTR_CUDA(cudaMemsetAsync(s.sumvx, 0, (size_t)cap * k * sizeof(float), c->stream));
        }
        s.cap_rows = cap;
    }
    if (nnz > s.cap_nnz) {
        int64_t cap = std::max<int64_t>(nnz, s.cap_nnz + s.cap_nnz / 2);
        dfree(s.fid); dfree(s.field); dfree(s.val);
        if (dalloc(&s.fid, (size_t)cap + 32)) return 1;
        if (dalloc(&s.field, (size_t)cap + 32)) return 1;
        if (dalloc(&s.val, (size_t)cap + 32)) return 1;
        s.cap_nnz = cap;
    }
    return 0;
}

__global__ void fill_value_kernel(float* p, size_t n, float v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void label_to_float_kernel(const int32_t* in, float* out, int64_t n_arg, const int64_t* hdr) {
    const int64_t n = hdr ? hdr[0] : n_arg;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// synthetic initialisation on the device (bench only): W = 0, V[f][j] = scale * N(0,1) from a counter-based hash of
// the GLOBAL element index, so the values do not depend on how the table is sharded
__global__ void fill_params_kernel(float* __restrict__ W, float* __restrict__ V, size_t Fl, size_t rowlen, int rank,
                                   int world, unsigned long long seed, float scale) {
    const size_t n = Fl * rowlen;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t l = i / rowlen, j = i % rowlen;
        const unsigned long long g = ((unsigned long long)(l * world + rank)) * rowlen + j;
        unsigned long long h = g * 0x9E3779B97F4A7C15ull + seed;
        h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
        const float u1 = ((unsigned)(h & 0xffffffu) + 1u) * (1.0f / 16777217.0f);
        const float u2 = (unsigned)((h >> 24) & 0xffffffu) * (1.0f / 16777216.0f);
        V[i] = scale * sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2);
        if (j == 0) W[l] = 0.f;
    }
}

// Feature-major view of a slot, built on the host from the caller's CSR arrays (one counting sort per row
// block; stable, so each fid's entries stay in ascending row order -- the accumulation order of the
// reference's canonical single-thread run, train_fm_algo.cpp:101-116).
static int build_csc(lctr_ctx* c, Slot& s, int64_t rows, int64_t nnz, const int64_t* row_ptr, const uint32_t* fid,
                     const float* val) {
    const int64_t block = c->cfg.csc_row_block ? (int64_t)c->cfg.csc_row_block : std::max<int64_t>(rows, 1);
    const int64_t nblocks = (rows + block - 1) / block;
    std::vector<int64_t> blk_seg_ptr(nblocks + 1, 0), seg_ptr;
    std::vector<uint32_t> seg_fid, ent_row((size_t)nnz);
    std::vector<float> ent_x(val ? (size_t)nnz : 0);
    seg_ptr.reserve((size_t)nnz / 4 + 16);
    seg_fid.reserve((size_t)nnz / 4 + 16);
    std::vector<uint32_t> cnt(c->F + 1, 0);
    std::vector<uint32_t> touched_list;
    int64_t out = 0;
    for (int64_t bi = 0; bi < nblocks; bi++) {
        const int64_t rb = bi * block, re = std::min(rows, rb + block);
        const int64_t eb = row_ptr[rb], ee = row_ptr[re];
        touched_list.clear();
        for (int64_t e = eb; e < ee; e++) {
            const uint32_t f = fid[e];
            if (f >= c->F) { set_error("upload_batch: fid %u >= feature_cnt %zu", f, c->F); return 1; }
            if (cnt[f]++ == 0) touched_list.push_back(f);
        }
        std::sort(touched_list.begin(), touched_list.end());
        // segment offsets for this block; cnt[f] becomes the write cursor
        for (uint32_t f : touched_list) {
            seg_fid.push_back(f);
            seg_ptr.push_back(out);
            const uint32_t n = cnt[f];
            cnt[f] = (uint32_t)(out - eb);  // cursor relative to the block (fits u32: block nnz < 2^32)
            out += n;
        }
        for (int64_t r = rb; r < re; r++)
            for (int64_t e = row_ptr[r]; e < row_ptr[r + 1]; e++) {
                const uint32_t f = fid[e];
                const int64_t pos = eb + cnt[f]++;
                ent_row[(size_t)pos] = (uint32_t)r;
                if (val) ent_x[(size_t)pos] = val[e];
            }
        for (uint32_t f : touched_list) cnt[f] = 0;
        blk_seg_ptr[bi + 1] = (int64_t)seg_fid.size();
    }
    seg_ptr.push_back(out);
    const int64_t nseg = (int64_t)seg_fid.size();
    if (nseg > s.cap_segs) {
        dfree(s.seg_ptr); dfree(s.seg_fid);
        if (dalloc(&s.seg_ptr, (size_t)nseg + 1) || dalloc(&s.seg_fid, (size_t)nseg + 1)) return 1;
        s.cap_segs = nseg;
    }
    if (nblocks > s.cap_blocks) {
        dfree(s.blk_seg_ptr);
        if (dalloc(&s.blk_seg_ptr, (size_t)nblocks + 1)) return 1;
        s.cap_blocks = nblocks;
    }
    if (nnz > s.cap_ent) {
        dfree(s.ent_row); dfree(s.ent_x);
        if (dalloc(&s.ent_row, (size_t)nnz + 32) || dalloc(&s.ent_x, (size_t)nnz + 32)) return 1;
        s.cap_ent = nnz;
    }
    LCTR_CUDA(cudaMemcpyAsync(s.seg_ptr, seg_ptr.data(), (size_t)(nseg + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    if (nseg) LCTR_CUDA(cudaMemcpyAsync(s.seg_fid, seg_fid.data(), (size_t)nseg * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    LCTR_CUDA(cudaMemcpyAsync(s.blk_seg_ptr, blk_seg_ptr.data(), (size_t)(nblocks + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    if (nnz) {
        LCTR_CUDA(cudaMemcpyAsync(s.ent_row, ent_row.data(), (size_t)nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        if (val) LCTR_CUDA(cudaMemcpyAsync(s.ent_x, ent_x.data(), (size_t)nnz * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    LCTR_CUDA(cudaStreamSynchronize(c->stream));  // the staging vectors die with this frame
    if (!s.h_blk_seg_ptr) s.h_blk_seg_ptr = new std::vector<int64_t>();
    *s.h_blk_seg_ptr = blk_seg_ptr;
    s.csc_block = block; s.n_blocks = nblocks; s.n_segs = nseg;
    return 0;
}
}  // namespace lctr

using namespace lctr;

extern "C" {

const char* lctr_last_error(void) { return g_err.c_str(); }
int lctr_abi_version(void) { return LCTR_ABI_VERSION; }

int lctr_create(const lctr_cfg* cfg, lctr_ctx** out) {
    LCTR_CHECK(cfg && out, "lctr_create: null argument");
    LCTR_CHECK(cfg->abi_version == LCTR_ABI_VERSION, "lctr_create: abi_version %u != %u", cfg->abi_version,
               LCTR_ABI_VERSION);
    LCTR_CHECK(cfg->model >= LCTR_MODEL_FM && cfg->model <= LCTR_MODEL_WND, "lctr_create: bad model %d", cfg->model);
    LCTR_CHECK(cfg->optimizer >= LCTR_OPT_ADAGRAD && cfg->optimizer <= LCTR_OPT_PS_DCASGDA, "lctr_create: bad optimizer %d",
               cfg->optimizer);
    LCTR_CHECK(cfg->feature_cnt > 0 && cfg->feature_cnt < (1ull << 32), "lctr_create: feature_cnt out of range");
    LCTR_CHECK(cfg->factor_cnt > 0, "lctr_create: factor_cnt must be > 0");
    LCTR_CHECK(cfg->model != LCTR_MODEL_FFM || cfg->field_cnt > 0, "lctr_create: FFM needs field_cnt > 0");
    if (cfg->model == LCTR_MODEL_WND) {
        LCTR_CHECK(cfg->field_cnt > 0 && cfg->field_cnt <= 2048, "lctr_create: Wide&Deep needs 0 < field_cnt <= 2048");
        LCTR_CHECK(cfg->deterministic == 0, "lctr_create: Wide&Deep uses the RED scatter (deterministic = 0)");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("lctr_create: no CUDA device available (%s); this library has no CPU fallback",
                  cudaGetErrorString(e));
        return 2;
    }
    LCTR_CHECK(cfg->device >= 0 && cfg->device < ndev, "lctr_create: device %d of %d", cfg->device, ndev);
    LCTR_CUDA(cudaSetDevice(cfg->device));
    lctr_ctx* c = new lctr_ctx();
    c->cfg = *cfg;
    if (c->cfg.ftrl_alpha == 0.f) {  // gradientUpdater.h:275
        c->cfg.ftrl_alpha = 0.15f; c->cfg.ftrl_lambda1 = 1.0f; c->cfg.ftrl_beta = 1.0f; c->cfg.ftrl_lambda2 = 1.0f;
    }
    if (c->cfg.world <= 0) { c->cfg.world = 1; c->cfg.rank = 0; }
    c->F = cfg->feature_cnt;
    c->Fl = (c->F + (size_t)c->cfg.world - 1) / (size_t)c->cfg.world;
    LCTR_CHECK(c->cfg.world == 1 || !cfg->deterministic, "lctr_create: deterministic modes are single-GPU only");
    LCTR_CHECK(cfg->deterministic >= 0 && cfg->deterministic <= 2, "lctr_create: deterministic must be 0, 1 or 2");
    c->rowlen = cfg->model == LCTR_MODEL_FFM ? (size_t)cfg->field_cnt * cfg->factor_cnt : cfg->factor_cnt;
    cudaDeviceProp prop;
    LCTR_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    c->sm_count = prop.multiProcessorCount;
    LCTR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const size_t FL = c->Fl;
    const size_t nv = FL * c->rowlen;
    const bool two = cfg->optimizer == LCTR_OPT_FTRL || cfg->optimizer == LCTR_OPT_ADAM || cfg->optimizer == LCTR_OPT_ADADELTA ||
                     cfg->optimizer == LCTR_OPT_PS_DCASGD || cfg->optimizer == LCTR_OPT_PS_DCASGDA;
    int rc = 0;
    rc |= dalloc(&c->W, FL); rc |= dalloc(&c->V, nv);
    rc |= dalloc(&c->gW, FL); rc |= dalloc(&c->gV, nv);
    rc |= dalloc(&c->s1W, FL); rc |= dalloc(&c->s1V, nv);
    if (two) { rc |= dalloc(&c->s2W, FL); rc |= dalloc(&c->s2V, nv); }
    rc |= dalloc(&c->touched, FL + 512);
    rc |= dalloc(&c->touch_list, FL + 32);
    rc |= dalloc(&c->n_touch, 1);
    rc |= dalloc(&c->apply_done, 1);
    rc |= dalloc(&c->stats, (size_t)2 * kStatRing);
    rc |= dalloc(&c->stat_partial, 2);
    rc |= dalloc(&c->stat_done, 1);
    if (rc) { lctr_destroy(c); return 1; }
    LCTR_CUDA(cudaMallocHost((void**)&c->h_stats, 2 * sizeof(double)));
    LCTR_CUDA(cudaMemsetAsync(c->W, 0, FL * sizeof(float), c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->V, 0, nv * sizeof(float), c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->gW, 0, FL * sizeof(float), c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->gV, 0, nv * sizeof(float), c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->s1W, 0, FL * sizeof(float), c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->s1V, 0, nv * sizeof(float), c->stream));
    if (two) {
        LCTR_CUDA(cudaMemsetAsync(c->s2W, 0, FL * sizeof(float), c->stream));
        LCTR_CUDA(cudaMemsetAsync(c->s2V, 0, nv * sizeof(float), c->stream));
    }
    if (cfg->optimizer == LCTR_OPT_PS_ADAGRAD || cfg->optimizer == LCTR_OPT_PS_DCASGDA) {  // data_accum = 1e-7 (paramserver.h:323)
        fill_value_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->s1W, FL, 1e-7f);
        fill_value_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->s1V, nv, 1e-7f);
    }
    LCTR_CUDA(cudaMemsetAsync(c->touched, 0, FL + 512, c->stream));
    if (c->cfg.world > 1) {
        if (dist_alloc(c)) { lctr_destroy(c); return 1; }
    } else {
        c->cW = c->W; c->cV = c->V; c->cgW = c->gW; c->cgV = c->gV;
    }
    LCTR_CUDA(cudaMemsetAsync(c->n_touch, 0, sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->apply_done, 0, sizeof(unsigned int), c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->stats, 0, sizeof(double) * 2 * kStatRing, c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->stat_partial, 0, sizeof(double) * 2, c->stream));
    LCTR_CUDA(cudaMemsetAsync(c->stat_done, 0, sizeof(unsigned int), c->stream));
    if (cfg->model == LCTR_MODEL_NFM || cfg->model == LCTR_MODEL_WND) {
        if (mlp_alloc(c)) { lctr_destroy(c); return 1; }
    }
    { const char* e = getenv("LCTR_CSC_IN_STEP"); c->csc_in_step = e && e[0] == '1'; }
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    *out = c;
    return 0;
}

int lctr_destroy(lctr_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    dfree(c->W); dfree(c->V); dfree(c->gW); dfree(c->gV); dfree(c->s1W); dfree(c->s1V); dfree(c->s2W); dfree(c->s2V);
    dfree(c->touched); dfree(c->touch_list); dfree(c->n_touch); dfree(c->apply_done); dfree(c->stats); dfree(c->stat_partial); dfree(c->stat_done);
    for (auto& s : c->slots) {
        dfree(s.row_ptr); dfree(s.fid); dfree(s.field); dfree(s.val); dfree(s.label); dfree(s.pred); dfree(s.sumvx);
        dfree(s.wide);
        dfree(s.blk_seg_ptr); dfree(s.seg_ptr); dfree(s.seg_fid); dfree(s.ent_row); dfree(s.ent_x); dfree(s.ent_field);
        dfree(s.ent_slot); dfree(s.ent_pslot); dfree(s.hot_of); dfree(s.hot_slot); dfree(s.n_hot);
        dfree(s.uniq); dfree(s.n_uniq); dfree(s.short_list); dfree(s.long_list); dfree(s.csc_totals); dfree(s.csc_acc); dfree(s.csc_arrived);
        delete s.h_blk_seg_ptr; s.h_blk_seg_ptr = nullptr;
    }
    mlp_free(c);
    ffm_grouped_free(c);
    metrics_free(c);
    wnd_free(c);
    dist_free(c);
    fused_free(c);
    csc_scratch_free(c);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    if (c->h_stat_ring) cudaFreeHost(c->h_stat_ring);
    for (int p = 0; p < kPipe; p++) {
        PipeGraph& g = c->pipe_graph[p];
        if (g.build) cudaGraphExecDestroy(g.build);
        if (g.step) cudaGraphExecDestroy(g.step);
        if (g.d_hdr) cudaFree(g.d_hdr); if (g.h_hdr) cudaFreeHost(g.h_hdr);
        if (g.d_opt) cudaFree(g.d_opt); if (g.h_opt) cudaFreeHost(g.h_opt);
        if (g.d_stat) cudaFree(g.d_stat); if (g.h_stat) cudaFreeHost(g.h_stat);
    }
    if (c->copy_stream) {
        for (int i = 0; i < kPipe; i++) { cudaEventDestroy(c->ev_copied[i]); cudaEventDestroy(c->ev_computed[i]); cudaEventDestroy(c->ev_h2d[i]); }
        if (c->build_stream) cudaStreamDestroy(c->build_stream);
        for (int i = 0; i < kStatRing; i++) cudaEventDestroy(c->ev_stat[i]);
        cudaStreamDestroy(c->copy_stream);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int lctr_sync(lctr_ctx* c) {
    LCTR_CHECK(c, "null ctx");
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// world > 1: W / V are the FULL (global) arrays; each rank keeps / returns only the rows it owns (fid % world == rank)
int lctr_upload_params(lctr_ctx* c, const float* W, const float* V) {
    LCTR_CHECK(c, "null ctx");
    const int R = c->cfg.world, me = c->cfg.rank;
    if (R == 1) {
        if (W) LCTR_CUDA(cudaMemcpyAsync(c->W, W, c->F * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        if (V) LCTR_CUDA(cudaMemcpyAsync(c->V, V, c->F * c->rowlen * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    } else {
        std::vector<float> w, v;
        if (W) {
            w.assign(c->Fl, 0.f);
            for (size_t f = (size_t)me, l = 0; f < c->F; f += R, l++) w[l] = W[f];
            LCTR_CUDA(cudaMemcpyAsync(c->W, w.data(), c->Fl * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        }
        if (V) {
            v.assign(c->Fl * c->rowlen, 0.f);
            for (size_t f = (size_t)me, l = 0; f < c->F; f += R, l++)
                memcpy(&v[l * c->rowlen], V + f * c->rowlen, c->rowlen * sizeof(float));
            LCTR_CUDA(cudaMemcpyAsync(c->V, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        }
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    }
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int lctr_download_params(lctr_ctx* c, float* W, float* V) {
    LCTR_CHECK(c, "null ctx");
    const int R = c->cfg.world, me = c->cfg.rank;
    if (R == 1) {
        if (W) LCTR_CUDA(cudaMemcpyAsync(W, c->W, c->F * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (V) LCTR_CUDA(cudaMemcpyAsync(V, c->V, c->F * c->rowlen * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    }
    std::vector<float> w(c->Fl), v(c->Fl * c->rowlen);
    LCTR_CUDA(cudaMemcpyAsync(w.data(), c->W, w.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    LCTR_CUDA(cudaMemcpyAsync(v.data(), c->V, v.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t f = (size_t)me, l = 0; f < c->F; f += R, l++) {
        if (W) W[f] = w[l];
        if (V) memcpy(V + f * c->rowlen, &v[l * c->rowlen], c->rowlen * sizeof(float));
    }
    return 0;
}
int lctr_fill_params(lctr_ctx* c, uint64_t seed, float scale) {
    LCTR_CHECK(c, "null ctx");
    fill_params_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->W, c->V, c->Fl, c->rowlen, c->cfg.rank, c->cfg.world,
                                                               (unsigned long long)seed, scale);
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int lctr_download_opt_state(lctr_ctx* c, float* s1, float* s2) {
    LCTR_CHECK(c, "null ctx");
    LCTR_CHECK(c->cfg.world == 1, "optimizer-state transfer is single-GPU only");
    const size_t nv = c->F * c->rowlen;
    if (s1) {
        LCTR_CUDA(cudaMemcpyAsync(s1, c->s1W, c->F * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        LCTR_CUDA(cudaMemcpyAsync(s1 + c->F, c->s1V, nv * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    if (s2 && c->s2W) {
        LCTR_CUDA(cudaMemcpyAsync(s2, c->s2W, c->F * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        LCTR_CUDA(cudaMemcpyAsync(s2 + c->F, c->s2V, nv * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int lctr_upload_opt_state(lctr_ctx* c, const float* s1, const float* s2) {
    LCTR_CHECK(c, "null ctx");
    LCTR_CHECK(c->cfg.world == 1, "optimizer-state transfer is single-GPU only (the state arrays are sharded by owner)");
    const size_t nv = c->F * c->rowlen;
    if (s1) {
        LCTR_CUDA(cudaMemcpyAsync(c->s1W, s1, c->F * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        LCTR_CUDA(cudaMemcpyAsync(c->s1V, s1 + c->F, nv * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    if (s2 && c->s2W) {
        LCTR_CUDA(cudaMemcpyAsync(c->s2W, s2, c->F * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        LCTR_CUDA(cudaMemcpyAsync(c->s2V, s2 + c->F, nv * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static int upload_batch_on(lctr_ctx* c, cudaStream_t st, int slot, int64_t rows, int64_t nnz, const int64_t* row_ptr,
                           const uint32_t* fid, const uint16_t* field, const float* val, const int32_t* label) {
    LCTR_CHECK(c, "null ctx");
    LCTR_CHECK(slot >= 0 && slot < kNumSlots, "slot %d out of range", slot);
    LCTR_CHECK(rows >= 0 && nnz >= 0 && row_ptr && (nnz == 0 || fid) && (rows == 0 || label), "upload_batch: null input");
    LCTR_CHECK((c->cfg.model != LCTR_MODEL_FFM && c->cfg.model != LCTR_MODEL_WND) || field || nnz == 0,
               "upload_batch: FFM / Wide&Deep need the field array");
    Slot& s = c->slots[slot];
    if (rows > s.cap_rows || nnz > s.cap_nnz) {
        LCTR_CUDA(cudaStreamSynchronize(c->stream));  // buffers about to be reallocated may still be in use
        if (slot_reserve(c, s, rows, nnz)) return 1;
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
    }
    s.rows = rows; s.nnz = nnz;
    s.has_val = val != nullptr;
    s.has_field = field != nullptr;
    LCTR_CUDA(cudaMemcpyAsync(s.row_ptr, row_ptr, (size_t)(rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (nnz) {
        LCTR_CUDA(cudaMemcpyAsync(s.fid, fid, (size_t)nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        if (field) LCTR_CUDA(cudaMemcpyAsync(s.field, field, (size_t)nnz * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
        if (val) LCTR_CUDA(cudaMemcpyAsync(s.val, val, (size_t)nnz * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    const bool grouped = c->cfg.deterministic == 2 && rows > 0 && nnz > 0 && c->cfg.world == 1;
    int32_t* tmp = reinterpret_cast<int32_t*>(s.pred);  // pred is overwritten by the next forward anyway
    if (rows) {
        // labels travel as int32 and are widened on device (the reference compares a `float target`)
        LCTR_CUDA(cudaMemcpyAsync(tmp, label, (size_t)rows * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        if (!grouped) {
            label_to_float_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(tmp, s.label, rows, nullptr);
            c->launches++;
            LCTR_CUDA(cudaGetLastError());
        }
    }
    s.fused_valid = false;
    if ((fused_supported(c) || c->cfg.world > 1) && rows > 0 && nnz > 0) {  // fused_supported: FM and NFM, one GPU
        // slot map of the batch: the gradient rows of the order-free fused step; on several GPUs also the key set of the
        // pull / push exchange, whose per-owner lists go out right away (posted stores, overlapping the previous step)
        if (fused_reserve(c, s, nnz) || fused_build_slot(c, s, st, nullptr, rows, nnz)) return 1;
        if (c->cfg.world > 1 && dist_send_keys(c, s, slot, st)) return 1;
    }
    s.csc_block = 0;
    s.dev_csc = false;
    if (grouped) {
        LCTR_CHECK(csc_device_supported(c), "cfg.deterministic=2 (device-grouped backward) needs FM with k in {4,8,16,32} "
                                            "or FFM with k %% 4 == 0 and field_cnt * k <= 512");
        if (csc_build_device(c, s, st, tmp, nullptr, rows, nnz)) return 1;  // widens the labels in its first kernel
    } else if (c->cfg.deterministic == 1 && c->cfg.model != LCTR_MODEL_FFM && rows > 0) {
        LCTR_CUDA(cudaStreamSynchronize(st));
        if (build_csc(c, s, rows, nnz, row_ptr, fid, val)) return 1;
    }
    return 0;
}

int lctr_upload_batch(lctr_ctx* c, int slot, int64_t rows, int64_t nnz, const int64_t* row_ptr, const uint32_t* fid,
                      const uint16_t* field, const float* val, const int32_t* label) {
    LCTR_CHECK(c, "null ctx");
    // Resident datasets are validated once, on the host: an out-of-range id would otherwise surface as an illegal
    // address inside a gather (the reference indexes W / V unchecked too, fm_algo_abst.h:146-151, but there feature_cnt
    // is derived from the same file).  The streamed entry points (lctr_train_batch[_async]) trust their caller.
    LCTR_CHECK(rows >= 0 && nnz >= 0 && row_ptr && (nnz == 0 || fid), "upload_batch: null input");
    LCTR_CHECK(row_ptr[0] == 0 && row_ptr[rows] == nnz, "upload_batch: row_ptr must run from 0 to nnz (%lld .. %lld, nnz %lld)",
               (long long)row_ptr[0], (long long)row_ptr[rows], (long long)nnz);
    for (int64_t r = 0; r < rows; r++)
        LCTR_CHECK(row_ptr[r] <= row_ptr[r + 1], "upload_batch: row_ptr decreases at row %lld", (long long)r);
    for (int64_t i = 0; i < nnz; i++)
        LCTR_CHECK(fid[i] < c->F, "upload_batch: fid %u at entry %lld >= feature_cnt %zu", fid[i], (long long)i, c->F);
    if ((c->cfg.model == LCTR_MODEL_FFM || c->cfg.model == LCTR_MODEL_WND) && field)
        for (int64_t i = 0; i < nnz; i++)
            LCTR_CHECK(field[i] < c->cfg.field_cnt, "upload_batch: field %u at entry %lld >= field_cnt %u", (unsigned)field[i],
                       (long long)i, c->cfg.field_cnt);
    return upload_batch_on(c, c->stream, slot, rows, nnz, row_ptr, fid, field, val, label);
}

static int read_stats(lctr_ctx* c, uint64_t step, float* loss_sum, float* acc_cnt) {
    if (!loss_sum && !acc_cnt) return 0;
    if (c->cfg.world > 1 && dist_check_overflow(c)) return 1;
    LCTR_CUDA(cudaMemcpyAsync(c->h_stats, c->stats + 2 * (step % kStatRing), 2 * sizeof(double), cudaMemcpyDeviceToHost,
                              c->stream));
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    if (loss_sum) *loss_sum = (float)c->h_stats[0];
    if (acc_cnt) *acc_cnt = (float)c->h_stats[1];
    return 0;
}

int lctr_train_step(lctr_ctx* c, int slot, int64_t rb, int64_t re, float* loss_sum, float* acc_cnt) {
    LCTR_CHECK(c, "null ctx");
    LCTR_CHECK(slot >= 0 && slot < kNumSlots, "slot %d out of range", slot);
    Slot& s = c->slots[slot];
    LCTR_CHECK(rb >= 0 && re <= s.rows && rb <= re, "train_step: rows [%lld,%lld) outside slot (%lld rows)",
               (long long)rb, (long long)re, (long long)s.rows);
    LCTR_CHECK(c->cfg.world == 1 || c->cfg.minibatch_size > 0,
               "train_step: multi-GPU contexts need cfg.minibatch_size = the GLOBAL batch (the updater's divisor)");
    const uint64_t step = c->step;
    int rc = 0;
    if (c->csc_in_step && c->cfg.deterministic == 2 && c->cfg.world == 1 && rb == 0 && re == s.rows && s.nnz > 0) {
        // bench mode: the grouping of the batch (count / scan / fill) is part of the timed step instead of the upload
        ProfScope prof(c, PROF_CSC_BUILD);
        if (csc_build_device(c, s, c->stream, nullptr, nullptr, s.rows, s.nnz)) return 1;
    }
    switch (c->cfg.model) {
        case LCTR_MODEL_FM:
            if (c->cfg.world > 1 && fused_kernels_ok(c))
                rc = dist_pre_step(c, s, slot, true) || launch_fm_fused(c, s, rb, re, true, nullptr, nullptr) ||
                     dist_post_step(c, s, slot, re - rb);
            else if (c->cfg.world > 1)
                rc = dist_pre_step(c, s, slot, false) || launch_fm_forward(c, s, rb, re, false, true) ||
                     launch_fm_backward(c, s, rb, re, false) || dist_post_step(c, s, slot, re - rb);
            else if (c->cfg.deterministic == 2)
                rc = launch_fm_forward(c, s, rb, re, false, true) || launch_fm_backward_devcsc(c, s, rb, re);
            else if (c->cfg.deterministic)
                rc = launch_fm_forward(c, s, rb, re, false, true) || launch_fm_backward_csc(c, s, rb, re, false);
            else if (s.fused_valid)  // order-free: one gather, RED scatter into the batch-compact buffer, compact updater
                rc = launch_fm_fused(c, s, rb, re, true, nullptr, nullptr) || launch_apply_compact(c, s, re - rb, nullptr, nullptr);
            else
                rc = launch_fm_forward(c, s, rb, re, false, true) || launch_fm_backward(c, s, rb, re, false) ||
                     launch_apply(c, re - rb);
            break;
        case LCTR_MODEL_FFM:
            if (c->cfg.world > 1)
                rc = dist_pre_step(c, s, slot, false) || launch_ffm_forward(c, s, rb, re, true) || dist_post_step(c, s, slot, re - rb);
            else if (c->cfg.deterministic == 2)
                rc = ffm_grouped_reserve(c, s.rows) || launch_ffm_forward_tiles(c, s, rb, re) ||
                     launch_ffm_backward_grouped(c, s, rb, re);
            else
                rc = launch_ffm_forward(c, s, rb, re, true) || launch_ffm_backward(c, s, rb, re) || launch_apply(c, re - rb);
            break;
        case LCTR_MODEL_WND:
            if (c->cfg.world > 1)  // rows from the owners' shards into the batch-compact cache, gradients back to the owners
                rc = dist_pre_step(c, s, slot, false) || mlp_reserve(c, re - rb) || wnd_reserve(c, re - rb) ||
                     launch_wnd_forward(c, s, rb, re) || launch_nfm_mlp(c, s, rb, re, re - rb) || launch_wnd_backward(c, s, rb, re) ||
                     dist_post_step(c, s, slot, re - rb);
            else
                rc = mlp_reserve(c, re - rb) || wnd_reserve(c, re - rb) || launch_wnd_forward(c, s, rb, re) ||
                     launch_nfm_mlp(c, s, rb, re, re - rb) || launch_wnd_backward(c, s, rb, re) || launch_apply(c, re - rb);
            break;
        case LCTR_MODEL_NFM:
            if (fused_kernels_ok(c) && s.fused_valid) {
                // order-free embedding side (fm_fused.cu) around the dense layers; on several GPUs the rows come from the
                // batch-compact cache and the gradient rows go to their owners (dist.cu)
                const bool multi = c->cfg.world > 1;
                rc = (multi && dist_pre_step(c, s, slot, true)) || mlp_reserve(c, re - rb) || launch_nfm_forward_fused(c, s, rb, re) ||
                     launch_nfm_mlp(c, s, rb, re, re - rb) || launch_nfm_backward_fused(c, s, rb, re) ||
                     (multi ? dist_post_step(c, s, slot, re - rb) : launch_apply_compact(c, s, re - rb, nullptr, nullptr));
                break;
            }
            if (c->cfg.world > 1) {
                // embeddings: owner-sharded pull / push like FM; dense layers: replicated, gradients all-reduced
                rc = dist_pre_step(c, s, slot, false) || mlp_reserve(c, re - rb) || launch_fm_forward(c, s, rb, re, true, false) ||
                     launch_nfm_mlp(c, s, rb, re, re - rb) || launch_fm_backward(c, s, rb, re, true) ||
                     dist_post_step(c, s, slot, re - rb);
                break;
            }
            rc = mlp_reserve(c, re - rb) || launch_fm_forward(c, s, rb, re, true, false) ||
                 launch_nfm_mlp(c, s, rb, re, re - rb);
            if (!rc) {
                if (c->cfg.deterministic == 1) rc = launch_fm_backward_csc(c, s, rb, re, true);
                else rc = launch_fm_backward(c, s, rb, re, true) || launch_apply(c, re - rb);
            }
            break;
    }
    if (rc) return 1;
    c->step++;
    return read_stats(c, step, loss_sum, acc_cnt);
}

int lctr_train_batch(lctr_ctx* c, int64_t rows, int64_t nnz, const int64_t* row_ptr, const uint32_t* fid,
                     const uint16_t* field, const float* val, const int32_t* label, float* loss_sum, float* acc_cnt) {
    if (lctr_upload_batch(c, 0, rows, nnz, row_ptr, fid, field, val, label)) return 1;
    return lctr_train_step(c, 0, 0, rows, loss_sum, acc_cnt);
}

// ---- streamed training: copy stream + compute stream, two pipeline slots -------------------------------------
static int pipe_init(lctr_ctx* c) {
    if (c->copy_stream) return 0;
    LCTR_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    LCTR_CUDA(cudaStreamCreateWithFlags(&c->build_stream, cudaStreamNonBlocking));
    for (int i = 0; i < kPipe; i++) {
        LCTR_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
        LCTR_CUDA(cudaEventCreateWithFlags(&c->ev_computed[i], cudaEventDisableTiming));
        LCTR_CUDA(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < kStatRing; i++) LCTR_CUDA(cudaEventCreateWithFlags(&c->ev_stat[i], cudaEventDisableTiming));
    LCTR_CUDA(cudaMallocHost((void**)&c->h_stat_ring, sizeof(double) * 2 * kStatRing));
    return 0;
}

// Graph path of the streamed pipeline (FM, cfg.deterministic == 2, one GPU): per pipeline slot two captured graphs --
// `build` (slot header copy + the five grouping kernels, on the copy stream) and `step` (updater-parameter copy,
// forward, grouped backward + update, result copy, on the compute stream).  Kernels take the batch size from the
// device-side slot header and the updater parameters from device memory, so the graphs are static: a streamed step
// costs the host three cudaMemcpyAsync, two graph launches and four event calls instead of ~20 launches.
static int pipe_graph_capture(lctr_ctx* c, int p, bool has_val) {
    PipeGraph& g = c->pipe_graph[p];
    Slot& s = c->slots[kNumSlots - kPipe + p];
    if (g.build) { cudaGraphExecDestroy(g.build); g.build = nullptr; }
    if (g.step) { cudaGraphExecDestroy(g.step); g.step = nullptr; }
    if (!g.d_hdr) {
        LCTR_CUDA(cudaMalloc((void**)&g.d_hdr, 2 * sizeof(int64_t)));
        LCTR_CUDA(cudaMallocHost((void**)&g.h_hdr, 2 * sizeof(int64_t)));
        LCTR_CUDA(cudaMalloc((void**)&g.d_opt, csc_opt_params_size()));
        LCTR_CUDA(cudaMallocHost((void**)&g.h_opt, csc_opt_params_size()));
        LCTR_CUDA(cudaMalloc((void**)&g.d_stat, 2 * sizeof(double)));
        LCTR_CUDA(cudaMallocHost((void**)&g.h_stat, 2 * sizeof(double)));
    }
    s.has_val = has_val;
    {   // the launchers pick their updater instance from the host copy at capture time: give it the context's updater
        OptParams* hp = reinterpret_cast<OptParams*>(g.h_opt);
        memset(hp, 0, csc_opt_params_size());
        hp->opt = c->cfg.optimizer;
    }
    cudaGraph_t graph;
    const bool fused = fused_supported(c);
    // ---- build graph (copy stream)
    LCTR_CUDA(cudaStreamBeginCapture(c->copy_stream, cudaStreamCaptureModeThreadLocal));
    int rc = cudaMemcpyAsync(g.d_hdr, g.h_hdr, 2 * sizeof(int64_t), cudaMemcpyHostToDevice, c->copy_stream) != cudaSuccess;
    if (!rc) {
        if (fused) {  // labels widened, then the slot map of the batch (fm_fused.cu)
            label_to_float_kernel<<<(unsigned)((s.cap_rows + 255) / 256), 256, 0, c->copy_stream>>>(
                reinterpret_cast<int32_t*>(s.pred), s.label, s.cap_rows, g.d_hdr);
            c->launches++;
            rc = fused_build_slot(c, s, c->copy_stream, g.d_hdr, s.cap_rows, s.cap_nnz);
        } else {
            rc = csc_build_device(c, s, c->copy_stream, reinterpret_cast<int32_t*>(s.pred), g.d_hdr, s.cap_rows, s.cap_nnz);
        }
    }
    cudaError_t ce = cudaStreamEndCapture(c->copy_stream, &graph);  // always closes the capture, also on error paths
    if (rc || ce != cudaSuccess) { set_error("streamed pipeline: capture of the build graph failed (%s)", cudaGetErrorString(ce)); return 1; }
    LCTR_CUDA(cudaGraphInstantiate(&g.build, graph, 0));
    cudaGraphDestroy(graph);
    // ---- step graph (compute stream)
    LCTR_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    rc = cudaMemcpyAsync(g.d_opt, g.h_opt, csc_opt_params_size(), cudaMemcpyHostToDevice, c->stream) != cudaSuccess;
    if (!rc) {
        if (fused)
            rc = launch_fm_fused(c, s, 0, s.cap_rows, true, g.d_hdr, g.d_stat) ||
                 launch_apply_compact(c, s, s.cap_rows, reinterpret_cast<const OptParams*>(g.h_opt),
                                      reinterpret_cast<const OptParams*>(g.d_opt));
        else
            rc = launch_fm_forward_ex(c, s, 0, s.cap_rows, false, true, g.d_hdr, g.d_stat) ||
                 launch_fm_backward_devcsc_ex(c, s, 0, s.cap_rows, reinterpret_cast<const OptParams*>(g.h_opt), g.d_opt);
    }
    if (!rc) rc = cudaMemcpyAsync(g.h_stat, g.d_stat, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess;
    ce = cudaStreamEndCapture(c->stream, &graph);
    if (rc || ce != cudaSuccess) { set_error("streamed pipeline: capture of the step graph failed (%s)", cudaGetErrorString(ce)); return 1; }
    LCTR_CUDA(cudaGraphInstantiate(&g.step, graph, 0));
    cudaGraphDestroy(graph);
    g.cap_rows = s.cap_rows; g.cap_nnz = s.cap_nnz; g.has_val = has_val;
    return 0;
}

static int train_batch_async_graph(lctr_ctx* c, int64_t rows, int64_t nnz, const int64_t* row_ptr, const uint32_t* fid,
                                   const float* val, const int32_t* label, uint64_t* ticket) {
    const int p = (int)(c->pipe_issued % kPipe);
    const int slot = kNumSlots - kPipe + p;
    Slot& s = c->slots[slot];
    PipeGraph& g = c->pipe_graph[p];
    if (rows > s.cap_rows || nnz > s.cap_nnz || !g.build || g.has_val != (val != nullptr) || g.cap_rows != s.cap_rows ||
        g.cap_nnz != s.cap_nnz) {
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        LCTR_CUDA(cudaStreamSynchronize(c->copy_stream));
        LCTR_CUDA(cudaStreamSynchronize(c->build_stream));
        if (rows > s.cap_rows || nnz > s.cap_nnz)  // head-room so that slightly larger batches do not re-capture
            if (slot_reserve(c, s, std::max(rows, s.cap_rows) + rows / 8 + 64, std::max(nnz, s.cap_nnz) + nnz / 8 + 1024)) return 1;
        if (fused_supported(c) ? fused_reserve(c, s, s.cap_nnz) : csc_reserve(c, s, s.cap_nnz)) return 1;
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        if (pipe_graph_capture(c, p, val != nullptr)) return 1;
    }
    s.rows = rows; s.nnz = nnz; s.has_val = val != nullptr; s.has_field = false;
    if (c->pipe_issued >= (uint64_t)kPipe) LCTR_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_computed[p], 0));
    LCTR_CUDA(cudaMemcpyAsync(s.row_ptr, row_ptr, (size_t)(rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->copy_stream));
    LCTR_CUDA(cudaMemcpyAsync(s.fid, fid, (size_t)nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_stream));
    if (val) LCTR_CUDA(cudaMemcpyAsync(s.val, val, (size_t)nnz * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
    LCTR_CUDA(cudaMemcpyAsync(s.pred, label, (size_t)rows * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy_stream));
    g.h_hdr[0] = rows; g.h_hdr[1] = nnz;
    // the batch's slot map is built on its own stream: the copy engine moves batch t+1 while the SMs build batch t's map (the
    // build kernels share scratch buffers, so they stay serialised among themselves -- on this one stream)
    LCTR_CUDA(cudaEventRecord(c->ev_h2d[p], c->copy_stream));
    LCTR_CUDA(cudaStreamWaitEvent(c->build_stream, c->ev_h2d[p], 0));
    LCTR_CUDA(cudaGraphLaunch(g.build, c->build_stream));
    LCTR_CUDA(cudaEventRecord(c->ev_copied[p], c->build_stream));
    LCTR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[p], 0));
    const bool fused = fused_supported(c);
    if (fused) fused_opt_params(c, rows, g.h_opt); else csc_opt_params(c, rows, g.h_opt);
    LCTR_CUDA(cudaGraphLaunch(g.step, c->stream));
    LCTR_CUDA(cudaEventRecord(c->ev_computed[p], c->stream));
    s.dev_csc = !fused;
    s.fused_valid = fused;
    c->launches += 8;  // kernels inside the two graphs (6 + 2 on the fused path, 5 + 3 on the grouped one)
    g.ticket = c->step;
    *ticket = c->step++;
    c->pipe_issued++;
    return 0;
}

int lctr_train_batch_async(lctr_ctx* c, int64_t rows, int64_t nnz, const int64_t* row_ptr, const uint32_t* fid,
                           const uint16_t* field, const float* val, const int32_t* label, uint64_t* ticket) {
    LCTR_CHECK(c && ticket, "null argument");
    LCTR_CHECK(c->cfg.deterministic != 1, "streamed batches need cfg.deterministic 0 (RED scatter) or 2 (device grouping)");
    if (pipe_init(c)) return 1;
    LCTR_CHECK(c->pipe_issued - c->pipe_waited < (uint64_t)kPipe, "more than %d streamed batches outstanding: call lctr_wait first", kPipe);
    if ((c->cfg.deterministic == 2 || fused_supported(c)) && c->cfg.world == 1 && c->cfg.model == LCTR_MODEL_FM && !c->profiling &&
        rows > 0 && nnz > 0)
        return train_batch_async_graph(c, rows, nnz, row_ptr, fid, val, label, ticket);
    const int p = (int)(c->pipe_issued % kPipe);
    const int slot = kNumSlots - kPipe + p;
    // the copy may only overwrite the slot once the step that last used it has finished
    if (c->pipe_issued >= (uint64_t)kPipe) LCTR_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_computed[p], 0));
    if (c->pipe_issued >= 1) LCTR_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_copied[(p + kPipe - 1) % kPipe], 0));  // a graph-path build of the
                                                                   // other slot (build_stream) shares the slot-map scratch
    if (upload_batch_on(c, c->copy_stream, slot, rows, nnz, row_ptr, fid, field, val, label)) return 1;
    LCTR_CUDA(cudaEventRecord(c->ev_copied[p], c->copy_stream));
    LCTR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[p], 0));
    const uint64_t step = c->step;
    if (lctr_train_step(c, slot, 0, rows, nullptr, nullptr)) return 1;
    LCTR_CUDA(cudaEventRecord(c->ev_computed[p], c->stream));
    const int ri = (int)(step % kStatRing);
    LCTR_CUDA(cudaMemcpyAsync(c->h_stat_ring + 2 * ri, c->stats + 2 * ri, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LCTR_CUDA(cudaEventRecord(c->ev_stat[ri], c->stream));
    c->pipe_graph[p].ticket = ~0ull;
    *ticket = step;
    c->pipe_issued++;
    return 0;
}

int lctr_wait(lctr_ctx* c, uint64_t ticket, float* loss_sum, float* acc_cnt) {
    LCTR_CHECK(c && c->copy_stream, "lctr_wait: no streamed batch was issued");
    LCTR_CHECK(ticket < c->step && c->step - ticket <= (uint64_t)kStatRing, "lctr_wait: ticket %llu is not outstanding",
               (unsigned long long)ticket);
    for (int p = 0; p < kPipe; p++) {
        if (c->pipe_graph[p].build && c->pipe_graph[p].ticket == ticket) {
            LCTR_CUDA(cudaEventSynchronize(c->ev_computed[p]));
            if (loss_sum) *loss_sum = (float)c->pipe_graph[p].h_stat[0];
            if (acc_cnt) *acc_cnt = (float)c->pipe_graph[p].h_stat[1];
            if (c->pipe_waited < c->pipe_issued) c->pipe_waited++;
            return 0;
        }
    }
    const int ri = (int)(ticket % kStatRing);
    LCTR_CUDA(cudaEventSynchronize(c->ev_stat[ri]));
    if (loss_sum) *loss_sum = (float)c->h_stat_ring[2 * ri];
    if (acc_cnt) *acc_cnt = (float)c->h_stat_ring[2 * ri + 1];
    if (c->pipe_waited < c->pipe_issued) c->pipe_waited++;
    return 0;
}

int lctr_predict(lctr_ctx* c, int slot, int quirk_sumvx_slot, float* pctr) {
    LCTR_CHECK(c, "null ctx");
    LCTR_CHECK(slot >= 0 && slot < kNumSlots, "slot %d out of range", slot);
    Slot& s = c->slots[slot];
    int rc = 0;
    if (c->cfg.model == LCTR_MODEL_WND) {
        // Distributed_Algo_Abst::Predict (distributed_algo_abst.h:163-174): a forward pass over the slot; with several
        // ranks a collective call (every rank serves the rows its peers need)
        const float* out = nullptr;
        rc = (c->cfg.world > 1 && dist_pre_step(c, s, slot, false)) || mlp_reserve(c, s.rows) || wnd_reserve(c, s.rows) ||
             launch_wnd_forward(c, s, 0, s.rows) || mlp_forward_only(c, s.rows, &out) || launch_wnd_pred(c, s, out, 0, s.rows);
        if (rc) return 1;
        if (pctr) {
            LCTR_CUDA(cudaMemcpyAsync(pctr, s.pred, (size_t)s.rows * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            LCTR_CUDA(cudaStreamSynchronize(c->stream));
        }
        return 0;
    }
    // world > 1: the compute view only holds the rows pulled by the last train step, at their pre-update values
    LCTR_CHECK(c->cfg.world == 1, "lctr_predict: multi-GPU contexts keep sharded tables; download the parameters "
                                  "(lctr_download_params) into a single-GPU context to predict");
    if (c->cfg.model == LCTR_MODEL_FFM) {
        // parity mode: the reference's own pair loop order; otherwise the field-pair factorised forward
        rc = c->cfg.deterministic == 1 ? launch_ffm_predict_inorder(c, s) : launch_ffm_forward(c, s, 0, s.rows, false);
    } else if (c->cfg.model == LCTR_MODEL_FM) {
        if (quirk_sumvx_slot >= 0) {
            LCTR_CHECK(quirk_sumvx_slot < kNumSlots, "quirk slot out of range");
            rc = launch_predict_quirk(c, s, c->slots[quirk_sumvx_slot]);
        } else {
            // order-free contexts predict with the shuffle-tree forward; parity contexts with the in-order one
            rc = fused_supported(c) ? launch_fm_forward_tree(c, s, 0, s.rows, false) : launch_fm_forward(c, s, 0, s.rows, false, false);
        }
    } else {
        set_error("lctr_predict: the reference ships no NFM predictor (main.cpp:230-233)");
        return 1;
    }
    if (rc) return 1;
    if (pctr) {
        LCTR_CUDA(cudaMemcpyAsync(pctr, s.pred, (size_t)s.rows * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int lctr_download_sumvx(lctr_ctx* c, int slot, float* out) {
    LCTR_CHECK(c && out, "null argument");
    LCTR_CHECK(slot >= 0 && slot < kNumSlots, "slot %d out of range", slot);
    Slot& s = c->slots[slot];
    LCTR_CHECK(s.sumvx, "model has no sumVX (FFM keeps it NULL, fm_predict.cpp:20)");
    LCTR_CUDA(cudaMemcpyAsync(out, s.sumvx, (size_t)s.rows * c->cfg.factor_cnt * sizeof(float), cudaMemcpyDeviceToHost,
                              c->stream));
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int lctr_download_pred(lctr_ctx* c, int slot, float* out) {
    LCTR_CHECK(c && out, "null argument");
    LCTR_CHECK(slot >= 0 && slot < kNumSlots, "slot %d out of range", slot);
    Slot& s = c->slots[slot];
    LCTR_CUDA(cudaMemcpyAsync(out, s.pred, (size_t)s.rows * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int lctr_dense_grad_buffer(lctr_ctx* c, void** dev_ptr, size_t* n_floats) {
    LCTR_CHECK(c && dev_ptr && n_floats, "null argument");
    *dev_ptr = c->dense_grad;
    *n_floats = c->dense_grad_n;
    return 0;
}

int lctr_profile(lctr_ctx* c, int enable) {
    LCTR_CHECK(c, "null ctx");
    if (enable && !c->prof_ev) { c->prof_ev = new std::vector<cudaEvent_t>(); c->prof_id = new std::vector<int>(); }
    c->profiling = enable ? 1 : 0;
    return 0;
}
int lctr_profile_read(lctr_ctx* c, double* ms, int64_t* counts, int n, int reset) {
    LCTR_CHECK(c && ms && counts, "null argument");
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    if (c->prof_ev) {
        for (size_t i = 0; i < c->prof_id->size(); i++) {
            float t = 0.f;
            cudaEventElapsedTime(&t, (*c->prof_ev)[2 * i], (*c->prof_ev)[2 * i + 1]);
            const int id = (*c->prof_id)[i];
            if (id >= 0 && id < kNumProf) { c->prof_ms[id] += t; c->prof_cnt[id]++; }
            cudaEventDestroy((*c->prof_ev)[2 * i]); cudaEventDestroy((*c->prof_ev)[2 * i + 1]);
        }
        c->prof_ev->clear(); c->prof_id->clear();
    }
    for (int i = 0; i < n && i < kNumProf; i++) { ms[i] = c->prof_ms[i]; counts[i] = c->prof_cnt[i]; }
    if (reset) for (int i = 0; i < kNumProf; i++) { c->prof_ms[i] = 0; c->prof_cnt[i] = 0; }
    return 0;
}

int64_t lctr_launch_count(const lctr_ctx* c) { return c ? c->launches : 0; }
void* lctr_stream(lctr_ctx* c) { return c ? (void*)c->stream : nullptr; }

}  // extern "C"
