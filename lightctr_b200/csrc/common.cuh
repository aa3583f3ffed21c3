// lightctr_b200/csrc/common.cuh -- shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/lightctr_b200.h"
#include "ref_expf.h"

namespace lctr {

void set_error(const char* fmt, ...);

#define LCTR_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::lctr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

#define LCTR_CHECK(cond, ...)                  \
    do {                                       \
        if (!(cond)) {                         \
            ::lctr::set_error(__VA_ARGS__);    \
            return 1;                          \
        }                                      \
    } while (0)

constexpr int kNumSlots = 8;
constexpr int kPipe = LCTR_PIPE_DEPTH;  // streamed pipeline: batches in flight (the last kPipe slots are its buffers)
constexpr int kNumProf = 16;  // per-kernel timing buckets
enum { PROF_FM_FWD = 0, PROF_FM_BWD_RED = 1, PROF_APPLY = 2, PROF_FFM_FUSED = 3, PROF_FM_BWD_CSC = 4, PROF_MLP = 5, PROF_DIST_MARK = 6, PROF_DIST_COMPACT = 7, PROF_DIST_PULL = 8, PROF_DIST_PUSH = 9, PROF_DIST_BAR0 = 10, PROF_DIST_MERGE = 11, PROF_DIST_BAR1 = 12, PROF_CSC_BUILD = 13, PROF_FM_FUSED = 14, PROF_APPLY_COMPACT = 15 };
constexpr int kStatRing = 64;
constexpr int kHotRep = 32;      // fm_fused: replica rows per hot slot of the batch-compact gradient buffer
constexpr int kHotMax = 2048;    // hot slots per batch (ids beyond the cap stay ordinary slots)
constexpr uint32_t kHotBit = 0x80000000u;  // gradient index of an entry of a hot slot: kHotBit | replica block
constexpr unsigned kFull = 0xffffffffu;

// One resident CSR batch / dataset (FM_Algo_Abst::dataSet + label, fm_algo_abst.h:156,170).
struct Slot {
    int64_t rows = 0, nnz = 0, cap_rows = 0, cap_nnz = 0;
    int64_t* row_ptr = nullptr;  // rows+1
    uint32_t* fid = nullptr;     // nnz
    uint16_t* field = nullptr;   // nnz (FFM)
    float* val = nullptr;        // nnz, or unused when !has_val (all 1.0f)
    float* label = nullptr;      // rows, as float (the reference compares `float target`)
    float* pred = nullptr;       // rows: sigmoid(pred) of the last forward
    float* sumvx = nullptr;      // rows*k: FM_Algo_Abst::sumVX (fm_algo_abst.h:145)
    float* wide = nullptr;       // rows (NFM: wide part)
    bool has_val = false, has_field = false;
    // feature-major view (built on the host at upload when cfg.deterministic): per row block, the segments
    // (one per distinct fid of the block) list that fid's entries in ascending row order
    int64_t csc_block = 0;          // rows per block (0: none built)
    int64_t n_blocks = 0, n_segs = 0;
    int64_t* blk_seg_ptr = nullptr; // n_blocks+1 (host copy in h_blk_seg_ptr)
    int64_t* seg_ptr = nullptr;     // n_segs+1 offsets into ent_*
    uint32_t* seg_fid = nullptr;    // n_segs
    uint32_t* ent_row = nullptr;    // nnz: row index of the entry
    float* ent_x = nullptr;         // nnz (only when has_val)
    uint16_t* ent_field = nullptr;  // nnz (FFM, device-built view only)
    std::vector<int64_t>* h_blk_seg_ptr = nullptr;
    int64_t cap_segs = 0, cap_blocks = 0, cap_ent = 0;
    // device-built feature-major view (cfg.deterministic == 2): work lists of short / long segments and
    // totals[4] = {nnz, n_segs, n_short, n_long} (device)
    bool dev_csc = false;
    uint32_t *short_list = nullptr, *long_list = nullptr;  // long_list holds uint2 {segment, first entry} tasks
    int64_t cap_long = 0;
    double* csc_acc = nullptr;            // meeting point of multi-task segments
    unsigned int* csc_arrived = nullptr;
    unsigned int* csc_totals = nullptr;
    // multi-GPU: unique fids of the whole slot (the pull/push key set), built once at upload on the upload stream
    uint32_t* uniq = nullptr;
    unsigned int* n_uniq = nullptr;
    int64_t cap_uniq = 0;
    bool uniq_valid = false;
    // order-free fused FM step (fm_fused.cu): per-entry slot of the gradient row (or kHotBit | replica block), the hot
    // slots' replica-block index per slot and their list; `uniq` / `n_uniq` above hold the key set
    uint32_t* ent_slot = nullptr;
    uint32_t* ent_pslot = nullptr;  // multi-GPU: plain slot per entry = row of the batch-compact parameter cache
    int64_t cap_ent_slot = 0;
    uint32_t *hot_of = nullptr, *hot_slot = nullptr;
    unsigned int* n_hot = nullptr;
    bool fused_valid = false;
};

// captured graphs of one slot of the streamed pipeline (capi.cu)
struct PipeGraph {
    cudaGraphExec_t build = nullptr, step = nullptr;
    int64_t cap_rows = 0, cap_nnz = 0;
    bool has_val = false;
    int64_t *d_hdr = nullptr, *h_hdr = nullptr;  // {rows, nnz} of the batch in the slot
    void *d_opt = nullptr, *h_opt = nullptr;     // updater parameters of the step
    double *d_stat = nullptr, *h_stat = nullptr; // (loss, correct)
    uint64_t ticket = ~0ull;
};

struct MlpLayer {
    int in = 0, out = 0;
    float *w = nullptr, *b = nullptr, *mask = nullptr;  // [out][in], [out], [out]
    float *dw = nullptr, *db = nullptr;                 // views into the fused dense-grad buffer
    float *acc_w = nullptr, *acc_b = nullptr;           // Adagrad state
    float* act = nullptr;                               // [B][out] activations (post-activation for hidden layers)
    float* delta = nullptr;                             // [B][out] dL/d(pre-activation)
    void* w16 = nullptr;                                // bf16 copy of w (tensor-core mode, hidden layers)
    void* w16t = nullptr;                               // the same, as the chunk-major tile mlp_umma.cu stages by bulk copy
};

}  // namespace lctr

namespace lctr {
struct DistState;
struct OptParams;
// slot map scratch + batch-compact gradient buffers of the order-free fused FM step (fm_fused.cu); the slot map part is
// also what the multi-GPU exchange is keyed by (dist.cu)
struct FusedState {
    uint8_t* mark = nullptr;      // 128 * T permuted byte marks
    size_t T = 0;
    uint32_t* slot_of = nullptr;  // F: fid -> slot of the batch being built
    unsigned int* cnt = nullptr;  // sampled multiplicities (zero between builds)
    size_t cnt_cap = 0;
    float* G = nullptr;           // [G_rows][GS] compact gradient rows (zero between steps)
    size_t G_rows = 0;
    float* Ghot = nullptr;        // [kHotMax][kHotRep][GS] replica rows of the hot slots (zero between steps)
    OptParams* d_opt = nullptr;   // updater parameters in device memory (graph launches)
    int GS = 0;                   // 0: slot map only (models without the fused kernels)
};
}  // namespace lctr
struct lctr_ctx {
    lctr_cfg cfg;
    cudaStream_t stream = nullptr;
    size_t F = 0, rowlen = 0;  // rowlen = k (FM/NFM) or Fc*k (FFM)
    size_t Fl = 0;             // rows of this rank's table shard (== F when world == 1)
    // parameters, gradient accumulators (update_g layout: W part, V part), optimizer state
    float *W = nullptr, *V = nullptr, *gW = nullptr, *gV = nullptr;
    float *s1W = nullptr, *s1V = nullptr, *s2W = nullptr, *s2V = nullptr;
    uint8_t* touched = nullptr;  // Fl bytes: 1 = fid (shard-local index) received gradient this step
    // compute view used by the forward/backward kernels, indexed by GLOBAL fid: aliases of the arrays above when
    // world == 1; in multi-GPU mode a full-size local cache of the rows pulled this step + local gradient buffers
    float *cW = nullptr, *cV = nullptr, *cgW = nullptr, *cgV = nullptr;
    lctr::DistState* dist = nullptr;
    lctr::FusedState* fused = nullptr;  // order-free fused FM step (fm_fused.cu)
    size_t dist_rows = 0;               // world > 1: rows of the exchange index space (gradient / cache rows are indexed by it)
    uint32_t* touch_list = nullptr;      // compacted fids of the step (stage A of the sparse apply)
    unsigned int* n_touch = nullptr;     // list length (device)
    unsigned int* apply_done = nullptr;  // block-completion counter of stage B
    // per-step statistics ring: [kStatRing][2] doubles (loss sum, correct count) + scratch
    double* stats = nullptr;
    double* stat_partial = nullptr;     // [2] running accumulation of the current step
    unsigned int* stat_done = nullptr;  // block-completion counter
    double* h_stats = nullptr;          // pinned host mirror [2]
    uint64_t step = 0;
    size_t adam_iter = 0;
    lctr::Slot slots[lctr::kNumSlots];
    // MLP
    int n_layers = 0;
    lctr::MlpLayer layers[LCTR_MAX_LAYERS + 1];
    float* dense_grad = nullptr;  // fused [dW0, db0, dW1, db1, ...]
    size_t dense_grad_n = 0;
    float *z = nullptr, *dz = nullptr;  // [B][k] NFM bi-interaction output and its gradient
    float* mlp_out = nullptr;           // [B]
    size_t mlp_cap_rows = 0;
    int64_t mlp_fwd_rows = 0;      // rows of the last lctr_mlp_forward (lctr_mlp_backward must match)
    uint32_t* wnd_src = nullptr;   // Wide&Deep: fid of the first entry of each field, [rows][Fc]
    size_t wnd_cap_rows = 0;
    void* auc_scratch = nullptr;  // metrics.cu: histograms + lists of lctr_eval
    int csc_in_step = 0;        // LCTR_CSC_IN_STEP=1: rebuild the feature-major view inside every train step (bench)
    float* ffm_T = nullptr;     // FFM grouped step: per-sample field-pair tiles [rows][Fc][Fc][k]
    uint16_t* ffm_cnt = nullptr; // [rows][Fc] features per field
    size_t ffm_T_rows = 0;
    lctr_allreduce_fn dense_allreduce = nullptr;  // world > 1: sums dense_grad over the ranks on c->stream
    void* dense_allreduce_user = nullptr;
    int mlp_tm = 0;             // bf16 mode: samples per CTA tile (128 or 64)
    size_t mlp_smem = 0;        // bf16 mode: dynamic shared memory per CTA
    int mlp_umma = 0;           // bf16 mode: the tcgen05 kernel (mlp_umma.cu) takes this chain
    size_t mlp_umma_smem = 0;
    int mlp_has_mask = 0;       // any dropout mask entry == 0
    int mlp_skip_update = 0;    // LCTR_MLP_SKIP_UPDATE=1: leave the dense gradients in place (tests read them)
    int sm_count = 148;
    int64_t launches = 0;
    const unsigned long long* apply_wait_flags = nullptr;  // multi-GPU owner: flags the sparse apply polls before it starts
    int apply_wait_n = 0;
    unsigned long long apply_wait_epoch = 0;
    const float* fwd_quirk_sumvx = nullptr;  // FM_Predict quirk: training sumVX rows used by the next forward launch
    int64_t fwd_quirk_rows = 0;
    void* csc_scratch = nullptr;  // csc.cu: dense count / offset arrays of the device-side grouping
    // optional per-kernel timing (lctr_profile): events bracket every launch on the ctx stream
    int profiling = 0;
    std::vector<cudaEvent_t>* prof_ev = nullptr;   // flat list of (start, stop) pairs
    std::vector<int>* prof_id = nullptr;
    double prof_ms[lctr::kNumProf] = {0};
    int64_t prof_cnt[lctr::kNumProf] = {0};
    // streamed training pipeline (lctr_train_batch_async): copy stream, per-slot events, pinned result ring
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[lctr::kPipe] = {}, ev_computed[lctr::kPipe] = {};
    cudaStream_t build_stream = nullptr;                  // graph pipeline: the slot-map kernels of batch t run here while the copy
    cudaEvent_t ev_h2d[lctr::kPipe] = {};                   // engine already moves batch t+1 on copy_stream
    cudaEvent_t ev_stat[lctr::kStatRing] = {nullptr};
    double* h_stat_ring = nullptr;
    uint64_t pipe_issued = 0, pipe_waited = 0;
    lctr::PipeGraph pipe_graph[lctr::kPipe];
};

namespace lctr {

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
// vectorised no-return atomic add: one 16-byte RED per 4 floats (sm_90+: red.global.add.v4.f32)
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// Gather loads whose ISSUE ORDER matters: as volatile asm they stay where the source puts them (all gathers of a pass
// back to back), where plain __ldg loads were sunk next to their uses -- one dependent round trip per row instead of
// one per pass (seen in the SASS of the forward kernel, profiles/README.md).
__device__ __forceinline__ float4 ldg_f4_pinned(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_f32_pinned(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// avx_dotProduct(x, y, n) (common/avx.h:102-127) with strided operands, evaluated by ONE thread in the
// reference's order: 8 lane accumulators over the full 8-chunks, the hsum tree, then the scalar tail.
template <typename FX, typename FY>
__device__ __forceinline__ float avx_dot_seq(FX x, FY y, int n) {
    float result = 0.f;
    int i = 0;
    if (n > 7) {
        float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (; i + 8 <= n; i += 8) {
#pragma unroll
            for (int l = 0; l < 8; l++) d[l] = d[l] + x(i + l) * y(i + l);
        }
        const float a0 = d[4] + d[0], a1 = d[5] + d[1], a2 = d[6] + d[2], a3 = d[7] + d[3];
        const float b0 = a0 + a2, b1 = a1 + a3;
        result = result + (b0 + b1);
    }
    for (; i < n; i++) result = result + x(i) * y(i);
    return result;
}

// Sigmoid::forward, util/activations.h:65-72 (clamps at +-16; accurate expf, no fast-math)
__device__ __forceinline__ float ref_sigmoid(float x) {
    if (x < -16.f) return 1e-7f;
    if (x > 16.f) return 0.99999988f;  // (float)(1.0 - 1e-7)
    return 1.0f / (1.0f + lctr_ref_expf(-x));
}
// std::exp(float) of the reference (glibc expf) for unbounded arguments (Tanh, activations.h:134)
__device__ __forceinline__ float ref_exp_any(float x) { return fabsf(x) < 87.f ? lctr_ref_expf(x) : expf(x); }
// loss term + accuracy, train_fm_algo.cpp:93-98: y==1 ? -logf(p) : -log(1.0 - p) (double)
__device__ __forceinline__ void loss_terms(float p, float y, double& loss, double& correct) {
    loss = (y == 1.f) ? (double)(-logf(p)) : -log(1.0 - (double)p);
    correct = ((p > 0.5f && y == 1.f) || (p < 0.5f && y == 0.f)) ? 1.0 : 0.0;
}

// Block-level accumulation of (loss, correct) into ctx statistics; the last block to finish publishes
// the step totals into the ring slot and re-arms the accumulators (no memset launches between steps).
__device__ __forceinline__ void publish_stats(double loss, double correct, double* partial, unsigned int* done,
                                               double* out_slot, bool accumulate_out) {
    __shared__ double sh[2][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    loss = warp_sum_d(loss);
    correct = warp_sum_d(correct);
    if (lane == 0) { sh[0][wid] = loss; sh[1][wid] = correct; }
    __syncthreads();
    if (wid == 0) {
        double a = lane < nw ? sh[0][lane] : 0.0, b = lane < nw ? sh[1][lane] : 0.0;
        a = warp_sum_d(a);
        b = warp_sum_d(b);
        if (lane == 0) {
            atomicAdd(&partial[0], a);
            atomicAdd(&partial[1], b);
            __threadfence();
            unsigned int prev = atomicAdd(done, 1u);
            if (prev == gridDim.x - 1) {
                __threadfence();
                double l = atomicAdd(&partial[0], 0.0), c = atomicAdd(&partial[1], 0.0);
                if (accumulate_out) { out_slot[0] += l; out_slot[1] += c; }
                else { out_slot[0] = l; out_slot[1] = c; }
                partial[0] = 0.0; partial[1] = 0.0;
                *done = 0u;
                __threadfence();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// kernel launchers (defined in the .cu files)
// ---------------------------------------------------------------------------------------------
// RAII bracket: records a (start, stop) event pair around one launch when profiling is on
struct ProfScope {
    lctr_ctx* c; int id; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(lctr_ctx* c_, int id_) : c(c_), id(id_) {
        if (!c->profiling) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, c->stream);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEventRecord(b, c->stream);
        c->prof_ev->push_back(a); c->prof_ev->push_back(b); c->prof_id->push_back(id);
    }
};
int launch_fm_forward(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool nfm, bool stats);
int launch_fm_backward(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool nfm);
int launch_ffm_forward(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool stats);
int launch_ffm_backward(lctr_ctx* c, Slot& s, int64_t rb, int64_t re);
int launch_apply(lctr_ctx* c, int64_t rows_in_step);
int launch_fm_backward_csc(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool nfm);
// csc.cu: feature-major view built on the device at upload + atomic-free backward with fused updater
int csc_reserve(lctr_ctx* c, Slot& s, int64_t max_nnz);
int csc_build_device(lctr_ctx* c, Slot& s, cudaStream_t st, const int32_t* label_i32, const int64_t* hdr,
                     int64_t rows_cap, int64_t nnz_cap);
int launch_fm_backward_devcsc(lctr_ctx* c, Slot& s, int64_t rb, int64_t re);
struct OptParams;
int launch_fm_backward_devcsc_ex(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, const OptParams* P_host, const void* dP);
void csc_opt_params(lctr_ctx* c, int64_t rows, void* out);
size_t csc_opt_params_size();
int launch_fm_forward_ex(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool nfm, bool stats, const int64_t* hdr,
                         double* out_slot_override);
bool csc_device_supported(const lctr_ctx* c);
void csc_scratch_free(lctr_ctx* c);
int launch_predict_quirk(lctr_ctx* c, Slot& s, Slot& train);
// fm_fused.cu: order-free FM step over a batch-compact gradient buffer (cfg.deterministic == 0, one GPU)
bool fused_supported(const lctr_ctx* c);
bool fused_kernels_ok(const lctr_ctx* c);
void fused_free(lctr_ctx* c);
int fused_reserve(lctr_ctx* c, Slot& s, int64_t nnz);
int fused_build_slot(lctr_ctx* c, Slot& s, cudaStream_t st, const int64_t* hdr, int64_t rows_cap, int64_t nnz_cap);
void dist_wait_info(lctr_ctx* c, const unsigned long long** flags, int* n, unsigned long long* epoch);
int launch_fm_fused(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool stats, const int64_t* hdr, double* out_slot_override);
int launch_fm_forward_tree(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool stats);
int launch_nfm_forward_fused(lctr_ctx* c, Slot& s, int64_t rb, int64_t re);
int launch_nfm_backward_fused(lctr_ctx* c, Slot& s, int64_t rb, int64_t re);
int launch_apply_compact(lctr_ctx* c, Slot& s, int64_t rows_in_step, const OptParams* P_host, const OptParams* dP);
void fused_opt_params(lctr_ctx* c, int64_t rows, void* out);
void* fused_dev_opt(lctr_ctx* c);
int launch_ffm_predict_inorder(lctr_ctx* c, Slot& s);
// multi-GPU (dist.cu)
int dist_alloc(lctr_ctx* c);
int dist_free(lctr_ctx* c);
int dist_send_keys(lctr_ctx* c, Slot& s, int slot, cudaStream_t st);               // at upload: key lists -> the owners' inboxes
int dist_pre_step(lctr_ctx* c, Slot& s, int slot, bool in_kernel_wait);           // owner-driven pull of the step's rows
int dist_post_step(lctr_ctx* c, Slot& s, int slot, int64_t rows_divisor);         // push gradients, owner-side merge + update
int dist_check_overflow(lctr_ctx* c);
size_t dist_bytes(const lctr_ctx* c);
__global__ void compact_touched_kernel(uint8_t* touched, size_t F, uint32_t* list, unsigned int* n_list,
                                       const unsigned long long* wait_flags, int n_wait, unsigned long long wait_epoch);
int mlp_alloc(lctr_ctx* c);
int mlp_free(lctr_ctx* c);
int mlp_reserve(lctr_ctx* c, int64_t rows);
bool ffm_grouped_supported(const lctr_ctx* c);
int ffm_grouped_reserve(lctr_ctx* c, int64_t rows);
void ffm_grouped_free(lctr_ctx* c);
int launch_ffm_forward_tiles(lctr_ctx* c, Slot& s, int64_t rb, int64_t re);
int launch_ffm_backward_grouped(lctr_ctx* c, Slot& s, int64_t rb, int64_t re);
void metrics_free(lctr_ctx* c);
int wnd_reserve(lctr_ctx* c, int64_t rows);
void wnd_free(lctr_ctx* c);
int launch_wnd_forward(lctr_ctx* c, Slot& s, int64_t rb, int64_t re);
int launch_wnd_backward(lctr_ctx* c, Slot& s, int64_t rb, int64_t re);
int launch_wnd_pred(lctr_ctx* c, Slot& s, const float* mlp_out, int64_t rb, int64_t re);
int mlp_forward_only(lctr_ctx* c, int64_t rows, const float** out);
// width of the dense chain's input: k (NFM bi-interaction) or Fc * d (Wide&Deep concat)
inline size_t mlp_in0(const lctr_cfg& cf) {
    return cf.model == LCTR_MODEL_WND ? (size_t)cf.field_cnt * cf.factor_cnt : (size_t)cf.factor_cnt;
}
int mlp_sync_dense_grad(lctr_ctx* c);
bool pdl_on();  // LCTR_PDL != 0 (fm_fused.cu)
// scan + clear of a permuted byte map (fm_fused.cuh): appends the ids of the set positions to uniq, counts in *n_uniq
void launch_slotmap_compact(lctr_ctx* c, uint8_t* mark, size_t T, uint32_t* uniq, unsigned int* n_uniq, cudaStream_t st);
int launch_ffm_warp(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool stats);  // 0 launched, -1 shape not covered, 1 error
int mlp_bf16_prepare(lctr_ctx* c);
int mlp_bf16_refresh(lctr_ctx* c, int layer);
int launch_nfm_mlp_bf16(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, int64_t rows_divisor);
int launch_nfm_mlp(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, int64_t rows_divisor);

}  // namespace lctr
