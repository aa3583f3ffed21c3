/* include/lightctr_b200.h -- the drop-in boundary (C ABI) of the B200-native LightCTR hot path.
 *
 * The reference (cnkuangshi/LightCTR) has no FFI: its "API" is C++ inheritance
 * (FM_Algo_Abst::Train() fm_algo_abst.h:137, Train_FM_Algo/Train_FFM_Algo/Train_NFM_Algo ctors
 * train_fm_algo.h:23, train_ffm_algo.h:25, train_nfm_algo.h:21, Fully_Conn_Layer
 * train/layer/fullyconnLayer.h:80-206, updaters util/gradientUpdater.h:128-278,
 * util/momentumUpdater.h:172-215).  The host shims in lightctr_b200/host/ keep those class
 * surfaces; each of their methods lowers to the entry points below.  Plain pointers and sizes
 * only, no C++/torch types, no exceptions across the boundary.  Every function returns 0 on
 * success, non-zero on failure (lctr_last_error() gives the message; the host shims turn that
 * into the reference's own style: print + exit(1), fm_algo_abst.h:79-82).
 *
 * Threading: one ctx per trainer, driven from one host thread (the reference's Train() is
 * synchronous and single-caller).  All device work is issued on the ctx's
 * stream; calls that return host-visible results synchronise that stream only.
 */
#ifndef LIGHTCTR_B200_H
#define LIGHTCTR_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LCTR_ABI_VERSION 1

typedef struct lctr_ctx lctr_ctx;

/* LCTR_MODEL_WND: Wide&Deep with the per-field concat input of Distributed_Algo_Abst (distributed_algo_abst.h:176-280):
 * factor_cnt = the tensor width d (the reference's factor_dim 4), field_cnt > 0, hidden[] = the dense chain on the
 * field_cnt * d input; single GPU, RED scatter. */
enum { LCTR_MODEL_FM = 1, LCTR_MODEL_FFM = 2, LCTR_MODEL_NFM = 3, LCTR_MODEL_WND = 4 };
/* updater = the reference's `_Num` family member (gradientUpdater.h:128-154 Adagrad, :200-233 RMSprop,
 * :235-278 FTRL; momentumUpdater.h:74-111 Adadelta, :172-215 Adam) */
enum { LCTR_OPT_ADAGRAD = 0, LCTR_OPT_FTRL = 1, LCTR_OPT_ADAM = 2, LCTR_OPT_RMSPROP = 3, LCTR_OPT_ADADELTA = 4,
       /* the parameter server's own update rules (distribut/paramserver.h:232-300), applied by the owner of a row to the
        * step's summed gradient as ONE push of worker 0: SGD (:295-300, the PS default :49; Wide&Deep tensors use the tensor
        * form :232-237), Adagrad (:288-294), DCASGD / DCASGDA (:252-286; shadow copy and accumulator kept per coordinate).
        * The synchronous exchange has no staleness, so the delay-compensation term of DCASGD* sees the parameter change
        * since the previous step. */
       LCTR_OPT_PS_SGD = 5, LCTR_OPT_PS_ADAGRAD = 6, LCTR_OPT_PS_DCASGD = 7, LCTR_OPT_PS_DCASGDA = 8 };
enum { LCTR_ACT_SIGMOID = 0, LCTR_ACT_TANH = 1 };
enum { LCTR_MLP_FP32 = 0, LCTR_MLP_BF16 = 1 };

#define LCTR_MAX_LAYERS 8

/* Snapshot of the reference's process-global statics (main.cpp:64-73; SURVEY.md 8b "Hyper-parameter
 * sources") plus the trainer ctor arguments.  Zero-initialise, then fill. */
typedef struct lctr_cfg {
    uint32_t abi_version;     /* LCTR_ABI_VERSION */
    int32_t model;            /* LCTR_MODEL_* */
    int32_t optimizer;        /* LCTR_OPT_* for W and V (and the MLP, which the reference fixes to Adagrad) */
    int32_t device;           /* CUDA device ordinal */
    uint64_t feature_cnt;     /* F = max fid + 1 (fm_algo_abst.h:95) */
    uint32_t field_cnt;       /* Fc; 0 for FM/NFM (fm_algo_abst.h:57-59) */
    uint32_t factor_cnt;      /* k */
    float learning_rate;      /* GradientUpdater::__global_learning_rate */
    float l2_reg;             /* L2Reg_ratio (train_fm_algo.cpp:13) */
    uint64_t minibatch_size;  /* updater divisor; 0 => rows of the step (FM/FFM: train_fm_algo.cpp:38) */
    float momentum;           /* MomentumUpdater::__global_momentum       (Adam beta1 -- used for BOTH moments) */
    float momentum_adam2;     /* MomentumUpdater::__global_momentum_adam2 (Adam bias correction only) */
    float ftrl_alpha, ftrl_beta, ftrl_lambda1, ftrl_lambda2; /* gradientUpdater.h:275; 0 => reference defaults */
    /* MLP (NFM deep part): dims k -> hidden[0] -> ... -> hidden[n_hidden-1] -> 1 */
    int32_t n_hidden;
    uint32_t hidden[LCTR_MAX_LAYERS];
    int32_t activation;       /* LCTR_ACT_* of the hidden layers */
    int32_t mlp_precision;    /* LCTR_MLP_FP32 (parity) or LCTR_MLP_BF16 (tensor-core perf mode) */
    /* capacity hints (0 => grow on demand) */
    uint64_t max_rows, max_nnz;
    /* multi-GPU: this process' rank / world (1 process per GPU); tables are owner-sharded by fid % world */
    int32_t rank, world;
    /* Backward scatter-add strategy (FFM: 0 and 2 only):
     *  0  vector REDs into update_g + sparse apply (sum order arbitrary, like the reference's Hogwild threads);
     *  1  feature-major (CSC) view of the slot built on the HOST at upload: every gradient is summed in ascending
     *     row order -- the order of the reference's canonical single-thread run -- no atomics, updater fused.
     *     csc_row_block = rows per train_step range (0 => whole slot; NFM: the minibatch size);
     *  2  the same view built on the DEVICE at upload (overlaps the previous step); entries of a feature arrive in
     *     arbitrary order and are accumulated in double precision, so the fp32 result is order-independent.
     *     Whole-slot steps only; FM with k in {4, 8, 16, 32}, FFM with k % 4 == 0 and field_cnt * k <= 512. */
    int32_t deterministic;
    int32_t reserved0;
    uint64_t csc_row_block;
    float ema_rate;           /* GradientUpdater::__global_ema_rate (RMSpropUpdater_Num, gradientUpdater.h:200-233); 0 => 0.99 (main.cpp:66) */
    uint32_t reserved[3];
} lctr_cfg;

const char* lctr_last_error(void);
int lctr_abi_version(void);

/* ---- lifetime ------------------------------------------------------------------------------- */
int lctr_create(const lctr_cfg* cfg, lctr_ctx** out);
int lctr_destroy(lctr_ctx* ctx);
int lctr_sync(lctr_ctx* ctx);

/* ---- parameters (replaces direct pokes at FM_Algo_Abst::W / V, fm_algo_abst.h:141,145) ------- */
/* V layout == reference: FM/NFM V[fid*k + f] (fm_algo_abst.h:146-148); FFM V[fid*Fc*k + field*k + f] (:149-151) */
int lctr_upload_params(lctr_ctx* ctx, const float* W, const float* V);
int lctr_download_params(lctr_ctx* ctx, float* W, float* V);
/* synthetic initialisation on the device: W = 0, V ~ scale * N(0,1) (hash of the global element index; independent
 * of the sharding).  Used by bench.py for tables too large to stage through host memory. */
int lctr_fill_params(lctr_ctx* ctx, uint64_t seed, float scale);
/* optimizer state: s1 = adagrad accum | ftrl z | adam m ; s2 = ftrl n | adam v ; each F + |V| floats, W part first
 * (same concatenation as update_g, train_fm_algo.h:52-57).  NULL pointers are skipped. */
int lctr_download_opt_state(lctr_ctx* ctx, float* s1, float* s2);
int lctr_upload_opt_state(lctr_ctx* ctx, const float* s1, const float* s2);

/* ---- data: CSR form of FM_Algo_Abst::dataSet / label (fm_algo_abst.h:29-35,156,170) ---------- */
/* Copies one CSR batch (or a whole dataset) host->device into `slot` (0..3), asynchronously on the
 * ctx stream.  row_ptr has rows+1 entries; field may be NULL for FM/NFM; val may be NULL meaning all
 * 1.0f (the shipped data).  Indexing is bit-exact w.r.t. the reference parser. */
int lctr_upload_batch(lctr_ctx* ctx, int slot, int64_t rows, int64_t nnz, const int64_t* row_ptr,
                      const uint32_t* fid, const uint16_t* field, const float* val, const int32_t* label);

/* ---- the hot path --------------------------------------------------------------------------- */
/* One reference "batch": forward (gather + interaction [+ MLP]) -> loss -> backward scatter-add ->
 * per-coordinate update on the rows [row_begin,row_end) of the resident slot.
 *   FM : Train_FM_Algo::batchGradCompute + accumWVGrad + ApplyGrad   (train_fm_algo.cpp:63-126)
 *   FFM: Train_FFM_Algo::...                                         (train_ffm_algo.cpp:51-126)
 *   NFM: Train_NFM_Algo::batchGradCompute + ApplyGrad                (train_nfm_algo.cpp:56-169)
 * loss_sum / acc_cnt (may be NULL: then no host sync happens) receive the summed logloss and the
 * number of correct rows of this step, the quantities the reference prints per epoch. */
int lctr_train_step(lctr_ctx* ctx, int slot, int64_t row_begin, int64_t row_end, float* loss_sum, float* acc_cnt);
/* End-to-end convenience: upload_batch(slot 0) + train_step + result readback, host buffers in. */
int lctr_train_batch(lctr_ctx* ctx, int64_t rows, int64_t nnz, const int64_t* row_ptr, const uint32_t* fid,
                     const uint16_t* field, const float* val, const int32_t* label, float* loss_sum,
                     float* acc_cnt);
/* Streamed training (host buffers in, results out, every step) with the copy of batch i+1 overlapping the kernels of
 * batch i: the batch is copied on a second stream into one of LCTR_PIPE_DEPTH pipeline slots (its slot map is built on a
 * third stream), the step is enqueued behind it and its (loss, acc) are copied back into a pinned ring; returns immediately
 * with a ticket.  At most LCTR_PIPE_DEPTH tickets may be outstanding; the host buffers must stay valid until the ticket has
 * been waited for.  lctr_wait blocks until that step's results are on the host.  (Depth 3: while step t computes, batch
 * t+1 has its slot map built and batch t+2 is on the copy engine.) */
#define LCTR_PIPE_DEPTH 3
int lctr_train_batch_async(lctr_ctx* ctx, int64_t rows, int64_t nnz, const int64_t* row_ptr, const uint32_t* fid,
                           const uint16_t* field, const float* val, const int32_t* label, uint64_t* ticket);
int lctr_wait(lctr_ctx* ctx, uint64_t ticket, float* loss_sum, float* acc_cnt);
/* Forward only on a resident slot; pctr (rows floats, host) receives sigmoid(pred).
 * quirk_sumvx_slot >= 0 reproduces FM_Predict's use of the TRAINING sumVX of the same row index
 * (predict/fm_predict.cpp:27-32); -1 computes the proper FM prediction. */
int lctr_predict(lctr_ctx* ctx, int slot, int quirk_sumvx_slot, float* pctr);
/* sumVX[rows*k] of a slot as left by the last forward over it (FM_Algo_Abst::sumVX, fm_algo_abst.h:145). */
int lctr_download_sumvx(lctr_ctx* ctx, int slot, float* sumVX);
/* per-row sigmoid(pred) of the last forward over the slot */
int lctr_download_pred(lctr_ctx* ctx, int slot, float* pred);

/* ---- MLP (Fully_Conn_Layer chain, fullyconnLayer.h) ------------------------------------------ */
/* The chain as a stand-alone operator -- what a DL_Algo_Abst subclass calls per minibatch (dl_algo_abst.h:46-49:
 * Predict / BP / applyBP) and what Layer_Base::forward / backward / applyBatchGradient do per sample
 * (layer_abst.h:45-67, fullyconnLayer.h:80-206), batched over `rows` samples, fp32 reference-order arithmetic:
 *   forward   x [rows][in0] (in0 = factor_cnt, or field_cnt * factor_cnt for Wide&Deep) -> out [rows] = the last layer's
 *             linear output (:116); hidden activations stay on the device for the backward.  out may be NULL.
 *   backward  dout [rows] = outputDelta of the last layer; clips to +-15 (:129-131), accumulates weightDelta / biasDelta
 *             (:165-179) in the fused dense-gradient buffer, returns the first layer's inputDelta in dx [rows][in0]
 *             (NULL to skip the copy).  Must follow a forward of the same row count.
 *   apply     applyBatchGradient (:194-197): Adagrad on bias then weights with divisor `minibatch`, deltas zeroed.
 *             With world > 1 the registered dense all-reduce runs first.  (Dropout masks: lctr_mlp_set_mask.) */
int lctr_mlp_forward(lctr_ctx* ctx, int64_t rows, const float* x, float* out);
int lctr_mlp_backward(lctr_ctx* ctx, int64_t rows, const float* dout, float* dx);
int lctr_mlp_apply(lctr_ctx* ctx, uint64_t minibatch);
/* layer l: weight [out][in] row-major (fullyconnLayer.h:211-216), bias[out], dropout mask[out] (1/0). */
int lctr_mlp_upload(lctr_ctx* ctx, int layer, const float* weight, const float* bias);
int lctr_mlp_download(lctr_ctx* ctx, int layer, float* weight, float* bias);
int lctr_mlp_set_mask(lctr_ctx* ctx, int layer, const float* mask);
/* Fully_Conn_Layer::weightDelta / biasDelta (fullyconnLayer.h:165-179) as they stand in the fused dense-gradient
 * buffer; zero after a step unless the process runs with LCTR_MLP_SKIP_UPDATE=1 (test hook: gradients are left in
 * place and the Adagrad update of the dense layers is skipped). */
int lctr_mlp_download_grad(lctr_ctx* ctx, int layer, float* dweight, float* dbias);

/* ---- multi-GPU (replaces distribut/ring_collect.h + pull.h/push.h) ------------ */
/* Exchange of CUDA IPC handles is done by the caller's process group (torch.distributed / MPI):
 * export this rank's table handles, gather them, import all peers'. */
int lctr_ipc_export(lctr_ctx* ctx, void* handles_out, size_t cap, size_t* bytes);
int lctr_ipc_import(lctr_ctx* ctx, const void* all_handles, size_t bytes_per_rank);
/* device memory of the context in bytes: table shard + updater state, and (world > 1) the exchange arena, caches and
 * inboxes -- owner-sharding keeps the second number O(keys of a batch), not O(feature_cnt) */
int lctr_device_bytes(lctr_ctx* ctx, uint64_t* shard_bytes, uint64_t* exchange_bytes);
/* Data-parallel dense layers (world > 1, NFM): the per-rank weightDelta / biasDelta of the batch must be summed over
 * the ranks before the updater runs -- Worker_RingReduce::syncGradient (distribut/ring_collect.h:48-72) on the
 * BufferFusion of Fully_Conn_Layer::registerGradient (fullyconnLayer.h:69-75).  The library calls `fn` once per train
 * step, between the MLP backward and the dense Adagrad, with the fused gradient buffer and the CUDA stream the step
 * runs on; `fn` must enqueue an in-place SUM all-reduce on that stream (ncclAllReduce(buf, buf, n, ncclFloat, ncclSum,
 * comm, stream), or torch.distributed.all_reduce under that stream) and return 0.  Required when world > 1. */
typedef int (*lctr_allreduce_fn)(void* user, float* dev_buf, size_t n_floats, void* cuda_stream);
int lctr_set_dense_allreduce(lctr_ctx* ctx, lctr_allreduce_fn fn, void* user);
/* device pointer + element count of the fused dense-gradient buffer (the BufferFusion of
 * fullyconnLayer.h:69-75) for an external ncclAllReduce; 0 elements when the model has no MLP */
int lctr_dense_grad_buffer(lctr_ctx* ctx, void** dev_ptr, size_t* n_floats);

/* ---- host-side ingest (fm_algo_abst.h:70-107), bit-exact indexing ----------------------------- */
typedef struct lctr_dataset {
    int64_t rows, nnz, label_cnt;
    uint64_t feature_cnt, field_cnt;
    int64_t* row_ptr;
    uint32_t* fid;
    uint16_t* field;
    float* val;
    int32_t* label;
} lctr_dataset;
int lctr_load_libffm(const char* path, uint64_t field_cnt_in, uint64_t feature_cnt_in, lctr_dataset** out);
int lctr_free_dataset(lctr_dataset* d);
/* binary CSR cache of a parsed file: the sscanf-per-token parse is paid once */
int lctr_save_dataset_bin(const lctr_dataset* d, const char* path);
int lctr_load_dataset_bin(const char* path, lctr_dataset** out);

/* ---- test-set metrics on the device ----------------------------------------- */
/* Summed logloss and correct count of FM_Predict::Predict (predict/fm_predict.cpp:63-72) and AucEvaluator's AUC
 * (util/evaluator.h:51-104, 2^24 - 1 buckets, fp32 trapezoid walk from the top bucket) over the pCTR / label arrays
 * that lctr_predict left in `slot`.  The AUC is bit-identical to the reference's for the same pCTR values. */
int lctr_eval(lctr_ctx* ctx, int slot, float* loss_sum, int64_t* correct, float* auc);
/* test hook: overwrite the slot's pCTR array with host values */
int lctr_upload_pred(lctr_ctx* ctx, int slot, const float* pctr);

/* ---- checkpoint / resume ---------------------------------------------------- */
/* FM_Algo_Abst::saveModel (fm_algo_abst.h:109-135) writes W and V as text (kept in the host shims); these dump and
 * restore the complete trainer state in binary -- W, V, updater state (Adagrad accumulators | FTRL z, n | Adam m, v and
 * its call counter), dense layers with their Adagrad state and dropout masks, step counter -- so that a restored
 * trainer continues exactly where the saved one stopped.  The restoring ctx must have been created with the same cfg. */
int lctr_save_checkpoint(lctr_ctx* ctx, const char* path);
int lctr_load_checkpoint(lctr_ctx* ctx, const char* path);

/* per-kernel device timing for bench.py's roofline: when enabled, every launch on the ctx stream is bracketed by
 * CUDA events; lctr_profile_read sums the elapsed ms and launch counts per kernel class
 * (0 fm_forward, 1 fm_backward(RED), 2 sparse apply, 3 ffm_fused, 4 fm_backward(CSC)+update, 5 mlp). */
int lctr_profile(lctr_ctx* ctx, int enable);
int lctr_profile_read(lctr_ctx* ctx, double* ms, int64_t* counts, int n, int reset);

/* introspection for tests / bench: number of kernels this library has launched on ctx so far */
int64_t lctr_launch_count(const lctr_ctx* ctx);
/* raw CUDA stream handle (cudaStream_t) of the ctx, for event timing on the launching stream */
void* lctr_stream(lctr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* LIGHTCTR_B200_H */
