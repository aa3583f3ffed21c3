// lightctr_b200/csrc/fm.cu -- FM / NFM embedding gather (forward) and scatter-add (backward), sm_100a.
//
// Reference semantics: Train_FM_Algo::batchGradCompute / accumWVGrad (train/train_fm_algo.cpp:63-118)
// and the wide + bi-interaction / accumWideGrad / accumDeepGrad parts of Train_NFM_Algo
// (train/train_nfm_algo.cpp:78-94,126-159).
//
// Mapping: one warp per sample.  A V row (k floats) is covered by LPR lanes, each holding VEC
// contiguous floats (k = 16 -> 4 lanes x float4: one 64 B row = two 32 B sectors, fully used), so a
// warp gathers G = 32/LPR rows per step.  The CSR columns of the sample are read 32 at a time
// with one coalesced 128 B load and redistributed by shuffle.  All gathers of a 32-entry chunk are
// issued before any is consumed (LPR independent 16 B loads in flight per lane).  Interaction sums
// are reduced with xor-shuffles.  HBM/L2-bound integer+fp32 work: no tensor cores here by design.
#include "common.cuh"

namespace lctr {

template <int VEC>
struct Vec {
    float a[VEC];
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> load_row(const float* p, bool active) {
    Vec<VEC> r;
    if (VEC == 4) {
        float4 t = active ? ldg_f4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
        r.a[0] = t.x; r.a[1 % VEC] = t.y; r.a[2 % VEC] = t.z; r.a[3 % VEC] = t.w;
    } else {
        r.a[0] = active ? __ldg(p) : 0.f;
    }
    return r;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int LPR, int VEC, bool HAS_VAL, bool NFM>
__global__ void __launch_bounds__(256)
fm_forward_kernel(const int64_t* __restrict__ row_ptr, const uint32_t* __restrict__ fid,
                  const float* __restrict__ val, const float* __restrict__ label, const float* __restrict__ W,
                  const float* __restrict__ V, int k, float* __restrict__ pred, float* __restrict__ sumvx,
                  float* __restrict__ z_out, float* __restrict__ wide_out, int64_t rb, int64_t re, double* partial,
                  unsigned int* done, double* out_slot, int do_stats) {
    constexpr int G = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int q = lane % LPR, g = lane / LPR;
    const int64_t r = rb + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const bool qa = q * VEC < k;  // lane covers real factors
    double loss = 0.0, correct = 0.0;
    if (r < re) {
        const int64_t b = row_ptr[r], e = row_ptr[r + 1];
        float s[VEC], zq[VEC];
#pragma unroll
        for (int c = 0; c < VEC; c++) { s[c] = 0.f; zq[c] = 0.f; }
        float sq = 0.f, wsum = 0.f;
        for (int64_t base = b; base < e; base += 32) {
            const int cnt = (int)min((int64_t)32, e - base);
            const uint32_t my_f = lane < cnt ? __ldg(fid + base + lane) : 0u;
            const float my_x = HAS_VAL ? (lane < cnt ? __ldg(val + base + lane) : 0.f) : 1.f;
            Vec<VEC> v[LPR];
            float w[LPR], x[LPR];
            // issue every gather of this chunk first
#pragma unroll
            for (int j = 0; j < LPR; j++) {
                const int idx = j * G + g;
                const uint32_t f = __shfl_sync(kFull, my_f, idx);
                x[j] = __shfl_sync(kFull, my_x, idx);
                const bool ok = idx < cnt;
                v[j] = load_row<VEC>(V + (size_t)f * k + q * VEC, ok && qa);
                w[j] = (ok && q == 0) ? __ldg(W + f) : 0.f;
                if (!ok) x[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < LPR; j++) {
                wsum += w[j] * x[j];  // fm_pred += W[fid] * X           (train_fm_algo.cpp:74)
#pragma unroll
                for (int c = 0; c < VEC; c++) {
                    const float t = v[j].a[c] * x[j];  // avx_vecScale(V, tmp, X) (:76)
                    s[c] += t;                         // sumVX += tmp           (:77)
                    if (NFM) zq[c] += t * (t * -0.5f); // train_nfm_algo.cpp:87-91
                    else sq += t * t;                  // dot(tmp,tmp)           (:78)
                }
            }
        }
        // reduce over the G lane-groups that share the same factor slice
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
            for (int c = 0; c < VEC; c++) {
                s[c] += __shfl_xor_sync(kFull, s[c], o);
                if (NFM) zq[c] += __shfl_xor_sync(kFull, zq[c], o);
            }
        }
        wsum = warp_sum(wsum);
        if (g == 0 && qa) {
#pragma unroll
            for (int c = 0; c < VEC; c++) sumvx[(size_t)r * k + q * VEC + c] = s[c];
        }
        if (NFM) {
            // z = sum_i t*(-0.5 t) + s*(0.5 s)   (train_nfm_algo.cpp:93-94)
            if (g == 0 && qa) {
#pragma unroll
                for (int c = 0; c < VEC; c++) z_out[(size_t)(r - rb) * k + q * VEC + c] = zq[c] + s[c] * (s[c] * 0.5f);
            }
            if (lane == 0) wide_out[r] = wsum;
        } else {
            sq = warp_sum(sq);
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < VEC; c++) ss += s[c] * s[c];
#pragma unroll
            for (int o = 1; o < LPR; o <<= 1) ss += __shfl_xor_sync(kFull, ss, o);
            // fm_pred = sum w x - 0.5 sum|vx|^2 + 0.5 |sum vx|^2  (:74,78,82; the 0.5* terms are double there)
            const float fm_pred = (float)((double)wsum + 0.5 * ((double)ss - (double)sq));
            const float p = ref_sigmoid(fm_pred);
            if (lane == 0) {
                pred[r] = p;
                if (do_stats) loss_terms(p, label[r], loss, correct);
            }
        }
    }
    if (!NFM && do_stats) publish_stats(loss, correct, partial, done, out_slot, false);
}

// ------------------------------------------------------------------------------------------------
// backward: scatter-add into the dense update_g (W part, V part) with vector REDs + touched marks
// ------------------------------------------------------------------------------------------------
template <int LPR, int VEC, bool HAS_VAL, bool NFM>
__global__ void __launch_bounds__(256)
fm_backward_kernel(const int64_t* __restrict__ row_ptr, const uint32_t* __restrict__ fid,
                   const float* __restrict__ val, const float* __restrict__ label, const float* __restrict__ W,
                   const float* __restrict__ V, int k, const float* __restrict__ pred,
                   const float* __restrict__ sumvx, const float* __restrict__ dz, float* __restrict__ gW,
                   float* __restrict__ gV, uint8_t* __restrict__ touched, float l2, int64_t rb, int64_t re) {
    constexpr int G = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int q = lane % LPR, g = lane / LPR;
    const int64_t r = rb + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= re) return;
    const bool qa = q * VEC < k;
    const int64_t b = row_ptr[r], e = row_ptr[r + 1];
    const float d = pred[r] - label[r];  // LogisticGradW: (pred - label) * x  (fm_algo_abst.h:159-161)
    float s[VEC], dzv[VEC];
#pragma unroll
    for (int c = 0; c < VEC; c++) {
        s[c] = qa ? sumvx[(size_t)r * k + q * VEC + c] : 0.f;
        dzv[c] = (NFM && qa) ? dz[(size_t)(r - rb) * k + q * VEC + c] : 0.f;
    }
    for (int64_t base = b; base < e; base += 32) {
        const int cnt = (int)min((int64_t)32, e - base);
        const uint32_t my_f = lane < cnt ? __ldg(fid + base + lane) : 0u;
        const float my_x = HAS_VAL ? (lane < cnt ? __ldg(val + base + lane) : 0.f) : 1.f;
        Vec<VEC> v[LPR];
        float w[LPR], x[LPR];
        uint32_t ff[LPR];
#pragma unroll
        for (int j = 0; j < LPR; j++) {
            const int idx = j * G + g;
            ff[j] = __shfl_sync(kFull, my_f, idx);
            x[j] = __shfl_sync(kFull, my_x, idx);
            const bool ok = idx < cnt;
            v[j] = load_row<VEC>(V + (size_t)ff[j] * k + q * VEC, ok && qa);
            w[j] = ok ? __ldg(W + ff[j]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < LPR; j++) {
            const int idx = j * G + g;
            if (idx >= cnt) continue;
            const float gw = d * x[j] + l2 * w[j];  // train_fm_algo.cpp:108 / train_nfm_algo.cpp:135
            float gv[VEC];
#pragma unroll
            for (int c = 0; c < VEC; c++) {
                const float t = s[c] + v[j].a[c] * (-x[j]);  // sumVX - x*V  (:112-113 / nfm :152-153)
                if (NFM) gv[c] = t * (dzv[c] * x[j]) + v[j].a[c] * l2;  // train_nfm_algo.cpp:154-157
                else gv[c] = t * gw + v[j].a[c] * l2;                   // train_fm_algo.cpp:114-115
            }
            if (qa) {
                float* dst = gV + (size_t)ff[j] * k + q * VEC;
                if (VEC == 4) red_add_v4(dst, make_float4(gv[0], gv[1 % VEC], gv[2 % VEC], gv[3 % VEC]));
                else red_add_f32(dst, gv[0]);
            }
            if (q == 0) {
                red_add_f32(gW + ff[j], gw);  // *update_W(fid) += gradW  (:109)
                touched[ff[j]] = 1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// FM_Predict quirk mode (predict/fm_predict.cpp:20-33): pred = sum w x - 0.5 sum|vx|^2 + 0.5|sumVX_train[rid]|^2
// ------------------------------------------------------------------------------------------------
__global__ void fm_predict_quirk_kernel(const int64_t* __restrict__ row_ptr, const uint32_t* __restrict__ fid,
                                        const float* __restrict__ val, const float* __restrict__ W,
                                        const float* __restrict__ V, int k, const float* __restrict__ train_sumvx,
                                        int64_t train_rows, float* __restrict__ pred, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float wsum = 0.f, sq = 0.f;
    for (int64_t i = row_ptr[r] + lane; i < row_ptr[r + 1]; i += 32) {
        const uint32_t f = fid[i];
        const float x = val ? val[i] : 1.f;
        wsum += W[f] * x;
        for (int c = 0; c < k; c++) { const float t = V[(size_t)f * k + c] * x; sq += t * t; }
    }
    wsum = warp_sum(wsum);
    sq = warp_sum(sq);
    float ss = 0.f;
    if (r < train_rows)
        for (int c = lane; c < k; c += 32) { const float t = train_sumvx[(size_t)r * k + c]; ss += t * t; }
    ss = warp_sum(ss);
    if (lane == 0) pred[r] = ref_sigmoid((float)((double)wsum + 0.5 * ((double)ss - (double)sq)));
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
struct Shape { int lpr, vec; };
static bool pick_shape(int k, Shape& sh) {
    if (k % 4 == 0 && k <= 128 && (k / 4 & (k / 4 - 1)) == 0) { sh.vec = 4; sh.lpr = k / 4; return true; }
    if (k <= 32) { int l = 1; while (l < k) l <<= 1; sh.vec = 1; sh.lpr = l; return true; }
    return false;
}

#define FM_DISPATCH_LPR(KERNEL, VECN, HV, NF, ...)                                     \
    switch (sh.lpr) {                                                                  \
        case 1: KERNEL<1, VECN, HV, NF><<<grid, 256, 0, c->stream>>>(__VA_ARGS__); break;   \
        case 2: KERNEL<2, VECN, HV, NF><<<grid, 256, 0, c->stream>>>(__VA_ARGS__); break;   \
        case 4: KERNEL<4, VECN, HV, NF><<<grid, 256, 0, c->stream>>>(__VA_ARGS__); break;   \
        case 8: KERNEL<8, VECN, HV, NF><<<grid, 256, 0, c->stream>>>(__VA_ARGS__); break;   \
        case 16: KERNEL<16, VECN, HV, NF><<<grid, 256, 0, c->stream>>>(__VA_ARGS__); break; \
        case 32: KERNEL<32, VECN, HV, NF><<<grid, 256, 0, c->stream>>>(__VA_ARGS__); break; \
    }
#define FM_DISPATCH(KERNEL, ...)                                                        \
    do {                                                                                \
        if (sh.vec == 4) {                                                              \
            if (s.has_val) { if (nfm) { FM_DISPATCH_LPR(KERNEL, 4, true, true, __VA_ARGS__) } else { FM_DISPATCH_LPR(KERNEL, 4, true, false, __VA_ARGS__) } } \
            else { if (nfm) { FM_DISPATCH_LPR(KERNEL, 4, false, true, __VA_ARGS__) } else { FM_DISPATCH_LPR(KERNEL, 4, false, false, __VA_ARGS__) } } \
        } else {                                                                        \
            if (s.has_val) { if (nfm) { FM_DISPATCH_LPR(KERNEL, 1, true, true, __VA_ARGS__) } else { FM_DISPATCH_LPR(KERNEL, 1, true, false, __VA_ARGS__) } } \
            else { if (nfm) { FM_DISPATCH_LPR(KERNEL, 1, false, true, __VA_ARGS__) } else { FM_DISPATCH_LPR(KERNEL, 1, false, false, __VA_ARGS__) } } \
        }                                                                               \
    } while (0)

int launch_fm_forward(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool nfm, bool stats) {
    Shape sh;
    const int k = (int)c->cfg.factor_cnt;
    LCTR_CHECK(pick_shape(k, sh), "factor_cnt=%d unsupported (need k<=32, or k in {4,8,16,32,64,128})", k);
    const int64_t rows = re - rb;
    if (rows <= 0) return 0;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    double* out_slot = c->stats + 2 * (c->step % kStatRing);
    FM_DISPATCH(fm_forward_kernel, s.row_ptr, s.fid, s.val, s.label, c->W, c->V, k, s.pred, s.sumvx, c->z, s.wide,
                rb, re, c->stat_partial, c->stat_done, out_slot, stats ? 1 : 0);
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

int launch_fm_backward(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool nfm) {
    Shape sh;
    const int k = (int)c->cfg.factor_cnt;
    LCTR_CHECK(pick_shape(k, sh), "factor_cnt=%d unsupported", k);
    const int64_t rows = re - rb;
    if (rows <= 0) return 0;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    FM_DISPATCH(fm_backward_kernel, s.row_ptr, s.fid, s.val, s.label, c->W, c->V, k, s.pred, s.sumvx, c->dz, c->gW,
                c->gV, c->touched, c->cfg.l2_reg, rb, re);
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

int launch_predict_quirk(lctr_ctx* c, Slot& s, Slot& train) {
    if (s.rows <= 0) return 0;
    const unsigned grid = (unsigned)((s.rows + 7) / 8);
    fm_predict_quirk_kernel<<<grid, 256, 0, c->stream>>>(s.row_ptr, s.fid, s.has_val ? s.val : nullptr, c->W, c->V,
                                                         (int)c->cfg.factor_cnt, train.sumvx, train.rows, s.pred,
                                                         s.rows);
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lctr
