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
#include "opt.cuh"

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
// forward -- arithmetic in the reference's own order
// ------------------------------------------------------------------------------------------------
// avx_dotProduct(x, x, k) of common/avx.h:102-127 evaluated across the LR lanes of one sample group:
// lane c holds p = x_c*x_c.  Full 8-chunks are accumulated lane-wise (d_l = p_l + p_{l+8} + ...), then
// the hsum tree (d_i + d_{i+4}; a_0 + a_2, a_1 + a_3; b_0 + b_1), then the scalar tail in order.
// Result valid in lane c == 0 of the group.  All 32 lanes must call this.
template <int LR>
__device__ __forceinline__ float avx_dot_lanes(float p, int k, int c) {
    const int nfull = k >> 3;
    float result = 0.f;
    if (nfull > 0) {
        float d = p;
        if (LR > 8) {
            for (int m = 1; m < nfull; m++) {
                const float o = __shfl_down_sync(kFull, p, 8 * m, LR);
                d = d + o;
            }
        }
        float a = d + __shfl_down_sync(kFull, d, 4, LR);
        float b = a + __shfl_down_sync(kFull, a, 2, LR);
        result = b + __shfl_down_sync(kFull, b, 1, LR);
    }
    for (int t = nfull * 8; t < k; t++) {
        const float o = __shfl_sync(kFull, p, t, LR);
        result = result + o;
    }
    (void)c;
    return result;
}

// One sample per group of LR lanes (LR = next power of two >= k); lane c of the group owns factor c and
// walks the sample's features IN ORDER, so sumVX[c], the bi-interaction z[c] and -- on lane 0 -- the
// float/double fm_pred chain of train_fm_algo.cpp:69-84 are evaluated exactly as the reference does.
// The LR gathers of a chunk are all issued before the first is consumed.
template <int LR, bool HAS_VAL, bool NFM>
__global__ void __launch_bounds__(256)
fm_forward_kernel(const int64_t* __restrict__ row_ptr, const uint32_t* __restrict__ fid,
                  const float* __restrict__ val, const float* __restrict__ label, const float* __restrict__ W,
                  const float* __restrict__ V, int k, float* __restrict__ pred, float* __restrict__ sumvx,
                  float* __restrict__ z_out, float* __restrict__ wide_out, int64_t rb, int64_t re, double* partial,
                  unsigned int* done, double* out_slot, int do_stats) {
    constexpr int RPW = 32 / LR;
    const int lane = threadIdx.x & 31;
    const int c = lane % LR, grp = lane / LR;
    const int64_t r = rb + ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + grp;
    const bool rv = r < re, act = c < k;
    const int64_t b = rv ? row_ptr[r] : 0;
    const int n = rv ? (int)(row_ptr[r + 1] - b) : 0;
    int nmax = n;
#pragma unroll
    for (int o = LR; o < 32; o <<= 1) nmax = max(nmax, __shfl_xor_sync(kFull, nmax, o));
    float s = 0.f, z = 0.f, fm = 0.f;
    for (int base = 0; base < nmax; base += LR) {
        const int cnt = n - base;  // entries of this chunk that exist for my group (may be <= 0)
        const uint32_t my_f = c < cnt ? __ldg(fid + b + base + c) : 0u;
        const float my_x = HAS_VAL ? (c < cnt ? __ldg(val + b + base + c) : 0.f) : 1.f;
        float v[LR], w[LR];
#pragma unroll
        for (int j = 0; j < LR; j++) {
            const uint32_t f = __shfl_sync(kFull, my_f, j, LR);
            const bool ok = j < cnt;
            v[j] = (ok && act) ? __ldg(V + (size_t)f * k + c) : 0.f;
            w[j] = (ok && c == 0) ? __ldg(W + f) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < LR; j++) {
            if (j >= nmax - base) break;  // warp-uniform
            const bool ok = j < cnt;
            const float x = __shfl_sync(kFull, my_x, j, LR);
            const float t = v[j] * x;                    // avx_vecScale(V, tmp, X)      train_fm_algo.cpp:76
            s = s + t;                                   // sumVX += tmp                 :77
            if (NFM) {
                z = z + t * (t * -0.5f);                 // train_nfm_algo.cpp:87-91
                if (ok) fm = fm + w[j] * x;              // wide part                    train_nfm_algo.cpp:83
            } else {
                const float dot = avx_dot_lanes<LR>(t * t, k, c);                      // :78
                if (ok) {
                    fm = fm + w[j] * x;                                                // fm_pred += W[fid] * X  :74
                    fm = (float)((double)fm - 0.5 * (double)dot);                      // fm_pred -= 0.5 * dot   :78
                }
            }
        }
    }
    double loss = 0.0, correct = 0.0;
    if (rv && act) sumvx[(size_t)r * k + c] = s;
    if (NFM) {
        // z = z + sumVX * (sumVX * 0.5)    (train_nfm_algo.cpp:93-94)
        if (rv && act) z_out[(size_t)(r - rb) * k + c] = z + s * (s * 0.5f);
        if (rv && c == 0) wide_out[r] = fm;
    } else {
        const float dot = avx_dot_lanes<LR>(act ? s * s : 0.f, k, c);
        if (rv && c == 0) {
            fm = (float)((double)fm + 0.5 * (double)dot);  // :82
            const float p = ref_sigmoid(fm);               // :84
            pred[r] = p;
            if (do_stats) loss_terms(p, label[r], loss, correct);
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

template <int LR>
static void fwd_go(lctr_ctx* c, Slot& s, bool nfm, unsigned grid, int k, int64_t rb, int64_t re, double* out_slot,
                   int stats) {
#define FWD_ARGS s.row_ptr, s.fid, s.val, s.label, c->W, c->V, k, s.pred, s.sumvx, c->z, s.wide, rb, re, \
                 c->stat_partial, c->stat_done, out_slot, stats
    if (s.has_val) {
        if (nfm) fm_forward_kernel<LR, true, true><<<grid, 256, 0, c->stream>>>(FWD_ARGS);
        else fm_forward_kernel<LR, true, false><<<grid, 256, 0, c->stream>>>(FWD_ARGS);
    } else {
        if (nfm) fm_forward_kernel<LR, false, true><<<grid, 256, 0, c->stream>>>(FWD_ARGS);
        else fm_forward_kernel<LR, false, false><<<grid, 256, 0, c->stream>>>(FWD_ARGS);
    }
#undef FWD_ARGS
}

int launch_fm_forward(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool nfm, bool stats) {
    const int k = (int)c->cfg.factor_cnt;
    LCTR_CHECK(k <= 32, "factor_cnt=%d unsupported by the in-order forward (need k <= 32)", k);
    const int64_t rows = re - rb;
    if (rows <= 0) return 0;
    int lr = 4;
    while (lr < k) lr <<= 1;
    const int rows_per_cta = 8 * (32 / lr);
    const unsigned grid = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
    double* out_slot = c->stats + 2 * (c->step % kStatRing);
    switch (lr) {
        case 4: fwd_go<4>(c, s, nfm, grid, k, rb, re, out_slot, stats ? 1 : 0); break;
        case 8: fwd_go<8>(c, s, nfm, grid, k, rb, re, out_slot, stats ? 1 : 0); break;
        case 16: fwd_go<16>(c, s, nfm, grid, k, rb, re, out_slot, stats ? 1 : 0); break;
        default: fwd_go<32>(c, s, nfm, grid, k, rb, re, out_slot, stats ? 1 : 0); break;
    }
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


// ------------------------------------------------------------------------------------------------
// deterministic backward + fused updater over the feature-major (CSC) view
// ------------------------------------------------------------------------------------------------
// One group of LR lanes per segment (= one fid of the row block); lane c owns factor c.  The fid's
// entries are visited in ascending row order and accumulated with the reference's exact expression
// sequence (train_fm_algo.cpp:108-115 / train_nfm_algo.cpp:131-157), starting from the zeroed update_g,
// so the summed gradient equals the reference's canonical single-thread result bit for bit (given equal
// inputs).  The updater then runs on the register-resident gradient: no update_g traffic, no atomics,
// no touched map.
template <int LR, bool HAS_VAL, bool NFM>
__global__ void __launch_bounds__(256)
fm_backward_csc_kernel(const int64_t* __restrict__ seg_ptr, const uint32_t* __restrict__ seg_fid,
                       const uint32_t* __restrict__ ent_row, const float* __restrict__ ent_x, int64_t seg_begin,
                       int64_t seg_end, const float* __restrict__ label, const float* __restrict__ pred,
                       const float* __restrict__ sumvx, const float* __restrict__ dz, int64_t rb, float* __restrict__ W,
                       float* __restrict__ V, float* __restrict__ s1W, float* __restrict__ s1V, float* __restrict__ s2W,
                       float* __restrict__ s2V, int k, float l2, OptParams P) {
    constexpr int SPW = 32 / LR;
    const int lane = threadIdx.x & 31;
    const int c = lane % LR, grp = lane / LR;
    const int64_t seg = seg_begin + ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * SPW + grp;
    const bool sv = seg < seg_end, act = c < k;
    const int64_t eb = sv ? seg_ptr[seg] : 0;
    const int n = sv ? (int)(seg_ptr[seg + 1] - eb) : 0;
    int nmax = n;
#pragma unroll
    for (int o = LR; o < 32; o <<= 1) nmax = max(nmax, __shfl_xor_sync(kFull, nmax, o));
    const uint32_t f = sv ? seg_fid[seg] : 0u;
    const float w = sv ? W[f] : 0.f;
    const float v = (sv && act) ? V[(size_t)f * k + c] : 0.f;
    float u = 0.f, gwsum = 0.f;
    for (int base = 0; base < nmax; base += LR) {
        const int cnt = n - base;
        const uint32_t my_row = c < cnt ? __ldg(ent_row + eb + base + c) : 0u;
        const float my_x = HAS_VAL ? (c < cnt ? __ldg(ent_x + eb + base + c) : 0.f) : 1.f;
        float sv_[LR], d_[LR], dz_[LR];
#pragma unroll
        for (int j = 0; j < LR; j++) {
            const uint32_t row = __shfl_sync(kFull, my_row, j, LR);
            const bool ok = j < cnt;
            sv_[j] = (ok && act) ? __ldg(sumvx + (size_t)row * k + c) : 0.f;
            d_[j] = ok ? (__ldg(pred + row) - __ldg(label + row)) : 0.f;  // LogisticGradW  fm_algo_abst.h:159-161
            dz_[j] = (NFM && ok && act) ? __ldg(dz + (size_t)(row - rb) * k + c) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < LR; j++) {
            if (j >= nmax - base) break;  // warp-uniform
            const float x = __shfl_sync(kFull, my_x, j, LR);
            if (j < cnt) {
                const float gw = d_[j] * x + l2 * w;         // train_fm_algo.cpp:108 / train_nfm_algo.cpp:135
                gwsum = gwsum + gw;                          // *update_W(fid) += gradW            :109
                const float t = sv_[j] + v * (-x);           // avx_vecScalerAdd(sumVX, V, tmp, -x) :112-113
                if (NFM) u = u + t * (dz_[j] * x);           // train_nfm_algo.cpp:154-156
                else u = u + t * gw;                         // avx_vecScalerAdd(ptr, tmp, ptr, gradW) :114
                u = u + v * l2;                              // avx_vecScalerAdd(ptr, V, ptr, L2)      :115
            }
        }
    }
    if (!sv) return;
    const bool two = P.opt != LCTR_OPT_ADAGRAD;
    if (c == 0) {
        float ww = w, a = s1W[f], b2 = two ? s2W[f] : 0.f;
        update_one(P, P.corrW, ww, gwsum, a, b2);
        W[f] = ww; s1W[f] = a;
        if (two) s2W[f] = b2;
    }
    if (act) {
        const size_t o = (size_t)f * k + c;
        float vv = v, a = s1V[o], b2 = two ? s2V[o] : 0.f;
        update_one(P, P.corrV, vv, u, a, b2);
        V[o] = vv; s1V[o] = a;
        if (two) s2V[o] = b2;
    }
}

template <int LR>
static void bwd_csc_go(lctr_ctx* c, Slot& s, bool nfm, unsigned grid, int k, int64_t sb, int64_t se, int64_t rb,
                       const OptParams& P) {
#define CSC_ARGS s.seg_ptr, s.seg_fid, s.ent_row, s.ent_x, sb, se, s.label, s.pred, s.sumvx, c->dz, rb, c->W, c->V, \
                 c->s1W, c->s1V, c->s2W, c->s2V, k, c->cfg.l2_reg, P
    if (s.has_val) {
        if (nfm) fm_backward_csc_kernel<LR, true, true><<<grid, 256, 0, c->stream>>>(CSC_ARGS);
        else fm_backward_csc_kernel<LR, true, false><<<grid, 256, 0, c->stream>>>(CSC_ARGS);
    } else {
        if (nfm) fm_backward_csc_kernel<LR, false, true><<<grid, 256, 0, c->stream>>>(CSC_ARGS);
        else fm_backward_csc_kernel<LR, false, false><<<grid, 256, 0, c->stream>>>(CSC_ARGS);
    }
#undef CSC_ARGS
}

int launch_fm_backward_csc(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, bool nfm) {
    const int k = (int)c->cfg.factor_cnt;
    LCTR_CHECK(k <= 32, "factor_cnt=%d unsupported by the feature-major backward (need k <= 32)", k);
    LCTR_CHECK(s.csc_block > 0 && s.h_blk_seg_ptr, "deterministic step on a slot uploaded without the CSC view");
    LCTR_CHECK(rb % s.csc_block == 0 && (re == rb + s.csc_block || re == s.rows) && re - rb <= s.csc_block,
               "deterministic train_step rows [%lld,%lld) do not match the slot's row blocks of %lld",
               (long long)rb, (long long)re, (long long)s.csc_block);
    const int64_t bi = rb / s.csc_block;
    const int64_t sb = (*s.h_blk_seg_ptr)[bi], se = (*s.h_blk_seg_ptr)[bi + 1];
    if (se <= sb) return 0;
    const OptParams P = make_opt_params(c, re - rb);
    int lr = 4;
    while (lr < k) lr <<= 1;
    const int segs_per_cta = 8 * (32 / lr);
    const unsigned grid = (unsigned)((se - sb + segs_per_cta - 1) / segs_per_cta);
    switch (lr) {
        case 4: bwd_csc_go<4>(c, s, nfm, grid, k, sb, se, rb, P); break;
        case 8: bwd_csc_go<8>(c, s, nfm, grid, k, sb, se, rb, P); break;
        case 16: bwd_csc_go<16>(c, s, nfm, grid, k, sb, se, rb, P); break;
        default: bwd_csc_go<32>(c, s, nfm, grid, k, sb, se, rb, P); break;
    }
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
