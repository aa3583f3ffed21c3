// lightctr_b200/csrc/opt.cu -- sparse per-coordinate updaters (Adagrad / FTRL / Adam), sm_100a.
//
// Reference: AdagradUpdater_Num::update (util/gradientUpdater.h:139-150), FTRLUpdater::update
// (:252-273), AdamUpdater_Num::update (util/momentumUpdater.h:187-210), applied by
// Train_*_Algo::ApplyGrad as a DENSE sweep over all F*(rowlen+1) coordinates.  Every updater
// skips coordinates whose batch gradient is exactly 0, and a gradient is only ever produced
// for the fids present in the batch (L2 terms are added inside the nnz loops), so visiting only
// the touched fids is exactly equivalent.  The backward kernels mark
// touched[fid] = 1 (plain byte store, idempotent); this kernel scans the byte map 16 B per lane,
// pops set entries 32 at a time and updates G = 32/LPR rows concurrently, LPR lanes per row,
// each lane owning VEC-float slices q, q+LPR, ...  It also zeroes the gradient and the mark, which
// replaces the reference's memset(grad) (gradientUpdater.h:149).
#include <algorithm>
#include <vector>

#include "opt.cuh"

namespace lctr {

template <int LPR, int VEC>
__global__ void __launch_bounds__(256)
apply_kernel(uint8_t* __restrict__ touched, size_t F, int rowlen, float* __restrict__ W, float* __restrict__ V,
             float* __restrict__ gW, float* __restrict__ gV, float* __restrict__ s1W, float* __restrict__ s1V,
             float* __restrict__ s2W, float* __restrict__ s2V, OptParams P) {
    constexpr int G = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int q = lane % LPR, g = lane / LPR;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    const size_t ntiles = (F + 511) / 512;  // 512 marks per warp-tile (16 B per lane)
    const bool two_state = P.opt != LCTR_OPT_ADAGRAD;
    for (size_t tile = warp; tile < ntiles; tile += nwarps) {
        const size_t base = tile * 512 + (size_t)lane * 16;
        uint4 m = make_uint4(0, 0, 0, 0);
        if (base + 16 <= F) {
            m = *reinterpret_cast<const uint4*>(touched + base);
        } else if (base < F) {
            unsigned char tmp[16];
            for (int i = 0; i < 16; i++) tmp[i] = base + i < F ? touched[base + i] : 0;
            m = *reinterpret_cast<uint4*>(tmp);
        }
        const bool any = (m.x | m.y | m.z | m.w) != 0;
        if (!__any_sync(kFull, any)) continue;
        // 16-bit mask of set marks of this lane
        unsigned bits = 0;
        {
            const unsigned wv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if ((wv[i] >> (8 * b)) & 0xffu) bits |= 1u << (i * 4 + b);
        }
        if (any) {  // clear the marks (memset(grad) analogue)
            if (base + 16 <= F) *reinterpret_cast<uint4*>(touched + base) = make_uint4(0, 0, 0, 0);
            else for (int i = 0; i < 16 && base + i < F; i++) touched[base + i] = 0;
        }
        while (__any_sync(kFull, bits != 0)) {
            // every lane pops one fid
            const bool has = bits != 0;
            const int bit = has ? __ffs(bits) - 1 : 0;
            if (has) bits &= bits - 1;
            const size_t my_f = base + bit;
            const unsigned havemask = __ballot_sync(kFull, has);
            // process the up-to-32 popped fids, G rows at a time
            for (int j = 0; j < 32; j += G) {
                if (((havemask >> j) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u))) == 0) continue;
                const int src = j + g;
                const size_t f = __shfl_sync(kFull, (unsigned long long)my_f, src);
                const bool ok = (havemask >> src) & 1u;
                if (!ok) continue;
                if (q == 0) {
                    float w = W[f], s1 = s1W[f], s2 = two_state ? s2W[f] : 0.f;
                    update_one(P, P.corrW, w, gW[f], s1, s2);
                    W[f] = w; s1W[f] = s1; gW[f] = 0.f;
                    if (two_state) s2W[f] = s2;
                }
                const size_t ro = f * (size_t)rowlen;
                for (int o = q * VEC; o < rowlen; o += LPR * VEC) {
                    if (VEC == 4) {
                        float4 gv = *reinterpret_cast<float4*>(gV + ro + o);
                        if (gv.x == 0.f && gv.y == 0.f && gv.z == 0.f && gv.w == 0.f) continue;  // untouched slice (FFM)
                        float4 wv = *reinterpret_cast<float4*>(V + ro + o);
                        float4 a = *reinterpret_cast<float4*>(s1V + ro + o);
                        float4 b2 = two_state ? *reinterpret_cast<float4*>(s2V + ro + o) : make_float4(0, 0, 0, 0);
                        update_one(P, P.corrV, wv.x, gv.x, a.x, b2.x);
                        update_one(P, P.corrV, wv.y, gv.y, a.y, b2.y);
                        update_one(P, P.corrV, wv.z, gv.z, a.z, b2.z);
                        update_one(P, P.corrV, wv.w, gv.w, a.w, b2.w);
                        *reinterpret_cast<float4*>(V + ro + o) = wv;
                        *reinterpret_cast<float4*>(s1V + ro + o) = a;
                        if (two_state) *reinterpret_cast<float4*>(s2V + ro + o) = b2;
                        *reinterpret_cast<float4*>(gV + ro + o) = make_float4(0, 0, 0, 0);
                    } else {
                        const float gv = gV[ro + o];
                        if (gv == 0.f) continue;
                        float wv = V[ro + o], a = s1V[ro + o], b2 = two_state ? s2V[ro + o] : 0.f;
                        update_one(P, P.corrV, wv, gv, a, b2);
                        V[ro + o] = wv; s1V[ro + o] = a; gV[ro + o] = 0.f;
                        if (two_state) s2V[ro + o] = b2;
                    }
                }
            }
        }
    }
}

int launch_apply(lctr_ctx* c, int64_t rows_in_step) {
    OptParams P = make_opt_params(c, rows_in_step);
    const int rowlen = (int)c->rowlen;
    int vec = (rowlen % 4 == 0) ? 4 : 1;
    int slices = rowlen / vec;
    int lpr = 1;
    while (lpr < slices && lpr < 32) lpr <<= 1;
    const size_t ntiles = (c->F + 511) / 512;
    unsigned grid = (unsigned)std::min<size_t>((ntiles + 7) / 8, (size_t)c->sm_count * 8);
    if (grid == 0) grid = 1;
#define APPLY_CASE(L, VV) \
    apply_kernel<L, VV><<<grid, 256, 0, c->stream>>>(c->touched, c->F, rowlen, c->W, c->V, c->gW, c->gV, c->s1W, \
                                                     c->s1V, c->s2W, c->s2V, P)
    if (vec == 4) {
        switch (lpr) {
            case 1: APPLY_CASE(1, 4); break;
            case 2: APPLY_CASE(2, 4); break;
            case 4: APPLY_CASE(4, 4); break;
            case 8: APPLY_CASE(8, 4); break;
            case 16: APPLY_CASE(16, 4); break;
            default: APPLY_CASE(32, 4); break;
        }
    } else {
        switch (lpr) {
            case 1: APPLY_CASE(1, 1); break;
            case 2: APPLY_CASE(2, 1); break;
            case 4: APPLY_CASE(4, 1); break;
            case 8: APPLY_CASE(8, 1); break;
            case 16: APPLY_CASE(16, 1); break;
            default: APPLY_CASE(32, 1); break;
        }
    }
#undef APPLY_CASE
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lctr
