// lightctr_b200/csrc/mlp_umma.cu -- the NFM dense layers on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// tensor memory).  Same Fully_Conn_Layer chain and the same bf16 rounding points as mlp_bf16.cu (train/layer/
// fullyconnLayer.h:80-180; operands bf16, accumulation fp32, masters fp32), different machine mapping:
//
//   * one CTA = 128 samples = the M of every UMMA; 8 warps; all of tensor memory (512 columns) belongs to the CTA;
//   * every matrix that is ever an MMA operand lives in shared memory as an un-swizzled "chunk-major" tile
//         byte offset of element (r, c) of an R x C matrix = (c / 8) * (R * 16) + r * 16 + (c % 8) * 2
//     (8 x 16 B core matrices; 8-row groups 128 B apart, 8-column chunks R*16 B apart).  The tile is a K-major operand
//     when its columns are the reduction index and an MN-major operand when its rows are, so ONE copy of the activations
//     X_l [sample][feature], of the deltas (written over X_{l+1} in place) and of the weights W_l [out][in] serves all
//     three products of a layer -- only LBO/SBO and the major bits of the descriptors differ:
//         forward  Y  = X_l . W_l^T          A = X_l     K-major    B = W_l     K-major    K = in_l
//         dX       dX = delta_l . W_l        A = delta_l K-major    B = W_l     MN-major   K = out_l
//         dW       dW = delta_l^T . X_l      A = delta_l MN-major   B = X_l     MN-major   K = 128 samples
//     (dW is computed transposed, X_l^T . delta_l, when out_l is not a multiple of 128 but in_l is);
//   * weights arrive by three bulk-copy (TMA) transfers from chunk-major bf16 copies kept in global memory next to the
//     fp32 masters (mlp_bf16.cu maintains them in the dense Adagrad kernel);
//   * one elected thread issues the MMAs of a phase and commits them to an mbarrier; the epilogues read the accumulators
//     with tcgen05.ld (thread = one TMEM lane = one sample for forward/dX, = one weight row for dW), apply bias /
//     activation / activation' / clipping and write the next operand straight back into its chunk-major tile (16 B per
//     thread, conflict free), or send dW to the dense-gradient buffer with vector REDs;
//   * db_l (column sums of delta_l over the samples) is taken where delta_l is produced, by a 31-shuffle
//     transpose-reduce per 32 columns, not by an extra MMA.
// Shapes outside the support test below (and any dropout mask) run on the mma.sync kernel of mlp_bf16.cu.
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "mlp_umma.cuh"

namespace lctr {
namespace umma {

constexpr int kTM = 128;        // samples per CTA
constexpr int kThreads = 256;   // 8 warps: lane quadrant q = warp & 3, column half h = warp >> 2
constexpr int kChunk = kTM * 16;  // bytes between 8-column chunks of a [128 x C] tile
constexpr uint32_t kColsDX = 0, kColsDW = 256;  // TMEM columns of the backward accumulators

__device__ __forceinline__ void stamp(const Dev& P, int& n) {
    if (P.trace && blockIdx.x == 0 && threadIdx.x == 0) {
        P.trace[n] = (unsigned long long)clock64();
    }
    n++;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// A wait that cannot hang the device: a descriptor or protocol bug traps instead.
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; spin++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared-memory matrix descriptor, no swizzle (cute/arch/mma_sm100_desc.hpp: start >> 4 at [0,14), LBO >> 4 at [16,30),
// SBO >> 4 at [32,46), version 1 at [46,48), layout type 0 at [61,64))
__device__ __forceinline__ uint64_t sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor, kind::f16: D fp32, A = B = bf16, M = 128
__device__ __forceinline__ uint32_t idesc(int n, uint32_t a_mn_major, uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t id, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(id), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float a) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory"); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
// activations.h:71-90,126-143 on the MUFU unit, branch free and without the denormal fix-ups of __expf / __fdividef (the
// epilogues are issue bound: 18 instructions per element with the intrinsics, 6 here).
//   Sigmoid: 1 / (1 + 2^t), t = -log2(e) * (acc + b) clamped to +-16 log2(e); the bias arrives pre-multiplied by -log2(e)
//            so that t is one FFMA.  The reference's clamp values (1e-7, 1 - 1e-7 beyond |x| = 16) become sigmoid(+-16)
//            = 1.1e-7 / 0.9999999: the same bf16 number above, 1.13e-7 instead of 1.0e-7 below.
//   Tanh:    tanh.approx (relative error 2^-11, an eighth of a bf16 ulp); saturates to +-1 by itself.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float kLog2e = 1.4426950408889634f;
template <int ACT>
__device__ __forceinline__ float bias_scale() { return ACT == LCTR_ACT_SIGMOID ? -kLog2e : 1.0f; }
template <int ACT>
__device__ __forceinline__ float fwd_act(float acc, float b_scaled) {
    if (ACT == LCTR_ACT_SIGMOID) {
        const float t = fminf(fmaxf(fmaf(acc, -kLog2e, b_scaled), -16.f * kLog2e), 16.f * kLog2e);
        return rcp_ftz(1.0f + ex2_ftz(t));
    } else {
        float y;
        asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(acc + b_scaled));
        return y;
    }
}
template <int ACT>
__device__ __forceinline__ float bwd_act(float fo) { return ACT == LCTR_ACT_SIGMOID ? fo * (1.0f - fo) : 1.0f - fo * fo; }
__device__ __forceinline__ float clip15(float v) { return fminf(fmaxf(v, -15.f), 15.f); }

// Column sums over the 32 lanes of a warp for 32 columns held one row per lane: on return lane j holds the sum of
// column j.  Each step folds the column set in half (lanes with the step bit keep the upper half), 31 shuffles in all.
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool upper = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < step; i++) {
            const float send = upper ? v[i] : v[i + step];
            const float keep = upper ? v[i + step] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return v[0];
}

template <int ACT>
__global__ void __launch_bounds__(kThreads, 1)
nfm_mlp_umma_kernel(Dev P, const float* __restrict__ z, float* __restrict__ dz, const float* __restrict__ wide,
                    const float* __restrict__ label, float* __restrict__ pred, int64_t rb, int B, double* partial,
                    unsigned int* done, double* out_slot) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_addr(smem);
    int ns = 0;
    stamp(P, ns);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, q = wid & 3, h = wid >> 2;
    const int row = q * 32 + lane;  // the TMEM lane this thread reads
    const int nh = P.nh;
    const int row0 = blockIdx.x * kTM;
    const int valid = min(kTM, B - row0);
    float* s_wl = reinterpret_cast<float*>(smem + P.wl_off);
    float* s_bias = reinterpret_cast<float*>(smem + P.bias_off);
    float* s_part = reinterpret_cast<float*>(smem + P.part_off);  // [2][128] partial output-layer dot products
    const uint32_t bar_w = sbase + P.bar_off, bar_m = bar_w + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P.bar_off + 16);
    const int64_t gi = rb + row0 + row;
    const float b_last = P.bias[nh][0];
    cudaTriggerProgrammaticLaunchCompletion();  // the dense updater behind this kernel may be scheduled as our CTAs retire

    if (tid == 0) {
        bar_init(bar_w, 1);
        bar_init(bar_m, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t total = 0;
        for (int l = 0; l < nh; l++) total += (uint32_t)P.out[l] * P.in[l] * 2;
        bar_expect_tx(bar_w, total);
        for (int l = 0; l < nh; l++) bulk_g2s(sbase + P.w_off[l], P.w16t[l], (uint32_t)P.out[l] * P.in[l] * 2, bar_w);
    }
    if (wid == 1) {  // the whole tensor memory; nothing else shares the SM (the CTA's shared memory sees to that)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int l = 0; l < nh; l++)
        for (int j = tid; j < P.out[l]; j += kThreads) s_bias[P.vec_off[l] + j] = P.bias[l][j] * bias_scale<ACT>();
    for (int i = tid; i < P.in[nh]; i += kThreads) s_wl[i] = P.w32_last[i];
    // Everything above is independent of the kernel in front (the embedding forward that writes z and the wide term): when
    // launched programmatically dependent, barriers, tensor memory, the weight copies and the bias vectors are under way
    // before that kernel has finished.  (No-op for an ordinary launch.)
    cudaGridDependencySynchronize();
    // the row's wide term and label: requested now, used after the forward pass
    const float wide_r = row < valid ? wide[gi] : 0.f;
    const float label_r = row < valid ? label[gi] : 0.f;
    {   // z tile -> bf16, chunk-major
        const int k = P.in[0], chunks = k / 8;
        for (int idx = tid; idx < kTM * chunks; idx += kThreads) {
            const int r = idx & (kTM - 1), ch = idx >> 7;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (r < valid) {
                const float4* src = reinterpret_cast<const float4*>(z + (size_t)(row0 + r) * k + ch * 8);
                a = src[0]; b = src[1];
            }
            *reinterpret_cast<uint4*>(smem + P.x_off[0] + ch * kChunk + r * 16) =
                make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w));
        }
    }
    fence_async_smem();
    tc_before();
    __syncthreads();
    tc_after();
    stamp(P, ns);
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t phase = 0;

    // ---- forward through the hidden layers (fullyconnLayer.h:80-118)
    float part = 0.f;
    for (int l = 0; l < nh; l++) {
        const int K = P.in[l], N = P.out[l];
        if (tid == 0) {
            if (l == 0) bar_wait(bar_w, 0);
            tc_after();
            const uint32_t id = idesc(N, 0, 0);
            for (int k = 0; k < K / 16; k++)
                mma(tmem, sdesc(sbase + P.x_off[l] + k * 2 * kChunk, kChunk, 128),
                    sdesc(sbase + P.w_off[l] + k * 2 * N * 16, N * 16, 128), id, k > 0);
            commit(bar_m);
        }
        __syncwarp();
        bar_wait(bar_m, phase); phase ^= 1;
        __syncwarp();
        tc_after();
        stamp(P, ns);
        const float* bias = s_bias + P.vec_off[l];
        unsigned char* y = smem + P.x_off[l + 1];
        const bool last = l == nh - 1;
        for (int c0 = h * (N / 2); c0 < (h + 1) * (N / 2); c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tlane + c0, r);
            float a[32];
#pragma unroll
            for (int g = 0; g < 8; g++) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias + c0 + g * 4);
                a[g * 4 + 0] = fwd_act<ACT>(__uint_as_float(r[g * 4 + 0]), b4.x);
                a[g * 4 + 1] = fwd_act<ACT>(__uint_as_float(r[g * 4 + 1]), b4.y);
                a[g * 4 + 2] = fwd_act<ACT>(__uint_as_float(r[g * 4 + 2]), b4.z);
                a[g * 4 + 3] = fwd_act<ACT>(__uint_as_float(r[g * 4 + 3]), b4.w);
            }
#pragma unroll
            for (int g = 0; g < 4; g++) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; j++) pk[j] = pack2(a[g * 8 + 2 * j], a[g * 8 + 2 * j + 1]);
                if (last) {  // output layer (linear, out = 1) on the rounded activations, like the other operands
                    const float4 w0 = *reinterpret_cast<const float4*>(s_wl + c0 + g * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(s_wl + c0 + g * 8 + 4);
                    const float2 a0 = unpack2(pk[0]), a1 = unpack2(pk[1]), a2 = unpack2(pk[2]), a3 = unpack2(pk[3]);
                    part += a0.x * w0.x + a0.y * w0.y + a1.x * w0.z + a1.y * w0.w + a2.x * w1.x + a2.y * w1.y + a3.x * w1.z + a3.y * w1.w;
                }
                *reinterpret_cast<uint4*>(y + (c0 / 8 + g) * kChunk + row * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
        if (last) s_part[h * kTM + row] = part;
        fence_async_smem();
        tc_before();
        __syncthreads();
        stamp(P, ns);
    }

    // ---- output layer, loss, delta of the last hidden layer in place, dW/db of the output layer, db of the last hidden
    float p_row = 0.5f;
    double loss = 0.0, correct = 0.0;
    {
        const int K = P.in[nh];
        float d3 = 0.f;
        if (row < valid) {
            const float o = s_part[row] + s_part[kTM + row] + b_last;
            p_row = ref_sigmoid(wide_r + o);  // train_nfm_algo.cpp:101-116
            if (h == 0) pred[gi] = p_row;
            d3 = clip15(p_row - label_r);
        }
        unsigned char* x = smem + P.x_off[nh];
        for (int c0 = h * (K / 2); c0 < (h + 1) * (K / 2); c0 += 32) {
            float gw[32], gd[32];
#pragma unroll
            for (int g = 0; g < 4; g++) {
                uint4* px = reinterpret_cast<uint4*>(x + (c0 / 8 + g) * kChunk + row * 16);
                const uint4 u = *px;
                const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = c0 + g * 8 + 2 * j;
                    const float2 a = unpack2(uu[j]);
                    gw[g * 8 + 2 * j] = d3 * a.x;  // weightDelta of the output layer (:165-178)
                    gw[g * 8 + 2 * j + 1] = d3 * a.y;
                    // no mask on the output layer's dX; previous activation' (:139-156)
                    pk[j] = pack2(clip15(d3 * s_wl[c] * bwd_act<ACT>(a.x)), clip15(d3 * s_wl[c + 1] * bwd_act<ACT>(a.y)));
                    const float2 e = unpack2(pk[j]);
                    gd[g * 8 + 2 * j] = e.x;
                    gd[g * 8 + 2 * j + 1] = e.y;
                }
                *px = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            const float sw = transpose_reduce32(gw, lane);
            const float sd = transpose_reduce32(gd, lane);
            red_add(P.dw[nh] + c0 + lane, sw);
            red_add(P.db[nh - 1] + c0 + lane, sd);  // biasDelta of the last hidden layer (:179)
        }
        const float dbl = warp_sum(d3);
        if (h == 0 && lane == 0) red_add(P.db[nh], dbl);
        fence_async_smem();
        tc_before();
        __syncthreads();
        stamp(P, ns);
    }

    // ---- backward through the hidden layers (fullyconnLayer.h:120-180)
    // TMEM columns of a layer's accumulators: dX at cx, dW at cw.  By default (0, 256); a layer whose two accumulators fit
    // into the 256 columns the previous layer's dX has vacated is issued EARLY -- before the previous layer's dW leaves
    // for global memory -- so that its MMAs run under that epilogue.
    auto dw_cols = [&](int l) { return (P.out[l] % 128 == 0) ? (P.out[l] / 128) * P.in[l] : (P.in[l] / 128) * P.out[l]; };
    auto issue_backward = [&](int l, uint32_t cx, uint32_t cw) {  // one thread
        const int K = P.in[l], N = P.out[l];
        tc_after();
        const uint32_t xd = sbase + P.x_off[l + 1], xl = sbase + P.x_off[l], wl = sbase + P.w_off[l];
        if (N % 128 == 0) {  // dW_l = delta_l^T . X_l, rows = out
            const uint32_t id = idesc(K, 1, 1);
            for (int m = 0; m < N / 128; m++)
                for (int k = 0; k < kTM / 16; k++)
                    mma(tmem + cw + m * K, sdesc(xd + m * 16 * kChunk + k * 256, 128, kChunk), sdesc(xl + k * 256, 128, kChunk), id, k > 0);
        } else {             // transposed: X_l^T . delta_l, rows = in
            const uint32_t id = idesc(N, 1, 1);
            for (int m = 0; m < K / 128; m++)
                for (int k = 0; k < kTM / 16; k++)
                    mma(tmem + cw + m * N, sdesc(xl + m * 16 * kChunk + k * 256, 128, kChunk), sdesc(xd + k * 256, 128, kChunk), id, k > 0);
        }
        {   // dX_l = delta_l . W_l
            const uint32_t id = idesc(K, 0, 1);
            for (int k = 0; k < N / 16; k++)
                mma(tmem + cx, sdesc(xd + k * 2 * kChunk, kChunk, 128), sdesc(wl + k * 256, 128, N * 16), id, k > 0);
        }
        commit(bar_m);
    };
    uint32_t cx = kColsDX, cw = kColsDW;
    bool issued = false;
    for (int l = nh - 1; l >= 0; l--) {
        const int K = P.in[l], N = P.out[l];
        const bool dw_normal = (N % 128) == 0;  // else transposed: rows = in_l
        if (!issued && tid == 0) issue_backward(l, cx, cw);
        __syncwarp();
        if (l == nh - 1 && h == 0 && row < valid) loss_terms(p_row, label_r, loss, correct);  // double-precision log, under the MMAs
        bar_wait(bar_m, phase); phase ^= 1;
        __syncwarp();
        tc_after();
        stamp(P, ns);
        if (l > 0) {  // delta_{l-1} = clip(dX_l * act'(x_l)) over x_l in place (:153-156), and its column sums = db_{l-1}
            unsigned char* xl = smem + P.x_off[l];
            for (int c0 = h * (K / 2); c0 < (h + 1) * (K / 2); c0 += 32) {
                uint32_t r[32];
                float gd[32];
                tmem_ld32(tlane + cx + c0, r);
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    uint4* px = reinterpret_cast<uint4*>(xl + (c0 / 8 + g) * kChunk + row * 16);
                    const uint4 u = *px;
                    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float2 a = unpack2(uu[j]);
                        pk[j] = pack2(clip15(__uint_as_float(r[g * 8 + 2 * j]) * bwd_act<ACT>(a.x)),
                                      clip15(__uint_as_float(r[g * 8 + 2 * j + 1]) * bwd_act<ACT>(a.y)));
                        const float2 e = unpack2(pk[j]);
                        gd[g * 8 + 2 * j] = e.x;
                        gd[g * 8 + 2 * j + 1] = e.y;
                    }
                    *px = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                const float sd = transpose_reduce32(gd, lane);
                red_add(P.db[l - 1] + c0 + lane, sd);
            }
        } else {  // dz: the gradient handed back to the embedding backward, fp32
            for (int c0 = h * (K / 2); c0 < (h + 1) * (K / 2); c0 += 8) {
                uint32_t r[8];
                tmem_ld8(tlane + cx + c0, r);
                if (row < valid) {
                    float4* dst = reinterpret_cast<float4*>(dz + (size_t)(row0 + row) * K + c0);
                    dst[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                    dst[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
                }
            }
        }
        stamp(P, ns);
        const uint32_t cw_l = cw;
        issued = false;
        if (l > 0) {
            const uint32_t ncw = (uint32_t)((P.in[l - 1] + 31) & ~31);
            if (cw_l >= kColsDW && ncw + dw_cols(l - 1) <= kColsDW) {  // uniform over the CTA; this layer's dW is outside [0, 256)
                fence_async_smem();
                tc_before();
                __syncthreads();
                cx = 0; cw = ncw;
                if (tid == 0) issue_backward(l - 1, cx, cw);
                __syncwarp();
                issued = true;
            } else {
                cx = kColsDX; cw = kColsDW;
            }
        }
        // dW_l -> dense gradient buffer (weightDelta, :165-178); fire-and-forget REDs
        if (dw_normal) {
            // thread = weight row: a direct RED would touch 32 lines per warp instruction.  32 x 32 tiles go through a
            // per-warp staging buffer (the delta_l tile, dead once this layer's MMAs have completed; 16 B chunks XOR-ed
            // with the row so that both sides are conflict free) and leave as 4 full 128 B lines per instruction.
            float* stage = reinterpret_cast<float*>(smem + P.x_off[l + 1] + wid * 4096);
            for (int m = 0; m < N / 128; m++) {
                float* dbase = P.dw[l] + (size_t)(m * 128 + q * 32) * K;
                if ((K / 2) % 32 == 0) {
                    for (int c0 = h * (K / 2); c0 < (h + 1) * (K / 2); c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(tlane + cw_l + m * K + c0, r);
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            *reinterpret_cast<uint4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int rr = 4 * i + (lane >> 3), j = lane & 7;
                            const float4 v = *reinterpret_cast<const float4*>(stage + rr * 32 + ((j ^ (rr & 7)) << 2));
                            red_add_v4(dbase + (size_t)rr * K + c0 + j * 4, v.x, v.y, v.z, v.w);
                        }
                        __syncwarp();
                    }
                } else {
                    float* drow = dbase + (size_t)lane * K;
                    for (int c0 = h * (K / 2); c0 < (h + 1) * (K / 2); c0 += 8) {
                        uint32_t r[8];
                        tmem_ld8(tlane + cw_l + m * K + c0, r);
                        red_add_v4(drow + c0, __uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                        red_add_v4(drow + c0 + 4, __uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
                    }
                }
            }
        } else {
            for (int m = 0; m < K / 128; m++) {
                float* dcol = P.dw[l] + m * 128 + row;  // thread = input index; lanes are contiguous in memory
                for (int c0 = h * (N / 2); c0 < (h + 1) * (N / 2); c0 += 8) {
                    uint32_t r[8];
                    tmem_ld8(tlane + cw_l + m * N + c0, r);
#pragma unroll
                    for (int j = 0; j < 8; j++) red_add(dcol + (size_t)(c0 + j) * K, __uint_as_float(r[j]));
                }
            }
        }
        fence_async_smem();
        tc_before();
        __syncthreads();
        stamp(P, ns);
    }
    if (wid == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    publish_stats(loss, correct, partial, done, out_slot, false);
    stamp(P, ns);
}

static size_t layout(const lctr_ctx* c, Dev& P) {
    const int nl = c->n_layers, nh = nl - 1;
    size_t off = 0;
    auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) & ~(align - 1); size_t o = off; off += bytes; return (int)o; };
    int voff = 0;
    P.nh = nh;
    for (int l = 0; l < nl; l++) {
        P.in[l] = c->layers[l].in; P.out[l] = c->layers[l].out;
        P.x_off[l] = take((size_t)kTM * P.in[l] * 2, 128);
    }
    for (int l = 0; l < nh; l++) {
        P.w_off[l] = take((size_t)P.out[l] * P.in[l] * 2, 128);
        P.vec_off[l] = voff; voff += P.out[l];
    }
    P.wl_off = take((size_t)P.in[nh] * 4, 16);
    P.bias_off = take((size_t)voff * 4, 16);
    P.part_off = take((size_t)2 * kTM * 4, 16);
    P.bar_off = take(32, 16);
    return off;
}

}  // namespace umma

// Shapes the tcgen05 kernel takes: every hidden width a multiple of 64 (<= 256), input width a multiple of 16, each
// layer's dW expressible with M = 128 (out or in a multiple of 128) inside 256 TMEM columns, everything in 227 KB.
bool mlp_umma_supported(const lctr_ctx* c) {
    const char* e = getenv("LCTR_MLP_UMMA");
    if (e && e[0] == '0') return false;
    const int nl = c->n_layers, nh = nl - 1;
    if (nh < 1 || c->layers[nh].out != 1 || c->layers[nh].in > 256) return false;
    if (c->layers[0].in % 16 != 0 || c->layers[0].in > 256) return false;
    for (int l = 0; l < nh; l++) {
        const int in = c->layers[l].in, out = c->layers[l].out;
        if (out % 64 != 0 || out > 256 || in > 256) return false;
        const int dw_cols = (out % 128 == 0) ? (out / 128) * in : (in % 128 == 0 ? (in / 128) * out : 1 << 20);
        if (dw_cols > 256) return false;
    }
    umma::Dev P;
    return umma::layout(c, P) <= (size_t)227 * 1024;
}

int mlp_umma_prepare(lctr_ctx* c) {
    umma::Dev P;
    const size_t need = umma::layout(c, P);
    c->mlp_umma_smem = need;
    LCTR_CUDA(cudaFuncSetAttribute(umma::nfm_mlp_umma_kernel<LCTR_ACT_SIGMOID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    LCTR_CUDA(cudaFuncSetAttribute(umma::nfm_mlp_umma_kernel<LCTR_ACT_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    return 0;
}

int launch_mlp_umma(lctr_ctx* c, Slot& s, int64_t rb, int B, double* out_slot) {
    umma::Dev P;
    umma::layout(c, P);
    const int nl = c->n_layers, nh = nl - 1;
    P.act = c->cfg.activation;
    for (int l = 0; l < nl; l++) {
        MlpLayer& L = c->layers[l];
        P.w16t[l] = (const __nv_bfloat16*)L.w16t; P.bias[l] = L.b; P.dw[l] = L.dw; P.db[l] = L.db;
    }
    P.w32_last = c->layers[nh].w;
    P.trace = nullptr;
    const bool trace = getenv("LCTR_MLP_UMMA_TRACE") && getenv("LCTR_MLP_UMMA_TRACE")[0] == '1';
    static unsigned long long* d_trace = nullptr;
    if (trace) {
        if (!d_trace) LCTR_CUDA(cudaMalloc((void**)&d_trace, 64 * sizeof(unsigned long long)));
        LCTR_CUDA(cudaMemsetAsync(d_trace, 0, 64 * sizeof(unsigned long long), c->stream));
        P.trace = d_trace;
    }
    const unsigned grid = (unsigned)((B + umma::kTM - 1) / umma::kTM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(umma::kThreads); cfg.dynamicSmemBytes = c->mlp_umma_smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (c->cfg.world == 1 && pdl_on()) ? 1 : 0;  // behind the embedding forward (fm_fused.cu, MODE 2)
    if (P.act == LCTR_ACT_SIGMOID)
        cudaLaunchKernelEx(&cfg, umma::nfm_mlp_umma_kernel<LCTR_ACT_SIGMOID>, P, (const float*)c->z, c->dz, (const float*)s.wide,
                           (const float*)s.label, s.pred, rb, B, c->stat_partial, c->stat_done, out_slot);
    else
        cudaLaunchKernelEx(&cfg, umma::nfm_mlp_umma_kernel<LCTR_ACT_TANH>, P, (const float*)c->z, c->dz, (const float*)s.wide,
                           (const float*)s.label, s.pred, rb, B, c->stat_partial, c->stat_done, out_slot);
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    if (trace) {  // phase boundaries of CTA 0: setup | per layer (mma wait, epilogue) | output | per layer (mma wait, dX, dW) | stats
        unsigned long long h[64];
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        LCTR_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[mlp_umma trace, SM cycles]");
        for (int i = 1; i < 64 && h[i]; i++) fprintf(stderr, " %llu", h[i] - h[i - 1]);
        fprintf(stderr, "  total %llu\n", h[0] ? [&] { int i = 1; while (i < 64 && h[i]) i++; return h[i - 1] - h[0]; }() : 0ull);
    }
    return 0;
}

}  // namespace lctr
