// lightctr_b200/csrc/mlp.cu -- the dense part of NFM / Wide&Deep: Fully_Conn_Layer chain, batched.
//
// Reference semantics (train/layer/fullyconnLayer.h, per SAMPLE): forward y_i = <x, W[i,:]> + b_i with
// masked hidden neurons forced to 0 BEFORE the activation (:96-99,110-113; sigmoid(0)=0.5 flows on),
// last layer linear (:116); backward clips delta to +-15 (:129-131), dX_i = sum_j W[j,i] mask_j delta_j
// (:139-147, mask only on hidden layers), previous activation' (:153-156), dW[j,:] += delta_j x
// (:165-178, UNmasked delta), db += delta (:179); Adagrad on bias then weights (:194-197).
// Here the per-sample GEMVs become batched GEMMs over the B rows of the step:
//     fwd  Y = act(mask .* (X W^T + b))        NT gemm   [B,in]x[out,in]^T
//     dX   = (D .* mask) W                      NN gemm   [B,out]x[out,in]
//     dW   = D^T X   (sum over the batch)       TN gemm   [out,B]x[B,in]
// fp32 SIMT tiles here are the PARITY mode (fp32 like the reference, no FMA contraction).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace lctr {

constexpr int TM = 64, TN = 64, TK = 16;

// C[M][N] (+)= op(A)[M][K] * op(B)[K][N];   element accessors via strides so NT/NN/TN share one kernel.
// A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].  Epilogue: 0 none, 1 forward (bias+mask+act).
struct GemmEpi {
    int mode;           // 0: store; 1: forward epilogue
    const float* bias;  // [N]
    const float* mask;  // [N] or nullptr
    int act;            // -1 none (last layer), 0 sigmoid, 1 tanh
    int atomic;         // 1: atomicAdd into C (split-K)
};

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K,
                long sam, long sak, long sbk, long sbn, int ksplit, GemmEpi epi) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int kchunk = (K + ksplit - 1) / ksplit;
    const int kb = blockIdx.z * kchunk, ke = min(K, kb + kchunk);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = kb; k0 < ke; k0 += TK) {
        for (int i = threadIdx.x; i < TM * TK; i += 256) {
            int mm, kk;
            if (sak == 1) { kk = i % TK; mm = i / TK; } else { mm = i % TM; kk = i / TM; }
            const int gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < M && gk < ke) ? A[(long)gm * sam + (long)gk * sak] : 0.f;
        }
        for (int i = threadIdx.x; i < TN * TK; i += 256) {
            int nn, kk;
            if (sbk == 1) { kk = i % TK; nn = i / TK; } else { nn = i % TN; kk = i / TN; }
            const int gn = n0 + nn, gk = k0 + kk;
            Bs[kk][nn] = (gn < N && gk < ke) ? B[(long)gk * sbk + (long)gn * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (epi.mode == 1) {
                v += epi.bias[gn];
                if (epi.mask && epi.mask[gn] == 0.f) v = 0.f;  // masked neuron: pre-activation forced to 0
                if (epi.act == LCTR_ACT_SIGMOID) {
                    v = v < -16.f ? 1e-7f : (v > 16.f ? 0.99999988f : 1.0f / (1.0f + expf(-v)));  // activations.h:73-84
                } else if (epi.act == LCTR_ACT_TANH) {
                    const float t1 = expf(v), t2 = expf(-v);  // activations.h:132-138
                    v = (t1 - t2) / (t1 + t2);
                }
            }
            if (epi.atomic) atomicAdd(&C[(long)gm * N + gn], v);
            else C[(long)gm * N + gn] = v;
        }
    }
}

// p = sigmoid(wide + mlp_out); loss/acc; delta_L = p - y  (train_nfm_algo.cpp:101-116)
__global__ void nfm_loss_kernel(const float* __restrict__ wide, const float* __restrict__ out,
                                const float* __restrict__ label, float* __restrict__ pred, float* __restrict__ delta,
                                int64_t rb, int64_t n, double* partial, unsigned int* done, double* out_slot) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double loss = 0.0, correct = 0.0;
    if (i < n) {
        const float p = ref_sigmoid(wide[rb + i] + out[i]);
        pred[rb + i] = p;
        const float y = label[rb + i];
        loss_terms(p, y, loss, correct);
        delta[i] = p - y;
    }
    publish_stats(loss, correct, partial, done, out_slot, false);
}

// in-place clip to +-15 (matrix.h:152-162) and optional masked copy for the dX gemm
__global__ void clip_mask_kernel(float* __restrict__ d, float* __restrict__ dm, const float* __restrict__ mask,
                                 int64_t n, int out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = d[i];
    v = v < -15.f ? -15.f : (v > 15.f ? 15.f : v);
    d[i] = v;
    if (dm) dm[i] = mask ? v * mask[i % out] : v;
}
// delta_prev = dX .* act'(act_prev)   (activations.h:85-90 sigmoid: (d*f)*(1-f); :139-143 tanh: d*(1-f*f))
__global__ void act_backward_kernel(float* __restrict__ dx, const float* __restrict__ f, int64_t n, int act) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float fo = f[i], d = dx[i];
    dx[i] = act == LCTR_ACT_SIGMOID ? (d * fo) * (1.0f - fo) : d * (1.0f - fo * fo);
}
__global__ void colsum_kernel(const float* __restrict__ d, float* __restrict__ db, int64_t rows, int out) {
    // db[j] += sum_r d[r][j]
    const int j = blockIdx.x;
    double acc = 0.0;
    for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) acc += (double)d[r * out + j];
    __shared__ double sh[32];
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double a = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        a = warp_sum_d(a);
        if (threadIdx.x == 0) db[j] += (float)a;
    }
}
// AdagradUpdater_Num::update on a dense array (gradientUpdater.h:139-150)
__global__ void adagrad_dense_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ acc, size_t n,
                                     float invB, float lr) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g1 = g[i] * invB;
    if (g1 != 0.f) {
        const float a = acc[i] + g1 * g1;
        acc[i] = a;
        w[i] = (float)((double)w[i] - (double)(lr * g1) / sqrt((double)a + 1e-7));
    }
    g[i] = 0.f;
}

static int gemm(lctr_ctx* c, const float* A, const float* B, float* C, int M, int N, int K, long sam, long sak,
                long sbk, long sbn, int ksplit, GemmEpi epi) {
    dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, ksplit);
    gemm_f32_kernel<<<grid, 256, 0, c->stream>>>(A, B, C, M, N, K, sam, sak, sbk, sbn, ksplit, epi);
    c->launches++;
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

int mlp_alloc(lctr_ctx* c) {
    const lctr_cfg& cf = c->cfg;
    LCTR_CHECK(cf.n_hidden >= 1 && cf.n_hidden <= LCTR_MAX_LAYERS, "NFM needs 1..%d hidden layers", LCTR_MAX_LAYERS);
    c->n_layers = cf.n_hidden + 1;
    size_t total = 0;
    int in = (int)cf.factor_cnt;
    for (int l = 0; l < c->n_layers; l++) {
        MlpLayer& L = c->layers[l];
        L.in = in;
        L.out = l < cf.n_hidden ? (int)cf.hidden[l] : 1;
        LCTR_CHECK(L.out > 0, "hidden[%d] must be > 0", l);
        total += (size_t)L.out * L.in + L.out;
        in = L.out;
    }
    c->dense_grad_n = total;
    LCTR_CUDA(cudaMalloc((void**)&c->dense_grad, total * sizeof(float)));
    LCTR_CUDA(cudaMemsetAsync(c->dense_grad, 0, total * sizeof(float), c->stream));
    size_t off = 0;
    for (int l = 0; l < c->n_layers; l++) {
        MlpLayer& L = c->layers[l];
        const size_t nw = (size_t)L.out * L.in;
        LCTR_CUDA(cudaMalloc((void**)&L.w, nw * sizeof(float)));
        LCTR_CUDA(cudaMalloc((void**)&L.b, L.out * sizeof(float)));
        LCTR_CUDA(cudaMalloc((void**)&L.mask, L.out * sizeof(float)));
        LCTR_CUDA(cudaMalloc((void**)&L.acc_w, nw * sizeof(float)));
        LCTR_CUDA(cudaMalloc((void**)&L.acc_b, L.out * sizeof(float)));
        LCTR_CUDA(cudaMemsetAsync(L.w, 0, nw * sizeof(float), c->stream));
        LCTR_CUDA(cudaMemsetAsync(L.b, 0, L.out * sizeof(float), c->stream));
        LCTR_CUDA(cudaMemsetAsync(L.acc_w, 0, nw * sizeof(float), c->stream));
        LCTR_CUDA(cudaMemsetAsync(L.acc_b, 0, L.out * sizeof(float), c->stream));
        std::vector<float> ones(L.out, 1.f);
        LCTR_CUDA(cudaMemcpyAsync(L.mask, ones.data(), L.out * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        LCTR_CUDA(cudaStreamSynchronize(c->stream));
        L.dw = c->dense_grad + off; off += nw;   // registerGradient order: weightDelta then biasDelta (fullyconnLayer.h:69-75)
        L.db = c->dense_grad + off; off += L.out;
    }
    return 0;
}

int mlp_free(lctr_ctx* c) {
    for (int l = 0; l < c->n_layers; l++) {
        MlpLayer& L = c->layers[l];
        if (L.w) cudaFree(L.w); if (L.b) cudaFree(L.b); if (L.mask) cudaFree(L.mask);
        if (L.acc_w) cudaFree(L.acc_w); if (L.acc_b) cudaFree(L.acc_b);
        if (L.act) cudaFree(L.act); if (L.delta) cudaFree(L.delta);
        L = MlpLayer();
    }
    if (c->dense_grad) cudaFree(c->dense_grad);
    if (c->z) cudaFree(c->z); if (c->dz) cudaFree(c->dz); if (c->mlp_out) cudaFree(c->mlp_out);
    c->dense_grad = c->z = c->dz = c->mlp_out = nullptr;
    c->n_layers = 0;
    return 0;
}

int mlp_reserve(lctr_ctx* c, int64_t rows) {
    if ((size_t)rows <= c->mlp_cap_rows) return 0;
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    const size_t cap = (size_t)rows;
    const size_t k = c->cfg.factor_cnt;
    if (c->z) cudaFree(c->z); if (c->dz) cudaFree(c->dz); if (c->mlp_out) cudaFree(c->mlp_out);
    LCTR_CUDA(cudaMalloc((void**)&c->z, cap * k * sizeof(float)));
    LCTR_CUDA(cudaMalloc((void**)&c->dz, cap * k * sizeof(float)));
    LCTR_CUDA(cudaMalloc((void**)&c->mlp_out, cap * sizeof(float)));
    for (int l = 0; l < c->n_layers; l++) {
        MlpLayer& L = c->layers[l];
        if (L.act) cudaFree(L.act); if (L.delta) cudaFree(L.delta);
        LCTR_CUDA(cudaMalloc((void**)&L.act, cap * L.out * sizeof(float)));
        LCTR_CUDA(cudaMalloc((void**)&L.delta, cap * (size_t)std::max(L.out, L.in) * sizeof(float)));
    }
    c->mlp_cap_rows = cap;
    return 0;
}

// forward MLP on c->z, loss, backward to c->dz, accumulate dW/db, Adagrad on the MLP.
int launch_nfm_mlp(lctr_ctx* c, Slot& s, int64_t rb, int64_t re, int64_t rows_divisor) {
    const int B = (int)(re - rb);
    const int nl = c->n_layers;
    // ---- forward
    const float* x = c->z;
    for (int l = 0; l < nl; l++) {
        MlpLayer& L = c->layers[l];
        GemmEpi e{1, L.b, l + 1 < nl ? L.mask : nullptr, l + 1 < nl ? c->cfg.activation : -1, 0};
        // Y[B][out] = X[B][in] * W[out][in]^T : A=X (sam=in, sak=1), B(k,n)=W[n][k] (sbk=1, sbn=in)
        if (gemm(c, x, L.w, L.act, B, L.out, L.in, L.in, 1, 1, L.in, 1, e)) return 1;
        x = L.act;
    }
    // ---- loss, delta of the output layer
    double* out_slot = c->stats + 2 * (c->step % kStatRing);
    MlpLayer& last = c->layers[nl - 1];
    nfm_loss_kernel<<<(B + 255) / 256, 256, 0, c->stream>>>(s.wide, last.act, s.label, s.pred, last.delta, rb, B,
                                                           c->stat_partial, c->stat_done, out_slot);
    c->launches++;
    // ---- backward
    for (int l = nl - 1; l >= 0; l--) {
        MlpLayer& L = c->layers[l];
        const bool hidden = l + 1 < nl;
        const int64_t n = (int64_t)B * L.out;
        // scratch for the masked delta: reuse act of this layer?  no -- act is needed by layer l+1's dW (already done)
        // and by act_backward of THIS layer's output (done when processing l+1).  So L.act is free now for hidden l,
        // but the last layer's act is tiny; use a dedicated region: the upper half of next-lower delta buffer is not
        // safe, so masked delta goes into L.act (its consumers have all run).
        float* dm = hidden ? L.act : nullptr;
        const float* xin = l == 0 ? c->z : c->layers[l - 1].act;
        // clip (in place) + masked copy
        clip_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(L.delta, dm, hidden ? L.mask : nullptr, n, L.out);
        c->launches++;
        // dW[out][in] += D^T X : A(m=j,k=r) = D[r*out + j] (sam=1, sak=out); B(k=r,n=i) = X[r*in + i] (sbk=in, sbn=1)
        {
            GemmEpi e{0, nullptr, nullptr, -1, 1};
            int ksplit = std::max(1, std::min(64, B / 256));
            if (gemm(c, L.delta, xin, L.dw, L.out, L.in, B, 1, L.out, L.in, 1, ksplit, e)) return 1;
        }
        colsum_kernel<<<L.out, 256, 0, c->stream>>>(L.delta, L.db, B, L.out);
        c->launches++;
        // dX[B][in] = Dm[B][out] * W[out][in] : A=Dm (sam=out, sak=1); B(k=j,n=i)=W[j*in+i] (sbk=in, sbn=1)
        float* dx = l == 0 ? c->dz : c->layers[l - 1].delta;
        {
            GemmEpi e{0, nullptr, nullptr, -1, 0};
            if (gemm(c, hidden ? dm : L.delta, L.w, dx, B, L.in, L.out, L.out, 1, L.in, 1, 1, e)) return 1;
        }
        if (l > 0) {
            const int64_t m = (int64_t)B * L.in;
            act_backward_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(dx, c->layers[l - 1].act, m,
                                                                                    c->cfg.activation);
            c->launches++;
        }
    }
    // ---- Adagrad on bias then weights, per layer (fullyconnLayer.h:194-197)
    const uint64_t mb = c->cfg.minibatch_size ? c->cfg.minibatch_size : (uint64_t)rows_divisor;
    const float invB = (float)(1.0 / (double)mb);
    for (int l = 0; l < nl; l++) {
        MlpLayer& L = c->layers[l];
        const size_t nw = (size_t)L.out * L.in;
        adagrad_dense_kernel<<<(unsigned)((L.out + 255) / 256), 256, 0, c->stream>>>(L.b, L.db, L.acc_b, L.out, invB,
                                                                                     c->cfg.learning_rate);
        adagrad_dense_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, c->stream>>>(L.w, L.dw, L.acc_w, nw, invB,
                                                                                  c->cfg.learning_rate);
        c->launches += 2;
    }
    LCTR_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lctr

using namespace lctr;
extern "C" {
int lctr_mlp_upload(lctr_ctx* c, int layer, const float* weight, const float* bias) {
    LCTR_CHECK(c && layer >= 0 && layer < c->n_layers, "mlp layer %d out of range", layer);
    MlpLayer& L = c->layers[layer];
    if (weight) LCTR_CUDA(cudaMemcpyAsync(L.w, weight, (size_t)L.out * L.in * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (bias) LCTR_CUDA(cudaMemcpyAsync(L.b, bias, (size_t)L.out * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int lctr_mlp_download(lctr_ctx* c, int layer, float* weight, float* bias) {
    LCTR_CHECK(c && layer >= 0 && layer < c->n_layers, "mlp layer %d out of range", layer);
    MlpLayer& L = c->layers[layer];
    if (weight) LCTR_CUDA(cudaMemcpyAsync(weight, L.w, (size_t)L.out * L.in * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (bias) LCTR_CUDA(cudaMemcpyAsync(bias, L.b, (size_t)L.out * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int lctr_mlp_set_mask(lctr_ctx* c, int layer, const float* mask) {
    LCTR_CHECK(c && mask && layer >= 0 && layer < c->n_layers, "mlp layer %d out of range", layer);
    MlpLayer& L = c->layers[layer];
    LCTR_CUDA(cudaMemcpyAsync(L.mask, mask, (size_t)L.out * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    LCTR_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
}
