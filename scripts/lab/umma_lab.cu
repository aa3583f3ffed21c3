// scripts/lab/umma_lab.cu -- descriptor lab for the tcgen05 dense-layer kernel (lightctr_b200/csrc/mlp_umma.cu).
// One CTA, one tcgen05.mma chain per test, operands in the un-swizzled "chunk-major" shared-memory tile the kernel uses:
//     byte offset of element (r, c) of an R x C bf16 matrix  =  (c / 8) * (R * 16) + r * 16 + (c % 8) * 2
// i.e. 8 x 16 B core matrices, 8-row groups 128 B apart, 8-column chunks R*16 B apart.  The same tile is a K-major operand
// (rows = M or N, columns = K) and an MN-major operand (rows = K, columns = M or N); only LBO/SBO and the major bits of
// the descriptors change.  The lab checks all three products the kernel needs against a CPU fp32 reference and tries
// both assignments of LBO/SBO so that one run settles the encoding.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/lab/umma_lab scripts/lab/umma_lab.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // layout type 0 = no swizzle, base offset 0
}

struct Op { uint32_t lbo, sbo, kstep, major; };

__global__ void __launch_bounds__(128, 1)
umma_test(const __nv_bfloat16* __restrict__ A, int a_bytes, const __nv_bfloat16* __restrict__ B, int b_bytes, float* __restrict__ D,
          int N, int K, Op oa, Op ob, int* status) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    unsigned char* sa = smem;
    unsigned char* sb = smem + ((a_bytes + 127) & ~127);
    for (int i = tid; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(A)[i];
    for (int i = tid; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(B)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (oa.major << 15) | (ob.major << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        for (int k = 0; k < K / 16; k++) {
            const uint64_t da = make_desc(smem_u32(sa) + k * oa.kstep, oa.lbo, oa.sbo);
            const uint64_t db = make_desc(smem_u32(sb) + k * ob.kstep, ob.lbo, ob.sbo);
            const uint32_t acc = k > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    {   // bounded wait: a wrong descriptor must not hang the box
        uint32_t ok = 0;
        for (int spin = 0; spin < (1 << 22) && !ok; spin++) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        if (!ok && tid == 0) *status = 1;
    }
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tmem + ((uint32_t)(wid * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) D[(size_t)(wid * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// chunk-major tile of an R x C matrix
static std::vector<__nv_bfloat16> tile(const std::vector<float>& m, int R, int C) {
    std::vector<__nv_bfloat16> t((size_t)R * C);
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) t[(size_t)(c / 8) * R * 8 + (size_t)r * 8 + (c % 8)] = __float2bfloat16(m[(size_t)r * C + c]);
    return t;
}

static Op as_kmajor(int R, bool swap) {  // rows = M/N, columns = K
    Op o; o.lbo = R * 16; o.sbo = 128; o.kstep = 2 * R * 16; o.major = 0;
    if (swap) std::swap(o.lbo, o.sbo);
    return o;
}
static Op as_mnmajor(int R, bool swap) {  // rows = K, columns = M/N
    Op o; o.lbo = 128; o.sbo = R * 16; o.kstep = 256; o.major = 1;
    if (swap) std::swap(o.lbo, o.sbo);
    return o;
}

static double run(const char* name, const std::vector<__nv_bfloat16>& A, const std::vector<__nv_bfloat16>& B, const std::vector<float>& ref,
                  int N, int K, Op oa, Op ob) {
    __nv_bfloat16 *dA, *dB; float* dD; int* dS;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, (size_t)128 * N * 4)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, (size_t)128 * N * 4)); CK(cudaMemset(dS, 0, 4));
    const int smem = (int)(((A.size() * 2 + 127) & ~127) + B.size() * 2 + 128);
    CK(cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_test<<<1, 128, smem>>>(dA, (int)(A.size() * 2), dB, (int)(B.size() * 2), dD, N, K, oa, ob, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s launch failed: %s\n", name, cudaGetErrorString(e)); exit(3); }
    std::vector<float> D((size_t)128 * N); int st;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (size_t i = 0; i < D.size(); i++) { double d = std::fabs((double)D[i] - ref[i]); if (!(d <= worst)) worst = std::isnan(d) ? 1e30 : d; }
    printf("%-44s lboA=%5u sboA=%5u lboB=%5u sboB=%5u  timeout=%d  max|err|=%.3g  %s\n", name, oa.lbo, oa.sbo, ob.lbo, ob.sbo, st, worst,
           worst < 1e-2 ? "MATCH" : "");
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
    return worst;
}

int main() {
    srand(7);
    auto rnd = [](int n) { std::vector<float> v(n); for (auto& x : v) x = bf((float)rand() / RAND_MAX - 0.5f); return v; };
    // T1 forward: D[s][o] = sum_i X[s][i] W[o][i]   (A = X 128 x 64 K-major, B = W 64 x 64 K-major)
    {
        const int S = 128, I = 64, O = 64;
        auto X = rnd(S * I), W = rnd(O * I);
        std::vector<float> ref((size_t)S * O, 0.f);
        for (int s = 0; s < S; s++) for (int o = 0; o < O; o++) { float a = 0; for (int i = 0; i < I; i++) a += X[s * I + i] * W[o * I + i]; ref[s * O + o] = a; }
        auto tX = tile(X, S, I), tW = tile(W, O, I);
        for (int sw = 0; sw < 4; sw++) run("T1 fwd  A=K-major B=K-major", tX, tW, ref, O, I, as_kmajor(S, sw & 1), as_kmajor(O, sw >> 1));
    }
    // T2 dX: D[s][i] = sum_o Dl[s][o] W[o][i]        (A = Dl 128 x 64 K-major, B = W 64 x 32: rows = K -> MN-major)
    {
        const int S = 128, O = 64, I = 32;
        auto Dl = rnd(S * O), W = rnd(O * I);
        std::vector<float> ref((size_t)S * I, 0.f);
        for (int s = 0; s < S; s++) for (int i = 0; i < I; i++) { float a = 0; for (int o = 0; o < O; o++) a += Dl[s * O + o] * W[o * I + i]; ref[s * I + i] = a; }
        auto tD = tile(Dl, S, O), tW = tile(W, O, I);
        for (int sw = 0; sw < 2; sw++) run("T2 dX   A=K-major B=MN-major", tD, tW, ref, I, O, as_kmajor(S, 0), as_mnmajor(O, sw));
    }
    // T3 dW: D[o][i] = sum_s Dl[s][o] X[s][i]         (A = Dl 128 x 128: rows = K -> MN-major, B = X 128 x 64 MN-major)
    {
        const int S = 128, O = 128, I = 64;
        auto Dl = rnd(S * O), X = rnd(S * I);
        std::vector<float> ref((size_t)O * I, 0.f);
        for (int o = 0; o < O; o++) for (int i = 0; i < I; i++) { float a = 0; for (int s = 0; s < S; s++) a += Dl[s * O + o] * X[s * I + i]; ref[o * I + i] = a; }
        auto tD = tile(Dl, S, O), tX = tile(X, S, I);
        for (int sw = 0; sw < 4; sw++) run("T3 dW   A=MN-major B=MN-major", tD, tX, ref, I, S, as_mnmajor(S, sw & 1), as_mnmajor(S, sw >> 1));
    }
    // T4: N = 256, K = 16 (layer 0 of C4) and N = 16 (dX0 / dW0)
    {
        const int S = 128, I = 16, O = 256;
        auto X = rnd(S * I), W = rnd(O * I);
        std::vector<float> ref((size_t)S * O, 0.f);
        for (int s = 0; s < S; s++) for (int o = 0; o < O; o++) { float a = 0; for (int i = 0; i < I; i++) a += X[s * I + i] * W[o * I + i]; ref[s * O + o] = a; }
        run("T4 fwd  N=256 K=16", tile(X, S, I), tile(W, O, I), ref, O, I, as_kmajor(S, 0), as_kmajor(O, 0));
    }
    return 0;
}
