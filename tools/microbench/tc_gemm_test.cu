// tc_gemm_test.cu -- unit test + latency probe for the tcgen05 building block of the tensor-core
// UDE kernels:  D[128 x 32] (fp32, TMEM) = A[128 x 32] * B[32 x 32]^T  with a 3xTF32 split
// (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo), operands K-major / no swizzle in shared memory (SS form) or A in
// tensor memory (TS form).  One CTA of 128 threads, thread t owns row t (one trajectory).
// Prints the max error against an fp64 CPU product and the round-trip latency of one
// stage -> MMA -> commit -> wait -> tcgen05.ld cycle.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_rna(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   bits [0,14) start>>4, [16,30) leading byte offset>>4 (between the two 16-byte K chunks of one MMA),
//   [32,46) stride byte offset>>4 (between 8-row groups), [46,48) version = 1, [61,64) layout type = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @4, a/b format TF32 = 2 @7/@10,
// a_major/b_major = K (0) @15/@16, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

constexpr uint32_t LBO = 128, SBO = 1024;   // [row/8][k/4][row%8][k%4] : 8 K-chunks of a row group are adjacent

// MODE 0: A from shared memory (SS), MODE 1: A from tensor memory (TS)
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_test(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D,
                                                 int iters, long long *cycles)
{
    __shared__ __align__(1024) float sAhi[128 * 32], sAlo[128 * 32], sBhi[32 * 32], sBlo[32 * 32];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t d_tmem = tmem_base;             // columns [0,32): accumulator
    const uint32_t a_hi_tmem = tmem_base + 32;     // columns [32,64): A_hi (TS form)
    const uint32_t a_lo_tmem = tmem_base + 64;     // columns [64,96): A_lo
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;   // this warp's TMEM lane quadrant
    // ---- stage B (weights): element (n, k) at (n/8)*SBO + (k/4)*LBO + (n%8)*16 + (k%4)*4 bytes ----
    for (int e = t; e < 32 * 32; e += 128) {
        const int n = e >> 5, k = e & 31;
        const float x = B[n * 32 + k], hi = tf32_rna(x), lo = tf32_rna(x - hi);
        const int off = (n >> 3) * (SBO / 4) + (k >> 2) * (LBO / 4) + (n & 7) * 4 + (k & 3);
        sBhi[off] = hi;
        sBlo[off] = lo;
    }
    float a[32], ahi[32], alo[32], d[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = A[t * 32 + k];
    uint32_t parity = 0;
    const uint32_t idesc = make_idesc(128, 32);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // ---- stage A: split and write this thread's row ----
#pragma unroll
        for (int k = 0; k < 32; ++k) { ahi[k] = tf32_rna(a[k]); alo[k] = tf32_rna(a[k] - ahi[k]); }
        if (MODE == 0) {
            const int rowoff = (t >> 3) * (SBO / 4) + (t & 7) * 4;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                *reinterpret_cast<float4 *>(&sAhi[rowoff + c * (LBO / 4)]) = make_float4(ahi[4 * c], ahi[4 * c + 1], ahi[4 * c + 2], ahi[4 * c + 3]);
                *reinterpret_cast<float4 *>(&sAlo[rowoff + c * (LBO / 4)]) = make_float4(alo[4 * c], alo[4 * c + 1], alo[4 * c + 2], alo[4 * c + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        } else {
            tmem_st32(a_hi_tmem + lane_base, ahi);
            tmem_st32(a_lo_tmem + lane_base, alo);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {   // K = 32 = 4 MMAs of K = 8 (two 16-byte chunks each)
                const uint64_t dbh = make_desc(smem_u32(sBhi) + kb * 2 * LBO, LBO, SBO);
                const uint64_t dbl = make_desc(smem_u32(sBlo) + kb * 2 * LBO, LBO, SBO);
                if (MODE == 0) {
                    const uint64_t dah = make_desc(smem_u32(sAhi) + kb * 2 * LBO, LBO, SBO);
                    const uint64_t dal = make_desc(smem_u32(sAlo) + kb * 2 * LBO, LBO, SBO);
                    mma_ss(d_tmem, dah, dbh, idesc, kb > 0);
                    mma_ss(d_tmem, dal, dbh, idesc, 1);
                    mma_ss(d_tmem, dah, dbl, idesc, 1);
                } else {
                    mma_ts(d_tmem, a_hi_tmem + kb * 8, dbh, idesc, kb > 0);
                    mma_ts(d_tmem, a_lo_tmem + kb * 8, dbh, idesc, 1);
                    mma_ts(d_tmem, a_hi_tmem + kb * 8, dbl, idesc, 1);
                }
            }
            mma_commit(smem_u32(&mbar));
        }
        mbar_wait(smem_u32(&mbar), parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_ld32(d_tmem + lane_base, d);
        if (it + 1 < iters) {
#pragma unroll
            for (int k = 0; k < 32; ++k) a[k] = a[k] + 1e-30f * d[k];   // dependency chain between iterations
        }
    }
    long long t1 = clock64();
#pragma unroll
    for (int k = 0; k < 32; ++k) D[t * 32 + k] = d[k];
    if (t == 0 && cycles) *cycles = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
}

int main()
{
    std::vector<float> A(128 * 32), B(32 * 32), D(128 * 32);
    srand(1);
    for (auto &x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto &x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    long long *dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    int fails = 0;
    for (int mode = 0; mode < 2; ++mode) {
        for (int iters : {1, 200}) {
            cudaMemset(dD, 0, D.size() * 4);
            if (mode == 0) k_test<0><<<1, 128>>>(dA, dB, dD, iters, dC);
            else k_test<1><<<1, 128>>>(dA, dB, dD, iters, dC);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d iters %d: CUDA error %s\n", mode, iters, cudaGetErrorString(e)); return 2; }
            long long cyc = 0;
            cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0, maxref = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 32; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 32; ++k) ref += (double)A[m * 32 + k] * (double)B[n * 32 + k];
                    maxerr = fmax(maxerr, fabs(ref - D[m * 32 + n]));
                    maxref = fmax(maxref, fabs(ref));
                }
            printf("%s iters=%3d  max|err|=%.3e (max|ref|=%.2f)  cycles/iter=%.0f\n", mode == 0 ? "SS (A in smem)" : "TS (A in tmem)", iters,
                   maxerr, maxref, (double)cyc / iters);
            if (iters == 1 && maxerr > 1e-5) ++fails;
        }
    }
    printf(fails ? "FAILED\n" : "OK\n");
    return fails;
}
