// umma_test.cu -- validates the tcgen05 building blocks of the network kernel in isolation (one CTA):
//   * K-major SWIZZLE_NONE shared-memory descriptors over the [kchunk][row][8 x f16] activation layout, with the
//     +-8-row (dy) shift applied through the descriptor start address,
//   * instruction descriptor for kind::f16 (f16 x f16 -> f32), M=128, N=96,
//   * TMEM alloc / tcgen05.mma / tcgen05.commit -> mbarrier / tcgen05.ld 32x32b,
// and times back-to-back MMAs to measure the achievable issue rate at this (small-N) shape.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#define ROWS 160          // activation rows in smem (tile rows are [16, 144))
#define KC 4              // 16-byte k-chunks (C = 32 channels)
#define NN 96             // N = 3 dx x 32 co

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;       // version = 1 (sm100)
    return d;                     // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}

// instruction descriptor: c_format F32 (1) @4, a_format F16 (0) @7, b_format F16 (0) @10, K-major both, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int fmt /*0 f16, 1 bf16*/)
{
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(160, 1) k_test(const __half *gA, const __half *gB, float *gD, long long *cycles, int reps)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sA = smem;                                  // [KC][ROWS][16 B]
    unsigned char *sB = smem + KC * ROWS * 16;                 // [3][KC][NN][16 B]
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + 3 * KC * NN * 16);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // global (row-major [ROWS][32], [3][96][32]) -> smem core-matrix layouts
    for (int i = threadIdx.x; i < ROWS * KC; i += blockDim.x) {
        int r = i / KC, kc = i % KC;
        *reinterpret_cast<uint4 *>(sA + (kc * ROWS + r) * 16) = *reinterpret_cast<const uint4 *>(gA + r * 32 + kc * 8);
    }
    for (int i = threadIdx.x; i < 3 * NN * KC; i += blockDim.x) {
        int dy = i / (NN * KC), n = (i / KC) % NN, kc = i % KC;
        *reinterpret_cast<uint4 *>(sB + ((dy * KC + kc) * NN + n) * 16) =
            *reinterpret_cast<const uint4 *>(gB + (dy * NN + n) * 32 + kc * 8);
    }
    if (threadIdx.x == 0) {
        mbar_init(s32(&bar[0]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" :: "r"(s32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = make_idesc(128, NN, 0);

    if (warp == 4 && lane == 0) {
        long long t0 = clock64();
        for (int rep = 0; rep < reps; rep++) {
            uint32_t acc = 0;
            for (int dy = 0; dy < 3; dy++)
                for (int ks = 0; ks < 2; ks++) {
                    uint64_t ad = make_desc(s32(sA) + (2 * ks) * ROWS * 16 + (16 + 8 * (dy - 1)) * 16, ROWS * 16, 128);
                    uint64_t bd = make_desc(s32(sB) + (dy * KC + 2 * ks) * NN * 16, NN * 16, 128);
                    mma_f16(tmem, ad, bd, idesc, acc);
                    acc = 1;
                }
        }
        mma_commit(s32(&bar[0]));
        mbar_wait(s32(&bar[0]), 0);
        long long t1 = clock64();
        cycles[0] = t1 - t0;
    }
    if (warp < 4) {
        mbar_wait(s32(&bar[0]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int blk = 0; blk < 3; blk++) {
            uint32_t r[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + blk * 32, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 32; j++) gD[(warp * 32 + lane) * NN + blk * 32 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" :: "r"(tmem) : "memory");
}

// TMEM read / write throughput: `nw` warps (nw = 4 or 8; warp%4 = lane quadrant) each issue `reps` x (32x32b.x32) loads
__global__ void __launch_bounds__(288, 1) k_tmem_bw(long long *cycles, float *sink, int nw, int reps, int do_store)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" :: "r"(s32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = slot;
    float acc = 0.f;
    long long t0 = clock64();
    if (warp < nw) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
        uint32_t r[32];
        for (int j = 0; j < 32; j++) r[j] = lane + j;
        for (int i = 0; i < reps; i++) {
            if (do_store) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                             "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n"
                             :: "r"(base + (i & 7) * 32), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                                "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                                "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                                "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
            } else {
                tmem_ld32(base + (i & 7) * 32, r);
                if ((i & 3) == 3) asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            }
        }
        if (do_store) asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        else asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int j = 0; j < 32; j++) acc += __uint_as_float(r[j]);
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tmem) : "memory");
}

int main()
{
    std::vector<__half> A(ROWS * 32), B(3 * NN * 32);
    std::vector<float> Af(ROWS * 32), Bf(3 * NN * 32);
    srand(1);
    for (size_t i = 0; i < A.size(); i++) { float v = (rand() % 17 - 8) / 8.0f; A[i] = __float2half(v); Af[i] = __half2float(A[i]); }
    for (size_t i = 0; i < B.size(); i++) { float v = (rand() % 13 - 6) / 16.0f; B[i] = __float2half(v); Bf[i] = __half2float(B[i]); }
    __half *dA, *dB; float *dD; long long *dC;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * NN * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    size_t smem = KC * ROWS * 16 + 3 * KC * NN * 16 + 64;
    cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_test<<<1, 160, smem>>>(dA, dB, dD, dC, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> D(128 * NN);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0;
    for (int r = 0; r < 128; r++)
        for (int n = 0; n < NN; n++) {
            double ref = 0;
            for (int dy = 0; dy < 3; dy++)
                for (int k = 0; k < 32; k++) ref += (double)Af[(16 + r + 8 * (dy - 1)) * 32 + k] * Bf[(dy * NN + n) * 32 + k];
            double err = fabs(ref - D[r * NN + n]);
            if (err > maxerr) maxerr = err;
            if (err > 1e-3 && bad++ < 5) printf("mismatch r=%d n=%d got %f want %f\n", r, n, D[r * NN + n], ref);
        }
    printf("max |err| = %g  (%s)\n", maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
    for (int reps : {100, 1000}) {
        k_test<<<1, 160, smem>>>(dA, dB, dD, dC, reps);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
        printf("reps %d: %lld cycles -> %.1f cycles per M128 N96 K16 MMA (tensor floor 48, smem-operand bound ~56)\n", reps, c,
               (double)c / (reps * 6));
    }
    float *sink;
    cudaMalloc(&sink, 4096);
    for (int st = 0; st < 2; st++)
        for (int nw : {4, 8}) {
            k_tmem_bw<<<1, 288>>>(dC, sink, nw, 2000, st);
            cudaError_t e2 = cudaDeviceSynchronize();
            long long c;
            cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
            printf("tmem %s, %d warps: %s, %.1f bytes/cycle/SM (%.1f cycles per 32x32b.x32 per warp)\n", st ? "st" : "ld", nw,
                   cudaGetErrorString(e2), (double)nw * 2000 * 4096 / c, (double)c / 2000);
        }
    return 0;
}
