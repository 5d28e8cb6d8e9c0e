// Microbenchmark behind DESIGN.md section 3 (profiles/r2b_mma_microbench.txt): cost of back-to-back tcgen05.mma (kind::f16, K16) as a
// function of N, accumulator dependency, CTA pairs, concurrent tcgen05.ld (+ 32 LDS per load) and the A operand in tensor memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_microbench tools/mma_microbench.cu && ./mma_microbench
// (main() holds the last sweep that was run; the three sweeps of the log differ only in the run<>() calls.)
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
template <int CG>
__device__ __forceinline__ void mma_e(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc)
{
    if (CG == 2)
        asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\telect.sync _|e, 0xffffffff;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\telect.sync _|e, 0xffffffff;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void mma_ts_e(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc)
{
    if (CG == 2)
        asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\telect.sync _|e, 0xffffffff;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\telect.sync _|e, 0xffffffff;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit_e(uint32_t bar)
{
    if (CG == 2)
        asm volatile("{\n\t.reg .pred e;\n\t.reg .b16 m;\n\telect.sync _|e, 0xffffffff;\n\tmov.b16 m, 3;\n\t@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar) : "memory");
    else
        asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t *r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
struct Args {
    int N, reps, nacc, ld_warps, a_mode, nb, lds, ts; // nacc: accumulators cycled; ld_warps: warps looping tcgen05.ld meanwhile; a_mode 0: same A every MMA, 1: A advances 32 B per MMA (Hankel-like walk); nb: distinct B blocks cycled
    unsigned long long *out;  // per CTA: mma clocks, ld count
    volatile int *dummy;
};
template <int CG>
__global__ void __launch_bounds__(320, 1) k_bench(Args a)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    __half *sA = reinterpret_cast<__half *>(sm);             // 32 KB: Hankel-style operand
    __half *sB = reinterpret_cast<__half *>(sm + 32768);     // nb blocks of N/CG rows x 16
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 32768 + 8 * 8192);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar + 4);
    volatile int *s_stop = reinterpret_cast<volatile int *>(bar + 6);
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sA[i] = __float2half(0.25f + 0.001f * (float)((i * 37) % 211));
    for (int i = threadIdx.x; i < 8 * 4096; i += blockDim.x) sB[i] = __float2half(0.5f - 0.002f * (float)((i * 53) % 197));
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        *s_stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    if (warp == 0) {
        if (rank == 0) {
            const uint32_t hi = (128u >> 4) | (1u << 14);
            const uint32_t a0 = ((smem_u32(sA) >> 4) & 0x3FFF) | ((128u >> 4) << 16);
            const int nh = a.N / CG;
            const uint32_t b0 = ((smem_u32(sB) >> 4) & 0x3FFF) | ((uint32_t)(nh * 16 >> 4) << 16);
            const uint32_t idesc = (1u << 4) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
            const uint32_t bstep = (uint32_t)(nh * 32) >> 4;
            // warm-up
            for (int i = 0; i < 64; i++) mma_e<CG>(tmem + (uint32_t)((i % a.nacc) * a.N), a0, b0, hi, idesc, i >= a.nacc);
            commit_e<CG>(smem_u32(bar));
            mbar_wait(smem_u32(bar), 0);
            const long long t0 = clock64();
            int ia = 0, ib = 0, ic = 0;
            if (a.ts) {
                for (int i = 0; i < a.reps; i += 3) {
                    const uint32_t d = tmem + (uint32_t)(ic * a.N), bb = b0 + bstep * (uint32_t)ib, at = tmem + 320u + 8u * (uint32_t)ia;
                    mma_ts_e<CG>(d, at, bb, hi, idesc, 1u);
                    mma_ts_e<CG>(d, at, bb + bstep, hi, idesc, 1u);
                    mma_e<CG>(d, a0 + 2u * (uint32_t)ia, bb, hi, idesc, 1u);
                    if (++ia == 24) ia = 0;
                    if (++ib >= a.nb - 1) ib = 0;
                    if (++ic == a.nacc) ic = 0;
                }
            } else {
                for (int i = 0; i < a.reps; i += 3) {
                    const uint32_t d = tmem + (uint32_t)(ic * a.N), bb = b0 + bstep * (uint32_t)ib, aa = a0 + 2u * (uint32_t)ia;
                    mma_e<CG>(d, aa, bb, hi, idesc, 1u);
                    mma_e<CG>(d, aa, bb + bstep, hi, idesc, 1u);
                    mma_e<CG>(d, aa + 80u, bb, hi, idesc, 1u);
                    if (++ia == 24) ia = 0;
                    if (++ib >= a.nb - 1) ib = 0;
                    if (++ic == a.nacc) ic = 0;
                }
            }
            commit_e<CG>(smem_u32(bar));
            const long long t1 = clock64();
            mbar_wait(smem_u32(bar), 1);
            const long long t2 = clock64();
            if (lane == 0) {
                a.out[4 * blockIdx.x + 0] = (unsigned long long)(t2 - t0);
                a.out[4 * blockIdx.x + 1] = (unsigned long long)(t1 - t0);
            }
        }
        else {
            mbar_wait(smem_u32(bar), 0);
            mbar_wait(smem_u32(bar), 1);
        }
        __syncwarp();
        if (lane == 0) *s_stop = 1;
    } else if (warp >= 2 && warp < 2 + a.ld_warps) {
        // concurrent accumulator read-back: lane quarter (warp & 3), columns cycling over the first 384
        const int wq = warp & 3;
        uint32_t r[32];
        unsigned long long cnt = 0;
        float s = 0.f;
        int col = 0;
        const long long t0 = clock64();
        while (!*s_stop) {
            tc_ld32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)col, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (a.lds) {
                const float *Ew = reinterpret_cast<const float *>(sm) + (warp & 3) * 32 + lane + col;
#pragma unroll
                for (int i = 0; i < 32; i++) s = fmaf(__uint_as_float(r[i]), Ew[i], s);
            } else {
#pragma unroll
                for (int i = 0; i < 32; i++) s += __uint_as_float(r[i]);
            }
            col += 32;
            if (col >= 192) col = 0;
            cnt++;
        }
        const long long t1 = clock64();
        if (s == 123.456f) *a.dummy = 1;
        if (lane == 0 && warp == 2) {
            a.out[4 * blockIdx.x + 2] = cnt;
            a.out[4 * blockIdx.x + 3] = (unsigned long long)(t1 - t0);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}
template <int CG>
static void run(int grid, int N, int reps, int nacc, int ldw, int a_mode, int nb, unsigned long long *d_out, int *d_dummy, int lds = 0, int ts = 0)
{
    Args a{N, reps, nacc, ldw, a_mode, nb, lds, ts, d_out, d_dummy};
    const size_t smem = 32768 + 8 * 8192 + 128;
    cudaFuncSetAttribute(k_bench<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(d_out, 0, sizeof(unsigned long long) * 4 * 148);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_bench<CG>, a);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) {
        printf("CG %d N %d: error %s / %s\n", CG, N, cudaGetErrorString(e), cudaGetErrorString(e2));
        exit(1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> h(4 * 148);
    cudaMemcpy(h.data(), d_out, sizeof(unsigned long long) * 4 * 148, cudaMemcpyDeviceToHost);
    double tot = 0, iss = 0, cnt = 0, ldt = 0;
    int n = 0;
    for (int i = 0; i < grid; i += CG) {
        tot += (double)h[4 * i];
        iss += (double)h[4 * i + 1];
        n++;
    }
    for (int i = 0; i < grid; i++) {
        cnt += (double)h[4 * i + 2];
        ldt += (double)h[4 * i + 3];
    }
    const double clk_mma = tot / n / reps;
    printf("CG %d grid %3d N %3d nacc %d ldw %d amode %d nb %d : %.1f clk/MMA (nominal %.0f; issue %.1f)  wall %.3f ms -> %.0f MHz", CG, grid, N, nacc, ldw, a_mode, nb, clk_mma,
           N / 2.0, iss / n / reps, ms, tot / n / (ms * 1e3));
    if (lds) printf("  [+32 LDS per ld]");
    if (ts) printf("  [2 of 3 A from TMEM]");
    if (ldw) printf("  | ld.x32 per warp: %.0f clk each, %.1f B/clk/SM", ldt / cnt, (double)ldw * 4096.0 * cnt / ldt);
    printf("\n");
}
int main()
{
    unsigned long long *d_out;
    int *d_dummy;
    cudaMalloc(&d_out, sizeof(unsigned long long) * 4 * 148);
    cudaMalloc(&d_dummy, 4);
    const int reps = 21000;
    for (int ts : {0, 1})
        for (int N : {192, 128, 96, 64, 32}) run<2>(148, N, reps, 1, 8, 1, 8, d_out, d_dummy, 1, ts);
    for (int N : {192, 128, 64}) run<1>(148, N, reps, 1, 8, 1, 8, d_out, d_dummy, 1, 1);
    for (int N : {192, 128, 64, 32}) run<2>(148, N, reps, 1, 0, 1, 8, d_out, d_dummy, 0, 1);
    return 0;
}
