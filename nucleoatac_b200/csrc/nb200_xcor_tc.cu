// nb200_xcor_tc.cu -- the dense background cross-correlation (BiasTrack.calculateBackgroundSignal,
// nucleoatac/NucleosomeCalling.py:60-63) as a tcgen05 tensor-core contraction.
//
//   bx[x] = sum_i sum_k f_i V[i,k] Bp[i, x-w+k],   Bp[i,c] = E[c-(i-1)//2] * E[c+i//2]   (chunkmat2d.py:140-156)
//
// Every cell of the bias matrix is a product of two taps of the SAME 1-D track, so with a = left-tap offset and
// b = right-tap offset relative to x the sum is a windowed bilinear form
//
//   bx[x] = sum_a E[x+a] * H[x,a],      H[x,a] = sum_b G[a,b] * E[x+b],      G[a,b] = f_i V[i,k]  ((a,b) <-> (i,k), 1:1)
//
// H = Hankel(E) x G^T is a GEMM whose A operand (M = 128 output positions, K = b) is a Hankel matrix of the track:
// A[m,kb] = Es[m + kb].  It is never expanded: shared memory holds the track once per 8 element shifts
// (Z[t][s][0..7] = Es[8t+s .. 8t+s+7], 16 B rows) and the UMMA shared-memory descriptor walks it with
// SBO = LBO = 128 B, i.e. the 8x16B core matrices of neighbouring row groups / K chunks overlap in memory.
// B = G (constant per run: VMat x fragment-size distribution) is pre-split, pre-tiled and block-sparsified on the
// host (the non-zero region of G is a parallelogram) and streamed L2 -> SMEM with cp.async.bulk through an mbarrier
// ring.  fp64-grade accuracy of the operands comes from a 2-term fp16 split (hi + lo, 22 bits) of both operands,
// scaled by powers of two into the fp16 range: three MMAs hi*hi + hi*lo + lo*hi accumulate in fp32 TMEM.
//
// Two kernels, both persistent (one CTA per SM, CTA pairs on a TPC contracting their x-tiles with cta_group::2 MMAs of
// M = 256), warp-specialised and connected by mbarriers only:
//   k_nuc_bx_ts  the default whenever a CTA's half of the G image fits in shared memory (up to ~251 x 251): G resident, the hi
//                half of the Hankel operand of one x-tile expanded in TENSOR MEMORY (two of a block's three MMAs read A from
//                TMEM), issue loop on the uniform datapath, slabs of 128 columns through two accumulators;
//   k_nuc_bx_tc  both operands in shared memory, G streamed L2 -> SMEM through a ring (or resident), slabs of 192 columns,
//                one accumulator per x-tile: larger VMats, and NB200_TC_TS=0.
// H is produced in slabs (a taps); every K16 block of a slab is contracted by MMAs whose N is TRIMMED to the 16-row granules
// of G that are non-zero in that block (the non-zero region of G is a hexagon inside the NA x NB box), so the tensor pipe
// neither multiplies nor fetches the zero corners.  The epilogue warps read a finished slab back (tcgen05.ld) and contract
// it with the fp32 track: runs of 8 products in fp32, a 32-column chunk's runs in fp32, the chunks as an error-free float
// pair.  What each kernel is bound by, and what was measured on the way, is in DESIGN.md section 3.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nb200_dev.cuh"

#define TC_XT 2                           // x-tiles of 128 outputs per work item (they share every G stage)
#define TC_M 128
#define TC_TX (TC_XT * TC_M)
#ifndef TC_N
#define TC_N 192                          // slab width (a taps): one accumulator of TC_N TMEM columns per x-tile
#endif
#define TC_MAX_STAGES 5
#define TC_TMEM_COLS 512                  // allocation (power of two >= TC_XT * TC_N); one persistent CTA per SM
#define TC_BLOCK_BYTES (TC_N * 16 * 2 * 2)      // one untrimmed K16 block of G: hi + lo images
#define TC_SLOT_BYTES (3 * TC_BLOCK_BYTES)      // ring slot = one G stage (as many trimmed K16 blocks as fit)
#define TC_MAX_STAGE_BLOCKS 8
#define TC_EPI_WARPS (4 * TC_XT)          // 4 warps (one TMEM lane quarter each) per x-tile
#define TC_WARP_PROD TC_EPI_WARPS
#define TC_WARP_MMA (TC_EPI_WARPS + 1)
#define TC_WARP_PREP (TC_EPI_WARPS + 1 + TC_XT)   // first of the operand-generation warps (one MMA warp per x-tile before them)
#define TC_PREP_WARPS 4
#define TC_PREP_THREADS (32 * TC_PREP_WARPS)
#define TC_THREADS (32 * (TC_EPI_WARPS + 1 + TC_XT + TC_PREP_WARPS))

struct TcPlan {
    bool ok = false;
    int A0 = 0, B0 = 0, NA = 0, NB = 0, NAp = 0, NBp = 0, gmin = 0, span = 0;
    int n_stages = 0, n_achunks = 0, n_blocks = 0;
    int sG = 0;           // G scaled by 2^sG
    int has_row1 = 0;     // insert size 1 has a single tap (linear term), kept out of G
    int pair = 1;         // 1: images laid out for CTA pairs (cta_group::2: each CTA stages half of the rows of a block)
    size_t img_bytes = 0; // whole G image (all stages, both halves)
    double density = 0.0; // MMA columns issued / (NAp * NBp / 16)
    DevBuf g_img, stage_tab, block_tab, t_row1, emax;
    std::vector<int4> h_tab, h_blk;
    // second plan, for k_nuc_bx_ts (hi half of the Hankel operand in tensor memory): slabs of TS_N rows, every block image resident
    bool ts_ok = false;
    int ts_slabs = 0, ts_blocks = 0, ts_rank_bytes = 0, ts_c_split = 0, ts_c_end = 0, ts_q_need2 = -1, ts_t_need2 = 0;
    DevBuf ts_img;
    std::vector<int4> ts_blk;     // block and slab tables: passed to the kernel by value (TsTab)
    std::vector<int2> ts_slab;
};

static TcPlan *plan_of(nb200_ctx *ctx, bool create)
{
    if (!ctx->tc_plan && create) ctx->tc_plan = new TcPlan();
    return static_cast<TcPlan *>(ctx->tc_plan);
}

void nb200_tc_release(nb200_ctx *ctx)
{
    TcPlan *pl = static_cast<TcPlan *>(ctx->tc_plan);
    if (!pl) return;
    pl->g_img.release();
    pl->stage_tab.release();
    pl->block_tab.release();
    pl->t_row1.release();
    pl->emax.release();
    pl->ts_img.release();
    delete pl;
    ctx->tc_plan = nullptr;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Protocol-bug aid: a wait that gives up records who waited on what in host-mapped memory (if the host installed some,
// NB200_TC_DEBUG) before it traps, so that the host can say which barrier starved.
__device__ int *g_tc_trap_info = nullptr;
__device__ __noinline__ void tc_trap(uint32_t bar, uint32_t parity, int kind)
{
    int *t = g_tc_trap_info;
    if (t) {
        if ((threadIdx.x & 31) == 0) {   // one record per starving warp
            const int slot = atomicAdd(t, 1);
            if (slot < 500) {
                int *r = t + 16 + 4 * slot;
                r[0] = (int)blockIdx.x;
                r[1] = (int)threadIdx.x;
                r[2] = (int)bar;
                r[3] = (int)parity | (kind << 8);
            }
            __threadfence_system();
        }
        for (int i = 0; i < 200000; i++) __nanosleep(1000);   // let the other starving warps report before the context dies
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    int spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1 << 20)) tc_trap(bar, parity, 0);  // never hang the GPU on a protocol bug
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// barrier shared by the two CTAs of a pair: waits acquire at cluster scope, the peer arrives through its cluster address
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    int spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1 << 20)) tc_trap(bar, parity, 1);
    }
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta)   // arrive on `bar` of CTA `cta` of the cluster
{
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// Same without memory ordering: for hand-offs whose payload is tensor memory (ordered by tcgen05.fence / tcgen05.wait), not
// generic memory.  The .release form costs a MEMBAR.ALL.GPU per arrive (SASS), ~1 k clk on the read-back path of every slab.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar, uint32_t cta)
{
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(bar), "r"(cta));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit_e(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
        : "memory");
}
// Shared-memory operands are K-major, no-swizzle matrix descriptors (cute::UMMA::SmemDescriptor): 8x16B core matrices, low
// word = start address >> 4 | (LBO >> 4) << 16 (LBO = byte stride between the 16-byte K chunks), high word = SBO >> 4 (byte
// stride between 8-row groups) | version 1 << 14.  CTA-pair forms (cta_group::2): one MMA of M = 256 spans the two CTAs of
// the pair -- rows 0..127 from the leader's A operand and accumulator, rows 128..255 from the peer's, each CTA's shared
// memory holding half of the N rows of B; issued by the leader only; a commit can arrive on the same barrier of both CTAs.
// The three MMAs of one K16 block -- hi*hi (accumulate = acc0), hi*lo, lo*hi -- behind ONE election: the issuing warp
// pays for every operand it moves into uniform registers and for every election / vote, so that set-up is shared by
// the block's three instructions (SASS: ~35 instructions per block instead of ~70).  CG = 1: single CTA, 2: CTA pair.
template <int CG>
__device__ __forceinline__ void tc_mma3_e(uint32_t d_tmem, uint32_t a_hi_lo, uint32_t a_lo_lo, uint32_t b_hi_lo, uint32_t b_lo_lo, uint32_t desc_hi,
                                          uint32_t idesc, uint32_t acc0)
{
    if (CG == 2)
        asm volatile(
            "{\n\t.reg .pred p, q, e;\n\t.reg .b64 dah, dal, dbh, dbl;\n\t"
            "elect.sync _|e, 0xffffffff;\n\t"
            "mov.b64 dah, {%1, %5};\n\t"
            "mov.b64 dal, {%2, %5};\n\t"
            "mov.b64 dbh, {%3, %5};\n\t"
            "mov.b64 dbl, {%4, %5};\n\t"
            "setp.ne.b32 p, %7, 0;\n\t"
            "setp.eq.b32 q, %5, %5;\n\t"
            "@e tcgen05.mma.cta_group::2.kind::f16 [%0], dah, dbh, %6, p;\n\t"
            "@e tcgen05.mma.cta_group::2.kind::f16 [%0], dah, dbl, %6, q;\n\t"
            "@e tcgen05.mma.cta_group::2.kind::f16 [%0], dal, dbh, %6, q;\n\t}" ::"r"(d_tmem),
            "r"(a_hi_lo), "r"(a_lo_lo), "r"(b_hi_lo), "r"(b_lo_lo), "r"(desc_hi), "r"(idesc), "r"(acc0)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p, q, e;\n\t.reg .b64 dah, dal, dbh, dbl;\n\t"
            "elect.sync _|e, 0xffffffff;\n\t"
            "mov.b64 dah, {%1, %5};\n\t"
            "mov.b64 dal, {%2, %5};\n\t"
            "mov.b64 dbh, {%3, %5};\n\t"
            "mov.b64 dbl, {%4, %5};\n\t"
            "setp.ne.b32 p, %7, 0;\n\t"
            "setp.eq.b32 q, %5, %5;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, %6, p;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, %6, q;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], dal, dbh, %6, q;\n\t}" ::"r"(d_tmem),
            "r"(a_hi_lo), "r"(a_lo_lo), "r"(b_hi_lo), "r"(b_lo_lo), "r"(desc_hi), "r"(idesc), "r"(acc0)
            : "memory");
}
__device__ __forceinline__ void tc_commit_e2_both(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\t.reg .b16 m;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b16 m, 3;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_commit_e2_local(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
        : "memory");
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
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t *r)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// hi + lo += s, error-free (Knuth's two-sum): the epilogues accumulate their fp32 chunk sums as an unevaluated pair of floats
// (~48 bits) instead of in fp64 -- on this part a warp's DADD waits for the fp64 pipe long enough to show as a fifth of
// the kernel's issue stalls (ncu source page), and the pair costs seven FADDs on the idle fp32 pipe
__device__ __forceinline__ void two_sum_acc(float &hi, float &lo, float s)
{
    const float t = hi + s;
    const float bp = t - hi;
    lo += (hi - (t - bp)) + (s - bp);
    hi = t;
}
// The three MMAs of one K16 block of a CTA pair with the hi half of the A operand in TENSOR MEMORY (tcgen05.mma [d], [a], b-desc):
// hi * hi and hi * lo read A from TMEM, lo * hi reads the lo Hankel rows from shared memory.
__device__ __forceinline__ void tc_mma3_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t a_lo_lo, uint32_t b_hi_lo, uint32_t b_lo_lo, uint32_t desc_hi,
                                           uint32_t idesc, uint32_t acc0)
{
    asm volatile(
        "{\n\t.reg .pred p, q, e;\n\t.reg .b64 dal, dbh, dbl;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 dal, {%2, %5};\n\t"
        "mov.b64 dbh, {%3, %5};\n\t"
        "mov.b64 dbl, {%4, %5};\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.eq.b32 q, %5, %5;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], dbh, %6, p;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], dbl, %6, q;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], dal, dbh, %6, q;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(a_lo_lo), "r"(b_hi_lo), "r"(b_lo_lo), "r"(desc_hi), "r"(idesc), "r"(acc0)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
struct TcArgs {
    const int32_t *start;
    const int64_t *out_off, *bias_off;
    const int32_t *seq_start;
    const double *E;
    const double *emax;       // device scalar: max of E over the batch (bit pattern max, E > 0)
    const int4 *tab;          // per stage {first block, #blocks | flags << 8 (bit0 first of slab, bit1 last), bytes, image offset / 16}
    const int4 *blk;          // per K16 block {A descriptor advance (16 B units), first row n_lo, rows N_t, byte offset in the stage}
    const unsigned char *g_img;
    const double *t_row1;     // f_1 * V[1 - lv, :] (size-1 fragments: single tap)
    double *bx;
    unsigned long long *dbg;  // NB200_TC_DEBUG: per-CTA clock sums (8 words), else null
    int pwm_up, A0, B0, NAp, NBp, gmin, span, n_stages, n_achunks, n_blocks, sG, has_row1, W, w, epad, stagger;
    int n_chunks, tiles_per_chunk, ring;   // work items = n_chunks * tiles_per_chunk; ring = G stages resident in smem
    int slot_bytes;                        // bytes of one ring slot in this CTA's shared memory
    int res_bytes;                         // resident mode: bytes of this CTA's share of the whole G image
};

// Persistent kernel: one CTA per SM walks the work items (chunk, pair of x-tiles) it = blockIdx.x + i * gridDim.x.
// Four warp roles, all connected by mbarriers only (no CTA-wide barrier inside the item loop):
//   prep  (4 warps)  E window of item n+1 -> fp32 smem copy + fp16 hi/lo Hankel operand Z, double-buffered sets
//   prod  (1 thread) cp.async.bulk of the G stages through the ring (continues across items)
//   mma   (2 warps)  tcgen05.mma of x-tile j into its TMEM accumulator, N trimmed per K16 block
//   epi   (8 warps)  tcgen05.ld of a finished slab, contraction with E, bx store at the end of the item
// PAIR: the CTAs 2p, 2p + 1 form a cluster on one TPC and contract their x-tiles together with cta_group::2 MMAs of
// M = 256: the leader (cluster rank 0) issues for both, every CTA stages only half of the rows of each block of G (B is
// split across the pair's shared memories), so the shared-memory traffic per MMA and SM falls from A 4 KB + B 6 KB to
// A 4 KB + B 3 KB and G crosses L2 -> SMEM once per pair.  A work item of a pair is 4 x-tiles (512 outputs): rank r
// takes outputs [x0 + 256 r, x0 + 256 r + 256).  Barriers that gather both CTAs live in the leader: the peer arrives
// through the cluster address space (its two otherwise idle MMA warps relay "stage landed"), the leader's commits
// arrive on both CTAs at once (multicast).
// RES (with PAIR): a CTA's half of the whole G image fits its shared memory next to the operands (251 x 251: 154 KB), so it
// is loaded ONCE per kernel instead of being streamed from L2 for every work item: no producer loop, no ring, no "stage
// landed" waits.  With every block always at hand one issuing warp takes the slabs of the two x-tiles in turn --
// (tile 0, slab 0), (tile 1, slab 0), (tile 0, slab 1), ... -- so the read-back of a tile's accumulator runs under the other
// tile's MMAs.
template <bool DBG, bool PAIR, bool RES>
__global__ void __launch_bounds__(TC_THREADS, 1) k_nuc_bx_tc(TcArgs a)
{
    extern __shared__ __align__(128) unsigned char sm_tc[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;        // 0 = leader (issues the MMAs)
    const int it_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, it_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int item_w = PAIR ? 2 * TC_TX : TC_TX;                // outputs per work item
    const int x_rank = (int)rank * TC_TX;                       // this CTA's offset inside the item

    // ---- shared memory carve-up
    const int nZ = (TC_TX + a.NBp) / 8;                     // 128-byte chunks per Z part
    const int spanp = (a.epad + a.span + 31) & ~31;
    unsigned char *p = sm_tc;
    unsigned char *s_stage = p;            p += RES ? (size_t)a.res_bytes : (size_t)a.ring * a.slot_bytes;
    unsigned char *s_z = p;                p += (size_t)4 * nZ * 128;             // [set][hi|lo][nZ * 128]
    float *s_Eb = reinterpret_cast<float *>(p);             // [set][epad + span] E over genomic [g0 + gmin, ...), rounded to fp32
    p += sizeof(float) * 2 * spanp;                         //   (epad: the epilogue's 32-float runs start on 128-byte lines)
    // size-1 fragments have a single tap (a term linear in E): f_1 * V[1,:] as fp32 (zero padded), a 16-byte aligned
    // copy of the E window whose element x + k is tap k of output x, and the finished term per output  [only if f_1 V[1,:] != 0]
    const int Wp = (a.W + 3) & ~3, e1p = TC_TX + Wp + 8;
    float *s_t1 = reinterpret_cast<float *>(p);             p += sizeof(float) * (a.has_row1 ? Wp : 0);
    float *s_E1b = reinterpret_cast<float *>(p);            p += sizeof(float) * (a.has_row1 ? 2 * e1p : 0);
    float *s_linb = reinterpret_cast<float *>(p);           p += sizeof(float) * (a.has_row1 ? 2 * TC_TX : 0);
    int4 *s_tab = reinterpret_cast<int4 *>(p);              // stage + block tables (kept out of the issue loop's global-load latency)
    p += sizeof(int4) * a.n_stages;
    int4 *s_blk = reinterpret_cast<int4 *>(p);
    p += sizeof(int4) * (a.n_blocks + 1);                   // one spare entry: the resident issue loop reads one block ahead
    int *s_soff = reinterpret_cast<int *>(p);               // resident mode: byte offset of every stage's image in s_stage
    p += sizeof(int) * ((a.n_stages + 3) & ~3);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(p);      // full[ring], empty[ring], tfull[tile], tempty[tile], zfull[2], zempty[2], stag, zpair[2]
    p += sizeof(uint64_t) * (2 * TC_MAX_STAGES + 2 * TC_XT + 4 + 1 + 2);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(p);
    const uint32_t bar_full = smem_u32(s_bar), bar_empty = bar_full + 8 * TC_MAX_STAGES, bar_tfull = bar_empty + 8 * TC_MAX_STAGES,
                   bar_tempty = bar_tfull + 8 * TC_XT, bar_zfull = bar_tempty + 8 * TC_XT, bar_zempty = bar_zfull + 16,
                   bar_stag = bar_zempty + 16, bar_zpair = bar_stag + 8;   // zpair (leader): the operands of BOTH CTAs are ready

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.ring; i++) {
            mbar_init(bar_full + 8 * i, (PAIR && rank == 0) ? 2 : 1);   // own bulk copy (+ the peer's "landed" relay)
            mbar_init(bar_empty + 8 * i, TC_XT);                // one commit per MMA warp
        }
        for (int i = 0; i < TC_XT; i++) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, PAIR ? 8 : 4);        // the 4 epilogue warps of the tile (of both CTAs)
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(bar_zfull + 8 * i, TC_PREP_WARPS);
            mbar_init(bar_zempty + 8 * i, TC_EPI_WARPS + (RES ? 1 : TC_XT));   // epilogue warps (E window) + the MMA warps' commits (Z)
            mbar_init(bar_zpair + 8 * i, 2 * TC_PREP_WARPS);
        }
        mbar_init(bar_stag, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_WARP_MMA) {  // TMEM: all 512 columns (one CTA per SM by shared-memory footprint)
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(TC_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(TC_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (a.has_row1)
        for (int i = threadIdx.x; i < Wp; i += TC_THREADS) s_t1[i] = (i < a.W) ? (float)a.t_row1[i] : 0.f;
    for (int i = threadIdx.x; i < a.n_stages; i += TC_THREADS) s_tab[i] = a.tab[i];
    for (int i = threadIdx.x; i <= a.n_blocks; i += TC_THREADS) s_blk[i] = (i < a.n_blocks) ? a.blk[i] : make_int4(0, 0, 0, 0);
    if (RES && threadIdx.x == 0) {
        int off = 0;
        for (int i = 0; i < a.n_stages; i++) {
            s_soff[i] = off;
            off += a.tab[i].z >> 1;
        }
    }
    int eexp = 1;
    const double emax = a.emax[0];
    if (emax > 0.0) frexp(32768.0 / emax, &eexp);
    const int sE = eexp - 1;
    tc_fence_before();
    if (PAIR)
        cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
    else
        __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    if (DBG && threadIdx.x == 0 && blockIdx.x < 2 && g_tc_trap_info) {   // where the barriers of CTA 0 / 1 live, for reading a trap record
        g_tc_trap_info[8 + 2 * blockIdx.x] = (int)bar_full;
        g_tc_trap_info[9 + 2 * blockIdx.x] = (int)tmem;
    }
    const long long t_begin = DBG ? clock64() : 0;
    unsigned long long *dbg = DBG ? a.dbg + 8 * (size_t)blockIdx.x : nullptr;
    const int n_items = a.n_chunks * a.tiles_per_chunk;

    if (warp >= TC_WARP_PREP) {
        // ===== operand generation: E window (fp32) and Z[t][s][0..7] = fp16 hi/lo of 2^sE * E[b-window start + 8t + s + j]
        const int tid = threadIdx.x - 32 * TC_WARP_PREP;
        const float scale = (float)ldexp(1.0, sE);
        const int boff = a.B0 - a.gmin;  // s_E index of the first b tap of output 0
        long long w_ze = 0;
        int n = 0;
        for (int it = it_first; it < n_items; it += it_step) {
            const int c = it / a.tiles_per_chunk, xb = (it - c * a.tiles_per_chunk) * item_w, x0 = xb + x_rank;
            const int64_t oo = a.out_off[c];
            if (xb >= (int)(a.out_off[c + 1] - oo)) continue;
            const int set = n & 1;
            const long long t0 = DBG ? clock64() : 0;
            mbar_wait(bar_zempty + 8 * set, ((n >> 1) & 1) ^ 1);
            if (DBG) w_ze += clock64() - t0;
            float *s_E = s_Eb + (size_t)set * spanp + a.epad;
            const int64_t e_lo = a.bias_off[c], e_hi = a.bias_off[c + 1];
            const int64_t ebase = e_lo - (int64_t)(a.seq_start[c] + a.pwm_up) + (int64_t)a.start[c] + x0 + a.gmin;
#pragma unroll 4
            for (int i = tid; i < a.span; i += TC_PREP_THREADS) {   // out of track -> 0 (only under zero columns of G / unused outputs)
                const int64_t idx = ebase + i;
                s_E[i] = (idx >= e_lo && idx < e_hi) ? (float)a.E[idx] : 0.f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PREP_THREADS) : "memory");
            unsigned char *zh = s_z + (size_t)set * 2 * nZ * 128, *zl = zh + (size_t)nZ * 128;
            for (int e = tid; e < nZ * 8; e += TC_PREP_THREADS) {
                __align__(16) __half hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int idx = boff + e + j;                       // = boff + 8t + s + j with e = 8t + s
                    const float v = (idx < a.span) ? s_E[idx] * scale : 0.f;   // power-of-two scale: exact
                    hi[j] = __float2half_rn(v);
                    lo[j] = __float2half_rn(v - __half2float(hi[j]));
                }
                *reinterpret_cast<uint4 *>(zh + (size_t)e * 16) = *reinterpret_cast<uint4 *>(hi);
                *reinterpret_cast<uint4 *>(zl + (size_t)e * 16) = *reinterpret_cast<uint4 *>(lo);
            }
            if (PAIR)
                asm volatile("fence.proxy.async;" ::: "memory");              // ... also to the leader's MMAs reading this CTA's memory
            else
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            if (a.has_row1) {
                // insert size 1: Bp[1,c] = E[c] (one tap) -> lin[x] = sum_k t1[k] E[x - w + k], a plain 1-D correlation.  Two warps,
                // 4 consecutive outputs per lane, 4 taps per step from one aligned 16-byte load (64 FMAs per smem wavefront).
                float *s_E1 = s_E1b + (size_t)set * e1p;
                const int sh = -a.w - a.gmin;                    // >= 0
                for (int i = tid; i < e1p; i += TC_PREP_THREADS) {
                    const int idx = i + sh;
                    s_E1[i] = (idx < a.span) ? s_E[idx] : 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(TC_PREP_THREADS) : "memory");
                if (tid < TC_TX / 4) {
                    const float4 *e4 = reinterpret_cast<const float4 *>(s_E1 + 4 * tid);
                    const float4 *t4 = reinterpret_cast<const float4 *>(s_t1);
                    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                    float4 e0 = e4[0];
                    for (int k4 = 0; k4 < Wp / 4; k4++) {
                        const float4 t = t4[k4], e1 = e4[k4 + 1];
                        l0 = fmaf(t.x, e0.x, fmaf(t.y, e0.y, fmaf(t.z, e0.z, fmaf(t.w, e0.w, l0))));
                        l1 = fmaf(t.x, e0.y, fmaf(t.y, e0.z, fmaf(t.z, e0.w, fmaf(t.w, e1.x, l1))));
                        l2 = fmaf(t.x, e0.z, fmaf(t.y, e0.w, fmaf(t.z, e1.x, fmaf(t.w, e1.y, l2))));
                        l3 = fmaf(t.x, e0.w, fmaf(t.y, e1.x, fmaf(t.z, e1.y, fmaf(t.w, e1.z, l3))));
                        e0 = e1;
                    }
                    *reinterpret_cast<float4 *>(s_linb + (size_t)set * TC_TX + 4 * tid) = make_float4(l0, l1, l2, l3);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_zfull + 8 * set);
                if (PAIR) mbar_arrive_cluster(bar_zpair + 8 * set, 0);
            }
            n++;
        }
        if (DBG && tid == 0) dbg[6] = (unsigned long long)w_ze;
    } else if (warp == TC_WARP_PROD) {
        // ===== producer: stream the G stages (hi + lo images of the trimmed K16 blocks) through the ring, once per item =====
        if (RES) {
            if (lane == 0) {   // this CTA's half of every stage image, once
                mbar_expect_tx(bar_full, (uint32_t)a.res_bytes);
                for (int s = 0; s < a.n_stages; s++) {
                    const int4 st = s_tab[s];
                    const uint32_t bytes = (uint32_t)st.z >> 1;
                    bulk_g2s(smem_u32(s_stage + s_soff[s]), a.g_img + (size_t)st.w * 16 + (size_t)rank * bytes, bytes, bar_full);
                }
            }
        } else if (lane == 0) {
            const unsigned char *g_img = a.g_img;
            int slot = 0;
            uint32_t ph = 0;
            for (int it = it_first; it < n_items; it += it_step) {
                const int c = it / a.tiles_per_chunk, xb = (it - c * a.tiles_per_chunk) * item_w;
                if (xb >= (int)(a.out_off[c + 1] - a.out_off[c])) continue;
                for (int s = 0; s < a.n_stages; s++) {
                    const int4 st = s_tab[s];
                    const uint32_t bytes = PAIR ? (uint32_t)st.z >> 1 : (uint32_t)st.z;   // a pair's stage image: rank 0's half, then rank 1's
                    mbar_wait(bar_empty + 8 * slot, ph ^ 1);
                    mbar_expect_tx(bar_full + 8 * slot, bytes);
                    bulk_g2s(smem_u32(s_stage + (size_t)slot * a.slot_bytes), g_img + (size_t)st.w * 16 + (size_t)rank * bytes, bytes, bar_full + 8 * slot);
                    if (++slot == a.ring) {
                        slot = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (PAIR && rank != 0 && warp >= TC_WARP_MMA) {
        // ===== peer of a pair: its MMA warps issue nothing; warp 0 of them tells the leader when this CTA's half of a stage has landed
        if (RES) {
            if (warp == TC_WARP_MMA) {
                mbar_wait(bar_full, 0);
                if (lane == 0) mbar_arrive_cluster(bar_full, 0);
                __syncwarp();
            }
        } else if (warp == TC_WARP_MMA) {
            int slot = 0;
            uint32_t ph = 0;
            for (int it = it_first; it < n_items; it += it_step) {
                const int c = it / a.tiles_per_chunk, xb = (it - c * a.tiles_per_chunk) * item_w;
                if (xb >= (int)(a.out_off[c + 1] - a.out_off[c])) continue;
                for (int s = 0; s < a.n_stages; s++) {
                    mbar_wait(bar_full + 8 * slot, ph);
                    if (lane == 0) mbar_arrive_cluster(bar_full + 8 * slot, 0);
                    __syncwarp();
                    if (++slot == a.ring) {
                        slot = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (RES && warp >= TC_WARP_MMA) {
        // ===== resident G: one issuing warp (of the leader), slabs of the two x-tiles in turn
        if (warp == TC_WARP_MMA) {
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);
            const uint32_t sbase = (smem_u32(s_stage) >> 4) & 0x3FFF;
            long long w_zf = 0, w_te = 0;
            int n = 0, gq0 = 0, gq1 = 0;
            mbar_wait_cluster(bar_full, 0);   // both CTAs' halves of G have landed
            tc_fence_after();
            for (int it = it_first; it < n_items; it += it_step) {
                const int c = it / a.tiles_per_chunk, xb = (it - c * a.tiles_per_chunk) * item_w;
                if (xb >= (int)(a.out_off[c + 1] - a.out_off[c])) continue;
                const int set = n & 1;
                {
                    const long long t0 = DBG ? clock64() : 0;
                    mbar_wait_cluster(bar_zpair + 8 * set, (n >> 1) & 1);
                    tc_fence_after();
                    if (DBG) w_zf += clock64() - t0;
                }
                const uint32_t zb = smem_u32(s_z + (size_t)set * 2 * nZ * 128);
                for (int s0 = 0; s0 < a.n_stages;) {
                    int s1 = s0;
                    while (!((s_tab[s1].y >> 9) & 1)) s1++;   // last stage of this slab
#pragma unroll
                    for (int j = 0; j < TC_XT; j++) {
                        const uint32_t zj = zb + 2048u * j;
                        const uint32_t a_hi0 = ((zj >> 4) & 0x3FFF) | ((128u >> 4) << 16);
                        const uint32_t a_lo0 = (((zj + (uint32_t)nZ * 128u) >> 4) & 0x3FFF) | ((128u >> 4) << 16);
                        const uint32_t d0 = tmem + (uint32_t)(j * TC_N);
                        {
                            const long long t0 = DBG ? clock64() : 0;
                            mbar_wait_cluster(bar_tempty + 8 * j, ((j ? gq1 : gq0) & 1) ^ 1);
                            tc_fence_after();
                            if (DBG) w_te += clock64() - t0;
                        }
                        uint32_t acc0 = 0u;   // the first block of a slab is stored untrimmed: it initialises all TC_N columns
                        for (int s = s0; s <= s1; s++) {
                            const int4 st = s_tab[s];
                            const int b0 = st.x, nblk = st.y & 0xff;
                            const uint32_t sb = sbase + ((uint32_t)s_soff[s] >> 4);
                            int4 bk = s_blk[b0];
                            for (int t = 0; t < nblk; t++) {
                                const int4 nx = s_blk[b0 + t + 1];   // next entry in flight while this block is issued (the table has a spare entry)
                                const uint32_t bh = (uint32_t)bk.w + sb;
                                const uint32_t bl = bh + 2u * ((uint32_t)bk.w >> 16);
                                tc_mma3_e<2>(d0 + (uint32_t)bk.y, a_hi0 + (uint32_t)bk.x, a_lo0 + (uint32_t)bk.x, bh, bl, desc_hi, (uint32_t)bk.z, acc0);
                                acc0 = 1u;
                                bk = nx;
                            }
                        }
                        tc_commit_e2_both(bar_tfull + 8 * j);   // slab of H complete in both CTAs' TMEM
                        if (j)
                            gq1++;
                        else
                            gq0++;
                    }
                    s0 = s1 + 1;
                }
                tc_commit_e2_both(bar_zempty + 8 * set);        // Z set free once every MMA of the item has retired
                n++;
            }
            if (DBG && lane == 0) {
                dbg[1] = (unsigned long long)w_zf;
                dbg[2] = (unsigned long long)w_te;
                dbg[3] = 0;
                dbg[7] = (unsigned long long)n;
            }
        }
    } else if (warp >= TC_WARP_MMA) {
        // ===== MMA issuers: warp TC_WARP_MMA + j contracts x-tile j.  The whole warp runs the loop (convergent); one elected
        // lane issues.  Two issuing warps keep two independent accumulation streams in the tensor pipe.
        const int j = warp - TC_WARP_MMA;
        // descriptor words: low = start address >> 4 | (LBO >> 4) << 16, high = SBO >> 4 | version 1 << 14
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);                                // SBO = 128 B for both operands
        const uint32_t d0 = tmem + (uint32_t)(j * TC_N);
        long long w_zf = 0, w_te = 0, w_full = 0;
        int n = 0, gq = 0, slot = 0;
        uint32_t ph = 0;
        for (int it = it_first; it < n_items; it += it_step) {
            const int c = it / a.tiles_per_chunk, xb = (it - c * a.tiles_per_chunk) * item_w;
            if (xb >= (int)(a.out_off[c + 1] - a.out_off[c])) continue;
            const int set = n & 1;
            {
                const long long t0 = DBG ? clock64() : 0;
                if (PAIR)
                    mbar_wait_cluster(bar_zpair + 8 * set, (n >> 1) & 1);   // both CTAs' operands
                else
                    mbar_wait(bar_zfull + 8 * set, (n >> 1) & 1);
                tc_fence_after();
                if (DBG) w_zf += clock64() - t0;
            }
            const uint32_t zb = smem_u32(s_z + (size_t)set * 2 * nZ * 128) + 2048u * j;      // x-tile j: 128 elements = 16 chunks of 128 B on
            const uint32_t a_hi0 = ((zb >> 4) & 0x3FFF) | ((128u >> 4) << 16);            // Hankel view: LBO = SBO = 128 B
            const uint32_t a_lo0 = (((zb + (uint32_t)nZ * 128u) >> 4) & 0x3FFF) | ((128u >> 4) << 16);
            // The two tiles share every G stage but each has a single accumulator: if they finished their slabs together the
            // tensor pipe would idle while both are read back.  Tile 1 therefore starts a few stages behind tile 0 (once; the
            // offset persists because both streams have the same period), so one tile is contracted while the other drains.
            if (n == 0 && j == 1 && a.stagger >= 0) mbar_wait(bar_stag, 0);
            for (int s = 0; s < a.n_stages; s++) {
                const int4 st = s_tab[s];
                const int b0 = st.x, nblk = st.y & 0xff;
                const bool first = (st.y >> 8) & 1, last = (st.y >> 9) & 1;
                if (first) {  // the accumulator of this tile is free once its previous slab has been read back
                    const long long t0 = DBG ? clock64() : 0;
                    if (PAIR)
                        mbar_wait_cluster(bar_tempty + 8 * j, (gq & 1) ^ 1);
                    else
                        mbar_wait(bar_tempty + 8 * j, (gq & 1) ^ 1);
                    tc_fence_after();
                    if (DBG) w_te += clock64() - t0;
                }
                const long long t1 = DBG ? clock64() : 0;
                if (PAIR)
                    mbar_wait_cluster(bar_full + 8 * slot, ph);
                else
                    mbar_wait(bar_full + 8 * slot, ph);
                tc_fence_after();
                if (DBG) w_full += clock64() - t1;
                const uint32_t sb = (smem_u32(s_stage + (size_t)slot * a.slot_bytes) >> 4) & 0x3FFF;
                uint32_t acc0 = first ? 0u : 1u;   // the first block of a slab is stored untrimmed: it initialises all TC_N columns
                for (int t = 0; t < nblk; t++) {
                    const int4 bk = s_blk[b0 + t];   // {A advance, n_lo, instruction descriptor, B descriptor low word (relative)}
                    const uint32_t bh = (uint32_t)bk.w + sb;                       // hi image: N_t rows, LBO = N_t * 16 B
                    const uint32_t bl = bh + 2u * ((uint32_t)bk.w >> 16);           // lo image follows (N_t * 32 B)
                    const uint32_t ah = a_hi0 + (uint32_t)bk.x, al = a_lo0 + (uint32_t)bk.x;
                    const uint32_t d = d0 + (uint32_t)bk.y;
                    if (PAIR)
                        tc_mma3_e<2>(d, ah, al, bh, bl, desc_hi, (uint32_t)bk.z, acc0);   // hi * hi, hi * lo, lo * hi
                    else
                        tc_mma3_e<1>(d, ah, al, bh, bl, desc_hi, (uint32_t)bk.z, acc0);
                    acc0 = 1u;
                }
                if (PAIR) {
                    tc_commit_e2_both(bar_empty + 8 * slot);           // both CTAs' slots reusable once these MMAs retire
                    if (n == 0 && j == 0 && s == a.stagger) tc_commit_e2_local(bar_stag);
                    if (last) {
                        tc_commit_e2_both(bar_tfull + 8 * j);          // slab of H complete in both CTAs' TMEM
                        gq++;
                    }
                } else {
                    tc_commit_e(bar_empty + 8 * slot);                 // smem slot reusable once these MMAs retire
                    if (n == 0 && j == 0 && s == a.stagger) tc_commit_e(bar_stag);
                    if (last) {
                        tc_commit_e(bar_tfull + 8 * j);                // slab of H complete in TMEM
                        gq++;
                    }
                }
                if (++slot == a.ring) {
                    slot = 0;
                    ph ^= 1;
                }
            }
            if (PAIR)
                tc_commit_e2_both(bar_zempty + 8 * set);
            else
                tc_commit_e(bar_zempty + 8 * set);                     // Z set free once every MMA of the item has retired
            n++;
        }
        if (DBG && j == 0 && lane == 0) {
            dbg[1] = (unsigned long long)w_zf;
            dbg[2] = (unsigned long long)w_te;
            dbg[3] = (unsigned long long)w_full;
            dbg[7] = (unsigned long long)n;
        }
    } else {
        // ===== epilogue: bx[x] = sum_n E[x + A0 + n] * H[x, n]; warps 4j..4j+3 own x-tile j (one TMEM lane quarter each)
        const int j = warp >> 2, wq = warp & 3;
        const int m = wq * 32 + lane;                          // TMEM lane = output position within the x-tile
        const int aoff = a.A0 - a.gmin;
        const double unscale = ldexp(1.0, -(sE + a.sG));
        const uint32_t t0addr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(j * TC_N);
        long long w_tf = 0, t_epi = 0;
        int n = 0, gq = 0;
        for (int it = it_first; it < n_items; it += it_step) {
            const int c = it / a.tiles_per_chunk, xb = (it - c * a.tiles_per_chunk) * item_w, x0 = xb + x_rank;
            const int64_t oo = a.out_off[c];
            const int L = (int)(a.out_off[c + 1] - oo);
            if (xb >= L) continue;
            const int set = n & 1;
            mbar_wait(bar_zfull + 8 * set, (n >> 1) & 1);
            const float *s_E = s_Eb + (size_t)set * spanp + a.epad;
            const double lin = a.has_row1 ? (double)s_linb[(size_t)set * TC_TX + TC_M * j + m] : 0.0;  // size-1 term (prep warps)
            // H (fp32 from TMEM) x E (fp32): runs of 8 products are summed in fp32 (4 independent chains per 32 columns), the
            // four runs of a chunk in fp32 too, the chunks as an error-free float pair -- rounding ~1e-7 of a chunk of positive
            // terms, below the fp16-split error of H itself (no fp64 in the loop: see two_sum_acc)
            float acc_hi = 0.f, acc_lo = 0.f;
            for (int q = 0; q < a.n_achunks; q++, gq++) {
                const long long t0 = DBG ? clock64() : 0;
                mbar_wait(bar_tfull + 8 * j, gq & 1);
                tc_fence_after();
                const long long t1 = DBG ? clock64() : 0;
                if (DBG) w_tf += t1 - t0;
                const float *Ew = s_E + aoff + TC_M * j + m + TC_N * q;   // lane 0: a multiple of 32 floats from a 128-byte line
                // two register sets: the tcgen05.ld of columns 32(h+1).. is in flight while columns 32h.. are contracted;
                // the accumulator is handed back to the MMA warp as soon as its last columns have landed in registers
                uint32_t r[2][32];
                tc_ld32(t0addr, r[0]);
#pragma unroll
                for (int h = 0; h < TC_N / 32; h++) {
                    tc_wait_ld();
                    if (h + 1 < TC_N / 32) {
                        tc_ld32(t0addr + (uint32_t)(32 * (h + 1)), r[(h + 1) & 1]);
                    } else {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (PAIR)
                                mbar_arrive_cluster_relaxed(bar_tempty + 8 * j, 0);   // the leader's barrier gathers both CTAs
                            else
                                mbar_arrive(bar_tempty + 8 * j);
                        }
                    }
                    float f[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int nn = 0; nn < 32; nn++) f[nn >> 3] = fmaf(__uint_as_float(r[h & 1][nn]), Ew[32 * h + nn], f[nn >> 3]);
                    two_sum_acc(acc_hi, acc_lo, (f[0] + f[1]) + (f[2] + f[3]));
                }
                if (DBG) t_epi += clock64() - t1;
            }
            const int x = x0 + TC_M * j + m;
            if (x < L) a.bx[oo + x] = ((double)acc_hi + (double)acc_lo) * unscale + lin;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_zempty + 8 * set);      // this warp no longer reads the E window of the set
            n++;
        }
        if (DBG && threadIdx.x == 0) {
            dbg[4] = (unsigned long long)w_tf;
            dbg[5] = (unsigned long long)t_epi;
            dbg[0] = (unsigned long long)(clock64() - t_begin);
        }
    }
    tc_fence_before();
    if (PAIR)
        cluster_sync_all();   // no CTA leaves (or frees its TMEM) while its partner's MMAs / arrives may still touch it
    else
        __syncthreads();
    if (warp == TC_WARP_MMA) {
        tc_fence_after();
        if (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
    }
}

// =============================================================================================
// k_nuc_bx_ts -- the same contraction with the hi half of the Hankel operand in TENSOR MEMORY
// =============================================================================================
// What bounds k_nuc_bx_tc (measured with a stand-alone issue loop, profiles/r2b_mma_microbench.txt): the 128 B/clk shared-memory
// pipe of an SM is shared by the MMAs' operand fetches -- which win -- and the epilogue's loads of E.  An MMA with both
// operands in shared memory costs max(N/2, (4096 + 16 N)/128) clk in a pair (the A tile alone is 32 clk of the pipe), and
// while MMAs of N <= 128 run the epilogue gets 11-23 B/clk of E, so a slab drains more slowly than the next one is
// contracted.  Two of the three MMAs of a block read the SAME hi rows: with those in tensor memory (tcgen05.mma [d], [a],
// b-desc) they fetch only G, MMAs run at N/2 clk down to N = 32 and the epilogue keeps 55-70 B/clk.
//
// TMEM (512 columns): two accumulators of TS_N = 128 columns + the expanded hi operand of ONE x-tile, A[m][k] = Es_hi[m + k]
// as packed halves (column 256 + k/2; NBp/2 <= 192 columns at 251 x 251).  The x-tiles of a CTA are therefore contracted
// one after the other, slab by slab through the two accumulators (unit u -> accumulator u & 1); the two groups of four
// epilogue warps take alternate units and add their partial sums per tile through shared memory.  The operand is
// written by the operand-generation warps (tcgen05.st from the hi Hankel rows they produced in shared memory) in two
// parts: the K blocks the tile's last slab does not read as soon as the previous tile's last MMA that reads them has
// retired (a commit a whole slab before its end), the rest when the previous tile is complete -- under the part-1 blocks
// of slab 0 of the new tile, which are issued first.  CTA pairs, resident G and the fused issue of a block's three MMAs are as in k_nuc_bx_tc<.., true, true>.
#ifndef TS_N
#define TS_N 128                          // slab width = accumulator columns
#define TS_NACC 2                         // accumulators the slabs go round (measured: 128 x 2 6.16 ms, 96 x 3 6.53 ms per 20 Mbp)
#endif
#ifndef TS_FIN32
#define TS_FIN32 1                        // the last additions of an output in fp32 too (A/B in one run: 5.64 -> 5.54 ms; with TS_EPI_BUF 3: 5.49)
#endif
#ifndef TS_EPI_BUF
#define TS_EPI_BUF 3                      // 32-register sets an epilogue thread reads a slab through
#endif
#define TS_ACOL (TS_NACC * TS_N)          // first TMEM column of the hi operand
#define TS_MAXCOL (TC_TMEM_COLS - TS_ACOL)
#define TS_BARS 23                        // gfull, zfull[2], zempty[2], zpair[2], tfull[3], tempty[3], afull[2], afree[2], linfull[2], pfull[4]
// 13 warps (152 registers per thread, so that an epilogue thread holds a whole 128-column slab and hands the accumulator back
// as soon as its four tcgen05.ld have landed): 0-7 epilogue, 8 issuing warp (+ the one-off load of G), 9-12 operand
// generation and expansion into tensor memory (warp & 3 = 1, 2, 3, 0: the four TMEM lane quarters)
#define TS_WARP_MMA TC_EPI_WARPS
#define TS_WARP_PREP (TC_EPI_WARPS + 1)
#define TS_THREADS (32 * (TC_EPI_WARPS + 1 + TC_PREP_WARPS))

struct TsArgs {
    const int32_t *start;
    const int64_t *out_off, *bias_off;
    const int32_t *seq_start;
    const double *E;
    const double *emax;
    const unsigned char *g_img;   // [rank]: that CTA's half of every block image (hi image, then lo image)
    const double *t_row1;
    double *bx;
    unsigned long long *dbg;
    int pwm_up, A0, B0, NBp, gmin, span, n_slabs, n_blocks, sG, has_row1, W, w, epad;
    int n_chunks, tiles_per_chunk;
    int rank_bytes;           // bytes of one CTA's image
    int c_split, c_end;       // hi-operand columns of part 1 / in all (multiples of 32)
    int q_need2, t_need2;     // the first block (in issue order) that reads part 2 of the operand: its slab and its position in the slab
    int nZ;                   // 128-byte chunks per Hankel part
};

// The block list travels as a kernel parameter (constant bank): with its entries and every loop counter of the issuing warp
// warp-uniform, ptxas keeps the whole issue loop on the uniform datapath (LDCU / UIADD3 / UMOV / UTCHMMA, no election, no
// R2UR) -- ~15 instructions per block instead of ~45, which matters because a block trimmed to N <= 96 occupies the tensor
// pipe for less time than the 45-instruction form takes to issue.
#define TS_MAX_BLOCKS 192
#define TS_MAX_SLABS 8
struct TsTab {
    int4 blk[TS_MAX_BLOCKS];   // per K16 block {K block | flags << 16 (bit 0: part 1 of the hi operand is read no more after it, bit 2: the first block that reads part 2), first row n_lo, instruction descriptor, image offset / 16 | rows per CTA << 16}
    int2 slab[TS_MAX_SLABS];   // per slab {first block, blocks}
};

template <bool DBG>
__global__ void __launch_bounds__(TS_THREADS, 1) k_nuc_bx_ts(const __grid_constant__ TsTab tab, const __grid_constant__ TsArgs a)
{
    extern __shared__ __align__(128) unsigned char sm_tc[];
    // the warp index through a lane-0 shuffle: the value the compiler can prove warp-uniform (the branch on it is what lets
    // the issuing warp's loop stay on the uniform datapath)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = blockIdx.x & 1u;                      // = %cluster_ctarank for clusters of (2, 1, 1), but provably warp-uniform; 0 = leader
    const int it_first = (int)(blockIdx.x >> 1), it_step = (int)(gridDim.x >> 1);
    const int item_w = 2 * TC_TX;                               // outputs per work item of the pair
    const int x_rank = (int)rank * TC_TX;

    // ---- shared memory carve-up
    const int nZ = a.nZ;
    // The fp32 E window of an item is kept FOUR times, copy s shifted by s elements (cp_s[i] = E[i + s]): whatever a thread's
    // first element, one of the copies has it on a 16-byte boundary, so the epilogue reads its 32 consecutive values per chunk
    // with 8 LDS.128 instead of 32 LDS.32 (with every register holding accumulator columns the loads cannot be batched, and
    // the one-at-a-time LDS latency was the epilogue's whole cost).  Copy stride = 8 mod 32 words: the 8 lanes of a
    // quarter-warp (4 copies x 2 consecutive 16-byte groups) fall into 8 different bank quads.
    const int cplen = ((a.span + 3 + 31) & ~31) + 8;
    unsigned char *p = sm_tc;
    unsigned char *s_stage = p;            p += (size_t)a.rank_bytes;
    unsigned char *s_zhi = p;              p += (size_t)nZ * 128;                 // hi Hankel rows: only the source of the expansion into TMEM (one set)
    unsigned char *s_zlo = p;              p += (size_t)2 * nZ * 128;             // [set][nZ * 128] lo Hankel rows (MMA operand)
    float *s_cpb = reinterpret_cast<float *>(p);            p += sizeof(float) * 2 * 4 * cplen;   // [set][copy][cplen]
    const int Wp = (a.W + 3) & ~3;
    float *s_t1 = reinterpret_cast<float *>(p);             p += sizeof(float) * (a.has_row1 ? Wp : 0);
    float *s_linb = reinterpret_cast<float *>(p);           p += sizeof(float) * (a.has_row1 ? 2 * TC_TX : 0);
    double *s_part = reinterpret_cast<double *>(p);         p += sizeof(double) * 4 * TC_M;       // [tile & 3][128]: the partial sums of the group that does not finish the tile (a double or a float pair)
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(p);      p += sizeof(uint64_t) * (TS_BARS + 1);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(p);
    const uint32_t bar_gfull = smem_u32(s_bar), bar_zfull = bar_gfull + 8, bar_zempty = bar_zfull + 16, bar_zpair = bar_zempty + 16,
                   bar_tfull = bar_zpair + 16, bar_tempty = bar_tfull + 8 * TS_NACC, bar_afull = bar_tempty + 8 * TS_NACC, bar_afree = bar_afull + 16,
                   bar_linfull = bar_afree + 16, bar_pfull = bar_linfull + 16;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        mbar_init(bar_gfull, rank == 0 ? 2 : 1);                // own bulk copies (+ the peer's "landed" relay)
        for (int i = 0; i < 2; i++) {
            mbar_init(bar_zfull + 8 * i, TC_PREP_WARPS);
            mbar_init(bar_zempty + 8 * i, TC_EPI_WARPS + 1);    // epilogue warps (E window) + the commit of the item's MMAs (lo rows)
            mbar_init(bar_zpair + 8 * i, 2 * TC_PREP_WARPS);
            mbar_init(bar_afull + 8 * i, 2 * TC_PREP_WARPS);    // part i of the hi operand written, in both CTAs
            mbar_init(bar_afree + 8 * i, 1);
            mbar_init(bar_linfull + 8 * i, TC_PREP_WARPS);      // the size-1 term of the item
        }
        for (int i = 0; i < 4; i++) mbar_init(bar_pfull + 8 * i, 4);   // the 4 warps of the depositing group
        for (int i = 0; i < TS_NACC; i++) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 8);                   // the 4 warps of the group that reads the slab, in both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TS_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (a.has_row1)
        for (int i = threadIdx.x; i < Wp; i += TS_THREADS) s_t1[i] = (i < a.W) ? (float)a.t_row1[i] : 0.f;
    int eexp = 1;
    const double emax = a.emax[0];
    if (emax > 0.0) frexp(32768.0 / emax, &eexp);
    const int sE = eexp - 1;
    tc_fence_before();
    cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    if (DBG && threadIdx.x == 0 && blockIdx.x < 2 && g_tc_trap_info) {
        g_tc_trap_info[8 + 2 * blockIdx.x] = (int)bar_gfull;
        g_tc_trap_info[9 + 2 * blockIdx.x] = (int)tmem;
    }
    const long long t_begin = DBG ? clock64() : 0;
    unsigned long long *dbg = DBG ? a.dbg + 16 * (size_t)blockIdx.x : nullptr;
    const int n_items = a.n_chunks * a.tiles_per_chunk;
    auto item_valid = [&](int it) {
        const int c = it / a.tiles_per_chunk, xb = (it - c * a.tiles_per_chunk) * item_w;
        return xb < (int)(a.out_off[c + 1] - a.out_off[c]);
    };
    auto next_valid = [&](int it) {
        while (it < n_items && !item_valid(it)) it += it_step;
        return it;
    };

    if (warp >= TS_WARP_PREP) {
        // ===== operand generation (for the next item) and the hi operand of every tile into tensor memory
        const int tid = threadIdx.x - 32 * TS_WARP_PREP;
        const int wq = warp & 3, m = wq * 32 + lane;          // the TMEM lane quarter this warp may touch; row of the tile
        const float scale = (float)ldexp(1.0, sE);
        const int boff = a.B0 - a.gmin;
        long long w_ze = 0, w_af = 0, t_prep = 0, t_fill = 0, t_eload = 0, t_lin = 0;
        auto prep_item = [&](int it, int n) {
            const int c = it / a.tiles_per_chunk, xb = (it - c * a.tiles_per_chunk) * item_w, x0 = xb + x_rank;
            const int set = n & 1;
            // the E window first into registers (the loads do not depend on the set being free), then wait for the set
            const int64_t e_lo = a.bias_off[c], e_hi = a.bias_off[c + 1];
            const int64_t ebase = e_lo - (int64_t)(a.seq_start[c] + a.pwm_up) + (int64_t)a.start[c] + x0 + a.gmin;
            constexpr int EPF = 8;                               // covers spans up to 8 * 128 = 1024 positions; longer ones in a second round
            float ev[EPF];
#pragma unroll
            for (int k = 0; k < EPF; k++) {
                const int i = tid + k * TC_PREP_THREADS;
                const int64_t idx = ebase + i;
                ev[k] = (i < a.span && idx >= e_lo && idx < e_hi) ? (float)a.E[idx] : 0.f;   // out of track / past the window -> 0
            }
            const long long t0 = DBG ? clock64() : 0;
            mbar_wait(bar_zempty + 8 * set, ((n >> 1) & 1) ^ 1);
            const long long tp0 = DBG ? clock64() : 0;
            if (DBG) w_ze += tp0 - t0;
            float *s_E = s_cpb + (size_t)set * 4 * cplen;   // copy 0; copy s at + s * cplen
            auto put = [&](int i, float v) {   // element i of the window into the four copies (zeros past the window: every copy is defined up to cplen)
#pragma unroll
                for (int sft = 0; sft < 4; sft++)
                    if (i >= sft && i - sft < cplen) s_E[sft * cplen + i - sft] = v;
            };
#pragma unroll
            for (int k = 0; k < EPF; k++) {
                const int i = tid + k * TC_PREP_THREADS;
                if (i < cplen + 3) put(i, ev[k]);
            }
            for (int i = tid + EPF * TC_PREP_THREADS; i < cplen + 3; i += TC_PREP_THREADS) {
                const int64_t idx = ebase + i;
                put(i, (i < a.span && idx >= e_lo && idx < e_hi) ? (float)a.E[idx] : 0.f);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PREP_THREADS) : "memory");
            if (DBG) t_eload += clock64() - tp0;
            unsigned char *zh = s_zhi, *zl = s_zlo + (size_t)set * nZ * 128;
            for (int e = tid; e < nZ * 8; e += TC_PREP_THREADS) {
                __align__(16) __half hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int idx = boff + e + j;
                    const float v = (idx < a.span) ? s_E[idx] * scale : 0.f;
                    hi[j] = __float2half_rn(v);
                    lo[j] = __float2half_rn(v - __half2float(hi[j]));
                }
                *reinterpret_cast<uint4 *>(zh + (size_t)e * 16) = *reinterpret_cast<uint4 *>(hi);
                *reinterpret_cast<uint4 *>(zl + (size_t)e * 16) = *reinterpret_cast<uint4 *>(lo);
            }
            asm volatile("fence.proxy.async;" ::: "memory");   // the lo rows are read by the leader's MMAs
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PREP_THREADS) : "memory");   // every hi row is in place before any thread expands its tile row from them
            if (DBG) t_prep += clock64() - tp0;
            if (lane == 0) {
                mbar_arrive(bar_zfull + 8 * set);
                mbar_arrive_cluster(bar_zpair + 8 * set, 0);
            }
        };
        // insert size 1: Bp[1,c] = E[c] (one tap) -> lin[x] = sum_k t1[k] E[x - w + k], a plain 1-D correlation in fp32: two
        // consecutive outputs per thread, 4 taps per step.  Only the epilogue's last addition needs it, so it runs when these
        // warps have nothing else due (after the operand of the item's first tile is in tensor memory).
        auto lin_item = [&](int n) {
            const int set = n & 1;
            const long long tl0 = DBG ? clock64() : 0;
            const int sh = -a.w - a.gmin;                    // >= 0: element x + k of the window shifted by sh is tap k of output x
            const float *s_E1 = s_cpb + (size_t)(set * 4 + (sh & 3)) * cplen + (sh & ~3);   // the copy that has it 16-byte aligned
            {
                const float2 *e2 = reinterpret_cast<const float2 *>(s_E1 + 2 * tid);
                const float4 *t4 = reinterpret_cast<const float4 *>(s_t1);
                float l0a = 0.f, l0b = 0.f, l1a = 0.f, l1b = 0.f;
                float2 ea = e2[0], eb = e2[1];
#pragma unroll 4
                for (int k4 = 0; k4 < Wp / 4; k4++) {
                    const float4 t = t4[k4];
                    const float2 ec = e2[2 * k4 + 2], ed = e2[2 * k4 + 3];
                    l0a = fmaf(t.x, ea.x, fmaf(t.y, ea.y, l0a));
                    l0b = fmaf(t.z, eb.x, fmaf(t.w, eb.y, l0b));
                    l1a = fmaf(t.x, ea.y, fmaf(t.y, eb.x, l1a));
                    l1b = fmaf(t.z, eb.y, fmaf(t.w, ec.x, l1b));
                    ea = ec;
                    eb = ed;
                }
                *reinterpret_cast<float2 *>(s_linb + (size_t)set * TC_TX + 2 * tid) = make_float2(l0a + l0b, l1a + l1b);
            }
            if (DBG) t_lin += clock64() - tl0;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_linfull + 8 * set);
        };
        // columns [c0, c1) of the hi operand of tile j: row m holds Es_hi[128 j + m + k], k = 2 c, 2 c + 1 -> the 16-byte Hankel rows
        // 128 j + m + 8 i are its columns 4 i .. 4 i + 3
        auto fill = [&](int set, int j, int part, int tcnt) {
            const int c0 = part ? a.c_split : 0, c1 = part ? a.c_end : a.c_split;
            const long long t0 = DBG ? clock64() : 0;
            mbar_wait(bar_afree + 8 * part, (tcnt & 1) ^ 1);   // the previous tile's MMAs that read these columns have retired
            tc_fence_after();
            const long long tf0 = DBG ? clock64() : 0;
            if (DBG) w_af += tf0 - t0;
            const unsigned char *zrow = s_zhi + (size_t)(TC_M * j + m) * 16;
            const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)TS_ACOL;
            for (int c = c0; c < c1; c += 32) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(zrow + (size_t)(c / 4 + i) * 128);
                    v[4 * i] = q.x;
                    v[4 * i + 1] = q.y;
                    v[4 * i + 2] = q.z;
                    v[4 * i + 3] = q.w;
                }
                tc_st32(ta + (uint32_t)c, v);
            }
            tc_wait_st();
            if (DBG) t_fill += clock64() - tf0;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(bar_afull + 8 * part, 0);   // the stores have completed (tcgen05.wait::st)
        };
        int it = next_valid(it_first), n = 0, tcnt = 0;
        if (it < n_items) prep_item(it, 0);
        while (it < n_items) {
            const int nx = next_valid(it + it_step);
            const int set = n & 1;
            for (int j = 0; j < TC_XT; j++, tcnt++) {
                fill(set, j, 0, tcnt);
                fill(set, j, 1, tcnt);
                if (j == 0 && a.has_row1) lin_item(n);   // nothing else is due before the last slab of tile 0
            }
            // under the MMAs of tile 1: the set it overwrites was released an item ago, and nothing is due from these warps
            // before the last slab of tile 1
            if (nx < n_items) prep_item(nx, n + 1);
            it = nx;
            n++;
        }
        if (DBG && tid == 0) {
            dbg[8] = (unsigned long long)w_ze;
            dbg[9] = (unsigned long long)w_af;
            dbg[10] = (unsigned long long)t_prep;
            dbg[11] = (unsigned long long)t_eload;
            dbg[12] = (unsigned long long)t_lin;
            dbg[13] = (unsigned long long)t_fill;
        }
    } else if (warp == TS_WARP_MMA && rank != 0) {
        // ===== peer: loads its half of every block image and tells the leader when it has landed
        if (lane == 0) {
            mbar_expect_tx(bar_gfull, (uint32_t)a.rank_bytes);
            const unsigned char *src = a.g_img + (size_t)rank * a.rank_bytes;
            for (int off = 0; off < a.rank_bytes; off += 32768) bulk_g2s(smem_u32(s_stage + off), src + off, (uint32_t)min(32768, a.rank_bytes - off), bar_gfull);
        }
        __syncwarp();
        mbar_wait(bar_gfull, 0);
        if (lane == 0) mbar_arrive_cluster(bar_gfull, 0);
        __syncwarp();
    } else if (warp == TS_WARP_MMA) {
        // ===== the issuing warp (leader): tiles one after the other, slabs through the two accumulators.  Everything in this
        // loop derives from kernel parameters and loop counters (warp-uniform): see TsTab.
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t sbase = (smem_u32(s_stage) >> 4) & 0x3FFF;
        long long w_zf = 0, w_te = 0, w_af = 0;
        int n = 0, tcnt = 0;
        uint32_t ab = 0, aph = 0;   // accumulator of the current slab (slabs go round the TS_NACC accumulators) and its use parity
        if (lane == 0) {   // this CTA's half of every block image, once
            mbar_expect_tx(bar_gfull, (uint32_t)a.rank_bytes);
            for (int off = 0; off < a.rank_bytes; off += 32768) bulk_g2s(smem_u32(s_stage + off), a.g_img + off, (uint32_t)min(32768, a.rank_bytes - off), bar_gfull);
        }
        __syncwarp();
        mbar_wait_cluster(bar_gfull, 0);
        tc_fence_after();
        for (int it = it_first; it < n_items; it += it_step) {
            if (!item_valid(it)) continue;
            const int set = n & 1;
            {
                const long long t0 = DBG ? clock64() : 0;
                mbar_wait_cluster(bar_zpair + 8 * set, (n >> 1) & 1);   // the lo rows of both CTAs
                tc_fence_after();
                if (DBG) w_zf += clock64() - t0;
            }
            const uint32_t zb = smem_u32(s_zlo + (size_t)set * nZ * 128);
            for (int j = 0; j < TC_XT; j++, tcnt++) {
                const uint32_t a_lo0 = (((zb + 2048u * j) >> 4) & 0x3FFF) | ((128u >> 4) << 16);
                for (int q = 0; q < a.n_slabs; q++) {
                    if (q == 0) {   // part 1 of the tile's hi operand is written
                        const long long t0 = DBG ? clock64() : 0;
                        mbar_wait_cluster(bar_afull, tcnt & 1);
                        tc_fence_after();
                        if (DBG) w_af += clock64() - t0;
                    }
                    {
                        const long long t0 = DBG ? clock64() : 0;
                        mbar_wait_cluster(bar_tempty + 8 * ab, aph ^ 1u);
                        tc_fence_after();
                        if (DBG) w_te += clock64() - t0;
                    }
                    const uint32_t d0 = tmem + ab * TS_N;
                    const int2 sl = tab.slab[q];
                    uint32_t acc0 = 0u;   // the first block of a slab is stored untrimmed: it initialises all TS_N columns
                    auto issue = [&](int t0, int t1) {
                        for (int t = t0; t < t1; t++) {
                            const int4 bk = tab.blk[sl.x + t];
                            const uint32_t kb = (uint32_t)bk.x & 0xffffu;
                            const uint32_t bh = (uint32_t)bk.w + sbase;
                            const uint32_t bl = bh + 2u * ((uint32_t)bk.w >> 16);
#ifndef TS_NO_MMA
                            tc_mma3_ts(d0 + (uint32_t)bk.y, tmem + (uint32_t)TS_ACOL + 8u * kb, a_lo0 + 16u * kb, bh, bl, desc_hi, (uint32_t)bk.z, acc0);
#endif
                            acc0 = 1u;
                            if ((bk.x >> 16) & 1) tc_commit_e2_both(bar_afree);   // nothing issued after this block reads part 1 of this tile's operand
                        }
                    };
                    // the slab that holds the tile's first reader of part 2 is issued in two runs with the wait in between (a wait
                    // INSIDE the block loop puts the loop back on per-thread registers)
                    const int tsplit = (q == a.q_need2) ? a.t_need2 : sl.y;
                    issue(0, tsplit);
                    if (q == a.q_need2) {
                        const long long t0 = DBG ? clock64() : 0;
                        mbar_wait_cluster(bar_afull + 8, tcnt & 1);
                        tc_fence_after();
                        if (DBG) w_af += clock64() - t0;
                        issue(tsplit, sl.y);
                    }
                    tc_commit_e2_both(bar_tfull + 8 * ab);
                    if (++ab == TS_NACC) {
                        ab = 0;
                        aph ^= 1u;
                    }
                }
                tc_commit_e2_both(bar_afree + 8);
            }
            tc_commit_e2_both(bar_zempty + 8 * set);
            n++;
        }
        if (DBG && lane == 0) {
            dbg[1] = (unsigned long long)w_zf;
            dbg[2] = (unsigned long long)w_te;
            dbg[3] = (unsigned long long)w_af;
            dbg[7] = (unsigned long long)n;
        }
    } else if (warp < TC_EPI_WARPS) {
        // ===== epilogue: bx[x] = sum_n E[x + A0 + n] * H[x, n]; group g (4 warps, one TMEM lane quarter each) reads every other
        // slab, keeps a partial sum per tile and adds the other group's through shared memory
        const int grp = warp >> 2, wq = warp & 3;
        const int m = wq * 32 + lane;
        const int aoff = a.A0 - a.gmin;
        const double unscale = ldexp(1.0, -(sE + a.sG));
#if TS_FIN32
        const bool unscale_f32 = sE + a.sG > -100 && sE + a.sG < 100;
        const float unscale_f = unscale_f32 ? (float)unscale : 0.f;
#endif
        const uint32_t tqaddr = tmem + ((uint32_t)(wq * 32) << 16);
        long long w_tf = 0, t_epi = 0, w_ld = 0;
        int n = 0, tcnt = 0;
        uint32_t u = 0, ab = 0, aph = 0;
        for (int it = it_first; it < n_items; it += it_step) {
            const int c = it / a.tiles_per_chunk, xb = (it - c * a.tiles_per_chunk) * item_w, x0 = xb + x_rank;
            const int64_t oo = a.out_off[c];
            const int L = (int)(a.out_off[c + 1] - oo);
            if (xb >= L) continue;
            const int set = n & 1;
            mbar_wait(bar_zfull + 8 * set, (n >> 1) & 1);
            // this thread's window starts at element aoff + m (+ multiples of 32): the copy shifted by (aoff + m) & 3 has it aligned
            const float *s_E = s_cpb + (size_t)(set * 4 + ((aoff + m) & 3)) * cplen + ((aoff + m) & ~3);
            for (int j = 0; j < TC_XT; j++, tcnt++) {
                float acc_hi = 0.f, acc_lo = 0.f;
                for (int q = 0; q < a.n_slabs; q++, u++) {
                    const uint32_t cab = ab, cph = aph;   // this slab's accumulator and use parity
                    if (++ab == TS_NACC) {
                        ab = 0;
                        aph ^= 1u;
                    }
                    if ((int)(u & 1u) != grp) continue;
                    const uint32_t t0addr = tqaddr + cab * TS_N;
                    const long long t0 = DBG ? clock64() : 0;
                    mbar_wait(bar_tfull + 8 * cab, cph);
                    tc_fence_after();
                    const long long t1 = DBG ? clock64() : 0;
                    if (DBG) w_tf += t1 - t0;
                    const float4 *Ew = reinterpret_cast<const float4 *>(s_E + TC_M * j + TS_N * q);
                    // TS_EPI_BUF register sets for the 32-column chunks of a slab: the tcgen05.ld of chunk h + TS_EPI_BUF goes out as soon as
                    // chunk h has been contracted; the accumulator returns to the issuing warp when the last chunk has landed
                    constexpr int NCH = TS_N / 32, NB = TS_EPI_BUF;
                    static_assert(TS_N % 32 == 0 && NCH >= NB, "slab = whole 32-column chunks");
                    uint32_t r[NB][32];
#pragma unroll
                    for (int h = 0; h < NB; h++) tc_ld32(t0addr + (uint32_t)(32 * h), r[h]);
#pragma unroll
                    for (int h = 0; h < NCH; h++) {
                        if (h <= NCH - NB) {   // something new has been requested since the last wait
                            const long long tw = DBG ? clock64() : 0;
                            tc_wait_ld();
                            if (DBG) w_ld += clock64() - tw;
                        }
                        if (h == NCH - NB) {   // the last chunk has landed too: the rest of the slab is in registers
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster_relaxed(bar_tempty + 8 * cab, 0);
                        }
                        float f[4] = {0.f, 0.f, 0.f, 0.f};   // four runs of 8 products in fp32
#pragma unroll
                        for (int g4 = 0; g4 < 8; g4++) {
                            const float4 e = Ew[8 * h + g4];
                            float v = f[g4 >> 1];
                            v = fmaf(__uint_as_float(r[h % NB][4 * g4]), e.x, v);
                            v = fmaf(__uint_as_float(r[h % NB][4 * g4 + 1]), e.y, v);
                            v = fmaf(__uint_as_float(r[h % NB][4 * g4 + 2]), e.z, v);
                            v = fmaf(__uint_as_float(r[h % NB][4 * g4 + 3]), e.w, v);
                            f[g4 >> 1] = v;
                        }
                        // the chunk's four runs in fp32 (two more roundings: 1e-7 of a chunk of positive terms), the chunks as a float pair
                        two_sum_acc(acc_hi, acc_lo, (f[0] + f[1]) + (f[2] + f[3]));
                        if (h + NB < NCH) tc_ld32(t0addr + (uint32_t)(32 * (h + NB)), r[h % NB]);
                    }
                    if (DBG) t_epi += clock64() - t1;
                }
                // The group that read the tile's LAST slab finishes the tile; the other one leaves its partial sums in shared memory
                // and goes on (no rendezvous: it may be up to TS_NACC slabs ahead, less than the 4 tiles the buffers cover).
#if TS_FIN32
                float2 *sp = reinterpret_cast<float2 *>(s_part) + (size_t)(tcnt & 3) * TC_M;
                const uint32_t bar_p = bar_pfull + 8 * (uint32_t)(tcnt & 3);
                if (grp != (int)((u - 1u) & 1u)) {
                    sp[m] = make_float2(acc_hi, acc_lo);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_p);   // release: the partial sums are visible to the waiting group
                } else {
                    if (a.has_row1 && j == 0) mbar_wait(bar_linfull + 8 * set, (n >> 1) & 1);
                    mbar_wait(bar_p, (uint32_t)(tcnt >> 2) & 1u);
                    const int x = x0 + TC_M * j + m;
                    const float lin = a.has_row1 ? s_linb[(size_t)set * TC_TX + TC_M * j + m] : 0.f;  // size-1 term (prep warps)
                    const float2 o = sp[m];
                    // group 0's pair first, whoever adds: one order of summation; sum and power-of-two unscaling in fp32 when the scale is an fp32 normal
                    const float h0 = grp ? o.x : acc_hi, l0 = grp ? o.y : acc_lo, h1 = grp ? acc_hi : o.x, l1 = grp ? acc_lo : o.y;
                    if (x < L) {
                        if (unscale_f32)
                            a.bx[oo + x] = (double)fmaf((h0 + h1) + (l0 + l1), unscale_f, lin);
                        else
                            a.bx[oo + x] = (((double)h0 + (double)h1) + ((double)l0 + (double)l1)) * unscale + (double)lin;
                    }
                }
            }
#else
                const double part = (double)acc_hi + (double)acc_lo;
                double *sp = s_part + (size_t)(tcnt & 3) * TC_M;
                const uint32_t bar_p = bar_pfull + 8 * (uint32_t)(tcnt & 3);
                if (grp != (int)((u - 1u) & 1u)) {
                    sp[m] = part;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_p);   // release: the partial sums are visible to the waiting group
                } else {
                    if (a.has_row1 && j == 0) mbar_wait(bar_linfull + 8 * set, (n >> 1) & 1);
                    mbar_wait(bar_p, (uint32_t)(tcnt >> 2) & 1u);
                    const int x = x0 + TC_M * j + m;
                    const double lin = a.has_row1 ? (double)s_linb[(size_t)set * TC_TX + TC_M * j + m] : 0.0;  // size-1 term (prep warps)
                    const double p0 = grp ? sp[m] : part, p1 = grp ? part : sp[m];   // group 0's part first, whoever adds: one order of summation
                    if (x < L) a.bx[oo + x] = (p0 + p1) * unscale + lin;
                }
            }
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_zempty + 8 * set);      // this warp no longer reads the E window of the set
            n++;
        }
        if (DBG && threadIdx.x == 0) {
            dbg[4] = (unsigned long long)w_tf;
            dbg[5] = (unsigned long long)t_epi;
            dbg[6] = (unsigned long long)w_ld;
            dbg[0] = (unsigned long long)(clock64() - t_begin);
        }
    }
    tc_fence_before();
    cluster_sync_all();   // no CTA leaves (or frees its TMEM) while its partner's MMAs / arrives may still touch it
    if (warp == TS_WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side: G images + stage table (once per VMat / fragment-size change)
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Host-only planning (no CUDA calls: also reachable through nb200_tc_plan_describe for the CPU tests)
// ---------------------------------------------------------------------------------------------
// Geometry of the bilinear form and the matrix G of a VMat x fragment-size distribution.  false: nothing for the tensor core
// (a VMat with only the single-tap size 1, or a zero / non-finite G).
struct TcGeom {
    int A0 = 0, B0 = 0, NA = 0, NB = 0, NAp = 0, NBp = 0, has_row1 = 0, gmin = 0, span = 0, sG = 0;
    double sc = 1.0;
    std::vector<double> G;   // [NAp][NBp], unscaled
};
static bool tc_build_geometry(const double *vmat, int lv, int uv, int W, int w, const double *sizes, TcGeom &g)
{
    auto fd2 = [](int v) { return v >> 1; };  // floor division by 2
    // tap offsets relative to the output position: a = (k-w) - (i-1)//2, b = (k-w) + i//2   (i != 1)
    int amin = INT32_MAX, amax = INT32_MIN, bmin = INT32_MAX, bmax = INT32_MIN;
    for (int i = lv; i < uv; i++) {
        if (i == 1) continue;
        amin = std::min(amin, -w - fd2(i - 1));
        amax = std::max(amax, w - fd2(i - 1));
        bmin = std::min(bmin, -w + fd2(i));
        bmax = std::max(bmax, w + fd2(i));
    }
    if (amin > amax) return false;
    g.A0 = amin;
    g.B0 = bmin;
    g.NA = amax - amin + 1;
    g.NB = bmax - bmin + 1;
    g.NAp = (g.NA + TC_N - 1) / TC_N * TC_N;
    g.NBp = (g.NB + 15) / 16 * 16;
    g.has_row1 = 0;  // size-1 row: a linear term, only paid for when f_1 * V[1,:] is not identically zero
    if (lv <= 1 && 1 < uv)
        for (int k = 0; k < W; k++)
            if (sizes[1] * vmat[(size_t)(1 - lv) * W + k] != 0.0) g.has_row1 = 1;
    // shared E window: genomic offsets [gmin, gmax] relative to the CTA's first output
    const int gmin = std::min(std::min(amin, bmin), -w);
    const int gmax = std::max(std::max(TC_TX - 1 + g.A0 + g.NAp - 1, TC_TX + g.B0 + g.NBp + 8), TC_TX - 1 + w);
    g.gmin = gmin;
    g.span = gmax - gmin + 1;
    g.G.assign((size_t)g.NAp * g.NBp, 0.0);
    double gmaxv = 0.0;
    for (int i = lv; i < uv; i++) {
        if (i == 1) continue;
        const double f = sizes[i];
        for (int k = 0; k < W; k++) {
            const int ia = (k - w) - fd2(i - 1) - g.A0, ib = (k - w) + fd2(i) - g.B0;
            const double v = f * vmat[(size_t)(i - lv) * W + k];
            g.G[(size_t)ia * g.NBp + ib] += v;
            gmaxv = std::max(gmaxv, fabs(v));
        }
    }
    if (!(gmaxv > 0.0) || !std::isfinite(gmaxv)) return false;
    int ex;
    frexp(32768.0 / gmaxv, &ex);
    g.sG = ex - 1;
    g.sc = ldexp(1.0, g.sG);
    return true;
}

struct TsHostPlan {
    bool ok = false;
    int n_slabs = 0, n_blocks = 0, rank_bytes = 0, c_split = 0, c_end = 0, q_need2 = -1, t_need2 = 0;
    long long mma_cols = 0, uncovered = 0, nonzero_cols = 0;   // MMA columns issued per x-tile and precision pass; non-zeros of G outside every block; non-zero (row, K block) pairs
    std::vector<int4> blk;
    std::vector<int2> slab;
    std::vector<unsigned char> img;   // [rank][blocks]: hi image then lo image of that CTA's half of every block
};

// Second plan (k_nuc_bx_ts): slabs of TS_N rows of G (NAp x NBp, row-major, scaled by `sc` in the images), all block images
// resident, one contiguous region per CTA of the pair.
static void ts_build_plan(const std::vector<double> &G, int NA, int NAp, int NBp, double sc, TsHostPlan &out)
{
    out.ok = false;
    const int nkb = NBp / 16;
    const int n_sl = (NA + TS_N - 1) / TS_N;
    const int c_end = (NBp / 2 + 31) / 32 * 32;
    struct Blk { int q, kb, n_lo, n_t; };
    std::vector<Blk> blks;
    std::vector<int2> &slabs = out.slab;
    slabs.clear();
    std::vector<std::vector<Blk>> per_slab(n_sl);
    for (int q = 0; q < n_sl; q++) {
        for (int kb = 0; kb < nkb; kb++) {
            int r_lo = TS_N, r_hi = -1;
            for (int n = 0; n < TS_N && q * TS_N + n < NAp; n++)
                for (int k = 0; k < 16; k++)
                    if (G[(size_t)(q * TS_N + n) * NBp + kb * 16 + k] != 0.0) {
                        r_lo = std::min(r_lo, n);
                        r_hi = std::max(r_hi, n);
                    }
            if (r_hi < 0) continue;
            per_slab[q].push_back({q, kb, r_lo / 16 * 16, (r_hi + 16) / 16 * 16 - r_lo / 16 * 16});
        }
        if (per_slab[q].empty()) per_slab[q].push_back({q, 0, 0, TS_N});
    }
    // The hi operand is written in two parts (32-column = 4-K-block granularity).  Part 1 = the K blocks the LAST slab does
    // not read: they are free a whole slab before the tile ends, so the next tile's part 1 is in place long before it
    // starts; slab 0 contracts its part-1 blocks first, and part 2 -- free only when the tile is complete -- is written
    // under them.  Without such a range (one slab, or a last slab that reads K block 0) part 1 = what slab 0 reads.
    int kmax0 = 0, kmin_last = nkb;
    for (const Blk &b : per_slab[0]) kmax0 = std::max(kmax0, b.kb);
    for (const Blk &b : per_slab[n_sl - 1]) kmin_last = std::min(kmin_last, b.kb);
    const bool early = n_sl >= 2 && kmin_last / 4 * 4 >= 4 && !(getenv("NB200_TC_EARLY") && atoi(getenv("NB200_TC_EARLY")) == 0);
    const int kb_split = early ? kmin_last / 4 * 4 : std::min((kmax0 + 1 + 3) / 4 * 4, c_end / 8);
    for (int q = 0; q < n_sl; q++) {
        std::vector<Blk> &v = per_slab[q];
        if (q == 0 && early)   // part-1 blocks first (K order is kept inside each group)
            std::stable_partition(v.begin(), v.end(), [&](const Blk &b) { return b.kb < kb_split; });
        // the first MMA of a slab initialises every accumulator column, so it is issued untrimmed: take the block that is
        // (closest to) full width anyway instead of the first in K order, whose rows are a corner of the hexagon
        // (251 x 251: 320 of 4912 MMA columns per x-tile saved; 5.44 -> 5.27 ms in an interleaved A/B)
        size_t cand_end = v.size();
        if (q == 0 && early) {
            cand_end = 0;
            while (cand_end < v.size() && v[cand_end].kb < kb_split) cand_end++;
            if (cand_end == 0) cand_end = v.size();
        }
        size_t widest = 0;
        for (size_t i = 1; i < cand_end; i++)
            if (v[i].n_t > v[widest].n_t) widest = i;
        std::rotate(v.begin(), v.begin() + widest, v.begin() + widest + 1);
        v[0].n_lo = 0;
        v[0].n_t = TS_N;
        slabs.push_back(make_int2((int)blks.size(), (int)v.size()));
        blks.insert(blks.end(), v.begin(), v.end());
    }
    int last_p1 = 0, first_p2 = -1;   // in issue order: the last block that reads part 1, the first that reads part 2
    for (size_t i = 0; i < blks.size(); i++) {
        if (blks[i].kb < kb_split) last_p1 = (int)i;
        else if (first_p2 < 0) first_p2 = (int)i;
    }
    if (first_p2 < 0) first_p2 = (int)blks.size() - 1;   // nobody reads part 2: its barrier phase is consumed before the last block
    size_t rank_bytes = 0;
    for (const Blk &b : blks) rank_bytes += (size_t)b.n_t * 32;   // half of the rows, hi + lo, 16 halves each
    std::vector<unsigned char> &timg = out.img;
    timg.assign(2 * rank_bytes, 0);
    std::vector<int4> &tblk = out.blk;
    tblk.clear();
    size_t off = 0;
    for (size_t i = 0; i < blks.size(); i++) {
        const Blk &b = blks[i];
        const int nh = b.n_t / 2;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(b.n_t >> 3) << 17) | ((uint32_t)((TC_M * 2) >> 4) << 24);
        const int flags = ((int)i == last_p1 ? 1 : 0) | ((int)i == first_p2 ? 4 : 0);
        tblk.push_back(make_int4(b.kb | (flags << 16), b.n_lo, (int)idesc, (int)(off >> 4) | (nh << 16)));
        for (int rk = 0; rk < 2; rk++) {
            __half *hi = reinterpret_cast<__half *>(timg.data() + (size_t)rk * rank_bytes + off);
            __half *lo = hi + (size_t)nh * 16;
            for (int n = 0; n < nh; n++)
                for (int k = 0; k < 16; k++) {
                    const int row = b.q * TS_N + b.n_lo + rk * nh + n;
                    const double g = (row < NAp ? G[(size_t)row * NBp + b.kb * 16 + k] : 0.0) * sc;
                    const float gf = (float)g;
                    const __half h = __float2half_rn(gf);
                    const __half l = __float2half_rn(gf - __half2float(h));
                    const size_t o = ((size_t)(k / 8) * (nh / 8) + n / 8) * 64 + (n % 8) * 8 + (k % 8);
                    hi[o] = h;
                    lo[o] = l;
                }
        }
        off += (size_t)nh * 64;
    }
    out.n_slabs = n_sl;
    out.n_blocks = (int)tblk.size();
    out.rank_bytes = (int)rank_bytes;
    out.c_end = c_end;
    out.c_split = kb_split * 8;
    out.q_need2 = blks[first_p2].q;
    out.t_need2 = first_p2 - slabs[blks[first_p2].q].x;
    out.mma_cols = 0;
    for (const Blk &b : blks) out.mma_cols += b.n_t;
    // self-check: every non-zero of G lies inside a block of its slab
    std::vector<char> covered((size_t)NAp * nkb, 0);
    for (const Blk &b : blks)
        for (int n = 0; n < b.n_t; n++)
            if (b.q * TS_N + b.n_lo + n < NAp) covered[(size_t)(b.q * TS_N + b.n_lo + n) * nkb + b.kb] = 1;
    out.uncovered = out.nonzero_cols = 0;
    for (int row = 0; row < NAp; row++)
        for (int kb = 0; kb < nkb; kb++) {
            bool nz = false;
            for (int k = 0; k < 16; k++) nz = nz || G[(size_t)row * NBp + kb * 16 + k] != 0.0;
            if (nz) {
                out.nonzero_cols++;
                if (!covered[(size_t)row * nkb + kb]) out.uncovered++;
            }
        }
    out.ok = c_end <= TS_MAXCOL && rank_bytes < (size_t)200 * 1024 && tblk.size() <= TS_MAX_BLOCKS && slabs.size() <= TS_MAX_SLABS && out.uncovered == 0;
}

int nb200_tc_setup(nb200_ctx *ctx)
{
    RunConst &r = ctx->rc;
    TcPlan *pl = plan_of(ctx, true);
    if (!pl) return NB200_OK;
    pl->ok = false;
    if (!r.have_vmat || !r.have_sizes) return NB200_OK;
    const int lv = r.v_lower, uv = r.v_upper, W = r.v_cols, w = r.v_w;
    TcGeom geo;
    if (!tc_build_geometry(r.h_vmat.data(), lv, uv, W, w, r.h_sizes.data(), geo)) return NB200_OK;   // nothing for the tensor core
    pl->A0 = geo.A0;
    pl->B0 = geo.B0;
    pl->NA = geo.NA;
    pl->NB = geo.NB;
    pl->NAp = geo.NAp;
    pl->NBp = geo.NBp;
    pl->n_achunks = pl->NAp / TC_N;
    pl->has_row1 = geo.has_row1;
    pl->gmin = geo.gmin;
    pl->span = geo.span;
    pl->sG = geo.sG;
    const std::vector<double> &G = geo.G;
    const double sc = geo.sc;
    // Per slab (TC_N rows of G = a taps) the K16 blocks that hold a non-zero, each trimmed to the 16-row granules that are
    // non-zero in it; the first block of a slab stays untrimmed (its MMA initialises every accumulator column).  Blocks
    // are packed into stages of at most TC_SLOT_BYTES.  Block image (hi, then lo): canonical K-major no-swizzle core
    // matrices, (n/8, k/8) at ((k/8) * (N_t/8) + n/8) * 128 B, row n%8, element k%8.
    pl->h_tab.clear();
    pl->h_blk.clear();
    pl->pair = (getenv("NB200_TC_PAIR") && atoi(getenv("NB200_TC_PAIR")) == 0) ? 0 : 1;   // developer switch: single-CTA MMAs
    const int NR = pl->pair ? 2 : 1;   // CTAs sharing a block of G
    std::vector<unsigned char> img;
    const int nkb = pl->NBp / 16;
    long long cols_issued = 0;
    for (int q = 0; q < pl->n_achunks; q++) {
        struct Blk { int kb, n_lo, n_t; };
        std::vector<Blk> blks;
        for (int kb = 0; kb < nkb; kb++) {
            int r_lo = TC_N, r_hi = -1;
            for (int n = 0; n < TC_N; n++)
                for (int k = 0; k < 16; k++)
                    if (G[(size_t)(q * TC_N + n) * pl->NBp + kb * 16 + k] != 0.0) {
                        r_lo = std::min(r_lo, n);
                        r_hi = std::max(r_hi, n);
                    }
            if (r_hi < 0) continue;
            const int n_lo = r_lo / 16 * 16, n_hi = (r_hi + 16) / 16 * 16;
            blks.push_back({kb, n_lo, n_hi - n_lo});
        }
        if (blks.empty()) blks.push_back({0, 0, TC_N});  // keep every slab present (one zero block)
        {   // the untrimmed first block of the slab: the widest one instead of the first in K order (see the second plan below)
            size_t widest = 0;
            for (size_t i = 1; i < blks.size(); i++)
                if (blks[i].n_t > blks[widest].n_t) widest = i;
            std::rotate(blks.begin(), blks.begin() + widest, blks.begin() + widest + 1);
        }
        blks[0].n_lo = 0;
        blks[0].n_t = TC_N;
        size_t bi = 0;
        while (bi < blks.size()) {
            const int first_blk = (int)pl->h_blk.size();
            const size_t base = img.size();
            int bytes = 0, cnt = 0;
            const size_t bi_start = bi;
            // which blocks go into this stage (their images together fit a ring slot)
            size_t bj = bi;
            while (bj < blks.size() && cnt < TC_MAX_STAGE_BLOCKS && bytes + blks[bj].n_t * 64 <= TC_SLOT_BYTES) {
                bytes += blks[bj].n_t * 64;
                cnt++;
                bj++;
            }
            img.resize(base + bytes, 0);
            // stage image: for every CTA of the pair (one without pairs) its share of the rows of each block -- rank r holds
            // rows [r n_t / NR, (r + 1) n_t / NR) as a canonical hi image followed by the lo image
            const int rank_bytes = bytes / NR;
            for (int rk = 0; rk < NR; rk++) {
                int off_r = 0;
                for (size_t bb = bi; bb < bj; bb++) {
                    const Blk &bk = blks[bb];
                    const int nh = bk.n_t / NR;   // n_t is a multiple of 16: the halves are multiples of 8 rows
                    if (rk == 0) {
                        const uint32_t idesc = (1u << 4) | ((uint32_t)(bk.n_t >> 3) << 17) | ((uint32_t)((TC_M * NR) >> 4) << 24);  // f16 x f16 -> f32, M = 128 per CTA
                        pl->h_blk.push_back(make_int4(16 * bk.kb, bk.n_lo, (int)idesc, (off_r >> 4) | (nh << 16)));
                        cols_issued += bk.n_t;
                    }
                    __half *hi = reinterpret_cast<__half *>(img.data() + base + (size_t)rk * rank_bytes + off_r);
                    __half *lo = hi + (size_t)nh * 16;
                    for (int n = 0; n < nh; n++)
                        for (int k = 0; k < 16; k++) {
                            const double g = G[(size_t)(q * TC_N + bk.n_lo + rk * nh + n) * pl->NBp + bk.kb * 16 + k] * sc;
                            const float gf = (float)g;
                            const __half h = __float2half_rn(gf);
                            const __half l = __float2half_rn(gf - __half2float(h));
                            const size_t off = ((size_t)(k / 8) * (nh / 8) + n / 8) * 64 + (n % 8) * 8 + (k % 8);
                            hi[off] = h;
                            lo[off] = l;
                        }
                    off_r += nh * 64;
                }
            }
            bi = bj;
            const int flags = (bi_start == 0 ? 1 : 0) | (bi == blks.size() ? 2 : 0);
            pl->h_tab.push_back(make_int4(first_blk, cnt | (flags << 8), bytes, (int)(base >> 4)));
        }
    }
    pl->n_blocks = (int)pl->h_blk.size();
    pl->n_stages = (int)pl->h_tab.size();
    pl->density = (double)cols_issued / ((double)pl->NAp * nkb);
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    pl->img_bytes = img.size();
    NB_CUDA(ctx, pl->g_img.reserve(img.size()));
    NB_CUDA(ctx, cudaMemcpy(pl->g_img.p, img.data(), img.size(), cudaMemcpyHostToDevice));
    NB_CUDA(ctx, pl->stage_tab.reserve(sizeof(int4) * pl->h_tab.size()));
    NB_CUDA(ctx, cudaMemcpy(pl->stage_tab.p, pl->h_tab.data(), sizeof(int4) * pl->h_tab.size(), cudaMemcpyHostToDevice));
    NB_CUDA(ctx, pl->block_tab.reserve(sizeof(int4) * pl->h_blk.size()));
    NB_CUDA(ctx, cudaMemcpy(pl->block_tab.p, pl->h_blk.data(), sizeof(int4) * pl->h_blk.size(), cudaMemcpyHostToDevice));
    // ---- second plan (k_nuc_bx_ts): slabs of TS_N rows, all block images resident, one contiguous region per CTA of the pair
    pl->ts_ok = false;
    {
        TsHostPlan hp;
        ts_build_plan(G, pl->NA, pl->NAp, pl->NBp, sc, hp);
        if (hp.ok) {
            pl->ts_slabs = hp.n_slabs;
            pl->ts_blocks = hp.n_blocks;
            pl->ts_rank_bytes = hp.rank_bytes;
            pl->ts_c_end = hp.c_end;
            pl->ts_c_split = hp.c_split;
            pl->ts_q_need2 = hp.q_need2;
            pl->ts_t_need2 = hp.t_need2;
            NB_CUDA(ctx, pl->ts_img.reserve(hp.img.size()));
            NB_CUDA(ctx, cudaMemcpy(pl->ts_img.p, hp.img.data(), hp.img.size(), cudaMemcpyHostToDevice));
            pl->ts_blk = hp.blk;
            pl->ts_slab = hp.slab;
            pl->ts_ok = true;
        }
    }
    std::vector<double> t1(W, 0.0);
    if (pl->has_row1)
        for (int k = 0; k < W; k++) t1[k] = r.h_sizes[1] * r.h_vmat[(size_t)(1 - lv) * W + k];
    NB_CUDA(ctx, pl->t_row1.reserve(sizeof(double) * W));
    NB_CUDA(ctx, cudaMemcpy(pl->t_row1.p, t1.data(), sizeof(double) * W, cudaMemcpyHostToDevice));
    NB_CUDA(ctx, pl->emax.reserve(sizeof(double)));
    pl->ok = true;
    return NB200_OK;
}

int nb200_tc_available(nb200_ctx *ctx)
{
    TcPlan *pl = plan_of(ctx, false);
    return pl && pl->ok;
}

// Developer / test aid, no device needed: the block plan k_nuc_bx_ts would run for this VMat (rows = insert sizes
// [lower, upper), `cols` columns) and fragment-size distribution (at least `upper` values).  stats[16] = {eligible, slabs,
// blocks, hi-operand columns of part 1, columns in all, slab and position of the first block that reads part 2, bytes of one
// CTA's image, NA, NB, MMA columns per x-tile and pass, non-zero (row, K block) pairs of G, non-zeros outside every block,
// size-1 term present, 0, 0}; blocks[4 * i ..] = the table entry of block i (at most max_blocks are copied).
extern "C" int nb200_tc_plan_describe(const double *vmat, int32_t lower, int32_t upper, int32_t cols, const double *sizes, int32_t n_sizes,
                                      int32_t *stats, int32_t *blocks, int32_t max_blocks)
{
    if (!vmat || !sizes || !stats || upper <= lower || cols < 1 || n_sizes < upper || lower < 0) return NB200_ERR_ARG;
    for (int i = 0; i < 16; i++) stats[i] = 0;
    TcGeom geo;
    if (!tc_build_geometry(vmat, lower, upper, cols, cols / 2, sizes, geo)) return NB200_OK;
    TsHostPlan hp;
    ts_build_plan(geo.G, geo.NA, geo.NAp, geo.NBp, geo.sc, hp);
    stats[0] = hp.ok ? 1 : 0;
    stats[1] = hp.n_slabs;
    stats[2] = hp.n_blocks;
    stats[3] = hp.c_split;
    stats[4] = hp.c_end;
    stats[5] = hp.q_need2;
    stats[6] = hp.t_need2;
    stats[7] = hp.rank_bytes;
    stats[8] = geo.NA;
    stats[9] = geo.NB;
    stats[10] = (int32_t)hp.mma_cols;
    stats[11] = (int32_t)hp.nonzero_cols;
    stats[12] = (int32_t)hp.uncovered;
    stats[13] = geo.has_row1;
    if (blocks)
        for (int i = 0; i < hp.n_blocks && i < max_blocks; i++) {
            blocks[4 * i] = hp.blk[i].x;
            blocks[4 * i + 1] = hp.blk[i].y;
            blocks[4 * i + 2] = hp.blk[i].z;
            blocks[4 * i + 3] = hp.blk[i].w;
        }
    return NB200_OK;
}

// max over the batch's E track (E > 0: IEEE bit patterns order like unsigned integers)
__global__ void k_emax(const double *__restrict__ E, int64_t n, unsigned long long *out)
{
    unsigned long long m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = E[i];
        if (v == v && v < CUDART_INF) m = max(m, (unsigned long long)__double_as_longlong(v));
    }
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(NB_FULL, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

int nb200_nuc_bx_tc(nb200_ctx *ctx, nb200_dbatch *b)
{
    TcPlan *pl = plan_of(ctx, false);
    if (!pl || !pl->ok) return nb200_fail(ctx, NB200_ERR_STATE, "tcgen05 xcor plan is not available");
    RunConst &r = ctx->rc;
    NB_CUDA(ctx, cudaMemsetAsync(pl->emax.p, 0, sizeof(double), b->stream));
    {
        ProfScope ps(ctx, b->stream, "k_emax");
        k_emax<<<ctx->sm_count * 4, 256, 0, b->stream>>>(b->d_E.as<double>(), b->n_bias, pl->emax.as<unsigned long long>());
        NB_LAUNCH_CHECK(ctx);
    }
    static const bool tc_debug = getenv("NB200_TC_DEBUG") != nullptr;
    static int *h_trap = nullptr;
    if (tc_debug && !h_trap) {
        NB_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&h_trap), 2100 * sizeof(int), cudaHostAllocMapped));
        memset(h_trap, 0, 2100 * sizeof(int));
        int *d_trap = nullptr;
        NB_CUDA(ctx, cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_trap), h_trap, 0));
        NB_CUDA(ctx, cudaMemcpyToSymbol(g_tc_trap_info, &d_trap, sizeof(d_trap)));
    }
    // NB200_TC_TS = 1 / 0 forces / forbids the tensor-memory kernel (read per call: the tests compare the two kernels in one
    // process).  By default it is used from ~24 K16 blocks per x-tile on: below that (VMats under ~151 x 151) an x-tile is
    // so short that expanding its operand into tensor memory costs more than it saves (101 x 101: 9.6 % against 12.9 % of peak).
    const char *ts_e = getenv("NB200_TC_TS");
    const bool ts_env = ts_e ? atoi(ts_e) != 0 : pl->ts_blocks >= 24;
    if (pl->ts_ok && ts_env && ctx->sm_count >= 2) {
        // ---- hi operand in tensor memory (k_nuc_bx_ts) whenever the whole image fits next to the operands
        TsArgs t;
        t.start = b->d_start.as<int32_t>();
        t.out_off = b->d_out_off.as<int64_t>();
        t.bias_off = b->d_bias_off.as<int64_t>();
        t.seq_start = b->d_seq_start.as<int32_t>();
        t.E = b->d_E.as<double>();
        t.emax = pl->emax.as<double>();
        t.g_img = pl->ts_img.as<unsigned char>();
        t.t_row1 = pl->t_row1.as<double>();
        t.bx = b->n_bx.as<double>();
        t.dbg = nullptr;
        t.pwm_up = r.pwm_up;
        t.A0 = pl->A0;
        t.B0 = pl->B0;
        t.NBp = pl->NBp;
        t.gmin = pl->gmin;
        t.n_slabs = pl->ts_slabs;
        t.n_blocks = pl->ts_blocks;
        t.sG = pl->sG;
        t.has_row1 = pl->has_row1;
        t.W = r.v_cols;
        t.w = r.v_w;
        t.epad = (32 - ((pl->A0 - pl->gmin) & 31)) & 31;
        t.n_chunks = b->n_chunks;
        t.tiles_per_chunk = (int)div_up64(b->max_len, 2 * TC_TX);
        t.rank_bytes = pl->ts_rank_bytes;
        t.c_split = pl->ts_c_split;
        t.c_end = pl->ts_c_end;
        t.q_need2 = pl->ts_q_need2;
        t.t_need2 = pl->ts_t_need2;
        // the E window must also cover the a taps of the last slab (TS_N granularity) and the rows the hi operand is expanded from
        const int gmax = std::max(std::max(TC_TX - 1 + pl->A0 + pl->ts_slabs * TS_N - 1, TC_TX + pl->B0 + std::max(pl->NBp, 2 * pl->ts_c_end) + 8), TC_TX - 1 + r.v_w);
        t.span = std::max(pl->span, gmax - pl->gmin + 1);
        t.nZ = (TC_TX + std::max(pl->NBp, 2 * pl->ts_c_end) + 7) / 8;
        const int Wp4 = (r.v_cols + 3) & ~3;
        const int cplen = ((t.span + 3 + 31) & ~31) + 8;
        const size_t smem_ts = ((size_t)t.rank_bytes + 3 * (size_t)t.nZ * 128 + sizeof(float) * 2 * 4 * cplen +
                                sizeof(float) * (pl->has_row1 ? (Wp4 + 2 * TC_TX) : 0) + sizeof(double) * 4 * TC_M + 8 * (TS_BARS + 1) + 16 + 127) / 128 * 128;
        if (smem_ts <= 227 * 1024) {
            void (*kern)(TsTab, TsArgs) = tc_debug ? k_nuc_bx_ts<true> : k_nuc_bx_ts<false>;
            TsTab tab;   // 3 KB, passed by value: the launch copies the parameter
            memset(&tab, 0, sizeof(tab));
            memcpy(tab.blk, pl->ts_blk.data(), sizeof(int4) * pl->ts_blk.size());
            memcpy(tab.slab, pl->ts_slab.data(), sizeof(int2) * pl->ts_slab.size());
            NB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ts));
            ProfScope ps(ctx, b->stream, "k_nuc_bx_ts");
            const int n_items = t.n_chunks * t.tiles_per_chunk;
            dim3 grid((unsigned)(2 * std::max(1, std::min(ctx->sm_count / 2, n_items))));
            unsigned long long *d_dbg = nullptr;
            if (tc_debug) {
                NB_CUDA(ctx, cudaMalloc(&d_dbg, (size_t)grid.x * 16 * sizeof(unsigned long long)));
                NB_CUDA(ctx, cudaMemsetAsync(d_dbg, 0, (size_t)grid.x * 16 * sizeof(unsigned long long), b->stream));
                t.dbg = d_dbg;
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid;
            cfg.blockDim = dim3(TS_THREADS);
            cfg.dynamicSmemBytes = smem_ts;
            cfg.stream = b->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            NB_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, tab, t));
            NB_LAUNCH_CHECK(ctx);
            if (tc_debug) {
                std::vector<unsigned long long> h((size_t)grid.x * 16);
                cudaError_t e = cudaStreamSynchronize(b->stream);
                if (e != cudaSuccess) {
                    fprintf(stderr, "[ts debug] kernel failed (%s); %d starving warps | first barrier of CTA0 0x%x CTA1 0x%x tmem 0x%x 0x%x; slabs %d blocks %d "
                                    "c_split %d c_end %d q_need2 %d grid %u\n", cudaGetErrorString(e), h_trap[0], h_trap[8], h_trap[10], h_trap[9], h_trap[11],
                            t.n_slabs, t.n_blocks, t.c_split, t.c_end, t.q_need2, grid.x);
                    for (int i = 0; i < std::min(h_trap[0], 60); i++) {
                        const int *rr = h_trap + 16 + 4 * i;
                        fprintf(stderr, "[ts debug]   CTA %d warp %d waits on barrier %d parity %d\n", rr[0], rr[1] >> 5,
                                ((rr[2] & 0xffffff) - (h_trap[8 + 2 * (rr[0] & 1)] & 0xffffff)) / 8, rr[3] & 1);
                    }
                    return nb200_cuda_fail(ctx, e, "k_nuc_bx_ts (debug sync)", __FILE__, __LINE__);
                }
                NB_CUDA(ctx, cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                cudaFree(d_dbg);
                double mm[16] = {0};
                for (size_t i = 0; i < grid.x; i++)
                    for (int k = 0; k < 16; k++) mm[k] += (double)h[16 * i + k];
                static const char *nm[16] = {"cta_total", "mma_wait_lo_rows", "mma_wait_tmem_empty", "mma_wait_hi_operand", "epi_wait_tmem_full",
                                            "epi_compute", "epi_wait_ld", "items",
                                            "prep_wait_set_empty", "fill_wait_free", "prep_item", "prep_E_load", "prep_row1", "fill_body", "-", "-"};
                fprintf(stderr, "[ts debug] %u CTAs, image %d bytes per CTA, %d slabs, %d blocks, hi operand %d + %d columns, smem %zu:", grid.x, t.rank_bytes,
                        t.n_slabs, t.n_blocks, t.c_split, t.c_end - t.c_split, smem_ts);
                for (int k = 0; k < 14; k++) fprintf(stderr, " %s=%.0f", nm[k], mm[k] / (double)grid.x);
                fprintf(stderr, "\n");
            }
            return NB200_OK;
        }
    }
    TcArgs a;
    a.start = b->d_start.as<int32_t>();
    a.out_off = b->d_out_off.as<int64_t>();
    a.bias_off = b->d_bias_off.as<int64_t>();
    a.seq_start = b->d_seq_start.as<int32_t>();
    a.E = b->d_E.as<double>();
    a.emax = pl->emax.as<double>();
    a.tab = pl->stage_tab.as<int4>();
    a.blk = pl->block_tab.as<int4>();
    a.g_img = pl->g_img.as<unsigned char>();
    a.t_row1 = pl->t_row1.as<double>();
    a.bx = b->n_bx.as<double>();
    a.dbg = nullptr;
    a.pwm_up = r.pwm_up;
    a.A0 = pl->A0;
    a.B0 = pl->B0;
    a.NAp = pl->NAp;
    a.NBp = pl->NBp;
    a.gmin = pl->gmin;
    a.span = pl->span;
    a.n_stages = pl->n_stages;
    a.n_achunks = pl->n_achunks;
    a.n_blocks = pl->n_blocks;
    a.epad = (32 - ((pl->A0 - pl->gmin) & 31)) & 31;
    a.sG = pl->sG;
    a.has_row1 = pl->has_row1;
    a.W = r.v_cols;
    a.w = r.v_w;
    a.n_chunks = b->n_chunks;
    const int pair = pl->pair;
    a.tiles_per_chunk = (int)div_up64(b->max_len, TC_TX * (pair ? 2 : 1));
    a.slot_bytes = TC_SLOT_BYTES / (pair ? 2 : 1);
    const int nZ = (TC_TX + pl->NBp) / 8;
    const size_t fixed = 4 * (size_t)nZ * 128 + sizeof(float) * 2 * ((a.epad + pl->span + 31) & ~31) +
                         sizeof(float) * (pl->has_row1 ? (((r.v_cols + 3) & ~3) + 2 * (TC_TX + ((r.v_cols + 3) & ~3) + 8) + 2 * TC_TX) : 0) + sizeof(int4) * (pl->n_stages + pl->n_blocks + 1) +
                         8 * (2 * TC_MAX_STAGES + 2 * TC_XT + 4 + 1 + 2) + 16 + 128 + sizeof(int) * ((pl->n_stages + 3) & ~3);
    // resident mode: this CTA's half of the whole image stays in shared memory for the life of the kernel (when it fits)
    static const bool res_env = !(getenv("NB200_TC_RES") && atoi(getenv("NB200_TC_RES")) == 0);
    const bool res = pair && res_env && fixed + pl->img_bytes / 2 + 256 <= 227 * 1024;
    a.res_bytes = res ? (int)(pl->img_bytes / 2) : 0;
    int ring = TC_MAX_STAGES;
    while (ring > 2 && fixed + (size_t)ring * a.slot_bytes > 227 * 1024) ring--;
    const size_t smem = ((res ? fixed + (size_t)a.res_bytes : fixed + (size_t)ring * a.slot_bytes) + 127) / 128 * 128;
    if (smem > 227 * 1024) return nb200_fail(ctx, NB200_ERR_ARG, "VMat too large for the tcgen05 background kernel");
    a.ring = ring;
    static const int stag_env = getenv("NB200_TC_STAGGER") ? atoi(getenv("NB200_TC_STAGGER")) : 2;
    a.stagger = stag_env < 0 ? -1 : std::min(std::min(stag_env, ring - 2), pl->n_stages - 1);
    void (*kern)(TcArgs) = res    ? (tc_debug ? k_nuc_bx_tc<true, true, true> : k_nuc_bx_tc<false, true, true>)
                           : pair ? (tc_debug ? k_nuc_bx_tc<true, true, false> : k_nuc_bx_tc<false, true, false>)
                                  : (tc_debug ? k_nuc_bx_tc<true, false, false> : k_nuc_bx_tc<false, false, false>);
    NB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(ctx, b->stream, "k_nuc_bx_tc");
    const int n_items = a.n_chunks * a.tiles_per_chunk;
    // one persistent CTA per SM; with pairs one cluster of 2 CTAs per TPC, both walking the same items
    dim3 grid(pair ? (unsigned)(2 * std::max(1, std::min(ctx->sm_count / 2, n_items))) : (unsigned)std::max(1, std::min(ctx->sm_count, n_items)));
    unsigned long long *d_dbg = nullptr;
    const size_t n_cta = grid.x;
    if (tc_debug) {
        NB_CUDA(ctx, cudaMalloc(&d_dbg, n_cta * 8 * sizeof(unsigned long long)));
        NB_CUDA(ctx, cudaMemsetAsync(d_dbg, 0, n_cta * 8 * sizeof(unsigned long long), b->stream));
        a.dbg = d_dbg;
    }
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = b->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = pair ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        NB_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, a));
    }
    NB_LAUNCH_CHECK(ctx);
    if (tc_debug) {  // developer aid: mean clock counts per CTA of the waits of each warp role
        std::vector<unsigned long long> h(n_cta * 8);
        {
            cudaError_t e = cudaStreamSynchronize(b->stream);
            if (e != cudaSuccess) {   // which barriers starved (records of tc_trap)
                fprintf(stderr, "[tc debug] kernel failed (%s); %d starving warps | bar_full of CTA0 0x%x CTA1 0x%x tmem 0x%x 0x%x; pair %d resident %d "
                                "ring %d stages %d slabs %d grid %u\n", cudaGetErrorString(e), h_trap[0], h_trap[8], h_trap[10], h_trap[9], h_trap[11], pair,
                        (int)res, ring, pl->n_stages, pl->n_achunks, grid.x);
                int hist[2][16][32][2] = {};   // rank, warp, barrier slot, parity -> count
                for (int i = 0; i < std::min(h_trap[0], 500); i++) {
                    const int *r = h_trap + 16 + 4 * i;
                    const int sl = ((r[2] & 0xffffff) - (h_trap[8] & 0xffffff)) / 8;
                    if (sl >= 0 && sl < 32 && (r[1] >> 5) < 16) hist[r[0] & 1][r[1] >> 5][sl][r[3] & 1]++;
                }
                for (int rk = 0; rk < 2; rk++)
                    for (int w = 0; w < 16; w++)
                        for (int sl = 0; sl < 32; sl++)
                            for (int pa = 0; pa < 2; pa++)
                                if (hist[rk][w][sl][pa])
                                    fprintf(stderr, "[tc debug]   rank %d warp %2d waits on barrier slot %2d parity %d: %d CTAs\n", rk, w, sl, pa, hist[rk][w][sl][pa]);
                return nb200_cuda_fail(ctx, e, "k_nuc_bx_tc (debug sync)", __FILE__, __LINE__);
            }
        }
        NB_CUDA(ctx, cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        cudaFree(d_dbg);
        double m[8] = {0};
        for (size_t i = 0; i < n_cta; i++)
            for (int k = 0; k < 8; k++) m[k] += (double)h[8 * i + k];
        static const char *nm[8] = {"cta_total", "mma_wait_operands", "mma_wait_tmem_empty", "mma_wait_stage_full", "epi_wait_tmem_full",
                                    "epi_compute", "prep_wait_set_empty", "items"};
        fprintf(stderr, "[tc debug] %zu CTAs, pair %d resident %d (%d bytes), ring %d (%d stages, %d slabs per item), smem %zu:", n_cta, pair, (int)res,
                a.res_bytes, ring, pl->n_stages, pl->n_achunks, smem);
        for (int k = 0; k < 8; k++) fprintf(stderr, " %s=%.0f", nm[k], m[k] / (double)n_cta);
        fprintf(stderr, "\n");
    }
    return NB200_OK;
}
