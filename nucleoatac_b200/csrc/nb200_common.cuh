// nb200_common.cuh -- shared declarations of libnucleo_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/nucleo_b200.h"

#define NB200_MAX_PWM_WIDTH 64
#define NB200_MAX_NUC 8

// ---------------------------------------------------------------------------------------------
// device buffer that only ever grows (no per-batch cudaMalloc on the steady-state path)
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct ProfEntry {
    std::string name;
    int64_t launches = 0;
    double ms = 0.0;
};

// run constants living on the device
struct RunConst {
    // PWM
    int pwm_up = 0, pwm_down = 0, pwm_width = 0, n_nuc = 0;
    bool have_pwm = false;
    DevBuf log_pwm;   // f64 [n_nuc][width]
    DevBuf nuc_code;  // int8 [256]: byte -> PWM row or -1
    // VMat
    bool have_vmat = false;
    int v_rows = 0, v_cols = 0, v_lower = 0, v_upper = 0, v_w = 0;
    bool v_has_zero = false;
    DevBuf vmat;      // f64 [R][W]
    DevBuf vmat_f;    // f64 [R][W]  f_i * V      (needs fragment sizes)
    DevBuf vmat_f2;   // f64 [R][W]  f_i * V^2
    std::vector<double> h_vmat;
    // fragment sizes
    bool have_sizes = false;
    int sizes_upper = 0;
    DevBuf sizes;     // f64 [upper]
    std::vector<double> h_sizes;
    // occupancy model
    bool have_occ_model = false;
    int occ_upper = 0, n_alpha = 0;
    double cutoff = 0.0;
    int pn_has_zero = 0, pf_has_zero = 0;
    DevBuf nuc_probs, nfr_probs, alphas;
    // jitter
    int64_t n_jitter = 0;
    DevBuf jitter;
    // smoothing windows
    DevBuf occ_win, nuc_win;
};

struct nb200_ctx {
    int device = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t hbm_bytes = 0;
    cudaStream_t stream = nullptr;  // stream of the primitive calls
    std::string err;
    RunConst rc;
    nb200_occ_params occ{};
    nb200_nuc_params nuc{};
    bool occ_configured = false, nuc_configured = false;
    // scratch for the primitive (single-call) paths
    DevBuf s0, s1, s2, s3, s4;
    // profiling
    bool prof_on = false;
    std::vector<ProfEntry> prof;
    std::vector<cudaEvent_t> ev_pool;
    struct PendingEv {
        int idx;
        cudaEvent_t a, b;
    };
    std::vector<PendingEv> pending;
    DevBuf flush;  // L2 flush target
    // nccl
    void *nccl_lib = nullptr;
    void *nccl_comm = nullptr;
    int nccl_rank = 0, nccl_world = 1;
};

int nb200_fail(nb200_ctx *ctx, int code, const char *fmt, ...);
int nb200_cuda_fail(nb200_ctx *ctx, cudaError_t e, const char *what, const char *file, int line);

#define NB_CUDA(ctx, call)                                                             \
    do {                                                                               \
        cudaError_t _e = (call);                                                       \
        if (_e != cudaSuccess) return nb200_cuda_fail((ctx), _e, #call, __FILE__, __LINE__); \
    } while (0)

#define NB_CHECK(call)                 \
    do {                               \
        int _s = (call);               \
        if (_s != NB200_OK) return _s; \
    } while (0)

// profiling bracket around a kernel launch
struct ProfScope {
    nb200_ctx *ctx;
    cudaStream_t st;
    int idx;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(nb200_ctx *c, cudaStream_t s, const char *name);
    ~ProfScope();
};

void nb200_prof_collect(nb200_ctx *ctx);  // drain pending events (sync)

// ---------------------------------------------------------------------------------------------
// batch on the device
// ---------------------------------------------------------------------------------------------
struct nb200_dbatch {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    int n_chunks = 0;
    int64_t total_len = 0;   // sum of chunk lengths
    int64_t n_frag = 0;
    int64_t n_seq = 0;
    int64_t h2d_bytes = 0;
    bool have_seq = false;
    int max_len = 0;
    // geometry (host copies for launch configuration)
    std::vector<int32_t> h_start, h_end;
    std::vector<int64_t> h_out_off, h_frag_off, h_seq_off;
    std::vector<int32_t> h_seq_start;
    // device inputs
    DevBuf d_start, d_end, d_frag_off, d_pos, d_tlen, d_seq_off, d_seq_start, d_seq, d_out_off;
    // derived: bias track (log-bias b and E = exp(b)), offsets per chunk
    DevBuf d_bias_off;  // int64 [n+1]
    DevBuf d_bias0;     // int32 [n] genomic coordinate of b[bias_off[c]]
    DevBuf d_b, d_E;    // f64 packed
    std::vector<int64_t> h_bias_off;
    std::vector<int32_t> h_bias0;
    bool prep_done = false;
    // fragment matrix in CSC form (columns = genomic positions over [start-pad, end+pad))
    int csc_pad = 0, csc_upper = 0, csc_atac = -1;
    DevBuf d_col_off;   // int64 [n+1] offsets into col arrays (each chunk ncol+1 entries)
    std::vector<int64_t> h_col_off;
    DevBuf d_col_ptr;   // int32 packed [ncol+1] per chunk: all entries
    DevBuf d_cursor;    // int32 scratch same shape
    DevBuf d_ent;       // int2 {col,row} packed at frag_off
    // occ outputs (device)
    DevBuf o_vals, o_lower, o_upper, o_svals, o_slower, o_supper, o_cov, o_nuc_dist;
    DevBuf o_peak_count, o_peak_pos, o_peak_occ, o_peak_lower, o_peak_upper, o_peak_reads;
    DevBuf o_cn, o_cf;          // per-column sums of pn*Bp, pf*Bp
    DevBuf o_peak_off;          // int64 [n+1]
    std::vector<int64_t> h_opeak_off;
    bool occ_done = false;
    // nuc outputs (device)
    DevBuf n_signal, n_bg, n_norm, n_smooth, n_nuc_cov, n_nfr_cov, n_bx, n_colsum;
    DevBuf n_cand_count, n_cand_pos, n_cand_flag, n_cand_z, n_cand_lr, n_cand_norm, n_cand_sig, n_cand_cov,
        n_cand_nfr, n_cand_smooth;
    DevBuf n_cand_off;          // int64 [n+1]
    std::vector<int64_t> h_ncand_off;
    DevBuf n_work;              // candidate work list
    DevBuf n_work_count;
    // tensor-core path operands
    DevBuf t_a_hi, t_a_lo;      // fp16 materialised prenorm bias rows (when not generated in-kernel)
    bool nuc_done = false;
    DevBuf misc;
};

// kernels / stage drivers implemented across the .cu files
int nb200_prep_batch(nb200_ctx *ctx, nb200_dbatch *b, int pad, int upper, int atac, bool need_bias);
int nb200_background_fp64(nb200_ctx *ctx, nb200_dbatch *b);
int nb200_background_tc(nb200_ctx *ctx, nb200_dbatch *b);
int nb200_tc_setup(nb200_ctx *ctx);  // (re)build the band operand after vmat / sizes change

static inline int64_t div_up64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// floor division helpers matching Python semantics for the (i-1)//2 taps
__host__ __device__ static inline int floordiv2(int a) { return a >> 1; }  // arithmetic shift == floor for /2
