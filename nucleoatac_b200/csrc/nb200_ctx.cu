// nb200_ctx.cu -- context, run constants, profiling brackets, pinned memory, NCCL plumbing.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "nb200_common.cuh"

static std::string g_create_error;

int nb200_fail(nb200_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_error = buf;
    return code;
}

int nb200_cuda_fail(nb200_ctx *ctx, cudaError_t e, const char *what, const char *file, int line)
{
    return nb200_fail(ctx, NB200_ERR_CUDA, "CUDA error %s (%d) at %s:%d in %s", cudaGetErrorString(e), (int)e, file,
                      line, what);
}

ProfScope::ProfScope(nb200_ctx *c, cudaStream_t s, const char *name) : ctx(c), st(s), idx(-1)
{
    for (size_t i = 0; i < c->prof.size(); i++)
        if (c->prof[i].name == name) {
            idx = (int)i;
            break;
        }
    if (idx < 0) {
        ProfEntry e;
        e.name = name;
        c->prof.push_back(e);
        idx = (int)c->prof.size() - 1;
    }
    c->prof[idx].launches++;
    if (c->prof_on) {
        auto get = [&]() {
            cudaEvent_t ev;
            if (!c->ev_pool.empty()) {
                ev = c->ev_pool.back();
                c->ev_pool.pop_back();
            } else
                cudaEventCreate(&ev);
            return ev;
        };
        a = get();
        b = get();
        cudaEventRecord(a, st);
    }
}

ProfScope::~ProfScope()
{
    if (a) {
        cudaEventRecord(b, st);
        ctx->pending.push_back({idx, a, b});
    }
}

void nb200_prof_collect(nb200_ctx *ctx)
{
    for (auto &p : ctx->pending) {
        float ms = 0.f;
        cudaEventSynchronize(p.b);
        cudaEventElapsedTime(&ms, p.a, p.b);
        ctx->prof[p.idx].ms += ms;
        ctx->ev_pool.push_back(p.a);
        ctx->ev_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

int nb200_fifo_enter(nb200_ctx *ctx, cudaStream_t st)
{
    if (ctx->fifo) NB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_fifo, 0));   // a never-recorded event is complete
    return NB200_OK;
}

int nb200_fifo_leave(nb200_ctx *ctx, cudaStream_t st)
{
    if (ctx->fifo) NB_CUDA(ctx, cudaEventRecord(ctx->ev_fifo, st));
    return NB200_OK;
}

extern "C" {

int nb200_ctx_create(int device, nb200_ctx **out)
{
    if (!out) return nb200_fail(nullptr, NB200_ERR_ARG, "nb200_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return nb200_fail(nullptr, NB200_ERR_CUDA,
                          "nb200_ctx_create: no CUDA device available (%s); libnucleo_b200 has no CPU fallback",
                          cudaGetErrorString(e));
    if (device < 0 || device >= n) return nb200_fail(nullptr, NB200_ERR_ARG, "nb200_ctx_create: bad device %d", device);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return nb200_cuda_fail(nullptr, e, "cudaSetDevice", __FILE__, __LINE__);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return nb200_cuda_fail(nullptr, e, "cudaGetDeviceProperties", __FILE__, __LINE__);
    if (prop.major != 10)
        return nb200_fail(nullptr, NB200_ERR_CUDA,
                          "nb200_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                          prop.major, prop.minor);
    nb200_ctx *c = new nb200_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->hbm_bytes = prop.totalGlobalMem;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return nb200_cuda_fail(nullptr, e, "cudaStreamCreate", __FILE__, __LINE__);
    }
    e = cudaEventCreateWithFlags(&c->ev_fifo, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cudaStreamDestroy(c->stream);
        delete c;
        return nb200_cuda_fail(nullptr, e, "cudaEventCreate", __FILE__, __LINE__);
    }
    c->fifo = getenv("NB200_NO_FIFO") == nullptr;   // developer switch: let the passes of different batches time-slice
    *out = c;
    return NB200_OK;
}

int nb200_ctx_destroy(nb200_ctx *c)
{
    if (!c) return NB200_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    nb200_prof_collect(c);
    nb200_tc_release(c);
    for (auto ev : c->ev_pool) cudaEventDestroy(ev);
    RunConst &r = c->rc;
    DevBuf *bufs[] = {&r.log_pwm, &r.nuc_code, &r.vmat,   &r.vmat_fp, &r.sizes,  &r.nuc_probs, &r.nfr_probs,
                      &r.alphas,  &r.jitter,   &r.occ_win, &r.nuc_win, &c->s0,     &c->s1,    &c->s2,       &c->s3,
                      &c->s4,     &c->flush,   &r.vp_pair, &r.vp_one, &r.vp_pair32, &r.vp_one32};
    for (auto b : bufs) b->release();
    if (c->ev_fifo) cudaEventDestroy(c->ev_fifo);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return NB200_OK;
}

const char *nb200_last_error(nb200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int nb200_device_info(nb200_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, int64_t *hbm_bytes)
{
    if (!ctx) return NB200_ERR_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (hbm_bytes) *hbm_bytes = (int64_t)ctx->hbm_bytes;
    return NB200_OK;
}

int nb200_host_alloc(nb200_ctx *ctx, int64_t bytes, void **out)
{
    if (!ctx || !out || bytes < 0) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_host_alloc: bad argument");
    NB_CUDA(ctx, cudaHostAlloc(out, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault));
    return NB200_OK;
}

int nb200_host_free(nb200_ctx *ctx, void *p)
{
    if (p) NB_CUDA(ctx, cudaFreeHost(p));
    return NB200_OK;
}

// ---------------------------------------------------------------------------------------------
// run constants
// ---------------------------------------------------------------------------------------------
static int upload(nb200_ctx *ctx, DevBuf &dst, const void *src, size_t bytes)
{
    NB_CUDA(ctx, dst.reserve(bytes));
    NB_CUDA(ctx, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

static int rebuild_scaled_vmat(nb200_ctx *ctx)
{
    RunConst &r = ctx->rc;
    if (!r.have_vmat || !r.have_sizes) return NB200_OK;
    if (r.sizes_upper < r.v_upper)
        return nb200_fail(ctx, NB200_ERR_ARG, "fragment sizes cover [0,%d) but the VMat needs sizes up to %d",
                          r.sizes_upper, r.v_upper);
    // T = f_i * V, rows zero padded to a multiple of 16 columns (operand of the dense background xcor)
    r.v_wpad = (r.v_cols + 15) / 16 * 16;
    std::vector<double> vf((size_t)r.v_rows * r.v_wpad, 0.0);
    r.f_has_zero = false;
    r.f_sum_v = 0.0;
    r.f_max_v = 0.0;
    for (int i = 0; i < r.v_rows; i++) {
        double f = r.h_sizes[r.v_lower + i];
        if (!(f != 0.0)) r.f_has_zero = true;
        r.f_sum_v += f;
        if (!(f >= 0.0 && f < 1e300)) r.f_max_v = -1.0;
        else if (r.f_max_v >= 0.0 && f > r.f_max_v) r.f_max_v = f;
        for (int k = 0; k < r.v_cols; k++) vf[(size_t)i * r.v_wpad + k] = f * r.h_vmat[(size_t)i * r.v_cols + k];
    }
    NB_CHECK(upload(ctx, r.vmat_fp, vf.data(), vf.size() * sizeof(double)));
    {
        // Window sums sum_i T[i,k] Bp[i,c] regrouped by the left tap (sizes 2j+1 and 2j+2 share l = c - j):
        //   sum_j E[c-j] * (T[2j+1,k] E[c+j] + T[2j+2,k] E[c+j+1]);  size 0 has the taps of size 2, size 1 a single tap.
        const int uv = r.v_upper, lv = r.v_lower, W = r.v_cols;
        const int J2 = (std::max(1, uv / 2) + 1) & ~1, W2 = (W + 1) & ~1;
        std::vector<double> pair((size_t)3 * J2 * W2 * 2, 0.0), one((size_t)3 * W2, 0.0);
        auto T = [&](int t, int i, int k) -> double {
            if (i < lv || i >= uv) return 0.0;
            const double v = r.h_vmat[(size_t)(i - lv) * W + k], f = r.h_sizes[i];
            return t == 0 ? v : (t == 1 ? f * v : f * v * v);
        };
        for (int t = 0; t < 3; t++)
            for (int k = 0; k < W; k++) {
                one[(size_t)t * W2 + k] = T(t, 1, k);
                for (int j = 0; j < J2; j++) {
                    double *c = &pair[(((size_t)t * J2 + j) * W2 + k) * 2];
                    c[0] = j == 0 ? 0.0 : T(t, 2 * j + 1, k);
                    c[1] = T(t, 2 * j + 2, k) + (j == 0 ? T(t, 0, k) : 0.0);
                }
            }
        NB_CHECK(upload(ctx, r.vp_pair, pair.data(), pair.size() * sizeof(double)));
        NB_CHECK(upload(ctx, r.vp_one, one.data(), one.size() * sizeof(double)));
        // template 0 (V) once more in fp32 for the candidate screen
        std::vector<float> pair32((size_t)J2 * W2 * 2), one32((size_t)W2);
        for (size_t i = 0; i < pair32.size(); i++) pair32[i] = (float)pair[i];
        for (size_t i = 0; i < one32.size(); i++) one32[i] = (float)one[i];
        NB_CHECK(upload(ctx, r.vp_pair32, pair32.data(), pair32.size() * sizeof(float)));
        NB_CHECK(upload(ctx, r.vp_one32, one32.data(), one32.size() * sizeof(float)));
        r.vp_J2 = J2;
        r.vp_W2 = W2;
    }
    return nb200_tc_setup(ctx);
}

int nb200_set_pwm(nb200_ctx *ctx, const double *log_pwm, int n_nuc, int up, int down, const char *nucleotides)
{
    if (!ctx || !log_pwm || !nucleotides) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_set_pwm: NULL argument");
    int width = up + down + 1;
    if (n_nuc < 1 || n_nuc > NB200_MAX_NUC || width < 1 || width > NB200_MAX_PWM_WIDTH || (int)strlen(nucleotides) != n_nuc)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_set_pwm: unsupported PWM shape %d x %d", n_nuc, width);
    RunConst &r = ctx->rc;
    int8_t code[256];
    memset(code, -1, sizeof(code));
    for (int i = 0; i < n_nuc; i++) {
        unsigned char ch = (unsigned char)nucleotides[i];
        code[ch] = (int8_t)i;
        // the reference upper-cases the sequence (seq.py:22) before comparing with the PWM letters
        if (ch >= 'A' && ch <= 'Z') code[ch - 'A' + 'a'] = (int8_t)i;
    }
    NB_CHECK(upload(ctx, r.log_pwm, log_pwm, sizeof(double) * n_nuc * width));
    NB_CHECK(upload(ctx, r.nuc_code, code, 256));
    r.pwm_up = up;
    r.pwm_down = down;
    r.pwm_width = width;
    r.n_nuc = n_nuc;
    r.have_pwm = true;
    ctx->occ_gen++;
    return NB200_OK;
}

int nb200_set_vmat(nb200_ctx *ctx, const double *mat, int nrow, int ncol, int lower, int upper)
{
    if (!ctx || !mat) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_set_vmat: NULL argument");
    if (nrow != upper - lower)  // VMat.py:33-34
        return nb200_fail(ctx, NB200_ERR_ARG, "mat shape is not consistent with insert limits");
    if (nrow < 1 || ncol < 1 || (ncol % 2) != 1 || lower < 0)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_set_vmat: need an odd number of columns and lower >= 0");
    RunConst &r = ctx->rc;
    size_t n = (size_t)nrow * ncol;
    r.h_vmat.assign(mat, mat + n);
    r.v_has_zero = false;
    r.v_nonneg = true;
    for (size_t i = 0; i < n; i++) {
        if (mat[i] == 0.0) r.v_has_zero = true;
        if (!(mat[i] >= 0.0 && mat[i] < 1e300)) r.v_nonneg = false;
    }
    NB_CHECK(upload(ctx, r.vmat, mat, n * sizeof(double)));
    r.v_rows = nrow;
    r.v_cols = ncol;
    r.v_lower = lower;
    r.v_upper = upper;
    r.v_w = ncol / 2;
    r.have_vmat = true;
    return rebuild_scaled_vmat(ctx);
}

int nb200_set_fragment_sizes(nb200_ctx *ctx, const double *freq, int upper)
{
    if (!ctx || !freq || upper < 1) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_set_fragment_sizes: bad argument");
    RunConst &r = ctx->rc;
    r.h_sizes.assign(freq, freq + upper);
    NB_CHECK(upload(ctx, r.sizes, freq, sizeof(double) * upper));
    r.sizes_upper = upper;
    r.have_sizes = true;
    return rebuild_scaled_vmat(ctx);
}

int nb200_set_occ_model(nb200_ctx *ctx, const double *nuc_probs, const double *nfr_probs, int upper,
                        const double *alphas, int n_alpha, double cutoff)
{
    if (!ctx || !nuc_probs || !nfr_probs || !alphas || upper < 1 || n_alpha < 1 || n_alpha > 128)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_set_occ_model: bad argument (n_alpha must be in [1,128])");
    for (int i = 0; i + 1 < n_alpha; i++)
        if (alphas[i] == 1.0)
            return nb200_fail(ctx, NB200_ERR_ARG, "nb200_set_occ_model: alpha == 1 is only supported as the last grid value");
    RunConst &r = ctx->rc;
    NB_CHECK(upload(ctx, r.nuc_probs, nuc_probs, sizeof(double) * upper));
    NB_CHECK(upload(ctx, r.nfr_probs, nfr_probs, sizeof(double) * upper));
    NB_CHECK(upload(ctx, r.alphas, alphas, sizeof(double) * n_alpha));
    r.pn_has_zero = r.pf_has_zero = r.both_zero = 0;
    r.pn_sum = r.pf_sum = 0.0;
    for (int i = 0; i < upper; i++) {
        if (!(nuc_probs[i] != 0.0)) r.pn_has_zero = 1;
        if (!(nfr_probs[i] != 0.0)) r.pf_has_zero = 1;
        if (!(nuc_probs[i] != 0.0) && !(nfr_probs[i] != 0.0)) r.both_zero = 1;
        r.pn_sum += nuc_probs[i];
        r.pf_sum += nfr_probs[i];
    }
    r.occ_upper = upper;
    r.n_alpha = n_alpha;
    r.cutoff = cutoff;
    r.have_occ_model = true;
    ctx->occ_gen++;
    return NB200_OK;
}

int nb200_set_jitter(nb200_ctx *ctx, const double *jitter, int64_t n)
{
    if (!ctx || !jitter || n < 1) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_set_jitter: bad argument");
    NB_CHECK(upload(ctx, ctx->rc.jitter, jitter, sizeof(double) * (size_t)n));
    ctx->rc.n_jitter = n;
    return NB200_OK;
}

int nb200_occ_configure(nb200_ctx *ctx, const nb200_occ_params *p)
{
    if (!ctx || !p) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_occ_configure: NULL argument");
    if (p->upper < 1 || p->flank < 0 || p->step < 1 || p->sep < 1 || !p->smooth_win || p->smooth_len < 1 ||
        (p->smooth_len % 2) != 1)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_occ_configure: bad parameter");
    ctx->occ = *p;
    if (ctx->occ.step % 2 == 0) ctx->occ.step -= 1;  // Occupancy.py:190-191
    NB_CHECK(upload(ctx, ctx->rc.occ_win, p->smooth_win, sizeof(double) * p->smooth_len));
    ctx->occ.smooth_win = nullptr;
    ctx->occ_configured = true;
    ctx->occ_gen++;
    return NB200_OK;
}

int nb200_nuc_configure(nb200_ctx *ctx, const nb200_nuc_params *p)
{
    if (!ctx || !p) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_nuc_configure: NULL argument");
    if (!p->smooth_win || p->smooth_len < 1 || (p->smooth_len % 2) != 1 || p->redundant_sep < 1 ||
        p->nonredundant_sep < 1)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_nuc_configure: bad parameter");
    ctx->nuc = *p;
    NB_CHECK(upload(ctx, ctx->rc.nuc_win, p->smooth_win, sizeof(double) * p->smooth_len));
    ctx->nuc.smooth_win = nullptr;
    ctx->nuc_configured = true;
    return NB200_OK;
}

// ---------------------------------------------------------------------------------------------
// measurement
// ---------------------------------------------------------------------------------------------
int nb200_timer_start(nb200_ctx *ctx, nb200_dbatch *b)
{
    cudaStream_t st = b ? b->stream : ctx->stream;
    if (b) {
        if (!b->ev_start) NB_CUDA(ctx, cudaEventCreate(&b->ev_start));
        if (!b->ev_stop) NB_CUDA(ctx, cudaEventCreate(&b->ev_stop));
        NB_CUDA(ctx, cudaEventRecord(b->ev_start, st));
    }
    return NB200_OK;
}

int nb200_timer_stop(nb200_ctx *ctx, nb200_dbatch *b)
{
    if (!b || !b->ev_stop) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_timer_stop: timer not started");
    NB_CUDA(ctx, cudaEventRecord(b->ev_stop, b->stream));
    return NB200_OK;
}

int nb200_timer_elapsed_ms(nb200_ctx *ctx, nb200_dbatch *b, float *ms)
{
    if (!b || !b->ev_stop || !ms) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_timer_elapsed_ms: timer not started");
    NB_CUDA(ctx, cudaEventSynchronize(b->ev_stop));
    NB_CUDA(ctx, cudaEventElapsedTime(ms, b->ev_start, b->ev_stop));
    return NB200_OK;
}

int nb200_profile_enable(nb200_ctx *ctx, int on)
{
    nb200_prof_collect(ctx);
    ctx->prof_on = on != 0;
    return NB200_OK;
}

int nb200_profile_reset(nb200_ctx *ctx)
{
    nb200_prof_collect(ctx);
    for (auto &e : ctx->prof) {
        e.launches = 0;
        e.ms = 0.0;
    }
    return NB200_OK;
}

int nb200_profile_count(nb200_ctx *ctx)
{
    nb200_prof_collect(ctx);
    return (int)ctx->prof.size();
}

int nb200_profile_get(nb200_ctx *ctx, int i, const char **name, int64_t *launches, double *ms)
{
    if (i < 0 || i >= (int)ctx->prof.size()) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_profile_get: bad index");
    if (name) *name = ctx->prof[i].name.c_str();
    if (launches) *launches = ctx->prof[i].launches;
    if (ms) *ms = ctx->prof[i].ms;
    return NB200_OK;
}

__global__ void k_flush_fill(float4 *p, size_t n, float v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = make_float4(v, v, v, v);
}

int nb200_flush_l2(nb200_ctx *ctx, nb200_dbatch *b)
{
    const size_t bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    NB_CUDA(ctx, ctx->flush.reserve(bytes));
    cudaStream_t st = b ? b->stream : ctx->stream;
    k_flush_fill<<<ctx->sm_count * 4, 256, 0, st>>>(ctx->flush.as<float4>(), bytes / 16, 1.0f);
    NB_CUDA(ctx, cudaGetLastError());
    return NB200_OK;
}

// ---------------------------------------------------------------------------------------------
// NCCL (dlopen'd so the library loads without it; used only for end-of-run reductions)
// ---------------------------------------------------------------------------------------------
typedef struct {
    char internal[128];
} nccl_uid_t;
typedef int (*fn_getuid)(nccl_uid_t *);
typedef int (*fn_init)(void **, int, nccl_uid_t, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_destroy)(void *);

static void *nccl_open()
{
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (auto n : names) {
        void *h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) return h;
    }
    return nullptr;
}

int nb200_nccl_unique_id(void *out128)
{
    void *h = nccl_open();
    if (!h) return nb200_fail(nullptr, NB200_ERR_STATE, "libnccl.so.2 not found");
    fn_getuid f = (fn_getuid)dlsym(h, "ncclGetUniqueId");
    if (!f) return nb200_fail(nullptr, NB200_ERR_STATE, "ncclGetUniqueId not found");
    nccl_uid_t id;
    int s = f(&id);
    if (s != 0) return nb200_fail(nullptr, NB200_ERR_STATE, "ncclGetUniqueId failed (%d)", s);
    memcpy(out128, &id, 128);
    return NB200_OK;
}

int nb200_nccl_init(nb200_ctx *ctx, const void *unique_id128, int rank, int world)
{
    if (!ctx || !unique_id128) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_nccl_init: NULL argument");
    ctx->nccl_lib = nccl_open();
    if (!ctx->nccl_lib) return nb200_fail(ctx, NB200_ERR_STATE, "libnccl.so.2 not found");
    fn_init f = (fn_init)dlsym(ctx->nccl_lib, "ncclCommInitRank");
    if (!f) return nb200_fail(ctx, NB200_ERR_STATE, "ncclCommInitRank not found");
    nccl_uid_t id;
    memcpy(&id, unique_id128, 128);
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    int s = f(&ctx->nccl_comm, world, id, rank);
    if (s != 0) return nb200_fail(ctx, NB200_ERR_STATE, "ncclCommInitRank failed (%d)", s);
    ctx->nccl_rank = rank;
    ctx->nccl_world = world;
    return NB200_OK;
}

static int allreduce_impl(nb200_ctx *ctx, void *host, int64_t n, int nccl_dtype, size_t elt)
{
    if (!ctx->nccl_comm) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_allreduce: call nb200_nccl_init first");
    fn_allreduce f = (fn_allreduce)dlsym(ctx->nccl_lib, "ncclAllReduce");
    if (!f) return nb200_fail(ctx, NB200_ERR_STATE, "ncclAllReduce not found");
    NB_CUDA(ctx, ctx->s0.reserve((size_t)n * elt));
    NB_CUDA(ctx, cudaMemcpyAsync(ctx->s0.p, host, (size_t)n * elt, cudaMemcpyHostToDevice, ctx->stream));
    int s = f(ctx->s0.p, ctx->s0.p, (size_t)n, nccl_dtype, 0 /*ncclSum*/, ctx->nccl_comm, ctx->stream);
    if (s != 0) return nb200_fail(ctx, NB200_ERR_STATE, "ncclAllReduce failed (%d)", s);
    NB_CUDA(ctx, cudaMemcpyAsync(host, ctx->s0.p, (size_t)n * elt, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_allreduce_f64(nb200_ctx *ctx, double *host_inout, int64_t n) { return allreduce_impl(ctx, host_inout, n, 8 /*ncclFloat64*/, 8); }
int nb200_allreduce_i64(nb200_ctx *ctx, int64_t *host_inout, int64_t n) { return allreduce_impl(ctx, host_inout, n, 4 /*ncclInt64*/, 8); }

int nb200_nccl_finalize(nb200_ctx *ctx)
{
    if (ctx && ctx->nccl_comm) {
        fn_destroy f = (fn_destroy)dlsym(ctx->nccl_lib, "ncclCommDestroy");
        if (f) f(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return NB200_OK;
}

}  // extern "C"
