// api.cu -- C ABI of libfoto_b200.so (include/foto_b200.h): contexts, workspaces, the outer
// ALG2 loop of benamou_brenier.solve, the GN driver, operator application and the per-device
// work-queue batch driver.  No CPU fallback: every compute entry point needs a CUDA device.
#include <atomic>
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <thread>

#include "foto_kernels.cuh"

namespace foto {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace foto

using namespace foto;

// ------------------------------------------------------------------------------- context
struct DevResult {          // written by kernels, mirrored in pinned host memory
    double crit[2];         // numerator / denominator sums of the stopping criterion
    int cg_iters, cg_info;  // scipy-style (iterations, info)
    int error, pad;
    OuterState outer;       // ALG2 outer loop as decided on the device (k_outer_decide)
};
static const int kLookSlots = 4;                          // pinned mirrors / events of the last outer iterations

enum Cat { CAT_RHS = 0, CAT_CG, CAT_PROX, CAT_FLOW, CAT_GN, CAT_COUNT };

struct foto_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;          // stream every kernel of the context is launched on
    cudaStream_t own_stream = nullptr;      // created with the context; `stream` may be redirected (foto_ctx_set_stream)
    char *ws = nullptr;     size_t ws_bytes = 0;      // solver workspace
    char *io = nullptr;     size_t io_bytes = 0;      // staging for the host-buffer API
    unsigned int *sync_counter = nullptr;
    double *sync_partials = nullptr;
    double *prox_partials = nullptr;
    DevResult *d_res = nullptr;
    DevResult *h_res = nullptr;                       // pinned, kLookSlots entries ([0]: fetch_result)
    cudaEvent_t look_ev[kLookSlots] = {};
    char *d_trace = nullptr, *h_trace = nullptr;      // per outer iteration: crit (double), cg_iters, cg_info (int)
    int trace_cap = 0;
    int cg_grid = 0, cg_block = 0, gn_grid = 0, gn_block = 0;
    int cg_variant = -1;
    bool profiling = false;
    foto_stats stats{};
    std::vector<cudaEvent_t> ev_pool;
    struct Span { cudaEvent_t a, b; int cat; };
    std::vector<Span> spans;
    cudaEvent_t open_a = nullptr; int open_cat = -1;
    cudaEvent_t watch[2] = {nullptr, nullptr};
    OnchipScratch onchip;
    DctTables dct;
    DctTables gn_dct64;                               // fp64 DCT tables of the GN image shape (source of gn_tb)
    GnDctTables gn_tb;
    char *gn_state = nullptr, *h_gn_state = nullptr;  // device state of the spectral GN solve + pinned mirrors (kLookSlots)
    double *metric_partials = nullptr;                // foto_flow_metrics_dev / foto_warp_dev
    char *warp_tmp = nullptr; size_t warp_tmp_bytes = 0;
    char *pin = nullptr; size_t pin_bytes = 0;        // pinned staging of foto_solve_batch_u8
};

static const int kProxMaxBlocks = 148 * 8;

static int ctx_bind(const foto_ctx *c)
{
    CUDA_TRY(cudaSetDevice(c->device));
    return FOTO_OK;
}

static int ensure(char **buf, size_t *have, size_t need)
{
    if (*have >= need) return FOTO_OK;
    if (*buf) { CUDA_TRY(cudaFree(*buf)); *buf = nullptr; *have = 0; }
    need = (need + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    CUDA_TRY(cudaMalloc((void **)buf, need));
    *have = need;
    return FOTO_OK;
}

struct Carver {             // 256-byte aligned sub-allocation of a workspace
    char *base; size_t off = 0;
    explicit Carver(char *b) : base(b) {}
    double *take(size_t n) { double *p = (double *)(base + off); off += (n * sizeof(double) + 255) & ~size_t(255); return p; }
    static size_t bytes(size_t n) { return (n * sizeof(double) + 255) & ~size_t(255); }
};

static void prof_begin(foto_ctx *c, int cat)
{
    if (!c->profiling) return;
    cudaEvent_t e;
    if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); } else cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    c->open_a = e; c->open_cat = cat;
}

static void prof_end(foto_ctx *c)
{
    if (!c->profiling || !c->open_a) return;
    cudaEvent_t e;
    if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); } else cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    c->spans.push_back({c->open_a, e, c->open_cat});
    c->open_a = nullptr;
}

static void prof_resolve(foto_ctx *c)      // call after the stream has been synchronised
{
    for (auto &s : c->spans) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.a, s.b);
        switch (s.cat) {
        case CAT_RHS: c->stats.rhs_ms += ms; break;
        case CAT_CG: c->stats.cg_ms += ms; break;
        case CAT_PROX: c->stats.prox_ms += ms; break;
        case CAT_FLOW: c->stats.flow_ms += ms; break;
        case CAT_GN: c->stats.gn_ms += ms; break;
        }
        c->ev_pool.push_back(s.a); c->ev_pool.push_back(s.b);
    }
    c->spans.clear();
}

extern "C" const char *foto_last_error(void) { return g_err; }
extern "C" int foto_version(void) { return 100; }

extern "C" int foto_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return FOTO_ERR_NODEV; }
    return n;
}

extern "C" int foto_ctx_create(int device, foto_ctx **out)
{
    if (!out) { set_error("foto_ctx_create: out is NULL"); return FOTO_ERR_ARG; }
    *out = nullptr;
    int n = foto_device_count();
    if (n <= 0) { if (n == 0) set_error("no CUDA device: libfoto_b200 has no CPU fallback"); return FOTO_ERR_NODEV; }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return FOTO_ERR_ARG; }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libfoto_b200 is built for sm_100a only", device, prop.major, prop.minor);
        return FOTO_ERR_NODEV;
    }
    foto_ctx *c = new foto_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    int rc = [&]() -> int {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        CUDA_TRY(cudaMalloc((void **)&c->sync_counter, 256));
        CUDA_TRY(cudaMalloc((void **)&c->sync_partials, sizeof(double) * 2 * kMaxVals * kMaxBlocks));
        CUDA_TRY(cudaMalloc((void **)&c->prox_partials, sizeof(double) * 2 * kProxMaxBlocks));
        CUDA_TRY(cudaMalloc((void **)&c->d_res, sizeof(DevResult)));
        CUDA_TRY(cudaMemset(c->d_res, 0, sizeof(DevResult)));
        CUDA_TRY(cudaMemset(c->sync_counter, 0, 256));
        CUDA_TRY(cudaMallocHost((void **)&c->h_res, kLookSlots * sizeof(DevResult)));
        for (int i = 0; i < kLookSlots; i++) CUDA_TRY(cudaEventCreateWithFlags(&c->look_ev[i], cudaEventDisableTiming));
        FOTO_TRY(cg_stream_config(device, &c->cg_grid, &c->cg_block));
        FOTO_TRY(gn_pcg_config(device, &c->gn_grid, &c->gn_block));
        return FOTO_OK;
    }();
    if (rc != FOTO_OK) { foto_ctx_destroy(c); return rc; }
    *out = c;
    return FOTO_OK;
}

extern "C" void foto_ctx_destroy(foto_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->stream = c->own_stream;
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (auto e : c->watch) if (e) cudaEventDestroy(e);
    for (auto e : c->look_ev) if (e) cudaEventDestroy(e);
    cudaFree(c->d_trace);
    if (c->h_trace) cudaFreeHost(c->h_trace);
    cudaFree(c->ws); cudaFree(c->io); cudaFree(c->sync_counter); cudaFree(c->sync_partials);
    cudaFree(c->prox_partials); cudaFree(c->d_res); cudaFree(c->metric_partials); cudaFree(c->warp_tmp);
    cudaFree(c->gn_dct64.base); cudaFree(c->gn_tb.base); cudaFree(c->gn_state);
    if (c->h_gn_state) cudaFreeHost(c->h_gn_state);
    if (c->pin) cudaFreeHost(c->pin);
    onchip_release(c->onchip);
    cudaFree(c->dct.base);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int foto_ctx_device(const foto_ctx *c) { return c ? c->device : FOTO_ERR_ARG; }
extern "C" int foto_ctx_set_profiling(foto_ctx *c, int on) { if (!c) return FOTO_ERR_ARG; c->profiling = on != 0; return FOTO_OK; }
extern "C" int foto_ctx_reset_stats(foto_ctx *c) { if (!c) return FOTO_ERR_ARG; c->stats = foto_stats{}; return FOTO_OK; }
extern "C" int foto_ctx_get_stats(foto_ctx *c, foto_stats *out) { if (!c || !out) return FOTO_ERR_ARG; *out = c->stats; return FOTO_OK; }
extern "C" int foto_ctx_set_cg_variant(foto_ctx *c, int v)
{
    if (!c || v < -1 || v > 3 || v == 1) { set_error("cg variant must be -1 (auto), 0 (streaming), 2 (on-chip single-reduction) or 3 (GN: spectral preconditioner)"); return FOTO_ERR_ARG; }
    c->cg_variant = v;
    return FOTO_OK;
}

extern "C" int foto_ctx_event_record(foto_ctx *c, int which)
{
    if (!c || which < 0 || which > 1) { set_error("foto_ctx_event_record: bad argument"); return FOTO_ERR_ARG; }
    FOTO_TRY(ctx_bind(c));
    if (!c->watch[which]) CUDA_TRY(cudaEventCreate(&c->watch[which]));
    CUDA_TRY(cudaEventRecord(c->watch[which], c->stream));
    return FOTO_OK;
}

extern "C" int foto_ctx_event_elapsed_ms(foto_ctx *c, double *ms)
{
    if (!c || !ms || !c->watch[0] || !c->watch[1]) { set_error("foto_ctx_event_elapsed_ms: record start and stop first"); return FOTO_ERR_ARG; }
    FOTO_TRY(ctx_bind(c));
    CUDA_TRY(cudaEventSynchronize(c->watch[1]));
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, c->watch[0], c->watch[1]));
    *ms = f;
    return FOTO_OK;
}

// Debugging aid: cycle counters of every CTA of the on-chip CG kernel (cg_fused.cu), accumulated over launches:
// [0] halo import, [1] stencil, [2] all-reduce (+ x/2), [3] p, s, r update + edge export, [4] x/2,
// [6] iterations.  enable != 0 (re)starts counting; out may be NULL.
extern "C" int foto_debug_onchip_prof(foto_ctx *c, int enable, long long *out)
{
    if (!c) return FOTO_ERR_ARG;
    FOTO_TRY(ctx_bind(c));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (out && c->onchip.prof) CUDA_TRY(cudaMemcpy(out, c->onchip.prof, 8 * 1024 * sizeof(long long), cudaMemcpyDeviceToHost));
    else if (out) memset(out, 0, 8 * 1024 * sizeof(long long));
    if (enable) {
        if (!c->onchip.prof) CUDA_TRY(cudaMalloc((void **)&c->onchip.prof, 8 * 1024 * sizeof(long long)));
        CUDA_TRY(cudaMemset(c->onchip.prof, 0, 8 * 1024 * sizeof(long long)));
    } else if (c->onchip.prof) {
        CUDA_TRY(cudaFree(c->onchip.prof));
        c->onchip.prof = nullptr;
    }
    return FOTO_OK;
}

// ------------------------------------------------------------------------------- helpers
static int make_dims(int Nt, int Nx, int Ny, Dims *d)
{
    if (Nt < 2 || Nx < 2 || Ny < 2) { set_error("grid must be at least 2 in every direction (Nt=%d Nx=%d Ny=%d)", Nt, Nx, Ny); return FOTO_ERR_ARG; }
    unsigned long long P = (unsigned long long)Nx * Ny, N = P * Nt;
    if (3ull * N >= (1ull << 32)) { set_error("grid too large for 32-bit cell indices (3N = %llu)", 3ull * N); return FOTO_ERR_ARG; }
    d->Nt = Nt; d->Ny = Ny; d->Nx = Nx; d->P = (unsigned int)P; d->N = (unsigned int)N;
    d->n0 = 0; d->gNt = Nt; d->cs = (unsigned int)N;
    return FOTO_OK;
}

static int fetch_result(foto_ctx *c)
{
    CUDA_TRY(cudaMemcpyAsync(c->h_res, c->d_res, sizeof(DevResult), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->h_res->error) {
        cudaMemsetAsync(&c->d_res->error, 0, sizeof(int), c->stream);
        set_error("grid-barrier watchdog fired inside a persistent kernel");
        return FOTO_ERR_TIMEOUT;
    }
    return FOTO_OK;
}

static int ensure_dct_tables(foto_ctx *c, const Dims &d)
{
    DctTables &t = c->dct;
    const bool split = d.Nx % 4 == 0 && d.Ny % 4 == 0 && getenv("FOTO_DCT_DENSE") == nullptr;     // even / odd folded x and y transforms
    // second folding level: from 1 M pixels per plane on (1080x1920x16: 7.0 -> 5.7 ms per solve; 388x584x4: 0.154 -> 0.171 ms, the
    // extra launches cost more than the smaller GEMMs save).  FOTO_DCT_LEVELS=1|2 forces it (A/B, tests).
    const char *lv = getenv("FOTO_DCT_LEVELS");
    const int want2 = lv ? (atoi(lv) >= 2 ? 2 : 1) : ((long long)d.Nx * d.Ny >= (1ll << 20) ? 2 : 1);
    if (t.base && t.Nt == d.Nt && t.Ny == d.Ny && t.Nx == d.Nx && t.split == split &&
        (!split || ((d.Nx % 8 ? 1 : want2) == t.lx && (d.Ny % 8 ? 1 : want2) == t.ly))) return FOTO_OK;
    if (t.base) { CUDA_TRY(cudaFree(t.base)); t = DctTables(); }
    const int n[3] = {d.Nx, d.Ny, d.Nt};
    std::vector<double> host;
    std::vector<size_t> off;
    int level[3] = {1, 1, 1};
    auto push = [&](const std::vector<double> &v) {
        if (host.size() & 1) host.push_back(0.0);        // every table starts on a 16-byte boundary
        off.push_back(host.size()); host.insert(host.end(), v.begin(), v.end());
    };
    for (int a = 0; a < 3; a++) {
        std::vector<double> C, Ct, lam;
        dct_host_tables(n[a], C, Ct, lam);
        push(C); push(Ct); push(lam);
        if (split && a < 2) {
            std::vector<double> E, ET, lam_p;
            dct_host_folded(n[a], C, lam, E, ET, lam_p);
            const bool two = n[a] % 8 == 0 && want2 == 2;
            std::vector<double> E2, E2T, lam_p2;
            if (two) dct_host_folded2(n[a], C, lam, E2, E2T, lam_p2);
            push(E); push(ET); push(two ? lam_p2 : lam_p);
            if (two) { push(E2); push(E2T); }
            level[a] = two ? 2 : 1;
        }
    }
    const size_t total = host.size();
    CUDA_TRY(cudaMalloc((void **)&t.base, total * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(t.base, host.data(), total * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));          // host vector goes out of scope
    int k = 0;
    t.Cx = t.base + off[k++]; t.CxT = t.base + off[k++]; t.lam_x = t.base + off[k++];
    if (split) { t.Ex = t.base + off[k++]; t.ExT = t.base + off[k++]; t.lam_xp = t.base + off[k++]; }
    if (split && level[0] == 2) { t.E2x = t.base + off[k++]; t.E2xT = t.base + off[k++]; }
    t.Cy = t.base + off[k++]; t.CyT = t.base + off[k++]; t.lam_y = t.base + off[k++];
    if (split) { t.Ey = t.base + off[k++]; t.EyT = t.base + off[k++]; t.lam_yp = t.base + off[k++]; }
    if (split && level[1] == 2) { t.E2y = t.base + off[k++]; t.E2yT = t.base + off[k++]; }
    t.Ct = t.base + off[k++]; t.CtT = t.base + off[k++]; t.lam_t = t.base + off[k++];
    t.split = split; t.lx = level[0]; t.ly = level[1];
    t.Nt = d.Nt; t.Ny = d.Ny; t.Nx = d.Nx;
    return FOTO_OK;
}

// One Poisson solve A phi = F on device buffers.  work: 4N doubles (r, p0, p1, q).
static int run_cg(foto_ctx *c, const Dims &d, const double *F, double *phi, double *work, double r, double eps,
                  int backend)
{
    if (backend == FOTO_POISSON_DCT_EXACT) {
        FOTO_TRY(ensure_dct_tables(c, d));
        prof_begin(c, CAT_CG);
        FOTO_TRY(launch_poisson_dct(c->stream, c->dct, d.Nt, d.Ny, d.Nx, r, eps, F, phi, work, work + d.N));
        prof_end(c);
        // report "0 iterations, converged" through the same result block the CG kernels write
        CUDA_TRY(cudaMemsetAsync(&c->d_res->cg_iters, 0, 2 * sizeof(int), c->stream));
        c->stats.cg_variant = 2;
        c->stats.launches += c->dct.split ? 9 : 5; c->stats.cg_launches++;
        return FOTO_OK;
    }
    CgArgs a;
    a.b = F; a.x = phi;
    a.r = work; a.p0 = work + d.N; a.p1 = work + 2ull * d.N; a.q = work + 3ull * d.N;
    a.Nt = d.Nt; a.Ny = d.Ny; a.Nx = d.Nx;
    a.rcoef = r; a.eps = eps;
    if (backend == FOTO_POISSON_CG_PARITY) { a.rtol = 1e-6; a.maxiter = 1000; }          // benamou_brenier.py:85
    else if (backend == FOTO_POISSON_CG_TIGHT) { a.rtol = 1e-13; a.maxiter = 100000; }
    else { set_error("unknown Poisson back-end %d", backend); return FOTO_ERR_ARG; }
    a.sync.counter = c->sync_counter; a.sync.partials = c->sync_partials; a.sync.error = &c->d_res->error;
    a.out = &c->d_res->cg_iters;
    a.skip = d.skip;
    // kernel choice: 0 streaming (textbook recurrences), 2 on-chip single-reduction, -1 auto = on-chip when the grid
    // fits (only for the truncated cg_parity solve the single-reduction form was validated on)
    const bool fits_fused = cg_fused_fits(c->onchip, c->device, d.Nt, d.Ny, d.Nx);
    if (c->cg_variant == 2 && !fits_fused) { set_error("grid %dx%dx%d does not fit the single-reduction on-chip CG variant", d.Nt, d.Ny, d.Nx); return FOTO_ERR_ARG; }
    const bool want_auto = c->cg_variant == -1 || c->cg_variant == 3;       // 3 selects a GN solver only
    const int kind = (c->cg_variant == 2 || (want_auto && fits_fused && backend == FOTO_POISSON_CG_PARITY)) ? 3 : 0;
    prof_begin(c, CAT_CG);
    if (kind == 3) FOTO_TRY(launch_cg_fused(c->stream, a, c->device, c->onchip));
    else FOTO_TRY(launch_cg_stream(c->stream, a, c->cg_grid, c->cg_block));
    prof_end(c);
    c->stats.cg_variant = kind;
    c->stats.launches++; c->stats.cg_launches++;
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

static size_t solve_ws_bytes(const Dims &d)
{
    return 2 * Carver::bytes(3ull * d.N) + 2 * Carver::bytes(d.N) + Carver::bytes(4ull * d.N);
}

// ------------------------------------------------------------------------------- FOTO solve
extern "C" int foto_solve_dev(foto_ctx *c, const double *d_rho0, const double *d_rhoT, int Nt, int Nx, int Ny,
                              double r, double tol, double eps, int max_it, int backend, double *d_u, double *d_v,
                              double *d_m, double *crit_trace, int *n_outer, int *cg_iters, int *cg_info)
{
    if (!c || !d_rho0 || !d_rhoT || !d_u || !d_v || !d_m) { set_error("foto_solve_dev: NULL argument"); return FOTO_ERR_ARG; }
    if (max_it < 1) { set_error("max_it must be >= 1"); return FOTO_ERR_ARG; }
    Dims d;
    FOTO_TRY(make_dims(Nt, Nx, Ny, &d));
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->ws, &c->ws_bytes, solve_ws_bytes(d)));
    Carver cv(c->ws);
    double *mu = cv.take(3ull * d.N), *q = cv.take(3ull * d.N), *F = cv.take(d.N), *phi = cv.take(d.N);
    double *work = cv.take(4ull * d.N);          // CG vectors r, p0, p1, q (contiguous, stride N)

    // trace of the outer loop on the device + pinned mirror
    if (c->trace_cap < max_it) {
        if (c->d_trace) { CUDA_TRY(cudaFree(c->d_trace)); c->d_trace = nullptr; }
        if (c->h_trace) { CUDA_TRY(cudaFreeHost(c->h_trace)); c->h_trace = nullptr; }
        c->trace_cap = 0;
        const int cap = max_it < 128 ? 128 : max_it;
        CUDA_TRY(cudaMalloc((void **)&c->d_trace, (size_t)cap * 16));
        CUDA_TRY(cudaMallocHost((void **)&c->h_trace, (size_t)cap * 16));
        c->trace_cap = cap;
    }
    OuterTrace tr;
    tr.crit = (double *)c->d_trace; tr.cg_iters = (int *)(c->d_trace + (size_t)c->trace_cap * 8); tr.cg_info = tr.cg_iters + c->trace_cap;
    CUDA_TRY(cudaMemsetAsync(&c->d_res->outer, 0, sizeof(OuterState), c->stream));

    launch_init_state(c->stream, d, d_rho0, d_rhoT, mu, q);
    c->stats.launches++;
    // The ALG2 loop (benamou_brenier.py:204-258).  The stopping rule is evaluated on the device (k_outer_decide), and
    // the host stays `look` iterations ahead: iteration it+1 is enqueued before the result of iteration it is known
    // and its kernels return at once if iteration it stopped the loop (Dims::skip), so the device never waits for a
    // host round trip between outer iterations.  dct_exact (several library kernels per solve) keeps look = 0.
    const int look = backend == FOTO_POISSON_DCT_EXACT ? 0 : 1;
    d.skip = look ? &c->d_res->outer.done : nullptr;
    int outer = 0, enq = 0;
    bool done = false;
    auto check = [&](int it) -> int {                     // result of outer iteration `it` (waits for it)
        DevResult *h = c->h_res + (it % kLookSlots);
        CUDA_TRY(cudaEventSynchronize(c->look_ev[it % kLookSlots]));
        if (h->error) {
            cudaMemsetAsync(&c->d_res->error, 0, sizeof(int), c->stream);
            set_error("grid-barrier watchdog fired inside a persistent kernel");
            return FOTO_ERR_TIMEOUT;
        }
        if (h->outer.done) { done = true; outer = h->outer.n_outer; }
        return FOTO_OK;
    };
    for (int it = 0; it < max_it && !done; it++) {
        if (it - look - 1 >= 0) FOTO_TRY(check(it - look - 1));
        if (done) break;
        prof_begin(c, CAT_RHS);
        launch_rhs(c->stream, d, mu, q, d_rho0, d_rhoT, r, F);                 // stepA, right-hand side
        prof_end(c);
        FOTO_TRY(run_cg(c, d, F, phi, work, r, eps, backend));                 // stepA, Poisson solve
        prof_begin(c, CAT_PROX);
        int blocks = launch_prox_dual(c->stream, d, phi, mu, q, r, c->prox_partials, kProxMaxBlocks, c->num_sms, &c->stats.prox_variant);   // stepB + stepC
        if (blocks < 0) return FOTO_ERR_CUDA;
        launch_outer_decide(c->stream, c->prox_partials, blocks, c->d_res->crit, &c->d_res->cg_iters, &c->d_res->outer, tr,
                            it, tol, max_it);
        prof_end(c);
        c->stats.launches += 3;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(c->h_res + (it % kLookSlots), c->d_res, sizeof(DevResult), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaEventRecord(c->look_ev[it % kLookSlots], c->stream));
        enq = it + 1;
    }
    for (int it = enq - look - 1 < 0 ? 0 : enq - look - 1; it < enq && !done; it++) FOTO_TRY(check(it));
    if (!done) { set_error("outer loop ended without a decision"); return FOTO_ERR_CUDA; }
    CUDA_TRY(cudaMemcpyAsync(c->h_trace, c->d_trace, (size_t)c->trace_cap * 16, cudaMemcpyDeviceToHost, c->stream));
    if (n_outer) *n_outer = outer;
    prof_begin(c, CAT_FLOW);
    launch_flow(c->stream, d, phi, d_u, d_v, d_m);
    prof_end(c);
    c->stats.launches += 2;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    prof_resolve(c);
    {
        const double *hc = (const double *)c->h_trace;
        const int *hi = (const int *)(c->h_trace + (size_t)c->trace_cap * 8), *hf = hi + c->trace_cap;
        for (int it = 0; it < outer; it++) {
            c->stats.cg_iterations += hi[it];
            c->stats.cg_cells += (long long)hi[it] * d.N;
            if (crit_trace) crit_trace[it] = hc[it];
            if (cg_iters) cg_iters[it] = hi[it];
            if (cg_info) cg_info[it] = hf[it];
        }
        c->stats.rhs_cells += (long long)outer * d.N; c->stats.prox_cells += (long long)outer * d.N;
    }
    return FOTO_OK;
}

// ------------------------------------------------------------------------------- GN solve
static int ensure_dct_tables_2d(foto_ctx *c, DctTables &t, int h, int w)
{
    if (t.base && t.Ny == h && t.Nx == w) return FOTO_OK;
    if (t.base) { CUDA_TRY(cudaFree(t.base)); t = DctTables(); }
    const int n[2] = {w, h};
    std::vector<double> host;
    std::vector<size_t> off;
    for (int a = 0; a < 2; a++) {
        std::vector<double> C, Ct, lam;
        dct_host_tables(n[a], C, Ct, lam);
        off.push_back(host.size()); host.insert(host.end(), C.begin(), C.end());
        off.push_back(host.size()); host.insert(host.end(), Ct.begin(), Ct.end());
        off.push_back(host.size()); host.insert(host.end(), lam.begin(), lam.end());
    }
    CUDA_TRY(cudaMalloc((void **)&t.base, host.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(t.base, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));          // host vector goes out of scope
    t.Cx = t.base + off[0]; t.CxT = t.base + off[1]; t.lam_x = t.base + off[2];
    t.Cy = t.base + off[3]; t.CyT = t.base + off[4]; t.lam_y = t.base + off[5];
    t.Nt = 0; t.Ny = h; t.Nx = w;
    return FOTO_OK;
}

// classical.py:113-130 with the spectral preconditioner (gn_dct.cu): iterations are enqueued in chunks, the host
// reads the device's `done` flag one chunk behind
static int gn_solve_spectral(foto_ctx *c, const double *d_f1, const double *d_f2, int w, int h, double alpha, double lambda,
                             double rtol, int max_it, double *d_u, double *d_v, double *d_m, int *iters, int *info)
{
    const size_t P = (size_t)w * h;
    FOTO_TRY(ensure_dct_tables_2d(c, c->gn_dct64, h, w));
    FOTO_TRY(gn_dct_prepare_tables(c->stream, c->gn_dct64, w, h, c->gn_tb));
    if (!c->gn_state) {
        CUDA_TRY(cudaMalloc((void **)&c->gn_state, 256));
        CUDA_TRY(cudaMallocHost((void **)&c->h_gn_state, kLookSlots * 256));
    }
    const size_t n32 = (c->gn_tb.volume_floats() + 1) / 2;                  // one padded fp32 volume, in doubles
    FOTO_TRY(ensure(&c->ws, &c->ws_bytes, 2 * Carver::bytes(P) + 7 * Carver::bytes(3 * P) + Carver::bytes(8) + Carver::bytes(6 * 1184) +
                                              Carver::bytes(3 * 592) + 4 * Carver::bytes(n32)));
    Carver cv(c->ws);
    double *fx = cv.take(P), *fy = cv.take(P), *dinv = cv.take(3 * P), *b = cv.take(3 * P);
    GnDctArgs a;
    a.fx = fx; a.fy = fy; a.f2 = d_f2; a.b = b; a.lam_x = c->gn_dct64.lam_x; a.lam_y = c->gn_dct64.lam_y; a.tb = &c->gn_tb;
    a.x = cv.take(3 * P); a.r = cv.take(3 * P); a.p = cv.take(3 * P); a.s = cv.take(3 * P); a.wv = cv.take(3 * P);
    a.gbar = cv.take(8); a.partials6 = cv.take(6 * 1184); a.partials3 = cv.take(3 * 592);
    a.r32 = (float *)cv.take(n32); a.t1 = (float *)cv.take(n32); a.t2 = (float *)cv.take(n32); a.u32 = (float *)cv.take(n32);
    a.state = c->gn_state; a.w = w; a.h = h; a.maxiter = max_it; a.alpha = alpha; a.lam = lambda; a.rtol = rtol;
    launch_gn_coeffs(c->stream, w, h, d_f1, d_f2, alpha, lambda, fx, fy, dinv, b);
    prof_begin(c, CAT_GN);
    FOTO_TRY(gn_dct_begin(c->stream, a));
    c->stats.launches += 4;
    const int chunk = 8;
    int launches = 0, done = 0, n_it = 0, inf = 0, enq = 0, nchunk = 0;
    auto check = [&](int ci) -> int {
        CUDA_TRY(cudaEventSynchronize(c->look_ev[ci % kLookSlots]));
        gn_dct_read_state(c->h_gn_state + (ci % kLookSlots) * 256, &done, &n_it, &inf);
        return FOTO_OK;
    };
    while (!done && enq < max_it) {
        if (nchunk >= 2) FOTO_TRY(check(nchunk - 2));
        if (done) break;
        const int cnt = max_it - enq < chunk ? max_it - enq : chunk;
        FOTO_TRY(gn_dct_enqueue_iterations(c->stream, a, enq, cnt, &launches));
        CUDA_TRY(cudaMemcpyAsync(c->h_gn_state + (nchunk % kLookSlots) * 256, c->gn_state, gn_dct_state_bytes(), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaEventRecord(c->look_ev[nchunk % kLookSlots], c->stream));
        enq += cnt; nchunk++;
    }
    for (int ci = nchunk - 2 < 0 ? 0 : nchunk - 2; ci < nchunk && !done; ci++) FOTO_TRY(check(ci));
    if (!done) { set_error("spectral GN solve ended without a decision"); return FOTO_ERR_CUDA; }
    gn_dct_copy_out(c->stream, a, d_u, d_v, d_m);
    prof_end(c);
    c->stats.launches += launches + 1; c->stats.gn_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    prof_resolve(c);
    c->stats.gn_iterations += n_it;
    c->stats.gn_pixels += (long long)n_it * (long long)P;
    if (iters) *iters = n_it;
    if (info) *info = inf;
    return FOTO_OK;
}

extern "C" int foto_gn_solve_dev(foto_ctx *c, const double *d_f1, const double *d_f2, int w, int h, double alpha,
                                 double lambda, double rtol, int max_it, double *d_u, double *d_v, double *d_m,
                                 int *iters, int *info)
{
    if (!c || !d_f1 || !d_f2 || !d_u || !d_v || !d_m) { set_error("foto_gn_solve_dev: NULL argument"); return FOTO_ERR_ARG; }
    if (w < 2 || h < 2) { set_error("image must be at least 2x2"); return FOTO_ERR_ARG; }
    unsigned long long P = (unsigned long long)w * h;
    if (3ull * P >= (1ull << 31)) { set_error("image too large"); return FOTO_ERR_ARG; }
    if (rtol <= 0) rtol = 1e-13;
    if (max_it <= 0) max_it = 20000;
    FOTO_TRY(ctx_bind(c));
    // solver choice: 0 streaming Jacobi-PCG, 2 on-chip single-reduction Jacobi-PCG, 3 spectral preconditioner (TF32
    // tensor-core DCT, fp64 recurrences), -1 auto = spectral from 64 x 64 pixels on, else on-chip / streaming
    if (c->cg_variant == 3 || (c->cg_variant == -1 && P >= 4096 && w <= 4096 && h <= 4096))
        return gn_solve_spectral(c, d_f1, d_f2, w, h, alpha, lambda, rtol, max_it, d_u, d_v, d_m, iters, info);
    FOTO_TRY(ensure(&c->ws, &c->ws_bytes, 2 * Carver::bytes(P) + 8 * Carver::bytes(3 * P)));
    Carver cv(c->ws);
    GnArgs a;
    double *fx = cv.take(P), *fy = cv.take(P), *dinv = cv.take(3 * P), *b = cv.take(3 * P);
    a.fx = fx; a.fy = fy; a.f2 = d_f2; a.dinv = dinv; a.b = b;
    a.x = cv.take(3 * P); a.r = cv.take(3 * P); a.z = cv.take(3 * P);
    a.p0 = cv.take(3 * P); a.p1 = cv.take(3 * P); a.q = cv.take(3 * P);
    a.w = w; a.h = h; a.alpha = alpha; a.lam = lambda; a.rtol = rtol; a.maxiter = max_it;
    a.sync.counter = c->sync_counter; a.sync.partials = c->sync_partials; a.sync.error = &c->d_res->error;
    a.out = &c->d_res->cg_iters;
    launch_gn_coeffs(c->stream, w, h, d_f1, d_f2, alpha, lambda, fx, fy, dinv, b);
    prof_begin(c, CAT_GN);
    const bool gnf_fits = gn_fused_fits(c->onchip, c->device, h, w);
    if (c->cg_variant == 2 && !gnf_fits) { set_error("image %dx%d does not fit the single-reduction on-chip GN variant", h, w); return FOTO_ERR_ARG; }
    if (c->cg_variant == 2 || (c->cg_variant == -1 && gnf_fits)) { FOTO_TRY(launch_gn_fused(c->stream, a, c->device, c->onchip)); c->stats.launches++; }
    else FOTO_TRY(launch_gn_pcg(c->stream, a, c->gn_grid, c->gn_block));
    prof_end(c);
    c->stats.launches += 2; c->stats.gn_launches++;
    CUDA_TRY(cudaMemcpyAsync(d_u, a.x, P * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_v, a.x + P, P * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_m, a.x + 2 * P, P * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaGetLastError());
    FOTO_TRY(fetch_result(c));
    prof_resolve(c);
    c->stats.gn_iterations += c->h_res->cg_iters;
    c->stats.gn_pixels += (long long)c->h_res->cg_iters * (long long)P;
    if (iters) *iters = c->h_res->cg_iters;
    if (info) *info = c->h_res->cg_info;
    return FOTO_OK;
}

// ------------------------------------------------------------------------------- host API
static std::mutex g_ctx_mutex;
static std::map<std::pair<std::thread::id, int>, foto_ctx *> g_ctx;

static std::atomic<int> g_default_variant(-2);      // -2: not yet read from the environment

static int default_variant()
{
    int v = g_default_variant.load();
    if (v == -2) {
        const char *e = getenv("FOTO_CG_VARIANT");
        v = e ? atoi(e) : -1;
        if (v < -1 || v > 3 || v == 1) v = -1;
        g_default_variant.store(v);
    }
    return v;
}

extern "C" int foto_set_default_cg_variant(int v)
{
    if (v < -1 || v > 3 || v == 1) { set_error("cg variant must be -1 (auto), 0 (streaming), 2 (on-chip single-reduction) or 3 (GN: spectral preconditioner)"); return FOTO_ERR_ARG; }
    g_default_variant.store(v);
    return FOTO_OK;
}

static int default_ctx(foto_ctx **out)
{
    int n = foto_device_count();
    if (n <= 0) { if (n == 0) set_error("no CUDA device: libfoto_b200 has no CPU fallback"); return FOTO_ERR_NODEV; }
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    auto key = std::make_pair(std::this_thread::get_id(), dev);
    auto it = g_ctx.find(key);
    if (it != g_ctx.end()) { *out = it->second; (*out)->cg_variant = default_variant(); return FOTO_OK; }
    foto_ctx *c = nullptr;
    FOTO_TRY(foto_ctx_create(dev, &c));
    g_ctx[key] = c;
    c->cg_variant = default_variant();
    *out = c;
    return FOTO_OK;
}

struct IoPlan {             // host<->device staging of doubles through ctx->io
    foto_ctx *c; size_t off = 0;
    explicit IoPlan(foto_ctx *ctx) : c(ctx) {}
    double *slot(size_t n) { double *p = (double *)(c->io + off); off += Carver::bytes(n); return p; }
};

static int h2d(foto_ctx *c, double *dst, const double *src, size_t n)
{
    CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return FOTO_OK;
}
static int d2h(foto_ctx *c, double *dst, const double *src, size_t n)
{
    CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return FOTO_OK;
}

static int solve_on_ctx(foto_ctx *c, const double *rho0, const double *rhoT, int Nt, int Nx, int Ny, double r,
                        double tol, double eps, int max_it, int backend, double *u, double *v, double *m,
                        double *crit_trace, int *n_outer, int *cg_iters, int *cg_info)
{
    if (!rho0 || !rhoT || !u || !v || !m) { set_error("foto_solve: NULL argument"); return FOTO_ERR_ARG; }
    Dims d;
    FOTO_TRY(make_dims(Nt, Nx, Ny, &d));
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, 5 * Carver::bytes(d.P)));
    IoPlan io(c);
    double *d0 = io.slot(d.P), *dT = io.slot(d.P), *du = io.slot(d.P), *dv = io.slot(d.P), *dm = io.slot(d.P);
    FOTO_TRY(h2d(c, d0, rho0, d.P));
    FOTO_TRY(h2d(c, dT, rhoT, d.P));
    FOTO_TRY(foto_solve_dev(c, d0, dT, Nt, Nx, Ny, r, tol, eps, max_it, backend, du, dv, dm, crit_trace, n_outer,
                            cg_iters, cg_info));
    FOTO_TRY(d2h(c, u, du, d.P));
    FOTO_TRY(d2h(c, v, dv, d.P));
    FOTO_TRY(d2h(c, m, dm, d.P));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FOTO_OK;
}

extern "C" int foto_solve(const double *rho0, const double *rhoT, int Nt, int Nx, int Ny, double r, double tol,
                          double eps, int max_it, int backend, double *u, double *v, double *m, double *crit_trace,
                          int *n_outer, int *cg_iters, int *cg_info)
{
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    return solve_on_ctx(c, rho0, rhoT, Nt, Nx, Ny, r, tol, eps, max_it, backend, u, v, m, crit_trace, n_outer,
                        cg_iters, cg_info);
}

extern "C" int foto_solve_host(foto_ctx *c, const double *rho0, const double *rhoT, int Nt, int Nx, int Ny, double r,
                               double tol, double eps, int max_it, int backend, double *u, double *v, double *m,
                               double *crit_trace, int *n_outer, int *cg_iters, int *cg_info)
{
    if (!c) { set_error("foto_solve_host: NULL context"); return FOTO_ERR_ARG; }
    return solve_on_ctx(c, rho0, rhoT, Nt, Nx, Ny, r, tol, eps, max_it, backend, u, v, m, crit_trace, n_outer,
                        cg_iters, cg_info);
}

extern "C" int foto_stepB(const double *p, int Nt, int Nx, int Ny, double *q)
{
    if (!p || !q) { set_error("foto_stepB: NULL argument"); return FOTO_ERR_ARG; }
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    if (Nt < 1 || Nx < 1 || Ny < 1) { set_error("foto_stepB: empty grid"); return FOTO_ERR_ARG; }
    unsigned long long N = (unsigned long long)Nt * Nx * Ny;
    if (3ull * N >= (1ull << 32)) { set_error("grid too large"); return FOTO_ERR_ARG; }
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, 2 * Carver::bytes(3 * N)));
    IoPlan io(c);
    double *dp = io.slot(3 * N), *dq = io.slot(3 * N);
    FOTO_TRY(h2d(c, dp, p, 3 * N));
    launch_stepB(c->stream, (unsigned int)N, dp, dq);
    c->stats.launches++;
    CUDA_TRY(cudaGetLastError());
    FOTO_TRY(d2h(c, q, dq, 3 * N));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FOTO_OK;
}

extern "C" int foto_rhs(const double *mu, const double *q, const double *rho0, const double *rhoT, double r, int Nt,
                        int Nx, int Ny, double *F)
{
    if (!mu || !q || !rho0 || !rhoT || !F) { set_error("foto_rhs: NULL argument"); return FOTO_ERR_ARG; }
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    Dims d;
    FOTO_TRY(make_dims(Nt, Nx, Ny, &d));
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, 2 * Carver::bytes(3ull * d.N) + 2 * Carver::bytes(d.P) + Carver::bytes(d.N)));
    IoPlan io(c);
    double *dmu = io.slot(3ull * d.N), *dq = io.slot(3ull * d.N), *d0 = io.slot(d.P), *dT = io.slot(d.P), *dF = io.slot(d.N);
    FOTO_TRY(h2d(c, dmu, mu, 3ull * d.N)); FOTO_TRY(h2d(c, dq, q, 3ull * d.N));
    FOTO_TRY(h2d(c, d0, rho0, d.P)); FOTO_TRY(h2d(c, dT, rhoT, d.P));
    launch_rhs(c->stream, d, dmu, dq, d0, dT, r, dF);
    c->stats.launches++;
    CUDA_TRY(cudaGetLastError());
    FOTO_TRY(d2h(c, F, dF, d.N));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FOTO_OK;
}

extern "C" int foto_stepA(const double *mu, const double *q, const double *rho0, const double *rhoT, double r,
                          double eps, int Nt, int Nx, int Ny, int backend, double *phi, int *cg_iters, int *cg_info)
{
    if (!mu || !q || !rho0 || !rhoT || !phi) { set_error("foto_stepA: NULL argument"); return FOTO_ERR_ARG; }
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    Dims d;
    FOTO_TRY(make_dims(Nt, Nx, Ny, &d));
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, 2 * Carver::bytes(3ull * d.N) + 2 * Carver::bytes(d.P) + 2 * Carver::bytes(d.N)));
    FOTO_TRY(ensure(&c->ws, &c->ws_bytes, 4ull * d.N * sizeof(double) + 256));
    IoPlan io(c);
    double *dmu = io.slot(3ull * d.N), *dq = io.slot(3ull * d.N), *d0 = io.slot(d.P), *dT = io.slot(d.P);
    double *dF = io.slot(d.N), *dphi = io.slot(d.N);
    FOTO_TRY(h2d(c, dmu, mu, 3ull * d.N)); FOTO_TRY(h2d(c, dq, q, 3ull * d.N));
    FOTO_TRY(h2d(c, d0, rho0, d.P)); FOTO_TRY(h2d(c, dT, rhoT, d.P));
    launch_rhs(c->stream, d, dmu, dq, d0, dT, r, dF);
    c->stats.launches++;
    FOTO_TRY(run_cg(c, d, dF, dphi, (double *)c->ws, r, eps, backend));
    FOTO_TRY(fetch_result(c));
    if (cg_iters) *cg_iters = c->h_res->cg_iters;
    if (cg_info) *cg_info = c->h_res->cg_info;
    FOTO_TRY(d2h(c, phi, dphi, d.N));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    prof_resolve(c);
    return FOTO_OK;
}

extern "C" int foto_flow_from_phi(const double *phi, int Nt, int Nx, int Ny, double *u, double *v, double *m)
{
    if (!phi || !u || !v || !m) { set_error("foto_flow_from_phi: NULL argument"); return FOTO_ERR_ARG; }
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    Dims d;
    FOTO_TRY(make_dims(Nt, Nx, Ny, &d));
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, Carver::bytes(d.N) + 3 * Carver::bytes(d.P)));
    IoPlan io(c);
    double *dphi = io.slot(d.N), *du = io.slot(d.P), *dv = io.slot(d.P), *dm = io.slot(d.P);
    FOTO_TRY(h2d(c, dphi, phi, d.N));
    launch_flow(c->stream, d, dphi, du, dv, dm);
    c->stats.launches += 2;
    CUDA_TRY(cudaGetLastError());
    FOTO_TRY(d2h(c, u, du, d.P)); FOTO_TRY(d2h(c, v, dv, d.P)); FOTO_TRY(d2h(c, m, dm, d.P));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FOTO_OK;
}

// 1-D finite-difference builders of operators.py:5-110 as tridiagonal rows (lo, di, up).
// The reference patches boundary rows after dividing by h, so those entries stay unscaled.
extern "C" int foto_tri_coeffs(int kind, int n, double h, int bc, double *lo, double *di, double *up)
{
    if (bc != FOTO_BC_N && bc != FOTO_BC_D) { set_error("These boundary conditions are not implemented"); return FOTO_ERR_NOTIMPL; }
    if (kind < 0 || kind > FOTO_1D_LAP) { set_error("unknown 1-D builder %d", kind); return FOTO_ERR_NOTIMPL; }
    if (n < 2 || !lo || !di || !up) { set_error("1-D operator needs n >= 2"); return FOTO_ERR_ARG; }
    const bool neumann = bc == FOTO_BC_N;
    const double ih = 1.0 / h, hh = 0.5 / h, i2 = 1.0 / (h * h);
    for (int i = 0; i < n; i++) {
        const bool first = i == 0, last = i == n - 1;
        double l = 0, d = 0, u = 0;
        switch (kind) {
        case FOTO_1D_FORWARD_WEIRD:  d = -ih; u = ih; if (last) { l = -1.0; d = 1.0; } break;
        case FOTO_1D_BACKWARD_WEIRD: l = -ih; d = ih; if (first) { d = -1.0; u = 1.0; } break;
        case FOTO_1D_CENTRAL_WEIRD:
            l = -hh; u = hh;
            if (neumann && first) { d = -1.0; u = 1.0; }
            if (neumann && last) { d = 1.0; l = -1.0; }
            break;
        case FOTO_1D_CENTRAL: l = -hh; u = hh; if (neumann && (first || last)) { l = 0; u = 0; } break;
        case FOTO_1D_FORWARD: d = -ih; u = ih; if (neumann && last) d = 0; break;
        case FOTO_1D_BACKWARD: l = -ih; d = ih; if (neumann && first) d = 0; break;
        case FOTO_1D_LAP:
            l = i2; d = -2.0 * i2; u = i2;
            if (neumann && (first || last)) d = -i2;
            break;
        }
        if (first) l = 0;
        if (last) u = 0;
        lo[i] = l; di[i] = d; up[i] = u;
    }
    return FOTO_OK;
}

extern "C" int foto_op_apply(int op, int bc, int Nt, int Nx, int Ny, double dt, double dx, double dy, int transpose,
                             const double *in, double *out)
{
    if (!in || !out) { set_error("foto_op_apply: NULL argument"); return FOTO_ERR_ARG; }
    if (op < 0 || op > FOTO_OP_GRAD_FORWARD) { set_error("unknown operator %d", op); return FOTO_ERR_NOTIMPL; }
    const bool three_d = op <= FOTO_OP_LAPLACIAN_ST;
    if (!three_d) Nt = 1;
    if (Nx < 2 || Ny < 2 || (three_d && Nt < 2)) { set_error("operator grid must be at least 2 per axis"); return FOTO_ERR_ARG; }
    const int kind = (op == FOTO_OP_GRAD_ST || op == FOTO_OP_DIV_ST) ? FOTO_1D_CENTRAL_WEIRD
                   : op == FOTO_OP_LAPLACIAN_ST ? FOTO_1D_LAP
                   : op == FOTO_OP_GRAD_FORWARD ? FOTO_1D_FORWARD : FOTO_1D_CENTRAL;
    const int lens[3] = {Nx, Ny, Nt};
    const double hs[3] = {dx, dy, dt};
    int mx = Nx > Ny ? Nx : Ny; if (Nt > mx) mx = Nt;
    std::vector<double> hc(9 * (size_t)mx, 0.0);
    const int naxes = three_d ? 3 : 2;
    for (int ax = 0; ax < naxes; ax++)
        FOTO_TRY(foto_tri_coeffs(kind, lens[ax], hs[ax], bc, &hc[(3 * ax) * mx], &hc[(3 * ax + 1) * mx], &hc[(3 * ax + 2) * mx]));
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    const unsigned long long P = (unsigned long long)Nx * Ny, N = P * Nt;
    if (3ull * N >= (1ull << 32)) { set_error("grid too large"); return FOTO_ERR_ARG; }
    const int ncomp = naxes;
    const bool stack_out = (op == FOTO_OP_GRAD_ST || op == FOTO_OP_GRAD || op == FOTO_OP_GRAD_FORWARD);
    const bool single = op == FOTO_OP_LAPLACIAN_ST;
    const bool one_to_many = !single && (stack_out != (transpose != 0));
    const size_t n_in = single ? N : (one_to_many ? N : ncomp * N), n_out = single ? N : (one_to_many ? ncomp * N : N);
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, Carver::bytes(n_in) + Carver::bytes(n_out) + Carver::bytes(hc.size())));
    IoPlan io(c);
    double *din = io.slot(n_in), *dout = io.slot(n_out), *dc = io.slot(hc.size());
    FOTO_TRY(h2d(c, din, in, n_in));
    FOTO_TRY(h2d(c, dc, hc.data(), hc.size()));
    const unsigned int strides[3] = {1u, (unsigned int)Nx, (unsigned int)P};
    // block order of the reference: [t, x, y] for space-time operators, [x, y] in 2-D
    const int order3[3] = {2, 0, 1}, order2[2] = {0, 1};
    const int *order = three_d ? order3 : order2;
    for (int b = 0; b < naxes; b++) {
        const int ax = order[b];
        const double *lo = dc + (3 * ax) * mx, *di = dc + (3 * ax + 1) * mx, *up = dc + (3 * ax + 2) * mx;
        const double *src = (single || one_to_many) ? din : din + (size_t)b * N;
        double *dst = (single || !one_to_many) ? dout : dout + (size_t)b * N;
        const int accumulate = (single || !one_to_many) && b > 0;
        launch_axis_apply(c->stream, src, dst, lo, di, up, transpose, strides[ax], lens[ax], (unsigned int)N, accumulate);
        c->stats.launches++;
    }
    CUDA_TRY(cudaGetLastError());
    FOTO_TRY(d2h(c, out, dout, n_out));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FOTO_OK;
}

static int gn_on_ctx(foto_ctx *c, const double *f1, const double *f2, int w, int h, double alpha, double lambda,
                     double rtol, int max_it, double *u, double *v, double *m, int *iters, int *info)
{
    if (!f1 || !f2 || !u || !v || !m) { set_error("foto_gn_solve: NULL argument"); return FOTO_ERR_ARG; }
    if (w < 2 || h < 2) { set_error("image must be at least 2x2"); return FOTO_ERR_ARG; }
    const size_t P = (size_t)w * h;
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, 5 * Carver::bytes(P)));
    IoPlan io(c);
    double *d1 = io.slot(P), *d2 = io.slot(P), *du = io.slot(P), *dv = io.slot(P), *dm = io.slot(P);
    FOTO_TRY(h2d(c, d1, f1, P)); FOTO_TRY(h2d(c, d2, f2, P));
    FOTO_TRY(foto_gn_solve_dev(c, d1, d2, w, h, alpha, lambda, rtol, max_it, du, dv, dm, iters, info));
    FOTO_TRY(d2h(c, u, du, P)); FOTO_TRY(d2h(c, v, dv, P)); FOTO_TRY(d2h(c, m, dm, P));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FOTO_OK;
}

extern "C" int foto_gn_solve(const double *f1, const double *f2, int w, int h, double alpha, double lambda,
                             double rtol, int max_it, double *u, double *v, double *m, int *iters, int *info)
{
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    return gn_on_ctx(c, f1, f2, w, h, alpha, lambda, rtol, max_it, u, v, m, iters, info);
}

extern "C" int foto_gn_solve_host(foto_ctx *c, const double *f1, const double *f2, int w, int h, double alpha,
                                  double lambda, double rtol, int max_it, double *u, double *v, double *m, int *iters,
                                  int *info)
{
    if (!c) { set_error("foto_gn_solve_host: NULL context"); return FOTO_ERR_ARG; }
    return gn_on_ctx(c, f1, f2, w, h, alpha, lambda, rtol, max_it, u, v, m, iters, info);
}

extern "C" int foto_gn_system(const double *f1, const double *f2, int w, int h, double alpha, double lambda,
                              const double *x, double *y, double *b)
{
    if (!f1 || !f2 || !x || !y || !b) { set_error("foto_gn_system: NULL argument"); return FOTO_ERR_ARG; }
    if (w < 2 || h < 2) { set_error("image must be at least 2x2"); return FOTO_ERR_ARG; }
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    const size_t P = (size_t)w * h;
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, 4 * Carver::bytes(P) + 4 * Carver::bytes(3 * P)));
    IoPlan io(c);
    double *d1 = io.slot(P), *d2 = io.slot(P), *fx = io.slot(P), *fy = io.slot(P);
    double *dinv = io.slot(3 * P), *db = io.slot(3 * P), *dx = io.slot(3 * P), *dy = io.slot(3 * P);
    FOTO_TRY(h2d(c, d1, f1, P)); FOTO_TRY(h2d(c, d2, f2, P)); FOTO_TRY(h2d(c, dx, x, 3 * P));
    launch_gn_coeffs(c->stream, w, h, d1, d2, alpha, lambda, fx, fy, dinv, db);
    launch_gn_apply(c->stream, w, h, fx, fy, d2, alpha, lambda, dx, dy);
    c->stats.launches += 2;
    CUDA_TRY(cudaGetLastError());
    FOTO_TRY(d2h(c, y, dy, 3 * P)); FOTO_TRY(d2h(c, b, db, 3 * P));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FOTO_OK;
}

extern "C" int foto_warp_apply(const double *f1, const double *u, const double *v, int w, int h, const double *m,
                               double *out)
{
    if (!f1 || !u || !v || !out) { set_error("foto_warp_apply: NULL argument"); return FOTO_ERR_ARG; }
    if (w < 1 || h < 1) { set_error("empty image"); return FOTO_ERR_ARG; }
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    const size_t P = (size_t)w * h;
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, 6 * Carver::bytes(P)));
    IoPlan io(c);
    double *df = io.slot(P), *du = io.slot(P), *dv = io.slot(P), *dm = io.slot(P), *dg = io.slot(P), *dout = io.slot(P);
    FOTO_TRY(h2d(c, df, f1, P)); FOTO_TRY(h2d(c, du, u, P)); FOTO_TRY(h2d(c, dv, v, P));
    if (m) FOTO_TRY(h2d(c, dm, m, P));
    launch_warp(c->stream, w, h, df, du, dv, m ? dm : nullptr, dg, dout);
    c->stats.launches += 2;
    CUDA_TRY(cudaGetLastError());
    FOTO_TRY(d2h(c, out, dout, P));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FOTO_OK;
}

extern "C" int foto_pack_flo(const double *u, const double *v, int n, float *out)
{
    if (!u || !v || !out || n < 1) { set_error("foto_pack_flo: bad argument"); return FOTO_ERR_ARG; }
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, 3 * Carver::bytes(n)));
    IoPlan io(c);
    double *du = io.slot(n), *dv = io.slot(n); float *dout = (float *)io.slot(n);
    FOTO_TRY(h2d(c, du, u, n)); FOTO_TRY(h2d(c, dv, v, n));
    launch_pack_flo(c->stream, (unsigned int)n, du, dv, dout);
    c->stats.launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, dout, (size_t)n * 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FOTO_OK;
}

extern "C" int foto_flow_metrics(const double *u, const double *v, const double *ug, const double *vg, int n, double *out6)
{
    if (!u || !v || !ug || !vg || !out6 || n < 1) { set_error("foto_flow_metrics: bad argument"); return FOTO_ERR_ARG; }
    foto_ctx *c = nullptr;
    FOTO_TRY(default_ctx(&c));
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure(&c->io, &c->io_bytes, 4 * Carver::bytes(n) + Carver::bytes(6 * 148 * 8 + 8)));
    IoPlan io(c);
    double *d[4] = {io.slot(n), io.slot(n), io.slot(n), io.slot(n)}, *part = io.slot(6 * 148 * 8 + 8);
    const double *h[4] = {u, v, ug, vg};
    for (int i = 0; i < 4; i++) FOTO_TRY(h2d(c, d[i], h[i], n));
    launch_flow_metrics(c->stream, (unsigned int)n, d[0], d[1], d[2], d[3], part, part + 6 * 148 * 8);
    c->stats.launches += 2;
    CUDA_TRY(cudaGetLastError());
    FOTO_TRY(d2h(c, out6, part + 6 * 148 * 8, 6));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FOTO_OK;
}

// ------------------------------------------------------------------------------- device-resident ingest / egress
extern "C" int foto_ingest_u8_dev(foto_ctx *c, const unsigned char *d_u8, int n, double *d_out)
{
    if (!c || !d_u8 || !d_out || n < 1) { set_error("foto_ingest_u8_dev: bad argument"); return FOTO_ERR_ARG; }
    FOTO_TRY(ctx_bind(c));
    launch_ingest_u8(c->stream, (unsigned int)n, d_u8, d_out);
    c->stats.launches++;
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

extern "C" int foto_pack_flo_dev(foto_ctx *c, const double *d_u, const double *d_v, int n, float *d_out)
{
    if (!c || !d_u || !d_v || !d_out || n < 1) { set_error("foto_pack_flo_dev: bad argument"); return FOTO_ERR_ARG; }
    FOTO_TRY(ctx_bind(c));
    launch_pack_flo(c->stream, (unsigned int)n, d_u, d_v, d_out);
    c->stats.launches++;
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

static const int kMetricPartials = 6 * 148 * 8 + 8;

extern "C" int foto_flow_metrics_dev(foto_ctx *c, const double *d_u, const double *d_v, const double *d_ug, const double *d_vg,
                                     int n, double *d_out6)
{
    if (!c || !d_u || !d_v || !d_ug || !d_vg || !d_out6 || n < 1) { set_error("foto_flow_metrics_dev: bad argument"); return FOTO_ERR_ARG; }
    FOTO_TRY(ctx_bind(c));
    if (!c->metric_partials) CUDA_TRY(cudaMalloc((void **)&c->metric_partials, kMetricPartials * sizeof(double)));
    launch_flow_metrics(c->stream, (unsigned int)n, d_u, d_v, d_ug, d_vg, c->metric_partials, c->metric_partials + 6 * 148 * 8);
    CUDA_TRY(cudaMemcpyAsync(d_out6, c->metric_partials + 6 * 148 * 8, 6 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->stats.launches += 2;
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

extern "C" int foto_warp_dev(foto_ctx *c, const double *d_f1, const double *d_u, const double *d_v, int w, int h,
                             const double *d_m, double *d_out, const double *d_igt, double *d_ie)
{
    if (!c || !d_f1 || !d_u || !d_v || !d_out || w < 1 || h < 1) { set_error("foto_warp_dev: bad argument"); return FOTO_ERR_ARG; }
    FOTO_TRY(ctx_bind(c));
    const size_t P = (size_t)w * h;
    FOTO_TRY(ensure(&c->warp_tmp, &c->warp_tmp_bytes, Carver::bytes(P)));
    launch_warp(c->stream, w, h, d_f1, d_u, d_v, d_m, (double *)c->warp_tmp, d_out);
    c->stats.launches += 2;
    if (d_igt && d_ie) {
        if (!c->metric_partials) CUDA_TRY(cudaMalloc((void **)&c->metric_partials, kMetricPartials * sizeof(double)));
        launch_ie_sumsq(c->stream, (unsigned int)P, d_out, d_igt, c->metric_partials, d_ie);
        c->stats.launches += 2;
    }
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

// ------------------------------------------------------------------------------- time-slab building blocks
// One huge volume split into contiguous time slabs, one per rank (second row).  The
// exchange steps (1-plane halos, the t <-> y all-to-all of the DCT, the 2-scalar all-reduce of the criterion)
// are NCCL calls made by the host driver (foto_b200/slab.py, torch.distributed); these entry points are the
// compute between them.  Arrays are device pointers to the first OWNED plane; 3-component fields have
// component stride cs (in doubles) and addressable halo planes at -1 and nloc where those exist globally.
extern "C" int foto_ctx_set_stream(foto_ctx *c, void *stream, int use_own_stream)
{
    if (!c) return FOTO_ERR_ARG;
    FOTO_TRY(ctx_bind(c));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->stream = use_own_stream ? c->own_stream : (cudaStream_t)stream;     // NULL = the legacy default stream
    return FOTO_OK;
}

static int slab_dims(int gNt, int n0, int nloc, int Nx, int Ny, unsigned long long cs, Dims *d)
{
    if (gNt < 2 || nloc < 1 || n0 < 0 || n0 + nloc > gNt || Nx < 2 || Ny < 2) { set_error("bad slab geometry"); return FOTO_ERR_ARG; }
    const unsigned long long P = (unsigned long long)Nx * Ny;
    if (cs < P * nloc || 3ull * cs >= (1ull << 32)) { set_error("bad slab component stride"); return FOTO_ERR_ARG; }
    d->Nt = nloc; d->Ny = Ny; d->Nx = Nx; d->P = (unsigned int)P; d->N = (unsigned int)(P * nloc);
    d->n0 = n0; d->gNt = gNt; d->cs = (unsigned int)cs;
    return FOTO_OK;
}

extern "C" int foto_slab_rhs_dev(foto_ctx *c, const double *mu, const double *q, unsigned long long cs, const double *rho0,
                                 const double *rhoT, double r, int gNt, int n0, int nloc, int Nx, int Ny, double *F)
{
    if (!c || !mu || !q || !rho0 || !rhoT || !F) { set_error("foto_slab_rhs_dev: NULL argument"); return FOTO_ERR_ARG; }
    Dims d;
    FOTO_TRY(slab_dims(gNt, n0, nloc, Nx, Ny, cs, &d));
    FOTO_TRY(ctx_bind(c));
    launch_rhs(c->stream, d, mu, q, rho0, rhoT, r, F);
    c->stats.launches++;
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

extern "C" int foto_slab_prox_dev(foto_ctx *c, const double *phi, double *mu, double *q, unsigned long long cs, double r,
                                  int gNt, int n0, int nloc, int Nx, int Ny, double *d_out2)
{
    if (!c || !phi || !mu || !q || !d_out2) { set_error("foto_slab_prox_dev: NULL argument"); return FOTO_ERR_ARG; }
    Dims d;
    FOTO_TRY(slab_dims(gNt, n0, nloc, Nx, Ny, cs, &d));
    FOTO_TRY(ctx_bind(c));
    const int blocks = launch_prox_dual(c->stream, d, phi, mu, q, r, c->prox_partials, kProxMaxBlocks, c->num_sms, &c->stats.prox_variant);
    if (blocks < 0) return FOTO_ERR_CUDA;
    launch_crit_final(c->stream, c->prox_partials, blocks, d_out2);
    c->stats.launches += 2;
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

// One step of the reference's truncated CG on a time slab (cg_slab.cu); the caller issues the collectives in between.
extern "C" int foto_slab_cg_dev(foto_ctx *c, int op, int gNt, int n0, int nloc, int Ny, int Nx, double r, double eps, double rtol, int it,
                                int maxiter, const double *d_b, double *d_x, double *d_r, double *d_p_old, double *d_p_new, double *d_q,
                                double *d_state)
{
    if (!c || !d_state || gNt < 2 || nloc < 1 || n0 < 0 || n0 + nloc > gNt || Nx < 2 || Ny < 2) { set_error("foto_slab_cg_dev: bad argument"); return FOTO_ERR_ARG; }
    if ((unsigned long long)Nx * Ny * ((unsigned long long)nloc + 2) >= (1ull << 32)) { set_error("foto_slab_cg_dev: slab too large"); return FOTO_ERR_ARG; }
    if ((op == 0 && (!d_b || !d_x || !d_r || !d_p_old)) || (op == 3 && (!d_r || !d_p_old || !d_p_new || !d_q)) ||
        (op == 5 && (!d_x || !d_r || !d_p_new || !d_q))) { set_error("foto_slab_cg_dev: NULL argument"); return FOTO_ERR_ARG; }
    static_assert(2 * kMaxVals * kMaxBlocks >= 148 * 8, "partials buffer");
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(launch_cg_slab(c->stream, op, gNt, n0, nloc, Ny, Nx, r, eps, rtol, it, maxiter, d_b, d_x, d_r, d_p_old, d_p_new, d_q,
                            c->sync_partials, d_state));
    c->stats.launches++;
    return FOTO_OK;
}

extern "C" int foto_slab_cg_state_words(void) { return (int)cg_slab_state_words(); }

extern "C" int foto_dct_xy_dev(foto_ctx *c, const double *in, double *out, double *tmp, int nplanes, int gNt, int Ny, int Nx,
                               int inverse)
{
    if (!c || !in || !out || !tmp || nplanes < 1) { set_error("foto_dct_xy_dev: bad argument"); return FOTO_ERR_ARG; }
    Dims d;
    FOTO_TRY(make_dims(gNt, Nx, Ny, &d));
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure_dct_tables(c, d));
    FOTO_TRY(launch_dct_xy(c->stream, c->dct, nplanes, Ny, Nx, in, out, tmp, inverse));
    c->stats.launches += c->dct.split ? 4 : 2;
    return FOTO_OK;
}

extern "C" int foto_dct_t_solve_dev(foto_ctx *c, const double *in, double *out, int gNt, int Ny, int Nx, int y_off, int ny_loc,
                                    double r, double eps)
{
    if (!c || !in || !out || y_off < 0 || ny_loc < 1 || y_off + ny_loc > Ny) { set_error("foto_dct_t_solve_dev: bad argument"); return FOTO_ERR_ARG; }
    Dims d;
    FOTO_TRY(make_dims(gNt, Nx, Ny, &d));
    FOTO_TRY(ctx_bind(c));
    FOTO_TRY(ensure_dct_tables(c, d));
    FOTO_TRY(launch_dct_t_solve(c->stream, c->dct, gNt, ny_loc, Nx, y_off, r, eps, in, out));
    c->stats.launches++;
    return FOTO_OK;
}

// t-slab <-> y-slab transpose buffers of the DCT all-to-all: direction 0 packs nloc planes [nloc][Ny][Nx] into the send
// buffer (block for rank g contiguous: [nloc][rows of g][Nx]), direction 1 unpacks a received buffer into planes
extern "C" int foto_slab_pack_dev(foto_ctx *c, int direction, int nloc, int Ny, int Nx, int world, const double *d_in, double *d_out)
{
    if (!c || !d_in || !d_out || nloc < 1 || Ny < world || world < 1 || Nx < 1) { set_error("foto_slab_pack_dev: bad argument"); return FOTO_ERR_ARG; }
    FOTO_TRY(ctx_bind(c));
    if (direction == 0) launch_slab_pack(c->stream, nloc, Ny, Nx, world, d_in, d_out, nullptr, nullptr);
    else launch_slab_pack(c->stream, nloc, Ny, Nx, world, nullptr, nullptr, d_out, d_in);
    c->stats.launches++;
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

extern "C" int foto_flow_dev(foto_ctx *c, const double *d_phi, int Nt, int Nx, int Ny, double *d_u, double *d_v, double *d_m)
{
    if (!c || !d_phi || !d_u || !d_v || !d_m) { set_error("foto_flow_dev: NULL argument"); return FOTO_ERR_ARG; }
    Dims d;
    FOTO_TRY(make_dims(Nt, Nx, Ny, &d));
    FOTO_TRY(ctx_bind(c));
    launch_flow(c->stream, d, d_phi, d_u, d_v, d_m);
    c->stats.launches += 2;
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

// ------------------------------------------------------------------------------- batch driver
// Pairs are independent: one host thread per device pulls pair indices
// from an atomic counter (dynamic load balance: iteration counts are data dependent), runs the
// whole solve on its device and writes straight into the caller's output rows.  No collective.
template <class Fn>
static int run_batch(int n_items, const int *device_ids, int n_dev, Fn per_item)
{
    if (n_items < 0 || n_dev < 1) { set_error("batch: need n_pairs >= 0 and n_dev >= 1"); return FOTO_ERR_ARG; }
    int have = foto_device_count();
    if (have <= 0) { if (have == 0) set_error("no CUDA device: libfoto_b200 has no CPU fallback"); return FOTO_ERR_NODEV; }
    std::atomic<int> next(0), first_rc(FOTO_OK);
    std::mutex err_mutex;
    std::string err_msg;
    std::vector<std::thread> threads;
    for (int t = 0; t < n_dev; t++) {
        const int dev = device_ids ? device_ids[t] : t;
        threads.emplace_back([&, dev]() {
            foto_ctx *c = nullptr;
            int rc = foto_ctx_create(dev, &c);
            if (rc == FOTO_OK) c->cg_variant = default_variant();
            while (rc == FOTO_OK) {
                const int i = next.fetch_add(1);
                if (i >= n_items || first_rc.load() != FOTO_OK) break;
                rc = per_item(c, i);
            }
            if (rc != FOTO_OK) {
                std::lock_guard<std::mutex> lk(err_mutex);
                if (first_rc.load() == FOTO_OK) { first_rc.store(rc); err_msg = foto_last_error(); }
            }
            foto_ctx_destroy(c);
        });
    }
    for (auto &th : threads) th.join();
    if (first_rc.load() != FOTO_OK) set_error("%s", err_msg.c_str());
    return first_rc.load();
}

extern "C" int foto_solve_batch(int n_pairs, const double *rho0s, const double *rhoTs, int Nt, int Nx, int Ny,
                                double r, double tol, double eps, int max_it, int backend, const int *device_ids,
                                int n_dev, double *us, double *vs, double *ms, int *n_outer)
{
    if (!rho0s || !rhoTs || !us || !vs || !ms) { set_error("foto_solve_batch: NULL argument"); return FOTO_ERR_ARG; }
    const size_t P = (size_t)Nx * Ny;
    return run_batch(n_pairs, device_ids, n_dev, [&](foto_ctx *c, int i) {
        int outer = 0;
        int rc = solve_on_ctx(c, rho0s + i * P, rhoTs + i * P, Nt, Nx, Ny, r, tol, eps, max_it, backend, us + i * P,
                              vs + i * P, ms + i * P, nullptr, &outer, nullptr, nullptr);
        if (n_outer) n_outer[i] = outer;
        return rc;
    });
}

extern "C" int foto_gn_solve_batch(int n_pairs, const double *f1s, const double *f2s, int w, int h, double alpha,
                                   double lambda, double rtol, int max_it, const int *device_ids, int n_dev,
                                   double *us, double *vs, double *ms, int *iters)
{
    if (!f1s || !f2s || !us || !vs || !ms) { set_error("foto_gn_solve_batch: NULL argument"); return FOTO_ERR_ARG; }
    const size_t P = (size_t)w * h;
    return run_batch(n_pairs, device_ids, n_dev, [&](foto_ctx *c, int i) {
        int it = 0, info = 0;
        int rc = gn_on_ctx(c, f1s + i * P, f2s + i * P, w, h, alpha, lambda, rtol, max_it, us + i * P, vs + i * P,
                           ms + i * P, &it, &info);
        if (iters) iters[i] = it;
        return rc;
    });
}

// Batched ingest / egress: 8-bit frames in, .flo payload out, pinned staging per context.
extern "C" int foto_solve_batch_u8(int n_pairs, const unsigned char *f0s, const unsigned char *f1s, int Nt, int Nx, int Ny,
                                   double r, double tol, double eps, int max_it, int backend, const int *device_ids,
                                   int n_dev, float *flo, double *ms, int *n_outer)
{
    if (!f0s || !f1s || !flo) { set_error("foto_solve_batch_u8: NULL argument"); return FOTO_ERR_ARG; }
    Dims d;
    FOTO_TRY(make_dims(Nt, Nx, Ny, &d));
    const size_t P = d.P;
    return run_batch(n_pairs, device_ids, n_dev, [&](foto_ctx *c, int i) -> int {
        FOTO_TRY(ctx_bind(c));
        // device: [u8 f0 | u8 f1] [rho0] [rhoT] [u] [v] [m] [flo payload]; pinned: [u8 f0 | u8 f1] [flo payload] [m]
        const size_t u8b = (2 * P + 255) & ~size_t(255);
        FOTO_TRY(ensure(&c->io, &c->io_bytes, u8b + 6 * Carver::bytes(P)));
        const size_t pin_need = u8b + Carver::bytes(P) + Carver::bytes(P);
        if (c->pin_bytes < pin_need) {
            if (c->pin) { CUDA_TRY(cudaFreeHost(c->pin)); c->pin = nullptr; c->pin_bytes = 0; }
            CUDA_TRY(cudaMallocHost((void **)&c->pin, pin_need));
            c->pin_bytes = pin_need;
        }
        unsigned char *d8 = (unsigned char *)c->io;
        double *base = (double *)(c->io + u8b);
        const size_t st = Carver::bytes(P) / sizeof(double);
        double *d0 = base, *dT = base + st, *du = base + 2 * st, *dv = base + 3 * st, *dm = base + 4 * st;
        float *dflo = (float *)(base + 5 * st);
        unsigned char *p8 = (unsigned char *)c->pin;
        float *pflo = (float *)(c->pin + u8b);
        double *pm = (double *)(c->pin + u8b + Carver::bytes(P));
        memcpy(p8, f0s + (size_t)i * P, P); memcpy(p8 + P, f1s + (size_t)i * P, P);
        CUDA_TRY(cudaMemcpyAsync(d8, p8, 2 * P, cudaMemcpyHostToDevice, c->stream));
        launch_ingest_u8(c->stream, (unsigned int)P, d8, d0);
        launch_ingest_u8(c->stream, (unsigned int)P, d8 + P, dT);
        c->stats.launches += 2;
        int outer = 0;
        FOTO_TRY(foto_solve_dev(c, d0, dT, Nt, Nx, Ny, r, tol, eps, max_it, backend, du, dv, dm, nullptr, &outer, nullptr, nullptr));
        launch_pack_flo(c->stream, (unsigned int)P, du, dv, dflo);
        c->stats.launches++;
        CUDA_TRY(cudaMemcpyAsync(pflo, dflo, 2 * P * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (ms) CUDA_TRY(cudaMemcpyAsync(pm, dm, P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        memcpy(flo + (size_t)i * 2 * P, pflo, 2 * P * sizeof(float));
        if (ms) memcpy(ms + (size_t)i * P, pm, P * sizeof(double));
        if (n_outer) n_outer[i] = outer;
        return FOTO_OK;
    });
}
