// gn_onchip.cu -- K6, on-chip resident variant of the Gennert-Negahdaripour Jacobi-PCG solve.
//
// Same idea as cg_onchip.cu (see there for the measurements behind the design): at 388x584 the PCG
// vectors fit in L2, the streaming kernel (gn_kernels.cu) is L2-bandwidth bound (15.7 us / iteration,
// ~1 280 iterations per solve), and the whole state of 1.5 k pixels x 3 unknowns x (x, r, p, q, D^-1, g)
// fits in one SM.  One CTA per (y, x) tile; a thread owns up to PPT pixels with all three unknowns
// (u, v, m), so the pointwise coupling g (g . p) is evaluated once per pixel.  r and q live in
// registers, p (halo ring, zero outside the domain), x, D^-1 and g in shared memory.  Per iteration only the
// tile-edge values of z = D^-1 r cross L2 (the neighbours advance their halo copy of p = z + beta p with
// the owner's arithmetic) plus the slots of the two grid all-reduces (root gather, as in cg_onchip.cu).
//
// System (classical.py:102-110): A = diag(alpha, alpha, lambda) (x) (-Lap_Neumann) + g g^T, g = (fx, fy, -f2).
// The reference factorises A (SuperLU); any solve converged far below 1e-9 is a valid stand-in.
#include "foto_kernels.cuh"
#include "grid_sync.cuh"

namespace foto {

namespace {

struct Geom {
    int gy, gx, maxlen;
    double *edges;               // [ncta][4 (N,S,W,E)][3 * maxlen] tile-edge values of z
    unsigned long long *slots;
};

using namespace gsync;

// pixel descriptor: bits 0-15 index of the pixel in one plane of ps, 16-18 neighbour count, 19-22 export N/S/W/E, 23 valid
constexpr int kSiMask = 0xFFFF, kCntShift = 16;
constexpr int kEdgeN = 1 << 19, kEdgeS = 1 << 20, kEdgeW = 1 << 21, kEdgeE = 1 << 22, kEdgeAny = 0xF << 19, kValid = 1 << 23;
constexpr int kHaloPerThread = 4;

template <int NTHREADS, int PPT>
__global__ void __launch_bounds__(NTHREADS, 1) gn_onchip_kernel(GnArgs a, Geom g)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
    const int w = a.w, h = a.h;
    const int by = cta / g.gx, bx = cta - by * g.gx;
    const int y0 = (int)((long long)by * h / g.gy), y1 = (int)((long long)(by + 1) * h / g.gy);
    const int x0 = (int)((long long)bx * w / g.gx), x1 = (int)((long long)(bx + 1) * w / g.gx);
    const int ty = y1 - y0, tx = x1 - x0, PX = tx + 2, PY = ty + 2, plane = PY * PX;
    const int psz = (3 * plane + 1) & ~1;
    double *ps = smem;                                  // [3][PY][PX]   p with halo ring
    double *xs = ps + psz;                              // [3][PPT][NTHREADS]  x
    double *gs = xs + 3 * PPT * NTHREADS;               // [3][PPT][NTHREADS]  fx, fy, f2 of the owned pixels
    double *ds = gs + 3 * PPT * NTHREADS;               // [3][PPT][NTHREADS]  Jacobi D^-1 of the owned pixels
    double *red = ds + 3 * PPT * NTHREADS;              // 64 + 4
    int *hsrc = (int *)(red + 68);
    int *hdst = hsrc + 6 * (tx + ty);
    const bool hasN = by > 0, hasS = by < g.gy - 1, hasW = bx > 0, hasE = bx < g.gx - 1;
    const int edge_stride = 3 * g.maxlen;
    double *my_edges = g.edges + (size_t)cta * 4 * edge_stride;
    const int lx = tid % tx, r0 = tid / tx, RPP = NTHREADS / tx;
    const size_t P = (size_t)w * h;
    const int sdead = PX + 1;                           // (ly = 0, lx = 0) of a tile is a real pixel: use masking instead
    (void)sdead;

    double rj[PPT][3], qj[PPT][3];
    int info[PPT];

    for (int i = tid; i < psz; i += NTHREADS) ps[i] = 0.0;
    int nhalo = 0;
    {   // halo table: for each neighbour tile and component, edge value index -> ps index
        const int segNS = 3 * tx, segWE = 3 * ty;
        for (int e = tid; e < segNS; e += NTHREADS) {
            const int c = e / tx, pos = e - c * tx;
            if (hasN) { hsrc[nhalo + e] = ((cta - g.gx) * 4 + 1) * edge_stride + c * g.maxlen + pos; hdst[nhalo + e] = c * plane + pos + 1; }
        }
        if (hasN) nhalo += segNS;
        for (int e = tid; e < segNS; e += NTHREADS) {
            const int c = e / tx, pos = e - c * tx;
            if (hasS) { hsrc[nhalo + e] = ((cta + g.gx) * 4 + 0) * edge_stride + c * g.maxlen + pos; hdst[nhalo + e] = c * plane + (ty + 1) * PX + pos + 1; }
        }
        if (hasS) nhalo += segNS;
        for (int e = tid; e < segWE; e += NTHREADS) {
            const int c = e / ty, pos = e - c * ty;
            if (hasW) { hsrc[nhalo + e] = ((cta - 1) * 4 + 3) * edge_stride + c * g.maxlen + pos; hdst[nhalo + e] = c * plane + (pos + 1) * PX; }
        }
        if (hasW) nhalo += segWE;
        for (int e = tid; e < segWE; e += NTHREADS) {
            const int c = e / ty, pos = e - c * ty;
            if (hasE) { hsrc[nhalo + e] = ((cta + 1) * 4 + 2) * edge_stride + c * g.maxlen + pos; hdst[nhalo + e] = c * plane + (pos + 1) * PX + tx + 1; }
        }
        if (hasE) nhalo += segWE;
    }
    // export helper: z of component c of a pixel at (ly, lx)
    auto export_z = [&](int inf, int ly, int c, double z) {
        if (inf & kEdgeN) my_edges[0 * edge_stride + c * g.maxlen + lx] = z;
        if (inf & kEdgeS) my_edges[1 * edge_stride + c * g.maxlen + lx] = z;
        if (inf & kEdgeW) my_edges[2 * edge_stride + c * g.maxlen + ly] = z;
        if (inf & kEdgeE) my_edges[3 * edge_stride + c * g.maxlen + ly] = z;
    };
    double acc[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < PPT; j++) {
        const int ly = j * RPP + r0;
        int inf = 0;
#pragma unroll
        for (int c = 0; c < 3; c++) { rj[j][c] = 0.0; qj[j][c] = 0.0; ds[(c * PPT + j) * NTHREADS + tid] = 0.0; xs[(c * PPT + j) * NTHREADS + tid] = 0.0; gs[(c * PPT + j) * NTHREADS + tid] = 0.0; }
        if (r0 < RPP && ly < ty) {
            const int gy_ = y0 + ly, gx_ = x0 + lx;
            const size_t gk = (size_t)gy_ * w + gx_;
            const int cnt = (gy_ > 0) + (gy_ < h - 1) + (gx_ > 0) + (gx_ < w - 1);
            inf = ((ly + 1) * PX + lx + 1) | (cnt << kCntShift) | kValid;
            if (ly == 0 && hasN) inf |= kEdgeN;
            if (ly == ty - 1 && hasS) inf |= kEdgeS;
            if (lx == 0 && hasW) inf |= kEdgeW;
            if (lx == tx - 1 && hasE) inf |= kEdgeE;
            gs[(0 * PPT + j) * NTHREADS + tid] = a.fx[gk];
            gs[(1 * PPT + j) * NTHREADS + tid] = a.fy[gk];
            gs[(2 * PPT + j) * NTHREADS + tid] = a.f2[gk];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double bk = a.b[c * P + gk], dk = a.dinv[c * P + gk], zk = dk * bk;
                rj[j][c] = bk; ds[(c * PPT + j) * NTHREADS + tid] = dk;
                acc[0] = fma(bk, bk, acc[0]); acc[1] = fma(bk, zk, acc[1]);
                export_z(inf, ly, c, zk);
            }
        }
        info[j] = inf;
    }
    unsigned int gen = 0;
    bool abort = false;
    // grid-wide sums of two values + barrier, split in two halves so that work that does not feed the sums can
    // run between them; results in every thread
    auto sum_begin = [&](double *v, bool publish) {
        block_sum<2>(reinterpret_cast<double(&)[2]>(*v), red);
        if (tid == 0) grid_arrive<2>(g.slots, gen, v, publish);
    };
    auto sum_end = [&](double *v) {
        if (cta == 0 && tid < 32) grid_root<2>(g.slots, gen, ncta, tid);
        if (tid < 2) {
            const unsigned long long bits = grid_wait(g.slots, gen, tid);
            red[64 + tid] = __longlong_as_double((long long)bits);
            if (bits == kAbort) red[66] = 1.0;
        }
        __syncthreads();
        gen++;
        v[0] = red[64]; v[1] = red[65]; abort = red[66] != 0.0;
    };
    auto grid_sum2 = [&](double *v, bool publish) { sum_begin(v, publish); sum_end(v); };
    if (tid == 0) red[66] = 0.0;
    grid_sum2(acc, true);
    const double bb = acc[0];
    int it = 0, status = bb == 0.0 ? 0 : a.maxiter;
    if (!abort && bb != 0.0) {
        const double stop = a.rtol * sqrt(bb);
        double rr = bb, rz = acc[1], rz_prev = 0.0;
        for (; it < a.maxiter; it++) {
            const double beta = it > 0 ? rz / rz_prev : 0.0;
            double hv[kHaloPerThread];
#pragma unroll
            for (int e = 0; e < kHaloPerThread; e++) {
                const int hh = tid + e * NTHREADS;
                hv[e] = hh < nhalo ? __ldcg(g.edges + hsrc[hh]) : 0.0;
            }
            if (sqrt(rr) <= stop) { status = 0; break; }
            // own pixels: p = z + beta p, z = D^-1 r
#pragma unroll
            for (int j = 0; j < PPT; j++) {
                if (info[j] & kValid) {
                    const int si = info[j] & kSiMask;
#pragma unroll
                    for (int c = 0; c < 3; c++) ps[c * plane + si] = ds[(c * PPT + j) * NTHREADS + tid] * rj[j][c] + beta * ps[c * plane + si];
                }
            }
#pragma unroll
            for (int e = 0; e < kHaloPerThread; e++) {
                const int hh = tid + e * NTHREADS;
                if (hh < nhalo) { const int i = hdst[hh]; ps[i] = hv[e] + beta * ps[i]; }
            }
            __syncthreads();
            // q = A p, partial p.q
            acc[0] = 0.0; acc[1] = 0.0;
#pragma unroll
            for (int j = 0; j < PPT; j++) {
                const int inf = info[j];
                if (inf & kValid) {
                    const int si = inf & kSiMask;
                    const double cnt = (double)((inf >> kCntShift) & 7);
                    const double fx = gs[(0 * PPT + j) * NTHREADS + tid], fy = gs[(1 * PPT + j) * NTHREADS + tid], f2 = gs[(2 * PPT + j) * NTHREADS + tid];
                    double pc[3], nl[3];
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const double *pp = ps + c * plane + si;
                        pc[c] = pp[0];
                        nl[c] = cnt * pc[c] - (((pp[-PX] + pp[-1]) + pp[1]) + pp[PX]);
                    }
                    const double gp = fx * pc[0] + fy * pc[1] - f2 * pc[2];
                    qj[j][0] = a.alpha * nl[0] + fx * gp;
                    qj[j][1] = a.alpha * nl[1] + fy * gp;
                    qj[j][2] = a.lam * nl[2] - f2 * gp;
                    acc[0] += pc[0] * qj[j][0] + pc[1] * qj[j][1] + pc[2] * qj[j][2];
                }
            }
            grid_sum2(acc, false);
            if (abort) break;
            const double alpha = rz / acc[0];
            // r -= alpha q, z = D^-1 r (exported at tile edges), partial r.r and r.z; x += alpha p overlaps the barrier
            acc[0] = 0.0; acc[1] = 0.0;
#pragma unroll
            for (int j = 0; j < PPT; j++) {
                const int inf = info[j];
                if (inf & kValid) {
                    const int ly = j * RPP + r0;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const double rk = rj[j][c] - alpha * qj[j][c];
                        const double zk = ds[(c * PPT + j) * NTHREADS + tid] * rk;
                        rj[j][c] = rk;
                        acc[0] = fma(rk, rk, acc[0]); acc[1] = fma(rk, zk, acc[1]);
                        if (inf & kEdgeAny) export_z(inf, ly, c, zk);
                    }
                }
            }
            sum_begin(acc, true);
#pragma unroll
            for (int j = 0; j < PPT; j++) {
                if (info[j] & kValid) {
                    const int si = info[j] & kSiMask;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const int xi = (c * PPT + j) * NTHREADS + tid;
                        xs[xi] = xs[xi] + alpha * ps[c * plane + si];
                    }
                }
            }
            sum_end(acc);
            if (abort) break;
            rz_prev = rz; rr = acc[0]; rz = acc[1];
        }
    }
    if (abort) { if (tid == 0) *a.sync.error = 1; return; }
#pragma unroll
    for (int j = 0; j < PPT; j++) {
        if (info[j] & kValid) {
            const int ly = j * RPP + r0;
            const size_t gk = (size_t)(y0 + ly) * w + (x0 + lx);
#pragma unroll
            for (int c = 0; c < 3; c++) a.x[c * P + gk] = xs[(c * PPT + j) * NTHREADS + tid];
        }
    }
    if (cta == 0 && tid == 0) { a.out[0] = it; a.out[1] = status; }
}

__global__ void k_fill_u64(unsigned long long *p, int n, unsigned long long v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

constexpr int kThreads = 512, kPPT = 4;

struct Plan { bool ok = false; int gy = 0, gx = 0, maxlen = 0, ncta = 0; size_t smem = 0; };

Plan make_plan(OnchipScratch &d, int device, int h, int w)
{
    Plan best;
    if (!d.num_sms) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return best;
        d.num_sms = prop.multiProcessorCount; d.smem_optin = prop.sharedMemPerBlockOptin;
    }
    long long best_key = -1;
    for (int gy = 1; gy <= d.num_sms && gy <= h; gy++)
        for (int gx = 1; gx <= d.num_sms / gy && gx <= w; gx++) {
            const int ty = (h + gy - 1) / gy, tx = (w + gx - 1) / gx, ty_min = h / gy, tx_min = w / gx;
            if (tx > kThreads || ty_min < 1 || tx_min < 1) continue;
            int passes = 0;
            for (int txx = tx_min; txx <= tx; txx++) { const int rpp = kThreads / txx; const int p = (ty + rpp - 1) / rpp; if (p > passes) passes = p; }
            if (passes > kPPT) continue;
            if ((long long)(ty + 2) * (tx + 2) > kSiMask) continue;
            if (6LL * (tx + ty) > (long long)kHaloPerThread * kThreads) continue;
            const size_t smem = ((((size_t)3 * (ty + 2) * (tx + 2) + 1) & ~size_t(1)) + (size_t)9 * kPPT * kThreads + 68) * 8 + (size_t)12 * (tx + ty) * sizeof(int);
            if (smem > d.smem_optin) continue;
            const long long key = ((long long)passes * 10000 + 32000 / tx + 2000 / ty) * 100000 + (tx + ty);
            if (best_key < 0 || key < best_key) {
                best_key = key; best.ok = true; best.gy = gy; best.gx = gx; best.ncta = gy * gx;
                best.maxlen = tx > ty ? tx : ty; best.smem = smem;
            }
        }
    return best;
}

}  // namespace

bool gn_onchip_fits(OnchipScratch &s, int device, int h, int w) { return make_plan(s, device, h, w).ok; }

int launch_gn_onchip(cudaStream_t st, const GnArgs &a, int device, OnchipScratch &d)
{
    Plan p = make_plan(d, device, a.h, a.w);
    if (!p.ok) { set_error("image %dx%d does not fit the on-chip GN variant", a.h, a.w); return FOTO_ERR_ARG; }
    const size_t need = (size_t)p.ncta * 4 * 3 * p.maxlen * sizeof(double);
    if (d.gn_edges_bytes < need) {
        if (d.gn_edges) CUDA_TRY(cudaFree(d.gn_edges));
        CUDA_TRY(cudaMalloc((void **)&d.gn_edges, need));
        d.gn_edges_bytes = need;
    }
    if (!d.gn_slots) CUDA_TRY(cudaMalloc((void **)&d.gn_slots, kSlotWords * sizeof(unsigned long long)));
    if (!d.gn_attr_set) {
        CUDA_TRY(cudaFuncSetAttribute((const void *)gn_onchip_kernel<kThreads, kPPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_optin));
        d.gn_attr_set = true;
    }
    k_fill_u64<<<(kSlotWords + 255) / 256, 256, 0, st>>>(d.gn_slots, kSlotWords, kSentinel);
    Geom g;
    g.gy = p.gy; g.gx = p.gx; g.maxlen = p.maxlen; g.edges = d.gn_edges; g.slots = d.gn_slots;
    void *args[] = {(void *)&a, (void *)&g};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void *)gn_onchip_kernel<kThreads, kPPT>, dim3(p.ncta), dim3(kThreads), args, p.smem, st));
    return FOTO_OK;
}

}  // namespace foto
