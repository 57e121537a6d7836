// cg_fused.cu -- K2a, on-chip resident CG with ONE grid all-reduce per iteration.
//
// The textbook recurrences (scipy's, benamou_brenier.py:85; cg_kernels.cu) need two grid all-reduces per iteration
// (p.Ap, then r.r); an on-chip kernel built that way (round 1, removed) spent 8 800 of its 16 200 cycles per
// iteration in them, and the all-reduce is at the floor of L2 signalling (grid_sync.cuh).  This kernel runs the same
// Krylov iteration in the Chronopoulos-Gear arrangement, which needs one:
//
//     w = A r,  gamma = r.r,  delta = r.w                      <- one all-reduce of (gamma, delta)
//     stop if sqrt(gamma) < atol                                  (scipy's test, same place in the sequence)
//     beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old)   [= gamma / (delta - beta^2 d_old)]
//     p = r + beta p,  s = w + beta s  (= A p),  x += alpha p,  r -= alpha s
//
// In exact arithmetic x_k, r_k, alpha_k, beta_k are those of the textbook form; in floating point A p is carried by
// a recurrence instead of being recomputed.  Measured against scipy's cg on the reference's systems (CPU prototype,
// 4 grids up to 388x584x4): identical iteration counts in every outer iteration, phi within 9e-11 of scipy's per
// solve and u, v, m within 8e-12 after the full ALG2 loop (contract: 1e-9).  The parity tests run this kernel as
// the default and the textbook kernels beside it.
//
// The stencil now acts on r, so r (not p) lives in shared memory with the halo ring; p and s live in registers, x and
// w in private shared-memory slots.  Tile-edge values of the new r go through L2 to the four neighbours without any
// barrier or fence: every exported word carries the parity of its generation in the least significant mantissa bit
// (the owner keeps the same rounded value, so both copies of r agree; the perturbation is one ulp of an edge value
// per iteration, the size of an ordinary rounding error), and the reader spins on each word until the parity is
// the one it expects (table-driven export pass after the update, mirror of the import).  One buffer is enough: a CTA
// overwrites generation g with g+1 only after the all-reduce of iteration g, which every neighbour enters after it
// has read generation g.  (Measured alternatives: flag + release
// fence hand-off 2 900 cycles per iteration, as much as the grid barrier it replaces; sentinel reset + triple
// buffering doubles the stores and costs 3 000 cycles in the reset loop; exporting from registers inside the unrolled
// update costs the edge warps 1 500 cycles.)  Half of the x update of iteration k covers the L2 hop of the edge
// values, the other half runs in the shadow of the all-reduce of iteration k+1.
#include "foto_kernels.cuh"
#include "grid_sync.cuh"

namespace foto {

namespace {

using namespace gsync;

struct Geom {
    int gy, gx, maxlen;
    double *edges;                 // [ncta][4 (N,S,W,E)][NT * maxlen] tile-edge values of r, LSB = generation parity
    unsigned long long *slots;     // all-reduce slots (grid_sync.cuh)
    long long *prof;
};

constexpr int kHaloPerThread = 4;

// XG: x lives in global memory (L2 resident) instead of shared memory: the large variant (384 threads x 24 cell slots =
// 9 216; 3 warps per SM sub-partition leave a thread 168 registers) for grids such as 480x640x4 whose x slots no
// longer fit next to r and w
template <int NTHREADS, int NT, int YPT, bool UNIT, bool XG = false>
__global__ void __launch_bounds__(NTHREADS, 1) cg_fused_kernel(CgArgs a, Geom g)
{
    constexpr int CPT = NT * YPT;
    extern __shared__ double smem[];
    if (a.skip && *a.skip) return;
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
    const int Ny = a.Ny, Nx = a.Nx;
    const int by = cta / g.gx, bx = cta - by * g.gx;
    const int y0 = (int)((long long)by * Ny / g.gy), y1 = (int)((long long)(by + 1) * Ny / g.gy);
    const int x0 = (int)((long long)bx * Nx / g.gx), x1 = (int)((long long)(bx + 1) * Nx / g.gx);
    const int ty = y1 - y0, tx = x1 - x0, PX = tx + 2, PY = ty + 2, plane = PY * PX;
    const int psz = (NT * plane + 1) & ~1;
    double *rs = smem;                                  // [NT][PY][PX]  r, halo ring (zero outside the domain)
    double *xs = rs + psz;                              // [CPT][NTHREADS] x
    double *ws = xs + (XG ? 0 : CPT * NTHREADS);        // [CPT][NTHREADS] w = A r
    double *red = ws + CPT * NTHREADS;                  // reduction scratch: 64 block_sum, 64..66 totals, 72..77 profile
    double *dtab = red + 80;                            // diagonal entries for 3..6 neighbours
    int *hsrc = (int *)(dtab + 4);                      // halo import table: offset into g.edges
    int *hdst = hsrc + 2 * NT * (tx + ty);              //                    index into rs
    int *esrc = hdst + 2 * NT * (tx + ty);              // edge export table: index into rs
    int *edst = esrc + 2 * NT * (tx + ty);              //                    offset into my_edges
    const bool hasN = by > 0, hasS = by < g.gy - 1, hasW = bx > 0, hasE = bx < g.gx - 1;
    const double off = -a.rcoef * 1.0;
    const int edge_stride = NT * g.maxlen;
    double *my_edges = g.edges + (size_t)cta * 4 * edge_stride;
    const int lx = tid % tx, r0 = tid / tx, RPP = NTHREADS / tx;

    // ---- setup
    for (int i = tid; i < psz; i += NTHREADS) rs[i] = 0.0;
    if (tid < 4) dtab[tid] = -a.rcoef * (-(double)(tid + 3)) + a.rcoef * a.eps * 1.0;    // -r*L_ii + r*eps
    int nhalo = 0;
    {
        const int segNS = NT * tx, segWE = NT * ty;
        for (int e = tid; e < segNS; e += NTHREADS) {
            const int t = e / tx, pos = e - t * tx;
            if (hasN) { hsrc[nhalo + e] = ((cta - g.gx) * 4 + 1) * edge_stride + e; hdst[nhalo + e] = (t * PY) * PX + pos + 1; }
        }
        if (hasN) nhalo += segNS;
        for (int e = tid; e < segNS; e += NTHREADS) {
            const int t = e / tx, pos = e - t * tx;
            if (hasS) { hsrc[nhalo + e] = ((cta + g.gx) * 4 + 0) * edge_stride + e; hdst[nhalo + e] = (t * PY + ty + 1) * PX + pos + 1; }
        }
        if (hasS) nhalo += segNS;
        for (int e = tid; e < segWE; e += NTHREADS) {
            const int t = e / ty, pos = e - t * ty;
            if (hasW) { hsrc[nhalo + e] = ((cta - 1) * 4 + 3) * edge_stride + e; hdst[nhalo + e] = (t * PY + pos + 1) * PX; }
        }
        if (hasW) nhalo += segWE;
        for (int e = tid; e < segWE; e += NTHREADS) {
            const int t = e / ty, pos = e - t * ty;
            if (hasE) { hsrc[nhalo + e] = ((cta + 1) * 4 + 2) * edge_stride + e; hdst[nhalo + e] = (t * PY + pos + 1) * PX + tx + 1; }
        }
        if (hasE) nhalo += segWE;
    }
    int nexp = 0;
    {
        const int segNS = NT * tx, segWE = NT * ty;
        for (int e = tid; e < segNS; e += NTHREADS) {
            const int t = e / tx, pos = e - t * tx;
            if (hasN) { esrc[nexp + e] = (t * PY + 1) * PX + pos + 1; edst[nexp + e] = 0 * edge_stride + e; }
        }
        if (hasN) nexp += segNS;
        for (int e = tid; e < segNS; e += NTHREADS) {
            const int t = e / tx, pos = e - t * tx;
            if (hasS) { esrc[nexp + e] = (t * PY + ty) * PX + pos + 1; edst[nexp + e] = 1 * edge_stride + e; }
        }
        if (hasS) nexp += segNS;
        for (int e = tid; e < segWE; e += NTHREADS) {
            const int t = e / ty, pos = e - t * ty;
            if (hasW) { esrc[nexp + e] = (t * PY + pos + 1) * PX + 1; edst[nexp + e] = 2 * edge_stride + e; }
        }
        if (hasW) nexp += segWE;
        for (int e = tid; e < segWE; e += NTHREADS) {
            const int t = e / ty, pos = e - t * ty;
            if (hasE) { esrc[nexp + e] = (t * PY + pos + 1) * PX + tx; edst[nexp + e] = 3 * edge_stride + e; }
        }
        if (hasE) nexp += segWE;
    }
    // patch ownership: first tile row, number of owned rows, index of cell (t = 0, jy = 0) in rs
    const int ly0 = r0 * YPT;
    const int nval = r0 < RPP ? min(YPT, max(ty - ly0, 0)) : 0;
    const int sb = (ly0 + 1) * PX + lx + 1;
    const int xmiss = (x0 + lx == 0) + (x0 + lx == Nx - 1);
    const int jTop = (by == 0 && r0 == 0) ? 0 : -1;            // owned row on the global y = 0 boundary
    const int jBot = (by == g.gy - 1) ? ty - 1 - ly0 : -1;      //                        y = Ny-1
    const double dg_ti = -a.rcoef * (-(double)(6 - xmiss)) + a.rcoef * a.eps * 1.0;
    const double dg_tb = -a.rcoef * (-(double)(5 - xmiss)) + a.rcoef * a.eps * 1.0;
    auto fresh = [](int v) { asm volatile("" : "+r"(v)); return v; };

    double pj[CPT], sj[CPT];
    __syncthreads();                                     // rs zeroed before the owners fill it
#pragma unroll
    for (int j = 0; j < CPT; j++) { pj[j] = 0.0; sj[j] = 0.0; if (!XG) xs[j * NTHREADS + tid] = 0.0; ws[j * NTHREADS + tid] = 0.0; }
#pragma unroll
    for (int jy = 0; jy < YPT; jy++) {
        if (jy < nval) {
#pragma unroll
            for (int t = 0; t < NT; t++) {
                const double v = a.b[((size_t)t * Ny + (y0 + ly0 + jy)) * Nx + (x0 + lx)];
                rs[sb + t * plane + jy * PX] = v;
                if (XG) a.x[((size_t)t * Ny + (y0 + ly0 + jy)) * Nx + (x0 + lx)] = 0.0;
            }
        }
    }
    // export pass (all threads, after a __syncthreads that follows the writes of rs): the tile-edge values of
    // generation gn are rounded to its parity in place (so the owner and the neighbour use the same value) and stored;
    // a corner cell sits in two lists and is rounded twice to the same value
    auto export_edges = [&](unsigned int gn) {
        const long long par = (long long)(gn & 1u);
#pragma unroll
        for (int e = 0; e < kHaloPerThread; e++) {
            const int h = tid + e * NTHREADS;
            if (h < nexp) {
                const int si = esrc[h];
                const double v = __longlong_as_double((__double_as_longlong(rs[si]) & ~1ll) | par);
                rs[si] = v;
                st_relaxed_u64((unsigned long long *)(my_edges + edst[h]), (unsigned long long)__double_as_longlong(v));
            }
        }
    };
    __syncthreads();
    export_edges(0u);

    // x slot j += c * p_j  (slots [j0, j1)); XG: read-modify-write of the owned cell in global memory
    double *const xg = a.x + ((size_t)(y0 + ly0)) * Nx + (x0 + lx);
    const size_t Pst = (size_t)Ny * Nx;
    auto x_update = [&](int j0, int j1, double c) {
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (j < j0 || j >= j1) continue;
            if (XG) {
                const int t = j / YPT, jy = j - t * YPT;
                if (jy < nval) { double *px = xg + t * Pst + (size_t)jy * Nx; *px = *px + c * pj[j]; }
            } else {
                const int xi = j * NTHREADS + tid;
                xs[xi] = xs[xi] + c * pj[j];
            }
        }
    };
    unsigned int gen = 0;
    bool abort = false;
    long long tmark = 0;
    const bool prof = g.prof != nullptr && tid == 0;
    long long *sprof = (long long *)(red + 72);           // shared-memory accumulators (thread 0 only)
    if (tid == 0) { for (int k = 0; k < 6; k++) sprof[k] = 0; red[66] = 0.0; }
    auto lap = [&](int k) { if (prof) { long long now = clock64(); sprof[k] += now - tmark; tmark = now; } };

    int it = 0, status = a.maxiter;
    double gam_stop = 0.0, rgam_prev = 0.0, d_prev = 0.0, alpha_prev = 0.0;
    bool pend = false;                                   // x += alpha_prev p not yet applied
    if (prof) tmark = clock64();
    for (; it < a.maxiter; it++) {
        // ---- import the neighbours' edge values of generation `it`: spin on every word until its parity is it & 1
        {
            const double *eg = g.edges;
            const unsigned long long par = (unsigned long long)(it & 1);
            double hv[kHaloPerThread];
            const long long t0 = clock64();
            bool ready;
            do {
                ready = true;
#pragma unroll
                for (int e = 0; e < kHaloPerThread; e++) {
                    const int h = tid + e * NTHREADS;
                    hv[e] = 0.0;
                    if (h < nhalo) {
                        const unsigned long long bits = ld_relaxed_u64((const unsigned long long *)(eg + hsrc[h]));
                        ready = ready && (bits & 1ull) == par;
                        hv[e] = __longlong_as_double((long long)bits);
                    }
                }
                if (!ready && clock64() - t0 > kWatchdogCycles) { red[66] = 1.0; break; }     // a neighbour is stuck: abort below
            } while (!ready);
#pragma unroll
            for (int e = 0; e < kHaloPerThread; e++) {
                const int h = tid + e * NTHREADS;
                if (h < nhalo) rs[hdst[h]] = hv[e];
            }
        }
        __syncthreads();
        lap(0);
        // ---- w = A r (csr_matvec order), partial r.r and r.w
        double acc[2] = {0.0, 0.0};
        if (nval > 0) {
            double up[NT], cur[NT], nxt[NT];             // rolling window over the owned rows, all NT levels
            const double *pb = rs + fresh(sb);
#pragma unroll
            for (int t = 0; t < NT; t++) { up[t] = pb[t * plane - PX]; cur[t] = pb[t * plane]; }
#pragma unroll
            for (int jy = 0; jy < YPT; jy++) {
                if (jy < nval) {
#pragma unroll
                    for (int t = 0; t < NT; t++) nxt[t] = pb[t * plane + (jy + 1) * PX];
                    const int ym = (jy == jTop) + (jy == jBot);
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        const double *px = pb + t * plane + jy * PX;
                        const bool tb = t == 0 || t == NT - 1;
                        double dg = tb ? dg_tb : dg_ti;
                        if (ym) dg = dtab[(tb ? 5 : 6) - xmiss - ym - 3];
                        const double c = cur[t];
                        double s = 0.0;
                        if (UNIT) {                      // r == 1: products with -1.0 are exact negations
                            if (t > 0) s -= cur[t - 1];
                            s -= up[t]; s -= px[-1];
                            s += dg * c;
                            s -= px[1]; s -= nxt[t];
                            if (t < NT - 1) s -= cur[t + 1];
                        } else {
                            if (t > 0) s += off * cur[t - 1];
                            s += off * up[t]; s += off * px[-1];
                            s += dg * c;
                            s += off * px[1]; s += off * nxt[t];
                            if (t < NT - 1) s += off * cur[t + 1];
                        }
                        ws[(t * YPT + jy) * NTHREADS + tid] = s;
                        acc[0] = fma(c, c, acc[0]);
                        acc[1] = fma(c, s, acc[1]);
                    }
#pragma unroll
                    for (int t = 0; t < NT; t++) { up[t] = cur[t]; cur[t] = nxt[t]; }
                }
            }
        }
        lap(1);
        // ---- the one all-reduce; the x update of the previous iteration runs in its shadow
        block_sum<2>(acc, red);
        if (tid == 0) grid_arrive<2>(g.slots, gen, acc, red[66] != 0.0);    // an import watchdog of this CTA aborts the whole grid
        if (cta == 0 && tid < 32) grid_root<2>(g.slots, gen, ncta, tid);
        if (pend) {                                      // second half of x += alpha_prev p (first half: after the export)
            x_update(CPT / 2, CPT, alpha_prev);
            pend = false;
        }
        if (tid == 0) {
            if (!grid_wait<2>(g.slots, gen, red + 64)) red[66] = 1.0;
        }
        __syncthreads();
        gen++;
        const double gam = red[64], del = red[65];
        abort = red[66] != 0.0;
        lap(2);
        if (abort) break;
        if (it == 0) {
            if (gam == 0.0) { status = 0; break; }       // scipy: "if bnrm2 == 0: return b, 0"
            // scipy stops when sqrt(gamma) < atol, atol = rtol sqrt(gamma_0).  sqrt is monotone and correctly rounded,
            // so that is gamma < gam_stop with gam_stop the smallest double whose square root reaches atol: found once,
            // the per-iteration test is then a comparison (no sqrt on the critical path)
            const double atol = a.rtol * sqrt(gam);
            double t = atol * atol;
            for (int k = 0; k < 8 && sqrt(t) < atol; k++) t = __longlong_as_double(__double_as_longlong(t) + 1);
            for (int k = 0; k < 8 && t > 0.0 && sqrt(__longlong_as_double(__double_as_longlong(t) - 1)) >= atol; k++)
                t = __longlong_as_double(__double_as_longlong(t) - 1);
            gam_stop = t;
        }
        if (gam < gam_stop) { status = 0; break; }       // scipy's "||r|| < atol" at the top of the iteration
        // beta = gamma / gamma_old with the reciprocal taken in the shadow of the all-reduce; with d = gamma / alpha:
        // d = delta - beta^2 d_old (= delta - beta gamma / alpha_old), alpha = gamma / d: one division after the barrier
        const double beta = gam * rgam_prev;
        const double dk = del - (beta * beta) * d_prev;
        const double alpha = gam / dk;
        // ---- p = r + beta p, s = w + beta s, r -= alpha s; tile-edge values exported as generation it+1
#pragma unroll
        for (int jy = 0; jy < YPT; jy++) {
            if (jy < nval) {
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const int j = t * YPT + jy, si = fresh(sb) + t * plane + jy * PX;
                    const double rv = rs[si], wv = ws[j * NTHREADS + tid];
                    const double sv = sj[j] * beta + wv;
                    pj[j] = pj[j] * beta + rv;
                    sj[j] = sv;
                    rs[si] = rv - alpha * sv;
                }
            }
        }
        __syncthreads();
        export_edges((unsigned int)it + 1u);
        lap(3);
        // first half of x += alpha p while the edge values travel
        x_update(0, CPT / 2, alpha);
        lap(4);
        pend = true; alpha_prev = alpha; d_prev = dk;
        rgam_prev = 1.0 / gam;                           // not needed before the next all-reduce has completed
    }
    if (abort) { if (tid == 0) *a.sync.error = 1; return; }
    // ---- write phi (apply the update that was still waiting for a barrier shadow: only after maxiter iterations)
#pragma unroll
    for (int jy = 0; jy < YPT; jy++) {
        if (jy < nval) {
#pragma unroll
            for (int t = 0; t < NT; t++) {
                const int j = t * YPT + jy;
                double *px = a.x + ((size_t)t * Ny + (y0 + ly0 + jy)) * Nx + (x0 + lx);
                double xv = XG ? *px : xs[j * NTHREADS + tid];
                if (pend && j >= CPT / 2) xv = xv + alpha_prev * pj[j];
                if (!XG || (pend && j >= CPT / 2)) *px = xv;
            }
        }
    }
    if (cta == 0 && tid == 0) { a.out[0] = it; a.out[1] = status; }
    if (prof) { for (int k = 0; k < 6; k++) g.prof[cta * 8 + k] += sprof[k]; g.prof[cta * 8 + 6] += it; }
}

__global__ void k_fill2_u64(unsigned long long *p, int n, unsigned long long v, unsigned long long *q, int m, unsigned long long w)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
    if (i < m) q[i] = w;
}

// patch shapes: NT time levels x YPT rows per thread (<= 16 cell slots: p and s take 4 registers per slot, the
// stencil window 6 NT); Nt = 4 is the CLI default and the benchmark configuration
struct Shape { int nt, ypt, threads; bool xg; const void *unit, *general; };
#define FOTO_FUSED_SHAPE(NT, YPT, T, XG) {NT, YPT, T, XG, (const void *)cg_fused_kernel<T, NT, YPT, true, XG>, (const void *)cg_fused_kernel<T, NT, YPT, false, XG>}
// listed fastest first per Nt.  Nt = 4: 448 threads (14 warps: 2 % faster than 16 at 388x584x4, 7 168 cell slots), 512
// threads (8 192), then 384 threads x 24 slots with x in global memory (9 216; 576 threads x 16 slots are capped at 96
// registers and spill: 10.2 against 7.7 us per iteration at 480x640x4)
const Shape kShapes[] = {FOTO_FUSED_SHAPE(2, 8, 512, false), FOTO_FUSED_SHAPE(3, 5, 512, false), FOTO_FUSED_SHAPE(4, 4, 448, false),
                         FOTO_FUSED_SHAPE(4, 4, 512, false), FOTO_FUSED_SHAPE(5, 3, 512, false), FOTO_FUSED_SHAPE(4, 6, 384, true),
                         FOTO_FUSED_SHAPE(6, 3, 384, false), FOTO_FUSED_SHAPE(7, 2, 384, false), FOTO_FUSED_SHAPE(8, 2, 384, false),
                         FOTO_FUSED_SHAPE(16, 1, 256, false)};
constexpr int kNumShapes = sizeof(kShapes) / sizeof(kShapes[0]);

struct Plan { bool ok = false; int shape = 0, gy = 0, gx = 0, maxlen = 0, ncta = 0; size_t smem = 0; };

// Tile grid: gy*gx <= #SMs, every tile fits the 4 x 4 patches of 512 threads, the halo tables and shared memory.
// Every grid that fits costs an active thread the same 16 cell slots; what differs is measured (tools/sweep_grid.py,
// 388x584x4, SWEEP_VARIANT=2: 5.87 us per iteration for 12x12, 5.96 for 8x18, 6.2-6.6 for the rest): tiles whose
// width is a multiple of 16 (+1) keep a half-warp inside one row group, i.e. free of shared-memory bank conflicts;
// then short halos and few idle SMs; then wide rows.
Plan make_plan(OnchipScratch &d, int device, int Nt, int Ny, int Nx)
{
    Plan best;
    if (!d.num_sms) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return best;
        d.num_sms = prop.multiProcessorCount; d.smem_optin = prop.sharedMemPerBlockOptin;
    }
    for (int shape = 0; shape < kNumShapes && !best.ok; shape++) {
    if (kShapes[shape].nt != Nt) continue;
    const int kNT = Nt, kYPT = kShapes[shape].ypt, kThreads = kShapes[shape].threads;
    const int xslots = kShapes[shape].xg ? 1 : 2;       // private shared-memory arrays of kNT*kYPT*kThreads doubles
    best.shape = shape;
    int force_gy = 0, force_gx = 0;                      // FOTO_ONCHIP_GRID=gy,gx: pin the tile grid (experiments)
    if (const char *e = getenv("FOTO_ONCHIP_GRID")) sscanf(e, "%d,%d", &force_gy, &force_gx);
    long long best_key = -1;
    for (int gy = 1; gy <= d.num_sms && gy <= Ny; gy++)
        for (int gx = 1; gx <= d.num_sms / gy && gx <= Nx; gx++) {
            if (force_gy > 0 && (gy != force_gy || gx != force_gx)) continue;
            const int ty = (Ny + gy - 1) / gy, tx = (Nx + gx - 1) / gx, ty_min = Ny / gy, tx_min = Nx / gx;
            if (tx > kThreads || ty_min < 1 || tx_min < 1) continue;
            bool fits = true;
            for (int txx = tx_min; txx <= tx; txx++) if (ty > kYPT * (kThreads / txx)) fits = false;
            if (!fits) continue;
            if (2LL * kNT * (tx + ty) > (long long)kHaloPerThread * kThreads) continue;
            const size_t smem = ((((size_t)kNT * (ty + 2) * (tx + 2) + 1) & ~size_t(1)) + (size_t)xslots * kNT * kYPT * kThreads + 80 + 4) * 8
                              + (size_t)8 * kNT * (tx + ty) * sizeof(int);
            if (smem > d.smem_optin) continue;
            const int straddle = (tx % 16) > 1 ? 1 : 0;
            const long long key = ((straddle * 100000LL + 2LL * kNT * (tx + ty) + 8LL * (d.num_sms - gy * gx)) * 1000) + (999 - tx);
            if (best_key < 0 || key < best_key) {
                best_key = key; best.ok = true; best.gy = gy; best.gx = gx; best.ncta = gy * gx;
                best.maxlen = tx > ty ? tx : ty; best.smem = smem;
            }
        }
    }
    return best;
}

}  // namespace

bool cg_fused_fits(OnchipScratch &s, int device, int Nt, int Ny, int Nx) { return make_plan(s, device, Nt, Ny, Nx).ok; }

void onchip_release(OnchipScratch &s)
{
    cudaFree(s.prof); cudaFree(s.fused_edges); cudaFree(s.fused_slots); cudaFree(s.gnf_edges); cudaFree(s.gnf_slots);
    s.prof = nullptr;
    s.fused_edges = nullptr; s.fused_slots = nullptr; s.fused_edges_bytes = 0;
    s.gnf_edges = nullptr; s.gnf_slots = nullptr; s.gnf_edges_bytes = 0;
}

int launch_cg_fused(cudaStream_t st, const CgArgs &a, int device, OnchipScratch &d)
{
    Plan p = make_plan(d, device, a.Nt, a.Ny, a.Nx);
    if (!p.ok) { set_error("grid %dx%dx%d does not fit the single-reduction on-chip CG variant", a.Nt, a.Ny, a.Nx); return FOTO_ERR_ARG; }
    const size_t need = (size_t)p.ncta * 4 * a.Nt * p.maxlen * sizeof(double);
    if (d.fused_edges_bytes < need) {
        if (d.fused_edges) CUDA_TRY(cudaFree(d.fused_edges));
        CUDA_TRY(cudaMalloc((void **)&d.fused_edges, need));
        d.fused_edges_bytes = need;
    }
    if (!d.fused_slots) CUDA_TRY(cudaMalloc((void **)&d.fused_slots, kSlotWords * sizeof(unsigned long long)));
    const void *fn = a.rcoef == 1.0 ? kShapes[p.shape].unit : kShapes[p.shape].general;
    if (!d.fused_attr_set) {
        for (int i = 0; i < kNumShapes; i++) {
            CUDA_TRY(cudaFuncSetAttribute(kShapes[i].unit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_optin));
            CUDA_TRY(cudaFuncSetAttribute(kShapes[i].general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_optin));
        }
        d.fused_attr_set = true;
    }
    const int nedge = (int)(need / sizeof(double));
    // all-reduce slots and edge words <- odd parity ("generation 0 not yet written")
    k_fill2_u64<<<((nedge > kSlotWords ? nedge : kSlotWords) + 255) / 256, 256, 0, st>>>(d.fused_slots, kSlotWords, kSlotInit, (unsigned long long *)d.fused_edges, nedge, ~0ull);
    Geom g;
    g.gy = p.gy; g.gx = p.gx; g.maxlen = p.maxlen; g.edges = d.fused_edges; g.slots = d.fused_slots; g.prof = d.prof;
    void *args[] = {(void *)&a, (void *)&g};
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(p.ncta), dim3(kShapes[p.shape].threads), args, p.smem, st));
    return FOTO_OK;
}

}  // namespace foto
