// cg_onchip.cu -- K2a, on-chip resident variant: the whole CG state of one Poisson solve lives in
// the shared memory and registers of the SMs for the ~670 iterations of a solve.
//
// Why: at Middlebury size (N = 0.9 M cells) the four CG vectors (29 MB) fit in L2, so the
// streaming kernel is bound by L2 bandwidth and grid-barrier latency (16 us per iteration),
// not by HBM.  6 k cells per SM x (x, r, p, q) = 200 KB does fit in one SM's 227 KB of shared
// memory + 256 KB of registers, so this kernel
//   * cuts the (y, x) plane into gy x gx tiles, one CTA (= one SM) per tile, all Nt planes,
//   * keeps r and q in registers, p (with a one-cell halo ring and two zero planes) and x in
//     shared memory,
//   * per iteration touches global memory only for the tile-edge values of r (written after
//     the r update, read by the four neighbours after the barrier) and for the all-reduce
//     slots -- about 10 KB per SM per iteration instead of 0.5 MB,
//   * synchronises the grid with a two-hop reduction: every CTA publishes its partial dot
//     product with one release store; one warp of CTA 0 gathers the partials, sums them in a
//     fixed order and publishes the total; one thread per CTA polls that single word.  Only
//     148 + 32 threads poll L2 and every CTA reads the same total.  Measured alternatives, all
//     slower (profiles/README.md): all-to-all polling of one slot array (148 x 148 threads on ten
//     hot L2 lines, 8 k cycles per barrier), 16 replicated slot arrays polled by 9 CTAs each
//     (one hop, but 9.6 instead of 8.9 us per iteration), one private total word per CTA.
// The kernel is instruction-issue and shared-memory bound, so the per-cell code is branch-free:
// out-of-domain neighbours are zero cells (subtracting +0.0 is exact, the csr_matvec running
// sum is unchanged), unused cell slots point at a dead cell and their q is masked to zero.
// The recurrence, the summation order of A p and the stopping rule are those of
// cg_stream_kernel (cg_kernels.cu); halo cells of p are advanced with the same p = p*beta + r
// arithmetic as their owner, so both variants produce the same iterates up to the summation
// order of the dot products.
#include "foto_kernels.cuh"

namespace foto {

namespace {

constexpr unsigned long long kSentinel = 0x7FF8DEADBEEF0001ull;   // NaN payloads no sum produces
constexpr unsigned long long kAbort = 0x7FF8DEADBEEF0002ull;
constexpr unsigned long long kPlainNaN = 0x7FF8000000000000ull;
constexpr int kBcastOff = 3 * kMaxBlocks;     // slots: [3][kMaxBlocks] partials, then 3 x 16 words broadcast
constexpr int kSlotWords = kBcastOff + 3 * 16;

struct OnchipGeom {
    int gy, gx;                  // tile grid (gy * gx CTAs)
    int maxlen;                  // longest tile edge, in cells
    double *edges;               // [ncta][4 (N,S,W,E)][Nt * maxlen] tile-edge values of r
    unsigned long long *slots;   // all kSentinel at launch
    long long *prof;             // optional: [ncta][8] cycle counters, thread 0 of each CTA (debugging aid)
};

__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

constexpr long long kWatchdogCycles = 8000000000ll;    // ~4 s

// Thread 0 of every CTA, after a __syncthreads that follows the CTA's global writes (the release
// is cumulative over them).  Re-arms this CTA's slot of generation gen+1, which nobody reads
// before this CTA has arrived there.
__device__ __forceinline__ void grid_arrive(const OnchipGeom &g, unsigned int gen, double v, bool publish)
{
    unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    if (bits == kSentinel || bits == kAbort) bits = kPlainNaN;
    st_relaxed_u64(g.slots + ((gen + 1u) % 3u) * kMaxBlocks + blockIdx.x, kSentinel);
    // Release fence (~900 cycles) only when the CTA has published tile-edge values since the last
    // barrier.  The re-arm store above needs no fence: it is ordered before the value store by
    // same-thread program order to the same L2 slice set and nobody reads that slot for two barriers.
    if (publish) fence_acq_rel_gpu();
    st_relaxed_u64(g.slots + (gen % 3u) * kMaxBlocks + blockIdx.x, bits);
}

// Warp 0 of CTA 0: gather all partials of generation gen (lanes poll their slots with all loads
// in flight), sum in a fixed order, publish the total.
__device__ __forceinline__ void grid_root(const OnchipGeom &g, unsigned int gen, int ncta, int lane)
{
    const unsigned long long *cur = g.slots + (gen % 3u) * kMaxBlocks;
    const long long t0 = clock64();
    double s = 0.0;
    bool abort = false;
    for (int base = 0; base < ncta; base += 256) {          // 8 slots per lane per round
        unsigned long long v[8];
        bool ready;
        do {
            ready = true;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int b = base + k * 32 + lane;
                v[k] = b < ncta ? ld_relaxed_u64(cur + b) : 0ull;        // relaxed: the 8 polls pipeline
                ready = ready && v[k] != kSentinel;
            }
            if (!ready && clock64() - t0 > kWatchdogCycles) { abort = true; break; }
        } while (!ready);
#pragma unroll
        for (int k = 0; k < 8; k++) s += __longlong_as_double((long long)v[k]);
    }
    s = warp_sum(s);
    abort = __any_sync(0xffffffffu, abort);
    if (lane == 0) {
        unsigned long long bits = (unsigned long long)__double_as_longlong(s);
        if (bits == kSentinel || bits == kAbort) bits = kPlainNaN;
        st_relaxed_u64(g.slots + kBcastOff + ((gen + 1u) % 3u) * 16, kSentinel);
        // no fence: the total is data-dependent on the polled partials, and every partial was
        // stored after its CTA's release fence, so the edge values are already performed at L2
#ifdef FOTO_PARANOID_FENCES
        fence_acq_rel_gpu();
#endif
        st_relaxed_u64(g.slots + kBcastOff + (gen % 3u) * 16, abort ? kAbort : bits);
    }
}

// Thread 0 of every CTA: wait for the total of generation gen.
__device__ __forceinline__ double grid_wait(const OnchipGeom &g, unsigned int gen, bool &abort)
{
    const unsigned long long *p = g.slots + kBcastOff + (gen % 3u) * 16;
    const long long t0 = clock64();
    unsigned long long bits;
    while ((bits = ld_relaxed_u64(p)) == kSentinel) {
        if (clock64() - t0 > 2 * kWatchdogCycles) { bits = kAbort; break; }
    }
    // Readers fetch the neighbours' edge values with ld.global.cg (L2, never L1) after this
    // control-dependent loop and a __syncthreads, so they observe the released values without an
    // acquire fence (another ~800 cycles).  -DFOTO_PARANOID_FENCES restores it.
#ifdef FOTO_PARANOID_FENCES
    fence_acq_rel_gpu();
#endif
    abort = bits == kAbort;
    return __longlong_as_double((long long)bits);
}

// per-cell descriptor (one int per owned cell slot, computed once):
//   bits 0-15  index of the cell in the shared p array      bits 16-21  t
//   bits 22-23 number of existing neighbours minus 3        bits 24-27  export to N/S/W/E edge
//   bit 28     slot holds a cell
constexpr int kSiMask = 0xFFFF;
constexpr int kTShift = 16, kTMask = 0x3F;
constexpr int kCntShift = 22;
constexpr int kEdgeN = 1 << 24, kEdgeS = 1 << 25, kEdgeW = 1 << 26, kEdgeE = 1 << 27, kEdgeAny = 0xF << 24;
constexpr int kValid = 1 << 28;
constexpr int kHaloPerThread = 4;      // halo cells per thread the plan guarantees (nhalo <= 4 * threads)

// NT == 0: generic cell ownership, cell slot j of a thread is row j*RPP + r0 of the (t, y) rows of the tile.
// NT > 0 (requires Nt == NT, CPT == NT*YPT): patch ownership, a thread owns all NT time levels of YPT consecutive
// tile rows at its column, so the t and inner y neighbours of the stencil come from registers (3.5 shared-memory
// loads per cell instead of 7; the stencil phase is shared-memory-bandwidth bound).
template <int NTHREADS, int CPT, bool UNIT, int NT = 0, int YPT = 0>
__global__ void __launch_bounds__(NTHREADS, 1) cg_onchip_kernel(CgArgs a, OnchipGeom g)
{
    constexpr bool PATCH = NT > 0;
    static_assert(!PATCH || CPT == NT * YPT, "patch ownership: CPT = NT * YPT");
    extern __shared__ double smem[];
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
    const int Nt = a.Nt, Ny = a.Ny, Nx = a.Nx;
    const int by = cta / g.gx, bx = cta - by * g.gx;
    const int y0 = (int)((long long)by * Ny / g.gy), y1 = (int)((long long)(by + 1) * Ny / g.gy);
    const int x0 = (int)((long long)bx * Nx / g.gx), x1 = (int)((long long)(bx + 1) * Nx / g.gx);
    const int ty = y1 - y0, tx = x1 - x0, PX = tx + 2, PY = ty + 2, plane = PY * PX;
    const int psz = ((Nt + 2) * plane + 1) & ~1;
    double *ps = smem;                                  // [Nt+2][PY][PX]  p, halo ring, zero planes
    double *xs = ps + psz;                              // [CPT][NTHREADS] x
    double *red = xs + CPT * NTHREADS;                  // reduction scratch (64)
    double *dtab = red + 64;                            // diagonal entries for 3..6 neighbours
    int *hsrc = (int *)(dtab + 4);                      // halo import table: offset into g.edges
    int *hdst = hsrc + 2 * Nt * (tx + ty);              //                    index into ps
    const bool hasN = by > 0, hasS = by < g.gy - 1, hasW = bx > 0, hasE = bx < g.gx - 1;
    const double off = -a.rcoef * 1.0;
    const int edge_stride = Nt * g.maxlen;
    double *my_edges = g.edges + (size_t)cta * 4 * edge_stride;
    const int lx = tid % tx, r0 = tid / tx, RPP = NTHREADS / tx;
    // dead cell: in the upper zero plane; its 7-point neighbourhood stays inside the shared
    // allocation (the +plane read lands in xs), and everything derived from it is masked.
    const int sdead = ((Nt + 1) * PY + 1) * PX + 1;

    // export pointers: N/S edges are indexed [t*tx + lx], W/E edges by the (t, y) row j*RPP + r0
    double *const pN = my_edges + lx, *const pS = my_edges + edge_stride + lx;
    double *const pW = my_edges + 2 * edge_stride + r0, *const pE = my_edges + 3 * edge_stride + r0;
    const double dg_interior = -a.rcoef * (-6.0) + a.rcoef * a.eps * 1.0;
    double rj[CPT], qj[CPT];        // p lives in shared memory: the register file cannot hold a third vector
    int info[CPT];

    // ---- setup
    for (int i = tid; i < psz; i += NTHREADS) ps[i] = 0.0;
    if (tid < 4) dtab[tid] = -a.rcoef * (-(double)(tid + 3)) + a.rcoef * a.eps * 1.0;    // -r*L_ii + r*eps
    int nhalo = 0;
    {
        const int segNS = Nt * tx, segWE = Nt * ty;
        for (int e = tid; e < segNS; e += NTHREADS) {
            const int t = e / tx, pos = e - t * tx;
            if (hasN) { hsrc[nhalo + e] = ((cta - g.gx) * 4 + 1) * edge_stride + e; hdst[nhalo + e] = ((t + 1) * PY) * PX + pos + 1; }
        }
        if (hasN) nhalo += segNS;
        for (int e = tid; e < segNS; e += NTHREADS) {
            const int t = e / tx, pos = e - t * tx;
            if (hasS) { hsrc[nhalo + e] = ((cta + g.gx) * 4 + 0) * edge_stride + e; hdst[nhalo + e] = ((t + 1) * PY + ty + 1) * PX + pos + 1; }
        }
        if (hasS) nhalo += segNS;
        for (int e = tid; e < segWE; e += NTHREADS) {
            const int t = e / ty, pos = e - t * ty;
            if (hasW) { hsrc[nhalo + e] = ((cta - 1) * 4 + 3) * edge_stride + e; hdst[nhalo + e] = ((t + 1) * PY + pos + 1) * PX; }
        }
        if (hasW) nhalo += segWE;
        for (int e = tid; e < segWE; e += NTHREADS) {
            const int t = e / ty, pos = e - t * ty;
            if (hasE) { hsrc[nhalo + e] = ((cta + 1) * 4 + 2) * edge_stride + e; hdst[nhalo + e] = ((t + 1) * PY + pos + 1) * PX + tx + 1; }
        }
        if (hasE) nhalo += segWE;
    }
    double acc = 0.0;
    // patch ownership: first tile row, number of owned rows, base index of cell (t = 0, jy = 0) in ps
    const int ly0 = r0 * (PATCH ? YPT : 1);
    const int nval = (PATCH && r0 < RPP) ? min(YPT, max(ty - ly0, 0)) : 0;
    const int sb = (PY + ly0 + 1) * PX + lx + 1;
    const int xmiss = (x0 + lx == 0) + (x0 + lx == Nx - 1);
    const int jTop = (by == 0 && r0 == 0) ? 0 : -1;            // owned row on the global y = 0 boundary
    const int jBot = (by == g.gy - 1) ? ty - 1 - ly0 : -1;      //                        y = Ny-1
    const int jS = hasS ? ty - 1 - ly0 : -1;                    // owned row exported to the southern neighbour
    const bool expN = hasN && r0 == 0, expW = hasW && lx == 0, expE = hasE && lx == tx - 1;
    if (PATCH) {
#pragma unroll
        for (int j = 0; j < CPT; j++) { rj[j] = 0.0; qj[j] = 0.0; info[j] = 0; xs[j * NTHREADS + tid] = 0.0; }
#pragma unroll
        for (int jy = 0; jy < (PATCH ? YPT : 0); jy++) {
            if (jy < nval) {
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const int ly = ly0 + jy;
                    const double v = a.b[((size_t)t * Ny + (y0 + ly)) * Nx + (x0 + lx)];
                    if (expN && jy == 0) my_edges[0 * edge_stride + t * tx + lx] = v;
                    if (jy == jS) my_edges[1 * edge_stride + t * tx + lx] = v;
                    if (expW) my_edges[2 * edge_stride + t * ty + ly] = v;
                    if (expE) my_edges[3 * edge_stride + t * ty + ly] = v;
                    rj[t * YPT + jy] = v;
                    acc = fma(v, v, acc);
                }
            }
        }
    } else {
        const int rows = Nt * ty;
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const int row = j * RPP + r0;
            int inf = sdead;
            double v = 0.0;
            if (r0 < RPP && row < rows) {
                const int t = row / ty, ly = row - t * ty;
                const int cnt = (t > 0) + (t < Nt - 1) + (y0 + ly > 0) + (y0 + ly < Ny - 1) + (x0 + lx > 0) + (x0 + lx < Nx - 1);
                inf = (((t + 1) * PY + ly + 1) * PX + lx + 1) | (t << kTShift) | ((cnt - 3) << kCntShift) | kValid;
                if (ly == 0 && hasN) inf |= kEdgeN;
                if (ly == ty - 1 && hasS) inf |= kEdgeS;
                if (lx == 0 && hasW) inf |= kEdgeW;
                if (lx == tx - 1 && hasE) inf |= kEdgeE;
                v = a.b[((size_t)t * Ny + (y0 + ly)) * Nx + (x0 + lx)];
                if (inf & kEdgeN) my_edges[0 * edge_stride + t * tx + lx] = v;
                if (inf & kEdgeS) my_edges[1 * edge_stride + t * tx + lx] = v;
                if (inf & kEdgeW) my_edges[2 * edge_stride + row] = v;
                if (inf & kEdgeE) my_edges[3 * edge_stride + row] = v;
            }
            info[j] = inf;
            rj[j] = v; qj[j] = 0.0;
            xs[j * NTHREADS + tid] = 0.0;
            acc = fma(v, v, acc);
        }
    }
    // patch ownership: diagonal of a cell with 6 - xmiss (t interior) or 5 - xmiss (t = 0, NT-1) neighbours
    const double dg_ti = -a.rcoef * (-(double)(6 - xmiss)) + a.rcoef * a.eps * 1.0;
    const double dg_tb = -a.rcoef * (-(double)(5 - xmiss)) + a.rcoef * a.eps * 1.0;
    double *const qW = my_edges + 2 * edge_stride + ly0, *const qE = my_edges + 3 * edge_stride + ly0;
    unsigned int gen = 0;
    bool abort = false;
    long long tmark = 0;
    const bool prof = g.prof != nullptr && tid == 0;
    long long *sprof = (long long *)(red + 40);        // shared-memory accumulators (thread 0 only)
    if (tid == 0) { for (int k = 0; k < 6; k++) sprof[k] = 0; }
    auto lap = [&](int k) { if (prof) { long long now = clock64(); sprof[k] += now - tmark; tmark = now; } };
    // opaque copy of a descriptor: keeps the per-cell address arithmetic inside the loop instead
    // of hoisted into spilled registers
    auto fresh = [](int v) { asm volatile("" : "+r"(v)); return v; };

    // grid-wide sum of `val` (+ barrier).  `overlap` runs between arrive and wait.
    auto grid_sum = [&](double val, bool publish, auto overlap) -> double {
        double v1[1] = {val};
        block_sum<1>(v1, red);
        if (tid == 0) grid_arrive(g, gen, v1[0], publish);
        if (cta == 0 && tid < 32) grid_root(g, gen, ncta, tid);     // before the overlap work: all CTAs wait on it
        overlap();
        if (tid == 0) {
            bool ab;
            const double total = grid_wait(g, gen, ab);
            red[32] = total; red[33] = ab ? 1.0 : 0.0;
        }
        __syncthreads();
        gen++;
        abort = red[33] != 0.0;
        return red[32];
    };
    auto nothing = [] {};

    const double bb = grid_sum(acc, true, nothing);
    int it = 0, status = bb == 0.0 ? 0 : a.maxiter;      // scipy: "if bnrm2 == 0: return b, 0"
    if (!abort && bb != 0.0) {
        const double atol = a.rtol * sqrt(bb);
        double rr = bb, rr_prev = 0.0;
        if (prof) tmark = clock64();
        for (; it < a.maxiter; it++) {
            const double beta = it > 0 ? rr / rr_prev : 0.0;
            // ---- A1 (first half): request the neighbours' freshly published r (L2 round trip ~700 cycles)
            // before touching own cells; halo tables hold at most kHaloPerThread entries per thread
            double hv[kHaloPerThread];
#pragma unroll
            for (int e = 0; e < kHaloPerThread; e++) {
                const int h = tid + e * NTHREADS;
                hv[e] = h < nhalo ? __ldcg(g.edges + hsrc[h]) : 0.0;
            }
            // scipy's stopping test "||r|| < atol" (top of the loop).  p may already be advanced when we
            // leave: only x is returned.  Placed here so that the sqrt overlaps the loads above.
            if (sqrt(rr) < atol) { status = 0; break; }
            // ---- A2: own cells  p = p*beta + r
            if (PATCH) {
#pragma unroll
                for (int jy = 0; jy < (PATCH ? YPT : 0); jy++) {
                    if (jy < nval) {
#pragma unroll
                        for (int t = 0; t < NT; t++) {
                            const int si = fresh(sb) + t * plane + jy * PX;
                            ps[si] = ps[si] * beta + rj[t * YPT + jy];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    const int si = fresh(info[j]) & kSiMask;
                    ps[si] = ps[si] * beta + rj[j];
                }
            }
            // ---- A1 (second half): advance the halo copy of p
#pragma unroll
            for (int e = 0; e < kHaloPerThread; e++) {
                const int h = tid + e * NTHREADS;
                if (h < nhalo) { const int i = hdst[h]; ps[i] = ps[i] * beta + hv[e]; }
            }
            __syncthreads();
            lap(0);
            // ---- A3: q = A p (csr_matvec order), partial p.q
            acc = 0.0;
            if (PATCH) {
                if (nval > 0) {
                    // rolling window over the owned rows: up = row jy-1, cur = row jy, nxt = row jy+1, all NT levels
                    constexpr int NTP = PATCH ? NT : 1;
                    double up[NTP], cur[NTP], nxt[NTP];
                    const double *pb = ps + fresh(sb);
#pragma unroll
                    for (int t = 0; t < NTP; t++) { up[t] = pb[t * plane - PX]; cur[t] = pb[t * plane]; }
#pragma unroll
                    for (int jy = 0; jy < (PATCH ? YPT : 0); jy++) {
                        if (jy < nval) {
#pragma unroll
                            for (int t = 0; t < NTP; t++) nxt[t] = pb[t * plane + (jy + 1) * PX];
                            const int ym = (jy == jTop) + (jy == jBot);
#pragma unroll
                            for (int t = 0; t < NTP; t++) {
                                const double *px = pb + t * plane + jy * PX;
                                const bool tb = t == 0 || t == NTP - 1;
                                double dg = tb ? dg_tb : dg_ti;
                                if (ym) dg = dtab[(tb ? 5 : 6) - xmiss - ym - 3];
                                const double c = cur[t];
                                double s = 0.0;
                                if (UNIT) {
                                    if (t > 0) s -= cur[t - 1];
                                    s -= up[t]; s -= px[-1];
                                    s += dg * c;
                                    s -= px[1]; s -= nxt[t];
                                    if (t < NTP - 1) s -= cur[t + 1];
                                } else {
                                    if (t > 0) s += off * cur[t - 1]; else s += off * 0.0;
                                    s += off * up[t]; s += off * px[-1];
                                    s += dg * c;
                                    s += off * px[1]; s += off * nxt[t];
                                    if (t < NTP - 1) s += off * cur[t + 1]; else s += off * 0.0;
                                }
                                qj[t * YPT + jy] = s;
                                acc = fma(c, s, acc);
                            }
#pragma unroll
                            for (int t = 0; t < NTP; t++) { up[t] = cur[t]; cur[t] = nxt[t]; }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    const int inf = fresh(info[j]);
                    const double *pc = ps + (inf & kSiMask);
                    const int code = (inf >> kCntShift) & 3;
                    double dg = dg_interior;                 // 6 neighbours: no shared-memory lookup
                    if (code != 3) dg = dtab[code];
                    const double c = pc[0];
                    double s = 0.0;
                    if (UNIT) {                           // r == 1: products with -1.0 are exact negations
                        s -= pc[-plane]; s -= pc[-PX]; s -= pc[-1];
                        s += dg * c;
                        s -= pc[1]; s -= pc[PX]; s -= pc[plane];
                    } else {
                        s += off * pc[-plane]; s += off * pc[-PX]; s += off * pc[-1];
                        s += dg * c;
                        s += off * pc[1]; s += off * pc[PX]; s += off * pc[plane];
                    }
                    s = (inf & kValid) ? s : 0.0;
                    qj[j] = s;
                    acc = fma(c, s, acc);
                }
            }
            lap(1);
            const double pq = grid_sum(acc, false, nothing);     // nothing published since the last barrier
            lap(2);
            if (abort) break;
            const double alpha = rr / pq;
            // ---- B: r -= alpha q, export tile edges, partial r.r
            acc = 0.0;
            if (PATCH) {
#pragma unroll
                for (int jy = 0; jy < (PATCH ? YPT : 0); jy++) {
                    if (jy < nval) {
#pragma unroll
                        for (int t = 0; t < NT; t++) {
                            const int j = t * YPT + jy;
                            const double v = rj[j] - alpha * qj[j];
                            rj[j] = v;
                            acc = fma(v, v, acc);
                            if (expW) __stcg(qW + t * ty + jy, v);
                            if (expE) __stcg(qE + t * ty + jy, v);
                            if (jy == 0 && expN) __stcg(pN + t * tx, v);
                            if (jy == jS) __stcg(pS + t * tx, v);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    const double v = rj[j] - alpha * qj[j];
                    rj[j] = v;
                    acc = fma(v, v, acc);
                    const int inf = info[j];
                    if (inf & kEdgeAny) {                    // rare path; precomputed pointers keep it short
                        if (inf & kEdgeW) __stcg(pW + j * RPP, v);
                        if (inf & kEdgeE) __stcg(pE + j * RPP, v);
                        if (inf & (kEdgeN | kEdgeS)) {
                            const int o = ((inf >> kTShift) & kTMask) * tx;
                            if (inf & kEdgeN) __stcg(pN + o, v);
                            if (inf & kEdgeS) __stcg(pS + o, v);
                        }
                    }
                }
            }
            lap(3);
            // split-phase barrier: publish r.r, update x while the other CTAs arrive
            const double rr_new = grid_sum(acc, true, [&] {
                if (PATCH) {
#pragma unroll
                    for (int jy = 0; jy < (PATCH ? YPT : 0); jy++) {
                        if (jy < nval) {
#pragma unroll
                            for (int t = 0; t < NT; t++) {
                                const int xi = (t * YPT + jy) * NTHREADS + tid;
                                xs[xi] = xs[xi] + alpha * ps[fresh(sb) + t * plane + jy * PX];
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CPT; j++) {
                        const int xi = j * NTHREADS + tid;
                        xs[xi] = xs[xi] + alpha * ps[fresh(info[j]) & kSiMask];
                    }
                }
                lap(4);
            });
            lap(5);
            if (abort) break;
            rr_prev = rr; rr = rr_new;
        }
    }
    if (abort) { if (tid == 0) *a.sync.error = 1; return; }
    // ---- write phi
    if (PATCH) {
#pragma unroll
        for (int jy = 0; jy < (PATCH ? YPT : 0); jy++) {
            if (jy < nval) {
#pragma unroll
                for (int t = 0; t < NT; t++)
                    a.x[((size_t)t * Ny + (y0 + ly0 + jy)) * Nx + (x0 + lx)] = xs[(t * YPT + jy) * NTHREADS + tid];
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (info[j] & kValid) {
                const int row = j * RPP + r0;
                const int t = row / ty, ly = row - t * ty;
                a.x[((size_t)t * Ny + (y0 + ly)) * Nx + (x0 + lx)] = xs[j * NTHREADS + tid];
            }
        }
    }
    if (cta == 0 && tid == 0) { a.out[0] = it; a.out[1] = status; }
    if (prof) { for (int k = 0; k < 6; k++) g.prof[cta * 8 + k] += sprof[k]; g.prof[cta * 8 + 6] += it; }
}

__global__ void k_fill_u64(unsigned long long *p, int n, unsigned long long v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- configurations: threads per CTA x cell slots per thread (same per-SM capacity) ----------
struct Config { int threads, cpt; const void *unit, *general; int nt, ypt; };   // nt > 0: patch ownership, needs Nt == nt
const Config kConfigs[] = {
    // patch ownership for the CLI default Nt = 4
    {512, 16, (const void *)cg_onchip_kernel<512, 16, true, 4, 4>, (const void *)cg_onchip_kernel<512, 16, false, 4, 4>, 4, 4},
    // measured on B200 at 388x584x4: 8.9 / 9.8 / 10.6 us per CG iteration (ties go to the first)
    {512, 14, (const void *)cg_onchip_kernel<512, 14, true>, (const void *)cg_onchip_kernel<512, 14, false>, 0, 0},
    {1024, 7, (const void *)cg_onchip_kernel<1024, 7, true>, (const void *)cg_onchip_kernel<1024, 7, false>, 0, 0},
    {256, 28, (const void *)cg_onchip_kernel<256, 28, true>, (const void *)cg_onchip_kernel<256, 28, false>, 0, 0},
    // larger per-SM capacity (9 216 / 10 240 cells) for grids such as 480x640x4 that the first three cannot hold
    {512, 18, (const void *)cg_onchip_kernel<512, 18, true>, (const void *)cg_onchip_kernel<512, 18, false>, 0, 0},
    {256, 40, (const void *)cg_onchip_kernel<256, 40, true>, (const void *)cg_onchip_kernel<256, 40, false>, 0, 0},
};
constexpr int kNumConfigs = sizeof(kConfigs) / sizeof(kConfigs[0]);

struct Plan {
    bool ok = false;
    int cfg = 0, gy = 0, gx = 0, maxlen = 0, ncta = 0, passes = 0;
    size_t smem = 0;
};

int dev_init(OnchipScratch &d, int device)
{
    if (d.num_sms) return FOTO_OK;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    d.num_sms = prop.multiProcessorCount;
    d.smem_optin = prop.sharedMemPerBlockOptin;
    const char *e = getenv("FOTO_ONCHIP_CONFIG");
    d.forced_cfg = e ? atoi(e) : -1;
    if (d.forced_cfg >= kNumConfigs) d.forced_cfg = -1;
    return FOTO_OK;
}

// Choose configuration and tile grid: gy*gx <= #SMs, every tile fits (passes <= cell slots, shared
// memory, descriptor bit fields); minimise the number of passes of the busiest CTA, then the
// largest tile, then the edge length.
Plan make_plan(OnchipScratch &d, int device, int Nt, int Ny, int Nx)
{
    Plan best;
    if (dev_init(d, device) != FOTO_OK || Nt > kTMask) return best;
    long long best_key = -1;
    int force_gy = 0, force_gx = 0;                      // FOTO_ONCHIP_GRID=gy,gx: pin the tile grid (experiments)
    if (const char *e = getenv("FOTO_ONCHIP_GRID")) sscanf(e, "%d,%d", &force_gy, &force_gx);
    for (int c = 0; c < kNumConfigs; c++) {
        if (d.forced_cfg >= 0 && c != d.forced_cfg) continue;
        const int T = kConfigs[c].threads, CPT = kConfigs[c].cpt;
        if (kConfigs[c].nt > 0 && kConfigs[c].nt != Nt) continue;
        for (int gy = 1; gy <= d.num_sms && gy <= Ny; gy++) {
            const int gx_max = d.num_sms / gy;
            for (int gx = 1; gx <= gx_max && gx <= Nx; gx++) {
                if (force_gy > 0 && (gy != force_gy || gx != force_gx)) continue;
                const int ty = (Ny + gy - 1) / gy, tx = (Nx + gx - 1) / gx;      // largest tile
                const int ty_min = Ny / gy, tx_min = Nx / gx;                   // smallest tile
                if (tx > T || ty_min < 1 || tx_min < 1) continue;
                int passes = 0;                                                 // worst over the tile shapes that occur
                for (int txx = tx_min; txx <= tx; txx++)
                    for (int tyy = ty_min; tyy <= ty; tyy++) {
                        const int rpp = T / txx;
                        const int p = (Nt * tyy + rpp - 1) / rpp;
                        if (p > passes) passes = p;
                    }
                if (passes > CPT) continue;
                if ((long long)(Nt + 2) * (ty + 2) * (tx + 2) > kSiMask) continue;
                const size_t smem = ((((size_t)(Nt + 2) * (ty + 2) * (tx + 2) + 1) & ~size_t(1)) + (size_t)CPT * T + 64 + 4) * 8
                                  + (size_t)4 * Nt * (tx + ty) * sizeof(int);
                if (smem > d.smem_optin) continue;
                if (2LL * Nt * (tx + ty) > (long long)kHaloPerThread * T) continue;
                // work per SM ~ passes * threads (issue slots); then prefer wide tiles: a warp that
                // spans several tile rows takes the W/E edge-export path in every cell and has
                // shared-memory bank conflicts at the row breaks; then short halos
                long long edge_warps = 32000 / tx + 2000 / ty;            // ~ per-mille of warps on the edge path
                long long work = (long long)passes * T;
                if (kConfigs[c].nt > 0) {
                    // patch ownership: an active thread always works on all of its cell slots, so every grid that
                    // fits costs about the same; measured at 388x584x4 (tools/sweep_grid.py): 7.9-8.1 us for tiles
                    // wider than tall with ty >= 11, 8.3-8.8 us for very flat (ty = 8) or narrow (tx < 40) tiles; idle SMs cost a little
                    work = 0;
                    edge_warps = 32000 / tx + 8000 / ty + 3 * (d.num_sms - gy * gx);
                }
                const long long key = (work * 10000 + edge_warps) * 100000 + (long long)(tx + ty);
                if (best_key < 0 || key < best_key) {
                    best_key = key;
                    best.ok = true; best.cfg = c; best.gy = gy; best.gx = gx; best.ncta = gy * gx;
                    best.maxlen = tx > ty ? tx : ty; best.smem = smem; best.passes = passes;
                }
            }
        }
        if (best.ok) break;        // configurations are listed fastest first: take the first that fits
    }
    return best;
}

}  // namespace

bool cg_onchip_fits(OnchipScratch &s, int device, int Nt, int Ny, int Nx)
{
    return make_plan(s, device, Nt, Ny, Nx).ok;
}

void cg_onchip_release(OnchipScratch &s)
{
    cudaFree(s.edges); cudaFree(s.slots); cudaFree(s.prof); cudaFree(s.gn_edges); cudaFree(s.gn_slots); cudaFree(s.fused_edges); cudaFree(s.fused_slots); cudaFree(s.gnf_edges); cudaFree(s.gnf_slots);
    s.gnf_edges = nullptr; s.gnf_slots = nullptr; s.gnf_edges_bytes = 0;
    s.fused_edges = nullptr; s.fused_slots = nullptr; s.fused_edges_bytes = 0;
    s.gn_edges = nullptr; s.gn_slots = nullptr; s.gn_edges_bytes = 0;
    s.prof = nullptr;
    s.edges = nullptr; s.slots = nullptr; s.edges_bytes = 0;
}

int launch_cg_onchip(cudaStream_t st, const CgArgs &a, int device, OnchipScratch &d)
{
    Plan p = make_plan(d, device, a.Nt, a.Ny, a.Nx);
    if (!p.ok) { set_error("grid %dx%dx%d does not fit the on-chip CG variant", a.Nt, a.Ny, a.Nx); return FOTO_ERR_ARG; }
    const size_t need = (size_t)p.ncta * 4 * a.Nt * p.maxlen * sizeof(double);
    if (d.edges_bytes < need) {
        if (d.edges) CUDA_TRY(cudaFree(d.edges));
        CUDA_TRY(cudaMalloc((void **)&d.edges, need));
        d.edges_bytes = need;
    }
    if (!d.slots) CUDA_TRY(cudaMalloc((void **)&d.slots, kSlotWords * sizeof(unsigned long long)));
    if (!d.attr_set) {
        for (int c = 0; c < kNumConfigs; c++) {
            CUDA_TRY(cudaFuncSetAttribute(kConfigs[c].unit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_optin));
            CUDA_TRY(cudaFuncSetAttribute(kConfigs[c].general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_optin));
        }
        d.attr_set = true;
    }
    const void *fn = a.rcoef == 1.0 ? kConfigs[p.cfg].unit : kConfigs[p.cfg].general;
    k_fill_u64<<<(kSlotWords + 255) / 256, 256, 0, st>>>(d.slots, kSlotWords, kSentinel);
    OnchipGeom g;
    g.gy = p.gy; g.gx = p.gx; g.maxlen = p.maxlen; g.edges = d.edges; g.slots = d.slots; g.prof = d.prof;
    void *args[] = {(void *)&a, (void *)&g};
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(p.ncta), dim3(kConfigs[p.cfg].threads), args, p.smem, st));
    return FOTO_OK;
}

}  // namespace foto
