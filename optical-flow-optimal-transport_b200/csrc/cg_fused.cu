// cg_fused.cu -- K2a, on-chip resident CG with ONE grid all-reduce per iteration and NOTHING ELSE on the critical path
// between two stencils but the all-reduce and the vector update.
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
// the default and the streaming textbook kernel beside it.
//
// The stencil acts on r, so r lives in shared memory with a one-cell halo ring; p and s live in registers, x and w in
// private shared-memory slots.  Tile-edge exchange (round 2, second form): what crosses L2 is the tile edge of
// w = A r, NOT of the new r.  w is known before the all-reduce, so its L2 hop (~1 000-1 400 cycles store -> visible)
// runs in the shadow of the all-reduce instead of after the update (the first form exported r after the update and
// spun ~1 000 cycles per iteration at the top of the next one).  Each CTA keeps, for the halo ring, its own copies of
// r (in the ring of rs) and s (registers) and advances them with the owner's arithmetic -- s = s beta + w,
// r = r - alpha s, separately rounded products and sums, the same alpha and beta -- so both copies stay bit-identical.
// The ring of r_0 = b is read from global memory at setup.
// No barrier and no fence order the exchange: every exported word carries a generation tag in the least significant
// mantissa bit (the owner writes the same rounded w back into its slot, so both sides use one value; the perturbation
// is one ulp of an edge value of w per iteration, the size of an ordinary rounding error) and the reader spins on
// each word until the tag is the one it expects.  Edge words are double buffered on the iteration parity, tag =
// bit 1 of the iteration number: a CTA overwrites buffer (k & 1) with generation k+2 only after the all-reduce of
// iteration k+1, which every neighbour enters after it has consumed generation k.
// (Measured alternatives: flag + release fence hand-off 2 900 cycles per iteration, as much as the grid barrier it
// replaces; sentinel reset + triple buffering doubles the stores and costs 3 000 cycles in the reset loop; exporting
// from registers inside the unrolled update costs the edge warps 1 500 cycles.)
// Shadow of the all-reduce of iteration k: export of w_k's edges, the whole x += alpha_{k-1} p_{k-1}, import of the
// neighbours' edges into registers.
#include "foto_kernels.cuh"
#include "grid_sync.cuh"

namespace foto {

namespace {

using namespace gsync;

struct Geom {
    int gy, gx, maxlen;
    double *edges;                 // [2 (iteration parity)][ncta][4 (N,S,W,E)][NT * maxlen] tile-edge values of w, LSB = tag
    unsigned long long *slots;     // all-reduce words: kGranules 2 KB granules + the placement record (see place_allreduce)
    unsigned int launch_seq;       // distinguishes the placement broadcast of this launch from the previous one's
    unsigned int force_choice;     // experiments: granules given by FOTO_AR_PLACE=a,b,t (bit 31 set), bit 30: print the classification
    long long *prof;
};

constexpr int kHaloPerThread = 4;

// x[k] += v at the L2 (one adder per address and iteration: deterministic; IEEE addition, i.e. the value a load, an add
// and a store would give, without the round trip and with half the traffic)
__device__ __forceinline__ void red_add_f64(double *p, double v)
{
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// ---- where the all-reduce words live.  B200 is two dies of 74 SMs with the L2 split between them, and the home of an
// address changes every 2 KB.  A word that one SM stores and another polls becomes visible in ~480 cycles when its home is
// on the die of both SMs and in ~900 when it is not (tools/ubench_die.cu); for the all-reduce that is 4.75 against up to
// 5.27 us per CG iteration depending on where cudaMalloc happened to put the buffer (profiles/r2_allreduce_placement.md:
// best with the partial slots AND the totals on the root's die).  So the buffer is kGranules 2 KB granules, the root
// classifies them once by a self ping-pong (store a value, poll it back: ~300 cycles on its own die, ~720 on the other),
// keeps the partial slots and the totals in granules of its own die (plus a second copy of the totals in a granule of the
// other die: a waiter polls both and takes whichever arrives first, 0.8 % faster), remembers the choice per SM id, and
// tells the other CTAs at the start of every launch.
constexpr int kGranules = 32, kGranWords = 256, kSlotsPerGran = 128;       // 16-byte slots
constexpr int kPlaceOff = kGranules * kGranWords;       // words: [0] choice of this launch, [1] cached choice, [2] its SM id + 1
constexpr int kFusedSlotWords = kPlaceOff + 16;

struct Placement { unsigned long long *part_a, *part_b, *tot, *tot2; };   // CTAs 0..127 / 128..255 / totals / copy on the other die

__device__ __forceinline__ Placement placement_of(unsigned long long *base, unsigned long long choice)
{
    Placement p;
    p.part_a = base + (choice & 255) * kGranWords;
    p.part_b = base + ((choice >> 8) & 255) * kGranWords;
    p.tot = base + ((choice >> 16) & 255) * kGranWords;
    p.tot2 = base + ((choice >> 24) & 255) * kGranWords;
    return p;
}

// thread 0 of every CTA, before the first all-reduce
__device__ Placement place_allreduce(unsigned long long *base, unsigned int seq, bool root, bool *timed_out, unsigned int force)
{
    unsigned long long *rec = base + kPlaceOff;
    unsigned long long choice;
    if (root) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        choice = ld_relaxed_u64(rec + 1);
        if (ld_relaxed_u64(rec + 2) != (unsigned long long)smid + 1 || (force & 0x40000000u)) {
            int t[kGranules], lo = 1 << 30, hi = 0;
            for (int g = 0; g < kGranules; g++) {
                unsigned long long *w = base + g * kGranWords + 2 * kSlotsPerGran - 2;      // last slot of the granule: never used below 128 CTAs
                int m = 1 << 30;
                for (int rep = 0; rep < 5; rep++) {
                    const unsigned long long v = 0xC0DE0000ull + (unsigned long long)rep * 2 + 1;        // odd: reads as "not written" afterwards
                    const long long c0 = clock64();
                    st_relaxed_u64(w, v);
                    while (ld_relaxed_u64(w) != v && clock64() - c0 < 1000000) { }
                    const int d = (int)(clock64() - c0);
                    m = d < m ? d : m;
                }
                st_relaxed_u64(w, ~0ull);
                t[g] = m; lo = m < lo ? m : lo; hi = m > hi ? m : hi;
            }
            if (force & 0x40000000u) {
                printf("root smid %u, self ping-pong cycles per granule:", smid);
                for (int g = 0; g < kGranules; g++) printf(" %d", t[g]);
                printf("\n");
            }
            int pick[3] = {0, 1, 2}, n = 0, far = -1;
            for (int g = 0; g < kGranules; g++) {
                if (2 * t[g] <= lo + hi) { if (n < 3) pick[n++] = g; }
                else if (far < 0) far = g;
            }
            if (far < 0) far = pick[2];
            choice = (unsigned long long)pick[0] | ((unsigned long long)pick[1] << 8) | ((unsigned long long)pick[2] << 16) |
                     ((unsigned long long)far << 24);
            st_relaxed_u64(rec + 1, choice);
            st_relaxed_u64(rec + 2, (unsigned long long)smid + 1);
        }
        if (force & 0x80000000u) choice = (force & 0xFFFFFFu) | ((unsigned long long)(force & 0xFF0000u) << 8);   // forced: one copy of the totals
        if (force & 0x20000000u) choice = (choice & 0xFFFFFFull) | ((choice & 0xFF0000ull) << 8);      // one copy of the totals only
        choice = (choice & 0xFFFFFFFFull) | ((unsigned long long)seq << 32);
        st_relaxed_u64(rec, choice);
    } else {
        const long long c0 = clock64();
        do { choice = ld_relaxed_u64(rec); } while ((unsigned int)(choice >> 32) != seq && clock64() - c0 < kWatchdogCycles);
        if ((unsigned int)(choice >> 32) != seq) { *timed_out = true; choice = 0x02020100ull; }      // the root never showed up
    }
    return placement_of(base, choice);
}

// the all-reduce of grid_sync.cuh (two values, tagged words) with explicit addresses
__device__ __forceinline__ void arrive2(unsigned long long *slot, unsigned int gen, const double *v, bool abort = false)
{
    const unsigned long long par = (unsigned long long)(gen & 1u);
    st_relaxed_v2(slot, abort ? (kAbort | par) : tagged(v[0], gen), abort ? (kAbort | par) : tagged(v[1], gen));
}
__device__ __forceinline__ void root2(const Placement &pl, unsigned int gen, int ncta, int lane)
{
    constexpr int PER_LANE = 5;
    const long long t0 = clock64();
    const unsigned long long par = (unsigned long long)(gen & 1u);
    bool abort = false;
    unsigned long long w[PER_LANE][2];
    bool ready;
    do {
#pragma unroll
        for (int k = 0; k < PER_LANE; k++) {
            int b = k * 32 + lane; if (b >= ncta) b = 0;
            ld_relaxed_v2((b < kSlotsPerGran ? pl.part_a : pl.part_b) + 2 * (b & (kSlotsPerGran - 1)), w[k][0], w[k][1]);
        }
        ready = true;
#pragma unroll
        for (int k = 0; k < PER_LANE; k++) ready = ready & ((w[k][0] & 1ull) == par) & ((w[k][1] & 1ull) == par);
        if (!ready && clock64() - t0 > kWatchdogCycles) { abort = true; break; }
    } while (!ready);
    double tot[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < PER_LANE; k++)
        if (k * 32 + lane < ncta) {
            tot[0] += __longlong_as_double((long long)w[k][0]); tot[1] += __longlong_as_double((long long)w[k][1]);
            abort = abort || is_abort(w[k][0]) || is_abort(w[k][1]);
        }
    tot[0] = warp_sum(tot[0]); tot[1] = warp_sum(tot[1]);
    abort = __any_sync(0xffffffffu, abort);
    if (lane < 2) st_relaxed_v2(lane == 0 ? pl.tot : pl.tot2, abort ? (kAbort | par) : tagged(tot[0], gen), abort ? (kAbort | par) : tagged(tot[1], gen));
}
__device__ __forceinline__ bool wait2(const unsigned long long *tot, const unsigned long long *tot2, unsigned int gen, double *out)
{
    const long long t0 = clock64();
    const unsigned long long par = (unsigned long long)(gen & 1u);
    unsigned long long a, b, c, d;
    bool ready, ok = true;
    do {                                                 // both copies in flight; whichever shows the new generation first
        ld_relaxed_v2(tot, a, b);
        ld_relaxed_v2(tot2, c, d);
        ready = ((a & 1ull) == par) & ((b & 1ull) == par);
        if (!ready && ((c & 1ull) == par) & ((d & 1ull) == par)) { a = c; b = d; ready = true; }
        if (!ready && clock64() - t0 > 2 * kWatchdogCycles) { ok = false; break; }
    } while (!ready);
    out[0] = __longlong_as_double((long long)a); out[1] = __longlong_as_double((long long)b);
    return ok && !is_abort(a) && !is_abort(b);
}

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
    // layout: scratch and the small tables first (constant offsets), then r, x, w, then the per-entry halo state
    double *red = smem;                                 // reduction scratch: 64 block_sum, 64..67 totals / flags, 72..77 profile
    double *dtab = red + 80;                            // diagonal entries for 3..6 neighbours
    double *rs = dtab + 4;                              // [NT][PY][PX]  r, halo ring (zero outside the domain)
    double *xs = rs + psz;                              // [CPT][NTHREADS] x
    double *ws = xs + (XG ? 0 : CPT * NTHREADS);        // [CPT][NTHREADS] w = A r
    double *shs = ws + CPT * NTHREADS;                  // s on the halo-ring cells (copy of the neighbours' s), by entry
    // one table entry per exchanged edge cell (import and export lists have the same sides and lengths):
    //   .x offset of the neighbour's word in a parity buffer of g.edges   .y ring cell of rs it feeds
    //   .z slot of ws (of the owning thread) that is exported              .w offset of the exported word in my_edges
    int4 *htab = (int4 *)(shs + 2 * NT * (tx + ty));
    const bool hasN = by > 0, hasS = by < g.gy - 1, hasW = bx > 0, hasE = bx < g.gx - 1;
    const double off = -a.rcoef * 1.0;
    const int edge_stride = NT * g.maxlen;
    const size_t ebuf = (size_t)ncta * 4 * edge_stride;           // one parity buffer
    double *my_edges = g.edges + (size_t)cta * 4 * edge_stride;
    const int lx = tid % tx, r0 = tid / tx, RPP = NTHREADS / tx;
    // warp 0 of every CTA owns the all-reduce (block total, arrive, root duty in CTA 0, poll of the totals) and takes no
    // export / import entries, so that its poll starts long before the totals can arrive
    constexpr int eoff = 32, estride = NTHREADS - eoff;
    const int etid = tid - eoff;

    // ---- setup
    for (int i = tid; i < psz; i += NTHREADS) rs[i] = 0.0;
    if (tid < 4) dtab[tid] = -a.rcoef * (-(double)(tid + 3)) + a.rcoef * a.eps * 1.0;    // -r*L_ii + r*eps
    __syncthreads();                                     // rs zeroed before the ring and the owners fill it
    // ws slot of the tile cell (t, ly, lx): owner thread (ly / YPT) * tx + lx, slot t * YPT + ly % YPT
    auto wslot = [&](int t, int ly, int lxx) { return (t * YPT + ly % YPT) * NTHREADS + (ly / YPT) * tx + lxx; };
    int nhalo = 0;
    {
        const int segNS = NT * tx, segWE = NT * ty;
        // per side: e -> (t, pos); import from the neighbour's facing edge list, ring cell of rs, ring value of r_0 = b
        // from global memory; export of the own edge row / column
        for (int e = tid; e < segNS; e += NTHREADS) {
            const int t = e / tx, pos = e - t * tx;
            if (hasN) {
                htab[nhalo + e] = make_int4(((cta - g.gx) * 4 + 1) * edge_stride + e, (t * PY) * PX + pos + 1, wslot(t, 0, pos), 0 * edge_stride + e);
                rs[(t * PY) * PX + pos + 1] = a.b[((size_t)t * Ny + (y0 - 1)) * Nx + (x0 + pos)];
            }
        }
        if (hasN) nhalo += segNS;
        for (int e = tid; e < segNS; e += NTHREADS) {
            const int t = e / tx, pos = e - t * tx;
            if (hasS) {
                htab[nhalo + e] = make_int4(((cta + g.gx) * 4 + 0) * edge_stride + e, (t * PY + ty + 1) * PX + pos + 1, wslot(t, ty - 1, pos), 1 * edge_stride + e);
                rs[(t * PY + ty + 1) * PX + pos + 1] = a.b[((size_t)t * Ny + y1) * Nx + (x0 + pos)];
            }
        }
        if (hasS) nhalo += segNS;
        for (int e = tid; e < segWE; e += NTHREADS) {
            const int t = e / ty, pos = e - t * ty;
            if (hasW) {
                htab[nhalo + e] = make_int4(((cta - 1) * 4 + 3) * edge_stride + e, (t * PY + pos + 1) * PX, wslot(t, pos, 0), 2 * edge_stride + e);
                rs[(t * PY + pos + 1) * PX] = a.b[((size_t)t * Ny + (y0 + pos)) * Nx + (x0 - 1)];
            }
        }
        if (hasW) nhalo += segWE;
        for (int e = tid; e < segWE; e += NTHREADS) {
            const int t = e / ty, pos = e - t * ty;
            if (hasE) {
                htab[nhalo + e] = make_int4(((cta + 1) * 4 + 2) * edge_stride + e, (t * PY + pos + 1) * PX + tx + 1, wslot(t, pos, tx - 1), 3 * edge_stride + e);
                rs[(t * PY + pos + 1) * PX + tx + 1] = a.b[((size_t)t * Ny + (y0 + pos)) * Nx + x1];
            }
        }
        if (hasE) nhalo += segWE;
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
    for (int i = tid; i < 2 * NT * (tx + ty); i += NTHREADS) shs[i] = 0.0;
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
    // export pass (after a __syncthreads that follows the writes of ws): the tile-edge values of w of iteration k are
    // rounded to the tag in place (so the owner and the neighbour use the same value) and stored into parity buffer
    // k & 1; a corner cell sits in two lists and is rounded twice to the same value
    auto export_edges = [&](unsigned int k) {
        const long long tag = (long long)((k >> 1) & 1u);
        double *dst = my_edges + (size_t)(k & 1u) * ebuf;
        if (etid >= 0) {
#pragma unroll
            for (int e = 0; e < kHaloPerThread; e++) {
                const int h = etid + e * estride;
                if (h < nhalo) {
                    const int4 te = htab[h];
                    const int si = te.z;
                    const double v = __longlong_as_double((__double_as_longlong(ws[si]) & ~1ll) | tag);
                    ws[si] = v;
                    st_relaxed_u64((unsigned long long *)(dst + te.w), (unsigned long long)__double_as_longlong(v));
                }
            }
        }
    };

    // x slot j += c * p_j; XG: reduction into the owned cell in global memory (L2)
    double *const xg = a.x + ((size_t)(y0 + ly0)) * Nx + (x0 + lx);
    const size_t Pst = (size_t)Ny * Nx;
    auto x_update = [&](double c) {
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            if (XG) {
                const int t = j / YPT, jy = j - t * YPT;
                if (jy < nval) red_add_f64(xg + t * Pst + (size_t)jy * Nx, c * pj[j]);
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
    if (tid == 0) { for (int k = 0; k < 6; k++) sprof[k] = 0; red[66] = 0.0; red[67] = 0.0; }
    // all-reduce addresses (only warp 0 uses them): red[68..71] = partial granules a, b, totals, this CTA's slot; red[78] totals copy
    unsigned long long **arp = (unsigned long long **)(red + 68);
    if (tid == 0) {
        bool timed_out = false;
        const Placement pl = place_allreduce(g.slots, g.launch_seq, cta == 0, &timed_out, g.force_choice);
        if (timed_out) red[66] = 1.0;
        arp[0] = pl.part_a; arp[1] = pl.part_b; arp[2] = pl.tot; arp[10] = pl.tot2;
        arp[3] = (cta < kSlotsPerGran ? pl.part_a : pl.part_b) + 2 * (cta & (kSlotsPerGran - 1));
    }
    auto lap = [&](int k) { if (prof) { long long now = clock64(); sprof[k] += now - tmark; tmark = now; } };

    int it = 0, status = a.maxiter;
    double gam_stop = 0.0, rgam_prev = 0.0, d_prev = 0.0, alpha_prev = 0.0;
    bool pend = false;                                   // x += alpha_prev p not yet applied
    __syncthreads();                                     // rs (interior and ring) complete
    if (prof) tmark = clock64();
    for (; it < a.maxiter; it++) {
        // ---- w = A r (csr_matvec order), partial r.r and r.w
        double acc[2] = {0.0, 0.0};
        // (a branch-free copy of this block for warps whose threads all own YPT rows saves 48 register moves per thread and
        // 160 in-kernel cycles per iteration under the phase counters, but the plain build is 1 % slower with it: A/B on one box)
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
                    double dgb = dg_tb, dgi = dg_ti;
                    if (ym) { dgb = dtab[2 - xmiss - ym]; dgi = dtab[3 - xmiss - ym]; }      // rows on the global y boundary only
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        const double *px = pb + t * plane + jy * PX;
                        const bool tb = t == 0 || t == NT - 1;
                        const double dg = tb ? dgb : dgi;
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
        lap(0);
        // ---- the one all-reduce; in its shadow: export of the edges of w, x update of the previous iteration,
        //      import of the neighbours' edges of w (into registers)
        {                                                // block total in warp 0 only (block_sum of common.cuh without the broadcast)
            acc[0] = warp_sum(acc[0]); acc[1] = warp_sum(acc[1]);
            if ((tid & 31) == 0) { red[tid >> 5] = acc[0]; red[32 + (tid >> 5)] = acc[1]; }
            __syncthreads();                             // also orders the writes of ws before the export
            if (tid < 32) {
                acc[0] = warp_sum(tid < NTHREADS / 32 ? red[tid] : 0.0);
                acc[1] = warp_sum(tid < NTHREADS / 32 ? red[32 + tid] : 0.0);
                if (tid == 0) arrive2(arp[3], gen, acc);
                if (cta == 0) root2(Placement{arp[0], arp[1], arp[2], arp[10]}, gen, ncta, tid);
            }
        }
        export_edges((unsigned int)it);                  // first: the neighbours' imports wait for these words (x update first: 4.92
                                                         // against 4.69 us; the root CTA taking the totals from its own root warp
                                                         // instead of polling them: 4.75, same-box A/B)
        if (pend) { x_update(alpha_prev); pend = false; }
        double hv[kHaloPerThread] = {0.0, 0.0, 0.0, 0.0};
        auto import_edges = [&]() {
            if (etid >= 0) {
                const double *eg = g.edges + (size_t)(it & 1) * ebuf;
                const unsigned long long tag = (unsigned long long)((it >> 1) & 1);
                const long long t0 = clock64();
                int rounds = 0;
                bool ready;
                do {
                    ready = true;
#pragma unroll
                    for (int e = 0; e < kHaloPerThread; e++) {
                        const int h = etid + e * estride;
                        hv[e] = 0.0;
                        if (h < nhalo) {
                            const unsigned long long bits = ld_relaxed_u64((const unsigned long long *)(eg + htab[h].x));
                            ready = ready && (bits & 1ull) == tag;
                            hv[e] = __longlong_as_double((long long)bits);
                        }
                    }
                    rounds++;
                    if (!ready && clock64() - t0 > kWatchdogCycles) { red[66] = 1.0; break; }     // a neighbour is stuck: abort below
                } while (!ready);
                if (g.prof != nullptr && tid == 32) { sprof[4] += rounds; sprof[5] += clock64() - t0; }
            }
        };
        import_edges();
        lap(1);
        if (tid == 0) {
            if (!wait2(arp[2], arp[10], gen, red + 64)) red[67] = 1.0;
        }
        __syncthreads();
        gen++;
        const double gam = red[64], del = red[65];
        abort = red[67] != 0.0;
        lap(2);
        if (abort) break;
        if (red[66] != 0.0) {                            // import watchdog of this CTA: one more round that tells every CTA
            __syncthreads();
            if (tid == 0) arrive2(arp[3], gen, acc, true);
            if (cta == 0 && tid < 32) root2(Placement{arp[0], arp[1], arp[2], arp[10]}, gen, ncta, tid);
            if (tid == 0) wait2(arp[2], arp[10], gen, red + 64);
            abort = true;
            break;
        }
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
        // ---- p = r + beta p, s = w + beta s, r -= alpha s on the owned cells ...
        // (a branch-free copy of this loop for warps whose threads all own YPT rows, as in the stencil, is slower: 4.83
        // against 4.72 us per iteration, A/B on one box)
        {
            double *pr[NT];
#pragma unroll
            for (int t = 0; t < NT; t++) pr[t] = rs + fresh(sb) + t * plane;
            const double *wp = ws + tid;
#pragma unroll
            for (int jy = 0; jy < YPT; jy++) {
                if (jy < nval) {
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        const int j = t * YPT + jy;
                        const double rv = pr[t][0], wv = wp[j * NTHREADS];
                        const double sv = sj[j] * beta + wv;
                        pj[j] = pj[j] * beta + rv;
                        sj[j] = sv;
                        pr[t][0] = rv - alpha * sv;
                        pr[t] += PX;
                    }
                }
            }
        }
        // ... and, with the same arithmetic, s and r on the halo ring (the neighbours' edge cells)
        if (etid >= 0) {
#pragma unroll
            for (int e = 0; e < kHaloPerThread; e++) {
                const int h = etid + e * estride;
                if (h < nhalo) {
                    const int di = htab[h].y;
                    const double sv = shs[h] * beta + hv[e];
                    shs[h] = sv;
                    rs[di] = rs[di] - alpha * sv;
                }
            }
        }
        __syncthreads();
        lap(3);
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
                if (XG) {
                    if (pend) red_add_f64(px, alpha_prev * pj[j]);
                } else {
                    double xv = xs[j * NTHREADS + tid];
                    if (pend) xv = xv + alpha_prev * pj[j];
                    *px = xv;
                }
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
// listed fastest first per Nt.  Nt = 4 at 388x584 (us per iteration, this round's kernel): 448 threads x 16 slots 4.76, 384 x 20
// (168 registers, no spills) 4.78, 512 x 16 4.89, 320 x 24 5.11.  448 threads (14 warps, 7 168 cell slots), 512
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
// 388x584x4, SWEEP_VARIANT=2: 5.87 us per iteration for 12x12, 5.96 for 8x18, 6.2-6.6 for the rest): short halos and few idle SMs first, then wide rows.  (Padding the rows of r so that a half-warp running off a thread row
// continues on the following banks -- conflict free for any tile width -- was measured: no gain, 5.00 against 4.97 us.)
Plan make_plan(OnchipScratch &d, int device, int Nt, int Ny, int Nx)
{
    Plan best;
    if (!d.num_sms) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return best;
        d.num_sms = prop.multiProcessorCount; d.smem_optin = prop.sharedMemPerBlockOptin;
    }
    int force_shape = -1;                                // FOTO_FUSED_SHAPE=index into kShapes (experiments)
    if (const char *e = getenv("FOTO_FUSED_SHAPE")) force_shape = atoi(e);
    for (int shape = 0; shape < kNumShapes && !best.ok; shape++) {
    if (kShapes[shape].nt != Nt || (force_shape >= 0 && shape != force_shape)) continue;
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
            if (2LL * kNT * (tx + ty) > (long long)kHaloPerThread * (kThreads - 32)) continue;
            const size_t smem = ((((size_t)kNT * (ty + 2) * (tx + 2) + 1) & ~size_t(1)) + (size_t)xslots * kNT * kYPT * kThreads + 80 + 4
                                 + (size_t)2 * kNT * (tx + ty)) * 8 + (size_t)8 * kNT * (tx + ty) * sizeof(int);
            if (smem > d.smem_optin) continue;
            const long long key = ((2LL * kNT * (tx + ty) + 8LL * (d.num_sms - gy * gx)) * 1000) + (999 - tx);
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
    cudaFree(s.prof); cudaFree(s.fused_edges); cudaFree(s.fused_slots_raw); cudaFree(s.gnf_edges); cudaFree(s.gnf_slots);
    s.prof = nullptr;
    s.fused_edges = nullptr; s.fused_slots = nullptr; s.fused_slots_raw = nullptr; s.fused_edges_bytes = 0;
    s.gnf_edges = nullptr; s.gnf_slots = nullptr; s.gnf_edges_bytes = 0;
}

int launch_cg_fused(cudaStream_t st, const CgArgs &a, int device, OnchipScratch &d)
{
    Plan p = make_plan(d, device, a.Nt, a.Ny, a.Nx);
    if (!p.ok) { set_error("grid %dx%dx%d does not fit the single-reduction on-chip CG variant", a.Nt, a.Ny, a.Nx); return FOTO_ERR_ARG; }
    const size_t need = (size_t)2 * p.ncta * 4 * a.Nt * p.maxlen * sizeof(double);     // two parity buffers
    if (d.fused_edges_bytes < need) {
        if (d.fused_edges) CUDA_TRY(cudaFree(d.fused_edges));
        CUDA_TRY(cudaMalloc((void **)&d.fused_edges, need));
        d.fused_edges_bytes = need;
    }
    if (!d.fused_slots) {                                // 2 KB aligned granules + the placement record (zeroed once: "nothing cached")
        CUDA_TRY(cudaMalloc((void **)&d.fused_slots_raw, (kFusedSlotWords + kGranWords) * sizeof(unsigned long long)));
        d.fused_slots = (unsigned long long *)(((size_t)d.fused_slots_raw + 2047) & ~(size_t)2047);
        CUDA_TRY(cudaMemsetAsync(d.fused_slots + kPlaceOff, 0, 16 * sizeof(unsigned long long), st));
    }
    const void *fn = a.rcoef == 1.0 ? kShapes[p.shape].unit : kShapes[p.shape].general;
    if (!d.fused_attr_set) {
        for (int i = 0; i < kNumShapes; i++) {
            CUDA_TRY(cudaFuncSetAttribute(kShapes[i].unit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_optin));
            CUDA_TRY(cudaFuncSetAttribute(kShapes[i].general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_optin));
        }
        d.fused_attr_set = true;
    }
    const int nedge = (int)(need / sizeof(double));
    // all-reduce slots and edge words <- tag 1 ("generations 0 and 1 not yet written")
    k_fill2_u64<<<((nedge > kPlaceOff ? nedge : kPlaceOff) + 255) / 256, 256, 0, st>>>(d.fused_slots, kPlaceOff, kSlotInit, (unsigned long long *)d.fused_edges, nedge, ~0ull);
    Geom g;
    g.gy = p.gy; g.gx = p.gx; g.maxlen = p.maxlen; g.edges = d.fused_edges; g.slots = d.fused_slots; g.prof = d.prof;
    g.launch_seq = ++d.fused_launch_seq;
    g.force_choice = 0;
    if (const char *e = getenv("FOTO_AR_PLACE")) {
        int ga = 0, gb = 1, gt = 2;
        if (sscanf(e, "%d,%d,%d", &ga, &gb, &gt) == 3) g.force_choice = 0x80000000u | (ga & 31) | ((gb & 31) << 8) | ((gt & 31) << 16);
    }
    if (getenv("FOTO_AR_DEBUG") && d.fused_launch_seq == 1) g.force_choice |= 0x40000000u;
    if (getenv("FOTO_AR_ONECOPY")) g.force_choice |= 0x20000000u;
    void *args[] = {(void *)&a, (void *)&g};
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(p.ncta), dim3(kShapes[p.shape].threads), args, p.smem, st));
    return FOTO_OK;
}

}  // namespace foto
