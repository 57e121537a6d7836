// grid_sync.cuh -- grid-wide all-reduce of NV <= 4 doubles for the persistent on-chip kernels (one CTA per SM).
//
// Two-hop root gather through L2: thread 0 of every CTA stores its partial sums, warp 0 of CTA 0 sums the partials
// in a fixed order and stores the totals, thread 0 of every CTA polls the totals.  Every CTA gets bit-identical
// totals.  2 260-2 750 cycles for 144-148 CTAs and two values (tools/ubench_allreduce2.cu, variant A); one-hop
// all-gathers, thread-block clusters with DSMEM pre-reduction, integer-atomic accumulation and several staggered polls
// in flight per waiter were measured slower or equal (profiles/r2_allreduce_variants.md): the cost is two L2 traversals, partly across the two dies.
//
// Memory model.  Every communicated word validates itself: bit 0 of the mantissa carries the parity of the
// generation the word belongs to (a one-ulp rounding of a partial sum / total, the same value for every reader), and
// every access to such a word is a single 8-byte (or 2 x 8-byte vector) relaxed.gpu access.  A consumer uses only the
// bits it loaded, so no ordering between different addresses is needed: no sentinel reset, no fence, no flag.
// One buffer per word suffices: a CTA overwrites its generation-g partial only after it has read the total of
// generation g, which the root stores after it has read that partial; the root overwrites the total of generation g
// only after every CTA's partial of generation g+1 has arrived, i.e. after every CTA has read total g.  Each
// overwrite therefore depends, through values, on the load it must not overtake.
//   layout (unsigned long long words): partials [kMaxBlocks][4], then totals [4]
// The tile-edge words of cg_fused.cu / gn_fused.cu follow the same rule (LSB = parity of their generation).
#pragma once
#include "common.cuh"

namespace foto {
namespace gsync {

constexpr unsigned long long kAbort = 0x7FF8DEADBEEF0002ull;     // bit 0 is the tag: compared with bit 0 masked
constexpr unsigned long long kPlainNaN = 0x7FF8000000000000ull;
constexpr int kNV = 4;                                   // words per CTA slot (32 bytes)
constexpr int kTotOff = kNV * kMaxBlocks;
constexpr int kSlotWords = kTotOff + 16;
constexpr unsigned long long kSlotInit = ~0ull;          // tag 1: "generation 0 not yet written"
constexpr long long kWatchdogCycles = 8000000000ll;

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
__device__ __forceinline__ void st_relaxed_v2(unsigned long long *p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_relaxed_v2(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// bits of v with the generation parity in bit 0 (a NaN that happens to look like the abort code is replaced)
__device__ __forceinline__ unsigned long long tagged(double v, unsigned int gen)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v) & ~1ull;
    if (b == kAbort) b = kPlainNaN;
    return b | (unsigned long long)(gen & 1u);
}
__device__ __forceinline__ bool is_abort(unsigned long long bits) { return (bits & ~1ull) == kAbort; }

template <int NV>
__device__ __forceinline__ void store_words(unsigned long long *p, const unsigned long long (&w)[NV])
{
    static_assert(NV >= 1 && NV <= kNV, "1..4 values");
    if (NV == 1) st_relaxed_u64(p, w[0]);
    if (NV >= 2) st_relaxed_v2(p, w[0], w[1]);
    if (NV == 3) st_relaxed_u64(p + 2, w[2]);
    if (NV == 4) st_relaxed_v2(p + 2, w[2], w[NV - 1]);
}
template <int NV>
__device__ __forceinline__ void load_words(const unsigned long long *p, unsigned long long (&w)[NV])
{
    if (NV == 1) w[0] = ld_relaxed_u64(p);
    if (NV >= 2) ld_relaxed_v2(p, w[0], w[1]);
    if (NV == 3) w[2] = ld_relaxed_u64(p + 2);
    if (NV == 4) ld_relaxed_v2(p + 2, w[2], w[NV - 1]);
}

// arrive: thread 0 of every CTA.  abort = true publishes the abort code instead (a watchdog fired in this CTA).
template <int NV>
__device__ __forceinline__ void grid_arrive(unsigned long long *slots, unsigned int gen, const double *v, bool abort = false)
{
    unsigned long long w[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) w[i] = abort ? (kAbort | (unsigned long long)(gen & 1u)) : tagged(v[i], gen);
    store_words<NV>(slots + kNV * blockIdx.x, w);
}

// root: warp 0 of CTA 0.  All loads of a poll round are issued before the first one is tested.
template <int NV, int PER_LANE = 5>
__device__ __forceinline__ void grid_root(unsigned long long *slots, unsigned int gen, int ncta, int lane)
{
    const long long t0 = clock64();
    const unsigned long long par = (unsigned long long)(gen & 1u);
    bool abort = false;
    double tot[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) tot[i] = 0.0;
    for (int base = 0; base < ncta; base += 32 * PER_LANE) {
        unsigned long long w[PER_LANE][NV];
        bool ready;
        do {
#pragma unroll
            for (int k = 0; k < PER_LANE; k++) {
                const int b = base + k * 32 + lane;
                load_words<NV>(slots + kNV * (b < ncta ? b : 0), w[k]);
            }
            ready = true;
#pragma unroll
            for (int k = 0; k < PER_LANE; k++)
#pragma unroll
                for (int i = 0; i < NV; i++) ready = ready & ((w[k][i] & 1ull) == par);
            if (!ready && clock64() - t0 > kWatchdogCycles) { abort = true; break; }
        } while (!ready);
#pragma unroll
        for (int k = 0; k < PER_LANE; k++)
#pragma unroll
            for (int i = 0; i < NV; i++) {
                if (base + k * 32 + lane < ncta) { tot[i] += __longlong_as_double((long long)w[k][i]); abort = abort || is_abort(w[k][i]); }
            }
    }
#pragma unroll
    for (int i = 0; i < NV; i++) tot[i] = warp_sum(tot[i]);
    abort = __any_sync(0xffffffffu, abort);
    if (lane == 0) {
        unsigned long long w[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) w[i] = abort ? (kAbort | par) : tagged(tot[i], gen);
        store_words<NV>(slots + kTotOff, w);
    }
}

// wait: one thread per CTA; out[i] = total i; returns false on abort (root or watchdog)
template <int NV>
__device__ __forceinline__ bool grid_wait(unsigned long long *slots, unsigned int gen, double *out)
{
    const long long t0 = clock64();
    const unsigned long long par = (unsigned long long)(gen & 1u);
    unsigned long long w[NV];
    bool ready, ok = true;
    do {
        load_words<NV>(slots + kTotOff, w);
        ready = true;
#pragma unroll
        for (int i = 0; i < NV; i++) ready = ready & ((w[i] & 1ull) == par);
        if (!ready && clock64() - t0 > 2 * kWatchdogCycles) { ok = false; break; }
    } while (!ready);
#pragma unroll
    for (int i = 0; i < NV; i++) { out[i] = __longlong_as_double((long long)w[i]); ok = ok && !is_abort(w[i]); }
    return ok;
}

}  // namespace gsync
}  // namespace foto
