// grid_sync.cuh -- grid-wide all-reduce of NV <= 3 doubles for persistent cooperative kernels (one CTA per SM).
//
// Two-hop root gather through L2 (tools/ubench_barrier.cu: 2 450 cycles for 148 CTAs, the floor of L2 signalling;
// a store -> poll hand-off between two SMs costs 830-1 000 cycles).  Slots hold the raw bits of the partial sums;
// a sentinel NaN pattern marks "not yet written", so a value is its own ready flag and no counter is needed.
// Slots are triple buffered by generation: a CTA resets its slot of generation g+1 before publishing generation g.
//   layout: [3 generations][kGsNV values][kMaxBlocks] partials, then [3][16] words with the totals.
// arrive: thread 0 of every CTA (after a __syncthreads that follows the CTA's global writes when `publish`);
// root:   warp 0 of CTA 0 sums the partials in a fixed order and writes the totals;
// wait:   lane i of warp 0 polls total i.   Every CTA gets bit-identical totals.
// Readers rely on a control-dependent poll followed by __syncthreads; -DFOTO_PARANOID_FENCES adds the acquire fences.
#pragma once
#include "common.cuh"

namespace foto {
namespace gsync {

constexpr unsigned long long kSentinel = 0x7FF8DEADBEEF0001ull;
constexpr unsigned long long kAbort = 0x7FF8DEADBEEF0002ull;
constexpr unsigned long long kPlainNaN = 0x7FF8000000000000ull;
constexpr int kNV = 3;                                   // most values per all-reduce (slot layout stride)
constexpr int kBcastOff = 3 * kNV * kMaxBlocks;          // slots: [3 gen][kNV][kMaxBlocks] partials, then [3][16] totals
constexpr int kSlotWords = kBcastOff + 3 * 16;
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
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned long long enc(double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b == kSentinel || b == kAbort) ? kPlainNaN : b;
}

// grid all-reduce of NV (<= kNV) values: arrive (thread 0), root gather (warp 0 of CTA 0), wait (thread 0).
template <int NV>
__device__ __forceinline__ void grid_arrive(unsigned long long *slots, unsigned int gen, const double *v, bool publish)
{
#pragma unroll
    for (int i = 0; i < NV; i++) st_relaxed_u64(slots + (((gen + 1u) % 3u) * kNV + i) * kMaxBlocks + blockIdx.x, kSentinel);
    if (publish) fence_acq_rel_gpu();
#pragma unroll
    for (int i = 0; i < NV; i++) st_relaxed_u64(slots + ((gen % 3u) * kNV + i) * kMaxBlocks + blockIdx.x, enc(v[i]));
}

template <int NV, int PER_LANE = 8>
__device__ __forceinline__ void grid_root(unsigned long long *slots, unsigned int gen, int ncta, int lane)
{
    const long long t0 = clock64();
    bool abort = false;
    double tot[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) tot[i] = 0.0;
    const unsigned long long *cur = slots + (gen % 3u) * kNV * kMaxBlocks;
    for (int base = 0; base < ncta; base += 32 * PER_LANE) {
        unsigned long long v[NV][PER_LANE];
        bool ready;
        do {                                             // all NV x PER_LANE polls of a lane are in flight together
            ready = true;
#pragma unroll
            for (int i = 0; i < NV; i++)
#pragma unroll
                for (int k = 0; k < PER_LANE; k++) {
                    const int b = base + k * 32 + lane;
                    v[i][k] = b < ncta ? ld_relaxed_u64(cur + i * kMaxBlocks + b) : 0ull;
                    ready = ready && v[i][k] != kSentinel;
                }
            if (!ready && clock64() - t0 > kWatchdogCycles) { abort = true; break; }
        } while (!ready);
#pragma unroll
        for (int i = 0; i < NV; i++)
#pragma unroll
            for (int k = 0; k < PER_LANE; k++) tot[i] += __longlong_as_double((long long)v[i][k]);
    }
#ifdef FOTO_PARANOID_FENCES
    fence_acq_rel_gpu();                                 // acquire the partials, release the totals
#endif
#pragma unroll
    for (int i = 0; i < NV; i++) tot[i] = warp_sum(tot[i]);
    abort = __any_sync(0xffffffffu, abort);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) st_relaxed_u64(slots + kBcastOff + ((gen + 1u) % 3u) * 16 + i, kSentinel);
#pragma unroll
        for (int i = 0; i < NV; i++) st_relaxed_u64(slots + kBcastOff + (gen % 3u) * 16 + i, abort ? kAbort : enc(tot[i]));
    }
}

// one polling lane per value; returns the bits of value i (kAbort on a watchdog / root abort)
__device__ __forceinline__ unsigned long long grid_wait(unsigned long long *slots, unsigned int gen, int i)
{
    const long long t0 = clock64();
    const unsigned long long *p = slots + kBcastOff + (gen % 3u) * 16 + i;
    unsigned long long bits;
    while ((bits = ld_relaxed_u64(p)) == kSentinel) {
        if (clock64() - t0 > 2 * kWatchdogCycles) { bits = kAbort; break; }
    }
#ifdef FOTO_PARANOID_FENCES
    fence_acq_rel_gpu();
#endif
    return bits;
}

}  // namespace gsync
}  // namespace foto
