// rlr_step.cuh — the step kernel of the batched packet-routing simulator (sm_100a) and everything it needs on the device:
// parameter block, Philox streams, deterministic exp/log2, NumPy-order sums, shared-memory layout.  Included by
// rlr_kernels.cu (host side of the C ABI + the small kernels) and by one step_<policy>.cu per policy family, which
// instantiates the kernel template for that policy only — so the seven families compile in parallel and a tuning build
// of one of them takes seconds.  Everything lives in an anonymous namespace (one copy per translation unit).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rlrouting_b200.h"


namespace {

#ifndef RLR_MAXT
#define RLR_MAXT 512
#endif
constexpr int kMaxThreads = RLR_MAXT;
#ifndef RLR_MIN_BLOCKS
#define RLR_MIN_BLOCKS 1
#endif
#ifndef RLR_LB1_BLOCKS
#define RLR_LB1_BLOCKS 6
#endif
#ifndef RLR_LB1_BLOCKS_QROUTE
#define RLR_LB1_BLOCKS_QROUTE 7     // 72 registers: 8 bytes of spills with the four-word window code, and the seventh resident
#endif                              // CTA is worth 6 % on lata.net (the step waits on HBM row reads); 64 registers spill 68 bytes: -19 %
constexpr int kSmemLimit = 227 * 1024;

// ------------------------------------------------------------------------------------------------
// Device parameter block of the step kernel
// ------------------------------------------------------------------------------------------------
struct StepParams {
    int N, sdeg, G, R;
    int cap, K, clock0, ntab;
    long long tsize;
    const int *row_ptr;
    const int *col;
    const unsigned char *heap_pos;
    const unsigned char *next_hop;
    const unsigned short *choice_mask;    // Shortest(random=True): [N*N] bit j = choice[x][d][j]
    int shortest_random;
    unsigned short *gr_choice_mask;   // GlobalRoute: [R][N*N] per-replica choice masks, bit j = choice[x][d][j] (persist across launches)
    int gr_multiway;              // GlobalRoute(multiway=True): keep every optimal neighbour (shortest.py:57-58)
    int *gr_distance;             // GlobalRoute: [R][N*N] distances (0x3FFFFFFF = inf)
    const int *gr_qs_off;         // GlobalRoute: [R][N] queue_size - len(queue) (non-zero only after Network.reset)
    double *Q, *Theta, *Trace;
    uint4 *ring;
    unsigned *qhead, *qtail;
    long long *counters;
    int *status;
    int inj_cap, add_entropy;
    int is_drop;                  // Network(is_drop=True), env.py:355-360
    int phase;                    // bit 0: inject + send + arrive (Network.step); bit 1: agent.learn
    uint4 *rewards;               // [R][N][4] saved Reward records (phase 1 writes, phase 2 reads)
    const unsigned *inj_events;
    unsigned long long seed;
    long long replica_base;
    double lr_q, lr_p, lr_e, discount, discount_trace, conf_decay;
    uint4 *metrics;
    int *sample;                  // Network.sample_route_time: [R][sample_size] routing times in delivery order
    int *sample_idx;              // [R] Network._sample_idx (persists across calls)
    int *sample_clock;            // [R] clock at which the replica reached sample_size deliveries
    int sample_size;
    uint2 *metrics2;              // [K][R] {all_packets, drop_packets} after each step (for the droprate vector)
    unsigned *sendlog;
    // timing model (SPLIT instantiation only): clock advances `slot` per step, arrivals at clock + transtime
    int slot, freq, transtime, bandwidth;
    int general;                  // events outlive a step: persistent heap, busy links, queue scan (env.py:171-190)
    int ecap;                     // event capacity per replica = N * (transtime + 1)
    int *link_sent;               // [R][sdeg]
    unsigned long long *event_heap;   // [R][ecap]
    int *event_count;             // [R]
    uint4 *event_rec;             // [R][ecap]
    int *q_dead;                  // [R][N]
    int max_deg;                  // largest node degree (uniform loop bound of the MAXD > 4 instantiations)
    // paged queues (PAGED instantiations): a per-replica pool of pages of 2^page_log records, linked per queue
    uint4 *pool;                  // [R][n_pages << page_log] records
    unsigned short *page_next;    // [R][n_pages] page that follows in its queue
    unsigned short *page_free;    // [R][n_pages] stack of free pages
    int *page_top;                // [R] entries on that stack
    unsigned short *q_pages;      // [R][N][4] {head page, tail page, spare page, -}
    int page_log, n_pages;
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  Stream law shared with oracle/rlr_oracle.c:
//   counter = (clock, node | purpose<<16 | block<<24, replica_lo, replica_hi), key = seed.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

struct PhiloxStream {
    uint4 w;
    unsigned c0, c1, c2, c3;
    uint2 key;
    int have, block;
    __device__ __forceinline__ PhiloxStream(unsigned long long seed, unsigned long long replica, int clock,
                                            int node, int purpose)
        : c0((unsigned)clock), c1((unsigned)node | ((unsigned)purpose << 16)), c2((unsigned)replica),
          c3((unsigned)(replica >> 32)), key(make_uint2((unsigned)seed, (unsigned)(seed >> 32))), have(0), block(0)
    {
    }
    __device__ __forceinline__ unsigned next()
    {
        if (have == 0) {
            w = philox4x32_10(make_uint4(c0, c1 | ((unsigned)block << 24), c2, c3), key);
            have = 4;
            ++block;
        }
        const unsigned r = w.x;
        w.x = w.y; w.y = w.z; w.z = w.w;
        --have;
        return r;
    }
};

// ------------------------------------------------------------------------------------------------
// Deterministic exp / log2: the algorithm documented in DESIGN.md §5 and restated independently in
// oracle/rlr_oracle.c — IEEE +,-,*,/ and rint only, so CPU and GPU agree bit for bit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double pow2i(int k) { return __longlong_as_double((long long)(k + 1023) << 52); }

__device__ __forceinline__ double hstep(double p, double r, double c) { return __dadd_rn(__dmul_rn(p, r), c); }

__device__ double det_exp(double x)
{
    if (x != x) return x;
    if (x > 709.782712893384) return CUDART_INF;
    if (x < -745.2) return 0.0;
    const double kf = rint(__dmul_rn(x, 1.4426950408889634));
    double r = __dsub_rn(x, __dmul_rn(kf, 0.6931471803691238));
    r = __dsub_rn(r, __dmul_rn(kf, 1.9082149292705877e-10));
    double p = 1.0 / 6227020800.0;
    p = hstep(p, r, 1.0 / 479001600.0);
    p = hstep(p, r, 1.0 / 39916800.0);
    p = hstep(p, r, 1.0 / 3628800.0);
    p = hstep(p, r, 1.0 / 362880.0);
    p = hstep(p, r, 1.0 / 40320.0);
    p = hstep(p, r, 1.0 / 5040.0);
    p = hstep(p, r, 1.0 / 720.0);
    p = hstep(p, r, 1.0 / 120.0);
    p = hstep(p, r, 1.0 / 24.0);
    p = hstep(p, r, 1.0 / 6.0);
    p = hstep(p, r, 0.5);
    p = hstep(p, r, 1.0);
    p = hstep(p, r, 1.0);
    const int k = (int)kf, k1 = k / 2, k2 = k - k1;
    return __dmul_rn(__dmul_rn(p, pow2i(k1)), pow2i(k2));
}

__device__ double det_log2(double x)
{
    if (x != x || x < 0.0) return CUDART_NAN;
    if (x == 0.0) return -CUDART_INF;
    if (x == CUDART_INF) return x;
    int e = 0;
    if (x < 2.2250738585072014e-308) { x = __dmul_rn(x, 18014398509481984.0); e = -54; }
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    e += (int)(b >> 52) - 1023;
    double m = __longlong_as_double((long long)((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
    if (m > 1.4142135623730951) { m = __dmul_rn(m, 0.5); e += 1; }
    const double s = __ddiv_rn(__dsub_rn(m, 1.0), __dadd_rn(m, 1.0)), z = __dmul_rn(s, s);
    double q = 1.0 / 25.0;
    q = hstep(q, z, 1.0 / 23.0);
    q = hstep(q, z, 1.0 / 21.0);
    q = hstep(q, z, 1.0 / 19.0);
    q = hstep(q, z, 1.0 / 17.0);
    q = hstep(q, z, 1.0 / 15.0);
    q = hstep(q, z, 1.0 / 13.0);
    q = hstep(q, z, 1.0 / 11.0);
    q = hstep(q, z, 1.0 / 9.0);
    q = hstep(q, z, 1.0 / 7.0);
    q = hstep(q, z, 1.0 / 5.0);
    q = hstep(q, z, 1.0 / 3.0);
    q = hstep(q, z, 1.0);
    const double lnm = __dmul_rn(__dmul_rn(2.0, s), q);
    return __dadd_rn((double)e, __dmul_rn(lnm, 1.4426950408889634));
}

// Four independent evaluations of det_exp / det_log2 run side by side: the same operations per element (so the same
// bits), but the four Horner chains interleave instead of running one after the other behind a call.
__device__ __forceinline__ void det_exp4(const double (&x)[4], double (&y)[4])
{
    double kf[4], r[4], p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        kf[j] = rint(__dmul_rn(x[j], 1.4426950408889634));
        r[j] = __dsub_rn(x[j], __dmul_rn(kf[j], 0.6931471803691238));
        r[j] = __dsub_rn(r[j], __dmul_rn(kf[j], 1.9082149292705877e-10));
        p[j] = 1.0 / 6227020800.0;
    }
    const double c[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
                          1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};
#pragma unroll
    for (int i = 0; i < 13; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = hstep(p[j], r[j], c[i]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = (int)kf[j], k1 = k / 2, k2 = k - k1;
        double v = __dmul_rn(__dmul_rn(p[j], pow2i(k1)), pow2i(k2));
        if (x[j] < -745.2) v = 0.0;                      // the special cases of det_exp, applied last
        if (x[j] > 709.782712893384) v = CUDART_INF;
        if (x[j] != x[j]) v = x[j];
        y[j] = v;
    }
}

__device__ __forceinline__ void det_log2_4(const double (&xin)[4], double (&y)[4])
{
    double s[4], z[4], q[4];
    int e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double x = xin[j];
        e[j] = 0;
        if (x < 2.2250738585072014e-308) { x = __dmul_rn(x, 18014398509481984.0); e[j] = -54; }
        const unsigned long long b = (unsigned long long)__double_as_longlong(x);
        e[j] += (int)(b >> 52) - 1023;
        double m = __longlong_as_double((long long)((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
        if (m > 1.4142135623730951) { m = __dmul_rn(m, 0.5); e[j] += 1; }
        s[j] = __ddiv_rn(__dsub_rn(m, 1.0), __dadd_rn(m, 1.0));
        z[j] = __dmul_rn(s[j], s[j]);
        q[j] = 1.0 / 25.0;
    }
    const double c[12] = {1.0 / 23.0, 1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0,
                          1.0 / 5.0, 1.0 / 3.0, 1.0};
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = hstep(q[j], z[j], c[i]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double lnm = __dmul_rn(__dmul_rn(2.0, s[j]), q[j]);
        double v = __dadd_rn((double)e[j], __dmul_rn(lnm, 1.4426950408889634));
        const double x = xin[j];
        if (x == CUDART_INF) v = x;                      // the special cases of det_log2, applied last
        if (x == 0.0) v = -CUDART_INF;
        if (x != x || x < 0.0) v = CUDART_NAN;
        y[j] = v;
    }
}

// ndarray.sum() order for a short row (NumPy pairwise sum, n < 128): sequential below 8 elements,
// else 8 running accumulators combined as ((0+1)+(2+3))+((4+5)+(6+7)) plus a sequential tail.
template <int MAXD>
__device__ __forceinline__ double np_sum_row(const double (&a)[MAXD], int n)
{
    if (MAXD < 8 || n < 8) {
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < MAXD; ++j)
            if (j < n) r = __dadd_rn(r, a[j]);
        return r;
    }
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = a[j < MAXD ? j : 0];
    const int nb = n - (n % 8);
#pragma unroll
    for (int j = 8; j < MAXD; ++j)
        if (j < nb) acc[j & 7] = __dadd_rn(acc[j & 7], a[j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3])),
                           __dadd_rn(__dadd_rn(acc[4], acc[5]), __dadd_rn(acc[6], acc[7])));
#pragma unroll
    for (int j = 8; j < MAXD; ++j)
        if (j >= nb && j < n) res = __dadd_rn(res, a[j]);
    return res;
}

// ndarray.sum() in NumPy's order over the values of one replica's senders (MaHybridQ.learn's r.sum(), max_Q_y.sum(),
// max_Q_x_d.sum(), multi_agent.py:28-29), computed by one whole warp: the three reductions run on helper warps while
// the node warps process arrivals.  v is indexed by thread slot; the replica's senders are the set bits of its window of the
// CTA-wide "sent" bit array (words wm[wlo .. wlo+nw), first/last word masked).  n < 128 (NumPy's single pairwise block).
template <int MAXW>
__device__ double np_sum_warp(const double *v, const unsigned *wm, int wlo, int nw, unsigned mfirst, unsigned mlast, int lane)
{
    unsigned m[MAXW], cum[MAXW + 1];
    cum[0] = 0;
#pragma unroll
    for (int i = 0; i < MAXW; ++i) {
        m[i] = (i < nw) ? wm[wlo + i] : 0u;
        if (i == 0) m[i] &= mfirst;
        if (i == nw - 1) m[i] &= mlast;
        cum[i + 1] = cum[i] + __popc(m[i]);
    }
    const int n = (int)cum[MAXW];
    double a[4];                                    // element e = lane + 32 q of the compacted sender list
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int e = lane + 32 * q;
        a[q] = 0.0;
        if (e < n) {
#pragma unroll
            for (int i = 0; i < MAXW; ++i)
                if ((unsigned)e >= cum[i] && (unsigned)e < cum[i + 1])
                    a[q] = v[(wlo + i) * 32 + __fns(m[i], 0, e - (int)cum[i] + 1)];
        }
    }
    const unsigned full = 0xFFFFFFFFu;
    double res = 0.0;
    if (n < 8) {
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, __shfl_sync(full, a[0], i));
        return res;
    }
    const int nb = n - (n % 8);
    double acc = a[0];                              // lanes 0..7 hold the eight running accumulators
    for (int mm = 1; 8 * mm < nb; ++mm) {
        const int q = mm >> 2, src = (lane + 8 * (mm & 3)) & 31;
        const double val = __shfl_sync(full, q == 0 ? a[0] : q == 1 ? a[1] : q == 2 ? a[2] : a[3], src);
        if (lane < 8 && lane + 8 * mm < nb) acc = __dadd_rn(acc, val);
    }
    const double t1 = __dadd_rn(acc, __shfl_down_sync(full, acc, 1));          // (r0+r1) at lane 0, (r2+r3) at lane 2, ...
    const double t2 = __dadd_rn(t1, __shfl_down_sync(full, t1, 2));            // ((r0+r1)+(r2+r3)) at lane 0, ... at lane 4
    res = __dadd_rn(t2, __shfl_down_sync(full, t2, 4));
    res = __shfl_sync(full, res, 0);
    for (int i = nb; i < n; ++i) {
        const int q = i >> 5;
        res = __dadd_rn(res, __shfl_sync(full, q == 0 ? a[0] : q == 1 ? a[1] : q == 2 ? a[2] : a[3], i & 31));
    }
    return res;
}

__device__ __forceinline__ void set_status(int *status, long long r, int code, int clock)
{
    if (atomicCAS(&status[2 * r], 0, code) == 0) status[2 * r + 1] = clock;
}

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) for staging the per-replica tables: one thread moves a whole
// table between HBM and shared memory; sizes and addresses are multiples of 16 bytes ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- predicated memory operations.  The step is a latency-bound dependent chain: a short `if` around a
// load or store becomes a divergent branch region that ptxas will not schedule across, so the loop-free arrival phase
// issues its accesses as predicated instructions instead (lanes whose predicate is off do not touch memory). ----
// (outputs are write-only: a lane whose predicate is off gets an undefined value, which every caller masks)
__device__ __forceinline__ uint4 lds128_if(const uint4 *ptr, bool pred)
{
    uint4 v;
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %5, 0;\n@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n}"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(ptr)), "r"((int)pred) : "memory");
    return v;
}
__device__ __forceinline__ void stg128_if(uint4 *ptr, const uint4 &v, bool pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %5, 0;\n@p st.global.v4.u32 [%0], {%1, %2, %3, %4};\n}"
                 :: "l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void sts128_if(uint4 *ptr, const uint4 &v, bool pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %5, 0;\n@p st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n}"
                 :: "r"(smem_u32(ptr)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((int)pred) : "memory");
}
// ---- asynchronous global -> shared copies (cp.async, SASS LDGSTS): they complete in commit groups, not on a register
// scoreboard, so a load issued now can be consumed D steps later without any instruction in between waiting for it ----
__device__ __forceinline__ void cp_async16_if(void *dst_smem, const void *src_gmem, bool pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p cp.async.cg.shared.global [%0], [%1], 16;\n}"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void cp_async8_if(void *dst_smem, const void *src_gmem, bool pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p cp.async.ca.shared.global [%0], [%1], 8;\n}"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"((int)pred) : "memory");
}
// the same on 32-bit shared-window addresses (the staged queue heads are addressed as base + (i & (D-1)) * row bytes, see
// stage_at in the step kernel; mad_u32 keeps ptxas from turning the multiply by a 0/1 value into compare + select + add)
__device__ __forceinline__ unsigned mad_u32(unsigned a, unsigned b, unsigned c)
{
    unsigned r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint4 lds128_at(unsigned addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// (volatile: the loads of one row keep their order, i.e. leave back to back instead of one before each comparison)
__device__ __forceinline__ double lds_row(const double *ptr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_u32(ptr)) : "memory");
    return v;
}
__device__ __forceinline__ double lds_row(const float *ptr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(ptr)) : "memory");
    return (double)v;
}
__device__ __forceinline__ unsigned lds32_at(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128_at(unsigned addr, const uint4 &v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts128_at_if(unsigned addr, const uint4 &v, bool pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %5, 0;\n@p st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n}"
                 :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void cp_async16_at_if(unsigned dst, const void *src_gmem, bool pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p cp.async.cg.shared.global [%0], [%1], 16;\n}"
                 :: "r"(dst), "l"(src_gmem), "r"((int)pred) : "memory");
}
// Loop constants of the step kernel kept in ordinary registers: xor-ing with a zero that was read from shared memory hides
// where the value came from, so the compiler cannot go back to the parameter block and reload it (LDCU into a uniform
// register) in every iteration.
__device__ __forceinline__ unsigned opaque(unsigned v, unsigned zero) { return v ^ zero; }
__device__ __forceinline__ int opaque(int v, unsigned zero) { return (int)((unsigned)v ^ zero); }
__device__ __forceinline__ double opaque(double v, unsigned zero)
{
    return __hiloint2double(__double2hiint(v) ^ (int)zero, __double2loint(v) ^ (int)zero);
}
template <typename T>
__device__ __forceinline__ T *opaque(T *v, unsigned zero)
{
    return reinterpret_cast<T *>(reinterpret_cast<unsigned long long>(v) ^ (((unsigned long long)zero << 32) | zero));
}
// One slot of a row scan over a row read whole from shared memory: best = max(best, v) (and k = j) if slot j exists (j < deg)
// and v > best — the degree test is the predicate input of the fp64 compare, the selects follow: 4 (5) instructions per slot
// (volatile, like the row loads before them: all loads of the row first, then the comparisons)
__device__ __forceinline__ void max_slot(double &best, double v, int deg, int j)
{
    asm volatile("{\n.reg .pred c, p;\nsetp.gt.s32 c, %2, %3;\nsetp.gt.and.f64 p, %1, %0, c;\nselp.f64 %0, %1, %0, p;\n}"
        : "+d"(best) : "d"(v), "r"(deg), "r"(j));
}
__device__ __forceinline__ void argmax_slot(double &best, int &k, double v, int deg, int j)
{
    asm volatile("{\n.reg .pred c, p;\nsetp.gt.s32 c, %3, %4;\nsetp.gt.and.f64 p, %2, %0, c;\nselp.f64 %0, %2, %0, p;\nselp.s32 %1, %4, %1, p;\n}"
        : "+d"(best), "+r"(k) : "d"(v), "r"(deg), "r"(j));
}
// acc += v under a predicate: one predicated add, where `acc += pred ? v : 0` compiles to a select and an add
__device__ __forceinline__ void add_if(unsigned &acc, unsigned v, bool pred)
{
    asm("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p add.u32 %0, %0, %1;\n}" : "+r"(acc) : "r"(v), "r"((int)pred));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }   // committed or not
__device__ __forceinline__ void sts32f_if(float *ptr, float v, bool pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p st.shared.f32 [%0], %1;\n}"
                 :: "r"(smem_u32(ptr)), "f"(v), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void sts64_if(double *ptr, double v, bool pred)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p st.shared.f64 [%0], %1;\n}"
                 :: "r"(smem_u32(ptr)), "d"(v), "r"((int)pred) : "memory");
}

// Shared-memory carve-up, identical on host (sizing) and device (pointers).
// Depth of the head-record staging of the fused path: the record a node will pop D steps from now is already on its way
// into shared memory (one cp.async per pop), so DRAM latency (~1 us on B200, more in the tail of 32 lanes) is off the step's
// dependent chain.  Power of two.
// Tables in shared memory: depth 2.  Tables in HBM (lata.net: 312 KB per table): depth 4 with one more step of slack, so that
// right after a pop the record that will be popped NEXT step is guaranteed to have landed; its destination tells which row of
// this node's table the next step's `choose` will read, and that row is prefetched a whole step ahead of its use.
__host__ __device__ constexpr int stage_depth(bool smem_tables) { return smem_tables ? 2 : 4; }
__device__ __forceinline__ void prefetch_l1(const void *ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

struct SmemLayout {
    size_t tables, rec, tp, wmask, rowptr, col, nexthop, heappos, ull, u32, f64, gr, samp, evheap, evorder, evsent, evlen, evfrom, mbar,
        stage, swin, nbr, total;
    bool heap_in_smem;
    // fused = the Network.train instantiation (SPLIT = false): it stages queue heads and the injection schedule in shared
    // memory, has no sample_route_time staging, and when the tables fill shared memory it reads heap_pos through L1 instead
    // (the table is read-only and shared by every CTA of the SM; nothing else of the fused loop allocates in L1)
    __host__ __device__ SmemLayout(int N, int sdeg, int G, int threads, int policy, bool smem_tables, long long tsize, int ntab,
                                   int ecap, bool fused, int esize = 8)
    {
        const int GN = G * N;
        size_t o = 0;
        tables = o; o += smem_tables ? align16((size_t)G * ntab * tsize * esize) : 0;      // esize: 8 (fp64) or 4 (fp32 Qroute)
        rec = o; o += align16((size_t)GN * 16);
        f64 = o; o += (policy == RLR_POLICY_MAHYBRIDQ) ? align16((size_t)(3 * GN + 4 * G) * 8)
                       : (policy == RLR_POLICY_CDRQ) ? align16((size_t)GN * 32) : 0;   // CDRQ: {C_b, max_Q_b}, len, q_x, d|k, src
        ull = o; o += align16((size_t)G * 3 * 8);            // route_time, sends, injected
        u32 = o; o += align16((size_t)G * 9 * 4);            // end, hops, dead, drop, injected-so-far, ..., free pages
        wmask = o; o += align16((size_t)(threads / 32 + 3) * 4);    // (+ 2: a replica's window is read as three words, masked)
        rowptr = o; o += align16((size_t)(N + 1) * 4);
        col = o; o += align16((size_t)sdeg * 2);
        tp = o; o += align16((size_t)(threads > GN ? threads : GN) * 2);      // one entry per thread: the threads behind the last
                                                                              // replica publish "no send" like an idle node
        heap_in_smem = !(fused && smem_tables);
        heappos = o; o += heap_in_smem ? align16((size_t)(N + 1) * N) : 0;
        nexthop = o; o += (policy == RLR_POLICY_SHORTEST) ? align16((size_t)N * N * 2)      // u8 next hops or u16 choice masks
                           : (policy == RLR_POLICY_GLOBALROUTE) ? align16((size_t)G * N * N * 2) : 0;
        samp = o; o += fused ? 0 : align16((size_t)G * (ecap > N ? ecap : N) * 8);   // sample_route_time staging: (heap position, routing time) per delivery
        gr = o; o += (policy == RLR_POLICY_GLOBALROUTE) ? align16((size_t)G * (N * N + N) * 4) : 0;   // distances + weights
        // general timing model: the event heap, this step's popped events, Node.sent per link, heap length + pop count
        evheap = o; o += ecap ? align16((size_t)G * ((ecap + 3) & ~1) * 4) : 0;
        evorder = o; o += align16((size_t)G * ecap * 4);
        evsent = o; o += ecap ? align16((size_t)G * sdeg * 4) : 0;
        evlen = o; o += ecap ? align16((size_t)G * 2 * 4) : 0;
        evfrom = o; o += ecap ? align16((size_t)sdeg) : 0;
        mbar = o; o += smem_tables ? 16 : 0;               // mbarrier of the bulk table load
        stage = o; o += fused ? align16((size_t)stage_depth(smem_tables) * GN * 16) : 0;   // [D][G*N] records at / after the queue heads
        swin = o; o += fused ? align16((size_t)GN * 16) : 0;                   // [2][G*N][2] injection events ahead
        // fused loop with the tables in shared memory: per directed link (x, k) the next hop y, its degree and the offset of
        // its table block in one word, so that get_info finds the row Q[y][d] with one lookup instead of three dependent ones
        nbr = o; o += (fused && smem_tables) ? align16((size_t)sdeg * 4) : 0;
        total = o;
    }
};

// PolicyGradient._softmax (hybrid.py:16-18): exp(theta) / exp(theta).sum(), no max-subtraction; NumPy sum order.
template <int MAXD>
__device__ __forceinline__ void softmax_row(const double *th, int deg, double (&sm)[MAXD])
{
    double e[MAXD];
    static_assert(MAXD % 4 == 0, "softmax_row works on groups of four");
#pragma unroll
    for (int g4 = 0; g4 < MAXD; g4 += 4) {
        double xin[4], yo[4];
        bool any = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) { xin[j] = (g4 + j < deg) ? th[g4 + j] : 0.0; any |= g4 + j < deg; }
        if (MAXD > 4 && !__any_sync(__activemask(), any)) {          // nobody in the warp has neighbours this far out
#pragma unroll
            for (int j = 0; j < 4; ++j) e[g4 + j] = 0.0;
            continue;
        }
        det_exp4(xin, yo);
#pragma unroll
        for (int j = 0; j < 4; ++j) e[g4 + j] = (g4 + j < deg) ? yo[j] : 0.0;
    }
    const double ssum = np_sum_row<MAXD>(e, deg);
#pragma unroll
    for (int j = 0; j < MAXD; ++j) sm[j] = (j < deg) ? __ddiv_rn(e[j], ssum) : 0.0;
}

#ifdef RLR_EXP_TIMING
__device__ unsigned long long g_phase_cycles[8];
#define RLR_T(i) do { const long long now_ = clock64(); tacc_[i] += now_ - tmark_; tmark_ = now_; } while (0)
#else
#define RLR_T(i) do { } while (0)
#endif

constexpr int kGrInf = 0x3FFFFFFF;

// GlobalRoute._calc_distance (shortest.py:43-58 with mask = all off-diagonal, unit(y) = 1 + queue_size[y]) for ONE
// destination column z: the columns of the distance matrix are independent, so thread z replays the reference's
// sequential Gauss-Seidel (x, y in links[x]) sweeps on its own column.  dist/nh are [N][N] in shared memory, w[y] the
// node weights.  mask[x][z] holds `choice[x][z]` as bits: reset to the improving neighbour on a strict improvement, and with
// multiway every neighbour that ties the current distance is added, like the reference.
__device__ __forceinline__ void gr_column(int z, int N, const int *rowptr, const unsigned short *col, const int *w,
                                          int *dist, unsigned short *mask, bool multiway)
{
    for (int x = 0; x < N; ++x) { dist[x * N + z] = (x == z) ? 0 : kGrInf; mask[x * N + z] = 0; }
    bool changing = true;
    while (changing) {
        changing = false;
        for (int x = 0; x < N; ++x) {
            const int rp = rowptr[x], deg = rowptr[x + 1] - rp;
            int cur = dist[x * N + z];
            unsigned m = mask[x * N + z];
            for (int j = 0; j < deg; ++j) {
                const int y = col[rp + j];
                const int nd = dist[y * N + z] + w[y];
                if (cur > nd) { cur = nd; dist[x * N + z] = nd; m = 1u << j; changing = true; }
                if (multiway && nd < kGrInf && cur == nd) m |= 1u << j;          // shortest.py:57-58
            }
            mask[x * N + z] = (unsigned short)m;
        }
    }
}

constexpr unsigned kNoKey = 0xFFFFFFFFu;   // "no arrival" sort key (real keys are heap position << 16 | send slot)
constexpr unsigned kNoSend = 0x00FFu;      // s_tp entry of a node that did not send this step: target 255 is no node (N <= 255),
                                           // rank 0 keeps the heap_pos index formed from it inside the table
constexpr unsigned kDeadRec = 0xFFFFFFFFu; // word 0 of a ring entry that was sent from the middle of the queue (general timing)

// ------------------------------------------------------------------------------------------------
// The step kernel: n_steps iterations of Network.train's loop body (env.py:400-413, slot=freq=1).
//   MAXD  >= max degree (row loops are fully unrolled and predicated)
//   MAXW  >= 32-bit words a replica's N consecutive thread slots can span (3 for N <= 64, else 9)
// ------------------------------------------------------------------------------------------------
//   LB    launch-bound class: 0 = any CTA up to kMaxThreads threads; 1 = CTAs of at most 128 threads with tables in HBM
//         (one 65..128-node replica per CTA, e.g. lata.net), compiled for RLR_LB1_BLOCKS (Qroute on rings: RLR_LB1_BLOCKS_QROUTE) resident CTAs per SM: the step is
//         bound by the latency of the table rows it reads from L2 / HBM, so what pays is warps in flight (80 registers
//         instead of 126, no spills once the row loops are sized 12 and not 16)
//   QT    element type of the Q table in memory: double, or float for Qroute's fp32 STORAGE mode (Network(dtype='float32')):
//         rows are widened to fp64 when they are read, the TD update is evaluated in fp64 with the same unfused operations
//         and rounded once when it is stored — half the bytes per table, so twice the replicas per SM
//   PAGED queues live in a per-replica pool of small pages instead of one fixed ring per node (fused instantiation only): the
//         reference's queues are unbounded (env.py:100,130) and a ring must be sized for the worst node of the whole run,
//         which is an order of magnitude more than the packets alive at any time.  A queue is a linked list of pages; the
//         node thread keeps the pages under its head and tail and one spare in registers, so a record's address costs a
//         select and a shift more than in a ring, and pages are taken from / returned to the replica's free stack only when
//         a cursor crosses a page boundary (pops return pages before barrier 1, appends take pages after it).
template <int POLICY, int MAXD, int MAXW, bool SMEM_TABLES, bool SPLIT, int LB = 0, typename QT = double, bool PAGED = false>
__global__ void __launch_bounds__(LB ? 128 : kMaxThreads, LB ? (POLICY == RLR_POLICY_QROUTE && !PAGED ? RLR_LB1_BLOCKS_QROUTE : RLR_LB1_BLOCKS) : RLR_MIN_BLOCKS)
rlr_step_kernel(const __grid_constant__ StepParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = p.N, G = p.G, GN = G * N;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr bool FUSED = !SPLIT;            // the Network.train instantiation: staged queue heads, injections at step end
#ifdef RLR_NOFLAT
    constexpr bool kFlatArrivals = false;
#else
    constexpr bool kFlatArrivals = FUSED && (MAXD == 4 || MAXD == 8);
#endif
    // MAXD == 8 (a graph that is mostly a grid, with a few nodes of degree 5..8: 6x6-modified.net): the first four links of every
    // node go through the loop-free arrival code; a node that finds a packet on one of its further links in some step handles
    // that step's arrivals — all of them — with the general sorted-list code instead.  (Not for the wide instantiations: on
    // lata.net — average degree 2.9, hubs of up to 9 — some node of nearly every warp has such a packet, both paths run, and the
    // four unconditional slots cost more than the mark-then-insert pass: 1.19e10 -> 0.84e10 hops/s when tried.)
    constexpr bool kFlatPlus = kFlatArrivals && MAXD > 4;
    constexpr bool kLearnInArrivals = kFlatArrivals && POLICY == RLR_POLICY_QROUTE && SMEM_TABLES;
    // Qroute's and Shortest's fused loops with rings: a replica cannot stop in the middle of a launch there, so "this thread is a node of a
    // running replica" (`act`) is a constant of the launch.  Instead of branching around every phase (a warp never skips
    // them: even the second warp of a 36-node replica holds four nodes), `act` is one more input of the three predicates
    // that start any work — queue not empty, a neighbour's packet is for me, an injection is due — and the phases run
    // unconditionally: an inactive lane keeps head == tail, matches nothing and has no schedule.
    constexpr bool kStaticLive = kFlatArrivals && (POLICY == RLR_POLICY_QROUTE || POLICY == RLR_POLICY_SHORTEST) && !PAGED;
    // Kernels with registers to spare — the degree <= 4 fused loops of the policies without softmax rows — keep loop constants,
    // the sender's scratch and a few addresses in registers across the steps; the others run at their register limit, where
    // a reloaded constant is cheaper than a spill.
    constexpr bool kRegRich = kFlatArrivals && !(POLICY == RLR_POLICY_HYBRIDQ || POLICY == RLR_POLICY_MAHYBRIDQ) && LB == 0;
    constexpr int D = stage_depth(SMEM_TABLES);
    constexpr int kStageWait = SMEM_TABLES ? D - 1 : D - 2;    // commit groups that may still be in flight at the top of a step
    static_assert(sizeof(QT) == 8 || POLICY == RLR_POLICY_QROUTE, "fp32 table storage is implemented for Qroute");
    static_assert(!PAGED || !SPLIT, "paged queues are implemented for the fused Network.train instantiation");
    const SmemLayout L(N, p.sdeg, G, blockDim.x, POLICY, SMEM_TABLES, p.tsize, p.ntab, SPLIT ? p.ecap : 0, FUSED, (int)sizeof(QT));
    constexpr bool kHeapInSmem = !(FUSED && SMEM_TABLES);

    double *s_tables = reinterpret_cast<double *>(smem + L.tables);
    uint4 *s_rec = reinterpret_cast<uint4 *>(smem + L.rec);
    double *s_f64 = reinterpret_cast<double *>(smem + L.f64);
    unsigned long long *s_ull = reinterpret_cast<unsigned long long *>(smem + L.ull);
    unsigned *s_u32 = reinterpret_cast<unsigned *>(smem + L.u32);
    unsigned *s_wmask = reinterpret_cast<unsigned *>(smem + L.wmask);
    int *s_rowptr = reinterpret_cast<int *>(smem + L.rowptr);
    unsigned short *s_col = reinterpret_cast<unsigned short *>(smem + L.col);
    unsigned short *s_tp = reinterpret_cast<unsigned short *>(smem + L.tp);   // target | warp-local rank << 8
    unsigned char *s_heappos = smem + L.heappos;
    unsigned char *s_nexthop = smem + L.nexthop;
    // [D][G*N] staged head records: record i of thread t's queue at [(i & (D-1)) * GN + t].  stage_at(i) forms its 32-bit shared
    // address as (this thread's slot) + (i & (D-1)) * (bytes per row): a mask and a multiply-add
    // (a thread behind the last replica is given slot 0: it only ever reads there — see kStaticLive — and slot t may not exist)
    const unsigned stage_t_u32 = smem_u32(smem + L.stage) + (unsigned)(t < GN ? t : 0) * 16u, stage_row = (unsigned)GN * 16u;
    auto stage_ptr = [&](unsigned i) -> uint4 * { return reinterpret_cast<uint4 *>(smem + L.stage) + (i & (unsigned)(D - 1)) * GN + t; };
    auto stage_at = [&](unsigned i) -> unsigned {
        if (!kRegRich) return smem_u32(stage_ptr(i));              // (uniform operands instead of two more registers)
        return mad_u32(i & (unsigned)(D - 1), stage_row, stage_t_u32);
    };
    unsigned *s_nbr = reinterpret_cast<unsigned *>(smem + L.nbr);   // [sdeg]: y | deg_y << 8 | (N * row_ptr[y]) << 16 of link (x, k) at row_ptr[x] + k
    constexpr bool kPackedNbr = FUSED && SMEM_TABLES;               // (table blocks in shared memory hold fewer than 2^16 entries)
    unsigned *s_win = reinterpret_cast<unsigned *>(smem + L.swin);  // [2][G*N][2]: event e of thread t at [((e >> 1) & 1) * 2 * GN + 2 * t + (e & 1)]
    // loop constants of the fused kernel (made opaque, i.e. register-resident, once the CTA's shared memory is initialised)
    double c_discount = p.discount, c_lr_q = p.lr_q;
    unsigned c_N = (unsigned)N, c_cap = (unsigned)p.cap, c_inj_cap = (unsigned)p.inj_cap;
    const unsigned char *c_heap_pos = p.heap_pos;
    int clock_run = p.clock0;
    auto heap_pos_at = [&](unsigned i) -> unsigned { return kHeapInSmem ? (unsigned)s_heappos[i] : (unsigned)__ldg(c_heap_pos + i); };
    // timing model (compile-time defaults in the fused instantiation)
    const bool general = SPLIT && p.general != 0;
    const int slotw = SPLIT ? p.slot : 1, ttime = SPLIT ? p.transtime : 1, freq = SPLIT ? p.freq : 1;
    const int ecap = SPLIT ? p.ecap : 0, ecap1 = ecap > 0 ? ecap : 1;
    const int sstride = ecap > N ? ecap : N;                        // deliveries one step can stage per replica
    int *s_samp_lat = reinterpret_cast<int *>(smem + L.samp);        // [G][sstride] routing times of this step's deliveries
    unsigned *s_samp_pos = reinterpret_cast<unsigned *>(smem + L.samp) + G * sstride;   // their heap pop positions
    // general timing: Network.event_queue as the heapq array of 32-bit items (arrive_time - clock0) << 16 | link, item i
    // of replica g at s_evheap[g * hstride + 1 + i] so that the children 2i+1, 2i+2 form one aligned 8-byte pair
    unsigned *s_evheap = reinterpret_cast<unsigned *>(smem + L.evheap);
    const int hstride = (ecap + 3) & ~1;
    unsigned *s_evorder = reinterpret_cast<unsigned *>(smem + L.evorder);   // [G][ecap] items popped this step, in pop order
    unsigned char *s_evfrom = smem + L.evfrom;                       // [sdeg] sender of each directed link
    int *s_evsent = reinterpret_cast<int *>(smem + L.evsent);        // [G][sdeg] Node.sent per directed link
    int *s_evlen = reinterpret_cast<int *>(smem + L.evlen);          // [G] {len(event_queue), pops of this step}
    int *s_grdist = reinterpret_cast<int *>(smem + L.gr);          // GlobalRoute: [G][N*N] distances, then [G][N] weights
    int *s_grw = s_grdist + G * N * N;
    // per-replica accumulators of this launch
    unsigned long long *s_sends = s_ull + G, *s_inj = s_ull + 2 * G;
    unsigned *s_end = s_u32, *s_hops = s_u32 + G, *s_dead = s_u32 + 2 * G, *s_drop = s_u32 + 3 * G, *s_injs = s_u32 + 4 * G,
             *s_rt32 = s_u32 + 5 * G, *s_nsamp = s_u32 + 6 * G, *s_sidx = s_u32 + 7 * G;
    int *s_ptop = reinterpret_cast<int *>(s_u32 + 8 * G);          // paged queues: [G] free pages of each replica of the CTA        // delivery latencies of the current step (drained every step by the metric thread)
    // MaHybridQ scratch: r, max_Q_y, max_Q_x_d per node; three sums + spare per replica
    double *s_r = s_f64, *s_qy = s_f64 + GN, *s_qx = s_f64 + 2 * GN, *s_sum = s_f64 + 3 * GN;
    // CDRQ scratch (same region): backward-update inputs of each sender and the queue lengths of 'dual' mode
    double2 *s_bk = reinterpret_cast<double2 *>(s_f64);                          // {C_b, max_Q_b}
    unsigned *s_len = reinterpret_cast<unsigned *>(s_f64 + 2 * GN);              // queue length after injection
    int *s_qxi = reinterpret_cast<int *>(s_f64 + 2 * GN) + GN;                   // q_x = max(1, own length after the pop)
    unsigned short *s_dk = reinterpret_cast<unsigned short *>(s_f64 + 3 * GN);   // dest | action index << 8
    unsigned short *s_src = s_dk + GN;                                           // packet.source
    constexpr bool IS_PG = POLICY == RLR_POLICY_HYBRIDQ || POLICY == RLR_POLICY_MAHYBRIDQ;   // softmax-sampled
    constexpr bool IS_CQ = POLICY == RLR_POLICY_CQ || POLICY == RLR_POLICY_CDRQ;             // confidence tables
    constexpr bool DUAL = POLICY == RLR_POLICY_CDRQ;                                         // Node mode 'dual'

    const long long rblock = (long long)blockIdx.x * G;            // first replica of this CTA
    const int g = t / N, x = t - g * N;
    const long long r = rblock + g;
    const bool is_node = (t < GN) && (r < p.R);
    const int gbase = g * N;
    const long long tstride = (long long)p.ntab * p.tsize;         // smem doubles per replica

    // ---- stage topology (+ tables) into shared memory ----
    // the tables come by TMA bulk copy: thread 0 arms an mbarrier with the byte count and issues one copy per table, the
    // other threads stage the small topology arrays meanwhile and everybody waits on the mbarrier after the CTA barrier
    unsigned long long *s_mbar = reinterpret_cast<unsigned long long *>(smem + L.mbar);
    if (SMEM_TABLES && t == 0) {
        mbar_init(s_mbar, 1);
        int nrep = 0;
        for (int gg = 0; gg < G; ++gg) nrep += (rblock + gg < p.R) ? 1 : 0;
        const unsigned tbytes = (unsigned)(p.tsize * sizeof(QT));
        mbar_expect_tx(s_mbar, (unsigned)(nrep * p.ntab) * tbytes);
        for (int gg = 0; gg < nrep; ++gg)
            for (int tb = 0; tb < p.ntab; ++tb)
                bulk_g2s(reinterpret_cast<char *>(s_tables) + ((size_t)gg * p.ntab + tb) * tbytes,
                         reinterpret_cast<const char *>(tb == 0 ? p.Q : tb == 1 ? p.Theta : p.Trace) + (size_t)(rblock + gg) * tbytes,
                         tbytes, s_mbar);
    }
    for (int i = t; i <= N; i += blockDim.x) s_rowptr[i] = p.row_ptr[i];
    for (int i = t; i < p.sdeg; i += blockDim.x) s_col[i] = (unsigned short)p.col[i];
    if (kPackedNbr)
        for (int i = t; i < p.sdeg; i += blockDim.x) {
            const int yy = p.col[i], ry = p.row_ptr[yy];
            s_nbr[i] = (unsigned)yy | ((unsigned)(p.row_ptr[yy + 1] - ry) << 8) | ((unsigned)(N * ry) << 16);
        }
    if (kHeapInSmem)
        for (int i = t; i < (N + 1) * N; i += blockDim.x) s_heappos[i] = p.heap_pos[i];
    if (POLICY == RLR_POLICY_SHORTEST) {
        if (p.shortest_random)
            for (int i = t; i < N * N; i += blockDim.x) reinterpret_cast<unsigned short *>(s_nexthop)[i] = p.choice_mask[i];
        else
            for (int i = t; i < N * N; i += blockDim.x) s_nexthop[i] = p.next_hop[i];
    }
    if (POLICY == RLR_POLICY_GLOBALROUTE)
        for (int i = t; i < G * N * N; i += blockDim.x) {
            const long long rr = rblock + i / (N * N);
            reinterpret_cast<unsigned short *>(s_nexthop)[i] = rr < p.R ? p.gr_choice_mask[rr * N * N + i % (N * N)] : (unsigned short)0;
        }
    for (int i = t; i < G; i += blockDim.x) {
        s_rt32[i] = 0; s_sends[i] = 0; s_inj[i] = 0; s_end[i] = 0; s_hops[i] = 0; s_drop[i] = 0; s_injs[i] = 0;
        s_dead[i] = (rblock + i < p.R) ? (unsigned)(p.status[2 * (rblock + i)] != 0) : 1u;
        s_nsamp[i] = 0;
        s_ptop[i] = (PAGED && rblock + i < p.R) ? p.page_top[rblock + i] : 0;
        s_sidx[i] = (p.sample && rblock + i < p.R) ? (unsigned)p.sample_idx[rblock + i] : 0u;
        if (p.sample && s_sidx[i] >= (unsigned)p.sample_size) s_dead[i] = 2u;     // already has its samples: frozen
    }
    if (general) {
        for (int i = t; i < G * ecap; i += blockDim.x) {
            const int gg = i / ecap1, j = i - gg * ecap1;
            const unsigned long long e = rblock + gg < p.R ? p.event_heap[(rblock + gg) * ecap + j] : 0ull;
            s_evheap[gg * hstride + 1 + j] = (((unsigned)(e >> 32) - (unsigned)p.clock0) << 16) | ((unsigned)(e >> 16) & 0xFFFFu);
        }
        for (int i = t; i < N; i += blockDim.x)
            for (int j = p.row_ptr[i]; j < p.row_ptr[i + 1]; ++j) s_evfrom[j] = (unsigned char)i;
        for (int i = t; i < G * p.sdeg; i += blockDim.x) {
            const long long rr = rblock + i / p.sdeg;
            s_evsent[i] = rr < p.R ? p.link_sent[rr * p.sdeg + i % p.sdeg] : 0;
        }
        for (int i = t; i < G; i += blockDim.x) {
            s_evlen[2 * i] = rblock + i < p.R ? p.event_count[rblock + i] : 0;
            s_evlen[2 * i + 1] = 0;
        }
    }
    __syncthreads();
    if (SMEM_TABLES) mbar_wait(s_mbar, 0);
    if (kRegRich) {
        const unsigned zero = s_drop[0];            // (set to 0 above)
        c_discount = opaque(c_discount, zero); c_lr_q = opaque(c_lr_q, zero);
        c_N = opaque(c_N, zero); c_cap = opaque(c_cap, zero); c_inj_cap = opaque(c_inj_cap, zero);
        c_heap_pos = opaque(c_heap_pos, zero);
        clock_run = opaque(clock_run, zero);
    }

    // ---- per-thread state, constant over the launch ----
    QT *Qr = nullptr;                                              // this replica's tables
    double *Thr = nullptr, *Trr = nullptr;
    if (POLICY != RLR_POLICY_SHORTEST && POLICY != RLR_POLICY_GLOBALROUTE && t < GN && r < p.R) {
        if (SMEM_TABLES) {
            Qr = reinterpret_cast<QT *>(s_tables) + g * tstride;
            Thr = s_tables + g * tstride + p.tsize;
            Trr = s_tables + g * tstride + 2 * p.tsize;
        } else {
            Qr = reinterpret_cast<QT *>(p.Q) + r * p.tsize;
            Thr = p.Theta ? p.Theta + r * p.tsize : nullptr;
            Trr = p.Trace ? p.Trace + r * p.tsize : nullptr;
        }
    }
    const unsigned capmask = (unsigned)p.cap - 1u;
    unsigned head = 0, tail = 0;
    uint4 pref = make_uint4(0, 0, 0, 0), fresh = pref;            // head record: prefetched from the ring / just appended
    bool use_fresh = false;
    uint4 *myring = nullptr;
    // paged queues: the pages under the head and the tail, the page after the head's (0xFFFF until it is linked) and a spare
    unsigned hpage = 0, hnext = 0xFFFFu, tpage = 0, spare = 0;
    const unsigned plog = PAGED ? (unsigned)p.page_log : 0u, pmask = (1u << plog) - 1u;
    uint4 *pool = nullptr;
    unsigned short *pnext = nullptr, *pfree = nullptr;
    int rp = 0, deg = 0;
    unsigned my_sends = 0, my_inj = 0;
    unsigned ndead = 0;                 // general timing: ring entries in [head, tail) already sent out of order
    long long base_rt = 0, base_end = 0, base_hops = 0, base_all = 0, base_drop = 0, run_rt = 0;
    QT *Qx = nullptr;                                              // table[x] of this node
    double *Thx = nullptr, *Trx = nullptr;
    if (is_node) {
        head = p.qhead[r * N + x];
        tail = p.qtail[r * N + x];
        myring = PAGED ? nullptr : p.ring + ((size_t)r * N + x) * (size_t)p.cap;
        if (PAGED) {
            pool = p.pool + ((size_t)r * p.n_pages << plog);
            pnext = p.page_next + (size_t)r * p.n_pages;
            pfree = p.page_free + (size_t)r * p.n_pages;
            const unsigned short *qp = p.q_pages + ((size_t)r * N + x) * 4;
            hpage = qp[0]; tpage = qp[1]; spare = qp[2];
            hnext = hpage == tpage ? 0xFFFFu : (unsigned)pnext[hpage];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const unsigned idx = head + i;
                const unsigned pg = ((idx ^ head) >> plog) ? hnext : hpage;
                cp_async16_at_if(stage_at(idx), pool + ((pg << plog) | (idx & pmask)), tail - head > (unsigned)i);
            }
        } else if (FUSED) {                     // stage the first D records of the queue
#pragma unroll
            for (int i = 0; i < D; ++i)
                cp_async16_at_if(stage_at(head + i), myring + ((head + i) & capmask), tail - head > (unsigned)i);
        } else if (head != tail) {
            pref = myring[head & capmask];
        }
        if (general) ndead = (unsigned)p.q_dead[r * N + x];
        rp = s_rowptr[x];
        deg = s_rowptr[x + 1] - rp;
        if (POLICY != RLR_POLICY_SHORTEST && POLICY != RLR_POLICY_GLOBALROUTE) {
            Qx = Qr + (long long)N * rp;
            Thx = Thr ? Thr + (long long)N * rp : nullptr;
            Trx = Trr ? Trr + (long long)N * rp : nullptr;
        }
        if (x == N - 1) {                      // the last node's thread keeps the per-step metric record
            base_rt = p.counters[r * RLR_NUM_COUNTERS + RLR_CTR_ROUTE_TIME];
            base_end = p.counters[r * RLR_NUM_COUNTERS + RLR_CTR_END_PACKETS];
            base_hops = p.counters[r * RLR_NUM_COUNTERS + RLR_CTR_HOPS];
            base_all = p.counters[r * RLR_NUM_COUNTERS + RLR_CTR_ALL_PACKETS];
            base_drop = p.counters[r * RLR_NUM_COUNTERS + RLR_CTR_DROP_PACKETS];
        }
    }
    const unsigned long long stream_id = (unsigned long long)(p.replica_base + r);
    // the replica's N thread slots as a window of the CTA-wide "sent this step" bit array (one word per warp)
    // (a thread beyond the CTA's replicas sees an empty window and its own send slot, which always says "no send", as every
    // neighbour: whatever it reads stays inside the arrays and nothing it reads can match — see kStaticLive)
    const bool in_win = is_node || !kStaticLive;
    const int wlo = in_win ? gbase >> 5 : 0, nw = in_win ? ((gbase + N - 1) >> 5) - wlo + 1 : 1;
    const unsigned mfirst = in_win ? 0xFFFFFFFFu << (gbase & 31) : 0u, mlast = in_win ? 0xFFFFFFFFu >> (31 - ((gbase + N - 1) & 31)) : 0u;
    // short windows (MAXW <= 3): all three words are always read and masked — words past the replica's last one with zero
    unsigned wmsk[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        wmsk[i] = (i < nw) ? ((i == 0 ? mfirst : 0xFFFFFFFFu) & (i == nw - 1 ? mlast : 0xFFFFFFFFu)) : 0u;
    const unsigned prank_mask = ((1u << lane) - 1u) & ((gbase > (warp << 5)) ? (0xFFFFFFFFu << (gbase - (warp << 5))) : 0xFFFFFFFFu);
    // neighbour send slots (own slot for j >= deg: a node never sends to itself, so it never matches) and the
    // window word each one falls in
    unsigned nb_slot[MAXD], nb_word[MAXD];
#pragma unroll
    for (int j = 0; j < MAXD; ++j) {
        const int z = (is_node && j < deg) ? s_col[rp + j] : x;
        nb_slot[j] = in_win ? (unsigned)(gbase + z) : (unsigned)t;          // (its own slot: always "no send")
        nb_word[j] = in_win ? (unsigned)(((gbase + z) >> 5) - wlo) : 0u;
    }
    // injection schedule of this node: events (local step << 8 | dest), ascending, 0xFFFFFFFF-terminated
    const unsigned *evp = p.inj_events + ((size_t)(is_node ? r : 0) * N + (is_node ? x : 0)) * (size_t)p.inj_cap;
    unsigned ev = is_node ? evp[0] : 0xFFFFFFFFu;                   // current event and the one after it: the list is
    unsigned ev_next = (!FUSED && is_node && ev != 0xFFFFFFFFu) ? evp[1] : 0xFFFFFFFFu;   // read two entries ahead of its use
    // Fused path: the list is read through a window of two 8-byte chunks (events 2c, 2c + 1 in half c & 1) that cp.async
    // refills a chunk ahead, so no load of the schedule ever sits on a register scoreboard inside the step loop.
    const unsigned *ev_base = evp;
    unsigned eidx = 2u;                                             // next event of the list to take from the window
    int win_t_old = -(1 << 20), win_t_new = -(1 << 20);             // steps at which the two resident chunks were requested
    if (FUSED) {
        cp_async8_if(s_win + 2 * t, ev_base, is_node);
        cp_async8_if(s_win + 2 * GN + 2 * t, ev_base + 2, is_node && p.inj_cap > 2);
        cp_async_wait_all();
        ev_next = is_node ? s_win[2 * t + 1] : 0xFFFFFFFFu;         // (ev, ev_next) in registers, the rest behind them in the window
        cp_async8_if(s_win + 2 * t, ev_base + 4, is_node && p.inj_cap > 4);     // chunk 0 is used up: chunk 2 takes its half
        win_t_new = -1;
    }
    evp += 1;
    bool dead = is_node ? (s_dead[g] != 0u) : true;                 // only sampled policies can die mid-launch
    const bool act = !dead;                                         // (kStaticLive: constant over the launch)
    if (kStaticLive && !act) { ev = 0xFFFFFFFFu; ev_next = 0xFFFFFFFFu; }
    const int gr_off = (POLICY == RLR_POLICY_GLOBALROUTE && is_node) ? p.gr_qs_off[r * N + x] : 0;
    // rarely used options (send log, is_drop, droprate records, sampling, split phases) live in the SPLIT instantiation only
    unsigned *sendlog = (SPLIT && p.sendlog && is_node) ? p.sendlog + r * (long long)p.K * N + x : nullptr;

    // Ring append.  Overflow is detected once per step (tail - head > cap) and reported through the sticky
    // status word; the masked index keeps a wrapped write inside this node's own ring.
    // paged queues: a page from the replica's free stack (only called after barrier 1 of a step, when nobody returns pages)
    auto take_page = [&]() -> unsigned {
        const int i = atomicSub(&s_ptop[g], 1) - 1;
        if (i < 0) {                    // pool exhausted: sticky status, the replica stops stepping from the next step on;
            atomicAdd(&s_ptop[g], 1);   // page 0 keeps this step's remaining writes inside the pool
            set_status(p.status, r, RLR_STATUS_QUEUE_OVERFLOW, p.clock0);
            s_dead[g] = 1u;
            return 0u;
        }
        return (unsigned)pfree[i];
    };
    // the tail has filled its page: link the spare behind it, write there from now on, and fetch a new spare
    auto next_tail_page = [&]() {
        pnext[tpage] = (unsigned short)spare;
        if (tpage == hpage) hnext = spare;
        tpage = spare;
        spare = take_page();
    };
    auto append = [&](const uint4 &rec) {
        if (PAGED) pool[(tpage << plog) | (tail & pmask)] = rec;
        else myring[tail & capmask] = rec;
        if (PAGED) {
            if (tail - head < (unsigned)D) { if (!kRegRich) *stage_ptr(tail) = rec; else sts128_at(stage_at(tail), rec); }
            ++tail;
            if ((tail & pmask) == 0u) next_tail_page();
            return;
        }
        if (FUSED) {                    // a record that lands within D of the head is staged directly (nothing will fetch it)
            if (tail - head < (unsigned)D) { if (!kRegRich) *stage_ptr(tail) = rec; else sts128_at(stage_at(tail), rec); }
        } else if (head == tail) {
            fresh = rec;
            use_fresh = true;
        }
        ++tail;
    };

    // SPLIT = false is the fused Network.train path (both phases, no flag tests in the loop)
    const bool do_env = !SPLIT || (p.phase & 1) != 0, do_learn = !SPLIT || (p.phase & 2) != 0;
    // Fused path: the packets of step s + 1 (new_packet + inject, env.py:298-319) are appended at the END of step s, right
    // after the arrivals of step s — the same queue order as injecting at the start of step s + 1, but off the head of the
    // dependent chain, and with the schedule read from the shared-memory window.
    constexpr bool kInjectAtStart = SPLIT;
    auto inject_step = [&](int s_now, unsigned step, int clk) {
        while ((ev >> 8) == step) {
            append(make_uint4((ev & 0xFFu) | ((unsigned)x << 16), 0u, (unsigned)clk, (unsigned)clk));
            ++my_inj;
            ev = ev_next;
            // the look-ahead register is refilled from the window (nothing waits for this read before the node's next
            // packet); the first event of a chunk requested fewer than D + 2 steps ago (a burst of packets at one node)
            // may still be in flight
            if ((eidx & 1u) == 0u && s_now - win_t_old < D + 2) cp_async_wait_all();
            ev_next = s_win[((eidx >> 1) & 1u) * 2u * GN + 2u * t + (eidx & 1u)];
            ++eidx;
            if ((eidx & 1u) == 0u) {    // chunk eidx/2 - 1 is used up: request chunk eidx/2 + 1 into its half
                const unsigned c = (eidx >> 1) + 1u;
                cp_async8_if(s_win + (c & 1u) * 2u * GN + 2u * t, ev_base + 2u * c, 2u * c < c_inj_cap);
                win_t_old = win_t_new;
                win_t_new = s_now;
            }
        }
    };
    if (!kInjectAtStart && !dead) inject_step(-1, 0u, p.clock0);
    if (PAGED) __syncthreads();                 // these appends took pages; the first pops below return pages
    const bool sampling = SPLIT && p.sample != nullptr;      // sample_route_time also runs on the SPLIT instantiation
    // Very wide nodes (MAXD > 8, e.g. lata.net): the row scans run in groups of four slots, and a group is skipped (one uniform
    // branch per group, so the loads inside a group still issue back to back) when no node of the warp has a neighbour that
    // far: lata Qroute +10 %.  (With MAXD = 8 the branch costs more than the second group: 6x6-modified.net -3 %.)
    const int wdeg = (MAXD > 8 || kFlatPlus) ? __reduce_max_sync(0xFFFFFFFFu, is_node ? deg : 1) : MAXD;
#ifdef RLR_EXP_TIMING
    long long tacc_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tmark_ = clock64();
#endif
    // what a sender carries from the send phase to its learn step: declared outside the step loop, so that a node which does
    // not send keeps stale values (every use is behind `sending`) instead of zeroing eleven registers every step
    int d = 0, k = 0, y = 0;
    double oldq = 0.0, max_qy = 0.0, max_qxd = 0.0, rwd = 0.0, c_f = 0.0;
    // the metric thread of a replica (its last node) and whether it writes per-step records: kept as two bits of one opaque
    // register — left to itself the compiler re-derives both predicates from r, R, x, N and the pointer in every step
    unsigned metric_bits = (is_node && x == N - 1) ? (p.metrics ? 3u : 1u) : 0u;
    if (kRegRich) asm volatile("mov.u32 %0, %0;" : "+r"(metric_bits));
    uint4 *metric_row = p.metrics ? p.metrics + r : nullptr;       // this replica's record of the current step
    for (int s = 0; s < p.K; ++s) {
        if (!kRegRich) { d = 0; k = 0; y = 0; oldq = 0.0; max_qy = 0.0; max_qxd = 0.0; rwd = 0.0; c_f = 0.0; }   // (register-poor builds: nothing lives across steps)
        const int clock = kRegRich ? clock_run : p.clock0 + s * slotw;
        if (kRegRich) clock_run += slotw;
        if (FUSED) cp_async_wait<kStageWait>(); // every copy of the groups before the latest one (two) has landed (see the commit below)
        if ((IS_PG || sampling || PAGED) && is_node) dead = (s_dead[g] != 0u);
        const bool live = !dead;
        bool sending = false;
        int qy_dual = 1;                            // 'dual' mode q_y = max(1, len(nodes[y].queue)), env.py:156
        double sm[MAXD];
        if (live && !do_env) {
            // learn-only call (agent.learn after Network.step): restore this node's Reward of that step
            const uint4 *rw = p.rewards + ((size_t)r * N + x) * 4;
            const uint4 meta = rw[1];
            sending = (meta.x & 1u) != 0u;
            if (sending) {
                k = (int)((meta.x >> 1) & 0xFFu); y = (int)((meta.x >> 9) & 0xFFu); d = (int)(meta.x >> 17);
                rwd = (double)(-(int)meta.y - 1);
                const double2 mq = *reinterpret_cast<const double2 *>(rw + 2);
                max_qy = mq.x; max_qxd = mq.y;
                if (IS_CQ) c_f = mq.y;
                if (DUAL) {                         // what CDRQ.learn reads from the other senders' rewards
                    qy_dual = (int)meta.y;
                    s_bk[t] = *reinterpret_cast<const double2 *>(rw + 3);
                    s_qxi[t] = (int)meta.z;
                    s_dk[t] = (unsigned short)(d | (k << 8));
                    s_src[t] = (unsigned short)(rw[0].x >> 16);
                }
                if (POLICY != RLR_POLICY_SHORTEST && POLICY != RLR_POLICY_GLOBALROUTE) oldq = Qx[d * deg + k];
                if (IS_PG) softmax_row<MAXD>(Thx + d * deg, deg, sm);
                if (POLICY == RLR_POLICY_MAHYBRIDQ) { s_r[t] = rwd; s_qy[t] = max_qy; s_qx[t] = max_qxd; }
            }
        }
        if ((kStaticLive || live) && do_env) {
            // ---- A. new_packet + inject (env.py:298-319): consume this step's events of the schedule ----
            // (the fused instantiation has done this at the end of the previous step, see inject_step below)
            while (kInjectAtStart && (ev >> 8) == (unsigned)s) {
                append(make_uint4((ev & 0xFFu) | ((unsigned)x << 16), 0u, (unsigned)clock, (unsigned)clock));
                ++my_inj;
                if (SPLIT && p.metrics2) atomicAdd(&s_injs[g], 1u);
                ev = ev_next;
                if (ev != 0xFFFFFFFFu) ev_next = *++evp;
            }
            // ---- B. Node._send_default (env.py:162-191): pop the head, choose, get_info ----
            if (DUAL) s_len[t] = tail - head - ndead;                           // len(queue) before this node's own send
            bool have = false;
            uint4 rec;
            if (!kStaticLive) rec = make_uint4(0u, 0u, 0u, 0u);
            if (FUSED) {
                // the head record was staged in shared memory at least D steps ago (or by the append that created it); the
                // pop requests the record that is now D - 1 behind the new head — the loads never touch a register
                const bool nonempty = kStaticLive ? (head != tail) & act : head != tail;
                const unsigned head_slot = stage_at(head);              // the slot this pop vacates is the one its refill lands in:
                                                                        // (head + 1 + (D - 1)) & (D - 1) == head & (D - 1)
                if (kStaticLive) rec = lds128_at(head_slot);            // (read whether or not there is a head: nothing to zero)
                const uint4 *head_ptr = stage_ptr(head);
                if (nonempty) {
                    have = true;
                    if (!kStaticLive) rec = !kRegRich ? *head_ptr : lds128_at(head_slot);
                    ++head;
                    if (PAGED && (head & pmask) == 0u) {        // the head left its page: back onto the free stack
                        pfree[atomicAdd(&s_ptop[g], 1)] = (unsigned short)hpage;
                        hpage = hnext;
                        hnext = hpage == tpage ? 0xFFFFu : (unsigned)pnext[hpage];
                    }
                }
                if (PAGED) {
                    const unsigned idx = head + (D - 1);
                    const unsigned pg = ((idx ^ head) >> plog) ? hnext : hpage;
                    cp_async16_at_if(stage_at(idx), pool + ((pg << plog) | (idx & pmask)),
                                  nonempty && tail - head > (unsigned)(D - 1));
                } else
                cp_async16_at_if(head_slot, myring + ((head + (D - 1)) & capmask),       // (issued only after a pop)
                              nonempty && tail - head > (unsigned)(D - 1));
                if (!SMEM_TABLES && POLICY != RLR_POLICY_SHORTEST && POLICY != RLR_POLICY_GLOBALROUTE && head != tail) {
                    // tables in HBM: the record behind the one just popped is next step's head (staged at least three commit
                    // groups ago, or written by the append that created it); bring the rows `choose` will read for it into L1
                    const unsigned dn = (!kRegRich ? stage_ptr(head)->x : lds32_at(stage_at(head))) & 0xFFFFu;
                    const QT *rowq = Qx + dn * deg;
                    prefetch_l1(rowq);
                    prefetch_l1(rowq + (deg - 1));
                    if (IS_PG) { prefetch_l1(Thx + dn * deg); prefetch_l1(Thx + dn * deg + (deg - 1)); }
                }
            } else if (!general) {
                if (head != tail) {
                    have = true;
                    rec = use_fresh ? fresh : pref;
                    use_fresh = false;
                    ++head;
                    if (head != tail) pref = myring[head & capmask];           // prefetch the next head
                }
            } else if (head != tail) {
                // env.py:171-190 with busy links: avaliable_path is evaluated once, then the queue is scanned in order
                // for the first packet whose chosen link is free; agent.choose runs (and, for the sampled policies,
                // draws from this node's stream of the step) once per packet looked at.
                unsigned am = 0u;
#pragma unroll
                for (int j = 0; j < MAXD; ++j)
                    if (j < deg && s_evsent[g * p.sdeg + rp + j] < p.bandwidth) am |= 1u << j;
                PhiloxStream ps(p.seed, stream_id, clock, x, 1);
                unsigned pos = head;
                while (am != 0u && pos != tail) {
                    // the head record is kept in registers (prefetched when the head last moved, or just appended)
                    const uint4 cand = (pos == head) ? (use_fresh ? fresh : pref) : myring[pos & capmask];
                    if (cand.x != kDeadRec) {
                        const int dd = (int)(cand.x & 0xFFFFu);
                        int kk = 0;
                        if (POLICY == RLR_POLICY_GLOBALROUTE) {
                            const unsigned cm = reinterpret_cast<const unsigned short *>(s_nexthop)[(g * N + x) * N + dd];
                            const int cnt = __popc(cm);
                            const int idx = (p.shortest_random && cnt > 1) ? (int)__umulhi(ps.next(), (unsigned)cnt) : 0;
                            kk = cm ? __fns(cm, 0, idx + 1) : 0;
                        } else if (POLICY == RLR_POLICY_SHORTEST) {
                            if (!p.shortest_random) {
                                kk = s_nexthop[x * N + dd];
                            } else {
                                const unsigned cm = reinterpret_cast<const unsigned short *>(s_nexthop)[x * N + dd];
                                const int cnt = __popc(cm);
                                const int idx = cnt > 1 ? (int)__umulhi(ps.next(), (unsigned)cnt) : 0;
                                kk = __fns(cm, 0, idx + 1);
                            }
                        } else if (POLICY == RLR_POLICY_QROUTE || IS_CQ) {
                            const QT *row = Qx + dd * deg;
                            double best = row[0];
#pragma unroll
                            for (int j = 1; j < MAXD; ++j)
                                if (j < deg) { const double v = row[j]; if (v > best) { best = v; kk = j; } }
                        } else {
                            softmax_row<MAXD>(Thx + dd * deg, deg, sm);
                            bool bad = false;
                            double c = 0.0, cdf[MAXD];
#pragma unroll
                            for (int j = 0; j < MAXD; ++j) {
                                if (j < deg) { bad |= (sm[j] != sm[j]); c = __dadd_rn(c, sm[j]); }
                                cdf[j] = c;
                            }
                            if (bad) {
                                set_status(p.status, r, RLR_STATUS_NAN_PROB, clock);
                                s_dead[g] = 1u;
                                break;
                            }
                            const unsigned a = ps.next() >> 5, b = ps.next() >> 6;
                            const double u = __dmul_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 1.0 / 9007199254740992.0);
#pragma unroll
                            for (int j = 0; j < MAXD - 1; ++j)
                                if (j < deg - 1 && kk == j && __ddiv_rn(cdf[j], c) <= u) kk = j + 1;
                        }
                        if ((am >> kk) & 1u) { have = true; rec = cand; k = kk; break; }
                    }
                    ++pos;
                }
                if (have) {                                                     // queue.pop(i), env.py:177
                    if (pos == head) {
                        ++head;
                        use_fresh = false;
                        while (head != tail) {                                  // next live entry becomes the head
                            pref = myring[head & capmask];
                            if (pref.x != kDeadRec) break;
                            ++head;
                            --ndead;
                        }
                    } else {
                        reinterpret_cast<unsigned *>(myring + (pos & capmask))[0] = kDeadRec;
                        ++ndead;
                    }
                }
            }
            if (have) {
                sending = true;
                d = (int)(rec.x & 0xFFFFu);
                if (!general) k = 0;                // (the queue scan of the general timing model has chosen k already)
                RLR_T(6);                           // A + pop
                if (general) {
                    // k was chosen by the scan; pick up the values Qroute/HybridQ.get_info and learn need
                    if (POLICY != RLR_POLICY_SHORTEST && POLICY != RLR_POLICY_GLOBALROUTE) {
                        const QT *row = Qx + d * deg;
                        double best = row[0];
                        oldq = row[0];
#pragma unroll
                        for (int j = 1; j < MAXD; ++j)
                            if (j < deg) { const double v = row[j]; if (v > best) best = v; if (j == k) oldq = v; }
                        max_qxd = best;
                    }
                } else if (POLICY == RLR_POLICY_GLOBALROUTE) {
                    // shortest.py:38-41 on this replica's table: choices[0], or np.random.choice(choices) with random=True
                    const unsigned cm = reinterpret_cast<const unsigned short *>(s_nexthop)[(g * N + x) * N + d];
                    const int cnt = __popc(cm);
                    int idx = 0;
                    if (p.shortest_random && cnt > 1) {
                        PhiloxStream ps(p.seed, stream_id, clock, x, 1);
                        idx = (int)__umulhi(ps.next(), (unsigned)cnt);
                    }
                    k = cm ? __fns(cm, 0, idx + 1) : 0;
                } else if (POLICY == RLR_POLICY_SHORTEST) {
                    if (!p.shortest_random) {
                        k = s_nexthop[x * N + d];                               // shortest.py:38-41, choices[0]
                    } else {                                                    // np.random.choice(choices), shortest.py:40-41
                        const unsigned cm = reinterpret_cast<const unsigned short *>(s_nexthop)[x * N + d];
                        const int cnt = __popc(cm);
                        int idx = 0;
                        if (cnt > 1) {
                            PhiloxStream ps(p.seed, stream_id, clock, x, 1);
                            idx = (int)__umulhi(ps.next(), (unsigned)cnt);
                        }
                        k = __fns(cm, 0, idx + 1);
                    }
                } else if (POLICY == RLR_POLICY_QROUTE || IS_CQ) {
                    const QT *row = Qx + d * deg;                           // qroute.py:22-27
                    double best = row[0];
                    if constexpr (MAXD > 8) {
#pragma unroll
                        for (int g4 = 0; g4 < MAXD; g4 += 4) {
                            if (g4 > 0 && g4 >= wdeg) break;
#pragma unroll
                            for (int j = g4 ? g4 : 1; j < g4 + 4; ++j)
                                if (j < deg) { const double v = row[j]; if (v > best) { best = v; k = j; } }
                        }
                    } else if constexpr (SMEM_TABLES) {
                        // the row is read whole (a slot past the node's degree is the next row's entry, or — behind the last
                        // row — whatever follows the table in shared memory) and the degree test joins the comparison:
                        // compare + selects instead of a predicated block per slot
                        double rv[MAXD];
#pragma unroll
                        for (int j = 0; j < MAXD; ++j) rv[j] = lds_row(row + j);
                        if constexpr (MAXD == 4) {
                            // first-index argmax as a two-level tournament: (0 vs 1), (2 vs 3), then left vs right — a strict
                            // comparison keeps the lower index on a tie at every level, like the sequential scan, and the two
                            // first-level comparisons need the whole row at once, so its four loads leave together
                            double b = rv[2];
                            int kb = 2;
                            best = rv[0];
                            argmax_slot(best, k, rv[1], deg, 1);
                            argmax_slot(b, kb, rv[3], deg, 3);
                            const bool right = (deg > 2) & (b > best);
                            best = right ? b : best;
                            k = right ? kb : k;
                        } else {
                            best = rv[0];
#pragma unroll
                            for (int j = 1; j < MAXD; ++j) argmax_slot(best, k, rv[j], deg, j);
                        }
                    } else {
#pragma unroll
                        for (int j = 1; j < MAXD; ++j)
                            if (j < deg) { const double v = row[j]; if (v > best) { best = v; k = j; } }
                    }
                    max_qxd = best;
                    oldq = best;
                } else {
                    // PolicyGradient.choose (hybrid.py:16-23): softmax(theta) sampled by inverse CDF
                    softmax_row<MAXD>(Thx + d * deg, deg, sm);
                    bool bad = false;
                    double c = 0.0, cdf[MAXD];
#pragma unroll
                    for (int j = 0; j < MAXD; ++j) {
                        if (j < deg) { bad |= (sm[j] != sm[j]); c = __dadd_rn(c, sm[j]); }
                        cdf[j] = c;
                    }
                    if (bad) {
                        set_status(p.status, r, RLR_STATUS_NAN_PROB, clock);
                        s_dead[g] = 1u;        // takes effect next step; this step's state is void
                    }
                    PhiloxStream ps(p.seed, stream_id, clock, x, 1);
                    const unsigned a = ps.next() >> 5, b = ps.next() >> 6;
                    const double u = __dmul_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 1.0 / 9007199254740992.0);
#pragma unroll
                    for (int j = 0; j < MAXD - 1; ++j)
                        if (j < deg - 1 && k == j && __ddiv_rn(cdf[j], c) <= u) k = j + 1;
                    const QT *row = Qx + d * deg;                           // hybrid.py:49-53
                    double best = row[0];
                    oldq = row[0];
                    if constexpr (MAXD > 8) {
#pragma unroll
                        for (int g4 = 0; g4 < MAXD; g4 += 4) {
                            if (g4 > 0 && g4 >= wdeg) break;
#pragma unroll
                            for (int j = g4 ? g4 : 1; j < g4 + 4; ++j)
                                if (j < deg) { const double v = row[j]; if (v > best) best = v; if (j == k) oldq = v; }
                        }
                    } else {
#pragma unroll
                        for (int j = 1; j < MAXD; ++j)
                            if (j < deg) { const double v = row[j]; if (v > best) best = v; if (j == k) oldq = v; }
                    }
                    max_qxd = best;
                }
                int qoff_y, deg_y;                                              // Q[y] offset and degree of the next hop
                if constexpr (kPackedNbr) {
                    const unsigned nb = s_nbr[rp + k];
                    y = (int)(nb & 0xFFu); deg_y = (int)((nb >> 8) & 0xFFu); qoff_y = (int)(nb >> 16);
                } else {
                    y = s_col[rp + k];
                    qoff_y = N * s_rowptr[y]; deg_y = s_rowptr[y + 1] - s_rowptr[y];
                }
                rec.y += 1u;                                                    // p.hops += 1, env.py:141
                rwd = (double)(-(clock - (int)rec.w) - 1);                      // r = -q_y - t_y, qroute.py:37
                if (IS_CQ) {                        // CQ.get_info (qroute.py:72-78): z = argmax Q[y][d], max_Q_f, C_f
                    const int degy = deg_y;
                    const QT *ry = Qr + qoff_y + d * degy;
                    double m = ry[0];
                    int z = 0;
#pragma unroll
                    for (int j = 1; j < MAXD; ++j)
                        if (j < degy) { const double v = ry[j]; if (v > m) { m = v; z = j; } }
                    max_qy = m;
                    c_f = Thr[qoff_y + d * degy + z];
                    if (DUAL) {                     // CDRQ.get_info (qroute.py:107-115) + _build_info_dual (env.py:154-160)
                        const int src = (int)(rec.x >> 16);
                        const QT *rb = Qx + src * deg;
                        double mb = rb[0];
                        int w = 0;
#pragma unroll
                        for (int j = 1; j < MAXD; ++j)
                            if (j < deg) { const double v = rb[j]; if (v > mb) { mb = v; w = j; } }
                        s_bk[t] = make_double2(Thx[src * deg + w], mb);
                        s_qxi[t] = (int)max(1u, tail - head - ndead);
                        s_dk[t] = (unsigned short)(d | (k << 8));
                        s_src[t] = (unsigned short)src;
                    }
                } else if (POLICY != RLR_POLICY_SHORTEST && POLICY != RLR_POLICY_GLOBALROUTE) {   // max_Q_y, qroute.py:29-30
                    const int degy = deg_y;
                    const int wdegy = MAXD > 8 ? __reduce_max_sync(__activemask(), degy) : MAXD;   // of the lanes sending now
                    const QT *ry = Qr + qoff_y + d * degy;
                    double m = ry[0];
                    if constexpr (MAXD > 8) {
#pragma unroll
                        for (int g4 = 0; g4 < MAXD; g4 += 4) {
                            if (g4 > 0 && g4 >= wdegy) break;
#pragma unroll
                            for (int j = g4 ? g4 : 1; j < g4 + 4; ++j)
                                if (j < degy) { const double v = ry[j]; if (v > m) m = v; }
                        }
                    } else if constexpr (SMEM_TABLES) {
                        double rv[MAXD];
#pragma unroll
                        for (int j = 0; j < MAXD; ++j) rv[j] = lds_row(ry + j);
                        if constexpr (MAXD == 4) {
                            double b = rv[2];
                            m = rv[0];
                            max_slot(m, rv[1], degy, 1);
                            max_slot(b, rv[3], degy, 3);
                            m = ((degy > 2) & (b > m)) ? b : m;
                        } else {
                            m = rv[0];
#pragma unroll
                            for (int j = 1; j < MAXD; ++j) max_slot(m, rv[j], degy, j);
                        }
                    } else {
#pragma unroll
                        for (int j = 1; j < MAXD; ++j)
                            if (j < degy) { const double v = ry[j]; if (v > m) m = v; }
                    }
                    max_qy = m;
                }
                // published with the arrival time already in start_queue (env.py:129), so that the receiver appends it as is
                s_rec[t] = make_uint4(rec.x, rec.y, rec.z, general ? rec.w : (unsigned)(clock + ttime));
                if (general) {                      // Node._send_packet (env.py:135-145): the event outlives the step
                    p.event_rec[(size_t)r * ecap + ((unsigned)(clock + ttime) % (unsigned)(ttime + 1)) * N + x] = rec;
                    s_evsent[g * p.sdeg + rp + k] += 1;
                }
                if (POLICY == RLR_POLICY_MAHYBRIDQ) { s_r[t] = rwd; s_qy[t] = max_qy; s_qx[t] = max_qxd; }
                ++my_sends;
                if (!do_learn && p.rewards) {                                   // Network.step(): hand the Reward to the host
                    uint4 *rw = p.rewards + ((size_t)r * N + x) * 4;
                    rw[0] = rec;
                    rw[1] = make_uint4(1u | ((unsigned)k << 1) | ((unsigned)y << 9) | ((unsigned)d << 17),
                                       (unsigned)(clock - (int)rec.w), DUAL ? (unsigned)s_qxi[t] : 0u, 0u);
                    *reinterpret_cast<double2 *>(rw + 2) = make_double2(max_qy, IS_CQ ? c_f : max_qxd);
                    if (DUAL) *reinterpret_cast<double2 *>(rw + 3) = s_bk[t];
                }
            } else if (!do_learn && p.rewards) {
                p.rewards[((size_t)r * N + x) * 4 + 1] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        if (POLICY == RLR_POLICY_MAHYBRIDQ && t >= GN && do_learn) {
            // MaHybridQ.learn, multi_agent.py:32-33: Trace *= discount_trace.  Nothing reads Trace during the send
            // phase, so the helper threads (t >= G*N) do this dense pass while the node threads send.
            const int h = t - GN, nh = blockDim.x - GN;
            for (int gg = 0; gg < G; ++gg) {
                if (rblock + gg >= p.R || s_dead[gg]) continue;
                double2 *tr = reinterpret_cast<double2 *>(SMEM_TABLES ? s_tables + gg * tstride + 2 * p.tsize
                                                                      : p.Trace + (rblock + gg) * p.tsize);
                const int n2 = (int)(p.tsize >> 1);
                for (int i = h; i < n2; i += nh) {
                    double2 v = tr[i];
                    v.x = __dmul_rn(v.x, p.discount_trace);
                    v.y = __dmul_rn(v.y, p.discount_trace);
                    tr[i] = v;
                }
            }
        }
        RLR_T(0);                                   // A + B
        if (sendlog) sendlog[(long long)s * N] = sending ? ((unsigned)y | ((unsigned)d << 16)) : 0xFFFFFFFFu;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, sending);
        if (lane == 0) s_wmask[warp] = bal;
        s_tp[t] = sending ? (unsigned short)(y | ((general ? k : __popc(bal & prank_mask)) << 8)) : (unsigned short)kNoSend;
        __syncthreads();                                                        // ---- barrier 1 ----
        RLR_T(1);

        // Qroute's learn step only needs what this thread computed before the barrier, and every read of the tables by the
        // other threads' get_info is behind us now: do it first, so its fp64 chain overlaps the shared-memory loads of C
        if (POLICY == RLR_POLICY_QROUTE && do_learn && sending && !kLearnInArrivals) {  // qroute.py:40-44
            const double tq = __dmul_rn(p.discount, max_qy);
            const double u1 = __dadd_rn(rwd, tq);
            const double u2 = __dsub_rn(u1, oldq);
            Qx[d * deg + k] = (QT)__dadd_rn(oldq, __dmul_rn(p.lr_q, u2));
        }
        if (general) {
            // ---- C (general timing). Network.event_queue restated literally (heapq on arrive_time only, env.py:48-49):
            // one thread per replica pushes this step's events in node order and pops everything due by clock + slot;
            // the receivers then take their arrivals from the popped list in pop order.
            // heappush (env.py:144) never sifts here: a new event's arrive_time is >= every one in the heap and `<` is strict,
            // so this step's events are appended in node order — each sender writes its own item at its rank
            unsigned *hp1 = s_evheap + g * hstride;                             // item i at hp1[i + 1]
            unsigned nsend = 0;
            if (live && do_env) {
                unsigned rank = 0;
#pragma unroll
                for (int i = 0; i < MAXW; ++i) {
                    unsigned m = (i < nw) ? s_wmask[wlo + i] : 0u;
                    if (i == 0) m &= mfirst;
                    if (i == nw - 1) m &= mlast;
                    nsend += __popc(m);
                    if (wlo + i < warp) rank += __popc(m);
                    else if (wlo + i == warp) rank += __popc(m & ((1u << lane) - 1u));
                }
                if (sending) hp1[1 + s_evlen[2 * g] + rank] = ((unsigned)(clock + ttime - p.clock0) << 16) | (unsigned)(rp + k);
            }
            __syncthreads();
            RLR_T(7);                               // general: parallel append + barrier
            if (live && do_env && x == 0) {
                unsigned *ord = s_evorder + g * ecap;
                int len = s_evlen[2 * g] + (int)nsend, npop = 0;
                const unsigned end_rel = (unsigned)(clock + slotw - p.clock0);
                while (len > 0 && (hp1[1] >> 16) <= end_rel) {                  // heappop, env.py:350-353 (heapq restated)
                    const unsigned last = hp1[len];
                    --len;
                    unsigned ret = last;
                    if (len > 0) {
                        ret = hp1[1];
                        int pos = 0, child = 1;                                 // _siftup: bubble the smaller child up to a leaf
                        while (child < len) {
                            const uint2 pr = *reinterpret_cast<const uint2 *>(hp1 + child + 1);    // items child, child + 1
                            const bool right = (child + 1 < len) && !(pr.x < (pr.y & 0xFFFF0000u));   // not (left < right)
                            hp1[pos + 1] = right ? pr.y : pr.x;
                            pos = child + (right ? 1 : 0);
                            child = 2 * pos + 1;
                        }
                        while (pos > 0) {                                       // _siftdown(heap, 0, pos) of the former last item
                            const int parent = (pos - 1) >> 1;
                            const unsigned pv = hp1[parent + 1];
                            if (last < (pv & 0xFFFF0000u)) { hp1[pos + 1] = pv; pos = parent; continue; }
                            break;
                        }
                        hp1[pos + 1] = last;
                    }
                    ord[npop++] = ret;
                }
                s_evlen[2 * g] = len;
                s_evlen[2 * g + 1] = npop;
            }
            __syncthreads();
            RLR_T(5);                               // general: heap pops + barrier
            if (live && do_env) {
                // arrivals in pop order: first find this node's events in the popped list (shared-memory reads only), fetch
                // their records together, then process them in order; a node with more than MAXD events (several arrival
                // times due in one drain) repeats the pass from where it stopped
                const int npop = s_evlen[2 * g + 1];
                for (int skip = 0;; skip += MAXD) {
                    unsigned mine[MAXD];
                    int nm = 0, seen = 0;
#pragma unroll 4
                    for (int i = 0; i < npop; ++i) {                            // branch-free: the list is the same for every lane
                        const unsigned e = s_evorder[g * ecap + i];
                        const bool m = (int)s_col[e & 0xFFFFu] == x;
                        const bool take = m && seen >= skip && nm < MAXD;
#pragma unroll
                        for (int j = 0; j < MAXD; ++j)
                            if (take && j == nm) mine[j] = (e & 0xFFFF0000u) | (unsigned)i;      // arrive offset | pop index
                        nm += take ? 1 : 0;
                        seen += m ? 1 : 0;
                    }
                    uint4 recs[MAXD];
#pragma unroll
                    for (int j = 0; j < MAXD; ++j)
                        if (j < nm) {
                            const unsigned e = s_evorder[g * ecap + (mine[j] & 0xFFFFu)];
                            const int at = p.clock0 + (int)(e >> 16);
                            s_evsent[g * p.sdeg + (e & 0xFFFFu)] -= 1;         // env.py:352
                            recs[j] = p.event_rec[(size_t)r * ecap + ((unsigned)at % (unsigned)(ttime + 1)) * N + s_evfrom[e & 0xFFFFu]];
                        }
#pragma unroll
                    for (int j = 0; j < MAXD; ++j)
                        if (j < nm) {
                            uint4 rec = recs[j];
                            const int at = p.clock0 + (int)(mine[j] >> 16);
                            if (p.is_drop && rec.y >= (unsigned)N) {            // env.py:355-360
                                atomicAdd(&s_drop[g], 1u);
                            } else if ((int)(rec.x & 0xFFFFu) == x) {           // Network.end_packet, env.py:321-331
                                if (sampling) {
                                    const unsigned slot = atomicAdd(&s_nsamp[g], 1u);
                                    s_samp_pos[g * sstride + slot] = mine[j] & 0xFFFFu;
                                    s_samp_lat[g * sstride + slot] = at - (int)rec.z;
                                }
                                atomicAdd(&s_end[g], 1u);
                                atomicAdd(&s_rt32[g], (unsigned)(at - (int)rec.z));
                                atomicAdd(&s_hops[g], rec.y);
                            } else {                                            // Node.receive, env.py:127-130
                                rec.w = (unsigned)at;
                                append(rec);
                            }
                        }
                    if (seen <= skip + MAXD) break;
                }
                if (tail - head > (unsigned)p.cap) set_status(p.status, r, RLR_STATUS_QUEUE_OVERFLOW, clock);
            }
        } else if ((kStaticLive || live) && do_env) {
            // ---- C. event drain (env.py:349-364): gather arrivals from neighbours in heap-tie order ----
            unsigned cum[MAXW + 1];         // senders of this replica in the window words before word i
            cum[0] = 0;
#pragma unroll
            for (int i = 0; i < MAXW; ++i) {
                unsigned m;
                if (MAXW <= 3) {
                    m = s_wmask[wlo + i] & wmsk[i < 3 ? i : 0];
                } else {
                    m = (i < nw) ? s_wmask[wlo + i] : 0u;
                    if (i == 0) m &= mfirst;
                    if (i == nw - 1) m &= mlast;
                }
                cum[i + 1] = cum[i] + __popc(m);
            }
            const unsigned hp_row = cum[MAXW] * c_N;                    // heap_pos row k = number of senders
            // cum[] as bytes (N <= 255) of up to three registers, looked up with byte_perm: no indexed local array
            unsigned cumpack = 0, cumpack1 = 0, cumpack2 = 0;
            cumpack = cum[0] | (cum[1] << 8) | (cum[2] << 16) | ((MAXW > 3 ? cum[MAXW > 3 ? 3 : 0] : 0u) << 24);
            if (MAXW > 4) cumpack1 = cum[MAXW > 4 ? 4 : 0] | (cum[MAXW > 5 ? 5 : 0] << 8) | (cum[MAXW > 6 ? 6 : 0] << 16) | (cum[MAXW > 7 ? 7 : 0] << 24);
            if (MAXW > 8) cumpack2 = cum[MAXW > 8 ? 8 : 0];
            auto cum_at = [&](unsigned w) -> unsigned {
                if (MAXW <= 4) return __byte_perm(cumpack, 0u, w);       // (selector w = bytes {w, 0, 0, 0}, and byte 0 is cum[0] = 0)
                const unsigned reg = w < 4u ? cumpack : (w < 8u ? cumpack1 : cumpack2);
                return __byte_perm(reg, 0u, w & 3u) & 0xFFu;
            };
            // sort key per arrival: heap pop position << 16 | send slot, or kNoKey.
            // MAXD == 4: one key per neighbour slot, sorted by a 5-comparator network.  Wider nodes: a first pass only marks
            // the slots that hold a packet for this node (a node receives about one packet a step whatever its degree);
            // the marked slots' keys are then inserted, sorted, into four registers.  Only a node with more than four
            // arrivals in one step falls back to a list in local memory.
            unsigned key[MAXD];
            unsigned q4[4] = {kNoKey, kNoKey, kNoKey, kNoKey};
            int nkey = 0;
            bool use_list = false;
            bool far = false;               // kFlatPlus: a packet arrives on a link beyond this node's fourth
            if (kFlatPlus && wdeg > 4) {    // (uniform: some node of the warp has such links)
#pragma unroll
                for (int j = 4; j < MAXD; ++j) far |= ((s_tp[nb_slot[j]] ^ (unsigned)x) & 0xFFu) == 0u;
            }
            if (MAXD == 4 && !kFlatArrivals) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned tp = s_tp[nb_slot[j]];
                    const bool match = (tp & 0xFFu) == (unsigned)x;             // kNoSend's low byte is 255 > any node id
                    const unsigned pos = heap_pos_at(hp_row + (match ? (tp >> 8) + cum_at(nb_word[j]) : 0u));
                    q4[j] = match ? ((pos << 16) | nb_slot[j]) : kNoKey;
                }
#define RLR_CE(a, b) { const unsigned lo_ = min(q4[a], q4[b]), hi_ = max(q4[a], q4[b]); q4[a] = lo_; q4[b] = hi_; }
                RLR_CE(0, 1) RLR_CE(2, 3) RLR_CE(0, 2) RLR_CE(1, 3) RLR_CE(1, 2)
#undef RLR_CE
            } else if (MAXD > 4 && (!kFlatPlus || far)) {
                unsigned marks = 0u;
#pragma unroll
                for (int j = 0; j < MAXD; ++j) {
                    if (j >= p.max_deg) continue;                               // uniform: no node has a j-th neighbour
                    marks |= ((s_tp[nb_slot[j]] & 0xFFu) == (unsigned)x ? 1u : 0u) << j;
                }
                use_list = __popc(marks) > 4;
                if (!use_list) {
                    while (marks) {
                        const int j = __ffs((int)marks) - 1;
                        marks &= marks - 1u;
                        const unsigned slot = (unsigned)gbase + s_col[rp + j];
                        const unsigned tp = s_tp[slot];
                        const unsigned pos = heap_pos_at(hp_row + (tp >> 8) + cum_at((slot >> 5) - (unsigned)wlo));
                        unsigned tk = (pos << 16) | slot, lo_;                  // sorted insert (ascending)
                        lo_ = min(q4[0], tk); tk = max(q4[0], tk); q4[0] = lo_;
                        lo_ = min(q4[1], tk); tk = max(q4[1], tk); q4[1] = lo_;
                        lo_ = min(q4[2], tk); tk = max(q4[2], tk); q4[2] = lo_;
                        q4[3] = min(q4[3], tk);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < MAXD; ++j) {
                        if (j >= p.max_deg) continue;
                        const unsigned tp = s_tp[nb_slot[j]];
                        if ((tp & 0xFFu) == (unsigned)x)
                            key[nkey++] = (heap_pos_at(hp_row + (tp >> 8) + cum_at(nb_word[j])) << 16) | nb_slot[j];
                    }
                }
            }
            if (kLearnInArrivals) {
                // Qroute.learn (qroute.py:40-44) in the same straight-line block as the arrival loads, so that the fp64
                // chain fills their latency; only the store is predicated on `sending`
                const double tq = __dmul_rn(c_discount, max_qy);
                const double u1 = __dadd_rn(rwd, tq);
                const double u2 = __dsub_rn(u1, oldq);
                const double newq = __dadd_rn(oldq, __dmul_rn(c_lr_q, u2));
                if constexpr (sizeof(QT) == 8) sts64_if(reinterpret_cast<double *>(Qx) + (d * deg + k), newq, sending);
                else sts32f_if(reinterpret_cast<float *>(Qx) + (d * deg + k), (float)newq, sending);
            }
            if (kFlatArrivals) {
                // Loop-free arrivals (degree <= 4): every matching neighbour's record and heap position are fetched side
                // by side, a record's place in the queue is its rank among the appended ones (pairwise compares of the
                // positions), and the <= 4 ring writes are independent of each other.
                uint4 rc[4];
                unsigned posv[4];
                bool mt[4], ap[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned tp = s_tp[nb_slot[j]];
                    mt[j] = ((tp ^ (unsigned)x) & 0xFFu) == 0u;      // (a stopped replica's nodes publish "no send" too)
                    if (kFlatPlus) mt[j] = mt[j] & !far;
                    const unsigned base = cum_at(nb_word[j]);
                    posv[j] = heap_pos_at(hp_row + (tp >> 8) + base);      // (in the table for any slot: kNoSend carries rank 0)
                    rc[j] = lds128_if(s_rec + nb_slot[j], mt[j]);
                }
                unsigned ndel = 0u, rtsum = 0u, hopsum = 0u, ndrop = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool dropped = SPLIT && p.is_drop && rc[j].y >= (unsigned)N;      // env.py:355-360
                    const bool del = mt[j] && !dropped && (int)(rc[j].x & 0xFFFFu) == x;    // Network.end_packet, env.py:321-331
                    ap[j] = mt[j] && !dropped && !del;                                      // Node.receive, env.py:127-130
                    if (SPLIT && mt[j] && dropped) ++ndrop;
                    add_if(ndel, 1u, del);
                    add_if(rtsum, (unsigned)(clock + ttime - (int)rc[j].z), del);
                    add_if(hopsum, rc[j].y, del);
                    if (sampling && del) {                                                  // env.py:325-328, ordered later
                        const unsigned slot = atomicAdd(&s_nsamp[g], 1u);
                        s_samp_pos[g * sstride + slot] = posv[j];
                        s_samp_lat[g * sstride + slot] = clock + ttime - (int)rc[j].z;
                    }
                }
                if (SPLIT && ndrop) atomicAdd(&s_drop[g], ndrop);
                if (ndel) {
                    atomicAdd(&s_end[g], ndel);
                    atomicAdd(&s_rt32[g], rtsum);
                    atomicAdd(&s_hops[g], hopsum);
                }
                const unsigned tail0 = tail;
                // rank among the appended ones by heap position.  Six comparisons, one per pair: the positions of two
                // appended records are distinct, and a pair with a record that is not appended (key 0x100, above every
                // position) only ever raises the rank of that record, which nobody uses
                unsigned keyv[4], rkv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) keyv[j] = ap[j] ? posv[j] : 0x100u;
                {   // m_ij = -1 if key i < key j else 0 (keys are below 2^9: the sign of the difference), sums of three terms
                    const int m01 = (int)(keyv[0] - keyv[1]) >> 31, m02 = (int)(keyv[0] - keyv[2]) >> 31, m03 = (int)(keyv[0] - keyv[3]) >> 31;
                    const int m12 = (int)(keyv[1] - keyv[2]) >> 31, m13 = (int)(keyv[1] - keyv[3]) >> 31, m23 = (int)(keyv[2] - keyv[3]) >> 31;
                    rkv[0] = (unsigned)(3 + m01 + m02 + m03);
                    rkv[1] = (unsigned)(2 - m01 + m12 + m13);
                    rkv[2] = (unsigned)(1 - m02 - m12 + m23);
                    rkv[3] = (unsigned)(-m03 - m13 - m23);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned rk = rkv[j];
                    const unsigned at = tail0 + rk;
                    if (PAGED) {                    // at most one page boundary inside the batch (pages hold >= 4 records)
                        const unsigned pg = ((at ^ tail0) >> plog) ? spare : tpage;
                        stg128_if(pool + ((pg << plog) | (at & pmask)), rc[j], ap[j]);
                    } else {
                        stg128_if(myring + (at & capmask), rc[j], ap[j]);
                    }
                    sts128_at_if(stage_at(at), rc[j], ap[j] && at - head < (unsigned)D);
                    add_if(tail, 1u, ap[j]);
                }
                if (PAGED && ((tail ^ tail0) >> plog)) next_tail_page();
            }
            if (!kFlatArrivals || (kFlatPlus && far)) {
            RLR_T(5);                               // C: window + keys + sort
            for (int taken = 0;; ++taken) {
                unsigned best;
                if (!use_list) {
                    best = q4[0];
                    q4[0] = q4[1]; q4[1] = q4[2]; q4[2] = q4[3]; q4[3] = kNoKey;
                    if (best == kNoKey) break;
                } else {
                    if (taken >= nkey) break;
                    best = key[taken];
                    int bi = taken;
                    for (int b = taken + 1; b < nkey; ++b) {
                        const unsigned v = key[b];
                        if (v < best) { best = v; bi = b; }
                    }
                    key[bi] = key[taken];
                }
                uint4 rec = s_rec[best & 0xFFFFu];
                if (SPLIT && p.is_drop && rec.y >= (unsigned)N) {                        // env.py:355-360: too many hops, dropped
                    atomicAdd(&s_drop[g], 1u);
                    continue;
                }
                if ((int)(rec.x & 0xFFFFu) == x) {                              // Network.end_packet, env.py:321-331
                    if (sampling) {                                             // env.py:325-328, ordered later by pop position
                        const unsigned slot = atomicAdd(&s_nsamp[g], 1u);
                        s_samp_pos[g * sstride + slot] = best >> 16;
                        s_samp_lat[g * sstride + slot] = clock + ttime - (int)rec.z;
                    }
                    atomicAdd(&s_end[g], 1u);
                    atomicAdd(&s_rt32[g], (unsigned)(clock + ttime - (int)rec.z));
                    atomicAdd(&s_hops[g], rec.y);
                } else {                                                        // Node.receive, env.py:127-130
                    rec.w = (unsigned)(clock + ttime);
                    append(rec);
                }
            }
            }
            if (kInjectAtStart && tail - head > (unsigned)p.cap) set_status(p.status, r, RLR_STATUS_QUEUE_OVERFLOW, clock);
        }
        RLR_T(2);                                   // C
        // ---- D. agent.learn ----
        if (!do_learn) {
            if (DUAL && live && sending && p.rewards) {     // Network.step() in 'dual' mode: q_y needs every node's send (env.py:156)
                const unsigned ly = s_len[gbase + y], tpy = s_tp[gbase + y];
                p.rewards[((size_t)r * N + x) * 4 + 1].y = max(1u, ly - ((y < x && tpy != kNoSend) ? 1u : 0u));
            }
        } else if (POLICY == RLR_POLICY_QROUTE) {
            // (done right after barrier 1)
        } else if (POLICY == RLR_POLICY_HYBRIDQ) {
            if (sending) {                                                      // hybrid.py:55-62
                double rr = rwd;
                if (p.add_entropy) {                                            // hybrid.py:38-39
                    double tmp[MAXD];
#pragma unroll
                    for (int g4 = 0; g4 < MAXD; g4 += 4) {
                        double xin[4], yo[4];
                        bool any = false;
#pragma unroll
                        for (int j = 0; j < 4; ++j) { xin[j] = (g4 + j < deg) ? sm[g4 + j] : 1.0; any |= g4 + j < deg; }
                        if (MAXD > 4 && !__any_sync(__activemask(), any)) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) tmp[g4 + j] = 0.0;
                            continue;
                        }
                        det_log2_4(xin, yo);
#pragma unroll
                        for (int j = 0; j < 4; ++j) tmp[g4 + j] = (g4 + j < deg) ? __dmul_rn(sm[g4 + j], yo[j]) : 0.0;
                    }
                    rr = __dsub_rn(rr, __dmul_rn(p.lr_e, np_sum_row<MAXD>(tmp, deg)));
                }
                const double tq = __dmul_rn(p.discount, max_qy);
                const double u1 = __dadd_rn(rr, tq);
                Qx[d * deg + k] = (QT)__dadd_rn(oldq, __dmul_rn(p.lr_q, __dsub_rn(u1, oldq)));
                const double adv = __dsub_rn(u1, max_qxd);                      // hybrid.py:32-36
                double *th = Thx + d * deg;
#pragma unroll
                for (int j = 0; j < MAXD; ++j)
                    if (j < deg) {
                        double gj = -sm[j];
                        if (j == k) gj = __dadd_rn(gj, 1.0);
                        th[j] = __dadd_rn(th[j], __dmul_rn(__dmul_rn(p.lr_p, gj), adv));
                    }
            }
        } else if (POLICY == RLR_POLICY_GLOBALROUTE) {
            // GlobalRoute.learn (shortest.py:99-104): recompute every distance with node weights 1 + queue_size[y]
            if (live) s_grw[gbase + x] = 1 + (int)(tail - head - ndead) + gr_off;
            __syncthreads();
            if (live) gr_column(x, N, s_rowptr, s_col, s_grw + gbase, s_grdist + g * N * N,
                                reinterpret_cast<unsigned short *>(s_nexthop) + g * N * N, p.gr_multiway != 0);
        } else if (IS_CQ) {
            // CQ._update_qtable (qroute.py:80-93): eta = max(C, 1 - old_conf); Q += eta * (r + discount * max_Q - Q);
            // conf += eta * (C - conf); conf /= decay.  CDRQ (qroute.py:117-122) adds the backward update of
            // Q[y][packet.source][idx(x)] with r_b = -q_x; rewards are applied in sender order, forward then backward.
            auto cq_update = [&](double &q, double &c, double rr, double C, double maxQ) {
                const double one_minus = __dsub_rn(1.0, c);
                const double eta = (one_minus > C) ? one_minus : C;
                q = __dadd_rn(q, __dmul_rn(eta, __dsub_rn(__dadd_rn(rr, __dmul_rn(p.discount, maxQ)), q)));
                c = __ddiv_rn(__dadd_rn(c, __dmul_rn(eta, __dsub_rn(C, c))), p.conf_decay);
            };
            if (sending) {
                const int cell = d * deg + k;
                double q = oldq, c = Thx[cell];
                if (!DUAL) {
                    cq_update(q, c, rwd, c_f, max_qy);
                } else {
                    const int yt = gbase + y;                                   // my next hop's thread slot
                    const unsigned ly = s_len[yt];
                    const unsigned tpy = s_tp[yt];
                    // len(nodes[y].queue) as the sequential node loop sees it: y already popped its packet if y < x and y sent
                    // (with busy links a non-empty queue does not imply a send); a learn-only call gets it from the saved reward
                    if (do_env) qy_dual = (int)max(1u, ly - ((y < x && tpy != kNoSend) ? 1u : 0u));
                    const double r_f = -(double)qy_dual;
                    // does y's backward update land on this very cell?  (y sends to me a packet whose source is my dest)
                    const bool hit = tpy != kNoSend && (tpy & 0xFFu) == (unsigned)x && s_src[yt] == (unsigned short)d;
                    const double2 bk = s_bk[yt];
                    const double r_by = -(double)s_qxi[yt];
                    if (hit && y < x) cq_update(q, c, r_by, bk.x, bk.y);        // reward y precedes reward x
                    cq_update(q, c, r_f, c_f, max_qy);
                    if (hit && y > x) cq_update(q, c, r_by, bk.x, bk.y);
                }
                Qx[cell] = (QT)q;
                Thx[cell] = c;
                if (DUAL) {                                                     // my backward update on y's table
                    const int src = (int)s_src[t];
                    const int rpy = s_rowptr[y], degy = s_rowptr[y + 1] - rpy;
                    int idx = 0;
                    for (int j = 1; j < degy; ++j)
                        if (s_col[rpy + j] == (unsigned short)x) idx = j;
                    const unsigned tpy = s_tp[gbase + y];
                    const bool owned_by_y = tpy != kNoSend && s_dk[gbase + y] == (unsigned short)(src | (idx << 8));
                    if (!owned_by_y) {                                          // else y applies it with its own forward update
                        const int cb = N * rpy + src * degy + idx;
                        double qb = Qr[cb], cbv = Thr[cb];
                        const double2 bk = s_bk[t];
                        cq_update(qb, cbv, -(double)s_qxi[t], bk.x, bk.y);
                        Qr[cb] = (QT)qb;
                        Thr[cb] = cbv;
                    }
                }
            }
            __syncthreads();
            // CQ.confidence_decay (qroute.py:99-101): every confidence entry *= decay, every step
            for (int gg = 0; gg < G; ++gg) {
                if (rblock + gg >= p.R || s_dead[gg]) continue;
                double2 *cf = reinterpret_cast<double2 *>(SMEM_TABLES ? s_tables + gg * tstride + p.tsize
                                                                      : p.Theta + (rblock + gg) * p.tsize);
                const int n2 = (int)(p.tsize >> 1);
                for (int i = t; i < n2; i += blockDim.x) {
                    double2 v = cf[i];
                    v.x = __dmul_rn(v.x, p.conf_decay);
                    v.y = __dmul_rn(v.y, p.conf_decay);
                    cf[i] = v;
                }
            }
        } else if (POLICY == RLR_POLICY_MAHYBRIDQ) {
            // MaHybridQ.learn (multi_agent.py:17-43).  Trace was decayed by the helper threads before barrier 1.
            if (sending) {                                                      // multi_agent.py:35-40
                double *trow = Trx + d * deg;
#pragma unroll
                for (int j = 0; j < MAXD; ++j)
                    if (j < deg) {
                        double gj = -sm[j];
                        if (j == k) gj = __dadd_rn(gj, 1.0);
                        trow[j] = __dadd_rn(trow[j], gj);
                    }
                const double u1 = __dadd_rn(rwd, __dmul_rn(p.discount, max_qy));
                Qx[d * deg + k] = (QT)__dadd_rn(oldq, __dmul_rn(p.lr_q, __dsub_rn(u1, oldq)));
            }
            {   // r.sum(), max_Q_y.sum(), max_Q_x_d.sum() (multi_agent.py:28-29), one helper warp per (replica, sum)
                const int hw0 = (GN + 31) >> 5, nhw = (int)(blockDim.x >> 5) - hw0;
                if (warp >= hw0) {
                    for (int pair = warp - hw0; pair < 3 * G; pair += nhw) {
                        const int gg = pair / 3, q = pair - 3 * gg;
                        if (rblock + gg >= p.R || s_dead[gg]) continue;
                        const int gb = gg * N, wl = gb >> 5, nwg = ((gb + N - 1) >> 5) - wl + 1;
                        const double *v = (q == 0) ? s_r : (q == 1) ? s_qy : s_qx;
                        const double sum = np_sum_warp<MAXW>(v, s_wmask, wl, nwg, 0xFFFFFFFFu << (gb & 31),
                                                             0xFFFFFFFFu >> (31 - ((gb + N - 1) & 31)), lane);
                        if (lane == 0) s_sum[gg * 4 + q] = sum;
                    }
                }
            }
            __syncthreads();
            // dense pass: Theta += (lr_p * delta) * Trace                          (multi_agent.py:28-30,42-43)
            for (int gg = 0; gg < G; ++gg) {
                if (rblock + gg >= p.R || s_dead[gg]) continue;
                const double d1 = __dadd_rn(s_sum[gg * 4 + 0], 0.0);            // + reward_shape (always 0)
                const double d3 = __dadd_rn(d1, __dmul_rn(p.discount, s_sum[gg * 4 + 1]));
                const double ld = __dmul_rn(p.lr_p, __dsub_rn(d3, s_sum[gg * 4 + 2]));
                double2 *th = reinterpret_cast<double2 *>(SMEM_TABLES ? s_tables + gg * tstride + p.tsize
                                                                      : p.Theta + (rblock + gg) * p.tsize);
                const double2 *tr = reinterpret_cast<const double2 *>(SMEM_TABLES ? s_tables + gg * tstride + 2 * p.tsize
                                                                                  : p.Trace + (rblock + gg) * p.tsize);
                const int n2 = (int)(p.tsize >> 1);
                for (int i = t; i < n2; i += blockDim.x) {
                    double2 a2 = th[i];
                    const double2 b2 = tr[i];
                    a2.x = __dadd_rn(a2.x, __dmul_rn(ld, b2.x));
                    a2.y = __dadd_rn(a2.y, __dmul_rn(ld, b2.y));
                    th[i] = a2;
                }
            }
        }
        // One commit group per step, closed here: group s holds the schedule chunks requested at the end of step s - 1 and the
        // record requested by the pop of step s, so when the top of step s + 2 waits for it every copy in it is two steps old.
        if (FUSED) cp_async_commit();
        if (!kInjectAtStart && (kStaticLive || (is_node && live))) {   // next step's new_packet + inject (env.py:298-319); slot = 1 here
            inject_step(s, (unsigned)(s + 1), clock + 1);
            if (!PAGED && tail - head > c_cap && (!kStaticLive || act)) set_status(p.status, r, RLR_STATUS_QUEUE_OVERFLOW, clock);
        }
        RLR_T(3);                                   // D
        __syncthreads();                                                        // ---- barrier 2 ----
        RLR_T(4);
        if (kRegRich ? (metric_bits & 1u) != 0u : (is_node && x == N - 1)) {    // the metric thread: 64-bit running route_time, per-step record
            run_rt += (long long)s_rt32[g];
            s_rt32[g] = 0u;                         // next deliveries come after the next barrier 1
        }
        if (sampling) {                             // Network.sample_route_time (env.py:416-438): file this step's deliveries
            if (is_node && x == N - 1 && !dead) {
                const int n = (int)s_nsamp[g];
                for (int i = 1; i < n; ++i) {       // insertion sort by heap pop position = the reference's delivery order
                    const unsigned kp = s_samp_pos[g * sstride + i];
                    const int kl = s_samp_lat[g * sstride + i];
                    int j = i - 1;
                    while (j >= 0 && s_samp_pos[g * sstride + j] > kp) {
                        s_samp_pos[g * sstride + j + 1] = s_samp_pos[g * sstride + j];
                        s_samp_lat[g * sstride + j + 1] = s_samp_lat[g * sstride + j];
                        --j;
                    }
                    s_samp_pos[g * sstride + j + 1] = kp;
                    s_samp_lat[g * sstride + j + 1] = kl;
                }
                int idx = (int)s_sidx[g];
                for (int i = 0; i < n; ++i, ++idx)
                    p.sample[r * (long long)p.sample_size + (idx < p.sample_size - 1 ? idx : p.sample_size - 1)] = s_samp_lat[g * sstride + i];
                s_sidx[g] = (unsigned)idx;
                s_nsamp[g] = 0u;
                // the reference finishes the iteration (all `freq` steps) in which the sample filled up, env.py:426-428
                if (idx >= p.sample_size && (s + 1) % freq == 0) { s_dead[g] = 2u; p.sample_clock[r] = clock + slotw; }
            }
            __syncthreads();
        }
        if (kRegRich ? (metric_bits & 2u) != 0u : (is_node && x == N - 1 && p.metrics)) {     // env.py:409-413 (cumulative counters)
            const long long rt = base_rt + run_rt;
            const unsigned en = (unsigned)(base_end + s_end[g]), hps = (unsigned)(base_hops + s_hops[g]);
            if (kRegRich) {
                *metric_row = make_uint4((unsigned)rt, (unsigned)((unsigned long long)rt >> 32), en, hps);
                metric_row += p.R;
            } else {
                p.metrics[(long long)s * p.R + r] = make_uint4((unsigned)rt, (unsigned)((unsigned long long)rt >> 32), en, hps);
            }
            if (SPLIT && p.metrics2)
                p.metrics2[(long long)s * p.R + r] = make_uint2((unsigned)(base_all + s_injs[g]), (unsigned)(base_drop + s_drop[g]));
        }
    }

#ifdef RLR_EXP_TIMING
    if (t == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(&g_phase_cycles[i], (unsigned long long)tacc_[i]);
#endif
    // ---- write back ----
    if (is_node) {
        p.qhead[r * N + x] = head;
        p.qtail[r * N + x] = tail;
        if (PAGED) {
            unsigned short *qp = p.q_pages + ((size_t)r * N + x) * 4;
            qp[0] = (unsigned short)hpage; qp[1] = (unsigned short)tpage; qp[2] = (unsigned short)spare;
        }
        if (general) p.q_dead[r * N + x] = (int)ndead;
        if (my_sends) atomicAdd(&s_sends[g], (unsigned long long)my_sends);
        if (my_inj) atomicAdd(&s_inj[g], (unsigned long long)my_inj);
    }
    __syncthreads();
    if (PAGED && t < G && rblock + t < p.R) p.page_top[rblock + t] = s_ptop[t];
    if (is_node && x == N - 1) {
        long long *c = p.counters + r * RLR_NUM_COUNTERS;
        c[RLR_CTR_ALL_PACKETS] += (long long)s_inj[g];
        c[RLR_CTR_END_PACKETS] += (long long)s_end[g];
        c[RLR_CTR_DROP_PACKETS] += (long long)s_drop[g];
        c[RLR_CTR_HOPS] += (long long)s_hops[g];
        c[RLR_CTR_ROUTE_TIME] += run_rt;
        if (sampling) p.sample_idx[r] = (int)s_sidx[g];
        c[RLR_CTR_SENDS] += (long long)s_sends[g];
    }
    if (general) {
        for (int i = t; i < G * ecap; i += blockDim.x) {
            const int gg = i / ecap1, j = i - gg * ecap1;
            if (rblock + gg < p.R && j < s_evlen[2 * gg]) {
                const unsigned e = s_evheap[gg * hstride + 1 + j], link = e & 0xFFFFu;
                const unsigned at = (unsigned)p.clock0 + (e >> 16);
                p.event_heap[(rblock + gg) * ecap + j] = ((unsigned long long)at << 32) | (link << 16) |
                                                         ((at % (unsigned)(ttime + 1)) * N + s_evfrom[link]);
            }
        }
        for (int i = t; i < G * p.sdeg; i += blockDim.x) {
            const long long rr = rblock + i / p.sdeg;
            if (rr < p.R) p.link_sent[rr * p.sdeg + i % p.sdeg] = s_evsent[i];
        }
        for (int i = t; i < G; i += blockDim.x)
            if (rblock + i < p.R) p.event_count[rblock + i] = s_evlen[2 * i];
    }
    if (POLICY == RLR_POLICY_GLOBALROUTE && do_learn) {
        for (int i = t; i < G * N * N; i += blockDim.x) {
            const long long rr = rblock + i / (N * N);
            if (rr < p.R) {
                p.gr_choice_mask[rr * N * N + i % (N * N)] = reinterpret_cast<const unsigned short *>(s_nexthop)[i];
                p.gr_distance[rr * N * N + i % (N * N)] = s_grdist[i];
            }
        }
    }
    if (SMEM_TABLES) {
        // tables go back by TMA bulk store: every thread publishes its shared-memory writes to the async proxy, then one
        // thread issues a copy per table and waits for the group before the CTA (and its shared memory) goes away
        fence_proxy_async();
        __syncthreads();
        if (t == 0) {
            const unsigned tbytes = (unsigned)(p.tsize * sizeof(QT));
            for (int gg = 0; gg < G; ++gg) {
                if (rblock + gg >= p.R) break;
                for (int tb = 0; tb < p.ntab; ++tb)
                    bulk_s2g(reinterpret_cast<char *>(tb == 0 ? p.Q : tb == 1 ? p.Theta : p.Trace) + (size_t)(rblock + gg) * tbytes,
                             reinterpret_cast<const char *>(s_tables) + ((size_t)gg * p.ntab + tb) * tbytes, tbytes);
            }
            bulk_commit_wait();
        }
    }
}

using StepKernel = void (*)(const StepParams);

template <int POLICY, int MAXD, int MAXW, bool SPLIT, int LB = 0, typename QT = double, bool PAGED = false>
StepKernel pick_smem(bool smem_tables)
{
    if (POLICY == RLR_POLICY_SHORTEST || POLICY == RLR_POLICY_GLOBALROUTE || LB != 0)
        return (StepKernel)rlr_step_kernel<POLICY, MAXD, MAXW, false, SPLIT, LB, QT, PAGED>;
    return smem_tables ? (StepKernel)rlr_step_kernel<POLICY, MAXD, MAXW, true, SPLIT, 0, QT, PAGED>
                       : (StepKernel)rlr_step_kernel<POLICY, MAXD, MAXW, false, SPLIT, 0, QT, PAGED>;
}

template <int POLICY, bool SPLIT, typename QT = double, bool PAGED = false>
StepKernel pick_deg(int n_nodes, int max_deg, bool smem_tables, int threads)
{
    if (n_nodes <= 64) {
        if (max_deg <= 4) return pick_smem<POLICY, 4, 3, SPLIT, 0, QT, PAGED>(smem_tables);
        if (max_deg <= 8) return pick_smem<POLICY, 8, 3, SPLIT, 0, QT, PAGED>(smem_tables);
        return pick_smem<POLICY, 16, 3, SPLIT, 0, QT, PAGED>(smem_tables);
    }
    // small CTAs reading their tables from HBM (lata.net: 116 nodes, 128 threads): the high-occupancy build
    constexpr bool kHasLb1 = POLICY == RLR_POLICY_QROUTE || POLICY == RLR_POLICY_HYBRIDQ || POLICY == RLR_POLICY_SHORTEST;
    // (a CTA of at most 128 threads has four window words, so no replica in it spans more: MAXW = 4, one packed register of
    // prefix counts; with MAXW = 9 the window code was 29 % of the instructions lata.net's Qroute step executed)
    if (kHasLb1 && !SPLIT && !smem_tables && threads <= 128 && max_deg > 8 && max_deg <= 12)
        return pick_smem<POLICY, 12, 4, SPLIT, kHasLb1 && !SPLIT ? 1 : 0, QT, PAGED>(false);
    if (!SPLIT && n_nodes <= 128 && max_deg > 8 && max_deg <= 12)           // 65..128 nodes at any alignment: at most five words
        return pick_smem<POLICY, 12, 5, SPLIT, 0, QT, PAGED>(smem_tables);
    if (max_deg <= 8) return pick_smem<POLICY, 8, 9, SPLIT, 0, QT, PAGED>(smem_tables);
    if (max_deg <= 12) return pick_smem<POLICY, 12, 9, SPLIT, 0, QT, PAGED>(smem_tables);      // lata.net: degree 9 (row loops sized 12, not 16)
    return pick_smem<POLICY, 16, 9, SPLIT, 0, QT, PAGED>(smem_tables);
}

}  // namespace

// One definition per policy family (step_<policy>.cu): the instantiation serving (n_nodes, max_deg, tables in shared memory),
// fused Network.train loop (split = false) or the split-phase / options variant (split = true).
#define RLR_DEFINE_STEP_PICK(NAME, POLICY)                                                                        \
    void *rlr_pick_step_##NAME(bool split, int n_nodes, int max_deg, bool smem_tables, int threads, bool paged)    \
    {                                                                                                             \
        if (paged && !split) return (void *)pick_deg<POLICY, false, double, true>(n_nodes, max_deg, smem_tables, threads);   \
        return split ? (void *)pick_deg<POLICY, true>(n_nodes, max_deg, smem_tables, threads)                      \
                     : (void *)pick_deg<POLICY, false>(n_nodes, max_deg, smem_tables, threads);                    \
    }

// The same for Qroute with its table stored in fp32 (step_qroute_f32.cu).
#define RLR_DEFINE_STEP_PICK_F32(NAME, POLICY)                                                                    \
    void *rlr_pick_step_##NAME(bool split, int n_nodes, int max_deg, bool smem_tables, int threads, bool paged)    \
    {                                                                                                             \
        if (paged && !split) return (void *)pick_deg<POLICY, false, float, true>(n_nodes, max_deg, smem_tables, threads);    \
        return split ? (void *)pick_deg<POLICY, true, float>(n_nodes, max_deg, smem_tables, threads)               \
                     : (void *)pick_deg<POLICY, false, float>(n_nodes, max_deg, smem_tables, threads);             \
    }
