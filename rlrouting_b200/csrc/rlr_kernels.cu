// rlr_kernels.cu — sm_100a kernels + C ABI (include/rlrouting_b200.h) of the batched packet-routing
// simulator.  Written from scratch for B200; the reference (xuxife/rlrouting) is pure Python.
//
// Execution model
//   * one thread per (replica, node); a CTA owns G whole replicas and runs ALL n_steps of a launch,
//     so the send -> arrive -> learn ordering of Network.step (env.py:333-365) is two CTA barriers;
//   * per-node queue cursors and the record at the head of the queue live in registers for the
//     whole launch (the head record is prefetched one step ahead of its use);
//   * the per-replica tables (Q / Theta / Trace, fp64) are staged into shared memory when they fit
//     (6x6: 28.8 KB per table) so the per-hop row reads never leave the SM; otherwise (lata, fp64)
//     they stay in HBM and the same code reads them through L1/L2;
//   * arrivals are GATHERED by the receiving node from its neighbours' send slots and appended in
//     the heap-tie order of the reference, so no atomics touch the queues and the
//     result is deterministic;
//   * injections come from a precomputed per-node schedule (Philox kernel one chunk ahead on a side
//     stream, or the host's replay of NumPy's legacy stream), so no RNG sits on the step's critical path;
//   * the fused Network.train path and the rarely used options (split step/learn phases, send log,
//     is_drop, droprate records, sample_route_time) are separate instantiations (SPLIT), because the
//     fused loop is latency-bound and every extra test in it costs throughput.
// Tuning aids: -DRLR_FAST_BUILD (only the 6x6 Qroute instantiations, seconds to build) and
// -DRLR_EXP_TIMING (per-phase clock64 accounting read through rlr_debug_phase_cycles).
//
// All fp64 arithmetic on tables uses the explicitly rounded intrinsics (__dmul_rn/__dadd_rn/...),
// never FMA, because the reference evaluates every NumPy scalar operation separately.
#include "rlr_step.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define RLR_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(RLR_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));            \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Qroute.__init__ structure overrides (qroute.py:16-20), Theta = initP, Trace = 0.
// ------------------------------------------------------------------------------------------------
// conf_mode: Theta is CQ.confidence and receives CQ.clean's structure (qroute.py:66-71: 0, conf[links[x]] = eye).
template <typename QT>
__global__ void rlr_init_tables_kernel(QT *Q, double *Theta, double *Trace, const int *row_ptr, const int *col,
                                       int N, int sdeg, long long R, double initP, int conf_mode)
{
    const long long tsize = (long long)N * sdeg;
    const long long total = R * sdeg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / sdeg;
        const int e = (int)(i - r * sdeg);
        int x = 0;
        while (row_ptr[x + 1] <= e) ++x;          // small N: linear search
        const int rp = row_ptr[x], deg = row_ptr[x + 1] - rp, j = e - rp, y = col[e];
        if (Q) {
            QT *tab = Q + r * tsize + (long long)N * rp;
            // -eye: -1 on the diagonal and NEGATIVE zero off it (NumPy's -0.0).  The float table is written as bit patterns:
            // the compiler folds `cond ? -1.0f : -0.0f` into an integer-to-float conversion that loses the sign of zero
            for (int k = 0; k < deg; ++k) {
                if constexpr (sizeof(QT) == 4)
                    reinterpret_cast<unsigned *>(tab)[(long long)y * deg + k] = (j == k) ? 0xBF800000u : 0x80000000u;
                else
                    tab[(long long)y * deg + k] = (QT)((j == k) ? -1.0 : -0.0);
            }
            if (j == 0)
                for (int k = 0; k < deg; ++k) tab[(long long)x * deg + k] = (QT)0.0;              // table[x] = 0
        }
    }
    const long long tot2 = R * tsize;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tot2; i += (long long)gridDim.x * blockDim.x) {
        if (Theta) {
            double v = initP;
            if (conf_mode) {                       // element i = (r, x, d, k): 1 where d is x's k-th neighbour
                const long long e = i % tsize;
                int x = 0;
                while ((long long)N * row_ptr[x + 1] <= e) ++x;
                const int rp = row_ptr[x], deg = row_ptr[x + 1] - rp;
                const int dd = (int)((e - (long long)N * rp) / deg), k = (int)((e - (long long)N * rp) % deg);
                v = (col[rp + k] == dd) ? 1.0 : 0.0;
            }
            Theta[i] = v;
        }
        if (Trace) Trace[i] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Shortest.__init__ + _calc_distance (shortest.py:20-58), multiway=False.  Columns z of the
// distance matrix are mutually independent (distance[.,z] reads only distance[.,z]), so one thread
// replays the reference's sequential Gauss-Seidel (x, y in links[x]) sweeps for its destination z.
// ------------------------------------------------------------------------------------------------
__global__ void rlr_apsp_kernel(double *distance, unsigned char *choice, unsigned char *next_hop, unsigned short *choice_mask,
                                const int *row_ptr, const int *col, int N, int multiway)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= N) return;
    for (int x = 0; x < N; ++x) distance[x * N + z] = CUDART_INF;                  // shortest.py:24
    for (int x = 0; x < N; ++x) {                                                   // shortest.py:28-34
        const int rp = row_ptr[x], deg = row_ptr[x + 1] - rp;
        unsigned char *c = choice + (long long)N * rp + (long long)z * deg;
        for (int j = 0; j < deg; ++j) c[j] = 0;
        if (x == z) distance[x * N + z] = 0.0;
        for (int j = 0; j < deg; ++j)
            if (col[rp + j] == z) { distance[x * N + z] = 1.0; c[j] = 1; }
    }
    bool changing = true;                                                           // shortest.py:43-58
    while (changing) {
        changing = false;
        for (int x = 0; x < N; ++x) {
            const int rp = row_ptr[x], deg = row_ptr[x + 1] - rp;
            unsigned char *c = choice + (long long)N * rp + (long long)z * deg;
            for (int j = 0; j < deg; ++j) {
                const double nd = distance[col[rp + j] * N + z] + 1.0;
                if (distance[x * N + z] > nd) {
                    distance[x * N + z] = nd;
                    for (int q = 0; q < deg; ++q) c[q] = 0;
                    c[j] = 1;
                    changing = true;
                }
                if (multiway && nd < CUDART_INF && distance[x * N + z] == nd) c[j] = 1;         // shortest.py:57-58
            }
        }
    }
    for (int x = 0; x < N; ++x) {                                                   // choose(): first True
        const int rp = row_ptr[x], deg = row_ptr[x + 1] - rp;
        const unsigned char *c = choice + (long long)N * rp + (long long)z * deg;
        int k = 0;
        while (k < deg - 1 && !c[k]) ++k;
        next_hop[x * N + z] = (unsigned char)k;
        unsigned m = 0;
        for (int j = 0; j < deg; ++j) m |= (unsigned)(c[j] != 0) << j;
        if (choice_mask) choice_mask[x * N + z] = (unsigned short)m;
    }
}

// GlobalRoute.__init__ (shortest.py:85-91): after Shortest.__init__ the mask is widened to everything off the diagonal
// and the distances are recomputed from inf with queue_size = 0; every reachable (x, z) is strictly improved at least
// once in that pass, so its `choice` entry is the one this kernel computes.  One CTA per replica, thread z = column z.
__global__ void rlr_globalroute_init_kernel(unsigned short *gr_choice_mask, int *gr_distance, const int *qs_off,
                                            const int *row_ptr, const int *col, int N, int sdeg, int multiway)
{
    extern __shared__ __align__(16) unsigned char smem[];
    int *dist = reinterpret_cast<int *>(smem);
    int *w = dist + N * N;
    int *rowptr = w + N;
    unsigned short *scol = reinterpret_cast<unsigned short *>(rowptr + N + 1);
    unsigned short *mask = scol + sdeg;
    const long long r = blockIdx.x;
    for (int i = threadIdx.x; i <= N; i += blockDim.x) rowptr[i] = row_ptr[i];
    for (int i = threadIdx.x; i < sdeg; i += blockDim.x) scol[i] = (unsigned short)col[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) w[i] = 1 + qs_off[r * N + i];
    __syncthreads();
    for (int z = threadIdx.x; z < N; z += blockDim.x) gr_column(z, N, rowptr, scol, w, dist, mask, multiway != 0);
    __syncthreads();
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
        gr_choice_mask[r * N * N + i] = mask[i];
        gr_distance[r * N * N + i] = dist[i];
    }
}


// Paged queues after Network.reset: node x of every replica owns page x (head = tail page) and spare N + x; pages
// 2N .. n_pages-1 are on the replica's free stack.
__global__ void rlr_paged_reset_kernel(unsigned short *page_free, int *page_top, unsigned short *q_pages, long long R, int N, int n_pages)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= R * n_pages) return;
    const long long r = i / n_pages;
    const int j = (int)(i - r * n_pages);
    if (j < n_pages - 2 * N) page_free[i] = (unsigned short)(2 * N + j);
    if (j == 0) page_top[r] = n_pages - 2 * N;
    if (j < N) {
        unsigned short *qp = q_pages + (r * N + j) * 4;
        qp[0] = (unsigned short)j; qp[1] = (unsigned short)j; qp[2] = (unsigned short)(N + j); qp[3] = 0;
    }
}

__global__ void rlr_reduce_stats_kernel(const long long *counters, long long R, long long *out)
{
    __shared__ long long acc[RLR_NUM_COUNTERS];
    if (threadIdx.x < RLR_NUM_COUNTERS) acc[threadIdx.x] = 0;
    __syncthreads();
    long long loc[RLR_NUM_COUNTERS];
#pragma unroll
    for (int c = 0; c < RLR_NUM_COUNTERS; ++c) loc[c] = 0;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x)
#pragma unroll
        for (int c = 0; c < RLR_NUM_COUNTERS; ++c) loc[c] += counters[r * RLR_NUM_COUNTERS + c];
#pragma unroll
    for (int c = 0; c < RLR_NUM_COUNTERS; ++c) {
        long long v = loc[c];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd((unsigned long long *)&acc[c], (unsigned long long)v);
    }
    __syncthreads();
    if (threadIdx.x < RLR_NUM_COUNTERS && acc[threadIdx.x])
        atomicAdd((unsigned long long *)&out[threadIdx.x], (unsigned long long)acc[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// Injection schedule (Philox mode): Network.new_packet (env.py:298-312) for K steps, one thread per
// (replica, node).  Poisson splitting: node x draws n ~ Poisson(lambda/N) by NumPy's multiplication method
// (count uniforms until their product <= exp(-lambda/N)); the first uniform u0 = (w0 + 0.5) * 2^-32 gives
// n >= 1 iff w0 >= thr (threshold computed exactly on the host), so the common n == 0 case is one integer
// compare; each packet's destination is uniform over the other N-1 nodes.  Same word stream as the oracle.
// Output: per (replica, node) an ascending list of events (local step << 8 | dest), 0xFFFFFFFF-terminated.
// ------------------------------------------------------------------------------------------------
__global__ void rlr_inject_schedule_kernel(unsigned *events, int cap, long long R, int N, int K, int clock0, int freq, int slot,
                                           const double *enlam, const unsigned long long *thr, unsigned long long seed,
                                           long long replica_base, int *status)
{
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= R * N) return;
    const long long r = idx / N;
    const int x = (int)(idx - r * N);
    const double en = enlam[r];
    const unsigned long long th = thr[r];
    unsigned *ev = events + idx * cap;
    int n_ev = 0;
    bool overflow = false;
    // The common case — no packet at this node and step — is one Philox block and one compare.  Its ten round keys are
    // computed once per thread and laundered through an opaque move: left to itself the compiler keeps them on the uniform
    // datapath and re-derives most of them inside the loop (19 of the loop's 65 issue slots).
    const unsigned long long rep = (unsigned long long)(replica_base + r);
    const unsigned rlo = (unsigned)rep, rhi = (unsigned)(rep >> 32);
    unsigned kx[10], ky[10];
    kx[0] = (unsigned)seed; ky[0] = (unsigned)(seed >> 32);
#pragma unroll
    for (int i = 1; i < 10; ++i) { kx[i] = kx[i - 1] + 0x9E3779B9u; ky[i] = ky[i - 1] + 0xBB67AE85u; }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        asm volatile("mov.u32 %0, %0;" : "+r"(kx[i]));
        asm volatile("mov.u32 %0, %0;" : "+r"(ky[i]));
    }
    // One new_packet per `freq` steps, each step `slot` long (env.py:400-403).  A node injects in ~lambda/N of its steps, but
    // some lane of a warp does in most of them, so a loop that handles a hit where it finds it walks the slow path nearly
    // every step with one or two lanes active.  Instead a thread scans kBatch steps with the common-case block only, notes
    // its hits in a bit mask, and then takes the hits one by one: the slow path runs max-over-lanes(hits) times per batch.
    constexpr int kBatch = 32;
    for (int s0 = 0; s0 < K; s0 += kBatch * freq) {
        unsigned hits = 0u;
#pragma unroll 2
        for (int i = 0; i < kBatch; ++i) {
            const int s = s0 + i * freq;
            if (s >= K) break;
            uint4 c = make_uint4((unsigned)(clock0 + s * slot), (unsigned)x, rlo, rhi);     // block 0 of purpose 0
#pragma unroll
            for (int q = 0; q < 10; ++q) {
                const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
                const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
                c = make_uint4(hi1 ^ c.y ^ kx[q], lo1, hi0 ^ c.w ^ ky[q], lo0);
            }
            if ((unsigned long long)c.x >= th) hits |= 1u << i;
        }
        while (hits) {
            const int i = __ffs((int)hits) - 1;
            hits &= hits - 1u;
            const int s = s0 + i * freq;
            PhiloxStream ps(seed, rep, clock0 + s * slot, x, 0);
            const unsigned w0 = ps.next();              // the block of the scan again: only a hit pays for it twice
            int n = 1;
            double prod = __dmul_rn(__dadd_rn((double)w0, 0.5), 1.0 / 4294967296.0);
            for (;;) {
                const double u = __dmul_rn(__dadd_rn((double)ps.next(), 0.5), 1.0 / 4294967296.0);
                prod = __dmul_rn(prod, u);
                if (prod > en) ++n; else break;
            }
            for (int j = 0; j < n; ++j) {
                unsigned dd = __umulhi(ps.next(), (unsigned)(N - 1));
                if (dd >= (unsigned)x) ++dd;
                if (n_ev < cap - 1) ev[n_ev++] = ((unsigned)s << 8) | dd; else overflow = true;
            }
        }
    }
    ev[n_ev] = 0xFFFFFFFFu;
    if (overflow) set_status(status, r, RLR_STATUS_INJECT_OVERFLOW, clock0);
}

// ave_route_time / ave_hops after each step (env.py:440-446): route_time / end_packets if end_packets > 0 else 0.
// IEEE division of two exactly converted integers is Python's correctly rounded int / int.
__global__ void rlr_ratio_kernel(const uint4 *metrics, const uint2 *metrics2, long long n, double *route_time_out,
                                 double *hop_out, double *droprate_out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint4 m = metrics[i];
        const long long rt = (long long)(((unsigned long long)m.y << 32) | m.x);
        const double en = (double)m.z;
        if (route_time_out) route_time_out[i] = m.z ? __ddiv_rn((double)rt, en) : 0.0;
        if (hop_out) hop_out[i] = m.z ? __ddiv_rn((double)m.w, en) : 0.0;
        if (droprate_out) {                          // drop_packets / all_packets if all_packets > 0 else 0, env.py:448-450
            const uint2 m2 = metrics2[i];
            droprate_out[i] = m2.x ? __ddiv_rn((double)m2.y, (double)m2.x) : 0.0;
        }
    }
}

// Per-load-level curves (north star: only reduced statistics leave the device).  The step kernel leaves one cumulative
// record (route_time, end_packets, hops) per (step, replica); this kernel pools them over the replicas of each load level
// — integer sums, so the result does not depend on the order — and forms the level's ave_route_time / ave_hops
// (env.py:440-446) per step.  One CTA per step; replicas of a level are contiguous, so a warp normally adds one value.
__global__ void rlr_reduce_levels_kernel(const uint4 *metrics, long long R, const int *level_idx, int n_levels,
                                         double *route_time_out, double *hop_out, long long *sums_out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(smem);       // [n_levels][3]
    const long long s = blockIdx.x;
    for (int i = threadIdx.x; i < 3 * n_levels; i += blockDim.x) acc[i] = 0ull;
    __syncthreads();
    const uint4 *row = metrics + s * R;
    for (long long r0 = 0; r0 < R; r0 += blockDim.x) {
        const long long r = r0 + threadIdx.x;
        const bool ok = r < R;
        const int lv = ok ? level_idx[r] : -1;
        uint4 m = make_uint4(0u, 0u, 0u, 0u);
        if (ok) m = row[r];
        unsigned long long rt = ((unsigned long long)m.y << 32) | m.x, en = m.z, hp = m.w;
        const int lv0 = __shfl_sync(0xFFFFFFFFu, lv, 0);
        if (__all_sync(0xFFFFFFFFu, lv == lv0)) {                                  // the whole warp is one level
            for (int o = 16; o > 0; o >>= 1) {
                rt += __shfl_down_sync(0xFFFFFFFFu, rt, o);
                en += __shfl_down_sync(0xFFFFFFFFu, en, o);
                hp += __shfl_down_sync(0xFFFFFFFFu, hp, o);
            }
            if ((threadIdx.x & 31) == 0 && lv0 >= 0 && lv0 < n_levels) {
                atomicAdd(&acc[3 * lv0], rt); atomicAdd(&acc[3 * lv0 + 1], en); atomicAdd(&acc[3 * lv0 + 2], hp);
            }
        } else if (ok && lv >= 0 && lv < n_levels) {
            atomicAdd(&acc[3 * lv], rt); atomicAdd(&acc[3 * lv + 1], en); atomicAdd(&acc[3 * lv + 2], hp);
        }
    }
    __syncthreads();
    for (int l = threadIdx.x; l < n_levels; l += blockDim.x) {
        const long long rt = (long long)acc[3 * l], en = (long long)acc[3 * l + 1], hp = (long long)acc[3 * l + 2];
        if (route_time_out) route_time_out[s * n_levels + l] = en ? __ddiv_rn((double)rt, (double)en) : 0.0;
        if (hop_out) hop_out[s * n_levels + l] = en ? __ddiv_rn((double)hp, (double)en) : 0.0;
        if (sums_out) { sums_out[(s * n_levels + l) * 3] = rt; sums_out[(s * n_levels + l) * 3 + 1] = en; sums_out[(s * n_levels + l) * 3 + 2] = hp; }
    }
}

// Q-convergence statistics (north star: "final allreduce of delivery-time and Q-convergence statistics"): per replica the
// sum of a table's entries and, against a previous copy of it, the sum of |table - prev| — what the TD updates of
// qroute.py:40-44 moved since the copy was taken.  One warp per replica, lane-strided partial sums combined by a fixed
// shuffle tree, so the fp64 result is reproducible; a second pass pools the replicas of a load level in index order.
template <typename QT>
__global__ void rlr_table_stats_kernel(const QT *table, const QT *prev, long long R, long long tsize, double *per_replica)
{
    const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const QT *t = table + r * tsize, *q = prev ? prev + r * tsize : nullptr;
    double s = 0.0, d = 0.0;
    for (long long i = lane; i < tsize; i += 32) {
        const double v = t[i];
        s = __dadd_rn(s, v);
        if (q) d = __dadd_rn(d, fabs(__dsub_rn(v, (double)q[i])));
    }
    for (int o = 16; o > 0; o >>= 1) {
        s = __dadd_rn(s, __shfl_down_sync(0xFFFFFFFFu, s, o));
        d = __dadd_rn(d, __shfl_down_sync(0xFFFFFFFFu, d, o));
    }
    if (lane == 0) { per_replica[2 * r] = s; per_replica[2 * r + 1] = d; }
}

__global__ void rlr_level_sums_kernel(const double *per_replica, long long R, const int *level_idx, int n_levels, double *out)
{
    const int l = blockIdx.x, lane = threadIdx.x;            // one warp per level, replicas in index order
    if (l >= n_levels) return;
    double s = 0.0, d = 0.0;
    for (long long r0 = 0; r0 < R; r0 += 32) {
        const long long r = r0 + lane;
        const bool mine = r < R && (level_idx ? level_idx[r] == l : true);
        if (mine) { s = __dadd_rn(s, per_replica[2 * r]); d = __dadd_rn(d, per_replica[2 * r + 1]); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s = __dadd_rn(s, __shfl_down_sync(0xFFFFFFFFu, s, o));
        d = __dadd_rn(d, __shfl_down_sync(0xFFFFFFFFu, d, o));
    }
    if (lane == 0) { out[2 * l] = s; out[2 * l + 1] = d; }
}

}  // namespace

// the per-policy instantiations live in step_<policy>.cu (RLR_DEFINE_STEP_PICK); a -DRLR_FAST_BUILD tuning build is one
// translation unit holding only the degree-4 instantiations of one policy (-DRLR_FAST_POLICY=<id>, Qroute by default)
#ifdef RLR_FAST_BUILD
#ifndef RLR_FAST_POLICY
#define RLR_FAST_POLICY RLR_POLICY_QROUTE
#endif
static void *rlr_pick_fast(bool split, bool smem_tables)
{
    return split ? (void *)pick_smem<RLR_FAST_POLICY, 4, 3, true>(smem_tables) : (void *)pick_smem<RLR_FAST_POLICY, 4, 3, false>(smem_tables);
}
#else
void *rlr_pick_step_shortest(bool, int, int, bool, int, bool);
void *rlr_pick_step_qroute(bool, int, int, bool, int, bool);
void *rlr_pick_step_qroute_f32(bool, int, int, bool, int, bool);
void *rlr_pick_step_hybridq(bool, int, int, bool, int, bool);
void *rlr_pick_step_mahybridq(bool, int, int, bool, int, bool);
void *rlr_pick_step_cq(bool, int, int, bool, int, bool);
void *rlr_pick_step_cdrq(bool, int, int, bool, int, bool);
void *rlr_pick_step_globalroute(bool, int, int, bool, int, bool);
#endif

namespace {
template <bool SPLIT>
StepKernel pick_kernel(int policy, int n_nodes, int max_deg, bool smem_tables, int threads, bool f32, bool paged)
{
#ifdef RLR_FAST_BUILD
    (void)policy; (void)n_nodes; (void)max_deg; (void)threads; (void)f32; (void)paged;
    return (StepKernel)rlr_pick_fast(SPLIT, smem_tables);
#else
    switch (policy) {
    case RLR_POLICY_SHORTEST: return (StepKernel)rlr_pick_step_shortest(SPLIT, n_nodes, max_deg, false, threads, paged);
    case RLR_POLICY_QROUTE: return f32 ? (StepKernel)rlr_pick_step_qroute_f32(SPLIT, n_nodes, max_deg, smem_tables, threads, paged)
                                       : (StepKernel)rlr_pick_step_qroute(SPLIT, n_nodes, max_deg, smem_tables, threads, paged);
    case RLR_POLICY_HYBRIDQ: return (StepKernel)rlr_pick_step_hybridq(SPLIT, n_nodes, max_deg, smem_tables, threads, paged);
    case RLR_POLICY_MAHYBRIDQ: return (StepKernel)rlr_pick_step_mahybridq(SPLIT, n_nodes, max_deg, smem_tables, threads, paged);
    case RLR_POLICY_CQ: return (StepKernel)rlr_pick_step_cq(SPLIT, n_nodes, max_deg, smem_tables, threads, paged);
    case RLR_POLICY_GLOBALROUTE: return (StepKernel)rlr_pick_step_globalroute(SPLIT, n_nodes, max_deg, false, threads, paged);
    default: return (StepKernel)rlr_pick_step_cdrq(SPLIT, n_nodes, max_deg, smem_tables, threads, paged);
    }
#endif
}
}  // namespace


// ------------------------------------------------------------------------------------------------
// Host-side replay of NumPy's legacy RandomState stream for Network.new_packet (env.py:298-312), so that trace
// mode scales to tens of thousands of replicas.  Restates, for the calls the reference makes:
//   np.random.poisson(lam), lam < 10  -> legacy random_poisson_mult: count doubles until their product <= exp(-lam)
//   np.random.randint(0, N[, size=2]) -> masked rejection on 32-bit outputs (mask = smallest 2^b - 1 >= N - 1)
//   doubles                           -> (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53
// on the MT19937 state exported by RandomState.get_state().  Checked against NumPy in tests/test_host_logic.py.
// ------------------------------------------------------------------------------------------------
namespace {
struct Mt19937 {
    uint32_t *key;
    int pos;
    void refill()
    {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7FFFFFFFu, MATRIX = 0x9908B0DFu;
        int i;
        for (i = 0; i < 624 - 397; ++i) {
            const uint32_t y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
        }
        for (; i < 623; ++i) {
            const uint32_t y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
        }
        const uint32_t y = (key[623] & UPPER) | (key[0] & LOWER);
        key[623] = key[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX : 0u);
        pos = 0;
    }
    uint32_t next()
    {
        if (pos == 624) refill();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9D2C5680u;
        y ^= (y << 15) & 0xEFC60000u;
        y ^= (y >> 18);
        return y;
    }
    double next_double()
    {
        const int32_t a = (int32_t)(next() >> 5), b = (int32_t)(next() >> 6);
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    uint32_t bounded(uint32_t rng, uint32_t mask)
    {
        uint32_t v;
        do { v = next() & mask; } while (v > rng);
        return v;
    }
};
}  // namespace

#ifdef RLR_EXP_TIMING
extern "C" int rlr_debug_phase_cycles(unsigned long long *out, int reset)
{
    cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
    return 0;
}
#endif

extern "C" int rlr_trace_replay(uint32_t *mt_key, int32_t *mt_pos, int32_t n_nodes, int64_t n_steps, double lam,
                                int32_t *cnt_out, uint32_t *pairs_out, int64_t pairs_cap, int64_t *n_pairs_out)
{
    if (!mt_key || !mt_pos || !cnt_out || !pairs_out || !n_pairs_out) return fail(RLR_E_INVALID, "rlr_trace_replay: null argument");
    if (n_nodes < 2 || n_nodes > 65535) return fail(RLR_E_INVALID, "rlr_trace_replay: n_nodes out of range");
    if (!(lam >= 0.0) || lam >= 10.0) return fail(RLR_E_UNSUPPORTED, "rlr_trace_replay: needs 0 <= lambda < 10 (NumPy switches to PTRS at 10)");
    Mt19937 mt{mt_key, *mt_pos};
    const uint32_t rng = (uint32_t)n_nodes - 1u;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    const double enlam = exp(-lam);
    int64_t np = 0;
    for (int64_t s = 0; s < n_steps; ++s) {
        int32_t n = 0;
        if (lam != 0.0) {                             // legacy_random_poisson: lam == 0 returns 0 without drawing
            double prod = 1.0;
            for (;;) {
                prod *= mt.next_double();
                if (prod > enlam) ++n; else break;
            }
        }
        cnt_out[s] = n;
        for (int32_t i = 0; i < n; ++i) {
            const uint32_t src = mt.bounded(rng, mask);
            uint32_t dst = mt.bounded(rng, mask);
            while (dst == src) dst = mt.bounded(rng, mask);
            if (np >= pairs_cap) return fail(RLR_E_INVALID, "rlr_trace_replay: pairs buffer too small");
            pairs_out[np++] = src | (dst << 16);
        }
    }
    *mt_pos = mt.pos;
    *n_pairs_out = np;
    return RLR_OK;
}

// The same replay for R independent streams on `nthreads` host threads (replica r uses lam[r], writes cnt_out[r][*]
// and at most pairs_cap entries of pairs_out[r][*]).
extern "C" int rlr_trace_replay_many(uint32_t *mt_keys, int32_t *mt_pos, int64_t n_streams, int32_t n_nodes, int64_t n_steps,
                                     const double *lam, int32_t *cnt_out, uint32_t *pairs_out, int64_t pairs_cap,
                                     int64_t *n_pairs_out, int32_t nthreads)
{
    if (!mt_keys || !mt_pos || !lam || !cnt_out || !pairs_out || !n_pairs_out || n_streams < 0)
        return fail(RLR_E_INVALID, "rlr_trace_replay_many: null argument");
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next(0);
    std::atomic<int> rc(RLR_OK);
    std::string msg;
    auto work = [&]() {
        for (;;) {
            const int64_t r = next.fetch_add(1);
            if (r >= n_streams) return;
            const int one = rlr_trace_replay(mt_keys + r * 624, mt_pos + r, n_nodes, n_steps, lam[r], cnt_out + r * n_steps,
                                             pairs_out + r * pairs_cap, pairs_cap, n_pairs_out + r);
            if (one != RLR_OK) { int expect = RLR_OK; if (rc.compare_exchange_strong(expect, one)) msg = g_last_error; }
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < nthreads; ++i) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (rc.load() != RLR_OK) return fail(rc.load(), msg.empty() ? "rlr_trace_replay_many: a stream failed" : msg);
    return RLR_OK;
}

// ------------------------------------------------------------------------------------------------
// Host side of the C ABI
// ------------------------------------------------------------------------------------------------
// Per-node injection event lists (the layout rlr_run_params_t.inj_events takes) from replayed traces, for the iterations
// [it0, it1) of every stream: an event is ((it - it0) * freq) << 8 | dest, in (node, iteration, draw) order, 0xFFFFFFFF ends
// a list.  With events == NULL only *cap_out = largest list length + 2 is computed.
extern "C" int rlr_trace_schedule(const int32_t *cnt, const uint32_t *pairs, int64_t pairs_cap, int64_t n_streams, int32_t n_nodes,
                                  int64_t n_steps, int64_t it0, int64_t it1, int32_t freq, uint32_t *events, int32_t cap,
                                  int32_t *cap_out, int32_t nthreads)
{
    if (!cnt || !pairs || n_streams < 0 || n_nodes < 1 || n_nodes > 255 || it0 < 0 || it1 > n_steps || it0 > it1 || freq < 1)
        return fail(RLR_E_INVALID, "rlr_trace_schedule: invalid argument");
    if (events && cap < 2) return fail(RLR_E_INVALID, "rlr_trace_schedule: cap must be >= 2");
    if ((it1 - it0) * freq >= (1 << 24)) return fail(RLR_E_INVALID, "rlr_trace_schedule: too many steps for the 24-bit step field");
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next(0);
    std::atomic<int> worst(0), bad(0);
    auto work = [&]() {
        std::vector<int32_t> fill((size_t)n_nodes);
        int local_max = 0;
        for (;;) {
            const int64_t r = next.fetch_add(1);
            if (r >= n_streams) break;
            const int32_t *c = cnt + r * n_steps;
            const uint32_t *pk = pairs + r * pairs_cap;
            int64_t off = 0;
            for (int64_t i = 0; i < it0; ++i) off += c[i];
            std::fill(fill.begin(), fill.end(), 0);
            uint32_t *ev = events ? events + (size_t)r * n_nodes * cap : nullptr;
            for (int64_t i = it0; i < it1; ++i) {
                for (int32_t j = 0; j < c[i]; ++j, ++off) {
                    if (off >= pairs_cap) { bad.store(1); break; }
                    const uint32_t src = pk[off] & 0xFFFFu, dst = pk[off] >> 16;
                    if (src >= (uint32_t)n_nodes) { bad.store(1); continue; }
                    int32_t &f = fill[src];
                    if (ev) { if (f < cap - 1) ev[(size_t)src * cap + f] = ((uint32_t)((i - it0) * freq) << 8) | dst; else bad.store(2); }
                    ++f;
                }
            }
            for (int32_t x = 0; x < n_nodes; ++x) {
                if (fill[x] > local_max) local_max = fill[x];
                if (ev) for (int32_t f = fill[x] < cap ? fill[x] : cap - 1; f < cap; ++f) ev[(size_t)x * cap + f] = 0xFFFFFFFFu;
            }
        }
        int cur = worst.load();
        while (local_max > cur && !worst.compare_exchange_weak(cur, local_max)) {}
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < nthreads; ++i) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (bad.load() == 1) return fail(RLR_E_INVALID, "rlr_trace_schedule: trace arrays are inconsistent (pairs_cap / node ids)");
    if (bad.load() == 2) return fail(RLR_E_INVALID, "rlr_trace_schedule: cap is smaller than the longest list + 2");
    if (cap_out) *cap_out = worst.load() + 2;
    return RLR_OK;
}

struct rlr_handle {
    int N, sdeg, max_deg, policy, ntab;
    long long R, replica_base, tsize;
    int cap, G, threads, smem_bytes, smem_bytes_split, device, ecap;   // dynamic shared memory of the fused / split kernel
    int esize;                           // bytes per Q-table entry: 8, or 4 for Qroute's fp32 storage mode
    int page_log, n_pages;               // paged queues: log2(records per page) (0 = rings), pages per replica
    bool smem_tables, bound;
    std::vector<int> row_ptr, col;
    rlr_buffers_t bufs;
    StepKernel kernel, kernel_split;     // fused Network.train path / Network.step + agent.learn phases
};

extern "C" {

int rlr_abi_version(void) { return RLR_ABI_VERSION; }

const char *rlr_last_error_string(void) { return g_last_error.c_str(); }

int rlr_struct_sizes(int32_t out[4])
{
    if (!out) return fail(RLR_E_INVALID, "rlr_struct_sizes: null output");
    out[0] = (int32_t)sizeof(rlr_topology_t);
    out[1] = (int32_t)sizeof(rlr_config_t);
    out[2] = (int32_t)sizeof(rlr_buffers_t);
    out[3] = (int32_t)sizeof(rlr_run_params_t);
    return RLR_OK;
}

int rlr_create(const rlr_topology_t *topo, const rlr_config_t *cfg, rlr_handle_t **out)
{
    if (!topo || !cfg || !out) return fail(RLR_E_INVALID, "rlr_create: null argument");
    const int N = topo->n_nodes;
    if (N < 2 || N > 255) return fail(RLR_E_UNSUPPORTED, "rlr_create: n_nodes must be in [2, 255]");
    if (topo->max_deg < 1 || topo->max_deg > 16) return fail(RLR_E_UNSUPPORTED, "rlr_create: max_deg must be in [1, 16]");
    if (!topo->row_ptr || !topo->col || topo->row_ptr[0] != 0 || topo->row_ptr[N] != topo->sum_deg)
        return fail(RLR_E_INVALID, "rlr_create: malformed CSR");
    if (topo->sum_deg % 2) return fail(RLR_E_INVALID, "rlr_create: sum_deg must be even (undirected graph)");
    for (int x = 0; x < N; ++x) {
        const int dg = topo->row_ptr[x + 1] - topo->row_ptr[x];
        if (dg < 1 || dg > topo->max_deg) return fail(RLR_E_INVALID, "rlr_create: node degree outside [1, max_deg]");
    }
    if (cfg->replicas < 1) return fail(RLR_E_INVALID, "rlr_create: replicas must be >= 1");
    if (cfg->policy < RLR_POLICY_SHORTEST || cfg->policy > RLR_POLICY_GLOBALROUTE)
        return fail(RLR_E_UNSUPPORTED, "rlr_create: unknown policy id");
    const int cap = cfg->queue_capacity;
    if (cap < 2 || (cap & (cap - 1))) return fail(RLR_E_INVALID, "rlr_create: queue_capacity must be a power of two >= 2");
    int page_log = 0;
    if (cfg->page_records != 0) {
        const int pr = cfg->page_records;
        if (pr < 4 || pr > 4096 || (pr & (pr - 1))) return fail(RLR_E_INVALID, "rlr_create: page_records must be 0 or a power of two in [4, 4096]");
        while ((1 << page_log) < pr) ++page_log;
        if (cfg->pages_per_replica < 2 * N + 1 || cfg->pages_per_replica > 65535)
            return fail(RLR_E_INVALID, "rlr_create: pages_per_replica must be in [2 * n_nodes + 1, 65535]");
        if (cfg->event_capacity > 0) return fail(RLR_E_UNSUPPORTED, "rlr_create: paged queues do not serve the general timing model");
    }

    RLR_CUDA(cudaSetDevice(cfg->device));
    if (cfg->event_capacity < 0 || cfg->event_capacity > 65535 || topo->sum_deg > 65535)
        return fail(RLR_E_UNSUPPORTED, "rlr_create: event_capacity must be in [0, 65535] (N * (transtime + 1)) and sum_deg <= 65535");
    rlr_handle *h = new rlr_handle();
    h->N = N; h->sdeg = topo->sum_deg; h->max_deg = topo->max_deg; h->policy = cfg->policy;
    // tables per replica: Shortest 0, Qroute 1 (Q), HybridQ 2 (+Theta), MaHybridQ 3 (+Trace), CQ/CDRQ 2 (Q + confidence)
    h->ntab = cfg->policy == RLR_POLICY_GLOBALROUTE ? 0 : cfg->policy >= RLR_POLICY_CQ ? 2 : cfg->policy;
    h->R = cfg->replicas; h->replica_base = cfg->replica_base;
    h->tsize = (long long)N * topo->sum_deg;
    h->cap = cap; h->bound = false; h->device = cfg->device; h->ecap = cfg->event_capacity;
    h->esize = cfg->table_dtype == 1 ? 4 : 8;
    h->page_log = page_log; h->n_pages = page_log ? cfg->pages_per_replica : 0;
    if (cfg->table_dtype != 0 && cfg->table_dtype != 1) { delete h; return fail(RLR_E_INVALID, "rlr_create: table_dtype must be 0 (fp64) or 1 (fp32)"); }
    if (h->esize == 4 && cfg->policy != RLR_POLICY_QROUTE) {
        delete h;
        return fail(RLR_E_UNSUPPORTED, "rlr_create: fp32 table storage is implemented for Qroute only");
    }
#ifdef RLR_FAST_BUILD
    if (h->esize == 4) { delete h; return fail(RLR_E_UNSUPPORTED, "rlr_create: this tuning build has no fp32 instantiations"); }
#endif
    h->row_ptr.assign(topo->row_ptr, topo->row_ptr + N + 1);
    h->col.assign(topo->col, topo->col + topo->sum_deg);
    memset(&h->bufs, 0, sizeof(h->bufs));

    // launch geometry
    const size_t table_bytes = (size_t)h->ntab * h->tsize * h->esize;
    int G = cfg->replicas_per_block;
    int smem_tables = cfg->tables_in_smem;
    if (h->policy == RLR_POLICY_SHORTEST || h->policy == RLR_POLICY_GLOBALROUTE) smem_tables = 0;
    // footprint of a launch: the larger of the two instantiations (fused Network.train / split phases and options)
    auto smem_for = [&](int g, int thr, bool st) {
        const size_t a = SmemLayout(N, h->sdeg, g, thr, h->policy, st, h->tsize, h->ntab, 0, true, h->esize).total;
        const size_t b = SmemLayout(N, h->sdeg, g, thr, h->policy, st, h->tsize, h->ntab, h->ecap, false, h->esize).total;
        return a > b ? a : b;
    };
    // General timing model: one thread per replica replays the event heap serially and bounds the step, so what pays is
    // more resident replicas, not faster table reads: tables stay in HBM/L2 and a CTA carries ~256 threads (measured).
    if (smem_tables < 0 && h->ecap > 0) smem_tables = 0;
    // the bulk copies that stage a table move multiples of 16 bytes
    if (table_bytes % 16 != 0) {
        if (smem_tables > 0) { delete h; return fail(RLR_E_UNSUPPORTED, "rlr_create: a table in shared memory must be a multiple of 16 bytes"); }
        smem_tables = 0;
    }
    // HybridQ's step is dominated by the softmax arithmetic, not by table reads, and its two fp64 tables would leave room
    // for only 3 replicas per SM: tables in HBM/L2 and 256-thread CTAs of several replicas run 14 per SM (measured on
    // 6x6, 4,096 replicas: 3.6e9 -> 8.0e9 hops/s).  MaHybridQ and CQ / CDRQ keep theirs on chip: they sweep whole tables
    // every step.
    const bool hybrid_hbm = smem_tables < 0 && h->policy == RLR_POLICY_HYBRIDQ;
    if (hybrid_hbm) smem_tables = 0;
    if (smem_tables < 0) {                               // automatic: in shared memory if they fit (for the requested G, if any)
        const int g_try = G > 0 ? G : 1;
        const int thr_try = cfg->threads_per_block > 0 ? cfg->threads_per_block : (g_try * N + 31) / 32 * 32 + 128;
        smem_tables = (table_bytes > 0 && smem_for(g_try, thr_try, true) <= (size_t)kSmemLimit) ? 1 : 0;
    }
    if (G <= 0 && h->ecap > 0 && !smem_tables) G = 256 / N > 0 ? 256 / N : 1;
    // GlobalRoute's per-step all-pairs recomputation is one serial thread per destination column: a 256-thread CTA of
    // several replicas keeps more columns going per SM (measured on 6x6: 1.9e8 -> 2.3e8 hops/s)
    if (G <= 0 && ((h->policy == RLR_POLICY_GLOBALROUTE && N < 96) || hybrid_hbm)) G = 256 / N > 0 ? 256 / N : 1;
    // fp32 storage: the tables are no longer what limits residency, so whole replicas are packed into 128-thread CTAs and
    // fewer, fuller warps issue the same work (measured on 6x6: 2.9e10 with one replica per CTA -> 3.8e10 hops/s with three)
    if (G <= 0 && h->esize == 4 && smem_tables && N <= 64) G = 128 / N > 0 ? 128 / N : 1;
    if (G <= 0) G = smem_tables ? 1 : (N >= 96 ? 1 : (128 / N > 0 ? 128 / N : 1));
    if ((long long)G > h->R) G = (int)h->R;
    int threads = cfg->threads_per_block;
    if (threads <= 0) {
        threads = ((G * N + 31) / 32) * 32;
        if (h->policy == RLR_POLICY_MAHYBRIDQ) {         // + helper warps for the trace pass and the three reductions
            threads = ((G * N + 31) / 32) * 32 + 128;
            if (threads < 256) threads = 256;
            if (threads > kMaxThreads) threads = kMaxThreads;
        }
        if (h->policy == RLR_POLICY_CQ || h->policy == RLR_POLICY_CDRQ) {
            // the confidence decay is one short dense pass a step: two more warps are enough, and 128-thread CTAs let a
            // third one fit in the register file (measured on 6x6: CQ 4.6e9 -> 6.3e9, CDRQ 3.6e9 -> 4.7e9 hops/s)
            threads = ((G * N + 31) / 32) * 32 + 64;
            if (threads > kMaxThreads) threads = kMaxThreads;
        }
    }
    if (h->policy == RLR_POLICY_MAHYBRIDQ && (N > 127 || threads < ((G * N + 31) / 32) * 32 + 96)) {
        delete h;
        return fail(RLR_E_UNSUPPORTED, "rlr_create: MaHybridQ needs n_nodes <= 127 and threads_per_block >= roundup32(G*N) + 96");
    }
    if (threads % 32 || threads < G * N || threads > kMaxThreads) {
        delete h;
        return fail(RLR_E_INVALID, "rlr_create: threads_per_block must be a multiple of 32 in [G*N, 512]");
    }
    const size_t smem = smem_for(G, threads, smem_tables != 0);
    if (smem > (size_t)kSmemLimit) {
        delete h;
        return fail(RLR_E_UNSUPPORTED, "rlr_create: shared-memory footprint exceeds 227 KB (lower replicas_per_block or set tables_in_smem=0)");
    }
    h->G = G; h->threads = threads; h->smem_tables = smem_tables != 0;
    h->smem_bytes = (int)SmemLayout(N, h->sdeg, G, threads, h->policy, h->smem_tables, h->tsize, h->ntab, 0, true, h->esize).total;
    h->smem_bytes_split = (int)SmemLayout(N, h->sdeg, G, threads, h->policy, h->smem_tables, h->tsize, h->ntab, h->ecap, false, h->esize).total;
    h->kernel = pick_kernel<false>(h->policy, h->N, h->max_deg, h->smem_tables, h->threads, h->esize == 4, h->page_log != 0);
    h->kernel_split = pick_kernel<true>(h->policy, h->N, h->max_deg, h->smem_tables, h->threads, h->esize == 4, false);
    (void)smem;                 // (the per-function shared-memory limit is set right before every launch, rlr_run_steps)
    *out = h;
    return RLR_OK;
}

int rlr_destroy(rlr_handle_t *h)
{
    delete h;
    return RLR_OK;
}

int rlr_bind_buffers(rlr_handle_t *h, const rlr_buffers_t *b)
{
    if (!h || !b) return fail(RLR_E_INVALID, "rlr_bind_buffers: null argument");
    if ((!b->ring && !h->page_log) || !b->q_head || !b->q_tail || !b->counters || !b->status || !b->topo_row_ptr || !b->topo_col || !b->heap_pos)
        return fail(RLR_E_INVALID, "rlr_bind_buffers: ring, q_head, q_tail, counters, status, topo_* and heap_pos are required");
    if (h->page_log && (!b->pool || !b->page_next || !b->page_free || !b->page_top || !b->q_pages || ((uintptr_t)b->pool & 15)))
        return fail(RLR_E_INVALID, "rlr_bind_buffers: paged queues need pool (16-byte aligned), page_next, page_free, page_top and q_pages");
    if (h->policy == RLR_POLICY_SHORTEST && (!b->next_hop || !b->distance || !b->choice))
        return fail(RLR_E_INVALID, "rlr_bind_buffers: Shortest needs next_hop, distance and choice");
    const bool gr = h->policy == RLR_POLICY_GLOBALROUTE;
    if (gr && (!b->gr_choice_mask || !b->gr_distance || !b->gr_qs_off))
        return fail(RLR_E_INVALID, "rlr_bind_buffers: GlobalRoute needs gr_choice_mask, gr_distance and gr_qs_off");
    if (!gr && h->policy >= RLR_POLICY_QROUTE && !b->q_table) return fail(RLR_E_INVALID, "rlr_bind_buffers: q_table required");
    if (!gr && h->policy >= RLR_POLICY_HYBRIDQ && !b->theta) return fail(RLR_E_INVALID, "rlr_bind_buffers: theta required");
    if (h->policy == RLR_POLICY_MAHYBRIDQ && !b->trace) return fail(RLR_E_INVALID, "rlr_bind_buffers: trace required");
    if (h->ecap > 0 && (!b->link_sent || !b->event_heap || !b->event_count || !b->event_rec || !b->q_dead))
        return fail(RLR_E_INVALID, "rlr_bind_buffers: event_capacity > 0 needs link_sent, event_heap, event_count, event_rec and q_dead");
    if (((uintptr_t)b->event_rec & 15) || ((uintptr_t)b->event_heap & 7))
        return fail(RLR_E_INVALID, "rlr_bind_buffers: event_rec must be 16-byte and event_heap 8-byte aligned");
    if (((uintptr_t)b->ring & 15) || ((uintptr_t)b->q_table & 15) || ((uintptr_t)b->theta & 15) || ((uintptr_t)b->trace & 15))
        return fail(RLR_E_INVALID, "rlr_bind_buffers: ring and tables must be 16-byte aligned");
    h->bufs = *b;
    h->bound = true;
    return RLR_OK;
}

int rlr_launch_info(const rlr_handle_t *h, int32_t *threads, int32_t *replicas_per_block, int32_t *smem_bytes,
                    int32_t *tables_in_smem)
{
    if (!h) return fail(RLR_E_INVALID, "rlr_launch_info: null handle");
    if (threads) *threads = h->threads;
    if (replicas_per_block) *replicas_per_block = h->G;
    if (smem_bytes) *smem_bytes = h->smem_bytes;
    if (tables_in_smem) *tables_in_smem = h->smem_tables ? 1 : 0;
    return RLR_OK;
}

int rlr_reset(rlr_handle_t *h, void *stream)
{
    if (!h || !h->bound) return fail(RLR_E_STATE, "rlr_reset: buffers not bound");
    RLR_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t RN = (size_t)h->R * h->N;
    RLR_CUDA(cudaMemsetAsync(h->bufs.q_head, 0, RN * 4, st));
    RLR_CUDA(cudaMemsetAsync(h->bufs.q_tail, 0, RN * 4, st));
    RLR_CUDA(cudaMemsetAsync(h->bufs.counters, 0, (size_t)h->R * RLR_NUM_COUNTERS * 8, st));
    RLR_CUDA(cudaMemsetAsync(h->bufs.status, 0, (size_t)h->R * 2 * 4, st));
    if (h->policy == RLR_POLICY_MAHYBRIDQ) RLR_CUDA(cudaMemsetAsync(h->bufs.trace, 0, (size_t)h->R * h->tsize * 8, st));
    if (h->ecap > 0) {                                   // event_queue = [], Node.sent = 0 (env.py:269, 100-101)
        RLR_CUDA(cudaMemsetAsync(h->bufs.link_sent, 0, (size_t)h->R * h->sdeg * 4, st));
        RLR_CUDA(cudaMemsetAsync(h->bufs.event_count, 0, (size_t)h->R * 4, st));
        RLR_CUDA(cudaMemsetAsync(h->bufs.q_dead, 0, RN * 4, st));
    }
    if (h->page_log) {                                   // every queue gets its first page and a spare, the rest is free
        const long long n = h->R * h->n_pages;
        rlr_paged_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->bufs.page_free, h->bufs.page_top, h->bufs.q_pages,
                                                                            h->R, h->N, h->n_pages);
        RLR_CUDA(cudaGetLastError());
    }
    if (h->policy == RLR_POLICY_CQ || h->policy == RLR_POLICY_CDRQ) {              // CQ.clean (qroute.py:66-71)
        const long long total = h->R * h->tsize;
        int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        rlr_init_tables_kernel<double><<<blocks, 256, 0, st>>>(nullptr, h->bufs.theta, nullptr, h->bufs.topo_row_ptr, h->bufs.topo_col,
                                                       h->N, h->sdeg, h->R, 0.0, 1);
        RLR_CUDA(cudaGetLastError());
    }
    return RLR_OK;
}

static int globalroute_init(rlr_handle_t *h, int multiway, void *stream)
{
    const size_t sm = (size_t)(h->N * h->N + h->N + h->N + 1) * 4 + (size_t)h->sdeg * 2 + (size_t)h->N * h->N * 2;
    RLR_CUDA(cudaFuncSetAttribute((const void *)rlr_globalroute_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    rlr_globalroute_init_kernel<<<(unsigned)h->R, 128, sm, (cudaStream_t)stream>>>(h->bufs.gr_choice_mask, h->bufs.gr_distance,
                                                                                    h->bufs.gr_qs_off, h->bufs.topo_row_ptr,
                                                                                    h->bufs.topo_col, h->N, h->sdeg, multiway);
    RLR_CUDA(cudaGetLastError());
    return RLR_OK;
}

int rlr_init_tables(rlr_handle_t *h, double initP, void *stream)
{
    if (!h || !h->bound) return fail(RLR_E_STATE, "rlr_init_tables: buffers not bound");
    if (h->policy == RLR_POLICY_SHORTEST) return RLR_OK;
    RLR_CUDA(cudaSetDevice(h->device));
    if (h->policy == RLR_POLICY_GLOBALROUTE) return globalroute_init(h, 0, stream);       // GlobalRoute.__init__, shortest.py:85-91
    const long long total = h->R * h->tsize;
    int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    const int conf = (h->policy == RLR_POLICY_CQ || h->policy == RLR_POLICY_CDRQ) ? 1 : 0;
    if (h->esize == 4)
        rlr_init_tables_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float *)h->bufs.q_table, h->bufs.theta, h->bufs.trace,
                                                                                 h->bufs.topo_row_ptr, h->bufs.topo_col, h->N, h->sdeg,
                                                                                 h->R, initP, conf);
    else
        rlr_init_tables_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(h->bufs.q_table, h->bufs.theta, h->bufs.trace,
                                                                                  h->bufs.topo_row_ptr, h->bufs.topo_col, h->N, h->sdeg,
                                                                                  h->R, initP, conf);
    RLR_CUDA(cudaGetLastError());
    return RLR_OK;
}

int rlr_build_shortest_tables(rlr_handle_t *h, int32_t multiway, void *stream)
{
    if (!h || !h->bound) return fail(RLR_E_STATE, "rlr_build_shortest_tables: buffers not bound");
    if (h->policy == RLR_POLICY_GLOBALROUTE) {           // per-replica tables with weights 1 + queue_size (shortest.py:84-104)
        RLR_CUDA(cudaSetDevice(h->device));
        return globalroute_init(h, multiway ? 1 : 0, stream);
    }
    if (!h->bufs.next_hop || !h->bufs.distance || !h->bufs.choice)
        return fail(RLR_E_INVALID, "rlr_build_shortest_tables: next_hop, distance and choice are required");
    RLR_CUDA(cudaSetDevice(h->device));
    rlr_apsp_kernel<<<(h->N + 63) / 64, 64, 0, (cudaStream_t)stream>>>(h->bufs.distance, h->bufs.choice, h->bufs.next_hop,
                                                                        h->bufs.choice_mask, h->bufs.topo_row_ptr,
                                                                        h->bufs.topo_col, h->N, multiway ? 1 : 0);
    RLR_CUDA(cudaGetLastError());
    return RLR_OK;
}

int rlr_build_inject_schedule(rlr_handle_t *h, int64_t n_steps, const rlr_run_params_t *rp, void *stream)
{
    if (!h || !h->bound) return fail(RLR_E_STATE, "rlr_build_inject_schedule: buffers not bound");
    if (!rp || !rp->inj_events || rp->inj_cap < 2 || !rp->enlam || !rp->inj_thr)
        return fail(RLR_E_INVALID, "rlr_build_inject_schedule: inj_events, inj_cap, enlam and inj_thr are required");
    if (n_steps < 1 || n_steps >= (1 << 24) - 1) return fail(RLR_E_INVALID, "rlr_build_inject_schedule: n_steps out of range");
    RLR_CUDA(cudaSetDevice(h->device));
    const long long nthr = h->R * h->N;
    rlr_inject_schedule_kernel<<<(unsigned)((nthr + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        rp->inj_events, rp->inj_cap, h->R, h->N, (int)n_steps, rp->clock0, rp->freq > 0 ? rp->freq : 1, rp->slot > 0 ? rp->slot : 1, rp->enlam,
        (const unsigned long long *)rp->inj_thr, rp->seed, h->replica_base, h->bufs.status);
    RLR_CUDA(cudaGetLastError());
    return RLR_OK;
}

int rlr_run_steps(rlr_handle_t *h, int64_t n_steps, const rlr_run_params_t *rp, void *stream)
{
    if (!h || !h->bound) return fail(RLR_E_STATE, "rlr_run_steps: buffers not bound");
    if (!rp) return fail(RLR_E_INVALID, "rlr_run_steps: null params");
    if (n_steps < 0 || n_steps > (1 << 30)) return fail(RLR_E_INVALID, "rlr_run_steps: n_steps out of range");
    if (n_steps == 0) return RLR_OK;
    if (!rp->inj_events || rp->inj_cap < 2) return fail(RLR_E_INVALID, "rlr_run_steps: inj_events / inj_cap are required");
    if (rp->inject_mode == RLR_INJECT_PHILOX) {
        if (!rp->enlam || !rp->inj_thr) return fail(RLR_E_INVALID, "rlr_run_steps: Philox mode needs enlam and inj_thr");
    } else if (rp->inject_mode != RLR_INJECT_TRACE) {
        return fail(RLR_E_INVALID, "rlr_run_steps: unknown inject_mode");
    }
    if (n_steps >= (1 << 24) - 1) return fail(RLR_E_INVALID, "rlr_run_steps: at most 2^24 - 2 steps per call");
    if (rp->phase < 0 || rp->phase > 3) return fail(RLR_E_INVALID, "rlr_run_steps: phase must be 0 (= 3), 1, 2 or 3");
    if ((rp->phase == 1 || rp->phase == 2) && (n_steps != 1 || !rp->rewards))
        return fail(RLR_E_INVALID, "rlr_run_steps: the split phases (Network.step / agent.learn) run one step and need the rewards buffer");
    if ((rp->route_time_out || rp->hop_out || rp->droprate_out) && !rp->metrics)
        return fail(RLR_E_INVALID, "rlr_run_steps: route_time_out / hop_out / droprate_out need the metrics record buffer");
    if (rp->droprate_out && !rp->metrics2) return fail(RLR_E_INVALID, "rlr_run_steps: droprate_out needs the metrics2 buffer");
    const int slot = rp->slot > 0 ? rp->slot : 1, freq = rp->freq > 0 ? rp->freq : 1;
    const int transtime = (rp->slot == 0 && rp->transtime == 0) ? 1 : rp->transtime;     // an all-zero block means the defaults
    const int bandwidth = (rp->slot == 0 && rp->bandwidth == 0) ? 1 : rp->bandwidth;
    if (rp->slot < 0 || rp->freq < 0 || transtime < 0) return fail(RLR_E_INVALID, "rlr_run_steps: slot, freq and transtime must be >= 0");
    if ((long long)rp->clock0 + n_steps * slot + transtime > 0x7FFFFFFFll) return fail(RLR_E_INVALID, "rlr_run_steps: clock overflows int32");
    if (rp->inject_mode == RLR_INJECT_PHILOX && n_steps % freq != 0)
        return fail(RLR_E_INVALID, "rlr_run_steps: n_steps must be a multiple of freq (whole iterations of Network.train's loop)");
    const bool need_general = transtime > slot || bandwidth < 1;
    if (need_general && !rp->general)
        return fail(RLR_E_INVALID, "rlr_run_steps: transtime > slot or bandwidth < 1 needs `general` (events outlive the step)");
    if (rp->general) {
        if (h->ecap < h->N * (transtime + 1))
            return fail(RLR_E_INVALID, "rlr_run_steps: `general` needs a handle created with event_capacity >= N * (transtime + 1)");
        if (h->max_deg > 32) return fail(RLR_E_UNSUPPORTED, "rlr_run_steps: the general timing model supports degrees up to 32");
        if (n_steps * slot + transtime > 65535)
            return fail(RLR_E_INVALID, "rlr_run_steps: with `general`, n_steps * slot + transtime must stay below 65536 per call");
    }
    RLR_CUDA(cudaSetDevice(h->device));
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.N = h->N; p.sdeg = h->sdeg; p.G = h->G; p.R = (int)h->R; p.max_deg = h->max_deg;
    p.cap = h->cap; p.K = (int)n_steps; p.clock0 = rp->clock0; p.ntab = h->ntab;
    p.tsize = h->tsize;
    p.row_ptr = h->bufs.topo_row_ptr; p.col = h->bufs.topo_col; p.heap_pos = h->bufs.heap_pos; p.next_hop = h->bufs.next_hop;
    p.choice_mask = h->bufs.choice_mask; p.shortest_random = rp->shortest_random;
    if (p.shortest_random && h->policy != RLR_POLICY_GLOBALROUTE && (h->policy != RLR_POLICY_SHORTEST || !p.choice_mask))
        return fail(RLR_E_INVALID, "rlr_run_steps: shortest_random needs Shortest with the choice_mask buffer, or GlobalRoute");
    p.gr_choice_mask = h->bufs.gr_choice_mask; p.gr_multiway = rp->multiway; p.gr_distance = h->bufs.gr_distance; p.gr_qs_off = h->bufs.gr_qs_off;
    p.Q = h->bufs.q_table; p.Theta = h->bufs.theta; p.Trace = h->bufs.trace;
    p.ring = (uint4 *)h->bufs.ring; p.qhead = h->bufs.q_head; p.qtail = h->bufs.q_tail;
    p.counters = (long long *)h->bufs.counters; p.status = h->bufs.status;
    p.inj_cap = rp->inj_cap; p.add_entropy = rp->add_entropy; p.is_drop = rp->is_drop;
    p.phase = rp->phase == 0 ? 3 : rp->phase; p.rewards = (uint4 *)rp->rewards;
    p.inj_events = rp->inj_events; p.seed = rp->seed; p.replica_base = h->replica_base;
    p.lr_q = rp->lr_q; p.lr_p = rp->lr_p; p.lr_e = rp->lr_e; p.discount = rp->discount; p.discount_trace = rp->discount_trace; p.conf_decay = rp->conf_decay;
    p.metrics = (uint4 *)rp->metrics; p.metrics2 = (uint2 *)rp->metrics2; p.sendlog = rp->sendlog;
    p.slot = slot; p.freq = freq; p.transtime = transtime; p.bandwidth = bandwidth; p.general = rp->general != 0; p.ecap = h->ecap;
    p.link_sent = h->bufs.link_sent; p.event_heap = (unsigned long long *)h->bufs.event_heap; p.event_count = h->bufs.event_count;
    p.event_rec = (uint4 *)h->bufs.event_rec; p.q_dead = h->bufs.q_dead;
    p.sample = rp->sample; p.sample_idx = rp->sample_idx; p.sample_clock = rp->sample_clock; p.sample_size = rp->sample_size;
    if (p.sample && (!p.sample_idx || !p.sample_clock || p.sample_size < 1))
        return fail(RLR_E_INVALID, "rlr_run_steps: sample needs sample_idx, sample_clock and sample_size >= 1");
    if (rp->inject_mode == RLR_INJECT_PHILOX) {
        const int rc = rlr_build_inject_schedule(h, n_steps, rp, stream);
        if (rc != RLR_OK) return rc;
    }
    const long long blocks = (h->R + h->G - 1) / h->G;
    if (blocks > 0x7FFFFFFFll) return fail(RLR_E_INVALID, "rlr_run_steps: too many blocks");
    // the fused instantiation reads the schedule in 8-byte chunks (cp.async): it needs an even, 8-byte aligned list layout
    const bool plain = p.phase == 3 && !p.sample && !p.sendlog && !p.is_drop && !p.metrics2 && !p.general && slot == 1 &&
                       freq == 1 && transtime == 1 && p.inj_cap % 2 == 0 && ((uintptr_t)p.inj_events & 7) == 0;
    if (h->page_log && !plain)
        return fail(RLR_E_UNSUPPORTED, "rlr_run_steps: paged queues are served by the fused Network.train path only (no split phases, "
                                       "send log, is_drop, sampling or non-default timing)");
    p.pool = (uint4 *)h->bufs.pool; p.page_next = h->bufs.page_next; p.page_free = h->bufs.page_free; p.page_top = h->bufs.page_top;
    p.q_pages = h->bufs.q_pages; p.page_log = h->page_log; p.n_pages = h->n_pages;
    const StepKernel kern = plain ? h->kernel : h->kernel_split;
    const int smem = plain ? h->smem_bytes : h->smem_bytes_split;
    // the attribute belongs to the function, not to the handle: another handle of the same instantiation may have set less
    RLR_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<(unsigned)blocks, h->threads, smem, (cudaStream_t)stream>>>(p);
    RLR_CUDA(cudaGetLastError());
    if (rp->route_time_out || rp->hop_out || rp->droprate_out) {
        const long long n = n_steps * h->R;
        const int rb = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
        rlr_ratio_kernel<<<rb, 256, 0, (cudaStream_t)stream>>>((const uint4 *)rp->metrics, (const uint2 *)rp->metrics2, n,
                                                               rp->route_time_out, rp->hop_out, rp->droprate_out);
        RLR_CUDA(cudaGetLastError());
    }
    return RLR_OK;
}

int rlr_reduce_stats(rlr_handle_t *h, int64_t *out, void *stream)
{
    if (!h || !h->bound || !out) return fail(RLR_E_STATE, "rlr_reduce_stats: buffers not bound or null output");
    RLR_CUDA(cudaSetDevice(h->device));
    RLR_CUDA(cudaMemsetAsync(out, 0, RLR_NUM_COUNTERS * 8, (cudaStream_t)stream));
    const int blocks = (int)((h->R + 255) / 256 < 148 * 4 ? (h->R + 255) / 256 : 148 * 4);
    rlr_reduce_stats_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const long long *)h->bufs.counters, h->R, (long long *)out);
    RLR_CUDA(cudaGetLastError());
    return RLR_OK;
}

int rlr_reduce_levels(rlr_handle_t *h, const void *metrics, int64_t n_steps, const int32_t *level_idx, int32_t n_levels,
                      double *route_time_out, double *hop_out, int64_t *sums_out, void *stream)
{
    if (!h || !h->bound) return fail(RLR_E_STATE, "rlr_reduce_levels: buffers not bound");
    if (!metrics || !level_idx || n_steps < 1 || n_levels < 1 || n_levels > 2048)
        return fail(RLR_E_INVALID, "rlr_reduce_levels: metrics, level_idx, n_steps >= 1 and 1 <= n_levels <= 2048 are required");
    if (!route_time_out && !hop_out && !sums_out) return fail(RLR_E_INVALID, "rlr_reduce_levels: no output requested");
    if (n_steps > 0x7FFFFFFFll) return fail(RLR_E_INVALID, "rlr_reduce_levels: too many steps");
    RLR_CUDA(cudaSetDevice(h->device));
    rlr_reduce_levels_kernel<<<(unsigned)n_steps, 256, (size_t)n_levels * 24, (cudaStream_t)stream>>>(
        (const uint4 *)metrics, h->R, level_idx, n_levels, route_time_out, hop_out, (long long *)sums_out);
    RLR_CUDA(cudaGetLastError());
    return RLR_OK;
}

int rlr_table_stats(rlr_handle_t *h, int32_t table_id, const double *prev, double *per_replica, const int32_t *level_idx,
                    int32_t n_levels, double *level_out, void *stream)
{
    if (!h || !h->bound) return fail(RLR_E_STATE, "rlr_table_stats: buffers not bound");
    const double *tab = table_id == 0 ? h->bufs.q_table : table_id == 1 ? h->bufs.theta : table_id == 2 ? h->bufs.trace : nullptr;
    if (!tab || h->ntab <= table_id) return fail(RLR_E_INVALID, "rlr_table_stats: the policy has no such table");
    if (!per_replica) return fail(RLR_E_INVALID, "rlr_table_stats: per_replica ([R][2] doubles) is required");
    if (level_out && n_levels < 1) return fail(RLR_E_INVALID, "rlr_table_stats: level_out needs n_levels >= 1");
    RLR_CUDA(cudaSetDevice(h->device));
    const long long warps = h->R, blocks = (warps * 32 + 255) / 256;
    if (h->esize == 4 && table_id == 0)
        rlr_table_stats_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float *)tab, (const float *)prev, h->R, h->tsize, per_replica);
    else
        rlr_table_stats_kernel<double><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(tab, prev, h->R, h->tsize, per_replica);
    RLR_CUDA(cudaGetLastError());
    if (level_out) {
        rlr_level_sums_kernel<<<(unsigned)n_levels, 32, 0, (cudaStream_t)stream>>>(per_replica, h->R, level_idx, n_levels, level_out);
        RLR_CUDA(cudaGetLastError());
    }
    return RLR_OK;
}

}  // extern "C"
