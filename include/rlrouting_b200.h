/*
 * rlrouting_b200.h — C ABI of the B200-native batched packet-routing simulator.
 *
 * The reference (xuxife/rlrouting) is pure Python and has NO FFI; its plugin boundary is the two
 * Python interfaces `Network` (env.py:213-450) and `Policy` (base_policy.py:5-66).  This header
 * is the boundary a maintainer binds UNDER those interfaces (ctypes stub in INTEGRATION.md): every
 * entry point below names the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative RLR_E_* code, and
 *     rlr_last_error_string() describes the last failure on the calling thread;
 *   - the caller (PyTorch in the shipped host layer) allocates and owns EVERY device buffer; the
 *     library borrows the pointers in rlr_buffers_t for the lifetime of the handle and allocates
 *     nothing on the device;
 *   - kernels are launched on the stream passed in (a cudaStream_t cast to void*); nothing here
 *     synchronises the device;
 *   - one host thread per handle; no callbacks into the host language;
 *   - asynchronous device conditions (queue ring overflow; NaN softmax probabilities, where the
 *     reference raises ValueError out of np.random.choice, hybrid.py:20-23) are reported through
 *     the sticky per-replica status words in rlr_buffers_t.status, never by dropping a packet.
 */
#ifndef RLROUTING_B200_H
#define RLROUTING_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLR_ABI_VERSION 4

enum { RLR_OK = 0, RLR_E_INVALID = -1, RLR_E_UNSUPPORTED = -2, RLR_E_CUDA = -3, RLR_E_STATE = -4 };

/* Policy ids: Shortest (shortest.py:6-58), Qroute (qroute.py:6-53), HybridQ (hybrid.py:41-62),
 * MaHybridQ (multi_agent.py:6-50). */
enum {
    RLR_POLICY_SHORTEST = 0, RLR_POLICY_QROUTE = 1, RLR_POLICY_HYBRIDQ = 2, RLR_POLICY_MAHYBRIDQ = 3,
    RLR_POLICY_CQ = 4,      /* qroute.py:56-101  confidence-based Q-routing                              */
    RLR_POLICY_CDRQ = 5,    /* qroute.py:104-122 + Node 'dual' mode env.py:92-94,154-160                  */
    RLR_POLICY_GLOBALROUTE = 6 /* shortest.py:84-104 all-pairs distances recomputed every step, weights 1 + queue size */
};

/* Injection: RLR_INJECT_TRACE replays host-generated draws of the reference's NumPy legacy stream
 * (env.py:307-310); RLR_INJECT_PHILOX draws on the device (Philox4x32-10 keyed per replica, clock,
 * node; Poisson splitting, SURVEY §8c law ii). */
enum { RLR_INJECT_TRACE = 0, RLR_INJECT_PHILOX = 1 };

/* Sticky per-replica status codes (status[2*r] = code, status[2*r+1] = clock of the first event). */
enum { RLR_STATUS_OK = 0, RLR_STATUS_NAN_PROB = 1, RLR_STATUS_QUEUE_OVERFLOW = 2, RLR_STATUS_INJECT_OVERFLOW = 3 };

/* Per-replica counters, int64 each (Network attributes, env.py:267-276 / 321-331). */
enum {
    RLR_CTR_ALL_PACKETS = 0, RLR_CTR_END_PACKETS = 1, RLR_CTR_DROP_PACKETS = 2, RLR_CTR_HOPS = 3,
    RLR_CTR_ROUTE_TIME = 4, RLR_CTR_SENDS = 5, RLR_CTR_RESERVED0 = 6, RLR_CTR_RESERVED1 = 7,
    RLR_NUM_COUNTERS = 8
};

typedef struct rlr_handle rlr_handle_t;

/* Graph in the order Network.read_network builds it (env.py:282-296); HOST pointers, copied. */
typedef struct {
    int32_t n_nodes;          /* N                                                          */
    int32_t sum_deg;          /* sum of degrees                                             */
    int32_t max_deg;
    const int32_t *row_ptr;   /* [N+1]  links[x] = col[row_ptr[x] .. row_ptr[x+1])          */
    const int32_t *col;       /* [sum_deg]                                                  */
} rlr_topology_t;

typedef struct {
    int64_t replicas;         /* environments simulated by this handle (on one device)      */
    int64_t replica_base;     /* global index of local replica 0: Philox stream id = base+r */
    int32_t policy;           /* RLR_POLICY_*                                               */
    int32_t queue_capacity;   /* ring slots per (replica, node); power of two               */
    int32_t replicas_per_block; /* 0 = choose                                               */
    int32_t threads_per_block;  /* 0 = choose; >= replicas_per_block * N, multiple of 32    */
    int32_t tables_in_smem;   /* -1 = choose, 0 = tables stay in HBM, 1 = staged in smem    */
    int32_t device;           /* CUDA device ordinal the buffers live on                    */
    int32_t event_capacity;   /* events that can be in flight per replica when they outlive a step (transtime > slot,
                                 env.py:143-145): N * (transtime + 1); 0 = the default timing model only */
    int32_t table_dtype;          /* 0: Q table stored as float64 (the reference's precision, the default); 1: float32 storage
                                   * (Qroute only): rows are widened to fp64 when read, updates computed in fp64 and rounded
                                   * on store — `prev` of rlr_table_stats and the q_table buffer are then float arrays */
    int32_t page_records;         /* 0: one ring of queue_capacity slots per (replica, node).  A power of two >= 4: PAGED queues —
                                   * the reference's unbounded lists (env.py:100,130) as linked pages of this many records taken
                                   * from a per-replica pool, so memory follows the packets alive, not the worst node of the run;
                                   * served by the fused Network.train path (rlr_run_steps with default options) */
    int32_t pages_per_replica;    /* pool size (<= 65535; >= 2 * N + 1); an empty pool sets RLR_STATUS_QUEUE_OVERFLOW */
} rlr_config_t;

/* Raw DEVICE pointers borrowed from the caller.  R = replicas, N = n_nodes,
 * TS = N * sum_deg (table[x][d][k] at N*row_ptr[x] + d*deg[x] + k, the reference's per-node
 * (N, deg_x) arrays of qroute.py:13-15 packed back to back). */
typedef struct {
    double *q_table;          /* [R][TS]  Qroute.Qtable            (NULL for Shortest)       */
    double *theta;            /* [R][TS]  PolicyGradient.Theta (HybridQ, MaHybridQ) or CQ.confidence (CQ, CDRQ) */
    double *trace;            /* [R][TS]  MaHybridQ.Trace                                     */
    void *ring;               /* [R][N][cap] 16-byte packet records (Node.queue, env.py:100) */
    uint32_t *q_head;         /* [R][N] monotonically increasing pop counter                 */
    uint32_t *q_tail;         /* [R][N] monotonically increasing push counter                */
    int64_t *counters;        /* [R][RLR_NUM_COUNTERS]                                       */
    int32_t *status;          /* [R][2]                                                      */
    uint8_t *next_hop;        /* [N][N] Shortest: action index chosen at x for dest d        */
    double *distance;         /* [N][N] Shortest.distance (float64, inf = unreachable)       */
    uint8_t *choice;          /* [TS]   Shortest.choice masks, packed like the tables        */
    int32_t *topo_row_ptr;    /* [N+1]   device copy of the CSR (uploaded by the caller)     */
    int32_t *topo_col;        /* [sum_deg]                                                   */
    uint8_t *heap_pos;        /* [(N+1)*N] heap_pos[k*N + rank]: pop position among k ties   */
    /* GlobalRoute only (per replica, persist across calls) */
    uint16_t *gr_choice_mask; /* [R][N*N] GlobalRoute.choice[x][d] as bits (bit j = links[x][j]); one bit unless multiway */
    int32_t *gr_distance;     /* [R][N*N] GlobalRoute.distance as integers, 0x3FFFFFFF = inf   */
    uint16_t *choice_mask;    /* [N][N] Shortest: bit j = choice[x][d][j] (all optimal neighbours with multiway=True);
                                 needed when shortest_random is set */
    int32_t *gr_qs_off;       /* [R][N] queue_size[x] - len(queue[x]): 0 until Network.reset, which clears the queues
                                 but not the agent's queue_size (env.py:267-280, shortest.py:89) */
    /* general timing model only (event_capacity > 0; all persist across calls, zeroed by rlr_reset) */
    int32_t *link_sent;       /* [R][sum_deg] Node.sent[y] per directed link (env.py:100,115-116,142,352)       */
    uint64_t *event_heap;     /* [R][event_capacity] Network.event_queue as the heapq array (env.py:144,349-353):
                                 arrive_time << 32 | link << 16 | record slot, ordered by arrive_time only      */
    int32_t *event_count;     /* [R] len(event_queue)                                                           */
    void *event_rec;          /* [R][event_capacity] 16-byte packet records of the events in flight, at slot
                                 (arrive_time mod (transtime + 1)) * N + sender                                 */
    int32_t *q_dead;          /* [R][N] ring entries between head and tail already sent out of order (queue.pop(i),
                                 env.py:177): len(queue) = tail - head - q_dead                                 */
    /* paged queues (page_records > 0; `ring` may then be NULL).  Record i of a queue (q_head <= i < q_tail) is at offset
     * i mod page_records of the (i / page_records - q_head / page_records)-th page of the list that starts at q_pages[..][0]
     * and follows page_next.  rlr_reset initialises all of them. */
    void *pool;               /* [R][pages_per_replica * page_records] 16-byte packet records                           */
    uint16_t *page_next;      /* [R][pages_per_replica] the page that follows in its queue                              */
    uint16_t *page_free;      /* [R][pages_per_replica] stack of free pages                                             */
    int32_t *page_top;        /* [R] entries on that stack                                                              */
    uint16_t *q_pages;        /* [R][N][4] {page under q_head, page under q_tail, the node's spare page, 0}            */
} rlr_buffers_t;

typedef struct {
    int32_t inject_mode;      /* RLR_INJECT_*                                                */
    int32_t add_entropy;      /* HybridQ(add_entropy=...)  hybrid.py:44                      */
    /* Injection schedule, device [R][N][inj_cap] uint32: for each (replica, node) the ascending list of its
     * injections over this call, one entry (local_step << 8 | dest) per packet, terminated by 0xFFFFFFFF.
     * RLR_INJECT_TRACE: filled by the caller from the host replay of env.py:307-310.
     * RLR_INJECT_PHILOX: scratch filled by the library (rlr_inject_schedule_kernel) from enlam / inj_thr / seed;
     *   a node with more than inj_cap - 1 injections sets RLR_STATUS_INJECT_OVERFLOW. */
    uint32_t *inj_events;
    int32_t inj_cap;
    int32_t is_drop;          /* Network(is_drop=True): a packet arriving with hops >= N is dropped, env.py:355-360 */
    /* Philox mode */
    const double *enlam;      /* device [R]: exp(-lambda_r / N)                              */
    const uint64_t *inj_thr;  /* device [R]: smallest w in [0, 2^32] with (w + 0.5) * 2^-32 > enlam[r] */
    uint64_t seed;
    /* learning rates lr['q'], lr['p'], lr['e'] (qroute.py:46, hybrid.py:55, multi_agent.py:17) */
    double lr_q, lr_p, lr_e;
    double discount;          /* qroute.py:9 / hybrid.py:44                                  */
    double discount_trace;    /* multi_agent.py:10                                           */
    double conf_decay;        /* CQ(decay=...) qroute.py:59                                  */
    int32_t clock0;           /* Network.clock before the first step                         */
    int32_t shortest_random;  /* Shortest / GlobalRoute(random=True): next hop drawn uniformly from the choice mask (shortest.py:40-41) */
    int32_t multiway;         /* GlobalRoute(multiway=True): the per-step recomputation keeps every optimal neighbour */
    int32_t phase;            /* 0 or 3: fused inject+step+learn (Network.train); 1: Network.step only (env.py:333-365,
                                 rewards saved, nothing learned); 2: agent.learn of the rewards saved by phase 1.
                                 Phases 1 and 2 need n_steps == 1 and `rewards`. */
    void *rewards;            /* device [R][N] 64-byte Reward records (env.py:52-70): packet record (16 B);
                                 {sending | k << 1 | y << 9 | dest << 17, q_y, q_x ('dual' mode), 0};
                                 {max_Q_y (CQ/CDRQ: max_Q_f), max_Q_x_d (CQ/CDRQ: C_f)}; {C_b, max_Q_b} (CDRQ) */
    /* outputs (device pointers, may be NULL) */
    void *metrics;            /* [n_steps][R] 16-byte records {int64 route_time, uint32 end_packets, uint32 hops}:
                                 cumulative counters after each step (env.py:409-413); step-major so a chunk
                                 of steps is one contiguous block.  Required by route_time_out / hop_out. */
    double *route_time_out;   /* [n_steps][R] Network.ave_route_time after each step (env.py:409, 444-446); step-major
                                 so a chunk of steps is one contiguous block for the device->host copy */
    double *hop_out;          /* [n_steps][R] Network.ave_hops after each step (env.py:412-413, 440-442) */
    void *metrics2;           /* [n_steps][R] {uint32 all_packets, uint32 drop_packets} after each step; for droprate_out */
    double *droprate_out;     /* [n_steps][R] Network.drop_rate after each step (env.py:410-411, 448-450) */
    /* Network.sample_route_time (env.py:416-438): when `sample` is set, every delivery files its routing time at
     * sample[r][min(sample_idx[r], sample_size-1)] in the reference's delivery order, and a replica stops stepping
     * (sample_clock[r] = its clock) once sample_idx[r] >= sample_size */
    int32_t *sample;          /* [R][sample_size] */
    int32_t *sample_idx;      /* [R], persists across calls */
    int32_t *sample_clock;    /* [R] */
    int32_t sample_size;
    int32_t reserved2;
    uint32_t *sendlog;        /* [R][n_steps][N]: y | dest << 16 for a send by node x, else 0xFFFFFFFF */
    /* timing model, integers; 0 means 1.  n_steps counts Network.step calls: one injection
     * (Poisson(lambda * slot), the caller folds slot into enlam / inj_thr) every `freq` steps, each step advancing the
     * clock by `slot`; a packet sent at clock c arrives at c + transtime; a link carries at most `bandwidth` packets.
     * transtime > slot or bandwidth < 1 needs a handle created with event_capacity > 0 and `general` set. */
    int32_t slot;             /* Network.train(slot=...) / Network.step(duration), env.py:333,372,403 */
    int32_t freq;             /* Network.train(freq=...), env.py:373,402 */
    int32_t transtime;        /* Network(transtime=...), env.py:143 */
    int32_t bandwidth;        /* Network(bandwidth=...), env.py:115-116 */
    int32_t general;          /* 1: events persist in event_heap, links can be busy, the send loop scans the queue
                                 (env.py:171-190); 0: every event arrives inside its step (transtime <= slot) */
    int32_t reserved3;
} rlr_run_params_t;

/* ---- introspection ---- */
int rlr_abi_version(void);
const char *rlr_last_error_string(void);
/* sizeof(rlr_topology_t), sizeof(rlr_config_t), sizeof(rlr_buffers_t), sizeof(rlr_run_params_t): lets a
 * foreign-language binding verify its struct layout against the library it loaded. */
int rlr_struct_sizes(int32_t out[4]);

/* ---- lifecycle: replaces Network.__init__ (env.py:237-250) for R replicas ---- */
int rlr_create(const rlr_topology_t *topo, const rlr_config_t *cfg, rlr_handle_t **out);
int rlr_destroy(rlr_handle_t *h);
int rlr_bind_buffers(rlr_handle_t *h, const rlr_buffers_t *bufs);

/* Chosen launch geometry, for reporting: threads per block, replicas per block, dynamic shared
 * memory bytes, and whether the tables are staged in shared memory. */
int rlr_launch_info(const rlr_handle_t *h, int32_t *threads, int32_t *replicas_per_block,
                    int32_t *smem_bytes, int32_t *tables_in_smem);

/* ---- Network.reset (env.py:267-280) + Node.reset (env.py:99-101) + agent.clean()
 *      (multi_agent.py:48-50): zero queues, counters and status; MaHybridQ traces are zeroed;
 *      Q / Theta are kept. ---- */
int rlr_reset(rlr_handle_t *h, void *stream);

/* ---- Qroute.__init__ structure overrides (qroute.py:16-20) applied to q_table already holding
 *      the N(initQ,1) draws; Theta = initP (hybrid.py:13-14); Trace = 0 (multi_agent.py:14-15) ---- */
int rlr_init_tables(rlr_handle_t *h, double initP, void *stream);

/* ---- Shortest.__init__ + _calc_distance (shortest.py:20-58): fills distance, choice, next_hop (first True of the
 *      mask) and, if bound, choice_mask; multiway != 0 keeps every optimal neighbour (shortest.py:57-58).
 *      For a GlobalRoute handle: GlobalRoute.__init__ (shortest.py:84-91), every replica's gr_choice_mask / gr_distance
 *      with node weights 1 + queue_size ---- */
int rlr_build_shortest_tables(rlr_handle_t *h, int32_t multiway, void *stream);

/* ---- host helper (no device work): replays NumPy's legacy RandomState stream for n_steps calls of
 *      Network.new_packet(lam) (env.py:298-312; np.random.poisson + np.random.randint) on the MT19937 state
 *      exported by RandomState.get_state() (mt_key[624], *mt_pos; both updated).  cnt_out[n_steps] receives the
 *      packet count of each step, pairs_out the packets as source | dest << 16.  Trace mode uses it instead of
 *      calling NumPy once per draw. ---- */
int rlr_trace_replay(uint32_t *mt_key, int32_t *mt_pos, int32_t n_nodes, int64_t n_steps, double lam,
                     int32_t *cnt_out, uint32_t *pairs_out, int64_t pairs_cap, int64_t *n_pairs_out);

/* the same for n_streams independent RandomState streams (mt_keys[n_streams][624], mt_pos[n_streams], lam[n_streams];
 * cnt_out[n_streams][n_steps], pairs_out[n_streams][pairs_cap], n_pairs_out[n_streams]) on nthreads host threads */
int rlr_trace_replay_many(uint32_t *mt_keys, int32_t *mt_pos, int64_t n_streams, int32_t n_nodes, int64_t n_steps,
                          const double *lam, int32_t *cnt_out, uint32_t *pairs_out, int64_t pairs_cap,
                          int64_t *n_pairs_out, int32_t nthreads);

/* ---- host helper (no device work): turns the traces of rlr_trace_replay_many (cnt[n_streams][n_steps], packed pairs
 *      [n_streams][pairs_cap]) into the per-node event lists rlr_run_params_t.inj_events takes, for the iterations
 *      [it0, it1) of Network.train's loop: events[n_streams][n_nodes][cap], entry ((it - it0) * freq) << 8 | dest,
 *      0xFFFFFFFF-terminated.  With events == NULL only *cap_out (longest list + 2) is computed. ---- */
int rlr_trace_schedule(const int32_t *cnt, const uint32_t *pairs, int64_t pairs_cap, int64_t n_streams, int32_t n_nodes,
                       int64_t n_steps, int64_t it0, int64_t it1, int32_t freq, uint32_t *events, int32_t cap,
                       int32_t *cap_out, int32_t nthreads);

/* ---- Network.new_packet (env.py:298-312) for n_steps steps ahead of time, Philox mode: fills params->inj_events
 *      from enlam / inj_thr / seed / clock0.  rlr_run_steps does this itself for RLR_INJECT_PHILOX; calling it
 *      separately (on another stream, then running with RLR_INJECT_TRACE) overlaps it with the previous chunk. ---- */
int rlr_build_inject_schedule(rlr_handle_t *h, int64_t n_steps, const rlr_run_params_t *params, void *stream);

/* ---- the hot path: n_steps calls of Network.step + agent.learn inside Network.train's loop (env.py:400-413),
 *      with new_packet + inject before every freq-th of them (slot = freq = transtime = bandwidth = 1 by default) ---- */
int rlr_run_steps(rlr_handle_t *h, int64_t n_steps, const rlr_run_params_t *params, void *stream);

/* ---- sums the per-replica counters into out[RLR_NUM_COUNTERS] (device pointer, int64) so a
 *      load level's statistics can be all-reduced across GPUs ---- */
int rlr_reduce_stats(rlr_handle_t *h, int64_t *out, void *stream);

/* ---- per-load-level curves: pools the per-(step, replica) records `metrics` ([n_steps][R] 16-byte records written by
 *      rlr_run_steps: cumulative route_time u64, end_packets u32, hops u32) over the replicas of each level
 *      (level_idx[r] in [0, n_levels)) and writes, per step and level, Network.ave_route_time / ave_hops of the pooled
 *      level (env.py:440-446) and/or the three integer sums ([n_steps][n_levels][3] int64).  This is what leaves the
 *      device in a load sweep instead of the [n_steps][R] vectors of env.py:394-414.  All pointers are device pointers. ---- */
int rlr_reduce_levels(rlr_handle_t *h, const void *metrics, int64_t n_steps, const int32_t *level_idx, int32_t n_levels,
                      double *route_time_out, double *hop_out, int64_t *sums_out, void *stream);

/* ---- Q-convergence statistics of a table (0 = Qtable, 1 = Theta / confidence, 2 = Trace): per replica
 *      {sum of entries, sum of |entry - prev entry|} into per_replica[R][2] (prev: an earlier copy of the table, or NULL
 *      for zeros in the second slot) and, if level_out is given, the same pooled per load level into level_out[n_levels][2]
 *      (level_idx == NULL: one level).  Fixed summation order: reproducible fp64 (what qroute.py:40-44 moved). ---- */
int rlr_table_stats(rlr_handle_t *h, int32_t table_id, const double *prev, double *per_replica, const int32_t *level_idx,
                    int32_t n_levels, double *level_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RLROUTING_B200_H */
