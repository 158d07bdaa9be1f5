/*
 * rlr_oracle.c — CPU restatement of the xuxife/rlrouting hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rlrouting_b200/ may import, link or call this
 * file.  Its users are tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs, always as the checker or the timed CPU baseline, never as the
 * product path.
 *
 * Parity pin: the reference has no tests or golden vectors of its own, so this
 * restatement is pinned against outputs of the UNMODIFIED reference run in the build
 * container (tests/golden/gen_golden.py -> tests/golden/ fixtures, tests/test_oracle_golden.py)
 * and against the known-answer table in SURVEY §8(c).
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * One orc_t is one environment replica == one `Network` + one bound agent.
 *
 * Scope (SURVEY §8a + row f2): the reference's full timing model with INTEGER bandwidth, transtime, slot and freq
 * (orc_set_timing; defaults 1).  Under the defaults every event of a step arrives inside the step (env.py:349-353),
 * `sent` returns to 0 and `avaliable_path` is all-True (env.py:171-174), so the head of the queue is the packet sent
 * (env.py:175-180).  With transtime > slot events outlive the step: the literal event heap persists, links stay busy
 * (env.py:115-116) and the send loop scans the queue for the first packet whose chosen link is free (env.py:174-190).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { POL_SHORTEST = 0, POL_QROUTE = 1, POL_HYBRIDQ = 2, POL_MAHYBRIDQ = 3, POL_CQ = 4, POL_CDRQ = 5, POL_GLOBALROUTE = 6 };
enum { INJ_TRACE = 0, INJ_PHILOX = 1 };
enum { ST_OK = 0, ST_NAN_PROB = 1 };

/* env.py:9-27 Packet (+ start_queue, env.py:129).  trans_time is the constant 1. */
typedef struct { int32_t dest, source, hops, birth, start_queue; } pkt_t;

/* Node.queue (env.py:100): an unbounded Python list used as a FIFO. */
typedef struct { pkt_t *buf; int64_t head, len, cap; } fifo_t;

/* env.py:30-49 Event; ordered by arrive_time ONLY (env.py:48-49). */
typedef struct { pkt_t p; int32_t from, to, link; int64_t arrive; } event_t;   /* link = row_ptr[from] + action index */

/* env.py:52-70 Reward + the agent_info dict flattened. */
typedef struct {
    int32_t x, y, d, k, src; int64_t q_y, q_x;
    double max_Q_y, max_Q_x_d;          /* for CQ/CDRQ max_Q_y holds max_Q_f */
    double C_f, C_b, max_Q_b;           /* qroute.py:72-78, 107-115 */
} reward_t;

typedef struct orc {
    int N, sdeg, policy;
    int *row_ptr, *col;       /* links[x] = col[row_ptr[x] .. row_ptr[x+1])   base_policy.py:19-20 */
    int *aidx;                /* action_idx[x][y] (N*N, -1 if absent)          base_policy.py:21-23 */
    int64_t *toff, tsize;     /* table[x] is (N, deg_x) at toff[x]             qroute.py:13-15      */
    double *Q, *Theta, *Trace;
    double *Conf;             /* CQ.confidence (qroute.py:62-63); shares storage with Theta (never both) */
    double decay;             /* CQ(decay=0.9), qroute.py:59 */
    uint8_t *choice;          /* shortest.py:26, same packing as the tables */
    double *distance;         /* shortest.py:24 */
    fifo_t *queue;
    event_t *heap; int heap_len, heap_cap;
    reward_t *rewards; int n_rewards;
    double *scratch;          /* 3*N doubles for MaHybridQ.learn's gathered vectors */
    int64_t clock, all_packets, end_packets, drop_packets, active_packets, hops, route_time, sends;
    double discount, discount_trace, reward_shape;
    int add_entropy, status;
    int is_drop;              /* Network(is_drop=True): env.py:355-360 */
    int bandwidth, transtime; /* Network(bandwidth=1, transtime=1), env.py:237-239 (integers) */
    int slot, freq;           /* Network.train(slot=1, freq=1), env.py:367-373 (integers) */
    int *sent;                /* Node.sent[y] per directed link (env.py:100,115-116,142,352) */
    int64_t probes;           /* diagnostic: agent.choose calls made by the send loops (queue entries looked at) */
    int q_f32;                /* checker of the product's fp32 STORAGE mode (not in the reference): every value written to the
                               * Q table is rounded to float; arithmetic stays fp64, as the device widens rows when it reads them */
    int multiway, random_choice; /* Shortest(multiway=..., random=...), shortest.py:20-23,38-41 */
    int64_t *sample; int64_t sample_size, sample_idx;   /* Network.sample / _sample_idx (env.py:242, 325-328, 424-425) */
    int64_t *queue_size;      /* GlobalRoute.queue_size (shortest.py:89): +1 per receive callback, -1 per send callback */
    int64_t status_clock;
    uint64_t seed, replica;   /* Philox key / stream id */
} orc_t;

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11; the published Random123 constants).  Not in the
 * reference: it replaces the reference's process-wide MT19937 stream (SURVEY §8c "Philox
 * mode") with a counter-based one keyed per (replica, clock, node, purpose).
 * ------------------------------------------------------------------------------------------ */
static inline void philox4x32_10(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4])
{
    uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3];
    uint32_t k0 = k_in[0], k1 = k_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr, key, out);
}

/* Word stream for (replica, clock, node, purpose): block b supplies words 4b..4b+3. */
typedef struct { uint32_t c[4], k[2], w[4]; int have, block; } pstream_t;

static void ps_init(pstream_t *s, uint64_t seed, uint64_t replica, int64_t clock, int node, int purpose)
{
    s->c[0] = (uint32_t)clock;
    s->c[1] = (uint32_t)node | ((uint32_t)purpose << 16);
    s->c[2] = (uint32_t)replica;
    s->c[3] = (uint32_t)(replica >> 32);   /* low 8 bits of c[1]'s top byte carry the block index */
    s->k[0] = (uint32_t)seed; s->k[1] = (uint32_t)(seed >> 32);
    s->have = 0; s->block = 0;
}

static uint32_t ps_next(pstream_t *s)
{
    if (s->have == 0) {
        uint32_t c[4] = { s->c[0], s->c[1] | ((uint32_t)s->block << 24), s->c[2], s->c[3] };
        philox4x32_10(c, s->k, s->w);
        s->have = 4; s->block++;
    }
    return s->w[4 - s->have--];
}

/* ------------------------------------------------------------------------------------------ */
static void fifo_push(fifo_t *q, const pkt_t *p)
{
    if (q->head + q->len == q->cap) {
        if (q->head > 0) {
            memmove(q->buf, q->buf + q->head, (size_t)q->len * sizeof(pkt_t));
            q->head = 0;
        }
        if (q->len * 2 >= q->cap) {
            q->cap = q->cap ? q->cap * 2 : 16;
            q->buf = (pkt_t *)realloc(q->buf, (size_t)q->cap * sizeof(pkt_t));
        }
    }
    q->buf[q->head + q->len++] = *p;
}

orc_t *orc_create(int N, const int *row_ptr, const int *col, int policy)
{
    orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
    o->N = N; o->sdeg = row_ptr[N]; o->policy = policy;
    o->row_ptr = (int *)malloc(sizeof(int) * (N + 1));
    memcpy(o->row_ptr, row_ptr, sizeof(int) * (N + 1));
    o->col = (int *)malloc(sizeof(int) * o->sdeg);
    memcpy(o->col, col, sizeof(int) * o->sdeg);
    o->aidx = (int *)malloc(sizeof(int) * N * N);
    for (int i = 0; i < N * N; ++i) o->aidx[i] = -1;
    o->toff = (int64_t *)malloc(sizeof(int64_t) * (N + 1));
    for (int x = 0; x < N; ++x) {
        o->toff[x] = (int64_t)N * row_ptr[x];
        for (int k = row_ptr[x]; k < row_ptr[x + 1]; ++k) o->aidx[x * N + col[k]] = k - row_ptr[x];
    }
    o->toff[N] = o->tsize = (int64_t)N * o->sdeg;
    o->Q = (double *)calloc((size_t)o->tsize, sizeof(double));
    o->Theta = (double *)calloc((size_t)o->tsize, sizeof(double));
    o->Trace = (double *)calloc((size_t)o->tsize, sizeof(double));
    o->choice = (uint8_t *)calloc((size_t)o->tsize, 1);
    o->distance = (double *)malloc(sizeof(double) * N * N);
    o->queue = (fifo_t *)calloc((size_t)N, sizeof(fifo_t));
    o->heap_cap = N + 8; o->heap = (event_t *)malloc(sizeof(event_t) * o->heap_cap);
    o->rewards = (reward_t *)malloc(sizeof(reward_t) * N);
    o->queue_size = (int64_t *)calloc((size_t)N, sizeof(int64_t));
    o->scratch = (double *)malloc(sizeof(double) * 3 * N);
    o->sent = (int *)calloc((size_t)o->sdeg, sizeof(int));
    o->bandwidth = o->transtime = o->slot = o->freq = 1;
    o->discount = 0.99;        /* qroute.py:9, hybrid.py:44, multi_agent.py:10 */
    o->discount_trace = 0.6;   /* multi_agent.py:10 */
    o->add_entropy = 1;        /* hybrid.py:44 */
    o->Conf = o->Theta;
    o->decay = 0.9;
    if (policy == POL_CQ || policy == POL_CDRQ) o->discount = 0.9;     /* qroute.py:59 */
    return o;
}

void orc_destroy(orc_t *o)
{
    if (!o) return;
    for (int x = 0; x < o->N; ++x) free(o->queue[x].buf);
    free(o->row_ptr); free(o->col); free(o->aidx); free(o->toff); free(o->Q); free(o->Theta);
    free(o->Trace); free(o->choice); free(o->distance); free(o->queue); free(o->heap);
    free(o->rewards); free(o->scratch); free(o->queue_size); free(o->sent); free(o);
}

void orc_set_hparams(orc_t *o, double discount, double discount_trace, int add_entropy)
{
    o->discount = discount; o->discount_trace = discount_trace; o->add_entropy = add_entropy;
}

void orc_set_decay(orc_t *o, double decay) { o->decay = decay; }
void orc_set_is_drop(orc_t *o, int is_drop) { o->is_drop = is_drop; }
void orc_set_q_f32(orc_t *o, int on) { o->q_f32 = on; }
void orc_set_timing(orc_t *o, int bandwidth, int transtime, int slot, int freq)
{
    o->bandwidth = bandwidth; o->transtime = transtime; o->slot = slot; o->freq = freq;
}
int orc_in_flight(const orc_t *o) { return o->heap_len; }
int64_t orc_probes(const orc_t *o) { return o->probes; }
void orc_set_shortest_flags(orc_t *o, int multiway, int random_choice) { o->multiway = multiway; o->random_choice = random_choice; }

void orc_set_stream(orc_t *o, uint64_t seed, uint64_t replica) { o->seed = seed; o->replica = replica; }

int64_t orc_table_size(const orc_t *o) { return o->tsize; }
double *orc_table(orc_t *o, int which) { return which == 0 ? o->Q : which == 1 ? o->Theta : o->Trace; }
uint8_t *orc_choice(orc_t *o) { return o->choice; }
double *orc_distance(orc_t *o) { return o->distance; }
int orc_status(const orc_t *o, int64_t *clock_out) { if (clock_out) *clock_out = o->status_clock; return o->status; }

void orc_counters(const orc_t *o, int64_t out[8])
{
    out[0] = o->clock; out[1] = o->all_packets; out[2] = o->end_packets; out[3] = o->drop_packets;
    out[4] = o->active_packets; out[5] = o->hops; out[6] = o->route_time; out[7] = o->sends;
}

int64_t orc_queue_len(const orc_t *o, int x) { return o->queue[x].len; }

/* Copies queue x (head first) as rows (dest, source, hops, birth, start_queue). */
void orc_queue_get(const orc_t *o, int x, int32_t *out)
{
    const fifo_t *q = &o->queue[x];
    for (int64_t i = 0; i < q->len; ++i) memcpy(out + 5 * i, &q->buf[q->head + i], sizeof(pkt_t));
}

/* Qroute.__init__ structure overrides (qroute.py:16-20), applied to a table already holding
 * the N(initQ,1) draws (qroute.py:13-15, supplied by the caller):
 *   table[x] = 0 ; table[links[x]] = -eye(deg_x)                                            */
void orc_qtable_structure(orc_t *o)
{
    for (int x = 0; x < o->N; ++x) {
        int deg = o->row_ptr[x + 1] - o->row_ptr[x];
        double *t = o->Q + o->toff[x];
        for (int k = 0; k < deg; ++k) t[(int64_t)x * deg + k] = 0.0;
        for (int j = 0; j < deg; ++j) {
            int y = o->col[o->row_ptr[x] + j];
            for (int k = 0; k < deg; ++k) t[(int64_t)y * deg + k] = (j == k) ? -1.0 : -0.0;
        }
    }
}

/* CQ.clean (qroute.py:66-71): confidence = 0, then conf[links[x]] = eye(deg_x). */
void orc_cq_clean(orc_t *o)
{
    memset(o->Conf, 0, sizeof(double) * (size_t)o->tsize);
    for (int x = 0; x < o->N; ++x) {
        int deg = o->row_ptr[x + 1] - o->row_ptr[x];
        for (int j = 0; j < deg; ++j) {
            int y = o->col[o->row_ptr[x] + j];
            o->Conf[o->toff[x] + (int64_t)y * deg + j] = 1.0;
        }
    }
}

/* Network.reset (env.py:267-280) + Node.reset (env.py:99-101) + agent.clean():
 * Policy.clean is a no-op (base_policy.py:37-39); MaHybridQ.clean zeroes the traces
 * (multi_agent.py:48-50).  Tables are NOT reset. */
void orc_reset(orc_t *o)
{
    o->clock = o->all_packets = o->end_packets = o->drop_packets = 0;
    o->active_packets = o->hops = o->route_time = o->sends = 0;
    o->heap_len = 0; o->status = ST_OK; o->status_clock = 0;
    for (int x = 0; x < o->N; ++x) { o->queue[x].head = 0; o->queue[x].len = 0; }
    memset(o->sent, 0, sizeof(int) * (size_t)o->sdeg);
    if (o->policy == POL_MAHYBRIDQ) memset(o->Trace, 0, sizeof(double) * (size_t)o->tsize);
    if (o->policy == POL_CQ || o->policy == POL_CDRQ) orc_cq_clean(o);
}

/* Shortest.__init__ + _calc_distance (shortest.py:20-58), multiway=False.
 * Gauss-Seidel sweeps in (x, y in links[x], z) order; the strict `>` makes the stored next hop
 * "the neighbour through which distance[x,z] last strictly improved". */
void orc_shortest_build(orc_t *o)
{
    int N = o->N;
    memset(o->choice, 0, (size_t)o->tsize);
    for (int i = 0; i < N * N; ++i) o->distance[i] = INFINITY;          /* shortest.py:24 */
    for (int x = 0; x < N; ++x) {                                        /* shortest.py:28-34 */
        int deg = o->row_ptr[x + 1] - o->row_ptr[x];
        o->distance[x * N + x] = 0;
        for (int j = 0; j < deg; ++j) {
            int y = o->col[o->row_ptr[x] + j];
            o->distance[x * N + y] = 1;
            o->choice[o->toff[x] + (int64_t)y * deg + j] = 1;
        }
    }
    int changing = 1;                                                    /* shortest.py:43-58 */
    while (changing) {
        changing = 0;
        for (int x = 0; x < N; ++x) {
            int deg = o->row_ptr[x + 1] - o->row_ptr[x];
            for (int j = 0; j < deg; ++j) {
                int y = o->col[o->row_ptr[x] + j];
                for (int z = 0; z < N; ++z) {
                    double nd = o->distance[y * N + z] + 1.0;            /* unit(y) == 1, shortest.py:35 */
                    if (o->distance[x * N + z] > nd) {
                        o->distance[x * N + z] = nd;
                        uint8_t *c = o->choice + o->toff[x] + (int64_t)z * deg;
                        memset(c, 0, (size_t)deg);
                        c[j] = 1;
                        changing = 1;
                    }
                    if (o->multiway && isfinite(nd) && o->distance[x * N + z] == nd)      /* shortest.py:57-58 */
                        o->choice[o->toff[x] + (int64_t)z * deg + j] = 1;
                }
            }
        }
    }
}

/* GlobalRoute (shortest.py:84-104).  `_calc_distance` (shortest.py:43-58) with mask = everything off the diagonal
 * (shortest.py:87-88: all distances restart from inf) and unit(y) = 1 + queue_size[y] (shortest.py:90).  The
 * caller clears `choice` first, as GlobalRoute.learn does (shortest.py:101-103). */
void orc_globalroute_calc(orc_t *o)
{
    int N = o->N;
    for (int x = 0; x < N; ++x)
        for (int z = 0; z < N; ++z)
            if (x != z) o->distance[x * N + z] = INFINITY;               /* distance[mask] = inf */
    int changing = 1;
    while (changing) {
        changing = 0;
        for (int x = 0; x < N; ++x) {
            int deg = o->row_ptr[x + 1] - o->row_ptr[x];
            for (int j = 0; j < deg; ++j) {
                int y = o->col[o->row_ptr[x] + j];
                double unit = (double)(1 + o->queue_size[y]);
                for (int z = 0; z < N; ++z) {
                    double nd = o->distance[y * N + z] + unit;
                    if (o->distance[x * N + z] > nd) {
                        o->distance[x * N + z] = nd;
                        uint8_t *c = o->choice + o->toff[x] + (int64_t)z * deg;
                        memset(c, 0, (size_t)deg);
                        c[j] = 1;
                        changing = 1;
                    }
                    if (o->multiway && isfinite(nd) && o->distance[x * N + z] == nd)     /* shortest.py:57-58 */
                        o->choice[o->toff[x] + (int64_t)z * deg + j] = 1;
                }
            }
        }
    }
}

/* GlobalRoute.__init__ (shortest.py:85-91): Shortest.__init__ (tables with unit 1), then the mask is widened, the
 * queue sizes zeroed and the distances recomputed WITHOUT clearing `choice` first. */
void orc_globalroute_init(orc_t *o)
{
    orc_shortest_build(o);
    memset(o->queue_size, 0, sizeof(int64_t) * (size_t)o->N);
    orc_globalroute_calc(o);
}

int64_t *orc_queue_size(orc_t *o) { return o->queue_size; }

/* ---- heapq with keys compared by arrive_time only (CPython Lib/heapq.py restated) ---- */
static inline int ev_lt(const event_t *a, const event_t *b) { return a->arrive < b->arrive; }  /* env.py:48-49 */

static void hq_siftdown(event_t *h, int startpos, int pos)
{
    event_t newitem = h[pos];
    while (pos > startpos) {
        int parentpos = (pos - 1) >> 1;
        if (ev_lt(&newitem, &h[parentpos])) { h[pos] = h[parentpos]; pos = parentpos; continue; }
        break;
    }
    h[pos] = newitem;
}

static void hq_siftup(event_t *h, int endpos, int pos)
{
    int startpos = pos;
    event_t newitem = h[pos];
    int childpos = 2 * pos + 1;
    while (childpos < endpos) {
        int rightpos = childpos + 1;
        if (rightpos < endpos && !ev_lt(&h[childpos], &h[rightpos])) childpos = rightpos;
        h[pos] = h[childpos];
        pos = childpos;
        childpos = 2 * pos + 1;
    }
    h[pos] = newitem;
    hq_siftdown(h, startpos, pos);
}

static void hq_push(orc_t *o, const event_t *e)
{
    if (o->heap_len == o->heap_cap) {
        o->heap_cap *= 2;
        o->heap = (event_t *)realloc(o->heap, sizeof(event_t) * o->heap_cap);
    }
    o->heap[o->heap_len++] = *e;
    hq_siftdown(o->heap, 0, o->heap_len - 1);
}

static event_t hq_pop(orc_t *o)
{
    event_t last = o->heap[--o->heap_len];
    if (o->heap_len > 0) {
        event_t ret = o->heap[0];
        o->heap[0] = last;
        hq_siftup(o->heap, o->heap_len, 0);
        return ret;
    }
    return last;
}

/* Exported so tests can derive the same-key pop permutation independently of the GPU table. */
void orc_heap_perm(int k, int *perm_out)
{
    orc_t tmp; memset(&tmp, 0, sizeof(tmp));
    tmp.heap_cap = k + 1; tmp.heap = (event_t *)malloc(sizeof(event_t) * tmp.heap_cap);
    for (int i = 0; i < k; ++i) { event_t e; memset(&e, 0, sizeof(e)); e.from = i; e.arrive = 1; hq_push(&tmp, &e); }
    for (int i = 0; i < k; ++i) perm_out[i] = hq_pop(&tmp).from;
    free(tmp.heap);
}


/* ------------------------------------------------------------------------------------------
 * Deterministic exp / log2.
 * hybrid.py:17 calls np.exp and hybrid.py:39 np.log2.  NumPy's SIMD kernels, glibc and CUDA's
 * libdevice each round these differently in the last bit, and MaHybridQ's dynamics amplify a
 * 1-ulp difference exponentially (measured: 1e-15 -> 1e-2 in Theta over 1,700 steps on 6x6), so
 * no two libraries keep the same trajectory.  The oracle and the GPU kernels therefore share ONE
 * published algorithm built only from IEEE-754 +,-,*,/ and rint (no FMA), which gives the same
 * bits on every machine; it is within 2 ulp of the correctly rounded value (checked against
 * libm in tests/test_oracle_golden.py).  ref_harness.py can patch np.exp/np.log2 with these two
 * functions from outside, which makes the unmodified reference bit-comparable with the oracle.
 *   exp : k = rint(x/ln2); r = x - k*ln2_hi - k*ln2_lo; Taylor degree 13 (Horner); *2^k in two
 *         exact-or-once-rounded steps.
 *   log2: x = m*2^e, m in (sqrt(.5), sqrt(2)]; s=(m-1)/(m+1); ln m = 2s*sum_{n<=12} s^2n/(2n+1);
 *         log2 x = e + ln m / ln 2.
 * ------------------------------------------------------------------------------------------ */
static inline double u64_as_double(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static inline uint64_t double_as_u64(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }
static inline double pow2i(int k) { return u64_as_double((uint64_t)(k + 1023) << 52); }  /* -1022<=k<=1023 */

double orc_exp(double x)
{
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.2) return 0.0;
    double kf = rint(x * 1.4426950408889634);
    double r = x - kf * 0.6931471803691238;        /* ln2_hi: 21 trailing zero bits, product exact */
    r = r - kf * 1.9082149292705877e-10;           /* ln2_lo */
    double p = 1.0 / 6227020800.0;                 /* 1/13! */
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    int k = (int)kf, k1 = k / 2, k2 = k - k1;
    return (p * pow2i(k1)) * pow2i(k2);
}

double orc_log2(double x)
{
    if (x != x || x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    int e = 0;
    if (x < 2.2250738585072014e-308) { x = x * 18014398509481984.0; e = -54; }
    uint64_t b = double_as_u64(x);
    e += (int)(b >> 52) - 1023;
    double m = u64_as_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0), z = s * s;
    double q = 1.0 / 25.0;
    q = q * z + 1.0 / 23.0;
    q = q * z + 1.0 / 21.0;
    q = q * z + 1.0 / 19.0;
    q = q * z + 1.0 / 17.0;
    q = q * z + 1.0 / 15.0;
    q = q * z + 1.0 / 13.0;
    q = q * z + 1.0 / 11.0;
    q = q * z + 1.0 / 9.0;
    q = q * z + 1.0 / 7.0;
    q = q * z + 1.0 / 5.0;
    q = q * z + 1.0 / 3.0;
    q = q * z + 1.0;
    double lnm = (2.0 * s) * q;
    return (double)e + lnm * 1.4426950408889634;
}

/* ---- numpy reductions restated (numpy/core/src/umath/loops_utils.h.src pairwise sum,
 *      verified against ndarray.sum() for n < 128 in tests/test_oracle_golden.py) ---- */
static double np_sum(const double *a, int n)
{
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r += a[i]; return r; }
    double r[8]; int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

#define MAXDEG 64

/* PolicyGradient._softmax (hybrid.py:16-18): exp(theta)/exp(theta).sum(), no max-subtraction. */
static void softmax_row(const double *theta, int deg, double *out)
{
    double e[MAXDEG];
    for (int j = 0; j < deg; ++j) e[j] = orc_exp(theta[j]);
    double s = np_sum(e, deg);
    for (int j = 0; j < deg; ++j) out[j] = e[j] / s;
}

/* Node.receive (env.py:121-133) + Network.end_packet (env.py:321-331). */
static void node_receive(orc_t *o, int x, pkt_t *p)
{
    if (x == p->dest) {
        o->active_packets -= 1;
        o->end_packets += 1;
        if (o->sample_size > 0) {                       /* env.py:325-328 */
            int64_t at = o->sample_idx < o->sample_size - 1 ? o->sample_idx : o->sample_size - 1;
            o->sample[at] = o->clock - p->birth;
            o->sample_idx += 1;
        }
        o->route_time += o->clock - p->birth;
        o->hops += p->hops;
    } else {
        p->start_queue = (int32_t)o->clock;
        fifo_push(&o->queue[x], p);
        o->queue_size[x] += 1;                           /* agent.receive (env.py:131-133, shortest.py:93-94) */
    }
}

/* Network.new_packet (env.py:298-312) + inject (env.py:314-319).
 * Trace mode: the (count, (src,dst)...) of this step were drawn by NumPy's legacy stream on the
 * host exactly as env.py:307-310 would, so only the enqueue is restated here.
 * Philox mode: Poisson splitting — node x draws n_x ~ Poisson(lambda/N) by the multiplication
 * method (NumPy legacy `random_poisson_mult`: count uniforms until their product <= exp(-lam)),
 * each packet's dest uniform over the other N-1 nodes; distributionally identical to
 * env.py:307-310 (SURVEY §8c law ii).  `enlam` = exp(-lambda/N), computed once by the caller so
 * CPU and GPU compare the same double. */
static void inject_step(orc_t *o, int mode, const int32_t *pairs, int cnt, double enlam)
{
    if (mode == INJ_TRACE) {
        o->all_packets += cnt; o->active_packets += cnt;
        for (int i = 0; i < cnt; ++i) {
            pkt_t p = { pairs[2 * i + 1], pairs[2 * i], 0, (int32_t)o->clock, 0 };
            node_receive(o, p.source, &p);
        }
        return;
    }
    for (int x = 0; x < o->N; ++x) {
        pstream_t s; ps_init(&s, o->seed, o->replica, o->clock, x, 0);
        int n = 0; double prod = 1.0;
        for (;;) {
            double u = ((double)ps_next(&s) + 0.5) * (1.0 / 4294967296.0);
            prod = prod * u;
            if (prod > enlam) n++; else break;
        }
        o->all_packets += n; o->active_packets += n;
        for (int i = 0; i < n; ++i) {
            uint32_t w = ps_next(&s);
            int d = (int)(((uint64_t)w * (uint32_t)(o->N - 1)) >> 32);
            if (d >= x) d++;
            pkt_t p = { d, x, 0, (int32_t)o->clock, 0 };
            node_receive(o, x, &p);
        }
    }
}

/* Node._send_default (env.py:162-191) for every node in ID order (env.py:343-347), with the
 * policy callbacks choose / get_info inlined:
 *   Shortest.choose  shortest.py:38-41 (random=False: first True of choice[x][d])
 *   Qroute.choose    qroute.py:22-27   (np.argmax: first maximum)
 *   PolicyGradient.choose hybrid.py:20-23 (np.random.choice(links, p=softmax): legacy
 *       RandomState.choice = cdf=cumsum(p); cdf/=cdf[-1]; idx=searchsorted(cdf,u,'right');
 *       u comes from the Philox stream (purpose 1) instead of the global MT19937)
 *   Qroute.get_info  qroute.py:29-30 ; HybridQ.get_info hybrid.py:49-53
 *   _build_info_default env.py:147-152 ; _send_packet env.py:135-145                        */
/* agent.choose(x, d) for the built-in policies; returns the action index, or -1 when the softmax row holds a NaN
 * (np.random.choice raises ValueError).  `s` is node x's sampling stream of this step (purpose 1), opened on first
 * use: a scan that calls choose several times (env.py:174-190) keeps drawing from the same stream. */
static int choose_action(orc_t *o, int x, int d, pstream_t *s, int *s_open)
{
    int deg = o->row_ptr[x + 1] - o->row_ptr[x], k = 0;
    if (o->policy == POL_SHORTEST || o->policy == POL_GLOBALROUTE) {
        const uint8_t *c = o->choice + o->toff[x] + (int64_t)d * deg;
        while (k < deg && !c[k]) k++;
        if (o->random_choice) {
            /* shortest.py:40-41 np.random.choice(choices): a uniform index into the True entries; the index comes
             * from the Philox stream (purpose 1) as mulhi(word, count), nothing is drawn for a single choice */
            int cnt = 0;
            for (int j = 0; j < deg; ++j) cnt += c[j];
            if (cnt > 1) {
                if (!*s_open) { ps_init(s, o->seed, o->replica, o->clock, x, 1); *s_open = 1; }
                int idx = (int)(((uint64_t)ps_next(s) * (uint32_t)cnt) >> 32);
                for (k = 0; k < deg; ++k) if (c[k] && idx-- == 0) break;
            }
        }
    } else if (o->policy == POL_QROUTE || o->policy == POL_CQ || o->policy == POL_CDRQ) {
        const double *row = o->Q + o->toff[x] + (int64_t)d * deg;
        for (int j = 1; j < deg; ++j) if (row[j] > row[k]) k = j;
    } else {
        double sm[MAXDEG], cdf[MAXDEG];
        softmax_row(o->Theta + o->toff[x] + (int64_t)d * deg, deg, sm);
        int bad = 0; double c = 0.0;
        for (int j = 0; j < deg; ++j) { if (isnan(sm[j])) bad = 1; c += sm[j]; cdf[j] = c; }
        if (bad) return -1;
        for (int j = 0; j < deg; ++j) cdf[j] /= c;
        if (!*s_open) { ps_init(s, o->seed, o->replica, o->clock, x, 1); *s_open = 1; }
        uint32_t a = ps_next(s) >> 5, b = ps_next(s) >> 6;
        double u = ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
        k = 0; while (k < deg - 1 && cdf[k] <= u) k++;
    }
    return k;
}

static void send_phase(orc_t *o)
{
    o->n_rewards = 0;
    for (int x = 0; x < o->N; ++x) {
        fifo_t *q = &o->queue[x];
        if (q->len == 0) continue;
        int deg = o->row_ptr[x + 1] - o->row_ptr[x];
        /* env.py:171-190: avaliable_path is evaluated once; scan the queue in order for the first packet whose chosen
         * link has sent[action] < bandwidth.  Under the default timing every link is free and i stays 0. */
        int any_free = 0;
        for (int j = 0; j < deg; ++j) any_free |= (o->sent[o->row_ptr[x] + j] < o->bandwidth);
        pstream_t ps; int ps_open = 0;
        int64_t i = 0; int k = -1;
        while (i < q->len && any_free) {
            k = choose_action(o, x, q->buf[q->head + i].dest, &ps, &ps_open);
            o->probes++;
            if (k < 0) { o->status = ST_NAN_PROB; o->status_clock = o->clock; return; }
            if (o->sent[o->row_ptr[x] + k] < o->bandwidth) break;
            k = -1; i++;
        }
        if (k < 0) continue;
        pkt_t p = q->buf[q->head + i];
        int d = p.dest;
        int y = o->col[o->row_ptr[x] + k];
        if (i > 0) memmove(q->buf + q->head + 1, q->buf + q->head, (size_t)i * sizeof(pkt_t));   /* queue.pop(i) */
        q->head++; q->len--;
        o->queue_size[x] -= 1;                           /* agent.send (env.py:181, shortest.py:96-97) */
        p.hops += 1;
        o->sent[o->row_ptr[x] + k] += 1;                 /* env.py:142 */
        event_t e; e.p = p; e.from = x; e.to = y; e.link = o->row_ptr[x] + k; e.arrive = o->clock + o->transtime;
        hq_push(o, &e);
        o->sends += 1;
        reward_t *r = &o->rewards[o->n_rewards++];
        r->x = x; r->y = y; r->d = d; r->k = k; r->src = p.source; r->q_y = o->clock - p.start_queue;
        r->max_Q_y = r->max_Q_x_d = 0.0;
        r->q_x = 0; r->C_f = r->C_b = r->max_Q_b = 0.0;
        if (o->policy == POL_CQ || o->policy == POL_CDRQ) {
            /* CQ.get_info (qroute.py:72-78): z_idx, max_Q_f = choose(y, dest, idx=True); C_f = conf[y][dest][z_idx] */
            int degy = o->row_ptr[y + 1] - o->row_ptr[y];
            const double *ry = o->Q + o->toff[y] + (int64_t)d * degy;
            int z = 0; for (int j = 1; j < degy; ++j) if (ry[j] > ry[z]) z = j;
            r->max_Q_y = ry[z];
            r->C_f = o->Conf[o->toff[y] + (int64_t)d * degy + z];
            if (o->policy == POL_CDRQ) {
                /* CDRQ.get_info (qroute.py:107-115) + Node._build_info_dual (env.py:154-160): queue lengths as they
                 * are at this point of the sequential node loop (own packet already popped) */
                const double *rb = o->Q + o->toff[x] + (int64_t)p.source * deg;
                int w = 0; for (int j = 1; j < deg; ++j) if (rb[j] > rb[w]) w = j;
                r->max_Q_b = rb[w];
                r->C_b = o->Conf[o->toff[x] + (int64_t)p.source * deg + w];
                r->q_y = o->queue[y].len > 1 ? o->queue[y].len : 1;
                r->q_x = q->len > 1 ? q->len : 1;
            }
        } else if (o->policy != POL_SHORTEST && o->policy != POL_GLOBALROUTE) {
            int degy = o->row_ptr[y + 1] - o->row_ptr[y];
            const double *ry = o->Q + o->toff[y] + (int64_t)d * degy;
            double m = ry[0]; for (int j = 1; j < degy; ++j) if (ry[j] > m) m = ry[j];
            r->max_Q_y = m;
            const double *rx = o->Q + o->toff[x] + (int64_t)d * deg;
            m = rx[0]; for (int j = 1; j < deg; ++j) if (rx[j] > m) m = rx[j];
            r->max_Q_x_d = m;
        }
    }
}

/* Event drain of Network.step (env.py:349-364). */
static void drain_events(orc_t *o)
{
    int64_t end_time = o->clock + o->slot;
    while (o->heap_len > 0 && o->heap[0].arrive <= end_time) {
        event_t e = hq_pop(o);
        o->sent[e.link] -= 1;                            /* env.py:352 */
        if (o->is_drop && e.p.hops >= o->N) {          /* env.py:355-360; drop_penalty is a no-op for every in-scope agent */
            o->drop_packets += 1;
            o->active_packets -= 1;
            continue;
        }
        o->clock = e.arrive;
        node_receive(o, e.to, &e.p);
    }
    o->clock = end_time;
}

/* Qroute._update_qtable (qroute.py:40-44): Q += lr*((r + discount*max_Q_y) - old). */
static inline void q_update(orc_t *o, const reward_t *rw, double r, double lr)
{
    int deg = o->row_ptr[rw->x + 1] - o->row_ptr[rw->x];
    double *cell = o->Q + o->toff[rw->x] + (int64_t)rw->d * deg + rw->k;
    double old = *cell;
    double t = o->discount * rw->max_Q_y;   /* built with -ffp-contract=off: no FMA */
    double u = r + t;
    double v = u - old;
    double w = lr * v;
    *cell = old + w;
    if (o->q_f32) *cell = (double)(float)*cell;
}

/* agent.learn(rewards, lr) for the four in-scope policies. */
static void learn_phase(orc_t *o, double lr_q, double lr_p, double lr_e)
{
    int n = o->n_rewards;
    if (o->policy == POL_SHORTEST) return;                       /* Policy.learn: no-op */
    if (o->policy == POL_GLOBALROUTE) {                          /* shortest.py:99-104 */
        memset(o->choice, 0, (size_t)o->tsize);
        orc_globalroute_calc(o);
        return;
    }
    if (o->policy == POL_CQ || o->policy == POL_CDRQ) {          /* qroute.py:80-101, 117-122 */
        for (int i = 0; i < n; ++i) {
            const reward_t *rw = &o->rewards[i];
            for (int pass = 0; pass < (o->policy == POL_CDRQ ? 2 : 1); ++pass) {
                /* forward: table x, row dest, column idx(y); backward (CDRQ): table y, row packet.source, column idx(x) */
                int tx = pass == 0 ? rw->x : rw->y, ty = pass == 0 ? rw->y : rw->x, row = pass == 0 ? rw->d : rw->src;
                double C = pass == 0 ? rw->C_f : rw->C_b, maxQ = pass == 0 ? rw->max_Q_y : rw->max_Q_b;
                double r = pass == 0 ? (double)(-rw->q_y - (o->policy == POL_CDRQ ? 0 : 1)) : (double)(-rw->q_x);
                int deg = o->row_ptr[tx + 1] - o->row_ptr[tx];
                int64_t cell = o->toff[tx] + (int64_t)row * deg + o->aidx[tx * o->N + ty];
                double old_Q = o->Q[cell], old_conf = o->Conf[cell];
                double one_minus = 1.0 - old_conf;
                double eta = (one_minus > C) ? one_minus : C;            /* max(C, 1 - old_conf) */
                double t = o->discount * maxQ;
                double u = r + t;
                double v = u - old_Q;
                double w = eta * v;
                o->Q[cell] = old_Q + w;
                double dc = C - old_conf;
                double ec = eta * dc;
                double nc = old_conf + ec;
                o->Conf[cell] = nc / o->decay;
            }
        }
        for (int64_t i = 0; i < o->tsize; ++i) o->Conf[i] = o->Conf[i] * o->decay;   /* confidence_decay, qroute.py:99-101 */
        return;
    }
    if (o->policy == POL_QROUTE) {                               /* qroute.py:32-53 */
        for (int i = 0; i < n; ++i) {
            const reward_t *rw = &o->rewards[i];
            double r = (double)(-rw->q_y - 1);                   /* qroute.py:37, t_y = 1 */
            q_update(o, rw, r, lr_q);
        }
        return;
    }
    if (o->policy == POL_HYBRIDQ) {                              /* hybrid.py:55-62 via qroute.py:51-53 */
        for (int i = 0; i < n; ++i) {
            const reward_t *rw = &o->rewards[i];
            int deg = o->row_ptr[rw->x + 1] - o->row_ptr[rw->x];
            double *th = o->Theta + o->toff[rw->x] + (int64_t)rw->d * deg;
            double sm[MAXDEG], tmp[MAXDEG];
            softmax_row(th, deg, sm);
            double r = (double)(-rw->q_y - 1);
            if (o->add_entropy) {                                /* hybrid.py:38-39 */
                for (int j = 0; j < deg; ++j) tmp[j] = sm[j] * orc_log2(sm[j]);
                double ent = np_sum(tmp, deg);
                double le = lr_e * ent;
                r = r - le;
            }
            q_update(o, rw, r, lr_q);
            double t = o->discount * rw->max_Q_y;       /* hybrid.py:32-36 */
            double u = r + t;
            double adv = u - rw->max_Q_x_d;
            for (int j = 0; j < deg; ++j) {
                double g = -sm[j]; if (j == rw->k) g += 1.0;     /* hybrid.py:25-30 */
                double a = lr_p * g;
                double b = a * adv;
                th[j] = th[j] + b;
            }
        }
        return;
    }
    /* MaHybridQ.learn (multi_agent.py:17-43) */
    double *rs = o->scratch, *qy = o->scratch + o->N, *qx = o->scratch + 2 * o->N;
    for (int i = 0; i < n; ++i) {
        rs[i] = (double)(-o->rewards[i].q_y - 1);
        qy[i] = o->rewards[i].max_Q_y; qx[i] = o->rewards[i].max_Q_x_d;
    }
    double d1 = np_sum(rs, n) + o->reward_shape;        /* multi_agent.py:28-29 */
    double d2 = o->discount * np_sum(qy, n);
    double d3 = d1 + d2;
    double delta = d3 - np_sum(qx, n);
    o->reward_shape = 0;
    for (int64_t i = 0; i < o->tsize; ++i) o->Trace[i] = o->Trace[i] * o->discount_trace;  /* :32-33 */
    for (int i = 0; i < n; ++i) {
        const reward_t *rw = &o->rewards[i];
        int deg = o->row_ptr[rw->x + 1] - o->row_ptr[rw->x];
        int64_t off = o->toff[rw->x] + (int64_t)rw->d * deg;
        double sm[MAXDEG];
        softmax_row(o->Theta + off, deg, sm);
        for (int j = 0; j < deg; ++j) {                          /* :35-36 */
            double g = -sm[j]; if (j == rw->k) g += 1.0;
            o->Trace[off + j] = o->Trace[off + j] + g;
        }
        q_update(o, rw, rs[i], lr_q);                            /* :38-40 */
    }
    double ld = lr_p * delta;                           /* :42-43 */
    for (int64_t i = 0; i < o->tsize; ++i) {
        double b = ld * o->Trace[i];
        o->Theta[i] = o->Theta[i] + b;
    }
}

/* Network.train (env.py:367-414) for `n_steps` iterations (inject once, then `freq` calls of step(slot) + learn).
 *   inj_cnt[n_steps], inj_pairs[2*sum(cnt)]  trace-mode injections (NULL in Philox mode)
 *   enlam                                     exp(-lambda/N) for Philox mode
 *   m_route_time/m_end/m_hops[n_steps]        cumulative counters after each step (env.py:409-413)
 *   sendlog[4*sendlog_cap]                    rows (clock_after_step, x, y, dest) per send
 * Returns the number of steps completed (< n_steps only if the run stopped on ST_NAN_PROB,
 * where the reference raises ValueError out of np.random.choice). */
int64_t orc_run(orc_t *o, int64_t n_steps, int inj_mode, const int32_t *inj_cnt,
                const int32_t *inj_pairs, double enlam, double lr_q, double lr_p, double lr_e,
                int64_t *m_route_time, int64_t *m_end, int64_t *m_hops,
                int32_t *sendlog, int64_t sendlog_cap, int64_t *n_sendlog)
{
    int64_t off = 0, nlog = 0;
    for (int64_t s = 0; s < n_steps; ++s) {
        int cnt = inj_mode == INJ_TRACE ? inj_cnt[s] : 0;
        inject_step(o, inj_mode, inj_pairs ? inj_pairs + 2 * off : NULL, cnt, enlam);
        off += cnt;
        for (int f = 0; f < o->freq; ++f) {              /* env.py:402-408: freq calls of step(slot) + learn per injection */
            int64_t sends_before = o->sends;
            send_phase(o);
            if (o->status != ST_OK) { o->sends = sends_before; if (n_sendlog) *n_sendlog = nlog; return s; }
            if (sendlog) {
                for (int i = 0; i < o->n_rewards && nlog < sendlog_cap; ++i, ++nlog) {
                    sendlog[4 * nlog + 0] = (int32_t)(o->clock + o->slot);
                    sendlog[4 * nlog + 1] = o->rewards[i].x;
                    sendlog[4 * nlog + 2] = o->rewards[i].y;
                    sendlog[4 * nlog + 3] = o->rewards[i].d;
                }
            }
            drain_events(o);
            learn_phase(o, lr_q, lr_p, lr_e);
        }
        if (m_route_time) m_route_time[s] = o->route_time;
        if (m_end) m_end[s] = o->end_packets;
        if (m_hops) m_hops[s] = o->hops;
    }
    if (n_sendlog) *n_sendlog = nlog;
    return n_steps;
}

/* Network.sample_route_time (env.py:416-438): runs whole iterations until `size` packets have been
 * delivered; sample[i] is the routing time of the i-th delivery (deliveries past `size` in the last step overwrite the
 * last entry).  Philox mode if inj_cnt is NULL; in trace mode the caller supplies at least `max_steps` steps of draws.
 * Returns the number of steps run (or -1 if max_steps was not enough). */
int64_t orc_sample_route_time(orc_t *o, int64_t size, int64_t max_steps, const int32_t *inj_cnt, const int32_t *inj_pairs,
                              double enlam, double lr_q, double lr_p, double lr_e, int64_t *sample_out)
{
    o->sample = sample_out; o->sample_size = size; o->sample_idx = 0;
    for (int64_t i = 0; i < size; ++i) sample_out[i] = 0;
    int64_t off = 0, steps = 0;
    while (o->sample_idx < size) {
        if (steps >= max_steps) { o->sample_size = 0; o->sample = NULL; return -1; }
        int cnt = inj_cnt ? inj_cnt[steps] : 0;
        inject_step(o, inj_cnt ? INJ_TRACE : INJ_PHILOX, inj_pairs ? inj_pairs + 2 * off : NULL, cnt, enlam);
        off += cnt;
        for (int f = 0; f < o->freq && o->status == ST_OK; ++f) {
            send_phase(o);
            if (o->status != ST_OK) break;
            drain_events(o);
            learn_phase(o, lr_q, lr_p, lr_e);
        }
        if (o->status != ST_OK) break;
        steps++;
    }
    o->sample_size = 0; o->sample = NULL;
    return steps;
}

/* Runs `nrep` independent replicas (one orc_t each) for n_steps in Philox mode, one OpenMP
 * thread per replica chunk.  Used as the all-host-cores CPU baseline (bench.py) and for
 * bulk parity checks; per-replica counters are read back with orc_counters. */
int64_t orc_run_many(orc_t **os, int nrep, int64_t n_steps, const double *enlam,
                     double lr_q, double lr_p, double lr_e, int nthreads,
                     int64_t *m_route_time, int64_t *m_end)
{
    int64_t total_sends = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_sends)
    for (int r = 0; r < nrep; ++r) {
        int64_t before = os[r]->sends;
        orc_run(os[r], n_steps, INJ_PHILOX, NULL, NULL, enlam[r], lr_q, lr_p, lr_e,
                m_route_time ? m_route_time + (int64_t)r * n_steps : NULL,
                m_end ? m_end + (int64_t)r * n_steps : NULL, NULL, NULL, 0, NULL);
        total_sends += os[r]->sends - before;
    }
    return total_sends;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
