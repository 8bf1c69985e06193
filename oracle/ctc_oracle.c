/* TEST INFRASTRUCTURE ONLY -- CPU oracle.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product path never does.
 *
 * Restates, in plain C, the TensorFlow-1.x CTC kernels the reference calls:
 *   tf.nn.ctc_loss                 /root/reference/src/weinman/model.py:224-229
 *   tf.nn.ctc_greedy_decoder       /root/reference/src/weinman/validate.py:81-92
 *   tf.nn.ctc_beam_search_decoder  /root/reference/src/weinman/test.py:84-88 (merge_repeated=True)
 *                                  /root/reference/src/weinman/client.py:227-231 (merge_repeated=False)
 *   tf.edit_distance               /root/reference/src/weinman/test.py:90
 * The arithmetic lives in un-vendored TensorFlow (README.md:62, no pinned version); the
 * algorithm restated here is upstream's tensorflow/core/util/ctc/{ctc_loss_calculator,
 * ctc_beam_search,ctc_beam_entry}.h and core/kernels/ctc_decoder_ops.cc.
 *
 * PARITY PIN: the reference holds no golden vectors for this path.  This oracle is pinned on
 * upstream TensorFlow's own unit-test vectors (ctc_loss_op_test.py testBasic,
 * ctc_decoder_ops_test.py testCTCGreedyDecoder / testCTCDecoderBeamSearch) committed under
 * tests/golden/, on brute-force path enumeration, and on torch.nn.functional.ctc_loss.
 * It has never been run against a live TensorFlow: parity with TF itself is "unpinned".
 *
 * Layouts: logits [T,B,C] float32 time-major; blank = C-1; labels flat int32 with offsets[B+1].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include "det_math.h"

/* minimal parallel-for over the batch (pthreads; this image's gcc has no usable libgomp) */
typedef void (*pf_body)(int b, void* ctx);
typedef struct { pf_body fn; void* ctx; int n; volatile int* next; } pf_job;
static void* pf_worker(void* arg) {
    pf_job* j = (pf_job*)arg;
    for (;;) {
        int b = __sync_fetch_and_add(j->next, 1);
        if (b >= j->n) break;
        j->fn(b, j->ctx);
    }
    return NULL;
}
static void parallel_for(int n, int nthreads, pf_body fn, void* ctx) {
    volatile int next = 0;
    pf_job job; job.fn = fn; job.ctx = ctx; job.n = n; job.next = &next;
    if (nthreads > n) nthreads = n;
    if (nthreads <= 1) { pf_worker(&job); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    int i, started = 0;
    for (i = 0; i < nthreads - 1; ++i) if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) started++;
    pf_worker(&job);
    for (i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(th);
}

#define LOG0 (-INFINITY)

/* ctc_loss_util.h LogSumExp restated (libm flavour, as TF computes it). */
static inline float lse_libm(float a, float b) {
    if (a == LOG0 && b == LOG0) return LOG0;
    return (a > b) ? a + log1pf(expf(b - a)) : b + log1pf(expf(a - b));
}

/* ------------------------------------------------------------------ CTC loss + gradient */

/* One example.  x: logits of this example, row t at x + t*stride.  Follows
 * CTCLossCalculator::CalculateLoss / CalculateForwardVariables / CalculateBackwardVariables /
 * CalculateGradient, in float32 with libm exp/log/log1p exactly as TF evaluates it. */
static int ctc_one(const float* x, long stride, int T, int C, const int* lab, int L,
                   float* loss, float* grad, long gstride, int Tfull, float* y, float* la, float* lb)
{
    const int blank = C - 1;
    const int U = 2 * L + 1;
    int t, u, k;
    /* gradient rows beyond the sequence length are zero */
    if (grad) for (t = 0; t < Tfull; ++t) memset(grad + t * gstride, 0, sizeof(float) * C);
    *loss = 0.0f;
    if (T == 0) return 0;
    /* required time: labels plus one blank between adjacent repeats */
    int need = L;
    for (u = 1; u < L; ++u) if (lab[u] == lab[u - 1]) need++;
    if (need > T) return 2; /* "Not enough time for target transition sequence" */

    /* softmax per frame, max-subtracted, linear domain (as TF) */
    for (t = 0; t < T; ++t) {
        const float* r = x + t * stride;
        float m = r[0];
        for (k = 1; k < C; ++k) if (r[k] > m) m = r[k];
        float s = 0.0f;
        for (k = 0; k < C; ++k) { float e = expf(r[k] - m); y[t * C + k] = e; s += e; }
        for (k = 0; k < C; ++k) y[t * C + k] /= s;
    }
#define LP(uu) ((uu) & 1 ? lab[(uu) >> 1] : blank)
#define LY(uu, tt) logf(y[(tt) * C + LP(uu)])
    for (u = 0; u < U * T; ++u) { la[u] = LOG0; lb[u] = LOG0; }
    /* forward */
    la[0 * T + 0] = LY(0, 0);
    if (U > 1) la[1 * T + 0] = LY(1, 0);
    for (t = 1; t < T; ++t) {
        int lo = U - 2 * (T - t); if (lo < 0) lo = 0;
        int hi = 2 * (t + 1); if (hi > U) hi = U;
        for (u = lo; u < hi; ++u) {
            float s = la[u * T + t - 1];
            if (u > 0) s = lse_libm(s, la[(u - 1) * T + t - 1]);
            if (u > 1 && LP(u) != blank && LP(u) != LP(u - 2)) s = lse_libm(s, la[(u - 2) * T + t - 1]);
            la[u * T + t] = LY(u, t) + s;
        }
    }
    /* backward (beta excludes y_t) */
    for (u = U - 2; u < U; ++u) if (u >= 0) lb[u * T + T - 1] = 0.0f;
    for (t = T - 2; t >= 0; --t) {
        int lo = U - 2 * (T - t); if (lo < 0) lo = 0;
        int hi = 2 * (t + 1); if (hi > U) hi = U;
        for (u = lo; u < hi; ++u) {
            float s = lse_libm(lb[u * T + t], lb[u * T + t + 1] + LY(u, t + 1));
            if (u + 1 < U) s = lse_libm(s, lb[(u + 1) * T + t + 1] + LY(u + 1, t + 1));
            if (u + 2 < U && LP(u) != blank && LP(u) != LP(u + 2))
                s = lse_libm(s, lb[(u + 2) * T + t + 1] + LY(u + 2, t + 1));
            lb[u * T + t] = s;
        }
    }
    /* log p(z|x) from the t=0 column, as CalculateLoss does */
    float logp = LOG0;
    for (u = 0; u < U; ++u) logp = lse_libm(logp, la[u * T + 0] + lb[u * T + 0]);
    *loss = -logp;
    if (!grad) return logp == LOG0 ? 1 : 0;
    if (logp == LOG0) { /* "No valid path found": dy = y */
        for (t = 0; t < T; ++t) for (k = 0; k < C; ++k) grad[t * gstride + k] = y[t * C + k];
        return 1;
    }
    float* ps = (float*)malloc(sizeof(float) * C);
    for (t = 0; t < T; ++t) {
        for (k = 0; k < C; ++k) ps[k] = LOG0;
        for (u = 0; u < U; ++u) { int l = LP(u); ps[l] = lse_libm(ps[l], la[u * T + t] + lb[u * T + t]); }
        for (k = 0; k < C; ++k) grad[t * gstride + k] = y[t * C + k] - expf(ps[k] - logp);
    }
    free(ps);
    return 0;
#undef LP
#undef LY
}

/* Same recursion evaluated in float64 (inputs/outputs float32): the "exact" answer used to state
 * tolerances -- the float32 log-domain recursion above carries ~1 ulp(|alpha|) ~ 1e-5 of noise. */
static inline double lse_d(double a, double b) {
    if (a == -INFINITY && b == -INFINITY) return -INFINITY;
    return (a > b) ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}
static int ctc_one_f64(const float* x, long stride, int T, int C, const int* lab, int L,
                       float* loss, float* grad, long gstride, int Tfull)
{
    const int blank = C - 1;
    const int U = 2 * L + 1;
    int t, u, k;
    if (grad) for (t = 0; t < Tfull; ++t) memset(grad + t * gstride, 0, sizeof(float) * C);
    *loss = 0.0f;
    if (T == 0) return 0;
    int need = L;
    for (u = 1; u < L; ++u) if (lab[u] == lab[u - 1]) need++;
    if (need > T) return 2;
    double* ly = (double*)malloc(sizeof(double) * (size_t)T * C); /* log softmax */
    double* la = (double*)malloc(sizeof(double) * (size_t)T * U);
    double* lb = (double*)malloc(sizeof(double) * (size_t)T * U);
    for (t = 0; t < T; ++t) {
        const float* r = x + t * stride;
        double m = r[0], s = 0.0;
        for (k = 1; k < C; ++k) if (r[k] > m) m = r[k];
        for (k = 0; k < C; ++k) s += exp((double)r[k] - m);
        for (k = 0; k < C; ++k) ly[t * C + k] = (double)r[k] - m - log(s);
    }
#define LP(uu) ((uu) & 1 ? lab[(uu) >> 1] : blank)
#define LY(uu, tt) ly[(tt) * C + LP(uu)]
    for (u = 0; u < U * T; ++u) { la[u] = -INFINITY; lb[u] = -INFINITY; }
    la[0] = LY(0, 0);
    if (U > 1) la[1 * T] = LY(1, 0);
    for (t = 1; t < T; ++t) for (u = 0; u < U; ++u) {
        double s = la[u * T + t - 1];
        if (u > 0) s = lse_d(s, la[(u - 1) * T + t - 1]);
        if (u > 1 && LP(u) != blank && LP(u) != LP(u - 2)) s = lse_d(s, la[(u - 2) * T + t - 1]);
        la[u * T + t] = LY(u, t) + s;
    }
    for (u = U - 2; u < U; ++u) if (u >= 0) lb[u * T + T - 1] = 0.0;
    for (t = T - 2; t >= 0; --t) for (u = 0; u < U; ++u) {
        double s = lb[u * T + t + 1] + LY(u, t + 1);
        if (u + 1 < U) s = lse_d(s, lb[(u + 1) * T + t + 1] + LY(u + 1, t + 1));
        if (u + 2 < U && LP(u) != blank && LP(u) != LP(u + 2)) s = lse_d(s, lb[(u + 2) * T + t + 1] + LY(u + 2, t + 1));
        lb[u * T + t] = s;
    }
    double logp = lse_d(la[(U - 1) * T + T - 1], U > 1 ? la[(U - 2) * T + T - 1] : -INFINITY);
    *loss = (float)(-logp);
    int st = 0;
    if (grad) {
        if (logp == -INFINITY) {
            for (t = 0; t < T; ++t) for (k = 0; k < C; ++k) grad[t * gstride + k] = (float)exp(ly[t * C + k]);
            st = 1;
        } else {
            double* ps = (double*)malloc(sizeof(double) * C);
            for (t = 0; t < T; ++t) {
                for (k = 0; k < C; ++k) ps[k] = 0.0;
                for (u = 0; u < U; ++u) ps[LP(u)] += exp(la[u * T + t] + lb[u * T + t] - logp);
                for (k = 0; k < C; ++k) grad[t * gstride + k] = (float)(exp(ly[t * C + k]) - ps[k]);
            }
            free(ps);
        }
    } else if (logp == -INFINITY) st = 1;
    free(ly); free(la); free(lb);
    return st;
#undef LP
#undef LY
}

/* status[b]: 0 ok, 1 no valid path (loss=+inf, grad=softmax), 2 label does not fit (TF raises
 * InvalidArgument; loss/grad left zero here).  Returns the number of non-zero statuses. */
typedef struct {
    const float* logits; int T, B, C; const int* labels; const int* offsets; const int* seq_len;
    float* loss; float* grad; int* status; int bad; int wide;
} loss_ctx;
static void loss_body(int b, void* vc) {
    loss_ctx* c = (loss_ctx*)vc;
    int L = c->offsets[b + 1] - c->offsets[b];
    int U = 2 * L + 1;
    int Tb = c->seq_len[b];
    int C = c->C;
    if (c->wide) {
        int st64 = ctc_one_f64(c->logits + (size_t)b * C, (long)c->B * C, Tb, C, c->labels + c->offsets[b], L,
                               c->loss + b, c->grad ? c->grad + (size_t)b * C : NULL, (long)c->B * C, c->T);
        if (c->status) c->status[b] = st64;
        if (st64) __sync_fetch_and_add(&c->bad, 1);
        return;
    }
    float* y = (float*)malloc(sizeof(float) * ((size_t)(Tb > 0 ? Tb : 1) * C));
    float* la = (float*)malloc(sizeof(float) * ((size_t)(Tb > 0 ? Tb : 1) * U));
    float* lb = (float*)malloc(sizeof(float) * ((size_t)(Tb > 0 ? Tb : 1) * U));
    int st = ctc_one(c->logits + (size_t)b * C, (long)c->B * C, Tb, C, c->labels + c->offsets[b], L, c->loss + b,
                     c->grad ? c->grad + (size_t)b * C : NULL, (long)c->B * C, c->T, y, la, lb);
    if (c->status) c->status[b] = st;
    if (st) __sync_fetch_and_add(&c->bad, 1);
    free(y); free(la); free(lb);
}
int oracle_ctc_loss(const float* logits, int T, int B, int C, const int* labels, const int* offsets,
                    const int* seq_len, float* loss, float* grad, int* status, int nthreads)
{
    loss_ctx c = { logits, T, B, C, labels, offsets, seq_len, loss, grad, status, 0, 0 };
    parallel_for(B, nthreads, loss_body, &c);
    return c.bad;
}
int oracle_ctc_loss_f64(const float* logits, int T, int B, int C, const int* labels, const int* offsets,
                        const int* seq_len, float* loss, float* grad, int* status, int nthreads)
{
    loss_ctx c = { logits, T, B, C, labels, offsets, seq_len, loss, grad, status, 0, 1 };
    parallel_for(B, nthreads, loss_body, &c);
    return c.bad;
}

/* ------------------------------------------------------------------ greedy decoder
 * CTCGreedyDecoderOp: per frame first-maximum arg-max (strict '>' scan from class 0),
 * neg_sum_logits += -max, emit unless blank or (merge_repeated and same as previous frame). */
int oracle_ctc_greedy(const float* logits, int T, int B, int C, const int* seq_len, int merge_repeated,
                      int64_t* decoded /*[B,T]*/, int* decoded_len /*[B]*/, float* neg_sum_logits /*[B]*/,
                      int nthreads)
{
    int b;
    const int blank = C - 1;
    (void)nthreads; /* a read-once scan: single thread is already memory-speed */
    for (b = 0; b < B; ++b) {
        int prev = -1, n = 0, t, k;
        float acc = 0.0f;
        for (t = 0; t < seq_len[b]; ++t) {
            const float* r = logits + ((size_t)t * B + b) * C;
            float p = r[0]; int c = 0;
            for (k = 1; k < C; ++k) if (r[k] > p) { p = r[k]; c = k; }
            acc += -p;
            if (c != blank && !(merge_repeated && c == prev)) decoded[(size_t)b * T + n++] = c;
            prev = c;
        }
        for (t = n; t < T; ++t) decoded[(size_t)b * T + t] = -1;
        decoded_len[b] = n;
        neg_sum_logits[b] = acc;
    }
    return 0;
}

/* ------------------------------------------------------------------ beam search decoder
 * Restates CTCBeamSearchDecoder<>::Step / TopPaths with BeamEntry prefix-tree nodes.
 * Tie rule (documented deviation, SURVEY.md section 7 hard part 2): TF orders equal-score beams
 * through gtl::TopN + libstdc++ heap internals (implementation-defined).  Here the order is
 * total: score descending, then push order ascending (branches in their sorted order first,
 * then children in (branch, class) order) -- a stable top-k.  On tie-free inputs it is TF's.
 * math: 0 = libm (what TF calls), 1 = det_math.h (bit-reproducible on the GPU). */
typedef struct Node {
    struct Node* parent;
    struct Node** children; /* C-1 slots, lazily allocated */
    int label;
    float o_total, o_blank, o_label; /* oldp */
    float n_total, n_blank, n_label; /* newp */
    long seq;                        /* push order within the current step */
    struct Node* pool_next;
} Node;

typedef struct { Node* all; int C; } Tree;

static Node* node_new(Tree* tr, Node* parent, int label) {
    Node* n = (Node*)calloc(1, sizeof(Node));
    n->parent = parent; n->label = label;
    n->o_total = n->o_blank = n->o_label = LOG0;
    n->n_total = n->n_blank = n->n_label = LOG0;
    n->pool_next = tr->all; tr->all = n;
    return n;
}
static Node* node_child(Tree* tr, Node* b, int k) {
    if (!b->children) b->children = (Node**)calloc((size_t)tr->C, sizeof(Node*));
    if (!b->children[k]) b->children[k] = node_new(tr, b, k);
    return b->children[k];
}
static void tree_free(Tree* tr) {
    Node* n = tr->all;
    while (n) { Node* nx = n->pool_next; free(n->children); free(n); n = nx; }
    tr->all = NULL;
}
static int node_before(const Node* a, const Node* b) { /* a ranks ahead of b */
    if (a->n_total != b->n_total) return a->n_total > b->n_total;
    return a->seq < b->seq;
}
/* bounded best-k list kept sorted (insertion); k <= a few hundred so O(k) insert is fine */
typedef struct { Node** v; int n, cap; } Leaves;
static Node* leaves_bottom(Leaves* l) { return l->v[l->n - 1]; }
static void leaves_push(Leaves* l, Node* e) {
    int i;
    if (l->n == l->cap) { /* caller verified e beats the bottom */ l->n--; }
    i = l->n++;
    while (i > 0 && node_before(e, l->v[i - 1])) { l->v[i] = l->v[i - 1]; --i; }
    l->v[i] = e;
}

/* experiment switch (tests only): 1 = skip the oldp reset of a rejected re-created child, i.e.
 * the "top-k of the union" model without TF's order-dependent side effect. */
static long g_stat[8]; /* steps, sum M, max M, resets, accepted pushes, sum nb */
void oracle_beam_stats(long* out, int clear) { int i; for (i = 0; i < 8; ++i) { out[i] = g_stat[i]; if (clear) g_stat[i] = 0; } }
static int g_beam_no_reset = 0;
void oracle_beam_set_no_reset(int v) { g_beam_no_reset = v; }
static inline float LSE2(int math, float a, float b) { return math ? det_lse2(a, b) : lse_libm(a, b); }

static void beam_one(const float* x, long stride, int T, int C, int beam_width, int top_paths,
                     int merge_repeated, int normalize, int math,
                     int64_t* decoded /*[top_paths,Tmax]*/, int Tmax, int* dec_len, float* log_prob)
{
    const int blank = C - 1;
    Tree tr; tr.all = NULL; tr.C = C;
    Leaves lv; lv.cap = beam_width; lv.n = 0; lv.v = (Node**)malloc(sizeof(Node*) * beam_width);
    Node** br = (Node**)malloc(sizeof(Node*) * beam_width);
    float* in = (float*)malloc(sizeof(float) * C);
    Node* root = node_new(&tr, NULL, -1);
    int t, i, k, p;
    root->n_total = 0.0f; root->n_blank = 0.0f;
    root->seq = 0;
    leaves_push(&lv, root);
    for (t = 0; t < T; ++t) {
        const float* r = x + t * stride;
        float m = r[0];
        for (k = 1; k < C; ++k) if (r[k] > m) m = r[k];
        float off = m;
        if (normalize) {
            float s = 0.0f;
            for (k = 0; k < C; ++k) s += math ? det_expf(r[k] - m) : expf(r[k] - m);
            off = m + (math ? det_logf(s) : logf(s));
        }
        for (k = 0; k < C; ++k) in[k] = r[k] - off;
        int nb = lv.n;
        for (i = 0; i < nb; ++i) br[i] = lv.v[i];
        lv.n = 0;
        for (i = 0; i < nb; ++i) { Node* b = br[i]; b->o_total = b->n_total; b->o_blank = b->n_blank; b->o_label = b->n_label; }
        long seq = 0;
        for (i = 0; i < nb; ++i) {
            Node* b = br[i];
            if (b->parent) {
                if (b->parent->n_total != LOG0) { /* parent->Active() */
                    float prev = (b->label == b->parent->label) ? b->parent->o_blank : b->parent->o_total;
                    b->n_label = LSE2(math, b->n_label, prev);
                }
                b->n_label += in[b->label];
            }
            b->n_blank = b->o_total + in[blank];
            b->n_total = LSE2(math, b->n_blank, b->n_label);
            b->seq = seq++;
            leaves_push(&lv, b);
        }
#ifdef ORACLE_STATS
        {
            float thmin = LOG0; long M = 0;
            if (nb == beam_width) { thmin = br[0]->n_total; for (i = 0; i < nb; ++i) if (br[i]->n_total < thmin) thmin = br[i]->n_total; }
            for (i = 0; i < nb; ++i) for (k = 0; k < C - 1; ++k) {
                float prev = (k == br[i]->label) ? br[i]->o_blank : br[i]->o_total;
                if (in[k] + prev > thmin) M++;
            }
            __sync_fetch_and_add(&g_stat[0], 1); __sync_fetch_and_add(&g_stat[1], M);
            if (M > g_stat[2]) g_stat[2] = M;
            __sync_fetch_and_add(&g_stat[5], nb);
        }
#endif
        for (i = 0; i < nb; ++i) {
            Node* b = br[i];
            /* is_candidate(b->oldp) */
            if (!(b->o_total > LOG0 && (lv.n < beam_width || b->o_total > leaves_bottom(&lv)->n_total))) {
                seq += C; /* keep the push numbering independent of this shortcut */
                continue;
            }
            for (k = 0; k < C; ++k, ++seq) {
                if (k == blank) continue;
                Node* c = node_child(&tr, b, k);
                if (c->n_total != LOG0) continue; /* already active */
                float prev = (k == b->label) ? b->o_blank : b->o_total;
                c->n_blank = LOG0;
                c->n_label = in[k] + prev;
                c->n_total = c->n_label;
                c->seq = seq;
                if (c->n_total > LOG0 && (lv.n < beam_width || c->n_total > leaves_bottom(&lv)->n_total)) {
                    if (lv.n == beam_width) {
                        Node* bot = leaves_bottom(&lv);
                        bot->n_total = bot->n_blank = bot->n_label = LOG0;
                    }
                    leaves_push(&lv, c);
                } else {
#ifdef ORACLE_STATS
                    if (c->o_total != LOG0) __sync_fetch_and_add(&g_stat[3], 1);
#endif
                    if (!g_beam_no_reset) c->o_total = c->o_blank = c->o_label = LOG0;
                    c->n_total = c->n_blank = c->n_label = LOG0;
                }
            }
        }
    }
    /* TopPaths */
    for (p = 0; p < top_paths; ++p) {
        int64_t* out = decoded + (size_t)p * Tmax;
        int n = 0;
        if (p < lv.n) {
            Node* c = lv.v[p];
            int prev = -1;
            log_prob[p] = c->n_total;
            while (c->parent) {
                if (!merge_repeated || c->label != prev) out[n++] = c->label;
                prev = c->label;
                c = c->parent;
            }
            for (i = 0; i < n / 2; ++i) { int64_t tmp = out[i]; out[i] = out[n - 1 - i]; out[n - 1 - i] = tmp; }
        } else {
            log_prob[p] = LOG0; /* TF errors: fewer leaves than requested paths */
        }
        dec_len[p] = n;
        for (i = n; i < Tmax; ++i) out[i] = -1;
    }
    tree_free(&tr);
    free(lv.v); free(br); free(in);
}

typedef struct {
    const float* logits; int T, B, C; const int* seq_len; int beam_width, top_paths, merge_repeated, normalize, math;
    int64_t* decoded; int* decoded_len; float* log_prob;
} beam_ctx;
static void beam_body(int b, void* vc) {
    beam_ctx* c = (beam_ctx*)vc;
    beam_one(c->logits + (size_t)b * c->C, (long)c->B * c->C, c->seq_len[b], c->C, c->beam_width, c->top_paths,
             c->merge_repeated, c->normalize, c->math, c->decoded + (size_t)b * c->top_paths * c->T, c->T,
             c->decoded_len + (size_t)b * c->top_paths, c->log_prob + (size_t)b * c->top_paths);
}
int oracle_ctc_beam(const float* logits, int T, int B, int C, const int* seq_len, int beam_width,
                    int top_paths, int merge_repeated, int normalize, int math,
                    int64_t* decoded /*[B,top_paths,T]*/, int* decoded_len /*[B,top_paths]*/,
                    float* log_prob /*[B,top_paths]*/, int nthreads)
{
    beam_ctx c = { logits, T, B, C, seq_len, beam_width, top_paths, merge_repeated, normalize, math,
                   decoded, decoded_len, log_prob };
    if (top_paths > beam_width) return -1;
    parallel_for(B, nthreads, beam_body, &c);
    return 0;
}

/* ------------------------------------------------------------------ edit distance
 * tf.edit_distance(hyp, truth, normalize=False): Levenshtein distance per example. */
int oracle_edit_distance(const int64_t* hyp, const int* hyp_off, const int64_t* truth, const int* truth_off,
                         int B, float* dist)
{
    int b;
    for (b = 0; b < B; ++b) {
        const int64_t* h = hyp + hyp_off[b]; int n = hyp_off[b + 1] - hyp_off[b];
        const int64_t* g = truth + truth_off[b]; int m = truth_off[b + 1] - truth_off[b];
        int* row = (int*)malloc(sizeof(int) * (m + 1));
        int i, j;
        for (j = 0; j <= m; ++j) row[j] = j;
        for (i = 1; i <= n; ++i) {
            int diag = row[0];
            row[0] = i;
            for (j = 1; j <= m; ++j) {
                int up = row[j];
                int best = diag + (h[i - 1] != g[j - 1]);
                if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
                if (up + 1 < best) best = up + 1;
                diag = up;
                row[j] = best;
            }
        }
        dist[b] = (float)row[m];
        free(row);
    }
    return 0;
}

/* exposed so tests can measure det_math accuracy against libm */
float oracle_det_expf(float x) { return det_expf(x); }
float oracle_det_logf(float x) { return det_logf(x); }
float oracle_det_lse2(float a, float b) { return det_lse2(a, b); }
int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
