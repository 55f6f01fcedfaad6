/*
 * tree_model.c -- CPU model of the DEVICE tree algorithm (ataxzero_b200/csrc/az_tree.cu, "compact" node layout).
 *
 * TEST INFRASTRUCTURE ONLY (same rule as ataxx_oracle.c: only tests/ may load this).
 *
 * The reference's select_action (cpp/self_play_client.cpp:333-366) scores EVERY legal move of a node on every
 * visit.  The device kernel does not: per node it keeps
 *   - a list of the moves that already have an edge ("visited", in order of creation) with their P / W / n, and
 *   - ONE candidate among the moves without an edge: the one with the largest prior (exactly equal priors: the
 *     one that comes last in the reference's unordered_map iteration order), plus the next smaller prior value.
 * A move without an edge scores fl(sqrt(1+N) * P) + 0 (:313-315,322), which is monotone in P, so only the candidate
 * can win among them -- except when two different priors round to the same product, which the kernel detects with
 * the second prior value and resolves by a full scan.  This file restates that algorithm sequentially so that
 * tests/test_tree_model.py can check it against the reference-order oracle (ao_mcts_*) on the CPU, including
 * exact ties (uniform evaluator), re-rooted trees (rank order changes, :155,489-490) and adversarially perturbed
 * priors (adjacent doubles, heavy quantisation).
 */
#include "ataxx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct tm_node tm_node;
typedef struct {
    double P, W;
    int n;
    tm_node *child;
    int idx;                       /* movegen index of the move */
} tm_entry;

struct tm_node {
    ao_position board;
    int terminal, populated;
    double value;
    int L, N, k;
    int cand;                      /* movegen index of the candidate, -1 = every move has an edge */
    double candP, cand2P;          /* cand2P < 0: no second prior value */
    int ranked, repopulated;
    int32_t from[AO_MAX_MOVES], to[AO_MAX_MOVES];
    double P[AO_MAX_MOVES];
    unsigned char visited[AO_MAX_MOVES];
    int rank[AO_MAX_MOVES];
    tm_entry ent[AO_MAX_MOVES];
};

typedef struct {
    tm_node *root;
    ao_eval_fn fn; void *ctx;
    long evals, slow_selects, rank_computations, near_ties;
    int force_slow;
} tm_tree;

/* test hook shared with the oracle (ataxx_oracle.c): perturbs the priors of a freshly populated node */
extern void (*ao_prior_hook)(double *prior, int n);

static int buckets_after(int n) { return n <= 13 ? 13 : n <= 29 ? 29 : n <= 59 ? 59 : n <= 127 ? 127 : 257; }

static tm_node *tm_node_new(const ao_position *b)
{
    tm_node *n = (tm_node *)calloc(1, sizeof(tm_node));
    n->board = *b;
    n->cand = -1;
    return n;
}

static void tm_node_free(tm_node *n)
{
    if (!n) return;
    for (int e = 0; e < n->k; e++) tm_node_free(n->ent[e].child);
    free(n);
}

static void ensure_ranked(tm_tree *t, tm_node *nd)
{
    if (nd->ranked) return;
    int32_t order[AO_MAX_MOVES];
    ao_umap_order(nd->from, nd->to, nd->L, nd->repopulated ? buckets_after(nd->L) : 0, order);
    for (int k = 0; k < nd->L; k++) nd->rank[order[k]] = k;
    nd->ranked = 1;
    t->rank_computations++;
}

/* candidate among the moves without an edge: max prior, exact ties -> last in map order; second prior value */
static void rescan(tm_tree *t, tm_node *nd)
{
    double pmax = -1.0, p2 = -1.0;
    int ties = 0, best = -1;
    for (int i = 0; i < nd->L; i++) {
        if (nd->visited[i]) continue;
        if (nd->P[i] > pmax) { pmax = nd->P[i]; ties = 1; best = i; }
        else if (nd->P[i] == pmax) ties++;
    }
    if (ties > 1) {
        ensure_ranked(t, nd);
        for (int i = 0; i < nd->L; i++)
            if (!nd->visited[i] && nd->P[i] == pmax && nd->rank[i] > nd->rank[best]) best = i;
    }
    for (int i = 0; i < nd->L; i++)
        if (!nd->visited[i] && nd->P[i] < pmax && nd->P[i] > p2) p2 = nd->P[i];
    nd->cand = best;
    nd->candP = best >= 0 ? pmax : 0.0;
    nd->cand2P = p2;
}

static void tm_populate(tm_tree *t, tm_node *n)
{
    if (n->populated) return;
    int result = ao_result(&n->board);
    if (result != 0) {
        n->terminal = 1;
        n->L = 0;
        n->value = result == 1 ? 1.0 : -1.0;
        if (n->board.turn == 1) n->value *= -1;
        n->populated = 1;
        return;
    }
    n->L = ao_movegen(&n->board, n->from, n->to);
    float feats[AO_FEATURES], logits[AO_LOGITS], v;
    ao_features(&n->board, feats);
    t->fn(t->ctx, feats, logits, &v);
    t->evals++;
    n->value = (double)v;
    ao_priors(logits, n->from, n->to, n->L, n->P);
    if (ao_prior_hook) ao_prior_hook(n->P, n->L);
    n->populated = 1;
    rescan(t, n);
}

/* full scan of the moves without an edge: max PRODUCT, ties -> last in map order */
static int slow_candidate(tm_tree *t, tm_node *nd, double s)
{
    double best = -1.0;
    int ties = 0, idx = -1;
    for (int i = 0; i < nd->L; i++) {
        if (nd->visited[i]) continue;
        double prod = s * nd->P[i];
        if (prod > best) { best = prod; ties = 1; idx = i; }
        else if (prod == best) ties++;
    }
    if (ties > 1) {
        ensure_ranked(t, nd);
        for (int i = 0; i < nd->L; i++)
            if (!nd->visited[i] && s * nd->P[i] == best && nd->rank[i] > nd->rank[idx]) idx = i;
    }
    t->slow_selects++;
    return idx;
}

/* returns: entry index e >= 0 (descend), -2 - idx (expand move idx), or -1 (NO_MOVE) */
static int tm_select(tm_tree *t, tm_node *nd)
{
    if (nd->terminal || nd->L == 0) return -1;
    const double s = sqrt((double)(1 + nd->N));
    double score[AO_MAX_MOVES + 1];
    int who[AO_MAX_MOVES + 1];       /* movegen index of contender */
    int cnt = 0;
    for (int e = 0; e < nd->k; e++) {
        const tm_entry *en = &nd->ent[e];
        double u = s / (double)(1 + en->n);
        double q = en->n == 0 ? 0.0 : en->W / (double)en->n;
        u *= en->P;
        score[cnt] = u + q;
        who[cnt] = en->idx;
        cnt++;
    }
    int cand_pos = -1;
    if (nd->cand >= 0) {
        int c = nd->cand;
        double sc = s * nd->candP;
        int near = nd->cand2P >= 0.0 && s * nd->cand2P == sc;
        if (near) t->near_ties++;
        if (near || t->force_slow) c = slow_candidate(t, nd, s);
        cand_pos = cnt;
        score[cnt] = s * nd->P[c] + 0.0;
        who[cnt] = c;
        cnt++;
    }
    double m = -1.0;
    int at = 0, win = -1;
    for (int j = 0; j < cnt; j++) {
        if (score[j] > m) { m = score[j]; at = 1; win = j; }
        else if (score[j] == m) at++;
    }
    if (at > 1) {
        ensure_ranked(t, nd);
        for (int j = 0; j < cnt; j++)
            if (score[j] == m && nd->rank[who[j]] > nd->rank[who[win]]) win = j;
    }
    if (win == cand_pos) return -2 - who[win];
    return win;
}

void *tm_new(const ao_position *root, ao_eval_fn fn, void *ctx, int force_slow)
{
    tm_tree *t = (tm_tree *)calloc(1, sizeof(tm_tree));
    t->fn = fn; t->ctx = ctx; t->force_slow = force_slow;
    t->root = tm_node_new(root);
    tm_populate(t, t->root);
    return t;
}

void tm_free(void *tv)
{
    tm_tree *t = (tm_tree *)tv;
    if (!t) return;
    tm_node_free(t->root);
    free(t);
}

void tm_step(void *tv)
{
    tm_tree *t = (tm_tree *)tv;
    enum { MAX_PATH = 4096 };
    tm_node *pn[MAX_PATH]; int pe[MAX_PATH]; int depth = 0;
    tm_node *node = t->root, *leaf;
    int sel;
    for (;;) {
        sel = tm_select(t, node);
        if (sel < 0) break;
        pn[depth] = node; pe[depth] = sel; depth++;
        node = node->ent[sel].child;
    }
    leaf = node;
    if (sel <= -2) {
        const int idx = -2 - sel;
        ao_position nb = node->board;
        ao_makemove(&nb, node->from[idx], node->to[idx]);
        leaf = tm_node_new(&nb);
        tm_entry *en = &node->ent[node->k];
        en->P = node->P[idx]; en->W = 0.0; en->n = 0; en->child = leaf; en->idx = idx;
        node->visited[idx] = 1;
        pn[depth] = node; pe[depth] = node->k; depth++;
        node->k++;
        rescan(t, node);
    }
    tm_populate(t, leaf);
    double score = (leaf->value + 1.0) / 2.0;
    for (int d = depth - 1; d >= 0; d--) {
        score = 1.0 - score;
        pn[d]->ent[pe[d]].n += 1;
        pn[d]->ent[pe[d]].W += score;
        pn[d]->N++;
    }
}

int tm_root_visits(const void *tv) { return ((const tm_tree *)tv)->root->N; }
long tm_evals(const void *tv) { return ((const tm_tree *)tv)->evals; }
long tm_counter(const void *tv, int which)
{
    const tm_tree *t = (const tm_tree *)tv;
    return which == 0 ? t->slow_selects : which == 1 ? t->rank_computations : t->near_ties;
}

int tm_root_dist(const void *tv, int32_t *visits, double *total_score, double *prior)
{
    const tm_node *r = ((const tm_tree *)tv)->root;
    for (int i = 0; i < r->L; i++) {
        visits[i] = 0; total_score[i] = 0.0; prior[i] = r->P[i];
    }
    for (int e = 0; e < r->k; e++) {
        visits[r->ent[e].idx] = r->ent[e].n;
        total_score[r->ent[e].idx] = r->ent[e].W;
    }
    return r->L;
}

int tm_play(void *tv, int from, int to)
{
    tm_tree *t = (tm_tree *)tv;
    tm_node *r = t->root;
    int idx = -1, ent = -1;
    for (int i = 0; i < r->L; i++)
        if (r->from[i] == from && r->to[i] == to) idx = i;
    for (int e = 0; e < r->k; e++)
        if (r->ent[e].idx == idx) ent = e;
    if (idx < 0 || ent < 0) {
        ao_position nb = r->board;
        ao_makemove(&nb, from, to);
        tm_node_free(r);
        t->root = tm_node_new(&nb);
        tm_populate(t, t->root);
        return 0;
    }
    tm_node *keep = r->ent[ent].child;
    r->ent[ent].child = NULL;
    tm_node_free(r);
    t->root = keep;
    /* the reference re-populates the new root: same priors (deterministic evaluator), but the posterior map is
     * clear()ed and refilled, which changes its iteration order -> ranks are stale, the candidate may change */
    if (!keep->terminal) {
        float feats[AO_FEATURES], logits[AO_LOGITS], v;     /* the evaluation the reference repeats here */
        ao_features(&keep->board, feats);
        t->fn(t->ctx, feats, logits, &v);
        t->evals++;
        keep->repopulated = 1;
        keep->ranked = 0;
        rescan(t, keep);
    }
    return 1;
}

/* Runs the model and the reference-order oracle side by side: `visits` root visits, then `plays` times
 * (play the most visited root move -- first in movegen order on equal counts -- and search again).
 * Returns the number of root edges whose visit count or total score (bitwise) differ, summed over all
 * comparisons; counters_out[0..3] = slow selects, rank computations, near ties, evaluations of the model. */
long tm_selftest(const char *fen, int visits, int evaluator, int plays, int force_slow, long *counters_out)
{
    ao_position p;
    if (ao_set_board(&p, fen) != 0) return -1;
    ao_eval_fn fn = evaluator == 1 ? ao_uniform_eval : ao_probe_eval;
    ao_mcts *ref = ao_mcts_new(&p, fn, NULL);
    tm_tree *t = (tm_tree *)tm_new(&p, fn, NULL, force_slow);
    long bad = 0;
    for (int round = 0; round <= plays; round++) {
        while (ao_mcts_root_visits(ref) < visits) {
            const int before = ao_mcts_root_visits(ref);
            ao_mcts_step(ref);
            if (ao_mcts_root_visits(ref) == before) break;       /* terminal root */
        }
        while (tm_root_visits(t) < visits) {
            const int before = tm_root_visits(t);
            tm_step(t);
            if (tm_root_visits(t) == before) break;
        }
        int32_t from[AO_MAX_MOVES], to[AO_MAX_MOVES], v1[AO_MAX_MOVES], v2[AO_MAX_MOVES];
        double w1[AO_MAX_MOVES], w2[AO_MAX_MOVES], p1[AO_MAX_MOVES], p2[AO_MAX_MOVES];
        const int n1 = ao_mcts_root_dist(ref, from, to, v1, w1, p1);
        const int n2 = tm_root_dist(t, v2, w2, p2);
        if (n1 != n2) { bad += 1000; break; }
        int best = -1;
        for (int i = 0; i < n1; i++) {
            if (v1[i] != v2[i] || memcmp(&w1[i], &w2[i], sizeof(double)) != 0 || memcmp(&p1[i], &p2[i], sizeof(double)) != 0) bad++;
            if (v1[i] > 0 && (best < 0 || v1[i] > v1[best])) best = i;
        }
        if (ao_mcts_eval_count(ref) != tm_evals(t)) bad++;
        if (round == plays || best < 0) break;
        ao_mcts_play(ref, from[best], to[best]);
        tm_play(t, from[best], to[best]);
    }
    if (counters_out) {
        counters_out[0] = t->slow_selects; counters_out[1] = t->rank_computations;
        counters_out[2] = t->near_ties; counters_out[3] = t->evals;
    }
    ao_mcts_free(ref);
    tm_free(t);
    return bad;
}

/* ------------------------------------------------------------------------------------------------
 * Tick simulation of SPECULATIVE leaf evaluation for single-tree search (BASELINE config 5): how many net round trips
 * ("ticks") does a bit-exact sequential search need when, each time a node's evaluation is consumed, the evaluations of
 * its `top_k` highest-prior children are requested in the same batch (engine.py:387-392 queues children with prior > 0.15
 * the same way)?  A child that was requested in an EARLIER tick links without waiting; everything else costs a tick.
 * Evaluator 2 mimics the random-init net of the benchmark: nearly uniform priors, values within +-0.005.
 * Returns the number of ticks; *evals_out = evaluations requested (speculative ones included).
 * ------------------------------------------------------------------------------------------------ */
static void flat_eval(void *ctx, const float feats[AO_FEATURES], float logits[AO_LOGITS], float *value)
{
    ao_probe_eval(ctx, feats, logits, value);
    for (int i = 0; i < AO_LOGITS; i++) logits[i] *= 0.025f;       /* logit std ~0.03 like model-001 */
    *value *= 0.006f;
}

long tm_spec_sim(const char *fen, int visits, int evaluator, int top_k, long *evals_out)
{
    ao_position p;
    if (ao_set_board(&p, fen) != 0) return -1;
    ao_eval_fn fn = evaluator == 1 ? ao_uniform_eval : evaluator == 2 ? flat_eval : ao_probe_eval;
    tm_tree *t = (tm_tree *)tm_new(&p, fn, NULL, 0);
    long now = 1, requested = 1;       /* the root evaluation: requested in tick 0, ready in tick 1 */
    /* per node: the tick in which its evaluation was CONSUMED (side table in order of creation; recent nodes are found first) */
    typedef struct { tm_node *n; long consumed; } rec;
    rec *recs = (rec *)calloc((size_t)visits + 8, sizeof(rec));
    long n_recs = 0;
    recs[n_recs].n = t->root; recs[n_recs].consumed = 1; n_recs++;
    while (t->root->N < visits) {
        /* one step, with timing: find the parent of the node this step will expand */
        tm_node *node = t->root;
        int sel;
        for (;;) {
            sel = tm_select(t, node);
            if (sel < 0 || sel <= -2) break;
            node = node->ent[sel].child;
        }
        if (sel <= -2) {
            const int idx = -2 - sel;
            /* parent's consumption tick */
            long consumed = 0;
            for (long i = n_recs - 1; i >= 0; i--) if (recs[i].n == node) { consumed = recs[i].consumed; break; }
            /* was this child among the parent's top_k priors? */
            int better = 0;
            for (int i = 0; i < node->L; i++) if (node->P[i] > node->P[idx]) better++;
            long ready;
            if (better < top_k) ready = consumed + 1;            /* requested when the parent was consumed */
            else { ready = now + 1; requested++; }               /* requested only now */
            if (ready > now) now = ready;
            tm_step(t);
            tm_node *leaf = node->ent[node->k - 1].child;
            if (!leaf->terminal) {
                recs[n_recs].n = leaf; recs[n_recs].consumed = now; n_recs++;
                int spec = leaf->L < top_k ? leaf->L : top_k;
                requested += spec;
            }
        } else {
            tm_step(t);                                          /* terminal re-visit: no evaluation */
        }
    }
    if (evals_out) *evals_out = requested;
    free(recs);
    tm_free(t);
    return now;
}

/* ------------------------------------------------------------------------------------------------
 * Design study: how often does a simulation walk THROUGH the node the previous simulation created?  (A pool that backs a
 * leaf's value up at once but computes its priors off the critical path would have to hold such a simulation back for a
 * tick.)  `plays` > 0: after `visits` root visits play the most visited move and search on, like self-play with tree reuse.
 * Returns simulations counted; *through_out = those that passed through the previous simulation's new node.
 * ------------------------------------------------------------------------------------------------ */
long tm_revisit_sim(const char *fen, int visits, int evaluator, int plays, long *through_out)
{
    ao_position p;
    if (ao_set_board(&p, fen) != 0) return -1;
    ao_eval_fn fn = evaluator == 1 ? ao_uniform_eval : evaluator == 2 ? flat_eval : ao_probe_eval;
    tm_tree *t = (tm_tree *)tm_new(&p, fn, NULL, 0);
    long sims = 0, through = 0;
    tm_node *last = NULL;
    for (int round = 0; round <= plays; round++) {
        while (t->root->N < visits) {
            tm_node *node = t->root;
            int sel, hit = 0;
            for (;;) {
                if (node == last) hit = 1;
                sel = tm_select(t, node);
                if (sel < 0 || sel <= -2) break;
                node = node->ent[sel].child;
            }
            const int before = t->root->N;
            tm_node *parent = node;
            const int k_before = parent->k;
            tm_step(t);
            if (t->root->N == before) goto done;                 /* terminal root */
            sims++;
            through += hit;
            last = (sel <= -2 && parent->k > k_before && !parent->ent[parent->k - 1].child->terminal) ? parent->ent[parent->k - 1].child : NULL;
        }
        if (round == plays) break;
        int best = -1;
        for (int e = 0; e < t->root->k; e++)
            if (best < 0 || t->root->ent[e].n > t->root->ent[best].n) best = e;
        if (best < 0) break;
        const int idx = t->root->ent[best].idx;
        last = NULL;
        if (tm_play(t, t->root->from[idx], t->root->to[idx]) != 0) break;
    }
done:
    if (through_out) *through_out = through;
    tm_free(t);
    return sims;
}
