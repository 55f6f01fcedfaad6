/*
 * ataxx_oracle.c -- CPU restatement of the AtaxxZero self-play hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ataxx_oracle.h).  Written from the behaviour of
 * the reference, not from its text: tables are derived from board geometry, the
 * tree is an array-of-nodes restatement of the reference's hash-map tree, and
 * the libstdc++ hash-map iteration order the reference's tie-breaking silently
 * depends on is modelled explicitly (ao_umap_order).
 *
 * Pinned against: SURVEY App. C goldens and the compiled reference in
 * oracle/_ref (tests/test_oracle_pinned.py).
 */
#include "ataxx_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define BOARD_MASK 0x1FFFFFFFFFFFFULL   /* 49 bits, cpp/bitboards.hpp:7 */

/* ------------------------------------------------------------------ */
/* geometry-derived tables                                            */
/* ------------------------------------------------------------------ */

static uint64_t g_ring1[49], g_ring2[49];
static uint64_t g_file_mask[7];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void build_tables(void)
{
    for (int f = 0; f < 7; f++) {
        uint64_t m = 0;
        for (int r = 0; r < 7; r++) m |= 1ULL << (r * 7 + f);
        g_file_mask[f] = m;
    }
    for (int sq = 0; sq < 49; sq++) {
        int f = sq % 7, r = sq / 7;
        uint64_t a = 0, b = 0;
        for (int dr = -2; dr <= 2; dr++)
            for (int df = -2; df <= 2; df++) {
                int nf = f + df, nr = r + dr;
                if ((df == 0 && dr == 0) || nf < 0 || nf > 6 || nr < 0 || nr > 6) continue;
                int cheb = abs(df) > abs(dr) ? abs(df) : abs(dr);
                if (cheb == 1) a |= 1ULL << (nr * 7 + nf);
                else           b |= 1ULL << (nr * 7 + nf);
            }
        g_ring1[sq] = a;
        g_ring2[sq] = b;
    }
}

static inline void tables(void) { pthread_once(&g_once, build_tables); }

uint64_t ao_single_ring(int sq) { tables(); return g_ring1[sq]; }
uint64_t ao_double_ring(int sq) { tables(); return g_ring2[sq]; }

/* shift a bitboard by (df, dr), dropping bits that would wrap across a file edge */
static uint64_t shift_bb(uint64_t bb, int df, int dr)
{
    tables();
    uint64_t keep = BOARD_MASK;
    /* a piece can move df files only if it stays on the board: mask the SOURCE files */
    for (int f = 0; f < 7; f++)
        if (f + df < 0 || f + df > 6) keep &= ~g_file_mask[f];
    bb &= keep;
    int s = dr * 7 + df;
    bb = s >= 0 ? bb << s : bb >> (-s);
    return bb & BOARD_MASK;
}

uint64_t ao_single_jump_bb(uint64_t bb)
{
    uint64_t out = 0;
    for (int dr = -1; dr <= 1; dr++)
        for (int df = -1; df <= 1; df++)
            if (df || dr) out |= shift_bb(bb, df, dr);
    return out;
}

uint64_t ao_double_jump_bb(uint64_t bb)
{
    uint64_t out = 0;
    for (int dr = -2; dr <= 2; dr++)
        for (int df = -2; df <= 2; df++)
            if (abs(df) == 2 || abs(dr) == 2) out |= shift_bb(bb, df, dr);
    return out;
}

/* ------------------------------------------------------------------ */
/* position                                                           */
/* ------------------------------------------------------------------ */

int ao_invalid(const ao_position *p)
{
    if (p->pieces[1] & p->pieces[0]) return 1;
    if (p->pieces[1] & p->blockers)  return 2;
    if (p->pieces[0] & p->blockers)  return 3;
    if (p->pieces[1] & ~BOARD_MASK)  return 4;
    if (p->pieces[0] & ~BOARD_MASK)  return 5;
    if (p->blockers  & ~BOARD_MASK)  return 6;
    return 0;
}

/* FEN: rows rank 7 -> 1, 'x'/'o' pieces, '-' blocker, digits = empties, then side
 * to move.  Error codes follow cpp/ataxx.cpp:14-92. */
int ao_set_board(ao_position *pos, const char *fen)
{
    static const char *START = "x5o/7/3-3/2-1-2/3-3/7/o5x x";
    if (strcmp(fen, "startpos") == 0) fen = START;

    /* split on single spaces the way the reference's getline-split does */
    const char *tok[4]; int len[4]; int nt = 0;
    const char *s = fen;
    if (*s == 0) return 1;
    for (;;) {
        const char *e = strchr(s, ' ');
        if (nt < 4) { tok[nt] = s; len[nt] = e ? (int)(e - s) : (int)strlen(s); }
        nt++;
        if (!e) break;
        s = e + 1;
        if (*s == 0) break;           /* trailing delimiter yields no extra token */
    }
    if (nt == 0) return 1;
    if (nt > 2) return 2;
    if (len[0] < 13) return 3;
    if (len[0] > 55) return 4;

    pos->pieces[0] = pos->pieces[1] = pos->blockers = 0;
    pos->turn = 0;
    pos->ply = 0;
    int sq = 42;
    for (int i = 0; i < len[0]; i++) {
        char c = tok[0][i];
        if (c == 'x' || c == 'X')      { pos->pieces[0] ^= 1ULL << sq; sq++; }
        else if (c == 'o' || c == 'O') { pos->pieces[1] ^= 1ULL << sq; sq++; }
        else if (c == '-')             { pos->blockers  ^= 1ULL << sq; sq++; }
        else if (c >= '1' && c <= '7') sq += c - '0';
        else if (c == '/')             sq -= 14;
        else return 5;
    }
    if (nt > 1) {
        if (len[1] == 1 && (tok[1][0] == 'x' || tok[1][0] == 'X')) pos->turn = 0;
        else if (len[1] == 1 && (tok[1][0] == 'o' || tok[1][0] == 'O')) pos->turn = 1;
        else return 6;
    }
    if (sq != 7) return 7;
    if (ao_invalid(pos)) return 8;
    return 0;
}

int ao_movegen(const ao_position *pos, int32_t *from, int32_t *to)
{
    tables();
    const uint64_t own = pos->pieces[pos->turn];
    const uint64_t empty = BOARD_MASK & ~(pos->pieces[0] | pos->pieces[1] | pos->blockers);
    int n = 0;
    /* jumps first: source ascending, destination ascending */
    for (uint64_t src = own; src; src &= src - 1) {
        int f = __builtin_ctzll(src);
        for (uint64_t dst = g_ring2[f] & empty; dst; dst &= dst - 1) {
            from[n] = f; to[n] = __builtin_ctzll(dst); n++;
        }
    }
    /* then one clone per reachable destination, ascending; from == to */
    for (uint64_t dst = ao_single_jump_bb(own) & empty; dst; dst &= dst - 1) {
        int t = __builtin_ctzll(dst);
        from[n] = t; to[n] = t; n++;
    }
    return n;
}

void ao_makemove(ao_position *pos, int from, int to)
{
    tables();
    const int me = pos->turn, you = !pos->turn;
    const uint64_t flipped = g_ring1[to] & pos->pieces[you];
    pos->pieces[me] &= ~(1ULL << from);     /* no-op for a clone: from == to is empty */
    pos->pieces[me] ^= 1ULL << to;
    pos->pieces[me] ^= flipped;
    pos->pieces[you] ^= flipped;
    pos->turn = you;
    pos->ply++;
}

int ao_legal_move(const ao_position *pos, int from, int to)
{
    tables();
    const uint64_t empty = BOARD_MASK & ~(pos->pieces[0] | pos->pieces[1] | pos->blockers);
    if (!((1ULL << to) & empty)) return 0;
    if (from == to) return (g_ring1[to] & pos->pieces[pos->turn]) != 0;
    if (!((1ULL << from) & pos->pieces[pos->turn])) return 0;
    return ((1ULL << to) & g_ring2[from]) != 0;
}

int ao_move_string(int from, int to, char out[5])
{
    int n = 0;
    if (from != to) { out[n++] = (char)('a' + from % 7); out[n++] = (char)('1' + from / 7); }
    out[n++] = (char)('a' + to % 7);
    out[n++] = (char)('1' + to / 7);
    out[n] = 0;
    return n;
}

int ao_result(const ao_position *pos)
{
    int p1 = __builtin_popcountll(pos->pieces[0]);
    int p2 = __builtin_popcountll(pos->pieces[1]);
    int bl = __builtin_popcountll(pos->blockers);
    int empty = 49 - p1 - p2 - bl;
    if (p1 == 0) return 2;
    if (p2 == 0) return 1;
    int32_t f[AO_MAX_MOVES], t[AO_MAX_MOVES];
    if (ao_movegen(pos, f, t) == 0) {      /* stuck: opponent is credited every empty cell */
        if (pos->turn == 0) p2 += empty; else p1 += empty;
    }
    if (p1 + p2 + bl == 49) return p1 < p2 ? 2 : 1;
    return 0;
}

void ao_features(const ao_position *pos, float out[AO_FEATURES])
{
    memset(out, 0, sizeof(float) * AO_FEATURES);
    for (int x = 0; x < 7; x++)
        for (int y = 0; y < 7; y++) {
            float *cell = out + 28 * x + 4 * y;
            uint64_t bit = 1ULL << (x + 7 * (6 - y));
            cell[0] = 1.0f;
            if (pos->pieces[pos->turn] & bit)  cell[1] = 1.0f;
            if (pos->pieces[!pos->turn] & bit) cell[2] = 1.0f;
            if (pos->blockers & bit)           cell[3] = 1.0f;
        }
}

void ao_board_json(const ao_position *pos, int32_t out[49])
{
    for (int y = 0; y < 7; y++)
        for (int x = 0; x < 7; x++) {
            uint64_t bit = 1ULL << (x + 7 * (6 - y));
            out[y * 7 + x] = (pos->pieces[0] & bit) ? 1 : (pos->pieces[1] & bit) ? 2 : 0;
        }
}

int ao_policy_index(int from, int to)
{
    int tx = to % 7, ty = 6 - to / 7;
    if (from == to) return 119 * tx + 17 * ty + 16;
    int dx = tx - from % 7, dy = ty - (6 - from / 7);
    /* plane order: all (dx,dy) with max(|dx|,|dy|)==2, dx-major then dy */
    int plane = 0;
    for (int a = -2; a <= 2; a++)
        for (int b = -2; b <= 2; b++) {
            if (abs(a) != 2 && abs(b) != 2) continue;
            if (a == dx && b == dy) return 119 * tx + 17 * ty + plane;
            plane++;
        }
    return -1;
}

/* ------------------------------------------------------------------ */
/* perft                                                              */
/* ------------------------------------------------------------------ */

uint64_t ao_perft(const ao_position *pos, int depth)
{
    if (depth <= 0) return 1;
    int32_t f[AO_MAX_MOVES], t[AO_MAX_MOVES];
    int n = ao_movegen(pos, f, t);
    if (depth == 1) return (uint64_t)n;
    uint64_t total = 0;
    for (int i = 0; i < n; i++) {
        ao_position c = *pos;
        ao_makemove(&c, f[i], t[i]);
        total += ao_perft(&c, depth - 1);
    }
    return total;
}

typedef struct {
    const ao_position *items; int n_items; int depth;
    int next; uint64_t total; pthread_mutex_t mu;
} perft_job;

static void *perft_worker(void *arg)
{
    perft_job *job = (perft_job *)arg;
    uint64_t local = 0;
    for (;;) {
        pthread_mutex_lock(&job->mu);
        int i = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (i >= job->n_items) break;
        local += ao_perft(&job->items[i], job->depth);
    }
    pthread_mutex_lock(&job->mu);
    job->total += local;
    pthread_mutex_unlock(&job->mu);
    return NULL;
}

uint64_t ao_perft_mt(const ao_position *pos, int depth, int threads)
{
    tables();
    if (depth < 4 || threads <= 1) return ao_perft(pos, depth);
    /* depth-2 frontier as work items */
    int cap = AO_MAX_MOVES * AO_MAX_MOVES, n_items = 0;
    ao_position *items = (ao_position *)malloc(sizeof(ao_position) * (size_t)cap);
    int32_t f1[AO_MAX_MOVES], t1[AO_MAX_MOVES], f2[AO_MAX_MOVES], t2[AO_MAX_MOVES];
    int n1 = ao_movegen(pos, f1, t1);
    for (int i = 0; i < n1; i++) {
        ao_position a = *pos; ao_makemove(&a, f1[i], t1[i]);
        int n2 = ao_movegen(&a, f2, t2);
        for (int j = 0; j < n2; j++) {
            ao_position b = a; ao_makemove(&b, f2[j], t2[j]);
            items[n_items++] = b;
        }
    }
    perft_job job = { items, n_items, depth - 2, 0, 0, PTHREAD_MUTEX_INITIALIZER };
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, perft_worker, &job);
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    free(items);
    return job.total;
}

/* ------------------------------------------------------------------ */
/* priors                                                             */
/* ------------------------------------------------------------------ */

void ao_priors(const float logits[AO_LOGITS], const int32_t *from, const int32_t *to,
               int n_moves, double *prior_out)
{
    double soft[AO_LOGITS];
    for (int i = 0; i < AO_LOGITS; i++) soft[i] = exp((double)logits[i]);  /* no max-subtraction */
    double total = 0.0;
    for (int i = 0; i < AO_LOGITS; i++) total += soft[i];                  /* sequential */
    if (total != 0.0)
        for (int i = 0; i < AO_LOGITS; i++) soft[i] /= total;
    double legal = 0.0;
    for (int i = 0; i < n_moves; i++) {
        prior_out[i] = soft[ao_policy_index(from[i], to[i])];
        legal += prior_out[i];                                             /* movegen order */
    }
    if (legal != 0.0)
        for (int i = 0; i < n_moves; i++) prior_out[i] /= legal;
}

/* ------------------------------------------------------------------ */
/* libstdc++ unordered_map iteration-order model                      */
/* ------------------------------------------------------------------ */

/* Bucket growth of libstdc++'s _Prime_rehash_policy for max_load_factor 1 starting
 * from the single-bucket empty table: 1 -> 13 -> 29 -> 59 -> 127 -> 257 (each is the
 * first entry of its prime table >= max(n+1, 2*buckets)).  Enough for < 256 keys;
 * validated against the real container in tests/test_oracle_pinned.py. */
static int grow_buckets(int buckets, int want)
{
    static const int ladder[] = { 13, 29, 59, 127, 257, 541, 1109 };
    int need = want > 2 * buckets ? want : 2 * buckets;
    for (unsigned i = 0; i < sizeof(ladder) / sizeof(ladder[0]); i++)
        if (ladder[i] >= need) return ladder[i];
    return -1;
}

int ao_umap_order(const int32_t *from, const int32_t *to, int n, int start_buckets,
                  int32_t *order_out)
{
    /* singly linked list with a sentinel "before begin" (index n); bucket b stores the
     * index of the node BEFORE its first node, or -1 when empty. */
    enum { CAP = 1200 };
    int nxt[AO_MAX_MOVES + 1];
    int bucket[CAP];
    const int SENT = n;
    int buckets = start_buckets > 0 ? start_buckets : 1;
    int next_resize = start_buckets > 0 ? start_buckets : 0;
    int count = 0;
    if (n > AO_MAX_MOVES) return -1;
    for (int b = 0; b < buckets; b++) bucket[b] = -1;
    nxt[SENT] = -1;
#define HASH(i) ((unsigned)(from[i] + 49 * to[i]))

    for (int i = 0; i < n; i++) {
        if (count + 1 > next_resize) {
            int min_bkts = count + 1;
            if (next_resize == 0 && min_bkts < 11) min_bkts = 11;
            if (min_bkts >= buckets) {
                int nb = grow_buckets(buckets, min_bkts + 1);
                /* rehash: walk the old list in order, re-thread into the new buckets */
                int p = nxt[SENT], begin_bkt = 0;
                for (int b = 0; b < nb; b++) bucket[b] = -1;
                nxt[SENT] = -1;
                while (p >= 0) {
                    int following = nxt[p];
                    int b = (int)(HASH(p) % (unsigned)nb);
                    if (bucket[b] < 0) {
                        nxt[p] = nxt[SENT];
                        nxt[SENT] = p;
                        bucket[b] = SENT;
                        if (nxt[p] >= 0) bucket[begin_bkt] = p;
                        begin_bkt = b;
                    } else {
                        nxt[p] = nxt[bucket[b]];
                        nxt[bucket[b]] = p;
                    }
                    p = following;
                }
                buckets = nb;
                next_resize = nb;
            } else {
                next_resize = buckets;
            }
        }
        int b = (int)(HASH(i) % (unsigned)buckets);
        if (bucket[b] >= 0) {                 /* non-empty: becomes first of its bucket */
            nxt[i] = nxt[bucket[b]];
            nxt[bucket[b]] = i;
        } else {                              /* empty: becomes head of the whole list */
            nxt[i] = nxt[SENT];
            nxt[SENT] = i;
            if (nxt[i] >= 0) bucket[HASH(nxt[i]) % (unsigned)buckets] = i;
            bucket[b] = SENT;
        }
        count++;
    }
#undef HASH
    int k = 0;
    for (int p = nxt[SENT]; p >= 0; p = nxt[p]) order_out[k++] = p;
    return buckets;
}

/* ------------------------------------------------------------------ */
/* MCTS                                                               */
/* ------------------------------------------------------------------ */

/* Test hook (NULL in normal use): perturbs the priors of every freshly populated node, in this oracle and in the
 * device-algorithm model (tree_model.c) alike, to provoke the tie situations real evaluators almost never produce. */
void (*ao_prior_hook)(double *prior, int n) = NULL;

static void perturb_quantise(double *p, int n)       /* many exactly equal priors among unequal logits */
{
    for (int i = 0; i < n; i++) p[i] = floor(p[i] * 64.0 + 0.5) / 64.0 + 1.0 / 1024.0;
}
static void perturb_adjacent(double *p, int n)       /* pairs of ADJACENT doubles: different priors, often equal products */
{
    for (int i = 0; i + 1 < n; i += 2) p[i + 1] = nextafter(p[i], 0.0);
}
static void perturb_adjacent_groups(double *p, int n) /* runs of 4 consecutive doubles sharing one leading value */
{
    for (int i = 0; i < n; i++)
        if (i % 4) p[i] = nextafter(p[i - 1], 0.0);
}
void ao_set_prior_perturbation(int mode)
{
    ao_prior_hook = mode == 1 ? perturb_quantise : mode == 2 ? perturb_adjacent : mode == 3 ? perturb_adjacent_groups : NULL;
}

typedef struct ao_node {
    ao_position board;
    int populated, game_over;
    double value;
    int n_moves;                 /* posterior.size(); 0 for terminal nodes */
    int map_buckets;             /* bucket count of the posterior map (for re-population) */
    int32_t from[AO_MAX_MOVES], to[AO_MAX_MOVES];
    int32_t order[AO_MAX_MOVES]; /* posterior iteration order */
    double prior[AO_MAX_MOVES];
    int all_edge_visits;
    struct ao_node *child[AO_MAX_MOVES];
    double edge_visits[AO_MAX_MOVES];
    double edge_total[AO_MAX_MOVES];
} ao_node;

struct ao_mcts {
    ao_node *root;
    ao_eval_fn fn; void *ctx;
    long evals, ties;
};

static ao_node *node_new(const ao_position *b)
{
    ao_node *n = (ao_node *)calloc(1, sizeof(ao_node));
    n->board = *b;
    return n;
}

static void node_free(ao_node *n)
{
    if (!n) return;
    for (int i = 0; i < n->n_moves; i++) node_free(n->child[i]);
    free(n);
}

static void node_populate(ao_mcts *m, ao_node *n)
{
    if (n->populated) return;
    n->game_over = 0;
    int result = ao_result(&n->board);
    if (result != 0) {
        n->game_over = 1;
        n->n_moves = 0;
        n->value = result == 1 ? 1.0 : -1.0;
        if (n->board.turn == 1) n->value *= -1;
        n->populated = 1;
        return;
    }
    n->n_moves = ao_movegen(&n->board, n->from, n->to);
    float feats[AO_FEATURES], logits[AO_LOGITS], v;
    ao_features(&n->board, feats);
    m->fn(m->ctx, feats, logits, &v);
    m->evals++;
    n->value = (double)v;
    ao_priors(logits, n->from, n->to, n->n_moves, n->prior);
    if (ao_prior_hook) ao_prior_hook(n->prior, n->n_moves);
    n->map_buckets = ao_umap_order(n->from, n->to, n->n_moves, n->map_buckets, n->order);
    n->populated = 1;
}

/* returns the move index or -1 for the reference's NO_MOVE */
static int node_select(ao_mcts *m, ao_node *n)
{
    if (n->n_moves == 0 || n->game_over) return -1;
    int best = -1, at_best = 0;
    double best_score = -1;
    for (int k = 0; k < n->n_moves; k++) {
        int i = n->order[k];
        double u, q;
        if (!n->child[i]) {
            u = sqrt(1 + n->all_edge_visits);
            q = 0;
        } else {
            u = sqrt(1 + n->all_edge_visits) / (1 + n->edge_visits[i]);
            q = n->edge_visits[i] == 0 ? 0 : n->edge_total[i] / n->edge_visits[i];
        }
        u *= 1.0 * n->prior[i];
        double score = u + q;
        if (score > best_score) at_best = 1; else if (score == best_score) at_best++;
        if (score >= best_score) { best = i; best_score = score; }   /* last maximal wins */
    }
    if (at_best > 1) m->ties++;
    return best;
}

ao_mcts *ao_mcts_new(const ao_position *root, ao_eval_fn fn, void *ctx)
{
    tables();
    ao_mcts *m = (ao_mcts *)calloc(1, sizeof(ao_mcts));
    m->fn = fn; m->ctx = ctx;
    m->root = node_new(root);
    node_populate(m, m->root);
    return m;
}

void ao_mcts_free(ao_mcts *m)
{
    if (!m) return;
    node_free(m->root);
    free(m);
}

void ao_mcts_step(ao_mcts *m)
{
    enum { MAX_PATH = 4096 };
    ao_node *pn[MAX_PATH]; int pi[MAX_PATH]; int depth = 0;
    ao_node *node = m->root;
    int mv;
    for (;;) {
        mv = node_select(m, node);
        if (mv < 0 || !node->child[mv]) break;
        pn[depth] = node; pi[depth] = mv; depth++;
        node = node->child[mv];
    }
    ao_node *leaf = node;
    if (mv >= 0) {
        ao_position nb = node->board;
        ao_makemove(&nb, node->from[mv], node->to[mv]);
        leaf = node_new(&nb);
        node->child[mv] = leaf;
        pn[depth] = node; pi[depth] = mv; depth++;
    }
    node_populate(m, leaf);
    double score = (leaf->value + 1.0) / 2.0;
    for (int d = depth - 1; d >= 0; d--) {
        score = 1.0 - score;
        pn[d]->edge_visits[pi[d]] += 1;
        pn[d]->edge_total[pi[d]] += score;
        pn[d]->all_edge_visits++;
    }
}

int  ao_mcts_root_visits(const ao_mcts *m) { return m->root->all_edge_visits; }
long ao_mcts_eval_count(const ao_mcts *m)  { return m->evals; }
long ao_mcts_tie_count(const ao_mcts *m)   { return m->ties; }
void ao_mcts_root_position(const ao_mcts *m, ao_position *out) { *out = m->root->board; }

int ao_mcts_root_dist(const ao_mcts *m, int32_t *from, int32_t *to, int32_t *visits,
                      double *total_score, double *prior)
{
    const ao_node *r = m->root;
    for (int i = 0; i < r->n_moves; i++) {
        if (from) from[i] = r->from[i];
        if (to) to[i] = r->to[i];
        if (visits) visits[i] = (int32_t)r->edge_visits[i];
        if (total_score) total_score[i] = r->edge_total[i];
        if (prior) prior[i] = r->prior[i];
    }
    return r->n_moves;
}

int ao_mcts_play(ao_mcts *m, int from, int to)
{
    ao_node *r = m->root;
    int idx = -1;
    for (int i = 0; i < r->n_moves; i++)
        if (r->from[i] == from && r->to[i] == to) idx = i;
    if (idx < 0 || !r->child[idx]) {
        /* miss: throw everything away and start from the moved board */
        ao_position nb = r->board;
        ao_makemove(&nb, from, to);
        node_free(r);
        m->root = node_new(&nb);
        node_populate(m, m->root);
        return 0;
    }
    ao_node *keep = r->child[idx];
    r->child[idx] = NULL;
    node_free(r);
    m->root = keep;
    /* the reference re-populates the new root (clear() keeps the map's bucket count) */
    keep->populated = 0;
    node_populate(m, keep);
    return 1;
}

/* ------------------------------------------------------------------ */
/* fixture evaluators                                                 */
/* ------------------------------------------------------------------ */

void ao_probe_eval(void *ctx, const float feats[AO_FEATURES], float logits[AO_LOGITS], float *value)
{
    (void)ctx;
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < AO_FEATURES; i++) {
        h ^= (uint64_t)(feats[i] != 0.0f) + (uint64_t)i * 2;
        h *= 1099511628211ULL;
    }
    for (int i = 0; i < AO_LOGITS; i++) {
        h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33;
        logits[i] = (float)((double)(h & 0xffffff) / (double)0x1000000 * 4.0 - 2.0);
    }
    h ^= h >> 29; h *= 0xc4ceb9fe1a85ec53ULL;
    *value = (float)((double)(h & 0xffffff) / (double)0x1000000 * 1.6 - 0.8);
}

void ao_probe_eval_batch(const float *feats, int n, float *logits, float *values)
{
    for (int i = 0; i < n; i++) ao_probe_eval(NULL, feats + (size_t)i * AO_FEATURES, logits + (size_t)i * AO_LOGITS, values + i);
}

void ao_uniform_eval(void *ctx, const float feats[AO_FEATURES], float logits[AO_LOGITS], float *value)
{
    (void)ctx; (void)feats;
    for (int i = 0; i < AO_LOGITS; i++) logits[i] = 0.0f;
    *value = 0.0f;
}

int ao_search_fen(const char *fen, int visits, int evaluator, int32_t *from, int32_t *to,
                  int32_t *visit_out, long *evals_out)
{
    ao_position p;
    if (ao_set_board(&p, fen) != 0) return -1;
    ao_mcts *m = ao_mcts_new(&p, evaluator == 1 ? ao_uniform_eval : ao_probe_eval, NULL);
    while (ao_mcts_root_visits(m) < visits) ao_mcts_step(m);
    int n = ao_mcts_root_dist(m, from, to, visit_out, NULL, NULL);
    if (evals_out) *evals_out = ao_mcts_eval_count(m);
    ao_mcts_free(m);
    return n;
}
