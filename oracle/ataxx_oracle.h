/*
 * ataxx_oracle.h -- CPU restatement of the AtaxxZero self-play hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ataxxzero_b200/ links, imports or
 * executes this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may.  Each function cites the reference
 * file:line (relative to /root/reference) whose behaviour it restates.
 *
 * Parity status: rules / perft / adjudication / features / MCTS are PINNED
 * against the compiled reference (oracle/_ref, see oracle/Makefile) and the
 * golden vectors in SURVEY.md App. C (tests/test_oracle_pinned.py).
 */
#ifndef ATAXX_ORACLE_H
#define ATAXX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cpp/ataxx.hpp:28-34 -- same field order and meaning; turn 0 = x, 1 = o. */
typedef struct {
    int32_t  ply;
    int32_t  turn;
    uint64_t blockers;
    uint64_t pieces[2];
} ao_position;

#define AO_MAX_MOVES   256
#define AO_FEATURES    196   /* 7*7*4  */
#define AO_LOGITS      833   /* 7*7*17 */

/* bitboards */
uint64_t ao_single_ring(int sq);                 /* cpp/bitboards.hpp:32 */
uint64_t ao_double_ring(int sq);                 /* cpp/bitboards.hpp:33 */
uint64_t ao_single_jump_bb(uint64_t bb);         /* cpp/bitboards.cpp:6-16 */
uint64_t ao_double_jump_bb(uint64_t bb);         /* cpp/bitboards.cpp:18-39 */

/* position / rules */
int  ao_set_board(ao_position *pos, const char *fen);          /* cpp/ataxx.cpp:14-92 */
int  ao_invalid(const ao_position *pos);                       /* cpp/invalid.cpp:7-16 */
int  ao_movegen(const ao_position *pos, int32_t *from, int32_t *to); /* cpp/movegen.cpp:10-79 */
void ao_makemove(ao_position *pos, int from, int to);          /* cpp/makemove.cpp:56-76 */
int  ao_legal_move(const ao_position *pos, int from, int to);  /* cpp/move.cpp:62-98 */
int  ao_move_string(int from, int to, char out[5]);            /* cpp/move.cpp:11-21 */
int  ao_result(const ao_position *pos);                        /* cpp/self_play_client.cpp:109-144 */
void ao_features(const ao_position *pos, float out[AO_FEATURES]);      /* self_play_client.cpp:174-202 */
void ao_board_json(const ao_position *pos, int32_t out[49]);   /* self_play_client.cpp:88-107 */
int  ao_policy_index(int from, int to);                        /* self_play_client.cpp:77-86,224-237 */

/* perft: recursion over movegen+makemove, bulk count at depth 1. */
uint64_t ao_perft(const ao_position *pos, int depth);
/* same, root moves split over `threads` pthreads (for the cpu_baseline leg). */
uint64_t ao_perft_mt(const ao_position *pos, int depth, int threads);

/* priors: softmax in double + legal gather + renormalise (self_play_client.cpp:208-245). */
void ao_priors(const float logits[AO_LOGITS], const int32_t *from, const int32_t *to,
               int n_moves, double *prior_out);

/* iteration order of a libstdc++ std::unordered_map<Move,...> (hash = from + 49*to,
 * self_play_client.cpp:49-55) after inserting the moves in the given order.
 * start_buckets = 0 -> a fresh map; otherwise the bucket count retained by clear()
 * (root re-population, self_play_client.cpp:155,489-490).  order_out[k] = index (into
 * the insertion sequence) of the k-th element visited by iteration.  Returns the
 * final bucket count. */
int ao_umap_order(const int32_t *from, const int32_t *to, int n, int start_buckets,
                  int32_t *order_out);

/* ---- MCTS restatement (self_play_client.cpp:148-493) ---- */
typedef void (*ao_eval_fn)(void *ctx, const float feats[AO_FEATURES],
                           float logits[AO_LOGITS], float *value);

typedef struct ao_mcts ao_mcts;

ao_mcts *ao_mcts_new(const ao_position *root, ao_eval_fn fn, void *ctx);
void     ao_mcts_free(ao_mcts *m);
void     ao_mcts_step(ao_mcts *m);                         /* :419-473 */
int      ao_mcts_root_visits(const ao_mcts *m);            /* root.all_edge_visits */
long     ao_mcts_eval_count(const ao_mcts *m);
/* root edges in movegen order: visits (0 when no edge); returns n_moves at root */
int      ao_mcts_root_dist(const ao_mcts *m, int32_t *from, int32_t *to,
                           int32_t *visits, double *total_score, double *prior);
int      ao_mcts_play(ao_mcts *m, int from, int to);       /* :475-492 (noise off) */
void     ao_mcts_root_position(const ao_mcts *m, ao_position *out);
long     ao_mcts_tie_count(const ao_mcts *m);              /* #select_action calls that saw an exact tie at the max */

/* The deterministic probe evaluator of SURVEY App. D (FNV-style hash of the 196
 * feature non-zero flags -> 833 logits in [-2,2), value in [-0.8,0.8)). */
void ao_probe_eval(void *ctx, const float feats[AO_FEATURES], float logits[AO_LOGITS], float *value);
void ao_probe_eval_batch(const float *feats, int n, float *logits, float *values);
/* Degenerate evaluator: all logits 0, value 0 -> every prior ties (exercises map order). */
void ao_uniform_eval(void *ctx, const float feats[AO_FEATURES], float logits[AO_LOGITS], float *value);

/* test hook: 0 none, 1 heavy quantisation (exact ties), 2 adjacent-double pairs, 3 runs of 4 adjacent doubles */
void ao_set_prior_perturbation(int mode);

/* ---- tree_model.c: CPU model of the device tree algorithm (candidate + visited list), checked against ao_mcts ---- */
long tm_selftest(const char *fen, int visits, int evaluator, int plays, int force_slow, long *counters_out);
/* net round trips a bit-exact single-tree search needs when the top_k children of every consumed node are evaluated
 * speculatively (evaluator 2: nearly uniform priors, values ~0, like the benchmark's random-init net) */
long tm_spec_sim(const char *fen, int visits, int evaluator, int top_k, long *evals_out);

/* convenience: search `visits` from fen with a named evaluator (0 probe, 1 uniform),
 * fill root distribution; returns n_moves or <0 on error. */
int ao_search_fen(const char *fen, int visits, int evaluator, int32_t *from, int32_t *to,
                  int32_t *visit_out, long *evals_out);

#ifdef __cplusplus
}
#endif
#endif
