// ref_shim.cpp -- OUR C-ABI shim around the REFERENCE's own search core.
//
// oracle/Makefile concatenates  [cpp/self_play_client.cpp lines 1..582]  +  this
// file  into one translation unit (nothing of the reference is copied into the
// repo), so everything above the reference's "Threaded Workload" banner -- rules
// adjudication, Evaluations::populate, MCTSNode/MCTSEdge/MCTS -- is the real
// thing, and only request_evaluation() (declared at self_play_client.cpp:146)
// is ours: it forwards to an injected evaluator so a search is reproducible.
// Output: oracle/_ref/libref.so.  Test infrastructure only.
#include <cstring>
#include "bitboards.hpp"
#include "invalid.hpp"

extern "C" {
typedef void (*ref_eval_fn)(void *ctx, const float *feats, float *logits, float *value);
}

static ref_eval_fn g_eval = nullptr;
static void *g_eval_ctx = nullptr;
static float g_logits[7 * 7 * 17];
static float g_last_features[7 * 7 * 4];
static long g_eval_calls = 0;

std::pair<const float*, double> request_evaluation(int, const float* feature_string) {
	std::memcpy(g_last_features, feature_string, sizeof(g_last_features));
	float v = 0.0f;
	g_eval(g_eval_ctx, feature_string, g_logits, &v);
	g_eval_calls++;
	return {g_logits, (double)v};
}

struct RefPos { int ply, turn; uint64_t blockers, pieces[2]; };
static Position to_pos(const RefPos* p) { Position q; q.ply = p->ply; q.turn = p->turn; q.blockers = p->blockers; q.pieces[0] = p->pieces[0]; q.pieces[1] = p->pieces[1]; return q; }
static void from_pos(const Position& q, RefPos* p) { p->ply = q.ply; p->turn = q.turn; p->blockers = q.blockers; p->pieces[0] = q.pieces[0]; p->pieces[1] = q.pieces[1]; }

extern "C" {

int ref_set_board(RefPos* out, const char* fen) { Position p; int r = set_board(p, fen); from_pos(p, out); return r; }
uint64_t ref_single_jump_bb(uint64_t bb) { return single_jump_bb(bb); }
uint64_t ref_double_jump_bb(uint64_t bb) { return double_jump_bb(bb); }
uint64_t ref_single_ring(int sq) { return single_jump_sq(sq); }
uint64_t ref_double_ring(int sq) { return double_jump_sq(sq); }

int ref_movegen(const RefPos* p, int* from, int* to) {
	Move buf[256];
	int n = movegen(to_pos(p), buf);
	for (int i = 0; i < n; i++) { from[i] = buf[i].from; to[i] = buf[i].to; }
	return n;
}

void ref_makemove(RefPos* p, int from, int to) { Position q = to_pos(p); makemove(q, Move(from, to)); from_pos(q, p); }
int ref_legal_move(const RefPos* p, int from, int to) { return legal_move(to_pos(p), Move(from, to)); }
int ref_result(const RefPos* p) { return get_board_result(to_pos(p)); }
int ref_move_string(int from, int to, char* out) { std::string s = move_string(Move(from, to)); std::strcpy(out, s.c_str()); return (int)s.size(); }

void ref_board_json(const RefPos* p, int* out) {
	std::vector<int> v = serialize_board_for_json(to_pos(p));
	for (int i = 0; i < 49; i++) out[i] = v[i];
}

// features + priors exactly as Evaluations::populate produces them (noise off).
// Returns n_moves, or -1 when the position is terminal (then *value is the terminal value).
int ref_populate(const RefPos* p, ref_eval_fn fn, void* ctx, float* features_out, int* from, int* to, double* prior, double* value) {
	g_eval = fn; g_eval_ctx = ctx;
	Evaluations ev;
	Position q = to_pos(p);
	ev.populate(0, q, false);
	*value = ev.value;
	if (ev.game_over) return -1;
	std::memcpy(features_out, g_last_features, sizeof(g_last_features));
	Move buf[256];
	int n = movegen(q, buf);
	for (int i = 0; i < n; i++) { from[i] = buf[i].from; to[i] = buf[i].to; prior[i] = ev.posterior.at(buf[i]); }
	return n;
}

// iteration order of a real std::unordered_map<Move,double>, fresh or after clear()+reinsert
int ref_umap_order(const int* from, const int* to, int n, int reinsert, int* order_out) {
	std::unordered_map<Move, double> m;
	for (int pass = 0; pass <= (reinsert ? 1 : 0); pass++) {
		m.clear();
		for (int i = 0; i < n; i++) m.insert({Move(from[i], to[i]), (double)i});
	}
	int k = 0;
	for (auto& kv : m) order_out[k++] = (int)kv.second;
	return (int)m.bucket_count();
}

// One search from `fen` (noise off): step until root.all_edge_visits >= visits.
// Root distribution is returned in movegen order (visits 0 where no edge exists).
int ref_search(const char* fen, int visits, ref_eval_fn fn, void* ctx, int* from, int* to, int* visit_out, double* score_out, long* evals_out) {
	g_eval = fn; g_eval_ctx = ctx; g_eval_calls = 0;
	Position b;
	if (set_board(b, fen) != 0) return -1;
	MCTS mcts(0, b, false);
	while (mcts.root_node->all_edge_visits < visits) mcts.step();
	Move buf[256];
	int n = movegen(b, buf);
	for (int i = 0; i < n; i++) {
		from[i] = buf[i].from; to[i] = buf[i].to; visit_out[i] = 0;
		if (score_out) score_out[i] = 0;
		auto it = mcts.root_node->outgoing_edges.find(buf[i]);
		if (it != mcts.root_node->outgoing_edges.end()) {
			visit_out[i] = (int)it->second.edge_visits;
			if (score_out) score_out[i] = it->second.edge_total_score;
		}
	}
	if (evals_out) *evals_out = g_eval_calls;
	return n;
}

// Multi-ply game with tree reuse (noise off): at each ply search to `visits`, record the
// root distribution, then play the most-visited move (ties: first in movegen order).
// dist_out is [max_plies][256] visits in movegen order, n_moves_out[ply] the move count,
// played_out[ply] = from | to<<8.  Returns plies played; *result_out = get_board_result.
int ref_selfplay_greedy(const char* fen, int visits, int max_plies, ref_eval_fn fn, void* ctx,
                        int* n_moves_out, int* dist_out, int* played_out, int* result_out, long* evals_out) {
	g_eval = fn; g_eval_ctx = ctx; g_eval_calls = 0;
	Position b;
	if (set_board(b, fen) != 0) return -1;
	MCTS mcts(0, b, false);
	int ply = 0;
	for (; ply < max_plies; ply++) {
		while (mcts.root_node->all_edge_visits < visits) mcts.step();
		Move buf[256];
		int n = movegen(mcts.root_board, buf);
		n_moves_out[ply] = n;
		int best = -1, best_visits = -1;
		for (int i = 0; i < n; i++) {
			int v = 0;
			auto it = mcts.root_node->outgoing_edges.find(buf[i]);
			if (it != mcts.root_node->outgoing_edges.end()) v = (int)it->second.edge_visits;
			dist_out[ply * 256 + i] = v;
			if (v > best_visits) { best_visits = v; best = i; }
		}
		played_out[ply] = buf[best].from | (buf[best].to << 8);
		mcts.play(buf[best]);
		if (get_board_result(mcts.root_node->board) != 0) { ply++; break; }
	}
	*result_out = get_board_result(mcts.root_node->board);
	if (evals_out) *evals_out = g_eval_calls;
	return ply;
}

} // extern "C"
