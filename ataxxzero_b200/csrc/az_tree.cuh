// az_tree.cuh -- device-resident PUCT trees: memory layout shared by az_tree.cu (kernels)
// and az_pool.cu (host runtime).  Internal.
//
// One tree per game, one fixed-stride slot per node (no allocator metadata on the hot path):
//
//   node slot (9216 B, 128-B aligned):
//     [   0,   64)  header: position, value, child count L, visit count N, flags
//     [  64, 2112)  P[256]      f64  prior                      (self_play_client.cpp:151 posterior)
//     [2112, 4160)  W[256]      f64  edge_total_score           (:283)
//     [4160, 6208)  Q[256]      f64  W/n, refreshed by backup   (:288-292 get_edge_score; select never divides W)
//     [6208, 7232)  n[256]      u32  edge_visits                (:282; integral, so exact as u32)
//     [7232, 8256)  child[256]  i32  child node index, -1 = no edge yet (:281)
//     [8256, 8768)  move[256]   u16  from | to<<8, reference movegen order (cpp/movegen.cpp:16-66)
//     [8768, 9024)  rank[256]   u8   position of the move in the reference's hash-map iteration order
//                                    (select_action's `>=` tie-break walks that order, :345-358); filled in
//                                    lazily, the first time a node actually sees an exact tie at its maximum
//   Children of a node are a struct-of-arrays inside its slot, so a warp reads P/W/n of 32 children
//   with three coalesced loads.  L < 256 is the reference's own bound (movegen.cpp:69).
//
// Dead subtrees are recycled lazily: re-rooting pushes the discarded siblings on a per-game stack;
// allocating a node pops one entry and pushes that node's children.  The pool therefore never holds
// more than (live nodes + 1) slots and is never compacted or copied.
#pragma once
#include "az_common.h"

namespace aztree {

constexpr int kNodeStride = 9216;
constexpr int kOffP = 64, kOffW = 2112, kOffQ = 4160, kOffN = 6208, kOffChild = 7232, kOffMove = 8256, kOffRank = 8768;
constexpr int kMaxPath = 1024;
// child[] entries: bits 0..23 node index, bits 24..27 ceil(L_child / 32) -- how many 32-child groups a descent must
// load for that child, so a level fetches the child's real fan-out instead of a fixed 128 entries; -1 = no edge yet
constexpr int32_t kChildMask = 0x00ffffff;
constexpr int kChildGroupShift = 24;
constexpr int kRecWordsPerPly = 6 + 2 * 256;    // worst-case record words per ply

enum : uint32_t {
    NF_TERMINAL = 1u,
    NF_POPULATED = 2u,
    NF_RANKED = 4u,      // rank[] is valid for the node's current population
    NF_REPOPULATED = 8u  // the node became a root: the reference re-inserted its posterior into a clear()ed map (:489-490)
};
enum : int32_t { ST_IDLE = 0, ST_WAIT = 1, ST_DONE = 2, ST_STALL = 3, ST_ERROR = 4, ST_DESCEND = 5 };
enum : int32_t { ERR_NODES = 1, ERR_PATH = 2, ERR_MOVES = 3 };

struct __align__(16) NodeHdr {
    uint64_t own, opp;       // pieces of the side to move / of the opponent
    double value;            // evals.value: from the side to move's point of view
    int32_t n_moves;         // L; 0 for adjudicated (terminal) nodes
    int32_t N;               // all_edge_visits
    int32_t turn;            // absolute side to move: 0 = x, 1 = o
    uint32_t flags;
    int32_t reserved;
    int32_t pad[5];
};
static_assert(sizeof(NodeHdr) == 64, "node header is 64 bytes");

struct __align__(16) Game {
    int32_t status;
    int32_t root;
    int32_t pending;         // node awaiting an evaluation (ST_WAIT)
    int32_t path_len;
    int32_t n_alloc;         // bump pointer of never-used node slots
    int32_t gsp;             // garbage stack pointer
    int32_t ply;
    int32_t req_slot;
    uint64_t blockers;
    uint32_t rec_words;      // words written to the current record buffer
    int32_t rec_buf;         // 0/1
    int32_t rec_plies;
    uint32_t games_started;  // RNG stream selector
    int32_t rec_busy[2];     // record buffer handed to the host and not yet released
    int32_t error;
    int32_t start_turn;
    uint64_t start_own, start_opp;    // starting position of this slot (side to move / opponent)
    // statistics (summed by the host on demand)
    unsigned long long steps, evals, terminal_steps, positions, finished, skipped, max_depth;
    int32_t pad2[2];
};

struct DoneEntry {
    int32_t game, buf, words, plies, result, pad[3];
};

struct PoolDev {
    uint8_t *nodes;          // [G][C][kNodeStride]
    Game *games;             // [G]
    uint32_t *path;          // [G][kMaxPath]  node << 8 | slot
    uint32_t *gstack;        // [G][C]
    az_position *req_pos;    // [G]
    int32_t *req_game;       // [G]
    int32_t *req_count;      // [3 slots][2]: {requests, games that still have work but made no request} of a tick.  Ticks
                             // rotate through the slots (tick_slot); a tick reads its predecessor's request count (of which
                             // the first `cap` were evaluated) and zeroes its successor's slot, so the host never memsets.
    float *logits;           // [G][833]
    float *values;           // [G]
    uint32_t *records;       // [G][2][rec_cap_words]
    DoneEntry *done;         // [2G]
    int32_t *done_count;     // [1]
    int32_t G, C, visits, max_plies, noise, auto_play, steps_per_tick;
    int32_t levels_per_tick; // bound on tree levels a game may descend per tick (tail latency of deep endgame lines)
    int32_t cap;             // evaluations served per tick: a whole number of net-kernel rounds; later requests are re-queued
    int32_t consume;         // 1: evaluations of the previous requests are in logits/values; 0: top-up tick, leave waiting games alone
    int32_t tick_slot;       // slot of req_count this tick appends its requests to
    int32_t game_base;       // global index of this group's game 0 (RNG streams are keyed by the global game index)
    uint32_t rec_cap_words;
    uint64_t seed;
    int32_t full_fetch;      // experiment (AZ_TREE_FULL_FETCH=1): always load 128 children per level, ignoring the edge's group count
    unsigned long long *prof;   // AZ_POOL_PROFILE=1: [G][8] clock cycles per phase of the last tick (debug aid, normally nullptr)
};

}  // namespace aztree
