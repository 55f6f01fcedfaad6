// az_rules.cuh -- 7x7 Ataxx bitboard rules as inline host/device functions.
//
// Behavioural contract (not a translation): cpp/bitboards.cpp:6-39 (dilations),
// cpp/movegen.cpp:10-79 (move order), cpp/makemove.cpp:56-76, cpp/self_play_client.cpp:
// 109-144 (adjudication), :174-202 (features), :77-86/:224-237 (policy planes).
// A board is 49 bits of a u64, bit = rank*7 + file.  Everything here is register-only
// integer work: no tables in memory except the two ring tables (49 x u64 each), which the
// kernels stage into shared memory.
#pragma once
#include "az_common.h"

namespace az {

constexpr uint64_t kBoard = 0x1FFFFFFFFFFFFULL;
// file masks are generated from the one-file pattern (bit every 7)
constexpr uint64_t kFileA = 0x0040810204081ULL;
constexpr uint64_t file_mask(int f) { return kFileA << f; }
constexpr uint64_t kNotA = kBoard & ~file_mask(0);
constexpr uint64_t kNotG = kBoard & ~file_mask(6);
constexpr uint64_t kNotAB = kNotA & ~file_mask(1);
constexpr uint64_t kNotFG = kNotG & ~file_mask(5);

// 8-neighbourhood of every set bit (the bit itself only if it neighbours another one)
AZ_HD uint64_t ring1_bb(uint64_t bb)
{
    const uint64_t e = (bb << 1) & kNotA;          // one file right
    const uint64_t w = (bb >> 1) & kNotG;          // one file left
    const uint64_t row = bb | e | w;
    return (e | w | (row << 7) | (row >> 7)) & kBoard;
}

// squares at Chebyshev distance exactly 2 of any set bit
AZ_HD uint64_t ring2_bb(uint64_t bb)
{
    const uint64_t e1 = (bb << 1) & kNotA, w1 = (bb >> 1) & kNotG;
    const uint64_t e2 = (bb << 2) & kNotAB, w2 = (bb >> 2) & kNotFG;
    const uint64_t far = e2 | w2;                  // the two outer columns
    const uint64_t row5 = bb | e1 | w1 | far;      // all five columns
    return (far | (far << 7) | (far >> 7) | (row5 << 14) | (row5 >> 14)) & kBoard;
}

AZ_HD uint64_t ring1_sq(int sq) { return ring1_bb(1ULL << sq); }
AZ_HD uint64_t ring2_sq(int sq) { return ring2_bb(1ULL << sq); }

AZ_HD int popc64(uint64_t v)
{
#ifdef __CUDA_ARCH__
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
AZ_HD int lsb64(uint64_t v)
{
#ifdef __CUDA_ARCH__
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}

// number of (source, destination) jump pairs: one shifted copy of `own` per direction
AZ_HD int count_jumps(uint64_t own, uint64_t empty)
{
    const uint64_t a1 = own & kNotG, a2 = own & kNotFG;   // may move 1 / 2 files right
    const uint64_t b1 = own & kNotA, b2 = own & kNotAB;   // may move 1 / 2 files left
    int n = 0;
    // two ranks up / down, five file offsets each
    n += popc64((b2 << 12) & empty) + popc64((b1 << 13) & empty) + popc64((own << 14) & empty) +
         popc64((a1 << 15) & empty) + popc64((a2 << 16) & empty);
    n += popc64((a2 >> 12) & empty) + popc64((a1 >> 13) & empty) + popc64((own >> 14) & empty) +
         popc64((b1 >> 15) & empty) + popc64((b2 >> 16) & empty);
    // two files left / right, rank offsets -1, 0, +1
    n += popc64((a2 << 2) & empty) + popc64((a2 << 9) & empty) + popc64((a2 >> 5) & empty);
    n += popc64((b2 >> 2) & empty) + popc64((b2 << 5) & empty) + popc64((b2 >> 9) & empty);
    return n;
}

AZ_HD int count_moves(uint64_t own, uint64_t empty)
{
    return popc64(ring1_bb(own) & empty) + count_jumps(own, empty);
}

// state after the side `own` plays from->to (from == to: clone).  Returns the new
// (own, opp) from the MOVER's point of view; callers swap for the next side to move.
AZ_HD void apply_move(uint64_t &own, uint64_t &opp, int from, int to, uint64_t ring1_to)
{
    const uint64_t flipped = ring1_to & opp;
    own = (own & ~(1ULL << from)) | (1ULL << to) | flipped;
    opp &= ~flipped;
}

AZ_HD void makemove(az_position &p, int from, int to)
{
    uint64_t own = p.pieces[p.turn], opp = p.pieces[p.turn ^ 1];
    apply_move(own, opp, from, to, ring1_sq(to));
    p.pieces[p.turn] = own;
    p.pieces[p.turn ^ 1] = opp;
    p.turn ^= 1;
    p.ply += 1;
}

// reference move order: jumps by source then destination ascending, then clones by destination
AZ_HD int movegen(const az_position &p, az_move *out)
{
    const uint64_t own = p.pieces[p.turn];
    const uint64_t empty = kBoard & ~(p.pieces[0] | p.pieces[1] | p.blockers);
    int n = 0;
    for (uint64_t src = own; src; src &= src - 1) {
        const int f = lsb64(src);
        for (uint64_t dst = ring2_sq(f) & empty; dst; dst &= dst - 1) out[n++] = AZ_MOVE(f, lsb64(dst));
    }
    for (uint64_t dst = ring1_bb(own) & empty; dst; dst &= dst - 1) {
        const int t = lsb64(dst);
        out[n++] = AZ_MOVE(t, t);
    }
    return n;
}

// 0 ongoing, 1 x wins, 2 o wins; n_moves (of the side to move) is returned through *moves_out
AZ_HD int board_result(const az_position &p, int *moves_out)
{
    int x = popc64(p.pieces[0]), o = popc64(p.pieces[1]);
    const int blocked = popc64(p.blockers);
    const uint64_t empty = kBoard & ~(p.pieces[0] | p.pieces[1] | p.blockers);
    const int n = count_moves(p.pieces[p.turn], empty);
    if (moves_out) *moves_out = n;
    if (x == 0) return 2;
    if (o == 0) return 1;
    if (n == 0) {                       // stuck: the opponent is credited every empty cell
        if (p.turn == 0) o += popc64(empty); else x += popc64(empty);
    }
    if (x + o + blocked == 49) return x < o ? 2 : 1;
    return 0;
}

// flat logit index 119*to_x + 17*to_y + plane with y counted from the top rank
AZ_HD int policy_index(int from, int to)
{
    const int tx = to % 7, ty = 6 - to / 7;
    if (from == to) return 119 * tx + 17 * ty + 16;
    const int dx = tx - from % 7, dy = ty - (6 - from / 7);
    // planes enumerate (dx,dy) with max(|dx|,|dy|) == 2 in dx-major order:
    // dx=-2: dy=-2..2 -> 0..4 ; dx=-1: dy=-2 -> 5, dy=2 -> 6 ; dx=0: 7, 8 ; dx=1: 9, 10 ; dx=2: 11..15
    int plane;
    if (dx == -2) plane = dy + 2;
    else if (dx == 2) plane = 13 + dy;
    else plane = 5 + 2 * (dx + 1) + (dy > 0 ? 1 : 0);
    return 119 * tx + 17 * ty + plane;
}

// features [x][y][c] (index 28x+4y+c), y = 6 - rank
AZ_HD void feature_cell(const az_position &p, int x, int y, float out[4])
{
    const uint64_t bit = 1ULL << (x + 7 * (6 - y));
    out[0] = 1.0f;
    out[1] = (p.pieces[p.turn] & bit) ? 1.0f : 0.0f;
    out[2] = (p.pieces[p.turn ^ 1] & bit) ? 1.0f : 0.0f;
    out[3] = (p.blockers & bit) ? 1.0f : 0.0f;
}

}  // namespace az
