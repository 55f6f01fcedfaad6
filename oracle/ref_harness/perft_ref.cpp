// perft_ref.cpp -- OUR driver around the REFERENCE's own movegen()/makemove()
// (cpp/movegen.cpp:10, cpp/makemove.cpp:56), compiled by oracle/Makefile into
// oracle/_ref/perft_ref.  Test/baseline infrastructure only.
//
// usage: perft_ref "<fen>" <depth> [threads]
// prints: nodes <N> seconds <S> threads <T>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <thread>
#include <vector>

#include "ataxx.hpp"
#include "movegen.hpp"
#include "makemove.hpp"

static uint64_t walk(const Position &pos, int depth)
{
    if (depth <= 0) return 1;
    Move moves[256];
    int n = movegen(pos, moves);
    if (depth == 1) return (uint64_t)n;
    uint64_t total = 0;
    for (int i = 0; i < n; ++i) {
        Position child = pos;
        makemove(child, moves[i]);
        total += walk(child, depth - 1);
    }
    return total;
}

extern "C" uint64_t ref_perft(const char *fen, int depth, int threads)
{
    Position root;
    if (set_board(root, fen) != 0) return ~0ull;
    if (threads <= 1 || depth < 4) return walk(root, depth);
    std::vector<Position> frontier;
    Move m1[256], m2[256];
    int n1 = movegen(root, m1);
    for (int i = 0; i < n1; ++i) {
        Position a = root; makemove(a, m1[i]);
        int n2 = movegen(a, m2);
        for (int j = 0; j < n2; ++j) { Position b = a; makemove(b, m2[j]); frontier.push_back(b); }
    }
    std::atomic<size_t> cursor{0};
    std::atomic<uint64_t> total{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            uint64_t local = 0;
            for (size_t i; (i = cursor.fetch_add(1)) < frontier.size();) local += walk(frontier[i], depth - 2);
            total += local;
        });
    for (auto &t : pool) t.join();
    return total;
}

#ifndef PERFT_REF_NO_MAIN
int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s fen depth [threads]\n", argv[0]); return 2; }
    int depth = atoi(argv[2]);
    int threads = argc > 3 ? atoi(argv[3]) : 1;
    auto t0 = std::chrono::steady_clock::now();
    uint64_t nodes = ref_perft(argv[1], depth, threads);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("nodes %llu seconds %.6f threads %d\n", (unsigned long long)nodes, s, threads);
    return 0;
}
#endif
