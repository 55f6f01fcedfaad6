// az_playout.cu -- uniformly random games played entirely on the device (BASELINE configs[0]:
// generate_games.py --random-play, generate_games.py:20-24,50-57): one thread per game, every ply
// draws uniformly from the side to move's legal moves (jumps by source / destination, then one clone per
// destination -- the move set of ataxx_rules.legal_moves / cpp movegen) with a Philox stream keyed by
// (seed, game), and records the position before the move and the move.
//
// Integer-ALU work; HBM traffic is the records only (24 bytes per ply, written once).
#include "az_common.h"
#include "az_rules.cuh"

namespace {

__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

struct PlayoutPly {
    uint64_t x, o;          // pieces before the move
    uint32_t move;          // from | to << 8
    uint32_t pad;
};
static_assert(sizeof(PlayoutPly) == 24, "24-byte ply records");

__global__ void k_random_playouts(az_position start, int n_games, int max_plies, uint64_t seed, PlayoutPly *__restrict__ plies,
                                  int32_t *__restrict__ n_plies, int32_t *__restrict__ result)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_games) return;
    az_position p = start;
    p.turn &= 1;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    PlayoutPly *rec = plies + (size_t)g * max_plies;
    int ply = 0, res = az::board_result(p, nullptr);
    uint4 rnd = make_uint4(0, 0, 0, 0);
    while (res == 0 && ply < max_plies) {
        const uint64_t own = p.pieces[p.turn];
        const uint64_t empty = az::kBoard & ~(p.pieces[0] | p.pieces[1] | p.blockers);
        const int n_jumps = az::count_jumps(own, empty);
        const uint64_t clones = az::ring1_bb(own) & empty;
        const int n = n_jumps + az::popc64(clones);          // > 0: board_result() would have adjudicated otherwise
        if ((ply & 3) == 0) rnd = philox4x32(make_uint4((uint32_t)g, (uint32_t)(ply >> 2), 0x504c4159u, 0u), key);
        const uint32_t r32 = (ply & 3) == 0 ? rnd.x : (ply & 3) == 1 ? rnd.y : (ply & 3) == 2 ? rnd.z : rnd.w;
        int k = (int)(((unsigned long long)r32 * (unsigned)n) >> 32);      // uniform index into the move list
        int from = -1, to = -1;
        if (k < n_jumps) {
            for (uint64_t src = own; src; src &= src - 1) {
                const int f = az::lsb64(src);
                uint64_t dst = az::ring2_sq(f) & empty;
                const int c = az::popc64(dst);
                if (k >= c) { k -= c; continue; }
                for (; k > 0; --k) dst &= dst - 1;
                from = f;
                to = az::lsb64(dst);
                break;
            }
        } else {
            uint64_t dst = clones;
            for (k -= n_jumps; k > 0; --k) dst &= dst - 1;
            from = to = az::lsb64(dst);
        }
        rec[ply].x = p.pieces[0];
        rec[ply].o = p.pieces[1];
        rec[ply].move = (uint32_t)from | ((uint32_t)to << 8);
        rec[ply].pad = 0;
        az::makemove(p, from, to);
        ++ply;
        res = az::board_result(p, nullptr);
    }
    n_plies[g] = ply;
    result[g] = res;
}

}  // namespace

extern "C" int az_random_playouts(az_context *ctx, const az_position *start, int n_games, int max_plies, uint64_t seed, void *plies_out,
                                  int32_t *n_plies_out, int32_t *result_out)
{
    AZ_REQUIRE(ctx && start && n_games >= 0 && max_plies >= 1 && max_plies <= 4096, AZ_ERR_ARG, "az_random_playouts: bad argument");
    if (n_games == 0) return AZ_OK;
    AZ_REQUIRE(plies_out && n_plies_out && result_out, AZ_ERR_ARG, "az_random_playouts: null output");
    const uint64_t all = start->pieces[0] | start->pieces[1] | start->blockers;
    AZ_REQUIRE(!(start->pieces[0] & start->pieces[1]) && !((start->pieces[0] | start->pieces[1]) & start->blockers) && !(all >> 49) &&
                   (start->pieces[0] | start->pieces[1]),
               AZ_ERR_ARG, "az_random_playouts: invalid start position");
    const size_t pb = sizeof(PlayoutPly) * (size_t)n_games * max_plies, ib = sizeof(int32_t) * (size_t)n_games;
    AzBuffer *b = ctx->scratch;
    AZ_REQUIRE(b[0].reserve(pb) == 0 && b[1].reserve(ib) == 0 && b[2].reserve(ib) == 0, AZ_ERR_CUDA, "az_random_playouts: device scratch alloc");
    cudaStream_t s = ctx->stream;
    k_random_playouts<<<(n_games + 127) / 128, 128, 0, s>>>(*start, n_games, max_plies, seed, b[0].as<PlayoutPly>(), b[1].as<int32_t>(),
                                                            b[2].as<int32_t>());
    ctx->launches++;
    AZ_CUDA(cudaMemcpyAsync(plies_out, b[0].ptr, pb, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemcpyAsync(n_plies_out, b[1].ptr, ib, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemcpyAsync(result_out, b[2].ptr, ib, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}
