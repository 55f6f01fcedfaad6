// az_samples.cu -- training-sample extraction on the GPU: the step immediately downstream of
// self-play.  Replaces get_sample_from_entries of train.py:43-77 (board -> feature planes, visit distribution ->
// policy heat-map, dihedral symmetry of both, value target) for a whole minibatch at once, reading the binary
// ply records the self-play kernels write (az_tree.cu make_move) -- or the same layout packed from JSON files.
//
// Ply record (uint32 words): [0,1] x bitboard, [2,3] o bitboard, [4] = played move | entries << 16,
// [5] = N (sum of visit counts; 0: the second word of each pair is a float32 probability), then `entries`
// pairs (move, count | probability bits).  bit sq = x + 7*(6-y), moves are from | to << 8 (include/ataxxzero.h).
//
// One warp per sample; the 833-float heat-map and the 196 feature bytes are assembled in shared memory and leave
// with coalesced stores.  HBM-bound: 3532 output bytes + (24 + 8*entries) input bytes per sample.
#include "az_common.h"
#include "az_rules.cuh"

namespace {

constexpr int kWarps = 8;

// train.py:25-40 apply_symmetry_to_move / :11-23 apply_symmetry: where cell (x, y) of the original lands
__device__ __forceinline__ void sym_forward(int sym, int &x, int &y)
{
    if (sym & 1) x = 6 - x;
    if (sym & 2) y = 6 - y;
    if (sym & 4) { const int t = x; x = y; y = t; }
}

__global__ void __launch_bounds__(kWarps * 32) k_extract_samples(const uint32_t *__restrict__ plies, const unsigned long long *__restrict__ offsets,
                                                                 const uint32_t *__restrict__ meta, int n, int8_t *__restrict__ features,
                                                                 float *__restrict__ policy, float *__restrict__ value)
{
    __shared__ float s_policy[kWarps][AZ_LOGITS + 3];
    __shared__ uint32_t s_feat[kWarps][49];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * kWarps + w;
    if (s >= n) return;
    const uint32_t *rec = plies + offsets[s];
    const uint32_t m = meta[s];
    const int to_move = (int)(m & 1u);            // 0: x (player 1) to move, 1: o
    const int result = (int)((m >> 1) & 3u);      // 1 | 2
    const int sym = (int)((m >> 3) & 7u);
    const bool use_dist = ((m >> 6) & 1u) != 0;
    const uint64_t bx = (uint64_t)rec[0] | ((uint64_t)rec[1] << 32), bo = (uint64_t)rec[2] | ((uint64_t)rec[3] << 32);
    const uint64_t own = to_move ? bo : bx, opp = to_move ? bx : bo;
    for (int i = lane; i < AZ_LOGITS; i += 32) s_policy[w][i] = 0.f;
    // engine.board_to_features (engine.py:53-73) + symmetry: plane 0 ones, 1 side to move, 2 opponent, 3 blocked cells
    // (always 0 here: the reference's training boards carry no blockers)
    for (int c = lane; c < 49; c += 32) {
        int x = c % 7, y = c / 7;                 // source cell
        const uint64_t bit = 1ULL << (x + 7 * (6 - y));
        sym_forward(sym, x, y);
        s_feat[w][x * 7 + y] = 1u | ((own & bit) ? 1u << 8 : 0u) | ((opp & bit) ? 1u << 16 : 0u);
    }
    __syncwarp();
    // visit distribution -> heat-map (engine.add_move_to_heatmap, engine.py:79-88), moves transformed by the symmetry
    const int entries = use_dist ? (int)(rec[4] >> 16) : 1;
    const double total = (double)rec[5];
    for (int e = lane; e < entries; e += 32) {
        uint32_t mv;
        float p;
        if (use_dist) {
            mv = rec[6 + 2 * e] & 0xffffu;
            const uint32_t v = rec[7 + 2 * e];
            p = rec[5] ? (float)((double)v / total) : __uint_as_float(v);      // n/N as the C++ client wrote it (double), then float32
        } else {
            mv = rec[4] & 0xffffu;                // {played move: 1} (train.py:62-63)
            p = 1.f;
        }
        const int from = (int)(mv & 0xff), to = (int)(mv >> 8);
        int ex = to % 7, ey = 6 - to / 7;
        sym_forward(sym, ex, ey);
        int plane = 16;
        if (from != to) {
            int sx = from % 7, sy = 6 - from / 7;
            sym_forward(sym, sx, sy);
            const int dx = ex - sx, dy = ey - sy;
            plane = dx == -2 ? dy + 2 : dx == 2 ? 13 + dy : 5 + 2 * (dx + 1) + (dy > 0 ? 1 : 0);
        }
        atomicAdd(&s_policy[w][119 * ex + 17 * ey + plane], p);
    }
    __syncwarp();
    float *po = policy + (size_t)s * AZ_LOGITS;
    for (int i = lane; i < AZ_LOGITS; i += 32) po[i] = s_policy[w][i];
    uint32_t *fo = reinterpret_cast<uint32_t *>(features + (size_t)s * AZ_FEATURES);
    for (int c = lane; c < 49; c += 32) fo[c] = s_feat[w][c];
    if (lane == 0) value[s] = result == to_move + 1 ? 1.f : -1.f;          // train.py:57
}

}  // namespace

extern "C" int az_samples_extract_dev(az_context *ctx, const void *d_plies, const void *d_offsets, const void *d_meta, int n,
                                      void *d_features, void *d_policy, void *d_value)
{
    AZ_REQUIRE(ctx && n >= 0, AZ_ERR_ARG, "az_samples_extract: bad argument");
    if (n == 0) return AZ_OK;
    AZ_REQUIRE(d_plies && d_offsets && d_meta && d_features && d_policy && d_value, AZ_ERR_ARG, "az_samples_extract: null pointer");
    k_extract_samples<<<(n + kWarps - 1) / kWarps, kWarps * 32, 0, ctx->stream>>>(
        static_cast<const uint32_t *>(d_plies), static_cast<const unsigned long long *>(d_offsets), static_cast<const uint32_t *>(d_meta), n,
        static_cast<int8_t *>(d_features), static_cast<float *>(d_policy), static_cast<float *>(d_value));
    ctx->launches++;
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

extern "C" int az_samples_extract(az_context *ctx, const uint32_t *plies, size_t ply_words, const uint64_t *offsets, const uint32_t *meta,
                                  int n, int8_t *features, float *policy, float *value)
{
    AZ_REQUIRE(ctx && n >= 0, AZ_ERR_ARG, "az_samples_extract: bad argument");
    if (n == 0) return AZ_OK;
    AZ_REQUIRE(plies && offsets && meta && features && policy && value, AZ_ERR_ARG, "az_samples_extract: null pointer");
    for (int i = 0; i < n; ++i) {               // a record must lie inside the table, entries included
        AZ_REQUIRE(offsets[i] + 6 <= ply_words, AZ_ERR_ARG, "az_samples_extract: sample %d points outside the ply table", i);
        const size_t need = ((meta[i] >> 6) & 1u) ? 6 + 2 * (size_t)(plies[offsets[i] + 4] >> 16) : 6;
        AZ_REQUIRE(offsets[i] + need <= ply_words, AZ_ERR_ARG, "az_samples_extract: sample %d has a truncated record", i);
        const int result = (int)((meta[i] >> 1) & 3u);
        AZ_REQUIRE(result == 1 || result == 2, AZ_ERR_ARG, "az_samples_extract: sample %d has result %d", i, result);
    }
    const size_t nn = (size_t)n;
    AzBuffer *b = ctx->scratch;
    AZ_REQUIRE(b[0].reserve(ply_words * 4) == 0 && b[1].reserve(nn * 8) == 0 && b[2].reserve(nn * 4) == 0 && b[3].reserve(nn * AZ_FEATURES) == 0 &&
                   b[4].reserve(nn * AZ_LOGITS * 4) == 0 && b[5].reserve(nn * 4) == 0,
               AZ_ERR_CUDA, "az_samples_extract: device scratch alloc");
    cudaStream_t s = ctx->stream;
    AZ_CUDA(cudaMemcpyAsync(b[0].ptr, plies, ply_words * 4, cudaMemcpyHostToDevice, s));
    AZ_CUDA(cudaMemcpyAsync(b[1].ptr, offsets, nn * 8, cudaMemcpyHostToDevice, s));
    AZ_CUDA(cudaMemcpyAsync(b[2].ptr, meta, nn * 4, cudaMemcpyHostToDevice, s));
    int rc = az_samples_extract_dev(ctx, b[0].ptr, b[1].ptr, b[2].ptr, n, b[3].ptr, b[4].ptr, b[5].ptr);
    if (rc) return rc;
    AZ_CUDA(cudaMemcpyAsync(features, b[3].ptr, nn * AZ_FEATURES, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemcpyAsync(policy, b[4].ptr, nn * AZ_LOGITS * 4, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemcpyAsync(value, b[5].ptr, nn * 4, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}
