// az_api.cu -- context, error plumbing, host-side helpers and the batched rule kernels
// behind the C ABI (include/ataxxzero.h).
#include "az_rules.cuh"

#include <cstring>

static thread_local std::string g_last_error;

int az_fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int AzBuffer::reserve(size_t need)
{
    if (need <= bytes) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    size_t want = need + need / 4;
    if (cudaMalloc(&ptr, want) != cudaSuccess) {
        ptr = nullptr;
        cudaGetLastError();
        return -1;
    }
    bytes = want;
    return 0;
}

void AzBuffer::release()
{
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
}

void az_perft_release(az_context *ctx);
void az_net_release(az_context *ctx);

extern "C" const char *az_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *az_version(void) { return "ataxxzero-b200 0.1 (sm_100a)"; }

extern "C" int az_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int az_create(int device, uint64_t seed, az_context **out)
{
    AZ_REQUIRE(out, AZ_ERR_ARG, "az_create: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return az_fail(AZ_ERR_CUDA, "az_create: no CUDA device (%s); this library has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    AZ_REQUIRE(device >= 0 && device < n, AZ_ERR_ARG, "az_create: device %d out of range [0,%d)", device, n);
    AZ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    AZ_CUDA(cudaGetDeviceProperties(&prop, device));
    AZ_REQUIRE(prop.major >= 10, AZ_ERR_CUDA, "az_create: device %d is sm_%d%d; this build is sm_100a only", device,
               prop.major, prop.minor);
    az_context *ctx = new az_context();
    ctx->device = device;
    ctx->seed = seed;
    ctx->sm_count = prop.multiProcessorCount;
    AZ_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    *out = ctx;
    return AZ_OK;
}

extern "C" void az_destroy(az_context *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    az_perft_release(ctx);
    az_net_release(ctx);
    for (auto &b : ctx->scratch) b.release();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int az_sync(az_context *ctx)
{
    AZ_REQUIRE(ctx, AZ_ERR_ARG, "az_sync: null context");
    AZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return AZ_OK;
}

extern "C" void *az_stream(az_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

// ---------------------------------------------------------------------------------------
// host-side text helpers
// ---------------------------------------------------------------------------------------

// FEN grammar and status codes of cpp/ataxx.cpp:14-92 (rows rank 7..1; x/o pieces, '-' blocker,
// digits = empties; optional side-to-move token; "startpos" alias).
extern "C" int az_set_board(az_position *pos, const char *fen)
{
    if (!pos || !fen) return az_fail(AZ_ERR_ARG, "az_set_board: null argument");
    if (std::strcmp(fen, "startpos") == 0) fen = "x5o/7/3-3/2-1-2/3-3/7/o5x x";
    // tokens separated by single spaces; a trailing space adds no token
    std::vector<std::string> tok;
    {
        std::string cur;
        bool any = false;
        for (const char *c = fen; *c; ++c) {
            any = true;
            if (*c == ' ') { tok.push_back(cur); cur.clear(); }
            else cur.push_back(*c);
        }
        if (!cur.empty() || (any && fen[std::strlen(fen) - 1] != ' ')) tok.push_back(cur);
    }
    if (tok.empty()) return 1;
    if (tok.size() > 2) return 2;
    if (tok[0].size() < 13) return 3;
    if (tok[0].size() > 55) return 4;
    az_position p{};
    int sq = 42;
    for (char c : tok[0]) {
        if (c == 'x' || c == 'X') p.pieces[0] ^= 1ULL << (sq++ & 63);
        else if (c == 'o' || c == 'O') p.pieces[1] ^= 1ULL << (sq++ & 63);
        else if (c == '-') p.blockers ^= 1ULL << (sq++ & 63);
        else if (c >= '1' && c <= '7') sq += c - '0';
        else if (c == '/') sq -= 14;
        else return 5;
    }
    if (tok.size() > 1) {
        if (tok[1] == "x" || tok[1] == "X") p.turn = 0;
        else if (tok[1] == "o" || tok[1] == "O") p.turn = 1;
        else return 6;
    }
    *pos = p;
    if (sq != 7) return 7;
    if ((p.pieces[0] & p.pieces[1]) || ((p.pieces[0] | p.pieces[1]) & p.blockers) ||
        ((p.pieces[0] | p.pieces[1] | p.blockers) & ~az::kBoard))
        return 8;
    return 0;
}

extern "C" int az_move_string(az_move m, char out[5])
{
    const int f = AZ_MOVE_FROM(m), t = AZ_MOVE_TO(m);
    int n = 0;
    if (f != t) { out[n++] = char('a' + f % 7); out[n++] = char('1' + f / 7); }
    out[n++] = char('a' + t % 7);
    out[n++] = char('1' + t / 7);
    out[n] = 0;
    return n;
}

extern "C" az_move az_parse_move(const char *s)
{
    if (!s) return AZ_NO_MOVE;
    const size_t len = std::strlen(s);
    auto sq = [](const char *c) -> int {
        if (c[0] < 'a' || c[0] > 'g' || c[1] < '1' || c[1] > '7') return -1;
        return (c[1] - '1') * 7 + (c[0] - 'a');
    };
    if (len == 2) {
        int t = sq(s);
        return t < 0 ? AZ_NO_MOVE : AZ_MOVE(t, t);
    }
    if (len == 4) {
        int f = sq(s), t = sq(s + 2);
        return (f < 0 || t < 0) ? AZ_NO_MOVE : AZ_MOVE(f, t);
    }
    return AZ_NO_MOVE;
}

extern "C" int az_fen(const az_position *pos, char *out, size_t cap)
{
    if (!pos || !out || cap < 64) return az_fail(AZ_ERR_ARG, "az_fen: need a 64-byte buffer");
    int n = 0;
    for (int rank = 6; rank >= 0; --rank) {
        int run = 0;
        for (int file = 0; file < 7; ++file) {
            const uint64_t bit = 1ULL << (rank * 7 + file);
            char c = (pos->pieces[0] & bit) ? 'x' : (pos->pieces[1] & bit) ? 'o' : (pos->blockers & bit) ? '-' : 0;
            if (!c) { run++; continue; }
            if (run) { out[n++] = char('0' + run); run = 0; }
            out[n++] = c;
        }
        if (run) out[n++] = char('0' + run);
        if (rank) out[n++] = '/';
    }
    out[n++] = ' ';
    out[n++] = pos->turn ? 'o' : 'x';
    out[n] = 0;
    return n;
}

// ---------------------------------------------------------------------------------------
// batched rule kernels
// ---------------------------------------------------------------------------------------
namespace {

// One WARP per board: the warp ballots over the 49 squares / the per-source jump sets so the
// 256-entry move list of a board is written with coalesced stores in reference order.
__global__ void k_movegen(const az_position *pos, int n, az_move *moves, int32_t *counts)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const az_position p = pos[warp];
    const uint64_t own = p.pieces[p.turn & 1];
    const uint64_t empty = az::kBoard & ~(p.pieces[0] | p.pieces[1] | p.blockers);
    az_move *out = moves + (size_t)warp * AZ_MAX_MOVES;
    // jumps: lanes take sources round-robin; exclusive offsets come from a warp scan of counts
    int base = 0;
    for (uint64_t rest = own; rest;) {
        // peel up to 32 sources, one per lane (ascending)
        uint64_t mine = 0;
        int f = -1;
        uint64_t r = rest;
        for (int k = 0; k < 32 && r; ++k) {
            const int s = az::lsb64(r);
            r &= r - 1;
            if (k == lane) { f = s; mine = az::ring2_sq(s) & empty; }
        }
        rest = r;
        const int cnt = az::popc64(mine);
        int incl = cnt;
        for (int s = 1; s < 32; s <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += v;
        }
        int o = base + incl - cnt;
        for (; mine; mine &= mine - 1) out[o++] = AZ_MOVE(f, az::lsb64(mine));
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
    // clones: ascending destinations, lane k writes the k-th, k+32-th ... set bit
    uint64_t clones = az::ring1_bb(own) & empty;
    const int n_clones = az::popc64(clones);
    int k = 0;
    for (uint64_t c = clones; c; c &= c - 1, ++k)
        if ((k & 31) == lane) { const int t = az::lsb64(c); out[base + k] = AZ_MOVE(t, t); }
    if (lane == 0) counts[warp] = base + n_clones;
}

__global__ void k_makemove(az_position *pos, const az_move *moves, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    az_position p = pos[i];
    az::makemove(p, AZ_MOVE_FROM(moves[i]), AZ_MOVE_TO(moves[i]));
    pos[i] = p;
}

__global__ void k_result(const az_position *pos, int n, int32_t *result)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    az_position p = pos[i];
    p.turn &= 1;
    result[i] = az::board_result(p, nullptr);
}

// one thread per (board, cell): writes one float4 -> fully coalesced 16-byte stores
__global__ void k_features(const az_position *pos, int n, float4 *features)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 49) return;
    const int b = i / 49, cell = i % 49;
    az_position p = pos[b];
    p.turn &= 1;
    float v[4];
    az::feature_cell(p, cell / 7, cell % 7, v);
    features[i] = make_float4(v[0], v[1], v[2], v[3]);
}

__global__ void k_jump_bb(const uint64_t *bb, int n, uint64_t *s_out, uint64_t *d_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    s_out[i] = az::ring1_bb(bb[i]);
    d_out[i] = az::ring2_bb(bb[i]);
}

}  // namespace

#define AZ_STAGE_IN(slot, src, bytes)                                                                   \
    AZ_REQUIRE(ctx->scratch[slot].reserve(bytes) == 0, AZ_ERR_CUDA, "device scratch alloc of %zu bytes", (size_t)(bytes)); \
    AZ_CUDA(cudaMemcpyAsync(ctx->scratch[slot].ptr, src, bytes, cudaMemcpyHostToDevice, ctx->stream))
#define AZ_STAGE_OUT(slot, bytes) \
    AZ_REQUIRE(ctx->scratch[slot].reserve(bytes) == 0, AZ_ERR_CUDA, "device scratch alloc of %zu bytes", (size_t)(bytes))
#define AZ_FETCH(dst, slot, bytes) \
    AZ_CUDA(cudaMemcpyAsync(dst, ctx->scratch[slot].ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream))
#define AZ_FINISH()                               \
    AZ_CUDA(cudaStreamSynchronize(ctx->stream)); \
    AZ_CUDA(cudaGetLastError())

extern "C" int az_movegen_batch(az_context *ctx, const az_position *pos, int n, az_move *moves, int32_t *counts)
{
    AZ_REQUIRE(ctx && n >= 0 && (n == 0 || (pos && moves && counts)), AZ_ERR_ARG, "az_movegen_batch: bad argument");
    if (n == 0) return AZ_OK;
    AZ_STAGE_IN(0, pos, sizeof(az_position) * (size_t)n);
    AZ_STAGE_OUT(1, sizeof(az_move) * AZ_MAX_MOVES * (size_t)n);
    AZ_STAGE_OUT(2, sizeof(int32_t) * (size_t)n);
    const int warps_per_block = 8;
    k_movegen<<<(n + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, ctx->stream>>>(
        ctx->scratch[0].as<az_position>(), n, ctx->scratch[1].as<az_move>(), ctx->scratch[2].as<int32_t>());
    ctx->launches++;
    AZ_FETCH(counts, 2, sizeof(int32_t) * (size_t)n);
    AZ_FETCH(moves, 1, sizeof(az_move) * AZ_MAX_MOVES * (size_t)n);
    AZ_FINISH();
    return AZ_OK;
}

extern "C" int az_makemove_batch(az_context *ctx, az_position *pos, const az_move *moves, int n)
{
    AZ_REQUIRE(ctx && n >= 0 && (n == 0 || (pos && moves)), AZ_ERR_ARG, "az_makemove_batch: bad argument");
    if (n == 0) return AZ_OK;
    for (int i = 0; i < n; ++i)
        AZ_REQUIRE(AZ_MOVE_FROM(moves[i]) < 49 && AZ_MOVE_TO(moves[i]) < 49, AZ_ERR_ARG,
                   "az_makemove_batch: move %d is off the board", i);
    AZ_STAGE_IN(0, pos, sizeof(az_position) * (size_t)n);
    AZ_STAGE_IN(1, moves, sizeof(az_move) * (size_t)n);
    k_makemove<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->scratch[0].as<az_position>(), ctx->scratch[1].as<az_move>(), n);
    ctx->launches++;
    AZ_FETCH(pos, 0, sizeof(az_position) * (size_t)n);
    AZ_FINISH();
    return AZ_OK;
}

extern "C" int az_result_batch(az_context *ctx, const az_position *pos, int n, int32_t *result)
{
    AZ_REQUIRE(ctx && n >= 0 && (n == 0 || (pos && result)), AZ_ERR_ARG, "az_result_batch: bad argument");
    if (n == 0) return AZ_OK;
    AZ_STAGE_IN(0, pos, sizeof(az_position) * (size_t)n);
    AZ_STAGE_OUT(1, sizeof(int32_t) * (size_t)n);
    k_result<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->scratch[0].as<az_position>(), n, ctx->scratch[1].as<int32_t>());
    ctx->launches++;
    AZ_FETCH(result, 1, sizeof(int32_t) * (size_t)n);
    AZ_FINISH();
    return AZ_OK;
}

extern "C" int az_features_batch(az_context *ctx, const az_position *pos, int n, float *features)
{
    AZ_REQUIRE(ctx && n >= 0 && (n == 0 || (pos && features)), AZ_ERR_ARG, "az_features_batch: bad argument");
    if (n == 0) return AZ_OK;
    AZ_STAGE_IN(0, pos, sizeof(az_position) * (size_t)n);
    AZ_STAGE_OUT(1, sizeof(float) * AZ_FEATURES * (size_t)n);
    k_features<<<(n * 49 + 255) / 256, 256, 0, ctx->stream>>>(ctx->scratch[0].as<az_position>(), n, ctx->scratch[1].as<float4>());
    ctx->launches++;
    AZ_FETCH(features, 1, sizeof(float) * AZ_FEATURES * (size_t)n);
    AZ_FINISH();
    return AZ_OK;
}

extern "C" int az_jump_bb_batch(az_context *ctx, const uint64_t *bb, int n, uint64_t *single_out, uint64_t *double_out)
{
    AZ_REQUIRE(ctx && n >= 0 && (n == 0 || (bb && single_out && double_out)), AZ_ERR_ARG, "az_jump_bb_batch: bad argument");
    if (n == 0) return AZ_OK;
    AZ_STAGE_IN(0, bb, sizeof(uint64_t) * (size_t)n);
    AZ_STAGE_OUT(1, sizeof(uint64_t) * (size_t)n);
    AZ_STAGE_OUT(2, sizeof(uint64_t) * (size_t)n);
    k_jump_bb<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->scratch[0].as<uint64_t>(), n, ctx->scratch[1].as<uint64_t>(),
                                                         ctx->scratch[2].as<uint64_t>());
    ctx->launches++;
    AZ_FETCH(single_out, 1, sizeof(uint64_t) * (size_t)n);
    AZ_FETCH(double_out, 2, sizeof(uint64_t) * (size_t)n);
    AZ_FINISH();
    return AZ_OK;
}
