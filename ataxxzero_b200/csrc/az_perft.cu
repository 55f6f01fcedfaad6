// az_perft.cu -- perft over thousands of concurrent positions, sm_100a.
//
// Replaces the recursion a maintainer would write over cpp/movegen.cpp:10 + cpp/makemove.cpp:56
// (perft.py:5-26 is the Python twin).  Design:
//   * the frontier lives in HBM as 32-byte items (own, opp, blockers, root index), expanded
//     level by level ON THE DEVICE (count -> warp-aggregated slot reservation -> scatter) until
//     it is wide enough to fill 148 SMs x 2048 threads several times over;
//   * the remaining plies run as a register-only depth-first walk, one thread per frontier
//     item, with a flattened move iterator (one loop, no nested trip-count divergence) and a
//     popcount-only bulk count at the last ply (no move list is ever materialised);
//   * per-root totals are combined with warp shuffles and one 64-bit atomic per warp.
// The path is integer-ALU bound: HBM traffic is only the frontier (32 B per item, once).
#include "az_rules.cuh"

namespace {

struct __align__(32) PerftItem {
    uint64_t own, opp, blockers;
    uint32_t root, pad;
};

struct RingTables {
    uint64_t r1[49], r2[49];
};

__device__ __forceinline__ void load_tables(RingTables &t)
{
    for (int i = threadIdx.x; i < 49; i += blockDim.x) {
        t.r1[i] = az::ring1_sq(i);
        t.r2[i] = az::ring2_sq(i);
    }
    __syncthreads();
}

// flattened enumeration of the legal moves of `own`: jumps (by source), then clones
struct MoveIter {
    uint64_t src, dst, clones, empty;
    int from;
    __device__ __forceinline__ MoveIter(uint64_t own, uint64_t empty_) : src(own), dst(0), empty(empty_), from(0)
    {
        clones = az::ring1_bb(own) & empty_;
    }
    __device__ __forceinline__ bool next(const RingTables &t, int &f, int &to)
    {
        while (dst == 0 && src != 0) {
            from = az::lsb64(src);
            src &= src - 1;
            dst = t.r2[from] & empty;
        }
        if (dst) {
            to = az::lsb64(dst);
            dst &= dst - 1;
            f = from;
            return true;
        }
        if (clones) {
            to = az::lsb64(clones);
            clones &= clones - 1;
            f = to;
            return true;
        }
        return false;
    }
};

template <int R>
__device__ __forceinline__ uint64_t walk(uint64_t own, uint64_t opp, uint64_t blockers, const RingTables &t,
                                         unsigned long long &count_nodes)
{
    const uint64_t empty = az::kBoard & ~(own | opp | blockers);
    if constexpr (R <= 1) {
        count_nodes++;
        return (uint64_t)az::count_moves(own, empty);
    } else {
        uint64_t total = 0;
        MoveIter it(own, empty);
        int f, to;
        while (it.next(t, f, to)) {
            uint64_t o = own, p = opp;
            az::apply_move(o, p, f, to, t.r1[to]);
            total += walk<R - 1>(p, o, blockers, t, count_nodes);
        }
        return total;
    }
}

__global__ void k_pack(const az_position *pos, int n, PerftItem *items)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    az_position p = pos[i];
    PerftItem it;
    it.own = p.pieces[p.turn & 1];
    it.opp = p.pieces[(p.turn & 1) ^ 1];
    it.blockers = p.blockers;
    it.root = i;
    it.pad = 0;
    items[i] = it;
}

// total number of children of all items (grid-stride); one atomic per warp
__global__ void k_count(const PerftItem *items, int n, unsigned long long *total)
{
    unsigned long long local = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        PerftItem it = items[i];
        local += az::count_moves(it.own, az::kBoard & ~(it.own | it.opp | it.blockers));
    }
    for (int s = 16; s; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(total, local);
}

// one level of breadth-first expansion.  Output order is arbitrary (sums do not care).
__global__ void k_expand(const PerftItem *in, int n, PerftItem *out, unsigned long long *cursor)
{
    __shared__ RingTables t;
    load_tables(t);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    PerftItem it{};
    uint64_t empty = 0;
    int cnt = 0;
    if (i < n) {
        it = in[i];
        empty = az::kBoard & ~(it.own | it.opp | it.blockers);
        cnt = az::count_moves(it.own, empty);
    }
    // warp-inclusive scan of cnt, one reservation per warp
    int incl = cnt;
    for (int s = 1; s < 32; s <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += v;
    }
    unsigned long long base = 0;
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 31 && warp_total) base = atomicAdd(cursor, (unsigned long long)warp_total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (cnt == 0) return;
    PerftItem *dst = out + base + (incl - cnt);
    MoveIter mi(it.own, empty);
    int f, to;
    while (mi.next(t, f, to)) {
        uint64_t o = it.own, p = it.opp;
        az::apply_move(o, p, f, to, t.r1[to]);
        PerftItem c;
        c.own = p; c.opp = o; c.blockers = it.blockers; c.root = it.root; c.pad = 0;
        *dst++ = c;
    }
}

template <int R>
__global__ void __launch_bounds__(256) k_walk(const PerftItem *items, int n, unsigned long long *nodes,
                                              unsigned long long *count_nodes)
{
    __shared__ RingTables t;
    load_tables(t);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0, counted = 0;
    uint32_t root = 0xffffffffu;
    if (i < n) {
        PerftItem it = items[i];
        root = it.root;
        mine = walk<R>(it.own, it.opp, it.blockers, t, counted);
    }
    // combine per root: the common case is a warp whose lanes all share one root
    const uint32_t root0 = __shfl_sync(0xffffffffu, root, 0);
    const bool uniform = __all_sync(0xffffffffu, root == root0 || root == 0xffffffffu);
    for (int s = 16; s; s >>= 1) counted += __shfl_xor_sync(0xffffffffu, counted, s);
    if ((threadIdx.x & 31) == 0 && counted) atomicAdd(count_nodes, counted);
    if (uniform && root0 != 0xffffffffu) {
        for (int s = 16; s; s >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, s);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&nodes[root0], mine);
    } else if (root != 0xffffffffu && mine) {
        atomicAdd(&nodes[root], mine);
    }
}

}  // namespace

struct AzPerftState {
    AzBuffer frontier[2];
    AzBuffer counters;          // [0] cursor / child total, [1] count_nodes
    unsigned long long *h_counters = nullptr;   // pinned mirror
    uint64_t last_count_nodes = 0;
    int last_launches = 0;
    int last_frontier = 0;
};

static int perft_state(az_context *ctx, AzPerftState **out)
{
    if (!ctx->perft) {
        ctx->perft = new AzPerftState();
        AZ_CUDA(cudaMallocHost(&ctx->perft->h_counters, 4 * sizeof(unsigned long long)));
    }
    AZ_REQUIRE(ctx->perft->counters.reserve(4 * sizeof(unsigned long long)) == 0, AZ_ERR_CUDA, "perft counters alloc");
    *out = ctx->perft;
    return AZ_OK;
}

void az_perft_release(az_context *ctx)
{
    if (!ctx->perft) return;
    ctx->perft->frontier[0].release();
    ctx->perft->frontier[1].release();
    ctx->perft->counters.release();
    if (ctx->perft->h_counters) cudaFreeHost(ctx->perft->h_counters);
    delete ctx->perft;
    ctx->perft = nullptr;
}

template <int R>
static void launch_walk(az_context *ctx, const PerftItem *items, int n, unsigned long long *nodes,
                        unsigned long long *count_nodes)
{
    k_walk<R><<<(n + 255) / 256, 256, 0, ctx->stream>>>(items, n, nodes, count_nodes);
}

// Core: d_pos (az_position[n], device) -> d_nodes (u64[n], device).
extern "C" int az_perft_batch_dev(az_context *ctx, const void *d_pos, int n, int depth, void *d_nodes)
{
    AZ_REQUIRE(ctx && d_pos && d_nodes, AZ_ERR_ARG, "az_perft_batch_dev: null argument");
    AZ_REQUIRE(n >= 0 && depth >= 0 && depth <= 16, AZ_ERR_ARG, "az_perft_batch_dev: bad n=%d depth=%d", n, depth);
    AzPerftState *st;
    int rc = perft_state(ctx, &st);
    if (rc) return rc;
    st->last_count_nodes = 0;
    st->last_launches = 0;
    st->last_frontier = n;
    if (n == 0) return AZ_OK;
    cudaStream_t s = ctx->stream;
    unsigned long long *nodes = static_cast<unsigned long long *>(d_nodes);
    unsigned long long *ctr = st->counters.as<unsigned long long>();

    if (depth == 0) {   // perft(0) = 1 per position
        std::vector<unsigned long long> ones((size_t)n, 1ULL);
        AZ_CUDA(cudaMemcpyAsync(nodes, ones.data(), sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, s));
        AZ_CUDA(cudaStreamSynchronize(s));
        return AZ_OK;
    }

    // frontier growth policy: expand while the frontier is too narrow to fill the machine and
    // the next level fits the cap; leave 1..4 plies for the register-only walk.
    const size_t kCapItems = (size_t)12 << 20;                       // 12 Mi items = 384 MiB per buffer
    const size_t kWide = (size_t)ctx->sm_count * 2048 * 8;           // ~8 full waves of threads
    AZ_REQUIRE(st->frontier[0].reserve(sizeof(PerftItem) * (size_t)n) == 0, AZ_ERR_CUDA, "perft frontier alloc");
    AZ_CUDA(cudaMemsetAsync(nodes, 0, sizeof(unsigned long long) * n, s));
    AZ_CUDA(cudaMemsetAsync(ctr, 0, 4 * sizeof(unsigned long long), s));
    k_pack<<<(n + 255) / 256, 256, 0, s>>>(static_cast<const az_position *>(d_pos), n, st->frontier[0].as<PerftItem>());
    st->last_launches++;

    int cur = 0, remaining = depth;
    size_t items = (size_t)n;
    while (remaining > 1) {
        if (items >= kWide && remaining <= 4) break;
        // how many children would the next level have?
        AZ_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), s));
        int grid = (int)((items + 255) / 256);
        if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
        k_count<<<grid, 256, 0, s>>>(st->frontier[cur].as<PerftItem>(), (int)items, ctr);
        st->last_launches++;
        AZ_CUDA(cudaMemcpyAsync(st->h_counters, ctr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        AZ_CUDA(cudaStreamSynchronize(s));
        const size_t children = (size_t)st->h_counters[0];
        if (children == 0) { remaining = 0; items = 0; break; }       // every line is stuck: all zero
        if (children > kCapItems) {
            AZ_REQUIRE(remaining <= 4, AZ_ERR_CAPACITY,
                       "perft frontier would need %zu items (> %zu) with %d plies left", children, kCapItems, remaining);
            break;
        }
        AZ_REQUIRE(st->frontier[cur ^ 1].reserve(sizeof(PerftItem) * children) == 0, AZ_ERR_CUDA, "perft frontier alloc");
        AZ_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), s));
        k_expand<<<(int)((items + 255) / 256), 256, 0, s>>>(st->frontier[cur].as<PerftItem>(), (int)items,
                                                             st->frontier[cur ^ 1].as<PerftItem>(), ctr);
        st->last_launches++;
        cur ^= 1;
        items = children;
        remaining--;
    }
    st->last_frontier = (int)items;
    if (items > 0) {
        const PerftItem *f = st->frontier[cur].as<PerftItem>();
        AZ_REQUIRE(items <= 0x7fffffffULL, AZ_ERR_CAPACITY, "frontier too large");
        switch (remaining) {
            case 1: launch_walk<1>(ctx, f, (int)items, nodes, ctr + 1); break;
            case 2: launch_walk<2>(ctx, f, (int)items, nodes, ctr + 1); break;
            case 3: launch_walk<3>(ctx, f, (int)items, nodes, ctr + 1); break;
            case 4: launch_walk<4>(ctx, f, (int)items, nodes, ctr + 1); break;
            default: return az_fail(AZ_ERR_CAPACITY, "perft: %d plies left after expansion", remaining);
        }
        st->last_launches++;
    }
    AZ_CUDA(cudaMemcpyAsync(st->h_counters + 1, ctr + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    AZ_CUDA(cudaGetLastError());
    st->last_count_nodes = st->h_counters[1];
    ctx->launches += st->last_launches;
    return AZ_OK;
}

extern "C" int az_perft_batch(az_context *ctx, const az_position *pos, int n, int depth, uint64_t *nodes)
{
    AZ_REQUIRE(ctx && (n == 0 || (pos && nodes)), AZ_ERR_ARG, "az_perft_batch: null argument");
    AZ_REQUIRE(n >= 0, AZ_ERR_ARG, "az_perft_batch: n=%d", n);
    if (n == 0) return AZ_OK;
    AZ_REQUIRE(ctx->scratch[0].reserve(sizeof(az_position) * (size_t)n) == 0, AZ_ERR_CUDA, "scratch alloc");
    AZ_REQUIRE(ctx->scratch[1].reserve(sizeof(uint64_t) * (size_t)n) == 0, AZ_ERR_CUDA, "scratch alloc");
    AZ_CUDA(cudaMemcpyAsync(ctx->scratch[0].ptr, pos, sizeof(az_position) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = az_perft_batch_dev(ctx, ctx->scratch[0].ptr, n, depth, ctx->scratch[1].ptr);
    if (rc) return rc;
    AZ_CUDA(cudaMemcpyAsync(nodes, ctx->scratch[1].ptr, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    AZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return AZ_OK;
}

extern "C" int az_perft(az_context *ctx, const az_position *root, int depth, uint64_t *nodes)
{
    return az_perft_batch(ctx, root, 1, depth, nodes);
}

extern "C" int az_perft_last_stats(az_context *ctx, uint64_t *count_nodes, int32_t *launches, int32_t *frontier_items)
{
    AZ_REQUIRE(ctx && ctx->perft, AZ_ERR_STATE, "no perft call made on this context");
    if (count_nodes) *count_nodes = ctx->perft->last_count_nodes;
    if (launches) *launches = ctx->perft->last_launches;
    if (frontier_items) *frontier_items = ctx->perft->last_frontier;
    return AZ_OK;
}
