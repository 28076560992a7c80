// engine.cu — CUDA kernels (sm_100a) and the C ABI of include/ismcts_b200.h.
//
// Kernel map (reference symbols in SURVEY.md section 8a):
//   prepare_kernel<Game>                       a3 (set-up)   root record -> the register image iterations start from
//   search_kernel<Game, Slot, kind, occupancy> a1-a15        one lane = one SO tree / one set of MO trees
//   playout_kernel<Game>                       a3-a6,a8      determinise + rollout only
// No tensor cores are involved: the path has no dense contraction; it is
// bound by SM instruction issue.
#include "lane_search.cuh"
#include "kw_game.cuh"
#include "mnk_game.cuh"
#include "goof_game.cuh"
#include "phantom_game.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

namespace ismcts {

constexpr int kBlock = 128;
static_assert(kBlock == (int)kBlockLanes, "games that find their scratch from the thread index assume this block size");

// Shared memory of a block: the per-thread game scratch (hands / stones, one
// 8-byte column per thread) followed by the per-thread Philox rings (one 4-byte
// column per thread): column layouts are bank-conflict free.
template <class Game>
__host__ __device__ constexpr size_t block_smem()
{
    // (a game whose loops keep their Philox blocks in registers needs no ring: more of the SM's memory stays L1)
    return (size_t)kBlock * (Game::kHandWords * sizeof(uint64_t) + (game_is_phased<Game>(0) ? 0 : LaneRng::kRingWords) * sizeof(uint32_t))
         + game_block_extra_bytes<Game>(0);   // (per-block tables of the game, behind the scratch: Knockout Whist has no ring)
}

// Root records -> the register image every iteration starts from (one thread per position).
// (Games that find their scratch themselves — Knockout Whist — use the block's dynamic shared memory.)
template <class Game>
__global__ void prepare_kernel(const typename Game::Root* roots, typename Game::Prepared* prepared, uint32_t n,
                               uint64_t* start_ns)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && start_ns) *start_ns = global_timer_ns();    // time mode: the clock every lane's deadline counts from
    if (i >= n) return;
    uint64_t scratch[game_home_in_scratch<Game>(0) ? 1 : Game::kHandWords];
    Game::prepare(roots + i, prepared + i, scratch);
}

// kMinBlocks (resident CTAs per SM the register allocation must allow) trades registers per lane
// for warps per scheduler.
//
// Private trees (root parallelisation): lane = tree while the trees of the launch fit the resident lanes
// (P.unit_queue == nullptr: tree `gtid` is built in table `gtid` of the pool, which the host has zeroed).  With
// more trees than that the grid is PERSISTENT: one resident wave of blocks whose warps claim 32 trees at a time
// from an atomic counter until none are left, zeroing and reusing their 32 tables — a warp that finishes
// early takes the next trees instead of leaving its slot idle while the slowest warps of a single wave finish (a
// sixth of the warp slots in round 1, profiles/r02_persistent.md), and the pool no longer grows with the
// number of positions.  Either way a lane adds the root statistics of its tree to the sums of its position as
// soon as the tree is complete (RootParallel::compileVisitCounts, execution.h:219-231).
template <class Game, class Slot, uint32_t kKind, bool kShared, int kMinBlocks>
__global__ void __launch_bounds__(kBlock, kMinBlocks) search_kernel(SearchParams params)
{
    // the tree operations are real calls that reach the parameters through a pointer: one copy per block
    // in shared memory instead of one per thread in local memory
    __shared__ SearchParams P;
    if (threadIdx.x == 0) P = params;
    extern __shared__ uint64_t smem[];
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem + Game::kHandWords * kBlock) + threadIdx.x;
    Game::init_block(threadIdx.x, blockDim.x);
    __syncthreads();
    uint32_t path[Lane<Game, Slot, kKind, kShared>::kPathWords];   // (local memory: only the levels a descent reaches are touched)
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t units = P.n_positions * P.trees * (kShared ? P.lanes_per_tree : 1u);
    if constexpr (kShared) {
        lane_tree_search<Game, Slot, kKind, true>(P, gtid, gtid < units, smem, kBlock, threadIdx.x, ring, kBlock, path, gtid);
    } else {
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t* const queue = P.unit_queue;
        for (uint32_t pass = 0;; ++pass) {
            uint32_t base = gtid - lane;
            if (queue) {
                if (lane == 0) base = atomicAdd(queue, 32u);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
            } else if (pass) {
                break;
            }
            if (base >= units) break;
            if (queue) {
                // the 32 tables of this warp are consecutive: zeroed together, 512 bytes per store instruction
                uint4* const first = reinterpret_cast<uint4*>(reinterpret_cast<Slot*>(P.pool) + ((size_t)(gtid - lane) << P.log2cap));
                const size_t words = ((size_t)32 * sizeof(Slot) / sizeof(uint4)) << P.log2cap;
                for (size_t i = lane; i < words; i += 32) __stcs(first + i, make_uint4(0, 0, 0, 0));   // (streaming: not kept in L2)
                __syncwarp();
            }
            const uint32_t unit = base + lane;
            lane_tree_search<Game, Slot, kKind, false>(P, unit, unit < units, smem, kBlock, threadIdx.x, ring, kBlock, path, gtid);
            if (unit < units) collect_roots<Game, Slot, kKind>(P, unit, gtid);
            __syncwarp();
        }
    }
}

// Tree parallelisation: root node and node counter of every shared tree (one thread per tree).
template <class Slot>
__global__ void init_shared_roots_kernel(SearchParams P, uint32_t n_roots)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.n_positions * P.trees) init_shared_root<Slot>(P, i, n_roots);
}

// Shared trees: root statistics of every tree -> per-position sums (one thread per tree), after the search.
template <class Game, class Slot, uint32_t kKind>
__global__ void collect_roots_kernel(SearchParams P)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.n_positions * P.trees) collect_roots<Game, Slot, kKind>(P, i, i);
}

template <class Game>
__global__ void __launch_bounds__(kBlock) playout_kernel(SearchParams P, uint32_t per_position)
{
    extern __shared__ uint64_t smem[];
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem + Game::kHandWords * kBlock) + threadIdx.x;
    Game::init_block(threadIdx.x, blockDim.x);
    __syncthreads();
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    lane_playout<Game>(P, gtid, gtid < P.n_positions * per_position, per_position, smem, kBlock, threadIdx.x, ring, kBlock);
}

} // namespace ismcts

using namespace ismcts;

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------

struct ismcts_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string error;
    uint64_t launches = 0;

    // device buffers, grown on demand
    void* pool = nullptr;        size_t pool_bytes = 0;
    double* ln_table = nullptr;  uint32_t ln_count = 0;     // EXP3's ln(); entries of the two UCB1 tables
    UcbVisitFactors* visit_tab = nullptr; double* sqrt_ln_tab = nullptr;
    void* roots = nullptr;       size_t roots_bytes = 0;
    void* prepared = nullptr;    size_t prepared_bytes = 0;
    uint64_t* out = nullptr;     size_t out_bytes = 0;   // visits | score2 | avail | iters
    uint32_t* play_out = nullptr; size_t play_bytes = 0;
    uint64_t* start_ns = nullptr;                             // time mode: device clock at launch (prepare_kernel)
    uint32_t* queue = nullptr;                                // persistent grid: next tree to claim
    std::vector<std::pair<const void*, uint32_t>> resident;   // blocks of one resident wave, per search kernel

    int min_blocks = 8;          // resident CTAs per SM the SO/UCB1 Knockout Whist kernel is compiled for; ISMCTS_MIN_BLOCKS
    int min_blocks_mo = 0;       // ... the MO/EXP3 Knockout Whist kernel (0 = chosen per launch); ISMCTS_MIN_BLOCKS_MO

    // geometry of the last search (for ismcts_dump_tree)
    uint32_t last_log2cap = 0, last_lanes = 0 /* tables */, last_game = 0, last_solver = 0, last_policy = 0;
    std::vector<uint8_t> last_root_player; // per position
};

static thread_local std::string g_create_error;

static int fail(ismcts_ctx* ctx, int code, const std::string& what)
{
    if (ctx) ctx->error = what; else g_create_error = what;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                     \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(ctx, ISMCTS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

template <class T>
static int ensure(ismcts_ctx* ctx, T*& ptr, size_t& have, size_t need)
{
    if (need <= have) return ISMCTS_OK;
    if (ptr) { CUDA_TRY(ctx, cudaFree(ptr)); ptr = nullptr; have = 0; }
    CUDA_TRY(ctx, cudaMalloc((void**)&ptr, need));
    have = need;
    return ISMCTS_OK;
}

extern "C" {

const char* ismcts_version(void) { return "ismcsolver_b200 0.2 (sm_100a)"; }

size_t ismcts_state_size(uint32_t game)
{
    return game == ISMCTS_GAME_KW ? sizeof(ismcts_kw_state) : game == ISMCTS_GAME_MNK ? sizeof(ismcts_mnk_state)
         : game == ISMCTS_GAME_GOOF ? sizeof(ismcts_goof_state) : game == ISMCTS_GAME_PHANTOM ? sizeof(ismcts_phantom_state) : 0;
}

uint32_t ismcts_move_stride(uint32_t game)
{
    return game == ISMCTS_GAME_KW ? KwGame::kMoveStride : game == ISMCTS_GAME_MNK ? MnkGame::kMoveStride
         : game == ISMCTS_GAME_GOOF ? GoofGame::kMoveStride : game == ISMCTS_GAME_PHANTOM ? PhantomGame::kMoveStride : 0;
}

int ismcts_create(int device, ismcts_ctx** out)
{
    if (!out) return fail(nullptr, ISMCTS_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, ISMCTS_ERR_NO_DEVICE,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        " (this engine has no CPU fallback)");
    if (device < 0 || device >= count) return fail(nullptr, ISMCTS_ERR_INVALID, "device index out of range");
    ismcts_ctx* ctx = new (std::nothrow) ismcts_ctx;
    if (!ctx) return fail(nullptr, ISMCTS_ERR_INVALID, "out of host memory");
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, ISMCTS_ERR_NO_DEVICE, std::string("cannot initialise device: ") +
                                                      cudaGetErrorString(cudaGetLastError()));
    }
    ctx->stream = ctx->own_stream;
    if (cudaMalloc((void**)&ctx->start_ns, sizeof(uint64_t)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->queue, sizeof(uint32_t)) != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return fail(nullptr, ISMCTS_ERR_CUDA, "cannot allocate device memory");
    }
    if (const char* v = std::getenv("ISMCTS_MIN_BLOCKS")) ctx->min_blocks = std::atoi(v);
    if (const char* v = std::getenv("ISMCTS_MIN_BLOCKS_MO")) ctx->min_blocks_mo = std::atoi(v);
    *out = ctx;
    return ISMCTS_OK;
}

void ismcts_destroy(ismcts_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->pool); cudaFree(ctx->ln_table); cudaFree(ctx->visit_tab); cudaFree(ctx->sqrt_ln_tab); cudaFree(ctx->roots); cudaFree(ctx->prepared);
    cudaFree(ctx->out); cudaFree(ctx->play_out); cudaFree(ctx->start_ns); cudaFree(ctx->queue);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* ismcts_last_error(const ismcts_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int ismcts_set_stream(ismcts_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return ISMCTS_ERR_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return ISMCTS_OK;
}

uint64_t ismcts_launch_count(const ismcts_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ismcts_synchronize(ismcts_ctx* ctx)
{
    if (!ctx) return ISMCTS_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISMCTS_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------

static uint32_t ceil_log2(uint64_t x)
{
    uint32_t l = 0;
    while (((uint64_t)1 << l) < x) ++l;
    return l;
}

// Runs f(tag) with tag.game() of the device game class for a game kind.
template <class G> struct GameTag { using Game = G; };
template <class F>
static auto with_game(uint32_t game, F&& f)
{
    if (game == ISMCTS_GAME_KW) return f(GameTag<KwGame>{});
    if (game == ISMCTS_GAME_MNK) return f(GameTag<MnkGame>{});
    if (game == ISMCTS_GAME_PHANTOM) return f(GameTag<PhantomGame>{});
    return f(GameTag<GoofGame>{});
}

static size_t slot_bytes(uint32_t game, uint32_t policy)
{
    return with_game(game, [&](auto tag) {
        using Mask = typename decltype(tag)::Game::Mask;
        return policy == ISMCTS_POLICY_UCB1 ? sizeof(SlotUcb<Mask>) : sizeof(SlotExp<Mask>);
    });
}

static uint32_t trees_per_lane(uint32_t game, uint32_t solver)
{
    if (solver == ISMCTS_SOLVER_SO) return 1;
    return with_game(game, [](auto tag) { return (uint32_t) decltype(tag)::Game::kMaxPlayers; });
}

static int validate(ismcts_ctx* ctx, const ismcts_search_desc* d)
{
    if (!ctx || !d) return ISMCTS_ERR_INVALID;
    if (d->game > ISMCTS_GAME_PHANTOM) return fail(ctx, ISMCTS_ERR_INVALID, "unknown game");
    if (d->solver != ISMCTS_SOLVER_SO && d->solver != ISMCTS_SOLVER_MO) return fail(ctx, ISMCTS_ERR_INVALID, "unknown solver kind");
    if (d->policy != ISMCTS_POLICY_UCB1 && d->policy != ISMCTS_POLICY_EXP3) return fail(ctx, ISMCTS_ERR_INVALID, "unknown tree policy");
    if (d->flags & ISMCTS_FLAG_SHARED_TREE) {
        if ((d->flags & 0xF8u) != 0) return fail(ctx, ISMCTS_ERR_INVALID, "unknown flags");   // (lanes: bits 8..31)
        if ((d->flags >> ISMCTS_FLAG_LANES_SHIFT) == 0) return fail(ctx, ISMCTS_ERR_INVALID, "shared tree: lanes must be > 0");
    } else if ((d->flags & ~(ISMCTS_FLAG_KW_MAX4 | ISMCTS_FLAG_KEEP_TREES)) != 0) {
        return fail(ctx, ISMCTS_ERR_INVALID, "unknown flags");
    }
    if (d->n_positions == 0 || d->trees == 0) return fail(ctx, ISMCTS_ERR_INVALID, "n_positions and trees must be > 0");
    if ((uint64_t)d->n_positions * d->trees > 0x7FFFFFFFull) return fail(ctx, ISMCTS_ERR_INVALID, "too many trees");
    if (d->trees_total < d->tree_base + d->trees) return fail(ctx, ISMCTS_ERR_INVALID, "trees_total < tree_base + trees");
    if (d->iterations == 0 && d->time_ns == 0) return fail(ctx, ISMCTS_ERR_INVALID, "iterations or time_ns must be > 0");
    return ISMCTS_OK;
}

// Host records of root positions: everything the kernels index with must be in range (a malformed record
// would address the scratch of another lane).  ismcts_search_device takes device records and cannot look at
// them: its callers guarantee the same.
static int validate_roots(ismcts_ctx* ctx, uint32_t game, const void* roots, uint32_t n)
{
    const size_t state = ismcts_state_size(game);
    auto bad = [&](uint32_t p, const char* what) {
        return fail(ctx, ISMCTS_ERR_INVALID, "position " + std::to_string(p) + ": " + what);
    };
    for (uint32_t p = 0; p < n; ++p) {
        const uint8_t* rec = (const uint8_t*)roots + p * state;
        if (game == ISMCTS_GAME_KW) {
            const ismcts_kw_state* s = (const ismcts_kw_state*)rec;
            if (s->num_players < 2 || s->num_players > ISMCTS_KW_MAX_PLAYERS) return bad(p, "Knockout Whist: 2..7 players");
            const uint32_t all = (1u << s->num_players) - 1u;
            if (s->alive_mask == 0 || (s->alive_mask & ~all)) return bad(p, "Knockout Whist: alive_mask names no player or a player that does not exist");
            if (s->player >= s->num_players || s->dealer >= s->num_players) return bad(p, "Knockout Whist: player or dealer out of range");
            if (s->trump > 3 || s->tricks_left > 7 || s->request_trump > 1) return bad(p, "Knockout Whist: trump, tricks_left or request_trump out of range");
            if (s->trick_len >= 8 || s->trick_len > (uint32_t)__builtin_popcount(s->alive_mask)) return bad(p, "Knockout Whist: trick longer than the players still in");
            for (uint32_t i = 0; i < s->trick_len; ++i)
                if (s->trick_player[i] >= s->num_players || s->trick_card[i] >= ISMCTS_KW_CARDS) return bad(p, "Knockout Whist: trick entry out of range");
            for (uint32_t q = 0; q < ISMCTS_KW_MAX_PLAYERS; ++q) {
                // a hand has seven places (the game never deals more than seven cards)
                if (__builtin_popcountll(s->hands[q]) > 7) return bad(p, "Knockout Whist: a hand holds more than seven cards");
                if (s->hands[q] >> ISMCTS_KW_CARDS) return bad(p, "Knockout Whist: a hand holds a card that does not exist");
                if (q >= s->num_players && s->hands[q]) return bad(p, "Knockout Whist: a player that does not exist holds cards");
                if (s->tricks_taken[q] > 7) return bad(p, "Knockout Whist: more than seven tricks taken");
            }
            if (s->unknown >> ISMCTS_KW_CARDS) return bad(p, "Knockout Whist: an unknown card that does not exist");
        } else if (game == ISMCTS_GAME_MNK || game == ISMCTS_GAME_PHANTOM) {
            // (both records begin with stones[2][4]; m, n, k, player, result2 follow the boards)
            const uint8_t* f = game == ISMCTS_GAME_MNK ? &((const ismcts_mnk_state*)rec)->m : &((const ismcts_phantom_state*)rec)->m;
            const uint32_t m = f[0], nn = f[1], k = f[2], player = f[3];
            const int r2 = (int8_t)f[4];
            if (m == 0 || nn == 0 || m * nn > ISMCTS_MNK_MAX_CELLS) return bad(p, "m,n,k: the board must have 1..256 cells");
            if (k == 0 || player > 1 || r2 < -1 || r2 > 2) return bad(p, "m,n,k: k, player or result2 out of range");
        } else {
            const ismcts_goof_state* s = (const ismcts_goof_state*)rec;
            if (s->n_prizes > 13 || s->player > 2 || s->draw_prize > 1 || s->current_prize > 12) return bad(p, "Goofspiel: field out of range");
            if ((s->hands[0] | s->hands[1]) >> 13) return bad(p, "Goofspiel: a hand holds a rank that does not exist");
            if (s->moves[0] > 12 || s->moves[1] > 12) return bad(p, "Goofspiel: bid out of range");
        }
    }
    return ISMCTS_OK;
}

// The tables the bandits read (include/ismcts_detmath.h), built with the host libm — the oracle uses the same libm
// and the same expressions, so UCB values agree bit for bit; CUDA's log() does not.  ln(K), K < kLnSmall, for EXP3,
// once; the UCB1 factors {1 / v, 1 / sqrt(v)} and sqrt(ln(a)) for v, a < count, regrown on demand.
static int ensure_ln_table(ismcts_ctx* ctx, uint32_t count)
{
    if (!ctx->ln_table) {
        std::vector<double> host(kLnSmall);
        host[0] = 0.0;
        for (uint32_t i = 1; i < kLnSmall; ++i) host[i] = std::log((double)i);
        CUDA_TRY(ctx, cudaMalloc((void**)&ctx->ln_table, (size_t)kLnSmall * sizeof(double)));
        CUDA_TRY(ctx, cudaMemcpy(ctx->ln_table, host.data(), (size_t)kLnSmall * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (count <= ctx->ln_count) return ISMCTS_OK;
    count = std::max<uint32_t>(count, 4096);
    std::vector<UcbVisitFactors> by_visits(count);
    std::vector<double> sqrt_ln(count);
    for (uint32_t i = 0; i < count; ++i) {
        by_visits[i].inv = ismcts_tab_inv_visits(i);
        by_visits[i].inv_sqrt = ismcts_tab_inv_sqrt_visits(i);
        sqrt_ln[i] = ismcts_tab_sqrt_ln(i);
    }
    if (ctx->visit_tab) { CUDA_TRY(ctx, cudaFree(ctx->visit_tab)); ctx->visit_tab = nullptr; }
    if (ctx->sqrt_ln_tab) { CUDA_TRY(ctx, cudaFree(ctx->sqrt_ln_tab)); ctx->sqrt_ln_tab = nullptr; }
    ctx->ln_count = 0;
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->visit_tab, (size_t)count * sizeof(UcbVisitFactors)));
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->sqrt_ln_tab, (size_t)count * sizeof(double)));
    CUDA_TRY(ctx, cudaMemcpy(ctx->visit_tab, by_visits.data(), (size_t)count * sizeof(UcbVisitFactors), cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(ctx->sqrt_ln_tab, sqrt_ln.data(), (size_t)count * sizeof(double), cudaMemcpyHostToDevice));
    ctx->ln_count = count;
    return ISMCTS_OK;
}

// Runs prepare_kernel over `n` root records at d_roots into ctx->prepared.
template <class Game>
static int launch_prepare(ismcts_ctx* ctx, const void* d_roots, uint32_t n)
{
    int rc = ensure(ctx, ctx->prepared, ctx->prepared_bytes, (size_t)n * sizeof(typename Game::Prepared));
    if (rc) return rc;
    const size_t smem = game_home_in_scratch<Game>(0) ? (size_t)128 * Game::kHandWords * sizeof(uint64_t) : 0;
    prepare_kernel<Game><<<(n + 127) / 128, 128, smem, ctx->stream>>>((const typename Game::Root*)d_roots,
                                                                  (typename Game::Prepared*)ctx->prepared, n, ctx->start_ns);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return ISMCTS_OK;
}

// The kernels of one search: what the descriptor selects.
struct KernelSet {
    const void* search = nullptr;       // search_kernel<...>
    const void* collect = nullptr;      // shared trees: collect_roots_kernel<...>
    const void* init_shared = nullptr;  // shared trees: init_shared_roots_kernel<...>
    size_t smem = 0;                    // dynamic shared memory of a block of the search kernel
    uint32_t roots_per_table = 1;
};

template <class Game, class Slot, uint32_t kKind, bool kShared, int kMinBlocks, class RootGame = Game>
static KernelSet kernel_set()
{
    KernelSet k;
    k.search = (const void*)search_kernel<Game, Slot, kKind, kShared, kMinBlocks>;
    k.collect = (const void*)collect_roots_kernel<RootGame, Slot, kKind>;
    k.init_shared = (const void*)init_shared_roots_kernel<Slot>;
    k.smem = block_smem<Game>();
    k.roots_per_table = kKind == kMO ? Game::kMaxPlayers : 1u;
    return k;
}

template <class Game>
static KernelSet choose_kernels(const ismcts_ctx* ctx, const ismcts_search_desc* d)
{
    using Mask = typename Game::Mask;
    const bool so = d->solver == ISMCTS_SOLVER_SO, ucb = d->policy == ISMCTS_POLICY_UCB1;
    if (d->flags & ISMCTS_FLAG_SHARED_TREE) {
        // tree parallelisation: the lanes of a group share the tree (SO) or the trees of all players (MO)
        if constexpr (Game::kGameId == ISMCTS_GAME_KW) {
            // Knockout Whist, SO, UCB1 — root parallelisation with tree parallelisation inside (a block of lanes per
            // tree) is the engine's way of running the reference's own geometry (few deep trees) at full occupancy:
            // the same register budget and scratch layouts as the private-tree kernel
            if (so && ucb) {
                const bool max4 = (d->flags & ISMCTS_FLAG_KW_MAX4) != 0;
                if (ctx->min_blocks == 4) return max4 ? kernel_set<KwGame4, SlotUcb<Mask>, kSO, true, 4, Game>() : kernel_set<Game, SlotUcb<Mask>, kSO, true, 4>();
                return max4 ? kernel_set<KwGame4, SlotUcb<Mask>, kSO, true, 8, Game>() : kernel_set<Game, SlotUcb<Mask>, kSO, true, 8>();
            }
        }
        if (so && ucb) return kernel_set<Game, SlotUcb<Mask>, kSO, true, 4>();
        if (so) return kernel_set<Game, SlotExp<Mask>, kSO, true, 4>();
        if (ucb) return kernel_set<Game, SlotUcb<Mask>, kMO, true, 4>();
        return kernel_set<Game, SlotExp<Mask>, kMO, true, 4>();
    }
    if (so && ucb) {
        if constexpr (Game::kGameId == ISMCTS_GAME_KW) {
            // the headline path (Knockout Whist, SO, UCB1) exists at several register budgets, and for positions
            // with at most four players with less scratch per lane (the same prepared roots)
            const bool max4 = (d->flags & ISMCTS_FLAG_KW_MAX4) != 0;
            switch (ctx->min_blocks) {
            case 4: return max4 ? kernel_set<KwGame4, SlotUcb<Mask>, kSO, false, 4, Game>() : kernel_set<Game, SlotUcb<Mask>, kSO, false, 4>();
            case 5: return max4 ? kernel_set<KwGame4, SlotUcb<Mask>, kSO, false, 5, Game>() : kernel_set<Game, SlotUcb<Mask>, kSO, false, 5>();
            case 6: return max4 ? kernel_set<KwGame4, SlotUcb<Mask>, kSO, false, 6, Game>() : kernel_set<Game, SlotUcb<Mask>, kSO, false, 6>();
            default: return max4 ? kernel_set<KwGame4, SlotUcb<Mask>, kSO, false, 8, Game>() : kernel_set<Game, SlotUcb<Mask>, kSO, false, 8>();
            }
        } else {
            return kernel_set<Game, SlotUcb<Mask>, kSO, false, 4>();
        }
    }
    if (so) return kernel_set<Game, SlotExp<Mask>, kSO, false, 4>();
    if (ucb) return kernel_set<Game, SlotUcb<Mask>, kMO, false, 4>();
    if constexpr (Game::kGameId == ISMCTS_GAME_KW) {
        // BASELINE config 4 (Knockout Whist, MO-ISMCTS, EXP3) at several register budgets (ISMCTS_MIN_BLOCKS_MO)
        // (0 = by the size of the launch: with enough lanes to fill eight blocks per SM the 64-register kernel wins,
        // 4.4e8 against 4.0e8 it/s; with fewer lanes the registers are worth more than the occupancy nobody uses)
        const bool max4 = (d->flags & ISMCTS_FLAG_KW_MAX4) != 0;
        int blocks = ctx->min_blocks_mo;
        if (blocks == 0) blocks = (uint64_t)d->n_positions * d->trees >= 148u * 7u * kBlock ? 8 : 4;
        if (blocks == 6) return kernel_set<Game, SlotExp<Mask>, kMO, false, 6>();
        if (blocks == 8) return max4 ? kernel_set<KwGame4, SlotExp<Mask>, kMO, false, 8, Game>() : kernel_set<Game, SlotExp<Mask>, kMO, false, 8>();
    }
    return kernel_set<Game, SlotExp<Mask>, kMO, false, 4>();
}

// Blocks of `kernel` one resident wave holds (occupancy x SMs), asked of the runtime once per kernel.
static int resident_blocks(ismcts_ctx* ctx, const KernelSet& k, uint32_t* out)
{
    for (const auto& e : ctx->resident)
        if (e.first == k.search) { *out = e.second; return ISMCTS_OK; }
    int per_sm = 0, sms = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k.search, kBlock, k.smem));
    CUDA_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (per_sm < 1 || sms < 1) return fail(ctx, ISMCTS_ERR_CUDA, "the search kernel does not fit an SM");
    *out = (uint32_t)per_sm * (uint32_t)sms;
    ctx->resident.emplace_back(k.search, *out);
    return ISMCTS_OK;
}

static int launch(ismcts_ctx* ctx, const void* kernel, uint32_t grid, uint32_t block, size_t smem, void** args)
{
    CUDA_TRY(ctx, cudaLaunchKernel(kernel, dim3(grid), dim3(block), args, smem, ctx->stream));
    ctx->launches++;
    return ISMCTS_OK;
}

extern "C" int ismcts_search_device(ismcts_ctx* ctx, const ismcts_search_desc* d, const void* d_roots,
                                    uint64_t* d_visits, uint64_t* d_score2, uint64_t* d_avail, uint64_t* d_iters)
{
    int rc = validate(ctx, d);
    if (rc) return rc;
    if (!d_roots || !d_visits || !d_score2 || !d_avail || !d_iters) return fail(ctx, ISMCTS_ERR_INVALID, "null device pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));

    const bool shared = (d->flags & ISMCTS_FLAG_SHARED_TREE) != 0;
    const uint32_t width = shared ? d->flags >> ISMCTS_FLAG_LANES_SHIFT : 1u;
    const uint32_t trees = d->n_positions * d->trees;            // trees (MO: sets of trees) of this call
    if ((uint64_t)trees * width > 0x7FFFFFFFull) return fail(ctx, ISMCTS_ERR_INVALID, "too many lanes");
    const uint32_t lanes = trees * width;                        // threads with work
    const uint32_t stride = ismcts_move_stride(d->game);
    const uint32_t per_lane = trees_per_lane(d->game, d->solver);
    const size_t slot = slot_bytes(d->game, d->policy);
    const KernelSet kernels = with_game(d->game, [&](auto tag) { return choose_kernels<typename decltype(tag)::Game>(ctx, d); });

    // Private trees in count mode that do not fit one resident wave: a persistent grid whose warps claim trees and
    // reuse their tables (search_kernel).  The trees can then not be dumped afterwards; ISMCTS_FLAG_KEEP_TREES
    // (or a dump asked for in the descriptor) keeps one table per tree instead.
    uint32_t resident = 0;
    rc = resident_blocks(ctx, kernels, &resident);
    if (rc) return rc;
    const bool keep = (d->flags & ISMCTS_FLAG_KEEP_TREES) != 0 || d->dump_tree != 0xFFFFFFFFu;
    const bool recycle = !shared && d->iterations > 0 && !keep && lanes > resident * (uint32_t)kBlock;
    const uint32_t tables = recycle ? resident * (uint32_t)kBlock : trees;   // node pools

    // nodes per table: every iteration adds at most one node to each of its trees
    uint64_t max_it;
    uint32_t log2cap;
    if (d->iterations > 0) {
        max_it = d->iterations / d->trees_total + 1;
        log2cap = std::max<uint32_t>(4, ceil_log2(2 * (per_lane * (max_it + 1) + (shared ? (uint64_t)width * per_lane : 0))));
        if (log2cap > kMaxSlotsLog2) return fail(ctx, ISMCTS_ERR_CAPACITY, "too many iterations per tree for the node pool");
    } else {
        // time mode: the iteration count is unknown; give every lane an equal share of the free memory
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
        const size_t budget = (free_b + ctx->pool_bytes) / 2;
        log2cap = 4;
        while (log2cap < 21 && ((size_t)tables << (log2cap + 1)) * slot <= budget) ++log2cap;
        max_it = (((uint64_t)1 << (log2cap - 1)) - 1) / per_lane;
    }
    const size_t pool_need = ((size_t)tables << log2cap) * slot;
    if (pool_need > ctx->pool_bytes) {                    // (asking the driver costs milliseconds: only when the pool grows)
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
        if (pool_need > free_b + ctx->pool_bytes)
            return fail(ctx, ISMCTS_ERR_CAPACITY, "node pool does not fit device memory: need " + std::to_string(pool_need) + " bytes");
    }
    rc = ensure(ctx, ctx->pool, ctx->pool_bytes, pool_need);
    if (rc) return rc;
    // (shared trees: every lane in flight counts towards the availability of a node before its iteration is complete,
    // and max_nodes below must not stop a tree short of its budget)
    rc = ensure_ln_table(ctx, (uint32_t)std::min<uint64_t>(max_it + 3 + (shared ? (uint64_t)width + 1 : 0), 0x7FFFFFFFull));
    if (rc) return rc;

    if (!recycle) CUDA_TRY(ctx, cudaMemsetAsync(ctx->pool, 0, pool_need, ctx->stream));   // (recycled tables are zeroed by their warps)
    CUDA_TRY(ctx, cudaMemsetAsync(d_visits, 0, (size_t)d->n_positions * stride * sizeof(uint64_t), ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(d_score2, 0, (size_t)d->n_positions * stride * sizeof(uint64_t), ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(d_avail, 0, (size_t)d->n_positions * stride * sizeof(uint64_t), ctx->stream));
    if (recycle) CUDA_TRY(ctx, cudaMemsetAsync(ctx->queue, 0, sizeof(uint32_t), ctx->stream));
    if (shared) {
        CUDA_TRY(ctx, cudaMemsetAsync(d_iters, 0, (size_t)trees * sizeof(uint64_t), ctx->stream));
    }

    SearchParams P;
    std::memset(&P, 0, sizeof P);
    P.n_positions = d->n_positions; P.trees = d->trees; P.trees_total = d->trees_total;
    P.tree_base = d->tree_base; P.pos_base = d->pos_base;
    P.iterations = d->iterations; P.time_ns = d->time_ns; P.seed = d->seed;
    P.exploration = d->exploration < 0.0 ? 0.0 : d->exploration; // UCB1 ctor clamps, ucb1.h:65-67
    P.pool = ctx->pool; P.log2cap = log2cap;
    P.max_nodes = (uint32_t)std::min<uint64_t>(((uint64_t)1 << (log2cap - 1)), (uint64_t)per_lane * (ctx->ln_count - 2));
    P.ln_table = ctx->ln_table; P.ln_count = ctx->ln_count;
    P.visit_tab = ctx->visit_tab; P.sqrt_ln_tab = ctx->sqrt_ln_tab;
    P.visits = d_visits; P.score2 = d_score2; P.avail = d_avail; P.iters_done = d_iters;
    P.playout_out = nullptr;
    P.lanes_per_tree = width;
    P.start_ns = ctx->start_ns;
    P.unit_queue = recycle ? ctx->queue : nullptr;

    ctx->last_log2cap = log2cap; ctx->last_lanes = recycle ? 0 : trees; ctx->last_game = d->game;
    ctx->last_solver = d->solver; ctx->last_policy = d->policy;

    rc = with_game(d->game, [&](auto tag) { return launch_prepare<typename decltype(tag)::Game>(ctx, d_roots, P.n_positions); });
    if (rc) return rc;
    P.prepared = ctx->prepared;
    void* args[] = {&P};
    const uint32_t tgrid = (trees + 127) / 128;
    if (shared) {
        uint32_t n_roots = kernels.roots_per_table;
        void* init_args[] = {&P, &n_roots};
        rc = launch(ctx, kernels.init_shared, tgrid, 128, 0, init_args);
        if (rc) return rc;
    }
    rc = launch(ctx, kernels.search, recycle ? resident : (lanes + kBlock - 1) / kBlock, kBlock, kernels.smem, args);
    if (rc) return rc;
    if (shared) rc = launch(ctx, kernels.collect, tgrid, 128, 0, args);
    return rc;
}

// One lane's table -> node records of the tree rooted at slot `root`, in creation order.
template <class Slot>
static int dump_tree_t(ismcts_ctx* ctx, uint32_t lane, uint32_t root, ismcts_node_rec* nodes, uint32_t capacity,
                       uint32_t* count)
{
    const size_t cap = (size_t)1 << ctx->last_log2cap;
    std::vector<Slot> table(cap);
    CUDA_TRY(ctx, cudaMemcpy(table.data(), (const Slot*)ctx->pool + (size_t)lane * cap, cap * sizeof(Slot),
                             cudaMemcpyDeviceToHost));
    const uint32_t n_roots = trees_per_lane(ctx->last_game, ctx->last_solver);
    // the root slot of every node (follow the parent links), then the members of the wanted tree by creation index
    auto root_of = [&](size_t i) {
        while (i >= n_roots) i = slot_parent_of(table[i].key);
        return (uint32_t)i;
    };
    std::vector<std::pair<uint32_t, uint32_t>> members; // (created, slot)
    for (size_t i = 0; i < cap; ++i) {
        if (table[i].key == 0) continue;
        if (i < n_roots ? i == root : root_of(i) == root) members.emplace_back(i < n_roots ? 0u : table[i].created, (uint32_t)i);
    }
    std::sort(members.begin(), members.end());
    *count = (uint32_t)members.size();
    if (members.size() > capacity) return fail(ctx, ISMCTS_ERR_CAPACITY, "dump buffer too small");
    std::vector<uint32_t> index_of_slot(cap, 0);
    for (size_t k = 0; k < members.size(); ++k) index_of_slot[members[k].second] = (uint32_t)k;
    for (size_t k = 0; k < members.size(); ++k) {
        const Slot& s = table[members[k].second];
        ismcts_node_rec& r = nodes[k];
        if (s.key == kRootKey) { r.parent = 0; r.move = 0; }
        else { r.parent = index_of_slot[slot_parent_of(s.key)]; r.move = slot_move_of(s.key); }
        r.player = s.player; r.visits = s.visits;
        if constexpr (Slot::kExp3) { r.score2 = 0; r.avail = 1; r.score = s.score; r.prob = s.prob; }
        else { r.score2 = s.score2; r.avail = s.avail; r.score = 0.5 * (double)s.score2; r.prob = 1.0; }
    }
    return ISMCTS_OK;
}

extern "C" int ismcts_dump_tree_of(ismcts_ctx* ctx, uint32_t dump_index, uint32_t player, ismcts_node_rec* nodes,
                                   uint32_t capacity, uint32_t* count)
{
    if (!ctx || !nodes || !count) return ISMCTS_ERR_INVALID;
    if (ctx->last_lanes == 0) return fail(ctx, ISMCTS_ERR_INVALID, "the trees of the last search were not kept (ISMCTS_FLAG_KEEP_TREES)");
    if (dump_index >= ctx->last_lanes) return fail(ctx, ISMCTS_ERR_INVALID, "dump index out of range");
    const uint32_t n_roots = trees_per_lane(ctx->last_game, ctx->last_solver);
    if (ctx->last_solver == ISMCTS_SOLVER_SO) player = 0;
    if (player >= n_roots) return fail(ctx, ISMCTS_ERR_INVALID, "player out of range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const bool ucb = ctx->last_policy == ISMCTS_POLICY_UCB1;
    return with_game(ctx->last_game, [&](auto tag) {
        using Mask = typename decltype(tag)::Game::Mask;
        return ucb ? dump_tree_t<SlotUcb<Mask>>(ctx, dump_index, player, nodes, capacity, count)
                   : dump_tree_t<SlotExp<Mask>>(ctx, dump_index, player, nodes, capacity, count);
    });
}

extern "C" int ismcts_dump_tree(ismcts_ctx* ctx, uint32_t dump_index, ismcts_node_rec* nodes, uint32_t capacity,
                                uint32_t* count)
{
    if (!ctx) return ISMCTS_ERR_INVALID;
    uint32_t player = 0;
    if (ctx->last_solver == ISMCTS_SOLVER_MO && ctx->last_lanes) {
        // the tree that decides: the one of the root's player to move (mosolver.h:42-46)
        const size_t pos = dump_index / (ctx->last_lanes / std::max<size_t>(1, ctx->last_root_player.size()));
        if (pos < ctx->last_root_player.size()) player = ctx->last_root_player[pos];
    }
    return ismcts_dump_tree_of(ctx, dump_index, player, nodes, capacity, count);
}

extern "C" int ismcts_search(ismcts_ctx* ctx, const ismcts_search_desc* d, const void* roots, ismcts_search_out* out)
{
    int rc = validate(ctx, d);
    if (rc) return rc;
    if (!roots || !out || !out->visits) return fail(ctx, ISMCTS_ERR_INVALID, "null host pointer");
    rc = validate_roots(ctx, d->game, roots, d->n_positions);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint32_t stride = ismcts_move_stride(d->game);
    const size_t state = ismcts_state_size(d->game);
    const size_t root_bytes = (size_t)d->n_positions * state;
    const size_t per_arr = (size_t)d->n_positions * stride;
    const size_t n_lanes = (size_t)d->n_positions * d->trees;
    rc = ensure(ctx, ctx->roots, ctx->roots_bytes, root_bytes);
    if (rc) return rc;
    rc = ensure(ctx, ctx->out, ctx->out_bytes, (3 * per_arr + n_lanes) * sizeof(uint64_t));
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->roots, roots, root_bytes, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t* dv = ctx->out; uint64_t* ds = dv + per_arr; uint64_t* da = ds + per_arr; uint64_t* di = da + per_arr;
    ismcts_search_desc desc = *d;
    if (d->game == ISMCTS_GAME_KW) {
        // positions with at most four players use the search kernel with four seats of scratch per lane
        bool max4 = true;
        for (uint32_t p = 0; p < d->n_positions; ++p)
            max4 = max4 && ((const ismcts_kw_state*)((const uint8_t*)roots + p * state))->num_players <= 4;
        if (max4) desc.flags |= ISMCTS_FLAG_KW_MAX4; else desc.flags &= ~ISMCTS_FLAG_KW_MAX4;
    }
    rc = ismcts_search_device(ctx, &desc, ctx->roots, dv, ds, da, di);
    if (rc) return rc;
    // who moves at each root (for dumps of MO searches)
    ctx->last_root_player.resize(d->n_positions);
    for (uint32_t p = 0; p < d->n_positions; ++p) {
        const uint8_t* rec = (const uint8_t*)roots + p * state;
        ctx->last_root_player[p] = d->game == ISMCTS_GAME_KW ? ((const ismcts_kw_state*)rec)->player
                                 : d->game == ISMCTS_GAME_MNK ? ((const ismcts_mnk_state*)rec)->player
                                 : d->game == ISMCTS_GAME_PHANTOM ? ((const ismcts_phantom_state*)rec)->player
                                                                  : ((const ismcts_goof_state*)rec)->player;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(out->visits, dv, per_arr * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (out->score2) CUDA_TRY(ctx, cudaMemcpyAsync(out->score2, ds, per_arr * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (out->avail) CUDA_TRY(ctx, cudaMemcpyAsync(out->avail, da, per_arr * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (out->iterations_done) CUDA_TRY(ctx, cudaMemcpyAsync(out->iterations_done, di, n_lanes * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (out->best_move) {
        // RootParallel::bestMove, execution.h:209-216: first maximum in ascending move order
        for (uint32_t p = 0; p < d->n_positions; ++p) {
            uint32_t best = 0xFFFFFFFFu; uint64_t best_v = 0;
            for (uint32_t m = 0; m < stride; ++m)
                if (out->visits[(size_t)p * stride + m] > best_v) { best_v = out->visits[(size_t)p * stride + m]; best = m; }
            out->best_move[p] = best;
        }
    }
    out->dump_count = 0;
    if (d->dump_tree != 0xFFFFFFFFu && out->dump_nodes)
        return ismcts_dump_tree(ctx, d->dump_tree, out->dump_nodes, out->dump_capacity, &out->dump_count);
    return ISMCTS_OK;
}

// The trees of one request sharded over several devices (one context each): host threads launch the shards, the
// root arrays are summed on the host (RootParallel::compileVisitCounts, execution.h:219-231, across devices).  Tree
// ids stay global, so the result equals that of the same request on one device.
extern "C" int ismcts_search_multi(ismcts_ctx* const* ctxs, uint32_t n_ctx, const ismcts_search_desc* d, const void* roots,
                                   ismcts_search_out* out)
{
    if (!ctxs || n_ctx == 0 || !ctxs[0]) return ISMCTS_ERR_INVALID;
    ismcts_ctx* first = ctxs[0];
    int rc = validate(first, d);
    if (rc) return rc;
    if (!roots || !out || !out->visits) return fail(first, ISMCTS_ERR_INVALID, "null host pointer");
    if (d->dump_tree != 0xFFFFFFFFu) return fail(first, ISMCTS_ERR_INVALID, "ismcts_search_multi does not dump trees");
    for (uint32_t i = 0; i < n_ctx; ++i)
        if (!ctxs[i]) return fail(first, ISMCTS_ERR_INVALID, "null context");
    const uint32_t used = std::min<uint32_t>(n_ctx, d->trees);
    if (used == 1) return ismcts_search(first, d, roots, out);
    const uint32_t stride = ismcts_move_stride(d->game);
    const size_t per_arr = (size_t)d->n_positions * stride;
    struct Shard {
        ismcts_search_desc desc;
        std::vector<uint64_t> visits, score2, avail, iters;
        int rc = ISMCTS_OK;
    };
    std::vector<Shard> shards(used);
    uint32_t base = d->tree_base;
    for (uint32_t i = 0; i < used; ++i) {
        Shard& s = shards[i];
        s.desc = *d;
        s.desc.trees = d->trees / used + (i < d->trees % used ? 1u : 0u);
        s.desc.tree_base = base;
        base += s.desc.trees;
        s.visits.assign(per_arr, 0); s.score2.assign(per_arr, 0); s.avail.assign(per_arr, 0);
        s.iters.assign((size_t)d->n_positions * s.desc.trees, 0);
    }
    std::vector<std::thread> workers;
    for (uint32_t i = 0; i < used; ++i)
        workers.emplace_back([&, i] {
            Shard& s = shards[i];
            ismcts_search_out o;
            std::memset(&o, 0, sizeof o);
            o.visits = s.visits.data(); o.score2 = s.score2.data(); o.avail = s.avail.data(); o.iterations_done = s.iters.data();
            s.rc = ismcts_search(ctxs[i], &s.desc, roots, &o);
        });
    for (auto& w : workers) w.join();
    for (uint32_t i = 0; i < used; ++i)
        if (shards[i].rc) return fail(first, shards[i].rc, "device " + std::to_string(ctxs[i]->device) + ": " + ctxs[i]->error);
    for (size_t k = 0; k < per_arr; ++k) {
        uint64_t v = 0, s2 = 0, a = 0;
        for (const Shard& s : shards) { v += s.visits[k]; s2 += s.score2[k]; a += s.avail[k]; }
        out->visits[k] = v;
        if (out->score2) out->score2[k] = s2;
        if (out->avail) out->avail[k] = a;
    }
    if (out->iterations_done)
        for (uint32_t p = 0; p < d->n_positions; ++p) {
            uint64_t* dst = out->iterations_done + (size_t)p * d->trees;
            for (const Shard& s : shards) {
                std::copy(s.iters.begin() + (size_t)p * s.desc.trees, s.iters.begin() + (size_t)(p + 1) * s.desc.trees, dst);
                dst += s.desc.trees;
            }
        }
    if (out->best_move)
        for (uint32_t p = 0; p < d->n_positions; ++p) {
            uint32_t best = 0xFFFFFFFFu; uint64_t best_v = 0;
            for (uint32_t m = 0; m < stride; ++m)
                if (out->visits[(size_t)p * stride + m] > best_v) { best_v = out->visits[(size_t)p * stride + m]; best = m; }
            out->best_move[p] = best;
        }
    out->dump_count = 0;
    return ISMCTS_OK;
}

extern "C" int ismcts_playouts(ismcts_ctx* ctx, uint32_t game, const void* roots, uint32_t n_positions,
                               uint32_t per_position, uint64_t seed, uint32_t* out_result)
{
    if (!ctx || !roots || !out_result || n_positions == 0 || per_position == 0) return ISMCTS_ERR_INVALID;
    if (game > ISMCTS_GAME_PHANTOM) return fail(ctx, ISMCTS_ERR_INVALID, "unknown game");
    if (int bad = validate_roots(ctx, game, roots, n_positions)) return bad;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t root_bytes = (size_t)n_positions * ismcts_state_size(game);
    const size_t n = (size_t)n_positions * per_position;
    int rc = ensure(ctx, ctx->roots, ctx->roots_bytes, root_bytes);
    if (rc) return rc;
    rc = ensure(ctx, ctx->play_out, ctx->play_bytes, n * sizeof(uint32_t));
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->roots, roots, root_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t grid = (uint32_t)((n + kBlock - 1) / kBlock);
    SearchParams P;
    std::memset(&P, 0, sizeof P);
    P.n_positions = n_positions; P.seed = seed; P.playout_out = ctx->play_out;
    rc = with_game(game, [&](auto tag) { return launch_prepare<typename decltype(tag)::Game>(ctx, ctx->roots, n_positions); });
    if (rc) return rc;
    P.prepared = ctx->prepared;
    with_game(game, [&](auto tag) {
        using Game = typename decltype(tag)::Game;
        playout_kernel<Game><<<grid, kBlock, block_smem<Game>(), ctx->stream>>>(P, per_position);
        return 0;
    });
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(out_result, ctx->play_out, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISMCTS_OK;
}
