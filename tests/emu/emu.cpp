// emu.cpp — host build of the DEVICE search headers.  TEST INFRASTRUCTURE ONLY.
//
// The per-thread search body (ismcsolver_b200/csrc/*.cuh) is written as inline
// functions; without nvcc they compile as plain C++ so that the device logic can
// be compared with the CPU oracle on a machine without a GPU
// (tests/test_emu_vs_oracle.py).  The product library never loads this file and
// has no CPU path: the real kernels are in ismcsolver_b200/csrc/engine.cu.
//
// A warp is emulated by fibers (ucontext): every lane runs the unmodified
// lane_tree_search / lane_playout on its own stack until its next warp vote,
// the votes are combined when all lanes of the warp have arrived, and the lanes
// resume with the result — the lock-step loop of lane_search.cuh sees the same
// votes it sees on the GPU, in a fixed interleaving.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC
#include "../../ismcsolver_b200/csrc/lane_search.cuh"
#include "../../ismcsolver_b200/csrc/kw_game.cuh"
#include "../../ismcsolver_b200/csrc/mnk_game.cuh"
#include "../../ismcsolver_b200/csrc/goof_game.cuh"
#include "../../ismcsolver_b200/csrc/phantom_game.cuh"

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

using namespace ismcts;

namespace {

class FiberWarp {
public:
    explicit FiberWarp(size_t lanes) : fibers_(lanes) {}

    // The scratch of games that find it themselves (devdefs.cuh: block_scratch): lanes 128 k .. 128 k + 127
    // are the threads of the k-th "block" and use the k-th piece of `words_per_block` words.
    void set_scratch(uint64_t* base, size_t words_per_block) { scratch_ = base; words_per_warp_ = words_per_block; }

    // Runs body(lane) for every lane as one warp.
    void run(const std::function<void(uint32_t)>& body)
    {
        body_ = &body;
        active_ = this;
        g_host_warp_any = &FiberWarp::vote;
        for (size_t i = 0; i < fibers_.size(); ++i) {
            Fiber& f = fibers_[i];
            f.stack.resize(kStackBytes);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.data();
            f.ctx.uc_stack.ss_size = f.stack.size();
            f.ctx.uc_link = &main_;
            makecontext(&f.ctx, &FiberWarp::entry, 0);
        }
        for (;;) {
            bool alive = false, any = false;
            for (size_t i = 0; i < fibers_.size(); ++i) {
                Fiber& f = fibers_[i];
                if (f.finished) continue;
                current_ = i;
                g_host_block_scratch = scratch_ ? scratch_ + (i / kBlockLanes) * words_per_warp_ : nullptr;
                g_host_block_lane = (uint32_t)(i % kBlockLanes);
                swapcontext(&main_, &f.ctx);          // until its next vote, or its end
                if (!f.finished) { alive = true; any = any || f.ballot; }
            }
            if (!alive) break;
            result_ = any;
        }
        g_host_warp_any = nullptr;
        g_host_block_scratch = nullptr;
        active_ = nullptr;
    }

private:
    static constexpr size_t kStackBytes = 256 * 1024;
    struct Fiber {
        ucontext_t ctx;
        std::vector<char> stack;
        bool finished = false, ballot = false;
    };

    static bool vote(bool p)
    {
        FiberWarp* w = active_;
        Fiber& f = w->fibers_[w->current_];
        f.ballot = p;
        swapcontext(&f.ctx, &w->main_);
        return w->result_;
    }

    static void entry()
    {
        FiberWarp* w = active_;
        const size_t lane = w->current_;
        (*w->body_)((uint32_t)lane);
        w->fibers_[lane].finished = true;             // uc_link returns to main_
    }

    static FiberWarp* active_;
    std::vector<Fiber> fibers_;
    uint64_t* scratch_ = nullptr;
    size_t words_per_warp_ = 0;
    ucontext_t main_;
    size_t current_ = 0;
    bool result_ = false;
    const std::function<void(uint32_t)>* body_ = nullptr;
};

FiberWarp* FiberWarp::active_ = nullptr;

std::vector<double> ln_table(size_t n)
{
    std::vector<double> ln(std::max<size_t>(n, kLnSmall));
    ln[0] = 0.0;
    for (size_t i = 1; i < ln.size(); ++i) ln[i] = std::log((double)i);
    return ln;
}

// The UCB1 tables of the engine (engine.cu: ensure_ln_table), installed in P.
struct UcbTables {
    std::vector<UcbVisitFactors> by_visits;
    std::vector<double> sqrt_ln;
    explicit UcbTables(size_t n) : by_visits(n), sqrt_ln(n)
    {
        for (size_t i = 0; i < n; ++i) {
            by_visits[i].inv = ismcts_tab_inv_visits((uint32_t)i);
            by_visits[i].inv_sqrt = ismcts_tab_inv_sqrt_visits((uint32_t)i);
            sqrt_ln[i] = ismcts_tab_sqrt_ln((uint32_t)i);
        }
    }
    void install(SearchParams& P) const { P.visit_tab = by_visits.data(); P.sqrt_ln_tab = sqrt_ln.data(); P.ln_count = (uint32_t)sqrt_ln.size(); }
};

// One table -> the node records of the tree rooted at slot `root_slot`, in creation order.
template <class Slot>
int dump_table(const Slot* table, size_t cap, uint32_t n_roots, uint32_t root_slot, ismcts_node_rec* nodes, uint32_t capacity,
               uint32_t* count)
{
    auto root_of = [&](size_t i) { while (i >= n_roots) i = slot_parent_of(table[i].key); return (uint32_t)i; };
    std::vector<std::pair<uint32_t, uint32_t>> members;
    for (size_t i = 0; i < cap; ++i) {
        if (table[i].key == 0) continue;
        if (i < n_roots ? i == root_slot : root_of(i) == root_slot) members.emplace_back(i < n_roots ? 0u : table[i].created, (uint32_t)i);
    }
    std::sort(members.begin(), members.end());
    *count = (uint32_t)members.size();
    if (members.size() > capacity) return 5;
    std::vector<uint32_t> index_of_slot(cap, 0);
    for (size_t k = 0; k < members.size(); ++k) index_of_slot[members[k].second] = (uint32_t)k;
    for (size_t k = 0; k < members.size(); ++k) {
        const Slot& sl = table[members[k].second];
        ismcts_node_rec& r = nodes[k];
        if (sl.key == kRootKey) { r.parent = 0; r.move = 0; }
        else { r.parent = index_of_slot[slot_parent_of(sl.key)]; r.move = slot_move_of(sl.key); }
        r.player = sl.player; r.visits = sl.visits;
        if constexpr (Slot::kExp3) { r.score2 = 0; r.avail = 1; r.score = sl.score; r.prob = sl.prob; }
        else { r.score2 = sl.score2; r.avail = sl.avail; r.score = 0.5 * (double)sl.score2; r.prob = 1.0; }
    }
    return 0;
}

// Per-lane scratch of a warp of `width` lanes, laid out as in a CTA (one column per lane).
template <class Game>
struct WarpScratch {
    static constexpr size_t kWordsPerWarp = (size_t)Game::kHandWords * kBlockLanes;   // (per block of 128 lanes)
    std::vector<uint64_t> blocked;            // the scratch of the lanes, block of 128 lanes after block (as in the kernels)
    std::vector<uint32_t> ring, path;
    explicit WarpScratch(uint32_t w)
        : blocked(kWordsPerWarp * ((w + kBlockLanes - 1) / kBlockLanes)),
          ring((size_t)LaneRng::kRingWords * w), path((size_t)kPathWords * w) {}
    static constexpr size_t kPathWords = (size_t)Game::kMaxDepth * Game::kMaxPlayers;   // (enough for every kind of lane)
};

// Game::prepare with the scratch of one lane installed.
template <class Game>
void prepare_root(const void* root, typename Game::Prepared* prepared)
{
    std::vector<uint64_t> blocked(WarpScratch<Game>::kWordsPerWarp);
    uint64_t prep_scratch[Game::kHandWords];
    g_host_block_scratch = blocked.data();
    g_host_block_lane = 0;
    Game::prepare((const typename Game::Root*)root, prepared, prep_scratch);
    g_host_block_scratch = nullptr;
}

template <class Game>
uint32_t playout_t(const void* root, uint64_t seed, uint32_t pos, uint32_t iteration)
{
    SearchParams P;
    std::memset(&P, 0, sizeof P);
    uint32_t out = 0;
    // position `pos` of the caller is record 0 here: shift pos_base instead of the pointer
    typename Game::Prepared prepared;
    prepare_root<Game>(root, &prepared);
    P.prepared = &prepared; P.n_positions = 1; P.pos_base = pos; P.seed = seed; P.playout_out = &out - iteration;
    WarpScratch<Game> ws(1);
    // thread `iteration` of a launch with per_position > iteration plays playout `iteration` of record 0
    FiberWarp warp(1);
    warp.set_scratch(ws.blocked.data(), WarpScratch<Game>::kWordsPerWarp);
    warp.run([&](uint32_t) { lane_playout<Game>(P, iteration, true, 0x7FFFFFFFu, ws.blocked.data(), kBlockLanes, 0, ws.ring.data(), 1); });
    return out;
}

// `width` private trees (global ids first_tree ...) of one position searched by ONE warp of `width`
// lanes, `iterations` each; writes the tree of lane `dump` (MO: the tree of `player`).
template <class Game, class Slot, uint32_t kKind>
int search_t(const void* root, uint64_t iterations, uint64_t seed, uint32_t pos, uint32_t first_tree, uint32_t width,
             uint32_t dump, double exploration, uint32_t player, ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count)
{
    using L = Lane<Game, Slot, kKind>;
    uint32_t log2cap = 4;
    while (((uint64_t)1 << log2cap) < 2 * L::kTrees * (iterations + 2)) ++log2cap;
    const size_t cap = (size_t)1 << log2cap;
    std::vector<Slot> tables(cap * width);
    std::memset((void*)tables.data(), 0, tables.size() * sizeof(Slot));
    const std::vector<double> ln = ln_table(iterations + 300);
    const uint32_t stride = Game::kMoveStride;
    std::vector<uint64_t> v(stride), s(stride), a(stride), it(width);

    SearchParams P;
    std::memset(&P, 0, sizeof P);
    typename Game::Prepared prepared;
    prepare_root<Game>(root, &prepared);
    P.prepared = &prepared; P.n_positions = 1; P.trees = width; P.trees_total = first_tree + width; P.tree_base = first_tree;
    P.pos_base = pos; P.iterations = iterations * P.trees_total; P.seed = seed; P.exploration = exploration;
    P.pool = tables.data(); P.log2cap = log2cap; P.max_nodes = 1u << (log2cap - 1);
    P.ln_table = ln.data();
    const UcbTables ucb(iterations + 300);
    ucb.install(P);
    P.visits = v.data(); P.score2 = s.data(); P.avail = a.data(); P.iters_done = it.data();
    P.lanes_per_tree = 1;
    WarpScratch<Game> ws(width);
    FiberWarp warp(width);
    warp.set_scratch(ws.blocked.data(), WarpScratch<Game>::kWordsPerWarp);
    warp.run([&](uint32_t lane) {
        lane_tree_search<Game, Slot, kKind, false>(P, lane, true, ws.blocked.data() + (lane / kBlockLanes) * WarpScratch<Game>::kWordsPerWarp,
                                                   kBlockLanes, lane % kBlockLanes, ws.ring.data() + lane, width,
                                                   ws.path.data() + (size_t)lane * WarpScratch<Game>::kPathWords);
    });
    for (uint32_t i = 0; i < width; ++i) if (it[i] != iterations) return 6;
    return dump_table(tables.data() + cap * dump, cap, L::kTrees, kKind == kMO ? player : 0, nodes, capacity, count);
}

template <class Game>
int search_game(const void* root, uint32_t solver, uint32_t policy, uint64_t iterations, uint64_t seed, uint32_t pos,
                uint32_t tree, uint32_t width, uint32_t dump, double exploration, uint32_t player, ismcts_node_rec* nodes,
                uint32_t capacity, uint32_t* count)
{
    using Mask = typename Game::Mask;
    if (solver == ISMCTS_SOLVER_SO && policy == ISMCTS_POLICY_UCB1)
        return search_t<Game, SlotUcb<Mask>, kSO>(root, iterations, seed, pos, tree, width, dump, exploration, player, nodes, capacity, count);
    if (solver == ISMCTS_SOLVER_SO)
        return search_t<Game, SlotExp<Mask>, kSO>(root, iterations, seed, pos, tree, width, dump, exploration, player, nodes, capacity, count);
    if (policy == ISMCTS_POLICY_UCB1)
        return search_t<Game, SlotUcb<Mask>, kMO>(root, iterations, seed, pos, tree, width, dump, exploration, player, nodes, capacity, count);
    return search_t<Game, SlotExp<Mask>, kMO>(root, iterations, seed, pos, tree, width, dump, exploration, player, nodes, capacity, count);
}

// Tree parallelisation on the host: the `width` lanes of one warp share ONE tree (MO-ISMCTS: the
// trees of all players); on the GPU the interleaving of the lanes is whatever the hardware does, here
// it is the fibers' fixed one.  Writes the tree of `player` (SO: the tree).
template <class Game, class Slot, uint32_t kKind>
int shared_search_t(const void* root, uint64_t iterations, uint32_t width, uint64_t seed, uint32_t player,
                    ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count, uint64_t* done_out)
{
    constexpr uint32_t n_roots = kKind == kMO ? Game::kMaxPlayers : 1;
    uint32_t log2cap = 4;
    while (((uint64_t)1 << log2cap) < 2 * n_roots * (iterations + 2 + width)) ++log2cap;
    std::vector<Slot> table((size_t)1 << log2cap);
    std::memset((void*)table.data(), 0, table.size() * sizeof(Slot));
    const std::vector<double> ln = ln_table(iterations + 300);
    std::vector<uint64_t> v(Game::kMoveStride), s2(Game::kMoveStride), a(Game::kMoveStride), it(1);
    typename Game::Prepared prepared;
    prepare_root<Game>(root, &prepared);
    SearchParams P;
    std::memset(&P, 0, sizeof P);
    P.prepared = &prepared; P.n_positions = 1; P.trees = 1; P.trees_total = 1; P.iterations = iterations;
    P.seed = seed; P.exploration = 0.7; P.pool = table.data(); P.log2cap = log2cap; P.max_nodes = 1u << (log2cap - 1);
    P.ln_table = ln.data();
    const UcbTables ucb(iterations + 300);
    ucb.install(P);
    P.visits = v.data(); P.score2 = s2.data(); P.avail = a.data(); P.iters_done = it.data();
    P.lanes_per_tree = width;
    init_shared_root<Slot>(P, 0, n_roots);
    WarpScratch<Game> ws(width);
    FiberWarp warp(width);
    warp.set_scratch(ws.blocked.data(), WarpScratch<Game>::kWordsPerWarp);
    warp.run([&](uint32_t lane) {
        lane_tree_search<Game, Slot, kKind, true>(P, lane, true, ws.blocked.data() + (lane / kBlockLanes) * WarpScratch<Game>::kWordsPerWarp,
                                                  kBlockLanes, lane % kBlockLanes, ws.ring.data() + lane, width,
                                                  ws.path.data() + (size_t)lane * WarpScratch<Game>::kPathWords);
    });
    *done_out = it[0];
    return dump_table(table.data(), table.size(), n_roots, kKind == kMO ? player : 0, nodes, capacity, count);
}

template <class Game>
int shared_search_game(const void* root, uint32_t solver, uint32_t policy, uint64_t iterations, uint32_t width, uint64_t seed,
                       uint32_t player, ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count, uint64_t* done)
{
    using Mask = typename Game::Mask;
    if (solver == ISMCTS_SOLVER_SO && policy == ISMCTS_POLICY_UCB1)
        return shared_search_t<Game, SlotUcb<Mask>, kSO>(root, iterations, width, seed, player, nodes, capacity, count, done);
    if (solver == ISMCTS_SOLVER_SO)
        return shared_search_t<Game, SlotExp<Mask>, kSO>(root, iterations, width, seed, player, nodes, capacity, count, done);
    if (policy == ISMCTS_POLICY_UCB1)
        return shared_search_t<Game, SlotUcb<Mask>, kMO>(root, iterations, width, seed, player, nodes, capacity, count, done);
    return shared_search_t<Game, SlotExp<Mask>, kMO>(root, iterations, width, seed, player, nodes, capacity, count, done);
}

} // namespace

extern "C" {

int emu_shared_search(uint32_t game, const void* root, uint32_t solver, uint32_t policy, uint64_t iterations, uint32_t width,
                      uint64_t seed, uint32_t player, ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count, uint64_t* done)
{
    if (game == ISMCTS_GAME_KW)
        return shared_search_game<KwGame>(root, solver, policy, iterations, width, seed, player, nodes, capacity, count, done);
    if (game == ISMCTS_GAME_GOOF)
        return shared_search_game<GoofGame>(root, solver, policy, iterations, width, seed, player, nodes, capacity, count, done);
    if (game == ISMCTS_GAME_PHANTOM)
        return shared_search_game<PhantomGame>(root, solver, policy, iterations, width, seed, player, nodes, capacity, count, done);
    return shared_search_game<MnkGame>(root, solver, policy, iterations, width, seed, player, nodes, capacity, count, done);
}

uint32_t emu_playout(uint32_t game, const void* root, uint64_t seed, uint32_t pos, uint32_t iteration)
{
    if (game == ISMCTS_GAME_KW) return playout_t<KwGame>(root, seed, pos, iteration);
    if (game == ISMCTS_GAME_GOOF) return playout_t<GoofGame>(root, seed, pos, iteration);
    if (game == ISMCTS_GAME_PHANTOM) return playout_t<PhantomGame>(root, seed, pos, iteration);
    return playout_t<MnkGame>(root, seed, pos, iteration);
}

// Any solver / tree policy.  A warp of `width` lanes searches the trees tree ... tree + width - 1 of
// position `pos`, `iterations` each; the tree of lane `dump` is written (MO: the tree of `player`).
int emu_search(uint32_t game, const void* root, uint32_t solver, uint32_t policy, uint64_t iterations, uint64_t seed,
               uint32_t pos, uint32_t tree, uint32_t width, uint32_t dump, double exploration, uint32_t player,
               ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count)
{
    if (width == 0 || dump >= width) return 1;
    if (game == ISMCTS_GAME_KW)
        return search_game<KwGame>(root, solver, policy, iterations, seed, pos, tree, width, dump, exploration, player, nodes, capacity, count);
    if (game == ISMCTS_GAME_GOOF)
        return search_game<GoofGame>(root, solver, policy, iterations, seed, pos, tree, width, dump, exploration, player, nodes, capacity, count);
    if (game == ISMCTS_GAME_PHANTOM)
        return search_game<PhantomGame>(root, solver, policy, iterations, seed, pos, tree, width, dump, exploration, player, nodes, capacity, count);
    return search_game<MnkGame>(root, solver, policy, iterations, seed, pos, tree, width, dump, exploration, player, nodes, capacity, count);
}

} // extern "C"
