// emu.cpp — host build of the DEVICE search headers.  TEST INFRASTRUCTURE ONLY.
//
// The per-thread search body (ismcsolver_b200/csrc/*.cuh) is written as inline
// functions; with ISMCTS_HOST_EMU they compile as plain C++ so that the device
// logic can be compared with the CPU oracle on a machine without a GPU
// (tests/test_emu_vs_oracle.py).  The product library never loads this file and
// has no CPU path: the real kernels are in ismcsolver_b200/csrc/engine.cu.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC
#include "../../ismcsolver_b200/csrc/lane_search.cuh"
#include "../../ismcsolver_b200/csrc/kw_game.cuh"
#include "../../ismcsolver_b200/csrc/mnk_game.cuh"
#include "../../ismcsolver_b200/csrc/goof_game.cuh"
#include "../../ismcsolver_b200/csrc/phantom_game.cuh"

#include <chrono>
#include <cmath>
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

using namespace ismcts;

static uint32_t g_thresh = 8, g_period = 8, g_mode = 6, g_split = 8;

template <class Game>
static uint32_t playout_t(const void* root, uint64_t seed, uint32_t pos, uint32_t iteration)
{
    SearchParams P;
    std::memset(&P, 0, sizeof P);
    uint32_t out = 0;
    // position `pos` of the caller is record 0 here: shift pos_base instead of the pointer
    typename Game::Prepared prepared;
    uint64_t prep_scratch[Game::kHandWords];
    Game::prepare((const typename Game::Root*)root, &prepared, prep_scratch);
    P.prepared = &prepared; P.n_positions = 1; P.pos_base = pos; P.seed = seed; P.playout_out = &out - iteration;
    P.tree_thresh = g_thresh; P.tree_period = g_period; P.sched_mode = g_mode; P.sched_split_min = g_split;
    uint64_t scratch[Game::kHandWords];
    uint32_t ring[LaneRng::kRingWords];
    // thread `iteration` of a launch with per_position > iteration plays playout `iteration` of record 0
    if (g_mode == 6) lane_playout<Game, true>(P, iteration, true, 0x7FFFFFFFu, scratch, 1, ring, 1);
    else lane_playout<Game, false>(P, iteration, true, 0x7FFFFFFFu, scratch, 1, ring, 1);
    return out;
}

// One lane (one SO tree or one set of MO trees); writes the tree of `player` in creation order.
template <class Game, class Slot, uint32_t kKind>
static int search_t(const void* root, uint64_t iterations, uint64_t seed, uint32_t pos, uint32_t tree_id,
                    double exploration, uint32_t player, ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count)
{
    using L = Lane<Game, Slot, kKind>;
    uint32_t log2cap = 4;
    while (((uint64_t)1 << log2cap) < 2 * L::kTrees * (iterations + 2)) ++log2cap;
    std::vector<Slot> table((size_t)1 << log2cap);
    std::memset((void*)table.data(), 0, table.size() * sizeof(Slot));
    std::vector<double> ln(iterations + 300);
    ln[0] = 0.0;
    for (size_t i = 1; i < ln.size(); ++i) ln[i] = std::log((double)i);
    const uint32_t stride = Game::kMoveStride;
    std::vector<uint64_t> v(stride), s(stride), a(stride), it(1);

    SearchParams P;
    std::memset(&P, 0, sizeof P);
    typename Game::Prepared prepared;
    uint64_t prep_scratch[Game::kHandWords];
    Game::prepare((const typename Game::Root*)root, &prepared, prep_scratch);
    P.prepared = &prepared; P.n_positions = 1; P.trees = 1; P.trees_total = tree_id + 1; P.tree_base = tree_id;
    P.pos_base = pos; P.iterations = iterations * (tree_id + 1); P.seed = seed; P.exploration = exploration;
    P.pool = table.data(); P.log2cap = log2cap; P.max_nodes = 1u << (log2cap - 1);
    P.ln_table = ln.data(); P.ln_count = (uint32_t)ln.size();
    P.visits = v.data(); P.score2 = s.data(); P.avail = a.data(); P.iters_done = it.data();
    P.tree_thresh = g_thresh; P.tree_period = g_period; P.sched_mode = g_mode; P.sched_split_min = g_split;
    uint64_t scratch[Game::kHandWords];
    uint32_t ring[LaneRng::kRingWords];
    std::vector<uint32_t> path(Game::kMaxDepth);
    if (g_mode == 6) lane_tree_search<Game, Slot, kKind, false, true>(P, 0, true, scratch, 1, ring, 1, path.data());
    else lane_tree_search<Game, Slot, kKind, false, false>(P, 0, true, scratch, 1, ring, 1, path.data());

    const uint32_t n_roots = L::kTrees;
    const uint32_t root_slot = kKind == kMO ? player : 0;
    auto root_of = [&](size_t i) { while (i >= n_roots) i = slot_parent_of(table[i].key); return (uint32_t)i; };
    std::vector<std::pair<uint32_t, uint32_t>> members;
    for (size_t i = 0; i < table.size(); ++i) {
        if (table[i].key == 0) continue;
        if (i < n_roots ? i == root_slot : root_of(i) == root_slot) members.emplace_back(i < n_roots ? 0u : table[i].created, (uint32_t)i);
    }
    std::sort(members.begin(), members.end());
    *count = (uint32_t)members.size();
    if (members.size() > capacity) return 5;
    std::vector<uint32_t> index_of_slot(table.size(), 0);
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

template <class Game>
static int search_game(const void* root, uint32_t solver, uint32_t policy, uint64_t iterations, uint64_t seed, uint32_t pos,
                       uint32_t tree, double exploration, uint32_t player, ismcts_node_rec* nodes, uint32_t capacity,
                       uint32_t* count)
{
    using Mask = typename Game::Mask;
    if (solver == ISMCTS_SOLVER_SO && policy == ISMCTS_POLICY_UCB1)
        return search_t<Game, SlotUcb<Mask>, kSO>(root, iterations, seed, pos, tree, exploration, player, nodes, capacity, count);
    if (solver == ISMCTS_SOLVER_SO)
        return search_t<Game, SlotExp<Mask>, kSO>(root, iterations, seed, pos, tree, exploration, player, nodes, capacity, count);
    if (policy == ISMCTS_POLICY_UCB1)
        return search_t<Game, SlotUcb<Mask>, kMO>(root, iterations, seed, pos, tree, exploration, player, nodes, capacity, count);
    return search_t<Game, SlotExp<Mask>, kMO>(root, iterations, seed, pos, tree, exploration, player, nodes, capacity, count);
}

// Tree parallelisation on the host: `width` lanes share ONE tree and are advanced pass by pass in a
// fixed interleaving (on the GPU the interleaving is whatever the hardware does).  Writes the tree
// in creation order.
template <class Game>
static int shared_search_t(const void* root, uint64_t iterations, uint32_t width, uint64_t seed, ismcts_node_rec* nodes,
                           uint32_t capacity, uint32_t* count, uint64_t* done_out)
{
    using Slot = SlotUcb<typename Game::Mask>;
    using L = Lane<Game, Slot, kSO, true>;
    uint32_t log2cap = 4;
    while (((uint64_t)1 << log2cap) < 2 * (iterations + 2 + width)) ++log2cap;
    std::vector<Slot> table((size_t)1 << log2cap);
    std::memset((void*)table.data(), 0, table.size() * sizeof(Slot));
    std::vector<double> ln(iterations + 300);
    ln[0] = 0.0;
    for (size_t i = 1; i < ln.size(); ++i) ln[i] = std::log((double)i);
    std::vector<uint64_t> v(Game::kMoveStride), s2(Game::kMoveStride), a(Game::kMoveStride), it(1);
    uint32_t counter = 0;
    typename Game::Prepared prepared;
    uint64_t prep_scratch[Game::kHandWords];
    Game::prepare((const typename Game::Root*)root, &prepared, prep_scratch);
    SearchParams P;
    std::memset(&P, 0, sizeof P);
    P.prepared = &prepared; P.n_positions = 1; P.trees = 1; P.trees_total = 1; P.iterations = iterations;
    P.seed = seed; P.exploration = 0.7; P.pool = table.data(); P.log2cap = log2cap; P.max_nodes = 1u << (log2cap - 1);
    P.ln_table = ln.data(); P.ln_count = (uint32_t)ln.size();
    P.visits = v.data(); P.score2 = s2.data(); P.avail = a.data(); P.iters_done = it.data();
    P.tree_thresh = g_thresh; P.tree_period = g_period; P.sched_mode = g_mode; P.sched_split_min = g_split;
    P.lanes_per_tree = width; P.node_counters = &counter;
    init_shared_root<Slot>(P, 0);
    std::vector<std::unique_ptr<L>> lanes;
    std::vector<std::vector<uint64_t>> scratch(width, std::vector<uint64_t>(Game::kHandWords));
    std::vector<std::vector<uint32_t>> ring(width, std::vector<uint32_t>(LaneRng::kRingWords)), path(width, std::vector<uint32_t>(Game::kMaxDepth));
    for (uint32_t i = 0; i < width; ++i) {
        lanes.emplace_back(new L(P));
        L& l = *lanes.back();
        l.path = path[i].data();
        l.g.bind(scratch[i].data(), 1);
        l.phase = kDone; l.done = 0; l.deadline = 0; l.root_children = Game::Mask::none();
        l.prep = &prepared;
        l.tree.bind(table.data(), log2cap, 1);
        l.tree_players = 1;
        l.rng.init(seed, 0, 0, ring[i].data(), 1);
        l.quota = (uint32_t)(iterations / width + (i < iterations % width ? 1 : 0));
        l.first_iteration = i; l.iteration_stride = width; l.node_counter = &counter;
        if (l.budget_left()) l.start_iteration();
    }
    Schedule sched;
    sched.init(P);
    for (uint32_t pass = 0;; ++pass) {
        std::vector<typename L::Wants> w(width);
        uint32_t nc = 0, nm = 0, nt = 0, alive = 0;
        for (uint32_t i = 0; i < width; ++i) {
            alive += lanes[i]->phase != kDone;
            w[i] = lanes[i]->classify();
            nc += w[i].chance; nm += w[i].move; nt += w[i].tree;
        }
        if (!alive) break;
        const PassPlan plan = sched.plan_pass(nc, nm, nt, pass);
        for (uint32_t i = 0; i < width; ++i) lanes[i]->act(w[i], plan, pass, i);
    }
    *done_out = 0;
    for (uint32_t i = 0; i < width; ++i) *done_out += lanes[i]->done;
    // dump in creation order
    std::vector<std::pair<uint32_t, uint32_t>> members;
    for (size_t i = 0; i < table.size(); ++i) if (table[i].key != 0) members.emplace_back(i == 0 ? 0u : table[i].created, (uint32_t)i);
    std::sort(members.begin(), members.end());
    *count = (uint32_t)members.size();
    if (members.size() > capacity) return 5;
    std::vector<uint32_t> index_of_slot(table.size(), 0);
    for (size_t k = 0; k < members.size(); ++k) index_of_slot[members[k].second] = (uint32_t)k;
    for (size_t k = 0; k < members.size(); ++k) {
        const Slot& sl = table[members[k].second];
        ismcts_node_rec& r = nodes[k];
        if (sl.key == kRootKey) { r.parent = 0; r.move = 0; }
        else { r.parent = index_of_slot[slot_parent_of(sl.key)]; r.move = slot_move_of(sl.key); }
        r.player = sl.player; r.visits = sl.visits; r.score2 = sl.score2; r.avail = sl.avail;
        r.score = 0.5 * (double)sl.score2; r.prob = 1.0;
    }
    return 0;
}

extern "C" {

// Schedule of the search loop (when pending tree operations run); results must not depend on it.
void emu_set_schedule(uint32_t thresh, uint32_t period) { g_thresh = thresh; g_period = period; }
void emu_set_schedule_mode(uint32_t mode, uint32_t split_min) { g_mode = mode; g_split = split_min; }

int emu_shared_search(uint32_t game, const void* root, uint64_t iterations, uint32_t width, uint64_t seed,
                      ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count, uint64_t* done)
{
    if (game == ISMCTS_GAME_KW) return shared_search_t<KwGame>(root, iterations, width, seed, nodes, capacity, count, done);
    return shared_search_t<MnkGame>(root, iterations, width, seed, nodes, capacity, count, done);
}

uint32_t emu_playout(uint32_t game, const void* root, uint64_t seed, uint32_t pos, uint32_t iteration)
{
    if (game == ISMCTS_GAME_KW) return playout_t<KwGame>(root, seed, pos, iteration);
    if (game == ISMCTS_GAME_GOOF) return playout_t<GoofGame>(root, seed, pos, iteration);
    if (game == ISMCTS_GAME_PHANTOM) return playout_t<PhantomGame>(root, seed, pos, iteration);
    return playout_t<MnkGame>(root, seed, pos, iteration);
}

int emu_so_search(uint32_t game, const void* root, uint64_t iterations, uint64_t seed, uint32_t pos, uint32_t tree,
                  double exploration, ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count)
{
    if (game == ISMCTS_GAME_KW)
        return search_game<KwGame>(root, ISMCTS_SOLVER_SO, ISMCTS_POLICY_UCB1, iterations, seed, pos, tree, exploration, 0, nodes, capacity, count);
    return search_game<MnkGame>(root, ISMCTS_SOLVER_SO, ISMCTS_POLICY_UCB1, iterations, seed, pos, tree, exploration, 0, nodes, capacity, count);
}

// Any solver / tree policy; MO: the tree of `player`.
int emu_search(uint32_t game, const void* root, uint32_t solver, uint32_t policy, uint64_t iterations, uint64_t seed,
               uint32_t pos, uint32_t tree, double exploration, uint32_t player, ismcts_node_rec* nodes, uint32_t capacity,
               uint32_t* count)
{
    if (game == ISMCTS_GAME_KW)
        return search_game<KwGame>(root, solver, policy, iterations, seed, pos, tree, exploration, player, nodes, capacity, count);
    if (game == ISMCTS_GAME_GOOF)
        return search_game<GoofGame>(root, solver, policy, iterations, seed, pos, tree, exploration, player, nodes, capacity, count);
    if (game == ISMCTS_GAME_PHANTOM)
        return search_game<PhantomGame>(root, solver, policy, iterations, seed, pos, tree, exploration, player, nodes, capacity, count);
    return search_game<MnkGame>(root, solver, policy, iterations, seed, pos, tree, exploration, player, nodes, capacity, count);
}

} // extern "C"
