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

#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

using namespace ismcts;

static uint32_t g_thresh = 8, g_period = 8;

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
    P.tree_thresh = g_thresh; P.tree_period = g_period;
    uint64_t scratch[Game::kHandWords];
    uint32_t ring[LaneRng::kRingWords];
    // thread `iteration` of a launch with per_position > iteration plays playout `iteration` of record 0
    lane_playout<Game>(P, iteration, true, 0x7FFFFFFFu, scratch, 1, ring, 1);
    return out;
}

template <class Game>
static int so_search_t(const void* root, uint64_t iterations, uint64_t seed, uint32_t pos, uint32_t tree_id,
                       double exploration, ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count)
{
    using Mask = typename Game::Mask;
    using Slot = SlotT<Mask>;
    uint32_t log2cap = 4;
    while (((uint64_t)1 << log2cap) < 2 * (iterations + 2)) ++log2cap;
    std::vector<Slot> table((size_t)1 << log2cap);
    std::memset(table.data(), 0, table.size() * sizeof(Slot));
    std::vector<double> ln(iterations + 4);
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

    P.tree_thresh = g_thresh; P.tree_period = g_period;
    uint64_t scratch[Game::kHandWords];
    uint32_t ring[LaneRng::kRingWords];
    std::vector<uint32_t> path(Game::kMaxDepth);
    lane_tree_search<Game>(P, 0, true, scratch, 1, ring, 1, path.data());

    uint32_t n = 0;
    for (const Slot& sl : table) if (sl.key != 0) ++n;
    *count = n;
    if (n > capacity) return 5;
    for (const Slot& sl : table) {
        if (sl.key == 0) continue;
        ismcts_node_rec& r = nodes[sl.created];
        if (sl.key == kRootKey) { r.parent = 0; r.move = 0; }
        else { r.parent = table[slot_parent_of(sl.key)].created; r.move = slot_move_of(sl.key); }
        r.player = sl.player; r.visits = sl.visits; r.score2 = sl.score2; r.avail = sl.avail;
        r.score = 0.5 * (double)sl.score2; r.prob = 1.0;
    }
    return 0;
}

extern "C" {

// Schedule of the search loop (when pending tree operations run); results must not depend on it.
void emu_set_schedule(uint32_t thresh, uint32_t period) { g_thresh = thresh; g_period = period; }

uint32_t emu_playout(uint32_t game, const void* root, uint64_t seed, uint32_t pos, uint32_t iteration)
{
    if (game == ISMCTS_GAME_KW) return playout_t<KwGame>(root, seed, pos, iteration);
    return playout_t<MnkGame>(root, seed, pos, iteration);
}

int emu_so_search(uint32_t game, const void* root, uint64_t iterations, uint64_t seed, uint32_t pos, uint32_t tree,
                  double exploration, ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count)
{
    if (game == ISMCTS_GAME_KW)
        return so_search_t<KwGame>(root, iterations, seed, pos, tree, exploration, nodes, capacity, count);
    return so_search_t<MnkGame>(root, iterations, seed, pos, tree, exploration, nodes, capacity, count);
}

} // extern "C"
