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
#include "../../ismcsolver_b200/csrc/kw_jobs.cuh"

#include <chrono>
#include <cmath>
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

using namespace ismcts;

static uint32_t g_thresh = 8, g_period = 8, g_mode = 0, g_split = 8;

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
    lane_playout<Game>(P, iteration, true, 0x7FFFFFFFu, scratch, 1, ring, 1);
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
    lane_tree_search<Game, Slot, kKind, false>(P, 0, true, scratch, 1, ring, 1, path.data());

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

// kw_jobs.cuh on the host: one CTA of `threads` lanes with K jobs each (jobs = trees 0.. of position 0),
// advanced pass by pass with the warp votes computed over all lanes.  Writes tree `dump` in creation order.
template <uint32_t K, uint32_t NP>
static int kw_jobs_search_t(const void* root, uint64_t iterations_per_tree, uint32_t threads, uint32_t trees, uint64_t seed,
                            uint32_t dump, ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count, uint64_t* passes)
{
    using Jobs = KwJobs<K, NP>;
    using Slot = typename Jobs::Slot;
    uint32_t log2cap = 4;
    while (((uint64_t)1 << log2cap) < 2 * (iterations_per_tree + 2)) ++log2cap;
    std::vector<Slot> pool((size_t)trees << log2cap);
    std::memset((void*)pool.data(), 0, pool.size() * sizeof(Slot));
    std::vector<double> ln(iterations_per_tree + 300);
    ln[0] = 0.0;
    for (size_t i = 1; i < ln.size(); ++i) ln[i] = std::log((double)i);
    std::vector<uint64_t> v(64), s2(64), a(64), it(trees);
    KwGame::Prepared prepared;
    uint64_t prep_scratch[KwGame::kHandWords];
    KwGame::prepare((const KwGame::Root*)root, &prepared, prep_scratch);
    SearchParams P;
    std::memset(&P, 0, sizeof P);
    P.prepared = &prepared; P.n_positions = 1; P.trees = trees; P.trees_total = trees; P.iterations = iterations_per_tree * trees;
    P.seed = seed; P.exploration = 0.7; P.pool = pool.data(); P.log2cap = log2cap; P.max_nodes = 1u << (log2cap - 1);
    P.ln_table = ln.data(); P.ln_count = (uint32_t)ln.size();
    P.visits = v.data(); P.score2 = s2.data(); P.avail = a.data(); P.iters_done = it.data();
    P.tree_thresh = g_thresh; P.tree_period = g_period; P.lanes_per_tree = 1;
    std::vector<uint64_t> smem((size_t)Jobs::kWordsPerJob * K * threads / 2 + 8);
    std::vector<uint32_t> paths((size_t)trees * KwGame::kMaxDepth);
    Jobs jobs(P);
    jobs.bind(reinterpret_cast<uint32_t*>(smem.data()), threads, 0);
    jobs.path_base = paths.data();
    jobs.deadline = 0;
    for (uint32_t t = 0; t < threads; ++t) jobs.init_lane(t);
    *passes = 0;
    for (uint32_t pass = 0;; ++pass) {
        uint32_t nc = 0, nm = 0, nt = 0, live = 0;
        for (uint32_t t = 0; t < threads; ++t) {
            const uint32_t bits = jobs.classify(t);
            nc += bits & 1u; nm += (bits >> 1) & 1u; nt += (bits >> 2) & 1u; live += (bits >> 3) & 1u;
        }
        if (!live) break;
        const uint32_t kind = Jobs::choose_kind(nc, nm, nt, P.tree_thresh, P.tree_period, pass);
        for (uint32_t t = 0; t < threads; ++t) jobs.act(t, kind);
        if ((pass & 3u) == 3u) for (uint32_t t = 0; t < threads; ++t) jobs.refill_pass(t);
        ++*passes;
    }
    for (uint32_t t = 0; t < threads; ++t) jobs.finish_lane(t);
    // dump tree `dump`
    const Slot* table = pool.data() + ((size_t)dump << log2cap);
    const size_t cap = (size_t)1 << log2cap;
    std::vector<std::pair<uint32_t, uint32_t>> members;
    for (size_t i = 0; i < cap; ++i) if (table[i].key != 0) members.emplace_back(i == 0 ? 0u : table[i].created, (uint32_t)i);
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
        r.player = sl.player; r.visits = sl.visits; r.score2 = sl.score2; r.avail = sl.avail;
        r.score = 0.5 * (double)sl.score2; r.prob = 1.0;
    }
    return it[dump] == iterations_per_tree ? 0 : 6;
}

extern "C" {

// Schedule of the search loop (when pending tree operations run); results must not depend on it.
void emu_set_schedule(uint32_t thresh, uint32_t period) { g_thresh = thresh; g_period = period; }
void emu_set_schedule_mode(uint32_t mode, uint32_t split_min) { g_mode = mode; g_split = split_min; }

int emu_kw_jobs_search(const void* root, uint64_t iterations_per_tree, uint32_t jobs_per_lane, uint32_t players_stored,
                       uint32_t threads, uint32_t trees, uint64_t seed, uint32_t dump, ismcts_node_rec* nodes,
                       uint32_t capacity, uint32_t* count, uint64_t* passes)
{
    if (jobs_per_lane == 2 && players_stored == 4)
        return kw_jobs_search_t<2, 4>(root, iterations_per_tree, threads, trees, seed, dump, nodes, capacity, count, passes);
    if (jobs_per_lane == 4 && players_stored == 4)
        return kw_jobs_search_t<4, 4>(root, iterations_per_tree, threads, trees, seed, dump, nodes, capacity, count, passes);
    if (jobs_per_lane == 2 && players_stored == 7)
        return kw_jobs_search_t<2, 7>(root, iterations_per_tree, threads, trees, seed, dump, nodes, capacity, count, passes);
    if (jobs_per_lane == 1 && players_stored == 7)
        return kw_jobs_search_t<1, 7>(root, iterations_per_tree, threads, trees, seed, dump, nodes, capacity, count, passes);
    return 2;
}

int emu_shared_search(uint32_t game, const void* root, uint64_t iterations, uint32_t width, uint64_t seed,
                      ismcts_node_rec* nodes, uint32_t capacity, uint32_t* count, uint64_t* done)
{
    if (game == ISMCTS_GAME_KW) return shared_search_t<KwGame>(root, iterations, width, seed, nodes, capacity, count, done);
    return shared_search_t<MnkGame>(root, iterations, width, seed, nodes, capacity, count, done);
}

// Run-queue scheduling experiment: J jobs (trees) per CTA, `warps` warps of 32 lanes.  A warp works on
// ONE kind of run at a time (0 chance run, 1 move run, 2 tree operation); a lane whose job's run ends
// parks the job in the queue of its next kind and idles until `swap_thresh` lanes of its warp idle (or
// all do), then the idle lanes take jobs of the warp's kind from the queue; a warp that cannot fill any
// lane switches to the kind with the longest queue.  stats: [0] warp-passes chance, [1] warp-passes move,
// [2] tree batches, [3] chance lane-steps, [4] move lane-steps, [5] tree ops, [6] swap batches,
// [7] jobs swapped in, [8] iterations, [9] idle warp-passes.
void emu_runqueue_sim(const void* root, uint32_t iters_per_tree, uint32_t jobs, uint32_t warps, uint32_t swap_thresh,
                      uint64_t seed, uint64_t* stats)
{
    using Game = KwGame;
    using Slot = SlotUcb<Game::Mask>;
    using L = Lane<Game, Slot, kSO>;
    uint32_t log2cap = 4;
    while (((uint64_t)1 << log2cap) < 2 * ((uint64_t)iters_per_tree + 2)) ++log2cap;
    std::vector<Slot> pool((size_t)jobs << log2cap);
    std::memset((void*)pool.data(), 0, pool.size() * sizeof(Slot));
    std::vector<double> ln(iters_per_tree + 4);
    ln[0] = 0.0;
    for (size_t i = 1; i < ln.size(); ++i) ln[i] = std::log((double)i);
    std::vector<uint64_t> v(64), s2(64), a(64), it(jobs);
    Game::Prepared prepared;
    uint64_t prep_scratch[Game::kHandWords];
    Game::prepare((const Game::Root*)root, &prepared, prep_scratch);
    SearchParams P;
    std::memset(&P, 0, sizeof P);
    P.prepared = &prepared; P.n_positions = 1; P.trees = jobs; P.trees_total = jobs; P.iterations = (uint64_t)iters_per_tree * jobs;
    P.seed = seed; P.exploration = 0.7; P.pool = pool.data(); P.log2cap = log2cap; P.max_nodes = 1u << (log2cap - 1);
    P.ln_table = ln.data(); P.ln_count = (uint32_t)ln.size();
    P.visits = v.data(); P.score2 = s2.data(); P.avail = a.data(); P.iters_done = it.data();
    P.tree_thresh = 1; P.tree_period = 1; P.sched_mode = 0; P.sched_split_min = 0; P.lanes_per_tree = 1;
    std::vector<std::unique_ptr<L>> job;
    std::vector<std::vector<uint64_t>> scratch(jobs, std::vector<uint64_t>(Game::kHandWords));
    std::vector<std::vector<uint32_t>> ring(jobs, std::vector<uint32_t>(LaneRng::kRingWords)), path(jobs, std::vector<uint32_t>(Game::kMaxDepth));
    for (uint32_t i = 0; i < jobs; ++i) {
        job.emplace_back(new L(P));
        L& l = *job.back();
        l.path = path[i].data();
        l.g.bind(scratch[i].data(), 1);
        l.done = 0; l.deadline = 0; l.first_iteration = 0; l.iteration_stride = 1; l.node_counter = nullptr;
        l.root_children = Game::Mask::none();
        l.prep = &prepared;
        l.tree.bind(pool.data() + ((size_t)i << log2cap), log2cap, 1);
        l.tree.init_roots();
        l.tree_players = 1;
        l.rng.init(seed, 0, i, ring[i].data(), 1);
        l.quota = iters_per_tree;
        l.start_iteration();
    }
    auto kind_of = [&](uint32_t j) -> int {
        const auto w = job[j]->classify();
        return w.chance ? 0 : w.move ? 1 : w.tree ? 2 : -1;
    };
    std::vector<std::vector<uint32_t>> queue(3);
    for (uint32_t j = 0; j < jobs; ++j) queue[kind_of(j)].push_back(j);
    std::vector<int> warp_kind(warps, 0);
    std::vector<std::vector<int>> lane_job(warps, std::vector<int>(32, -1));
    for (int i = 0; i < 10; ++i) stats[i] = 0;
    uint64_t remaining = jobs; // jobs not finished
    for (uint64_t round = 0; remaining > 0 && round < 100000000ull; ++round) {
        for (uint32_t w = 0; w < warps; ++w) {
            int idle = 0;
            for (int l = 0; l < 32; ++l) idle += lane_job[w][l] < 0;
            // refill when enough lanes idle
            if (idle >= (int)swap_thresh || idle == 32) {
                if (idle == 32 && queue[warp_kind[w]].empty()) {
                    int best = 0;
                    for (int k = 1; k < 3; ++k) if (queue[k].size() > queue[best].size()) best = k;
                    warp_kind[w] = best;
                }
                auto& q = queue[warp_kind[w]];
                int taken = 0;
                for (int l = 0; l < 32 && !q.empty(); ++l)
                    if (lane_job[w][l] < 0) { lane_job[w][l] = (int)q.back(); q.pop_back(); ++taken; }
                if (taken) { stats[6]++; stats[7] += taken; }
            }
            int active = 0;
            for (int l = 0; l < 32; ++l) active += lane_job[w][l] >= 0;
            if (!active) { stats[9]++; continue; }
            const int kind = warp_kind[w];
            if (kind == 0) { stats[0]++; stats[3] += active; } else if (kind == 1) { stats[1]++; stats[4] += active; } else { stats[2]++; stats[5] += active; }
            PassPlan plan; plan.chance = kind == 0; plan.move = kind == 1; plan.tree = kind == 2;
            for (int l = 0; l < 32; ++l) {
                const int j = lane_job[w][l];
                if (j < 0) continue;
                L& jl = *job[j];
                const auto wants = jl.classify();
                jl.act(wants, plan, 3 /* refill every pass */, j);
                const int next = kind_of(j);
                if (next != kind) {
                    lane_job[w][l] = -1;
                    if (next < 0) --remaining; else queue[next].push_back(j);
                }
            }
        }
    }
    for (uint32_t j = 0; j < jobs; ++j) stats[8] += job[j]->done;
}

// Multi-job lanes experiment: every lane owns `k` jobs (trees); a pass runs ONE kind of step for the whole
// warp — the kind most lanes can serve — and a lane steps one (or, with `all` set, every) job of that kind.
// stats: [0] chance passes, [1] move passes, [2] tree passes, [3] chance lane-slots used (lanes with >= 1
// job stepped), [4] move lane-slots, [5] tree lane-slots, [6] chance job-steps, [7] move job-steps,
// [8] tree ops, [9] iterations.
void emu_multijob_sim(const void* root, uint32_t iters_per_tree, uint32_t k, uint32_t all, uint32_t tree_thresh,
                      uint64_t seed, uint64_t* stats)
{
    using Game = KwGame;
    using Slot = SlotUcb<Game::Mask>;
    using L = Lane<Game, Slot, kSO>;
    const uint32_t jobs = 32 * k;
    uint32_t log2cap = 4;
    while (((uint64_t)1 << log2cap) < 2 * ((uint64_t)iters_per_tree + 2)) ++log2cap;
    std::vector<Slot> pool((size_t)jobs << log2cap);
    std::memset((void*)pool.data(), 0, pool.size() * sizeof(Slot));
    std::vector<double> ln(iters_per_tree + 4);
    ln[0] = 0.0;
    for (size_t i = 1; i < ln.size(); ++i) ln[i] = std::log((double)i);
    std::vector<uint64_t> v(64), s2(64), a(64), it(jobs);
    Game::Prepared prepared;
    uint64_t prep_scratch[Game::kHandWords];
    Game::prepare((const Game::Root*)root, &prepared, prep_scratch);
    SearchParams P;
    std::memset(&P, 0, sizeof P);
    P.prepared = &prepared; P.n_positions = 1; P.trees = jobs; P.trees_total = jobs; P.iterations = (uint64_t)iters_per_tree * jobs;
    P.seed = seed; P.exploration = 0.7; P.pool = pool.data(); P.log2cap = log2cap; P.max_nodes = 1u << (log2cap - 1);
    P.ln_table = ln.data(); P.ln_count = (uint32_t)ln.size();
    P.visits = v.data(); P.score2 = s2.data(); P.avail = a.data(); P.iters_done = it.data();
    P.tree_thresh = 1; P.tree_period = 1; P.sched_mode = 0; P.sched_split_min = 0; P.lanes_per_tree = 1;
    std::vector<std::unique_ptr<L>> job;
    std::vector<std::vector<uint64_t>> scratch(jobs, std::vector<uint64_t>(Game::kHandWords));
    std::vector<std::vector<uint32_t>> ring(jobs, std::vector<uint32_t>(LaneRng::kRingWords)), path(jobs, std::vector<uint32_t>(Game::kMaxDepth));
    for (uint32_t i = 0; i < jobs; ++i) {
        job.emplace_back(new L(P));
        L& l = *job.back();
        l.path = path[i].data();
        l.g.bind(scratch[i].data(), 1);
        l.done = 0; l.deadline = 0; l.first_iteration = 0; l.iteration_stride = 1; l.node_counter = nullptr;
        l.root_children = Game::Mask::none();
        l.prep = &prepared;
        l.tree.bind(pool.data() + ((size_t)i << log2cap), log2cap, 1);
        l.tree.init_roots();
        l.tree_players = 1;
        l.rng.init(seed, 0, i, ring[i].data(), 1);
        l.quota = iters_per_tree;
        l.start_iteration();
    }
    for (int i = 0; i < 10; ++i) stats[i] = 0;
    for (uint64_t pass = 0; pass < 100000000ull; ++pass) {
        // what every job wants
        std::vector<int> want(jobs);
        uint32_t lanes_with[3] = {0, 0, 0}, alive = 0;
        for (uint32_t j = 0; j < jobs; ++j) {
            const auto w = job[j]->classify();
            want[j] = w.chance ? 0 : w.move ? 1 : w.tree ? 2 : -1;
            alive += want[j] >= 0;
        }
        if (!alive) break;
        for (uint32_t l = 0; l < 32; ++l) {
            bool has[3] = {false, false, false};
            for (uint32_t q = 0; q < k; ++q) if (want[l * k + q] >= 0) has[want[l * k + q]] = true;
            for (int x = 0; x < 3; ++x) lanes_with[x] += has[x];
        }
        int kind;
        if (lanes_with[2] >= tree_thresh || lanes_with[0] + lanes_with[1] == 0) kind = 2;
        else kind = lanes_with[0] >= lanes_with[1] ? 0 : 1;
        stats[kind]++;
        PassPlan plan; plan.chance = kind == 0; plan.move = kind == 1; plan.tree = kind == 2;
        for (uint32_t l = 0; l < 32; ++l) {
            bool used = false;
            for (uint32_t q = 0; q < k; ++q) {
                const uint32_t j = l * k + q;
                if (want[j] != kind) continue;
                if (used && !all) break;
                const auto w = job[j]->classify();
                job[j]->act(w, plan, 3, j);
                stats[6 + kind]++;
                used = true;
            }
            stats[3 + kind] += used;
        }
    }
    for (uint32_t j = 0; j < jobs; ++j) stats[9] += job[j]->done;
}

// A warp of 32 lanes (32 trees of one Knockout Whist position) run pass by pass with real warp votes:
// measures how busy the lanes are under a schedule.  stats: [0] passes, [1] passes doing chance only,
// [2] move only, [3] both, [4] passes with tree operations, [5] chance lane-steps, [6] move lane-steps,
// [7] tree operations, [8] iterations.
void emu_warp_sim(const void* root, uint32_t iters_per_tree, uint64_t seed, uint64_t* stats)
{
    using Game = KwGame;
    using Slot = SlotUcb<Game::Mask>;
    using L = Lane<Game, Slot, kSO>;
    constexpr int W = 32;
    uint32_t log2cap = 4;
    while (((uint64_t)1 << log2cap) < 2 * ((uint64_t)iters_per_tree + 2)) ++log2cap;
    std::vector<Slot> pool((size_t)W << log2cap);
    std::memset((void*)pool.data(), 0, pool.size() * sizeof(Slot));
    std::vector<double> ln(iters_per_tree + 4);
    ln[0] = 0.0;
    for (size_t i = 1; i < ln.size(); ++i) ln[i] = std::log((double)i);
    std::vector<uint64_t> v(64), s2(64), a(64), it(W);
    Game::Prepared prepared;
    uint64_t prep_scratch[Game::kHandWords];
    Game::prepare((const Game::Root*)root, &prepared, prep_scratch);
    SearchParams P;
    std::memset(&P, 0, sizeof P);
    P.prepared = &prepared; P.n_positions = 1; P.trees = W; P.trees_total = W; P.iterations = (uint64_t)iters_per_tree * W;
    P.seed = seed; P.exploration = 0.7; P.pool = pool.data(); P.log2cap = log2cap; P.max_nodes = 1u << (log2cap - 1);
    P.ln_table = ln.data(); P.ln_count = (uint32_t)ln.size();
    P.visits = v.data(); P.score2 = s2.data(); P.avail = a.data(); P.iters_done = it.data();
    P.tree_thresh = g_thresh; P.tree_period = g_period; P.sched_mode = g_mode; P.sched_split_min = g_split;
    std::vector<std::unique_ptr<L>> lanes;
    std::vector<std::vector<uint64_t>> scratch(W, std::vector<uint64_t>(Game::kHandWords));
    std::vector<std::vector<uint32_t>> ring(W, std::vector<uint32_t>(LaneRng::kRingWords)), path(W, std::vector<uint32_t>(Game::kMaxDepth));
    for (int i = 0; i < W; ++i) {
        lanes.emplace_back(new L(P));
        L& l = *lanes.back();
        l.path = path[i].data();
        l.g.bind(scratch[i].data(), 1);
        l.done = 0; l.deadline = 0; l.first_iteration = 0; l.iteration_stride = 1; l.node_counter = nullptr;
        l.root_children = Game::Mask::none();
        l.prep = &prepared;
        l.tree.bind(pool.data() + ((size_t)i << log2cap), log2cap, 1);
        l.tree.init_roots();
        l.tree_players = 1;
        l.rng.init(seed, 0, i, ring[i].data(), 1);
        l.quota = iters_per_tree;
        l.start_iteration();
    }
    for (int i = 0; i < 9; ++i) stats[i] = 0;
    Schedule sched;
    sched.init(P);
    for (uint32_t pass = 0;; ++pass) {
        L::Wants w[W];
        uint32_t nc = 0, nm = 0, nt = 0, alive = 0;
        for (int i = 0; i < W; ++i) {
            alive += lanes[i]->phase != kDone;
            w[i] = lanes[i]->classify();
            nc += w[i].chance; nm += w[i].move; nt += w[i].tree;
        }
        if (!alive) break;
        const PassPlan plan = sched.plan_pass(nc, nm, nt, pass);
        stats[0]++;
        if (plan.chance && plan.move) stats[3]++; else if (plan.chance) stats[1]++; else if (plan.move) stats[2]++;
        if (plan.tree) { stats[4]++; stats[7] += nt; }
        if (plan.chance) stats[5] += nc;
        if (plan.move) stats[6] += nm;
        for (int i = 0; i < W; ++i) lanes[i]->act(w[i], plan, pass, i);
        if (P.sched_mode == 4 && nc + nm + nt == 0) for (int i = 0; i < W; ++i) lanes[i]->leave_barrier();
    }
    for (int i = 0; i < W; ++i) stats[8] += lanes[i]->done;
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
