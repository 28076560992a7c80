// lane_search.cuh — the search loop: one lane = one SO-ISMCTS tree, one warp = 32
// trees advancing together through ONE loop of micro-steps.
//
// What it restates: SOSolver::search/select/expand
// (sosolver.h:44-71), SolverBase::simulate/backPropagate/selectNode
// (solverbase.h:36-56), Node::selectChild/untriedMoves/update
// (tree/node.h:58-91), the UCB1 policy (tree/ucb1.h:36-76) and the iteration
// budget of ExecutionPolicy (execution.h:42-63,112-124, utility.h:43-60), with
// exactly the reference's sequential semantics per tree; root parallelism
// (execution.h:181-232) comes from the trees being independent.
//
// How it is mapped to the GPU.  An iteration is a sequence of micro-steps of two
// kinds.  (1) Random micro-steps: "remove a uniformly random member from a bit
// set" — a determinisation or re-deal card, a tie-break, a rollout move; about
// 150 per Knockout Whist iteration, a few dozen instructions each, identical
// code for every lane whatever iteration or phase it is in.  (2) Tree
// operations: the UCB descent with the expansion (once per iteration) and the
// backpropagation with the start of the next iteration (once per iteration);
// memory-bound and divergent.  Every pass of the loop performs one random
// micro-step for all lanes that need one; lanes that need a tree operation wait
// until enough of them are pending (or a few passes have gone by) and then do
// them together.  Lanes never wait for the longest playout of their warp: a
// lane that finishes an iteration starts its next one while its neighbours are
// still rolling out.  The result of a tree does not depend on this schedule —
// each lane performs exactly the sequential sequence of operations.
#pragma once

#include "devdefs.cuh"
#include "philox.cuh"
#include "pool.cuh"

namespace ismcts {

struct SearchParams {
    const void* prepared;    // n_positions Game::Prepared records (prepare_kernel)
    uint32_t n_positions;
    uint32_t trees;          // trees per position in this launch
    uint32_t trees_total;    // trees per position over all devices
    uint32_t tree_base;      // global id of this launch's first tree
    uint32_t pos_base;       // global id of this launch's first position
    uint64_t iterations;     // total per position (count mode), 0 = time mode
    uint64_t time_ns;        // time-mode budget
    uint64_t seed;
    double exploration;      // UCB1 constant, ucb1.h:65-67
    void* pool;              // [n_positions * trees] tables of (1 << log2cap) slots, zeroed
    uint32_t log2cap;
    uint32_t max_nodes;      // a tree stops iterating before exceeding this many nodes
    const double* ln_table;  // ln(i) for i < ln_count, built on the host (bit-exact with the oracle)
    uint32_t ln_count;
    uint64_t* visits;        // [n_positions * stride] root-child sums (compileVisitCounts, execution.h:219-231)
    uint64_t* score2;
    uint64_t* avail;
    uint64_t* iters_done;    // [n_positions * trees]
    uint32_t* playout_out;   // playout-only launches: one packed result per thread
    uint32_t tree_thresh;    // run the pending tree operations when this many lanes of the warp wait ...
    uint32_t tree_period;    // ... or at the latest every tree_period passes (a power of two)
    uint32_t sched_mode;     // Schedule::mode
    uint32_t sched_split_min;
};

// Iterations of one tree in count mode: the reference drains one shared counter
// (execution.h:143-154); here the split is fixed so it can be replayed.
ISMCTS_DEV uint64_t tree_iterations(uint64_t total, uint32_t trees_total, uint32_t tree_global)
{
    return total / trees_total + ((uint64_t)tree_global < total % trees_total ? 1u : 0u);
}

enum LanePhase : uint32_t { kSelect = 0, kRollout = 1, kDone = 2 };

// Which kinds of micro-step a pass performs (the same for all lanes of the warp).
struct PassPlan { bool chance, move, tree; };

// The schedule of a warp: decides from the numbers of lanes waiting for a chance
// event, a rollout move and a tree operation what the next pass does.  A pass
// that does one kind only runs one short code path with the waiting lanes
// active; a pass that does several kinds runs the paths one after the other.
// The result of a search does not depend on the schedule (tests/emu checks that).
struct Schedule {
    uint32_t tree_thresh, tree_period, mode, split_min;
    uint32_t current; // kind the warp is running (hysteresis modes)

    ISMCTS_DEV void init(const SearchParams& P)
    {
        tree_thresh = P.tree_thresh; tree_period = P.tree_period; mode = P.sched_mode; split_min = P.sched_split_min;
        current = 0;
    }

    ISMCTS_DEV PassPlan plan_pass(uint32_t n_chance, uint32_t n_move, uint32_t n_tree, uint32_t pass)
    {
        PassPlan plan;
        plan.tree = n_tree != 0 && (n_tree >= tree_thresh || (pass & (tree_period - 1u)) == tree_period - 1u ||
                                    n_chance + n_move == 0);
        if (mode == 0) {                       // every kind every pass
            plan.chance = n_chance != 0; plan.move = n_move != 0;
        } else if (mode == 1) {                // the kind more lanes wait for
            plan.chance = n_chance >= n_move && n_chance != 0;
            plan.move = !plan.chance && n_move != 0;
        } else if (mode == 2) {                // majority, but both when the minority is large enough
            const uint32_t lo = n_chance < n_move ? n_chance : n_move;
            const bool both = lo >= split_min;
            plan.chance = (both || n_chance >= n_move) && n_chance != 0;
            plan.move = (both || n_move > n_chance) && n_move != 0;
        } else {                               // stay with a kind until few lanes want it
            const uint32_t n_cur = current == 0 ? n_chance : n_move, n_other = current == 0 ? n_move : n_chance;
            if (n_cur < split_min && n_other > n_cur) current ^= 1u;
            plan.chance = current == 0 && n_chance != 0;
            plan.move = current == 1 && n_move != 0;
            if (!plan.chance && !plan.move) { plan.chance = n_chance != 0; plan.move = n_move != 0 && !plan.chance; }
        }
        return plan;
    }
};

// kTree = false: determinise + rollout only (SolverBase::simulate after
// cloneAndRandomise), one iteration per lane, no tree — the playout kernel.
template <class Game, bool kTree>
struct Lane {
    using Mask = typename Game::Mask;
    using Prepared = typename Game::Prepared;
    using Slot = SlotT<Mask>;

    const SearchParams& P;
    Game g;
    LaneRng rng;
    TreePool<Mask> tree;
    const Prepared* prep;    // this lane's position, ready to start an iteration from
    uint32_t* path;          // slot | player << 24 of the nodes on the current descent
    uint32_t phase, node, depth, plies;
    Mask node_children, root_children;
    uint64_t done, quota, deadline;
    uint32_t first_iteration; // index of this lane's first iteration (playout kernel: the playout id)

    ISMCTS_DEV explicit Lane(const SearchParams& p) : P(p) {}

    ISMCTS_DEV void start_iteration()
    {
        rng.begin_iteration(first_iteration + (uint32_t)done);
        g.start(prep);                                        // cloneAndRandomise, sosolver.h:46
        node = 0; depth = 0; plies = 0;
        node_children = root_children;
        phase = kTree ? kSelect : kRollout;
    }

    ISMCTS_DEV bool budget_left() const
    {
        if (kTree && tree.count >= P.max_nodes) return false;
        return P.iterations > 0 || !kTree ? done < quota : global_timer_ns() < deadline;
    }

    // The game is over (validMoves() empty): backPropagate, solverbase.h:45-51;
    // Node::update, node.h:74-80 (the root is never updated); UCBNode::updateData,
    // ucb1.h:53-56.  One 64-bit reduction per node on the path: visits += 1 in
    // the low word, score2 += 2 * reward in the high word.  Then the next iteration.
    ISMCTS_DEV void finish_iteration(uint32_t gtid)
    {
        if (kTree) {
            for (uint32_t i = 0; i < depth; ++i) {
                const uint32_t e = path[i];
                const uint64_t delta = 1u | ((uint64_t)g.result2(e >> 24) << 32);
                red_add_u64(reinterpret_cast<uint64_t*>(&tree.slots[e & 0xFFFFFFu]), delta);
            }
        } else {
            P.playout_out[gtid] = (g.outcome() & 0xFFu) | ((plies & 0xFFFu) << 8) | ((rng.consumed() & 0xFFFu) << 20);
        }
        ++done;
        if (budget_left()) start_iteration(); else phase = kDone;
    }

    // select + expand, sosolver.h:53-71, from the current node until a node is
    // added, the game ends or a move triggers chance events (they are drained by
    // the random micro-steps of the following passes, then the descent resumes).
    ISMCTS_DEV void descend()
    {
        for (;;) {
            const Mask legal = g.legal();
            if (!legal.any()) return;                             // terminal: selectNode, solverbase.h:53-56
            const Mask untried = legal.minus(node_children);      // untriedMoves, node.h:82-91
            if (untried.any()) {                                  // expand, sosolver.h:63-71
                const uint32_t move = untried.kth(rng.below(untried.count()));
                const uint32_t who = g.current_player();          // newChild, solverbase.h:74-80: before the move
                const Mask grown = node_children.with(move);
                tree.slots[node].child_mask = grown;
                if (node == 0) root_children = grown;
                node = tree.insert(node, move, who);
                path[depth++] = node | (who << 24);
                g.step(move);
                phase = kRollout;
                return;
            }
            // Node::selectChild + UCB1::operator(), node.h:58-72, ucb1.h:69-76: every
            // legal child becomes available, then the first maximum in creation order.
            uint32_t best = 0, best_created = 0xFFFFFFFFu, best_move = 0, best_player = 0;
            double best_u = -1.0;
            Mask best_children = Mask::none();
            for (Mask rest = legal; rest.any();) {
                const uint32_t move = rest.pop_lowest();
                Slot c;
                const uint32_t ci = tree.find(node, move, c);
                c.avail += 1;
                tree.slots[ci].avail = c.avail;
                const double u = ismcts_ucb_value(c.score2, c.visits, P.ln_table[c.avail], P.exploration);
                if (u > best_u || (u == best_u && c.created < best_created)) {
                    best_u = u; best = ci; best_created = c.created; best_move = move; best_player = c.player;
                    best_children = c.child_mask;
                }
            }
            node = best;
            node_children = best_children;
            path[depth++] = node | (best_player << 24);
            g.step(best_move);
            if (g.chance_pending()) return;
        }
    }

    // ---- one pass of the loop, in two halves so that the warp votes sit between them

    struct Wants { bool chance, move, tree, over; };

    ISMCTS_DEV Wants classify() const
    {
        Wants w;
        const bool playing = phase != kDone;
        w.chance = playing && g.chance_pending();
        w.over = playing && !w.chance && g.terminal();
        w.tree = w.over || (playing && !w.chance && phase == kSelect);
        w.move = playing && !w.chance && !w.tree;
        return w;
    }

    ISMCTS_DEV void act(const Wants& w, const PassPlan& plan, uint32_t pass, uint32_t gtid)
    {
        const bool do_chance = plan.chance && w.chance, do_move = plan.move && w.move;
        if (do_chance || do_move) {
            // One random micro-step.  Chance: a determinisation (cloneAndRandomise) or re-deal card or
            // a tie-break.  Move: simulate, solverbase.h:36-43 with RandomElement, utility.h:69-83.
            // Both are "the k-th member of a set", drawn by code the two kinds share.
            const Mask set = do_chance ? g.chance_set() : g.legal();
            const uint32_t b = set.kth(rng.below(set.count()));
            if (do_chance) g.chance_step(b);
            else { ++plies; g.play(b); }
        }
        if (plan.tree && w.tree) {
            if (w.over) finish_iteration(gtid);
            else descend();
        }
        if ((pass & 3u) == 3u && phase != kDone && rng.low()) rng.refill();
    }

    // The loop.  Every pass performs, for the whole warp, the kinds of micro-step the
    // schedule (plan_pass) picks from the numbers of lanes waiting for each kind.
    ISMCTS_DEV void run(uint32_t gtid)
    {
        Schedule sched;
        sched.init(P);
        for (uint32_t pass = 0;; ++pass) {
            if (!warp_any(phase != kDone)) break;
            const Wants w = classify();
            const uint32_t n_chance = popc32(warp_ballot(w.chance)), n_move = popc32(warp_ballot(w.move)),
                           n_tree = popc32(warp_ballot(w.tree));
            const PassPlan plan = sched.plan_pass(n_chance, n_move, n_tree, pass);
            act(w, plan, pass, gtid);
        }
    }
};

// The whole search of tree `gtid` of this launch (`valid` = false for the lanes
// that only pad the last warp: they take part in the warp votes and do nothing).
template <class Game>
ISMCTS_DEV void lane_tree_search(const SearchParams& P, uint32_t gtid, bool valid, uint64_t* scratch,
                                 uint32_t scratch_stride, uint32_t* ring, uint32_t ring_stride, uint32_t* path)
{
    using Mask = typename Game::Mask;
    using Slot = SlotT<Mask>;
    Lane<Game, true> L(P);
    L.path = path;
    L.g.bind(scratch, scratch_stride);
    L.phase = kDone; L.done = 0; L.quota = 0; L.deadline = 0; L.first_iteration = 0;
    L.root_children = Mask::none();
    uint32_t pos = 0;
    if (valid) {
        pos = gtid / P.trees;
        const uint32_t tree_global = P.tree_base + gtid % P.trees;
        L.prep = reinterpret_cast<const typename Game::Prepared*>(P.prepared) + pos;
        L.tree.bind(reinterpret_cast<Slot*>(P.pool) + ((size_t)gtid << P.log2cap), P.log2cap);
        L.tree.init_root();
        L.rng.init(P.seed, P.pos_base + pos, tree_global, ring, ring_stride);
        if (P.iterations > 0) L.quota = tree_iterations(P.iterations, P.trees_total, tree_global);
        else L.deadline = global_timer_ns() + P.time_ns;   // time mode, utility.h:51-60
        if (L.budget_left()) L.start_iteration();
    }
    L.run(gtid);
    if (!valid) return;
    P.iters_done[gtid] = L.done;

    // Root children -> per-position sums (RootParallel::compileVisitCounts, execution.h:219-231).
    uint64_t* v = P.visits + (size_t)pos * Game::kMoveStride;
    uint64_t* s = P.score2 + (size_t)pos * Game::kMoveStride;
    uint64_t* a = P.avail + (size_t)pos * Game::kMoveStride;
    for (Mask rest = L.root_children; rest.any();) {
        const uint32_t move = rest.pop_lowest();
        Slot c;
        L.tree.find(0, move, c);
        atomic_add_u64(v + move, c.visits);
        atomic_add_u64(s + move, c.score2);
        atomic_add_u64(a + move, c.avail);
    }
}

// Determinise + rollout only: playout `gtid % per_position` of position `gtid / per_position`.
template <class Game>
ISMCTS_DEV void lane_playout(const SearchParams& P, uint32_t gtid, bool valid, uint32_t per_position, uint64_t* scratch,
                             uint32_t scratch_stride, uint32_t* ring, uint32_t ring_stride)
{
    using Mask = typename Game::Mask;
    Lane<Game, false> L(P);
    L.path = nullptr;
    L.g.bind(scratch, scratch_stride);
    L.phase = kDone; L.done = 0; L.quota = 1; L.deadline = 0; L.first_iteration = 0;
    L.root_children = Mask::none();
    if (valid) {
        const uint32_t pos = gtid / per_position;
        L.first_iteration = gtid % per_position;
        L.prep = reinterpret_cast<const typename Game::Prepared*>(P.prepared) + pos;
        L.rng.init(P.seed, P.pos_base + pos, 0, ring, ring_stride);
        L.start_iteration();
    }
    L.run(gtid);
}

} // namespace ismcts
