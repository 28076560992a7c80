// so_search.cuh — one SO-ISMCTS tree searched by one GPU thread.
//
// Restates SOSolver::search/select/expand (sosolver.h:44-71) with
// SolverBase::simulate/backPropagate/selectNode (solverbase.h:36-56),
// Node::selectChild/untriedMoves/update (tree/node.h:58-91) and the UCB1 policy
// (tree/ucb1.h:36-76) on the flat node pool of pool.cuh.  Semantics are exactly
// the reference's sequential ones; root parallelism
// (execution.h:181-232) comes from running one such tree per thread.
#pragma once

#include "devdefs.cuh"
#include "philox.cuh"
#include "pool.cuh"

namespace ismcts {

struct SearchParams {
    const void* roots;       // n_positions packed Game::Root records
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
};

// Iterations of one tree in count mode: the reference drains one shared counter
// (execution.h:143-154); here the split is fixed so it can be replayed.
ISMCTS_DEV uint64_t tree_iterations(uint64_t total, uint32_t trees_total, uint32_t tree_global)
{
    return total / trees_total + ((uint64_t)tree_global < total % trees_total ? 1u : 0u);
}

// Uniform random playout to the end of the game: SolverBase::simulate,
// solverbase.h:36-43 with RandomElement (utility.h:69-83).  Chance events inside
// the game (Knockout Whist re-deals and tie-breaks) are micro-steps of the same
// loop, so that the lanes of a warp, each in its own game, run one loop body.
template <class Game>
ISMCTS_DEV uint32_t rollout(Game& g, Philox& rng)
{
    uint32_t plies = 0;
    for (;;) {
        const typename Game::Mask c = g.choices();
        if (!c.any()) break;
        plies += g.chance_pending() ? 0u : 1u;
        g.step(c.kth(rng.below(c.count())));
    }
    return plies;
}

template <class Game>
ISMCTS_DEV void so_iteration(Game& g, const typename Game::Root* root, TreePool<typename Game::Mask>& tree,
                             Philox& rng, const SearchParams& P, uint32_t* path)
{
    using Mask = typename Game::Mask;
    using Slot = SlotT<Mask>;
    g.load(root);
    g.determinise(g.current_player(), rng, root);            // sosolver.h:46

    uint32_t node = 0, depth = 0;
    Mask node_children = load_slot(&tree.slots[0]).child_mask;
    for (;;) {                                                // select, sosolver.h:53-61
        const Mask legal = g.legal();
        if (!legal.any()) break;                              // terminal: selectNode, solverbase.h:53-56
        const Mask untried = legal.minus(node_children);      // untriedMoves, node.h:82-91
        if (untried.any()) {                                  // expand, sosolver.h:63-71
            const uint32_t move = untried.kth(rng.below(untried.count()));
            const uint32_t who = g.current_player();          // newChild, solverbase.h:74-80: before the move
            tree.slots[node].child_mask = node_children.with(move);
            node = tree.insert(node, move, who);
            path[depth++] = node | (who << 24);
            g.apply(move, rng);
            break;
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
        g.apply(best_move, rng);
    }

    rng.align();
    rollout(g, rng);                                          // simulate, solverbase.h:36-43

    // backPropagate, solverbase.h:45-51; Node::update, node.h:74-80 (the root is
    // never updated); UCBNode::updateData, ucb1.h:53-56.  One 64-bit reduction per
    // node: visits += 1 in the low word, score2 += 2*reward in the high word.
    for (uint32_t i = 0; i < depth; ++i) {
        const uint32_t e = path[i];
        const uint64_t delta = 1u | ((uint64_t)g.result2(e >> 24) << 32);
        red_add_u64(reinterpret_cast<uint64_t*>(&tree.slots[e & 0xFFFFFFu]), delta);
    }
}

// The whole search of tree `gtid` of this launch.
template <class Game>
ISMCTS_DEV void so_tree_search(const SearchParams& P, uint32_t gtid, uint64_t* scratch, uint32_t scratch_stride)
{
    using Mask = typename Game::Mask;
    using Slot = SlotT<Mask>;
    const uint32_t pos = gtid / P.trees, local_tree = gtid % P.trees;
    const uint32_t tree_global = P.tree_base + local_tree;
    const typename Game::Root* root = reinterpret_cast<const typename Game::Root*>(P.roots) + pos;

    Game g;
    g.bind(scratch, scratch_stride);
    TreePool<Mask> tree;
    tree.bind(reinterpret_cast<Slot*>(P.pool) + ((size_t)gtid << P.log2cap), P.log2cap);
    tree.init_root();
    Philox rng;
    rng.init(P.seed, P.pos_base + pos, tree_global);
    uint32_t path[Game::kMaxDepth];

    uint64_t done = 0;
    if (P.iterations > 0) {
        uint64_t n = tree_iterations(P.iterations, P.trees_total, tree_global);
        if (n > P.max_nodes - 1u) n = P.max_nodes - 1u;
        for (; done < n; ++done) {
            rng.begin_iteration((uint32_t)done);
            so_iteration<Game>(g, root, tree, rng, P, path);
        }
    } else {
        // time mode, utility.h:51-60: iterate until the budget is spent
        const uint64_t deadline = global_timer_ns() + P.time_ns;
        while (tree.count < P.max_nodes && global_timer_ns() < deadline) {
            rng.begin_iteration((uint32_t)done);
            so_iteration<Game>(g, root, tree, rng, P, path);
            ++done;
        }
    }
    P.iters_done[gtid] = done;

    // Root children -> per-position sums (RootParallel::compileVisitCounts, execution.h:219-231).
    uint64_t* v = P.visits + (size_t)pos * Game::kMoveStride;
    uint64_t* s = P.score2 + (size_t)pos * Game::kMoveStride;
    uint64_t* a = P.avail + (size_t)pos * Game::kMoveStride;
    for (Mask rest = load_slot(&tree.slots[0]).child_mask; rest.any();) {
        const uint32_t move = rest.pop_lowest();
        Slot c;
        tree.find(0, move, c);
        atomic_add_u64(v + move, c.visits);
        atomic_add_u64(s + move, c.score2);
        atomic_add_u64(a + move, c.avail);
    }
}

// Determinise + rollout only (no tree): SolverBase::simulate after cloneAndRandomise.
template <class Game>
ISMCTS_DEV uint32_t playout_only(const typename Game::Root* root, uint64_t seed, uint32_t pos, uint32_t iteration,
                                 uint64_t* scratch, uint32_t scratch_stride)
{
    Game g;
    g.bind(scratch, scratch_stride);
    Philox rng;
    rng.init(seed, pos, 0);
    rng.begin_iteration(iteration);
    g.load(root);
    g.determinise(g.current_player(), rng, root);
    rng.align();
    const uint32_t plies = rollout(g, rng);
    return (g.outcome() & 0xFFu) | ((plies & 0xFFFu) << 8) | ((rng.d & 0xFFFu) << 20);
}

} // namespace ismcts
