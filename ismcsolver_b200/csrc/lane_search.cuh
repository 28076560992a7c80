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
// How it is mapped to the GPU: the lock-step loop.  The 32 lanes of a warp start their
// iterations together and stay in the same PHASE of the iteration, so that the warp runs one
// short code path at a time with (nearly) all lanes active:
//   1. chance events — the determinisation (cloneAndRandomise) at the start, later the
//      re-deals and tie-breaks a move triggers: every lane draws until nothing is pending;
//   2. descent — UCB1 / EXP3 selection and the expansion, level by level (lanes leave the
//      loop at different depths; a move that triggers chance events sends the warp to 1);
//   3. rollout — every lane plays uniformly random moves until its round or game ends
//      (chance events of the round end: back to 1);
//   4. backpropagation and the start of the next iteration, together.
// A lane with nothing to do in a phase idles: in Knockout Whist the lanes of a position deal
// the same number of cards per player and play the same number of tricks per round, they
// differ in the number of players still in and in the number of rounds (measured: 73 % of
// the lane-steps are useful).  The loop this replaced let every lane do whatever step it
// needed in every pass; a pass then executed the union of all code paths with half of the
// lanes active in each and was 2.2x (Knockout Whist, SO) to 3.6x (MO, EXP3; m,n,k) slower
// (profiles/r01_lockstep.md).  The result of a tree does not depend on the schedule — each
// lane performs exactly the sequential sequence of operations of its tree.
#pragma once

#include <new>
#include <type_traits>
#include "devdefs.cuh"
#include "philox.cuh"
#include "pool.cuh"

namespace ismcts {

struct alignas(16) UcbVisitFactors { double inv, inv_sqrt; };
constexpr uint32_t kLnSmall = 4096;   // entries of the ln() table EXP3 reads (a node has at most 256 children)

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
    const double* ln_table;  // EXP3: ln(K) for K < kLnSmall children, built on the host (bit-exact with the oracle)
    // UCB1 (include/ismcts_detmath.h: ismcts_ucb_value): {1 / v, 1 / sqrt(v)} for v < ln_count visits and sqrt(ln(a)) for
    // a < ln_count availabilities, built on the host
    const UcbVisitFactors* visit_tab;
    const double* sqrt_ln_tab;
    uint32_t ln_count;
    uint64_t* visits;        // [n_positions * stride] root-child sums (compileVisitCounts, execution.h:219-231)
    uint64_t* score2;
    uint64_t* avail;
    uint64_t* iters_done;    // [n_positions * trees]
    uint32_t* playout_out;   // playout-only launches: one packed result per thread
    // tree parallelisation: `lanes_per_tree` lanes search ONE tree together (1 = private trees)
    uint32_t lanes_per_tree;
    // time mode: the device clock when the search was launched (written by prepare_kernel): ONE deadline for
    // all lanes, whichever wave of blocks they run in (null: every lane starts its own clock)
    const uint64_t* start_ns;
    // private trees, count mode, more trees than resident lanes: the warps of a persistent grid claim 32 trees at a
    // time from this counter and reuse their tables (null: lane = tree, one pass)
    uint32_t* unit_queue;
};

// Iterations of one tree in count mode: the reference drains one shared counter
// (execution.h:143-154); here the split is fixed so it can be replayed.
ISMCTS_DEV uint64_t tree_iterations(uint64_t total, uint32_t trees_total, uint32_t tree_global)
{
    return total / trees_total + ((uint64_t)tree_global < total % trees_total ? 1u : 0u);
}

enum LanePhase : uint32_t { kSelect = 0, kRollout = 1, kDone = 2 };

// Games that bring their own phase loops for the lock-step schedule (chance_phase, play_phase).
template <class G> ISMCTS_HDI constexpr auto game_is_phased(int) -> decltype(G::kPhased) { return G::kPhased; }
template <class G> ISMCTS_HDI constexpr bool game_is_phased(long) { return false; }

// Draws a word of the stream serves in a game (philox.cuh): Game::kDrawsPerWord, else two.
template <class G> ISMCTS_HDI constexpr auto game_draws_per_word(int) -> decltype(G::kDrawsPerWord) { return G::kDrawsPerWord; }
template <class G> ISMCTS_HDI constexpr uint32_t game_draws_per_word(long) { return LaneRng::kDrawsPerWord; }

// Bytes of per-block tables a game keeps behind the scratch of the block's games (Game::kBlockExtraBytes).
template <class G> ISMCTS_HDI constexpr auto game_block_extra_bytes(int) -> decltype(G::kBlockExtraBytes) { return G::kBlockExtraBytes; }
template <class G> ISMCTS_HDI constexpr uint32_t game_block_extra_bytes(long) { return 0; }

// Games whose object lives in their own scratch while the tree operations run (Game::home(), kw_game.cuh):
// shared memory instead of the thread's local memory.
template <class G> ISMCTS_HDI constexpr auto game_home_in_scratch(int) -> decltype(G::kHomeInScratch) { return G::kHomeInScratch; }
template <class G> ISMCTS_HDI constexpr bool game_home_in_scratch(long) { return false; }

// The copy of a game the loops of the phases work on (Game::Hot where a game has one).
template <class G, class = void> struct GameHot { using type = G; };
template <class G> struct GameHot<G, std::void_t<typename G::Hot>> { using type = typename G::Hot; };

template <class Game, bool kHome>
struct LaneGameStore {
    Game g_;
    ISMCTS_DEV Game& game() { return g_; }
    ISMCTS_DEV const Game& game() const { return g_; }
};
template <class Game>
struct LaneGameStore<Game, true> {
    ISMCTS_DEV Game& game() { return *Game::home(); }
    ISMCTS_DEV const Game& game() const { return *Game::home(); }
};

// Solver kinds of a lane.  kPlayout: determinise + rollout only
// (SolverBase::simulate after cloneAndRandomise), one iteration per lane, no tree.
enum LaneKind : uint32_t { kPlayout = 0, kSO = 1, kMO = 2 };

// kShared: tree parallelisation (TreeParallel, execution.h:169-179) — the lanes of a tree group
// work on ONE tree: nodes are reserved with compare-and-swap, statistics are updated with
// reductions, and every node a descent passes through is counted as visited at once (virtual
// loss: the visit is in, the reward follows at backpropagation), which steers the other
// lanes away from the path while its playout is still running.  MO-ISMCTS shares the trees of all
// players the same way; EXP3 updates its scores with floating-point atomics.
template <class Game, class SlotType, uint32_t kKind, bool kShared = false>
struct Lane : LaneGameStore<Game, game_home_in_scratch<Game>(0)> {
    using LaneGameStore<Game, game_home_in_scratch<Game>(0)>::game;
    static constexpr bool kHome = game_home_in_scratch<Game>(0);
    using Mask = typename Game::Mask;
    using Prepared = typename Game::Prepared;
    using Slot = SlotType;
    static constexpr bool kTree = kKind != kPlayout;
    static constexpr bool kMulti = kKind == kMO;
    static constexpr bool kExp3 = Slot::kExp3;
    // MO-ISMCTS: one tree per player (mosolver.h:31-34); the root of player p's tree is slot p
    static constexpr uint32_t kTrees = kMulti ? Game::kMaxPlayers : 1;
    // The descent records the nodes it passes (slot | player << 24; one entry per tree and level) for the
    // backpropagation: private trees of every kind, and shared SO trees with UCB1.  The other shared trees walk
    // the parent links instead.
    static constexpr bool kPathed = kTree && (!kShared || (!kMulti && !kExp3));
    static constexpr uint32_t kPathWords = kPathed ? Game::kMaxDepth * kTrees : 1;

    const SearchParams& P;
    LaneRng rng;
    TreePool<Slot> tree;
    const Prepared* prep;    // this lane's position, ready to start an iteration from
    uint32_t* path;          // SO + UCB1: slot | player << 24 of the nodes on the current descent
    uint32_t phase, depth, plies;
    uint32_t node[kTrees];   // the cursor of every tree
    Mask node_children[kTrees];
    Mask root_children[kTrees]; // private trees: children of the root(s) (saves their reload every iteration)
    uint32_t done, quota;
    uint64_t deadline;
    uint32_t first_iteration; // index of this lane's first iteration (playout kernel: the playout id)
    uint32_t iteration_stride; // shared tree: the lanes of a tree interleave their iteration ids
    uint32_t my_nodes;        // shared tree: nodes this lane has created; its share of the table is max_nodes / lanes
    uint32_t created_seq;     // shared tree: nodes this lane has created in the current iteration
    uint32_t mask_known;      // shared SO tree: node_children[0] holds the child mask of the cursor (descend_shared)
    uint32_t tree_players;    // MO: bit p = player p has a tree (the players of the root position)

    ISMCTS_DEV explicit Lane(const SearchParams& p) : P(p) {}

    // A draw of a tree operation (the loops of the phases draw for themselves).
    ISMCTS_DEV uint32_t draw_below(uint32_t n)
    {
        if constexpr (game_is_phased<Game>(0)) return rng.below_direct(n); else return rng.below(n);
    }
    ISMCTS_DEV uint32_t draw_word()
    {
        if constexpr (game_is_phased<Game>(0)) return rng.word_direct(); else return rng.word();
    }

    ISMCTS_DEV void start_iteration()
    {
        rng.begin_iteration(first_iteration + done * iteration_stride);
        game().start(prep);                                        // cloneAndRandomise, sosolver.h:46 / mosolver.h:60
        depth = 0; plies = 0; created_seq = 0; mask_known = 0;
        if (kMulti) {
            // (private trees: the lane is the only writer of its roots' masks; shared trees re-read masks at every level)
            for (uint32_t t = 0; t < kTrees; ++t) {
                node[t] = t;
                if constexpr (kShared) node_children[t] = load_slot(&tree.slots[t]).child_mask;
                else node_children[t] = root_children[t];
            }
        } else {
            node[0] = 0; node_children[0] = root_children[0];
        }
        phase = kTree ? kSelect : kRollout;
    }

    ISMCTS_DEV bool budget_left() const
    {
        if (kTree && !kShared && tree.count + kTrees > P.max_nodes) return false;
        if (kShared && (uint64_t)(my_nodes + kTrees) * P.lanes_per_tree + kTrees > P.max_nodes) return false;
        if (P.iterations > 0 || !kTree) return done < quota;
        // time mode: the availability of a node grows by one per iteration even when the tree has stopped
        // growing (a root with one legal move), and ln() is a table: stop before its end
        return done + 3u < P.ln_count && global_timer_ns() < deadline;
    }

    // Node::update + updateData of one node (node.h:74-80; ucb1.h:53-56; exp3.h:45-48).
    ISMCTS_DEV void update_node(uint32_t slot, uint32_t who, double prob)
    {
        const uint32_t r2 = game().result2(who);
        if (!kExp3) {
            // one 64-bit reduction: visits += 1 in the low word, score2 += 2 * reward in the high word
            red_add_u64(reinterpret_cast<uint64_t*>(&tree.slots[slot]), 1u | ((uint64_t)r2 << 32));
        } else {
            update_exp(&tree.slots[slot], r2, prob);
        }
    }
    template <class S>
    ISMCTS_DEV static void update_exp(S* s, uint32_t r2, double prob)
    {
        if constexpr (S::kExp3) {
            // (reductions: an IEEE addition each, performed where the node lives — nothing of the node is loaded)
            red_add_u32(&s->visits, 1u);
            atomic_add_f64(&s->score, ISMCTS_DDIV(ISMCTS_DMUL((double)r2, 0.5), prob));
        }
    }

    // The game is over (validMoves() empty): backPropagate, solverbase.h:45-51 /
    // mosolver.h:97-101 — every node from the cursor up to but excluding the root.
    // Then the next iteration.
    ISMCTS_DEV_CALL void finish_iteration(uint32_t gtid)
    {
        if (kTree) {
            if (kShared && !kMulti && !kExp3) {
                // the visits were counted on the way down (virtual loss); now the rewards
                for (uint32_t i = 0; i < depth; ++i)
                    red_add_u64(reinterpret_cast<uint64_t*>(&tree.slots[path[i] & 0xFFFFFFu]),
                                (uint64_t)game().result2(path[i] >> 24) << 32);
            } else if (kShared) {
                // the same along the parent links of every tree
                for (uint32_t t = 0; t < kTrees; ++t) {
                    uint32_t n = node[t];
                    while (n >= tree.roots) {
                        const Slot s = load_slot(&tree.slots[n]);
                        add_reward_shared(&tree.slots[n], game().result2(s.player), slot_prob(s));
                        n = slot_parent_of(s.key);
                    }
                }
            } else if (!kMulti && !kExp3) {
                // (four path entries at a time: their loads are in flight together)
                uint32_t i = 0;
                for (; i + 4 <= depth; i += 4) {
                    const uint32_t a = path[i], b = path[i + 1], c = path[i + 2], d = path[i + 3];
                    update_node(a & 0xFFFFFFu, a >> 24, 1.0); update_node(b & 0xFFFFFFu, b >> 24, 1.0);
                    update_node(c & 0xFFFFFFu, c >> 24, 1.0); update_node(d & 0xFFFFFFu, d >> 24, 1.0);
                }
                for (; i < depth; ++i) update_node(path[i] & 0xFFFFFFu, path[i] >> 24, 1.0);
            } else {
                // MO-ISMCTS and EXP3 on private trees: the recorded path of every tree, level by level.  Round 1 walked
                // the parent links (load a node, find its parent in the key, load that ...): a chain of dependent
                // loads per tree, the largest single source of stalls of the MO kernel (profiles/r02_extra_benches.md).
                // Here the probabilities of a level's nodes (EXP3) are loaded together and the updates are reductions.
                for (uint32_t i = 0; i < depth; ++i) {
                    const uint32_t* const level = path + i * kTrees;
                    double prob[kTrees];
#pragma unroll
                    for (uint32_t t = 0; t < kTrees; ++t) {
                        const uint32_t n = level[t] & 0xFFFFFFu;
                        prob[t] = 1.0;
                        if constexpr (kExp3) { if (n >= tree.roots) prob[t] = load_prob(&tree.slots[n]); }
                    }
#pragma unroll
                    for (uint32_t t = 0; t < kTrees; ++t) {
                        const uint32_t n = level[t] & 0xFFFFFFu;
                        if (n >= tree.roots) update_node(n, level[t] >> 24, prob[t]);
                    }
                }
            }
        } else {
            P.playout_out[gtid] = (game().outcome() & 0xFFu) | ((plies & 0xFFFu) << 8) | ((rng.consumed() & 0xFFFu) << 20);
        }
        ++done;
        if (budget_left()) start_iteration();
        else phase = kDone;
    }

    // updateData on a shared node whose visit is already counted (ucb1.h:53-56, exp3.h:45-48).
    template <class S>
    ISMCTS_DEV static void add_reward_shared(S* s, uint32_t r2, double prob)
    {
        if constexpr (!S::kExp3) red_add_u64(reinterpret_cast<uint64_t*>(s), (uint64_t)r2 << 32);
        else atomic_add_f64(&s->score, ISMCTS_DDIV(ISMCTS_DMUL((double)r2, 0.5), prob > 0.0 ? prob : 1.0)); // (a node another lane is still filling in)
    }

    template <class S> ISMCTS_DEV static double slot_prob(const S& s) { if constexpr (S::kExp3) return s.prob; else return 1.0; }
    template <class S> ISMCTS_DEV static double load_prob(const S* s)
    {
        if constexpr (S::kExp3) return load_f64_cg(&s->prob); else return 1.0;
    }

    // Node::selectChild + UCB1::operator(), node.h:58-72, ucb1.h:69-76: every legal
    // child becomes available, then the first maximum in creation order.
    template <class S>
    ISMCTS_DEV uint32_t select_child(uint32_t parent, const Mask& legal, S& chosen, uint32_t& chosen_move)
    {
        for (Mask rest = legal; rest.any();) tree.prefetch(parent, rest.pop_lowest());
        if constexpr (!S::kExp3) {
            uint32_t best = 0;
            double best_u = -1.0;
            chosen.created = 0xFFFFFFFFu;
            // copies: what is read through `this` or P would be read again after every store to a node
            const TreePool<S> pool = tree;
            const UcbVisitFactors* const by_visits = P.visit_tab;
            const double* const sqrt_ln = P.sqrt_ln_tab;
            const double exploration = P.exploration;
            const uint32_t last = P.ln_count - 1u;                    // (never reached: budget_left())
            for (Mask rest = legal; rest.any();) {
                const uint32_t move = rest.pop_lowest();
                S c;
                const uint32_t ci = pool.find(parent, move, c);
                c.avail += 1;
                pool.slots[ci].avail = c.avail;
                const UcbVisitFactors f = load_factors_readonly(by_visits + (c.visits < last ? c.visits : last));
                const double u = ismcts_ucb_value(c.score2, f.inv, f.inv_sqrt, load_f64_readonly(sqrt_ln + (c.avail < last ? c.avail : last)), exploration);
                if (u > best_u || (u == best_u && c.created < chosen.created)) { best_u = u; best = ci; chosen = c; chosen_move = move; }
            }
            return best;
        } else {
            // EXP3::operator(), exp3.h:57-95: p_i = eps_t + (1 - K eps_t) exp(eps_{t-1} s_i) / sum_j exp(eps_{t-1} s_j),
            // stored in the nodes, then one sample.  Children are taken in ascending move order.
            uint32_t idx[Game::kMaxLegal];
            double val[Game::kMaxLegal];
            uint32_t K = 0;
            uint64_t t = 0;
            double smax = 0.0;
            for (Mask rest = legal; rest.any(); ++K) {
                const uint32_t move = rest.pop_lowest();
                S c;
                idx[K] = tree.find(parent, move, c);
                val[K] = c.score;
                t += c.visits;
                smax = (K == 0 || c.score > smax) ? c.score : smax;
            }
            const double lnK = P.ln_table[K];
            const double e_t = ismcts_exp3_epsilon(K, lnK, (double)t);
            const double e_tm1 = ismcts_exp3_epsilon(K, lnK, (double)(uint64_t)(t - 1));
            double sum = 0.0;
            for (uint32_t i = 0; i < K; ++i) {
                val[i] = ismcts_det_exp(ISMCTS_DMUL(e_tm1, ISMCTS_DSUB(val[i], smax)));
                sum = ISMCTS_DADD(sum, val[i]);
            }
            const double rest_mass = ISMCTS_DSUB(1.0, ISMCTS_DMUL((double)K, e_t));
            const double x = ISMCTS_DMUL((double)draw_word(), 1.0 / 4294967296.0);
            double cum = 0.0;
            uint32_t pick = K - 1;
            bool picked = false;
            for (uint32_t i = 0; i < K; ++i) {
                const double p = ISMCTS_DADD(e_t, ISMCTS_DDIV(ISMCTS_DMUL(rest_mass, val[i]), sum));
                tree.slots[idx[i]].prob = p;                  // setProbability, exp3.h:84
                cum = ISMCTS_DADD(cum, p);
                if (!picked && x < cum) { pick = i; picked = true; }
            }
            chosen = load_slot(&tree.slots[idx[pick]]);
            chosen_move = slot_move_of(chosen.key);
            return idx[pick];
        }
    }

    // select + expand (sosolver.h:53-71; mosolver.h:69-95) from the current cursor(s)
    // until a node is added, the game ends or a move triggers chance events (they
    // are drained by the random micro-steps of the following passes, then the
    // descent resumes).  MO: the tree of the player to move decides, then the
    // cursor of EVERY tree moves to the child for that move, which is created where
    // it does not exist yet (findOrAddChild, node.h:50-56).
    //
    // Every lane of the warp makes the call (`active`: the lane has a descent to make) and the warp goes down level
    // by level: on each level the lanes first decide — some select, some expand, their paths through that code
    // differ — then, together again, make their moves.  Left to itself the compiler joins the expanding and the
    // selecting lanes only at the end of the descent and runs the move once for each group with a handful of lanes
    // (7 of 32 measured, profiles/r02_level_sync.md); the barrier between the two halves of a level keeps them one
    // instruction stream.  A tree is not affected: each lane still performs the sequential operations of its tree.
    ISMCTS_DEV_CALL void descend(bool active)
    {
        const uint32_t ahead = expansion_word_ahead(active);
        for (;;) {
            if (!warp_any(active)) return;
            uint32_t move = 0;
            bool expand = false;
            if (active) {
                const Mask legal = game().legal();
                if (!legal.any()) {
                    active = false;                                   // terminal: selectNode, solverbase.h:53-56
                } else {
                    const uint32_t who = game().current_player();          // newChild, solverbase.h:74-80: before the move
                    const uint32_t a = kMulti ? who : 0;                  // the deciding tree
                    const Mask untried = legal.minus(node_children[a]);   // untriedMoves, node.h:82-91
                    expand = untried.any();
                    uint32_t chosen_slot = 0;
                    Slot chosen;
                    chosen.child_mask = Mask::none();
                    if (expand) move = untried.kth(draw_expansion(untried.count(), ahead));  // expand, sosolver.h:63-71
                    else chosen_slot = select_child(node[a], legal, chosen, move);
                    // the reward is credited from the point of view of the player stored in the node (ucb1.h:53-56),
                    // which is the mover when the node was created — in games where the same move can be made by
                    // different players in different determinisations (phantom m,n,k) not necessarily today's mover
                    uint32_t* const level = path + depth * kTrees;
                    for (uint32_t t = 0; t < kTrees; ++t) {
                        if (kMulti && !((tree_players >> t) & 1u)) { level[t] = t; continue; }   // (a root: no update)
                        uint32_t stored = who;
                        if (t == a && !expand) {
                            node[t] = chosen_slot; node_children[t] = chosen.child_mask; stored = chosen.player;
                        } else if (node_children[t].test(move)) {         // MO: the child exists in this tree
                            Slot c;
                            node[t] = tree.find(node[t], move, c);
                            node_children[t] = c.child_mask; stored = c.player;
                        } else {
                            const Mask grown = node_children[t].with(move);
                            tree.slots[node[t]].child_mask = grown;
                            if (node[t] < kTrees) root_children[node[t]] = grown;
                            node[t] = tree.insert(node[t], move, who, tree.count);
                            node_children[t] = Mask::none();
                        }
                        level[t] = node[t] | (stored << 24);
                    }
                    ++depth;
                }
            }
            warp_sync();
            if (active) {
                game().step(move);
                if (expand) { phase = kRollout; active = false; }
                else if (game().chance_pending()) active = false;
            }
        }
    }

    // The one draw a descent of SO-ISMCTS with UCB1 makes is the pick of its expansion.  In Knockout Whist the word it
    // comes from is always the first of a fresh Philox block (every phase leaves the stream aligned), generated in
    // registers: the lanes generate it here, together, before they part ways in the tree, instead of each group of
    // expanding lanes for itself on its level.  Other games and policies draw when they get there (0 is returned
    // and not used).
    static constexpr bool kWordAhead = game_is_phased<Game>(0) && !kExp3 && !kMulti;
    ISMCTS_DEV uint32_t expansion_word_ahead(bool active) const
    {
        if constexpr (kWordAhead) return active ? rng.peek_word_direct() : 0u;
        else return 0u;
    }
    ISMCTS_DEV uint32_t draw_expansion(uint32_t n, uint32_t ahead)
    {
        if constexpr (kWordAhead) return rng.below_with_word(ahead, n);
        else return draw_below(n);
    }

    // A visit of `slot` counted now (virtual loss): the reward follows at backpropagation.
    ISMCTS_DEV void count_visit(uint32_t slot)
    {
        if constexpr (!kExp3) red_add_u64(reinterpret_cast<uint64_t*>(&tree.slots[slot]), 1u);
        else red_add_u32(&tree.slots[slot].visits, 1u);
    }

    // UCB1 on a node whose children other lanes update (selectChild as in select_child()).
    template <class S>
    ISMCTS_DEV uint32_t select_child_shared(uint32_t parent, const Mask& legal, uint32_t& chosen_move, uint32_t& chosen_player,
                                            Mask& chosen_children)
    {
        if constexpr (!S::kExp3) {
            for (Mask rest = legal; rest.any();) tree.prefetch(parent, rest.pop_lowest());
            double best_u = -1.0;
            uint32_t best_created = 0xFFFFFFFFu, best = 0;
            const TreePool<S> pool = tree;
            const UcbVisitFactors* const by_visits = P.visit_tab;
            const double* const sqrt_ln = P.sqrt_ln_tab;
            const double exploration = P.exploration;
            const uint32_t last = P.ln_count - 1u;
            // The children one after the other, two loads ahead of the arithmetic: while the value of a child is
            // computed its table factors (tables of tens of thousands of entries for a deep tree: L2 accesses that depend
            // on the slot just read) and the slot of the NEXT child are in flight.  Evaluated strictly one after the
            // other, a level cost two dependent L2 round trips per child — 57 % of this kernel's long-scoreboard
            // stalls (profiles/r02_hybrid.md).
            Mask rest = legal;
            uint32_t m = rest.pop_lowest();
            uint32_t i = pool.home_slot(parent, m);
            S c = load_slot(&pool.slots[i]);
            for (;;) {
                const uint32_t ci = pool.resolve(parent, m, i, c);
                const uint32_t avail = c.avail + 1u < last ? c.avail + 1u : last;
                const double sqrt_ln_avail = load_f64_readonly(sqrt_ln + avail);
                const UcbVisitFactors f = load_factors_readonly(by_visits + (c.visits < last ? c.visits : last));
                const bool more = rest.any();
                uint32_t m_next = 0, i_next = 0;
                S c_next = c;
                if (more) {
                    m_next = rest.pop_lowest();
                    i_next = pool.home_slot(parent, m_next);
                    c_next = load_slot(&pool.slots[i_next]);
                }
                red_add_u32(&pool.slots[ci].avail, 1u);                    // markAvailable, ucb1.h:71-72
                const double u = ismcts_ucb_value(c.score2, f.inv, f.inv_sqrt, sqrt_ln_avail, exploration);
                if (u > best_u || (u == best_u && c.created < best_created)) {
                    best_u = u; best_created = c.created; best = ci; chosen_move = m; chosen_player = c.player;
                    chosen_children = c.child_mask;
                }
                if (!more) break;
                m = m_next; i = i_next; c = c_next;
            }
            return best;
        } else {
            // EXP3 reads the scores and visits the other lanes are updating and stores its probabilities,
            // as the reference does with its atomics (exp3.h:42-48,84)
            S chosen;
            const uint32_t ci = select_child(parent, legal, chosen, chosen_move);
            chosen_player = chosen.player;
            chosen_children = chosen.child_mask;
            return ci;
        }
    }

    // select + expand on trees shared with other lanes (tree parallelisation): SO-ISMCTS on one tree,
    // MO-ISMCTS on the trees of all players (the children of the other players' trees are found or
    // created for every move, findOrAddChild, node.h:50-56).
    ISMCTS_DEV_CALL void descend_shared(bool active)
    {
        if constexpr (kShared) {
            const uint32_t ahead = expansion_word_ahead(active);
            for (;;) {
                if (!warp_any(active)) return;
                uint32_t move = 0;
                bool expand = false;
                if (active) {
                    const Mask legal = game().legal();
                    if (!legal.any()) {
                        active = false;
                    } else {
                        const uint32_t who = game().current_player();
                        const uint32_t a = kMulti ? who : 0;                  // the deciding tree
                        // The children of the cursor: other lanes add children, so the mask is read from the node — except
                        // (SO) right after a selection, which has just read the chosen child's slot, mask included: a
                        // mask a few hundred cycles old instead of one more dependent round trip per level.  A child
                        // added since then is found by find_or_insert_shared() when this lane tries to create it.
                        const Mask kids = (!kMulti && mask_known) ? node_children[0] : Mask::load_shared(&tree.slots[node[a]].child_mask);
                        const Mask untried = legal.minus(kids);
                        expand = untried.any();
                        uint32_t chosen = 0, stored = who;
                        Mask chosen_children = Mask::none();
                        if (expand) move = untried.kth(draw_expansion(untried.count(), ahead));
                        else chosen = select_child_shared<Slot>(node[a], legal, move, stored, chosen_children);
                        if constexpr (!kMulti) { node_children[0] = chosen_children; mask_known = !expand; }
                        for (uint32_t t = 0; t < kTrees; ++t) {
                            if (kMulti && !((tree_players >> t) & 1u)) continue;
                            uint32_t next;
                            if (t == a && !expand) {
                                next = chosen;
                                count_visit(next);
                            } else {
                                // Creation order within a tree = (iteration, node of that iteration): unique without a
                                // counter the lanes would have to share (a second atomic round trip per expansion).
                                bool won;
                                next = tree.find_or_insert_shared(node[t], move, who, ((rng.c1 + 1u) << 8) | (created_seq & 0xFFu), won);
                                // count the visit before the child becomes visible: a published child has visits >= 1
                                count_visit(next);
                                if (won) {
                                    ++created_seq; ++my_nodes;
                                    Mask::set_shared(&tree.slots[node[t]].child_mask, move);   // publishes the child (a release)
                                } else {
                                    // another lane is creating this child: wait until it has published it
                                    while (!Mask::test_shared_acquire(&tree.slots[node[t]].child_mask, move)) {}
                                }
                            }
                            node[t] = next;
                        }
                        if (!kMulti && !kExp3) path[depth++] = node[0] | ((expand ? who : stored) << 24);
                    }
                }
                warp_sync();                                              // (see descend())
                if (active) {
                    game().step(move);
                    if (expand) { phase = kRollout; active = false; }
                    else if (game().chance_pending()) active = false;
                }
            }
        }
    }

    // ---- the lock-step loop (see the head of this file)
    //
    // The tree operations (descend, finish_iteration) are real calls: the lane object lives in local
    // memory and the registers those operations need (64-bit node fields, the double-precision
    // bandit values) do not count against the loops that run most of the time.  Those loops —
    // chance events and rollout moves — work on copies of the game and of the stream, which
    // are registers, between two tree operations.
    ISMCTS_DEV bool wants_descent() const { return phase == kSelect && !game().chance_pending() && !game().terminal(); }

    // `scratch` + `column`, `ring`: the columns the game and the stream are bound to, handed in again so
    // that the copies the loops work on address them as shared memory (the lane object itself lives in
    // local memory, where a pointer is just 64 bits).
    ISMCTS_DEV void run(uint32_t gtid, uint64_t* scratch, uint32_t scratch_stride, uint32_t column, uint32_t* ring, uint32_t ring_stride)
    {
        for (;;) {
            if (!warp_any(phase != kDone)) break;
            {
                typename GameHot<Game>::type hot(game());
                LaneRng stream = rng;
                hot.rebind(scratch, scratch_stride, column);
                stream.ring = ring; stream.stride = ring_stride;
                const uint32_t ph = phase;
                uint32_t moves = 0;
                for (;;) {
                    // every pending chance event of the warp
                    if constexpr (game_is_phased<Game>(0)) {
                        hot.chance_phase(stream, ph != kDone, [this] { return prep; });
                    } else {
                        for (uint32_t s = 0; warp_any(ph != kDone && hot.chance_pending()); ++s) {
                            if ((s & 3u) == 0u && ph != kDone && stream.low()) stream.refill();
                            if (ph != kDone && hot.chance_pending()) hot.chance_draw(stream.below(hot.chance_count()));
                        }
                    }
                    // the tree decides next, for some lane
                    if (kTree && warp_any(ph == kSelect && !hot.chance_pending() && !hot.terminal())) break;
                    // every game is over
                    if (!warp_any(ph == kRollout && !hot.chance_pending() && !hot.terminal())) break;
                    // rollout moves (simulate, solverbase.h:36-43) until a chance event or the end of the game
                    if constexpr (game_is_phased<Game>(0)) {
                        hot.play_phase(stream, ph == kRollout, moves);
                    } else {
                        for (uint32_t s = 0; warp_any(ph == kRollout && !hot.chance_pending() && !hot.terminal()); ++s) {
                            if ((s & 3u) == 0u && ph == kRollout && stream.low()) stream.refill();
                            if (ph == kRollout && !hot.chance_pending() && !hot.terminal()) {
                                const Mask set = hot.legal();
                                ++moves;
                                hot.play(set.kth(stream.below(set.count())));
                            }
                        }
                    }
                }
                game() = hot;
                rng = stream;
                plies += moves;
            }
            if (kTree) {
                const bool down = wants_descent();
                if (warp_any(down)) {
                    if (kShared) descend_shared(down); else descend(down);
                    continue;
                }
            }
            // every lane still at work has reached the end of its game
            if (phase != kDone) finish_iteration(gtid);
        }
    }
};

// The body of lane_tree_search on a lane object that lives wherever the caller put it.
template <class Game, class Slot, uint32_t kKind, bool kShared>
ISMCTS_DEV void lane_tree_search_on(Lane<Game, Slot, kKind, kShared>& lane, const SearchParams& P, uint32_t gtid, bool valid,
                                    uint64_t* scratch, uint32_t scratch_stride, uint32_t column, uint32_t* ring,
                                    uint32_t ring_stride, uint32_t* path, uint32_t table)
{
    using Mask = typename Game::Mask;
    using L = Lane<Game, Slot, kKind, kShared>;
    lane.path = path;
    lane.game().bind(scratch, scratch_stride, column);
    lane.phase = kDone; lane.done = 0; lane.quota = 0; lane.deadline = 0; lane.first_iteration = 0;
    lane.iteration_stride = 1; lane.my_nodes = 0; lane.created_seq = 0; lane.mask_known = 0;
    for (uint32_t t = 0; t < L::kTrees; ++t) lane.root_children[t] = Mask::none();
    uint32_t tree_index = 0;
    if (valid) {
        const uint32_t width = kShared ? P.lanes_per_tree : 1u;
        tree_index = gtid / width;                            // position * trees + tree
        const uint32_t lane_in_tree = gtid % width;
        const uint32_t pos = tree_index / P.trees;
        const uint32_t tree_global = P.tree_base + tree_index % P.trees;
        lane.prep = reinterpret_cast<const typename Game::Prepared*>(P.prepared) + pos;
        lane.tree.bind(reinterpret_cast<Slot*>(P.pool) + ((size_t)(kShared ? tree_index : table) << P.log2cap), P.log2cap, L::kTrees);
        if (!kShared) lane.tree.init_roots();                 // shared trees: init_shared_roots
        lane.tree_players = Game::root_players(lane.prep);
        lane.rng.init(P.seed, P.pos_base + pos, tree_global, ring, ring_stride);
        lane.rng.per_word = game_draws_per_word<Game>(0);
        if (P.iterations > 0) {
            // the iterations of a tree, dealt round-robin to its lanes
            const uint64_t of_tree = tree_iterations(P.iterations, P.trees_total, tree_global);
            lane.quota = (uint32_t)(of_tree / width + (lane_in_tree < of_tree % width ? 1u : 0u));
        } else {
            lane.deadline = (P.start_ns ? load_u64_cg(P.start_ns) : global_timer_ns()) + P.time_ns;    // time mode, utility.h:51-60
        }
        if (kShared) {
            lane.first_iteration = lane_in_tree; lane.iteration_stride = width;
        }
        if (lane.budget_left()) lane.start_iteration();
    }
    lane.run(gtid, scratch, scratch_stride, column, ring, ring_stride);
    if (!valid) return;
    if (kShared) atomic_add_u64(P.iters_done + tree_index, lane.done);
    else P.iters_done[tree_index] = lane.done;
}

// The whole search of lane `gtid` of this launch: one SO tree or one set of MO trees, or — tree
// parallelisation — one of the `lanes_per_tree` lanes of a shared tree (`valid` = false for the
// lanes that only pad the last warp: they take part in the warp votes and do nothing).  The
// statistics of the root children are collected afterwards by collect_roots.  `table`: the table of the
// pool a private tree is built in (a lane that searches several trees one after the other reuses its table).
template <class Game, class Slot, uint32_t kKind, bool kShared>
ISMCTS_DEV void lane_tree_search(const SearchParams& P, uint32_t gtid, bool valid, uint64_t* scratch,
                                 uint32_t scratch_stride, uint32_t column, uint32_t* ring, uint32_t ring_stride, uint32_t* path,
                                 uint32_t table)
{
    Lane<Game, Slot, kKind, kShared> lane(P);
    lane_tree_search_on<Game, Slot, kKind, kShared>(lane, P, gtid, valid, scratch, scratch_stride, column, ring, ring_stride, path, table);
}
template <class Game, class Slot, uint32_t kKind, bool kShared>
ISMCTS_DEV void lane_tree_search(const SearchParams& P, uint32_t gtid, bool valid, uint64_t* scratch,
                                 uint32_t scratch_stride, uint32_t column, uint32_t* ring, uint32_t ring_stride, uint32_t* path)
{
    lane_tree_search<Game, Slot, kKind, kShared>(P, gtid, valid, scratch, scratch_stride, column, ring, ring_stride, path, gtid);
}

// Shared trees: the root node(s) and the node counter of tree `index`, before the search starts.
template <class Slot>
ISMCTS_DEV void init_shared_root(const SearchParams& P, uint32_t index, uint32_t n_roots = 1)
{
    TreePool<Slot> tree;
    tree.bind(reinterpret_cast<Slot*>(P.pool) + ((size_t)index << P.log2cap), P.log2cap, n_roots);
    tree.init_roots();
}

// Root children of tree `index` -> per-position sums (RootParallel::compileVisitCounts,
// execution.h:219-231).  MO: the tree of the root's player to move decides (mosolver.h:42-46).
// `table`: where tree `index` was built.
template <class Game, class Slot, uint32_t kKind>
ISMCTS_DEV void collect_roots(const SearchParams& P, uint32_t index, uint32_t table)
{
    using Mask = typename Game::Mask;
    const uint32_t pos = index / P.trees;
    const typename Game::Prepared* prep = reinterpret_cast<const typename Game::Prepared*>(P.prepared) + pos;
    TreePool<Slot> tree;
    tree.bind(reinterpret_cast<Slot*>(P.pool) + ((size_t)table << P.log2cap), P.log2cap, kKind == kMO ? Game::kMaxPlayers : 1);
    const uint32_t root = kKind == kMO ? Game::root_player(prep) : 0;
    uint64_t* v = P.visits + (size_t)pos * Game::kMoveStride;
    uint64_t* s = P.score2 + (size_t)pos * Game::kMoveStride;
    uint64_t* a = P.avail + (size_t)pos * Game::kMoveStride;
    for (Mask rest = load_slot(&tree.slots[root]).child_mask; rest.any();) {
        const uint32_t move = rest.pop_lowest();
        Slot c;
        tree.find(root, move, c);
        atomic_add_u64(v + move, c.visits);
        if constexpr (!Slot::kExp3) {
            atomic_add_u64(s + move, c.score2);
            atomic_add_u64(a + move, c.avail);
        } else {
            // EXP3: fixed point, so that the sums do not depend on the order of the additions
            atomic_add_u64(s + move, (uint64_t)(long long)llrint(c.score * 65536.0));
            atomic_add_u64(a + move, (uint64_t)(long long)llrint(c.prob * 4294967296.0));
        }
    }
}

// Determinise + rollout only: playout `gtid % per_position` of position `gtid / per_position`.
template <class Game>
ISMCTS_DEV void lane_playout(const SearchParams& P, uint32_t gtid, bool valid, uint32_t per_position, uint64_t* scratch,
                             uint32_t scratch_stride, uint32_t column, uint32_t* ring, uint32_t ring_stride)
{
    using Mask = typename Game::Mask;
    Lane<Game, SlotUcb<Mask>, kPlayout> lane(P);
    lane.path = nullptr;
    lane.game().bind(scratch, scratch_stride, column);
    lane.phase = kDone; lane.done = 0; lane.quota = 1; lane.deadline = 0; lane.first_iteration = 0;
    lane.iteration_stride = 1; lane.my_nodes = 0; lane.created_seq = 0; lane.mask_known = 0;
    lane.root_children[0] = Mask::none();
    if (valid) {
        const uint32_t pos = gtid / per_position;
        lane.first_iteration = gtid % per_position;
        lane.prep = reinterpret_cast<const typename Game::Prepared*>(P.prepared) + pos;
        lane.rng.init(P.seed, P.pos_base + pos, 0, ring, ring_stride);
        lane.rng.per_word = game_draws_per_word<Game>(0);
        lane.start_iteration();
    }
    lane.run(gtid, scratch, scratch_stride, column, ring, ring_stride);
}

} // namespace ismcts
