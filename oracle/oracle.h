/*
 * oracle.h — CPU oracle for the ISMCTS hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithm (sfranzen/ismcsolver) for
 * the search iteration loop, driven by the same counter-based Philox4x32-10
 * streams as the CUDA kernels so the two can be compared bit for bit.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (ismcsolver_b200/,
 * include/ismcts/) never does.
 *
 * Parity status: the game rules, node/bandit semantics and full-search
 * behaviour are PINNED against the reference's own known-answer tests
 * (tests/test_oracle_kat.py) and, statistically, against the real reference
 * compiled into oracle/_ref/ (tests/golden/ref_stats_*.json).  The random
 * streams themselves cannot be pinned: the reference draws from unseeded
 * std::mt19937 + implementation-defined std::shuffle (utility.h:62-74,
 * knockoutwhist.cpp:28-32,138-140).
 *
 * The POD state layouts are shared with the product through
 * include/ismcts_b200.h (a header only; no product code is linked here).
 */
#ifndef ISMCTS_ORACLE_H
#define ISMCTS_ORACLE_H

#include <stdint.h>
#include "../include/ismcts_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Philox4x32-10 block (Salmon et al. 2011, "Parallel random numbers: as easy as 1, 2, 3"). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* Word `d` of the stream (seed; pos, tree, iteration). */
uint32_t orc_draw_word(uint64_t seed, uint32_t pos, uint32_t tree, uint32_t iteration, uint32_t d);

/* Game helpers on the POD states (same contracts as the ismcts_kw_ / ismcts_mnk_ helpers). */
int orc_kw_deal(uint64_t seed, uint32_t stream, uint32_t num_players, ismcts_kw_state* out);
uint64_t orc_kw_legal(const ismcts_kw_state* s);
int orc_kw_apply(ismcts_kw_state* s, uint32_t move, uint64_t seed, uint32_t stream, uint32_t* draw_counter);
int orc_kw_result2(const ismcts_kw_state* s, uint32_t player);
int orc_kw_determinise(ismcts_kw_state* s, uint32_t observer, uint64_t seed, uint32_t pos, uint32_t tree,
                       uint32_t iteration, uint32_t* draw_counter);
int orc_mnk_init(uint32_t m, uint32_t n, uint32_t k, ismcts_mnk_state* out);
void orc_mnk_legal(const ismcts_mnk_state* s, uint64_t legal[4]);
int orc_mnk_apply(ismcts_mnk_state* s, uint32_t move);
int orc_mnk_result2(const ismcts_mnk_state* s, uint32_t player);

int orc_phantom_init(uint32_t m, uint32_t n, uint32_t k, ismcts_phantom_state* out);
void orc_phantom_legal(const ismcts_phantom_state* s, uint64_t legal[4]);
int orc_phantom_apply(ismcts_phantom_state* s, uint32_t move);
int orc_phantom_result2(const ismcts_phantom_state* s, uint32_t player);
int orc_goof_init(uint64_t seed, uint32_t stream, ismcts_goof_state* out);
uint64_t orc_goof_legal(const ismcts_goof_state* s);
int orc_goof_apply(ismcts_goof_state* s, uint32_t move);
int orc_goof_result2(const ismcts_goof_state* s, uint32_t player);

/* One determinise + rollout (no tree); result packed as ismcts_playouts does. */
uint32_t orc_playout(uint32_t game, const void* root, uint64_t seed, uint32_t pos, uint32_t tree,
                     uint32_t iteration);

/* Work counters accumulated by the searches (SURVEY 8d "algorithmic work per unit"). */
typedef struct orc_counters {
    uint64_t iterations;
    uint64_t plies;          /* state transitions (doMove) */
    uint64_t draws;          /* random words consumed */
    uint64_t chance_picks;   /* cards dealt + tie-breaks */
    uint64_t select_steps;   /* tree levels descended through selectChild */
    uint64_t child_evals;    /* UCB/EXP3 evaluations */
    uint64_t nodes_created;
    uint64_t backprop_updates;
} orc_counters;

/*
 * One tree, `iterations` sequential iterations (SOSolver::search,
 * sosolver.h:44-71).  Writes the tree in creation order.  Returns 0, or
 * ISMCTS_ERR_CAPACITY when max_nodes is too small.
 */
int orc_so_search(uint32_t game, const void* root, uint32_t policy, double exploration, uint64_t seed,
                  uint32_t pos, uint32_t tree, uint64_t iterations, ismcts_node_rec* nodes,
                  uint32_t max_nodes, uint32_t* n_nodes, orc_counters* counters);

/*
 * MO-ISMCTS (MOSolver::search, mosolver.h:57-101): one tree per player.  The
 * trees are written back to back: nodes of player p's tree start at
 * tree_offsets[p] and there are tree_counts[p] of them (p < num players).
 */
int orc_mo_search(uint32_t game, const void* root, uint32_t policy, double exploration, uint64_t seed,
                  uint32_t pos, uint32_t tree, uint64_t iterations, ismcts_node_rec* nodes,
                  uint32_t max_nodes, uint32_t tree_offsets[8], uint32_t tree_counts[8],
                  orc_counters* counters);

/*
 * Root-parallel search of one position (RootParallel::execute + bestMove,
 * execution.h:193-231): trees [tree_base, tree_base+trees) of trees_total,
 * `iterations` total per position split as ismcts_search_desc documents.
 * Sums root-child statistics into visits/score2/avail[stride]; returns the
 * best move in *best_move (first maximum, ascending move order).
 */
int orc_root_parallel(uint32_t game, uint32_t solver, const void* root, uint32_t policy, double exploration,
                      uint64_t seed, uint32_t pos, uint32_t tree_base, uint32_t trees, uint32_t trees_total,
                      uint64_t iterations, uint64_t* visits, uint64_t* score2, uint64_t* avail,
                      uint32_t* best_move, orc_counters* counters);

/* Per-tree iteration split used by count mode. */
uint64_t orc_tree_iterations(uint64_t iterations, uint32_t trees_total, uint32_t tree_global);

/* The UCB1 value exactly as evaluated (ucb1.h:36-39, utility.h:96-99). */
double orc_ucb(uint32_t score2, uint32_t visits, uint32_t avail, double exploration);

/* EXP3 probabilities (exp3.h:73-95) for K children with the given scores/visits. */
void orc_exp3_probabilities(const double* scores, const uint32_t* visits, uint32_t K, double* p_out);

#ifdef __cplusplus
}
#endif
#endif
