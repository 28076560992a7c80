/*
 * ismcts_b200.h — C ABI of the B200-native ISMCTS search engine.
 *
 * This is the drop-in boundary for ONE hot path of sfranzen/ismcsolver: the
 * search iteration loop (determinise -> select -> expand -> rollout ->
 * backpropagate) that the reference runs on std::async workers.  Every entry
 * point below cites the reference interface it replaces (paths relative to the
 * reference checkout).  Plain C types only: no C++ types, no exceptions, no
 * torch types.  Buffers are caller-owned.  A context is not thread-safe (the
 * reference solver object is not either: execution.h:143-154 shares m_counter).
 *
 * There is no CPU fallback: every compute entry point fails with
 * ISMCTS_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef ISMCTS_B200_H
#define ISMCTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
enum {
    ISMCTS_OK = 0,
    ISMCTS_ERR_NO_DEVICE = 1,   /* no CUDA device / CUDA runtime failure at create */
    ISMCTS_ERR_INVALID = 2,     /* bad descriptor (sizes, kinds, null pointers) */
    ISMCTS_ERR_ILLEGAL_MOVE = 3,/* host game helpers: move not legal (reference throws std::out_of_range,
                                   knockoutwhist.cpp:105-109, mnkgame.cpp:41-43) */
    ISMCTS_ERR_CUDA = 4,        /* a CUDA call failed; see ismcts_last_error */
    ISMCTS_ERR_CAPACITY = 5     /* node pool / ln-table capacity exceeded */
};

/* ------------------------------------------------------------------- games */
enum { ISMCTS_GAME_KW = 0, ISMCTS_GAME_MNK = 1, ISMCTS_GAME_GOOF = 2, ISMCTS_GAME_PHANTOM = 3 };
enum { ISMCTS_SOLVER_SO = 0, ISMCTS_SOLVER_MO = 1 };
enum { ISMCTS_POLICY_UCB1 = 0, ISMCTS_POLICY_EXP3 = 1 };

#define ISMCTS_KW_MAX_PLAYERS 7
#define ISMCTS_KW_CARDS 52
#define ISMCTS_MNK_MAX_CELLS 256

/*
 * Knockout Whist position, POD mirror of the reference's KnockoutWhist members
 * (test/common/knockoutwhist.h:32-48).  A card is the integer
 * id = rank + 13*suit, rank 0..12 (Two..Ace), suit 0..3 (clubs, diamonds,
 * hearts, spades) — the reference's `explicit operator int` (card.h:38-41).
 * A set of cards is a 64-bit mask with bit `id` set.
 */
typedef struct ismcts_kw_state {
    uint64_t hands[ISMCTS_KW_MAX_PLAYERS]; /* m_playerCards */
    uint64_t unknown;                      /* m_unknownCards: undealt and unseen by everyone */
    uint8_t num_players;                   /* m_numPlayers, 2..7 */
    uint8_t player;                        /* m_player: to move */
    uint8_t dealer;                        /* m_dealer */
    uint8_t trump;                         /* m_trumpSuit */
    uint8_t tricks_left;                   /* m_tricksLeft: 7..1, 0 = game over */
    uint8_t request_trump;                 /* m_requestTrump */
    uint8_t alive_mask;                    /* m_players as a bit set */
    uint8_t trick_len;                     /* m_currentTrick.size() */
    uint8_t tricks_taken[8];               /* m_tricksTaken (index 7 unused) */
    uint8_t trick_player[8];               /* m_currentTrick[i].first, play order */
    uint8_t trick_card[8];                 /* m_currentTrick[i].second */
} ismcts_kw_state;                         /* 96 bytes */

/*
 * m,n,k-game position (test/common/mnkgame.h:32-39).  Board is n rows by m
 * columns; move = cell index, row = move / m, column = move % m
 * (mnkgame.cpp:81-89).  stones[p] bit `cell` set = player p occupies it.
 * result2 = 2 * m_result from player 0's view, or -1 while running.
 */
typedef struct ismcts_mnk_state {
    uint64_t stones[2][4];
    uint8_t m, n, k;
    uint8_t player;
    int8_t result2;
    uint8_t pad[3];
} ismcts_mnk_state;                        /* 72 bytes */

/*
 * Two-player Goofspiel position (test/common/goofspiel.h:26-49): prize suit hearts, player 0 bids
 * spades, player 1 clubs, player 2 is the environment that turns over the prizes and reveals the bids.
 * Every move is simultaneous (goofspiel.cpp:80-83), so searches use the simultaneous-move policy
 * (EXP3, config.h:20-26).  Ranks are 0..12 (Two..Ace, card.h:13-15); a card's value is Ace 1, Two 2,
 * ... King 13 (goofspiel.cpp:17-36).  Move ids are card ids: a bid of player 0 is rank + 39, of
 * player 1 rank + 0, a prize rank + 26, and the environment's reveal move is Card{} = 0
 * (goofspiel.cpp:140-146).
 */
typedef struct ismcts_goof_state {
    uint64_t prize_order;   /* 4 bits per prize not turned over yet, entry n_prizes-1 is next (m_prizes.back()) */
    uint16_t hands[2];      /* bit r: the player still holds rank r (m_hands) */
    uint8_t n_prizes;
    uint8_t player;         /* m_player: 0, 1 or 2 */
    uint8_t draw_prize;     /* m_drawPrize */
    uint8_t current_prize;  /* rank of m_currentPrize */
    uint8_t moves[2];       /* ranks bid this turn (m_moves) */
    uint8_t scores[2];      /* m_scores */
    uint8_t pad[4];
} ismcts_goof_state;        /* 24 bytes */

/*
 * Phantom m,n,k-game position (test/common/phantommnkgame.h:12-36): the m,n,k-game in which the players
 * do not see each other's stones.  avail[p] are the cells player p still believes free (m_available);
 * trying an occupied one only removes it from that set and the same player moves again
 * (phantommnkgame.cpp:83-105).
 */
typedef struct ismcts_phantom_state {
    uint64_t stones[2][4];
    uint64_t avail[2][4];
    uint8_t m, n, k;
    uint8_t player;
    int8_t result2;
    uint8_t pad[3];
} ismcts_phantom_state;    /* 136 bytes */

/* Size in bytes of one root position record for a game kind (0 if unknown). */
size_t ismcts_state_size(uint32_t game);
/* Number of move ids the per-position output arrays are strided by:
 * 64 for Knockout Whist and Goofspiel (52 used), 256 for m,n,k and phantom m,n,k. */
uint32_t ismcts_move_stride(uint32_t game);

/* ---------------------------------------------------------- host game helpers
 * From-scratch host implementations of the reference's Game<Move> interface
 * (include/ismcts/game.h:15-46) on the POD states.  They run on the host, use
 * no device, and exist so callers (the C++ policies, the bench's synthetic
 * deal/position streams) can build and advance root positions.
 */

/* KnockoutWhist::KnockoutWhist(players) (knockoutwhist.cpp:36-54) with the
 * deal drawn from Philox stream `stream` of `seed` (SURVEY 8d "deal stream"). */
int ismcts_kw_deal(uint64_t seed, uint32_t stream, uint32_t num_players, ismcts_kw_state* out);
/* validMoves (knockoutwhist.cpp:122-136): returns the legal set as a mask. */
uint64_t ismcts_kw_legal(const ismcts_kw_state* s);
/* doMove (knockoutwhist.cpp:95-115); chance events (tie-break, re-deal) draw
 * from Philox(seed, stream, *draw_counter ...), advancing *draw_counter. */
int ismcts_kw_apply(ismcts_kw_state* s, uint32_t move, uint64_t seed, uint32_t stream, uint32_t* draw_counter);
/* getResult*2 (knockoutwhist.cpp:117-120). */
int ismcts_kw_result2(const ismcts_kw_state* s, uint32_t player);

/* MnkGame::MnkGame(m,n,k) (mnkgame.cpp:24-32). */
int ismcts_mnk_init(uint32_t m, uint32_t n, uint32_t k, ismcts_mnk_state* out);
/* validMoves (mnkgame.cpp:61-64): legal[0..3] = mask of empty cells, or all 0 when finished. */
void ismcts_mnk_legal(const ismcts_mnk_state* s, uint64_t legal[4]);
/* doMove (mnkgame.cpp:39-54). */
int ismcts_mnk_apply(ismcts_mnk_state* s, uint32_t move);
/* getResult*2 (mnkgame.cpp:56-59). */
int ismcts_mnk_result2(const ismcts_mnk_state* s, uint32_t player);

/* Goofspiel::Goofspiel (goofspiel.cpp:52-63): full hands, the prizes shuffled by Philox stream `stream`. */
int ismcts_goof_init(uint64_t seed, uint32_t stream, ismcts_goof_state* out);
/* validMoves (goofspiel.cpp:130-153) as a mask of move ids. */
uint64_t ismcts_goof_legal(const ismcts_goof_state* s);
/* doMove (goofspiel.cpp:90-117). */
int ismcts_goof_apply(ismcts_goof_state* s, uint32_t move);
/* getResult*2 (goofspiel.cpp:119-128). */
int ismcts_goof_result2(const ismcts_goof_state* s, uint32_t player);

/* PhantomMnkGame::PhantomMnkGame(m,n,k) (phantommnkgame.cpp:12-19). */
int ismcts_phantom_init(uint32_t m, uint32_t n, uint32_t k, ismcts_phantom_state* out);
/* validMoves (phantommnkgame.cpp:107-110): the cells the player to move believes free. */
void ismcts_phantom_legal(const ismcts_phantom_state* s, uint64_t legal[4]);
/* doMove (phantommnkgame.cpp:83-105). */
int ismcts_phantom_apply(ismcts_phantom_state* s, uint32_t move);
/* getResult*2 (mnkgame.cpp:56-59). */
int ismcts_phantom_result2(const ismcts_phantom_state* s, uint32_t player);

/* ------------------------------------------------------------------ context */
typedef struct ismcts_ctx ismcts_ctx;

/* Creates a search context on CUDA device `device`.  Replaces nothing in the
 * reference (its workers are std::async threads, execution.h:112-124); it owns
 * the device node pools, the ln() table and a stream. */
int ismcts_create(int device, ismcts_ctx** out);
void ismcts_destroy(ismcts_ctx* ctx);
/* Last error text of a context (or of the failed create when ctx == NULL). */
const char* ismcts_last_error(const ismcts_ctx* ctx);
/* Use an external CUDA stream (a cudaStream_t passed as void*) for all work of
 * this context; NULL restores the context's own stream. */
int ismcts_set_stream(ismcts_ctx* ctx, void* cuda_stream);

/* ------------------------------------------------------------------- search */
/*
 * One search request = what ExecutionPolicy::execute + bestMove do for one
 * operator() call (sosolver.h:30-36, execution.h:193-231), batched over
 * `n_positions` independent root positions.
 *
 * Root-parallel layout: every position is searched by `trees` independent
 * trees (reference: one per std::async worker, execution.h:193-204).  Tree j of
 * position i has the global id  tree_base + j  and draws all its randomness
 * from Philox4x32-10 keyed by `seed` with counter
 * (draw/4, iteration, global tree id, pos_base + i), so the result does not
 * depend on how trees are spread over threads, CTAs or GPUs.
 *
 * Budget: count mode (iterations > 0): `iterations` is the TOTAL per position,
 * as in the reference (execution.h:42-52,143-154); tree j runs
 * iterations/trees_total (+1 when j_global < iterations % trees_total)
 * iterations, where trees_total is the number of trees of the position over
 * all devices.  Time mode (iterations == 0): every tree iterates until
 * `time_ns` nanoseconds have passed on the device clock (utility.h:51-60).
 */
/* Tree parallelisation (TreeParallel, execution.h:169-179), alone or inside root parallelisation: EACH of the
 * `trees` trees of a position is searched by `lanes` lanes together — iterations in flight on one tree, with virtual
 * loss; lanes in bits 8..31 (bits 3..7 must be zero).  trees = 1: the reference's TreeParallel.  trees = 16, lanes = 128:
 * the geometry of the reference's RootParallel on a 16-thread host with a block of lanes on every tree (what bench.py
 * measures).  Keep iterations / trees / lanes >= 128 (profiles/r02_tree_parallel_quality.md).  With lanes = 1 the search
 * equals the sequential one bit for bit; with more the interleaving inside a tree is the hardware's. */
#define ISMCTS_FLAG_SHARED_TREE 1u
#define ISMCTS_FLAG_LANES_SHIFT 8
/* Knockout Whist: the caller guarantees that no position has more than four players, which lets the
 * search keep more games per SM (ismcts_search sets it itself from the host records; callers of
 * ismcts_search_device may set it). */
#define ISMCTS_FLAG_KW_MAX4 2u
/* Keep every tree of the search in device memory for ismcts_dump_tree afterwards.  Without it, a count-mode
 * search with more private trees than the GPU holds lanes lets its warps claim trees one after the other and
 * reuse their node tables (memory then does not grow with n_positions * trees), and only the root statistics
 * survive.  A descriptor that asks for a dump (dump_tree != 0xffffffff) implies it. */
#define ISMCTS_FLAG_KEEP_TREES 4u

typedef struct ismcts_search_desc {
    uint32_t game;          /* ISMCTS_GAME_* */
    uint32_t solver;        /* ISMCTS_SOLVER_* (sosolver.h / mosolver.h) */
    uint32_t policy;        /* ISMCTS_POLICY_* for sequential moves (config.h:20-26) */
    uint32_t n_positions;
    uint32_t trees;         /* trees per position searched by THIS call (this device's shard) */
    uint32_t trees_total;   /* trees per position over all devices (>= tree_base + trees) */
    uint32_t tree_base;     /* global id of this call's first tree */
    uint32_t pos_base;      /* global id of this call's first position */
    uint64_t iterations;    /* total per position, 0 = time mode */
    uint64_t time_ns;       /* time mode budget */
    uint64_t seed;          /* Philox key */
    double exploration;     /* UCB1 exploration constant (ucb1.h:65-67), default 0.7 */
    uint32_t dump_tree;     /* 0xffffffff = none; else local tree index (of position 0..) to dump:
                               dump index = position * trees + tree */
    uint32_t flags;         /* 0 (ISMCTS_FLAG_*) */
} ismcts_search_desc;

/* One tree node as the reference keeps it (tree/node.h:113-118, ucb1.h:50-51),
 * in creation order; index 0 is the root. */
typedef struct ismcts_node_rec {
    uint32_t parent;   /* creation index of the parent (root: 0) */
    uint32_t move;     /* m_move */
    uint32_t player;   /* m_playerJustMoved */
    uint32_t visits;   /* m_visits */
    uint32_t score2;   /* UCB1: 2 * m_score (rewards are 0, 1/2, 1); EXP3: 0 */
    uint32_t avail;    /* UCB1: m_available (ucb1.h:51); EXP3: 1 */
    double score;      /* m_score as a double (EXP3: sum of reward / probability, exp3.h:45-48) */
    double prob;       /* EXP3: m_probability (exp3.h:42); UCB1: 1 */
} ismcts_node_rec;     /* 40 bytes */

typedef struct ismcts_search_out {
    /* Per position, strided by ismcts_move_stride(game): root-child statistics
     * summed over this call's trees — RootParallel::compileVisitCounts
     * (execution.h:219-231) plus the reward and availability sums. */
    uint64_t* visits;       /* [n_positions * stride] */
    uint64_t* score2;       /* [n_positions * stride], UCB1: 2*score; EXP3: score (exp3.h:45-48) in 48.16 fixed point */
    uint64_t* avail;        /* [n_positions * stride], UCB1: availability; EXP3: probability in 32.32 fixed point */
    uint64_t* iterations_done; /* [n_positions * trees]: iterations each tree ran */
    uint32_t* best_move;    /* [n_positions]: first maximum of visits in ascending move order
                               (execution.h:209-216); 0xffffffff when the root is terminal */
    /* Optional tree dump (desc.dump_tree != 0xffffffff). */
    ismcts_node_rec* dump_nodes; /* capacity dump_capacity records */
    uint32_t dump_capacity;
    uint32_t dump_count;    /* out: nodes in the dumped tree */
} ismcts_search_out;

/* Host-buffer entry: `roots` points to n_positions packed state records on the
 * host.  Copies them to the device, runs the search, copies results back. */
int ismcts_search(ismcts_ctx* ctx, const ismcts_search_desc* desc, const void* roots,
                  ismcts_search_out* out);

/* Device-resident entries (for callers that keep positions in HBM and overlap
 * copies themselves).  d_roots, d_visits, d_score2, d_avail, d_iters are DEVICE
 * pointers laid out as in ismcts_search_out.  Launch is asynchronous on the
 * context's stream.  The records at d_roots cannot be checked from the host: the caller guarantees
 * what ismcts_search checks (player, dealer, trick entries below num_players <= 7, hands of at most
 * seven existing cards, boards of at most 256 cells; ISMCTS_FLAG_KW_MAX4 only with num_players <= 4). */
int ismcts_search_device(ismcts_ctx* ctx, const ismcts_search_desc* desc, const void* d_roots,
                         uint64_t* d_visits, uint64_t* d_score2, uint64_t* d_avail,
                         uint64_t* d_iters);
int ismcts_synchronize(ismcts_ctx* ctx);

/* Several devices: the `desc->trees` trees of the request (global ids tree_base ...) are split over the `n_ctx`
 * contexts, one CUDA device each, searched concurrently and the root arrays summed on the host —
 * RootParallel::compileVisitCounts (execution.h:219-231) across devices.  The result is that of ismcts_search
 * with the same descriptor on one device (tree ids are global).  `out->iterations_done` is
 * [n_positions * desc->trees] as for one device.  Errors are reported on ctxs[0].  No tree dump. */
int ismcts_search_multi(ismcts_ctx* const* ctxs, uint32_t n_ctx, const ismcts_search_desc* desc, const void* roots,
                        ismcts_search_out* out);

/* Reads back tree `dump_index` (= position * trees + tree of the last search)
 * as node records in creation order. */
int ismcts_dump_tree(ismcts_ctx* ctx, uint32_t dump_index, ismcts_node_rec* nodes,
                     uint32_t capacity, uint32_t* count);

/* Same for MO-ISMCTS searches: the tree of `player` (mosolver.h:31-34 keeps one tree per player;
 * ismcts_dump_tree reads the one that decides, the tree of the root's player to move). */
int ismcts_dump_tree_of(ismcts_ctx* ctx, uint32_t dump_index, uint32_t player, ismcts_node_rec* nodes,
                        uint32_t capacity, uint32_t* count);

/* ------------------------------------------------------------------ playouts
 * Determinise + uniform random rollout only (SolverBase::simulate,
 * solverbase.h:36-43, after cloneAndRandomise) — no tree.  Playout j of
 * position i uses the Philox stream (pos = i, tree = 0, iteration = j).
 * out_result[i * playouts_per_position + j] packs
 *   winner (bits 0..7: KW sole survivor id; m,n,k: result2 of player 0),
 *   plies  (bits 8..19), draws consumed (bits 20..31).
 */
int ismcts_playouts(ismcts_ctx* ctx, uint32_t game, const void* roots, uint32_t n_positions,
                    uint32_t playouts_per_position, uint64_t seed, uint32_t* out_result);

/* Number of kernels this context has launched so far (bench: gpu_launches). */
uint64_t ismcts_launch_count(const ismcts_ctx* ctx);
/* Library version string. */
const char* ismcts_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ISMCTS_B200_H */
