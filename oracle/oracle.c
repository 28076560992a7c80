/*
 * oracle.c — CPU oracle for the ISMCTS hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h for who may load it and for the parity status).
 *
 * Written from the reference's behaviour, not from its code: games keep
 * explicit membership arrays (the CUDA side uses bit masks), trees are arrays
 * of nodes with child lists in creation order (the CUDA side uses a hash pool).
 * Every function cites the reference lines whose behaviour it restates; paths
 * are relative to the reference checkout.
 *
 * Build: gcc -O2 -ffp-contract=off -std=c11 -shared -fPIC (see Makefile).
 */
#include "oracle.h"
#include "../include/ismcts_detmath.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ================================================================== Philox */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

uint32_t orc_draw_word(uint64_t seed, uint32_t pos, uint32_t tree, uint32_t iteration, uint32_t d)
{
    const uint32_t ctr[4] = { d >> 2, iteration, tree, pos };
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    return out[d & 3];
}

/* A random stream position.  The words of one iteration are d = 0, 1, 2, ... of the stream
 * (seed; pos, tree, iteration).  A word serves several consecutive draws: "uniform integer below
 * n" is the high half of word * n and the low half is the word of the next draw.  TWO draws per
 * word in general (the deviation from uniform of the pair is below n1 * n2 / 2^32 < 1e-6 for the
 * choices of m,n,k, n <= 225); THREE in Knockout Whist, whose choices are among at most 52 (the
 * triple deviates by less than n1 * n2 * n3 / 2^32 < 3.3e-5, a rollout triple, n <= 7, by less than
 * 1e-7).  rng_align() moves on to
 * the start of the next Philox block of four words: Knockout Whist aligns the stream around
 * every sequence of chance events and at the start of the rollout, so that all playouts of a
 * position consume their words in the same rhythm (on the GPU: the lanes of a warp generate
 * their blocks in the same steps). */
typedef struct rng {
    uint64_t seed;
    uint32_t pos, tree, iteration, d;
    orc_counters* cnt;
    uint32_t w, left;      /* the word being used up and the draws it still serves */
    uint32_t per_word;     /* draws a word serves (0 = RNG_DRAWS_PER_WORD) */
} rng;

enum { RNG_DRAWS_PER_WORD = 2, RNG_DRAWS_PER_WORD_KW = 3 };

static uint32_t draws_per_word_of(uint32_t game) { return game == ISMCTS_GAME_KW ? RNG_DRAWS_PER_WORD_KW : RNG_DRAWS_PER_WORD; }

/* The next whole word (EXP3 samples with all 32 bits); drops what is left of the current one. */
static uint32_t rng_word(rng* r)
{
    r->left = 0;
    if (r->cnt) r->cnt->draws++;
    return orc_draw_word(r->seed, r->pos, r->tree, r->iteration, r->d++);
}

/* Uniform integer in [0, n). */
static uint32_t rng_below(rng* r, uint32_t n)
{
    if (r->left == 0) { r->w = rng_word(r); r->left = r->per_word ? r->per_word : RNG_DRAWS_PER_WORD; }
    const uint64_t p = (uint64_t)r->w * n;
    r->w = (uint32_t)p;
    r->left--;
    return (uint32_t)(p >> 32);
}

static void rng_align(rng* r)
{
    r->left = 0;
    r->d = (r->d + 3u) & ~3u;
}

static void rng_restart(rng* r) { r->d = 0; r->left = 0; }

/* ========================================================== Knockout Whist */

typedef struct okw {
    int np, player, dealer, trump, tricks_left, request_trump;
    int alive[7], tricks[7];
    unsigned char hand[7][52];          /* the hands as sets ... */
    unsigned char hcard[7][7], hlive[7][7]; /* ... and as the seven places of a hand: the card put there, and whether it is still held */
    unsigned char unknown[52];
    int trick_n, trick_player[7], trick_card[7];
} okw;

/* The places of a hand.  A hand holds at most seven cards (the game never deals more); a card is
 * put on the first place after the last card still held, a played card leaves its place empty.
 * The default policy picks "the j-th legal card in place order", which costs the kernels a table
 * look-up instead of a search through a 64-bit set; any fixed order gives the same distribution
 * as the reference's RandomElement (utility.h:69-83). */
static void hand_clear(okw* g, int p)
{
    memset(g->hand[p], 0, sizeof g->hand[p]);
    memset(g->hlive[p], 0, sizeof g->hlive[p]);
}
static void hand_append(okw* g, int p, int card)
{
    int k = 0;
    for (int i = 0; i < 7; ++i) if (g->hlive[p][i]) k = i + 1;
    if (k >= 7) return;
    g->hcard[p][k] = (unsigned char)card; g->hlive[p][k] = 1; g->hand[p][card] = 1;
}
static void hand_remove(okw* g, int p, int card)
{
    for (int i = 0; i < 7; ++i)
        if (g->hlive[p][i] && g->hcard[p][i] == card) { g->hlive[p][i] = 0; break; }
    g->hand[p][card] = 0;
}

static void okw_from_pod(okw* g, const ismcts_kw_state* s)
{
    memset(g, 0, sizeof *g);
    g->np = s->num_players; g->player = s->player; g->dealer = s->dealer; g->trump = s->trump;
    g->tricks_left = s->tricks_left; g->request_trump = s->request_trump;
    for (int p = 0; p < 7; ++p) {
        g->alive[p] = (s->alive_mask >> p) & 1;
        g->tricks[p] = s->tricks_taken[p];
        for (int c = 0; c < 52; ++c) if ((s->hands[p] >> c) & 1) hand_append(g, p, c);   /* ascending */
    }
    for (int c = 0; c < 52; ++c) g->unknown[c] = (unsigned char)((s->unknown >> c) & 1);
    g->trick_n = s->trick_len;
    for (int i = 0; i < 7; ++i) { g->trick_player[i] = s->trick_player[i]; g->trick_card[i] = s->trick_card[i]; }
}

static void okw_to_pod(const okw* g, ismcts_kw_state* s)
{
    memset(s, 0, sizeof *s);
    s->num_players = (uint8_t)g->np; s->player = (uint8_t)g->player; s->dealer = (uint8_t)g->dealer;
    s->trump = (uint8_t)g->trump; s->tricks_left = (uint8_t)g->tricks_left;
    s->request_trump = (uint8_t)g->request_trump;
    for (int p = 0; p < 7; ++p) {
        if (g->alive[p]) s->alive_mask |= (uint8_t)(1u << p);
        s->tricks_taken[p] = (uint8_t)g->tricks[p];
        for (int c = 0; c < 52; ++c) if (g->hand[p][c]) s->hands[p] |= (uint64_t)1 << c;
    }
    for (int c = 0; c < 52; ++c) if (g->unknown[c]) s->unknown |= (uint64_t)1 << c;
    s->trick_len = (uint8_t)g->trick_n;
    for (int i = 0; i < g->trick_n; ++i) {
        s->trick_player[i] = (uint8_t)g->trick_player[i];
        s->trick_card[i] = (uint8_t)g->trick_card[i];
    }
}

static int set_size(const unsigned char* set, int n)
{
    int k = 0;
    for (int i = 0; i < n; ++i) k += set[i] != 0;
    return k;
}

/* Removes and returns a uniformly random member of `set` (ascending order index). */
static int set_draw(unsigned char* set, int n, rng* r)
{
    const int size = set_size(set, n);
    int k = (int)rng_below(r, (uint32_t)size);
    if (r->cnt) r->cnt->chance_picks++;
    for (int i = 0; i < n; ++i)
        if (set[i] && k-- == 0) { set[i] = 0; return i; }
    return -1;
}

/* nextPlayer, knockoutwhist.cpp:83-88: next id (cyclic) that is still in the game. */
static int okw_next(const okw* g, int p)
{
    do { p = (p + 1) % g->np; } while (!g->alive[p]);
    return p;
}

static int okw_alive_count(const okw* g)
{
    int k = 0;
    for (int p = 0; p < g->np; ++p) k += g->alive[p];
    return k;
}

/* A pile of cards to deal from: a uniformly random card leaves it by taking the place of...
 * rather, by the LAST card of the pile taking its place (one step of a Fisher-Yates shuffle
 * that only keeps the dealt prefix).  The pile starts in ascending card order. */
typedef struct okw_pile { unsigned char card[52]; int n; } okw_pile;

static void pile_from_set(okw_pile* d, const unsigned char* set)
{
    d->n = 0;
    for (int c = 0; c < 52; ++c) if (set[c]) d->card[d->n++] = (unsigned char)c;
}

static int pile_draw(okw_pile* d, rng* r)
{
    const int j = (int)rng_below(r, (uint32_t)d->n);
    if (r->cnt) r->cnt->chance_picks++;
    const int c = d->card[j];
    d->card[j] = d->card[--d->n];
    return c;
}

/* deal, knockoutwhist.cpp:138-152.  The reference shuffles all 52 cards and hands tricks_left
 * to each remaining player in order; drawing the same cards one by one without replacement
 * gives the same distribution. */
static void okw_deal(okw* g, rng* r)
{
    unsigned char all[52];
    memset(all, 1, sizeof all);
    okw_pile pile;
    pile_from_set(&pile, all);
    for (int p = 0; p < g->np; ++p) {
        if (!g->alive[p]) continue;
        for (int j = 0; j < g->tricks_left; ++j)
            hand_append(g, p, pile_draw(&pile, r));
    }
    memset(g->unknown, 0, sizeof g->unknown);
    for (int i = 0; i < pile.n; ++i) g->unknown[pile.card[i]] = 1;
}

/* KnockoutWhist::KnockoutWhist, knockoutwhist.cpp:36-54. */
static void okw_new(okw* g, int players, rng* r)
{
    memset(g, 0, sizeof *g);
    g->np = players < 2 ? 2 : players > 7 ? 7 : players;
    g->tricks_left = 7;
    g->dealer = g->np - 1;
    for (int p = 0; p < g->np; ++p) g->alive[p] = 1;
    okw_deal(g, r);
    /* one of the remaining cards is turned up for trumps and leaves the unknown set (:50-53) */
    rng_align(r);
    g->trump = set_draw(g->unknown, 52, r) / 13;
}

/* validMoves, knockoutwhist.cpp:122-136.  Returns the count, ascending ids. */
static int okw_legal(const okw* g, uint16_t* moves)
{
    int n = 0;
    if (g->request_trump) {
        for (int s = 0; s < 4; ++s) moves[n++] = (uint16_t)(13 * s); /* (Two, suit s), :18-24 */
        return n;
    }
    const unsigned char* hand = g->hand[g->player];
    const int hand_n = set_size(hand, 52);
    if (g->trick_n > 0 && hand_n >= 2) {
        const int lead = g->trick_card[0] / 13;
        for (int c = 13 * lead; c < 13 * lead + 13; ++c) if (hand[c]) moves[n++] = (uint16_t)c;
        if (n > 0) return n;
    }
    for (int c = 0; c < 52; ++c) if (hand[c]) moves[n++] = (uint16_t)c;
    return n;
}

/* The move of the default policy (simulate, solverbase.h:36-43): uniform over validMoves(), taken as
 * the j-th legal card in place order (the round winner's call of trumps: the j-th suit). */
static int okw_rollout_move(const okw* g, rng* r)
{
    if (g->request_trump) return 13 * (int)rng_below(r, 4);
    const int p = g->player;
    int place[7], n = 0;
    if (g->trick_n > 0) {
        const int lead = g->trick_card[0] / 13;
        for (int i = 0; i < 7; ++i) if (g->hlive[p][i] && g->hcard[p][i] / 13 == lead) place[n++] = i;
    }
    if (n == 0) for (int i = 0; i < 7; ++i) if (g->hlive[p][i]) place[n++] = i;
    return g->hcard[p][place[rng_below(r, (uint32_t)n)]];
}

/* trickWinner, knockoutwhist.cpp:182-196. */
static int okw_trick_winner(const okw* g)
{
    int w = 0;
    for (int i = 1; i < g->trick_n; ++i) {
        const int c = g->trick_card[i], wc = g->trick_card[w];
        if (c / 13 == wc / 13) { if (c % 13 > wc % 13) w = i; }
        else if (c / 13 == g->trump) w = i;
    }
    return g->trick_player[w];
}

/* finishRound, knockoutwhist.cpp:163-180 (+ roundWinner :198-210). */
static void okw_finish_round(okw* g, rng* r)
{
    for (int p = 0; p < g->np; ++p) if (g->alive[p] && g->tricks[p] == 0) g->alive[p] = 0;
    if (okw_alive_count(g) > 1) {
        g->tricks_left--;
        g->request_trump = 1;
        /* roundWinner: uniform among the players tied on most tricks */
        int best = 0;
        for (int p = 0; p < g->np; ++p) if (g->alive[p] && g->tricks[p] > best) best = g->tricks[p];
        unsigned char tied[7] = {0};
        for (int p = 0; p < g->np; ++p) tied[p] = (unsigned char)(g->alive[p] && g->tricks[p] == best);
        rng_align(r);                  /* the chance events of a round end start a Philox block ... */
        g->player = set_draw(tied, 7, r);
        g->dealer = okw_next(g, g->dealer);
        for (int p = 0; p < g->np; ++p) g->tricks[p] = 0;
        okw_deal(g, r);
        rng_align(r);                  /* ... and the next decision starts one */
    } else {
        g->tricks_left = 0; /* game over; the reference's final deal() hands out 0 cards */
    }
}

/* doMove, knockoutwhist.cpp:95-115.  Returns 0 or ISMCTS_ERR_ILLEGAL_MOVE. */
static int okw_apply(okw* g, int move, rng* r)
{
    if (move < 0 || move >= 52) return ISMCTS_ERR_ILLEGAL_MOVE;
    if (r->cnt) r->cnt->plies++;
    if (g->request_trump) {
        g->trump = move / 13;
        g->request_trump = 0;
        g->player = okw_next(g, g->dealer);
        return 0;
    }
    if (!g->hand[g->player][move]) return ISMCTS_ERR_ILLEGAL_MOVE; /* std::out_of_range, :108-109 */
    g->trick_player[g->trick_n] = g->player;
    g->trick_card[g->trick_n] = move;
    g->trick_n++;
    hand_remove(g, g->player, move);
    g->player = okw_next(g, g->player);
    if (g->trick_n == okw_alive_count(g)) { /* finishTrick, :155-161 */
        const int w = okw_trick_winner(g);
        g->tricks[w]++;
        g->trick_n = 0;
        g->player = w;
    }
    if (set_size(g->hand[g->player], 52) == 0) okw_finish_round(g, r);
    return 0;
}

/* getResult, knockoutwhist.cpp:117-120: 1 for the first remaining player. */
static int okw_result2(const okw* g, int player)
{
    for (int p = 0; p < g->np; ++p) if (g->alive[p]) return p == player ? 2 : 0;
    return 0;
}

/* cloneAndRandomise, knockoutwhist.cpp:56-76: everything the observer cannot
 * see (unknown cards and the other players' hands) is pooled and redealt to
 * the other players, keeping their hand sizes; m_unknownCards is left alone. */
static void okw_determinise(okw* g, int observer, rng* r)
{
    unsigned char pool[52];
    memcpy(pool, g->unknown, sizeof pool);
    int size[7] = {0};
    for (int p = 0; p < g->np; ++p) {
        if (!g->alive[p] || p == observer) continue;
        for (int c = 0; c < 52; ++c) if (g->hand[p][c]) { pool[c] = 1; size[p]++; }
        hand_clear(g, p);
    }
    okw_pile pile;
    pile_from_set(&pile, pool);
    rng_align(r);
    for (int p = 0; p < g->np; ++p) {
        if (!g->alive[p] || p == observer) continue;
        for (int j = 0; j < size[p]; ++j) hand_append(g, p, pile_draw(&pile, r));
    }
    rng_align(r);
}

/* ==================================================================== m,n,k */

typedef struct omnk {
    int m, n, k, player, result2;
    signed char board[256]; /* -1 empty, else player */
} omnk;

static void omnk_from_pod(omnk* g, const ismcts_mnk_state* s)
{
    g->m = s->m; g->n = s->n; g->k = s->k; g->player = s->player; g->result2 = s->result2;
    for (int c = 0; c < 256; ++c) {
        g->board[c] = -1;
        for (int p = 0; p < 2; ++p) if ((s->stones[p][c >> 6] >> (c & 63)) & 1) g->board[c] = (signed char)p;
    }
}

static void omnk_to_pod(const omnk* g, ismcts_mnk_state* s)
{
    memset(s, 0, sizeof *s);
    s->m = (uint8_t)g->m; s->n = (uint8_t)g->n; s->k = (uint8_t)g->k;
    s->player = (uint8_t)g->player; s->result2 = (int8_t)g->result2;
    for (int c = 0; c < g->m * g->n; ++c)
        if (g->board[c] >= 0) s->stones[(int)g->board[c]][c >> 6] |= (uint64_t)1 << (c & 63);
}

/* validMoves, mnkgame.cpp:61-64. */
static int omnk_legal(const omnk* g, uint16_t* moves)
{
    int n = 0;
    if (g->result2 != -1) return 0;
    for (int c = 0; c < g->m * g->n; ++c) if (g->board[c] < 0) moves[n++] = (uint16_t)c;
    return n;
}

/* hasWinningSequence / continueCounting, mnkgame.cpp:109-130: a run of EXACTLY k. */
static int omnk_run(const omnk* g, int move, int dr, int dc, int player)
{
    int count = 1;
    for (int sign = -1; sign <= 1; sign += 2) {
        int r = move / g->m + sign * dr, c = move % g->m + sign * dc;
        while (r >= 0 && c >= 0 && r < g->n && c < g->m && g->board[r * g->m + c] == player) {
            ++count; r += sign * dr; c += sign * dc;
        }
    }
    return count == g->k;
}

/* doMove, mnkgame.cpp:39-54. */
static int omnk_apply(omnk* g, int move, rng* r)
{
    if (move < 0 || move >= g->m * g->n || g->board[move] >= 0) return ISMCTS_ERR_ILLEGAL_MOVE;
    if (r && r->cnt) r->cnt->plies++;
    g->board[move] = (signed char)g->player;
    const int p = g->player;
    if (omnk_run(g, move, 0, 1, p) || omnk_run(g, move, 1, 0, p) || omnk_run(g, move, 1, 1, p) ||
        omnk_run(g, move, -1, 1, p)) {
        g->result2 = 2 * (1 - p);
    } else {
        int empty = 0;
        for (int c = 0; c < g->m * g->n; ++c) empty += g->board[c] < 0;
        if (empty) g->player = 1 - p; else g->result2 = 1;
    }
    return 0;
}

/* getResult, mnkgame.cpp:56-59. */
static int omnk_result2(const omnk* g, int player) { return player == 0 ? g->result2 : 2 - g->result2; }

/* ============================================================ phantom m,n,k */

typedef struct ophantom {
    omnk b;                    /* board, m, n, k, player, result2 */
    unsigned char avail[2][256];
} ophantom;

static void ophantom_from_pod(ophantom* g, const ismcts_phantom_state* s)
{
    ismcts_mnk_state base;
    memset(&base, 0, sizeof base);
    memcpy(base.stones, s->stones, sizeof base.stones);
    base.m = s->m; base.n = s->n; base.k = s->k; base.player = s->player; base.result2 = s->result2;
    omnk_from_pod(&g->b, &base);
    for (int p = 0; p < 2; ++p)
        for (int c = 0; c < 256; ++c) g->avail[p][c] = (unsigned char)((s->avail[p][c >> 6] >> (c & 63)) & 1);
}

static void ophantom_to_pod(const ophantom* g, ismcts_phantom_state* s)
{
    ismcts_mnk_state base;
    omnk_to_pod(&g->b, &base);
    memset(s, 0, sizeof *s);
    memcpy(s->stones, base.stones, sizeof base.stones);
    s->m = base.m; s->n = base.n; s->k = base.k; s->player = base.player; s->result2 = base.result2;
    for (int p = 0; p < 2; ++p)
        for (int c = 0; c < 256; ++c) if (g->avail[p][c]) s->avail[p][c >> 6] |= (uint64_t)1 << (c & 63);
}

/* validMoves, phantommnkgame.cpp:107-110. */
static int ophantom_legal(const ophantom* g, uint16_t* moves)
{
    int n = 0;
    for (int c = 0; c < g->b.m * g->b.n; ++c) if (g->avail[g->b.player][c]) moves[n++] = (uint16_t)c;
    return n;
}

static int ophantom_wins(const omnk* b, int move, int p)
{
    return omnk_run(b, move, 0, 1, p) || omnk_run(b, move, 1, 0, p) || omnk_run(b, move, 1, 1, p) || omnk_run(b, move, -1, 1, p);
}

/* doMove, phantommnkgame.cpp:83-105. */
static int ophantom_apply(ophantom* g, int move, rng* r)
{
    omnk* b = &g->b;
    const int p = b->player;
    if (move < 0 || move >= b->m * b->n || !g->avail[p][move]) return ISMCTS_ERR_ILLEGAL_MOVE;
    if (r && r->cnt) r->cnt->plies++;
    g->avail[p][move] = 0;
    if (b->board[move] >= 0) return 0;            /* occupied by the opponent: noted, the same player moves again */
    b->board[move] = (signed char)p;
    if (ophantom_wins(b, move, p)) {
        b->result2 = 2 * (1 - p);
        memset(g->avail, 0, sizeof g->avail);
    } else {
        int empty = 0;
        for (int c = 0; c < b->m * b->n; ++c) empty += b->board[c] < 0;
        if (empty) b->player = 1 - p; else b->result2 = 1;
    }
    return 0;
}

/* cloneAndRandomise, phantommnkgame.cpp:21-81.  The observer knows its own stones, the cells it
 * still believes free and how many stones the opponent has placed.  The opponent's stones it has not
 * run into are taken back and as many are put on random cells it cannot rule out, such that no stone
 * completes a winning line when it is placed (phantommnkgame.cpp:70-73 "ensure a non-winning state");
 * a sample that would is discarded and drawn again (the reference restarts recursively). */
static void ophantom_determinise(ophantom* g, int observer, rng* r)
{
    omnk* b = &g->b;
    const int opp = 1 - observer, cells = b->m * b->n;
    unsigned char cand[256] = {0};
    int num = 0;
    for (int c = 0; c < cells; ++c) {
        if (b->board[c] == opp && g->avail[observer][c]) { b->board[c] = -1; g->avail[opp][c] = 1; ++num; } /* undoMoves :37-55 */
        if (b->board[c] < 0) cand[c] = 1;
    }
    if (num == 0) return;
    for (;;) {
        unsigned char pool[256], placed[256] = {0};
        memcpy(pool, cand, sizeof pool);
        int ok = 1;
        for (int i = 0; i < num && ok; ++i) {
            const int c = set_draw(pool, 256, r);
            b->board[c] = (signed char)opp;
            placed[c] = 1;
            if (ophantom_wins(b, c, opp)) ok = 0;
        }
        if (ok) {
            for (int c = 0; c < cells; ++c) if (placed[c]) g->avail[opp][c] = 0;
            return;
        }
        for (int c = 0; c < cells; ++c) if (placed[c]) b->board[c] = -1;
    }
}

/* ================================================================ Goofspiel */

typedef struct ogoof {
    unsigned char hand[2][13];
    unsigned char prize_left[13];
    int order[13], n;          /* prizes not turned over, order[n-1] next (m_prizes.back()) */
    int known_order;           /* the observer is the environment: no reshuffle (goofspiel.cpp:69-71) */
    int next_prize;            /* the prize the environment turns over next, -1 = not determined yet */
    int current_prize, moves[2], scores[2], player, draw_prize;
    int hidden_bid_pending;    /* observer 1: player 0's bid is still to be randomised (goofspiel.cpp:76-78) */
} ogoof;

static int goof_value(int rank) { return rank == 12 ? 1 : rank + 2; } /* goofspiel.cpp:17-36 */

static void ogoof_from_pod(ogoof* g, const ismcts_goof_state* s)
{
    memset(g, 0, sizeof *g);
    for (int p = 0; p < 2; ++p)
        for (int r = 0; r < 13; ++r) g->hand[p][r] = (unsigned char)((s->hands[p] >> r) & 1);
    g->n = s->n_prizes;
    for (int i = 0; i < g->n; ++i) { g->order[i] = (int)((s->prize_order >> (4 * i)) & 15); g->prize_left[g->order[i]] = 1; }
    g->known_order = 1; /* a position as it is: the order is what it is until a determinisation hides it */
    g->next_prize = -1;
    g->current_prize = s->current_prize; g->moves[0] = s->moves[0]; g->moves[1] = s->moves[1];
    g->scores[0] = s->scores[0]; g->scores[1] = s->scores[1]; g->player = s->player; g->draw_prize = s->draw_prize;
}

static void ogoof_to_pod(const ogoof* g, ismcts_goof_state* s)
{
    memset(s, 0, sizeof *s);
    for (int p = 0; p < 2; ++p)
        for (int r = 0; r < 13; ++r) if (g->hand[p][r]) s->hands[p] |= (uint16_t)(1u << r);
    s->n_prizes = (uint8_t)g->n;
    for (int i = 0; i < g->n; ++i) s->prize_order |= (uint64_t)g->order[i] << (4 * i);
    s->current_prize = (uint8_t)g->current_prize; s->moves[0] = (uint8_t)g->moves[0]; s->moves[1] = (uint8_t)g->moves[1];
    s->scores[0] = (uint8_t)g->scores[0]; s->scores[1] = (uint8_t)g->scores[1];
    s->player = (uint8_t)g->player; s->draw_prize = (uint8_t)g->draw_prize;
}

/* The environment's next prize: the back of the known order, or — after a determinisation for a
 * player, who does not know the order (goofspiel.cpp:69-71) — a uniformly random remaining prize
 * (shuffling all remaining prizes at once and turning them over one by one is the same distribution). */
static void ogoof_draw_prize(ogoof* g, rng* r)
{
    int prize;
    if (g->known_order) {
        (void)rng_below(r, 1);
        if (r->cnt) r->cnt->chance_picks++;
        prize = g->order[g->n - 1];
        g->prize_left[prize] = 0;
    } else {
        prize = set_draw(g->prize_left, 13, r);
    }
    int k = 0;
    for (int i = 0; i < g->n; ++i) if (g->order[i] != prize) g->order[k++] = g->order[i];
    g->n = k;
    g->next_prize = prize;
}

/* validMoves, goofspiel.cpp:130-153. */
static int ogoof_legal(const ogoof* g, uint16_t* moves)
{
    int n = 0;
    if (g->player == 0) { for (int r = 0; r < 13; ++r) if (g->hand[0][r]) moves[n++] = (uint16_t)(39 + r); }
    else if (g->player == 1) { for (int r = 0; r < 13; ++r) if (g->hand[1][r]) moves[n++] = (uint16_t)r; }
    else if (!g->draw_prize) moves[n++] = 0;                       /* the dummy reveal move Card{} */
    else if (g->next_prize >= 0) moves[n++] = (uint16_t)(26 + g->next_prize);
    return n;
}

/* doMove + handleP2Turn, goofspiel.cpp:90-117. */
static int ogoof_apply(ogoof* g, int move, rng* r)
{
    if (r && r->cnt) r->cnt->plies++;
    if (g->player < 2) {
        const int rank = move - (g->player == 0 ? 39 : 0);
        if (rank < 0 || rank > 12 || !g->hand[g->player][rank]) return ISMCTS_ERR_ILLEGAL_MOVE;
        g->moves[g->player] = rank;
        g->player++;
        return 0;
    }
    if (g->draw_prize) {
        if (g->next_prize < 0 || move != 26 + g->next_prize) return ISMCTS_ERR_ILLEGAL_MOVE;
        g->current_prize = g->next_prize;
        g->next_prize = -1;
        g->player = 0;
        g->draw_prize = 0;
        return 0;
    }
    if (g->moves[0] != g->moves[1]) {
        const int w = goof_value(g->moves[0]) > goof_value(g->moves[1]) ? 0 : 1;
        g->scores[w] += goof_value(g->current_prize);
    }
    g->hand[0][g->moves[0]] = 0;
    g->hand[1][g->moves[1]] = 0;
    g->draw_prize = 1;
    if (g->n > 0 && r) ogoof_draw_prize(g, r);
    return 0;
}

/* getResult, goofspiel.cpp:119-128. */
static int ogoof_result2(const ogoof* g, int player)
{
    if (player == 2) return 0;
    if (g->scores[0] == g->scores[1]) return 1;
    return g->scores[player] > g->scores[1 - player] ? 2 : 0;
}

/* cloneAndRandomise, goofspiel.cpp:65-81. */
static void ogoof_determinise(ogoof* g, int observer, rng* r)
{
    g->known_order = observer == 2;
    if (observer == 1) {
        unsigned char pick[13];
        memcpy(pick, g->hand[0], sizeof pick);
        g->moves[0] = set_draw(pick, 13, r);
    }
    if (g->player == 2 && g->draw_prize && g->next_prize < 0 && g->n > 0) ogoof_draw_prize(g, r);
}

/* ============================================================ generic game */

typedef struct ogame {
    uint32_t kind;
    union { okw kw; omnk mnk; ogoof goof; ophantom ph; } u;
} ogame;

static int og_load(ogame* g, uint32_t kind, const void* root)
{
    g->kind = kind;
    if (kind == ISMCTS_GAME_KW) okw_from_pod(&g->u.kw, (const ismcts_kw_state*)root);
    else if (kind == ISMCTS_GAME_MNK) omnk_from_pod(&g->u.mnk, (const ismcts_mnk_state*)root);
    else if (kind == ISMCTS_GAME_GOOF) ogoof_from_pod(&g->u.goof, (const ismcts_goof_state*)root);
    else if (kind == ISMCTS_GAME_PHANTOM) ophantom_from_pod(&g->u.ph, (const ismcts_phantom_state*)root);
    else return ISMCTS_ERR_INVALID;
    return 0;
}
static int og_player(const ogame* g)
{
    return g->kind == ISMCTS_GAME_KW ? g->u.kw.player : g->kind == ISMCTS_GAME_MNK ? g->u.mnk.player
         : g->kind == ISMCTS_GAME_GOOF ? g->u.goof.player : g->u.ph.b.player;
}
static int og_legal(const ogame* g, uint16_t* mv)
{
    if (g->kind == ISMCTS_GAME_KW) return okw_legal(&g->u.kw, mv);
    if (g->kind == ISMCTS_GAME_MNK) return omnk_legal(&g->u.mnk, mv);
    if (g->kind == ISMCTS_GAME_PHANTOM) return ophantom_legal(&g->u.ph, mv);
    return ogoof_legal(&g->u.goof, mv);
}
static void og_apply(ogame* g, int mv, rng* r)
{
    if (g->kind == ISMCTS_GAME_KW) okw_apply(&g->u.kw, mv, r);
    else if (g->kind == ISMCTS_GAME_MNK) omnk_apply(&g->u.mnk, mv, r);
    else if (g->kind == ISMCTS_GAME_PHANTOM) ophantom_apply(&g->u.ph, mv, r);
    else ogoof_apply(&g->u.goof, mv, r);
}
static int og_result2(const ogame* g, int p)
{
    if (g->kind == ISMCTS_GAME_KW) return okw_result2(&g->u.kw, p);
    if (g->kind == ISMCTS_GAME_MNK) return omnk_result2(&g->u.mnk, p);
    if (g->kind == ISMCTS_GAME_PHANTOM) return omnk_result2(&g->u.ph.b, p);
    return ogoof_result2(&g->u.goof, p);
}
static void og_determinise(ogame* g, int obs, rng* r)
{
    if (g->kind == ISMCTS_GAME_KW) okw_determinise(&g->u.kw, obs, r);
    else if (g->kind == ISMCTS_GAME_GOOF) ogoof_determinise(&g->u.goof, obs, r);
    else if (g->kind == ISMCTS_GAME_PHANTOM) ophantom_determinise(&g->u.ph, obs, r);
    /* m,n,k: plain copy, mnkgame.cpp:34-37 */
}
/* players(), knockoutwhist.cpp:90-93 / mnkgame.cpp:76-79 */
static int og_players(const ogame* g, int* out)
{
    int n = 0;
    if (g->kind == ISMCTS_GAME_KW) { for (int p = 0; p < g->u.kw.np; ++p) if (g->u.kw.alive[p]) out[n++] = p; }
    else if (g->kind == ISMCTS_GAME_GOOF) { out[0] = 0; out[1] = 1; out[2] = 2; n = 3; } /* goofspiel.cpp:124-128 */
    else { out[0] = 0; out[1] = 1; n = 2; }
    return n;
}
static uint32_t og_stride(uint32_t kind) { return kind == ISMCTS_GAME_MNK || kind == ISMCTS_GAME_PHANTOM ? 256u : 64u; }

/* ======================================================== exported game API */

static rng generator_rng(uint64_t seed, uint32_t stream, uint32_t tree, uint32_t iteration, uint32_t d, uint32_t game)
{
    rng r = { seed, stream, tree, iteration, d, NULL, 0, 0, draws_per_word_of(game) };
    return r;
}

int orc_kw_deal(uint64_t seed, uint32_t stream, uint32_t num_players, ismcts_kw_state* out)
{
    okw g;
    rng r = generator_rng(seed, stream, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, ISMCTS_GAME_KW);
    okw_new(&g, (int)num_players, &r);
    okw_to_pod(&g, out);
    return 0;
}

uint64_t orc_kw_legal(const ismcts_kw_state* s)
{
    okw g; uint16_t mv[64]; uint64_t mask = 0;
    okw_from_pod(&g, s);
    const int n = okw_legal(&g, mv);
    for (int i = 0; i < n; ++i) mask |= (uint64_t)1 << mv[i];
    return mask;
}

int orc_kw_apply(ismcts_kw_state* s, uint32_t move, uint64_t seed, uint32_t stream, uint32_t* draw_counter)
{
    okw g;
    okw_from_pod(&g, s);
    rng r = generator_rng(seed, stream, 0xFFFFFFFEu, 0, draw_counter ? *draw_counter : 0, ISMCTS_GAME_KW);
    const int rc = okw_apply(&g, (int)move, &r);
    if (rc) return rc;
    if (draw_counter) *draw_counter = r.d;
    okw_to_pod(&g, s);
    return 0;
}

int orc_kw_result2(const ismcts_kw_state* s, uint32_t player)
{
    okw g; okw_from_pod(&g, s);
    return okw_result2(&g, (int)player);
}

int orc_kw_determinise(ismcts_kw_state* s, uint32_t observer, uint64_t seed, uint32_t pos, uint32_t tree,
                       uint32_t iteration, uint32_t* draw_counter)
{
    okw g; okw_from_pod(&g, s);
    rng r = generator_rng(seed, pos, tree, iteration, draw_counter ? *draw_counter : 0, ISMCTS_GAME_KW);
    okw_determinise(&g, (int)observer, &r);
    if (draw_counter) *draw_counter = r.d;
    okw_to_pod(&g, s);
    return 0;
}

int orc_mnk_init(uint32_t m, uint32_t n, uint32_t k, ismcts_mnk_state* out)
{
    if (m == 0 || n == 0 || m * n > 256) return ISMCTS_ERR_INVALID;
    memset(out, 0, sizeof *out);
    out->m = (uint8_t)m; out->n = (uint8_t)n; out->k = (uint8_t)k; out->player = 0; out->result2 = -1;
    return 0;
}

void orc_mnk_legal(const ismcts_mnk_state* s, uint64_t legal[4])
{
    omnk g; uint16_t mv[256];
    omnk_from_pod(&g, s);
    const int n = omnk_legal(&g, mv);
    legal[0] = legal[1] = legal[2] = legal[3] = 0;
    for (int i = 0; i < n; ++i) legal[mv[i] >> 6] |= (uint64_t)1 << (mv[i] & 63);
}

int orc_mnk_apply(ismcts_mnk_state* s, uint32_t move)
{
    omnk g; omnk_from_pod(&g, s);
    /* the reference looks the move up in m_moves (mnkgame.cpp:41-43); a finished game keeps its
       remaining cells there, so doMove itself does not reject moves after the end */
    const int rc = omnk_apply(&g, (int)move, NULL);
    if (rc) return rc;
    omnk_to_pod(&g, s);
    return 0;
}

int orc_mnk_result2(const ismcts_mnk_state* s, uint32_t player)
{
    omnk g; omnk_from_pod(&g, s);
    return omnk_result2(&g, (int)player);
}

int orc_phantom_init(uint32_t m, uint32_t n, uint32_t k, ismcts_phantom_state* out)
{
    if (m == 0 || n == 0 || m * n > 256) return ISMCTS_ERR_INVALID;
    memset(out, 0, sizeof *out);
    out->m = (uint8_t)m; out->n = (uint8_t)n; out->k = (uint8_t)k; out->result2 = -1;
    for (uint32_t c = 0; c < m * n; ++c) for (int p = 0; p < 2; ++p) out->avail[p][c >> 6] |= (uint64_t)1 << (c & 63);
    return 0;
}

void orc_phantom_legal(const ismcts_phantom_state* s, uint64_t legal[4])
{
    for (int i = 0; i < 4; ++i) legal[i] = s->avail[s->player & 1][i];
}

int orc_phantom_apply(ismcts_phantom_state* s, uint32_t move)
{
    ophantom g; ophantom_from_pod(&g, s);
    const int rc = ophantom_apply(&g, (int)move, NULL);
    if (rc) return rc;
    ophantom_to_pod(&g, s);
    return 0;
}

int orc_phantom_result2(const ismcts_phantom_state* s, uint32_t player)
{
    return player == 0 ? s->result2 : 2 - s->result2;
}

int orc_goof_init(uint64_t seed, uint32_t stream, ismcts_goof_state* out)
{
    /* Goofspiel::Goofspiel, goofspiel.cpp:52-63: the prizes are shuffled; entry i of the order is a
       uniformly random prize not placed yet */
    ogoof g;
    memset(&g, 0, sizeof g);
    rng r = generator_rng(seed, stream, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, ISMCTS_GAME_GOOF);
    unsigned char left[13];
    memset(left, 1, sizeof left);
    for (int p = 0; p < 2; ++p) for (int k = 0; k < 13; ++k) g.hand[p][k] = 1;
    for (int i = 0; i < 13; ++i) { g.order[i] = set_draw(left, 13, &r); g.prize_left[g.order[i]] = 1; }
    g.n = 13; g.player = 2; g.draw_prize = 1; g.next_prize = -1;
    ogoof_to_pod(&g, out);
    return 0;
}

/* A position as its own player to move sees it: the environment knows which prize comes next. */
static void ogoof_face_up(ogoof* g)
{
    if (g->player == 2 && g->draw_prize && g->n > 0) {
        g->next_prize = g->order[g->n - 1];
    }
}

uint64_t orc_goof_legal(const ismcts_goof_state* s)
{
    ogoof g; uint16_t mv[16]; uint64_t mask = 0;
    ogoof_from_pod(&g, s);
    ogoof_face_up(&g);
    const int n = ogoof_legal(&g, mv);
    for (int i = 0; i < n; ++i) mask |= (uint64_t)1 << mv[i];
    return mask;
}

int orc_goof_apply(ismcts_goof_state* s, uint32_t move)
{
    ogoof g;
    ogoof_from_pod(&g, s);
    ogoof_face_up(&g);
    if (g.player == 2 && g.draw_prize && g.next_prize >= 0) {
        /* turning the prize over removes it from the order */
        g.prize_left[g.next_prize] = 0;
        g.n--;
    }
    const int rc = ogoof_apply(&g, (int)move, NULL);
    if (rc) return rc;
    ogoof_to_pod(&g, s);
    return 0;
}

int orc_goof_result2(const ismcts_goof_state* s, uint32_t player)
{
    ogoof g; ogoof_from_pod(&g, s);
    return ogoof_result2(&g, (int)player);
}

/* ================================================================= playout */

/* simulate, solverbase.h:36-43 with RandomElement (utility.h:69-83): uniform
 * random legal moves until there are none.  Returns plies played. */
static uint32_t og_rollout(ogame* g, rng* r)
{
    uint16_t mv[256];
    uint32_t plies = 0;
    int n;
    if (g->kind == ISMCTS_GAME_KW) rng_align(r);   /* Knockout Whist: the rollout starts a Philox block */
    while ((n = og_legal(g, mv)) > 0) {
        og_apply(g, g->kind == ISMCTS_GAME_KW ? okw_rollout_move(&g->u.kw, r) : mv[rng_below(r, (uint32_t)n)], r);
        ++plies;
    }
    return plies;
}

uint32_t orc_playout(uint32_t game, const void* root, uint64_t seed, uint32_t pos, uint32_t tree,
                     uint32_t iteration)
{
    ogame g;
    if (og_load(&g, game, root)) return 0xFFFFFFFFu;
    rng r = { seed, pos, tree, iteration, 0, NULL, 0, 0, draws_per_word_of(game) };
    og_determinise(&g, og_player(&g), &r);
    const uint32_t plies = og_rollout(&g, &r);
    uint32_t winner;
    if (game == ISMCTS_GAME_KW) {
        winner = 0xFF;
        for (int p = 0; p < g.u.kw.np; ++p) if (g.u.kw.alive[p]) { winner = (uint32_t)p; break; }
    } else if (game == ISMCTS_GAME_GOOF) {
        winner = (uint32_t)ogoof_result2(&g.u.goof, 0);
    } else if (game == ISMCTS_GAME_PHANTOM) {
        winner = (uint32_t)(g.u.ph.b.result2 & 0xFF);
    } else {
        winner = (uint32_t)(g.u.mnk.result2 & 0xFF);
    }
    return (winner & 0xFF) | ((plies & 0xFFF) << 8) | ((r.d & 0xFFF) << 20);
}

/* ==================================================================== tree */

typedef struct onode {
    uint32_t parent, move, player, visits, score2, avail;
    double fscore, prob;
    uint32_t* child;
    uint32_t nchild, cap;
} onode;

typedef struct otree { onode* n; uint32_t count, cap; } otree;

static void otree_init(otree* t)
{
    t->cap = 1024; t->count = 1;
    t->n = (onode*)calloc(t->cap, sizeof(onode));
    t->n[0].avail = 1; t->n[0].prob = 1.0; /* root: default-constructed node, node.h:30-33 */
}

static void otree_free(otree* t)
{
    for (uint32_t i = 0; i < t->count; ++i) free(t->n[i].child);
    free(t->n);
}

/* addChild, node.h:44-48 / newChild, solverbase.h:74-80. */
static uint32_t otree_add(otree* t, uint32_t parent, uint32_t move, uint32_t player)
{
    if (t->count == t->cap) {
        t->cap *= 2;
        t->n = (onode*)realloc(t->n, t->cap * sizeof(onode));
    }
    const uint32_t id = t->count++;
    onode* c = &t->n[id];
    memset(c, 0, sizeof *c);
    c->parent = parent; c->move = move; c->player = player;
    c->avail = 1;  /* ucb1.h:51 */
    c->prob = 1.0; /* exp3.h:42 */
    onode* p = &t->n[parent];
    if (p->nchild == p->cap) {
        p->cap = p->cap ? 2 * p->cap : 4;
        p->child = (uint32_t*)realloc(p->child, p->cap * sizeof(uint32_t));
    }
    p->child[p->nchild++] = id;
    return id;
}

static int otree_find(const otree* t, uint32_t parent, uint32_t move)
{
    const onode* p = &t->n[parent];
    for (uint32_t i = 0; i < p->nchild; ++i) if (t->n[p->child[i]].move == move) return (int)p->child[i];
    return -1;
}

/* untriedMoves, node.h:82-91: legal moves without a child, in legal-move order. */
static int otree_untried(const otree* t, uint32_t node, const uint16_t* legal, int n, uint16_t* out)
{
    int k = 0;
    for (int i = 0; i < n; ++i) if (otree_find(t, node, legal[i]) < 0) out[k++] = legal[i];
    return k;
}

double orc_ucb(uint32_t score2, uint32_t visits, uint32_t avail, double exploration)
{
    return ismcts_ucb_value(score2, ismcts_tab_inv_visits(visits), ismcts_tab_inv_sqrt_visits(visits),
                            ismcts_tab_sqrt_ln(avail), exploration);
}

void orc_exp3_probabilities(const double* scores, const uint32_t* visits, uint32_t K, double* p_out)
{
    /* exp3.h:73-95.  t-1 wraps as size_t when t == 0 (never in a sequential search). */
    uint64_t t = 0;
    for (uint32_t i = 0; i < K; ++i) t += visits[i];
    const double lnK = log((double)K);
    const double e_t = ismcts_exp3_epsilon(K, lnK, (double)t);
    const double e_tm1 = ismcts_exp3_epsilon(K, lnK, (double)(uint64_t)(t - 1));
    double smax = scores[0];
    for (uint32_t i = 1; i < K; ++i) if (scores[i] > smax) smax = scores[i];
    double sum = 0.0;
    for (uint32_t i = 0; i < K; ++i) {
        p_out[i] = ismcts_det_exp(ISMCTS_DMUL(e_tm1, ISMCTS_DSUB(scores[i], smax)));
        sum = ISMCTS_DADD(sum, p_out[i]);
    }
    const double rest = ISMCTS_DSUB(1.0, ISMCTS_DMUL((double)K, e_t));
    for (uint32_t i = 0; i < K; ++i)
        p_out[i] = ISMCTS_DADD(e_t, ISMCTS_DDIV(ISMCTS_DMUL(rest, p_out[i]), sum));
}

/* Node::selectChild + policy, node.h:58-72.  Children whose move is legal, in
 * creation order; UCB1 (ucb1.h:69-76) marks each available and takes the first
 * maximum; EXP3 (exp3.h:57-62) stores the probabilities and samples. */
static uint32_t otree_select(otree* t, uint32_t node, const uint16_t* legal, int n, uint32_t policy,
                             double exploration, rng* r)
{
    uint32_t cand[256];
    uint32_t K = 0;
    const onode* p = &t->n[node];
    for (uint32_t i = 0; i < p->nchild; ++i) {
        const uint32_t c = p->child[i];
        for (int j = 0; j < n; ++j) if (legal[j] == t->n[c].move) { cand[K++] = c; break; }
    }
    if (r->cnt) { r->cnt->select_steps++; r->cnt->child_evals += K; }
    if (policy == ISMCTS_POLICY_UCB1) {
        for (uint32_t i = 0; i < K; ++i) t->n[cand[i]].avail++;
        uint32_t best = cand[0];
        double best_u = orc_ucb(t->n[best].score2, t->n[best].visits, t->n[best].avail, exploration);
        for (uint32_t i = 1; i < K; ++i) {
            const onode* c = &t->n[cand[i]];
            const double u = orc_ucb(c->score2, c->visits, c->avail, exploration);
            if (best_u < u) { best_u = u; best = cand[i]; }
        }
        return best;
    }
    /* EXP3: the reference samples a std::discrete_distribution over the legal children in
     * child-vector order (exp3.h:57-62); any fixed order gives the same distribution.  The oracle
     * and the kernels take them in ascending move order (legal[] is ascending). */
    K = 0;
    for (int j = 0; j < n; ++j) {
        const int c = otree_find(t, node, legal[j]);
        if (c >= 0) cand[K++] = (uint32_t)c;
    }
    double scores[256], probs[256];
    uint32_t visits[256];
    for (uint32_t i = 0; i < K; ++i) { scores[i] = t->n[cand[i]].fscore; visits[i] = t->n[cand[i]].visits; }
    orc_exp3_probabilities(scores, visits, K, probs);
    for (uint32_t i = 0; i < K; ++i) t->n[cand[i]].prob = probs[i];
    /* sample: x uniform in [0,1) from one word; first i with x < cumulative p */
    const uint32_t w = rng_word(r);
    const double x = ISMCTS_DMUL((double)w, 1.0 / 4294967296.0);
    double cum = 0.0;
    for (uint32_t i = 0; i < K; ++i) {
        cum = ISMCTS_DADD(cum, probs[i]);
        if (x < cum) return cand[i];
    }
    return cand[K - 1];
}

/* Node::update + updateData, node.h:74-80, ucb1.h:53-56, exp3.h:45-48; walk of
 * backPropagate, solverbase.h:45-51 (the root is never updated). */
static void otree_backprop(otree* t, uint32_t node, const ogame* g, uint32_t policy, orc_counters* cnt)
{
    while (node != 0) {
        onode* n = &t->n[node];
        const int r2 = og_result2(g, (int)n->player);
        n->visits++;
        if (policy == ISMCTS_POLICY_UCB1) {
            n->score2 += (uint32_t)r2;
            n->fscore = ISMCTS_DMUL((double)n->score2, 0.5);
        } else {
            n->fscore = ISMCTS_DADD(n->fscore, ISMCTS_DDIV(ISMCTS_DMUL((double)r2, 0.5), n->prob));
        }
        if (cnt) cnt->backprop_updates++;
        node = n->parent;
    }
}

static void otree_export(const otree* t, ismcts_node_rec* out)
{
    for (uint32_t i = 0; i < t->count; ++i) {
        const onode* n = &t->n[i];
        out[i].parent = n->parent; out[i].move = n->move; out[i].player = n->player;
        out[i].visits = n->visits; out[i].score2 = n->score2; out[i].avail = n->avail;
        out[i].score = n->fscore; out[i].prob = n->prob;
    }
}

/* ============================================================== SO-ISMCTS */

/* SOSolver::search, sosolver.h:44-71. */
static void so_iteration(otree* t, const ogame* root, uint32_t policy, double exploration, rng* r)
{
    ogame g = *root;
    uint16_t legal[256], untried[256];
    rng_restart(r);
    og_determinise(&g, og_player(root), r);
    uint32_t node = 0;
    for (;;) {
        const int n = og_legal(&g, legal);
        if (n == 0) break;                                   /* selectNode: terminal, solverbase.h:53-56 */
        const int u = otree_untried(t, node, legal, n, untried);
        if (u > 0) {                                         /* expand, sosolver.h:63-71 */
            const uint32_t mv = untried[rng_below(r, (uint32_t)u)];
            node = otree_add(t, node, mv, (uint32_t)og_player(&g));
            if (r->cnt) r->cnt->nodes_created++;
            og_apply(&g, (int)mv, r);
            break;
        }
        node = otree_select(t, node, legal, n, policy, exploration, r); /* select, sosolver.h:53-61 */
        og_apply(&g, (int)t->n[node].move, r);
    }
    og_rollout(&g, r);
    otree_backprop(t, node, &g, policy, r->cnt);
    if (r->cnt) r->cnt->iterations++;
}

int orc_so_search(uint32_t game, const void* root, uint32_t policy, double exploration, uint64_t seed,
                  uint32_t pos, uint32_t tree, uint64_t iterations, ismcts_node_rec* nodes,
                  uint32_t max_nodes, uint32_t* n_nodes, orc_counters* counters)
{
    ogame g;
    if (og_load(&g, game, root)) return ISMCTS_ERR_INVALID;
    otree t;
    otree_init(&t);
    rng r = { seed, pos, tree, 0, 0, counters, 0, 0, draws_per_word_of(game) };
    for (uint64_t it = 0; it < iterations; ++it) {
        r.iteration = (uint32_t)it;
        so_iteration(&t, &g, policy, exploration, &r);
    }
    int rc = 0;
    if (n_nodes) *n_nodes = t.count;
    if (nodes) {
        if (t.count > max_nodes) rc = ISMCTS_ERR_CAPACITY; else otree_export(&t, nodes);
    }
    otree_free(&t);
    return rc;
}

/* ============================================================== MO-ISMCTS */

typedef struct motrees { otree t[8]; int players[8]; int np; } motrees;

static int mo_index(const motrees* m, int player)
{
    for (int i = 0; i < m->np; ++i) if (m->players[i] == player) return i;
    return -1;
}

/* findOrAddChild, node.h:50-56. */
static uint32_t otree_find_or_add(otree* t, uint32_t node, uint32_t move, uint32_t player, orc_counters* cnt)
{
    const int c = otree_find(t, node, move);
    if (c >= 0) return (uint32_t)c;
    if (cnt) cnt->nodes_created++;
    return otree_add(t, node, move, player);
}

/* MOSolver::search, mosolver.h:57-101. */
static void mo_iteration(motrees* m, const ogame* root, uint32_t policy, double exploration, rng* r)
{
    ogame g = *root;
    uint16_t legal[256], untried[256];
    uint32_t cur[8] = {0};
    rng_restart(r);
    og_determinise(&g, og_player(root), r);
    for (;;) {
        const int n = og_legal(&g, legal);
        if (n == 0) break;
        const int player = og_player(&g);
        const int ti = mo_index(m, player);
        const int u = otree_untried(&m->t[ti], cur[ti], legal, n, untried);
        uint32_t mv;
        if (u > 0) {                                         /* expand, mosolver.h:84-95 */
            mv = untried[rng_below(r, (uint32_t)u)];
        } else {                                             /* select, mosolver.h:69-82 */
            const uint32_t c = otree_select(&m->t[ti], cur[ti], legal, n, policy, exploration, r);
            mv = m->t[ti].n[c].move;
        }
        for (int i = 0; i < m->np; ++i)
            cur[i] = otree_find_or_add(&m->t[i], cur[i], mv, (uint32_t)player, r->cnt);
        og_apply(&g, (int)mv, r);
        if (u > 0) break;
    }
    og_rollout(&g, r);
    for (int i = 0; i < m->np; ++i) otree_backprop(&m->t[i], cur[i], &g, policy, r->cnt); /* :97-101 */
    if (r->cnt) r->cnt->iterations++;
}

static void mo_init(motrees* m, const ogame* g)
{
    m->np = og_players(g, m->players); /* newTree, mosolver.h:106-112 */
    for (int i = 0; i < m->np; ++i) otree_init(&m->t[i]);
}

int orc_mo_search(uint32_t game, const void* root, uint32_t policy, double exploration, uint64_t seed,
                  uint32_t pos, uint32_t tree, uint64_t iterations, ismcts_node_rec* nodes,
                  uint32_t max_nodes, uint32_t tree_offsets[8], uint32_t tree_counts[8],
                  orc_counters* counters)
{
    ogame g;
    if (og_load(&g, game, root)) return ISMCTS_ERR_INVALID;
    motrees m;
    mo_init(&m, &g);
    rng r = { seed, pos, tree, 0, 0, counters, 0, 0, draws_per_word_of(game) };
    for (uint64_t it = 0; it < iterations; ++it) {
        r.iteration = (uint32_t)it;
        mo_iteration(&m, &g, policy, exploration, &r);
    }
    int rc = 0;
    uint32_t off = 0;
    for (int i = 0; i < 8; ++i) { tree_offsets[i] = 0; tree_counts[i] = 0; }
    for (int i = 0; i < m.np; ++i) {
        const int p = m.players[i];
        tree_offsets[p] = off; tree_counts[p] = m.t[i].count;
        if (nodes) {
            if (off + m.t[i].count > max_nodes) rc = ISMCTS_ERR_CAPACITY;
            else otree_export(&m.t[i], nodes + off);
        }
        off += m.t[i].count;
    }
    for (int i = 0; i < m.np; ++i) otree_free(&m.t[i]);
    return rc;
}

/* =========================================================== root parallel */

uint64_t orc_tree_iterations(uint64_t iterations, uint32_t trees_total, uint32_t tree_global)
{
    return iterations / trees_total + (tree_global < iterations % trees_total ? 1 : 0);
}

/* compileVisitCounts, execution.h:219-231, plus reward/availability sums. */
static void add_root_children(const otree* t, uint64_t* visits, uint64_t* score2, uint64_t* avail)
{
    const onode* root = &t->n[0];
    for (uint32_t i = 0; i < root->nchild; ++i) {
        const onode* c = &t->n[root->child[i]];
        visits[c->move] += c->visits;
        score2[c->move] += c->score2;
        avail[c->move] += c->avail;
    }
}

int orc_root_parallel(uint32_t game, uint32_t solver, const void* root, uint32_t policy, double exploration,
                      uint64_t seed, uint32_t pos, uint32_t tree_base, uint32_t trees, uint32_t trees_total,
                      uint64_t iterations, uint64_t* visits, uint64_t* score2, uint64_t* avail,
                      uint32_t* best_move, orc_counters* counters)
{
    ogame g;
    if (og_load(&g, game, root)) return ISMCTS_ERR_INVALID;
    const uint32_t stride = og_stride(game);
    memset(visits, 0, stride * sizeof(uint64_t));
    memset(score2, 0, stride * sizeof(uint64_t));
    memset(avail, 0, stride * sizeof(uint64_t));
    for (uint32_t j = 0; j < trees; ++j) {
        const uint32_t tg = tree_base + j;
        const uint64_t n_it = orc_tree_iterations(iterations, trees_total, tg);
        rng r = { seed, pos, tg, 0, 0, counters, 0, 0, draws_per_word_of(game) };
        if (solver == ISMCTS_SOLVER_SO) {
            otree t;
            otree_init(&t);
            for (uint64_t it = 0; it < n_it; ++it) { r.iteration = (uint32_t)it; so_iteration(&t, &g, policy, exploration, &r); }
            add_root_children(&t, visits, score2, avail);
            otree_free(&t);
        } else {
            motrees m;
            mo_init(&m, &g);
            for (uint64_t it = 0; it < n_it; ++it) { r.iteration = (uint32_t)it; mo_iteration(&m, &g, policy, exploration, &r); }
            /* mosolver.h:42-46: the tree of the root's current player decides */
            add_root_children(&m.t[mo_index(&m, og_player(&g))], visits, score2, avail);
            for (int i = 0; i < m.np; ++i) otree_free(&m.t[i]);
        }
    }
    /* RootParallel::bestMove, execution.h:209-216: first maximum in ascending move order */
    uint32_t best = 0xFFFFFFFFu;
    uint64_t best_v = 0;
    for (uint32_t mv = 0; mv < stride; ++mv)
        if (visits[mv] > best_v) { best_v = visits[mv]; best = mv; }
    if (best_move) *best_move = best;
    return 0;
}
