// phantom_game.cuh — the phantom m,n,k-game as a device-side POD state machine.
//
// Device mirror of the reference's PhantomMnkGame (test/common/phantommnkgame.{h,cpp}): the m,n,k-game
// (mnk_game.cuh) in which the players do not see each other's stones.  A player moves on a cell it
// still believes free; if the opponent is there the cell is only struck off its list and it moves again
// (doMove, phantommnkgame.cpp:83-105).  Per thread, in shared memory: the two stone sets, the two
// "believed free" sets and the replay pool (five 256-bit masks).
//
// cloneAndRandomise (phantommnkgame.cpp:21-81) as chance micro-steps: the opponent's stones the observer
// has not run into are taken back (done once per position, in prepare()) and as many are put, one per
// micro-step, on uniformly random cells the observer cannot rule out; a stone that would complete a
// winning line voids the sample, which is drawn again from the start ("ensure a non-winning state",
// :70-73; the reference restarts recursively).
#pragma once

#include "mnk_game.cuh"

namespace ismcts {

// What an iteration starts from: the observer's view, hidden stones removed.
struct alignas(16) PhantomPrepared {
    uint64_t stones[2][4];   // the observer's stones and the opponent's stones it knows
    uint64_t avail[2][4];    // believed-free sets, the opponent's with the removed stones added back
    uint64_t cand[4];        // cells the hidden stones may be on
    uint32_t hidden;         // how many stones are hidden
    uint32_t m, n, k, player;
    int32_t result2;
    uint32_t pad[2];
};

struct PhantomGame : MnkGame {
    static constexpr uint32_t kGameId = ISMCTS_GAME_PHANTOM;
    static constexpr uint32_t kHandWords = 20;  // stones[2][4], avail[2][4], pool[4]
    static constexpr uint32_t kMaxDepth = 512;  // every player may try every cell
    using Root = ismcts_phantom_state;
    using Prepared = PhantomPrepared;

    const Prepared* prep;    // to restart a voided sample from
    uint32_t replay_left;    // hidden stones still to place (chance pending while > 0)
    uint32_t over;           // validMoves() is empty

    // believed-free set of player p / the replay pool: words 8.. of the scratch column
    ISMCTS_DEV uint64_t& avail_word(uint32_t p, uint32_t i) const { return stones[(8 + p * 4 + i) * stride]; }
    ISMCTS_DEV uint64_t& pool_word(uint32_t i) const { return stones[(16 + i) * stride]; }

    ISMCTS_DEV void bind(uint64_t* scratch, uint32_t scratch_stride, uint32_t column) { MnkGame::bind(scratch, scratch_stride, column); replay_left = 0; over = 1; }

    ISMCTS_DEV Mask avail_of(uint32_t p) const { return Mask{{avail_word(p, 0), avail_word(p, 1), avail_word(p, 2), avail_word(p, 3)}}; }

    ISMCTS_DEV void load(const Root* r)
    {
        for (uint32_t i = 0; i < 8; ++i) { stones[i * stride] = r->stones[i >> 2][i & 3]; avail_word(i >> 2, i & 3) = r->avail[i >> 2][i & 3]; }
        for (uint32_t i = 0; i < 4; ++i) pool_word(i) = 0;
        m = r->m; n = r->n; k = r->k; player = r->player; res2 = r->result2;
        replay_left = 0; prep = nullptr;
        over = !avail_of(player).any();
    }

    // undoMoves, phantommnkgame.cpp:37-55, for observer = the player to move.
    ISMCTS_DEV static void prepare(const Root* r, Prepared* out, uint64_t*)
    {
        Prepared p;
        const uint32_t obs = r->player & 1u, opp = 1u - obs;
        p.hidden = 0;
        for (uint32_t i = 0; i < 4; ++i) {
            const uint64_t unseen = r->stones[opp][i] & r->avail[obs][i];   // opponent stones the observer has not run into
            p.hidden += popc64(unseen);
            p.stones[obs][i] = r->stones[obs][i];
            p.stones[opp][i] = r->stones[opp][i] & ~unseen;
            p.avail[obs][i] = r->avail[obs][i];
            p.avail[opp][i] = r->avail[opp][i] | unseen;
        }
        const uint32_t cells = (uint32_t)r->m * r->n;
        for (uint32_t i = 0; i < 4; ++i) {
            const uint32_t lo = 64u * i;
            const uint64_t board = cells >= lo + 64u ? ~(uint64_t)0 : cells > lo ? (((uint64_t)1 << (cells - lo)) - 1u) : 0;
            p.cand[i] = board & ~(p.stones[0][i] | p.stones[1][i]);
        }
        p.m = r->m; p.n = r->n; p.k = r->k; p.player = r->player; p.result2 = r->result2;
        p.pad[0] = p.pad[1] = 0;
        *out = p;
    }

    ISMCTS_DEV void restart_sample()
    {
        const uint32_t opp = 1u - prep->player;
        for (uint32_t i = 0; i < 4; ++i) {
            stones[(opp * 4 + i) * stride] = prep->stones[opp][i];
            avail_word(opp, i) = prep->avail[opp][i];
            pool_word(i) = prep->cand[i];
        }
        replay_left = prep->hidden;
    }

    ISMCTS_DEV void start(const Prepared* p)
    {
        prep = p;
        for (uint32_t i = 0; i < 8; ++i) { stones[i * stride] = p->stones[i >> 2][i & 3]; avail_word(i >> 2, i & 3) = p->avail[i >> 2][i & 3]; }
        for (uint32_t i = 0; i < 4; ++i) pool_word(i) = p->cand[i];
        m = p->m; n = p->n; k = p->k; player = p->player; res2 = p->result2;
        replay_left = p->hidden;
        over = !avail_of(player).any();
    }

    ISMCTS_DEV static uint32_t root_player(const Prepared* p) { return p->player; }
    ISMCTS_DEV static uint32_t root_players(const Prepared*) { return 3u; }

    ISMCTS_DEV bool chance_pending() const { return replay_left != 0; }
    ISMCTS_DEV Mask chance_set() const { return Mask{{pool_word(0), pool_word(1), pool_word(2), pool_word(3)}}; }

    // a pending chance event is one uniform draw: the j-th member of chance_set()
    ISMCTS_DEV uint32_t chance_count() const { return chance_set().count(); }
    ISMCTS_DEV void chance_draw(uint32_t j) { chance_step(chance_set().kth(j)); }

    // randomReplay, phantommnkgame.cpp:57-81: one hidden stone of the opponent.
    ISMCTS_DEV void chance_step(uint32_t cell)
    {
        const uint32_t me = player, opp = 1u - prep->player;
        const uint64_t bit = (uint64_t)1 << (cell & 63u);
        stones[(opp * 4 + (cell >> 6)) * stride] |= bit;
        player = opp;                                          // exact_run looks at the stones of `player`
        const bool wins = exact_run(cell, 0, 1) || exact_run(cell, 1, 0) || exact_run(cell, 1, 1) || exact_run(cell, -1, 1);
        player = me;
        if (wins) { restart_sample(); return; }                // a finished game cannot be the position: draw again
        avail_word(opp, cell >> 6) &= ~bit;
        pool_word(cell >> 6) &= ~bit;
        --replay_left;
    }

    // validMoves, phantommnkgame.cpp:107-110
    ISMCTS_DEV Mask legal() const { return avail_of(player); }
    ISMCTS_DEV bool terminal() const { return over != 0; }

    // doMove, phantommnkgame.cpp:83-105
    ISMCTS_DEV void play(uint32_t cell)
    {
        const uint64_t bit = (uint64_t)1 << (cell & 63u);
        const uint32_t w = cell >> 6;
        avail_word(player, w) &= ~bit;
        const bool occupied = ((stones[w * stride] | stones[(4 + w) * stride]) & bit) != 0;
        if (!occupied) {
            stones[(player * 4 + w) * stride] |= bit;
            if (exact_run(cell, 0, 1) || exact_run(cell, 1, 0) || exact_run(cell, 1, 1) || exact_run(cell, -1, 1)) {
                res2 = 2 * (1 - (int32_t)player);
                for (uint32_t i = 0; i < 8; ++i) avail_word(i >> 2, i & 3) = 0;
            } else {
                uint32_t filled = 0;
                for (uint32_t i = 0; i < 8; ++i) filled += popc64(stones[i * stride]);
                if (filled < m * n) player = 1u - player; else res2 = 1;
            }
        }
        over = !avail_of(player).any();
    }

    ISMCTS_DEV void step(uint32_t b) { if (chance_pending()) chance_step(b); else play(b); }
};

} // namespace ismcts
