// goof_game.cuh — two-player Goofspiel as a device-side POD state machine.
//
// Device mirror of ISMCTS::POMGame<Card> for the reference's Goofspiel (test/common/goofspiel.{h,cpp}):
// the prizes (hearts) are turned over one by one by the environment (player 2), players 0 (spades) and 1
// (clubs) bid a card each, the higher value takes the prize; every move counts as simultaneous
// (goofspiel.cpp:80-83), so the searches select with EXP3.  The state fits in registers.
//
// Chance events as micro-steps (see kw_game.cuh): the order of the remaining prizes is hidden from the
// players (cloneAndRandomise reshuffles them, goofspiel.cpp:69-71), which is restated as "the next prize
// is a uniformly random remaining prize", drawn when the environment is about to turn it over; for
// player 1 the bid player 0 has just made is hidden too (goofspiel.cpp:76-78) and is drawn at the start
// of the iteration.  When the environment itself is the observer the order is known: the draw is from the
// one-element set holding the next prize.
#pragma once

#include "devdefs.cuh"
#include "bitset.cuh"
#include "../../include/ismcts_b200.h"

namespace ismcts {

struct GoofGame {
    static constexpr uint32_t kGameId = ISMCTS_GAME_GOOF;
    static constexpr uint32_t kMoveStride = 64;
    static constexpr uint32_t kHandWords = 1;    // no shared-memory scratch needed
    static constexpr uint32_t kMaxPlayers = 3;   // players(), goofspiel.cpp:124-128
    static constexpr uint32_t kMaxLegal = 13;
    static constexpr uint32_t kMaxDepth = 64;    // 13 turns of 4 moves
    using Mask = Mask64;
    using Root = ismcts_goof_state;
    using Prepared = ismcts_goof_state;

    static constexpr uint32_t kBid0 = 39, kBid1 = 0, kPrize = 26; // move id = rank + offset; the reveal move is 0

    uint32_t hands;         // player 0 in bits 0..12, player 1 in bits 16..28
    uint32_t prizes;        // ranks not turned over yet
    uint64_t order;         // 4 bits per remaining prize, entry n-1 next; meaningful while the order is known
    uint32_t known_order;   // the observer is the environment
    uint32_t next_prize;    // rank + 1 of the prize about to be turned over, 0 = not determined
    uint32_t hidden_bid;    // observer 1: player 0's bid is still to be drawn
    uint32_t current_prize, move0, move1, score0, score1, player, draw_prize;

    ISMCTS_DEV void bind(uint64_t*, uint32_t, uint32_t) { hidden_bid = 0; next_prize = 0; player = 2; draw_prize = 1; prizes = 0; }
    ISMCTS_DEV void rebind(uint64_t*, uint32_t, uint32_t) {}
    ISMCTS_DEV static void init_block(uint32_t, uint32_t) {}   // no per-block tables

    ISMCTS_DEV static uint32_t value(uint32_t rank) { return rank == 12u ? 1u : rank + 2u; } // goofspiel.cpp:17-36

    ISMCTS_DEV void load(const Root* r)
    {
        hands = (uint32_t)r->hands[0] | ((uint32_t)r->hands[1] << 16);
        order = r->prize_order;
        prizes = 0;
        for (uint32_t i = 0; i < r->n_prizes; ++i) prizes |= 1u << ((order >> (4 * i)) & 15u);
        known_order = 1; next_prize = 0; hidden_bid = 0;
        current_prize = r->current_prize; move0 = r->moves[0]; move1 = r->moves[1];
        score0 = r->scores[0]; score1 = r->scores[1]; player = r->player; draw_prize = r->draw_prize;
    }

    ISMCTS_DEV static void prepare(const Root* r, Prepared* out, uint64_t*) { *out = *r; }

    // cloneAndRandomise(observer = the player to move), goofspiel.cpp:65-81
    ISMCTS_DEV void start(const Prepared* p)
    {
        load(p);
        known_order = player == 2u;
        hidden_bid = player == 1u;
    }

    ISMCTS_DEV uint32_t current_player() const { return player; }
    ISMCTS_DEV static uint32_t root_player(const Prepared* p) { return p->player; }
    ISMCTS_DEV static uint32_t root_players(const Prepared*) { return 7u; }

    ISMCTS_DEV uint32_t n_prizes() const { return popc32(prizes); }
    ISMCTS_DEV bool draw_pending() const { return player == 2u && draw_prize && next_prize == 0u && prizes != 0u; }
    ISMCTS_DEV bool chance_pending() const { return hidden_bid != 0u || draw_pending(); }

    ISMCTS_DEV Mask chance_set() const
    {
        if (hidden_bid) return Mask{(uint64_t)(hands & 0x1FFFu)};
        if (known_order) return Mask{(uint64_t)1 << ((order >> (4 * (n_prizes() - 1u))) & 15u)};
        return Mask{(uint64_t)prizes};
    }

    // a pending chance event is one uniform draw: the j-th member of chance_set()
    ISMCTS_DEV uint32_t chance_count() const { return chance_set().count(); }
    ISMCTS_DEV void chance_draw(uint32_t j) { chance_step(chance_set().kth(j)); }

    ISMCTS_DEV void chance_step(uint32_t b)
    {
        if (hidden_bid) { move0 = b; hidden_bid = 0; return; }
        prizes &= ~(1u << b);
        next_prize = b + 1u;
    }

    // validMoves, goofspiel.cpp:130-153
    ISMCTS_DEV Mask legal() const
    {
        if (player == 0u) return Mask{(uint64_t)(hands & 0x1FFFu) << kBid0};
        if (player == 1u) return Mask{(uint64_t)((hands >> 16) & 0x1FFFu) << kBid1};
        if (!draw_prize) return Mask{1};
        return Mask{next_prize ? (uint64_t)1 << (kPrize + next_prize - 1u) : 0};
    }

    ISMCTS_DEV bool terminal() const { return player == 2u && draw_prize && next_prize == 0u && prizes == 0u; }

    // doMove + handleP2Turn, goofspiel.cpp:90-117
    ISMCTS_DEV void play(uint32_t b)
    {
        if (player == 0u) { move0 = b - kBid0; player = 1; return; }
        if (player == 1u) { move1 = b - kBid1; player = 2; return; }
        if (draw_prize) {
            current_prize = next_prize - 1u;
            next_prize = 0; player = 0; draw_prize = 0;
            return;
        }
        if (move0 != move1) {
            if (value(move0) > value(move1)) score0 += value(current_prize); else score1 += value(current_prize);
        }
        hands &= ~((1u << move0) | (1u << (16u + move1)));
        draw_prize = 1;
    }

    ISMCTS_DEV void step(uint32_t b) { if (chance_pending()) chance_step(b); else play(b); }

    // getResult * 2, goofspiel.cpp:119-128
    ISMCTS_DEV uint32_t result2(uint32_t who) const
    {
        if (who == 2u) return 0;
        if (score0 == score1) return 1;
        return (who == 0u) == (score0 > score1) ? 2u : 0u;
    }
    ISMCTS_DEV uint32_t outcome() const { return result2(0); }
};

} // namespace ismcts
