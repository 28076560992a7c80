// kw_game.cuh — Knockout Whist as a device-side POD state machine.
//
// Device mirror of ISMCTS::Game<Card> for the reference's KnockoutWhist
// (test/common/knockoutwhist.{h,cpp}; rules restated in SURVEY.md appendix B).
// One GPU thread owns one game.  Card sets are 64-bit masks (bit id =
// rank + 13*suit, card.h:38-41); the seven hands live in shared memory, one
// 8-byte column per thread, everything else in registers.
//
// Every random event of the reference — the determinisation shuffle
// (cloneAndRandomise, knockoutwhist.cpp:56-76), the re-deal shuffle (deal,
// :138-152), the round-winner tie-break (roundWinner, :198-210) and the
// uniformly random move of the default policy (utility.h:69-83) — is expressed
// as ONE kind of micro-step: "remove a uniformly random member from a bit set".
// A micro-step draws exactly one Philox word.  Chance events that the reference
// resolves inside doMove become pending micro-steps (mode != kPlay) which are
// drained before the next decision, in the order the reference draws them
// (tie-break, then the deal player by player in ascending id).  That is what lets
// the 32 lanes of a warp, each in its own game and phase, run one loop body.
#pragma once

#include "devdefs.cuh"
#include "bitset.cuh"
#include "../../include/ismcts_b200.h"

namespace ismcts {

struct KwGame {
    static constexpr uint32_t kGameId = ISMCTS_GAME_KW;
    static constexpr uint32_t kMoveStride = 64;
    static constexpr uint32_t kHandWords = ISMCTS_KW_MAX_PLAYERS; // 8-byte words of per-thread scratch
    static constexpr uint32_t kMaxDepth = 208;  // plies are bounded by 6 + players * 28 (gametest.cpp:166-178)
    using Mask = Mask64;
    using Root = ismcts_kw_state;

    enum Mode : uint32_t { kPlay = 0, kDeal = 1, kTie = 2 };

    static constexpr uint64_t kFullDeck = (uint64_t(1) << 52) - 1;
    static constexpr uint64_t kTrumpChoices = uint64_t(1) | (uint64_t(1) << 13) | (uint64_t(1) << 26) | (uint64_t(1) << 39);
    static constexpr uint64_t kSuit = 0x1FFF;

    // per-thread hands: hand of player p is hands[p * stride]
    uint64_t* hands;
    uint32_t stride;

    uint32_t player, dealer, trump, tricks_left, request_trump;
    uint32_t alive;      // bit p: player p is still in (m_players)
    uint32_t tricks;     // 4 bits per player (m_tricksTaken)
    uint32_t trick_len;  // cards in the current trick
    uint32_t lead_suit;  // suit of its first card
    uint32_t win_card;   // card currently winning it
    uint32_t win_player; // who played that card
    uint32_t mode;       // pending chance event, if any
    uint64_t pool;       // kDeal: cards not dealt yet; kTie: tied players
    uint32_t deal_player, deal_left; // kDeal: who is being dealt to and how many more cards
    uint32_t need;       // kDeal: 4 bits per player, cards the players after deal_player still get

    ISMCTS_DEV uint64_t hand(uint32_t p) const { return hands[p * stride]; }
    ISMCTS_DEV void set_hand(uint32_t p, uint64_t v) { hands[p * stride] = v; }

    ISMCTS_DEV void bind(uint64_t* scratch, uint32_t scratch_stride) { hands = scratch; stride = scratch_stride; }

    // Loads a root position (knockoutwhist.h:36-48 as POD).
    ISMCTS_DEV void load(const Root* r)
    {
        for (uint32_t p = 0; p < ISMCTS_KW_MAX_PLAYERS; ++p) set_hand(p, r->hands[p]);
        player = r->player; dealer = r->dealer; trump = r->trump;
        tricks_left = r->tricks_left; request_trump = r->request_trump;
        alive = r->alive_mask;
        tricks = 0;
        for (uint32_t p = 0; p < ISMCTS_KW_MAX_PLAYERS; ++p) tricks |= (uint32_t)(r->tricks_taken[p] & 15u) << (4 * p);
        trick_len = r->trick_len;
        lead_suit = 0; win_card = 0; win_player = 0;
        // fold the trick so far into its running winner (trickWinner, knockoutwhist.cpp:182-196)
        for (uint32_t i = 0; i < trick_len; ++i) fold_trick(i, r->trick_player[i], r->trick_card[i]);
        mode = kPlay; pool = 0; deal_player = 0; deal_left = 0; need = 0;
    }

    ISMCTS_DEV void fold_trick(uint32_t index, uint32_t who, uint32_t card)
    {
        const uint32_t suit = card / 13u;
        if (index == 0) { lead_suit = suit; win_card = card; win_player = who; return; }
        const uint32_t win_suit = win_card / 13u;
        const bool beats = (suit == win_suit) ? (card > win_card) : (suit == trump);
        if (beats) { win_card = card; win_player = who; }
    }

    ISMCTS_DEV uint32_t current_player() const { return player; }
    ISMCTS_DEV bool chance_pending() const { return mode != kPlay; }

    // nextPlayer, knockoutwhist.cpp:83-88: next id, cyclically, that is still in.
    ISMCTS_DEV uint32_t next_alive(uint32_t p) const
    {
        const uint32_t above = alive & ~((2u << p) - 1u);
        return ctz32(above ? above : alive);
    }

    // validMoves, knockoutwhist.cpp:122-136 (empty <=> the game is over).  Only
    // meaningful while no chance event is pending.
    ISMCTS_DEV Mask legal() const
    {
        if (request_trump) return Mask{kTrumpChoices};
        const uint64_t h = hand(player);
        if (trick_len == 0 || popc64(h) < 2) return Mask{h};
        const uint64_t follow = h & (kSuit << (13u * lead_suit));
        return Mask{follow ? follow : h};
    }

    // The set a pending chance micro-step draws from.
    ISMCTS_DEV Mask chance_set() const { return Mask{pool}; }

    // getResult * 2, knockoutwhist.cpp:117-120: the first remaining player has won.
    ISMCTS_DEV uint32_t result2(uint32_t who) const { return who == ctz32(alive) ? 2u : 0u; }
    ISMCTS_DEV uint32_t outcome() const { return ctz32(alive); }

    // Starts dealing from `pool`: `sizes` holds 4 bits per player, the number of
    // cards each gets, ascending id.
    ISMCTS_DEV void begin_deal(uint64_t from, uint32_t sizes)
    {
        if (sizes == 0) { mode = kPlay; return; }
        pool = from;
        const uint32_t first = ctz32(sizes) >> 2;
        deal_player = first;
        deal_left = (sizes >> (4 * first)) & 15u;
        need = sizes & ~(15u << (4 * first));
        mode = kDeal;
    }

    // deal, knockoutwhist.cpp:138-152: a fresh deck, tricks_left cards to each remaining player.
    ISMCTS_DEV void start_deal()
    {
        uint32_t sizes = 0;
        for (uint32_t p = 0; p < ISMCTS_KW_MAX_PLAYERS; ++p)
            if ((alive >> p) & 1u) sizes |= tricks_left << (4 * p);
        begin_deal(kFullDeck, sizes);
    }

    // cloneAndRandomise, knockoutwhist.cpp:56-76: what the observer cannot see (the
    // unknown cards and the other players' hands) is pooled and dealt back to the
    // other players, ascending id, keeping their hand sizes.
    ISMCTS_DEV void begin_determinise(uint32_t observer, const Root* r)
    {
        uint64_t unseen = r->unknown;
        uint32_t sizes = 0;
        for (uint32_t p = 0; p < ISMCTS_KW_MAX_PLAYERS; ++p) {
            if (!((alive >> p) & 1u) || p == observer) continue;
            const uint64_t h = hand(p);
            unseen |= h;
            sizes |= popc64(h) << (4 * p);
            set_hand(p, 0);
        }
        begin_deal(unseen, sizes);
    }

    // finishRound, knockoutwhist.cpp:163-180
    ISMCTS_DEV void finish_round()
    {
        uint32_t still = 0, best = 0;
        for (uint32_t p = 0; p < ISMCTS_KW_MAX_PLAYERS; ++p) {
            const uint32_t t = (tricks >> (4 * p)) & 15u;
            if (((alive >> p) & 1u) && t > 0) { still |= 1u << p; best = t > best ? t : best; }
        }
        alive = still;
        if (popc32(alive) > 1) {
            --tricks_left;
            request_trump = 1;
            uint32_t tied = 0; // roundWinner, :198-210: uniform among the players with most tricks
            for (uint32_t p = 0; p < ISMCTS_KW_MAX_PLAYERS; ++p)
                if (((alive >> p) & 1u) && ((tricks >> (4 * p)) & 15u) == best) tied |= 1u << p;
            pool = tied;
            mode = kTie;
        } else {
            tricks_left = 0; // over: nobody holds cards, validMoves() is empty
        }
    }

    // One micro-step: `b` is the drawn member of chance_set() (chance pending) or the move played.
    ISMCTS_DEV void step(uint32_t b)
    {
        if (mode == kDeal) {
            const uint64_t bit = (uint64_t)1 << b;
            pool &= ~bit;
            set_hand(deal_player, hand(deal_player) | bit);
            if (--deal_left == 0) {
                if (need) {
                    const uint32_t next = ctz32(need) >> 2;
                    deal_player = next;
                    deal_left = (need >> (4 * next)) & 15u;
                    need &= ~(15u << (4 * next));
                } else {
                    mode = kPlay;
                }
            }
            return;
        }
        if (mode == kTie) {
            player = b;                    // the round winner calls trumps
            dealer = next_alive(dealer);
            tricks = 0;
            start_deal();
            return;
        }
        // doMove, knockoutwhist.cpp:95-115
        if (request_trump) {
            trump = b / 13u;
            request_trump = 0;
            player = next_alive(dealer);
            return;
        }
        set_hand(player, hand(player) & ~((uint64_t)1 << b));
        fold_trick(trick_len, player, b);
        ++trick_len;
        player = next_alive(player);
        if (trick_len == popc32(alive)) { // finishTrick, :155-161
            tricks += 1u << (4 * win_player);
            trick_len = 0;
            player = win_player;
        }
        if (hand(player) == 0) finish_round();
    }
};

} // namespace ismcts
