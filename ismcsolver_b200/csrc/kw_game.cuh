// kw_game.cuh — Knockout Whist as a device-side POD state machine.
//
// Device mirror of ISMCTS::Game<Card> for the reference's KnockoutWhist
// (test/common/knockoutwhist.{h,cpp}; rules restated in SURVEY.md appendix B).
// One GPU thread owns one game.  Everything but the cards lives in registers;
// sets of players and the trick counts are one 4-bit field per player, so that
// the round-end bookkeeping (knock-outs, most tricks, ties, deal sizes) is a
// handful of word-parallel operations instead of loops over the players.
//
// Cards.  Moves of the tree and of the C ABI are card ids rank + 13 * suit
// (card.h:38-41); inside the game a card is the code (suit << 4) | rank.  A HAND
// is seven places (the game never deals more than seven cards): `list` holds the
// code put on each place (one byte per place), `held` has bit 8 * suit + place
// set while the card on that place has not been played.  A card is put on the
// first place after the last card still held, a played card leaves its place
// empty.  With this layout the cards that follow suit are one shift and mask of
// `held`, "the j-th legal card" is a look-up in a 128-entry table of bit
// positions, and playing a card clears one bit — the default policy (simulate,
// solverbase.h:36-43) never searches a 64-bit set.  The hands live in shared
// memory, one column per thread.
//
// The chance events the reference resolves inside its functions — the
// determinisation shuffle (cloneAndRandomise, knockoutwhist.cpp:56-76), the
// re-deal shuffle (deal, :138-152) and the round-winner tie-break (roundWinner,
// :198-210) — become pending steps (mode != kPlay) that are drained before the
// next decision, in the order the reference draws (tie-break, then the deal
// player by player in ascending id), one uniform draw each.  Cards are dealt from
// a PILE, a byte array in shared memory that starts in ascending card order: a
// uniformly random card leaves it and the last card of the pile takes its place
// (the steps of a Fisher-Yates shuffle of which only the dealt cards are kept; a
// shuffle followed by "take the first k" has the same distribution).  The default
// policy plays the j-th legal card in place order (any fixed order is uniform).
// The stream is aligned to a Philox block before and after every sequence of
// chance events and at the start of the rollout (philox.cuh), so that the lanes of
// a warp, which go through the phases of an iteration together (lane_search.cuh),
// generate their blocks in the same steps: chance_phase() and play_phase() below
// keep the current block in registers.  oracle/oracle.c states all of this
// sequentially (hand_append, okw_rollout_move, pile_draw, rng_align).
//
// Where a game lives.  The game object is 14 words and holds no pointer: its scratch (hands,
// pile) is the column of its thread in the block's scratch (devdefs.cuh: block_scratch).
// While a deal is in progress its state is a separate `Deal` that lives in the loop that deals
// (chance_phase); between two such loops no deal is ever in progress.  That makes the pile dead
// whenever the tree operations of the search run, and the search keeps the game object in the 56
// bytes of its pile meanwhile (home()), i.e. in shared memory at no cost, instead of in the thread's
// local memory.
#pragma once

#include <type_traits>
#include "devdefs.cuh"
#include "bitset.cuh"
#include "../../include/ismcts_b200.h"

namespace ismcts {

// What every iteration of a search starts from: the root position as the lane's
// registers want it, with the determinisation (cloneAndRandomise for the root's
// player to move) already set up as a pending deal.  Built once per position.
struct alignas(16) KwPrepared {
    uint64_t list_obs;   // the observer's own hand (the only one it can see): the places ...
    uint32_t held_obs;   // ... and which of them hold a card
    uint32_t sizes;      // 4 bits per player: cards to deal back (0 for the observer)
    uint32_t observer;
    uint32_t player, dealer, trump, tricks_left, request_trump;
    uint32_t alive4, tricks, best;
    uint32_t trick_len, lead_suit, win_key, win_player;
    uint32_t pile_n;     // unknown cards + the other players' hands: the pile to deal back from
    uint64_t pile[7];    // its cards (codes), one per byte, ascending
};
static_assert(sizeof(KwPrepared) == 128, "eight 16-byte loads");

// The place of the j-th set bit of a 7-bit set, 3 bits per j: the table behind "the j-th legal
// card".  On the device it lives in the block's dynamic shared memory right behind the scratch of the
// games (filled by init_block at the start of a kernel), so that its address derives from the same base as
// theirs; host builds use one static table.
struct KwPlaceTable {
    static constexpr uint32_t kBytes = 128 * sizeof(uint32_t);
    ISMCTS_DEV static uint32_t entry(uint32_t set)
    {
        uint32_t e = 0, j = 0;
        for (uint32_t place = 0; place < 7; ++place)
            if ((set >> place) & 1u) e |= place << (3u * j++);
        return e;
    }
    // `behind`: the end of the games' scratch
    ISMCTS_DEV static uint32_t* table(uint64_t* behind)
    {
#if defined(__CUDA_ARCH__)
        return reinterpret_cast<uint32_t*>(behind);
#else
        (void)behind;
        static uint32_t t[128];
        static bool ready = false;
        if (!ready) { for (uint32_t s = 0; s < 128; ++s) t[s] = entry(s); ready = true; }
        return t;
#endif
    }
    // Called by every thread of a block before any game runs (followed by a block barrier).
    ISMCTS_DEV static void init_block(uint64_t* behind, uint32_t thread, uint32_t threads)
    {
#if defined(__CUDA_ARCH__)
        uint32_t* t = table(behind);
        for (uint32_t s = thread; s < 128; s += threads) t[s] = entry(s);
#else
        (void)behind; (void)thread; (void)threads;
#endif
    }
};

// kSeats: the highest player id + 1 the scratch has room for (7; 4 for positions whose records say so:
// less shared memory per block leaves more of it to L1).
// kCarry: a copy of the game that carries the addresses of its scratch with it (the loops of the phases work
// on such copies, `Hot`: the addresses are values the compiler keeps or derives from the thread index as it
// likes); the game proper (`Core`) finds its scratch from the thread index whenever it needs it.
struct KwNoAddresses {};
struct KwAddresses { uint64_t* hands_; uint32_t* helds_; uint64_t* pile_; const uint32_t* table_; };

template <uint32_t kSeats, bool kCarry = false>
struct KwGameT : std::conditional_t<kCarry, KwAddresses, KwNoAddresses> {
    using Core = KwGameT<kSeats, false>;
    using Hot = KwGameT<kSeats, true>;
    static constexpr uint32_t kGameId = ISMCTS_GAME_KW;
    static constexpr uint32_t kMoveStride = 64;
    static constexpr uint32_t kPlaces = 7;                          // cards a hand can hold
    static constexpr uint32_t kPileWords = 7;                       // 52 cards, one per byte
    // 8-byte words of per-thread scratch: one list per player, the pile, `held` of the players (4 bytes each)
    static constexpr uint32_t kHeldRow = kSeats;
    static constexpr uint32_t kPileRow = kHeldRow + (kSeats + 1) / 2;   // (the last rows: home())
    static constexpr uint32_t kHandWords = kPileRow + kPileWords;
    static constexpr uint32_t kMaxPlayers = ISMCTS_KW_MAX_PLAYERS;
    static constexpr uint32_t kMaxLegal = ISMCTS_KW_CARDS;
    static constexpr uint32_t kMaxDepth = 208;  // plies are bounded by 6 + players * 28 (gametest.cpp:166-178)
    using Mask = Mask64;
    using Root = ismcts_kw_state;
    using Prepared = KwPrepared;

    // kStart: the determinisation has not begun (start()); kRoundEnd: inside play_phase only, the round is over
    enum Mode : uint32_t { kPlay = 0, kDeal = 1, kTie = 2, kStart = 3, kRoundEnd = 4 };

    static constexpr uint64_t kFullDeck = (uint64_t(1) << 52) - 1;
    static constexpr uint64_t kTrumpChoices = uint64_t(1) | (uint64_t(1) << 13) | (uint64_t(1) << 26) | (uint64_t(1) << 39);
    static constexpr uint32_t kOnes = 0x1111111u; // bit 4p for each of the 7 players

    // The scratch of this thread's game.  Lists and `held`: rows of one column per thread of the block —
    // list of player p = hands()[p * kStride], `held` of player p = helds()[p * kStride] (rows of 4-byte words
    // behind the rows of lists).  The pile: 56 consecutive bytes per thread behind those (card i = byte i:
    // no row arithmetic for a card picked at random; consecutive threads are seven 8-byte words apart, which
    // keeps the 8-byte accesses of a warp free of bank conflicts).
    static constexpr uint32_t kStride = kBlockLanes;
    ISMCTS_DEV static uint64_t* hands_of(uint64_t* scratch, uint32_t column) { return scratch + column; }
    ISMCTS_DEV static uint32_t* helds_of(uint64_t* scratch, uint32_t column) { return reinterpret_cast<uint32_t*>(scratch + kHeldRow * kStride) + column; }
    ISMCTS_DEV static uint64_t* pile_of(uint64_t* scratch, uint32_t column) { return scratch + kPileRow * kStride + column * kPileWords; }
    ISMCTS_DEV uint64_t* hands() const { if constexpr (kCarry) return this->hands_; else return hands_of(block_scratch(), block_lane()); }
    ISMCTS_DEV uint32_t* helds() const { if constexpr (kCarry) return this->helds_; else return helds_of(block_scratch(), block_lane()); }
    ISMCTS_DEV uint64_t* pile() const { if constexpr (kCarry) return this->pile_; else return pile_of(block_scratch(), block_lane()); }

    uint32_t player, dealer, trump, tricks_left, request_trump;
    uint32_t alive4;     // bit 4p: player p is still in (m_players)
    uint32_t tricks;     // 4 bits per player (m_tricksTaken)
    uint32_t best;       // most tricks taken by one player in this round
    uint32_t trick_len;  // cards in the current trick
    uint32_t lead_suit;  // suit of its first card
    uint32_t win_key;    // strength of the card currently winning it (card_key)
    uint32_t win_player; // who played that card
    uint32_t mode;       // pending chance event, if any
    uint32_t tied;       // kTie: the players tied on most tricks (bit 4p)

    // A deal in progress (mode == kDeal): lives in the loop that deals, see the head of this file.
    struct Deal {
        uint32_t pile_n;                  // cards left in the pile
        uint32_t deal_player, deal_left;  // who is being dealt to and how many more cards
        uint32_t need;                    // 4 bits per player, cards the players after deal_player still get
        uint32_t deal_place;              // the place the next card of deal_player goes to
        uint32_t deal_held;               // `held` of deal_player while it is being dealt, stored when it is complete
    };

    // Where the search keeps this thread's game while no deal can be in progress: the 56 bytes of its pile.
    ISMCTS_DEV static Core* home() { return reinterpret_cast<Core*>(pile_of(block_scratch(), block_lane())); }
    static constexpr bool kHomeInScratch = true;

    // the state of another copy of the game
    template <bool kOther>
    ISMCTS_DEV void take(const KwGameT<kSeats, kOther>& o)
    {
        player = o.player; dealer = o.dealer; trump = o.trump; tricks_left = o.tricks_left; request_trump = o.request_trump;
        alive4 = o.alive4; tricks = o.tricks; best = o.best; trick_len = o.trick_len; lead_suit = o.lead_suit;
        win_key = o.win_key; win_player = o.win_player; mode = o.mode; tied = o.tied;
    }
    KwGameT() = default;
    template <bool kOther, bool kMine = kCarry, class = std::enable_if_t<kMine != kOther>>
    ISMCTS_DEV explicit KwGameT(const KwGameT<kSeats, kOther>& o) { take(o); }
    template <bool kOther, bool kMine = kCarry, class = std::enable_if_t<kMine != kOther>>
    ISMCTS_DEV KwGameT& operator=(const KwGameT<kSeats, kOther>& o) { take(o); return *this; }

    // ---- cards and hands

    ISMCTS_DEV static uint32_t code_of_id(uint32_t id)
    {
        const uint32_t suit = (id * 79u) >> 10;   // id / 13, exact below 64
        return (suit << 4) | (id - 13u * suit);
    }
    ISMCTS_DEV static uint32_t id_of_code(uint32_t code) { return (code >> 4) * 13u + (code & 15u); }
    // the places that hold a card, whatever its suit
    ISMCTS_DEV static uint32_t occupied(uint32_t held)
    {
        held |= held >> 16;
        held |= held >> 8;
        return held & 0x7Fu;
    }
    // one past the last place that holds a card
    ISMCTS_DEV static uint32_t top_place(uint32_t held) { return 32u - clz32(occupied(held)); }
    ISMCTS_DEV static uint32_t code_at(uint64_t list, uint32_t place) { return (uint32_t)(list >> (8u * place)) & 0xFFu; }

    ISMCTS_DEV uint64_t list_of(uint32_t p) const { return hands()[p * kStride]; }
    ISMCTS_DEV void set_list(uint32_t p, uint64_t v) const { hands()[p * kStride] = v; }
    ISMCTS_DEV uint32_t held_of(uint32_t p) const { return helds()[p * kStride]; }
    ISMCTS_DEV void set_held(uint32_t p, uint32_t v) const { helds()[p * kStride] = v; }

    // a hand as a set of card ids (the C ABI, the tree)
    ISMCTS_DEV static uint64_t ids_of(uint64_t list, uint32_t places)
    {
        uint64_t out = 0;
        for (; places; places &= places - 1) out |= (uint64_t)1 << id_of_code(code_at(list, ctz32(places)));
        return out;
    }
    ISMCTS_DEV uint64_t hand_ids(uint32_t p) const { return ids_of(list_of(p), occupied(held_of(p))); }
    // ... and the other way round: ascending ids on the places 0, 1, ... (a hand holds at most seven cards)
    ISMCTS_DEV void set_hand_ids(uint32_t p, uint64_t ids)
    {
        uint64_t list = 0;
        uint32_t held = 0, place = 0;
        for (ids &= kFullDeck; ids && place < kPlaces; ids &= ids - 1, ++place) {
            const uint32_t code = code_of_id(ctz64(ids));
            list |= (uint64_t)code << (8u * place);
            held |= 1u << (8u * (code >> 4) + place);
        }
        set_list(p, list);
        set_held(p, held);
    }

    // card i of the pile: byte i % 8 of word i / 8
    ISMCTS_DEV uint8_t* pile_at(uint32_t i) const
    {
        return reinterpret_cast<uint8_t*>(pile()) + i;
    }
    ISMCTS_DEV void set_pile_word(uint32_t i, uint64_t v) const { pile()[i] = v; }
    // the cards left in the pile as a set of ids (after a deal: the cards nobody holds)
    ISMCTS_DEV uint64_t pile_set(const Deal& d) const
    {
        uint64_t out = 0;
        for (uint32_t i = 0; i < d.pile_n; ++i) out |= (uint64_t)1 << id_of_code(*pile_at(i));
        return out;
    }
    // word i of the full deck in ascending order (codes)
    ISMCTS_DEV static constexpr uint64_t deck_word(uint32_t i)
    {
        uint64_t w = 0;
        for (uint32_t b = 0; b < 8; ++b) {
            const uint32_t id = 8u * i + b;
            if (id < ISMCTS_KW_CARDS) w |= (uint64_t)(((id / 13u) << 4) | (id % 13u)) << (8u * b);
        }
        return w;
    }

    static constexpr uint32_t kBlockExtraBytes = KwPlaceTable::kBytes;       // behind the scratch of the block's games
    ISMCTS_DEV const uint32_t* place_table() const
    {
        if constexpr (kCarry) return this->table_; else return KwPlaceTable::table(block_scratch() + kHandWords * kStride);
    }
    ISMCTS_DEV static void init_block(uint32_t thread, uint32_t threads) { KwPlaceTable::init_block(block_scratch() + kHandWords * kStride, thread, threads); }

    // (the scratch is found from the thread index: nothing to bind)
    ISMCTS_DEV void bind(uint64_t*, uint32_t, uint32_t) { mode = kPlay; }
    // kCarry: `scratch` is the block's scratch and `column` the index of the thread in its block
    ISMCTS_DEV void rebind(uint64_t* scratch, uint32_t, uint32_t column)
    {
        if constexpr (kCarry) {
#if defined(__CUDA_ARCH__)
            // (the column as a value the compiler cannot read again from the thread index: S2R has a long latency)
            asm volatile("" : "+r"(column));
#endif
            this->hands_ = hands_of(scratch, column); this->helds_ = helds_of(scratch, column); this->pile_ = pile_of(scratch, column);
            this->table_ = KwPlaceTable::table(scratch + kHandWords * kStride);
        }
    }

    // bit p -> bit 4p
    ISMCTS_DEV static uint32_t spread4(uint32_t bits)
    {
        uint32_t out = 0;
        for (uint32_t p = 0; p < ISMCTS_KW_MAX_PLAYERS; ++p) out |= ((bits >> p) & 1u) << (4 * p);
        return out;
    }
    // bit 4p set for every non-zero 4-bit field of x
    ISMCTS_DEV static uint32_t nonzero4(uint32_t x)
    {
        x |= x >> 1;
        x |= x >> 2;
        return x & kOnes;
    }

    // trickWinner, knockoutwhist.cpp:182-196, as a key: a later card beats the
    // winning one if it has the same suit and a higher rank, or another suit that
    // is trumps.  With (trumps 2, suit led 1, other 0) * 16 + rank that is
    // "strictly larger key" (the first card of a trick is of the suit led).
    ISMCTS_DEV void fold_trick(uint32_t index, uint32_t who, uint32_t code)
    {
        const uint32_t suit = code >> 4, rank = code & 15u;
        if (index == 0) lead_suit = suit;
        const uint32_t key = rank + (suit == trump ? 32u : suit == lead_suit ? 16u : 0u);
        if (index == 0 || key > win_key) { win_key = key; win_player = who; }
    }

    // Loads a root position (knockoutwhist.h:36-48 as POD).
    ISMCTS_DEV void load(const Root* r)
    {
        for (uint32_t p = 0; p < kSeats; ++p) set_hand_ids(p, r->hands[p]);
        player = r->player; dealer = r->dealer; trump = r->trump;
        tricks_left = r->tricks_left; request_trump = r->request_trump;
        alive4 = spread4(r->alive_mask);
        tricks = 0; best = 0;
        for (uint32_t p = 0; p < ISMCTS_KW_MAX_PLAYERS; ++p) {
            const uint32_t t = r->tricks_taken[p] & 15u;
            tricks |= t << (4 * p);
            if (((r->alive_mask >> p) & 1u) && t > best) best = t;
        }
        trick_len = r->trick_len;
        lead_suit = 0; win_key = 0; win_player = 0;
        for (uint32_t i = 0; i < trick_len; ++i) fold_trick(i, r->trick_player[i], code_of_id(r->trick_card[i]));
        mode = kPlay; tied = 0;
    }

    // The start of an iteration for `observer` = the root's player to move.
    ISMCTS_DEV static void prepare(const Root* r, Prepared* out, uint64_t* scratch)
    {
        (void)scratch;                                 // (the game finds its scratch itself: block_scratch)
        Core g;
        g.load(r);
        Prepared p;
        p.observer = g.player;
        p.list_obs = g.list_of(g.player);
        p.held_obs = g.held_of(g.player);
        // cloneAndRandomise, knockoutwhist.cpp:56-76: what the observer cannot see (the unknown
        // cards and the other players' hands) is pooled and dealt back to the other players,
        // ascending id, keeping their hand sizes
        uint64_t unseen = r->unknown;
        p.sizes = 0;
        for (uint32_t q = 0; q < kSeats; ++q) {
            if (!((g.alive4 >> (4 * q)) & 1u) || q == g.player) continue;
            const uint64_t ids = g.hand_ids(q);
            unseen |= ids;
            p.sizes |= popc64(ids) << (4 * q);
        }
        for (uint32_t i = 0; i < kPileWords; ++i) p.pile[i] = 0;
        p.pile_n = 0;
        for (unseen &= kFullDeck; unseen; unseen &= unseen - 1) {
            p.pile[p.pile_n >> 3] |= (uint64_t)code_of_id(ctz64(unseen)) << (8 * (p.pile_n & 7u));
            ++p.pile_n;
        }
        if (!g.legal().any()) g.tricks_left = 0;      // a finished game, however it is recorded
        p.player = g.player; p.dealer = g.dealer; p.trump = g.trump; p.tricks_left = g.tricks_left;
        p.request_trump = g.request_trump; p.alive4 = g.alive4; p.tricks = g.tricks; p.best = g.best;
        p.trick_len = g.trick_len; p.lead_suit = g.lead_suit; p.win_key = g.win_key; p.win_player = g.win_player;
        *out = p;
    }

    // Clone of the root with the determinisation pending (sosolver.h:46).  Does not touch the pile
    // (see home()): the loop that deals begins with begin_start().
    ISMCTS_DEV void start(const Prepared* q)
    {
        for (uint32_t p = 0; p < kSeats; ++p) set_held(p, 0);                     // nobody holds a card ...
        set_list(q->observer, q->list_obs);                                        // ... but the observer
        set_held(q->observer, q->held_obs);
        player = q->player; dealer = q->dealer; trump = q->trump; tricks_left = q->tricks_left;
        request_trump = q->request_trump; alive4 = q->alive4; tricks = q->tricks; best = q->best;
        trick_len = q->trick_len; lead_suit = q->lead_suit; win_key = q->win_key; win_player = q->win_player;
        mode = kStart; tied = 0;
    }
    // mode == kStart: the pile of the determinisation and its deal (cloneAndRandomise, knockoutwhist.cpp:56-76)
    ISMCTS_DEV void begin_start(Deal& d, const Prepared* q)
    {
        for (uint32_t i = 0; i < kPileWords; ++i) set_pile_word(i, q->pile[i]);
        d.pile_n = q->pile_n;
        begin_deal(d, q->sizes);
    }

    ISMCTS_DEV uint32_t current_player() const { return player; }
    ISMCTS_DEV static uint32_t root_player(const Prepared* p) { return p->player; }
    // players(), knockoutwhist.cpp:90-93, of the root position as a bit set
    ISMCTS_DEV static uint32_t root_players(const Prepared* p)
    {
        uint32_t out = 0;
        for (uint32_t q = 0; q < ISMCTS_KW_MAX_PLAYERS; ++q) out |= ((p->alive4 >> (4 * q)) & 1u) << q;
        return out;
    }
    ISMCTS_DEV bool chance_pending() const { return mode != kPlay; }

    // nextPlayer, knockoutwhist.cpp:83-88: next id, cyclically, that is still in.
    ISMCTS_DEV uint32_t next_alive(uint32_t p) const
    {
        const uint32_t above = alive4 & ~((2u << (4 * p)) - 1u);
        return ctz32(above ? above : alive4) >> 2;
    }

    // The places of the legal cards of a hand: those that follow suit if there are any, else all.
    ISMCTS_DEV uint32_t legal_places(uint32_t held) const
    {
        const uint32_t follow = (held >> (8u * lead_suit)) & 0x7Fu;
        return trick_len != 0 && follow != 0 ? follow : occupied(held);
    }

    // validMoves, knockoutwhist.cpp:122-136, as card ids (empty <=> the game is over).  Only
    // meaningful while no chance event is pending.
    ISMCTS_DEV Mask legal() const
    {
        if (request_trump) return Mask{kTrumpChoices};
        return Mask{ids_of(list_of(player), legal_places(held_of(player)))};
    }

    // getResult * 2, knockoutwhist.cpp:117-120: the first remaining player has won.
    ISMCTS_DEV uint32_t result2(uint32_t who) const { return 4u * who == ctz32(alive4) ? 2u : 0u; }
    ISMCTS_DEV uint32_t outcome() const { return ctz32(alive4) >> 2; }

    // ---- chance events

    // The hand the next cards go to: what deal_player holds stays, the first free place is the one
    // after the last card held.  The cards go straight to their places in shared memory (a played
    // card stays on its place, only `held` says whether a place counts), `held` is stored when the
    // hand is complete.
    ISMCTS_DEV void open_hand(Deal& d) const
    {
        d.deal_held = held_of(d.deal_player);
        d.deal_place = top_place(d.deal_held);
    }
    ISMCTS_DEV void close_hand(const Deal& d) const { set_held(d.deal_player, d.deal_held); }
    // (A hand never gets an eighth card in a legal game; one that did would land on the unused eighth byte of
    // the list and set no `held` bit: dropped, without a branch.)
    ISMCTS_DEV void deal_card(Deal& d, uint32_t code) const
    {
        const uint32_t place = d.deal_place < kPlaces ? d.deal_place : kPlaces;
        reinterpret_cast<uint8_t*>(hands() + d.deal_player * kStride)[place] = (uint8_t)code;
        d.deal_held |= (1u << (8u * (code >> 4) + place)) & 0x7F7F7F7Fu;
        ++d.deal_place;
    }

    // Starts dealing from the pile: `sizes` holds 4 bits per player, the number of cards each
    // gets, ascending id.
    ISMCTS_DEV void begin_deal(Deal& d, uint32_t sizes)
    {
        d.need = sizes;
        next_hand(d);
    }

    // The next player to deal to, or the end of the deal.
    ISMCTS_DEV void next_hand(Deal& d)
    {
        if (d.need) {
            const uint32_t shift = ctz32(d.need) & ~3u;
            d.deal_player = shift >> 2;
            d.deal_left = (d.need >> shift) & 15u;
            d.need &= ~(15u << shift);
            mode = kDeal;
            open_hand(d);
        } else {
            mode = kPlay;
        }
    }

    // deal, knockoutwhist.cpp:138-152: a fresh deck, tricks_left cards to each remaining player.
    ISMCTS_DEV void start_deal(Deal& d)
    {
#pragma unroll
        for (uint32_t i = 0; i < kPileWords; ++i) set_pile_word(i, deck_word(i));
        d.pile_n = ISMCTS_KW_CARDS;
        begin_deal(d, alive4 * tricks_left);
    }

    // finishRound, knockoutwhist.cpp:163-180
    ISMCTS_DEV void finish_round()
    {
        alive4 &= nonzero4(tricks);                       // players without a trick are knocked out
        if (alive4 & (alive4 - 1u)) {                     // more than one player left
            --tricks_left;
            request_trump = 1;
            // roundWinner, :198-210: uniform among the players with most tricks
            tied = alive4 & ~nonzero4(tricks ^ (best * kOnes));
            mode = kTie;
        } else {
            tricks_left = 0; // over: nobody holds cards, validMoves() is empty
        }
    }

    // validMoves() is empty: the game is over.  Every transition leaves either a
    // player with cards to move, a pending chance event, or tricks_left == 0
    // (prepare() normalises root records that are finished in some other way).
    ISMCTS_DEV bool terminal() const { return tricks_left == 0; }

    // A pending chance event (mode kDeal or kTie) is one uniform draw below chance_count().
    ISMCTS_DEV uint32_t chance_count(const Deal& d) const { return mode == kDeal ? d.pile_n : popc32(tied); }

    // The last card of the pile takes the place of card `j`, which is dealt.
    ISMCTS_DEV uint32_t take_from_pile(Deal& d, uint32_t j) const
    {
        uint8_t* at = pile_at(j);
        const uint32_t code = *at;
        *at = *pile_at(--d.pile_n);
        return code;
    }

    // The tie-break: the j-th of the tied players has won the round and calls trumps
    // (roundWinner, knockoutwhist.cpp:198-210), then the cards are dealt.
    ISMCTS_DEV void break_tie(Deal& d, uint32_t j)
    {
        player = kth_bit32(tied, j) >> 2;
        dealer = next_alive(dealer);
        tricks = 0; best = 0;
        start_deal(d);
    }

    // The pending chance event with draw `j` (uniform below chance_count()).
    ISMCTS_DEV void chance_draw(Deal& d, uint32_t j)
    {
        if (mode == kDeal) {
            deal_card(d, take_from_pile(d, j));
            if (--d.deal_left == 0) { close_hand(d); next_hand(d); }
        } else {
            break_tie(d, j);
        }
    }

    // ---- moves

    // The card on `place` of the hand (`list`, `held`) of the player to move is played:
    // doMove, knockoutwhist.cpp:95-115.  Leaves the hand of the next player to move in (`list`, `held`).
    // (The state is read before the first store and written after the last load: when the game lives in
    // memory — a tree operation — the compiler must assume that a store to a hand changes it.)
    // kDefer: the end of a round is only recorded (mode = kRoundEnd); the caller runs finish_round() later.
    template <bool kDefer = false>
    ISMCTS_DEV void play_place(uint32_t place, uint64_t& list, uint32_t& held)
    {
        uint64_t* const lists = hands();
        uint32_t* const held_rows = helds();
        const uint32_t pitch = kStride, alive = alive4, trumps = trump;
        uint32_t who = player, len = trick_len, lead = lead_suit, key_won = win_key, who_won = win_player;
        uint32_t taken_all = tricks, most = best;

        const uint32_t code = code_at(list, place);
        const uint32_t suit = code >> 4, rank = code & 15u;
        held_rows[who * pitch] = held & ~(1u << (8u * suit + place));
        // trickWinner, :182-196, as in fold_trick()
        if (len == 0) lead = suit;
        const uint32_t key = rank + (suit == trumps ? 32u : suit == lead ? 16u : 0u);
        if (len == 0 || key > key_won) { key_won = key; who_won = who; }
        ++len;
        {   // nextPlayer, :83-88
            const uint32_t above = alive & ~((2u << (4 * who)) - 1u);
            who = ctz32(above ? above : alive) >> 2;
        }
        if (len == popc32(alive)) {                       // finishTrick, :155-161
            taken_all += 1u << (4 * who_won);
            const uint32_t taken = (taken_all >> (4 * who_won)) & 15u;
            most = taken > most ? taken : most;
            len = 0;
            who = who_won;
        }
        list = lists[who * pitch];
        held = held_rows[who * pitch];
        player = who; trick_len = len; lead_suit = lead; win_key = key_won; win_player = who_won;
        tricks = taken_all; best = most;
        if (held == 0) { if constexpr (kDefer) mode = kRoundEnd; else finish_round(); }
    }

    // doMove for a card id (a move of the tree or of a caller; no chance event pending).
    ISMCTS_DEV void play(uint32_t id)
    {
        if (request_trump) {
            trump = id / 13u;
            request_trump = 0;
            player = next_alive(dealer);
            return;
        }
        uint64_t list = list_of(player);
        uint32_t held = held_of(player);
        const uint32_t code = code_of_id(id);
        uint32_t place = 0;
        for (uint32_t rest = (held >> (8u * (code >> 4))) & 0x7Fu; rest; rest &= rest - 1) {
            place = ctz32(rest);
            if (code_at(list, place) == code) break;
        }
        play_place(place, list, held);
    }
    ISMCTS_DEV void step(uint32_t id) { play(id); }

    // ---- phases of the lock-step loop (lane_search.cuh): the same transitions as chance_draw() and
    // play() with the same draws, as tight loops over ONE kind of step.  Every lane that takes
    // part makes one draw per step from an aligned stream, so draw s of the phase is third s % 3 of
    // word (s / 3) % 4 of block s / 12 for all of them: the block is generated in registers by all
    // lanes in the same steps and no ring is involved.
    static constexpr bool kPhased = true;
    static constexpr uint32_t kDrawsPerWord = 3;      // choices are among at most 52 (philox.cuh)

    // A block of the stream in registers.  A turn of a phase loop is three steps (the three draws
    // of one word); every fourth turn generates the next block.
    struct BlockWords {
        uint32_t w, b1, b2, b3;   // the word being used up, the words not started yet
        template <class Rng>
        ISMCTS_DEV void turn(const Rng& rng, uint32_t first_block, uint32_t t)
        {
            if ((t & 3u) == 0u) {
                uint32_t b[4];
                rng.block(first_block + (t >> 2), b);
                w = b[0]; b1 = b[1]; b2 = b[2]; b3 = b[3];
            } else {
                w = b1; b1 = b2; b2 = b3;
            }
        }
        // uniform integer below n from the word in use
        ISMCTS_DEV uint32_t below(uint32_t n)
        {
            const uint64_t product = (uint64_t)w * n;
            w = (uint32_t)product;
            return (uint32_t)(product >> 32);
        }
    };

    // Every pending chance event of the warp: the tie-break (it starts a deal), then the deal
    // card by card.  `active`: the lane takes part in the search (it may have nothing pending).
    // `prep_of()`: the lane's position (asked for when mode == kStart: the determinisation begins here).
    template <class Rng, class PrepOf>
    ISMCTS_DEV void chance_phase(Rng& rng, bool active, PrepOf prep_of)
    {
        const bool mine = active && mode != kPlay;
        if (!warp_any(mine)) return;
        Deal deal;
        deal.pile_n = 0; deal.deal_player = 0; deal.deal_left = 0; deal.need = 0; deal.deal_place = 0; deal.deal_held = 0;
        if (mine && mode == kStart) begin_start(deal, prep_of());
        const uint32_t first_block = mine ? rng.phase_begin() : 0u;
        BlockWords words;
        words.turn(rng, first_block, 0);
        // A tie-break is the first draw of its lane (the first third of word 0) and is made here, once, so that
        // the loop below is the deal step alone (three copies of it: the shorter they are the better they are
        // fetched).  `used`: thirds of the current word the lane has used before the loop.
        uint32_t draws = 0, used = 0;
        if (mine && mode == kTie) { draws = 1; used = 1; break_tie(deal, words.below(popc32(tied))); }
        for (uint32_t t = 0; warp_any(active && mode != kPlay); ++t) {
            if (t != 0) words.turn(rng, first_block, t);
#pragma unroll
            for (uint32_t part = 0; part < kDrawsPerWord; ++part) {
                if (active && mode == kDeal && (part != 0 || used == 0)) {
                    ++draws;
                    deal_card(deal, take_from_pile(deal, words.below(deal.pile_n)));
                    if (--deal.deal_left == 0) { close_hand(deal); next_hand(deal); }
                }
            }
            used = 0;
        }
        if (mine) { rng.phase_end(first_block, (draws + kDrawsPerWord - 1u) / kDrawsPerWord); rng.align(); }
    }

    struct PlaceTableRef {
#if defined(__CUDA_ARCH__)
        uint32_t address;
        ISMCTS_DEV explicit PlaceTableRef(const uint32_t* t) : address((uint32_t)__cvta_generic_to_shared(t)) { asm volatile("" : "+r"(address)); }
        ISMCTS_DEV uint32_t at(uint32_t i) const
        {
            uint32_t v;
            asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(address + 4u * i));
            return v;
        }
#else
        const uint32_t* table;
        ISMCTS_DEV explicit PlaceTableRef(const uint32_t* t) : table(t) {}
        ISMCTS_DEV uint32_t at(uint32_t i) const { return table[i]; }
#endif
    };

    // Rollout moves (simulate, solverbase.h:36-43: uniform over validMoves(), knockoutwhist.cpp:122-136) for the
    // lanes with `rolling` until their round ends (a chance event becomes pending) or their game does.
    // `plies` counts the moves.
    template <class Rng>
    ISMCTS_DEV void play_phase(Rng& rng, bool rolling, uint32_t& plies)
    {
        const bool mine = rolling && mode == kPlay && tricks_left != 0;
        if (!warp_any(mine)) return;
        const uint32_t first_block = mine ? rng.phase_begin() : 0u;   // the rollout and every round start a block
        // (the table is addressed through a 32-bit shared-memory address the compiler cannot recompute: ptxas
        // would otherwise derive the address of the block's shared memory again in every step, four instructions)
        const PlaceTableRef table(place_table());
        BlockWords words;
        words.turn(rng, first_block, 0);
        // The round winner calls trumps: one of the four suits.  Only the first move of a phase can be that
        // (the request is made by the end of a round, which ends the phase), so it is made here, once, and the
        // loop below is the card step alone; the end of a round is recorded there and run after the loop.
        uint32_t draws = 0, used = 0;
        if (mine && request_trump) { draws = 1; used = 1; play(13u * words.below(4)); }
        // the hand of the player to move (lanes that do not take part may hold no game at all)
        uint64_t list = mine ? list_of(player) : 0;
        uint32_t held = mine ? held_of(player) : 0;
        for (uint32_t t = 0; warp_any(rolling && mode == kPlay && tricks_left != 0); ++t) {
            if (t != 0) words.turn(rng, first_block, t);
#pragma unroll
            for (uint32_t part = 0; part < kDrawsPerWord; ++part) {
                if (rolling && mode == kPlay && tricks_left != 0 && (part != 0 || used == 0)) {
                    ++draws;
                    const uint32_t places = legal_places(held);
                    const uint32_t j = words.below(popc32(places));
                    play_place<true>((table.at(places) >> (3u * j)) & 7u, list, held);
                }
            }
            used = 0;
        }
        if (mode == kRoundEnd) { mode = kPlay; finish_round(); }
        plies += draws;
        if (mine) rng.phase_end(first_block, (draws + kDrawsPerWord - 1u) / kDrawsPerWord);
    }
};

static_assert(std::is_trivially_copyable<KwGameT<4>>::value && sizeof(KwGameT<4>) == 56 && sizeof(KwGameT<ISMCTS_KW_MAX_PLAYERS>) == 56, "home(): a game fits the 56 bytes of its pile");

using KwGame = KwGameT<ISMCTS_KW_MAX_PLAYERS>;
using KwGame4 = KwGameT<4>;   // positions with at most four players (ISMCTS_FLAG_KW_MAX4)

} // namespace ismcts
