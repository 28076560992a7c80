// games_host.cu — host game helpers of the C ABI (include/ismcts_b200.h).
//
// The reference's games are host objects that callers construct and advance
// between searches (KnockoutWhist / MnkGame in test/common, driven by
// test/benchmark.cpp:73-89).  These helpers do the same on the POD position
// records, with the SAME rule code the kernels run (kw_game.cuh, mnk_game.cuh
// are __host__ __device__).  They contain no search: the search has no host path.
#include "philox.cuh"
#include "kw_game.cuh"
#include "mnk_game.cuh"
#include "goof_game.cuh"
#include "phantom_game.cuh"

#include <cstring>

using namespace ismcts;

namespace {

// Streams of the position generators (never used by a search: tree ids of a
// search are < 2^31).
constexpr uint32_t kDealTree = 0xFFFFFFFFu, kDealIteration = 0xFFFFFFFFu;
constexpr uint32_t kApplyTree = 0xFFFFFFFEu;

// The scratch of ONE host-side Knockout Whist game (the game code finds it through devdefs.cuh: block_scratch).
struct KwHostScratch {
    uint64_t words[(size_t)KwGame::kHandWords * kBlockLanes];
    KwHostScratch() { std::memset(words, 0, sizeof words); g_host_block_scratch = words; g_host_block_lane = 0; }
    ~KwHostScratch() { g_host_block_scratch = nullptr; }
    KwHostScratch(const KwHostScratch&) = delete;
    KwHostScratch& operator=(const KwHostScratch&) = delete;
};

void kw_store(const KwGame& g, uint64_t unknown, ismcts_kw_state* s)
{
    for (uint32_t p = 0; p < ISMCTS_KW_MAX_PLAYERS; ++p) s->hands[p] = g.hand_ids(p);
    s->unknown = unknown;
    s->player = (uint8_t)g.player; s->dealer = (uint8_t)g.dealer; s->trump = (uint8_t)g.trump;
    s->tricks_left = (uint8_t)g.tricks_left; s->request_trump = (uint8_t)g.request_trump;
    uint32_t alive = 0;
    for (uint32_t p = 0; p < ISMCTS_KW_MAX_PLAYERS; ++p) alive |= ((g.alive4 >> (4 * p)) & 1u) << p;
    s->alive_mask = (uint8_t)alive;
    s->trick_len = (uint8_t)g.trick_len;
    for (uint32_t p = 0; p < 8; ++p) s->tricks_taken[p] = p < 7 ? (uint8_t)((g.tricks >> (4 * p)) & 15u) : 0;
}

} // namespace

extern "C" {

int ismcts_kw_deal(uint64_t seed, uint32_t stream, uint32_t num_players, ismcts_kw_state* out)
{
    if (!out) return ISMCTS_ERR_INVALID;
    // KnockoutWhist::KnockoutWhist, knockoutwhist.cpp:36-54 (player count clamped to 2..7, :38)
    const uint32_t np = num_players < 2 ? 2 : num_players > 7 ? 7 : num_players;
    std::memset(out, 0, sizeof *out);
    out->num_players = (uint8_t)np;
    KwHostScratch scratch;
    KwGame g;
    ismcts_kw_state blank;
    std::memset(&blank, 0, sizeof blank);
    blank.num_players = (uint8_t)np; blank.alive_mask = (uint8_t)((1u << np) - 1u);
    blank.tricks_left = 7; blank.dealer = (uint8_t)(np - 1);
    g.load(&blank);
    uint32_t ring[LaneRng::kRingWords];
    LaneRng rng;
    rng.init(seed, stream, kDealTree, ring, 1);
    rng.per_word = KwGame::kDrawsPerWord;
    rng.begin_iteration(kDealIteration);
    KwGame::Deal deal;
    g.start_deal(deal);
    while (g.chance_pending()) g.chance_draw(deal, rng.below(g.chance_count(deal)));
    // one of the remaining cards is turned up for trumps and leaves the unknown set (:50-53)
    uint64_t unknown = g.pile_set(deal);
    rng.align();
    const uint32_t up = kth_bit64(unknown, rng.below(popc64(unknown)));
    unknown &= ~((uint64_t)1 << up);
    g.trump = up / 13u;
    kw_store(g, unknown, out);
    return ISMCTS_OK;
}

uint64_t ismcts_kw_legal(const ismcts_kw_state* s)
{
    if (!s) return 0;
    KwHostScratch scratch;
    KwGame g;
    g.load(s);
    return g.legal().w;
}

int ismcts_kw_apply(ismcts_kw_state* s, uint32_t move, uint64_t seed, uint32_t stream, uint32_t* draw_counter)
{
    if (!s) return ISMCTS_ERR_INVALID;
    if (move >= ISMCTS_KW_CARDS) return ISMCTS_ERR_ILLEGAL_MOVE;
    KwHostScratch scratch;
    KwGame g;
    g.load(s);
    // doMove accepts any card while trumps are requested (knockoutwhist.cpp:97-102); otherwise
    // the card must be in the mover's hand or std::out_of_range is thrown (:105-109)
    const bool play = !g.request_trump;
    if (play && !((g.hand_ids(g.player) >> move) & 1u)) return ISMCTS_ERR_ILLEGAL_MOVE;
    const uint32_t mover = g.player;
    uint32_t ring[LaneRng::kRingWords];
    LaneRng rng;
    rng.init(seed, stream, kApplyTree, ring, 1);
    rng.per_word = KwGame::kDrawsPerWord;
    rng.begin_iteration(0, draw_counter ? *draw_counter : 0);
    g.step(move);
    uint64_t unknown = s->unknown;
    if (g.chance_pending()) {
        // the chance events of a round end (tie-break, deal) start a Philox block, and so does what follows
        rng.align();
        KwGame::Deal deal{};
        while (g.chance_pending()) g.chance_draw(deal, rng.below(g.chance_count(deal)));
        rng.align();
        unknown = g.pile_set(deal); // deal, :138-152: the undealt rest of the deck is unknown to everyone
    }
    if (draw_counter) *draw_counter = rng.consumed();
    // the cards of the trick in progress (m_currentTrick)
    if (play) {
        if (g.trick_len == 0) {
            std::memset(s->trick_player, 0, sizeof s->trick_player);
            std::memset(s->trick_card, 0, sizeof s->trick_card);
        } else {
            s->trick_player[g.trick_len - 1] = (uint8_t)mover;
            s->trick_card[g.trick_len - 1] = (uint8_t)move;
        }
    }
    kw_store(g, unknown, s);
    return ISMCTS_OK;
}

int ismcts_kw_result2(const ismcts_kw_state* s, uint32_t player)
{
    if (!s || s->alive_mask == 0) return 0;
    return player == ctz32(s->alive_mask) ? 2 : 0; // getResult, knockoutwhist.cpp:117-120
}

// ---- phantom m,n,k (test/common/phantommnkgame.{h,cpp})

int ismcts_phantom_init(uint32_t m, uint32_t n, uint32_t k, ismcts_phantom_state* out)
{
    if (!out || m == 0 || n == 0 || m * n > ISMCTS_MNK_MAX_CELLS || m > 255 || n > 255 || k > 255) return ISMCTS_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    out->m = (uint8_t)m; out->n = (uint8_t)n; out->k = (uint8_t)k; out->result2 = -1;
    for (uint32_t c = 0; c < m * n; ++c)
        for (int p = 0; p < 2; ++p) out->avail[p][c >> 6] |= (uint64_t)1 << (c & 63);
    return ISMCTS_OK;
}

void ismcts_phantom_legal(const ismcts_phantom_state* s, uint64_t legal[4])
{
    for (int i = 0; i < 4; ++i) legal[i] = s ? s->avail[s->player & 1][i] : 0;
}

int ismcts_phantom_apply(ismcts_phantom_state* s, uint32_t move)
{
    if (!s) return ISMCTS_ERR_INVALID;
    uint64_t scratch[PhantomGame::kHandWords];
    PhantomGame g;
    g.bind(scratch, 1, 0);
    g.load(s);
    // phantommnkgame.cpp:85-88: the move must be one the player still believes possible
    if (move >= (uint32_t)s->m * s->n || !g.legal().test(move)) return ISMCTS_ERR_ILLEGAL_MOVE;
    g.play(move);
    for (uint32_t i = 0; i < 8; ++i) { s->stones[i >> 2][i & 3] = scratch[i]; s->avail[i >> 2][i & 3] = scratch[8 + i]; }
    s->player = (uint8_t)g.player; s->result2 = (int8_t)g.res2;
    return ISMCTS_OK;
}

int ismcts_phantom_result2(const ismcts_phantom_state* s, uint32_t player)
{
    if (!s) return 0;
    return player == 0 ? s->result2 : 2 - s->result2;
}

// ---- Goofspiel (test/common/goofspiel.{h,cpp})

namespace {
void goof_store(const GoofGame& g, uint32_t n, ismcts_goof_state* s)
{
    std::memset(s, 0, sizeof *s);
    s->hands[0] = (uint16_t)(g.hands & 0x1FFFu); s->hands[1] = (uint16_t)((g.hands >> 16) & 0x1FFFu);
    s->n_prizes = (uint8_t)n;
    s->prize_order = n < 16 ? g.order & ((((uint64_t)1 << (4 * n)) - 1u)) : g.order;
    s->player = (uint8_t)g.player; s->draw_prize = (uint8_t)g.draw_prize; s->current_prize = (uint8_t)g.current_prize;
    s->moves[0] = (uint8_t)g.move0; s->moves[1] = (uint8_t)g.move1;
    s->scores[0] = (uint8_t)g.score0; s->scores[1] = (uint8_t)g.score1;
}
} // namespace

int ismcts_goof_init(uint64_t seed, uint32_t stream, ismcts_goof_state* out)
{
    if (!out) return ISMCTS_ERR_INVALID;
    // Goofspiel::Goofspiel, goofspiel.cpp:52-63: entry i of the shuffled prizes is a uniformly random prize not placed yet
    std::memset(out, 0, sizeof *out);
    uint32_t ring[LaneRng::kRingWords];
    LaneRng rng;
    rng.init(seed, stream, kDealTree, ring, 1);
    rng.begin_iteration(kDealIteration);
    uint32_t left = 0x1FFFu;
    for (uint32_t i = 0; i < 13; ++i) {
        const uint32_t r = kth_bit32(left, rng.below(popc32(left)));
        left &= ~(1u << r);
        out->prize_order |= (uint64_t)r << (4 * i);
        if (rng.low()) rng.refill();
    }
    out->hands[0] = out->hands[1] = 0x1FFF;
    out->n_prizes = 13; out->player = 2; out->draw_prize = 1;
    return ISMCTS_OK;
}

uint64_t ismcts_goof_legal(const ismcts_goof_state* s)
{
    if (!s) return 0;
    GoofGame g;
    g.load(s);
    // the position as its own player to move sees it: the environment knows which prize comes next
    if (g.draw_pending()) g.chance_step((uint32_t)((g.order >> (4 * (g.n_prizes() - 1u))) & 15u));
    return g.legal().w;
}

int ismcts_goof_apply(ismcts_goof_state* s, uint32_t move)
{
    if (!s) return ISMCTS_ERR_INVALID;
    GoofGame g;
    g.load(s);
    uint32_t n = s->n_prizes;
    if (g.draw_pending()) { g.chance_step((uint32_t)((g.order >> (4 * (n - 1u))) & 15u)); --n; }
    if (move >= 64 || !((g.legal().w >> move) & 1u)) return ISMCTS_ERR_ILLEGAL_MOVE;
    g.play(move);
    goof_store(g, n, s);
    return ISMCTS_OK;
}

int ismcts_goof_result2(const ismcts_goof_state* s, uint32_t player)
{
    if (!s) return 0;
    GoofGame g;
    g.load(s);
    return (int)g.result2(player);
}

int ismcts_mnk_init(uint32_t m, uint32_t n, uint32_t k, ismcts_mnk_state* out)
{
    if (!out || m == 0 || n == 0 || m * n > ISMCTS_MNK_MAX_CELLS || m > 255 || n > 255 || k > 255) return ISMCTS_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    out->m = (uint8_t)m; out->n = (uint8_t)n; out->k = (uint8_t)k; out->player = 0; out->result2 = -1;
    return ISMCTS_OK;
}

void ismcts_mnk_legal(const ismcts_mnk_state* s, uint64_t legal[4])
{
    uint64_t stones[8];
    MnkGame g;
    g.bind(stones, 1, 0);
    g.load(s);
    const MnkGame::Mask l = g.legal();
    for (int i = 0; i < 4; ++i) legal[i] = l.w[i];
}

int ismcts_mnk_apply(ismcts_mnk_state* s, uint32_t move)
{
    if (!s) return ISMCTS_ERR_INVALID;
    uint64_t stones[8];
    MnkGame g;
    g.bind(stones, 1, 0);
    g.load(s);
    // mnkgame.cpp:41-43: the move must be one of the remaining cells
    if (move >= (uint32_t)s->m * s->n || g.occupied_by(0, move) || g.occupied_by(1, move)) return ISMCTS_ERR_ILLEGAL_MOVE;
    g.step(move);
    for (uint32_t i = 0; i < 8; ++i) s->stones[i >> 2][i & 3] = stones[i];
    s->player = (uint8_t)g.player; s->result2 = (int8_t)g.res2;
    return ISMCTS_OK;
}

int ismcts_mnk_result2(const ismcts_mnk_state* s, uint32_t player)
{
    if (!s) return 0;
    return player == 0 ? s->result2 : 2 - s->result2; // getResult, mnkgame.cpp:56-59
}

} // extern "C"
