/*
 * knockoutwhist.h — Knockout Whist as a host game object for the solvers.
 *
 * Public behaviour of the reference's KnockoutWhist (test/common/knockoutwhist.{h,cpp}):
 * 2-7 players, rounds of 7..1 tricks, follow suit, trumps, players without a trick are
 * knocked out, the round winner (ties broken at random) calls trumps.  The state is the
 * plain-data record the kernels search (ismcts_kw_state) and every rule is evaluated by the
 * engine's host helpers (ismcts_kw_*), i.e. by the same code the kernels run.
 */
#ifndef ISMCTS_GAMES_KNOCKOUTWHIST_H
#define ISMCTS_GAMES_KNOCKOUTWHIST_H

#include "../game.h"
#include "card.h"
#include "../../ismcts_b200.h"

#include <algorithm>
#include <cstring>
#include <random>
#include <stdexcept>

class KnockoutWhist : public ISMCTS::POMGame<Card>, public ISMCTS::PodGame
{
public:
    /// A new game with a random deal (seeded from std::random_device unless a seed is given).
    explicit KnockoutWhist(unsigned players = 4) : KnockoutWhist{players, std::random_device{}()} {}

    KnockoutWhist(unsigned players, std::uint64_t seed) : m_seed{seed}
    {
        if (ismcts_kw_deal(m_seed, 0, players, &m_state) != ISMCTS_OK)
            throw std::runtime_error("KnockoutWhist: deal failed");
    }

    /// A game put into a given position (test fixtures, position streams).
    explicit KnockoutWhist(ismcts_kw_state const &state, std::uint64_t seed = 1) : m_state(state), m_seed{seed} {}

    Clone cloneAndRandomise(Player observer) const override
    {
        // knockoutwhist.cpp:56-76: everything the observer cannot see is pooled and dealt back
        auto clone = std::make_unique<KnockoutWhist>(*this);
        ismcts_kw_state &s = clone->m_state;
        std::vector<int> pool;
        for (int c = 0; c < ISMCTS_KW_CARDS; ++c) {
            bool hidden = (s.unknown >> c) & 1;
            for (unsigned p = 0; p < s.num_players; ++p)
                if (p != observer && ((s.alive_mask >> p) & 1) && ((s.hands[p] >> c) & 1))
                    hidden = true;
            if (hidden)
                pool.push_back(c);
        }
        std::mt19937_64 rng{m_seed ^ (0x9E3779B97F4A7C15ull * ++m_clones)};
        std::shuffle(pool.begin(), pool.end(), rng);
        std::size_t next = 0;
        for (unsigned p = 0; p < s.num_players; ++p) {
            if (p == observer || !((s.alive_mask >> p) & 1))
                continue;
            int const size = __builtin_popcountll(s.hands[p]);
            s.hands[p] = 0;
            for (int i = 0; i < size; ++i)
                s.hands[p] |= std::uint64_t(1) << pool[next++];
        }
        return clone;
    }

    Player currentPlayer() const override { return m_state.player; }

    std::vector<Player> players() const override
    {
        std::vector<Player> out;
        for (unsigned p = 0; p < m_state.num_players; ++p)
            if ((m_state.alive_mask >> p) & 1)
                out.push_back(p);
        return out;
    }

    std::vector<Card> validMoves() const override
    {
        std::vector<Card> out;
        std::uint64_t const legal = ismcts_kw_legal(&m_state);
        for (int c = 0; c < ISMCTS_KW_CARDS; ++c)
            if ((legal >> c) & 1)
                out.push_back(Card::fromId(c));
        return out;
    }

    void doMove(Card const move) override
    {
        if (ismcts_kw_apply(&m_state, static_cast<std::uint32_t>(int(move)), m_seed, 1, &m_draws) != ISMCTS_OK)
            throw std::out_of_range("Invalid move");
    }

    double getResult(Player player) const override { return 0.5 * ismcts_kw_result2(&m_state, player); }

    // PodGame
    std::uint32_t podKind() const override { return ISMCTS_GAME_KW; }
    std::size_t podSize() const override { return sizeof m_state; }
    void exportPod(void *destination) const override { std::memcpy(destination, &m_state, sizeof m_state); }

    ismcts_kw_state const &state() const { return m_state; }

protected:
    ismcts_kw_state m_state;
    std::uint64_t m_seed;
    std::uint32_t m_draws {0};
    mutable std::uint64_t m_clones {0};
};

#endif // ISMCTS_GAMES_KNOCKOUTWHIST_H
