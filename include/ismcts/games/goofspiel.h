/*
 * goofspiel.h — two-player Goofspiel as a host game object.
 *
 * Public behaviour of the reference's Goofspiel (test/common/goofspiel.{h,cpp}): hearts are the prizes,
 * player 0 bids spades, player 1 clubs, player 2 is the environment that turns over the prizes and
 * reveals the bids; all moves are simultaneous, so the solvers use their simultaneous-move policy (EXP3).
 * State and rules are the engine's (ismcts_goof_state, ismcts_goof_*).
 */
#ifndef ISMCTS_GAMES_GOOFSPIEL_H
#define ISMCTS_GAMES_GOOFSPIEL_H

#include "../game.h"
#include "card.h"
#include "../../ismcts_b200.h"

#include <algorithm>
#include <cstring>
#include <random>
#include <stdexcept>

class Goofspiel : public ISMCTS::POMGame<Card>, public ISMCTS::PodGame
{
public:
    Goofspiel() : Goofspiel{std::random_device{}()} {}

    explicit Goofspiel(std::uint64_t seed) : m_seed{seed}
    {
        if (ismcts_goof_init(seed, 0, &m_state) != ISMCTS_OK)
            throw std::runtime_error("Goofspiel: initialisation failed");
    }

    explicit Goofspiel(ismcts_goof_state const &state, std::uint64_t seed = 1) : m_state(state), m_seed{seed} {}

    Clone cloneAndRandomise(Player observer) const override
    {
        // goofspiel.cpp:65-81: the players do not know the order of the remaining prizes; player 1 does
        // not know the bid player 0 has just made
        auto clone = std::make_unique<Goofspiel>(*this);
        std::mt19937_64 rng{m_seed ^ (0x9E3779B97F4A7C15ull * ++m_clones)};
        ismcts_goof_state &s = clone->m_state;
        if (observer != 2) {
            std::vector<unsigned> order;
            for (unsigned i = 0; i < s.n_prizes; ++i)
                order.push_back((s.prize_order >> (4 * i)) & 15);
            std::shuffle(order.begin(), order.end(), rng);
            s.prize_order = 0;
            for (unsigned i = 0; i < order.size(); ++i)
                s.prize_order |= std::uint64_t(order[i]) << (4 * i);
        }
        if (observer == 1) {
            std::vector<unsigned> hand;
            for (unsigned r = 0; r < 13; ++r)
                if ((s.hands[0] >> r) & 1)
                    hand.push_back(r);
            s.moves[0] = hand[rng() % hand.size()];
        }
        return clone;
    }

    Player currentPlayer() const override { return m_state.player; }
    std::vector<Player> players() const override { return {0, 1, 2}; }
    bool currentMoveSimultaneous() const override { return true; }

    std::vector<Card> validMoves() const override
    {
        std::vector<Card> out;
        std::uint64_t const legal = ismcts_goof_legal(&m_state);
        for (int c = 0; c < 64; ++c)
            if ((legal >> c) & 1)
                out.push_back(Card::fromId(c));
        return out;
    }

    void doMove(Card const move) override
    {
        if (ismcts_goof_apply(&m_state, static_cast<std::uint32_t>(int(move))) != ISMCTS_OK)
            throw std::out_of_range("Invalid move");
    }

    double getResult(Player player) const override { return 0.5 * ismcts_goof_result2(&m_state, player); }

    std::uint32_t podKind() const override { return ISMCTS_GAME_GOOF; }
    std::size_t podSize() const override { return sizeof m_state; }
    void exportPod(void *destination) const override { std::memcpy(destination, &m_state, sizeof m_state); }

    ismcts_goof_state const &state() const { return m_state; }

protected:
    ismcts_goof_state m_state;
    std::uint64_t m_seed;
    mutable std::uint64_t m_clones {0};
};

#endif // ISMCTS_GAMES_GOOFSPIEL_H
