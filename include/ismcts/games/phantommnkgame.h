/*
 * phantommnkgame.h — the phantom m,n,k-game as a host game object.
 *
 * Public behaviour of the reference's PhantomMnkGame (test/common/phantommnkgame.{h,cpp}): the m,n,k-game in
 * which the players do not see each other's stones; trying an occupied cell strikes it off the player's
 * list and the same player moves again.  State and rules are the engine's (ismcts_phantom_state,
 * ismcts_phantom_*).
 */
#ifndef ISMCTS_GAMES_PHANTOMMNKGAME_H
#define ISMCTS_GAMES_PHANTOMMNKGAME_H

#include "../game.h"
#include "../../ismcts_b200.h"

#include <algorithm>
#include <cstring>
#include <random>
#include <stdexcept>

class PhantomMnkGame : public ISMCTS::POMGame<int>, public ISMCTS::PodGame
{
public:
    explicit PhantomMnkGame(int m = 3, int n = 3, int k = 3, std::uint64_t seed = std::random_device{}()) : m_seed{seed}
    {
        if (ismcts_phantom_init(m, n, k, &m_state) != ISMCTS_OK)
            throw std::invalid_argument("PhantomMnkGame: board must have 1..256 cells");
    }

    explicit PhantomMnkGame(ismcts_phantom_state const &state, std::uint64_t seed = 1) : m_state(state), m_seed{seed} {}

    /// phantommnkgame.cpp:21-81: the opponent's stones the observer has not run into are taken back and as
    /// many are put on random cells the observer cannot rule out, without completing a line.
    Clone cloneAndRandomise(Player observer) const override
    {
        auto clone = std::make_unique<PhantomMnkGame>(*this);
        ismcts_phantom_state &s = clone->m_state;
        unsigned const opp = 1 - observer;
        std::vector<int> candidates;
        unsigned hidden = 0;
        for (int c = 0; c < s.m * s.n; ++c) {
            std::uint64_t const bit = std::uint64_t(1) << (c & 63);
            if ((s.stones[opp][c >> 6] & bit) && (s.avail[observer][c >> 6] & bit)) {
                s.stones[opp][c >> 6] &= ~bit;
                s.avail[opp][c >> 6] |= bit;
                ++hidden;
            }
            if (!((s.stones[0][c >> 6] | s.stones[1][c >> 6]) & bit))
                candidates.push_back(c);
        }
        std::mt19937_64 rng{m_seed ^ (0x9E3779B97F4A7C15ull * ++m_clones)};
        ismcts_phantom_state const base = s;
        for (;;) {
            s = base;
            s.player = opp;
            std::shuffle(candidates.begin(), candidates.end(), rng);
            bool ok = true;
            for (unsigned i = 0; i < hidden && ok; ++i) {
                // place through the rules so that a completed line is noticed
                ismcts_phantom_state trial = s;
                trial.player = opp;
                trial.avail[opp][candidates[i] >> 6] |= std::uint64_t(1) << (candidates[i] & 63);
                ismcts_phantom_apply(&trial, candidates[i]);
                ok = trial.result2 < 0 || trial.result2 == 1;
                s.stones[opp][candidates[i] >> 6] |= std::uint64_t(1) << (candidates[i] & 63);
                s.avail[opp][candidates[i] >> 6] &= ~(std::uint64_t(1) << (candidates[i] & 63));
            }
            if (ok)
                break;
        }
        s.player = m_state.player;
        s.result2 = m_state.result2;
        return clone;
    }

    Player currentPlayer() const override { return m_state.player; }
    std::vector<Player> players() const override { return {0, 1}; }

    std::vector<int> validMoves() const override
    {
        std::uint64_t legal[4];
        ismcts_phantom_legal(&m_state, legal);
        std::vector<int> out;
        for (int c = 0; c < m_state.m * m_state.n; ++c)
            if ((legal[c >> 6] >> (c & 63)) & 1)
                out.push_back(c);
        return out;
    }

    void doMove(int const move) override
    {
        if (move < 0 || ismcts_phantom_apply(&m_state, static_cast<std::uint32_t>(move)) != ISMCTS_OK)
            throw std::out_of_range("Illegal move");
    }

    double getResult(Player player) const override { return 0.5 * ismcts_phantom_result2(&m_state, player); }

    std::uint32_t podKind() const override { return ISMCTS_GAME_PHANTOM; }
    std::size_t podSize() const override { return sizeof m_state; }
    void exportPod(void *destination) const override { std::memcpy(destination, &m_state, sizeof m_state); }

    ismcts_phantom_state const &state() const { return m_state; }

protected:
    ismcts_phantom_state m_state;
    std::uint64_t m_seed;
    mutable std::uint64_t m_clones {0};
};

#endif // ISMCTS_GAMES_PHANTOMMNKGAME_H
