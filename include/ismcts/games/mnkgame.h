/*
 * mnkgame.h — the m,n,k-game (generalised tic-tac-toe) as a host game object.
 *
 * Public behaviour of the reference's MnkGame (test/common/mnkgame.{h,cpp}): n rows by m
 * columns, move = cell index, a win is a run of exactly k through the last stone, results
 * 0 / 0.5 / 1.  State and rules are the engine's (ismcts_mnk_state, ismcts_mnk_*).
 */
#ifndef ISMCTS_GAMES_MNKGAME_H
#define ISMCTS_GAMES_MNKGAME_H

#include "../game.h"
#include "../../ismcts_b200.h"

#include <cstring>
#include <stdexcept>

class MnkGame : public ISMCTS::POMGame<int>, public ISMCTS::PodGame
{
public:
    explicit MnkGame(int m = 3, int n = 3, int k = 3)
    {
        if (ismcts_mnk_init(m, n, k, &m_state) != ISMCTS_OK)
            throw std::invalid_argument("MnkGame: board must have 1..256 cells");
    }

    explicit MnkGame(ismcts_mnk_state const &state) : m_state(state) {}

    Clone cloneAndRandomise(Player) const override { return std::make_unique<MnkGame>(*this); }
    Player currentPlayer() const override { return m_state.player; }
    std::vector<Player> players() const override { return {0, 1}; }

    std::vector<int> validMoves() const override
    {
        std::uint64_t legal[4];
        ismcts_mnk_legal(&m_state, legal);
        std::vector<int> out;
        for (int c = 0; c < m_state.m * m_state.n; ++c)
            if ((legal[c >> 6] >> (c & 63)) & 1)
                out.push_back(c);
        return out;
    }

    void doMove(int const move) override
    {
        if (move < 0 || ismcts_mnk_apply(&m_state, static_cast<std::uint32_t>(move)) != ISMCTS_OK)
            throw std::out_of_range("Illegal move");
    }

    double getResult(Player player) const override { return 0.5 * ismcts_mnk_result2(&m_state, player); }

    std::uint32_t podKind() const override { return ISMCTS_GAME_MNK; }
    std::size_t podSize() const override { return sizeof m_state; }
    void exportPod(void *destination) const override { std::memcpy(destination, &m_state, sizeof m_state); }

    ismcts_mnk_state const &state() const { return m_state; }

protected:
    ismcts_mnk_state m_state;
};

#endif // ISMCTS_GAMES_MNKGAME_H
