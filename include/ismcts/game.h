/*
 * game.h — the game concept of the solvers.
 *
 * Same interface as the reference's ISMCTS::Game<Move> / POMGame<Move>
 * (include/ismcts/game.h:15-46) so that existing game classes and call sites
 * compile unchanged.  A game that can be searched by the CUDA execution policies
 * additionally implements PodGame: the kernels cannot call virtual host functions,
 * they work on the plain-data position records of include/ismcts_b200.h.
 */
#ifndef ISMCTS_GAME_H
#define ISMCTS_GAME_H

#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace ISMCTS
{

template<class Move>
struct Game
{
    using Clone = std::unique_ptr<Game>;
    using Player = unsigned int;
    using MoveType = Move;

    virtual ~Game() = default;

    /// A copy of the game in which everything `observer` cannot see has been randomised.
    virtual Clone cloneAndRandomise(Player observer) const = 0;
    /// The player who makes the next move.
    virtual Player currentPlayer() const = 0;
    /// The moves available to that player; empty when the game is over.
    virtual std::vector<Move> validMoves() const = 0;
    /// Applies a move; throws std::out_of_range when it cannot be made.
    virtual void doMove(Move const move) = 0;
    /// Result for `player` in [0, 1] once the game is over.
    virtual double getResult(Player player) const = 0;
    /// Whether the next move is made simultaneously by several players.
    virtual bool currentMoveSimultaneous() const { return false; }
};

/// Game with partially observable moves: the multiple-observer solver needs the player list.
template<class Move>
struct POMGame : public Game<Move>
{
    using typename Game<Move>::Player;
    virtual std::vector<Player> players() const = 0;
};

template<class G>
using MoveType = typename G::MoveType;

/*
 * Device form of a game: which kernel family plays it (ISMCTS_GAME_*) and its
 * current position as the plain-data record of include/ismcts_b200.h, plus the
 * mapping between the game's Move type and the kernels' integer move ids.
 */
struct PodGame
{
    virtual ~PodGame() = default;
    virtual std::uint32_t podKind() const = 0;
    virtual std::size_t podSize() const = 0;
    virtual void exportPod(void *destination) const = 0;
};

/// Move <-> integer move id of the kernels; specialised next to each move type.
template<class Move>
struct MoveCodec;

template<>
struct MoveCodec<int>
{
    static std::uint32_t toId(int move) { return static_cast<std::uint32_t>(move); }
    static int fromId(std::uint32_t id) { return static_cast<int>(id); }
};

} // ISMCTS

#endif // ISMCTS_GAME_H
