/*
 * mosolver.h — multiple-observer ISMCTS solver.
 *
 * Same public interface as the reference's MOSolver (include/ismcts/mosolver.h:22-114):
 * MOSolver<Move, ExecutionPolicy, TreePolicies...>, constructors inherited from the policy,
 * Move operator()(POMGame<Move> const&), currentTrees() returning one map player -> root per
 * searched tree set.  One tree per player; in every iteration the tree of the player to move
 * decides and the cursors of all trees follow, creating the node where it is missing
 * (mosolver.h:57-101) — in the engine's kernels.
 */
#ifndef ISMCTS_MOSOLVER_H
#define ISMCTS_MOSOLVER_H

#include "cuda_execution.h"
#include "solverbase.h"

#include <map>

namespace ISMCTS
{

template<class Move, class _ExecutionPolicy = CudaRootParallel, template<class> class... Ps>
class MOSolver : public _ExecutionPolicy, public SolverBase<Move, Ps...>
{
public:
    using _ExecutionPolicy::_ExecutionPolicy;
    using Config = ISMCTS::Config<Move, Ps...>;
    using RootNode = typename Config::RootNode;
    using Player = typename Game<Move>::Player;
    /// One tree for each player (mosolver.h:31-34)
    using TreeMap = std::map<Player, RootNode>;
    using TreeList = std::vector<TreeMap>;

    /// Searches from `rootState`; the decision comes from the tree of the player to move
    /// (mosolver.h:42-46).  Returns Move{} when the game is over.
    Move operator()(POMGame<Move> const &rootState)
    {
        PodGame const &pod = MOSolver::podOf(rootState);
        std::vector<unsigned char> state(pod.podSize());
        pod.exportPod(state.data());
        auto const policy = this->kernelPolicy(rootState);
        m_policyKind = policy.kind;
        m_players = rootState.players();
        auto const &result = this->search(pod.podKind(), state.data(), ISMCTS_SOLVER_MO,
                                          policy.kind, policy.exploration);
        m_treesValid = true;
        if (result.bestMove == 0xFFFFFFFFu)
            return Move{};
        return MoveCodec<Move>::fromId(result.bestMove);
    }

    /// Batched form: every game of `rootStates` searched with the full budget, all in ONE engine call
    /// (see SOSolver::operator()(std::vector<...>)).
    std::vector<Move> operator()(std::vector<POMGame<Move> const *> const &rootStates)
    {
        std::vector<Move> moves;
        if (rootStates.empty())
            return moves;
        std::uint32_t kind = 0;
        auto const states = MOSolver::podsOf(rootStates, kind);
        auto const policy = this->kernelPolicy(*rootStates.front());
        for (auto const *g : rootStates)
            if (g->currentMoveSimultaneous() != rootStates.front()->currentMoveSimultaneous())
                throw std::invalid_argument("ISMCTS: the games of a batch must use the same tree policy");
        m_policyKind = policy.kind;
        auto const &results = this->searchBatch(kind, states.data(), rootStates.size(), ISMCTS_SOLVER_MO,
                                                policy.kind, policy.exploration);
        m_treesValid = false;
        for (auto const &r : results)
            moves.push_back(r.bestMove == 0xFFFFFFFFu ? Move{} : MoveCodec<Move>::fromId(r.bestMove));
        return moves;
    }

    /// The trees of the last search: for each materialised tree set, the root of every player's tree.
    TreeList currentTrees() const
    {
        TreeList trees;
        if (!m_treesValid)
            return trees;
        for (unsigned int t = 0; t < this->materialisedTrees(); ++t) {
            TreeMap map;
            for (Player p : m_players)
                map.emplace(p, MOSolver::materialiseAs(m_policyKind, this->dumpTree(t, p)));
            trees.push_back(std::move(map));
        }
        return trees;
    }

private:
    std::vector<Player> m_players;
    bool m_treesValid {false};
    std::uint32_t m_policyKind {ISMCTS_POLICY_UCB1};
};

} // ISMCTS

#endif // ISMCTS_MOSOLVER_H
