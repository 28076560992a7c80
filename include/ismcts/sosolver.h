/*
 * sosolver.h — single-observer ISMCTS solver.
 *
 * Same public interface as the reference's SOSolver (include/ismcts/sosolver.h:21-76):
 * SOSolver<Move, ExecutionPolicy, TreePolicies...>, constructors inherited from the
 * policy, Move operator()(Game<Move> const&), currentTrees().  One tree for the player to
 * move at the root; every iteration determinises the root for that player, descends with
 * the sequential tree policy, expands one node, plays out uniformly at random and
 * backpropagates (sosolver.h:44-71) — in the engine's kernels.
 */
#ifndef ISMCTS_SOSOLVER_H
#define ISMCTS_SOSOLVER_H

#include "cuda_execution.h"
#include "solverbase.h"

namespace ISMCTS
{

template<class Move, class _ExecutionPolicy = CudaRootParallel, template<class> class... Ps>
class SOSolver : public _ExecutionPolicy, public SolverBase<Move, Ps...>
{
public:
    using _ExecutionPolicy::_ExecutionPolicy;
    using Config = ISMCTS::Config<Move, Ps...>;
    using RootNode = typename Config::RootNode;
    using TreeList = typename Config::TreeList;

    /// Searches from `rootState` and returns the most visited root move (ties: the first in
    /// ascending move order).  The game is not modified.  Returns Move{} when it is over.
    Move operator()(Game<Move> const &rootState)
    {
        PodGame const &pod = SOSolver::podOf(rootState);
        std::vector<unsigned char> state(pod.podSize());
        pod.exportPod(state.data());
        auto const policy = this->kernelPolicy(rootState);
        m_policyKind = policy.kind;
        auto const &result = this->search(pod.podKind(), state.data(), ISMCTS_SOLVER_SO,
                                          policy.kind, policy.exploration);
        m_treesValid = true;
        if (result.bestMove == 0xFFFFFFFFu)
            return Move{};
        return MoveCodec<Move>::fromId(result.bestMove);
    }

    /// Batched form: searches every game of `rootStates` with the full budget each, all of them in ONE engine
    /// call (the GPU is filled by positions x trees; a single position only reaches a few per cent of it), and
    /// returns their moves in order.  Equivalent to calling operator() on each game in turn, sosolver.h:30-36.
    /// The games must be of one kind and agree on currentMoveSimultaneous().
    std::vector<Move> operator()(std::vector<Game<Move> const *> const &rootStates)
    {
        std::vector<Move> moves;
        if (rootStates.empty())
            return moves;
        std::uint32_t kind = 0;
        auto const states = SOSolver::podsOf(rootStates, kind);
        auto const policy = this->kernelPolicy(*rootStates.front());
        for (auto const *g : rootStates)
            if (g->currentMoveSimultaneous() != rootStates.front()->currentMoveSimultaneous())
                throw std::invalid_argument("ISMCTS: the games of a batch must use the same tree policy");
        m_policyKind = policy.kind;
        auto const &results = this->searchBatch(kind, states.data(), rootStates.size(), ISMCTS_SOLVER_SO,
                                                policy.kind, policy.exploration);
        m_treesValid = false;
        for (auto const &r : results)
            moves.push_back(r.bestMove == 0xFFFFFFFFu ? Move{} : MoveCodec<Move>::fromId(r.bestMove));
        return moves;
    }

    /// The trees of the last search, one root per tree (a copy; the first
    /// setMaterialisedTrees() trees are read back from the device).
    TreeList currentTrees() const
    {
        TreeList trees;
        if (!m_treesValid)
            return trees;
        for (unsigned int t = 0; t < this->materialisedTrees(); ++t)
            trees.push_back(SOSolver::materialiseAs(m_policyKind, this->dumpTree(t)));
        return trees;
    }

private:
    bool m_treesValid {false};
    std::uint32_t m_policyKind {ISMCTS_POLICY_UCB1};
};

} // ISMCTS

#endif // ISMCTS_SOSOLVER_H
