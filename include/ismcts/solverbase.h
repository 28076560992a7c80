/*
 * solverbase.h — what SOSolver and MOSolver share on the host side.
 *
 * Counterpart of the reference's SolverBase (include/ismcts/solverbase.h:18-88).  There
 * the base class holds the per-iteration steps (simulate, backPropagate, selectNode,
 * selectChild, newRoot, newChild) that run on host threads; here those steps are the
 * kernels' (ismcsolver_b200/csrc/lane_search.cuh) and the base class keeps the policy
 * configuration (setConfig, solverbase.h:23-26), marshals the root position and turns the
 * device node pool into host Node objects.
 */
#ifndef ISMCTS_SOLVERBASE_H
#define ISMCTS_SOLVERBASE_H

#include "game.h"
#include "tree/policies.h"

#include <map>
#include <memory>
#include <stdexcept>
#include <vector>

namespace ISMCTS
{

/// Tree policy for sequential moves (default UCB1) and for simultaneous moves (default EXP3),
/// as in the reference's Config (include/ismcts/config.h:20-48).
template<class Move, template<class> class SeqTree = UCB1, template<class> class SimTree = EXP3>
struct Config
{
    using SeqTreePolicy = SeqTree<Move>;
    using SimTreePolicy = SimTree<Move>;
    using RootNode = std::shared_ptr<Node<Move>>;
    using ChildNode = typename Node<Move>::ChildPtr;
    using TreeList = std::vector<RootNode>;

    SeqTreePolicy seqTreePolicy;
    SimTreePolicy simTreePolicy;

    Config(SeqTreePolicy seq = SeqTreePolicy{}, SimTreePolicy sim = SimTreePolicy{})
        : seqTreePolicy{seq}, simTreePolicy{sim}
    {}
};

template<class Move, template<class> class... Ps>
class SolverBase
{
public:
    /// Replaces the tree policies, e.g. setConfig(UCB1<Card>{1.2}) (solverbase.h:23-26).
    void setConfig(Ps<Move>... policies) { m_config = Config(policies...); }

protected:
    using Config = ISMCTS::Config<Move, Ps...>;
    using RootNode = typename Config::RootNode;
    using SeqPolicy = typename Config::SeqTreePolicy;

    Config const &config() const { return m_config; }

    /// The device form of a game, or a clear error: opaque host games cannot run in a kernel.
    static PodGame const &podOf(Game<Move> const &game)
    {
        auto const *pod = dynamic_cast<PodGame const *>(&game);
        if (!pod)
            throw std::invalid_argument("ISMCTS CUDA policies need a game with a device form (ISMCTS::PodGame); "
                                        "there is no CPU fallback");
        return *pod;
    }

    /// The position records of a list of games, one after the other (all games must be of the same kind).
    template<class GameT>
    static std::vector<unsigned char> podsOf(std::vector<GameT const *> const &games, std::uint32_t &kind)
    {
        std::vector<unsigned char> states;
        kind = 0;
        for (std::size_t i = 0; i < games.size(); ++i) {
            if (!games[i])
                throw std::invalid_argument("ISMCTS: null game in a batch");
            PodGame const &pod = podOf(*games[i]);
            if (i == 0)
                kind = pod.podKind();
            else if (pod.podKind() != kind)
                throw std::invalid_argument("ISMCTS: the games of a batch must be of the same kind");
            states.resize(states.size() + pod.podSize());
            pod.exportPod(states.data() + states.size() - pod.podSize());
        }
        return states;
    }

    /// Which bandit the kernels run and with which constant: the simultaneous-move policy when the
    /// game's moves are simultaneous, else the sequential one (selectChild / newRoot / newChild,
    /// solverbase.h:58-80).  The choice is made once per search, from the root position.
    struct KernelPolicy { std::uint32_t kind; double exploration; };
    KernelPolicy kernelPolicy(Game<Move> const &game) const
    {
        if (game.currentMoveSimultaneous())
            return {Config::SimTreePolicy::kernelPolicy, m_config.simTreePolicy.exploration()};
        return {Config::SeqTreePolicy::kernelPolicy, m_config.seqTreePolicy.exploration()};
    }

    /// Builds host nodes from one tree of the device pool, with the node type of the policy used.
    static RootNode materialiseAs(std::uint32_t kernelPolicyKind, std::vector<ismcts_node_rec> const &records)
    {
        if (kernelPolicyKind == ISMCTS_POLICY_EXP3)
            return materialiseWith<EXPNode<Move>>(records);
        return materialiseWith<UCBNode<Move>>(records);
    }

    /// Builds host nodes from one tree of the device pool (records in creation order, root first).
    static RootNode materialise(std::vector<ismcts_node_rec> const &records)
    {
        return materialiseWith<typename SeqPolicy::Node>(records);
    }

private:
    template<class SeqNode>
    static RootNode materialiseWith(std::vector<ismcts_node_rec> const &records)
    {
        if (records.empty())
            return std::make_shared<SeqNode>();
        std::vector<Node<Move> *> byIndex(records.size(), nullptr);
        auto root = std::make_shared<SeqNode>();
        byIndex[0] = root.get();
        fill(root.get(), records[0]);
        for (std::size_t i = 1; i < records.size(); ++i) {
            ismcts_node_rec const &r = records[i];
            auto node = std::make_unique<SeqNode>(MoveCodec<Move>::fromId(r.move), r.player);
            fill(node.get(), r);
            byIndex[i] = byIndex[r.parent]->addChild(std::move(node)); // parents are created before children
        }
        return root;
    }

    static void fill(UCBNode<Move> *node, ismcts_node_rec const &r)
    {
        node->setVisits(r.visits);
        node->setStatistics(r.score, r.avail);
    }
    static void fill(EXPNode<Move> *node, ismcts_node_rec const &r)
    {
        node->setVisits(r.visits);
        node->setStatistics(r.score, r.prob);
    }

    Config m_config;
};

} // ISMCTS

#endif // ISMCTS_SOLVERBASE_H
