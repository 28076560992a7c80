/*
 * policies.h — tree policies as the CUDA policies see them.
 *
 * In the reference a tree policy is a host functor that picks a child
 * (tree/ucb1.h:60-79, tree/exp3.h:51-97).  Here the selection itself runs in the
 * search kernels; these classes keep the reference's names and constructor
 * parameters and tell the kernels which bandit to run and with which constants.
 */
#ifndef ISMCTS_TREE_POLICIES_H
#define ISMCTS_TREE_POLICIES_H

#include "node.h"
#include "../../ismcts_b200.h"

#include <algorithm>

namespace ISMCTS
{

/// UCB1: score/visits + C * sqrt(ln(available) / visits), first maximum (tree/ucb1.h:36-39,69-76).
template<class Move>
class UCB1
{
public:
    using Node = UCBNode<Move>;
    static constexpr std::uint32_t kernelPolicy = ISMCTS_POLICY_UCB1;

    /// The exploration constant is clamped to be non-negative (tree/ucb1.h:65-67).
    explicit UCB1(double exploration = 0.7) : m_exploration{std::max(exploration, 0.0)} {}
    double exploration() const { return m_exploration; }

private:
    double m_exploration;
};

/// EXP3: sample from the exponentially weighted distribution (tree/exp3.h:57-95).
template<class Move>
class EXP3
{
public:
    using Node = EXPNode<Move>;
    static constexpr std::uint32_t kernelPolicy = ISMCTS_POLICY_EXP3;
    double exploration() const { return 0.7; } // unused by EXP3
};

} // ISMCTS

#endif // ISMCTS_TREE_POLICIES_H
