/*
 * node.h — host-side view of a search tree.
 *
 * The CUDA policies keep the trees in a flat node pool in device memory
 * (ismcsolver_b200/csrc/pool.cuh).  currentTrees() materialises them into these
 * objects, which offer the read interface of the reference's nodes
 * (include/ismcts/tree/node.h:23-151, tree/ucb1.h:23-58, tree/exp3.h:23-49): parent,
 * children in creation order, move, player, visits, the policy statistics, depth /
 * height, untriedMoves and the textual dump "[M:<move> by <player>, V/S/A: v/s/a]".
 */
#ifndef ISMCTS_TREE_NODE_H
#define ISMCTS_TREE_NODE_H

#include <algorithm>
#include <cstddef>
#include <iomanip>
#include <memory>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

namespace ISMCTS
{

template<class Move>
class Node
{
public:
    using ChildPtr = std::unique_ptr<Node>;

    explicit Node(Move const &move = {}, unsigned int player = 0) : m_move{move}, m_player{player} {}
    virtual ~Node() = default;
    Node(Node const &) = delete;
    Node &operator=(Node const &) = delete;

    Node *parent() const { return m_parent; }
    std::vector<ChildPtr> const &children() const { return m_children; }
    Move const &move() const { return m_move; }
    unsigned int player() const { return m_player; }
    unsigned int visits() const { return m_visits; }

    /// Adopts `child` as the last child of this node and returns it.
    Node *addChild(ChildPtr child)
    {
        child->m_parent = this;
        m_children.push_back(std::move(child));
        return m_children.back().get();
    }

    /// The legal moves that have no child node yet, in the order given.
    std::vector<Move> untriedMoves(std::vector<Move> const &legalMoves) const
    {
        std::vector<Move> untried;
        for (auto const &move : legalMoves) {
            bool const tried = std::any_of(m_children.begin(), m_children.end(),
                                           [&](ChildPtr const &c) { return c->m_move == move; });
            if (!tried)
                untried.push_back(move);
        }
        return untried;
    }

    virtual operator std::string() const { return ""; }

    std::string treeToString(unsigned int indent = 0) const
    {
        std::string text;
        for (unsigned int i = 0; i < indent; ++i)
            text += "| ";
        text += std::string(*this) + "\n";
        for (auto const &child : m_children)
            text += child->treeToString(indent + 1);
        return text;
    }

    std::size_t depth() const
    {
        std::size_t d = 0;
        for (Node const *n = m_parent; n; n = n->m_parent)
            ++d;
        return d;
    }

    std::size_t height() const
    {
        std::size_t h = 0;
        for (auto const &child : m_children)
            h = std::max(h, 1 + child->height());
        return h;
    }

    friend std::ostream &operator<<(std::ostream &out, Node const &node) { return out << std::string(node); }

    /// Statistics as read back from the device pool.
    void setVisits(unsigned int visits) { m_visits = visits; }

private:
    Node *m_parent {nullptr};
    std::vector<ChildPtr> m_children;
    Move m_move;
    unsigned int m_player;
    unsigned int m_visits {0};
};

/// Node of the UCB1 policy: total reward and availability count (tree/ucb1.h:23-58).
template<class Move>
class UCBNode : public Node<Move>
{
public:
    using Node<Move>::Node;
    double score() const { return m_score; }
    unsigned int available() const { return m_available; }
    void setStatistics(double score, unsigned int available) { m_score = score; m_available = available; }

    operator std::string() const override
    {
        std::ostringstream text;
        text << "[M:" << this->move() << " by " << this->player() << ", V/S/A: " << std::fixed << std::setprecision(1)
             << this->visits() << "/" << m_score << "/" << m_available << "]";
        return text.str();
    }

private:
    double m_score {0};
    unsigned int m_available {1};
};

/// Node of the EXP3 policy: importance-weighted reward and last selection probability (tree/exp3.h:23-49).
template<class Move>
class EXPNode : public Node<Move>
{
public:
    using Node<Move>::Node;
    double score() const { return m_score; }
    double probability() const { return m_probability; }
    void setStatistics(double score, double probability) { m_score = score; m_probability = probability; }

    operator std::string() const override
    {
        std::ostringstream text;
        text << "[M:" << this->move() << " by " << this->player() << ", V/S/P: " << std::fixed << std::setprecision(1)
             << this->visits() << "/" << m_score << "/" << std::setprecision(2) << m_probability << "]";
        return text.str();
    }

private:
    double m_score {0};
    double m_probability {1};
};

} // ISMCTS

#endif // ISMCTS_TREE_NODE_H
