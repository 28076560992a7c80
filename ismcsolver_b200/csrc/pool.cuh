// pool.cuh — flat node pool of the search tree(s) of one lane, in HBM.
//
// Replaces the reference's pointer tree (tree/node.h:113-118: parent pointer,
// vector<unique_ptr<Node>> children, one std::mutex per node; UCBNode adds
// atomic<double> score and atomic_uint available, ucb1.h:50-51; EXPNode adds
// atomic<double> probability and score, exp3.h:42-43).  A pool is an
// open-addressing hash table of fixed-size slots keyed by (parent slot, move):
// finding the child for a move is one slot load (one or two 32-byte DRAM
// sectors) instead of a walk over the child vector (node.h:58-72,
// O(children x legal)), and the per-node `child_mask` answers untriedMoves
// (node.h:82-91) with one AND.  The first `roots` slots are the root nodes: one
// for SO-ISMCTS, one per player for MO-ISMCTS (mosolver.h:31-34,106-112) — the
// trees of all players share one table, their keys cannot collide because their
// parent slots differ.
//
// UCB1 slot: (visits, score2) share one aligned 64-bit word so that
// backpropagation is ONE fire-and-forget 64-bit reduction per node
// (visits += 1, score2 += reward*2 — vectorised backpropagation).
#pragma once

#include "devdefs.cuh"
#include "bitset.cuh"

namespace ismcts {

template <class MaskType>
struct alignas(32) SlotUcb {
    using Mask = MaskType;
    static constexpr bool kExp3 = false;
    uint32_t visits;     // m_visits                      } one 64-bit word:
    uint32_t score2;     // 2 * m_score (rewards 0,1/2,1) } low = visits, high = score2
    uint32_t key;        // ((parent slot << 8) | move) + 1; 0 = empty; roots = kRootKey
    uint32_t avail;      // m_available (ucb1.h:51)
    uint32_t created;    // creation index within its tree (root 0): order of the reference's child vector
    uint32_t player;     // m_playerJustMoved
    Mask child_mask;     // moves that have a child

    ISMCTS_DEV void set_fresh() { visits = 0; score2 = 0; avail = 1; }
};

template <class MaskType>
struct alignas(32) SlotExp {
    using Mask = MaskType;
    static constexpr bool kExp3 = true;
    uint32_t visits;     // m_visits
    uint32_t pad0;
    uint32_t key;
    uint32_t pad1;
    uint32_t created;
    uint32_t player;
    Mask child_mask;
    double score;        // m_score: sum of reward / probability (exp3.h:45-48)
    double prob;         // m_probability, 1 until the node is first offered (exp3.h:42)

    ISMCTS_DEV void set_fresh() { visits = 0; pad0 = 0; pad1 = 0; score = 0.0; prob = 1.0; }
};

constexpr uint32_t kRootKey = 0xFFFFFFFFu;
constexpr uint32_t kMaxSlotsLog2 = 23; // parent slot index must fit 24 bits minus one

ISMCTS_HDI uint32_t slot_make_key(uint32_t parent, uint32_t move) { return ((parent << 8) | move) + 1u; }
ISMCTS_HDI uint32_t slot_parent_of(uint32_t key) { return (key - 1u) >> 8; }
ISMCTS_HDI uint32_t slot_move_of(uint32_t key) { return (key - 1u) & 0xFFu; }

// Slot loads bypass L1 (ld.global.cg): statistics are updated by reductions that
// are performed in L2, possibly by other lanes sharing the tree.
template <class Slot>
ISMCTS_DEV Slot load_slot(const Slot* p)
{
#if defined(__CUDA_ARCH__)
    static_assert(sizeof(Slot) % 16 == 0, "slot is a whole number of 16-byte words");
    Slot s;
    const uint4* src = reinterpret_cast<const uint4*>(p);
    uint4* dst = reinterpret_cast<uint4*>(&s);
#pragma unroll
    for (uint32_t i = 0; i < sizeof(Slot) / 16; ++i) dst[i] = __ldcg(src + i);
    return s;
#else
    return *p;
#endif
}

template <class SlotType>
struct TreePool {
    using Slot = SlotType;
    using Mask = typename Slot::Mask;
    Slot* slots;     // this lane's table
    uint32_t mask;   // capacity - 1 (capacity is a power of two)
    uint32_t shift;  // 32 - log2(capacity)
    uint32_t count;  // nodes created so far in all trees of the table, roots included
    uint32_t roots;  // slots [0, roots) are root nodes

    ISMCTS_DEV void bind(Slot* base, uint32_t log2cap, uint32_t n_roots)
    {
        slots = base; mask = (1u << log2cap) - 1u; shift = 32u - log2cap; count = n_roots; roots = n_roots;
    }

    ISMCTS_DEV uint32_t home(uint32_t key) const { return (key * 2654435769u) >> shift; }

    // Writes the roots (default-constructed nodes, node.h:30-33).  The caller
    // guarantees the table memory is zero.
    ISMCTS_DEV void init_roots()
    {
        Slot r;
        r.set_fresh();
        r.key = kRootKey; r.created = 0; r.player = 0; r.child_mask = Mask::none();
        for (uint32_t i = 0; i < roots; ++i) slots[i] = r;
        count = roots;
    }

    // Slot index (and contents) of the child of `parent` reached by `move`; the child must exist.
    ISMCTS_DEV uint32_t find(uint32_t parent, uint32_t move, Slot& out) const
    {
        const uint32_t key = slot_make_key(parent, move);
        uint32_t i = home(key);
        for (;;) {
            if (i >= roots) {
                out = load_slot(&slots[i]);
                if (out.key == key) return i;
            }
            i = (i + 1u) & mask;
        }
    }

    // find() in two halves, for callers that keep the load of one child in flight while they work on another:
    // home_slot() is where the search for the child starts, resolve() finishes it from the contents of that
    // slot (loaded by the caller), looking further only if the home slot holds another node.
    ISMCTS_DEV uint32_t home_slot(uint32_t parent, uint32_t move) const { return home(slot_make_key(parent, move)); }
    ISMCTS_DEV uint32_t resolve(uint32_t parent, uint32_t move, uint32_t i, Slot& contents) const
    {
        const uint32_t key = slot_make_key(parent, move);
        if (i >= roots && contents.key == key) return i;
        for (;;) {
            i = (i + 1u) & mask;
            if (i >= roots) {
                contents = load_slot(&slots[i]);
                if (contents.key == key) return i;
            }
        }
    }

    // Starts loading the home slot of the child of `parent` for `move`: a descent looks its children up
    // one after the other, and without this their DRAM latencies add up.
    ISMCTS_DEV void prefetch(uint32_t parent, uint32_t move) const { prefetch_l2(&slots[home(slot_make_key(parent, move))]); }

    // Shared tree (tree parallelisation): the child of `parent` for `move`, created if no lane has
    // created it yet (Node::addChild under the node mutex, node.h:44-48,121-126, without the
    // mutex).  A slot is reserved by a compare-and-swap on its key.  The lane that wins (`won`) fills in the
    // fields nobody else writes; the caller counts its visit and then PUBLISHES the child with a release on the
    // parent's child_mask, so that whoever sees the bit sees the whole node.  A lane that finds the key reserved
    // by another (`won` = false) gets the slot too — its visit is a reduction on a field that started at zero —
    // but must wait for that bit before it reads anything else of the node.  The availability starts as an
    // increment of the zeroed field: increments of other lanes commute with it.  `created`: the node's place
    // in the creation order of its tree (never 0 for a child).
    ISMCTS_DEV uint32_t find_or_insert_shared(uint32_t parent, uint32_t move, uint32_t player, uint32_t created, bool& won)
    {
        const uint32_t key = slot_make_key(parent, move);
        uint32_t i = home(key);
        for (;;) {
            if (i >= roots) {
                const uint32_t old = atomic_cas_u32(&slots[i].key, 0u, key);
                if (old == 0u) {
                    if constexpr (!Slot::kExp3) red_add_u32(&slots[i].avail, 1u);      // ucb1.h:51
                    else slots[i].prob = 1.0;                                          // exp3.h:42
                    slots[i].player = player;
                    slots[i].created = created;
                    won = true;
                    return i;
                }
                if (old == key) { won = false; return i; }    // another lane was first
            }
            i = (i + 1u) & mask;
        }
    }

    // addChild + newChild, node.h:44-48 / solverbase.h:74-80: a fresh node
    // (available = 1, ucb1.h:51; probability = 1, exp3.h:42).  `created` is its
    // creation index within its own tree.  Returns its slot index.
    ISMCTS_DEV uint32_t insert(uint32_t parent, uint32_t move, uint32_t player, uint32_t created)
    {
        const uint32_t key = slot_make_key(parent, move);
        uint32_t i = home(key);
        while (i < roots || load_u32_cg(&slots[i].key) != 0) i = (i + 1u) & mask;
        Slot s;
        s.set_fresh();
        s.key = key; s.created = created; s.player = player; s.child_mask = Mask::none();
        slots[i] = s;
        ++count;
        return i;
    }
};

} // namespace ismcts
