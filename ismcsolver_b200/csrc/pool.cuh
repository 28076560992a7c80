// pool.cuh — flat node pool of one search tree in HBM.
//
// Replaces the reference's pointer tree (tree/node.h:113-118: parent pointer,
// vector<unique_ptr<Node>> children, one std::mutex per node; UCBNode adds
// atomic<double> score and atomic_uint available, ucb1.h:50-51).  A tree is an
// open-addressing hash table of fixed-size slots keyed by (parent slot, move):
// finding the child for a move is one slot load (one or two 32-byte DRAM
// sectors) instead of a walk over the child vector (node.h:58-72,
// O(children x legal)), and the per-node `child_mask` answers untriedMoves
// (node.h:82-91) with one AND.
//
// Slot layout: (visits, score2) share one aligned 64-bit word so that
// backpropagation is ONE fire-and-forget 64-bit reduction per node
// (visits += 1, score2 += reward*2 — vectorised backpropagation), which is also
// what makes a tree shareable between the lanes of a wave.
#pragma once

#include "devdefs.cuh"
#include "bitset.cuh"

namespace ismcts {

template <class Mask>
struct alignas(32) SlotT {
    uint32_t visits;     // m_visits                      } one 64-bit word:
    uint32_t score2;     // 2 * m_score (rewards 0,1/2,1) } low = visits, high = score2
    uint32_t key;        // ((parent slot << 8) | move) + 1; 0 = empty; root = kRootKey
    uint32_t avail;      // m_available (ucb1.h:51)
    uint32_t created;    // creation index (root 0): the order of the reference's child vector
    uint32_t player;     // m_playerJustMoved
    Mask child_mask;     // moves that have a child
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

template <class Mask>
struct TreePool {
    using Slot = SlotT<Mask>;
    Slot* slots;     // this tree's table
    uint32_t mask;   // capacity - 1 (capacity is a power of two)
    uint32_t shift;  // 32 - log2(capacity)
    uint32_t count;  // nodes created so far, root included

    ISMCTS_DEV void bind(Slot* base, uint32_t log2cap)
    {
        slots = base; mask = (1u << log2cap) - 1u; shift = 32u - log2cap; count = 1;
    }

    ISMCTS_DEV uint32_t home(uint32_t key) const { return (key * 2654435769u) >> shift; }

    // Writes the root (a default-constructed node, node.h:30-33) into slot 0.
    // The caller guarantees the table memory is zero.
    ISMCTS_DEV void init_root()
    {
        Slot r;
        r.visits = 0; r.score2 = 0; r.key = kRootKey; r.avail = 1; r.created = 0; r.player = 0;
        r.child_mask = Mask::none();
        slots[0] = r;
        count = 1;
    }

    // Slot index (and contents) of the child of `parent` reached by `move`; the child must exist.
    ISMCTS_DEV uint32_t find(uint32_t parent, uint32_t move, Slot& out) const
    {
        const uint32_t key = slot_make_key(parent, move);
        uint32_t i = home(key);
        for (;;) {
            if (i != 0) { // slot 0 is the root
                out = load_slot(&slots[i]);
                if (out.key == key) return i;
            }
            i = (i + 1u) & mask;
        }
    }

    // addChild + newChild, node.h:44-48 / solverbase.h:74-80: a fresh node with
    // available = 1 (ucb1.h:51).  Returns its slot index.
    ISMCTS_DEV uint32_t insert(uint32_t parent, uint32_t move, uint32_t player)
    {
        const uint32_t key = slot_make_key(parent, move);
        uint32_t i = home(key);
        while (i == 0 || load_slot(&slots[i]).key != 0) i = (i + 1u) & mask;
        Slot s;
        s.visits = 0; s.score2 = 0; s.key = key; s.avail = 1; s.created = count++; s.player = player;
        s.child_mask = Mask::none();
        slots[i] = s;
        return i;
    }
};

} // namespace ismcts
