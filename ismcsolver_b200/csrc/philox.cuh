// philox.cuh — counter-based Philox4x32-10 stream of one search iteration.
//
// Replaces the reference's thread_local std::mt19937 (utility.h:62-67 and
// knockoutwhist.cpp:28-32).  Word d of an iteration is
//   Philox4x32-10(counter = (d/4, iteration, global tree id, position id),
//                 key = seed)[d % 4],
// (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), so any
// random choice can be recomputed from its coordinates alone and the result of a
// search does not depend on how trees map to threads, warps or GPUs.  Every
// random choice of the engine is "uniform integer below n" = high word of
// word * n, and consumes exactly one word.
#pragma once

#include "devdefs.cuh"

namespace ismcts {

ISMCTS_DEV void philox4x32_10(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t a, uint32_t b,
                              uint32_t out[4])
{
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        x0 = hi1 ^ x1 ^ a;
        x1 = lo1;
        x2 = hi0 ^ x3 ^ b;
        x3 = lo0;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}

// The stream of one lane.  Generated words wait in a small ring in shared memory
// (one 4-byte column per thread, 8 rows; word d of the iteration sits in row
// d % 8): the search loop refills all lanes of a warp in the same steps, so the
// ten Philox rounds are executed with all lanes active although the lanes
// consume their words at different rates.
struct LaneRng {
    static constexpr uint32_t kRingWords = 8;
    uint32_t k0, k1;      // key = seed
    uint32_t c1, c2, c3;  // iteration, tree, position
    uint32_t next_block;  // next block (d / 4) to generate
    uint32_t d;           // next word to hand out; words [d, 4 * next_block) wait in the ring
    uint32_t* ring;       // ring[(d % 8) * stride]
    uint32_t stride;

    ISMCTS_DEV void init(uint64_t seed, uint32_t pos, uint32_t tree, uint32_t* ring_base, uint32_t ring_stride)
    {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        c1 = 0; c2 = tree; c3 = pos; next_block = 0; d = 0;
        ring = ring_base; stride = ring_stride;
    }

    // Generates the next block of four words into its half of the ring.
    ISMCTS_DEV void refill()
    {
        uint32_t w[4];
        philox4x32_10(next_block, c1, c2, c3, k0, k1, w);
        uint32_t* half = ring + (next_block & 1u) * 4u * stride;
        ++next_block;
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) half[i * stride] = w[i];
    }

    // Starts the stream of `iteration` at word 0 (or at word `at` for the host helpers).
    ISMCTS_DEV void begin_iteration(uint32_t iteration, uint32_t at = 0)
    {
        c1 = iteration; next_block = at >> 2; d = at;
        refill();
    }

    ISMCTS_DEV uint32_t waiting() const { return 4u * next_block - d; }
    ISMCTS_DEV bool low() const { return waiting() <= 4u; }   // room for one more block
    ISMCTS_DEV uint32_t consumed() const { return d; }         // words handed out

    // Next word when one is known to wait: the loops refill every lane that is low() at least
    // every fourth step and draw at most one word per step.
    ISMCTS_DEV uint32_t word_waiting()
    {
        const uint32_t w = ring[(d & (kRingWords - 1)) * stride];
        ++d;
        return w;
    }

    // Next 32-bit word of the stream; generates the block it needs when nothing waits (a lane
    // that draws several words in one step: an EXP3 descent draws one per level).
    ISMCTS_DEV uint32_t word()
    {
        if (waiting() == 0) refill();
        return word_waiting();
    }

    // Uniform integer in [0, n) — always consumes exactly one word.
    ISMCTS_DEV uint32_t below(uint32_t n) { return mulhi32(word(), n); }
    ISMCTS_DEV uint32_t below_waiting(uint32_t n) { return mulhi32(word_waiting(), n); }
};

} // namespace ismcts
