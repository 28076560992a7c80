// philox.cuh — counter-based Philox4x32-10 stream of one search iteration.
//
// Replaces the reference's thread_local std::mt19937 (utility.h:62-67 and
// knockoutwhist.cpp:28-32).  Word d of an iteration is
//   Philox4x32-10(counter = (d/4, iteration, global tree id, position id),
//                 key = seed)[d % 4],
// (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), so any
// random choice can be recomputed from its coordinates alone and the result of a
// search does not depend on how trees map to threads, warps or GPUs.  A random
// choice is "uniform integer below n" = the high half of word * n; the low half
// is the word of the next choice, so a word serves several consecutive choices: TWO
// in general (the pair deviates from uniform by less than n1 * n2 / 2^32 < 1e-6
// for n <= 225), THREE in Knockout Whist (n <= 52: n1 * n2 * n3 / 2^32 < 3.3e-5; a
// rollout triple, n <= 7, less than 1e-7) — `per_word`.  align() moves
// on to the next block of four words: Knockout Whist aligns around every sequence
// of chance events and at the start of the rollout, which makes the lanes of a
// warp generate their blocks in the same steps (kw_game.cuh).  The oracle
// (oracle/oracle.c: rng_below, rng_align) states the same protocol sequentially.
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

// The stream of one lane.  Words generated ahead wait in a small ring in shared memory (one
// 4-byte column per thread, 8 rows; word d of the iteration sits in row d % 8).  The phase
// loops of Knockout Whist do not use the ring: they keep the current block in registers
// (phase_begin / block / phase_end).
struct LaneRng {
    static constexpr uint32_t kRingWords = 8;
    static constexpr uint32_t kDrawsPerWord = 2;   // unless a game says otherwise (Game::kDrawsPerWord)
    uint32_t k0, k1;      // key = seed
    uint32_t c1, c2, c3;  // iteration, tree, position
    uint32_t next_block;  // next block (d / 4) to generate into the ring
    uint32_t d;           // next word to hand out; words [d, 4 * next_block) wait in the ring
    uint32_t w, left;     // the word being used up and the draws it still serves
    uint32_t per_word;    // draws a word serves
    uint32_t* ring;       // ring[(d % 8) * stride]
    uint32_t stride;

    ISMCTS_DEV void init(uint64_t seed, uint32_t pos, uint32_t tree, uint32_t* ring_base, uint32_t ring_stride)
    {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        c1 = 0; c2 = tree; c3 = pos; next_block = 0; d = 0; w = 0; left = 0; per_word = kDrawsPerWord;
        ring = ring_base; stride = ring_stride;
    }

    // Block `index` of the current iteration.
    ISMCTS_DEV void block(uint32_t index, uint32_t out[4]) const
    {
        uint32_t a = k0, b = k1;
#if defined(__CUDA_ARCH__)
        // (the key as the compiler cannot see through it: the round keys are then derived inside the block — two
        // additions per round — instead of twenty registers held for them across the loops that call this)
        asm volatile("" : "+r"(a), "+r"(b));
#endif
        philox4x32_10(index, c1, c2, c3, a, b, out);
    }

    // Generates the next block of four words into its half of the ring.
    ISMCTS_DEV void refill()
    {
        uint32_t b[4];
        block(next_block, b);
        uint32_t* half = ring + (next_block & 1u) * 4u * stride;
        ++next_block;
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) half[i * stride] = b[i];
    }

    // Starts the stream of `iteration` at word 0 (or at word `at` for the host helpers); nothing
    // is generated until a word is asked for.
    ISMCTS_DEV void begin_iteration(uint32_t iteration, uint32_t at = 0)
    {
        c1 = iteration; d = at; left = 0;
        next_block = at >> 2;
        if (at & 3u) refill();
    }

    ISMCTS_DEV uint32_t waiting() const { return 4u * next_block - d; }
    ISMCTS_DEV bool low() const { return waiting() <= 4u; }   // room for one more block
    ISMCTS_DEV uint32_t consumed() const { return d; }         // words handed out (or skipped by align)

    // The next whole word (EXP3 samples with all 32 bits); drops what is left of the current one.
    ISMCTS_DEV uint32_t word()
    {
        left = 0;
        if (waiting() == 0) refill();
        const uint32_t x = ring[(d & (kRingWords - 1)) * stride];
        ++d;
        return x;
    }

    // Uniform integer in [0, n).
    ISMCTS_DEV uint32_t below(uint32_t n)
    {
        if (left == 0) { w = word(); left = per_word; }
        const uint64_t p = (uint64_t)w * n;
        w = (uint32_t)p;
        --left;
        return (uint32_t)(p >> 32);
    }

    // The same two draws without the ring: the block of word d is generated in registers.  For the
    // streams of games whose loops keep their blocks in registers anyway (Knockout Whist: the only draws
    // left are the few of the tree operations) — a stream uses either these or the ring functions.
    ISMCTS_DEV uint32_t word_direct()
    {
        uint32_t b[4];
        block(d >> 2, b);
        const uint32_t i = d & 3u;
        ++d;
        next_block = (d + 3u) >> 2;   // (the ring stays empty)
        left = 0;
        return i == 0 ? b[0] : i == 1 ? b[1] : i == 2 ? b[2] : b[3];
    }
    ISMCTS_DEV uint32_t below_direct(uint32_t n)
    {
        if (left == 0) { w = word_direct(); left = per_word; }
        const uint64_t p = (uint64_t)w * n;
        w = (uint32_t)p;
        --left;
        return (uint32_t)(p >> 32);
    }

    // word_direct() without taking the word (the stream does not move): for a draw whose word is generated ahead of
    // its use.  below_with_word(word, n) then makes the draw below_direct(n) would have made at the same position.
    ISMCTS_DEV uint32_t peek_word_direct() const
    {
        uint32_t b[4];
        block(d >> 2, b);
        const uint32_t i = d & 3u;
        return i == 0 ? b[0] : i == 1 ? b[1] : i == 2 ? b[2] : b[3];
    }
    ISMCTS_DEV uint32_t below_with_word(uint32_t word, uint32_t n)
    {
        if (left == 0) {
            ++d;
            next_block = (d + 3u) >> 2;
            w = word; left = per_word;
        }
        const uint64_t p = (uint64_t)w * n;
        w = (uint32_t)p;
        --left;
        return (uint32_t)(p >> 32);
    }

    // On to the start of the next block (a no-op at the start of a block with nothing used).
    ISMCTS_DEV void align()
    {
        left = 0;
        d = (d + 3u) & ~3u;
    }

    // A phase that generates its blocks in registers: aligns and returns the index of its first
    // block; `words` words later (four words per block) phase_end() moves the stream past them.  The ring is empty afterwards and the words up to the
    // next block boundary are NOT in it: the next use of the stream must be an align() — which is
    // what follows every phase of Knockout Whist — or the end of the iteration.
    ISMCTS_DEV uint32_t phase_begin() { align(); return d >> 2; }
    ISMCTS_DEV void phase_end(uint32_t first_block, uint32_t words)
    {
        d = 4u * first_block + words;
        next_block = (d + 3u) >> 2;
        left = 0;
    }
};

} // namespace ismcts
