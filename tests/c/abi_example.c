/* abi_example.c — the C ABI from plain C11: what a foreign-function binding of another language would call.
 *
 * Deals a few Knockout Whist positions with the host helpers, searches them with the reference's geometry for this
 * box (16 trees, a block of 128 lanes on each), prints the decisions.  Without a CUDA device ismcts_create must fail
 * with ISMCTS_ERR_NO_DEVICE (there is no CPU path) and the program says so and exits 0 — that is what the CPU test
 * (tests/test_capi_exports.py) runs; tests/test_gpu_parity.py runs it on the GPU. */
#include <ismcts_b200.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define N_POS 4u

int main(void)
{
    ismcts_kw_state roots[N_POS];
    for (uint32_t i = 0; i < N_POS; ++i)
        if (ismcts_kw_deal(2026u, i, 4u, &roots[i]) != ISMCTS_OK) return 2;
    if (sizeof roots[0] != ismcts_state_size(ISMCTS_GAME_KW)) return 3;

    ismcts_ctx *ctx = NULL;
    int rc = ismcts_create(0, &ctx);
    if (rc == ISMCTS_ERR_NO_DEVICE) {
        printf("no device: %s\n", ismcts_last_error(NULL));
        return 0;
    }
    if (rc != ISMCTS_OK) return 4;

    const uint32_t stride = ismcts_move_stride(ISMCTS_GAME_KW);
    ismcts_search_desc d;
    memset(&d, 0, sizeof d);
    d.game = ISMCTS_GAME_KW; d.solver = ISMCTS_SOLVER_SO; d.policy = ISMCTS_POLICY_UCB1;
    d.n_positions = N_POS; d.trees = 16; d.trees_total = 16; d.iterations = 16u * 128u * 130u;
    d.seed = 42; d.exploration = 0.7; d.dump_tree = 0xFFFFFFFFu;
    d.flags = ISMCTS_FLAG_SHARED_TREE | (128u << ISMCTS_FLAG_LANES_SHIFT);

    uint64_t *visits = calloc((size_t)N_POS * stride, sizeof *visits);
    uint64_t *score2 = calloc((size_t)N_POS * stride, sizeof *score2);
    uint64_t *avail = calloc((size_t)N_POS * stride, sizeof *avail);
    uint64_t *iters = calloc((size_t)N_POS * d.trees, sizeof *iters);
    uint32_t best[N_POS];
    ismcts_search_out out;
    memset(&out, 0, sizeof out);
    out.visits = visits; out.score2 = score2; out.avail = avail; out.iterations_done = iters; out.best_move = best;

    rc = ismcts_search(ctx, &d, roots, &out);
    if (rc != ISMCTS_OK) { printf("search failed: %s\n", ismcts_last_error(ctx)); return 5; }
    for (uint32_t i = 0; i < N_POS; ++i) {
        uint64_t total = 0;
        for (uint32_t m = 0; m < stride; ++m) total += visits[(size_t)i * stride + m];
        const uint64_t legal = ismcts_kw_legal(&roots[i]);
        if (total != d.iterations || !((legal >> best[i]) & 1u)) return 6;
        printf("position %u: move %u, %.3f of the visits, root win rate %.3f\n", i, best[i],
               (double)visits[(size_t)i * stride + best[i]] / (double)total,
               0.5 * (double)score2[(size_t)i * stride + best[i]] / (double)visits[(size_t)i * stride + best[i]]);
    }
    printf("kernels launched: %llu (%s)\n", (unsigned long long)ismcts_launch_count(ctx), ismcts_version());
    free(visits); free(score2); free(avail); free(iters);
    ismcts_destroy(ctx);
    return 0;
}
