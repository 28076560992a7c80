/*
 * cuda_execution.h — CUDA execution policies: ISMCTS::CudaRootParallel and
 * ISMCTS::CudaTreeParallel.
 *
 * Drop-in counterparts of the reference's std::async execution policies
 * (include/ismcts/execution.h:26-232).  They keep the policy interface the solvers
 * and their callers rely on —
 *   explicit Policy(std::size_t iterationCount = 1000, unsigned n = default)
 *   explicit Policy(Duration iterationTime, unsigned n = default)
 *   iterationCount() / setIterationCount() / iterationTime() / setIterationTime()
 *   (setting one budget zeroes the other, execution.h:42-63), numThreads(),
 *   non-copyable (execution.h:34-35)
 * — and run the search iterations in the engine's sm_100a kernels through the C ABI
 * of include/ismcts_b200.h instead of on host threads.  `n` is the number of
 * independent trees (root parallel; the reference: one tree per thread) or the
 * number of iterations in flight on the one shared tree (tree parallel; the
 * reference: threads sharing tree 0).
 *
 * There is no CPU path: without a usable CUDA device every search throws.
 */
#ifndef ISMCTS_CUDA_EXECUTION_H
#define ISMCTS_CUDA_EXECUTION_H

#include "../ismcts_b200.h"

#include <algorithm>
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace ISMCTS
{

/// Thrown for engine failures (no device, CUDA errors, capacity); illegal moves keep std::out_of_range.
struct CudaError : std::runtime_error
{
    CudaError(int errorCode, std::string const &what) : std::runtime_error{what}, code{errorCode} {}
    int code;
};

namespace detail
{

/// What one search call produced for the root position (summed over trees and devices).
struct RootStatistics
{
    std::vector<std::uint64_t> visits, score2, available; // per move id
    std::vector<std::uint64_t> iterationsPerTree;
    std::uint32_t bestMove {0xFFFFFFFFu};                 // first maximum of visits, ascending move id
    std::uint64_t iterations() const
    {
        std::uint64_t n = 0;
        for (auto i : iterationsPerTree)
            n += i;
        return n;
    }
};

/// RAII owner of one engine context (one CUDA device).
class Context
{
public:
    explicit Context(int device) : m_device{device}
    {
        int const rc = ismcts_create(device, &m_ctx);
        if (rc != ISMCTS_OK)
            throw CudaError{rc, std::string{"ISMCTS CUDA engine: "} + ismcts_last_error(nullptr)};
    }
    ~Context() { ismcts_destroy(m_ctx); }
    Context(Context const &) = delete;
    Context &operator=(Context const &) = delete;

    ismcts_ctx *get() const { return m_ctx; }
    int device() const { return m_device; }

    void check(int rc) const
    {
        if (rc != ISMCTS_OK)
            throw CudaError{rc, std::string{"ISMCTS CUDA engine: "} + ismcts_last_error(m_ctx)};
    }

private:
    ismcts_ctx *m_ctx {nullptr};
    int m_device;
};

} // detail

class CudaExecutionPolicy
{
public:
    using Duration = std::chrono::duration<double>;

    CudaExecutionPolicy(CudaExecutionPolicy const &) = delete;
    CudaExecutionPolicy &operator=(CudaExecutionPolicy const &) = delete;
    virtual ~CudaExecutionPolicy() = default;

    std::size_t iterationCount() const { return m_iterCount; }
    void setIterationCount(std::size_t count)
    {
        m_iterCount = count;
        m_iterTime = Duration::zero();
    }

    Duration iterationTime() const { return m_iterTime; }
    void setIterationTime(Duration time)
    {
        m_iterTime = time;
        m_iterCount = 0;
    }

    /// Trees (root parallel) or iterations in flight on the shared tree (tree parallel).
    /// 0 was given to the constructor: chosen per search from the budget (see treesFor()).
    unsigned int numThreads() const { return m_parallelism ? m_parallelism : treesFor(m_iterCount); }

    // ---- CUDA-specific knobs (defaulted so that reference call sites compile unchanged)

    /// Fixes the Philox seed: searches become reproducible (the reference is unseeded, utility.h:62-67).
    void setSeed(std::uint64_t seed) { m_seed = seed; m_seedFixed = true; }
    std::uint64_t seed() const { return m_seed; }

    /// CUDA devices to search on; root-parallel trees are sharded over them.
    void setDevices(std::vector<int> devices)
    {
        if (devices.empty())
            throw std::invalid_argument("setDevices: at least one device");
        m_devices = std::move(devices);
        m_contexts.clear();
    }
    std::vector<int> const &devices() const { return m_devices; }

    /// Root parallel: lanes that search each tree together, with virtual loss.  0 (default): chosen per search — a
    /// block of 128 when the positions of the call fill the GPU, more when they do not (a single 1M-iteration
    /// decision: 488 lanes on each of its 16 trees), never so many that a lane keeps fewer than
    /// minIterationsPerLane iterations.  1: every tree is searched sequentially by one lane, exactly as a thread of
    /// the reference's RootParallel does, and a seeded search is reproducible bit for bit; with more lanes the
    /// interleaving inside a tree is the hardware's.
    void setLanesPerTree(unsigned int lanes) { m_lanesPerTree = lanes; }
    unsigned int lanesPerTree() const { return m_lanesPerTree; }

    /// How many of the searched trees currentTrees() materialises on the host (default 8).
    void setMaterialisedTrees(unsigned int count) { m_materialised = count; }

    /// Statistics of the last search (root visit / reward / availability sums, iterations run).
    detail::RootStatistics const &lastStatistics() const { return m_last; }
    /// ... of every position of the last batched search (operator() on a list of games).
    std::vector<detail::RootStatistics> const &lastBatchStatistics() const { return m_lastBatch; }

protected:
    /// Fewest iterations a tree gets when the number of trees is chosen automatically.  Spreading a
    /// small budget over many trees leaves every root child with one visit (solvertest.cpp:131-138
    /// expects a decision from 16 iterations), so small budgets use one tree.
    static constexpr std::size_t minIterationsPerTree = 1024;

    CudaExecutionPolicy(std::size_t iterationCount, unsigned int parallelism, bool sharedTree)
        : m_parallelism{parallelism}, m_sharedTree{sharedTree}
    {
        setIterationCount(iterationCount);
        m_seed = std::random_device{}();
        m_seed = (m_seed << 32) ^ std::random_device{}();
    }

    CudaExecutionPolicy(Duration iterationTime, unsigned int parallelism, bool sharedTree)
        : CudaExecutionPolicy{std::size_t{0}, parallelism, sharedTree}
    {
        setIterationTime(iterationTime);
    }

    /// Root parallel with the number of trees left open: as many trees as the reference's RootParallel builds by
    /// default — one per hardware thread of the host (execution.h:160-167,181-189) — so that the search has the
    /// reference's geometry (a 1M-iteration decision on a 16-core host: 16 trees of 62,500 iterations, whose
    /// decisions agree with the reference's, profiles/r02_head_to_head.md), but not more than the budget feeds.
    unsigned int treesFor(std::size_t iterations) const
    {
        if (m_sharedTree)
            return 1;
        unsigned int const host = std::max(1u, std::thread::hardware_concurrency());
        if (iterations == 0)
            return host;  // time mode: every tree iterates until the deadline
        return static_cast<unsigned int>(std::min<std::size_t>(std::max<std::size_t>(iterations / minIterationsPerTree, 1), host));
    }

    /// One resident wave of the search kernel on a B200: 148 SMs x 8 blocks x 128 lanes.
    static constexpr std::size_t residentLanes = 148 * 8 * 128;

    /// Root parallel: the lanes that search one tree (setLanesPerTree), capped like the lanes of tree parallelisation.
    unsigned int lanesPerTreeFor(std::size_t iterations, unsigned int trees, std::size_t positions) const
    {
        std::size_t lanes = m_lanesPerTree;
        if (lanes == 0)   // a block per tree, or what it takes to fill the GPU with the positions at hand
            lanes = std::max<std::size_t>(128, residentLanes / std::max<std::size_t>(1, positions * trees));
        if (iterations == 0)
            return static_cast<unsigned int>(std::min<std::size_t>(lanes, 1024));
        std::size_t const cap = iterations / trees / minIterationsPerLane;
        return static_cast<unsigned int>(std::max<std::size_t>(1, std::min<std::size_t>(lanes, cap)));
    }

    bool sharedTree() const { return m_sharedTree; }

    /// Tree parallel: iterations in flight on the shared tree.  Every lane in flight adds a virtual loss and
    /// searches on statistics that are `lanes` iterations old, so the lanes are capped at
    /// iterations / minIterationsPerLane: measured on m,n,k boards from 5x5 to 13x13 and on Knockout Whist at 1M
    /// iterations (profiles/r02_tree_parallel_quality.md), the decision equals the sequential one down to about
    /// 120 iterations per lane, begins to drift at 30 and is lost at 7 (16 iterations -> 1 lane, 1M -> 7812).
    static constexpr std::size_t minIterationsPerLane = 128;
    unsigned int lanesFor(std::size_t iterations) const
    {
        if (iterations == 0)
            return m_parallelism;
        return static_cast<unsigned int>(std::max<std::size_t>(1, std::min<std::size_t>(m_parallelism, iterations / minIterationsPerLane)));
    }

    /// Runs one search of `podState` (a position record of include/ismcts_b200.h) on the configured
    /// devices and merges the per-device root arrays (RootParallel::compileVisitCounts,
    /// execution.h:219-231, across devices).
    detail::RootStatistics const &search(std::uint32_t game, void const *podState, std::uint32_t solver,
                                         std::uint32_t policy, double exploration)
    {
        searchBatch(game, podState, 1, solver, policy, exploration);
        m_last = m_lastBatch.front();
        return m_last;
    }

    /// The same for `count` positions (consecutive records at `podStates`) in ONE engine call per device: every
    /// position gets the whole budget (iterationCount() iterations, or iterationTime()), as if operator() had
    /// been called on each of them in turn (sosolver.h:30-36), but the GPU works on all of them together.
    std::vector<detail::RootStatistics> const &searchBatch(std::uint32_t game, void const *podStates, std::size_t count,
                                                           std::uint32_t solver, std::uint32_t policy, double exploration)
    {
        if (count == 0) {
            m_lastBatch.clear();
            return m_lastBatch;
        }
        if (m_contexts.empty())
            for (int device : m_devices)
                m_contexts.push_back(std::make_unique<detail::Context>(device));
        if (!m_seedFixed)
            ++m_seed;

        std::uint32_t const stride = ismcts_move_stride(game);
        unsigned int treesTotal = m_parallelism && !m_sharedTree ? m_parallelism : treesFor(m_iterCount);
        if (m_iterCount > 0)
            treesTotal = static_cast<unsigned int>(std::min<std::size_t>(treesTotal, m_iterCount));

        ismcts_search_desc desc {};
        desc.game = game;
        desc.solver = solver;
        desc.policy = policy;
        desc.n_positions = static_cast<std::uint32_t>(count);
        desc.trees = treesTotal;
        desc.trees_total = treesTotal;
        desc.tree_base = 0;
        desc.pos_base = 0;
        desc.iterations = m_iterCount;
        desc.time_ns = static_cast<std::uint64_t>(m_iterTime.count() * 1e9);
        desc.seed = m_seed;
        desc.exploration = exploration;
        desc.dump_tree = 0xFFFFFFFFu;
        unsigned int const lanes = m_sharedTree ? lanesFor(m_iterCount) : lanesPerTreeFor(m_iterCount, treesTotal, count);
        desc.flags = (m_sharedTree || lanes > 1) ? (ISMCTS_FLAG_SHARED_TREE | (lanes << ISMCTS_FLAG_LANES_SHIFT)) : 0;
        if (m_materialised > 0 && count == 1 && desc.flags == 0)
            desc.flags |= ISMCTS_FLAG_KEEP_TREES;   // currentTrees() reads trees back after the search

        std::vector<std::uint64_t> visits(count * stride), score2(count * stride), avail(count * stride), iters(count * treesTotal);
        std::vector<std::uint32_t> best(count);
        ismcts_search_out out {};
        out.visits = visits.data();
        out.score2 = score2.data();
        out.avail = avail.data();
        out.iterations_done = iters.data();
        out.best_move = best.data();
        // the trees are sharded over the configured devices by the library (ismcts_search_multi)
        std::vector<ismcts_ctx *> contexts;
        for (auto const &c : m_contexts)
            contexts.push_back(c->get());
        m_contexts.front()->check(ismcts_search_multi(contexts.data(), static_cast<std::uint32_t>(contexts.size()), &desc, podStates, &out));

        m_lastBatch.assign(count, detail::RootStatistics{});
        for (std::size_t p = 0; p < count; ++p) {
            detail::RootStatistics &r = m_lastBatch[p];
            r.visits.assign(visits.begin() + p * stride, visits.begin() + (p + 1) * stride);
            r.score2.assign(score2.begin() + p * stride, score2.begin() + (p + 1) * stride);
            r.available.assign(avail.begin() + p * stride, avail.begin() + (p + 1) * stride);
            r.iterationsPerTree.assign(iters.begin() + p * treesTotal, iters.begin() + (p + 1) * treesTotal);
            r.bestMove = best[p];   // most visits, the first maximum in ascending move order (execution.h:126-135,209-216)
        }
        std::size_t const nDev = std::min<std::size_t>(m_contexts.size(), treesTotal);
        unsigned int const treesOnFirst = treesTotal / nDev + (treesTotal % nDev ? 1 : 0);
        m_lastTreesOnFirstDevice = count == 1 ? treesOnFirst : 0;   // (batches keep no trees)
        return m_lastBatch;
    }

    /// Reads tree `index` of the last search back from the first device (creation order);
    /// MO: the tree of `player`.
    std::vector<ismcts_node_rec> dumpTree(std::uint32_t index, std::uint32_t player = 0) const
    {
        if (m_contexts.empty())
            return {};
        std::uint32_t count = 0;
        std::vector<ismcts_node_rec> nodes(1024);
        int rc = ismcts_dump_tree_of(m_contexts[0]->get(), index, player, nodes.data(), static_cast<std::uint32_t>(nodes.size()), &count);
        if (rc == ISMCTS_ERR_CAPACITY) {
            nodes.resize(count);
            rc = ismcts_dump_tree_of(m_contexts[0]->get(), index, player, nodes.data(), count, &count);
        }
        m_contexts[0]->check(rc);
        nodes.resize(count);
        return nodes;
    }

    unsigned int materialisedTrees() const { return std::min(m_materialised, m_lastTreesOnFirstDevice); }

private:
    std::size_t m_iterCount {0};
    Duration m_iterTime {Duration::zero()};
    unsigned int const m_parallelism;
    bool const m_sharedTree;
    std::uint64_t m_seed {0};
    bool m_seedFixed {false};
    std::vector<int> m_devices {0};
    std::vector<std::unique_ptr<detail::Context>> m_contexts;
    unsigned int m_lanesPerTree {0};
    unsigned int m_materialised {8};
    unsigned int m_lastTreesOnFirstDevice {0};
    detail::RootStatistics m_last;
    std::vector<detail::RootStatistics> m_lastBatch;
};

/// Root parallelisation on the GPU: `numTrees` independent trees (0 = one per hardware thread of the host, as the
/// reference's default, fewer for small budgets), each searched by a block of lanes (setLanesPerTree); the
/// decision sums the root visits per move over all trees (RootParallel, execution.h:181-232).  The throughput
/// geometry of bench.py's `extra.wide` leg is CudaRootParallel{1000000, 2368} with setLanesPerTree(1).
class CudaRootParallel : public CudaExecutionPolicy
{
public:
    explicit CudaRootParallel(std::size_t iterationCount = 1000, unsigned int numTrees = 0)
        : CudaExecutionPolicy{iterationCount, numTrees, false}
    {}
    explicit CudaRootParallel(Duration iterationTime, unsigned int numTrees = 0)
        : CudaExecutionPolicy{iterationTime, numTrees, false}
    {}
};

/// Tree parallelisation on the GPU: one shared tree, `numThreads` iterations in flight with
/// virtual loss (TreeParallel, execution.h:169-179; the reference has no virtual loss).
class CudaTreeParallel : public CudaExecutionPolicy
{
public:
    explicit CudaTreeParallel(std::size_t iterationCount = 1000, unsigned int numThreads = 32)
        : CudaExecutionPolicy{iterationCount, std::max(numThreads, 1u), true}
    {}
    explicit CudaTreeParallel(Duration iterationTime, unsigned int numThreads = 32)
        : CudaExecutionPolicy{iterationTime, std::max(numThreads, 1u), true}
    {}
};

} // ISMCTS

#endif // ISMCTS_CUDA_EXECUTION_H
