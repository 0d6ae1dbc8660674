// C++ host side above the C ABI: adapter classes with the method names, argument meaning and
// error behaviour of the reference's plugin interfaces, so that
//   kaminpar-shm/factories.cc:66-67   case ClusteringAlgorithm::LABEL_PROPAGATION  -> b200::LPClustering
//   kaminpar-shm/factories.cc:108-109 case RefinementAlgorithm::LABEL_PROPAGATION  -> b200::LabelPropagationRefiner
// become one-line swaps (INTEGRATION.md shows the glue that maps kaminpar::shm::Graph /
// PartitionedGraph / PartitionContext onto the views below).
//
// Mirrors:
//   class Clusterer  kaminpar-shm/coarsening/clusterer.h:19-47
//   class Refiner    kaminpar-shm/refinement/refiner.h:18-57
//   LPClustering     kaminpar-shm/coarsening/clustering/lp_clusterer.h:19 / lp_clusterer.cc:376-399
//   LabelPropagationRefiner  kaminpar-shm/refinement/lp/lp_refiner.h:19 / lp_refiner.cc:357-376
//   CoarseGraph / contract_clustering  kaminpar-shm/coarsening/contraction/cluster_contraction.h:22-56
//
// Error convention: the reference's path has no error codes (KASSERT aborts); here a non-zero
// status of the C ABI becomes std::runtime_error. There is no CPU fallback.
#pragma once

#include <cstdint>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

#include "kaminpar_b200_contraction.h"
#include "kaminpar_b200_lp.h"

namespace kaminpar_b200 {

using NodeID = std::uint32_t;     // include/kaminpar-shm/kaminpar.h:32-57 (default build)
using EdgeID = std::uint32_t;
using BlockID = std::uint32_t;
using NodeWeight = std::int32_t;
using EdgeWeight = std::int32_t;
using BlockWeight = std::int32_t;

// Borrowed view of a CSRGraph (csr_graph.h:35-482). Empty weight spans mean unit weights.
struct CSRGraphView {
  std::span<const EdgeID> nodes;           // raw_nodes(),   n + 1
  std::span<const NodeID> edges;           // raw_edges(),   m
  std::span<const NodeWeight> node_weights; // raw_node_weights(), n or empty
  std::span<const EdgeWeight> edge_weights; // raw_edge_weights(), m or empty
  bool sorted = false;                      // CSRGraph::sorted(); read by the seq_strict schedule only
  [[nodiscard]] NodeID n() const { return nodes.empty() ? 0 : static_cast<NodeID>(nodes.size() - 1); }
  [[nodiscard]] EdgeID m() const { return static_cast<EdgeID>(edges.size()); }
  [[nodiscard]] const void *identity() const { return nodes.data(); }
};

// The parts of PartitionedGraph (partitioned_graph.h:50-456) the refiner reads and writes.
struct PartitionedGraphView {
  CSRGraphView graph;
  BlockID k = 0;
  std::span<BlockID> partition;          // raw_partition(), refined in place
  std::span<BlockWeight> block_weights;  // k entries, updated to match the refined partition
};

// The parts of PartitionContext (kaminpar.h:417-531) the refiner reads.
struct PartitionContextView {
  BlockID k = 0;
  std::span<const BlockWeight> max_block_weights; // max_block_weight(b)
  std::span<const BlockWeight> min_block_weights; // min_block_weight(b); empty -> 0
};

struct LabelPropagationCoarseningContext { // kaminpar.h:140-154, defaults presets.cc:140-153
  std::size_t num_iterations = 5;
  NodeID large_degree_threshold = 0xFFFFFFFFu;
  NodeID max_num_neighbors = 0xFFFFFFFFu;
  int impl = KMP_LP_TWO_PHASE;
  bool relabel_before_second_phase = false;
  int two_hop_strategy = KMP_TWO_HOP_MATCH_THREADWISE;
  double two_hop_threshold = 0.5;
  int isolated_nodes_strategy = KMP_ISOLATED_MATCH_DURING_TWO_HOP;
  int tie_breaking_strategy = KMP_TIE_UNIFORM;
};

struct LabelPropagationRefinementContext { // kaminpar.h:221-228, defaults presets.cc:339-347
  std::size_t num_iterations = 5;
  NodeID large_degree_threshold = 0xFFFFFFFFu;
  NodeID max_num_neighbors = 0xFFFFFFFFu;
  int impl = KMP_LP_SINGLE_PHASE;
  int tie_breaking_strategy = KMP_TIE_UNIFORM;
};

struct EngineContext { // engine knobs without a reference counterpart
  int seed = 0;        // Random::reseed analogue
  unsigned sync_subrounds = 8;
  unsigned sync_granule_log2 = 4;
  int device = -1;
  int schedule = KMP_SCHEDULE_SYNC; // KMP_SCHEDULE_SEQ_STRICT: the reference's one-thread order, bit-identical, small inputs
};

namespace detail {
inline void check(int rc) {
  if (rc != KMP_OK) {
    throw std::runtime_error(std::string("kaminpar_b200: ") + kmp_last_error());
  }
}
class Handle {
public:
  explicit Handle(const kmp_lp_config &cfg) {
    // struct layouts (kmp_lp_config / kmp_lp_stats) are part of the ABI: refuse a library built from another header
    if (kmp_lp_abi_version() != KMP_LP_ABI_VERSION) {
      throw std::runtime_error("kaminpar_b200: library ABI version " + std::to_string(kmp_lp_abi_version()) +
                               " != header ABI version " + std::to_string(KMP_LP_ABI_VERSION));
    }
    check(kmp_lp_create(&cfg, &_h));
  }
  Handle(const Handle &) = delete;
  Handle &operator=(const Handle &) = delete;
  ~Handle() { kmp_lp_destroy(_h); }
  void set_graph(const CSRGraphView &g) {
    if (g.identity() == _graph_id && g.n() == _n && g.m() == _m) {
      return; // same borrowed graph as in the previous call
    }
    check(kmp_lp_set_graph(_h, g.n(), g.m(), g.nodes.data(), g.edges.data(),
                           g.node_weights.empty() ? nullptr : g.node_weights.data(),
                           g.edge_weights.empty() ? nullptr : g.edge_weights.data()));
    if (g.sorted) {
      check(kmp_lp_set_graph_sorted(_h, 1));
    }
    _graph_id = g.identity();
    _n = g.n();
    _m = g.m();
  }
  [[nodiscard]] kmp_lp_handle *get() const { return _h; }
  // The borrowed graph is recognised by (address, n, m): call this when its arrays were rewritten in place (or
  // freed and reallocated at the same address) so that the next call uploads it again.
  void invalidate_graph() { _graph_id = nullptr; }

private:
  kmp_lp_handle *_h = nullptr;
  const void *_graph_id = nullptr;
  NodeID _n = 0;
  EdgeID _m = 0;
};
} // namespace detail

// Same surface as kaminpar::shm::Clusterer (clusterer.h:35-46).
class LPClustering {
public:
  explicit LPClustering(const LabelPropagationCoarseningContext &lp_ctx, const EngineContext &engine = {})
      : _handle(make_config(lp_ctx, engine)) {}

  void set_max_cluster_weight(const NodeWeight weight) { _max_cluster_weight = weight; }
  void set_desired_cluster_count(const NodeID count) { _desired = count; }
  void set_communities(std::span<const NodeID> communities) { _communities = communities; }

  // clustering: caller-allocated, size graph.n(), fully overwritten with ids of cluster leaders in
  // [0, n), not compacted (basic_cluster_coarsener.cc:29, cluster_contraction_preprocessing.cc:17-51).
  void compute_clustering(std::span<NodeID> clustering, const CSRGraphView &graph, const bool free_memory_afterwards) {
    _handle.set_graph(graph);
    detail::check(kmp_lp_cluster(_handle.get(), _max_cluster_weight, _desired,
                                 _communities.empty() ? nullptr : _communities.data(), clustering.data(), &_stats));
    if (free_memory_afterwards) { // lp_clusterer.cc:330-333
      detail::check(kmp_lp_free_scratch(_handle.get()));
    }
  }
  [[nodiscard]] const kmp_lp_stats &last_stats() const { return _stats; }
  [[nodiscard]] kmp_lp_handle *handle() const { return _handle.get(); } // graph holder for contract_clustering
  void invalidate_graph() { _handle.invalidate_graph(); }                // the borrowed graph changed in place

private:
  static kmp_lp_config make_config(const LabelPropagationCoarseningContext &c, const EngineContext &e) {
    kmp_lp_config cfg;
    kmp_lp_default_config(0, &cfg);
    cfg.num_iterations = static_cast<std::uint32_t>(c.num_iterations);
    cfg.large_degree_threshold = c.large_degree_threshold;
    cfg.max_num_neighbors = c.max_num_neighbors;
    cfg.impl = c.impl;
    cfg.relabel_before_second_phase = c.relabel_before_second_phase;
    cfg.two_hop_strategy = c.two_hop_strategy;
    cfg.two_hop_threshold = c.two_hop_threshold;
    cfg.isolated_nodes_strategy = c.isolated_nodes_strategy;
    cfg.tie_breaking_strategy = c.tie_breaking_strategy;
    cfg.seed = e.seed;
    cfg.sync_subrounds = e.sync_subrounds;
    cfg.sync_granule_log2 = e.sync_granule_log2;
    cfg.device = e.device;
    cfg.schedule = e.schedule;
    return cfg;
  }
  detail::Handle _handle;
  NodeWeight _max_cluster_weight = 0;
  NodeID _desired = 0;
  std::span<const NodeID> _communities;
  kmp_lp_stats _stats{};
};

// Same surface as kaminpar::shm::Refiner (refiner.h:34-56).
class LabelPropagationRefiner {
public:
  explicit LabelPropagationRefiner(const LabelPropagationRefinementContext &lp_ctx, const EngineContext &engine = {})
      : _handle(make_config(lp_ctx, engine)) {}

  [[nodiscard]] std::string name() const { return "Label Propagation"; }
  void set_communities(std::span<const NodeID> communities) { _communities = communities; }

  void initialize(const PartitionedGraphView &p_graph) { _handle.set_graph(p_graph.graph); }

  // Mutates p_graph in place (labels and block weights stay consistent,
  // partitioned_graph.h:117-135); always returns true like lp_refiner.cc:88.
  bool refine(PartitionedGraphView &p_graph, const PartitionContextView &p_ctx) {
    if (p_graph.k > p_ctx.k || p_ctx.max_block_weights.size() != p_ctx.k) {
      throw std::invalid_argument("kaminpar_b200: inconsistent k / max_block_weights");
    }
    _handle.set_graph(p_graph.graph);
    detail::check(kmp_lp_refine(_handle.get(), p_ctx.k, p_ctx.max_block_weights.data(),
                                p_ctx.min_block_weights.empty() ? nullptr : p_ctx.min_block_weights.data(),
                                _communities.empty() ? nullptr : _communities.data(), p_graph.partition.data(),
                                p_graph.block_weights.empty() ? nullptr : p_graph.block_weights.data(), &_stats));
    return true;
  }
  [[nodiscard]] const kmp_lp_stats &last_stats() const { return _stats; }
  [[nodiscard]] kmp_lp_handle *handle() const { return _handle.get(); } // e.g. for kmp_lp_dist_init
  void invalidate_graph() { _handle.invalidate_graph(); }

private:
  static kmp_lp_config make_config(const LabelPropagationRefinementContext &c, const EngineContext &e) {
    kmp_lp_config cfg;
    kmp_lp_default_config(1, &cfg);
    cfg.num_iterations = static_cast<std::uint32_t>(c.num_iterations);
    cfg.large_degree_threshold = c.large_degree_threshold;
    cfg.max_num_neighbors = c.max_num_neighbors;
    cfg.impl = c.impl;
    cfg.tie_breaking_strategy = c.tie_breaking_strategy;
    cfg.seed = e.seed;
    cfg.sync_subrounds = e.sync_subrounds;
    cfg.sync_granule_log2 = e.sync_granule_log2;
    cfg.device = e.device;
    cfg.schedule = e.schedule;
    return cfg;
  }
  detail::Handle _handle;
  std::span<const NodeID> _communities;
  kmp_lp_stats _stats{};
};

// Same surface as kaminpar::shm::CoarseGraph (cluster_contraction.h:22-32). The coarse graph lives on the
// device; get() copies it into host vectors once (a maintainer's glue wraps them into a CSRGraph).
class CoarseGraph {
public:
  struct HostCSR {
    std::vector<EdgeID> nodes;
    std::vector<NodeID> edges;
    std::vector<NodeWeight> node_weights;
    std::vector<EdgeWeight> edge_weights;
  };
  explicit CoarseGraph(kmp_coarse_graph *g, const kmp_contraction_stats &stats) : _g(g), _stats(stats) {}
  CoarseGraph(const CoarseGraph &) = delete;
  CoarseGraph &operator=(const CoarseGraph &) = delete;
  ~CoarseGraph() { kmp_coarse_destroy(_g); }

  [[nodiscard]] NodeID n() const { return kmp_coarse_n(_g); }
  [[nodiscard]] EdgeID m() const { return kmp_coarse_m(_g); }
  const HostCSR &get() {
    if (_host.nodes.empty()) {
      _host.nodes.resize(static_cast<std::size_t>(n()) + 1);
      _host.edges.resize(m());
      _host.node_weights.resize(n());
      _host.edge_weights.resize(m());
      detail::check(kmp_coarse_download(_g, _host.nodes.data(), _host.edges.data(), _host.node_weights.data(),
                                        _host.edge_weights.data(), nullptr));
    }
    return _host;
  }
  // fine[u] = coarse[mapping[u]] (cluster_contraction_preprocessing.h:36-40)
  void project_up(std::span<const BlockID> coarse, std::span<BlockID> fine) const {
    if (coarse.size() != n() || fine.size() != kmp_coarse_fine_n(_g)) {
      throw std::invalid_argument("project_up: wrong span size");
    }
    detail::check(kmp_coarse_project_up(_g, coarse.data(), fine.data()));
  }
  // coarse[mapping[u]] = fine[u] (:42-46)
  void project_down(std::span<const BlockID> fine, std::span<BlockID> coarse) const {
    if (coarse.size() != n() || fine.size() != kmp_coarse_fine_n(_g)) {
      throw std::invalid_argument("project_down: wrong span size");
    }
    detail::check(kmp_coarse_project_down(_g, fine.data(), coarse.data()));
  }
  [[nodiscard]] const kmp_contraction_stats &stats() const { return _stats; }
  [[nodiscard]] const kmp_coarse_graph *device() const { return _g; } // kmp_coarse_device_arrays for the next level

private:
  kmp_coarse_graph *_g;
  kmp_contraction_stats _stats;
  HostCSR _host;
};

// contract_clustering(graph, clustering, con_ctx) (cluster_contraction.h:47-50). `clusterer` is the LP
// clusterer that already holds `graph` on the device (no second H2D copy); an empty `clustering` span
// contracts by the clustering its last compute_clustering() left on the device.
inline std::unique_ptr<CoarseGraph> contract_clustering(kmp_lp_handle *graph_holder, std::span<const NodeID> clustering) {
  kmp_coarse_graph *g = nullptr;
  kmp_contraction_stats stats{};
  detail::check(kmp_contract_clustering(graph_holder, clustering.empty() ? nullptr : clustering.data(), &g, &stats));
  return std::make_unique<CoarseGraph>(g, stats);
}

} // namespace kaminpar_b200
