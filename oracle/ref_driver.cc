// C-ABI driver around the UNMODIFIED KaMinPar reference sources (compiled from where they lie in
// /root/reference against the serial oneTBB stand-in in oracle/ref_shim/). TEST INFRASTRUCTURE
// ONLY: the product never links or loads this. Output: oracle/_ref/libkaminpar_ref.so.
//
// It drives the reference's own objects the way its benchmark apps do
// (apps/benchmarks/shm_label_propagation_benchmark.cc:76-123,
//  apps/benchmarks/shm_refinement_benchmark.cc:125-130):
//   Random::reseed -> [rearrange_by_degree_buckets -> remove_isolated_nodes]
//   -> LPClustering(ctx.coarsening).compute_clustering(...)
//   -> LabelPropagationRefiner(ctx).initialize/refine(p_graph, p_ctx)
// and, for the contraction row, contract_clustering(graph, clustering, con_ctx)
// (coarsening/contraction/cluster_contraction.h:47-56, driven like basic_cluster_coarsener.cc:27-45).
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "kaminpar-shm/coarsening/clustering/lp_clusterer.h"
#include "kaminpar-shm/coarsening/contraction/cluster_contraction.h"
#include "kaminpar-shm/coarsening/max_cluster_weights.h"
#include "kaminpar-shm/datastructures/csr_graph.h"
#include "kaminpar-shm/datastructures/graph.h"
#include "kaminpar-shm/datastructures/partitioned_graph.h"
#include "kaminpar-shm/graphutils/permutator.h"
#include "kaminpar-shm/kaminpar.h"
#include "kaminpar-shm/metrics.h"
#include "kaminpar-shm/refinement/lp/lp_refiner.h"

#include "kaminpar-common/datastructures/static_array.h"
#include "kaminpar-common/random.h"
#include "kaminpar-common/timer.h"

#include "kaminpar-common/environment.h"

// The reference generates environment.cc from environment.cc.in with CMake (configure_file); the
// three string constants are supplied here instead so that no generated file is needed.
namespace kaminpar {
const std::string_view Environment::GIT_SHA1 = "d82640e123ddc71386476c584421a50f0256ac8e";
const std::string_view Environment::GIT_MODIFIED_FILES = "";
const std::string_view Environment::HOSTNAME = "oracle-ref";
} // namespace kaminpar

using namespace kaminpar;
using namespace kaminpar::shm;

namespace {

template <typename T> StaticArray<T> copy_array(const T *src, std::size_t n) {
  StaticArray<T> a(n);
  if (n > 0) {
    std::memcpy(a.data(), src, n * sizeof(T));
  }
  return a;
}

Graph make_graph(
    std::uint32_t n,
    std::uint32_t m,
    const std::uint32_t *xadj,
    const std::uint32_t *adjncy,
    const std::int32_t *vwgt,
    const std::int32_t *adjwgt,
    bool sorted
) {
  StaticArray<EdgeID> nodes = copy_array<EdgeID>(xadj, static_cast<std::size_t>(n) + 1);
  StaticArray<NodeID> edges = copy_array<NodeID>(adjncy, m);
  StaticArray<NodeWeight> nw = vwgt ? copy_array<NodeWeight>(vwgt, n) : StaticArray<NodeWeight>();
  StaticArray<EdgeWeight> ew = adjwgt ? copy_array<EdgeWeight>(adjwgt, m) : StaticArray<EdgeWeight>();
  return Graph(std::make_unique<CSRGraph>(
      std::move(nodes), std::move(edges), std::move(nw), std::move(ew), sorted
  ));
}

} // namespace

#ifdef KMP_SHIM_PARALLEL
#include <omp.h>
#endif

extern "C" {

// number of OpenMP threads the parallel stand-in uses (no effect in the serial build)
void kmpref_set_num_threads(int threads) {
#ifdef KMP_SHIM_PARALLEL
  omp_set_num_threads(threads > 0 ? threads : 1);
#else
  (void)threads;
#endif
}

struct kmpref_lp_params {
  std::uint32_t num_iterations;        // 5
  std::uint32_t large_degree_threshold; // UINT32_MAX
  std::uint32_t max_num_neighbors;      // UINT32_MAX
  std::int32_t impl;                    // 0 single phase, 1 two phase, 2 growing hash tables
  std::int32_t tie_breaking;            // 0 geometric, 1 uniform
  std::int32_t two_hop_strategy;        // enum TwoHopStrategy
  double two_hop_threshold;             // 0.5
  std::int32_t isolated_nodes_strategy; // enum IsolatedNodesClusteringStrategy
};

int kmpref_abi_version() {
  return 1;
}

const char *kmpref_version() {
  static const std::string v = std::to_string(KAMINPAR_VERSION_MAJOR) + "." +
                               std::to_string(KAMINPAR_VERSION_MINOR) + "." +
                               std::to_string(KAMINPAR_VERSION_PATCH);
  return v.c_str();
}

// Reference `rearrange_by_degree_buckets` (graphutils/permutator.cc) followed, optionally, by
// `remove_isolated_nodes` as in kaminpar.cc:369-396. Outputs the permuted CSR (caller-allocated,
// same sizes as the input), old_to_new, and the degree-bucket prefix array (34 entries) as seen
// by the LP after isolated-node removal. Returns the number of nodes the LP sees.
std::uint32_t kmpref_rearrange_by_degree_buckets(
    std::uint32_t n,
    std::uint32_t m,
    const std::uint32_t *xadj,
    const std::uint32_t *adjncy,
    const std::int32_t *vwgt,
    const std::int32_t *adjwgt,
    int remove_isolated,
    std::uint32_t *out_xadj,
    std::uint32_t *out_adjncy,
    std::int32_t *out_vwgt,
    std::int32_t *out_adjwgt,
    std::uint32_t *out_old_to_new,
    std::uint32_t *out_buckets /* 34 */,
    std::uint32_t *out_num_buckets
) {
  Graph graph = make_graph(n, m, xadj, adjncy, vwgt, adjwgt, false);
  CSRGraph &csr = graph.csr_graph();
  graph = graph::rearrange_by_degree_buckets(csr);
  CSRGraph &sorted = graph.csr_graph();

  std::memcpy(out_xadj, sorted.raw_nodes().data(), (static_cast<std::size_t>(n) + 1) * 4);
  std::memcpy(out_adjncy, sorted.raw_edges().data(), static_cast<std::size_t>(m) * 4);
  if (vwgt && out_vwgt) {
    std::memcpy(out_vwgt, sorted.raw_node_weights().data(), static_cast<std::size_t>(n) * 4);
  }
  if (adjwgt && out_adjwgt) {
    std::memcpy(out_adjwgt, sorted.raw_edge_weights().data(), static_cast<std::size_t>(m) * 4);
  }
  if (out_old_to_new) {
    for (std::uint32_t u = 0; u < n; ++u) {
      out_old_to_new[u] = sorted.map_original_node(u);
    }
  }

  if (remove_isolated) {
    const NodeID num_isolated = graph::count_isolated_nodes(graph);
    sorted.remove_isolated_nodes(num_isolated);
  }
  for (std::size_t b = 0; b < 34; ++b) {
    out_buckets[b] = (b < 33) ? sorted.first_node_in_bucket(b) : sorted.first_invalid_node_in_bucket(32);
  }
  *out_num_buckets = static_cast<std::uint32_t>(sorted.number_of_buckets());
  return sorted.n();
}

// Reference LP clusterer at 1 thread. `sorted` = whether the CSR is degree-bucket sorted (then the
// reference computes degree buckets itself, csr_graph.cc:199-248); n is the LP-visible node count
// (isolated nodes already cut off when sorted, as the facade does).
int kmpref_lp_cluster(
    std::uint32_t n,
    std::uint32_t m,
    const std::uint32_t *xadj,
    const std::uint32_t *adjncy,
    const std::int32_t *vwgt,
    const std::int32_t *adjwgt,
    int sorted,
    int seed,
    std::int32_t max_cluster_weight,
    std::uint32_t desired_cluster_count,
    const kmpref_lp_params *params,
    int num_calls, // >1: call compute_clustering repeatedly on the same object (overlay coarsener)
    std::uint32_t *out_clustering /* n * num_calls */
) {
  Random::reseed(seed);
  DISABLE_TIMERS();

  Graph graph = make_graph(n, m, xadj, adjncy, vwgt, adjwgt, sorted != 0);

  Context ctx = create_default_context();
  auto &lp = ctx.coarsening.clustering.lp;
  if (params) {
    lp.num_iterations = params->num_iterations;
    lp.large_degree_threshold = params->large_degree_threshold;
    lp.max_num_neighbors = params->max_num_neighbors;
    lp.impl = static_cast<LabelPropagationImplementation>(params->impl);
    lp.tie_breaking_strategy = static_cast<TieBreakingStrategy>(params->tie_breaking);
    lp.two_hop_strategy = static_cast<TwoHopStrategy>(params->two_hop_strategy);
    lp.two_hop_threshold = params->two_hop_threshold;
    lp.isolated_nodes_strategy =
        static_cast<IsolatedNodesClusteringStrategy>(params->isolated_nodes_strategy);
  }

  LPClustering clusterer(ctx.coarsening);
  clusterer.set_max_cluster_weight(max_cluster_weight);
  clusterer.set_desired_cluster_count(desired_cluster_count);

  for (int call = 0; call < num_calls; ++call) {
    StaticArray<NodeID> clustering(n);
    clusterer.compute_clustering(clustering, graph, false);
    for (std::uint32_t u = 0; u < n; ++u) {
      out_clustering[static_cast<std::size_t>(call) * n + u] = clustering[u];
    }
  }
  return 0;
}

// Reference LP refiner at 1 thread on a given k-way partition with per-block max weights.
int kmpref_lp_refine(
    std::uint32_t n,
    std::uint32_t m,
    const std::uint32_t *xadj,
    const std::uint32_t *adjncy,
    const std::int32_t *vwgt,
    const std::int32_t *adjwgt,
    int sorted,
    int seed,
    std::uint32_t k,
    const std::int32_t *max_block_weights,
    const std::int32_t *min_block_weights, // nullable
    const kmpref_lp_params *params,
    std::uint32_t *partition_inout,
    std::int32_t *out_block_weights /* k, nullable */
) {
  Random::reseed(seed);
  DISABLE_TIMERS();

  Graph graph = make_graph(n, m, xadj, adjncy, vwgt, adjwgt, sorted != 0);

  Context ctx = create_default_context();
  if (params) {
    auto &lp = ctx.refinement.lp;
    lp.num_iterations = params->num_iterations;
    lp.large_degree_threshold = params->large_degree_threshold;
    lp.max_num_neighbors = params->max_num_neighbors;
    lp.impl = static_cast<LabelPropagationImplementation>(params->impl);
    lp.tie_breaking_strategy = static_cast<TieBreakingStrategy>(params->tie_breaking);
  }

  std::vector<BlockWeight> max_bw(max_block_weights, max_block_weights + k);
  ctx.partition.setup(graph, std::move(max_bw), false);
  if (min_block_weights) {
    ctx.partition.setup_min_block_weights(
        std::vector<BlockWeight>(min_block_weights, min_block_weights + k)
    );
  }

  StaticArray<BlockID> partition = copy_array<BlockID>(partition_inout, n);
  PartitionedGraph p_graph(graph, k, std::move(partition));

  LabelPropagationRefiner refiner(ctx);
  refiner.initialize(p_graph);
  refiner.refine(p_graph, ctx.partition);

  for (std::uint32_t u = 0; u < n; ++u) {
    partition_inout[u] = p_graph.block(u);
  }
  if (out_block_weights) {
    for (std::uint32_t b = 0; b < k; ++b) {
      out_block_weights[b] = p_graph.block_weight(b);
    }
  }
  return 0;
}

std::int64_t kmpref_edge_cut(
    std::uint32_t n,
    std::uint32_t m,
    const std::uint32_t *xadj,
    const std::uint32_t *adjncy,
    const std::int32_t *vwgt,
    const std::int32_t *adjwgt,
    std::uint32_t k,
    const std::uint32_t *partition
) {
  Graph graph = make_graph(n, m, xadj, adjncy, vwgt, adjwgt, false);
  StaticArray<BlockID> part = copy_array<BlockID>(partition, n);
  PartitionedGraph p_graph(graph, k, std::move(part));
  return metrics::edge_cut(p_graph);
}

// compute_max_cluster_weight (coarsening/max_cluster_weights.h:17-46) with the default context and
// PartitionContext::setup(graph, k, eps) (context.cc:27-39).
std::int32_t kmpref_max_cluster_weight(
    std::uint32_t n, std::uint32_t m, const std::uint32_t *xadj, const std::uint32_t *adjncy,
    const std::int32_t *vwgt, std::uint32_t k, double epsilon
) {
  Graph graph = make_graph(n, m, xadj, adjncy, vwgt, nullptr, false);
  Context ctx = create_default_context();
  ctx.partition.setup(graph, k, epsilon);
  return compute_max_cluster_weight<NodeWeight>(
      ctx.coarsening, ctx.partition, graph.n(), graph.total_node_weight()
  );
}

void kmpref_max_block_weights(
    std::uint32_t n, std::uint32_t m, const std::uint32_t *xadj, const std::uint32_t *adjncy,
    const std::int32_t *vwgt, std::uint32_t k, double epsilon, std::int32_t *out /* k */
) {
  Graph graph = make_graph(n, m, xadj, adjncy, vwgt, nullptr, false);
  Context ctx = create_default_context();
  ctx.partition.setup(graph, k, epsilon);
  for (std::uint32_t b = 0; b < k; ++b) {
    out[b] = ctx.partition.max_block_weight(b);
  }
}

// contract_clustering(graph, clustering, con_ctx) (cluster_contraction.cc:22-50). algorithm: 0 BUFFERED,
// 1 UNBUFFERED (default preset, presets.cc:181-183), 2 UNBUFFERED_NAIVE. Outputs are caller-allocated
// upper bounds: c_xadj[n+1], c_adjncy[m], c_vwgt[n], c_adjwgt[m], mapping[n] (fine -> coarse, via
// CoarseGraph::project_up of the identity). Returns c_n; *c_m_out = directed coarse edges.
std::uint32_t kmpref_contract(
    std::uint32_t n, std::uint32_t m, const std::uint32_t *xadj, const std::uint32_t *adjncy,
    const std::int32_t *vwgt, const std::int32_t *adjwgt, const std::uint32_t *clustering, int algorithm,
    std::uint32_t *c_xadj, std::uint32_t *c_adjncy, std::int32_t *c_vwgt, std::int32_t *c_adjwgt,
    std::uint32_t *mapping, std::uint32_t *c_m_out
) {
  Graph graph = make_graph(n, m, xadj, adjncy, vwgt, adjwgt, false);
  Context ctx = create_default_context();
  ContractionCoarseningContext con_ctx = ctx.coarsening.contraction;
  con_ctx.algorithm = algorithm == 0   ? ContractionAlgorithm::BUFFERED
                      : algorithm == 1 ? ContractionAlgorithm::UNBUFFERED
                                       : ContractionAlgorithm::UNBUFFERED_NAIVE;
  StaticArray<NodeID> cl = copy_array<NodeID>(clustering, n);
  std::unique_ptr<CoarseGraph> coarse = contract_clustering(graph, std::move(cl), con_ctx);
  const CSRGraph &cg = coarse->get().csr_graph();
  const std::uint32_t c_n = cg.n();
  const std::uint32_t c_m = cg.m();
  for (std::uint32_t u = 0; u <= c_n; ++u) {
    c_xadj[u] = cg.raw_nodes()[u];
  }
  for (std::uint32_t u = 0; u < c_n; ++u) {
    c_vwgt[u] = cg.node_weight(u);
  }
  for (std::uint32_t e = 0; e < c_m; ++e) {
    c_adjncy[e] = cg.raw_edges()[e];
    c_adjwgt[e] = cg.edge_weight(e);
  }
  std::vector<BlockID> ident(c_n);
  for (std::uint32_t u = 0; u < c_n; ++u) {
    ident[u] = u;
  }
  std::vector<BlockID> fine(n);
  coarse->project_up(ident, fine);
  for (std::uint32_t u = 0; u < n; ++u) {
    mapping[u] = fine[u];
  }
  *c_m_out = c_m;
  return c_n;
}

} // extern "C"
