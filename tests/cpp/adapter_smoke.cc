// Compiles against the C++ adapters + C ABI; used by tests to check that the header is valid C++20
// and that the library links. Running it needs a GPU (tests/test_gpu_cpp_adapter.py).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kaminpar_b200_adapters.hpp"

using namespace kaminpar_b200;

int main() {
  // 4x4 grid
  const int R = 4, C = 4;
  std::vector<EdgeID> xadj{0};
  std::vector<NodeID> adj;
  for (int r = 0; r < R; ++r) {
    for (int c = 0; c < C; ++c) {
      if (r > 0) adj.push_back((r - 1) * C + c);
      if (c > 0) adj.push_back(r * C + c - 1);
      if (c + 1 < C) adj.push_back(r * C + c + 1);
      if (r + 1 < R) adj.push_back((r + 1) * C + c);
      xadj.push_back(static_cast<EdgeID>(adj.size()));
    }
  }
  CSRGraphView g{xadj, adj, {}, {}};
  try {
    LPClustering clusterer(LabelPropagationCoarseningContext{});
    clusterer.set_max_cluster_weight(4);
    clusterer.set_desired_cluster_count(0);
    std::vector<NodeID> clustering(g.n());
    clusterer.compute_clustering(clustering, g, false);
    for (NodeID c : clustering) {
      if (c >= g.n()) return 2;
    }
    // contraction by the clustering that is still on the device, then projections
    auto coarse = contract_clustering(clusterer.handle(), {});
    NodeWeight total = 0;
    for (NodeWeight w : coarse->get().node_weights) total += w;
    if (total != static_cast<NodeWeight>(g.n()) || coarse->get().nodes.size() != coarse->n() + 1u) return 5;
    std::vector<BlockID> cpart(coarse->n()), fine(g.n()), back(coarse->n());
    for (NodeID c = 0; c < coarse->n(); ++c) cpart[c] = c % 3;
    coarse->project_up(cpart, fine);
    coarse->project_down(fine, back);
    if (back != cpart) return 6;
    std::vector<BlockID> part(g.n());
    for (NodeID u = 0; u < g.n(); ++u) part[u] = u % 2;
    std::vector<BlockWeight> bw(2), maxw{9, 9};
    PartitionedGraphView pg{g, 2, part, bw};
    PartitionContextView pc{2, maxw, {}};
    LabelPropagationRefiner refiner(LabelPropagationRefinementContext{});
    refiner.initialize(pg);
    if (!refiner.refine(pg, pc)) return 3;
    if (bw[0] + bw[1] != static_cast<BlockWeight>(g.n()) || bw[0] > 9 || bw[1] > 9) return 4;
    std::printf("adapter ok: clusters via C++ adapters, block weights %d/%d\n", bw[0], bw[1]);
    // parity hook: ADAPTER_DUMP=<file> gets "n, xadj, adjncy, clustering, partition, block weights" as text so that
    // tests/test_cpp_adapter.py can compare the adapter's results with the oracle's sync schedule bit for bit
    if (const char *path = std::getenv("ADAPTER_DUMP")) {
      if (std::FILE *f = std::fopen(path, "w")) {
        std::fprintf(f, "%u %u\n", g.n(), g.m());
        for (EdgeID x : xadj) std::fprintf(f, "%u ", x);
        std::fprintf(f, "\n");
        for (NodeID v : adj) std::fprintf(f, "%u ", v);
        std::fprintf(f, "\n");
        for (NodeID c : clustering) std::fprintf(f, "%u ", c);
        std::fprintf(f, "\n");
        for (BlockID b : part) std::fprintf(f, "%u ", b);
        std::fprintf(f, "\n%d %d\n", bw[0], bw[1]);
        std::fclose(f);
      }
    }
  } catch (const std::exception &e) {
    std::printf("exception: %s\n", e.what());
    return 1;
  }
  return 0;
}
