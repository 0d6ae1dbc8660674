// TEST / DEMO DRIVER for the drop-in claim: the UNMODIFIED reference partitioner (KaMinPar::compute_partition,
// include/kaminpar-shm/kaminpar.h:857-997) linked with (a) its own factories.cc -> libkaminpar_ref_full.so, or
// (b) the generated factories_b200.cc whose LABEL_PROPAGATION cases return the glue classes of this directory
// -> libkaminpar_ref_b200.so. Same C entry point in both; built by `make -C oracle ref_full ref_b200`.
#include <cstdint>
#include <span>
#include <string_view>

#include "kaminpar-shm/kaminpar.h"

#include "kaminpar-common/environment.h"

namespace kaminpar {
// the three strings CMake generates from kaminpar-common/environment.cc.in
const std::string_view Environment::GIT_SHA1 = "d82640e123ddc71386476c584421a50f0256ac8e";
const std::string_view Environment::GIT_MODIFIED_FILES = "";
const std::string_view Environment::HOSTNAME = "b200-integration";
} // namespace kaminpar

extern "C" {

// Returns the edge cut KaMinPar::compute_partition reports (shm_endtoend_test.cc:142-173 checks it against a
// recomputed cut); partition_out[n]. preset: 0 = default. threads = 1 for the deterministic configuration.
long long kmpfull_compute_partition(std::uint32_t n, const std::uint32_t *xadj, const std::uint32_t *adjncy,
                                    const std::int32_t *vwgt, const std::int32_t *adjwgt, std::uint32_t k,
                                    double epsilon, int seed, int threads, std::uint32_t *partition_out) {
  using namespace kaminpar;
  using namespace kaminpar::shm;
  KaMinPar::reseed(seed);
  KaMinPar partitioner(threads, create_default_context());
  partitioner.set_output_level(OutputLevel::QUIET);
  const std::size_t m = xadj[n];
  partitioner.copy_graph({xadj, static_cast<std::size_t>(n) + 1}, {adjncy, m},
                         vwgt != nullptr ? std::span<const NodeWeight>(vwgt, n) : std::span<const NodeWeight>(),
                         adjwgt != nullptr ? std::span<const EdgeWeight>(adjwgt, m) : std::span<const EdgeWeight>());
  partitioner.set_k(k);
  partitioner.set_uniform_max_block_weights(epsilon);
  return static_cast<long long>(partitioner.compute_partition({partition_out, n}));
}

} // extern "C"
