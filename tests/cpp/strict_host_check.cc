// TEST INFRASTRUCTURE ONLY (never linked into libkaminpar_b200.so).
//
// Compiles kaminpar_b200/csrc/lp_strict.cuh -- the source of the KMP_SCHEDULE_SEQ_STRICT kernel -- with g++
// so that (1) its restatements of libstdc++'s mt19937 / uniform_int_distribution / std::shuffle can be
// pinned against the real std:: facilities, and (2) the sequential engine can be checked against the
// reference goldens on a box without a GPU. The product runs the same code inside strict_kernel only.
//   g++ -O2 -std=c++17 -shared -fPIC -o libstrict_host_check.so strict_host_check.cc
#include <algorithm>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include "../../kaminpar_b200/csrc/lp_strict.cuh"

using namespace kmp_strict;

extern "C" {

// returns 0 if the restated generator / distributions agree with libstdc++ on `trials` mixed draws
int strict_check_rng(int seed, int trials) {
  Rng *r = new Rng();
  mt_seed(*r, static_cast<uint32_t>(seed));
  std::mt19937 gen(seed);
  int bad = 0;
  for (int t = 0; t < trials; ++t) {
    // ranges: small, medium, near 2^32
    const uint32_t ranges[6] = {2u, 3u, 64u, 4032u + static_cast<uint32_t>(t % 977), 1u << 20, 0xFFFFFFF0u - static_cast<uint32_t>(t)};
    for (uint32_t range : ranges) {
      const uint32_t mine = uniform_below(*r, range);
      const std::size_t theirs = std::uniform_int_distribution<std::size_t>(0, static_cast<std::size_t>(range) - 1)(gen);
      bad += mine != theirs;
    }
    const int b = std::uniform_int_distribution<int>(0, 1)(gen);
    bad += static_cast<int>(uniform_below(*r, 2)) != b;
    // shuffles of several sizes incl. even / odd / above the two-at-a-time limit (65535)
    const std::size_t sizes[7] = {1, 2, 5, 64, 1000, 65535, 70001};
    const std::size_t sz = sizes[t % 7];
    std::vector<uint32_t> a(sz), b2(sz);
    std::iota(a.begin(), a.end(), 0u);
    std::iota(b2.begin(), b2.end(), 0u);
    shuffle(*r, a.data(), sz);
    std::shuffle(b2.begin(), b2.end(), gen);
    bad += a != b2;
    // pairs (chunks are {start, end})
    std::vector<std::pair<uint32_t, uint32_t>> pc(37);
    std::vector<uint32_t> flat(74);
    for (uint32_t i = 0; i < 37; ++i) {
      pc[i] = {i, 1000 + i};
      flat[2 * i] = i;
      flat[2 * i + 1] = 1000 + i;
    }
    shuffle(*r, flat.data(), 37, 2);
    std::shuffle(pc.begin(), pc.end(), gen);
    for (uint32_t i = 0; i < 37; ++i) {
      bad += flat[2 * i] != pc[i].first || flat[2 * i + 1] != pc[i].second;
    }
  }
  delete r;
  return bad;
}

// Random::reseed + RandomPermutations as the reference builds them (random.cc:45-56, random.h:138-143)
int strict_check_rng_init(int seed) {
  Rng *r = new Rng();
  rng_init(*r, seed);
  std::mt19937 gen(seed);
  std::uniform_int_distribution<int> bd(0, 1);
  int bad = 0;
  for (uint32_t i = 0; i < kBools; ++i) {
    bad += (bd(gen) != 0) != (r->bools[i] != 0);
  }
  for (uint32_t p = 0; p < kNumPerms; ++p) {
    std::vector<uint32_t> v(kPermSize);
    std::iota(v.begin(), v.end(), 0u);
    std::shuffle(v.begin(), v.end(), gen);
    for (uint32_t i = 0; i < kPermSize; ++i) {
      bad += v[i] != r->perms[p][i];
    }
  }
  bad += random_index(*r, 0, 64) != std::uniform_int_distribution<std::size_t>(0, 63)(gen);
  delete r;
  return bad;
}

struct HostRun {
  std::vector<int32_t> weight, slot, ent_val, slot2, ent2_val, concurrent;
  std::vector<uint32_t> favored, ent_key, ent2_key, used, second, tie_best, tie_fav, chunks, sub_perm, match_map, buckets;
  std::vector<uint8_t> active;
};

// num_calls consecutive compute_clustering calls on one object (one Random, one set of permutations)
int strict_host_cluster(uint32_t n, uint32_t m, const uint32_t *xadj, const uint32_t *adjncy, const int32_t *vwgt,
                        const int32_t *adjwgt, int sorted, int seed, int32_t max_cluster_weight, uint32_t desired,
                        const uint32_t *communities, uint32_t num_iterations, uint32_t large_degree_threshold,
                        uint32_t max_num_neighbors, int impl, int tie_uniform, int two_hop_strategy,
                        double two_hop_threshold, int isolated_nodes_strategy, int num_calls, uint32_t *out,
                        Stats *stats_out) {
  Rng *rng = new Rng();
  rng_init(*rng, seed);
  const std::size_t keys = std::max<uint32_t>(n, 1);
  HostRun h;
  h.weight.resize(keys); h.slot.resize(keys); h.ent_val.resize(keys + 1); h.slot2.resize(keys); h.ent2_val.resize(keys + 1);
  h.concurrent.resize(keys); h.favored.resize(keys); h.ent_key.resize(keys + 1); h.ent2_key.resize(keys + 1); h.used.resize(keys);
  h.second.resize(keys); h.tie_best.resize(keys + 1); h.tie_fav.resize(keys + 1); h.chunks.resize(2 * (keys + 64));
  h.sub_perm.resize(keys / 64 + 2); h.match_map.resize(keys); h.buckets.resize(36); h.active.resize(keys);
  for (int call = 0; call < num_calls; ++call) {
    Args a{};
    a.n = n; a.m = m; a.xadj = xadj; a.adjncy = adjncy; a.vwgt = vwgt; a.adjwgt = adjwgt; a.sorted = sorted;
    a.buckets = h.buckets.data();
    a.num_iterations = num_iterations; a.large_degree_threshold = large_degree_threshold; a.max_num_neighbors = max_num_neighbors;
    a.impl = impl; a.tie_uniform = tie_uniform; a.two_hop_strategy = two_hop_strategy; a.isolated_nodes_strategy = isolated_nodes_strategy;
    a.two_hop_threshold = two_hop_threshold; a.mode = 0; a.max_cluster_weight = max_cluster_weight; a.desired_num_clusters = desired;
    a.k = 0; a.max_bw = nullptr; a.min_bw = nullptr; a.communities = communities;
    a.label = out + static_cast<std::size_t>(call) * n; a.weight = h.weight.data(); a.favored = h.favored.data(); a.active = h.active.data();
    a.slot = h.slot.data(); a.ent_key = h.ent_key.data(); a.ent_val = h.ent_val.data(); a.slot2 = h.slot2.data();
    a.ent2_key = h.ent2_key.data(); a.ent2_val = h.ent2_val.data(); a.concurrent = h.concurrent.data(); a.used_entries = h.used.data();
    a.second_phase_nodes = h.second.data(); a.tie_best = h.tie_best.data(); a.tie_fav = h.tie_fav.data(); a.chunks = h.chunks.data();
    a.sub_perm = h.sub_perm.data(); a.match_map = h.match_map.data(); a.rng = rng; a.stats = &stats_out[call];
    Engine e(a);
    e.run();
  }
  delete rng;
  return 0;
}

int strict_host_refine(uint32_t n, uint32_t m, const uint32_t *xadj, const uint32_t *adjncy, const int32_t *vwgt,
                       const int32_t *adjwgt, int sorted, int seed, uint32_t k, const int32_t *max_bw, const int32_t *min_bw,
                       const uint32_t *communities, uint32_t num_iterations, uint32_t large_degree_threshold,
                       uint32_t max_num_neighbors, int impl, int tie_uniform, uint32_t *partition, int32_t *bw_out,
                       Stats *stats_out) {
  Rng *rng = new Rng();
  rng_init(*rng, seed);
  const std::size_t keys = std::max<uint32_t>(std::max(n, k), 1);
  HostRun h;
  h.weight.resize(keys); h.slot.resize(keys); h.ent_val.resize(keys + 1); h.slot2.resize(keys); h.ent2_val.resize(keys + 1);
  h.concurrent.resize(keys); h.favored.resize(keys); h.ent_key.resize(keys + 1); h.ent2_key.resize(keys + 1); h.used.resize(keys);
  h.second.resize(keys); h.tie_best.resize(keys + 1); h.tie_fav.resize(keys + 1); h.chunks.resize(2 * (keys + 64));
  h.sub_perm.resize(keys / 64 + 2); h.match_map.resize(keys); h.buckets.resize(36); h.active.resize(keys);
  Args a{};
  a.n = n; a.m = m; a.xadj = xadj; a.adjncy = adjncy; a.vwgt = vwgt; a.adjwgt = adjwgt; a.sorted = sorted;
  a.buckets = h.buckets.data();
  a.num_iterations = num_iterations; a.large_degree_threshold = large_degree_threshold; a.max_num_neighbors = max_num_neighbors;
  a.impl = impl; a.tie_uniform = tie_uniform; a.two_hop_strategy = 0; a.isolated_nodes_strategy = 0; a.two_hop_threshold = 0.5;
  a.mode = 1; a.max_cluster_weight = 0; a.desired_num_clusters = 0; a.k = k; a.max_bw = max_bw; a.min_bw = min_bw; a.communities = communities;
  a.label = partition; a.weight = h.weight.data(); a.favored = h.favored.data(); a.active = h.active.data();
  a.slot = h.slot.data(); a.ent_key = h.ent_key.data(); a.ent_val = h.ent_val.data(); a.slot2 = h.slot2.data();
  a.ent2_key = h.ent2_key.data(); a.ent2_val = h.ent2_val.data(); a.concurrent = h.concurrent.data(); a.used_entries = h.used.data();
  a.second_phase_nodes = h.second.data(); a.tie_best = h.tie_best.data(); a.tie_fav = h.tie_fav.data(); a.chunks = h.chunks.data();
  a.sub_perm = h.sub_perm.data(); a.match_map = h.match_map.data(); a.rng = rng; a.stats = stats_out;
  Engine e(a);
  e.run();
  std::memcpy(bw_out, h.weight.data(), k * sizeof(int32_t));
  delete rng;
  return 0;
}

} // extern "C"
