// TEST INFRASTRUCTURE ONLY (never linked into libkaminpar_b200.so).
//
// Compiles kaminpar_b200/csrc/lp_sortnet.cuh -- the register sorting network of the sweep_thread<N> kernels -- with
// g++ and checks it on the host:
//   * 0-1 principle: a comparator network sorts every input iff it sorts every 0/1 input. Exhaustive for N = 8 and
//     N = 16 (2^16 inputs), 2^20 random 0/1 inputs plus every "threshold" input for N = 32 and 64;
//   * the weighted variant moves each weight with its key: the multiset of (key, weight) pairs is preserved and
//     the per-key weight sums (= the ratings the kernel reads off the run lengths) are unchanged;
//   * the compare-exchange counts are Batcher's (19 / 63 / 191 / 543).
//   g++ -O2 -std=c++17 -shared -fPIC -o libsortnet_host_check.so sortnet_host_check.cc
#include <algorithm>
#include <cstdint>
#include <map>
#include <random>
#include <vector>

#include "../../kaminpar_b200/csrc/lp_sortnet.cuh"

namespace {

template <int N> bool sorted_after(uint64_t bits) {
  uint32_t k[N];
  int32_t w[N];
  int ones = 0;
  for (int i = 0; i < N; ++i) {
    k[i] = (bits >> i) & 1u ? 0xFFFFFFFFu : 7u; // "empty slot" vs a label
    w[i] = 1;
    ones += static_cast<int>((bits >> i) & 1u);
  }
  kmp::sort_registers<N, false>(k, w);
  for (int i = 0; i < N; ++i) {
    if (k[i] != (i < N - ones ? 7u : 0xFFFFFFFFu)) {
      return false;
    }
  }
  return true;
}

template <int N> int zero_one(bool exhaustive, uint32_t seed) {
  if (exhaustive) {
    for (uint64_t b = 0; b < (1ull << N); ++b) {
      if (!sorted_after<N>(b)) {
        return 1;
      }
    }
    return 0;
  }
  std::mt19937_64 gen(seed);
  for (int t = 0; t < (1 << 20); ++t) {
    // mix densities: AND / OR of draws gives sparse and dense patterns as well
    uint64_t b = gen();
    if (t % 3 == 1) {
      b &= gen();
    } else if (t % 3 == 2) {
      b |= gen();
    }
    if (N < 64) {
      b &= (1ull << N) - 1;
    }
    if (!sorted_after<N>(b)) {
      return 1;
    }
  }
  for (int ones = 0; ones <= N; ++ones) { // all rotations of every threshold pattern
    for (int rot = 0; rot < N; ++rot) {
      uint64_t b = 0;
      for (int i = 0; i < ones; ++i) {
        b |= 1ull << ((i + rot) % N);
      }
      if (!sorted_after<N>(b)) {
        return 1;
      }
    }
  }
  return 0;
}

template <int N> int weighted(uint32_t seed, int trials) {
  std::mt19937 gen(seed);
  for (int t = 0; t < trials; ++t) {
    uint32_t k[N], k2[N];
    int32_t w[N], w2[N];
    std::map<uint32_t, long long> sum;
    std::vector<std::pair<uint32_t, int32_t>> before, after;
    const uint32_t labels = 1 + gen() % (t % 5 == 0 ? 2 : N);
    const uint32_t deg = gen() % (N + 1);
    for (int i = 0; i < N; ++i) {
      const bool used = static_cast<uint32_t>(i) < deg;
      k[i] = used ? gen() % labels + (t % 7 == 0 ? 0xFFFFFF00u : 0u) : 0xFFFFFFFFu;
      w[i] = used ? static_cast<int32_t>(1 + gen() % 1000) : 0;
      k2[i] = k[i];
      w2[i] = w[i];
      sum[k[i]] += w[i];
      before.emplace_back(k[i], w[i]);
    }
    kmp::sort_registers<N, true>(k, w);
    kmp::sort_registers<N, false>(k2, w2);
    std::map<uint32_t, long long> sum_after;
    for (int i = 0; i < N; ++i) {
      if (i > 0 && k[i - 1] > k[i]) {
        return 2;
      }
      if (k2[i] != k[i]) {
        return 3; // keys-only variant must give the same key order
      }
      sum_after[k[i]] += w[i];
      after.emplace_back(k[i], w[i]);
    }
    std::sort(before.begin(), before.end());
    std::sort(after.begin(), after.end());
    if (before != after || sum != sum_after) {
      return 4;
    }
  }
  return 0;
}

} // namespace

extern "C" {

int sortnet_pairs(int n) {
  switch (n) {
  case 8: return kmp::SortNetworkOf<8>::net.count;
  case 16: return kmp::SortNetworkOf<16>::net.count;
  case 32: return kmp::SortNetworkOf<32>::net.count;
  case 64: return kmp::SortNetworkOf<64>::net.count;
  default: return -1;
  }
}

// 0 = every checked 0/1 input comes out sorted
int sortnet_zero_one(int n, unsigned seed) {
  switch (n) {
  case 8: return zero_one<8>(true, seed);
  case 16: return zero_one<16>(true, seed);
  case 32: return zero_one<32>(false, seed);
  case 64: return zero_one<64>(false, seed);
  default: return -1;
  }
}

// 0 = sorted, same key order as the keys-only network, (key, weight) multiset and per-key sums preserved
int sortnet_weighted(int n, unsigned seed, int trials) {
  switch (n) {
  case 8: return weighted<8>(seed, trials);
  case 16: return weighted<16>(seed, trials);
  case 32: return weighted<32>(seed, trials);
  case 64: return weighted<64>(seed, trials);
  default: return -1;
  }
}
}
