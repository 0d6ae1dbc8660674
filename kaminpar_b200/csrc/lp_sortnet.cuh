// Register sorting network of the thread-per-vertex sweep kernels (lp_sweep.cuh: sweep_thread<N>).
//
// Batcher's odd-even merge sort as a fixed compare-exchange network on N = 8, 16, 32 (or 64) values held in
// registers: 19 / 63 / 191 / 543 compare-exchanges, no data-dependent control flow, so 32 vertices of different
// degree sort in lockstep in one warp (unused slots hold 0xFFFFFFFF and sort to the end). The compare-exchange list
// is computed at compile time and applied through index sequences so that every array index is a constant
// expression -- a run-time loop nest makes ptxas put the arrays into local memory.
//
// Plain integer code marked KMP_SORT_HD so that tests/cpp/sortnet_host_check.cc can compile the same source with
// g++ and check it (0-1 principle) on a box without a GPU; the product only runs it inside sweep_thread.
#pragma once

#include <cstdint>
#include <utility>

#if defined(__CUDACC__)
#define KMP_SORT_HD __host__ __device__ __forceinline__
#else
#define KMP_SORT_HD inline
#endif

namespace kmp {

template <bool EW>
KMP_SORT_HD void compare_exchange(uint32_t &ka, uint32_t &kb, int32_t &wa, int32_t &wb) {
  if (EW) {
    const bool sw = ka > kb;
    const uint32_t k0 = sw ? kb : ka, k1 = sw ? ka : kb;
    const int32_t w0 = sw ? wb : wa, w1 = sw ? wa : wb;
    ka = k0;
    kb = k1;
    wa = w0;
    wb = w1;
  } else {
    const uint32_t lo = ka < kb ? ka : kb, hi = ka < kb ? kb : ka;
    ka = lo;
    kb = hi;
  }
}
// The compare-exchange list of Batcher's odd-even merge sort, computed at compile time so that every register
// index below is a constant (a run-time loop nest would push the arrays into local memory).
template <int N> struct SortNetwork {
  static constexpr int kMaxPairs = 576; // N = 64 needs 543
  int count = 0;
  unsigned char lo[kMaxPairs] = {}, hi[kMaxPairs] = {};
};
template <int N> constexpr SortNetwork<N> make_sort_network() {
  SortNetwork<N> net{};
  for (int p = 1; p < N; p <<= 1) {
    for (int q = p; q >= 1; q >>= 1) {
      for (int j = q % p; j + q < N; j += 2 * q) {
        for (int i = 0; i < q; ++i) {
          if (i + j + q < N && (i + j) / (2 * p) == (i + j + q) / (2 * p)) {
            net.lo[net.count] = static_cast<unsigned char>(i + j);
            net.hi[net.count] = static_cast<unsigned char>(i + j + q);
            ++net.count;
          }
        }
      }
    }
  }
  return net;
}
template <int N> struct SortNetworkOf {
  static constexpr SortNetwork<N> net = make_sort_network<N>();
};
template <int N, bool EW, int IDX>
KMP_SORT_HD void sort_step(uint32_t (&k)[N], int32_t (&w)[N]) {
  if constexpr (IDX < SortNetworkOf<N>::net.count) {
    constexpr int A = SortNetworkOf<N>::net.lo[IDX];
    constexpr int B = SortNetworkOf<N>::net.hi[IDX];
    compare_exchange<EW>(k[A], k[B], w[A], w[B]);
  }
}
template <int N, bool EW, int BASE, int... I>
KMP_SORT_HD void sort_steps(uint32_t (&k)[N], int32_t (&w)[N], std::integer_sequence<int, I...>) {
  (sort_step<N, EW, BASE + I>(k, w), ...);
}
template <int N, bool EW, int... C>
KMP_SORT_HD void sort_chunks(uint32_t (&k)[N], int32_t (&w)[N], std::integer_sequence<int, C...>) {
  (sort_steps<N, EW, C * 64>(k, w, std::make_integer_sequence<int, 64>{}), ...);
}
// ascending by key; the weights travel with their keys (EW only)
template <int N, bool EW> KMP_SORT_HD void sort_registers(uint32_t (&k)[N], int32_t (&w)[N]) {
  static_assert((N & (N - 1)) == 0 && N <= 64, "network size must be a power of two <= 64");
  sort_chunks<N, EW>(k, w, std::make_integer_sequence<int, (SortNetworkOf<N>::net.count + 63) / 64>{});
}

} // namespace kmp
