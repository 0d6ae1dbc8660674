// Device-side building blocks of the LP sweep: hashes, candidate ordering, warp/block arg-max.
//
// The selection rule is the reference's (lp_clusterer.cc:181-250, lp_refiner.cc:151-245) restated as
// an order-independent reduction: among feasible candidates maximise
//   clusterer: (rating desc, tie_hash asc, cluster id asc)
//   refiner:   (rating desc, overload asc, tie_hash asc, block id asc)
// where tie_hash is a counter-based hash replacing the reference's "uniform draw among the ties in
// rating-map insertion order" (SURVEY.md §7 hard parts).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace kmp {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kLadderLevels = 16;

enum : uint32_t { SALT_SUBROUND = 1, SALT_TIE = 2, SALT_FAV = 3, SALT_COMMIT = 4 };

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t tie_hash(uint32_t base, uint32_t u, uint32_t c) {
  return lowbias32((u * 0x9E3779B1u) ^ (c * 0x85EBCA77u) ^ base);
}
// permutation of [0, 2^32): unique commit priorities
__host__ __device__ __forceinline__ uint32_t bijective32(uint32_t x, uint32_t base) {
  x ^= base;
  x *= 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  x *= 0xC2B2AE3Du;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t subround_of(uint32_t u, uint32_t granule_log2,
                                                         uint32_t base, uint32_t S) {
  const uint32_t h = lowbias32((u >> granule_log2) ^ base);
  return static_cast<uint32_t>((static_cast<uint64_t>(h) * S) >> 32);
}
__device__ __forceinline__ uint32_t ladder_level(uint32_t prio) {
  const uint32_t z = static_cast<uint32_t>(__clz(static_cast<int>(prio))); // __clz(0) == 32
  return z < static_cast<uint32_t>(kLadderLevels - 1) ? z : static_cast<uint32_t>(kLadderLevels - 1);
}
// degree group of the sync schedule: degree buckets {<=3} {4,5} {6..8} {>=9}
// (bucket(d) = floor(log2 d) + 1, kaminpar-common/degree_buckets.h:23-25)
__host__ __device__ __forceinline__ uint32_t degree_group(uint32_t d) {
  return d < 8 ? 0u : (d < 32 ? 1u : (d < 256 ? 2u : 3u));
}

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline uint32_t sync_base(int seed, uint32_t call, uint32_t iter, uint32_t salt) {
  const uint64_t a = splitmix64((static_cast<uint64_t>(static_cast<uint32_t>(seed)) << 32) | call);
  const uint64_t b = splitmix64((static_cast<uint64_t>(iter) << 32) | salt);
  return static_cast<uint32_t>(splitmix64(a ^ b) >> 32);
}

// ---- candidates ---------------------------------------------------------------------------------
struct Cand {
  int32_t gain;  // -1: none
  int32_t over;  // refiner only
  uint32_t hash;
  uint32_t key;
};
__device__ __forceinline__ Cand cand_none() { return Cand{-1, 0, kEmpty, kEmpty}; }

template <int MODE> __device__ __forceinline__ bool cand_better(const Cand &a, const Cand &b) {
  if (a.gain != b.gain) {
    return a.gain > b.gain;
  }
  if (MODE == 1 && a.over != b.over) {
    return a.over < b.over;
  }
  if (a.hash != b.hash) {
    return a.hash < b.hash;
  }
  return a.key < b.key;
}

// arg-max over the lanes in `mask` (all of them must call with the same mask). Uses redux.sync.
template <int MODE> __device__ __forceinline__ Cand warp_argmax(unsigned mask, Cand c) {
  const int gmax = __reduce_max_sync(mask, c.gain);
  bool in = (c.gain == gmax) && (gmax > 0);
  int omin = 0;
  if (MODE == 1) {
    omin = __reduce_min_sync(mask, in ? c.over : INT32_MAX);
    in = in && (c.over == omin);
  }
  const unsigned hmin = __reduce_min_sync(mask, in ? c.hash : kEmpty);
  in = in && (c.hash == hmin);
  const unsigned kmin = __reduce_min_sync(mask, in ? c.key : kEmpty);
  Cand r;
  r.gain = gmax > 0 ? gmax : -1;
  r.over = omin;
  r.hash = hmin;
  r.key = kmin;
  return r;
}

// ---- packed gather array: neighbour label + "moved" stamp in ONE random access -------------------
// The sweep reads label[v] for every edge. Packing a stamp of v's last move next to the label lets the
// sweep derive the reference's active flag ("a neighbour moved since my last visit",
// label_propagation.h:848-870) from the gather it does anyway, instead of a second pass over the
// adjacency of every moved vertex that scatters active[v] = 1 (pull instead of push).
//   P64 = false: 32-bit word, label in bits 0..23 (n <= 2^24), stamp in bits 24..31
//   P64 = true : 64-bit word, label in bits 0..31, stamp in bits 32..63
// Stamp = 0: not moved in the current or the previous round; otherwise 1 + (round parity << kStampBits)
// + sub-round index of the move. A vertex u visited in sub-round q of round r has to see the moves of
// round r (all of them happened before its visit: sub-rounds < q) and the moves of round r-1 in
// sub-rounds >= q (after its previous visit). Stamps of round r-2 are cleared at the start of round r.
constexpr uint32_t kStampBits = 6;   // sub-round index < 64 (4 degree groups x S <= 16 sub-rounds)
constexpr uint32_t kMaxStampSubrounds = 1u << kStampBits;
template <bool P64> struct LabG;
template <> struct LabG<false> {
  using word = uint32_t;
  static __host__ __device__ __forceinline__ uint32_t label(word w) { return w & 0x00FFFFFFu; }
  static __host__ __device__ __forceinline__ uint32_t stamp(word w) { return w >> 24; }
  static __host__ __device__ __forceinline__ word pack(uint32_t label, uint32_t stamp) { return label | (stamp << 24); }
};
template <> struct LabG<true> {
  using word = unsigned long long;
  static __host__ __device__ __forceinline__ uint32_t label(word w) { return static_cast<uint32_t>(w); }
  static __host__ __device__ __forceinline__ uint32_t stamp(word w) { return static_cast<uint32_t>(w >> 32); }
  static __host__ __device__ __forceinline__ word pack(uint32_t label, uint32_t stamp) {
    return static_cast<word>(label) | (static_cast<word>(stamp) << 32);
  }
};
__host__ __device__ __forceinline__ uint32_t make_stamp(uint32_t round, uint32_t subround) {
  return 1u + ((round & 1u) << kStampBits) + subround;
}
// window test of the visit in (round parity `par`, sub-round `q`): win_start = ((par ^ 1) << kStampBits) + q,
// win_len = 2^(kStampBits+1) - q; the codes of the window are cyclically contiguous
struct StampWindow {
  uint32_t start, len;
};
__host__ __device__ __forceinline__ StampWindow make_window(uint32_t round, uint32_t subround) {
  return StampWindow{(((round & 1u) ^ 1u) << kStampBits) + subround, (2u << kStampBits) - subround};
}
__device__ __forceinline__ bool stamp_hit(uint32_t stamp, const StampWindow &w) {
  return stamp != 0 && ((stamp - 1u - w.start) & ((2u << kStampBits) - 1u)) < w.len;
}

// ---- per-vertex context -------------------------------------------------------------------------
constexpr int kCounterNodesOffset = 16; // scan counters: edges at [tier], visited vertices at [16 + tier]
struct SweepArgs {
  // graph
  const uint32_t *__restrict__ xadj;
  const uint32_t *__restrict__ adjncy;
  const int32_t *__restrict__ vwgt;   // nullable
  const int32_t *__restrict__ adjwgt; // nullable
  // state
  const uint32_t *__restrict__ label;  // frozen during the sweep
  const void *__restrict__ labg;       // packed (label, stamp) gather array, LabG<P64>::word[n]
  StampWindow window;                  // which stamps count as "moved since my last visit"
  bool pull;                           // derive the active flag from the stamps (else: active[] only)
  uint32_t *__restrict__ queue;        // work-queue cursor of this launch (team kernels), zeroed per round
  const int32_t *__restrict__ weight;  // cluster weights [n] / block weights [k], frozen
  const int32_t *__restrict__ max_w;   // refiner: per block; clusterer: nullptr
  const int32_t *__restrict__ min_w;   // refiner: nullable
  const uint32_t *__restrict__ communities; // nullable
  uint8_t *__restrict__ active;        // nullable in select_all mode
  uint32_t *__restrict__ favored;      // clusterer
  int32_t max_cluster_weight;          // clusterer
  uint32_t num_labels;                 // n (clusterer) or k (refiner)
  uint32_t max_num_neighbors;
  // work list of this sub-round
  const uint32_t *__restrict__ list;
  uint32_t list_size;
  // hashing
  uint32_t base_tie, base_fav, base_commit;
  // outputs
  uint32_t *__restrict__ mv_u;
  uint32_t *__restrict__ mv_t;
  uint32_t *__restrict__ mover_count;
  bool accumulate;                     // add to incoming[] / hist[] while emitting proposals
  int32_t *__restrict__ incoming;      // clusterer: [n]
  int32_t *__restrict__ hist;          // refiner: [k][16]
  unsigned long long *__restrict__ counters; // [0] edges scanned, [kCounterNodesOffset] nodes visited (of this kernel tier)
  // select_all mode (T0 parity hook): write decisions instead of proposing
  uint32_t *__restrict__ sel_target;
  uint32_t *__restrict__ sel_favored;
};

// Evaluate one (key, rating) candidate of vertex u. Returns the "best" candidate; fills `fav`.
// kw = weight[key] (the caller may have batched the gather)
template <int MODE>
__device__ __forceinline__ Cand eval_candidate_w(const SweepArgs &a, uint32_t u, uint32_t own, int32_t uw,
                                                 int32_t own_w, uint32_t key, int32_t rating, int32_t kw,
                                                 bool store_fav, Cand &fav) {
  Cand c = cand_none();
  fav = cand_none();
  if (rating <= 0) {
    return c;
  }
  if (MODE == 0) {
    bool feasible = (kw + uw <= a.max_cluster_weight) || (key == own);
    if (a.communities != nullptr) {
      feasible = feasible && (a.communities[key] == a.communities[own]);
    }
    if (feasible) {
      c.gain = rating;
      c.over = 0;
      c.hash = tie_hash(a.base_tie, u, key);
      c.key = key;
    }
    if (store_fav) {
      fav.gain = rating;
      fav.over = 0;
      fav.hash = tie_hash(a.base_fav, u, key);
      fav.key = key;
    }
  } else {
    const int32_t kmax = a.max_w[key];
    const int32_t over = kw - kmax;
    const int32_t init_over = own_w - a.max_w[own];
    const bool feasible = (kw + uw <= kmax) || (over < init_over) || (key == own);
    if (feasible) {
      c.gain = rating;
      c.over = over;
      c.hash = tie_hash(a.base_tie, u, key);
      c.key = key;
    }
  }
  return c;
}

template <int MODE>
__device__ __forceinline__ Cand eval_candidate(const SweepArgs &a, uint32_t u, uint32_t own, int32_t uw,
                                               int32_t own_w, uint32_t key, int32_t rating, bool store_fav,
                                               Cand &fav) {
  if (rating <= 0) {
    fav = cand_none();
    return cand_none();
  }
  return eval_candidate_w<MODE>(a, u, own, uw, own_w, key, rating, a.weight[key], store_fav, fav);
}

// Final per-vertex action (one thread): propose a move or store the favored cluster.
template <int MODE>
__device__ __forceinline__ bool finish_vertex(const SweepArgs &a, uint32_t u, uint32_t own, bool store_fav,
                                              const Cand &best, const Cand &fav, uint32_t &target_out) {
  const uint32_t target = best.gain > 0 ? best.key : own;
  target_out = target;
  if (a.sel_target != nullptr) { // select_all mode
    a.sel_target[u] = target;
    if (MODE == 0 && a.sel_favored != nullptr) {
      a.sel_favored[u] = store_fav ? (fav.gain > 0 ? fav.key : own) : kEmpty;
    }
    return false;
  }
  if (target != own) {
    return true;
  }
  if (MODE == 0 && store_fav) {
    a.favored[u] = fav.gain > 0 ? fav.key : own;
  }
  return false;
}

} // namespace kmp
