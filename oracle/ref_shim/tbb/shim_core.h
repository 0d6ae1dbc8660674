// Serial stand-in for the subset of oneTBB that KaMinPar's shared-memory code uses.
//
// TEST INFRASTRUCTURE ONLY. This is *our* code, not a copy of oneTBB: it exists so that the
// UNMODIFIED reference sources under /root/reference (which hard-depend on oneTBB v2022.2.0,
// absent from this image) can be compiled into oracle/_ref/ and run with exactly one worker
// thread. Every "parallel" construct executes sequentially, in ascending index order, on the
// calling thread; this_task_arena::current_thread_index() is pinned to 0 and max_concurrency()
// to 1. That is the configuration the reference's own tests call deterministic
// (tests/endtoend/shm_endtoend_test.cc:152 "1 thread: deterministic").
//
// Assumption flagged in DESIGN.md: real oneTBB may report a different slot index for a thread
// that calls into the library outside an arena; the reference only uses that index to offset the
// RNG seed (kaminpar-common/random.cc:45-50), and 0 is the value for the master thread of the
// arena in which KaMinPar runs its algorithms.
#pragma once

// Two build modes:
//   default              : serial (deterministic; used to pin the oracle, oracle/_ref/libkaminpar_ref.so)
//   -DKMP_SHIM_PARALLEL  : parallel_for / enumerable_thread_specific / concurrent_vector backed by OpenMP
//                          threads, so that the reference's LP runs on all host cores for the CPU
//                          baseline (oracle/_ref/libkaminpar_ref_omp.so). Like with real oneTBB the
//                          result then depends on thread timing.
#ifdef KMP_SHIM_PARALLEL
#include <omp.h>
#endif

#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <functional>
#include <iterator>
#include <memory>
#include <mutex>
#include <type_traits>
#include <utility>
#include <array>
#include <vector>

namespace tbb {

struct split {};

template <typename Value> class blocked_range {
public:
  using const_iterator = Value;
  using size_type = std::size_t;

  blocked_range() = default;
  blocked_range(Value b, Value e, size_type grain = 1) : _b(b), _e(e), _grain(grain) {}
  blocked_range(blocked_range &r, split) : _b(r._e), _e(r._e), _grain(r._grain) {}

  const_iterator begin() const { return _b; }
  const_iterator end() const { return _e; }
  size_type size() const { return static_cast<size_type>(_e - _b); }
  size_type grainsize() const { return _grain; }
  bool empty() const { return !(_b < _e); }
  bool is_divisible() const { return false; }

private:
  Value _b{};
  Value _e{};
  size_type _grain = 1;
};

struct auto_partitioner {};
struct simple_partitioner {};
struct static_partitioner {};

// ---- parallel_for ---------------------------------------------------------------------------
namespace shim_detail {
#ifdef KMP_SHIM_PARALLEL
inline bool in_parallel() { return omp_in_parallel() != 0; }
inline int num_threads() { return omp_get_max_threads(); }
inline int thread_index() { return omp_get_thread_num(); }
#else
inline bool in_parallel() { return true; }
inline int num_threads() { return 1; }
inline int thread_index() { return 0; }
#endif
} // namespace shim_detail

template <typename Value, typename Body> void parallel_for(const blocked_range<Value> &range, const Body &body) {
  if (range.empty()) {
    return;
  }
#ifdef KMP_SHIM_PARALLEL
  if constexpr (std::is_integral_v<Value>) {
    if (!shim_detail::in_parallel() && shim_detail::num_threads() > 1) {
      const std::size_t n = range.size();
      const std::size_t target = (n + 8 * shim_detail::num_threads() - 1) / (8 * shim_detail::num_threads());
      const std::size_t grain = std::max<std::size_t>(std::max<std::size_t>(range.grainsize(), 1), target);
      const std::size_t chunks = (n + grain - 1) / grain;
#pragma omp parallel for schedule(dynamic, 1)
      for (std::size_t c = 0; c < chunks; ++c) {
        const Value b = static_cast<Value>(range.begin() + c * grain);
        const Value e = static_cast<Value>(std::min<std::size_t>(n, (c + 1) * grain) + range.begin());
        body(blocked_range<Value>(b, e, range.grainsize()));
      }
      return;
    }
  }
#endif
  body(range);
}

// any other range type (e.g. enumerable_thread_specific::range_type): one sequential call; the
// blocked_range overload above is more specialised and wins for blocked ranges
template <typename Range, typename Body,
          typename = decltype(std::declval<const Range &>().begin())>
void parallel_for(const Range &range, const Body &body) {
  if (!range.empty()) {
    body(range);
  }
}
template <typename Range, typename Body, typename Partitioner,
          typename = decltype(std::declval<const Range &>().begin()),
          typename = std::enable_if_t<std::is_empty_v<Partitioner>>>
void parallel_for(const Range &range, const Body &body, const Partitioner &) {
  if (!range.empty()) {
    body(range);
  }
}
template <typename Index, typename Function,
          typename = std::enable_if_t<std::is_integral_v<Index>>>
void parallel_for(Index first, Index last, const Function &f) {
#ifdef KMP_SHIM_PARALLEL
  if (!shim_detail::in_parallel() && shim_detail::num_threads() > 1 && first < last) {
    const long long n = static_cast<long long>(last) - static_cast<long long>(first);
    const long long chunk = std::max<long long>(1, n / (16LL * shim_detail::num_threads()));
#pragma omp parallel for schedule(dynamic, chunk)
    for (long long i = 0; i < n; ++i) {
      f(static_cast<Index>(first + i));
    }
    return;
  }
#endif
  for (Index i = first; i < last; ++i) {
    f(i);
  }
}
template <typename Index, typename Function,
          typename = std::enable_if_t<std::is_integral_v<Index>>>
void parallel_for(Index first, Index last, Index step, const Function &f) {
  for (Index i = first; i < last; i += step) {
    f(i);
  }
}

// ---- parallel_invoke ------------------------------------------------------------------------
template <typename... Fs> void parallel_invoke(Fs &&...fs) {
  (std::forward<Fs>(fs)(), ...);
}

// ---- parallel_reduce ------------------------------------------------------------------------
template <typename Range, typename Value, typename RealBody, typename Reduction>
Value parallel_reduce(
    const Range &range, const Value &identity, const RealBody &real_body, const Reduction &
) {
  if (range.empty()) {
    return identity;
  }
  return real_body(range, identity);
}
template <typename Range, typename Body> void parallel_reduce(const Range &range, Body &body) {
  if (!range.empty()) {
    body(range);
  }
}

// ---- parallel_scan --------------------------------------------------------------------------
struct pre_scan_tag {
  static bool is_final_scan() { return false; }
  operator bool() const { return false; }
};
struct final_scan_tag {
  static bool is_final_scan() { return true; }
  operator bool() const { return true; }
};
template <typename Range, typename Body> void parallel_scan(const Range &range, Body &body) {
  if (!range.empty()) {
    body(range, final_scan_tag{});
  }
}
template <typename Range, typename Value, typename Scan, typename ReverseJoin>
Value parallel_scan(const Range &range, const Value &identity, const Scan &scan, const ReverseJoin &) {
  if (range.empty()) {
    return identity;
  }
  return scan(range, identity, true);
}

// ---- arena ----------------------------------------------------------------------------------
namespace this_task_arena {
inline int max_concurrency() { return shim_detail::num_threads(); }
inline int current_thread_index() { return shim_detail::thread_index(); }
template <typename F> auto isolate(F &&f) { return f(); }
} // namespace this_task_arena

class task_arena {
public:
  static constexpr int automatic = -1;
  static constexpr int not_initialized = -2;
  task_arena(int = automatic, unsigned = 1) {}
  void initialize() {}
  void initialize(int, unsigned = 1) {}
  void terminate() {}
  int max_concurrency() const { return 1; }
  template <typename F> auto execute(F &&f) { return f(); }
  template <typename F> void enqueue(F &&f) { f(); }
};

class task_scheduler_observer {
public:
  task_scheduler_observer() = default;
  explicit task_scheduler_observer(task_arena &) {}
  virtual ~task_scheduler_observer() = default;
  void observe(bool = true) {}
  virtual void on_scheduler_entry(bool) {}
  virtual void on_scheduler_exit(bool) {}
};

class global_control {
public:
  enum parameter { max_allowed_parallelism, thread_stack_size, terminate_on_exception };
  global_control(parameter, std::size_t) {}
  static std::size_t active_value(parameter) { return 1; }
};

class task_group {
public:
  template <typename F> void run(F &&f) { f(); }
  template <typename F> void run_and_wait(F &&f) { f(); }
  void wait() {}
  void cancel() {}
};

class spin_mutex {
public:
  class scoped_lock {
  public:
    scoped_lock() = default;
    explicit scoped_lock(spin_mutex &) {}
    void acquire(spin_mutex &) {}
    bool try_acquire(spin_mutex &) { return true; }
    void release() {}
  };
  void lock() {}
  bool try_lock() { return true; }
  void unlock() {}
};

// ---- allocators -----------------------------------------------------------------------------
template <typename T> using cache_aligned_allocator = std::allocator<T>;
template <typename T> using scalable_allocator = std::allocator<T>;
template <typename T> using tbb_allocator = std::allocator<T>;

// ---- enumerable_thread_specific (exactly one slot, created lazily) ---------------------------
enum ets_key_usage_type { ets_key_per_instance, ets_no_key, ets_suspend_aware };

template <typename T, typename Allocator = std::allocator<T>, ets_key_usage_type = ets_no_key>
class enumerable_thread_specific {
public:
  using value_type = T;
  using reference = T &;
  using const_reference = const T &;
  using iterator = typename std::vector<std::unique_ptr<T>>::iterator;

  // Iterator that dereferences to T& (std::vector<unique_ptr<T>> keeps addresses stable).
  template <bool Const> class iter {
    using base = std::conditional_t<Const, typename std::vector<std::unique_ptr<T>>::const_iterator,
                                    typename std::vector<std::unique_ptr<T>>::iterator>;

  public:
    using iterator_category = std::random_access_iterator_tag;
    using value_type = T;
    using difference_type = std::ptrdiff_t;
    using pointer = std::conditional_t<Const, const T *, T *>;
    using reference = std::conditional_t<Const, const T &, T &>;
    iter() = default;
    explicit iter(base it) : _it(it) {}
    reference operator*() const { return **_it; }
    pointer operator->() const { return _it->get(); }
    iter &operator++() { ++_it; return *this; }
    iter operator++(int) { iter t = *this; ++_it; return t; }
    iter &operator--() { --_it; return *this; }
    iter &operator+=(difference_type d) { _it += d; return *this; }
    iter operator+(difference_type d) const { return iter(_it + d); }
    difference_type operator-(const iter &o) const { return _it - o._it; }
    bool operator==(const iter &o) const { return _it == o._it; }
    bool operator!=(const iter &o) const { return _it != o._it; }
    bool operator<(const iter &o) const { return _it < o._it; }

  private:
    base _it{};
  };
  using it_type = iter<false>;
  using const_it_type = iter<true>;

  class range_type {
  public:
    range_type(it_type b, it_type e) : _b(b), _e(e) {}
    it_type begin() const { return _b; }
    it_type end() const { return _e; }
    bool empty() const { return _b == _e; }
    std::size_t size() const { return static_cast<std::size_t>(_e - _b); }
    bool is_divisible() const { return false; }

  private:
    it_type _b, _e;
  };

  enumerable_thread_specific() : _init([] { return std::make_unique<T>(); }) {}

  template <typename Finit,
            typename = std::enable_if_t<std::is_invocable_r_v<T, Finit> &&
                                        !std::is_same_v<std::decay_t<Finit>, enumerable_thread_specific>>>
  explicit enumerable_thread_specific(Finit finit)
      : _init([finit]() mutable { return std::unique_ptr<T>(new T(finit())); }) {}

  template <typename U = T, typename = std::enable_if_t<std::is_copy_constructible_v<U>>>
  explicit enumerable_thread_specific(const T &exemplar)
      : _init([exemplar] { return std::make_unique<T>(exemplar); }) {}

  template <typename A0, typename A1, typename... Args>
  enumerable_thread_specific(A0 &&a0, A1 &&a1, Args &&...args)
      : _init([=] { return std::make_unique<T>(a0, a1, args...); }) {}

  enumerable_thread_specific(enumerable_thread_specific &&o) noexcept
      : _init(std::move(o._init)), _slots(std::move(o._slots)), _by_thread(o._by_thread) {
    o._by_thread.fill(nullptr);
  }
  enumerable_thread_specific &operator=(enumerable_thread_specific &&o) noexcept {
    if (this != &o) {
      _init = std::move(o._init);
      _slots = std::move(o._slots);
      _by_thread = o._by_thread;
      o._by_thread.fill(nullptr);
    }
    return *this;
  }
  enumerable_thread_specific(const enumerable_thread_specific &o) : _init(o._init) {  // slots copied, thread map rebuilt lazily
    for (const auto &p : o._slots) {
      if constexpr (std::is_copy_constructible_v<T>) {
        _slots.push_back(std::make_unique<T>(*p));
      }
    }
  }
  enumerable_thread_specific &operator=(const enumerable_thread_specific &o) {
    if (this != &o) {
      enumerable_thread_specific tmp(o);
      *this = std::move(tmp);
    }
    return *this;
  }

  reference local() {
#ifdef KMP_SHIM_PARALLEL
    // the slot table has a fixed size and exists from construction on: every thread of a team calls
    // local() at the same moment when a parallel loop starts, and a lazily (re)allocated table raced
    const int t = shim_detail::thread_index();
    if (static_cast<std::size_t>(t) >= kMaxSlots) {
      std::abort();
    }
    T *p = _by_thread[static_cast<std::size_t>(t)];
    if (p == nullptr) {
      std::unique_ptr<T> fresh = _init();
      p = fresh.get();
#pragma omp critical(kmp_shim_ets)
      { _slots.push_back(std::move(fresh)); }
      _by_thread[static_cast<std::size_t>(t)] = p;
    }
    return *p;
#else
    if (_slots.empty()) {
      _slots.push_back(_init());
    }
    return *_slots.front();
#endif
  }
  reference local(bool &exists) {
#ifdef KMP_SHIM_PARALLEL
    const int t = shim_detail::thread_index();
    exists = static_cast<std::size_t>(t) < kMaxSlots && _by_thread[static_cast<std::size_t>(t)] != nullptr;
#else
    exists = !_slots.empty();
#endif
    return local();
  }

  std::size_t size() const { return _slots.size(); }
  bool empty() const { return _slots.empty(); }
  void clear() {
    _slots.clear();
    _by_thread.fill(nullptr);
  }

  it_type begin() { return it_type(_slots.begin()); }
  it_type end() { return it_type(_slots.end()); }
  const_it_type begin() const { return const_it_type(_slots.begin()); }
  const_it_type end() const { return const_it_type(_slots.end()); }

  range_type range(std::size_t = 1) { return range_type(begin(), end()); }

  template <typename BinaryOp> T combine(BinaryOp op) {
    if (_slots.empty()) {
      return *_init();
    }
    auto it = _slots.begin();
    T acc = **it;
    for (++it; it != _slots.end(); ++it) {
      acc = op(acc, **it);
    }
    return acc;
  }
  template <typename UnaryOp> void combine_each(UnaryOp op) {
    for (auto &p : _slots) {
      op(*p);
    }
  }

private:
  std::function<std::unique_ptr<T>()> _init;
  std::vector<std::unique_ptr<T>> _slots;
  static constexpr std::size_t kMaxSlots = 1024;
  std::array<T *, kMaxSlots> _by_thread{}; // parallel mode: slot of each OpenMP thread (never reallocated)
};

// ---- combinable -----------------------------------------------------------------------------
template <typename T> class combinable {
public:
  combinable() : _init([] { return T(); }) {}
  template <typename Finit> explicit combinable(Finit finit) : _init(finit) {}
  T &local() {
    if (!_value) {
      _value = std::make_unique<T>(_init());
    }
    return *_value;
  }
  T &local(bool &exists) {
    exists = static_cast<bool>(_value);
    return local();
  }
  void clear() { _value.reset(); }
  template <typename BinaryOp> T combine(BinaryOp) { return _value ? *_value : _init(); }
  template <typename UnaryOp> void combine_each(UnaryOp op) {
    if (_value) {
      op(*_value);
    }
  }

private:
  std::function<T()> _init;
  std::unique_ptr<T> _value;
};

// ---- concurrent_vector ----------------------------------------------------------------------
template <typename T, typename Allocator = std::allocator<T>>
class concurrent_vector : public std::vector<T> {
  using base = std::vector<T>;

public:
  using base::base;
  using iterator = typename base::iterator;
  using const_iterator = typename base::const_iterator;

  class range_type {
  public:
    range_type(iterator b, iterator e) : _b(b), _e(e) {}
    iterator begin() const { return _b; }
    iterator end() const { return _e; }
    bool empty() const { return _b == _e; }
    std::size_t size() const { return static_cast<std::size_t>(_e - _b); }
    bool is_divisible() const { return false; }

  private:
    iterator _b, _e;
  };

  // growth is serialised; like with tbb::concurrent_vector callers must not rely on iterators
  // obtained before a concurrent growth
  iterator push_back(const T &v) {
    iterator it;
#ifdef KMP_SHIM_PARALLEL
#pragma omp critical(kmp_shim_cvec)
#endif
    {
      base::push_back(v);
      it = base::end() - 1;
    }
    return it;
  }
  iterator push_back(T &&v) {
    iterator it;
#ifdef KMP_SHIM_PARALLEL
#pragma omp critical(kmp_shim_cvec)
#endif
    {
      base::push_back(std::move(v));
      it = base::end() - 1;
    }
    return it;
  }
  template <typename... Args> iterator emplace_back(Args &&...args) {
    iterator it;
#ifdef KMP_SHIM_PARALLEL
#pragma omp critical(kmp_shim_cvec)
#endif
    {
      base::emplace_back(std::forward<Args>(args)...);
      it = base::end() - 1;
    }
    return it;
  }
  iterator grow_by(std::size_t delta) {
    const std::size_t old = base::size();
    base::resize(old + delta);
    return base::begin() + static_cast<std::ptrdiff_t>(old);
  }
  iterator grow_by(std::size_t delta, const T &v) {
    const std::size_t old = base::size();
    base::resize(old + delta, v);
    return base::begin() + static_cast<std::ptrdiff_t>(old);
  }
  iterator grow_to_at_least(std::size_t n) {
    const std::size_t old = base::size();
    if (n > old) {
      base::resize(n);
    }
    return base::begin() + static_cast<std::ptrdiff_t>(old);
  }
  range_type range(std::size_t = 1) { return range_type(base::begin(), base::end()); }
};

} // namespace tbb

namespace oneapi {
namespace tbb = ::tbb;
}

// tbbmalloc C API (only referenced when KAMINPAR_ENABLE_TBB_MALLOC is defined; we leave it off).
inline void *scalable_malloc(std::size_t n) { return std::malloc(n); }
inline void scalable_free(void *p) { std::free(p); }
inline int scalable_posix_memalign(void **p, std::size_t a, std::size_t n) {
  return posix_memalign(p, a, n);
}
