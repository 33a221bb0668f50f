// Shim: boost::detail::spinlock over std::atomic_flag (engine/db/index/knn/nndescent_common.hpp:40).
#pragma once
#include <atomic>
namespace boost {
namespace detail {
class spinlock {
  std::atomic_flag f_ = ATOMIC_FLAG_INIT;
 public:
  void lock() { while (f_.test_and_set(std::memory_order_acquire)) {} }
  void unlock() { f_.clear(std::memory_order_release); }
};
}  // namespace detail
}  // namespace boost
