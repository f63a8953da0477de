// SpinMutex: test-and-test-and-set lock that spins briefly and then yields.
//
// The control plane's critical sections (free-list splice, ledger insert) are a few hundred
// nanoseconds long but are taken by every client thread on every object; a futex-based mutex
// turns each collision into a sleep + wake-up (tens of microseconds), which is what made 8
// concurrent clients slower than one.  Usable with std::lock_guard / std::unique_lock.
#pragma once
#include <atomic>
#include <thread>

namespace bb {

class SpinMutex {
 public:
  void lock() noexcept {
    for (;;) {
      if (!flag_.exchange(true, std::memory_order_acquire)) return;
      int spins = 0;
      while (flag_.load(std::memory_order_relaxed)) {
        if (++spins < 128) {
#if defined(__x86_64__) || defined(__i386__)
          __builtin_ia32_pause();
#endif
        } else {
          std::this_thread::yield();  // owner was descheduled: stop burning its core
          spins = 0;
        }
      }
    }
  }
  bool try_lock() noexcept { return !flag_.load(std::memory_order_relaxed) && !flag_.exchange(true, std::memory_order_acquire); }
  void unlock() noexcept { flag_.store(false, std::memory_order_release); }
  // shared_lock compatibility: readers are exclusive too (sections are too short for a reader count to pay off)
  void lock_shared() noexcept { lock(); }
  bool try_lock_shared() noexcept { return try_lock(); }
  void unlock_shared() noexcept { unlock(); }

 private:
  std::atomic<bool> flag_{false};
};

}  // namespace bb
