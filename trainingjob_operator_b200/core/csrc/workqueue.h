// Rate-limited, delaying, de-duplicating work queue + controller expectations.
//
// Behavioural spec: the client-go workqueue the reference builds at
// pkg/controller/controller.go:113 (NewNamedRateLimitingQueue(DefaultControllerRateLimiter()))
// and the ControllerExpectations cache at controller.go:112 -- neither is
// vendored in the reference tree, so this is a fresh implementation of the documented
// semantics:
//   * a key that is added while queued is coalesced; a key added while being processed is
//     re-queued after Done();
//   * AddRateLimited delay = max(per-item 5ms*2^n capped at 1000s, token bucket 10qps/100);
//   * AddAfter(key, d) delayed add; Forget(key) resets the per-item back-off;
//   * Expectations: {add,del} counters per key with a 5 minute TTL.  Unlike the reference
//     (quirk Q3: ExpectCreations overwrites per pod), Raise/Lower accumulate.
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <map>
#include <mutex>
#include <optional>
#include <queue>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace aitj {

using Clock = std::chrono::steady_clock;

class WorkQueue {
 public:
  explicit WorkQueue(std::string name, double base_delay_s = 0.005, double max_delay_s = 1000.0, double qps = 10.0,
                     int burst = 100)
      : name_(std::move(name)), base_delay_(base_delay_s), max_delay_(max_delay_s), qps_(qps), burst_(burst),
        tokens_(burst), last_refill_(Clock::now()) {
    timer_ = std::thread([this] { timer_loop(); });
  }
  ~WorkQueue() {
    shutdown();
    if (timer_.joinable()) timer_.join();
  }

  void add(const std::string& key) {
    std::lock_guard<std::mutex> lk(mu_);
    add_locked(key);
  }

  void add_after(const std::string& key, double seconds) {
    if (seconds <= 0) { add(key); return; }
    std::lock_guard<std::mutex> lk(mu_);
    if (shutting_down_) return;
    auto when = Clock::now() + std::chrono::duration_cast<Clock::duration>(std::chrono::duration<double>(seconds));
    auto it = waiting_.find(key);
    if (it != waiting_.end()) {
      if (it->second <= when) return;  // an earlier wake-up is already scheduled
      it->second = when;
    } else {
      waiting_[key] = when;
    }
    heap_.push({when, key});
    timer_cv_.notify_all();
  }

  // Returns the delay that was applied (seconds).
  double add_rate_limited(const std::string& key) {
    double d;
    {
      std::lock_guard<std::mutex> lk(mu_);
      d = when_locked(key);
    }
    add_after(key, d);
    return d;
  }

  void forget(const std::string& key) {
    std::lock_guard<std::mutex> lk(mu_);
    failures_.erase(key);
  }

  int num_requeues(const std::string& key) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = failures_.find(key);
    return it == failures_.end() ? 0 : it->second;
  }

  // Blocks up to timeout_s (<0: forever). nullopt on timeout or shutdown-with-empty-queue.
  std::optional<std::string> get(double timeout_s) {
    std::unique_lock<std::mutex> lk(mu_);
    auto ready = [this] { return !queue_.empty() || shutting_down_; };
    if (timeout_s < 0) {
      cv_.wait(lk, ready);
    } else if (!cv_.wait_for(lk, std::chrono::duration<double>(timeout_s), ready)) {
      return std::nullopt;
    }
    if (queue_.empty()) return std::nullopt;
    std::string key = std::move(queue_.front());
    queue_.pop_front();
    processing_.insert(key);
    dirty_.erase(key);
    return key;
  }

  void done(const std::string& key) {
    std::lock_guard<std::mutex> lk(mu_);
    processing_.erase(key);
    if (dirty_.count(key)) {
      queue_.push_back(key);
      cv_.notify_one();
    }
  }

  size_t len() {
    std::lock_guard<std::mutex> lk(mu_);
    return queue_.size();
  }
  size_t len_waiting() {
    std::lock_guard<std::mutex> lk(mu_);
    return waiting_.size();
  }
  void shutdown() {
    std::lock_guard<std::mutex> lk(mu_);
    shutting_down_ = true;
    cv_.notify_all();
    timer_cv_.notify_all();
  }
  bool shutting_down() {
    std::lock_guard<std::mutex> lk(mu_);
    return shutting_down_;
  }
  const std::string& name() const { return name_; }

 private:
  struct Waiting {
    Clock::time_point when;
    std::string key;
    bool operator>(const Waiting& o) const { return when > o.when; }
  };

  void add_locked(const std::string& key) {
    if (shutting_down_) return;
    if (dirty_.count(key)) return;
    dirty_.insert(key);
    if (processing_.count(key)) return;
    queue_.push_back(key);
    cv_.notify_one();
  }

  double when_locked(const std::string& key) {
    // per-item exponential back-off
    int n = failures_[key]++;
    double item = base_delay_;
    for (int i = 0; i < n && item < max_delay_; ++i) item *= 2.0;
    if (item > max_delay_) item = max_delay_;
    // overall token bucket
    auto now = Clock::now();
    double elapsed = std::chrono::duration<double>(now - last_refill_).count();
    tokens_ = std::min<double>(burst_, tokens_ + elapsed * qps_);
    last_refill_ = now;
    double bucket = 0.0;
    tokens_ -= 1.0;
    if (tokens_ < 0) bucket = -tokens_ / qps_;
    return std::max(item, bucket);
  }

  void timer_loop() {
    std::unique_lock<std::mutex> lk(mu_);
    while (!shutting_down_) {
      if (heap_.empty()) {
        timer_cv_.wait(lk);
        continue;
      }
      auto top = heap_.top();
      auto now = Clock::now();
      if (top.when > now) {
        timer_cv_.wait_until(lk, top.when);
        continue;
      }
      heap_.pop();
      auto it = waiting_.find(top.key);
      if (it == waiting_.end() || it->second != top.when) continue;  // superseded entry
      waiting_.erase(it);
      add_locked(top.key);
    }
  }

  std::string name_;
  double base_delay_, max_delay_, qps_;
  int burst_;
  double tokens_;
  Clock::time_point last_refill_;

  std::mutex mu_;
  std::condition_variable cv_, timer_cv_;
  std::deque<std::string> queue_;
  std::unordered_set<std::string> dirty_, processing_;
  std::unordered_map<std::string, int> failures_;
  std::unordered_map<std::string, Clock::time_point> waiting_;
  std::priority_queue<Waiting, std::vector<Waiting>, std::greater<Waiting>> heap_;
  bool shutting_down_ = false;
  std::thread timer_;
};

class Expectations {
 public:
  explicit Expectations(double ttl_s = 300.0) : ttl_(ttl_s) {}

  void set(const std::string& key, int64_t add, int64_t del) {
    std::lock_guard<std::mutex> lk(mu_);
    recs_[key] = Rec{add, del, Clock::now()};
  }
  void expect_creations(const std::string& key, int64_t n) { set(key, n, 0); }
  void expect_deletions(const std::string& key, int64_t n) { set(key, 0, n); }
  void raise(const std::string& key, int64_t add, int64_t del) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = recs_.find(key);
    if (it == recs_.end() || expired(it->second)) {
      recs_[key] = Rec{add, del, Clock::now()};
    } else {
      it->second.add = std::max<int64_t>(it->second.add, 0) + add;
      it->second.del = std::max<int64_t>(it->second.del, 0) + del;
      it->second.ts = Clock::now();
    }
  }
  void creation_observed(const std::string& key) { lower(key, 1, 0); }
  void deletion_observed(const std::string& key) { lower(key, 0, 1); }
  void lower(const std::string& key, int64_t add, int64_t del) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = recs_.find(key);
    if (it == recs_.end()) return;
    it->second.add -= add;
    it->second.del -= del;
  }
  bool satisfied(const std::string& key) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = recs_.find(key);
    if (it == recs_.end()) return true;
    if (it->second.add <= 0 && it->second.del <= 0) return true;
    return expired(it->second);
  }
  void erase(const std::string& key) {
    std::lock_guard<std::mutex> lk(mu_);
    recs_.erase(key);
  }
  std::optional<std::pair<int64_t, int64_t>> peek(const std::string& key) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = recs_.find(key);
    if (it == recs_.end()) return std::nullopt;
    return std::make_pair(it->second.add, it->second.del);
  }

 private:
  struct Rec {
    int64_t add, del;
    Clock::time_point ts;
  };
  bool expired(const Rec& r) const { return std::chrono::duration<double>(Clock::now() - r.ts).count() > ttl_; }
  double ttl_;
  std::mutex mu_;
  std::unordered_map<std::string, Rec> recs_;
};

}  // namespace aitj
