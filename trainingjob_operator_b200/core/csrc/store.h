// Versioned, watchable object store: the kube-apiserver + etcd analogue of the single-box
// design.  The reference talks to a real apiserver through
// client-go (pkg/client/clientset/versioned/clientset.go:61-78); here the same contract --
// monotonically increasing resourceVersion, optimistic concurrency on update, LIST + WATCH
// from a resourceVersion, owner-reference cascade on delete -- is provided in-process with an
// optional write-ahead log so a restarted daemon resumes from disk.
//
// Objects are opaque serialized bytes (JSON produced by the Python API layer) plus the few
// indexed attributes the store itself needs: labels (selector filtering), uid and owner uids
// (cascading delete).  resourceVersion is owned by the store and injected by the caller on
// the way out.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unistd.h>
#include <unordered_map>
#include <utility>
#include <vector>

namespace aitj {

struct StoreError : std::runtime_error {
  std::string reason;  // NotFound | AlreadyExists | Conflict | Gone
  StoreError(const std::string& r, const std::string& msg) : std::runtime_error(msg), reason(r) {}
};

using Labels = std::map<std::string, std::string>;

struct StoredObject {
  std::string kind, ns, name, uid, data;
  Labels labels;
  std::vector<std::string> owner_uids;
  uint64_t rv = 0;
};

struct WatchEvent {
  std::string type;  // ADDED | MODIFIED | DELETED
  StoredObject obj;
};

class Store {
 public:
  explicit Store(const std::string& wal_path = "", size_t history = 16384) : history_cap_(history) {
    if (!wal_path.empty()) {
      wal_path_ = wal_path;
      replay();
      rebuild_owner_index();
      // what was logged before this start cannot be replayed to a watcher: a watch from an older resourceVersion must
      // be told to re-list (Gone) instead of silently missing the events in between
      history_floor_ = rv_;
      bool need_nl = false;
      {
        std::ifstream f(wal_path_, std::ios::binary | std::ios::ate);
        if (f.good() && f.tellg() > 0) {
          f.seekg(-1, std::ios::end);
          need_nl = f.get() != '\n';
        }
      }
      wal_.open(wal_path_, std::ios::app | std::ios::binary);
      // a crash can cut the log right after a record's last digit; without a separator the next record would be glued to it
      if (need_nl) { wal_ << '\n'; wal_.flush(); }
    }
  }

  uint64_t current_rv() {
    std::lock_guard<std::mutex> lk(mu_);
    return rv_;
  }

  StoredObject create(StoredObject o) {
    std::unique_lock<std::mutex> lk(mu_);
    auto& m = objs_[o.kind];
    const std::string key = o.ns + "/" + o.name;
    if (m.count(key)) throw StoreError("AlreadyExists", o.kind + " \"" + o.name + "\" already exists");
    o.rv = ++rv_;
    m[key] = o;
    index_owners(o, key, true);
    log_put(o);
    publish("ADDED", o);
    return o;
  }

  // expected_rv == 0: unconditional.
  StoredObject update(StoredObject o, uint64_t expected_rv) {
    std::unique_lock<std::mutex> lk(mu_);
    auto& m = objs_[o.kind];
    const std::string key = o.ns + "/" + o.name;
    auto it = m.find(key);
    if (it == m.end()) throw StoreError("NotFound", o.kind + " \"" + o.name + "\" not found");
    if (expected_rv != 0 && it->second.rv != expected_rv)
      throw StoreError("Conflict", "Operation cannot be fulfilled on " + o.kind + " \"" + o.name +
                                       "\": the object has been modified; please apply your changes to the latest "
                                       "version and try again");
    if (o.uid.empty()) o.uid = it->second.uid;
    o.rv = ++rv_;
    if (it->second.owner_uids != o.owner_uids) {
      index_owners(it->second, key, false);
      index_owners(o, key, true);
    }
    it->second = o;
    log_put(o);
    publish("MODIFIED", o);
    return o;
  }

  StoredObject get(const std::string& kind, const std::string& ns, const std::string& name) {
    std::lock_guard<std::mutex> lk(mu_);
    auto kit = objs_.find(kind);
    if (kit != objs_.end()) {
      auto it = kit->second.find(ns + "/" + name);
      if (it != kit->second.end()) return it->second;
    }
    throw StoreError("NotFound", kind + " \"" + name + "\" not found");
  }

  // ns == "" lists every namespace. Equality selector on labels. Returns list resourceVersion.
  std::pair<std::vector<StoredObject>, uint64_t> list(const std::string& kind, const std::string& ns,
                                                      const Labels& selector) {
    std::lock_guard<std::mutex> lk(mu_);
    std::vector<StoredObject> out;
    auto kit = objs_.find(kind);
    if (kit != objs_.end()) {
      for (auto& kv : kit->second) {
        const StoredObject& o = kv.second;
        if (!ns.empty() && o.ns != ns) continue;
        if (!matches(o.labels, selector)) continue;
        out.push_back(o);
      }
    }
    return {out, rv_};
  }

  // Deletes the object and (background-GC semantics) every object that lists its uid as owner.
  // Returns the deleted objects, the requested one first.
  std::vector<StoredObject> remove(const std::string& kind, const std::string& ns, const std::string& name,
                                   const std::string& expected_uid = "") {
    std::unique_lock<std::mutex> lk(mu_);
    auto kit = objs_.find(kind);
    if (kit == objs_.end()) throw StoreError("NotFound", kind + " \"" + name + "\" not found");
    auto it = kit->second.find(ns + "/" + name);
    if (it == kit->second.end()) throw StoreError("NotFound", kind + " \"" + name + "\" not found");
    if (!expected_uid.empty() && it->second.uid != expected_uid)
      throw StoreError("Conflict", "uid precondition failed for " + kind + " \"" + name + "\"");
    std::vector<StoredObject> removed;
    std::vector<std::string> frontier;
    StoredObject victim = it->second;
    index_owners(victim, it->first, false);
    kit->second.erase(it);
    victim.rv = ++rv_;
    log_del(victim);
    publish("DELETED", victim);
    removed.push_back(victim);
    if (!victim.uid.empty()) frontier.push_back(victim.uid);
    // dependents come from the owner index (uid -> objects that list it), not from a sweep over every stored object:
    // under the throughput benchmark a sweep visited thousands of pods / services / events per delete, under the lock
    while (!frontier.empty()) {
      const std::string owner = frontier.back();
      frontier.pop_back();
      auto oi = owned_by_.find(owner);
      if (oi == owned_by_.end()) continue;
      const std::set<std::pair<std::string, std::string>> deps = oi->second;   // erased from below
      for (auto& kk : deps) {
        auto dk = objs_.find(kk.first);
        if (dk == objs_.end()) continue;
        auto oit = dk->second.find(kk.second);
        if (oit == dk->second.end()) continue;
        StoredObject dep = oit->second;
        index_owners(dep, oit->first, false);
        dk->second.erase(oit);
        dep.rv = ++rv_;
        log_del(dep);
        publish("DELETED", dep);
        removed.push_back(dep);
        if (!dep.uid.empty()) frontier.push_back(dep.uid);
      }
    }
    return removed;
  }

  // ---- watch ------------------------------------------------------------------------------
  // since_rv == 0: start from "now" (no replay).  Otherwise replay history with rv > since_rv;
  // throws Gone if that history has been trimmed.
  int64_t watch_open(const std::string& kind, const std::string& ns, uint64_t since_rv) {
    std::lock_guard<std::mutex> lk(mu_);
    auto w = std::make_shared<Watcher>();
    w->kind = kind;
    w->ns = ns;
    if (since_rv != 0 && since_rv < rv_) {
      // history_floor_: newest resourceVersion whose event is no longer (or was never) in history_
      if (since_rv < history_floor_)
        throw StoreError("Gone", "too old resource version: " + std::to_string(since_rv));
      for (auto& ev : history_) {
        if (ev.obj.rv <= since_rv) continue;
        if (ev.obj.kind != kind) continue;
        if (!ns.empty() && ev.obj.ns != ns) continue;
        w->q.push_back(ev);
      }
    }
    const int64_t id = ++watch_id_;
    watchers_[id] = w;
    return id;
  }

  // Blocks up to timeout_s. nullopt on timeout / closed watcher.
  std::optional<WatchEvent> watch_next(int64_t id, double timeout_s) {
    std::shared_ptr<Watcher> w;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = watchers_.find(id);
      if (it == watchers_.end()) return std::nullopt;
      w = it->second;
    }
    std::unique_lock<std::mutex> lk(w->mu);
    if (!w->cv.wait_for(lk, std::chrono::duration<double>(timeout_s), [&] { return !w->q.empty() || w->closed; }))
      return std::nullopt;
    if (w->q.empty()) return std::nullopt;
    WatchEvent ev = std::move(w->q.front());
    w->q.pop_front();
    return ev;
  }

  // Up to `max` queued events: waits (at most timeout_s) only when nothing is queued.  `wait == false` never blocks --
  // the binding calls it that way first, with the interpreter lock held, and gives the lock up only for a real wait.
  std::vector<WatchEvent> watch_next_many(int64_t id, double timeout_s, size_t max, bool wait = true) {
    std::vector<WatchEvent> out;
    std::shared_ptr<Watcher> w;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = watchers_.find(id);
      if (it == watchers_.end()) return out;
      w = it->second;
    }
    std::unique_lock<std::mutex> lk(w->mu);
    if (w->q.empty()) {
      if (!wait) return out;
      w->cv.wait_for(lk, std::chrono::duration<double>(timeout_s), [&] { return !w->q.empty() || w->closed; });
    }
    while (!w->q.empty() && out.size() < max) {
      out.push_back(std::move(w->q.front()));
      w->q.pop_front();
    }
    return out;
  }

  void watch_close(int64_t id) {
    std::shared_ptr<Watcher> w;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = watchers_.find(id);
      if (it == watchers_.end()) return;
      w = it->second;
      watchers_.erase(it);
    }
    std::lock_guard<std::mutex> lk(w->mu);
    w->closed = true;
    w->cv.notify_all();
  }

  size_t num_watchers() {
    std::lock_guard<std::mutex> lk(mu_);
    return watchers_.size();
  }

  size_t count(const std::string& kind) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = objs_.find(kind);
    return it == objs_.end() ? 0 : it->second.size();
  }

  // Rewrite the WAL as a snapshot of the live objects.
  void compact() {
    std::lock_guard<std::mutex> lk(mu_);
    if (wal_path_.empty()) return;
    wal_.close();
    const std::string tmp = wal_path_ + ".tmp";
    {
      std::ofstream f(tmp, std::ios::trunc | std::ios::binary);
      // the snapshot only holds live objects: when the newest operations were deletes the highest rv in it is below
      // rv_, and resourceVersions would be handed out twice after a restart -- so the counter itself is recorded
      f << "V " << rv_ << '\n';
      for (auto& kk : objs_)
        for (auto& kv : kk.second) write_put(f, kv.second);
    }
    std::rename(tmp.c_str(), wal_path_.c_str());
    wal_.open(wal_path_, std::ios::app | std::ios::binary);
  }

 private:
  // owner uid -> (kind, "ns/name") of the objects whose ownerReferences name it.  Caller holds mu_.
  void index_owners(const StoredObject& o, const std::string& key, bool add) {
    for (auto& u : o.owner_uids) {
      if (add) {
        owned_by_[u].insert({o.kind, key});
      } else {
        auto it = owned_by_.find(u);
        if (it == owned_by_.end()) continue;
        it->second.erase({o.kind, key});
        if (it->second.empty()) owned_by_.erase(it);
      }
    }
  }

  void rebuild_owner_index() {
    owned_by_.clear();
    for (auto& kk : objs_)
      for (auto& kv : kk.second) index_owners(kv.second, kv.first, true);
  }

  struct Watcher {
    std::string kind, ns;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<WatchEvent> q;
    bool closed = false;
  };

  static bool matches(const Labels& labels, const Labels& sel) {
    for (auto& kv : sel) {
      auto it = labels.find(kv.first);
      if (it == labels.end() || it->second != kv.second) return false;
    }
    return true;
  }

  void publish(const char* type, const StoredObject& o) {
    WatchEvent ev{type, o};
    history_.push_back(ev);
    if (history_.size() > history_cap_) {
      history_floor_ = history_.front().obj.rv;
      history_.pop_front();
    }
    for (auto& kv : watchers_) {
      Watcher& w = *kv.second;
      if (w.kind != o.kind) continue;
      if (!w.ns.empty() && w.ns != o.ns) continue;
      std::lock_guard<std::mutex> lk(w.mu);
      w.q.push_back(ev);
      w.cv.notify_all();
    }
  }

  // ---- WAL: length-prefixed text records -------------------------------------------------
  static void write_str(std::ostream& f, const std::string& s) { f << s.size() << ' ' << s; }
  static bool read_str(std::istream& f, std::string& s) {
    size_t n;
    if (!(f >> n)) return false;
    f.get();
    s.resize(n);
    f.read(&s[0], static_cast<std::streamsize>(n));
    return static_cast<size_t>(f.gcount()) == n;
  }
  static void write_put(std::ostream& f, const StoredObject& o) {
    f << "P " << o.rv << ' ';
    write_str(f, o.kind); write_str(f, o.ns); write_str(f, o.name); write_str(f, o.uid); write_str(f, o.data);
    f << ' ' << o.labels.size() << ' ';
    for (auto& kv : o.labels) { write_str(f, kv.first); write_str(f, kv.second); }
    f << ' ' << o.owner_uids.size() << ' ';
    for (auto& u : o.owner_uids) write_str(f, u);
    f << '\n';
  }
  void log_put(const StoredObject& o) {
    if (!wal_.is_open()) return;
    write_put(wal_, o);
    wal_.flush();
  }
  void log_del(const StoredObject& o) {
    if (!wal_.is_open()) return;
    wal_ << "D " << o.rv << ' ';
    write_str(wal_, o.kind); write_str(wal_, o.ns); write_str(wal_, o.name);
    wal_ << '\n';
    wal_.flush();
  }
  // Replays the log; a record torn by a crash in the middle of an append ends the replay, and the file is cut back to
  // the last complete record so that what is appended from now on is not hidden behind the torn bytes at the next start.
  void replay() {
    std::ifstream f(wal_path_, std::ios::binary);
    if (!f.good()) return;
    std::string op;
    std::streamoff good = 0;
    struct Trunc {
      const std::string& path; std::ifstream& f; std::streamoff& good;
      ~Trunc() {
        f.clear();
        f.seekg(0, std::ios::end);
        const std::streamoff size = f.tellg();
        f.close();
        if (size > good && good >= 0) { int rc = ::truncate(path.c_str(), good); (void)rc; }
      }
    } trunc{wal_path_, f, good};
    while (f >> op) {
      uint64_t rv;
      if (!(f >> rv)) break;
      f.get();
      if (op == "P") {
        StoredObject o;
        o.rv = rv;
        if (!read_str(f, o.kind) || !read_str(f, o.ns) || !read_str(f, o.name) || !read_str(f, o.uid) ||
            !read_str(f, o.data))
          break;
        size_t nl, no;
        if (!(f >> nl)) break;
        f.get();
        bool bad = false;
        for (size_t i = 0; i < nl; ++i) {
          std::string k, v;
          if (!read_str(f, k) || !read_str(f, v)) { bad = true; break; }
          o.labels[k] = v;
        }
        if (bad || !(f >> no)) break;
        f.get();
        for (size_t i = 0; i < no; ++i) {
          std::string u;
          if (!read_str(f, u)) { bad = true; break; }
          o.owner_uids.push_back(u);
        }
        if (bad) break;
        objs_[o.kind][o.ns + "/" + o.name] = o;
      } else if (op == "V") {
        // resourceVersion counter at the time of a compaction (no object)
      } else if (op == "D") {
        std::string kind, ns, name;
        if (!read_str(f, kind) || !read_str(f, ns) || !read_str(f, name)) break;
        auto it = objs_.find(kind);
        if (it != objs_.end()) it->second.erase(ns + "/" + name);
      } else {
        break;
      }
      if (rv > rv_) rv_ = rv;
      if (f.peek() == '\n') f.get();
      good = f.tellg();
    }
  }

  std::mutex mu_;
  uint64_t rv_ = 0;
  std::map<std::string, std::map<std::string, StoredObject>> objs_;
  std::unordered_map<std::string, std::set<std::pair<std::string, std::string>>> owned_by_;
  std::deque<WatchEvent> history_;
  size_t history_cap_;
  uint64_t history_floor_ = 0;
  std::unordered_map<int64_t, std::shared_ptr<Watcher>> watchers_;
  int64_t watch_id_ = 0;
  std::string wal_path_;
  std::ofstream wal_;
};

}  // namespace aitj
