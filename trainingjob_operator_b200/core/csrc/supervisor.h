// Process supervisor: the kubelet analogue of the single-box design.
// The reference never runs processes itself -- it creates Pods (pkg/controller/pod.go:483-546)
// and reads their phase / container exit codes back (pod.go:339-379).  Here one OS process
// (its own process group) is spawned per replica and observed directly:
//   * fork + execve with explicit env / cwd / stdio redirection / CPU affinity,
//   * exec failure is reported synchronously with errno (the CreateContainerError analogue,
//     pkg/apis/aitrainingjob/v1/constants.go:46-56),
//   * exit is observed through a pidfd + poll() reaper thread; a signal N is reported as exit
//     code 128+N (so `kill -9` => 137, matching example/paddle-mnist.yaml:7),
//   * kill() signals the whole process group.
#pragma once
#include <fcntl.h>
#include <poll.h>
#include <spawn.h>
#include <sched.h>
#include <signal.h>
#include <sys/eventfd.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace aitj {

struct SpawnError : std::runtime_error {
  int err;
  SpawnError(int e, const std::string& msg) : std::runtime_error(msg), err(e) {}
};

struct ExitEvent {
  std::string id;
  int pid = 0;
  int exit_code = 0;  // 128+signal when killed by a signal
  int signal = 0;
  double wall_time_s = 0.0;
  bool status_unknown = false;  // an adopted (non-child) process: its exit is seen through the pidfd, its status is not
};

struct ProcInfo {
  std::string id;
  int pid = 0;
  int pidfd = -1;
  bool adopted = false;
  std::chrono::steady_clock::time_point started;
};

// Kernel start time of `pid` in clock ticks since boot (field 22 of /proc/<pid>/stat), 0 if the pid is gone.  With the
// pid it identifies a process across agent restarts (pids are recycled, start times are not).
inline long long proc_start_time(int pid) {
  char path[64];
  snprintf(path, sizeof(path), "/proc/%d/stat", pid);
  FILE* f = fopen(path, "r");
  if (!f) return 0;
  char buf[1024];
  size_t n = fread(buf, 1, sizeof(buf) - 1, f);
  fclose(f);
  buf[n] = 0;
  const char* p = strrchr(buf, ')');        // the command name may contain spaces and parentheses
  if (!p) return 0;
  ++p;
  long long v = 0;
  // after ")": state(3) ppid(4) ... starttime is field 22 => the 20th token after the ')'
  int field = 2;
  while (*p) {
    while (*p == ' ') ++p;
    if (!*p) break;
    ++field;
    const char* q = p;
    while (*q && *q != ' ') ++q;
    if (field == 22) {
      v = atoll(p);
      break;
    }
    p = q;
  }
  return v;
}

class Supervisor {
 public:
  Supervisor() {
    wake_fd_ = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
    reaper_ = std::thread([this] { reap_loop(); });
  }
  ~Supervisor() {
    stop_ = true;
    wake();
    if (reaper_.joinable()) reaper_.join();
    if (wake_fd_ >= 0) close(wake_fd_);
  }

  // Spawns argv[0] (PATH lookup) as a new process-group leader. Throws SpawnError(errno).
  int spawn(const std::string& id, const std::vector<std::string>& argv, const std::map<std::string, std::string>& env,
            const std::string& cwd, const std::string& stdout_path, const std::string& stderr_path,
            const std::vector<int>& cpus) {
    if (argv.empty()) throw SpawnError(EINVAL, "empty argv");
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (procs_.count(id)) throw SpawnError(EEXIST, "process id already supervised: " + id);
    }
    std::vector<std::string> env_strs;
    env_strs.reserve(env.size());
    for (auto& kv : env) env_strs.push_back(kv.first + "=" + kv.second);
    std::vector<char*> c_argv, c_env;
    for (auto& a : argv) c_argv.push_back(const_cast<char*>(a.c_str()));
    c_argv.push_back(nullptr);
    for (auto& e : env_strs) c_env.push_back(const_cast<char*>(e.c_str()));
    c_env.push_back(nullptr);

    // posix_spawn (CLONE_VFORK | CLONE_VM in glibc) instead of fork + exec: the agent is a large process and copying
    // its page tables dominated the per-replica start cost (2.7 ms -> well under 1 ms).  The stages that would fail
    // inside the child are checked here first so that the error still says which one it was.
    if (!cwd.empty()) {
      struct stat st;
      if (stat(cwd.c_str(), &st) != 0) throw SpawnError(errno, std::string("chdir failed: ") + strerror(errno) + " (" + cwd + ")");
      if (!S_ISDIR(st.st_mode)) throw SpawnError(ENOTDIR, std::string("chdir failed: ") + strerror(ENOTDIR) + " (" + cwd + ")");
    }
    for (const std::string* path : {&stdout_path, &stderr_path}) {
      if (path->empty()) continue;
      int fd = open(path->c_str(), O_WRONLY | O_CREAT | O_APPEND | O_CLOEXEC, 0644);
      if (fd < 0) throw SpawnError(errno, std::string("open log failed: ") + strerror(errno) + " (" + *path + ")");
      close(fd);
    }
    posix_spawn_file_actions_t fa;
    posix_spawnattr_t attr;
    posix_spawn_file_actions_init(&fa);
    posix_spawnattr_init(&attr);
    if (!cwd.empty()) posix_spawn_file_actions_addchdir_np(&fa, cwd.c_str());
    posix_spawn_file_actions_addopen(&fa, 0, "/dev/null", O_RDONLY, 0);
    if (!stdout_path.empty()) {
      posix_spawn_file_actions_addopen(&fa, 1, stdout_path.c_str(), O_WRONLY | O_CREAT | O_APPEND, 0644);
      if (stderr_path.empty() || stderr_path == stdout_path) posix_spawn_file_actions_adddup2(&fa, 1, 2);
    }
    if (!stderr_path.empty() && stderr_path != stdout_path)
      posix_spawn_file_actions_addopen(&fa, 2, stderr_path.c_str(), O_WRONLY | O_CREAT | O_APPEND, 0644);
    sigset_t none, all;
    sigemptyset(&none);
    sigfillset(&all);
    posix_spawnattr_setsigmask(&attr, &none);
    posix_spawnattr_setsigdefault(&attr, &all);
    posix_spawnattr_setpgroup(&attr, 0);          // own process group: signals reach the whole replica, nothing else
    posix_spawnattr_setflags(&attr, POSIX_SPAWN_SETPGROUP | POSIX_SPAWN_SETSIGMASK | POSIX_SPAWN_SETSIGDEF);
    pid_t pid = -1;
    const int rc = posix_spawnp(&pid, c_argv[0], &fa, &attr, c_argv.data(), c_env.data());
    posix_spawn_file_actions_destroy(&fa);
    posix_spawnattr_destroy(&attr);
    if (rc != 0) throw SpawnError(rc, std::string("exec failed: ") + strerror(rc) + " (" + argv[0] + ")");
    if (!cpus.empty()) {
      cpu_set_t mask;
      CPU_ZERO(&mask);
      for (int c : cpus) if (c >= 0 && c < CPU_SETSIZE) CPU_SET(c, &mask);
      sched_setaffinity(pid, sizeof(mask), &mask);   // inherited by every thread the replica creates from here on
    }
    ProcInfo info;
    info.id = id;
    info.pid = pid;
    info.pidfd = static_cast<int>(syscall(SYS_pidfd_open, pid, 0));
    info.started = std::chrono::steady_clock::now();
    {
      std::lock_guard<std::mutex> lk(mu_);
      procs_[id] = info;
    }
    wake();
    return pid;
  }

  // Signal the process group (or only the leader). Returns false if unknown / already gone.
  bool kill_proc(const std::string& id, int sig, bool group) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = procs_.find(id);
    if (it == procs_.end()) return false;
    int rc = group ? ::kill(-it->second.pid, sig) : ::kill(it->second.pid, sig);
    if (rc != 0 && group) rc = ::kill(it->second.pid, sig);
    return rc == 0;
  }

  // Take over a process this supervisor did not spawn (the agent was restarted and finds its predecessor's workers
  // still running).  The caller has verified the pid's identity (start time).  Exit is observed through a pidfd; the
  // exit status of a non-child cannot be read, so its ExitEvent carries status_unknown.  Throws if the pid is gone.
  void adopt(const std::string& id, int pid) {
    int fd = static_cast<int>(syscall(SYS_pidfd_open, pid, 0));
    if (fd < 0) throw SpawnError(errno, std::string("pidfd_open failed: ") + strerror(errno));
    ProcInfo info;
    info.id = id;
    info.pid = pid;
    info.pidfd = fd;
    info.adopted = true;
    info.started = std::chrono::steady_clock::now();
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (procs_.count(id)) {
        close(fd);
        throw SpawnError(EEXIST, "process id already supervised: " + id);
      }
      procs_[id] = info;
    }
    wake();
  }

  // Re-key a supervised process (a pre-warmed worker adopted as a pod's container).  The exit event of the
  // process is reported under the new id.  False if `from` is gone or `to` is taken.
  bool rename(const std::string& from, const std::string& to) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = procs_.find(from);
      if (it == procs_.end() || procs_.count(to)) return false;
      ProcInfo info = it->second;
      info.id = to;
      info.started = std::chrono::steady_clock::now();
      procs_.erase(it);
      procs_[to] = info;
    }
    wake();
    return true;
  }

  bool alive(const std::string& id) {
    std::lock_guard<std::mutex> lk(mu_);
    return procs_.count(id) > 0;
  }
  int pid_of(const std::string& id) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = procs_.find(id);
    return it == procs_.end() ? -1 : it->second.pid;
  }
  std::vector<std::pair<std::string, int>> list() {
    std::lock_guard<std::mutex> lk(mu_);
    std::vector<std::pair<std::string, int>> out;
    for (auto& kv : procs_) out.emplace_back(kv.first, kv.second.pid);
    return out;
  }

  // Blocks up to timeout_s for reaped children.
  std::vector<ExitEvent> poll_exits(double timeout_s) {
    std::unique_lock<std::mutex> lk(ev_mu_);
    if (events_.empty())
      ev_cv_.wait_for(lk, std::chrono::duration<double>(timeout_s), [&] { return !events_.empty() || stop_.load(); });
    std::vector<ExitEvent> out(events_.begin(), events_.end());
    events_.clear();
    return out;
  }

 private:
  void wake() {
    uint64_t one = 1;
    ssize_t w = write(wake_fd_, &one, sizeof(one));
    (void)w;
  }

  void reap_loop() {
    while (!stop_) {
      std::vector<pollfd> fds;
      std::vector<std::string> ids;
      fds.push_back({wake_fd_, POLLIN, 0});
      {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto& kv : procs_) {
          if (kv.second.pidfd >= 0) {
            fds.push_back({kv.second.pidfd, POLLIN, 0});
            ids.push_back(kv.first);
          }
        }
      }
      int rc = poll(fds.data(), fds.size(), 500);
      if (rc < 0 && errno != EINTR) { usleep(1000); continue; }
      if (fds[0].revents & POLLIN) {
        uint64_t v;
        ssize_t r = read(wake_fd_, &v, sizeof(v));
        (void)r;
      }
      for (size_t i = 1; i < fds.size(); ++i) {
        if (!(fds[i].revents & (POLLIN | POLLHUP | POLLERR))) continue;
        reap(ids[i - 1]);
      }
      // pidfd_open can fail on exotic kernels: fall back to WNOHANG polling for those.
      std::vector<std::string> nofd;
      {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto& kv : procs_) if (kv.second.pidfd < 0) nofd.push_back(kv.first);
      }
      for (auto& id : nofd) reap(id, true);
    }
  }

  void reap(const std::string& id, bool nohang = false) {
    ProcInfo info;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = procs_.find(id);
      if (it == procs_.end()) return;
      info = it->second;
    }
    int st = 0;
    pid_t r = waitpid(info.pid, &st, (nohang || info.adopted) ? WNOHANG : 0);
    if (r == 0 && !info.adopted) return;
    ExitEvent ev;
    ev.id = id;
    ev.pid = info.pid;
    if (r <= 0) {
      if (info.adopted && r < 0 && errno == ECHILD) {
        // not our child: the readable pidfd (or a vanished pid) is all we get
        if (::kill(info.pid, 0) == 0 && nohang) return;
        ev.status_unknown = true;
        ev.exit_code = 137;
      } else if (r == 0) {
        return;
      } else {
        ev.exit_code = 255;
      }
    } else if (WIFSIGNALED(st)) {
      ev.signal = WTERMSIG(st);
      ev.exit_code = 128 + ev.signal;
    } else {
      ev.exit_code = WEXITSTATUS(st);
    }
    ev.wall_time_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - info.started).count();
    {
      // One critical section over both tables: an observer never sees the process gone from `list()` while its exit
      // event is not queued yet.  The entry is looked up again by pid -- `rename` may have re-keyed it while waitpid ran,
      // and the event must carry the id the owner knows it by now.
      std::scoped_lock lk(mu_, ev_mu_);
      auto it = procs_.find(id);
      if (it == procs_.end() || it->second.pid != info.pid) {
        it = procs_.end();
        for (auto jt = procs_.begin(); jt != procs_.end(); ++jt)
          if (jt->second.pid == info.pid && jt->second.pidfd == info.pidfd) { it = jt; break; }
      }
      if (it != procs_.end()) {
        ev.id = it->first;
        procs_.erase(it);
      }
      events_.push_back(ev);
    }
    if (info.pidfd >= 0) close(info.pidfd);
    ev_cv_.notify_all();
  }

  std::mutex mu_;
  std::map<std::string, ProcInfo> procs_;
  std::mutex ev_mu_;
  std::condition_variable ev_cv_;
  std::deque<ExitEvent> events_;
  std::atomic<bool> stop_{false};
  int wake_fd_ = -1;
  std::thread reaper_;
};

}  // namespace aitj
