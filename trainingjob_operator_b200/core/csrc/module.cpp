// pybind11 bindings of the native control-plane runtime (work-queue, expectations, object
// store, process supervisor).  Blocking calls release the GIL so reconcile workers, informer
// threads and the supervisor's reaper run truly concurrently.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>

#include "store.h"
#include "supervisor.h"
#include "workqueue.h"

namespace py = pybind11;
using namespace aitj;

namespace {

py::dict obj_to_dict(const StoredObject& o) {
  py::dict d;
  d["kind"] = o.kind;
  d["namespace"] = o.ns;
  d["name"] = o.name;
  d["uid"] = o.uid;
  d["data"] = py::bytes(o.data);
  d["labels"] = o.labels;
  d["owner_uids"] = o.owner_uids;
  d["rv"] = o.rv;
  return d;
}

StoredObject make_obj(const std::string& kind, const std::string& ns, const std::string& name, const std::string& uid,
                      const py::bytes& data, const Labels& labels, const std::vector<std::string>& owners) {
  StoredObject o;
  o.kind = kind;
  o.ns = ns;
  o.name = name;
  o.uid = uid;
  o.data = static_cast<std::string>(data);
  o.labels = labels;
  o.owner_uids = owners;
  return o;
}

// Deep copy of a JSON-shaped tree (dict / list of dict, list, str, int, float, bool, None).  API objects travel as
// such trees and every lister read, claim pass and typed conversion copies one; copy.deepcopy spends most of its time
// on memo bookkeeping that a tree of immutable leaves does not need.  Anything that is not a dict or a list (tuples,
// dataclasses, ...) is handed to `fallback` (copy.deepcopy) so the function is a drop-in replacement.
PyObject* jcopy_impl(PyObject* x, PyObject* fallback, int depth) {
  if (depth > 200) {
    PyErr_SetString(PyExc_RecursionError, "jcopy: object nested too deeply (or cyclic)");
    return nullptr;
  }
  if (PyDict_CheckExact(x)) {
    PyObject* out = PyDict_New();
    if (!out) return nullptr;
    PyObject *k, *v;
    Py_ssize_t pos = 0;
    while (PyDict_Next(x, &pos, &k, &v)) {
      PyObject* kc;
      if (PyUnicode_CheckExact(k) || PyLong_CheckExact(k)) {
        kc = k;
        Py_INCREF(kc);
      } else {
        kc = jcopy_impl(k, fallback, depth + 1);
        if (!kc) { Py_DECREF(out); return nullptr; }
      }
      PyObject* c = jcopy_impl(v, fallback, depth + 1);
      if (!c || PyDict_SetItem(out, kc, c) < 0) {
        Py_XDECREF(c);
        Py_DECREF(kc);
        Py_DECREF(out);
        return nullptr;
      }
      Py_DECREF(c);
      Py_DECREF(kc);
    }
    return out;
  }
  if (PyList_CheckExact(x)) {
    const Py_ssize_t n = PyList_GET_SIZE(x);
    PyObject* out = PyList_New(n);
    if (!out) return nullptr;
    for (Py_ssize_t i = 0; i < n; ++i) {
      PyObject* c = jcopy_impl(PyList_GET_ITEM(x, i), fallback, depth + 1);
      if (!c) { Py_DECREF(out); return nullptr; }
      PyList_SET_ITEM(out, i, c);   // steals the reference
    }
    return out;
  }
  if (x == Py_None || PyUnicode_CheckExact(x) || PyLong_CheckExact(x) || PyFloat_CheckExact(x) || PyBool_Check(x) ||
      PyBytes_CheckExact(x)) {
    Py_INCREF(x);
    return x;
  }
  return PyObject_CallFunctionObjArgs(fallback, x, nullptr);
}

// n distinct loopback TCP ports that are free right now: every probe socket is bound (to port 0) before any is closed,
// so the kernel cannot hand one port out twice within a call.  One native call instead of 4 system calls per port
// from Python -- under a busy interpreter lock each of those cost the caller a switch interval (a measured 38 ms per
// port in the throughput benchmark).
std::vector<int> free_loopback_ports(int n) {
  std::vector<int> fds, ports;
  for (int i = 0; i < n; ++i) {
    int fd = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) break;
    fds.push_back(fd);
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
    a.sin_port = 0;
    socklen_t len = sizeof(a);
    if (::bind(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) != 0) continue;
    if (::getsockname(fd, reinterpret_cast<sockaddr*>(&a), &len) != 0) continue;
    ports.push_back(ntohs(a.sin_port));
  }
  for (int fd : fds) ::close(fd);
  return ports;
}

py::object store_error_type;

[[noreturn]] void raise_store_error(const StoreError& e) {
  py::object exc = store_error_type(py::str(e.what()));
  exc.attr("reason") = e.reason;
  PyErr_SetObject(store_error_type.ptr(), exc.ptr());
  throw py::error_already_set();
}

}  // namespace

PYBIND11_MODULE(_aitj_core, m) {
  m.doc() = "native control-plane core: work-queue, expectations, watchable store, process supervisor";

  store_error_type = py::reinterpret_borrow<py::object>(
      PyErr_NewException("trainingjob_operator_b200.core._aitj_core.StoreError", PyExc_RuntimeError, nullptr));
  m.attr("StoreError") = store_error_type;
  static py::exception<SpawnError> spawn_exc(m, "SpawnError", PyExc_OSError);
  py::register_exception_translator([](std::exception_ptr p) {
    try {
      if (p) std::rethrow_exception(p);
    } catch (const SpawnError& e) {
      py::object exc = py::reinterpret_steal<py::object>(
          PyObject_CallFunction(spawn_exc.ptr(), "is", e.err, e.what()));
      if (exc) PyErr_SetObject(spawn_exc.ptr(), exc.ptr());
    }
  });

  py::class_<WorkQueue>(m, "WorkQueue")
      .def(py::init<std::string, double, double, double, int>(), py::arg("name") = "queue",
           py::arg("base_delay") = 0.005, py::arg("max_delay") = 1000.0, py::arg("qps") = 10.0,
           py::arg("burst") = 100)
      .def("add", &WorkQueue::add)
      .def("add_after", &WorkQueue::add_after)
      .def("add_rate_limited", &WorkQueue::add_rate_limited)
      .def("forget", &WorkQueue::forget)
      .def("num_requeues", &WorkQueue::num_requeues)
      .def("get",
           [](WorkQueue& q, double timeout) -> std::optional<std::string> {
             // an item that is already queued is taken with the interpreter lock held; the lock is given up only to wait
             std::optional<std::string> key = q.get(0.0);
             if (key || timeout == 0.0) return key;
             py::gil_scoped_release rel;
             return q.get(timeout);
           },
           py::arg("timeout") = -1.0)
      .def("done", &WorkQueue::done)
      .def("__len__", &WorkQueue::len)
      .def("len_waiting", &WorkQueue::len_waiting)
      .def("shutdown", &WorkQueue::shutdown)
      .def("shutting_down", &WorkQueue::shutting_down)
      .def_property_readonly("name", &WorkQueue::name);

  py::class_<Expectations>(m, "Expectations")
      .def(py::init<double>(), py::arg("ttl") = 300.0)
      .def("expect_creations", &Expectations::expect_creations)
      .def("expect_deletions", &Expectations::expect_deletions)
      .def("set", &Expectations::set)
      .def("raise_expectations", &Expectations::raise)
      .def("lower_expectations", &Expectations::lower)
      .def("creation_observed", &Expectations::creation_observed)
      .def("deletion_observed", &Expectations::deletion_observed)
      .def("satisfied", &Expectations::satisfied)
      .def("delete", &Expectations::erase)
      .def("peek", &Expectations::peek);

  py::class_<Store>(m, "Store")
      .def(py::init<std::string, size_t>(), py::arg("wal_path") = "", py::arg("history") = 16384)
      .def("current_rv", &Store::current_rv)
      .def("create",
           [](Store& s, const std::string& kind, const std::string& ns, const std::string& name,
              const std::string& uid, const py::bytes& data, const Labels& labels,
              const std::vector<std::string>& owners) {
             try {
               return obj_to_dict(s.create(make_obj(kind, ns, name, uid, data, labels, owners)));
             } catch (const StoreError& e) { raise_store_error(e); }
           })
      .def("update",
           [](Store& s, const std::string& kind, const std::string& ns, const std::string& name,
              const std::string& uid, const py::bytes& data, const Labels& labels,
              const std::vector<std::string>& owners, uint64_t expected_rv) {
             try {
               return obj_to_dict(s.update(make_obj(kind, ns, name, uid, data, labels, owners), expected_rv));
             } catch (const StoreError& e) { raise_store_error(e); }
           })
      .def("get",
           [](Store& s, const std::string& kind, const std::string& ns, const std::string& name) {
             try {
               return obj_to_dict(s.get(kind, ns, name));
             } catch (const StoreError& e) { raise_store_error(e); }
           })
      .def("list",
           [](Store& s, const std::string& kind, const std::string& ns, const Labels& selector) {
             auto r = s.list(kind, ns, selector);
             py::list items;
             for (auto& o : r.first) items.append(obj_to_dict(o));
             return py::make_tuple(items, r.second);
           },
           py::arg("kind"), py::arg("namespace") = "", py::arg("selector") = Labels{})
      .def("remove",
           [](Store& s, const std::string& kind, const std::string& ns, const std::string& name,
              const std::string& expected_uid) {
             try {
               py::list out;
               for (auto& o : s.remove(kind, ns, name, expected_uid)) out.append(obj_to_dict(o));
               return out;
             } catch (const StoreError& e) { raise_store_error(e); }
           },
           py::arg("kind"), py::arg("namespace"), py::arg("name"), py::arg("expected_uid") = "")
      .def("watch_open",
           [](Store& s, const std::string& kind, const std::string& ns, uint64_t since_rv) {
             try {
               return s.watch_open(kind, ns, since_rv);
             } catch (const StoreError& e) { raise_store_error(e); }
           },
           py::arg("kind"), py::arg("namespace") = "", py::arg("since_rv") = 0)
      .def("watch_next",
           [](Store& s, int64_t id, double timeout) -> py::object {
             // queued event: taken with the interpreter lock held (giving it up for a call that returns at once costs
             // the caller a switch interval under load); the lock is released only for a real wait
             std::vector<WatchEvent> evs = s.watch_next_many(id, 0.0, 1, false);
             if (evs.empty() && timeout > 0) {
               py::gil_scoped_release rel;
               evs = s.watch_next_many(id, timeout, 1, true);
             }
             if (evs.empty()) return py::none();
             return py::make_tuple(evs[0].type, obj_to_dict(evs[0].obj));
           },
           py::arg("id"), py::arg("timeout") = 1.0)
      .def("watch_next_many",
           [](Store& s, int64_t id, double timeout, size_t max) {
             std::vector<WatchEvent> evs = s.watch_next_many(id, 0.0, max, false);
             if (evs.empty() && timeout > 0) {
               py::gil_scoped_release rel;
               evs = s.watch_next_many(id, timeout, max, true);
             }
             py::list out;
             for (auto& ev : evs) out.append(py::make_tuple(ev.type, obj_to_dict(ev.obj)));
             return out;
           },
           py::arg("id"), py::arg("timeout") = 1.0, py::arg("max") = 64,
           "up to `max` queued events; waits up to `timeout` seconds only when none is queued")
      .def("watch_close", &Store::watch_close)
      .def("num_watchers", &Store::num_watchers)
      .def("count", &Store::count)
      .def("compact", &Store::compact);

  m.def("jcopy", [](py::object x, py::object fallback) {
    PyObject* r = jcopy_impl(x.ptr(), fallback.ptr(), 0);
    if (!r) throw py::error_already_set();
    return py::reinterpret_steal<py::object>(r);
  }, py::arg("obj"), py::arg("fallback"),
        "deep copy of a JSON-shaped tree; objects other than dict / list / scalars are copied by fallback(obj)");
  // (the interpreter lock is kept on purpose: the call takes microseconds, giving the lock up would cost the caller a
  //  switch interval to get it back)
  m.def("free_loopback_ports", &free_loopback_ports, py::arg("n"),
        "n distinct loopback TCP ports that are free right now (bound to port 0 simultaneously, then closed)");
  m.def("proc_start_time", &proc_start_time, py::arg("pid"),
        "kernel start time (clock ticks since boot) of a pid, 0 if it does not exist");
  py::class_<Supervisor>(m, "Supervisor")
      .def(py::init<>())
      .def("spawn", &Supervisor::spawn, py::arg("id"), py::arg("argv"), py::arg("env"), py::arg("cwd") = "",
           py::arg("stdout_path") = "", py::arg("stderr_path") = "", py::arg("cpus") = std::vector<int>{})
      .def("kill", &Supervisor::kill_proc, py::arg("id"), py::arg("sig") = 15, py::arg("group") = true)
      .def("adopt", &Supervisor::adopt, py::arg("id"), py::arg("pid"))
      .def("rename", &Supervisor::rename, py::arg("from_id"), py::arg("to_id"))
      .def("alive", &Supervisor::alive)
      .def("pid_of", &Supervisor::pid_of)
      .def("list", &Supervisor::list)
      .def("poll_exits",
           [](Supervisor& s, double timeout) {
             std::vector<ExitEvent> evs;
             {
               py::gil_scoped_release rel;
               evs = s.poll_exits(timeout);
             }
             py::list out;
             for (auto& e : evs) {
               py::dict d;
               d["id"] = e.id;
               d["pid"] = e.pid;
               d["exit_code"] = e.exit_code;
               d["signal"] = e.signal;
               d["wall_time_s"] = e.wall_time_s;
               d["status_unknown"] = e.status_unknown;
               out.append(d);
             }
             return out;
           },
           py::arg("timeout") = 1.0);
}
