// flash_join_py.cpp — the pybind11 module `flash_join` over the C ABI (include/flashjoin_b200.h).
//
// Mirrors PYBIND11_MODULE(flash_join, m) of /root/reference/hash_join.cpp:598-640: the same 12 join
// entry points + initialize(), the same keyword names (build_keys, build_values, probe_keys) and the
// same return value (int num_matches, float seconds).  Deliberate, documented deviations from the
// reference's undefined behaviour: non-1-D input and len(build_values) !=
// len(build_keys) raise ValueError (the reference reads out of bounds / ignores strides).
// Non-breaking additions: last_stats(), last_pairs(), configure(), get_config(), pinned_empty().
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <stdexcept>
#include <string>

#include "../../include/flashjoin_b200.h"

namespace py = pybind11;

namespace {

fj_stats g_last_stats;
bool g_have_stats = false;

[[noreturn]] void raise_status(fj_status s) {
  const std::string msg = fj_last_error();
  switch (s) {
    case FJ_ERR_BAD_ARG: throw py::value_error(msg);
    case FJ_ERR_OOM: throw std::bad_alloc();
    default: throw std::runtime_error(msg);
  }
}

// uint64 / int64 1-D C-contiguous arrays are taken zero-copy (same bits); anything else is
// converted by copy like the reference's py::array_t<uint64_t> forcecast (hash_join.cpp:316).
py::array_t<uint64_t, py::array::c_style> as_u64(const py::object& o, const char* name) {
  py::array a = py::array::ensure(o);
  if (!a) throw py::type_error(std::string(name) + " is not convertible to a numpy array");
  if (a.ndim() != 1) throw py::value_error(std::string(name) + " must be 1-D (got ndim=" + std::to_string(a.ndim()) + ")");
  if (a.dtype().is(py::dtype::of<int64_t>())) a = a.attr("view")(py::dtype::of<uint64_t>());
  auto r = py::array_t<uint64_t, py::array::c_style | py::array::forcecast>::ensure(a);
  if (!r) throw py::type_error(std::string(name) + " cannot be converted to uint64");
  return r;
}

py::tuple run(int algo, unsigned flags, const py::object& bk_o, const py::object& bv_o, const py::object& pk_o) {
  auto bk = as_u64(bk_o, "build_keys");
  auto bv = as_u64(bv_o, "build_values");
  auto pk = as_u64(pk_o, "probe_keys");
  if (bv.size() != bk.size())
    throw py::value_error("build_values has " + std::to_string(bv.size()) + " rows, build_keys has " +
                          std::to_string(bk.size()));
  uint64_t n = 0;
  double sec = 0.0;
  fj_stats st;
  fj_status s;
  {
    py::gil_scoped_release nogil;
    s = fj_join_u64(algo, flags, bk.data(), bv.data(), (size_t)bk.size(), pk.data(), (size_t)pk.size(), &n, &sec, &st);
  }
  if (s != FJ_OK) raise_status(s);
  g_last_stats = st;
  g_have_stats = true;
  return py::make_tuple(py::int_(n), sec);
}

template <int ALGO, unsigned FLAGS>
py::tuple entry(const py::object& bk, const py::object& bv, const py::object& pk) {
  return run(ALGO, FLAGS, bk, bv, pk);
}

py::dict last_stats() {
  py::dict d;
  if (!g_have_stats) return d;
  const fj_stats& s = g_last_stats;
  d["h2d_s"] = s.h2d_s; d["clear_s"] = s.clear_s; d["build_s"] = s.build_s; d["partition_s"] = s.partition_s;
  d["probe_s"] = s.probe_s; d["comm_s"] = s.comm_s; d["device_s"] = s.device_s; d["wall_s"] = s.wall_s;
  d["matches"] = s.matches; d["table_bytes"] = s.table_bytes; d["algorithmic_bytes"] = s.algorithmic_bytes;
  d["h2d_bytes"] = s.h2d_bytes;
  d["path"] = s.path == FJ_ALGO_RADIX ? "radix" : "scalar";
  d["narrow"] = (bool)s.narrow;
  d["bloom_kind"] = s.bloom_kind == 0 ? "none" : (s.bloom_kind == 1 ? "smem" : (s.bloom_kind == 2 ? "global" : (s.bloom_kind == 3 ? "bitmap" : "partition")));
  d["attempts"] = s.attempts; d["dedup_exact"] = (bool)s.dedup_exact; d["kernel_launches"] = s.kernel_launches;
  d["radix_bits"] = py::make_tuple(s.radix_bits1, s.radix_bits2);
  d["n_gpus"] = s.n_gpus;
  d["dense"] = s.dense;
  d["part_build_us"] = s.part_build_us;
  d["part_probe_us"] = s.part_probe_us;
  return d;
}

py::object last_pairs(bool with_probe_idx) {
  uint64_t n = 0;
  fj_status s = fj_pairs_count(&n);
  if (s != FJ_OK) raise_status(s);
  py::array_t<uint64_t> k((py::ssize_t)n), v((py::ssize_t)n), ix((py::ssize_t)(with_probe_idx ? n : 0));
  {
    py::gil_scoped_release nogil;
    s = fj_pairs_fetch(k.mutable_data(), v.mutable_data(), with_probe_idx ? ix.mutable_data() : nullptr, (size_t)n);
  }
  if (s != FJ_OK) raise_status(s);
  if (with_probe_idx) return py::make_tuple(k, v, ix);
  return py::make_tuple(k, v);
}

void configure(const py::kwargs& kw) {
  for (auto item : kw) {
    const std::string key = py::cast<std::string>(item.first);
    const int64_t val = py::cast<int64_t>(item.second);
    fj_status s = fj_config_set(key.c_str(), val);
    if (s != FJ_OK) raise_status(s);
  }
}
int64_t get_config(const std::string& key) {
  int64_t v = 0;
  fj_status s = fj_config_get(key.c_str(), &v);
  if (s != FJ_OK) raise_status(s);
  return v;
}

// a uint64 numpy array backed by page-locked host memory (H2D at full PCIe rate, async capable)
py::array_t<uint64_t> pinned_empty(size_t n) {
  void* p = nullptr;
  fj_status s = fj_host_alloc_pinned(&p, n * sizeof(uint64_t));
  if (s != FJ_OK) raise_status(s);
  py::capsule owner(p, [](void* q) { fj_host_free_pinned(q); });
  return py::array_t<uint64_t>({(py::ssize_t)n}, {(py::ssize_t)sizeof(uint64_t)}, static_cast<uint64_t*>(p), owner);
}

void initialize() {
  fj_status s = fj_init(-1);
  if (s != FJ_OK) raise_status(s);
}

constexpr unsigned B = FJ_FLAG_BLOOM, M = FJ_FLAG_MATERIALIZE;

}  // namespace

PYBIND11_MODULE(flash_join, m) {
  m.doc() = "B200-native hash join library with adaptive and explicit strategies (flash_join API).";
  auto a = [](const char* n) { return py::arg(n); };
#define FJ_DEF(name, ALGO, FLAGS, doc) \
  m.def(name, &entry<ALGO, FLAGS>, doc, a("build_keys"), a("build_values"), a("probe_keys"))
  // adaptive API (hash_join.cpp:603-617)
  FJ_DEF("adaptive_join", FJ_ALGO_ADAPTIVE, M, "Adaptively chooses between the global-table and radix join; materializes pairs.");
  FJ_DEF("adaptive_join_bloom", FJ_ALGO_ADAPTIVE, M | B, "Adaptive join with Bloom filter; materializes pairs.");
  FJ_DEF("adaptive_join_count", FJ_ALGO_ADAPTIVE, 0u, "Adaptively chooses between the global-table and radix join; counts.");
  FJ_DEF("adaptive_join_count_bloom", FJ_ALGO_ADAPTIVE, B, "Adaptive join with Bloom filter; counts.");
  // explicit APIs (hash_join.cpp:621-637)
  FJ_DEF("hash_join_radix", FJ_ALGO_RADIX, M, "Forces the radix join; materializes pairs.");
  FJ_DEF("hash_join", FJ_ALGO_SCALAR, M, "Forces the non-partitioned (global table) join; materializes pairs.");
  FJ_DEF("hash_join_radix_bloom", FJ_ALGO_RADIX, M | B, "");
  FJ_DEF("hash_join_bloom", FJ_ALGO_SCALAR, M | B, "");
  FJ_DEF("hash_join_count_radix", FJ_ALGO_RADIX, 0u, "Forces the radix join; counts.");
  FJ_DEF("hash_join_count", FJ_ALGO_SCALAR, 0u, "Forces the non-partitioned (global table) join; counts.");
  FJ_DEF("hash_join_count_radix_bloom", FJ_ALGO_RADIX, B, "");
  FJ_DEF("hash_join_count_bloom", FJ_ALGO_SCALAR, B, "");
#undef FJ_DEF
  m.def("initialize", &initialize, "Creates the CUDA context, stream and device arena (hash_join.cpp:639).");
  // additions
  m.def("last_stats", &last_stats, "Statistics of the last join call (fj_stats).");
  m.def("last_pairs", &last_pairs, py::arg("with_probe_idx") = false,
        "(probe_keys, build_values[, probe_idx]) materialized by the last materialize call.");
  m.def("configure", &configure, "Set engine tunables, e.g. configure(load_pct=50, smem_bloom=1).");
  m.def("get_config", &get_config, py::arg("key"));
  m.def("pinned_empty", &pinned_empty, py::arg("n"), "uint64 array in page-locked host memory.");
  m.def("version", []() { return std::string(fj_version()); });
  m.def("join_flags", [](const std::string& algo, bool bloom, bool materialize, const py::object& bk, const py::object& bv,
                         const py::object& pk, bool force_wide, bool probe_idx) {
          int al = algo == "adaptive" ? FJ_ALGO_ADAPTIVE : algo == "scalar" ? FJ_ALGO_SCALAR : algo == "radix" ? FJ_ALGO_RADIX : -1;
          if (al < 0) throw py::value_error("algo must be adaptive|scalar|radix");
          unsigned f = (bloom ? B : 0u) | (materialize ? M : 0u) | (force_wide ? FJ_FLAG_FORCE_WIDE : 0u) |
                       (probe_idx ? FJ_FLAG_PROBE_IDX : 0u);
          return run(al, f, bk, bv, pk);
        },
        py::arg("algo"), py::arg("bloom"), py::arg("materialize"), py::arg("build_keys"), py::arg("build_values"),
        py::arg("probe_keys"), py::arg("force_wide") = false, py::arg("probe_idx") = false,
        "Generic entry point exposing the extra C-ABI flags (force_wide, probe_idx).");
}
