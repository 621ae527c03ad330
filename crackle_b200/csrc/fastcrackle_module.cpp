// fastcrackle_module.cpp -- pybind11 drop-in for the two hot-path functions of the reference's `fastcrackle`
// extension module, implemented above the C-ABI (include/crackle_b200.h).  Host code only; all compute happens in
// libcrackle_b200.so (CUDA, sm_100a).
//
// Mirrors src/fastcrackle.cpp:
//   compress(labels, allow_pins, fortran_order, markov_model_order, optimize_pins, auto_bgcolor, manual_bgcolor, parallel)
//        :163-210  (positional, same defaults; dtype dispatch by itemsize/kind :173-208)
//   decompress(buffer, z_start, z_end, parallel, label)    :84-129  (1-D output array; uint8 when label is given)
// Every other name of the reference module (remap, point_cloud, ...) is outside the hot path; INTEGRATION.md shows
// how the reference package re-exports them from its own build.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <optional>
#include <stdexcept>
#include <string>

#include "../../include/crackle_b200.h"

namespace py = pybind11;

static py::bytes compress(const py::array& labels, const bool allow_pins = false, const bool fortran_order = true,
                          const uint64_t markov_model_order = 0, const bool optimize_pins = false,
                          const bool auto_bgcolor = true, const int64_t manual_bgcolor = 0, const size_t parallel = 1) {
  (void)optimize_pins; (void)auto_bgcolor; (void)manual_bgcolor; (void)parallel;
  if (allow_pins) throw std::runtime_error("crackle_b200: allow_pins is outside the flat-label hot path; use the reference module");
  if (labels.dtype().kind() == 'i') throw std::runtime_error("crackle_b200: signed labels are not supported on this path");
  const int width = (int)labels.dtype().itemsize();
  const uint64_t sx = labels.ndim() < 1 ? 1 : (uint64_t)labels.shape()[0];
  const uint64_t sy = labels.ndim() < 2 ? 1 : (uint64_t)labels.shape()[1];
  const uint64_t sz = labels.ndim() < 3 ? 1 : (uint64_t)labels.shape()[2];
  uint8_t* out = nullptr;
  uint64_t n = 0;
  char err[512] = {0};
  int rc;
  const void* src = labels.data();          // every pybind11 / numpy accessor is called while the GIL is still held
  {
    py::gil_scoped_release nogil;
    rc = crackle_b200_compress(src, width, sx, sy, sz, fortran_order ? 1 : 0, (int)markov_model_order, &out, &n, err, sizeof err);
  }
  if (rc) throw std::runtime_error(err);
  py::bytes b(reinterpret_cast<const char*>(out), n);
  crackle_b200_free(out);
  return b;
}

static py::array decompress(const py::buffer buffer, int64_t z_start = 0, int64_t z_end = -1, const size_t parallel = 1,
                            const std::optional<uint64_t> label = std::nullopt) {
  (void)parallel;
  py::buffer_info info = buffer.request();
  if (info.ndim != 1) throw std::runtime_error("Expected a 1D buffer");
  const uint8_t* data = static_cast<const uint8_t*>(info.ptr);
  const uint64_t nbytes = (uint64_t)info.size * (uint64_t)info.itemsize;
  ckl_header_info h;
  char err[512] = {0};
  if (crackle_b200_header(data, nbytes, &h, err, sizeof err)) throw std::runtime_error(err);
  // same clamping as src/fastcrackle.cpp:50-60
  int64_t zs = std::max<int64_t>(z_start, 0);
  int64_t ze = z_end == -1 ? (int64_t)h.sz : z_end;
  ze = std::min<int64_t>(std::max<int64_t>(ze, 0), (int64_t)h.sz);
  const int64_t voxels = (int64_t)h.sx * (int64_t)h.sy * std::max<int64_t>(ze - zs, 0);
  py::array arr;
  if (label.has_value()) arr = py::array_t<uint8_t>(voxels);
  else if (h.data_width == 1) arr = py::array_t<uint8_t>(voxels);
  else if (h.data_width == 2) arr = py::array_t<uint16_t>(voxels);
  else if (h.data_width == 4) arr = py::array_t<uint32_t>(voxels);
  else arr = py::array_t<uint64_t>(voxels);
  int rc;
  void* dst = arr.mutable_data();
  const uint64_t dst_bytes = (uint64_t)arr.nbytes();
  {
    py::gil_scoped_release nogil;
    rc = crackle_b200_decompress(data, nbytes, z_start, z_end, label.has_value() ? 1 : 0, label.value_or(0), dst, dst_bytes, err,
                                 sizeof err);
  }
  if (rc) throw std::runtime_error(err);
  return arr;
}

PYBIND11_MODULE(fastcrackle, m) {
  m.doc() = "B200-native drop-in for fastcrackle.compress / fastcrackle.decompress (crackle hot path).";
  m.def("compress", &compress, "Compress a Fortran-ordered unsigned label array into a .ckl stream (flat labels).");
  m.def("decompress", &decompress, "Decompress a .ckl stream (optionally a z-range / single-label mask).");
  m.attr("__backend__") = "crackle_b200 (sm_100a, C-ABI)";
}
