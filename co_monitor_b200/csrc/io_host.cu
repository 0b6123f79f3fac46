// Host-side writer of the Stage-1 file for 1e7-1e8 row results.
//
// The reference writes its result with dask/pandas ``to_csv(sep=";", header=True, index=False)``
// (co_monitor/geometry/simulation.py:616-630): one header line, then per visible voxel
//     idx_sel_plas_points;plasma_x;plasma_y;plasma_z;total_intensity_fraction
// with integers in decimal and floats in Python's shortest round-trip ``repr`` form (e.g. 2.188959843266816e-06,
// -1234.5, 0.0001, 1e-05).  pandas formats about 1e6 such rows per second; a 1 mm grid has 3-5e7 visible voxels per
// channel.  This writer produces the same bytes, formatting row blocks on all host cores.
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "scene.h"

namespace com {

// Python float.__repr__: shortest digits that round-trip; fixed notation when -4 <= exponent10 < 16, else scientific
// with a sign and at least two exponent digits; integers keep a trailing ".0".
static char* format_repr(char* out, double v) {
  if (std::isnan(v)) {
    std::memcpy(out, "nan", 3);
    return out + 3;
  }
  if (std::isinf(v)) {
    if (v < 0) *out++ = '-';
    std::memcpy(out, "inf", 3);
    return out + 3;
  }
  char sci[40];
  const auto r = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);  // d[.ddd]e[+-]XX, shortest
  const char* p = sci;
  if (*p == '-') *out++ = *p++;
  char digits[24];
  int nd = 0;
  digits[nd++] = *p++;
  if (*p == '.') {
    ++p;
    while (*p != 'e') digits[nd++] = *p++;
  }
  ++p;  // 'e'
  int e10 = 0;
  const bool eneg = *p == '-';
  ++p;
  for (; p < r.ptr; ++p) e10 = e10 * 10 + (*p - '0');
  if (eneg) e10 = -e10;
  if (v == 0.0) e10 = 0;
  if (e10 < -4 || e10 >= 16) {
    *out++ = digits[0];
    if (nd > 1) {
      *out++ = '.';
      std::memcpy(out, digits + 1, nd - 1);
      out += nd - 1;
    }
    *out++ = 'e';
    *out++ = e10 < 0 ? '-' : '+';
    int a = e10 < 0 ? -e10 : e10;
    if (a >= 100) *out++ = (char)('0' + a / 100);
    *out++ = (char)('0' + (a / 10) % 10);
    *out++ = (char)('0' + a % 10);
    return out;
  }
  if (e10 < 0) {  // 0.000ddd
    *out++ = '0';
    *out++ = '.';
    for (int i = 0; i < -e10 - 1; ++i) *out++ = '0';
    std::memcpy(out, digits, nd);
    return out + nd;
  }
  // integer part has e10 + 1 digits
  for (int i = 0; i <= e10; ++i) *out++ = i < nd ? digits[i] : '0';
  *out++ = '.';
  if (nd > e10 + 1) {
    std::memcpy(out, digits + e10 + 1, nd - e10 - 1);
    return out + (nd - e10 - 1);
  }
  *out++ = '0';
  return out;
}

static void format_rows(std::string& buf, const int64_t* idx, const double* x, const double* y, const double* z,
                        const double* w, int64_t lo, int64_t hi) {
  buf.resize((size_t)(hi - lo) * 128);
  char* o = buf.data();
  for (int64_t i = lo; i < hi; ++i) {
    o = std::to_chars(o, o + 24, idx[i]).ptr;
    *o++ = ';';
    o = format_repr(o, x[i]);
    *o++ = ';';
    o = format_repr(o, y[i]);
    *o++ = ';';
    o = format_repr(o, z[i]);
    *o++ = ';';
    o = format_repr(o, w[i]);
    *o++ = '\n';
  }
  buf.resize((size_t)(o - buf.data()));
}

}  // namespace com

extern "C" {

int com_format_float_repr(double value, char* out, int32_t out_bytes) {
  char tmp[48];
  char* e = com::format_repr(tmp, value);
  const int n = (int)(e - tmp);
  if (!out || out_bytes <= n) return COM_E_BADARG;
  std::memcpy(out, tmp, n);
  out[n] = 0;
  return n;
}

int com_write_stage1_csv(const char* path, const int64_t* idx, const double* x, const double* y, const double* z,
                         const double* w, int64_t n, int32_t write_header, int32_t n_threads) {
  using namespace com;
  if (!path || n < 0 || (n > 0 && (!idx || !x || !y || !z || !w))) {
    set_error("com_write_stage1_csv: bad arguments");
    return COM_E_BADARG;
  }
  FILE* f = std::fopen(path, "wb");
  if (!f) {
    set_error("com_write_stage1_csv: cannot open %s", path);
    return COM_E_IO;
  }
  bool ok = true;
  if (write_header) {
    static const char header[] = "idx_sel_plas_points;plasma_x;plasma_y;plasma_z;total_intensity_fraction\n";
    ok = std::fwrite(header, 1, sizeof header - 1, f) == sizeof header - 1;
  }
  int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  if (T > 64) T = 64;
  const int64_t block = 1 << 16;  // rows per formatted block
  const int64_t n_blocks = (n + block - 1) / block;
  // rounds of T blocks: formatted in parallel, written in order
  std::vector<std::string> bufs((size_t)T);
  for (int64_t b0 = 0; b0 < n_blocks && ok; b0 += T) {
    const int64_t nb = std::min<int64_t>(T, n_blocks - b0);
    std::vector<std::thread> th;
    for (int64_t t = 1; t < nb; ++t)
      th.emplace_back([&, t] {
        const int64_t lo = (b0 + t) * block;
        format_rows(bufs[(size_t)t], idx, x, y, z, w, lo, std::min(n, lo + block));
      });
    format_rows(bufs[0], idx, x, y, z, w, b0 * block, std::min(n, (b0 + 1) * block));
    for (auto& t : th) t.join();
    for (int64_t t = 0; t < nb && ok; ++t)
      ok = std::fwrite(bufs[(size_t)t].data(), 1, bufs[(size_t)t].size(), f) == bufs[(size_t)t].size();
  }
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) {
    set_error("com_write_stage1_csv: write to %s failed", path);
    return COM_E_IO;
  }
  return COM_OK;
}

}  // extern "C"
