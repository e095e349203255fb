// vmk_io.hpp -- the callers' snapshot format, host code.
//
// The reference scripts dump fields as text, one grid node per line (vm.jl:81-85,132-136,142-146, hybrid.jl:79-83,
// pseudospectral_23_rule.jl:78-82, ...):
//     for j in 1:ny+1, i in 1:nx+1:  write(io, "$(x[i]) $(y[j]) $(ut[i, j])\n")
// i.e. j outer / i inner, three Float64 per line printed by Julia's `print(::Float64)`, and plotting.jl:14-28 reads
// them back with `readdlm` (whitespace-separated columns) and reshapes column 3 to (nx+1, ny+1).
//
// Julia prints a Float64 with the shortest digit string that round-trips (Ryu; base/ryu/shortest.jl `writeshortest`
// with hash = true, precision = -1), laid out as follows, with d = the decimal digits, n = their count and pt = the
// position of the decimal point relative to the first digit (value = 0.d1d2... x 10^pt):
//     -4 < pt <= 6 : positional --  pt <= 0: "0." + (-pt zeros) + d;  0 < pt < n: d with '.' after pt digits;
//                                   pt >= n: d + (pt - n zeros) + ".0"
//     otherwise    : d1 + "." + (d2.. or "0") + "e" + (pt - 1)   -- no '+', no zero padding of the exponent
//     specials     : "NaN", "Inf", "-Inf"; zeros "0.0" / "-0.0"
// (C's "%.17g" and Python's repr differ from this in the exponent form: "1e-05" vs Julia "1.0e-5", the switch-over at
// 1e16 instead of 1e6.)  std::to_chars(scientific) supplies the same shortest round-trip digits (Ryu-equivalent).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <charconv>
#include <string>
#include <vector>

namespace vmk {

// writes Julia's print(::Float64) of v into buf (at least 32 bytes), returns the length (no terminator)
inline int julia_print_f64(double v, char* buf) {
  if (isnan(v)) {
    memcpy(buf, "NaN", 3);
    return 3;
  }
  char* o = buf;
  if (signbit(v)) {
    *o++ = '-';
    v = -v;
  }
  if (isinf(v)) {
    memcpy(o, "Inf", 3);
    return (int)(o + 3 - buf);
  }
  if (v == 0.0) {
    memcpy(o, "0.0", 3);
    return (int)(o + 3 - buf);
  }
  char sci[40];
  const auto r = std::to_chars(sci, sci + sizeof(sci) - 1, v, std::chars_format::scientific);  // d[.ddd]e[+-]XX, shortest
  *r.ptr = 0;
  char digits[24];
  int n = 0;
  const char* p = sci;
  for (; p < r.ptr && *p != 'e'; p++)
    if (*p != '.') digits[n++] = *p;
  const int e10 = (int)strtol(p + 1, nullptr, 10);
  const int pt = e10 + 1;
  if (-4 < pt && pt <= 6) {
    if (pt <= 0) {
      *o++ = '0';
      *o++ = '.';
      for (int i = 0; i < -pt; i++) *o++ = '0';
      memcpy(o, digits, n);
      o += n;
    } else if (pt < n) {
      memcpy(o, digits, pt);
      o += pt;
      *o++ = '.';
      memcpy(o, digits + pt, n - pt);
      o += n - pt;
    } else {
      memcpy(o, digits, n);
      o += n;
      for (int i = 0; i < pt - n; i++) *o++ = '0';
      *o++ = '.';
      *o++ = '0';
    }
  } else {
    *o++ = digits[0];
    *o++ = '.';
    if (n > 1) {
      memcpy(o, digits + 1, n - 1);
      o += n - 1;
    } else {
      *o++ = '0';
    }
    *o++ = 'e';
    o += snprintf(o, 8, "%d", pt - 1);
  }
  return (int)(o - buf);
}

// "x[i] y[j] ut[i,j]\n" for j outer, i inner; ut is column-major nx1 x ny1.  Coordinates are formatted once.
inline int write_field_text(const char* path, const double* x, const double* y, const double* ut, int64_t nx1,
                            int64_t ny1, std::string* err) {
  FILE* f = fopen(path, "wb");
  if (!f) {
    *err = std::string("cannot open ") + path + " for writing";
    return 1;
  }
  std::vector<std::string> xs((size_t)nx1);
  char tmp[40];
  for (int64_t i = 0; i < nx1; i++) xs[(size_t)i].assign(tmp, (size_t)julia_print_f64(x[i], tmp));
  std::vector<char> line;
  line.reserve((size_t)nx1 * 80);
  bool ok = true;
  for (int64_t j = 0; j < ny1 && ok; j++) {
    const int ylen = julia_print_f64(y[j], tmp);
    const std::string ys(tmp, (size_t)ylen);
    line.clear();
    for (int64_t i = 0; i < nx1; i++) {
      const std::string& xi = xs[(size_t)i];
      line.insert(line.end(), xi.begin(), xi.end());
      line.push_back(' ');
      line.insert(line.end(), ys.begin(), ys.end());
      line.push_back(' ');
      const int l = julia_print_f64(ut[i + nx1 * j], tmp);
      line.insert(line.end(), tmp, tmp + l);
      line.push_back('\n');
    }
    ok = fwrite(line.data(), 1, line.size(), f) == line.size();
  }
  if (fclose(f) != 0) ok = false;
  if (!ok) {
    *err = std::string("write error on ") + path;
    return 1;
  }
  return 0;
}

// plotting.jl:14-28 (`readdlm`): three whitespace-separated Float64 columns, one row per line; at most `cap` rows are
// stored (x, y, w may each be NULL), *nrows receives the number of rows in the file
inline int read_field_text(const char* path, double* x, double* y, double* w, int64_t cap, int64_t* nrows,
                           std::string* err) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    *err = std::string("cannot open ") + path;
    return 1;
  }
  char buf[512];
  int64_t n = 0;
  while (fgets(buf, sizeof(buf), f)) {
    char* p = buf;
    while (*p == ' ' || *p == '\t') p++;
    if (*p == '\n' || *p == '\r' || *p == 0) continue;  // readdlm skips blank lines
    double v[3];
    for (int c = 0; c < 3; c++) {
      char* end = nullptr;
      v[c] = strtod(p, &end);  // accepts Julia's forms incl. "NaN", "Inf", "-Inf"
      if (end == p) {
        fclose(f);
        *err = std::string(path) + ": line " + std::to_string((long long)n + 1) + ": expected three numbers";
        return 1;
      }
      p = end;
    }
    if (n < cap) {
      if (x) x[n] = v[0];
      if (y) y[n] = v[1];
      if (w) w[n] = v[2];
    }
    n++;
  }
  fclose(f);
  if (nrows) *nrows = n;
  return 0;
}

}  // namespace vmk
