// json_io.cpp -- trajectory files with the reference's keys, for one problem or a batch.
//
// The reference reads its tracking reference from test/scotty.json (ReadScottyTrajectory,
// test/test_utils.cpp:240-289: keys "N", "tf", "state_trajectory" [knots][n], "input_trajectory"
// [knots][m]) and writes the closed-loop MPC run to test/scotty_mpc.json (test/bicycle_test.cpp:
// 344-359: the same keys plus "solve_iters" [steps] and "tracking_error" [steps], std::setw(4)).
// This is the same wire format behind the C ABI (include/altro_b200.h, section D), extended to B
// problems: a batch file carries "batch": B and one more leading dimension on every array.  A file
// without "batch" is a single problem, so the reference's own files load unchanged.
//
// Host-only code: a small recursive-descent JSON reader (objects, arrays, numbers, strings,
// true/false/null -- everything nlohmann::json would accept in these files) and a writer.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/altro_b200.h"

namespace {

struct JValue {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  double num = 0.0;
  bool b = false;
  std::string str;
  std::vector<JValue> arr;
  std::map<std::string, JValue> obj;
};

struct Parser {
  const char* p;
  const char* end;
  bool ok = true;

  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  bool eat(char c) {
    ws();
    if (p < end && *p == c) {
      ++p;
      return true;
    }
    return false;
  }
  JValue value() {
    JValue v;
    ws();
    if (p >= end) {
      ok = false;
      return v;
    }
    const char c = *p;
    if (c == '{') {
      ++p;
      v.kind = JValue::Object;
      if (eat('}')) return v;
      do {
        ws();
        JValue key = value();
        if (!ok || key.kind != JValue::String || !eat(':')) {
          ok = false;
          return v;
        }
        v.obj[key.str] = value();
        if (!ok) return v;
      } while (eat(','));
      if (!eat('}')) ok = false;
    } else if (c == '[') {
      ++p;
      v.kind = JValue::Array;
      if (eat(']')) return v;
      do {
        v.arr.push_back(value());
        if (!ok) return v;
      } while (eat(','));
      if (!eat(']')) ok = false;
    } else if (c == '"') {
      ++p;
      v.kind = JValue::String;
      while (p < end && *p != '"') {
        if (*p == '\\' && p + 1 < end) {
          ++p;
          switch (*p) {
            case 'n': v.str.push_back('\n'); break;
            case 't': v.str.push_back('\t'); break;
            case 'r': v.str.push_back('\r'); break;
            case 'b': v.str.push_back('\b'); break;
            case 'f': v.str.push_back('\f'); break;
            case 'u': p += (end - p > 4) ? 4 : 0; v.str.push_back('?'); break;
            default: v.str.push_back(*p);
          }
          ++p;
        } else {
          v.str.push_back(*p++);
        }
      }
      if (p >= end) ok = false; else ++p;
    } else if (!std::strncmp(p, "true", 4) && end - p >= 4) {
      v.kind = JValue::Bool;
      v.b = true;
      p += 4;
    } else if (!std::strncmp(p, "false", 5) && end - p >= 5) {
      v.kind = JValue::Bool;
      p += 5;
    } else if (!std::strncmp(p, "null", 4) && end - p >= 4) {
      p += 4;
    } else {
      char* q = nullptr;
      v.num = std::strtod(p, &q);  // round-trips the 17 significant digits the writer emits
      if (q == p) {
        ok = false;
        return v;
      }
      v.kind = JValue::Number;
      p = q;
    }
    return v;
  }
};

// depth of nested arrays and the extent of each level (ragged input -> false)
bool shape_of(const JValue& v, std::vector<int>& shape, size_t level = 0) {
  if (v.kind != JValue::Array) return v.kind == JValue::Number && level == shape.size();
  if (level == shape.size()) shape.push_back((int)v.arr.size());
  if (shape[level] != (int)v.arr.size()) return false;
  for (const JValue& e : v.arr)
    if (!shape_of(e, shape, level + 1)) return false;
  return true;
}
void flatten(const JValue& v, std::vector<double>& out) {
  if (v.kind == JValue::Number) out.push_back(v.num);
  for (const JValue& e : v.arr) flatten(e, out);
}

}  // namespace

struct altro_b200_traj_file {
  int batch = 1, N = 0;
  float tf = 0.f;
  int knots_x = 0, n = 0, knots_u = 0, m = 0, steps = 0;
  std::vector<double> X, U, err;
  std::vector<double> iters;
};

extern "C" {

altro_b200_traj_file* altro_b200_traj_open(const char* path, int* err) {
  auto fail = [&](int code) -> altro_b200_traj_file* {
    if (err) *err = code;
    return nullptr;
  };
  if (!path) return fail(ALTRO_B200_INVALID_POINTER);
  FILE* fp = std::fopen(path, "rb");
  if (!fp) return fail(ALTRO_B200_FILE_ERROR);
  std::string text;
  char buf[1 << 16];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, got);
  std::fclose(fp);
  Parser ps{text.data(), text.data() + text.size()};
  JValue root = ps.value();
  if (!ps.ok || root.kind != JValue::Object) return fail(ALTRO_B200_FILE_ERROR);
  std::unique_ptr<altro_b200_traj_file> f(new altro_b200_traj_file());
  auto num = [&](const char* key, double dflt) {
    auto it = root.obj.find(key);
    return (it != root.obj.end() && it->second.kind == JValue::Number) ? it->second.num : dflt;
  };
  const bool is_batch = root.obj.count("batch") > 0;
  f->batch = is_batch ? (int)num("batch", 1) : 1;
  f->N = (int)num("N", 0);
  f->tf = (float)num("tf", 0.0);
  if (f->batch <= 0) return fail(ALTRO_B200_FILE_ERROR);
  // [knots][dim] for one problem, [batch][knots][dim] for a batch file
  auto matrix = [&](const char* key, std::vector<double>& out, int* knots, int* dim) {
    auto it = root.obj.find(key);
    if (it == root.obj.end()) return true;
    std::vector<int> shape;
    if (!shape_of(it->second, shape)) return false;
    const size_t lead = is_batch ? 1 : 0;
    if (shape.size() != 2 + lead) return shape.size() == 1 + lead && shape.back() == 0;
    if (is_batch && shape[0] != f->batch) return false;
    *knots = shape[lead];
    *dim = shape[lead + 1];
    flatten(it->second, out);
    return true;
  };
  auto series = [&](const char* key, std::vector<double>& out) {
    auto it = root.obj.find(key);
    if (it == root.obj.end()) return true;
    std::vector<int> shape;
    if (!shape_of(it->second, shape)) return false;
    const size_t lead = is_batch ? 1 : 0;
    if (shape.size() != 1 + lead) return false;
    if (is_batch && shape[0] != f->batch) return false;
    if (f->steps && f->steps != shape[lead]) return false;
    f->steps = shape[lead];
    flatten(it->second, out);
    return true;
  };
  if (!matrix("state_trajectory", f->X, &f->knots_x, &f->n)) return fail(ALTRO_B200_DIMENSION_MISMATCH);
  if (!matrix("input_trajectory", f->U, &f->knots_u, &f->m)) return fail(ALTRO_B200_DIMENSION_MISMATCH);
  if (!series("solve_iters", f->iters)) return fail(ALTRO_B200_DIMENSION_MISMATCH);
  if (!series("tracking_error", f->err)) return fail(ALTRO_B200_DIMENSION_MISMATCH);
  if (err) *err = ALTRO_B200_NO_ERROR;
  return f.release();
}

int altro_b200_traj_dims(const altro_b200_traj_file* f, int* batch, int* N, float* tf, int* knots_x,
                         int* n, int* knots_u, int* m, int* steps) {
  if (!f) return ALTRO_B200_INVALID_POINTER;
  if (batch) *batch = f->batch;
  if (N) *N = f->N;
  if (tf) *tf = f->tf;
  if (knots_x) *knots_x = f->knots_x;
  if (n) *n = f->n;
  if (knots_u) *knots_u = f->knots_u;
  if (m) *m = f->m;
  if (steps) *steps = f->steps;
  return ALTRO_B200_NO_ERROR;
}

int altro_b200_traj_read(const altro_b200_traj_file* f, double* X, double* U, int* solve_iters,
                         double* tracking_error) {
  if (!f) return ALTRO_B200_INVALID_POINTER;
  if (X) std::memcpy(X, f->X.data(), sizeof(double) * f->X.size());
  if (U) std::memcpy(U, f->U.data(), sizeof(double) * f->U.size());
  if (solve_iters)
    for (size_t i = 0; i < f->iters.size(); ++i) solve_iters[i] = (int)std::lround(f->iters[i]);
  if (tracking_error) std::memcpy(tracking_error, f->err.data(), sizeof(double) * f->err.size());
  return ALTRO_B200_NO_ERROR;
}

void altro_b200_traj_close(altro_b200_traj_file* f) { delete f; }

int altro_b200_traj_write(const char* path, int batch, int N, float tf, int knots_x, int n,
                          const double* X, int knots_u, int m, const double* U, int steps,
                          const int* solve_iters, const double* tracking_error) {
  if (!path || !X || !U) return ALTRO_B200_INVALID_POINTER;
  if (batch < 0 || knots_x < 0 || knots_u < 0 || n <= 0 || m <= 0) return ALTRO_B200_DIMENSION_MISMATCH;
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return ALTRO_B200_FILE_ERROR;
  const bool is_batch = batch > 0;  // batch == 0: the reference's single-problem layout
  const int B = is_batch ? batch : 1;
  auto matrix = [&](const char* key, const double* A, int knots, int dim) {
    std::fprintf(fp, "    \"%s\": [", key);
    for (int b = 0; b < B; ++b) {
      if (is_batch) std::fprintf(fp, "%s\n        [", b ? "," : "");
      for (int k = 0; k < knots; ++k) {
        std::fprintf(fp, "%s\n            [", k ? "," : "");
        for (int i = 0; i < dim; ++i)
          std::fprintf(fp, "%s%.17g", i ? ", " : "", A[((size_t)b * knots + k) * dim + i]);
        std::fprintf(fp, "]");
      }
      if (is_batch) std::fprintf(fp, "\n        ]");
    }
    std::fprintf(fp, "\n    ]");
  };
  std::fprintf(fp, "{\n    \"N\": %d,\n", N);
  if (is_batch) std::fprintf(fp, "    \"batch\": %d,\n", batch);
  matrix("input_trajectory", U, knots_u, m);
  if (solve_iters && steps > 0) {
    std::fprintf(fp, ",\n    \"solve_iters\": [");
    for (int b = 0; b < B; ++b) {
      if (is_batch) std::fprintf(fp, "%s[", b ? ", " : "");
      for (int i = 0; i < steps; ++i) std::fprintf(fp, "%s%d", i ? ", " : "", solve_iters[(size_t)b * steps + i]);
      if (is_batch) std::fprintf(fp, "]");
    }
    std::fprintf(fp, "]");
  }
  std::fprintf(fp, ",\n");
  matrix("state_trajectory", X, knots_x, n);
  std::fprintf(fp, ",\n    \"tf\": %.9g", (double)tf);
  if (tracking_error && steps > 0) {
    std::fprintf(fp, ",\n    \"tracking_error\": [");
    for (int b = 0; b < B; ++b) {
      if (is_batch) std::fprintf(fp, "%s[", b ? ", " : "");
      for (int i = 0; i < steps; ++i)
        std::fprintf(fp, "%s%.17g", i ? ", " : "", tracking_error[(size_t)b * steps + i]);
      if (is_batch) std::fprintf(fp, "]");
    }
    std::fprintf(fp, "]");
  }
  std::fprintf(fp, "\n}\n");
  const bool bad = std::ferror(fp) != 0;
  if (std::fclose(fp) != 0 || bad) return ALTRO_B200_FILE_ERROR;
  return ALTRO_B200_NO_ERROR;
}

}  // extern "C"
