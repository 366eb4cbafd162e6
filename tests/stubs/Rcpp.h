// Rcpp.h — TEST STUB, not Rcpp.  The smallest API-shaped stand-in for the parts of Rcpp that
// r/src/harmony_shim.cpp uses, so that the shim can be type-checked against include/harmony_b200.h, linked to the
// in-tree library and driven up to its first library call in an image without R (tests/test_host_cpu.py).
// It says nothing about R's memory model or Rcpp modules' dispatch; building the real module needs R + Rcpp.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

typedef std::ptrdiff_t R_xlen_t;

namespace R {
inline double unif_rand() { return 0.5; }
}  // namespace R

namespace Rcpp {

template <class T>
class Vector {
 public:
  Vector() {}
  explicit Vector(R_xlen_t n) : v_((size_t)n) {}
  Vector(std::initializer_list<T> l) : v_(l) {}
  R_xlen_t size() const { return (R_xlen_t)v_.size(); }
  T* begin() { return v_.data(); }
  const T* begin() const { return v_.data(); }
  T& operator[](R_xlen_t i) { return v_[(size_t)i]; }
  const T& operator[](R_xlen_t i) const { return v_[(size_t)i]; }

 private:
  std::vector<T> v_;
};
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;

class NumericMatrix {
 public:
  NumericMatrix() {}
  NumericMatrix(int nr, int nc) : nr_(nr), nc_(nc), v_((size_t)nr * (size_t)nc) {}
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
  double* begin() { return v_.data(); }
  const double* begin() const { return v_.data(); }

 private:
  int nr_ = 0, nc_ = 0;
  std::vector<double> v_;
};

class S4 {  // a dgCMatrix as far as the shim looks at it: integer slots "i" and "p"
 public:
  IntegerVector slot(const std::string& name) const { return slots_.at(name); }
  void set_slot(const std::string& name, const IntegerVector& v) { slots_[name] = v; }

 private:
  std::map<std::string, IntegerVector> slots_;
};

inline std::vector<std::string>& stub_warnings() {
  static std::vector<std::string> w;
  return w;
}
// warning / stop take a printf-style format like Rcpp's (tinyformat); the stub knows "%s" with one argument only
inline void warning(const char* msg) { stub_warnings().push_back(msg); }
inline void warning(const char* fmt, const char* arg) { stub_warnings().push_back(std::string(fmt) == "%s" ? arg : fmt); }
[[noreturn]] inline void stop(const char* msg) { throw std::runtime_error(msg); }
[[noreturn]] inline void stop(const char* fmt, const char* arg) { throw std::runtime_error(std::string(fmt) == "%s" ? arg : fmt); }
inline void checkUserInterrupt() {}
struct RNGScope {
  RNGScope() {}
  ~RNGScope() {}
};

// module registration: accepts what Rcpp's class_<T> accepts for the calls the shim makes, records the names
template <class T>
class class_ {
 public:
  explicit class_(const char* name) { names().push_back(std::string("class ") + name); }
  class_& constructor() { return *this; }
  template <class R>
  class_& property(const char* n, R (T::*)()) {
    names().push_back(std::string("property ") + n);
    return *this;
  }
  template <class P>  // like Rcpp: the setter takes the getter's PROP by value
  class_& property(const char* n, P (T::*)(), void (T::*)(P)) {
    names().push_back(std::string("property(rw) ") + n);
    return *this;
  }
  template <class R, class... A>
  class_& method(const char* n, R (T::*)(A...)) {
    names().push_back(std::string("method ") + n);
    return *this;
  }
  static std::vector<std::string>& names() {
    static std::vector<std::string> v;
    return v;
  }
};

}  // namespace Rcpp

#define RCPP_EXPOSED_CLASS(x)
#define RCPP_MODULE(name) void rcpp_stub_module_##name()
