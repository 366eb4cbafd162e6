// Drives r/src/harmony_shim.cpp against the Rcpp TEST STUB (tests/stubs/Rcpp.h): registers the module (every
// .property / .method pointer must type-check), prints the exposed names, then constructs `harmony`.
// Exit 3 = the constructor stopped with the library's "no usable CUDA device" (expected in an image without a GPU).
#include <cstdio>
#include "harmony_shim.cpp"

int main() {
  rcpp_stub_module_harmony_module();
  for (const std::string& n : Rcpp::class_<harmony>::names()) std::printf("%s\n", n.c_str());
  try {
    harmony h;
    std::printf("constructed; N = %d\n", h.get_N());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "stop: %s\n", e.what());
    return 3;
  }
  return 0;
}
