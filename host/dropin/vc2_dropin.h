// Shared by the drop-in Library bodies (host/dropin/*.cpp): the process-wide GPU context and the mapping of C-ABI
// status codes back to the exceptions the reference's callers expect.
//
// These three files are what a maintainer of bbc/vc2-reference would put in place of src/Library/src/
// {WaveletTransform,Quantisation,Slices}.cpp: they include the reference's OWN headers (they are compiled with
// -I<reference>/src/Library) and implement the functions declared there over include/vc2_cabi.h, so that the
// reference's unmodified EncodeStream.cpp / DecodeStream.cpp link against the CUDA library.  Every array-valued
// function runs on the GPU through the C-ABI; there is no CPU path for them (vc2_create fails without a device).
// Scalar int -> int helpers (quant, scale, adjust_quant_index, predictDC, paddedSize ...) are plain host code.
#ifndef VC2_DROPIN_H
#define VC2_DROPIN_H
#include <cstdlib>
#include <stdexcept>
#include <string>
#include "vc2_cabi.h"

namespace vc2dropin {

inline vc2_ctx* ctx() {
  static vc2_ctx* c = nullptr;
  if (!c) {
    const char* dev = std::getenv("VC2_DEVICE");
    c = vc2_create(dev ? std::atoi(dev) : 0);
    if (!c) throw std::runtime_error("vc2: no usable CUDA device (this Library has no CPU fallback)");
  }
  return c;
}

// the C-ABI keeps the reference's exception texts (vc2_status_message); argument errors are invalid_argument
inline void check(int status) {
  if (status == VC2_OK) return;
  const std::string msg = vc2_last_error(ctx());
  if (status == VC2_ERR_ARG) throw std::invalid_argument(msg.empty() ? "vc2: invalid argument" : msg);
  throw std::logic_error(msg.empty() ? vc2_status_message(status) : msg);
}

}  // namespace vc2dropin
#endif
