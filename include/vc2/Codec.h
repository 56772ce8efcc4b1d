// vc2/Codec.h - RAII C++ face of the fused, batched picture codec (vc2_codec_* in include/vc2_cabi.h):
// raw planar picture bytes <-> HQ slice payloads for a batch of pictures per call, host buffers in and out.
// It is what the drop-in EncodeStream / DecodeStream run on: one Codec per GPU, pictures dealt round robin.
#ifndef VC2_CODEC_H
#define VC2_CODEC_H
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../vc2_cabi.h"

namespace vc2 {

class Codec {
 public:
  Codec(int device, const vc2_codec_params& p) : ctx_(vc2_create(device)), k_(nullptr) {
    if (!ctx_) throw std::runtime_error("vc2: cannot open CUDA device (the hot path has no CPU fallback)");
    k_ = vc2_codec_create(ctx_, &p);
    if (!k_) { const std::string m = vc2_last_error(ctx_); vc2_destroy(ctx_); throw std::invalid_argument("vc2 codec: " + m); }
  }
  ~Codec() { vc2_codec_destroy(k_); vc2_destroy(ctx_); }
  size_t pictureBytes() const { return vc2_codec_picture_in_bytes(k_); }
  size_t payloadCapacity() const { return vc2_codec_payload_capacity(k_); }
  // EncodeStream.cpp:456-565 for n pictures; throws the reference's std::logic_error texts
  void encode(int n, const void* const* pictures, uint8_t* const* payloads, size_t cap, size_t* lens) {
    check(vc2_codec_encode_host(k_, n, pictures, payloads, cap, lens));
  }
  // DecodeStream.cpp:512-605 for n pictures
  void decode(int n, const uint8_t* const* payloads, const size_t* lens, void* const* pictures) {
    check(vc2_codec_decode_host(k_, n, payloads, lens, pictures));
  }
  vc2_codec* handle() { return k_; }
  vc2_ctx* context() { return ctx_; }
  void check(int st) {
    if (st == VC2_OK) return;
    const std::string m = vc2_last_error(ctx_);
    if (st == VC2_ERR_ARG) throw std::invalid_argument(m);
    if (st == VC2_ERR_CUDA) throw std::runtime_error(m);
    throw std::logic_error(m);
  }
 private:
  Codec(const Codec&);
  Codec& operator=(const Codec&);
  vc2_ctx* ctx_;
  vc2_codec* k_;
};

}  // namespace vc2
#endif
