// C-ABI implementation (include/vc2_cabi.h): context, host-side helpers, Library-surface
// operations and the batched fused picture codec.  All compute goes through the CUDA kernels in
// dwt.cu / slices.cu; there is deliberately no CPU path for any of it.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <new>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "dwt.cuh"
#include <sched.h>
#include <cctype>
#include <cstdio>
#include "slices.cuh"

namespace vc2 {
cudaError_t dwt_fwd_launch(cudaStream_t s, int kernel, int sample_kind, const DwtParams& p, int npictures, int seg_rows);
cudaError_t dwt_inv_launch(cudaStream_t s, int kernel, int sample_kind, const DwtParams& p, int npictures, int seg_rows);
cudaError_t dwt_tile_fwd_launch(cudaStream_t s, int kernel, int sample_kind, const DwtParams& p, int npictures, int cfg);
cudaError_t dwt_tile_inv_launch(cudaStream_t s, int kernel, int sample_kind, const DwtParams& p, int npictures, int cfg);
}  // namespace vc2

using namespace vc2;

// ================================================================================================
// host-side helpers (pure host code)
// ================================================================================================

extern "C" int vc2_padded_size(int size, int depth) {
  const int cell = 1 << depth;   // WaveletTransform.cpp:74-77
  return cell * ((size + cell - 1) / cell);
}

extern "C" int vc2_slice_size_is_valid(int depth, int luma, int chroma, int n) {
  // WaveletTransform.cpp:116-136
  if (depth <= 0 || depth > 31) return 0;
  const int cell = 1 << depth;
  const int max_slices = std::min(luma, chroma) / cell;
  if (n <= 0 || n > max_slices) return 0;
  const int transform_size = n * cell;
  const int pl = vc2_padded_size(luma, depth), pc = vc2_padded_size(chroma, depth);
  const int ns = (pl + transform_size - 1) / transform_size;
  if (pl % ns == 0 && (pl / ns) % cell == 0 && pc % ns == 0 && (pc / ns) % cell == 0) return ns;
  return 0;
}

extern "C" int vc2_quant_matrix(int kernel, int depth, int32_t* out) {
  // WaveletTransform.cpp:345-423.  Same float expression order as the reference so that the
  // rounding of 4*log2(gain/minGain) lands on the same integers.
  if (depth < 0 || depth > VC2_MAX_DEPTH || !out) return VC2_ERR_ARG;
  if (depth == 0) { out[0] = 0; return VC2_OK; }
  float alpha, beta;
  int shift;
  switch (kernel) {
    case VC2_DD97: alpha = 1.280868846f; beta = 0.820572875f; shift = 1; break;
    case VC2_LEGALL: alpha = 1.224744871f; beta = 0.847791248f; shift = 1; break;
    case VC2_DD137: alpha = 1.280868846f; beta = 0.809253958f; shift = 1; break;
    case VC2_HAAR0: alpha = 1.414213562f; beta = 0.707106871f; shift = 0; break;
    case VC2_HAAR1: alpha = 1.414213562f; beta = 0.707106871f; shift = 1; break;
    case VC2_FIDELITY: alpha = 0.682408629f; beta = 1.367856979f; shift = 0; break;
    case VC2_DAUB97: alpha = 1.139917028f; beta = 0.887168005f; shift = 1; break;
    default: return VC2_ERR_ARG;
  }
  const float a2 = alpha * alpha, ab = alpha * beta, b2 = beta * beta;
  std::vector<float> ll(depth + 1), lh(depth + 1), hh(depth + 1);
  float min_gain = 3.402823466e+38f;
  for (int level = depth; level > 0; --level) {
    const float scale = pow(a2, depth - level) / pow(2.0f, shift * (depth - level + 1));
    ll[level] = scale * a2;
    lh[level] = scale * ab;
    hh[level] = scale * b2;
    min_gain = std::min(std::min(std::min(ll[level], lh[level]), hh[level]), min_gain);
  }
  std::vector<int> llq(depth + 1), lhq(depth + 1), hhq(depth + 1);
  for (int level = depth; level > 0; --level) {
    llq[level] = static_cast<int>(floor(4.0f * log(ll[level] / min_gain) / log(2.0f) + 0.5f));
    lhq[level] = static_cast<int>(floor(4.0f * log(lh[level] / min_gain) / log(2.0f) + 0.5f));
    hhq[level] = static_cast<int>(floor(4.0f * log(hh[level] / min_gain) / log(2.0f) + 0.5f));
  }
  int i = 0;
  out[i++] = llq[1];
  for (int level = 1; level <= depth; ++level) {
    out[i++] = lhq[level];
    out[i++] = lhq[level];
    out[i++] = hhq[level];
  }
  return VC2_OK;
}

extern "C" int vc2_quant_factor(int q) {
  // ST 2042-1 closed form; equals the 120-entry table of Quantisation.cpp:42-59 (checked in tests)
  if (q < 0) q = 0;
  const unsigned long long b = 1ull << (q / 4);
  switch (q & 3) {
    case 0: return (int)(unsigned)(4 * b);
    case 1: return (int)(unsigned)((503829ull * b + 52958ull) / 105917ull);
    case 2: return (int)(unsigned)((665857ull * b + 58854ull) / 117708ull);
    default: return (int)(unsigned)((440253ull * b + 32722ull) / 65444ull);
  }
}

extern "C" int vc2_quant_offset(int q) {
  // Quantisation.cpp:78-83
  if (q < 0) q = 0;
  if (q == 0) return 1;
  if (q == 1) return 2;
  return (vc2_quant_factor(q) + 1) / 2;
}

static int gcd_int(int a, int b) {
  if (a < 0) a = -a;
  if (b < 0) b = -b;
  while (b) { const int t = a % b; a = b; b = t; }
  return a;
}

extern "C" int vc2_slice_bytes(int ny, int nx, int total, int scalar, int32_t* out) {
  // Slices.cpp:28-49
  if (ny <= 0 || nx <= 0 || scalar <= 0 || !out) return VC2_ERR_ARG;
  int num = total / scalar - 4 * (ny * nx), den = ny * nx;
  const int g = gcd_int(num, den);
  if (g) { num /= g; den /= g; }
  const int ratio = num / den;
  const int remainder = num - ratio * den;
  int residue = 0;
  for (int i = 0; i < ny * nx; ++i) {
    residue += remainder;
    if (residue < den) out[i] = ratio * scalar + 4;
    else { out[i] = (ratio + 1) * scalar + 4; residue -= den; }
  }
  return VC2_OK;
}

extern "C" int vc2_make_geom(int height, int width, int chroma_format, int kernel, int depth, int v_slice, int h_slice,
                             int prefix, int scalar, vc2_geom* g) {
  if (!g || height < 1 || width < 1 || depth < 1 || depth > VC2_MAX_DEPTH || kernel < 0 || kernel > VC2_DAUB97 ||
      chroma_format < 0 || chroma_format > 2 || prefix < 0 || scalar < 1)
    return VC2_ERR_ARG;
  g->luma_h = height; g->luma_w = width;
  // PictureFormat::construct, Picture.cpp:49-73
  g->chroma_w = chroma_format == 0 ? width : width / 2;
  g->chroma_h = chroma_format == 2 ? height / 2 : height;
  g->kernel = kernel; g->depth = depth;
  g->slices_y = vc2_slice_size_is_valid(depth, g->luma_h, g->chroma_h, v_slice);
  g->slices_x = vc2_slice_size_is_valid(depth, g->luma_w, g->chroma_w, h_slice);
  g->prefix = prefix; g->scalar = scalar;
  if (g->slices_y == 0 || g->slices_x == 0) return VC2_ERR_ARG;
  return VC2_OK;
}

extern "C" int vc2_hq_index_slices(const uint8_t* p, size_t len, int n_slices, int prefix, int scalar, uint32_t* off) {
  // the read order of HQSliceIO_VBR(istream), Slices.cpp:544-605: lengths are the only framing
  if (!p || !off || n_slices < 0 || prefix < 0 || scalar < 1) return VC2_ERR_ARG;
  size_t pos = 0;
  for (int s = 0; s < n_slices; ++s) {
    off[s] = (uint32_t)pos;
    size_t q = pos + prefix + 1;   // after prefix and qindex
    for (int c = 0; c < 3; ++c) {
      if (q >= len) return VC2_ERR_STREAM;
      q += 1 + (size_t)p[q] * scalar;
    }
    if (q > len) return VC2_ERR_STREAM;
    pos = q;
  }
  off[n_slices] = (uint32_t)pos;
  return VC2_OK;
}

extern "C" const char* vc2_status_message(int st) {
  switch (st) {
    case VC2_OK: return "";
    case VC2_ERR_ARG: return "invalid argument";
    case VC2_ERR_CUDA: return "CUDA error (no usable GPU or runtime failure); this library has no CPU fallback";
    case VC2_ERR_SCALAR_TOO_SMALL: return "Slice scalar is too small, consider using a larger slice scalar.";
    case VC2_ERR_QUANT_INDEX: return "quantization index exceeds maximum implemented value.";
    case VC2_ERR_CBR_TOO_MANY_BYTES: return "SliceIO, HQ CBR mode: Too many bytes for the slice";
    case VC2_ERR_CBR_COMPONENT_LENGTH:
      return "Slice component length exceeds 1 byte when divided by slice size scalar. See above for suggestions to prevent this.";
    case VC2_ERR_CAPACITY: return "output buffer too small";
    case VC2_ERR_VLC_RANGE: return "quantised coefficient outside the 32-bit VLC range of the reference";
    case VC2_ERR_STREAM: return "malformed or truncated slice data";
    case VC2_ERR_LD_TOO_MANY_BYTES: return "SliceIO, LD mode: Too many bytes for the U and V slices";
    default: return "unknown error";
  }
}

// Run the calling thread (and the threads it starts later) on the CPUs next to a GPU, so that the host buffers it then
// allocates and fills are on the GPU's NUMA node: on a multi-socket box copies from the far node share the inter-socket
// link with every other rank.  Linux sysfs; VC2_ERR_ARG when the topology cannot be read (nothing is changed then).
extern "C" int vc2_bind_thread_to_device(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return VC2_ERR_CUDA; }
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE* f = fopen(path, "r");
  if (!f) return VC2_ERR_ARG;
  char list[4096] = {0};
  const bool got = fgets(list, (int)sizeof(list), f) != nullptr;
  fclose(f);
  if (!got) return VC2_ERR_ARG;
  cpu_set_t set;
  CPU_ZERO(&set);
  int n = 0;
  for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {   // "0-31,64-95"
    int a = 0, b = 0;
    const int k = sscanf(tok, "%d-%d", &a, &b);
    if (k < 1) continue;
    if (k == 1) b = a;
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &set); ++n; }
  }
  if (n == 0) return VC2_ERR_ARG;
  // keep to the CPUs this process is allowed on (containers, taskset)
  cpu_set_t allowed;
  if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
    cpu_set_t both;
    CPU_AND(&both, &set, &allowed);
    if (CPU_COUNT(&both) == 0) return VC2_ERR_ARG;
    set = both;
  }
  return sched_setaffinity(0, sizeof(set), &set) == 0 ? VC2_OK : VC2_ERR_ARG;
}

// page-locked host memory for the picture / payload buffers handed to vc2_codec_encode_host / _decode_host: copies from
// and to such buffers run at the full PCIe rate and overlap the kernels (pageable buffers are staged by the driver)
extern "C" void* vc2_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
extern "C" void vc2_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" int vc2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// ================================================================================================
// context
// ================================================================================================

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, n + 256);   // slack: the bit reader may touch 7 bytes past the data
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct vc2_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  long launches = 0;
  DevBuf tmp[8];   // scratch for the Library-surface (host pointer) calls
  int dwt_min_warps = 148 * 48;   // streaming DWT: cut rows into segments until the grid has at least this many warps
  int dwt_pd = 2;                 // prefetch distance in row pairs (VC2_DWT_PD)
  int dwt_fast = 1;               // fast loop of the lifting kernels (VC2_DWT_FAST=0 turns it off)
  int dwt_seg_rows = 0;           // > 0: forced rows per warp (tuning, VC2_DWT_SEG_ROWS)
  int dwt_tile = 1;               // lifting kernels: 1 = shared-memory tiles (dwt_tile.cu), 0 = streaming register rings (dwt.cu; VC2_DWT_TILE=0)
  int fuse_gather = 0;            // narrow packer: scan + gather inside the packing kernel (VC2_FUSE_GATHER=1).  Off: measured slower - the
                                  // images of all resident warps (3700 x 25 KB) do not stay in L2 until their warp gathers them
  int narrow = 1;                 // codecs keep quantised 16-bit coefficients between the lifting kernels and the slice coders (VC2_NARROW=0: 32-bit)
  DevBuf scale_tab;               // [128] (quant_factor, quant_offset + 2) for the inverse lifting kernels of the narrow path
  // optional per-kernel timing with CUDA events on the launch stream (vc2_profile_*)
  bool profiling = false;
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
};

static cudaEvent_t prof_event(vc2_ctx* c) {
  cudaEvent_t e = nullptr;
  if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}
struct ProfScope {
  vc2_ctx* c; int idx = -1;
  ProfScope(vc2_ctx* c_, int stage) : c(c_) {
    if (!c->profiling) return;
    vc2_ctx::Span sp; sp.stage = stage; sp.a = prof_event(c); sp.b = prof_event(c);
    cudaEventRecord(sp.a, c->stream);
    c->spans.push_back(sp); idx = (int)c->spans.size() - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(c->spans[idx].b, c->stream); }
};

static int fail(vc2_ctx* c, int st, const char* extra = nullptr) {
  if (c) {
    c->err = vc2_status_message(st);
    if (extra) { c->err += ": "; c->err += extra; }
  }
  return st;
}
static int cuda_fail(vc2_ctx* c, cudaError_t e) { return fail(c, VC2_ERR_CUDA, cudaGetErrorString(e)); }
#define CU(call)                                         \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return cuda_fail(ctx, e_);    \
  } while (0)

// Truncating division by quant_factor(q) for dividends below 2^31 with ONE multiply: a / d == mulhi(a, m) >> shift,
// m = ceil(2^(31 + l) / d), shift = l - 1, l = ceil(log2 d).  d >= 4 so l >= 2; 2^(l-1) < d <= 2^l so m < 2^32;
// m * d - 2^(31+l) < d <= 2^l, so the error term a * (m * d - 2^(31+l)) / (d * 2^(31+l)) stays below 1 / d for
// a < 2^31.  The dividend of the dead-zone quantiser is |v| << 2 computed in int (Quantisation.cpp:69-76): the
// reference itself is only defined for |v| < 2^29, i.e. for dividends below 2^31.
extern "C" int vc2_quant_magic31(int q, uint32_t* m, uint32_t* shift) {
  if (q < 0 || q > 119 || !m || !shift) return VC2_ERR_ARG;
  const uint64_t d = (uint64_t)(uint32_t)vc2_quant_factor(q);   // the top entries exceed INT_MAX (as in make_quant_tables)
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  *m = (uint32_t)(((1ull << (31 + l)) + d - 1) / d);
  *shift = l - 1;
  return VC2_OK;
}

struct NarrowMagic { uint32_t mul; int sh; };
static const NarrowMagic& narrow_magic(int q);
static void make_quant_tables(QuantTables& t) {
  for (int q = 0; q < 128; ++q) {
    const uint32_t d = (uint32_t)vc2_quant_factor(std::min(q, 119));
    t.qf[q] = d;
    t.qo[q] = (uint32_t)vc2_quant_offset(std::min(q, 119));
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;   // ceil(log2 d); d >= 4 so l >= 2
    t.ql[q] = l;
    t.qm[q] = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    vc2_quant_magic31(std::min(q, 119), &t.qm31[q], &t.ql31[q]);
    const NarrowMagic& m = narrow_magic(std::min(q, 119));
    t.qmul16[q] = m.sh < 0 ? 0u : m.mul;
    t.qsh16[q] = m.sh < 0 ? 0u : (uint32_t)m.sh;
  }
}

extern "C" vc2_ctx* vc2_create(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  vc2_ctx* c = new (std::nothrow) vc2_ctx();
  if (!c) return nullptr;
  c->device = device;
  if (const char* e = getenv("VC2_DWT_SEG_ROWS")) c->dwt_seg_rows = atoi(e) & ~1;
  if (const char* e = getenv("VC2_DWT_MIN_WARPS")) c->dwt_min_warps = atoi(e);
  if (const char* e = getenv("VC2_DWT_PD")) c->dwt_pd = atoi(e);
  if (const char* e = getenv("VC2_DWT_FAST")) c->dwt_fast = atoi(e) != 0;
  if (const char* e = getenv("VC2_DWT_TILE")) c->dwt_tile = atoi(e);
  if (const char* e = getenv("VC2_NARROW")) c->narrow = atoi(e) != 0;
  if (const char* e = getenv("VC2_FUSE_GATHER")) c->fuse_gather = atoi(e) != 0;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
  c->own_stream = true;
  QuantTables t;
  make_quant_tables(t);
  if (upload_quant_tables(t) != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return nullptr; }
  uint32_t st[256];
  for (int q = 0; q < 128; ++q) { st[2 * q] = t.qf[q]; st[2 * q + 1] = t.qo[q] + 2u; }
  if (c->scale_tab.reserve(sizeof(st)) != cudaSuccess || cudaMemcpy(c->scale_tab.p, st, sizeof(st), cudaMemcpyHostToDevice) != cudaSuccess) {
    c->scale_tab.release(); cudaStreamDestroy(c->stream); delete c; return nullptr;
  }
  return c;
}

extern "C" void vc2_destroy(vc2_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& b : c->tmp) b.release();
  c->scale_tab.release();
  for (auto& sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto e : c->event_pool) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int vc2_set_stream(vc2_ctx* c, void* s) {
  if (!c) return VC2_ERR_ARG;
  if (c->own_stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); c->own_stream = false; }
  c->stream = reinterpret_cast<cudaStream_t>(s);
  return VC2_OK;
}

extern "C" int vc2_synchronize(vc2_ctx* ctx) {
  if (!ctx) return VC2_ERR_ARG;
  CU(cudaStreamSynchronize(ctx->stream));
  return VC2_OK;
}

extern "C" const char* vc2_last_error(vc2_ctx* c) { return c ? c->err.c_str() : "no context"; }

extern "C" int vc2_kernel_launches(vc2_ctx* c, int reset) {
  if (!c) return 0;
  const long n = c->launches;
  if (reset) c->launches = 0;
  return (int)n;
}

extern "C" int vc2_profile_enable(vc2_ctx* ctx, int on) {
  if (!ctx) return VC2_ERR_ARG;
  ctx->profiling = on != 0;
  return VC2_OK;
}

// sum of CUDA-event durations (ms) and launch counts per stage since the last read; synchronises the stream
extern "C" int vc2_profile_read(vc2_ctx* ctx, float* ms, int* launches, int nstages) {
  if (!ctx || !ms || !launches) return VC2_ERR_ARG;
  CU(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < nstages; ++i) { ms[i] = 0.f; launches[i] = 0; }
  for (auto& sp : ctx->spans) {
    float t = 0.f;
    cudaEventElapsedTime(&t, sp.a, sp.b);
    if (sp.stage >= 0 && sp.stage < nstages) { ms[sp.stage] += t; launches[sp.stage]++; }
    ctx->event_pool.push_back(sp.a);
    ctx->event_pool.push_back(sp.b);
  }
  ctx->spans.clear();
  return VC2_OK;
}

// ================================================================================================
// pipeline helpers shared by the Library-surface calls and the fused codec
// ================================================================================================

struct CompBuf {
  PlaneGeom pg;
  void* pix = nullptr;              // dense picture plane (int32 or raw)
  long long pix_pic_stride = 0;     // elements (int32) or bytes (raw)
  int pix_pitch = 0;
  int comp = 0;                     // component index inside the slice-major block
  int32_t* coef = nullptr;          // slice-major coefficient block of picture 0
  long long coef_pic_stride = 0;
  int32_t* scratch[2] = {nullptr, nullptr};
  long long scratch_pic_stride[2] = {0, 0};
  int sshift = 0, soffset = 0, clip_min = 0, clip_max = 0;
};

static PlaneGeom make_plane(int h, int w, int depth) {
  PlaneGeom g;
  g.h = h; g.w = w; g.depth = depth;
  g.ph = vc2_padded_size(h, depth);
  g.pw = vc2_padded_size(w, depth);
  return g;
}

// all levels of the forward or inverse transform for ncomp components of npictures pictures
static int ilog2_or_neg(int v) {
  if (v <= 0 || (v & (v - 1))) return -1;
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// The narrow path's forward quantiser: |q| = (|v| * mul) >> sh == (|v| << 2) / quant_factor(q) for every |v| below
// VC2_NARROW_FAST_MAX, with a product that stays inside 32 bits.  Found by trying the shifts in turn and checking every
// dividend (once per index and process); an index without such a pair keeps mul = 0, sh = -1 and the kernels never take
// the fast form for it (cannot happen for the table of Quantisation.cpp:42-59, but nothing relies on that).
static const NarrowMagic& narrow_magic(int q) {
  static NarrowMagic tab[120];
  static std::once_flag once;
  std::call_once(once, [] {
    for (int i = 0; i < 120; ++i) {
      const uint64_t d = (uint64_t)(uint32_t)vc2_quant_factor(i);
      tab[i].mul = 0; tab[i].sh = -1;
      for (int sh = 0; sh < 32 && tab[i].sh < 0; ++sh) {
        const uint64_t mul = ((4ull << sh) + d - 1) / d;
        if (mul * (VC2_NARROW_FAST_MAX - 1) >= (1ull << 32)) break;
        bool ok = true;
        for (uint64_t v = 0; v < VC2_NARROW_FAST_MAX && ok; ++v) ok = ((v * mul) >> sh) == (4 * v) / d;
        if (ok) { tab[i].mul = (uint32_t)mul; tab[i].sh = sh; }
      }
    }
  });
  return tab[std::min(std::max(q, 0), 119)];
}

// narrow coefficient block (dwt.cuh): off, forward with one quantisation index (HQ_ConstQ), or inverse with the index of every slice
struct NarrowCfg {
  int on = 0;
  int qindex = 0;                 // forward
  const int32_t* qidx = nullptr;  // inverse: [picture][slice]
  uint32_t* ovf = nullptr;        // [picture]
  const BandScale* band_scale = nullptr; // inverse: [picture], see UnpackParams::band_scale
};

static cudaError_t run_dwt(vc2_ctx* ctx, bool inverse, int kernel, int depth, int sample_kind, const SliceGeom& g, const CompBuf* cb,
                           int ncomp, int npictures, const NarrowCfg& nw = NarrowCfg()) {
  if (nw.on && !ctx->dwt_tile) return cudaErrorInvalidValue;   // only the tile kernels know the narrow block
  for (int step = 0; step < depth; ++step) {
    const int l = inverse ? depth - 1 - step : step;   // 0 = finest
    const int L = depth - l;                           // VC-2 level of the bands touched
    DwtParams p;
    memset(&p, 0, sizeof(p));
    p.ncomp = ncomp;
    p.pd = ctx->dwt_pd;
    p.fast = ctx->dwt_fast;
    p.narrow = nw.on;
    p.narrow_ovf = nw.ovf;
    p.qidx = nw.qidx;
    p.band_scale = nw.band_scale;
    p.nslices = g.slices_x * g.slices_y;
    p.scale_tab = ctx->scale_tab.as<uint2>();
    for (int c = 0; c < ncomp; ++c) {
      const CompBuf& B = cb[c];
      DwtComp& C = p.c[c];
      C.lat_h = B.pg.ph >> l;
      C.lat_w = B.pg.pw >> l;
      if (l == 0) {
        C.pix = B.pix; C.pix_pic_stride = B.pix_pic_stride; C.pix_h = B.pg.h; C.pix_w = B.pg.w; C.pix_pitch = B.pix_pitch;
      } else {
        C.pix = B.scratch[(l - 1) & 1]; C.pix_pic_stride = B.scratch_pic_stride[(l - 1) & 1];
        C.pix_h = C.lat_h; C.pix_w = C.lat_w; C.pix_pitch = C.lat_w;
      }
      if (l == depth - 1) {
        C.ll = nullptr; C.ll_pic_stride = 0;   // LL of the last level is band 0 of the coefficient block
      } else {
        C.ll = B.scratch[l & 1]; C.ll_pic_stride = B.scratch_pic_stride[l & 1];
      }
      C.ll_pitch = C.lat_w / 2;
      const int cc = B.comp, bhl = 3 * (L - 1) + 1;
      C.coef = B.coef; C.coef_pic_stride = B.coef_pic_stride;
      C.bh = g.part_h[cc][bhl]; C.bw = g.part_w[cc][bhl];
      C.lgbh = ilog2_or_neg(C.bh); C.lgbw = ilog2_or_neg(C.bw);
      C.nx = g.slices_x; C.NC = g.comp_start[3];
      C.base_ll = g.comp_start[cc] + g.band_start[cc][0];
      C.base_hl = g.comp_start[cc] + g.band_start[cc][bhl];
      C.base_lh = g.comp_start[cc] + g.band_start[cc][bhl + 1];
      C.base_hh = g.comp_start[cc] + g.band_start[cc][bhl + 2];
      C.sshift = B.sshift; C.soffset = B.soffset; C.clip_min = B.clip_min; C.clip_max = B.clip_max;
      if (nw.on) {
        const int band[4] = {0, bhl, bhl + 1, bhl + 2};
        for (int i = 0; i < 4; ++i) {
          C.qmat[i] = g.qmatrix[band[i]];
          C.band[i] = band[i];
          const int q = std::min(std::max(nw.qindex - g.qmatrix[band[i]], 0), 119);   // Quantisation.cpp:16-20; beyond 119 the packer raises the error
          const NarrowMagic& m = narrow_magic(q);
          C.qmul[i] = m.mul; C.qsh[i] = m.sh < 0 ? 0 : m.sh;
          uint32_t mm = 0, ll = 0;
          vc2_quant_magic31(q, &mm, &ll);
          C.qm31[i] = mm; C.ql31[i] = (int)ll;
          if (m.sh < 0) { C.qmul[i] = 0; C.qsh[i] = 0; }
        }
      }
    }
    ProfScope ps(ctx, inverse ? (l == 0 ? VC2_STAGE_IDWT_L0 : VC2_STAGE_IDWT_DEEP) : (l == 0 ? VC2_STAGE_DWT_L0 : VC2_STAGE_DWT_DEEP));
    // rows per warp: long segments amortise the warm-up rows of the streaming kernels, short ones fill the GPU
    long long cols = 0;
    int maxh = 0;
    for (int c = 0; c < ncomp; ++c) { cols += p.c[c].lat_w; maxh = std::max(maxh, p.c[c].lat_h); }
    int seg_rows = 256;
    while (seg_rows > 32 && (cols / 240 + ncomp) * ((maxh + seg_rows - 1) / seg_rows) * npictures < ctx->dwt_min_warps) seg_rows >>= 1;
    if (ctx->dwt_seg_rows > 0) seg_rows = ctx->dwt_seg_rows;
    const int kind = l == 0 ? sample_kind : SAMPLE_I32;
    cudaError_t e;
    if (ctx->dwt_tile) e = inverse ? dwt_tile_inv_launch(ctx->stream, kernel, kind, p, npictures, ctx->dwt_tile)
                                   : dwt_tile_fwd_launch(ctx->stream, kernel, kind, p, npictures, ctx->dwt_tile);
    else e = inverse ? dwt_inv_launch(ctx->stream, kernel, kind, p, npictures, seg_rows)
                     : dwt_fwd_launch(ctx->stream, kernel, kind, p, npictures, seg_rows);
    if (e != cudaSuccess) return e;
    ctx->launches++;
  }
  return cudaSuccess;
}

static void fill_slice_geom(SliceGeom& g, const vc2_geom& vg, const int32_t* qmatrix) {
  memset(&g, 0, sizeof(g));
  g.depth = vg.depth;
  g.nbands = 3 * vg.depth + 1;
  g.slices_x = vg.slices_x; g.slices_y = vg.slices_y;
  g.prefix = vg.prefix; g.scalar = vg.scalar;
  g.plane[0] = make_plane(vg.luma_h, vg.luma_w, vg.depth);
  g.plane[1] = g.plane[2] = make_plane(vg.chroma_h, vg.chroma_w, vg.depth);
  int cs = 0;
  for (int c = 0; c < 3; ++c) {
    g.comp_start[c] = cs;
    int bs = 0;
    for (int b = 0; b < g.nbands; ++b) {
      const int lev = b == 0 ? g.depth : g.depth - ((b - 1) / 3 + 1) + 1;   // band dims = padded dims >> lev
      g.part_h[c][b] = (g.plane[c].ph >> lev) / g.slices_y;
      g.part_w[c][b] = (g.plane[c].pw >> lev) / g.slices_x;
      g.band_start[c][b] = bs;
      bs += g.part_h[c][b] * g.part_w[c][b];
    }
    g.band_start[c][g.nbands] = bs;
    cs += bs;
  }
  g.comp_start[3] = cs;
  g.coef_pic_stride = (long long)((g.slices_x * g.slices_y + 31) / 32) * 32 * cs;   // whole groups of 32 slices
  for (int b = 0; b < g.nbands; ++b) g.qmatrix[b] = qmatrix ? qmatrix[b] : 0;
}

// one plane treated as a single component cut into the smallest slices (one 2^depth x 2^depth cell each),
// so that the groups of the interleaved layout are full (stand-alone transforms)
static void fill_plane_geom(SliceGeom& g, int h, int w, int ph, int pw, int depth) {
  memset(&g, 0, sizeof(g));
  g.depth = depth;
  g.nbands = 3 * depth + 1;
  g.slices_x = pw >> depth; g.slices_y = ph >> depth;
  g.scalar = 1;
  g.plane[0].h = h; g.plane[0].w = w; g.plane[0].ph = ph; g.plane[0].pw = pw; g.plane[0].depth = depth;
  g.plane[1] = g.plane[2] = g.plane[0];
  g.plane[1].h = g.plane[1].w = g.plane[1].ph = g.plane[1].pw = 0;
  g.plane[2] = g.plane[1];
  int bs = 0;
  for (int b = 0; b < g.nbands; ++b) {
    const int side = b == 0 ? 1 : 1 << ((b - 1) / 3);   // level L band part = 2^(L-1) square
    g.part_h[0][b] = side;
    g.part_w[0][b] = side;
    g.band_start[0][b] = bs;
    bs += side * side;
  }
  g.band_start[0][g.nbands] = bs;
  g.comp_start[0] = 0;
  g.comp_start[1] = g.comp_start[2] = g.comp_start[3] = bs;
  g.coef_pic_stride = (long long)((g.slices_x * g.slices_y + 31) / 32) * 32 * bs;
}

static bool geom_ok(const vc2_geom* g) {
  if (!g) return false;
  if (g->depth < 1 || g->depth > VC2_MAX_DEPTH || g->kernel < 0 || g->kernel > VC2_DAUB97) return false;
  if (g->luma_h < 1 || g->luma_w < 1 || g->chroma_h < 1 || g->chroma_w < 1) return false;
  if (g->slices_x < 1 || g->slices_y < 1 || g->prefix < 0 || g->scalar < 1) return false;
  const int cell = 1 << g->depth;
  const int ph = vc2_padded_size(g->luma_h, g->depth), pw = vc2_padded_size(g->luma_w, g->depth);
  const int ch = vc2_padded_size(g->chroma_h, g->depth), cw = vc2_padded_size(g->chroma_w, g->depth);
  if (ph % g->slices_y || pw % g->slices_x || ch % g->slices_y || cw % g->slices_x) return false;
  if ((ph / g->slices_y) % cell || (pw / g->slices_x) % cell || (ch / g->slices_y) % cell || (cw / g->slices_x) % cell) return false;
  return true;
}

// worst-case payload of one slice: every coefficient at the 32-bit VLC limit, or the length-byte limit
static int max_slice_bytes(const SliceGeom& g, int mode, const int32_t* slice_bytes) {
  const int nslices = g.slices_x * g.slices_y;
  int m = 0;
  if (mode == VC2_HQ_CBR && slice_bytes) {
    for (int i = 0; i < nslices; ++i) m = std::max(m, slice_bytes[i]);
    m += g.prefix;
  }
  int v = g.prefix + 4;
  for (int c = 0; c < 3; ++c) {
    const int n = g.band_start[c][g.nbands];
    const int by_bits = ((4 * n + g.scalar - 1) / g.scalar) * g.scalar;   // 32 bits per coefficient
    v += std::min(by_bits, 255 * g.scalar);
  }
  return std::max(m, v);
}

static int flags_to_status(unsigned f) {
  if (f & VC2_FLAG_QUANT_INDEX) return VC2_ERR_QUANT_INDEX;
  if (f & VC2_FLAG_SCALAR_TOO_SMALL) return VC2_ERR_SCALAR_TOO_SMALL;
  if (f & VC2_FLAG_CBR_TOO_MANY_BYTES) return VC2_ERR_CBR_TOO_MANY_BYTES;
  if (f & VC2_FLAG_CBR_COMP_LENGTH) return VC2_ERR_CBR_COMPONENT_LENGTH;
  if (f & VC2_FLAG_LD_TOO_MANY_BYTES) return VC2_ERR_LD_TOO_MANY_BYTES;
  if (f & VC2_FLAG_VLC_RANGE) return VC2_ERR_VLC_RANGE;
  if (f & VC2_FLAG_STREAM) return VC2_ERR_STREAM;
  return VC2_OK;
}
// reference throw order: rate control for every slice (raster) runs before any slice is written
// `ignore`: flag bits the caller does not treat as errors (the decoders pass VC2_FLAG_VLC_RANGE: a parsed value beyond the
// encoder's domain is still decoded, but a malformed slice behind it must not hide behind that)
static int first_error(const uint32_t* flags, int n, uint32_t ignore = 0) {
  for (int i = 0; i < n; ++i)
    if (flags[i] & VC2_FLAG_SEARCH_PHASE & ~ignore) return flags_to_status(flags[i] & ~ignore);
  for (int i = 0; i < n; ++i)
    if (flags[i] & ~ignore) return flags_to_status(flags[i] & ~ignore);
  return VC2_OK;
}

struct PackBuffers {
  uint8_t* out; long long out_stride; long long out_capacity;
  uint32_t* slice_off; uint32_t* err_flags; int32_t* qidx;
  uint32_t* staging; uint32_t* sizes;
  const int32_t* slice_bytes_dev; const uint32_t* fixed_off_dev;
  uint32_t* total_len = nullptr;   // [pic] payload bytes, device memory (optional)
  // fused scan + gather of the narrow packer: [pic][tiles] 64-bit tile states followed by [pic] 32-bit tickets (NULL: separate scan and gather)
  unsigned long long* tile_state = nullptr;
};
static size_t tile_state_bytes(int nslices, int npictures) { return ((size_t)((nslices + 127) / 128) * 8 + 8) * npictures; }

// staging words per slice: every coefficient at the 32-bit VLC limit, plus header bytes and slack
// worst case (every code 32 bits) plus slack for the assembler's look-ahead, in whole 16-byte units: the HQ packer
// stores four words at a time (WideBitWriter)
static int staging_words(const SliceGeom& g) { return ((g.prefix + 4 + 4 * g.comp_start[3] + 3) / 4 + 4 + 3) & ~3; }

static cudaError_t run_pack(vc2_ctx* ctx, const SliceGeom& g, const int32_t* coef, int npictures, int mode, int quantise,
                            int search, int const_q, int emit, const PackBuffers& B, int narrow = 0) {
  PackParams p;
  memset(&p, 0, sizeof(p));
  p.g = g;
  p.coef = coef;
  p.narrow = narrow;
  p.mode = mode; p.quantise = quantise; p.search = search; p.const_q = const_q; p.emit = emit;
  p.qidx = B.qidx; p.slice_bytes = B.slice_bytes_dev;
  p.staging = B.staging; p.wcap = staging_words(g); p.sizes = B.sizes; p.err_flags = B.err_flags;
  cudaError_t e;
  const bool fuse = narrow && B.tile_state && ctx->fuse_gather;
  if (fuse) {
    const int nslices = g.slices_x * g.slices_y;
    p.fuse = 1;
    p.tiles = (nslices + 127) / 128;
    p.tile_state = B.tile_state;
    p.tile_ticket = reinterpret_cast<uint32_t*>(B.tile_state + (size_t)p.tiles * npictures);
    p.out = B.out; p.out_pic_stride = B.out_stride; p.out_capacity = B.out_capacity;
    p.slice_off = B.slice_off; p.total_len = B.total_len;
    e = cudaMemsetAsync(B.tile_state, 0, tile_state_bytes(nslices, npictures), ctx->stream);
    if (e != cudaSuccess) return e;
  }
  if (search && emit) {
    // HQ_CBR: rate control and packing as two launches - the search leaves the index of every slice (and its error
    // flags), the packer reads them - so that the two are timed apart (VC2_STAGE_SEARCH)
    PackParams q = p;
    q.emit = 0;
    {
      ProfScope ps(ctx, VC2_STAGE_SEARCH);
      e = pack_launch(ctx->stream, q, npictures);
    }
    if (e != cudaSuccess) return e;
    ctx->launches++;
    p.search = 0; p.const_q = -1; p.after_search = 1;
  }
  {
    ProfScope ps(ctx, VC2_STAGE_PACK);
    e = pack_launch(ctx->stream, p, npictures);
  }
  if (e != cudaSuccess) return e;
  ctx->launches++;
  if (!emit || fuse) return cudaSuccess;
  AssembleParams a;
  memset(&a, 0, sizeof(a));
  a.nslices = g.slices_x * g.slices_y;
  a.sizes = B.sizes; a.fixed_off = B.fixed_off_dev; a.slice_off = B.slice_off; a.total_len = B.total_len;
  a.staging = B.staging; a.wcap = p.wcap;
  a.out = B.out; a.out_pic_stride = B.out_stride; a.out_capacity = B.out_capacity; a.err_flags = B.err_flags;
  {
    ProfScope ps(ctx, VC2_STAGE_ASSEMBLE);
    e = assemble_launch(ctx->stream, a, npictures);
  }
  if (e == cudaSuccess) ctx->launches += 2;
  return e;
}

// ================================================================================================
// Library-surface operations (host buffers)
// ================================================================================================

extern "C" int vc2_dwt_forward(vc2_ctx* ctx, const int32_t* src, int h, int w, int kernel, int depth, int32_t* dst) {
  if (!ctx || !src || !dst || h < 1 || w < 1 || depth < 1 || depth > VC2_MAX_DEPTH || kernel < 0 || kernel > VC2_DAUB97)
    return fail(ctx, VC2_ERR_ARG);
  CU(cudaSetDevice(ctx->device));
  CompBuf B;
  B.pg = make_plane(h, w, depth);
  const size_t n_in = (size_t)h * w * 4, n_pad = (size_t)B.pg.size() * 4;
  CU(ctx->tmp[0].reserve(n_in));
  SliceGeom g;
  fill_plane_geom(g, h, w, B.pg.ph, B.pg.pw, depth);
  CU(ctx->tmp[1].reserve((size_t)g.coef_pic_stride * 4));
  CU(ctx->tmp[2].reserve(n_pad / 4));
  CU(ctx->tmp[3].reserve(n_pad / 16 + 4));
  CU(ctx->tmp[4].reserve(n_pad));
  B.pix = ctx->tmp[0].p; B.pix_pitch = w; B.pix_pic_stride = (long long)h * w;
  B.coef = ctx->tmp[1].as<int32_t>(); B.coef_pic_stride = g.coef_pic_stride; B.comp = 0;
  B.scratch[0] = ctx->tmp[2].as<int32_t>(); B.scratch[1] = ctx->tmp[3].as<int32_t>();
  CU(cudaMemcpyAsync(B.pix, src, n_in, cudaMemcpyHostToDevice, ctx->stream));
  CU(run_dwt(ctx, false, kernel, depth, SAMPLE_I32, g, &B, 1, 1));
  CU(layout_launch(ctx->stream, false, B.coef, ctx->tmp[4].as<int32_t>(), g, 0));
  ctx->launches++;
  CU(cudaMemcpyAsync(dst, ctx->tmp[4].p, n_pad, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return VC2_OK;
}

extern "C" int vc2_dwt_inverse(vc2_ctx* ctx, const int32_t* src, int ph, int pw, int kernel, int depth, int32_t* dst, int h,
                               int w) {
  if (!ctx || !src || !dst || h < 1 || w < 1 || depth < 1 || depth > VC2_MAX_DEPTH || kernel < 0 || kernel > VC2_DAUB97)
    return fail(ctx, VC2_ERR_ARG);
  if (ph % (1 << depth) || pw % (1 << depth) || h > ph || w > pw) return fail(ctx, VC2_ERR_ARG);
  CU(cudaSetDevice(ctx->device));
  CompBuf B;
  B.pg.h = h; B.pg.w = w; B.pg.ph = ph; B.pg.pw = pw; B.pg.depth = depth;
  const size_t n_out = (size_t)h * w * 4, n_pad = (size_t)ph * pw * 4;
  CU(ctx->tmp[0].reserve(n_out));
  SliceGeom g;
  fill_plane_geom(g, h, w, ph, pw, depth);
  CU(ctx->tmp[1].reserve((size_t)g.coef_pic_stride * 4));
  CU(ctx->tmp[2].reserve(n_pad / 4));
  CU(ctx->tmp[3].reserve(n_pad / 16 + 4));
  CU(ctx->tmp[4].reserve(n_pad));
  B.pix = ctx->tmp[0].p; B.pix_pitch = w; B.pix_pic_stride = (long long)h * w;
  B.coef = ctx->tmp[1].as<int32_t>(); B.coef_pic_stride = g.coef_pic_stride; B.comp = 0;
  B.scratch[0] = ctx->tmp[2].as<int32_t>(); B.scratch[1] = ctx->tmp[3].as<int32_t>();
  CU(cudaMemcpyAsync(ctx->tmp[4].p, src, n_pad, cudaMemcpyHostToDevice, ctx->stream));
  CU(layout_launch(ctx->stream, true, ctx->tmp[4].as<int32_t>(), B.coef, g, 0));
  ctx->launches++;
  CU(run_dwt(ctx, true, kernel, depth, SAMPLE_I32, g, &B, 1, 1));
  CU(cudaMemcpyAsync(dst, B.pix, n_out, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return VC2_OK;
}

static int quant_host(vc2_ctx* ctx, const int32_t* coef, int ph, int pw, int depth, const int32_t* qmatrix, const int32_t* qidx,
                      int ny, int nx, int32_t* out, int inverse, int ld) {
  if (!ctx || !coef || !qmatrix || !qidx || !out || depth < 1 || depth > VC2_MAX_DEPTH || ny < 1 || nx < 1 || ph < 1 || pw < 1)
    return fail(ctx, VC2_ERR_ARG);
  if (ph % ny || pw % nx || ph % (1 << depth) || pw % (1 << depth)) return fail(ctx, VC2_ERR_ARG);
  // quant_factor throws for adjusted indices above 119 (Quantisation.cpp:60-63)
  int minm = qmatrix[0];
  for (int b = 0; b < 3 * depth + 1; ++b) minm = std::min(minm, qmatrix[b]);
  for (int i = 0; i < ny * nx; ++i)
    if (qidx[i] - minm > 119) return fail(ctx, VC2_ERR_QUANT_INDEX);
  CU(cudaSetDevice(ctx->device));
  const size_t n = (size_t)ph * pw * 4;
  CU(ctx->tmp[0].reserve(n));
  CU(ctx->tmp[1].reserve(n));
  CU(ctx->tmp[2].reserve((size_t)ny * nx * 4));
  CU(cudaMemcpyAsync(ctx->tmp[0].p, coef, n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->tmp[2].p, qidx, (size_t)ny * nx * 4, cudaMemcpyHostToDevice, ctx->stream));
  QuantParams p;
  memset(&p, 0, sizeof(p));
  p.src = ctx->tmp[0].as<int32_t>(); p.dst = ctx->tmp[1].as<int32_t>(); p.qidx = ctx->tmp[2].as<int32_t>();
  p.ph = ph; p.pw = pw; p.depth = depth; p.slices_y = ny; p.slices_x = nx; p.inverse = inverse; p.skip_ll = ld;
  for (int b = 0; b < 3 * depth + 1; ++b) p.qmatrix[b] = qmatrix[b];
  CU(quant_launch(ctx->stream, p));
  ctx->launches++;
  if (ld && inverse) {
    LdDcParams d;
    memset(&d, 0, sizeof(d));
    d.base = p.dst; d.qidx = p.qidx; d.H = ph >> depth; d.W = pw >> depth;
    d.interleaved = 0; d.pitch = pw; d.depth = depth;
    d.slices_y = ny; d.slices_x = nx; d.qm0 = qmatrix[0];
    CU(ld_dc_launch(ctx->stream, d));
    ctx->launches++;
  } else if (ld) {
    LdDcQuantParams d;
    CU(ctx->tmp[3].reserve((size_t)(ph >> depth) * (pw >> depth) * 4));
    d.src = p.src; d.dst = p.dst; d.restored = ctx->tmp[3].as<int32_t>(); d.qidx = p.qidx;
    d.H = ph >> depth; d.W = pw >> depth; d.pitch = pw; d.depth = depth;
    d.slices_y = ny; d.slices_x = nx; d.qm0 = qmatrix[0];
    CU(ld_dc_quant_launch(ctx->stream, d));
    ctx->launches++;
  }
  CU(cudaMemcpyAsync(out, ctx->tmp[1].p, n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return VC2_OK;
}

extern "C" int vc2_quantise_np(vc2_ctx* ctx, const int32_t* coef, int ph, int pw, int depth, const int32_t* qmatrix,
                               const int32_t* qidx, int ny, int nx, int32_t* out) {
  return quant_host(ctx, coef, ph, pw, depth, qmatrix, qidx, ny, nx, out, 0, 0);
}
extern "C" int vc2_dequantise_np(vc2_ctx* ctx, const int32_t* coef, int ph, int pw, int depth, const int32_t* qmatrix,
                                 const int32_t* qidx, int ny, int nx, int32_t* out) {
  return quant_host(ctx, coef, ph, pw, depth, qmatrix, qidx, ny, nx, out, 1, 0);
}
extern "C" int vc2_dequantise_ld(vc2_ctx* ctx, const int32_t* coef, int ph, int pw, int depth, const int32_t* qmatrix,
                                 const int32_t* qidx, int ny, int nx, int32_t* out) {
  return quant_host(ctx, coef, ph, pw, depth, qmatrix, qidx, ny, nx, out, 1, 1);
}

// upload three in-place planes and convert them to one slice-major coefficient block in tmp[1]
static int upload_planes(vc2_ctx* ctx, const SliceGeom& g, const int32_t* y, const int32_t* u, const int32_t* v) {
  const int32_t* src[3] = {y, u, v};
  CU(ctx->tmp[0].reserve((size_t)g.plane[0].size() * 4));
  CU(ctx->tmp[1].reserve((size_t)g.coef_pic_stride * 4));
  for (int c = 0; c < 3; ++c) {
    const size_t n = (size_t)g.plane[c].size() * 4;
    CU(cudaMemcpyAsync(ctx->tmp[0].p, src[c], n, cudaMemcpyHostToDevice, ctx->stream));
    CU(layout_launch(ctx->stream, true, ctx->tmp[0].as<int32_t>(), ctx->tmp[1].as<int32_t>(), g, c));
    ctx->launches++;
  }
  return VC2_OK;
}
static int download_planes(vc2_ctx* ctx, const SliceGeom& g, const int32_t* coef, int32_t* y, int32_t* u, int32_t* v) {
  int32_t* dst[3] = {y, u, v};
  CU(ctx->tmp[0].reserve((size_t)g.plane[0].size() * 4));
  for (int c = 0; c < 3; ++c) {
    if (!dst[c]) continue;
    const size_t n = (size_t)g.plane[c].size() * 4;
    CU(layout_launch(ctx->stream, false, coef, ctx->tmp[0].as<int32_t>(), g, c));
    ctx->launches++;
    CU(cudaMemcpyAsync(dst[c], ctx->tmp[0].p, n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  return VC2_OK;
}

// shared body of vc2_hq_pack / vc2_cbr_qindices
static int pack_host(vc2_ctx* ctx, const int32_t* Y, const int32_t* U, const int32_t* V, const vc2_geom* vg, const int32_t* qmatrix,
                     const int32_t* qidx_in, int32_t* qidx_out, int mode, int search, int emit, const int32_t* slice_bytes,
                     uint8_t* out, size_t cap, size_t* out_len, uint32_t* slice_off, uint32_t* err_flags_out) {
  if (!ctx || !Y || !U || !V || !geom_ok(vg)) return fail(ctx, VC2_ERR_ARG);
  if ((mode == VC2_HQ_CBR || search) && !slice_bytes) return fail(ctx, VC2_ERR_ARG);
  CU(cudaSetDevice(ctx->device));
  SliceGeom g;
  fill_slice_geom(g, *vg, qmatrix);
  const int nslices = g.slices_x * g.slices_y;
  int st = upload_planes(ctx, g, Y, U, V);
  if (st) return st;
  const int img_bytes = max_slice_bytes(g, mode, slice_bytes);
  std::vector<uint32_t> fixed(nslices + 1, 0);
  size_t total_cap = 0;
  if (mode == VC2_HQ_CBR) {
    for (int i = 0; i < nslices; ++i) fixed[i + 1] = fixed[i] + (uint32_t)(slice_bytes[i] + g.prefix);
    total_cap = fixed[nslices];
  } else {
    total_cap = (size_t)img_bytes * nslices;
  }
  // layout of tmp[2]: payload | slice_off | err | sizes | qidx | slice_bytes | fixed ; tmp[3]: staging
  const size_t o_pay = 0, o_off = (total_cap + 259) / 256 * 256, o_err = o_off + (size_t)(nslices + 1) * 4,
               o_sz = o_err + (size_t)nslices * 4, o_q = o_sz + (size_t)nslices * 4, o_sb = o_q + (size_t)nslices * 4,
               o_fx = o_sb + (size_t)nslices * 4, o_end = o_fx + (size_t)(nslices + 1) * 4;
  CU(ctx->tmp[2].reserve(o_end));
  CU(ctx->tmp[3].reserve((size_t)staging_words(g) * 4 * nslices));
  uint8_t* base = ctx->tmp[2].as<uint8_t>();
  PackBuffers B;
  B.out = base + o_pay; B.out_stride = 0; B.out_capacity = (long long)total_cap;
  B.slice_off = (uint32_t*)(base + o_off); B.err_flags = (uint32_t*)(base + o_err); B.qidx = (int32_t*)(base + o_q);
  B.sizes = (uint32_t*)(base + o_sz); B.staging = ctx->tmp[3].as<uint32_t>();
  B.slice_bytes_dev = slice_bytes ? (const int32_t*)(base + o_sb) : nullptr;
  B.fixed_off_dev = mode == VC2_HQ_CBR ? (const uint32_t*)(base + o_fx) : nullptr;
  CU(cudaMemsetAsync(base + o_off, 0, o_q - o_off, ctx->stream));
  if (qidx_in) CU(cudaMemcpyAsync(B.qidx, qidx_in, (size_t)nslices * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (slice_bytes) CU(cudaMemcpyAsync(base + o_sb, slice_bytes, (size_t)nslices * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (mode == VC2_HQ_CBR) CU(cudaMemcpyAsync(base + o_fx, fixed.data(), (size_t)(nslices + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(run_pack(ctx, g, ctx->tmp[1].as<int32_t>(), 1, mode, search ? 1 : 0, search, -1, emit, B));
  std::vector<uint32_t> flags(nslices), offs(nslices + 1);
  CU(cudaMemcpyAsync(flags.data(), B.err_flags, (size_t)nslices * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(offs.data(), B.slice_off, (size_t)(nslices + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (qidx_out) CU(cudaMemcpyAsync(qidx_out, B.qidx, (size_t)nslices * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (err_flags_out) memcpy(err_flags_out, flags.data(), (size_t)nslices * 4);
  st = first_error(flags.data(), nslices);
  if (st) return fail(ctx, st);
  if (emit) {
    const size_t len = offs[nslices];
    if (out_len) *out_len = len;
    if (len > cap) return fail(ctx, VC2_ERR_CAPACITY);
    CU(cudaMemcpy(out, B.out, len, cudaMemcpyDeviceToHost));
    if (slice_off) memcpy(slice_off, offs.data(), (size_t)(nslices + 1) * 4);
  }
  return VC2_OK;
}

extern "C" int vc2_hq_pack(vc2_ctx* ctx, const int32_t* qY, const int32_t* qU, const int32_t* qV, const vc2_geom* g,
                           const int32_t* qidx, int mode, const int32_t* slice_bytes, uint8_t* out, size_t cap, size_t* out_len,
                           uint32_t* slice_off) {
  if (!qidx || !out || (mode != VC2_HQ_VBR && mode != VC2_HQ_CBR)) return fail(ctx, VC2_ERR_ARG);
  return pack_host(ctx, qY, qU, qV, g, nullptr, qidx, nullptr, mode, 0, 1, slice_bytes, out, cap, out_len, slice_off, nullptr);
}

extern "C" int vc2_cbr_qindices(vc2_ctx* ctx, const int32_t* cY, const int32_t* cU, const int32_t* cV, const vc2_geom* g,
                                const int32_t* qmatrix, const int32_t* slice_bytes, int32_t* qidx_out, uint32_t* err_flags) {
  if (!qmatrix || !qidx_out) return fail(ctx, VC2_ERR_ARG);
  return pack_host(ctx, cY, cU, cV, g, qmatrix, nullptr, qidx_out, VC2_HQ_CBR, 1, 0, slice_bytes, nullptr, 0, nullptr, nullptr, err_flags);
}

// operator<<(ostream&, Slices) with the LD writer (Slices.cpp:195-244, 645-660) on already quantised planes
extern "C" int vc2_ld_pack(vc2_ctx* ctx, const int32_t* qY, const int32_t* qU, const int32_t* qV, const vc2_geom* vg,
                           const int32_t* qidx, const int32_t* slice_bytes, uint8_t* out, size_t cap, size_t* out_len) {
  if (!ctx || !qY || !qU || !qV || !geom_ok(vg) || !qidx || !slice_bytes || !out) return fail(ctx, VC2_ERR_ARG);
  CU(cudaSetDevice(ctx->device));
  vc2_geom lg = *vg;
  lg.prefix = 0; lg.scalar = 1;
  SliceGeom g;
  fill_slice_geom(g, lg, nullptr);
  const int nslices = g.slices_x * g.slices_y;
  std::vector<uint32_t> fixed(nslices + 1, 0);
  for (int i = 0; i < nslices; ++i) {
    if (slice_bytes[i] < 1) return fail(ctx, VC2_ERR_ARG);
    fixed[i + 1] = fixed[i] + (uint32_t)slice_bytes[i];
  }
  const size_t total = fixed[nslices];
  if (out_len) *out_len = total;
  if (total > cap) return fail(ctx, VC2_ERR_CAPACITY);
  int st = upload_planes(ctx, g, qY, qU, qV);
  if (st) return st;
  int maxb = 0;
  for (int i = 0; i < nslices; ++i) maxb = std::max(maxb, slice_bytes[i]);
  const int wcap = std::max(staging_words(g), (maxb + 3) / 4 + 4);
  // tmp[2]: payload | slice_off | err | sizes | qidx | slice_bytes | fixed ; tmp[3]: staging
  const size_t o_off = (total + 259) / 256 * 256, o_err = o_off + (size_t)(nslices + 1) * 4, o_sz = o_err + (size_t)nslices * 4,
               o_q = o_sz + (size_t)nslices * 4, o_sb = o_q + (size_t)nslices * 4, o_fx = o_sb + (size_t)nslices * 4,
               o_end = o_fx + (size_t)(nslices + 1) * 4;
  CU(ctx->tmp[2].reserve(o_end));
  CU(ctx->tmp[3].reserve((size_t)wcap * 4 * nslices));
  uint8_t* base = ctx->tmp[2].as<uint8_t>();
  CU(cudaMemsetAsync(base + o_off, 0, o_q - o_off, ctx->stream));
  CU(cudaMemcpyAsync(base + o_q, qidx, (size_t)nslices * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(base + o_sb, slice_bytes, (size_t)nslices * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(base + o_fx, fixed.data(), (size_t)(nslices + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  LdEncParams p;
  memset(&p, 0, sizeof(p));
  p.g = g;
  p.coef = ctx->tmp[1].as<int32_t>(); p.qcoef = ctx->tmp[1].as<int32_t>(); p.prequantised = 1;
  p.qidx = (int32_t*)(base + o_q); p.slice_bytes = (const int32_t*)(base + o_sb);
  p.staging = ctx->tmp[3].as<uint32_t>(); p.wcap = wcap;
  p.sizes = (uint32_t*)(base + o_sz); p.err_flags = (uint32_t*)(base + o_err);
  CU(ld_pack_launch(ctx->stream, p, 1));
  ctx->launches++;
  AssembleParams a;
  memset(&a, 0, sizeof(a));
  a.nslices = nslices; a.sizes = p.sizes; a.fixed_off = (const uint32_t*)(base + o_fx);
  a.slice_off = (uint32_t*)(base + o_off); a.staging = p.staging; a.wcap = wcap;
  a.out = base; a.out_pic_stride = 0; a.out_capacity = (long long)total; a.err_flags = p.err_flags;
  CU(assemble_launch(ctx->stream, a, 1));
  ctx->launches += 2;
  std::vector<uint32_t> flags(nslices);
  CU(cudaMemcpyAsync(flags.data(), p.err_flags, (size_t)nslices * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  st = first_error(flags.data(), nslices, VC2_FLAG_VLC_RANGE);
  if (st) return fail(ctx, st);
  CU(cudaMemcpy(out, base, total, cudaMemcpyDeviceToHost));
  return VC2_OK;
}

// luma_slice_bits / chroma_slice_bits (Slices.cpp:51-96) for every slice of one plane (or of a U/V pair walked interleaved)
extern "C" int vc2_slice_bits(vc2_ctx* ctx, const int32_t* q, const int32_t* q2, int ph, int pw, int depth, int ny, int nx,
                              int32_t* bits) {
  if (!ctx || !q || !bits || depth < 1 || depth > VC2_MAX_DEPTH || ny < 1 || nx < 1 || ph < 1 || pw < 1) return fail(ctx, VC2_ERR_ARG);
  if (ph % ny || pw % nx || (ph / ny) % (1 << depth) || (pw / nx) % (1 << depth)) return fail(ctx, VC2_ERR_ARG);
  CU(cudaSetDevice(ctx->device));
  const size_t n = (size_t)ph * pw * 4;
  CU(ctx->tmp[0].reserve(n));
  CU(ctx->tmp[1].reserve(n));
  CU(ctx->tmp[2].reserve((size_t)ny * nx * 4));
  CU(cudaMemcpyAsync(ctx->tmp[0].p, q, n, cudaMemcpyHostToDevice, ctx->stream));
  if (q2) CU(cudaMemcpyAsync(ctx->tmp[1].p, q2, n, cudaMemcpyHostToDevice, ctx->stream));
  SliceBitsParams p;
  p.q = ctx->tmp[0].as<int32_t>(); p.q2 = q2 ? ctx->tmp[1].as<int32_t>() : nullptr;
  p.ph = ph; p.pw = pw; p.depth = depth; p.slices_y = ny; p.slices_x = nx; p.bits = ctx->tmp[2].as<int32_t>();
  CU(slice_bits_launch(ctx->stream, p));
  ctx->launches++;
  CU(cudaMemcpyAsync(bits, p.bits, (size_t)ny * nx * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return VC2_OK;
}

// component_slice_bytes (Slices.cpp:97-119) for every slice of one quantised in-place plane
extern "C" int vc2_hq_slice_sizes(vc2_ctx* ctx, const int32_t* q, int ph, int pw, int depth, int ny, int nx, int scalar,
                                  int32_t* bytes) {
  if (scalar < 1) return fail(ctx, VC2_ERR_ARG);
  const int st = vc2_slice_bits(ctx, q, nullptr, ph, pw, depth, ny, nx, bytes);
  if (st) return st;
  for (int i = 0; i < ny * nx; ++i) {
    const int scaled = ((bytes[i] + 7) / 8 + scalar - 1) / scalar;
    if (scaled > 0xFF) return fail(ctx, VC2_ERR_SCALAR_TOO_SMALL);
    bytes[i] = scaled * scalar;
  }
  return VC2_OK;
}

// quantise_transform(Array2D, qIndices, qMatrix) - the LD quantiser with DC prediction of the LL band
// (Quantisation.cpp:213-282, 353-367)
extern "C" int vc2_quantise_ld(vc2_ctx* ctx, const int32_t* coef, int ph, int pw, int depth, const int32_t* qmatrix,
                               const int32_t* qidx, int ny, int nx, int32_t* out) {
  return quant_host(ctx, coef, ph, pw, depth, qmatrix, qidx, ny, nx, out, 0, 1);
}

static int unpack_host(vc2_ctx* ctx, const uint8_t* in, size_t len, const vc2_geom* vg, const int32_t* slice_bytes, int ld,
                       int32_t* qY, int32_t* qU, int32_t* qV, int32_t* qidx) {
  if (!ctx || !in || !geom_ok(vg) || !qY || !qU || !qV || !qidx) return fail(ctx, VC2_ERR_ARG);
  if (ld && !slice_bytes) return fail(ctx, VC2_ERR_ARG);
  CU(cudaSetDevice(ctx->device));
  SliceGeom g;
  fill_slice_geom(g, *vg, nullptr);
  const int nslices = g.slices_x * g.slices_y;
  std::vector<uint32_t> offs(nslices + 1, 0);
  if (ld) {
    for (int i = 0; i < nslices; ++i) offs[i + 1] = offs[i] + (uint32_t)slice_bytes[i];
    if (offs[nslices] > len) return fail(ctx, VC2_ERR_STREAM);
  } else {
    const int st = vc2_hq_index_slices(in, len, nslices, g.prefix, g.scalar, offs.data());
    if (st) return fail(ctx, st);
  }
  const size_t o_off = (len + 259) / 256 * 256, o_err = o_off + (size_t)(nslices + 1) * 4, o_q = o_err + (size_t)nslices * 4,
               o_end = o_q + (size_t)nslices * 4;
  CU(ctx->tmp[2].reserve(o_end));
  CU(ctx->tmp[1].reserve((size_t)g.coef_pic_stride * 4));
  uint8_t* base = ctx->tmp[2].as<uint8_t>();
  CU(cudaMemcpyAsync(base, in, len, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(base + o_off, offs.data(), (size_t)(nslices + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(base + o_err, 0, (size_t)nslices * 8, ctx->stream));
  UnpackParams p;
  memset(&p, 0, sizeof(p));
  p.g = g; p.in = base; p.in_pic_stride = 0; p.slice_off = (const uint32_t*)(base + o_off); p.slice_off_pic_stride = 0;
  p.coef = ctx->tmp[1].as<int32_t>(); p.qidx = (int32_t*)(base + o_q); p.err_flags = (uint32_t*)(base + o_err);
  p.dequantise = 0; p.ld = ld;
  CU(unpack_launch(ctx->stream, p, 1));
  ctx->launches++;
  int st = download_planes(ctx, g, p.coef, qY, qU, qV);
  if (st) return st;
  std::vector<uint32_t> flags(nslices);
  CU(cudaMemcpyAsync(flags.data(), p.err_flags, (size_t)nslices * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(qidx, p.qidx, (size_t)nslices * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  st = first_error(flags.data(), nslices, VC2_FLAG_VLC_RANGE);
  if (st) return fail(ctx, st);
  return VC2_OK;
}

extern "C" int vc2_hq_unpack(vc2_ctx* ctx, const uint8_t* in, size_t len, const vc2_geom* g, int32_t* qY, int32_t* qU, int32_t* qV,
                             int32_t* qidx) {
  return unpack_host(ctx, in, len, g, nullptr, 0, qY, qU, qV, qidx);
}
extern "C" int vc2_ld_unpack(vc2_ctx* ctx, const uint8_t* in, size_t len, const vc2_geom* g, const int32_t* slice_bytes, int32_t* qY,
                             int32_t* qU, int32_t* qV, int32_t* qidx) {
  return unpack_host(ctx, in, len, g, slice_bytes, 1, qY, qU, qV, qidx);
}

// ================================================================================================
// fused, batched picture codec
// ================================================================================================

struct vc2_codec {
  vc2_ctx* ctx = nullptr;
  vc2_codec_params prm;
  SliceGeom g;
  int nslices = 0;
  int sample_kind = SAMPLE_U16BE;
  size_t pic_bytes = 0;          // raw planar bytes per picture
  size_t comp_bytes[3] = {0, 0, 0};
  size_t payload_cap = 0;        // per picture
  std::vector<int32_t> slice_bytes;   // CBR / LD
  std::vector<uint32_t> fixed_off;
  // device buffers
  DevBuf samples, recon, coef, scratch0, scratch1, payload, slice_off, err, qidx, staging, sizes, sbytes, fixed, tmp_plane, tmp_q;
  DevBuf dev_len;                     // [B] payload bytes of each slot: written by the encoder's scan or by upload_payload
  // narrow coefficient block (dwt.cuh): HQ_ConstQ encodes and HQ decodes keep 16-bit quantised coefficients between the
  // lifting kernels and the slice coders.  A magnitude beyond 15 bits raises narrow_ovf[slot]; whoever looks at the
  // slot's result next (status, downloads, taps, the host-buffer calls) first runs the slot again through the 32-bit path.
  bool narrow_enc = false, narrow_dec = false;
  DevBuf narrow_ovf;                  // [2][B]: overflow flags of the encodes, of the decodes
  DevBuf band_scale;                  // [B] BandScale of the decodes
  DevBuf tile_state;                  // fused scan + gather of the narrow packer (PackBuffers::tile_state), one region per slot
  struct SlotState {
    bool enc_unchecked = false;       // the payload comes from a narrow encode whose overflow flag has not been looked at
    bool dec_unchecked = false;       // the same for the reconstructed picture and a narrow decode
    bool dec_after_enc = false;       // a decode read that payload
    uint8_t coef = 0;                 // what the coefficient block holds: 0 = 32-bit, 1 = narrow (encode), 2 = narrow (decode)
  };
  std::vector<SlotState> slot_state;
  uint32_t* host_novf = nullptr;      // pinned [2 * B]: overflow flags of the host-buffer pipelines
  DevBuf ld_qcoef, ld_acbits, ld_restored;   // LD encoder only, allocated at its first use
  long long ld_ll_stride = 0;
  int ld_ll_off[3] = {0, 0, 0}, ld_ll_w[3] = {0, 0, 0};
  std::vector<uint32_t> len32;        // host copy the uploads are staged from
  long long scratch_stride[2] = {0, 0};
  long long scratch_off[2][3];
  std::vector<size_t> payload_len;    // host copy per slot (decode)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  // device-resident entry points: the batch is cut into sub-batches that run on their own streams, so the
  // small-grid kernels of one sub-batch (deep DWT levels, the slice scan) overlap the wide kernels of another
  static constexpr int MAX_SUB = 8;
  int nsub = 4;                        // VC2_CODEC_SUBBATCH
  cudaStream_t sub_stream[MAX_SUB] = {};   // [0] is only used in pipelined mode (else part 0 runs on the context stream)
  cudaEvent_t sub_fork = nullptr, sub_join[MAX_SUB] = {};
  cudaEvent_t sub_stagger[MAX_SUB] = {};
  // pipelined mode (vc2_codec_set_pipelined): consecutive device-resident calls are ordered per sub-batch only -
  // sub-batch p of a call waits for sub-batch p of the previous call (same stream, same slots), not for the whole
  // previous call - so the slice index walk of one sub-batch (a millisecond of one SM per picture) hides behind the
  // kernels of the others.  The context stream still joins every sub-batch of every call.
  bool pipelined = false;
  bool main_dirty = true;              // the context stream carries work the sub-batch streams have not waited for
  int last_split_n = -1;
  // software pipeline of the host-buffer entry points: one event triple per slot
  std::vector<cudaEvent_t> ev_in, ev_done, ev_out;
  // decode_host: the slice index of every payload is built on the device (hq_index_kernel, one CTA walking one
  // picture for about a millisecond) on a per-slot stream, between the payload upload and the parse kernel, so
  // the walks of all pictures in flight overlap each other and the other stages
  std::vector<cudaStream_t> index_stream;
  std::vector<cudaEvent_t> ev_idx_fork;   // [slot] device-resident decode: the sub-batch stream has reached its index walk
  bool index_prio = true;                 // run the device-resident index walks on the priority streams too
  std::vector<cudaEvent_t> ev_idx;
  uint32_t* host_flags = nullptr;     // pinned: per-slot error flags [B][nslices]
  uint32_t* host_len = nullptr;       // pinned: per-slot payload length [B]
};

static void codec_free(vc2_codec* k) {
  if (!k) return;
  cudaSetDevice(k->ctx->device);
  cudaStreamSynchronize(k->ctx->stream);
  DevBuf* all[] = {&k->samples, &k->recon, &k->coef, &k->scratch0, &k->scratch1, &k->payload, &k->slice_off, &k->err,
                   &k->qidx, &k->staging, &k->sizes, &k->sbytes, &k->fixed, &k->tmp_plane, &k->tmp_q, &k->dev_len,
                   &k->ld_qcoef, &k->ld_acbits, &k->ld_restored, &k->narrow_ovf, &k->tile_state, &k->band_scale};
  for (DevBuf* b : all) b->release();
  if (k->host_flags) cudaFreeHost(k->host_flags);
  if (k->host_len) cudaFreeHost(k->host_len);
  if (k->host_novf) cudaFreeHost(k->host_novf);
  for (auto e : k->ev_in) cudaEventDestroy(e);
  for (auto e : k->ev_done) cudaEventDestroy(e);
  for (auto e : k->ev_out) cudaEventDestroy(e);
  for (auto e : k->ev_idx) cudaEventDestroy(e);
  for (auto st : k->index_stream) cudaStreamDestroy(st);
  for (auto e : k->ev_idx_fork) cudaEventDestroy(e);
  if (k->copy_in) cudaStreamDestroy(k->copy_in);
  if (k->copy_out) cudaStreamDestroy(k->copy_out);
  for (int i = 0; i < vc2_codec::MAX_SUB; ++i) {
    if (k->sub_stream[i]) cudaStreamDestroy(k->sub_stream[i]);
    if (k->sub_join[i]) cudaEventDestroy(k->sub_join[i]);
    if (k->sub_stagger[i]) cudaEventDestroy(k->sub_stagger[i]);
  }
  if (k->sub_fork) cudaEventDestroy(k->sub_fork);
  delete k;
}

extern "C" vc2_codec* vc2_codec_create(vc2_ctx* ctx, const vc2_codec_params* prm) {
  if (!ctx || !prm || !geom_ok(&prm->geom) || prm->max_pictures < 1) { fail(ctx, VC2_ERR_ARG); return nullptr; }
  const vc2_sample_format& f = prm->fmt;
  if ((f.bytes_per_sample != 1 && f.bytes_per_sample != 2) || f.luma_depth < 1 || f.luma_depth > 8 * f.bytes_per_sample ||
      f.chroma_depth < 1 || f.chroma_depth > 8 * f.bytes_per_sample) { fail(ctx, VC2_ERR_ARG); return nullptr; }
  if (prm->mode == VC2_HQ_VBR && (prm->qindex < 0 || prm->qindex > 119)) { fail(ctx, VC2_ERR_ARG); return nullptr; }
  if (prm->mode != VC2_HQ_VBR && prm->picture_bytes < 1) { fail(ctx, VC2_ERR_ARG); return nullptr; }
  if (cudaSetDevice(ctx->device) != cudaSuccess) { fail(ctx, VC2_ERR_CUDA); return nullptr; }
  vc2_codec* k = new (std::nothrow) vc2_codec();
  if (!k) return nullptr;
  k->ctx = ctx;
  k->prm = *prm;
  int32_t qm[VC2_MAX_BANDS];
  vc2_quant_matrix(prm->geom.kernel, prm->geom.depth, qm);
  if (prm->mode == VC2_LD) { k->prm.geom.prefix = 0; k->prm.geom.scalar = 1; }
  fill_slice_geom(k->g, k->prm.geom, qm);
  const SliceGeom& g = k->g;
  k->nslices = g.slices_x * g.slices_y;
  k->sample_kind = f.bytes_per_sample == 2 ? SAMPLE_U16BE : SAMPLE_U8;
  for (int c = 0; c < 3; ++c) k->comp_bytes[c] = (size_t)g.plane[c].h * g.plane[c].w * f.bytes_per_sample;
  k->pic_bytes = k->comp_bytes[0] + k->comp_bytes[1] + k->comp_bytes[2];
  const int B = prm->max_pictures;
  if (prm->mode != VC2_HQ_VBR) {
    k->slice_bytes.resize(k->nslices);
    vc2_slice_bytes(g.slices_y, g.slices_x, prm->picture_bytes, prm->mode == VC2_LD ? 1 : g.scalar, k->slice_bytes.data());
    k->fixed_off.assign(k->nslices + 1, 0);
    for (int i = 0; i < k->nslices; ++i) {
      if (k->slice_bytes[i] < (prm->mode == VC2_LD ? 1 : 4)) { fail(ctx, VC2_ERR_ARG, "compressed bytes too small for this many slices"); delete k; return nullptr; }
      k->fixed_off[i + 1] = k->fixed_off[i] + (uint32_t)(k->slice_bytes[i] + (prm->mode == VC2_LD ? 0 : g.prefix));
    }
  }
  if (prm->mode == VC2_HQ_VBR) k->payload_cap = (size_t)max_slice_bytes(g, prm->mode, nullptr) * k->nslices;
  else k->payload_cap = k->fixed_off[k->nslices];
  k->payload_cap = (k->payload_cap + 255) / 256 * 256 + 256;

  bool ok = true;
  auto R = [&](DevBuf& b, size_t n) { if (ok && b.reserve(n) != cudaSuccess) ok = false; };
  R(k->samples, k->pic_bytes * B + 64);
  R(k->recon, k->pic_bytes * B + 64);
  R(k->coef, (size_t)g.coef_pic_stride * 4 * B);
  long long s0 = 0, s1 = 0;
  for (int c = 0; c < 3; ++c) {
    k->scratch_off[0][c] = s0; k->scratch_off[1][c] = s1;
    s0 += g.plane[c].size() / 4;
    s1 += (g.plane[c].size() / 16 + 4) & ~3ll;   // every plane starts on a 16-byte boundary
  }
  k->scratch_stride[0] = s0; k->scratch_stride[1] = s1;
  R(k->scratch0, (size_t)s0 * 4 * B);
  R(k->scratch1, (size_t)s1 * 4 * B);
  R(k->payload, k->payload_cap * B);
  R(k->slice_off, (size_t)(k->nslices + 1) * 4 * B);
  R(k->err, (size_t)k->nslices * 4 * B);
  R(k->qidx, (size_t)k->nslices * 4 * B);
  if (prm->mode != VC2_LD) R(k->staging, (size_t)staging_words(g) * 4 * k->nslices * B);
  R(k->sizes, (size_t)k->nslices * 4 * B);
  R(k->sbytes, (size_t)k->nslices * 4);
  R(k->fixed, (size_t)(k->nslices + 1) * 4);
  R(k->tmp_plane, (size_t)g.plane[0].size() * 4);
  R(k->tmp_q, (size_t)g.plane[0].size() * 4);
  R(k->dev_len, (size_t)B * 4);
  R(k->narrow_ovf, (size_t)B * 4 * 2);
  R(k->band_scale, sizeof(BandScale) * B);
  R(k->tile_state, tile_state_bytes(k->nslices, 1) * B);
  // (HQ_CBR keeps the 32-bit block: an unquantised 16-bit block for its rate control was measured 7 % SLOWER - the search
  // is bound by instruction issue, not by the bytes it reads, and the sign extension adds instructions)
  k->narrow_enc = ctx->narrow && ctx->dwt_tile && prm->mode == VC2_HQ_VBR;
  k->narrow_dec = ctx->narrow && ctx->dwt_tile && prm->mode != VC2_LD;
  k->slot_state.assign(B, vc2_codec::SlotState());
  if (ok && cudaMallocHost((void**)&k->host_novf, (size_t)4 * 2 * B) != cudaSuccess) ok = false;
  k->len32.assign(B, 0);
  if (ok && cudaMallocHost((void**)&k->host_flags, (size_t)k->nslices * 4 * 2 * B) != cudaSuccess) ok = false;
  if (ok && cudaMallocHost((void**)&k->host_len, (size_t)4 * B) != cudaSuccess) ok = false;
  for (int i = 0; ok && i < B; ++i) {
    cudaEvent_t e[3];
    for (int j = 0; j < 3; ++j) if (cudaEventCreateWithFlags(&e[j], cudaEventDisableTiming) != cudaSuccess) ok = false;
    if (ok) { k->ev_in.push_back(e[0]); k->ev_done.push_back(e[1]); k->ev_out.push_back(e[2]); }
  }
  for (int i = 0; ok && i < B; ++i) {   // decode_host keeps two chunks in flight: a second set of "picture has left" events
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) ok = false; else k->ev_out.push_back(e);
  }
  for (int i = 0; ok && i < B; ++i) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) ok = false; else k->ev_idx.push_back(e);
  }
  // the slice index walks run on streams of the highest priority: a walk is 1.4 ms of latency on one SM per picture, its CTAs
  // should take the first SM that has room instead of queueing behind the thousands of CTAs of the lifting kernels
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  for (int i = 0; ok && i < std::min(B, 8); ++i) {
    cudaStream_t st;
    if (cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_hi) != cudaSuccess) ok = false; else k->index_stream.push_back(st);
  }
  for (int i = 0; ok && i < B; ++i) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) ok = false; else k->ev_idx_fork.push_back(e);
  }
  k->index_prio = getenv("VC2_INDEX_PRIO") ? atoi(getenv("VC2_INDEX_PRIO")) != 0 : true;
  if (ok && cudaStreamCreateWithFlags(&k->copy_in, cudaStreamNonBlocking) != cudaSuccess) ok = false;
  if (ok && cudaStreamCreateWithFlags(&k->copy_out, cudaStreamNonBlocking) != cudaSuccess) ok = false;
  if (const char* e = getenv("VC2_CODEC_SUBBATCH")) k->nsub = std::min(std::max(atoi(e), 1), (int)vc2_codec::MAX_SUB);
  if (ok && cudaEventCreateWithFlags(&k->sub_fork, cudaEventDisableTiming) != cudaSuccess) ok = false;
  for (int i = 0; ok && i < k->nsub; ++i) {
    if (cudaStreamCreateWithFlags(&k->sub_stream[i], cudaStreamNonBlocking) != cudaSuccess) ok = false;
    if (ok && cudaEventCreateWithFlags(&k->sub_join[i], cudaEventDisableTiming) != cudaSuccess) ok = false;
    if (ok && cudaEventCreateWithFlags(&k->sub_stagger[i], cudaEventDisableTiming) != cudaSuccess) ok = false;
  }
  if (ok && !k->slice_bytes.empty()) {
    ok = cudaMemcpy(k->sbytes.p, k->slice_bytes.data(), (size_t)k->nslices * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(k->fixed.p, k->fixed_off.data(), (size_t)(k->nslices + 1) * 4, cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (ok) ok = cudaMemset(k->err.p, 0, (size_t)k->nslices * 4 * B) == cudaSuccess;
  if (ok) ok = cudaMemset(k->dev_len.p, 0, (size_t)B * 4) == cudaSuccess;
  if (ok) ok = cudaMemset(k->narrow_ovf.p, 0, (size_t)B * 4 * 2) == cudaSuccess;
  if (!ok) { fail(ctx, VC2_ERR_CUDA, "codec allocation failed"); codec_free(k); return nullptr; }
  k->payload_len.assign(B, 0);
  return k;
}

extern "C" void vc2_codec_destroy(vc2_codec* k) { codec_free(k); }
extern "C" size_t vc2_codec_picture_in_bytes(const vc2_codec* k) { return k ? k->pic_bytes : 0; }
extern "C" size_t vc2_codec_payload_capacity(const vc2_codec* k) { return k ? k->payload_cap : 0; }

static void codec_compbufs(vc2_codec* k, CompBuf cb[3], int first_slot, bool recon) {
  const SliceGeom& g = k->g;
  const vc2_sample_format& f = k->prm.fmt;
  size_t off = 0;
  for (int c = 0; c < 3; ++c) {
    CompBuf& B = cb[c];
    B.pg = g.plane[c];
    B.pix = (recon ? k->recon : k->samples).as<uint8_t>() + (size_t)first_slot * k->pic_bytes + off;
    B.pix_pic_stride = (long long)k->pic_bytes;
    B.pix_pitch = g.plane[c].w;
    off += k->comp_bytes[c];
    B.comp = c;
    B.coef = k->coef.as<int32_t>() + (long long)first_slot * g.coef_pic_stride;
    B.coef_pic_stride = g.coef_pic_stride;
    B.scratch[0] = k->scratch0.as<int32_t>() + (long long)first_slot * k->scratch_stride[0] + k->scratch_off[0][c];
    B.scratch[1] = k->scratch1.as<int32_t>() + (long long)first_slot * k->scratch_stride[1] + k->scratch_off[1][c];
    B.scratch_pic_stride[0] = k->scratch_stride[0];
    B.scratch_pic_stride[1] = k->scratch_stride[1];
    const int depth = c == 0 ? f.luma_depth : f.chroma_depth;
    B.sshift = 8 * f.bytes_per_sample - depth;      // Arrays.cpp:170-172 (left justified)
    B.soffset = 1 << (depth - 1);                   // Arrays.cpp:174-178 (offset binary)
    B.clip_min = -(1 << (depth - 1));               // DecodeStream.cpp:593-596
    B.clip_max = (1 << (depth - 1)) - 1;
  }
}

// LD encoder buffers: quantised coefficients, the per-slice bit tables of every index, the locally decoded LL bands
static int codec_ld_buffers(vc2_codec* k) {
  vc2_ctx* ctx = k->ctx;
  if (k->ld_qcoef.p) return VC2_OK;
  const SliceGeom& g = k->g;
  const int B = k->prm.max_pictures;
  long long off = 0;
  for (int c = 0; c < 3; ++c) {
    k->ld_ll_off[c] = (int)off;
    k->ld_ll_w[c] = g.plane[c].pw >> g.depth;
    off += (long long)(g.plane[c].ph >> g.depth) * (g.plane[c].pw >> g.depth);
  }
  k->ld_ll_stride = off;
  CU(k->ld_qcoef.reserve((size_t)g.coef_pic_stride * 4 * B));
  CU(k->ld_acbits.reserve((size_t)k->nslices * 256 * 4 * B));
  CU(k->ld_restored.reserve((size_t)off * 4 * B));
  CU(k->staging.reserve((size_t)staging_words(g) * 4 * k->nslices * B));
  return VC2_OK;
}

static int codec_encode_range(vc2_codec* k, int first, int n, bool wide = false) {
  vc2_ctx* ctx = k->ctx;
  CompBuf cb[3];
  codec_compbufs(k, cb, first, false);
  const bool narrow = k->narrow_enc && !wide;
  NarrowCfg nw;
  if (narrow) {
    nw.on = 1; nw.qindex = k->prm.qindex; nw.ovf = k->narrow_ovf.as<uint32_t>() + first;
    CU(cudaMemsetAsync(nw.ovf, 0, (size_t)n * 4, ctx->stream));
  }
  for (int i = 0; i < n; ++i) {
    vc2_codec::SlotState& ss = k->slot_state[first + i];
    ss.enc_unchecked = narrow; ss.dec_after_enc = false; ss.coef = narrow ? 1 : 0;
  }
  CU(run_dwt(ctx, false, k->prm.geom.kernel, k->prm.geom.depth, k->sample_kind, k->g, cb, 3, n, nw));
  if (k->prm.mode == VC2_LD) {
    // EncodeStream.cpp:141-245 (rate control), Quantisation.cpp:213-282 (predictive quantiser), Slices.cpp:195-244 (writer)
    const int st = codec_ld_buffers(k);
    if (st) return st;
    LdEncParams p;
    memset(&p, 0, sizeof(p));
    p.g = k->g;
    p.coef = k->coef.as<int32_t>() + (long long)first * k->g.coef_pic_stride;
    p.qcoef = k->ld_qcoef.as<int32_t>() + (long long)first * k->g.coef_pic_stride;
    p.acbits = k->ld_acbits.as<uint32_t>() + (size_t)first * k->nslices * 256;
    p.qidx = k->qidx.as<int32_t>() + (size_t)first * k->nslices;
    p.restored = k->ld_restored.as<int32_t>() + (long long)first * k->ld_ll_stride;
    p.ll_stride = k->ld_ll_stride;
    for (int c = 0; c < 3; ++c) { p.ll_off[c] = k->ld_ll_off[c]; p.ll_w[c] = k->ld_ll_w[c]; }
    p.slice_bytes = k->sbytes.as<int32_t>();
    p.staging = k->staging.as<uint32_t>() + (size_t)first * k->nslices * staging_words(k->g);
    p.wcap = staging_words(k->g);
    p.sizes = k->sizes.as<uint32_t>() + (size_t)first * k->nslices;
    p.err_flags = k->err.as<uint32_t>() + (size_t)first * k->nslices;
    {
      ProfScope ps(ctx, VC2_STAGE_PACK);
      CU(ld_encode_launch(ctx->stream, p, n));
    }
    ctx->launches += 3;
    AssembleParams a;
    memset(&a, 0, sizeof(a));
    a.nslices = k->nslices;
    a.sizes = p.sizes; a.fixed_off = k->fixed.as<uint32_t>();
    a.slice_off = k->slice_off.as<uint32_t>() + (size_t)first * (k->nslices + 1);
    a.total_len = k->dev_len.as<uint32_t>() + first;
    a.staging = p.staging; a.wcap = p.wcap;
    a.out = k->payload.as<uint8_t>() + (size_t)first * k->payload_cap;
    a.out_pic_stride = (long long)k->payload_cap; a.out_capacity = (long long)k->payload_cap; a.err_flags = p.err_flags;
    {
      ProfScope ps(ctx, VC2_STAGE_ASSEMBLE);
      CU(assemble_launch(ctx->stream, a, n));
    }
    ctx->launches += 2;
    return VC2_OK;
  }
  PackBuffers B;
  B.out = k->payload.as<uint8_t>() + (size_t)first * k->payload_cap;
  B.out_stride = (long long)k->payload_cap; B.out_capacity = (long long)k->payload_cap;
  B.slice_off = k->slice_off.as<uint32_t>() + (size_t)first * (k->nslices + 1);
  B.err_flags = k->err.as<uint32_t>() + (size_t)first * k->nslices;
  B.qidx = k->qidx.as<int32_t>() + (size_t)first * k->nslices;
  B.staging = k->staging.as<uint32_t>() + (size_t)first * k->nslices * staging_words(k->g);
  B.sizes = k->sizes.as<uint32_t>() + (size_t)first * k->nslices;
  const bool cbr = k->prm.mode == VC2_HQ_CBR;
  B.slice_bytes_dev = cbr ? k->sbytes.as<int32_t>() : nullptr;
  B.fixed_off_dev = cbr ? k->fixed.as<uint32_t>() : nullptr;
  B.total_len = k->dev_len.as<uint32_t>() + first;
  B.tile_state = reinterpret_cast<unsigned long long*>(k->tile_state.as<uint8_t>() + tile_state_bytes(k->nslices, 1) * first);
  const int32_t* coef = k->coef.as<int32_t>() + (long long)first * k->g.coef_pic_stride;
  CU(run_pack(ctx, k->g, coef, n, k->prm.mode, 1, cbr ? 1 : 0, cbr ? -1 : k->prm.qindex, 1, B, narrow ? nw.on : 0));
  return VC2_OK;
}

// run fn(first, count) over the batch: whole on the context stream, or - when the batch is big enough and the
// per-kernel profiler is off (its event pairs would time overlapped kernels) - as sub-batches on forked streams
// that are joined back into the context stream
template <class Fn>
static int codec_run_split(vc2_codec* k, int n, Fn fn) {
  vc2_ctx* ctx = k->ctx;
  const bool pipe = k->pipelined && !ctx->profiling;
  const int parts = ctx->profiling ? 1 : std::max(1, std::min(k->nsub, n / 2));
  if (parts <= 1 && !pipe) { k->main_dirty = true; k->last_split_n = -1; return fn(0, n); }
  cudaStream_t main_stream = ctx->stream;
  // the fork: sub-batch streams wait for what the context stream carries.  In pipelined mode that is only needed
  // when something other than the previous split call (same n, same slot -> stream map) went onto it
  const bool fork = !pipe || k->main_dirty || k->last_split_n != n;
  if (fork) CU(cudaEventRecord(k->sub_fork, main_stream));
  int st = VC2_OK;
  for (int p = 0, first = 0; p < parts && st == VC2_OK; ++p) {
    const int cnt = n / parts + (p < n % parts ? 1 : 0);
    cudaStream_t s = (p > 0 || pipe) ? k->sub_stream[p] : main_stream;
    if (s != main_stream) {
      if (fork && cudaStreamWaitEvent(s, k->sub_fork, 0) != cudaSuccess) { st = VC2_ERR_CUDA; break; }
      // pipelined mode, first call after a fork: run the sub-batches one after the other.  The later, unforked
      // calls then keep that phase shift, so the streams are never all inside their index walk at the same time
      if (pipe && fork && p > 0 && cudaStreamWaitEvent(s, k->sub_stagger[p - 1], 0) != cudaSuccess) { st = VC2_ERR_CUDA; break; }
      ctx->stream = s;
    }
    st = fn(first, cnt);
    ctx->stream = main_stream;
    if (pipe && fork && cudaEventRecord(k->sub_stagger[p], s) != cudaSuccess) st = st == VC2_OK ? VC2_ERR_CUDA : st;
    if (s != main_stream) {
      // always join, also after a failure, so the context stream never runs ahead of a sub-batch
      if (cudaEventRecord(k->sub_join[p], s) != cudaSuccess ||
          cudaStreamWaitEvent(main_stream, k->sub_join[p], 0) != cudaSuccess) st = st == VC2_OK ? VC2_ERR_CUDA : st;
    }
    first += cnt;
  }
  k->main_dirty = !pipe;
  k->last_split_n = pipe ? n : -1;
  if (st == VC2_ERR_CUDA) return fail(ctx, VC2_ERR_CUDA, "sub-batch stream");
  return st;
}

extern "C" int vc2_codec_set_pipelined(vc2_codec* k, int on) {
  if (!k) return VC2_ERR_ARG;
  k->pipelined = on != 0;
  k->main_dirty = true;
  return VC2_OK;
}

extern "C" int vc2_codec_encode_dev(vc2_codec* k, int n) {
  if (!k || n < 1 || n > k->prm.max_pictures) return fail(k ? k->ctx : nullptr, VC2_ERR_ARG);
  vc2_ctx* ctx = k->ctx;
  CU(cudaSetDevice(ctx->device));
  return codec_run_split(k, n, [&](int first, int cnt) { return codec_encode_range(k, first, cnt); });
}

static int codec_decode_range(vc2_codec* k, int first, int n, bool index_here, bool wide = false) {
  vc2_ctx* ctx = k->ctx;
  const SliceGeom& g = k->g;
  const bool ld = k->prm.mode == VC2_LD;
  const bool narrow = k->narrow_dec && !wide;
  NarrowCfg nw;
  if (narrow) {
    nw.on = 1; nw.qidx = k->qidx.as<int32_t>() + (size_t)first * k->nslices; nw.ovf = k->narrow_ovf.as<uint32_t>() + k->prm.max_pictures + first;
    BandScale* bs = k->band_scale.as<BandScale>() + first;
    nw.band_scale = bs;
    CU(cudaMemsetAsync(nw.ovf, 0, (size_t)n * 4, ctx->stream));
    CU(cudaMemsetAsync(bs, 0, sizeof(BandScale) * n, ctx->stream));
  }
  for (int i = 0; i < n; ++i) {
    vc2_codec::SlotState& ss = k->slot_state[first + i];
    ss.dec_unchecked = narrow; ss.dec_after_enc = true; ss.coef = narrow ? 2 : 0;
  }
  const bool shared_off = ld;   // HQ pictures are always indexed from their own length bytes (DecodeStream.cpp:512)
  CU(cudaMemsetAsync(k->err.as<uint32_t>() + (size_t)first * k->nslices, 0, (size_t)k->nslices * 4 * n, ctx->stream));
  UnpackParams p;
  memset(&p, 0, sizeof(p));
  p.g = g;
  p.in = k->payload.as<uint8_t>() + (size_t)first * k->payload_cap;
  p.in_pic_stride = (long long)k->payload_cap;
  p.slice_off = shared_off ? k->fixed.as<uint32_t>() : k->slice_off.as<uint32_t>() + (size_t)first * (k->nslices + 1);
  p.slice_off_pic_stride = shared_off ? 0 : (k->nslices + 1);
  p.coef = k->coef.as<int32_t>() + (long long)first * g.coef_pic_stride;
  p.qidx = k->qidx.as<int32_t>() + (size_t)first * k->nslices;
  p.err_flags = k->err.as<uint32_t>() + (size_t)first * k->nslices;
  p.dequantise = 1;
  p.ld = ld ? 1 : 0;
  p.narrow = narrow ? 1 : 0;
  p.narrow_ovf = nw.ovf;
  p.band_scale = const_cast<BandScale*>(nw.band_scale);
  if (index_here && !ld) {
    // the reader's walk over the length bytes of every slice (Slices.cpp:544-605): the slice offsets are always
    // derived from the payload itself, never taken from the encoder
    IndexParams ip;
    memset(&ip, 0, sizeof(ip));
    ip.in = p.in; ip.in_pic_stride = p.in_pic_stride;
    ip.len_dev = k->dev_len.as<uint32_t>() + first;
    ip.slice_off = k->slice_off.as<uint32_t>() + (size_t)first * (k->nslices + 1);
    ip.nslices = k->nslices; ip.prefix = g.prefix; ip.scalar = g.scalar;
    if (k->index_prio && !ctx->profiling && !k->index_stream.empty()) {
      // on a priority stream, between two events of the sub-batch's own stream
      const int B = k->prm.max_pictures;
      cudaStream_t is = k->index_stream[(size_t)((long long)first * (long long)k->index_stream.size() / B) % k->index_stream.size()];
      CU(cudaEventRecord(k->ev_idx_fork[first], ctx->stream));
      CU(cudaStreamWaitEvent(is, k->ev_idx_fork[first], 0));
      CU(index_launch(is, ip, n));
      CU(cudaEventRecord(k->ev_idx[first], is));
      CU(cudaStreamWaitEvent(ctx->stream, k->ev_idx[first], 0));
    } else {
      ProfScope ps(ctx, VC2_STAGE_INDEX);
      CU(index_launch(ctx->stream, ip, n));
    }
    ctx->launches++;
  }
  {
    ProfScope ps(ctx, VC2_STAGE_UNPACK);
    CU(unpack_launch(ctx->stream, p, n, ctx->scale_tab.as<uint2>()));
  }
  ctx->launches += narrow ? 2 : 1;
  if (ld) {
    // LL band: DC-predicted reconstruction, one wavefront CTA per (picture, component), the whole batch in one launch
    LdDcBatch bt;
    memset(&bt, 0, sizeof(bt));
    bt.nplanes = 3;
    for (int c = 0; c < 3; ++c) {
      LdDcParams& d = bt.c[c];
      d.base = p.coef;
      d.qidx = p.qidx;
      d.base_pic_stride = g.coef_pic_stride; d.qidx_pic_stride = k->nslices;
      d.H = g.plane[c].ph >> g.depth; d.W = g.plane[c].pw >> g.depth;
      d.interleaved = 1; d.bh = g.part_h[c][0]; d.bw = g.part_w[c][0];
      d.k0 = g.comp_start[c]; d.nc4 = g.comp_start[3] >> 2;   // band 0 starts each component
      d.slices_y = g.slices_y; d.slices_x = g.slices_x; d.qm0 = g.qmatrix[0];
    }
    {
      ProfScope ps(ctx, VC2_STAGE_LD_DC);
      CU(ld_dc_batch_launch(ctx->stream, bt, n));
    }
    ctx->launches++;
  }
  CompBuf cb[3];
  codec_compbufs(k, cb, first, true);
  CU(run_dwt(ctx, true, k->prm.geom.kernel, k->prm.geom.depth, k->sample_kind, k->g, cb, 3, n, nw));
  return VC2_OK;
}

// Slots that went through the narrow block: did a coefficient overflow it?  Then the slot runs again through the 32-bit
// path - the inputs (samples, or payload and slice offsets) are still in the slot - and a decode that consumed the payload
// of such an encode is repeated as well.  Synchronises the stream when there is something to look at.
static int codec_fix_narrow(vc2_codec* k, int slot) {
  vc2_ctx* ctx = k->ctx;
  vc2_codec::SlotState& ss = k->slot_state[slot];
  const int B = k->prm.max_pictures;
  if (ss.enc_unchecked) {
    uint32_t ovf = 0;
    CU(cudaMemcpyAsync(&ovf, k->narrow_ovf.as<uint32_t>() + slot, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ss.enc_unchecked = false;
    if (ovf) {
      k->main_dirty = true;
      const bool again = ss.dec_after_enc;
      int st = codec_encode_range(k, slot, 1, true);
      if (st) return st;
      if (again) {
        st = codec_decode_range(k, slot, 1, true);
        if (st) return st;
      }
    }
  }
  if (ss.dec_unchecked) {
    uint32_t ovf = 0;
    CU(cudaMemcpyAsync(&ovf, k->narrow_ovf.as<uint32_t>() + B + slot, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ss.dec_unchecked = false;
    if (ovf) {
      k->main_dirty = true;
      const int st = codec_decode_range(k, slot, 1, false, true);
      if (st) return st;
    }
  }
  return VC2_OK;
}
// the coefficient taps read the 32-bit block: rebuild it for a slot that was run narrow (forward transform, or parse + scale)
static int codec_make_wide(vc2_codec* k, int slot) {
  vc2_ctx* ctx = k->ctx;
  int st = codec_fix_narrow(k, slot);
  if (st) return st;
  const int state = k->slot_state[slot].coef;
  if (state == 0) return VC2_OK;
  k->main_dirty = true;
  if (state == 1) {
    CompBuf cb[3];
    codec_compbufs(k, cb, slot, false);
    CU(run_dwt(ctx, false, k->prm.geom.kernel, k->prm.geom.depth, k->sample_kind, k->g, cb, 3, 1));
  } else {
    const SliceGeom& g = k->g;
    UnpackParams p;
    memset(&p, 0, sizeof(p));
    p.g = g;
    p.in = k->payload.as<uint8_t>() + (size_t)slot * k->payload_cap;
    p.in_pic_stride = (long long)k->payload_cap;
    p.slice_off = k->slice_off.as<uint32_t>() + (size_t)slot * (k->nslices + 1);
    p.slice_off_pic_stride = k->nslices + 1;
    p.coef = k->coef.as<int32_t>() + (long long)slot * g.coef_pic_stride;
    p.qidx = k->qidx.as<int32_t>() + (size_t)slot * k->nslices;
    p.err_flags = k->err.as<uint32_t>() + (size_t)slot * k->nslices;
    p.dequantise = 1;
    CU(unpack_launch(ctx->stream, p, 1));
    ctx->launches++;
  }
  k->slot_state[slot].coef = 0;
  return VC2_OK;
}

extern "C" int vc2_codec_decode_dev(vc2_codec* k, int n) {
  if (!k || n < 1 || n > k->prm.max_pictures) return fail(k ? k->ctx : nullptr, VC2_ERR_ARG);
  vc2_ctx* ctx = k->ctx;
  CU(cudaSetDevice(ctx->device));
  return codec_run_split(k, n, [&](int first, int cnt) { return codec_decode_range(k, first, cnt, true); });
}

extern "C" void* vc2_codec_samples_dev(vc2_codec* k, int slot) {
  return (k && slot >= 0 && slot < k->prm.max_pictures) ? k->samples.as<uint8_t>() + (size_t)slot * k->pic_bytes : nullptr;
}
extern "C" void* vc2_codec_recon_dev(vc2_codec* k, int slot) {
  return (k && slot >= 0 && slot < k->prm.max_pictures) ? k->recon.as<uint8_t>() + (size_t)slot * k->pic_bytes : nullptr;
}
extern "C" uint8_t* vc2_codec_payload_dev(vc2_codec* k, int slot) {
  return (k && slot >= 0 && slot < k->prm.max_pictures) ? k->payload.as<uint8_t>() + (size_t)slot * k->payload_cap : nullptr;
}
extern "C" int32_t* vc2_codec_coeffs_dev(vc2_codec* k, int slot) {
  if (!k || slot < 0 || slot >= k->prm.max_pictures) return nullptr;
  return k->coef.as<int32_t>() + (long long)slot * k->g.coef_pic_stride;
}
extern "C" uint32_t* vc2_codec_slice_offsets_dev(vc2_codec* k, int slot) {
  return (k && slot >= 0 && slot < k->prm.max_pictures) ? k->slice_off.as<uint32_t>() + (size_t)slot * (k->nslices + 1) : nullptr;
}

#define KARG(cond)                                              \
  do {                                                          \
    if (!(cond)) return fail(k ? k->ctx : nullptr, VC2_ERR_ARG); \
  } while (0)

extern "C" int vc2_codec_upload_picture(vc2_codec* k, int slot, const void* raw) {
  KARG(k && raw && slot >= 0 && slot < k->prm.max_pictures);
  vc2_ctx* ctx = k->ctx;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(vc2_codec_samples_dev(k, slot), raw, k->pic_bytes, cudaMemcpyHostToDevice, ctx->stream));
  k->main_dirty = true;
  return VC2_OK;
}
extern "C" int vc2_codec_download_picture(vc2_codec* k, int slot, void* raw) {
  KARG(k && raw && slot >= 0 && slot < k->prm.max_pictures);
  vc2_ctx* ctx = k->ctx;
  CU(cudaSetDevice(ctx->device));
  { const int st = codec_fix_narrow(k, slot); if (st) return st; }
  CU(cudaMemcpyAsync(raw, vc2_codec_recon_dev(k, slot), k->pic_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return VC2_OK;
}

extern "C" int vc2_codec_upload_payload(vc2_codec* k, int slot, const uint8_t* payload, size_t len) {
  KARG(k && payload && slot >= 0 && slot < k->prm.max_pictures);
  vc2_ctx* ctx = k->ctx;
  if (len > k->payload_cap) return fail(ctx, VC2_ERR_CAPACITY);
  if (k->prm.mode == VC2_LD && len < k->fixed_off[k->nslices]) return fail(ctx, VC2_ERR_STREAM);
  CU(cudaSetDevice(ctx->device));
  // bytes and length only: decode_dev walks the slice length bytes on the device (hq_index_kernel)
  CU(cudaMemcpyAsync(vc2_codec_payload_dev(k, slot), payload, len, cudaMemcpyHostToDevice, ctx->stream));
  k->payload_len[slot] = len;
  k->slot_state[slot].enc_unchecked = false;   // the payload is the caller's now
  k->len32[slot] = (uint32_t)len;
  CU(cudaMemcpyAsync(k->dev_len.as<uint32_t>() + slot, &k->len32[slot], 4, cudaMemcpyHostToDevice, ctx->stream));
  k->main_dirty = true;
  return VC2_OK;
}

static int codec_slot_status(vc2_codec* k, int slot, uint32_t ignore) {
  KARG(k && slot >= 0 && slot < k->prm.max_pictures);
  vc2_ctx* ctx = k->ctx;
  CU(cudaSetDevice(ctx->device));
  { const int st = codec_fix_narrow(k, slot); if (st) return st; }
  std::vector<uint32_t> flags(k->nslices);
  CU(cudaMemcpyAsync(flags.data(), k->err.as<uint32_t>() + (size_t)slot * k->nslices, (size_t)k->nslices * 4,
                     cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const int st = first_error(flags.data(), k->nslices, ignore);
  if (st) return fail(ctx, st);
  return VC2_OK;
}
extern "C" int vc2_codec_slot_status(vc2_codec* k, int slot) { return codec_slot_status(k, slot, 0); }

extern "C" int vc2_codec_download_payload(vc2_codec* k, int slot, uint8_t* payload, size_t cap, size_t* len, int32_t* qidx,
                                          uint32_t* slice_off) {
  KARG(k && payload && len && slot >= 0 && slot < k->prm.max_pictures);
  vc2_ctx* ctx = k->ctx;
  const int st = vc2_codec_slot_status(k, slot);
  if (st) return st;
  std::vector<uint32_t> offs(k->nslices + 1);
  CU(cudaMemcpy(offs.data(), vc2_codec_slice_offsets_dev(k, slot), (size_t)(k->nslices + 1) * 4, cudaMemcpyDeviceToHost));
  *len = offs[k->nslices];
  if (*len > cap) return fail(ctx, VC2_ERR_CAPACITY);
  CU(cudaMemcpy(payload, vc2_codec_payload_dev(k, slot), *len, cudaMemcpyDeviceToHost));
  if (qidx) CU(cudaMemcpy(qidx, k->qidx.as<int32_t>() + (size_t)slot * k->nslices, (size_t)k->nslices * 4, cudaMemcpyDeviceToHost));
  if (slice_off) memcpy(slice_off, offs.data(), (size_t)(k->nslices + 1) * 4);
  return VC2_OK;
}

extern "C" int vc2_codec_read_transform(vc2_codec* k, int slot, int32_t* y, int32_t* u, int32_t* v) {
  KARG(k && slot >= 0 && slot < k->prm.max_pictures);
  vc2_ctx* ctx = k->ctx;
  CU(cudaSetDevice(ctx->device));
  int32_t* dst[3] = {y, u, v};
  { const int st = codec_make_wide(k, slot); if (st) return st; }
  for (int c = 0; c < 3; ++c) {
    if (!dst[c]) continue;
    CU(layout_launch(ctx->stream, false, vc2_codec_coeffs_dev(k, slot), k->tmp_plane.as<int32_t>(), k->g, c));
    ctx->launches++;
    CU(cudaMemcpyAsync(dst[c], k->tmp_plane.p, (size_t)k->g.plane[c].size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return VC2_OK;
}

extern "C" int vc2_codec_read_quantised(vc2_codec* k, int slot, int32_t* y, int32_t* u, int32_t* v) {
  KARG(k && slot >= 0 && slot < k->prm.max_pictures);
  vc2_ctx* ctx = k->ctx;
  CU(cudaSetDevice(ctx->device));
  int32_t* dst[3] = {y, u, v};
  { const int st = codec_make_wide(k, slot); if (st) return st; }
  for (int c = 0; c < 3; ++c) {
    if (!dst[c]) continue;
    const PlaneGeom& pg = k->g.plane[c];
    if (k->prm.mode == VC2_LD) {
      // the LD encoder keeps its quantised coefficients (LL band: prediction residuals, Quantisation.cpp:213-231)
      if (!k->ld_qcoef.p) return fail(ctx, VC2_ERR_ARG, "no LD picture has been encoded");
      CU(layout_launch(ctx->stream, false, k->ld_qcoef.as<int32_t>() + (long long)slot * k->g.coef_pic_stride, k->tmp_plane.as<int32_t>(), k->g, c));
      ctx->launches++;
      CU(cudaMemcpyAsync(dst[c], k->tmp_plane.p, (size_t)pg.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      continue;
    }
    CU(layout_launch(ctx->stream, false, vc2_codec_coeffs_dev(k, slot), k->tmp_plane.as<int32_t>(), k->g, c));
    QuantParams p;
    memset(&p, 0, sizeof(p));
    p.src = k->tmp_plane.as<int32_t>(); p.dst = k->tmp_q.as<int32_t>();
    p.qidx = k->qidx.as<int32_t>() + (size_t)slot * k->nslices;
    p.ph = pg.ph; p.pw = pg.pw; p.depth = pg.depth; p.slices_y = k->g.slices_y; p.slices_x = k->g.slices_x;
    for (int b = 0; b < k->g.nbands; ++b) p.qmatrix[b] = k->g.qmatrix[b];
    CU(quant_launch(ctx->stream, p));
    ctx->launches += 2;
    CU(cudaMemcpyAsync(dst[c], k->tmp_q.p, (size_t)pg.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return VC2_OK;
}

extern "C" int vc2_codec_read_indices(vc2_codec* k, int slot, int32_t* qidx) {
  KARG(k && qidx && slot >= 0 && slot < k->prm.max_pictures);
  vc2_ctx* ctx = k->ctx;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(qidx, k->qidx.as<int32_t>() + (size_t)slot * k->nslices, (size_t)k->nslices * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return VC2_OK;
}

// ---- end to end with host buffers -------------------------------------------------------------------
// Software pipeline over the B = max_pictures slots, one picture per stage: the H2D copy of picture i+1
// (copy_in stream), the kernels of picture i (context stream) and the D2H copy of picture i-1 (copy_out
// stream) overlap; streams are ordered with events, the host only blocks to learn a payload length
// (encode) or to reuse a pinned staging table (decode).
// pictures per pipeline stage: kernels on a single picture leave most of the GPU idle, so once the batch has
// room for it a stage carries two pictures (the copy of a stage still overlaps the kernels of the previous one)
static int stage_pictures(int B) { return (B >= 4 && B % 2 == 0) ? 2 : 1; }

extern "C" int vc2_codec_encode_host(vc2_codec* k, int n, const void* const* pictures, uint8_t* const* payloads, size_t cap,
                                     size_t* payload_len) {
  KARG(k && pictures && payloads && payload_len && n >= 1);
  vc2_ctx* ctx = k->ctx;
  CU(cudaSetDevice(ctx->device));
  k->main_dirty = true;
  const int B = k->prm.max_pictures, ns = k->nslices;
  const int sub = stage_pictures(B), nstage_slots = B / sub;
  const int lag = std::min(2, nstage_slots - 1);   // stages in flight behind the newest upload
  const int nstages = (n + sub - 1) / sub;
  int status = VC2_OK;
  std::vector<int> redo;                // pictures that overflowed the narrow coefficient block: encoded again at the end, 32-bit
  auto drain = [&](int st) -> int {     // stage st: wait for its kernels, then send its payloads home
    const int slot0 = (st % nstage_slots) * sub, first = st * sub, m = std::min(sub, n - first);
    cudaError_t e = cudaEventSynchronize(k->ev_done[slot0]);
    if (e != cudaSuccess) return cuda_fail(ctx, e);
    for (int j = 0; j < m; ++j) {
      if (k->narrow_enc && k->host_novf[slot0 + j]) { redo.push_back(first + j); payload_len[first + j] = 0; continue; }
      const int er = first_error(k->host_flags + (size_t)(slot0 + j) * ns, ns);
      if (er) return fail(ctx, er);
      payload_len[first + j] = k->host_len[slot0 + j];
      if (payload_len[first + j] > cap) return fail(ctx, VC2_ERR_CAPACITY);
      e = cudaMemcpyAsync(payloads[first + j], vc2_codec_payload_dev(k, slot0 + j), payload_len[first + j], cudaMemcpyDeviceToHost, k->copy_out);
      if (e != cudaSuccess) return cuda_fail(ctx, e);
    }
    e = cudaEventRecord(k->ev_out[slot0], k->copy_out);
    return e == cudaSuccess ? VC2_OK : cuda_fail(ctx, e);
  };
  CU(cudaStreamSynchronize(ctx->stream));   // earlier work of the caller on the slots
  for (int st = 0; st < nstages && status == VC2_OK; ++st) {
    const int slot0 = (st % nstage_slots) * sub, first = st * sub, m = std::min(sub, n - first);
    if (st >= nstage_slots) CU(cudaStreamWaitEvent(k->copy_in, k->ev_done[slot0], 0));   // the samples in these slots have been consumed
    for (int j = 0; j < m; ++j)
      CU(cudaMemcpyAsync(vc2_codec_samples_dev(k, slot0 + j), pictures[first + j], k->pic_bytes, cudaMemcpyHostToDevice, k->copy_in));
    CU(cudaEventRecord(k->ev_in[slot0], k->copy_in));
    CU(cudaStreamWaitEvent(ctx->stream, k->ev_in[slot0], 0));
    if (st >= nstage_slots) CU(cudaStreamWaitEvent(ctx->stream, k->ev_out[slot0], 0));   // the payloads have left these slots
    status = codec_encode_range(k, slot0, m);
    if (status) break;
    CU(cudaMemcpy2DAsync(k->host_len + slot0, 4, vc2_codec_slice_offsets_dev(k, slot0) + ns, (size_t)(ns + 1) * 4, 4, m,
                         cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(k->host_flags + (size_t)slot0 * ns, k->err.as<uint32_t>() + (size_t)slot0 * ns, (size_t)ns * 4 * m,
                       cudaMemcpyDeviceToHost, ctx->stream));
    if (k->narrow_enc) CU(cudaMemcpyAsync(k->host_novf + slot0, k->narrow_ovf.as<uint32_t>() + slot0, (size_t)4 * m, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(k->ev_done[slot0], ctx->stream));
    if (st >= lag) status = drain(st - lag);
  }
  for (int st = std::max(0, nstages - lag); st < nstages && status == VC2_OK; ++st) status = drain(st);
  cudaStreamSynchronize(k->copy_out);
  cudaStreamSynchronize(ctx->stream);
  for (size_t r = 0; r < redo.size() && status == VC2_OK; ++r) {   // slot 0, one at a time, 32-bit coefficient block
    const int i = redo[r];
    CU(cudaMemcpyAsync(vc2_codec_samples_dev(k, 0), pictures[i], k->pic_bytes, cudaMemcpyHostToDevice, ctx->stream));
    status = codec_encode_range(k, 0, 1, true);
    if (status) break;
    status = vc2_codec_download_payload(k, 0, payloads[i], cap, &payload_len[i], nullptr, nullptr);
  }
  return status;
}

extern "C" int vc2_codec_decode_host(vc2_codec* k, int n, const uint8_t* const* payloads, const size_t* payload_len,
                                     void* const* pictures) {
  KARG(k && pictures && payloads && payload_len && n >= 1);
  vc2_ctx* ctx = k->ctx;
  CU(cudaSetDevice(ctx->device));
  k->main_dirty = true;
  const int B = k->prm.max_pictures, ns = k->nslices;
  const bool hq = k->prm.mode != VC2_LD;
  for (int i = 0; i < n; ++i) {
    if (payload_len[i] > k->payload_cap) return fail(ctx, VC2_ERR_CAPACITY);
    if (!hq && payload_len[i] < k->fixed_off[ns]) return fail(ctx, VC2_ERR_STREAM);
  }
  CU(cudaStreamSynchronize(ctx->stream));
  auto flags = [&](int set) { return k->host_flags + (size_t)set * B * ns; };
  std::vector<int> redo;             // pictures that overflowed the narrow coefficient block: decoded again at the end, 32-bit
  auto check = [&](int c) -> int {   // wait for chunk c and look at its error flags
    const int m = std::min(B, n - c * B);
    const int sub = stage_pictures(B);
    cudaError_t e = cudaEventSynchronize(k->ev_out[(c & 1) * B + (m - 1) / sub * sub]);
    if (e != cudaSuccess) return cuda_fail(ctx, e);
    for (int i = 0; i < m; ++i) {
      if (k->narrow_dec && k->host_novf[(c & 1) * B + i]) { redo.push_back(c * B + i); continue; }
      const int st = first_error(flags(c & 1) + (size_t)i * ns, ns, VC2_FLAG_VLC_RANGE);
      if (st) return fail(ctx, st);
    }
    return VC2_OK;
  };
  const int chunks = (n + B - 1) / B;
  int status = VC2_OK;
  for (int c = 0; c < chunks && status == VC2_OK; ++c) {
    const int base = c * B, m = std::min(B, n - base);
    const int sub = stage_pictures(B);
    for (int i = 0; i < m; i += sub) {   // slots i .. i+mm-1; every wait below is stream side, the host does not block
      const int mm = std::min(sub, m - i);
      if (c > 0) CU(cudaStreamWaitEvent(k->copy_in, k->ev_done[i], 0));      // the previous payloads in these slots have been parsed
      for (int j = i; j < i + mm; ++j) {
        CU(cudaMemcpyAsync(vc2_codec_payload_dev(k, j), payloads[base + j], payload_len[base + j], cudaMemcpyHostToDevice, k->copy_in));
        if (!hq) continue;
        // slice offsets of this payload: walked on the device as soon as its bytes have landed
        cudaStream_t is = k->index_stream[j % (int)k->index_stream.size()];
        CU(cudaEventRecord(k->ev_idx[j], k->copy_in));
        CU(cudaStreamWaitEvent(is, k->ev_idx[j], 0));
        IndexParams ip;
        memset(&ip, 0, sizeof(ip));
        ip.in = vc2_codec_payload_dev(k, j); ip.in_pic_stride = (long long)k->payload_cap;
        ip.len[0] = (uint32_t)payload_len[base + j];
        ip.slice_off = vc2_codec_slice_offsets_dev(k, j);
        ip.nslices = ns; ip.prefix = k->g.prefix; ip.scalar = k->g.scalar;
        CU(index_launch(is, ip, 1));
        ctx->launches++;
        CU(cudaEventRecord(k->ev_idx[j], is));
        CU(cudaStreamWaitEvent(ctx->stream, k->ev_idx[j], 0));
      }
      CU(cudaEventRecord(k->ev_in[i], k->copy_in));
      CU(cudaStreamWaitEvent(ctx->stream, k->ev_in[i], 0));
      if (c > 0) CU(cudaStreamWaitEvent(ctx->stream, k->ev_out[((c - 1) & 1) * B + i], 0));   // the previous pictures have left these slots
      status = codec_decode_range(k, i, mm, false);
      if (status) break;
      if (k->narrow_dec) CU(cudaMemcpyAsync(k->host_novf + (c & 1) * B + i, k->narrow_ovf.as<uint32_t>() + B + i, (size_t)4 * mm, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaEventRecord(k->ev_done[i], ctx->stream));
      CU(cudaStreamWaitEvent(k->copy_out, k->ev_done[i], 0));
      for (int j = i; j < i + mm; ++j)
        CU(cudaMemcpyAsync(pictures[base + j], vc2_codec_recon_dev(k, j), k->pic_bytes, cudaMemcpyDeviceToHost, k->copy_out));
      CU(cudaMemcpyAsync(flags(c & 1) + (size_t)i * ns, k->err.as<uint32_t>() + (size_t)i * ns, (size_t)ns * 4 * mm, cudaMemcpyDeviceToHost,
                         k->copy_out));
      CU(cudaEventRecord(k->ev_out[(c & 1) * B + i], k->copy_out));
    }
    if (status) break;
    if (c + 1 < chunks && c >= 1) status = check(c - 1);   // the flag set of chunk c+1 was last used by chunk c-1
  }
  if (status == VC2_OK && chunks >= 2) status = check(chunks - 2);
  if (status == VC2_OK) status = check(chunks - 1);
  cudaStreamSynchronize(k->copy_out);
  cudaStreamSynchronize(ctx->stream);
  for (size_t r = 0; r < redo.size() && status == VC2_OK; ++r) {   // slot 0, one at a time, 32-bit coefficient block
    const int i = redo[r];
    status = vc2_codec_upload_payload(k, 0, payloads[i], payload_len[i]);
    if (status) break;
    status = codec_decode_range(k, 0, 1, hq, true);
    if (status) break;
    status = codec_slot_status(k, 0, VC2_FLAG_VLC_RANGE);
    if (status) break;
    status = vc2_codec_download_picture(k, 0, pictures[i]);
  }
  return status;
}
