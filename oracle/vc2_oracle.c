/* vc2_oracle.c - plain-C CPU restatement of the bbc/vc2-reference HQ/LD hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
 * load this (oracle/_build/libvc2oracle.so); the product path (vc2_reference_b200/) never does.
 *
 * It restates, from the arithmetic spec extracted in SURVEY.md Appendix A, the algorithms of
 * (paths relative to /root/reference/src):
 *   Library/src/WaveletTransform.cpp:74-94, 116-136, 224-342, 345-450, 478-1265
 *   Library/src/Quantisation.cpp:16-20, 40-95, 287-306, 479-558, 627-642
 *   Library/src/VLC.cpp:21-94, 151-257
 *   Library/src/Slices.cpp:28-49, 97-119, 246-303, 305-382, 469-612
 *   EncodeStream/EncodeStream.cpp:73-125
 * Nothing here is copied from the reference: the lifting engine is table driven, the bit IO is a
 * flat MSB-first cursor.  PINNED: tests/test_oracle.py checks every function against the compiled,
 * unmodified reference (oracle/_ref/libvc2ref.so, built by oracle/build_ref.sh) on random inputs,
 * against the reference's own known answers (tests/Quantisation.cpp:30-36) and against the golden
 * digests produced by running the reference (tests/golden/md5.json).
 *
 * Conventions: planes are row-major int32 [y][x]; transformed planes are padded and hold the
 * reference's in-place interleaved coefficient order.  Every function returns 0 or a negative
 * ORC_ERR_* code standing for the reference exception named beside it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_ARG (-1)
#define ORC_ERR_SCALAR_TOO_SMALL (-3) /* Slices.cpp:115-117 */
#define ORC_ERR_QUANT_INDEX (-4)      /* Quantisation.cpp:60-63 */
#define ORC_ERR_CBR_TOO_MANY_BYTES (-5) /* Slices.cpp:356-358 */
#define ORC_ERR_CBR_COMP_LENGTH (-6)  /* Slices.cpp:359-366 */
#define ORC_ERR_CAPACITY (-7)
#define ORC_ERR_STREAM (-9)
#define ORC_ERR_LD_TOO_MANY_BYTES (-10) /* Slices.cpp:209-211 */

enum { K_DD97 = 0, K_LEGALL, K_DD137, K_HAAR0, K_HAAR1, K_FIDELITY, K_DAUB97 };
#define MAX_DEPTH 6
#define MAX_BANDS (3 * MAX_DEPTH + 1)

/* ------------------------------------------------------------------------------------------------
 * geometry helpers
 * ---------------------------------------------------------------------------------------------- */
int orc_padded_size(int size, int depth) { /* WaveletTransform.cpp:74-77 */
  const int unit = 1 << depth;
  return ((size + unit - 1) / unit) * unit;
}

int orc_slice_size_is_valid(int depth, int luma, int chroma, int n) { /* WaveletTransform.cpp:116-136 */
  const int unit = 1 << depth;
  const int most = (luma < chroma ? luma : chroma) / unit;
  if (n <= 0 || n > most) return 0;
  const int pl = orc_padded_size(luma, depth), pc = orc_padded_size(chroma, depth);
  const int span = n * unit;
  const int count = (pl + span - 1) / span;
  if (pl % count || pc % count) return 0;
  if ((pl / count) % unit || (pc / count) % unit) return 0;
  return count;
}

/* ------------------------------------------------------------------------------------------------
 * quantiser (Quantisation.cpp:40-95).  The factor table is generated from the ST 2042-1 closed form.
 * ---------------------------------------------------------------------------------------------- */
int orc_quant_factor(int q) {
  if (q < 0) q = 0;
  if (q > 119) return -1; /* the reference throws (Quantisation.cpp:60-63) */
  const uint64_t base = (uint64_t)1 << (q >> 2);
  uint64_t f;
  switch (q & 3) {
    case 0: f = 4 * base; break;
    case 1: f = (503829 * base + 52958) / 105917; break;
    case 2: f = (665857 * base + 58854) / 117708; break;
    default: f = (440253 * base + 32722) / 65444; break;
  }
  return (int)(uint32_t)f; /* the reference table is int: entries >= 2^31 wrap (SURVEY C-2) */
}
int orc_quant_offset(int q) {
  if (q < 0) q = 0;
  if (q > 119) return -1;
  if (q == 0) return 1;
  if (q == 1) return 2;
  return (orc_quant_factor(q) + 1) / 2;
}
int orc_quant(int v, int q, int* out) {
  const int f = orc_quant_factor(q);
  if (f < 0 && q > 119) return ORC_ERR_QUANT_INDEX;
  const int mag = (int)(((uint32_t)abs(v) << 2) / (uint32_t)f); /* truncating */
  *out = v < 0 ? -mag : mag;
  return ORC_OK;
}
int orc_scale(int v, int q, int* out) {
  const int f = orc_quant_factor(q), o = orc_quant_offset(q);
  if (q > 119) return ORC_ERR_QUANT_INDEX;
  if (v == 0) { *out = 0; return ORC_OK; }
  const int mag = (int)(((uint32_t)abs(v) * (uint32_t)f + (uint32_t)o + 2u) >> 2);
  *out = v < 0 ? -mag : mag;
  return ORC_OK;
}

int orc_quant_matrix(int kernel, int depth, int* out) { /* WaveletTransform.cpp:345-423 */
  /* low-pass / high-pass analysis gains of each kernel; float arithmetic in the reference's order */
  static const float A[7] = {1.280868846f, 1.224744871f, 1.280868846f, 1.414213562f, 1.414213562f, 0.682408629f, 1.139917028f};
  static const float B[7] = {0.820572875f, 0.847791248f, 0.809253958f, 0.707106871f, 0.707106871f, 1.367856979f, 0.887168005f};
  static const int SH[7] = {1, 1, 1, 0, 1, 0, 1};
  if (depth < 0 || depth > MAX_DEPTH || kernel < 0 || kernel > 6) return ORC_ERR_ARG;
  if (depth == 0) { out[0] = 0; return ORC_OK; }
  const float a2 = A[kernel] * A[kernel], ab = A[kernel] * B[kernel], b2 = B[kernel] * B[kernel];
  float gl[MAX_DEPTH + 1], gm[MAX_DEPTH + 1], gh[MAX_DEPTH + 1], lo = 3.402823466e+38f;
  for (int lev = depth; lev >= 1; --lev) {
    const float s = (float)(pow(a2, depth - lev) / pow(2.0f, SH[kernel] * (depth - lev + 1)));
    gl[lev] = s * a2; gm[lev] = s * ab; gh[lev] = s * b2;
    float m = gl[lev] < gm[lev] ? gl[lev] : gm[lev];
    m = m < gh[lev] ? m : gh[lev];
    lo = m < lo ? m : lo;
  }
#define QSTEP(g) ((int)floor(4.0f * log((g) / lo) / log(2.0f) + 0.5f))
  int n = 0;
  out[n++] = QSTEP(gl[1]);
  for (int lev = 1; lev <= depth; ++lev) {
    out[n++] = QSTEP(gm[lev]);
    out[n++] = QSTEP(gm[lev]);
    out[n++] = QSTEP(gh[lev]);
  }
#undef QSTEP
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * integer lifting (WaveletTransform.cpp:478-1265), table driven.
 * A step updates every sample of parity `par` from the 2*ntaps nearest samples of the other parity:
 *    x[i] += sign * ((add + sum_k c[k] * (x[i-(2k+1)] + x[i+(2k+1)])) >> sh)
 * except Haar whose steps are one sided.  A tap index outside [0,n) is replaced by the nearest
 * in-range index of the same parity (the reference's explicit edge statements).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int par;        /* parity of the updated samples: 1 = odd (predict), 0 = even (update) */
  int ntaps;      /* taps per side */
  int cl[4], cr[4]; /* coefficients of x[i-(2k+1)] and x[i+(2k+1)] */
  int add, sh, sign;
} lift_step;

typedef struct { int nsteps, shift; lift_step s[4]; } lift_kernel;

static const lift_kernel KERNELS[7] = {
  /* DD97 (:492-511) */
  {2, 1, {{1, 2, {9, -1}, {9, -1}, 8, 4, -1}, {0, 1, {1}, {1}, 2, 2, +1}}},
  /* LeGall (:609-625) */
  {2, 1, {{1, 1, {1}, {1}, 1, 1, -1}, {0, 1, {1}, {1}, 2, 2, +1}}},
  /* DD137 (:714-736) */
  {2, 1, {{1, 2, {9, -1}, {9, -1}, 8, 4, -1}, {0, 2, {9, -1}, {9, -1}, 16, 5, +1}}},
  /* Haar0 (:843-855): o -= e[x] ; e += (o[x+1] + 1) >> 1 */
  {2, 0, {{1, 1, {1}, {0}, 0, 0, -1}, {0, 1, {0}, {1}, 1, 1, +1}}},
  /* Haar1 */
  {2, 1, {{1, 1, {1}, {0}, 0, 0, -1}, {0, 1, {0}, {1}, 1, 1, +1}}},
  /* Fidelity (:933-964): update first */
  {2, 0, {{0, 4, {161, -46, 21, -8}, {161, -46, 21, -8}, 128, 8, +1}, {1, 4, {81, -25, 10, -2}, {81, -25, 10, -2}, 128, 8, -1}}},
  /* Daub97 (:1104-1137) */
  {4, 1, {{1, 1, {6497}, {6497}, 2048, 12, -1}, {0, 1, {217}, {217}, 2048, 12, -1},
          {1, 1, {3616}, {3616}, 2048, 12, +1}, {0, 1, {1817}, {1817}, 2048, 12, +1}}},
};

static int clamp_same_parity(int i, int n, int par) {
  /* n is even; indices of parity par run par, par+2, ..., n-2+par */
  if (i < 0) return par;
  if (i >= n) return n - 2 + par;
  return i;
}

/* one lifting step along a line of n samples with element stride `st`; dir = +1 forward, -1 inverse */
static void lift_line(int32_t* p, long st, int n, const lift_step* s, int dir) {
  const int src_par = 1 - s->par;
  for (int i = s->par; i < n; i += 2) {
    uint32_t acc = (uint32_t)s->add;
    for (int k = 0; k < s->ntaps; ++k) {
      const int a = clamp_same_parity(i - (2 * k + 1), n, src_par);
      const int b = clamp_same_parity(i + (2 * k + 1), n, src_par);
      acc += (uint32_t)s->cl[k] * (uint32_t)p[a * st] + (uint32_t)s->cr[k] * (uint32_t)p[b * st];
    }
    const int32_t d = ((int32_t)acc) >> s->sh; /* arithmetic shift, as the reference's >> on int */
    if (s->sign * dir > 0) p[i * st] = (int32_t)((uint32_t)p[i * st] + (uint32_t)d);
    else p[i * st] = (int32_t)((uint32_t)p[i * st] - (uint32_t)d);
  }
}

/* one level on the lattice of samples at multiples of `sub` of a ph x pw plane */
static void level_forward(int32_t* pl, int ph, int pw, int sub, const lift_kernel* K) {
  const int h = ph / sub, w = pw / sub;
  if (K->shift)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        int32_t* e = pl + (long)y * sub * pw + (long)x * sub;
        *e = (int32_t)((uint32_t)*e << K->shift);
      }
  for (int y = 0; y < h; ++y)
    for (int s = 0; s < K->nsteps; ++s) lift_line(pl + (long)y * sub * pw, sub, w, &K->s[s], +1);
  for (int x = 0; x < w; ++x)
    for (int s = 0; s < K->nsteps; ++s) lift_line(pl + (long)x * sub, (long)sub * pw, h, &K->s[s], +1);
}
static void level_inverse(int32_t* pl, int ph, int pw, int sub, const lift_kernel* K) {
  const int h = ph / sub, w = pw / sub;
  for (int x = 0; x < w; ++x)
    for (int s = K->nsteps - 1; s >= 0; --s) lift_line(pl + (long)x * sub, (long)sub * pw, h, &K->s[s], -1);
  for (int y = 0; y < h; ++y)
    for (int s = K->nsteps - 1; s >= 0; --s) lift_line(pl + (long)y * sub * pw, sub, w, &K->s[s], -1);
  if (K->shift) {
    const int r = 1 << (K->shift - 1);
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        int32_t* e = pl + (long)y * sub * pw + (long)x * sub;
        *e = (*e + r) >> K->shift;
      }
  }
}

/* waveletTransform (WaveletTransform.cpp:262-281) incl. waveletPad (:79-94); dst is ph x pw */
int orc_dwt_forward(const int32_t* src, int h, int w, int kernel, int depth, int32_t* dst) {
  if (kernel < 0 || kernel > 6 || depth < 0 || depth > MAX_DEPTH || h < 1 || w < 1) return ORC_ERR_ARG;
  const int ph = orc_padded_size(h, depth), pw = orc_padded_size(w, depth);
  for (int y = 0; y < ph; ++y) {
    const int32_t* row = src + (long)(y < h ? y : h - 1) * w;
    for (int x = 0; x < pw; ++x) dst[(long)y * pw + x] = row[x < w ? x : w - 1];
  }
  for (int lev = 0; lev < depth; ++lev) level_forward(dst, ph, pw, 1 << lev, &KERNELS[kernel]);
  return ORC_OK;
}
/* inverseWaveletTransform (:321-342): levels coarse to fine, then crop to h x w (:340) */
int orc_dwt_inverse(const int32_t* src, int ph, int pw, int kernel, int depth, int32_t* dst, int h, int w) {
  if (kernel < 0 || kernel > 6 || depth < 0 || depth > MAX_DEPTH || h > ph || w > pw) return ORC_ERR_ARG;
  int32_t* t = (int32_t*)malloc(sizeof(int32_t) * (size_t)ph * pw);
  if (!t) return ORC_ERR_ARG;
  memcpy(t, src, sizeof(int32_t) * (size_t)ph * pw);
  for (int lev = depth - 1; lev >= 0; --lev) level_inverse(t, ph, pw, 1 << lev, &KERNELS[kernel]);
  for (int y = 0; y < h; ++y) memcpy(dst + (long)y * w, t + (long)y * pw, sizeof(int32_t) * (size_t)w);
  free(t);
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * in-place layout: which subband a coefficient belongs to (WaveletTransform.cpp:428-450)
 * ---------------------------------------------------------------------------------------------- */
static int band_of(int y, int x, int depth) {
  const int unit = 1 << depth;
  const int t = (y | x) & (unit - 1);
  if (t == 0) return 0;
  int lev = 0; /* 0 = finest */
  while (!((t >> lev) & 1)) ++lev;
  const int L = depth - lev; /* VC-2 level, 1 = coarsest */
  const int hx = (x >> lev) & 1, hy = (y >> lev) & 1;
  return 3 * (L - 1) + (hx ? (hy ? 3 : 1) : 2);
}

/* quantise_transform_np / inverse_quantise_transform_np (Quantisation.cpp:479-489, 534-544) */
static int quant_plane(const int32_t* c, int ph, int pw, const int32_t* qidx, int ny, int nx, const int32_t* qm, int nbands,
                       int32_t* out, int inverse, int skip_ll) {
  const int depth = (nbands - 1) / 3;
  if (ph % ny || pw % nx) return ORC_ERR_ARG;
  const int sh = ph / ny, sw = pw / nx;
  for (int y = 0; y < ph; ++y)
    for (int x = 0; x < pw; ++x) {
      const int b = band_of(y, x, depth);
      const long i = (long)y * pw + x;
      if (b == 0 && skip_ll) { out[i] = c[i]; continue; }
      int q = qidx[(y / sh) * nx + x / sw] - qm[b];
      if (q < 0) q = 0;
      const int rc = inverse ? orc_scale(c[i], q, &out[i]) : orc_quant(c[i], q, &out[i]);
      if (rc) return rc;
    }
  return ORC_OK;
}
int orc_quantise_np(const int32_t* c, int ph, int pw, const int32_t* qidx, int ny, int nx, const int32_t* qm, int nbands, int32_t* out) {
  return quant_plane(c, ph, pw, qidx, ny, nx, qm, nbands, out, 0, 0);
}
int orc_dequantise_np(const int32_t* c, int ph, int pw, const int32_t* qidx, int ny, int nx, const int32_t* qm, int nbands, int32_t* out) {
  return quant_plane(c, ph, pw, qidx, ny, nx, qm, nbands, out, 1, 0);
}
/* LD inverse quantisation: LL band DC predicted in raster order (Quantisation.cpp:191-208, 287-306, 369-379) */
int orc_dequantise_ld(const int32_t* c, int ph, int pw, const int32_t* qidx, int ny, int nx, const int32_t* qm, int nbands, int32_t* out) {
  const int depth = (nbands - 1) / 3;
  int rc = quant_plane(c, ph, pw, qidx, ny, nx, qm, nbands, out, 1, 1);
  if (rc) return rc;
  const int H = ph >> depth, W = pw >> depth;
  const long sy = (long)pw << depth, sx = 1L << depth;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int yb = ((y + 1) * ny - 1) / H, xb = ((x + 1) * nx - 1) / W;
      int q = qidx[yb * nx + xb] - qm[0];
      if (q < 0) q = 0;
      int pred;
      if (y > 0 && x > 0) {
        const int s = out[(y - 1) * sy + (x - 1) * sx] + out[(y - 1) * sy + x * sx] + out[y * sy + (x - 1) * sx];
        pred = s >= 0 ? (s + 1) / 3 : (s - 1) / 3;
      } else if (y > 0) pred = out[(y - 1) * sy + x * sx];
      else if (x > 0) pred = out[y * sy + (x - 1) * sx];
      else pred = 0;
      int v;
      rc = orc_scale(c[y * sy + x * sx], q, &v);
      if (rc) return rc;
      out[y * sy + x * sx] = v + pred;
    }
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * slice geometry and coefficient scan order
 * ---------------------------------------------------------------------------------------------- */
int orc_slice_bytes(int ny, int nx, int total, int scalar, int32_t* out) { /* Slices.cpp:28-49 */
  const int n = ny * nx;
  long num = total / scalar - 4L * n, den = n;
  long a = num < 0 ? -num : num, b = den;
  while (b) { const long t = a % b; a = b; b = t; }
  if (a) { num /= a; den /= a; }
  const long whole = num / den, rest = num - whole * den;
  long carry = 0;
  for (int i = 0; i < n; ++i) {
    carry += rest;
    if (carry >= den) { out[i] = (int)((whole + 1) * scalar + 4); carry -= den; }
    else out[i] = (int)(whole * scalar + 4);
  }
  return ORC_OK;
}

/* visit the coefficients of slice (sy,sx) of a ph x pw in-place plane in coding order:
 * band 0..3*depth, raster inside the slice's part of each band; idx[] receives plane offsets */
static int slice_scan(int ph, int pw, int depth, int ny, int nx, int sy, int sx, long* idx) {
  const int sh = ph / ny, sw = pw / nx, y0 = sy * sh, x0 = sx * sw;
  int n = 0;
  for (int b = 0; b <= 3 * depth; ++b) {
    int stride, oy = 0, ox = 0;
    if (b == 0) stride = 1 << depth;
    else {
      const int L = (b - 1) / 3 + 1, kind = (b - 1) % 3;
      stride = 1 << (depth + 1 - L);
      if (kind == 0 || kind == 2) ox = stride / 2;
      if (kind == 1 || kind == 2) oy = stride / 2;
    }
    for (int y = oy; y < sh; y += stride)
      for (int x = ox; x < sw; x += stride) idx[n++] = (long)(y0 + y) * pw + (x0 + x);
  }
  return n;
}

/* SignedVLC (VLC.cpp:21-52, 78-94): number of bits and right-justified code */
int orc_signed_vlc(int v, unsigned* nbits, unsigned* code) {
  if (v == 0) { *nbits = 1; *code = 1; return ORC_OK; }
  const uint32_t m = (uint32_t)abs(v) + 1u;
  int k = 0;
  while ((m >> (k + 1)) != 0) ++k;
  uint32_t c = 0;
  for (int i = k - 1; i >= 0; --i) c = (c << 2) | ((m >> i) & 1u); /* 0 b(i) pairs */
  c = (c << 2) | 2u | (v < 0 ? 1u : 0u);                         /* stop bit, sign */
  *nbits = (unsigned)(2 * k + 2);
  *code = c;
  return ORC_OK;
}

static int comp_bits_to_last_nonzero(const int32_t* pl, const long* idx, int n) {
  int gross = 0, count = 0;
  for (int i = 0; i < n; ++i) {
    unsigned nb, code;
    orc_signed_vlc(pl[idx[i]], &nb, &code);
    gross += (int)nb;
    if (nb > 1) count = gross;
  }
  return count;
}
static int scaled_length(int bits, int scalar, int* too_big) { /* Slices.cpp:114-118 */
  const int units = ((bits + 7) / 8 + scalar - 1) / scalar;
  if (units > 0xFF) *too_big = 1;
  return units * scalar;
}
/* component_slice_bytes on one stand-alone slice array (Slices.cpp:97-119) */
int orc_component_slice_bytes(const int32_t* slice, int h, int w, int depth, int scalar, int* out) {
  long* idx = (long*)malloc(sizeof(long) * (size_t)h * w);
  const int n = slice_scan(h, w, depth, 1, 1, 0, 0, idx);
  int big = 0;
  *out = scaled_length(comp_bits_to_last_nonzero(slice, idx, n), scalar, &big);
  free(idx);
  return big ? ORC_ERR_SCALAR_TOO_SMALL : ORC_OK;
}

/* ---- MSB-first bit cursor over a byte buffer ----------------------------------------------------- */
typedef struct { uint8_t* p; long bit, end; } bitw; /* writes beyond `end` drop (only '1's can be cut, VLC.cpp:151-155) */
static void put_bits(bitw* w, uint32_t code, int n) {
  for (int i = n - 1; i >= 0; --i, ++w->bit) {
    if (w->bit >= w->end) continue;
    if ((code >> i) & 1u) w->p[w->bit >> 3] |= (uint8_t)(0x80u >> (w->bit & 7));
  }
}
typedef struct { const uint8_t* p; long bit, end; } bitr; /* reads beyond `end` return 1 (VLC.cpp:182-185) */
static int get_bit(bitr* r) {
  int b = 1;
  if (r->bit < r->end) b = (r->p[r->bit >> 3] >> (7 - (r->bit & 7))) & 1;
  ++r->bit;
  return b;
}
static int get_signed_vlc(bitr* r) { /* VLC.cpp:283-317 */
  uint32_t m = 1;
  while (!get_bit(r)) m = (m << 1) | (uint32_t)get_bit(r);
  const int v = (int)(m - 1);
  if (v == 0) return 0;
  return get_bit(r) ? -v : v;
}

/* ---- HQ slice writer / reader --------------------------------------------------------------------- */
typedef struct { int ph, pw; const int32_t* pl; } plane_in;

/* operator<<(ostream&, Slices) with HQSliceIO_VBR (mode 0, Slices.cpp:469-533) or HQSliceIO_CBR
 * (mode 1, :305-382); planes hold QUANTISED coefficients */
int orc_pack_slices(const int32_t* y, const int32_t* u, const int32_t* v, int lh, int lw, int ch, int cw, int depth,
                    const int32_t* qidx, int ny, int nx, int mode, int prefix, int scalar, const int32_t* sbytes,
                    uint8_t* out, long cap, long* out_len) {
  const plane_in P[3] = {{lh, lw, y}, {ch, cw, u}, {ch, cw, v}};
  long* idx = (long*)malloc(sizeof(long) * (size_t)(lh / ny) * (lw / nx));
  long pos = 0;
  int rc = ORC_OK;
  for (int sy = 0; sy < ny && !rc; ++sy)
    for (int sx = 0; sx < nx && !rc; ++sx) {
      int len[3], n[3], big = 0;
      for (int c = 0; c < 3; ++c) {
        n[c] = slice_scan(P[c].ph, P[c].pw, depth, ny, nx, sy, sx, idx);
        len[c] = scaled_length(comp_bits_to_last_nonzero(P[c].pl, idx, n[c]), scalar, &big);
      }
      if (big) { rc = ORC_ERR_SCALAR_TOO_SMALL; break; }
      if (mode == 1) {
        const int vb = sbytes[sy * nx + sx] - 4 - len[0] - len[1];
        if (vb < len[2]) { rc = ORC_ERR_CBR_TOO_MANY_BYTES; break; }
        if (vb / scalar > 255) { rc = ORC_ERR_CBR_COMP_LENGTH; break; }
        len[2] = vb;
      }
      const long total = prefix + 4 + len[0] + len[1] + len[2];
      if (pos + total > cap) { rc = ORC_ERR_CAPACITY; break; }
      memset(out + pos, 0, (size_t)total);
      long at = pos + prefix;
      out[at++] = (uint8_t)qidx[sy * nx + sx];
      for (int c = 0; c < 3; ++c) {
        out[at++] = (uint8_t)(len[c] / scalar);
        slice_scan(P[c].ph, P[c].pw, depth, ny, nx, sy, sx, idx);
        bitw w = {out + at, 0, 8L * len[c]};
        for (int i = 0; i < n[c]; ++i) {
          unsigned nb, code;
          orc_signed_vlc(P[c].pl[idx[i]], &nb, &code);
          put_bits(&w, code, (int)nb);
        }
        at += len[c];
      }
      pos += total;
    }
  free(idx);
  *out_len = pos;
  return rc;
}

/* operator>>(istream&, Slices) with HQSliceIO_VBR (mode 0, Slices.cpp:535-612) or LDSliceIO (mode 2, :246-303) */
int orc_unpack_slices(const uint8_t* in, long len, int lh, int lw, int ch, int cw, int depth, int ny, int nx, int mode,
                      int prefix, int scalar, const int32_t* sbytes, int32_t* y, int32_t* u, int32_t* v, int32_t* qidx) {
  int32_t* const out[3] = {y, u, v};
  const int PH[3] = {lh, ch, ch}, PW[3] = {lw, cw, cw};
  long* idx = (long*)malloc(sizeof(long) * (size_t)(lh / ny) * (lw / nx));
  long* idx2 = (long*)malloc(sizeof(long) * (size_t)(lh / ny) * (lw / nx));
  long pos = 0;
  int rc = ORC_OK;
  for (int sy = 0; sy < ny && !rc; ++sy)
    for (int sx = 0; sx < nx && !rc; ++sx) {
      if (mode == 0) {
        if (pos + prefix + 1 > len) { rc = ORC_ERR_STREAM; break; }
        pos += prefix;
        qidx[sy * nx + sx] = in[pos++];
        for (int c = 0; c < 3; ++c) {
          if (pos + 1 > len) { rc = ORC_ERR_STREAM; break; }
          const long L = (long)in[pos++] * scalar;
          if (pos + L > len) { rc = ORC_ERR_STREAM; break; }
          const int n = slice_scan(PH[c], PW[c], depth, ny, nx, sy, sx, idx);
          bitr r = {in + pos, 0, 8 * L};
          for (int i = 0; i < n; ++i) out[c][idx[i]] = get_signed_vlc(&r);
          pos += L;
        }
      } else {
        const long B = sbytes[sy * nx + sx];
        if (pos + B > len) { rc = ORC_ERR_STREAM; break; }
        bitr r = {in + pos, 0, 8 * B};
        int q = 0;
        for (int i = 0; i < 7; ++i) q = (q << 1) | get_bit(&r);
        qidx[sy * nx + sx] = q;
        int lb = 0; /* utils::intlog2(8B-7): bits needed to express the luma length */
        while ((1L << lb) < 8 * B - 7) ++lb;
        long ybits = 0;
        for (int i = 0; i < lb; ++i) ybits = (ybits << 1) | get_bit(&r);
        const long ystart = r.bit;
        int n = slice_scan(PH[0], PW[0], depth, ny, nx, sy, sx, idx);
        bitr ry = {in + pos, ystart, ystart + ybits};
        for (int i = 0; i < n; ++i) out[0][idx[i]] = get_signed_vlc(&ry);
        n = slice_scan(PH[1], PW[1], depth, ny, nx, sy, sx, idx2);
        bitr rc2 = {in + pos, ystart + ybits, 8 * B};
        for (int i = 0; i < n; ++i) {
          out[1][idx2[i]] = get_signed_vlc(&rc2);
          out[2][idx2[i]] = get_signed_vlc(&rc2);
        }
        pos += B;
      }
    }
  free(idx);
  free(idx2);
  return rc;
}

/* ------------------------------------------------------------------------------------------------
 * HQ_CBR rate control: quantIndicesCBR (EncodeStream.cpp:73-125), yss_for_slice (Quantisation.cpp:627-642)
 * ---------------------------------------------------------------------------------------------- */
static int quantise_slice(const int32_t* pl, const long* idx, int n, int pw, int depth, int q, const int32_t* qm, int32_t* dst) {
  for (int i = 0; i < n; ++i) {
    const int yy = (int)(idx[i] / pw), xx = (int)(idx[i] % pw);
    int aq = q - qm[band_of(yy, xx, depth)];
    if (aq < 0) aq = 0;
    const int rc = orc_quant(pl[idx[i]], aq, &dst[i]);
    if (rc) return rc;
  }
  return ORC_OK;
}
static int bits_of_list(const int32_t* v, int n) {
  int gross = 0, count = 0;
  for (int i = 0; i < n; ++i) {
    unsigned nb, code;
    orc_signed_vlc(v[i], &nb, &code);
    gross += (int)nb;
    if (nb > 1) count = gross;
  }
  return count;
}
static int luma_sq_error(const int32_t* pl, const long* idx, int n, int pw, int depth, int q, const int32_t* qm, long long* out) {
  long long acc = 0;
  for (int i = 0; i < n; ++i) {
    const int yy = (int)(idx[i] / pw), xx = (int)(idx[i] % pw);
    int aq = q - qm[band_of(yy, xx, depth)];
    if (aq < 0) aq = 0;
    int qv, rv;
    int rc = orc_quant(pl[idx[i]], aq, &qv);
    if (!rc) rc = orc_scale(qv, aq, &rv);
    if (rc) return rc;
    const int d = pl[idx[i]] - rv;
    acc += (long long)(int32_t)((uint32_t)d * (uint32_t)d); /* product in int, sum in long long */
  }
  *out = acc;
  return ORC_OK;
}

int orc_cbr_qindices(const int32_t* y, const int32_t* u, const int32_t* v, int lh, int lw, int ch, int cw, const int32_t* qm,
                     int nbands, const int32_t* sbytes, int ny, int nx, int scalar, int32_t* out) {
  const int depth = (nbands - 1) / 3;
  const plane_in P[3] = {{lh, lw, y}, {ch, cw, u}, {ch, cw, v}};
  const size_t cap = (size_t)(lh / ny) * (lw / nx);
  long* idx = (long*)malloc(sizeof(long) * cap);
  int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * cap);
  int rc = ORC_OK;
  for (int sy = 0; sy < ny && !rc; ++sy)
    for (int sx = 0; sx < nx && !rc; ++sx) {
      const int avail = sbytes[sy * nx + sx] - 4;
      int trial = 63, best = 127, delta = 64;
      while (delta > 0 && !rc) {
        delta >>= 1;
        int need = 0, big = 0;
        for (int c = 0; c < 3 && !rc; ++c) {
          const int n = slice_scan(P[c].ph, P[c].pw, depth, ny, nx, sy, sx, idx);
          rc = quantise_slice(P[c].pl, idx, n, P[c].pw, depth, trial, qm, tmp);
          need += scaled_length(bits_of_list(tmp, n), scalar, &big);
        }
        if (!rc && big) rc = ORC_ERR_SCALAR_TOO_SMALL;
        if (rc) break;
        if (need <= avail) { if (trial < best) best = trial; trial -= delta; }
        else trial += delta;
      }
      if (rc) break;
      /* keep raising the index while the luma squared error strictly falls */
      const int n = slice_scan(lh, lw, depth, ny, nx, sy, sx, idx);
      trial = best;
      long long prev, cur;
      rc = luma_sq_error(y, idx, n, lw, depth, trial, qm, &prev);
      while (!rc) {
        ++trial;
        rc = luma_sq_error(y, idx, n, lw, depth, trial, qm, &cur);
        if (rc) break;
        const long long d = cur - prev;
        prev = cur;
        if (!(d < 0)) break;
      }
      out[sy * nx + sx] = trial - 1;
    }
  free(idx);
  free(tmp);
  return rc;
}

/* ------------------------------------------------------------------------------------------------
 * LD encoder: predictive quantiser (Quantisation.cpp:213-282, 357-367), quantIndicesLD with SliceQuantiserRef
 * (EncodeStream.cpp:139-245), LD slice writer (Slices.cpp:51-96, 195-244)
 * ---------------------------------------------------------------------------------------------- */
static int predict_dc(const int32_t* ll, int w, int y, int x) { /* predictDC, Quantisation.cpp:191-208 */
  if (y > 0 && x > 0) {
    const int s = ll[(y - 1) * w + x - 1] + ll[(y - 1) * w + x] + ll[y * w + x - 1];
    return s >= 0 ? (s + 1) / 3 : (s - 1) / 3;
  }
  if (y > 0) return ll[(y - 1) * w + x];
  if (x > 0) return ll[y * w + x - 1];
  return 0;
}

/* quantise_transform: every band with quant(), the LL band against the prediction from the restored LL band */
int orc_quantise_ld(const int32_t* c, int ph, int pw, const int32_t* qidx, int ny, int nx, const int32_t* qm, int nbands, int32_t* out) {
  const int depth = (nbands - 1) / 3;
  int rc = quant_plane(c, ph, pw, qidx, ny, nx, qm, nbands, out, 0, 0);
  if (rc) return rc;
  const int H = ph >> depth, W = pw >> depth;
  const long sy = (long)pw << depth, sx = 1L << depth;
  int32_t* restored = (int32_t*)calloc((size_t)H * W, sizeof(int32_t));
  for (int y = 0; y < H && !rc; ++y)
    for (int x = 0; x < W; ++x) {
      const int yb = ((y + 1) * ny - 1) / H, xb = ((x + 1) * nx - 1) / W;
      int q = qidx[yb * nx + xb] - qm[0];
      if (q < 0) q = 0;
      const int pred = predict_dc(restored, W, y, x);
      int qv, rv;
      rc = orc_quant(c[y * sy + x * sx] - pred, q, &qv);
      if (!rc) rc = orc_scale(qv, q, &rv);
      if (rc) break;
      out[y * sy + x * sx] = qv;
      restored[y * W + x] = rv + pred;
    }
  free(restored);
  return rc;
}

/* one component of one slice at trial index q, SliceQuantiserRef::quantise_slice (EncodeStream.cpp:172-191): the LL
 * samples (first n_ll entries of the scan list, raster inside the slice) are predicted from `restored` and update it */
static int ld_quantise_slice(const int32_t* pl, const long* idx, int n, int n_ll, int pw, int depth, int q, const int32_t* qm,
                             int32_t* restored, int llw, int32_t* dst) {
  for (int i = 0; i < n; ++i) {
    const int yy = (int)(idx[i] / pw), xx = (int)(idx[i] % pw);
    int aq = q - qm[band_of(yy, xx, depth)];
    if (aq < 0) aq = 0;
    int rc;
    if (i < n_ll) {
      const int yl = yy >> depth, xl = xx >> depth;
      const int pred = predict_dc(restored, llw, yl, xl);
      int rv;
      rc = orc_quant(pl[idx[i]] - pred, aq, &dst[i]);
      if (!rc) rc = orc_scale(dst[i], aq, &rv);
      if (rc) return rc;
      restored[yl * llw + xl] = rv + pred;
    } else {
      rc = orc_quant(pl[idx[i]], aq, &dst[i]);
      if (rc) return rc;
    }
  }
  return ORC_OK;
}
static int ld_intlog2(int v) { /* Utils.cpp:40-48 */
  int lg = 0;
  --v;
  while (v > 0) { v >>= 1; ++lg; }
  return lg;
}
/* chroma_slice_bits (Slices.cpp:70-96): U and V coefficient by coefficient */
static int bits_of_pair_list(const int32_t* a, const int32_t* b, int n) {
  int gross = 0, count = 0;
  for (int i = 0; i < n; ++i) {
    unsigned nb, code;
    orc_signed_vlc(a[i], &nb, &code);
    gross += (int)nb;
    if (nb > 1) count = gross;
    orc_signed_vlc(b[i], &nb, &code);
    gross += (int)nb;
    if (nb > 1) count = gross;
  }
  return count;
}

int orc_ld_qindices(const int32_t* y, const int32_t* u, const int32_t* v, int lh, int lw, int ch, int cw, const int32_t* qm,
                    int nbands, const int32_t* sbytes, int ny, int nx, int32_t* out) {
  const int depth = (nbands - 1) / 3;
  const plane_in P[3] = {{lh, lw, y}, {ch, cw, u}, {ch, cw, v}};
  const size_t cap = (size_t)(lh / ny) * (lw / nx);
  long* idx = (long*)malloc(sizeof(long) * cap);
  int32_t* tmp[3];
  int32_t* restored[3];
  for (int c = 0; c < 3; ++c) {
    tmp[c] = (int32_t*)malloc(sizeof(int32_t) * cap);
    restored[c] = (int32_t*)calloc((size_t)(P[c].ph >> depth) * (P[c].pw >> depth), sizeof(int32_t));
  }
  int rc = ORC_OK;
  for (int sy = 0; sy < ny && !rc; ++sy)
    for (int sx = 0; sx < nx && !rc; ++sx) {
      const int bytes = sbytes[sy * nx + sx];
      const int avail = 8 * bytes - 7 - ld_intlog2(8 * bytes - 7);
      int trial = 63, best = 127, delta = 64, n[3];
      for (int pass = 0; pass < 8 && !rc; ++pass) {   /* seven probes, then the slice again with the chosen index (:232-236) */
        const int q = pass < 7 ? trial : best;
        for (int c = 0; c < 3 && !rc; ++c) {
          n[c] = slice_scan(P[c].ph, P[c].pw, depth, ny, nx, sy, sx, idx);
          const int n_ll = ((P[c].ph / ny) >> depth) * ((P[c].pw / nx) >> depth);
          rc = ld_quantise_slice(P[c].pl, idx, n[c], n_ll, P[c].pw, depth, q, qm, restored[c], P[c].pw >> depth, tmp[c]);
        }
        if (rc || pass == 7) break;
        delta >>= 1;
        const int need = bits_of_list(tmp[0], n[0]) + bits_of_pair_list(tmp[1], tmp[2], n[1]);
        if (need <= avail) { if (trial < best) best = trial; trial -= delta; }
        else trial += delta;
      }
      out[sy * nx + sx] = best;
    }
  free(idx);
  for (int c = 0; c < 3; ++c) { free(tmp[c]); free(restored[c]); }
  return rc;
}

/* operator<<(ostream&, Slices) with LDSliceIO (Slices.cpp:195-244); planes hold QUANTISED coefficients */
int orc_pack_slices_ld(const int32_t* y, const int32_t* u, const int32_t* v, int lh, int lw, int ch, int cw, int depth,
                       const int32_t* qidx, int ny, int nx, const int32_t* sbytes, uint8_t* out, long cap, long* out_len) {
  const plane_in P[3] = {{lh, lw, y}, {ch, cw, u}, {ch, cw, v}};
  const size_t lcap = (size_t)(lh / ny) * (lw / nx);
  long* idx = (long*)malloc(sizeof(long) * lcap);
  int32_t* val[3];
  for (int c = 0; c < 3; ++c) val[c] = (int32_t*)malloc(sizeof(int32_t) * lcap);
  long pos = 0;
  int rc = ORC_OK;
  for (int sy = 0; sy < ny && !rc; ++sy)
    for (int sx = 0; sx < nx && !rc; ++sx) {
      const int size = sbytes[sy * nx + sx];
      int n[3];
      for (int c = 0; c < 3; ++c) {
        n[c] = slice_scan(P[c].ph, P[c].pw, depth, ny, nx, sy, sx, idx);
        for (int i = 0; i < n[c]; ++i) val[c][i] = P[c].pl[idx[i]];
      }
      const int ybits = bits_of_list(val[0], n[0]);
      const int split = ld_intlog2(8 * size - 7);
      const int uvbits = 8 * size - 7 - split - ybits;
      if (uvbits < bits_of_pair_list(val[1], val[2], n[1])) { rc = ORC_ERR_LD_TOO_MANY_BYTES; break; }
      if (pos + size > cap) { rc = ORC_ERR_CAPACITY; break; }
      memset(out + pos, 0, (size_t)size);
      bitw w = {out + pos, 0, 8L * size};
      put_bits(&w, (uint32_t)qidx[sy * nx + sx] & 0x7Fu, 7);
      put_bits(&w, (uint32_t)ybits, split);
      /* vlc::bounded(ybits): what lies beyond the bound (only the ones of trailing zeros) is dropped */
      w.end = 7 + split + ybits;
      for (int i = 0; i < n[0]; ++i) {
        unsigned nb, code;
        orc_signed_vlc(val[0][i], &nb, &code);
        put_bits(&w, code, (int)nb);
      }
      w.bit = w.end;                         /* vlc::flush */
      w.end = 7 + split + ybits + uvbits;    /* vlc::bounded(uvbits) */
      for (int i = 0; i < n[1]; ++i) {
        unsigned nb, code;
        orc_signed_vlc(val[1][i], &nb, &code);
        put_bits(&w, code, (int)nb);
        orc_signed_vlc(val[2][i], &nb, &code);
        put_bits(&w, code, (int)nb);
      }
      pos += size;
    }
  free(idx);
  for (int c = 0; c < 3; ++c) free(val[c]);
  *out_len = pos;
  return rc;
}

/* ------------------------------------------------------------------------------------------------
 * whole-picture convenience used by the CPU baseline "port" leg and by smoke():
 *   raw planar big-endian samples -> payload (HQ ConstQ), and back (Arrays.cpp:333-426, Picture.cpp:284-292)
 * ---------------------------------------------------------------------------------------------- */
static void read_plane_u16be(const uint8_t* raw, int n, int depth, int32_t* out) {
  for (int i = 0; i < n; ++i) out[i] = (int32_t)((((uint32_t)raw[2 * i] << 8) | raw[2 * i + 1]) >> (16 - depth)) - (1 << (depth - 1));
}
static void write_plane_u16be(const int32_t* v, int n, int depth, uint8_t* raw) {
  const int lo = -(1 << (depth - 1)), hi = (1 << (depth - 1)) - 1;
  for (int i = 0; i < n; ++i) {
    int s = v[i] < lo ? lo : (v[i] > hi ? hi : v[i]);
    const uint32_t w = (uint32_t)(s + (1 << (depth - 1))) << (16 - depth);
    raw[2 * i] = (uint8_t)(w >> 8);
    raw[2 * i + 1] = (uint8_t)w;
  }
}

int orc_encode_picture_hq_constq(const uint8_t* raw, int lh, int lw, int ch, int cw, int bits, int kernel, int depth, int ny, int nx,
                                 int q, int prefix, int scalar, uint8_t* out, long cap, long* out_len) {
  const int H[3] = {lh, ch, ch}, W[3] = {lw, cw, cw};
  int32_t qm[MAX_BANDS];
  orc_quant_matrix(kernel, depth, qm);
  int32_t* qplane[3] = {0, 0, 0};
  int32_t* qi = (int32_t*)malloc(sizeof(int32_t) * (size_t)ny * nx);
  for (int i = 0; i < ny * nx; ++i) qi[i] = q;
  int rc = ORC_OK;
  int PH[3], PW[3];
  for (int c = 0; c < 3 && !rc; ++c) {
    PH[c] = orc_padded_size(H[c], depth); PW[c] = orc_padded_size(W[c], depth);
    int32_t* s = (int32_t*)malloc(sizeof(int32_t) * (size_t)H[c] * W[c]);
    int32_t* t = (int32_t*)malloc(sizeof(int32_t) * (size_t)PH[c] * PW[c]);
    qplane[c] = (int32_t*)malloc(sizeof(int32_t) * (size_t)PH[c] * PW[c]);
    read_plane_u16be(raw, H[c] * W[c], bits, s);
    raw += 2L * H[c] * W[c];
    rc = orc_dwt_forward(s, H[c], W[c], kernel, depth, t);
    if (!rc) rc = orc_quantise_np(t, PH[c], PW[c], qi, ny, nx, qm, 3 * depth + 1, qplane[c]);
    free(s); free(t);
  }
  if (!rc) rc = orc_pack_slices(qplane[0], qplane[1], qplane[2], PH[0], PW[0], PH[1], PW[1], depth, qi, ny, nx, 0, prefix, scalar, 0, out, cap, out_len);
  for (int c = 0; c < 3; ++c) free(qplane[c]);
  free(qi);
  return rc;
}

int orc_decode_picture_hq(const uint8_t* in, long len, int lh, int lw, int ch, int cw, int bits, int kernel, int depth, int ny, int nx,
                          int prefix, int scalar, uint8_t* raw) {
  const int H[3] = {lh, ch, ch}, W[3] = {lw, cw, cw};
  int PH[3], PW[3];
  int32_t qm[MAX_BANDS];
  orc_quant_matrix(kernel, depth, qm);
  int32_t* qp[3];
  for (int c = 0; c < 3; ++c) {
    PH[c] = orc_padded_size(H[c], depth); PW[c] = orc_padded_size(W[c], depth);
    qp[c] = (int32_t*)malloc(sizeof(int32_t) * (size_t)PH[c] * PW[c]);
  }
  int32_t* qi = (int32_t*)malloc(sizeof(int32_t) * (size_t)ny * nx);
  int rc = orc_unpack_slices(in, len, PH[0], PW[0], PH[1], PW[1], depth, ny, nx, 0, prefix, scalar, 0, qp[0], qp[1], qp[2], qi);
  for (int c = 0; c < 3 && !rc; ++c) {
    int32_t* t = (int32_t*)malloc(sizeof(int32_t) * (size_t)PH[c] * PW[c]);
    int32_t* s = (int32_t*)malloc(sizeof(int32_t) * (size_t)H[c] * W[c]);
    rc = orc_dequantise_np(qp[c], PH[c], PW[c], qi, ny, nx, qm, 3 * depth + 1, t);
    if (!rc) rc = orc_dwt_inverse(t, PH[c], PW[c], kernel, depth, s, H[c], W[c]);
    if (!rc) write_plane_u16be(s, H[c] * W[c], bits, raw);
    raw += 2L * H[c] * W[c];
    free(t); free(s);
  }
  for (int c = 0; c < 3; ++c) free(qp[c]);
  free(qi);
  return rc;
}
