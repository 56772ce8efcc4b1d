// Slice coding kernels (see slices.cuh for the reference line citations).
#include "slices.cuh"

namespace vc2 {

__constant__ QuantTables c_qt;

cudaError_t upload_quant_tables(const QuantTables& t) { return cudaMemcpyToSymbol(c_qt, &t, sizeof(t)); }

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;

// ---- exact dead-zone quantiser (Quantisation.cpp:69-76): sign(v) * ((|v| << 2) / qf) --------------
// The truncating division by the table constant uses the round-up multiply-shift of Granlund &
// Montgomery, exact for every 32-bit dividend:  t = mulhi(m', a);  q = (t + ((a - t) >> 1)) >> (l - 1)
struct QParam {
  uint32_t qf, qo, qm, ql;
};
__device__ __forceinline__ QParam qparam(int q) {
  QParam r;
  q = min(max(q, 0), 127);
  r.qf = c_qt.qf[q];
  r.qo = c_qt.qo[q];
  r.qm = c_qt.qm[q];
  r.ql = c_qt.ql[q];
  return r;
}
__device__ __forceinline__ uint32_t udiv_magic(uint32_t a, uint32_t m, uint32_t l) {
  const uint32_t t = __umulhi(m, a);
  return (t + ((a - t) >> 1)) >> (l - 1);
}
__device__ __forceinline__ int quant_one(int v, uint32_t qm, uint32_t ql) {
  const uint32_t a = (uint32_t)abs(v) << 2;
  const int q = (int)udiv_magic(a, qm, ql);
  return v < 0 ? -q : q;
}
// inverse quantiser (Quantisation.cpp:86-95)
__device__ __forceinline__ int scale_one(int v, uint32_t qf, uint32_t qo) {
  if (v == 0) return 0;
  const uint32_t m = ((uint32_t)abs(v) * qf + qo + 2u) >> 2;
  return v < 0 ? -(int)m : (int)m;
}
// SignedVLC length (VLC.cpp:78-85): 1 for zero, else 2*floor(log2(|v|+1)) + 2
__device__ __forceinline__ int vlc_bits(int v) {
  if (v == 0) return 1;
  const int k = 31 - __clz(abs(v) + 1);
  return 2 * k + 2;
}
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
__device__ __forceinline__ uint32_t compress16(uint32_t x) {
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0F0F0F0Fu;
  x = (x | (x >> 4)) & 0x00FF00FFu;
  x = (x | (x >> 8)) & 0x0000FFFFu;
  return x;
}
// SignedVLC code word (VLC.cpp:21-52, 78-85), valid for |v| < 65535 (reference VLC is 32-bit)
__device__ __forceinline__ void vlc_code(int v, uint32_t& code, int& nb) {
  if (v == 0) { code = 1; nb = 1; return; }
  const uint32_t m = (uint32_t)abs(v) + 1u;
  const int k = 31 - __clz(m);
  const uint32_t low = m & ((1u << k) - 1u);
  code = (spread16(low) << 2) | 2u | (v < 0 ? 1u : 0u);
  nb = 2 * k + 2;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += n;
  }
  total = __shfl_sync(FULL, incl, 31);
  return incl - v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ unsigned warp_sum_u(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}


// Per-warp band table: quantiser parameters of every band for the slice's current index.
struct BandQ {
  uint32_t qm, ql, qf, qo;
};

// fill bq[0..nbands) for quantiser index q (lane b handles band b); returns false when the reference
// would throw "quantization index exceeds maximum implemented value" (Quantisation.cpp:60-63)
__device__ __forceinline__ bool fill_bandq(BandQ* bq, const SliceGeom& g, int q, int lane) {
  bool ok = true;
  __syncwarp();
  if (lane < g.nbands) {
    const int aq = max(q - g.qmatrix[lane], 0);
    ok = aq <= 119;
    const QParam p = qparam(aq);
    bq[lane].qm = p.qm; bq[lane].ql = p.ql; bq[lane].qf = p.qf; bq[lane].qo = p.qo;
  }
  ok = __all_sync(FULL, ok);
  __syncwarp();
  return ok;
}

// One pass over a slice component held in shared memory as 32 lane runs (run-major, odd stride):
// lane L owns coefficients [L*rl, min(n, (L+1)*rl)) of the component's coding order.
//  MODE 0: length only, coefficients stay unquantised (rate-control probe)
//  MODE 1: length, and the quantised values replace the coefficients (final pass)
//  MODE 2: luma squared error of quantise + inverse quantise (yss_for_slice); returns 0
// Returns component_slice_bytes' "count" = bits up to and including the last non-zero coefficient;
// lane_off = this lane's bit offset; range_err set if |quantised| >= 65535.
template <int MODE, bool QUANT>
__device__ __forceinline__ int comp_pass(int* run, int rl, int n, const int* bstart, int nbands, const BandQ* bq, int lane,
                                         int& lane_off, bool& range_err, long long& sse) {
  const int i0 = min(lane * rl, n), i1 = min(i0 + rl, n);
  int b = 0;
  BandQ q;
  int bend = 0;
  if (QUANT) {
    while (b + 1 < nbands && i0 >= bstart[b + 1]) ++b;
    bend = bstart[b + 1];
    q = bq[b];
  }
  int bits = 0, last = 0;
  long long acc = 0;
  for (int i = i0; i < i1; ++i) {
    if (QUANT && i >= bend) {
      ++b;
      bend = bstart[b + 1];
      q = bq[b];
    }
    const int v = run[i - i0];
    int qv = QUANT ? quant_one(v, q.qm, q.ql) : v;
    if (MODE == 2) {
      const int d = v - scale_one(qv, q.qf, q.qo);
      acc += (long long)(int)((unsigned)d * (unsigned)d);   // product in int, sum in long long (Quantisation.cpp:637-641)
    } else {
      if (abs(qv) >= 65535) {   // outside the reference's 32-bit VLC domain (VLC.h:27-28): flag, keep the bit IO sane
        range_err = true;
        qv = qv < 0 ? -65534 : 65534;
      }
      const int nb = vlc_bits(qv);
      bits += nb;
      if (nb > 1) last = bits;
      if (MODE == 1) run[i - i0] = qv;
    }
  }
  if (MODE == 2) {
    sse = warp_sum_ll(acc);
    return 0;
  }
  int total;
  lane_off = warp_excl_scan(bits, lane, total);
  return warp_max(last > 0 ? lane_off + last : 0);
}

// component_slice_bytes (Slices.cpp:114-118): whole scalar units; err when the length byte overflows
__device__ __forceinline__ int scaled_bytes(int count, int scalar, bool& too_big) {
  const int units = ((count + 7) / 8 + scalar - 1) / scalar;
  if (units > 0xFF) too_big = true;
  return units * scalar;
}

__device__ __forceinline__ void img_or(uint32_t* img, int w, uint32_t word, int endbit) {
  const int b0 = w << 5;
  if (b0 >= endbit) return;
  if (b0 + 32 > endbit) word &= ~0u << (b0 + 32 - endbit);
  if (word) atomicOr(&img[w], word);
}
__device__ __forceinline__ void img_put_byte(uint32_t* img, int byte_pos, uint32_t value) {
  atomicOr(&img[byte_pos >> 2], (value & 0xFFu) << (24 - 8 * (byte_pos & 3)));
}

// write the VLC codes of this lane's run at bit position startbit; bits >= endbit are dropped
// (they can only be the '1' codes of trailing zeros, VLC.cpp:151-155)
__device__ __forceinline__ void emit_run(uint32_t* img, const int* run, int cnt, int startbit, int endbit) {
  int w = startbit >> 5;
  int nacc = startbit & 31;
  unsigned long long acc = 0;
  for (int i = 0; i < cnt; ++i) {
    uint32_t code;
    int nb;
    vlc_code(run[i], code, nb);
    acc = (acc << nb) | code;
    nacc += nb;
    if (nacc >= 32) {
      img_or(img, w, (uint32_t)(acc >> (nacc - 32)), endbit);
      ++w;
      nacc -= 32;
      acc &= (1ull << nacc) - 1ull;
    }
  }
  if (nacc > 0) img_or(img, w, (uint32_t)(acc << (32 - nacc)), endbit);
}

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// ------------------------------------------------------------------------------------------
// HQ slice encoder: one warp per slice.
//   stream the slice's coefficients (contiguous in the slice-major layout) -> quantise ->
//   [quantIndicesCBR] -> per-lane code lengths + warp scan -> bit-pack into a shared-memory
//   slice image -> global offset (decoupled look-back over CTAs, or a priori for CBR) -> copy out.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hq_pack_kernel(const PackParams p) {
  extern __shared__ uint32_t smem[];
  __shared__ unsigned s_ticket;
  __shared__ unsigned s_wtot[32];
  __shared__ unsigned s_cta_prefix;
  __shared__ BandQ s_bq[8][VC2_MAX_BANDS];

  const SliceGeom& g = p.g;
  const int pic = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = p.warps_per_cta;
  const int nslices = g.slices_x * g.slices_y;

  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.ticket[pic], 1u);
  __syncthreads();
  const unsigned ticket = s_ticket;

  const int s = ticket * W + warp;
  const bool active = s < nslices;
  int* cf = reinterpret_cast<int*>(smem) + (size_t)warp * (p.coef_words + p.img_words);
  uint32_t* img = reinterpret_cast<uint32_t*>(cf + p.coef_words);
  BandQ* bq = s_bq[warp];

  unsigned flags = 0;
  int total = 0;
  int qi = 0;
  int Lc[3] = {0, 0, 0};
  int loff[3] = {0, 0, 0};

  if (active) {
    if (!p.search) qi = p.const_q >= 0 ? p.const_q : p.qidx[(long long)pic * nslices + s];
    const bool quant_on_load = p.quantise && !p.search;
    if (quant_on_load && !fill_bandq(bq, g, qi, lane)) flags |= VC2_FLAG_QUANT_INDEX;

    // ---- stream the slice in: coefficient i of component c -> run i / rl, slot i % rl (conflict-free both ways)
    const int32_t* src = p.coef + (long long)pic * g.coef_pic_stride + (long long)s * g.comp_start[3];
    for (int c = 0; c < 3; ++c) {
      const int n = g.band_start[c][g.nbands];
      const int rl = p.run_len[c], rs = p.run_stride[c];
      const int lg = (rl & (rl - 1)) == 0 ? 31 - __clz(rl) : -1;
      int* dst = cf + p.run_base[c];
      const int32_t* sc = src + g.comp_start[c];
      int b = 0, bend = g.band_start[c][1];
      BandQ q = bq[0];
      for (int i = lane; i < n; i += 32) {
        int v = sc[i];
        if (quant_on_load) {
          while (i >= bend) { ++b; bend = g.band_start[c][b + 1]; q = bq[b]; }
          v = quant_one(v, q.qm, q.ql);
        }
        const int r = lg >= 0 ? (i >> lg) : (i / rl);
        dst[r * rs + (i - r * rl)] = v;
      }
    }
    __syncwarp();

    int lane_off;
    bool range_err = false;
    long long sse = 0;

    // ---- rate control: literal replay of quantIndicesCBR (EncodeStream.cpp:85-122)
    if (p.search) {
      const int avail = p.slice_bytes[s] - 4;
      int trialQ = 63, q = 127, delta = 64;
      bool dead = false;
      while (delta > 0 && !dead) {
        delta >>= 1;
        if (!fill_bandq(bq, g, trialQ, lane)) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; dead = true; break; }
        int need = 0;
        bool too_big = false;
        for (int c = 0; c < 3; ++c) {
          const int n = g.band_start[c][g.nbands];
          const int count = comp_pass<0, true>(cf + p.run_base[c] + lane * p.run_stride[c], p.run_len[c], n, g.band_start[c], g.nbands,
                                               bq, lane, lane_off, range_err, sse);
          need += scaled_bytes(count, g.scalar, too_big);
        }
        if (too_big) { flags |= VC2_FLAG_SCALAR_TOO_SMALL | VC2_FLAG_SEARCH_PHASE; dead = true; break; }
        if (need <= avail) { if (trialQ < q) q = trialQ; trialQ -= delta; }
        else trialQ += delta;
      }
      if (!dead) {
        // "try a few higher quantisers": keep going while the luma squared error strictly drops
        trialQ = q;
        const int ny = g.band_start[0][g.nbands];
        int* yrun = cf + p.run_base[0] + lane * p.run_stride[0];
        if (!fill_bandq(bq, g, trialQ, lane)) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; dead = true; }
        long long prev = 0;
        if (!dead) comp_pass<2, true>(yrun, p.run_len[0], ny, g.band_start[0], g.nbands, bq, lane, lane_off, range_err, prev);
        while (!dead) {
          ++trialQ;
          if (!fill_bandq(bq, g, trialQ, lane)) { flags |= VC2_FLAG_QUANT_INDEX | VC2_FLAG_SEARCH_PHASE; dead = true; break; }
          long long cur;
          comp_pass<2, true>(yrun, p.run_len[0], ny, g.band_start[0], g.nbands, bq, lane, lane_off, range_err, cur);
          const long long d = cur - prev;
          prev = cur;
          if (!(d < 0)) break;
        }
        q = trialQ - 1;
      }
      qi = dead ? 0 : q;
      range_err = false;
    }
    if ((p.search || p.const_q >= 0) && lane == 0) p.qidx[(long long)pic * nslices + s] = qi;

    if (p.emit) {
      // ---- final quantisation (CBR: the staged coefficients are still raw) and component lengths
      const bool quant_now = p.quantise && p.search;
      if (quant_now && !fill_bandq(bq, g, qi, lane)) flags |= VC2_FLAG_QUANT_INDEX;
      bool too_big = false;
      int need[3];
      for (int c = 0; c < 3; ++c) {
        const int n = g.band_start[c][g.nbands];
        int* run = cf + p.run_base[c] + lane * p.run_stride[c];
        int count;
        if (quant_now) count = comp_pass<1, true>(run, p.run_len[c], n, g.band_start[c], g.nbands, bq, lane, lane_off, range_err, sse);
        else count = comp_pass<1, false>(run, p.run_len[c], n, g.band_start[c], g.nbands, bq, lane, lane_off, range_err, sse);
        loff[c] = lane_off;
        need[c] = scaled_bytes(count, g.scalar, too_big);
        Lc[c] = need[c];
      }
      if (too_big) flags |= VC2_FLAG_SCALAR_TOO_SMALL;
      if (__any_sync(FULL, range_err)) flags |= VC2_FLAG_VLC_RANGE;
      if (p.mode == VC2_HQ_CBR) {
        // V takes all remaining bytes (Slices.cpp:355-366)
        const int vBytes = p.slice_bytes[s] - 4 - Lc[0] - Lc[1];
        if (vBytes < need[2]) flags |= VC2_FLAG_CBR_TOO_MANY_BYTES;
        else if (vBytes / g.scalar > 255) flags |= VC2_FLAG_CBR_COMP_LENGTH;
        Lc[2] = max(vBytes, 0);
      }
      total = g.prefix + 4 + Lc[0] + Lc[1] + Lc[2];
      if (total > p.img_words * 4 - 8) {   // cannot happen for well-formed parameters; keep shared memory safe
        flags |= VC2_FLAG_SCALAR_TOO_SMALL;
        total = 0;
      }
      if (flags & ~VC2_FLAG_VLC_RANGE) total = (p.mode == VC2_HQ_CBR && p.fixed_off) ? total : 0;

      // ---- build the slice image: prefix | qindex | len Y | Y | len U | U | len V | V   (Slices.cpp:478-530)
      const int words = ((total + 3) >> 2) + 2;
      for (int w = lane; w < words; w += 32) img[w] = 0;
      __syncwarp();
      if (total > 0 && !(flags & ~VC2_FLAG_VLC_RANGE)) {
        int pos = g.prefix;
        if (lane == 0) img_put_byte(img, pos, (uint32_t)qi);
        ++pos;
        for (int c = 0; c < 3; ++c) {
          if (lane == 0) img_put_byte(img, pos, (uint32_t)(Lc[c] / g.scalar));
          ++pos;
          const int n = g.band_start[c][g.nbands];
          const int rl = p.run_len[c];
          const int i0 = min(lane * rl, n), cnt = min(i0 + rl, n) - i0;
          emit_run(img, cf + p.run_base[c] + lane * p.run_stride[c], cnt, 8 * pos + loff[c], 8 * (pos + Lc[c]));
          pos += Lc[c];
        }
      }
      __syncwarp();
    }
    if (lane == 0) p.err_flags[(long long)pic * nslices + s] = flags;
  }

  if (!p.emit) return;

  // ---- where does this slice go?  CBR: known a priori.  VBR: exclusive scan of slice sizes in
  //      raster order = decoupled look-back across CTAs (tickets make CTA order == launch order).
  unsigned offset;
  if (p.fixed_off) {
    offset = active ? p.fixed_off[s] : 0u;
  } else {
    if (lane == 0) s_wtot[warp] = active ? (unsigned)total : 0u;
    __syncthreads();
    if (warp == 0) {
      const unsigned v = lane < W ? s_wtot[lane] : 0u;
      int tot;
      const unsigned excl = (unsigned)warp_excl_scan((int)v, lane, tot);
      const unsigned agg = (unsigned)tot;
      if (lane < W) s_wtot[lane] = excl;
      unsigned long long* st = p.tile_state + (long long)pic * p.ctas_per_pic;
      unsigned prefix = 0;
      if (ticket == 0) {
        if (lane == 0) atomicExch(&st[0], (2ull << 32) | agg);
      } else {
        if (lane == 0) atomicExch(&st[ticket], (1ull << 32) | agg);
        int pos = (int)ticket - 1;
        while (true) {
          const int idx = pos - lane;
          const unsigned long long val = idx >= 0 ? ld_state(&st[idx]) : (2ull << 32);
          const unsigned f = (unsigned)(val >> 32);
          if (__any_sync(FULL, f == 0)) { __nanosleep(40); continue; }
          const unsigned m = __ballot_sync(FULL, f == 2);
          const int first = m ? (__ffs(m) - 1) : 32;
          prefix += warp_sum_u(lane <= first ? (unsigned)val : 0u);
          if (m) break;
          pos -= 32;
        }
        if (lane == 0) atomicExch(&st[ticket], (2ull << 32) | (unsigned long long)(prefix + agg));
      }
      if (lane == 0) s_cta_prefix = prefix;
    }
    __syncthreads();
    offset = s_cta_prefix + s_wtot[warp];
  }

  if (active) {
    uint32_t* so = p.slice_off + (long long)pic * (nslices + 1);
    if (lane == 0) {
      so[s] = offset;
      if (s == nslices - 1) so[nslices] = offset + (unsigned)total;
    }
    if ((long long)offset + total > p.out_capacity) {
      if (lane == 0) p.err_flags[(long long)pic * nslices + s] = flags | VC2_FLAG_STREAM;
    } else if (total > 0) {
      // copy out: head bytes up to 4-byte alignment, aligned 32-bit words, tail bytes
      uint8_t* dst = p.out + (long long)pic * p.out_pic_stride + offset;
      const int head = min((int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3), total);
      if (lane < head) dst[lane] = (uint8_t)(img[0] >> (24 - 8 * lane));
      const int nwords = (total - head) >> 2;
      uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
      for (int m = lane; m < nwords; m += 32) {
        const int qb = head + 4 * m;   // image byte index; qb & 3 == head
        const uint32_t be = __funnelshift_l(img[(qb >> 2) + 1], img[qb >> 2], 8 * head);
        dw[m] = __byte_perm(be, 0, 0x0123);
      }
      const int done = head + 4 * nwords, tail = total - done;
      if (lane < tail) {
        const int i = done + lane;
        dst[i] = (uint8_t)(img[i >> 2] >> (24 - 8 * (i & 3)));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Bit reader over shared memory, MSB first: 64-bit buffer refilled 32 bits at a time;
// bytes at or beyond the component length read as 0xFF (VLC.cpp:182-185).
// ------------------------------------------------------------------------------------------
struct BitReader {
  const uint32_t* words;   // aligned words containing the data (shared or global)
  unsigned long long buf;  // valid bits are the top nb bits
  int nb;
  int nextw;               // index of the next aligned word to load
  int sh;                  // 8 * (start byte & 3): data is shifted inside the aligned words
  int pos;                 // bytes of data already pulled into buf (multiple of 4)
  int len;                 // bytes of real data
  uint32_t carry;          // previous aligned word (big endian)
  int bits_left;           // BOUNDED reads only (LD blocks end on arbitrary bits): readable bits from here

  __device__ __forceinline__ uint32_t ldw(int i) const { return __byte_perm(words[i], 0, 0x0123); }

  __device__ __forceinline__ void init(const uint32_t* aligned_base, int start_byte, int nbytes) {
    words = aligned_base;
    nextw = start_byte >> 2;
    sh = 8 * (start_byte & 3);
    len = nbytes;
    pos = 0;
    buf = 0;
    nb = 0;
    carry = ldw(nextw++);
    refill();
    refill();
  }
  // next 32 data bits (big endian), 0xFF beyond len
  __device__ __forceinline__ uint32_t next32() {
    uint32_t w;
    if (pos >= len) w = 0xFFFFFFFFu;
    else {
      const uint32_t nw = ldw(nextw++);
      w = __funnelshift_l(nw, carry, sh);
      carry = nw;
      const int left = len - pos;
      if (left < 4) w |= 0xFFFFFFFFu >> (8 * left);
    }
    pos += 4;
    return w;
  }
  __device__ __forceinline__ void refill() {
    if (nb <= 32) {
      buf |= (unsigned long long)next32() << (32 - nb);
      nb += 32;
    }
  }
  __device__ __forceinline__ uint32_t peek() const { return (uint32_t)(buf >> 32); }
  __device__ __forceinline__ void skip(int n) { buf <<= n; nb -= n; }
  // one signed interleaved exp-Golomb value (VLC.cpp:283-317); nb >= 32 on entry.
  // BOUNDED: bits beyond bits_left read as ones (vlc::bounded, VLC.cpp:182-185)
  template <bool BOUNDED>
  __device__ __forceinline__ int get_vlc(bool& range_err) {
    uint32_t w = peek();
    if (BOUNDED && bits_left < 32) w |= bits_left <= 0 ? 0xFFFFFFFFu : (0xFFFFFFFFu >> bits_left);
    const uint32_t f = w & 0xAAAAAAAAu;
    if (f == 0) {   // more than 16 magnitude bits: outside the reference's 32-bit VLC domain
      range_err = true;
      skip(32);
      if (BOUNDED) bits_left -= 32;
      return 0;
    }
    const int k = __clz(f) >> 1;
    if (k == 0) { skip(1); if (BOUNDED) bits_left -= 1; return 0; }
    const uint32_t t = w >> (32 - 2 * k);
    const uint32_t m = (1u << k) | compress16(t);
    const int v = (int)m - 1;
    const int neg = (w >> (30 - 2 * k)) & 1u;
    skip(2 * k + 2);
    if (BOUNDED) bits_left -= 2 * k + 2;
    return neg ? -v : v;
  }
  __device__ __forceinline__ uint32_t get_bits(int n) {   // n in 1..32, nb >= 32
    const uint32_t w = peek();
    skip(n);
    return n == 32 ? w : (w >> (32 - n));
  }
};

// 8-bit prefix table: entry = (value << 4) | length for codes of at most 8 bits, 0 when the code is longer
__device__ __forceinline__ int16_t vlc_lut_entry(int idx) {
  const uint32_t w = (uint32_t)idx << 24;
  const uint32_t f = w & 0xAA000000u;
  if (f == 0) return 0;
  const int k = __clz(f) >> 1;
  if (k == 0) return 1;                  // value 0, 1 bit
  const int nbits = 2 * k + 2;
  if (nbits > 8) return 0;
  const uint32_t t = w >> (32 - 2 * k);
  const int v = (int)((1u << k) | compress16(t)) - 1;
  const int neg = (w >> (30 - 2 * k)) & 1u;
  return (int16_t)(((neg ? -v : v) << 4) | nbits);
}

// ------------------------------------------------------------------------------------------
// HQ / LD slice decoder.  A CTA takes G = 32 consecutive slices: their bytes are one contiguous
// run of the payload and are staged in shared memory with coalesced loads.  One thread decodes one
// (slice, component) bitstream; a warp owns the same component of the 32 slices, so every thread
// of a warp decodes the same number of coefficients (no divergence on the loop structure).
// Output is contiguous per thread in the slice-major layout: 128-bit stores.
// ------------------------------------------------------------------------------------------
constexpr int UNPACK_G = 32;

template <bool LD>
__global__ void __launch_bounds__(96) slice_unpack_kernel(const UnpackParams p, int stage_bytes) {
  extern __shared__ uint32_t stage[];
  __shared__ int16_t s_lut[256];
  const SliceGeom& g = p.g;
  const int nslices = g.slices_x * g.slices_y;
  const int pic = blockIdx.y;
  const int s0 = blockIdx.x * UNPACK_G;
  const int lane = threadIdx.x & 31, strm = threadIdx.x >> 5;   // warp = stream class: Y,U,V (HQ) / Y,UV (LD)
  const int nact = min(UNPACK_G, nslices - s0);

  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = vlc_lut_entry(i);

  const uint32_t* so = p.slice_off + (long long)pic * p.slice_off_pic_stride;
  const uint32_t r0 = so[s0], r1 = so[s0 + nact];
  const uint8_t* base = p.in + (long long)pic * p.in_pic_stride;
  // stage [r0, r1) at its own 4-byte phase so that aligned global words map to aligned shared words
  const uintptr_t ga = reinterpret_cast<uintptr_t>(base) + r0;
  const int phase = (int)(ga & 3);
  const uint32_t* gw = reinterpret_cast<const uint32_t*>(ga - phase);
  const int nwords = (int)((phase + (r1 - r0) + 3) >> 2) + 2;   // + slack for the reader's look-ahead (buffer has slack too)
  const bool staged = nwords * 4 <= stage_bytes;
  if (staged)
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) stage[i] = gw[i];
  __syncthreads();
  const uint32_t* words = staged ? stage : gw;

  const int s = s0 + lane;
  if (s >= nslices || strm >= (LD ? 2 : 3)) return;
  const int off = phase + (int)(so[s] - r0), size = (int)(so[s + 1] - so[s]);
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(words);
  int32_t* out = p.coef + (long long)pic * g.coef_pic_stride + (long long)s * g.comp_start[3];
  unsigned flags = 0;
  bool range_err = false;
  int qi;
  BitReader br;
  int c0 = strm, ncomp_here = 1;

  if (!LD) {
    // prefix | qindex | len | data | len | data | len | data   (Slices.cpp:544-605)
    int pos = g.prefix;
    int len = 0, start = 0;
    bool bad = size < g.prefix + 4;
    qi = bad ? 0 : bytes[off + pos];
    ++pos;
    for (int c = 0; c <= strm && !bad; ++c) {
      len = bytes[off + pos] * g.scalar;
      start = pos + 1;
      pos = start + len;
      if (pos + (2 - c) > size) bad = true;
    }
    if (bad) { flags |= VC2_FLAG_STREAM; len = 0; start = 0; }
    br.init(words, off + start, len);
  } else {
    // qindex (7 bits) | luma length | luma (bounded) | U/V interleaved (bounded)   (Slices.cpp:253-296)
    br.init(words, off, size);
    qi = (int)br.get_bits(7);
    br.refill();
    int lb = 0;   // utils::intlog2(8*bytes-7)
    { const int v = 8 * size - 7; while ((1 << lb) < v) ++lb; }
    const int ybits = (int)br.get_bits(lb);
    const int uvbits = 8 * size - 7 - lb - ybits;
    if (uvbits < 0) flags |= VC2_FLAG_STREAM;
    // the two blocks end on arbitrary bits: restart the reader at the byte holding the block's first bit,
    // drop the leading bits, and bound the readable bits (everything beyond reads as ones)
    const int ystart = 7 + lb;
    const int bstart = strm == 0 ? ystart : ystart + ybits, bbits = strm == 0 ? ybits : max(uvbits, 0);
    const int rb = min(bstart >> 3, size);
    br.init(words, off + rb, size - rb);
    br.skip(bstart & 7);
    br.refill();
    br.bits_left = bbits;
    if (strm == 1) { ncomp_here = 2; c0 = 1; }
  }
  if (strm == 0) p.qidx[(long long)pic * nslices + s] = qi;

  int32_t* dstc[2];
  dstc[0] = out + g.comp_start[c0];
  dstc[1] = ncomp_here == 2 ? out + g.comp_start[2] : dstc[0];

  for (int b = 0; b < g.nbands; ++b) {
    const int aq = max(qi - g.qmatrix[b], 0);
    if (aq > 119 && p.dequantise) flags |= VC2_FLAG_QUANT_INDEX;
    const QParam qp = qparam(aq);
    const bool deq = p.dequantise && !(LD && b == 0);
    const int n = g.band_start[c0][b + 1] - g.band_start[c0][b];
    int i = g.band_start[c0][b];
    const int iend = i + n;
    auto one = [&]() -> int {
      br.refill();
      int v;
      if (LD) {
        v = br.get_vlc<true>(range_err);
      } else {
        const int ent = s_lut[br.peek() >> 24];
        if (ent) { v = ent >> 4; br.skip(ent & 15); }
        else v = br.get_vlc<false>(range_err);
      }
      return deq ? scale_one(v, qp.qf, qp.qo) : v;
    };
    if (ncomp_here == 1) {
      // vector part when the band start is 16-byte aligned inside the slice block
      if (((i | n) & 3) == 0 && ((g.comp_start[3] | g.comp_start[c0]) & 3) == 0) {
        for (; i < iend; i += 4) {
          int4 q;
          q.x = one(); q.y = one(); q.z = one(); q.w = one();
          *reinterpret_cast<int4*>(dstc[0] + i) = q;
        }
      } else {
        for (; i < iend; ++i) dstc[0][i] = one();
      }
    } else {
      for (; i < iend; ++i) {   // LD chroma: u0 v0 u1 v1 ... (Slices.cpp:287-294)
        dstc[0][i] = one();
        dstc[1][i] = one();
      }
    }
  }
  if (range_err) flags |= VC2_FLAG_VLC_RANGE;
  if (flags) atomicOr(&p.err_flags[(long long)pic * nslices + s], flags);
}

// ------------------------------------------------------------------------------------------
// stand-alone quantise / dequantise of an in-place ordered plane (Library surface)
// ------------------------------------------------------------------------------------------
__global__ void quant_kernel(const QuantParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= p.pw || y >= p.ph) return;
  const int d = p.depth;
  const int t = (y | x) & ((1 << d) - 1);
  int band = 0;
  if (t) {
    const int l = __ffs(t) - 1, L = d - l;
    const int hx = (x >> l) & 1, hy = (y >> l) & 1;
    band = 3 * (L - 1) + (hx ? (hy ? 3 : 1) : 2);
  }
  const long long i = (long long)y * p.pw + x;
  if (band == 0 && p.skip_ll) { p.dst[i] = p.src[i]; return; }
  const int sy = y / (p.ph / p.slices_y), sx = x / (p.pw / p.slices_x);
  const int q = max(p.qidx[sy * p.slices_x + sx] - p.qmatrix[band], 0);
  const QParam qp = qparam(q);
  const int v = p.src[i];
  p.dst[i] = p.inverse ? scale_one(v, qp.qf, qp.qo) : quant_one(v, qp.qm, qp.ql);
}

// ------------------------------------------------------------------------------------------
// LD LL-band reconstruction: rec[y][x] = scale(c, q'(slice)) + predictDC(rec, y, x), raster
// recurrence over the whole band (Quantisation.cpp:191-208, 287-306).  One CTA per plane runs
// the anti-diagonal wavefront: element (y, x) depends on (y-1, x-1), (y-1, x), (y, x-1).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) ld_dc_kernel(const LdDcParams p) {
  const int H = p.H, Wd = p.W;
  auto at = [&](int y, int x) -> int32_t* {
    const int sy = y / p.bh, sx = x / p.bw;
    return p.base + sy * p.A + (y - sy * p.bh) * p.B + sx * p.C + (x - sx * p.bw) * p.D;
  };
  for (int diag = 0; diag < H + Wd - 1; ++diag) {
    const int ylo = max(0, diag - (Wd - 1)), yhi = min(H - 1, diag);
    for (int y = ylo + (int)threadIdx.x; y <= yhi; y += blockDim.x) {
      const int x = diag - y;
      int32_t* e = at(y, x);
      const int yb = ((y + 1) * p.slices_y - 1) / H, xb = ((x + 1) * p.slices_x - 1) / Wd;
      const int q = max(p.qidx[yb * p.slices_x + xb] - p.qm0, 0);
      const QParam qp = qparam(q);
      int pred;
      if (y > 0 && x > 0) {
        const int sum = *at(y - 1, x - 1) + *at(y - 1, x) + *at(y, x - 1);
        pred = sum >= 0 ? (sum + 1) / 3 : (sum - 1) / 3;
      } else if (y > 0) pred = *at(y - 1, x);
      else if (x > 0) pred = *at(y, x - 1);
      else pred = 0;
      *e = scale_one(*e, qp.qf, qp.qo) + pred;
    }
    __syncthreads();
  }
}

}  // namespace

cudaError_t pack_launch(cudaStream_t s, const PackParams& p, int npictures, size_t smem_bytes) {
  static size_t attr_set = 0;
  if (smem_bytes > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(hq_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    attr_set = smem_bytes;
  }
  hq_pack_kernel<<<dim3(p.ctas_per_pic, npictures), p.warps_per_cta * 32, smem_bytes, s>>>(p);
  return cudaGetLastError();
}

cudaError_t unpack_launch(cudaStream_t s, const UnpackParams& p, int npictures) {
  const int nslices = p.g.slices_x * p.g.slices_y;
  const int stage_bytes = 40 * 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(slice_unpack_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(slice_unpack_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage_bytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const dim3 grid((nslices + UNPACK_G - 1) / UNPACK_G, npictures);
  if (p.ld) slice_unpack_kernel<true><<<grid, 96, stage_bytes, s>>>(p, stage_bytes);
  else slice_unpack_kernel<false><<<grid, 96, stage_bytes, s>>>(p, stage_bytes);
  return cudaGetLastError();
}

cudaError_t quant_launch(cudaStream_t s, const QuantParams& p) {
  const dim3 block(32, 8), grid((p.pw + 31) / 32, (p.ph + 7) / 8);
  quant_kernel<<<grid, block, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t ld_dc_launch(cudaStream_t s, const LdDcParams& p) {
  ld_dc_kernel<<<1, 1024, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace vc2
