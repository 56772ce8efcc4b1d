// Forward / inverse lifting level kernels (streaming, register resident) and the in-place <->
// group-interleaved layout kernels.  See dwt.cuh for the reference line citations.
//
// One WARP owns a vertical strip of 32*V lattice columns (lane = V consecutive columns, V/2 sample
// pairs) and walks down a segment of rows two at a time.  Nothing is staged in shared memory:
//   * horizontal lifting of a row runs in registers, neighbouring pairs come from the adjacent
//     lanes by warp shuffle;
//   * vertical lifting keeps, per lane and column, two small rings of rows in registers (rows of
//     the parity that feeds the first step, and rows of the other parity) and applies step i to
//     the row that lags the newest loaded row by lag(i) - the earliest moment at which all of its
//     taps have completed step i-1 AND nobody needs its own previous value any more.  Ring rows are
//     indexed by age (compile-time registers) and age by one slot per row pair.
// The reference's edge rule (a tap outside the array uses the nearest sample of the same parity)
// is reproduced by an "edge" variant of each step taken only for rows near the top/bottom (the tap
// is redirected to the first row of that parity, or to a register copy of the last one), and by
// extending the source-parity sequence past the left/right border before each horizontal step.
#include "dwt_lift.cuh"

#ifndef VC2_DWT_PART
#error "compile with -DVC2_DWT_PART=1 (forward) or =2 (inverse + layout)"
#endif

namespace vc2 {

namespace {

constexpr int WARPS = 4;   // warps per CTA: four vertically adjacent row segments of one strip
#ifndef VC2_DWT_NO_SWPIPE
#define VC2_DWT_SWPIPE 1   // forward fast loop, 16-bit samples: the next row pair is loaded before the current one is used
#endif
#ifndef VC2_DWT_MINB
#define VC2_DWT_MINB 4     // resident CTAs per SM the register allocation is held to (4 -> 128 registers per thread: the fast loop spills at 5)
#endif

// ---- compile-time schedule of the vertical pipeline ------------------------------------------------
template <int K, int DIR>
struct Sched {
  static constexpr int n = Wavelet<K>::NSTEPS;
  // i-th step executed: forward = declared order, inverse = reversed
  __host__ __device__ static constexpr int sidx(int i) { return DIR > 0 ? i : n - 1 - i; }
  __host__ __device__ static constexpr int par(int i) {
    return sidx(i) == 0 ? Step<K, 0>::P : sidx(i) == 1 ? Step<K, 1>::P : sidx(i) == 2 ? Step<K, (n > 2 ? 2 : 0)>::P : Step<K, (n > 3 ? 3 : 0)>::P;
  }
  __host__ __device__ static constexpr int taps(int i) {
    return sidx(i) == 0 ? Step<K, 0>::N : sidx(i) == 1 ? Step<K, 1>::N : sidx(i) == 2 ? Step<K, (n > 2 ? 2 : 0)>::N : Step<K, (n > 3 ? 3 : 0)>::N;
  }
  __host__ __device__ static constexpr int reach(int i) { return 2 * taps(i) - 1; }
  static constexpr int PB = par(0);        // parity of the rows the first step updates ("B rows")
  static constexpr int PA = 1 - PB;        // parity of the rows the first step reads ("A rows")
  // step i targets row a - lag(i), a = newest A row
  __host__ __device__ static constexpr int lag(int i) { return i == 0 ? reach(0) : lag(i - 1) + cmax(reach(i), reach(i - 1)); }
  __host__ __device__ static constexpr int total_reach() { return reach(0) + reach(1) + (n > 2 ? reach(n > 2 ? 2 : 0) : 0) + (n > 3 ? reach(n > 3 ? 3 : 0) : 0); }
  static constexpr bool last_targets_A = ((n - 1) % 2) == 1;
  static constexpr int oldestA = last_targets_A ? lag(n - 1) : lag(n - 1) + reach(n - 1);   // back from a
  static constexpr int oldestB = last_targets_A ? lag(n - 1) + reach(n - 1) : lag(n - 1);
  static constexpr int WA = oldestA / 2 + 1;
  static constexpr int WB = (oldestB - reach(0)) / 2 + 1;
  static constexpr int WR = cmax(WA, WB);            // both rings are allocated with WR rows: the fast loop rotates them with one period
  // smallest newest-row index for which no step of the pair needs the edge rule at the top
  __host__ __device__ static constexpr int amin_i(int i) { return lag(i) + reach(i); }
  static constexpr int AMIN = cmax(cmax(amin_i(0), amin_i(1)), cmax(n > 2 ? amin_i(n > 2 ? 2 : 0) : 0, n > 3 ? amin_i(n > 3 ? 3 : 0) : 0));
};

template <int K> struct Vec { static constexpr int V = (K == VC2_FIDELITY) ? 4 : 8; };

template <int K>
struct Geo {
  static constexpr int V = Vec<K>::V, PPL = V / 2;
  static constexpr int R = Wavelet<K>::R;
  static constexpr int HX = (R + V - 1) / V * V;      // horizontal halo, whole lanes
  static constexpr int XW = 32 * V;                   // columns per warp
  static constexpr int XU = XW - 2 * HX;              // useful columns per warp
};


// ---- the per-lane state of the vertical pipeline ---------------------------------------------------
// Rings are indexed by AGE: A[i] = A row (a - 2i), B[i] = B row (r0 - 2i), a = newest A row of the current
// row pair, r0 = a - reach(0).  shift() ages every row by one pair (register moves).
template <int K, int DIR>
struct Rings {
  using SC = Sched<K, DIR>;
  static constexpr int V = Vec<K>::V;
  int A[SC::WR][V];
  int B[SC::WR][V];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int v = 0; v < V; ++v) {
#pragma unroll
      for (int i = 0; i < SC::WR; ++i) A[i][v] = 0;
#pragma unroll
      for (int i = 0; i < SC::WR; ++i) B[i][v] = 0;
    }
  }
  __device__ __forceinline__ void shift() {
#pragma unroll
    for (int v = 0; v < V; ++v) {
#pragma unroll
      for (int i = SC::WR - 1; i > 0; --i) A[i][v] = A[i - 1][v];
#pragma unroll
      for (int i = SC::WR - 1; i > 0; --i) B[i][v] = B[i - 1][v];
    }
  }
};

// ring row of a (warp uniform) run-time age: select chain, only used by the edge variant
template <int W, int V>
__device__ __forceinline__ int pick_row(const int (&ring)[W][V], int age, int v) {
  int x = ring[0][v];
#pragma unroll
  for (int i = 1; i < W; ++i) x = (age == i) ? ring[i][v] : x;
  return x;
}

// vertical step I of the schedule on target row r = a - lag(I).
// EDGE: taps above row 0 / below row `last` are redirected to the first / last row of their parity
// (its current state is still in the ring: the window of needed rows always contains it).
template <int K, int DIR, int I, bool EDGE>
__device__ __forceinline__ void vstep(Rings<K, DIR>& g, int a, int r, int last) {
  using SC = Sched<K, DIR>;
  constexpr int S = SC::sidx(I);
  using ST = Step<K, S>;
  constexpr int V = Vec<K>::V, N = ST::N;
  constexpr bool TA = (I % 2) == 1;               // target is an A row
  constexpr int L = SC::lag(I);
  constexpr int tage = TA ? L / 2 : (L - SC::reach(0)) / 2;
  // ages of the first and of the last row of the source parity
  const int newest = TA ? a - SC::reach(0) : a, spar = TA ? SC::PB : SC::PA;
  const int age_first = (newest - spar) >> 1, age_last = (newest - (last - 1 + spar)) >> 1;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    int lv[N], rv[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      int l, rr;
      if constexpr (TA) {   // sources are B rows r -/+ (2k+1)
        l = g.B[(L + (2 * k + 1) - SC::reach(0)) / 2][v];
        rr = g.B[(L - (2 * k + 1) - SC::reach(0)) / 2][v];
        if (EDGE) {
          if (r - (2 * k + 1) < 0) l = pick_row<SC::WR, V>(g.B, age_first, v);
          if (r + (2 * k + 1) > last) rr = pick_row<SC::WR, V>(g.B, age_last, v);
        }
      } else {              // sources are A rows
        l = g.A[(L + (2 * k + 1)) / 2][v];
        rr = g.A[(L - (2 * k + 1)) / 2][v];
        if (EDGE) {
          if (r - (2 * k + 1) < 0) l = pick_row<SC::WR, V>(g.A, age_first, v);
          if (r + (2 * k + 1) > last) rr = pick_row<SC::WR, V>(g.A, age_last, v);
        }
      }
      lv[k] = l; rv[k] = rr;
    }
    int& t = TA ? g.A[tage][v] : g.B[tage][v];
    t = lift_update<K, S, DIR>(t, [&](int k, bool right) { return right ? rv[k] : lv[k]; });
  }
}

// run step I for the current row pair if its target row exists
template <int K, int DIR, int I>
__device__ __forceinline__ void vstep_guarded(Rings<K, DIR>& g, int a, int H) {
  using SC = Sched<K, DIR>;
  const int r = a - SC::lag(I);
  if (r < 0 || r > H - 1) return;
  const int last = H - 1;
  if (r - SC::reach(I) < 0 || r + SC::reach(I) > last) vstep<K, DIR, I, true>(g, a, r, last);
  else vstep<K, DIR, I, false>(g, a, r, last);
}

// ---- the fast loop: interior rows of interior strips ------------------------------------------------
// No edge rule, no bounds or alignment tests, addresses from running pointers and 32-bit band indices.  The ring
// accesses are written for a ROTATED ring (row of age i in slot (i + R) % WR, R a template constant); the kernels
// use R = 0 after an ordinary shift, see the note at the loop.
template <int K, int DIR, int I, int R>
__device__ __forceinline__ void vstep_rot(Rings<K, DIR>& g) {
  using SC = Sched<K, DIR>;
  constexpr int S = SC::sidx(I);
  using ST = Step<K, S>;
  constexpr int V = Vec<K>::V, N = ST::N, W = SC::WR;
  constexpr bool TA = (I % 2) == 1;
  constexpr int L = SC::lag(I);
  constexpr int tage = TA ? L / 2 : (L - SC::reach(0)) / 2;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    int& t = TA ? g.A[(tage + R) % W][v] : g.B[(tage + R) % W][v];
    t = lift_update<K, S, DIR>(t, [&](int k, bool right) {
      if constexpr (TA) return right ? g.B[((L - (2 * k + 1) - SC::reach(0)) / 2 + R) % W][v] : g.B[((L + (2 * k + 1) - SC::reach(0)) / 2 + R) % W][v];
      else return right ? g.A[((L - (2 * k + 1)) / 2 + R) % W][v] : g.A[((L + (2 * k + 1)) / 2 + R) % W][v];
    });
  }
}
template <int K, int DIR, int R>
__device__ __forceinline__ void vsteps_rot(Rings<K, DIR>& g) {
  vstep_rot<K, DIR, 0, R>(g);
  vstep_rot<K, DIR, 1, R>(g);
  if constexpr (Sched<K, DIR>::n == 4) {
    vstep_rot<K, DIR, 2, R>(g);
    vstep_rot<K, DIR, 3, R>(g);
  }
}

// per-lane constants of the fast loop
struct FastCtx {
  int32_t* coefpic;          // this picture's coefficient block
  int sx, kx4;               // slice column of this lane's band samples, (column inside the slice part) / 4
  int lgbh, bhm1, bw4, nx, nc4;
  int o_ll, o_hl, o_lh, o_hh;   // band starts, as element offsets ((base / 4) * 128)
  int32_t* llp;              // compact LL plane at this lane's band column, or NULL (LL = band 0 of the block)
  int ll_pitch;
  bool store;                // this lane's columns are useful
  __device__ __forceinline__ int idx(int by) const {   // element index of the lane's 16-byte piece in band row by (band start 0)
    const int sy = by >> lgbh, ry = by & bhm1;
    const int s = sy * nx + sx;
    return ((((s >> 5) * nc4 + ry * bw4 + kx4) << 5) + (s & 31)) << 2;
  }
};


// per-warp constants of one strip segment
struct StripCtx {
  int pic, lane;
  int xs;            // lattice column of lane 0's first sample
  int x0;            // first useful lattice column of this strip
  int plo, phi;      // valid pair range inside the warp segment
  bool hedge;
  int H;             // lattice rows
  int y0, y1;        // rows to output: [y0, y1)
  int bxmax;         // last band column = lat_w / 2 - 1
  bool mine;         // this lane's columns are useful (not halo) and inside the lattice
  int pd;            // prefetch distance in row pairs
};

// ------------------------------------------------------------------------------------------
// forward level:  pix (dense plane)  ->  LL (compact plane or band 0), HL, LH, HH (interleaved)
// ------------------------------------------------------------------------------------------
template <int K, int KIND, int PPL>
__device__ __forceinline__ void fwd_fetch_row(const DwtComp& C, const StripCtx& S, int row, int (&dst)[2 * PPL]) {
  constexpr int V = 2 * PPL, SHIFT = Wavelet<K>::SHIFT;
  int x[V];
  load_pix<KIND, V>(C, S.pic, row, S.xs + V * S.lane, x);
  int e[PPL], o[PPL];
#pragma unroll
  for (int a = 0; a < PPL; ++a) {
    e[a] = (int)((unsigned)x[2 * a] << SHIFT);
    o[a] = (int)((unsigned)x[2 * a + 1] << SHIFT);
  }
  hsteps<K, +1, PPL>(e, o, S.lane, S.hedge, S.plo, S.phi);
#pragma unroll
  for (int a = 0; a < PPL; ++a) { dst[a] = e[a]; dst[PPL + a] = o[a]; }   // columns: [even samples | odd samples]
}

template <int K, int PPL>
__device__ __forceinline__ void fwd_emit_row(const DwtComp& C, const StripCtx& S, const BandAddr& ba, int row, const int (&src)[2 * PPL]) {
  if (!S.mine || row < S.y0 || row >= S.y1) return;
  int32_t* coef = C.coef + (long long)S.pic * C.coef_pic_stride;
  const int by = row >> 1, bx0 = ((S.xs >> 1) + PPL * S.lane);
  int lo[PPL], hi[PPL];
#pragma unroll
  for (int a = 0; a < PPL; ++a) { lo[a] = src[a]; hi[a] = src[PPL + a]; }
  if (row & 1) {
    band_access<PPL, true>(coef, ba, C.base_lh, by, bx0, S.bxmax, lo);
    band_access<PPL, true>(coef, ba, C.base_hh, by, bx0, S.bxmax, hi);
  } else {
    if (C.ll) ll_access<PPL, true>(C.ll + (long long)S.pic * C.ll_pic_stride, C.ll_pitch, by, bx0, S.bxmax, lo);
    else band_access<PPL, true>(coef, ba, C.base_ll, by, bx0, S.bxmax, lo);
    band_access<PPL, true>(coef, ba, C.base_hl, by, bx0, S.bxmax, hi);
  }
}

template <int K, int KIND>
__device__ __forceinline__ void fwd_pair(const DwtComp& C, const StripCtx& S, const BandAddr& ba, Rings<K, +1>& g, int tau) {
  using SC = Sched<K, +1>;
  constexpr int V = Vec<K>::V, PPL = V / 2, n = SC::n;
  const int H = S.H;
  const int a = 2 * tau + SC::PA;
  prefetch_pix<KIND, V>(C, S.pic, a + 2 * S.pd, S.xs + V * S.lane);
  prefetch_pix<KIND, V>(C, S.pic, a + 2 * S.pd - SC::reach(0), S.xs + V * S.lane);
  g.shift();
  if (a <= H - 1) fwd_fetch_row<K, KIND, PPL>(C, S, a, g.A[0]);
  const int r0 = a - SC::reach(0);
  if (r0 >= 0 && r0 <= H - 1) fwd_fetch_row<K, KIND, PPL>(C, S, r0, g.B[0]);
  vstep_guarded<K, +1, 0>(g, a, H);
  vstep_guarded<K, +1, 1>(g, a, H);
  if constexpr (n == 4) {
    vstep_guarded<K, +1, 2>(g, a, H);
    vstep_guarded<K, +1, 3>(g, a, H);
  }
  // rows that just became final: the targets of the last two steps
  constexpr int L1 = SC::lag(n - 1), L2 = SC::lag(n - 2);
  {
    const int r = a - L1;
    if (r >= 0 && r <= H - 1) {
      if constexpr (SC::last_targets_A) fwd_emit_row<K, PPL>(C, S, ba, r, g.A[L1 / 2]);
      else fwd_emit_row<K, PPL>(C, S, ba, r, g.B[(L1 - SC::reach(0)) / 2]);
    }
  }
  {
    const int r = a - L2;
    if (r >= 0 && r <= H - 1) {
      if constexpr (!SC::last_targets_A) fwd_emit_row<K, PPL>(C, S, ba, r, g.A[L2 / 2]);
      else fwd_emit_row<K, PPL>(C, S, ba, r, g.B[(L2 - SC::reach(0)) / 2]);
    }
  }
}

// ---- forward fast loop -------------------------------------------------------------------------------
template <int K, int KIND, int PPL>
__device__ __forceinline__ void fwd_fast_fetch(const DwtComp& C, const StripCtx& S, const uint8_t* p, int (&dst)[2 * PPL]) {
  constexpr int V = 2 * PPL, SHIFT = Wavelet<K>::SHIFT;
  static_assert(PPL == 4, "the fast loop moves 16-byte pieces");
  int x[V];
  if (KIND == SAMPLE_I32) {
#pragma unroll
    for (int j = 0; j < V; j += 4) {
      const int4 q = __ldg(reinterpret_cast<const int4*>(p) + j / 4);
      x[j] = q.x; x[j + 1] = q.y; x[j + 2] = q.z; x[j + 3] = q.w;
    }
  } else {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x[2 * j] = (int)(__byte_perm(w[j], 0, 0x4401) >> C.sshift) - C.soffset;
      x[2 * j + 1] = (int)(__byte_perm(w[j], 0, 0x4423) >> C.sshift) - C.soffset;
    }
  }
  int e[PPL], o[PPL];
#pragma unroll
  for (int a = 0; a < PPL; ++a) {
    e[a] = (int)((unsigned)x[2 * a] << SHIFT);
    o[a] = (int)((unsigned)x[2 * a + 1] << SHIFT);
  }
  hsteps<K, +1, PPL>(e, o, S.lane, false, 0, 0);
#pragma unroll
  for (int a = 0; a < PPL; ++a) { dst[a] = e[a]; dst[PPL + a] = o[a]; }
}

// the same from words that are already in registers (the loop loads one row pair ahead)
template <int K, int PPL>
__device__ __forceinline__ void fwd_fast_convert(const DwtComp& C, const StripCtx& S, const uint4 q, int (&dst)[2 * PPL]) {
  constexpr int V = 2 * PPL, SHIFT = Wavelet<K>::SHIFT;
  int x[V];
  const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    x[2 * j] = (int)(__byte_perm(w[j], 0, 0x4401) >> C.sshift) - C.soffset;
    x[2 * j + 1] = (int)(__byte_perm(w[j], 0, 0x4423) >> C.sshift) - C.soffset;
  }
  int e[PPL], o[PPL];
#pragma unroll
  for (int a = 0; a < PPL; ++a) {
    e[a] = (int)((unsigned)x[2 * a] << SHIFT);
    o[a] = (int)((unsigned)x[2 * a + 1] << SHIFT);
  }
  hsteps<K, +1, PPL>(e, o, S.lane, false, 0, 0);
#pragma unroll
  for (int a = 0; a < PPL; ++a) { dst[a] = e[a]; dst[PPL + a] = o[a]; }
}

template <int PPL, bool ODD>
__device__ __forceinline__ void fwd_fast_emit(const FastCtx& F, int row, const int (&src)[2 * PPL]) {
  if (!F.store) return;
  const int by = row >> 1, i = F.idx(by);
  const int4 lo = make_int4(src[0], src[1], src[2], src[3]), hi = make_int4(src[PPL], src[PPL + 1], src[PPL + 2], src[PPL + 3]);
  if (ODD) {
    *reinterpret_cast<int4*>(F.coefpic + i + F.o_lh) = lo;
    *reinterpret_cast<int4*>(F.coefpic + i + F.o_hh) = hi;
  } else {
    if (F.llp) *reinterpret_cast<int4*>(F.llp + (long long)by * F.ll_pitch) = lo;
    else *reinterpret_cast<int4*>(F.coefpic + i + F.o_ll) = lo;
    *reinterpret_cast<int4*>(F.coefpic + i + F.o_hl) = hi;
  }
}

// iteration T of a chunk; pA / pB run over the picture rows a and a - reach(0) of this lane
template <int K, int KIND, int T>
__device__ __forceinline__ void fwd_fast_iter(const DwtComp& C, const StripCtx& S, const FastCtx& F, Rings<K, +1>& g, int tau,
                                              const uint8_t*& pA, const uint8_t*& pB, long long step, long long pf, int pf_last,
                                              uint4& qa, uint4& qb, bool more) {
  using SC = Sched<K, +1>;
  constexpr int V = Vec<K>::V, PPL = V / 2, n = SC::n, W = SC::WR, R = (W - 1 - T % W) % W;
  const int a = 2 * tau + SC::PA;
  if (a <= pf_last) { prefetch_l1(pA + pf); prefetch_l1(pB + pf); }
#ifdef VC2_DWT_SWPIPE
  if constexpr (KIND == SAMPLE_U16BE) {
    const uint4 ca = qa, cb = qb;
    pA += step; pB += step;
    if (more) { qa = __ldg(reinterpret_cast<const uint4*>(pA)); qb = __ldg(reinterpret_cast<const uint4*>(pB)); }
    fwd_fast_convert<K, PPL>(C, S, ca, g.A[R]);
    fwd_fast_convert<K, PPL>(C, S, cb, g.B[R]);
  } else
#endif
  {
    fwd_fast_fetch<K, KIND, PPL>(C, S, pA, g.A[R]);
    fwd_fast_fetch<K, KIND, PPL>(C, S, pB, g.B[R]);
    pA += step; pB += step;
  }
  vsteps_rot<K, +1, R>(g);
  constexpr int L1 = SC::lag(n - 1), L2 = SC::lag(n - 2);
  if constexpr (SC::last_targets_A) {
    fwd_fast_emit<PPL, ((SC::PA - L1) & 1) != 0>(F, a - L1, g.A[(L1 / 2 + R) % W]);
    fwd_fast_emit<PPL, ((SC::PA - L2) & 1) != 0>(F, a - L2, g.B[((L2 - SC::reach(0)) / 2 + R) % W]);
  } else {
    fwd_fast_emit<PPL, ((SC::PA - L1) & 1) != 0>(F, a - L1, g.B[((L1 - SC::reach(0)) / 2 + R) % W]);
    fwd_fast_emit<PPL, ((SC::PA - L2) & 1) != 0>(F, a - L2, g.A[(L2 / 2 + R) % W]);
  }
}
// band side of the fast loop: false when this level / strip has to stay on the general loop
template <int K>
__device__ __forceinline__ bool fast_setup(const DwtComp& C, const StripCtx& S, FastCtx& F) {
  using G = Geo<K>;
  constexpr int PPL = G::V / 2;
  if (PPL != 4 || S.hedge) return false;
  if (C.lgbh < 0 || C.lgbw < 0 || (C.bw & 3) || ((C.base_ll | C.base_hl | C.base_lh | C.base_hh) & 3)) return false;
  const int bx0 = (S.xs >> 1) + PPL * S.lane;
  F.coefpic = C.coef + (long long)S.pic * C.coef_pic_stride;
  F.sx = bx0 >> C.lgbw;
  F.kx4 = (bx0 & (C.bw - 1)) >> 2;
  F.lgbh = C.lgbh; F.bhm1 = C.bh - 1; F.bw4 = C.bw >> 2; F.nx = C.nx; F.nc4 = C.NC >> 2;
  F.o_ll = (C.base_ll >> 2) * 128; F.o_hl = (C.base_hl >> 2) * 128; F.o_lh = (C.base_lh >> 2) * 128; F.o_hh = (C.base_hh >> 2) * 128;
  F.llp = nullptr; F.ll_pitch = C.ll_pitch;
  if (C.ll) {
    F.llp = C.ll + (long long)S.pic * C.ll_pic_stride + bx0;
    if ((C.ll_pitch & 3) || (reinterpret_cast<uintptr_t>(C.ll + (long long)S.pic * C.ll_pic_stride) & 15)) return false;
  }
  if (reinterpret_cast<uintptr_t>(F.coefpic) & 15) return false;
  F.store = S.mine;
  return true;
}

// common strip set-up; returns false when this warp has nothing to do
template <int K>
__device__ __forceinline__ bool strip_setup(const DwtComp& C, int seg_rows, StripCtx& S, bool inverse) {
  using G = Geo<K>;
  S.lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  S.x0 = blockIdx.x * G::XU;
  S.y0 = (blockIdx.y * WARPS + warp) * seg_rows;
  if (S.x0 >= C.lat_w || S.y0 >= C.lat_h) return false;
  if (inverse && (S.x0 >= C.pix_w || S.y0 >= C.pix_h)) return false;   // nothing of this strip survives the crop (WaveletTransform.cpp:340)
  S.y1 = min(S.y0 + seg_rows, C.lat_h);
  S.xs = S.x0 - G::HX;
  S.H = C.lat_h;
  S.plo = max(0, -S.xs / 2);
  S.phi = min(G::XW / 2 - 1, (C.lat_w - 2 - S.xs) / 2);
  S.hedge = S.xs < 0 || S.xs + G::XW > C.lat_w;
  S.bxmax = C.lat_w / 2 - 1;
  const int gx = S.xs + G::V * S.lane;
  S.mine = gx >= S.x0 && gx < S.x0 + G::XU && gx < C.lat_w;
  return true;
}

template <int K, int KIND>
__global__ void __launch_bounds__(32 * WARPS, VC2_DWT_MINB) dwt_fwd_kernel(const DwtParams p, int seg_rows) {
  using SC = Sched<K, +1>;
  const int comp = blockIdx.z % p.ncomp;
  const DwtComp& C = p.c[comp];
  StripCtx S;
  S.pic = blockIdx.z / p.ncomp;
  S.pd = p.pd;
  if (!strip_setup<K>(C, seg_rows, S, false)) return;
  const BandAddr ba = {C.bh, C.bw, C.lgbh, C.lgbw, C.nx, C.NC >> 2};
  Rings<K, +1> g;
  g.clear();
  // first row pair: early enough that every row of [y0, y1) has its whole dependency cone inside the walk;
  // last pair: until the last step has reached row y1 - 1
  const int tau0 = max((S.y0 - SC::total_reach() - SC::PA) >> 1, 0);
  const int tau_end = (S.y1 - 1 + SC::lag(SC::n - 1) - SC::PA + 2) >> 1;
  int tau = tau0;
  if constexpr (Vec<K>::V == 8 && KIND != SAMPLE_U8) {
    constexpr int W = SC::WR, esz = KIND == SAMPLE_I32 ? 4 : 2;
    FastCtx F;
    const uint8_t* base = (const uint8_t*)C.pix + (long long)S.pic * C.pix_pic_stride * (KIND == SAMPLE_I32 ? 4 : 1);
    bool fast = p.fast && fast_setup<K>(C, S, F) && S.xs + Geo<K>::XW <= C.pix_w && ((C.pix_pitch * esz) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(base) & 15) == 0;
    // newest rows a for which the pair needs no edge rule, no row replication, and emits two rows of [y0, y1)
    const int a_lo = max(SC::AMIN, S.y0 + SC::lag(SC::n - 1));
    const int a_hi = min(min(S.H, C.pix_h) - 1, S.y1 - 1 + SC::lag(SC::n - 2));
    const int tf0 = max(tau0, (a_lo - SC::PA + 1) >> 1);
    const int nfast = ((a_hi - SC::PA) >> 1) + 1 - tf0;
    int stop = (fast && nfast > 0) ? tf0 : tau_end;
    for (;;) {   // general loop up to the fast range, the fast range, general loop again: ONE copy of the general body
#pragma unroll 1
      for (; tau < stop; ++tau) fwd_pair<K, KIND>(C, S, ba, g, tau);
      if (stop == tau_end) break;
      stop = tau_end;
      const long long step = 2ll * C.pix_pitch * esz, pf = step * S.pd;
      const int a0 = 2 * tau + SC::PA;
      const uint8_t* pA = base + ((long long)a0 * C.pix_pitch + S.xs + 8 * S.lane) * esz;
      const uint8_t* pB = pA - (long long)SC::reach(0) * C.pix_pitch * esz;
      const int pf_last = C.pix_h - 1 - 2 * S.pd;
      // one iteration per trip, rings aged by register moves (rotation index 0): unrolling a whole rotation period
      // makes every slot index a constant and saves the moves, but the loop then outgrows the instruction caches
      // (measured: no_instruction became the first stall reason)
      uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
#ifdef VC2_DWT_SWPIPE
      if constexpr (KIND == SAMPLE_U16BE) { qa = __ldg(reinterpret_cast<const uint4*>(pA)); qb = __ldg(reinterpret_cast<const uint4*>(pB)); }
#endif
#pragma unroll 1
      for (int i = 0; i < nfast; ++i, ++tau) {
        g.shift();
        fwd_fast_iter<K, KIND, W - 1>(C, S, F, g, tau, pA, pB, step, pf, pf_last, qa, qb, i + 1 < nfast);
      }
    }
  } else {
#pragma unroll 1
    for (; tau < tau_end; ++tau) fwd_pair<K, KIND>(C, S, ba, g, tau);
  }
}

// ------------------------------------------------------------------------------------------
// inverse level:  LL, HL, LH, HH  ->  pix (dense plane; cropped / clipped / packed at level 0)
// ------------------------------------------------------------------------------------------
template <int K, int PPL>
__device__ __forceinline__ void inv_fetch_row(const DwtComp& C, const StripCtx& S, const BandAddr& ba, int row, int (&dst)[2 * PPL]) {
  int32_t* coef = C.coef + (long long)S.pic * C.coef_pic_stride;
  const int by = row >> 1, bx0 = ((S.xs >> 1) + PPL * S.lane);
  int lo[PPL], hi[PPL];
  if (row & 1) {
    band_access<PPL, false>(coef, ba, C.base_lh, by, bx0, S.bxmax, lo);
    band_access<PPL, false>(coef, ba, C.base_hh, by, bx0, S.bxmax, hi);
  } else {
    if (C.ll) ll_access<PPL, false>(C.ll + (long long)S.pic * C.ll_pic_stride, C.ll_pitch, by, bx0, S.bxmax, lo);
    else band_access<PPL, false>(coef, ba, C.base_ll, by, bx0, S.bxmax, lo);
    band_access<PPL, false>(coef, ba, C.base_hl, by, bx0, S.bxmax, hi);
  }
#pragma unroll
  for (int a = 0; a < PPL; ++a) { dst[a] = lo[a]; dst[PPL + a] = hi[a]; }
}

template <int PPL>
__device__ __forceinline__ void inv_prefetch_row(const DwtComp& C, const StripCtx& S, const BandAddr& ba, int row) {
  if (row > S.H - 1) return;
  const int32_t* coef = C.coef + (long long)S.pic * C.coef_pic_stride;
  const int by = row >> 1, bx0 = ((S.xs >> 1) + PPL * S.lane);
  if (row & 1) {
    band_prefetch<PPL>(coef, ba, C.base_lh, by, bx0, S.bxmax);
    band_prefetch<PPL>(coef, ba, C.base_hh, by, bx0, S.bxmax);
  } else {
    if (C.ll) { if (bx0 >= 0 && bx0 <= S.bxmax) prefetch_l1(C.ll + (long long)S.pic * C.ll_pic_stride + (long long)by * C.ll_pitch + bx0); }
    else band_prefetch<PPL>(coef, ba, C.base_ll, by, bx0, S.bxmax);
    band_prefetch<PPL>(coef, ba, C.base_hl, by, bx0, S.bxmax);
  }
}

template <int K, int KIND, int PPL>
__device__ __forceinline__ void inv_emit_row(const DwtComp& C, const StripCtx& S, int row, const int (&src)[2 * PPL]) {
  constexpr int V = 2 * PPL, SHIFT = Wavelet<K>::SHIFT;
  if (row < S.y0 || row >= S.y1) return;   // warp uniform
  int e[PPL], o[PPL];
#pragma unroll
  for (int a = 0; a < PPL; ++a) { e[a] = src[a]; o[a] = src[PPL + a]; }
  hsteps<K, -1, PPL>(e, o, S.lane, S.hedge, S.plo, S.phi);
  const int gx = S.xs + V * S.lane;
  if (!S.mine || row >= C.pix_h || gx >= C.pix_w) return;
  int v[V];
#pragma unroll
  for (int a = 0; a < PPL; ++a) { v[2 * a] = e[a]; v[2 * a + 1] = o[a]; }
  store_pix<K, KIND, V>(C, S.pic, row, gx, v);
}

template <int K, int KIND>
__device__ __forceinline__ void inv_pair(const DwtComp& C, const StripCtx& S, const BandAddr& ba, Rings<K, -1>& g, int tau) {
  using SC = Sched<K, -1>;
  constexpr int V = Vec<K>::V, PPL = V / 2, n = SC::n;
  const int H = S.H;
  const int a = 2 * tau + SC::PA;
  inv_prefetch_row<PPL>(C, S, ba, a + 2 * S.pd);
  inv_prefetch_row<PPL>(C, S, ba, a + 2 * S.pd - SC::reach(0));
  g.shift();
  if (a <= H - 1) inv_fetch_row<K, PPL>(C, S, ba, a, g.A[0]);
  const int r0 = a - SC::reach(0);
  if (r0 >= 0 && r0 <= H - 1) inv_fetch_row<K, PPL>(C, S, ba, r0, g.B[0]);
  vstep_guarded<K, -1, 0>(g, a, H);
  vstep_guarded<K, -1, 1>(g, a, H);
  if constexpr (n == 4) {
    vstep_guarded<K, -1, 2>(g, a, H);
    vstep_guarded<K, -1, 3>(g, a, H);
  }
  constexpr int L1 = SC::lag(n - 1), L2 = SC::lag(n - 2);
  {
    const int r = a - L1;
    if (r >= 0 && r <= H - 1) {
      if constexpr (SC::last_targets_A) inv_emit_row<K, KIND, PPL>(C, S, r, g.A[L1 / 2]);
      else inv_emit_row<K, KIND, PPL>(C, S, r, g.B[(L1 - SC::reach(0)) / 2]);
    }
  }
  {
    const int r = a - L2;
    if (r >= 0 && r <= H - 1) {
      if constexpr (!SC::last_targets_A) inv_emit_row<K, KIND, PPL>(C, S, r, g.A[L2 / 2]);
      else inv_emit_row<K, KIND, PPL>(C, S, r, g.B[(L2 - SC::reach(0)) / 2]);
    }
  }
}

// ---- inverse fast loop -------------------------------------------------------------------------------
template <int PPL, bool ODD>
__device__ __forceinline__ void inv_fast_fetch(const FastCtx& F, int row, int (&dst)[2 * PPL]) {
  const int by = row >> 1, i = F.idx(by);
  int4 lo, hi;
  if (ODD) {
    lo = __ldg(reinterpret_cast<const int4*>(F.coefpic + i + F.o_lh));
    hi = __ldg(reinterpret_cast<const int4*>(F.coefpic + i + F.o_hh));
  } else {
    lo = F.llp ? __ldg(reinterpret_cast<const int4*>(F.llp + (long long)by * F.ll_pitch)) : __ldg(reinterpret_cast<const int4*>(F.coefpic + i + F.o_ll));
    hi = __ldg(reinterpret_cast<const int4*>(F.coefpic + i + F.o_hl));
  }
  dst[0] = lo.x; dst[1] = lo.y; dst[2] = lo.z; dst[3] = lo.w;
  dst[PPL] = hi.x; dst[PPL + 1] = hi.y; dst[PPL + 2] = hi.z; dst[PPL + 3] = hi.w;
}
template <bool ODD>
__device__ __forceinline__ void inv_fast_prefetch(const FastCtx& F, int row) {
  const int by = row >> 1, i = F.idx(by);
  if (ODD) { prefetch_l1(F.coefpic + i + F.o_lh); prefetch_l1(F.coefpic + i + F.o_hh); }
  else {
    if (F.llp) prefetch_l1(F.llp + (long long)by * F.ll_pitch); else prefetch_l1(F.coefpic + i + F.o_ll);
    prefetch_l1(F.coefpic + i + F.o_hl);
  }
}

template <int K, int KIND, int PPL>
__device__ __forceinline__ void inv_fast_emit(const DwtComp& C, const StripCtx& S, const FastCtx& F, uint8_t* dst, const int (&src)[2 * PPL]) {
  constexpr int V = 2 * PPL, SHIFT = Wavelet<K>::SHIFT;
  int e[PPL], o[PPL];
#pragma unroll
  for (int a = 0; a < PPL; ++a) { e[a] = src[a]; o[a] = src[PPL + a]; }
  hsteps<K, -1, PPL>(e, o, S.lane, false, 0, 0);
  if (!F.store) return;
  int v[V];
  const int rnd = SHIFT ? (1 << (SHIFT - 1)) : 0;
#pragma unroll
  for (int a = 0; a < PPL; ++a) { v[2 * a] = e[a]; v[2 * a + 1] = o[a]; }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    if (SHIFT) v[k] = (v[k] + rnd) >> SHIFT;
    if (KIND != SAMPLE_I32) v[k] = (int)((unsigned)(min(max(v[k], C.clip_min), C.clip_max) + C.soffset) << C.sshift);
  }
  if (KIND == SAMPLE_I32) {
#pragma unroll
    for (int j = 0; j < V; j += 4) reinterpret_cast<int4*>(dst)[j / 4] = make_int4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else {
    unsigned w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = __byte_perm((unsigned)v[2 * j], (unsigned)v[2 * j + 1], 0x4501);
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// iteration T of a chunk; p1 / p2 run over the output rows a - L1 and a - L2 of this lane
template <int K, int KIND, int T>
__device__ __forceinline__ void inv_fast_iter(const DwtComp& C, const StripCtx& S, const FastCtx& F, Rings<K, -1>& g, int tau,
                                              uint8_t*& p1, uint8_t*& p2, long long step) {
  using SC = Sched<K, -1>;
  constexpr int V = Vec<K>::V, PPL = V / 2, n = SC::n, W = SC::WR, R = (W - 1 - T % W) % W;
  const int a = 2 * tau + SC::PA;
  if (a + 2 * S.pd <= S.H - 1) {
    inv_fast_prefetch<(SC::PA & 1) != 0>(F, a + 2 * S.pd);
    inv_fast_prefetch<((SC::PA - SC::reach(0)) & 1) != 0>(F, a + 2 * S.pd - SC::reach(0));
  }
  inv_fast_fetch<PPL, (SC::PA & 1) != 0>(F, a, g.A[R]);
  inv_fast_fetch<PPL, ((SC::PA - SC::reach(0)) & 1) != 0>(F, a - SC::reach(0), g.B[R]);
  vsteps_rot<K, -1, R>(g);
  constexpr int L1 = SC::lag(n - 1), L2 = SC::lag(n - 2);
  if constexpr (SC::last_targets_A) {
    inv_fast_emit<K, KIND, PPL>(C, S, F, p1, g.A[(L1 / 2 + R) % W]);
    inv_fast_emit<K, KIND, PPL>(C, S, F, p2, g.B[((L2 - SC::reach(0)) / 2 + R) % W]);
  } else {
    inv_fast_emit<K, KIND, PPL>(C, S, F, p1, g.B[((L1 - SC::reach(0)) / 2 + R) % W]);
    inv_fast_emit<K, KIND, PPL>(C, S, F, p2, g.A[(L2 / 2 + R) % W]);
  }
  p1 += step; p2 += step;
}
template <int K, int KIND>
__global__ void __launch_bounds__(32 * WARPS, VC2_DWT_MINB) dwt_inv_kernel(const DwtParams p, int seg_rows) {
  using SC = Sched<K, -1>;
  const int comp = blockIdx.z % p.ncomp;
  const DwtComp& C = p.c[comp];
  StripCtx S;
  S.pic = blockIdx.z / p.ncomp;
  S.pd = p.pd;
  if (!strip_setup<K>(C, seg_rows, S, true)) return;
  S.y1 = min(S.y1, C.pix_h + (C.pix_h & 1));   // rows beyond the crop are never needed
  const BandAddr ba = {C.bh, C.bw, C.lgbh, C.lgbw, C.nx, C.NC >> 2};
  Rings<K, -1> g;
  g.clear();
  const int tau0 = max((S.y0 - SC::total_reach() - SC::PA) >> 1, 0);
  const int tau_end = (S.y1 - 1 + SC::lag(SC::n - 1) - SC::PA + 2) >> 1;
  int tau = tau0;
  if constexpr (Vec<K>::V == 8 && KIND != SAMPLE_U8) {
    constexpr int W = SC::WR, esz = KIND == SAMPLE_I32 ? 4 : 2;
    FastCtx F;
    uint8_t* base = (uint8_t*)C.pix + (long long)S.pic * C.pix_pic_stride * (KIND == SAMPLE_I32 ? 4 : 1);
    bool fast = p.fast && fast_setup<K>(C, S, F) && S.xs + Geo<K>::XW <= C.pix_w && ((C.pix_pitch * esz) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(base) & 15) == 0;
    // newest rows a for which the pair needs no edge rule and emits two rows of [y0, y1) that survive the crop
    const int a_lo = max(SC::AMIN, S.y0 + SC::lag(SC::n - 1));
    const int a_hi = min(S.H - 1, min(S.y1, C.pix_h) - 1 + SC::lag(SC::n - 2));
    const int tf0 = max(tau0, (a_lo - SC::PA + 1) >> 1);
    const int nfast = ((a_hi - SC::PA) >> 1) + 1 - tf0;
    int stop = (fast && nfast > 0) ? tf0 : tau_end;
    for (;;) {
#pragma unroll 1
      for (; tau < stop; ++tau) inv_pair<K, KIND>(C, S, ba, g, tau);
      if (stop == tau_end) break;
      stop = tau_end;
      const long long step = 2ll * C.pix_pitch * esz;
      const int a0 = 2 * tau + SC::PA;
      uint8_t* p1 = base + ((long long)(a0 - SC::lag(SC::n - 1)) * C.pix_pitch + S.xs + 8 * S.lane) * esz;
      uint8_t* p2 = base + ((long long)(a0 - SC::lag(SC::n - 2)) * C.pix_pitch + S.xs + 8 * S.lane) * esz;
#pragma unroll 1
      for (int i = 0; i < nfast; ++i, ++tau) {
        g.shift();
        inv_fast_iter<K, KIND, W - 1>(C, S, F, g, tau, p1, p2, step);
      }
    }
  } else {
#pragma unroll 1
    for (; tau < tau_end; ++tau) inv_pair<K, KIND>(C, S, ba, g, tau);
  }
}

// ------------------------------------------------------------------------------------------
// layout kernels: reference in-place interleaved plane <-> group-interleaved coefficient block
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ long long slice_major_index(const SliceGeom& g, int c, int y, int x) {
  const int d = g.depth;
  const int t = (y | x) & ((1 << d) - 1);
  int b, by, bx;
  if (t == 0) { b = 0; by = y >> d; bx = x >> d; }
  else {
    const int l = __ffs(t) - 1;          // 0 = finest level
    const int L = d - l;                 // VC-2 level number
    const int hx = (x >> l) & 1, hy = (y >> l) & 1;
    b = 3 * (L - 1) + (hx ? (hy ? 3 : 1) : 2);
    by = y >> (l + 1); bx = x >> (l + 1);
  }
  const int bh = g.part_h[c][b], bw = g.part_w[c][b];
  const int sy = by / bh, sx = bx / bw;
  return coef_index(sy * g.slices_x + sx, g.comp_start[c] + g.band_start[c][b] + (by - sy * bh) * bw + (bx - sx * bw), g.comp_start[3] >> 2);
}

__global__ void layout_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ dst, const SliceGeom g, int c, bool to_slice_major) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const PlaneGeom& pg = g.plane[c];
  if (x >= pg.pw || y >= pg.ph) return;
  const long long a = (long long)y * pg.pw + x, b = slice_major_index(g, c, y, x);
  if (to_slice_major) dst[b] = src[a];
  else dst[a] = src[b];
}

template <int K, int KIND>
cudaError_t launch_level(cudaStream_t s, const DwtParams& p, int npictures, int seg_rows) {
  int mw = 0, mh = 0;
  for (int c = 0; c < p.ncomp; ++c) {
    mw = p.c[c].lat_w > mw ? p.c[c].lat_w : mw;
    mh = p.c[c].lat_h > mh ? p.c[c].lat_h : mh;
  }
  const int segs = (mh + seg_rows - 1) / seg_rows;
  const dim3 grid((mw + Geo<K>::XU - 1) / Geo<K>::XU, (segs + WARPS - 1) / WARPS, npictures * p.ncomp);
#if VC2_DWT_PART == 1
  dwt_fwd_kernel<K, KIND><<<grid, 32 * WARPS, 0, s>>>(p, seg_rows);
#else
  dwt_inv_kernel<K, KIND><<<grid, 32 * WARPS, 0, s>>>(p, seg_rows);
#endif
  return cudaGetLastError();
}

template <int KIND>
cudaError_t dispatch(cudaStream_t s, int kernel, const DwtParams& p, int npictures, int seg_rows) {
#define VC2_CASE(K) case K: return launch_level<K, KIND>(s, p, npictures, seg_rows);
  switch (kernel) {
    VC2_CASE(VC2_DD97) VC2_CASE(VC2_LEGALL) VC2_CASE(VC2_DD137) VC2_CASE(VC2_HAAR0)
    VC2_CASE(VC2_HAAR1) VC2_CASE(VC2_FIDELITY) VC2_CASE(VC2_DAUB97)
    default: return cudaErrorInvalidValue;
  }
#undef VC2_CASE
}

}  // namespace

// seg_rows: lattice rows per warp (even); the host picks it so that the grid fills the GPU.
// This file is compiled twice (Makefile): VC2_DWT_PART=1 forward kernels, =2 inverse kernels + layout.
#if VC2_DWT_PART == 1
cudaError_t dwt_fwd_launch(cudaStream_t s, int kernel, int sample_kind, const DwtParams& p, int npictures, int seg_rows) {
#else
cudaError_t dwt_inv_launch(cudaStream_t s, int kernel, int sample_kind, const DwtParams& p, int npictures, int seg_rows) {
#endif
  if (seg_rows < 2 || (seg_rows & 1)) return cudaErrorInvalidValue;
  switch (sample_kind) {
    case SAMPLE_I32: return dispatch<SAMPLE_I32>(s, kernel, p, npictures, seg_rows);
    case SAMPLE_U16BE: return dispatch<SAMPLE_U16BE>(s, kernel, p, npictures, seg_rows);
    case SAMPLE_U8: return dispatch<SAMPLE_U8>(s, kernel, p, npictures, seg_rows);
    default: return cudaErrorInvalidValue;
  }
}

#if VC2_DWT_PART == 2
// src/dst: one in-place plane (ph x pw) and one picture's group-interleaved block
cudaError_t layout_launch(cudaStream_t s, bool to_slice_major, const int32_t* src, int32_t* dst, const SliceGeom& g, int c) {
  const PlaneGeom& pg = g.plane[c];
  const dim3 block(32, 8), grid((pg.pw + 31) / 32, (pg.ph + 7) / 8);
  layout_kernel<<<grid, block, 0, s>>>(src, dst, g, c, to_slice_major);
  return cudaGetLastError();
}

#endif

}  // namespace vc2
