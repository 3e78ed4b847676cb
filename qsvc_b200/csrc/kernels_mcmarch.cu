// kernels_mcmarch.cu -- decorrelate / correlate as ONE line-based pass per temporal level.
//
// Reference pipeline per pair (decorrelate.cpp:732-861, 920-1066): predict() averages the two
// displaced up-sampled references block by block into prediction[c] (luma size << a), the planes
// are clipped, analysed `a` levels (chroma one more) in place with the integer 5/3 lifting
// (dwt2d.cpp:76-119: rows, then columns, per level), and only the LL band meets the odd frame
// (residue) or the high frame (reconstruction).
//
// Here nothing between the up-sampled reference bytes (V planes, k_upsample2x) and the output
// frame touches memory.  One WARP owns a strip of 29 x 8 prediction columns (plus two halo lanes
// on the left and one on the right) and marches down the rows:
//   * each lane fetches its 8 prediction bytes of a row from the two displaced references
//     (unaligned 8-byte window = three aligned words + funnel shift) and averages them with
//     packed-byte arithmetic; blocks whose footprint leaves the picture evaluate the reference's
//     border rule (bordered_ref_u8) per sample; rows below the last whole block come from the
//     chained tail state (k_tail_state);
//   * the row pass of every level exchanges one sample and one high-pass value with the
//     neighbouring lanes by warp shuffles;
//   * the column pass of every level is a streaming lifting step per column held in registers
//     (state: last even sample, last odd sample, last high-pass value); level k+1 consumes the
//     rows level k emits, so the whole multi-level LL analysis is a register pipeline;
//   * the LL rows that fall into the warp's row segment are turned into residue / reconstruction
//     bytes (and histograms) on the spot.
// No shared memory (except the optional histograms), no block barriers, no intermediate planes.
#include <cstdlib>

#include "kernels.cuh"

#define COUNT(L) (++*(L).counter)
#define FULL 0xffffffffu

static constexpr int MARCH_HL = 2;   // halo lanes on the left  (16 columns >= 14 = reach of 3 levels)
static constexpr int MARCH_UL = 29;
static constexpr int MARCH_PF = 6;   // L2 prefetch distance in steps (two prediction rows each)  // lanes that own outputs  (232 columns per warp)

__device__ __forceinline__ int tq2(int v) { return (v + (int)((unsigned)v >> 31)) >> 1; }  // C "/ 2"
__device__ __forceinline__ int tq4(int v) { return (v + (int)((unsigned)v >> 30)) >> 2; }  // C "/ 4", |v| < 2^30

__device__ __forceinline__ int bref_u8(const uint8_t *U, int pitch, int Yd, int Xd, int b, int padh, int y,
                                       int x) {
  // closed form of texture::alloc + fill_border (common.cuh bordered_ref) on a byte plane
  if ((unsigned)y < (unsigned)Yd && (unsigned)x < (unsigned)Xd) return U[(long long)y * pitch + x];
  if (b > padh) {
    int y0 = Yd - b;
    if (y0 != 0) {
      if (y == y0 && x < -padh) return U[(long long)iclamp(y0 - 1, 0, Yd - 1) * pitch + Xd - 1];
    } else if (y == -1 && x >= Xd + padh) {
      return U[0];
    }
  }
  if (y >= Yd && x < 0) return U[(long long)(Yd - 1) * pitch + Xd - 1];
  return U[(long long)iclamp(y, 0, Yd - 1) * pitch + iclamp(x, 0, Xd - 1)];
}

// Streaming 5/3 analysis of one line of 2*half samples (5_3.cpp:39-52).  Step k consumes
// x[2k], x[2k+1] and returns l[k-1]; step k == half is virtual and flushes l[half-1]
// (h[half-1] = x[n-1] - x[n-2] is the generic formula with x[n] := x[n-2]); l[0] = x[0] + h[0]/2
// is the generic formula with h[-1] := h[0].  SP = false: the caller knows 1 < k < half.
struct VState {
  int e, o, hp;
};
template <bool SP>
__device__ __forceinline__ int vstep(VState &s, int k, int half, int xe, int xo) {
  if (SP && k == half) xe = s.e;
  const int h = s.o - tq2(s.e + xe);
  const int l = s.e + tq4(h + ((SP && k == 1) ? h : s.hp));
  s.hp = h;
  s.e = xe;
  s.o = xo;
  return l;
}

// Row pass of one level across the warp: every lane holds N consecutive samples of the row and
// produces the N/2 low-pass samples of its columns.
template <int N>
__device__ __forceinline__ void hpass(const int *v, int *out, bool first, bool last) {
  int nxt = __shfl_down_sync(FULL, v[0], 1);
  if (last) nxt = v[N - 2];
  int h[N / 2];
#pragma unroll
  for (int i = 0; i < N / 2; i++) h[i] = v[2 * i + 1] - tq2(v[2 * i] + (i + 1 < N / 2 ? v[2 * i + 2] : nxt));
  int hp = __shfl_up_sync(FULL, h[N / 2 - 1], 1);
  if (first) hp = h[0];
  out[0] = v[0] + tq4(h[0] + hp);
#pragma unroll
  for (int i = 1; i < N / 2; i++) out[i] = v[2 * i] + tq4(h[i] + h[i - 1]);
}

// Level-0 row pass on the 8 prediction bytes of a lane (s[0..7] = lo, hi; s[-2], s[-1] from the lane
// on the left, s[8] from the lane on the right).  The kernel is bound by the ALU pipe (LOP3 / SHF /
// PRMT / IADD3 / LEA.HI share it, profiles/r2_pipe_probe.txt), so the byte sums go through IDP.4A on
// the FMA pipe: for output i the window U = (s[2i-1], s[2i], s[2i+1], s[2i+2]) gives the halved even
// pair sum A_i, the odd pair sum d_{i-1} + d_i and the centre sample:
//   l_i = s[2i] + (h_{i-1} + h_i) / 4,  h_{i-1} + h_i = d_{i-1} + d_i - A_{i-1} - A_i.
// The line ends are mirror extensions: x[n] := x[n-2] on the right, and h[-1] := h[0] on the left is
// s[-1] := s[1], s[-2] := s[2].
__device__ __forceinline__ int dp4(unsigned a, unsigned w, int acc) { return (int)__dp4a(a, w, (unsigned)acc); }
__device__ __forceinline__ void hpass_u8(unsigned lo, unsigned hi, int *out, bool first, bool last) {
  unsigned nlo = __shfl_down_sync(FULL, lo, 1), phi = __shfl_up_sync(FULL, hi, 1);
  if (last) nlo = hi >> 16;                      // s[8] := s[6]
  if (first) phi = __byte_perm(lo, 0, 0x1200);   // s[-2] := s[2], s[-1] := s[1]
  const unsigned U[4] = {__byte_perm(phi, lo, 0x6543), __byte_perm(lo, hi, 0x4321), __byte_perm(lo, hi, 0x6543),
                         __byte_perm(hi, nlo, 0x4321)};
  int ap = dp4(__byte_perm(phi, lo, 0x5432), 0x00010001u, 0) >> 1;  // A_{-1} = (s[-2] + s[0]) >> 1
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int a = dp4(U[i], 0x01000100u, 0) >> 1;
    const int t = dp4(U[i], 0x00010001u, 0) - a - ap;
    out[i] = dp4(U[i], 0x00000100u, tq4(t));
    ap = a;
  }
}

// Raw words of the two prediction rows of one step (both rows lie in the same block row).
struct RawPair {
  unsigned a[6], b[6];  // [row][word]: three aligned words per row and reference
  int sa, sb;           // funnel shifts (bits) that align them
};

// Eight samples V(y, x0 .. x0+7) of a bordered reference plane whose window leaves the picture
// on the left or right (rare: edge blocks with outward vectors).
__device__ __noinline__ uint2 bref_row8(const uint8_t *U, int pitch, int Yd, int Xd, int b, int padh, int y,
                                        int x0) {
  unsigned lo = 0, hi = 0;
  for (int k = 0; k < 4; k++) {
    lo |= (unsigned)bref_u8(U, pitch, Yd, Xd, b, padh, y, x0 + k) << (8 * k);
    hi |= (unsigned)bref_u8(U, pitch, Yd, Xd, b, padh, y, x0 + 4 + k) << (8 * k);
  }
  return make_uint2(lo, hi);
}

// EXTRA: histograms and / or the prediction side output are requested
template <int NLEV, bool EXTRA>
struct Marcher {
  const MarchParams &q;
  int pair, c, x, xs, og0, og1, OW;
  bool in_pic, first, last, owner, do_hist;
  const uint8_t *in;
  uint8_t *out;
  int *h_pred, *h_res;
  int is_I;
  unsigned in_pre;  // input bytes of the next output row, loaded one output row ahead
  // plane pointers are recomputed where the (rare) general paths need them
  __device__ __forceinline__ const uint8_t *V0() const {
    return q.v + ((long long)(q.f0 + pair) * 3 + c) * q.v_plane_stride;
  }
  __device__ __forceinline__ const uint8_t *V1() const { return V0() + 3 * q.v_plane_stride; }
  __device__ __forceinline__ const uint8_t *TL() const {
    return q.tail + ((long long)pair * 3 + c) * q.tail_plane_stride;
  }
  __device__ __forceinline__ uint8_t *pout() const {
    if (!EXTRA || !q.prediction) return nullptr;
    const long long coff = c == 0 ? 0 : (long long)q.X * q.Y + (long long)(c - 1) * (q.X / 2) * (q.Y / 2);
    return q.prediction + (long long)pair * q.pred_stride + coff;
  }
  // block-row cache
  int cur_by, my0, my1, col0, col1, sa, sb;
  bool xin0, xin1, fastrow;
  const unsigned *pa, *pb;  // fastrow: word pointers of prediction row 0 in the two references
  // lifting pipeline
  int half1, half2, half3, t_fetch;
  VState s1[4], s2[2], s3[1];
  int st1[2], st2[1];
  RawPair nx;

  __device__ __forceinline__ Marcher(const MarchParams &qq) : q(qq) {}

  __device__ __forceinline__ void new_block_row(int by) {
    cur_by = by;
    const int plane = q.BY * q.BX;
    const short *m = q.mv + (long long)pair * 4 * plane + by * q.BX + (xs >> q.bs_shift);
    int mx0 = 0, mx1 = 0;
    my0 = my1 = 0;
    if (in_pic) {  // lanes outside the picture read column 0 undisplaced (values never used)
      mx0 = __ldg(m + MV_PREV_X * plane);
      my0 = __ldg(m + MV_PREV_Y * plane);
      mx1 = __ldg(m + MV_NEXT_X * plane);
      my1 = __ldg(m + MV_NEXT_Y * plane);
    }
    col0 = xs + mx0;
    col1 = xs + mx1;
    // the planes carry a materialised border ring of q.ring samples (the reference's border rule,
    // k_fill_ring): windows that stay inside it are plain loads
    xin0 = col0 >= -q.ring && col0 + 8 <= q.Xa + q.ring;
    xin1 = col1 >= -q.ring && col1 + 8 <= q.Xa + q.ring;
    const int y0 = by << q.bs_shift, y1 = min(y0 + q.bsa, q.cy) - 1;
    const bool noclamp = y0 + my0 >= -q.ring && y1 + my0 < q.Ya + q.ring && y0 + my1 >= -q.ring &&
                         y1 + my1 < q.Ya + q.ring;
    fastrow = __all_sync(FULL, xin0 && xin1 && noclamp);
    const int o0 = my0 * q.v_pitch + col0, o1 = my1 * q.v_pitch + col1;  // pitch % 4 == 0
    sa = 8 * (o0 & 3);
    sb = 8 * (o1 & 3);
    pa = reinterpret_cast<const unsigned *>(V0() + (o0 & ~3));
    pb = reinterpret_cast<const unsigned *>(V1() + (o1 & ~3));
  }

  // One row, any case.  Columns inside the picture: V(y, x) = U[clamp(y)][x] (the border quirks
  // of texture::fill_border need x < 0 or x >= Xd); otherwise the closed-form border rule.
  __device__ __forceinline__ void fetch_row_general(int r, unsigned *a, unsigned *b) {
    // columns inside the ring: rows beyond it repeat its outermost row (the border rule is constant
    // along y there); otherwise the closed-form rule per sample
    if (xin0) {
      const int off = min(max(r + my0, -q.ring), q.Ya + q.ring - 1) * q.v_pitch + col0;
      const unsigned *a4 = reinterpret_cast<const unsigned *>(V0() + (off & ~3));
      a[0] = __ldg(a4);
      a[1] = __ldg(a4 + 1);
      a[2] = __ldg(a4 + 2);
    } else {
      const uint2 t = bref_row8(V0(), q.v_pitch, q.Ya, q.Xa, q.ba, q.padh, r + my0, col0);
      a[0] = t.x;
      a[1] = t.y;
      a[2] = 0;
    }
    if (xin1) {
      const int off = min(max(r + my1, -q.ring), q.Ya + q.ring - 1) * q.v_pitch + col1;
      const unsigned *b4 = reinterpret_cast<const unsigned *>(V1() + (off & ~3));
      b[0] = __ldg(b4);
      b[1] = __ldg(b4 + 1);
      b[2] = __ldg(b4 + 2);
    } else {
      const uint2 t = bref_row8(V1(), q.v_pitch, q.Ya, q.Xa, q.ba, q.padh, r + my1, col1);
      b[0] = t.x;
      b[1] = t.y;
      b[2] = 0;
    }
  }

  // fastrow: every lane's windows of this block row lie inside the picture
  template <bool PF>
  __device__ __forceinline__ void fetch2_fast(int r, RawPair &w) {
    const unsigned pw = (unsigned)q.v_pitch >> 2;
    const unsigned *a4 = pa + r * pw, *b4 = pb + r * pw;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      w.a[k] = __ldg(a4 + k);
      w.a[3 + k] = __ldg(a4 + pw + k);
      w.b[k] = __ldg(b4 + k);
      w.b[3 + k] = __ldg(b4 + pw + k);
    }
    // pull the rows MARCH_PF steps further down into L2 (same vectors assumed: a hint only;
    // the planes have Ya + 2 rows and the caller stays MARCH_PF steps away from the last row)
    if (PF) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a4 + 2 * MARCH_PF * pw));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a4 + (2 * MARCH_PF + 1) * pw));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(b4 + 2 * MARCH_PF * pw));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(b4 + (2 * MARCH_PF + 1) * pw));
    }
  }

  // Raw words of prediction rows r, r + 1 (r even) for this lane's 8 columns.
  __device__ __forceinline__ void fetch2(int r, RawPair &w) {
    if (r >= q.cy) {  // rows below the last whole block: chained state (A.2.6)
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const uint2 t = __ldg(reinterpret_cast<const uint2 *>(TL() + (unsigned)((r + k) * q.v_pitch + xs)));
        w.a[3 * k] = w.b[3 * k] = t.x;
        w.a[3 * k + 1] = w.b[3 * k + 1] = t.y;
        w.a[3 * k + 2] = w.b[3 * k + 2] = 0;
      }
      w.sa = w.sb = 0;
      fastrow = false;
      return;
    }
    const int by = r >> q.bs_shift;
    if (by != cur_by) new_block_row(by);
    if (fastrow) {
      fetch2_fast<false>(r, w);
      w.sa = sa;
      w.sb = sb;
    } else {
      fetch_row_general(r, w.a, w.b);
      fetch_row_general(r + 1, w.a + 3, w.b + 3);
      w.sa = xin0 ? sa : 0;
      w.sb = xin1 ? sb : 0;
    }
  }

  // (r0 + r1) / 2 on bytes; the [0,255] clip of decorrelate.cpp:841-848 is a no-op
  static __device__ __forceinline__ void combine(const RawPair &w, int sa_, int sb_, int k, unsigned &lo,
                                                 unsigned &hi) {
    lo = __vhaddu4(__funnelshift_r(w.a[3 * k], w.a[3 * k + 1], sa_), __funnelshift_r(w.b[3 * k], w.b[3 * k + 1], sb_));
    hi = __vhaddu4(__funnelshift_r(w.a[3 * k + 1], w.a[3 * k + 2], sa_),
                   __funnelshift_r(w.b[3 * k + 1], w.b[3 * k + 2], sb_));
  }

  // the NOUT <= 4 input bytes of this lane in output row e
  template <int NOUT>
  __device__ __forceinline__ unsigned load_in(int e) const {
    if (!owner) return 0;
    const uint8_t *p = in + (unsigned)(e * OW + (x >> NLEV));  // a component plane is far below 4 GB
    if (NOUT == 1) return *p;
    if (NOUT == 2) return *reinterpret_cast<const unsigned short *>(p);
    return *reinterpret_cast<const unsigned *>(p);
  }

  // NOUT = 8 >> NLEV consecutive LL samples of output row e (component resolution)
  template <int NOUT>
  __device__ __forceinline__ void out_row(int e, const int *p) {
    if (e < og0 || e >= og1) return;
    unsigned sw[2] = {0, 0}, ow[2] = {0, 0}, pw[2] = {0, 0};
    if (NOUT == 8) {
      if (owner) {
        const uint2 t = *reinterpret_cast<const uint2 *>(in + (unsigned)(e * OW + x));
        sw[0] = t.x;
        sw[1] = t.y;
      }
    } else {
      // output rows arrive in order from og0 (run() loaded that one): the next row's input is in flight
      // until it is needed, and no load sits between the scoreboard wait and this row's bytes
      sw[0] = in_pre;
      if (e + 1 < og1) in_pre = load_in<NOUT>(e + 1);
    }
    if (!owner) return;
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
      const int s = (sw[k >> 2] >> (8 * (k & 3))) & 0xff;
      int o;
      if (!q.synth) {
        int rr = s - p[k];
        rr = rr < -128 ? -128 : (rr > 127 ? 127 : rr);
        o = rr + 128;
        if (EXTRA && do_hist) {
          atomicAdd(&h_pred[s], 1);
          atomicAdd(&h_res[o], 1);
        }
      } else if (is_I) {
        o = s;
      } else {
        o = s - 128 + p[k];
        o = o < 0 ? 0 : (o > 255 ? 255 : o);
      }
      ow[k >> 2] |= (unsigned)o << (8 * (k & 3));
      pw[k >> 2] |= (unsigned)(p[k] & 0xff) << (8 * (k & 3));
    }
    const unsigned idx = (unsigned)(e * OW + (x >> NLEV));
    uint8_t *po = pout();
    if (NOUT == 1) {
      out[idx] = (uint8_t)ow[0];
      if (po) po[idx] = (uint8_t)pw[0];
    } else if (NOUT == 2) {
      *reinterpret_cast<unsigned short *>(out + idx) = (unsigned short)ow[0];
      if (po) *reinterpret_cast<unsigned short *>(po + idx) = (unsigned short)pw[0];
    } else if (NOUT == 4) {
      *reinterpret_cast<unsigned *>(out + idx) = ow[0];
      if (po) *reinterpret_cast<unsigned *>(po + idx) = pw[0];
    } else {
      *reinterpret_cast<uint2 *>(out + idx) = make_uint2(ow[0], ow[1]);
      if (po) *reinterpret_cast<uint2 *>(po + idx) = make_uint2(pw[0], pw[1]);
    }
  }

  // One step t of the pipeline: prediction rows 2t, 2t+1 enter level 1; every level emits the
  // row it completes into the next one.  SP = true handles the first outputs of a line and the
  // virtual steps at its end; PH = t & 3 when known at compile time (-1 otherwise); FAST_FETCH:
  // the rows of step t + 1 lie in the current block row and that block row is `fastrow`.
  template <bool SP, int PH, bool FAST_FETCH>
  __device__ __forceinline__ void step(int t) {
    int xe[4] = {0, 0, 0, 0}, xo[4] = {0, 0, 0, 0};
    if (!SP || t < half1) {
      unsigned lo0, hi0, lo1, hi1;
      combine(nx, nx.sa, nx.sb, 0, lo0, hi0);
      combine(nx, nx.sa, nx.sb, 1, lo1, hi1);
      // rows of the next step are in flight while this one is computed
      if (FAST_FETCH) fetch2_fast<true>(2 * t + 2, nx);
      else if (!SP || t < t_fetch) fetch2(2 * t + 2, nx);
      hpass_u8(lo0, hi0, xe, first, last);
      hpass_u8(lo1, hi1, xo, first, last);
    }
    int a[4];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = vstep<SP>(s1[i], t, half1, xe[i], xo[i]);
    const int e1 = t - 1;  // LL1 row just completed
    if (NLEV == 1) {
      out_row<4>(e1, a);
      return;
    }
    const bool e1_odd = PH >= 0 ? !(PH & 1) : (e1 & 1);
    if (!e1_odd) {  // even row of the level-1 image: keep its row-passed form
      hpass<4>(a, st1, first, last);
      return;
    }
    int x2o[2];
    hpass<4>(a, x2o, first, last);
    const int j = (e1 - 1) >> 1;
#pragma unroll 1
    for (int rep = 0; rep < (SP ? 2 : 1); rep++) {  // rep 1: virtual step after the last real one
      if (rep == 1 && j != half2 - 1) break;
      const int jj = j + rep;
      int b[2];
#pragma unroll
      for (int i = 0; i < 2; i++) b[i] = vstep<SP>(s2[i], jj, half2, st1[i], x2o[i]);
      const int e2 = jj - 1;  // LL2 row just completed
      if (NLEV == 2) {
        out_row<2>(e2, b);
        continue;
      }
      const bool e2_odd = (PH >= 0 && !SP) ? (PH == 2) : (e2 & 1);
      if (!e2_odd) {
        hpass<2>(b, st2, first, last);
        continue;
      }
      int x3o[1];
      hpass<2>(b, x3o, first, last);
      const int i3 = (e2 - 1) >> 1;
#pragma unroll 1
      for (int rep3 = 0; rep3 < (SP ? 2 : 1); rep3++) {
        if (rep3 == 1 && i3 != half3 - 1) break;
        const int ii = i3 + rep3;
        int cc[1];
        cc[0] = vstep<SP>(s3[0], ii, half3, st2[0], x3o[0]);
        out_row<1>(ii - 1, cc);
      }
    }
  }

  __device__ void run() {
    half1 = q.Ya >> 1;
    half2 = q.Ya >> 2;
    half3 = q.Ya >> 3;
    cur_by = -1;
    in_pre = 0;
    my0 = my1 = col0 = col1 = sa = sb = 0;
    xin0 = xin1 = true;
    fastrow = false;
    pa = pb = nullptr;
    if (NLEV == 0) {
      for (int r = og0; r < og1; r++) {
        unsigned lo, hi;
        if (r >= q.cy) {
          const uint2 t = __ldg(reinterpret_cast<const uint2 *>(TL() + (unsigned)(r * q.v_pitch + xs)));
          lo = t.x;
          hi = t.y;
        } else {
          const int by = r >> q.bs_shift;
          if (by != cur_by) new_block_row(by);
          fetch_row_general(r, nx.a, nx.b);
          nx.a[3] = nx.a[4] = nx.a[5] = nx.b[3] = nx.b[4] = nx.b[5] = 0;
          nx.sa = xin0 ? sa : 0;
          nx.sb = xin1 ? sb : 0;
          combine(nx, nx.sa, nx.sb, 0, lo, hi);
        }
        int p[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          p[k] = (lo >> (8 * k)) & 0xff;
          p[4 + k] = (hi >> (8 * k)) & 0xff;
        }
        out_row<8>(r, p);
      }
      return;
    }
    in_pre = load_in<(8 >> NLEV)>(og0);
    int t0, t_last;
    if (NLEV == 1) {
      t0 = max(0, og0 - 1);
      t_last = og1;
    } else if (NLEV == 2) {
      t0 = max(0, (og0 - 2) * 2);
      t_last = 2 * (og1 - 1) + 4;
    } else {
      t0 = max(0, (og0 - 2) * 4);
      t_last = 4 * (og1 - 1) + 10;
    }
    t_last = min(t_last, half1);
    t_fetch = min(t_last, half1 - 1);  // last step that reads rows
#pragma unroll
    for (int i = 0; i < 4; i++) s1[i].e = s1[i].o = s1[i].hp = 0;
    s2[0] = s2[1] = s3[0] = s1[0];
    st1[0] = st1[1] = st2[0] = 0;
    fetch2(2 * t0, nx);
    // Steps 0..10 of a line produce the first output of some level (SP).  Groups of four
    // regular steps starting at a multiple of four run unrolled: their rows 8u .. 8u+7 lie in one
    // block row (block sizes are multiples of 8 rows), only the fetch for step 4u+4 can enter a
    // new one.
    const bool can4 = q.bs_shift >= 3;
#pragma unroll 1
    for (int t = t0; t <= t_last;) {
      if (can4 && fastrow && !(t & 3) && t > 10 && t + 3 + MARCH_PF < half1) {
        step<false, 0, true>(t);
        step<false, 1, true>(t + 1);
        step<false, 2, true>(t + 2);
        step<false, 3, false>(t + 3);
        t += 4;
      } else {
        step<true, -1, false>(t);
        t++;
      }
    }
  }
};

template <int NLEV, bool EXTRA>
__device__ __forceinline__ void mc_march(const MarchParams &q, int pair, int c, int strip, int seg, int *h_pred,
                                         int *h_res, bool do_hist) {
  Marcher<NLEV, EXTRA> m(q);
  const int lane = threadIdx.x & 31;
  m.pair = pair;
  m.c = c;
  m.x = (strip * MARCH_UL - MARCH_HL + lane) * 8;
  m.in_pic = m.x >= 0 && m.x < q.Xa;
  m.xs = m.in_pic ? m.x : 0;
  m.first = m.x == 0;
  m.last = m.x + 8 == q.Xa;
  m.owner = m.in_pic && lane >= MARCH_HL && lane < MARCH_HL + MARCH_UL;
  m.do_hist = do_hist;
  m.OW = c ? q.X >> 1 : q.X;
  const int OH = q.Ya >> NLEV;
  m.og0 = (int)(((long long)seg * q.seg_p) >> NLEV);
  m.og1 = min(OH, (int)(((long long)(seg + 1) * q.seg_p) >> NLEV));
  if (m.og0 >= m.og1) return;
  const long long coff = c == 0 ? 0 : (long long)q.X * q.Y + (long long)(c - 1) * (q.X / 2) * (q.Y / 2);
  m.in = q.in + (long long)pair * q.in_stride + coff;
  m.out = q.out + (long long)pair * q.out_stride + coff;
  m.h_pred = h_pred;
  m.h_res = h_res;
  m.is_I = q.synth && q.types[pair] == 'I';
  m.run();
}

// 128 threads: one warp = one (strip, segment) of component c0 + .. % nc (luma and chroma have
// different level counts: two launches).  A (pair, component) plane takes nb CTAs in
// segment-major order.  Block order inside a group of G consecutive pairs (blockIdx.z = group):
// chunks of KC CTAs (about one segment) -> pair of the group -> component -> CTA of the chunk, so
// that the even frame shared by pairs i-1 (NEXT) and i (PREV) is fetched from DRAM once.
template <int NLEV, bool EXTRA>
__global__ void __launch_bounds__(128, 6) k_mc_march(MarchParams q, int c0, int nc, int npairs, int G, int KC,
                                                     int nb) {
  __shared__ int h_pred[EXTRA ? 256 : 1], h_res[EXTRA ? 256 : 1];
  const int per_chunk = G * nc * KC;
  const int chunk = blockIdx.x / per_chunk, rem = blockIdx.x - chunk * per_chunk;
  const int pgc = rem / KC, k = chunk * KC + (rem - pgc * KC);
  const int pair = blockIdx.z * G + pgc / nc, c = c0 + pgc % nc;
  if (pair >= npairs || k >= nb) return;
  const bool do_hist = EXTRA && q.hist && c == 0 && !q.synth;
  if (do_hist) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h_pred[i] = h_res[i] = 0;
    __syncthreads();
  }
  const int idx = k * 4 + (threadIdx.x >> 5);
  const int strip = idx % q.nstrips, seg = idx / q.nstrips;
  if (seg < q.nsegs) mc_march<NLEV, EXTRA>(q, pair, c, strip, seg, h_pred, h_res, do_hist);
  if (do_hist) {
    __syncthreads();
    int *hist = q.hist + (long long)pair * q.hist_stride;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
      if (h_pred[i]) atomicAdd(&hist[i], h_pred[i]);
      if (h_res[i]) atomicAdd(&hist[256 + i], h_res[i]);
    }
  }
}

template <int NLEV>
static void launch_march_n(const Launch &L, const MarchParams &q, int npairs, int c0, int nc) {
  const int nb = (q.nstrips * q.nsegs + 3) / 4, KC = (q.nstrips + 3) / 4;
  const int G = npairs < 4 ? npairs : 4;
  const int nchunks = (nb + KC - 1) / KC;
  dim3 grid(G * nc * KC * nchunks, 1, (npairs + G - 1) / G);
  ProfScope ps_(L, KC_RESIDUE);
  if (q.hist || q.prediction) k_mc_march<NLEV, true><<<grid, 128, 0, L.stream>>>(q, c0, nc, npairs, G, KC, nb);
  else k_mc_march<NLEV, false><<<grid, 128, 0, L.stream>>>(q, c0, nc, npairs, G, KC, nb);
  COUNT(L);
}

void launch_mc_march(const Launch &L, MarchParams q, int npairs) {
  if (npairs <= 0) return;
  q.bs_shift = 0;
  while ((1 << q.bs_shift) < q.bsa) q.bs_shift++;
  q.nstrips = (q.Xa + 8 * MARCH_UL - 1) / (8 * MARCH_UL);
  // enough warps to fill the machine several times over (measured with 24 warps per SM resident: flat between
  // 5 and 12 rounds of 20 warps per SM, 1 % worse at 24 where the warm-up rows of the extra segments cost more),
  // segments as tall as that allows
  const long long cols = (long long)npairs * 3 * q.nstrips;
  static const int seg_waves = getenv("QSVC_MARCH_WAVES") ? atoi(getenv("QSVC_MARCH_WAVES")) : 8;
  long long want = (148LL * 20 * seg_waves + cols - 1) / cols;
  int seg_p = (int)((q.Ya + want - 1) / want);
  seg_p = (seg_p + 7) & ~7;
  if (seg_p < 64) seg_p = 64;
  if (seg_p > 512) seg_p = 512;
  q.seg_p = seg_p;
  q.nsegs = (q.Ya + seg_p - 1) / seg_p;
  switch (q.a) {
    case 0: launch_march_n<0>(L, q, npairs, 0, 1); launch_march_n<1>(L, q, npairs, 1, 2); break;
    case 1: launch_march_n<1>(L, q, npairs, 0, 1); launch_march_n<2>(L, q, npairs, 1, 2); break;
    default: launch_march_n<2>(L, q, npairs, 0, 1); launch_march_n<3>(L, q, npairs, 1, 2); break;
  }
}
