// kernels_mctile.cu -- decorrelate / correlate as a banded shared-memory pipeline.
//
// Reference pipeline per pair (decorrelate.cpp:732-861, 920-1066): predict() averages the two
// displaced up-sampled references block by block into prediction[c] (luma size << a), the planes
// are clipped, analysed N levels in place with the integer 5/3 lifting (dwt2d.cpp:76-119: rows,
// then columns, per level; N = a for luma, a + 1 for chroma) and only the LL band meets the odd
// frame (residue) or the high frame (reconstruction).
//
// One CTA owns a tile of 480 prediction columns (+ a ring of 16 on either side: the reach of
// three analysis levels is 14) of one (pair, component) and marches down a segment of rows in
// bands of 16 prediction rows.  Per band:
//   G   the 16 x 512 prediction bytes: two displaced 8-byte windows of the reference planes per
//       thread and row, averaged as packed bytes; blocks whose footprint leaves the picture take
//       the reference's border rule per sample; rows below the last whole block come from the
//       chained tail state (k_tail_state);
//   R1  row pass of level 1 on the bytes (one warp per row, 8 outputs per lane, byte sums on the
//       FMA pipe with IDP.4A) -> 256 int32 low-pass columns per row;
//   C1  column pass of level 1: one thread per column streams the 16 rows through the lifting
//       step (state in registers across bands) and emits the 8 LL1 rows it completes;
//   R2 / C2 / R3 / C3  the same on the LL rows of the level above (128 / 64 columns);
//   OUT the LL rows of the last level become residue / reconstruction bytes (and histograms,
//       prediction side output) with 32-bit accesses.
// G, R1, C1 and R2 use the whole CTA and are separated by block barriers; from C2 on only the
// first four warps have work and meet on a named barrier while the other four already fetch
// the next band.  Nothing but the two reference planes, the input frame and the output frame
// touches HBM.
#include <cstdlib>

#include "kernels.cuh"

#define COUNT(L) (++*(L).counter)
#define FULL 0xffffffffu

namespace {

constexpr int TWP = 512;            // prediction columns per tile, ring included
constexpr int RING = 16;            // ring on either side (>= 14 = reach of three levels; multiple of 8)
constexpr int TW = TWP - 2 * RING;  // columns a tile produces output for
constexpr int RB = 16;              // prediction rows per band
constexpr int P_PITCH = 528;        // bytes per prediction row in shared memory
constexpr int NT = 256;

__device__ __forceinline__ int tq2(int v) { return (v + (int)((unsigned)v >> 31)) >> 1; }  // C "/ 2"
__device__ __forceinline__ int tq4(int v) { return (v + (int)((unsigned)v >> 30)) >> 2; }  // C "/ 4", |v| < 2^30
__device__ __forceinline__ int dp4(unsigned a, unsigned w, int acc) { return (int)__dp4a(a, w, (unsigned)acc); }

__device__ __forceinline__ int bref8(const uint8_t *U, int pitch, int Yd, int Xd, int b, int padh, int y, int x) {
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

// Eight samples V(y, x0 .. x0+7) of a bordered reference plane, any position.
__device__ __noinline__ uint2 bref_row8(const uint8_t *U, int pitch, int Yd, int Xd, int b, int padh, int y, int x0) {
  if (x0 >= 0 && x0 + 8 <= Xd) {  // columns inside: V(y, x) = U[clamp(y)][x]
    const uint8_t *p = U + (long long)min(max(y, 0), Yd - 1) * pitch + x0;
    const unsigned *p4 = reinterpret_cast<const unsigned *>((uintptr_t)p & ~(uintptr_t)3);
    const int s = 8 * (int)((uintptr_t)p & 3);
    const unsigned w0 = __ldg(p4), w1 = __ldg(p4 + 1), w2 = __ldg(p4 + 2);
    return make_uint2(__funnelshift_r(w0, w1, s), __funnelshift_r(w1, w2, s));
  }
  unsigned lo = 0, hi = 0;
  for (int k = 0; k < 4; k++) {
    lo |= (unsigned)bref8(U, pitch, Yd, Xd, b, padh, y, x0 + k) << (8 * k);
    hi |= (unsigned)bref8(U, pitch, Yd, Xd, b, padh, y, x0 + 4 + k) << (8 * k);
  }
  return make_uint2(lo, hi);
}

// Streaming 5/3 analysis of one column (5_3.cpp:39-52).  Step k consumes x[2k], x[2k+1] and
// returns l[k-1]; step k == half is virtual and flushes l[half-1] (h[half-1] = x[n-1] - x[n-2] is
// the generic formula with x[n] := x[n-2]); l[0] = x[0] + h[0]/2 is the generic formula with
// h[-1] := h[0].
struct VState {
  int e, o, hp;
};
__device__ __forceinline__ int vstep(VState &s, int xe, int xo) {
  const int h = s.o - tq2(s.e + xe);
  const int l = s.e + tq4(h + s.hp);
  s.hp = h;
  s.e = xe;
  s.o = xo;
  return l;
}
__device__ __forceinline__ int vstep_sp(VState &s, bool first, bool flush, int xe, int xo) {
  if (flush) xe = s.e;
  const int h = s.o - tq2(s.e + xe);
  const int l = s.e + tq4(h + (first ? h : s.hp));
  s.hp = h;
  s.e = xe;
  s.o = xo;
  return l;
}

// Row pass: OUTS consecutive low-pass samples from v[0 .. 2*OUTS+2] = x[2c0-2 .. 2c0+2*OUTS].
// BORDER: output `ifirst` is column 0 of the line (h[-1] := h[0]), output `ilast` is its last
// column (x[n] := x[n-2]); -1 / -2 when the run does not contain them.
template <int OUTS, bool BORDER>
__device__ __forceinline__ void rowpass(const int *v, int *out, int ifirst, int ilast) {
  int h[OUTS + 1];  // h[i] = hh(c0 - 1 + i)
#pragma unroll
  for (int i = 0; i <= OUTS; i++) {
    int nx = v[2 * i + 2];
    if (BORDER && i - 1 == ilast) nx = v[2 * i];
    h[i] = v[2 * i + 1] - tq2(v[2 * i] + nx);
  }
#pragma unroll
  for (int i = 0; i < OUTS; i++) {
    int hp = h[i];
    if (BORDER && i == ifirst) hp = h[i + 1];
    out[i] = v[2 * i + 2] + tq4(h[i + 1] + hp);
  }
}

// Level-1 row pass on bytes, away from the line ends: eight outputs from the 16 bytes in w
// (s[0..15]) plus s[-2], s[-1] (top of wp) and s[16] (bottom of wn).  For output i the window
// U = (s[2i-1], s[2i], s[2i+1], s[2i+2]) feeds three byte dot products: the even pair sum (halved:
// A_i), the odd pair sum and the centre sample; l_i = s[2i] + (d_{i-1} + d_i - A_{i-1} - A_i) / 4.
__device__ __forceinline__ void rowpass_bytes(unsigned wp, const uint4 &w, unsigned wn, int *o) {
  const unsigned ws[6] = {wp, w.x, w.y, w.z, w.w, wn};
  int aprev = dp4(__byte_perm(ws[0], ws[1], 0x5432), 0x00010001u, 0) >> 1;  // (s[-2] + s[0]) >> 1
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const unsigned U = (i & 1) ? __byte_perm(ws[1 + (i >> 1)], ws[2 + (i >> 1)], 0x4321)
                               : __byte_perm(ws[i >> 1], ws[1 + (i >> 1)], 0x6543);
    const int a = dp4(U, 0x01000100u, 0) >> 1;
    const int t = dp4(U, 0x00010001u, 0) - a - aprev;
    o[i] = dp4(U, 0x00000100u, tq4(t));
    aprev = a;
  }
}

__device__ __forceinline__ void bar_low128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int N, bool EXTRA>
struct Smem {
  uint8_t p[RB][P_PITCH];        // prediction bytes of the band
  int a1[RB][256];               // level-1 row pass of the band's rows
  int b1[9][256];                // LL1 rows completed in this band
  int a2[N >= 2 ? 9 : 1][128];   // level-2 row pass of those rows
  int b2[N >= 2 ? 6 : 1][128];   // LL2 rows completed in this band
  int a3[N >= 3 ? 6 : 1][64];
  int b3[N >= 3 ? 4 : 1][64];
  int h_pred[EXTRA ? 256 : 1], h_res[EXTRA ? 256 : 1];
};

template <int N, bool EXTRA>
struct Tile {
  const MarchParams &q;
  Smem<N, EXTRA> &sm;
  int tid, lane, warp;
  int c, pair, tx, X0, OW, og0, og1;
  bool border, do_hist, is_I;
  const uint8_t *in, *V0, *V1, *TL;
  uint8_t *out, *pout;
  const short *mvp;
  int plane;
  // G: this thread's 8 columns x 4 rows of the band
  int gx, grow;                 // absolute column, row offset inside the band (4 * (tid >> 6))
  bool g_on;
  // lifting states: level-1 column `tid`, level-2 column `tid` (< 128), level-3 column `tid` (< 64)
  VState s1, s2, s3;
  int carry2, carry3;           // row-passed LL row left without its odd partner at the end of a band

  __device__ __forceinline__ Tile(const MarchParams &qq, Smem<N, EXTRA> &s) : q(qq), sm(s) {}

  // ---------------- G: prediction bytes of rows [r0, r0 + nb)
  __device__ __forceinline__ void gen(int r0, int nb) {
    if (!g_on || grow >= nb) return;
    const int ra = r0 + grow;
    uint2 *dst = reinterpret_cast<uint2 *>(&sm.p[grow][gx - X0]);
    constexpr int DP = P_PITCH / 8;
    if (ra >= q.cy) {
      // rows below the last whole block: chained state (A.2.6)
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (grow + i < nb) dst[i * DP] = __ldg(reinterpret_cast<const uint2 *>(TL + (long long)(ra + i) * q.v_pitch + gx));
      return;
    }
    const short *m = mvp + (ra >> q.bs_shift) * q.BX + (gx >> q.bs_shift);
    const int mx0 = __ldg(m + MV_PREV_X * plane), my0 = __ldg(m + MV_PREV_Y * plane);
    const int mx1 = __ldg(m + MV_NEXT_X * plane), my1 = __ldg(m + MV_NEXT_Y * plane);
    const int col0 = gx + mx0, col1 = gx + mx1;
    const bool inside = col0 >= 0 && col0 + 8 <= q.Xa && col1 >= 0 && col1 + 8 <= q.Xa && ra + my0 >= 0 &&
                        ra + 3 + my0 < q.Ya && ra + my1 >= 0 && ra + 3 + my1 < q.Ya;
    if (inside) {
      const unsigned o0 = (unsigned)((ra + my0) * q.v_pitch + col0), o1 = (unsigned)((ra + my1) * q.v_pitch + col1);
      const int sa = 8 * (int)(o0 & 3), sb = 8 * (int)(o1 & 3);
      const unsigned *pa = reinterpret_cast<const unsigned *>(V0 + (o0 & ~3u));
      const unsigned *pb = reinterpret_cast<const unsigned *>(V1 + (o1 & ~3u));
      const int pw = q.v_pitch >> 2;
      unsigned a[4][3], b[4][3];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
          a[i][k] = __ldg(pa + i * pw + k);
          b[i][k] = __ldg(pb + i * pw + k);
        }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        // (r0 + r1) / 2 on bytes; the [0,255] clip of decorrelate.cpp:841-848 is a no-op
        const unsigned lo = __vhaddu4(__funnelshift_r(a[i][0], a[i][1], sa), __funnelshift_r(b[i][0], b[i][1], sb));
        const unsigned hi = __vhaddu4(__funnelshift_r(a[i][1], a[i][2], sa), __funnelshift_r(b[i][1], b[i][2], sb));
        dst[i * DP] = make_uint2(lo, hi);
      }
    } else {
      for (int i = 0; i < 4; i++) {
        const uint2 ta = bref_row8(V0, q.v_pitch, q.Ya, q.Xa, q.ba, q.padh, ra + i + my0, col0);
        const uint2 tb = bref_row8(V1, q.v_pitch, q.Ya, q.Xa, q.ba, q.padh, ra + i + my1, col1);
        dst[i * DP] = make_uint2(__vhaddu4(ta.x, tb.x), __vhaddu4(ta.y, tb.y));
      }
    }
  }

  // ---------------- R1: level-1 row pass; warp = row, lane = 16 bytes -> 8 outputs
  __device__ __forceinline__ void r1(int nb) {
    for (int row = warp; row < nb; row += 8) {
      const uint4 w = *reinterpret_cast<const uint4 *>(&sm.p[row][16 * lane]);
      const unsigned wp = __shfl_up_sync(FULL, w.w, 1), wn = __shfl_down_sync(FULL, w.x, 1);
      int o[8];
      if (!border) {
        rowpass_bytes(wp, w, wn, o);
      } else {
        int v[19];
        v[0] = (wp >> 16) & 0xff;
        v[1] = wp >> 24;
        const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 16; k++) v[2 + k] = (ws[k >> 2] >> (8 * (k & 3))) & 0xff;
        v[18] = wn & 0xff;
        // absolute level-1 column of output i: (X0 >> 1) + 8 * lane + i
        const int cabs = (X0 >> 1) + 8 * lane;
        const int ifirst = -cabs, ilast = ((q.Xa >> 1) - 1) - cabs;
        rowpass<8, true>(v, o, (ifirst >= 0 && ifirst < 8) ? ifirst : -1, (ilast >= 0 && ilast < 8) ? ilast : -2);
      }
      int *dst = &sm.a1[row][8 * lane];
      *reinterpret_cast<int4 *>(dst) = make_int4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<int4 *>(dst + 4) = make_int4(o[4], o[5], o[6], o[7]);
    }
  }

  // ---------------- R2: level-2 row pass on n1 LL1 rows; warp = row, lane = 8 values -> 4 outputs
  __device__ __forceinline__ void r2(int n1) {
    for (int i = warp; i < n1; i += 8) {
      const int *x = &sm.b1[i][8 * lane];
      const int4 p0 = *reinterpret_cast<const int4 *>(x), p1 = *reinterpret_cast<const int4 *>(x + 4);
      int v[11];
      v[0] = __shfl_up_sync(FULL, p1.z, 1);
      v[1] = __shfl_up_sync(FULL, p1.w, 1);
      v[2] = p0.x, v[3] = p0.y, v[4] = p0.z, v[5] = p0.w, v[6] = p1.x, v[7] = p1.y, v[8] = p1.z, v[9] = p1.w;
      v[10] = __shfl_down_sync(FULL, p0.x, 1);
      int o[4];
      if (!border) {
        rowpass<4, false>(v, o, -1, -2);
      } else {
        const int cabs = (X0 >> 2) + 4 * lane;
        const int ifirst = -cabs, ilast = ((q.Xa >> 2) - 1) - cabs;
        rowpass<4, true>(v, o, (ifirst >= 0 && ifirst < 4) ? ifirst : -1, (ilast >= 0 && ilast < 4) ? ilast : -2);
      }
      *reinterpret_cast<int4 *>(&sm.a2[i][4 * lane]) = make_int4(o[0], o[1], o[2], o[3]);
    }
  }

  // ---------------- R3: level-3 row pass on n2 LL2 rows (first four warps); lane = 4 values -> 2 outputs
  __device__ __forceinline__ void r3(int n2) {
    for (int i = warp; i < n2; i += 4) {
      const int4 p0 = *reinterpret_cast<const int4 *>(&sm.b2[i][4 * lane]);
      int v[7];
      v[0] = __shfl_up_sync(FULL, p0.z, 1);
      v[1] = __shfl_up_sync(FULL, p0.w, 1);
      v[2] = p0.x, v[3] = p0.y, v[4] = p0.z, v[5] = p0.w;
      v[6] = __shfl_down_sync(FULL, p0.x, 1);
      int o[2];
      const int cabs = (X0 >> 3) + 2 * lane;
      const int ifirst = -cabs, ilast = ((q.Xa >> 3) - 1) - cabs;
      rowpass<2, true>(v, o, (ifirst >= 0 && ifirst < 2) ? ifirst : -1, (ilast >= 0 && ilast < 2) ? ilast : -2);
      *reinterpret_cast<int2 *>(&sm.a3[i][2 * lane]) = make_int2(o[0], o[1]);
    }
  }

  // ---------------- OUT: LL rows [an, bn) of the last level (buffer row 0 = row an), 4 samples per item
  __device__ __forceinline__ void emit(int an, int bn, int nthreads) {
    constexpr int LR = RING >> N, LV = TW >> N;
    const int e0 = max(an, og0), e1 = min(bn, og1);
    if (e0 >= e1) return;
    const int cbase = tx * LV;  // component column of local column LR
    const int per_row = min(LV, OW - cbase) >> 2;
    for (int it = tid; it < (e1 - e0) * per_row; it += nthreads) {
      const int er = it / per_row, k4 = (it - er * per_row) * 4, e = e0 + er;
      const int *p = (N == 1 ? &sm.b1[e - an][0] : (N == 2 ? &sm.b2[e - an][0] : &sm.b3[e - an][0])) + LR + k4;
      const long long o = (long long)e * OW + cbase + k4;
      const unsigned sw = *reinterpret_cast<const unsigned *>(in + o);
      unsigned ow = 0, pw = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int s = (sw >> (8 * k)) & 0xff, pv = p[k];
        int v;
        if (!q.synth) {
          int rr = s - pv;
          rr = rr < -128 ? -128 : (rr > 127 ? 127 : rr);
          v = rr + 128;
          if (EXTRA && do_hist) {
            atomicAdd(&sm.h_pred[s], 1);
            atomicAdd(&sm.h_res[v], 1);
          }
        } else if (is_I) {
          v = s;
        } else {
          v = s - 128 + pv;
          v = v < 0 ? 0 : (v > 255 ? 255 : v);
        }
        ow |= (unsigned)v << (8 * k);
        pw |= (unsigned)(pv & 0xff) << (8 * k);
      }
      *reinterpret_cast<unsigned *>(out + o) = ow;
      if (EXTRA && pout) *reinterpret_cast<unsigned *>(pout + o) = pw;
    }
  }

  // One band.  STEADY: 16 rows, r0 >= 32, not the last band of the picture -- every count below is
  // a constant and no lifting step is the first or the virtual last one of its line.
  template <bool STEADY>
  __device__ __forceinline__ void band(int r0, int nb) {
    gen(r0, nb);
    __syncthreads();
    r1(STEADY ? RB : nb);
    __syncthreads();
    // ---- C1: rows [r0, r0 + nb) -> steps [ta1, tb1) (+ the virtual step at the end of the picture)
    const bool fin = !STEADY && r0 + nb == q.Ya;
    const int ta1 = r0 >> 1, tb1 = STEADY ? ta1 + 8 : (r0 + nb) >> 1;
    const int a1 = STEADY ? ta1 - 1 : max(ta1 - 1, 0), b1 = tb1 - 1 + (fin ? 1 : 0);  // new LL1 rows [a1, b1)
    if (STEADY) {
#pragma unroll
      for (int k = 0; k < 8; k++) sm.b1[k][tid] = vstep(s1, sm.a1[2 * k][tid], sm.a1[2 * k + 1][tid]);
    } else {
      for (int t = ta1; t < tb1 + (fin ? 1 : 0); t++) {
        const bool flush = t == (q.Ya >> 1);
        const int xe = flush ? 0 : sm.a1[2 * (t - ta1)][tid], xo = flush ? 0 : sm.a1[2 * (t - ta1) + 1][tid];
        const int l = vstep_sp(s1, t == 1, flush, xe, xo);
        if (t >= 1) sm.b1[t - 1 - a1][tid] = l;
      }
    }
    __syncthreads();
    if (N == 1) {
      emit(a1, b1, NT);  // b1 is rewritten by the next band's C1, two block barriers from here
      return;
    }
    r2(b1 - a1);
    __syncthreads();
    if (warp >= 4) return;  // the upper four warps go on to the next band's G
    // ---- C2 (128 threads): LL1 rows [a1, b1) -> steps [a1 >> 1, b1 >> 1)
    const int ta2 = a1 >> 1, tb2 = b1 >> 1;
    const int a2 = STEADY ? ta2 - 1 : max(ta2 - 1, 0), b2 = tb2 - 1 + (fin ? 1 : 0);
    if (STEADY) {
      // a1 is odd: the first step pairs the carried row with buffer row 0
      sm.b2[0][tid] = vstep(s2, carry2, sm.a2[0][tid]);
#pragma unroll
      for (int k = 1; k < 4; k++) sm.b2[k][tid] = vstep(s2, sm.a2[2 * k - 1][tid], sm.a2[2 * k][tid]);
      carry2 = sm.a2[7][tid];
    } else {
      for (int t = ta2; t < tb2 + (fin ? 1 : 0); t++) {
        const bool flush = t == (q.Ya >> 2);
        const int xe = flush ? 0 : (2 * t < a1 ? carry2 : sm.a2[2 * t - a1][tid]);
        const int xo = flush ? 0 : sm.a2[2 * t + 1 - a1][tid];
        const int l = vstep_sp(s2, t == 1, flush, xe, xo);
        if (t >= 1) sm.b2[t - 1 - a2][tid] = l;
      }
      if ((b1 & 1) && b1 > a1) carry2 = sm.a2[b1 - 1 - a1][tid];
    }
    bar_low128();
    if (N == 2) {
      emit(a2, b2, 128);  // b2 is rewritten by the next band's C2, block barriers away
      return;
    }
    r3(b2 - a2);
    bar_low128();
    // ---- C3 (64 threads)
    const int ta3 = a2 >> 1, tb3 = b2 >> 1;
    const int a3 = STEADY ? ta3 - 1 : max(ta3 - 1, 0), b3 = tb3 - 1 + (fin ? 1 : 0);
    if (tid < 64) {
      if (STEADY) {
        // a2 is even: both steps pair rows of this band
        sm.b3[0][tid] = vstep(s3, sm.a3[0][tid], sm.a3[1][tid]);
        sm.b3[1][tid] = vstep(s3, sm.a3[2][tid], sm.a3[3][tid]);
      } else {
        for (int t = ta3; t < tb3 + (fin ? 1 : 0); t++) {
          const bool flush = t == (q.Ya >> 3);
          const int xe = flush ? 0 : (2 * t < a2 ? carry3 : sm.a3[2 * t - a2][tid]);
          const int xo = flush ? 0 : sm.a3[2 * t + 1 - a2][tid];
          const int l = vstep_sp(s3, t == 1, flush, xe, xo);
          if (t >= 1) sm.b3[t - 1 - a3][tid] = l;
        }
        if ((b2 & 1) && b2 > a2) carry3 = sm.a3[b2 - 1 - a2][tid];
      }
    }
    bar_low128();
    emit(a3, b3, 128);
  }
};

template <int N, bool EXTRA>
__global__ void __launch_bounds__(NT, 4) k_mc_tile(MarchParams q, int c0, int nc, int npairs, int G, int tiles_x) {
  __shared__ __align__(16) Smem<N, EXTRA> sm;
  Tile<N, EXTRA> t(q, sm);
  t.tid = threadIdx.x;
  t.lane = t.tid & 31;
  t.warp = t.tid >> 5;
  int idx = blockIdx.x;
  t.c = c0 + idx % nc;
  idx /= nc;
  const int pg = idx % G;
  idx /= G;
  t.tx = idx % tiles_x;
  const int seg = idx / tiles_x;
  t.pair = blockIdx.z * G + pg;
  if (t.pair >= npairs) return;

  t.X0 = t.tx * TW - RING;  // absolute prediction column of local column 0
  t.OW = t.c ? q.X >> 1 : q.X;
  const int OH = q.Ya >> N;
  t.og0 = (int)(((long long)seg * q.seg_p) >> N);
  t.og1 = min(OH, (int)(((long long)(seg + 1) * q.seg_p) >> N));
  if (t.og0 >= t.og1) return;
  const int rstart = max(0, (t.og0 << N) - RB), rend = min(q.Ya, (t.og1 << N) + RB);
  t.border = t.tx == 0 || (t.X0 + TWP >= q.Xa);
  t.do_hist = EXTRA && q.hist && t.c == 0 && !q.synth;
  if (EXTRA) {
    for (int i = t.tid; i < 256; i += NT) sm.h_pred[i] = sm.h_res[i] = 0;
  }
  const long long coff = t.c == 0 ? 0 : (long long)q.X * q.Y + (long long)(t.c - 1) * (q.X / 2) * (q.Y / 2);
  t.in = q.in + (long long)t.pair * q.in_stride + coff;
  t.out = q.out + (long long)t.pair * q.out_stride + coff;
  t.pout = (EXTRA && q.prediction) ? q.prediction + (long long)t.pair * q.pred_stride + coff : nullptr;
  t.is_I = q.synth && q.types[t.pair] == 'I';
  t.V0 = q.v + ((long long)(q.f0 + t.pair) * 3 + t.c) * q.v_plane_stride;
  t.V1 = t.V0 + 3 * q.v_plane_stride;
  t.TL = q.tail ? q.tail + ((long long)t.pair * 3 + t.c) * q.tail_plane_stride : nullptr;
  t.plane = q.BY * q.BX;
  t.mvp = q.mv + (long long)t.pair * 4 * t.plane;
  t.gx = t.X0 + 8 * (t.tid & 63);
  t.grow = 4 * (t.tid >> 6);
  t.g_on = t.gx >= 0 && t.gx < q.Xa;
  t.s1 = t.s2 = t.s3 = VState{0, 0, 0};
  t.carry2 = t.carry3 = 0;

  for (int r0 = rstart; r0 < rend; r0 += RB) {
    const int nb = min(RB, q.Ya - r0);
    if (nb == RB && r0 >= 2 * RB && r0 + RB < q.Ya) t.template band<true>(r0, nb);
    else t.template band<false>(r0, nb);
  }
  if (EXTRA && t.do_hist) {
    __syncthreads();
    int *hist = q.hist + (long long)t.pair * q.hist_stride;
    for (int i = t.tid; i < 256; i += NT) {
      if (sm.h_pred[i]) atomicAdd(&hist[i], sm.h_pred[i]);
      if (sm.h_res[i]) atomicAdd(&hist[256 + i], sm.h_res[i]);
    }
  }
}

template <int N>
void launch_tile_n(const Launch &L, const MarchParams &q, int npairs, int c0, int nc) {
  const int tiles_x = (q.Xa + TW - 1) / TW;
  const int G = npairs < 4 ? npairs : 4;
  dim3 grid(nc * G * tiles_x * q.nsegs, 1, (npairs + G - 1) / G);
  ProfScope ps_(L, KC_RESIDUE);
  if (q.hist || q.prediction) k_mc_tile<N, true><<<grid, NT, 0, L.stream>>>(q, c0, nc, npairs, G, tiles_x);
  else k_mc_tile<N, false><<<grid, NT, 0, L.stream>>>(q, c0, nc, npairs, G, tiles_x);
  COUNT(L);
}

}  // namespace

bool mc_tile_supported(int a, int bsa, int X) { return (a == 1 || a == 2) && bsa % RB == 0 && X % 8 == 0; }

void launch_mc_tile(const Launch &L, MarchParams q, int npairs) {
  if (npairs <= 0) return;
  q.bs_shift = 0;
  while ((1 << q.bs_shift) < q.bsa) q.bs_shift++;
  // segments: enough CTAs to fill the machine a few times over, as tall as that allows (every
  // segment pays two extra bands: one to warm the lifting states up, one for the reach at its end)
  const int tiles_x = (q.Xa + TW - 1) / TW;
  static const int want_ctas = getenv("QSVC_TILE_CTAS") ? atoi(getenv("QSVC_TILE_CTAS")) : 148 * 5 * 3;
  const long long per_seg = (long long)tiles_x * npairs * 3;
  long long nsegs = (want_ctas + per_seg - 1) / per_seg;
  int seg_p = (int)((q.Ya + nsegs - 1) / nsegs);
  seg_p = (seg_p + 15) & ~15;
  if (seg_p < 64) seg_p = 64;
  q.seg_p = seg_p;
  q.nsegs = (q.Ya + seg_p - 1) / seg_p;
  if (q.a == 1) {
    launch_tile_n<1>(L, q, npairs, 0, 1);
    launch_tile_n<2>(L, q, npairs, 1, 2);
  } else {
    launch_tile_n<2>(L, q, npairs, 0, 1);
    launch_tile_n<3>(L, q, npairs, 1, 2);
  }
}
