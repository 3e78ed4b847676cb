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
// bands of 16 prediction rows.  Per band, with a block barrier between the phases:
//   G   the 16 x 512 prediction bytes: two displaced 8-byte windows of the reference planes per
//       thread and row, averaged as packed bytes; blocks whose footprint leaves the picture take
//       the reference's border rule per sample; rows below the last whole block come from the
//       chained tail state (k_tail_state);
//   R1  row pass of level 1 on the bytes -> 256 int32 low-pass columns per row;
//   C1  column pass of level 1: one thread per column streams the 16 rows through the lifting
//       step (state in registers across bands) and emits the LL1 rows it completes;
//   R2 / C2 / R3 / C3  the same on the LL rows of the level above (128 / 64 columns);
//   OUT the LL rows of the last level become residue / reconstruction bytes (and histograms,
//       prediction side output) with 32-bit accesses.
// Every phase is a flat loop over independent items; nothing but the two reference planes, the
// input frame and the output frame touches HBM.
#include <cstdlib>

#include "kernels.cuh"

#define COUNT(L) (++*(L).counter)

namespace {

constexpr int TWP = 512;            // prediction columns per tile, ring included
constexpr int RING = 16;            // ring on either side (>= 14 = reach of three levels; multiple of 8)
constexpr int TW = TWP - 2 * RING;  // columns a tile produces output for
constexpr int RB = 16;              // prediction rows per band
constexpr int P_PITCH = 576;        // bytes per prediction row in shared memory (>= TWP + 4)
constexpr int NT = 256;

__device__ __forceinline__ int tq2(int v) { return (v + (int)((unsigned)v >> 31)) >> 1; }  // C "/ 2"
__device__ __forceinline__ int tq4(int v) { return (v + (int)((unsigned)v >> 30)) >> 2; }  // C "/ 4", |v| < 2^30

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

// Eight samples V(y, x0 .. x0+7) of a bordered reference plane whose window leaves the picture on
// the left or right (rare: edge blocks with outward vectors).
__device__ __noinline__ uint2 bref_row8(const uint8_t *U, int pitch, int Yd, int Xd, int b, int padh, int y, int x0) {
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
__device__ __forceinline__ int vstep(VState &s, bool first, bool flush, int xe, int xo) {
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
// column (x[n] := x[n-2]); -1 when the run does not contain them.
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

template <int N, bool EXTRA>
struct Smem {
  // the prediction bytes of a band are dead once R1 has run; the LL1 rows of C1 take their place
  union {
    uint8_t p[RB][P_PITCH];
    int b1[9][256];
  } u;
  int a1[RB][256];
  int a2[N >= 2 ? 16 : 1][128];  // ring over LL1 row index
  int b2[N >= 2 ? 6 : 1][128];
  int a3[N >= 3 ? 8 : 1][64];    // ring over LL2 row index
  int b3[N >= 3 ? 4 : 1][64];
  int h_pred[EXTRA ? 256 : 1], h_res[EXTRA ? 256 : 1];
};

template <int N, bool EXTRA>
__global__ void __launch_bounds__(NT, 4) k_mc_tile(MarchParams q, int c0, int nc, int npairs, int G, int tiles_x) {
  __shared__ __align__(16) Smem<N, EXTRA> sm;
  const int tid = threadIdx.x;
  int idx = blockIdx.x;
  const int c = c0 + idx % nc;
  idx /= nc;
  const int pg = idx % G;
  idx /= G;
  const int tx = idx % tiles_x, seg = idx / tiles_x;
  const int pair = blockIdx.z * G + pg;
  if (pair >= npairs) return;

  const int X0 = tx * TW - RING;       // absolute prediction column of local column 0
  const int OW = c ? q.X >> 1 : q.X;   // component width
  const int OH = q.Ya >> N;            // component height
  const int og0 = (int)(((long long)seg * q.seg_p) >> N);
  const int og1 = min(OH, (int)(((long long)(seg + 1) * q.seg_p) >> N));
  if (og0 >= og1) return;
  const int rstart = max(0, (og0 << N) - RB), rend = min(q.Ya, (og1 << N) + RB);
  const bool border = tx == 0 || (X0 + TWP >= q.Xa);

  const bool do_hist = EXTRA && q.hist && c == 0 && !q.synth;
  if (EXTRA) {
    for (int i = tid; i < 256; i += NT) sm.h_pred[i] = sm.h_res[i] = 0;
  }
  const long long coff = c == 0 ? 0 : (long long)q.X * q.Y + (long long)(c - 1) * (q.X / 2) * (q.Y / 2);
  const uint8_t *in = q.in + (long long)pair * q.in_stride + coff;
  uint8_t *out = q.out + (long long)pair * q.out_stride + coff;
  uint8_t *pout = (EXTRA && q.prediction) ? q.prediction + (long long)pair * q.pred_stride + coff : nullptr;
  const bool is_I = q.synth && q.types[pair] == 'I';
  const uint8_t *V0 = q.v + ((long long)(q.f0 + pair) * 3 + c) * q.v_plane_stride;
  const uint8_t *V1 = V0 + 3 * q.v_plane_stride;
  const uint8_t *TL = q.tail ? q.tail + ((long long)pair * 3 + c) * q.tail_plane_stride : nullptr;
  const int plane = q.BY * q.BX;
  const short *mvp = q.mv + (long long)pair * 4 * plane;

  VState s1 = {0, 0, 0}, s2 = {0, 0, 0}, s3 = {0, 0, 0};
  const int half1 = q.Ya >> 1, half2 = q.Ya >> 2, half3 = q.Ya >> 3;
  const int w1 = q.Xa >> 1, w2 = q.Xa >> 2, w3 = q.Xa >> 3;  // line lengths met by the row passes 2, 3 (and L3 width)
  (void)w3;

  for (int r0 = rstart; r0 < rend; r0 += RB) {
    const int nb = min(RB, q.Ya - r0);
    __syncthreads();  // the previous band's consumers are done with every buffer
    // ---------------- G: prediction bytes of rows [r0, r0 + nb)
    {
      const int gc = tid & 63, rc = tid >> 6;
      const int x = X0 + 8 * gc;
      if (x >= 0 && x < q.Xa) {
        const int ra = r0 + 4 * rc;
        if (ra >= q.cy) {
          // rows below the last whole block: chained state (A.2.6)
#pragma unroll
          for (int i = 0; i < 4; i++)
            if (4 * rc + i < nb) {
              const uint2 t = __ldg(reinterpret_cast<const uint2 *>(TL + (long long)(ra + i) * q.v_pitch + x));
              *reinterpret_cast<uint2 *>(&sm.u.p[4 * rc + i][8 * gc]) = t;
            }
        } else if (4 * rc < nb) {
          const int by = ra >> q.bs_shift, bx = x >> q.bs_shift;
          const short *m = mvp + by * q.BX + bx;
          const int mx0 = __ldg(m + MV_PREV_X * plane), my0 = __ldg(m + MV_PREV_Y * plane);
          const int mx1 = __ldg(m + MV_NEXT_X * plane), my1 = __ldg(m + MV_NEXT_Y * plane);
          const int col0 = x + mx0, col1 = x + mx1;
          const bool xin0 = col0 >= 0 && col0 + 8 <= q.Xa, xin1 = col1 >= 0 && col1 + 8 <= q.Xa;
          const bool yin = ra + my0 >= 0 && ra + 3 + my0 < q.Ya && ra + my1 >= 0 && ra + 3 + my1 < q.Ya;
          if (xin0 && xin1 && yin) {
            const long long o0 = (long long)(ra + my0) * q.v_pitch + col0, o1 = (long long)(ra + my1) * q.v_pitch + col1;
            const int sa = 8 * (int)(o0 & 3), sb = 8 * (int)(o1 & 3);
            const unsigned *pa = reinterpret_cast<const unsigned *>(V0 + (o0 & ~3LL));
            const unsigned *pb = reinterpret_cast<const unsigned *>(V1 + (o1 & ~3LL));
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
              *reinterpret_cast<uint2 *>(&sm.u.p[4 * rc + i][8 * gc]) = make_uint2(lo, hi);
            }
          } else {
            for (int i = 0; i < 4; i++) {
              const int r = ra + i;
              uint2 ta, tb;
              if (xin0) {
                const uint8_t *p = V0 + (long long)min(max(r + my0, 0), q.Ya - 1) * q.v_pitch + col0;
                const unsigned *p4 = reinterpret_cast<const unsigned *>((uintptr_t)p & ~(uintptr_t)3);
                const int s = 8 * (int)((uintptr_t)p & 3);
                const unsigned w0 = __ldg(p4), w1_ = __ldg(p4 + 1), w2_ = __ldg(p4 + 2);
                ta = make_uint2(__funnelshift_r(w0, w1_, s), __funnelshift_r(w1_, w2_, s));
              } else {
                ta = bref_row8(V0, q.v_pitch, q.Ya, q.Xa, q.ba, q.padh, r + my0, col0);
              }
              if (xin1) {
                const uint8_t *p = V1 + (long long)min(max(r + my1, 0), q.Ya - 1) * q.v_pitch + col1;
                const unsigned *p4 = reinterpret_cast<const unsigned *>((uintptr_t)p & ~(uintptr_t)3);
                const int s = 8 * (int)((uintptr_t)p & 3);
                const unsigned w0 = __ldg(p4), w1_ = __ldg(p4 + 1), w2_ = __ldg(p4 + 2);
                tb = make_uint2(__funnelshift_r(w0, w1_, s), __funnelshift_r(w1_, w2_, s));
              } else {
                tb = bref_row8(V1, q.v_pitch, q.Ya, q.Xa, q.ba, q.padh, r + my1, col1);
              }
              *reinterpret_cast<uint2 *>(&sm.u.p[4 * rc + i][8 * gc]) = make_uint2(__vhaddu4(ta.x, tb.x), __vhaddu4(ta.y, tb.y));
            }
          }
        }
      }
    }
    __syncthreads();
    // ---------------- R1: level-1 row pass, 8 outputs per item
    {
      const int nitems = nb * 32;
      for (int it = tid; it < nitems; it += NT) {
        const int row = it >> 5, j = it & 31;
        const uint8_t *pr = &sm.u.p[row][0];
        const uint4 w = *reinterpret_cast<const uint4 *>(pr + 16 * j);
        const unsigned wp = j > 0 ? *reinterpret_cast<const unsigned *>(pr + 16 * j - 4) : 0u;
        const unsigned wn = *reinterpret_cast<const unsigned *>(pr + 16 * j + 16);  // j == 31: the row's padding
        int v[19];
        v[0] = (wp >> 16) & 0xff;
        v[1] = wp >> 24;
        const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 16; k++) v[2 + k] = (ws[k >> 2] >> (8 * (k & 3))) & 0xff;
        v[18] = wn & 0xff;
        int o[8];
        if (border) {
          // absolute level-1 column of output i: (X0 >> 1) + 8j + i
          const int cabs = (X0 >> 1) + 8 * j;
          const int ifirst = -cabs, ilast = (w1 - 1) - cabs;
          rowpass<8, true>(v, o, (ifirst >= 0 && ifirst < 8) ? ifirst : -1, (ilast >= 0 && ilast < 8) ? ilast : -2);
        } else {
          rowpass<8, false>(v, o, -1, -2);
        }
        int *dst = &sm.a1[row][8 * j];
        *reinterpret_cast<int4 *>(dst) = make_int4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<int4 *>(dst + 4) = make_int4(o[4], o[5], o[6], o[7]);
      }
    }
    __syncthreads();
    // ---------------- C1: level-1 column pass (one column per thread)
    const bool fin1 = r0 + nb == q.Ya;
    const int ta1 = r0 >> 1, tb1 = (r0 + nb) >> 1;
    const int a1 = max(ta1 - 1, 0), b1 = tb1 - 1 + (fin1 ? 1 : 0);  // new LL1 rows [a1, b1)
    for (int t = ta1; t < tb1 + (fin1 ? 1 : 0); t++) {
      const bool flush = t == half1;
      const int xe = flush ? 0 : sm.a1[2 * (t - ta1)][tid], xo = flush ? 0 : sm.a1[2 * (t - ta1) + 1][tid];
      const int l = vstep(s1, t == 1, flush, xe, xo);
      if (t >= 1) sm.u.b1[t - 1 - a1][tid] = l;
    }
    // NB: C1 writes b1 (aliases p) while reading a1 only; R1 finished before the barrier above.
    __syncthreads();
    int an = a1, bn = b1;  // rows of the last level completed in this band
    if (N >= 2) {
      // ---------------- R2: level-2 row pass on the new LL1 rows, 4 outputs per item
      const int nrows = b1 - a1;
      for (int it = tid; it < nrows * 32; it += NT) {
        const int i = it >> 5, m = it & 31;
        const int *x = &sm.u.b1[i][0];
        int v[11];
        if (m > 0) {
          const int2 t = *reinterpret_cast<const int2 *>(x + 8 * m - 2);
          v[0] = t.x;
          v[1] = t.y;
        } else {
          v[0] = v[1] = 0;
        }
        const int4 p0 = *reinterpret_cast<const int4 *>(x + 8 * m), p1 = *reinterpret_cast<const int4 *>(x + 8 * m + 4);
        v[2] = p0.x, v[3] = p0.y, v[4] = p0.z, v[5] = p0.w, v[6] = p1.x, v[7] = p1.y, v[8] = p1.z, v[9] = p1.w;
        v[10] = m < 31 ? x[8 * m + 8] : 0;
        int o[4];
        if (border) {
          const int cabs = (X0 >> 2) + 4 * m;
          const int ifirst = -cabs, ilast = (w2 - 1) - cabs;
          rowpass<4, true>(v, o, (ifirst >= 0 && ifirst < 4) ? ifirst : -1, (ilast >= 0 && ilast < 4) ? ilast : -2);
        } else {
          rowpass<4, false>(v, o, -1, -2);
        }
        *reinterpret_cast<int4 *>(&sm.a2[(a1 + i) & 15][4 * m]) = make_int4(o[0], o[1], o[2], o[3]);
      }
      __syncthreads();
      // ---------------- C2
      const bool fin2 = fin1;
      const int ta2 = a1 >> 1, tb2 = b1 >> 1;
      const int a2 = max(ta2 - 1, 0), b2 = tb2 - 1 + (fin2 ? 1 : 0);
      if (tid < 128) {
        for (int t = ta2; t < tb2 + (fin2 ? 1 : 0); t++) {
          const bool flush = t == half2;
          const int xe = flush ? 0 : sm.a2[(2 * t) & 15][tid], xo = flush ? 0 : sm.a2[(2 * t + 1) & 15][tid];
          const int l = vstep(s2, t == 1, flush, xe, xo);
          if (t >= 1) sm.b2[t - 1 - a2][tid] = l;
        }
      }
      __syncthreads();
      an = a2, bn = b2;
      if (N >= 3) {
        // ---------------- R3: 1 output per item
        const int nrows3 = b2 - a2;
        for (int it = tid; it < nrows3 * 64; it += NT) {
          const int i = it >> 6, m = it & 63;
          const int *x = &sm.b2[i][0];
          int v[5];
          v[0] = m > 0 ? x[2 * m - 2] : 0;
          v[1] = m > 0 ? x[2 * m - 1] : 0;
          v[2] = x[2 * m];
          v[3] = x[2 * m + 1];
          v[4] = m < 63 ? x[2 * m + 2] : 0;
          int o[1];
          const int cabs = (X0 >> 3) + m;
          rowpass<1, true>(v, o, cabs == 0 ? 0 : -1, cabs == w3 - 1 ? 0 : -2);
          sm.a3[(a2 + i) & 7][m] = o[0];
        }
        __syncthreads();
        // ---------------- C3
        const int ta3 = a2 >> 1, tb3 = b2 >> 1;
        const int a3 = max(ta3 - 1, 0), b3 = tb3 - 1 + (fin2 ? 1 : 0);
        if (tid < 64) {
          for (int t = ta3; t < tb3 + (fin2 ? 1 : 0); t++) {
            const bool flush = t == half3;
            const int xe = flush ? 0 : sm.a3[(2 * t) & 7][tid], xo = flush ? 0 : sm.a3[(2 * t + 1) & 7][tid];
            const int l = vstep(s3, t == 1, flush, xe, xo);
            if (t >= 1) sm.b3[t - 1 - a3][tid] = l;
          }
        }
        __syncthreads();
        an = a3, bn = b3;
      }
    }
    // ---------------- OUT: LL rows [an, bn) of the last level, 4 samples per item
    {
      constexpr int LW = TWP >> N;          // columns per tile at this level
      constexpr int LR = RING >> N, LV = TW >> N;
      const int e0 = max(an, og0), e1 = min(bn, og1);
      const int cbase = tx * LV;            // component column of local column LR
      const int ncols = min(LV, OW - cbase);
      const int per_row = (ncols + 3) >> 2;
      for (int it = tid; it < (e1 - e0) * per_row; it += NT) {
        const int e = e0 + it / per_row, k4 = (it % per_row) * 4;
        const int *p = (N == 1 ? &sm.u.b1[e - an][0] : (N == 2 ? &sm.b2[e - an][0] : &sm.b3[e - an][0])) + LR + k4;
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
      (void)LW;
    }
  }
  if (EXTRA && do_hist) {
    __syncthreads();
    int *hist = q.hist + (long long)pair * q.hist_stride;
    for (int i = tid; i < 256; i += NT) {
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
