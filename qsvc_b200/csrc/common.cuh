// common.cuh -- shared device/host helpers for the MCTF kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// Motion-field plane order inside one field (reference motion.cpp:9-15,93-101).
#define MV_PREV_X 0
#define MV_PREV_Y 1
#define MV_NEXT_X 2
#define MV_NEXT_Y 3

// An int16 image whose addressing reproduces the reference's
// texture::alloc (texture.cpp:34-46) on a fresh glibc heap:
//   - every row is its own malloc chunk of `S` shorts (16-byte granules, 8-byte
//     size field right before the row's user pointer);
//   - data[y][x] = rowptr[y + b][x], and only the first `y_dim` physical row
//     pointers are shifted right by the border `b` (reference bug, kept), so
//     logical rows [y_dim - b, y_dim + b) address physical column x directly
//     and x < 0 lands in the size field / the tail of the previous row.
// `base` is the user pointer of physical row 0 of slot 0.  With b == 0 and an
// arbitrary S this degenerates to a dense image (used by decorrelate/update).
struct Plane {
  short *base;
  long long slot_stride;  // shorts between consecutive slots
  int S;                  // shorts between consecutive row user pointers
  int y_dim, x_dim, b;

  __host__ __device__ __forceinline__ short *row(int slot, int y) const {
    int r = y + b;
    return base + (long long)slot * slot_stride + (long long)r * S + (r < y_dim ? b : 0);
  }
};

__host__ __device__ __forceinline__ int iclamp(int v, int lo, int hi) {
  return v < lo ? lo : (v > hi ? hi : v);
}

// ---- 5/3 integer lifting on a strided line (reference 5_3.cpp:39-115) ----
// C `/` truncates toward zero; results are stored as short (wraps mod 2^16).
// The truncating divisions by 2 and 4 are spelled as shift sequences: with a
// plain `/` nvcc merges the /2 and /4 sites of different branches into one
// division by a *variable* and calls its 16-bit division subroutine per sample.
__host__ __device__ __forceinline__ int tdiv2(int x) { return (x + (int)((unsigned)x >> 31)) >> 1; }
__host__ __device__ __forceinline__ int tdiv4(int x) { return (x + ((x >> 31) & 3)) >> 2; }

// h[i] of an n-sample line s (analysis).
__device__ __forceinline__ short l53_ana_h(const short *s, int st, int i, int n) {
  int half = n >> 1;
  if (!(n & 1) && i == half - 1) return (short)(s[(n - 1) * st] - s[(n - 2) * st]);
  return (short)(s[(2 * i + 1) * st] - tdiv2(s[(2 * i) * st] + s[(2 * i + 2) * st]));
}
// l[i] (analysis); i in [0, n - n/2).
__device__ __forceinline__ short l53_ana_l(const short *s, int st, int i, int n) {
  int half = n >> 1;
  if (i == 0) return (short)(s[0] + tdiv2(l53_ana_h(s, st, 0, n)));
  if (i < half)
    return (short)(s[(2 * i) * st] + tdiv4(l53_ana_h(s, st, i, n) + l53_ana_h(s, st, i - 1, n)));
  return (short)(s[(n - 1) * st] + tdiv2(l53_ana_h(s, st, half - 1, n)));  // odd n tail
}
// Output sample j of the analysed line laid out [lows | highs].
__device__ __forceinline__ short l53_ana_out(const short *s, int st, int j, int n) {
  int nlow = n - (n >> 1);
  return j < nlow ? l53_ana_l(s, st, j, n) : l53_ana_h(s, st, j - nlow, n);
}

// Synthesis: l = s (lows at 0), h at offset hoff = n - n/2.
__device__ __forceinline__ short l53_syn_even(const short *l, const short *h, int st, int i, int n) {
  int half = n >> 1;
  if (i == 0) return (short)(l[0] - tdiv2(h[0]));
  if (i < half) return (short)(l[i * st] - tdiv4(h[i * st] + h[(i - 1) * st]));
  return (short)(l[half * st] - tdiv2(h[(half - 1) * st]));  // odd n tail
}
// Output sample j of the synthesised line.
__device__ __forceinline__ short l53_syn_out(const short *s, int st, int j, int n) {
  int half = n >> 1;
  const short *l = s;
  const short *h = s + (long long)(n - half) * st;
  if (!(j & 1)) return l53_syn_even(l, h, st, j >> 1, n);
  int i = j >> 1;
  int e0 = l53_syn_even(l, h, st, i, n);
  if (!(n & 1) && i == half - 1) return (short)(h[i * st] + e0);
  int e1 = l53_syn_even(l, h, st, i + 1, n);
  return (short)(h[i * st] + tdiv2(e0 + e1));
}

// Value seen through data[y][x] of a bordered texture whose alloc and
// fill_border got the same (Yd, Xd, b) -- decorrelate's references
// (decorrelate.cpp:539-553,681-684).  U is the dense Yd x Xd interior with row
// stride S.  Closed form of the heap model (validated against the oracle):
// edge replication, except (i) the bottom-left corner replicates the
// bottom-RIGHT pixel (texture.cpp:92-97), (ii) the left ring of the first
// unshifted row y0 = Yd - b aliases the right ring of the row above it.
// Coordinates further than b outside the image are undefined in the reference
// (out-of-bounds reads); here they are clamped like the ring.
__device__ __forceinline__ short bordered_ref(const short *U, int S, int Yd, int Xd, int b,
                                              int padh, int y, int x) {
  if ((unsigned)y < (unsigned)Yd && (unsigned)x < (unsigned)Xd) return U[(long long)y * S + x];
  if (b > padh) {
    int y0 = Yd - b;
    if (y0 != 0) {
      if (y == y0 && x < -padh) return U[(long long)iclamp(y0 - 1, 0, Yd - 1) * S + Xd - 1];
    } else if (y == -1 && x >= Xd + padh) {
      return U[0];
    }
  }
  if (y >= Yd && x < 0) return U[(long long)(Yd - 1) * S + Xd - 1];
  return U[(long long)iclamp(y, 0, Yd - 1) * S + iclamp(x, 0, Xd - 1)];
}
