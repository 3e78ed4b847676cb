// kernels_mcfused.cu -- motion-compensated prediction on byte planes.
//
// Reference pipeline per pair (decorrelate.cpp:732-861): both even frames are
// brought to luma size << a by zero-high-band 5/3 synthesis, predict() averages
// the two displaced references per block, the result is clipped to [0,255],
// analysed `a` levels (chroma one more) in place, and only the LL band is used.
//
// On byte data every step before the analysis stays inside [0,255], so it runs on
// u8 planes with packed-byte arithmetic:
//   V_a  = up^a(component)          k_upsample2x (kernels_subpel.cu), once per even frame
//   P_a  = (V_a0[+mv0] + V_a1[+mv1]) >> 1   k_predict_u8: __vhaddu4 on 32-bit words; blocks
//          whose displaced footprint leaves the picture use the closed-form border
//          rule of the reference's texture::alloc/fill_border (bordered_ref)
//   LL   = LL-only multi-level integer 5/3 analysis of P_a fused with the residue /
//          reconstruction and the I/B histograms: k_mc_march (kernels_mcmarch.cu), which also
//          folds the prediction in, so that P_a is only materialised for the last block row
//          that feeds the tail chain below.
// Rows of P_a below the last whole block (Y % block_size != 0) are never written by
// predict(); the reference keeps there what the previous pair's in-place analysis
// left.  k_tail_state reproduces that chain: it is the only
// sequential step and touches 2*(Ya - cy) rows per pair.
#include <algorithm>

#include "kernels.cuh"

#define COUNT(L) (++*(L).counter)

__device__ __forceinline__ int bordered_ref_u8(const uint8_t *U, int pitch, int Yd, int Xd, int b,
                                               int padh, int y, int x) {
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

// Materialises the reference's border rule (texture::alloc + fill_border, bordered_ref_u8) as a
// ring of `ring` samples around the interior of every plane, so that displaced windows which
// leave the picture by less than that are plain loads.  U = interior origin of plane 0.
// Above and below the picture the columns inside it repeat the first / last row (none of the
// rule's quirks applies there): 8-byte copies.  The columns left and right of the picture, for
// every ringed row, go through the closed form, four cells per thread and store.
__global__ void __launch_bounds__(256) k_ring_bands(uint8_t *U, long long plane_stride, int pitch, int Yd, int Xd,
                                                    int ring) {
  uint8_t *P = U + (long long)blockIdx.z * plane_stride;
  const int per_row = Xd >> 3, total = 2 * ring * per_row;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / per_row, k = i - r * per_row;
    const int y = r < ring ? r - ring : Yd + (r - ring);
    reinterpret_cast<uint2 *>(P + (long long)y * pitch)[k] =
        reinterpret_cast<const uint2 *>(P + (long long)(y < 0 ? 0 : Yd - 1) * pitch)[k];
  }
}
__global__ void __launch_bounds__(256) k_ring_sides(uint8_t *U, long long plane_stride, int pitch, int Yd, int Xd,
                                                    int ring, int b, int padh) {
  uint8_t *P = U + (long long)blockIdx.z * plane_stride;
  const int per_row = ring >> 3, total = (Yd + 2 * ring) * per_row;  // groups of 16 cells (ring % 16 == 0)
  const int y0 = Yd - b;  // the first row whose pointer the reference's alloc does not shift (heap alias, A.3)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / per_row, g = i - r * per_row, y = r - ring;
    // the first ring / 16 groups on the left, the others on the right
    const bool left = g < (ring >> 4);
    const int x0 = left ? 16 * g - ring : Xd + 16 * (g - (ring >> 4));
    // Along a row the rule is constant on either side of the picture, except in the one row where the
    // heap alias (left, x < -padh) or its mirror case (b == Yd: row -1, right, x >= Xd + padh) applies.
    const bool quirk = b > padh && (left ? (y0 != 0 && y == y0) : (y0 == 0 && y == -1));
    unsigned w[4];
    if (!quirk) {
      const int cy = min(max(y, 0), Yd - 1);
      const unsigned v = left ? (y >= Yd ? P[(long long)(Yd - 1) * pitch + Xd - 1] : P[(long long)cy * pitch])
                              : P[(long long)cy * pitch + Xd - 1];
      w[0] = w[1] = w[2] = w[3] = v * 0x01010101u;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        w[j] = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
          w[j] |= (unsigned)bordered_ref_u8(P, pitch, Yd, Xd, b, padh, y, x0 + 4 * j + k) << (8 * k);
      }
    }
    *reinterpret_cast<uint4 *>(P + (long long)y * pitch + x0) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

void launch_fill_ring(const Launch &L, uint8_t *U, long long plane_stride, int pitch, int nplanes, int Yd, int Xd,
                      int ring, int b, int padh) {
  if (nplanes <= 0 || ring <= 0) return;
  auto blocks_for = [](long long items) { return (int)std::max<long long>(1, std::min<long long>((items + 1023) / 1024, 64)); };
  for (int z0 = 0; z0 < nplanes; z0 += 65535) {
    const int nz = nplanes - z0 < 65535 ? nplanes - z0 : 65535;
    uint8_t *Uz = U + (long long)z0 * plane_stride;
    {
      ProfScope ps_(L, KC_IMG);
      k_ring_bands<<<dim3(blocks_for(2LL * ring * (Xd >> 3)), 1, nz), 256, 0, L.stream>>>(Uz, plane_stride, pitch, Yd, Xd, ring);
      COUNT(L);
    }
    ProfScope ps_(L, KC_IMG);
    k_ring_sides<<<dim3(blocks_for((long long)(Yd + 2 * ring) * (ring >> 3)), 1, nz), 256, 0, L.stream>>>(
        Uz, plane_stride, pitch, Yd, Xd, ring, b, padh);
    COUNT(L);
  }
}

__device__ __forceinline__ unsigned load_u32_unaligned(const uint8_t *p) {
  const uintptr_t a = (uintptr_t)p;
  const unsigned *w = reinterpret_cast<const unsigned *>(a & ~(uintptr_t)3);
  return __funnelshift_r(w[0], w[1], 8 * (int)(a & 3));
}

// grid (BX, BY, pairs * 3); one CTA = one (block, component) of bsa x bsa samples.
__global__ void __launch_bounds__(256) k_predict_u8(PredU8Params q) {
  const int bx = blockIdx.x, by = q.by0 + blockIdx.y;
  const int pair = blockIdx.z / 3, c = blockIdx.z % 3;
  const long long plane = (long long)q.BY * q.BX;
  const short *mv = q.mv + (long long)pair * 4 * plane + (long long)by * q.BX + bx;
  const int mx0 = mv[MV_PREV_X * plane], my0 = mv[MV_PREV_Y * plane];
  const int mx1 = mv[MV_NEXT_X * plane], my1 = mv[MV_NEXT_Y * plane];
  const uint8_t *V0 = q.v + ((long long)(q.f0 + pair) * 3 + c) * q.v_plane_stride;
  const uint8_t *V1 = q.v + ((long long)(q.f0 + pair + 1) * 3 + c) * q.v_plane_stride;
  uint8_t *P = q.p + ((long long)pair * 3 + c) * q.p_plane_stride;
  const int y0 = by * q.bsa, x0 = bx * q.bsa;
  const int wpr = q.bsa >> 2;
  auto inside = [&](int my, int mx) {
    return y0 + my >= 0 && y0 + my + q.bsa <= q.Ya && x0 + mx >= 0 && x0 + mx + q.bsa <= q.Xa;
  };
  if (inside(my0, mx0) && inside(my1, mx1)) {
    // thread = (row within a pass, word column); pointers advance by whole passes so the
    // loop body is two unaligned word loads, one halving add and one store
    const int wl = 31 - __clz(wpr);  // wpr is a power of two here (checked by the launcher)
    const int w = threadIdx.x & (wpr - 1), r0 = threadIdx.x >> wl, rstep = blockDim.x >> wl;
    const uint8_t *a = V0 + (long long)(y0 + r0 + my0) * q.v_pitch + x0 + 4 * w + mx0;
    const uint8_t *b = V1 + (long long)(y0 + r0 + my1) * q.v_pitch + x0 + 4 * w + mx1;
    uint8_t *o = P + (long long)(y0 + r0) * q.p_pitch + x0 + 4 * w;
    const long long vs = (long long)rstep * q.v_pitch, ps = (long long)rstep * q.p_pitch;
    const int sa = 8 * (int)((uintptr_t)a & 3), sb = 8 * (int)((uintptr_t)b & 3);
    const unsigned *a4 = reinterpret_cast<const unsigned *>((uintptr_t)a & ~(uintptr_t)3);
    const unsigned *b4 = reinterpret_cast<const unsigned *>((uintptr_t)b & ~(uintptr_t)3);
    const long long vs4 = vs >> 2;  // pitches are multiples of 16 bytes
    for (int r = r0; r < q.bsa; r += rstep) {
      unsigned va = __funnelshift_r(a4[0], a4[1], sa);
      unsigned vb = __funnelshift_r(b4[0], b4[1], sb);
      // (r0 + r1) / 2 of bytes; the [0,255] clip of decorrelate.cpp:841-848 is a no-op
      *reinterpret_cast<unsigned *>(o) = __vhaddu4(va, vb);
      a4 += vs4;
      b4 += vs4;
      o += ps;
    }
  } else {
    for (int i = threadIdx.x; i < q.bsa * q.bsa; i += blockDim.x) {
      const int y = y0 + i / q.bsa, x = x0 + i % q.bsa;
      int a = bordered_ref_u8(V0, q.v_pitch, q.Ya, q.Xa, q.ba, q.padh, y + my0, x + mx0);
      int b = bordered_ref_u8(V1, q.v_pitch, q.Ya, q.Xa, q.ba, q.padh, y + my1, x + mx1);
      P[(long long)y * q.p_pitch + x] = (uint8_t)((a + b) >> 1);
    }
  }
}

void launch_predict_u8(const Launch &L, const PredU8Params &q, int npairs) {
  if (npairs <= 0 || q.BY <= 0 || q.BX <= 0) return;
  dim3 grid(q.BX, q.nby > 0 ? q.nby : q.BY, npairs * 3);
  ProfScope ps_(L, KC_PREDICT);
  const int wpr = q.bsa >> 2;
  int threads = q.bsa >= 32 ? 256 : 64;
  if (wpr > threads) threads = wpr <= 1024 ? wpr : 1024;
  k_predict_u8<<<grid, threads, 0, L.stream>>>(q);
  COUNT(L);
}

// ---- LL-only integer 5/3 analysis (5_3.cpp:39-52, even lengths) ----
// l[m] of a line of n samples reached through S(g), g = global sample index.
template <typename F>
__device__ __forceinline__ int ll53(F S, int m, int n) {
  auto H = [&](int i) -> int {
    return (i == (n >> 1) - 1) ? (short)(S(n - 1) - S(n - 2))
                               : (short)(S(2 * i + 1) - tdiv2(S(2 * i) + S(2 * i + 2)));
  };
  if (m == 0) return (short)(S(0) + tdiv2(H(0)));
  return (short)(S(2 * m) + tdiv4(H(m) + H(m - 1)));
}
template <typename F>
__device__ __forceinline__ int hh53(F S, int i, int n) {
  return (i == (n >> 1) - 1) ? (short)(S(n - 1) - S(n - 2))
                             : (short)(S(2 * i + 1) - tdiv2(S(2 * i) + S(2 * i + 2)));
}

// ---- chained tail rows (Y % block_size != 0) ----
// After pair i the reference's prediction buffer holds, in rows [cy, Ya), the
// column-pass high band of the first analysis level of the rows [Ya - 2R, Ya)
// (R = Ya - cy): rows [Ya - 2R, cy) come from predict(), rows [cy, Ya) are the
// previous state clipped to [0,255].  This kernel computes that state for one pair
// and stores it, clipped, as the uncovered rows of the next pair's P_a planes.
// grid (ceil(Xa/256), R, 3)
__global__ void __launch_bounds__(256) k_tail_state(const uint8_t *P, long long plane_stride,
                                                    int pitch, uint8_t *Pnext, int Ya, int Xa,
                                                    int cy, int first_comp) {
  const int c = first_comp + blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= Xa) return;
  const int row = cy + blockIdx.y;  // buffer row that receives h_col[row - Ya/2]
  const uint8_t *Pc = P + (long long)c * plane_stride;
  // sample (r, x) of the row-analysed image: [lows | highs] of row r
  auto rowpass = [&](int r) -> int {
    const uint8_t *line = Pc + (long long)r * pitch;
    auto S = [&](int g) { return (int)line[g]; };
    return x < (Xa >> 1) ? ll53(S, x, Xa) : hh53(S, x - (Xa >> 1), Xa);
  };
  const int v = hh53(rowpass, row - (Ya >> 1), Ya);
  const int clipped = v < 0 ? 0 : (v > 255 ? 255 : v);
  Pnext[(long long)c * plane_stride + (long long)row * pitch + x] = (uint8_t)clipped;
}

void launch_tail_state(const Launch &L, const uint8_t *P, long long plane_stride, int pitch,
                       uint8_t *Pnext, int Ya, int Xa, int cy, int first_comp, int ncomp) {
  if (cy >= Ya || ncomp <= 0) return;
  dim3 grid((Xa + 255) / 256, Ya - cy, ncomp);
  ProfScope ps_(L, KC_PREDICT);
  k_tail_state<<<grid, 256, 0, L.stream>>>(P, plane_stride, pitch, Pnext, Ya, Xa, cy, first_comp);
  COUNT(L);
}

// copies `rows` rows of `width` bytes between pitched planes (3 components)
__global__ void k_copy_rows(const uint8_t *src, uint8_t *dst, long long plane_stride, int pitch,
                            int row0, int rows, int width) {
  const int c = blockIdx.z;
  for (int r = blockIdx.y; r < rows; r += gridDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < width; x += gridDim.x * blockDim.x)
      dst[(long long)c * plane_stride + (long long)(row0 + r) * pitch + x] =
          src[(long long)c * plane_stride + (long long)(row0 + r) * pitch + x];
}

void launch_copy_rows(const Launch &L, const uint8_t *src, uint8_t *dst, long long plane_stride,
                      int pitch, int row0, int rows, int width) {
  if (rows <= 0) return;
  dim3 grid((width + 255) / 256, rows, 3);
  ProfScope ps_(L, KC_IMG);
  k_copy_rows<<<grid, 256, 0, L.stream>>>(src, dst, plane_stride, pitch, row0, rows, width);
  COUNT(L);
}
