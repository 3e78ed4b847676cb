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
//   LL   = LL-only multi-level integer 5/3 analysis of P_a, tile by tile in shared
//          memory, fused with the residue / reconstruction and the I/B histograms
//          (k_ll_residue).
// Rows of P_a below the last whole block (Y % block_size != 0) are never written by
// predict(); the reference keeps there what the previous pair's in-place analysis
// left.  k_tail_state reproduces that chain: it is the only
// sequential step and touches 2*(Ya - cy) rows per pair.
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

// Flattened 2-D loop over h x w elements by 256 threads without a division per element.
#define LL_NT 512
#define FOR_2D(r, c, h, w)                                                                      \
  for (int i_ = threadIdx.x, r = i_ / (w), c = i_ - r * (w), dc_ = LL_NT % (w), dr_ = LL_NT / (w); \
       i_ < (h) * (w); i_ += LL_NT, c += dc_, r += dr_, r += (c >= (w)), c -= (c >= (w)) ? (w) : 0)

// l[m] from a line reached through s[g * st] (g = global index minus `base`), line
// length n.  Branch-free: out-of-line taps of the first / last sample are re-pointed at
// valid cells and masked by selects.
template <typename T>
__device__ __forceinline__ int ll53_line(const T *s, int st, int base, int m, int n) {
  const T *p = s + (2 * m - base) * st;
  const bool last = (m == (n >> 1) - 1), first = (m == 0);
  const int s0 = p[0], s1 = p[st];
  const int s2 = p[last ? st : 2 * st];
  const int sm1 = p[first ? 0 : -st], sm2 = p[first ? 0 : -2 * st];
  const int hm = (short)(s1 - (last ? s0 : tdiv2(s0 + s2)));
  const int hm1 = (short)(sm1 - tdiv2(sm2 + s0));
  const int hs = first ? 2 * hm : hm + hm1;  // l[0] = s[0] + h[0]/2 == s[0] + (2 h[0])/4
  return (short)(s0 + tdiv4(hs));
}

// Sliding LL pass along one line: outputs l[m0 .. m0+cnt) of a line of n samples read
// through src[(g - base) * sst] (g = global sample index) into dst[k * dst_st].  Each step
// loads two new samples and reuses s[2m] and h[m-1] from the previous step.
__device__ __forceinline__ void ll53_slide(const short *src, int sst, int base, int n, int m0, int cnt,
                                           short *dst, int dst_st) {
  if (cnt <= 0) return;
  const int half = n >> 1;
  const short *p = src + (2 * m0 - base) * sst;
  int s0 = p[0];
  int hprev = 0;
  if (m0 > 0) hprev = (short)(p[-sst] - tdiv2(p[-2 * sst] + s0));
  for (int k = 0; k < cnt; k++) {
    const int m = m0 + k;
    const int s1 = p[sst];
    int hm, s2 = 0;
    if (m == half - 1) {
      hm = (short)(s1 - s0);
    } else {
      s2 = p[2 * sst];
      hm = (short)(s1 - tdiv2(s0 + s2));
    }
    const int hs = (m == 0) ? 2 * hm : hm + hprev;
    dst[k * dst_st] = (short)(s0 + tdiv4(hs));
    hprev = hm;
    s0 = s2;
    p += 2 * sst;
  }
}

// Level 0 -> 1 row pass on bytes: one thread produces the four outputs m = 4g+1 .. 4g+4
// of one tile row from the 11 bytes [8g, 8g+11) held in three aligned words.
__device__ __forceinline__ void ll53_row_u8x4(const unsigned *row_words, int xs, int nwords, int g,
                                              int n, int mx0, int mx1, short *dst /* row, index m - mx0 */) {
  const int half = n >> 1;
  // word index of byte 8g inside the staged row (xs is a multiple of 8)
  const int wi = (8 * g - xs) >> 2;
  unsigned w[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int idx = wi + k;
    idx = idx < 0 ? 0 : (idx >= nwords ? nwords - 1 : idx);
    w[k] = row_words[idx];
  }
  int s[11];
#pragma unroll
  for (int k = 0; k < 11; k++) s[k] = (w[k >> 2] >> (8 * (k & 3))) & 0xff;
  // h_rel[j] = h[4g + j], j = 0..4 (bytes are non-negative: /2 is a shift)
  int h[5];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const int gi = 4 * g + j;
    h[j] = (gi == half - 1) ? s[2 * j + 1] - s[2 * j] : s[2 * j + 1] - ((s[2 * j] + s[2 * j + 2]) >> 1);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int m = 4 * g + 1 + i;
    if (m >= mx0 && m < mx1 && m > 0) {
      const int hs = h[i + 1] + h[i];  // m >= 1 here: l[m] = s[2m] + (h[m] + h[m-1]) / 4
      dst[m - mx0] = (short)(s[2 * i + 2] + tdiv4(hs));
    }
  }
}

// Output tile [oy0,oy0+T) x [ox0,ox0+T) of the level-nlev LL band of one P_a plane,
// then residue (analysis) or reconstruction (synthesis) of that tile.
// grid (ceil(X/32), ceil(Y/32), pairs * 3)
__global__ void __launch_bounds__(LL_NT) k_ll_residue(LLParams q) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ int h_pred[256], h_res[256];
  const int pair = blockIdx.z / 3, c = blockIdx.z % 3;
  const int nlev = q.a + (c ? 1 : 0);
  const int T = nlev >= 3 ? 16 : 32;
  const int OW = c ? q.X / 2 : q.X, OH = c ? q.Y / 2 : q.Y;  // component size = LL size
  const int ox0 = blockIdx.x * T, oy0 = blockIdx.y * T;
  if (ox0 >= OW || oy0 >= OH) return;
  const int ox1 = min(ox0 + T, OW), oy1 = min(oy0 + T, OH);
  const uint8_t *P = q.p + ((long long)pair * 3 + c) * q.p_plane_stride;
  const long long coff = c == 0 ? 0 : (long long)q.X * q.Y + (long long)(c - 1) * (q.X / 2) * (q.Y / 2);
  const bool do_hist = q.hist && c == 0 && !q.synth;
  if (do_hist) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h_pred[i] = h_res[i] = 0;
  }
  // required index ranges per level, top-down: l[m] needs s[2m-2 .. 2m+2]
  int ry0[4], ry1[4], rx0[4], rx1[4];
  ry0[nlev] = oy0; ry1[nlev] = oy1; rx0[nlev] = ox0; rx1[nlev] = ox1;
  for (int k = nlev; k > 0; k--) {
    const int ny = (q.Y << q.a) >> (k - 1), nx = (q.X << q.a) >> (k - 1);
    ry0[k - 1] = max(0, 2 * ry0[k] - 2); ry1[k - 1] = min(ny, 2 * ry1[k] + 1);
    rx0[k - 1] = max(0, 2 * rx0[k] - 2); rx1[k - 1] = min(nx, 2 * rx1[k] + 1);
  }
  short *A = reinterpret_cast<short *>(smraw);  // row-pass output
  short *B = A + q.smem_a;                      // column-pass output (level image)
  const short *LL = nullptr;
  int ll_w = 0;
  if (nlev > 0) {
    // level-0 bytes are read straight from the plane (aligned 32-bit loads through L1)
    const int xs = rx0[0] & ~7;
    const int uw = (rx1[0] - xs + 3) >> 2;  // words per tile row
    // level 0 -> 1
    {
      const int h0 = ry1[0] - ry0[0], w1 = rx1[1] - rx0[1], h1 = ry1[1] - ry0[1];
      const int nx = q.X << q.a, ny = q.Y << q.a;
      const int mx0 = rx0[1], my0 = ry0[1];
      {
        // row pass: groups of four outputs m = 4g+1 .. 4g+4; m = 0 (first column of the
        // picture) is outside every group and handled by the generic tap code
        const int mx1 = rx1[1];
        const int g_lo = (max(mx0, 1) - 1) >> 2, g_hi = (mx1 - 2) >> 2;
        const int ng = g_hi - g_lo + 1;
        FOR_2D(r, gg, h0, ng)
          ll53_row_u8x4(reinterpret_cast<const unsigned *>(P + (long long)(ry0[0] + r) * q.p_pitch + xs), xs, uw,
                        g_lo + gg, nx, mx0, mx1, A + r * w1);
        if (mx0 == 0)
          for (int r = threadIdx.x; r < h0; r += LL_NT)
            A[r * w1] = (short)ll53_line(P + (long long)(ry0[0] + r) * q.p_pitch, 1, 0, 0, nx);
      }
      __syncthreads();
      {
        // column pass: thread = (column, row segment), sliding down the column
        const int nseg = max(1, LL_NT / w1), per = (h1 + nseg - 1) / nseg;
        const int col = threadIdx.x % w1, seg = threadIdx.x / w1;
        if (seg < nseg) {
          const int k0 = seg * per, cnt = min(per, h1 - k0);
          ll53_slide(A + col, w1, ry0[0], ny, my0 + k0, cnt, B + k0 * w1 + col, w1);
        }
      }
      __syncthreads();
    }
    for (int k = 1; k < nlev; k++) {
      const int hk = ry1[k] - ry0[k], wk = rx1[k] - rx0[k];
      const int w1 = rx1[k + 1] - rx0[k + 1], h1 = ry1[k + 1] - ry0[k + 1];
      const int nx = (q.X << q.a) >> k, ny = (q.Y << q.a) >> k;
      const int mx0 = rx0[k + 1], my0 = ry0[k + 1], bx0 = rx0[k], by0 = ry0[k];
      {
        // row pass: thread = (row, column segment), sliding along the row
        const int nseg = max(1, LL_NT / hk), per = (w1 + nseg - 1) / nseg;
        const int row = threadIdx.x % hk, seg = threadIdx.x / hk;
        if (seg < nseg) {
          const int k0 = seg * per, cnt = min(per, w1 - k0);
          ll53_slide(B + row * wk, 1, bx0, nx, mx0 + k0, cnt, A + row * w1 + k0, 1);
        }
      }
      __syncthreads();
      {
        const int nseg = max(1, LL_NT / w1), per = (h1 + nseg - 1) / nseg;
        const int col = threadIdx.x % w1, seg = threadIdx.x / w1;
        if (seg < nseg) {
          const int k0 = seg * per, cnt = min(per, h1 - k0);
          ll53_slide(A + col, w1, by0, ny, my0 + k0, cnt, B + k0 * w1 + col, w1);
        }
      }
      __syncthreads();
    }
    LL = B;
    ll_w = ox1 - ox0;
  } else if (do_hist) {
    __syncthreads();
  }
  const uint8_t *in = q.in + (long long)pair * q.in_stride + coff;
  uint8_t *out = q.out + (long long)pair * q.out_stride + coff;
  uint8_t *pout = q.prediction ? q.prediction + (long long)pair * q.pred_stride + coff : nullptr;
  const int is_I = q.synth && q.types[pair] == 'I';
  const int tw = ox1 - ox0, th = oy1 - oy0;
  FOR_2D(r, cc, th, tw) {
    const int y = oy0 + r, x = ox0 + cc;
    const int p = LL ? (int)LL[r * ll_w + cc] : (int)P[(long long)y * q.p_pitch + x];
    const int s = in[(long long)y * OW + x];
    int o;
    if (!q.synth) {
      int rr = s - p;
      rr = rr < -128 ? -128 : (rr > 127 ? 127 : rr);
      o = rr + 128;
      if (do_hist) {
        atomicAdd(&h_pred[s], 1);
        atomicAdd(&h_res[o], 1);
      }
    } else if (is_I) {
      o = s;
    } else {
      o = s - 128 + p;
      o = o < 0 ? 0 : (o > 255 ? 255 : o);
    }
    out[(long long)y * OW + x] = (uint8_t)o;
    if (pout) pout[(long long)y * OW + x] = (uint8_t)p;
  }
  if (do_hist) {
    __syncthreads();
    int *hist = q.hist + (long long)pair * q.hist_stride;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
      if (h_pred[i]) atomicAdd(&hist[i], h_pred[i]);
      if (h_res[i]) atomicAdd(&hist[256 + i], h_res[i]);
    }
  }
}

// shared-memory footprint (shorts) of the row-pass (A) and level (B) buffers
static void ll_smem(int a, int *sa, int *sb, int *su) {
  int A = 0, B = 0, U = 0;
  for (int c = 0; c < 2; c++) {
    int nlev = a + c;
    if (nlev == 0) continue;
    int T = nlev >= 3 ? 16 : 32;
    int ext[4];
    ext[nlev] = T;
    for (int k = nlev; k > 0; k--) ext[k - 1] = 2 * ext[k] + 3;
    U = U > ext[0] * (ext[0] + 12) ? U : ext[0] * (ext[0] + 12);
    for (int k = 0; k < nlev; k++) {
      A = A > ext[k] * ext[k + 1] ? A : ext[k] * ext[k + 1];
      B = B > ext[k + 1] * ext[k + 1] ? B : ext[k + 1] * ext[k + 1];
    }
  }
  *sa = (A + 7) & ~7;
  *sb = (B + 7) & ~7;
  *su = (U + 15) & ~15;
}

void launch_ll_residue(const Launch &L, LLParams q, int npairs) {
  if (npairs <= 0) return;
  int sa, sb, su;
  ll_smem(q.a, &sa, &sb, &su);
  q.smem_a = sa;
  q.smem_b = sb;
  size_t smem = (size_t)(sa + sb) * sizeof(short);
  (void)su;
  static size_t s_attr = 0;
  if (smem > 48 * 1024 && smem > s_attr) {
    cudaFuncSetAttribute(k_ll_residue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    s_attr = smem;
  }
  dim3 grid((q.X + 31) / 32, (q.Y + 31) / 32, npairs * 3);
  ProfScope ps_(L, KC_RESIDUE);
  k_ll_residue<<<grid, LL_NT, smem, L.stream>>>(q);
  COUNT(L);
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
