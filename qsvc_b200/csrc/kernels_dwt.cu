// kernels_dwt.cu -- in-place 2-D integer 5/3 transform in the reference's
// Mallat layout (dwt2d.cpp:76-175 over 5_3.cpp:39-115).
//
// Analysis of one level: rows first, then columns; synthesis: columns first,
// then rows.  Each line is staged in shared memory, so the in-place
// de-interleave ([lows | highs]) has no read/write hazard; a column CTA owns a
// strip of CW adjacent columns (32-byte or 64-byte row segments) over all rows.
#include "kernels.cuh"

#define COUNT(L) (++*(L).counter)

static const int kMaxDynSmem = 200 * 1024;

// One line of the lifting, "pair" formulation: task i produces l[i] and h[i] (analysis)
// or s[2i] and s[2i+1] (synthesis) of an n-sample line held in shared memory with
// stride st.  half = n/2 pairs, plus the unpaired last low sample when n is odd.
__device__ __forceinline__ void ana_pair(const short *s, int st, int i, int n, int &l, int &h) {
  const int half = n >> 1;
  const int s0 = s[(2 * i) * st], s1 = s[(2 * i + 1) * st];
  const bool last_even = !(n & 1) && i == half - 1;
  h = (short)(last_even ? s1 - s0 : s1 - tdiv2(s0 + s[(2 * i + 2) * st]));
  if (i == 0) {
    l = (short)(s0 + tdiv2(h));
  } else {
    const int hp = (short)(s[(2 * i - 1) * st] - tdiv2(s[(2 * i - 2) * st] + s0));
    l = (short)(s0 + tdiv4(h + hp));
  }
}
// even sample e(i) of the synthesis; lo = lows, hi = highs (stride st)
__device__ __forceinline__ int syn_e(const short *lo, const short *hi, int st, int i, int n) {
  const int half = n >> 1;
  if (i == 0) return (short)(lo[0] - tdiv2(hi[0]));
  if (i < half) return (short)(lo[i * st] - tdiv4(hi[i * st] + hi[(i - 1) * st]));
  return (short)(lo[half * st] - tdiv2(hi[(half - 1) * st]));
}

// Row pass.  One CTA per row at a time: the row is staged in shared memory (128-bit
// loads when aligned), thread i produces one output pair and stores it straight to
// the row (coalesced 16-bit stores for analysis, one 32-bit store for synthesis).
template <bool SYNTH>
__global__ void __launch_bounds__(256) k_dwt_rows(Plane p, int slot0, int ny, int nx) {
  extern __shared__ __align__(16) short sm[];
  const int slot = slot0 + blockIdx.z;
  const int nvec = nx >> 3;
  const int half = nx >> 1, nlow = nx - half;
  for (int y = blockIdx.x; y < ny; y += gridDim.x) {
    short *row = p.row(slot, y);
    if ((((uintptr_t)row) & 15) == 0) {
      const uint4 *r4 = reinterpret_cast<const uint4 *>(row);
      uint4 *s4 = reinterpret_cast<uint4 *>(sm);
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) s4[i] = r4[i];
      for (int i = (nvec << 3) + threadIdx.x; i < nx; i += blockDim.x) sm[i] = row[i];
    } else {
      for (int i = threadIdx.x; i < nx; i += blockDim.x) sm[i] = row[i];
    }
    __syncthreads();
    if (!SYNTH) {
      if (!(nx & 3) && (((uintptr_t)row) & 3) == 0) {
        // two output pairs per thread: the high-pass value in the middle is shared, the stores are 32-bit
        unsigned *lo2 = reinterpret_cast<unsigned *>(row), *hi2 = reinterpret_cast<unsigned *>(row + nlow);
        for (int k = threadIdx.x; k < (half >> 1); k += blockDim.x) {
          const int i = 2 * k;
          const short *s = sm + 2 * i;  // s[0..4]: samples 2i .. 2i+4
          const int s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3];
          const int h0 = (short)(s1 - tdiv2(s0 + s2));
          const int h1 = (short)(i + 1 == half - 1 ? s3 - s2 : s3 - tdiv2(s2 + s[4]));
          int l0;
          if (i == 0) {
            l0 = (short)(s0 + tdiv2(h0));
          } else {
            const int hp = (short)(s[-1] - tdiv2(s[-2] + s0));
            l0 = (short)(s0 + tdiv4(h0 + hp));
          }
          const int l1 = (short)(s2 + tdiv4(h1 + h0));
          lo2[k] = ((unsigned)(unsigned short)l0) | ((unsigned)(unsigned short)l1 << 16);
          hi2[k] = ((unsigned)(unsigned short)h0) | ((unsigned)(unsigned short)h1 << 16);
        }
      } else {
        for (int i = threadIdx.x; i < half; i += blockDim.x) {
          int l, h;
          ana_pair(sm, 1, i, nx, l, h);
          row[i] = (short)l;
          row[nlow + i] = (short)h;
          if ((nx & 1) && i == half - 1) row[half] = (short)(sm[nx - 1] + tdiv2(h));
        }
      }
    } else {
      const short *lo = sm, *hi = sm + nlow;
      const bool al4 = (((uintptr_t)row) & 3) == 0;
      for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const int e0 = syn_e(lo, hi, 1, i, nx);
        int o;
        if (!(nx & 1) && i == half - 1) o = (short)(hi[i] + e0);
        else o = (short)(hi[i] + tdiv2(e0 + syn_e(lo, hi, 1, i + 1, nx)));
        if (al4) {
          reinterpret_cast<unsigned *>(row)[i] = ((unsigned)(unsigned short)e0) | ((unsigned)(unsigned short)o << 16);
        } else {
          row[2 * i] = (short)e0;
          row[2 * i + 1] = (short)o;
        }
        if ((nx & 1) && i == half - 1) row[nx - 1] = (short)syn_e(lo, hi, 1, half, nx);
      }
    }
    __syncthreads();
  }
}

// Column pass: the CTA stages a strip of CW = 2^cw_log2 columns x ny rows in shared
// memory (in-place de-interleave needs the whole column), then thread (column, row
// segment) slides down its segment reusing the previous step's samples.
template <bool SYNTH>
__global__ void __launch_bounds__(1024) k_dwt_cols(Plane p, int slot0, int ny, int nx, int cw_log2) {
  extern __shared__ __align__(16) short sm[];
  const int slot = slot0 + blockIdx.z;
  const int CW = 1 << cw_log2;
  const int x0 = blockIdx.x << cw_log2;
  const int cw = min(CW, nx - x0);
  if (cw_log2 >= 3) {
    const int upr_log2 = cw_log2 - 3;  // 8-column units per row
    const int units = ny << upr_log2;
    for (int u = threadIdx.x; u < units; u += blockDim.x) {
      int y = u >> upr_log2, c0 = (u & ((1 << upr_log2) - 1)) << 3;
      if (c0 >= cw) continue;
      const short *src = p.row(slot, y) + x0 + c0;
      short *dst = sm + (y << cw_log2) + c0;
      if (c0 + 8 <= cw && (((uintptr_t)src) & 15) == 0) {
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src);
      } else {
        for (int k = 0; k < 8 && c0 + k < cw; k++) dst[k] = src[k];
      }
    }
  } else {
    const int total = ny << cw_log2;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      int y = idx >> cw_log2, c = idx & (CW - 1);
      if (c < cw) sm[idx] = p.row(slot, y)[x0 + c];
    }
  }
  __syncthreads();
  const int c = threadIdx.x & (CW - 1), seg = threadIdx.x >> cw_log2, nseg = blockDim.x >> cw_log2;
  if (c >= cw) return;
  const int half = ny >> 1, nlow = ny - half;
  const int per = (half + nseg - 1) / nseg;
  const int i0 = seg * per, i1 = min(half, i0 + per);
  const short *s = sm + c;
  // one row pointer walk per output stream
  if (!SYNTH) {
    // sliding form: h[i-1] and the sample s[2i] carry over from the previous output pair
    if (i0 < i1) {
      int s0 = s[(2 * i0) * CW];
      int hp = i0 > 0 ? (short)(s[(2 * i0 - 1) * CW] - tdiv2(s[(2 * i0 - 2) * CW] + s0)) : 0;
      for (int i = i0; i < i1; i++) {
        const int s1 = s[(2 * i + 1) * CW];
        const bool last_even = !(ny & 1) && i == half - 1;
        const int s2 = last_even ? 0 : s[(2 * i + 2) * CW];
        const int h = (short)(last_even ? s1 - s0 : s1 - tdiv2(s0 + s2));
        const int l = (short)(i == 0 ? s0 + tdiv2(h) : s0 + tdiv4(h + hp));
        p.row(slot, i)[x0 + c] = (short)l;
        p.row(slot, nlow + i)[x0 + c] = (short)h;
        if ((ny & 1) && i == half - 1) p.row(slot, half)[x0 + c] = (short)(s[(ny - 1) * CW] + tdiv2(h));
        hp = h;
        s0 = s2;
      }
    }
  } else {
    const short *lo = s, *hi = s + (long long)nlow * CW;
    int e0 = i0 < i1 ? syn_e(lo, hi, CW, i0, ny) : 0;
    for (int i = i0; i < i1; i++) {
      int o, e1 = 0;
      if (!(ny & 1) && i == half - 1) {
        o = (short)(hi[i * CW] + e0);
      } else {
        e1 = syn_e(lo, hi, CW, i + 1, ny);
        o = (short)(hi[i * CW] + tdiv2(e0 + e1));
      }
      p.row(slot, 2 * i)[x0 + c] = (short)e0;
      p.row(slot, 2 * i + 1)[x0 + c] = (short)o;
      if ((ny & 1) && i == half - 1) p.row(slot, ny - 1)[x0 + c] = (short)e1;
      e0 = e1;
    }
  }
}

// ---- first pyramid level straight from the frames ----
// One level of dwt2d::analyze (rows, then all columns; dwt2d.cpp:76-119, 5_3.cpp:39-52, even
// sizes) of the luma of frame f0 + z, read as bytes and written as the four sub-bands of the
// in-place Mallat layout of slot slot0 + z.  Replaces load + row pass + column pass of the level (and the
// snapshot the descent would restore it from: the frame itself is that snapshot).
// One thread owns coefficient column gi (pixel columns 2 gi - 2 .. 2 gi + 2) and walks down D0_SEG coefficient
// rows: the row pass of a pixel row at that column is five bytes (two 16-bit loads and a byte; neighbouring
// threads share them through L1) and yields one low and one high sample, each of which runs through its own
// streaming column lifting (last even row and last high-pass value in registers); four 16-bit stores per step,
// coalesced across the warp.  No shared memory, no barriers.
static constexpr int D0_SEG = 32;
__global__ void __launch_bounds__(128) k_dwt0_u8(Plane p, int slot0, const uint8_t *__restrict__ src,
                                                  long long frame_stride, int f0, int Y, int X) {
  const int halfx = X >> 1, halfy = Y >> 1;
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= halfx) return;
  const int slot = slot0 + blockIdx.z;
  const uint8_t *frame = src + (long long)(f0 + blockIdx.z) * frame_stride + 2 * gi;
  const int j0 = blockIdx.y * D0_SEG, j1 = min(j0 + D0_SEG, halfy);
  // low (l) and high (h) sample gi of pixel row y
  auto rowpass = [&](int y, int &l, int &h) {
    const uint8_t *r = frame + (unsigned)(y * X);
    const unsigned c = *reinterpret_cast<const unsigned short *>(r);
    const int s0 = c & 0xff, s1 = c >> 8;
    h = gi == halfx - 1 ? s1 - s0 : s1 - tdiv2(s0 + (int)r[2]);
    if (gi == 0) {
      l = s0 + tdiv2(h);
    } else {
      const unsigned q = *reinterpret_cast<const unsigned short *>(r - 2);
      const int hp = (int)(q >> 8) - tdiv2((int)(q & 0xff) + s0);
      l = s0 + tdiv4(h + hp);
    }
  };
  int el, eh, hpl = 0, hph = 0;  // per band: last even row, last column high-pass value
  rowpass(2 * j0, el, eh);
  if (j0 > 0) {
    int al, ah, bl, bh;
    rowpass(2 * j0 - 1, al, ah);
    rowpass(2 * j0 - 2, bl, bh);
    hpl = (short)(al - tdiv2(bl + el));
    hph = (short)(ah - tdiv2(bh + eh));
  }
  for (int j = j0; j < j1; j++) {
    int ol, oh, nl = 0, nh = 0;
    rowpass(2 * j + 1, ol, oh);
    int hl, hh;
    if (j == halfy - 1) {
      hl = (short)(ol - el);
      hh = (short)(oh - eh);
    } else {
      rowpass(2 * j + 2, nl, nh);
      hl = (short)(ol - tdiv2(el + nl));
      hh = (short)(oh - tdiv2(eh + nh));
    }
    const int ll = j == 0 ? el + tdiv2(hl) : el + tdiv4(hl + hpl);
    const int lh = j == 0 ? eh + tdiv2(hh) : eh + tdiv4(hh + hph);
    short *top = p.row(slot, j), *bot = p.row(slot, halfy + j);
    top[gi] = (short)ll;
    top[halfx + gi] = (short)lh;
    bot[gi] = (short)hl;
    bot[halfx + gi] = (short)hh;
    hpl = hl;
    hph = hh;
    el = nl;
    eh = nh;
  }
}

void launch_dwt0_u8(const Launch &L, Plane p, int slot0, int nslots, const uint8_t *src, long long frame_stride,
                    int f0, int Y, int X) {
  if (nslots <= 0) return;
  dim3 grid(((X >> 1) + 127) / 128, ((Y >> 1) + D0_SEG - 1) / D0_SEG, nslots);
  ProfScope ps_(L, KC_DWT_ROWS);
  k_dwt0_u8<<<grid, 128, 0, L.stream>>>(p, slot0, src, frame_stride, f0, Y, X);
  COUNT(L);
}

// The same level from an int16 snapshot of the LL region (invertible pyramids: the region is copied out before it
// is transformed, to be restored by the descent): snap (ny x nx, row pitch `pitch`, slot s at snap + s *
// snap_slot_stride) -> the four sub-bands of the region, in place in plane p.  ny, nx even, pitch even.
__global__ void __launch_bounds__(128) k_dwt_snap(Plane p, int slot0, const short *__restrict__ snap,
                                                   long long snap_slot_stride, int pitch, int ny, int nx) {
  const int halfx = nx >> 1, halfy = ny >> 1;
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= halfx) return;
  const int slot = slot0 + blockIdx.z;
  const short *col = snap + (long long)slot * snap_slot_stride + 2 * gi;
  const int j0 = blockIdx.y * D0_SEG, j1 = min(j0 + D0_SEG, halfy);
  auto rowpass = [&](int y, int &l, int &h) {
    const unsigned *w = reinterpret_cast<const unsigned *>(col + (long long)y * pitch);
    const unsigned c = w[0];
    const int s0 = (short)(c & 0xffffu), s1 = (int)c >> 16;
    h = (short)(gi == halfx - 1 ? s1 - s0 : s1 - tdiv2(s0 + (int)(short)(w[1] & 0xffffu)));
    if (gi == 0) {
      l = (short)(s0 + tdiv2(h));
    } else {
      const unsigned q = w[-1];
      const int hp = (short)(((int)q >> 16) - tdiv2((int)(short)(q & 0xffffu) + s0));
      l = (short)(s0 + tdiv4(h + hp));
    }
  };
  int el, eh, hpl = 0, hph = 0;
  rowpass(2 * j0, el, eh);
  if (j0 > 0) {
    int al, ah, bl, bh;
    rowpass(2 * j0 - 1, al, ah);
    rowpass(2 * j0 - 2, bl, bh);
    hpl = (short)(al - tdiv2(bl + el));
    hph = (short)(ah - tdiv2(bh + eh));
  }
  for (int j = j0; j < j1; j++) {
    int ol, oh, nl = 0, nh = 0;
    rowpass(2 * j + 1, ol, oh);
    int hl, hh;
    if (j == halfy - 1) {
      hl = (short)(ol - el);
      hh = (short)(oh - eh);
    } else {
      rowpass(2 * j + 2, nl, nh);
      hl = (short)(ol - tdiv2(el + nl));
      hh = (short)(oh - tdiv2(eh + nh));
    }
    const int ll = (short)(j == 0 ? el + tdiv2(hl) : el + tdiv4(hl + hpl));
    const int lh = (short)(j == 0 ? eh + tdiv2(hh) : eh + tdiv4(hh + hph));
    short *top = p.row(slot, j), *bot = p.row(slot, halfy + j);
    top[gi] = (short)ll;
    top[halfx + gi] = (short)lh;
    bot[gi] = (short)hl;
    bot[halfx + gi] = (short)hh;
    hpl = hl;
    hph = hh;
    el = nl;
    eh = nh;
  }
}

bool dwt_snap_supported(int ny, int nx, int pitch, const short *snap, long long snap_slot_stride) {
  return ny >= 2 && nx >= 2 && !(ny & 1) && !(nx & 1) && !(pitch & 1) && !(snap_slot_stride & 1) &&
         (((uintptr_t)snap) & 3) == 0;
}

void launch_dwt_snap(const Launch &L, Plane p, int slot0, int nslots, const short *snap, long long snap_slot_stride,
                     int pitch, int ny, int nx) {
  if (nslots <= 0) return;
  dim3 grid(((nx >> 1) + 127) / 128, ((ny >> 1) + D0_SEG - 1) / D0_SEG, nslots);
  ProfScope ps_(L, KC_DWT_ROWS);
  k_dwt_snap<<<grid, 128, 0, L.stream>>>(p, slot0, snap, snap_slot_stride, pitch, ny, nx);
  COUNT(L);
}

// One synthesis level (dwt2d.cpp:139-172: columns, then rows; 5_3.cpp:81-94) of the region whose copy is `snap`,
// written in place in plane p.  A lane owns coefficient column gi: it runs the column synthesis of the two columns
// it contributes to a row (gi of the row-low half, gi of the row-high half) as streaming lifting down D0_SEG
// coefficient rows, and the row synthesis of a finished row takes the neighbours' samples by shuffle (30 owner
// lanes per warp + one halo lane on either side).  ny, nx even.
__global__ void __launch_bounds__(128) k_syn_snap(Plane p, int slot0, const short *__restrict__ snap,
                                                   long long snap_slot_stride, int pitch, int ny, int nx) {
  const int hx = nx >> 1, hy = ny >> 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gi = (blockIdx.x * 4 + warp) * 30 + lane - 1;
  const bool owner = lane >= 1 && lane <= 30 && gi < hx;  // gi >= 0 for every lane >= 1
  const int gc = min(max(gi, 0), hx - 1);
  const int slot = slot0 + blockIdx.z;
  const short *base = snap + (long long)slot * snap_slot_stride;
  const short *cA = base + gc, *cB = base + hx + gc;
  const int j0 = blockIdx.y * D0_SEG, j1 = min(j0 + D0_SEG, hy);
  if ((blockIdx.x * 4 + warp) * 30 >= hx) return;  // whole warp beyond the region
  auto ecol = [&](const short *c, int j) -> int {  // even sample 2j of the column synthesis
    const int lo = c[(long long)j * pitch], hi = c[(long long)(hy + j) * pitch];
    if (j == 0) return (short)(lo - tdiv2(hi));
    return (short)(lo - tdiv4(hi + (int)c[(long long)(hy + j - 1) * pitch]));
  };
  const bool al4 = (((uintptr_t)p.row(slot, 0)) & 3) == 0 && !(p.S & 1);
  auto emit = [&](int y, int a, int b) {  // row synthesis of row y from this lane's (low, high) samples
    const int a_next = __shfl_down_sync(0xffffffffu, a, 1), b_next = __shfl_down_sync(0xffffffffu, b, 1);
    const int b_prev = __shfl_up_sync(0xffffffffu, b, 1);
    const int e0 = (short)(gi == 0 ? a - tdiv2(b) : a - tdiv4(b + b_prev));
    int o;
    if (gi == hx - 1) {
      o = (short)(b + e0);
    } else {
      const int e1 = (short)(a_next - tdiv4(b_next + b));
      o = (short)(b + tdiv2(e0 + e1));
    }
    if (owner) {
      short *row = p.row(slot, y) + 2 * gi;
      if (al4) {
        *reinterpret_cast<unsigned *>(row) = ((unsigned)(unsigned short)e0) | ((unsigned)(unsigned short)o << 16);
      } else {
        row[0] = (short)e0;
        row[1] = (short)o;
      }
    }
  };
  int eA = ecol(cA, j0), eB = ecol(cB, j0);
  for (int j = j0; j < j1; j++) {
    const int hiA = cA[(long long)(hy + j) * pitch], hiB = cB[(long long)(hy + j) * pitch];
    int nA = 0, nB = 0, oA, oB;
    if (j == hy - 1) {
      oA = (short)(hiA + eA);
      oB = (short)(hiB + eB);
    } else {
      nA = ecol(cA, j + 1);
      nB = ecol(cB, j + 1);
      oA = (short)(hiA + tdiv2(eA + nA));
      oB = (short)(hiB + tdiv2(eB + nB));
    }
    emit(2 * j, eA, eB);
    emit(2 * j + 1, oA, oB);
    eA = nA;
    eB = nB;
  }
}

void launch_syn_snap(const Launch &L, Plane p, int slot0, int nslots, const short *snap, long long snap_slot_stride,
                     int pitch, int ny, int nx) {
  if (nslots <= 0) return;
  dim3 grid(((nx >> 1) + 119) / 120, ((ny >> 1) + D0_SEG - 1) / D0_SEG, nslots);
  ProfScope ps_(L, KC_DWT_COLS);
  k_syn_snap<<<grid, 128, 0, L.stream>>>(p, slot0, snap, snap_slot_stride, pitch, ny, nx);
  COUNT(L);
}

int dwt_init_attributes() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_dwt_cols<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_dwt_cols<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_dwt_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_dwt_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  return (int)e;
}

static void launch_rows(const Launch &L, Plane p, int slot0, int nslots, int ny, int nx, bool synth) {
  dim3 grid(ny < 2048 ? ny : 2048, 1, nslots);
  size_t smem = (size_t)nx * sizeof(short);
  ProfScope ps_(L, KC_DWT_ROWS);
  if (synth)
    k_dwt_rows<true><<<grid, 256, smem, L.stream>>>(p, slot0, ny, nx);
  else
    k_dwt_rows<false><<<grid, 256, smem, L.stream>>>(p, slot0, ny, nx);
  COUNT(L);
}

static void launch_cols(const Launch &L, Plane p, int slot0, int nslots, int ny, int nx, bool synth) {
  int cw_log2 = 5;
  while (cw_log2 > 0 && ((size_t)ny << cw_log2) * sizeof(short) > (size_t)kMaxDynSmem) cw_log2--;
  // at least two CTAs per SM: a strip is loaded, transformed and stored in phases that only overlap across CTAs
  while (cw_log2 > 4 && ((size_t)ny << cw_log2) * sizeof(short) > (size_t)100 * 1024) cw_log2--;
  // keep at least ~2 CTAs per SM worth of strips when the image is narrow
  while (cw_log2 > 3 && ((nx + (1 << cw_log2) - 1) >> cw_log2) * nslots < 296) cw_log2--;
  dim3 grid((nx + (1 << cw_log2) - 1) >> cw_log2, 1, nslots);
  size_t smem = ((size_t)ny << cw_log2) * sizeof(short);
  int threads = (ny << cw_log2) >= 8192 ? 1024 : 256;
  ProfScope ps_(L, KC_DWT_COLS);
  if (synth)
    k_dwt_cols<true><<<grid, threads, smem, L.stream>>>(p, slot0, ny, nx, cw_log2);
  else
    k_dwt_cols<false><<<grid, threads, smem, L.stream>>>(p, slot0, ny, nx, cw_log2);
  COUNT(L);
}

void launch_dwt_level(const Launch &L, Plane p, int slot0, int nslots, int ny, int nx, bool synth) {
  if (nslots <= 0) return;
  if (synth) {
    launch_cols(L, p, slot0, nslots, ny, nx, true);
    launch_rows(L, p, slot0, nslots, ny, nx, true);
  } else {
    launch_rows(L, p, slot0, nslots, ny, nx, false);
    launch_cols(L, p, slot0, nslots, ny, nx, false);
  }
}

// dwt2d.cpp:76-119
void dwt_analyze(const Launch &L, Plane p, int slot0, int nslots, int y, int x, int levels) {
  for (int lv = 0; lv < levels; lv++) {
    int nx = x, ny = y;
    x >>= 1;
    y >>= 1;
    if (y == 0) y = 1;
    if (x == 0) x = 1;
    launch_dwt_level(L, p, slot0, nslots, ny, nx, false);
  }
}

// dwt2d.cpp:128-175
void dwt_synthesize(const Launch &L, Plane p, int slot0, int nslots, int y, int x, int levels) {
  for (int lv = levels - 1; lv >= 0; lv--) {
    int nx = x >> lv, ny = y >> lv;
    if (nx == 0) nx = 1;
    if (ny == 0) ny = 1;
    launch_dwt_level(L, p, slot0, nslots, ny, nx, true);
  }
}
