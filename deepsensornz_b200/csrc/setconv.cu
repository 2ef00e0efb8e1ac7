// SetConv encoder (context sets -> internal grid) and decoder (grid -> off-grid targets).
//
// Replaces, behind the DeepSensor ConvNP API (nzdownscale/downscaler/train.py:238-241,:370), the
// upstream neuralprocesses PrependDensityChannel + SetConv + DivideByFirstChannel encoder and the
// grid->target SetConv decoder.  Gather formulation: every output element sums
// its own contributions from shared-memory staged weights; no global atomics anywhere.
//
// Truncation: a term whose exponent 0.5*d^2/s^2 exceeds 104 is exactly 0.0f in fp32, so restricting
// the sums to |d| <= R = s*sqrt(208) reproduces the dense einsum up to summation order.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int TI = 8;    // output tile rows   (grid dim 1)
constexpr int TJ = 32;   // output tile cols   (grid dim 2)
constexpr int MAXC1 = 9; // max channels incl. density for one context set

// ---- monotone search helpers ----------------------------------------------------------------
// first index p in [0,n) with key(p) >= v where key is ascending (dir=+1) or descending mirrored.
__device__ int lower_bound_f(const float* __restrict__ x, int n, float v, int asc) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    float xv = x[mid];
    bool before = asc ? (xv < v) : (xv > v);
    if (before) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// index range [p0,p1) of coordinates within [a-R, b+R]; full range when not monotone.
__device__ void window_of(const float* __restrict__ x, int n, float a, float b, float R, int mono,
                          int* p0, int* p1) {
  if (mono == 0) { *p0 = 0; *p1 = n; return; }
  float lo = a - R, hi = b + R;
  if (mono > 0) {
    *p0 = lower_bound_f(x, n, lo, 1);
    int q = lower_bound_f(x, n, hi, 1);
    while (q < n && x[q] <= hi) ++q;
    *p1 = q;
  } else {
    // descending: entries > hi come first
    *p0 = lower_bound_f(x, n, hi, 0);
    int q = lower_bound_f(x, n, lo, 0);
    while (q < n && x[q] >= lo) ++q;
    *p1 = q;
  }
  if (*p1 < *p0) *p1 = *p0;
}

// =============================================================================================
// (1a) off-grid context set -> [density ; data/(density+eps)] on the internal grid
// =============================================================================================
constexpr int OG_CHUNK = 128;

__global__ void __launch_bounds__(256)
enc_offgrid_kernel(const float* __restrict__ x, const float* __restrict__ y,
                   const float* __restrict__ mask, int C, int N,
                   double start1, int n1, double start2, int n2, double res,
                   float scale2, float eps, float* __restrict__ out, int ch_off, int c_total) {
  __shared__ float g1s[TI], g2s[TJ];
  __shared__ float w1s[OG_CHUNK][TI];
  __shared__ float w2s[OG_CHUNK][TJ + 1];
  __shared__ float ys[MAXC1][OG_CHUNK];
  __shared__ int warp_cnt[8];
  __shared__ int sel[OG_CHUNK];

  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TJ + tx;
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;
  if (tid < TI) g1s[tid] = cnp_grid_pt(start1, res, min(i0 + tid, n1 - 1));
  if (tid >= 32 && tid < 32 + TJ) g2s[tid - 32] = cnp_grid_pt(start2, res, min(j0 + tid - 32, n2 - 1));
  __syncthreads();
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  const float a1 = g1s[0], b1 = g1s[TI - 1], a2 = g2s[0], b2 = g2s[TJ - 1];

  float acc[MAXC1];
#pragma unroll
  for (int c = 0; c < MAXC1; ++c) acc[c] = 0.f;

  const float* xb = x + (size_t)b * 2 * N;
  const float* yb = y + (size_t)b * C * N;
  const float* mb = mask ? mask + (size_t)b * N : nullptr;

  for (int c0 = 0; c0 < N; c0 += OG_CHUNK) {
    // --- order-preserving compaction of the points that can touch this tile ---
    int n = c0 + tid;
    float p1 = 0.f, p2 = 0.f;
    bool keep = false;
    if (tid < OG_CHUNK && n < N) {
      p1 = xb[n]; p2 = xb[N + n];
      keep = (p1 >= a1 - R) && (p1 <= b1 + R) && (p2 >= a2 - R) && (p2 <= b2 + R);
    }
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (tx == 0) warp_cnt[ty] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < ty) base += warp_cnt[w]; total += warp_cnt[w]; }
    if (keep) sel[base + __popc(bal & ((1u << tx) - 1u))] = n;
    __syncthreads();
    // --- stage weights and (masked) values of the kept points ---
    for (int m = ty; m < total; m += 8) {
      int nn = sel[m];
      float q1 = xb[nn], q2 = xb[N + nn];
      w2s[m][tx] = cnp_rbf(q2, g2s[tx], scale2);
      if (tx < TI) w1s[m][tx] = cnp_rbf(q1, g1s[tx], scale2);
    }
    for (int m = tid; m < total; m += 256) {
      int nn = sel[m];
      float valid = mb ? mb[nn] : 1.f;
      bool nan_any = false;
      for (int c = 0; c < C; ++c) nan_any |= isnan(yb[(size_t)c * N + nn]);
      if (nan_any) valid = 0.f;
      ys[0][m] = valid;
      for (int c = 0; c < C; ++c) {
        float v = yb[(size_t)c * N + nn];
        ys[1 + c][m] = nan_any ? 0.f : v * valid;
      }
    }
    __syncthreads();
    for (int m = 0; m < total; ++m) {
      float w = w1s[m][ty] * w2s[m][tx];
#pragma unroll
      for (int c = 0; c < MAXC1; ++c)
        if (c <= C) acc[c] = fmaf(ys[c][m], w, acc[c]);
    }
    __syncthreads();
  }
  const int i = i0 + ty, j = j0 + tx;
  if (i < n1 && j < n2) {
    float* ob = out + ((size_t)b * c_total + ch_off) * n1 * n2 + (size_t)i * n2 + j;
    const float dens = acc[0];
    ob[0] = dens;
    const float den = dens + eps;
#pragma unroll
    for (int c = 1; c < MAXC1; ++c)
      if (c <= C) ob[(size_t)c * n1 * n2] = acc[c] / den;
  }
}

// =============================================================================================
// (1b) gridded context set -> internal grid, separable two-pass inside one block
// =============================================================================================
constexpr int PCH = 32;  // input rows per chunk
constexpr int QCH = 64;  // input cols per chunk

__global__ void __launch_bounds__(256)
enc_grid_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int x_bstride1, int x_bstride2,
                const float* __restrict__ y, const float* __restrict__ mask, int C, int N1, int N2,
                int mono1, int mono2, double start1, int n1, double start2, int n2, double res,
                float scale2, float eps, float* __restrict__ out, int ch_off, int c_total) {
  __shared__ float g1s[TI], g2s[TJ];
  __shared__ float w1t[PCH][TI];
  __shared__ float w2t[QCH][TJ + 1];
  __shared__ float T[MAXC1][PCH][TJ];
  __shared__ int win[4];

  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TJ + tx;
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;
  const float* x1b = x1 + (size_t)b * x_bstride1;
  const float* x2b = x2 + (size_t)b * x_bstride2;
  if (tid < TI) g1s[tid] = cnp_grid_pt(start1, res, min(i0 + tid, n1 - 1));
  if (tid >= 32 && tid < 32 + TJ) g2s[tid - 32] = cnp_grid_pt(start2, res, min(j0 + tid - 32, n2 - 1));
  __syncthreads();
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  if (tid == 0) window_of(x1b, N1, g1s[0], g1s[TI - 1], R, mono1, &win[0], &win[1]);
  if (tid == 32) window_of(x2b, N2, g2s[0], g2s[TJ - 1], R, mono2, &win[2], &win[3]);
  __syncthreads();
  const int p0 = win[0], p1 = win[1], q0 = win[2], q1 = win[3];
  // per-thread column band (prunes exact-zero terms only)
  int bq0, bq1;
  window_of(x2b, N2, g2s[tx], g2s[tx], R, mono2, &bq0, &bq1);

  const float* yb = y + (size_t)b * C * N1 * N2;
  const float* mb = mask ? mask + (size_t)b * N1 * N2 : nullptr;
  const float my_g1 = g1s[ty];

  float acc[MAXC1];
#pragma unroll
  for (int c = 0; c < MAXC1; ++c) acc[c] = 0.f;

  for (int pc = p0; pc < p1; pc += PCH) {
    const int np = min(PCH, p1 - pc);
    for (int e = tid; e < PCH * TI; e += 256) {
      int pp = e / TI, ii = e % TI;
      w1t[pp][ii] = (pp < np) ? cnp_rbf(x1b[pc + pp], g1s[ii], scale2) : 0.f;
    }
    float Tr[PCH / TI][MAXC1];
#pragma unroll
    for (int m = 0; m < PCH / TI; ++m)
#pragma unroll
      for (int c = 0; c < MAXC1; ++c) Tr[m][c] = 0.f;

    for (int qc = q0; qc < q1; qc += QCH) {
      const int nq = min(QCH, q1 - qc);
      __syncthreads();
      for (int e = tid; e < QCH * TJ; e += 256) {
        int qq = e / TJ, jj = e % TJ;
        w2t[qq][jj] = (qq < nq) ? cnp_rbf(x2b[qc + qq], g2s[jj], scale2) : 0.f;
      }
      __syncthreads();
      const int ks = max(bq0, qc) - qc, ke = min(bq1, qc + nq) - qc;
#pragma unroll
      for (int m = 0; m < PCH / TI; ++m) {
        const int pp = ty + TI * m;
        if (pp < np) {
          const size_t rowoff = (size_t)(pc + pp) * N2 + qc;
          for (int k = ks; k < ke; ++k) {
            const float w = w2t[k][tx];
            float valid = mb ? __ldg(mb + rowoff + k) : 1.f;
            float v[MAXC1 - 1];
            bool nan_any = false;
#pragma unroll
            for (int c = 0; c < MAXC1 - 1; ++c)
              if (c < C) { v[c] = __ldg(yb + (size_t)c * N1 * N2 + rowoff + k); nan_any |= isnan(v[c]); }
            if (nan_any) valid = 0.f;
            const float wv = w * valid;
            Tr[m][0] += wv;
#pragma unroll
            for (int c = 0; c < MAXC1 - 1; ++c)
              if (c < C) Tr[m][1 + c] = fmaf(nan_any ? 0.f : v[c], wv, Tr[m][1 + c]);
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < PCH / TI; ++m)
#pragma unroll
      for (int c = 0; c < MAXC1; ++c)
        if (c <= C) T[c][ty + TI * m][tx] = Tr[m][c];
    __syncthreads();
    for (int pp = 0; pp < np; ++pp) {
      const float w = w1t[pp][ty];
      if (w != 0.f) {
#pragma unroll
        for (int c = 0; c < MAXC1; ++c)
          if (c <= C) acc[c] = fmaf(w, T[c][pp][tx], acc[c]);
      }
    }
    __syncthreads();
  }
  (void)my_g1;
  const int i = i0 + ty, j = j0 + tx;
  if (i < n1 && j < n2) {
    float* ob = out + ((size_t)b * c_total + ch_off) * n1 * n2 + (size_t)i * n2 + j;
    const float dens = acc[0];
    ob[0] = dens;
    const float den = dens + eps;
#pragma unroll
    for (int c = 1; c < MAXC1; ++c)
      if (c <= C) ob[(size_t)c * n1 * n2] = acc[c] / den;
  }
}

// =============================================================================================
// (1c) gridded context set, fast path: precomputed band tables + two separable passes.
//   band table (per output index j of one dimension): first contributing input index q0[j] and the
//   weights w[k][j], k < KB, of inputs q0[j]+k (exact zeros beyond the band).
//   pass 1: T[b,c,p,j] = sum_k y~[b,c,p,q0[j]+k] * w2[k][j]        (y~ = [valid ; y*valid])
//   pass 2: out[b,c,i,j] = sum_k w1[k][i] * T[b,c,p0[i]+k,j], then density normalisation.
// =============================================================================================
constexpr int KB_MAX = 32;

__global__ void __launch_bounds__(128)
band_table_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int xb1, int xb2, int N1, int N2,
                  int mono1, int mono2, double start1, int n1, double start2, int n2, double res, float scale2,
                  int KB, int* __restrict__ tab_i, float* __restrict__ tab_w) {
  // layout per batch-of-coordinates bx: tab_i = [q0_1 (n1) | len_1 (n1) | q0_2 (n2) | len_2 (n2)],
  //                                      tab_w = [w1 (KB x n1) | w2 (KB x n2)]
  const int bx = blockIdx.y;
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= n1 + n2) return;
  const bool dim2 = e >= n1;
  const int j = dim2 ? e - n1 : e, n = dim2 ? n2 : n1, N = dim2 ? N2 : N1;
  const float* x = dim2 ? x2 + (size_t)bx * xb2 : x1 + (size_t)bx * xb1;
  const float g = cnp_grid_pt(dim2 ? start2 : start1, res, j);
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  int lo, hi;
  window_of(x, N, g, g, R, dim2 ? mono2 : mono1, &lo, &hi);
  const int len = min(hi - lo, KB);
  int* ti = tab_i + (size_t)bx * 2 * (n1 + n2) + (dim2 ? 2 * n1 : 0);
  float* tw = tab_w + (size_t)bx * KB * (n1 + n2) + (dim2 ? (size_t)KB * n1 : 0);
  ti[j] = lo; ti[n + j] = len;
  for (int k = 0; k < KB; ++k) tw[(size_t)k * n + j] = (k < len) ? cnp_rbf(x[lo + k], g, scale2) : 0.f;
}

constexpr int P1_ROWS = 4;

__global__ void __launch_bounds__(256)
enc_grid_pass1_kernel(const float* __restrict__ y, const float* __restrict__ mask, int C, int N1, int N2, int n1,
                      int n2, int KB, int tab_bstride_i, int tab_bstride_w, const int* __restrict__ tab_i,
                      const float* __restrict__ tab_w, float* __restrict__ T) {
  extern __shared__ float ys[];  // [C+1][N2]
  const int b = blockIdx.y;
  const int* q0 = tab_i + (size_t)b * tab_bstride_i + 2 * n1;
  const int* qlen = q0 + n2;
  const float* w2 = tab_w + (size_t)b * tab_bstride_w + (size_t)KB * n1;
  const float* yb = y + (size_t)b * C * N1 * N2;
  const float* mb = mask ? mask + (size_t)b * N1 * N2 : nullptr;
  for (int r = 0; r < P1_ROWS; ++r) {
    const int p = blockIdx.x * P1_ROWS + r;
    if (p >= N1) break;
    __syncthreads();
    for (int q = threadIdx.x; q < N2; q += 256) {
      float valid = mb ? mb[(size_t)p * N2 + q] : 1.f;
      bool nan_any = false;
      for (int c = 0; c < C; ++c) nan_any |= isnan(yb[((size_t)c * N1 + p) * N2 + q]);
      if (nan_any) valid = 0.f;
      ys[q] = valid;
      for (int c = 0; c < C; ++c) ys[(c + 1) * N2 + q] = nan_any ? 0.f : yb[((size_t)c * N1 + p) * N2 + q] * valid;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n2; j += 256) {
      const int s = q0[j], len = qlen[j];
      float acc[MAXC1];
#pragma unroll
      for (int c = 0; c < MAXC1; ++c) acc[c] = 0.f;
      for (int k = 0; k < len; ++k) {
        const float w = __ldg(w2 + (size_t)k * n2 + j);
#pragma unroll
        for (int c = 0; c < MAXC1; ++c)
          if (c <= C) acc[c] = fmaf(ys[c * N2 + s + k], w, acc[c]);
      }
#pragma unroll
      for (int c = 0; c < MAXC1; ++c)
        if (c <= C) T[(((size_t)b * (C + 1) + c) * N1 + p) * n2 + j] = acc[c];
    }
  }
}

__global__ void __launch_bounds__(256)
enc_grid_pass2_kernel(const float* __restrict__ T, int C, int N1, int n1, int n2, int KB, int tab_bstride_i,
                      int tab_bstride_w, const int* __restrict__ tab_i, const float* __restrict__ tab_w, float eps,
                      float* __restrict__ out, int ch_off, int c_total) {
  const int tx = threadIdx.x, ty = threadIdx.y, b = blockIdx.z;
  const int i = blockIdx.y * TI + ty, j = blockIdx.x * TJ + tx;
  if (i >= n1 || j >= n2) return;
  const int* p0 = tab_i + (size_t)b * tab_bstride_i;
  const int s = p0[i], len = p0[n1 + i];
  const float* w1 = tab_w + (size_t)b * tab_bstride_w;
  float acc[MAXC1];
#pragma unroll
  for (int c = 0; c < MAXC1; ++c) acc[c] = 0.f;
  for (int k = 0; k < len; ++k) {
    const float w = __ldg(w1 + (size_t)k * n1 + i);
    const float* Tr = T + ((size_t)b * (C + 1) * N1 + s + k) * n2 + j;
#pragma unroll
    for (int c = 0; c < MAXC1; ++c)
      if (c <= C) acc[c] = fmaf(w, __ldg(Tr + (size_t)c * N1 * n2), acc[c]);
  }
  float* ob = out + ((size_t)b * c_total + ch_off) * n1 * n2 + (size_t)i * n2 + j;
  const float dens = acc[0];
  ob[0] = dens;
  const float den = dens + eps;
#pragma unroll
  for (int c = 1; c < MAXC1; ++c)
    if (c <= C) ob[(size_t)c * n1 * n2] = acc[c] / den;
}


// =============================================================================================
// (1d) gridded context set, fused banded path (used when the band tables exist): one block owns FR_I output rows
//   and up to FR_J output columns (one per thread) and streams the input rows of its band through shared memory:
//     row p:  Trow[c]   = sum_k y~[c, p, q0[j]+k] w2[k][j]            (horizontal band, <= KB taps)
//             acc[i][c] += w1[p - p0[i]][i] * Trow[c]                  (the <= FR_I output rows that row p touches)
//   Same summation order as the two-pass path (bit-identical results) but no intermediate T tensor: the input is
//   read ~1.3x (row halo between neighbouring blocks) and the output written once.  Rows are staged FR_RS at a time,
//   double-buffered, so the next rows' loads are in flight while the current ones are consumed.
// =============================================================================================
constexpr int FR_I = 8;       // output rows per block
constexpr int FR_J = 320;     // output columns per block = threads
constexpr int FR_MAXROWS = 160;  // max input rows in the band of FR_I output rows

// LPT = loads per thread per staged row segment (seg_max <= LPT * FR_J), FR_RS = input rows staged per step,
// KBT = compile-time bound of the horizontal band (taps kept in registers)
template <int C, int LPT, int FR_RS, int KBT>
__global__ void __launch_bounds__(FR_J)
enc_grid_fused_kernel(const float* __restrict__ y, const float* __restrict__ mask, int N1, int N2, int n1, int n2, int KB,
                      int tab_bstride_i, int tab_bstride_w, const int* __restrict__ tab_i,
                      const float* __restrict__ tab_w, float eps, float* __restrict__ out, int ch_off, int c_total,
                      int seg_max) {
  extern __shared__ float fr_smem[];
  float* ys = fr_smem;                                     // [2][FR_RS][C+1][seg_max]
  float* w1blk = fr_smem + 2 * FR_RS * (C + 1) * seg_max;  // [FR_MAXROWS][FR_I]
  __shared__ int sh[4];
  const int tid = threadIdx.x, b = blockIdx.z;
  const int i0 = blockIdx.y * FR_I, j0 = blockIdx.x * FR_J;
  const int ni = min(FR_I, n1 - i0), nj = min(FR_J, n2 - j0);
  const int* t1 = tab_i + (size_t)b * tab_bstride_i;       // p0 (n1) | len1 (n1) | q0 (n2) | len2 (n2)
  const int* t2 = t1 + 2 * n1;
  const float* w1 = tab_w + (size_t)b * tab_bstride_w;
  const float* w2 = w1 + (size_t)KB * n1;
  if (tid == 0) { sh[0] = N1; sh[1] = 0; sh[2] = N2; sh[3] = 0; }
  __syncthreads();
  if (tid < ni) {
    const int s = t1[i0 + tid], l = t1[n1 + i0 + tid];
    if (l > 0) { atomicMin(&sh[0], s); atomicMax(&sh[1], s + l); }
  }
  if (tid < nj) {
    const int s = t2[j0 + tid], l = t2[n2 + j0 + tid];
    if (l > 0) { atomicMin(&sh[2], s); atomicMax(&sh[3], s + l); }
  }
  __syncthreads();
  if (tid == 0) { sh[1] = max(sh[1], sh[0]); sh[3] = max(sh[3], sh[2]); }
  __syncthreads();
  const int plo = sh[0], phi = min(sh[1], sh[0] + FR_MAXROWS), qlo = sh[2], nseg = min(sh[3] - sh[2], seg_max);
  // vertical weights of this block: w1blk[p - plo][i]
  for (int e = tid; e < (phi - plo) * FR_I; e += FR_J) {
    const int pp = e / FR_I, i = e - pp * FR_I;
    float w = 0.f;
    if (i < ni) {
      const int s = t1[i0 + i], l = t1[n1 + i0 + i], k = plo + pp - s;
      if (k >= 0 && k < l) w = w1[(size_t)k * n1 + i0 + i];
    }
    w1blk[e] = w;
  }
  const bool active = tid < nj;
  const int j = j0 + tid;
  const int myq = active ? t2[j] - qlo : 0, mylen = active ? min(t2[n2 + j], KBT) : 0;
  float w2r[KBT];
#pragma unroll
  for (int k = 0; k < KBT; ++k) w2r[k] = (k < mylen) ? __ldg(w2 + (size_t)k * n2 + j) : 0.f;
  float acc[FR_I][C + 1];
#pragma unroll
  for (int i = 0; i < FR_I; ++i)
#pragma unroll
    for (int c = 0; c <= C; ++c) acc[i][c] = 0.f;
  const float* yb = y + (size_t)b * C * N1 * N2;
  const float* mb = mask ? mask + (size_t)b * N1 * N2 : nullptr;
  float pre[FR_RS][C + 1][LPT];
  auto prefetch = [&](int p) {
#pragma unroll
    for (int r = 0; r < FR_RS; ++r)
#pragma unroll
      for (int l = 0; l < LPT; ++l) {
        const int q = tid + l * FR_J;
        const bool ok = (p + r < phi) && (q < nseg);
        const size_t off = (size_t)(p + r) * N2 + qlo + q;
        pre[r][0][l] = ok ? (mb ? __ldg(mb + off) : 1.f) : 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) pre[r][1 + c][l] = ok ? __ldg(yb + (size_t)c * N1 * N2 + off) : 0.f;
      }
  };
  prefetch(plo);
  int buf = 0;
  for (int p = plo; p < phi; p += FR_RS, buf ^= 1) {
    float* yd = ys + (size_t)buf * FR_RS * (C + 1) * seg_max;
#pragma unroll
    for (int r = 0; r < FR_RS; ++r)
#pragma unroll
      for (int l = 0; l < LPT; ++l) {
        const int q = tid + l * FR_J;
        if (q < nseg) {
          float valid = pre[r][0][l];
          bool nan_any = false;
#pragma unroll
          for (int c = 0; c < C; ++c) nan_any |= isnan(pre[r][1 + c][l]);
          if (nan_any) valid = 0.f;
          yd[(r * (C + 1)) * seg_max + q] = valid;
#pragma unroll
          for (int c = 0; c < C; ++c) yd[(r * (C + 1) + 1 + c) * seg_max + q] = nan_any ? 0.f : pre[r][1 + c][l] * valid;
        }
      }
    __syncthreads();
    if (p + FR_RS < phi) prefetch(p + FR_RS);
#pragma unroll
    for (int r = 0; r < FR_RS; ++r) {
      if (p + r >= phi) break;
      float Tr[C + 1];
#pragma unroll
      for (int c = 0; c <= C; ++c) Tr[c] = 0.f;
      const float* yr = yd + (r * (C + 1)) * seg_max + myq;
#pragma unroll
      for (int k = 0; k < KBT; ++k) {
        if (k < mylen) {
#pragma unroll
          for (int c = 0; c <= C; ++c) Tr[c] = fmaf(yr[c * seg_max + k], w2r[k], Tr[c]);
        }
      }
      const float* wv = w1blk + (p + r - plo) * FR_I;
#pragma unroll
      for (int i = 0; i < FR_I; ++i) {
        const float w = wv[i];
        if (w != 0.f) {
#pragma unroll
          for (int c = 0; c <= C; ++c) acc[i][c] = fmaf(w, Tr[c], acc[i][c]);
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int i = 0; i < FR_I; ++i) {
    if (i < ni) {
      float* ob = out + ((size_t)b * c_total + ch_off) * n1 * n2 + (size_t)(i0 + i) * n2 + j;
      const float dens = acc[i][0];
      ob[0] = dens;
      const float den = dens + eps;
#pragma unroll
      for (int c = 1; c <= C; ++c) ob[(size_t)c * n1 * n2] = acc[i][c] / den;
    }
  }
}

template <int C, int LPT, int FR_RS, int KBT>
int launch_enc_fused(const float* y, const float* mask, int B, int N1, int N2, int n1, int n2, int KB, int tbi, int tbw,
                     const int* tab_i, const float* tab_w, float eps, float* out, int ch_off, int c_total, int seg_max,
                     cudaStream_t stream) {
  const size_t smem = ((size_t)2 * FR_RS * (C + 1) * seg_max + (size_t)FR_MAXROWS * FR_I) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr && smem > 48 * 1024) {
    cudaFuncSetAttribute((enc_grid_fused_kernel<C, LPT, FR_RS, KBT>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  dim3 grid(cnp_cdiv(n2, FR_J), cnp_cdiv(n1, FR_I), B);
  enc_grid_fused_kernel<C, LPT, FR_RS, KBT><<<grid, FR_J, smem, stream>>>(y, mask, N1, N2, n1, n2, KB, tbi, tbw, tab_i, tab_w, eps, out, ch_off,
                                                         c_total, seg_max);
  return 0;
}

// =============================================================================================
// (1e) gridded context set, output-tile path for inputs no finer than the internal grid (ERA5 / aux fields):
//   one block = ET_I x ET_J output pixels.  The input window of the tile (band union, ~(ET*dx_in/dx_out + band)^2
//   cells) is staged in shared memory once, the horizontal band pass writes T[c][row][j] to shared memory and the
//   vertical pass finishes the tile -- no global intermediate, one launch, same summation order as the two-pass
//   path (bit-identical results).
// =============================================================================================
constexpr int ET_I = 16, ET_J = 32;

template <int KBT>
__global__ void __launch_bounds__(256)
enc_grid_tile_kernel(const float* __restrict__ y, const float* __restrict__ mask, int C, int N1, int N2, int n1, int n2,
                     int KB, int tab_bstride_i, int tab_bstride_w, const int* __restrict__ tab_i,
                     const float* __restrict__ tab_w, float eps, float* __restrict__ out, int ch_off, int c_total,
                     int max_rows, int max_cols) {
  extern __shared__ float et_smem[];
  float* yw = et_smem;                                        // [C+1][max_rows][max_cols]
  float* T = et_smem + (size_t)(C + 1) * max_rows * max_cols; // [C+1][max_rows][ET_J]
  __shared__ int sh[4];
  const int tid = threadIdx.x, b = blockIdx.z;
  const int i0 = blockIdx.y * ET_I, j0 = blockIdx.x * ET_J;
  const int ni = min(ET_I, n1 - i0), nj = min(ET_J, n2 - j0);
  const int* t1 = tab_i + (size_t)b * tab_bstride_i;
  const int* t2 = t1 + 2 * n1;
  const float* w1 = tab_w + (size_t)b * tab_bstride_w;
  const float* w2 = w1 + (size_t)KB * n1;
  // band union of the tile, in parallel (a serial scan by one thread costs ~1 us of L2 latency per table entry)
  if (tid == 0) { sh[0] = N1; sh[1] = 0; sh[2] = N2; sh[3] = 0; }
  __syncthreads();
  if (tid < ni) {
    const int s = t1[i0 + tid], l = t1[n1 + i0 + tid];
    if (l > 0) { atomicMin(&sh[0], s); atomicMax(&sh[1], s + l); }
  } else if (tid >= 64 && tid < 64 + nj) {
    const int s = t2[j0 + tid - 64], l = t2[n2 + j0 + tid - 64];
    if (l > 0) { atomicMin(&sh[2], s); atomicMax(&sh[3], s + l); }
  }
  __syncthreads();
  const int plo = sh[0], nr = max(min(sh[1] - sh[0], max_rows), 0), qlo = sh[2], nc = max(min(sh[3] - sh[2], max_cols), 0);
  const float* yb = y + (size_t)b * C * N1 * N2;
  const float* mb = mask ? mask + (size_t)b * N1 * N2 : nullptr;
  // 1. stage y~ = [valid ; y * valid] of the window
  for (int e = tid; e < nr * nc; e += 256) {
    const int r = e / nc, q = e - r * nc;
    const size_t off = (size_t)(plo + r) * N2 + qlo + q;
    float valid = mb ? __ldg(mb + off) : 1.f;
    bool nan_any = false;
    for (int c = 0; c < C; ++c) nan_any |= isnan(__ldg(yb + (size_t)c * N1 * N2 + off));
    if (nan_any) valid = 0.f;
    yw[(size_t)r * max_cols + q] = valid;
    for (int c = 0; c < C; ++c)
      yw[((size_t)(1 + c) * max_rows + r) * max_cols + q] = nan_any ? 0.f : __ldg(yb + (size_t)c * N1 * N2 + off) * valid;
  }
  __syncthreads();
  // 2. horizontal band pass: thread = (tile column j, rows r = tid/32, +8, ...)
  const int tj = tid & 31, tr = tid >> 5;
  const bool colok = tj < nj;
  const int jg = j0 + tj;
  const int myq = colok ? t2[jg] - qlo : 0, mylen = colok ? min(t2[n2 + jg], KBT) : 0;
  float w2r[KBT];
#pragma unroll
  for (int k = 0; k < KBT; ++k) w2r[k] = (k < mylen) ? __ldg(w2 + (size_t)k * n2 + jg) : 0.f;
  for (int c = 0; c <= C; ++c)
    for (int r = tr; r < nr; r += 8) {
      const float* src = yw + ((size_t)c * max_rows + r) * max_cols + myq;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < KBT; ++k)
        if (k < mylen) acc = fmaf(src[k], w2r[k], acc);
      T[((size_t)c * max_rows + r) * ET_J + tj] = acc;
    }
  __syncthreads();
  // 3. vertical band pass + density normalisation: thread = (column j, output rows i = tid/32 and +8)
#pragma unroll
  for (int m = 0; m < ET_I / 8; ++m) {
    const int i = tr + 8 * m;
    if (i < ni && colok) {
      const int ig = i0 + i;
      const int s = t1[ig] - plo, len = min(t1[n1 + ig], KBT);
      float w1r[KBT];
#pragma unroll
      for (int k = 0; k < KBT; ++k) w1r[k] = (k < len) ? __ldg(w1 + (size_t)k * n1 + ig) : 0.f;
      float* ob = out + ((size_t)b * c_total + ch_off) * n1 * n2 + (size_t)ig * n2 + jg;
      float dens = 0.f;
      for (int c = 0; c <= C; ++c) {
        const float* src = T + ((size_t)c * max_rows + s) * ET_J + tj;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < KBT; ++k)
          if (k < len) acc = fmaf(w1r[k], src[(size_t)k * ET_J], acc);
        if (c == 0) { dens = acc; ob[0] = acc; }
        else ob[(size_t)c * n1 * n2] = acc / (dens + eps);
      }
    }
  }
}

// =============================================================================================
// (3) decoder: grid -> off-grid targets, forward.  One block per (target, batch).
//     z is NCHW fp32 [B, C, n1, n2] with batch stride z_bstride (elements).
// =============================================================================================
__global__ void __launch_bounds__(256)
dec_offgrid_fwd_kernel(const float* __restrict__ z, long long z_bstride, const float* __restrict__ xt,
                       int C, int Nt, double start1, int n1, double start2, int n2, double res,
                       float scale2, float* __restrict__ f, int f_ctotal) {
  __shared__ float w1s[64], w2s[64];
  __shared__ int rng[4];
  const int t = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float p1 = xt[((size_t)b * 2 + 0) * Nt + t], p2 = xt[((size_t)b * 2 + 1) * Nt + t];
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  if (threadIdx.x == 0) {
    // conservative index window (one extra cell each side), clipped to the grid
    int ilo = (int)floor(((double)p1 - R - start1) / res) - 1, ihi = (int)ceil(((double)p1 + R - start1) / res) + 1;
    int jlo = (int)floor(((double)p2 - R - start2) / res) - 1, jhi = (int)ceil(((double)p2 + R - start2) / res) + 1;
    rng[0] = max(ilo, 0); rng[1] = min(ihi + 1, n1); rng[2] = max(jlo, 0); rng[3] = min(jhi + 1, n2);
  }
  __syncthreads();
  const int ilo = rng[0], ihi = rng[1], jlo = rng[2], jhi = rng[3];
  const float* zb = z + (size_t)b * z_bstride;
  for (int c = warp; c < C; c += 8) {
    float acc = 0.f;
    const float* zc = zb + (size_t)c * n1 * n2;
    for (int ic = ilo; ic < ihi; ic += 64) {
      for (int jc = jlo; jc < jhi; jc += 64) {
        __syncthreads();
        if (threadIdx.x < 64) {
          int i = ic + threadIdx.x;
          w1s[threadIdx.x] = (i < ihi) ? cnp_rbf(p1, cnp_grid_pt(start1, res, i), scale2) : 0.f;
        } else if (threadIdx.x < 128) {
          int j = jc + threadIdx.x - 64;
          w2s[threadIdx.x - 64] = (j < jhi) ? cnp_rbf(p2, cnp_grid_pt(start2, res, j), scale2) : 0.f;
        }
        __syncthreads();
        const int ni = min(64, ihi - ic), nj = min(64, jhi - jc);
        for (int ii = 0; ii < ni; ++ii) {
          const float wi = w1s[ii];
          const float* zr = zc + (size_t)(ic + ii) * n2 + jc;
          float part = 0.f;
          for (int jj = lane; jj < nj; jj += 32) part = fmaf(__ldg(zr + jj), w2s[jj], part);
          acc = fmaf(wi, part, acc);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) f[((size_t)b * f_ctotal + c) * Nt + t] = acc;
  }
}

// =============================================================================================
// (3) decoder backward: dz[b,c,i,j] = sum_t df[b,c,t] w1[i,t] w2[j,t]   (dense write of dz)
//     Tile 8x32 pixels; targets touching the tile are compacted into shared memory.
// =============================================================================================
constexpr int DB_MAXT = 64;   // targets staged per pass
constexpr int DB_CCH = 16;    // channels per register pass

__global__ void __launch_bounds__(256)
dec_offgrid_bwd_kernel(const float* __restrict__ df, int f_ctotal, const float* __restrict__ xt,
                       int C, int Nt, double start1, int n1, double start2, int n2, double res,
                       float scale2, float* __restrict__ dz, long long dz_bstride) {
  __shared__ float g1s[TI], g2s[TJ];
  __shared__ float w1s[DB_MAXT][TI];
  __shared__ float w2s[DB_MAXT][TJ + 1];
  __shared__ float dfs[DB_MAXT][DB_CCH + 1];
  __shared__ int sel[DB_MAXT];
  __shared__ int nsel_s, tnext_s;

  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TJ + tx;
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;
  if (tid < TI) g1s[tid] = cnp_grid_pt(start1, res, min(i0 + tid, n1 - 1));
  if (tid >= 32 && tid < 32 + TJ) g2s[tid - 32] = cnp_grid_pt(start2, res, min(j0 + tid - 32, n2 - 1));
  __syncthreads();
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  const float a1 = g1s[0], b1 = g1s[TI - 1], a2 = g2s[0], b2 = g2s[TJ - 1];
  const float* xb = xt + (size_t)b * 2 * Nt;
  const int i = i0 + ty, j = j0 + tx;
  const bool inb = (i < n1 && j < n2);
  float* dzb = dz + (size_t)b * dz_bstride + (size_t)i * n2 + j;

  for (int cc = 0; cc < C; cc += DB_CCH) {
    float acc[DB_CCH];
#pragma unroll
    for (int c = 0; c < DB_CCH; ++c) acc[c] = 0.f;
    int tnext = 0;
    while (tnext < Nt) {
      // serial, order-preserving selection of up to DB_MAXT touching targets (Nt is small)
      __syncthreads();
      if (tid == 0) {
        int cnt = 0, t = tnext;
        for (; t < Nt && cnt < DB_MAXT; ++t) {
          float p1 = xb[t], p2 = xb[Nt + t];
          if (p1 >= a1 - R && p1 <= b1 + R && p2 >= a2 - R && p2 <= b2 + R) sel[cnt++] = t;
        }
        nsel_s = cnt;
        tnext_s = t;
      }
      __syncthreads();
      const int total = nsel_s;
      tnext = tnext_s;
      for (int m = ty; m < total; m += 8) {
        int tt = sel[m];
        w2s[m][tx] = cnp_rbf(xb[Nt + tt], g2s[tx], scale2);
        if (tx < TI) w1s[m][tx] = cnp_rbf(xb[tt], g1s[tx], scale2);
      }
      for (int e = tid; e < total * DB_CCH; e += 256) {
        int m = e / DB_CCH, c = e % DB_CCH;
        dfs[m][c] = (cc + c < C) ? df[((size_t)b * f_ctotal + cc + c) * Nt + sel[m]] : 0.f;
      }
      __syncthreads();
      for (int m = 0; m < total; ++m) {
        const float w = w1s[m][ty] * w2s[m][tx];
#pragma unroll
        for (int c = 0; c < DB_CCH; ++c) acc[c] = fmaf(dfs[m][c], w, acc[c]);
      }
    }
    if (inb) {
#pragma unroll
      for (int c = 0; c < DB_CCH; ++c)
        if (cc + c < C) dzb[(size_t)(cc + c) * n1 * n2] = acc[c];
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
CNP_API int cnp_setconv_enc_offgrid_fwd(const float* x, const float* y, const float* mask, int B, int C, int N,
                                        double start1, int n1, double start2, int n2, double res,
                                        float scale2, float eps, float* out, int ch_off, int c_total,
                                        cudaStream_t stream) {
  CNP_REQUIRE(B > 0 && C >= 1 && C + 1 <= MAXC1, "enc_offgrid: need 1 <= C <= %d (got %d)", MAXC1 - 1, C);
  CNP_REQUIRE(n1 > 0 && n2 > 0 && N >= 0, "enc_offgrid: bad sizes");
  CNP_REQUIRE(ch_off >= 0 && ch_off + C + 1 <= c_total, "enc_offgrid: channel window out of range");
  dim3 grid(cnp_cdiv(n2, TJ), cnp_cdiv(n1, TI), B), block(TJ, TI);
  enc_offgrid_kernel<<<grid, block, 0, stream>>>(x, y, mask, C, N, start1, n1, start2, n2, res, scale2, eps,
                                                 out, ch_off, c_total);
  CNP_LAUNCH_CHECK("enc_offgrid_kernel");
  return 0;
}

// Workspace (bytes) of the banded fast path of cnp_setconv_enc_grid_fwd.
CNP_API long long cnp_setconv_enc_grid_workspace_bytes(int B, int C, int N1, int n1, int n2, int band) {
  if (band < 1 || band > KB_MAX) return 0;
  long long tabs = (long long)B * (2LL * (n1 + n2) * sizeof(int) + (long long)band * (n1 + n2) * sizeof(float));
  long long T = (long long)B * (C + 1) * N1 * n2 * sizeof(float);
  return ((tabs + 255) / 256) * 256 + T;
}

// band: upper bound on the number of inputs within the truncation radius of any grid point along
// either dimension (host-computed from the coordinates); 1..32 with monotone coordinates and a
// workspace selects the two-pass banded path, anything else the generic single-kernel path.
CNP_API int cnp_setconv_enc_grid_fwd(const float* x1, const float* x2, int x_batched, const float* y,
                                     const float* mask, int B, int C, int N1, int N2, int mono1, int mono2,
                                     double start1, int n1, double start2, int n2, double res, float scale2,
                                     float eps, float* out, int ch_off, int c_total, int band, void* workspace,
                                     long long workspace_bytes, cudaStream_t stream) {
  CNP_REQUIRE(B > 0 && C >= 1 && C + 1 <= MAXC1, "enc_grid: need 1 <= C <= %d (got %d)", MAXC1 - 1, C);
  CNP_REQUIRE(N1 > 0 && N2 > 0 && n1 > 0 && n2 > 0, "enc_grid: bad sizes");
  CNP_REQUIRE(ch_off >= 0 && ch_off + C + 1 <= c_total, "enc_grid: channel window out of range");
  const size_t p1_smem = (size_t)(C + 1) * N2 * sizeof(float);
  if (band >= 1 && band <= KB_MAX && mono1 != 0 && mono2 != 0 && workspace && p1_smem <= 200 * 1024) {
    const long long need = cnp_setconv_enc_grid_workspace_bytes(B, C, N1, n1, n2, band);
    CNP_REQUIRE(workspace_bytes >= need, "enc_grid: workspace too small (%lld < %lld)", workspace_bytes, need);
    const int Bx = x_batched ? B : 1;
    int* tab_i = reinterpret_cast<int*>(workspace);
    float* tab_w = reinterpret_cast<float*>(tab_i + (size_t)B * 2 * (n1 + n2));
    const long long tabs = (long long)B * (2LL * (n1 + n2) * sizeof(int) + (long long)band * (n1 + n2) * sizeof(float));
    float* T = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((tabs + 255) / 256) * 256);
    dim3 gb(cnp_cdiv(n1 + n2, 128), Bx);
    band_table_kernel<<<gb, 128, 0, stream>>>(x1, x2, x_batched ? N1 : 0, x_batched ? N2 : 0, N1, N2, mono1, mono2,
                                              start1, n1, start2, n2, res, scale2, band, tab_i, tab_w);
    CNP_LAUNCH_CHECK("band_table_kernel");
    const int tbi = x_batched ? 2 * (n1 + n2) : 0, tbw = x_batched ? band * (n1 + n2) : 0;
    // inputs no finer than ~2x the grid: output-tile kernel (window + T in shared memory, one launch)
    if (N1 < 2 * n1 && N2 < 2 * n2) {
      const int max_rows = (int)((double)N1 / n1 * ET_I) + band + 4, max_cols = (int)((double)N2 / n2 * ET_J) + band + 4;
      const size_t tsmem = ((size_t)(C + 1) * max_rows * max_cols + (size_t)(C + 1) * max_rows * ET_J) * sizeof(float);
      if (tsmem <= 96 * 1024) {
        dim3 gt(cnp_cdiv(n2, ET_J), cnp_cdiv(n1, ET_I), B);
        static size_t attr16 = 0, attr32 = 0;
        if (band <= 16) {
          if (tsmem > attr16 && tsmem > 48 * 1024) { cudaFuncSetAttribute(enc_grid_tile_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem); attr16 = tsmem; }
          enc_grid_tile_kernel<16><<<gt, 256, tsmem, stream>>>(y, mask, C, N1, N2, n1, n2, band, tbi, tbw, tab_i, tab_w, eps, out,
                                                              ch_off, c_total, max_rows, max_cols);
        } else {
          if (tsmem > attr32 && tsmem > 48 * 1024) { cudaFuncSetAttribute(enc_grid_tile_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem); attr32 = tsmem; }
          enc_grid_tile_kernel<32><<<gt, 256, tsmem, stream>>>(y, mask, C, N1, N2, n1, n2, band, tbi, tbw, tab_i, tab_w, eps, out,
                                                              ch_off, c_total, max_rows, max_cols);
        }
        CNP_LAUNCH_CHECK("enc_grid_tile_kernel");
        return 0;
      }
    }
    // fused single pass when the staged row segment and the row band fit (input no finer than ~20x the grid)
    {
      const int seg_max = ((N2 < 6 * FR_J ? N2 : 6 * FR_J) + 3) & ~3;
      // one column block sees the whole input row; only worth it when the input is finer than the grid (the
      // streamed rows are then the dominant traffic); coarse inputs keep the two-pass path
      const bool seg_ok = (n2 <= FR_J) && (N2 <= 6 * FR_J) && (N1 >= 2 * n1);
      const bool rows_ok = (double)N1 / n1 * FR_I + band + 10 <= FR_MAXROWS;
      const int lpt = cnp_cdiv(seg_max, FR_J);
      // (LPT, rows per step): narrow inputs stage 8 rows per step, wide ones 2; register budget limits C
      const int rs = lpt == 1 ? 8 : 2;
      const size_t fsmem = ((size_t)2 * rs * (C + 1) * seg_max + (size_t)FR_MAXROWS * FR_I) * sizeof(float);
      if (seg_ok && rows_ok && fsmem <= 200 * 1024 && C <= 8) {
        int rc = -1;
#define CNP_ENC_FUSED(CC, LL, RR)                                                                                        \
  rc = band <= 16 ? launch_enc_fused<CC, LL, RR, 16>(y, mask, B, N1, N2, n1, n2, band, tbi, tbw, tab_i, tab_w, eps, out,    \
                                                     ch_off, c_total, seg_max, stream)                                    \
                  : launch_enc_fused<CC, LL, RR, 32>(y, mask, B, N1, N2, n1, n2, band, tbi, tbw, tab_i, tab_w, eps, out,    \
                                                     ch_off, c_total, seg_max, stream)
        if (lpt == 1) {
          switch (C) {
            case 1: CNP_ENC_FUSED(1, 1, 8); break; case 2: CNP_ENC_FUSED(2, 1, 8); break; case 3: CNP_ENC_FUSED(3, 1, 8); break;
            case 4: CNP_ENC_FUSED(4, 1, 8); break; case 5: CNP_ENC_FUSED(5, 1, 8); break; case 6: CNP_ENC_FUSED(6, 1, 8); break;
            case 7: CNP_ENC_FUSED(7, 1, 8); break; case 8: CNP_ENC_FUSED(8, 1, 8); break;
          }
        } else if (lpt <= 3 && C <= 4) {
          switch (C) {
            case 1: CNP_ENC_FUSED(1, 3, 2); break; case 2: CNP_ENC_FUSED(2, 3, 2); break; case 3: CNP_ENC_FUSED(3, 3, 2); break;
            case 4: CNP_ENC_FUSED(4, 3, 2); break;
          }
        } else if (lpt <= 6 && C <= 2) {
          switch (C) { case 1: CNP_ENC_FUSED(1, 6, 2); break; case 2: CNP_ENC_FUSED(2, 6, 2); break; }
        }
#undef CNP_ENC_FUSED
        if (rc == 0) { CNP_LAUNCH_CHECK("enc_grid_fused_kernel"); return 0; }
      }
    }
    static size_t attr = 0;
    if (p1_smem > attr && p1_smem > 48 * 1024) {
      cudaFuncSetAttribute(enc_grid_pass1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p1_smem);
      attr = p1_smem;
    }
    dim3 g1(cnp_cdiv(N1, P1_ROWS), B);
    enc_grid_pass1_kernel<<<g1, 256, p1_smem, stream>>>(y, mask, C, N1, N2, n1, n2, band, tbi, tbw, tab_i, tab_w, T);
    CNP_LAUNCH_CHECK("enc_grid_pass1_kernel");
    dim3 g2(cnp_cdiv(n2, TJ), cnp_cdiv(n1, TI), B), block(TJ, TI);
    enc_grid_pass2_kernel<<<g2, block, 0, stream>>>(T, C, N1, n1, n2, band, tbi, tbw, tab_i, tab_w, eps, out, ch_off,
                                                    c_total);
    CNP_LAUNCH_CHECK("enc_grid_pass2_kernel");
    return 0;
  }
  dim3 grid(cnp_cdiv(n2, TJ), cnp_cdiv(n1, TI), B), block(TJ, TI);
  enc_grid_kernel<<<grid, block, 0, stream>>>(x1, x2, x_batched ? N1 : 0, x_batched ? N2 : 0, y, mask, C, N1, N2,
                                              mono1, mono2, start1, n1, start2, n2, res, scale2, eps, out, ch_off,
                                              c_total);
  CNP_LAUNCH_CHECK("enc_grid_kernel");
  return 0;
}

CNP_API int cnp_setconv_dec_offgrid_fwd(const float* z, long long z_bstride, const float* xt, int B, int C, int Nt,
                                        double start1, int n1, double start2, int n2, double res, float scale2,
                                        float* f, int f_ctotal, cudaStream_t stream) {
  CNP_REQUIRE(B > 0 && C > 0 && Nt >= 0 && f_ctotal >= C, "dec_offgrid_fwd: bad sizes");
  if (Nt == 0) return 0;
  dim3 grid(Nt, B);
  dec_offgrid_fwd_kernel<<<grid, 256, 0, stream>>>(z, z_bstride, xt, C, Nt, start1, n1, start2, n2, res, scale2,
                                                   f, f_ctotal);
  CNP_LAUNCH_CHECK("dec_offgrid_fwd_kernel");
  return 0;
}

CNP_API int cnp_setconv_dec_offgrid_bwd(const float* df, int f_ctotal, const float* xt, int B, int C, int Nt,
                                        double start1, int n1, double start2, int n2, double res, float scale2,
                                        float* dz, long long dz_bstride, cudaStream_t stream) {
  CNP_REQUIRE(B > 0 && C > 0 && Nt >= 0 && f_ctotal >= C, "dec_offgrid_bwd: bad sizes");
  dim3 grid(cnp_cdiv(n2, TJ), cnp_cdiv(n1, TI), B), block(TJ, TI);
  dec_offgrid_bwd_kernel<<<grid, block, 0, stream>>>(df, f_ctotal, xt, C, Nt, start1, n1, start2, n2, res, scale2,
                                                     dz, dz_bstride);
  CNP_LAUNCH_CHECK("dec_offgrid_bwd_kernel");
  return 0;
}
