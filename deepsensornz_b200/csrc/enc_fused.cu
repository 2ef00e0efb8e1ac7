// Fused SetConv encoder: every context set of a task -> the UNet's input tensor, one launch.
//
// Replaces, behind ConvNP.loss_fn / predict (nzdownscale/downscaler/train.py:370, validate_ERA.py:88-92), upstream
// neuralprocesses' PrependDensityChannel + SetConv (per context set) + DivideByFirstChannel + Concatenate
// (SURVEY.md A.3): per set k
//     h_k[c,i,j] = sum_n y~[c,n] exp(-(x1_n-g1_i)^2 / 2 s_k^2) exp(-(x2_n-g2_j)^2 / 2 s_k^2),   y~ = [valid ; y * valid]
//     out = [h_k[0] ; h_k[1:] / (h_k[0] + eps)]
// with NaN observations turned into valid = 0 on the fly (ConvNP.modify_task's host scan) and terms whose exponent
// exceeds 104 (exactly 0.0f in fp32) left out of the sums.
//
// One CTA = one 8 x 32 tile of the internal grid of one task; it walks the task's context sets:
//   * gridded set: the input rows inside the tile's band are streamed through shared memory 16 / 32 at a time; a
//     horizontal band pass (<= 32 taps per output column) writes T[c][row][j] to shared memory and the vertical pass
//     accumulates the thread's pixel.  Band starts and weights come from per-set tables (cnp_encode_tables) that depend
//     only on (coordinates, internal grid, length scale): the caller builds them once and keeps them across steps, so a
//     CTA starts with two coalesced table reads instead of binary searches and expf.  Gather form, no atomics.
//   * off-grid set: points that can touch the tile are compacted in order, their separable weights staged in shared
//     memory, every thread sums its own pixel.
//   * precomputed planes: channels of sets that the whole batch shares (topography aux, land mask) are encoded ONCE
//     per step by a first launch of this same kernel (B = 1, fp32 planes) and copied in here from L2.
// The tile's channels meet in shared memory and leave as either fp32 NCHW (parity mode) or -- bf16 UNet -- directly
// as the blocked bf16 tensor [B][chunk][H+4][W+4][8] with the constant-1 channel of the folded first layer
// (fold_in.cu): no fp32 encoder tensor, no layout-conversion kernel.  Bound: HBM.
#include "tc_common.cuh"
#include <math.h>

struct cnp_enc_set {
  int kind;          // 0 off-grid, 1 gridded, 2 precomputed fp32 planes [C][n1][n2] (already normalised)
  int C;             // data channels (kind 2: number of planes)
  int ch_off;        // first output channel (the density channel for kinds 0 / 1)
  int batched;       // y (and mask) carry a batch axis; 0: one field shared by every task
  const float* x1;   // gridded: [N1]; off-grid: x [B,2,N]
  const float* x2;   // gridded: [N2]
  const float* y;    // gridded [B or 1, C, N1, N2]; off-grid [B, C, N]
  const float* mask; // gridded [B or 1, 1, N1, N2] | off-grid [B, 1, N] | NULL
  int N1, N2;        // off-grid: N1 = N
  int mono1, mono2;  // +1 ascending, -1 descending (gridded coordinates)
  float scale2;      // s_k^2
  int KB;            // gridded: band width of the tables (<= 32)
  const int* tab_i;  // gridded: [p0 (n1) | len1 (n1) | q0 (n2) | len2 (n2)]   (cnp_encode_tables)
  const float* tab_w;// gridded: [w1 (KB x n1) | w2 (KB x n2)]
};
struct cnp_enc_sets {
  int n_sets;
  int pad_;
  cnp_enc_set s[8];
};

namespace {

constexpr int TI = 8, TJ = 32, NT = 256;
constexpr int KBMAX = 32;    // max band (inputs within R of one grid row / column)
constexpr int OGC = 256;     // off-grid points per chunk (one per thread)
constexpr int OGS = 64;      // off-grid points whose weights are staged at a time
constexpr int MAXC1 = 9;     // max channels incl. density of one set

__device__ int lower_bound_f(const float* __restrict__ x, int n, float v, int asc) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const float xv = __ldg(x + mid);
    const bool before = asc ? (xv < v) : (xv > v);
    if (before) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// index range [p0,p1) of the monotone coordinates within [a-R, b+R]
__device__ void window_of(const float* __restrict__ x, int n, float a, float b, float R, int mono, int* p0, int* p1) {
  const float lo = a - R, hi = b + R;
  if (mono > 0) {
    *p0 = lower_bound_f(x, n, lo, 1);
    int q = lower_bound_f(x, n, hi, 1);
    while (q < n && __ldg(x + q) <= hi) ++q;
    *p1 = q;
  } else {
    *p0 = lower_bound_f(x, n, hi, 0);
    int q = lower_bound_f(x, n, lo, 0);
    while (q < n && __ldg(x + q) >= lo) ++q;
    *p1 = q;
  }
  if (*p1 < *p0) *p1 = *p0;
}

template <int MODE>   // 0: fp32 NCHW output, 1: blocked bf16 output with the constant-1 channel
__global__ void __launch_bounds__(NT)
enc_fused_kernel(const __grid_constant__ cnp_enc_sets S, double start1, int n1, double start2, int n2, double res,
                 float eps, float* __restrict__ out_f32, long long out_bs, int c_total, cnp_blk ob, int n_chunks,
                 int max_cols, int cmax1, int CP, int ROWS) {
  extern __shared__ float sm[];
  float* outs = sm;                               // [CP][NT]
  float* scr = sm + (size_t)CP * NT;              // per-set scratch
  __shared__ float g1s[TI], g2s[TJ];
  __shared__ int p0s[TI], p1s[TI], q0s[TJ], q1s[TJ], win[4], warp_cnt[8];

  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int b = blockIdx.z, i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;
  const int i = i0 + ty, j = j0 + tx;
  if (tid < TI) g1s[tid] = cnp_grid_pt(start1, res, min(i0 + tid, n1 - 1));
  if (tid >= 32 && tid < 32 + TJ) g2s[tid - 32] = cnp_grid_pt(start2, res, min(j0 + tid - 32, n2 - 1));
  for (int c = 0; c < CP; ++c) outs[c * NT + tid] = 0.f;

  for (int k = 0; k < S.n_sets; ++k) {
    const cnp_enc_set& st = S.s[k];
    const int C = st.C;
    __syncthreads();                               // scratch of the previous set is free; g1s / g2s are visible
    if (st.kind == 2) {
      if (i < n1 && j < n2)
        for (int c = 0; c < C; ++c)
          outs[(st.ch_off + c) * NT + tid] = __ldg(st.y + ((size_t)c * n1 + i) * n2 + j);
      continue;
    }
    const float scale2 = st.scale2;
    const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
    float acc[MAXC1];
#pragma unroll
    for (int c = 0; c < MAXC1; ++c) acc[c] = 0.f;

    if (st.kind == 1) {
      // ---------------- gridded set ----------------
      const int N1 = st.N1, N2 = st.N2, KBs = st.KB;
      float* yv = scr;                                           // [cmax1][ROWS][max_cols]
      float* T = yv + (size_t)cmax1 * ROWS * max_cols;           // [cmax1][ROWS][TJ]
      float* w2b = T + (size_t)cmax1 * ROWS * TJ;                // [KBMAX][TJ]
      float* w1c = w2b + KBMAX * TJ;                             // [ROWS][TI]
      const int* ti = st.tab_i;
      const float* tw1 = st.tab_w;
      const float* tw2 = st.tab_w + (size_t)KBs * n1;
      if (ty < 2) {                                  // warp 0: row bands of the tile, warp 1: column bands
        int lo = 0x7fffffff, hi = 0;
        if (ty == 0) {
          if (tx < TI) {
            const int ii = min(i0 + tx, n1 - 1), s0 = __ldg(ti + ii), l = __ldg(ti + n1 + ii);
            p0s[tx] = s0; p1s[tx] = s0 + l;
            if (i0 + tx < n1 && l > 0) { lo = s0; hi = s0 + l; }
          }
        } else {
          const int jj = min(j0 + tx, n2 - 1), s0 = __ldg(ti + 2 * n1 + jj), l = __ldg(ti + 2 * n1 + n2 + jj);
          q0s[tx] = s0; q1s[tx] = s0 + l;
          if (j0 + tx < n2 && l > 0) { lo = s0; hi = s0 + l; }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (tx == 0) { win[2 * ty] = (hi > 0) ? lo : 0; win[2 * ty + 1] = (hi > 0) ? hi : 0; }
      }
      for (int e = tid; e < KBMAX * TJ; e += NT) {
        const int kk = e / TJ, jj = min(j0 + (e - kk * TJ), n2 - 1);
        w2b[e] = (kk < KBs) ? __ldg(tw2 + (size_t)kk * n2 + jj) : 0.f;      // zero beyond the band (table padding)
      }
      __syncthreads();
      const int plo = win[0], phi = win[1], qlo = win[2];
      const int ncols = min(win[3] - win[2], max_cols);
      const int myq = q0s[tx] - qlo;
      const int mylen = max(0, min(min(q1s[tx] - q0s[tx], KBs), ncols - myq));
      const size_t plane = (size_t)N1 * N2;
      const float* yb = st.y + (st.batched ? (size_t)b * C * plane : 0);
      const float* mb = st.mask ? st.mask + (st.batched ? (size_t)b * plane : 0) : nullptr;
      const size_t cstride = (size_t)ROWS * max_cols;
      for (int pc = plo; pc < phi; pc += ROWS) {
        const int np = min(ROWS, phi - pc);
        if (pc > plo) __syncthreads();             // the previous chunk's T / yv / w1c readers are done
        for (int e = tid; e < np * ncols; e += NT) {
          const int r = e / ncols, q = e - r * ncols;
          const size_t off = (size_t)(pc + r) * N2 + qlo + q;
          float valid = mb ? __ldg(mb + off) : 1.f;
          float v[MAXC1 - 1];
          bool nan_any = false;
#pragma unroll
          for (int c = 0; c < MAXC1 - 1; ++c)
            if (c < C) { v[c] = __ldg(yb + (size_t)c * plane + off); nan_any |= isnan(v[c]); }
          if (nan_any) valid = 0.f;
          float* d = yv + (size_t)r * max_cols + q;
          d[0] = valid;
#pragma unroll
          for (int c = 0; c < MAXC1 - 1; ++c)
            if (c < C) d[(size_t)(1 + c) * cstride] = nan_any ? 0.f : v[c] * valid;
        }
        for (int e = tid; e < np * TI; e += NT) {
          const int r = e / TI, ii = e - r * TI, kk = pc + r - p0s[ii];
          w1c[e] = (kk >= 0 && kk < p1s[ii] - p0s[ii] && kk < KBs)
                       ? __ldg(tw1 + (size_t)kk * n1 + min(i0 + ii, n1 - 1)) : 0.f;
        }
        __syncthreads();
        // horizontal band pass: T[c][r][tx] = sum_k yv[c][r][myq + k] w2b[k][tx]
        for (int r = ty; r < np; r += TI) {
          float Tr[MAXC1];
#pragma unroll
          for (int c = 0; c < MAXC1; ++c) Tr[c] = 0.f;
          const float* yr = yv + (size_t)r * max_cols + myq;
          for (int kk = 0; kk < mylen; ++kk) {
            const float w = w2b[kk * TJ + tx];
#pragma unroll
            for (int c = 0; c < MAXC1; ++c)
              if (c <= C) Tr[c] = fmaf(yr[(size_t)c * cstride + kk], w, Tr[c]);
          }
#pragma unroll
          for (int c = 0; c < MAXC1; ++c)
            if (c <= C) T[((size_t)c * ROWS + r) * TJ + tx] = Tr[c];
        }
        __syncthreads();
        // vertical pass: this thread's pixel
        for (int r = 0; r < np; ++r) {
          const float w = w1c[r * TI + ty];
          if (w != 0.f) {
#pragma unroll
            for (int c = 0; c < MAXC1; ++c)
              if (c <= C) acc[c] = fmaf(w, T[((size_t)c * ROWS + r) * TJ + tx], acc[c]);
          }
        }
      }
    } else {
      // ---------------- off-grid set ----------------
      // one global round trip per chunk of 256 points: every thread loads its point (coordinates, all channels, mask),
      // the points that can touch the tile are compacted in order into shared memory, their separable weights are
      // staged 64 points at a time and every thread sums its own pixel
      const int N = st.N1;
      float* px = scr;                     // [2][OGC] compacted coordinates
      float* ys = px + 2 * OGC;            // [MAXC1][OGC] compacted [valid ; y * valid]
      float* w1s = ys + MAXC1 * OGC;       // [OGS][TI]
      float* w2s = w1s + OGS * TI;         // [OGS][TJ + 1]
      const float a1 = g1s[0], b1 = g1s[TI - 1], a2 = g2s[0], b2 = g2s[TJ - 1];
      const float* xb = st.x1 + (size_t)b * 2 * N;
      const float* yb = st.y + (size_t)b * C * N;
      const float* mb = st.mask ? st.mask + (size_t)b * N : nullptr;
      for (int c0 = 0; c0 < N; c0 += OGC) {
        const int n = c0 + tid;
        bool keep = false;
        float p1 = 0.f, p2 = 0.f, valid = 0.f, v[MAXC1 - 1];
        if (n < N) {
          p1 = __ldg(xb + n); p2 = __ldg(xb + N + n);
          valid = mb ? __ldg(mb + n) : 1.f;
          bool nan_any = false;
#pragma unroll
          for (int c = 0; c < MAXC1 - 1; ++c)
            if (c < C) { v[c] = __ldg(yb + (size_t)c * N + n); nan_any |= isnan(v[c]); }
          if (nan_any) valid = 0.f;
#pragma unroll
          for (int c = 0; c < MAXC1 - 1; ++c)
            if (c < C) v[c] = nan_any ? 0.f : v[c] * valid;
          keep = (p1 >= a1 - R) && (p1 <= b1 + R) && (p2 >= a2 - R) && (p2 <= b2 + R);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (c0 > 0) __syncthreads();       // the previous chunk's consumers are done
        if (tx == 0) warp_cnt[ty] = __popc(bal);
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { if (w < ty) base += warp_cnt[w]; total += warp_cnt[w]; }
        if (keep) {
          const int m = base + __popc(bal & ((1u << tx) - 1u));
          px[m] = p1; px[OGC + m] = p2; ys[m] = valid;
#pragma unroll
          for (int c = 0; c < MAXC1 - 1; ++c)
            if (c < C) ys[(1 + c) * OGC + m] = v[c];
        }
        for (int m0 = 0; m0 < total; m0 += OGS) {
          const int nm = min(OGS, total - m0);
          __syncthreads();                 // compacted points visible / previous weights consumed
          for (int e = tid; e < nm * TJ; e += NT) {
            const int m = e / TJ, jj = e - m * TJ;
            w2s[m * (TJ + 1) + jj] = cnp_rbf(px[OGC + m0 + m], g2s[jj], scale2);
            if (jj < TI) w1s[m * TI + jj] = cnp_rbf(px[m0 + m], g1s[jj], scale2);
          }
          __syncthreads();
          for (int m = 0; m < nm; ++m) {
            const float w = w1s[m * TI + ty] * w2s[m * (TJ + 1) + tx];
#pragma unroll
            for (int c = 0; c < MAXC1; ++c)
              if (c <= C) acc[c] = fmaf(ys[c * OGC + m0 + m], w, acc[c]);
          }
        }
      }
    }
    // density first, data divided by (density + eps)
    const float dens = acc[0], den = dens + eps;
    outs[st.ch_off * NT + tid] = dens;
#pragma unroll
    for (int c = 1; c < MAXC1; ++c)
      if (c <= C) outs[(st.ch_off + c) * NT + tid] = acc[c] / den;
  }

  if (i >= n1 || j >= n2) return;
  if (MODE == 0) {
    float* o = out_f32 + (size_t)b * out_bs + (size_t)i * n2 + j;
    for (int c = 0; c < c_total; ++c) o[(size_t)c * n1 * n2] = outs[c * NT + tid];
  } else {
    outs[c_total * NT + tid] = 1.f;       // constant-1 channel of the folded first layer (0 in the pad, like the image)
    const int Hp = ob.H + 4, Wp = ob.W + 4;
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(ob.base) + (size_t)b * ob.bstride;
    for (int ch = 0; ch < n_chunks; ++ch) {
      __align__(16) __nv_bfloat16 pk[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) pk[e] = __float2bfloat16(outs[(ch * 8 + e) * NT + tid]);
      *reinterpret_cast<uint4*>(base + (((size_t)(ob.cb_off + ch) * Hp + i + 2) * Wp + j + 2) * 8) =
          *reinterpret_cast<const uint4*>(pk);
    }
  }
}

// ---- band tables: depend only on (coordinates, internal grid, length scale) ------------------------------------
__global__ void __launch_bounds__(128)
enc_tables_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int N1, int N2, int mono1, int mono2,
                  double start1, int n1, double start2, int n2, double res, float scale2, int KBs,
                  int* __restrict__ tab_i, float* __restrict__ tab_w) {
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= n1 + n2) return;
  const bool dim2 = e >= n1;
  const int j = dim2 ? e - n1 : e, n = dim2 ? n2 : n1, N = dim2 ? N2 : N1;
  const float* x = dim2 ? x2 : x1;
  const float g = cnp_grid_pt(dim2 ? start2 : start1, res, j);
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  int lo, hi;
  window_of(x, N, g, g, R, dim2 ? mono2 : mono1, &lo, &hi);
  const int len = min(hi - lo, KBs);
  int* ti = tab_i + (dim2 ? 2 * n1 : 0);
  float* tw = tab_w + (dim2 ? (size_t)KBs * n1 : 0);
  ti[j] = lo; ti[n + j] = len;
  for (int k = 0; k < KBs; ++k) tw[(size_t)k * n + j] = (k < len) ? cnp_rbf(x[lo + k], g, scale2) : 0.f;
}

}  // namespace

// tab_i: 2 * (n1 + n2) ints, tab_w: band * (n1 + n2) floats.  band = upper bound (<= 32) of the number of inputs within
// the truncation radius of any grid point along either dimension (host-side, from the coordinates).
CNP_API int cnp_encode_tables(const float* x1, const float* x2, int N1, int N2, int mono1, int mono2, double start1, int n1,
                              double start2, int n2, double res, float scale2, int band, int* tab_i, float* tab_w,
                              cudaStream_t st) {
  CNP_REQUIRE(x1 && x2 && tab_i && tab_w && N1 > 0 && N2 > 0 && n1 > 0 && n2 > 0, "encode_tables: bad arguments");
  CNP_REQUIRE(mono1 != 0 && mono2 != 0, "encode_tables: coordinates must be monotone");
  CNP_REQUIRE(band >= 1 && band <= KBMAX, "encode_tables: band %d outside 1..%d", band, KBMAX);
  enc_tables_kernel<<<cnp_cdiv(n1 + n2, 128), 128, 0, st>>>(x1, x2, N1, N2, mono1, mono2, start1, n1, start2, n2, res,
                                                            scale2, band, tab_i, tab_w);
  CNP_LAUNCH_CHECK("enc_tables_kernel");
  return 0;
}

static long long ef_smem(int channels_staged, int cmax1, int max_cols, int rows) {
  if (cmax1 < 1) cmax1 = 1;
  const long long grid_f = (long long)cmax1 * rows * max_cols + (long long)cmax1 * rows * TJ + KBMAX * TJ + rows * TI;
  const long long og_f = 2LL * OGC + (long long)MAXC1 * OGC + OGS * TI + OGS * (TJ + 1);
  return ((long long)channels_staged * NT + (grid_f > og_f ? grid_f : og_f)) * 4;
}
// rows staged per step: 32 when the launch then still fits 3 CTAs per SM, else 16
static int ef_rows(int channels_staged, int cmax1, int max_cols) {
  return ef_smem(channels_staged, cmax1, max_cols, 32) <= 72 * 1024 ? 32 : 16;
}

// Shared memory of one launch (bytes) for the given staging geometry; -1 when it cannot fit.
CNP_API long long cnp_encode_fused_smem_bytes(int channels_staged, int cmax1, int max_cols) {
  const long long bytes = ef_smem(channels_staged, cmax1, max_cols, ef_rows(channels_staged, cmax1, max_cols));
  return bytes <= 200 * 1024 ? bytes : -1;
}

// mode 0: out_f32 [B][c_total][n1][n2] (batch stride out_bstride floats); mode 1: out_blk = blocked bf16 view with
// n_chunks >= (c_total + 1 + 7) / 8 chunks: channels [0, c_total) = the encoder output, channel c_total = 1, rest 0.
// max_cols: upper bound of the number of input columns inside the band of any 32 consecutive grid columns, over the
// gridded sets (host-side, from the coordinates; the kernel clamps to it).
CNP_API int cnp_encode_fused(const cnp_enc_sets* sets, int B, double start1, int n1, double start2, int n2, double res,
                             float eps, int mode, float* out_f32, long long out_bstride, int c_total,
                             const cnp_blk* out_blk, int n_chunks, int max_cols, cudaStream_t st) {
  CNP_REQUIRE(sets && sets->n_sets >= 1 && sets->n_sets <= 8 && B > 0 && n1 > 0 && n2 > 0, "encode_fused: bad arguments");
  CNP_REQUIRE(mode == 0 ? out_f32 != nullptr : (out_blk != nullptr && n_chunks * 8 >= c_total + 1),
              "encode_fused: output does not match mode %d", mode);
  int cmax1 = 1;
  for (int k = 0; k < sets->n_sets; ++k) {
    const cnp_enc_set& s = sets->s[k];
    CNP_REQUIRE(s.kind >= 0 && s.kind <= 2 && s.C >= 0 && s.ch_off >= 0, "encode_fused: set %d malformed", k);
    CNP_REQUIRE(s.ch_off + s.C + (s.kind == 2 ? 0 : 1) <= c_total, "encode_fused: set %d exceeds %d channels", k, c_total);
    if (s.kind != 2) CNP_REQUIRE(s.C <= MAXC1 - 1, "encode_fused: set %d has more than %d channels", k, MAXC1 - 1);
    if (s.kind == 1) {
      CNP_REQUIRE(s.y && s.tab_i && s.tab_w && s.KB >= 1 && s.KB <= KBMAX, "encode_fused: gridded set %d needs band tables", k);
      if (s.C + 1 > cmax1) cmax1 = s.C + 1;
    }
    if (s.kind == 0) CNP_REQUIRE(s.N1 == 0 || (s.x1 && s.y), "encode_fused: off-grid set %d has null inputs", k);
  }
  if (mode == 1) CNP_REQUIRE(out_blk->H == n1 && out_blk->W == n2, "encode_fused: blocked output geometry mismatch");
  if (max_cols < 1) max_cols = 1;
  const int CP = mode == 1 ? n_chunks * 8 : c_total;
  const int rows = ef_rows(CP, cmax1, max_cols);
  const long long smem = ef_smem(CP, cmax1, max_cols, rows);
  CNP_REQUIRE(smem <= 200 * 1024, "encode_fused: staging of %d channels x %d columns does not fit in shared memory", cmax1, max_cols);
  static long long attr[2] = {0, 0};
  if (smem > attr[mode] && smem > 48 * 1024) {
    cudaError_t e = mode == 0
        ? cudaFuncSetAttribute(enc_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        : cudaFuncSetAttribute(enc_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cnp_set_error("encode_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr[mode] = smem;
  }
  dim3 grid(cnp_cdiv(n2, TJ), cnp_cdiv(n1, TI), B);
  cnp_blk ob;
  memset(&ob, 0, sizeof(ob));
  if (out_blk) ob = *out_blk;
  if (mode == 0)
    enc_fused_kernel<0><<<grid, NT, smem, st>>>(*sets, start1, n1, start2, n2, res, eps, out_f32, out_bstride, c_total, ob,
                                                n_chunks, max_cols, cmax1, CP, rows);
  else
    enc_fused_kernel<1><<<grid, NT, smem, st>>>(*sets, start1, n1, start2, n2, res, eps, out_f32, out_bstride, c_total, ob,
                                                n_chunks, max_cols, cmax1, CP, rows);
  CNP_LAUNCH_CHECK("enc_fused_kernel");
  return 0;
}
