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
// Three launches per step, each a flat data-parallel sweep (the work is tiny -- 60 MB of algorithmic traffic per 16-task
// step -- so what matters is exposing all of it at once and keeping every thread's instruction stream short):
//   1. cnp_encode_hpass: horizontal band pass of EVERY gridded set, T_k[b][c][p][j] = sum_q y~[b][c][p][q] w2[q][j]
//      (<= 32 taps, one thread per (b, p, j), all channels) into a workspace that stays in L2.
//   2. cnp_encode_vpass: vertical band pass + density normalisation, one thread per (b, i, j), all channels:
//      V_k[b][c][i][j] (fp32 planes; in fp32 mode these ARE the output channels).
//   3. cnp_encode_fused: one CTA = one 16 x 32 tile of one task: the off-grid sets (every thread loads one point --
//      coordinates, channels, mask --, the points that can touch the tile are compacted in order, their separable weights
//      staged in shared memory, every thread sums its own 4 pixels), then -- bf16 UNet -- the tile's channels are
//      gathered from the V planes (coalesced, L2) and leave directly as the blocked bf16 tensor [B][chunk][H+4][W+4][8]
//      with the constant-1 channel of the folded first layer (fold_in.cu): no fp32 encoder tensor, no layout-conversion
//      kernel.
// Band starts and weights come from per-set tables (cnp_encode_tables) that depend only on (coordinates, internal grid,
// length scale): the caller builds them once and keeps them across steps.  Gather form throughout, no atomics.
// Bound: HBM.  History (B200, B = 16, 304^2, ncu): per-set kernels of round 1 240 us; fused per-tile
// kernel recomputing the horizontal pass in shared memory 312 us (166 M warp instructions, instruction-bound: every
// 8 x 32 tile redid ~3x the horizontal work in 9-channel predicated loops); horizontal pass split off, 4 rows per
// thread 158 us; 4 tasks per CTA 119 us (latency-bound at 12 resident warps per SM); this version: profiles/.
#include "tc_common.cuh"
#include <math.h>

struct cnp_enc_set {
  int kind;          // 0 off-grid, 1 gridded
  int C;             // data channels (1..8)
  int ch_off;        // first output channel (the density channel)
  int batched;       // y (and mask, T, V) carry a batch axis; 0: one field shared by every task
  const float* x1;   // off-grid: x [B,2,N]
  const float* x2;   // unused by the kernels (gridded coordinates enter through the tables)
  const float* y;    // gridded [B or 1, C, N1, N2]; off-grid [B, C, N]
  const float* mask; // gridded [B or 1, 1, N1, N2] | off-grid [B, 1, N] | NULL
  int N1, N2;        // off-grid: N1 = N
  int mono1, mono2;  // +1 ascending, -1 descending (gridded coordinates)
  float scale2;      // s_k^2
  int KB;            // gridded: band width of the tables (<= 32)
  const int* tab_i;  // gridded: [p0 (n1) | len1 (n1) | q0 (n2) | len2 (n2)]   (cnp_encode_tables)
  const float* tab_w;// gridded: [w1 (KB x n1) | w2 (KB x n2)]
  float* T;          // gridded: horizontal-pass workspace [B or 1][C+1][N1][n2]
  float* V;          // gridded: normalised output planes, channel stride n1*n2, batch stride V_bs
  long long V_bs;
};
struct cnp_enc_sets {
  int n_sets;
  int pad_;
  cnp_enc_set s[8];
};

namespace {

constexpr int TI = 16, TJ = 32, NT = 128, RPT = 4;   // tile rows / cols, threads, rows per thread (launch 3)
constexpr int NPIX = TI * TJ;
constexpr int KBMAX = 32;    // max band (inputs within R of one grid row / column)
constexpr int OGC = NT;      // off-grid points per chunk (one per thread)
constexpr int OGS = 64;      // off-grid points whose weights are staged at a time
constexpr int MAXC1 = 9;     // max channels incl. density of one set
constexpr int SW = 160;      // threads per block of the sweeps (launches 1, 2): 304 columns = 2 blocks, 95 % of the lanes
constexpr int MAXCH = 72;    // max output channels (8 sets x 9)

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

// =====================================================================================================================
// launches 1 and 2: one block = SW consecutive columns of one (set, batch entry, row)
// =====================================================================================================================
struct cnp_sweep_args {
  cnp_enc_sets S;    // gridded sets only
  int blk0[9];       // first block of every set (prefix sums); blk0[n_sets] = grid size
  int rows[8];       // rows swept per batch entry: N1 (horizontal pass) or n1 (vertical pass)
  int n1, n2, bpr;   // bpr = blocks per row
};

// block -> (set, batch entry, row, column); block-uniform integer work only
__device__ __forceinline__ const cnp_enc_set& sweep_decode(const cnp_sweep_args& a, int* b, int* row, int* j) {
  int k = 0;
  while (k + 1 < a.S.n_sets && (int)blockIdx.x >= a.blk0[k + 1]) ++k;
  const int lb = (int)blockIdx.x - a.blk0[k];
  const int pr = lb / a.bpr;
  *j = (lb - pr * a.bpr) * (int)blockDim.x + threadIdx.x;
  *b = pr / a.rows[k];
  *row = pr - *b * a.rows[k];
  return a.S.s[k];
}

// one thread = one grid column j of HR consecutive input rows: the band of a column (start, length, tap weights) is
// looked up once and shared by the rows
constexpr int HR = 1;   // (2 rows per thread measured no faster: 42 vs 39 us, more registers)
template <int C, bool MASK>
__device__ __forceinline__ void hpass_elem(const cnp_enc_set& st, int b, int p0r, int j, int n1, int n2) {
  const int N1 = st.N1, N2 = st.N2;
  const int q0 = __ldg(st.tab_i + 2 * n1 + j), len = __ldg(st.tab_i + 2 * n1 + n2 + j);
  const float* w2 = st.tab_w + st.KB * n1 + j;
  const int plane = N1 * N2;                       // (a context field has far fewer than 2^30 cells)
  const int nr = min(HR, N1 - p0r);
  const float* yb = st.y + (size_t)b * C * plane + p0r * N2 + q0;
  const float* mb = MASK ? st.mask + (size_t)b * plane + p0r * N2 + q0 : nullptr;
  float acc[HR][C + 1];
#pragma unroll
  for (int r = 0; r < HR; ++r)
#pragma unroll
    for (int c = 0; c <= C; ++c) acc[r][c] = 0.f;
#pragma unroll 2
  for (int k = 0; k < len; ++k) {
    const float w = __ldg(w2 + k * n2);
#pragma unroll
    for (int r = 0; r < HR; ++r) {
      if (r < nr) {
        float v[C];
        bool nan_any = false;
#pragma unroll
        for (int c = 0; c < C; ++c) { v[c] = __ldg(yb + c * plane + r * N2 + k); nan_any |= isnan(v[c]); }
        float valid = MASK ? __ldg(mb + r * N2 + k) : 1.f;
        if (nan_any) valid = 0.f;
        acc[r][0] = fmaf(valid, w, acc[r][0]);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[r][1 + c] = fmaf(nan_any ? 0.f : (MASK ? v[c] * valid : v[c]), w, acc[r][1 + c]);
      }
    }
  }
  const int tplane = N1 * n2;
  float* T = st.T + (size_t)b * (C + 1) * tplane + p0r * n2 + j;
#pragma unroll
  for (int r = 0; r < HR; ++r)
    if (r < nr) {
#pragma unroll
      for (int c = 0; c <= C; ++c) T[c * tplane + r * n2] = acc[r][c];
    }
}

#define CNP_SWITCH_C(Cv, ...)                     \
  switch (Cv) {                                   \
    case 1: { constexpr int CC = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int CC = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int CC = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int CC = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int CC = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int CC = 6; __VA_ARGS__; } break; \
    case 7: { constexpr int CC = 7; __VA_ARGS__; } break; \
    default: { constexpr int CC = 8; __VA_ARGS__; } break; \
  }

// (One instantiation per channel count, launched back to back, was measured SLOWER: the launches of the small shared
// sets serialise behind the big one -- 54 + 51 us against 39 + 37 us for one launch per sweep that switches over 1..8
// channels at the 8-channel register allocation.)
__global__ void __launch_bounds__(SW)
enc_hpass_kernel(const __grid_constant__ cnp_sweep_args a) {
  int b, p, j;
  const cnp_enc_set& st = sweep_decode(a, &b, &p, &j);
  if (j >= a.n2) return;
  if (st.mask) { CNP_SWITCH_C(st.C, hpass_elem<CC, true>(st, b, p * HR, j, a.n1, a.n2)) }
  else { CNP_SWITCH_C(st.C, hpass_elem<CC, false>(st, b, p * HR, j, a.n1, a.n2)) }
}

// one thread = 4 adjacent columns of one (task, grid row): the T rows are read with 16 B loads (n2 % 4 == 0), the
// weight of a tap is the same for the 4 outputs
template <int C>
__device__ __forceinline__ void vpass_elem4(const cnp_enc_set& st, int b, int i, int j, int n1, int n2, float eps) {
  const int p0 = __ldg(st.tab_i + i), len = __ldg(st.tab_i + n1 + i);
  const float* w1 = st.tab_w + i;
  const int tplane = st.N1 * n2;
  const float* T = st.T + (size_t)b * (C + 1) * tplane + p0 * n2 + j;
  float4 acc[C + 1];
#pragma unroll
  for (int c = 0; c <= C; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (int k = 0; k < len; ++k, T += n2) {
    const float w = __ldg(w1 + k * n1);
#pragma unroll
    for (int c = 0; c <= C; ++c) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(T + c * tplane));
      acc[c].x = fmaf(w, t.x, acc[c].x); acc[c].y = fmaf(w, t.y, acc[c].y);
      acc[c].z = fmaf(w, t.z, acc[c].z); acc[c].w = fmaf(w, t.w, acc[c].w);
    }
  }
  // density first, data divided by (density + eps)
  const int plane = n1 * n2;
  float* V = st.V + (size_t)b * st.V_bs + i * n2 + j;
  const float4 d = acc[0];
  const float4 inv = make_float4(1.0f / (d.x + eps), 1.0f / (d.y + eps), 1.0f / (d.z + eps), 1.0f / (d.w + eps));
  *reinterpret_cast<float4*>(V) = d;
#pragma unroll
  for (int c = 1; c <= C; ++c)
    *reinterpret_cast<float4*>(V + c * plane) =
        make_float4(acc[c].x * inv.x, acc[c].y * inv.y, acc[c].z * inv.z, acc[c].w * inv.w);
}

template <int C>
__device__ __forceinline__ void vpass_elem(const cnp_enc_set& st, int b, int i, int j, int n1, int n2, float eps) {
  const int p0 = __ldg(st.tab_i + i), len = __ldg(st.tab_i + n1 + i);
  const float* w1 = st.tab_w + i;
  const int tplane = st.N1 * n2;
  const float* T = st.T + (size_t)b * (C + 1) * tplane + p0 * n2 + j;
  float acc[C + 1];
#pragma unroll
  for (int c = 0; c <= C; ++c) acc[c] = 0.f;
#pragma unroll 4
  for (int k = 0; k < len; ++k) {
    const float w = __ldg(w1 + k * n1);
#pragma unroll
    for (int c = 0; c <= C; ++c) acc[c] = fmaf(w, __ldg(T + c * tplane + k * n2), acc[c]);
  }
  const int plane = n1 * n2;
  float* V = st.V + (size_t)b * st.V_bs + i * n2 + j;
  const float dens = acc[0], inv = 1.0f / (dens + eps);
  V[0] = dens;
#pragma unroll
  for (int c = 1; c <= C; ++c) V[c * plane] = acc[c] * inv;
}

// VEC: the block sweeps 4 columns per thread (needs n2 % 4 == 0 and 16 B aligned T / V planes)
template <bool VEC>
__global__ void __launch_bounds__(SW, 4)
enc_vpass_kernel(const __grid_constant__ cnp_sweep_args a, float eps) {
  int b, i, j;
  const cnp_enc_set& st = sweep_decode(a, &b, &i, &j);
  if (VEC) {
    j *= 4;
    if (j >= a.n2) return;
    CNP_SWITCH_C(st.C, vpass_elem4<CC>(st, b, i, j, a.n1, a.n2, eps))
  } else {
    if (j >= a.n2) return;
    CNP_SWITCH_C(st.C, vpass_elem<CC>(st, b, i, j, a.n1, a.n2, eps))
  }
}

// =====================================================================================================================
// launch 3: off-grid sets + assembly of the UNet input
// =====================================================================================================================
struct cnp_asm_args {
  cnp_enc_sets S;                 // off-grid sets only; ch_off = REAL output channel, og_off[] = slot in the staging tile
  int og_off[8];
  int n_og_ch;                    // staged off-grid channels
  const float* cptr[MAXCH];       // per output channel: V plane of a gridded set, or NULL
  long long cbs[MAXCH];           // its batch stride (0: shared by every task)
  short og_map[MAXCH];            // per output channel: slot in the off-grid staging tile, or -1
};

template <int C>
__device__ __forceinline__ void store_set(float* __restrict__ o, const float (&acc)[RPT][MAXC1], float eps) {
#pragma unroll
  for (int e = 0; e < RPT; ++e) {
    const float dens = acc[e][0], inv = 1.0f / (dens + eps);
    o[e * TJ] = dens;
#pragma unroll
    for (int c = 1; c <= C; ++c) o[c * NPIX + e * TJ] = acc[e][c] * inv;
  }
}

template <int C>
__device__ __forceinline__ void ogacc(const float* __restrict__ ys, const float* __restrict__ w1s,
                                      const float* __restrict__ w2s, int m0, int nm, int row0, int tx,
                                      float (&acc)[RPT][MAXC1]) {
  for (int m = 0; m < nm; ++m) {
    const float w2 = w2s[m * (TJ + 1) + tx];
    float yv[C + 1];
#pragma unroll
    for (int c = 0; c <= C; ++c) yv[c] = ys[c * OGC + m0 + m];
#pragma unroll
    for (int e = 0; e < RPT; ++e) {
      const float w = w1s[m * TI + row0 + e] * w2;
#pragma unroll
      for (int c = 0; c <= C; ++c) acc[e][c] = fmaf(yv[c], w, acc[e][c]);
    }
  }
}

template <int MODE>   // 0: fp32 NCHW output (off-grid channels only; launch 2 wrote the rest), 1: blocked bf16 output
__global__ void __launch_bounds__(NT, 6)
enc_fused_kernel(const __grid_constant__ cnp_asm_args A, double start1, int n1, double start2, int n2, double res,
                 float eps, float* __restrict__ out_f32, long long out_bs, int c_total, cnp_blk ob, int n_chunks) {
  extern __shared__ __align__(16) float sm[];
  float* og = sm;                                  // [n_og_ch][NPIX] staging of the off-grid channels of the tile
  float* scr = sm + A.n_og_ch * NPIX;              // per-set scratch
  __shared__ float g1s[TI], g2s[TJ];
  __shared__ int warp_cnt[NT / 32];

  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int b = blockIdx.z, i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;
  const int row0 = ty * RPT, j = j0 + tx;
  const bool col_ok = j < n2;
  const int pix0 = row0 * TJ + tx;                 // tile-local pixel of the thread's first row
  if (tid < TI) g1s[tid] = cnp_grid_pt(start1, res, min(i0 + tid, n1 - 1));
  if (tid >= 32 && tid < 32 + TJ) g2s[tid - 32] = cnp_grid_pt(start2, res, min(j0 + tid - 32, n2 - 1));

  for (int k = 0; k < A.S.n_sets; ++k) {
    const cnp_enc_set& st = A.S.s[k];
    const int C = st.C;
    __syncthreads();                               // scratch of the previous set is free; g1s / g2s are visible
    float acc[RPT][MAXC1];
#pragma unroll
    for (int e = 0; e < RPT; ++e)
#pragma unroll
      for (int c = 0; c < MAXC1; ++c) acc[e][c] = 0.f;
    const float scale2 = st.scale2;
    const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
    const int N = st.N1;
    float* px = scr;                     // [2][OGC] compacted coordinates
    float* ys = px + 2 * OGC;            // [MAXC1][OGC] compacted [valid ; y * valid]
    float* w1s = ys + MAXC1 * OGC;       // [OGS][TI]
    float* w2s = w1s + OGS * TI;         // [OGS][TJ + 1]
    const float a1 = g1s[0], b1 = g1s[TI - 1], a2 = g2s[0], b2 = g2s[TJ - 1];
    const float lo1 = fminf(a1, b1) - R, hi1 = fmaxf(a1, b1) + R, lo2 = fminf(a2, b2) - R, hi2 = fmaxf(a2, b2) + R;
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
          if (c < C) { v[c] = __ldg(yb + c * N + n); nan_any |= isnan(v[c]); }
        if (nan_any) valid = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC1 - 1; ++c)
          if (c < C) v[c] = nan_any ? 0.f : v[c] * valid;
        keep = (p1 >= lo1) && (p1 <= hi1) && (p2 >= lo2) && (p2 <= hi2);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (c0 > 0) __syncthreads();       // the previous chunk's consumers are done
      if (tx == 0) warp_cnt[ty] = __popc(bal);
      __syncthreads();
      int base = 0, total = 0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) { if (w < ty) base += warp_cnt[w]; total += warp_cnt[w]; }
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
          const int m = e >> 5, jj = e & 31;
          w2s[m * (TJ + 1) + jj] = cnp_rbf(px[OGC + m0 + m], g2s[jj], scale2);
          if (jj < TI) w1s[m * TI + jj] = cnp_rbf(px[m0 + m], g1s[jj], scale2);
        }
        __syncthreads();
        CNP_SWITCH_C(C, ogacc<CC>(ys, w1s, w2s, m0, nm, row0, tx, acc))
      }
    }
    CNP_SWITCH_C(C, store_set<CC>(og + A.og_off[k] * NPIX + pix0, acc, eps))
  }
  // ---- write the tile (every thread reads back only what it staged itself: no barrier) ----
  if (!col_ok) return;
  if (MODE == 0) {
    const int plane = n1 * n2;
    for (int k = 0; k < A.S.n_sets; ++k) {
      const cnp_enc_set& st = A.S.s[k];
#pragma unroll
      for (int e = 0; e < RPT; ++e) {
        const int i = i0 + row0 + e;
        if (i >= n1) continue;
        float* o = out_f32 + (size_t)b * out_bs + (size_t)st.ch_off * plane + i * n2 + j;
        for (int c = 0; c <= st.C; ++c) o[c * plane] = og[(A.og_off[k] + c) * NPIX + pix0 + e * TJ];
      }
    }
  } else {
    const int Hp = ob.H + 4, Wp = ob.W + 4;
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(ob.base) + (size_t)b * ob.bstride;
    int poff[RPT];
#pragma unroll
    for (int e = 0; e < RPT; ++e) poff[e] = min(i0 + row0 + e, n1 - 1) * n2 + j;
    for (int ch = 0; ch < n_chunks; ++ch) {
      float v[8][RPT];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = ch * 8 + q;
        const float* p = A.cptr[c];
        if (p != nullptr) {                        // channel of a gridded set: V plane (coalesced along the row, L2)
          p += (size_t)b * A.cbs[c];
#pragma unroll
          for (int e = 0; e < RPT; ++e) v[q][e] = __ldg(p + poff[e]);
        } else {
          const int slot = A.og_map[c];
          const float cst = (c == c_total) ? 1.f : 0.f;       // constant-1 channel of the folded first layer
#pragma unroll
          for (int e = 0; e < RPT; ++e) v[q][e] = slot >= 0 ? og[slot * NPIX + pix0 + e * TJ] : cst;
        }
      }
#pragma unroll
      for (int e = 0; e < RPT; ++e) {
        const int i = i0 + row0 + e;
        if (i >= n1) continue;
        uint4 pk;
        __nv_bfloat162 h;
        h = __floats2bfloat162_rn(v[0][e], v[1][e]); pk.x = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(v[2][e], v[3][e]); pk.y = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(v[4][e], v[5][e]); pk.z = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(v[6][e], v[7][e]); pk.w = *reinterpret_cast<uint32_t*>(&h);
        *reinterpret_cast<uint4*>(base + (((size_t)(ob.cb_off + ch) * Hp + i + 2) * Wp + j + 2) * 8) = pk;
      }
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

static long long ef_smem(int og_channels) {
  const long long og_f = 2LL * OGC + (long long)MAXC1 * OGC + OGS * TI + OGS * (TJ + 1);
  return ((long long)og_channels * NPIX + og_f) * 4;
}

static int ef_check_sets(const cnp_enc_sets* sets, int c_total, const char* who) {
  CNP_REQUIRE(sets && sets->n_sets >= 1 && sets->n_sets <= 8, "%s: need 1..8 context sets", who);
  for (int k = 0; k < sets->n_sets; ++k) {
    const cnp_enc_set& s = sets->s[k];
    CNP_REQUIRE((s.kind == 0 || s.kind == 1) && s.ch_off >= 0, "%s: set %d malformed", who, k);
    CNP_REQUIRE(s.C >= 1 && s.C <= MAXC1 - 1, "%s: set %d needs 1..%d channels", who, k, MAXC1 - 1);
    CNP_REQUIRE(s.ch_off + s.C + 1 <= c_total, "%s: set %d exceeds %d channels", who, k, c_total);
    if (s.kind == 1)
      CNP_REQUIRE(s.y && s.tab_i && s.tab_w && s.T && s.V && s.KB >= 1 && s.KB <= KBMAX && s.N1 > 0 && s.N2 > 0 &&
                      (long long)s.N1 * s.N2 < (1LL << 30),
                  "%s: gridded set %d needs band tables and the T / V workspaces", who, k);
    if (s.kind == 0) CNP_REQUIRE(s.N1 == 0 || (s.x1 && s.y), "%s: off-grid set %d has null inputs", who, k);
  }
  return 0;
}

static int ef_sweep_args(const cnp_enc_sets* sets, int B, int n1, int n2, bool vertical, int cols_per_block,
                         cnp_sweep_args* a) {
  memset(a, 0, sizeof(*a));
  a->n1 = n1; a->n2 = n2; a->bpr = cnp_cdiv(n2, cols_per_block);
  int n = 0;
  long long blocks = 0;
  for (int k = 0; k < sets->n_sets; ++k) {
    const cnp_enc_set& s = sets->s[k];
    if (s.kind != 1) continue;
    a->S.s[n] = s;
    a->rows[n] = vertical ? n1 : cnp_cdiv(s.N1, HR);      // the horizontal pass sweeps HR input rows per thread
    a->blk0[n] = (int)blocks;
    blocks += (long long)(s.batched ? B : 1) * a->rows[n] * a->bpr;
    CNP_REQUIRE(blocks < (1LL << 31) && (long long)(s.batched ? B : 1) * (s.C + 1) * s.N1 * n2 < (1LL << 40),
                "encode: set %d is too large", k);
    ++n;
  }
  a->S.n_sets = n;
  a->blk0[n] = (int)blocks;
  return 0;
}

// Launch 1: horizontal band pass of every gridded set of ``sets`` (off-grid sets are skipped) into their T workspaces
// ([B or 1][C+1][N1][n2] floats each).
CNP_API int cnp_encode_hpass(const cnp_enc_sets* sets, int B, int n1, int n2, cudaStream_t st) {
  CNP_REQUIRE(B > 0 && n1 > 0 && n2 > 0, "encode_hpass: bad arguments");
  if (int rc = ef_check_sets(sets, 1 << 30, "encode_hpass")) return rc;
  cnp_sweep_args a;
  if (int rc = ef_sweep_args(sets, B, n1, n2, false, SW, &a)) return rc;
  if (a.S.n_sets == 0) return 0;
  enc_hpass_kernel<<<(unsigned)a.blk0[a.S.n_sets], SW, 0, st>>>(a);
  CNP_LAUNCH_CHECK("enc_hpass_kernel");
  return 0;
}

// Launch 2: vertical band pass + density normalisation of every gridded set: T -> V (fp32 planes, channel stride n1*n2,
// batch stride V_bs; a set every task shares writes one batch entry).
CNP_API int cnp_encode_vpass(const cnp_enc_sets* sets, int B, int n1, int n2, float eps, cudaStream_t st) {
  CNP_REQUIRE(B > 0 && n1 > 0 && n2 > 0, "encode_vpass: bad arguments");
  if (int rc = ef_check_sets(sets, 1 << 30, "encode_vpass")) return rc;
  bool vec = (n2 % 4 == 0);
  for (int k = 0; k < sets->n_sets; ++k) {
    const cnp_enc_set& s = sets->s[k];
    if (s.kind != 1) continue;
    vec = vec && ((uintptr_t)s.T % 16 == 0) && ((uintptr_t)s.V % 16 == 0) && (s.V_bs % 4 == 0) &&
          (((long long)s.N1 * n2) % 4 == 0) && (((long long)n1 * n2) % 4 == 0);
  }
  // vector form: a block of `threads` lanes covers 4 * threads columns of one row (304 columns: 96 lanes, 79 % active)
  int threads = SW;
  if (vec) { threads = ((cnp_cdiv(n2, 4) + 31) / 32) * 32; if (threads > SW) threads = SW; }
  cnp_sweep_args a;
  if (int rc = ef_sweep_args(sets, B, n1, n2, true, vec ? 4 * threads : threads, &a)) return rc;
  if (a.S.n_sets == 0) return 0;
  if (vec) enc_vpass_kernel<true><<<(unsigned)a.blk0[a.S.n_sets], threads, 0, st>>>(a, eps);
  else enc_vpass_kernel<false><<<(unsigned)a.blk0[a.S.n_sets], threads, 0, st>>>(a, eps);
  CNP_LAUNCH_CHECK("enc_vpass_kernel");
  return 0;
}

// Launch 3.  mode 0: writes the channels of the OFF-GRID sets into out_f32 [B][c_total][n1][n2] (batch stride out_bstride
// floats; the gridded sets' V planes point into the same tensor and were written by launch 2).  mode 1: out_blk = blocked
// bf16 view with n_chunks >= (c_total + 1 + 7) / 8 chunks: channels [0, c_total) = the encoder output (gridded channels
// gathered from the V planes), channel c_total = 1, rest 0.
CNP_API int cnp_encode_fused(const cnp_enc_sets* sets, int B, double start1, int n1, double start2, int n2, double res,
                             float eps, int mode, float* out_f32, long long out_bstride, int c_total,
                             const cnp_blk* out_blk, int n_chunks, cudaStream_t st) {
  CNP_REQUIRE(B > 0 && n1 > 0 && n2 > 0 && (mode == 0 || mode == 1), "encode_fused: bad arguments");
  CNP_REQUIRE(mode == 0 ? out_f32 != nullptr : (out_blk != nullptr && n_chunks * 8 >= c_total + 1),
              "encode_fused: output does not match mode %d", mode);
  CNP_REQUIRE(c_total + 1 <= MAXCH && n_chunks * 8 <= MAXCH, "encode_fused: more than %d channels", MAXCH);
  if (int rc = ef_check_sets(sets, c_total, "encode_fused")) return rc;
  if (mode == 1) CNP_REQUIRE(out_blk->H == n1 && out_blk->W == n2, "encode_fused: blocked output geometry mismatch");
  cnp_asm_args A;
  memset(&A, 0, sizeof(A));
  for (int c = 0; c < MAXCH; ++c) A.og_map[c] = -1;
  int n = 0, slot = 0;
  for (int k = 0; k < sets->n_sets; ++k) {
    const cnp_enc_set& s = sets->s[k];
    if (s.kind == 1) {
      for (int c = 0; c <= s.C; ++c) {
        A.cptr[s.ch_off + c] = s.V + (size_t)c * n1 * n2;
        A.cbs[s.ch_off + c] = s.batched ? s.V_bs : 0;
      }
    } else {
      A.S.s[n] = s;
      A.og_off[n] = slot;
      for (int c = 0; c <= s.C; ++c) A.og_map[s.ch_off + c] = (short)(slot + c);
      slot += s.C + 1;
      ++n;
    }
  }
  A.S.n_sets = n;
  A.n_og_ch = slot;
  if (mode == 0 && n == 0) return 0;               // nothing left to write
  const long long smem = ef_smem(slot);
  CNP_REQUIRE(smem <= 200 * 1024, "encode_fused: %d off-grid channels do not fit in shared memory", slot);
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
    enc_fused_kernel<0><<<grid, NT, smem, st>>>(A, start1, n1, start2, n2, res, eps, out_f32, out_bstride, c_total, ob,
                                                n_chunks);
  else
    enc_fused_kernel<1><<<grid, NT, smem, st>>>(A, start1, n1, start2, n2, res, eps, out_f32, out_bstride, c_total, ob,
                                                n_chunks);
  CNP_LAUNCH_CHECK("enc_fused_kernel");
  return 0;
}
