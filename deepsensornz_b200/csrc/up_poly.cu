// Support kernels of the polyphase resize-convolution (the decomposition is stated and checked in float64
// in tools/polyphase_strips.py).
//
// Replaces, together with the conv_tc2 kinds UP_PHASE / UP_PHASE_DGRAD and the wgrad kind WG_UP_PHASE, the
// torch.nn.Upsample(scale_factor=2, mode="bilinear") + Conv2d(k=5, padding=2) pairs of upstream neuralprocesses'
// UNet decoder levels (resize_convs=True; SURVEY.md A.4 / U10, reached from ConvNP.loss_fn
// nzdownscale/downscaler/train.py:370) WITHOUT materialising the upsampled tensor:
//   * low-res pixels (i, j) with 2 <= i < H-2, 2 <= j < W-2 own high-res outputs that never see the clamped border of
//     the bilinear map nor the zero padding of the upsampled tensor: four 4x4 phase convolutions of x itself;
//   * the band of 2 low-res (4 high-res) pixels along every edge is computed by the STANDARD 5x5 kernels on strips of
//     the upsampled tensor -- 6 high-res rows (top, bottom) or 6 high-res columns stored transposed (left, right) -- so
//     that the zero padding and the clamping are exactly the reference's.
// Strip tensors are ordinary blocked bf16 images [2B][C/8][6+4][L+4][8]: image s*B + b, s = 0 the low edge (rows / columns
// 0..5), s = 1 the high edge (rows / columns 2H-6 .. 2H-1); row strips L = 2W, column strips L = 2H (strip row = column).
#include "tc_common.cuh"

#define UP_SR 6     // strip height (high-res pixels): 4 band outputs + the 2-pixel reach of the 5x5 kernel
#define UP_BAND 4   // band width in high-res pixels (2 low-res pixels)

namespace {

struct blkv {       // device-side view of a blocked tensor
  __nv_bfloat16* p; long long bs; int Hp, Wp;
  __device__ __forceinline__ __nv_bfloat16* at(int b, int chunk, int y, int x) const {   // interior coordinates
    return p + (size_t)b * bs + (((size_t)chunk * Hp + (y + 2)) * Wp + (x + 2)) * 8;
  }
};
inline blkv view_of(const cnp_blk* v) {
  blkv r;
  r.Hp = v->H + 4; r.Wp = v->W + 4; r.bs = v->bstride;
  r.p = reinterpret_cast<__nv_bfloat16*>(v->base) + (size_t)v->cb_off * r.Hp * r.Wp * 8;
  return r;
}

__device__ __forceinline__ void up_src(int Y, int H, int* y0, int* y1, float* lam) {   // align_corners = False, scale 2
  float src = ((float)Y + 0.5f) * 0.5f - 0.5f;
  src = fmaxf(src, 0.f);
  const int i0 = (int)src;
  *y0 = i0; *y1 = min(i0 + 1, H - 1); *lam = src - (float)i0;
}
__device__ __forceinline__ void unpk(const uint4& r, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint4 pk(const float* v) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return r;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// bilinear x2 value of 8 channels at high-res (Y, X), rounded to bf16 exactly where blk_upsample2x_fwd rounds
__device__ __forceinline__ uint4 up_value(const blkv& x, int b, int chunk, int H, int W, int Y, int X) {
  int y0, y1, x0, x1; float ly, lx;
  up_src(Y, H, &y0, &y1, &ly);
  up_src(X, W, &x0, &x1, &lx);
  float a00[8], a01[8], a10[8], a11[8], o[8];
  unpk(ldg16(x.at(b, chunk, y0, x0)), a00); unpk(ldg16(x.at(b, chunk, y0, x1)), a01);
  unpk(ldg16(x.at(b, chunk, y1, x0)), a10); unpk(ldg16(x.at(b, chunk, y1, x1)), a11);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float r0 = (1.f - lx) * a00[k] + lx * a01[k], r1 = (1.f - lx) * a10[k] + lx * a11[k];
    o[k] = (1.f - ly) * r0 + ly * r1;
  }
  return pk(o);
}

// strips of the upsampled tensor: thread = one strip pixel (8 channels)
__global__ void __launch_bounds__(256)
up_strips_fwd_kernel(blkv x, int H, int W, int CB, int B, blkv rows, blkv cols) {
  const int L = 2 * W + 2 * H;                   // row-strip pixels then column-strip pixels of one strip row
  const long long total = (long long)B * CB * 2 * UP_SR * L;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int pos = (int)(e % L);
    long long r = e / L;
    const int t = (int)(r % UP_SR); r /= UP_SR;
    const int s = (int)(r & 1); r >>= 1;
    const int chunk = (int)(r % CB), b = (int)(r / CB);
    if (pos < 2 * W) {
      const int Y = s ? 2 * H - UP_SR + t : t;
      *reinterpret_cast<uint4*>(rows.at(s * B + b, chunk, t, pos)) = up_value(x, b, chunk, H, W, Y, pos);
    } else {
      const int Yc = pos - 2 * W, X = s ? 2 * W - UP_SR + t : t;
      *reinterpret_cast<uint4*>(cols.at(s * B + b, chunk, t, Yc)) = up_value(x, b, chunk, H, W, Yc, X);
    }
  }
}

// band outputs of the strip convolutions -> the high-res destination (8 chunks at the view's chunk offset)
__global__ void __launch_bounds__(256)
up_strips_scatter_kernel(blkv rows, blkv cols, blkv dst, int H2, int W2, int B) {
  // per image and chunk: 2 row bands of UP_BAND x W2, then 2 column bands of UP_BAND x (H2 - 2*UP_BAND) (corners: rows)
  const int nr = 2 * UP_BAND * W2, nc = 2 * UP_BAND * (H2 - 2 * UP_BAND);
  const long long total = (long long)B * 8 * (nr + nc);
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    int i = (int)(e % (nr + nc));
    const int chunk = (int)((e / (nr + nc)) & 7), b = (int)(e / (nr + nc) / 8);
    if (i < nr) {
      const int X = i % W2, k = i / W2, s = k / UP_BAND, u = k % UP_BAND;
      const int t = s ? UP_SR - UP_BAND + u : u, Y = s ? H2 - UP_BAND + u : u;
      *reinterpret_cast<uint4*>(dst.at(b, chunk, Y, X)) = *reinterpret_cast<const uint4*>(rows.at(s * B + b, chunk, t, X));
    } else {
      i -= nr;
      const int Hc = H2 - 2 * UP_BAND;
      const int Y = UP_BAND + i % Hc, k = i / Hc, s = k / UP_BAND, u = k % UP_BAND;
      const int t = s ? UP_SR - UP_BAND + u : u, X = s ? W2 - UP_BAND + u : u;
      *reinterpret_cast<uint4*>(dst.at(b, chunk, Y, X)) = *reinterpret_cast<const uint4*>(cols.at(s * B + b, chunk, t, Y));
    }
  }
}

// backward, partition of dy: (1) phase planes at low resolution with the band zeroed, (2) band strips (corner squares in
// the row strips only).  Thread = one low-res pixel x chunk for (1), one strip pixel for (2).
__global__ void __launch_bounds__(256)
up_dy_split_kernel(blkv dy, int H, int W, int B, blkv s2d, blkv rows, blkv cols) {
  const int H2 = 2 * H, W2 = 2 * W;
  const long long n1 = s2d.p ? (long long)B * 8 * H * W : 0;
  const int L = W2 + H2;
  const long long n2 = (long long)B * 8 * 2 * UP_SR * L;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n1 + n2; e += (long long)gridDim.x * 256) {
    if (e < n1) {
      const int j = (int)(e % W), i = (int)((e / W) % H);
      const int chunk = (int)((e / ((long long)W * H)) & 7), b = (int)(e / ((long long)W * H) / 8);
      const bool inner = i >= 2 && i < H - 2 && j >= 2 && j < W - 2;
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (inner) v = ldg16(dy.at(b, chunk, 2 * i + (ph >> 1), 2 * j + (ph & 1)));
        *reinterpret_cast<uint4*>(s2d.at(b, ph * 8 + chunk, i, j)) = v;
      }
    } else {
      long long r = e - n1;
      const int pos = (int)(r % L); r /= L;
      const int t = (int)(r % UP_SR); r /= UP_SR;
      const int s = (int)(r & 1); r >>= 1;
      const int chunk = (int)(r & 7), b = (int)(r >> 3);
      const bool band = s ? t >= UP_SR - UP_BAND : t < UP_BAND;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (pos < W2) {
        if (band) v = ldg16(dy.at(b, chunk, s ? H2 - UP_SR + t : t, pos));
        *reinterpret_cast<uint4*>(rows.at(s * B + b, chunk, t, pos)) = v;
      } else {
        const int Y = pos - W2;
        if (band && Y >= UP_BAND && Y < H2 - UP_BAND) v = ldg16(dy.at(b, chunk, Y, s ? W2 - UP_SR + t : t));
        *reinterpret_cast<uint4*>(cols.at(s * B + b, chunk, t, Y)) = v;
      }
    }
  }
}

// weights with which high-res index Y feeds low-res index r (n low-res pixels): transposed bilinear x2
__device__ __forceinline__ float up_wgt_t(int Y, int r, int n) {
  const int d = Y - 2 * r;        // -1, 0, 1, 2
  float w = (d == 0 || d == 1) ? 0.75f : 0.25f;
  if ((r == 0 && Y == 0) || (r == n - 1 && Y == 2 * n - 1)) w = 1.f;
  return w;
}

// input gradient of the band: the strips' gradients w.r.t. the upsampled tensor, gathered through the transposed
// bilinear map onto the low-res pixels within 4 of the border, ReLU-masked by act > 0 and ADDED to dx.
// The map is separable, so a thread owns a LINE of four outputs across the band and sweeps its strip once:
//   part A (row bands, columns 4 .. W-5): thread = (edge s, column c): the 24 strip pixels around c in flight at once,
//           combined horizontally per strip row, then with the vertical weights of its 4 output rows;
//   part B (column bands, rows 4 .. H-5): thread = (edge s, row r), the same on the transposed strips;
//   part C (corner squares): one output per thread, gathered from the row strip AND the column strip of its corner.
// Needs H, W >= 12 (the two strips of a dimension must not overlap).
__device__ __forceinline__ void fold_store(const blkv& dx, const blkv& act, int b, int chunk, int r, int c, const float* acc) {
  float old[8], av[8], v[8];
  __nv_bfloat16* d = dx.at(b, chunk, r, c);
  unpk(*reinterpret_cast<const uint4*>(d), old);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = acc[k];
  if (act.p) {
    unpk(ldg16(act.at(b, chunk, r, c)), av);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = av[k] > 0.f ? v[k] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) old[k] += v[k];
  *reinterpret_cast<uint4*>(d) = pk(old);
}

// the four high-res neighbours 2u-1 .. 2u+2 of low-res index u: clamped indices and weights (0 outside the image)
__device__ __forceinline__ void fold_taps(int u, int n, int* idx, float* wgt) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int X = 2 * u - 1 + j;
    const bool ok = X >= 0 && X < 2 * n;
    idx[j] = ok ? X : 0;
    wgt[j] = ok ? up_wgt_t(X, u, n) : 0.f;
  }
}

__global__ void __launch_bounds__(128)
up_strips_bwd_fold_kernel(blkv rows, blkv cols, blkv dx, blkv act, int H, int W, int CB, int B) {
  // items per (image, chunk): part A = 2 (W - 8) row-band lines, part B = 2 (H - 8) column-band lines, part C = the 64
  // outputs of the four corner squares (they see BOTH strips: one output per thread)
  const int nA = 2 * (W - 8), nB = 2 * (H - 8), nC = 64;
  const long long total = (long long)B * CB * (nA + nB + nC);
  for (long long e = (long long)blockIdx.x * 128 + threadIdx.x; e < total; e += (long long)gridDim.x * 128) {
    const int i = (int)(e % (nA + nB + nC));
    const int chunk = (int)((e / (nA + nB + nC)) % CB), b = (int)(e / (nA + nB + nC) / CB);
    if (i >= nA + nB) {
      const int q = i - nA - nB, s = q >> 5, sc = (q >> 4) & 1, k = (q >> 2) & 3, l = q & 3;
      const int r = (s ? H - 4 : 0) + k, c = (sc ? W - 4 : 0) + l;
      int yi[4], xi[4]; float wy[4], wx[4];
      fold_taps(r, H, yi, wy);
      fold_taps(c, W, xi, wx);
      const int ty0 = s ? 2 * H - UP_SR : 0, tx0 = sc ? 2 * W - UP_SR : 0;
      float acc[8];
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) acc[c8] = 0.f;
#pragma unroll
      for (int jy = 0; jy < 4; ++jy)
#pragma unroll
        for (int jx = 0; jx < 4; ++jx) {
          const float wgt = wy[jy] * wx[jx];
          const int ty = yi[jy] - ty0, tx = xi[jx] - tx0;
          float v[8];
          if (wgt != 0.f && ty >= 0 && ty < UP_SR) {
            unpk(ldg16(rows.at(s * B + b, chunk, ty, xi[jx])), v);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) acc[c8] = fmaf(wgt, v[c8], acc[c8]);
          }
          if (wgt != 0.f && tx >= 0 && tx < UP_SR) {
            unpk(ldg16(cols.at(sc * B + b, chunk, tx, yi[jy])), v);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) acc[c8] = fmaf(wgt, v[c8], acc[c8]);
          }
        }
      fold_store(dx, act, b, chunk, r, c, acc);
      continue;
    }
    const bool partA = i < nA;
    // line coordinate u along the edge (extent n); the 4 outputs lie across the band at v0 .. v0+3 (extent m)
    const int s = partA ? i / (W - 8) : (i - nA) / (H - 8);
    const int u = 4 + (partA ? i % (W - 8) : (i - nA) % (H - 8));
    const int n = partA ? W : H, m = partA ? H : W;
    const blkv& strip = partA ? rows : cols;
    const int v0 = s ? m - 4 : 0;                 // first output across the band
    const int t0 = s ? 2 * m - UP_SR : 0;         // high-res index of strip row 0
    int xi[4]; float wx[4];
    fold_taps(u, n, xi, wx);
    uint4 raw[UP_SR][4];
#pragma unroll
    for (int t = 0; t < UP_SR; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) raw[t][j] = ldg16(strip.at(s * B + b, chunk, t, xi[j]));
    float acc[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) acc[k][c8] = 0.f;
#pragma unroll
    for (int t = 0; t < UP_SR; ++t) {
      float h[8];
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) h[c8] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[8];
        unpk(raw[t][j], v);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) h[c8] = fmaf(wx[j], v[c8], h[c8]);
      }
      const int Y = t0 + t;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int d = Y - 2 * (v0 + k);
        const float wy = (d >= -1 && d <= 2) ? up_wgt_t(Y, v0 + k, m) : 0.f;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) acc[k][c8] = fmaf(wy, h[c8], acc[k][c8]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (partA) fold_store(dx, act, b, chunk, v0 + k, u, acc[k]);
      else fold_store(dx, act, b, chunk, u, v0 + k, acc[k]);
    }
  }
}

// 1-D map from the 5 kernel taps to the 4 low-res taps of output phase a (same as fold_in.cu: up_fold)
__device__ __forceinline__ float fold1(int a, int k, int p) {
  const int s = a + k - 2;
  const int m = (s >= 0) ? s / 2 : -((-s + 1) / 2), r = s - 2 * m;
  const int p0 = m + (r == 0 ? -1 : 0) + 2 - a, p1 = p0 + 1;
  const float w0 = r == 0 ? 0.25f : 0.75f, w1 = r == 0 ? 0.75f : 0.25f;
  return (p == p0 ? w0 : 0.f) + (p == p1 ? w1 : 0.f);
}

// dw5[co][ci][k][l] += sum_{a,b,p,q} dwp[a][b][co][ci][p][q] fold(a,k,p) fold(b,l,q)  +  dwt[co][ci][l][k]
// (dwt: the column strips' weight gradient, taken with the taps transposed).  One thread per (co, ci): 64 + 25 loads.
__global__ void __launch_bounds__(128)
up_wgrad_fold_kernel(const float* __restrict__ dwp, const float* __restrict__ dwt, int Cout, int Cin,
                     float* __restrict__ dw5) {
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= Cout * Cin) return;
  float acc[25];
#pragma unroll
  for (int i = 0; i < 25; ++i) acc[i] = 0.f;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const float4* src = reinterpret_cast<const float4*>(dwp + ((size_t)(a * 2 + b) * Cout * Cin + e) * 16);
      float g[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float4 v = __ldg(src + i); g[4 * i] = v.x; g[4 * i + 1] = v.y; g[4 * i + 2] = v.z; g[4 * i + 3] = v.w; }
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float fk = fold1(a, k, p);
          if (fk != 0.f) {
#pragma unroll
            for (int l = 0; l < 5; ++l)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float fl = fold1(b, l, q);
                if (fl != 0.f) acc[k * 5 + l] = fmaf(g[p * 4 + q] * fk, fl, acc[k * 5 + l]);
              }
          }
        }
    }
  float* dst = dw5 + (size_t)e * 25;
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int l = 0; l < 5; ++l) dst[k * 5 + l] += acc[k * 5 + l] + (dwt ? __ldg(dwt + (size_t)e * 25 + l * 5 + k) : 0.f);
}

int grid_for(long long total) {
  long long g = (total + 255) / 256;
  return (int)(g < 148 * 16 ? (g < 1 ? 1 : g) : 148 * 16);
}

bool strips_ok(const cnp_blk* rows, const cnp_blk* cols, int H, int W) {
  return rows && cols && rows->H == UP_SR && rows->W == 2 * W && cols->H == UP_SR && cols->W == 2 * H;
}

}  // namespace

// rows / cols <- strips of bilinear_up2x(x) (n_chunks chunks of x; strip images s*B + b).
CNP_API int cnp_up_strips_fwd(const cnp_blk* x, int n_chunks, const cnp_blk* rows, const cnp_blk* cols, int B,
                              cudaStream_t st) {
  CNP_REQUIRE(x && n_chunks > 0 && B > 0 && x->H >= 8 && x->W >= 8 && strips_ok(rows, cols, x->H, x->W),
              "up_strips_fwd: needs an input of at least 8 x 8 and strips of 6 x 2W / 6 x 2H");
  const long long total = (long long)B * n_chunks * 2 * UP_SR * (2 * x->W + 2 * x->H);
  up_strips_fwd_kernel<<<grid_for(total), 256, 0, st>>>(view_of(x), x->H, x->W, n_chunks, B, view_of(rows), view_of(cols));
  CNP_LAUNCH_CHECK("up_strips_fwd_kernel");
  return 0;
}

// dst (8 chunks, high resolution 2H x 2W) <- the band rows / columns of the strip convolutions' outputs.
CNP_API int cnp_up_strips_scatter(const cnp_blk* rows, const cnp_blk* cols, const cnp_blk* dst, int B, cudaStream_t st) {
  CNP_REQUIRE(dst && B > 0 && dst->H >= 16 && dst->W >= 16 && strips_ok(rows, cols, dst->H / 2, dst->W / 2),
              "up_strips_scatter: geometry mismatch");
  const long long total = (long long)B * 8 * (2 * UP_BAND * dst->W + 2 * UP_BAND * (dst->H - 2 * UP_BAND));
  up_strips_scatter_kernel<<<grid_for(total), 256, 0, st>>>(view_of(rows), view_of(cols), view_of(dst), dst->H, dst->W, B);
  CNP_LAUNCH_CHECK("up_strips_scatter_kernel");
  return 0;
}

// dy (8 chunks, 2H x 2W) -> s2d (32 chunks = phase (a*2+b)*8 + chunk at H x W, band zeroed) + band strips.
CNP_API int cnp_up_dy_split(const cnp_blk* dy, const cnp_blk* s2d, const cnp_blk* rows, const cnp_blk* cols, int B,
                            cudaStream_t st) {
  CNP_REQUIRE(dy && B > 0 && dy->H % 2 == 0 && dy->W % 2 == 0 && dy->H >= 16 && dy->W >= 16 &&
              (!s2d || (dy->H == 2 * s2d->H && dy->W == 2 * s2d->W && s2d->cb_off == 0)) &&
              strips_ok(rows, cols, dy->H / 2, dy->W / 2), "up_dy_split: geometry mismatch");
  const int H = dy->H / 2, W = dy->W / 2;
  blkv sv;
  sv.p = nullptr; sv.bs = 0; sv.Hp = sv.Wp = 0;
  if (s2d) sv = view_of(s2d);
  const long long total = (s2d ? (long long)B * 8 * H * W : 0) + (long long)B * 8 * 2 * UP_SR * (2 * W + 2 * H);
  up_dy_split_kernel<<<grid_for(total), 256, 0, st>>>(view_of(dy), H, W, B, sv, view_of(rows), view_of(cols));
  CNP_LAUNCH_CHECK("up_dy_split_kernel");
  return 0;
}

// dx (n_chunks chunks, H x W) += mask(act > 0) * bilinear_up2x^T(strip gradients); act may be NULL.
CNP_API int cnp_up_strips_bwd_fold(const cnp_blk* rows, const cnp_blk* cols, const cnp_blk* dx, const cnp_blk* act,
                                   int n_chunks, int B, cudaStream_t st) {
  CNP_REQUIRE(dx && n_chunks > 0 && B > 0 && dx->H >= 12 && dx->W >= 12 && strips_ok(rows, cols, dx->H, dx->W) &&
              (!act || (act->H == dx->H && act->W == dx->W)), "up_strips_bwd_fold: needs at least 12 x 12 low-res pixels");
  blkv a;
  a.p = nullptr; a.bs = 0; a.Hp = a.Wp = 0;
  if (act) a = view_of(act);
  const long long total = (long long)B * n_chunks * (2 * (dx->W - 8) + 2 * (dx->H - 8) + 64);
  up_strips_bwd_fold_kernel<<<(int)((total + 127) / 128), 128, 0, st>>>(view_of(rows), view_of(cols), view_of(dx), a, dx->H, dx->W,
                                                           n_chunks, B);
  CNP_LAUNCH_CHECK("up_strips_bwd_fold_kernel");
  return 0;
}

// dw5 [Cout][Cin][5][5] += fold^T(dwp [2][2][Cout][Cin][4][4]) + transpose_taps(dwt [Cout][Cin][5][5]) (dwt may be NULL).
CNP_API int cnp_up_wgrad_fold(const float* dwp, const float* dwt, int Cout, int Cin, float* dw5, cudaStream_t st) {
  CNP_REQUIRE(dwp && dw5 && Cout > 0 && Cin > 0, "up_wgrad_fold: bad arguments");
  up_wgrad_fold_kernel<<<cnp_cdiv(Cout * Cin, 128), 128, 0, st>>>(dwp, dwt, Cout, Cin, dw5);
  CNP_LAUNCH_CHECK("up_wgrad_fold_kernel");
  return 0;
}
