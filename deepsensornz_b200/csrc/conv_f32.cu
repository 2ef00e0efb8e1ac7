// fp32 UNet building blocks (the 1e-5 parity path): direct convolution forward / dgrad / wgrad on
// NCHW tensors, ReLU backward, bilinear x2 resize forward / backward, per-channel sums.
//
// Replaces the torch Conv2d / Upsample / ReLU autograd ops inside upstream neuralprocesses' UNet
//  as reached from ConvNP.loss_fn (nzdownscale/downscaler/train.py:370,:388-394).
// True fp32 FFMA on the CUDA cores: tensor cores would need bf16/TF32 and break the 1e-5 bound; the
// bf16 tcgen05 path lives in conv_bf16.cu.
#include "common.cuh"

namespace {

constexpr int CO_T = 16;   // output channels per block (forward / dgrad)
constexpr int TH = 8;      // tile rows
constexpr int TWX = 64;    // tile cols (two pixels per thread: tx and tx+32)

// ---------------------------------------------------------------------------------------------
// forward direct conv.  mode 0: w[co][ci][tap].  mode 1 (stride-1 dgrad): the "input" is dy with
// Cin_eff = Cout_orig channels, the output has Cout_eff = Cin_orig channels and the effective
// weight is w[ci_eff][co_eff][KK-1-tap] (transposed + flipped).
// ---------------------------------------------------------------------------------------------
template <int K, int S, int CI_T>
__global__ void __launch_bounds__(256)
conv_fwd_kernel(const float* __restrict__ x, long long x_bs, int Cin, int Hin, int Win,
                const float* __restrict__ w, const float* __restrict__ bias, int wmode,
                float* __restrict__ y, long long y_bs, int Cout, int Hout, int Wout,
                int relu, int accumulate, int tiles_x) {
  constexpr int P = K / 2;
  constexpr int IH = (TH - 1) * S + K;
  constexpr int IW = (TWX - 1) * S + K;
  constexpr int KK = K * K;
  __shared__ float in_s[CI_T][IH][IW];
  __shared__ __align__(16) float w_s[CI_T][KK][CO_T];

  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tile = blockIdx.x, tyi = tile / tiles_x, txi = tile % tiles_x;
  const int oy0 = tyi * TH, ox0 = txi * TWX;
  const int co0 = blockIdx.y * CO_T, b = blockIdx.z;
  const int iy0 = oy0 * S - P, ix0 = ox0 * S - P;
  const float* xb = x + (size_t)b * x_bs;

  float acc[2][CO_T];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc[p][c] = 0.f;

  for (int ci0 = 0; ci0 < Cin; ci0 += CI_T) {
    __syncthreads();
    for (int e = threadIdx.x; e < CI_T * IH * IW; e += 256) {
      int ci = e / (IH * IW), r = (e / IW) % IH, c = e % IW;
      int iy = iy0 + r, ix = ix0 + c;
      float v = 0.f;
      if (ci0 + ci < Cin && iy >= 0 && iy < Hin && ix >= 0 && ix < Win)
        v = __ldg(xb + ((size_t)(ci0 + ci) * Hin + iy) * Win + ix);
      in_s[ci][r][c] = v;
    }
    for (int e = threadIdx.x; e < CI_T * KK * CO_T; e += 256) {
      int ci = e / (KK * CO_T), t = (e / CO_T) % KK, co = e % CO_T;
      float v = 0.f;
      if (ci0 + ci < Cin && co0 + co < Cout) {
        if (wmode == 0) v = __ldg(w + ((size_t)(co0 + co) * Cin + (ci0 + ci)) * KK + t);
        else v = __ldg(w + ((size_t)(ci0 + ci) * Cout + (co0 + co)) * KK + (KK - 1 - t));
      }
      w_s[ci][t][co] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < CI_T; ++ci) {
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float v0 = in_s[ci][ty * S + ky][tx * S + kx];
          const float v1 = in_s[ci][ty * S + ky][(tx + 32) * S + kx];
          const float4* wp = reinterpret_cast<const float4*>(&w_s[ci][ky * K + kx][0]);
#pragma unroll
          for (int q = 0; q < CO_T / 4; ++q) {
            const float4 wv = wp[q];
            acc[0][4 * q + 0] = fmaf(v0, wv.x, acc[0][4 * q + 0]);
            acc[0][4 * q + 1] = fmaf(v0, wv.y, acc[0][4 * q + 1]);
            acc[0][4 * q + 2] = fmaf(v0, wv.z, acc[0][4 * q + 2]);
            acc[0][4 * q + 3] = fmaf(v0, wv.w, acc[0][4 * q + 3]);
            acc[1][4 * q + 0] = fmaf(v1, wv.x, acc[1][4 * q + 0]);
            acc[1][4 * q + 1] = fmaf(v1, wv.y, acc[1][4 * q + 1]);
            acc[1][4 * q + 2] = fmaf(v1, wv.z, acc[1][4 * q + 2]);
            acc[1][4 * q + 3] = fmaf(v1, wv.w, acc[1][4 * q + 3]);
          }
        }
      }
    }
  }
  const int oy = oy0 + ty;
  if (oy >= Hout) return;
  float* yb = y + (size_t)b * y_bs;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int ox = ox0 + tx + 32 * p;
    if (ox >= Wout) continue;
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
      if (co0 + c < Cout) {
        float v = acc[p][c] + (bias ? __ldg(bias + co0 + c) : 0.f);
        if (relu && v < 0.f) v = 0.f;  // NaN propagates like torch.relu
        float* dst = yb + ((size_t)(co0 + c) * Hout + oy) * Wout + ox;
        if (accumulate) v += *dst;
        *dst = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// stride-2 dgrad (transposed conv, k=5, p=2), gather form, one parity class per block.
//   dx[ci, 2u+py, 2v+px] (+)= sum_co sum_{ky = py (mod 2)} sum_{kx = px (mod 2)}
//        dy[co, u + (py+2-ky)/2, v + (px+2-kx)/2] * w[co][ci][ky][kx]
// ---------------------------------------------------------------------------------------------
constexpr int DG_CO_T = 8;

__global__ void __launch_bounds__(256)
conv_dgrad_s2_kernel(const float* __restrict__ dy, long long dy_bs, int Cout, int Hout, int Wout,
                     const float* __restrict__ w, float* __restrict__ dx, long long dx_bs,
                     int Cin, int Hin, int Win, int accumulate, int tiles_x) {
  constexpr int K = 5, KK = 25;
  __shared__ __align__(16) float w_s[DG_CO_T][KK][CO_T];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int phase = blockIdx.x & 3, tile = blockIdx.x >> 2;
  const int py = phase >> 1, px = phase & 1;
  const int u = (tile / tiles_x) * TH + ty, v = (tile % tiles_x) * 32 + tx;
  const int ci0 = blockIdx.y * CO_T, b = blockIdx.z;
  const int yi = 2 * u + py, xi = 2 * v + px;
  const float* dyb = dy + (size_t)b * dy_bs;

  float acc[CO_T];
#pragma unroll
  for (int c = 0; c < CO_T; ++c) acc[c] = 0.f;

  for (int co0 = 0; co0 < Cout; co0 += DG_CO_T) {
    __syncthreads();
    for (int e = threadIdx.x; e < DG_CO_T * KK * CO_T; e += 256) {
      int co = e / (KK * CO_T), t = (e / CO_T) % KK, ci = e % CO_T;
      float val = 0.f;
      if (co0 + co < Cout && ci0 + ci < Cin) val = __ldg(w + ((size_t)(co0 + co) * Cin + (ci0 + ci)) * KK + t);
      w_s[co][t][ci] = val;
    }
    __syncthreads();
    for (int co = 0; co < DG_CO_T && co0 + co < Cout; ++co) {
      const float* dyc = dyb + (size_t)(co0 + co) * Hout * Wout;
      for (int ky = py; ky < K; ky += 2) {
        const int yo = u + (py + 2 - ky) / 2;
        if (yo < 0 || yo >= Hout) continue;
        for (int kx = px; kx < K; kx += 2) {
          const int xo = v + (px + 2 - kx) / 2;
          if (xo < 0 || xo >= Wout) continue;
          const float g = __ldg(dyc + (size_t)yo * Wout + xo);
          const float4* wp = reinterpret_cast<const float4*>(&w_s[co][ky * K + kx][0]);
#pragma unroll
          for (int q = 0; q < CO_T / 4; ++q) {
            const float4 wv = wp[q];
            acc[4 * q + 0] = fmaf(g, wv.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(g, wv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(g, wv.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(g, wv.w, acc[4 * q + 3]);
          }
        }
      }
    }
  }
  if (yi >= Hin || xi >= Win) return;
  float* dxb = dx + (size_t)b * dx_bs;
#pragma unroll
  for (int c = 0; c < CO_T; ++c) {
    if (ci0 + c < Cin) {
      float* dst = dxb + ((size_t)(ci0 + c) * Hin + yi) * Win + xi;
      *dst = accumulate ? (*dst + acc[c]) : acc[c];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: dw[co][ci][tap] += sum_{b,y,x} dy[b,co,y,x] * x[b,ci,S*y+ky-P,S*x+kx-P]
// Block = (pixel-tile group, slice of 8 input channels); all Cout (<=64 per pass) x 8ci x KK outputs
// are register-tiled: thread = (group of 8 co) x up to NPAIR (ci,tap) pairs.
// ---------------------------------------------------------------------------------------------
constexpr int WG_CI = 8;
constexpr int WG_TH = 4, WG_TW = 32;  // 128 pixels per tile
constexpr int WG_CO = 64;             // output channels per pass
constexpr int WG_LD = WG_CO + 4;      // padded row stride of the staged dy tile (keeps float4 alignment)

template <int K, int S>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float* __restrict__ x, long long x_bs, int Cin, int Hin, int Win,
                  const float* __restrict__ dy, long long dy_bs, int Cout, int Hout, int Wout,
                  float* __restrict__ dw, int B, int tiles_x, int tiles_y) {
  constexpr int P = K / 2, KK = K * K;
  constexpr int IH = (WG_TH - 1) * S + K, IW = (WG_TW - 1) * S + K;
  constexpr int NPAIRS = WG_CI * KK;
  constexpr int NP = (NPAIRS + 31) / 32;
  extern __shared__ __align__(16) float smem[];
  float* x_s = smem;                                   // [WG_CI][IH][IW]
  float* dy_s = smem + WG_CI * IH * IW;                // [128 px][WG_CO]

  const int lane = threadIdx.x & 31, cg = threadIdx.x >> 5;  // cg: group of 8 output channels
  const int ci0 = blockIdx.y * WG_CI;
  int off[NP];
  bool pv[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    int pr = lane + 32 * k;
    pv[k] = pr < NPAIRS;
    int ci = pv[k] ? pr / KK : 0, t = pv[k] ? pr % KK : 0;
    off[k] = ci * IH * IW + (t / K) * IW + (t % K);
  }
  const int ntiles = B * tiles_x * tiles_y;
  for (int cob = 0; cob < Cout; cob += WG_CO) {
    float acc[NP][8];
#pragma unroll
    for (int k = 0; k < NP; ++k)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[k][c] = 0.f;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int b = tile / (tiles_x * tiles_y), tr = tile % (tiles_x * tiles_y);
      const int oy0 = (tr / tiles_x) * WG_TH, ox0 = (tr % tiles_x) * WG_TW;
      const int iy0 = oy0 * S - P, ix0 = ox0 * S - P;
      const float* xb = x + (size_t)b * x_bs;
      const float* dyb = dy + (size_t)b * dy_bs;
      __syncthreads();
      for (int e = threadIdx.x; e < WG_CI * IH * IW; e += 256) {
        int ci = e / (IH * IW), r = (e / IW) % IH, c = e % IW;
        int iy = iy0 + r, ix = ix0 + c;
        float v = 0.f;
        if (ci0 + ci < Cin && iy >= 0 && iy < Hin && ix >= 0 && ix < Win)
          v = __ldg(xb + ((size_t)(ci0 + ci) * Hin + iy) * Win + ix);
        x_s[e] = v;
      }
      for (int e = threadIdx.x; e < WG_TH * WG_TW * WG_CO; e += 256) {
        int co = e / (WG_TH * WG_TW), px = e % (WG_TH * WG_TW);
        int oy = oy0 + px / WG_TW, ox = ox0 + px % WG_TW;
        float v = 0.f;
        if (cob + co < Cout && oy < Hout && ox < Wout)
          v = __ldg(dyb + ((size_t)(cob + co) * Hout + oy) * Wout + ox);
        dy_s[px * WG_LD + co] = v;
      }
      __syncthreads();
#pragma unroll 2
      for (int px = 0; px < WG_TH * WG_TW; ++px) {
        const int base = (px / WG_TW) * S * IW + (px % WG_TW) * S;
        const float4 g0 = *reinterpret_cast<const float4*>(dy_s + px * WG_LD + cg * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(dy_s + px * WG_LD + cg * 8 + 4);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          const float xv = pv[k] ? x_s[base + off[k]] : 0.f;
          acc[k][0] = fmaf(xv, g0.x, acc[k][0]); acc[k][1] = fmaf(xv, g0.y, acc[k][1]);
          acc[k][2] = fmaf(xv, g0.z, acc[k][2]); acc[k][3] = fmaf(xv, g0.w, acc[k][3]);
          acc[k][4] = fmaf(xv, g1.x, acc[k][4]); acc[k][5] = fmaf(xv, g1.y, acc[k][5]);
          acc[k][6] = fmaf(xv, g1.z, acc[k][6]); acc[k][7] = fmaf(xv, g1.w, acc[k][7]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      if (!pv[k]) continue;
      int pr = lane + 32 * k, ci = pr / KK, t = pr % KK;
      if (ci0 + ci >= Cin) continue;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        int co = cob + cg * 8 + c;
        if (co < Cout) atomicAdd(dw + ((size_t)co * Cin + ci0 + ci) * KK + t, acc[k][c]);
      }
    }
  }
}

// per-channel sum over batch and pixels: out[c] += sum_{b,p} v[b,c,p]
__global__ void __launch_bounds__(256)
channel_sum_kernel(const float* __restrict__ v, long long bs, int C, int HW, int B, float* __restrict__ out) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  float s = 0.f;
  const long long total = (long long)B * HW;
  for (long long e = (long long)blockIdx.y * 256 + threadIdx.x; e < total; e += (long long)gridDim.y * 256) {
    int b = (int)(e / HW), p = (int)(e % HW);
    s += __ldg(v + (size_t)b * bs + (size_t)c * HW + p);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(out + c, t);
  }
}

__global__ void __launch_bounds__(256)
relu_bwd_kernel(float* __restrict__ dy, long long dy_bs, const float* __restrict__ y, long long y_bs,
                long long per_batch) {
  const int b = blockIdx.y;
  float* d = dy + (size_t)b * dy_bs;
  const float* a = y + (size_t)b * y_bs;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < per_batch; e += (long long)gridDim.x * 256)
    if (!(a[e] > 0.f)) d[e] = 0.f;
}

// ---- bilinear x2, align_corners=False (torch nn.Upsample(scale_factor=2, mode='bilinear')) ----
__device__ __forceinline__ void up2_src(int Y, int H, int* y0, int* y1, float* lam) {
  float src = ((float)Y + 0.5f) * 0.5f - 0.5f;
  src = fmaxf(src, 0.f);
  int i0 = (int)src;
  *y0 = i0;
  *y1 = min(i0 + 1, H - 1);
  *lam = src - (float)i0;
}

__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const float* __restrict__ x, long long x_bs, float* __restrict__ y, long long y_bs,
                      int C, int H, int W) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int H2 = 2 * H, W2 = 2 * W;
  const float* xc = x + (size_t)b * x_bs + (size_t)c * H * W;
  float* yc = y + (size_t)b * y_bs + (size_t)c * H2 * W2;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < H2 * W2; e += gridDim.x * 256) {
    int Y = e / W2, X = e % W2, y0, y1, x0, x1;
    float ly, lx;
    up2_src(Y, H, &y0, &y1, &ly);
    up2_src(X, W, &x0, &x1, &lx);
    float r0 = (1.f - lx) * xc[y0 * W + x0] + lx * xc[y0 * W + x1];
    float r1 = (1.f - lx) * xc[y1 * W + x0] + lx * xc[y1 * W + x1];
    yc[e] = (1.f - ly) * r0 + ly * r1;
  }
}

__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const float* __restrict__ dy, long long dy_bs, float* __restrict__ dx, long long dx_bs,
                      int C, int H, int W, int accumulate) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int H2 = 2 * H, W2 = 2 * W;
  const float* dyc = dy + (size_t)b * dy_bs + (size_t)c * H2 * W2;
  float* dxc = dx + (size_t)b * dx_bs + (size_t)c * H * W;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < H * W; e += gridDim.x * 256) {
    const int u = e / W, v = e % W;
    float s = 0.f;
    for (int Y = max(2 * u - 1, 0); Y <= min(2 * u + 2, H2 - 1); ++Y) {
      int y0, y1; float ly;
      up2_src(Y, H, &y0, &y1, &ly);
      float wy = (y0 == u ? (1.f - ly) : 0.f) + (y1 == u ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int X = max(2 * v - 1, 0); X <= min(2 * v + 2, W2 - 1); ++X) {
        int x0, x1; float lx;
        up2_src(X, W, &x0, &x1, &lx);
        float wx = (x0 == v ? (1.f - lx) : 0.f) + (x1 == v ? lx : 0.f);
        if (wx != 0.f) s = fmaf(wy * wx, dyc[Y * W2 + X], s);
      }
    }
    dxc[e] = accumulate ? dxc[e] + s : s;
  }
}

template <int K, int S, int CI_T>
int launch_fwd(const float* x, long long x_bs, int B, int Cin, int Hin, int Win, const float* w, const float* bias,
               int wmode, float* y, long long y_bs, int Cout, int relu, int accumulate, cudaStream_t st) {
  const int P = K / 2;
  const int Hout = (Hin + 2 * P - K) / S + 1, Wout = (Win + 2 * P - K) / S + 1;
  const int tiles_x = cnp_cdiv(Wout, TWX), tiles_y = cnp_cdiv(Hout, TH);
  dim3 grid(tiles_x * tiles_y, cnp_cdiv(Cout, CO_T), B);
  conv_fwd_kernel<K, S, CI_T><<<grid, 256, 0, st>>>(x, x_bs, Cin, Hin, Win, w, bias, wmode, y, y_bs, Cout, Hout,
                                                     Wout, relu, accumulate, tiles_x);
  CNP_LAUNCH_CHECK("conv_fwd_kernel");
  return 0;
}

template <int K, int S>
int launch_wgrad(const float* x, long long x_bs, int B, int Cin, int Hin, int Win, const float* dy, long long dy_bs,
                 int Cout, float* dw, cudaStream_t st) {
  const int P = K / 2;
  const int Hout = (Hin + 2 * P - K) / S + 1, Wout = (Win + 2 * P - K) / S + 1;
  const int IH = (WG_TH - 1) * S + K, IW = (WG_TW - 1) * S + K;
  const int tiles_x = cnp_cdiv(Wout, WG_TW), tiles_y = cnp_cdiv(Hout, WG_TH);
  const size_t smem = (size_t)(WG_CI * IH * IW + WG_TH * WG_TW * WG_LD) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv_wgrad_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  int ntiles = B * tiles_x * tiles_y;
  int gx = ntiles < 96 ? ntiles : 96;
  dim3 grid(gx, cnp_cdiv(Cin, WG_CI));
  conv_wgrad_kernel<K, S><<<grid, 256, smem, st>>>(x, x_bs, Cin, Hin, Win, dy, dy_bs, Cout, Hout, Wout, dw, B,
                                                    tiles_x, tiles_y);
  CNP_LAUNCH_CHECK("conv_wgrad_kernel");
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
CNP_API int cnp_conv2d_fwd_f32(const float* x, long long x_bstride, const float* w, const float* bias, float* y,
                               long long y_bstride, int B, int Cin, int Hin, int Win, int Cout, int k, int stride,
                               int relu, cudaStream_t st) {
  CNP_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Hin > 0 && Win > 0, "conv2d_fwd_f32: bad sizes");
  if (k == 5 && stride == 1) return launch_fwd<5, 1, 8>(x, x_bstride, B, Cin, Hin, Win, w, bias, 0, y, y_bstride, Cout, relu, 0, st);
  if (k == 5 && stride == 2) return launch_fwd<5, 2, 4>(x, x_bstride, B, Cin, Hin, Win, w, bias, 0, y, y_bstride, Cout, relu, 0, st);
  if (k == 1 && stride == 1) return launch_fwd<1, 1, 8>(x, x_bstride, B, Cin, Hin, Win, w, bias, 0, y, y_bstride, Cout, relu, 0, st);
  CNP_REQUIRE(false, "conv2d_fwd_f32: unsupported k=%d stride=%d (supported: k5 s1|s2, k1 s1)", k, stride);
}

// dx gets Cin channels on the Hin x Win input grid; dy has Cout channels on the output grid.
CNP_API int cnp_conv2d_dgrad_f32(const float* dy, long long dy_bstride, const float* w, float* dx,
                                 long long dx_bstride, int B, int Cin, int Hin, int Win, int Cout, int k,
                                 int stride, int accumulate, cudaStream_t st) {
  CNP_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Hin > 0 && Win > 0, "conv2d_dgrad_f32: bad sizes");
  if (stride == 1 && k == 5)
    return launch_fwd<5, 1, 8>(dy, dy_bstride, B, Cout, Hin, Win, w, nullptr, 1, dx, dx_bstride, Cin, 0, accumulate, st);
  if (stride == 1 && k == 1)
    return launch_fwd<1, 1, 8>(dy, dy_bstride, B, Cout, Hin, Win, w, nullptr, 1, dx, dx_bstride, Cin, 0, accumulate, st);
  if (stride == 2 && k == 5) {
    CNP_REQUIRE(Hin % 2 == 0 && Win % 2 == 0, "conv2d_dgrad_f32: stride-2 needs even input size");
    const int Hout = Hin / 2, Wout = Win / 2;
    const int tiles_x = cnp_cdiv(Wout, 32), tiles_y = cnp_cdiv(Hout, TH);
    dim3 grid(tiles_x * tiles_y * 4, cnp_cdiv(Cin, CO_T), B);
    conv_dgrad_s2_kernel<<<grid, 256, 0, st>>>(dy, dy_bstride, Cout, Hout, Wout, w, dx, dx_bstride, Cin, Hin, Win,
                                               accumulate, tiles_x);
    CNP_LAUNCH_CHECK("conv_dgrad_s2_kernel");
    return 0;
  }
  CNP_REQUIRE(false, "conv2d_dgrad_f32: unsupported k=%d stride=%d", k, stride);
}

// dw (+=) [Cout,Cin,k,k]; dbias (+=) [Cout] or NULL.
CNP_API int cnp_conv2d_wgrad_f32(const float* x, long long x_bstride, const float* dy, long long dy_bstride,
                                 float* dw, float* dbias, int B, int Cin, int Hin, int Win, int Cout, int k,
                                 int stride, cudaStream_t st) {
  CNP_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Hin > 0 && Win > 0, "conv2d_wgrad_f32: bad sizes");
  const int P = k / 2;
  const int Hout = (Hin + 2 * P - k) / stride + 1, Wout = (Win + 2 * P - k) / stride + 1;
  if (dbias) {
    dim3 grid(Cout, 16);
    channel_sum_kernel<<<grid, 256, 0, st>>>(dy, dy_bstride, Cout, Hout * Wout, B, dbias);
    CNP_LAUNCH_CHECK("channel_sum_kernel");
  }
  if (k == 5 && stride == 1) return launch_wgrad<5, 1>(x, x_bstride, B, Cin, Hin, Win, dy, dy_bstride, Cout, dw, st);
  if (k == 5 && stride == 2) return launch_wgrad<5, 2>(x, x_bstride, B, Cin, Hin, Win, dy, dy_bstride, Cout, dw, st);
  if (k == 1 && stride == 1) return launch_wgrad<1, 1>(x, x_bstride, B, Cin, Hin, Win, dy, dy_bstride, Cout, dw, st);
  CNP_REQUIRE(false, "conv2d_wgrad_f32: unsupported k=%d stride=%d", k, stride);
}

CNP_API int cnp_relu_bwd_f32(float* dy, long long dy_bstride, const float* y, long long y_bstride, int B,
                             long long per_batch, cudaStream_t st) {
  CNP_REQUIRE(B > 0 && per_batch > 0, "relu_bwd_f32: bad sizes");
  long long nb = (per_batch + 255) / 256;
  dim3 grid((unsigned)(nb < 1184 ? nb : 1184), B);
  relu_bwd_kernel<<<grid, 256, 0, st>>>(dy, dy_bstride, y, y_bstride, per_batch);
  CNP_LAUNCH_CHECK("relu_bwd_kernel");
  return 0;
}

CNP_API int cnp_upsample2x_fwd_f32(const float* x, long long x_bstride, float* y, long long y_bstride, int B, int C,
                                   int H, int W, cudaStream_t st) {
  CNP_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "upsample2x_fwd_f32: bad sizes");
  dim3 grid(min(cnp_cdiv(4 * H * W, 256), 64), C, B);
  upsample2x_fwd_kernel<<<grid, 256, 0, st>>>(x, x_bstride, y, y_bstride, C, H, W);
  CNP_LAUNCH_CHECK("upsample2x_fwd_kernel");
  return 0;
}

CNP_API int cnp_upsample2x_bwd_f32(const float* dy, long long dy_bstride, float* dx, long long dx_bstride, int B,
                                   int C, int H, int W, int accumulate, cudaStream_t st) {
  CNP_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "upsample2x_bwd_f32: bad sizes");
  dim3 grid(min(cnp_cdiv(H * W, 256), 64), C, B);
  upsample2x_bwd_kernel<<<grid, 256, 0, st>>>(dy, dy_bstride, dx, dx_bstride, C, H, W, accumulate);
  CNP_LAUNCH_CHECK("upsample2x_bwd_kernel");
  return 0;
}
