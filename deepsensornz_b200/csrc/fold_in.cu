// Folding the UNet's initial 1x1 convolution into the first 5x5 convolution (bf16 training / inference path).
//
// Upstream neuralprocesses' UNet starts with  h = initial_linear(x)  (1x1, Cin -> 64, no activation) followed by
// before_turn_layers[0] (5x5, 64 -> 64) + ReLU (coders/nn.py; reached from ConvNP.loss_fn,
// nzdownscale/downscaler/train.py:370 / train_epoch train.py:388-394).  Both are linear and h has no other reader,
// so on the tensor-core path the pair is ONE 5x5 convolution of the (Cin + 1)-channel encoder output:
//
//     y[co] = b5[co] + sum_{t,m} W5[co,m,t] * pad0( sum_c W1[m,c] x[c] + b1[m] )[t]
//           = b5[co] + sum_{t,c} Wf[co,c,t] * xa[c][t],      xa = [x ; 1] zero-padded,
//     Wf[co,c,t] = sum_m W5[co,m,t] W1[m,c]  (c < Cin),   Wf[co,Cin,t] = sum_m W5[co,m,t] b1[m].
//
// The constant-1 channel (1 inside the image, 0 in the pad) reproduces the border behaviour of the zero-padded
// bias exactly.  The 64-channel fp32->bf16 1x1 output (189 MB per 16-task step at 304^2), its 1x1 weight-gradient
// kernel and the dgrad of the first 5x5 disappear; the first layer's GEMM K shrinks from 1600 to 25 x 16.
// Backward: the tensor-core wgrad gives dWf; the chain rule back to the three real parameters is the tiny
// contraction below (64 x 64 x 25 x (Cin+1) MACs).
#include "common.cuh"

namespace {

// wf[co][c][t], c < Cp
__global__ void __launch_bounds__(256)
fold_in_fwd_kernel(const float* __restrict__ w5, const float* __restrict__ w1, const float* __restrict__ b1, int Cout,
                   int Cmid, int Cin, int Cp, int KK, float* __restrict__ wf) {
  const int total = Cout * Cp * KK;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
    const int t = e % KK, c = (e / KK) % Cp, co = e / (KK * Cp);
    float s = 0.f;
    if (c <= Cin) {
      for (int m = 0; m < Cmid; ++m) {
        const float a = c < Cin ? __ldg(w1 + (size_t)m * Cin + c) : __ldg(b1 + m);
        s = fmaf(__ldg(w5 + ((size_t)co * Cmid + m) * KK + t), a, s);
      }
    }
    wf[e] = s;
  }
}

// dw5[co][m][t] += sum_{c<=Cin} dwf[co][c][t] * A[m][c],   A = [W1 | b1]
__global__ void __launch_bounds__(256)
fold_in_bwd_w5_kernel(const float* __restrict__ dwf, const float* __restrict__ w1, const float* __restrict__ b1,
                      int Cout, int Cmid, int Cin, int Cp, int KK, float* __restrict__ dw5) {
  const int total = Cout * Cmid * KK;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
    const int t = e % KK, m = (e / KK) % Cmid, co = e / (KK * Cmid);
    float s = 0.f;
    for (int c = 0; c < Cin; ++c) s = fmaf(__ldg(dwf + ((size_t)co * Cp + c) * KK + t), __ldg(w1 + (size_t)m * Cin + c), s);
    s = fmaf(__ldg(dwf + ((size_t)co * Cp + Cin) * KK + t), __ldg(b1 + m), s);
    dw5[e] += s;
  }
}

// block (m, c): dA[m][c] += sum_{co,t} w5[co][m][t] * dwf[co][c][t];  c == Cin is the bias gradient
__global__ void __launch_bounds__(128)
fold_in_bwd_w1_kernel(const float* __restrict__ dwf, const float* __restrict__ w5, int Cout, int Cmid, int Cin, int Cp,
                      int KK, float* __restrict__ dw1, float* __restrict__ db1) {
  __shared__ float red[4];
  const int m = blockIdx.x, c = blockIdx.y;
  float s = 0.f;
  for (int e = threadIdx.x; e < Cout * KK; e += 128) {
    const int co = e / KK, t = e - co * KK;
    s = fmaf(__ldg(w5 + ((size_t)co * Cmid + m) * KK + t), __ldg(dwf + ((size_t)co * Cp + c) * KK + t), s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float tot = red[0] + red[1] + red[2] + red[3];
    if (c < Cin) dw1[(size_t)m * Cin + c] += tot;
    else if (db1) db1[m] += tot;
  }
}

// 1-D map from the 5 kernel taps to the 4 low-resolution taps of output phase a of a x2 bilinear upsampling
// (align_corners = False, interior formula): hi-res row 2i + a + k - 2 = 2m + r reads 0.25 x[m-1] + 0.75 x[m] (r = 0)
// or 0.75 x[m] + 0.25 x[m+1] (r = 1); low-res offset index = m + dm + 2 - a - i.  (tools/polyphase_check.py)
__device__ __forceinline__ float up_fold(int a, int k, int p) {
  const int s = a + k - 2;
  const int m = (s >= 0) ? s / 2 : -((-s + 1) / 2), r = s - 2 * m;
  const int p0 = m + (r == 0 ? -1 : 0) + 2 - a, p1 = p0 + 1;
  const float w0 = r == 0 ? 0.25f : 0.75f, w1 = r == 0 ? 0.75f : 0.25f;
  return (p == p0 ? w0 : 0.f) + (p == p1 ? w1 : 0.f);
}

// wp[a][b][co][ci][p][q] = sum_{k,l} w5[co][ci][k][l] fold(a,k,p) fold(b,l,q)
__global__ void __launch_bounds__(256)
up_phase_weights_kernel(const float* __restrict__ w5, int Cout, int Cin, float* __restrict__ wp) {
  const int total = 4 * Cout * Cin * 16;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
    const int q = e & 3, p = (e >> 2) & 3, rest = e >> 4;
    const int ci = rest % Cin, co = (rest / Cin) % Cout, ab = rest / (Cin * Cout);
    const int a = ab >> 1, b = ab & 1;
    const float* w = w5 + ((size_t)co * Cin + ci) * 25;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float fk = up_fold(a, k, p);
      if (fk != 0.f) {
#pragma unroll
        for (int l = 0; l < 5; ++l) s = fmaf(__ldg(w + k * 5 + l) * fk, up_fold(b, l, q), s);
      }
    }
    wp[e] = s;
  }
}

}  // namespace

// Phase weights of the polyphase resize-convolution (Upsample + Conv 5x5 of the UNet decoder levels without the upsampled
// tensor, DESIGN.md 4.5, up_poly.cu): wp [2 (row phase)][2 (x-phase)][Cout][Cin][4][4] fp32 from w5 [Cout][Cin][5][5].
CNP_API int cnp_up_phase_weights(const float* w5, int Cout, int Cin, float* wp, cudaStream_t st) {
  CNP_REQUIRE(w5 && wp && Cout > 0 && Cin > 0, "up_phase_weights: bad arguments");
  up_phase_weights_kernel<<<cnp_cdiv(4 * Cout * Cin * 16, 256), 256, 0, st>>>(w5, Cout, Cin, wp);
  CNP_LAUNCH_CHECK("up_phase_weights_kernel");
  return 0;
}

// wf [Cout][Cp][k*k] fp32 <- W5 [Cout][Cmid][k][k], W1 [Cmid][Cin], b1 [Cmid]; channels Cin+1 .. Cp-1 are zero.
CNP_API int cnp_fold_in_fwd(const float* w5, const float* w1, const float* b1, int Cout, int Cmid, int Cin, int Cp, int k,
                            float* wf, cudaStream_t st) {
  CNP_REQUIRE(w5 && w1 && b1 && wf && Cout > 0 && Cmid > 0 && Cin > 0 && Cp > Cin && k > 0, "fold_in_fwd: bad arguments");
  const int total = Cout * Cp * k * k;
  fold_in_fwd_kernel<<<cnp_cdiv(total, 256), 256, 0, st>>>(w5, w1, b1, Cout, Cmid, Cin, Cp, k * k, wf);
  CNP_LAUNCH_CHECK("fold_in_fwd_kernel");
  return 0;
}

// dw5 (+=) [Cout][Cmid][k][k], dw1 (+=) [Cmid][Cin], db1 (+=) [Cmid] from the folded gradient dwf [Cout][Cp][k*k].
CNP_API int cnp_fold_in_bwd(const float* dwf, const float* w5, const float* w1, const float* b1, int Cout, int Cmid,
                            int Cin, int Cp, int k, float* dw5, float* dw1, float* db1, cudaStream_t st) {
  CNP_REQUIRE(dwf && w5 && w1 && b1 && dw5 && dw1 && Cout > 0 && Cmid > 0 && Cin > 0 && Cp > Cin && k > 0,
              "fold_in_bwd: bad arguments");
  const int total = Cout * Cmid * k * k;
  fold_in_bwd_w5_kernel<<<cnp_cdiv(total, 256), 256, 0, st>>>(dwf, w1, b1, Cout, Cmid, Cin, Cp, k * k, dw5);
  CNP_LAUNCH_CHECK("fold_in_bwd_w5_kernel");
  fold_in_bwd_w1_kernel<<<dim3(Cmid, Cin + 1), 128, 0, st>>>(dwf, w5, Cout, Cmid, Cin, Cp, k * k, dw1, db1);
  CNP_LAUNCH_CHECK("fold_in_bwd_w1_kernel");
  return 0;
}
