// Inference decoder: SetConv from the internal grid onto an ON-GRID target (x1t x x2t, e.g. the
// 1400x1400 NZ grid) and the aux-at-target MLP + Gaussian head over every target pixel.
//
// Replaces, for ConvNP.predict(tasks, X_t=ds_elev) (nzdownscale/downscaler/validate_ERA.py:88-92,
// validate_WRF.py:227-231, validate.py:1106), the dense upstream einsum 'bcij,bip,bjq->bcpq'
// (98.6 GFLOP per task untruncated) and the cuBLAS MLP over 1.96 M rows (a6).
// The set-conv is separable and truncated at the fp32-exact radius (see common.cuh):
//   pass A: T[b,c,i,q] = sum_j z[b,c,i,j] w2[j,q]     (band of <= 32 grid columns per target column)
//   pass B: f[b,c,p,q] = sum_i w1[i,p] T[b,c,i,q]
// Band tables are analytic (the internal grid is uniform), so arbitrary target coordinates work.
#include "tc_common.cuh"   // cnp_blk
#include <math.h>

#include "mlp_params.cuh"

namespace {

constexpr int DKB = 40;  // max band (grid points within the truncation radius of one target coordinate)

// per target coordinate t: first grid index j0[t], count len[t], weights w[k][t]
__global__ void __launch_bounds__(128)
dec_band_kernel(const float* __restrict__ xt, int T, double start, int n, double res, float scale2,
                int* __restrict__ j0, int* __restrict__ len, float* __restrict__ w) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= T) return;
  const float x = xt[t];
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  int lo = (int)floor(((double)x - R - start) / res) - 1;
  int hi = (int)ceil(((double)x + R - start) / res) + 1;
  lo = max(lo, 0); hi = min(hi, n - 1);
  // shrink to the exact non-zero support so that the band fits DKB
  while (lo <= hi && cnp_rbf(x, cnp_grid_pt(start, res, lo), scale2) == 0.f) ++lo;
  while (hi >= lo && cnp_rbf(x, cnp_grid_pt(start, res, hi), scale2) == 0.f) --hi;
  const int L = min(max(hi - lo + 1, 0), DKB);
  j0[t] = lo; len[t] = L;
  for (int k = 0; k < DKB; ++k) w[(size_t)k * T + t] = (k < L) ? cnp_rbf(x, cnp_grid_pt(start, res, lo + k), scale2) : 0.f;
}

// pass A: one block per (grid row i, channel c, batch b): stage z row, produce Q outputs
__global__ void __launch_bounds__(256)
dec_grid_passA_kernel(const float* __restrict__ z, long long z_bs, int C, int n1, int n2, int Q,
                      const int* __restrict__ j0, const int* __restrict__ len, const float* __restrict__ w2,
                      float* __restrict__ T) {
  extern __shared__ float zr[];  // [n2]
  const int i = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
  const float* src = z + (size_t)b * z_bs + ((size_t)c * n1 + i) * n2;
  for (int j = threadIdx.x; j < n2; j += 256) zr[j] = src[j];
  __syncthreads();
  float* dst = T + (((size_t)b * C + c) * n1 + i) * Q;
  for (int q = threadIdx.x; q < Q; q += 256) {
    const int s = j0[q], L = len[q];
    float acc = 0.f;
    for (int k = 0; k < L; ++k) acc = fmaf(zr[s + k], __ldg(w2 + (size_t)k * Q + q), acc);
    dst[q] = acc;
  }
}

// pass B: f[b,c,p,q] = sum_k w1[k][p] T[b,c,i0[p]+k,q]; block = 8 rows p x 32 cols q, loops channels
__global__ void __launch_bounds__(256)
dec_grid_passB_kernel(const float* __restrict__ T, int C, int n1, int P, int Q, const int* __restrict__ i0,
                      const int* __restrict__ len, const float* __restrict__ w1, float* __restrict__ f,
                      long long f_bs) {
  const int q = blockIdx.x * 32 + (threadIdx.x & 31), p = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
  if (p >= P || q >= Q) return;
  const int s = i0[p], L = len[p];
  float wv[DKB];
#pragma unroll
  for (int k = 0; k < DKB; ++k) wv[k] = (k < L) ? __ldg(w1 + (size_t)k * P + p) : 0.f;
  for (int c = 0; c < C; ++c) {
    const float* Tc = T + (((size_t)b * C + c) * n1 + s) * Q + q;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < DKB; ++k)
      if (k < L) acc = fmaf(wv[k], __ldg(Tc + (size_t)k * Q), acc);
    f[(size_t)b * f_bs + ((size_t)c * P + p) * Q + q] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// per-point MLP + Gaussian head, one THREAD per target point (hidden width 64, input <= 72).
// Inputs of a layer live in registers, weights are broadcast float4 reads from shared memory,
// layer outputs go through a conflict-free [64][256] shared buffer.
// ---------------------------------------------------------------------------------------------
constexpr int GP_IN = 72;   // padded input width of layer 0
constexpr int GP_H = 64;

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

template <int IN>
__device__ __forceinline__ void gp_layer(const float* __restrict__ wl, const float* __restrict__ bl, int nout,
                                         const float (&a)[IN], float* __restrict__ hbuf, int tid) {
  for (int o = 0; o < nout; ++o) {
    const float4* wr = reinterpret_cast<const float4*>(wl + (size_t)o * IN);
    float s0 = bl[o], s1 = 0.f;
#pragma unroll
    for (int i = 0; i < IN / 4; ++i) {
      const float4 w = wr[i];
      s0 = fmaf(w.x, a[4 * i + 0], s0); s1 = fmaf(w.y, a[4 * i + 1], s1);
      s0 = fmaf(w.z, a[4 * i + 2], s0); s1 = fmaf(w.w, a[4 * i + 3], s1);
    }
    hbuf[o * 256 + tid] = s0 + s1;
  }
}

__global__ void __launch_bounds__(256, 1)
mlp_head_points_kernel(cnp_mlp_params p, const float* __restrict__ f, long long f_bs, int Cf,
                       const float* __restrict__ aux, long long aux_bs, int Ca, long long npts,
                       float* __restrict__ mean, float* __restrict__ stdv) {
  extern __shared__ __align__(16) float sm[];
  const int L = p.n_layers;
  // weights: layer 0 [64][GP_IN], hidden [64][64], last [2][64]; biases after each
  float* w0 = sm;
  float* b0 = w0 + GP_H * GP_IN;
  float* wh = b0 + GP_H;                         // (L-2) x ([64][64] + [64])
  float* wl = wh + (size_t)(L - 2) * (GP_H * GP_H + GP_H);
  float* bl = wl + 2 * GP_H;
  float* hbuf = bl + 4;                          // [64][256]
  const int in0 = p.dims[0];
  for (int e = threadIdx.x; e < GP_H * GP_IN; e += 256) {
    const int o = e / GP_IN, i = e % GP_IN;
    w0[e] = i < in0 ? p.W[0][(size_t)o * in0 + i] : 0.f;
  }
  for (int e = threadIdx.x; e < GP_H; e += 256) b0[e] = p.b[0][e];
  for (int l = 1; l < L - 1; ++l) {
    float* wd = wh + (size_t)(l - 1) * (GP_H * GP_H + GP_H);
    for (int e = threadIdx.x; e < GP_H * GP_H; e += 256) wd[e] = p.W[l][e];
    for (int e = threadIdx.x; e < GP_H; e += 256) wd[GP_H * GP_H + e] = p.b[l][e];
  }
  for (int e = threadIdx.x; e < 2 * GP_H; e += 256) wl[e] = p.W[L - 1][e];
  if (threadIdx.x < 2) bl[threadIdx.x] = p.b[L - 1][threadIdx.x];
  __syncthreads();
  const int tid = threadIdx.x, b = blockIdx.y;
  const float* fb = f + (size_t)b * f_bs;
  const float* ab = aux + (size_t)b * aux_bs;
  for (long long pt = (long long)blockIdx.x * 256 + tid; pt < npts; pt += (long long)gridDim.x * 256) {
    float a0[GP_IN];
#pragma unroll
    for (int i = 0; i < GP_IN; ++i) {
      float v = 0.f;
      if (i < Cf) v = __ldg(fb + (size_t)i * npts + pt);
      else if (i < Cf + Ca) v = __ldg(ab + (size_t)(i - Cf) * npts + pt);
      a0[i] = v;
    }
    gp_layer<GP_IN>(w0, b0, GP_H, a0, hbuf, tid);
    float a[GP_H];
    for (int l = 1; l < L - 1; ++l) {
#pragma unroll
      for (int i = 0; i < GP_H; ++i) { const float v = hbuf[i * 256 + tid]; a[i] = v < 0.f ? 0.f : v; }
      const float* wd = wh + (size_t)(l - 1) * (GP_H * GP_H + GP_H);
      gp_layer<GP_H>(wd, wd + GP_H * GP_H, GP_H, a, hbuf, tid);
    }
#pragma unroll
    for (int i = 0; i < GP_H; ++i) { const float v = hbuf[i * 256 + tid]; a[i] = v < 0.f ? 0.f : v; }
    float o0 = bl[0], o1 = bl[1];
#pragma unroll
    for (int i = 0; i < GP_H; ++i) { o0 = fmaf(wl[i], a[i], o0); o1 = fmaf(wl[GP_H + i], a[i], o1); }
    mean[(size_t)b * npts + pt] = o0;
    stdv[(size_t)b * npts + pt] = sqrtf(1e-6f + softplus_t(o1));
  }
}

}  // namespace

CNP_API long long cnp_setconv_dec_grid_workspace_bytes(int B, int C, int n1, int P, int Q) {
  long long tabs = (long long)(P + Q) * (2 * sizeof(int) + DKB * sizeof(float));
  long long T = (long long)B * C * n1 * Q * sizeof(float);
  return ((tabs + 255) / 256) * 256 + T;
}

// f [B,C,P,Q] (batch stride f_bstride) = SetConv of z [B,C,n1,n2] onto the target grid x1t[P] x x2t[Q].
CNP_API int cnp_setconv_dec_grid_fwd(const float* z, long long z_bstride, const float* x1t, const float* x2t, int B,
                                     int C, int P, int Q, double start1, int n1, double start2, int n2, double res,
                                     float scale2, float* f, long long f_bstride, void* workspace,
                                     long long workspace_bytes, cudaStream_t st) {
  CNP_REQUIRE(z && x1t && x2t && f && workspace && B > 0 && C > 0 && P > 0 && Q > 0, "dec_grid: bad arguments");
  CNP_REQUIRE(workspace_bytes >= cnp_setconv_dec_grid_workspace_bytes(B, C, n1, P, Q), "dec_grid: workspace too small");
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  CNP_REQUIRE(2.0 * R / res + 3 <= DKB, "dec_grid: decoder scale too large for the banded kernel (band > %d)", DKB);
  int* i0 = reinterpret_cast<int*>(workspace);
  int* len1 = i0 + P;
  int* j0 = len1 + P;
  int* len2 = j0 + Q;
  float* w1 = reinterpret_cast<float*>(len2 + Q);
  float* w2 = w1 + (size_t)DKB * P;
  const long long tabs = (long long)(P + Q) * (2 * sizeof(int) + DKB * sizeof(float));
  float* T = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ((tabs + 255) / 256) * 256);
  dec_band_kernel<<<cnp_cdiv(P, 128), 128, 0, st>>>(x1t, P, start1, n1, res, scale2, i0, len1, w1);
  dec_band_kernel<<<cnp_cdiv(Q, 128), 128, 0, st>>>(x2t, Q, start2, n2, res, scale2, j0, len2, w2);
  CNP_LAUNCH_CHECK("dec_band_kernel");
  dim3 ga(n1, C, B);
  dec_grid_passA_kernel<<<ga, 256, n2 * sizeof(float), st>>>(z, z_bstride, C, n1, n2, Q, j0, len2, w2, T);
  CNP_LAUNCH_CHECK("dec_grid_passA_kernel");
  dim3 gb(cnp_cdiv(Q, 32), cnp_cdiv(P, 8), B);
  dec_grid_passB_kernel<<<gb, 256, 0, st>>>(T, C, n1, P, Q, i0, len1, w1, f, f_bstride);
  CNP_LAUNCH_CHECK("dec_grid_passB_kernel");
  return 0;
}

// mean / std over npts target points per batch element; f [B,Cf,npts], aux [B or 1,Ca,npts].
// Specialised for hidden width 64 and Cf+Ca <= 72 (the configuration DeepSensor builds); returns -1 otherwise.
CNP_API int cnp_mlp_head_points_fwd(const cnp_mlp_params* p, const float* f, long long f_bstride, int Cf,
                                    const float* aux, long long aux_bstride, int Ca, int B, long long npts,
                                    float* mean, float* stdv, cudaStream_t st) {
  CNP_REQUIRE(p && f && aux && mean && stdv && B > 0 && npts > 0, "mlp_head_points: bad arguments");
  CNP_REQUIRE(p->n_layers >= 3 && p->n_layers <= CNP_MLP_MAX_LAYERS, "mlp_head_points: need >= 2 hidden layers");
  CNP_REQUIRE(p->dims[0] == Cf + Ca && p->dims[0] <= GP_IN, "mlp_head_points: input width %d > %d", p->dims[0], GP_IN);
  for (int l = 1; l < p->n_layers; ++l)
    CNP_REQUIRE(p->dims[l] == GP_H, "mlp_head_points: hidden width must be %d (got %d)", GP_H, p->dims[l]);
  CNP_REQUIRE(p->dims[p->n_layers] == 2, "mlp_head_points: last layer must have 2 outputs");
  const int L = p->n_layers;
  const size_t smem = (size_t)(GP_H * GP_IN + GP_H + (L - 2) * (GP_H * GP_H + GP_H) + 2 * GP_H + 4 + GP_H * 256) *
                      sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(mlp_head_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  long long nb = (npts + 255) / 256;
  dim3 grid((unsigned)(nb < 2 * sms ? nb : 2 * sms), B);
  mlp_head_points_kernel<<<grid, 256, smem, st>>>(*p, f, f_bstride, Cf, aux, aux_bstride, Ca, npts, mean, stdv);
  CNP_LAUNCH_CHECK("mlp_head_points_kernel");
  return 0;
}

// =============================================================================================
// Fused on-grid inference decoder on the blocked bf16 hidden activation (bf16 mode):
//   column pass   T[c,i,q]   = sum_j h[c,i,j] w2[j,q]                       (blocked bf16 in, blocked bf16 out)
//   fused tail    g[c,p,q]   = sum_i w1[i,p] T[c,i,q]                        (registers, never written)
//                 o          = MLP([Wf g + bf sw ; aux]) with the final 1x1 (Wf, bf) folded into MLP layer 0:
//                              W0'' = [W0f Wf | W0a | W0f bf],  input = [g ; aux ; sw],  sw = SetConv(1) = sw1[p] sw2[q]
//                 mean = o0, std = sqrt(1e-6 + softplus(o1))
// The 64 x P x Q decoder output (502 MB per 1400^2 task) and the full-grid 1x1 convolution are never materialised.
// Replaces, for ConvNP.predict (validate_ERA.py:88-92), final Conv1x1 + SetConv + Augment + MLP + likelihood.
// =============================================================================================
namespace {

constexpr int FZ_IN = 72;   // padded layer-0 input: 64 (g) + Ca (<= 6) + 1 (sw) <= 72

// per target coordinate: band start, length, weights and their sum (SetConv of ones)
__global__ void __launch_bounds__(128)
dec_band_sw_kernel(const float* __restrict__ xt, int T, double start, int n, double res, float scale2,
                   int* __restrict__ j0, int* __restrict__ len, float* __restrict__ w, float* __restrict__ sw) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= T) return;
  const float x = xt[t];
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  int lo = (int)floor(((double)x - R - start) / res) - 1;
  int hi = (int)ceil(((double)x + R - start) / res) + 1;
  lo = max(lo, 0); hi = min(hi, n - 1);
  while (lo <= hi && cnp_rbf(x, cnp_grid_pt(start, res, lo), scale2) == 0.f) ++lo;
  while (hi >= lo && cnp_rbf(x, cnp_grid_pt(start, res, hi), scale2) == 0.f) --hi;
  const int L = min(max(hi - lo + 1, 0), DKB);
  j0[t] = lo; len[t] = L;
  float s = 0.f;
  for (int k = 0; k < DKB; ++k) {
    const float v = (k < L) ? cnp_rbf(x, cnp_grid_pt(start, res, lo + k), scale2) : 0.f;
    w[(size_t)k * T + t] = v;
    s += v;
  }
  sw[t] = s;
}

// W0'' [64][FZ_IN] and b0 from (W0 [64][64+Ca], Wf [64][64], bf [64])
__global__ void __launch_bounds__(FZ_IN)
fold_final_kernel(const float* __restrict__ W0, const float* __restrict__ Wf, const float* __restrict__ bfin, int Ca,
                  float* __restrict__ Wout) {
  const int o = blockIdx.x, i = threadIdx.x, in0 = 64 + Ca;
  float v = 0.f;
  if (i < 64) {
    for (int c = 0; c < 64; ++c) v = fmaf(W0[(size_t)o * in0 + c], Wf[(size_t)c * 64 + i], v);
  } else if (i < 64 + Ca) {
    v = W0[(size_t)o * in0 + i];
  } else if (i == 64 + Ca) {
    for (int c = 0; c < 64; ++c) v = fmaf(W0[(size_t)o * in0 + c], bfin[c], v);
  }
  Wout[(size_t)o * FZ_IN + i] = v;
}

// column pass on the blocked layout; block = 32 target columns x 8 chunks, one grid row per blockIdx.y
__global__ void __launch_bounds__(256)
dec_cols_blk_kernel(const __nv_bfloat16* __restrict__ h, long long h_bs, int H, int W, int Q,
                    const int* __restrict__ j0, const int* __restrict__ len, const float* __restrict__ w2,
                    __nv_bfloat16* __restrict__ T) {
  const int q = blockIdx.x * 32 + (threadIdx.x & 31), chunk = threadIdx.x >> 5, i = blockIdx.y, b = blockIdx.z;
  if (q >= Q) return;
  const int Wp = W + 4;
  const __nv_bfloat16* src = h + (size_t)b * h_bs + (((size_t)chunk * (H + 4) + i + 2) * Wp + 2) * 8;
  const int s = j0[q], L = len[q];
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  for (int k = 0; k < L; ++k) {
    const float wv = __ldg(w2 + (size_t)k * Q + q);
    const uint4 pk = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(s + k) * 8));
    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float2 f2 = __bfloat1622float2(p2[c]);
      acc[2 * c] = fmaf(wv, f2.x, acc[2 * c]); acc[2 * c + 1] = fmaf(wv, f2.y, acc[2 * c + 1]);
    }
  }
  uint4 o;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int c = 0; c < 4; ++c) o2[c] = __floats2bfloat162_rn(acc[2 * c], acc[2 * c + 1]);
  *reinterpret_cast<uint4*>(T + ((((size_t)b * 8 + chunk) * H + i) * Q + q) * 8) = o;
}

// row pass + folded MLP + head; one thread per target pixel
__global__ void __launch_bounds__(256, 1)
dec_grid_mlp_fused_kernel(cnp_mlp_params p, const float* __restrict__ W0f /*[64][FZ_IN]*/, const __nv_bfloat16* __restrict__ T,
                          int H, int P, int Q, const int* __restrict__ i0, const int* __restrict__ len1,
                          const float* __restrict__ w1, const float* __restrict__ sw1, const float* __restrict__ sw2,
                          const float* __restrict__ aux, long long aux_bs, int Ca, float* __restrict__ mean,
                          float* __restrict__ stdv) {
  extern __shared__ __align__(16) float sm[];
  const int L = p.n_layers;
  float* w0 = sm;                                   // [64][FZ_IN]
  float* b0 = w0 + GP_H * FZ_IN;
  float* wh = b0 + GP_H;                            // (L-2) x ([64][64] + [64])
  float* wl = wh + (size_t)(L - 2) * (GP_H * GP_H + GP_H);
  float* bl = wl + 2 * GP_H;
  float* hbuf = bl + 4;                             // [64][256]
  for (int e = threadIdx.x; e < GP_H * FZ_IN; e += 256) w0[e] = W0f[e];
  for (int e = threadIdx.x; e < GP_H; e += 256) b0[e] = p.b[0][e];
  for (int l = 1; l < L - 1; ++l) {
    float* wd = wh + (size_t)(l - 1) * (GP_H * GP_H + GP_H);
    for (int e = threadIdx.x; e < GP_H * GP_H; e += 256) wd[e] = p.W[l][e];
    for (int e = threadIdx.x; e < GP_H; e += 256) wd[GP_H * GP_H + e] = p.b[l][e];
  }
  for (int e = threadIdx.x; e < 2 * GP_H; e += 256) wl[e] = p.W[L - 1][e];
  if (threadIdx.x < 2) bl[threadIdx.x] = p.b[L - 1][threadIdx.x];
  __syncthreads();
  const int tid = threadIdx.x, b = blockIdx.y;
  const long long npts = (long long)P * Q;
  const float* ab = aux + (size_t)b * aux_bs;
  const __nv_bfloat16* Tb = T + (size_t)b * 8 * H * Q * 8;
  for (long long pt = (long long)blockIdx.x * 256 + tid; pt < npts; pt += (long long)gridDim.x * 256) {
    const int pp = (int)(pt / Q), q = (int)(pt - (long long)pp * Q);
    const int s = i0[pp], Ln = len1[pp];
    float a0[FZ_IN];
#pragma unroll
    for (int i = 0; i < 64; ++i) a0[i] = 0.f;
    for (int k = 0; k < Ln; ++k) {
      const float wv = __ldg(w1 + (size_t)k * P + pp);
      const __nv_bfloat16* row = Tb + ((size_t)(s + k) * Q + q) * 8;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 pk = __ldg(reinterpret_cast<const uint4*>(row + (size_t)c * H * Q * 8));
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f2 = __bfloat1622float2(p2[e]);
          a0[c * 8 + 2 * e] = fmaf(wv, f2.x, a0[c * 8 + 2 * e]);
          a0[c * 8 + 2 * e + 1] = fmaf(wv, f2.y, a0[c * 8 + 2 * e + 1]);
        }
      }
    }
#pragma unroll
    for (int i = 64; i < FZ_IN; ++i) {
      float v = 0.f;
      if (i < 64 + Ca) v = __ldg(ab + (size_t)(i - 64) * npts + pt);
      else if (i == 64 + Ca) v = sw1[pp] * sw2[q];
      a0[i] = v;
    }
    gp_layer<FZ_IN>(w0, b0, GP_H, a0, hbuf, tid);
    float a[GP_H];
    for (int l = 1; l < L - 1; ++l) {
#pragma unroll
      for (int i = 0; i < GP_H; ++i) { const float v = hbuf[i * 256 + tid]; a[i] = v < 0.f ? 0.f : v; }
      const float* wd = wh + (size_t)(l - 1) * (GP_H * GP_H + GP_H);
      gp_layer<GP_H>(wd, wd + GP_H * GP_H, GP_H, a, hbuf, tid);
    }
#pragma unroll
    for (int i = 0; i < GP_H; ++i) { const float v = hbuf[i * 256 + tid]; a[i] = v < 0.f ? 0.f : v; }
    float o0 = bl[0], o1 = bl[1];
#pragma unroll
    for (int i = 0; i < GP_H; ++i) { o0 = fmaf(wl[i], a[i], o0); o1 = fmaf(wl[GP_H + i], a[i], o1); }
    mean[(size_t)b * npts + pt] = o0;
    stdv[(size_t)b * npts + pt] = sqrtf(1e-6f + softplus_t(o1));
  }
}

}  // namespace

CNP_API long long cnp_decode_grid_fused_workspace_bytes(int B, int n1, int P, int Q) {
  long long tabs = (long long)(P + Q) * (2 * sizeof(int) + (DKB + 1) * sizeof(float)) + (long long)64 * FZ_IN * sizeof(float);
  long long T = (long long)B * 64 * n1 * Q * 2;
  return ((tabs + 255) / 256) * 256 + T;
}

// mean / std [B,P,Q] on the target grid x1t[P] x x2t[Q] from the blocked 64-channel hidden activation h, the final 1x1
// (Wf [64][64], bf [64]) and the aux MLP (p: dims[0] = 64 + Ca, hidden width 64, >= 2 hidden layers).
CNP_API int cnp_decode_grid_fused_fwd(const cnp_blk* h, const float* x1t, const float* x2t, int B, int P, int Q,
                                      double start1, double start2, double res, float scale2, const float* Wf,
                                      const float* bfin, const cnp_mlp_params* p, const float* aux, long long aux_bstride,
                                      int Ca, float* mean, float* stdv, void* workspace, long long workspace_bytes,
                                      cudaStream_t st) {
  CNP_REQUIRE(h && x1t && x2t && Wf && bfin && p && aux && mean && stdv && workspace && B > 0 && P > 0 && Q > 0,
              "decode_grid_fused: bad arguments");
  const int n1 = h->H, n2 = h->W;
  CNP_REQUIRE(workspace_bytes >= cnp_decode_grid_fused_workspace_bytes(B, n1, P, Q), "decode_grid_fused: workspace too small");
  CNP_REQUIRE(p->n_layers >= 3 && p->n_layers <= CNP_MLP_MAX_LAYERS, "decode_grid_fused: need >= 2 hidden layers");
  CNP_REQUIRE(p->dims[0] == 64 + Ca && 64 + Ca + 1 <= FZ_IN, "decode_grid_fused: layer-0 width %d unsupported", p->dims[0]);
  for (int l = 1; l < p->n_layers; ++l)
    CNP_REQUIRE(p->dims[l] == GP_H, "decode_grid_fused: hidden width must be %d (got %d)", GP_H, p->dims[l]);
  CNP_REQUIRE(p->dims[p->n_layers] == 2, "decode_grid_fused: last layer must have 2 outputs");
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  CNP_REQUIRE(2.0 * R / res + 3 <= DKB, "decode_grid_fused: decoder scale too large for the banded kernel (band > %d)", DKB);
  int* i0 = reinterpret_cast<int*>(workspace);
  int* len1 = i0 + P;
  int* j0 = len1 + P;
  int* len2 = j0 + Q;
  float* w1 = reinterpret_cast<float*>(len2 + Q);
  float* w2 = w1 + (size_t)DKB * P;
  float* sw1 = w2 + (size_t)DKB * Q;
  float* sw2 = sw1 + P;
  float* W0f = sw2 + Q;
  const long long tabs = (long long)(P + Q) * (2 * sizeof(int) + (DKB + 1) * sizeof(float)) + (long long)64 * FZ_IN * sizeof(float);
  __nv_bfloat16* T = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(workspace) + ((tabs + 255) / 256) * 256);
  dec_band_sw_kernel<<<cnp_cdiv(P, 128), 128, 0, st>>>(x1t, P, start1, n1, res, scale2, i0, len1, w1, sw1);
  dec_band_sw_kernel<<<cnp_cdiv(Q, 128), 128, 0, st>>>(x2t, Q, start2, n2, res, scale2, j0, len2, w2, sw2);
  fold_final_kernel<<<64, FZ_IN, 0, st>>>(p->W[0], Wf, bfin, Ca, W0f);
  CNP_LAUNCH_CHECK("dec_band_sw_kernel");
  const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(h->base) + (size_t)h->cb_off * (n1 + 4) * (n2 + 4) * 8;
  dim3 gc(cnp_cdiv(Q, 32), n1, B);
  dec_cols_blk_kernel<<<gc, 256, 0, st>>>(hp, h->bstride, n1, n2, Q, j0, len2, w2, T);
  CNP_LAUNCH_CHECK("dec_cols_blk_kernel");
  const int L = p->n_layers;
  const size_t smem = (size_t)(GP_H * FZ_IN + GP_H + (L - 2) * (GP_H * GP_H + GP_H) + 2 * GP_H + 4 + GP_H * 256) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(dec_grid_mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const long long nb = ((long long)P * Q + 255) / 256;
  dim3 grid((unsigned)(nb < 2 * sms ? nb : 2 * sms), B);
  dec_grid_mlp_fused_kernel<<<grid, 256, smem, st>>>(*p, W0f, T, n1, P, Q, i0, len1, w1, sw1, sw2, aux, aux_bstride, Ca,
                                                     mean, stdv);
  CNP_LAUNCH_CHECK("dec_grid_mlp_fused_kernel");
  return 0;
}

// =============================================================================================
// Tensor-core inference decoder tail (bf16 mode).  For a tile of 128 consecutive target columns q of one target
// row p everything after the row pass is a chain of small GEMMs with the 128 pixels as M:
//     g   [128 x 64]  = W2band [128 q x 64 j]  x  T'[p] [64 j x 64 c]        SetConv column pass (W2 split hi + lo bf16)
//     a1  = relu(W0'' [g ; aux ; sw] + b0)    a2 = relu(W1 a1 + b1)    a3 = relu(W2 a2 + b2)      tcgen05, K = 80 / 64
//     mean, std from the 64 -> 2 head in registers
// where T'[b][p][chunk][j][8] (bf16) is the ROW pass of the hidden activation (dec_rows_blk_kernel).  A CTA owns one
// q tile (its W2band and the MLP weights stay in shared memory) and a segment of target rows; its 16 warps form 4
// independent "lanes" of 4 warps (one TMEM lane quadrant each, thread = pixel) that walk the rows p = lane, lane+4, ...:
// each lane issues its own MMAs (one elected thread), waits for them on its own mbarrier, reads the accumulator with
// tcgen05.ld, applies bias / ReLU, writes the next bf16 A operand back to its shared-memory buffer and repeats -- the
// tensor pipe is shared by the 4 lanes, so while one lane is in an epilogue the others' MMAs run.
// =============================================================================================
namespace {

constexpr int TCD_LANES = 4;
constexpr int TCD_KIN = 80;                 // layer-0 K: 64 (g) + 8 (aux, sw, pad) + 8 (pad)
constexpr int TCD_ABUF = 10 * 2048;         // [k8 (10)][px (128)][8] bf16
constexpr int TCD_TT = 8 * 1024;            // [chunk (8)][j (64)][8] bf16
// shared-memory map (bytes)
constexpr int TCD_OFF_WG = 0;                                   // hi, lo: 2 x [k8 (8)][q (128)][8] bf16
constexpr int TCD_OFF_W0 = TCD_OFF_WG + 2 * 16384;              // [k8 (10)][n (64)][8] bf16
constexpr int TCD_OFF_W1 = TCD_OFF_W0 + 10 * 1024;
constexpr int TCD_OFF_W2 = TCD_OFF_W1 + 8 * 1024;
constexpr int TCD_OFF_F32 = TCD_OFF_W2 + 8 * 1024;              // b0,b1,b2 [64] each, wl [2][64], bl [2] (+pad) fp32
constexpr int TCD_OFF_LANE = TCD_OFF_F32 + (3 * 64 + 128 + 8) * 4;
constexpr int TCD_LANE_BYTES = TCD_ABUF + 2 * TCD_TT;
constexpr int TCD_OFF_BAR = TCD_OFF_LANE + TCD_LANES * TCD_LANE_BYTES;
constexpr int TCD_SMEM = TCD_OFF_BAR + TCD_LANES * 3 * 8 + 16;

// row pass on the blocked layout: T'[b][p][chunk][j][8] = sum_k w1[k][p] h[b][chunk][i0[p]+k][j][8]
__global__ void __launch_bounds__(256)
dec_rows_blk_kernel(const __nv_bfloat16* __restrict__ h, long long h_bs, int H, int W, int P,
                    const int* __restrict__ i0, const int* __restrict__ len, const float* __restrict__ w1,
                    __nv_bfloat16* __restrict__ T) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31), chunk = threadIdx.x >> 5, p = blockIdx.y, b = blockIdx.z;
  if (j >= W) return;
  const int Wp = W + 4;
  const int s = i0[p], L = len[p];
  const __nv_bfloat16* src = h + (size_t)b * h_bs + (((size_t)chunk * (H + 4) + s + 2) * Wp + j + 2) * 8;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  for (int k = 0; k < L; ++k) {
    const float wv = __ldg(w1 + (size_t)k * P + p);
    const uint4 pk = __ldg(reinterpret_cast<const uint4*>(src + (size_t)k * Wp * 8));
    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float2 f2 = __bfloat1622float2(p2[c]);
      acc[2 * c] = fmaf(wv, f2.x, acc[2 * c]); acc[2 * c + 1] = fmaf(wv, f2.y, acc[2 * c + 1]);
    }
  }
  uint4 o;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int c = 0; c < 4; ++c) o2[c] = __floats2bfloat162_rn(acc[2 * c], acc[2 * c + 1]);
  *reinterpret_cast<uint4*>(T + ((((size_t)b * P + p) * 8 + chunk) * W + j) * 8) = o;
}

__device__ __forceinline__ void tcd_lane_bar(int lane_id) {   // named barrier of the 128 threads of one lane
  asm volatile("bar.sync %0, 128;" ::"r"(lane_id + 1) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// store 64 fp32 values of one pixel as bf16 into chunks 0..7 of a K-major A buffer
__device__ __forceinline__ void tcd_store_act(uint8_t* abuf, int px, const float* v) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 pk;
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int i = 0; i < 4; ++i) p2[i] = __floats2bfloat162_rn(v[c * 8 + 2 * i], v[c * 8 + 2 * i + 1]);
    *reinterpret_cast<uint4*>(abuf + c * 2048 + px * 16) = pk;
  }
}

__global__ void __launch_bounds__(128 * TCD_LANES, 1)
dec_tc_kernel(cnp_mlp_params mp, const float* __restrict__ W0f /*[64][FZ_IN] folded*/, const __nv_bfloat16* __restrict__ T,
              int n2, int P, int Q, const float* __restrict__ x2t, double start2, double res, float scale2,
              const int* __restrict__ j0tab, const float* __restrict__ sw1, const float* __restrict__ sw2,
              const float* __restrict__ aux, long long aux_bs, int Ca, float* __restrict__ mean, float* __restrict__ stdv,
              int rows_per_cta) {
  extern __shared__ __align__(128) uint8_t tsm[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ln = warp >> 2;                    // lane (group of 4 warps)
  const int px = tid & 127;                    // pixel of this thread inside the q tile = TMEM lane
  const int q0 = blockIdx.x * 128, b = blockIdx.z;
  const int p_begin = blockIdx.y * rows_per_cta, p_end = min(P, p_begin + rows_per_cta);
  float* f32 = reinterpret_cast<float*>(tsm + TCD_OFF_F32);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + TCD_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tsm + TCD_OFF_BAR + TCD_LANES * 3 * 8);
  // band start of this q tile (all 128 columns must fit in 64 grid columns; checked by the host wrapper's caller)
  int jlo = min(j0tab[min(q0, Q - 1)], j0tab[min(q0 + 127, Q - 1)]);   // ascending or descending target coordinates
  jlo = max(0, min(jlo, n2 - 64));
  // ---- one-time setup: W2 band (hi/lo), MLP weights as bf16 B operands, biases ----
  for (int e = tid; e < 128 * 64; e += 128 * TCD_LANES) {
    const int q = e >> 6, j = e & 63;
    const float xq = x2t[min(q0 + q, Q - 1)];
    const float w = (q0 + q < Q) ? cnp_rbf(xq, cnp_grid_pt(start2, res, jlo + j), scale2) : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    const int off = (j >> 3) * 2048 + q * 16 + (j & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(tsm + TCD_OFF_WG + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(tsm + TCD_OFF_WG + 16384 + off) = lo;
  }
  for (int e = tid; e < 64 * TCD_KIN; e += 128 * TCD_LANES) {
    const int n = e / TCD_KIN, k = e % TCD_KIN;
    const float v = k < FZ_IN ? W0f[(size_t)n * FZ_IN + k] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(tsm + TCD_OFF_W0 + (k >> 3) * 1024 + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(v);
  }
  for (int e = tid; e < 64 * 64; e += 128 * TCD_LANES) {
    const int n = e >> 6, k = e & 63;
    *reinterpret_cast<__nv_bfloat16*>(tsm + TCD_OFF_W1 + (k >> 3) * 1024 + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(mp.W[1][e]);
    *reinterpret_cast<__nv_bfloat16*>(tsm + TCD_OFF_W2 + (k >> 3) * 1024 + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(mp.W[2][e]);
  }
  if (tid < 64) { f32[tid] = mp.b[0][tid]; f32[64 + tid] = mp.b[1][tid]; f32[128 + tid] = mp.b[2][tid]; }
  if (tid < 128) f32[192 + tid] = mp.W[3][tid];
  if (tid < 2) f32[320 + tid] = mp.b[3][tid];
  uint8_t* lane_base = tsm + TCD_OFF_LANE + ln * TCD_LANE_BYTES;
  uint8_t* abuf = lane_base;
  uint8_t* tt = lane_base + TCD_ABUF;          // two T' tile buffers
  // chunks 8 and 9 of the A buffer (aux / sw / zero pad): zero once, chunk 8 is rewritten per tile
  *reinterpret_cast<uint4*>(abuf + 8 * 2048 + px * 16) = make_uint4(0, 0, 0, 0);
  *reinterpret_cast<uint4*>(abuf + 9 * 2048 + px * 16) = make_uint4(0, 0, 0, 0);
  uint64_t* tile_full = bars + ln * 3;         // [2]
  uint64_t* mma_done = bars + ln * 3 + 2;
  if ((tid & 127) == 0) { tc::mbar_init(tile_full, 1); tc::mbar_init(tile_full + 1, 1); tc::mbar_init(mma_done, 1); tc::mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot + ln * 128;                         // two 64-column accumulators per lane
  const uint32_t tl = ((uint32_t)((warp & 3) * 32)) << 16;             // this warp's TMEM lane quadrant
  const bool leader = (tid & 127) == 0;
  const uint32_t idesc_g = tc::make_idesc_bf16(128, 64, 0, 1);         // A K-major (W2 band), B MN-major (T' tile)
  const uint32_t idesc_m = tc::make_idesc_bf16(128, 64, 0, 0);         // A K-major (activations), B K-major (weights)
  const uint32_t hiA = (128u >> 4) | (1u << 14);                       // A: SBO 128 B
  const uint32_t lboA = (2048u >> 4) << 16;                            // A: LBO 2048 B (k8 planes)
  const uint32_t hiW = (128u >> 4) | (1u << 14), lboW = (1024u >> 4) << 16;   // weights: LBO 1024, SBO 128
  const uint32_t hiT = (1024u >> 4) | (1u << 14), lboT = (128u >> 4) << 16;   // T' tile (MN-major): SBO 1024 (chunk), LBO 128
  const size_t npts = (size_t)P * Q;
  const float* ab = aux + (size_t)b * aux_bs;
  const __nv_bfloat16* Tb = T + (size_t)b * P * 8 * n2 * 8;
  const int q = q0 + px;
  const bool q_ok = q < Q;
  const float swq = q_ok ? sw2[q] : 0.f;
  uint32_t mma_phase = 0;

  auto load_tile = [&](int p, int buf) {       // leader only: 8 chunk rows of 64 grid columns (1 KB each)
    tc::mbar_expect_tx(tile_full + buf, 8u * 1024u);
    const __nv_bfloat16* src = Tb + (((size_t)p * 8) * n2 + jlo) * 8;
    for (int c = 0; c < 8; ++c)
      tc::bulk_g2s(tt + buf * TCD_TT + c * 1024, src + (size_t)c * n2 * 8, 1024u, tile_full + buf);
  };
  auto wait_mma = [&]() {
    tc::mbar_wait(mma_done, mma_phase & 1);
    ++mma_phase;
    tc::fence_after_sync();
  };

  int it = 0;
  if (leader && p_begin + ln < p_end) load_tile(p_begin + ln, 0);
  for (int p = p_begin + ln; p < p_end; p += TCD_LANES, ++it) {
    const int buf = it & 1;
    // aux / sw for this pixel (global loads in flight during the column-pass MMAs)
    float av[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) av[c] = (c < Ca && q_ok) ? __ldg(ab + (size_t)c * npts + (size_t)p * Q + q) : 0.f;
    if (leader) {
      tc::mbar_wait(tile_full + buf, (it >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t t16 = tc::smem_u32(tt + buf * TCD_TT) >> 4, w16 = tc::smem_u32(tsm + TCD_OFF_WG) >> 4;
#pragma unroll
      for (int hl = 0; hl < 2; ++hl)
#pragma unroll
        for (int k = 0; k < 4; ++k)            // K = 16 grid columns per MMA: 2 k8 planes of A, 2 K groups of B
          tc::mma_bf16_ss_lohi(tmem, ((w16 + hl * 1024 + k * 256) & 0x3FFFu) | lboA, hiA, ((t16 + k * 16) & 0x3FFFu) | lboT, hiT,
                               idesc_g, (hl | k) ? 1u : 0u);
      tc::mma_commit(mma_done);
      if (p + TCD_LANES < p_end) load_tile(p + TCD_LANES, buf ^ 1);
    }
    wait_mma();
    float v[64];
    tc::tmem_ld32(tmem + tl, v);
    tc::tmem_ld32(tmem + tl + 32, v + 32);
    tc::tmem_ld_wait();
    // ---- A0 = [g ; aux ; sw ; 0] ----
    tcd_store_act(abuf, px, v);
    {
      uint4 pk;
      __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
      float ex[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) ex[c] = 0.f;
#pragma unroll
      for (int c = 0; c < 6; ++c) if (c < Ca) ex[c] = av[c];
      ex[Ca] = sw1[p] * swq;
#pragma unroll
      for (int i = 0; i < 4; ++i) p2[i] = __floats2bfloat162_rn(ex[2 * i], ex[2 * i + 1]);
      *reinterpret_cast<uint4*>(abuf + 8 * 2048 + px * 16) = pk;
    }
    // ---- three hidden layers on the tensor cores ----
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      fence_proxy_async();
      tc::fence_before_sync();
      tcd_lane_bar(ln);
      if (leader) {
        tc::fence_after_sync();
        const uint32_t a16 = tc::smem_u32(abuf) >> 4;
        const uint32_t w16 = tc::smem_u32(tsm + (l == 0 ? TCD_OFF_W0 : (l == 1 ? TCD_OFF_W1 : TCD_OFF_W2))) >> 4;
        const int nk = l == 0 ? TCD_KIN / 16 : 4;
        const uint32_t d = tmem + ((l & 1) ? 0u : 64u);
        for (int k = 0; k < nk; ++k)
          tc::mma_bf16_ss_lohi(d, ((a16 + k * 256) & 0x3FFFu) | lboA, hiA, ((w16 + k * 128) & 0x3FFFu) | lboW, hiW, idesc_m,
                               k ? 1u : 0u);
        tc::mma_commit(mma_done);
      }
      wait_mma();
      const uint32_t d = tmem + ((l & 1) ? 0u : 64u);
      tc::tmem_ld32(d + tl, v);
      tc::tmem_ld32(d + tl + 32, v + 32);
      tc::tmem_ld_wait();
      const float* bias = f32 + 64 * l;
#pragma unroll
      for (int i = 0; i < 64; ++i) { const float t = v[i] + bias[i]; v[i] = t < 0.f ? 0.f : t; }
      if (l < 2) tcd_store_act(abuf, px, v);
    }
    // ---- head ----
    float o0 = f32[320], o1 = f32[321];
#pragma unroll
    for (int i = 0; i < 64; ++i) { o0 = fmaf(f32[192 + i], v[i], o0); o1 = fmaf(f32[256 + i], v[i], o1); }
    if (q_ok) {
      mean[(size_t)b * npts + (size_t)p * Q + q] = o0;
      stdv[(size_t)b * npts + (size_t)p * Q + q] = sqrtf(1e-6f + softplus_t(o1));
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc::fence_after_sync(); tc::tmem_dealloc(*tmem_slot, 512); }
}

}  // namespace

CNP_API long long cnp_decode_grid_tc_workspace_bytes(int B, int n2, int P, int Q) {
  long long tabs = (long long)(P + Q) * (2 * sizeof(int) + (DKB + 1) * sizeof(float)) + (long long)64 * FZ_IN * sizeof(float);
  long long T = (long long)B * P * 64 * n2 * 2;
  return ((tabs + 255) / 256) * 256 + T;
}

// Tensor-core variant of cnp_decode_grid_fused_fwd (same arguments).  Requirements checked here: 3 hidden layers of
// width 64, 64 + Ca + 1 <= 72, n2 >= 64.  The caller guarantees that any 128 consecutive target columns x2t span at
// most 64 - band internal-grid columns (true whenever the target grid is at least ~4x finer than the internal grid).
CNP_API int cnp_decode_grid_tc_fwd(const cnp_blk* h, const float* x1t, const float* x2t, int B, int P, int Q,
                                   double start1, double start2, double res, float scale2, const float* Wf,
                                   const float* bfin, const cnp_mlp_params* p, const float* aux, long long aux_bstride,
                                   int Ca, float* mean, float* stdv, void* workspace, long long workspace_bytes,
                                   cudaStream_t st) {
  CNP_REQUIRE(h && x1t && x2t && Wf && bfin && p && aux && mean && stdv && workspace && B > 0 && P > 0 && Q > 0,
              "decode_grid_tc: bad arguments");
  const int n1 = h->H, n2 = h->W;
  CNP_REQUIRE(n2 >= 64, "decode_grid_tc: internal grid too narrow");
  CNP_REQUIRE(workspace_bytes >= cnp_decode_grid_tc_workspace_bytes(B, n2, P, Q), "decode_grid_tc: workspace too small");
  CNP_REQUIRE(p->n_layers == 4 && p->dims[0] == 64 + Ca && 64 + Ca + 1 <= FZ_IN && Ca <= 6, "decode_grid_tc: MLP shape unsupported");
  for (int l = 1; l < 4; ++l) CNP_REQUIRE(p->dims[l] == 64, "decode_grid_tc: hidden width must be 64");
  CNP_REQUIRE(p->dims[4] == 2, "decode_grid_tc: last layer must have 2 outputs");
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  CNP_REQUIRE(2.0 * R / res + 3 <= DKB, "decode_grid_tc: decoder scale too large (band > %d)", DKB);
  int* i0 = reinterpret_cast<int*>(workspace);
  int* len1 = i0 + P;
  int* j0 = len1 + P;
  int* len2 = j0 + Q;
  float* w1 = reinterpret_cast<float*>(len2 + Q);
  float* w2 = w1 + (size_t)DKB * P;
  float* sw1 = w2 + (size_t)DKB * Q;
  float* sw2 = sw1 + P;
  float* W0f = sw2 + Q;
  const long long tabs = (long long)(P + Q) * (2 * sizeof(int) + (DKB + 1) * sizeof(float)) + (long long)64 * FZ_IN * sizeof(float);
  __nv_bfloat16* T = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(workspace) + ((tabs + 255) / 256) * 256);
  dec_band_sw_kernel<<<cnp_cdiv(P, 128), 128, 0, st>>>(x1t, P, start1, n1, res, scale2, i0, len1, w1, sw1);
  dec_band_sw_kernel<<<cnp_cdiv(Q, 128), 128, 0, st>>>(x2t, Q, start2, n2, res, scale2, j0, len2, w2, sw2);
  fold_final_kernel<<<64, FZ_IN, 0, st>>>(p->W[0], Wf, bfin, Ca, W0f);
  CNP_LAUNCH_CHECK("dec_band_sw_kernel");
  const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(h->base) + (size_t)h->cb_off * (n1 + 4) * (n2 + 4) * 8;
  dim3 gr(cnp_cdiv(n2, 32), P, B);
  dec_rows_blk_kernel<<<gr, 256, 0, st>>>(hp, h->bstride, n1, n2, P, i0, len1, w1, T);
  CNP_LAUNCH_CHECK("dec_rows_blk_kernel");
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(dec_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCD_SMEM); attr = true; }
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const int nqt = cnp_cdiv(Q, 128);
  int nseg = sms / (nqt * B);
  if (nseg < 1) nseg = 1;
  int rows_per_cta = cnp_cdiv(P, nseg);
  rows_per_cta = cnp_cdiv(rows_per_cta, TCD_LANES) * TCD_LANES;
  dim3 grid(nqt, cnp_cdiv(P, rows_per_cta), B);
  dec_tc_kernel<<<grid, 128 * TCD_LANES, TCD_SMEM, st>>>(*p, W0f, T, n2, P, Q, x2t, start2, res, scale2, j0, sw1, sw2, aux,
                                                        aux_bstride, Ca, mean, stdv, rows_per_cta);
  CNP_LAUNCH_CHECK("dec_tc_kernel");
  return 0;
}
