// Training-path decoder on the blocked bf16 UNet activations: SetConv (grid -> off-grid targets) applied to the
// LAST HIDDEN activation h (64 channels, after ReLU) instead of to z = final_linear(h), with the 1x1 final
// convolution moved behind it -- both are linear, so
//     f[c,t] = sum_ij w1[i,t] w2[j,t] (sum_k Wf[c,k] h[k,i,j] + bf[c])
//            = sum_k Wf[c,k] g[k,t] + bf[c] sw[t],   g[k,t] = sum_ij w1 w2 h[k,i,j],  sw[t] = sum_ij w1 w2.
// The full-grid 1x1 convolution, its fp32 output z (378 MB per 16-task step), the dense dz and the 1x1
// dgrad / wgrad launches disappear; what is left touches only the ~31x31 windows around the targets plus one
// dense, coalesced write of d_h (ReLU mask fused).
//
// Replaces, for ConvNP.loss_fn (nzdownscale/downscaler/train.py:370, train_epoch train.py:388-394), the tail
// of upstream neuralprocesses' UNet (final Conv 1x1, coders/nn.py) + decoder SetConv (coders/setconv) and their
// autograd.
#include "tc_common.cuh"
#include <math.h>

namespace {

constexpr int DH_C = 64;        // hidden channels (8 chunks)
constexpr int DH_MAXW = 48;     // max window extent per dimension (2R/res + slack); larger scales are rejected

__device__ __forceinline__ void ld8f(const __nv_bfloat16* p, float* v) {
  const uint4 pk = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(p2[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// one block per (target t, batch b); warp c = channel chunk c, lanes stride over the window pixels
__global__ void __launch_bounds__(256)
dec_blk_fwd_kernel(const __nv_bfloat16* __restrict__ h, long long h_bs, int H, int W, const float* __restrict__ xt,
                   int Nt, double start1, double start2, double res, float scale2, const float* __restrict__ Wf,
                   const float* __restrict__ bfin, int Cz, float* __restrict__ g, float* __restrict__ sw,
                   float* __restrict__ f) {
  __shared__ float w1s[DH_MAXW], w2s[DH_MAXW];
  __shared__ float gs[DH_C];
  __shared__ float sws[2];
  __shared__ int rng[4];
  const int t = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float p1 = xt[((size_t)b * 2 + 0) * Nt + t], p2 = xt[((size_t)b * 2 + 1) * Nt + t];
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  if (threadIdx.x == 0) {
    int ilo = (int)floor(((double)p1 - R - start1) / res) - 1, ihi = (int)ceil(((double)p1 + R - start1) / res) + 1;
    int jlo = (int)floor(((double)p2 - R - start2) / res) - 1, jhi = (int)ceil(((double)p2 + R - start2) / res) + 1;
    ilo = max(ilo, 0); ihi = min(ihi + 1, H); jlo = max(jlo, 0); jhi = min(jhi + 1, W);
    rng[0] = ilo; rng[1] = min(max(ihi - ilo, 0), DH_MAXW); rng[2] = jlo; rng[3] = min(max(jhi - jlo, 0), DH_MAXW);
  }
  __syncthreads();
  const int ilo = rng[0], ni = rng[1], jlo = rng[2], nj = rng[3];
  if (threadIdx.x < DH_MAXW)
    w1s[threadIdx.x] = threadIdx.x < ni ? cnp_rbf(p1, cnp_grid_pt(start1, res, ilo + threadIdx.x), scale2) : 0.f;
  else if (threadIdx.x >= 64 && threadIdx.x < 64 + DH_MAXW) {
    const int j = threadIdx.x - 64;
    w2s[j] = j < nj ? cnp_rbf(p2, cnp_grid_pt(start2, res, jlo + j), scale2) : 0.f;
  }
  __syncthreads();
  const int Wp = W + 4;
  const __nv_bfloat16* hc = h + (size_t)b * h_bs + (size_t)warp * (H + 4) * Wp * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int npx = ni * nj;
  for (int e = lane; e < npx; e += 32) {
    const int ii = e / nj, jj = e - ii * nj;
    const float w = w1s[ii] * w2s[jj];
    float v[8];
    ld8f(hc + ((size_t)(ilo + ii + 2) * Wp + (jlo + jj + 2)) * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, v[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) gs[warp * 8 + i] = acc[i];
  }
  if (threadIdx.x == 0) {
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < ni; ++i) s1 += w1s[i];
    for (int j = 0; j < nj; ++j) s2 += w2s[j];
    sws[0] = s1 * s2;
  }
  __syncthreads();
  const size_t bt = (size_t)b * Nt + t;
  if (threadIdx.x < DH_C) g[bt * DH_C + threadIdx.x] = gs[threadIdx.x];
  if (threadIdx.x == 0) sw[bt] = sws[0];
  for (int c = threadIdx.x; c < Cz; c += 256) {
    const float* wr = Wf + (size_t)c * DH_C;
    float s = bfin[c] * sws[0];
#pragma unroll 8
    for (int k = 0; k < DH_C; ++k) s = fmaf(__ldg(wr + k), gs[k], s);
    f[((size_t)b * Cz + c) * Nt + t] = s;
  }
}

// dg[b,t,k] = sum_c Wf[c,k] df[b,c,t]                      one block per (t, b)
__global__ void __launch_bounds__(64)
dec_blk_dg_kernel(const float* __restrict__ df, const float* __restrict__ Wf, int Cz, int Nt, float* __restrict__ dg) {
  __shared__ float dfs[256];
  const int t = blockIdx.x, b = blockIdx.y, k = threadIdx.x;
  for (int c = threadIdx.x; c < Cz; c += 64) dfs[c] = df[((size_t)b * Cz + c) * Nt + t];
  __syncthreads();
  float s = 0.f;
  for (int c = 0; c < Cz; ++c) s = fmaf(__ldg(Wf + (size_t)c * DH_C + k), dfs[c], s);
  dg[((size_t)b * Nt + t) * DH_C + k] = s;
}

// dWf[c,k] += sum_{b,t} df[b,c,t] g[b,t,k];  dbf[c] += sum_{b,t} df[b,c,t] sw[b,t]      one block per output channel c:
// 8 slices of the (task, target) range per hidden channel k, each walked in order, then added 0..7 -- a fixed summation
// order (run-to-run identical gradients; the first version ran one block per (c, b) and met in dWf with atomics)
constexpr int DW_PARTS = 8;
__global__ void __launch_bounds__(64 * DW_PARTS)
dec_blk_dw_kernel(const float* __restrict__ df, const float* __restrict__ g, const float* __restrict__ sw, int B, int Cz,
                  int Nt, float* __restrict__ dWf, float* __restrict__ dbf) {
  __shared__ float ps[DW_PARTS][64], pb[DW_PARTS];
  const int c = blockIdx.x, k = threadIdx.x & 63, part = threadIdx.x >> 6;
  const int total = B * Nt, per = (total + DW_PARTS - 1) / DW_PARTS;
  const int e0 = part * per, e1 = min(total, e0 + per);
  float s = 0.f, sb = 0.f;
  for (int e = e0; e < e1; ++e) {
    const int b = e / Nt, t = e - b * Nt;
    const float d = __ldg(df + ((size_t)b * Cz + c) * Nt + t);
    s = fmaf(d, __ldg(g + ((size_t)b * Nt + t) * DH_C + k), s);
    if (k == 0) sb += d * __ldg(sw + (size_t)b * Nt + t);
  }
  ps[part][k] = s;
  if (k == 0) pb[part] = sb;
  __syncthreads();
  if (part == 0) {
    float a = 0.f, ab = 0.f;
#pragma unroll
    for (int q = 0; q < DW_PARTS; ++q) { a += ps[q][k]; ab += pb[q]; }
    dWf[(size_t)c * DH_C + k] += a;
    if (k == 0 && dbf) dbf[c] += ab;
  }
}

// d_h[b,k,i,j] = (h > 0) * sum_t dg[b,t,k] w1[i,t] w2[j,t], written densely in the blocked layout.
// Block = 8 rows x 32 pixels; targets whose support touches the tile are compacted into shared memory.
constexpr int DT_I = 8, DT_J = 32, DT_MAXT = 32;

__global__ void __launch_bounds__(256)
dec_blk_bwd_kernel(const float* __restrict__ dg, const float* __restrict__ xt, int Nt, double start1, double start2,
                   double res, float scale2, const __nv_bfloat16* __restrict__ h, long long h_bs,
                   __nv_bfloat16* __restrict__ dh, long long dh_bs, int H, int W) {
  __shared__ float g1s[DT_I], g2s[DT_J];
  __shared__ float w1s[DT_MAXT][DT_I];
  __shared__ float w2s[DT_MAXT][DT_J + 1];
  __shared__ __align__(16) float dgs[DT_MAXT][DH_C];
  __shared__ int sel[DT_MAXT];
  __shared__ int nsel_s;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, tid = threadIdx.x;
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * DT_I, j0 = blockIdx.x * DT_J;
  if (tid < DT_I) g1s[tid] = cnp_grid_pt(start1, res, min(i0 + tid, H - 1));
  if (tid >= 32 && tid < 32 + DT_J) g2s[tid - 32] = cnp_grid_pt(start2, res, min(j0 + tid - 32, W - 1));
  __syncthreads();
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  const float a1 = g1s[0], b1 = g1s[DT_I - 1], a2 = g2s[0], b2 = g2s[DT_J - 1];
  const float* xb = xt + (size_t)b * 2 * Nt;
  const int i = i0 + ty, j = j0 + tx;
  const bool inb = (i < H && j < W);
  float acc[DH_C];
#pragma unroll
  for (int k = 0; k < DH_C; ++k) acc[k] = 0.f;
  bool any = false;
  int tnext = 0;
  while (tnext < Nt) {
    __syncthreads();
    if (ty == 0) {   // warp 0 tests 32 targets at once; order-preserving compaction by ballot
      const int t = tnext + tx;
      bool hit = false;
      if (t < Nt) {
        const float p1 = xb[t], p2 = xb[Nt + t];
        hit = p1 >= a1 - R && p1 <= b1 + R && p2 >= a2 - R && p2 <= b2 + R;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) sel[__popc(m & ((1u << tx) - 1u))] = t;
      if (tx == 0) nsel_s = __popc(m);
    }
    __syncthreads();
    const int total = nsel_s;
    tnext += 32;
    if (total == 0) continue;
    any = true;
    for (int m = ty; m < total; m += 8) {
      const int tt = sel[m];
      w2s[m][tx] = cnp_rbf(xb[Nt + tt], g2s[tx], scale2);
      if (tx < DT_I) w1s[m][tx] = cnp_rbf(xb[tt], g1s[tx], scale2);
    }
    for (int e = tid; e < total * DH_C; e += 256) {
      const int m = e / DH_C, k = e % DH_C;
      dgs[m][k] = dg[((size_t)b * Nt + sel[m]) * DH_C + k];
    }
    __syncthreads();
    for (int m = 0; m < total; ++m) {
      const float w = w1s[m][ty] * w2s[m][tx];
      if (w != 0.f) {
        const float4* d4 = reinterpret_cast<const float4*>(dgs[m]);
#pragma unroll
        for (int k4 = 0; k4 < DH_C / 4; ++k4) {
          const float4 d = d4[k4];
          acc[4 * k4 + 0] = fmaf(d.x, w, acc[4 * k4 + 0]); acc[4 * k4 + 1] = fmaf(d.y, w, acc[4 * k4 + 1]);
          acc[4 * k4 + 2] = fmaf(d.z, w, acc[4 * k4 + 2]); acc[4 * k4 + 3] = fmaf(d.w, w, acc[4 * k4 + 3]);
        }
      }
    }
  }
  if (!inb) return;
  const int Wp = W + 4;
  const size_t plane = (size_t)(H + 4) * Wp * 8;
  const size_t pix = ((size_t)(i + 2) * Wp + (j + 2)) * 8;
  __nv_bfloat16* dhb = dh + (size_t)b * dh_bs + pix;
  if (!any) {
#pragma unroll
    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(dhb + c * plane) = make_uint4(0, 0, 0, 0);
    return;
  }
  const __nv_bfloat16* hb = h + (size_t)b * h_bs + pix;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float hv[8];
    ld8f(hb + c * plane, hv);
    uint4 pk;
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      p2[q] = __floats2bfloat162_rn(hv[2 * q] > 0.f ? acc[c * 8 + 2 * q] : 0.f, hv[2 * q + 1] > 0.f ? acc[c * 8 + 2 * q + 1] : 0.f);
    *reinterpret_cast<uint4*>(dhb + c * plane) = pk;
  }
}

}  // namespace

// g [B,Nt,64], sw [B,Nt], f [B,Cz,Nt] from the 64-channel blocked activation h (chunks [h->cb_off, +8)).
CNP_API int cnp_dec_blk_fwd(const cnp_blk* h, const float* xt, int B, int Nt, double start1, double start2, double res,
                            float scale2, const float* Wf, const float* bfin, int Cz, float* g, float* sw, float* f,
                            cudaStream_t st) {
  CNP_REQUIRE(h && xt && Wf && bfin && g && sw && f && B > 0 && Cz > 0 && Nt >= 0, "dec_blk_fwd: bad arguments");
  if (Nt == 0) return 0;
  const float R = sqrtf(2.0f * CNP_EXP_CUTOFF * scale2);
  CNP_REQUIRE(2.0 * R / res + 6 <= DH_MAXW, "dec_blk_fwd: decoder scale too large for the windowed kernel");
  const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(h->base) + (size_t)h->cb_off * (h->H + 4) * (h->W + 4) * 8;
  dim3 grid(Nt, B);
  dec_blk_fwd_kernel<<<grid, 256, 0, st>>>(hp, h->bstride, h->H, h->W, xt, Nt, start1, start2, res, scale2, Wf, bfin, Cz,
                                           g, sw, f);
  CNP_LAUNCH_CHECK("dec_blk_fwd_kernel");
  return 0;
}

// Parameter / target-side backward: dg [B,Nt,64] = Wf^T df, dWf += df g^T, dbf += df sw.
CNP_API int cnp_dec_blk_bwd_params(const float* df, const float* g, const float* sw, const float* Wf, int B, int Nt,
                                   int Cz, float* dg, float* dWf, float* dbf, cudaStream_t st) {
  CNP_REQUIRE(df && g && sw && Wf && dg && dWf && B > 0 && Cz > 0 && Cz <= 256 && Nt >= 0, "dec_blk_bwd_params: bad arguments");
  if (Nt == 0) return 0;
  dim3 grid(Nt, B);
  dec_blk_dg_kernel<<<grid, 64, 0, st>>>(df, Wf, Cz, Nt, dg);
  CNP_LAUNCH_CHECK("dec_blk_dg_kernel");
  dec_blk_dw_kernel<<<Cz, 64 * DW_PARTS, 0, st>>>(df, g, sw, B, Cz, Nt, dWf, dbf);
  CNP_LAUNCH_CHECK("dec_blk_dw_kernel");
  return 0;
}

// Dense d_h (blocked bf16, every interior pixel written) = ReLU'(h) * SetConv^T(dg).
CNP_API int cnp_dec_blk_bwd_data(const float* dg, const float* xt, int B, int Nt, double start1, double start2, double res,
                                 float scale2, const cnp_blk* h, const cnp_blk* dh, cudaStream_t st) {
  CNP_REQUIRE(dg && xt && h && dh && B > 0 && h->H == dh->H && h->W == dh->W, "dec_blk_bwd_data: bad arguments");
  const size_t plane = (size_t)(h->H + 4) * (h->W + 4) * 8;
  const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(h->base) + (size_t)h->cb_off * plane;
  __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dh->base) + (size_t)dh->cb_off * plane;
  dim3 grid(cnp_cdiv(h->W, DT_J), cnp_cdiv(h->H, DT_I), B);
  dec_blk_bwd_kernel<<<grid, 256, 0, st>>>(dg, xt, Nt, start1, start2, res, scale2, hp, h->bstride, dp, dh->bstride, h->H,
                                           h->W);
  CNP_LAUNCH_CHECK("dec_blk_bwd_kernel");
  return 0;
}
