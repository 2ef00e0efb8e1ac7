// Fused aux-at-target MLP + likelihood head + normalised NLL, forward and backward.
//
// Replaces upstream neuralprocesses Augment -> MLP -> likelihood -> logpdf -> nps.loglik(normalise=True)
// (U11) as reached from ConvNP.loss_fn (nzdownscale/downscaler/train.py:370) and train_epoch
// (train.py:388-394), for the likelihoods nzdownscale/dataprocess/config.py:162-169 selects per variable:
//   Gaussian ('cnp'):            o = MLP([f ; aux_t]);  mean = o0;  var = 1e-6 + softplus(o1);
//                                logp_b = -1/2 sum_t [log 2pi + log var + (y-mean)^2/var]
//   Bernoulli-Gamma (precip.):   k = 1e-6 + softplus(o0), scale = 1e-6 + softplus(o1), (l_zero, l_slab) = (o2, o3);
//                                log p(y) = log_softmax(l)_zero            if y == 0
//                                         = log_softmax(l)_slab + log Gamma(y; k, scale)   otherwise
//   spikes-Beta (humidity):      alpha = 1e-6 + softplus(o0), beta = 1e-6 + softplus(o1), (l_0, l_1, l_slab) = o2..4;
//                                log p(y) = log_softmax(l)_0 | _1          if y == 0 | y == 1
//                                         = log_softmax(l)_slab + log Beta(clamp(y, eps, 1-eps); alpha, beta)   otherwise
//   (upstream neuralprocesses SpikesSlab: spikes first, slab last -- [U], SURVEY 8(c): restated from memory)
// All log-pdfs in float64, NaN targets skipped.  The backward recomputes the (tiny) forward instead of saving activations.
#include "common.cuh"
#include <math.h>

#include "mlp_params.cuh"

namespace {

constexpr int MAXW = 160;  // max layer width (incl. input)

struct Offsets {
  int w[CNP_MLP_MAX_LAYERS];  // offset of padded W_l  ([out][in+1]) in the weight arena
  int b[CNP_MLP_MAX_LAYERS];
  int total;
};

__host__ __device__ inline Offsets make_offsets(const cnp_mlp_params& p) {
  Offsets o;
  int cur = 0;
  for (int l = 0; l < p.n_layers; ++l) {
    o.w[l] = cur; cur += p.dims[l + 1] * (p.dims[l] + 1);
    o.b[l] = cur; cur += p.dims[l + 1];
  }
  o.total = cur;
  return o;
}

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_t(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }

constexpr double LIK_EPS = 1e-6;

// digamma(x), x > 0: recurrence up to x >= 6, then the asymptotic series (|error| < 1e-12)
__device__ double digamma_d(double x) {
  double r = 0.0;
  while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + log(x) - 0.5 / x - f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f * (1.0 / 132.0)))));
}

__host__ __device__ inline int lik_channels(int lik) { return lik == CNP_LIK_GAUSS ? 2 : (lik == CNP_LIK_BERNOULLI_GAMMA ? 4 : 5); }

// Parameters of the spikes-and-slab likelihoods from the raw head inputs z: a = 1e-6 + softplus(z0), b = 1e-6 +
// softplus(z1) in fp32 like the model outputs, log-probabilities normalised in float64.  lp[] = spikes..., slab.
struct SpikeSlab {
  double a, b;        // Gamma: (k, scale); Beta: (alpha, beta)
  double lp[3];
  int n_spikes;
};
__device__ __forceinline__ SpikeSlab spike_slab_of(const float* z, int lik) {
  SpikeSlab s;
  s.a = (double)(1e-6f + softplus_t(z[0]));
  s.b = (double)(1e-6f + softplus_t(z[1]));
  s.n_spikes = lik == CNP_LIK_BERNOULLI_GAMMA ? 1 : 2;
  double mx = -INFINITY;
  for (int i = 0; i <= s.n_spikes; ++i) mx = fmax(mx, (double)z[2 + i]);
  double se = 0.0;
  for (int i = 0; i <= s.n_spikes; ++i) se += exp((double)z[2 + i] - mx);
  const double lse = mx + log(se);
  for (int i = 0; i <= s.n_spikes; ++i) s.lp[i] = (double)z[2 + i] - lse;
  return s;
}
// category of an observation: spike index, or n_spikes for the slab
__device__ __forceinline__ int spike_of(float y, int lik) {
  if (y == 0.f) return 0;
  if (lik == CNP_LIK_SPIKES_BETA && y == 1.f) return 1;
  return lik == CNP_LIK_BERNOULLI_GAMMA ? 1 : 2;
}
__device__ __forceinline__ double slab_logpdf(const SpikeSlab& s, float y, int lik) {
  if (lik == CNP_LIK_BERNOULLI_GAMMA) {
    const double x = (double)y;
    return (s.a - 1.0) * log(x) - x / s.b - lgamma(s.a) - s.a * log(s.b);
  }
  const double x = fmin(fmax((double)y, LIK_EPS), 1.0 - LIK_EPS);
  return (s.a - 1.0) * log(x) + (s.b - 1.0) * log1p(-x) - (lgamma(s.a) + lgamma(s.b) - lgamma(s.a + s.b));
}
// distribution mean and variance (what ConvNP.mean / .std / predict report)
__device__ __forceinline__ void spike_slab_moments(const SpikeSlab& s, int lik, float* mean, float* var) {
  double m1, m2;   // slab mean and second moment
  if (lik == CNP_LIK_BERNOULLI_GAMMA) { m1 = s.a * s.b; m2 = s.a * (s.a + 1.0) * s.b * s.b; }
  else { const double t = s.a + s.b; m1 = s.a / t; m2 = s.a * s.b / (t * t * (t + 1.0)) + m1 * m1; }
  const double ps = exp(s.lp[s.n_spikes]);
  double m = ps * m1, e2 = ps * m2;
  if (lik == CNP_LIK_SPIKES_BETA) { const double p1 = exp(s.lp[1]); m += p1; e2 += p1; }   // spike at 1 (spike at 0 adds 0)
  *mean = (float)m;
  *var = (float)fmax(e2 - m * m, 0.0);
}

// weights -> shared memory, rows padded to in + 1 floats.  One warp per output row, lanes along the row: coalesced, no
// integer division, and every load of a layer is independent of the others (the first version indexed e / in, e % in over
// the flat tensor: ~50 dependent-address iterations per thread, a third of the head kernels' 32 / 55 us)
__device__ void stage_weights(const cnp_mlp_params& p, const Offsets& o, float* ws) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int l = 0; l < p.n_layers; ++l) {
    const int in = p.dims[l], out = p.dims[l + 1];
    const float* __restrict__ W = p.W[l];
    for (int oo = warp; oo < out; oo += nwarp) {
      float* dst = ws + o.w[l] + oo * (in + 1);
      const float* src = W + oo * in;
      for (int i = lane; i < in; i += 32) dst[i] = __ldg(src + i);
    }
    for (int e = threadIdx.x; e < out; e += blockDim.x) ws[o.b[l] + e] = __ldg(p.b[l] + e);
  }
}

// one warp evaluates layer l for one point: hout[o] = act(b[o] + sum_i W[o][i] hin[i])
__device__ __forceinline__ void layer_fwd(const float* ws, const Offsets& o, const cnp_mlp_params& p, int l,
                                          const float* hin, float* hout, int lane, bool relu) {
  const int in = p.dims[l], out = p.dims[l + 1];
  for (int oo = lane; oo < out; oo += 32) {
    const float* wr = ws + o.w[l] + oo * (in + 1);
    float s = ws[o.b[l] + oo];
    for (int i = 0; i < in; ++i) s = fmaf(wr[i], hin[i], s);
    hout[oo] = (relu && s < 0.f) ? 0.f : s;  // NaN propagates like torch.relu
  }
}

// ---------------------------------------------------------------------------------------------
// forward: grid (chunks, B), 8 warps, PTS points per block
// ---------------------------------------------------------------------------------------------
constexpr int FW_PTS = 8;    // one target per warp per block: Nt ~ 40-200 targets x B tasks must fill 148 SMs

__global__ void __launch_bounds__(256)
mlp_head_fwd_kernel(cnp_mlp_params p, const float* __restrict__ f, int f_ctotal, int Cf,
                    const float* __restrict__ aux, int Ca, int Nt, float* __restrict__ mean, float* __restrict__ var,
                    float* __restrict__ zraw) {
  extern __shared__ __align__(16) float smem[];
  const Offsets o = make_offsets(p);
  float* ws = smem;
  float* act = smem + o.total;  // [8 warps][2][MAXW]
  stage_weights(p, o, ws);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
  float* ha = act + warp * 2 * MAXW;
  float* hb = ha + MAXW;
  const int t0 = blockIdx.x * FW_PTS;
  for (int t = t0 + warp; t < min(t0 + FW_PTS, Nt); t += 8) {
    for (int c = lane; c < Cf; c += 32) ha[c] = f[((size_t)b * f_ctotal + c) * Nt + t];
    for (int c = lane; c < Ca; c += 32) ha[Cf + c] = aux[((size_t)b * Ca + c) * Nt + t];
    __syncwarp();
    float* hin = ha; float* hout = hb;
    for (int l = 0; l < p.n_layers; ++l) {
      layer_fwd(ws, o, p, l, hin, hout, lane, l < p.n_layers - 1);
      __syncwarp();
      float* tmp = hin; hin = hout; hout = tmp;
    }
    if (lane == 0) {
      if (p.likelihood == CNP_LIK_GAUSS) {
        mean[(size_t)b * Nt + t] = hin[0];
        var[(size_t)b * Nt + t] = 1e-6f + softplus_t(hin[1]);
      } else {
        const int nz = lik_channels(p.likelihood);
        for (int c = 0; c < nz; ++c) zraw[((size_t)b * nz + c) * Nt + t] = hin[c];
        const SpikeSlab ss = spike_slab_of(hin, p.likelihood);
        spike_slab_moments(ss, p.likelihood, mean + (size_t)b * Nt + t, var + (size_t)b * Nt + t);
      }
    }
    __syncwarp();
  }
}

// logp[b] = sum_t log N(y_t; mean_t, var_t) in float64 over the non-NaN targets, count[b] = their number.  One warp per
// task, a FIXED summation order (lane-strided partial sums, then a shuffle tree): the loss is run-to-run identical
// (the first version added per-block partial sums with atomics, which reordered the float64 sum from run to run).
__global__ void __launch_bounds__(32)
head_logp_kernel(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ zraw,
                 int lik, const float* __restrict__ yt, int Nt, double* __restrict__ logp, int* __restrict__ count) {
  const int b = blockIdx.x, lane = threadIdx.x;
  double lp = 0.0;
  int cnt = 0;
  for (int t = lane; t < Nt; t += 32) {
    const float yv = yt[(size_t)b * Nt + t];
    if (!isnan(yv)) {
      if (lik == CNP_LIK_GAUSS) {
        const double dm = (double)yv - (double)mean[(size_t)b * Nt + t], dv = (double)var[(size_t)b * Nt + t];
        lp += -0.5 * (1.8378770664093453 + log(dv) + dm * dm / dv);
      } else {
        const int nz = lik_channels(lik);
        float z[5];
        for (int c = 0; c < nz; ++c) z[c] = zraw[((size_t)b * nz + c) * Nt + t];
        const SpikeSlab ss = spike_slab_of(z, lik);
        const int cat = spike_of(yv, lik);
        lp += ss.lp[cat] + (cat == ss.n_spikes ? slab_logpdf(ss, yv, lik) : 0.0);
      }
      cnt += 1;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lp += __shfl_xor_sync(0xffffffffu, lp, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) { logp[b] = lp; count[b] = cnt; }
}

// loss = -mean_b(logp_b / max(N_b, 1)) (or -mean_b logp_b) in float64, summed in task order, and the seed of the backward
// dlogp_b = d loss / d logp_b = -1 / (B max(N_b, 1)) as fp32 -- one launch instead of the dozen one-element torch kernels
// (clamp, casts, div, mean, neg; and their mirror images in the backward) that used to sit between the head and the backward.
__global__ void __launch_bounds__(32)
loss_mean_kernel(const double* __restrict__ logp, const int* __restrict__ count, int B, int normalise,
                 double* __restrict__ loss, float* __restrict__ dlogp) {
  if (threadIdx.x != 0) return;
  double s = 0.0;
  for (int b = 0; b < B; ++b) {
    const double den = normalise ? (double)max(count[b], 1) : 1.0;
    s += logp[b] / den;
    dlogp[b] = (float)(-1.0 / ((double)B * den));
  }
  *loss = -s / (double)B;
}

// ---------------------------------------------------------------------------------------------
// backward: grid (chunks, B), BW_PTS points per block
//   dlogp[b] = d loss / d logp_b (already includes -1/(B*N_b)); outputs df (first Cf channels),
//   dW_l, db_l (+= via atomics once per block)
// ---------------------------------------------------------------------------------------------
constexpr int BW_PTS = 8;

__global__ void __launch_bounds__(256)
mlp_head_bwd_kernel(cnp_mlp_params p, const float* __restrict__ f, int f_ctotal, int Cf,
                    const float* __restrict__ aux, int Ca, const float* __restrict__ yt, int Nt,
                    const float* __restrict__ dlogp, float* __restrict__ df, float* __restrict__ gws, int ws_stride) {
  extern __shared__ __align__(16) float smem[];
  const Offsets o = make_offsets(p);
  const int L = p.n_layers;
  float* ws = smem;                       // weights
  float* acts = ws + o.total;             // per layer input: sum_l dims[l] per point
  int aoff[CNP_MLP_MAX_LAYERS + 1], goff[CNP_MLP_MAX_LAYERS + 1];
  int asum = 0, gsum = 0;
  for (int l = 0; l < L; ++l) { aoff[l] = asum; asum += p.dims[l]; goff[l] = gsum; gsum += p.dims[l + 1]; }
  float* dpre = acts + asum * BW_PTS;     // per layer pre-activation grads
  float* scratch = dpre + gsum * BW_PTS;  // [8 warps][MAXW]
  stage_weights(p, o, ws);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
  const int t0 = blockIdx.x * BW_PTS;
  const int npts = min(BW_PTS, Nt - t0);
  const float gl = dlogp[b];
  float* tmp = scratch + warp * MAXW;

  for (int pt = warp; pt < npts; pt += 8) {
    const int t = t0 + pt;
    float* a0 = acts + aoff[0] * BW_PTS + pt * p.dims[0];
    for (int c = lane; c < Cf; c += 32) a0[c] = f[((size_t)b * f_ctotal + c) * Nt + t];
    for (int c = lane; c < Ca; c += 32) a0[Cf + c] = aux[((size_t)b * Ca + c) * Nt + t];
    __syncwarp();
    // forward, keeping every layer input
    for (int l = 0; l < L; ++l) {
      const float* hin = acts + aoff[l] * BW_PTS + pt * p.dims[l];
      float* hout = (l + 1 < L) ? acts + aoff[l + 1] * BW_PTS + pt * p.dims[l + 1] : tmp;
      layer_fwd(ws, o, p, l, hin, hout, lane, l < L - 1);
      __syncwarp();
    }
    // head gradient
    float* gL = dpre + goff[L - 1] * BW_PTS + pt * p.dims[L];
    if (lane == 0) {
      const float yv = yt[(size_t)b * Nt + t];
      if (p.likelihood == CNP_LIK_GAUSS) {
        const float m = tmp[0], z1 = tmp[1], v = 1e-6f + softplus_t(z1);
        float dm = 0.f, dz = 0.f;
        if (!isnan(yv)) {
          const float r = yv - m;
          dm = gl * (r / v);
          const float dvv = gl * (-0.5f) * (1.f / v - r * r / (v * v));
          dz = dvv * sigmoid_t(z1);
        }
        gL[0] = dm; gL[1] = dz;
      } else {
        const int lik = p.likelihood, nz = lik_channels(lik);
        for (int c = 0; c < nz; ++c) gL[c] = 0.f;
        if (!isnan(yv)) {
          const SpikeSlab ss = spike_slab_of(tmp, lik);
          const int cat = spike_of(yv, lik);
          // d log p / d l_j = [j == cat] - softmax(l)_j
          for (int i = 0; i <= ss.n_spikes; ++i) gL[2 + i] = gl * (float)((i == cat ? 1.0 : 0.0) - exp(ss.lp[i]));
          if (cat == ss.n_spikes) {
            double da, db;      // d slab_logpdf / d (a, b)
            if (lik == CNP_LIK_BERNOULLI_GAMMA) {
              const double x = (double)yv;
              da = log(x) - digamma_d(ss.a) - log(ss.b);
              db = x / (ss.b * ss.b) - ss.a / ss.b;
            } else {
              const double x = fmin(fmax((double)yv, LIK_EPS), 1.0 - LIK_EPS), dg = digamma_d(ss.a + ss.b);
              da = log(x) - digamma_d(ss.a) + dg;
              db = log1p(-x) - digamma_d(ss.b) + dg;
            }
            gL[0] = gl * (float)da * sigmoid_t(tmp[0]);
            gL[1] = gl * (float)db * sigmoid_t(tmp[1]);
          }
        }
      }
    }
    __syncwarp();
    // backward through the layers
    for (int l = L - 1; l >= 0; --l) {
      const int in = p.dims[l], out = p.dims[l + 1];
      const float* g = dpre + goff[l] * BW_PTS + pt * out;
      const float* hin = acts + aoff[l] * BW_PTS + pt * in;
      for (int i = lane; i < in; i += 32) {
        float s = 0.f;
        for (int oo = 0; oo < out; ++oo) s = fmaf(ws[o.w[l] + oo * (in + 1) + i], g[oo], s);
        if (l > 0) {
          dpre[goff[l - 1] * BW_PTS + pt * in + i] = (hin[i] > 0.f) ? s : 0.f;
        } else if (i < Cf) {
          df[((size_t)b * f_ctotal + i) * Nt + t] = s;
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // weight / bias gradients for this block's points: with a workspace every block stores its partial sums
  // ([block][W_0, b_0, W_1, b_1, ...], plain layout) and mlp_head_reduce_kernel adds the blocks in a fixed order
  // (run-to-run identical gradients); without one they meet in dW / db through atomics
  float* part = gws ? gws + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * ws_stride : nullptr;
  int poff = 0;
  for (int l = 0; l < L; ++l) {
    const int in = p.dims[l], out = p.dims[l + 1];
    const float* A = acts + aoff[l] * BW_PTS;
    const float* G = dpre + goff[l] * BW_PTS;
    for (int e = threadIdx.x; e < in * out; e += blockDim.x) {
      const int oo = e / in, i = e % in;
      float s = 0.f;
      for (int pt = 0; pt < npts; ++pt) s = fmaf(G[pt * out + oo], A[pt * in + i], s);
      if (part) part[poff + e] = s; else atomicAdd(p.dW[l] + e, s);
    }
    poff += in * out;
    for (int oo = threadIdx.x; oo < out; oo += blockDim.x) {
      float s = 0.f;
      for (int pt = 0; pt < npts; ++pt) s += G[pt * out + oo];
      if (part) part[poff + oo] = s; else atomicAdd(p.db[l] + oo, s);
    }
    poff += out;
  }
}

// dW_l / db_l += sum over the blocks of mlp_head_bwd_kernel, block 0 first: one thread per parameter
__global__ void __launch_bounds__(256)
mlp_head_reduce_kernel(cnp_mlp_params p, const float* __restrict__ ws, int ws_stride, int nblk) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= ws_stride) return;
  float s = 0.f;
  for (int k = 0; k < nblk; ++k) s += __ldg(ws + (size_t)k * ws_stride + e);
  int off = e;
  for (int l = 0; l < p.n_layers; ++l) {
    const int nw = p.dims[l] * p.dims[l + 1], nb = p.dims[l + 1];
    if (off < nw) { p.dW[l][off] += s; return; }
    off -= nw;
    if (off < nb) { p.db[l][off] += s; return; }
    off -= nb;
  }
}

int check_params(const cnp_mlp_params* p, int Cf, int Ca) {
  CNP_REQUIRE(p && p->n_layers >= 1 && p->n_layers <= CNP_MLP_MAX_LAYERS, "mlp_head: 1..%d layers", CNP_MLP_MAX_LAYERS);
  CNP_REQUIRE(p->dims[0] == Cf + Ca, "mlp_head: dims[0]=%d != Cf+Ca=%d", p->dims[0], Cf + Ca);
  CNP_REQUIRE(p->likelihood >= 0 && p->likelihood <= 2, "mlp_head: unknown likelihood %d", p->likelihood);
  CNP_REQUIRE(p->dims[p->n_layers] == lik_channels(p->likelihood), "mlp_head: likelihood %d needs %d head inputs, the MLP has %d",
              p->likelihood, lik_channels(p->likelihood), p->dims[p->n_layers]);
  for (int l = 0; l <= p->n_layers; ++l)
    CNP_REQUIRE(p->dims[l] >= 1 && p->dims[l] <= MAXW, "mlp_head: layer width %d out of range (<=%d)", p->dims[l], MAXW);
  return 0;
}

}  // namespace

CNP_API int cnp_mlp_head_fwd(const cnp_mlp_params* p, const float* f, int f_ctotal, int Cf, const float* aux, int Ca,
                             const float* yt, int B, int Nt, float* mean, float* var, float* zraw, double* logp, int* count,
                             cudaStream_t st) {
  if (int e = check_params(p, Cf, Ca)) return e;
  CNP_REQUIRE(p->likelihood == CNP_LIK_GAUSS || zraw != nullptr, "mlp_head_fwd: this likelihood needs the zraw buffer");
  CNP_REQUIRE(B > 0 && Nt >= 0 && f_ctotal >= Cf, "mlp_head_fwd: bad sizes");
  CNP_REQUIRE(logp == nullptr || (yt != nullptr && count != nullptr), "mlp_head_fwd: logp needs yt and count");
  if (Nt == 0) return 0;
  const Offsets o = make_offsets(*p);
  const size_t smem = (size_t)(o.total + 8 * 2 * MAXW) * sizeof(float);
  CNP_REQUIRE(smem <= 200 * 1024, "mlp_head_fwd: MLP too large for shared memory (%zu B)", smem);
  static size_t attr_f = 0;
  if (smem > attr_f) { cudaFuncSetAttribute(mlp_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_f = smem; }
  dim3 grid(cnp_cdiv(Nt, FW_PTS), B);
  mlp_head_fwd_kernel<<<grid, 256, smem, st>>>(*p, f, f_ctotal, Cf, aux, Ca, Nt, mean, var, zraw);
  CNP_LAUNCH_CHECK("mlp_head_fwd_kernel");
  if (logp) {
    head_logp_kernel<<<B, 32, 0, st>>>(mean, var, zraw, p->likelihood, yt, Nt, logp, count);
    CNP_LAUNCH_CHECK("head_logp_kernel");
  }
  return 0;
}

// Bytes of the optional partial-sum workspace of cnp_mlp_head_bwd (one slice of all MLP gradients per block).
CNP_API long long cnp_mlp_head_bwd_workspace_bytes(const cnp_mlp_params* p, int B, int Nt) {
  if (!p || p->n_layers < 1 || p->n_layers > CNP_MLP_MAX_LAYERS) return 0;
  long long per = 0;
  for (int l = 0; l < p->n_layers; ++l) per += (long long)p->dims[l] * p->dims[l + 1] + p->dims[l + 1];
  return per * cnp_cdiv(Nt, BW_PTS) * B * (long long)sizeof(float);
}

// workspace (or NULL): with it the parameter gradients are reduced over the blocks in a fixed order (run-to-run
// identical); without it they meet in dW / db through fp32 atomics.
CNP_API int cnp_mlp_head_bwd(const cnp_mlp_params* p, const float* f, int f_ctotal, int Cf, const float* aux, int Ca,
                             const float* yt, int B, int Nt, const float* dlogp, float* df, void* workspace,
                             long long workspace_bytes, cudaStream_t st) {
  if (int e = check_params(p, Cf, Ca)) return e;
  CNP_REQUIRE(B > 0 && Nt >= 0 && f_ctotal >= Cf && yt && dlogp && df, "mlp_head_bwd: bad arguments");
  for (int l = 0; l < p->n_layers; ++l) CNP_REQUIRE(p->dW[l] && p->db[l], "mlp_head_bwd: missing gradient buffers");
  if (Nt == 0) return 0;
  const Offsets o = make_offsets(*p);
  int asum = 0, gsum = 0;
  for (int l = 0; l < p->n_layers; ++l) { asum += p->dims[l]; gsum += p->dims[l + 1]; }
  const size_t smem = (size_t)(o.total + (asum + gsum) * BW_PTS + 8 * MAXW) * sizeof(float);
  CNP_REQUIRE(smem <= 220 * 1024, "mlp_head_bwd: MLP too large for shared memory (%zu B)", smem);
  static size_t attr_b = 0;
  if (smem > attr_b) { cudaFuncSetAttribute(mlp_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_b = smem; }
  dim3 grid(cnp_cdiv(Nt, BW_PTS), B);
  const long long need = cnp_mlp_head_bwd_workspace_bytes(p, B, Nt);
  float* ws = (workspace && workspace_bytes >= need) ? reinterpret_cast<float*>(workspace) : nullptr;
  const int nblk = (int)(grid.x * grid.y), stride = (int)(need / sizeof(float) / nblk);
  mlp_head_bwd_kernel<<<grid, 256, smem, st>>>(*p, f, f_ctotal, Cf, aux, Ca, yt, Nt, dlogp, df, ws, stride);
  CNP_LAUNCH_CHECK("mlp_head_bwd_kernel");
  if (ws) {
    mlp_head_reduce_kernel<<<cnp_cdiv(stride, 256), 256, 0, st>>>(*p, ws, stride, nblk);
    CNP_LAUNCH_CHECK("mlp_head_reduce_kernel");
  }
  return 0;
}

// loss (float64 scalar) = -mean_b(logp_b [/ max(count_b, 1)]); dlogp [B] fp32 = d loss / d logp.
CNP_API int cnp_loss_mean(const double* logp, const int* count, int B, int normalise, double* loss, float* dlogp,
                          cudaStream_t st) {
  CNP_REQUIRE(logp && count && loss && dlogp && B > 0, "loss_mean: bad arguments");
  loss_mean_kernel<<<1, 32, 0, st>>>(logp, count, B, normalise, loss, dlogp);
  CNP_LAUNCH_CHECK("loss_mean_kernel");
  return 0;
}
