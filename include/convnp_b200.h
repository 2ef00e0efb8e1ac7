/* convnp_b200.h -- C ABI of libconvnp_b200.so (sm_100a only).
 *
 * The reference (oriordanemily/deepsensorNZ) has no FFI layer: its hot path is reached through the
 * DeepSensor Python API and executed by torch ops inside
 * deepsensor==0.3.6 / neuralprocesses==0.2.6.  Each entry point below replaces the torch ops of one
 * stage of that path; the "replaces" line cites the reference call site that reaches the stage and
 * the upstream module that implements it today.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.  All pointers are DEVICE pointers unless
 *     stated otherwise; the caller owns every buffer, the callee never allocates, frees or synchronises.
 *   - every function returns int: 0 = ok, <0 = bad argument (see cnp_last_error()), >0 = cudaError_t.
 *   - work is enqueued on the cudaStream_t passed as the last argument.
 *   - fp32 activations are NCHW with an explicit batch stride (in elements) so that channel windows of a
 *     larger buffer (skip concatenations) can be addressed without copies.
 *   - bf16 activations use the "blocked" layout [B][C/8][H+4][W+4][8]: channel chunks of 8 (16 B per
 *     pixel-chunk), each plane physically zero-padded by 2 pixels; described by cnp_blk.  The pad must be
 *     zero and is never written by the library.  Allocate (16*(W+4)+512)*8 elements of zeroed slack after
 *     the last plane (tile over-reads).
 *   - the internal grid is (start1, n1, start2, n2, res): point i along a dimension is
 *     (float)(start + i*res), computed in double.
 */
#ifndef CONVNP_B200_H
#define CONVNP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* cnp_stream_t; /* == cudaStream_t */

/* ---- library ------------------------------------------------------------------------------- */
int cnp_version(void);
const char* cnp_last_error(void);          /* thread-local message of the last failing call */
int cnp_check_device(void);                /* 0 iff the current device is compute capability 10.x */

/* ---- (1) SetConv encoder -------------------------------------------------------------------
 * replaces: neuralprocesses PrependDensityChannel + SetConv + DivideByFirstChannel (coders/setconv/*.py),
 * reached from ConvNP.loss_fn / __call__ / predict -- nzdownscale/downscaler/train.py:370,:388-394,
 * validate_ERA.py:88-92.  Writes channels [ch_off, ch_off+1+C) of out [B,c_total,n1,n2] (fp32 NCHW):
 * density first, then data/(density+eps).  y may contain NaN: a point/cell with any NaN channel is
 * treated as missing (mask 0), exactly like Task.mask_nans_numpy/mask_nans_nps on the host.
 * scale2 = exp(2*log_scale). */
int cnp_setconv_enc_offgrid_fwd(const float* x /*[B,2,N]*/, const float* y /*[B,C,N]*/,
                                const float* mask /*[B,1,N] or NULL*/, int B, int C, int N,
                                double start1, int n1, double start2, int n2, double res,
                                float scale2, float eps, float* out, int ch_off, int c_total, cnp_stream_t s);
long long cnp_setconv_enc_grid_workspace_bytes(int B, int C, int N1, int n1, int n2, int band);
int cnp_setconv_enc_grid_fwd(const float* x1 /*[B or 1,N1]*/, const float* x2 /*[B or 1,N2]*/, int x_batched,
                             const float* y /*[B,C,N1,N2]*/, const float* mask /*[B,1,N1,N2] or NULL*/,
                             int B, int C, int N1, int N2, int mono1 /*+1 asc,-1 desc,0 unsorted*/, int mono2,
                             double start1, int n1, double start2, int n2, double res,
                             float scale2, float eps, float* out, int ch_off, int c_total,
                             int band /*0 = generic kernel*/, void* workspace, long long workspace_bytes,
                             cnp_stream_t s);

/* ---- (2) UNet, fp32 parity path ------------------------------------------------------------
 * replaces: torch Conv2d / Upsample(bilinear) / ReLU and their autograd inside neuralprocesses
 * coders/nn.py UNet, reached from train_epoch (train.py:388-394).  Weights are torch layout
 * [Cout][Cin][k][k]; supported (k,stride): (5,1) (5,2) (1,1); padding k/2. */
int cnp_conv2d_fwd_f32(const float* x, long long x_bstride, const float* w, const float* bias, float* y,
                       long long y_bstride, int B, int Cin, int Hin, int Win, int Cout, int k, int stride,
                       int relu, cnp_stream_t s);
int cnp_conv2d_dgrad_f32(const float* dy, long long dy_bstride, const float* w, float* dx, long long dx_bstride,
                         int B, int Cin, int Hin, int Win, int Cout, int k, int stride, int accumulate,
                         cnp_stream_t s);
int cnp_conv2d_wgrad_f32(const float* x, long long x_bstride, const float* dy, long long dy_bstride,
                         float* dw /*+=*/, float* dbias /*+= or NULL*/, int B, int Cin, int Hin, int Win, int Cout,
                         int k, int stride, cnp_stream_t s);
int cnp_relu_bwd_f32(float* dy /*in place*/, long long dy_bstride, const float* y, long long y_bstride, int B,
                     long long per_batch, cnp_stream_t s);
int cnp_upsample2x_fwd_f32(const float* x, long long x_bstride, float* y, long long y_bstride, int B, int C, int H,
                           int W, cnp_stream_t s);
int cnp_upsample2x_bwd_f32(const float* dy, long long dy_bstride, float* dx, long long dx_bstride, int B, int C,
                           int H, int W, int accumulate, cnp_stream_t s);

/* ---- (2) UNet, bf16 tensor-core path (tcgen05, TMEM accumulators, bulk-copy staging) ------------- */
typedef struct cnp_blk {
  void* base;          /* first element of batch 0, chunk 0 */
  long long bstride;   /* elements between batches */
  int cb_off;          /* first channel chunk of the view */
  int H, W;            /* interior size; planes are (H+4) x (W+4) */
} cnp_blk;

/* ---- (1') fused SetConv encoder: every context set of a task -> the UNet input ---------------------------
 * replaces: the whole upstream encoder stack (PrependDensityChannel + SetConv per set + DivideByFirstChannel +
 * Concatenate, neuralprocesses coders/setconv) as reached from ConvNP.loss_fn / predict --
 * nzdownscale/downscaler/train.py:370, validate_ERA.py:88-92.
 *   cnp_encode_tables : band tables of one gridded set; they depend only on (coordinates, internal grid, length scale)
 *                       -- build once, keep across steps.
 *   cnp_encode_hpass  : horizontal band pass of every gridded set of ``sets`` into its T workspace (one thread per
 *                       (task, input row, grid column), all channels).
 *   cnp_encode_vpass  : vertical band pass + density normalisation, T -> V fp32 planes (one thread per (task, grid row,
 *                       grid column)); a field every task shares (batched = 0: static topography / land mask) is done once.
 *   cnp_encode_fused  : one CTA = one 16 x 32 tile of one task: off-grid sets, then assembly of the output tile.
 * mode 0: fp32 NCHW out_f32 [B][c_total][n1][n2]: the gridded sets' V planes point INTO out_f32 (vpass writes the final
 * channels), cnp_encode_fused adds the off-grid channels.  mode 1: blocked bf16 out_blk with n_chunks chunks: channels
 * [0, c_total) = encoder output (gridded channels gathered from the V planes), channel c_total = 1 inside the image
 * (folded first layer, cnp_fold_in_fwd), rest 0.  Gridded sets need monotone coordinates shared by the batch. */
typedef struct cnp_enc_set {
  int kind;            /* 0 off-grid, 1 gridded */
  int C;               /* data channels (1..8) */
  int ch_off;          /* first output channel (density for kinds 0 / 1) */
  int batched;         /* gridded: y / mask / T carry a batch axis (0 = one field for every task) */
  const float* x1;     /* off-grid x [B,2,N] */
  const float* x2;     /* unused */
  const float* y;      /* gridded [B or 1,C,N1,N2]; off-grid [B,C,N]; may hold NaN (= missing) */
  const float* mask;   /* gridded [B or 1,1,N1,N2]; off-grid [B,1,N]; or NULL */
  int N1, N2;          /* off-grid: N1 = N */
  int mono1, mono2;    /* +1 ascending, -1 descending */
  float scale2;        /* exp(2 log_scale) */
  int KB;              /* gridded: band width of the tables (<= 32) */
  const int* tab_i;    /* gridded: [p0 (n1) | len1 (n1) | q0 (n2) | len2 (n2)], from cnp_encode_tables */
  const float* tab_w;  /* gridded: [w1 (KB x n1) | w2 (KB x n2)] */
  float* T;            /* gridded: horizontal-pass workspace [B or 1][C+1][N1][n2] */
  float* V;            /* gridded: normalised output planes, channel stride n1*n2 floats */
  long long V_bs;      /*          batch stride of V in floats */
} cnp_enc_set;
typedef struct cnp_enc_sets { int n_sets; int pad_; cnp_enc_set s[8]; } cnp_enc_sets;   /* HOST struct, passed by value to the kernels */
/* tab_i: 2 * (n1 + n2) ints, tab_w: band * (n1 + n2) floats; band = host-side upper bound (<= 32) of the number of inputs
 * within the truncation radius of any grid point along either dimension. */
int cnp_encode_tables(const float* x1 /*[N1]*/, const float* x2 /*[N2]*/, int N1, int N2, int mono1, int mono2,
                      double start1, int n1, double start2, int n2, double res, float scale2, int band,
                      int* tab_i, float* tab_w, cnp_stream_t s);
int cnp_encode_hpass(const cnp_enc_sets* sets, int B, int n1, int n2, cnp_stream_t s);
int cnp_encode_vpass(const cnp_enc_sets* sets, int B, int n1, int n2, float eps, cnp_stream_t s);
int cnp_encode_fused(const cnp_enc_sets* sets, int B, double start1, int n1, double start2, int n2, double res,
                     float eps, int mode, float* out_f32, long long out_bstride, int c_total, const cnp_blk* out_blk,
                     int n_chunks, cnp_stream_t s);

typedef struct cnp_conv_out {
  int mode;            /* 0: blocked bf16 (blk), 1: fp32 NCHW (f32, f32_bstride, f32_ch_off) */
  cnp_blk blk;
  float* f32; long long f32_bstride; int f32_ch_off;
  int sy, ay, sx, ax;  /* output pixel = (y*sy+ay, x*sx+ax); 1,0,1,0 for a plain conv */
  const float* bias;   /* [64] or NULL */
  int relu;
  const cnp_blk* mask; /* zero the result where mask <= 0 (ReLU backward), or NULL */
  int accumulate;      /* out += result */
  const cnp_blk* s2d;  /* cnp_conv_tc2 only: also write the space-to-depth copy (32 chunks, half size) or NULL */
  int s2d_c0;          /* first of the 8 output chunks copied there: 0, or 8 with n_out = 128 */
  int s2d_band;        /* half-res pixels along every edge left unwritten (zero): the band of the polyphase backward */
} cnp_conv_out;

enum { CNP_K5S1 = 0, CNP_K1 = 1, CNP_K5S2 = 2, CNP_K5S1_DGRAD = 3, CNP_K1_DGRAD = 4, CNP_K5S2_DGRAD = 5,
       CNP_UP_PHASE = 6 /* row phase py of conv5x5(bilinear_up2x(x)) as a 4x4 convolution of the low-res x (whatever its
                           2-pixel ring holds is the padding: zeros in the engine, exact 2 low-res pixels off the border);
                           both x-phases per call; weights from cnp_up_phase_weights()[py], packed with k = 4; output
                           sy = sx = 2, ay = py, ax = 0 */,
       CNP_UP_PHASE_DGRAD = 7 /* its input gradient at LOW resolution: source = the 32-chunk space-to-depth copy of dY
                           (cnp_up_dy_split), weights = all four phase tensors of cnp_up_phase_weights() packed with
                           k = 4, n_out = 128 */ };
enum { CNP_WG_K5S1 = 0, CNP_WG_K1 = 1, CNP_WG_K5S2 = 2, CNP_WG_K5S1_NARROW = 3 /* 1..8 source chunks */,
       CNP_WG_UP_PHASE = 4 /* x = low-res input (16 chunks), dy = space-to-depth dY (32 chunks);
                              dw += [2][2][64][128][4][4] phase gradients (fold back with cnp_up_wgrad_fold) */,
       CNP_WG_K5S1_T = 5  /* CNP_WG_K5S1 with the gradient stored tap-transposed: dw[co][ci][kx][ky] (column strips) */ };

#ifdef CNP_LEGACY_CONV_TC   /* first formulation (pixels = M operand), only built with `make LEGACY=1` */
long long cnp_conv_tc_packed_bytes(int kind, int n_chunks);
int cnp_conv_tc_pack(const float* w, int Cout, int Cin, int k, int kind, int n_chunks, int py, int px, int co_off,
                     void* wpk, cnp_stream_t s);
/* 64 output channels per call; x chunks [x->cb_off, +n_chunks) are the reduction dimension.
 * CNP_K5S2 reads the 32-chunk space-to-depth tensor; CNP_K5S2_DGRAD produces output phase (py,px). */
int cnp_conv_tc(const cnp_blk* x, int n_chunks, const void* wpk, int kind, int py, int px, const cnp_conv_out* out,
                int B, cnp_stream_t s);
#endif
/* Second formulation (weights = M operand, pixels = N operand; runs at the tcgen05 floor, see conv_tc2.cu).
 * n_out = 64: two output rows share one MMA (PAIR); n_out = 128: 128 output channels per call (WIDE, used for
 * the input gradient of the 128->64 layers).  n_chunks: 2, 4, .. 16 source chunks for the 5x5 kinds (2 = the folded
 * first layer).  CNP_K5S2_DGRAD computes one output phase (py, px) of the stride-2 input gradient per call
 * (output sy = sx = 2, ay = py, ax = px); px = 2 computes BOTH x-phases of row phase py in one call (lane group =
 * x-phase, adjacent output pixels: output ax = 0), which is what the engine uses. */
int cnp_conv_tc2_debug(long long* device_buf /*[148][8] or NULL*/, int flags);   /* per-CTA stall counters, profiling aid */
int cnp_conv_tc2_set_cluster(int cluster /*1 | 2: CTA pairs multicast the weight stream*/);
long long cnp_conv_tc2_packed_bytes(int kind, int n_chunks, int n_out);
int cnp_conv_tc2_pack(const float* w, int Cout, int Cin, int k, int kind, int n_chunks, int py, int px, int co_off,
                      int n_out, void* wpk, cnp_stream_t s);
int cnp_conv_tc2(const cnp_blk* x, int n_chunks, const void* wpk, int kind, int py, int px, int n_out,
                 const cnp_conv_out* out, int B, cnp_stream_t s);
/* the same launch with a second packed weight tensor for the images b >= w2_from_b (row strips + tap-transposed column
 * strips of a square level in one launch) */
int cnp_conv_tc2_w2(const cnp_blk* x, int n_chunks, const void* wpk, const void* wpk2, int w2_from_b, int kind, int py,
                    int px, int n_out, const cnp_conv_out* out, int B, cnp_stream_t s);
/* workspace (optional, cnp_conv_tc_wgrad_workspace_bytes()): the K-split partial sums are written there and folded
 * by a second kernel; without it they are reduced with fp32 atomics straight into dw. */
long long cnp_conv_tc_wgrad_workspace_bytes(void);
int cnp_conv_tc_wgrad(const cnp_blk* x, int n_chunks, const cnp_blk* dy, int kind, float* dw /*+= [64][Cin][k][k]*/,
                      float* dbias /*+= [64] or NULL*/, int Cin, int B, void* workspace, long long workspace_bytes,
                      cnp_stream_t s);
/* CNP_WG_K5S1 over two image groups in one launch: images [0, b_split) -> dw, [b_split, B) -> dw2 (both +=) */
int cnp_conv_tc_wgrad_pair(const cnp_blk* x, int n_chunks, const cnp_blk* dy, float* dw, float* dw2, int b_split,
                           float* dbias /*+= or NULL*/, int Cin, int B, void* workspace, long long workspace_bytes,
                           cnp_stream_t s);
int cnp_blk_channel_sum(const cnp_blk* v, int n_chunks, int B, float* out /*+=*/, cnp_stream_t s);
int cnp_conv1x1_in_bf16(const float* x /*fp32 NCHW*/, long long x_bstride, const float* w, const float* bias, int B,
                        int Cin, int Cout, const cnp_blk* out, cnp_stream_t s);
int cnp_conv1x1_in_wgrad(const float* x, long long x_bstride, int Cin, const cnp_blk* dy, int B, float* dw,
                         float* dbias, cnp_stream_t s);
int cnp_blk_upsample2x_fwd(const cnp_blk* x, int n_chunks, const cnp_blk* y, int B, cnp_stream_t s);
int cnp_blk_upsample2x_bwd(const cnp_blk* dy, int n_chunks, const cnp_blk* dx, const cnp_blk* act /*or NULL*/,
                           int accumulate, int B, cnp_stream_t s);
int cnp_blk_space_to_depth(const cnp_blk* x, int n_chunks, const cnp_blk* y /*32 chunks, half size*/, int B,
                           cnp_stream_t s);
int cnp_blk_from_nchw_f32(const float* src, long long src_bstride, int B, int C, int H, int W, const cnp_blk* dst,
                          cnp_stream_t s);
int cnp_blk_to_nchw_f32(const cnp_blk* src, int B, int C, float* dst, long long dst_bstride, cnp_stream_t s);
/* conversion into n_chunks chunks with channel C := 1 inside the image (input of the folded first layer) */
int cnp_blk_from_nchw_f32_ones(const float* src, long long src_bstride, int B, int C, int H, int W, const cnp_blk* dst,
                               int n_chunks, unsigned long long shared_mask /* bit c: channel c is read from batch 0 */,
                               cnp_stream_t s);
/* initial 1x1 (neuralprocesses UNet.initial_linear) folded into the first 5x5 (before_turn_layers[0]):
 * wf [Cout][Cp][k*k] = W5 . [W1 | b1] (channels Cin+1..Cp-1 zero); bwd maps the folded gradient dwf back (+=). */
/* wp [2 row phase][2 x-phase][Cout][Cin][4][4] <- w5 [Cout][Cin][5][5]: polyphase weights of Upsample(x2, bilinear) + Conv 5x5 */
int cnp_up_phase_weights(const float* w5, int Cout, int Cin, float* wp, cnp_stream_t s);
/* Polyphase resize-convolution, band handling (up_poly.cu; tools/polyphase_strips.py states the decomposition):
 * replaces torch.nn.Upsample(x2, bilinear) + Conv2d(5, padding=2) of neuralprocesses' UNet decoder levels.
 * Strip tensors are blocked images [2B][chunks][6+4][L+4][8]: image s*B + b, s = 0 low edge / 1 high edge;
 * rows: L = 2W high-res columns; cols: L = 2H, strip row = high-res column (transposed).
 *   cnp_up_strips_fwd     : strips of bilinear_up2x(x)
 *   cnp_up_strips_scatter : band outputs of the strip convolutions -> the high-res destination (8 chunks)
 *   cnp_up_dy_split       : dY -> space-to-depth copy with the band zeroed (32 chunks, low res; NULL when the producer's
 *                           epilogue wrote it: cnp_conv_out.s2d / s2d_c0 / s2d_band) + band strips
 *   cnp_up_strips_bwd_fold: dx += mask(act > 0) * up2x^T(strip gradients w.r.t. the upsampled tensor)
 *   cnp_up_wgrad_fold     : dw5 += fold^T(phase gradients) + tap-transposed column-strip gradient */
int cnp_up_strips_fwd(const cnp_blk* x, int n_chunks, const cnp_blk* rows, const cnp_blk* cols, int B, cnp_stream_t s);
int cnp_up_strips_scatter(const cnp_blk* rows, const cnp_blk* cols, const cnp_blk* dst, int B, cnp_stream_t s);
int cnp_up_dy_split(const cnp_blk* dy, const cnp_blk* s2d, const cnp_blk* rows, const cnp_blk* cols, int B, cnp_stream_t s);
int cnp_up_strips_bwd_fold(const cnp_blk* rows, const cnp_blk* cols, const cnp_blk* dx, const cnp_blk* act /*or NULL*/,
                           int n_chunks, int B, cnp_stream_t s);
int cnp_up_wgrad_fold(const float* dwp, const float* dwt /*or NULL*/, int Cout, int Cin, float* dw5 /*+=*/, cnp_stream_t s);
int cnp_fold_in_fwd(const float* w5, const float* w1, const float* b1, int Cout, int Cmid, int Cin, int Cp, int k,
                    float* wf, cnp_stream_t s);
int cnp_fold_in_bwd(const float* dwf, const float* w5, const float* w1, const float* b1, int Cout, int Cmid, int Cin,
                    int Cp, int k, float* dw5, float* dw1, float* db1 /*or NULL*/, cnp_stream_t s);

/* ---- (3) SetConv decoder, grid -> off-grid targets -------------------------------------------
 * replaces: neuralprocesses SetConv(scale=1/ppu) in the decoder chain + its autograd (A.5). */
int cnp_setconv_dec_offgrid_fwd(const float* z /*[B,C,n1,n2]*/, long long z_bstride, const float* xt /*[B,2,Nt]*/,
                                int B, int C, int Nt, double start1, int n1, double start2, int n2, double res,
                                float scale2, float* f /*[B,f_ctotal,Nt]*/, int f_ctotal, cnp_stream_t s);
int cnp_setconv_dec_offgrid_bwd(const float* df, int f_ctotal, const float* xt, int B, int C, int Nt,
                                double start1, int n1, double start2, int n2, double res, float scale2,
                                float* dz, long long dz_bstride, cnp_stream_t s);

/* Training path on the blocked bf16 activations (dec_blk.cu): the SetConv decoder is applied to the last hidden
 * activation h (64 ch) and the final 1x1 convolution (Wf [Cz][64], bf [Cz]) is applied to the Nt target vectors
 * afterwards -- both are linear, f = Wf g + bf sw with g = SetConv(h), sw = SetConv(1).  g [B,Nt,64], sw [B,Nt]. */
int cnp_dec_blk_fwd(const cnp_blk* h, const float* xt /*[B,2,Nt]*/, int B, int Nt, double start1, double start2,
                    double res, float scale2, const float* Wf, const float* bf, int Cz, float* g, float* sw,
                    float* f /*[B,Cz,Nt]*/, cnp_stream_t s);
int cnp_dec_blk_bwd_params(const float* df /*[B,Cz,Nt]*/, const float* g, const float* sw, const float* Wf, int B, int Nt,
                           int Cz, float* dg /*[B,Nt,64]*/, float* dWf /*+=*/, float* dbf /*+= or NULL*/, cnp_stream_t s);
int cnp_dec_blk_bwd_data(const float* dg, const float* xt, int B, int Nt, double start1, double start2, double res,
                         float scale2, const cnp_blk* h /*ReLU mask*/, const cnp_blk* dh /*out, dense*/, cnp_stream_t s);

/* on-grid targets (ConvNP.predict onto the 1400x1400 NZ grid, validate_ERA.py:88-92): separable, truncated */
long long cnp_setconv_dec_grid_workspace_bytes(int B, int C, int n1, int P, int Q);
int cnp_setconv_dec_grid_fwd(const float* z, long long z_bstride, const float* x1t /*[P]*/, const float* x2t /*[Q]*/,
                             int B, int C, int P, int Q, double start1, int n1, double start2, int n2, double res,
                             float scale2, float* f /*[B,C,P,Q]*/, long long f_bstride, void* workspace,
                             long long workspace_bytes, cnp_stream_t s);

/* Fused on-grid inference decoder on the blocked bf16 hidden activation h (64 ch): SetConv of h (column pass to a
 * bf16 workspace, row pass in registers), final 1x1 (Wf [64][64], bf) folded into MLP layer 0, MLP, Gaussian head.
 * The 64 x P x Q decoder output is never materialised.  p: dims[0] = 64 + Ca, hidden width 64, >= 2 hidden layers. */
long long cnp_decode_grid_fused_workspace_bytes(int B, int n1, int P, int Q);
struct cnp_mlp_params;
int cnp_decode_grid_fused_fwd(const cnp_blk* h, const float* x1t /*[P]*/, const float* x2t /*[Q]*/, int B, int P, int Q,
                              double start1, double start2, double res, float scale2, const float* Wf, const float* bf,
                              const struct cnp_mlp_params* p, const float* aux /*[B or 1,Ca,P,Q]*/, long long aux_bstride,
                              int Ca, float* mean /*[B,P,Q]*/, float* stdv, void* workspace, long long workspace_bytes,
                              cnp_stream_t s);

/* Tensor-core variant (same arguments): row pass to a bf16 workspace, then per 128-pixel tile the column pass and the
 * three hidden MLP layers as chained tcgen05 GEMMs (pixels = M), head in registers.  Needs 3 hidden layers of width
 * 64 and a target grid fine enough that 128 consecutive x2t span <= 64 - band internal-grid columns. */
long long cnp_decode_grid_tc_workspace_bytes(int B, int n2, int P, int Q);
int cnp_decode_grid_tc_fwd(const cnp_blk* h, const float* x1t, const float* x2t, int B, int P, int Q, double start1,
                           double start2, double res, float scale2, const float* Wf, const float* bf,
                           const struct cnp_mlp_params* p, const float* aux, long long aux_bstride, int Ca, float* mean,
                           float* stdv, void* workspace, long long workspace_bytes, cnp_stream_t s);

/* ---- (4) aux-at-target MLP + likelihood head + normalised NLL ---------------------------------------
 * replaces: neuralprocesses Augment -> MLP -> likelihood -> logpdf -> nps.loglik (A.6, A.7) reached from
 * ConvNP.loss_fn (train.py:370).  logp is float64.  Likelihoods (the ones nzdownscale/dataprocess/config.py:162-169
 * selects per variable): 0 heteroscedastic Gaussian ('cnp': mean = o0, var = 1e-6 + softplus(o1)); 1 Bernoulli-Gamma
 * ('bernoulli-gamma', precipitation: o = (k~, scale~, l_zero, l_slab)); 2 spikes-Beta ('cnp-spikes-beta', humidity:
 * o = (alpha~, beta~, l_0, l_1, l_slab)).  mean / var are the distribution's mean and variance in every case. */
#define CNP_MLP_MAX_LAYERS 6
typedef struct cnp_mlp_params {
  const float* W[CNP_MLP_MAX_LAYERS];  /* [out,in] row-major */
  const float* b[CNP_MLP_MAX_LAYERS];
  float* dW[CNP_MLP_MAX_LAYERS];       /* += (backward only) */
  float* db[CNP_MLP_MAX_LAYERS];
  int dims[CNP_MLP_MAX_LAYERS + 1];    /* dims[0] = Cf + Ca, dims[n_layers] = head inputs: 2 | 4 | 5 */
  int n_layers;
  int likelihood;                      /* 0 Gaussian, 1 Bernoulli-Gamma, 2 spikes-Beta */
} cnp_mlp_params;                      /* HOST struct holding DEVICE pointers */
int cnp_mlp_head_fwd(const cnp_mlp_params* p, const float* f, int f_ctotal, int Cf, const float* aux, int Ca,
                     const float* yt /*or NULL*/, int B, int Nt, float* mean, float* var,
                     float* zraw /*[B][dims[n_layers]][Nt] raw head inputs; needed for likelihoods 1, 2, else NULL*/,
                     double* logp /*[B] or NULL*/, int* count /*[B]*/, cnp_stream_t s);
/* inference head over npts points per batch element (mean, std = sqrt(var)); aux_bstride = 0 shares aux */
int cnp_mlp_head_points_fwd(const cnp_mlp_params* p, const float* f, long long f_bstride, int Cf, const float* aux,
                            long long aux_bstride, int Ca, int B, long long npts, float* mean, float* stdv,
                            cnp_stream_t s);
/* workspace (cnp_mlp_head_bwd_workspace_bytes) or NULL: with it dW / db are reduced over the blocks in a fixed order
 * (run-to-run identical), without it through fp32 atomics */
long long cnp_mlp_head_bwd_workspace_bytes(const cnp_mlp_params* p, int B, int Nt);
int cnp_mlp_head_bwd(const cnp_mlp_params* p, const float* f, int f_ctotal, int Cf, const float* aux, int Ca,
                     const float* yt, int B, int Nt, const float* dlogp /*[B]*/, float* df, void* workspace,
                     long long workspace_bytes, cnp_stream_t s);
/* loss (float64 scalar) = -mean_b(logp_b [/ max(count_b, 1)]) summed in task order, and dlogp [B] fp32 = d loss / d logp:
 * the scalar arithmetic of ConvNP.loss_fn(task, normalise=...) + train_epoch's batch mean (train.py:370, 388-394) */
int cnp_loss_mean(const double* logp, const int* count, int B, int normalise, double* loss, float* dlogp, cnp_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* CONVNP_B200_H */
