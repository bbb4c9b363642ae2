/*
 * b200cd.h — C ABI of the B200 (sm_100a) training-step kernels for multimodal siamese change detection.
 *
 * This is the drop-in boundary for ONE hot path of SebastianHafner/multimodal_siamese_cd: forward +
 * backward of the U-Net family in utils/networks.py and the losses in utils/loss_functions.py.
 * The reference is pure Python over torch ops; each entry point below replaces the torch op(s) cited
 * next to it (file:line relative to the reference root). The Python host layer
 * (multimodal_siamese_cd_b200/) binds these with ctypes and registers them as torch custom ops.
 *
 * Conventions
 *   - plain pointers and sizes; no torch / C++ types cross this boundary.
 *   - every pointer is a DEVICE pointer unless stated; the caller owns all memory (inputs, outputs,
 *     workspaces); the library allocates nothing in steady state (one 4-byte host-mapped error flag per
 *     device at b200cd_init).
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous on it.
 *   - activations are NHWC bf16 with an explicit element stride per pixel (`ld`, multiple of 8), so
 *     channel slices of a concatenation buffer are addressed without copies; base pointers must be
 *     16-byte aligned. Parameters, logits, targets, statistics and gradients of parameters are fp32 in
 *     the reference layouts.
 *   - return value: 0 = ok, otherwise a B200CD_ERR_* code; b200cd_last_error() gives the message
 *     (thread-local). Nothing throws.
 *   - sm_100a only: b200cd_init fails with B200CD_ERR_ARCH on anything else. There is no CPU path.
 */
#ifndef B200CD_H_
#define B200CD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CD_OK 0
#define B200CD_ERR_SHAPE 1
#define B200CD_ERR_ALIGN 2
#define B200CD_ERR_ARCH 3
#define B200CD_ERR_CUDA 4
#define B200CD_ERR_DEVICE 5 /* a kernel's pipeline timed out and trapped (see b200cd_device_status) */

#define B200CD_ABI_VERSION 2

int b200cd_abi_version(void);
const char* b200cd_last_error(void);

/* Binds the library to CUDA device `device` of the calling process: checks compute capability 10.x,
 * resolves cuTensorMapEncodeTiled, allocates the per-device error flag. Idempotent. */
int b200cd_init(int device);

/* Synchronises `stream` and returns 0, or B200CD_ERR_DEVICE with the time-out code if a tensor-core kernel waited
 * ~10 s on an mbarrier (wrong descriptor / byte count). Such a kernel records its code in a host-mapped flag and
 * TRAPS: the launch fails with a sticky CUDA error, every later CUDA call of the process fails, and nothing can train,
 * log or checkpoint on top of it. This call only adds the diagnosis. */
int b200cd_device_status(int device, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Input packing — replaces torch.cat((x_t1, x_t2), 1) utils/networks.py:74,106,113, the modality
 * slicing :105,112,233,246 and the implicit NCHW fp32 -> conv operand conversion of the first Conv2d
 * (InConv, :405-412). Writes bf16 im2col rows [n_img*H*W][kpad], k = tap*Cin + ci (zero padded), so the
 * first 3x3 conv is a plain GEMM.
 *   cat_mode 0: images = [src0 batch ; src1 batch], Cin = nc (shared-weight t1 / t2 calls, :141-145)
 *   cat_mode 1: channels = [src0 ; src1], Cin = 2*nc (early fusion, :74)
 *   src0/src1: fp32 NCHW [B][csrc][H][W]; channels c_lo .. c_lo+nc-1 of each are used.
 * ------------------------------------------------------------------------------------------------- */
int b200cd_pack_input(const float* src0, const float* src1, int csrc, int c_lo, int nc, int cat_mode, int B, int H,
                      int W, int kpad, void* out_bf16, void* stream);

/* Weight packing fp32 reference layouts -> bf16 GEMM operands (run every step: the optimizer owns the
 * fp32 master weights, nn.Conv2d.weight / nn.ConvTranspose2d.weight utils/networks.py:392,395,433).
 *   mode 0 conv3x3 forward   w[co=d0][ci=d1][3][3]  -> out[co][tap][ci]
 *   mode 1 conv3x3 dgrad     out[ci][tap'][co] = w[co][ci][2-ky'][2-kx']
 *   mode 2 first layer       out[co][kpad], k = tap*d1 + ci
 *   mode 3 convT forward     w[ci=d0][co=d1][2][2]  -> out[tap*d1 + co][ci]
 *   mode 4 convT dgrad       out[ci][tap*d1 + co] */
int b200cd_pack_weights(int mode, const float* w, void* out_bf16, int d0, int d1, int kpad, void* stream);

/* The same for every weight of a network in ONE launch. `jobs_dev` is a DEVICE array; job j owns the thread blocks
 * [start_j, start_j + b200cd_pack_job_blocks(mode_j, d0_j, d1_j, kpad_j)) (start = running sum, total_blocks = sum).
 * When out2_bf16 is not NULL the same weights are also written in layout mode2 (e.g. mode 0 + mode 1: the forward and
 * the input-gradient operand of one convolution) from a single read of w. */
typedef struct {
  const float* w;
  void* out_bf16;
  void* out2_bf16;
  int32_t mode, mode2, d0, d1, kpad, reserved;
  int64_t start;
} b200cd_pack_job;
int b200cd_pack_job_blocks(int mode, int d0, int d1, int kpad);
int b200cd_pack_weights_batched(const b200cd_pack_job* jobs_dev, int njobs, int64_t total_blocks, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * G1 — tcgen05 implicit GEMM, D[pixel, n] = sum_{tap,k} A_tap[pixel, k] * Bw[n, tap*ka + k].
 *   mode 0: 3x3 same-conv, 9 taps — nn.Conv2d(in,out,3,padding=1) utils/networks.py:392,395 (forward;
 *           with mode-1 packed weights also its input gradient).
 *   mode 1: single tap — the im2col'ed first conv; with out_mode 1 the forward of
 *           nn.ConvTranspose2d(c,c,2,stride=2) utils/networks.py:433 (N = 4*cout, scatter epilogue writes
 *           straight into the channel slice of the concat buffer, replacing F.pad + torch.cat :443,449).
 *   mode 2: 4 taps gathered at stride 2 from a (2H x 2W) tensor — input gradient of the transposed conv.
 *   A: bf16 NHWC view [n_img][H][W][ka] (mode 2: [n_img][2H][2W][ka]) with stride a_ld per pixel.
 *   Bw: bf16 [N][taps*ka] row-major.
 *   out: bf16 NHWC view [n_img][H][W][N] (out_mode 1: [n_img][2H][2W][cout]) with stride out_ld.
 *   bias: fp32 [N] (out_mode 1: [cout]) or NULL.
 *   stats: NULL or fp32 [num_tiles][N][2] per-tile (sum, sum of squares) of the stored bf16 values — the
 *          batch statistics nn.BatchNorm2d (utils/networks.py:393,396) needs; num_tiles from
 *          b200cd_conv_gemm_tiles. A tile never spans two images.
 *   flags: bit 0 (mode 0 only) = load the activation tile once per kx with a one-row halo and serve the three ky
 *          taps from it (less L2 -> SM traffic; same result bit for bit).
 *          bit 1 = use a 128 x 256 output tile when N (out_mode 1: cout) is a multiple of 256 and bit 0 is clear.
 *          bit 2 (every mode; out_mode 1 only with mode 1) = CTA-pair kernel: two SMs compute one 256-pixel tile with cta_group::2 MMAs,
 *          each loading half of the weight tile; persistent, weights resident in shared memory when they fit
 *          (same result bit for bit).
 *          bits 5..6 (with bit 2) = N tile of the CTA-pair kernel chosen by the caller: 1 = 64, 2 = 128, 3 = 256 output
 *          channels per work item (0 = the library's own rule: the widest tile that divides N and leaves >= 48 items);
 *          ignored when it does not divide N. The statistics row count depends on it: pass the same flags to
 *          b200cd_conv_gemm_stat_rows. Same math; per-CTA statistics rows are summed in a different grouping.
 *   requires ka % 64 == 0, N % 64 == 0, a_ld % 8 == 0, out_ld % 8 == 0.
 * ------------------------------------------------------------------------------------------------- */
int b200cd_conv_gemm(int mode, int out_mode, int flags, const void* A, int64_t a_ld, int n_img, int H, int W, int ka,
                     const void* Bw, int N, int cout, void* out, int64_t out_ld, const float* bias, float* stats,
                     void* stream);
/* The input-gradient convolution (mode 0 with dgrad weights, or mode 2) of a layer whose INPUT was produced by
 * BatchNorm+ReLU, with the reduce pass of that BatchNorm's backward fused into the epilogue: besides storing the gradient
 * tile dy it accumulates, per CTA, S1 = sum dy*[y > 0] and S2 = sum dy*[y > 0]*r, where r is the pre-BN tensor of the
 * producing layer (NHWC bf16 view [n_img][H][W][N], stride r_ld) and y = r*scale + shift (scale / shift fp32 [G][N], as
 * written by b200cd_bn_stats). Requires the CTA-pair kernel with per-CTA statistics (flags bits 2 and 3, G in bits
 * 8..15); sums is fp32 [G][rows][N][2] with rows = b200cd_conv_gemm_stat_rows(mode, 0, flags, ...). Feed the result to
 * b200cd_bn_bwd_from_sums. */
int b200cd_conv_gemm_bnbwd(int mode, int flags, const void* A, int64_t a_ld, int n_img, int H, int W, int ka, const void* Bw,
                           int N, void* out_bf16, int64_t out_ld, const void* r, int64_t r_ld, const float* scale,
                           const float* shift, float* sums, void* stream);
/* tiles_per_image for (H, W): stats has n_img * tiles_per_image rows. */
int b200cd_conv_gemm_tiles(int H, int W);
/* flags bit 3 (with bit 2, the CTA-pair kernel) switches the statistics to per-CTA running sums: bits 8..15 of flags hold
 * the number G (1 or 2) of BatchNorm stat-groups (image n belongs to group n / (n_img / G)), stats is then
 * fp32 [G][rows][N][2] with rows = b200cd_conv_gemm_stat_rows(...) (one row per CTA and epilogue group, at most 296
 * instead of one per 128-pixel tile). Without bit 3 the function returns the per-tile row count n_img * tiles. */
int b200cd_conv_gemm_stat_rows(int mode, int out_mode, int flags, int n_img, int H, int W, int ka, int N);

/* ---------------------------------------------------------------------------------------------------
 * G2 — tcgen05 weight-gradient GEMM, ws[split][tap][m][n] = sum_{pixels in split} U[pixel, m] * V_tap[pixel, n]
 * (the weight gradients autograd produces for nn.Conv2d / nn.ConvTranspose2d, utils/networks.py:392,395,433).
 *   mode 0: 3x3 conv, 9 taps; V is read shifted by sign*(kx-1, ky-1) (sign=+1: U = dOut, V = input;
 *           sign=-1: U = input, V = dOut).
 *   mode 1: single tap (first layer on the im2col rows).
 *   mode 2: transposed conv, 4 taps: U = low-res input [n_img][H][W][cu], V = full-res gradient
 *           [n_img][2H][2W][cv] gathered at (2y+dy, 2x+dx).
 *   ws element (split, tap, m, n) at split*split_stride + tap*tap_stride + m*m_stride + n*n_stride.
 *   requires cv % 64 == 0, cu % 64 == 0; H, W arbitrary (out-of-image pixels read as zero).
 *   halo: 1 = load the shifted operand once per pixel tile with a one-row halo (mode 0 only).
 *   splits2: 0, or (mode 0, halo, cv % 128 != 0) the number of CTAs per output block that own the kx = 2 taps: the
 *           `splits` CTAs then own kx = 0 and 1 together (one U tile feeds two V boxes), so taps with tap % 3 == 2 have
 *           splits2 partials and the others `splits` (b200cd_reduce_job.splits2). splits2 = splits / 2 balances the work.
 * ------------------------------------------------------------------------------------------------- */
int b200cd_wgrad_gemm(int mode, int sign, int halo, const void* U, int64_t u_ld, int cu, const void* V, int64_t v_ld,
                      int cv, int n_img, int H, int W, float* ws, int splits, int splits2, int64_t split_stride,
                      int64_t tap_stride, int64_t m_stride, int64_t n_stride, void* stream);
/* number of 8x8 pixel tiles (upper bound for `splits`) */
int b200cd_wgrad_tiles(int n_img, int H, int W);

/* CTAs one pixel-range split of b200cd_wgrad_gemm launches (the caller sizes `splits` so that splits * this fills the
 * 148 SMs): (cu / 128 rounded up) * (cv / N tile) * 3 filter columns for the 3x3 modes — except with halo, 64-wide N
 * tiles and cu <= 64, where one CTA owns all nine taps. -1 on bad arguments. */
int b200cd_wgrad_ctas_per_split(int mode, int halo, int cu, int cv);

/* Fixed-order sum over splits into the reference parameter layout.
 *   layout 0: ws[s][tap][d0][d1] -> grad[d0][d1][tap]  (Conv2d.weight [co][ci][3][3]; ConvTranspose2d.weight [ci][co][2][2])
 *   layout 1: ws[s][d0][ld1], k = tap*d1 + i -> grad[d0][d1][tap], split_stride = d0*ld1 (first layer) */
int b200cd_wgrad_reduce(const float* ws, int splits, int64_t split_stride, int layout, int d0, int d1, int taps,
                        float* grad, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * BatchNorm2d (train or eval), utils/networks.py:393,396 — statistics stage.
 *   partial: the `stats` output of b200cd_conv_gemm, ld = its N; stat-group g owns tiles
 *            [g*tiles_per_group, (g+1)*tiles_per_group) (a group = one reference module call, i.e. one
 *            timestamp of the shared-weight encoder, SURVEY §0 finding 1).
 *   count: elements per channel per group. ws: fp64 [spl][G][C][2].
 *   train=1: batch statistics; running_mean/var are updated once per group in order 0..G-1
 *            (order_rev=1: G-1..0, decoder_sem is called on t2 first, utils/networks.py:191-195) with the
 *            unbiased variance; *nbt += G. train=0: running statistics are used, nothing is updated.
 *   outputs fp32 [G][C]: mean, invstd, scale = gamma*invstd, shift = beta - mean*scale.
 * ------------------------------------------------------------------------------------------------- */
int b200cd_bn_stats(const float* partial, int ld, int C, int tiles_per_group, int G, double count, int spl, double* ws,
                    const float* gamma, const float* beta, float* running_mean, float* running_var, int64_t* nbt,
                    float momentum, float eps, int train, int order_rev, float* mean, float* invstd, float* scale,
                    float* shift, void* stream);

/* BN-apply + nn.ReLU (utils/networks.py:394,397) fused with nn.MaxPool2d(2) (:420), torch.sub(f_t2, f_t1)
 * (:147-150, 183-186, 223-228) and the skip half of torch.cat([x2, x1]) (:449).
 *   r: conv output bf16 [n_img][H][W][C]; diff=1: images [0,n_img/2) are t1, the rest t2, and
 *   dif[n] = a[n + n_img/2] - a[n]. Outputs (each nullable): a, a2 (second copy, e.g. a concat slice),
 *   pool [n_img][H/2][W/2][C], dif [n_img/2][H][W][C], pool_idx uint8 [n_img][H/2][W/2][C] (dense): position 0..3
 *   (row-major in the 2x2 window) of the first maximum of the stored activations — what MaxPool2d's backward needs. */
int b200cd_bn_apply(const void* r, int64_t ld_r, const float* scale, const float* shift, int n_img, int H, int W,
                    int C, int G, int diff, void* a, int64_t ld_a, void* a2, int64_t ld_a2, void* pool, int64_t ld_p,
                    void* dif, int64_t ld_d, void* pool_idx, void* stream);

/* Gradient sources summed on the fly by the BN backward kernels. */
typedef struct {
  int32_t kind;     /* 0 none; 1 bf16 NHWC at this resolution; 2 bf16 NHWC at half resolution, routed to the
                       arg-max of each 2x2 window (MaxPool2d backward); 3 fp32 dz[pixel] times w[channel]
                       (OutConv backward, utils/networks.py:457) */
  const void* ptr;
  const float* w;   /* kind 3: head weights; kind 2: the uint8 pool_idx tensor written by b200cd_bn_apply */
  int64_t ld;
  int32_t n_mod;    /* > 0: source image = n % n_mod and scale = n < n_mod ? scale_lo : scale_hi
                       (the t2 - t1 difference feeds +d to t2 and -d to t1) */
  float scale_lo, scale_hi;
} b200cd_grad_src;

/* BatchNorm2d + ReLU backward: dr = scale*(dy - mean_g(dy) - xhat*mean_g(dy*xhat)), dy = (sum of sources)*[y>0];
 * dgamma[C], dbeta[C] summed over groups. ws: fp32, b200cd_bn_bwd_ws_floats(...) floats. */
int b200cd_bn_bwd(const void* r, int64_t ld_r, const float* mean, const float* invstd, const float* scale,
                  const float* shift, const b200cd_grad_src* srcs /* [3] */, int n_img, int H, int W, int C, int G,
                  float* ws, float* dgamma, float* dbeta, void* dr, int64_t ld_dr, void* stream);
/* The same when the sums S1, S2 of the single gradient source were already accumulated by b200cd_conv_gemm_bnbwd
 * (sums fp32 [G][sum_rows][C][2]): only the finalize and dx kernels run; srcs must describe that one direct source. */
int b200cd_bn_bwd_from_sums(const void* r, int64_t ld_r, const float* mean, const float* invstd, const float* scale,
                            const float* shift, const b200cd_grad_src* srcs, int n_img, int H, int W, int C, int G,
                            const float* sums, int sum_rows, float* ws, float* dgamma, float* dbeta, void* dr_bf16,
                            int64_t ld_dr, void* stream);
size_t b200cd_bn_bwd_ws_floats(int n_img, int H, int W, int C, int G);

/* OutConv (nn.Conv2d(c, 1, 1), utils/networks.py:454-461) forward over one or two C-channel inputs
 * (the fusion heads read torch.concat((x_stream1, x_stream2), 1), :118,255,302, without materialising it). */
int b200cd_head_fwd(const void* a0, int64_t ld0, const void* a1, int64_t ld1, int C, const float* w, const float* b,
                    int64_t npix, float* logits, void* stream);

/* Centre pad of Up (utils/networks.py:440-443, F.pad of the transposed-conv output to the skip tensor's size):
 * dst[n][top + y][left + x][0..C) = src[n][y][x][0..C) for the h x w source, zero elsewhere in the H x W window.
 * Both tensors NHWC bf16 with pixel strides ld (dst is typically the upper half of a concat buffer). Used by
 * inference on tiles whose levels have odd sizes; training tiles (multiples of 16) never need it. */
int b200cd_pad_copy(const void* src, int64_t ld_src, int n_img, int h, int w, int C, void* dst, int64_t ld_dst, int H,
                    int W, int top, int left, void* stream);

/* out[c] = sum_pixels wgt[pixel] * x[pixel][c]  (x == NULL: C = 1, out = sum wgt; wgt == NULL: plain column sum).
 * OutConv weight/bias gradients and the ConvTranspose2d bias gradient. ws: fp32 [nblk*C]. */
int b200cd_colsum(const void* x, int64_t ld, int C, const float* wgt, int64_t npix, int nblk, float* ws, float* out,
                  void* stream);
/* out[j] = sum over rows of stats[row][c_off + j][0] for j < C, where stats is the fp32 [rows][ld][2] per-CTA
 * statistics buffer of one stat-group written by b200cd_conv_gemm (flags bit 3): the nn.ConvTranspose2d bias gradient
 * (utils/networks.py:433) is the pixel sum of the upper half of the concat-buffer gradient, which the input-gradient
 * convolution of the following DoubleConv has just stored. */
int b200cd_stat_rowsum(const float* stats, int rows, int ld, int c_off, int C, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * power_jaccard_loss, utils/loss_functions.py:141-150. sums = (sum p*t, sum p^2, sum t^2) in fp64 over the
 * selected rows (rowmask == NULL: all; else rows with (rowmask[row] != 0) == sel, the boolean row
 * indexing of train_semisupervised.py:85-87,102-104). t_is_logit: the target is sigmoid(t) and receives
 * a gradient (MMCR consistency term, :75-76,105). ws: fp64 [nblk*3].
 * A data-parallel caller all-reduces `sums` (SUM) between b200cd_pj_fwd and b200cd_pj_loss / _bwd.
 * ------------------------------------------------------------------------------------------------- */
int b200cd_pj_fwd(const float* z, const float* t, int t_is_logit, const unsigned char* rowmask, int sel, int rows,
                  int64_t per_row, int nblk, double* ws, double* sums, void* stream);
int b200cd_pj_loss(const double* sums, float* loss, void* stream);
/* dz (+)= g * dL/dz, dt (+)= g * dL/dt (dt nullable), g = gmul * (gptr ? *gptr : 1). accumulate=0 overwrites
 * (unselected rows get 0). */
int b200cd_pj_bwd(const float* z, const float* t, int t_is_logit, const unsigned char* rowmask, int sel, int rows,
                  int64_t per_row, const double* sums, const float* gptr, float gmul, int accumulate, float* dz,
                  float* dt, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * The split reduction of b200cd_wgrad_reduce for MANY layers in one launch (each layer has its own workspace region).
 * `jobs_dev` is a DEVICE array; job j owns thread blocks [start_j, start_j + b200cd_reduce_job_blocks(splits, d0, d1,
 * taps)); parts must be b200cd_reduce_job_parts(splits, d1, taps) (0 selects the row-transposing path used for few
 * splits: coalesced gradient stores). Layout 0 only (ws[split][tap][d0][d1] -> grad[d0][d1][tap]); requires
 * d1 % 4 == 0, split_stride % 4 == 0 and 16-byte aligned workspace and gradient pointers.
 * Deterministic: fixed summation order.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* ws;
  float* grad;
  int64_t split_stride;
  int64_t start;
  int32_t splits, layout, d0, d1, taps, parts;
  int32_t splits2;   /* > 0: taps with tap % 3 == 2 have splits2 partials (b200cd_wgrad_gemm splits2) */
  int32_t reserved;
} b200cd_reduce_job;
int b200cd_reduce_job_parts(int splits, int d1, int taps);
int64_t b200cd_reduce_job_blocks(int splits, int d0, int d1, int taps);
int b200cd_wgrad_reduce_batched(const b200cd_reduce_job* jobs_dev, int njobs, int64_t total_blocks, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Thresholded confusion counts — MultiThresholdMetric.add_sample, utils/metrics.py:23-31, as used by
 * utils/evaluation.py:23 — in one pass over the prediction and the label: for each of nthr (<= 8) thresholds
 * pred = round(p - thr + 0.5) != 0, where p = pred[i] (from_logits = 0) or sigmoid(pred[i]) (from_logits = 1), and
 * counts[t][0..3] += (TP, TN, FP, FN) in the REFERENCE's naming (its FP counts y_true & ~pred, its FN ~y_true & pred).
 * counts is uint64 [nthr][4] on the device and is ACCUMULATED into (zero it before the first sample).
 * ------------------------------------------------------------------------------------------------- */
int b200cd_confusion_counts(const float* pred, const float* truth, int64_t n, int from_logits, const float* thresholds,
                            int nthr, uint64_t* counts, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * AdamW step over every parameter tensor in ONE launch — optim.AdamW(net.parameters(), lr=cfg.TRAINER.LR,
 * weight_decay=0.01) train_supervised.py:32 (decoupled weight decay, betas (0.9, 0.999), eps 1e-8, no amsgrad).
 * `jobs_dev` is a DEVICE array; job j owns thread blocks [start_j, start_j + ceil(n_j / 1024)). Parameters without a
 * gradient (grad is None in the reference: outc_sem_change) get no job. step_count is the 1-based step number t
 * used for the bias corrections 1 - beta^t.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
  int64_t start;
  int32_t vec4;      /* 1 when all four pointers are 16-byte aligned */
  int32_t reserved;
} b200cd_adamw_job;
int b200cd_adamw_step(const b200cd_adamw_job* jobs_dev, int njobs, int64_t total_blocks, double lr, double beta1,
                      double beta2, double eps, double weight_decay, int64_t step_count, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Workspace queries: every entry point takes caller-allocated workspaces; this returns the bytes each needs (or -1).
 *   B200CD_WS_BN_BWD      d = {n_img, H, W, C, G}                         fp32 workspace of b200cd_bn_bwd*
 *   B200CD_WS_CONV_STATS  d = {mode, out_mode, flags, n_img, H, W, ka, N}  statistics side output of b200cd_conv_gemm
 *   B200CD_WS_WGRAD       d = {splits, taps, cu, cv}                       split partials of b200cd_wgrad_gemm*
 *   B200CD_WS_COLSUM      d = {nblk, C}                                    block partials of b200cd_colsum*
 *   B200CD_WS_PJ          d = {nblk}                                       block partials of b200cd_pj_fwd
 *   B200CD_WS_BN_STATS    d = {spl, G, C}                                  second-stage partials of b200cd_bn_stats
 * ------------------------------------------------------------------------------------------------- */
#define B200CD_WS_BN_BWD 1
#define B200CD_WS_CONV_STATS 2
#define B200CD_WS_WGRAD 3
#define B200CD_WS_COLSUM 4
#define B200CD_WS_PJ 5
#define B200CD_WS_BN_STATS 6
int64_t b200cd_query_workspace(int op, const int64_t* dims, int ndims);

/* ---------------------------------------------------------------------------------------------------
 * Multi-GPU: a library-owned NCCL communicator, one process per GPU — replaces the gather / reduce-add of
 * nn.DataParallel (utils/networks.py:27). NCCL is resolved with dlopen at run time (the copy already loaded in the
 * process, e.g. torch's, or `libnccl_path`); streams and buffers stay the caller's.
 *   rank 0: b200cd_comm_unique_id(id) -> ship the 128 bytes to the other ranks by any means
 *   all:    b200cd_comm_init(id, rank, nranks)  (CUDA device = the caller's current device)
 *   b200cd_allreduce_bucket: in-place SUM of a fp32 gradient bucket on `stream` (asynchronous, graph-capturable)
 *   b200cd_allreduce_f64:    in-place SUM of the power-Jaccard partial sums (3 per loss term)
 * ------------------------------------------------------------------------------------------------- */
int b200cd_comm_load(const char* libnccl_path);
int b200cd_comm_version(void);
int b200cd_comm_unique_id(void* id128);
int b200cd_comm_init(const void* id128, int rank, int nranks);
int b200cd_comm_size(void);
int b200cd_allreduce_bucket(float* buf, int64_t count, void* stream);
int b200cd_allreduce_f64(double* buf, int64_t count, void* stream);
int b200cd_comm_destroy(void);

/* ---------------------------------------------------------------------------------------------------
 * Plan graphs: a step's launches, captured by the caller into a cudaGraph_t (any capture API; the Python host layer
 * uses torch.cuda.graph), instantiated and launched by the library so that the per-node stream priorities recorded at
 * capture time are honoured (cudaGraphInstantiateFlagUseNodePriority): the plan's dependent chain runs on
 * high-priority streams, the weight-gradient GEMMs on a default-priority one. The caller keeps ownership of the
 * cudaGraph_t and of every buffer the graph touches; the executable graph is the library's until destroyed.
 * ------------------------------------------------------------------------------------------------- */
int b200cd_graph_instantiate(void* cuda_graph, int use_node_priority, void** exec_out);
int b200cd_graph_launch(void* graph_exec, void* stream);
int b200cd_graph_exec_destroy(void* graph_exec);

/* ---------------------------------------------------------------------------------------------------
 * Training-time augmentation + packing (next-row N2) — replaces the per-sample numpy transforms of
 * utils/augmentations.py:6-142 as composed by utils/datasets.py:111-181: crop (UniformCrop / ImportanceRandomCrop
 * :105-142) -> RandomFlip :44-62 -> RandomRotate :65-72 -> ColorShift :75-86 -> GammaCorrection :89-101 ->
 * Numpy2Torch :35-41 (HWC -> CHW), all samples of a batch in one launch. The random decisions are drawn on the host
 * (numpy, the reference's call order) and passed in: one job per sample. out: fp32 [n][cout][crop][crop];
 * output channel c of sample s = source channel cmap[c] of the crop at (x0, y0). `jobs_dev` is a DEVICE array.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* src; /* [H0][W0][C] fp32 (HWC) */
  int32_t H0, W0, C;
  int32_t x0, y0;
  int32_t hflip, vflip, rotk; /* rotk = number of counter-clockwise 90 degree rotations (np.rot90), 0..3 */
  int32_t use_mul, use_gamma;
  int32_t cmap[16];
  float mul[16];   /* per source channel */
  float gamma[16];
  int32_t reserved;
} b200cd_augment_job;
int b200cd_augment(const b200cd_augment_job* jobs_dev, int n, int crop, int cout, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Inference (utils/evaluation.py:7-23, net.eval()): BatchNorm is a fixed per-channel affine, so conv + BN + ReLU
 * (utils/networks.py:392-397) is ONE launch: b200cd_conv_gemm_affine = b200cd_conv_gemm (CTA-pair kernel, flags bit 2;
 * bit 4 for split-bf16 tensors) whose epilogue stores act((acc + bias[n]) * scale[n] + shift[n]), act = ReLU when
 * relu != 0, and writes no statistics. b200cd_bn_eval_affine_batched computes (mean, invstd, scale, shift) of EVERY
 * BatchNorm of a network from its running statistics in one launch: job j owns thread blocks
 * [start_j, start_j + ceil(C_j / 256)); `jobs_dev` is a DEVICE array.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* gamma;
  const float* beta;
  const float* running_mean;
  const float* running_var;
  float* mean;   /* [G][C] each */
  float* invstd;
  float* scale;
  float* shift;
  int32_t C, G;
  float eps;
  int32_t start;
} b200cd_bn_eval_job;
int b200cd_bn_eval_affine_batched(const b200cd_bn_eval_job* jobs_dev, int njobs, int total_blocks, void* stream);
int b200cd_conv_gemm_affine(int mode, int flags, const void* A, int64_t a_ld, int n_img, int H, int W, int ka,
                            const void* Bw, int N, void* out, int64_t out_ld, const float* bias, const float* scale,
                            const float* shift, int relu, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * "Precise" mode: split-bf16 storage (ABI version 2).
 *
 * north_star asks for logits / gradients within 1e-3 of the reference's fp32 path; single bf16 (or TF32) operands
 * cannot deliver that through train-mode BatchNorm and the siamese t2 - t1 differences. In this mode
 * every activation / gradient element is stored as TWO bf16 values, hi = bf16(x) and lo = bf16(x - hi) (16 mantissa
 * bits together), laid out per pixel as [hi (channels) | lo (channels)]: a tensor is passed as the pointer to its hi
 * half with its pixel stride `ld`, and its lo half lies exactly ld / 2 elements behind (so channel slices of a
 * concatenation buffer [hi skip | hi up | lo skip | lo up] keep working). The tensor-core kernels form every product
 * as three bf16 MMAs with fp32 accumulation (hi*hi + hi*lo + lo*hi):
 *   - b200cd_conv_gemm with flags bit 4 (16): A and out are split tensors (ld % 16 == 0, ld / 2 >= channels), `ka`
 *     stays the real channel count, Bw is [N][taps][3*ka] = [hi | lo | hi] per tap (b200cd_pack_weights_hp_batched),
 *     statistics are taken from the stored hi + lo values. Needs the CTA-pair kernel (flags bit 2).
 *   - b200cd_wgrad_gemm_hp: U and V split tensors; each pixel tile runs three times (U_hi*V_hi, U_hi*V_lo, U_lo*V_hi).
 * The memory-bound kernels below read hi + lo, compute in fp32 exactly like their bf16-storage counterparts and write
 * both halves. Everything that is fp32 already (statistics, loss, parameter gradients, AdamW) is shared.
 * ------------------------------------------------------------------------------------------------- */
int b200cd_wgrad_gemm_hp(int mode, int sign, int halo, const void* U, int64_t u_ld, int cu, const void* V, int64_t v_ld,
                         int cv, int n_img, int H, int W, float* ws, int splits, int splits2, int64_t split_stride,
                         int64_t tap_stride, int64_t m_stride, int64_t n_stride, void* stream);
/* b200cd_pack_input writing [pixel][hi kpad | lo kpad] rows (ld = 2 * kpad) */
int b200cd_pack_input_hp(const float* src0, const float* src1, int csrc, int c_lo, int nc, int cat_mode, int B, int H,
                         int W, int kpad, void* out_bf16, void* stream);
/* b200cd_pack_weights_batched writing the K-tripled [hi | lo | hi] operands; same job struct, blocks per job from
 * b200cd_pack_job_blocks_hp */
int b200cd_pack_job_blocks_hp(int mode, int d0, int d1, int kpad);
int b200cd_pack_weights_hp_batched(const b200cd_pack_job* jobs_dev, int njobs, int64_t total_blocks, void* stream);
/* b200cd_bn_apply / b200cd_bn_bwd / b200cd_head_fwd / b200cd_colsum on split tensors (same arguments) */
int b200cd_bn_apply_hp(const void* r, int64_t ld_r, const float* scale, const float* shift, int n_img, int H, int W,
                       int C, int G, int diff, void* a, int64_t ld_a, void* a2, int64_t ld_a2, void* pool, int64_t ld_p,
                       void* dif, int64_t ld_d, void* pool_idx, void* stream);
int b200cd_bn_bwd_hp(const void* r, int64_t ld_r, const float* mean, const float* invstd, const float* scale,
                     const float* shift, const b200cd_grad_src* srcs, int n_img, int H, int W, int C, int G, float* ws,
                     float* dgamma, float* dbeta, void* dr, int64_t ld_dr, void* stream);
int b200cd_head_fwd_hp(const void* a0, int64_t ld0, const void* a1, int64_t ld1, int C, const float* w, const float* b,
                       int64_t npix, float* logits, void* stream);
int b200cd_colsum_hp(const void* x, int64_t ld, int C, const float* wgt, int64_t npix, int nblk, float* ws, float* out,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CD_H_ */
