// Internal (C++) launch interface between the C-ABI layer (abi.cu) and the kernel translation units.
// Nothing here is exported; the exported surface is include/b200cd.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

namespace b200cd {

// thread-local message behind b200cd_last_error (abi.cu); other translation units report through it
void set_last_error(const char* msg);

// Launch helper. With B200CD_PDL=1 in the environment kernels are launched with the programmatic-dependent-launch
// attribute (see ptx.cuh: pdl_wait); by default they are plain stream-ordered launches and the kernels'
// griddepcontrol instructions are no-ops (measured on B200: PDL does not shorten the graph-replayed step).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ----------------------------------------------------------------------------------------------
// G1: implicit-GEMM "fprop-like" kernel.  D[pixel, n] = sum_{tap, k} A_tap[pixel, k] * B[n, tap*ka + k]
//   mode 0: 3x3 same-convolution (9 taps, zero padding through TMA out-of-bounds fill)
//   mode 1: single tap (plain row GEMM: transposed-conv forward, im2col'ed first layer)
//   mode 2: 4 taps gathered with stride 2 (transposed-conv input gradient)
//   out_mode 0: NHWC tile store at the GEMM's own resolution
//   out_mode 1: 2x2/stride-2 scatter (transposed-conv forward), N = 4*cout, tap = n / cout
// ----------------------------------------------------------------------------------------------
struct FpropParams {
  int mode, out_mode;
  int taps, kchunks, ka;  // K loop = taps * kchunks chunks of 64 channels; ka = channels per tap
  int tw, th, tiles_x, tiles_y;
  int rows;               // tw*th <= 128 rows of the tile are real (the A box); the rest of the MMA tile is ignored
  int H, W;               // pixel grid of the GEMM rows
  int N, cout;
  const float* bias;      // nullable
  float2* stats;          // nullable: per (tile, n) partial (sum, sum of squares) of the bf16-rounded output
  int ragged;             // 1 when H % th or W % tw != 0 (mask rows in the statistics)
  // CTA-pair kernel only: stat_groups > 0 switches the statistics to per-CTA running sums, stats[g][row][n] with
  // stat_rows rows per BatchNorm stat-group g = image / (n_img / stat_groups) (row = 2 * blockIdx.x + epilogue group)
  int stat_groups, stat_rows, n_img;
  // CTA-pair kernel, per-CTA statistics only: when bwd_r != nullptr the tile being stored is the gradient dy w.r.t. a
  // BatchNorm+ReLU output whose pre-BN tensor is bwd_r (same pixel grid, channel n of the tile = channel n of bwd_r):
  // the statistics become the BatchNorm-backward sums S1 = sum dy*[y > 0], S2 = sum dy*[y > 0]*r with
  // y = r*scale + shift (scale/shift: [stat_groups][N]) — the reduce pass of b200cd_bn_bwd, fused into the epilogue.
  const void* bwd_r;
  long long bwd_ld;
  const float* bwd_scale;
  const float* bwd_shift;
  // Split-bf16 ("precise") operands, CTA-pair kernel only. Every activation row holds its values twice: the bf16
  // rounding `hi` in channels [0, C) and the bf16-rounded remainder `lo = bf16(x - hi)` a_lo / o_lo elements further
  // (hi + lo carries 16 mantissa bits). The K loop then has 3 * kreal chunks per tap — A_hi * W_hi, A_hi * W_lo,
  // A_lo * W_hi against weights packed as [hi | lo | hi] per tap — and the epilogue stores both halves of the output.
  int prec, kreal, a_lo, o_lo;
  // Inference: BatchNorm folded into the epilogue (CTA-pair kernel, no statistics): out = act((acc + bias) * ep_scale[n]
  // + ep_shift[n]) with act = ReLU when ep_relu — one launch per conv + BN + ReLU stage (utils/evaluation.py:7-23).
  const float* ep_scale;
  const float* ep_shift;
  int ep_relu;
  int* err;
};
// CTAs the CTA-pair kernel launches for this problem (needs the current device: occupancy query on first use)
int fprop_pair_ctas(const FpropParams& p, int bn, int num_tiles);
cudaError_t launch_fprop(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapO,
                         const FpropParams& p, int bn, int halo, int num_tiles, cudaStream_t stream);

// true when the CTA-pair kernel for (mode, bn) wants mapB as the rank-4 view {k within tap, n, ky, kx} with box
// {64, bn / 2, 3, 1} (one load per step) instead of the rank-2 [N][taps * ka] view with box {64, bn / 2}
bool fprop_pair_stacked_weights(int mode, int bn);

// G1p (gemm_fprop2.cu): mode 0 / out_mode 0 on CTA pairs (cta_group::2), persistent, halo A boxes (p.rows = tw*(th+2)).
cudaError_t launch_fprop_pair(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapO,
                              const FpropParams& p, int bn, int num_tiles, cudaStream_t stream);

// ----------------------------------------------------------------------------------------------
// G2: weight-gradient kernel. D_tap[m, n] = sum_pixels U[pixel, m] * V_tap[pixel, n]
//   (both operands MN-major straight out of NHWC memory), split over pixel ranges.
//   mode 0: 3x3 conv (CTA = one kx, three ky taps; V shifted by sign*(kx-1, ky-1))
//   mode 1: single tap
//   mode 2: transposed conv (4 taps, V gathered with stride 2 from the full-resolution gradient)
// ----------------------------------------------------------------------------------------------
// true when launch_wgrad runs the M-stacked all-taps-per-CTA variant for (mode 0, halo, 64-wide N tiles, cu)
bool wgrad_mstack(int cu);

struct WgradParams {
  int mode, sign;
  int H, W, tiles_x, tiles_y, total_tiles, splits;
  int splits2;  // > 0 (64-wide N tiles, mode 0, halo): CTAs own two kx columns; splits2 CTAs own the third (gemm_wgrad.cu)
  int cu, cv;
  float* ws;
  long long split_stride, tap_stride, m_stride, n_stride;
  // split-bf16 operands (see FpropParams): passes = 3 runs every pixel tile three times, U_hi * V_hi, U_hi * V_lo and
  // U_lo * V_hi, into the same accumulators; u_lo / v_lo = channel offset of the lo halves. passes = 1 otherwise.
  int passes, u_lo, v_lo;
  int* err;
};
cudaError_t launch_wgrad(const CUtensorMap& mapU, const CUtensorMap& mapV, const WgradParams& p, int bn, int halo,
                         cudaStream_t stream);

// ----------------------------------------------------------------------------------------------
// Memory-bound kernels (elementwise.cu)
// ----------------------------------------------------------------------------------------------
struct GradSrc {
  int kind;          // 0 none, 1 direct bf16 NHWC, 2 pooled bf16 (half resolution, routed to the arg-max), 3 head (dz*w)
  const void* ptr;   // bf16 tensor (kinds 1, 2) or fp32 dz[pixel] (kind 3)
  const float* w;    // kind 3: 1x1 head weights for these channels; kind 2: the uint8 arg-max index tensor
  long long ld;      // elements per pixel
  int n_mod;         // > 0: source image = n % n_mod, scale = (n < n_mod) ? scale_lo : scale_hi
  float scale_lo, scale_hi;
};
struct GradSrcs {
  GradSrc s[3];
};

cudaError_t launch_pack_input(const float* src0, const float* src1, int csrc, int c_lo, int nc, int cat_mode, int B,
                              int H, int W, int kpad, void* out, cudaStream_t st);
struct PackJob {        // layout == b200cd_pack_job (include/b200cd.h)
  const float* w;
  void* out;
  void* out2;           // optional second operand layout (mode2) written from the same read of w
  int mode, mode2, d0, d1, kpad, reserved;
  long long start;      // first thread block (running sum of pack_job_blocks over the jobs) this job owns
};
int pack_job_blocks(int mode, int d0, int d1, int kpad);
cudaError_t launch_pack_weights_batched(const PackJob* jobs, int njobs, long long total_blocks, cudaStream_t st);
cudaError_t launch_pack_weights(int mode, const float* w, void* out, int d0, int d1, int kpad, cudaStream_t st);
cudaError_t launch_bn_stats_reduce(const float2* partial, int ld, int C, int tiles_per_group, int G, int spl,
                                   double* partial2, cudaStream_t st);
cudaError_t launch_bn_stats_fused(const float2* partial, int ld, int rows, int C, int G, double count, const float* gamma,
                                  const float* beta, float* running_mean, float* running_var, long long* nbt,
                                  float momentum, float eps, int order_rev, float* mean, float* invstd, float* scale,
                                  float* shift, cudaStream_t st);
cudaError_t launch_bn_finalize(const double* partial2, int spl, int C, int G, double count, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, long long* nbt,
                               float momentum, float eps, int train, int order_rev, float* mean, float* invstd,
                               float* scale, float* shift, cudaStream_t st);
struct BnEvalJob {      // layout == b200cd_bn_eval_job (include/b200cd.h)
  const float* gamma;
  const float* beta;
  const float* running_mean;
  const float* running_var;
  float* mean;          // [G][C] each: the four per-(group, channel) vectors the apply / fused-epilogue kernels read
  float* invstd;
  float* scale;
  float* shift;
  int C, G;
  float eps;
  int start;            // first thread block (256 channels per block)
};
cudaError_t launch_bn_eval_affine_batched(const BnEvalJob* jobs, int njobs, int total_blocks, cudaStream_t st);
cudaError_t launch_bn_apply(const void* r, long long ld_r, const float* scale, const float* shift, int n_img, int H,
                            int W, int C, int G, int diff, void* a, long long ld_a, void* a2, long long ld_a2,
                            void* pool, long long ld_p, void* dif, long long ld_d, void* pool_idx, cudaStream_t st);
cudaError_t launch_bn_bwd_reduce(const void* r, long long ld_r, const float* scale, const float* shift,
                                 const GradSrcs& srcs, int n_img, int H, int W, int C, int G, int nblk, float* partial,
                                 cudaStream_t st);
cudaError_t launch_bn_bwd_finalize(const float* partial, int nblk, int C, int G, double count, const float* mean,
                                   const float* invstd, const float* scale, float* dgamma, float* dbeta, float* coefA,
                                   float* coefB, cudaStream_t st);
cudaError_t launch_bn_bwd_dx(const void* r, long long ld_r, const float* scale, const float* shift, const float* coefA,
                             const float* coefB, const GradSrcs& srcs, int n_img, int H, int W, int C, int G, int nblk,
                             void* dr, long long ld_dr, cudaStream_t st);
cudaError_t launch_head_fwd(const void* a0, long long ld0, const void* a1, long long ld1, int C, const float* w,
                            const float* b, long long npix, float* logits, cudaStream_t st);
cudaError_t launch_pad_copy(const void* src, long long ld_s, int n_img, int h, int w, int C, void* dst, long long ld_d,
                            int H, int W, int top, int left, cudaStream_t st);
cudaError_t launch_colsum(const void* x, long long ld, int C, const float* wgt, long long npix, int nblk,
                          float* partial, cudaStream_t st);
cudaError_t launch_stat_rowsum(const float2* stats, int rows, int ld, int c_off, int C, float* out, cudaStream_t st);
cudaError_t launch_colsum_finalize(const float* partial, int nblk, int C, float* out, cudaStream_t st);
cudaError_t launch_wgrad_reduce(const float* ws, int splits, long long split_stride, int layout, int d0, int d1,
                                int taps, float* grad, cudaStream_t st);
cudaError_t launch_pj_reduce(const float* z, const float* t, int t_is_logit, const unsigned char* rowmask, int sel,
                             int rows, long long per_row, int nblk, double* partial, cudaStream_t st);
cudaError_t launch_pj_finalize(const double* partial, int nblk, double* sums, cudaStream_t st);
cudaError_t launch_pj_loss(const double* sums, float* loss, cudaStream_t st);
cudaError_t launch_pj_bwd(const float* z, const float* t, int t_is_logit, const unsigned char* rowmask, int sel,
                          int rows, long long per_row, const double* sums, const float* gptr, float gmul,
                          int accumulate, float* dz, float* dt, cudaStream_t st);

cudaError_t launch_confusion(const float* pred, const float* truth, long long n, int from_logits, const float* thr, int nthr,
                             unsigned long long* counts, cudaStream_t st);

struct ReduceJob {      // layout == b200cd_reduce_job (include/b200cd.h)
  const float* ws;
  float* grad;
  long long split_stride;
  long long start;      // first thread block of the job (reduce_job_blocks blocks)
  int splits, layout, d0, d1, taps, parts;
  int splits2, reserved;  // > 0: taps with tap % 3 == 2 (kx = 2) have only splits2 partials
};
int reduce_job_parts(int splits, int d1, int taps);   // 0 = row-transposing path
long long reduce_job_blocks(int splits, int d0, int d1, int taps);
cudaError_t launch_wgrad_reduce_batched(const ReduceJob* jobs, int njobs, long long total_blocks, cudaStream_t st);

struct AugmentJob {     // layout == b200cd_augment_job (include/b200cd.h)
  const float* src;     // [H0][W0][C] fp32, HWC as the dataset loads it
  int H0, W0, C;
  int x0, y0;           // crop origin
  int hflip, vflip, rotk;
  int use_mul, use_gamma;
  int cmap[16];         // output channel -> source channel
  float mul[16];        // per SOURCE channel
  float gamma[16];
  int reserved;
};
cudaError_t launch_augment(const AugmentJob* jobs, int n, int cs, int cout, float* out, cudaStream_t st);

struct AdamWJob {       // layout == b200cd_adamw_job (include/b200cd.h)
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  long long start;      // first thread block (1024 elements per block)
  int vec4;             // all four pointers 16-byte aligned
  int reserved;
};
cudaError_t launch_adamw(const AdamWJob* jobs, int njobs, long long total_blocks, float decay, float step_size, float omb1,
                         float b2, float omb2, float eps, float sqrt_bc2, cudaStream_t st);


// split-bf16 ("precise") variants (elementwise_hp.cu); same arguments as the bf16-storage launchers above
cudaError_t launch_pack_input_hp(const float* src0, const float* src1, int csrc, int c_lo, int nc, int cat_mode, int B,
                                 int H, int W, int kpad, void* out, cudaStream_t st);
int pack_job_blocks_hp(int mode, int d0, int d1, int kpad);
cudaError_t launch_pack_weights_hp_batched(const PackJob* jobs, int njobs, long long total_blocks, cudaStream_t st);
cudaError_t launch_bn_apply_hp(const void* r, long long ld_r, const float* scale, const float* shift, int n_img, int H,
                               int W, int C, int G, int diff, void* a, long long ld_a, void* a2, long long ld_a2,
                               void* pool, long long ld_p, void* dif, long long ld_d, void* pool_idx, cudaStream_t st);
cudaError_t launch_bn_bwd_reduce_hp(const void* r, long long ld_r, const float* scale, const float* shift,
                                    const GradSrcs& srcs, int n_img, int H, int W, int C, int G, int nblk,
                                    float* partial, cudaStream_t st);
cudaError_t launch_bn_bwd_dx_hp(const void* r, long long ld_r, const float* scale, const float* shift,
                                const float* coefA, const float* coefB, const GradSrcs& srcs, int n_img, int H, int W,
                                int C, int G, int nblk, void* dr, long long ld_dr, cudaStream_t st);
cudaError_t launch_head_fwd_hp(const void* a0, long long ld0, const void* a1, long long ld1, int C, const float* w,
                               const float* b, long long npix, float* logits, cudaStream_t st);
cudaError_t launch_colsum_hp(const void* x, long long ld, int C, const float* wgt, long long npix, int nblk,
                             float* partial, cudaStream_t st);

}  // namespace b200cd
