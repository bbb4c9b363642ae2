// C-ABI layer (include/b200cd.h): argument validation, TMA tensor-map construction, error codes.
// Everything exported is extern "C" with plain pointers and sizes; kernels live in the other .cu files.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/b200cd.h"
#include "kernels.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail(B200CD_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kMaxDevices = 16;
EncodeTiledFn g_encode = nullptr;
int* g_err_flag[kMaxDevices] = {nullptr};

int current_err_flag(int** out) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices || g_err_flag[dev] == nullptr)
    return fail(B200CD_ERR_ARCH, "b200cd_init(%d) has not been called for the current device", dev);
  *out = g_err_flag[dev];
  return 0;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// rank-2 or rank-5 bf16 tensor map with the 128-byte swizzle; dims/box innermost first, strides in bytes
// for dims 1..rank-1.
int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
             const uint32_t* box) {
  if (g_encode == nullptr) return fail(B200CD_ERR_ARCH, "b200cd_init has not been called");
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (box[i] == 0 || box[i] > 256) return fail(B200CD_ERR_SHAPE, "TMA box dim %d = %u out of range", i, box[i]);
  }
  for (int i = 0; i < rank - 1; ++i) {
    gs[i] = strides[i];
    if (strides[i] % 16 != 0) return fail(B200CD_ERR_ALIGN, "TMA stride %d = %llu not a multiple of 16 bytes", i,
                                          static_cast<unsigned long long>(strides[i]));
  }
  if (!aligned16(base)) return fail(B200CD_ERR_ALIGN, "TMA base pointer not 16-byte aligned");
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gd,
                        gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200CD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// NHWC view [n][h][w][c] with per-pixel stride ld (elements) as a rank-5 map (trailing unit dim).
int make_nhwc_map(CUtensorMap* m, const void* base, int64_t ld, int c, int w, int h, int n, uint32_t bw, uint32_t bh) {
  const uint64_t e = 2;
  const uint64_t dims[5] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n, 1};
  const uint64_t strides[4] = {(uint64_t)ld * e, (uint64_t)w * ld * e, (uint64_t)h * w * ld * e,
                               (uint64_t)n * h * w * ld * e};
  const uint32_t box[5] = {64, bw, bh, 1, 1};
  return make_map(m, base, 5, dims, strides, box);
}

// Full-resolution NHWC tensor [n][2h][2w][c] viewed as [n*h][dy][w][dx][c]: the 2x2/stride-2 gather/scatter
// of the transposed convolution as one box per (dy, dx).
int make_up2_map(CUtensorMap* m, const void* base, int64_t ld, int c, int w, int h, int n, uint32_t bw, uint32_t bh) {
  const uint64_t e = 2;
  const uint64_t dims[5] = {(uint64_t)c, 2, (uint64_t)w, 2, (uint64_t)n * h};
  const uint64_t strides[4] = {(uint64_t)ld * e, 2 * (uint64_t)ld * e, 2 * (uint64_t)w * ld * e,
                               4 * (uint64_t)w * ld * e};
  const uint32_t box[5] = {64, 1, bw, 1, bh};
  return make_map(m, base, 5, dims, strides, box);
}

// Pixel tile of the implicit GEMM: tw x th <= 128 pixels of ONE image. When the row dimension of a tensor map is the
// merged (image, row) index (2x2/stride-2 gather or scatter), th must divide H so a tile never spans two images.
void tile_shape(int W, int H, bool merged_rows, int* tw, int* th) {
  if (W >= 16) {
    *tw = 16;
    *th = 8;
  } else {
    *tw = 8;
    *th = 16;
  }
  if (merged_rows) {
    int t = *th < H ? *th : H;
    while (H % t != 0) --t;
    *th = t;
  }
}

// blocks per stat-group of the BN-backward kernels: up to ~4 waves of 2 CTAs/SM, at least 32 pixels per lane
int bn_bwd_nblk(int n_img, int H, int W, int C, int G) {
  const long long units = static_cast<long long>(n_img / G) * H * W;
  const int lanes = 256 / (C / 8);
  long long nblk = units / (static_cast<long long>(lanes) * 32);
  const long long cap = (148 * 2 * 4) / G;
  // small layers (16 x 16, 32 x 32 levels): down to 8 pixels per lane so that the launch still spreads over ~4 CTAs per
  // SM — a 4 MB layer on 64 CTAs is a chain of dependent DRAM round trips, not a bandwidth problem
  const long long want = (148 * 4) / G, fine = units / (static_cast<long long>(lanes) * 8);
  static const bool fine_on = [] { const char* e = getenv("B200CD_BN_FINE"); return !(e && e[0] == '0'); }();
  if (fine_on && nblk < want) nblk = fine < want ? fine : want;
  if (nblk > cap) nblk = cap;
  if (nblk < 1) nblk = 1;
  return static_cast<int>(nblk);
}

bool chan_ok(int C) { return C >= 64 && C % 64 == 0 && 256 % (C / 8) == 0; }

}  // namespace

namespace b200cd {
void set_last_error(const char* msg) { g_last_error = msg; }
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("B200CD_PDL");
    return e != nullptr && e[0] == '1';  // measured neutral-to-slightly-negative inside CUDA graphs: off by default
  }();
  return on;
}
}  // namespace b200cd

extern "C" {

int b200cd_abi_version(void) { return B200CD_ABI_VERSION; }
const char* b200cd_last_error(void) { return g_last_error.c_str(); }

int b200cd_init(int device) {
  if (device < 0 || device >= kMaxDevices) return fail(B200CD_ERR_SHAPE, "device index %d out of range", device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(B200CD_ERR_ARCH, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
                prop.minor);
  CUDA_TRY(cudaSetDevice(device));
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (fn == nullptr || q != cudaDriverEntryPointSuccess)
      return fail(B200CD_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  if (g_err_flag[device] == nullptr) {
    // pinned, mapped host memory (device-accessible through UVA): a kernel that times out records its code here and
    // traps, and the code is still readable after the trap has poisoned the context
    int* p = nullptr;
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&p), sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *p = 0;
    g_err_flag[device] = p;
  }
  return 0;
}

int b200cd_device_status(int device, void* stream) {
  if (device < 0 || device >= kMaxDevices || g_err_flag[device] == nullptr)
    return fail(B200CD_ERR_ARCH, "b200cd_init(%d) has not been called", device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const cudaError_t e = cudaStreamSynchronize(st);
  const int v = *reinterpret_cast<volatile int*>(g_err_flag[device]);
  if (v != 0)
    return fail(B200CD_ERR_DEVICE,
                "a tensor-core kernel timed out waiting on an mbarrier (device code %d) and trapped; the CUDA context "
                "of this process is no longer usable (%s)", v, cudaGetErrorString(e));
  if (e != cudaSuccess) return fail(B200CD_ERR_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
  return 0;
}

int b200cd_pack_input(const float* src0, const float* src1, int csrc, int c_lo, int nc, int cat_mode, int B, int H,
                      int W, int kpad, void* out, void* stream) {
  const int cin = cat_mode ? 2 * nc : nc;
  if (nc <= 0 || c_lo < 0 || c_lo + nc > csrc || B <= 0 || H <= 0 || W <= 0)
    return fail(B200CD_ERR_SHAPE, "pack_input: bad channel/batch arguments");
  if (kpad % 64 != 0 || 9 * cin > kpad || kpad > 256)
    return fail(B200CD_ERR_SHAPE, "pack_input: kpad=%d must be a multiple of 64 holding 9*Cin=%d", kpad, 9 * cin);
  if (!aligned16(out)) return fail(B200CD_ERR_ALIGN, "pack_input: output not 16-byte aligned");
  CUDA_TRY(b200cd::launch_pack_input(src0, src1, csrc, c_lo, nc, cat_mode, B, H, W, kpad, out,
                                     reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_pack_weights(int mode, const float* w, void* out, int d0, int d1, int kpad, void* stream) {
  if (mode < 0 || mode > 4 || d0 <= 0 || d1 <= 0) return fail(B200CD_ERR_SHAPE, "pack_weights: bad arguments");
  if (mode == 2 && (kpad % 64 != 0 || 9 * d1 > kpad)) return fail(B200CD_ERR_SHAPE, "pack_weights: bad kpad");
  CUDA_TRY(b200cd::launch_pack_weights(mode, w, out, d0, d1, kpad, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_pack_job_blocks(int mode, int d0, int d1, int kpad) {
  if (mode < 0 || mode > 4 || d0 <= 0 || d1 <= 0) return 0;
  return b200cd::pack_job_blocks(mode, d0, d1, kpad);
}

int b200cd_pack_weights_batched(const b200cd_pack_job* jobs_dev, int njobs, int64_t total_blocks, void* stream) {
  static_assert(sizeof(b200cd_pack_job) == sizeof(b200cd::PackJob), "pack job layout");
  if (jobs_dev == nullptr || njobs < 1 || total_blocks < 1 || total_blocks > 0x7fffffffll)
    return fail(B200CD_ERR_SHAPE, "pack_weights_batched: empty job table");
  CUDA_TRY(b200cd::launch_pack_weights_batched(reinterpret_cast<const b200cd::PackJob*>(jobs_dev), njobs, total_blocks,
                                               reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_conv_gemm_tiles(int H, int W) {
  int tw, th;
  tile_shape(W, H, false, &tw, &th);
  return ((W + tw - 1) / tw) * ((H + th - 1) / th);
}

// Tile / kernel selection shared by b200cd_conv_gemm and b200cd_conv_gemm_stat_rows.
struct ConvPlan {
  int tw, th, pair, halo, bn, num_tiles, stat_groups;
};
static ConvPlan plan_conv(int mode, int out_mode, int flags, int n_img, int H, int W, int N) {
  ConvPlan c;
  tile_shape(W, H, mode == 2 || out_mode == 1, &c.tw, &c.th);
  c.pair = ((flags & 4) && ((mode != 1 && out_mode == 0) || mode == 1)) ? 1 : 0;
  c.halo = (mode == 0 && ((flags & 1) || c.pair)) ? 1 : 0;
  c.num_tiles = n_img * ((W + c.tw - 1) / c.tw) * ((H + c.th - 1) / c.th);
  // N tile: 64 when the width is not a multiple of 128; 256 on request (flags bit 1) when it divides the width —
  // a 128 x 256 tile reads 96 B/clk of operands from shared memory per MMA instead of 128 B/clk (the SM's limit).
  // out_mode 1: every 64-channel slab of the tile belongs to one (dy, dx) tap (cout % 64 == 0), so the tile may span taps
  c.bn = (N % 128 == 0) ? 128 : 64;
  if ((flags & 2) && N % 256 == 0 && !c.halo) c.bn = 256;
  if (c.pair && N % 256 == 0) {
    // 256-wide tiles unless they leave most of the 74 CTA pairs idle (deep 16x16 / 32x32 layers at small batch):
    // then 128-wide tiles double the number of work items
    const int items256 = ((c.num_tiles + 1) / 2) * (N / 256);
    c.bn = items256 >= 48 ? 256 : 128;
  }
  // flags bits 5..6: the caller's choice of N tile for the CTA-pair kernel (1 = 64, 2 = 128, 3 = 256; measured per layer
  // shape, multimodal_siamese_cd_b200/tuning.py). Ignored when it does not divide N.
  const int want = (flags >> 5) & 3;
  if (c.pair && want != 0) {
    const int w = 32 << want;
    if (N % w == 0) c.bn = w;
  }
  c.stat_groups = (c.pair && (flags & 8)) ? ((flags >> 8) & 0xff) : 0;
  return c;
}

int b200cd_conv_gemm_stat_rows(int mode, int out_mode, int flags, int n_img, int H, int W, int ka, int N) {
  if (mode < 0 || mode > 2 || n_img <= 0 || H <= 0 || W <= 0 || N < 64 || N % 64 != 0 || ka < 64 || ka % 64 != 0) return -1;
  const ConvPlan c = plan_conv(mode, out_mode, flags, n_img, H, W, N);
  if (c.stat_groups == 0) return c.num_tiles;  // per-tile rows (all groups together)
  if (c.stat_groups > 2 || n_img % c.stat_groups != 0) return -1;
  b200cd::FpropParams p;
  memset(&p, 0, sizeof(p));
  p.mode = mode;
  p.out_mode = out_mode;
  p.taps = mode == 0 ? 9 : (mode == 1 ? 1 : 4);
  p.N = N;
  p.prec = (flags & 16) ? 1 : 0;
  p.kreal = ka / 64;
  p.kchunks = (p.prec ? 3 : 1) * (ka / 64);
  const int ctas = b200cd::fprop_pair_ctas(p, c.bn, c.num_tiles);
  return ctas < 0 ? -1 : 2 * ctas;            // rows per stat-group: one per (CTA, epilogue group)
}

static int conv_gemm_impl(int mode, int out_mode, int flags, const void* A, int64_t a_ld, int n_img, int H, int W, int ka,
                          const void* Bw, int N, int cout, void* out, int64_t out_ld, const float* bias, float* stats,
                          const void* bwd_r, int64_t bwd_ld, const float* bwd_scale, const float* bwd_shift,
                          void* stream, const float* ep_scale = nullptr, const float* ep_shift = nullptr, int ep_relu = 0) {
  if (mode < 0 || mode > 2 || out_mode < 0 || out_mode > 1) return fail(B200CD_ERR_SHAPE, "conv_gemm: bad mode");
  if (ka < 64 || ka % 64 != 0 || N < 64 || N % 64 != 0)
    return fail(B200CD_ERR_SHAPE, "conv_gemm: ka=%d and N=%d must be multiples of 64", ka, N);
  if (a_ld % 8 != 0 || out_ld % 8 != 0 || a_ld < ka) return fail(B200CD_ERR_ALIGN, "conv_gemm: ld must be a multiple of 8");
  if (n_img <= 0 || H <= 0 || W <= 0) return fail(B200CD_ERR_SHAPE, "conv_gemm: empty problem");
  if (out_mode == 1 && (cout < 64 || cout % 64 != 0 || N != 4 * cout))
    return fail(B200CD_ERR_SHAPE, "conv_gemm: scatter epilogue needs N = 4*cout, cout %% 64 == 0");
  if (out_mode == 1 && stats != nullptr) return fail(B200CD_ERR_SHAPE, "conv_gemm: no statistics with the scatter epilogue");
  int* err = nullptr;
  if (int rc = current_err_flag(&err)) return rc;

  const ConvPlan cp = plan_conv(mode, out_mode, flags, n_img, H, W, N);
  const int tw = cp.tw, th = cp.th, pair = cp.pair, halo = cp.halo, bn = cp.bn;
  // flags bit 4: split-bf16 operands. A / out rows hold [hi | lo] halves ld / 2 elements apart, Bw is [N][taps][3 * ka]
  // packed as [hi | lo | hi] per tap (b200cd_pack_weights_hp_batched); ka stays the real channel count.
  const int prec = (flags & 16) ? 1 : 0;
  if (prec && (!pair || bwd_r != nullptr))
    return fail(B200CD_ERR_SHAPE, "conv_gemm: split-bf16 operands need the CTA-pair kernel (flags bit 2) without the fused BN-backward sums");
  if (prec && (a_ld % 16 != 0 || out_ld % 16 != 0 || a_ld / 2 < ka || out_ld / 2 < (out_mode == 1 ? cout : N)))
    return fail(B200CD_ERR_ALIGN, "conv_gemm: split-bf16 rows need ld %% 16 == 0 and ld / 2 >= channels");
  const int kb = prec ? 3 * ka : ka;  // K per tap of the weight matrix
  if (cp.stat_groups > 2 || (cp.stat_groups > 0 && n_img % cp.stat_groups != 0))
    return fail(B200CD_ERR_SHAPE, "conv_gemm: stat groups must be 1 or 2 and divide n_img");
  b200cd::FpropParams p;
  memset(&p, 0, sizeof(p));
  p.mode = mode;
  p.out_mode = out_mode;
  p.taps = mode == 0 ? 9 : (mode == 1 ? 1 : 4);
  p.kchunks = kb / 64;
  p.ka = kb;
  p.prec = prec;
  p.kreal = ka / 64;
  p.a_lo = prec ? static_cast<int>(a_ld / 2) : 0;
  p.o_lo = prec ? static_cast<int>(out_ld / 2) : 0;
  p.tw = tw;
  p.th = th;
  p.rows = halo ? tw * (th + 2) : tw * th;  // rows the A box delivers (expect_tx); the tile itself is tw*th
  p.tiles_x = (W + tw - 1) / tw;
  p.tiles_y = (H + th - 1) / th;
  p.H = H;
  p.W = W;
  p.N = N;
  p.cout = out_mode == 1 ? cout : N;
  p.bias = bias;
  p.stats = reinterpret_cast<float2*>(stats);
  p.ragged = (H % th != 0 || W % tw != 0) ? 1 : 0;
  p.err = err;
  p.n_img = n_img;
  p.stat_groups = cp.stat_groups;
  if (ep_scale != nullptr) {
    if (!pair || stats != nullptr || bwd_r != nullptr || out_mode != 0 || ep_shift == nullptr)
      return fail(B200CD_ERR_SHAPE, "conv_gemm_affine: needs the CTA-pair kernel (flags bit 2), out_mode 0, no statistics");
    p.ep_scale = ep_scale;
    p.ep_shift = ep_shift;
    p.ep_relu = ep_relu;
  }
  if (bwd_r != nullptr) {
    if (!pair || cp.stat_groups == 0 || stats == nullptr || out_mode != 0 || bwd_scale == nullptr || bwd_shift == nullptr)
      return fail(B200CD_ERR_SHAPE, "conv_gemm_bnbwd: needs the CTA-pair kernel with per-CTA statistics (flags bits 2, 3)");
    if (bwd_ld % 2 != 0 || bwd_ld < N || (reinterpret_cast<uintptr_t>(bwd_r) & 3u))
      return fail(B200CD_ERR_ALIGN, "conv_gemm_bnbwd: pre-BN tensor stride / alignment");
    p.bwd_r = bwd_r;
    p.bwd_ld = bwd_ld;
    p.bwd_scale = bwd_scale;
    p.bwd_shift = bwd_shift;
  }
  const int num_tiles = cp.num_tiles;
  if (cp.stat_groups > 0 && stats != nullptr) {
    const int ctas = b200cd::fprop_pair_ctas(p, bn, num_tiles);
    if (ctas < 0) return fail(B200CD_ERR_CUDA, "conv_gemm: occupancy query for the CTA-pair kernel failed");
    p.stat_rows = 2 * ctas;
  }

  CUtensorMap mapA, mapB, mapO;
  int rc;
  const int ca = ka + p.a_lo;  // channel extent of the A map (the lo half lies a_lo channels behind the hi half)
  if (mode == 2) rc = make_up2_map(&mapA, A, a_ld, ca, W, H, n_img, tw, th);
  else rc = make_nhwc_map(&mapA, A, a_ld, ca, W, H, n_img, tw, halo ? th + 2 : th);
  if (rc) return rc;
  {
    const int ka = kb;  // the weight maps below see the tripled K of the split-bf16 layout
    const uint64_t ktot = static_cast<uint64_t>(p.taps) * ka;
    const uint64_t dims[2] = {ktot, (uint64_t)N};
    const uint64_t strides[1] = {ktot * 2};
    const uint32_t box[2] = {64, (uint32_t)(pair ? bn / 2 : bn)};  // pair: each CTA loads half of the weight tile
    if (pair && b200cd::fprop_pair_stacked_weights(mode, bn)) {
      // K index = (ky * 3 + kx) * ka + k: the three ky tiles of one kx as a single box, stacked tile after tile
      const uint64_t dims4[4] = {(uint64_t)ka, (uint64_t)N, 3, 3};
      const uint64_t strides4[3] = {ktot * 2, 3 * (uint64_t)ka * 2, (uint64_t)ka * 2};
      const uint32_t box4[4] = {64, (uint32_t)(bn / 2), 3, 1};
      rc = make_map(&mapB, Bw, 4, dims4, strides4, box4);
    } else {
      rc = make_map(&mapB, Bw, 2, dims, strides, box);
    }
    if (rc) return rc;
  }
  if (out_mode == 1) rc = make_up2_map(&mapO, out, out_ld, cout + p.o_lo, W, H, n_img, tw, th);
  else rc = make_nhwc_map(&mapO, out, out_ld, N + p.o_lo, W, H, n_img, tw, th);
  if (rc) return rc;
  if (pair)
    CUDA_TRY(b200cd::launch_fprop_pair(mapA, mapB, mapO, p, bn, num_tiles, reinterpret_cast<cudaStream_t>(stream)));
  else
    CUDA_TRY(b200cd::launch_fprop(mapA, mapB, mapO, p, bn, halo, num_tiles, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_conv_gemm(int mode, int out_mode, int flags, const void* A, int64_t a_ld, int n_img, int H, int W, int ka,
                     const void* Bw, int N, int cout, void* out, int64_t out_ld, const float* bias, float* stats,
                     void* stream) {
  return conv_gemm_impl(mode, out_mode, flags, A, a_ld, n_img, H, W, ka, Bw, N, cout, out, out_ld, bias, stats, nullptr, 0,
                        nullptr, nullptr, stream);
}

int b200cd_conv_gemm_affine(int mode, int flags, const void* A, int64_t a_ld, int n_img, int H, int W, int ka,
                            const void* Bw, int N, void* out, int64_t out_ld, const float* bias, const float* scale,
                            const float* shift, int relu, void* stream) {
  if (scale == nullptr || shift == nullptr) return fail(B200CD_ERR_SHAPE, "conv_gemm_affine: scale / shift are NULL");
  return conv_gemm_impl(mode, 0, flags, A, a_ld, n_img, H, W, ka, Bw, N, 0, out, out_ld, bias, nullptr, nullptr, 0, nullptr,
                        nullptr, stream, scale, shift, relu);
}

int b200cd_bn_eval_affine_batched(const b200cd_bn_eval_job* jobs_dev, int njobs, int total_blocks, void* stream) {
  static_assert(sizeof(b200cd_bn_eval_job) == sizeof(b200cd::BnEvalJob), "bn eval job layout");
  if (jobs_dev == nullptr || njobs < 1 || total_blocks < 1) return fail(B200CD_ERR_SHAPE, "bn_eval_affine_batched: empty job table");
  CUDA_TRY(b200cd::launch_bn_eval_affine_batched(reinterpret_cast<const b200cd::BnEvalJob*>(jobs_dev), njobs, total_blocks,
                                                 reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_conv_gemm_bnbwd(int mode, int flags, const void* A, int64_t a_ld, int n_img, int H, int W, int ka, const void* Bw,
                           int N, void* out, int64_t out_ld, const void* r, int64_t r_ld, const float* scale,
                           const float* shift, float* sums, void* stream) {
  if (r == nullptr) return fail(B200CD_ERR_SHAPE, "conv_gemm_bnbwd: r is NULL");
  return conv_gemm_impl(mode, 0, flags, A, a_ld, n_img, H, W, ka, Bw, N, 0, out, out_ld, nullptr, sums, r, r_ld, scale, shift,
                        stream);
}

int b200cd_wgrad_tiles(int n_img, int H, int W) { return n_img * ((W + 7) / 8) * ((H + 7) / 8); }

int b200cd_wgrad_ctas_per_split(int mode, int halo, int cu, int cv) {
  if (mode < 0 || mode > 2 || cu < 64 || cu % 64 != 0 || cv < 64 || cv % 64 != 0) return -1;
  const int bn = (cv % 128 == 0) ? 128 : 64;
  const int xy = ((cu + 127) / 128) * (cv / bn);
  if (mode != 0) return xy;
  return (bn == 64 && halo && b200cd::wgrad_mstack(cu)) ? xy : 3 * xy;
}

static int wgrad_gemm_impl(int prec, int mode, int sign, int halo, const void* U, int64_t u_ld, int cu, const void* V,
                           int64_t v_ld, int cv, int n_img, int H, int W, float* ws, int splits, int splits2,
                           int64_t split_stride, int64_t tap_stride, int64_t m_stride, int64_t n_stride, void* stream) {
  if (mode < 0 || mode > 2) return fail(B200CD_ERR_SHAPE, "wgrad_gemm: bad mode");
  if (prec && (u_ld % 16 != 0 || v_ld % 16 != 0 || u_ld / 2 < cu || v_ld / 2 < cv))
    return fail(B200CD_ERR_ALIGN, "wgrad_gemm_hp: split-bf16 rows need ld %% 16 == 0 and ld / 2 >= channels");
  if (cu < 64 || cu % 64 != 0 || cv < 64 || cv % 64 != 0)
    return fail(B200CD_ERR_SHAPE, "wgrad_gemm: cu=%d, cv=%d must be multiples of 64", cu, cv);
  if (u_ld % 8 != 0 || v_ld % 8 != 0) return fail(B200CD_ERR_ALIGN, "wgrad_gemm: ld must be a multiple of 8");
  if (sign != 1 && sign != -1) return fail(B200CD_ERR_SHAPE, "wgrad_gemm: sign must be +1 or -1");
  if (n_stride == 1 && (!aligned16(ws) || split_stride % 4 != 0 || tap_stride % 4 != 0 || m_stride % 4 != 0))
    return fail(B200CD_ERR_ALIGN, "wgrad_gemm: workspace strides must keep 16-byte alignment");
  const int total = b200cd_wgrad_tiles(n_img, H, W);
  if (splits < 1 || splits > total) return fail(B200CD_ERR_SHAPE, "wgrad_gemm: splits=%d outside [1, %d]", splits, total);
  if (splits2 != 0 && (mode != 0 || !halo || cv % 128 == 0 || splits2 < 1 || splits2 > total))
    return fail(B200CD_ERR_SHAPE, "wgrad_gemm: splits2=%d needs mode 0, halo, 64-wide N tiles and 1 <= splits2 <= %d",
                splits2, total);
  int* err = nullptr;
  if (int rc = current_err_flag(&err)) return rc;
  if (mode != 0) halo = 0;

  b200cd::WgradParams p;
  memset(&p, 0, sizeof(p));
  p.mode = mode;
  p.sign = sign;
  p.H = H;
  p.W = W;
  p.tiles_x = (W + 7) / 8;
  p.tiles_y = (H + 7) / 8;
  p.total_tiles = total;
  p.splits = splits;
  p.splits2 = splits2;
  p.cu = cu;
  p.cv = cv;
  p.ws = ws;
  p.split_stride = split_stride;
  p.tap_stride = tap_stride;
  p.m_stride = m_stride;
  p.n_stride = n_stride;
  p.err = err;
  p.passes = prec ? 3 : 1;
  p.u_lo = prec ? static_cast<int>(u_ld / 2) : 0;
  p.v_lo = prec ? static_cast<int>(v_ld / 2) : 0;
  const int bn = (cv % 128 == 0) ? 128 : 64;

  CUtensorMap mapU, mapV;
  int rc;
  if ((rc = make_nhwc_map(&mapU, U, u_ld, cu + p.u_lo, W, H, n_img, 8, 8))) return rc;
  if (mode == 2) rc = make_up2_map(&mapV, V, v_ld, cv + p.v_lo, W, H, n_img, 8, 8);
  else rc = make_nhwc_map(&mapV, V, v_ld, cv + p.v_lo, W, H, n_img, 8, halo ? 10 : 8);
  if (rc) return rc;
  CUDA_TRY(b200cd::launch_wgrad(mapU, mapV, p, bn, halo, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_wgrad_gemm(int mode, int sign, int halo, const void* U, int64_t u_ld, int cu, const void* V, int64_t v_ld,
                      int cv, int n_img, int H, int W, float* ws, int splits, int splits2, int64_t split_stride,
                      int64_t tap_stride, int64_t m_stride, int64_t n_stride, void* stream) {
  return wgrad_gemm_impl(0, mode, sign, halo, U, u_ld, cu, V, v_ld, cv, n_img, H, W, ws, splits, splits2, split_stride,
                         tap_stride, m_stride, n_stride, stream);
}

int b200cd_wgrad_gemm_hp(int mode, int sign, int halo, const void* U, int64_t u_ld, int cu, const void* V, int64_t v_ld,
                         int cv, int n_img, int H, int W, float* ws, int splits, int splits2, int64_t split_stride,
                         int64_t tap_stride, int64_t m_stride, int64_t n_stride, void* stream) {
  return wgrad_gemm_impl(1, mode, sign, halo, U, u_ld, cu, V, v_ld, cv, n_img, H, W, ws, splits, splits2, split_stride,
                         tap_stride, m_stride, n_stride, stream);
}

int b200cd_wgrad_reduce(const float* ws, int splits, int64_t split_stride, int layout, int d0, int d1, int taps,
                        float* grad, void* stream) {
  if (splits < 1 || d0 <= 0 || d1 <= 0 || taps <= 0 || layout < 0 || layout > 1)
    return fail(B200CD_ERR_SHAPE, "wgrad_reduce: bad arguments");
  CUDA_TRY(b200cd::launch_wgrad_reduce(ws, splits, split_stride, layout, d0, d1, taps, grad,
                                       reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_bn_stats(const float* partial, int ld, int C, int tiles_per_group, int G, double count, int spl, double* ws,
                    const float* gamma, const float* beta, float* running_mean, float* running_var, int64_t* nbt,
                    float momentum, float eps, int train, int order_rev, float* mean, float* invstd, float* scale,
                    float* shift, void* stream) {
  if (C <= 0 || G <= 0) return fail(B200CD_ERR_SHAPE, "bn_stats: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (train) {
    if (spl < 1 || tiles_per_group < 1 || partial == nullptr || ws == nullptr)
      return fail(B200CD_ERR_SHAPE, "bn_stats: training mode needs partial statistics and a workspace");
    if (tiles_per_group <= 1024) {  // few partial rows (per-CTA statistics): one kernel does reduce + finalize
      CUDA_TRY(b200cd::launch_bn_stats_fused(reinterpret_cast<const float2*>(partial), ld, tiles_per_group, C, G, count,
                                             gamma, beta, running_mean, running_var, reinterpret_cast<long long*>(nbt),
                                             momentum, eps, order_rev, mean, invstd, scale, shift, st));
      return 0;
    }
    if (spl > tiles_per_group) spl = tiles_per_group;
    CUDA_TRY(b200cd::launch_bn_stats_reduce(reinterpret_cast<const float2*>(partial), ld, C, tiles_per_group, G, spl,
                                            ws, st));
  }
  CUDA_TRY(b200cd::launch_bn_finalize(ws, spl, C, G, count, gamma, beta, running_mean, running_var,
                                      reinterpret_cast<long long*>(nbt), momentum, eps, train, order_rev, mean, invstd,
                                      scale, shift, st));
  return 0;
}

int b200cd_bn_apply(const void* r, int64_t ld_r, const float* scale, const float* shift, int n_img, int H, int W,
                    int C, int G, int diff, void* a, int64_t ld_a, void* a2, int64_t ld_a2, void* pool, int64_t ld_p,
                    void* dif, int64_t ld_d, void* pool_idx, void* stream) {
  if (C % 8 != 0 || n_img % G != 0 || (diff && (n_img % 2 != 0 || G != 2)))
    return fail(B200CD_ERR_SHAPE, "bn_apply: bad C/G/diff combination");
  if (ld_r % 8 || (a && ld_a % 8) || (a2 && ld_a2 % 8) || (pool && ld_p % 8) || (dif && ld_d % 8))
    return fail(B200CD_ERR_ALIGN, "bn_apply: ld must be a multiple of 8");
  if (!aligned16(r) || !aligned16(a) || !aligned16(a2) || !aligned16(pool) || !aligned16(dif))
    return fail(B200CD_ERR_ALIGN, "bn_apply: pointers must be 16-byte aligned");
  if (pool_idx != nullptr && (pool == nullptr || (reinterpret_cast<uintptr_t>(pool_idx) & 7u)))
    return fail(B200CD_ERR_ALIGN, "bn_apply: pool_idx needs pool and 8-byte alignment");
  CUDA_TRY(b200cd::launch_bn_apply(r, ld_r, scale, shift, n_img, H, W, C, G, diff, a, ld_a, a2, ld_a2, pool, ld_p, dif,
                                   ld_d, pool_idx, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

size_t b200cd_bn_bwd_ws_floats(int n_img, int H, int W, int C, int G) {
  if (!chan_ok(C) || G <= 0 || n_img % G != 0) return 0;
  return static_cast<size_t>(2) * G * C * (bn_bwd_nblk(n_img, H, W, C, G) + 1);
}

static int bn_bwd_impl(const void* r, int64_t ld_r, const float* mean, const float* invstd, const float* scale,
                       const float* shift, const b200cd_grad_src* srcs, int n_img, int H, int W, int C, int G, float* ws,
                       float* dgamma, float* dbeta, void* dr, int64_t ld_dr, const float* sums, int sum_rows, void* stream);

int b200cd_bn_bwd(const void* r, int64_t ld_r, const float* mean, const float* invstd, const float* scale,
                  const float* shift, const b200cd_grad_src* srcs, int n_img, int H, int W, int C, int G, float* ws,
                  float* dgamma, float* dbeta, void* dr, int64_t ld_dr, void* stream) {
  return bn_bwd_impl(r, ld_r, mean, invstd, scale, shift, srcs, n_img, H, W, C, G, ws, dgamma, dbeta, dr, ld_dr, nullptr, 0,
                     stream);
}

int b200cd_bn_bwd_from_sums(const void* r, int64_t ld_r, const float* mean, const float* invstd, const float* scale,
                            const float* shift, const b200cd_grad_src* srcs, int n_img, int H, int W, int C, int G,
                            const float* sums, int sum_rows, float* ws, float* dgamma, float* dbeta, void* dr,
                            int64_t ld_dr, void* stream) {
  if (sums == nullptr || sum_rows < 1) return fail(B200CD_ERR_SHAPE, "bn_bwd_from_sums: no partial sums");
  return bn_bwd_impl(r, ld_r, mean, invstd, scale, shift, srcs, n_img, H, W, C, G, ws, dgamma, dbeta, dr, ld_dr, sums,
                     sum_rows, stream);
}

static int bn_bwd_impl(const void* r, int64_t ld_r, const float* mean, const float* invstd, const float* scale,
                       const float* shift, const b200cd_grad_src* srcs, int n_img, int H, int W, int C, int G, float* ws,
                       float* dgamma, float* dbeta, void* dr, int64_t ld_dr, const float* sums, int sum_rows, void* stream) {
  if (!chan_ok(C)) return fail(B200CD_ERR_SHAPE, "bn_bwd: C=%d must be a multiple of 64 with 256 %% (C/8) == 0", C);
  if (G <= 0 || n_img % G != 0) return fail(B200CD_ERR_SHAPE, "bn_bwd: n_img must be a multiple of G");
  if (ld_r % 8 || ld_dr % 8 || !aligned16(r) || !aligned16(dr)) return fail(B200CD_ERR_ALIGN, "bn_bwd: alignment");
  b200cd::GradSrcs gs;
  memset(&gs, 0, sizeof(gs));
  if (static_cast<long long>(n_img) * H * W >= (1ll << 31)) return fail(B200CD_ERR_SHAPE, "bn_bwd: more than 2^31 pixels");
  for (int i = 0; i < 3; ++i) {
    gs.s[i].kind = srcs[i].kind;
    gs.s[i].ptr = srcs[i].ptr;
    gs.s[i].w = srcs[i].w;
    gs.s[i].ld = srcs[i].ld;
    gs.s[i].n_mod = srcs[i].n_mod;
    gs.s[i].scale_lo = srcs[i].scale_lo;
    gs.s[i].scale_hi = srcs[i].scale_hi;
    if (srcs[i].kind < 0 || srcs[i].kind > 3) return fail(B200CD_ERR_SHAPE, "bn_bwd: bad gradient source kind");
    if ((srcs[i].kind == 1 || srcs[i].kind == 2) && (srcs[i].ld % 8 != 0 || !aligned16(srcs[i].ptr)))
      return fail(B200CD_ERR_ALIGN, "bn_bwd: gradient source %d alignment", i);
    if (srcs[i].kind == 2 && srcs[i].w == nullptr)
      return fail(B200CD_ERR_SHAPE, "bn_bwd: a max-pool source needs the arg-max index tensor in .w");
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nblk = bn_bwd_nblk(n_img, H, W, C, G);
  float* partial = ws;
  float* coefA = ws + static_cast<size_t>(2) * G * C * nblk;
  float* coefB = coefA + static_cast<size_t>(G) * C;
  const double count = static_cast<double>(n_img / G) * H * W;
  if (sums != nullptr) {  // S1, S2 already reduced per CTA by the convolution that produced the gradient
    CUDA_TRY(b200cd::launch_bn_bwd_finalize(sums, sum_rows, C, G, count, mean, invstd, scale, dgamma, dbeta, coefA, coefB, st));
  } else {
    CUDA_TRY(b200cd::launch_bn_bwd_reduce(r, ld_r, scale, shift, gs, n_img, H, W, C, G, nblk, partial, st));
    CUDA_TRY(b200cd::launch_bn_bwd_finalize(partial, nblk, C, G, count, mean, invstd, scale, dgamma, dbeta, coefA, coefB, st));
  }
  CUDA_TRY(b200cd::launch_bn_bwd_dx(r, ld_r, scale, shift, coefA, coefB, gs, n_img, H, W, C, G, nblk, dr, ld_dr, st));
  return 0;
}

int b200cd_head_fwd(const void* a0, int64_t ld0, const void* a1, int64_t ld1, int C, const float* w, const float* b,
                    int64_t npix, float* logits, void* stream) {
  if (C % 64 != 0 || C <= 0) return fail(B200CD_ERR_SHAPE, "head_fwd: C must be a multiple of 64");
  if (ld0 % 8 || (a1 && ld1 % 8) || !aligned16(a0) || !aligned16(a1)) return fail(B200CD_ERR_ALIGN, "head_fwd: alignment");
  CUDA_TRY(b200cd::launch_head_fwd(a0, ld0, a1, ld1, C, w, b, npix, logits, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_pad_copy(const void* src, int64_t ld_src, int n_img, int h, int w, int C, void* dst, int64_t ld_dst, int H,
                    int W, int top, int left, void* stream) {
  if (n_img <= 0 || h <= 0 || w <= 0 || C <= 0 || C % 8 != 0 || top < 0 || left < 0 || top + h > H || left + w > W)
    return fail(B200CD_ERR_SHAPE, "pad_copy: %dx%d at (%d, %d) does not fit %dx%d (C=%d)", h, w, top, left, H, W, C);
  if (ld_src % 8 != 0 || ld_dst % 8 != 0 || ld_src < C || ld_dst < C || !aligned16(src) || !aligned16(dst))
    return fail(B200CD_ERR_ALIGN, "pad_copy: strides must be multiples of 8 elements and pointers 16-byte aligned");
  CUDA_TRY(b200cd::launch_pad_copy(src, ld_src, n_img, h, w, C, dst, ld_dst, H, W, top, left,
                                   reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_colsum(const void* x, int64_t ld, int C, const float* wgt, int64_t npix, int nblk, float* ws, float* out,
                  void* stream) {
  if (x != nullptr && !chan_ok(C)) return fail(B200CD_ERR_SHAPE, "colsum: unsupported C=%d", C);
  if (x == nullptr && (C != 1 || wgt == nullptr)) return fail(B200CD_ERR_SHAPE, "colsum: x == NULL needs C == 1 and wgt");
  if (nblk < 1 || npix < 1) return fail(B200CD_ERR_SHAPE, "colsum: empty");
  if (x != nullptr && (ld % 8 || !aligned16(x))) return fail(B200CD_ERR_ALIGN, "colsum: alignment");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUDA_TRY(b200cd::launch_colsum(x, ld, C, wgt, npix, nblk, ws, st));
  CUDA_TRY(b200cd::launch_colsum_finalize(ws, nblk, C, out, st));
  return 0;
}

int b200cd_stat_rowsum(const float* stats, int rows, int ld, int c_off, int C, float* out, void* stream) {
  if (stats == nullptr || rows < 1 || C < 1 || c_off < 0 || c_off + C > ld) return fail(B200CD_ERR_SHAPE, "stat_rowsum: bad arguments");
  CUDA_TRY(b200cd::launch_stat_rowsum(reinterpret_cast<const float2*>(stats), rows, ld, c_off, C, out,
                                      reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_pj_fwd(const float* z, const float* t, int t_is_logit, const unsigned char* rowmask, int sel, int rows,
                  int64_t per_row, int nblk, double* ws, double* sums, void* stream) {
  if (rows < 1 || per_row < 4 || per_row % 4 != 0 || nblk < 1)
    return fail(B200CD_ERR_SHAPE, "pj_fwd: per_row must be a positive multiple of 4");
  if (!aligned16(z) || !aligned16(t)) return fail(B200CD_ERR_ALIGN, "pj_fwd: alignment");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUDA_TRY(b200cd::launch_pj_reduce(z, t, t_is_logit, rowmask, sel, rows, per_row, nblk, ws, st));
  CUDA_TRY(b200cd::launch_pj_finalize(ws, nblk, sums, st));
  return 0;
}

int b200cd_pj_loss(const double* sums, float* loss, void* stream) {
  CUDA_TRY(b200cd::launch_pj_loss(sums, loss, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_pj_bwd(const float* z, const float* t, int t_is_logit, const unsigned char* rowmask, int sel, int rows,
                  int64_t per_row, const double* sums, const float* gptr, float gmul, int accumulate, float* dz,
                  float* dt, void* stream) {
  if (rows < 1 || per_row < 4 || per_row % 4 != 0) return fail(B200CD_ERR_SHAPE, "pj_bwd: per_row must be a multiple of 4");
  if (!aligned16(z) || !aligned16(t) || !aligned16(dz) || !aligned16(dt)) return fail(B200CD_ERR_ALIGN, "pj_bwd: alignment");
  CUDA_TRY(b200cd::launch_pj_bwd(z, t, t_is_logit, rowmask, sel, rows, per_row, sums, gptr, gmul, accumulate, dz, dt,
                                 reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_confusion_counts(const float* pred, const float* truth, int64_t n, int from_logits, const float* thresholds,
                            int nthr, uint64_t* counts, void* stream) {
  if (pred == nullptr || truth == nullptr || thresholds == nullptr || counts == nullptr || n < 1 || nthr < 1 || nthr > 8)
    return fail(B200CD_ERR_SHAPE, "confusion_counts: 1..8 thresholds and a non-empty input are required");
  CUDA_TRY(b200cd::launch_confusion(pred, truth, n, from_logits, thresholds, nthr,
                                    reinterpret_cast<unsigned long long*>(counts), reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_augment(const b200cd_augment_job* jobs_dev, int n, int crop, int cout, float* out, void* stream) {
  static_assert(sizeof(b200cd_augment_job) == sizeof(b200cd::AugmentJob), "augment job layout");
  if (jobs_dev == nullptr || n < 1 || crop < 1 || cout < 1 || cout > 16 || out == nullptr)
    return fail(B200CD_ERR_SHAPE, "augment: 1..16 output channels and a non-empty batch are required");
  CUDA_TRY(b200cd::launch_augment(reinterpret_cast<const b200cd::AugmentJob*>(jobs_dev), n, crop, cout, out,
                                  reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int64_t b200cd_query_workspace(int op, const int64_t* d, int nd) {
  // bytes of caller-allocated workspace / side-output buffers, per op (see include/b200cd.h for the dims of each op)
  if (d == nullptr) return -1;
  switch (op) {
    case B200CD_WS_BN_BWD:        // n_img, H, W, C, G -> fp32 workspace of b200cd_bn_bwd*
      return nd == 5 ? static_cast<int64_t>(4 * b200cd_bn_bwd_ws_floats((int)d[0], (int)d[1], (int)d[2], (int)d[3], (int)d[4])) : -1;
    case B200CD_WS_CONV_STATS: {  // mode, out_mode, flags, n_img, H, W, ka, N -> fp32 [groups][rows][N][2] statistics buffer
      if (nd != 8) return -1;
      const int rows = b200cd_conv_gemm_stat_rows((int)d[0], (int)d[1], (int)d[2], (int)d[3], (int)d[4], (int)d[5], (int)d[6], (int)d[7]);
      if (rows < 0) return -1;
      const int groups = ((int)d[2] & 8) ? (((int)d[2] >> 8) & 0xff) : 1;
      return static_cast<int64_t>(groups) * rows * d[7] * 2 * 4;
    }
    case B200CD_WS_WGRAD:         // splits, taps, cu, cv -> fp32 split partials of b200cd_wgrad_gemm*
      return nd == 4 ? 4 * d[0] * d[1] * d[2] * d[3] : -1;
    case B200CD_WS_COLSUM:        // nblk, C -> fp32 block partials of b200cd_colsum*
      return nd == 2 ? 4 * d[0] * d[1] : -1;
    case B200CD_WS_PJ:            // nblk -> fp64 block partials of b200cd_pj_fwd
      return nd == 1 ? 8 * 3 * d[0] : -1;
    case B200CD_WS_BN_STATS:      // spl, G, C -> fp64 second-stage partials of b200cd_bn_stats
      return nd == 3 ? 8 * 2 * d[0] * d[1] * d[2] : -1;
    default:
      return -1;
  }
}

int b200cd_graph_instantiate(void* cuda_graph, int use_node_priority, void** exec_out) {
  if (cuda_graph == nullptr || exec_out == nullptr) return fail(B200CD_ERR_SHAPE, "graph_instantiate: NULL graph / output");
  cudaGraphExec_t exec = nullptr;
  const unsigned long long flags = use_node_priority ? cudaGraphInstantiateFlagUseNodePriority : 0ull;
  CUDA_TRY(cudaGraphInstantiateWithFlags(&exec, reinterpret_cast<cudaGraph_t>(cuda_graph), flags));
  *exec_out = exec;
  return 0;
}

int b200cd_graph_launch(void* graph_exec, void* stream) {
  if (graph_exec == nullptr) return fail(B200CD_ERR_SHAPE, "graph_launch: NULL executable graph");
  CUDA_TRY(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_graph_exec_destroy(void* graph_exec) {
  if (graph_exec == nullptr) return 0;
  CUDA_TRY(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph_exec)));
  return 0;
}

int b200cd_reduce_job_parts(int splits, int d1, int taps) {
  return (splits < 1 || d1 < 4 || taps < 1) ? -1 : b200cd::reduce_job_parts(splits, d1, taps);
}

int64_t b200cd_reduce_job_blocks(int splits, int d0, int d1, int taps) {
  if (splits < 1 || d0 < 1 || d1 < 4 || d1 % 4 != 0 || taps < 1) return 0;
  return b200cd::reduce_job_blocks(splits, d0, d1, taps);
}

int b200cd_wgrad_reduce_batched(const b200cd_reduce_job* jobs_dev, int njobs, int64_t total_blocks, void* stream) {
  static_assert(sizeof(b200cd_reduce_job) == sizeof(b200cd::ReduceJob), "reduce job layout");
  if (jobs_dev == nullptr || njobs < 1 || total_blocks < 1 || total_blocks > 0x7fffffffll)
    return fail(B200CD_ERR_SHAPE, "wgrad_reduce_batched: empty job table");
  CUDA_TRY(b200cd::launch_wgrad_reduce_batched(reinterpret_cast<const b200cd::ReduceJob*>(jobs_dev), njobs, total_blocks,
                                               reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_adamw_step(const b200cd_adamw_job* jobs_dev, int njobs, int64_t total_blocks, double lr, double beta1,
                      double beta2, double eps, double weight_decay, int64_t step_count, void* stream) {
  static_assert(sizeof(b200cd_adamw_job) == sizeof(b200cd::AdamWJob), "adamw job layout");
  if (jobs_dev == nullptr || njobs < 1 || total_blocks < 1 || total_blocks > 0x7fffffffll || step_count < 1)
    return fail(B200CD_ERR_SHAPE, "adamw_step: empty job table or step_count < 1");
  if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0))
    return fail(B200CD_ERR_SHAPE, "adamw_step: betas must be in [0, 1) and eps >= 0");
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step_count));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step_count));
  CUDA_TRY(b200cd::launch_adamw(reinterpret_cast<const b200cd::AdamWJob*>(jobs_dev), njobs, total_blocks,
                                static_cast<float>(1.0 - lr * weight_decay), static_cast<float>(lr / bc1),
                                static_cast<float>(1.0 - beta1), static_cast<float>(beta2),
                                static_cast<float>(1.0 - beta2), static_cast<float>(eps), static_cast<float>(sqrt(bc2)),
                                reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// split-bf16 ("precise") entry points (include/b200cd.h, ABI version 2)
// ------------------------------------------------------------------------------------------------
static bool split_ok(const void* p, int64_t ld, int C) { return p == nullptr || (ld % 16 == 0 && ld / 2 >= C && aligned16(p)); }

int b200cd_pack_input_hp(const float* src0, const float* src1, int csrc, int c_lo, int nc, int cat_mode, int B, int H,
                         int W, int kpad, void* out, void* stream) {
  const int cin = cat_mode ? 2 * nc : nc;
  if (nc <= 0 || c_lo < 0 || c_lo + nc > csrc || B <= 0 || H <= 0 || W <= 0)
    return fail(B200CD_ERR_SHAPE, "pack_input_hp: bad channel/batch arguments");
  if (kpad % 64 != 0 || 9 * cin > kpad || kpad > 256)
    return fail(B200CD_ERR_SHAPE, "pack_input_hp: kpad=%d must be a multiple of 64 holding 9*Cin=%d", kpad, 9 * cin);
  if (!aligned16(out)) return fail(B200CD_ERR_ALIGN, "pack_input_hp: output not 16-byte aligned");
  CUDA_TRY(b200cd::launch_pack_input_hp(src0, src1, csrc, c_lo, nc, cat_mode, B, H, W, kpad, out,
                                        reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_pack_job_blocks_hp(int mode, int d0, int d1, int kpad) {
  if (mode < 0 || mode > 4 || d0 <= 0 || d1 <= 0) return 0;
  return b200cd::pack_job_blocks_hp(mode, d0, d1, kpad);
}

int b200cd_pack_weights_hp_batched(const b200cd_pack_job* jobs_dev, int njobs, int64_t total_blocks, void* stream) {
  if (jobs_dev == nullptr || njobs < 1 || total_blocks < 1 || total_blocks > 0x7fffffffll)
    return fail(B200CD_ERR_SHAPE, "pack_weights_hp_batched: empty job table");
  CUDA_TRY(b200cd::launch_pack_weights_hp_batched(reinterpret_cast<const b200cd::PackJob*>(jobs_dev), njobs, total_blocks,
                                                  reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_bn_apply_hp(const void* r, int64_t ld_r, const float* scale, const float* shift, int n_img, int H, int W,
                       int C, int G, int diff, void* a, int64_t ld_a, void* a2, int64_t ld_a2, void* pool, int64_t ld_p,
                       void* dif, int64_t ld_d, void* pool_idx, void* stream) {
  if (C % 8 != 0 || n_img % G != 0 || (diff && (n_img % 2 != 0 || G != 2)))
    return fail(B200CD_ERR_SHAPE, "bn_apply_hp: bad C/G/diff combination");
  if (!split_ok(r, ld_r, C) || !split_ok(a, ld_a, C) || !split_ok(a2, ld_a2, C) || !split_ok(pool, ld_p, C) ||
      !split_ok(dif, ld_d, C))
    return fail(B200CD_ERR_ALIGN, "bn_apply_hp: split tensors need ld %% 16 == 0, ld / 2 >= C and 16-byte alignment");
  if (pool_idx != nullptr && (pool == nullptr || (reinterpret_cast<uintptr_t>(pool_idx) & 7u)))
    return fail(B200CD_ERR_ALIGN, "bn_apply_hp: pool_idx needs pool and 8-byte alignment");
  CUDA_TRY(b200cd::launch_bn_apply_hp(r, ld_r, scale, shift, n_img, H, W, C, G, diff, a, ld_a, a2, ld_a2, pool, ld_p, dif,
                                      ld_d, pool_idx, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_bn_bwd_hp(const void* r, int64_t ld_r, const float* mean, const float* invstd, const float* scale,
                     const float* shift, const b200cd_grad_src* srcs, int n_img, int H, int W, int C, int G, float* ws,
                     float* dgamma, float* dbeta, void* dr, int64_t ld_dr, void* stream) {
  if (!chan_ok(C)) return fail(B200CD_ERR_SHAPE, "bn_bwd_hp: C=%d must be a multiple of 64 with 256 %% (C/8) == 0", C);
  if (G <= 0 || n_img % G != 0) return fail(B200CD_ERR_SHAPE, "bn_bwd_hp: n_img must be a multiple of G");
  if (!split_ok(r, ld_r, C) || !split_ok(dr, ld_dr, C)) return fail(B200CD_ERR_ALIGN, "bn_bwd_hp: alignment");
  if (static_cast<long long>(n_img) * H * W >= (1ll << 31)) return fail(B200CD_ERR_SHAPE, "bn_bwd_hp: more than 2^31 pixels");
  b200cd::GradSrcs gs;
  memset(&gs, 0, sizeof(gs));
  for (int i = 0; i < 3; ++i) {
    gs.s[i].kind = srcs[i].kind;
    gs.s[i].ptr = srcs[i].ptr;
    gs.s[i].w = srcs[i].w;
    gs.s[i].ld = srcs[i].ld;
    gs.s[i].n_mod = srcs[i].n_mod;
    gs.s[i].scale_lo = srcs[i].scale_lo;
    gs.s[i].scale_hi = srcs[i].scale_hi;
    if (srcs[i].kind < 0 || srcs[i].kind > 3) return fail(B200CD_ERR_SHAPE, "bn_bwd_hp: bad gradient source kind");
    if ((srcs[i].kind == 1 || srcs[i].kind == 2) && !split_ok(srcs[i].ptr, srcs[i].ld, C))
      return fail(B200CD_ERR_ALIGN, "bn_bwd_hp: gradient source %d alignment", i);
    if (srcs[i].kind == 2 && srcs[i].w == nullptr)
      return fail(B200CD_ERR_SHAPE, "bn_bwd_hp: a max-pool source needs the arg-max index tensor in .w");
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nblk = bn_bwd_nblk(n_img, H, W, C, G);
  float* partial = ws;
  float* coefA = ws + static_cast<size_t>(2) * G * C * nblk;
  float* coefB = coefA + static_cast<size_t>(G) * C;
  const double count = static_cast<double>(n_img / G) * H * W;
  CUDA_TRY(b200cd::launch_bn_bwd_reduce_hp(r, ld_r, scale, shift, gs, n_img, H, W, C, G, nblk, partial, st));
  CUDA_TRY(b200cd::launch_bn_bwd_finalize(partial, nblk, C, G, count, mean, invstd, scale, dgamma, dbeta, coefA, coefB, st));
  CUDA_TRY(b200cd::launch_bn_bwd_dx_hp(r, ld_r, scale, shift, coefA, coefB, gs, n_img, H, W, C, G, nblk, dr, ld_dr, st));
  return 0;
}

int b200cd_head_fwd_hp(const void* a0, int64_t ld0, const void* a1, int64_t ld1, int C, const float* w, const float* b,
                       int64_t npix, float* logits, void* stream) {
  if (C % 64 != 0 || C <= 0) return fail(B200CD_ERR_SHAPE, "head_fwd_hp: C must be a multiple of 64");
  if (!split_ok(a0, ld0, C) || !split_ok(a1, ld1, C)) return fail(B200CD_ERR_ALIGN, "head_fwd_hp: alignment");
  CUDA_TRY(b200cd::launch_head_fwd_hp(a0, ld0, a1, ld1, C, w, b, npix, logits, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int b200cd_colsum_hp(const void* x, int64_t ld, int C, const float* wgt, int64_t npix, int nblk, float* ws, float* out,
                     void* stream) {
  if (x == nullptr || !chan_ok(C)) return fail(B200CD_ERR_SHAPE, "colsum_hp: needs x and a supported C (got %d)", C);
  if (nblk < 1 || npix < 1) return fail(B200CD_ERR_SHAPE, "colsum_hp: empty");
  if (!split_ok(x, ld, C)) return fail(B200CD_ERR_ALIGN, "colsum_hp: alignment");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUDA_TRY(b200cd::launch_colsum_hp(x, ld, C, wgt, npix, nblk, ws, st));
  CUDA_TRY(b200cd::launch_colsum_finalize(ws, nblk, C, out, st));
  return 0;
}

}  // extern "C"
