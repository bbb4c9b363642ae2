// G2 — weight-gradient GEMM on tcgen05 / TMEM (sm_100a only).
//
// dW_tap[m, n] = sum over pixels of U[pixel, m] * V_tap[pixel, n].  The reduction dimension is the pixel
// index, and in NHWC memory the channel index is the contiguous one, so both operands are "MN-major":
// a TMA box of 64 pixels x 64 channels lands in shared memory as 64 rows of 128 bytes (128-byte swizzle)
// and is consumed as-is by tcgen05.mma with the MN-major bits set in the instruction descriptor.
// One CTA owns a 128 x BN block of the weight gradient for a group of taps (one accumulator per tap in
// TMEM) and a contiguous range of 8x8 pixel tiles (split over the pixel dimension); per-split results go
// to a workspace that wgrad_reduce sums in a fixed order (deterministic, no float atomics).
//
// For 3x3 convolutions a CTA handles the three ky taps of one kx. With HALO the shifted operand is
// loaded once per pixel tile as an 8 x 10 box (one halo row above and below); since a shift by one image
// row is a shift by 8 shared-memory rows = one 1024-byte swizzle atom, the three ky operands are plain
// address offsets into that box.
//
// Replaces (reference): the weight gradients autograd computes for nn.Conv2d utils/networks.py:392,395
// and nn.ConvTranspose2d utils/networks.py:433.
#include "kernels.h"
#include "ptx.cuh"

namespace b200cd {

namespace {

constexpr int kThreads = 224;  // U producer, MMA issuer, 4 epilogue warps, V producer (B200CD_WGRAD_SPLIT=0: 192, warp 0 loads both)
constexpr int kUBytes = 2 * 64 * 128;  // 64 pixels x 128 channels (two 64-channel slabs)

// NKX (3x3 mode with HALO only): kx columns per CTA. With 64-wide N tiles an MMA lasts 32 cycles and a (U, V box) pair
// of 26 KB feeds only 12 of them — 68 B/clk against the ~40 B/clk the L2 -> SM fabric delivers. NKX = 2 lets one U tile
// feed two V boxes (6 accumulators, 47 B/clk): CTAs z < splits own kx 0 and 1 over 1/splits of the pixel tiles, CTAs
// z >= splits own kx 2 over 1/splits2 of them (splits2 = splits / 2 balances the work).
template <int BN, int MODE, bool HALO, int NKX = 1>
struct WgradCfg {
  // NKX = 3 (64-wide N tiles, U <= 64 channels): ONE CTA owns all nine taps. The kx shift moves to the U side — three
  // 64-channel U tiles at x - sign, x, x + sign — and two of them stack along M (the M = 128 tile would otherwise be
  // half empty); the three ky taps stack along N out of one unshifted V halo box (N = 192).
  static constexpr int kTaps = MODE == 0 ? 3 * NKX : (MODE == 1 ? 1 : 4);
  static constexpr int kVRows = HALO ? 80 : 64;
  static constexpr int kVSlab = kVRows * 128;
  static constexpr int kVTiles = NKX == 3 ? 1 : (HALO ? NKX : kTaps);  // separately loaded tap tiles / boxes
  static constexpr int kVBytes = kVTiles * (BN / 64) * kVSlab;
  static constexpr int kUBytesC = NKX == 3 ? 3 * 64 * 128 : kUBytes;
  static constexpr int kStageBytes = kUBytesC + kVBytes;
  static constexpr int kStagesRaw = (196 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kCols = NKX == 3 ? 384 : kTaps * BN;
  static constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kTmemSlotOff = kBarOff + 8 * (2 * kStages + 1);
  static constexpr int kTotal = kTmemSlotOff + 16;
  static constexpr int kDynamic = kTotal + 1024;
  static_assert(kStages >= 2, "need at least a double buffer");
  static_assert(kCols <= 512, "accumulators exceed TMEM");
};

#ifdef B200CD_TRACE
__device__ long long g_trace_w[4][4096];
#define WTRACE_DECL(role)                                           \
  int tr_i = 0;                                                     \
  const bool tr_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0; \
  const int tr_role = (role);
#define WTR()                                                                   \
  do {                                                                          \
    if (tr_on && tr_i < 4096) g_trace_w[tr_role][tr_i++] = clock64();           \
  } while (0)
#else
#define WTRACE_DECL(role)
#define WTR()
#endif

}  // namespace

// 3x3 / halo / 64-wide N tiles with at most 64 U channels: one CTA per pixel range owns all nine taps (grid z = splits,
// not 3 * splits) — see WgradCfg. B200CD_WGRAD_MSTACK=0 keeps the one-kx-per-CTA kernel.
bool wgrad_mstack(int cu) {
  static const bool on = [] {
    const char* e = getenv("B200CD_WGRAD_MSTACK");
    return !(e && e[0] == '0');
  }();
  return on && cu <= 64;
}

namespace {

static int wgrad_threads() {
  static const int t = [] {
    const char* e = getenv("B200CD_WGRAD_SPLIT");
    return (e && e[0] == '0') ? 192 : kThreads;
  }();
  return t;
}

template <int BN, int MODE, bool HALO, int NKX>
__global__ void __launch_bounds__(kThreads) wgrad_kernel(const __grid_constant__ CUtensorMap mapU,
                                                         const __grid_constant__ CUtensorMap mapV,
                                                         const WgradParams p) {
  using C = WgradCfg<BN, MODE, HALO, NKX>;
  static_assert(NKX == 1 || (MODE == 0 && HALO), "NKX > 1 needs the 3x3 halo variant");
  static_assert(NKX != 3 || BN == 64, "the M-stacked variant is for 64-wide N tiles");
  constexpr int STAGES = C::kStages;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* accbar = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kTmemSlotOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int n0 = blockIdx.y * BN;
  int split, kx, nkx, nsplit;  // this CTA: kx columns [kx, kx + nkx) over pixel-tile range split / nsplit
  if (NKX == 3) {
    split = blockIdx.z;
    kx = 0;
    nkx = 3;
    nsplit = p.splits;
  } else if (NKX == 2) {
    const bool first = static_cast<int>(blockIdx.z) < p.splits;
    split = first ? blockIdx.z : blockIdx.z - p.splits;
    nsplit = first ? p.splits : p.splits2;
    kx = first ? 0 : 2;
    nkx = first ? 2 : 1;
  } else {
    split = blockIdx.z % p.splits;
    kx = blockIdx.z / p.splits;  // mode 0 only (0..2)
    nkx = 1;
    nsplit = p.splits;
  }
  const int t_begin = static_cast<int>(static_cast<long long>(p.total_tiles) * split / nsplit);
  const int t_end = static_cast<int>(static_cast<long long>(p.total_tiles) * (split + 1) / nsplit);
  const int iters = (t_end - t_begin) * p.passes;  // split-bf16 operands: every pixel tile runs p.passes = 3 times

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapU);
    tma_prefetch_desc(&mapV);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accbar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above is independent of the predecessor kernel's output

  if (warp == 0 || warp == 6) {
    // ---------------- TMA producers (whole warp runs the loop, one elected lane issues): warp 0 arms the stage's
    // barrier and loads U, warp 6 loads V — a thread can start a bulk-tensor load only every ~230 cycles
    // (tools/ubench/tma_rate.cu), so the four loads of a stage are spread over two issuing warps ----------------
    const bool split_prod = blockDim.x > 192;
    const bool do_u = warp == 0, do_v = split_prod ? warp == 6 : warp == 0;
    WTRACE_DECL(do_u ? 1 : 2)
    uint32_t s = 0, ph = 1;
    int tile = t_begin;
    int tx = tile % p.tiles_x;
    int ty = (tile / p.tiles_x) % p.tiles_y;
    int img = tile / (p.tiles_x * p.tiles_y);
    int pass = 0;  // split-bf16: 0 = U_hi * V_hi, 1 = U_hi * V_lo, 2 = U_lo * V_hi
    for (int it = 0; it < iters; ++it) {
      WTR();
      mbar_wait(&empty[s], ph, p.err, DEV_ERR_EMPTY_TIMEOUT);
      WTR();
      const int x0 = tx * 8, y0 = ty * 8;
      const int mu = m0 + (pass == 2 ? p.u_lo : 0), nv = n0 + (pass == 1 ? p.v_lo : 0);
      uint8_t* u_dst = smem + s * C::kStageBytes;
      uint8_t* v_dst = u_dst + C::kUBytesC;
      if (elect_one_sync()) {
        if (do_u) {
          // the byte count may be armed after warp 6's loads have begun to land: the phase cannot complete before this
          // arrival, and the transaction count is signed
          mbar_arrive_expect_tx(&full[s], NKX != 2 ? C::kStageBytes : kUBytes + nkx * (BN / 64) * C::kVSlab);
          if (NKX == 3) {
#pragma unroll
            for (int a = 0; a < 3; ++a)  // U tile a pairs with the unshifted V box as filter column kx = a
              tma_load_5d(u_dst + a * 8192, &mapU, &full[s], mu, x0 - p.sign * (a - 1), y0, img, 0);
          } else {
#pragma unroll
            for (int slab = 0; slab < 2; ++slab)
              tma_load_5d(u_dst + slab * 8192, &mapU, &full[s], mu + slab * 64, x0, y0, img, 0);
          }
        }
        if (!do_v) {
        } else if (NKX == 3) {
          tma_load_5d(v_dst, &mapV, &full[s], nv, x0, y0 - 1, img, 0);
        } else if (MODE == 0) {
          if (HALO) {
#pragma unroll
            for (int b = 0; b < NKX; ++b) {
              if (b < nkx) {
                const int sx = p.sign * (kx + b - 1);
#pragma unroll
                for (int slab = 0; slab < BN / 64; ++slab)
                  tma_load_5d(v_dst + (b * (BN / 64) + slab) * C::kVSlab, &mapV, &full[s], nv + slab * 64, x0 + sx,
                              y0 - 1, img, 0);
              }
            }
          } else {
            const int sx = p.sign * (kx - 1);
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
              for (int slab = 0; slab < BN / 64; ++slab)
                tma_load_5d(v_dst + (j * (BN / 64) + slab) * C::kVSlab, &mapV, &full[s], nv + slab * 64, x0 + sx,
                            y0 + j - 1, img, 0);
          }
        } else if (MODE == 1) {
#pragma unroll
          for (int slab = 0; slab < BN / 64; ++slab)
            tma_load_5d(v_dst + slab * C::kVSlab, &mapV, &full[s], nv + slab * 64, x0, y0, img, 0);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int slab = 0; slab < BN / 64; ++slab)
              tma_load_5d(v_dst + (j * (BN / 64) + slab) * C::kVSlab, &mapV, &full[s], nv + slab * 64, j & 1, x0,
                          j >> 1, img * p.H + y0);
        }
      }
      __syncwarp();
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
      if (++pass == p.passes) {
        pass = 0;
        if (++tx == p.tiles_x) {
          tx = 0;
          if (++ty == p.tiles_y) {
            ty = 0;
            ++img;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (whole warp runs the loop, one elected lane issues) ----------------
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
    const uint32_t desc_hi = smem_desc_hi(1024);
    const uint32_t u_lo0 = smem_desc_lo(smem_u32(smem), 8192);
    const uint32_t v_lo0 = smem_desc_lo(smem_u32(smem) + C::kUBytesC, C::kVSlab);
    uint32_t s = 0, ph = 0;
    WTRACE_DECL(0)
    for (int it = 0; it < iters; ++it) {
      WTR();
      mbar_wait(&full[s], ph, p.err, DEV_ERR_FULL_TIMEOUT);
      WTR();
      tc_fence_after();
      const uint32_t u_lo = u_lo0 + s * (C::kStageBytes >> 4);
      const uint32_t v_lo = v_lo0 + s * (C::kStageBytes >> 4);
      if (NKX == 3) {
        // rows 0..63 / 64..127 of the first accumulator: kx 0 / kx 1; rows 0..63 of the second: kx 2 (its upper half
        // multiplies whatever follows the third U tile — the V box, finite — and is never read); columns [ky][64]
        constexpr uint32_t idesc3 = make_idesc_bf16(128, 192, 1, 1);
        const uint32_t vp = (v_lo & 0xFFFFu) + ((1024u >> 4) << 16);
        if (elect_one_sync()) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lo(tmem_base + h * 192, u_lo + h * (16384 >> 4) + k * (2048 >> 4), vp + k * (2048 >> 4), desc_hi,
                           idesc3, it > 0 || k > 0);
          umma_commit(&empty[s]);
        }
      } else if (BN == 64 && HALO && MODE == 0) {
        // 64-wide N tiles: an N = 64 MMA costs as much tensor-pipe time as an N = 128 one (measured: ~70 cycles either
        // way), so two taps are issued as ONE N = 128 MMA — the B descriptor's leading byte offset is the distance
        // between the two taps' 64-channel blocks (next ky: 1024 bytes; last ky of a box -> first ky of the next box:
        // box size - 2048). Accumulator columns are unchanged: tap j still owns columns [64 j, 64 j + 64).
        constexpr uint32_t idesc2 = make_idesc_bf16(128, 128, 1, 1);
        const uint32_t v_base = (v_lo & 0xFFFFu);  // start-address field only (the LBO field is rebuilt per pair)
        const int ntaps = 3 * nkx;
        if (elect_one_sync()) {
#pragma unroll
          for (int j = 0; j + 1 < C::kTaps; j += 2) {
            if (j + 1 >= ntaps) break;
            const uint32_t off0 = (j / 3) * C::kVSlab + (j % 3) * 1024;
            const uint32_t off1 = ((j + 1) / 3) * C::kVSlab + ((j + 1) % 3) * 1024;
            const uint32_t vp = v_base + (off0 >> 4) + (((off1 - off0) >> 4) << 16);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lo(tmem_base + j * BN, u_lo + k * (2048 >> 4), vp + k * (2048 >> 4), desc_hi, idesc2, it > 0 || k > 0);
          }
          if (ntaps & 1) {
            const int j = ntaps - 1;
            const uint32_t vj = v_lo + (j / 3) * (C::kVSlab >> 4) + (j % 3) * (1024 >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lo(tmem_base + j * BN, u_lo + k * (2048 >> 4), vj + k * (2048 >> 4), desc_hi, idesc, it > 0 || k > 0);
          }
          umma_commit(&empty[s]);
        }
      } else if (BN == 128 && HALO && MODE == 0 && NKX == 1) {
        // 128-wide N tiles: the three ky taps of one 64-channel slab of the V box are 1024 bytes apart, so they form
        // ONE N = 192 operand (leading byte offset 1024) and U is read from shared memory twice per K step instead of
        // three times — this kernel is bound by the SM's 128 B/clk of shared-memory bandwidth (operand reads + TMA
        // writes), not by the tensor pipe. Accumulator columns: [slab][tap][64 channels].
        constexpr uint32_t idesc3 = make_idesc_bf16(128, 192, 1, 1);
        const uint32_t v_base = (v_lo & 0xFFFFu);
        if (elect_one_sync()) {
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) {
            const uint32_t vp = v_base + sl * (C::kVSlab >> 4) + ((1024u >> 4) << 16);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lo(tmem_base + sl * 192, u_lo + k * (2048 >> 4), vp + k * (2048 >> 4), desc_hi, idesc3, it > 0 || k > 0);
          }
          umma_commit(&empty[s]);
        }
      } else if (elect_one_sync()) {
#pragma unroll
        for (int j = 0; j < C::kTaps; ++j) {
          if (NKX == 2 && j >= 3 * nkx) break;
          // HALO: tap j = box j / 3, rows [8 (j % 3), 8 (j % 3) + 64) of its 80-row box (whole 1024-byte swizzle atoms)
          const uint32_t vj = HALO ? v_lo + (j / 3) * (((BN / 64) * C::kVSlab) >> 4) + (j % 3) * (1024 >> 4)
                                   : v_lo + j * (((BN / 64) * C::kVSlab) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 pixels (K) per MMA = 16 rows of 128 bytes
            umma_bf16_lo(tmem_base + j * BN, u_lo + k * (2048 >> 4), vj + k * (2048 >> 4), desc_hi, idesc, it > 0 || k > 0);
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
    }
    WTR();
    if (elect_one_sync()) umma_commit(accbar);
    __syncwarp();
  } else {
    // ---------------- epilogue: TMEM -> fp32 workspace ----------------
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    WTRACE_DECL(3)
    WTR();
    mbar_wait(accbar, 0, p.err, DEV_ERR_ACC_TIMEOUT);
    WTR();
    tc_fence_after();
    float* base = p.ws + static_cast<long long>(split) * p.split_stride;
    if (NKX == 3) {
      const int ch = (q & 1) * 32 + lane;  // U channel of this accumulator row
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && q >= 2) break;       // upper half of the second accumulator is padding
        const int kxx = h == 0 ? (q >> 1) : 2;
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
          const int tap = (p.sign > 0 ? j : 2 - j) * 3 + kxx;
          float* tbase = base + tap * p.tap_stride + static_cast<long long>(ch) * p.m_stride;
#pragma unroll 1
          for (int c32 = 0; c32 < 2; ++c32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * 192 + j * 64 + c32 * 32, v);
            tmem_ld_wait();
            if (ch < p.cu) {
              if (p.n_stride == 1) {
                float4* dst = reinterpret_cast<float4*>(tbase + n0 + c32 * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                       __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  tbase[static_cast<long long>(n0 + c32 * 32 + i) * p.n_stride] = __uint_as_float(v[i]);
              }
            }
          }
        }
      }
    } else
#pragma unroll 1
    for (int j = 0; j < C::kTaps; ++j) {
      if (NKX == 2 && j >= 3 * nkx) break;
      int tap = j;
      if (MODE == 0) tap = (p.sign > 0 ? j % 3 : 2 - j % 3) * 3 + kx + j / 3;
      float* tbase = base + tap * p.tap_stride + static_cast<long long>(m) * p.m_stride;
#pragma unroll 1
      for (int c32 = 0; c32 < BN / 32; ++c32) {
        uint32_t v[32];
        const int col = (BN == 128 && HALO && MODE == 0 && NKX == 1) ? (c32 >> 1) * 192 + j * 64 + (c32 & 1) * 32
                                                                     : j * BN + c32 * 32;
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col, v);
        tmem_ld_wait();
        if (m < p.cu) {
          if (p.n_stride == 1) {
            float4* dst = reinterpret_cast<float4*>(tbase + n0 + c32 * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) tbase[static_cast<long long>(n0 + c32 * 32 + i) * p.n_stride] = __uint_as_float(v[i]);
          }
        }
      }
    }
    WTR();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int BN, int MODE, bool HALO, int NKX = 1>
cudaError_t launch_one(const CUtensorMap& mapU, const CUtensorMap& mapV, const WgradParams& p, cudaStream_t stream) {
  using C = WgradCfg<BN, MODE, HALO, NKX>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<BN, MODE, HALO, NKX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kDynamic);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int z = NKX == 3 ? p.splits : NKX == 2 ? p.splits + p.splits2 : (MODE == 0 ? 3 : 1) * p.splits;
  dim3 grid((p.cu + 127) / 128, p.cv / BN, z);
  launch_k(wgrad_kernel<BN, MODE, HALO, NKX>, dim3(grid), dim3(wgrad_threads()), C::kDynamic, stream, mapU, mapV, p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_wgrad(const CUtensorMap& mapU, const CUtensorMap& mapV, const WgradParams& p, int bn, int halo,
                         cudaStream_t stream) {
  if (p.mode == 0) {
    if (bn == 128) return halo ? launch_one<128, 0, true>(mapU, mapV, p, stream) : launch_one<128, 0, false>(mapU, mapV, p, stream);
    if (bn == 64 && halo && p.splits2 > 0) return launch_one<64, 0, true, 2>(mapU, mapV, p, stream);
    if (bn == 64 && halo && wgrad_mstack(p.cu)) return launch_one<64, 0, true, 3>(mapU, mapV, p, stream);
    if (bn == 64) return halo ? launch_one<64, 0, true>(mapU, mapV, p, stream) : launch_one<64, 0, false>(mapU, mapV, p, stream);
  } else if (p.mode == 1) {
    if (bn == 128) return launch_one<128, 1, false>(mapU, mapV, p, stream);
    if (bn == 64) return launch_one<64, 1, false>(mapU, mapV, p, stream);
  } else if (p.mode == 2) {
    if (bn == 128) return launch_one<128, 2, false>(mapU, mapV, p, stream);
    if (bn == 64) return launch_one<64, 2, false>(mapU, mapV, p, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace b200cd

#ifdef B200CD_TRACE
extern "C" int b200cd_debug_trace_wgrad(long long* host) {
  return static_cast<int>(cudaMemcpyFromSymbol(host, b200cd::g_trace_w, sizeof(long long) * 4 * 4096));
}
#endif
