// G1 — implicit-GEMM convolution on tcgen05 / TMEM, fed by TMA (sm_100a only).
//
// One CTA computes a 128-pixel x BN-channel output tile. The pixel tile is a tw x th rectangle of one
// image, so the A operand of every filter tap is ONE TMA box load from the NHWC activation tensor at
// a shifted coordinate; zero padding is the TMA out-of-bounds fill. The box lands in shared memory as
// 128 rows of 128 bytes with the 128-byte swizzle = the canonical K-major UMMA operand, no im2col.
// K loop = taps x (channels / 64); accumulators live in TMEM; the epilogue adds the bias, rounds to
// bf16, stages the tile in (swizzled) shared memory, reduces the BatchNorm partial statistics of the
// rounded values and stores the tile with TMA.
//
// Replaces (reference): nn.Conv2d(.,.,3,padding=1) utils/networks.py:392,395 (forward and, with
// flipped/transposed weights, its input gradient), nn.ConvTranspose2d(c,c,2,stride=2)
// utils/networks.py:433 (forward: out_mode 1; input gradient: mode 2).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
#include "kernels.h"
#include "ptx.cuh"

namespace b200cd {

namespace {

constexpr int kThreads = 192;
constexpr int kABytes = 128 * 128;  // 128 pixels x 64 bf16

// HALO (3x3 mode only): one stage = one kx and one 64-channel chunk; the A box carries a halo row above and below
// the tile (th+2 rows, <= 160 rows = 20 KB) and serves the three ky taps as address offsets of ky*tw rows (whole
// 1024-byte swizzle atoms, tw % 8 == 0), so A crosses L2 -> SM 3.6 times per output tile instead of 9; B = three
// weight tiles (one per ky).
template <int BN, int STAGES, bool HALO>
struct FpropSmem {
  static constexpr int kABuf = HALO ? 160 * 128 : kABytes;
  static constexpr int kBTile = BN * 128;
  static constexpr int kBBytes = (HALO ? 3 : 1) * kBTile;
  static constexpr int kStageBytes = kABuf + kBBytes;
  static constexpr int kBarOff = STAGES * kStageBytes;
  static constexpr int kTmemSlotOff = kBarOff + 8 * (2 * STAGES + 1);
  static constexpr int kBiasOff = kTmemSlotOff + 16;
  static constexpr int kRedOff = kBiasOff + BN * 4;
  static constexpr int kTotal = kRedOff + 4 * BN * 2 * 4;
  static constexpr int kDynamic = kTotal + 1024;  // slack for the manual 1024-byte alignment
};

template <int BN, int STAGES, bool HALO>
__global__ void __launch_bounds__(kThreads) fprop_kernel(const __grid_constant__ CUtensorMap mapA,
                                                         const __grid_constant__ CUtensorMap mapB,
                                                         const __grid_constant__ CUtensorMap mapO,
                                                         const FpropParams p) {
  using L = FpropSmem<BN, STAGES, HALO>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* accbar = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);
  float* bias_s = reinterpret_cast<float*>(smem + L::kBiasOff);
  float* red = reinterpret_cast<float*>(smem + L::kRedOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * BN;
  const int tx = tile % p.tiles_x;
  const int ty = (tile / p.tiles_x) % p.tiles_y;
  const int img = tile / (p.tiles_x * p.tiles_y);
  const int x0 = tx * p.tw, y0 = ty * p.th;
  const int iters = (HALO ? 3 : p.taps) * p.kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapO);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accbar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  pdl_wait();  // everything above is independent of the predecessor kernel's output
  if (threadIdx.x >= 64) {
    for (int t = threadIdx.x - 64; t < BN; t += 128) {
      float b = 0.f;
      if (p.bias) b = p.bias[p.out_mode == 1 ? (n0 + t) % p.cout : (n0 + t)];
      bias_s[t] = b;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer (whole warp runs the loop, one elected lane issues) ----------------
    uint32_t s = 0, ph = 1;
    int tap = 0, kc = 0;  // HALO: tap == kx
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&empty[s], ph, p.err, DEV_ERR_EMPTY_TIMEOUT);
      uint8_t* a_dst = smem + s * L::kStageBytes;
      uint8_t* b_dst = a_dst + L::kABuf;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full[s], p.rows * 128 + L::kBBytes);
        if (HALO) {
          tma_load_5d(a_dst, &mapA, &full[s], kc * 64, x0 + tap - 1, y0 - 1, img, 0);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
            tma_load_2d(b_dst + ky * L::kBTile, &mapB, &full[s], (ky * 3 + tap) * p.ka + kc * 64, n0);
        } else {
          if (p.mode == 0) {
            const int ky = tap / 3, kx = tap - ky * 3;
            tma_load_5d(a_dst, &mapA, &full[s], kc * 64, x0 + kx - 1, y0 + ky - 1, img, 0);
          } else if (p.mode == 1) {
            tma_load_5d(a_dst, &mapA, &full[s], kc * 64, x0, y0, img, 0);
          } else {
            const int dy = tap >> 1, dx = tap & 1;
            tma_load_5d(a_dst, &mapA, &full[s], kc * 64, dx, x0, dy, img * p.H + y0);
          }
          tma_load_2d(b_dst, &mapB, &full[s], tap * p.ka + kc * 64, n0);
        }
      }
      __syncwarp();
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
      if (++kc == p.kchunks) {
        kc = 0;
        ++tap;
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (whole warp runs the loop, one elected lane issues) ----------------
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    const uint32_t desc_hi = smem_desc_hi(1024);
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), 16);
    const uint32_t ky_step = static_cast<uint32_t>(p.tw * 128) >> 4;
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[s], ph, p.err, DEV_ERR_FULL_TIMEOUT);
      tc_fence_after();
      const uint32_t a_lo = a_lo0 + s * (L::kStageBytes >> 4);
      const uint32_t b_lo = a_lo + (L::kABuf >> 4);
      if (elect_one_sync()) {
        if (HALO) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int k = 0; k < 4; ++k)  // +32 bytes along K inside the swizzled 128-byte row: +2 in the >>4 field
              umma_bf16_lo(tmem_base, a_lo + ky * ky_step + 2 * k, b_lo + ky * (L::kBTile >> 4) + 2 * k, desc_hi, idesc,
                           it > 0 || ky > 0 || k > 0);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lo(tmem_base, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, it > 0 || k > 0);
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
    }
    if (elect_one_sync()) umma_commit(accbar);
    __syncwarp();
  } else {
    // ---------------- epilogue (128 threads) ----------------
    const int q = warp & 3;
    const int m = q * 32 + lane;  // row of the tile = pixel (m / tw, m % tw)
    uint8_t* stg = smem;          // staging aliases stage 0: slab 0 = its A region, slab 1 = its B region
    mbar_wait(accbar, 0, p.err, DEV_ERR_ACC_TIMEOUT);
    tc_fence_after();
#pragma unroll 1
    for (int c32 = 0; c32 < BN / 32; ++c32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c32 * 32, v);
      tmem_ld_wait();
      const int slab = (c32 * 32) >> 6;
      const int chunk0 = ((c32 * 32) & 63) >> 3;
      uint8_t* row = stg + slab * L::kABuf + m * 128;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[cc * 8 + j]) + bias_s[c32 * 32 + cc * 8 + j];
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1]);
        o.y = pack_bf16x2(f[2], f[3]);
        o.z = pack_bf16x2(f[4], f[5]);
        o.w = pack_bf16x2(f[6], f[7]);
        const int phys = (chunk0 + cc) ^ (m & 7);
        *reinterpret_cast<uint4*>(row + phys * 16) = o;
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    named_barrier_sync(1, 128);

    const int t = threadIdx.x - 64;
    if (t == 0) {
#pragma unroll
      for (int slab = 0; slab < BN / 64; ++slab) {
        const int n = n0 + slab * 64;
        if (p.out_mode == 0) {
          tma_store_5d(&mapO, stg + slab * L::kABuf, n, x0, y0, img, 0);
        } else {
          const int tap = n / p.cout, co = n - tap * p.cout;
          tma_store_5d(&mapO, stg + slab * L::kABuf, co, tap & 1, x0, tap >> 1, img * p.H + y0);
        }
      }
      tma_store_commit();
    }

    if (p.stats != nullptr) {
      // per-channel partial sums over the 128 staged (bf16-rounded) rows; conflict-free swizzled reads
      const int cp = t & 31;  // channel pair inside a 64-channel slab
      const int rq = t >> 5;  // row quarter
#pragma unroll
      for (int slab = 0; slab < BN / 64; ++slab) {
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        for (int r = rq * 32; r < rq * 32 + 32; ++r) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(stg + slab * L::kABuf + r * 128 +
                                                                (((cp >> 2) ^ (r & 7)) << 4) + ((cp & 3) << 2));
          float lo = bf16_lo(w), hi = bf16_hi(w);
          if (p.ragged) {
            const bool ok = (r < p.rows) && (x0 + r % p.tw < p.W) && (y0 + r / p.tw < p.H);
            lo = ok ? lo : 0.f;
            hi = ok ? hi : 0.f;
          }
          s0 += lo;
          q0 += lo * lo;
          s1 += hi;
          q1 += hi * hi;
        }
        float* dst = red + ((rq * BN) + slab * 64 + 2 * cp) * 2;
        dst[0] = s0;
        dst[1] = q0;
        dst[2] = s1;
        dst[3] = q1;
      }
      named_barrier_sync(2, 128);
      for (int ch = t; ch < BN; ch += 128) {
        float s = 0.f, qq = 0.f;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          s += red[((r4 * BN) + ch) * 2];
          qq += red[((r4 * BN) + ch) * 2 + 1];
        }
        p.stats[static_cast<size_t>(tile) * p.N + n0 + ch] = make_float2(s, qq);
      }
    }
    if (t == 0) tma_store_wait_read0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

template <int BN, int STAGES, bool HALO>
cudaError_t launch_one(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapO,
                       const FpropParams& p, int num_tiles, cudaStream_t stream) {
  using L = FpropSmem<BN, STAGES, HALO>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fprop_kernel<BN, STAGES, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::kDynamic);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(num_tiles, p.N / BN, 1);
  launch_k(fprop_kernel<BN, STAGES, HALO>, dim3(grid), dim3(kThreads), L::kDynamic, stream, mapA, mapB, mapO, p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_fprop(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapO,
                         const FpropParams& p, int bn, int halo, int num_tiles, cudaStream_t stream) {
  // Shared memory is sized so that two CTAs stay resident per SM where possible: one CTA's epilogue then overlaps
  // the other's main loop.
  if (halo) {
    if (bn == 128) return launch_one<128, 3, true>(mapA, mapB, mapO, p, num_tiles, stream);  // 204 KB: 1 CTA / SM
    if (bn == 64) return launch_one<64, 2, true>(mapA, mapB, mapO, p, num_tiles, stream);    // 88 KB: 2 CTAs / SM
    return cudaErrorInvalidValue;
  }
  if (bn == 256) return launch_one<256, 4, false>(mapA, mapB, mapO, p, num_tiles, stream);  // 192 KB: 1 CTA / SM
  if (bn == 128) return launch_one<128, 3, false>(mapA, mapB, mapO, p, num_tiles, stream);
  if (bn == 64) return launch_one<64, 4, false>(mapA, mapB, mapO, p, num_tiles, stream);
  return cudaErrorInvalidValue;
}

}  // namespace b200cd
