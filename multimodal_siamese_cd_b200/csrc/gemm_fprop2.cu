// G1p — 3x3 implicit-GEMM convolution on CTA PAIRS (tcgen05 cta_group::2), persistent, sm_100a only.
//
// Why a second kernel: ncu on gemm_fprop.cu shows the convolutions bound by the L2 -> SM operand stream
// (l1tex__m_xbar2l1tex_read_bytes ~ 39 B/clk/SM, the fabric's limit), not by the tensor pipe, and most of that stream
// is the WEIGHT tile, which every CTA re-reads for every 128-pixel tile. Here two CTAs on the two SMs of a TPC compute
// one 256-pixel x BN tile with a single M = 256 MMA: each CTA loads its own 128 pixels of A (with the one-row halo, so
// the three ky taps share one box) and only HALF of the weight tile, which the tensor cores of both SMs read. The
// kernel is persistent (one pair per TPC, static round-robin over work items) with
//   * separate shared-memory rings for A boxes (one per kx and 64-channel chunk) and B half-tiles (one per tap),
//   * weights held resident in the B ring when all taps of the layer fit (64 -> 64, 128 -> 64, 64 -> 128 channels:
//     the weight stream then disappears entirely),
//   * two accumulators in TMEM, so the epilogue of item i (TMEM -> +bias -> bf16 -> swizzled staging -> TMA store,
//     BatchNorm partial sums of the rounded values) runs under the MMAs of item i+1.
//
// Replaces (reference): nn.Conv2d(.,.,3,padding=1) utils/networks.py:392,395 — forward and, with flipped/transposed
// weights, its input gradient. Same math and rounding points as fprop_kernel<., ., HALO> (results are bit-identical:
// the K loop runs in the same order).
//
// Warp roles (192 threads per CTA): warp 0 = TMA producer (both CTAs), warp 1 = TMEM owner + MMA issuer (leader CTA
// issues for the pair), warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
#include "kernels.h"
#include "ptx.cuh"

namespace b200cd {

namespace {

constexpr int kThreads = 352;  // A producer warp, MMA warp, two epilogue warpgroups, B producer warp
constexpr int kBProducerWarp = 10;
constexpr int kABox = 160 * 128;  // tw x (th + 2) <= 160 pixels x 64 bf16 (16 x 10 or 8 x 18 boxes)
constexpr int kStageSlab = 128 * 128;  // one 64-channel slab of the output tile

// NKY = filter rows served by one A box: 3 for the 3x3 convolution (halo box), 1 for the single-tap GEMMs (first-layer
// conv on im2col rows, transposed-conv forward with the 2x2 scatter epilogue).
template <int BN, int NKY = 3>
struct PairCfg {
  static constexpr int kBTile = (BN / 2) * 128;  // this CTA's half of a weight tile: BN/2 rows x 64 bf16
  // weight tiles per B-ring slot (= per barrier and per TMA instruction): up to 128-wide tiles the three ky tiles of a
  // step travel together as ONE rank-4 box [ky][n][64 k] (a thread can start a bulk-tensor load only every ~230 cycles
  // — tools/ubench/tma_rate.cu — and three 8 KB loads per step left the weight ring starved), and the issuer handles
  // one barrier per 12 MMAs
  static constexpr int kG = (BN <= 128 && NKY == 3) ? 3 : 1;
  static constexpr int kBSlot = kG * kBTile;
  static constexpr int kStg = BN == 64 ? 1 : 2;                   // staging slabs per epilogue group
  static constexpr int kSA = BN == 64 ? 5 : 3;
  static constexpr int kSB = BN == 256 ? 6 : (BN == 128 ? (kG == 3 ? 4 : 12) : 6);  // slots of kG tiles
  static constexpr int kAOff = 0;
  static constexpr int kBOff = kSA * kABox;
  static constexpr int kStgOff = kBOff + kSB * kBSlot;           // two staging slabs (ping-pong) per epilogue group
  static constexpr int kBarOff = kStgOff + 2 * kStg * kStageSlab;       // a_full, a_empty, b_full, b_empty, acc_full[2], acc_empty[2]
  static constexpr int kNumBars = 2 * kSA + 2 * kSB + 4;
  static constexpr int kTmemSlotOff = kBarOff + 8 * kNumBars;
  static constexpr int kBiasOff = kTmemSlotOff + 16;
  static constexpr int kRedOff = kBiasOff + 2 * BN * 4;          // one bias copy per epilogue group
  static constexpr int kTotal = kRedOff + 2 * 4 * 64 * 2 * 4;    // per group: [row quarter][64 channels][sum, sumsq]
  // slack for the manual 1024-byte alignment, capped at the 227 KB limit (the kernel checks that the aligned layout
  // still fits and reports DEV_ERR_SMEM_LAYOUT otherwise; the dynamic window starts 1024-aligned in practice)
  static constexpr int kDynamic = kTotal + 1024 <= 232448 ? kTotal + 1024 : 232448;
  static constexpr int kTmemCols = 2 * BN;
  static_assert(kTotal + 16 <= 232448, "shared memory budget");
};

#ifdef B200CD_TRACE
// debug build only (B200CD_NVCC_EXTRA=-DB200CD_TRACE): clock64 stamps of CTA 0's warps, read by tools/trace_pair.py
__device__ long long g_trace[6][4096];
#define TRACE_DECL(role)                                            \
  int tr_i = 0;                                                     \
  const bool tr_on = blockIdx.x == 0 && (threadIdx.x & 31) == 0;    \
  const int tr_role = (role);
#define TR()                                                                    \
  do {                                                                          \
    if (tr_on && tr_i < 4096) g_trace[tr_role][tr_i++] = clock64();             \
  } while (0)
#else
#define TRACE_DECL(role)
#define TR()
#endif

static int pair_threads() {
  static const int t = [] {
    const char* e = getenv("B200CD_PAIR_SPLIT");
    return (e && e[0] == '0') ? 320 : kThreads;
  }();
  return t;
}

// 32 values per lane, 32 lanes -> lane l ends up with the sum over the warp of value l (transposing butterfly: 31
// shuffles instead of 32 x 5). Fixed order, so the result is run-to-run deterministic.
__device__ __forceinline__ float warp_transpose_sum32(float (&x)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < n / 2) {
        const float send = up ? x[j] : x[j + n / 2];
        const float keep = up ? x[j + n / 2] : x[j];
        x[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
  return x[0];
}

template <int BN, bool RESIDENT, int NKY, bool PREC>
__global__ void __launch_bounds__(kThreads, 1) fprop_pair_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                 const __grid_constant__ CUtensorMap mapB,
                                                                 const __grid_constant__ CUtensorMap mapO,
                                                                 const FpropParams p, const int num_tiles) {
  using L = PairCfg<BN, NKY>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t align_pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  if (align_pad + L::kTotal > static_cast<uint32_t>(L::kDynamic)) {  // uniform over the grid: nobody touches a barrier
    if (threadIdx.x == 0) atomicCAS(p.err, 0, DEV_ERR_SMEM_LAYOUT);
    return;
  }
  uint8_t* smem = smem_raw + align_pad;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* a_empty = a_full + L::kSA;
  uint64_t* b_full = a_empty + L::kSA;
  uint64_t* b_empty = b_full + L::kSB;
  uint64_t* acc_full = b_empty + L::kSB;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmemSlotOff);
  float* bias_s = reinterpret_cast<float*>(smem + L::kBiasOff);
  float* red = reinterpret_cast<float*>(smem + L::kRedOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int num_pairs = (num_tiles + 1) >> 1;
  const int n_blocks = p.N / BN;
  const int num_items = num_pairs * n_blocks;  // item = n_block * num_pairs + pair (pairs fastest: all pairs share B)
  // A boxes per item: 3x3: one per kx and 64-channel chunk; otherwise one per tap (1, or the 4 gathered taps of the
  // transposed-conv input gradient) and chunk
  const int steps = (NKY == 3 ? 3 : p.taps) * p.kchunks;
  constexpr bool resident = RESIDENT;  // host: n_blocks == 1 and all 3 * steps weight tiles fit in the B ring

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapO);
    for (int s = 0; s < L::kSA; ++s) {
      mbar_init(&a_full[s], 1);  // the leader's producer arms it with the bytes of BOTH CTAs' loads
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < L::kSB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);  // one arrival per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, L::kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above is independent of the predecessor kernel's output

  if (warp == 0 || warp == kBProducerWarp) {
    // ---------------- TMA producers (warp 0: pixel boxes, warp 10: weight tiles; the two rings are independent, and
    // two issuing threads keep more loads in flight than one; full barriers live in the leader) ------
    const bool split = blockDim.x > 320;  // B200CD_PAIR_SPLIT=0 launches without warp 10: warp 0 feeds both rings
    const bool do_a = warp == 0, do_b = split ? !do_a : do_a;
    uint32_t sa = 0, pa = 1, sb = 0, pb = 1;  // ring slot and the parity to wait for on its empty barrier
    bool first = true;
    TRACE_DECL(do_a ? 1 : 2)
    for (int item = cluster_id; item < num_items; item += num_clusters) {
      const int nb = item / num_pairs;
      const int tile = 2 * (item - nb * num_pairs) + static_cast<int>(rank);
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int img = tile / (p.tiles_x * p.tiles_y);  // >= n_img for the phantom tile of an odd count: zero fill
      const int x0 = tx * p.tw, y0 = ty * p.th;
      const int nrow = nb * BN + static_cast<int>(rank) * (BN / 2);
      int kx = 0, kc = 0;  // same K order as fprop_kernel<., ., HALO>: kx outer, 64-channel chunk inner
      for (int st = 0; st < steps; ++st) {
        if (do_a) {
          TR();
          mbar_wait(&a_empty[sa], pa, p.err, DEV_ERR_EMPTY_TIMEOUT);
          TR();
          // split-bf16: chunk kc of a tap = (pass, real chunk); passes 0 and 1 read the hi half of A (against W_hi and
          // W_lo), pass 2 the lo half (against W_hi again)
          const int ac = PREC ? (kc % p.kreal) * 64 + (kc >= 2 * p.kreal ? p.a_lo : 0) : kc * 64;
          if (elect_one_sync()) {
            if (leader) mbar_arrive_expect_tx(&a_full[sa], 2 * p.rows * 128);
            if (NKY == 3) tma_load_5d_2cta(smem + L::kAOff + sa * kABox, &mapA, &a_full[sa], ac, x0 + kx - 1, y0 - 1, img, 0);
            else if (p.mode == 1) tma_load_5d_2cta(smem + L::kAOff + sa * kABox, &mapA, &a_full[sa], ac, x0, y0, img, 0);
            else  // mode 2: tap kx = (dy, dx) of the 2x2 / stride-2 gather from the full-resolution tensor
              tma_load_5d_2cta(smem + L::kAOff + sa * kABox, &mapA, &a_full[sa], ac, kx & 1, x0, kx >> 1, img * p.H + y0);
          }
          __syncwarp();
          if (++sa == L::kSA) {
            sa = 0;
            pa ^= 1;
          }
        }
        if (do_b && (!resident || first)) {
          if (L::kG == 3) {
            TR();
            if (!resident) mbar_wait(&b_empty[sb], pb, p.err, DEV_ERR_EMPTY_TIMEOUT);
            TR();
            if (elect_one_sync()) {
              if (leader) mbar_arrive_expect_tx(&b_full[sb], 2 * L::kBSlot);
              // mapB is the rank-4 view [kx][ky][n][k within tap]: one box = the three ky tiles, stacked
              tma_load_4d_2cta(smem + L::kBOff + sb * L::kBSlot, &mapB, &b_full[sb], kc * 64, nrow, 0, kx);
            }
            __syncwarp();
            if (++sb == L::kSB) {
              sb = 0;
              pb ^= 1;
            }
          } else {
#pragma unroll 1
            for (int ky = 0; ky < NKY; ++ky) {
              TR();
              if (!resident) mbar_wait(&b_empty[sb], pb, p.err, DEV_ERR_EMPTY_TIMEOUT);
              TR();
              if (elect_one_sync()) {
                if (leader) mbar_arrive_expect_tx(&b_full[sb], 2 * L::kBSlot);
                tma_load_2d_2cta(smem + L::kBOff + sb * L::kBSlot, &mapB, &b_full[sb],
                                 (NKY == 3 ? (ky * 3 + kx) : kx) * p.ka + kc * 64, nrow);
              }
              __syncwarp();
              if (++sb == L::kSB) {  // resident: the layer's tiles fill at most kSB slots, once, in order
                sb = 0;
                pb ^= 1;
              }
            }
          }
        }
        if (++kc == p.kchunks) {
          kc = 0;
          ++kx;
        }
      }
      first = false;
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (leader CTA; M = 256 across the pair) ----------------
      // The whole warp runs the loop (uniform control flow, operands in uniform registers); one elected lane issues.
      // Descriptors are (low word, constant high word) pairs so that stepping an operand is one 32-bit add.
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, 0, 0);
      const uint32_t desc_hi = smem_desc_hi(1024);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem + L::kAOff), 16);
      const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem + L::kBOff), 16);
      const uint32_t ky_step = static_cast<uint32_t>(p.tw * 128) >> 4;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;  // ring slot and the parity to wait for on its full barrier
      uint32_t n_item = 0;
      TRACE_DECL(0)
      for (int item = cluster_id; item < num_items; item += num_clusters, ++n_item) {
        const uint32_t buf = n_item & 1;
        TR();
        mbar_wait(&acc_empty[buf], ((n_item >> 1) & 1) ^ 1, p.err, DEV_ERR_ACC_TIMEOUT);
        TR();
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        if (resident) sb = 0;
#pragma unroll 1
        for (int st = 0; st < steps; ++st) {
          TR();
          mbar_wait(&a_full[sa], pa, p.err, DEV_ERR_FULL_TIMEOUT);
          TR();
          const uint32_t a_lo = a_lo0 + sa * (kABox >> 4);
          if (L::kG == 3 || resident) {
            // the three weight tiles of this step sit behind one barrier (kG == 3) or are resident: 12 MMAs and the
            // commits under one election
            if (L::kG == 3) {
              if (!resident || n_item == 0) mbar_wait(&b_full[sb], resident ? 0u : pb, p.err, DEV_ERR_FULL_TIMEOUT);
            } else if (n_item == 0) {
#pragma unroll
              for (int ky = 0; ky < NKY; ++ky) mbar_wait(&b_full[sb + ky], 0, p.err, DEV_ERR_FULL_TIMEOUT);
            }
            tc_fence_after();
            const uint32_t b_lo = b_lo0 + sb * (L::kBSlot >> 4);
            if (elect_one_sync()) {
#pragma unroll
              for (int ky = 0; ky < NKY; ++ky)
#pragma unroll
                for (int k = 0; k < 4; ++k)  // +32 bytes along K inside the swizzled 128-byte row: +2 in the >>4 field
                  umma_bf16_2cta_lo(tmem_d, a_lo + ky * ky_step + 2 * k, b_lo + ky * (L::kBTile >> 4) + 2 * k, desc_hi,
                                    idesc, st > 0 || ky > 0 || k > 0);
              if (!resident) umma_commit_2cta(&b_empty[sb], 3);
              umma_commit_2cta(&a_empty[sa], 3);
            }
            __syncwarp();
            sb += L::kG == 3 ? 1 : NKY;
            if (!resident && sb == L::kSB) {
              sb = 0;
              pb ^= 1;
            }
          } else {
            tc_fence_after();
#pragma unroll
            for (int ky = 0; ky < NKY; ++ky) {
              mbar_wait(&b_full[sb], pb, p.err, DEV_ERR_FULL_TIMEOUT);
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + sb * (L::kBSlot >> 4);
              if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_2cta_lo(tmem_d, a_lo + ky * ky_step + 2 * k, b_lo + 2 * k, desc_hi, idesc,
                                    st > 0 || ky > 0 || k > 0);
                umma_commit_2cta(&b_empty[sb], 3);
                if (ky == NKY - 1) umma_commit_2cta(&a_empty[sa], 3);
              }
              __syncwarp();
              if (++sb == L::kSB) {
                sb = 0;
                pb ^= 1;
              }
            }
          }
          if (++sa == L::kSA) {
            sa = 0;
            pa ^= 1;
          }
          TR();
        }
        if (elect_one_sync()) umma_commit_2cta(&acc_full[buf], 3);
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue: two groups of 128 threads, group g drains accumulator g (items g, g+2, ...) --------
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;  // row of this CTA's tile = pixel (m / tw, m % tw)
    const int t = threadIdx.x - 64 - grp * 128;
    const uint32_t bar1 = 1 + 2 * grp, bar2 = 2 + 2 * grp;
    uint8_t* stg = smem + L::kStgOff + grp * L::kStg * kStageSlab;
    bias_s += grp * BN;
    red += grp * 4 * 64 * 2;
    uint32_t n_item = grp, n_slab = 0;
    int bias_nb = -1;
    TRACE_DECL(q == 2 ? 3 + grp : 5)  // warps 2 and 6 (q == 2) stamp for their group
    // per-CTA BatchNorm statistics (p.stat_groups > 0): thread t < 64 keeps the running (sum, sum of squares) of
    // channel t of every slab of the current N block, per stat-group, and writes them once per N block
    const bool cta_stats = p.stats != nullptr && p.stat_groups > 0;
    // inference epilogue (host guarantees stats == nullptr then, so the statistics scratch `red` is free for the scales)
    const bool affine = p.ep_scale != nullptr;
    const float act_lo = p.ep_relu ? 0.f : -3.0e38f;
    const int per_group = cta_stats ? p.n_img / p.stat_groups : 1;
    const size_t my_row = 2 * blockIdx.x + grp;
    float acc_s[2][BN / 64], acc_q[2][BN / 64];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int sl = 0; sl < BN / 64; ++sl) acc_s[g][sl] = acc_q[g][sl] = 0.f;
    int acc_nb = -1;
    auto flush_stats = [&](int nb_done) {
      if (t < 64) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g >= p.stat_groups) break;
#pragma unroll
          for (int sl = 0; sl < BN / 64; ++sl) {
            p.stats[(static_cast<size_t>(g) * p.stat_rows + my_row) * p.N + nb_done * BN + sl * 64 + t] =
                make_float2(acc_s[g][sl], acc_q[g][sl]);
            acc_s[g][sl] = acc_q[g][sl] = 0.f;
          }
        }
      }
    };
    if (cta_stats) {  // rows of N blocks / groups this producer never meets stay zero
      for (int g = 0; g < p.stat_groups; ++g)
        for (int n = t; n < p.N; n += 128)
          p.stats[(static_cast<size_t>(g) * p.stat_rows + my_row) * p.N + n] = make_float2(0.f, 0.f);
    }
    for (int item = cluster_id + grp * num_clusters; item < num_items; item += 2 * num_clusters, n_item += 2) {
      const int nb = item / num_pairs;
      const int tile = 2 * (item - nb * num_pairs) + static_cast<int>(rank);
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int img = tile / (p.tiles_x * p.tiles_y);
      const int x0 = tx * p.tw, y0 = ty * p.th;
      const int n0 = nb * BN;
      const bool real = tile < num_tiles;
      if (cta_stats && nb != acc_nb) {
        if (acc_nb >= 0) flush_stats(acc_nb);
        acc_nb = nb;
      }
      const int sgrp = cta_stats ? img / per_group : 0;
      if (nb != bias_nb) {  // uniform across the 128 threads
        named_barrier_sync(bar1, 128);
        for (int i = t; i < BN; i += 128) {
          const float b = p.bias ? p.bias[p.out_mode == 1 ? (n0 + i) % p.cout : n0 + i] : 0.f;
          if (affine) {  // folded inference BatchNorm: (acc + b) * sc + sh = acc * sc + (b * sc + sh); `red` holds sc
            const float sc = __ldg(p.ep_scale + n0 + i);
            red[i] = sc;
            bias_s[i] = fmaf(b, sc, __ldg(p.ep_shift + n0 + i));
          } else {
            bias_s[i] = b;
          }
        }
        named_barrier_sync(bar1, 128);
        bias_nb = nb;
      }
      const uint32_t buf = grp;
      TR();
      mbar_wait(&acc_full[buf], (n_item >> 1) & 1, p.err, DEV_ERR_ACC_TIMEOUT);
      TR();
      tc_fence_after();
      if constexpr (PREC) {
        // ---- split-bf16 epilogue: every 64-channel slab is stored twice, hi = bf16(x) at channel n and
        // lo = bf16(x - hi) at channel n + o_lo; BatchNorm statistics are taken from the stored value hi + lo in
        // registers (transposing warp butterfly: lane l gets the column sum of channel l), not from the staged tile ----
        const int tw_shift = p.tw == 16 ? 4 : 3;
        const bool vrow = real && m < p.tw * p.th && (x0 + (m & (p.tw - 1)) < p.W) && (y0 + (m >> tw_shift) < p.H);
#pragma unroll 1
        for (int slab = 0; slab < BN / 64; ++slab) {
          float cs[2] = {0.f, 0.f}, cq[2] = {0.f, 0.f};
#pragma unroll 1
          for (int part = 0; part < 2; ++part, ++n_slab) {
            uint8_t* sbuf = stg + (n_slab % L::kStg) * kStageSlab;
            if (t == 0 && n_slab >= L::kStg) {
              if (L::kStg == 2) tma_store_wait_read1();
              else tma_store_wait_read0();
            }
            named_barrier_sync(bar1, 128);
            uint8_t* row = sbuf + m * 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t v[32];
              tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + slab * 64 + h * 32, v);
              tmem_ld_wait();
              float xs[32];
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                uint32_t w4[4];
#pragma unroll
                for (int j2 = 0; j2 < 4; ++j2) {
                  const int j = cc * 8 + 2 * j2;
                  const int cj = slab * 64 + h * 32 + j;
                  float f0, f1;
                  if (affine) {
                    f0 = fmaxf(fmaf(__uint_as_float(v[j]), red[cj], bias_s[cj]), act_lo);
                    f1 = fmaxf(fmaf(__uint_as_float(v[j + 1]), red[cj + 1], bias_s[cj + 1]), act_lo);
                  } else {
                    f0 = __uint_as_float(v[j]) + bias_s[cj];
                    f1 = __uint_as_float(v[j + 1]) + bias_s[cj + 1];
                  }
                  const uint32_t hi2 = pack_bf16x2(f0, f1);
                  const float l0 = f0 - bf16_lo(hi2), l1 = f1 - bf16_hi(hi2);
                  const uint32_t lo2 = pack_bf16x2(l0, l1);
                  w4[j2] = part == 0 ? hi2 : lo2;
                  xs[j] = vrow ? bf16_lo(hi2) + bf16_lo(lo2) : 0.f;
                  xs[j + 1] = vrow ? bf16_hi(hi2) + bf16_hi(lo2) : 0.f;
                }
                const int phys = (h * 4 + cc) ^ (m & 7);
                *reinterpret_cast<uint4*>(row + phys * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
              }
              if (part == 0 && p.stats != nullptr) {
                float xq[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) xq[j] = xs[j] * xs[j];
                cs[h] = warp_transpose_sum32(xs, lane);
                cq[h] = warp_transpose_sum32(xq, lane);
              }
            }
            if (slab == BN / 64 - 1 && part == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(&acc_empty[buf], 0);
            }
            fence_proxy_async_smem();
            named_barrier_sync(bar1, 128);
            if (t == 0) {
              const int lo_off = part ? p.o_lo : 0;
              if (p.out_mode == 0) {
                tma_store_5d(&mapO, sbuf, n0 + slab * 64 + lo_off, x0, y0, img, 0);
              } else {
                const int n = n0 + slab * 64;
                const int tap = n / p.cout, co = n - tap * p.cout;
                tma_store_5d(&mapO, sbuf, co + lo_off, tap & 1, x0, tap >> 1, img * p.H + y0);
              }
              tma_store_commit();
            }
            if (part == 0 && p.stats != nullptr) {
              float* dst = red + ((t >> 5) * 64 + lane) * 2;   // [warp][channel][sum, sum of squares]
              dst[0] = cs[0];
              dst[1] = cq[0];
              dst[64] = cs[1];
              dst[65] = cq[1];
              named_barrier_sync(bar2, 128);
              if (t < 64 && real) {
                float s = 0.f, qq = 0.f;
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4) {
                  s += red[((r4 * 64) + t) * 2];
                  qq += red[((r4 * 64) + t) * 2 + 1];
                }
                if (cta_stats) {
                  acc_s[0][slab] += sgrp == 0 ? s : 0.f;
                  acc_q[0][slab] += sgrp == 0 ? qq : 0.f;
                  acc_s[1][slab] += sgrp == 1 ? s : 0.f;
                  acc_q[1][slab] += sgrp == 1 ? qq : 0.f;
                } else {
                  p.stats[static_cast<size_t>(tile) * p.N + n0 + slab * 64 + t] = make_float2(s, qq);
                }
              }
            }
          }
        }
      } else
#pragma unroll
      for (int slab = 0; slab < BN / 64; ++slab, ++n_slab) {
        uint8_t* sbuf = stg + (n_slab % L::kStg) * kStageSlab;
        // fused BatchNorm-backward sums: this thread's 8 rows x 8 channels of the pre-BN tensor, requested before the
        // accumulator is read so that the global-load latency hides under the TMEM load / conversion / store issue
        uint4 rv[8];
        bool okr[8];
        if (p.bwd_r != nullptr) {
          const int oct = t & 7, rg = t >> 3;
          const int tw_shift = p.tw == 16 ? 4 : 3;
          const __nv_bfloat16* rbase = reinterpret_cast<const __nv_bfloat16*>(p.bwd_r) + n0 + slab * 64 + oct * 8;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = rg * 8 + u;
            const int yy = y0 + (r >> tw_shift), xx = x0 + (r & (p.tw - 1));
            okr[u] = real && r < p.tw * p.th && yy < p.H && xx < p.W;
            rv[u] = okr[u] ? __ldg(reinterpret_cast<const uint4*>(
                                 rbase + ((static_cast<long long>(img) * p.H + yy) * p.W + xx) * p.bwd_ld))
                           : make_uint4(0u, 0u, 0u, 0u);
          }
        }
        // the TMA store that last used this staging buffer has finished reading it
        if (t == 0 && n_slab >= L::kStg) {
          if (L::kStg == 2) tma_store_wait_read1();
          else tma_store_wait_read0();
        }
        named_barrier_sync(bar1, 128);
        uint8_t* row = sbuf + m * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + slab * 64 + h * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            float f[8];
            if (affine) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int cj = slab * 64 + h * 32 + cc * 8 + j;
                f[j] = fmaxf(fmaf(__uint_as_float(v[cc * 8 + j]), red[cj], bias_s[cj]), act_lo);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[cc * 8 + j]) + bias_s[slab * 64 + h * 32 + cc * 8 + j];
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            const int phys = (h * 4 + cc) ^ (m & 7);
            *reinterpret_cast<uint4*>(row + phys * 16) = o;
          }
        }
        if (slab == BN / 64 - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA issuer (leader's barrier)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&acc_empty[buf], 0);
          TR();
        }
        fence_proxy_async_smem();
        named_barrier_sync(bar1, 128);
        TR();
        if (t == 0) {
          if (p.out_mode == 0) {
            tma_store_5d(&mapO, sbuf, n0 + slab * 64, x0, y0, img, 0);  // clipped at the tensor bounds
          } else {  // 2x2 / stride-2 scatter of the transposed convolution: slab = 64 channels of one (dy, dx) tap
            const int n = n0 + slab * 64;
            const int tap = n / p.cout, co = n - tap * p.cout;
            tma_store_5d(&mapO, sbuf, co, tap & 1, x0, tap >> 1, img * p.H + y0);
          }
          tma_store_commit();
        }
        if (p.stats != nullptr) {
          // per-channel partial sums over the 128 staged (bf16-rounded) rows; conflict-free swizzled reads
          if (p.bwd_r != nullptr) {
            // BatchNorm-backward sums of the gradient tile being stored: per channel (sum dy*m, sum dy*m*r) with the
            // ReLU mask m recomputed from the pre-BN values exactly as the dx pass does. Thread = 8 rows x 8 channels;
            // the four row groups of a warp are combined with shuffles, the four warps through `red`.
            const int oct = t & 7, rg = t >> 3;
            const int c0 = n0 + slab * 64 + oct * 8;
            float sc[8], sh[8], ss[8], qq2[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              sc[jj] = __ldg(p.bwd_scale + sgrp * p.N + c0 + jj);
              sh[jj] = __ldg(p.bwd_shift + sgrp * p.N + c0 + jj);
              ss[jj] = qq2[jj] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int r = rg * 8 + u;
              const uint4 dv = *reinterpret_cast<const uint4*>(sbuf + r * 128 + ((oct ^ (r & 7)) << 4));
              const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
              const uint32_t rw[4] = {rv[u].x, rv[u].y, rv[u].z, rv[u].w};
#pragma unroll
              for (int h2 = 0; h2 < 4; ++h2) {
                const float rlo = bf16_lo(rw[h2]), rhi = bf16_hi(rw[h2]);
                const float dlo = (okr[u] && fmaf(rlo, sc[2 * h2], sh[2 * h2]) > 0.f) ? bf16_lo(dw[h2]) : 0.f;
                const float dhi = (okr[u] && fmaf(rhi, sc[2 * h2 + 1], sh[2 * h2 + 1]) > 0.f) ? bf16_hi(dw[h2]) : 0.f;
                ss[2 * h2] += dlo;
                qq2[2 * h2] = fmaf(dlo, rlo, qq2[2 * h2]);
                ss[2 * h2 + 1] += dhi;
                qq2[2 * h2 + 1] = fmaf(dhi, rhi, qq2[2 * h2 + 1]);
              }
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              ss[jj] += __shfl_xor_sync(0xffffffffu, ss[jj], 8);
              ss[jj] += __shfl_xor_sync(0xffffffffu, ss[jj], 16);
              qq2[jj] += __shfl_xor_sync(0xffffffffu, qq2[jj], 8);
              qq2[jj] += __shfl_xor_sync(0xffffffffu, qq2[jj], 16);
            }
            if ((t & 31) < 8) {
              float* dst = red + (((t >> 5) * 64) + oct * 8) * 2;   // [warp][channel][sum, sum*r]
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                dst[2 * jj] = ss[jj];
                dst[2 * jj + 1] = qq2[jj];
              }
            }
          } else {
            // thread = 8 rows x 8 channels (one 16-byte read per row); the four row groups of a warp are combined with
            // shuffles, the four warps through `red` — 8 shared-memory reads per thread instead of 32
            const int oct = t & 7, rg = t >> 3;
            const int tw_shift = p.tw == 16 ? 4 : 3;
            float ss[8], qq2[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) ss[jj] = qq2[jj] = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int r = rg * 8 + u;
              const uint4 dv = *reinterpret_cast<const uint4*>(sbuf + r * 128 + ((oct ^ (r & 7)) << 4));
              const bool ok = !p.ragged || ((r < p.tw * p.th) && (x0 + (r & (p.tw - 1)) < p.W) && (y0 + (r >> tw_shift) < p.H));
              const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
              for (int h2 = 0; h2 < 4; ++h2) {
                const float lo = ok ? bf16_lo(dw[h2]) : 0.f, hi = ok ? bf16_hi(dw[h2]) : 0.f;
                ss[2 * h2] += lo;
                qq2[2 * h2] = fmaf(lo, lo, qq2[2 * h2]);
                ss[2 * h2 + 1] += hi;
                qq2[2 * h2 + 1] = fmaf(hi, hi, qq2[2 * h2 + 1]);
              }
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              ss[jj] += __shfl_xor_sync(0xffffffffu, ss[jj], 8);
              ss[jj] += __shfl_xor_sync(0xffffffffu, ss[jj], 16);
              qq2[jj] += __shfl_xor_sync(0xffffffffu, qq2[jj], 8);
              qq2[jj] += __shfl_xor_sync(0xffffffffu, qq2[jj], 16);
            }
            if ((t & 31) < 8) {
              float* dst = red + (((t >> 5) * 64) + oct * 8) * 2;   // [warp][channel][sum, sum of squares]
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                dst[2 * jj] = ss[jj];
                dst[2 * jj + 1] = qq2[jj];
              }
            }
          }
          TR();
          named_barrier_sync(bar2, 128);
          TR();
          if (t < 64 && real) {
            float s = 0.f, qq = 0.f;
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) {
              s += red[((r4 * 64) + t) * 2];
              qq += red[((r4 * 64) + t) * 2 + 1];
            }
            if (cta_stats) {
              acc_s[0][slab] += sgrp == 0 ? s : 0.f;
              acc_q[0][slab] += sgrp == 0 ? qq : 0.f;
              acc_s[1][slab] += sgrp == 1 ? s : 0.f;
              acc_q[1][slab] += sgrp == 1 ? qq : 0.f;
            } else {
              p.stats[static_cast<size_t>(tile) * p.N + n0 + slab * 64 + t] = make_float2(s, qq);
            }
          }
          // `red` is rewritten only after the next slab's first named barrier, which every reader passes first
        }
      }
      TR();
    }
    if (cta_stats && acc_nb >= 0) flush_stats(acc_nb);
    if (t == 0) tma_store_wait_read0();
  }

  // teardown: both CTAs stay alive until every MMA / multicast arrive / store of the pair has retired
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, L::kTmemCols);
  }
}

template <int BN, bool RESIDENT, int NKY, bool PREC>
cudaError_t pair_max_clusters(int* out) {
  using L = PairCfg<BN, NKY>;
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cudaError_t e = cudaFuncSetAttribute(fprop_pair_kernel<BN, RESIDENT, NKY, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    cudaLaunchConfig_t qc{};
    qc.gridDim = dim3(sms & ~1);
    qc.blockDim = dim3(pair_threads());
    qc.dynamicSmemBytes = L::kDynamic;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    qc.attrs = qa;
    qc.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, fprop_pair_kernel<BN, RESIDENT, NKY, PREC>, &qc);
    if (e != cudaSuccess) return e;
    if (n < 1) return cudaErrorLaunchOutOfResources;
    max_clusters = n < sms / 2 ? n : sms / 2;
  }
  *out = max_clusters;
  return cudaSuccess;
}

template <int BN, bool RESIDENT, int NKY, bool PREC>
cudaError_t pair_clusters(const FpropParams& p, int num_tiles, int* clusters) {
  int max_clusters = 0;
  cudaError_t e = pair_max_clusters<BN, RESIDENT, NKY, PREC>(&max_clusters);
  if (e != cudaSuccess) return e;
  const int num_items = ((num_tiles + 1) / 2) * (p.N / BN);
  *clusters = num_items < max_clusters ? num_items : max_clusters;
  return cudaSuccess;
}

template <int BN, bool RESIDENT, int NKY, bool PREC>
cudaError_t launch_pair(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapO, const FpropParams& p,
                        int num_tiles, cudaStream_t stream) {
  using L = PairCfg<BN, NKY>;
  int clusters = 0;
  cudaError_t e = pair_clusters<BN, RESIDENT, NKY, PREC>(p, num_tiles, &clusters);
  if (e != cudaSuccess) return e;
  if (p.stats != nullptr && p.stat_groups > 0 && p.stat_rows != 4 * clusters) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(pair_threads());
  cfg.dynamicSmemBytes = L::kDynamic;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 2;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, fprop_pair_kernel<BN, RESIDENT, NKY, PREC>, mapA, mapB, mapO, p, num_tiles);
}

static bool pair_resident(const FpropParams& p, int bn) {
  // weights resident in shared memory when one N block covers the layer and all its half-tiles fit the B ring
  if (p.prec) return false;  // three times the weight tiles: streamed
  if (p.mode == 0) {
    const int cap = bn == 256 ? PairCfg<256>::kSB * PairCfg<256>::kG
                              : (bn == 128 ? PairCfg<128>::kSB * PairCfg<128>::kG : PairCfg<64>::kSB * PairCfg<64>::kG);
    return p.N == bn && 9 * p.kchunks <= cap;
  }
  const int cap = bn == 256 ? PairCfg<256, 1>::kSB : (bn == 128 ? PairCfg<128, 1>::kSB : PairCfg<64, 1>::kSB);
  return p.N == bn && p.taps * p.kchunks <= cap;
}

// (bn, resident, taps) -> instantiation
#define B200CD_PAIR_DISPATCH(FN, ...)                                                                          \
  do {                                                                                                          \
    const bool res = pair_resident(p, bn);                                                                      \
    if (p.prec) {                                                                                               \
      if (p.mode == 0) {                                                                                        \
        if (bn == 256) return FN<256, false, 3, true>(__VA_ARGS__);                                             \
        if (bn == 128) return FN<128, false, 3, true>(__VA_ARGS__);                                             \
        if (bn == 64) return FN<64, false, 3, true>(__VA_ARGS__);                                               \
      } else {                                                                                                  \
        if (bn == 256) return FN<256, false, 1, true>(__VA_ARGS__);                                             \
        if (bn == 128) return FN<128, false, 1, true>(__VA_ARGS__);                                             \
        if (bn == 64) return FN<64, false, 1, true>(__VA_ARGS__);                                               \
      }                                                                                                         \
    } else if (p.mode == 0) {                                                                                   \
      if (bn == 256) return res ? FN<256, true, 3, false>(__VA_ARGS__) : FN<256, false, 3, false>(__VA_ARGS__); \
      if (bn == 128) return res ? FN<128, true, 3, false>(__VA_ARGS__) : FN<128, false, 3, false>(__VA_ARGS__); \
      if (bn == 64) return res ? FN<64, true, 3, false>(__VA_ARGS__) : FN<64, false, 3, false>(__VA_ARGS__);    \
    } else {                                                                                                    \
      if (bn == 256) return res ? FN<256, true, 1, false>(__VA_ARGS__) : FN<256, false, 1, false>(__VA_ARGS__); \
      if (bn == 128) return res ? FN<128, true, 1, false>(__VA_ARGS__) : FN<128, false, 1, false>(__VA_ARGS__); \
      if (bn == 64) return res ? FN<64, true, 1, false>(__VA_ARGS__) : FN<64, false, 1, false>(__VA_ARGS__);    \
    }                                                                                                           \
  } while (0)

}  // namespace

bool fprop_pair_stacked_weights(int mode, int bn) { return mode == 0 && bn <= 128; }

cudaError_t launch_fprop_pair(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapO,
                              const FpropParams& p, int bn, int num_tiles, cudaStream_t stream) {
  if (!((p.mode == 0 && p.out_mode == 0) || p.mode == 1 || (p.mode == 2 && p.out_mode == 0))) return cudaErrorInvalidValue;
  if (p.stat_groups < 0 || p.stat_groups > 2) return cudaErrorInvalidValue;
  if (p.prec && (p.kreal < 1 || p.kchunks != 3 * p.kreal || p.bwd_r != nullptr)) return cudaErrorInvalidValue;
  if (p.ep_scale != nullptr && (p.stats != nullptr || p.ep_shift == nullptr || p.out_mode != 0)) return cudaErrorInvalidValue;
  B200CD_PAIR_DISPATCH(launch_pair, mapA, mapB, mapO, p, num_tiles, stream);
  return cudaErrorInvalidValue;
}

static cudaError_t pair_clusters_any(const FpropParams& p, int bn, int num_tiles, int* c) {
  B200CD_PAIR_DISPATCH(pair_clusters, p, num_tiles, c);
  return cudaErrorInvalidValue;
}

int fprop_pair_ctas(const FpropParams& p, int bn, int num_tiles) {
  int c = 0;
  return pair_clusters_any(p, bn, num_tiles, &c) == cudaSuccess ? 2 * c : -1;
}

}  // namespace b200cd

#ifdef B200CD_TRACE
extern "C" int b200cd_debug_trace(long long* host) {
  return static_cast<int>(cudaMemcpyFromSymbol(host, b200cd::g_trace, sizeof(long long) * 6 * 4096));
}
#endif
