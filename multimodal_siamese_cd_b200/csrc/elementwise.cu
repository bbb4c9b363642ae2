// Memory-bound kernels of the training step (sm_100a): everything that is not a tensor-core contraction.
// All activations are NHWC bf16 with an explicit element stride per pixel ("ld"), so producers can write
// into channel slices of a concatenation buffer and consumers can read them without copy kernels.
// Vector width is 8 channels = 16 bytes per thread access; reductions are two-stage (per-block partials,
// then a fixed-order fp64 finalize) so results are run-to-run deterministic.
//
// Reference ops replaced (file:line in /root/reference):
//   BatchNorm2d + ReLU (+ MaxPool2d, + torch.sub(t2, t1))   utils/networks.py:393-397, 420, 147-150
//   torch.cat / slicing of the inputs                         utils/networks.py:74, 105-113, 233-246
//   OutConv 1x1 heads                                         utils/networks.py:454-461
//   power_jaccard_loss                                        utils/loss_functions.py:141-150
#include "kernels.h"
#include "ptx.cuh"

namespace b200cd {

namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// ------------------------------------------------------------------------------------------------
// Input packing: fp32 NCHW inputs -> bf16 im2col rows [pixel][kpad], k = tap*Cin + ci, zero padded.
// The first-layer 3x3 conv (Cin in {2..8}) then runs as a plain GEMM on the tensor cores.
//   cat_mode 0: images = [src0 batch ; src1 batch] (shared-weight t1/t2 call), Cin = nc
//   cat_mode 1: channels = [src0 channels ; src1 channels] (early fusion), Cin = 2*nc
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pack_input_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                                         int csrc, int c_lo, int nc, int cat_mode, int B, int H, int W,
                                                         int kpad, __nv_bfloat16* __restrict__ out, long long npix) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint32_t sm32[];
  const int rowwords = kpad / 2 + 1;  // odd word stride -> conflict-free
  __nv_bfloat16* sm = reinterpret_cast<__nv_bfloat16*>(sm32);
  const int t = threadIdx.x;
  const long long pix0 = static_cast<long long>(blockIdx.x) * 128;
  const long long pix = pix0 + t;
  const int cin = cat_mode ? 2 * nc : nc;
  __nv_bfloat16* myrow = sm + static_cast<size_t>(t) * rowwords * 2;
  for (int k = 0; k < kpad; ++k) myrow[k] = __float2bfloat16_rn(0.f);
  if (pix < npix) {
    const int x = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    for (int ci = 0; ci < cin; ++ci) {
      const float* src;
      int b, c;
      if (cat_mode == 0) {
        src = n < B ? src0 : src1;
        b = n < B ? n : n - B;
        c = c_lo + ci;
      } else {
        src = ci < nc ? src0 : src1;
        b = n;
        c = c_lo + (ci < nc ? ci : ci - nc);
      }
      const float* plane = src + (static_cast<long long>(b) * csrc + c) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + kx - 1;
          float v = 0.f;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(plane + static_cast<long long>(yy) * W + xx);
          myrow[(ky * 3 + kx) * cin + ci] = __float2bfloat16_rn(v);
        }
      }
    }
  }
  __syncthreads();
  const int words = kpad / 2;
  uint32_t* out32 = reinterpret_cast<uint32_t*>(out);
  for (int w = t; w < 128 * words; w += 128) {
    const int row = w / words, col = w - row * words;
    if (pix0 + row < npix) out32[(pix0 + row) * words + col] = sm32[row * rowwords + col];
  }
}

// The same with the channel count known at compile time: a thread builds the im2col row of its pixel in registers
// (k = tap * CIN + ci is a compile-time index), writes it to shared memory as 16-byte chunks (row pitch kpad*2 + 16
// bytes: conflict-free) and the block copies the 128 rows out with coalesced 16-byte stores.
template <int CIN>
__global__ void __launch_bounds__(128) pack_input_t_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                                           int csrc, int c_lo, int nc, int cat_mode, int B, int H, int W,
                                                           __nv_bfloat16* __restrict__ out, long long npix) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int KPAD = 64 * ((9 * CIN + 63) / 64);
  constexpr int CH = KPAD / 8;        // 16-byte chunks per row
  constexpr int PITCH = CH + 1;       // in 16-byte units
  extern __shared__ uint4 smv[];
  const int t = threadIdx.x;
  const long long pix0 = static_cast<long long>(blockIdx.x) * 128;
  const long long pix = pix0 + t;
  float v[KPAD];
#pragma unroll
  for (int k = 0; k < KPAD; ++k) v[k] = 0.f;
  if (pix < npix) {
    const int x = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float* src;
      int b, c;
      if (cat_mode == 0) {
        src = n < B ? src0 : src1;
        b = n < B ? n : n - B;
        c = c_lo + ci;
      } else {
        src = ci < nc ? src0 : src1;
        b = n;
        c = c_lo + (ci < nc ? ci : ci - nc);
      }
      const float* plane = src + (static_cast<long long>(b) * csrc + c) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + kx - 1;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) v[(ky * 3 + kx) * CIN + ci] = __ldg(plane + static_cast<long long>(yy) * W + xx);
        }
      }
    }
  }
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = v[ch * 8 + j];
    smv[t * PITCH + ch] = pack8(f);
  }
  __syncthreads();
  uint4* o = reinterpret_cast<uint4*>(out) + pix0 * CH;
  const long long nvalid = (npix - pix0 < 128 ? npix - pix0 : 128) * CH;
  for (int i = t; i < 128 * CH; i += 128) {
    if (i < nvalid) o[i] = smv[(i / CH) * PITCH + (i % CH)];
  }
}

// ------------------------------------------------------------------------------------------------
// Weight packing fp32 (reference layouts) -> bf16 GEMM operands.
//   mode 0: conv3x3 forward   w[d0=co][d1=ci][3][3] -> out[co][tap][ci]
//   mode 1: conv3x3 dgrad     out[ci][tap'][co] = w[co][ci][2-ky'][2-kx']
//   mode 2: first layer       out[co][kpad], k = tap*ci_count + ci
//   mode 3: convT forward     w[d0=ci][d1=co][2][2] -> out[tap*co_count + co][ci]
//   mode 4: convT dgrad       out[ci][tap*co_count + co]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pack_weight_value(int mode, const float* __restrict__ w, int d0, int d1, int kpad,
                                                   long long i) {
  float v = 0.f;
  if (mode == 0) {
    const int ci = static_cast<int>(i % d1);
    const int tap = static_cast<int>((i / d1) % 9);
    const int co = static_cast<int>(i / (9ll * d1));
    v = w[(static_cast<long long>(co) * d1 + ci) * 9 + tap];
  } else if (mode == 1) {
    const int co = static_cast<int>(i % d0);
    const int tap = static_cast<int>((i / d0) % 9);
    const int ci = static_cast<int>(i / (9ll * d0));
    v = w[(static_cast<long long>(co) * d1 + ci) * 9 + (8 - tap)];
  } else if (mode == 2) {
    const int k = static_cast<int>(i % kpad);
    const int co = static_cast<int>(i / kpad);
    if (k < 9 * d1) {
      const int tap = k / d1, ci = k - tap * d1;
      v = w[(static_cast<long long>(co) * d1 + ci) * 9 + tap];
    }
  } else if (mode == 3) {
    const int ci = static_cast<int>(i % d0);
    const int n = static_cast<int>(i / d0);
    const int tap = n / d1, co = n - tap * d1;
    v = w[(static_cast<long long>(ci) * d1 + co) * 4 + tap];
  } else {
    const int n = static_cast<int>(i % (4ll * d1));
    const int ci = static_cast<int>(i / (4ll * d1));
    const int tap = n / d1, co = n - tap * d1;
    v = w[(static_cast<long long>(ci) * d1 + co) * 4 + tap];
  }
  return v;
}

__global__ void pack_weights_kernel(int mode, const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int d0,
                                    int d1, int kpad, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  out[i] = __float2bfloat16_rn(pack_weight_value(mode, w, d0, d1, kpad, i));
}

// All weights of a step in ONE launch. Job j owns thread blocks [start_j, start_{j+1}). For the conv / transposed-conv
// layouts a block moves one 32 (d0) x 32 (d1) x taps tile through shared memory: the fp32 source rows are read as
// contiguous runs of 32*taps floats and both operand layouts (forward and input-gradient) are written from the same
// tile in 64-byte runs, so the master weights are read once per step. Other shapes (first layer, odd sizes) use the
// element-wise path, 2048 outputs per block.
constexpr int kPackTile = 32;
constexpr int kPackGenericPerBlock = 2048;

template <int TAPS>
__device__ __forceinline__ void pack_tile_store(int mode, const float* tile, __nv_bfloat16* out, int d0, int d1, int r0,
                                                int c0) {
  // tile[r][cl * TAPS + tap] = w[r0 + r][c0 + cl][tap]; TAPS is a compile-time constant so that the index arithmetic
  // below is multiply-shift, not integer division (the kernel was instruction bound, not bandwidth bound)
  constexpr int pitch = kPackTile * TAPS + 1;
  constexpr int n2 = kPackTile * kPackTile * TAPS / 2;   // two consecutive outputs (one 4-byte store) per iteration
#pragma unroll 4
  for (int i = threadIdx.x; i < n2; i += 256) {
    const int fast = (i & 15) * 2;
    const int rest = i >> 4;
    const int tap = rest % TAPS;
    const int slow = rest / TAPS;
    float v0, v1;
    long long o;
    if (mode == 0) {         // out[co = d0][tap][ci = d1]
      o = (static_cast<long long>(r0 + slow) * TAPS + tap) * d1 + c0 + fast;
      v0 = tile[slow * pitch + fast * TAPS + tap];
      v1 = tile[slow * pitch + (fast + 1) * TAPS + tap];
    } else if (mode == 1) {  // out[ci = d1][tap'][co = d0] = w[co][ci][8 - tap']
      o = (static_cast<long long>(c0 + slow) * TAPS + tap) * d0 + r0 + fast;
      v0 = tile[fast * pitch + slow * TAPS + (TAPS - 1 - tap)];
      v1 = tile[(fast + 1) * pitch + slow * TAPS + (TAPS - 1 - tap)];
    } else if (mode == 3) {  // out[tap * d1 + co][ci = d0]
      o = (static_cast<long long>(tap) * d1 + c0 + slow) * d0 + r0 + fast;
      v0 = tile[fast * pitch + slow * TAPS + tap];
      v1 = tile[(fast + 1) * pitch + slow * TAPS + tap];
    } else {                 // mode 4: out[ci = d0][tap * d1 + co]
      o = static_cast<long long>(r0 + slow) * (TAPS * d1) + static_cast<long long>(tap) * d1 + c0 + fast;
      v0 = tile[slow * pitch + fast * TAPS + tap];
      v1 = tile[slow * pitch + (fast + 1) * TAPS + tap];
    }
    *reinterpret_cast<__nv_bfloat162*>(out + o) = __floats2bfloat162_rn(v0, v1);   // o is even, out 4-byte aligned
  }
}

template <int TAPS>
__device__ __forceinline__ void pack_tile(const PackJob& j, int lb, float* tile) {
  constexpr int pitch = kPackTile * TAPS + 1;
  constexpr int run = kPackTile * TAPS;
  const int tiles1 = j.d1 / kPackTile;
  const int r0 = (lb / tiles1) * kPackTile, c0 = (lb % tiles1) * kPackTile;
#pragma unroll 4
  for (int i = threadIdx.x; i < kPackTile * run; i += 256) {
    const int r = i / run, k = i - r * run;
    tile[r * pitch + k] = __ldg(j.w + (static_cast<long long>(r0 + r) * j.d1 + c0) * TAPS + k);
  }
  __syncthreads();
  pack_tile_store<TAPS>(j.mode, tile, reinterpret_cast<__nv_bfloat16*>(j.out), j.d0, j.d1, r0, c0);
  if (j.out2 != nullptr)
    pack_tile_store<TAPS>(j.mode2, tile, reinterpret_cast<__nv_bfloat16*>(j.out2), j.d0, j.d1, r0, c0);
}

__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const PackJob* __restrict__ jobs, int njobs) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[kPackTile * (kPackTile * 9 + 1)];
  const long long b = blockIdx.x;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].start <= b) lo = mid;
    else hi = mid - 1;
  }
  const PackJob j = jobs[lo];
  const int lb = static_cast<int>(b - j.start);
  const bool tiled = j.mode != 2 && j.d0 % kPackTile == 0 && j.d1 % kPackTile == 0;
  if (!tiled) {
    long long total;
    if (j.mode == 0 || j.mode == 1) total = 9ll * j.d0 * j.d1;
    else if (j.mode == 2) total = static_cast<long long>(j.d0) * j.kpad;
    else total = 4ll * j.d0 * j.d1;
    const long long e0 = static_cast<long long>(lb) * kPackGenericPerBlock;
    for (int k = threadIdx.x; k < kPackGenericPerBlock; k += blockDim.x) {
      const long long li = e0 + k;
      if (li >= total) break;
      reinterpret_cast<__nv_bfloat16*>(j.out)[li] = __float2bfloat16_rn(pack_weight_value(j.mode, j.w, j.d0, j.d1, j.kpad, li));
      if (j.out2 != nullptr)
        reinterpret_cast<__nv_bfloat16*>(j.out2)[li] = __float2bfloat16_rn(pack_weight_value(j.mode2, j.w, j.d0, j.d1, j.kpad, li));
    }
    return;
  }
  if (j.mode == 0 || j.mode == 1) pack_tile<9>(j, lb, tile);
  else pack_tile<4>(j, lb, tile);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics: stage 1 reduces the conv epilogue's per-tile (sum, sumsq) partials over a
// slice of the tiles of one stat-group in fp64; stage 2 finishes, produces mean / invstd and the
// affine (scale, shift) the apply kernel uses, and updates the running statistics group by group.
// Stat-group = one reference module call (timestamp t1 or t2 of the shared-weight encoder).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_reduce_kernel(const float2* __restrict__ partial, int ld, int C,
                                                              int tiles_per_group, int spl,
                                                              double* __restrict__ partial2) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double sh[8][32][2];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int l = threadIdx.x >> 5;
  const int g = blockIdx.y, sp = blockIdx.z;
  const int tb = static_cast<int>(static_cast<long long>(tiles_per_group) * sp / spl);
  const int te = static_cast<int>(static_cast<long long>(tiles_per_group) * (sp + 1) / spl);
  double s = 0.0, q = 0.0;
  if (c < C) {
    const float2* base = partial + static_cast<long long>(g) * tiles_per_group * ld + c;
    // four independent loads in flight per thread; the summation order is fixed (deterministic)
    double s1 = 0.0, q1 = 0.0, s2 = 0.0, q2 = 0.0, s3 = 0.0, q3 = 0.0;
    int t = tb + l;
    for (; t + 24 < te; t += 32) {
      const float2 v0 = __ldg(base + static_cast<long long>(t) * ld);
      const float2 v1 = __ldg(base + static_cast<long long>(t + 8) * ld);
      const float2 v2 = __ldg(base + static_cast<long long>(t + 16) * ld);
      const float2 v3 = __ldg(base + static_cast<long long>(t + 24) * ld);
      s += v0.x; q += v0.y;
      s1 += v1.x; q1 += v1.y;
      s2 += v2.x; q2 += v2.y;
      s3 += v3.x; q3 += v3.y;
    }
    for (; t < te; t += 8) {
      const float2 v = __ldg(base + static_cast<long long>(t) * ld);
      s += v.x;
      q += v.y;
    }
    s = (s + s1) + (s2 + s3);
    q = (q + q1) + (q2 + q3);
  }
  sh[l][threadIdx.x & 31][0] = s;
  sh[l][threadIdx.x & 31][1] = q;
  __syncthreads();
  if (l == 0 && c < C) {
    for (int i = 1; i < 8; ++i) {
      s += sh[i][threadIdx.x][0];
      q += sh[i][threadIdx.x][1];
    }
    double* o = partial2 + ((static_cast<long long>(sp) * gridDim.y + g) * C + c) * 2;
    o[0] = s;
    o[1] = q;
  }
}

// block = 8 channels x 32 lanes: lane sp of a channel's warp holds split sp (spl <= 32), a fixed-order shuffle tree
// sums them (one round of load latency instead of spl dependent ones); lane 0 finishes the channel.
__global__ void __launch_bounds__(256) bn_finalize_kernel(const double* __restrict__ partial2, int spl, int C, int G,
                                                          double count, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ running_mean,
                                                          float* __restrict__ running_var, long long* __restrict__ nbt,
                                                          float momentum, float eps, int train, int order_rev,
                                                          float* __restrict__ mean, float* __restrict__ invstd,
                                                          float* __restrict__ scale, float* __restrict__ shift) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;  // whole warp
  if (!train) {
    // eval: y = (x - running_mean) / sqrt(running_var + eps) * gamma + beta for every group
    if (lane != 0) return;
    const float is = 1.0f / sqrtf(running_var[c] + eps);
    for (int g = 0; g < G; ++g) {
      mean[g * C + c] = running_mean[c];
      invstd[g * C + c] = is;
      const float sc = gamma[c] * is;
      scale[g * C + c] = sc;
      shift[g * C + c] = beta[c] - running_mean[c] * sc;
    }
    return;
  }
  float rm = running_mean[c], rv = running_var[c];
  for (int gi = 0; gi < G; ++gi) {
    const int g = order_rev ? G - 1 - gi : gi;
    double s = 0.0, q = 0.0;
    for (int sp = lane; sp < spl; sp += 32) {
      const double* o = partial2 + ((static_cast<long long>(sp) * G + g) * C + c) * 2;
      s += o[0];
      q += o[1];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
      q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    const double mu = s / count;
    double var = q / count - mu * mu;
    if (var < 0.0) var = 0.0;
    const float is = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float muf = static_cast<float>(mu);
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    rm = (1.f - momentum) * rm + momentum * muf;
    rv = (1.f - momentum) * rv + momentum * static_cast<float>(unbiased);
    if (lane == 0) {
      mean[g * C + c] = muf;
      invstd[g * C + c] = is;
      const float sc = gamma[c] * is;
      scale[g * C + c] = sc;
      shift[g * C + c] = beta[c] - muf * sc;
    }
  }
  if (lane == 0) {
    running_mean[c] = rm;
    running_var[c] = rv;
    if (c == 0 && nbt != nullptr) *nbt += G;
  }
}

// Inference: the fixed affine of every BatchNorm of a network in ONE launch (running statistics only change when the
// network trains, but the engine cannot see that, so the forward plan starts with this launch). Job j owns thread
// blocks [start_j, start_j + ceil(C_j / 256)); y = (x - running_mean) / sqrt(running_var + eps) * gamma + beta for
// every stat-group (utils/evaluation.py:9-10: net.eval()).
__global__ void __launch_bounds__(256) bn_eval_affine_batched_kernel(const BnEvalJob* __restrict__ jobs, int njobs) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].start <= b) lo = mid;
    else hi = mid - 1;
  }
  const BnEvalJob j = jobs[lo];
  const int c = (b - j.start) * 256 + threadIdx.x;
  if (c >= j.C) return;
  const float rm = j.running_mean[c];
  const float is = 1.0f / sqrtf(j.running_var[c] + j.eps);
  const float sc = j.gamma[c] * is;
  const float sh = j.beta[c] - rm * sc;
  for (int g = 0; g < j.G; ++g) {
    j.mean[g * j.C + c] = rm;
    j.invstd[g * j.C + c] = is;
    j.scale[g * j.C + c] = sc;
    j.shift[g * j.C + c] = sh;
  }
}

// Few partial rows per stat-group (the CTA-pair convolution writes one row per CTA and epilogue group, <= 296): one
// kernel does both stages. block = 8 channels x 32 lanes; lane l sums rows l, l + 32, ... in fp64 (independent loads),
// a fixed-order shuffle tree combines the lanes, lane 0 finishes the channel exactly as bn_finalize_kernel does.
__global__ void __launch_bounds__(256) bn_stats_fused_kernel(const float2* __restrict__ partial, int ld, int rows, int C,
                                                             int G, double count, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             float* __restrict__ running_mean,
                                                             float* __restrict__ running_var, long long* __restrict__ nbt,
                                                             float momentum, float eps, int order_rev,
                                                             float* __restrict__ mean, float* __restrict__ invstd,
                                                             float* __restrict__ scale, float* __restrict__ shift) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;  // whole warp
  float rm = running_mean[c], rv = running_var[c];
  // The kernel sits on the critical path between a convolution and its apply pass and moves a few hundred KB: it is
  // pure latency. All partial rows of BOTH stat-groups are requested before anything is summed (rows <= 320: ten
  // independent loads per lane and group in flight), so the launch costs one memory round trip instead of ten.
  constexpr int kMaxK = 10;
  const bool all_at_once = rows <= 32 * kMaxK && G <= 2;
  float2 pre[2][kMaxK];
  if (all_at_once) {
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      const float2* base = partial + static_cast<long long>(gi) * rows * ld + c;
#pragma unroll
      for (int k = 0; k < kMaxK; ++k) {
        const int r = lane + 32 * k;
        pre[gi][k] = (gi < G && r < rows) ? __ldg(base + static_cast<long long>(r) * ld) : make_float2(0.f, 0.f);
      }
    }
  }
  for (int gi = 0; gi < G; ++gi) {
    const int g = order_rev ? G - 1 - gi : gi;
    const float2* base = partial + static_cast<long long>(g) * rows * ld + c;
    double s = 0.0, q = 0.0, s1 = 0.0, q1 = 0.0;
    if (all_at_once) {
      // same pairing of rows as the streaming loop below (rows l, l + 64, ... in one chain, l + 32, l + 96, ... in the
      // other), so both paths give bit-identical statistics
#pragma unroll
      for (int k = 0; k < kMaxK; k += 2) {
        const float2 v0 = g == 0 ? pre[0][k] : pre[1][k];          // static register indices
        const float2 v1 = g == 0 ? pre[0][k + 1] : pre[1][k + 1];
        s += v0.x; q += v0.y;
        s1 += v1.x; q1 += v1.y;
      }
    } else {
      int r = lane;
      for (; r + 32 < rows; r += 64) {
        const float2 v0 = __ldg(base + static_cast<long long>(r) * ld);
        const float2 v1 = __ldg(base + static_cast<long long>(r + 32) * ld);
        s += v0.x; q += v0.y;
        s1 += v1.x; q1 += v1.y;
      }
      if (r < rows) {
        const float2 v0 = __ldg(base + static_cast<long long>(r) * ld);
        s += v0.x; q += v0.y;
      }
    }
    s += s1;
    q += q1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
      q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    const double mu = s / count;
    double var = q / count - mu * mu;
    if (var < 0.0) var = 0.0;
    const float is = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float muf = static_cast<float>(mu);
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    rm = (1.f - momentum) * rm + momentum * muf;
    rv = (1.f - momentum) * rv + momentum * static_cast<float>(unbiased);
    if (lane == 0) {
      mean[g * C + c] = muf;
      invstd[g * C + c] = is;
      const float sc = gamma[c] * is;
      scale[g * C + c] = sc;
      shift[g * C + c] = beta[c] - muf * sc;
    }
  }
  if (lane == 0) {
    running_mean[c] = rm;
    running_var[c] = rv;
    if (c == 0 && nbt != nullptr) *nbt += G;
  }
}

// ------------------------------------------------------------------------------------------------
// BN-apply + ReLU, fused with MaxPool2d(2), the t2 - t1 feature difference and a second copy into a
// concat slice. One thread = one 2x2 pixel window x 8 channels (x both timestamps when diff).
// ------------------------------------------------------------------------------------------------
struct ApplyArgs {
  const __nv_bfloat16* r;
  long long ld_r;
  const float* scale;
  const float* shift;
  int n_img, H, W, C, G, diff;
  __nv_bfloat16 *a, *a2, *pool, *dif;
  long long ld_a, ld_a2, ld_p, ld_d;
  unsigned char* pidx;  // optional [n][H/2][W/2][C]: position (0..3) of the window maximum, for the backward routing
};

__global__ void __launch_bounds__(256) bn_apply_kernel(const ApplyArgs p) {
  pdl_launch_dependents();
  pdl_wait();
  const int cvecs = p.C >> 3;
  const int H2 = (p.H + 1) >> 1, W2 = (p.W + 1) >> 1;
  const int n_units = p.diff ? p.n_img / 2 : p.n_img;
  const long long total = static_cast<long long>(n_units) * H2 * W2 * cvecs;
  const int per_group = p.n_img / p.G;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % cvecs);
    long long w = i / cvecs;
    const int x2 = static_cast<int>(w % W2);
    w /= W2;
    const int y2 = static_cast<int>(w % H2);
    const int n = static_cast<int>(w / H2);
    const int c = cv << 3;
    float amax[2][8];
    float av[2][4][8];
    const int reps = p.diff ? 2 : 1;
    const bool full = (2 * y2 + 1 < p.H) && (2 * x2 + 1 < p.W);
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      if (rep >= reps) break;
      const int nn = n + rep * n_units;
      const int g = nn / per_group;
      float sc[8], sh[8];
      load8f(p.scale + g * p.C + c, sc);
      load8f(p.shift + g * p.C + c, sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) amax[rep][j] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int y = 2 * y2 + (k >> 1), x = 2 * x2 + (k & 1);
        if (y < p.H && x < p.W) {
          const long long pix = (static_cast<long long>(nn) * p.H + y) * p.W + x;
          float rv[8];
          unpack8(ldg16(p.r + pix * p.ld_r + c), rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v = fmaxf(fmaf(rv[j], sc[j], sh[j]), 0.f);
            av[rep][k][j] = v;
            amax[rep][j] = fmaxf(amax[rep][j], v);
          }
          const uint4 o = pack8(av[rep][k]);
          if (p.a) *reinterpret_cast<uint4*>(p.a + pix * p.ld_a + c) = o;
          if (p.a2) *reinterpret_cast<uint4*>(p.a2 + pix * p.ld_a2 + c) = o;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) av[rep][k][j] = 0.f;
        }
      }
      if (p.pool && full) {
        const long long ppix = (static_cast<long long>(nn) * (p.H >> 1) + y2) * (p.W >> 1) + x2;
        *reinterpret_cast<uint4*>(p.pool + ppix * p.ld_p + c) = pack8(amax[rep]);
        if (p.pidx) {
          // first maximum in row-major order of the STORED (bf16) activations, as ATen's max_pool2d_with_indices
          unsigned long long packed = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            int arg = 0;
            float best = round_bf16(av[rep][0][j]);
#pragma unroll
            for (int k = 1; k < 4; ++k) {
              const float v = round_bf16(av[rep][k][j]);
              if (v > best) {
                best = v;
                arg = k;
              }
            }
            packed |= static_cast<unsigned long long>(arg) << (8 * j);
          }
          *reinterpret_cast<unsigned long long*>(p.pidx + ppix * p.C + c) = packed;
        }
      }
    }
    if (p.diff && p.dif) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int y = 2 * y2 + (k >> 1), x = 2 * x2 + (k & 1);
        if (y < p.H && x < p.W) {
          const long long pix = (static_cast<long long>(n) * p.H + y) * p.W + x;
          float d[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = av[1][k][j] - av[0][k][j];
          *reinterpret_cast<uint4*>(p.dif + pix * p.ld_d + c) = pack8(d);
        }
      }
    }
  }
}

// Window variant for even H, W (every training shape): a thread keeps its 8 channels (scale/shift in registers) and
// walks over 2x2 windows of its stat-group with 32-bit index arithmetic; the four (eight with DIFF) 16-byte loads of a
// window are issued before any of them is used. Same arithmetic as bn_apply_kernel (bit-identical outputs).
//   DIFF: grid.y = 1, the block handles window w of image n (timestamp 1) and of image n + n_img/2 (timestamp 2)
//   else: grid.y = G, the block handles the images of one stat-group
template <bool DIFF>
__global__ void __launch_bounds__(256, 3) bn_apply_win_kernel(const ApplyArgs p) {
  pdl_launch_dependents();
  pdl_wait();
  const int cvecs = p.C >> 3;
  const int lanes = 256 / cvecs;
  const int cv = threadIdx.x % cvecs;
  const int l = threadIdx.x / cvecs;
  const int c = cv << 3;
  const int H2 = p.H >> 1, W2 = p.W >> 1;
  const int per_group = p.n_img / p.G;
  const int n_units = DIFF ? p.n_img / 2 : per_group;
  const int img0 = DIFF ? 0 : blockIdx.y * per_group;
  const int nwin = n_units * H2 * W2;
  const int wb = static_cast<int>(static_cast<long long>(nwin) * blockIdx.x / gridDim.x);
  const int we = static_cast<int>(static_cast<long long>(nwin) * (blockIdx.x + 1) / gridDim.x);
  constexpr int REPS = DIFF ? 2 : 1;
  float sc[REPS][8], sh[REPS][8];
#pragma unroll
  for (int rep = 0; rep < REPS; ++rep) {
    const int g = DIFF ? rep : blockIdx.y;
    load8f(p.scale + g * p.C + c, sc[rep]);
    load8f(p.shift + g * p.C + c, sh[rep]);
  }
  for (int w = wb + l; w < we; w += lanes) {
    const int x2 = w % W2;
    const int t = w / W2;
    const int y2 = t % H2;
    const int n = img0 + t / H2;
    uint4 rv[REPS][4];
    long long pix[REPS];
#pragma unroll
    for (int rep = 0; rep < REPS; ++rep) {
      const int nn = n + rep * n_units;
      pix[rep] = (static_cast<long long>(nn) * p.H + 2 * y2) * p.W + 2 * x2;
      const __nv_bfloat16* base = p.r + pix[rep] * p.ld_r + c;
      rv[rep][0] = ldg16(base);
      rv[rep][1] = ldg16(base + p.ld_r);
      rv[rep][2] = ldg16(base + static_cast<long long>(p.W) * p.ld_r);
      rv[rep][3] = ldg16(base + static_cast<long long>(p.W + 1) * p.ld_r);
    }
    float av[REPS][4][8];
#pragma unroll
    for (int rep = 0; rep < REPS; ++rep) {
      const int nn = n + rep * n_units;
      float amax[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) amax[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float r8[8];
        unpack8(rv[rep][k], r8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = fmaxf(fmaf(r8[j], sc[rep][j], sh[rep][j]), 0.f);
          av[rep][k][j] = v;
          amax[j] = fmaxf(amax[j], v);
        }
        const long long px = pix[rep] + (k & 1) + (k >> 1) * p.W;
        const uint4 o = pack8(av[rep][k]);
        if (p.a) *reinterpret_cast<uint4*>(p.a + px * p.ld_a + c) = o;
        if (p.a2) *reinterpret_cast<uint4*>(p.a2 + px * p.ld_a2 + c) = o;
      }
      if (p.pool) {
        const long long ppix = (static_cast<long long>(nn) * H2 + y2) * W2 + x2;
        *reinterpret_cast<uint4*>(p.pool + ppix * p.ld_p + c) = pack8(amax);
        if (p.pidx) {
          // first maximum in row-major order of the STORED (bf16) activations, as ATen's max_pool2d_with_indices
          unsigned long long packed = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            int arg = 0;
            float best = round_bf16(av[rep][0][j]);
#pragma unroll
            for (int k = 1; k < 4; ++k) {
              const float v = round_bf16(av[rep][k][j]);
              if (v > best) {
                best = v;
                arg = k;
              }
            }
            packed |= static_cast<unsigned long long>(arg) << (8 * j);
          }
          *reinterpret_cast<unsigned long long*>(p.pidx + ppix * p.C + c) = packed;
        }
      }
    }
    if (DIFF && p.dif) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = av[REPS - 1][k][j] - av[0][k][j];
        const long long px = pix[0] + (k & 1) + (k >> 1) * p.W;
        *reinterpret_cast<uint4*>(p.dif + px * p.ld_d + c) = pack8(d);
      }
    }
  }
}

// Per-pixel variant (no pooling, no difference): a thread keeps its 8 channels, so scale/shift stay in registers;
// four independent 16-byte loads in flight per thread.
__global__ void __launch_bounds__(256, 4) bn_apply_px_kernel(const ApplyArgs p) {
  pdl_launch_dependents();
  pdl_wait();
  const int cvecs = p.C >> 3;
  const int lanes = 256 / cvecs;
  const int cv = threadIdx.x % cvecs;
  const int l = threadIdx.x / cvecs;
  const int g = blockIdx.y;
  const int c = cv << 3;
  const long long npx = static_cast<long long>(p.n_img / p.G) * p.H * p.W;
  const long long pb = npx * blockIdx.x / gridDim.x, pe = npx * (blockIdx.x + 1) / gridDim.x;
  float sc[8], sh[8];
  load8f(p.scale + g * p.C + c, sc);
  load8f(p.shift + g * p.C + c, sh);
  const long long base = static_cast<long long>(g) * npx;
  const __nv_bfloat16* rb = p.r + base * p.ld_r + c;
  __nv_bfloat16* ab = p.a ? p.a + base * p.ld_a + c : nullptr;
  __nv_bfloat16* a2b = p.a2 ? p.a2 + base * p.ld_a2 + c : nullptr;
  long long i = pb + l;
  for (; i + 3 * lanes < pe; i += 4 * lanes) {
    uint4 rv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) rv[u] = ldg16(rb + (i + u * lanes) * p.ld_r);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float v[8];
      unpack8(rv[u], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
      const uint4 o = pack8(v);
      if (ab) *reinterpret_cast<uint4*>(ab + (i + u * lanes) * p.ld_a) = o;
      if (a2b) *reinterpret_cast<uint4*>(a2b + (i + u * lanes) * p.ld_a2) = o;
    }
  }
  for (; i < pe; i += lanes) {
    float v[8];
    unpack8(ldg16(rb + i * p.ld_r), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
    const uint4 o = pack8(v);
    if (ab) *reinterpret_cast<uint4*>(ab + i * p.ld_a) = o;
    if (a2b) *reinterpret_cast<uint4*>(a2b + i * p.ld_a2) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// BN + ReLU backward. dy = (sum of gradient sources) * [y > 0].
//   pass 1 (reduce): per (stat-group, channel) S1 = sum dy, S2 = sum dy*r   (block partials, fp32)
//   finalize (fp64): sum(dy*xhat) = invstd*(S2 - mean*S1); dgamma, dbeta; and the two per-channel
//                    coefficients of pass 2:  A = -scale*m2*invstd,  B = scale*(m2*mean*invstd - m1)
//                    with m1 = S1/N, m2 = sum(dy*xhat)/N
//   pass 2 (dx):     dr = scale*dy + r*A + B      (= scale*(dy - m1 - xhat*m2)), stored bf16
// Gradient sources are gathered on the fly (skip connections with sign, max-pool routing through the
// stored arg-max index, 1x1 head), so neither the summed gradient nor the ReLU mask nor the post-BN
// activation is ever materialised. A thread keeps its 8 channels for its whole pixel range (per-channel
// constants live in registers) and handles 4 pixels per iteration (independent 16-byte loads in flight).
// ------------------------------------------------------------------------------------------------
struct BwdArgs {
  const __nv_bfloat16* r;
  long long ld_r;
  const float *scale, *shift;  // [G][C]; y = r*scale + shift exactly as the forward apply kernel computes it
  GradSrcs srcs;
  int n_img, H, W, C, G;
};

// The kernels are instantiated for the source-kind triples the networks produce (K0, K1, K2 >= 0: known at compile
// time, the dead branches and their registers disappear) plus a generic variant (K < 0: kinds read at run time).
template <int K0, int K1, int K2>
__device__ __forceinline__ int src_kind(const BwdArgs& p, int si) {
  const int k = si == 0 ? K0 : (si == 1 ? K1 : K2);
  return k >= 0 ? k : p.srcs.s[si].kind;
}

// sum of the gradient sources at global pixel `gpix` (= n*H*W + y*W + x) -> dy[8] (unmasked)
template <int K0, int K1, int K2>
__device__ __forceinline__ void gather_px(const BwdArgs& p, int gpix, int hw, int c, float (&dy)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) dy[j] = 0.f;
#pragma unroll
  for (int si = 0; si < 3; ++si) {
    const GradSrc& s = p.srcs.s[si];
    const int kind = src_kind<K0, K1, K2>(p, si);
    if (kind == 0) continue;
    int sp = gpix;  // pixel index in the source tensor
    float scale = 1.f;
    if (s.n_mod > 0) {
      const int n = gpix / hw;
      scale = n < s.n_mod ? s.scale_lo : s.scale_hi;
      sp = gpix - (n - n % s.n_mod) * hw;
    }
    if (kind == 1) {
      float gv[8];
      unpack8(ldg16(reinterpret_cast<const __nv_bfloat16*>(s.ptr) + static_cast<long long>(sp) * s.ld + c), gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[j] = fmaf(scale, gv[j], dy[j]);
    } else if (kind == 2) {
      const int n = sp / hw, pix = sp - n * hw;
      const int y = pix / p.W, x = pix - y * p.W;
      const int H2 = p.H >> 1, W2 = p.W >> 1;
      if ((y >> 1) < H2 && (x >> 1) < W2) {  // MaxPool2d floors: the last odd row/column feeds no window
        const long long pp = (static_cast<long long>(n) * H2 + (y >> 1)) * W2 + (x >> 1);
        const unsigned long long idx =
            __ldg(reinterpret_cast<const unsigned long long*>(reinterpret_cast<const unsigned char*>(s.w) + pp * p.C + c));
        float gv[8];
        unpack8(ldg16(reinterpret_cast<const __nv_bfloat16*>(s.ptr) + pp * s.ld + c), gv);
        const unsigned me = ((y & 1) << 1) | (x & 1);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dy[j] += (((idx >> (8 * j)) & 0xffull) == me) ? scale * gv[j] : 0.f;
      }
    } else {
      const float d = scale * __ldg(reinterpret_cast<const float*>(s.ptr) + sp);
      float wv[8];
      load8f(s.w + c, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[j] = fmaf(d, wv[j], dy[j]);
    }
  }
}

constexpr int kBwdUnroll = 4;

// The same for kBwdUnroll pixels at once, source by source, with all loads of a source issued before any is used
// (memory-level parallelism: the kernels are pure streaming).
template <int K0, int K1, int K2>
__device__ __forceinline__ void gather_multi(const BwdArgs& p, const int (&gpix)[kBwdUnroll], int hw, int c,
                                             float (&dy)[kBwdUnroll][8]) {
#pragma unroll
  for (int u = 0; u < kBwdUnroll; ++u)
#pragma unroll
    for (int j = 0; j < 8; ++j) dy[u][j] = 0.f;
#pragma unroll
  for (int si = 0; si < 3; ++si) {
    const GradSrc& s = p.srcs.s[si];
    const int kind = src_kind<K0, K1, K2>(p, si);
    if (kind == 0) continue;
    int sp[kBwdUnroll];
    float scale[kBwdUnroll];
#pragma unroll
    for (int u = 0; u < kBwdUnroll; ++u) {
      sp[u] = gpix[u];
      scale[u] = 1.f;
      if (s.n_mod > 0) {
        const int n = gpix[u] / hw;
        scale[u] = n < s.n_mod ? s.scale_lo : s.scale_hi;
        sp[u] = gpix[u] - (n - n % s.n_mod) * hw;
      }
    }
    if (kind == 1) {
      uint4 g[kBwdUnroll];
#pragma unroll
      for (int u = 0; u < kBwdUnroll; ++u)
        g[u] = ldg16(reinterpret_cast<const __nv_bfloat16*>(s.ptr) + static_cast<long long>(sp[u]) * s.ld + c);
#pragma unroll
      for (int u = 0; u < kBwdUnroll; ++u) {
        float gv[8];
        unpack8(g[u], gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[u][j] = fmaf(scale[u], gv[j], dy[u][j]);
      }
    } else if (kind == 2) {
      const int H2 = p.H >> 1, W2 = p.W >> 1;
      uint4 g[kBwdUnroll];
      unsigned long long idx[kBwdUnroll];
      unsigned me[kBwdUnroll];
      bool in[kBwdUnroll];
#pragma unroll
      for (int u = 0; u < kBwdUnroll; ++u) {
        const int n = sp[u] / hw, pix = sp[u] - n * hw;
        const int y = pix / p.W, x = pix - y * p.W;
        in[u] = (y >> 1) < H2 && (x >> 1) < W2;  // MaxPool2d floors: the last odd row/column feeds no window
        const long long pp = in[u] ? (static_cast<long long>(n) * H2 + (y >> 1)) * W2 + (x >> 1) : 0;
        me[u] = ((y & 1) << 1) | (x & 1);
        idx[u] = __ldg(reinterpret_cast<const unsigned long long*>(reinterpret_cast<const unsigned char*>(s.w) + pp * p.C + c));
        g[u] = ldg16(reinterpret_cast<const __nv_bfloat16*>(s.ptr) + pp * s.ld + c);
      }
#pragma unroll
      for (int u = 0; u < kBwdUnroll; ++u) {
        float gv[8];
        unpack8(g[u], gv);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dy[u][j] += (in[u] && ((idx[u] >> (8 * j)) & 0xffull) == me[u]) ? scale[u] * gv[j] : 0.f;
      }
    } else {
      float d[kBwdUnroll];
#pragma unroll
      for (int u = 0; u < kBwdUnroll; ++u) d[u] = scale[u] * __ldg(reinterpret_cast<const float*>(s.ptr) + sp[u]);
      float wv[8];
      load8f(s.w + c, wv);
#pragma unroll
      for (int u = 0; u < kBwdUnroll; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[u][j] = fmaf(d[u], wv[j], dy[u][j]);
    }
  }
}

// grid = (nblk, G); block = 256 threads = (C/8 channel vectors) x (256/(C/8)) pixel lanes
template <int K0, int K1, int K2>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(const BwdArgs p, float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float shred[];  // [lanes][C][2]
  const int cvecs = p.C >> 3;
  const int lanes = 256 / cvecs;
  const int cv = threadIdx.x % cvecs;
  const int l = threadIdx.x / cvecs;
  const int g = blockIdx.y;
  const int c = cv << 3;
  const int hw = p.H * p.W;
  const int npx = (p.n_img / p.G) * hw;
  const int gbase = g * npx;
  const int pb = static_cast<int>(static_cast<long long>(npx) * blockIdx.x / gridDim.x);
  const int pe = static_cast<int>(static_cast<long long>(npx) * (blockIdx.x + 1) / gridDim.x);
  float sc[8], sh[8], s1[8], s2[8];
  load8f(p.scale + g * p.C + c, sc);
  load8f(p.shift + g * p.C + c, sh);
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  const __nv_bfloat16* rbase = p.r + static_cast<long long>(gbase) * p.ld_r + c;
  int i = pb + l;
  for (; i + (kBwdUnroll - 1) * lanes < pe; i += kBwdUnroll * lanes) {
    uint4 rr[kBwdUnroll];
    int gp[kBwdUnroll];
#pragma unroll
    for (int u = 0; u < kBwdUnroll; ++u) {
      rr[u] = ldg16(rbase + static_cast<long long>(i + u * lanes) * p.ld_r);
      gp[u] = gbase + i + u * lanes;
    }
    float d[kBwdUnroll][8];
    gather_multi<K0, K1, K2>(p, gp, hw, c, d);
#pragma unroll
    for (int u = 0; u < kBwdUnroll; ++u) {
      float v[8];
      unpack8(rr[u], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float m = fmaf(v[j], sc[j], sh[j]) > 0.f ? d[u][j] : 0.f;
        s1[j] += m;
        s2[j] = fmaf(m, v[j], s2[j]);
      }
    }
  }
  for (; i < pe; i += lanes) {
    float d[8], v[8];
    unpack8(ldg16(rbase + static_cast<long long>(i) * p.ld_r), v);
    gather_px<K0, K1, K2>(p, gbase + i, hw, c, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = fmaf(v[j], sc[j], sh[j]) > 0.f ? d[j] : 0.f;
      s1[j] += m;
      s2[j] = fmaf(m, v[j], s2[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    shred[(l * p.C + c + j) * 2] = s1[j];
    shred[(l * p.C + c + j) * 2 + 1] = s2[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < p.C; ch += 256) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < lanes; ++k) {
      a += shred[(k * p.C + ch) * 2];
      b += shred[(k * p.C + ch) * 2 + 1];
    }
    float* o = partial + ((static_cast<long long>(g) * gridDim.x + blockIdx.x) * p.C + ch) * 2;
    o[0] = a;
    o[1] = b;
  }
}

// Window variants for layers whose first gradient source is a max-pool (kind 2; H, W even): a thread walks over 2x2
// windows, so the pooled gradient and its arg-max index are loaded once per window instead of once per pixel, index
// arithmetic is per window and 32-bit, and all loads of a window (4 x r, pooled gradient, index, 4 per direct source)
// are in flight together. K1, K2 in {0, 1}. Same arithmetic per element as the per-pixel kernels.
template <int K1, int K2>
__device__ __forceinline__ void gather_window(const BwdArgs& p, int n, int y2, int x2, int c, const uint4 (&rr)[4],
                                              const float (&sc)[8], const float (&sh)[8], float (&m)[4][8]) {
  const int H2 = p.H >> 1, W2 = p.W >> 1;
  const int hw = p.H * p.W;
  // issue every load first
  const GradSrc& s0 = p.srcs.s[0];
  int n0 = n;
  float scale0 = 1.f;
  if (s0.n_mod > 0) {
    scale0 = n < s0.n_mod ? s0.scale_lo : s0.scale_hi;
    n0 = n % s0.n_mod;
  }
  const long long pp = (static_cast<long long>(n0) * H2 + y2) * W2 + x2;
  const unsigned long long idx =
      __ldg(reinterpret_cast<const unsigned long long*>(reinterpret_cast<const unsigned char*>(s0.w) + pp * p.C + c));
  const uint4 g0 = ldg16(reinterpret_cast<const __nv_bfloat16*>(s0.ptr) + pp * s0.ld + c);
  uint4 g1[4], g2[4];
  float scale1 = 1.f, scale2 = 1.f;
  if (K1 == 1) {
    const GradSrc& s = p.srcs.s[1];
    int nn = n;
    if (s.n_mod > 0) {
      scale1 = n < s.n_mod ? s.scale_lo : s.scale_hi;
      nn = n % s.n_mod;
    }
    const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(s.ptr) +
                             (static_cast<long long>(nn) * hw + 2 * y2 * p.W + 2 * x2) * s.ld + c;
    g1[0] = ldg16(b);
    g1[1] = ldg16(b + s.ld);
    g1[2] = ldg16(b + static_cast<long long>(p.W) * s.ld);
    g1[3] = ldg16(b + static_cast<long long>(p.W + 1) * s.ld);
  }
  if (K2 == 1) {
    const GradSrc& s = p.srcs.s[2];
    int nn = n;
    if (s.n_mod > 0) {
      scale2 = n < s.n_mod ? s.scale_lo : s.scale_hi;
      nn = n % s.n_mod;
    }
    const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(s.ptr) +
                             (static_cast<long long>(nn) * hw + 2 * y2 * p.W + 2 * x2) * s.ld + c;
    g2[0] = ldg16(b);
    g2[1] = ldg16(b + s.ld);
    g2[2] = ldg16(b + static_cast<long long>(p.W) * s.ld);
    g2[3] = ldg16(b + static_cast<long long>(p.W + 1) * s.ld);
  }
  float gp[8];
  unpack8(g0, gp);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float dy[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dy[j] = 0.f;
    // same accumulation order as gather_multi: source 0 (pool), then 1, then 2
#pragma unroll
    for (int j = 0; j < 8; ++j) dy[j] += (((idx >> (8 * j)) & 0xffull) == static_cast<unsigned>(k)) ? scale0 * gp[j] : 0.f;
    if (K1 == 1) {
      float gv[8];
      unpack8(g1[k], gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[j] = fmaf(scale1, gv[j], dy[j]);
    }
    if (K2 == 1) {
      float gv[8];
      unpack8(g2[k], gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[j] = fmaf(scale2, gv[j], dy[j]);
    }
    float v[8];
    unpack8(rr[k], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[k][j] = fmaf(v[j], sc[j], sh[j]) > 0.f ? dy[j] : 0.f;
  }
}

template <int K1, int K2>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_win_kernel(const BwdArgs p, float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float shred[];  // [lanes][C][2]
  const int cvecs = p.C >> 3;
  const int lanes = 256 / cvecs;
  const int cv = threadIdx.x % cvecs;
  const int l = threadIdx.x / cvecs;
  const int g = blockIdx.y;
  const int c = cv << 3;
  const int H2 = p.H >> 1, W2 = p.W >> 1;
  const int per_group = p.n_img / p.G;
  const int nwin = per_group * H2 * W2;
  const int wb = static_cast<int>(static_cast<long long>(nwin) * blockIdx.x / gridDim.x);
  const int we = static_cast<int>(static_cast<long long>(nwin) * (blockIdx.x + 1) / gridDim.x);
  float sc[8], sh[8], s1[8], s2[8];
  load8f(p.scale + g * p.C + c, sc);
  load8f(p.shift + g * p.C + c, sh);
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  for (int w = wb + l; w < we; w += lanes) {
    const int x2 = w % W2;
    const int t = w / W2;
    const int y2 = t % H2;
    const int n = g * per_group + t / H2;
    const __nv_bfloat16* rb = p.r + ((static_cast<long long>(n) * p.H + 2 * y2) * p.W + 2 * x2) * p.ld_r + c;
    uint4 rr[4];
    rr[0] = ldg16(rb);
    rr[1] = ldg16(rb + p.ld_r);
    rr[2] = ldg16(rb + static_cast<long long>(p.W) * p.ld_r);
    rr[3] = ldg16(rb + static_cast<long long>(p.W + 1) * p.ld_r);
    float m[4][8];
    gather_window<K1, K2>(p, n, y2, x2, c, rr, sc, sh, m);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8];
      unpack8(rr[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += m[k][j];
        s2[j] = fmaf(m[k][j], v[j], s2[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    shred[(l * p.C + c + j) * 2] = s1[j];
    shred[(l * p.C + c + j) * 2 + 1] = s2[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < p.C; ch += 256) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < lanes; ++k) {
      a += shred[(k * p.C + ch) * 2];
      b += shred[(k * p.C + ch) * 2 + 1];
    }
    float* o = partial + ((static_cast<long long>(g) * gridDim.x + blockIdx.x) * p.C + ch) * 2;
    o[0] = a;
    o[1] = b;
  }
}

template <int K1, int K2>
__global__ void __launch_bounds__(256, 2) bn_bwd_dx_win_kernel(const BwdArgs p, const float* __restrict__ coefA,
                                                               const float* __restrict__ coefB,
                                                               __nv_bfloat16* __restrict__ dr, long long ld_dr) {
  pdl_launch_dependents();
  pdl_wait();
  const int cvecs = p.C >> 3;
  const int lanes = 256 / cvecs;
  const int cv = threadIdx.x % cvecs;
  const int l = threadIdx.x / cvecs;
  const int g = blockIdx.y;
  const int c = cv << 3;
  const int H2 = p.H >> 1, W2 = p.W >> 1;
  const int per_group = p.n_img / p.G;
  const int nwin = per_group * H2 * W2;
  const int wb = static_cast<int>(static_cast<long long>(nwin) * blockIdx.x / gridDim.x);
  const int we = static_cast<int>(static_cast<long long>(nwin) * (blockIdx.x + 1) / gridDim.x);
  float sc[8], sh[8], ca[8], cb[8];
  load8f(p.scale + g * p.C + c, sc);
  load8f(p.shift + g * p.C + c, sh);
  load8f(coefA + g * p.C + c, ca);
  load8f(coefB + g * p.C + c, cb);
  for (int w = wb + l; w < we; w += lanes) {
    const int x2 = w % W2;
    const int t = w / W2;
    const int y2 = t % H2;
    const int n = g * per_group + t / H2;
    const long long pix = (static_cast<long long>(n) * p.H + 2 * y2) * p.W + 2 * x2;
    const __nv_bfloat16* rb = p.r + pix * p.ld_r + c;
    uint4 rr[4];
    rr[0] = ldg16(rb);
    rr[1] = ldg16(rb + p.ld_r);
    rr[2] = ldg16(rb + static_cast<long long>(p.W) * p.ld_r);
    rr[3] = ldg16(rb + static_cast<long long>(p.W + 1) * p.ld_r);
    float m[4][8];
    gather_window<K1, K2>(p, n, y2, x2, c, rr, sc, sh, m);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8], o[8];
      unpack8(rr[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(sc[j], m[k][j], fmaf(v[j], ca[j], cb[j]));
      *reinterpret_cast<uint4*>(dr + (pix + (k & 1) + (k >> 1) * p.W) * ld_dr + c) = pack8(o);
    }
  }
}

// block = 32 channels x 32 lanes: the lanes split the per-block partials, then a fixed-order tree in shared memory
__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, int G,
                                                             double count, const float* __restrict__ mean,
                                                             const float* __restrict__ invstd,
                                                             const float* __restrict__ scale, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, float* __restrict__ coefA,
                                                             float* __restrict__ coefB) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double sh[32][33][2];
  const int cx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double tg = 0.0, tb = 0.0;
  for (int g = 0; g < G; ++g) {
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
      const float2* base = reinterpret_cast<const float2*>(partial) + static_cast<long long>(g) * nblk * C + c;
      double a1 = 0.0, b1 = 0.0, a2 = 0.0, b2 = 0.0, a3 = 0.0, b3 = 0.0;
      int i = ly;
      for (; i + 96 < nblk; i += 128) {  // four independent loads in flight; fixed order
        const float2 o0 = __ldg(base + static_cast<long long>(i) * C);
        const float2 o1 = __ldg(base + static_cast<long long>(i + 32) * C);
        const float2 o2 = __ldg(base + static_cast<long long>(i + 64) * C);
        const float2 o3 = __ldg(base + static_cast<long long>(i + 96) * C);
        s1 += o0.x; s2 += o0.y;
        a1 += o1.x; b1 += o1.y;
        a2 += o2.x; b2 += o2.y;
        a3 += o3.x; b3 += o3.y;
      }
      for (; i < nblk; i += 32) {
        const float2 o = __ldg(base + static_cast<long long>(i) * C);
        s1 += o.x;
        s2 += o.y;
      }
      s1 = (s1 + a1) + (a2 + a3);
      s2 = (s2 + b1) + (b2 + b3);
    }
    sh[ly][cx][0] = s1;
    sh[ly][cx][1] = s2;
    __syncthreads();
    if (ly == 0 && c < C) {
      s1 = 0.0;
      s2 = 0.0;
      for (int k = 0; k < 32; ++k) {
        s1 += sh[k][cx][0];
        s2 += sh[k][cx][1];
      }
      const double mu = mean[g * C + c], is = invstd[g * C + c], sc = scale[g * C + c];
      const double sdyx = is * (s2 - mu * s1);  // sum dy * xhat
      const double m1 = s1 / count, m2 = sdyx / count;
      coefA[g * C + c] = static_cast<float>(-sc * m2 * is);
      coefB[g * C + c] = static_cast<float>(sc * (m2 * mu * is - m1));
      tb += s1;
      tg += sdyx;
    }
    __syncthreads();
  }
  if (ly == 0 && c < C) {
    dgamma[c] = static_cast<float>(tg);
    dbeta[c] = static_cast<float>(tb);
  }
}

template <int K0, int K1, int K2>
__global__ void __launch_bounds__(256, 2) bn_bwd_dx_kernel(const BwdArgs p, const float* __restrict__ coefA,
                                                           const float* __restrict__ coefB,
                                                           __nv_bfloat16* __restrict__ dr, long long ld_dr) {
  pdl_launch_dependents();
  pdl_wait();
  const int cvecs = p.C >> 3;
  const int lanes = 256 / cvecs;
  const int cv = threadIdx.x % cvecs;
  const int l = threadIdx.x / cvecs;
  const int g = blockIdx.y;
  const int c = cv << 3;
  const int hw = p.H * p.W;
  const int npx = (p.n_img / p.G) * hw;
  const int gbase = g * npx;
  const int pb = static_cast<int>(static_cast<long long>(npx) * blockIdx.x / gridDim.x);
  const int pe = static_cast<int>(static_cast<long long>(npx) * (blockIdx.x + 1) / gridDim.x);
  float sc[8], sh[8], ca[8], cb[8];
  load8f(p.scale + g * p.C + c, sc);
  load8f(p.shift + g * p.C + c, sh);
  load8f(coefA + g * p.C + c, ca);
  load8f(coefB + g * p.C + c, cb);
  const __nv_bfloat16* rbase = p.r + static_cast<long long>(gbase) * p.ld_r + c;
  __nv_bfloat16* obase = dr + static_cast<long long>(gbase) * ld_dr + c;
  int i = pb + l;
  for (; i + (kBwdUnroll - 1) * lanes < pe; i += kBwdUnroll * lanes) {
    uint4 rr[kBwdUnroll];
    int gp[kBwdUnroll];
#pragma unroll
    for (int u = 0; u < kBwdUnroll; ++u) {
      rr[u] = ldg16(rbase + static_cast<long long>(i + u * lanes) * p.ld_r);
      gp[u] = gbase + i + u * lanes;
    }
    float d[kBwdUnroll][8];
    gather_multi<K0, K1, K2>(p, gp, hw, c, d);
#pragma unroll
    for (int u = 0; u < kBwdUnroll; ++u) {
      float v[8], o[8];
      unpack8(rr[u], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float m = fmaf(v[j], sc[j], sh[j]) > 0.f ? d[u][j] : 0.f;
        o[j] = fmaf(sc[j], m, fmaf(v[j], ca[j], cb[j]));
      }
      *reinterpret_cast<uint4*>(obase + static_cast<long long>(i + u * lanes) * ld_dr) = pack8(o);
    }
  }
  for (; i < pe; i += lanes) {
    float d[8], v[8], o[8];
    unpack8(ldg16(rbase + static_cast<long long>(i) * p.ld_r), v);
    gather_px<K0, K1, K2>(p, gbase + i, hw, c, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = fmaf(v[j], sc[j], sh[j]) > 0.f ? d[j] : 0.f;
      o[j] = fmaf(sc[j], m, fmaf(v[j], ca[j], cb[j]));
    }
    *reinterpret_cast<uint4*>(obase + static_cast<long long>(i) * ld_dr) = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------------
// Centre pad of Up (utils/networks.py:440-443): the transposed-conv output [n][h][w][C] placed at (top, left) of a
// [n][H][W] window with row pitch W and pixel stride ld_d (the upper half of a concat buffer); the border is zero.
// Only inference on tiles whose size is not a multiple of 16 needs it (odd levels: MaxPool floors, Up pads).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pad_copy_kernel(const __nv_bfloat16* __restrict__ src, long long ld_s, int n_img,
                                                       int h, int w, int C, __nv_bfloat16* __restrict__ dst,
                                                       long long ld_d, int H, int W, int top, int left) {
  pdl_launch_dependents();
  pdl_wait();
  const int cvecs = C >> 3;
  const long long total = static_cast<long long>(n_img) * H * W * cvecs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % cvecs);
    long long px = i / cvecs;
    const int x = static_cast<int>(px % W);
    const int y = static_cast<int>((px / W) % H);
    const int n = static_cast<int>(px / (static_cast<long long>(W) * H));
    const int ys = y - top, xs = x - left;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ys >= 0 && ys < h && xs >= 0 && xs < w)
      v = ldg16(src + ((static_cast<long long>(n) * h + ys) * w + xs) * ld_s + cv * 8);
    *reinterpret_cast<uint4*>(dst + px * ld_d + cv * 8) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// 1x1 head (OutConv, C -> 1) over one or two 64..128-channel inputs; 8 lanes per pixel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(const __nv_bfloat16* __restrict__ a0, long long ld0,
                                                       const __nv_bfloat16* __restrict__ a1, long long ld1, int C,
                                                       const float* __restrict__ w, const float* __restrict__ b,
                                                       long long npix, float* __restrict__ logits) {
  pdl_launch_dependents();
  pdl_wait();
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long pix = gid >> 3;
  const int sub = static_cast<int>(gid & 7);
  float acc = 0.f;
  if (pix < npix) {
    for (int c = sub * 8; c < C; c += 64) {
      float av[8], wv[8];
      unpack8(ldg16(a0 + pix * ld0 + c), av);
      load8f(w + c, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(av[j], wv[j], acc);
    }
    if (a1 != nullptr) {
      for (int c = sub * 8; c < C; c += 64) {
        float av[8], wv[8];
        unpack8(ldg16(a1 + pix * ld1 + c), av);
        load8f(w + C + c, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(av[j], wv[j], acc);
      }
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (sub == 0 && pix < npix) logits[pix] = acc + __ldg(b);
}

// ------------------------------------------------------------------------------------------------
// Weighted column sums out[c] = sum_pixels wgt[pixel] * x[pixel, c]  (x bf16 or absent):
//   transposed-conv bias gradient (wgt = null), head weight gradient (wgt = dz), head bias gradient
//   (x = null, C = 1). Two-stage: partial[nblk][C] then fixed-order finalize.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int C,
                                                     const float* __restrict__ wgt, long long npix,
                                                     float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float shred[];
  const long long pb = npix * blockIdx.x / gridDim.x, pe = npix * (blockIdx.x + 1) / gridDim.x;
  if (x == nullptr) {
    float s = 0.f;
    for (long long i = pb + threadIdx.x; i < pe; i += 256) s += __ldg(wgt + i);
    shred[threadIdx.x] = s;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
      if (threadIdx.x < st) shred[threadIdx.x] += shred[threadIdx.x + st];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = shred[0];
    return;
  }
  const int cvecs = C >> 3;
  const int lanes = 256 / cvecs;
  const int cv = threadIdx.x % cvecs, l = threadIdx.x / cvecs;
  const int c = cv << 3;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (l < lanes) {
    long long i = pb + l;
    for (; i + 3 * lanes < pe; i += 4 * lanes) {  // four independent 16-byte loads in flight
      uint4 xv[4];
      float wg[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xv[u] = ldg16(x + (i + u * lanes) * ld + c);
        wg[u] = wgt ? __ldg(wgt + i + u * lanes) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[8];
        unpack8(xv[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = fmaf(wg[u], v[j], s[j]);
      }
    }
    for (; i < pe; i += lanes) {
      float v[8];
      unpack8(ldg16(x + i * ld + c), v);
      const float wg = wgt ? __ldg(wgt + i) : 1.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = fmaf(wg, v[j], s[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) shred[l * C + c + j] = s[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += 256) {
    float a = 0.f;
    for (int i = 0; i < lanes; ++i) a += shred[i * C + ch];
    partial[static_cast<long long>(blockIdx.x) * C + ch] = a;
  }
}

__global__ void __launch_bounds__(1024) colsum_finalize_kernel(const float* __restrict__ partial, int nblk, int C,
                                                             float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double sh[32][33];
  const int cx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double a = 0.0;
  if (c < C) {
    double a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int i = ly;
    for (; i + 96 < nblk; i += 128) {
      const float v0 = __ldg(partial + static_cast<long long>(i) * C + c);
      const float v1 = __ldg(partial + static_cast<long long>(i + 32) * C + c);
      const float v2 = __ldg(partial + static_cast<long long>(i + 64) * C + c);
      const float v3 = __ldg(partial + static_cast<long long>(i + 96) * C + c);
      a += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    for (; i < nblk; i += 32) a += __ldg(partial + static_cast<long long>(i) * C + c);
    a = (a + a1) + (a2 + a3);
  }
  sh[ly][cx] = a;
  __syncthreads();
  if (ly == 0 && c < C) {
    a = 0.0;
    for (int k = 0; k < 32; ++k) a += sh[k][cx];
    out[c] = static_cast<float>(a);
  }
}

// Column sums from the per-CTA statistics rows a convolution already produced: out[j] = sum_r stats[r][c_off + j].x
// (the transposed-conv bias gradient is the pixel sum of the upper half of the concat-buffer gradient, which the
// preceding input-gradient convolution wrote together with its per-channel sums). block = 8 channels x 32 lanes.
__global__ void __launch_bounds__(256) stat_rowsum_kernel(const float2* __restrict__ stats, int rows, int ld, int c_off,
                                                          int C, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;  // whole warp
  double s = 0.0, s1 = 0.0;
  const float2* base = stats + c_off + c;
  int r = lane;
  for (; r + 32 < rows; r += 64) {
    s += __ldg(base + static_cast<long long>(r) * ld).x;
    s1 += __ldg(base + static_cast<long long>(r + 32) * ld).x;
  }
  if (r < rows) s += __ldg(base + static_cast<long long>(r) * ld).x;
  s += s1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[c] = static_cast<float>(s);
}

// ------------------------------------------------------------------------------------------------
// Split reduction of the weight-gradient workspace into the reference parameter layouts.
//   layout 0: ws[s][tap][d0][d1]          -> grad[d0][d1][tap]   (conv3x3: [co][ci][3][3]; convT: [ci][co][2][2])
//   layout 1: ws[s][d0][ld1 >= taps*d1], k = tap*d1 + i -> grad[d0][d1][tap]  (first layer; split_stride = d0*ld1)
// ------------------------------------------------------------------------------------------------
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, long long split_stride, int layout,
                                    int d0, int d1, int taps, float* __restrict__ grad, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i enumerates the workspace order (coalesced reads); writes are the strided side
  long long src, dst;
  if (layout == 0) {
    const int b = static_cast<int>(i % d1);
    const int a = static_cast<int>((i / d1) % d0);
    const int tap = static_cast<int>(i / (static_cast<long long>(d1) * d0));
    src = i;
    dst = (static_cast<long long>(a) * d1 + b) * taps + tap;
  } else {
    const int k = static_cast<int>(i % (taps * d1));
    const int a = static_cast<int>(i / (taps * d1));
    const int tap = k / d1, b = k - tap * d1;
    const long long ld1 = split_stride / d0;
    src = static_cast<long long>(a) * ld1 + k;
    dst = (static_cast<long long>(a) * d1 + b) * taps + tap;
  }
  // fixed summation order (deterministic); four independent chains keep several loads in flight
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const float* q = ws + src;
  int sp = 0;
#pragma unroll 2
  for (; sp + 3 < splits; sp += 4) {
    a0 += __ldg(q + static_cast<long long>(sp) * split_stride);
    a1 += __ldg(q + static_cast<long long>(sp + 1) * split_stride);
    a2 += __ldg(q + static_cast<long long>(sp + 2) * split_stride);
    a3 += __ldg(q + static_cast<long long>(sp + 3) * split_stride);
  }
  for (; sp < splits; ++sp) a0 += __ldg(q + static_cast<long long>(sp) * split_stride);
  grad[dst] = (a0 + a1) + (a2 + a3);
}

// Vectorised variant (total % 4 == 0, 16-byte aligned workspace): block = 32 lanes x 8 parts; a lane owns four
// consecutive workspace elements (one 16-byte load per split), part p sums splits p, p+8, ... in two chains, then a
// fixed-order sum over the parts in shared memory. Same result on every run.
__global__ void __launch_bounds__(256) wgrad_reduce_v4_kernel(const float* __restrict__ ws, int splits,
                                                              long long split_stride, int layout, int d0, int d1,
                                                              int taps, float* __restrict__ grad, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float4 sh[8][33];
  const int e = threadIdx.x & 31, part = threadIdx.x >> 5;
  const long long i = (static_cast<long long>(blockIdx.x) * 32 + e) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  long long src = 0;
  if (i < total) {
    if (layout == 0) {
      src = i;
    } else {
      const int k = static_cast<int>(i % (taps * d1));
      const int a = static_cast<int>(i / (taps * d1));
      src = static_cast<long long>(a) * (split_stride / d0) + k;
    }
    const float* q = ws + src;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    int sp = part;
    for (; sp + 8 < splits; sp += 16) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(q + static_cast<long long>(sp) * split_stride));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(q + static_cast<long long>(sp + 8) * split_stride));
      acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
      b.x += v1.x; b.y += v1.y; b.z += v1.z; b.w += v1.w;
    }
    if (sp < splits) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(q + static_cast<long long>(sp) * split_stride));
      acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
    }
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  sh[part][e] = acc;
  __syncthreads();
  if (part == 0 && i < total) {
    float4 r = sh[0][e];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = sh[k][e];
      r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
    }
    const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long iu = i + u;
      long long dst;
      if (layout == 0) {
        const int bb = static_cast<int>(iu % d1);
        const int a = static_cast<int>((iu / d1) % d0);
        const int tap = static_cast<int>(iu / (static_cast<long long>(d1) * d0));
        dst = (static_cast<long long>(a) * d1 + bb) * taps + tap;
      } else {
        const int k = static_cast<int>(iu % (taps * d1));
        const int a = static_cast<int>(iu / (taps * d1));
        const int tap = k / d1, bb = k - tap * d1;
        dst = (static_cast<long long>(a) * d1 + bb) * taps + tap;
      }
      grad[dst] = rr[u];
    }
  }
}

// Batched split reduction: the per-layer kernels above run for ~15 us each, too short to reach the HBM roofline
// (~2 TB/s measured), and there are 44 of them per step. Here every layer of a backward segment is one job of ONE
// launch: job j owns thread blocks [start_j, start_{j+1}); a block is (256 / parts) lanes x parts, a lane owns four
// consecutive workspace elements (16-byte loads), part p sums splits p, p + parts, ... in two chains and the parts are
// summed in a fixed order in shared memory (deterministic). parts = min(8, largest power of two <= splits).
constexpr int kRedIter = 4;  // element groups per lane: 2 * kRedIter independent 16-byte loads in flight per thread
constexpr int kRedRowSplits = 16;  // jobs with fewer splits use the row-transposing path (coalesced gradient stores)

// row path tiling: a block owns `rows` d0 rows x `bch` d1 columns x all taps; the [rows][bch][taps] tile fits the
// kernel's 16 KB of shared memory. Whole rows when they fit (then several rows per block), else 256-column chunks.
__host__ __device__ inline int reduce_bchunk(int d1, int taps) {
  if (d1 * taps <= 4096) return d1;
  return (d1 % 256 == 0 && 256 * taps <= 4096) ? 256 : 0;
}
__host__ __device__ inline int reduce_rows_per_block(int bch, int taps) {
  const int r = 4096 / (bch * taps);
  return r < 1 ? 1 : (r > 8 ? 8 : r);
}

__global__ void __launch_bounds__(256) wgrad_reduce_batched_kernel(const ReduceJob* __restrict__ jobs, int njobs) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float4 sh[kRedIter][256];
  const long long b = blockIdx.x;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].start <= b) lo = mid;
    else hi = mid - 1;
  }
  const ReduceJob j = jobs[lo];
  if (j.parts == 0) {
    // Few splits: the strided 4-byte gradient stores of the path below would dominate (a 32-byte sector per element).
    // Here a block owns `rows` consecutive d0 rows with ALL taps: a thread sums one (row, tap, 4 x d1) unit over the
    // splits (four loads in flight), the block transposes through shared memory and writes the rows' contiguous
    // [d1][taps] gradient run with 16-byte stores.
    float* tile = reinterpret_cast<float*>(&sh[0][0]);          // [rows][bch][taps], <= 4096 floats
    const int bch = reduce_bchunk(j.d1, j.taps);                // d1 columns per block (whole rows when they fit)
    const int rows = reduce_rows_per_block(bch, j.taps);
    const int nchunks = j.d1 / bch;
    const int lb = static_cast<int>(b - j.start);
    const int a0 = (lb / nchunks) * rows;
    const int b0 = (lb % nchunks) * bch;
    const int q4 = bch >> 2;
    const int units = rows * j.taps * q4;
    const long long plane = static_cast<long long>(j.d0) * j.d1;
    for (int u = threadIdx.x; u < units; u += 256) {
      const int b4 = u % q4;
      const int rt = u / q4;
      const int tap = rt % j.taps, r = rt / j.taps;
      if (a0 + r >= j.d0) continue;
      const float* q = j.ws + tap * plane + static_cast<long long>(a0 + r) * j.d1 + b0 + b4 * 4;
      const int nsp = (j.splits2 > 0 && tap % 3 == 2) ? j.splits2 : j.splits;
      float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
      int sp = 0;
      for (; sp + 3 < nsp; sp += 4) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(q + static_cast<long long>(sp) * j.split_stride));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(q + static_cast<long long>(sp + 1) * j.split_stride));
        const float4 v2 = __ldg(reinterpret_cast<const float4*>(q + static_cast<long long>(sp + 2) * j.split_stride));
        const float4 v3 = __ldg(reinterpret_cast<const float4*>(q + static_cast<long long>(sp + 3) * j.split_stride));
        s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
        s1.x += v1.x; s1.y += v1.y; s1.z += v1.z; s1.w += v1.w;
        s2.x += v2.x; s2.y += v2.y; s2.z += v2.z; s2.w += v2.w;
        s3.x += v3.x; s3.y += v3.y; s3.z += v3.z; s3.w += v3.w;
      }
      for (; sp < nsp; ++sp) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(q + static_cast<long long>(sp) * j.split_stride));
        s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
      }
      float* t = tile + (static_cast<long long>(r) * bch + b4 * 4) * j.taps + tap;
      t[0] = (s0.x + s1.x) + (s2.x + s3.x);
      t[j.taps] = (s0.y + s1.y) + (s2.y + s3.y);
      t[2 * j.taps] = (s0.z + s1.z) + (s2.z + s3.z);
      t[3 * j.taps] = (s0.w + s1.w) + (s2.w + s3.w);
    }
    __syncthreads();
    const int nrows = min(rows, j.d0 - a0);                      // rows > 1 only when bch == d1: one contiguous run
    const int n4 = (nrows * bch * j.taps) >> 2;
    float4* dst = reinterpret_cast<float4*>(j.grad + (static_cast<long long>(a0) * j.d1 + b0) * j.taps);
    const float4* src = reinterpret_cast<const float4*>(tile);
    for (int k = threadIdx.x; k < n4; k += 256) dst[k] = src[k];
    return;
  }
  const int parts = j.parts;
  const int lanes = 256 / parts;
  const int e = threadIdx.x % lanes, part = threadIdx.x / lanes;
  const long long total = static_cast<long long>(j.d0) * j.d1 * j.taps;
  const long long i0 = ((b - j.start) * kRedIter * lanes + e) * 4;  // group it: i0 + it * lanes * 4
  float4 acc[kRedIter], c[kRedIter];
#pragma unroll
  for (int it = 0; it < kRedIter; ++it) acc[it] = c[it] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long plane_p = static_cast<long long>(j.d0) * j.d1;
  int nsp_it[kRedIter];  // splits of each group's tap (kx = 2 taps of a two-kx weight-gradient launch have fewer)
#pragma unroll
  for (int it = 0; it < kRedIter; ++it) {
    const long long i = i0 + static_cast<long long>(it) * lanes * 4;
    nsp_it[it] = (j.splits2 > 0 && i < total && (i / plane_p) % 3 == 2) ? j.splits2 : j.splits;
  }
  int sp = part;
  for (; sp + parts < j.splits; sp += 2 * parts) {
    float4 v0[kRedIter], v1[kRedIter];
#pragma unroll
    for (int it = 0; it < kRedIter; ++it) {
      const long long i = i0 + static_cast<long long>(it) * lanes * 4;
      v0[it] = v1[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < total && sp < nsp_it[it])
        v0[it] = __ldg(reinterpret_cast<const float4*>(j.ws + i + static_cast<long long>(sp) * j.split_stride));
      if (i < total && sp + parts < nsp_it[it])
        v1[it] = __ldg(reinterpret_cast<const float4*>(j.ws + i + static_cast<long long>(sp + parts) * j.split_stride));
    }
#pragma unroll
    for (int it = 0; it < kRedIter; ++it) {
      acc[it].x += v0[it].x; acc[it].y += v0[it].y; acc[it].z += v0[it].z; acc[it].w += v0[it].w;
      c[it].x += v1[it].x; c[it].y += v1[it].y; c[it].z += v1[it].z; c[it].w += v1[it].w;
    }
  }
  if (sp < j.splits) {
#pragma unroll
    for (int it = 0; it < kRedIter; ++it) {
      const long long i = i0 + static_cast<long long>(it) * lanes * 4;
      if (i < total && sp < nsp_it[it]) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(j.ws + i + static_cast<long long>(sp) * j.split_stride));
        acc[it].x += v.x; acc[it].y += v.y; acc[it].z += v.z; acc[it].w += v.w;
      }
    }
  }
#pragma unroll
  for (int it = 0; it < kRedIter; ++it) {
    acc[it].x += c[it].x; acc[it].y += c[it].y; acc[it].z += c[it].z; acc[it].w += c[it].w;
    sh[it][threadIdx.x] = acc[it];
  }
  __syncthreads();
  if (part == 0) {
#pragma unroll
    for (int it = 0; it < kRedIter; ++it) {
      const long long i = i0 + static_cast<long long>(it) * lanes * 4;
      if (i >= total) continue;
      float4 r = acc[it];
      for (int k = 1; k < parts; ++k) {
        const float4 v = sh[it][k * lanes + e];
        r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
      }
      const float rr[4] = {r.x, r.y, r.z, r.w};
      // layout 0 only: ws[tap][d0][d1] -> grad[d0][d1][tap]
      const int bb = static_cast<int>(i % j.d1);
      const int a = static_cast<int>((i / j.d1) % j.d0);
      const int tap = static_cast<int>(i / (static_cast<long long>(j.d1) * j.d0));
      float* dst = j.grad + (static_cast<long long>(a) * j.d1 + bb) * j.taps + tap;
#pragma unroll
      for (int u = 0; u < 4; ++u) dst[static_cast<long long>(u) * j.taps] = rr[u];  // d1 % 4 == 0: same (a, tap)
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Power-Jaccard loss (utils/loss_functions.py:141-150): p = sigmoid(z); I = sum p*t;
// D = sum p^2 + sum t^2 - I + 1e-6; L = 1 - I/D, over all selected batch rows at once.
// The target may itself be a logit (MMCR consistency term, train_semisupervised.py:75-105) and then
// receives a gradient too. sums = (I, sum p^2, sum t^2) in fp64 so that a data-parallel run can
// all-reduce them between the forward and backward kernels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + __expf(-z)); }

__global__ void __launch_bounds__(256) pj_reduce_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                        int t_is_logit, const unsigned char* __restrict__ rowmask,
                                                        int sel, int rows, long long per_row,
                                                        double* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sh[3][256];
  const long long total = static_cast<long long>(rows) * per_row;
  float a = 0.f, b = 0.f, c = 0.f;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
    const int row = static_cast<int>(i / per_row);
    if (rowmask != nullptr && (rowmask[row] != 0) != (sel != 0)) continue;
    const float4 zv = __ldg(reinterpret_cast<const float4*>(z + i));
    const float4 tv = __ldg(reinterpret_cast<const float4*>(t + i));
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
    const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float pp = sigmoidf_(zz[j]);
      const float tg = t_is_logit ? sigmoidf_(tt[j]) : tt[j];
      a = fmaf(pp, tg, a);
      b = fmaf(pp, pp, b);
      c = fmaf(tg, tg, c);
    }
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  sh[2][threadIdx.x] = c;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + st];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + st];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + st];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 3 + 0] = sh[0][0];
    partial[blockIdx.x * 3 + 1] = sh[1][0];
    partial[blockIdx.x * 3 + 2] = sh[2][0];
  }
}

__global__ void pj_finalize_kernel(const double* __restrict__ partial, int nblk, double* __restrict__ sums) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x < 3) {
    double a = 0.0;
    for (int i = 0; i < nblk; ++i) a += partial[i * 3 + threadIdx.x];
    sums[threadIdx.x] = a;
  }
}

__global__ void pj_loss_kernel(const double* __restrict__ sums, float* __restrict__ loss) {
  pdl_launch_dependents();
  pdl_wait();
  const double I = sums[0];
  const double D = sums[1] + sums[2] - I + 1e-6;
  *loss = static_cast<float>(1.0 - I / D);
}

__global__ void __launch_bounds__(256) pj_bwd_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                     int t_is_logit, const unsigned char* __restrict__ rowmask, int sel,
                                                     int rows, long long per_row, const double* __restrict__ sums,
                                                     const float* __restrict__ gptr, float gmul, int accumulate,
                                                     float* __restrict__ dz, float* __restrict__ dt) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(rows) * per_row;
  const float I = static_cast<float>(sums[0]);
  const float D = static_cast<float>(sums[1] + sums[2] - sums[0] + 1e-6);
  const float g = gmul * (gptr ? __ldg(gptr) : 1.f);
  const float invD2 = g / (D * D);
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
    const int row = static_cast<int>(i / per_row);
    const bool on = !(rowmask != nullptr && (rowmask[row] != 0) != (sel != 0));
    float4 oz = make_float4(0.f, 0.f, 0.f, 0.f), ot = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) {
      const float4 zv = __ldg(reinterpret_cast<const float4*>(z + i));
      const float4 tv = __ldg(reinterpret_cast<const float4*>(t + i));
      const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
      const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
      float rz[4], rt[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pp = sigmoidf_(zz[j]);
        const float tg = t_is_logit ? sigmoidf_(tt[j]) : tt[j];
        // dL/dp = -(t*D - I*(2p - t)) / D^2 ; dL/dt = -(p*D - I*(2t - p)) / D^2
        rz[j] = -(tg * D - I * (2.f * pp - tg)) * invD2 * pp * (1.f - pp);
        rt[j] = -(pp * D - I * (2.f * tg - pp)) * invD2 * (t_is_logit ? tg * (1.f - tg) : 1.f);
      }
      oz = make_float4(rz[0], rz[1], rz[2], rz[3]);
      ot = make_float4(rt[0], rt[1], rt[2], rt[3]);
    }
    if (accumulate) {
      if (on) {
        float4* pz = reinterpret_cast<float4*>(dz + i);
        float4 c = *pz;
        c.x += oz.x; c.y += oz.y; c.z += oz.z; c.w += oz.w;
        *pz = c;
        if (dt != nullptr) {
          float4* pt = reinterpret_cast<float4*>(dt + i);
          float4 d = *pt;
          d.x += ot.x; d.y += ot.y; d.z += ot.z; d.w += ot.w;
          *pt = d;
        }
      }
    } else {
      *reinterpret_cast<float4*>(dz + i) = oz;
      if (dt != nullptr) *reinterpret_cast<float4*>(dt + i) = ot;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW as constructed at train_supervised.py:32: decoupled weight decay, no amsgrad), every
// parameter tensor of the network in ONE launch. Job j owns thread blocks [start_j, start_{j+1}), 1024 elements per
// block; parameters whose gradient is None (outc_sem_change) simply have no job.
//   p <- p * (1 - lr * wd);  m <- m + (1 - b1)(g - m);  v <- b2 v + (1 - b2) g^2
//   p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The scalar factors are formed in double precision on the host, as torch forms them from Python floats.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(const AdamWJob* __restrict__ jobs, int njobs, float decay,
                                                    float step_size, float omb1, float b2, float omb2, float eps,
                                                    float sqrt_bc2) {
  pdl_launch_dependents();
  pdl_wait();
  const long long b = blockIdx.x;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].start <= b) lo = mid;
    else hi = mid - 1;
  }
  const AdamWJob j = jobs[lo];
  const long long i0 = (b - j.start) * 1024 + threadIdx.x * 4;
  if (i0 + 3 < j.n && j.vec4) {
    float4 pv = *reinterpret_cast<float4*>(j.p + i0);
    const float4 gv = __ldg(reinterpret_cast<const float4*>(j.g + i0));
    float4 mv = *reinterpret_cast<float4*>(j.m + i0);
    float4 vv = *reinterpret_cast<float4*>(j.v + i0);
    float* pp = &pv.x;
    const float* gg = &gv.x;
    float* mm = &mv.x;
    float* vq = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float g = gg[k];
      const float pd = pp[k] * decay;
      const float m = mm[k] + omb1 * (g - mm[k]);  // torch: exp_avg.lerp_(grad, 1 - beta1)
      const float v = b2 * vq[k] + omb2 * g * g;
      mm[k] = m;
      vq[k] = v;
      pp[k] = pd - step_size * (m / (sqrtf(v) / sqrt_bc2 + eps));
    }
    *reinterpret_cast<float4*>(j.p + i0) = pv;
    *reinterpret_cast<float4*>(j.m + i0) = mv;
    *reinterpret_cast<float4*>(j.v + i0) = vv;
  } else {
    for (long long i = i0; i < i0 + 4 && i < j.n; ++i) {
      const float g = j.g[i];
      const float pd = j.p[i] * decay;
      const float m = j.m[i] + omb1 * (g - j.m[i]);
      const float v = b2 * j.v[i] + omb2 * g * g;
      j.m[i] = m;
      j.v[i] = v;
      j.p[i] = pd - step_size * (m / (sqrtf(v) / sqrt_bc2 + eps));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Thresholded confusion counts of utils/metrics.py:23-31 (MultiThresholdMetric.add_sample) in one pass:
// pred = round(p - thr + 0.5) != 0 with p the probability (or sigmoid(logit) when from_logits), counted against
// y_true != 0 for up to 8 thresholds. counts[t] = (TP, TN, FP, FN) with the REFERENCE's naming: its "FP" is
// y_true & ~pred and its "FN" is ~y_true & pred (utils/metrics.py:29-30).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) confusion_kernel(const float* __restrict__ pred, const float* __restrict__ truth,
                                                        long long n, int from_logits, const float* __restrict__ thr,
                                                        int nthr, unsigned long long* __restrict__ counts) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ unsigned int sh[8][4];
  if (threadIdx.x < 32) sh[threadIdx.x >> 2][threadIdx.x & 3] = 0u;
  __syncthreads();
  float th[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) th[t] = t < nthr ? __ldg(thr + t) : 0.f;
  unsigned int c[8][4];
#pragma unroll
  for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = c[t][2] = c[t][3] = 0u;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float pv = __ldg(pred + i);
    if (from_logits) pv = 1.f / (1.f + expf(-pv));
    const bool yt = __ldg(truth + i) != 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (t < nthr) {
        const bool pd = rintf(pv - th[t] + 0.5f) != 0.f;
        c[t][0] += (yt && pd);
        c[t][1] += (!yt && !pd);
        c[t][2] += (yt && !pd);
        c[t][3] += (!yt && pd);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    if (t < nthr) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned int v = c[t][k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sh[t][k], v);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 4 * nthr) atomicAdd(&counts[threadIdx.x], static_cast<unsigned long long>(sh[threadIdx.x >> 2][threadIdx.x & 3]));
}

inline int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

cudaError_t launch_pack_input(const float* src0, const float* src1, int csrc, int c_lo, int nc, int cat_mode, int B,
                              int H, int W, int kpad, void* out, cudaStream_t st) {
  const int n_img = cat_mode ? B : 2 * B;
  const long long npix = static_cast<long long>(n_img) * H * W;
  const int grid = static_cast<int>((npix + 127) / 128);
  const int cin = cat_mode ? 2 * nc : nc;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (kpad == 64 * ((9 * cin + 63) / 64)) {
    const size_t sm = 128 * static_cast<size_t>(kpad / 8 + 1) * 16;
#define B200CD_PACK(CIN)                                                                                              \
  case CIN:                                                                                                          \
    launch_k(pack_input_t_kernel<CIN>, dim3(grid), dim3(128), sm, st, src0, src1, csrc, c_lo, nc, cat_mode, B, H, W, o, \
             npix);                                                                                                  \
    return cudaGetLastError();
    switch (cin) {
      B200CD_PACK(2)
      B200CD_PACK(3)
      B200CD_PACK(4)
      B200CD_PACK(6)
      B200CD_PACK(8)
      B200CD_PACK(12)
      default: break;
    }
#undef B200CD_PACK
  }
  const size_t smem = 128 * (kpad / 2 + 1) * 4;
  launch_k(pack_input_kernel, dim3(grid), dim3(128), smem, st, src0, src1, csrc, c_lo, nc, cat_mode, B, H, W, kpad, o, npix);
  return cudaGetLastError();
}

cudaError_t launch_pack_weights(int mode, const float* w, void* out, int d0, int d1, int kpad, cudaStream_t st) {
  long long total;
  if (mode == 0 || mode == 1) total = 9ll * d0 * d1;
  else if (mode == 2) total = static_cast<long long>(d0) * kpad;
  else total = 4ll * d0 * d1;
  launch_k(pack_weights_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, st, 
      mode, w, reinterpret_cast<__nv_bfloat16*>(out), d0, d1, kpad, total);
  return cudaGetLastError();
}

int pack_job_blocks(int mode, int d0, int d1, int kpad) {
  if (mode != 2 && d0 % kPackTile == 0 && d1 % kPackTile == 0) return (d0 / kPackTile) * (d1 / kPackTile);
  long long total;
  if (mode == 0 || mode == 1) total = 9ll * d0 * d1;
  else if (mode == 2) total = static_cast<long long>(d0) * kpad;
  else total = 4ll * d0 * d1;
  return static_cast<int>((total + kPackGenericPerBlock - 1) / kPackGenericPerBlock);
}

cudaError_t launch_pack_weights_batched(const PackJob* jobs, int njobs, long long total_blocks, cudaStream_t st) {
  launch_k(pack_weights_batched_kernel, dim3(static_cast<unsigned>(total_blocks)), dim3(256), 0, st, jobs, njobs);
  return cudaGetLastError();
}

cudaError_t launch_bn_stats_reduce(const float2* partial, int ld, int C, int tiles_per_group, int G, int spl,
                                   double* partial2, cudaStream_t st) {
  dim3 grid((C + 31) / 32, G, spl);
  launch_k(bn_stats_reduce_kernel, dim3(grid), dim3(256), 0, st, partial, ld, C, tiles_per_group, spl, partial2);
  return cudaGetLastError();
}

cudaError_t launch_bn_stats_fused(const float2* partial, int ld, int rows, int C, int G, double count, const float* gamma,
                                  const float* beta, float* running_mean, float* running_var, long long* nbt,
                                  float momentum, float eps, int order_rev, float* mean, float* invstd, float* scale,
                                  float* shift, cudaStream_t st) {
  launch_k(bn_stats_fused_kernel, dim3((C + 7) / 8), dim3(256), 0, st, partial, ld, rows, C, G, count, gamma, beta,
           running_mean, running_var, nbt, momentum, eps, order_rev, mean, invstd, scale, shift);
  return cudaGetLastError();
}

cudaError_t launch_bn_finalize(const double* partial2, int spl, int C, int G, double count, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, long long* nbt,
                               float momentum, float eps, int train, int order_rev, float* mean, float* invstd,
                               float* scale, float* shift, cudaStream_t st) {
  launch_k(bn_finalize_kernel, dim3((C + 7) / 8), dim3(256), 0, st, partial2, spl, C, G, count, gamma, beta, running_mean,
                                                      running_var, nbt, momentum, eps, train, order_rev, mean, invstd,
                                                      scale, shift);
  return cudaGetLastError();
}

cudaError_t launch_bn_eval_affine_batched(const BnEvalJob* jobs, int njobs, int total_blocks, cudaStream_t st) {
  launch_k(bn_eval_affine_batched_kernel, dim3(total_blocks), dim3(256), 0, st, jobs, njobs);
  return cudaGetLastError();
}

cudaError_t launch_bn_apply(const void* r, long long ld_r, const float* scale, const float* shift, int n_img, int H,
                            int W, int C, int G, int diff, void* a, long long ld_a, void* a2, long long ld_a2,
                            void* pool, long long ld_p, void* dif, long long ld_d, void* pool_idx, cudaStream_t st) {
  ApplyArgs p;
  p.r = reinterpret_cast<const __nv_bfloat16*>(r);
  p.ld_r = ld_r;
  p.scale = scale;
  p.shift = shift;
  p.n_img = n_img; p.H = H; p.W = W; p.C = C; p.G = G; p.diff = diff;
  p.a = reinterpret_cast<__nv_bfloat16*>(a);
  p.a2 = reinterpret_cast<__nv_bfloat16*>(a2);
  p.pool = reinterpret_cast<__nv_bfloat16*>(pool);
  p.dif = reinterpret_cast<__nv_bfloat16*>(dif);
  p.ld_a = ld_a; p.ld_a2 = ld_a2; p.ld_p = ld_p; p.ld_d = ld_d;
  p.pidx = reinterpret_cast<unsigned char*>(pool_idx);
  if (!diff && pool == nullptr && C % 64 == 0 && 256 % (C / 8) == 0) {
    const long long npx = static_cast<long long>(n_img / G) * H * W;
    long long nblk = npx / ((256 / (C / 8)) * 16);
    const long long cap = (148 * 4 * 4) / G;
    const long long want = (148 * 4) / G, fine = npx / ((256 / (C / 8)) * 4);  // small layers: >= 4 pixels per lane
    static const bool fine_on = [] { const char* e = getenv("B200CD_BN_FINE_APPLY"); return !(e && e[0] == '0'); }();
    if (fine_on && nblk < want) nblk = fine < want ? fine : want;
    nblk = nblk > cap ? cap : (nblk < 1 ? 1 : nblk);
    launch_k(bn_apply_px_kernel, dim3(dim3(static_cast<unsigned>(nblk), G)), dim3(256), 0, st, p);
    return cudaGetLastError();
  }
  if (H % 2 == 0 && W % 2 == 0 && C % 64 == 0 && 256 % (C / 8) == 0 &&
      static_cast<long long>(n_img) * H * W < (1ll << 31)) {
    const int lanes = 256 / (C / 8);
    const long long nwin = static_cast<long long>(diff ? n_img / 2 : n_img / G) * (H / 2) * (W / 2);
    long long nblk = nwin / (static_cast<long long>(lanes) * 4);
    const long long cap = (148 * 3 * 4) / (diff ? 1 : G);
    nblk = nblk > cap ? cap : (nblk < 1 ? 1 : nblk);
    if (diff) launch_k(bn_apply_win_kernel<true>, dim3(dim3(static_cast<unsigned>(nblk), 1)), dim3(256), 0, st, p);
    else launch_k(bn_apply_win_kernel<false>, dim3(dim3(static_cast<unsigned>(nblk), G)), dim3(256), 0, st, p);
    return cudaGetLastError();
  }
  const long long total = static_cast<long long>(diff ? n_img / 2 : n_img) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  launch_k(bn_apply_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, p);
  return cudaGetLastError();
}

static BwdArgs make_bwd_args(const void* r, long long ld_r, const float* scale, const float* shift, const GradSrcs& srcs,
                             int n_img, int H, int W, int C, int G) {
  BwdArgs p;
  p.r = reinterpret_cast<const __nv_bfloat16*>(r);
  p.ld_r = ld_r;
  p.scale = scale; p.shift = shift;
  p.srcs = srcs;
  p.n_img = n_img; p.H = H; p.W = W; p.C = C; p.G = G;
  return p;
}

// first source a max-pool, the others direct (or absent), even H and W: the 2x2-window kernels apply
static bool bn_bwd_use_windows(const GradSrcs& srcs, int H, int W) {
  const int k1 = srcs.s[1].kind, k2 = srcs.s[2].kind;
  return srcs.s[0].kind == 2 && H % 2 == 0 && W % 2 == 0 && (k1 == 0 || k1 == 1) && (k2 == 0 || (k2 == 1 && k1 == 1));
}

// source-kind triples with their own instantiation (everything the six network types produce); others run generic
#define B200CD_BWD_DISPATCH(CALL)                                  \
  do {                                                             \
    const int k0 = srcs.s[0].kind, k1 = srcs.s[1].kind, k2 = srcs.s[2].kind; \
    if (k0 == 1 && k1 == 0 && k2 == 0) { CALL(1, 0, 0); }          \
    else if (k0 == 1 && k1 == 1 && k2 == 0) { CALL(1, 1, 0); }     \
    else if (k0 == 2 && k1 == 1 && k2 == 0) { CALL(2, 1, 0); }     \
    else if (k0 == 2 && k1 == 1 && k2 == 1) { CALL(2, 1, 1); }     \
    else if (k0 == 1 && k1 == 3 && k2 == 0) { CALL(1, 3, 0); }     \
    else if (k0 == 3 && k1 == 0 && k2 == 0) { CALL(3, 0, 0); }     \
    else if (k0 == 3 && k1 == 3 && k2 == 0) { CALL(3, 3, 0); }     \
    else { CALL(-1, -1, -1); }                                     \
  } while (0)

cudaError_t launch_bn_bwd_reduce(const void* r, long long ld_r, const float* scale, const float* shift,
                                 const GradSrcs& srcs, int n_img, int H, int W, int C, int G, int nblk, float* partial,
                                 cudaStream_t st) {
  const BwdArgs p = make_bwd_args(r, ld_r, scale, shift, srcs, n_img, H, W, C, G);
  const int lanes = 256 / (C / 8);
  const size_t smem = static_cast<size_t>(lanes) * C * 2 * sizeof(float);
  dim3 grid(nblk, G);
  if (bn_bwd_use_windows(srcs, H, W)) {
    const int k1 = srcs.s[1].kind, k2 = srcs.s[2].kind;
    if (k1 == 1 && k2 == 1) launch_k(bn_bwd_reduce_win_kernel<1, 1>, dim3(grid), dim3(256), smem, st, p, partial);
    else if (k1 == 1) launch_k(bn_bwd_reduce_win_kernel<1, 0>, dim3(grid), dim3(256), smem, st, p, partial);
    else launch_k(bn_bwd_reduce_win_kernel<0, 0>, dim3(grid), dim3(256), smem, st, p, partial);
    return cudaGetLastError();
  }
#define B200CD_CALL(A, B, Cc) launch_k(bn_bwd_reduce_kernel<A, B, Cc>, dim3(grid), dim3(256), smem, st, p, partial)
  B200CD_BWD_DISPATCH(B200CD_CALL);
#undef B200CD_CALL
  return cudaGetLastError();
}

cudaError_t launch_bn_bwd_finalize(const float* partial, int nblk, int C, int G, double count, const float* mean,
                                   const float* invstd, const float* scale, float* dgamma, float* dbeta, float* coefA,
                                   float* coefB, cudaStream_t st) {
  launch_k(bn_bwd_finalize_kernel, dim3((C + 31) / 32), dim3(1024), 0, st, partial, nblk, C, G, count, mean, invstd, scale, dgamma, dbeta,
                                                         coefA, coefB);
  return cudaGetLastError();
}

cudaError_t launch_bn_bwd_dx(const void* r, long long ld_r, const float* scale, const float* shift, const float* coefA,
                             const float* coefB, const GradSrcs& srcs, int n_img, int H, int W, int C, int G, int nblk,
                             void* dr, long long ld_dr, cudaStream_t st) {
  const BwdArgs p = make_bwd_args(r, ld_r, scale, shift, srcs, n_img, H, W, C, G);
  dim3 grid(nblk, G);
  if (bn_bwd_use_windows(srcs, H, W)) {
    const int k1 = srcs.s[1].kind, k2 = srcs.s[2].kind;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dr);
    if (k1 == 1 && k2 == 1) launch_k(bn_bwd_dx_win_kernel<1, 1>, dim3(grid), dim3(256), 0, st, p, coefA, coefB, o, ld_dr);
    else if (k1 == 1) launch_k(bn_bwd_dx_win_kernel<1, 0>, dim3(grid), dim3(256), 0, st, p, coefA, coefB, o, ld_dr);
    else launch_k(bn_bwd_dx_win_kernel<0, 0>, dim3(grid), dim3(256), 0, st, p, coefA, coefB, o, ld_dr);
    return cudaGetLastError();
  }
#define B200CD_CALL(A, B, Cc) \
  launch_k(bn_bwd_dx_kernel<A, B, Cc>, dim3(grid), dim3(256), 0, st, p, coefA, coefB, reinterpret_cast<__nv_bfloat16*>(dr), ld_dr)
  B200CD_BWD_DISPATCH(B200CD_CALL);
#undef B200CD_CALL
  return cudaGetLastError();
}

cudaError_t launch_head_fwd(const void* a0, long long ld0, const void* a1, long long ld1, int C, const float* w,
                            const float* b, long long npix, float* logits, cudaStream_t st) {
  const long long threads = npix * 8;
  launch_k(head_fwd_kernel, dim3(static_cast<int>((threads + 255) / 256)), dim3(256), 0, st, 
      reinterpret_cast<const __nv_bfloat16*>(a0), ld0, reinterpret_cast<const __nv_bfloat16*>(a1), ld1, C, w, b, npix,
      logits);
  return cudaGetLastError();
}

cudaError_t launch_pad_copy(const void* src, long long ld_s, int n_img, int h, int w, int C, void* dst, long long ld_d,
                            int H, int W, int top, int left, cudaStream_t st) {
  const long long total = static_cast<long long>(n_img) * H * W * (C / 8);
  launch_k(pad_copy_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(src), ld_s,
           n_img, h, w, C, reinterpret_cast<__nv_bfloat16*>(dst), ld_d, H, W, top, left);
  return cudaGetLastError();
}

cudaError_t launch_colsum(const void* x, long long ld, int C, const float* wgt, long long npix, int nblk,
                          float* partial, cudaStream_t st) {
  size_t smem = 256 * sizeof(float);
  if (x != nullptr) smem = static_cast<size_t>(256 / (C / 8)) * C * sizeof(float);
  launch_k(colsum_kernel, dim3(nblk), dim3(256), smem, st, reinterpret_cast<const __nv_bfloat16*>(x), ld, C, wgt, npix, partial);
  return cudaGetLastError();
}

cudaError_t launch_stat_rowsum(const float2* stats, int rows, int ld, int c_off, int C, float* out, cudaStream_t st) {
  launch_k(stat_rowsum_kernel, dim3((C + 7) / 8), dim3(256), 0, st, stats, rows, ld, c_off, C, out);
  return cudaGetLastError();
}

cudaError_t launch_colsum_finalize(const float* partial, int nblk, int C, float* out, cudaStream_t st) {
  launch_k(colsum_finalize_kernel, dim3((C + 31) / 32), dim3(1024), 0, st, partial, nblk, C, out);
  return cudaGetLastError();
}

cudaError_t launch_wgrad_reduce(const float* ws, int splits, long long split_stride, int layout, int d0, int d1,
                                int taps, float* grad, cudaStream_t st) {
  const long long total = static_cast<long long>(d0) * d1 * taps;
  const bool v4 = total % 4 == 0 && split_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(ws) & 15u) == 0 &&
                  (layout == 0 || ((static_cast<long long>(taps) * d1) % 4 == 0 && (split_stride / d0) % 4 == 0));
  if (v4 && splits >= 16)
    launch_k(wgrad_reduce_v4_kernel, dim3(static_cast<unsigned>((total / 4 + 31) / 32)), dim3(256), 0, st, ws, splits,
             split_stride, layout, d0, d1, taps, grad, total);
  else
    launch_k(wgrad_reduce_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, st, ws, splits, split_stride,
             layout, d0, d1, taps, grad, total);
  return cudaGetLastError();
}

cudaError_t launch_pj_reduce(const float* z, const float* t, int t_is_logit, const unsigned char* rowmask, int sel,
                             int rows, long long per_row, int nblk, double* partial, cudaStream_t st) {
  launch_k(pj_reduce_kernel, dim3(nblk), dim3(256), 0, st, z, t, t_is_logit, rowmask, sel, rows, per_row, partial);
  return cudaGetLastError();
}

cudaError_t launch_pj_finalize(const double* partial, int nblk, double* sums, cudaStream_t st) {
  launch_k(pj_finalize_kernel, dim3(1), dim3(32), 0, st, partial, nblk, sums);
  return cudaGetLastError();
}

cudaError_t launch_pj_loss(const double* sums, float* loss, cudaStream_t st) {
  launch_k(pj_loss_kernel, dim3(1), dim3(1), 0, st, sums, loss);
  return cudaGetLastError();
}

cudaError_t launch_pj_bwd(const float* z, const float* t, int t_is_logit, const unsigned char* rowmask, int sel,
                          int rows, long long per_row, const double* sums, const float* gptr, float gmul,
                          int accumulate, float* dz, float* dt, cudaStream_t st) {
  const long long total = static_cast<long long>(rows) * per_row;
  launch_k(pj_bwd_kernel, dim3(grid_for(total / 4, 256)), dim3(256), 0, st, z, t, t_is_logit, rowmask, sel, rows, per_row, sums, gptr,
                                                          gmul, accumulate, dz, dt);
  return cudaGetLastError();
}

// parts == 0 selects the row path: few splits, d1 % 4 == 0 and a row (or a 256-column chunk of it) fits the tile
int reduce_job_parts(int splits, int d1, int taps) {
  if (splits < kRedRowSplits && d1 % 4 == 0 && reduce_bchunk(d1, taps) > 0) return 0;
  int p = 1;
  while (p < 8 && 2 * p <= splits) p *= 2;
  return p;
}

long long reduce_job_blocks(int splits, int d0, int d1, int taps) {
  const int parts = reduce_job_parts(splits, d1, taps);
  if (parts == 0) {
    const int bch = reduce_bchunk(d1, taps);
    const int rows = reduce_rows_per_block(bch, taps);
    return static_cast<long long>((d0 + rows - 1) / rows) * (d1 / bch);
  }
  const long long per_block = 4ll * kRedIter * (256 / parts);
  const long long total = static_cast<long long>(d0) * d1 * taps;
  return (total + per_block - 1) / per_block;
}

cudaError_t launch_wgrad_reduce_batched(const ReduceJob* jobs, int njobs, long long total_blocks, cudaStream_t st) {
  launch_k(wgrad_reduce_batched_kernel, dim3(static_cast<unsigned>(total_blocks)), dim3(256), 0, st, jobs, njobs);
  return cudaGetLastError();
}

cudaError_t launch_confusion(const float* pred, const float* truth, long long n, int from_logits, const float* thr, int nthr,
                             unsigned long long* counts, cudaStream_t st) {
  launch_k(confusion_kernel, dim3(grid_for(n, 256, 148 * 8)), dim3(256), 0, st, pred, truth, n, from_logits, thr, nthr, counts);
  return cudaGetLastError();
}

cudaError_t launch_adamw(const AdamWJob* jobs, int njobs, long long total_blocks, float decay, float step_size, float omb1,
                         float b2, float omb2, float eps, float sqrt_bc2, cudaStream_t st) {
  launch_k(adamw_kernel, dim3(static_cast<unsigned>(total_blocks)), dim3(256), 0, st, jobs, njobs, decay, step_size, omb1,
           b2, omb2, eps, sqrt_bc2);
  return cudaGetLastError();
}

}  // namespace b200cd

// ------------------------------------------------------------------------------------------------
// N2 — training-time augmentation + packing on the GPU (utils/augmentations.py:6-142 composed as in
// utils/datasets.py:111-181): per sample crop -> horizontal / vertical flip -> rot90(k) -> colour shift -> gamma
// -> HWC -> CHW, with the channel regrouping of datasets.py:156-162 (x_t1 = [S1 t1 | S2 t1], x_t2 = [S1 t2 | S2 t2]).
// The random DECISIONS stay on the host (numpy RNG, same call order as the reference so a seeded run draws the same
// augmentations; the importance crop needs only the label sums); the pixel work — 12 fp32 channels of a full tile per
// sample — happens here, all samples of a batch in one launch. One thread = one output pixel, all output channels.
// ------------------------------------------------------------------------------------------------
namespace b200cd {
namespace {

__global__ void __launch_bounds__(256) augment_kernel(const AugmentJob* __restrict__ jobs, int n, int cs, int cout,
                                                      float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(n) * cs * cs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % cs);
    const int y = static_cast<int>((i / cs) % cs);
    const int s = static_cast<int>(i / (static_cast<long long>(cs) * cs));
    const AugmentJob& j = jobs[s];
    // invert rot90(k) (counter-clockwise, square crop): R[i][j] = M[j][cs-1-i] (k=1), M[cs-1-i][cs-1-j] (2), M[cs-1-j][i] (3)
    int yy = y, xx = x;
    if (j.rotk == 1) { yy = x; xx = cs - 1 - y; }
    else if (j.rotk == 2) { yy = cs - 1 - y; xx = cs - 1 - x; }
    else if (j.rotk == 3) { yy = cs - 1 - x; xx = y; }
    if (j.vflip) yy = cs - 1 - yy;       // np.flip(axis=0) applied after the horizontal flip
    if (j.hflip) xx = cs - 1 - xx;       // np.flip(axis=1)
    const float* px = j.src + (static_cast<long long>(j.y0 + yy) * j.W0 + (j.x0 + xx)) * j.C;
    for (int c = 0; c < cout; ++c) {
      const int sc = j.cmap[c];
      float v = __ldg(px + sc);
      if (j.use_mul) v = fminf(fmaxf(v * j.mul[sc], 0.f), 1.f);            // ColorShift: clip(img * factor, 0, 1)
      if (j.use_gamma) v = fminf(fmaxf(powf(v, j.gamma[sc]), 0.f), 1.f);   // GammaCorrection: clip(img ** gamma, 0, 1)
      out[((static_cast<long long>(s) * cout + c) * cs + y) * cs + x] = v;
    }
  }
}

}  // namespace

cudaError_t launch_augment(const AugmentJob* jobs, int n, int cs, int cout, float* out, cudaStream_t st) {
  const long long total = static_cast<long long>(n) * cs * cs;
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  launch_k(augment_kernel, dim3(static_cast<unsigned>(g)), dim3(256), 0, st, jobs, n, cs, cout, out);
  return cudaGetLastError();
}

}  // namespace b200cd
