// Memory-bound kernels of the "precise" (split-bf16) mode, sm_100a.
//
// In this mode every activation / gradient tensor keeps 16 mantissa bits as TWO bf16 values per element: row layout
// [hi (C channels) | lo (C channels)] with hi = bf16(x), lo = bf16(x - hi) and the lo half exactly ld / 2 elements
// behind the hi half (ld = elements per pixel of the buffer the view belongs to). The tensor-core kernels consume the
// halves directly (three bf16 MMAs per product: hi*hi + hi*lo + lo*hi, fp32 accumulation); the kernels here read
// x = hi + lo, compute in fp32 exactly as their bf16-storage counterparts in elementwise.cu do, and write both halves.
// Same reference ops as elementwise.cu (utils/networks.py:393-397, 420, 147-150, 449, 454-461); same two-stage
// fixed-order reductions. They are deliberately the plain variants (no per-shape instantiations): the mode exists to
// meet the fp32 tolerance, the bf16-storage mode is the fast one.
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
// the value a split-bf16 store of x reads back as
__device__ __forceinline__ float q16(float x) {
  const float h = round_bf16(x);
  return h + round_bf16(x - h);
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// 8 channels of a split tensor: p points at the hi half, the lo half lies `lo` elements further
__device__ __forceinline__ void load8hl(const __nv_bfloat16* p, long long lo, float (&f)[8]) {
  const uint4 a = ldg16(p), b = ldg16(p + lo);
  float h[8], l[8];
  unpack8(a, h);
  unpack8(b, l);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = h[j] + l[j];
}
__device__ __forceinline__ void store8hl(__nv_bfloat16* p, long long lo, const float (&f)[8]) {
  const uint4 hi = pack8(f);
  float h[8], r[8];
  unpack8(hi, h);
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = f[j] - h[j];
  *reinterpret_cast<uint4*>(p) = hi;
  *reinterpret_cast<uint4*>(p + lo) = pack8(r);
}

// ------------------------------------------------------------------------------------------------
// Input packing: fp32 NCHW inputs -> split im2col rows [pixel][hi kpad | lo kpad], k = tap*Cin + ci (see
// pack_input_kernel in elementwise.cu for cat_mode). One thread = one pixel x two consecutive k.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_input_hp_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                                            int csrc, int c_lo, int nc, int cat_mode, int B, int H, int W,
                                                            int kpad, __nv_bfloat16* __restrict__ out, long long npix) {
  pdl_launch_dependents();
  pdl_wait();
  const int kp2 = kpad >> 1;
  const long long total = npix * kp2;
  const int cin = cat_mode ? 2 * nc : nc;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k0 = static_cast<int>(i % kp2) * 2;
    const long long pix = i / kp2;
    const int x = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float v[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = k0 + u;
      if (k < 9 * cin) {
        const int tap = k / cin, ci = k - tap * cin;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
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
          v[u] = __ldg(src + ((static_cast<long long>(b) * csrc + c) * H + yy) * W + xx);
        }
      }
    }
    const uint32_t hi = pack_bf16x2(v[0], v[1]);
    const uint32_t lo = pack_bf16x2(v[0] - bf16_lo(hi), v[1] - bf16_hi(hi));
    uint32_t* row = reinterpret_cast<uint32_t*>(out + pix * 2 * kpad);
    row[k0 >> 1] = hi;
    row[(kpad + k0) >> 1] = lo;
  }
}

// ------------------------------------------------------------------------------------------------
// Weight packing fp32 -> split-bf16 GEMM operand: the layouts of pack_weights_kernel (elementwise.cu, modes 0..4)
// with K tripled per tap: out[row][tap][ hi(K) | lo(K) | hi(K) ]. The convolution kernels pair these thirds with the
// hi, hi and lo halves of the activations: A_hi*W_hi + A_hi*W_lo + A_lo*W_hi.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pack_weight_value_hp(int mode, const float* __restrict__ w, int d0, int d1, int kpad,
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

constexpr int kPackHpPerBlock = 2048;

__device__ __forceinline__ long long pack_total(int mode, int d0, int d1, int kpad) {
  if (mode == 0 || mode == 1) return 9ll * d0 * d1;
  if (mode == 2) return static_cast<long long>(d0) * kpad;
  return 4ll * d0 * d1;
}
// K per tap of the packed layout (the unit that is tripled)
__device__ __forceinline__ int pack_k(int mode, int d0, int d1, int kpad) {
  return mode == 0 ? d1 : (mode == 1 ? d0 : (mode == 2 ? kpad : (mode == 3 ? d0 : d1)));
}

__device__ __forceinline__ void pack_hp_one(int mode, const float* w, __nv_bfloat16* out, int d0, int d1, int kpad,
                                            long long i) {
  const float v = pack_weight_value_hp(mode, w, d0, d1, kpad, i);
  const int K = pack_k(mode, d0, d1, kpad);
  const long long rt = i / K;            // (row, tap) index
  const int k = static_cast<int>(i - rt * K);
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16* o = out + rt * 3 * K + k;
  o[0] = hi;
  o[K] = lo;
  o[2 * K] = hi;
}

__global__ void __launch_bounds__(256) pack_weights_hp_batched_kernel(const PackJob* __restrict__ jobs, int njobs) {
  pdl_launch_dependents();
  pdl_wait();
  const long long b = blockIdx.x;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].start <= b) lo = mid;
    else hi = mid - 1;
  }
  const PackJob j = jobs[lo];
  const long long total = pack_total(j.mode, j.d0, j.d1, j.kpad);
  const long long e0 = (b - j.start) * kPackHpPerBlock;
  for (int k = threadIdx.x; k < kPackHpPerBlock; k += 256) {
    const long long li = e0 + k;
    if (li >= total) break;
    pack_hp_one(j.mode, j.w, reinterpret_cast<__nv_bfloat16*>(j.out), j.d0, j.d1, j.kpad, li);
    if (j.out2 != nullptr) pack_hp_one(j.mode2, j.w, reinterpret_cast<__nv_bfloat16*>(j.out2), j.d0, j.d1, j.kpad, li);
  }
}

// ------------------------------------------------------------------------------------------------
// BN-apply + ReLU (+ MaxPool2d + arg-max index, + t2 - t1, + second copy): bn_apply_kernel of elementwise.cu on split
// tensors. One thread = one 2x2 pixel window x 8 channels (x both timestamps when diff). Any H, W.
// ------------------------------------------------------------------------------------------------
struct ApplyArgsHp {
  const __nv_bfloat16* r;
  long long ld_r;
  const float* scale;
  const float* shift;
  int n_img, H, W, C, G, diff;
  __nv_bfloat16 *a, *a2, *pool, *dif;
  long long ld_a, ld_a2, ld_p, ld_d;
  unsigned char* pidx;
};

__global__ void __launch_bounds__(256) bn_apply_hp_kernel(const ApplyArgsHp p) {
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
    float av[2][4][8];
    const int reps = p.diff ? 2 : 1;
    const bool full = (2 * y2 + 1 < p.H) && (2 * x2 + 1 < p.W);
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      if (rep >= reps) break;
      const int nn = n + rep * n_units;
      const int g = nn / per_group;
      float sc[8], sh[8], amax[8];
      load8f(p.scale + g * p.C + c, sc);
      load8f(p.shift + g * p.C + c, sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) amax[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int y = 2 * y2 + (k >> 1), x = 2 * x2 + (k & 1);
        if (y < p.H && x < p.W) {
          const long long pix = (static_cast<long long>(nn) * p.H + y) * p.W + x;
          float rv[8];
          load8hl(p.r + pix * p.ld_r + c, p.ld_r >> 1, rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v = fmaxf(fmaf(rv[j], sc[j], sh[j]), 0.f);
            av[rep][k][j] = v;
            amax[j] = fmaxf(amax[j], v);
          }
          if (p.a) store8hl(p.a + pix * p.ld_a + c, p.ld_a >> 1, av[rep][k]);
          if (p.a2) store8hl(p.a2 + pix * p.ld_a2 + c, p.ld_a2 >> 1, av[rep][k]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) av[rep][k][j] = 0.f;
        }
      }
      if (p.pool && full) {
        const long long ppix = (static_cast<long long>(nn) * (p.H >> 1) + y2) * (p.W >> 1) + x2;
        store8hl(p.pool + ppix * p.ld_p + c, p.ld_p >> 1, amax);
        if (p.pidx) {
          // first maximum in row-major order of the STORED activations, as ATen's max_pool2d_with_indices
          unsigned long long packed = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            int arg = 0;
            float best = q16(av[rep][0][j]);
#pragma unroll
            for (int k = 1; k < 4; ++k) {
              const float v = q16(av[rep][k][j]);
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
          store8hl(p.dif + pix * p.ld_d + c, p.ld_d >> 1, d);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BN + ReLU backward on split tensors: the two passes of elementwise.cu (reduce S1 = sum dy, S2 = sum dy*r per block;
// dr = scale*dy + r*A + B) with the gradient sources gathered on the fly; kinds read at run time.
// The fp64 finalize between the passes is bn_bwd_finalize_kernel of elementwise.cu (it only sees fp32 partials).
// ------------------------------------------------------------------------------------------------
struct BwdArgsHp {
  const __nv_bfloat16* r;
  long long ld_r;
  const float *scale, *shift;
  GradSrcs srcs;
  int n_img, H, W, C, G;
};

__device__ __forceinline__ void gather_px_hp(const BwdArgsHp& p, int gpix, int hw, int c, float (&dy)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) dy[j] = 0.f;
#pragma unroll
  for (int si = 0; si < 3; ++si) {
    const GradSrc& s = p.srcs.s[si];
    const int kind = s.kind;
    if (kind == 0) continue;
    int sp = gpix;
    float scale = 1.f;
    if (s.n_mod > 0) {
      const int n = gpix / hw;
      scale = n < s.n_mod ? s.scale_lo : s.scale_hi;
      sp = gpix - (n - n % s.n_mod) * hw;
    }
    if (kind == 1) {
      float gv[8];
      load8hl(reinterpret_cast<const __nv_bfloat16*>(s.ptr) + static_cast<long long>(sp) * s.ld + c, s.ld >> 1, gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[j] = fmaf(scale, gv[j], dy[j]);
    } else if (kind == 2) {
      const int n = sp / hw, pix = sp - n * hw;
      const int y = pix / p.W, x = pix - y * p.W;
      const int H2 = p.H >> 1, W2 = p.W >> 1;
      if ((y >> 1) < H2 && (x >> 1) < W2) {
        const long long pp = (static_cast<long long>(n) * H2 + (y >> 1)) * W2 + (x >> 1);
        const unsigned long long idx =
            __ldg(reinterpret_cast<const unsigned long long*>(reinterpret_cast<const unsigned char*>(s.w) + pp * p.C + c));
        float gv[8];
        load8hl(reinterpret_cast<const __nv_bfloat16*>(s.ptr) + pp * s.ld + c, s.ld >> 1, gv);
        const unsigned me = ((y & 1) << 1) | (x & 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[j] += (((idx >> (8 * j)) & 0xffull) == me) ? scale * gv[j] : 0.f;
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

// grid = (nblk, G); block = 256 threads = (C/8 channel vectors) x (256/(C/8)) pixel lanes
__global__ void __launch_bounds__(256) bn_bwd_reduce_hp_kernel(const BwdArgsHp p, float* __restrict__ partial) {
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
  for (int i = pb + l; i < pe; i += lanes) {
    float d[8], v[8];
    load8hl(rbase + static_cast<long long>(i) * p.ld_r, p.ld_r >> 1, v);
    gather_px_hp(p, gbase + i, hw, c, d);
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

__global__ void __launch_bounds__(256) bn_bwd_dx_hp_kernel(const BwdArgsHp p, const float* __restrict__ coefA,
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
  for (int i = pb + l; i < pe; i += lanes) {
    float d[8], v[8], o[8];
    load8hl(rbase + static_cast<long long>(i) * p.ld_r, p.ld_r >> 1, v);
    gather_px_hp(p, gbase + i, hw, c, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = fmaf(v[j], sc[j], sh[j]) > 0.f ? d[j] : 0.f;
      o[j] = fmaf(sc[j], m, fmaf(v[j], ca[j], cb[j]));
    }
    store8hl(obase + static_cast<long long>(i) * ld_dr, ld_dr >> 1, o);
  }
}

// ------------------------------------------------------------------------------------------------
// 1x1 head (OutConv, C -> 1) over one or two split inputs; 8 lanes per pixel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fwd_hp_kernel(const __nv_bfloat16* __restrict__ a0, long long ld0,
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
      load8hl(a0 + pix * ld0 + c, ld0 >> 1, av);
      load8f(w + c, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(av[j], wv[j], acc);
    }
    if (a1 != nullptr) {
      for (int c = sub * 8; c < C; c += 64) {
        float av[8], wv[8];
        load8hl(a1 + pix * ld1 + c, ld1 >> 1, av);
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

// Weighted column sums out[c] = sum_pixels wgt[pixel] * x[pixel, c] of a split tensor (first stage; the fixed-order
// finalize is colsum_finalize_kernel of elementwise.cu).
__global__ void __launch_bounds__(256) colsum_hp_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int C,
                                                        const float* __restrict__ wgt, long long npix,
                                                        float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float shred[];
  const long long pb = npix * blockIdx.x / gridDim.x, pe = npix * (blockIdx.x + 1) / gridDim.x;
  const int cvecs = C >> 3;
  const int lanes = 256 / cvecs;
  const int cv = threadIdx.x % cvecs, l = threadIdx.x / cvecs;
  const int c = cv << 3;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (l < lanes) {
    for (long long i = pb + l; i < pe; i += lanes) {
      float v[8];
      load8hl(x + i * ld + c, ld >> 1, v);
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

inline int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

cudaError_t launch_pack_input_hp(const float* src0, const float* src1, int csrc, int c_lo, int nc, int cat_mode, int B,
                                 int H, int W, int kpad, void* out, cudaStream_t st) {
  const int n_img = cat_mode ? B : 2 * B;
  const long long npix = static_cast<long long>(n_img) * H * W;
  launch_k(pack_input_hp_kernel, dim3(grid_for(npix * (kpad / 2), 256, 148 * 32)), dim3(256), 0, st, src0, src1, csrc, c_lo,
           nc, cat_mode, B, H, W, kpad, reinterpret_cast<__nv_bfloat16*>(out), npix);
  return cudaGetLastError();
}

int pack_job_blocks_hp(int mode, int d0, int d1, int kpad) {
  long long total;
  if (mode == 0 || mode == 1) total = 9ll * d0 * d1;
  else if (mode == 2) total = static_cast<long long>(d0) * kpad;
  else total = 4ll * d0 * d1;
  return static_cast<int>((total + kPackHpPerBlock - 1) / kPackHpPerBlock);
}

cudaError_t launch_pack_weights_hp_batched(const PackJob* jobs, int njobs, long long total_blocks, cudaStream_t st) {
  launch_k(pack_weights_hp_batched_kernel, dim3(static_cast<unsigned>(total_blocks)), dim3(256), 0, st, jobs, njobs);
  return cudaGetLastError();
}

cudaError_t launch_bn_apply_hp(const void* r, long long ld_r, const float* scale, const float* shift, int n_img, int H,
                               int W, int C, int G, int diff, void* a, long long ld_a, void* a2, long long ld_a2,
                               void* pool, long long ld_p, void* dif, long long ld_d, void* pool_idx, cudaStream_t st) {
  ApplyArgsHp p;
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
  const long long total = static_cast<long long>(diff ? n_img / 2 : n_img) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  launch_k(bn_apply_hp_kernel, dim3(grid_for(total, 256, 148 * 32)), dim3(256), 0, st, p);
  return cudaGetLastError();
}

static BwdArgsHp make_bwd_args_hp(const void* r, long long ld_r, const float* scale, const float* shift,
                                  const GradSrcs& srcs, int n_img, int H, int W, int C, int G) {
  BwdArgsHp p;
  p.r = reinterpret_cast<const __nv_bfloat16*>(r);
  p.ld_r = ld_r;
  p.scale = scale; p.shift = shift;
  p.srcs = srcs;
  p.n_img = n_img; p.H = H; p.W = W; p.C = C; p.G = G;
  return p;
}

cudaError_t launch_bn_bwd_reduce_hp(const void* r, long long ld_r, const float* scale, const float* shift,
                                    const GradSrcs& srcs, int n_img, int H, int W, int C, int G, int nblk,
                                    float* partial, cudaStream_t st) {
  const BwdArgsHp p = make_bwd_args_hp(r, ld_r, scale, shift, srcs, n_img, H, W, C, G);
  const size_t smem = static_cast<size_t>(256 / (C / 8)) * C * 2 * sizeof(float);
  launch_k(bn_bwd_reduce_hp_kernel, dim3(nblk, G), dim3(256), smem, st, p, partial);
  return cudaGetLastError();
}

cudaError_t launch_bn_bwd_dx_hp(const void* r, long long ld_r, const float* scale, const float* shift,
                                const float* coefA, const float* coefB, const GradSrcs& srcs, int n_img, int H, int W,
                                int C, int G, int nblk, void* dr, long long ld_dr, cudaStream_t st) {
  const BwdArgsHp p = make_bwd_args_hp(r, ld_r, scale, shift, srcs, n_img, H, W, C, G);
  launch_k(bn_bwd_dx_hp_kernel, dim3(nblk, G), dim3(256), 0, st, p, coefA, coefB, reinterpret_cast<__nv_bfloat16*>(dr),
           ld_dr);
  return cudaGetLastError();
}

cudaError_t launch_head_fwd_hp(const void* a0, long long ld0, const void* a1, long long ld1, int C, const float* w,
                               const float* b, long long npix, float* logits, cudaStream_t st) {
  const long long threads = npix * 8;
  launch_k(head_fwd_hp_kernel, dim3(static_cast<int>((threads + 255) / 256)), dim3(256), 0, st,
           reinterpret_cast<const __nv_bfloat16*>(a0), ld0, reinterpret_cast<const __nv_bfloat16*>(a1), ld1, C, w, b, npix,
           logits);
  return cudaGetLastError();
}

cudaError_t launch_colsum_hp(const void* x, long long ld, int C, const float* wgt, long long npix, int nblk,
                             float* partial, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(256 / (C / 8)) * C * sizeof(float);
  launch_k(colsum_hp_kernel, dim3(nblk), dim3(256), smem, st, reinterpret_cast<const __nv_bfloat16*>(x), ld, C, wgt, npix,
           partial);
  return cudaGetLastError();
}

}  // namespace b200cd
