// RAFT-specific HBM-bound kernels: instance norm, correlation pyramid pooling + 9x9x4 lookup,
// GRU state plumbing, convex upsampling.  Reference call sites are cited per kernel.
#include "kernels.cuh"

namespace {

constexpr int TPB = 256;
inline int nblocks(long long n, int per = TPB) { return (int)((n + per - 1) / per); }

// ------------------------------------------------------------------------------------------------
// InstanceNorm2d(affine=False) statistics: per (image, channel) sum and sum of squares.
// grid (chunks, N); block = 256 threads; thread t owns channel pair (t % (C/2)) and strides pixels.
// (reference: RAFT/extractor.py fnet norm layers, F.instance_norm eps=1e-5, biased variance)
// ------------------------------------------------------------------------------------------------
// Deterministic: no floating-point atomics.  A block reduces its pixel lanes in lane order, writes its partial sums to
// `partial[n][block][2C]`, and the block that arrives last (integer counter) adds the partials in block order -- the
// result does not depend on scheduling, so RAFT (and everything after it) is bit-reproducible run to run and between
// the single-GPU and the sharded multi-GPU execution.
__global__ void instnorm_stats(const __half* __restrict__ x, int HW, int C, float* __restrict__ sums,
                               float* __restrict__ partial, unsigned int* __restrict__ counters, int pix_per_block) {
  extern __shared__ float sm[];  // [lanes][2C]
  __shared__ bool last;
  const int n = blockIdx.y;
  const int C2 = C >> 1;
  const int lanes = blockDim.x / C2;  // pixel lanes per block
  const int cp = threadIdx.x % C2;
  const int pl = threadIdx.x / C2;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  if (pl < lanes) {
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    const __half2* base = reinterpret_cast<const __half2*>(x + ((long long)n * HW) * C) + cp;
    for (int p = p0 + pl; p < p1; p += lanes) {
      const float2 v = __half22float2(base[(long long)p * C2]);
      s0 += v.x; s1 += v.y; q0 += v.x * v.x; q1 += v.y * v.y;
    }
    float* row = sm + pl * 2 * C;
    row[2 * cp] = s0; row[2 * cp + 1] = s1; row[C + 2 * cp] = q0; row[C + 2 * cp + 1] = q1;
  }
  __syncthreads();
  float* mine = partial + ((long long)n * gridDim.x + blockIdx.x) * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += sm[l * 2 * C + i];
    mine[i] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&counters[n], 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence();
    const float* all = partial + (long long)n * gridDim.x * 2 * C;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      float acc = 0.f;
      for (unsigned b = 0; b < gridDim.x; ++b) acc += __ldcg(all + (long long)b * 2 * C + i);
      sums[(long long)n * 2 * C + i] = acc;
    }
  }
}

// out = [relu]( (x - mean) * rstd );  if residual: out = relu(residual + out)   (ResidualBlock tail)
__global__ void __launch_bounds__(256) instnorm_apply(const __half* __restrict__ x, const float* __restrict__ sums,
                                                      const __half* __restrict__ residual, __half* __restrict__ out, int HW,
                                                      int C, int relu) {
  // grid = (chunks of one image's HW*C/2 channel pairs, images): 32-bit index math, statistics row uniform per block
  const int C2 = C >> 1;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)(HW * C2)) return;
  const int n = blockIdx.y;
  const int cp = i % (unsigned)C2;
  const long long idx = (long long)n * HW * C2 + i;
  const float inv = 1.f / (float)HW;
  const float* s = sums + (long long)n * 2 * C;
  const float m0 = s[2 * cp] * inv, m1 = s[2 * cp + 1] * inv;
  const float v0 = fmaxf(s[C + 2 * cp] * inv - m0 * m0, 0.f), v1 = fmaxf(s[C + 2 * cp + 1] * inv - m1 * m1, 0.f);
  const float r0 = rsqrtf(v0 + 1e-5f), r1 = rsqrtf(v1 + 1e-5f);
  const float2 xv = __half22float2(reinterpret_cast<const __half2*>(x)[idx]);
  float a = (xv.x - m0) * r0, b = (xv.y - m1) * r1;
  if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
  if (residual != nullptr) {
    const float2 rv = __half22float2(reinterpret_cast<const __half2*>(residual)[idx]);
    a = fmaxf(a + rv.x, 0.f);
    b = fmaxf(b + rv.y, 0.f);
  }
  reinterpret_cast<__half2*>(out)[idx] = __floats2half2_rn(a, b);
}

// ------------------------------------------------------------------------------------------------
// Pack a row-major [G][R][K] fp16 matrix into the swizzled B-operand tile image consumed by the
// tcgen05 GEMM ([G][K/64][R_pad] rows of 128 B, 16-byte chunk index XOR (row & 7)); rows >= R are zero.
// Used for the all-pairs correlation, where fmap2 plays the role of the weights (RAFT/corr.py:52-60).
// ------------------------------------------------------------------------------------------------
__global__ void pack_b_operand(const __half* __restrict__ src, __half* __restrict__ dst, int R, int R_pad, int K,
                               long long total) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // one 16-byte chunk each
  if (idx >= total) return;
  const int chunks_per_row = K / 8;
  const int ch = idx % chunks_per_row;
  long long t = idx / chunks_per_row;
  const int r = t % R_pad;
  const int g = t / R_pad;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (r < R) v = *reinterpret_cast<const uint4*>(src + ((long long)g * R + r) * K + ch * 8);
  const int kc = ch >> 3, c = ch & 7;
  const int num_kc = K / 64;
  __half* d = dst + (((long long)g * num_kc + kc) * R_pad + r) * 64 + ((c ^ (r & 7)) << 3);
  *reinterpret_cast<uint4*>(d) = v;
}

// 2x2 average pooling of every [h][w] correlation map (F.avg_pool2d(corr, 2, stride=2), corr.py:25-27).
// One block per query map (32-bit index math); each thread produces output pairs from two 8-byte row reads.
__global__ void __launch_bounds__(128) corr_pool(const __half* __restrict__ src, __half* __restrict__ dst, int h, int w) {
  const int oh = h >> 1, ow = w >> 1;
  const long long q = blockIdx.x;
  const __half* s = src + q * (long long)(h * w);
  __half* d = dst + q * (long long)(oh * ow);
  for (int i = threadIdx.x; i < oh * ow; i += blockDim.x) {
    const int oy = i / ow, ox = i - oy * ow;
    const __half* r = s + 2 * oy * w + 2 * ox;
    const float a = __half2float(r[0]) + __half2float(r[1]) + __half2float(r[w]) + __half2float(r[w + 1]);
    d[i] = __float2half_rn(0.25f * a);
  }
}

// ------------------------------------------------------------------------------------------------
// Correlation lookup (CorrBlock.__call__, corr.py:29-50; bilinear_sampler, RAFT/utils/utils.py:66-80).
// Output channel c = l*81 + i*9 + j samples level l at (x/2^l + (i-4), y/2^l + (j-4)), bilinear,
// zeros outside, align_corners=True.
// ------------------------------------------------------------------------------------------------
struct CorrLevels {
  const __half* p[4];
};

// One warp per query pixel.  For a given (pixel, level) all 81 outputs share the same bilinear fractions
// (the window offsets are integers), so the warp stages the tap window of each level in shared memory once
// (zero outside the map) and every lane then blends 4 staged taps per output:
//   out[l*81 + i*9 + j] = bilerp(T_l[j..j+1][i..i+1])      (i moves x, j moves y -- the meshgrid quirk)
// Taps are fetched as aligned fp16 pairs (12 columns starting at the even column <= x0-4; map widths are even at
// every level for the sizes ProPainter produces, odd widths take the scalar path), which halves the load
// instructions; the output loop runs level by level with the level's fractions in registers.
// Global traffic per pixel = the algorithmic 8 B coords + 4x100 taps + 324 outputs; stores are contiguous.
constexpr int LOOKUP_WARPS = 8;
constexpr int TAP_COLS = 12;   // staged columns per tap row
constexpr int TAP_STRIDE = 10 * TAP_COLS;

__global__ void __launch_bounds__(LOOKUP_WARPS * 32) corr_lookup(CorrLevels lv, const float* __restrict__ coords,
                                                                 __half* __restrict__ out, int out_cs, long long nq,
                                                                 int h8, int w8) {
  __shared__ float taps[LOOKUP_WARPS][4][TAP_STRIDE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * LOOKUP_WARPS + warp;
  if (q >= nq) return;
  const float cx = coords[q * 2], cy = coords[q * 2 + 1];
  float fa[4], fb[4];   // bilinear fractions and even-column phase of every level: registers of every lane
  int ph[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float inv = 1.f / (float)(1 << l);
    const int h = h8 >> l, w = w8 >> l;
    const float x = cx * inv, y = cy * inv;
    const float fx = floorf(x), fy = floorf(y);
    const int x0 = (int)fx - 4, y0 = (int)fy - 4;
    const int xa = x0 & ~1;                       // even column <= x0 (also for negative x0)
    fa[l] = x - fx; fb[l] = y - fy; ph[l] = x0 - xa;
    const __half* m = lv.p[l] + q * (long long)(h * w);
    float* T = taps[warp][l];
    if ((w & 1) == 0) {
#pragma unroll
      for (int t2 = lane; t2 < 60; t2 += 32) {    // 10 rows x 6 aligned pairs
        const int ty = t2 / 6, tp = t2 - ty * 6;
        const int yy = y0 + ty, xx = xa + 2 * tp;
        float2 v = make_float2(0.f, 0.f);
        if ((unsigned)yy < (unsigned)h && (unsigned)xx < (unsigned)w)
          v = __half22float2(*reinterpret_cast<const __half2*>(m + yy * w + xx));
        T[ty * TAP_COLS + 2 * tp] = v.x;
        T[ty * TAP_COLS + 2 * tp + 1] = v.y;
      }
    } else {
      for (int t = lane; t < TAP_STRIDE; t += 32) {
        const int ty = t / TAP_COLS, tx = t - ty * TAP_COLS;
        const int yy = y0 + ty, xx = xa + tx;
        float v = 0.f;
        if ((unsigned)yy < (unsigned)h && (unsigned)xx < (unsigned)w) v = __half2float(m[yy * w + xx]);
        T[t] = v;
      }
    }
  }
  __syncwarp();
  __half* o = out + q * out_cs;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float a = fa[l], b = fb[l];
    const float* T = taps[warp][l] + ph[l];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int r = lane + 32 * k;                 // output i*9 + j of this level
      if (r < 81) {
        const int i = (r * 57) >> 9, j = r - 9 * i;  // r / 9, r % 9 for r < 81
        const float* t = T + j * TAP_COLS + i;
        const float top = t[0] + a * (t[1] - t[0]);
        const float bot = t[TAP_COLS] + a * (t[TAP_COLS + 1] - t[TAP_COLS]);
        o[l * 81 + r] = __float2half_rn(top + b * (bot - top));
      }
    }
  }
  for (int c = 324 + lane; c < out_cs; c += 32) o[c] = __float2half_rn(0.f);   // padding channels
}

// cnet output -> GRU state: h = tanh(c[:, :128]) into hx[:, 0:128], inp = relu(c[:, 128:]) into hx[:, 128:256]
// (raft.py:119-122)
__global__ void cnet_split(const __half* __restrict__ c, __half* __restrict__ hx, int hx_cs, long long total) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = idx % 256;
  const long long p = idx / 256;
  const float v = __half2float(c[idx]);
  hx[p * hx_cs + ch] = __float2half_rn(ch < 128 ? tanhf(v) : fmaxf(v, 0.f));
}

// coords1 = coords0 (+ delta); flow = coords1 - coords0 written (fp16) to the motion-encoder input
// ([.,8], channels 2..7 zero) and to the last two channels of the GRU input (raft.py:124-140).
__global__ void raft_coords(const float* __restrict__ delta, float* __restrict__ coords1, __half* __restrict__ flow8,
                            __half* __restrict__ hx, int hx_cs, int hx_flow_co, long long total, int P, int w8,
                            int init) {
  long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= total) return;
  const int p = q % P;
  const float x0 = (float)(p % w8), y0 = (float)(p / w8);
  float x, y;
  if (init) { x = x0; y = y0; }
  else { x = coords1[2 * q] + delta[2 * q]; y = coords1[2 * q + 1] + delta[2 * q + 1]; }
  coords1[2 * q] = x;
  coords1[2 * q + 1] = y;
  const __half fx = __float2half_rn(x - x0), fy = __float2half_rn(y - y0);
  __half* f = flow8 + q * 8;
  f[0] = fx; f[1] = fy;
  if (init) for (int k = 2; k < 8; ++k) f[k] = __float2half_rn(0.f);
  hx[q * hx_cs + hx_flow_co] = fx;
  hx[q * hx_cs + hx_flow_co + 1] = fy;
}

// Convex 8x upsampling (RAFT.upsample_flow, raft.py:81-92): softmax over the 9 neighbours' logits
// mask[k*64 + sy*8 + sx], weighted sum of 8*flow (3x3 unfold, zero padding).  Output NCHW fp32.
__global__ void convex_upsample(const float* __restrict__ coords1, const __half* __restrict__ mask,
                                float* __restrict__ out, int B, int h8, int w8) {
  const int H = 8 * h8, W = 8 * w8;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int X = idx % W;
  long long t = idx / W;
  const int Y = t % H;
  const int b = t / H;
  const int x = X >> 3, sx = X & 7, y = Y >> 3, sy = Y & 7;
  const long long q = ((long long)b * h8 + y) * w8 + x;
  const __half* mk = mask + q * 576 + sy * 8 + sx;
  float lg[9], mx = -1e30f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { lg[k] = __half2float(mk[k * 64]); mx = fmaxf(mx, lg[k]); }
  float den = 0.f, ux = 0.f, uy = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float e = __expf(lg[k] - mx);
    den += e;
    const int ny = y + k / 3 - 1, nx = x + k % 3 - 1;
    if (ny >= 0 && ny < h8 && nx >= 0 && nx < w8) {
      const long long nq = ((long long)b * h8 + ny) * w8 + nx;
      ux += e * 8.f * (coords1[2 * nq] - (float)nx);
      uy += e * 8.f * (coords1[2 * nq + 1] - (float)ny);
    }
  }
  const long long HWl = (long long)H * W;
  out[((long long)b * 2) * HWl + (long long)Y * W + X] = ux / den;
  out[((long long)b * 2 + 1) * HWl + (long long)Y * W + X] = uy / den;
}

}  // namespace

size_t pp_k_instnorm_scratch_floats(int N, int HW, int C) {
  return (size_t)N * 2 * C * (pp_ceil_div(HW, 1024) + 1) + (size_t)N + 64;
}

// sums: scratch of pp_k_instnorm_scratch_floats(N, HW, C) floats; the statistics [N][2][C] are its first N*2*C entries
int pp_k_instnorm_stats(const __half* x, int N, int HW, int C, float* sums, cudaStream_t st) {
  PP_REQUIRE(C % 2 == 0 && C <= 256, "instnorm: unsupported C=%d", C);
  const int pix_per_block = 1024;
  const int nblk = pp_ceil_div(HW, pix_per_block);
  float* partial = sums + (size_t)N * 2 * C;
  unsigned int* counters = reinterpret_cast<unsigned int*>(partial + (size_t)N * nblk * 2 * C);
  PP_CUDA_CHECK(cudaMemsetAsync(counters, 0, (size_t)N * sizeof(unsigned int), st));
  dim3 grid(nblk, N);
  const int lanes = 256 / (C / 2);
  instnorm_stats<<<grid, 256, (size_t)lanes * 2 * C * sizeof(float), st>>>(x, HW, C, sums, partial, counters, pix_per_block);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_instnorm_apply(const __half* x, const float* sums, const __half* residual, __half* out, int N, int HW, int C,
                        int relu, cudaStream_t st) {
  if ((long long)N * HW == 0) return PP_OK;
  PP_REQUIRE(N <= 65535 && (long long)HW * (C / 2) < (1LL << 31), "instnorm: %d images of %d pixels exceed the grid limits", N, HW);
  instnorm_apply<<<dim3(pp_ceil_div(HW * (C / 2), 256), N), 256, 0, st>>>(x, sums, residual, out, HW, C, relu);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_pack_b_operand(const __half* src, __half* dst, int G, int R, int R_pad, int K, cudaStream_t st) {
  PP_REQUIRE(K % 64 == 0 && R_pad % 8 == 0, "pack_b_operand: K=%d must be a multiple of 64", K);
  const long long total = (long long)G * R_pad * (K / 8);
  pack_b_operand<<<nblocks(total), TPB, 0, st>>>(src, dst, R, R_pad, K, total);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_corr_pool(const __half* src, __half* dst, long long nq, int h, int w, cudaStream_t st) {
  if (nq * (h / 2) * (w / 2) == 0) return PP_OK;
  PP_REQUIRE(nq < (1LL << 31), "corr_pool: too many query maps");
  corr_pool<<<(unsigned)nq, 128, 0, st>>>(src, dst, h, w);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_corr_lookup(const __half* l0, const __half* l1, const __half* l2, const __half* l3, const float* coords,
                     __half* out, int out_cs, long long nq, int P, int h8, int w8, cudaStream_t st) {
  PP_REQUIRE(out_cs >= 324 && out_cs <= 352, "corr_lookup: out_cs=%d not in [324,352]", out_cs);
  CorrLevels lv;
  lv.p[0] = l0; lv.p[1] = l1; lv.p[2] = l2; lv.p[3] = l3;
  (void)P;
  corr_lookup<<<(unsigned)((nq + LOOKUP_WARPS - 1) / LOOKUP_WARPS), LOOKUP_WARPS * 32, 0, st>>>(lv, coords, out, out_cs,
                                                                                              nq, h8, w8);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_cnet_split(const __half* c, __half* hx, int hx_cs, long long npix, cudaStream_t st) {
  cnet_split<<<nblocks(npix * 256), TPB, 0, st>>>(c, hx, hx_cs, npix * 256);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

// im2col of the 2-channel flow for BasicMotionEncoder.convf1 (7x7, pad 3; update.py:97,105): per pixel the 49 taps x (dx, dy)
// = 98 values in (ky, kx, channel) order, zero outside the map, zero-padded to 128 -> [M][128] fp16, so that the layer runs
// as a K = 128 flat GEMM on the TMA kernel instead of a K = 49 x 8 (6 of 8 channels padding) implicit GEMM.
// One thread per (pixel, 16-byte unit = 4 taps).
__global__ void flow_patch7x7(const __half* __restrict__ flow8, uint4* __restrict__ out, int h, int w, long long total_units) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total_units) return;
  const int u = (int)(idx & 15);
  const long long pix = idx >> 4;
  const int hw = h * w;
  const long long img = pix / hw;
  const int p = (int)(pix - img * hw), y = p / w, x = p - y * w;
  __align__(16) __half2 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int tap = u * 4 + i;
    v[i] = __floats2half2_rn(0.f, 0.f);
    if (tap < 49) {
      const int yy = y + tap / 7 - 3, xx = x + tap % 7 - 3;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w)
        v[i] = *reinterpret_cast<const __half2*>(flow8 + ((img * hw + (long long)yy * w + xx) << 3));
    }
  }
  out[idx] = *reinterpret_cast<uint4*>(v);
}

int pp_k_flow_patch7x7(const __half* flow8, __half* out, int B, int h8, int w8, cudaStream_t st) {
  const long long total = (long long)B * h8 * w8 * 16;
  if (total == 0) return PP_OK;
  flow_patch7x7<<<nblocks(total), TPB, 0, st>>>(flow8, reinterpret_cast<uint4*>(out), h8, w8, total);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_raft_coords_init(float* coords1, __half* flow8, __half* hx, int hx_cs, int hx_flow_co, int B, int h8, int w8,
                          cudaStream_t st) {
  const long long total = (long long)B * h8 * w8;
  raft_coords<<<nblocks(total), TPB, 0, st>>>(nullptr, coords1, flow8, hx, hx_cs, hx_flow_co, total, h8 * w8, w8, 1);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_raft_coords_update(const float* delta, float* coords1, __half* flow8, __half* hx, int hx_cs, int hx_flow_co,
                            int B, int h8, int w8, cudaStream_t st) {
  const long long total = (long long)B * h8 * w8;
  raft_coords<<<nblocks(total), TPB, 0, st>>>(delta, coords1, flow8, hx, hx_cs, hx_flow_co, total, h8 * w8, w8, 0);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_convex_upsample(const float* coords1, const __half* mask, float* out_nchw, int B, int h8, int w8,
                         cudaStream_t st) {
  const long long total = (long long)B * 64 * h8 * w8;
  convex_upsample<<<nblocks(total), TPB, 0, st>>>(coords1, mask, out_nchw, B, h8, w8);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}
