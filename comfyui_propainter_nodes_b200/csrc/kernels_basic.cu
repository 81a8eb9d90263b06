// Layout conversion and small elementwise kernels (HBM-bound, coalesced / vectorised).
#include "kernels.cuh"

namespace {

constexpr int TPB = 256;
inline int nblocks(long long n, int per = TPB) { return (int)((n + per - 1) / per); }

// NCHW fp32 -> NHWC fp16 into a channel slice [co, co+C) of a pixel of dst_cs elements; channels
// [co+C, co+zero_to) are zero-filled (padding channels must be finite for the tensor-core path).
__global__ void nchw_f32_to_nhwc_f16(const float* __restrict__ src, __half* __restrict__ dst, int C, long long HW,
                                     long long npix, int dst_cs, int dst_co, int zero_to) {
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const long long n = p / HW, hw = p - n * HW;
  __half* d = dst + p * dst_cs + dst_co;
  const float* s = src + n * C * HW + hw;
  for (int c = 0; c < C; ++c) d[c] = __float2half_rn(s[c * HW]);
  for (int c = C; c < zero_to; ++c) d[c] = __float2half_rn(0.f);
}

__global__ void nhwc_f16_to_nchw_f32(const __half* __restrict__ src, int src_cs, int src_co, float* __restrict__ dst,
                                     int C, long long HW, long long npix) {
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const long long n = p / HW, hw = p - n * HW;
  const __half* s = src + p * src_cs + src_co;
  for (int c = 0; c < C; ++c) dst[(n * C + c) * HW + hw] = __half2float(s[c]);
}

// Bilinear x2 upsample, align_corners=True (reference deconv: F.interpolate(scale_factor=2, 'bilinear', True)).
// One thread per (output pixel, 8-channel vector); grid = (x chunks, output rows, images) so all index math is
// 32-bit and the row interpolation is uniform per block.
__global__ void __launch_bounds__(256) upsample2x_ac(const __half* __restrict__ src, int src_cs, int src_co,
                                                     __half* __restrict__ dst, int dst_cs, int dst_co, int H, int W, int C8) {
  const int OW = 2 * W, OH = 2 * H;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)(OW * C8)) return;
  const int c8 = idx % (unsigned)C8, ox = idx / (unsigned)C8;
  const int oy = blockIdx.y, n = blockIdx.z;
  const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
  const float sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  const float fy = sy * oy, fx = sx * ox;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const __half* base = src + (long long)n * H * W * src_cs + src_co + c8 * 8;
  const uint4 a = *reinterpret_cast<const uint4*>(base + (long long)(y0 * W + x0) * src_cs);
  const uint4 b = *reinterpret_cast<const uint4*>(base + (long long)(y0 * W + x1) * src_cs);
  const uint4 c = *reinterpret_cast<const uint4*>(base + (long long)(y1 * W + x0) * src_cs);
  const uint4 d = *reinterpret_cast<const uint4*>(base + (long long)(y1 * W + x1) * src_cs);
  const __half2* ah = reinterpret_cast<const __half2*>(&a);
  const __half2* bh = reinterpret_cast<const __half2*>(&b);
  const __half2* ch = reinterpret_cast<const __half2*>(&c);
  const __half2* dh = reinterpret_cast<const __half2*>(&d);
  __align__(16) __half2 o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __half22float2(ah[i]), fb = __half22float2(bh[i]), fc = __half22float2(ch[i]),
                 fd = __half22float2(dh[i]);
    o[i] = __floats2half2_rn(w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x,
                             w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y);
  }
  __half* dp = dst + (((long long)n * OH + oy) * OW + ox) * dst_cs + dst_co + c8 * 8;
  *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<uint4*>(o);
}

__global__ void copy_channels(const __half* __restrict__ src, int src_cs, int src_co, __half* __restrict__ dst,
                              int dst_cs, int dst_co, long long npix, int C8) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= npix * C8) return;
  const int c8 = idx % C8;
  const long long p = idx / C8;
  *reinterpret_cast<uint4*>(dst + p * dst_cs + dst_co + c8 * 8) =
      *reinterpret_cast<const uint4*>(src + p * src_cs + src_co + c8 * 8);
}

// dst block j <- src block idx[j]; blocks are `block16` 16-byte units (frame-sized gathers for window batching)
__global__ void gather_blocks(uint4* __restrict__ dst, const uint4* __restrict__ src, const int* __restrict__ idx,
                              long long n, long long block16) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n * block16) return;
  const long long j = i / block16, u = i - j * block16;
  dst[i] = src[(long long)idx[j] * block16 + u];
}

__global__ void fill_f16(__half* dst, long long n, float v) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(v);
}

}  // namespace

int pp_k_nchw_f32_to_nhwc_f16(const float* src, __half* dst, int N, int C, int H, int W, int dst_cs, int dst_co,
                              int zero_fill_to, cudaStream_t st) {
  const long long npix = (long long)N * H * W;
  if (npix == 0) return PP_OK;
  nchw_f32_to_nhwc_f16<<<nblocks(npix), TPB, 0, st>>>(src, dst, C, (long long)H * W, npix, dst_cs, dst_co,
                                                       zero_fill_to);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_nhwc_f16_to_nchw_f32(const __half* src, int src_cs, int src_co, float* dst, int N, int C, int H, int W,
                              cudaStream_t st) {
  const long long npix = (long long)N * H * W;
  if (npix == 0) return PP_OK;
  nhwc_f16_to_nchw_f32<<<nblocks(npix), TPB, 0, st>>>(src, src_cs, src_co, dst, C, (long long)H * W, npix);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_upsample2x(const __half* src, int src_cs, int src_co, __half* dst, int dst_cs, int dst_co, int N, int H,
                    int W, int C, cudaStream_t st) {
  PP_REQUIRE(C % 8 == 0 && src_cs % 8 == 0 && dst_cs % 8 == 0 && src_co % 8 == 0 && dst_co % 8 == 0,
             "upsample2x: channels must be multiples of 8");
  if ((long long)N * H * W == 0) return PP_OK;
  PP_REQUIRE(2 * H <= 65535 && N <= 65535, "upsample2x: %d rows / %d images exceed the grid limits", 2 * H, N);
  const dim3 grid(pp_ceil_div(2 * W * (C / 8), 256), 2 * H, N);
  upsample2x_ac<<<grid, 256, 0, st>>>(src, src_cs, src_co, dst, dst_cs, dst_co, H, W, C / 8);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_copy_channels(const __half* src, int src_cs, int src_co, __half* dst, int dst_cs, int dst_co, long long npix,
                       int C, cudaStream_t st) {
  PP_REQUIRE(C % 8 == 0 && src_cs % 8 == 0 && dst_cs % 8 == 0 && src_co % 8 == 0 && dst_co % 8 == 0,
             "copy_channels: channels must be multiples of 8");
  if (npix == 0) return PP_OK;
  copy_channels<<<nblocks(npix * (C / 8)), TPB, 0, st>>>(src, src_cs, src_co, dst, dst_cs, dst_co, npix, C / 8);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_gather_blocks(void* dst, const void* src, const int* idx_dev, long long n, long long block_bytes,
                       cudaStream_t st) {
  PP_REQUIRE(block_bytes % 16 == 0, "gather_blocks: block size must be a multiple of 16 bytes");
  if (n == 0) return PP_OK;
  const long long b16 = block_bytes / 16;
  gather_blocks<<<nblocks(n * b16), TPB, 0, st>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src), idx_dev, n,
                                                   b16);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_fill_f16(__half* dst, long long n, float v, cudaStream_t st) {
  if (n == 0) return PP_OK;
  fill_f16<<<nblocks(n), TPB, 0, st>>>(dst, n, v);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}
