// Device-side pre/post-processing with the reference's integer semantics (reference
// utils/image_utils.py:106-197, 276-290).  Used when no resize is needed (input size == processing size);
// resizing goes through PIL on the host so the bicubic filter stays bit-identical.
#include "kernels.cuh"

namespace {

constexpr int TPB = 256;
inline int nblocks(long long n, int per = TPB) { return (int)((n + per - 1) / per); }

// IMAGE [T,H,W,3] float 0..1 -> uint8 by *255, clip, truncate (image_utils.py:112) and
// frames [T,3,H,W] = u8/255*2-1 (image_utils.py:186-190)
__global__ void quantize_frames(const float* __restrict__ img, uint8_t* __restrict__ u8, float* __restrict__ frames,
                                long long HW, long long total) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // over T*H*W pixels
  if (idx >= total) return;
  const long long t = idx / HW, p = idx - t * HW;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = fminf(fmaxf(__fmul_rn(img[idx * 3 + c], 255.f), 0.f), 255.f);
    const uint8_t q = (uint8_t)(int)v;
    u8[idx * 3 + c] = q;
    frames[(t * 3 + c) * HW + p] = __fsub_rn(__fmul_rn(__fdiv_rn((float)q, 255.f), 2.f), 1.f);
  }
}

// MASK [Tm,H,W] float -> 8-bit ((m*255).clamp(0,255).byte(), image_utils.py:128-134), then N iterations of the
// cross-shaped binary dilation of scipy.ndimage.binary_dilation (== dilation by the L1 ball of radius N, zero
// border), or the > 0.1 threshold when N == 0 (image_utils.py:152-170).  Output float {0,1}, broadcast to T frames.
__global__ void mask_u8(const float* __restrict__ m, uint8_t* __restrict__ nz, long long total) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const float v = fminf(fmaxf(__fmul_rn(m[idx], 255.f), 0.f), 255.f);
  nz[idx] = ((uint8_t)(int)v) != 0 ? 1 : 0;
}

__global__ void dilate_diamond(const uint8_t* __restrict__ nz, float* __restrict__ out, int Tm, int T, int H, int W,
                               int radius) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)T * H * W) return;
  const int x = idx % W;
  long long r = idx / W;
  const int y = r % H;
  const int t = r / H;
  const uint8_t* src = nz + (long long)(Tm == 1 ? 0 : t) * H * W;
  int any = 0;
  for (int dy = -radius; dy <= radius && !any; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    const int span = radius - (dy < 0 ? -dy : dy);
    const int x0 = max(0, x - span), x1 = min(W - 1, x + span);
    for (int xx = x0; xx <= x1; ++xx)
      if (src[(long long)yy * W + xx]) { any = 1; break; }
  }
  out[idx] = any ? 1.f : 0.f;
}

// uint8 [T,H,W,3] -> float32 / 255 (handle_output, image_utils.py:281-283)
__global__ void u8_to_unit_float(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __fdiv_rn((float)src[i], 255.f);
}

}  // namespace

int pp_k_quantize_frames(const float* img, uint8_t* u8, float* frames, int T, int H, int W, cudaStream_t st) {
  const long long HW = (long long)H * W, total = HW * T;
  quantize_frames<<<nblocks(total), TPB, 0, st>>>(img, u8, frames, HW, total);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_prepare_masks(const float* mask, int Tm, int T, int H, int W, int iters_flow, int iters_dil, uint8_t* scratch,
                       float* flow_masks, float* masks_dilated, cudaStream_t st) {
  PP_REQUIRE(Tm == 1 || Tm == T, "prepare_masks: mask length %d must be 1 or %d", Tm, T);
  const long long n = (long long)Tm * H * W;
  mask_u8<<<nblocks(n), TPB, 0, st>>>(mask, scratch, n);
  dilate_diamond<<<nblocks((long long)T * H * W), TPB, 0, st>>>(scratch, flow_masks, Tm, T, H, W, iters_flow);
  dilate_diamond<<<nblocks((long long)T * H * W), TPB, 0, st>>>(scratch, masks_dilated, Tm, T, H, W, iters_dil);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_u8_to_unit_float(const uint8_t* src, float* dst, long long n, cudaStream_t st) {
  u8_to_unit_float<<<nblocks(n), TPB, 0, st>>>(src, dst, n);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}
