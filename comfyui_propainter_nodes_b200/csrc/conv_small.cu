// Direct 3x3 convolution for layers with a handful of output channels (RAFT flow head conv2: 256->2,
// generator decoder tail: 64->3 + tanh, flow-completion tail: 32->2).  These layers have no arithmetic
// intensity (<= 3 outputs per input vector): as implicit GEMMs they only stream the 9x im2col-amplified
// input through the tensor-core pipeline.  Here every input vector is read once per tap straight from
// L1/L2 (neighbouring pixels of a block share rows) and reduced with warp shuffles.
//
// NHWC fp16 input, weights fp16 [cout][9][C] (tap-major, (ky,kx) order), zero padding 1, stride 1.
// A group of C/8 lanes owns one output pixel: each lane multiplies its 8 channels for the 9 taps, then the
// group reduces with shuffles.
#include "kernels.cuh"

namespace {

template <int COUT>
__global__ void __launch_bounds__(256) conv3x3_small(const __half* __restrict__ x, int x_cs, int x_co,
                                                     const __half* __restrict__ w, const float* __restrict__ bias,
                                                     void* __restrict__ out, int out_cs, int out_co, int out_fp32,
                                                     int act_tanh, int N, int H, int W, int C) {
  extern __shared__ __half sw[];  // [COUT][9][C]
  for (int i = threadIdx.x; i < COUT * 9 * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int lpp = C >> 3;                 // lanes per pixel (4, 8, 16 or 32)
  const int ppw = 32 / lpp;               // pixels per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / lpp, cl = lane - sub * lpp;   // pixel slot inside the warp, channel-vector index
  const long long total = (long long)N * H * W;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long pix = warp_global * ppw + sub;
  const bool valid = pix < total;
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
  if (valid) {
    const int px = pix % W;
    const long long t = pix / W;
    const int py = t % H;
    const long long n = t / H;
    const __half* xb = x + (n * H * W) * x_cs + x_co + cl * 8;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const uint4 q = *reinterpret_cast<const uint4*>(xb + ((long long)yy * W + xx) * x_cs);
      const __half2* hq = reinterpret_cast<const __half2*>(&q);
      float xv[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(hq[e]); xv[2 * e] = f.x; xv[2 * e + 1] = f.y; }
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const uint4 wq = *reinterpret_cast<const uint4*>(sw + (o * 9 + tap) * C + cl * 8);
        const __half2* hw = reinterpret_cast<const __half2*>(&wq);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(hw[e]);
          acc[o] += xv[2 * e] * f.x + xv[2 * e + 1] * f.y;
        }
      }
    }
  }
  // reduce over the lanes of the pixel group
#pragma unroll
  for (int o = 0; o < COUT; ++o)
    for (int off = lpp >> 1; off > 0; off >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
  if (valid && cl == 0) {
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      float v = acc[o] + bias[o];
      if (act_tanh) v = tanhf(v);
      const long long oi = pix * out_cs + out_co + o;
      if (out_fp32) reinterpret_cast<float*>(out)[oi] = v;
      else reinterpret_cast<__half*>(out)[oi] = __float2half_rn(v);
    }
  }
}

}  // namespace

int pp_k_conv3x3_small(const __half* x, int x_cs, int x_co, const __half* w, const float* bias, int cout, void* out,
                       int out_cs, int out_co, int out_fp32, int act_tanh, int N, int H, int W, int C,
                       cudaStream_t st) {
  PP_REQUIRE(cout >= 1 && cout <= 3, "conv3x3_small: cout=%d not in [1,3]", cout);
  PP_REQUIRE(C == 32 || C == 64 || C == 128 || C == 256, "conv3x3_small: C=%d must be 32/64/128/256", C);
  PP_REQUIRE(x_cs % 8 == 0 && x_co % 8 == 0, "conv3x3_small: input not 16-byte aligned");
  const long long total = (long long)N * H * W;
  if (total == 0) return PP_OK;
  const int ppw = 32 / (C / 8);
  const long long warps = (total + ppw - 1) / ppw;
  const unsigned blocks = (unsigned)((warps + 7) / 8);
  const size_t smem = (size_t)cout * 9 * C * sizeof(__half);
  switch (cout) {
    case 1: conv3x3_small<1><<<blocks, 256, smem, st>>>(x, x_cs, x_co, w, bias, out, out_cs, out_co, out_fp32, act_tanh, N, H, W, C); break;
    case 2: conv3x3_small<2><<<blocks, 256, smem, st>>>(x, x_cs, x_co, w, bias, out, out_cs, out_co, out_fp32, act_tanh, N, H, W, C); break;
    default: conv3x3_small<3><<<blocks, 256, smem, st>>>(x, x_cs, x_co, w, bias, out, out_cs, out_co, out_fp32, act_tanh, N, H, W, C); break;
  }
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}
