// 3x3 convolutions with a handful of output channels (RAFT flow head conv2: 256->2, generator decoder tail:
// 64->3 + tanh, flow-completion tail: 32->2) have no arithmetic intensity: as 3x3 implicit GEMMs they only stream
// the 9x im2col-amplified input through the pipeline.  They are split into
//   (1) a 1x1 GEMM on the tensor-core kernel that reads every input pixel ONCE and produces the 9*cout per-tap
//       partial products  z[p][tap*cout + c] = sum_ch x[p][ch] * W[c][ch][tap]          (conv_igemm.cu)
//   (2) this kernel: out[p][c] = act(bias[c] + sum_tap z[p + d(tap)][tap*cout + c]), zero outside the image,
// which is exactly the zero-padded 3x3 convolution (the conv is linear in the taps).
#include "kernels.cuh"

namespace {

template <typename ZT>
__global__ void tap_sum3x3(const ZT* __restrict__ z, int z_cs, const float* __restrict__ bias, int cout,
                           void* __restrict__ out, int out_cs, int out_co, int out_fp32, int act_tanh, int N, int H,
                           int W) {
  const long long total = (long long)N * H * W;
  const long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int px = pix % W;
  const long long t = pix / W;
  const int py = t % H;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int yy = py + dy, xx = px + dx;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const ZT* zp = z + (pix + (long long)dy * W + dx) * z_cs + tap * cout;
    for (int c = 0; c < cout; ++c) acc[c] += (float)zp[c];
  }
  for (int c = 0; c < cout; ++c) {
    float v = acc[c] + bias[c];
    if (act_tanh) v = tanhf(v);
    const long long oi = pix * out_cs + out_co + c;
    if (out_fp32) reinterpret_cast<float*>(out)[oi] = v;
    else reinterpret_cast<__half*>(out)[oi] = __float2half_rn(v);
  }
}

}  // namespace

int pp_k_tap_sum3x3(const void* z, int z_fp32, int z_cs, const float* bias, int cout, void* out, int out_cs, int out_co,
                    int out_fp32, int act_tanh, int N, int H, int W, cudaStream_t st) {
  PP_REQUIRE(cout >= 1 && cout <= 3, "tap_sum3x3: cout=%d not in [1,3]", cout);
  const long long total = (long long)N * H * W;
  if (total == 0) return PP_OK;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (z_fp32)
    tap_sum3x3<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(z), z_cs, bias, cout, out, out_cs, out_co, out_fp32,
                                               act_tanh, N, H, W);
  else
    tap_sum3x3<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(z), z_cs, bias, cout, out, out_cs, out_co,
                                                out_fp32, act_tanh, N, H, W);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}
