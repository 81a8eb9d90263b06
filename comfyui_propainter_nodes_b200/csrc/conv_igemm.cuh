// Implicit-GEMM convolution / linear layer on tcgen05 tensor cores (sm_100a).
//
//   out[m][n] = epilogue( sum_k A[m][k] * Wt[n][k] + bias[n] )
//   m = (image, oy, ox) flattened, n = output channel, k = (ky, kx, ci)
//
// Activations are NHWC fp16 with channel strides that are multiples of 8 (16 B), so every 16-byte
// im2col vector lies inside one filter tap.  A conv input may be the channel-concatenation of up to
// four tensors (segments): the torch.cat calls of the reference become address arithmetic here.
#pragma once
#include "pp_common.cuh"

enum PPAct : int { PP_ACT_NONE = 0, PP_ACT_RELU = 1, PP_ACT_LRELU = 2, PP_ACT_SIGMOID = 3, PP_ACT_TANH = 4,
                   PP_ACT_GELU = 5 };

enum PPEpi : int {
  PP_EPI_STD = 0,      // v = act2( act1(acc + bias) * scale + residual )
  PP_EPI_GRU_ZR = 1,   // n < half: z = sigmoid(v) -> out ; n >= half: r = sigmoid(v), out2 = r * h
  PP_EPI_GRU_H = 2,    // q = tanh(v); out = (1 - z) * h + z * q   (h = aux0, z = aux1)
};

struct PPConvSeg {
  const __half* ptr;
  int cstride;   // elements between consecutive pixels
  int coff;      // first channel of this segment inside the pixel
  int gstep;     // added to coff per group index
  int cbegin;    // first conv-input channel (per group) covered by this segment
  int cend;      // one past the last conv-input channel covered (multiple of 8)
  int cvalid;    // 0, or the number of channels that really exist in memory (< cend - cbegin): the rest read as
                 // zero (TMA out-of-bounds fill; halo kernel only) -- lets a 32-channel tensor feed 64-wide K chunks
};

struct PPConvParams {
  PPConvSeg seg[4];
  int nseg;
  int N, H, W, OH, OW;
  int Cin;                // per-group input channels as seen by the kernel (multiple of 8)
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int pad_replicate;      // 0: zeros outside, 1: clamp coordinates (replicate padding)
  int ups2x;              // the input tensor is [N][H/2][W/2]: bilinear x2 (align_corners=True) on the fly (halo kernel only)
  int K_total;            // kh*kw*Cin
  int num_kc;             // ceil(K_total / 64)
  int M_total;            // N*OH*OW
  const __half* wpacked;  // [groups][num_kc][Cout_g_pad] rows of 64 fp16, 128B-swizzled tile image
  const float* bias;      // [groups*Cout_g] or nullptr
  int Cout_g, Cout_g_pad, BN, groups;
  int stages;
  int vec_ok;             // set by the launcher: every epilogue pointer/stride allows 16-byte accesses on full runs
  int vec32_ok;           // ... and fp16 epilogue operands are 32-byte aligned: one 256-bit access per 16 channels
  // epilogue
  int epi, act1, act2;
  float slope, scale;
  void* out; int out_cstride, out_coff, out_gstep, out_fp32;
  const __half* aux0; int aux0_cstride, aux0_coff;   // residual (STD) or h (GRU)
  const __half* aux1; int aux1_cstride, aux1_coff;   // z (GRU_H)
  __half* out2; int out2_cstride, out2_coff;          // r*h destination (GRU_ZR)
};

int pp_launch_conv(const PPConvParams& p, cudaStream_t stream);
// which kernel the last pp_launch_conv of this thread went to: 'h' TMA halo-tile kernel, 'i' cp.async implicit GEMM,
// 'p' recorded into a multi-layer program (profiling labels)
char pp_last_conv_kind();
// conv_halo.cu: TMA halo-tile kernel for stride-1 convolutions (dispatched from pp_launch_conv when eligible;
// PP_CONV_HALO=0 in the environment disables it).  `p` must already carry num_kc / vec_ok.
int pp_conv_halo_eligible(const PPConvParams& p);
int pp_launch_conv_halo(const PPConvParams& p, cudaStream_t stream);

// Multi-layer programs (conv_halo.cu): between pp_prog_begin() and pp_prog_end() every eligible convolution handed to
// pp_launch_conv and every pp_k_dcn_sample call of this thread is RECORDED instead of launched; pp_prog_end launches
// the recorded, mutually dependent layers as ONE persistent kernel with grid-wide barriers between them.
struct PPDcnArgs;
struct PPProgRecorder;
bool pp_prog_recording();
int pp_prog_begin();
void pp_prog_abort();
int pp_prog_eligible(const PPConvParams& p);
int pp_prog_record_conv(const PPConvParams& p);
int pp_prog_record_dcn(const PPDcnArgs& a);
int pp_prog_end(unsigned int* counter, unsigned int* arrivals, cudaStream_t stream);
