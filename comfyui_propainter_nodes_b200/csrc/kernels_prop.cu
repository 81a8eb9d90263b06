// Flow-guided propagation kernels: fused image-propagation step, flow warps, fb-consistency,
// modulated deformable sampling, flow-completion pack/combine, 1/4 downsampling.
#include <stdlib.h>

#include "kernels.cuh"
#include "conv_igemm.cuh"
#include "dcn_sample.cuh"

namespace {

constexpr int TPB = 256;
inline int nblocks(long long n, int per = TPB) { return (int)((n + per - 1) / per); }

// Sampling position of grid_sample(align_corners=True) for pixel coordinate (x + flow): the reference
// normalises with 2*g/max(W-1,1)-1 (flow_loss_utils.py:41-43) and ATen's CUDA sampler un-normalises with
// ((g+1)/2)*(W-1).  The round trip is kept (no FMA contraction) so `nearest` picks the same texel.
__device__ __forceinline__ float sample_coord(float g, int size) {
  const float d = (float)max(size - 1, 1);
  const float nrm = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, g), d), 1.0f);
  return __fmul_rn(__fdiv_rn(__fadd_rn(nrm, 1.0f), 2.0f), (float)(size - 1));
}

struct Bilin {
  int x0, y0;
  float w00, w01, w10, w11;  // weights, already zero for out-of-range corners
};

__device__ __forceinline__ Bilin bilin_setup(float sx, float sy, int W, int H) {
  Bilin b;
  const float fx = floorf(sx), fy = floorf(sy);
  b.x0 = (int)fx; b.y0 = (int)fy;
  const float ax = sx - fx, ay = sy - fy;
  const bool x0in = b.x0 >= 0 && b.x0 < W, x1in = b.x0 + 1 >= 0 && b.x0 + 1 < W;
  const bool y0in = b.y0 >= 0 && b.y0 < H, y1in = b.y0 + 1 >= 0 && b.y0 + 1 < H;
  b.w00 = (y0in && x0in) ? (1.f - ax) * (1.f - ay) : 0.f;
  b.w01 = (y0in && x1in) ? ax * (1.f - ay) : 0.f;
  b.w10 = (y1in && x0in) ? (1.f - ax) * ay : 0.f;
  b.w11 = (y1in && x1in) ? ax * ay : 0.f;
  return b;
}

// bilinear sample of a 2-channel fp16 flow field [H][W][2]
__device__ __forceinline__ float2 sample_flow2(const __half2* f, const Bilin& b, int W) {
  float2 r = make_float2(0.f, 0.f);
  if (b.w00 != 0.f) { const float2 v = __half22float2(f[b.y0 * W + b.x0]); r.x += b.w00 * v.x; r.y += b.w00 * v.y; }
  if (b.w01 != 0.f) { const float2 v = __half22float2(f[b.y0 * W + b.x0 + 1]); r.x += b.w01 * v.x; r.y += b.w01 * v.y; }
  if (b.w10 != 0.f) { const float2 v = __half22float2(f[(b.y0 + 1) * W + b.x0]); r.x += b.w10 * v.x; r.y += b.w10 * v.y; }
  if (b.w11 != 0.f) { const float2 v = __half22float2(f[(b.y0 + 1) * W + b.x0 + 1]); r.x += b.w11 * v.x; r.y += b.w11 * v.y; }
  return r;
}

// fbConsistencyCheck (model/propainter.py:27-36): |f_p + warp(f_c, f_p)|^2 < 0.01(|f_p|^2 + |warp|^2) + 0.5
__device__ __forceinline__ float fb_valid(float2 fp, float2 fcw) {
  const float dx = fp.x + fcw.x, dy = fp.y + fcw.y;
  const float mag = fp.x * fp.x + fp.y * fp.y + fcw.x * fcw.x + fcw.y * fcw.y;
  return (dx * dx + dy * dy) < (0.01f * mag + 0.5f) ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------------
// One time step of the non-learnable image propagation (BidirectionalPropagation(3, learnable=False),
// model/propainter.py:157-196), fully fused: fb check, nearest warp of the propagated pixels, bilinear
// warp + 0.1 threshold of the propagated mask, mask algebra, blend.
// Pixel layout: 4 x fp16 = (r, g, b, mask) so one 8-byte access moves a whole pixel.
// Algorithmic traffic: cur 8 B + prop gather 8 B (+ 3 more mask taps) + out 8 B + 2 flows 4 B each.
// ------------------------------------------------------------------------------------------------
__global__ void imgprop_step(const uint2* __restrict__ cur, const uint2* __restrict__ prop_in,
                             uint2* __restrict__ prop_out, const __half2* __restrict__ flow_prop,
                             const __half2* __restrict__ flow_check, int H, int W) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  const int x = idx % W, y = idx / W;
  const float2 fp = __half22float2(flow_prop[idx]);
  const float sx = sample_coord((float)x + fp.x, W), sy = sample_coord((float)y + fp.y, H);
  const Bilin b = bilin_setup(sx, sy, W, H);
  const float valid = fb_valid(fp, sample_flow2(flow_check, b, W));
  // nearest texel of the propagated frame
  const int nx = (int)nearbyintf(sx), ny = (int)nearbyintf(sy);
  float wr = 0.f, wg = 0.f, wb = 0.f;
  if (nx >= 0 && nx < W && ny >= 0 && ny < H) {
    const uint2 t = prop_in[ny * W + nx];
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    wr = a.x; wg = a.y; wb = c.x;
  }
  // bilinear sample of the propagated mask (4th channel)
  float mw = 0.f;
  if (b.w00 != 0.f) mw += b.w00 * __half2float(reinterpret_cast<const __half*>(&prop_in[b.y0 * W + b.x0])[3]);
  if (b.w01 != 0.f) mw += b.w01 * __half2float(reinterpret_cast<const __half*>(&prop_in[b.y0 * W + b.x0 + 1])[3]);
  if (b.w10 != 0.f) mw += b.w10 * __half2float(reinterpret_cast<const __half*>(&prop_in[(b.y0 + 1) * W + b.x0])[3]);
  if (b.w11 != 0.f) mw += b.w11 * __half2float(reinterpret_cast<const __half*>(&prop_in[(b.y0 + 1) * W + b.x0 + 1])[3]);
  const float mv = mw > 0.1f ? 1.f : 0.f;
  const uint2 cu = cur[idx];
  const float2 c01 = __half22float2(*reinterpret_cast<const __half2*>(&cu.x));
  const float2 c23 = __half22float2(*reinterpret_cast<const __half2*>(&cu.y));
  const float mcur = c23.y;
  const float u = (mcur * valid * (1.f - mv)) > 0.1f ? 1.f : 0.f;
  const float mnew = (mcur * (1.f - valid * (1.f - mv))) > 0.1f ? 1.f : 0.f;
  const float r = u * wr + (1.f - u) * c01.x, g = u * wg + (1.f - u) * c01.y, bb = u * wb + (1.f - u) * c23.x;
  uint2 o;
  *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(r, g);
  *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(bb, mnew);
  prop_out[idx] = o;
}

// ------------------------------------------------------------------------------------------------
// The whole bidirectional propagation (2(T-1) serial steps) as ONE persistent kernel.
//
// A step only changes pixels whose current mask is set: for m_cur = 0 the blend weight u and the new mask are 0 and
// the output is the current pixel (model/propainter.py:186-196).  So the kernel works on the bounding box of the
// hole (union over the clip, computed on the device), with `bwd` and `fwd` pre-initialised to the packed input, and
// a grid-wide barrier between steps replaces 158 kernel launches.  Everything that does not depend on the
// previous step (current pixel, both flows, the fb-consistency test, the sampling positions) is fetched for step
// s+1 before the barrier of step s, so the serial chain per step is one barrier + one gather from L2.
// ------------------------------------------------------------------------------------------------
struct PropPre {        // the step-independent half of one pixel's work
  int pix;              // pixel index in the frame, -1: nothing to do
  int near;             // nearest texel index of the propagated frame or -1
  int x0, y0;
  float w00, w01, w10, w11;
  float valid;
  uint2 cur;
};

__device__ __forceinline__ PropPre prop_pre(const uint2* __restrict__ cur, const __half2* __restrict__ flow_prop,
                                            const __half2* __restrict__ flow_check, int pix, int H, int W, bool cur_is_input) {
  PropPre r;
  r.pix = pix;
  r.near = -1;
  r.cur = cur_is_input ? __ldg(&cur[pix]) : __ldcg(&cur[pix]);
  const float mcur = __half2float(reinterpret_cast<const __half*>(&r.cur)[3]);
  r.valid = 0.f; r.x0 = r.y0 = 0; r.w00 = r.w01 = r.w10 = r.w11 = 0.f;
  if (mcur == 0.f) { r.near = -2; return r; }          // -2: pass-through pixel
  const int x = pix % W, y = pix / W;
  const float2 fp = __half22float2(__ldg(&flow_prop[pix]));
  const float sx = sample_coord((float)x + fp.x, W), sy = sample_coord((float)y + fp.y, H);
  const Bilin b = bilin_setup(sx, sy, W, H);
  r.valid = fb_valid(fp, sample_flow2(flow_check, b, W));
  const int nx = (int)nearbyintf(sx), ny = (int)nearbyintf(sy);
  if (nx >= 0 && nx < W && ny >= 0 && ny < H) r.near = ny * W + nx;
  r.x0 = b.x0; r.y0 = b.y0; r.w00 = b.w00; r.w01 = b.w01; r.w10 = b.w10; r.w11 = b.w11;
  return r;
}

__device__ __forceinline__ uint2 prop_post(const PropPre& r, const uint2* prop_in, int W) {
  float wr = 0.f, wg = 0.f, wb = 0.f;
  if (r.near >= 0) {
    const uint2 t = __ldcg(&prop_in[r.near]);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    wr = a.x; wg = a.y; wb = c.x;
  }
  float mw = 0.f;
  auto mk = [&](int i) { const uint2 t = __ldcg(&prop_in[i]); return __half2float(reinterpret_cast<const __half*>(&t)[3]); };
  if (r.w00 != 0.f) mw += r.w00 * mk(r.y0 * W + r.x0);
  if (r.w01 != 0.f) mw += r.w01 * mk(r.y0 * W + r.x0 + 1);
  if (r.w10 != 0.f) mw += r.w10 * mk((r.y0 + 1) * W + r.x0);
  if (r.w11 != 0.f) mw += r.w11 * mk((r.y0 + 1) * W + r.x0 + 1);
  const float mv = mw > 0.1f ? 1.f : 0.f;
  const float2 c01 = __half22float2(*reinterpret_cast<const __half2*>(&r.cur.x));
  const float2 c23 = __half22float2(*reinterpret_cast<const __half2*>(&r.cur.y));
  const float mcur = c23.y;
  const float u = (mcur * r.valid * (1.f - mv)) > 0.1f ? 1.f : 0.f;
  const float mnew = (mcur * (1.f - r.valid * (1.f - mv))) > 0.1f ? 1.f : 0.f;
  uint2 o;
  *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(u * wr + (1.f - u) * c01.x, u * wg + (1.f - u) * c01.y);
  *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(u * wb + (1.f - u) * c23.x, mnew);
  return o;
}

// bbox = {x0, y0, x1, y1} (inclusive) of mask > 0 over all frames; initialised to {W, H, -1, -1}
__global__ void mask_bbox(const float* __restrict__ masks, long long total, int HW, int W, int* __restrict__ bbox) {
  int x0 = 1 << 30, y0 = 1 << 30, x1 = -1, y1 = -1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (masks[i] != 0.f) {
      const int p = (int)(i % HW), x = p % W, y = p / W;
      x0 = min(x0, x); y0 = min(y0, y); x1 = max(x1, x); y1 = max(y1, y);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
    x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if ((threadIdx.x & 31) == 0 && x1 >= 0) {
    atomicMin(&bbox[0], x0); atomicMin(&bbox[1], y0); atomicMax(&bbox[2], x1); atomicMax(&bbox[3], y1);
  }
}

__global__ void __launch_bounds__(256) imgprop_persistent(const uint2* __restrict__ in4, uint2* bwd, uint2* fwd,
                                                          const __half2* __restrict__ ff, const __half2* __restrict__ fbk,
                                                          int T, int H, int W, const int* __restrict__ bbox,
                                                          unsigned int* counter) {
  const int bx0 = bbox[0], by0 = bbox[1], bw = bbox[2] - bbox[0] + 1, bh = bbox[3] - bbox[1] + 1;
  if (bbox[2] < 0) return;                                  // no hole anywhere: outputs are the pre-copied inputs
  const int npix = bw * bh;
  const int nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)H * W;
  const int steps = 2 * (T - 1);
  auto bufs = [&](int s, const uint2*& cur, const uint2*& pin, uint2*& out, const __half2*& fprop, const __half2*& fchk) {
    if (s < T - 1) {          // backward pass: idx = T-2 .. 0, prop flow = forward flow, check = backward flow
      const int idx = T - 2 - s;
      cur = in4 + idx * HW; pin = bwd + (idx + 1) * HW; out = bwd + idx * HW; fprop = ff + idx * HW; fchk = fbk + idx * HW;
    } else {                  // forward pass over the backward pass's outputs: idx = 1 .. T-1
      const int idx = s - (T - 1) + 1;
      cur = bwd + idx * HW; pin = fwd + (idx - 1) * HW; out = fwd + idx * HW; fprop = fbk + (idx - 1) * HW; fchk = ff + (idx - 1) * HW;
    }
  };
  auto pixel_of = [&](int j) { return (by0 + j / bw) * W + bx0 + j % bw; };
  const uint2 *cur, *pin;
  uint2* out;
  const __half2 *fprop, *fchk;
  PropPre pre;
  pre.pix = -1;
  if (gtid < npix) {
    bufs(0, cur, pin, out, fprop, fchk);
    pre = prop_pre(cur, fprop, fchk, pixel_of(gtid), H, W, true);
  }
  for (int s = 0; s < steps; ++s) {
    bufs(s, cur, pin, out, fprop, fchk);
    const bool fwd_pass = s >= T - 1;
    // first pixel of this thread: its step-independent half was fetched before the previous barrier
    if (pre.pix >= 0) {
      if (pre.near != -2) {
        const uint2 o = prop_post(pre, pin, W);
        out[pre.pix] = o;
        if (s == T - 2) fwd[pre.pix] = o;                   // frame 0 of the forward pass is the backward result
      } else if (fwd_pass) {
        out[pre.pix] = pre.cur;                             // pass-through (backward pass: already the pre-copied input)
      } else if (s == T - 2) {
        fwd[pre.pix] = pre.cur;
      }
    }
    for (int j = gtid + nthreads; j < npix; j += nthreads) {   // holes larger than the grid: remaining pixels
      const PropPre q = prop_pre(cur, fprop, fchk, pixel_of(j), H, W, !fwd_pass);
      if (q.near != -2) {
        const uint2 o = prop_post(q, pin, W);
        out[q.pix] = o;
        if (s == T - 2) fwd[q.pix] = o;
      } else if (fwd_pass) {
        out[q.pix] = q.cur;
      } else if (s == T - 2) {
        fwd[q.pix] = q.cur;
      }
    }
    if (s + 1 < steps && gtid < npix) {
      const uint2 *c2, *p2;
      uint2* o2;
      const __half2 *f2, *k2;
      bufs(s + 1, c2, p2, o2, f2, k2);
      // in the forward pass `cur` is a backward-pass output of THIS thread (same pixel mapping), complete by now
      pre = prop_pre(c2, f2, k2, pixel_of(gtid), H, W, s + 1 < T - 1);
    }
    ppx::grid_barrier(counter, (unsigned)(s + 1) * gridDim.x);
  }
}

// frames [T,3,H,W] f32 (already multiplied by (1-mask) here) + masks -> [T][H][W][4] fp16
__global__ void imgprop_pack(const float* __restrict__ frames, const float* __restrict__ masks,
                             uint2* __restrict__ dst, long long HW, long long total) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long t = idx / HW, p = idx - t * HW;
  const float m = masks[idx];
  const float* f = frames + t * 3 * HW + p;
  const float k = 1.f - m;
  uint2 o;
  *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(f[0] * k, f[HW] * k);
  *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(f[2 * HW] * k, m);
  dst[idx] = o;
}

// updated = frames*(1-m) + prop*m ; updated mask = propagated mask   (propainter_inference.py:213-219)
__global__ void imgprop_finish(const uint2* __restrict__ prop, const float* __restrict__ frames,
                               const float* __restrict__ masks, float* __restrict__ uf, float* __restrict__ um,
                               long long HW, long long total) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long t = idx / HW, p = idx - t * HW;
  const uint2 v = prop[idx];
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  const float m = masks[idx], k = 1.f - m;
  const float* f = frames + t * 3 * HW + p;
  float* o = uf + t * 3 * HW + p;
  o[0] = f[0] * k + a.x * m;
  o[HW] = f[HW] * k + a.y * m;
  o[2 * HW] = f[2 * HW] * k + c.x * m;
  um[idx] = c.y;
}

// [n,2,H,W] f32 -> [n][H][W][2] fp16
__global__ void flow_to_nhwc2(const float* __restrict__ src, __half2* __restrict__ dst, long long HW,
                              long long total) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / HW, p = idx - n * HW;
  dst[idx] = __floats2half2_rn(src[(n * 2) * HW + p], src[(n * 2 + 1) * HW + p]);
}

// ------------------------------------------------------------------------------------------------
// Flow completion input: cat(flow * (1 - mask), mask) per frame (recurrent_flow_completion.py:361-366, 320-323),
// optionally time-reversed (backward flows are flipped before the network, :375-376).  -> [T][H][W][8] fp16.
// ------------------------------------------------------------------------------------------------
__global__ void rfc_pack_input(const float* __restrict__ flows, const float* __restrict__ masks,
                               uint4* __restrict__ dst, int T, long long HW, int reverse, long long dst_tstride) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)T * HW) return;
  const long long t = idx / HW, p = idx - t * HW;
  const long long ts = reverse ? (T - 1 - t) : t;
  const float m = masks[ts * HW + p];
  const float k = 1.f - m;
  uint4 o = make_uint4(0, 0, 0, 0);
  *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(flows[(ts * 2) * HW + p] * k, flows[(ts * 2 + 1) * HW + p] * k);
  *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(m, 0.f);
  dst[t * dst_tstride + p] = o;
}

// combine_flow (recurrent_flow_completion.py:389-400): out = pred*m + gt*(1-m), un-reversing time.
__global__ void rfc_combine(const __half* __restrict__ pred, int pred_cs, const float* __restrict__ gt,
                            const float* __restrict__ masks, float* __restrict__ out, int T, long long HW,
                            int reverse, long long pred_tstride) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)T * HW) return;
  const long long t = idx / HW, p = idx - t * HW;
  const long long ts = reverse ? (T - 1 - t) : t;  // network time index holding frame t
  const __half* pr = pred + (ts * pred_tstride + p) * pred_cs;
  const float m = masks[idx], k = 1.f - m;
  out[(t * 2) * HW + p] = __half2float(pr[0]) * m + gt[(t * 2) * HW + p] * k;
  out[(t * 2 + 1) * HW + p] = __half2float(pr[1]) * m + gt[(t * 2 + 1) * HW + p] * k;
}

// ------------------------------------------------------------------------------------------------
// Modulated deformable sampling (torchvision.ops.deform_conv2d im2col stage; call sites
// recurrent_flow_completion.py:44-53 and propainter.py:73-82), 3x3, stride 1, pad 1, dil 1, 16 offset groups.
// offs row = raw output of conv_offset[-1]: channels [0,288) -> offsets, (g*9+k)*2+{0:dy,1:dx}, passed through
// max_mag*tanh (+ flow (dy,dx) for the feature path); channels [288,432) -> sigmoid modulation, g*9+k.
// cols[m][k*C + c] = mask * bilinear(x[c], y-1+ky+dy, x-1+kx+dx); zero outside (h<=-1 || h>=H ...).
// One thread per (pixel, tap, group): the 4 bilinear weights are computed once and applied to the
// group's C/16 contiguous channels with 16-byte loads.
// ------------------------------------------------------------------------------------------------
template <int CPG>  // channels per offset group: 8 or 16
__global__ void dcn_sample(const PPDcnArgs a) {
  // grid = (chunks of one image's H*W*144 (pixel, group, tap) items, images): 32-bit index math
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)(a.H * a.W * 144)) return;
  dcn_sample_item<CPG, false>(a, idx, blockIdx.y);
}

// ------------------------------------------------------------------------------------------------
// Learnable feature propagation, per step (model/propainter.py:157-176): fb check of the 1/4-res flows,
// bilinear warp of the propagated feature, and assembly of the DCN condition
//   cond = cat(cur[C], warped[C], flow[2], valid[1], mask_cur[2])  -> [H][W][2C+8] (3 pad channels = 0)
// One thread per (pixel, 8-channel vector of C); vector 0 also writes the 5 scalar channels.
// ------------------------------------------------------------------------------------------------
__global__ void featprop_cond(const __half* __restrict__ cur, int cur_cs, const __half* __restrict__ prop,
                              int prop_cs, const __half2* __restrict__ flow_prop,
                              const __half2* __restrict__ flow_check, const __half* __restrict__ mask2,
                              int mask_cs, __half* __restrict__ cond, int cond_cs, int N, int H, int W, int C) {
  const int C8 = C / 8;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)N * H * W * C8) return;
  const int c8 = idx % C8;
  const long long pg = idx / C8;          // pixel over all N images
  const int HWi = H * W;
  const int img = pg / HWi;
  const int p = pg - (long long)img * HWi;  // pixel inside the image
  const int x = p % W, y = p / W;
  // re-base every per-image tensor
  cur += (long long)img * HWi * cur_cs;
  prop += (long long)img * HWi * prop_cs;
  flow_prop += (long long)img * HWi;
  flow_check += (long long)img * HWi;
  mask2 += (long long)img * HWi * mask_cs;
  cond += (long long)img * HWi * cond_cs;
  const float2 fp = __half22float2(flow_prop[p]);
  const float sx = sample_coord((float)x + fp.x, W), sy = sample_coord((float)y + fp.y, H);
  const Bilin b = bilin_setup(sx, sy, W, H);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const float ws[4] = {b.w00, b.w01, b.w10, b.w11};
#pragma unroll
  for (int corner = 0; corner < 4; ++corner) {
    if (ws[corner] == 0.f) continue;
    const int yy = b.y0 + (corner >> 1), xx = b.x0 + (corner & 1);
    const uint4 q = *reinterpret_cast<const uint4*>(prop + ((long long)yy * W + xx) * prop_cs + c8 * 8);
    const __half2* hq = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(hq[e]);
      acc[2 * e] += ws[corner] * f.x;
      acc[2 * e + 1] += ws[corner] * f.y;
    }
  }
  __half* cp = cond + (long long)p * cond_cs;
  *reinterpret_cast<uint4*>(cp + c8 * 8) = *reinterpret_cast<const uint4*>(cur + (long long)p * cur_cs + c8 * 8);
  __align__(16) __half2 h[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
  *reinterpret_cast<uint4*>(cp + C + c8 * 8) = *reinterpret_cast<uint4*>(h);
  if (c8 == 0) {
    const float valid = fb_valid(fp, sample_flow2(flow_check, b, W));
    const float2 mk = __half22float2(*reinterpret_cast<const __half2*>(mask2 + (long long)p * mask_cs));
    __align__(16) __half2 s[4];
    s[0] = __floats2half2_rn(fp.x, fp.y);
    s[1] = __floats2half2_rn(valid, mk.x);
    s[2] = __floats2half2_rn(mk.y, 0.f);
    s[3] = __floats2half2_rn(0.f, 0.f);
    *reinterpret_cast<uint4*>(cp + 2 * C) = *reinterpret_cast<uint4*>(s);
  }
}

// F.interpolate(scale_factor=1/4, bilinear, align_corners=False) == mean of the centre 2x2 of each 4x4 block;
// the reference then divides the flow by 4 (propainter.py:389-406).  [n,2,H,W] f32 -> [n][H/4][W/4][2] fp16.
__global__ void downsample_flow4(const float* __restrict__ src, __half2* __restrict__ dst, int n, int H, int W) {
  const int h = H / 4, w = W / 4;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)n * h * w) return;
  const int x = idx % w;
  long long t = idx / w;
  const int y = t % h;
  const int i = t / h;
  float v[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const float* s = src + ((long long)(i * 2 + c) * H + 4 * y + 1) * W + 4 * x + 1;
    v[c] = 0.25f * (0.25f * (s[0] + s[1] + s[W] + s[W + 1]));
  }
  dst[idx] = __floats2half2_rn(v[0], v[1]);
}

// F.interpolate(scale_factor=1/4, 'nearest') picks source index 4*i.  [n,1,H,W] f32 -> fp16 slice.
__global__ void downsample_mask4(const float* __restrict__ src, __half* __restrict__ dst, int dst_cs, int dst_co,
                                 int n, int H, int W) {
  const int h = H / 4, w = W / 4;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)n * h * w) return;
  const int x = idx % w;
  long long t = idx / w;
  const int y = t % h;
  const int i = t / h;
  dst[idx * dst_cs + dst_co] = __float2half_rn(src[((long long)i * H + 4 * y) * W + 4 * x]);
}

}  // namespace

int pp_k_imgprop_step(const __half* cur, const __half* prop_in, __half* prop_out, const __half* flow_prop,
                      const __half* flow_check, int H, int W, cudaStream_t st) {
  imgprop_step<<<nblocks((long long)H * W), TPB, 0, st>>>(
      reinterpret_cast<const uint2*>(cur), reinterpret_cast<const uint2*>(prop_in), reinterpret_cast<uint2*>(prop_out),
      reinterpret_cast<const __half2*>(flow_prop), reinterpret_cast<const __half2*>(flow_check), H, W);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

// bwd / fwd must hold copies of in4; scratch = 8 ints (bbox[4], barrier counter, pad)
int pp_k_imgprop_run(const __half* in4, __half* bwd, __half* fwd, const __half* ff, const __half* fbk,
                     const float* masks, int T, int H, int W, int* scratch, cudaStream_t st) {
  static int grid_max = 0;
  if (grid_max == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    PP_CUDA_CHECK(cudaGetDevice(&dev));
    PP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, imgprop_persistent, 256, 0));
    PP_REQUIRE(per_sm >= 1, "imgprop: persistent kernel does not fit an SM");
    grid_max = sms * (per_sm < 2 ? per_sm : 2);
  }
  const int init[8] = {W, H, -1, -1, 0, 0, 0, 0};
  PP_CUDA_CHECK(cudaMemcpyAsync(scratch, init, sizeof(init), cudaMemcpyHostToDevice, st));
  const long long total = (long long)T * H * W;
  const long long want = (total + 256 * 16 - 1) / (256 * 16);
  mask_bbox<<<(int)(want < 1184 ? want : 1184), 256, 0, st>>>(masks, total, H * W, W, scratch);
  PP_CUDA_CHECK(cudaGetLastError());
  const uint2* a0 = reinterpret_cast<const uint2*>(in4);
  uint2* a1 = reinterpret_cast<uint2*>(bwd);
  uint2* a2 = reinterpret_cast<uint2*>(fwd);
  const __half2* a3 = reinterpret_cast<const __half2*>(ff);
  const __half2* a4 = reinterpret_cast<const __half2*>(fbk);
  const int* a8 = scratch;
  unsigned int* a9 = reinterpret_cast<unsigned int*>(scratch + 4);
  void* args[] = {&a0, &a1, &a2, &a3, &a4, &T, &H, &W, &a8, &a9};
  PP_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)imgprop_persistent, dim3(grid_max), dim3(256), args, 0, st));
  return PP_OK;
}

int pp_k_imgprop_pack(const float* frames, const float* masks, __half* dst, int T, int H, int W, cudaStream_t st) {
  const long long HW = (long long)H * W, total = HW * T;
  imgprop_pack<<<nblocks(total), TPB, 0, st>>>(frames, masks, reinterpret_cast<uint2*>(dst), HW, total);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_imgprop_finish(const __half* prop, const float* frames, const float* masks, float* upd_frames,
                        float* upd_masks, int T, int H, int W, cudaStream_t st) {
  const long long HW = (long long)H * W, total = HW * T;
  imgprop_finish<<<nblocks(total), TPB, 0, st>>>(reinterpret_cast<const uint2*>(prop), frames, masks, upd_frames,
                                                 upd_masks, HW, total);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_flow_to_nhwc2(const float* src, __half* dst, int n, int H, int W, cudaStream_t st) {
  const long long HW = (long long)H * W, total = HW * n;
  if (total == 0) return PP_OK;
  flow_to_nhwc2<<<nblocks(total), TPB, 0, st>>>(src, reinterpret_cast<__half2*>(dst), HW, total);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_rfc_pack_input(const float* flows, const float* masks, __half* dst, long long dst_tstride_pix, int T, int H,
                        int W, int reverse_time, cudaStream_t st) {
  const long long HW = (long long)H * W;
  rfc_pack_input<<<nblocks(HW * T), TPB, 0, st>>>(flows, masks, reinterpret_cast<uint4*>(dst), T, HW, reverse_time,
                                                  dst_tstride_pix);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_rfc_combine(const __half* pred, int pred_cs, long long pred_tstride_pix, const float* gt, const float* masks,
                     float* out, int T, int H, int W, int reverse_time, cudaStream_t st) {
  const long long HW = (long long)H * W;
  rfc_combine<<<nblocks(HW * T), TPB, 0, st>>>(pred, pred_cs, gt, masks, out, T, HW, reverse_time, pred_tstride_pix);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_dcn_sample(const __half* x0, int x0_cs, int x0_co, int C0, const __half* x1, int x1_cs, int x1_co, int C1,
                    const __half* offs, int offs_cs, const __half* flow, int flow_cs, int flow_co, float max_mag,
                    __half* cols, int N, int H, int W, cudaStream_t st) {
  const int C = C0 + C1;
  PP_REQUIRE(C == 128 || C == 256, "dcn_sample: C=%d must be 128 or 256 (16 offset groups)", C);
  PP_REQUIRE(C0 % 16 == 0, "dcn_sample: C0=%d", C0);
  if ((long long)N * H * W == 0) return PP_OK;
  PP_REQUIRE(N <= 65535 && (long long)H * W * 144 < (1LL << 31), "dcn_sample: %d images of %dx%d exceed the grid limits", N, W, H);
  PPDcnArgs a;
  a.x0 = x0; a.x0_cs = x0_cs; a.x0_co = x0_co; a.C0 = C0;
  a.x1 = x1; a.x1_cs = x1_cs; a.x1_co = x1_co;
  a.offs = offs; a.offs_cs = offs_cs;
  a.flow = flow; a.flow_cs = flow_cs; a.flow_co = flow_co;
  a.max_mag = max_mag; a.cols = cols; a.C = C; a.N = N; a.H = H; a.W = W;
  if (pp_prog_recording()) return pp_prog_record_dcn(a);     // multi-layer program (conv_halo.cu): runs inside it
  {
    // opt-in (PP_DCN_TILED=1): source tiles staged in shared memory by TMA (dcn_tiled.cu).  Bit-identical, but measured
    // no faster than the plain sampler (35 vs 36 us at the flow-completion size, 97 vs 102 us at the generator's): the
    // sampler is bound by L1 wavefronts of its uncoalesced accesses, the offsets / stores keep that cost
    const char* s = getenv("PP_DCN_TILED");       // read per call: tests compare both samplers in one process
    if (s != nullptr && atoi(s) != 0) {
      int handled = 0;
      PP_TRY(pp_k_dcn_sample_tiled(a, 3, st, &handled));
      if (handled) return PP_OK;
    }
  }
  return pp_k_dcn_sample_plain(a, st);
}

int pp_k_dcn_sample_plain(const PPDcnArgs& a, cudaStream_t st) {
  PP_REQUIRE(a.C == 128 || a.C == 256, "dcn_sample: C=%d must be 128 or 256 (16 offset groups)", a.C);
  if ((long long)a.N * a.H * a.W == 0) return PP_OK;
  PP_REQUIRE(a.N <= 65535 && (long long)a.H * a.W * 144 < (1LL << 31), "dcn_sample: %d images of %dx%d exceed the grid limits",
             a.N, a.W, a.H);
  const dim3 grid(pp_ceil_div(a.H * a.W * 144, TPB), a.N);
  if (a.C == 128) dcn_sample<8><<<grid, TPB, 0, st>>>(a);
  else dcn_sample<16><<<grid, TPB, 0, st>>>(a);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_featprop_cond(const __half* cur, int cur_cs, const __half* prop, int prop_cs, const __half* flow_prop,
                       const __half* flow_check, const __half* mask2, int mask_cs, __half* cond, int cond_cs, int N,
                       int H, int W, int C, cudaStream_t st) {
  PP_REQUIRE(C % 8 == 0 && cond_cs >= 2 * C + 8, "featprop_cond: bad channel counts");
  featprop_cond<<<nblocks((long long)N * H * W * (C / 8)), TPB, 0, st>>>(
      cur, cur_cs, prop, prop_cs, reinterpret_cast<const __half2*>(flow_prop),
      reinterpret_cast<const __half2*>(flow_check), mask2, mask_cs, cond, cond_cs, N, H, W, C);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_downsample_flow4(const float* flow, __half* dst, int n, int H, int W, cudaStream_t st) {
  const long long total = (long long)n * (H / 4) * (W / 4);
  if (total == 0) return PP_OK;
  downsample_flow4<<<nblocks(total), TPB, 0, st>>>(flow, reinterpret_cast<__half2*>(dst), n, H, W);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

int pp_k_downsample_mask4(const float* m, __half* dst, int dst_cs, int dst_co, int n, int H, int W, cudaStream_t st) {
  const long long total = (long long)n * (H / 4) * (W / 4);
  if (total == 0) return PP_OK;
  downsample_mask4<<<nblocks(total), TPB, 0, st>>>(m, dst, dst_cs, dst_co, n, H, W);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}
