// Epilogue shared by the tcgen05 conv kernels: 16 consecutive output channels of one output pixel
// (accumulators already in registers) -> bias / activation / residual / GRU gate fusions -> global memory.
#pragma once
#include "conv_igemm.cuh"

namespace ppconv {

template <int ACT>
__device__ __forceinline__ void act16_t(float (&v)[16], float slope) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (ACT == PP_ACT_RELU) v[i] = fmaxf(v[i], 0.f);
    else if (ACT == PP_ACT_LRELU) v[i] = v[i] > 0.f ? v[i] : v[i] * slope;
    else if (ACT == PP_ACT_SIGMOID) v[i] = ppx::sigmoidf_(v[i]);
    else if (ACT == PP_ACT_TANH) v[i] = tanhf(v[i]);
    else if (ACT == PP_ACT_GELU) v[i] = ppx::gelu_erf(v[i]);
  }
}
// one (uniform) branch per 16 values instead of one per value
__device__ __forceinline__ void act16(float (&v)[16], int act, float slope) {
  switch (act) {
    case PP_ACT_RELU: act16_t<PP_ACT_RELU>(v, slope); break;
    case PP_ACT_LRELU: act16_t<PP_ACT_LRELU>(v, slope); break;
    case PP_ACT_SIGMOID: act16_t<PP_ACT_SIGMOID>(v, slope); break;
    case PP_ACT_TANH: act16_t<PP_ACT_TANH>(v, slope); break;
    case PP_ACT_GELU: act16_t<PP_ACT_GELU>(v, slope); break;
    default: break;
  }
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): 16 fp16 channels of one pixel in ONE request, so a warp's
// store touches each of its 32 rows' sectors once instead of twice.
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// 16 consecutive fp16 values <-> registers.  `vec` (uniform per launch, checked on the host) says that full
// runs are 16-byte aligned, so they move as 2 x 16-byte accesses; partial runs take the scalar tail.
__device__ __forceinline__ void load16(const __half* src, int nvalid, bool vec, float (&r)[16]) {
  if (vec && nvalid == 16) {
    const uint4 a = reinterpret_cast<const uint4*>(src)[0], b = reinterpret_cast<const uint4*>(src)[1];
    const __half2* ha = reinterpret_cast<const __half2*>(&a);
    const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 fa = __half22float2(ha[i]), fb = __half22float2(hb[i]);
      r[2 * i] = fa.x; r[2 * i + 1] = fa.y; r[8 + 2 * i] = fb.x; r[8 + 2 * i + 1] = fb.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = i < nvalid ? __half2float(src[i]) : 0.f;
  }
}
__device__ __forceinline__ void store16(__half* dst, int nvalid, bool vec, const float (&v)[16], bool v32 = false) {
  if (vec && nvalid == 16) {
    __align__(16) __half2 h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    if (v32) {
      stg256(dst, reinterpret_cast<uint4*>(h)[0], reinterpret_cast<uint4*>(h)[1]);
      return;
    }
    reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<uint4*>(h)[0];
    reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<uint4*>(h)[1];
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) dst[i] = __float2half_rn(v[i]);
  }
}

__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float (&r)[16]) {
  const __half2* ha = reinterpret_cast<const __half2*>(&a);
  const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __half22float2(ha[i]), fb = __half22float2(hb[i]);
    r[2 * i] = fa.x; r[2 * i + 1] = fa.y; r[8 + 2 * i] = fb.x; r[8 + 2 * i + 1] = fb.y;
  }
}

// Epilogue operands that do not depend on the accumulator (residual / GRU h and z), fetched while the TMEM read of
// the same 16 columns is still in flight so the two latencies overlap instead of adding up.
struct EpiAux {
  uint4 a0[2], a1[2];
  bool have;
};
__device__ __forceinline__ void conv_epilogue_prefetch16(const PPConvParams& p, long long mrow, int ng0, int epi, bool vec,
                                                         EpiAux& x) {
  x.have = false;
  if (!vec || p.Cout_g - ng0 < 16) return;
  const __half* s0 = nullptr;
  const __half* s1 = nullptr;
  if (epi == PP_EPI_STD) {
    if (p.aux0 != nullptr) s0 = p.aux0 + mrow * p.aux0_cstride + p.aux0_coff + ng0;
  } else if (epi == PP_EPI_GRU_ZR) {
    const int half_c = p.Cout_g >> 1;
    if (ng0 >= half_c) s0 = p.aux0 + mrow * p.aux0_cstride + p.aux0_coff + (ng0 - half_c);
  } else {
    s0 = p.aux0 + mrow * p.aux0_cstride + p.aux0_coff + ng0;
    s1 = p.aux1 + mrow * p.aux1_cstride + p.aux1_coff + ng0;
  }
  if (s0 == nullptr) return;
  x.have = true;
  if (p.vec32_ok) {
    ldg256(s0, x.a0[0], x.a0[1]);
    if (s1 != nullptr) ldg256(s1, x.a1[0], x.a1[1]);
    return;
  }
  x.a0[0] = reinterpret_cast<const uint4*>(s0)[0];
  x.a0[1] = reinterpret_cast<const uint4*>(s0)[1];
  if (s1 != nullptr) {
    x.a1[0] = reinterpret_cast<const uint4*>(s1)[0];
    x.a1[1] = reinterpret_cast<const uint4*>(s1)[1];
  }
}

// `raw`: 16 fp32 accumulators (TMEM columns ng0-n0 .. +15) of output pixel `mrow` (flattened N*OH*OW index),
// group g, first channel ng0 (within the group; ng0 < Cout_g).  `epi`/`vec` are launch-uniform.
// sm0 / sm1 (STD epilogue, fp16 output only): when non-null the 16 results go to these two 16-byte shared-memory slots
// (a staging tile that a TMA store writes out) instead of global memory.
__device__ __forceinline__ void conv_epilogue16(const PPConvParams& p, const uint32_t (&raw)[16], long long mrow, int g,
                                                int ng0, int epi, bool vec, const EpiAux* pre = nullptr,
                                                uint4* sm0 = nullptr, uint4* sm1 = nullptr) {
    const int nvalid = min(16, p.Cout_g - ng0);
    const bool v32 = p.vec32_ok != 0;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
    if (p.bias != nullptr) {
      const float* bp = p.bias + g * p.Cout_g + ng0;
      if (vec && nvalid == 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + i);
          v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < nvalid) v[i] += __ldg(bp + i);
      }
    }
    if (epi == PP_EPI_STD) {
      act16(v, p.act1, p.slope);
      if (p.scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= p.scale;
      }
      if (p.aux0 != nullptr) {
        float r[16];
        if (pre != nullptr && pre->have) unpack16(pre->a0[0], pre->a0[1], r);
        else load16(p.aux0 + mrow * p.aux0_cstride + p.aux0_coff + ng0, nvalid, vec, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += r[i];
      }
      act16(v, p.act2, p.slope);
      const long long o = mrow * p.out_cstride + p.out_coff + (long long)g * p.out_gstep + ng0;
      if (p.out_fp32) {
        float* dst = reinterpret_cast<float*>(p.out) + o;
        if (vec && nvalid == 16) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < nvalid) dst[i] = v[i];
        }
      } else if (sm0 != nullptr) {
        __align__(16) __half2 h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        *sm0 = reinterpret_cast<uint4*>(h)[0];
        *sm1 = reinterpret_cast<uint4*>(h)[1];
      } else {
        store16(reinterpret_cast<__half*>(p.out) + o, nvalid, vec, v, v32);
      }
    } else if (epi == PP_EPI_GRU_ZR) {
      const int half_c = p.Cout_g >> 1;
      act16_t<PP_ACT_SIGMOID>(v, 0.f);
      if (ng0 < half_c) {
        store16(reinterpret_cast<__half*>(p.out) + mrow * p.out_cstride + p.out_coff + ng0, nvalid, vec, v, v32);
      } else {
        const int c = ng0 - half_c;
        float h[16];
        if (pre != nullptr && pre->have) unpack16(pre->a0[0], pre->a0[1], h);
        else load16(p.aux0 + mrow * p.aux0_cstride + p.aux0_coff + c, nvalid, vec, h);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= h[i];
        store16(p.out2 + mrow * p.out2_cstride + p.out2_coff + c, nvalid, vec, v, v32);
      }
    } else {  // PP_EPI_GRU_H
      float h[16], z[16];
      if (pre != nullptr && pre->have) {
        unpack16(pre->a0[0], pre->a0[1], h);
        unpack16(pre->a1[0], pre->a1[1], z);
      } else {
        load16(p.aux0 + mrow * p.aux0_cstride + p.aux0_coff + ng0, nvalid, vec, h);
        load16(p.aux1 + mrow * p.aux1_cstride + p.aux1_coff + ng0, nvalid, vec, z);
      }
      act16_t<PP_ACT_TANH>(v, 0.f);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (1.f - z[i]) * h[i] + z[i] * v[i];
      store16(reinterpret_cast<__half*>(p.out) + mrow * p.out_cstride + p.out_coff + ng0, nvalid, vec, v, v32);
    }
}

}  // namespace ppconv
