// Modulated deformable sampling with the source tile staged in shared memory by TMA.
//
// The plain sampler (dcn_sample in kernels_prop.cu) reads, per (pixel, offset group, tap), four bilinear corners of
// C/16 channels straight from L2: 128 B of 16-32-byte gathers per item, 18 KB per output pixel -- it is bound by L2
// sector bandwidth (ncu: long scoreboard 68 %, 1.4 TB/s algorithmic).  But the learned offsets are bounded
// (max_mag * tanh, recurrent_flow_completion.py:40, propainter.py:66) so every sample of a TH x TW tile of output
// pixels lies inside the tile grown by R = ceil(max_mag) + 2 pixels (+ a margin for the added flow in the feature
// path).  Here one CTA owns (tile, quad of 4 offset groups): ONE TMA box load lands [BH][BW][4*C/16] fp16 in shared
// memory (zero-filled outside the image = the sampler's zero padding) and all (pixel, group, tap) items of the quad
// sample from it.  L2 traffic per output pixel drops from 18 KB to (BH*BW / (TH*TW)) * 2C bytes (1.8-3.9 KB).  A sample whose corners leave the staged box (possible only through a
// large flow in the feature path) falls back to global loads, so the result never depends on R.
#include <cuda.h>

#include "dcn_sample.cuh"
#include "kernels.cuh"

namespace {

constexpr int NT = 256;

struct TiledParams {
  PPDcnArgs a;
  CUtensorMap tmap[2];     // x0, x1
  int TH, TW, R, BH, BW;   // tile, halo, box
  int tiles_x, tiles_y;
  int stage_bytes;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(ppx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// One CTA = (tile, quad of 4 offset groups, image): ONE TMA box load of the quad's 4*CPG channels, then all 256
// threads work through the tile's (pixel, group-in-quad, tap) items from shared memory.  No in-CTA pipeline: latency is
// hidden by 3-4 resident CTAs per SM and 4x more CTAs than tiles.
template <int CPG>
__global__ void __launch_bounds__(NT, 2) dcn_sample_tiled(const __grid_constant__ TiledParams P) {
  using namespace ppx;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint8_t* box = smem + 128;
  const PPDcnArgs& a = P.a;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x >> 2, quad = blockIdx.x & 3, n = blockIdx.y;
  const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
  const int y0 = ty * P.TH, x0 = tx * P.TW;              // tile origin
  const int by0 = y0 - P.R, bx0 = x0 - P.R;              // box origin (may be negative: zero-filled)
  const int H = a.H, W = a.W;
  constexpr int QC = 4 * CPG;                            // channels of a quad
  if (tid == 0) {
    mbar_init(full, 1);
    mbar_fence_init();
    const int c = quad * QC;                              // first channel inside cat(x0, x1)
    const int q = c < a.C0 ? 0 : 1;
    mbar_arrive_expect_tx(full, (uint32_t)(P.BH * P.BW * QC * 2));
    tma_load_4d(smem_u32(box), &P.tmap[q], q ? c - a.C0 : c, bx0, by0, n, full);
  }
  __syncthreads();
  mbar_wait(full, 0);
  const int items = P.TH * P.TW * 36;                    // (pixel, group in quad, tap)
  constexpr int U = 32 / CPG;                            // independent items per thread and pass (see dcn_sample.cuh)
  for (int base = tid; base < items; base += NT * U) {
    PPDcnItem item[U];
    int gq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = base + u * NT;
      item[u].m = -1;
      gq[u] = 0;
      if (it < items) {
        const int gk4 = it % 36, pp = it / 36;
        gq[u] = gk4 / 9;
        const int py_i = pp / P.TW, px_i = pp - py_i * P.TW;
        const int y = y0 + py_i, x = x0 + px_i;
        if (y < H && x < W) item[u] = dcn_item_setup<false>(a, n, y * W + x, quad * 4 + gq[u], gk4 - gq[u] * 9);
      }
    }
    uint4 q[U][4][CPG / 8];
    float w[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) dcn_item_gather<CPG, false>(a, item[u], box, by0, bx0, P.BH, P.BW, QC, gq[u], q[u], w[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) dcn_item_store<CPG>(a, item[u], q[u], w[u]);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult qr;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

// 1 = handled (kernel launched), 0 = not applicable (caller uses the plain sampler)
int pp_k_dcn_sample_tiled(const PPDcnArgs& a, int flow_margin, cudaStream_t st, int* handled) {
  *handled = 0;
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return PP_OK;
  const int cpg = a.C / 16;
  if ((cpg != 8 && cpg != 16) || a.C0 % (4 * cpg) != 0) return PP_OK;   // a quad of groups never straddles x0 | x1
  if ((a.x0_cs % 8) || (a.x0_co % 8) || (a.x1 != nullptr && ((a.x1_cs % 8) || (a.x1_co % 8)))) return PP_OK;
  TiledParams P;
  P.a = a;
  const long long px = (long long)a.N * a.H * a.W;
  P.R = (int)ceilf(a.max_mag) + 2 + (a.flow != nullptr ? flow_margin : 0);
  // 16x16 tiles when their box fits 96 KB and there are at least two waves of CTAs, else 8x8
  const bool big = (size_t)(16 + 2 * P.R) * (16 + 2 * P.R) * 4 * cpg * 2 + 256 <= 96 * 1024 &&
                   (long long)pp_ceil_div(a.H, 16) * pp_ceil_div(a.W, 16) * a.N * 4 >= 2 * 148;
  P.TH = P.TW = big ? 16 : 8;
  P.BH = P.TH + 2 * P.R;
  P.BW = P.TW + 2 * P.R;
  P.tiles_x = pp_ceil_div(a.W, P.TW);
  P.tiles_y = pp_ceil_div(a.H, P.TH);
  P.stage_bytes = pp_ceil_div(P.BH * P.BW * 4 * cpg * 2, 128) * 128;
  if (P.BH > 256 || P.BW > 256 || a.N > 65535 || px == 0) return PP_OK;
  const __half* base[2] = {a.x0 + a.x0_co, a.x1 != nullptr ? a.x1 + a.x1_co : nullptr};
  const int cs[2] = {a.x0_cs, a.x1_cs}, cn[2] = {a.C0, a.C - a.C0};
  for (int q = 0; q < 2; ++q) {
    if (cn[q] == 0 || base[q] == nullptr) { memset(&P.tmap[q], 0, sizeof(CUtensorMap)); continue; }
    if ((reinterpret_cast<uintptr_t>(base[q]) & 15) != 0) return PP_OK;
    cuuint64_t dims[4] = {(cuuint64_t)cn[q], (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N};
    cuuint64_t strides[3] = {(cuuint64_t)cs[q] * 2, (cuuint64_t)a.W * cs[q] * 2, (cuuint64_t)a.H * a.W * cs[q] * 2};
    cuuint32_t box[4] = {(cuuint32_t)(4 * cpg), (cuuint32_t)P.BW, (cuuint32_t)P.BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = enc(&P.tmap[q], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base[q]), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PP_REQUIRE(r == CUDA_SUCCESS, "dcn_sample_tiled: cuTensorMapEncodeTiled failed (%d)", (int)r);
  }
  const size_t smem = 128 + (size_t)P.stage_bytes;
  static bool attr = false;
  if (!attr) {
    PP_CUDA_CHECK(cudaFuncSetAttribute(dcn_sample_tiled<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    PP_CUDA_CHECK(cudaFuncSetAttribute(dcn_sample_tiled<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  if (smem > 96 * 1024) return PP_OK;
  const dim3 grid(P.tiles_x * P.tiles_y * 4, a.N);
  if (cpg == 8) dcn_sample_tiled<8><<<grid, NT, smem, st>>>(P);
  else dcn_sample_tiled<16><<<grid, NT, smem, st>>>(P);
  PP_CUDA_CHECK(cudaGetLastError());
  *handled = 1;
  return PP_OK;
}
