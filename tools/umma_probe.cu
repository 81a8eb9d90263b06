// Probe (B200 only): can a tcgen05 K-major SWIZZLE_128B A descriptor address 8-row groups that start at any
// 128-byte multiple (not 1024-aligned) with an arbitrary stride between groups (SBO)?  That is what a
// shared-memory *halo tile* convolution needs: the A operand of filter tap (ky,kx) is a shifted view of one
// input patch.  Part 1 fills the patch with generic stores (address-based swizzle); part 2 loads it with a
// 4-D TMA tensor-map box (SWIZZLE_128B, negative coordinates -> zero fill).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu -I comfyui_propainter_nodes_b200/csrc
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

void pp_set_error(const char*, ...) {}
#include "pp_common.cuh"

using namespace ppx;

constexpr int NPIX = 512;   // pixels of the patch (128 B each: 64 fp16 channels)
constexpr int BN = 64;

__device__ __forceinline__ uint64_t desc_a(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// mode 0: patch[p][64] copied from global `patch` with generic stores.  mode 1: TMA box load.
__global__ void __launch_bounds__(128) probe(const __half* patch, const __half* bmat, float* out, int start_pix,
                                             int sbo_pix, int base_off, int mode, const __grid_constant__ CUtensorMap tmap,
                                             int cx, int cy, int box_bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sA = smem;                       // NPIX * 128
  uint8_t* sB = smem + NPIX * 128;          // BN * 128
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + BN * 128);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  __syncthreads();
  if (mode == 0) {
    for (int i = tid; i < NPIX * 8; i += 128) {
      const int p = i >> 3, ch = i & 7;
      const uint4 v = reinterpret_cast<const uint4*>(patch)[i];
      *reinterpret_cast<uint4*>(sA + p * 128 + ((ch ^ (p & 7)) << 4)) = v;
    }
  } else {
    for (int i = tid; i < NPIX * 8; i += 128) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0x7e007e00u, 0x7e007e00u, 0x7e007e00u, 0x7e007e00u);  // NaN fill
    __syncthreads();
    fence_proxy_async();
    if (tid == 0) {
      mbar_arrive_expect_tx(bar2, (uint32_t)box_bytes);
      tma_load_4d(smem_u32(sA), &tmap, 0, cx, cy, 0, bar2);
    }
    mbar_wait(bar2, 0);
  }
  for (int i = tid; i < BN * 8; i += 128) {
    const int r = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(sB + r * 128 + ((ch ^ (r & 7)) << 4)) = reinterpret_cast<const uint4*>(bmat)[i];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(128, BN);
    const uint32_t a0 = smem_u32(sA) + start_pix * 128;
    for (int k = 0; k < 4; ++k)
      umma_f16(tb, desc_a(a0 + 32 * k, sbo_pix * 128, base_off), umma_desc_sw128_kmajor(smem_u32(sB) + 32 * k), idesc, k != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c = 0; c < BN; c += 16) {
    uint32_t rawv[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c, rawv);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[tid * BN + c + i] = __uint_as_float(rawv[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 64);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  std::vector<__half> hp(NPIX * 64), hb(BN * 64);
  std::vector<float> fp(NPIX * 64), fb(BN * 64);
  srand(1);
  for (size_t i = 0; i < hp.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hp[i] = __float2half(v); fp[i] = __half2float(hp[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hb[i] = __float2half(v); fb[i] = __half2float(hb[i]); }
  __half *dp, *db; float* dout;
  cudaMalloc(&dp, hp.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * BN * 4);
  cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = NPIX * 128 + BN * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  CUtensorMap dummy; memset(&dummy, 0, sizeof(dummy));
  std::vector<float> ho(128 * BN);
  auto check = [&](const char* tag, auto rowpix) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", tag, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0; int bad = 0;
    for (int m = 0; m < 128; ++m) {
      const float* a = rowpix(m);
      for (int n = 0; n < BN; ++n) {
        double acc = 0;
        if (a) for (int k = 0; k < 64; ++k) acc += (double)a[k] * fb[n * 64 + k];
        const double d = fabs(acc - ho[m * BN + n]);
        if (!(d < 1e-2)) ++bad;
        if (d > mx || d != d) mx = d;
      }
    }
    printf("%-44s max|err| = %-10.4g bad = %d  %s\n", tag, mx, bad, bad == 0 ? "OK" : "MISMATCH");
  };
  // ---- part 1: generic-store patch
  const int cfg[][3] = {{0, 8, 0}, {3, 8, 0}, {3, 8, 3}, {3, 10, 0}, {3, 10, 3}, {0, 10, 0}, {19, 18, 0}, {19, 18, 3}, {21, 18, 5}};
  for (auto& c : cfg) {
    probe<<<1, 128, smem>>>(dp, db, dout, c[0], c[1], c[2], 0, dummy, 0, 0, 0);
    char tag[96]; snprintf(tag, sizeof tag, "manual  start=%d px  SBO=%d px  base_off=%d", c[0], c[1], c[2]);
    check(tag, [&](int m) { return &fp[(c[0] + (m / 8) * c[1] + (m % 8)) * 64]; });
  }
  // ---- part 2: TMA box from an NHWC image [N=1][H=40][W=50][C=64(+pad: cstride 72)]
  {
    const int H = 40, W = 50, CS = 72, BW = 18, BH = 18;
    std::vector<__half> himg((size_t)H * W * CS);
    std::vector<float> fimg(himg.size());
    for (size_t i = 0; i < himg.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; himg[i] = __float2half(v); fimg[i] = __half2float(himg[i]); }
    __half* dimg; cudaMalloc(&dimg, himg.size() * 2);
    cudaMemcpy(dimg, himg.data(), himg.size() * 2, cudaMemcpyHostToDevice);
    EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tm;
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t strides[3] = {(cuuint64_t)CS * 2, (cuuint64_t)W * CS * 2, (cuuint64_t)H * W * CS * 2};
    cuuint32_t box[4] = {64, BW, BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dimg, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled -> %d\n", (int)r);
    static float zeros[64] = {0};
    // box origin (cx, cy); tap (ky,kx) of sub-tile s of a 16x16 output tile at (cx+1, cy+1)
    const int origins[][2] = {{5, 7}, {-1, -1}, {40, 30}};
    for (auto& o : origins) for (int tap = 0; tap < 9; tap += 4) for (int s = 0; s < 2; ++s) {
      const int ky = tap / 3, kx = tap % 3;
      const int start = ky * BW + kx + 8 * s;
      probe<<<1, 128, smem>>>(dp, db, dout, start, BW, 0, 1, tm, o[0], o[1], BW * BH * 128);
      char tag[96]; snprintf(tag, sizeof tag, "TMA box@(%d,%d) tap(%d,%d) sub%d", o[0], o[1], ky, kx, s);
      check(tag, [&](int m) -> const float* {
        const int y = o[1] + ky + m / 8, x = o[0] + kx + 8 * s + m % 8;
        if (y < 0 || y >= H || x < 0 || x >= W) return zeros;
        return &fimg[((size_t)y * W + x) * CS];
      });
    }
  }
  return 0;
}
