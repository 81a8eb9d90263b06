// tcgen05 implicit-GEMM convolution for sm_100a.  See conv_igemm.cuh for the contract.
//
// Persistent CTA (one per SM, 448 threads) looping over 128 x BN output tiles:
//   warps 0-3   epilogue: tcgen05.ld of the finished accumulator -> bias/activation/fusions -> global
//   warps 4-11  im2col producers: 16-byte cp.async gathers into a 128B-swizzled K-major A tile
//   warp  12    one lane issues tcgen05.mma (M=128, N=BN, K=16 x4 per 64-wide K chunk)
//   warp  13    TMEM allocation; one lane streams the pre-swizzled weight tile with a single
//               cp.async.bulk (TMA engine) per stage
// Pipelines: `stages` smem slots (full/empty mbarriers) that keep filling across tile boundaries, and two
// TMEM accumulators (acc_full/acc_empty) so the epilogue of tile i overlaps the main loop of tile i+1.
#include <string.h>

#include "conv_igemm.cuh"
#include "conv_epilogue.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int NUM_EPILOGUE = 128;   // warps 0-3
constexpr int NUM_PRODUCERS = 256;  // warps 4-11
constexpr int WARP_MMA = 12;
constexpr int WARP_TMA = 13;
constexpr int NUM_THREADS = 448;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 200 * 1024;

__global__ void __launch_bounds__(NUM_THREADS, 1) conv_igemm_kernel(const __grid_constant__ PPConvParams p) {
  using namespace ppx;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int S = p.stages;
  const int b_stage_bytes = p.BN * 128;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* acc_full = empty_bar + MAX_STAGES;   // [2] accumulator ready for the epilogue
  uint64_t* acc_empty = acc_full + 2;            // [2] accumulator drained, MMA may overwrite
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int num_kc = p.num_kc;
  const int m_tiles = (p.M_total + BM - 1) / BM;
  const int n_tiles = p.Cout_g_pad / p.BN;
  const int total_tiles = m_tiles * n_tiles * p.groups;

  // two accumulator buffers of `acc_cols` TMEM columns each
  uint32_t acc_cols = 32;
  while (acc_cols < (uint32_t)p.BN) acc_cols <<= 1;
  const uint32_t tmem_cols = acc_cols * 2;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], NUM_PRODUCERS + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], NUM_EPILOGUE);
    }
    mbar_fence_init();
  }
  if (warp == WARP_TMA) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation) overlapped the tail of the
  // previous kernel in the stream; from here on we read its outputs.  Let our own dependents start their prologue.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp < 4) {
    // ------------------------------------------------------------------ epilogue warps (TMEM lanes 32*warp..)
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int epi = p.epi;
    const bool vec = p.vec_ok != 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n_idx = tile % n_tiles;
      const int rest = tile / n_tiles;
      const int m0 = (rest % m_tiles) * BM;
      const int g = rest / m_tiles;
      const int n0 = n_idx * p.BN;
      const int buf = it & 1;
      mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const int m = m0 + warp * 32 + lane;
      const bool mvalid = m < p.M_total;
      const uint32_t t_row = tmem_base + lane_base + buf * acc_cols;
      const long long mrow = m;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t raw[16];
        tmem_ld16(t_row + c0, raw);
        tmem_ld_wait();
        if (c0 + 16 >= p.BN) {  // last read of this accumulator: hand the buffer back to the MMA warp
          tc_fence_before();
          mbar_arrive(&acc_empty[buf]);
        }
        const int ng0 = n0 + c0;  // channel within the group
        if (!mvalid || ng0 >= p.Cout_g) continue;
        ppconv::conv_epilogue16(p, raw, mrow, g, ng0, epi, vec);
      }
    }
  } else if (warp < WARP_MMA) {
    // ------------------------------------------------------------------ im2col producers (8 warps)
    const int ptid = tid - 128;
    const int j = ptid & 7;    // 16-byte chunk inside the 128-byte K row
    const int rb = ptid >> 3;  // rows rb, rb+32, rb+64, rb+96
    const uint32_t a_off = rb * 128 + ((j ^ (rb & 7)) << 4);
    int s = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int rest = tile / n_tiles;
      const int m0 = (rest % m_tiles) * BM;
      const int g = rest / m_tiles;
      int rpix[4], riy[4], rix[4];
      uint32_t rvalid = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + rb + 32 * i;
        if (m < p.M_total) {
          const int ox = m % p.OW;
          const int t = m / p.OW;
          const int oy = t % p.OH;
          const int img = t / p.OH;
          rpix[i] = img * p.H * p.W;
          riy[i] = oy * p.sh - p.ph;
          rix[i] = ox * p.sw - p.pw;
          rvalid |= 1u << i;
        } else {
          rpix[i] = 0; riy[i] = 0; rix[i] = 0;
        }
      }
      // this thread's position inside the K range: channel ci of tap (ky, kx); advances by 64 per chunk
      int k = j * 8;
      int tap0 = k / p.Cin;
      int ci = k - tap0 * p.Cin;
      int ky = tap0 / p.kw;
      int kx = tap0 - ky * p.kw;
      for (int kc = 0; kc < num_kc; ++kc) {
        mbar_wait(&empty_bar[s], phase ^ 1);
        const bool kvalid = k < p.K_total;
        int q = 0;
#pragma unroll
        for (int t = 1; t < 4; ++t)
          if (t < p.nseg && ci >= p.seg[t].cbegin) q = t;
        const __half* sbase = p.seg[q].ptr + p.seg[q].coff + g * p.seg[q].gstep + (ci - p.seg[q].cbegin);
        const long long cs = p.seg[q].cstride;
        const int dy = ky * p.dh, dx = kx * p.dw;
        const uint32_t a_dst = smem_u32(smem + s * stage_bytes) + a_off;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int iy = riy[i] + dy, ix = rix[i] + dx;
          bool v = kvalid && ((rvalid >> i) & 1u);
          if (p.pad_replicate) {
            iy = min(max(iy, 0), p.H - 1);
            ix = min(max(ix, 0), p.W - 1);
          } else {
            v = v && ((unsigned)iy < (unsigned)p.H) && ((unsigned)ix < (unsigned)p.W);
          }
          const __half* src = v ? sbase + (long long)(rpix[i] + iy * p.W + ix) * cs : p.seg[0].ptr;
          cp_async16(a_dst + i * (32 * 128), src, v ? 16u : 0u);
        }
        k += BK;
        ci += BK;
        while (ci >= p.Cin) {
          ci -= p.Cin;
          if (++kx == p.kw) { kx = 0; ++ky; }
        }
        // asynchronous arrival: fires when this thread's copies for the stage have landed, so up to `stages`
        // K chunks (across tile boundaries) are in flight without any wait in the producer loop
        cp_async_arrive_noinc(&full_bar[s]);
        if (++s == S) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp == WARP_MMA) {
    // ------------------------------------------------------------------ MMA issuer (elect.sync: one UTCHMMA
    // per tcgen05.mma instead of ptxas' per-lane loop around a uniform-datapath instruction)
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(BM, p.BN);
      int s = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + buf * acc_cols;
        for (int kc = 0; kc < num_kc; ++kc) {
          mbar_wait(&full_bar[s], phase);
          tc_fence_after();
          fence_proxy_async();
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint64_t adesc = umma_desc_sw128_kmajor(a_addr);
          const uint64_t bdesc = umma_desc_sw128_kmajor(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16(d_addr, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (++s == S) { s = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ weight-tile loader (TMA bulk copy)
    if (elect_one()) {
      int s = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % n_tiles) * p.BN;
        const int g = tile / (n_tiles * m_tiles);
        const __half* wsrc = p.wpacked + ((long long)g * num_kc * p.Cout_g_pad + n0) * BK;
        for (int kc = 0; kc < num_kc; ++kc) {
          mbar_wait(&empty_bar[s], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)b_stage_bytes);
          bulk_g2s(smem_u32(smem + s * stage_bytes + A_STAGE_BYTES), wsrc + (long long)kc * p.Cout_g_pad * BK,
                   (uint32_t)b_stage_bytes, &full_bar[s]);
          if (++s == S) { s = 0; phase ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_TMA) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace

namespace {
thread_local char g_last_kind = '?';
}
char pp_last_conv_kind() { return g_last_kind; }

int pp_launch_conv(const PPConvParams& pin, cudaStream_t stream) {
  PPConvParams p = pin;
  PP_REQUIRE(p.BN >= 16 && p.BN <= 256 && p.BN % 16 == 0, "conv: BN=%d must be a multiple of 16 in [16,256]", p.BN);
  PP_REQUIRE(p.Cout_g_pad % p.BN == 0, "conv: Cout_g_pad=%d not a multiple of BN=%d", p.Cout_g_pad, p.BN);
  PP_REQUIRE(p.Cin % 8 == 0, "conv: Cin=%d must be a multiple of 8", p.Cin);
  PP_REQUIRE(p.nseg >= 1 && p.nseg <= 4, "conv: nseg=%d", p.nseg);
  PP_REQUIRE(p.seg[p.nseg - 1].cend == p.Cin && p.seg[0].cbegin == 0, "conv: segments do not cover Cin=%d", p.Cin);
  for (int i = 0; i < p.nseg; ++i) {
    PP_REQUIRE(p.seg[i].cstride % 8 == 0 && p.seg[i].coff % 8 == 0 && p.seg[i].gstep % 8 == 0 &&
                   p.seg[i].cbegin % 8 == 0 && p.seg[i].cend % 8 == 0,
               "conv: segment %d is not 16-byte aligned (cstride=%d coff=%d)", i, p.seg[i].cstride, p.seg[i].coff);
    PP_REQUIRE((reinterpret_cast<uintptr_t>(p.seg[i].ptr) & 15) == 0, "conv: segment %d pointer misaligned", i);
  }
  PP_REQUIRE((reinterpret_cast<uintptr_t>(p.wpacked) & 15) == 0, "conv: weight pointer misaligned");
  p.K_total = p.kh * p.kw * p.Cin;
  p.num_kc = pp_ceil_div(p.K_total, BK);
  p.M_total = p.N * p.OH * p.OW;
  if (p.M_total <= 0) return PP_OK;
  {
    const int esz = p.out_fp32 ? 4 : 2, per16 = 16 / esz;
    auto al = [](const void* ptr, long long cs, long long co, long long gs, int per) {
      return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && cs % per == 0 && co % per == 0 && gs % per == 0);
    };
    bool ok = al(p.out, p.out_cstride, p.out_coff, p.out_gstep, per16) && al(p.aux0, p.aux0_cstride, p.aux0_coff, 0, 8) &&
              al(p.aux1, p.aux1_cstride, p.aux1_coff, 0, 8) && al(p.out2, p.out2_cstride, p.out2_coff, 0, 8);
    if (p.bias != nullptr) ok = ok && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0 && (p.groups == 1 || p.Cout_g % 4 == 0);
    if (p.epi == PP_EPI_GRU_ZR) ok = ok && ((p.Cout_g >> 1) % 16 == 0);
    p.vec_ok = ok ? 1 : 0;
    auto al32 = [](const void* ptr, long long cs, long long co, long long gs) {
      return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && cs % 16 == 0 && co % 16 == 0 && gs % 16 == 0);
    };
    p.vec32_ok = (ok && !p.out_fp32 && al32(p.out, p.out_cstride, p.out_coff, p.out_gstep) &&
                  al32(p.aux0, p.aux0_cstride, p.aux0_coff, 0) && al32(p.aux1, p.aux1_cstride, p.aux1_coff, 0) &&
                  al32(p.out2, p.out2_cstride, p.out2_coff, 0) && (p.epi != PP_EPI_GRU_ZR || ((p.Cout_g >> 1) % 16 == 0)))
                     ? 1 : 0;
  }
  if (pp_prog_recording()) { g_last_kind = 'p'; return pp_prog_record_conv(p); }
  if (pp_conv_halo_eligible(p)) { g_last_kind = 'h'; return pp_launch_conv_halo(p, stream); }
  g_last_kind = 'i';
  PP_REQUIRE(!p.ups2x, "conv: fused x2 upsampling needs the TMA halo kernel (stride 1, one input segment, even H and W)");
  for (int i = 0; i < p.nseg; ++i)
    PP_REQUIRE(p.seg[i].cvalid == 0, "conv: zero-extended input channels (cvalid=%d of %d) need the TMA halo kernel "
               "(stride 1, zero padding, Cin %% 64 == 0)", p.seg[i].cvalid, p.seg[i].cend - p.seg[i].cbegin);   // stride-1 k>1 layers: TMA halo-tile kernel
  const int stage_bytes = A_STAGE_BYTES + p.BN * 128;
  int stages = SMEM_BUDGET / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    PP_CUDA_CHECK(cudaGetDevice(&dev));
    PP_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    PP_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  }
  const long long total_tiles = (long long)pp_ceil_div(p.M_total, BM) * (p.Cout_g_pad / p.BN) * p.groups;
  PP_REQUIRE(total_tiles < (1LL << 31), "conv: too many tiles");
  const int grid = (int)(total_tiles < num_sms ? total_tiles : num_sms);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: see griddepcontrol.wait in the kernel
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_igemm_kernel, p));
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}
