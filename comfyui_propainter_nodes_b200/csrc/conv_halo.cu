// Halo-tile tcgen05 convolution for sm_100a: stride-1 convolutions whose A operand is fed by TMA.
//
// The implicit-GEMM kernel (conv_igemm.cu) fetches every input element once per filter tap (9x for a 3x3) with
// cp.async and streams the weight tile once per 128 output pixels; on the 3x3 / 1x5 / 5x1 layers with <= 256
// output channels that makes it L2->SM fill and instruction-issue bound.  Here one CTA tile is 16 rows x (8*MT)
// columns of output pixels (MT = 1 or 2 sub-tiles of 128 pixels):
//   * per 64-channel chunk, ONE 4-D TMA box load (cp.async.bulk.tensor, SWIZZLE_128B, out-of-image coordinates
//     zero-filled = the conv's zero padding) lands the input patch [(16+(kh-1)dh) x (8MT+(kw-1)dw)] pixels x 128 B
//     in shared memory; the A operand of filter tap (ky,kx) of sub-tile s is a *shifted view* of that patch:
//     UMMA descriptor start = patch + ((ky*dh)*BW + kx*dw + 8s)*128 B, stride between 8-pixel row groups
//     (SBO) = BW*128 B.  The 128B swizzle is a function of the absolute shared-memory address, so views that are
//     not 1024-byte aligned are consistent with what TMA wrote (verified on B200: tools/umma_probe.cu).
//   * the weight tile of (chunk, tap) [BN x 64] is streamed once per tile with cp.async.bulk and feeds both
//     sub-tiles, so weights move once per 256 output pixels.
// L2->SM bytes per output pixel drop ~3x on a 3x3 Cin=256 Cout=128 layer and no thread issues per-element loads.
//
// Warp roles (352 threads, one persistent CTA per SM): warps 0-7 epilogue (TMEM -> registers -> fused epilogue of
// conv_epilogue.cuh -> global), warp 8 patch producer (TMA), warp 9 weight producer (bulk copy), warp 10 MMA issuer
// + TMEM allocation.  Two accumulator sets in TMEM (2 x MT x BN columns) overlap epilogue i with main loop i+1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <memory>

#include "conv_epilogue.cuh"
#include "conv_igemm.cuh"
#include "dcn_sample.cuh"

namespace {

constexpr int NUM_THREADS = 352;
constexpr int NUM_EPI_THREADS = 256;
constexpr int WARP_A = 8, WARP_B = 9, WARP_MMA = 10;
// fused x2-upsample variant (UPS): warps 8-11 interpolate the input patch (warp 8 also issues the low-res TMA loads),
// the weight producer and the MMA issuer move to warps 12 / 13
constexpr int UPS_THREADS = 448, UPS_INTERP_THREADS = 128;
constexpr int UPS_WARP_B = 12, UPS_WARP_MMA = 13;
constexpr int MAX_SA = 4, MAX_SB = 8;
constexpr int SMEM_BUDGET = 208 * 1024;

struct HaloParams {
  PPConvParams c;
  CUtensorMap tmap[4];
  int MT;            // sub-tiles (128 pixels each) per CTA tile
  int BW, BH;        // patch size in pixels
  int tiles_x, tiles_y, n_tiles;
  int a_stage_bytes, b_stage_bytes, SA, SB;
  int tps;           // filter taps per weight stage: narrow N tiles pack several taps' [BN x 64] tiles into one stage, so
                     // the MMA issuer waits / commits once per group instead of once per tap (it is issue-bound there)
  int accw;          // TMEM columns per accumulator
  int chunks;        // Cin / 64
  int flat;          // 1x1 convs: tiles are runs of 128*MT consecutive pixels of the flattened [N*H*W] pixel list
  int n_img;         // images the tile index decomposes over (1 in flat mode)
  int sub_bytes;     // A-view offset between the sub-tiles: 8 pixels (spatial) or 128 pixels (flat)
  int ups;           // input tensor is half resolution: bilinear x2 (align_corners) on the fly (UPS kernel)
  int LH, LW;        // low-res source size
  int LBW, LBH;      // low-res staging box (pixels)
  int l_stage_bytes;
  int debug;         // bit 0: skip the epilogue math/stores (PP_CONV_NOEPI=1, mainloop-only timing experiments)
  // flat-mode layers with few K chunks are bound by the epilogue's per-thread 32-byte global stores (one L1 wavefront per
  // lane).  tstore: the epilogue writes the fp16 tile into a 128B-swizzled shared-memory staging tile and ONE thread
  // issues TMA stores ([128 rows x 64 columns] boxes, rows / columns beyond the tensor clipped by the TMA unit).
  int tstore;
  int out_stage_bytes;     // staging bytes per sub-tile: ceil(BN / 64) panels of [128 rows][128 B]
  CUtensorMap tmap_out;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(ppx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

struct TileCoord {
  int n_idx, tx, ty, img, g;
};
__device__ __forceinline__ TileCoord decode_tile(const HaloParams& h, int tile) {
  TileCoord t;
  t.n_idx = tile % h.n_tiles;
  int r = tile / h.n_tiles;
  t.tx = r % h.tiles_x; r /= h.tiles_x;
  t.ty = r % h.tiles_y; r /= h.tiles_y;
  t.img = r % h.n_img;
  t.g = r / h.n_img;
  return t;
}

template <bool UPS>
__global__ void __launch_bounds__(UPS ? UPS_THREADS : NUM_THREADS, 1) conv_halo_kernel(const __grid_constant__ HaloParams h) {
  using namespace ppx;
  const PPConvParams& p = h.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* smem_b = smem + h.SA * h.a_stage_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + h.SB * h.b_stage_bytes);
  uint64_t* a_empty = a_full + MAX_SA;
  uint64_t* b_full = a_empty + MAX_SA;
  uint64_t* b_empty = b_full + MAX_SB;
  uint64_t* acc_full = b_empty + MAX_SB;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* l_full = acc_empty + 2;     // UPS: low-res staging ring (2 stages)
  uint64_t* l_empty = l_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(l_empty + 2);
  uint8_t* stg_out = smem_b + h.SB * h.b_stage_bytes + 2048;     // TMA-store staging (flat layers), 1024-byte aligned
  constexpr int W_B = UPS ? UPS_WARP_B : WARP_B, W_MMA = UPS ? UPS_WARP_MMA : WARP_MMA;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total_tiles = h.n_tiles * h.tiles_x * h.tiles_y * h.n_img * p.groups;
  const int taps = p.kh * p.kw;
  const uint32_t set_cols = (uint32_t)(h.MT * h.accw);
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2 * set_cols) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < h.SA; ++s) { mbar_init(&a_full[s], UPS ? UPS_INTERP_THREADS : 1); mbar_init(&a_empty[s], 1); }
    if (UPS) for (int s = 0; s < 2; ++s) { mbar_init(&l_full[s], 1); mbar_init(&l_empty[s], UPS_INTERP_THREADS); }
    for (int s = 0; s < h.SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], NUM_EPI_THREADS); }
    mbar_fence_init();
  }
  if (warp == W_MMA) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp < 8) {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int r = quarter * 32 + lane;          // row of the 128-pixel sub-tile: 16 rows x 8 columns
    const int epi = p.epi;
    const bool vec = p.vec_ok != 0;
    const bool has_aux = p.aux0 != nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const TileCoord t = decode_tile(h, tile);
      const int n0 = t.n_idx * p.BN;
      const int bnt = min(p.BN, p.Cout_g_pad - n0);
      const int set = it & 1;
      // MT == 2: warps 0-3 own sub-tile 0, warps 4-7 sub-tile 1.  MT == 1: the two halves split the columns.
      int sub = 0, c_lo = 0, c_hi = bnt;
      if (h.MT == 2) sub = half;
      else {
        const int split = ((bnt / 16 + 1) / 2) * 16;
        c_lo = half ? split : 0;
        c_hi = half ? bnt : split;
      }
      bool mvalid;
      long long mrow;
      if (h.flat) {
        mrow = ((long long)t.tx * h.MT + sub) * 128 + r;
        mvalid = mrow < p.M_total;
      } else {
        const int oy = t.ty * 16 + (r >> 3), ox = t.tx * (8 * h.MT) + 8 * sub + (r & 7);
        mvalid = oy < p.OH && ox < p.OW;
        mrow = ((long long)t.img * p.OH + oy) * p.OW + ox;
      }
      mbar_wait(&acc_full[set], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_row = tmem_base + lane_base + set * set_cols + sub * h.accw;
      const bool skip = !mvalid || (h.debug & 1);
      // TMA-store path: staging tile of this sub-tile, its barrier (the 4 warps of a sub-tile, or all 8 when MT == 1)
      uint8_t* stg = nullptr;
      const int bar_id = 2 + (h.MT == 2 ? half : 0), bar_n = h.MT == 2 ? 128 : 256;
      const bool issuer = h.MT == 2 ? (tid == half * 128) : (tid == 0);
      if (h.tstore) {
        stg = stg_out + sub * h.out_stage_bytes;
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous tile's stores have read it
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory");
      }
      // 32 columns per round: both TMEM loads are in flight before the single wait
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        uint32_t raw0[16], raw1[16];
        const bool two = c0 + 16 < c_hi;
        tmem_ld16(t_row + c0, raw0);
        if (two) tmem_ld16(t_row + c0 + 16, raw1);
        const bool do0 = !skip && n0 + c0 < p.Cout_g, do1 = !skip && two && n0 + c0 + 16 < p.Cout_g;
        ppconv::EpiAux x0, x1;
        x0.have = x1.have = false;
        if (has_aux) {   // residual / GRU operands: issued while the TMEM reads are in flight
          if (do0) ppconv::conv_epilogue_prefetch16(p, mrow, n0 + c0, epi, vec, x0);
          if (do1) ppconv::conv_epilogue_prefetch16(p, mrow, n0 + c0 + 16, epi, vec, x1);
        }
        tmem_ld_wait();
        if (c0 + 32 >= c_hi) {   // last read of this accumulator set by this thread: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&acc_empty[set]);
        }
        if (stg != nullptr) {
          // 16 columns = two 16-byte units of panel c/64, row r, 128B swizzle (unit index XOR row & 7)
          auto slot = [&](int c, int u) {
            const int unit = ((c & 63) >> 3) + u;
            return reinterpret_cast<uint4*>(stg + (c >> 6) * 16384 + r * 128 + ((unit ^ (r & 7)) << 4));
          };
          if (do0) ppconv::conv_epilogue16(p, raw0, mrow, t.g, n0 + c0, epi, vec, &x0, slot(c0, 0), slot(c0, 1));
          if (do1) ppconv::conv_epilogue16(p, raw1, mrow, t.g, n0 + c0 + 16, epi, vec, &x1, slot(c0 + 16, 0), slot(c0 + 16, 1));
        } else {
          if (do0) ppconv::conv_epilogue16(p, raw0, mrow, t.g, n0 + c0, epi, vec, &x0);
          if (do1) ppconv::conv_epilogue16(p, raw1, mrow, t.g, n0 + c0 + 16, epi, vec, &x1);
        }
      }
      if (c_lo >= c_hi) {
        tc_fence_before();
        mbar_arrive(&acc_empty[set]);
      }
      if (stg != nullptr) {
        fence_proxy_async();                                  // generic-proxy smem writes -> visible to the TMA store
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory");
        if (issuer && !(h.debug & 1)) {
          const int row0 = (int)(((long long)t.tx * h.MT + sub) * 128);
          for (int pnl = 0; pnl * 64 < bnt; ++pnl)
            tma_store_2d(&h.tmap_out, smem_u32(stg + pnl * 16384), n0 + pnl * 64, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (h.tstore) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // outstanding stores of this thread complete
  } else if (UPS && warp < 12) {
    // ------------------------------------------------------------------ fused bilinear x2 (align_corners=True) producer
    // The conv's input is the x2 upsampling of a half-resolution tensor (reference deconv = F.interpolate + conv).
    // Per 64-channel chunk the elected lane of warp 8 TMA-loads the low-res patch that covers the 18x18 hi-res patch
    // into a 2-stage staging ring; the 128 threads interpolate it (fp32, one rounding, same expression as the
    // stand-alone upsample kernel) straight into the 128B-swizzled A stage.  Hi-res pixels outside the image are the
    // conv's zero padding.  The upsampled tensor (4x the pixels) never exists in memory.
    const int it_id = tid - 256;
    uint8_t* stg = smem_b + h.SB * h.b_stage_bytes + 1024;     // staging ring behind the barrier block
    const float sy = (float)(h.LH - 1) / (float)(2 * h.LH - 1), sx = (float)(h.LW - 1) / (float)(2 * h.LW - 1);
    const uint32_t lbytes = (uint32_t)(h.LBW * h.LBH * 128);
    const bool issuer = warp == 8 && elect_one();
    auto low_origin = [&](const TileCoord& t, int& xlo, int& ylo) {
      const int X0 = t.tx * (8 * h.MT) - p.pw, Y0 = t.ty * 16 - p.ph;
      xlo = (int)(sx * (float)max(X0, 0));
      ylo = (int)(sy * (float)max(Y0, 0));
    };
    auto issue = [&](int tile, int c, int ls, uint32_t lph) {
      const TileCoord t = decode_tile(h, tile);
      int xlo, ylo;
      low_origin(t, xlo, ylo);
      mbar_wait(&l_empty[ls], lph ^ 1);
      mbar_arrive_expect_tx(&l_full[ls], lbytes);
      tma_load_4d(smem_u32(stg + ls * h.l_stage_bytes), &h.tmap[0], c * 64, xlo, ylo, t.img, &l_full[ls]);
    };
    int ls = 0, sa = 0;
    uint32_t lph = 0, pa = 0;
    // producer-side schedule of staging slots: (stage, phase) advance once per issued chunk
    int is_ls = 0;
    uint32_t is_ph = 0;
    if (issuer && blockIdx.x < total_tiles) {
      issue(blockIdx.x, 0, is_ls, is_ph);
      if (++is_ls == 2) { is_ls = 0; is_ph ^= 1; }
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(h, tile);
      const int X0 = t.tx * (8 * h.MT) - p.pw, Y0 = t.ty * 16 - p.ph;
      int xlo, ylo;
      low_origin(t, xlo, ylo);
      for (int c = 0; c < h.chunks; ++c) {
        if (issuer) {   // prefetch the next chunk's low-res patch
          int nt = tile, nc = c + 1;
          if (nc == h.chunks) { nc = 0; nt += gridDim.x; }
          if (nt < total_tiles) {
            issue(nt, nc, is_ls, is_ph);
            if (++is_ls == 2) { is_ls = 0; is_ph ^= 1; }
          }
        }
        __syncwarp();
        mbar_wait(&l_full[ls], lph);
        mbar_wait(&a_empty[sa], pa ^ 1);
        const uint8_t* src = stg + ls * h.l_stage_bytes;
        uint8_t* dstA = smem + sa * h.a_stage_bytes;
        const int items = h.BW * h.BH * 8;
        for (int item = it_id; item < items; item += UPS_INTERP_THREADS) {
          const int ch = item & 7, pp = item >> 3;
          const int py = pp / h.BW, px = pp - py * h.BW;
          const int Y = Y0 + py, X = X0 + px;
          uint4 o = make_uint4(0, 0, 0, 0);
          if ((unsigned)Y < (unsigned)(2 * h.LH) && (unsigned)X < (unsigned)(2 * h.LW)) {
            const float fy = sy * (float)Y, fx = sx * (float)X;
            const int y0 = (int)fy, x0 = (int)fx;
            const int y1 = min(y0 + 1, h.LH - 1), x1 = min(x0 + 1, h.LW - 1);
            const float ly = fy - (float)y0, lx = fx - (float)x0;
            const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
            const uint8_t* r0 = src + ((y0 - ylo) * h.LBW - xlo) * 128 + ch * 16;
            const uint8_t* r1 = src + ((y1 - ylo) * h.LBW - xlo) * 128 + ch * 16;
            const uint4 qa = *reinterpret_cast<const uint4*>(r0 + x0 * 128), qb = *reinterpret_cast<const uint4*>(r0 + x1 * 128);
            const uint4 qc = *reinterpret_cast<const uint4*>(r1 + x0 * 128), qd = *reinterpret_cast<const uint4*>(r1 + x1 * 128);
            const __half2* ah = reinterpret_cast<const __half2*>(&qa);
            const __half2* bh = reinterpret_cast<const __half2*>(&qb);
            const __half2* chh = reinterpret_cast<const __half2*>(&qc);
            const __half2* dh = reinterpret_cast<const __half2*>(&qd);
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 fa = __half22float2(ah[i]), fb = __half22float2(bh[i]), fc = __half22float2(chh[i]),
                           fd = __half22float2(dh[i]);
              oh[i] = __floats2half2_rn(w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x,
                                        w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y);
            }
          }
          *reinterpret_cast<uint4*>(dstA + pp * 128 + ((ch ^ (pp & 7)) << 4)) = o;
        }
        fence_proxy_async();            // generic-proxy writes of the A stage -> visible to tcgen05.mma
        mbar_arrive(&a_full[sa]);
        mbar_arrive(&l_empty[ls]);
        if (++ls == 2) { ls = 0; lph ^= 1; }
        if (++sa == h.SA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (!UPS && warp == WARP_A) {
    // ------------------------------------------------------------------ input patch producer (TMA)
    if (ppx::elect_one()) {
      int s = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (uint32_t)(h.BW * h.BH * 128);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(h, tile);
        const int x0 = h.flat ? t.tx * (128 * h.MT) : t.tx * (8 * h.MT) - p.pw, y0 = h.flat ? 0 : t.ty * 16 - p.ph;
        for (int c = 0; c < h.chunks; ++c) {
          const int ci = c * 64;
          int q = 0;
#pragma unroll
          for (int k = 1; k < 4; ++k)
            if (k < p.nseg && ci >= p.seg[k].cbegin) q = k;
          const int ch0 = t.g * p.seg[q].gstep + (ci - p.seg[q].cbegin);
          mbar_wait(&a_empty[s], phase ^ 1);
          mbar_arrive_expect_tx(&a_full[s], bytes);
          tma_load_4d(smem_u32(smem + s * h.a_stage_bytes), &h.tmap[q], ch0, x0, y0, t.img, &a_full[s]);
          if (++s == h.SA) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == W_B) {
    // ------------------------------------------------------------------ weight tile producer (bulk copy)
    if (ppx::elect_one()) {
      int s = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(h, tile);
        const int n0 = t.n_idx * p.BN;
        const uint32_t bytes = (uint32_t)(min(p.BN, p.Cout_g_pad - n0) * 128);
        const __half* wbase = p.wpacked + ((long long)t.g * p.num_kc * p.Cout_g_pad + n0) * 64;
        for (int c = 0; c < h.chunks; ++c) {
          for (int tap0 = 0; tap0 < taps; tap0 += h.tps) {
            const int tn = min(h.tps, taps - tap0);
            mbar_wait(&b_empty[s], phase ^ 1);
            mbar_arrive_expect_tx(&b_full[s], bytes * (uint32_t)tn);
            for (int t = 0; t < tn; ++t) {
              const int kc = (tap0 + t) * h.chunks + c;
              bulk_g2s(smem_u32(smem_b + s * h.b_stage_bytes + t * p.BN * 128), wbase + (long long)kc * p.Cout_g_pad * 64, bytes,
                       &b_full[s]);
            }
            if (++s == h.SB) { s = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    // One elected lane (elect.sync lets ptxas emit each tcgen05.mma once instead of a per-lane loop).  The loop is
    // kept lean -- descriptors advance by precomputed 16-byte-unit steps -- because a single thread has to issue
    // one UTCHMMA per 32-64 tensor-pipe cycles.
    if (ppx::elect_one()) {
      int sa = 0, sb = 0, it = 0;
      uint32_t pa = 0, pb = 0;
      const uint32_t sbo = h.flat ? 1024u : (uint32_t)h.BW * 128;
      const uint64_t a_hi = ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t b_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t a_base0 = (smem_u32(smem) & 0x3FFFF) >> 4, b_base0 = (smem_u32(smem_b) & 0x3FFFF) >> 4;
      const uint32_t a_stage16 = (uint32_t)h.a_stage_bytes >> 4, b_stage16 = (uint32_t)h.b_stage_bytes >> 4;
      const uint32_t step_x = (uint32_t)p.dw * 8;                                           // next tap in the row
      const uint32_t step_row = (uint32_t)(p.dh * h.BW - (p.kw - 1) * p.dw) * 8;            // last tap of a row -> next row
      const uint32_t sub16 = (uint32_t)h.sub_bytes >> 4;
      const uint32_t tap16 = (uint32_t)p.BN * 8;                                            // next tap's weight tile in the stage
      const bool two = h.MT == 2;
      const int kw = p.kw;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int n0 = (tile % h.n_tiles) * p.BN;
        const uint32_t idesc = umma_idesc_f16(128, (uint32_t)min(p.BN, p.Cout_g_pad - n0));
        const int set = it & 1;
        mbar_wait(&acc_empty[set], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + set * set_cols, d1 = d0 + h.accw;
        uint32_t accum = 0;
        for (int c = 0; c < h.chunks; ++c) {
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          uint64_t adesc = a_hi | (uint64_t)(a_base0 + sa * a_stage16);
          int kx = 0;
          for (int tap0 = 0; tap0 < taps; tap0 += h.tps) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            uint64_t bdesc = b_hi | (uint64_t)(b_base0 + sb * b_stage16);
            const int tn = min(h.tps, taps - tap0);
            for (int t = 0; t < tn; ++t) {
              umma_f16(d0, adesc, bdesc, idesc, accum);
              umma_f16(d0, adesc + 2, bdesc + 2, idesc, 1u);
              umma_f16(d0, adesc + 4, bdesc + 4, idesc, 1u);
              umma_f16(d0, adesc + 6, bdesc + 6, idesc, 1u);
              if (two) {
                const uint64_t adesc1 = adesc + sub16;
                umma_f16(d1, adesc1, bdesc, idesc, accum);
                umma_f16(d1, adesc1 + 2, bdesc + 2, idesc, 1u);
                umma_f16(d1, adesc1 + 4, bdesc + 4, idesc, 1u);
                umma_f16(d1, adesc1 + 6, bdesc + 6, idesc, 1u);
              }
              accum = 1u;
              bdesc += tap16;
              adesc += step_x;
              if (++kx == kw) { kx = 0; adesc += step_row - step_x; }
            }
            umma_commit(&b_empty[sb]);
            if (++sb == h.SB) { sb = 0; pb ^= 1; }
          }
          umma_commit(&a_empty[sa]);
          if (++sa == h.SA) { sa = 0; pa ^= 1; }
        }
        umma_commit(&acc_full[set]);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// Multi-layer program: up to PROG_MAX_LAYERS dependent layers (stride-1 convolutions of the kind above, and modulated
// deformable sampling) executed by ONE launch of one CTA per SM, with a grid-wide barrier between consecutive layers.
// The recurrent propagation steps of flow completion are 8 dependent layers over 3,600-7,200 pixels: as separate
// launches each pays launch + prologue + pipeline fill/drain (~15-30 us, mostly fixed); here the fixed cost per layer
// is one barrier (an atomic counter in global memory) and one TMA round trip, the weight producer runs ahead across the
// barrier, and TMEM / mbarriers / tensor maps are set up once.
//
// Ordering between layers: the epilogue threads (the only writers of global memory) make their generic-proxy stores
// visible to the async proxy (fence.proxy.async.global), meet on a named barrier, and one thread publishes the CTA's
// arrival (threadfence + atomicAdd).  The TMA producer and the epilogue warps of the next layer spin on the counter
// (ld.acquire.gpu) before they touch that layer's inputs; the producer adds the consumer-side proxy fence before
// issuing TMA loads.  All 148 CTAs are co-resident (1 CTA per SM by shared memory), so the spin cannot deadlock.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PROG_MAX_LAYERS = 10;
enum { PROG_CONV = 0, PROG_DCN = 1 };

struct ProgParams {
  int n_layers;
  int SA, SB, a_stage_bytes, b_stage_bytes;   // one shared-memory carve-up for every layer
  unsigned int* counter;                      // arrivals since the counter was zeroed
  unsigned int base;                          // arrivals issued before this launch
  unsigned long long* ts;                     // debug (PP_PROG_TS=1): globaltimer of CTA 0 at [layer start, layer end]
  int kind[PROG_MAX_LAYERS];
  PPDcnArgs dcn[PROG_MAX_LAYERS];
  HaloParams layer[PROG_MAX_LAYERS];
};

__device__ __forceinline__ void prog_wait(const unsigned int* counter, unsigned int target) {
  unsigned int v, spins = 0;
  uint64_t t0 = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if ((int)(v - target) >= 0) return;
    __nanosleep(64);                      // one poller per CTA, backed off: the arrivals' atomics are not starved
    if ((++spins & 0xFFFu) == 0) {       // a lost arrival must not hang the GPU: trap after ~2 s
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__global__ void __launch_bounds__(NUM_THREADS, 1) conv_prog_kernel(const __grid_constant__ ProgParams P) {
  using namespace ppx;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* smem_b = smem + P.SA * P.a_stage_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + P.SB * P.b_stage_bytes);
  uint64_t* a_empty = a_full + MAX_SA;
  uint64_t* b_full = a_empty + MAX_SA;
  uint64_t* b_empty = b_full + MAX_SB;
  uint64_t* acc_full = b_empty + MAX_SB;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* layer_go = acc_empty + 2;     // the CTA's one poller (TMA producer thread) -> epilogue warps: layer li may start
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2 + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned int G = gridDim.x;
  if (tid == 0) {
    mbar_init(layer_go, 1);
    for (int s = 0; s < P.SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < P.SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], NUM_EPI_THREADS); }
    mbar_fence_init();
  }
  if (warp == WARP_MMA) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp < 8) {
    // ------------------------------------------------------------------ epilogue warps (+ the sampling layers)
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int r = quarter * 32 + lane;
    int it = 0;
    for (int li = 0; li < P.n_layers; ++li) {
      // inputs of this layer (residuals, sampling sources) were written by the previous one: the producer thread polls
      // the grid counter for the whole CTA and releases the epilogue warps through a shared-memory barrier
      mbar_wait(layer_go, (uint32_t)li & 1u);
      if (P.ts != nullptr && blockIdx.x == 0 && tid == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        P.ts[2 * li] = now;
      }
      if (P.kind[li] == PROG_DCN) {
        const PPDcnArgs& a = P.dcn[li];
        const unsigned per_img = (unsigned)(a.H * a.W * 144);
        const unsigned total = per_img * (unsigned)a.N;
        const unsigned stride = G * NUM_EPI_THREADS;
        // the sampling is bound by L1 wavefronts (one per uncoalesced 16/32-byte gather), not by latency: a two-phase
        // batched form (all offsets first, then all gathers) measured slower (78 vs 52 us per layer)
        for (unsigned idx = blockIdx.x * NUM_EPI_THREADS + tid; idx < total; idx += 4 * stride) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const unsigned i2 = idx + u * stride;
            if (i2 < total) {
              const unsigned n = i2 / per_img;
              if (a.C == 128) dcn_sample_item<8, true>(a, i2 - n * per_img, (int)n);
              else dcn_sample_item<16, true>(a, i2 - n * per_img, (int)n);
            }
          }
        }
      } else {
        const HaloParams& h = P.layer[li];
        const PPConvParams& p = h.c;
        const int total_tiles = h.n_tiles * h.tiles_x * h.tiles_y * h.n_img * p.groups;
        const uint32_t set_cols = (uint32_t)(h.MT * h.accw);
        const int epi = p.epi;
        const bool vec = p.vec_ok != 0;
        const bool has_aux = p.aux0 != nullptr;
        for (int tile = blockIdx.x; tile < total_tiles; tile += G, ++it) {
          const TileCoord t = decode_tile(h, tile);
          const int n0 = t.n_idx * p.BN;
          const int bnt = min(p.BN, p.Cout_g_pad - n0);
          const int set = it & 1;
          int sub = 0, c_lo = 0, c_hi = bnt;
          if (h.MT == 2) sub = half;
          else {
            const int split = ((bnt / 16 + 1) / 2) * 16;
            c_lo = half ? split : 0;
            c_hi = half ? bnt : split;
          }
          bool mvalid;
          long long mrow;
          if (h.flat) {
            mrow = ((long long)t.tx * h.MT + sub) * 128 + r;
            mvalid = mrow < p.M_total;
          } else {
            const int oy = t.ty * 16 + (r >> 3), ox = t.tx * (8 * h.MT) + 8 * sub + (r & 7);
            mvalid = oy < p.OH && ox < p.OW;
            mrow = ((long long)t.img * p.OH + oy) * p.OW + ox;
          }
          mbar_wait(&acc_full[set], (uint32_t)(it >> 1) & 1u);
          tc_fence_after();
          const uint32_t t_row = tmem_base + lane_base + set * 256u + sub * h.accw;
          (void)set_cols;
          for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            uint32_t raw0[16], raw1[16];
            const bool two = c0 + 16 < c_hi;
            tmem_ld16(t_row + c0, raw0);
            if (two) tmem_ld16(t_row + c0 + 16, raw1);
            const bool do0 = mvalid && n0 + c0 < p.Cout_g, do1 = mvalid && two && n0 + c0 + 16 < p.Cout_g;
            ppconv::EpiAux x0, x1;
            x0.have = x1.have = false;
            if (has_aux) {
              if (do0) ppconv::conv_epilogue_prefetch16(p, mrow, n0 + c0, epi, vec, x0);
              if (do1) ppconv::conv_epilogue_prefetch16(p, mrow, n0 + c0 + 16, epi, vec, x1);
            }
            tmem_ld_wait();
            if (c0 + 32 >= c_hi) {
              tc_fence_before();
              mbar_arrive(&acc_empty[set]);
            }
            if (do0) ppconv::conv_epilogue16(p, raw0, mrow, t.g, n0 + c0, epi, vec, &x0);
            if (do1) ppconv::conv_epilogue16(p, raw1, mrow, t.g, n0 + c0 + 16, epi, vec, &x1);
          }
          if (c_lo >= c_hi) {
            tc_fence_before();
            mbar_arrive(&acc_empty[set]);
          }
        }
      }
      // publish this CTA's part of the layer: stores -> async proxy, CTA-wide meet of the writers, one arrival
      fence_proxy_async_global();
      asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_THREADS) : "memory");
      if (tid == 0) {
        __threadfence();
        atomicAdd(P.counter, 1u);
        if (P.ts != nullptr && blockIdx.x == 0) {
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          P.ts[2 * li + 1] = now;
        }
      }
    }
  } else if (warp == WARP_A) {
    // ------------------------------------------------------------------ input patch producer (TMA)
    if (ppx::elect_one()) {
      int s = 0;
      uint32_t phase = 0;
      for (int li = 0; li < P.n_layers; ++li) {
        if (li > 0) {     // the CTA's only poller: every CTA finished layer li-1
          prog_wait(P.counter, P.base + (unsigned int)li * G);
          fence_proxy_async_global();
        }
        mbar_arrive(layer_go);
        if (P.kind[li] != PROG_CONV) continue;
        const HaloParams& h = P.layer[li];
        const PPConvParams& p = h.c;
        const int total_tiles = h.n_tiles * h.tiles_x * h.tiles_y * h.n_img * p.groups;
        const uint32_t bytes = (uint32_t)(h.BW * h.BH * 128);
        for (int tile = blockIdx.x; tile < total_tiles; tile += G) {
          const TileCoord t = decode_tile(h, tile);
          const int x0 = h.flat ? t.tx * (128 * h.MT) : t.tx * (8 * h.MT) - p.pw, y0 = h.flat ? 0 : t.ty * 16 - p.ph;
          for (int c = 0; c < h.chunks; ++c) {
            const int ci = c * 64;
            int q = 0;
#pragma unroll
            for (int k = 1; k < 4; ++k)
              if (k < p.nseg && ci >= p.seg[k].cbegin) q = k;
            const int ch0 = t.g * p.seg[q].gstep + (ci - p.seg[q].cbegin);
            mbar_wait(&a_empty[s], phase ^ 1);
            mbar_arrive_expect_tx(&a_full[s], bytes);
            tma_load_4d(smem_u32(smem + s * P.a_stage_bytes), &h.tmap[q], ch0, x0, y0, t.img, &a_full[s]);
            if (++s == P.SA) { s = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == WARP_B) {
    // ------------------------------------------------------------------ weight tile producer: never waits for a layer
    if (ppx::elect_one()) {
      int s = 0;
      uint32_t phase = 0;
      for (int li = 0; li < P.n_layers; ++li) {
        if (P.kind[li] != PROG_CONV) continue;
        const HaloParams& h = P.layer[li];
        const PPConvParams& p = h.c;
        const int total_tiles = h.n_tiles * h.tiles_x * h.tiles_y * h.n_img * p.groups;
        const int taps = p.kh * p.kw;
        for (int tile = blockIdx.x; tile < total_tiles; tile += G) {
          const TileCoord t = decode_tile(h, tile);
          const int n0 = t.n_idx * p.BN;
          const uint32_t bytes = (uint32_t)(min(p.BN, p.Cout_g_pad - n0) * 128);
          const __half* wbase = p.wpacked + ((long long)t.g * p.num_kc * p.Cout_g_pad + n0) * 64;
          for (int c = 0; c < h.chunks; ++c) {
            for (int tap0 = 0; tap0 < taps; tap0 += h.tps) {
              const int tn = min(h.tps, taps - tap0);
              mbar_wait(&b_empty[s], phase ^ 1);
              mbar_arrive_expect_tx(&b_full[s], bytes * (uint32_t)tn);
              for (int t = 0; t < tn; ++t) {
                const int kc = (tap0 + t) * h.chunks + c;
                bulk_g2s(smem_u32(smem_b + s * P.b_stage_bytes + t * p.BN * 128), wbase + (long long)kc * p.Cout_g_pad * 64, bytes,
                         &b_full[s]);
              }
              if (++s == P.SB) { s = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    if (ppx::elect_one()) {
      int sa = 0, sb = 0, it = 0;
      uint32_t pa = 0, pb = 0;
      const uint64_t b_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t a_base0 = (smem_u32(smem) & 0x3FFFF) >> 4, b_base0 = (smem_u32(smem_b) & 0x3FFFF) >> 4;
      const uint32_t a_stage16 = (uint32_t)P.a_stage_bytes >> 4, b_stage16 = (uint32_t)P.b_stage_bytes >> 4;
      for (int li = 0; li < P.n_layers; ++li) {
        if (P.kind[li] != PROG_CONV) continue;
        const HaloParams& h = P.layer[li];
        const PPConvParams& p = h.c;
        const int total_tiles = h.n_tiles * h.tiles_x * h.tiles_y * h.n_img * p.groups;
        const int taps = p.kh * p.kw;
        const uint32_t sbo = h.flat ? 1024u : (uint32_t)h.BW * 128;
        const uint64_t a_hi = ((uint64_t)1 << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint32_t step_x = (uint32_t)p.dw * 8;
        const uint32_t step_row = (uint32_t)(p.dh * h.BW - (p.kw - 1) * p.dw) * 8;
        const uint32_t sub16 = (uint32_t)h.sub_bytes >> 4;
        const uint32_t tap16 = (uint32_t)p.BN * 8;
        const bool two = h.MT == 2;
        const int kw = p.kw;
        for (int tile = blockIdx.x; tile < total_tiles; tile += G, ++it) {
          const int n0 = (tile % h.n_tiles) * p.BN;
          const uint32_t idesc = umma_idesc_f16(128, (uint32_t)min(p.BN, p.Cout_g_pad - n0));
          const int set = it & 1;
          mbar_wait(&acc_empty[set], ((uint32_t)(it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d0 = tmem_base + set * 256u, d1 = d0 + h.accw;
          uint32_t accum = 0;
          for (int c = 0; c < h.chunks; ++c) {
            mbar_wait(&a_full[sa], pa);
            tc_fence_after();
            uint64_t adesc = a_hi | (uint64_t)(a_base0 + sa * a_stage16);
            int kx = 0;
            for (int tap0 = 0; tap0 < taps; tap0 += h.tps) {
              mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              uint64_t bdesc = b_hi | (uint64_t)(b_base0 + sb * b_stage16);
              const int tn = min(h.tps, taps - tap0);
              for (int t = 0; t < tn; ++t) {
                umma_f16(d0, adesc, bdesc, idesc, accum);
                umma_f16(d0, adesc + 2, bdesc + 2, idesc, 1u);
                umma_f16(d0, adesc + 4, bdesc + 4, idesc, 1u);
                umma_f16(d0, adesc + 6, bdesc + 6, idesc, 1u);
                if (two) {
                  const uint64_t adesc1 = adesc + sub16;
                  umma_f16(d1, adesc1, bdesc, idesc, accum);
                  umma_f16(d1, adesc1 + 2, bdesc + 2, idesc, 1u);
                  umma_f16(d1, adesc1 + 4, bdesc + 4, idesc, 1u);
                  umma_f16(d1, adesc1 + 6, bdesc + 6, idesc, 1u);
                }
                accum = 1u;
                bdesc += tap16;
                adesc += step_x;
                if (++kx == kw) { kx = 0; adesc += step_row - step_x; }
              }
              umma_commit(&b_empty[sb]);
              if (++sb == P.SB) { sb = 0; pb ^= 1; }
            }
            umma_commit(&a_empty[sa]);
            if (++sa == P.SA) { sa = 0; pa ^= 1; }
          }
          umma_commit(&acc_full[set]);
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) tmem_dealloc(tmem_base, 512);
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

// 0 = not eligible (caller falls back to the cp.async implicit-GEMM kernel), 1 = eligible.
int pp_conv_halo_eligible(const PPConvParams& p) {
  static int enabled = -1, allow_1x1 = -1;
  if (enabled < 0) {
    const char* e = getenv("PP_CONV_HALO");
    enabled = (e == nullptr || atoi(e) != 0) ? 1 : 0;
    e = getenv("PP_HALO_1X1");
    allow_1x1 = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  if (!enabled) return 0;
  if (p.sh != 1 || p.sw != 1 || p.pad_replicate) return 0;
  const bool flat = p.kh * p.kw == 1;
  if (flat) {
    // 1x1 conv / linear layer: tiles are runs of consecutive pixels; a ragged channel tail is zero-filled by TMA
    if (!allow_1x1 || p.ph != 0 || p.pw != 0 || p.groups != 1) return 0;
  } else if (p.Cin % 64 != 0) {
    return 0;   // packed K order is (tap, ci): 64-channel chunks must not straddle taps
  }
  for (int i = 0; i < p.nseg; ++i) {
    if (p.seg[i].cbegin % 64 != 0) return 0;
    if (p.seg[i].cend % 64 != 0 && !(flat && i == p.nseg - 1)) return 0;
  }
  if ((p.kw - 1) * p.dw + 16 > 256 || (p.kh - 1) * p.dh + 16 > 256) return 0;
  if ((long long)p.N * p.OH * p.OW < 128) return 0;
  if (p.ups2x && (flat || p.nseg != 1 || p.groups != 1 || p.H % 2 != 0 || p.W % 2 != 0 || p.H < 4 || p.W < 4)) return 0;
  return encode_fn() != nullptr ? 1 : 0;
}

namespace {

int halo_num_sms(int* out) {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    PP_CUDA_CHECK(cudaGetDevice(&dev));
    PP_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    PP_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    PP_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    PP_CUDA_CHECK(cudaFuncSetAttribute(conv_prog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  }
  *out = num_sms;
  return PP_OK;
}

// Tile shape, pipeline depth and tensor maps of one layer.  one_wave: the layer is one of a multi-layer program whose
// layers are separated by grid-wide barriers -- prefer a tile count just below the SM count (a second, partial wave
// doubles the layer's latency) over fewer weight re-reads.
int halo_configure(const PPConvParams& pin, HaloParams& h, bool one_wave) {
  h.c = pin;
  PPConvParams& p = h.c;
  int num_sms = 0;
  PP_TRY(halo_num_sms(&num_sms));
  const bool flat = p.kh * p.kw == 1;
  // N tile: <= 128 columns (two accumulator sets x two sub-tiles fill the 512 TMEM columns)
  const int n_tiles0 = pp_ceil_div(p.Cout_g_pad, 128);
  int bn = pp_ceil_div(pp_ceil_div(p.Cout_g_pad, n_tiles0), 16) * 16;
  const int tiles_y = flat ? 1 : pp_ceil_div(p.OH, 16);
  auto tiles_x = [&](int mt) { return flat ? (int)pp_ceil_div64(p.M_total, 128 * mt) : pp_ceil_div(p.OW, 8 * mt); };
  const int n_img = flat ? 1 : p.N;
  auto count = [&](int mt, int bn_) {
    return (long long)pp_ceil_div(p.Cout_g_pad, bn_) * tiles_x(mt) * tiles_y * n_img * p.groups;
  };
  int mt = 2;
  if (!p.ups2x && count(2, bn) < num_sms) mt = 1;
  while (count(mt, bn) < num_sms && bn >= 64 && bn % 32 == 0) bn /= 2;   // small launches: more, narrower tiles
  if (one_wave && !p.ups2x) {
    // largest tile count that still fits one wave: 128-pixel tiles, N split into 1..8 tiles of <= 256 columns
    static int min_bn = -1;
    if (min_bn < 0) { const char* e = getenv("PP_PROG_MIN_BN"); min_bn = e != nullptr ? atoi(e) : 16; }
    for (int nt = 8; nt >= 1; --nt) {
      const int b = pp_ceil_div(pp_ceil_div(p.Cout_g_pad, nt), 16) * 16;
      if (b > 256 || b < 16 || (b < min_bn && nt > 1)) continue;
      if (count(1, b) <= num_sms) { mt = 1; bn = b; break; }
    }
  }
  p.BN = bn;
  h.MT = mt;
  h.flat = flat ? 1 : 0;
  h.n_img = n_img;
  h.BW = flat ? 128 * mt : 8 * mt + (p.kw - 1) * p.dw;
  h.BH = flat ? 1 : 16 + (p.kh - 1) * p.dh;
  h.sub_bytes = flat ? 128 * 128 : 8 * 128;
  h.tiles_x = tiles_x(mt);
  h.tiles_y = tiles_y;
  h.n_tiles = pp_ceil_div(p.Cout_g_pad, bn);
  h.chunks = pp_ceil_div(p.Cin, 64);
  h.accw = pp_ceil_div(bn, 32) * 32;
  h.a_stage_bytes = pp_ceil_div(h.BW * h.BH * 128, 1024) * 1024;
  {
    // narrow N tiles: several filter taps per weight stage (<= 16 KB), see HaloParams::tps
    static int max_tps = -1;
    if (max_tps < 0) { const char* e = getenv("PP_HALO_TPS"); max_tps = e != nullptr ? atoi(e) : 9; }
    int tps = 16384 / (bn * 128);
    if (tps > p.kh * p.kw) tps = p.kh * p.kw;
    if (tps > max_tps) tps = max_tps;
    if (tps < 1 || p.ups2x) tps = 1;
    h.tps = tps;
  }
  h.b_stage_bytes = h.tps * bn * 128;
  h.ups = p.ups2x ? 1 : 0;
  h.LH = p.H / 2; h.LW = p.W / 2;
  h.LBW = (h.BW - 1) / 2 + 3; h.LBH = (h.BH - 1) / 2 + 3;      // low-res pixels that can feed BW x BH hi-res ones
  h.l_stage_bytes = h.ups ? pp_ceil_div(h.LBW * h.LBH * 128, 1024) * 1024 : 0;
  // TMA-store epilogue (see HaloParams::tstore): flat layers with few K chunks, plain fp16 output
  h.tstore = 0;
  h.out_stage_bytes = 0;
  {
    static int ts_on = -1;
    if (ts_on < 0) { const char* e = getenv("PP_TMA_STORE"); ts_on = (e == nullptr || atoi(e) != 0) ? 1 : 0; }
    if (ts_on && !one_wave && flat && !h.ups && p.epi == PP_EPI_STD && !p.out_fp32 && p.groups == 1 && p.out_gstep == 0 &&
        bn % 64 == 0 && p.vec_ok && h.chunks <= 16 && p.out_cstride % 8 == 0 && p.out_coff % 8 == 0 &&
        (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
      h.tstore = 1;
      h.out_stage_bytes = (bn / 64) * 16384;
    }
  }
  const int budget = SMEM_BUDGET - 2 * h.l_stage_bytes - (h.ups ? 1024 : 0) - (h.tstore ? mt * h.out_stage_bytes + 1024 : 0);
  int sa = h.ups ? 2 : 3, sb = 0;
  for (; sa >= 2; --sa) {
    sb = (budget - sa * h.a_stage_bytes) / h.b_stage_bytes;
    if (sb >= 3) break;
  }
  if (h.tstore && !(sa >= 2 && sb >= 3)) {      // no room for the staging tile: plain epilogue
    h.tstore = 0;
    h.out_stage_bytes = 0;
    const int budget2 = SMEM_BUDGET - 2 * h.l_stage_bytes - (h.ups ? 1024 : 0);
    for (sa = 3; sa >= 2; --sa) {
      sb = (budget2 - sa * h.a_stage_bytes) / h.b_stage_bytes;
      if (sb >= 3) break;
    }
  }
  PP_REQUIRE(sa >= 2 && sb >= 3, "conv_halo: patch %dx%d does not fit shared memory", h.BW, h.BH);
  if (sb > MAX_SB) sb = MAX_SB;
  if (!h.ups && sa == 3 && sb == MAX_SB && (budget - 4 * h.a_stage_bytes) / h.b_stage_bytes >= MAX_SB) sa = 4;
  h.SA = sa; h.SB = sb;
  if (h.tstore) {
    cuuint64_t dims[2] = {(cuuint64_t)p.Cout_g, (cuuint64_t)p.M_total};
    cuuint64_t strides[1] = {(cuuint64_t)p.out_cstride * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    EncodeTiledFn enc_o = encode_fn();
    PP_REQUIRE(enc_o != nullptr, "conv_halo: cuTensorMapEncodeTiled is not available");
    const CUresult r = enc_o(&h.tmap_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, reinterpret_cast<__half*>(p.out) + p.out_coff, dims,
                             strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PP_REQUIRE(r == CUDA_SUCCESS, "conv_halo: cuTensorMapEncodeTiled (output) failed (%d)", (int)r);
  }
  { const char* e = getenv("PP_CONV_NOEPI"); h.debug = (e != nullptr && atoi(e) != 0) ? 1 : 0; }
  const long long total_tiles = count(mt, bn);
  PP_REQUIRE(total_tiles < (1LL << 31), "conv_halo: too many tiles");

  EncodeTiledFn enc = encode_fn();
  PP_REQUIRE(enc != nullptr, "conv_halo: cuTensorMapEncodeTiled is not available");
  for (int i = 0; i < p.nseg; ++i) {
    const PPConvSeg& s = p.seg[i];
    const cuuint64_t cacc = (cuuint64_t)(p.groups - 1) * s.gstep + (s.cvalid > 0 ? s.cvalid : s.cend - s.cbegin);
    cuuint64_t dims[4] = {cacc, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
    cuuint64_t strides[3] = {(cuuint64_t)s.cstride * 2, (cuuint64_t)p.W * s.cstride * 2, (cuuint64_t)p.H * p.W * s.cstride * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)h.BW, (cuuint32_t)h.BH, 1};
    if (flat) {   // pixels as one flat dimension; the last tile's tail is out of bounds -> zero-filled
      dims[1] = (cuuint64_t)p.M_total; dims[2] = 1; dims[3] = 1;
      strides[1] = strides[2] = (cuuint64_t)p.M_total * s.cstride * 2;
    }
    if (h.ups) {  // the tensor in memory is the half-resolution source; it lands unswizzled in the staging ring
      dims[1] = (cuuint64_t)h.LW; dims[2] = (cuuint64_t)h.LH;
      strides[1] = (cuuint64_t)h.LW * s.cstride * 2; strides[2] = (cuuint64_t)h.LH * h.LW * s.cstride * 2;
      box[1] = (cuuint32_t)h.LBW; box[2] = (cuuint32_t)h.LBH;
    }
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = enc(&h.tmap[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(s.ptr + s.coff), dims, strides, box,
                           es, CU_TENSOR_MAP_INTERLEAVE_NONE, h.ups ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PP_REQUIRE(r == CUDA_SUCCESS, "conv_halo: cuTensorMapEncodeTiled failed (%d) for segment %d (cstride=%d W=%d H=%d N=%d)",
               (int)r, i, s.cstride, p.W, p.H, p.N);
  }
  return PP_OK;
}

inline long long halo_total_tiles(const HaloParams& h) {
  return (long long)h.n_tiles * h.tiles_x * h.tiles_y * h.n_img * h.c.groups;
}

}  // namespace

int pp_launch_conv_halo(const PPConvParams& pin, cudaStream_t stream) {
  HaloParams h;
  PP_TRY(halo_configure(pin, h, false));
  int num_sms = 0;
  PP_TRY(halo_num_sms(&num_sms));
  const long long total_tiles = halo_total_tiles(h);
  const size_t smem = (size_t)h.SA * h.a_stage_bytes + (size_t)h.SB * h.b_stage_bytes + 1024 + 512 + 2 * (size_t)h.l_stage_bytes +
                      (h.ups ? 1024 : 0) + (h.tstore ? (size_t)h.MT * h.out_stage_bytes + 1536 : 0);
  const int grid = (int)(total_tiles < num_sms ? total_tiles : num_sms);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(h.ups ? UPS_THREADS : NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (h.ups) PP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_halo_kernel<true>, h));
  else PP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_halo_kernel<false>, h));
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}

// ---- multi-layer programs ------------------------------------------------------------------------------------------
namespace {
thread_local PPProgRecorder* g_recorder = nullptr;
}

struct PPProgRecorder {
  ProgParams prog;
  double flops = 0.0;
  int n_conv = 0;
};

bool pp_prog_recording() { return g_recorder != nullptr; }

int pp_prog_begin() {
  PP_REQUIRE(g_recorder == nullptr, "conv program: already recording");
  g_recorder = new PPProgRecorder();
  memset(&g_recorder->prog, 0, sizeof(ProgParams));
  return PP_OK;
}

void pp_prog_abort() {
  delete g_recorder;
  g_recorder = nullptr;
}

int pp_prog_eligible(const PPConvParams& p) { return pp_conv_halo_eligible(p) && !p.ups2x; }

int pp_prog_record_conv(const PPConvParams& p) {
  PPProgRecorder* r = g_recorder;
  PP_REQUIRE(r != nullptr, "conv program: not recording");
  PP_REQUIRE(r->prog.n_layers < PROG_MAX_LAYERS, "conv program: more than %d layers", PROG_MAX_LAYERS);
  PP_REQUIRE(pp_prog_eligible(p), "conv program: layer is not a stride-1 TMA halo-kernel convolution");
  const int li = r->prog.n_layers;
  PP_TRY(halo_configure(p, r->prog.layer[li], true));
  r->prog.kind[li] = PROG_CONV;
  r->prog.n_layers++;
  r->n_conv++;
  return PP_OK;
}

int pp_prog_record_dcn(const PPDcnArgs& a) {
  PPProgRecorder* r = g_recorder;
  PP_REQUIRE(r != nullptr, "conv program: not recording");
  PP_REQUIRE(r->prog.n_layers < PROG_MAX_LAYERS, "conv program: more than %d layers", PROG_MAX_LAYERS);
  const int li = r->prog.n_layers;
  r->prog.kind[li] = PROG_DCN;
  r->prog.dcn[li] = a;
  r->prog.n_layers++;
  return PP_OK;
}

// Launches the recorded layers as ONE kernel (grid = one CTA per SM); `counter` is a zero-initialised device word shared
// by all programs of a stream, `*arrivals` the host-side count of arrivals issued so far on it.
int pp_prog_end(unsigned int* counter, unsigned int* arrivals, cudaStream_t stream) {
  PPProgRecorder* r = g_recorder;
  PP_REQUIRE(r != nullptr, "conv program: not recording");
  g_recorder = nullptr;
  std::unique_ptr<PPProgRecorder> guard(r);
  ProgParams& P = r->prog;
  if (P.n_layers == 0) return PP_OK;
  int num_sms = 0;
  PP_TRY(halo_num_sms(&num_sms));
  int a_max = 1024, b_max = 2048;
  for (int i = 0; i < P.n_layers; ++i)
    if (P.kind[i] == PROG_CONV) {
      if (P.layer[i].a_stage_bytes > a_max) a_max = P.layer[i].a_stage_bytes;
      if (P.layer[i].b_stage_bytes > b_max) b_max = P.layer[i].b_stage_bytes;
    }
  int sa = 3, sb = 0;
  for (; sa >= 2; --sa) {
    sb = (SMEM_BUDGET - sa * a_max) / b_max;
    if (sb >= 3) break;
  }
  PP_REQUIRE(sa >= 2 && sb >= 3, "conv program: stages do not fit shared memory (A %d B, B %d B)", a_max, b_max);
  if (sb > MAX_SB) sb = MAX_SB;
  P.SA = sa; P.SB = sb; P.a_stage_bytes = a_max; P.b_stage_bytes = b_max;
  P.counter = counter;
  P.base = *arrivals;
  static int ts_mode = -1, ts_printed = 0;
  static unsigned long long* ts_dev = nullptr;
  if (ts_mode < 0) {
    const char* e = getenv("PP_PROG_TS");
    ts_mode = (e != nullptr && atoi(e) != 0) ? atoi(e) : 0;
    if (ts_mode) PP_CUDA_CHECK(cudaMalloc(&ts_dev, 2 * PROG_MAX_LAYERS * sizeof(unsigned long long)));
  }
  P.ts = (ts_mode && ts_printed < ts_mode) ? ts_dev : nullptr;
  const int grid = num_sms;
  *arrivals += (unsigned int)(P.n_layers * grid);
  const size_t smem = (size_t)sa * a_max + (size_t)sb * b_max + 1024 + 512;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_prog_kernel, P));
  PP_CUDA_CHECK(cudaGetLastError());
  if (P.ts != nullptr) {      // debug: per-layer wall time of CTA 0 (serialises the stream)
    unsigned long long h[2 * PROG_MAX_LAYERS];
    PP_CUDA_CHECK(cudaStreamSynchronize(stream));
    PP_CUDA_CHECK(cudaMemcpy(h, ts_dev, sizeof(h), cudaMemcpyDeviceToHost));
    ++ts_printed;
    fprintf(stderr, "[prog %d] %d layers:", ts_printed, P.n_layers);
    for (int i = 0; i < P.n_layers; ++i) {
      const HaloParams& L = P.layer[i];
      if (P.kind[i] == PROG_CONV)
        fprintf(stderr, " | conv K=%d N=%d bn=%d mt=%d tiles=%lld: %.1f us (gap %.1f)", L.c.K_total, L.c.Cout_g, L.c.BN, L.MT,
                halo_total_tiles(L), (h[2 * i + 1] - h[2 * i]) / 1e3, i ? (h[2 * i] - h[2 * i - 1]) / 1e3 : 0.0);
      else
        fprintf(stderr, " | dcn: %.1f us (gap %.1f)", (h[2 * i + 1] - h[2 * i]) / 1e3, i ? (h[2 * i] - h[2 * i - 1]) / 1e3 : 0.0);
    }
    fprintf(stderr, " | total %.1f us\n", (h[2 * P.n_layers - 1] - h[0]) / 1e3);
  }
  return PP_OK;
}
