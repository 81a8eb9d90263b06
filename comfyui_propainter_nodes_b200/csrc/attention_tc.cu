// Masked-window attention on tcgen05 tensor cores (SparseWindowAttention, sparse_transformer.py:327-357).
//
// One CTA (256 threads, two per SM) per (128-query tile, 5x9 window, head, sliding window).  Per 64-key tile:
//   S = Q K^T       tcgen05.mma  M=128 (queries) N=64 (keys) K=128 (d)    -> TMEM columns [0,64)
//   softmax         threads r and r+128 own query row r (= TMEM lane r), 32 key columns each: tcgen05.ld, online
//                   max/sum in registers (one shared-memory exchange of the row max per tile, no shuffles),
//                   P (fp16) written to shared memory as the next A operand
//   O += P V        tcgen05.mma  M=128 N=128 (d) K=64 (keys), V as the MN-major B operand  -> TMEM [128,256)
// O stays in TMEM for the whole key loop.  The running max is only raised when it grew by more than 8 (log2
// units), in which case the O rows are rescaled in TMEM (tcgen05.ld/st); softmax is invariant to that shift, so
// the result is exact while P stays within fp16 range (<= 2^8).
// Keys are gathered by index (own 45 + ring 148 + pooled tokens of every 2nd frame) with 16-byte cp.async into
// 128B-swizzled panels -- the window/rolled/pooled K,V tensors of the reference are never materialised.
#include "attention.cuh"

namespace {

constexpr int D = 128, BQ = 128, BKEY = 64, NT = 256, WIN_TOK = 45, RING = 193;
constexpr int CPT = BKEY / 2;
constexpr int TMEM_COLS = 256;                 // S: BKEY columns at 0, O: 128 columns at O_COL
constexpr int O_COL = 128;                   // key columns per thread (two threads per query row)
constexpr uint32_t QPANEL = BQ * 128;           // bytes of a [128 rows][64 halves] panel (Q, P)
constexpr uint32_t KPANEL = BKEY * 128;         // bytes of a [BKEY rows][64 halves] panel (K, V)
constexpr uint32_t QTILE = 2 * QPANEL;          // Q: [128][128 d]
constexpr uint32_t KTILE = 2 * KPANEL;          // K or V: [BKEY][128 d]
constexpr uint32_t PTILE = (BKEY / 64) * QPANEL;  // P: [128][BKEY keys]
// 112 KiB per CTA so that two CTAs share an SM: one CTA's softmax overlaps the other's MMAs and gathers
constexpr uint32_t SM_Q = 0, SM_K = QTILE, SM_V = SM_K + 2 * KTILE, SM_P = SM_V + 2 * KTILE, SM_END = SM_P + PTILE;

// byte offset of 16-byte chunk `chunk` (along the 64-wide panels) of row `row` in a tile whose panels are `panel` bytes
__device__ __forceinline__ uint32_t tile_off(uint32_t panel, int row, int chunk) {
  return (uint32_t)(chunk >> 3) * panel + (uint32_t)row * 128 + (uint32_t)(((chunk & 7) ^ (row & 7)) << 4);
}

// Key-row gather table: entry j of window `win` is the source row of key j of a masked window, as an offset (in
// 16-byte units, relative to the first frame of the sliding window) into the K/V token tensor, or -- top bit set --
// into the pooled K/V tensor.  Keys of T_ind frame fi (frame parity + 2*fi) are [own 45 | ring 148 | pooled n_pool].
// Built once per launch so that the gather loop of the attention kernel does no integer division or index math.
__global__ void attn_key_table(int* __restrict__ tab, const int* __restrict__ ring_idx, int nk_max, int kpf, int ntok,
                               int n_pool, int qkv_cs, int pool_cs, int parity) {
  const int win = blockIdx.x;
  const int* ring = ring_idx + win * RING;
  for (int j = threadIdx.x; j < nk_max; j += blockDim.x) {
    const int fi = j / kpf, w = j - fi * kpf;
    const int fr = parity + 2 * fi;
    unsigned e;
    if (w < RING) e = (unsigned)(((long long)fr * ntok + ring[w]) * qkv_cs / 8);
    else e = 0x80000000u | (unsigned)(((long long)fr * n_pool + (w - RING)) * pool_cs / 8);
    tab[(long long)win * nk_max + j] = (int)e;
  }
}

__global__ void __launch_bounds__(NT, 2) window_attention_tc(const PPAttnParams p) {
  using namespace ppx;
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzled tiles need 1024-byte alignment
  const uint32_t sbase = smem_u32(smem);
  uint64_t* mbar_s = reinterpret_cast<uint64_t*>(smem + SM_END);
  uint64_t* mbar_o = mbar_s + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar_o + 1);
  __half* xmax = reinterpret_cast<__half*>(smem + SM_END + 64);   // [2][128] per-tile row-max exchange
  float* xsum = reinterpret_cast<float*>(smem + SM_Q);            // [2][128] row-sum exchange (Q tile is dead by then)

  const int win = blockIdx.y >> 2, head = blockIdx.y & 3, sw = blockIdx.z;
  if (p.win_flags[sw * p.n_win + win] == 0) return;        // unmasked windows: mma.sync kernel
  const int t = p.sw_t[sw];
  const int frame_base = p.sw_frame_off[sw];
  const int nq = t * WIN_TOK;
  const int q0 = blockIdx.x * BQ;
  if (q0 >= nq) return;
  const int n_tind = (t - p.parity + 1) / 2;
  const int kpf = RING + p.n_pool;
  const int nk = n_tind * kpf;
  const int ntiles = (nk + BKEY - 1) / BKEY;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int* ring = p.ring_idx + win * RING;
  const long long ntok = (long long)p.nh * p.nw;

  if (tid == 0) {
    mbar_init(mbar_s, 1);
    mbar_init(mbar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }

  // ---- Q tile (rows beyond nq are clamped to a valid query; never stored)
  for (int i = tid; i < BQ * 16; i += NT) {
    const int r = i >> 4, ch = i & 15;
    const int qi = min(q0 + r, nq - 1);
    const int fr = frame_base + qi / WIN_TOK, pos = qi % WIN_TOK;
    const __half* src = p.q + ((long long)fr * ntok + ring[pos]) * p.qkv_cs + head * D + ch * 8;
    cp_async16(sbase + SM_Q + tile_off(QPANEL, r, ch), src, 16);
  }
  // K/V gather: thread owns 16-byte chunk `ch` of rows r0, r0+16, r0+32, r0+48 of every key tile; the source row of
  // key j comes from the per-window table (attn_key_table), so the loop body is a table load, two adds and two copies
  const int r0 = tid >> 4, ch = tid & 15;
  const long long fb = frame_base;
  const __half* kb = p.k + fb * ntok * p.qkv_cs + head * D + ch * 8;
  const __half* vb = p.v + fb * ntok * p.qkv_cs + head * D + ch * 8;
  const __half* pkb = p.pk + fb * p.n_pool * p.pool_cs + head * D + ch * 8;
  const __half* pvb = p.pv + fb * p.n_pool * p.pool_cs + head * D + ch * 8;
  const int* ktab = p.key_tab + (long long)win * p.key_tab_stride;
  const uint32_t kv_dst0 = sbase + tile_off(KPANEL, r0, ch);
  auto load_kv = [&](int tile, int stage) {
    const uint32_t dk = kv_dst0 + SM_K + stage * KTILE, dv = kv_dst0 + SM_V + stage * KTILE;
#pragma unroll
    for (int it = 0; it < BKEY / 16; ++it) {
      const int j = tile * BKEY + r0 + 16 * it;
      const __half* ks = kb; const __half* vs = vb;
      uint32_t nbytes = 0;
      if (j < nk) {
        nbytes = 16;
        const int e = __ldg(ktab + j);
        const long long off = (long long)(e & 0x7fffffff) * 8;
        ks = (e < 0 ? pkb : kb) + off;
        vs = (e < 0 ? pvb : vb) + off;
      }
      cp_async16(dk + it * 2048, ks, nbytes);     // rows r0 + 16*it: same swizzle phase, 16 rows * 128 B further
      cp_async16(dv + it * 2048, vs, nbytes);
    }
    cp_async_commit();
  };
  load_kv(0, 0);   // one group: Q + first K/V tile

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int row = tid & 127, half = tid >> 7;                             // query row, which 64 key/d columns
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);   // this thread's TMEM lane (= query row)
  const uint32_t idesc_s = umma_idesc_f16(128, BKEY);
  const uint32_t idesc_o = umma_idesc_f16_bmn(128, 128);

  float m_used = 0.f, m_run = -1e30f, row_sum = 0.f;
  for (int j = 0; j < ntiles; ++j) {
    const int stage = j & 1;
    cp_async_wait<0>();
    fence_proxy_async();
    __syncthreads();
    if (warp == 0 && elect_one()) {   // S = Q K^T (one elected lane: ptxas emits each UTCHMMA once)
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t sub = (uint32_t)(ks & 3) * 32;
        umma_f16(tmem_base, umma_desc_sw128_kmajor(sbase + SM_Q + (uint32_t)(ks >> 2) * QPANEL + sub),
                 umma_desc_sw128_kmajor(sbase + SM_K + stage * KTILE + (uint32_t)(ks >> 2) * KPANEL + sub), idesc_s,
                 ks != 0 ? 1u : 0u);
      }
      umma_commit(mbar_s);
    }
    // previous P.V must be done before its K/V stage and the P buffer are overwritten
    if (j > 0) { mbar_wait(mbar_o, (uint32_t)(j - 1) & 1u); tc_fence_after(); }
    if (j + 1 < ntiles) load_kv(j + 1, stage ^ 1);
    mbar_wait(mbar_s, (uint32_t)j & 1u);
    tc_fence_after();

    // ---- this thread's CPT columns of its S row
    float s[CPT];
#pragma unroll
    for (int c = 0; c < CPT / 16; ++c) {
      uint32_t raw[16];
      tmem_ld16(lane_addr + half * CPT + c * 16, raw);
#pragma unroll
      for (int i = 0; i < 16; ++i) s[c * 16 + i] = __uint_as_float(raw[i]);
    }
    tmem_ld_wait();
    const int kvalid = min(BKEY, nk - j * BKEY) - half * CPT;   // valid columns among this thread's
    if (kvalid < CPT) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) if (i >= kvalid) s[i] = -1e30f;
    }
    float mx[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
    for (int i = 0; i < CPT; i += 4) {
      mx[0] = fmaxf(mx[0], s[i]); mx[1] = fmaxf(mx[1], s[i + 1]); mx[2] = fmaxf(mx[2], s[i + 2]); mx[3] = fmaxf(mx[3], s[i + 3]);
    }
    // both threads of a row read the same two (fp16-rounded) values, so they agree on the reference max; any
    // reference works for softmax as long as numerator and denominator share it
    xmax[half * 128 + row] = __float2half_rn(fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * p.scale_log2, -60000.f));
    __syncthreads();
    const float m_tile = fmaxf(__half2float(xmax[row]), __half2float(xmax[128 + row]));
    const float m_new = fmaxf(m_run, m_tile);
    int need = 0;
    if (j == 0) m_used = m_new;
    else need = m_new > m_used + 8.f;
    m_run = m_new;
    if (__syncthreads_or(need)) {
      // rare: raise the reference max of the rows that need it and rescale their O rows in TMEM
      const float f = need ? exp2f(m_used - m_new) : 1.f;
      if (need) { m_used = m_new; row_sum *= f; }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[16];
        tmem_ld16(lane_addr + O_COL + half * 64 + c * 16, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * f);
        tmem_st16(lane_addr + O_COL + half * 64 + c * 16, raw);
      }
      tmem_st_wait();
    }
    // ---- P = exp2(s*scale - m_used) -> shared memory (A operand of P.V): this thread's 16-byte chunks
    float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ch = 0; ch < CPT / 8; ++ch) {
      __align__(16) __half2 h[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = exp2f(fmaf(s[ch * 8 + 2 * e], p.scale_log2, -m_used));
        const float b = exp2f(fmaf(s[ch * 8 + 2 * e + 1], p.scale_log2, -m_used));
        ps[e] += a + b;
        h[e] = __floats2half2_rn(a, b);
      }
      *reinterpret_cast<uint4*>(smem + SM_P + tile_off(QPANEL, row, half * (CPT / 8) + ch)) = *reinterpret_cast<uint4*>(h);
    }
    row_sum += (ps[0] + ps[1]) + (ps[2] + ps[3]);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) {   // O += P V
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < BKEY / 16; ++ks) {
        const uint64_t adesc = umma_desc_sw128_kmajor(sbase + SM_P + (uint32_t)(ks >> 2) * QPANEL + (uint32_t)(ks & 3) * 32);
        const uint64_t bdesc = umma_desc_sw128_mnmajor(sbase + SM_V + stage * KTILE + (uint32_t)ks * 2048, KPANEL);
        umma_f16(tmem_base + O_COL, adesc, bdesc, idesc_o, (j | ks) != 0 ? 1u : 0u);
      }
      umma_commit(mbar_o);
    }
  }

  // ---- epilogue: O / row_sum -> global (unpadded grid; padding queries are dropped)
  mbar_wait(mbar_o, (uint32_t)(ntiles - 1) & 1u);
  tc_fence_after();
  xsum[half * 128 + row] = row_sum;
  __syncthreads();
  const float inv = 1.f / (xsum[row] + xsum[128 + row]);
  const int qi = q0 + row;
  bool store = qi < nq;
  __half* dst = nullptr;
  if (store) {
    const int fr = frame_base + qi / WIN_TOK, pos = qi % WIN_TOK;
    const int tok = ring[pos];
    const int ty = tok / p.nw, tx = tok - ty * p.nw;
    store = ty < p.gh && tx < p.gw;
    dst = p.out + (((long long)fr * p.gh + ty) * p.gw + tx) * p.out_cs + head * D + half * 64;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t raw[16];
    tmem_ld16(lane_addr + O_COL + half * 64 + c * 16, raw);
    tmem_ld_wait();
    if (store) {
      __align__(16) __half2 h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        h[i] = __floats2half2_rn(__uint_as_float(raw[2 * i]) * inv, __uint_as_float(raw[2 * i + 1]) * inv);
      reinterpret_cast<uint4*>(dst + c * 16)[0] = reinterpret_cast<uint4*>(h)[0];
      reinterpret_cast<uint4*>(dst + c * 16)[1] = reinterpret_cast<uint4*>(h)[1];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int pp_launch_attention_tc(const PPAttnParams& p, int n_sliding, int t_max, cudaStream_t st) {
  const size_t smem = SM_END + 64 + 2 * 128 * sizeof(__half);   // 115,264 B: two CTAs per SM
  static bool attr_set = false;
  if (!attr_set) {
    PP_CUDA_CHECK(cudaFuncSetAttribute(window_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PP_CUDA_CHECK(cudaFuncSetAttribute(window_attention_tc, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  {
    const int kpf = RING + p.n_pool, nk_max = ((t_max - p.parity + 1) / 2) * kpf;
    PP_REQUIRE(p.key_tab != nullptr && p.key_tab_stride >= nk_max, "attention: key table scratch too small (%d < %d)",
               p.key_tab_stride, nk_max);
    attn_key_table<<<p.n_win, 256, 0, st>>>(const_cast<int*>(p.key_tab), p.ring_idx, p.key_tab_stride, kpf, p.nh * p.nw, p.n_pool,
                                           p.qkv_cs, p.pool_cs, p.parity);
    PP_CUDA_CHECK(cudaGetLastError());
  }
  dim3 grid(pp_ceil_div(t_max * WIN_TOK, BQ), p.n_win * 4, n_sliding);
  window_attention_tc<<<grid, NT, smem, st>>>(p);
  PP_CUDA_CHECK(cudaGetLastError());
  return PP_OK;
}
