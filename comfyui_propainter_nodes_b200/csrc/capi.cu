// extern "C" surface of libpropainter_b200.so (declared in include/propainter_b200.h).
#include <math.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/propainter_b200.h"
#include "dcn_sample.cuh"
#include "engine.cuh"

// Makes the engine's device current for the duration of one API call and restores the caller's device afterwards.
struct DeviceGuard {
  int prev = -1;
  bool changed = false, ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) {
      ok = cudaSetDevice(dev) == cudaSuccess;
      changed = ok;
    }
  }
  ~DeviceGuard() {
    if (changed && prev >= 0) cudaSetDevice(prev);
  }
};

#define PP_HANDLE(h)                                     \
  if ((h) == nullptr) {                                  \
    pp_set_error("null engine handle");                  \
    return PP_ERR_ARG;                                   \
  }                                                      \
  PPEngine& e = *reinterpret_cast<PPEngine*>(h);         \
  DeviceGuard _dev_guard(e.device);                      \
  if (!_dev_guard.ok) {                                  \
    pp_set_error("cudaSetDevice(%d) failed", e.device);  \
    return PP_ERR_CUDA;                                  \
  }

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Restores the arena to its state at entry on EVERY exit of a stage call (error paths included): a failed call
// ("workspace exhausted", bad argument after a partial allocation) must not shrink the workspace of the cached engine.
struct ArenaGuard {
  PPArena& a;
  size_t m;
  bool keep = false;
  explicit ArenaGuard(PPArena& arena) : a(arena), m(arena.mark()) {}
  ~ArenaGuard() { if (!keep) a.release(m); }
};

extern "C" {

const char* pp_version(void) { return "propainter_b200 1 sm_100a"; }

int pp_create(int device, void* workspace, size_t workspace_bytes, pp_handle* out) {
  PP_REQUIRE(out != nullptr, "pp_create: out is null");
  PP_REQUIRE(workspace != nullptr && workspace_bytes >= (64u << 20), "pp_create: workspace must be >= 64 MiB");
  PP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "pp_create: workspace must be 256-byte aligned");
  PP_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  PP_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  PP_REQUIRE(prop.major == 10, "pp_create: this library is built for sm_100a (Blackwell B200); device %d is sm_%d%d",
             device, prop.major, prop.minor);
  PPEngine* e = new PPEngine();
  e->device = device;
  e->arena.base = static_cast<uint8_t*>(workspace);
  e->arena.cap = workspace_bytes;
  if (cudaMalloc(&e->prog_counter, 256) != cudaSuccess || cudaMemset(e->prog_counter, 0, 256) != cudaSuccess) {
    delete e;
    pp_set_error("pp_create: cudaMalloc of the program barrier word failed");
    return PP_ERR_CUDA;
  }
  *out = reinterpret_cast<pp_handle>(e);
  return PP_OK;
}

int pp_set_workspace(pp_handle h, void* workspace, size_t workspace_bytes) {
  PP_HANDLE(h);
  PP_REQUIRE(!e.gen.active, "pp_set_workspace: a generator session is active (call pp_gen_end first)");
  PP_REQUIRE(e.arena.off == 0, "pp_set_workspace: the arena is in use");
  PP_REQUIRE(workspace != nullptr && workspace_bytes >= (64u << 20), "pp_set_workspace: workspace must be >= 64 MiB");
  PP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "pp_set_workspace: workspace must be 256-byte aligned");
  e.arena.base = static_cast<uint8_t*>(workspace);
  e.arena.cap = workspace_bytes;
  e.arena.peak = 0;
  return PP_OK;
}

int pp_destroy(pp_handle h) {
  if (h != nullptr) {
    PPEngine* e = reinterpret_cast<PPEngine*>(h);
    pp_comm_destroy_impl(*e);
    if (e->prog_counter != nullptr) cudaFree(e->prog_counter);
    delete e;
  }
  return PP_OK;
}

int pp_comm_unique_id(void* out128) {
  PP_REQUIRE(out128 != nullptr, "pp_comm_unique_id: null pointer");
  return pp_comm_unique_id_impl(out128);
}

int pp_comm_init(pp_handle h, const void* unique_id128, int rank, int world) {
  PP_HANDLE(h);
  PP_REQUIRE(unique_id128 != nullptr, "pp_comm_init: null id");
  return pp_comm_init_impl(e, unique_id128, rank, world);
}

int pp_comm_destroy(pp_handle h) {
  PP_HANDLE(h);
  return pp_comm_destroy_impl(e);
}

int pp_comm_all_gather_rows(pp_handle h, void* buf, const long long* rows_per_member, size_t row_bytes, int first_rank,
                            int n_members, void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(buf != nullptr && rows_per_member != nullptr && row_bytes > 0 && n_members >= 1,
             "pp_comm_all_gather_rows: bad argument");
  std::vector<long long> off(n_members, 0);
  for (int m = 1; m < n_members; ++m) off[m] = off[m - 1] + rows_per_member[m - 1];
  return pp_comm_all_gather_blocks_impl(e, buf, off.data(), rows_per_member, row_bytes, first_rank, n_members,
                                        as_stream(stream));
}

int pp_register_conv(pp_handle h, const char* name, const void* w, const float* bias, int cout_g, int cout_g_pad,
                     int bn, int cin_g, int kh, int kw, int groups) {
  PP_HANDLE(h);
  PP_REQUIRE(name != nullptr && w != nullptr, "pp_register_conv: null argument");
  PP_REQUIRE(bn % 16 == 0 && bn >= 16 && bn <= 256 && cout_g_pad % bn == 0 && cout_g <= cout_g_pad && cin_g % 8 == 0,
             "pp_register_conv(%s): invalid packing (cout_g=%d pad=%d bn=%d cin_g=%d)", name, cout_g, cout_g_pad, bn,
             cin_g);
  PPPackedConv c;
  c.w = static_cast<const __half*>(w); c.b = bias;
  c.cout_g = cout_g; c.cout_g_pad = cout_g_pad; c.bn = bn; c.cin_g = cin_g; c.kh = kh; c.kw = kw; c.groups = groups;
  e.convs[name] = c;
  return PP_OK;
}

int pp_set_conv_macs(pp_handle h, const char* name, double macs_per_pixel) {
  PP_HANDLE(h);
  PP_REQUIRE(name != nullptr, "pp_set_conv_macs: null name");
  auto it = e.convs.find(name);
  PP_REQUIRE(it != e.convs.end(), "pp_set_conv_macs: conv '%s' is not registered", name);
  it->second.macs_per_pixel = macs_per_pixel;
  return PP_OK;
}

int pp_register_tensor(pp_handle h, const char* name, const void* ptr, size_t bytes) {
  PP_HANDLE(h);
  PP_REQUIRE(name != nullptr && ptr != nullptr, "pp_register_tensor: null argument");
  PPTensor t;
  t.ptr = ptr; t.bytes = bytes;
  e.tensors[name] = t;
  return PP_OK;
}

int pp_raft_bidir(pp_handle h, const float* frames, int T, int H, int W, int iters, float* flows_f, float* flows_b,
                  void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(frames && flows_f && flows_b, "pp_raft_bidir: null pointer");
  ArenaGuard guard(e.arena);
  return pp_stage_raft(e, frames, T, H, W, iters, flows_f, flows_b, as_stream(stream));
}

int pp_flow_complete(pp_handle h, const float* flows_f, const float* flows_b, const float* flow_masks, int T, int H,
                     int W, float* out_f, float* out_b, void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(flows_f && flows_b && flow_masks && out_f && out_b, "pp_flow_complete: null pointer");
  ArenaGuard guard(e.arena);
  return pp_stage_flow_complete(e, flows_f, flows_b, flow_masks, T, H, W, out_f, out_b, 0, 1, as_stream(stream));
}

int pp_flow_complete_dist(pp_handle h, const float* flows_f, const float* flows_b, const float* flow_masks, int T, int H,
                          int W, float* out_f, float* out_b, int team_first, int team_size, void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(flows_f && flows_b && flow_masks && out_f && out_b, "pp_flow_complete_dist: null pointer");
  PP_REQUIRE(e.comm != nullptr || team_size <= 1, "pp_flow_complete_dist: pp_comm_init was not called");
  PP_REQUIRE(team_first >= 0 && team_size >= 1 && team_first + team_size <= e.world,
             "pp_flow_complete_dist: team [%d, %d) of %d ranks", team_first, team_first + team_size, e.world);
  ArenaGuard guard(e.arena);
  return pp_stage_flow_complete(e, flows_f, flows_b, flow_masks, T, H, W, out_f, out_b, team_first, team_size,
                                as_stream(stream));
}

int pp_image_propagate(pp_handle h, const float* frames, const float* masks, const float* flows_f,
                       const float* flows_b, int T, int H, int W, float* updated_frames, float* updated_masks,
                       void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(frames && masks && flows_f && flows_b && updated_frames && updated_masks,
             "pp_image_propagate: null pointer");
  ArenaGuard guard(e.arena);
  return pp_stage_image_propagate(e, frames, masks, flows_f, flows_b, T, H, W, updated_frames, updated_masks,
                                  as_stream(stream));
}

int pp_gen_begin_subset(pp_handle h, const float* updated_frames, const float* masks_dilated,
                        const float* updated_masks, const float* flows_f, const float* flows_b, int T, int H, int W,
                        const unsigned char* frames_needed, void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(updated_frames && masks_dilated && updated_masks && flows_f && flows_b, "pp_gen_begin: null pointer");
  const int r = pp_stage_gen_begin(e, updated_frames, masks_dilated, updated_masks, flows_f, flows_b, T, H, W,
                                   frames_needed, as_stream(stream));
  if (r != PP_OK) pp_stage_gen_end(e);   // a failed begin leaves no session and no allocation behind
  return r;
}

int pp_gen_begin(pp_handle h, const float* updated_frames, const float* masks_dilated, const float* updated_masks,
                 const float* flows_f, const float* flows_b, int T, int H, int W, void* stream) {
  return pp_gen_begin_subset(h, updated_frames, masks_dilated, updated_masks, flows_f, flows_b, T, H, W, nullptr, stream);
}

int pp_gen_window(pp_handle h, const int* frame_ids, int t, int l_t, void* pred_f16, void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(frame_ids && pred_f16, "pp_gen_window: null pointer");
  ArenaGuard guard(e.arena);
  return pp_stage_gen_window(e, frame_ids, t, l_t, static_cast<__half*>(pred_f16), as_stream(stream));
}

int pp_gen_run(pp_handle h, const int* frame_ids, const int* win_t, const int* win_lt, int n_windows, void* pred_f16,
               void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(frame_ids && win_t && win_lt && pred_f16, "pp_gen_run: null pointer");
  ArenaGuard guard(e.arena);
  return pp_stage_gen_run(e, frame_ids, win_t, win_lt, n_windows, static_cast<__half*>(pred_f16), as_stream(stream));
}

int pp_gen_end(pp_handle h) {
  PP_HANDLE(h);
  return pp_stage_gen_end(e);
}

int pp_composite(pp_handle h, const void* pred_f16, const float* masks_dilated, const uint8_t* orig, uint8_t* comp,
                 const int* frame_ids_dev, const int* first_visit_dev, int l_t, int H, int W, int half_math,
                 void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(pred_f16 && masks_dilated && orig && comp && frame_ids_dev && first_visit_dev, "pp_composite: null pointer");
  e.launches++;
  return pp_k_composite(static_cast<const __half*>(pred_f16), 4, masks_dilated, orig, comp, frame_ids_dev,
                        first_visit_dev, l_t, H, W, half_math, as_stream(stream));
}

int pp_preprocess(pp_handle h, const float* image, const float* mask, int mask_frames, int T, int H, int W,
                  int flow_mask_dilates, int mask_dilates, uint8_t* orig_u8, float* frames, float* flow_masks,
                  float* masks_dilated, void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(image && mask && orig_u8 && frames && flow_masks && masks_dilated, "pp_preprocess: null pointer");
  ArenaGuard guard(e.arena);
  uint8_t* scratch;
  PP_TRY(pp_alloc(e, &scratch, (size_t)mask_frames * H * W, "mask scratch"));
  PP_TRY(pp_k_quantize_frames(image, orig_u8, frames, T, H, W, as_stream(stream)));
  PP_TRY(pp_k_prepare_masks(mask, mask_frames, T, H, W, flow_mask_dilates, mask_dilates, scratch, flow_masks,
                            masks_dilated, as_stream(stream)));
  e.launches += 4;
  return PP_OK;
}

// Shared tail of the pre-processing entry points: 8-bit frames / masks at the input size -> outputs at the processing size
static int preprocess_from_u8(PPEngine& e, const uint8_t* u8_in, const uint8_t* m_in, int mask_frames, int T, int H, int W,
                              int out_h, int out_w, int flow_mask_dilates, int mask_dilates, uint8_t* orig_u8, float* frames,
                              float* flow_masks, float* masks_dilated, cudaStream_t st) {
  const bool resize = out_h != H || out_w != W;
  const uint8_t* m_use = m_in;
  if (resize) {
    const size_t mid_px = (size_t)T * H * out_w;
    uint8_t *tmp, *m_out;
    int* coef;
    const int mx = out_h > out_w ? out_h : out_w;
    const double sc = fmax(fmax((double)H / out_h, (double)W / out_w), 1.0);
    const size_t coef_ints = 2 * (size_t)mx * ((size_t)ceil(2.0 * sc) * 2 + 1 + 2) + 64;
    PP_TRY(pp_alloc(e, &tmp, mid_px * 3, "resize pass 1"));
    PP_TRY(pp_alloc(e, &m_out, (size_t)mask_frames * out_h * out_w, "mask resized"));
    PP_TRY(pp_alloc(e, &coef, coef_ints, "resize coefficients"));
    // frames: bicubic on the 8-bit image (image_utils.py:98-103); masks: the 8-bit 'L' image the same way (:142-150)
    PP_TRY(pp_k_resize_bicubic_u8(u8_in, orig_u8, tmp, coef, coef_ints, T, H, W, 3, out_h, out_w, st));
    PP_TRY(pp_k_resize_bicubic_u8(m_in, m_out, tmp, coef, coef_ints, mask_frames, H, W, 1, out_h, out_w, st));
    m_use = m_out;
    e.launches += 4;
  } else {
    PP_CUDA_CHECK(cudaMemcpyAsync(orig_u8, u8_in, (size_t)T * H * W * 3, cudaMemcpyDeviceToDevice, st));
  }
  PP_TRY(pp_k_u8_to_frames(orig_u8, frames, T, out_h, out_w, st));                       // u8/255*2-1 (:186-190)
  // any non-zero -> dilations (image_utils.py:152-170)
  PP_TRY(pp_k_dilate_masks_u8(m_use, mask_frames, T, out_h, out_w, flow_mask_dilates, mask_dilates, flow_masks,
                              masks_dilated, st));
  e.launches += 3;
  return PP_OK;
}

int pp_preprocess_resize(pp_handle h, const float* image, const float* mask, int mask_frames, int T, int H, int W,
                         int out_h, int out_w, int flow_mask_dilates, int mask_dilates, uint8_t* orig_u8, float* frames,
                         float* flow_masks, float* masks_dilated, void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(image && mask && orig_u8 && frames && flow_masks && masks_dilated, "pp_preprocess_resize: null pointer");
  PP_REQUIRE(out_h > 0 && out_w > 0 && H > 0 && W > 0, "pp_preprocess_resize: bad size");
  ArenaGuard guard(e.arena);
  cudaStream_t st = as_stream(stream);
  const size_t in_px = (size_t)T * H * W;
  uint8_t *u8_in, *m_in;
  PP_TRY(pp_alloc(e, &u8_in, in_px * 3, "resize input"));
  PP_TRY(pp_alloc(e, &m_in, (size_t)mask_frames * H * W, "mask input"));
  // float -> uint8 (truncate) for the frames (image_utils.py:106-114) and the mask images (:128-134)
  PP_TRY(pp_k_quantize_u8(image, u8_in, (long long)in_px * 3, st));
  PP_TRY(pp_k_quantize_u8(mask, m_in, (long long)mask_frames * H * W, st));
  e.launches += 2;
  return preprocess_from_u8(e, u8_in, m_in, mask_frames, T, H, W, out_h, out_w, flow_mask_dilates, mask_dilates, orig_u8,
                            frames, flow_masks, masks_dilated, st);
}

int pp_preprocess_u8(pp_handle h, const uint8_t* image_u8, const uint8_t* mask_u8, int mask_frames, int T, int H, int W,
                     int out_h, int out_w, int flow_mask_dilates, int mask_dilates, uint8_t* orig_u8, float* frames,
                     float* flow_masks, float* masks_dilated, void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(image_u8 && mask_u8 && orig_u8 && frames && flow_masks && masks_dilated, "pp_preprocess_u8: null pointer");
  PP_REQUIRE(out_h > 0 && out_w > 0 && H > 0 && W > 0, "pp_preprocess_u8: bad size");
  ArenaGuard guard(e.arena);
  return preprocess_from_u8(e, image_u8, mask_u8, mask_frames, T, H, W, out_h, out_w, flow_mask_dilates, mask_dilates,
                            orig_u8, frames, flow_masks, masks_dilated, as_stream(stream));
}

// Host helper (no GPU involved): float [0,1] -> uint8 with the reference's arithmetic -- x * 255 in float32, clip to
// [0, 255], truncate (utils/image_utils.py:106-114, 128-134) -- on `threads` host threads.  Lets a host caller ship 1/4 of
// the bytes over PCIe: quantise straight into a page-locked staging buffer, copy, then pp_preprocess_u8.
int pp_host_quantize_u8(const float* src, uint8_t* dst, long long n, int threads) {
  PP_REQUIRE(src != nullptr && dst != nullptr && n >= 0, "pp_host_quantize_u8: bad argument");
  if (threads < 1) threads = 1;
  if (threads > 64) threads = 64;
  auto work = [=](long long lo, long long hi) {
    for (long long i = lo; i < hi; ++i) {
      float v = src[i] * 255.0f;
      v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);     // NaN compares false twice and converts to 0 like the device path
      dst[i] = (uint8_t)(int)v;
    }
  };
  if (threads == 1 || n < (1 << 16)) { work(0, n); return PP_OK; }
  std::vector<std::thread> pool;
  const long long per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const long long lo = t * per, hi = lo + per < n ? lo + per : n;
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return PP_OK;
}

int pp_postprocess(pp_handle h, const uint8_t* comp_u8, float* image_out, long long n, void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(comp_u8 && image_out, "pp_postprocess: null pointer");
  e.launches++;
  return pp_k_u8_to_unit_float(comp_u8, image_out, n, as_stream(stream));
}

long long pp_launch_count(pp_handle h) { return h ? reinterpret_cast<PPEngine*>(h)->launches : 0; }
size_t pp_workspace_peak(pp_handle h) { return h ? reinterpret_cast<PPEngine*>(h)->arena.peak : 0; }

int pp_profile_enable(pp_handle h, int on) {
  PP_HANDLE(h);
  PP_CUDA_CHECK(cudaDeviceSynchronize());
  for (auto& r : e.prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  e.prof.clear();
  e.profile = on != 0;
  return PP_OK;
}

// Writes "name\tcount\tms\trows\tflops\tbytes\n" per kernel name (aggregated) into buf.
int pp_profile_dump(pp_handle h, char* buf, size_t cap) {
  PP_HANDLE(h);
  PP_REQUIRE(buf != nullptr && cap > 0, "pp_profile_dump: no buffer");
  PP_CUDA_CHECK(cudaDeviceSynchronize());
  struct Agg { long long n = 0; double ms = 0, rows = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : e.prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
    Agg& a = agg[r.name];
    a.n++; a.ms += ms; a.rows += r.rows; a.flops += r.flops; a.bytes += r.bytes;
  }
  size_t off = 0;
  buf[0] = 0;
  for (auto& kv : agg) {
    char line[512];
    int n = snprintf(line, sizeof(line), "%s\t%lld\t%.6f\t%.0f\t%.0f\t%.0f\n", kv.first.c_str(), kv.second.n,
                     kv.second.ms, kv.second.rows, kv.second.flops, kv.second.bytes);
    if (off + n + 1 > cap) { pp_set_error("pp_profile_dump: buffer too small"); return PP_ERR_ARG; }
    memcpy(buf + off, line, n + 1);
    off += n;
  }
  return PP_OK;
}

// ---- single-operator entry points --------------------------------------------------------------------
int pp_op_conv(pp_handle h, const char* name, const void* x_f16, int N, int H, int W, int stride, int pad, int dil,
               int replicate, int act, float slope, const void* residual_f16, void* out_f16, void* stream) {
  PP_HANDLE(h);
  const PPPackedConv* w = nullptr;
  PP_TRY(pp_get_conv(e, name, &w));
  PPConvCall c(e, name, N, H, W);
  c.in(static_cast<const __half*>(x_f16), w->cin_g * w->groups, 0, w->cin_g, w->groups > 1 ? w->cin_g : 0)
      .geom(stride, stride, pad, pad, dil, dil, replicate)
      .out(out_f16, w->cout_g * w->groups, 0, 0, w->groups > 1 ? w->cout_g : 0)
      .act(act, slope);
  if (residual_f16 != nullptr) c.residual(static_cast<const __half*>(residual_f16), w->cout_g * w->groups, 0);
  return c.run(as_stream(stream));
}

int pp_op_corr_lookup(pp_handle h, const void* l0, const void* l1, const void* l2, const void* l3,
                      const float* coords, void* out_f16, long long nq, int h8, int w8, void* stream) {
  PP_HANDLE(h);
  e.launches++;
  return pp_k_corr_lookup(static_cast<const __half*>(l0), static_cast<const __half*>(l1),
                          static_cast<const __half*>(l2), static_cast<const __half*>(l3), coords,
                          static_cast<__half*>(out_f16), 328, nq, h8 * w8, h8, w8, as_stream(stream));
}

int pp_op_imgprop_step(pp_handle h, const void* cur4_f16, const void* prop_in4_f16, void* prop_out4_f16,
                       const void* flow_prop_f16, const void* flow_check_f16, int H, int W, void* stream) {
  PP_HANDLE(h);
  e.launches++;
  return pp_k_imgprop_step(static_cast<const __half*>(cur4_f16), static_cast<const __half*>(prop_in4_f16),
                           static_cast<__half*>(prop_out4_f16), static_cast<const __half*>(flow_prop_f16),
                           static_cast<const __half*>(flow_check_f16), H, W, as_stream(stream));
}

int pp_op_dcn_sample(pp_handle h, const void* x_f16, const void* offs_f16, const void* flow_f16, float max_mag,
                     void* cols_f16, int N, int H, int W, int C, int tiled, void* stream) {
  PP_HANDLE(h);
  PP_REQUIRE(x_f16 && offs_f16 && cols_f16, "pp_op_dcn_sample: null pointer");
  PPDcnArgs a;
  const __half* x = static_cast<const __half*>(x_f16);
  // C == 256: cat(x0[128], x1[128]) like the flow-completion alignment; C == 128: one tensor (+ flow) like the generator's
  a.x0 = x; a.x0_cs = C; a.x0_co = 0; a.C0 = C == 256 ? 128 : C;
  a.x1 = C == 256 ? x : nullptr; a.x1_cs = C; a.x1_co = 128;
  a.offs = static_cast<const __half*>(offs_f16); a.offs_cs = 432;
  a.flow = static_cast<const __half*>(flow_f16); a.flow_cs = 2; a.flow_co = 0;
  a.max_mag = max_mag; a.cols = static_cast<__half*>(cols_f16); a.C = C; a.N = N; a.H = H; a.W = W;
  e.launches++;
  if (tiled) {
    int handled = 0;
    PP_TRY(pp_k_dcn_sample_tiled(a, 3, as_stream(stream), &handled));
    PP_REQUIRE(handled, "pp_op_dcn_sample: the tiled sampler does not handle this shape");
    return PP_OK;
  }
  return pp_k_dcn_sample_plain(a, as_stream(stream));
}

int pp_op_attention(pp_handle h, const void* qkv_f16, const void* pkv_f16, void* out_f16, const int* win_flags_dev,
                    int t, int gh, int gw, int n_pool, int parity, void* stream) {
  PP_HANDLE(h);
  const int nh = pp_ceil_div(gh, 5) * 5, nw = pp_ceil_div(gw, 9) * 9;
  std::vector<int> ring;
  pp_build_ring_indices(nh, nw, ring);
  ArenaGuard guard(e.arena);
  int* ring_dev;
  PP_TRY(pp_alloc(e, &ring_dev, ring.size(), "ring indices"));
  PP_CUDA_CHECK(cudaMemcpyAsync(ring_dev, ring.data(), ring.size() * sizeof(int), cudaMemcpyHostToDevice,
                                as_stream(stream)));
  const __half* qkv = static_cast<const __half*>(qkv_f16);
  const __half* pkv = static_cast<const __half*>(pkv_f16);
  int meta_host[2] = {0, t};
  int* meta_dev;
  PP_TRY(pp_alloc(e, &meta_dev, 2, "attention meta"));
  PP_CUDA_CHECK(cudaMemcpyAsync(meta_dev, meta_host, sizeof(meta_host), cudaMemcpyHostToDevice, as_stream(stream)));
  const int key_stride = ((t + 1) / 2) * (193 + n_pool);
  int* key_tab;
  PP_TRY(pp_alloc(e, &key_tab, (size_t)(nh / 5) * (nw / 9) * key_stride, "attention key table"));
  int r = pp_k_attention(qkv, qkv + 512, qkv + 1024, 1536, pkv, pkv + 512, 1024, static_cast<__half*>(out_f16), 512,
                         win_flags_dev, ring_dev, meta_dev, meta_dev + 1, 1, t, gh, gw, nh, nw, n_pool, parity,
                         key_tab, key_stride, as_stream(stream));
  PP_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
  e.launches++;
  return r;
}

}  // extern "C"
