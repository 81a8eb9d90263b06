/* C ABI of the B200-native ProPainter inference path (libpropainter_b200.so).
 *
 * Every entry point takes raw device pointers, sizes and a CUDA stream (as void*, 0 = legacy default
 * stream) and returns 0 on success; on failure pp_last_error() holds a message.  No torch types cross this
 * boundary.  Tensors are contiguous, float32, in the layouts the reference's own tensors have at the same
 * call sites (paths relative to daniabib/ComfyUI_ProPainter_Nodes):
 *
 *   pp_raft_bidir          replaces  raft_model(frames, iters)                propainter_inference.py:77-93
 *                                    (RAFT_bi.forward, model/modules/flow_comp_raft.py:39-58)
 *   pp_flow_complete       replaces  forward_bidirect_flow + combine_flow     propainter_inference.py:123-150
 *                                    (model/recurrent_flow_completion.py:356-400)
 *   pp_image_propagate     replaces  img_propagation + blend                  propainter_inference.py:186-219
 *                                    (model/propainter.py:350-356, 118-231)
 *   pp_gen_begin/window    replace   inpaint_model(selected_imgs, ...)        propainter_inference.py:272-281
 *                                    (InpaintGenerator.forward, model/propainter.py:358-453)
 *   pp_composite           replaces  the numpy composite                      propainter_inference.py:283-307
 *   pp_register_*          replace   load_state_dict of the three checkpoints utils/model_utils.py:49-59
 */
#ifndef PROPAINTER_B200_H
#define PROPAINTER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PP_API __attribute__((visibility("default")))
#else
#define PP_API
#endif

typedef struct PPEngine* pp_handle;

/* Message of the last failing call on this thread. */
PP_API const char* pp_last_error(void);
/* Library build id ("propainter_b200 <n> sm_100a"). */
PP_API const char* pp_version(void);

/* Create an engine on `device` whose scratch arena is the caller-allocated device buffer
 * [workspace, workspace + workspace_bytes) (256-byte aligned). */
PP_API int pp_create(int device, void* workspace, size_t workspace_bytes, pp_handle* out);
PP_API int pp_destroy(pp_handle h);
/* Replace the scratch arena (grow it for a larger clip, or shrink it after one); the previous buffer may be freed by
 * the caller once this returns.  Fails while a generator session is open. */
PP_API int pp_set_workspace(pp_handle h, void* workspace, size_t workspace_bytes);

/* Register packed weights (device memory owned by the caller, must outlive the handle).
 * `w` is the 128B-swizzled tile image produced by comfyui_propainter_nodes_b200.weights_pack,
 * [groups][ceil(kh*kw*cin_g/64)][cout_g_pad] rows of 64 fp16; `bias` is float32 [groups*cout_g] or NULL. */
PP_API int pp_register_conv(pp_handle h, const char* name, const void* w, const float* bias, int cout_g, int cout_g_pad,
                     int bn, int cin_g, int kh, int kw, int groups);
PP_API int pp_register_tensor(pp_handle h, const char* name, const void* ptr, size_t bytes);
/* Multiply-adds per output pixel of the reference layer behind a registered conv (unpadded channels, real groups);
 * only feeds the algorithmic flop count of pp_profile_dump. */
PP_API int pp_set_conv_macs(pp_handle h, const char* name, double macs_per_pixel);

/* ---- multi-GPU: one process per GPU, an NCCL communicator per engine ------------------------
 * Rank 0 calls pp_comm_unique_id and ships the 128 bytes to the other ranks (any host transport); every rank then
 * calls pp_comm_init(h, id, rank, world).  NCCL is resolved with dlopen at run time (the process' own copy first). */
PP_API int pp_comm_unique_id(void* out128);
PP_API int pp_comm_init(pp_handle h, const void* unique_id128, int rank, int world);
PP_API int pp_comm_destroy(pp_handle h);
/* In-place all-gather of uneven row blocks among ranks [first_rank, first_rank + n_members): `buf` holds
 * sum(rows_per_member) rows of row_bytes, member m has written its own block; on return (stream order) every member
 * holds all blocks.  One NCCL send/recv group over NVLink; ranks outside the range return at once. */
PP_API int pp_comm_all_gather_rows(pp_handle h, void* buf, const long long* rows_per_member, size_t row_bytes,
                                   int first_rank, int n_members, void* stream);

/* frames [T,3,H,W] in [-1,1]  ->  flows_f, flows_b [T-1,2,H,W]. */
PP_API int pp_raft_bidir(pp_handle h, const float* frames, int T, int H, int W, int iters, float* flows_f, float* flows_b,
                  void* stream);
/* flows [T-1,2,H,W], flow_masks [T,1,H,W]  ->  completed flows [T-1,2,H,W] (prediction inside the mask). */
PP_API int pp_flow_complete(pp_handle h, const float* flows_f, const float* flows_b, const float* flow_masks, int T, int H,
                     int W, float* out_f, float* out_b, void* stream);
/* The same call made collectively by the ranks [team_first, team_first + team_size) of the communicator on the same
 * (replicated) inputs; ranks outside the team return at once.  The two direction passes go to the two halves of the
 * team, the per-frame encoder / decoder is sharded inside a half (encoder with the +-8-frame temporal halo), the
 * serial recurrence runs on every rank of its half; NCCL all-gathers of the encoder features (inside a half) and of
 * the completed flows (whole team) leave the full result on every rank of the team.  team_size 1 = pp_flow_complete. */
PP_API int pp_flow_complete_dist(pp_handle h, const float* flows_f, const float* flows_b, const float* flow_masks, int T,
                                 int H, int W, float* out_f, float* out_b, int team_first, int team_size, void* stream);
/* frames [T,3,H,W], masks [T,1,H,W], completed flows  ->  updated frames [T,3,H,W], updated masks [T,1,H,W]. */
PP_API int pp_image_propagate(pp_handle h, const float* frames, const float* masks, const float* flows_f,
                       const float* flows_b, int T, int H, int W, float* updated_frames, float* updated_masks,
                       void* stream);
/* Generator session over one clip: encodes all T frames once. */
PP_API int pp_gen_begin(pp_handle h, const float* updated_frames, const float* masks_dilated, const float* updated_masks,
                 const float* flows_f, const float* flows_b, int T, int H, int W, void* stream);
/* Same, encoding only the frames with frames_needed[f] != 0 (host array of T bytes): a rank that runs a shard of the
 * sliding windows encodes the frames those windows touch; windows passed to pp_gen_run must stay inside that set. */
PP_API int pp_gen_begin_subset(pp_handle h, const float* updated_frames, const float* masks_dilated,
                               const float* updated_masks, const float* flows_f, const float* flows_b, int T, int H, int W,
                               const unsigned char* frames_needed, void* stream);
/* One sliding window: frame_ids[0..l_t) are consecutive local frames, frame_ids[l_t..t) reference frames
 * (host array).  pred is fp16 [l_t][H][W][4] (rgb in [-1,1] + 1 unused lane). */
PP_API int pp_gen_window(pp_handle h, const int* frame_ids, int t, int l_t, void* pred_f16, void* stream);
/* All sliding windows of the clip in one batched pass: frame_ids is the concatenation of every window's
 * [local..., reference...] ids, win_t / win_lt the per-window frame counts (host arrays).
 * pred is fp16 [sum(win_lt)][H][W][4] in window order. */
PP_API int pp_gen_run(pp_handle h, const int* frame_ids, const int* win_t, const int* win_lt, int n_windows,
                      void* pred_f16, void* stream);
PP_API int pp_gen_end(pp_handle h);
/* uint8 composite with the reference's truncation / 0.5-0.5 blending order.  frame_ids / first_visit are
 * device int32 arrays of length l_t; orig / comp are uint8 [T][H][W][3]; masks float32 [T,1,H,W].
 * half_math = 1 reproduces the roundings of the reference's fp16="enable" mode ((pred+1)/2 and *255 evaluated in
 * half precision before the uint8 truncation), 0 its fp32 mode. */
PP_API int pp_composite(pp_handle h, const void* pred_f16, const float* masks_dilated, const uint8_t* orig, uint8_t* comp,
                 const int* frame_ids_dev, const int* first_visit_dev, int l_t, int H, int W, int half_math,
                 void* stream);

/* Device pre-processing when no resize is needed (reference utils/image_utils.py:106-197): image [T,H,W,3]
 * float 0..1 -> uint8 (truncate) [T,H,W,3] + frames [T,3,H,W] in [-1,1]; mask [mask_frames,H,W] float ->
 * 8-bit -> cross dilation x N -> flow_masks / masks_dilated [T,1,H,W] in {0,1}.  All device pointers. */
PP_API int pp_preprocess(pp_handle h, const float* image, const float* mask, int mask_frames, int T, int H, int W,
                         int flow_mask_dilates, int mask_dilates, uint8_t* orig_u8, float* frames, float* flow_masks,
                         float* masks_dilated, void* stream);
/* The same with a resize to the processing size out_w x out_h (reference utils/image_utils.py:98-103: PIL
 * Image.resize, bicubic, on the 8-bit frames and on the 8-bit mask images) done on the device, bit-identical to
 * Pillow's 8-bit resampler (two passes, 22-bit fixed-point coefficients).  Outputs have the processing size. */
PP_API int pp_preprocess_resize(pp_handle h, const float* image, const float* mask, int mask_frames, int T, int H, int W,
                                int out_h, int out_w, int flow_mask_dilates, int mask_dilates, uint8_t* orig_u8,
                                float* frames, float* flow_masks, float* masks_dilated, void* stream);
/* The same from 8-bit inputs already on the device (image_u8 [T,H,W,3], mask_u8 [mask_frames,H,W]): what is left of the
 * pre-processing after the float -> uint8 truncation, which a host can do itself (pp_host_quantize_u8) to move 1/4 of the
 * bytes over PCIe.  out_h / out_w equal to H / W: no resize. */
PP_API int pp_preprocess_u8(pp_handle h, const uint8_t* image_u8, const uint8_t* mask_u8, int mask_frames, int T, int H,
                            int W, int out_h, int out_w, int flow_mask_dilates, int mask_dilates, uint8_t* orig_u8,
                            float* frames, float* flow_masks, float* masks_dilated, void* stream);
/* HOST helper, no GPU work: dst[i] = (uint8)trunc(clip(src[i] * 255, 0, 255)) in float32, the reference's conversion
 * (utils/image_utils.py:106-114, 128-134), on `threads` host threads; src / dst are host pointers. */
PP_API int pp_host_quantize_u8(const float* src, uint8_t* dst, long long n, int threads);
/* uint8 frames -> float32 / 255 (reference handle_output, utils/image_utils.py:276-290). */
PP_API int pp_postprocess(pp_handle h, const uint8_t* comp_u8, float* image_out, long long n, void* stream);

/* Kernels launched by this handle since creation (bench accounting), and the arena high-water mark. */
PP_API long long pp_launch_count(pp_handle h);
PP_API size_t pp_workspace_peak(pp_handle h);

/* Per-kernel timing with CUDA events on the launch stream (bench.py roofline): enable, run, dump a
 * tab-separated table "name count ms rows flops bytes" aggregated by kernel name. */
PP_API int pp_profile_enable(pp_handle h, int on);
PP_API int pp_profile_dump(pp_handle h, char* buf, size_t cap);

/* ---- single-operator entry points (unit tests and micro-benchmarks) ------------------------------------ */
/* Generic conv / linear through the tcgen05 implicit-GEMM kernel: x NHWC fp16 [N,H,W,cin_g*groups]. */
PP_API int pp_op_conv(pp_handle h, const char* name, const void* x_f16, int N, int H, int W, int stride, int pad, int dil,
               int replicate, int act, float slope, const void* residual_f16, void* out_f16, void* stream);
PP_API int pp_op_corr_lookup(pp_handle h, const void* l0, const void* l1, const void* l2, const void* l3,
                      const float* coords, void* out_f16, long long nq, int h8, int w8, void* stream);
PP_API int pp_op_imgprop_step(pp_handle h, const void* cur4_f16, const void* prop_in4_f16, void* prop_out4_f16,
                       const void* flow_prop_f16, const void* flow_check_f16, int H, int W, void* stream);
/* Modulated deformable sampling -> columns [N*H*W][9*C]: x [N,H,W,C] (C = 128, or 256 = two 128-channel halves),
 * offs [N,H,W,432] raw offset-head output, flow [N,H,W,2] (dx,dy) or NULL; tiled = 1: TMA-staged sampler. */
PP_API int pp_op_dcn_sample(pp_handle h, const void* x_f16, const void* offs_f16, const void* flow_f16, float max_mag,
                            void* cols_f16, int N, int H, int W, int C, int tiled, void* stream);
PP_API int pp_op_attention(pp_handle h, const void* qkv_f16, const void* pkv_f16, void* out_f16, const int* win_flags_dev,
                    int t, int gh, int gw, int n_pool, int parity, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROPAINTER_B200_H */
