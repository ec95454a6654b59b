/*
 * rmnet_b200.h -- C ABI of librmnet_b200.so: the B200 (sm_100a) implementation of hzxie/RMNet's
 * per-frame regional memory-read hot path.
 *
 * This is the drop-in boundary.  The reference has no C ABI of its own for this path: its
 * boundary is pybind11 (extensions/reg_att_map_generator/reg_att_map_generator_cuda.cpp:36-38),
 * the CPython C-API (extensions/flow_affine_transformation/flow_affine_transformation.cpp:87-99)
 * and plain Python methods (models/rmnet.py).  Each entry point below cites the reference
 * interface it replaces; INTEGRATION.md shows the reference-side binding (ctypes stubs that keep
 * the reference's Python signatures).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / pybind types.
 *   - every `const T* x` / `T* x` is a DEVICE pointer unless the parameter name ends in `_host`.
 *   - the caller owns all memory (inputs, outputs, workspaces, banks); the library never allocates
 *     device memory and keeps no pointer past the call.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no implicit host sync
 *     (exception: the *_host convenience entry points, which synchronise `stream` before returning).
 *   - return 0 on success, a negative RMNET_E_* code otherwise; rmnet_last_error() returns a
 *     thread-local message.  Functions are re-entrant; there is no global mutable state.
 *   - float tensors are float32, C-contiguous unless a stride parameter says otherwise.
 */
#ifndef RMNET_B200_H_
#define RMNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RMNET_API __attribute__((visibility("default")))
#else
#define RMNET_API
#endif

#define RMNET_ABI_VERSION 1

#define RMNET_OK 0
#define RMNET_E_INVALID (-1)   /* bad argument (null pointer, non-positive size, misaligned, unsupported shape) */
#define RMNET_E_CUDA (-2)      /* a CUDA runtime / driver call failed; see rmnet_last_error() */
#define RMNET_E_WORKSPACE (-3) /* workspace or bank too small */
#define RMNET_E_UNSUPPORTED (-4)

/* Channel counts fixed by the reference architecture (models/rmnet.py:185-186: keydim=128, valdim=512). */
#define RMNET_CK 128
#define RMNET_CV 512

/* precision modes of the memory read (see DESIGN.md "precision") */
#define RMNET_PREC_SPLIT3 0 /* strict: 16-bit hi/lo split operands, 3 tensor-core products per GEMM, fp32 accumulate */
#define RMNET_PREC_SINGLE 1 /* fast:   hi planes only, 1 product per GEMM                                            */
#define RMNET_PREC_MIXED 2  /* mixed:  3 products for the scores Q.K^T (their error is exponentiated), 1 for P.V        */
/* kernel selection (both are sm_100a CUDA; there is no CPU fallback) */
#define RMNET_IMPL_AUTO 0
#define RMNET_IMPL_SIMT 1   /* CUDA-core fp32 FFMA flash kernel (cross-check / odd shapes) */
#define RMNET_IMPL_UMMA 2   /* tcgen05 + TMEM + TMA kernel                                  */

RMNET_API int rmnet_abi_version(void);
RMNET_API const char *rmnet_last_error(void);
/* Number of kernels this thread has launched through the library since the last reset (bench.py's gpu_launches). */
RMNET_API long long rmnet_launch_count(void);
RMNET_API void rmnet_launch_count_reset(void);
/* 1 when the tcgen05 / TMEM / TMA memory-read kernel is compiled into this build, else 0. */
RMNET_API int rmnet_has_umma(void);

/* ---------------------------------------------------------------------------------------------
 * Regional attention-map generator.
 * Replaces reg_att_map_generator.forward(mask, prob_threshold, n_pts_threshold, n_bbox_loose_pixels)
 *   (reg_att_map_generator_cuda.cpp:26-38 -> reg_att_map_generator.cu:95-123, kernel :15-93).
 *   mask     [B,K,H,W] f32
 *   bboxes   [B,K,4]   i32 = (x_min, x_max, y_min, y_max) inclusive; channel 0 -> (0,0,0,0)
 *   att_full [B,K,H,W] f32 in {0,1}, nullable (every element is written: no pre-zeroing needed)
 *   workspace: rmnet_reg_att_map_workspace_bytes(B,K) bytes, 16-byte aligned, ZERO-FILLED ONCE by the
 *              caller at allocation; the kernels leave it zeroed again (self-cleaning), so it can be
 *              reused by consecutive calls on the same stream.
 * ------------------------------------------------------------------------------------------- */
RMNET_API size_t rmnet_reg_att_map_workspace_bytes(int B, int K);
RMNET_API int rmnet_reg_att_map_forward(const float *mask, int B, int K, int H, int W, float prob_threshold,
                              int n_pts_threshold, int n_bbox_loose_pixels, int *bboxes,
                              float *att_full, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * RMNet.warp(img0, flow) -> (img1, mask)                       (models/rmnet.py:252-278)
 *   img0 [B,C,H,W], flow [B,2,H,W] (ch 0 = x, ch 1 = y, pixels); img1 [B,C,H,W]; valid [B,C,H,W]
 *   (the reference's `mask` output: the same [H,W] validity plane broadcast over C), nullable.
 *   Bit-exact with the reference evaluated on torch's CUDA backend (reciprocal-multiply
 *   normalisation, FMA-chained bilinear taps, >= 0.9999 validity).  `sampler` selects the tap
 *   accumulation order: RMNET_SAMPLER_CUDNN (ne,nw,sw,se: cudnnSpatialTfSamplerForward, what
 *   F.grid_sample dispatches to by default on CUDA) or RMNET_SAMPLER_ATEN (nw,ne,sw,se: ATen's
 *   own kernel, used when torch.backends.cudnn.enabled is False).
 * ------------------------------------------------------------------------------------------- */
#define RMNET_SAMPLER_CUDNN 0
#define RMNET_SAMPLER_ATEN 1
RMNET_API int rmnet_warp_forward(const float *img0, const float *flow, int B, int C, int H, int W, int sampler,
                                 float *img1, float *valid, void *stream);

/* ---------------------------------------------------------------------------------------------
 * RMNet.get_att_map(prev_mask, flow) fused: warp + threshold + bbox in ONE pass, the warped mask
 * is never written to HBM.                                      (models/rmnet.py:280-287)
 *   outputs / workspace as rmnet_reg_att_map_forward.
 * ------------------------------------------------------------------------------------------- */
RMNET_API int rmnet_warp_att_map_forward(const float *prev_mask, const float *flow, int B, int K, int H, int W,
                               int sampler, float prob_threshold, int n_pts_threshold,
                               int n_bbox_loose_pixels, int *bboxes, float *att_full, void *workspace,
                               size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Region descriptors of one frame in ONE launch: bounding boxes + /16 cell rectangles, straight from the
 * UNPADDED soft masks -- replaces pad_divide_by (utils/helpers.py:105-124) + get_att_map + F.interpolate(1/16):
 *   flow == NULL, bbox_in_padded_frame = 1 : the memorise side, models/rmnet.py:212 + :244-245 (box of the zero-padded
 *       mask, in padded coordinates);
 *   flow != NULL, bbox_in_padded_frame = 0 : the segment side, :431 (warp + box in raw coordinates) + :307 + :356.
 *   pad_* = pad_divide_by amounts (lw, uw, lh, uh); bboxes [B,K,4]; cell_rects [B,K,4] (cx0,cx1,cy0,cy1), channel 0
 *   and empty boxes -> (0,-1,0,-1); workspace as rmnet_reg_att_map_forward.
 *   k_scan: 0 or K = scan every channel like the reference.  2 <= k_scan < K = read only channels [1, k_scan) and report
 *   channels >= k_scan as ABSENT objects (what the reference computes for an all-below-threshold channel).  Valid when
 *   those channels are known to stay below prob_threshold -- the reference's frame loop guarantees it for
 *   j > n_max_objects (their logit is forced to -16.1181, models/rmnet.py:444-448).
 * ------------------------------------------------------------------------------------------- */
RMNET_API int rmnet_regional_boxes_forward(const float *mask, const float *flow, int B, int K, int H, int W,
                                           int sampler, float prob_threshold, int n_pts_threshold,
                                           int n_bbox_loose_pixels, int pad_l, int pad_r, int pad_t, int pad_b,
                                           int bbox_in_padded_frame, int k_scan, int *bboxes, int *cell_rects,
                                           void *workspace, size_t workspace_bytes, void *stream);

/* Both sides of one frame in ONE pass over prev_mask (the memorise side and the segment side read the same
 * est_masks[t-1], models/rmnet.py:412-414 and :431): mem_* = box of the zero-padded mask in padded coordinates and
 * its cell rectangles; cur_* = box of the flow-warped mask in raw coordinates and its cell rectangles. */
RMNET_API int rmnet_frame_regions_forward(const float *prev_mask, const float *flow, int B, int K, int H, int W,
                                          int sampler, float prob_threshold, int n_pts_threshold,
                                          int n_bbox_loose_pixels, int pad_l, int pad_r, int pad_t, int pad_b,
                                          int k_scan, int *mem_bboxes, int *mem_rects, int *cur_bboxes, int *cur_rects,
                                          void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Low-resolution cell rectangles: the closed form of
 *   F.interpolate(pad(att_map), scale_factor=1/16)             (models/rmnet.py:245, :307+:356)
 * for a rectangular att_map:  cx in [ceil((x0+pad_l)/16), floor((x1+pad_l)/16)], same for y.
 *   bboxes [count,4] i32 (x_min,x_max,y_min,y_max) -> rects [count,4] i32 (cx0,cx1,cy0,cy1),
 *   clamped to the h x w grid; empty when lo > hi.  `skip_channel0_every` = K marks every K-th
 *   entry (the background channel, whose att_map is all zero) as empty; pass 0 to disable.
 * ------------------------------------------------------------------------------------------- */
RMNET_API int rmnet_cell_rects_from_bboxes(const int *bboxes, int count, int pad_l, int pad_t, int h, int w,
                                 int skip_channel0_every, int *rects, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Memory bank: the preallocated, region-compacted replacement of the reference's growing
 * `keys [B,K,128,T,h,w]` / `values [B,K,512,T,h,w]` tensors (models/rmnet.py:239-248 pad_memory +
 * regional multiply, :416-426 torch.cat growth, :348-349 per-object gather).
 *
 * One opaque device blob per clip; slot s holds object s+1.  Only in-region cells are stored
 * (16-bit hi/lo planes, keys position-major, values channel-major); masked cells are only counted.
 * A frame is first written as the TEMPORARY last frame (the reference's `this_keys`); `commit != 0`
 * makes it permanent (`keys = this_keys`, :424-426).  The next memorize overwrites a temporary frame.
 * ------------------------------------------------------------------------------------------- */
RMNET_API size_t rmnet_bank_bytes(int n_slots, int cap_cells);
RMNET_API int rmnet_bank_reset(void *bank, size_t bank_bytes, int n_slots, int cap_cells, void *stream);
/*   k4 [n_obj,128,h*w] (object stride k_obj_stride, channel stride k_ch_stride, in floats),
 *   v4 [n_obj,512,h*w] likewise; rects [n_obj,4] cell rectangles of this frame (device);
 *   elem_format: 0 = bf16 planes, 1 = fp16 planes. */
RMNET_API int rmnet_bank_memorize(void *bank, size_t bank_bytes, int n_slots, int cap_cells, const float *k4,
                        long long k_obj_stride, long long k_ch_stride, const float *v4,
                        long long v_obj_stride, long long v_ch_stride, const int *rects, int n_obj,
                        int h, int w, int elem_format, int commit, void *stream);
/* Host-visible copy of the per-slot counters (synchronises `stream`): out_host [n_slots,8] i32 =
 * (cells_committed, cells_temp, zeros_committed, zeros_temp, frames_committed, frames_temp, overflow, range).
 * Capacity: a memorize whose in-region cells do not fit behind the committed ones (cells_committed + r > cap_cells)
 * stores NOTHING for that slot (the frame then counts as fully masked) and sets the slot's sticky `overflow` flag, which
 * only rmnet_bank_reset clears.  The kernels cannot report that through a return code (they run asynchronously), so a
 * direct C caller must size cap_cells for the frames it commits or poll this function: it fills out_host and returns
 * RMNET_E_WORKSPACE when any slot has overflowed.
 * Range: elem_format 1 (fp16 hi/lo planes, 22 mantissa bits -- the default of the Python layer) represents |x| <= 65504;
 * a key / value / query key beyond that is stored saturated and sets the sticky `range` flag (-> RMNET_E_UNSUPPORTED
 * here).  elem_format 0 (bf16 planes, 16 mantissa bits) has fp32's range. */
RMNET_API int rmnet_bank_stats_host(const void *bank, int n_slots, int cap_cells, int *out_host, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Regional memory read against a bank: replaces the per-frame
 *   F.interpolate(att,1/16); k4e*=att; v4e*=att; MemoryReader.forward(key, value, k4e, v4e)
 *   (models/rmnet.py:355-361, :147-165) for all objects in one call.
 *   q_key [n_q?,128,h*w], q_val [.,512,h*w]: q_obj_stride = 0 when one query frame is shared by all
 *     objects (the reference's expand, :332-333), else the per-object stride in floats.
 *   q_rects [n_obj,4] cell rectangles of the query frame (device); NULL = dense (all cells).
 *   mem_val [n_obj,1024,h,w] f32: channels 0..511 the memory read, 512..1023 the (masked) q_val.
 *   workspace: rmnet_memory_read_workspace_bytes(...) bytes (partial results of the split-KV pass).
 *   stages: RMNET_STAGE_ALL normally; RMNET_STAGE_QUERY / _PARTIAL / _MERGE run only the query-side preparation
 *     (k4e*att16 packed for the tensor cores, v4e*att16 into mem_val[:,512:]) / the split-KV attention kernel / the
 *     merge+scatter kernel, each of which needs its predecessors' results from an earlier call with the same arguments
 *     (so a benchmark can time each launch alone).  The query stage also builds, on the device, the work plan that the
 *     tcgen05 attention kernel walks (which KV chunk of which object runs on which SM; rmnet_memory_read_plan_host
 *     copies it out): a PARTIAL-only call relies on the plan of the last QUERY stage run on this workspace, so the bank
 *     and q_rects must not have changed in between.
 * ------------------------------------------------------------------------------------------- */
#define RMNET_STAGE_PARTIAL 1
#define RMNET_STAGE_MERGE 2
#define RMNET_STAGE_QUERY 4 /* query-side preparation: packed query keys + the q_val passthrough half of mem_val */
#define RMNET_STAGE_ALL 7
RMNET_API size_t rmnet_memory_read_workspace_bytes(int n_obj, int h, int w, int cap_cells);
RMNET_API int rmnet_bank_memory_read(const void *bank, size_t bank_bytes, int n_slots, int cap_cells,
                           const float *q_key, const float *q_val, long long q_obj_stride,
                           const int *q_rects, int n_obj, int h, int w, int elem_format,
                           int precision, int impl, int stages, float *mem_val, void *workspace,
                           size_t workspace_bytes, void *stream);

/* Host copy of the work plan that the last query-side / pack launch on this workspace built for the tcgen05 read kernel
 * (introspection and tests; synchronises `stream`).  The plan cuts every object's stored cells into KV chunks of 64-cell
 * tiles and assigns (object, query tile, Cv half, chunk) pieces to the persistent CTAs:
 *   ns_out [n_obj]        partial slots (= chunks) per object, <= 16
 *   hdr_out [256][2]      per CTA: number of pieces, index of its first piece      (n_ctas_out CTAs are in use)
 *   pieces_out [max_pieces][4]   object | query_tile << 8 | half << 16 | slot << 20,  first tile,  tiles,  stored cells */
RMNET_API int rmnet_memory_read_plan_host(const void *workspace, int n_obj, int h, int w, int *ns_out, int *hdr_out,
                                          int *pieces_out, int max_pieces, int *n_ctas_out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * One frame of the reference's loop body in ONE call (models/rmnet.py:414-432 minus the convolutions), batch 1:
 *   rmnet_frame_regions_forward(prev_mask, flow)  ->  rmnet_bank_memorize(k4, v4, mem_rects[1..n])
 *   ->  rmnet_bank_memory_read(q_key, q_val shared by all objects, cur_rects[1..n])  ->  mem_val [n_obj,1024,h,w]
 *   boxes_out [4][K][4] i32 = mem_bboxes, mem_rects, cur_bboxes, cur_rects.  4 kernels (+1 on commit), chained with
 *   programmatic dependent launch: regions -> pack (memory + query side) [-> commit] -> tcgen05 read -> merge.
 * ------------------------------------------------------------------------------------------- */
RMNET_API int rmnet_frame_step(void *bank, size_t bank_bytes, int n_slots, int cap_cells, const float *prev_mask,
                               const float *flow, int K, int H, int W, int sampler, float prob_threshold,
                               int n_pts_threshold, int n_bbox_loose_pixels, int pad_l, int pad_r, int pad_t,
                               int pad_b, int k_scan, const float *k4, const float *v4, const float *q_key,
                               const float *q_val, int n_obj, int elem_format, int precision, int impl,
                               int commit, int *boxes_out, float *mem_val, void *box_workspace,
                               size_t box_workspace_bytes, void *read_workspace, size_t read_workspace_bytes,
                               void *stream);

/* ---------------------------------------------------------------------------------------------
 * The per-frame tail of RMNet.segment / RMNet.forward after the decoder, in ONE pass (batch 1):
 *   ps = F.softmax(decoder_logits, dim=1)[:, 1]                        (models/rmnet.py:368-370)
 *   logit = soft_aggregation(ps, K, n_obj)                             (:289-302, :373)
 *   un-pad by the pad_divide_by amounts                                (:376-380)
 *   whole-channel overrides of the frame loop                          (:436-448)
 *   est_mask = F.softmax(logit, dim=1)                                 (:450)
 *   dec_logits [n_obj,2,H+pad_t+pad_b,W+pad_l+pad_r] f32 (the decoder output);
 *   channel_mode_host [K] ints (HOST array, read at call time; NULL = all RMNET_CH_KEEP):
 *     RMNET_CH_KEEP   the soft-aggregation logit,
 *     RMNET_CH_ABSENT logit := -16.1181               (j <= n_max_objects not in existing_objects, :445-448),
 *     RMNET_CH_NEW    logit := new_mask[j] * 32.0605 - 16.1181   (object first annotated in this frame, :438-442);
 *   new_mask [K,H,W] i32 = masks[i,t] (only read for RMNET_CH_NEW channels, may be NULL otherwise);
 *   logit_out [K,H,W] f32 nullable (the return value of segment() after the overrides); est_mask [K,H,W] f32.
 * ------------------------------------------------------------------------------------------- */
#define RMNET_CH_KEEP 0
#define RMNET_CH_ABSENT 1
#define RMNET_CH_NEW 2
RMNET_API int rmnet_mask_epilogue_forward(const float *dec_logits, int n_obj, int K, int H, int W, int pad_l, int pad_r,
                                          int pad_t, int pad_b, const int *channel_mode_host, const int *new_mask,
                                          float *logit_out, float *est_mask, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Literal MemoryReader.forward(m_key, m_val, q_key, q_val) -> mem_val   (models/rmnet.py:147-165)
 *   m_key [n,128,T,h,w], m_val [n,512,T,h,w], q_key [n,128,h,w], q_val [n,512,h,w] (contiguous f32)
 *   mem_val [n,1024,h,w].  Region-agnostic (dense): packs the inputs into a scratch bank inside
 *   `workspace` and runs the same read kernel.  workspace >= rmnet_memory_reader_workspace_bytes().
 *   The reference's second output `p [n,T*h*w,h*w]` (softmax over the memory axis, :155-157) is produced only when
 *   p != NULL, by a separate fp32 FFMA kernel over the raw keys; the per-frame path never materialises it.
 * ------------------------------------------------------------------------------------------- */
RMNET_API size_t rmnet_memory_reader_workspace_bytes(int n, int T, int h, int w);
RMNET_API int rmnet_memory_reader_forward(const float *m_key, const float *m_val, const float *q_key,
                                const float *q_val, int n, int T, int h, int w, int elem_format,
                                int precision, int impl, float *mem_val, float *p, void *workspace,
                                size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * flow_affine_transformation.update_optical_flow(of, M1, M2)   (flow_affine_transformation.cpp:39-85)
 *   of [H,W,2] f32, m1_host / m2_host: 6 floats each (row-major 2x3, read on the host at call time),
 *   out [H,W,2] f32.  Bit-exact with the reference's -O2 (no-FMA) build.
 *   The *_host variant takes HOST arrays (what the NumPy-facing module passes), stages them through
 *   `dev_scratch` (>= 2*H*W*2*4 bytes of device memory) and synchronises before returning.
 *   The *_cpu variant is the same arithmetic in plain host C (no CUDA call at all; compiled with
 *   -ffp-contract=off): the reference calls this op inside forked DataLoader worker processes
 *   (utils/data_transforms.py:293-302), where a CUDA context cannot be created.  All pointers are HOST.
 * ------------------------------------------------------------------------------------------- */
RMNET_API int rmnet_update_optical_flow(const float *of, const float *m1_host, const float *m2_host, int H, int W,
                              float *out, void *stream);
RMNET_API int rmnet_update_optical_flow_cpu(const float *of_host, const float *m1_host, const float *m2_host, int H, int W,
                                  float *out_host);
RMNET_API int rmnet_update_optical_flow_host(const float *of_host, const float *m1_host, const float *m2_host,
                                   int H, int W, float *out_host, void *dev_scratch,
                                   size_t dev_scratch_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RMNET_B200_H_ */
