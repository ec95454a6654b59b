// memory_read.cu -- host-side dispatch of the regional memory read (C ABI entry points).
#include "common.cuh"

namespace rmnet {
int launch_memory_read_simt(const BankView &bank, const float *q_key, long long q_obj_stride, const int *q_rects,
                            int n_obj, int h, int w, int fmt, int precision, int n_splits, const ReadWorkspace &W,
                            cudaStream_t st);
int launch_memory_read_umma(const BankView &bank, int n_obj, int fmt, int precision, const ReadWorkspace &W, bool pdl,
                            cudaStream_t st);
int umma_grid_size();
int launch_merge(const BankView &bank, const int *q_rects, int n_obj, int h, int w, int n_splits, bool device_sched,
                 const ReadWorkspace &W, float *mem_val, bool pdl, cudaStream_t st);
bool umma_supported(int cap_cells);
int launch_attention_probs(const float *m_key, const float *q_key, int n, int M, int N, float *p, cudaStream_t st);

namespace {
__global__ void fill_dense_rects_kernel(int *rects, int n, int h, int w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) reinterpret_cast<int4 *>(rects)[i] = make_int4(0, w - 1, 0, h - 1);
}


// plan = true: the launch that prepares the query side also builds the work plan of the tcgen05 read (sched.cuh)
QuerySide make_query_side(const float *q_key, const float *q_val, long long q_key_obj_stride, const int *q_rects,
                          const ReadWorkspace &W, int N, float *mem_val, int *range_flag, bool plan, const BankView &bv,
                          int precision) {
  QuerySide qs = {};
  if (plan) {
    qs.plan_bank_meta = bv.meta;
    qs.plan_ns = W.sched;
    qs.plan_hdr = W.plan_hdr;
    qs.plan_pieces = W.plan_pieces;
    qs.plan_piece_cap = (int)W.plan_piece_cap;
    qs.plan_ctas = umma_grid_size();
    qs.plan_precision = precision;
    qs.plan_cap_cells = bv.cap;
  }
  qs.range_flag = range_flag;
  qs.q_key = q_key;
  qs.q_val = q_val;
  qs.q_key_obj_stride = q_key_obj_stride;
  qs.q_val_obj_stride = q_key_obj_stride ? (q_key_obj_stride / RMNET_CK) * RMNET_CV : 0;
  qs.q_rects = q_rects;
  qs.qhi = W.qhi;
  qs.qlo = W.qlo;
  qs.nq_pad = W.nq_pad;
  qs.vec4 = (N % 4 == 0 && (uintptr_t)q_val % 16 == 0 && (uintptr_t)mem_val % 16 == 0 && qs.q_val_obj_stride % 4 == 0) ? 1 : 0;
  qs.mem_val = mem_val;
  return qs;
}

// chained = true: called from rmnet_frame_step, the launches are programmatic dependents of the pack / commit kernel.
int bank_memory_read_impl(const void *bank, size_t bank_bytes, int n_slots, int cap_cells, const float *q_key,
                          const float *q_val, long long q_obj_stride, const int *q_rects, int n_obj, int h, int w,
                          int elem_format, int precision, int impl, int stages, float *mem_val, void *workspace,
                          size_t workspace_bytes, bool chained, void *stream) {
  RMNET_CHECK_ARG(bank && q_key && q_val && mem_val && workspace, "null pointer argument");
  RMNET_CHECK_ARG(n_obj > 0 && n_obj <= n_slots && h > 0 && w > 0, "bad shape");
  RMNET_CHECK_ARG(n_obj <= 65535, "too many objects");
  RMNET_CHECK_ARG(elem_format == 0 || elem_format == 1, "elem_format must be 0 (bf16) or 1 (fp16)");
  RMNET_CHECK_ARG(precision == RMNET_PREC_SPLIT3 || precision == RMNET_PREC_SINGLE || precision == RMNET_PREC_MIXED, "bad precision mode");
  RMNET_CHECK_ARG(q_rects == nullptr || (uintptr_t)q_rects % 16 == 0, "q_rects must be 16-byte aligned");
  BankLayout L = bank_layout(n_slots, cap_cells);
  if (bank_bytes < L.total) { set_error("bank too small"); return RMNET_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  BankView bv = bank_view(const_cast<void *>(bank), n_slots, cap_cells);
  if (impl == RMNET_IMPL_AUTO) impl = (umma_supported(cap_cells) && n_obj <= SCHED_MAX_OBJ) ? RMNET_IMPL_UMMA : RMNET_IMPL_SIMT;
  RMNET_CHECK_ARG(impl == RMNET_IMPL_UMMA || impl == RMNET_IMPL_SIMT, "unknown impl %d", impl);
  const int n_splits = impl == RMNET_IMPL_UMMA ? READ_MAX_SPLITS : pick_splits(n_obj, h * w, 64, cap_cells);
  ReadWorkspace W = read_workspace(workspace, n_obj, h * w, n_splits);
  if (workspace_bytes < W.total) { set_error("workspace too small: %zu < %zu", workspace_bytes, W.total); return RMNET_E_WORKSPACE; }
  RMNET_CHECK_ARG(stages >= 1 && stages <= RMNET_STAGE_ALL, "bad stages mask %d", stages);
  int rc = RMNET_OK;
  const bool umma = impl == RMNET_IMPL_UMMA;
  if (stages & RMNET_STAGE_QUERY) {  // (rmnet_frame_step folds this into its pack launch instead)
    QuerySide qs = make_query_side(q_key, q_val, q_obj_stride, q_rects, W, h * w, mem_val, bv.meta + META_RANGE, umma, bv, precision);
    if ((rc = launch_query_side(qs, n_obj, h, w, elem_format, st))) return rc;
  }
  if (!(stages & RMNET_STAGE_PARTIAL)) {
  } else if (umma)
    rc = launch_memory_read_umma(bv, n_obj, elem_format, precision, W, chained, st);
  else
    rc = launch_memory_read_simt(bv, q_key, q_obj_stride, q_rects, n_obj, h, w, elem_format, precision, n_splits, W, st);
  if (rc || !(stages & RMNET_STAGE_MERGE)) return rc;
  // the merge is a programmatic dependent of the tcgen05 kernel whenever both run in this call
  return launch_merge(bv, q_rects, n_obj, h, w, n_splits, umma, W, mem_val, umma && (stages & RMNET_STAGE_PARTIAL), st);
}

}  // namespace
}  // namespace rmnet

using namespace rmnet;
extern "C" {

int rmnet_has_umma(void) { return umma_supported(64) ? 1 : 0; }

size_t rmnet_memory_read_workspace_bytes(int n_obj, int h, int w, int cap_cells) {
  if (n_obj <= 0 || h <= 0 || w <= 0 || cap_cells <= 0) return 0;
  return read_workspace(nullptr, n_obj, h * w, READ_MAX_SPLITS).total;  // the plan may cut an object into that many KV chunks (partial slots)
}

int rmnet_bank_memory_read(const void *bank, size_t bank_bytes, int n_slots, int cap_cells, const float *q_key,
                           const float *q_val, long long q_obj_stride, const int *q_rects, int n_obj, int h, int w,
                           int elem_format, int precision, int impl, int stages, float *mem_val, void *workspace,
                           size_t workspace_bytes, void *stream) {
  return bank_memory_read_impl(bank, bank_bytes, n_slots, cap_cells, q_key, q_val, q_obj_stride, q_rects, n_obj, h, w, elem_format,
                               precision, impl, stages, mem_val, workspace, workspace_bytes, false, stream);
}

int rmnet_memory_read_plan_host(const void *workspace, int n_obj, int h, int w, int *ns_out, int *hdr_out, int *pieces_out,
                                int max_pieces, int *n_ctas_out, void *stream) {
  RMNET_CHECK_ARG(workspace && ns_out && hdr_out && pieces_out && n_ctas_out, "null pointer argument");
  RMNET_CHECK_ARG(n_obj > 0 && n_obj <= SCHED_MAX_OBJ && h > 0 && w > 0 && max_pieces > 0, "bad argument");
  ReadWorkspace W = rmnet::read_workspace(const_cast<void *>(workspace), n_obj, h * w, READ_MAX_SPLITS);
  cudaStream_t st = (cudaStream_t)stream;
  const int G = umma_grid_size();
  const size_t n_pieces = W.plan_piece_cap < (size_t)max_pieces ? W.plan_piece_cap : (size_t)max_pieces;
  RMNET_CUDA(cudaMemcpyAsync(ns_out, W.sched, (size_t)n_obj * sizeof(int), cudaMemcpyDeviceToHost, st));
  RMNET_CUDA(cudaMemcpyAsync(hdr_out, W.plan_hdr, (size_t)G * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  RMNET_CUDA(cudaMemcpyAsync(pieces_out, W.plan_pieces, n_pieces * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  RMNET_CUDA(cudaStreamSynchronize(st));
  *n_ctas_out = G;
  return RMNET_OK;
}

int rmnet_frame_step(void *bank, size_t bank_bytes, int n_slots, int cap_cells, const float *prev_mask, const float *flow,
                     int K, int H, int W, int sampler, float prob_threshold, int n_pts_threshold, int n_bbox_loose_pixels,
                     int pad_l, int pad_r, int pad_t, int pad_b, int k_scan, const float *k4, const float *v4, const float *q_key,
                     const float *q_val, int n_obj, int elem_format, int precision, int impl, int commit, int *boxes_out,
                     float *mem_val, void *box_workspace, size_t box_workspace_bytes, void *read_workspace,
                     size_t read_workspace_bytes, void *stream) {
  RMNET_CHECK_ARG(boxes_out && n_obj > 0 && n_obj < K, "bad argument (need 0 < n_obj < K)");
  const int h = (H + pad_t + pad_b) / 16, w = (W + pad_l + pad_r) / 16;
  const long long N = (long long)h * w;
  int *mem_bb = boxes_out, *mem_rc = boxes_out + 4 * K, *cur_bb = boxes_out + 8 * K, *cur_rc = boxes_out + 12 * K;
  RMNET_CHECK_ARG(bank && n_slots > 0 && cap_cells > 0 && n_obj <= n_slots, "bad bank argument");
  BankLayout L = bank_layout(n_slots, cap_cells);
  if (bank_bytes < L.total) { set_error("bank too small"); return RMNET_E_WORKSPACE; }
  BankView bv = bank_view(bank, n_slots, cap_cells);
  // One programmatic-dependent-launch chain: regions (also zeroes the temporary value sums) -> pack [-> commit] -> tcgen05
  // read -> merge.  Each kernel's launch latency and independent prologue overlap its predecessor's tail.
  int rc = frame_regions_chain_head(prev_mask, flow, 1, K, H, W, sampler, prob_threshold, n_pts_threshold, n_bbox_loose_pixels,
                                    pad_l, pad_r, pad_t, pad_b, k_scan, mem_bb, mem_rc, cur_bb, cur_rc, box_workspace,
                                    box_workspace_bytes, reinterpret_cast<float *>(bv.vsum + (size_t)n_slots * RMNET_CV),
                                    2 * n_slots * RMNET_CV /* 4-byte words of the temporary frame's i64 value sums */, stream);
  if (rc) return rc;
  RMNET_CHECK_ARG(q_key && q_val && mem_val && read_workspace, "null pointer argument");
  RMNET_CHECK_ARG(impl == RMNET_IMPL_AUTO || impl == RMNET_IMPL_UMMA || impl == RMNET_IMPL_SIMT, "unknown impl %d", impl);
  const bool umma = impl == RMNET_IMPL_UMMA || (impl == RMNET_IMPL_AUTO && umma_supported(cap_cells) && n_obj <= SCHED_MAX_OBJ);
  ReadWorkspace RW = rmnet::read_workspace(read_workspace, n_obj, (int)N, umma ? READ_MAX_SPLITS : pick_splits(n_obj, (int)N, 64, cap_cells));
  if (read_workspace_bytes < RW.total) { set_error("workspace too small: %zu < %zu", read_workspace_bytes, RW.total); return RMNET_E_WORKSPACE; }
  // the pack launch also prepares the query side (packed query keys, q_val passthrough) of this frame's read
  QuerySide qs = make_query_side(q_key, q_val, 0, cur_rc + 4, RW, (int)N, mem_val, bv.meta + META_RANGE, umma, bv, precision);
  rc = bank_memorize_impl(bank, bank_bytes, n_slots, cap_cells, k4, RMNET_CK * N, N, v4, RMNET_CV * N, N, mem_rc + 4, n_obj, h, w,
                          elem_format, commit, /*chained=*/true, &qs, stream);
  if (rc) return rc;
  return bank_memory_read_impl(bank, bank_bytes, n_slots, cap_cells, q_key, q_val, 0, cur_rc + 4, n_obj, h, w, elem_format,
                               precision, impl, RMNET_STAGE_PARTIAL | RMNET_STAGE_MERGE, mem_val, read_workspace,
                               read_workspace_bytes, /*chained=*/true, stream);
}

static size_t reader_scratch_layout(int n, int T, int h, int w, size_t *off_rects, size_t *off_read, int *cap) {
  const int N = h * w;
  *cap = cdiv(T * N, 64) * 64;
  size_t o = align_up(bank_layout(n, *cap).total, 1024);
  *off_rects = o; o = align_up(o + (size_t)n * 32, 1024);
  *off_read = o; o += rmnet_memory_read_workspace_bytes(n, h, w, *cap);
  return o;
}

size_t rmnet_memory_reader_workspace_bytes(int n, int T, int h, int w) {
  if (n <= 0 || T <= 0 || h <= 0 || w <= 0) return 0;
  size_t a, b; int cap;
  return reader_scratch_layout(n, T, h, w, &a, &b, &cap);
}

int rmnet_memory_reader_forward(const float *m_key, const float *m_val, const float *q_key, const float *q_val, int n,
                                int T, int h, int w, int elem_format, int precision, int impl, float *mem_val, float *p,
                                void *workspace, size_t workspace_bytes, void *stream) {
  RMNET_CHECK_ARG(m_key && m_val && q_key && q_val && mem_val && workspace, "null pointer argument");
  RMNET_CHECK_ARG(n > 0 && T > 0 && h > 0 && w > 0, "bad shape");
  RMNET_CHECK_ARG((uintptr_t)workspace % 1024 == 0, "workspace must be 1024-byte aligned");
  size_t off_rects, off_read; int cap;
  const size_t need = reader_scratch_layout(n, T, h, w, &off_rects, &off_read, &cap);
  if (workspace_bytes < need) { set_error("workspace too small: %zu < %zu", workspace_bytes, need); return RMNET_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h * w;
  char *ws = (char *)workspace;
  const size_t bank_bytes = bank_layout(n, cap).total;
  int *rects = (int *)(ws + off_rects);            // [n] query rectangles (dense h x w), then [n] memory rectangles
  int *mem_rects = rects + 4 * n;
  int rc = rmnet_bank_reset(ws, bank_bytes, n, cap, stream);
  if (rc) return rc;
  fill_dense_rects_kernel<<<cdiv(n, 128), 128, 0, st>>>(rects, n, h, w);
  fill_dense_rects_kernel<<<cdiv(n, 128), 128, 0, st>>>(mem_rects, n, T * h, w);
  RMNET_LAUNCH_CHECK();
  // ONE pack launch for the whole memory: [n,C,T,h,w] is a dense (T*h) x w cell grid with channel stride T*h*w (the
  // literal signature carries no region information, so every cell is stored; position t*N + y*w + x, the reference's
  // own flattening, :151)
  rc = rmnet_bank_memorize(ws, bank_bytes, n, cap, m_key, (long long)RMNET_CK * T * N, (long long)T * N, m_val,
                           (long long)RMNET_CV * T * N, (long long)T * N, mem_rects, n, T * h, w, elem_format, /*commit=*/1, stream);
  if (rc) return rc;
  rc = rmnet_bank_memory_read(ws, bank_bytes, n, cap, q_key, q_val, (long long)RMNET_CK * N, rects, n, h, w, elem_format,
                              precision, impl, RMNET_STAGE_ALL, mem_val, ws + off_read, workspace_bytes - off_read, stream);
  if (rc || p == nullptr) return rc;
  // the reference's second output, on request: fp32 scores from the raw keys (attention_probs.cu)
  return launch_attention_probs(m_key, q_key, n, T * N, N, p, st);
}
}
