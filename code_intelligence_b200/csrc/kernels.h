// Internal launch interfaces between the C-ABI layer (api.cu) and the sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ie {

// ---- TMA descriptor helper (tmap.cu) ---------------------------------------------------------
// 2-D bf16 tensor, inner (contiguous) dimension `inner` elements, `rows` rows, row pitch `ld` elements,
// box {box_inner, box_rows}, 128-byte swizzle (box_inner must be 64).
cudaError_t make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t ld,
                              uint32_t box_inner, uint32_t box_rows);

// ---- GEMM (gemm.cu) ----------------------------------------------------------------------------
struct GemmArgs {
  const __nv_bfloat16* a;  // [m_pad, lda]
  const __nv_bfloat16* b;  // [n_pad, ldb]
  void* d;                 // [m_store.., ldd] f32 or bf16
  const float* bias;       // [n_pad] or nullptr
  int m_pad, n_pad, k_pad;
  long long lda, ldb, ldd;
  int m_store, n_store;
  int bn;        // N tile (multiple of 16, <= 256, divides n_pad)
  int act;       // 0 none, 1 relu, 2 sigmoid
  int out_bf16;  // 0 -> f32 output, 1 -> bf16 output
  int num_sms;
};
cudaError_t launch_gemm_bf16(const GemmArgs& g, cudaStream_t stream);

// ---- recurrent LSTM step (lstm.cu) ---------------------------------------------------------
struct LstmStepArgs {
  CUtensorMap tm_h;  // hidden-state slots  [(T+1)*b_pad rows, kh_pad], box {64, 128}
  CUtensorMap tm_hs; // same tensor, box {64, 128/cluster}: the slice of a tile one CTA multicasts to its cluster
  CUtensorMap tm_w;  // sliced W_hh         [4*out_pad rows,  kh_pad], box {64, 4*u}
  const float* gx;   // [T*b_pad, 4*out_pad] sliced column order, bias folded in
  float* c;          // [b_pad, out_pad] cell state
  __nv_bfloat16* y;  // [(T+1)*b_pad, ldy] hidden-state slots (slot 0 = zeros)
  float* raw;        // optional [b_pad, T, raw_ld] f32 copy of h (get_raw_features), or nullptr
  float* pool_sum;   // optional [b_pad, out_pad] (last layer only)
  float* pool_max;
  float* pool_last;
  const int* lengths;  // [b_pad]
  int t, T;
  int b_pad;   // 128 or 256
  int u;       // hidden units per CTA (multiple of 4)
  int n_cta;   // out_pad / u
  int cluster; // CTAs per cluster sharing the h tiles by TMA multicast (1, 2, 4 or 8; divides n_cta)
  int fast_math; // 1: single-MUFU tanh.approx gates (same as the persistent kernel)
  int out_pad;
  int kh_pad;  // multiple of 64
  long long ldy;
  long long raw_ld;
};
cudaError_t launch_lstm_step(const LstmStepArgs& a, cudaStream_t stream);

// ---- persistent recurrent layer (lstm_seq.cu): all T steps in one launch, CTA pairs (cta_group::2) -------------
struct LstmSeqArgs {
  CUtensorMap tm_h;  // hidden-state slots  [(T+1)*256 rows, kh_pad], box {64, 128}
  CUtensorMap tm_w;  // sliced W_hh         [4*out_pad rows, kh_pad], box {64, 4*u}
  const float* gx;
  __nv_bfloat16* y;
  float* raw;
  float* pool_sum;
  float* pool_max;
  float* pool_last;
  const int* lengths;
  unsigned* step_done;  // [T] zero-initialised grid-barrier counters
  int T, b_pad, u, n_cta, out_pad, kh_pad;
  long long ldy, raw_ld;
  int check_only;  // 1: only check that the grid can be co-resident
  int fast_math;   // 1: single-MUFU tanh.approx gates
  long long* trace; // optional [n_cta][T][8] SM-clock timeline (debug), or nullptr
};
cudaError_t launch_lstm_seq(const LstmSeqArgs& a, cudaStream_t stream);

// ---- persistent recurrent layer, wide tiles (lstm_wide.cu): N = 256 per CTA pair, up to 3 batches per launch -------
struct LstmWideArgs {
  CUtensorMap tm_h;  // hidden-state slots  [(T+1)*256*ng rows, kh_pad], box {64, 128}
  CUtensorMap tm_w;  // sliced W_hh (u = 32) [4*out_pad rows, kh_pad], box {64, 128}
  const float* gx;
  float* c;          // [256*ng, out_pad] cell state
  __nv_bfloat16* y;
  float* raw;
  float* pool_sum;
  float* pool_max;
  float* pool_last;
  const int* lengths;
  unsigned* step_done;  // [T*ng] zero-initialised
  int T, ng, u, n_cta, out_pad, kh_pad;
  long long ldy, raw_ld;
  int fast_math, num_sms, check_only;
  long long* trace;  // optional [grid][T][12] timeline (debug)
  int trace_items;   // lstm_rot.cu only: the timeline is [grid][trace_items][12], one record per work item of the pair
  int variant;       // lstm_rot.cu only (IE_ROT_VARIANT): bit 0 proxy fence in the watcher warp, bit 1 four-stage h ring
  const int* tok;    // optional [T*256*ng] time-major token ids: Gx row of (t, row) is gx[tok[t*b_pad+row]] (per-token
                     // input-projection table of layer 0, api.cu IE_EMB_PROJ) instead of gx[t*b_pad+row]
};
cudaError_t launch_lstm_wide(const LstmWideArgs& a, cudaStream_t stream);

// ---- persistent recurrent layer, rotating item schedule (lstm_rot.cu; experimental, IE_ROT=1): same arguments, up to
//      kRotMaxBatches batches of 256 rows per launch; (timestep, batch, tile) items dealt round-robin over all CTA pairs
constexpr int kRotMaxBatches = 8;  // compile-time bound; api.cu uses 5 per launch unless IE_ROT_BATCHES says otherwise
cudaError_t launch_lstm_rot(const LstmWideArgs& a, cudaStream_t stream);
int lstm_rot_pairs(const LstmWideArgs& a);  // CTA pairs the launch will use

// ---- UMMA issue/throughput micro-benchmark (umma_bench.cu, debug) ------------------------------------------------
cudaError_t run_umma_rate(int mode, int n, int iters, int commit_every, int grid, int ntiles, long long* host_out);

// ---- small memory-bound kernels (misc.cu) --------------------------------------------------------
// ids [B, T] int64 (batch-first, right padded) -> x0 [(T*b_pad), ldx] bf16, time-major rows t*b_pad + b
cudaError_t launch_embed_gather(const int64_t* ids, int B, int T, int b_pad, const __nv_bfloat16* emb, int vocab,
                                int e_pad, __nv_bfloat16* x0, long long ldx, int pad_idx, int* err_flag,
                                cudaStream_t stream);
// out[b] = [sum/len | max | last], b < B, first `e` units
// ids [B, T] int64 -> tok [T*b_pad] int32 time-major (rows >= B: pad_idx), range-checked like launch_embed_gather
cudaError_t launch_tokens_time_major(const int64_t* ids, int B, int T, int b_pad, int vocab, int pad_idx, int* tok,
                                     int* err_flag, cudaStream_t stream);
cudaError_t launch_pool_finalize(const float* pool_sum, const float* pool_max, const float* pool_last,
                                 const int* lengths, int B, int e, int out_pad, float* out, cudaStream_t stream);
// same result from the last layer's f32 hidden states raw [.., T, raw_ld] (sequential over t: identical bits)
cudaError_t launch_pool_from_raw(const float* raw, const int* lengths, int B, int T, int e, long long raw_ld, float* out,
                                 cudaStream_t stream);
// f32 [rows, cols] (row pitch ld_src) -> bf16 [rows_pad, ld_dst] with optional row permutation (src row of dst row r
// = perm[r], or -1 for a zero row); columns >= cols zero filled.
cudaError_t launch_convert_rows(const float* src, long long ld_src, int cols, const int* perm, int rows_dst,
                                __nv_bfloat16* dst, long long ld_dst, cudaStream_t stream);
cudaError_t launch_fill_f32(float* p, size_t n, float v, cudaStream_t stream);

}  // namespace ie
