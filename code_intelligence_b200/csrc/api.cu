// C-ABI layer of libissue_emb_b200.so (declared in include/issue_emb_b200.h): handle management, weight
// re-layout, workspace, and the launch sequence of the encoder hot path and the MLP head.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/issue_emb_b200.h"
#include "kernels.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  const int code = (e == cudaErrorMemoryAllocation) ? IE_ERR_OOM : IE_ERR_CUDA;
  if (e == cudaErrorMemoryAllocation) cudaGetLastError();  // clear the sticky-free OOM
  return fail(code, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

#define CK(expr)                                              \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return cuda_fail(_e, #expr);       \
  } while (0)

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// grow-only device buffer; owns its allocation (freed on destruction, so error paths do not leak)
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  cudaError_t reserve(size_t bytes, bool zero = false) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { p = nullptr; return e; }
    cap = bytes;
    if (zero) {
      // the handle's streams are non-blocking: make the (legacy-stream) memset complete before anyone uses it
      e = cudaMemset(p, 0, bytes);
      if (e != cudaSuccess) return e;
      return cudaDeviceSynchronize();
    }
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Layer {
  int in = 0, out = 0;      // logical dims
  int u = 0, n_cta = 0;     // hidden units per CTA, CTAs per step
  int out_pad = 0;          // n_cta * u
  int kin_pad = 0;          // padded K of the input projection
  int kh_pad = 0;           // padded K of the recurrent projection
  int bn = 0;               // GEMM N tile for the input projection
  int cluster = 1;          // CTAs per cluster in the recurrent step (h tiles shared by TMA multicast)
  DevBuf w_ih, w_hh, bias;  // sliced layouts
  bool loaded = false;
};

}  // namespace

struct ie_encoder {
  ie_config cfg{};
  int num_sms = 148;
  int e_pad = 0;  // emb_sz rounded up to 64
  std::vector<Layer> layers;   // plan A: u ~ out/148 units per CTA (N = 160 pair tiles at H = 2400): B <= 512
  std::vector<Layer> layersB;  // plan B: u = 32 (N = 256 pair tiles): three batches per launch (512 < B <= 768)
  int use_wide = 1, wide_checked = 0;
  // experimental (IE_ROT=1): rotating item schedule (lstm_rot.cu), up to kRotMaxBatches batches per launch on plan B;
  // IE_ROT=2 also routes 256..768 rows through it (for testing).  max_batch is what ie_encoder_encode accepts.
  int use_rot = 0, rot_checked = 0, rot_variant = 0;  // rot_variant: IE_ROT_VARIANT (LstmWideArgs::variant)
  int max_batch = IE_MAX_BATCH;
  // experimental (IE_EMB_PROJ=1): layer 0's input projection W_ih0 . Emb[id] + b is a function of the token id alone,
  // so it is tabulated once per weight set (proj: [vocab_pad, 4*out_pad] f32, plan-B column order, computed by the same
  // GEMM from the same bf16 operands => the same bits) and the wide / rotating kernels read row tok[t, b] of it:
  // no embedding gather, no layer-0 GEMM, no Gx write for that layer.
  int use_proj = 0;
  // experimental (IE_POOL_RAW=1): in wide / rotating calls the last layer writes its f32 h_t to `raw` and a separate
  // kernel pools it (same sequential sums => same bits); the recurrent epilogue carries no pooling accumulators
  int use_pool_raw = 0;
  bool proj_built = false;
  DevBuf proj, tok;
  DevBuf emb;  // bf16 [vocab, e_pad]
  bool emb_loaded = false;
  // workspace
  DevBuf ids, lengths, x0, y[2], gx, c, pool_sum, pool_max, pool_last, out, raw, err, step_done;
  DevBuf trace;           // debug timeline of one layer of the persistent kernel (ie_debug_seq_trace)
  int trace_layer = -1;
  int trace_T = 0, trace_ctas = 0;
  int fast_math = 1;      // tanh.approx gates in the persistent kernel (ie_config.flags & IE_CFG_ACCURATE_GATES: off)
  int use_seq = 1;        // persistent per-layer kernel (lstm_seq.cu) when B_pad == 256 and the grid is co-resident
  int seq_checked = 0;    // co-residency verified for every layer
  long long y_ld = 0;
  cudaStream_t own_stream = nullptr;
  int64_t launches = 0;
  // phase boundary events of the last encode call: start, gather, (gemm_l, steps_l) x L, finalize
  std::vector<cudaEvent_t> ev;
  int ev_used = 0;
  int last_T = 0, last_b_pad = 0;
  std::mutex mu;
};

struct ie_mlp {
  int device = 0;
  int num_sms = 148;
  std::vector<int> dims;
  struct L {
    int k_pad = 0, n_pad = 0, bn = 0;
    DevBuf w, b;
    bool loaded = false;
  };
  std::vector<L> layers;
  DevBuf xf, act[2], probs;
  long long act_ld = 0;
  cudaStream_t own_stream = nullptr;
  std::mutex mu;
};

namespace {

constexpr int kStepStride = ie::kRotMaxBatches;  // step counters per (layer, timestep): one per batch of the launch

int plan_layers(ie_encoder* h, std::vector<Layer>& layers, int u_fixed) {
  const ie_config& c = h->cfg;
  h->e_pad = static_cast<int>(round_up(c.emb_sz, 64));
  layers.resize(c.n_layers);
  int prev_pad = h->e_pad;
  for (int l = 0; l < c.n_layers; ++l) {
    Layer& L = layers[l];
    L.in = (l == 0) ? c.emb_sz : c.n_hid;
    L.out = (l == c.n_layers - 1) ? c.emb_sz : c.n_hid;
    // u hidden units per CTA: multiple of 4, as many CTAs as fit on the SMs (one wave)
    L.u = u_fixed > 0 ? u_fixed : 4 * static_cast<int>((L.out + 4ll * h->num_sms - 1) / (4ll * h->num_sms));
    if (L.u > 32) return fail(IE_ERR_INVALID, "hidden size %d too large for %d SMs", L.out, h->num_sms);
    L.n_cta = (L.out + L.u - 1) / L.u;
    if (u_fixed > 0 && (L.n_cta & 1)) ++L.n_cta;  // CTA pairs: the last pair's second slice is pure padding
    L.out_pad = L.n_cta * L.u;
    L.kin_pad = prev_pad;
    L.kh_pad = static_cast<int>(round_up(L.out_pad, 64));
    prev_pad = L.kh_pad;
    // cluster size for the h-tile multicast in the per-step fallback kernel: largest of 8/4/2/1 <= want dividing n_cta.
    // Measured (profiles/README.md): multicast changes nothing there, so the default is no clusters.
    int want = 1;
    if (const char* e = getenv("IE_STEP_CLUSTER")) want = atoi(e);
    L.cluster = 1;
    for (int cs = 8; cs >= 1; cs >>= 1)
      if (cs <= want && L.n_cta % cs == 0) { L.cluster = cs; break; }
    // largest multiple of 16 <= 256 dividing 4*out_pad
    const int n = 4 * L.out_pad;
    L.bn = 16;
    for (int b = 256; b >= 16; b -= 16)
      if (n % b == 0) { L.bn = b; break; }
  }
  return IE_OK;
}

// torch gate-major rows [4*out] -> sliced rows [cta][unit][gate]; -1 marks zero padding rows
std::vector<int> slice_perm(const Layer& L) {
  std::vector<int> perm(4 * static_cast<size_t>(L.out_pad));
  for (int j = 0; j < L.n_cta; ++j)
    for (int i = 0; i < L.u; ++i)
      for (int g = 0; g < 4; ++g) {
        const int unit = j * L.u + i;
        perm[(static_cast<size_t>(j) * L.u + i) * 4 + g] = unit < L.out ? g * L.out + unit : -1;
      }
  return perm;
}

int upload_sliced(const float* host, int rows_src, int cols, const std::vector<int>& perm, int ld_dst, DevBuf& dst,
                  cudaStream_t s) {
  DevBuf tmp, dperm;
  CK(tmp.reserve(static_cast<size_t>(rows_src) * cols * sizeof(float)));
  CK(dperm.reserve(perm.size() * sizeof(int)));
  CK(cudaMemcpyAsync(tmp.p, host, static_cast<size_t>(rows_src) * cols * sizeof(float), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dperm.p, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  CK(dst.reserve(perm.size() * static_cast<size_t>(ld_dst) * sizeof(__nv_bfloat16)));
  CK(ie::launch_convert_rows(tmp.as<float>(), cols, cols, dperm.as<int>(), static_cast<int>(perm.size()),
                             dst.as<__nv_bfloat16>(), ld_dst, s));
  CK(cudaStreamSynchronize(s));
  tmp.release();
  dperm.release();
  return IE_OK;
}

int ensure_workspace(ie_encoder* h, int b_pad, int T, bool want_raw) {
  const ie_config& c = h->cfg;
  long long max_out_pad = 0, max_kh = 0;
  for (const std::vector<Layer>* plan : {&h->layers, &h->layersB})
    for (const Layer& L : *plan) {
      max_out_pad = std::max<long long>(max_out_pad, L.out_pad);
      max_kh = std::max<long long>(max_kh, L.kh_pad);
    }
  const long long rows = static_cast<long long>(T) * b_pad;
  CK(h->ids.reserve(static_cast<size_t>(h->max_batch) * T * sizeof(int64_t)));
  CK(h->lengths.reserve(h->max_batch * sizeof(int)));
  CK(h->err.reserve(sizeof(int), true));
  CK(h->x0.reserve(static_cast<size_t>(rows) * h->e_pad * sizeof(__nv_bfloat16)));
  // hidden-state rings: (T+1) slots of b_pad rows; slot 0 and the K padding columns must be zero
  h->y_ld = max_kh;
  const size_t ybytes = static_cast<size_t>(rows + b_pad) * max_kh * sizeof(__nv_bfloat16);
  for (int i = 0; i < 2; ++i) CK(h->y[i].reserve(ybytes, /*zero=*/true));
  CK(h->gx.reserve(static_cast<size_t>(rows) * 4 * max_out_pad * sizeof(float)));
  CK(h->c.reserve(static_cast<size_t>(h->max_batch) * max_out_pad * sizeof(float)));
  const size_t pb = static_cast<size_t>(h->max_batch) * max_out_pad * sizeof(float);
  CK(h->pool_sum.reserve(pb));
  CK(h->pool_max.reserve(pb));
  CK(h->pool_last.reserve(pb));
  CK(h->out.reserve(static_cast<size_t>(h->max_batch) * 3 * c.emb_sz * sizeof(float)));
  if (want_raw) CK(h->raw.reserve(static_cast<size_t>(b_pad) * T * max_out_pad * sizeof(float)));
  CK(h->step_done.reserve(static_cast<size_t>(c.n_layers) * T * kStepStride * sizeof(unsigned)));
  if (h->use_proj) CK(h->tok.reserve(static_cast<size_t>(rows) * sizeof(int)));
  return IE_OK;
}

int mark(ie_encoder* h, cudaStream_t s) {
  if (h->ev_used >= static_cast<int>(h->ev.size())) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    h->ev.push_back(e);
  }
  CK(cudaEventRecord(h->ev[h->ev_used++], s));
  return IE_OK;
}

// IE_EMB_PROJ: tabulate layer 0's input projection for every token id (once per weight set)
int build_proj_table(ie_encoder* h, cudaStream_t s) {
  const Layer& L = h->layersB[0];
  const long long v_pad = round_up(h->cfg.vocab_sz, 256);
  if (h->emb.cap < static_cast<size_t>(v_pad) * h->e_pad * sizeof(__nv_bfloat16))
    return fail(IE_ERR_STATE, "embedding was loaded without row padding (IE_EMB_PROJ must be set before loading)");
  CK(h->proj.reserve(static_cast<size_t>(v_pad) * 4 * L.out_pad * sizeof(float)));
  ie::GemmArgs g{};
  g.a = h->emb.as<__nv_bfloat16>();
  g.lda = h->e_pad;
  g.b = L.w_ih.as<__nv_bfloat16>();
  g.ldb = L.kin_pad;
  g.d = h->proj.p;
  g.ldd = 4ll * L.out_pad;
  g.bias = L.bias.as<float>();
  g.m_pad = static_cast<int>(v_pad);
  g.n_pad = 4 * L.out_pad;
  g.k_pad = L.kin_pad;
  g.m_store = static_cast<int>(v_pad);
  g.n_store = 4 * L.out_pad;
  g.bn = L.bn;
  g.act = 0;
  g.out_bf16 = 0;
  g.num_sms = h->num_sms;
  CK(ie::launch_gemm_bf16(g, s));
  h->launches++;
  h->proj_built = true;
  return IE_OK;
}

// the launch sequence shared by encode (pooled) and raw_features
int run_encoder(ie_encoder* h, const int64_t* ids, const int32_t* lengths, int B, int T, float* out, float* raw_out,
                int flags, cudaStream_t s) {
  const ie_config& c = h->cfg;
  if (!h->emb_loaded) return fail(IE_ERR_STATE, "embedding not loaded");
  for (const Layer& L : h->layers)
    if (!L.loaded) return fail(IE_ERR_STATE, "LSTM layer weights not loaded");
  for (const Layer& L : h->layersB)
    if (!L.loaded) return fail(IE_ERR_STATE, "LSTM layer weights not loaded");
  if (B < 1 || B > h->max_batch) return fail(IE_ERR_INVALID, "B=%d outside [1,%d]", B, h->max_batch);
  if (T < 1) return fail(IE_ERR_INVALID, "T=%d must be >= 1", T);
  if (ids == nullptr || (out == nullptr && raw_out == nullptr)) return fail(IE_ERR_INVALID, "null pointer");
  const bool dev = (flags & IE_FLAG_DEVICE_PTRS) != 0;
  const bool pooled = out != nullptr;
  const int b_pad = B <= 128 ? 128 : static_cast<int>(round_up(B, 256));
  if (h->use_rot && !h->rot_checked) {  // every CTA pair of the rotating-schedule kernel must be co-resident
    CK(cudaSetDevice(c.device));
    for (const Layer& L : h->layersB) {
      ie::LstmWideArgs q{};
      q.T = 1; q.ng = 1; q.u = L.u; q.n_cta = L.n_cta; q.out_pad = L.out_pad; q.kh_pad = L.kh_pad;
      q.num_sms = h->num_sms; q.check_only = 1; q.variant = h->rot_variant;
      if (ie::launch_lstm_rot(q, s) != cudaSuccess) h->use_rot = 0;
    }
    cudaGetLastError();
    h->rot_checked = 1;
  }
  const bool rot = h->use_rot && b_pad >= 256 && (b_pad > 768 || h->use_rot >= 2);
  const bool wide = (b_pad == 768) && !rot;
  if (b_pad > 768 && !rot) return fail(IE_ERR_STATE, "B > 768 needs the rotating-schedule kernel (caller splits the batch)");
  std::vector<Layer>& LS = (wide || rot) ? h->layersB : h->layers;
  // workspace cap (tokens per call); IE_MAX_TOKENS lowers it, e.g. to exercise the caller's batch-halving loop
  long long cap = 1ll << 21;
  if (const char* e = getenv("IE_MAX_TOKENS")) cap = std::max(128ll, atoll(e));
  if (static_cast<long long>(b_pad) * T > cap)
    return fail(IE_ERR_OOM, "B_pad*T = %lld tokens exceeds the workspace cap %lld; use a smaller batch",
                static_cast<long long>(b_pad) * T, cap);
  CK(cudaSetDevice(c.device));
  const bool pool_raw = h->use_pool_raw && pooled && raw_out == nullptr && (wide || rot);
  int rc = ensure_workspace(h, b_pad, T, raw_out != nullptr || pool_raw);
  if (rc != IE_OK) return rc;

  // lengths: validate on the host when we can; padded rows get length 1
  std::vector<int> len_host(h->max_batch, 1);
  if (pooled) {
    if (lengths == nullptr) return fail(IE_ERR_INVALID, "lengths is null");
    if (!dev) {
      for (int b = 0; b < B; ++b) {
        if (lengths[b] < 1 || lengths[b] > T)
          return fail(IE_ERR_INVALID, "lengths[%d]=%d outside [1,%d]", b, lengths[b], T);
        len_host[b] = lengths[b];
      }
      CK(cudaMemcpyAsync(h->lengths.p, len_host.data(), h->max_batch * sizeof(int), cudaMemcpyHostToDevice, s));
    } else {
      CK(cudaMemcpyAsync(h->lengths.p, len_host.data(), h->max_batch * sizeof(int), cudaMemcpyHostToDevice, s));
      CK(cudaMemcpyAsync(h->lengths.p, lengths, B * sizeof(int), cudaMemcpyDeviceToDevice, s));
    }
  }
  const int64_t* ids_dev = ids;
  if (!dev) {
    CK(cudaMemcpyAsync(h->ids.p, ids, static_cast<size_t>(B) * T * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    ids_dev = h->ids.as<int64_t>();
  }

  h->ev_used = 0;
  h->last_T = T;
  h->last_b_pad = b_pad;
  // layer 0 from the per-token projection table (only the wide / rotating kernels take the token indirection)
  const bool proj = h->use_proj && (wide || rot) && c.n_layers > 1;
  if (proj && !h->proj_built && (rc = build_proj_table(h, s)) != IE_OK) return rc;
  if ((rc = mark(h, s)) != IE_OK) return rc;
  if (proj)
    CK(ie::launch_tokens_time_major(ids_dev, B, T, b_pad, c.vocab_sz, c.pad_idx, h->tok.as<int>(), h->err.as<int>(), s));
  else
    CK(ie::launch_embed_gather(ids_dev, B, T, b_pad, h->emb.as<__nv_bfloat16>(), c.vocab_sz, h->e_pad,
                               h->x0.as<__nv_bfloat16>(), h->e_pad, c.pad_idx, h->err.as<int>(), s));
  h->launches++;
  if ((rc = mark(h, s)) != IE_OK) return rc;

  const long long rows = static_cast<long long>(T) * b_pad;
  // persistent per-layer recurrent kernel: needs both 128-row halves and every CTA of a layer co-resident
  bool seq = h->use_seq && b_pad >= 256 && !wide && !rot;
  if (wide) {
    if (h->use_wide && !h->wide_checked) {
      for (const Layer& L : h->layersB) {
        ie::LstmWideArgs q{};
        q.T = 1; q.ng = 3; q.u = L.u; q.n_cta = L.n_cta; q.out_pad = L.out_pad; q.kh_pad = L.kh_pad;
        q.num_sms = h->num_sms; q.check_only = 1;
        if (ie::launch_lstm_wide(q, s) != cudaSuccess) h->use_wide = 0;
      }
      cudaGetLastError();
      h->wide_checked = 1;
    }
    if (!h->use_wide) return fail(IE_ERR_STATE, "B > 512 needs the wide persistent kernel (caller splits the batch)");
  }
  if ((seq || wide) && !h->seq_checked) {  // (not needed by the rotating schedule, which runs every layer itself)
    for (const Layer& L : h->layers) {
      ie::LstmSeqArgs q{};
      q.T = 1; q.b_pad = 256; q.u = L.u; q.n_cta = L.n_cta; q.out_pad = L.out_pad; q.kh_pad = L.kh_pad;
      q.check_only = 1;
      if (L.n_cta % 2 || ie::launch_lstm_seq(q, s) != cudaSuccess) { h->use_seq = 0; seq = false; }
    }
    cudaGetLastError();
    h->seq_checked = 1;
  }
  if (b_pad == 512 && !seq && !rot) return fail(IE_ERR_STATE, "B > 256 needs the persistent kernel (caller splits the batch)");
  if (seq || wide || rot)
    CK(cudaMemsetAsync(h->step_done.p, 0, static_cast<size_t>(c.n_layers) * T * kStepStride * sizeof(unsigned), s));
  // slot 0 of both hidden-state rings is h_{-1} = 0; a previous call with another B_pad may have written these rows
  for (int i = 0; i < 2; ++i)
    CK(cudaMemsetAsync(h->y[i].p, 0, static_cast<size_t>(b_pad) * h->y_ld * sizeof(__nv_bfloat16), s));
  int cur = 0;
  const __nv_bfloat16* layer_in = h->x0.as<__nv_bfloat16>();
  long long layer_in_ld = h->e_pad;
  // with three batches per launch the pooled last layer (narrow, latency bound) still runs the lstm_seq.cu kernel on its
  // plan-A layout when that is possible; all other layers use the wide tiles of plan B
  const Layer& lastA = h->layers.back();
  const bool last_on_seq = wide && !rot && !pool_raw && pooled && raw_out == nullptr && h->use_seq && lastA.u <= 12 && lastA.n_cta % 2 == 0 &&
                           (c.n_layers < 2 || lastA.kin_pad == h->layersB[c.n_layers - 1].kin_pad);
  for (int l = 0; l < c.n_layers; ++l) {
    const bool last = (l == c.n_layers - 1);
    Layer& L = (last && last_on_seq) ? h->layers[l] : LS[l];
    // hoisted input projection over all T*b_pad rows
    ie::GemmArgs g{};
    g.a = layer_in;
    g.lda = layer_in_ld;
    g.b = L.w_ih.as<__nv_bfloat16>();
    g.ldb = L.kin_pad;
    g.d = h->gx.p;
    g.ldd = 4ll * L.out_pad;
    g.bias = L.bias.as<float>();
    g.m_pad = static_cast<int>(rows);
    g.n_pad = 4 * L.out_pad;
    g.k_pad = L.kin_pad;
    g.m_store = static_cast<int>(rows);
    g.n_store = 4 * L.out_pad;
    g.bn = L.bn;
    g.act = 0;
    g.out_bf16 = 0;
    g.num_sms = h->num_sms;
    const bool from_table = proj && l == 0;  // Gx rows of layer 0 are rows of the per-token table: no GEMM
    if (!from_table) {
      CK(ie::launch_gemm_bf16(g, s));
      h->launches++;
    }
    if ((rc = mark(h, s)) != IE_OK) return rc;

    // recurrence
    ie::LstmStepArgs a{};
    __nv_bfloat16* ybuf = h->y[cur].as<__nv_bfloat16>();
    CK(ie::make_tmap_bf16_2d(&a.tm_h, ybuf, L.kh_pad, static_cast<uint64_t>(rows + b_pad), h->y_ld, 64, 128));
    CK(ie::make_tmap_bf16_2d(&a.tm_w, L.w_hh.p, L.kh_pad, 4ull * L.out_pad, L.kh_pad, 64, 4 * L.u));
    a.cluster = L.cluster;
    a.fast_math = h->fast_math;
    CK(ie::make_tmap_bf16_2d(&a.tm_hs, ybuf, L.kh_pad, static_cast<uint64_t>(rows + b_pad), h->y_ld, 64,
                             128 / L.cluster));
    a.gx = from_table ? h->proj.as<float>() : h->gx.as<float>();
    a.c = h->c.as<float>();
    a.y = ybuf;
    a.raw = (last && (raw_out != nullptr || pool_raw)) ? h->raw.as<float>() : nullptr;
    a.pool_sum = (last && pooled && !pool_raw) ? h->pool_sum.as<float>() : nullptr;
    a.pool_max = h->pool_max.as<float>();
    a.pool_last = h->pool_last.as<float>();
    a.lengths = h->lengths.as<int>();
    a.T = T;
    a.b_pad = b_pad;
    a.u = L.u;
    a.n_cta = L.n_cta;
    a.out_pad = L.out_pad;
    a.kh_pad = L.kh_pad;
    a.ldy = h->y_ld;
    a.raw_ld = L.out_pad;
    if (rot) {
      ie::LstmWideArgs q{};
      q.tm_h = a.tm_h; q.tm_w = a.tm_w; q.gx = a.gx; q.c = a.c; q.y = a.y; q.raw = a.raw;
      q.pool_sum = a.pool_sum; q.pool_max = a.pool_max; q.pool_last = a.pool_last; q.lengths = a.lengths;
      q.step_done = h->step_done.as<unsigned>() + static_cast<size_t>(l) * T * kStepStride;
      q.T = T; q.ng = b_pad / 256; q.u = L.u; q.n_cta = L.n_cta; q.out_pad = L.out_pad; q.kh_pad = L.kh_pad;
      q.ldy = a.ldy; q.raw_ld = a.raw_ld; q.fast_math = h->fast_math; q.num_sms = h->num_sms; q.check_only = 0;
      q.trace = nullptr;
      q.variant = h->rot_variant;
      q.tok = from_table ? h->tok.as<int>() : nullptr;
      if (l == h->trace_layer) {
        const int pairs = ie::lstm_rot_pairs(q);
        const long long items = (static_cast<long long>(T) * q.ng * (L.n_cta / 2) + pairs - 1) / pairs;
        CK(h->trace.reserve(static_cast<size_t>(2 * pairs) * items * 12 * sizeof(long long), true));
        q.trace = h->trace.as<long long>();
        q.trace_items = static_cast<int>(items);
        h->trace_T = static_cast<int>(items);
        h->trace_ctas = 2 * pairs;
      }
      CK(ie::launch_lstm_rot(q, s));
      h->launches += 1;
    } else if (wide && !(last && last_on_seq)) {
      ie::LstmWideArgs q{};
      q.tm_h = a.tm_h; q.tm_w = a.tm_w; q.gx = a.gx; q.c = a.c; q.y = a.y; q.raw = a.raw;
      q.pool_sum = a.pool_sum; q.pool_max = a.pool_max; q.pool_last = a.pool_last; q.lengths = a.lengths;
      q.step_done = h->step_done.as<unsigned>() + static_cast<size_t>(l) * T * kStepStride;
      q.T = T; q.ng = 3; q.u = L.u; q.n_cta = L.n_cta; q.out_pad = L.out_pad; q.kh_pad = L.kh_pad;
      q.ldy = a.ldy; q.raw_ld = a.raw_ld; q.fast_math = h->fast_math; q.num_sms = h->num_sms; q.check_only = 0;
      q.trace = nullptr;
      q.tok = from_table ? h->tok.as<int>() : nullptr;
      if (l == h->trace_layer) {
        const int grid = 2 * std::min(h->num_sms / 2, 3 * (L.n_cta / 2));
        CK(h->trace.reserve(static_cast<size_t>(grid) * T * 12 * sizeof(long long), true));
        q.trace = h->trace.as<long long>();
        h->trace_T = T;
        h->trace_ctas = grid;
      }
      CK(ie::launch_lstm_wide(q, s));
      h->launches += 1;
    } else if (seq || (last && last_on_seq)) {
      ie::LstmSeqArgs q{};
      q.tm_h = a.tm_h; q.tm_w = a.tm_w; q.gx = a.gx; q.y = a.y; q.raw = a.raw;
      q.pool_sum = a.pool_sum; q.pool_max = a.pool_max; q.pool_last = a.pool_last; q.lengths = a.lengths;
      q.step_done = h->step_done.as<unsigned>() + static_cast<size_t>(l) * T * kStepStride;
      q.T = T; q.b_pad = b_pad; q.u = L.u; q.n_cta = L.n_cta; q.out_pad = L.out_pad; q.kh_pad = L.kh_pad;
      q.ldy = a.ldy; q.raw_ld = a.raw_ld; q.check_only = 0;
      q.fast_math = h->fast_math;
      q.trace = nullptr;
      if (l == h->trace_layer) {
        CK(h->trace.reserve(static_cast<size_t>(L.n_cta) * T * 12 * sizeof(long long), true));
        q.trace = h->trace.as<long long>();
        h->trace_T = T;
        h->trace_ctas = L.n_cta;
      }
      CK(ie::launch_lstm_seq(q, s));
      h->launches += 1;
    } else {
      for (int t = 0; t < T; ++t) {
        a.t = t;
        CK(ie::launch_lstm_step(a, s));
      }
      h->launches += T;
    }
    if ((rc = mark(h, s)) != IE_OK) return rc;
    layer_in = ybuf + static_cast<long long>(b_pad) * h->y_ld;  // slot 1 onwards
    layer_in_ld = h->y_ld;
    cur ^= 1;
  }

  const Layer& LL = (last_on_seq && !rot) ? h->layers.back() : LS.back();
  if (pooled) {
    float* out_dev = dev ? out : h->out.as<float>();
    if (pool_raw)
      CK(ie::launch_pool_from_raw(h->raw.as<float>(), h->lengths.as<int>(), B, T, c.emb_sz, LL.out_pad, out_dev, s));
    else
      CK(ie::launch_pool_finalize(h->pool_sum.as<float>(), h->pool_max.as<float>(), h->pool_last.as<float>(),
                                  h->lengths.as<int>(), B, c.emb_sz, LL.out_pad, out_dev, s));
    h->launches++;
    if ((rc = mark(h, s)) != IE_OK) return rc;
    if (!dev)
      CK(cudaMemcpyAsync(out, out_dev, static_cast<size_t>(B) * 3 * c.emb_sz * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  if (raw_out != nullptr) {
    // raw workspace is [b_pad, T, out_pad]; compact to [B, T, emb_sz]
    CK(cudaMemcpy2DAsync(raw_out, static_cast<size_t>(c.emb_sz) * sizeof(float), h->raw.p,
                         static_cast<size_t>(LL.out_pad) * sizeof(float), static_cast<size_t>(c.emb_sz) * sizeof(float),
                         static_cast<size_t>(B) * T, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  }
  if (!dev) {
    int err_host = 0;
    CK(cudaMemcpyAsync(&err_host, h->err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (err_host != 0) {
      cudaMemsetAsync(h->err.p, 0, sizeof(int), s);
      return fail(IE_ERR_TOKEN, "token id outside [0,%d) in ids", c.vocab_sz);
    }
  }
  return IE_OK;
}

}  // namespace

extern "C" {

int ie_version(void) { return 100; }

const char* ie_last_error(void) { return g_last_error.c_str(); }

int ie_encoder_create(const ie_config* cfg, ie_encoder** out) {
  if (cfg == nullptr || out == nullptr) return fail(IE_ERR_INVALID, "null argument");
  if (cfg->n_layers < 1 || cfg->n_layers > 16 || cfg->emb_sz < 1 || cfg->n_hid < 1 || cfg->vocab_sz < 1 ||
      cfg->pad_idx < 0 || cfg->pad_idx >= cfg->vocab_sz)
    return fail(IE_ERR_INVALID, "bad encoder config");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(IE_ERR_CUDA, "no CUDA device available (%s): this library has no CPU fallback",
                cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(IE_ERR_INVALID, "device %d not in [0,%d)", cfg->device, ndev);
  CK(cudaSetDevice(cfg->device));
  int major = 0, sms = 0;
  CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cfg->device));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
  if (major != 10) return fail(IE_ERR_CUDA, "device compute capability %d.x is not sm_100 (B200)", major);
  ie_encoder* h = new ie_encoder();
  h->cfg = *cfg;
  h->num_sms = sms;
  if (const char* e = getenv("IE_SEQ")) h->use_seq = atoi(e);
  if (const char* e = getenv("IE_ROT")) h->use_rot = atoi(e);
  if (const char* e = getenv("IE_ROT_VARIANT")) h->rot_variant = atoi(e) & 3;
  if (const char* e = getenv("IE_EMB_PROJ")) h->use_proj = atoi(e);
  if (const char* e = getenv("IE_POOL_RAW")) h->use_pool_raw = atoi(e);
  if (h->use_rot) {
    // five batches put an item's inputs two rounds back (C = 190 >= 2*74 + 38); IE_ROT_BATCHES=6..8 buys more slack
    int nb = 5;
    if (const char* e = getenv("IE_ROT_BATCHES")) nb = std::min(ie::kRotMaxBatches, std::max(1, atoi(e)));
    h->max_batch = std::max(IE_MAX_BATCH, 256 * nb);
  }
  h->fast_math = (cfg->flags & IE_CFG_ACCURATE_GATES) ? 0 : 1;
  if (const char* e = getenv("IE_FAST_MATH")) h->fast_math = atoi(e);
  int rc = plan_layers(h, h->layers, 0);
  if (rc == IE_OK) rc = plan_layers(h, h->layersB, 32);
  if (rc != IE_OK) { delete h; return rc; }
  e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaStreamCreate"); }
  *out = h;
  return IE_OK;
}

void ie_encoder_destroy(ie_encoder* h) {
  if (h == nullptr) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (Layer& L : h->layers) { L.w_ih.release(); L.w_hh.release(); L.bias.release(); }
  for (Layer& L : h->layersB) { L.w_ih.release(); L.w_hh.release(); L.bias.release(); }
  DevBuf* bufs[] = {&h->emb, &h->ids, &h->lengths, &h->x0, &h->y[0], &h->y[1], &h->gx, &h->c, &h->pool_sum,
                    &h->pool_max, &h->pool_last, &h->out, &h->raw, &h->err, &h->step_done, &h->trace, &h->proj, &h->tok};
  for (DevBuf* b : bufs) b->release();
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

int ie_encoder_load_embedding(ie_encoder* h, const float* emb) {
  if (h == nullptr || emb == nullptr) return fail(IE_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->cfg.device));
  // IE_EMB_PROJ: the table GEMM reads the embedding as its A operand, so its rows are padded to whole M tiles (zeros)
  const int v_rows = h->use_proj ? static_cast<int>(round_up(h->cfg.vocab_sz, 256)) : h->cfg.vocab_sz;
  std::vector<int> ident(v_rows);
  for (int i = 0; i < v_rows; ++i) ident[i] = i < h->cfg.vocab_sz ? i : -1;
  int rc = upload_sliced(emb, h->cfg.vocab_sz, h->cfg.emb_sz, ident, h->e_pad, h->emb, h->own_stream);
  if (rc != IE_OK) return rc;
  h->emb_loaded = true;
  h->proj_built = false;
  return IE_OK;
}

int ie_encoder_load_layer(ie_encoder* h, int32_t layer, const float* w_ih, const float* w_hh, const float* b_ih,
                          const float* b_hh) {
  if (h == nullptr || w_ih == nullptr || w_hh == nullptr || b_ih == nullptr || b_hh == nullptr)
    return fail(IE_ERR_INVALID, "null argument");
  if (layer < 0 || layer >= h->cfg.n_layers) return fail(IE_ERR_INVALID, "layer %d out of range", layer);
  std::lock_guard<std::mutex> lk(h->mu);
  CK(cudaSetDevice(h->cfg.device));
  for (std::vector<Layer>* plan : {&h->layers, &h->layersB}) {  // both weight layouts stay resident (2 x 266 MB at R4)
    Layer& L = (*plan)[layer];
    const std::vector<int> perm = slice_perm(L);
    int rc = upload_sliced(w_ih, 4 * L.out, L.in, perm, L.kin_pad, L.w_ih, h->own_stream);
    if (rc != IE_OK) return rc;
    rc = upload_sliced(w_hh, 4 * L.out, L.out, perm, L.kh_pad, L.w_hh, h->own_stream);
    if (rc != IE_OK) return rc;
    std::vector<float> bias(perm.size());
    for (size_t r = 0; r < perm.size(); ++r) bias[r] = perm[r] < 0 ? 0.0f : b_ih[perm[r]] + b_hh[perm[r]];
    CK(L.bias.reserve(bias.size() * sizeof(float)));
    CK(cudaMemcpy(L.bias.p, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    L.loaded = true;
  }
  if (layer == 0) h->proj_built = false;
  return IE_OK;
}

// rows beyond what the available persistent kernels take in one launch are run as consecutive sub-batches
static int encode_locked(ie_encoder* h, const int64_t* ids, const int32_t* lengths, int B, int T, float* out, int flags,
                         cudaStream_t s) {
  const long long ow = 3ll * h->cfg.emb_sz;
  if (B > IE_MAX_BATCH) {  // only reachable with the experimental rotating schedule (max_batch > IE_MAX_BATCH)
    if (h->use_rot) {
      const int rc = run_encoder(h, ids, lengths, B, T, out, nullptr, flags, s);
      if (rc != IE_ERR_STATE || h->use_rot) return rc;
    }
    const int rc = encode_locked(h, ids, lengths, IE_MAX_BATCH, T, out, flags, s);
    if (rc != IE_OK) return rc;
    return encode_locked(h, ids + static_cast<long long>(IE_MAX_BATCH) * T, lengths + IE_MAX_BATCH, B - IE_MAX_BATCH, T,
                         out + IE_MAX_BATCH * ow, flags, s);
  }
  if (B > 512) {
    if (h->use_wide) {
      const int rc = run_encoder(h, ids, lengths, B, T, out, nullptr, flags, s);
      if (rc != IE_ERR_STATE || h->use_wide) return rc;
    }
    const int rc = encode_locked(h, ids, lengths, 512, T, out, flags, s);
    if (rc != IE_OK) return rc;
    return encode_locked(h, ids + 512ll * T, lengths + 512, B - 512, T, out + 512 * ow, flags, s);
  }
  if (B > 256) {
    if (h->use_seq) {
      const int rc = run_encoder(h, ids, lengths, B, T, out, nullptr, flags, s);
      if (rc != IE_ERR_STATE || h->use_seq) return rc;
    }
    const int rc = encode_locked(h, ids, lengths, 256, T, out, flags, s);
    if (rc != IE_OK) return rc;
    return encode_locked(h, ids + 256ll * T, lengths + 256, B - 256, T, out + 256 * ow, flags, s);
  }
  return run_encoder(h, ids, lengths, B, T, out, nullptr, flags, s);
}

int ie_encoder_encode(ie_encoder* h, const int64_t* ids, const int32_t* lengths, int32_t B, int32_t T, float* out,
                      int32_t flags, void* stream) {
  if (h == nullptr) return fail(IE_ERR_INVALID, "null handle");
  if (out == nullptr) return fail(IE_ERR_INVALID, "out is null");
  if (ids == nullptr || lengths == nullptr) return fail(IE_ERR_INVALID, "null pointer");
  if (B < 1 || B > h->max_batch) return fail(IE_ERR_INVALID, "B=%d outside [1,%d]", B, h->max_batch);
  std::lock_guard<std::mutex> lk(h->mu);
  // device-pointer mode: `stream` is used verbatim (NULL = the legacy default stream, e.g. torch's default);
  // host-pointer mode: NULL selects the handle's own stream
  cudaStream_t s = (stream || (flags & IE_FLAG_DEVICE_PTRS)) ? static_cast<cudaStream_t>(stream) : h->own_stream;
  return encode_locked(h, ids, lengths, B, T, out, flags, s);
}

int ie_encoder_raw_features(ie_encoder* h, const int64_t* ids, int32_t B, int32_t T, float* raw, int32_t flags,
                            void* stream) {
  if (h == nullptr) return fail(IE_ERR_INVALID, "null handle");
  if (raw == nullptr) return fail(IE_ERR_INVALID, "raw is null");
  std::lock_guard<std::mutex> lk(h->mu);
  cudaStream_t s = (stream || (flags & IE_FLAG_DEVICE_PTRS)) ? static_cast<cudaStream_t>(stream) : h->own_stream;
  return run_encoder(h, ids, nullptr, B, T, nullptr, raw, flags, s);
}

int64_t ie_encoder_launch_count(const ie_encoder* h) { return h ? h->launches : 0; }

int32_t ie_encoder_max_batch(const ie_encoder* h) { return h ? h->max_batch : IE_MAX_BATCH; }

// debug: request a per-step timeline of `layer` in the persistent kernel on the next encode (layer < 0: off);
// with out != NULL copy the last recorded timeline [n_cta][T][8] (SM clocks) and return n_cta*T
int64_t ie_debug_seq_trace(ie_encoder* h, int32_t layer, long long* out, int64_t cap) {
  if (h == nullptr) return fail(IE_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  h->trace_layer = layer;
  if (out == nullptr) return 0;
  const int64_t n = static_cast<int64_t>(h->trace_ctas) * h->trace_T;
  if (n == 0 || n * 12 > cap) return fail(IE_ERR_STATE, "no trace recorded or buffer too small");
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  if (cudaMemcpy(out, h->trace.p, n * 12 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail(IE_ERR_CUDA, "trace copy failed");
  return n;
}

int ie_encoder_last_phase_ms(ie_encoder* h, float* ms, int32_t cap) {
  if (h == nullptr || ms == nullptr) return fail(IE_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->ev_used < 2) return fail(IE_ERR_STATE, "no encode call recorded");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaEventSynchronize(h->ev[h->ev_used - 1]));
  const int n = h->ev_used - 1;
  for (int i = 0; i < n && i < cap; ++i) CK(cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
  return n;
}

// debug: cycles to issue / execute iters*4 tcgen05.mma of shape M=128 (mode 0) or M=256 CTA pair (mode 1) x N=n x K=16
int ie_debug_umma_rate(int32_t mode, int32_t n, int32_t iters, int32_t commit_every, int32_t grid, int32_t ntiles,
                       long long* out2) {
  if (out2 == nullptr || n < 16 || n > 256 || n % 16 || iters < 1) return fail(IE_ERR_INVALID, "bad argument");
  CK(ie::run_umma_rate(mode, n, iters, commit_every, grid < 1 ? 1 : grid, ntiles, out2));
  return IE_OK;
}

// ---------------------------------------------------------------------------------------------
// MLP head
// ---------------------------------------------------------------------------------------------
int ie_mlp_create(int32_t n_layers, const int32_t* dims, int32_t device, ie_mlp** out) {
  if (dims == nullptr || out == nullptr || n_layers < 1 || n_layers > 16) return fail(IE_ERR_INVALID, "bad argument");
  for (int i = 0; i <= n_layers; ++i)
    if (dims[i] < 1) return fail(IE_ERR_INVALID, "dims[%d]=%d", i, dims[i]);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(IE_ERR_CUDA, "no CUDA device available (%s): this library has no CPU fallback",
                cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(IE_ERR_INVALID, "device %d not in [0,%d)", device, ndev);
  CK(cudaSetDevice(device));
  int major = 0, sms = 0;
  CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (major != 10) return fail(IE_ERR_CUDA, "device compute capability %d.x is not sm_100 (B200)", major);
  ie_mlp* m = new ie_mlp();
  m->device = device;
  m->num_sms = sms;
  m->dims.assign(dims, dims + n_layers + 1);
  m->layers.resize(n_layers);
  long long ld = 64;
  for (int l = 0; l < n_layers; ++l) {
    ie_mlp::L& L = m->layers[l];
    L.k_pad = static_cast<int>(round_up(dims[l], 64));
    const int n16 = static_cast<int>(round_up(dims[l + 1], 16));
    L.bn = n16 >= 128 ? 128 : n16;
    L.n_pad = static_cast<int>(round_up(dims[l + 1], L.bn));
    ld = std::max<long long>(ld, round_up(std::max(L.n_pad, L.k_pad), 64));
  }
  m->act_ld = ld;
  e = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete m; return cuda_fail(e, "cudaStreamCreate"); }
  *out = m;
  return IE_OK;
}

int ie_mlp_load_layer(ie_mlp* m, int32_t layer, const float* coef, const float* intercept) {
  if (m == nullptr || coef == nullptr || intercept == nullptr) return fail(IE_ERR_INVALID, "null argument");
  if (layer < 0 || layer >= static_cast<int>(m->layers.size())) return fail(IE_ERR_INVALID, "layer out of range");
  std::lock_guard<std::mutex> lk(m->mu);
  CK(cudaSetDevice(m->device));
  ie_mlp::L& L = m->layers[layer];
  const int fan_in = m->dims[layer], fan_out = m->dims[layer + 1];
  // sklearn coefs_[l] is [fan_in, fan_out]; the GEMM wants B = [fan_out rows, fan_in] (K-major)
  std::vector<float> wt(static_cast<size_t>(fan_out) * fan_in);
  for (int i = 0; i < fan_in; ++i)
    for (int o = 0; o < fan_out; ++o) wt[static_cast<size_t>(o) * fan_in + i] = coef[static_cast<size_t>(i) * fan_out + o];
  std::vector<int> perm(L.n_pad);
  for (int r = 0; r < L.n_pad; ++r) perm[r] = r < fan_out ? r : -1;
  int rc = upload_sliced(wt.data(), fan_out, fan_in, perm, L.k_pad, L.w, m->own_stream);
  if (rc != IE_OK) return rc;
  std::vector<float> b(L.n_pad, 0.0f);
  std::copy(intercept, intercept + fan_out, b.begin());
  CK(L.b.reserve(b.size() * sizeof(float)));
  CK(cudaMemcpy(L.b.p, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
  L.loaded = true;
  return IE_OK;
}

int ie_mlp_predict_proba(ie_mlp* m, const float* X, int32_t n, float* probs, int32_t flags, void* stream) {
  if (m == nullptr || X == nullptr || probs == nullptr) return fail(IE_ERR_INVALID, "null argument");
  if (n < 1) return fail(IE_ERR_INVALID, "n=%d", n);
  for (const auto& L : m->layers)
    if (!L.loaded) return fail(IE_ERR_STATE, "MLP layer weights not loaded");
  std::lock_guard<std::mutex> lk(m->mu);
  CK(cudaSetDevice(m->device));
  const bool dev = (flags & IE_FLAG_DEVICE_PTRS) != 0;
  cudaStream_t s = (stream || dev) ? static_cast<cudaStream_t>(stream) : m->own_stream;
  const int nl = static_cast<int>(m->layers.size());
  const int d_in = m->dims[0], n_labels = m->dims[nl];
  const int chunk = 1 << 16;
  const ie_mlp::L& LL = m->layers[nl - 1];
  CK(m->act[0].reserve(static_cast<size_t>(chunk) * m->act_ld * sizeof(__nv_bfloat16), true));
  CK(m->act[1].reserve(static_cast<size_t>(chunk) * m->act_ld * sizeof(__nv_bfloat16), true));
  CK(m->probs.reserve(static_cast<size_t>(chunk) * LL.n_pad * sizeof(float)));
  if (!dev) CK(m->xf.reserve(static_cast<size_t>(chunk) * d_in * sizeof(float)));
  for (long long r0 = 0; r0 < n; r0 += chunk) {
    const int rows = static_cast<int>(std::min<long long>(chunk, n - r0));
    const int m_pad = static_cast<int>(round_up(rows, 128));
    const float* xsrc = X + r0 * d_in;
    if (!dev) {
      CK(cudaMemcpyAsync(m->xf.p, xsrc, static_cast<size_t>(rows) * d_in * sizeof(float), cudaMemcpyHostToDevice, s));
      xsrc = m->xf.as<float>();
    }
    // f32 -> bf16, K padded with zeros (rows beyond `rows` keep stale finite data; they are never stored)
    CK(ie::launch_convert_rows(xsrc, d_in, d_in, nullptr, rows, m->act[0].as<__nv_bfloat16>(), m->act_ld, s));
    int cur = 0;
    for (int l = 0; l < nl; ++l) {
      const ie_mlp::L& L = m->layers[l];
      const bool last = (l == nl - 1);
      ie::GemmArgs g{};
      g.a = m->act[cur].as<__nv_bfloat16>();
      g.lda = m->act_ld;
      g.b = L.w.as<__nv_bfloat16>();
      g.ldb = L.k_pad;
      g.bias = L.b.as<float>();
      g.m_pad = m_pad;
      g.n_pad = L.n_pad;
      g.k_pad = L.k_pad;
      g.m_store = rows;
      g.n_store = L.n_pad;
      g.bn = L.bn;
      g.num_sms = m->num_sms;
      if (last) {
        g.d = m->probs.p;
        g.ldd = L.n_pad;
        g.act = 2;
        g.out_bf16 = 0;
      } else {
        g.d = m->act[cur ^ 1].p;
        g.ldd = m->act_ld;
        g.act = 1;
        g.out_bf16 = 1;
      }
      CK(ie::launch_gemm_bf16(g, s));
      cur ^= 1;
    }
    CK(cudaMemcpy2DAsync(probs + r0 * n_labels, static_cast<size_t>(n_labels) * sizeof(float), m->probs.p,
                         static_cast<size_t>(LL.n_pad) * sizeof(float), static_cast<size_t>(n_labels) * sizeof(float),
                         rows, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    if (!dev) CK(cudaStreamSynchronize(s));  // xf is reused by the next chunk
  }
  return IE_OK;
}

void ie_mlp_destroy(ie_mlp* m) {
  if (m == nullptr) return;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  for (auto& L : m->layers) { L.w.release(); L.b.release(); }
  m->xf.release(); m->act[0].release(); m->act[1].release(); m->probs.release();
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  delete m;
}

int ie_debug_gemm(const float* a, const float* b, const float* bias, int32_t M, int32_t N, int32_t K, int32_t act,
                  float* d, int32_t device) {
  if (a == nullptr || b == nullptr || d == nullptr || M < 1 || N < 1 || K < 1) return fail(IE_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(device));
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int m_pad = static_cast<int>(round_up(M, 128)), k_pad = static_cast<int>(round_up(K, 64));
  const int n16 = static_cast<int>(round_up(N, 16));
  int bn = 16;
  if (n16 >= 240 && n16 % 240 == 0) bn = 240;
  else if (n16 >= 128) bn = 128;
  else bn = n16;
  const int n_pad = static_cast<int>(round_up(N, bn));
  DevBuf fa, fb, ba, bb, dd, db;
  cudaStream_t s = nullptr;
  CK(fa.reserve(static_cast<size_t>(M) * K * 4));
  CK(fb.reserve(static_cast<size_t>(N) * K * 4));
  CK(ba.reserve(static_cast<size_t>(m_pad) * k_pad * 2, true));
  CK(bb.reserve(static_cast<size_t>(n_pad) * k_pad * 2, true));
  CK(dd.reserve(static_cast<size_t>(m_pad) * n_pad * 4));
  CK(cudaMemcpy(fa.p, a, static_cast<size_t>(M) * K * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(fb.p, b, static_cast<size_t>(N) * K * 4, cudaMemcpyHostToDevice));
  CK(ie::launch_convert_rows(fa.as<float>(), K, K, nullptr, M, ba.as<__nv_bfloat16>(), k_pad, s));
  CK(ie::launch_convert_rows(fb.as<float>(), K, K, nullptr, N, bb.as<__nv_bfloat16>(), k_pad, s));
  if (bias) {
    std::vector<float> bp(n_pad, 0.0f);
    std::copy(bias, bias + N, bp.begin());
    CK(db.reserve(n_pad * 4));
    CK(cudaMemcpy(db.p, bp.data(), n_pad * 4, cudaMemcpyHostToDevice));
  }
  ie::GemmArgs g{};
  g.a = ba.as<__nv_bfloat16>(); g.lda = k_pad;
  g.b = bb.as<__nv_bfloat16>(); g.ldb = k_pad;
  g.d = dd.p; g.ldd = n_pad;
  g.bias = bias ? db.as<float>() : nullptr;
  g.m_pad = m_pad; g.n_pad = n_pad; g.k_pad = k_pad;
  g.m_store = M; g.n_store = n_pad; g.bn = bn; g.act = act; g.out_bf16 = 0; g.num_sms = sms;
  CK(ie::launch_gemm_bf16(g, s));
  CK(cudaMemcpy2D(d, static_cast<size_t>(N) * 4, dd.p, static_cast<size_t>(n_pad) * 4, static_cast<size_t>(N) * 4, M,
                  cudaMemcpyDeviceToHost));
  fa.release(); fb.release(); ba.release(); bb.release(); dd.release(); db.release();
  return IE_OK;
}

}  // extern "C"
