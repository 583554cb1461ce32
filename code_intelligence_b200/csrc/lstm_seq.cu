// Persistent recurrent kernel: ALL T timesteps of one LSTM layer in one launch, for one or two independent batches of 256
// rows (the same arithmetic as lstm.cu, which stays as the fallback; reference call sites:
// Issue_Embeddings/flask_app/inference.py:56-57, :66-68, pooling :239).
//
// Why (profiles/README.md, round 1): with one launch per timestep the recurrence ran at 23.7 us per step: the
// cta_group::1 M=128,N=80 UMMAs use a third of the tensor pipe (an M=128 MMA costs ~128-170 cycles whatever N is) and
// ~9 us of launch / prologue / epilogue per step cannot overlap because step t+1 needs every CTA's h_t.  This kernel
//   * pairs CTAs (cluster of 2 on one TPC) into one M=256 tensor core (tcgen05 cta_group::2): CTA r of a pair
//     holds batch rows [128r, 128r+128) and HALF of the pair's W_hh slice, so per step an SM ingests
//     128 x K of h plus 2u x K of W instead of 256 x K plus 4u x K  (1.0 MB instead of 1.63 MB at H=2400);
//   * keeps c_t in registers, barriers / TMEM / tensor maps alive across steps, and prefetches the next step's W_hh
//     k-blocks (which do not depend on h_t) through their own ring while the epilogue and the step barrier run;
//   * replaces the kernel boundary by a grid-scope counter per (step, batch): epilogue threads store h_t, CTA barrier,
//     one thread fences and does red.add; the h producer of every CTA spins (bounded) with relaxed loads, fences,
//     fence.proxy.async, then issues the TMA loads of h_t;
//   * NG = 2: two independent 256-row batches alternate inside the launch (own TMEM accumulator, c registers and
//     counters), so the tensor pipe works on one batch while the other is in its epilogue / barrier phase.
// All CTAs must be co-resident (grid <= SMs, 1 CTA/SM; checked with cudaOccupancyMaxActiveClusters by the caller).
//
// Warp roles per CTA (384 threads): 0 = h (A) producer, 1 = UMMA issuer (leader CTA only), 2 = TMEM allocator,
// 3 = W producer, 4..11 = epilogue (two warps per TMEM lane quarter, each thread: one batch row x NCH chunks of 4 units).
#include "kernels.h"
#include "ptx.cuh"

namespace ie {

namespace {

constexpr int kSeqThreads = 384;
// The single UMMA-issuing thread is the bottleneck of the stream (measured: ~90 cycles per mbarrier wait, ~50 per commit,
// ~45 per MMA issue against 80 cycles of tensor time per M=256,N=160,K=16 MMA), so barriers guard GROUPS of 64-wide
// k-blocks: one wait / commit per kGA (h) or kGW (W_hh) k-blocks.
constexpr int kGA = 2;       // k-blocks per h stage (32 KB)
constexpr int kAStages = 4;  // 4 x 32 KB ring of h tiles
constexpr int kGW = 4;       // k-blocks per W_hh stage (ring mode)

template <int NCH, int NG, bool POOL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSeqThreads, 1)
lstm_seq_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w,
                const float* __restrict__ gx, __nv_bfloat16* __restrict__ y, float* __restrict__ raw,
                float* __restrict__ pool_sum, float* __restrict__ pool_max, float* __restrict__ pool_last,
                const int* __restrict__ lengths, unsigned* __restrict__ step_done, int T, int out_pad, int num_k_blocks,
                long long ldy, long long raw_ld, int w_stages, int w_resident, int tmem_cols, int fast_math,
                long long* __restrict__ trace) {
  // NG independent 256-row batches ride the same launch (ping-pong): while the epilogue / step barrier of one batch
  // runs, the tensor pipe works on the other one.  A time slot holds NG*256 rows.
  constexpr int kBPad = 256 * NG;
  constexpr int NH = NCH * 16;  // W rows this CTA contributes = accumulator columns of one slice
  constexpr int N = 2 * NH;     // accumulator columns of the pair
  extern __shared__ uint8_t smem_raw[];
  const uint32_t rawaddr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (rawaddr & 1023u)) & 1023u);

  constexpr uint32_t a_bytes = 128 * 64 * 2;
  constexpr uint32_t w_bytes = NH * 64 * 2;
  uint8_t* a_ring = smem;
  uint8_t* w_ring = smem + kAStages * kGA * a_bytes;
  // ring mode: w_stages stages of kGW k-blocks; resident mode: w_stages == num_k_blocks single k-block slots
  const uint32_t w_stage_bytes = w_resident ? w_bytes : kGW * w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_ring + static_cast<size_t>(w_stages) * w_stage_bytes);
  uint64_t* afull = bars;                     // [kAStages]  leader's copy is the live one
  uint64_t* aempty = afull + kAStages;        // [kAStages]
  uint64_t* wfull = aempty + kAStages;        // [w_stages]
  uint64_t* wempty = wfull + w_stages;        // [w_stages]
  uint64_t* tfull = wempty + w_stages;        // [NG]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + NG);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // optional per-CTA timeline (%globaltimer, ns) of every step: [cta][t][12]; null in production
#define IE_TRACE(slot, tt) do { if (trace) { unsigned long long _g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_g)); \
    trace[(static_cast<long long>(blockIdx.x) * T + (tt)) * 12 + (slot)] = static_cast<long long>(_g); } } while (0)
#define IE_TRACE_VAL(slot, tt, v) do { if (trace) trace[(static_cast<long long>(blockIdx.x) * T + (tt)) * 12 + (slot)] = (v); } while (0)
  const uint32_t crank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const unsigned total_ctas = gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_h);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(&afull[s], 2);   // leader's arrive.expect_tx + the peer's remote arrive
      mbar_init(&aempty[s], 1);  // the leader's multicast commit
    }
    for (int s = 0; s < w_stages; ++s) {
      mbar_init(&wfull[s], 2);
      mbar_init(&wempty[s], 1);
    }
    for (int g = 0; g < NG; ++g) mbar_init(&tfull[g], 1);
    fence_barrier_init();
  }
  cluster_sync();  // barrier inits visible to the peer before any remote arrive; both CTAs reach the 2-CTA alloc
  if (warp == 2) tmem_alloc_pair(tmem_slot, tmem_cols);
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- h producer: this CTA's 128 batch rows of h_{t-1}, k-block by k-block ----------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tg = 0; tg < T * NG; ++tg) {
        const int t = tg / NG, g = tg % NG;
        if (t > 0) {
          wait_flag_ge_relaxed(step_done + (tg - NG), total_ctas);  // every CTA has published h_{t-1} of batch g
          fence_proxy_async();                             // order the async-proxy (TMA) reads after the acquire
        }
        if (g == 0) IE_TRACE(0, t);
        const int row0 = t * kBPad + g * 256 + static_cast<int>(crank) * 128;
        for (int kb0 = 0; kb0 < num_k_blocks; kb0 += kGA) {
          const int n = min(kGA, num_k_blocks - kb0);
          mbar_wait(&aempty[stage], phase ^ 1);
          if (crank == 0) mbar_arrive_expect_tx(&afull[stage], 2 * n * a_bytes);
          else mbar_arrive_remote(&afull[stage], 0);
          for (int j = 0; j < n; ++j)
            tma_load_2d_pair(a_ring + (stage * kGA + j) * a_bytes, &tm_h, &afull[stage], (kb0 + j) * 64, row0,
                             kEvictNormal);
          if (++stage == kAStages) { stage = 0; phase ^= 1; }
        }
        if (g == 0) IE_TRACE(1, t);
      }
    }
  } else if (warp == 3) {
    // ---------------- W producer: this CTA's half of the pair's W_hh slice; free-running ahead of h -------------
    if (lane == 0) {
      const int wrow0 = blockIdx.x * NH;  // slices are laid out [cta][unit][gate]; CTA b owns rows [b*NH, +NH)
      if (w_resident) {
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          if (crank == 0) mbar_arrive_expect_tx(&wfull[kb], 2 * w_bytes);
          else mbar_arrive_remote(&wfull[kb], 0);
          tma_load_2d_pair(w_ring + kb * w_bytes, &tm_w, &wfull[kb], kb * 64, wrow0, kEvictLast);
        }
      } else {
        int stage = 0;
        uint32_t phase = 0;
        for (int tg = 0; tg < T * NG; ++tg) {
          for (int kb0 = 0; kb0 < num_k_blocks; kb0 += kGW) {
            const int n = min(kGW, num_k_blocks - kb0);
            mbar_wait(&wempty[stage], phase ^ 1);
            if (crank == 0) mbar_arrive_expect_tx(&wfull[stage], 2 * n * w_bytes);
            else mbar_arrive_remote(&wfull[stage], 0);
            for (int j = 0; j < n; ++j)
              tma_load_2d_pair(w_ring + (stage * kGW + j) * w_bytes, &tm_w, &wfull[stage], (kb0 + j) * 64, wrow0,
                               kEvictLast);
            if (++stage == w_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- UMMA issuer (leader CTA): M = 256 (both CTAs' rows), N = 2*NH, K = 16 per instruction -----
    if (crank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, N);
      const uint32_t a_base = smem_u32(a_ring);
      const uint32_t w_base = smem_u32(w_ring);
      int as = 0, ws = 0;
      uint32_t aph = 0, wph = 0;
      for (int tg = 0; tg < T * NG; ++tg) {
        const int t = tg / NG, g = tg % NG;
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(g * N);
        long long wa = 0, ww = 0, t_first = 0;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          const int ja = kb % kGA;   // position inside the current h stage
          const int jw = kb % kGW;   // position inside the current W stage (ring mode)
          if (ja == 0) {
            const long long c0 = trace ? clock64() : 0;
            mbar_wait(&afull[as], aph);
            if (kb == 0) { IE_TRACE(2, t); t_first = trace ? clock64() : 0; }
            else if (trace) wa += clock64() - c0;
          }
          if (w_resident) {
            if (tg == 0) mbar_wait(&wfull[kb], 0);
          } else if (jw == 0) {
            const long long c0 = trace ? clock64() : 0;
            mbar_wait(&wfull[ws], wph);
            if (trace) ww += clock64() - c0;
          }
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(a_base + (as * kGA + ja) * a_bytes);
          const uint64_t db = umma_desc_sw128(w_base + (w_resident ? kb : ws * kGW + jw) * w_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          const bool last = (kb == num_k_blocks - 1);
          if (ja == kGA - 1 || last) {
            umma_commit_pair_mc(&aempty[as], 0x3);
            if (++as == kAStages) { as = 0; aph ^= 1; }
          }
          if (!w_resident && (jw == kGW - 1 || last)) {
            umma_commit_pair_mc(&wempty[ws], 0x3);
            if (++ws == w_stages) { ws = 0; wph ^= 1; }
          }
        }
        umma_commit_pair_mc(&tfull[g], 0x3);
        if (g == 0) IE_TRACE(3, t);
        if (g == 0) IE_TRACE_VAL(8, t, wa);                                   // SM cycles waiting for h stages (after the first)
        if (g == 0) IE_TRACE_VAL(9, t, ww);                                   // SM cycles waiting for W stages
        if (g == 0) IE_TRACE_VAL(10, t, trace ? clock64() - t_first : 0);     // SM cycles first h stage -> all issued
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: gates, state update, h_t, pooling; then publish the step ----------------------
    const int e = warp - 4;
    const int q = e & 3;                      // TMEM lane quarter == warp % 4
    const int half = e >> 2;                  // which NH columns of the pair's N this warp handles
    const int row = static_cast<int>(crank) * 128 + q * 32 + lane;   // row inside a 256-row batch
    const int unit0 = pair * (2 * NCH * 4) + half * (NCH * 4);       // first hidden unit of this thread
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(half * NH);
    int len[NG];
    float cst[NG][NCH * 4];
    // last layer (POOL): the masked concat-pool accumulators of inference.py:239 live in registers for all T steps and
    // are written once at the end (a global read-modify-write per step put ~2.5 us on the step's critical path)
    constexpr int PN = POOL ? NCH * 4 : 1;
    float psum[NG][PN], pmax[NG][PN], plast[NG][PN];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      len[g] = POOL ? lengths[g * 256 + row] : 1;
#pragma unroll
      for (int i = 0; i < NCH * 4; ++i) cst[g][i] = 0.0f;
#pragma unroll
      for (int i = 0; i < PN; ++i) { psum[g][i] = 0.0f; pmax[g][i] = -INFINITY; plast[g][i] = 0.0f; }
    }

    for (int t = 0; t < T; ++t) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int brow = g * 256 + row;  // row inside the time slot
        const float4* gxp = reinterpret_cast<const float4*>(gx + (static_cast<long long>(t) * kBPad + brow) * (4ll * out_pad) +
                                                            4ll * unit0);
        float4 gxr[NCH * 4];
#pragma unroll
        for (int i = 0; i < NCH * 4; ++i) gxr[i] = __ldg(gxp + i);
        if (threadIdx.x == 128 && g == 0) IE_TRACE(7, t);
        if (t + 1 < T) {
          // pull the next step's Gx rows from HBM into L2 while this step streams, so that the loads at the top of
          // step t+1 are short and do not queue in front of the step barrier traffic
          const char* nx = reinterpret_cast<const char*>(gxp) + static_cast<long long>(kBPad) * out_pad * 16ll;
#pragma unroll
          for (int i = 0; i < (NCH * 64 + 127) / 128 + 1; ++i) prefetch_l2(nx + i * 128);
        }

        mbar_wait(&tfull[g], static_cast<uint32_t>(t & 1));
        tc_fence_after();
        if (threadIdx.x == 128 && g == 0) IE_TRACE(4, t);
        __nv_bfloat16* yrow = y + (static_cast<long long>(t + 1) * kBPad + brow) * ldy + unit0;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          uint32_t r[16];
          __syncwarp();
          tmem_ld16(taddr + g * N + ch * 16, r);
          tmem_ld_wait();
          float hn[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 gq = gxr[ch * 4 + j];
            const float zi = __uint_as_float(r[4 * j + 0]) + gq.x;
            const float zf = __uint_as_float(r[4 * j + 1]) + gq.y;
            const float zg = __uint_as_float(r[4 * j + 2]) + gq.z;
            const float zo = __uint_as_float(r[4 * j + 3]) + gq.w;
            float cn;
            if (fast_math) {
              cn = sigmoid_fast(zf) * cst[g][ch * 4 + j] + sigmoid_fast(zi) * tanh_fast(zg);
              hn[j] = sigmoid_fast(zo) * tanh_fast(cn);
            } else {
              cn = sigmoid_acc(zf) * cst[g][ch * 4 + j] + sigmoid_acc(zi) * tanh_acc(zg);
              hn[j] = sigmoid_acc(zo) * tanh_acc(cn);
            }
            cst[g][ch * 4 + j] = cn;
          }
          *reinterpret_cast<uint2*>(yrow + ch * 4) = make_uint2(pack_bf16x2(hn[0], hn[1]), pack_bf16x2(hn[2], hn[3]));
          if (raw != nullptr) {
            float4* rp = reinterpret_cast<float4*>(raw + (static_cast<long long>(brow) * T + t) * raw_ld + unit0 + ch * 4);
            *rp = make_float4(hn[0], hn[1], hn[2], hn[3]);
          }
          if constexpr (POOL) {
            if (t < len[g]) {
              const bool is_last = (t == len[g] - 1);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                psum[g][ch * 4 + j] += hn[j];
                pmax[g][ch * 4 + j] = fmaxf(pmax[g][ch * 4 + j], hn[j]);
                if (is_last) plast[g][ch * 4 + j] = hn[j];
              }
            }
          }
        }
        // publish: TMEM reads are done (the batch's next MMAs may overwrite its accumulator) and h_t is visible
        if (threadIdx.x == 128 && g == 0) IE_TRACE(5, t);
        tc_fence_before();
        named_bar_sync(1, 256);
        if (threadIdx.x == 128) {
          __threadfence();  // cumulative: covers the h stores of all 256 epilogue threads ordered by the barrier
          red_relaxed_add(step_done + t * NG + g, 1u);
          if (g == 0) IE_TRACE(6, t);
        }
      }
    }
    if constexpr (POOL) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const long long po = static_cast<long long>(g * 256 + row) * out_pad + unit0;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          *reinterpret_cast<float4*>(pool_sum + po + ch * 4) =
              make_float4(psum[g][ch * 4], psum[g][ch * 4 + 1], psum[g][ch * 4 + 2], psum[g][ch * 4 + 3]);
          *reinterpret_cast<float4*>(pool_max + po + ch * 4) =
              make_float4(pmax[g][ch * 4], pmax[g][ch * 4 + 1], pmax[g][ch * 4 + 2], pmax[g][ch * 4 + 3]);
          *reinterpret_cast<float4*>(pool_last + po + ch * 4) =
              make_float4(plast[g][ch * 4], plast[g][ch * 4 + 1], plast[g][ch * 4 + 2], plast[g][ch * 4 + 3]);
        }
      }
    }
  }

  __syncwarp();  // re-converge the single-lane role loops before the (aligned) cluster barrier
  tc_fence_before();
  cluster_sync();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, tmem_cols);
  }
}

template <int NCH, int NG, bool POOL>
cudaError_t launch_seq_t(const LstmSeqArgs& a, cudaStream_t stream) {
  const int nh = NCH * 16;
  const size_t a_ring = static_cast<size_t>(kAStages) * kGA * 128 * 64 * 2;
  const size_t w_bytes = static_cast<size_t>(nh) * 64 * 2;
  const int nkb = a.kh_pad / 64;
  const size_t budget = 227 * 1024 - 1024 - a_ring - 2048;  // alignment slack + barriers
  int w_stages;
  int resident = 0;
  if (static_cast<size_t>(nkb) * w_bytes <= budget && nkb <= 64) {
    w_stages = nkb;  // the whole slice stays in shared memory: loaded once
    resident = 1;
  } else {
    w_stages = static_cast<int>(budget / (kGW * w_bytes));
    if (w_stages > 4) w_stages = 4;
    if (w_stages < 2) return cudaErrorInvalidValue;
  }
  const size_t smem = 1024 + a_ring + static_cast<size_t>(w_stages) * (resident ? w_bytes : kGW * w_bytes) +
                      (2 * kAStages + 2 * w_stages + 2) * 8 + 16;
  int tmem_cols = 32;
  while (tmem_cols < NG * 2 * nh) tmem_cols <<= 1;
  auto kfn = lstm_seq_kernel<NCH, NG, POOL>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.n_cta);
  cfg.blockDim = dim3(kSeqThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  if (a.check_only) {
    // all CTAs must be co-resident: the step barrier would deadlock otherwise
    int max_clusters = 0;
    cfg.attrs = nullptr;  // the cluster shape is the kernel's compile-time __cluster_dims__(2,1,1)
    cfg.numAttrs = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &cfg);
    if (e != cudaSuccess) return e;
    return max_clusters * 2 >= a.n_cta ? cudaSuccess : cudaErrorCooperativeLaunchTooLarge;
  }
  kfn<<<a.n_cta, kSeqThreads, smem, stream>>>(a.tm_h, a.tm_w, a.gx, a.y, a.raw, a.pool_sum, a.pool_max, a.pool_last,
                                               a.lengths, a.step_done, a.T, a.out_pad, nkb, a.ldy, a.raw_ld, w_stages,
                                               resident, tmem_cols, a.fast_math, a.trace);
  return cudaGetLastError();
}

}  // namespace

// a.check_only != 0: only verify that the whole grid can be co-resident (no launch)
cudaError_t launch_lstm_seq(const LstmSeqArgs& a, cudaStream_t stream) {
  if (a.u % 4 || a.u < 4 || (a.b_pad != 256 && a.b_pad != 512 && a.b_pad != 768) || a.kh_pad % 64 || a.n_cta % 2)
    return cudaErrorInvalidValue;
  if (a.b_pad == 768) {
    // three batches per launch: only the narrow (<= 12 units per CTA) pooled last layer is run this way -- wider layers
    // with three batches go through lstm_wide.cu
    if (a.pool_sum == nullptr) return cudaErrorInvalidValue;
    switch (a.u / 4) {
      case 1: return launch_seq_t<1, 3, true>(a, stream);
      case 2: return launch_seq_t<2, 3, true>(a, stream);
      case 3: return launch_seq_t<3, 3, true>(a, stream);
      default: return cudaErrorInvalidValue;
    }
  }
#define IE_SEQ_CASE(n)                                                                                   \
  case n:                                                                                                \
    if (a.pool_sum != nullptr)                                                                           \
      return a.b_pad == 512 ? launch_seq_t<n, 2, true>(a, stream) : launch_seq_t<n, 1, true>(a, stream); \
    return a.b_pad == 512 ? launch_seq_t<n, 2, false>(a, stream) : launch_seq_t<n, 1, false>(a, stream);
  switch (a.u / 4) {
    IE_SEQ_CASE(1) IE_SEQ_CASE(2) IE_SEQ_CASE(3) IE_SEQ_CASE(4) IE_SEQ_CASE(5) IE_SEQ_CASE(6) IE_SEQ_CASE(7) IE_SEQ_CASE(8)
    default: return cudaErrorInvalidValue;
  }
#undef IE_SEQ_CASE
}

}  // namespace ie
