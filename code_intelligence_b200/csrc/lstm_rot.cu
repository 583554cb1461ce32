// Persistent recurrent kernel, "rotating schedule" variant of lstm_wide.cu (EXPERIMENTAL, opt-in with IE_ROT=1; the
// default paths do not reach this file).  Same arithmetic as lstm_wide.cu / lstm_seq.cu / lstm.cu (reference call
// sites: Issue_Embeddings/flask_app/inference.py:56-57, :66-68, pooling :239); checked on a B200 to give the same bits
// as the default kernels for 300..1280 rows (tools/gpu_rot.py) and measured at 6.8-7.0 k issues/s against 6.4-6.7 k of
// lstm_wide.cu on the same boxes (profiles/README.md, "Rotating schedule").
//
// Why: lstm_wide.cu deals the ng * tiles (batch, tile) chains of a timestep statically over the P = 74 CTA pairs; with
// 3 x 38 = 114 chains 40 pairs run two chains and 34 run one, so a timestep costs two chain times (2 x 13.8 us of MMAs)
// for 114 / 148 = 77 % of the pairs' capacity, and every chain's epilogue + step-counter latency (~8 us) is exposed
// once per timestep on the pairs that hold two chains of the same phase.  Here the work ITEMS
//     n = t * C + g * tiles + j        (C = ng * tiles; timestep t, batch g, tile j)
// are dealt round-robin in that global order: pair p runs items p, p + P, p + 2P, ...  Item n needs h_{t-1} of batch g,
// i.e. items n - C - j ... n - C + (tiles - 1 - j), all of which have a smaller index; since every pair walks its items
// in increasing order the item with the globally smallest index can always run => no wait cycle for any P, C.  With
// C >= 2 P + tiles (five batches at H = 2400: 190 >= 148 + 38) the items an item waits for were issued two rounds
// earlier, so their epilogue and counter latency hide behind the round in between and every pair issues MMAs
// back to back: 190 items / 74 pairs x 13.8 us = 35.4 us per timestep for five batches (7.1 us per batch-step against
// 12 us measured with lstm_wide.cu).
//
// Differences from lstm_wide.cu that the rotation forces:
//  * the cell state and the pooling accumulators of a chain are written by one SM and read by another one timestep
//    later: they are accessed with ld/st.global.cg (L2 only; L1 is not coherent) and read only after this CTA has
//    seen the (t-1, g) step counter (`cready`, a monotonic shared-memory sequence number advanced by a watcher warp,
//    which also takes the counter load and its gpu-scope fence off the h producer's path);
//  * the TMEM accumulator slot of item k (k & 1) is reused by item k + 2, which no longer belongs to the same batch,
//    so the drain is signalled explicitly (`tempty`, one arrive per CTA of the pair on the leader's barrier).
#include <cmath>

#include "kernels.h"
#include "ptx.cuh"

namespace ie {

namespace {

constexpr int kRThreads = 640;  // 4 role warps + 16 epilogue warps (4 per TMEM lane quarter, 64 columns each)
constexpr int kRGA = 2;                 // h ring: AST (3, variant: 4) stages x 2 k-blocks x 16 KB
constexpr int kRGW = 2, kRWStages = 3;  // W ring: 3 stages x 2 k-blocks x 16 KB
constexpr int kRTileN = 256;            // accumulator columns per tile = 64 hidden units
constexpr int kRHalfRows = 128;         // W rows each CTA of the pair contributes

__device__ __forceinline__ void st_release_cta(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
// bounded, backed-off spin of one lane until the shared sequence number reaches `target`
__device__ __forceinline__ void wait_seq_ge(const uint32_t* p, uint32_t target) {
  if (ld_acquire_cta(p) >= target) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (ld_acquire_cta(p) < target) {
    __nanosleep(64);
    if (((++spins) & 0xFFu) == 0 && (clock64() - t0) > 4000000000ll) __trap();
  }
}

// TOK: Gx rows indexed by token id (per-token projection table).  Variant knobs for round-2 experiments (both ran in
// the r7 session with identical results): WFENCE = the watcher, not the TMA-issuing thread, executes fence.proxy.async;
// AST = stages of the h ring.  <*, false, 3> is the validated default.
template <bool TOK, bool WFENCE, int AST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kRThreads, 1)
lstm_rot_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w,
                const float* __restrict__ gx, float* cstate, __nv_bfloat16* __restrict__ y, float* __restrict__ raw,
                float* pool_sum, float* pool_max, float* pool_last, const int* __restrict__ lengths,
                unsigned* __restrict__ step_done, int T, int ng, int tiles, int out_pad, int num_k_blocks, long long ldy,
                long long raw_ld, int fast_math, long long* __restrict__ trace, int trace_items,
                const int* __restrict__ tok) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t rawaddr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (rawaddr & 1023u)) & 1023u);

  constexpr uint32_t a_bytes = 128 * 64 * 2;
  constexpr uint32_t w_bytes = kRHalfRows * 64 * 2;
  uint8_t* a_ring = smem;
  uint8_t* w_ring = smem + AST * kRGA * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_ring + kRWStages * kRGW * w_bytes);
  uint64_t* afull = bars;                   // [AST] leader's copy is live
  uint64_t* aempty = afull + AST;
  uint64_t* wfull = aempty + AST;     // [kRWStages]
  uint64_t* wempty = wfull + kRWStages;
  uint64_t* tfull = wempty + kRWStages;     // [2] accumulator slot holds a finished item (both CTAs' copies live)
  uint64_t* tempty = tfull + 2;             // [2] accumulator slot drained by both CTAs (leader's copy is live)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* cready = tmem_slot + 1;         // number of this CTA's items whose (t-1, g) counter the watcher has seen

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // optional timeline of the pair's first `trace_items` items: [cta][k][12] (%globaltimer ns; slots 8-10 SM cycles)
#define IE_TRACE(slot, kk) do { if (trace && (kk) < trace_items) { unsigned long long _g; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_g)); \
    trace[(static_cast<long long>(blockIdx.x) * trace_items + (kk)) * 12 + (slot)] = static_cast<long long>(_g); } } while (0)
#define IE_TRACE_VAL(slot, kk, v) do { if (trace && (kk) < trace_items) \
    trace[(static_cast<long long>(blockIdx.x) * trace_items + (kk)) * 12 + (slot)] = (v); } while (0)
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int P = static_cast<int>(gridDim.x >> 1);
  const int C = ng * tiles;
  const long long total = static_cast<long long>(T) * C;
  const int b_pad = 256 * ng;
  const unsigned batch_ctas = 2u * static_cast<unsigned>(tiles);  // CTAs that publish a (step, batch)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_h);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < AST; ++s) {
      mbar_init(&afull[s], 2);
      mbar_init(&aempty[s], 1);
    }
    for (int s = 0; s < kRWStages; ++s) {
      mbar_init(&wfull[s], 2);
      mbar_init(&wempty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 2);
    }
    *cready = 0;
    fence_barrier_init();
  }
  cluster_sync();
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- h producer ------------------------------------------------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int k = 0;
      for (long long n = pair; n < total; n += P, ++k) {
        const int t = static_cast<int>(n / C);
        const int g = static_cast<int>(n - static_cast<long long>(t) * C) / tiles;
        wait_seq_ge(cready, static_cast<uint32_t>(k + 1));  // the watcher (warp 2) has seen counter (t-1, g)
        if (!WFENCE && t > 0) fence_proxy_async();  // h_{t-1} was written through the generic proxy, TMA reads it
        IE_TRACE(0, k);
        const int row0 = t * b_pad + g * 256 + static_cast<int>(crank) * 128;
        for (int kb0 = 0; kb0 < num_k_blocks; kb0 += kRGA) {
          const int nb = min(kRGA, num_k_blocks - kb0);
          mbar_wait(&aempty[stage], phase ^ 1);
          if (crank == 0) mbar_arrive_expect_tx(&afull[stage], 2 * nb * a_bytes);
          else mbar_arrive_remote(&afull[stage], 0);
          for (int q = 0; q < nb; ++q)
            tma_load_2d_pair(a_ring + (stage * kRGA + q) * a_bytes, &tm_h, &afull[stage], (kb0 + q) * 64, row0,
                             kEvictNormal);
          if (++stage == AST) { stage = 0; phase ^= 1; }
        }
        IE_TRACE(1, k);
      }
    }
  } else if (warp == 2) {
    // ---------------- counter watcher: runs ahead of the h producer and the epilogue ----------------------------
    // The counter load and the gpu-scope fence after it cost ~1-2 us next to the TMA streams; done here they are off
    // the h producer's path.  (Measured, profiles/README.md: moving the proxy fence here as well and a 4-stage h ring
    // changed nothing -- the kernel runs at the board's power cap, removing bubbles lowers the SM clock instead.)
    if (lane == 0) {
      int k = 0;
      for (long long n = pair; n < total; n += P, ++k) {
        const int t = static_cast<int>(n / C);
        const int g = static_cast<int>(n - static_cast<long long>(t) * C) / tiles;
        if (t > 0) {
          wait_flag_ge_relaxed(step_done + (t - 1) * ng + g, batch_ctas);  // ends with a gpu-scope fence
          if (WFENCE) fence_proxy_async();
        }
        st_release_cta(cready, static_cast<uint32_t>(k + 1));
      }
    }
  } else if (warp == 3) {
    // ---------------- W producer: free-running ahead of h ----------------------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long n = pair; n < total; n += P) {
        const int j = static_cast<int>(n % C) % tiles;
        const int wrow0 = (2 * j + static_cast<int>(crank)) * kRHalfRows;  // slices are [cta][unit][gate], 128 rows each
        for (int kb0 = 0; kb0 < num_k_blocks; kb0 += kRGW) {
          const int nb = min(kRGW, num_k_blocks - kb0);
          mbar_wait(&wempty[stage], phase ^ 1);
          if (crank == 0) mbar_arrive_expect_tx(&wfull[stage], 2 * nb * w_bytes);
          else mbar_arrive_remote(&wfull[stage], 0);
          for (int q = 0; q < nb; ++q)
            tma_load_2d_pair(w_ring + (stage * kRGW + q) * w_bytes, &tm_w, &wfull[stage], (kb0 + q) * 64, wrow0,
                             kEvictLast);
          if (++stage == kRWStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- UMMA issuer (leader CTA) --------------------------------------------------------------
    if (crank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, kRTileN);
      const uint32_t a_base = smem_u32(a_ring);
      const uint32_t w_base = smem_u32(w_ring);
      int as = 0, ws = 0;
      uint32_t aph = 0, wph = 0;
      int k = 0;
      for (long long n = pair; n < total; n += P, ++k) {
        const int slot = k & 1;
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(slot * kRTileN);
        long long wa = 0, ww = 0, t_first = 0;
        if (k >= 2) {  // the slot's previous item (k - 2) must have been read out of TMEM by both CTAs
          const long long c0 = trace ? clock64() : 0;
          mbar_wait(&tempty[slot], static_cast<uint32_t>(((k >> 1) - 1) & 1));
          IE_TRACE_VAL(11, k, trace ? clock64() - c0 : 0);
        }
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          const int ja = kb % kRGA, jw = kb % kRGW;
          if (ja == 0) {
            const long long c0 = trace ? clock64() : 0;
            mbar_wait(&afull[as], aph);
            if (kb == 0) { IE_TRACE(2, k); t_first = trace ? clock64() : 0; }
            else if (trace) wa += clock64() - c0;
          }
          if (jw == 0) {
            const long long c0 = trace ? clock64() : 0;
            mbar_wait(&wfull[ws], wph);
            if (trace) ww += clock64() - c0;
          }
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(a_base + (as * kRGA + ja) * a_bytes);
          const uint64_t db = umma_desc_sw128(w_base + (ws * kRGW + jw) * w_bytes);
#pragma unroll
          for (int q = 0; q < 4; ++q) umma_bf16_pair(tmem_d, da + 2 * q, db + 2 * q, idesc, (kb | q) != 0);
          const bool last = (kb == num_k_blocks - 1);
          if (ja == kRGA - 1 || last) {
            umma_commit_pair_mc(&aempty[as], 0x3);
            if (++as == AST) { as = 0; aph ^= 1; }
          }
          if (jw == kRGW - 1 || last) {
            umma_commit_pair_mc(&wempty[ws], 0x3);
            if (++ws == kRWStages) { ws = 0; wph ^= 1; }
          }
        }
        umma_commit_pair_mc(&tfull[slot], 0x3);
        IE_TRACE(3, k);
        IE_TRACE_VAL(8, k, wa);
        IE_TRACE_VAL(9, k, ww);
        IE_TRACE_VAL(10, k, trace ? clock64() - t_first : 0);
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ------------------------------------------------------------------------------
    const int e = warp - 4;
    const int q = e & 3;
    const int cq = e >> 2;  // which 64 of the tile's 256 columns (4 chunks of 16 = 16 hidden units per thread)
    const int row = static_cast<int>(crank) * 128 + q * 32 + lane;
    const bool pooled = pool_sum != nullptr;
    int k = 0;
    for (long long n = pair; n < total; n += P, ++k) {
      const int t = static_cast<int>(n / C);
      const int c = static_cast<int>(n - static_cast<long long>(t) * C);
      const int g = c / tiles, j = c % tiles;
      const int slot = k & 1;
      const int brow = g * 256 + row;
      const int unit0 = j * 64 + cq * 16;
      const int len = pooled ? lengths[brow] : 1;
      const long long grow = TOK ? static_cast<long long>(__ldg(tok + static_cast<long long>(t) * b_pad + brow))
                                   : static_cast<long long>(t) * b_pad + brow;  // TOK: per-token projection table
      const float4* gxp = reinterpret_cast<const float4*>(gx + grow * (4ll * out_pad) +
                                                          4ll * unit0);
      float4* cp = reinterpret_cast<float4*>(cstate + static_cast<long long>(brow) * out_pad + unit0);
      __nv_bfloat16* yrow = y + (static_cast<long long>(t + 1) * b_pad + brow) * ldy + unit0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(slot * kRTileN + cq * 64);
      // all of this thread's Gx (4 chunks x 4 units x 4 gates) and c are loaded while the MMAs still run
      constexpr int kCh = 4;
      float4 gxr[kCh][4];
      float4 cr[kCh];
#pragma unroll
      for (int ch = 0; ch < kCh; ++ch) {
        ldg_stream8(reinterpret_cast<const float*>(gxp + ch * 4), gxr[ch][0], gxr[ch][1]);
        ldg_stream8(reinterpret_cast<const float*>(gxp + ch * 4 + 2), gxr[ch][2], gxr[ch][3]);
      }
      // c_{t-1} of this chain was written by another pair: read it (from L2) only after (t-1, g) has been seen here
      if (lane == 0) wait_seq_ge(cready, static_cast<uint32_t>(k + 1));
      __syncwarp();
#pragma unroll
      for (int ch = 0; ch < kCh; ++ch) cr[ch] = (t == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcg(cp + ch);
      if (threadIdx.x == 128) IE_TRACE(7, k);
      mbar_wait(&tfull[slot], static_cast<uint32_t>((k >> 1) & 1));
      tc_fence_after();
      if (threadIdx.x == 128) IE_TRACE(4, k);
#pragma unroll
      for (int ch = 0; ch < kCh; ++ch) {
        uint32_t r[16];
        __syncwarp();
        tmem_ld16(taddr + ch * 16, r);
        tmem_ld_wait();
        const float cprev[4] = {cr[ch].x, cr[ch].y, cr[ch].z, cr[ch].w};
        float cnew[4], hn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float zi = __uint_as_float(r[4 * u + 0]) + gxr[ch][u].x;
          const float zf = __uint_as_float(r[4 * u + 1]) + gxr[ch][u].y;
          const float zg = __uint_as_float(r[4 * u + 2]) + gxr[ch][u].z;
          const float zo = __uint_as_float(r[4 * u + 3]) + gxr[ch][u].w;
          if (fast_math) {
            cnew[u] = sigmoid_fast(zf) * cprev[u] + sigmoid_fast(zi) * tanh_fast(zg);
            hn[u] = sigmoid_fast(zo) * tanh_fast(cnew[u]);
          } else {
            cnew[u] = sigmoid_acc(zf) * cprev[u] + sigmoid_acc(zi) * tanh_acc(zg);
            hn[u] = sigmoid_acc(zo) * tanh_acc(cnew[u]);
          }
        }
        __stcg(cp + ch, make_float4(cnew[0], cnew[1], cnew[2], cnew[3]));
        *reinterpret_cast<uint2*>(yrow + ch * 4) = make_uint2(pack_bf16x2(hn[0], hn[1]), pack_bf16x2(hn[2], hn[3]));
        if (raw != nullptr) {
          float4* rp = reinterpret_cast<float4*>(raw + (static_cast<long long>(brow) * T + t) * raw_ld + unit0 + ch * 4);
          *rp = make_float4(hn[0], hn[1], hn[2], hn[3]);
        }
        if (pooled && t < len) {
          const long long po = static_cast<long long>(brow) * out_pad + unit0 + ch * 4;
          float4* ps = reinterpret_cast<float4*>(pool_sum + po);
          float4* pm = reinterpret_cast<float4*>(pool_max + po);
          float4 s, m;
          if (t == 0) {
            s = make_float4(hn[0], hn[1], hn[2], hn[3]);
            m = s;
          } else {
            s = __ldcg(ps);
            m = __ldcg(pm);
            s.x += hn[0]; s.y += hn[1]; s.z += hn[2]; s.w += hn[3];
            m.x = fmaxf(m.x, hn[0]); m.y = fmaxf(m.y, hn[1]); m.z = fmaxf(m.z, hn[2]); m.w = fmaxf(m.w, hn[3]);
          }
          __stcg(ps, s);
          __stcg(pm, m);
          if (t == len - 1) __stcg(reinterpret_cast<float4*>(pool_last + po), make_float4(hn[0], hn[1], hn[2], hn[3]));
        }
      }
      // publish (step t, batch g): accumulator slot drained, h_t / c_t / pooling state visible
      if (threadIdx.x == 128) IE_TRACE(5, k);
      tc_fence_before();
      named_bar_sync(1, 512);
      if (threadIdx.x == 128) {
        mbar_arrive_remote(&tempty[slot], 0);
        __threadfence();
        red_relaxed_add(step_done + t * ng + g, 1u);
        IE_TRACE(6, k);
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
#undef IE_TRACE
#undef IE_TRACE_VAL
}

size_t rot_smem_bytes(int ast) {
  return 1024 + static_cast<size_t>(ast) * kRGA * 128 * 64 * 2 + static_cast<size_t>(kRWStages) * kRGW * kRHalfRows * 64 * 2 +
         (2 * ast + 2 * kRWStages + 4) * 8 + 16;
}

template <bool TOK, bool WFENCE, int AST>
cudaError_t launch_rot_t(const LstmWideArgs& a, int pairs, int tiles, cudaStream_t stream) {
  auto kfn = lstm_rot_kernel<TOK, WFENCE, AST>;
  const size_t smem = rot_smem_bytes(AST);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (a.check_only) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (a.num_sms / 2));
    cfg.blockDim = dim3(kRThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &cfg);
    if (e != cudaSuccess) return e;
    return max_clusters >= a.num_sms / 2 ? cudaSuccess : cudaErrorCooperativeLaunchTooLarge;
  }
  kfn<<<2 * pairs, kRThreads, smem, stream>>>(a.tm_h, a.tm_w, a.gx, a.c, a.y, a.raw, a.pool_sum, a.pool_max, a.pool_last,
                                               a.lengths, a.step_done, a.T, a.ng, tiles, a.out_pad, a.kh_pad / 64, a.ldy,
                                               a.raw_ld, a.fast_math, a.trace, a.trace_items, a.tok);
  return cudaGetLastError();
}

}  // namespace

int lstm_rot_pairs(const LstmWideArgs& a) {
  const long long total = static_cast<long long>(a.T) * a.ng * (a.n_cta / 2);
  long long pairs = a.num_sms / 2;
  if (pairs > total) pairs = total;
  return static_cast<int>(pairs);
}

// a.check_only: only verify co-residency of the grid.  Requires u == 32 per CTA (64 units per pair tile).
// a.variant (round-2 experiments, IE_ROT_VARIANT): bit 0 = proxy fence in the watcher warp, bit 1 = 4-stage h ring.
cudaError_t launch_lstm_rot(const LstmWideArgs& a, cudaStream_t stream) {
  if (a.u != 32 || a.n_cta % 2 || a.kh_pad % 64 || a.ng < 1 || a.ng > kRotMaxBatches || a.T < 1) return cudaErrorInvalidValue;
  const int tiles = a.n_cta / 2;
  const int pairs = lstm_rot_pairs(a);
  if (pairs < 1) return cudaErrorInvalidValue;
  const bool tok = a.tok != nullptr;  // layer 0 reading its input projection from the per-token table (IE_EMB_PROJ)
  switch (a.variant & 3) {
    case 0: return tok ? launch_rot_t<true, false, 3>(a, pairs, tiles, stream) : launch_rot_t<false, false, 3>(a, pairs, tiles, stream);
    case 1: return tok ? launch_rot_t<true, true, 3>(a, pairs, tiles, stream) : launch_rot_t<false, true, 3>(a, pairs, tiles, stream);
    case 2: return tok ? launch_rot_t<true, false, 4>(a, pairs, tiles, stream) : launch_rot_t<false, false, 4>(a, pairs, tiles, stream);
    default: return tok ? launch_rot_t<true, true, 4>(a, pairs, tiles, stream) : launch_rot_t<false, true, 4>(a, pairs, tiles, stream);
  }
}

}  // namespace ie
