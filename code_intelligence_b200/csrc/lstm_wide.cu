// Persistent recurrent kernel, "wide tile" variant: three independent batches of 256 rows per launch, CTA pairs that own
// N = 256 accumulator columns (64 hidden units) per tile.  Same arithmetic as lstm_seq.cu / lstm.cu (reference call
// sites: Issue_Embeddings/flask_app/inference.py:56-57, :66-68, pooling :239).
//
// Why (profiles/README.md, "UMMA issue rate"): on this hardware a tcgen05.mma costs ~91 ns to issue whatever its
// shape, so a recurrent step costs (K/16 = 152 instructions) x 91 ns = 13.8 us per (batch, tile) CHAIN however narrow
// the tile is.  lstm_seq.cu runs 60 pairs x 2 batches = 2 chains per pair per timestep on N = 160 tiles.  Here the
// tiles are N = 256 (38 per batch at H = 2400, the last one half padding) and 3 batches ride the launch: 114 chains
// are dealt round-robin over all 74 CTA pairs of the chip (40 pairs run 2 chains, 34 run 1), i.e. the same 2 chains
// per pair per timestep now serve 3 batches instead of 2.
//
// Chain c = g * tiles + j (batch g, tile j) belongs to pair c mod P; a pair's chains always belong to different
// batches (tiles <= P), and every pair walks its chains in increasing batch order, so the per-(step, batch) counters
// cannot form a wait cycle.  Each chain has its own TMEM accumulator (2 x 256 columns).  Everything else -- CTA pair
// UMMA (cta_group::2, M = 256), TMA rings for h and W_hh, step counters, fences -- is as in lstm_seq.cu.  The
// epilogue streams Gx and the cell state per 16-column chunk (it is off the critical path here: 2 x ~3 us per timestep
// against ~30 us of MMAs), c_t lives in global memory (L2 resident, 7.5 MB).
#include <cmath>

#include "kernels.h"
#include "ptx.cuh"

namespace ie {

namespace {

constexpr int kWThreads = 640;  // 4 role warps + 16 epilogue warps (4 per TMEM lane quarter, 64 columns each)
constexpr int kWGA = 2, kWAStages = 3;  // h ring: 3 stages x 2 k-blocks x 16 KB
constexpr int kWGW = 2, kWWStages = 3;  // W ring: 3 stages x 2 k-blocks x 16 KB
constexpr int kTileN = 256;             // accumulator columns per tile = 64 hidden units
constexpr int kHalfRows = 128;          // W rows each CTA of the pair contributes

template <bool TOK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWThreads, 1)
lstm_wide_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w,
                 const float* __restrict__ gx, float* __restrict__ cstate, __nv_bfloat16* __restrict__ y,
                 float* __restrict__ raw, float* __restrict__ pool_sum, float* __restrict__ pool_max,
                 float* __restrict__ pool_last, const int* __restrict__ lengths, unsigned* __restrict__ step_done, int T,
                 int ng, int tiles, int out_pad, int num_k_blocks, long long ldy, long long raw_ld, int fast_math,
                 long long* __restrict__ trace, const int* __restrict__ tok) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t rawaddr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (rawaddr & 1023u)) & 1023u);

  constexpr uint32_t a_bytes = 128 * 64 * 2;
  constexpr uint32_t w_bytes = kHalfRows * 64 * 2;
  uint8_t* a_ring = smem;
  uint8_t* w_ring = smem + kWAStages * kWGA * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_ring + kWWStages * kWGW * w_bytes);
  uint64_t* afull = bars;                   // [kWAStages] leader's copy is live
  uint64_t* aempty = afull + kWAStages;
  uint64_t* wfull = aempty + kWAStages;     // [kWWStages]
  uint64_t* wempty = wfull + kWWStages;
  uint64_t* tfull = wempty + kWWStages;     // [2] one per local chain
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // optional timeline of the pair's FIRST chain: [cta][t][12] (%globaltimer ns; slots 8-10 SM cycles); null in production
#define IE_TRACE(slot, tt) do { if (trace) { unsigned long long _g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_g)); \
    trace[(static_cast<long long>(blockIdx.x) * T + (tt)) * 12 + (slot)] = static_cast<long long>(_g); } } while (0)
#define IE_TRACE_VAL(slot, tt, v) do { if (trace) trace[(static_cast<long long>(blockIdx.x) * T + (tt)) * 12 + (slot)] = (v); } while (0)
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int P = static_cast<int>(gridDim.x >> 1);
  const int n_chains = ng * tiles;
  const int my_chains = (n_chains - pair + P - 1) / P;  // chains pair, pair + P, ...   (1 or 2)
  const int b_pad = 256 * ng;
  const unsigned batch_ctas = 2u * static_cast<unsigned>(tiles);  // CTAs that publish a (step, batch)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_h);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWAStages; ++s) {
      mbar_init(&afull[s], 2);
      mbar_init(&aempty[s], 1);
    }
    for (int s = 0; s < kWWStages; ++s) {
      mbar_init(&wfull[s], 2);
      mbar_init(&wempty[s], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
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
      for (int t = 0; t < T; ++t) {
        for (int ci = 0; ci < my_chains; ++ci) {
          const int g = (pair + ci * P) / tiles;
          if (t > 0) {
            wait_flag_ge_relaxed(step_done + (t - 1) * ng + g, batch_ctas);
            fence_proxy_async();
          }
          if (ci == 0) IE_TRACE(0, t);
          const int row0 = t * b_pad + g * 256 + static_cast<int>(crank) * 128;
          for (int kb0 = 0; kb0 < num_k_blocks; kb0 += kWGA) {
            const int n = min(kWGA, num_k_blocks - kb0);
            mbar_wait(&aempty[stage], phase ^ 1);
            if (crank == 0) mbar_arrive_expect_tx(&afull[stage], 2 * n * a_bytes);
            else mbar_arrive_remote(&afull[stage], 0);
            for (int j = 0; j < n; ++j)
              tma_load_2d_pair(a_ring + (stage * kWGA + j) * a_bytes, &tm_h, &afull[stage], (kb0 + j) * 64, row0,
                               kEvictNormal);
            if (++stage == kWAStages) { stage = 0; phase ^= 1; }
          }
          if (ci == 0) IE_TRACE(1, t);
        }
      }
    }
  } else if (warp == 3) {
    // ---------------- W producer: free-running ahead of h ----------------------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < T; ++t) {
        for (int ci = 0; ci < my_chains; ++ci) {
          const int j = (pair + ci * P) % tiles;
          const int wrow0 = (2 * j + static_cast<int>(crank)) * kHalfRows;  // slices are [cta][unit][gate], 128 rows each
          for (int kb0 = 0; kb0 < num_k_blocks; kb0 += kWGW) {
            const int n = min(kWGW, num_k_blocks - kb0);
            mbar_wait(&wempty[stage], phase ^ 1);
            if (crank == 0) mbar_arrive_expect_tx(&wfull[stage], 2 * n * w_bytes);
            else mbar_arrive_remote(&wfull[stage], 0);
            for (int q = 0; q < n; ++q)
              tma_load_2d_pair(w_ring + (stage * kWGW + q) * w_bytes, &tm_w, &wfull[stage], (kb0 + q) * 64, wrow0,
                               kEvictLast);
            if (++stage == kWWStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- UMMA issuer (leader CTA) --------------------------------------------------------------
    if (crank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, kTileN);
      const uint32_t a_base = smem_u32(a_ring);
      const uint32_t w_base = smem_u32(w_ring);
      int as = 0, ws = 0;
      uint32_t aph = 0, wph = 0;
      for (int t = 0; t < T; ++t) {
        for (int ci = 0; ci < my_chains; ++ci) {
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(ci * kTileN);
          long long wa = 0, ww = 0, t_first = 0;
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            const int ja = kb % kWGA, jw = kb % kWGW;
            if (ja == 0) {
              const long long c0 = trace ? clock64() : 0;
              mbar_wait(&afull[as], aph);
              if (kb == 0) { if (ci == 0) IE_TRACE(2, t); t_first = trace ? clock64() : 0; }
              else if (trace) wa += clock64() - c0;
            }
            if (jw == 0) {
              const long long c0 = trace ? clock64() : 0;
              mbar_wait(&wfull[ws], wph);
              if (trace) ww += clock64() - c0;
            }
            tc_fence_after();
            const uint64_t da = umma_desc_sw128(a_base + (as * kWGA + ja) * a_bytes);
            const uint64_t db = umma_desc_sw128(w_base + (ws * kWGW + jw) * w_bytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            const bool last = (kb == num_k_blocks - 1);
            if (ja == kWGA - 1 || last) {
              umma_commit_pair_mc(&aempty[as], 0x3);
              if (++as == kWAStages) { as = 0; aph ^= 1; }
            }
            if (jw == kWGW - 1 || last) {
              umma_commit_pair_mc(&wempty[ws], 0x3);
              if (++ws == kWWStages) { ws = 0; wph ^= 1; }
            }
          }
          umma_commit_pair_mc(&tfull[ci], 0x3);
          if (ci == 0) {
            IE_TRACE(3, t);
            IE_TRACE_VAL(8, t, wa);
            IE_TRACE_VAL(9, t, ww);
            IE_TRACE_VAL(10, t, trace ? clock64() - t_first : 0);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ------------------------------------------------------------------------------
    const int e = warp - 4;
    const int q = e & 3;
    const int cq = e >> 2;  // which 64 of the tile's 256 columns (4 chunks of 16 = 16 hidden units per thread)
    const int row = static_cast<int>(crank) * 128 + q * 32 + lane;
    const bool pooled = pool_sum != nullptr;
    for (int t = 0; t < T; ++t) {
      for (int ci = 0; ci < my_chains; ++ci) {
        const int c = pair + ci * P;
        const int g = c / tiles, j = c % tiles;
        const int brow = g * 256 + row;
        const int unit0 = j * 64 + cq * 16;
        const int len = pooled ? lengths[brow] : 1;
        const long long grow = TOK ? static_cast<long long>(__ldg(tok + static_cast<long long>(t) * b_pad + brow))
                                   : static_cast<long long>(t) * b_pad + brow;  // TOK: per-token projection table
        const float4* gxp = reinterpret_cast<const float4*>(gx + grow * (4ll * out_pad) +
                                                            4ll * unit0);
        float4* cp = reinterpret_cast<float4*>(cstate + static_cast<long long>(brow) * out_pad + unit0);
        __nv_bfloat16* yrow = y + (static_cast<long long>(t + 1) * b_pad + brow) * ldy + unit0;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(ci * kTileN + cq * 64);
        // all of this thread's Gx (4 chunks x 4 units x 4 gates) and c are loaded while the MMAs still run
        constexpr int kCh = 4;
        float4 gxr[kCh][4];
        float4 cr[kCh];
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
          ldg_stream8(reinterpret_cast<const float*>(gxp + ch * 4), gxr[ch][0], gxr[ch][1]);
          ldg_stream8(reinterpret_cast<const float*>(gxp + ch * 4 + 2), gxr[ch][2], gxr[ch][3]);
          cr[ch] = (t == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : cp[ch];
        }
        if (threadIdx.x == 128 && ci == 0) IE_TRACE(7, t);
        mbar_wait(&tfull[ci], static_cast<uint32_t>(t & 1));
        tc_fence_after();
        if (threadIdx.x == 128 && ci == 0) IE_TRACE(4, t);
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
          cp[ch] = make_float4(cnew[0], cnew[1], cnew[2], cnew[3]);
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
              s = *ps;
              m = *pm;
              s.x += hn[0]; s.y += hn[1]; s.z += hn[2]; s.w += hn[3];
              m.x = fmaxf(m.x, hn[0]); m.y = fmaxf(m.y, hn[1]); m.z = fmaxf(m.z, hn[2]); m.w = fmaxf(m.w, hn[3]);
            }
            *ps = s;
            *pm = m;
            if (t == len - 1) *reinterpret_cast<float4*>(pool_last + po) = make_float4(hn[0], hn[1], hn[2], hn[3]);
          }
        }
        // publish (step t, batch g): accumulator drained, h_t visible
        if (threadIdx.x == 128 && ci == 0) IE_TRACE(5, t);
        tc_fence_before();
        named_bar_sync(1, 512);
        if (threadIdx.x == 128) {
          __threadfence();
          red_relaxed_add(step_done + t * ng + g, 1u);
          if (ci == 0) IE_TRACE(6, t);
        }
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
}

size_t wide_smem_bytes() {
  return 1024 + static_cast<size_t>(kWAStages) * kWGA * 128 * 64 * 2 + static_cast<size_t>(kWWStages) * kWGW * kHalfRows * 64 * 2 +
         (2 * kWAStages + 2 * kWWStages + 2) * 8 + 16;
}

}  // namespace

// a.check_only: only verify co-residency of the grid.  Requires u == 32 per CTA (64 units per pair tile).
cudaError_t launch_lstm_wide(const LstmWideArgs& a, cudaStream_t stream) {
  if (a.u != 32 || a.n_cta % 2 || a.kh_pad % 64 || a.ng < 1 || a.ng > 3) return cudaErrorInvalidValue;
  const int tiles = a.n_cta / 2;
  const int n_chains = a.ng * tiles;
  int pairs = a.num_sms / 2;
  if (pairs > n_chains) pairs = n_chains;
  if (tiles > pairs || (n_chains + pairs - 1) / pairs > 2) return cudaErrorInvalidValue;  // <= 2 chains (TMEM), distinct batches
  const size_t smem = wide_smem_bytes();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(lstm_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lstm_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (a.check_only) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kWThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, lstm_wide_kernel<false>, &cfg);
    if (e != cudaSuccess) return e;
    return max_clusters >= pairs ? cudaSuccess : cudaErrorCooperativeLaunchTooLarge;
  }
  if (a.tok != nullptr)  // layer 0 reading its input projection from the per-token table (api.cu, IE_EMB_PROJ)
    lstm_wide_kernel<true><<<2 * pairs, kWThreads, smem, stream>>>(
        a.tm_h, a.tm_w, a.gx, a.c, a.y, a.raw, a.pool_sum, a.pool_max, a.pool_last, a.lengths, a.step_done, a.T, a.ng, tiles,
        a.out_pad, a.kh_pad / 64, a.ldy, a.raw_ld, a.fast_math, a.trace, a.tok);
  else
    lstm_wide_kernel<false><<<2 * pairs, kWThreads, smem, stream>>>(
        a.tm_h, a.tm_w, a.gx, a.c, a.y, a.raw, a.pool_sum, a.pool_max, a.pool_last, a.lengths, a.step_done, a.T, a.ng, tiles,
        a.out_pad, a.kh_pad / 64, a.ldy, a.raw_ld, a.fast_math, a.trace, nullptr);
  return cudaGetLastError();
}

}  // namespace ie
