// Fallback recurrent kernel: ONE LSTM timestep of one layer for one 256-row batch per launch.  Used only when the
// persistent kernel (lstm_layer.cu) cannot get its whole grid co-resident (or with IE_SEQ=0, for testing); same
// arithmetic through lstm_common.cuh, same weight layout ([slice][unit][gate], 32 units per slice), same Gx formats,
// hence the same bits.
//
//     z      = Gx[t] + h_{t-1} * W_hh^T          (Gx = x_t W_ih^T + b_ih + b_hh, hoisted: gemm.cu)
//     i,f,o  = sigmoid(z_i, z_f, z_o) ; g = tanh(z_g)
//     c_t    = f*c_{t-1} + i*g ; h_t = o*tanh(c_t)
//
// which is what torch nn.LSTM computes per step inside fastai's AWD_LSTM.forward, called by the reference at
// Issue_Embeddings/flask_app/inference.py:57 / :68 (encoder.forward(x)[-1][-1]) with zero initial state
// (inference.py:56,66 reset()).  On the last layer the masked concat-pool of inference.py:239 is accumulated in the same
// epilogue.
//
// Mapping.  CTA j owns 32 hidden units (N = 128 accumulator columns) for the batch's 256 rows (two M = 128 tiles, each
// row = one TMEM lane = one epilogue thread).
//   warp 0       TMA producer: per 64-wide K block, two h_{t-1} tiles 128x64 + the W_hh slice 128 x 64 (128B swizzle)
//   warp 1       UMMA issuer : tcgen05.mma kind::f16, M=128, N=128, bf16 operands, f32 accumulate in TMEM
//   warp 2       TMEM allocator
//   warps 4..11  epilogue
#include "kernels.h"
#include "lstm_common.cuh"
#include "ptx.cuh"

namespace ie {

namespace {

constexpr int kStepStages = 4;
constexpr int kStepN = 128;   // 32 units x 4 gates
constexpr int kStepThreads = 384;

template <bool TOK, bool GXBF>
__global__ void __launch_bounds__(kStepThreads, 1)
lstm_step_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w,
                 const void* __restrict__ gx, const int* __restrict__ tok, float* cstate, __nv_bfloat16* __restrict__ y,
                 float* __restrict__ raw, float* pool_sum, float* pool_max, float* pool_last,
                 const int* __restrict__ lengths, unsigned* abort_flag, long long spin_limit, int t, int t0, int T_total,
                 int b_pad, int g, int out_pad, int nkb, int segs, int kh_pad, long long ldy, long long raw_ld,
                 int gate_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t rawaddr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (rawaddr & 1023u)) & 1023u);

  constexpr uint32_t a_tile_bytes = 128 * 64 * 2;
  constexpr uint32_t a_bytes = 2 * a_tile_bytes;
  constexpr uint32_t b_bytes = kStepN * 64 * 2;
  constexpr uint32_t stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStepStages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStepStages;
  uint64_t* tfull_bar = bars + 2 * kStepStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStepStages + 1);
  uint32_t* abort_s = tmem_slot + 1;
  const Abort ab{abort_s, abort_flag, spin_limit};

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tg = t0 + t;          // global timestep
  const int nkt = nkb * segs;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_h);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStepStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    *abort_s = 0;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int row0 = t * b_pad + g * 256;  // ring slot t = h_{t-1} (chunk-local)
      for (int kb = 0; kb < nkt && !aborted(ab); ++kb) {
        const int seg = kb / nkb, r = kb - seg * nkb;   // split-bf16: [h_hi | h_lo | h_hi] x [W_hi | W_hi | W_lo]
        mbar_wait(&empty_bar[stage], phase ^ 1, ab);
        uint8_t* sa = smem + stage * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        // weights first: they do not depend on the previous step
        tma_load_2d(sa + a_bytes, &tm_w, &full_bar[stage], (seg == 2 ? kh_pad : 0) + r * 64, blockIdx.x * kStepN, kEvictLast);
        for (int mt = 0; mt < 2; ++mt)
          tma_load_2d(sa + mt * a_tile_bytes, &tm_h, &full_bar[stage], (seg == 1 ? kh_pad : 0) + r * 64, row0 + mt * 128,
                      kEvictFirst);
        if (++stage == kStepStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, kStepN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkt && !aborted(ab); ++kb) {
        mbar_wait(&full_bar[stage], phase, ab);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint64_t db = umma_desc_sw128(sa + a_bytes);
        for (int mt = 0; mt < 2; ++mt) {
          const uint64_t da = umma_desc_sw128(sa + mt * a_tile_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + mt * kStepN, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kStepStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    const int e = warp - 4;
    const int mt = e >> 2;
    const int q = e & 3;
    const int brow = g * 256 + mt * 128 + q * 32 + lane;  // row inside the time slot
    const int unit0 = blockIdx.x * 32;
    const bool pooled = pool_sum != nullptr;
    const long long lo_off = segs > 1 ? kh_pad : 0;
    const int len = pooled ? lengths[brow] : 1;
    const long long grow = TOK ? static_cast<long long>(__ldg(tok + static_cast<long long>(tg) * b_pad + brow))
                               : static_cast<long long>(t) * b_pad + brow;
    float* cp = cstate + static_cast<long long>(brow) * out_pad + unit0;
    __nv_bfloat16* yrow = y + (static_cast<long long>(t + 1) * b_pad + brow) * ldy + unit0;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mt * kStepN);
    // two halves of 16 units (4 chunks each) to bound the registers
    for (int hh = 0; hh < 2; ++hh) {
      constexpr int kCh = 4;
      constexpr int kGW = GXBF ? 8 : 16;
      uint32_t gxw[kCh][kGW];
      float4 cr[kCh];
      if constexpr (GXBF) {
        const __half* gxp = reinterpret_cast<const __half*>(gx) + grow * (4ll * out_pad) + 4ll * (unit0 + hh * 16);
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) ldg_stream8_b32(gxp + ch * 16, &gxw[ch][0]);
      } else {
        const float* gxp = reinterpret_cast<const float*>(gx) + grow * (4ll * out_pad) + 4ll * (unit0 + hh * 16);
#pragma unroll
        for (int ch = 0; ch < kCh; ++ch) {
          ldg_stream8_b32(gxp + ch * 16, &gxw[ch][0]);
          ldg_stream8_b32(gxp + ch * 16 + 8, &gxw[ch][kGW - 8]);
        }
      }
#pragma unroll
      for (int ch = 0; ch < kCh; ++ch)
        cr[ch] = (tg == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcg(reinterpret_cast<const float4*>(cp) + hh * 4 + ch);
      if (hh == 0) {
        mbar_wait(tfull_bar, 0, ab);
        tc_fence_after();
      }
#pragma unroll
      for (int ch = 0; ch < kCh; ++ch) {
        uint32_t r[16];
        __syncwarp();  // tcgen05.ld is .sync.aligned
        tmem_ld16(taddr + hh * 64 + ch * 16, r);
        tmem_ld_wait();
        float4 gx4[4];
        if constexpr (GXBF) {
          gx_unpack_f16(gxw[ch], gx4);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            gx4[u] = make_float4(__uint_as_float(gxw[ch][4 * u]), __uint_as_float(gxw[ch][4 * u + 1]),
                                 __uint_as_float(gxw[ch][4 * u + 2]), __uint_as_float(gxw[ch][4 * u + 3]));
        }
        const float cprev[4] = {cr[ch].x, cr[ch].y, cr[ch].z, cr[ch].w};
        float cnew[4], hn[4];
        lstm_cell4(r, gx4, cprev, cnew, hn, gate_mode);
        const int uo = hh * 16 + ch * 4;
        __stcg(reinterpret_cast<float4*>(cp + uo), make_float4(cnew[0], cnew[1], cnew[2], cnew[3]));
        store_h4(yrow + uo, hn, lo_off);
        if (raw != nullptr) {
          float4* rp = reinterpret_cast<float4*>(raw + (static_cast<long long>(brow) * T_total + tg) * raw_ld + unit0 + uo);
          *rp = make_float4(hn[0], hn[1], hn[2], hn[3]);
        }
        if (pooled) {
          const long long po = static_cast<long long>(brow) * out_pad + unit0 + uo;
          const float4 mprev = (tg > 0 && tg < len) ? __ldcg(reinterpret_cast<const float4*>(pool_max + po))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
          pool_accumulate4(pool_sum, pool_max, pool_last, po, hn, mprev, tg, len);
        }
      }
    }
  }

  __syncwarp();  // re-converge the single-lane role loops before the (aligned) barrier
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

size_t step_smem_bytes() {
  return 1024 + static_cast<size_t>(kStepStages) * (2 * 128 * 64 * 2 + kStepN * 64 * 2) + (2 * kStepStages + 1) * 8 + 32;
}

template <bool TOK, bool GXBF>
cudaError_t launch_step_t(const LstmStepArgs& a, cudaStream_t stream) {
  const size_t smem = step_smem_bytes();
  auto kfn = lstm_step_kernel<TOK, GXBF>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));  // per device
  if (e != cudaSuccess) return e;
  kfn<<<a.n_cta, kStepThreads, smem, stream>>>(a.tm_h, a.tm_w, a.gx, a.tok, a.c, a.y, a.raw, a.pool_sum, a.pool_max,
                                                a.pool_last, a.lengths, a.abort_flag,
                                                a.spin_limit > 0 ? a.spin_limit : kSpinLimitDefault, a.t, a.t0, a.T_total,
                                                a.b_pad, a.g, a.out_pad, a.kh_pad / 64, a.segs, a.kh_pad, a.ldy, a.raw_ld,
                                                a.gate_mode);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_lstm_step(const LstmStepArgs& a, cudaStream_t stream) {
  if (a.u != 32 || a.b_pad % 256 || a.kh_pad % 64 || (a.segs != 1 && a.segs != 3)) return cudaErrorInvalidValue;
  const bool tok = a.tok != nullptr;
  if (a.gx_bf16) return tok ? launch_step_t<true, true>(a, stream) : launch_step_t<false, true>(a, stream);
  return tok ? launch_step_t<true, false>(a, stream) : launch_step_t<false, false>(a, stream);
}

}  // namespace ie
