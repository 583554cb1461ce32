// One LSTM timestep of one layer for the whole (<= 256 row) batch:
//
//     z      = Gx[t] + h_{t-1} * W_hh^T          (Gx = x_t W_ih^T + b_ih + b_hh, hoisted: gemm.cu)
//     i,f,o  = sigmoid(z_i, z_f, z_o) ; g = tanh(z_g)
//     c_t    = f*c_{t-1} + i*g ; h_t = o*tanh(c_t)
//
// which is what torch nn.LSTM computes per step inside fastai's AWD_LSTM.forward, called by the reference at
// Issue_Embeddings/flask_app/inference.py:57 / :68 (encoder.forward(x)[-1][-1]) with zero initial state
// (inference.py:56,66 reset()).  On the last layer the masked concat-pool of inference.py:239
// ([mean | max | last] over t < len) is accumulated in the same epilogue, so the (B,T,800) hidden-state tensor the
// reference copies to the host (inference.py:57) never leaves the GPU.
//
// Mapping.  CTA j owns u hidden units (all four gates: N = 4u accumulator columns, rows of W_hh pre-permuted at
// load time to [cta][unit][gate]) for all batch rows.  The batch is the UMMA M dimension (one or two 128-row
// tiles, each row = one TMEM lane = one epilogue thread), so a thread finds i,f,g,o of a unit in four adjacent
// accumulator columns and the cell update needs no cross-thread traffic.
//   warp 0       TMA producer: per 64-wide K block, h_{t-1} tile(s) 128x64 + W_hh slice 4u x 64 (128B swizzle)
//   warp 1       UMMA issuer : tcgen05.mma kind::f16, M=128, N=4u, bf16 operands, f32 accumulate in TMEM
//   warp 2       TMEM allocator
//   warps 4..    epilogue    : prefetch Gx/c into registers while the MMAs run, tcgen05.ld, gates, state update,
//                              h_t (bf16) into slot t+1 of the hidden-state ring = next step's A operand and the
//                              next layer's GEMM input
#include "kernels.h"
#include "ptx.cuh"

namespace ie {

namespace {

constexpr int kStepStages = 4;

template <int NCH>
__global__ void __launch_bounds__(384, 1)
lstm_step_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_hs,
                 const __grid_constant__ CUtensorMap tm_w, const float* __restrict__ gx, float* __restrict__ cstate, __nv_bfloat16* __restrict__ y,
                 float* __restrict__ raw, float* __restrict__ pool_sum, float* __restrict__ pool_max,
                 float* __restrict__ pool_last, const int* __restrict__ lengths, int t, int T, int b_pad,
                 int out_pad, int num_k_blocks, long long ldy, long long raw_ld, int tmem_cols, int fast_math) {
  constexpr int U = NCH * 4;  // hidden units per CTA
  constexpr int N = U * 4;    // accumulator columns per M tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t rawaddr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (rawaddr & 1023u)) & 1023u);

  const int m_tiles = b_pad >> 7;
  const uint32_t a_tile_bytes = 128 * 64 * 2;
  const uint32_t a_bytes = a_tile_bytes * m_tiles;
  const uint32_t b_bytes = N * 64 * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStepStages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStepStages;
  uint64_t* tfull_bar = bars + 2 * kStepStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStepStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Cluster of `csize` CTAs (different unit slices, same batch rows): every h_{t-1} tile is fetched from L2 once
  // per cluster -- CTA r loads rows [r*128/csize, ...) of each tile and multicasts them to all CTAs of the cluster.
  const uint32_t csize = cluster_nctarank();
  const uint32_t crank = cluster_ctarank();
  const uint16_t cmask = static_cast<uint16_t>((1u << csize) - 1u);
  const uint32_t sub_rows = 128u / csize;
  const uint32_t sub_bytes = sub_rows * 128u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(csize > 1 ? &tm_hs : &tm_h);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStepStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], csize);  // one commit-arrive from every CTA that reads a stage we multicast into
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  if (csize > 1) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        // weights first: they do not depend on the previous step
        tma_load_2d(sa + a_bytes, &tm_w, &full_bar[stage], kb * 64, blockIdx.x * N, kEvictLast);
        if (csize == 1) {
          for (int mt = 0; mt < m_tiles; ++mt)
            tma_load_2d(sa + mt * a_tile_bytes, &tm_h, &full_bar[stage], kb * 64, t * b_pad + mt * 128, kEvictFirst);
        } else {
          for (int mt = 0; mt < m_tiles; ++mt)
            tma_load_2d_mc(sa + mt * a_tile_bytes + crank * sub_bytes, &tm_hs, &full_bar[stage], kb * 64,
                           t * b_pad + mt * 128 + static_cast<int>(crank * sub_rows), cmask, kEvictFirst);
        }
        if (++stage == kStepStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, N);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint64_t db = umma_desc_sw128(sa + a_bytes);
        for (int mt = 0; mt < m_tiles; ++mt) {
          const uint64_t da = umma_desc_sw128(sa + mt * a_tile_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + mt * N, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        if (csize == 1) umma_commit(&empty_bar[stage]); else umma_commit_mc(&empty_bar[stage], cmask);
        if (++stage == kStepStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else if (warp >= 4 && warp < 4 + 4 * m_tiles) {
    const int e = warp - 4;
    const int mt = e >> 2;
    const int q = e & 3;
    const int row = mt * 128 + q * 32 + lane;  // batch row
    const int unit0 = blockIdx.x * U;

    // prefetch the step's Gx slice and the cell state while the MMAs run
    const float4* gxp =
        reinterpret_cast<const float4*>(gx + (static_cast<long long>(t) * b_pad + row) * (4ll * out_pad) + 4ll * unit0);
    float4 gxr[NCH * 4];
#pragma unroll
    for (int i = 0; i < NCH * 4; ++i) gxr[i] = __ldg(gxp + i);
    float4* cp = reinterpret_cast<float4*>(cstate + static_cast<long long>(row) * out_pad + unit0);
    float4 cr[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) cr[i] = (t == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : cp[i];
    int len = 1;
    if (pool_sum != nullptr) len = lengths[row];

    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mt * N);
    __nv_bfloat16* yrow = y + (static_cast<long long>(t + 1) * b_pad + row) * ldy + unit0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      uint32_t r[16];
      __syncwarp();  // the pooling branch below may diverge; tcgen05.ld is .sync.aligned
      tmem_ld16(taddr + ch * 16, r);
      tmem_ld_wait();
      float cprev[4] = {cr[ch].x, cr[ch].y, cr[ch].z, cr[ch].w};
      float cn[4], hn[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 gq = gxr[ch * 4 + j];
        const float zi = __uint_as_float(r[4 * j + 0]) + gq.x;
        const float zf = __uint_as_float(r[4 * j + 1]) + gq.y;
        const float zg = __uint_as_float(r[4 * j + 2]) + gq.z;
        const float zo = __uint_as_float(r[4 * j + 3]) + gq.w;
        // same expressions (and association) as lstm_seq.cu so that both kernels give identical bits
        if (fast_math) {
          cn[j] = sigmoid_fast(zf) * cprev[j] + sigmoid_fast(zi) * tanh_fast(zg);
          hn[j] = sigmoid_fast(zo) * tanh_fast(cn[j]);
        } else {
          cn[j] = sigmoid_acc(zf) * cprev[j] + sigmoid_acc(zi) * tanh_acc(zg);
          hn[j] = sigmoid_acc(zo) * tanh_acc(cn[j]);
        }
      }
      cp[ch] = make_float4(cn[0], cn[1], cn[2], cn[3]);
      *reinterpret_cast<uint2*>(yrow + ch * 4) = make_uint2(pack_bf16x2(hn[0], hn[1]), pack_bf16x2(hn[2], hn[3]));
      if (raw != nullptr) {
        float4* rp = reinterpret_cast<float4*>(raw + (static_cast<long long>(row) * T + t) * raw_ld + unit0 + ch * 4);
        *rp = make_float4(hn[0], hn[1], hn[2], hn[3]);
      }
      if (pool_sum != nullptr && t < len) {
        const long long po = static_cast<long long>(row) * out_pad + unit0 + ch * 4;
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
  }

  __syncwarp();  // re-converge the single-lane role loops before the (aligned) barrier
  tc_fence_before();
  // no CTA may exit while a peer can still multicast into its shared memory or arrive on its barriers
  if (csize > 1) cluster_sync(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

size_t step_smem_bytes(int m_tiles, int n) {
  return 1024 + static_cast<size_t>(kStepStages) * (m_tiles * 128 * 64 * 2 + n * 64 * 2) + (2 * kStepStages + 1) * 8 + 16;
}

template <int NCH>
cudaError_t launch_step_t(const LstmStepArgs& a, cudaStream_t stream) {
  const int m_tiles = a.b_pad / 128;
  const int n = NCH * 16;
  int tmem_cols = 32;
  while (tmem_cols < m_tiles * n) tmem_cols <<= 1;
  const size_t smem = step_smem_bytes(m_tiles, n);
  auto kfn = lstm_step_kernel<NCH>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.n_cta);
  cfg.blockDim = dim3(128 + 128 * m_tiles);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.cluster > 0 ? a.cluster : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kfn, a.tm_h, a.tm_hs, a.tm_w, a.gx, a.c, a.y, a.raw, a.pool_sum, a.pool_max,
                            a.pool_last, a.lengths, a.t, a.T, a.b_pad, a.out_pad, a.kh_pad / 64, a.ldy, a.raw_ld,
                            tmem_cols, a.fast_math);
}

}  // namespace

cudaError_t launch_lstm_step(const LstmStepArgs& a, cudaStream_t stream) {
  if (a.u % 4 || a.u < 4 || (a.b_pad != 128 && a.b_pad != 256) || a.kh_pad % 64) return cudaErrorInvalidValue;
  if (a.cluster > 1 && (a.n_cta % a.cluster || 128 % a.cluster || a.cluster > 8)) return cudaErrorInvalidValue;
  switch (a.u / 4) {
    case 1: return launch_step_t<1>(a, stream);
    case 2: return launch_step_t<2>(a, stream);
    case 3: return launch_step_t<3>(a, stream);
    case 4: return launch_step_t<4>(a, stream);
    case 5: return launch_step_t<5>(a, stream);
    case 6: return launch_step_t<6>(a, stream);
    case 7: return launch_step_t<7>(a, stream);
    case 8: return launch_step_t<8>(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace ie
