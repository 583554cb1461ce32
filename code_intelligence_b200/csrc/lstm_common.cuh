// Arithmetic shared by the recurrent kernels (lstm_layer.cu = the persistent kernel, lstm.cu = the per-timestep
// fallback): the LSTM cell update on one 16-column accumulator chunk and the masked concat-pool accumulation.  Both
// kernels call exactly these functions with the same association of operations, so every path gives the same bits.
//
// Reference arithmetic: torch nn.LSTM as wrapped by fastai's AWD_LSTM, called at
// Issue_Embeddings/flask_app/inference.py:57,68 (gate rows i|f|g|o, c_t = f*c_{t-1} + i*g, h_t = o*tanh(c_t));
// pooling: inference.py:239 ([mean | max | last] over the first len_i steps).
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace ie {

// gate precision levels (LstmLayerArgs::gate_mode / LstmStepArgs::gate_mode)
constexpr int kGatesFast = 2;   // tanh.approx.f32 (1 MUFU per transcendental, rel err 2^-11): the bf16 default
constexpr int kGatesExp = 1;    // ex2.approx + rcp.approx (abs err ~1e-7): IE_CFG_ACCURATE_GATES
constexpr int kGatesIeee = 0;   // expf + IEEE division: the fp32-accurate mode (IE_CFG_FP32)

__device__ __forceinline__ float sigmoid_ieee(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float tanh_ieee(float x) {
  // 1 - 2/(e^{2x}+1) loses relative accuracy near 0: use expm1 there; odd in x
  const float ax = fabsf(x);
  float t;
  if (ax < 0.55f) {
    const float e = expm1f(2.0f * ax);
    t = e / (e + 2.0f);
  } else {
    t = 1.0f - 2.0f / (expf(2.0f * ax) + 1.0f);
  }
  return copysignf(t, x);
}

// One accumulator chunk: 16 TMEM columns = 4 hidden units x (i, f, g, o) of one batch row.
//   acc : the h_{t-1} W_hh^T part (f32 bits from tcgen05.ld)       gx : x_t W_ih^T + b_ih + b_hh (4 units x 4 gates)
__device__ __forceinline__ void lstm_cell4(const uint32_t (&acc)[16], const float4 (&gx)[4], const float (&cprev)[4],
                                           float (&cnew)[4], float (&hn)[4], int gate_mode) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float zi = __uint_as_float(acc[4 * u + 0]) + gx[u].x;
    const float zf = __uint_as_float(acc[4 * u + 1]) + gx[u].y;
    const float zg = __uint_as_float(acc[4 * u + 2]) + gx[u].z;
    const float zo = __uint_as_float(acc[4 * u + 3]) + gx[u].w;
    if (gate_mode == kGatesFast) {
      cnew[u] = sigmoid_fast(zf) * cprev[u] + sigmoid_fast(zi) * tanh_fast(zg);
      hn[u] = sigmoid_fast(zo) * tanh_fast(cnew[u]);
    } else if (gate_mode == kGatesExp) {
      cnew[u] = sigmoid_acc(zf) * cprev[u] + sigmoid_acc(zi) * tanh_acc(zg);
      hn[u] = sigmoid_acc(zo) * tanh_acc(cnew[u]);
    } else {
      cnew[u] = sigmoid_ieee(zf) * cprev[u] + sigmoid_ieee(zi) * tanh_ieee(zg);
      hn[u] = sigmoid_ieee(zo) * tanh_ieee(cnew[u]);
    }
  }
}

// fp16 x 16 (one 256-bit load) -> the 4 x float4 Gx operands of a chunk.  The hoisted input projections are stored as
// IEEE half (11-bit significand: 8x finer than bf16 at the same bytes; |Gx| is O(1), far inside the half range)
__device__ __forceinline__ void gx_unpack_f16(const uint32_t* p, float4 (&gx)[4]) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&p[2 * u]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&p[2 * u + 1]));
    gx[u] = make_float4(a.x, a.y, b.x, b.y);
  }
}

// ---- masked concat-pool accumulators in global memory ------------------------------------------------------------
// pool_sum : f32, sequential sum over t (one add per timestep, in timestep order): an L2 reduction (red.add.v4.f32) --
//            no load, no latency, no accumulator registers.  The (step, batch) counter protocol of the persistent kernel
//            orders step t's reduction after step t-1's (gpu-scope fence before the counter increment), so it is the same
//            sequential f32 sum a register accumulator would give: identical bits on every path.
// pool_max : f32 running max; it travels like the cell state (the caller loads the previous value from L2 before the
//            accumulator is ready and this function stores the new one) -- 16 scalar red.max per thread and item put
//            ~6 us of L2 atomic traffic on the last layer's step chain.
// pool_last: f32, h at t == len-1
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// po: offset of the 4 units in the [row, out_pad] accumulator arrays; mprev: pool_max[po..po+3] (read when 0 < tg < len);
// tg: global timestep; len: valid length of the row
__device__ __forceinline__ void pool_accumulate4(float* pool_sum, float* pool_max, float* pool_last, long long po,
                                                 const float (&hn)[4], const float4& mprev, int tg, int len) {
  if (tg >= len) return;
  const float4 h4 = make_float4(hn[0], hn[1], hn[2], hn[3]);
  if (tg == 0) {
    __stcg(reinterpret_cast<float4*>(pool_sum + po), h4);
    __stcg(reinterpret_cast<float4*>(pool_max + po), h4);
  } else {
    red_add_v4(pool_sum + po, hn[0], hn[1], hn[2], hn[3]);
    __stcg(reinterpret_cast<float4*>(pool_max + po),
           make_float4(fmaxf(mprev.x, hn[0]), fmaxf(mprev.y, hn[1]), fmaxf(mprev.z, hn[2]), fmaxf(mprev.w, hn[3])));
  }
  if (tg == len - 1) __stcg(reinterpret_cast<float4*>(pool_last + po), h4);
}

// sum / last part only (the persistent kernel writes the running max of two chunks with one 256-bit store)
__device__ __forceinline__ void pool_sum_last4(float* pool_sum, float* pool_last, long long po, const float (&hn)[4], int tg,
                                               int len) {
  if (tg >= len) return;
  const float4 h4 = make_float4(hn[0], hn[1], hn[2], hn[3]);
  if (tg == 0) __stcg(reinterpret_cast<float4*>(pool_sum + po), h4);
  else red_add_v4(pool_sum + po, hn[0], hn[1], hn[2], hn[3]);
  if (tg == len - 1) __stcg(reinterpret_cast<float4*>(pool_last + po), h4);
}

// order-preserving u32 encoding of f32 (used by pr_curve.cu to sort scores as integers)
__device__ __forceinline__ uint32_t enc_max(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float dec_max(uint32_t e) {
  const uint32_t b = (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

// h_t in the ring: bf16 (hi); with `lo_off` > 0 also the bf16 residual h - hi at column offset lo_off (the split-bf16
// fp32-accurate mode: h ~ hi + lo to ~16 mantissa bits)
__device__ __forceinline__ void store_h4(__nv_bfloat16* yp, const float (&hn)[4], long long lo_off) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(hn[0], hn[1]);
  const __nv_bfloat162 b = __floats2bfloat162_rn(hn[2], hn[3]);
  *reinterpret_cast<uint2*>(yp) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  if (lo_off > 0) {
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    *reinterpret_cast<uint2*>(yp + lo_off) =
        make_uint2(pack_bf16x2(hn[0] - fa.x, hn[1] - fa.y), pack_bf16x2(hn[2] - fb.x, hn[3] - fb.y));
  }
}

}  // namespace ie
