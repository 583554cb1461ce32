// Micro-benchmark (debug hook ie_debug_umma_rate): cycles per tcgen05.mma for the operand shapes the kernels use,
// operands resident in shared memory (no TMA in the loop).  Separates "issue / operand-fetch bound" from
// "tensor-pipe bound" when reading the profiles of the recurrent kernels.
#include "kernels.h"
#include "ptx.cuh"

namespace ie {
namespace {

// mode 0: cta_group::1, M=128.  One thread issues `iters` groups of 4 MMAs (K = 4 x 16) and one commit per group.
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int n, int iters, int commit_every, int ntiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  // zero the operand area: ntiles A tiles (16 KB apart) followed by ntiles B tiles (32 KB apart)
  for (int i = threadIdx.x; i < (ntiles * 49152) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tslot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint64_t da = umma_desc_sw128(smem_u32(smem));
    const uint64_t db = umma_desc_sw128(smem_u32(smem + ntiles * 16384));
    uint32_t phase = 0;
    // warm
    for (int k = 0; k < 4; ++k) umma_bf16(tm, da + 2 * k, db + 2 * k, idesc, 1);
    umma_commit(&bar);
    mbar_wait(&bar, phase); phase ^= 1;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // ntiles > 1: every group reads a different A / B tile (streaming operands), tiles 16 KB apart
      const uint64_t off = static_cast<uint64_t>((i % ntiles) * (16384 >> 4));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tm, da + off + 2 * k, db + 2 * off + 2 * k, idesc, 1);
      if ((i + 1) % commit_every == 0 && i + 1 < iters) umma_commit(&bar), mbar_wait(&bar, phase), phase ^= 1;
    }
    const long long t1 = clock64();   // issue done
    umma_commit(&bar);
    mbar_wait(&bar, phase);
    const long long t2 = clock64();   // execution done
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// mode 1: cta_group::2, M=256, N=n (each CTA holds n/2 rows of B)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma_rate_pair_kernel(int n, int iters, int ntiles, int mimic, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t done_bar;   // a barrier whose phase 0 has completed: waiting on it returns at once
  __shared__ uint64_t sink_bar;   // receives commits nobody waits for
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < (ntiles * 49152) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&done_bar, 1);
    mbar_init(&sink_bar, 1);
    fence_barrier_init();
    mbar_arrive(&done_bar);
  }
  fence_proxy_async();
  cluster_sync();
  if (threadIdx.x < 32) tmem_alloc_pair(&tslot, 512);
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tm = tslot;
  // bit5 of mimic: a second issuing thread (warp 1) runs the same loop on its own accumulator and barrier
  __shared__ uint64_t bar2;
  if (threadIdx.x == 0) { mbar_init(&bar2, 1); fence_barrier_init(); }
  cluster_sync();
  const bool issuer0 = cluster_ctarank() == 0 && threadIdx.x == 0;
  const bool issuer1 = false;  // a second issuing thread faults on sm_100a (tried in round 1): left disabled
  if (issuer0 || issuer1) {
    uint64_t* mybar = issuer0 ? &bar : &bar2;
    const uint32_t idesc = umma_idesc_bf16(256, n);
    const uint64_t da = umma_desc_sw128(smem_u32(smem));
    const uint64_t db = umma_desc_sw128(smem_u32(smem + ntiles * 16384));
    const uint32_t tacc = issuer0 ? tm : tm + 256;
    for (int k = 0; k < 4; ++k) umma_bf16_pair(tacc, da + 2 * k, db + 2 * k, idesc, 1);
    umma_commit_pair_mc(mybar, 0x3);
    mbar_wait(mybar, 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t off = static_cast<uint64_t>((i % ntiles) * (16384 >> 4));
      if (mimic & 1) { mbar_wait(&done_bar, 0); mbar_wait(&done_bar, 0); }
      if (mimic & 2) tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // bit4: two independent accumulators, alternating per MMA (two interleaved dependency chains)
        const uint32_t td = (mimic & 16) ? tacc + static_cast<uint32_t>(((i * 4 + k) & 1) * 128) : tacc;
        umma_bf16_pair(td, da + off + 2 * k, db + 2 * off + 2 * k, idesc, (i | k) != 0);
      }
      if (mimic & 4) umma_commit_pair_mc(&sink_bar, 0x3);
      if (mimic & 8) umma_commit_pair_mc(&sink_bar, 0x3);
    }
    const long long t1 = clock64();
    umma_commit_pair_mc(mybar, 0x3);
    mbar_wait(mybar, 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && issuer0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (threadIdx.x == 0) {
    mbar_wait(&bar, 0);
    mbar_wait(&bar, 1);
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc_pair(tm, 512); }
}

}  // namespace

// out[0] = cycles to ISSUE iters*4 MMAs, out[1] = cycles until they have all executed
cudaError_t run_umma_rate(int mode, int n, int iters, int commit_every, int grid, int ntiles, long long* host_out) {
  long long* d = nullptr;
  cudaError_t e = cudaMalloc(&d, 16);
  if (e != cudaSuccess) return e;
  cudaMemset(d, 0, 16);
  if (ntiles < 1) ntiles = 1;
  if (ntiles > 4) ntiles = 4;
  const size_t smem = 1024 + static_cast<size_t>(ntiles) * 49152 + 64;
  if (mode == 0) {
    cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    umma_rate_kernel<<<grid, 128, smem>>>(n, iters, commit_every > 0 ? commit_every : iters, ntiles, d);
  } else {
    cudaFuncSetAttribute(umma_rate_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    umma_rate_pair_kernel<<<grid * 2, 128, smem>>>(n, iters, ntiles, commit_every, d);
  }
  e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(host_out, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e;
}

}  // namespace ie
