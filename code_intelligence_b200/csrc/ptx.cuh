// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA + TMEM).
// Everything the kernels in this directory need and nothing else.  No CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ie {

// L2 cache-policy immediates accepted by cp.async.bulk.tensor ... .L2::cache_hint
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst  = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast   = 0x14F0000000000000ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// make generic-proxy writes visible to the async proxy (TMA / UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// ----------------------------------------------------------------------------------------------
// Abort protocol of every spin in this directory.  A wait that exceeds `limit` SM cycles (a protocol bug, or a
// persistent grid that lost co-residency) does NOT trap -- a trap poisons the CUDA context of every handle in the
// process, and include/issue_emb_b200.h promises error codes.  Instead the waiter raises the CTA's shared-memory flag
// and the launch's global flag (the handle's error word) and returns; from then on every wait of the CTA returns at
// once ("drain mode": role loops fall through, results are garbage, the kernel terminates), other CTAs adopt the
// global flag the next time one of their spins polls it, and the host turns the flag into IE_ERR_CUDA.
// ----------------------------------------------------------------------------------------------
struct Abort {
  uint32_t* s;      // this CTA's flag in shared memory (initialised to 0 by the thread that initialises the barriers)
  unsigned* g;      // the launch's flag in global memory (nullptr: none)
  long long limit;  // SM cycles a single wait may take
};
constexpr long long kSpinLimitDefault = 4000000000ll;  // ~2 s

__device__ __forceinline__ bool aborted(const Abort& a) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(a.s)) : "memory");
  return v != 0;
}
static __device__ __noinline__ void abort_raise(const Abort& a) {
  asm volatile("st.volatile.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(a.s)), "r"(1u) : "memory");
  if (a.g != nullptr) atomicExch(a.g, 1u);
}
// slow-path poll (every few hundred spins): true once this CTA is in drain mode
static __device__ __noinline__ bool abort_poll(const Abort& a, long long t0) {
  if (aborted(a)) return true;
  bool hit = (clock64() - t0) > a.limit;
  if (!hit && a.g != nullptr) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.g) : "memory");
    hit = v != 0;
  }
  if (hit) abort_raise(a);
  return hit;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, const Abort& ab) {
  if (mbar_try_wait(bar, parity)) return;
  if (aborted(ab)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0xFFu) == 0 && abort_poll(ab, t0)) return;
  }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier (complete_tx::bytes). c0 = inner coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                            int32_t c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

// multicast variant: the box is written to the same CTA-relative smem offset of every CTA in `mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                               int32_t c1, uint16_t mask, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%4, %5}], [%2], %3, %6;"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

// 2-D tiled store shared -> global (bulk async group); out-of-range parts of the box are clipped by the tensor map
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// same with an L2 cache-policy hint (evict-first: an output stream must not displace the operand panels in L2)
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1,
                                                  uint64_t hint) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(hint)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source / are incomplete
template <int N> __device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ----------------------------------------------------------------------------------------------
// thread-block clusters
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs in the cluster (also orders shared-memory accesses like __syncthreads)
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, UMMA issue, commit, TMEM loads
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; ncols power of two >= 32; the TMEM base address is written to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Shared-memory matrix descriptor: K-major operand tile, 128-byte swizzle, rows 128 B apart,
// 8-row core-matrix groups 1024 B apart (what a TMA box {64 bf16, rows} with SWIZZLE_128B writes).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // [0,14)  start address >> 4
  // [16,30) leading byte offset: unused for swizzled K-major
  d |= static_cast<uint64_t>(1024u >> 4) << 32;             // [32,46) stride byte offset >> 4
  d |= 1ull << 46;                                          // [46,48) descriptor version (sm_100)
  d |= 2ull << 61;                                          // [61,64) SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: A,B = bf16 K-major, D = f32, shape M x N (K = 16 per instruction).
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
  uint32_t d = 0;
  d |= 1u << 4;          // D format f32
  d |= 1u << 7;          // A format bf16
  d |= 1u << 10;         // B format bf16
  d |= (N >> 3) << 17;   // N / 8
  d |= (M >> 4) << 24;   // M / 16
  return d;
}
// D[tmem] (+)= A[smem] * B[smem]^T ; single issuing thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued UMMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// same, arriving on the barrier at the same offset in every CTA of `mask` (operand stages filled by multicast TMA
// may only be overwritten once every CTA of the cluster has consumed them)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: two CTAs of a cluster on one TPC act as one M=256 tensor core.
// Barriers that gate the MMA live in the even ("leader") CTA; shared-window addresses carry the pair rank in
// bit 24, so clearing it redirects a barrier operand to the leader's copy of the same smem offset.
// ----------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

// 2-D tiled load into THIS CTA's smem; the complete_tx goes to the barrier at the same offset in the leader CTA
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                                 int32_t c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// multicast variant of tma_load_2d_pair: the box lands at the same CTA-relative offset in every CTA of `mask`; with
// cta_group::2 the complete_tx of each destination goes to the barrier (same offset) of the CTA of ITS pair whose rank has
// the parity of the CTA `bar` points to -- `bar` is redirected to this pair's leader, so every destination pair's
// leader is signalled
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                                    int32_t c1, uint16_t mask, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      ".L2::cache_hint [%0], [%1, {%4, %5}], [%2], %3, %6;"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & kPeerBitMask), "h"(mask), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// plain arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256: rows 0..127 of A and columns 0..N/2-1 of B come from the leader's
// smem, the other halves from the peer's smem at the same offsets.  Issued by one thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of the pair's MMAs, arriving on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// grid-scope flags in global memory (persistent kernels)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bounded spin until *p >= target (abort protocol above instead of hanging the GPU on a protocol bug)
__device__ __forceinline__ void wait_flag_ge(const unsigned* p, unsigned target, const Abort& ab) {
  if (ld_acquire(p) >= target) return;
  if (aborted(ab)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (ld_acquire(p) < target) {
    if (((++spins) & 0x3Fu) == 0 && abort_poll(ab, t0)) return;
  }
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// TMEM -> registers: the warp's 32 lanes x 16 consecutive 32-bit columns (thread i <- lane base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// tanh(x) = 1 - 2/(exp(2x)+1): abs error ~1e-7, saturates correctly for |x| large
__device__ __forceinline__ float tanh_acc(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

// single-MUFU variants (tanh.approx.f32, max relative error 2^-11): the epilogue of the recurrent kernels is bound by
// the 16/clk/SM special-function unit (10 MUFU per cell with the accurate forms, 5 with these)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// 256-bit global store (sm_100): one full 32-byte sector per instruction, so the L2 never sees a partial-sector
// write (the 128-bit stores of the first GEMM epilogue caused read-modify-write fills: DRAM reads ~ output bytes)
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e,
                                             uint32_t f, uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d),
               "r"(e), "r"(f), "r"(g), "r"(h)
               : "memory");
}

// 256-bit store that stays out of L1 (state another SM will read from L2 with ld.global.cg)
__device__ __forceinline__ void st_global_cg_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e,
                                                uint32_t f, uint32_t g, uint32_t h) {
  asm volatile("st.global.cg.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d),
               "r"(e), "r"(f), "r"(g), "r"(h)
               : "memory");
}

// streaming 256-bit read-only load (sm_100): no L1 allocation, evict-first in L2 -- a once-read stream must not
// displace the L2-resident weights.  (The .L2::evict_first qualifier only exists for the 256-bit forms.)
__device__ __forceinline__ void ldg_stream8(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

// 256-bit coherent load that bypasses L1 (state written by another SM of the same launch: c, running max).  One full
// 32-byte sector per request -- two 128-bit ld.cg of the same sector are two requests and move the sector twice
// (ncu: 6.4 GB of the recurrent kernel's 25.7 GB of LSU reads per launch were the second halves of c sectors).
__device__ __forceinline__ void ldg_cg8(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p)
               : "memory");
}

// same load into eight 32-bit registers (f32 bits or packed bf16 pairs)
__device__ __forceinline__ void ldg_stream8_b32(const void* p, uint32_t* d) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
               : "l"(p));
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bounded spin with relaxed loads, one acquire fence at the end
__device__ __forceinline__ void wait_flag_ge_relaxed(const unsigned* p, unsigned target, const Abort& ab) {
  if (ld_relaxed(p) < target && !aborted(ab)) {
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (ld_relaxed(p) < target) {
      if (((++spins) & 0x3Fu) == 0 && abort_poll(ab, t0)) break;
    }
  }
  __threadfence();
}
__device__ __forceinline__ void red_relaxed_add(unsigned* p, unsigned v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace ie
