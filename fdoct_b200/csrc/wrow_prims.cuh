// Warp-level primitives of the warp-per-A-scan kernel (wrow_kernel.cuh): lane identity, shuffles, warp reductions, cache-hinted
// global accesses, TMA L2 prefetch and the few global atomics of the scheduler.
//
// Every primitive has two bodies.  On the device it is the sm_100a instruction.  When the translation unit is compiled with
// ABC_WROW_HOST_EMU (tests/native/test_wrow_host.cu only - never the product library) the very same kernel body is a
// __host__ __device__ function and the primitives call into a small lock-step emulator (32 host threads per warp, a barrier per
// shuffle), so that the index algebra, table layouts, pairing logic and the scheduling protocol of the kernel can be checked on
// a machine without a GPU.  The emulator is test infrastructure; the product path has no CPU fallback.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#ifdef ABC_WROW_HOST_EMU
#define WROW_HD __host__ __device__ __forceinline__
namespace wemu {  // implemented by the test harness
int lane();
int warp_in_cta();
int cta();
int ncta();
int nwarps();
unsigned shfl_u32(unsigned v, int src);
void syncwarp();
void syncthreads();
int redux_min(int v);
int redux_max(int v);
int atomic_add(int* p, int v);
void atomic_min(int* p, int v);
void atomic_max(int* p, int v);
int load_acquire(const int* p);
void check_smem(const void* p, int bytes, int align);
void backoff();
unsigned ballot(int pred);
unsigned* tmem();  // per-CTA emulation of the tensor memory: [128 lanes][512 columns] 32-bit words
}  // namespace wemu
#else
#define WROW_HD __device__ __forceinline__
#endif

namespace abcoct {

#if defined(__CUDA_ARCH__) || !defined(ABC_WROW_HOST_EMU)
#define WROW_DEVICE_BODY 1
#else
#define WROW_DEVICE_BODY 0
#endif

WROW_HD int w_lane() {
#if WROW_DEVICE_BODY
  return threadIdx.x & 31;
#else
  return wemu::lane();
#endif
}
WROW_HD int w_warp_in_cta() {
#if WROW_DEVICE_BODY
  return threadIdx.x >> 5;
#else
  return wemu::warp_in_cta();
#endif
}
WROW_HD int w_cta() {
#if WROW_DEVICE_BODY
  return blockIdx.x;
#else
  return wemu::cta();
#endif
}
WROW_HD int w_ncta() {
#if WROW_DEVICE_BODY
  return gridDim.x;
#else
  return wemu::ncta();
#endif
}
WROW_HD void w_syncwarp() {
#if WROW_DEVICE_BODY
  __syncwarp();
#else
  wemu::syncwarp();
#endif
}
WROW_HD void w_syncthreads() {
#if WROW_DEVICE_BODY
  __syncthreads();
#else
  wemu::syncthreads();
#endif
}
WROW_HD float w_shfl(float v, int src) {
#if WROW_DEVICE_BODY
  return __shfl_sync(0xffffffffu, v, src);
#else
  unsigned u;
  memcpy(&u, &v, 4);
  u = wemu::shfl_u32(u, src);
  memcpy(&v, &u, 4);
  return v;
#endif
}
WROW_HD int w_shfl_i(int v, int src) {
#if WROW_DEVICE_BODY
  return __shfl_sync(0xffffffffu, v, src);
#else
  return (int)wemu::shfl_u32((unsigned)v, src);
#endif
}
WROW_HD unsigned w_ballot(bool pred) {
#if WROW_DEVICE_BODY
  return __ballot_sync(0xffffffffu, pred);
#else
  return wemu::ballot(pred ? 1 : 0);
#endif
}
WROW_HD float w_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += w_shfl(v, w_lane() ^ o);
  return v;
}
// warp-wide min / max of order-preserving integer encodings: one CREDUX each
WROW_HD int w_redux_min(int v) {
#if WROW_DEVICE_BODY
  return __reduce_min_sync(0xffffffffu, v);
#else
  return wemu::redux_min(v);
#endif
}
WROW_HD int w_redux_max(int v) {
#if WROW_DEVICE_BODY
  return __reduce_max_sync(0xffffffffu, v);
#else
  return wemu::redux_max(v);
#endif
}
WROW_HD float w_min3(float a, float b, float c) {
#if WROW_DEVICE_BODY
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return r;
#else
  return fminf(a, fminf(b, c));
#endif
}
WROW_HD float w_max3(float a, float b, float c) {
#if WROW_DEVICE_BODY
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
#else
  return fmaxf(a, fmaxf(b, c));
#endif
}

// ---- global memory with cache hints -------------------------------------------------------------------------------
// raw pixels: read exactly once (prefetched into L2 by the TMA unit): streaming load, evict-first in L1 and L2 (SASS LDG.E.EF)
WROW_HD uint4 w_ldg_stream16(const void* p) {
#if WROW_DEVICE_BODY
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#else
  uint4 v;
  memcpy(&v, p, 16);
  return v;
#endif
}
// calibration rows: L2-resident, re-read by every B-scan, no reuse inside an SM (no L1 allocation)
WROW_HD float4 w_ldg_cal16(const float* p) {
#if WROW_DEVICE_BODY
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  float4 v;
  memcpy(&v, p, 16);
  return v;
#endif
}
WROW_HD float4 w_ld_cg16(const float* p) {
#if WROW_DEVICE_BODY
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  float4 v;
  memcpy(&v, p, 16);
  return v;
#endif
}
WROW_HD int w_ld_cg_i(const int* p) {
#if WROW_DEVICE_BODY
  return __ldcg(p);
#else
  return wemu::load_acquire(p);
#endif
}
// dB scratch: written once, consumed from L2 a few microseconds later and then discarded
WROW_HD void w_st_keep(float* p, float v) {
#if WROW_DEVICE_BODY
  asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
#else
  *p = v;
#endif
}
WROW_HD void w_st_global_u8(uint8_t* p, unsigned v) {  // explicit state space: the caller may have lost it (non-inlined function)
#if WROW_DEVICE_BODY
  asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  *p = (uint8_t)v;
#endif
}
WROW_HD void w_st_stream_u32(void* p, unsigned v) {
#if WROW_DEVICE_BODY
  __stcs(reinterpret_cast<unsigned*>(p), v);
#else
  memcpy(p, &v, 4);
#endif
}
WROW_HD void w_st_stream_f4(float* p, float4 v) {
#if WROW_DEVICE_BODY
  __stcs(reinterpret_cast<float4*>(p), v);
#else
  memcpy(p, &v, 16);
#endif
}
WROW_HD void w_discard128(const void* p) {
#if WROW_DEVICE_BODY
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
#else
  (void)p;
#endif
}
// TMA unit: bulk prefetch of one pixel row into L2 (SASS UBLKPF.L2); bytes a multiple of 16
WROW_HD void w_prefetch_l2(const void* p, unsigned bytes) {
#if WROW_DEVICE_BODY
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
  (void)p;
  (void)bytes;
#endif
}

// ---- tensor memory (TMEM) as a row store -----------------------------------------------------------------------------
// The kernel issues no MMA, so the 256 KB of tensor memory of the SM (128 lanes x 512 columns x 32 bit) are free: a warp parks
// the 32 dB values that each lane holds for a finished A-scan in 32 columns of its own lane quarter (tcgen05.st, shape 32x32b:
// thread i of warp w reaches lane 32 (w % 4) + i) and reads them back into the same registers when the B-scan is complete.
// Address = lane << 16 | column.  Allocation: the whole 512 columns, by warp 0 of the (only) CTA of the SM.
WROW_HD unsigned w_tmem_alloc512(unsigned* smem_word) {  // all threads of the CTA
#if WROW_DEVICE_BODY
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(smem_word)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *reinterpret_cast<volatile unsigned*>(smem_word);
#else
  (void)smem_word;
  wemu::syncthreads();
  return 0u;
#endif
}
WROW_HD void w_tmem_free512(unsigned base) {  // all threads of the CTA, after their last access
#if WROW_DEVICE_BODY
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
#else
  (void)base;
  wemu::syncthreads();
#endif
}
WROW_HD void w_tmem_st16(unsigned taddr, const float (&v)[16]) {  // whole warp; 16 consecutive columns of this thread's lane
#if WROW_DEVICE_BODY
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
      "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])),
      "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
      "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
#else
  unsigned* t = wemu::tmem() + (size_t)(((taddr >> 16) & 127u) + (unsigned)wemu::lane()) * 512u + (taddr & 511u);
  for (int i = 0; i < 16; ++i) memcpy(t + i, &v[i], 4);
#endif
}
WROW_HD void w_tmem_st_wait() {
#if WROW_DEVICE_BODY
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
}
WROW_HD void w_tmem_ld16(unsigned taddr, float (&v)[16]) {  // whole warp; returns when the values are in the registers
#if WROW_DEVICE_BODY
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"  // same statement: the registers are defined only after the wait
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
#else
  const unsigned* t = wemu::tmem() + (size_t)(((taddr >> 16) & 127u) + (unsigned)wemu::lane()) * 512u + (taddr & 511u);
  for (int i = 0; i < 16; ++i) memcpy(&v[i], t + i, 4);
#endif
}
WROW_HD void w_st_shared_u8(unsigned char* p, unsigned v) {
  *p = (unsigned char)v;
}

// ---- L2 cache policies (createpolicy): which = 0 evict_normal, 1 evict_first, 2 evict_last -----------------------------
WROW_HD unsigned long long w_policy(int which) {
#if WROW_DEVICE_BODY
  unsigned long long pn, pf, pl;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pn));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
  return which == 1 ? pf : (which == 2 ? pl : pn);
#else
  return (unsigned long long)which;
#endif
}
WROW_HD void w_prefetch_l2_pol(const void* p, unsigned bytes, unsigned long long pol) {
#if WROW_DEVICE_BODY
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p), "r"(bytes), "l"(pol) : "memory");
#else
  (void)p;
  (void)bytes;
  (void)pol;
#endif
}
WROW_HD void w_st_keep_pol(float* p, float v, unsigned long long pol) {
#if WROW_DEVICE_BODY
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
#else
  (void)pol;
  *p = v;
#endif
}
WROW_HD uint4 w_ldg_stream16_pol(const void* p, unsigned long long pol) {
#if WROW_DEVICE_BODY
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
#else
  (void)pol;
  uint4 v;
  memcpy(&v, p, 16);
  return v;
#endif
}

// ---- TMA bulk copies global -> shared with mbarrier completion (SASS UBLKCP + SYNCS) --------------------------------
// One mbarrier per warp; lane 0 arms it with the byte count and issues the copies, every lane waits on the phase parity.
WROW_HD void w_mbar_init(unsigned long long* bar) {
#if WROW_DEVICE_BODY
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
  *bar = 0;
#endif
}
WROW_HD void w_tma_arm(unsigned long long* bar, unsigned bytes) {  // also orders this warp's earlier generic accesses of the
#if WROW_DEVICE_BODY                                               // destination before the async-proxy writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
#else
  (void)bar;
  (void)bytes;
#endif
}
WROW_HD void w_tma_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
#if WROW_DEVICE_BODY
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
#else
  (void)bar;
  memcpy(dst, src, bytes);  // lock-step emulation: complete at issue; w_mbar_wait is a warp barrier
#endif
}
WROW_HD void w_mbar_wait(unsigned long long* bar, unsigned parity) {
#if WROW_DEVICE_BODY
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
        : "memory");
  } while (!ok);
#else
  (void)bar;
  (void)parity;
  wemu::syncwarp();
#endif
}

// ---- scheduler atomics --------------------------------------------------------------------------------------------
WROW_HD int w_atomic_add(int* p, int v) {
#if WROW_DEVICE_BODY
  // plain PTX: the compiler's warp-aggregation of atomicAdd (vote + leader + broadcast) would wait for the result on the
  // spot, and these tickets are deliberately consumed a row later
  int r;
  asm volatile("atom.relaxed.gpu.global.add.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
  return r;
#else
  return wemu::atomic_add(p, v);
#endif
}
WROW_HD void w_atomic_min(int* p, int v) {
#if WROW_DEVICE_BODY
  atomicMin(p, v);
#else
  wemu::atomic_min(p, v);
#endif
}
WROW_HD void w_atomic_max(int* p, int v) {
#if WROW_DEVICE_BODY
  atomicMax(p, v);
#else
  wemu::atomic_max(p, v);
#endif
}
// publish: everything this WARP wrote before the preceding w_syncwarp() becomes visible before the count (cumulative release)
WROW_HD void w_release_add(int* p, int v) {
#if WROW_DEVICE_BODY
  // release at gpu scope (fence.acq_rel, lighter than the fence.sc of __threadfence) + RED
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  wemu::atomic_add(p, v);
#endif
}
WROW_HD int w_ld_acquire(const int* p) {
#if WROW_DEVICE_BODY
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#else
  return wemu::load_acquire(p);
#endif
}
// cta-scope release / acquire on SHARED-memory words: the mailboxes between the worker warps and the service warp of a CTA
WROW_HD void w_st_release_cta(int* p, int v) {
#if WROW_DEVICE_BODY
  asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#else
  __atomic_store_n(p, v, __ATOMIC_SEQ_CST);
#endif
}
WROW_HD int w_ld_acquire_cta(const int* p) {
#if WROW_DEVICE_BODY
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
#else
  return __atomic_load_n(p, __ATOMIC_SEQ_CST);
#endif
}
// gpu-scope fence of the service warp: cumulative, i.e. it also orders the worker warps' stores that were observed through
// the cta-scope mailboxes before the counts that follow it
WROW_HD void w_fence_gpu() {
#if WROW_DEVICE_BODY
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
WROW_HD int w_ld_relaxed(const int* p) {
#if WROW_DEVICE_BODY
  return *reinterpret_cast<const volatile int*>(p);
#else
  return wemu::load_acquire(p);
#endif
}
WROW_HD void w_acquire_fence() {
#if WROW_DEVICE_BODY
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
WROW_HD void w_backoff() {
#if WROW_DEVICE_BODY
  __nanosleep(1000);  // pollers must not crowd the L2 lines that the workers' tickets and counts live in
#else
  wemu::backoff();
#endif
}
// ---- primitives of the resident-row kernel (wres_kernel.cuh) ---------------------------------------------------------
// tables that do not fit the shared memory of a plan: read through L1 (the same 16 KB for every warp of the SM)
WROW_HD float4 w_ldg_tbl16(const void* p) {
#if WROW_DEVICE_BODY
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  float4 v;
  memcpy(&v, p, 16);
  return v;
#endif
}
// display image, 4 A-scans of one depth bin: a partial sector - default policy, so that L2 merges the eight writers of a sector
WROW_HD void w_st_global_u32(void* p, unsigned v) {
#if WROW_DEVICE_BODY
  asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  memcpy(p, &v, 4);
#endif
}
WROW_HD void w_st_global_f4(float* p, float4 v) {
#if WROW_DEVICE_BODY
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#else
  memcpy(p, &v, 16);
#endif
}
// team counters in shared memory: add with cta-scope release (the rows / quarters counted are visible to whoever acquires)
WROW_HD void w_red_release_cta(int* p, int v) {
#if WROW_DEVICE_BODY
  asm volatile("red.release.cta.shared::cta.add.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#else
  __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
#endif
}
WROW_HD void w_red_relaxed(int* p, int v) {
#if WROW_DEVICE_BODY
  asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  wemu::atomic_add(p, v);
#endif
}
// shared-memory words of the per-CTA completion frontier (wres_kernel.cuh)
WROW_HD unsigned w_cas_smem(unsigned* p, unsigned expect, unsigned desired) {
#if WROW_DEVICE_BODY
  unsigned old;
  asm volatile("atom.relaxed.cta.shared::cta.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(expect), "r"(desired) : "memory");
  return old;
#else
  __atomic_compare_exchange_n(p, &expect, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return expect;
#endif
}
WROW_HD void w_max_release_smem(int* p, int v) {
#if WROW_DEVICE_BODY
  asm volatile("red.release.cta.shared::cta.max.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#else
  int cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > cur && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
#endif
}
WROW_HD unsigned w_now_ns32() {
#if WROW_DEVICE_BODY
  unsigned t;
  asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(t));
  return t;
#else
  static unsigned fake = 0;
  return __atomic_add_fetch(&fake, 1000u, __ATOMIC_SEQ_CST);
#endif
}
WROW_HD void w_sleep_ns(unsigned ns) {
#if WROW_DEVICE_BODY
  while (ns > 0u) {  // nanosleep takes at most about a millisecond; the argument here is a few microseconds
    const unsigned step = ns > 500000u ? 500000u : ns;
    __nanosleep(step);
    ns -= step;
  }
#else
  (void)ns;
#endif
}
WROW_HD void w_backoff_short() {
#if WROW_DEVICE_BODY
  __nanosleep(200);
#else
  wemu::backoff();
#endif
}
WROW_HD unsigned long long w_now_ns() {
#if WROW_DEVICE_BODY
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
#else
  return 0;
#endif
}
WROW_HD float w_inf(bool negative) {
  const unsigned u = negative ? 0xff800000u : 0x7f800000u;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
WROW_HD void w_trap() {
#if WROW_DEVICE_BODY
  __trap();
#endif
}
WROW_HD unsigned w_byte_perm(unsigned a, unsigned b, unsigned s) {
#if WROW_DEVICE_BODY
  return __byte_perm(a, b, s);
#else
  unsigned long long t = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((t >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
#endif
}
WROW_HD void w_check_smem(const void* p, int bytes, int align) {
#if !WROW_DEVICE_BODY
  wemu::check_smem(p, bytes, align);
#else
  (void)p;
  (void)bytes;
  (void)align;
#endif
}

}  // namespace abcoct
