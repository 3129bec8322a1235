// gemm_tcgen05.cu -- C = epilogue([A1|A2] . [B1;B2]) on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Same contract as gemm_tf32x3.cu (fp32 in / fp32 out, fp32-level accuracy through 3xTF32 split
// accumulation), but on the Blackwell-native path: the legacy mma.sync TF32 pipe tops out at
// ~240 TFLOP/s on B200 (profiles/r01_gemm_mma_sync.txt), i.e. 3xTF32 there is no faster than FFMA.
//
//   * accumulator: 128 x BN fp32 tile in TMEM (tcgen05.alloc), BN in {32..256} chosen per problem so the
//     grid is ~one wave of 148 CTAs;
//   * operands: K-major SWIZZLE_128B shared-memory images, one 128-byte row = 32 fp32 of K;
//       B (weights): split into TF32 hi/lo and pre-swizzled ONCE per call by pack_b_kernel into the exact
//         shared-memory image, then streamed with 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) that
//         complete on the stage's mbarrier;
//       A (node states): 4 producer warps load rows with coalesced 128-bit loads, split hi/lo in
//         registers and store both images swizzled (conflict-free), then fence.proxy.async + arrive;
//   * MMA: one elected thread issues, per 32-wide K chunk, 4 k-steps x {lo*hi, hi*lo, hi*hi}
//     tcgen05.mma.kind::tf32 into the same TMEM accumulator, and tcgen05.commit releases the stage;
//   * epilogue: the 4 producer warps read their 32 TMEM lanes with tcgen05.ld.32x32b.x16, apply
//     bias / activation / GRU gate math and store rows.
#include "gemm.cuh"

#include <stdlib.h>
#include <vector>
#include <mutex>
#include <algorithm>

namespace rgnn {

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                    // fp32 elements per 128-byte swizzle row
constexpr int TC_GROUPS = 2;                 // producer groups alternate K chunks (see the producer loop)
constexpr int TC_GROUP_WARPS = 4;            // one group covers the 128-row tile: 4 warps x 32 rows of TMEM lanes
constexpr int TC_PRODUCER_WARPS = TC_GROUPS * TC_GROUP_WARPS;
constexpr int TC_GROUP_THREADS = 32 * TC_GROUP_WARPS;
constexpr int TC_THREADS = 32 * (TC_PRODUCER_WARPS + 2);   // + MMA warp + B-loader warp
constexpr int A_IMG_BYTES = TC_BM * 128;     // 16 KB
constexpr size_t TC_SMEM_BUDGET = 200 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Explicit shared-space accesses on 32-bit shared addresses.  (Going through a generic pointer that was
// rounded up via uintptr_t made ptxas emit generic ST.E.128 + MEMBAR.ALL.CTA in front of the proxy fence,
// which waited on the prefetched global loads and serialised the whole pipeline -- r01 trace.)
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must fault the kernel (trap), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start>>4 | LBO | SBO=1024B | version 1 | layout 2.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // bits [0,14)  start address >> 4
  d |= (uint64_t)1 << 16;                                // bits [16,30) leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                      // bits [32,46) stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                                // bits [46,48) descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                                // bits [61,64) SWIZZLE_128B
  return d;
}

// kind::tf32 instruction descriptor: D=F32, A=B=TF32, both K-major, N>>3 at bit 17, M>>4 at bit 24.
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);   // exactly representable in TF32
  lo = x - hi;                                              // exact in fp32; the tensor core keeps its top 11 bits
}

// -------------------------------------------------------------------------------------------------
// pack_b: [K, N] fp32 weights -> per (n-tile, k-chunk) hi / lo shared-memory images (BN rows x 128 B,
// K-major, 128-byte swizzled).  Column blocks may come from different matrices (per-type kernels).
// -------------------------------------------------------------------------------------------------
struct PackParams {
  const float* b1[RGNN_MAX_EDGE_TYPES];   // column block j of segment 1: rows [0, K1)
  const float* b2[RGNN_MAX_EDGE_TYPES];   // column block j of segment 2: rows [0, K2) (may be null when K2 == 0)
  int ldb1, ldb2;
  int K1, K2;
  int block_cols;                         // width of one column block (N for a plain GEMM)
  int n_total;                            // total columns = blocks * block_cols
  int BN;
  int chunks1, chunks2;                   // ceil(K1/32), ceil(K2/32)
  float* out;                             // [n_tiles][chunks1+chunks2][2][BN*32]
};

// grid = (k-chunks, n-tiles, BN/32 row slices... ceil), block = 256 threads = 32 image rows x 8 16-byte chunks
__global__ void __launch_bounds__(256) pack_b_kernel(const __grid_constant__ PackParams p) {
  const int chunk = blockIdx.x;           // k-chunk over both segments
  const int tile = blockIdx.y;            // n-tile
  const int nchunks = p.chunks1 + p.chunks2;
  const bool seg2 = chunk >= p.chunks1;
  const int k0 = (seg2 ? chunk - p.chunks1 : chunk) * TC_BK;
  const int Kseg = seg2 ? p.K2 : p.K1;
  const int ld = seg2 ? p.ldb2 : p.ldb1;
  float* img_hi = p.out + ((size_t)tile * nchunks + chunk) * 2 * (p.BN * TC_BK);
  float* img_lo = img_hi + p.BN * TC_BK;
  const int nl = blockIdx.z * 32 + (threadIdx.x & 31);   // consecutive threads -> consecutive n: coalesced reads per k
  const int c16 = threadIdx.x >> 5;                      // 16-byte chunk (4 k values) inside the 128-byte row
  if (nl >= p.BN) return;
  const int n = tile * p.BN + nl;
  float hi[4], lo[4];
  float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (n < p.n_total) {
    const int blk = n / p.block_cols, col = n - blk * p.block_cols;
    const float* src = (seg2 ? p.b2[blk] : p.b1[blk]) + col;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + c16 * 4 + j;
      if (k < Kseg) x[j] = __ldg(src + (size_t)k * ld);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) split_tf32(x[j], hi[j], lo[j]);
  const int off = (nl >> 3) * 256 + (nl & 7) * 32 + ((c16 ^ (nl & 7)) << 2);   // float index inside the image
  *reinterpret_cast<float4*>(img_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<float4*>(img_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// -------------------------------------------------------------------------------------------------
struct TcParams {
  GemmParams g;
  const float* packed;     // pack_b output (per z for ROW_RANGES / COL_BLOCKS: z * packed_stride floats)
  size_t packed_stride;
  int BN, stages, chunks1, chunks2;
  int n_total;             // columns of C covered by this launch (batch*N for SHARED_A, else N)
  int tmem_cols;
  int ring_bytes;          // operand ring, at least as large as the epilogue's [128][BN+4] staging tile
  long long* trace;        // RGNN_GEMM_TRACE=1: per-CTA clock64 timeline (debug only), else nullptr
};

// trace slots (clock64 relative to kernel entry): 0 prologue done; 1+3c producer got stage; 2+3c loads landed;
// 3+3c arrived (c < 8); 32+2c MMA saw full; 33+2c MMA committed (c < 8); 56 epilogue start; 57 epilogue end
#define TC_TRACE(slot)                                                                          \
  do {                                                                                          \
    if (p.trace != nullptr) p.trace[(size_t)cta_lin * 64 + (slot)] = clock64() - t_entry;       \
  } while (0)

__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const long long t_entry = clock64();
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const GemmParams& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN, S = p.stages;
  const int b_img_bytes = BN * 128;
  const int stage_bytes = 2 * A_IMG_BYTES + 2 * b_img_bytes;
  // 1024-byte aligned operand ring (32-bit shared addresses), then barriers
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t ring = (smem_base + 1023u) & ~1023u;
  const uint32_t full0 = ring + (uint32_t)p.ring_bytes, empty0 = full0 + 8 * S, accum_bar = full0 + 16 * S;
  const uint32_t tmem_slot = accum_bar + 8;

  // ---- operands of this CTA ----
  const int z = blockIdx.z;
  const float* A1 = g.A1;
  float* C = g.C;
  int row_begin = 0, row_end = g.M;
  const float* packed = p.packed;
  if (g.batch_mode == BATCH_ROW_RANGES) {
    row_begin = g.row_off[z]; row_end = g.row_off[z + 1];
    packed += (size_t)z * p.packed_stride;
  } else if (g.batch_mode == BATCH_COL_BLOCKS) {
    A1 += (size_t)z * g.K1;
    C += (size_t)z * g.N;
    packed += (size_t)z * p.packed_stride;
  }
  const int m0 = row_begin + blockIdx.y * TC_BM;
  if (m0 >= row_end) return;                       // uniform across the CTA, before any barrier / allocation
  const int n_tile = blockIdx.x, n0 = n_tile * BN;
  const int nchunks = p.chunks1 + p.chunks2;

  // A-producer helper: the 8 float4 of chunk c this thread owns (rows ptid/8 + 16 i, 16-byte column ptid%8)
  const int group = warp / TC_GROUP_WARPS;          // producer group (meaningful for warps < TC_PRODUCER_WARPS)
  const int ptid = tid % TC_GROUP_THREADS;          // thread index inside the group
  auto load_a_chunk = [&](int c, float4 (&v)[8]) {
    const bool seg2 = c >= p.chunks1;
    const int k0 = (seg2 ? c - p.chunks1 : c) * TC_BK;
    const int Kseg = seg2 ? g.K2 : g.K1;
    const float* Abase = seg2 ? g.A2 : A1;
    const int lda = seg2 ? g.lda2 : g.lda1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = ptid + i * TC_GROUP_THREADS;
      const int row = f >> 3, c16 = f & 7;
      const int grow = m0 + row, gk = k0 + c16 * 4;
      v[i] = (grow < row_end && gk < Kseg) ? __ldg(reinterpret_cast<const float4*>(Abase + (size_t)grow * lda + gk))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float4 va[8];
  if (warp < TC_PRODUCER_WARPS && group < nchunks) load_a_chunk(group, va);   // in flight while barriers / TMEM are set up

  if (warp == TC_PRODUCER_WARPS && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, TC_GROUP_THREADS + 1);         // 128 producer arrivals + the B loader's expect_tx arrival
      mbar_init(empty0 + 8 * s, 1);                           // one tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == TC_PRODUCER_WARPS) {                  // TMEM allocation by the MMA warp (it also frees it)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = lds32(tmem_slot);
  if (tid == 0) TC_TRACE(0);

  if (warp < TC_PRODUCER_WARPS) {
    // =========================== A producers ===========================
    // Two producer groups alternate chunks (group g owns chunks g, g+2, ...).  A thread issues the loads of
    // its NEXT chunk right after publishing the current one, so they fly while the other group converts:
    // fence.proxy.async lowers to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC and the membar waits for the issuing
    // thread's outstanding loads -- prefetching inside one group would be serialised by it (r01 trace).
    for (int c = group; c < nchunks; c += TC_GROUPS) {
      const int s = c % S;
      const uint32_t use = c / S;
      if (c >= S) mbar_wait(empty0 + 8 * s, (use - 1) & 1);
      if (ptid == 0 && c < 8) TC_TRACE(1 + 3 * c);
      const uint32_t a_hi = ring + (uint32_t)(s * stage_bytes);
      const uint32_t a_lo = a_hi + A_IMG_BYTES;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int f = ptid + i * TC_GROUP_THREADS;
        const int row = f >> 3, c16 = f & 7;
        float4 hi, lo;
        split_tf32(va[i].x, hi.x, lo.x); split_tf32(va[i].y, hi.y, lo.y);
        split_tf32(va[i].z, hi.z, lo.z); split_tf32(va[i].w, hi.w, lo.w);
        const uint32_t off = (uint32_t)(row * 128 + ((c16 ^ (row & 7)) << 4));
        sts128(a_hi + off, hi);
        sts128(a_lo + off, lo);
      }
      fence_proxy_async_smem();                     // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(full0 + 8 * s);
      if (ptid == 0 && c < 8) TC_TRACE(3 + 3 * c);
      if (c + TC_GROUPS < nchunks) load_a_chunk(c + TC_GROUPS, va);
    }
    // =========================== epilogue ===========================
    // TMEM -> registers (thread = row) -> padded smem tile -> row-contiguous reads: bias / activation /
    // GRU math and 128-bit stores are coalesced along the row (the thread-per-row stores of v1 cost 6000
    // cycles per tile).  The operand ring is free: accum_bar fires after the last MMA has read it.
    mbar_wait(accum_bar, 0);
    if (tid == 0) TC_TRACE(56);
    __syncwarp();                                   // tcgen05.ld is warp-collective (.sync.aligned)
    tc_fence_after_sync();
    const uint32_t stile = ring;                    // [128][BN + 4] fp32
    const uint32_t lds = (uint32_t)(BN + 4) * 4;    // row pitch in bytes
    {
      const int quarter = warp % TC_GROUP_WARPS;    // a warp may only touch TMEM lanes [32*(warp%4), +32)
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t srow = stile + (uint32_t)(quarter * 32 + lane) * lds;
      for (int cb = group * 16; cb < BN; cb += 16 * TC_GROUPS) {   // the two groups split the 16-column blocks
        float v[16];
        tmem_ld16(lane_base + cb, v);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(srow + (uint32_t)(cb + q * 4) * 4, make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]));
      }
    }
    if (tid == 0) TC_TRACE(58);
    asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCER_WARPS * 32) : "memory");   // producer warps only
    if (tid == 0) TC_TRACE(59);
    const int dgru = (g.epi == EPI_GRU_ZR) ? g.N / 2 : g.N;
    constexpr int ROWS_PER_WARP = TC_BM / TC_PRODUCER_WARPS;
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
      const int r = m0 + warp * ROWS_PER_WARP + rr;
      if (r >= row_end) break;
      const uint32_t srow = stile + (uint32_t)(warp * ROWS_PER_WARP + rr) * lds;
      for (int c4 = lane; c4 < BN / 4; c4 += 32) {
        const int c = n0 + c4 * 4;
        if (c >= p.n_total) break;
        float4 o = lds128(srow + (uint32_t)c4 * 16);
        if (g.bias != nullptr) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + c));
          o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        if (g.epi == EPI_STORE) {
          o = act4(o, g.act);
          *reinterpret_cast<float4*>(C + (size_t)r * g.ldc + c) = o;
        } else if (g.epi == EPI_GRU_ZR) {
          o.x = hard_sigmoid(o.x); o.y = hard_sigmoid(o.y); o.z = hard_sigmoid(o.z); o.w = hard_sigmoid(o.w);
          if (c < dgru) {
            *reinterpret_cast<float4*>(C + (size_t)r * g.ldc + c) = o;                      // z gate
          } else {
            const int cc = c - dgru;
            const float4 h = __ldg(reinterpret_cast<const float4*>(g.aux_h + (size_t)r * g.ld_h + cc));
            *reinterpret_cast<float4*>(g.C2 + (size_t)r * g.ldc2 + cc) = make_float4(o.x * h.x, o.y * h.y, o.z * h.z, o.w * h.w);
          }
        } else {  // EPI_GRU_OUT: h' = z*h + (1-z)*act(.)
          o = act4(o, g.act);
          const float4 h = __ldg(reinterpret_cast<const float4*>(g.aux_h + (size_t)r * g.ld_h + c));
          const float4 zz = __ldg(reinterpret_cast<const float4*>(g.aux_z + (size_t)r * g.ld_z + c));
          *reinterpret_cast<float4*>(C + (size_t)r * g.ldc + c) =
              make_float4(zz.x * h.x + (1.0f - zz.x) * o.x, zz.y * h.y + (1.0f - zz.y) * o.y,
                          zz.z * h.z + (1.0f - zz.z) * o.z, zz.w * h.w + (1.0f - zz.w) * o.w);
        }
      }
    }
    if (tid == 0) TC_TRACE(57);
  } else if (warp == TC_PRODUCER_WARPS) {
    // =========================== MMA issuer (one thread) ===========================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TC_BM, BN);
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % S;
        mbar_wait(full0 + 8 * s, (c / S) & 1);
        if (c < 8) TC_TRACE(32 + 2 * c);
        tc_fence_after_sync();
        const uint32_t a_hi = ring + (uint32_t)(s * stage_bytes);
        const uint32_t a_lo = a_hi + A_IMG_BYTES;
        const uint32_t b_hi = a_lo + A_IMG_BYTES;
        const uint32_t b_lo = b_hi + b_img_bytes;
        const uint64_t da_hi = make_sw128_desc(a_hi), da_lo = make_sw128_desc(a_lo);
        const uint64_t db_hi = make_sw128_desc(b_hi), db_lo = make_sw128_desc(b_lo);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {       // UMMA_K = 8 tf32 = 32 bytes: advance the start address by 2 (>>4 units)
          const uint64_t adv = (uint64_t)(k * 2);
          umma_tf32(tmem_base, da_lo + adv, db_hi + adv, idesc, (c | k) != 0);   // small terms first
          umma_tf32(tmem_base, da_hi + adv, db_lo + adv, idesc, 1);
          umma_tf32(tmem_base, da_hi + adv, db_hi + adv, idesc, 1);
        }
        umma_commit(empty0 + 8 * s);                // stage reusable once these MMAs have read it
        if (c < 8) TC_TRACE(33 + 2 * c);
      }
      umma_commit(accum_bar);                       // accumulator complete -> epilogue
    }
    __syncwarp();
  } else {
    // =========================== B loader (TMA bulk copies of the packed images) ===========================
    if (lane == 0) {
      const float* src_tile = packed + (size_t)n_tile * nchunks * 2 * (BN * TC_BK);
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % S;
        if (c >= S) mbar_wait(empty0 + 8 * s, ((c / S) - 1) & 1);
        const uint32_t b_hi = ring + (uint32_t)(s * stage_bytes + 2 * A_IMG_BYTES);
        if (c < 8) TC_TRACE(2 + 3 * c);
        mbar_arrive_expect_tx(full0 + 8 * s, 2 * b_img_bytes);
        bulk_copy_g2s(b_hi, src_tile + (size_t)c * 2 * (BN * TC_BK), 2 * b_img_bytes, full0 + 8 * s);   // hi and lo are adjacent
      }
    }
    __syncwarp();
  }

  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (warp == TC_PRODUCER_WARPS) {
    tc_fence_after_sync();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

int pick_bn(long m_tiles, int n_total, int gz) {
  int best = 32;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 32; bn -= 16) {
    const long ctas = m_tiles * ((n_total + bn - 1) / bn) * gz;
    const long waves = (ctas + 147) / 148;
    const double cost = (double)waves * (96.0 + bn);   // per-tile time ~ fixed overhead + columns
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace

// ---- optional cache of packed weight images (static weights: inference / benchmarking) ----------------
// Off by default.  Keyed by every input of the packing (weight pointers, leading dims, K/N, batching, BN);
// the caller promises not to modify cached weights in place without calling rgnn_weight_cache_clear().
struct PackKey {
  uint64_t h[4];
  bool operator==(const PackKey& o) const { return h[0] == o.h[0] && h[1] == o.h[1] && h[2] == o.h[2] && h[3] == o.h[3]; }
};
struct PackEntry { PackKey key; float* images; size_t bytes; };
static std::vector<PackEntry> g_pack_cache;
static std::mutex g_pack_mutex;
static bool g_pack_cache_on = false;

static inline void mix(uint64_t& h, uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
static PackKey make_pack_key(const GemmParams& g, int BN, int device) {
  PackKey k = {{0x1234, 0x5678, 0x9abc, (uint64_t)device}};
  const int nb = (g.batch_mode == BATCH_NONE) ? 1 : g.batch;
  for (int j = 0; j < nb; ++j) {
    const uint64_t a = (uint64_t)(g.batch_mode == BATCH_NONE ? g.B1 : g.bptr[j]);
    const uint64_t b = (uint64_t)(g.batch_mode == BATCH_NONE ? g.B2 : g.bptr2[j]);
    mix(k.h[j & 1], a); mix(k.h[2], b); mix(k.h[3], a * 31 + b + j);
  }
  mix(k.h[0], ((uint64_t)g.K1 << 32) | (uint32_t)g.K2);
  mix(k.h[1], ((uint64_t)g.N << 32) | (uint32_t)BN);
  mix(k.h[2], ((uint64_t)g.ldb1 << 32) | (uint32_t)g.ldb2);
  mix(k.h[3], ((uint64_t)g.batch_mode << 32) | (uint32_t)g.batch);
  return k;
}

void gemm_weight_cache_enable(bool on) { std::lock_guard<std::mutex> l(g_pack_mutex); g_pack_cache_on = on; }
void gemm_weight_cache_clear() {
  std::lock_guard<std::mutex> l(g_pack_mutex);
  for (auto& e : g_pack_cache) cudaFree(e.images);
  g_pack_cache.clear();
}
bool gemm_weight_cache_enabled() { return g_pack_cache_on; }

size_t gemm_tc_pack_bytes(const GemmParams& g) {
  if (g_pack_cache_on) return 1024;   // images live in the cache, the caller's scratch is not used
  const int chunks = (g.K1 + TC_BK - 1) / TC_BK + (g.K2 + TC_BK - 1) / TC_BK;
  const int n_total = (g.batch_mode == BATCH_SHARED_A) ? g.batch * g.N : g.N;
  const int gz = (g.batch_mode == BATCH_ROW_RANGES || g.batch_mode == BATCH_COL_BLOCKS) ? g.batch : 1;
  const int rows = (g.batch_mode == BATCH_ROW_RANGES) ? g.max_rows : g.M;
  const int bn = pick_bn((rows + TC_BM - 1) / TC_BM, n_total, gz);
  const size_t tiles = (n_total + bn - 1) / bn;
  return align_up(tiles * chunks * 2 * (size_t)bn * TC_BK * sizeof(float) * gz, 1024);
}

int launch_gemm_tcgen05(const GemmParams& g, void* pack_ws, size_t pack_ws_bytes, cudaStream_t stream) {
  RGNN_REQUIRE(g.M >= 0 && g.N > 0 && g.K1 > 0 && g.K2 >= 0, "gemm: bad dims M=%d N=%d K1=%d K2=%d", g.M, g.N, g.K1, g.K2);
  RGNN_REQUIRE((g.N % 4) == 0 && (g.K1 % 4) == 0 && (g.K2 % 4) == 0, "gemm: N, K must be multiples of 4 (N=%d K1=%d K2=%d)", g.N, g.K1, g.K2);
  RGNN_REQUIRE((g.lda1 % 4) == 0 && (g.ldb1 % 4) == 0 && (g.ldc % 4) == 0, "gemm: leading dims must keep 16-byte rows");
  RGNN_REQUIRE(g.batch >= 1 && g.batch <= RGNN_MAX_EDGE_TYPES, "gemm: batch %d out of range", g.batch);
  RGNN_REQUIRE(aligned16(g.A1) && aligned16(g.C) && (g.K2 == 0 || aligned16(g.A2)), "gemm: operands must be 16-byte aligned");
  RGNN_REQUIRE(g.bias == nullptr || aligned16(g.bias), "gemm: bias must be 16-byte aligned");
  const int rows = (g.batch_mode == BATCH_ROW_RANGES) ? g.max_rows : g.M;
  if (rows <= 0) return RGNN_OK;
  TcParams p;
  p.g = g;
  p.chunks1 = (g.K1 + TC_BK - 1) / TC_BK;
  p.chunks2 = (g.K2 + TC_BK - 1) / TC_BK;
  const int nchunks = p.chunks1 + p.chunks2;
  p.n_total = (g.batch_mode == BATCH_SHARED_A) ? g.batch * g.N : g.N;
  const int gz = (g.batch_mode == BATCH_ROW_RANGES || g.batch_mode == BATCH_COL_BLOCKS) ? g.batch : 1;
  p.BN = pick_bn((rows + TC_BM - 1) / TC_BM, p.n_total, gz);
  const int n_tiles = (p.n_total + p.BN - 1) / p.BN;
  const size_t stage_bytes = 2 * (size_t)A_IMG_BYTES + 2 * (size_t)p.BN * 128;
  p.stages = (int)(TC_SMEM_BUDGET / stage_bytes);
  if (p.stages > 4) p.stages = 4;
  if (p.stages > nchunks) p.stages = nchunks;
  if (p.stages < 1) p.stages = 1;
  p.tmem_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
  p.packed_stride = (size_t)n_tiles * nchunks * 2 * p.BN * TC_BK;
  const size_t need = align_up(p.packed_stride * sizeof(float) * gz, 1024);
  bool need_pack = true;
  if (g_pack_cache_on) {
    int device = 0;
    cudaGetDevice(&device);
    const PackKey key = make_pack_key(g, p.BN, device);
    std::lock_guard<std::mutex> l(g_pack_mutex);
    float* images = nullptr;
    for (auto& e : g_pack_cache)
      if (e.key == key && e.bytes == need) { images = e.images; need_pack = false; break; }
    if (images == nullptr) {
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(stream, &cap);
      RGNN_REQUIRE(cap == cudaStreamCaptureStatusNone, "gemm: weight cache miss during CUDA-graph capture (run the layer once eagerly first)");
      RGNN_CHECK_CUDA(cudaMalloc(&images, need));
      g_pack_cache.push_back({key, images, need});
    }
    pack_ws = images;
  } else {
    RGNN_REQUIRE(pack_ws != nullptr && pack_ws_bytes >= need && (reinterpret_cast<uintptr_t>(pack_ws) & 15u) == 0,
                 "gemm: weight-image workspace too small (%zu < %zu)", pack_ws_bytes, need);
  }
  p.packed = static_cast<const float*>(pack_ws);

  // ---- pack the weights into shared-memory images ----
  for (int zz = 0; need_pack && zz < gz; ++zz) {
    PackParams q;
    q.ldb1 = g.ldb1; q.ldb2 = g.ldb2; q.K1 = g.K1; q.K2 = g.K2;
    q.BN = p.BN; q.chunks1 = p.chunks1; q.chunks2 = p.chunks2;
    q.n_total = p.n_total;
    if (g.batch_mode == BATCH_SHARED_A) {
      q.block_cols = g.N;
      for (int j = 0; j < g.batch; ++j) { q.b1[j] = g.bptr[j]; q.b2[j] = g.bptr2[j]; }
    } else if (g.batch_mode == BATCH_NONE) {
      q.block_cols = g.N; q.b1[0] = g.B1; q.b2[0] = g.B2;
    } else {
      q.block_cols = g.N; q.b1[0] = g.bptr[zz]; q.b2[0] = g.bptr2[zz];
    }
    for (int j = 0; j < (g.batch_mode == BATCH_SHARED_A ? g.batch : 1); ++j) {
      RGNN_REQUIRE(q.b1[j] != nullptr && (g.K2 == 0 || q.b2[j] != nullptr), "gemm: weight pointer %d is NULL", j);
    }
    q.out = static_cast<float*>(pack_ws) + (size_t)zz * p.packed_stride;
    pack_b_kernel<<<dim3(nchunks, n_tiles, (p.BN + 31) / 32), 256, 0, stream>>>(q);
    RGNN_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }

  const size_t epi_tile = (size_t)TC_BM * (p.BN + 4) * sizeof(float);
  p.ring_bytes = (int)align_up(std::max((size_t)p.stages * stage_bytes, epi_tile), 1024);
  const size_t smem = 1024 + (size_t)p.ring_bytes + (2 * p.stages + 1) * sizeof(uint64_t) + 16;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    RGNN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    attr_smem = 220 * 1024;
  }
  dim3 grid(n_tiles, (rows + TC_BM - 1) / TC_BM, gz);
  static const bool trace_on = getenv("RGNN_GEMM_TRACE") != nullptr;
  const size_t n_ctas = (size_t)grid.x * grid.y * grid.z;
  p.trace = nullptr;
  if (trace_on) {   // debug only: synchronous, prints one timeline summary per launch to stderr
    RGNN_CHECK_CUDA(cudaMalloc(&p.trace, n_ctas * 64 * sizeof(long long)));
    RGNN_CHECK_CUDA(cudaMemsetAsync(p.trace, 0, n_ctas * 64 * sizeof(long long), stream));
  }
  gemm_tcgen05_kernel<<<grid, TC_THREADS, smem, stream>>>(p);
  RGNN_CHECK_CUDA(cudaGetLastError());
  count_launch();
  if (trace_on) {
    std::vector<long long> h(n_ctas * 64);
    RGNN_CHECK_CUDA(cudaStreamSynchronize(stream));
    RGNN_CHECK_CUDA(cudaMemcpy(h.data(), p.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    fprintf(stderr, "[gemm trace] M=%d N=%d K=%d+%d BN=%d stages=%d ctas=%zu chunks=%d\n", g.M, p.n_total, g.K1, g.K2, p.BN, p.stages, n_ctas, nchunks);
    const size_t picks[3] = {0, n_ctas / 2, n_ctas - 1};
    for (size_t pi = 0; pi < 3; ++pi) {
      const long long* t = h.data() + picks[pi] * 64;
      fprintf(stderr, "  cta %zu: prologue %lld |", picks[pi], t[0]);
      for (int c = 0; c < 8 && c < nchunks; ++c) fprintf(stderr, " c%d got %lld Bissue %lld arrived %lld mma_full %lld mma_commit %lld |", c, t[1 + 3 * c], t[2 + 3 * c], t[3 + 3 * c], t[32 + 2 * c], t[33 + 2 * c]);
      fprintf(stderr, " epi %lld staged %lld bar %lld end %lld\n", t[56], t[58], t[59], t[57]);
    }
  }
  return RGNN_OK;
}

}  // namespace rgnn
