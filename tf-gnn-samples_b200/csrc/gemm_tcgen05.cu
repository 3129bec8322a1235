// gemm_tcgen05.cu -- C = epilogue([A1|A2] . [B1;B2]) on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Contract: fp32 in / fp32 out, fp32-level accuracy through 3xTF32 split accumulation (as round 1's mma.sync
// kernel gemm_tf32x3.cu, removed in round 2), on the Blackwell-native path: the legacy mma.sync TF32 pipe tops out at
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
#include "tc_ptx.cuh"

#include <stdlib.h>
#include <vector>
#include <mutex>
#include <algorithm>

namespace rgnn {

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                    // fp32 elements per 128-byte swizzle row
#ifndef RGNN_PAIR_DEFAULT
#define RGNN_PAIR_DEFAULT 0                  // 1: large-M, wide-tile contractions use CTA pairs unless RGNN_GEMM_PAIR=0
#endif
#ifndef RGNN_TC_GROUPS
#define RGNN_TC_GROUPS 3
#endif
constexpr int TC_GROUPS = RGNN_TC_GROUPS;    // producer groups take K chunks round-robin (see the producer loop): bytes of A in flight per CTA = TC_GROUPS x 16 KB
constexpr int TC_GROUP_WARPS = 4;            // one group covers the 128-row tile: 4 warps x 32 rows of TMEM lanes
constexpr int TC_PRODUCER_WARPS = TC_GROUPS * TC_GROUP_WARPS;
constexpr int TC_GROUP_THREADS = 32 * TC_GROUP_WARPS;
constexpr int A_IMG_BYTES = TC_BM * 128;     // 16 KB
constexpr size_t TC_SMEM_MAX = 227 * 1024;       // opt-in dynamic shared memory per CTA on sm_100
// Per-CTA cycle trace of the kernel's phases (tools/gemm_trace.py builds a second library with -DRGNN_GEMM_TRACE; the
// shipped library compiles these hooks away).  Slots: 0 entry, 1 set-up done, 2+q producer published chunk q, 18+q MMA saw
// chunk q full, 34+q MMA committed chunk q, 50 accumulator complete (seen by epilogue warp 0), 51 epilogue warp 0 done,
// 52 / 53 producer warp 4 starts / finishes helping, 54 teardown, 55 / 56 first / last weight-image copy issued.
#ifdef RGNN_GEMM_TRACE
__device__ long long g_gemm_trace[160 * 64];
#define TC_TRACE(slot) do { g_gemm_trace[(blockIdx.x % 160) * 64 + (slot)] = clock64(); } while (0)
#else
#define TC_TRACE(slot) do { } while (0)
#endif
constexpr size_t TC_RING_BUDGET = TC_SMEM_MAX - 1024 /*align*/ - 18432 /*epilogue staging*/ - 512 /*barriers*/;

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
  int transposed, k_block;                // transposed: B[(z, j), n] = b1[z][n * ldb1 + j], K = blocks * k_block
  float* out;                             // [n_tiles][chunks1+chunks2][2][BN*32]
};

// grid = (k-chunks, n-tiles, BN/32 row slices... ceil), block = 256 threads = 32 image rows x 8 16-byte chunks
__global__ void __launch_bounds__(256) pack_b_kernel(const __grid_constant__ PackParams p) {
  // Programmatic dependent launch: the image buffer may still be read by the GEMM enqueued before (the layers reuse one scratch
  // region), so nothing is written before the predecessor has completed; the GEMM that consumes these images may start its
  // set-up (barrier init, TMEM allocation) right away -- it waits for this grid before touching them.
  pdl_wait();
  pdl_launch_dependents();
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
    if (p.transposed) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + c16 * 4 + j;
        if (k < Kseg) {
          const int blk = k / p.k_block, jj = k - blk * p.k_block;
          x[j] = __ldg(p.b1[blk] + (size_t)n * ld + jj);
        }
      }
    } else {
      const int blk = n / p.block_cols, col = n - blk * p.block_cols;
      const float* src = (seg2 ? p.b2[blk] : p.b1[blk]) + col;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + c16 * 4 + j;
        if (k < Kseg) x[j] = __ldg(src + (size_t)k * ld);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) split_tf32(x[j], hi[j], lo[j]);
  const int off = (nl >> 3) * 256 + (nl & 7) * 32 + ((c16 ^ (nl & 7)) << 2);   // float index inside the image
  *reinterpret_cast<float4*>(img_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<float4*>(img_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// -------------------------------------------------------------------------------------------------
// Persistent, warp-specialised kernel.  Roles (14 warps):
//   warps 0-3   epilogue: TMEM -> registers -> padded smem staging -> coalesced global stores
//   warps 4-11  A producers, two groups of 4 warps alternating K chunks
//   warp 12     MMA issuer (one elected thread) + TMEM alloc / dealloc
//   warp 13     B loader (one elected thread, TMA bulk copies of the packed weight images)
// Each CTA walks tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...  The accumulator is double-buffered in TMEM
// (2 x BN_alloc columns), so the write-back of tile i overlaps the mainloop of tile i+1.
// -------------------------------------------------------------------------------------------------
constexpr int EPI_WARPS = 4;
constexpr int PROD_WARP0 = EPI_WARPS;
constexpr int MMA_WARP = PROD_WARP0 + TC_PRODUCER_WARPS;
constexpr int BLOAD_WARP = MMA_WARP + 1;
constexpr int TC_THREADS_V4 = 32 * (BLOAD_WARP + 1);
constexpr int EPI_COLS = 32;                       // columns staged per epilogue step
constexpr int EPI_PITCH = (EPI_COLS + 4) * 4;      // bytes; 144: 8 consecutive rows hit 8 distinct 16-byte bank groups
constexpr int EPI_STAGE_BYTES = TC_BM * EPI_PITCH; // 18 KB

struct TcParams {
  GemmParams g;
  const float* packed;     // pack_b output (per z for ROW_RANGES / COL_BLOCKS: z * packed_stride floats)
  size_t packed_stride;
  int BN, stages, chunks1, chunks2;
  int n_total;             // columns of C covered by this launch (batch*N for SHARED_A, else N)
  int n_tiles;
  int tmem_cols;           // total allocation (two accumulators of tmem_cols/2 columns)
  int ring_bytes;
  int total_tiles;         // number of tile GROUPS: a group = CL consecutive m-tiles x one n-tile, one per cluster step
  int tile_start[RGNN_MAX_EDGE_TYPES + 1];   // first group of batch entry z (ROW_RANGES / COL_BLOCKS), else {0, total}
};

struct TileInfo { int z, m0, row_end, n_tile; };

// group t, CTA rank cr inside its cluster of CL: the m-tile is (group row) * CL + cr; it may lie past the last row
// (dummy tile: loads are zero-filled, stores masked) so that every CTA of a cluster runs the same barrier protocol.
template <int CL>
__device__ __noinline__ TileInfo decode_tile(const TcParams& p, int t, int cr) {   // once per tile and role: keep it out of line (I-cache)
  TileInfo ti;
  const GemmParams& g = p.g;
  int z = 0;
  if (g.batch_mode == BATCH_ROW_RANGES || g.batch_mode == BATCH_COL_BLOCKS) {
    while (t >= p.tile_start[z + 1]) ++z;
  }
  const int local = t - p.tile_start[z];
  const int row_begin = (g.batch_mode == BATCH_ROW_RANGES) ? g.row_off[z] : 0;
  ti.z = z;
  ti.row_end = (g.batch_mode == BATCH_ROW_RANGES) ? g.row_off[z + 1] : g.M;
  ti.m0 = row_begin + ((local / p.n_tiles) * CL + cr) * TC_BM;
  ti.n_tile = local % p.n_tiles;
  return ti;
}

// One warp's share of a tile's write-back: for the 32-column blocks cb = cb_begin, cb_begin + cb_step, ... read this warp's 32
// TMEM lanes (thread = row) with tcgen05.ld, stage them in padded shared memory, and store rows so that 8 lanes write 128
// contiguous bytes (every store instruction covers 4 full 128-byte lines).  Called from ONE site that the epilogue
// warps (all blocks, or every third block of a CTA's last tile) and the producer warps (helping with the last tile) share.
template <int EPI>
__device__ __forceinline__ void epilogue_blocks(const TcParams& p, const TileInfo ti, uint32_t lane_base, uint32_t stage_q,
                                             int quarter, int lane, int cb_begin, int cb_step) {
  const GemmParams& g = p.g;
  const int BN = p.BN;
  const uint32_t srow_w = stage_q + (uint32_t)lane * EPI_PITCH;
  const int sub_row = lane >> 3, sub_c4 = lane & 7;  // read mapping: 4 rows x 8 float4 per instruction
  const int dgru = (EPI == EPI_GRU_ZR) ? g.N / 2 : g.N;
  float* C = g.C + (g.batch_mode == BATCH_COL_BLOCKS ? (size_t)ti.z * g.N : 0);
  const int n0 = ti.n_tile * BN;
  for (int cb = cb_begin; cb < BN; cb += cb_step) {
    {
#ifdef RGNN_EPI_LD16
      float v[16];
      tmem_ld16(lane_base + cb, v);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd)
        sts128(srow_w + (uint32_t)(qd * 16), make_float4(v[qd * 4], v[qd * 4 + 1], v[qd * 4 + 2], v[qd * 4 + 3]));
      tmem_ld16(lane_base + cb + 16, v);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd)
        sts128(srow_w + (uint32_t)(64 + qd * 16), make_float4(v[qd * 4], v[qd * 4 + 1], v[qd * 4 + 2], v[qd * 4 + 3]));
#else
      float v[32];                                   // the whole 32-column block of this lane's row: one TMEM round trip
      tmem_ld32(lane_base + cb, v);
#pragma unroll
      for (int qd = 0; qd < 8; ++qd)
        sts128(srow_w + (uint32_t)(qd * 16), make_float4(v[qd * 4], v[qd * 4 + 1], v[qd * 4 + 2], v[qd * 4 + 3]));
#endif
    }
    __syncwarp();
    const int c = n0 + cb + sub_c4 * 4;
    const bool col_ok = c < p.n_total;
    float4 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = lds128(stage_q + (uint32_t)(j * 4 + sub_row) * EPI_PITCH + (uint32_t)sub_c4 * 16);
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.bias != nullptr && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(g.bias + c));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = ti.m0 + quarter * 32 + j * 4 + sub_row;
      if (r < ti.row_end && col_ok) {
        float4 x = make_float4(o[j].x + bias4.x, o[j].y + bias4.y, o[j].z + bias4.z, o[j].w + bias4.w);
        if (EPI == EPI_STORE) {
          x = act4_cold(x, g.act);
          *reinterpret_cast<float4*>(C + (size_t)r * g.ldc + c) = x;
        } else if (EPI == EPI_GRU_ZR) {
          x.x = hard_sigmoid(x.x); x.y = hard_sigmoid(x.y); x.z = hard_sigmoid(x.z); x.w = hard_sigmoid(x.w);
          if (c < dgru) {
            *reinterpret_cast<float4*>(C + (size_t)r * g.ldc + c) = x;                      // z gate
          } else {
            const int cc = c - dgru;
            const float4 h = __ldg(reinterpret_cast<const float4*>(g.aux_h + (size_t)r * g.ld_h + cc));
            *reinterpret_cast<float4*>(g.C2 + (size_t)r * g.ldc2 + cc) = make_float4(x.x * h.x, x.y * h.y, x.z * h.z, x.w * h.w);
          }
        } else {  // EPI_GRU_OUT: h' = z*h + (1-z)*act(.)
          x = act4_cold(x, g.act);
          const float4 h = __ldg(reinterpret_cast<const float4*>(g.aux_h + (size_t)r * g.ld_h + c));
          const float4 zz = __ldg(reinterpret_cast<const float4*>(g.aux_z + (size_t)r * g.ld_z + c));
          *reinterpret_cast<float4*>(C + (size_t)r * g.ldc + c) =
              make_float4(zz.x * h.x + (1.0f - zz.x) * x.x, zz.y * h.y + (1.0f - zz.y) * x.y,
                          zz.z * h.z + (1.0f - zz.z) * x.z, zz.w * h.w + (1.0f - zz.w) * x.w);
        }
      }
    }
    __syncwarp();                               // staging rows are rewritten by the next column block
  }
}

// CL = thread-block cluster size (1, 2 or 4).  The CL CTAs of a cluster work on CL consecutive m-tiles of the SAME
// n-tile in lock step; each loads 1/CL of every weight-image chunk and TMA-multicasts it to all of them, so the
// L2->SM traffic of B (the bound of this kernel, profiles/r01_gemm_tcgen05_timeline.txt) drops by CL.
// GATHER: A rows follow an index list (GemmParams::a_rows: the compact (source, type) transform).  A separate instance,
// so that the common kernel stays under the ~2048-instruction L1.5 I-cache (the 16 extra index loads + selects of the
// gathered form pushed every variant over it: 86.9 -> 107 us on the headline step, gpurun_out r02 job B).
// PAIR (CL == 2): tcgen05 cta_group::2.  The two CTAs of a cluster work on two consecutive m-tiles of one n-tile as ONE
// 256 x BN MMA issued by the leader (cluster rank 0): each CTA produces the A images of its own 128 rows and streams only
// HALF of every weight-image chunk (rows [rank * BN/2, +BN/2) of the hi and of the lo image) into its own shared memory --
// the tensor cores read both halves.  Per CTA and chunk that is 16 KB of A + BN * 128 B of B instead of 16 KB + BN * 256 B:
// the per-SM operand ingest, which bounds this kernel on large M, nearly halves, and the stage shrinks so
// that a third ring stage fits at BN = 256.  Protocol: every CTA keeps its own full / empty barriers; the peer's MMA warp
// forwards "my stage is full" to the leader's barrier (remote arrive); the leader's commits are multicast to both CTAs'
// empty / accumulator-full barriers; the peer's epilogue warps release the accumulator on the leader's barrier.
template <int EPI, int CL, bool GATHER = false, bool PAIR = false>
__global__ void __launch_bounds__(TC_THREADS_V4, 1) gemm_tcgen05_kernel(const __grid_constant__ TcParams p) {
  static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const GemmParams& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN, S = p.stages;
  const int b_img_bytes = PAIR ? (BN / 2) * 128 : BN * 128;   // PAIR: this CTA holds half of the chunk's weight rows
  const int stage_bytes = 2 * A_IMG_BYTES + 2 * b_img_bytes;
  const int nchunks = p.chunks1 + p.chunks2;
  // shared-memory map (32-bit shared addresses): operand ring | epilogue staging | barriers | tmem slot
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t estage = ring + (uint32_t)p.ring_bytes;
  const uint32_t full0 = estage + EPI_STAGE_BYTES, empty0 = full0 + 8 * S;
  const uint32_t tfull0 = empty0 + 8 * S, tempty0 = tfull0 + 16;
  const uint32_t tmem_slot = tempty0 + 16;
  const int acc_cols = p.tmem_cols / 2;
  const int cr = (CL > 1) ? (int)cluster_cta_rank() : 0;
  const int cid = (int)blockIdx.x / CL, ncl = (int)gridDim.x / CL;
  const uint16_t cl_mask = (uint16_t)((1u << CL) - 1u);
  const int my_tiles = (p.total_tiles - cid + ncl - 1) / ncl;   // tile groups of this cluster (same for its CTAs)

  if (tid == 0) TC_TRACE(0);
  if (warp == MMA_WARP && lane == 0) {
    for (int s = 0; s < S; ++s) {
      // one producer group + the B loader's expect_tx arrival (+ the peer's "my stage is full" on the leader of a pair)
      mbar_init(full0 + 8 * s, TC_GROUP_THREADS + 1 + ((PAIR && cr == 0) ? 1 : 0));
      mbar_init(empty0 + 8 * s, PAIR ? 1 : CL);         // one tcgen05.commit from every issuing CTA (a pair: the leader only)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);                     // accumulator complete (tcgen05.commit)
      mbar_init(tempty0 + 8 * a, PAIR ? 2 * EPI_WARPS : EPI_WARPS);   // accumulator drained (one lane per epilogue warp, both CTAs of a pair)
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == MMA_WARP) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                       // every CTA's barriers are initialised before any remote arrive
  tc_fence_after_sync();
  const uint32_t tmem_base = lds32(tmem_slot);
  if (tid == 0) TC_TRACE(1);
  // Everything above (barrier init, TMEM allocation) touched only this CTA's own state and may have run while the previous
  // kernel of the stream was still finishing (programmatic dependent launch); A, the weight images and C must not be touched
  // before that kernel has completed.
  pdl_wait();
  pdl_launch_dependents();

  if (warp >= PROD_WARP0 && warp < MMA_WARP) {
    // =========================== A producers ===========================
    // Group g owns the chunks q = g, g+2, ... of this CTA's flat (tile, chunk) sequence.  A thread issues the
    // loads of its NEXT chunk right after publishing the current one: fence.proxy.async lowers to
    // MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC and the membar waits for the issuing thread's outstanding loads, so the
    // prefetch has to live in the other group's time slot, not inside this thread's store phase.
    const int group = (warp - PROD_WARP0) / TC_GROUP_WARPS;
    const int ptid = tid - 32 * PROD_WARP0 - group * TC_GROUP_THREADS;
    // Groups in use: never more than ring stages.  A group that has published chunk q waits for the stage of chunk q + G; with
    // G > S that stage's "empty" barrier can still be TWO phases behind the awaited one, and an mbarrier parity wait cannot
    // tell "two behind" from "done" (the producer would overwrite a stage the tensor core has not read yet).
    const int ngroups = S < TC_GROUPS ? S : TC_GROUPS;
    const int total_q = group < ngroups ? my_tiles * nchunks : 0;
    auto load_a_chunk = [&](int q, float4 (&v)[8]) {
      const int ti_idx = q / nchunks, c = q - ti_idx * nchunks;
      const TileInfo ti = decode_tile<CL>(p, cid + ti_idx * ncl, cr);
      const bool seg2 = c >= p.chunks1;
      const int k0 = (seg2 ? c - p.chunks1 : c) * TC_BK;
      const int Kseg = seg2 ? g.K2 : g.K1;
      const float* Abase = seg2 ? g.A2 : (g.batch_mode == BATCH_COL_BLOCKS ? g.A1 + (size_t)ti.z * g.K1 : g.A1);
      const int lda = seg2 ? g.lda2 : g.lda1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int f = ptid + i * TC_GROUP_THREADS;
        const int row = f >> 3, c16 = f & 7;
        const int grow = ti.m0 + row, gk = k0 + c16 * 4;
        if (GATHER) {
          const bool in = grow < ti.row_end && gk < Kseg;
          const int arow = in ? __ldg(g.a_rows + grow) : 0;   // the index list hits L1 after the tile's first chunk
          v[i] = in ? __ldg(reinterpret_cast<const float4*>(Abase + (size_t)arow * lda + gk)) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          v[i] = (grow < ti.row_end && gk < Kseg) ? __ldg(reinterpret_cast<const float4*>(Abase + (size_t)grow * lda + gk))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    float4 va[8];
    if (group < total_q) load_a_chunk(group, va);
    for (int q = group; q < total_q; q += ngroups) {
      const int s = q % S;
      const int use = q / S;
      if (use > 0) mbar_wait(empty0 + 8 * s, (use - 1) & 1);
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
      if (ptid == 0 && q < 16) TC_TRACE(2 + q);
      if (q + ngroups < total_q) load_a_chunk(q + ngroups, va);
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer (one thread) ===========================
    if (PAIR && cr != 0) {
      // peer of a pair: no MMAs to issue -- forward "this CTA's stage is full" (A images written and fenced by the producers,
      // weight half landed) to the leader's barrier, chunk by chunk
      if (lane == 0) {
        const int total_q = my_tiles * nchunks;
        for (int q = 0; q < total_q; ++q) {
          const int s = q % S;
          mbar_wait(full0 + 8 * s, (q / S) & 1);
          mbar_arrive_remote(full0 + 8 * s, 0);
        }
      }
    } else if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(PAIR ? 2 * TC_BM : TC_BM, BN);
      int q = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int a = it & 1;
        if (it >= 2) {                                                  // the epilogue(s) have drained this accumulator
          if (PAIR) mbar_wait_cluster(tempty0 + 8 * a, ((it >> 1) - 1) & 1);
          else mbar_wait(tempty0 + 8 * a, ((it >> 1) - 1) & 1);
        }
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * acc_cols);
        for (int c = 0; c < nchunks; ++c, ++q) {
          const int s = q % S;
          if (PAIR) mbar_wait_cluster(full0 + 8 * s, (q / S) & 1);
          else mbar_wait(full0 + 8 * s, (q / S) & 1);
          tc_fence_after_sync();
          if (q < 16) TC_TRACE(18 + q);
          const uint32_t a_hi = ring + (uint32_t)(s * stage_bytes);
          const uint32_t a_lo = a_hi + A_IMG_BYTES;
          const uint32_t b_hi = a_lo + A_IMG_BYTES;
          const uint32_t b_lo = b_hi + b_img_bytes;
          const uint64_t da_hi = make_sw128_desc(a_hi), da_lo = make_sw128_desc(a_lo);
          const uint64_t db_hi = make_sw128_desc(b_hi), db_lo = make_sw128_desc(b_lo);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {     // UMMA_K = 8 tf32 = 32 bytes: advance the start address by 2 (>>4 units)
            const uint64_t adv = (uint64_t)(k * 2);
            if (PAIR) {
              umma_tf32_pair(d_tmem, da_lo + adv, db_hi + adv, idesc, (c | k) != 0);
              umma_tf32_pair(d_tmem, da_hi + adv, db_lo + adv, idesc, 1);
              umma_tf32_pair(d_tmem, da_hi + adv, db_hi + adv, idesc, 1);
            } else {
              umma_tf32(d_tmem, da_lo + adv, db_hi + adv, idesc, (c | k) != 0);   // small terms first
              umma_tf32(d_tmem, da_hi + adv, db_lo + adv, idesc, 1);
              umma_tf32(d_tmem, da_hi + adv, db_hi + adv, idesc, 1);
            }
          }
          if (PAIR) umma_commit_pair(empty0 + 8 * s, cl_mask);          // both CTAs' loaders: stage s is reusable
          else if (CL > 1) umma_commit_multicast(empty0 + 8 * s, cl_mask);   // every CTA's loaders learn that this CTA is done with stage s
          else umma_commit(empty0 + 8 * s);         // stage reusable once these MMAs have read it
          if (q < 16) TC_TRACE(34 + q);
        }
        if (PAIR) umma_commit_pair(tfull0 + 8 * a, cl_mask);            // accumulator complete -> both CTAs' epilogues
        else umma_commit(tfull0 + 8 * a);           // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else if (warp == BLOAD_WARP) {
    // =========================== B loader (TMA bulk copies of the packed images) ===========================
    if (lane == 0) {
      int q = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const TileInfo ti = decode_tile<CL>(p, cid + it * ncl, cr);
        const float* src_tile = p.packed + (size_t)ti.z * p.packed_stride + (size_t)ti.n_tile * nchunks * 2 * (BN * TC_BK);
        for (int c = 0; c < nchunks; ++c, ++q) {
          const int s = q % S;
          const int use = q / S;
          if (use > 0) mbar_wait(empty0 + 8 * s, (use - 1) & 1);
          const uint32_t b_hi = ring + (uint32_t)(s * stage_bytes + 2 * A_IMG_BYTES);
          mbar_arrive_expect_tx(full0 + 8 * s, 2 * b_img_bytes);   // the whole chunk lands here, 1/CL from each CTA (a pair: this CTA's half)
          if (q == 0) TC_TRACE(55);
          TC_TRACE(56);
          if (PAIR) {   // rows [cr * BN/2, +BN/2) of the hi image and of the lo image, packed back to back in this CTA's stage
            const char* chunk = reinterpret_cast<const char*>(src_tile + (size_t)c * 2 * (BN * TC_BK));
            bulk_copy_g2s(b_hi, chunk + (size_t)cr * b_img_bytes, b_img_bytes, full0 + 8 * s);
            bulk_copy_g2s(b_hi + b_img_bytes, chunk + (size_t)BN * 128 + (size_t)cr * b_img_bytes, b_img_bytes, full0 + 8 * s);
          } else if (CL > 1) {
            const uint32_t part = (uint32_t)(2 * b_img_bytes) / CL;
            bulk_copy_g2s_multicast(b_hi + cr * part, reinterpret_cast<const char*>(src_tile + (size_t)c * 2 * (BN * TC_BK)) + (size_t)cr * part,
                                    part, full0 + 8 * s, cl_mask);
          } else {
            bulk_copy_g2s(b_hi, src_tile + (size_t)c * 2 * (BN * TC_BK), 2 * b_img_bytes, full0 + 8 * s);   // hi and lo are adjacent
          }
        }
      }
    }
    __syncwarp();
  }

  if (warp < MMA_WARP) {
    // =========================== write-back (epilogue warps; producer warps for the last tile) ===========================
    // Warps 0-3 write back every tile.  Once the producer warps have nothing left to produce they help with this CTA's LAST
    // tile: of its 32-column blocks the epilogue warps keep 0, 3, 6, ..., producer group g takes 1 + g, 4 + g, ...  A warp may
    // only read the TMEM lanes of its quarter (warp index % 4).  When tfull has fired every MMA has consumed its operands, so
    // the operand ring is free and serves as the helpers' staging area.  With one tile per CTA the write-back was ~37 % of
    // the kernel (profiles/r01_gemm_tcgen05_timeline.txt).
    const bool is_prod = warp >= PROD_WARP0;
    const int quarter = warp & 3;
    const int pgroup = is_prod ? (warp - PROD_WARP0) / TC_GROUP_WARPS : 0;
    // Only the producer groups that produced (pgroup < ngroups_wb) help: they reach this point after publishing the last
    // tile's chunks, i.e. at most one tfull phase early.  An idle group would arrive many phases early, and a parity wait
    // cannot tell those apart (it would read the accumulator of an earlier tile).
    const int ngroups_wb = S < TC_GROUPS ? S : TC_GROUPS;
    const uint32_t stage_q = is_prod ? ring + (uint32_t)((pgroup * 4 + quarter) * 32) * EPI_PITCH
                                     : estage + (uint32_t)(quarter * 32) * EPI_PITCH;
    const int cb_last = is_prod ? (1 + pgroup) * EPI_COLS : 0;
    for (int it = is_prod ? (pgroup < ngroups_wb ? my_tiles - 1 : my_tiles) : 0; it < my_tiles; ++it) {
      const int a = it & 1;
      const TileInfo ti = decode_tile<CL>(p, cid + it * ncl, cr);
      mbar_wait(tfull0 + 8 * a, (it >> 1) & 1);
      __syncwarp();
      tc_fence_after_sync();
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * acc_cols);
      const bool last = (it == my_tiles - 1);
      if (lane == 0 && warp == 0 && it == 0) TC_TRACE(50);
      if (lane == 0 && warp == PROD_WARP0) TC_TRACE(52);
      epilogue_blocks<EPI>(p, ti, lane_base, stage_q, quarter, lane, last ? cb_last : 0, last ? (ngroups_wb + 1) * EPI_COLS : EPI_COLS);
      if (lane == 0 && warp == 0 && it == 0) TC_TRACE(51);
      if (lane == 0 && warp == PROD_WARP0) TC_TRACE(53);
      tc_fence_before_sync();
      __syncwarp();
      if (!is_prod && lane == 0) {                               // this warp's TMEM lanes of accumulator a are drained
        if (PAIR && cr != 0) mbar_arrive_remote(tempty0 + 8 * a, 0);   // the leader issues the MMAs that overwrite it
        else mbar_arrive(tempty0 + 8 * a);
      }
    }
  }

  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) TC_TRACE(54);
  if (CL > 1) cluster_sync_all();                       // no CTA leaves while peers may still multicast into it / arrive on it
  if (warp == MMA_WARP) {
    tc_fence_after_sync();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

int pick_bn(long m_tiles, int n_total, int gz) {
  int best = 32;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 32; bn -= 32) {   // the epilogue stages 32-column blocks
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
  return gemm_tc_pack_bytes_uncached(g);
}
size_t gemm_tc_pack_bytes_uncached(const GemmParams& g) {
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
  RGNN_REQUIRE(g.batch_mode != BATCH_K_BLOCKS_T || (g.k_block > 0 && g.K1 == g.batch * g.k_block && g.K2 == 0),
               "gemm: BATCH_K_BLOCKS_T needs K1 == batch * k_block and no second segment");

  TcParams p;
  p.g = g;
  p.chunks1 = (g.K1 + TC_BK - 1) / TC_BK;
  p.chunks2 = (g.K2 + TC_BK - 1) / TC_BK;
  const int nchunks = p.chunks1 + p.chunks2;
  p.n_total = (g.batch_mode == BATCH_SHARED_A) ? g.batch * g.N : g.N;
  const int gz = (g.batch_mode == BATCH_ROW_RANGES || g.batch_mode == BATCH_COL_BLOCKS) ? g.batch : 1;
  p.BN = pick_bn((rows + TC_BM - 1) / TC_BM, p.n_total, gz);
  const int n_tiles = (p.n_total + p.BN - 1) / p.BN;
  p.n_tiles = n_tiles;
  // cluster size: CTAs of a cluster share the weight images of one n-tile (TMA multicast)
  static const int cl_env = getenv("RGNN_GEMM_CLUSTER") ? atoi(getenv("RGNN_GEMM_CLUSTER")) : 0;
  const int m_tiles_all = (rows + TC_BM - 1) / TC_BM;
  // Measured on B200 (profiles/r01_gemm_cluster_sweep.txt): multicasting the weight images over clusters of 2 / 4 does not
  // speed this kernel up -- the bound is what each SM can ingest (~30 B/clk) and feed to the tensor core from its own
  // shared memory, not the L2 read traffic -- so the default stays 1; RGNN_GEMM_CLUSTER=2|4 keeps the path testable.
  (void)m_tiles_all;
  int CL = 1;
  if (cl_env == 1 || cl_env == 2 || cl_env == 4) CL = cl_env;
  // CTA pairs (tcgen05 cta_group::2): RGNN_GEMM_PAIR=1 forces them, =0 forbids them; default: large-M contractions with wide
  // tiles, where the per-SM ingest of the weight images is the bound
  static const int pair_env = getenv("RGNN_GEMM_PAIR") ? atoi(getenv("RGNN_GEMM_PAIR")) : -1;
  bool pair = false;
  if (g.a_rows == nullptr && CL == 1 && (p.BN % 32) == 0) {
    if (pair_env == 1) pair = true;
    else if (pair_env != 0) pair = RGNN_PAIR_DEFAULT && m_tiles_all >= 2 * 148 && p.BN >= 128;
  }
  if (pair) CL = 2;
  auto groups_of = [&](int nrows) { return (((nrows + TC_BM - 1) / TC_BM + CL - 1) / CL) * n_tiles; };
  // group table: batch entry z owns groups [tile_start[z], tile_start[z+1])
  p.tile_start[0] = 0;
  if (g.batch_mode == BATCH_ROW_RANGES) {
    for (int z = 0; z < g.batch; ++z) p.tile_start[z + 1] = p.tile_start[z] + groups_of(g.row_off[z + 1] - g.row_off[z]);
    p.total_tiles = p.tile_start[g.batch];
  } else if (g.batch_mode == BATCH_COL_BLOCKS) {
    for (int z = 0; z < g.batch; ++z) p.tile_start[z + 1] = p.tile_start[z] + groups_of(g.M);
    p.total_tiles = p.tile_start[g.batch];
  } else {
    p.total_tiles = groups_of(g.M);
    p.tile_start[1] = p.total_tiles;
  }
  if (p.total_tiles <= 0) return RGNN_OK;

  const size_t stage_bytes = 2 * (size_t)A_IMG_BYTES + 2 * (size_t)(pair ? p.BN / 2 : p.BN) * 128;
  p.stages = (int)(TC_RING_BUDGET / stage_bytes);
  if (p.stages > 4) p.stages = 4;
  if (p.stages < 1) p.stages = 1;
  p.ring_bytes = (int)(p.stages * stage_bytes);
  const int acc_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
  p.tmem_cols = 2 * acc_cols;                                   // double-buffered accumulator
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
    q.transposed = 0; q.k_block = 0;
    if (g.batch_mode == BATCH_K_BLOCKS_T) {
      q.transposed = 1; q.k_block = g.k_block; q.block_cols = g.N;
      for (int j = 0; j < g.batch; ++j) { q.b1[j] = g.bptr[j]; q.b2[j] = nullptr; }
    } else if (g.batch_mode == BATCH_SHARED_A) {
      q.block_cols = g.N;
      for (int j = 0; j < g.batch; ++j) { q.b1[j] = g.bptr[j]; q.b2[j] = g.bptr2[j]; }
    } else if (g.batch_mode == BATCH_NONE) {
      q.block_cols = g.N; q.b1[0] = g.B1; q.b2[0] = g.B2;
    } else {
      q.block_cols = g.N; q.b1[0] = g.bptr[zz]; q.b2[0] = g.bptr2[zz];
    }
    for (int j = 0; j < ((g.batch_mode == BATCH_SHARED_A || g.batch_mode == BATCH_K_BLOCKS_T) ? g.batch : 1); ++j) {
      RGNN_REQUIRE(q.b1[j] != nullptr && (g.K2 == 0 || q.b2[j] != nullptr), "gemm: weight pointer %d is NULL", j);
    }
    q.out = static_cast<float*>(pack_ws) + (size_t)zz * p.packed_stride;
    RGNN_CHECK_CUDA(launch_pdl(pack_b_kernel, dim3(nchunks, n_tiles, (p.BN + 31) / 32), dim3(256), 0, stream, q));
    RGNN_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }

  const size_t smem = 1024 + (size_t)p.ring_bytes + EPI_STAGE_BYTES + (2 * p.stages + 4) * sizeof(uint64_t) + 16;
  // per-DEVICE launch state (a process may drive several GPUs): SM count, opt-in shared memory, resident clusters
  constexpr int MAX_DEV = 64;
  static int num_sms_of[MAX_DEV] = {};
  int device = 0;
  cudaGetDevice(&device);
  const int dv = (device >= 0 && device < MAX_DEV) ? device : 0;
  if (num_sms_of[dv] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    num_sms_of[dv] = n > 0 ? n : 148;
  }
  const int num_sms = num_sms_of[dv];
  int max_clusters = num_sms / CL;

  using KernelFn = void (*)(TcParams);
  auto pick = [&](int epi) -> KernelFn {
    switch (epi * 8 + CL) {
      case EPI_STORE * 8 + 1: return gemm_tcgen05_kernel<EPI_STORE, 1>;
      case EPI_STORE * 8 + 2: return gemm_tcgen05_kernel<EPI_STORE, 2>;
      case EPI_STORE * 8 + 4: return gemm_tcgen05_kernel<EPI_STORE, 4>;
      case EPI_GRU_ZR * 8 + 1: return gemm_tcgen05_kernel<EPI_GRU_ZR, 1>;
      case EPI_GRU_ZR * 8 + 2: return gemm_tcgen05_kernel<EPI_GRU_ZR, 2>;
      case EPI_GRU_ZR * 8 + 4: return gemm_tcgen05_kernel<EPI_GRU_ZR, 4>;
      case EPI_GRU_OUT * 8 + 1: return gemm_tcgen05_kernel<EPI_GRU_OUT, 1>;
      case EPI_GRU_OUT * 8 + 2: return gemm_tcgen05_kernel<EPI_GRU_OUT, 2>;
      default: return gemm_tcgen05_kernel<EPI_GRU_OUT, 4>;
    }
  };
  KernelFn fn = pick(g.epi);
  if (pair) {
    fn = g.epi == EPI_STORE ? gemm_tcgen05_kernel<EPI_STORE, 2, false, true>
       : g.epi == EPI_GRU_ZR ? gemm_tcgen05_kernel<EPI_GRU_ZR, 2, false, true> : gemm_tcgen05_kernel<EPI_GRU_OUT, 2, false, true>;
  }
  if (g.a_rows != nullptr) {
    RGNN_REQUIRE(g.epi == EPI_STORE && CL == 1 && g.K2 == 0, "gemm: gathered A rows support the plain store epilogue, one K segment, no cluster");
    fn = gemm_tcgen05_kernel<EPI_STORE, 1, true>;
  }
  static bool attr_done[MAX_DEV][7][5] = {};
  const int attr_slot = g.a_rows != nullptr ? 3 : (pair ? 4 + g.epi : g.epi);
  if (!attr_done[dv][attr_slot][CL]) {
    RGNN_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_MAX));
    attr_done[dv][attr_slot][CL] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TC_THREADS_V4);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  static const bool pdl_off = getenv("RGNN_NO_PDL") != nullptr;   // A/B knob
  int na = 0;
  if (!pdl_off) {   // the kernel calls pdl_wait() after its set-up: its prologue overlaps the previous kernel's tail
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (CL > 1) {   // clusters of 4 cannot use every SM (GPC sizes): size the persistent grid by what is co-resident
    static int cached_of[MAX_DEV][2][3][5] = {};
    int (&cached)[3][5] = cached_of[dv][pair ? 1 : 0];
    if (cached[g.epi][CL] == 0) {
      cfg.gridDim = dim3((unsigned)(max_clusters * CL));
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) == cudaSuccess && n > 0) cached[g.epi][CL] = n;
      else { cudaGetLastError(); cached[g.epi][CL] = max_clusters; }
    }
    if (cached[g.epi][CL] < max_clusters) max_clusters = cached[g.epi][CL];
  }
  const int clusters = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;   // persistent: at most one CTA per SM
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  RGNN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, p));
  RGNN_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return RGNN_OK;
}

}  // namespace rgnn

#ifdef RGNN_GEMM_TRACE
extern "C" __attribute__((visibility("default"))) int rgnn_debug_gemm_trace(long long* host_out, int count) {
  if (host_out == nullptr || count <= 0 || count > 160 * 64) return -1;
  return cudaMemcpyFromSymbol(host_out, rgnn::g_gemm_trace, sizeof(long long) * (size_t)count) == cudaSuccess ? 0 : -2;
}
#endif
