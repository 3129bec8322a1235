// gemm_tf32x3.cu -- C = epilogue([A1|A2] . [B1;B2]) with fp32 accuracy on the tensor cores.
//
// Replaces every node-level tf.keras.layers.Dense / tf.layers.Dense / GRUCell matmul of the
// reference hot path (gnns/rgcn.py:71,98; ggnn.py:61,92; rgat.py:70,95; gnn_film.py:70,75,102;
// utils/utils.py:112-125).
//
// Precision: the parity bar is 1e-4 max-norm against the fp32 reference; a single TF32 pass
// (10-bit mantissa) misses it, so every operand is split x = hi + lo with hi = x & 0xffffe000
// (exactly representable in TF32) and lo = x - hi (exact in fp32), and three MMAs accumulate
// lo*hi + hi*lo + hi*hi in fp32 -- relative error ~2^-21, i.e. fp32-level.
//
// Tiling: BM x 128 x 32 CTA tiles, 8 warps (2 x 4), mma.sync.m16n8k8.tf32, 3-stage cp.async
// pipeline, padded shared tiles (conflict-free fragment reads).  The A/B loaders take two K
// segments so [m | h] . [W ; U] (GRU / RNN cell) is one kernel, and blockIdx.z batches over edge
// types (shared A, per-type B: T = H . [W_0 | ... | W_{L-1}]).
#include "gemm.cuh"

namespace rgnn {

namespace {

constexpr int BK = 32;
constexpr int BN = 128;
constexpr int NTHREADS = 256;
constexpr int STAGES = 3;
constexpr int LDA_S = BK + 4;    // 36: (gid*36 + tig) % 32 distinct over a warp
constexpr int LDB_S = BN + 8;    // 136: (tig*136 + gid) % 32 distinct over a warp

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src, bool valid) {
  unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  int src_bytes = valid ? 16 : 0;   // src-size 0 => zero fill, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int BM>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_tf32x3_kernel(const __grid_constant__ GemmParams p) {
  constexpr int MT = BM / 32;   // m16 tiles per warp (warp rows = BM/2)
  constexpr int NT = 4;         // n8 tiles per warp (warp cols = 32)
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                                   // [STAGES][BM][LDA_S]
  float* Bs = smem + STAGES * BM * LDA_S;             // [STAGES][BK][LDB_S]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;

  // ---- resolve this CTA's operands (batching over edge types) ----
  const int z = blockIdx.z;
  const float* A1 = p.A1;
  const float* A2 = p.A2;
  const float* B1 = p.B1;
  const float* B2 = p.B2;
  float* C = p.C;
  int row_begin = 0, row_end = p.M;
  if (p.batch_mode == BATCH_SHARED_A) {
    B1 = p.bptr[z]; B2 = p.bptr2[z];
    C += (size_t)z * p.N;
  } else if (p.batch_mode == BATCH_ROW_RANGES) {
    B1 = p.bptr[z]; B2 = p.bptr2[z];
    row_begin = p.row_off[z]; row_end = p.row_off[z + 1];
  } else if (p.batch_mode == BATCH_COL_BLOCKS) {
    B1 = p.bptr[z]; B2 = p.bptr2[z];
    A1 += (size_t)z * p.K1;
    C += (size_t)z * p.N;
  }
  const int m0 = row_begin + blockIdx.y * BM;
  if (m0 >= row_end) return;
  const int n0 = blockIdx.x * BN;
  const int K1 = p.K1, K = p.K1 + p.K2;
  const int nk = (K + BK - 1) / BK;

  auto load_stage = [&](int stage, int kb) {
    float* as = As + stage * BM * LDA_S;
    float* bs = Bs + stage * BK * LDB_S;
    const int k0 = kb * BK;
#pragma unroll
    for (int i = 0; i < BM * 8 / NTHREADS; ++i) {
      const int c = tid + i * NTHREADS;
      const int row = c >> 3, kc = (c & 7) << 2;
      const int gk = k0 + kc, grow = m0 + row;
      const bool valid = (grow < row_end) && (gk < K);
      const float* src = A1;
      if (valid) src = (gk < K1) ? A1 + (size_t)grow * p.lda1 + gk : A2 + (size_t)grow * p.lda2 + (gk - K1);
      cp_async16(as + row * LDA_S + kc, src, valid);
    }
#pragma unroll
    for (int i = 0; i < BK * (BN / 4) / NTHREADS; ++i) {
      const int c = tid + i * NTHREADS;
      const int krow = c >> 5, nc = (c & 31) << 2;
      const int gk = k0 + krow, gn = n0 + nc;
      const bool valid = (gk < K) && (gn < p.N);
      const float* src = B1;
      if (valid) src = (gk < K1) ? B1 + (size_t)gk * p.ldb1 + gn : B2 + (size_t)(gk - K1) * p.ldb2 + gn;
      cp_async16(bs + krow * LDB_S + nc, src, valid);
    }
  };

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.0f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  for (int kb = 0; kb < nk; ++kb) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kb + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    const float* as = As + (kb % STAGES) * BM * LDA_S + (wm * (BM / 2)) * LDA_S;
    const float* bs = Bs + (kb % STAGES) * BK * LDB_S + wn * 32;
#pragma unroll
    for (int kk = 0; kk < BK / 8; ++kk) {
      uint32_t ahi[MT][4], alo[MT][4], bhi[NT][2], blo[NT][2];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float* ap = as + (mt * 16 + gid) * LDA_S + kk * 8 + tig;
        split_tf32(ap[0], ahi[mt][0], alo[mt][0]);
        split_tf32(ap[8 * LDA_S], ahi[mt][1], alo[mt][1]);
        split_tf32(ap[4], ahi[mt][2], alo[mt][2]);
        split_tf32(ap[8 * LDA_S + 4], ahi[mt][3], alo[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float* bp = bs + (kk * 8 + tig) * LDB_S + nt * 8 + gid;
        split_tf32(bp[0], bhi[nt][0], blo[nt][0]);
        split_tf32(bp[4 * LDB_S], bhi[nt][1], blo[nt][1]);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          mma_tf32(acc[mt][nt], alo[mt], bhi[nt]);   // small terms first
          mma_tf32(acc[mt][nt], ahi[mt], blo[nt]);
          mma_tf32(acc[mt][nt], ahi[mt], bhi[nt]);
        }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue ----
  const int dgru = (p.epi == EPI_GRU_ZR) ? p.N / 2 : p.N;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = m0 + wm * (BM / 2) + mt * 16 + gid + half * 8;
      if (r >= row_end) continue;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = n0 + wn * 32 + nt * 8 + 2 * tig;
        if (c >= p.N) continue;
        float v0 = acc[mt][nt][half * 2 + 0], v1 = acc[mt][nt][half * 2 + 1];
        if (p.bias != nullptr) { v0 += __ldg(p.bias + c); v1 += __ldg(p.bias + c + 1); }
        if (p.epi == EPI_STORE) {
          v0 = apply_act(v0, p.act); v1 = apply_act(v1, p.act);
          *reinterpret_cast<float2*>(C + (size_t)r * p.ldc + c) = make_float2(v0, v1);
        } else if (p.epi == EPI_GRU_ZR) {
          v0 = hard_sigmoid(v0); v1 = hard_sigmoid(v1);
          if (c < dgru) {
            *reinterpret_cast<float2*>(C + (size_t)r * p.ldc + c) = make_float2(v0, v1);
          } else {
            const int cc = c - dgru;
            const float2 h = *reinterpret_cast<const float2*>(p.aux_h + (size_t)r * p.ld_h + cc);
            *reinterpret_cast<float2*>(p.C2 + (size_t)r * p.ldc2 + cc) = make_float2(v0 * h.x, v1 * h.y);
          }
        } else {  // EPI_GRU_OUT
          v0 = apply_act(v0, p.act); v1 = apply_act(v1, p.act);
          const float2 h = *reinterpret_cast<const float2*>(p.aux_h + (size_t)r * p.ld_h + c);
          const float2 zz = *reinterpret_cast<const float2*>(p.aux_z + (size_t)r * p.ld_z + c);
          *reinterpret_cast<float2*>(C + (size_t)r * p.ldc + c) =
              make_float2(zz.x * h.x + (1.0f - zz.x) * v0, zz.y * h.y + (1.0f - zz.y) * v1);
        }
      }
    }
  }
}

template <int BM>
int launch_bm(const GemmParams& p, int rows, cudaStream_t stream) {
  const size_t smem = (size_t)STAGES * (BM * LDA_S + BK * LDB_S) * sizeof(float);
  static bool attr_set = false;   // per template instantiation
  if (!attr_set) {
    RGNN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (rows + BM - 1) / BM, p.batch);
  gemm_tf32x3_kernel<BM><<<grid, NTHREADS, smem, stream>>>(p);
  RGNN_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return RGNN_OK;
}

}  // namespace

int launch_gemm(const GemmParams& p, cudaStream_t stream) {
  RGNN_REQUIRE(p.M >= 0 && p.N > 0 && p.K1 > 0 && p.K2 >= 0, "gemm: bad dims M=%d N=%d K1=%d K2=%d", p.M, p.N, p.K1, p.K2);
  if (p.batch_mode == BATCH_K_BLOCKS_T) {
    set_error("gemm: the legacy mma.sync kernel has no transposed-weight mode (unset RGNN_GEMM_IMPL=mma)");
    return RGNN_E_UNSUPPORTED;
  }
  RGNN_REQUIRE((p.N % 4) == 0 && (p.K1 % 4) == 0 && (p.K2 % 4) == 0, "gemm: N, K must be multiples of 4 (N=%d K1=%d K2=%d)", p.N, p.K1, p.K2);
  RGNN_REQUIRE((p.lda1 % 4) == 0 && (p.ldb1 % 4) == 0 && (p.ldc % 2) == 0, "gemm: leading dims must keep 16-byte rows");
  RGNN_REQUIRE(p.batch >= 1 && p.batch <= RGNN_MAX_EDGE_TYPES, "gemm: batch %d out of range", p.batch);
  RGNN_REQUIRE(aligned16(p.A1) && aligned16(p.C) && (p.K2 == 0 || aligned16(p.A2)), "gemm: operands must be 16-byte aligned");
  if (p.batch_mode == BATCH_NONE) {
    RGNN_REQUIRE(aligned16(p.B1) && (p.K2 == 0 || aligned16(p.B2)), "gemm: B must be 16-byte aligned");
  } else {
    for (int i = 0; i < p.batch; ++i)
      RGNN_REQUIRE(p.bptr[i] != nullptr && aligned16(p.bptr[i]), "gemm: per-type weight %d is NULL or misaligned", i);
  }
  const int rows = (p.batch_mode == BATCH_ROW_RANGES) ? p.max_rows : p.M;
  if (rows <= 0) return RGNN_OK;
  const long tiles128 = (long)((rows + 127) / 128) * ((p.N + BN - 1) / BN) * p.batch;
  if (tiles128 >= 2 * 148) return launch_bm<128>(p, rows, stream);
  return launch_bm<64>(p, rows, stream);
}

}  // namespace rgnn
