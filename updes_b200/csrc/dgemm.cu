// FP64 trailing update of the blocked LU:  C -= A * B  on sub-blocks of one row-major matrix.
//
// This is the only dense contraction on the path (the reference reaches it through
// jnp.linalg.inv + `diffMat @ inv_A`, updes/assembly.py:90,:399, and lineax QR, operators.py:612).
// Blackwell's tcgen05 has no FP64 kind, so the contraction runs on the legacy FP64 tensor path:
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), measured 37.0 TFLOP/s raw on B200 (profiles/r01_ceilings.json).
//
// Structure (one persistent CTA per SM, static tile schedule):
//   * tiles are handed out dynamically (atomic counter), so CTAs delayed by co-running kernels take less;
//   * thread 0 doubles as the producer: it issues TMA (cp.async.bulk.tensor) loads of the
//     A tile (128 rows x 16 k) and the B tile (16 k x BN cols) STAGES-1 k-tiles ahead into a
//     shared-memory ring guarded by full/empty mbarriers;
//   * all 8 warps are consumers: each owns a (128/WARPS_M) x 32 accumulator tile in registers and
//     issues DMMAs from shared-memory fragments;
//   * the epilogue subtracts from C with fire-and-forget red.global.add.f64 (each C element belongs to exactly one
//     thread of one launch, so the bits equal a read-modify-write): the read-modify-write form cost MI dependent
//     HBM round trips per tile (~8 us: 30 TF at k = 512 vs 35 TF at k = 8192); measured with RED: 34.4 / 35.7 TF.
// Shared-memory layout: both operands use the 128-byte TMA swizzle.  An A row is 16 doubles
// (128 B); the fragment row order inside each 8-row group is permuted (0,4,1,5,2,6,3,7) and the
// two DMMAs of a k-pair take the even / odd k of each 16-byte chunk, which makes every
// LDS.128 (A) and LDS.64 (B) bank-conflict-free (derivation in DESIGN.md).
#include "lu.cuh"
#include <mutex>

namespace updes {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 16;

struct GemmParams {
  double *C;           // base of the buffer holding C
  long long ld;        // its leading dimension
  long long rc, cc;    // C block origin
  long long ra, ca;    // A block origin (m x k)
  long long rb, cb;    // B block origin (k x n)
  long long m, n, k;
  unsigned int *counter;       // dynamic tile scheduler: next tile index of THIS launch (starts at 0)
  unsigned int *next_counter;  // counter of the next launch on the stream, reset by this one
  int c_prefetch;              // 1: pull the C tile into L2 one pipeline stage before the epilogue
  int epilogue;                // 1: red.global.add (no C read by the SM; default); 0: read-modify-write of C
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Tile order: groups of GROUP_M row-tiles swept across all column-tiles, so one wave of CTAs
// touches ~GROUP_M A strips and ~num_sms/GROUP_M B strips (L2 reuse) instead of num_sms A strips.
constexpr int GEMM_GROUP_M = 16;
__device__ __forceinline__ void tile_coords(long long t, int tiles_m, int tiles_n, int &mt, int &nt) {
  const long long per_group = (long long)GEMM_GROUP_M * tiles_n;
  const int group = (int)(t / per_group);
  const int first_m = group * GEMM_GROUP_M;
  const int gsz = min(GEMM_GROUP_M, tiles_m - first_m);
  const int rem = (int)(t - (long long)group * per_group);
  mt = first_m + rem % gsz;
  nt = rem / gsz;
}

// NW = 8: one CTA per SM, 128 x BN tile (BN = 32 / 64 / 128).
// NW = 4: "ping-pong" -- two independent 4-warp CTAs per SM, each with a 128 x 64 tile and the same
//         64 x 32 per-warp accumulator block: while one CTA runs its epilogue (the exposed C
//         read-modify-write) the other keeps the FP64 tensor pipe busy.
// KSUB = 16-deep k sub-tiles per pipeline stage: 2 halves the number of barrier rounds per flop.
template <int BN, int NW, int KSUB>
struct GemmCfg {
  static constexpr int THREADS = NW * 32;
  static constexpr int MIN_CTAS = NW == 4 ? 2 : 1;
  static constexpr int WARPS_N = BN / 32;
  static constexpr int WARPS_M = NW / WARPS_N;
  static constexpr int WM = GEMM_BM / WARPS_M;     // rows per warp
  static constexpr int MI = WM / 8;                // 8-row fragments per warp
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 8;
  static constexpr int B_BYTES = BN * GEMM_BK * 8;
  static constexpr int SUB_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGE_BYTES = KSUB * SUB_BYTES;
  static constexpr int STAGES = (NW == 4 ? 4 : (BN == 128 ? 6 : 8)) / KSUB;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int NW, int KSUB>
__global__ void __launch_bounds__(NW * 32, (NW == 4 ? 2 : 1))
dgemm_sub_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmParams P) {
  using Cfg = GemmCfg<BN, NW, KSUB>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;   // full[s] at +8s, empty[s] at +8(STAGES+s)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; s++) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (Cfg::STAGES + s), NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int tiles_m = (int)((P.m + GEMM_BM - 1) / GEMM_BM);
  const int tiles_n = (int)((P.n + BN - 1) / BN);
  const long long ntiles = (long long)tiles_m * tiles_n;
  const int ktiles = (int)(P.k / (GEMM_BK * KSUB));

  // ------------------------------ producer state (thread 0) ------------------------------
  // No dedicated producer warp: a 9th warp would round the CTA up to 12 warps of register
  // allocation (168 regs/thread) and spill the 128 accumulator registers.  Thread 0 issues the TMA
  // loads STAGES-1 k-tiles ahead of the consumers, all warps (<= 255 regs) do the math.
  //
  // Tiles are handed out dynamically (one atomicAdd per tile on a per-launch counter) rather than
  // round-robin: when other kernels hold some SMs for a while -- NCCL's broadcast on the multi-GPU
  // path -- the CTAs that start late simply take fewer tiles instead of stretching the tail.
  const uint32_t stage_tile = bars + 16 * Cfg::STAGES;   // int[STAGES]: tile id the stage belongs to (-1 = no more work)
  long long p_t = -1, p_next = -1;
  int p_kt = 0, p_stage = 0;
  uint32_t p_phase = 0;
  bool p_done = false;
  auto fetch_tile = [&]() -> long long {
    const unsigned int t = atomicAdd(P.counter, 1u);
    return (long long)t < ntiles ? (long long)t : -1;
  };
  auto produce_one = [&]() {
    if (p_done) return;
    if (p_kt == 0) {
      p_t = p_next;
      if (p_t >= 0) p_next = fetch_tile();     // requested a whole tile ahead of its first use
    }
    mbar_wait(bars + 8 * (Cfg::STAGES + p_stage), p_phase ^ 1);
    const uint32_t full = bars + 8 * p_stage;
    if (p_t < 0) {                               // sentinel stage: tells the consumers to stop
      asm volatile("st.shared.s32 [%0], %1;" ::"r"(stage_tile + 4 * p_stage), "r"(-1) : "memory");
      mbar_arrive(full);
      p_done = true;
      return;
    }
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(stage_tile + 4 * p_stage), "r"((int)p_t) : "memory");
    int mt, nt;
    tile_coords(p_t, tiles_m, tiles_n, mt, nt);
    const int arow = (int)(P.ra + (long long)mt * GEMM_BM);
    const int bcol = (int)(P.cb + (long long)nt * BN);
    mbar_expect_tx(full, Cfg::STAGE_BYTES);
#pragma unroll
    for (int ks = 0; ks < KSUB; ks++) {
      const uint32_t sa = smem_base + p_stage * Cfg::STAGE_BYTES + ks * Cfg::SUB_BYTES;
      const int kk = (p_kt * KSUB + ks) * GEMM_BK;
      tma_load_2d(sa, &mapA, (int)(P.ca + kk), arow, full);
#pragma unroll
      for (int g = 0; g < BN / 16; g++)
        tma_load_2d(sa + Cfg::A_BYTES + g * 2048, &mapB, bcol + 16 * g, (int)(P.rb + kk), full);
    }
    if (++p_stage == Cfg::STAGES) { p_stage = 0; p_phase ^= 1; }
    if (++p_kt == ktiles) p_kt = 0;
  };
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) *P.next_counter = 0u;
    p_next = fetch_tile();
#pragma unroll 1
    for (int i = 0; i < Cfg::STAGES - 1; i++) produce_one();
  }

  // ------------------------------ consumers ------------------------------
  const int warp_m = warp % Cfg::WARPS_M, warp_n = warp / Cfg::WARPS_M;
  const int fr = lane >> 2, s4 = lane & 3;
  const int prow = (fr >> 1) | ((fr & 1) << 2);                 // permuted row inside an 8-row group
  // A: byte offset of (row = warp_m*WM + i*8 + prow, chunk) inside the stage = row*128 + ((4h+s4)^prow)*16
  const uint32_t a_row_off = (uint32_t)(warp_m * Cfg::WM + prow) * 128u;
  const uint32_t a_ch0 = (uint32_t)((s4 ^ prow) << 4), a_ch1 = (uint32_t)(((4 + s4) ^ prow) << 4);
  // B: element (k = 8h + 2 s4 + p, n = warp_n*32 + jn*8 + fr) lives at
  //    (n>>4)*2048 + k*128 + (((n&15)>>1) ^ (k&7))*16 + (n&1)*8.
  // With n&15 = (jn&1)*8 + fr the chunk index is ((fr>>1) ^ k7) ^ ((jn&1)<<2), so the four jn
  // offsets are  (base_p ^ ((jn&1)<<6)) + (jn>>1)*2048 : two base registers per parity p.
  uint32_t b_base[2][2];
#pragma unroll
  for (int p = 0; p < 2; p++) {
    const int k7 = 2 * s4 + p;
    const uint32_t base = (uint32_t)(warp_n * 2 * 2048 + k7 * 128 + ((((fr >> 1) ^ k7)) << 4) + (fr & 1) * 8);
    b_base[p][0] = base;
    b_base[p][1] = base ^ 64u;
  }

  int stage = 0;
  uint32_t phase = 0;
  while (true) {
    if (threadIdx.x == 0) produce_one();
    mbar_wait(bars + 8 * stage, phase);
    int t;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(t) : "r"(stage_tile + 4 * stage) : "memory");
    if (t < 0) break;
    int mt, nt;
    tile_coords(t, tiles_m, tiles_n, mt, nt);
    double acc[Cfg::MI][4][2];
#pragma unroll
    for (int i = 0; i < Cfg::MI; i++)
#pragma unroll
      for (int jn = 0; jn < 4; jn++) { acc[i][jn][0] = 0.0; acc[i][jn][1] = 0.0; }

    for (int kt = 0; kt < ktiles; kt++) {
      if (kt > 0) {
        if (threadIdx.x == 0) produce_one();
        mbar_wait(bars + 8 * stage, phase);
      }
      if (kt == ktiles - 1 && P.c_prefetch) {
        // The epilogue reads the C tile in MI dependent rounds (the accumulators leave no registers to batch them).
        // C was last written by an earlier launch, so each round would be a full HBM round trip (~8 us per tile,
        // the fixed per-tile cost behind 30 TF at k = 512 vs 35 TF at k = 8192): pull the tile into L2 one
        // pipeline stage ahead instead.
        constexpr int LPR = BN / 16, PER = GEMM_BM * LPR / Cfg::THREADS;
#pragma unroll
        for (int u = 0; u < PER; u++) {
          const int idx = threadIdx.x * PER + u;
          const long long row = (long long)mt * GEMM_BM + idx / LPR, col = (long long)nt * BN + (idx % LPR) * 16;
          if (row < P.m && col < P.n)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.C + (P.rc + row) * P.ld + P.cc + col));
        }
      }
#pragma unroll
      for (int ks = 0; ks < KSUB; ks++) {
      const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES + ks * Cfg::SUB_BYTES + a_row_off;
      const uint32_t sb = smem_base + stage * Cfg::STAGE_BYTES + ks * Cfg::SUB_BYTES + Cfg::A_BYTES;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        double b_ev[4], b_od[4];
#pragma unroll
        for (int jn = 0; jn < 4; jn++) {
          const uint32_t o = (uint32_t)((jn >> 1) * 2048 + h * 1024);
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b_ev[jn]) : "r"(sb + b_base[0][jn & 1] + o));
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b_od[jn]) : "r"(sb + b_base[1][jn & 1] + o));
        }
#pragma unroll
        for (int i = 0; i < Cfg::MI; i++) {
          double a_ev, a_od;
          const uint32_t addr = sa + (uint32_t)i * 1024u + (h ? a_ch1 : a_ch0);
          asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a_ev), "=d"(a_od) : "r"(addr));
#pragma unroll
          for (int jn = 0; jn < 4; jn++) {
            dmma(acc[i][jn][0], acc[i][jn][1], a_ev, b_ev[jn]);
            dmma(acc[i][jn][0], acc[i][jn][1], a_od, b_od[jn]);
          }
        }
      }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (Cfg::STAGES + stage));
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }

    // epilogue: C -= acc
    const long long row_base = (long long)mt * GEMM_BM + warp_m * Cfg::WM + prow;
    const long long col_base = (long long)nt * BN + warp_n * 32 + 2 * s4;
    if (P.epilogue == 1) {
      // fire-and-forget reductions: every C element is touched by exactly one thread of one launch, so
      // red.add(C, -acc) gives the same bits as C - acc, and the SM never waits for the C tile
#pragma unroll
      for (int i = 0; i < Cfg::MI; i++) {
        const long long r = row_base + i * 8;
        if (r >= P.m) continue;
        double *crow = P.C + (P.rc + r) * P.ld + P.cc;
#pragma unroll
        for (int jn = 0; jn < 4; jn++) {
          const long long c = col_base + jn * 8;
          if (c < P.n) asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(crow + c), "d"(-acc[i][jn][0]) : "memory");
          if (c + 1 < P.n) asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(crow + c + 1), "d"(-acc[i][jn][1]) : "memory");
        }
      }
      continue;
    }
    // read-modify-write, software-pipelined one row group deep (the loads of group i+1 are in flight while group i
    // is stored; the mainloop's fragment registers are dead here)
    constexpr int EPI_DEPTH = (NW == 4) ? 2 : 1;      // the 8-warp variants have no registers to spare
    double2 cv[EPI_DEPTH][4];
    auto load_c = [&](int i, double2 (&v)[4]) {
      const long long r = row_base + i * 8;
      if (r >= P.m) return;
      const double *crow = P.C + (P.rc + r) * P.ld + P.cc;
#pragma unroll
      for (int jn = 0; jn < 4; jn++) {
        const long long c = col_base + jn * 8;
        if (c + 1 < P.n) v[jn] = *reinterpret_cast<const double2 *>(crow + c);
        else if (c < P.n) v[jn] = make_double2(crow[c], 0.0);
      }
    };
    if (EPI_DEPTH == 2) load_c(0, cv[0]);
#pragma unroll
    for (int i = 0; i < Cfg::MI; i++) {
      if (EPI_DEPTH == 1) load_c(i, cv[0]);
      else if (i + 1 < Cfg::MI) load_c(i + 1, cv[(i + 1) % EPI_DEPTH]);
      const long long r = row_base + i * 8;
      if (r >= P.m) continue;
      double *crow = P.C + (P.rc + r) * P.ld + P.cc;
#pragma unroll
      for (int jn = 0; jn < 4; jn++) {
        const long long c = col_base + jn * 8;
        const double2 v = cv[i % EPI_DEPTH][jn];
        if (c + 1 < P.n) {
          *reinterpret_cast<double2 *>(crow + c) = make_double2(v.x - acc[i][jn][0], v.y - acc[i][jn][1]);
        } else if (c < P.n) {
          crow[c] = v.x - acc[i][jn][0];
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// Host side: tensor maps and launch
// --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int lu_bind_view(UpdesLU *h, int slot, const double *ptr, int64_t rows, int64_t ld_) {
  if (slot < 0 || slot >= UPDES_MAX_VIEWS) return -2;
  MatView &V = h->view[slot];
  if (V.ptr == ptr && V.rows == rows && V.ld == ld_) return 0;
  if (!ptr || rows <= 0 || ld_ < 16 || (ld_ % 16) || (((uintptr_t)ptr) & 127)) return -3;
  EncodeTiledFn enc = get_encode();
  if (!enc) return (int)cudaErrorNotSupported;
  const cuuint64_t ld = (cuuint64_t)ld_, n = (cuuint64_t)rows;
  cuuint64_t dims[2] = {ld, n};
  cuuint64_t strides[1] = {ld * 8};
  cuuint32_t estr[2] = {1, 1};
  {
    // left operand: box 16 columns (k) x 128 rows
    cuuint32_t box[2] = {GEMM_BK, GEMM_BM};
    CUresult r = enc(&V.mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
  }
  {
    // right operand: box 16 columns x 16 k-rows; a BN-wide tile is BN/16 such boxes laid out
    // [column group][k][16 columns], which keeps the k rows 128 B apart (conflict-free LDS.64)
    cuuint32_t box[2] = {16, GEMM_BK};
    CUresult r = enc(&V.mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
  }
  V.ptr = const_cast<double *>(ptr); V.rows = rows; V.ld = ld_;
  return 0;
}

template <int BN, int NW, int KSUB>
static int launch_gemm(UpdesLU *h, const MatView &VA, const MatView &VB, const GemmParams &P, cudaStream_t st) {
  using Cfg = GemmCfg<BN, NW, KSUB>;
  static bool attr_set = false;
  if (!attr_set) {
    UPDES_CUDA_TRY(cudaFuncSetAttribute(dgemm_sub_kernel<BN, NW, KSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const long long tiles = ((P.m + GEMM_BM - 1) / GEMM_BM) * ((P.n + BN - 1) / BN);
  const int cap = ((h->gemm_ctas > 0 && h->gemm_ctas < h->num_sms) ? h->gemm_ctas : h->num_sms) * Cfg::MIN_CTAS;
  const int grid = (int)(tiles < cap ? tiles : cap);
  GemmParams Q = P;
  Q.counter = h->gemm_counters + (h->gemm_launch_id % UPDES_GEMM_COUNTERS);
  Q.next_counter = h->gemm_counters + ((h->gemm_launch_id + 1) % UPDES_GEMM_COUNTERS);
  h->gemm_launch_id++;
  Q.c_prefetch = h->gemm_c_prefetch && h->gemm_epilogue == 0;
  Q.epilogue = h->gemm_epilogue;
  prof_begin(PROF_GEMM, 2.0 * (double)P.m * (double)P.n * (double)P.k, st);
  dgemm_sub_kernel<BN, NW, KSUB><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(VA.mapA, VB.mapB, Q);
  prof_end(st);
  UPDES_LAUNCH_CHECK();
  return 0;
}

int dgemm_sub(UpdesLU *h, int va, int64_t ra, int64_t ca, int vb, int64_t rb, int64_t cb, int vc, int64_t rc,
              int64_t cc, int64_t m, int64_t n, int64_t k, cudaStream_t st) {
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  if (k % GEMM_BK) return -11;
  if (cc & 1) return -4;   // 16-byte C accesses; TMA boxes may start at any column
  const MatView &VA = h->view[va], &VB = h->view[vb], &VC = h->view[vc];
  if (!VA.ptr || !VB.ptr || !VC.ptr) return -2;
  GemmParams P;
  P.C = VC.ptr; P.ld = VC.ld; P.rc = rc; P.cc = cc; P.ra = ra; P.ca = ca; P.rb = rb; P.cb = cb; P.m = m; P.n = n; P.k = k;
  if (n <= 32) return launch_gemm<32, 8, 1>(h, VA, VB, P, st);
  if (n <= 64) return launch_gemm<64, 8, 1>(h, VA, VB, P, st);
  // wide updates: ping-pong (two 128x64 CTAs per SM) once there are enough tiles to fill every slot;
  // 32-deep pipeline stages (KSUB = 2) when k allows and gemm_kdeep is set
  const long long tiles64 = ((m + GEMM_BM - 1) / GEMM_BM) * ((n + 63) / 64);
  const bool deep = h->gemm_kdeep && (k % (2 * GEMM_BK) == 0) && k >= 128;
  if (h->gemm_variant == 1 && tiles64 >= 2LL * h->num_sms)
    return deep ? launch_gemm<64, 4, 2>(h, VA, VB, P, st) : launch_gemm<64, 4, 1>(h, VA, VB, P, st);
  return deep ? launch_gemm<128, 8, 2>(h, VA, VB, P, st) : launch_gemm<128, 8, 1>(h, VA, VB, P, st);
}

}  // namespace updes

extern "C" int updes_dgemm_sub(UpdesLU *handle, double *K, int64_t rc, int64_t cc, int64_t ra, int64_t ca,
                               int64_t rb, int64_t cb, int64_t m, int64_t n, int64_t k, void *stream) {
  if (!handle) return -1;
  if (!K) return -2;
  int rc_ = updes::lu_bind_view(handle, 0, K, handle->n, handle->ld);
  if (rc_) return rc_;
  return updes::dgemm_sub(handle, 0, ra, ca, 0, rb, cb, 0, rc, cc, m, n, k, (cudaStream_t)stream);
}

extern "C" int updes_lu_gemm(UpdesLU *handle, int slot_a, int64_t ra, int64_t ca, int slot_b, int64_t rb, int64_t cb,
                             int slot_c, int64_t rc, int64_t cc, int64_t m, int64_t n, int64_t k, void *stream) {
  if (!handle) return -1;
  if (slot_a < 0 || slot_a >= UPDES_MAX_VIEWS) return -2;
  if (slot_b < 0 || slot_b >= UPDES_MAX_VIEWS) return -5;
  if (slot_c < 0 || slot_c >= UPDES_MAX_VIEWS) return -8;
  return updes::dgemm_sub(handle, slot_a, ra, ca, slot_b, rb, cb, slot_c, rc, cc, m, n, k, (cudaStream_t)stream);
}

extern "C" int updes_lu_set_gemm_ctas(UpdesLU *handle, int ctas) {
  if (!handle) return -1;
  if (ctas < 0) return -2;
  handle->gemm_ctas = ctas;
  return 0;
}

extern "C" int updes_lu_set_gemm_variant(UpdesLU *handle, int variant) {
  if (!handle) return -1;
  if (variant < 0 || variant > 15) return -2;
  handle->gemm_variant = variant & 1;      // bit 0: ping-pong schedule
  handle->gemm_kdeep = (variant >> 1) & 1; // bit 1: 32-deep pipeline stages
  handle->gemm_c_prefetch = ((variant >> 2) & 1) ^ 1;   // bit 2: 1 = do NOT prefetch the C tile before the epilogue
  handle->gemm_epilogue = ((variant >> 3) & 1) ^ 1;     // bit 3: 1 = read-modify-write epilogue instead of red.global.add
  return 0;
}

/* test hook: pretend the 32-wide panel kernel holds only `rows` rows, to exercise the 16- and 8-wide
 * base panels that very tall (multi-GPU, > 94 720-row) panels use */
extern "C" int updes_lu_set_panel_capacity(UpdesLU *handle, int64_t rows) {
  if (!handle) return -1;
  if (rows < 0) return -2;
  handle->panel_cap = rows;
  return 0;
}

/* tuning hook: largest block (rows, multiple of 32, <= 128) of the unit-lower solve handled by one substitution
 * kernel; 32 restores the first-generation recursion down to 32-row blocks */
extern "C" int updes_lu_set_trsm_base(UpdesLU *handle, int rows) {
  if (!handle) return -1;
  if (rows < 32 || rows > 128 || (rows % 32)) return -2;
  handle->trsm_base_rows = rows;
  return 0;
}

extern "C" int updes_lu_set_panel_variant(UpdesLU *handle, int variant) {
  if (!handle) return -1;
  if (variant < 0 || variant > 2) return -2;
  handle->panel_variant = variant;
  return 0;
}
