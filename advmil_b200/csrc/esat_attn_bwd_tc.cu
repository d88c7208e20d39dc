// ESAT self-attention BACKWARD on the Blackwell tensor path (tcgen05 / TMEM / TMA): one kernel template, two instantiations.
//   KV = false  "dQ pass"    rows = 128 queries of one (bag, head), streamed column tiles = 64 keys
//   KV = true   "dK/dV pass" rows = 128 keys,                        streamed column tiles = 64 queries
// Per column tile (flash style: the logits are recomputed, nothing of size R x R touches HBM):
//   T1 = R1 C1^T   (S = Q K^T,  or S^T = K Q^T)        tcgen05.mma kind::tf32, M = 128, N = 64 -> TMEM
//   T2 = R2 C2^T   (dP = dO V^T, or dP^T = V dO^T)                                             -> TMEM
//   P  = 2^(T1 c - lse),  Pd = keep P / (1-p),  dS = P (keep T2 / (1-p) - D)   one row per thread pair, written back over T1
//        (dS) and T2 (Pd) IN TENSOR MEMORY with tcgen05.st
//   ACC1 += dS C1    (dQ += dS K,  or dK += dS^T Q)    A = dS from tensor memory, B = the column tile read MN-major
//   ACC2 += Pd C2    (dV += Pd^T dO; KV only)          accumulators stay in TMEM over the whole column loop
// R1 / R2 (128 x hd, K-major) are loaded once; every column tile arrives by TMA in the layouts its two roles need: K-major
// (128-byte swizzle, 16-byte atoms) as the B operand of T1 / T2, and MN-major (32-byte atoms: the only swizzle the hardware has
// for MN-major 32-bit operands) as the B operand of the accumulating contractions.  Same ping-pong as the forward kernel:
// two groups of 8 element-wise warps own the even / odd column tiles, each with its own T1 / T2 columns (2 x 128 of the 512 TMEM
// columns; ACC1 / ACC2 take 2 x 64), so one group computes while the other waits for its MMAs and the next tile's TMA.
// D_i = dO_i . O_i is computed by the dQ pass (rows = queries) and handed to the dK/dV pass through `Dq` [heads, R].
// Dropout: the forward's generator (row = region * heads + head, column = key); in the dK/dV pass a thread owns ONE key, so
// the two lanes of a key pair share each 32-bit draw through a shuffle.
// Warp roles (608 threads): warps 0-15 element-wise, warp 16 MMA issuer, warp 17 TMA producer.
#include <stdlib.h>
#include "stages.cuh"
#include "tc_ptx.cuh"
#include "esat_attn.cuh"

namespace advmil {
using namespace tc;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// row-major fp32 matrix [rows, cols]; box = box_rows x 32 columns (128 bytes)
int make_map(CUtensorMap* m, const float* base, long long rows, long long cols, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ADVMIL_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("attention bwd: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld", (int)r, rows, cols); return ADVMIL_ERR_CUDA; }
  return ADVMIL_OK;
}

constexpr int BT_ROWS = 128;                 // rows per CTA
constexpr int BT_COLS = 64;                  // streamed columns per tile
constexpr int BT_EW = 16;                    // element-wise warps: 2 groups x 2 column halves x 4 lane quarters
constexpr int BT_THREADS = 32 * (BT_EW + 2);
constexpr int BT_RBLK = BT_ROWS * 128;       // bytes of one [128 rows x 128 B] block
constexpr int BT_CBLK = BT_COLS * 128;       // bytes of one [64 rows x 128 B] block
constexpr float kLog2eB = 1.4426950408889634f;

template <int HD, bool KV> struct BwdCfg {
  static_assert(HD % 16 == 0 && HD >= 16 && HD <= 64, "head width");
  static constexpr int NB = (HD + 31) / 32;
  static constexpr int R_BYTES = NB * BT_RBLK;                         // one resident operand
  static constexpr int C_BYTES = NB * BT_CBLK;                         // one column-tile operand in one layout
  static constexpr int NCOP = KV ? 4 : 3;                              // C1 (K-major), C1 (MN-major), C2 (K-major)[, C2 (MN-major)]
  static constexpr int STAGE_BYTES = NCOP * C_BYTES;
  static constexpr size_t USED = 2 * (size_t)R_BYTES + 2 * (size_t)STAGE_BYTES + 16 * 8;
  static constexpr size_t SMEM = USED + 1024;
  static constexpr int OQ = HD / 4;                                    // accumulator columns per thread at the read-out
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

__device__ __forceinline__ float ex2b(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// tf32 rounding of a finite value: the tensor core drops the low 13 mantissa bits, so half an ulp of the magnitude is added first
__device__ __forceinline__ uint32_t round_tf32(float x) { return __float_as_uint(x) + 0x1000u; }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N> __device__ __forceinline__ void tmem_ldq(uint32_t taddr, float (&v)[N]) {
  static_assert(N == 4 || N == 8 || N == 12 || N == 16, "column count");
  if constexpr (N == 4) tmem_ld4(taddr, v);
  else if constexpr (N == 8) tmem_ld8(taddr, v);
  else if constexpr (N == 12) { tmem_ld8(taddr, v); tmem_ld4(taddr + 8, v + 8); }
  else { tmem_ld8(taddr, v); tmem_ld8(taddr + 8, v + 8); }
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_tf32_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct BwdMaps { CUtensorMap qkvK128, qkvK64, qkvM64, dcK128, dcK64, dcM64; };

template <int HD, bool KV>
__global__ void __launch_bounds__(BT_THREADS, 1)
mha_bwd_tcgen05_kernel(const __grid_constant__ BwdMaps maps, const int32_t* __restrict__ ro, int d, float scale, AttDrop ad,
                       const float* __restrict__ lse, const float* __restrict__ ctx, const float* __restrict__ d_ctx,
                       float* __restrict__ Dq, float* __restrict__ d_qkv, int Rtot) {
  pdl_prologue();
  using Cfg = BwdCfg<HD, KV>;
  constexpr int NB = Cfg::NB, OQ = Cfg::OQ;
  const int b = blockIdx.y, head = blockIdx.z, r0 = ro[b], Rb = ro[b + 1] - r0;
  const int row0 = blockIdx.x * BT_ROWS;
  if (row0 >= Rb) return;
  const int T = (Rb + BT_COLS - 1) / BT_COLS;          // column tiles

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if ((size_t)(smem - smem_raw) + Cfg::USED > Cfg::SMEM) __trap();
  uint8_t* R1_s = smem;
  uint8_t* R2_s = R1_s + Cfg::R_BYTES;
  uint8_t* C_s = R2_s + Cfg::R_BYTES;                  // [2 stages][C1K | C1M | C2K | (C2M)]
  uint64_t* bars = (uint64_t*)(C_s + 2 * Cfg::STAGE_BYTES);
  uint64_t* rfull = bars;            // [1]
  uint64_t* cfull = bars + 1;        // [2]
  uint64_t* cfree = bars + 3;        // [2]
  uint64_t* tfull = bars + 5;        // [2]  T1 / T2 of the group's tile are in TMEM
  uint64_t* pfull = bars + 7;        // [2]  the group has written dS (and Pd) over them (8 warp arrivals)
  uint64_t* afull = bars + 9;        // [1]  the accumulators are complete
  uint32_t* tmem_ptr = (uint32_t*)(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&maps.qkvK128); prefetch_tmap(&maps.qkvK64); prefetch_tmap(&maps.qkvM64);
    prefetch_tmap(&maps.dcK128); prefetch_tmap(&maps.dcK64); prefetch_tmap(&maps.dcM64);
    mbar_init(rfull, 1);
    mbar_init(afull, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&cfull[i], 1); mbar_init(&cfree[i], 1); mbar_init(&tfull[i], 1); mbar_init(&pfull[i], BT_EW / 2); }
    fence_barrier_init();
  }
  if (warp == BT_EW) tmem_alloc(tmem_ptr, 512);        // T1 | T2: 2 groups x 128 columns; ACC1, ACC2: 64 columns each
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr uint32_t T_COL = 0, A1_COL = 256, A2_COL = 320;
  // columns of this head inside the packed projection [q | k | v] and inside d_ctx
  const int cq = head * HD, ck = d + head * HD, cv = 2 * d + head * HD, cg = head * HD;

  if (warp == BT_EW + 1) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      mbar_arrive_expect_tx(rfull, 2 * Cfg::R_BYTES);
#pragma unroll
      for (int g = 0; g < NB; ++g) {
        if (KV) {      // rows = keys: R1 = K, R2 = V
          tma_load_2d(R1_s + g * BT_RBLK, &maps.qkvK128, rfull, ck + 32 * g, r0 + row0);
          tma_load_2d(R2_s + g * BT_RBLK, &maps.qkvK128, rfull, cv + 32 * g, r0 + row0);
        } else {       // rows = queries: R1 = Q, R2 = dO
          tma_load_2d(R1_s + g * BT_RBLK, &maps.qkvK128, rfull, cq + 32 * g, r0 + row0);
          tma_load_2d(R2_s + g * BT_RBLK, &maps.dcK128, rfull, cg + 32 * g, r0 + row0);
        }
      }
      for (int j = 0; j < T; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait(&cfree[s], (uint32_t)(((j >> 1) & 1) ^ 1));
        mbar_arrive_expect_tx(&cfull[s], Cfg::STAGE_BYTES);
        uint8_t* st = C_s + s * Cfg::STAGE_BYTES;
        const int crow = r0 + j * BT_COLS;
#pragma unroll
        for (int g = 0; g < NB; ++g) {
          if (KV) {    // columns = queries: C1 = Q (both layouts), C2 = dO (both layouts)
            tma_load_2d(st + 0 * Cfg::C_BYTES + g * BT_CBLK, &maps.qkvK64, &cfull[s], cq + 32 * g, crow);
            tma_load_2d(st + 1 * Cfg::C_BYTES + g * BT_CBLK, &maps.qkvM64, &cfull[s], cq + 32 * g, crow);
            tma_load_2d(st + 2 * Cfg::C_BYTES + g * BT_CBLK, &maps.dcK64, &cfull[s], cg + 32 * g, crow);
            tma_load_2d(st + 3 * Cfg::C_BYTES + g * BT_CBLK, &maps.dcM64, &cfull[s], cg + 32 * g, crow);
          } else {     // columns = keys: C1 = K (both layouts), C2 = V (K-major)
            tma_load_2d(st + 0 * Cfg::C_BYTES + g * BT_CBLK, &maps.qkvK64, &cfull[s], ck + 32 * g, crow);
            tma_load_2d(st + 1 * Cfg::C_BYTES + g * BT_CBLK, &maps.qkvM64, &cfull[s], ck + 32 * g, crow);
            tma_load_2d(st + 2 * Cfg::C_BYTES + g * BT_CBLK, &maps.qkvK64, &cfull[s], cv + 32 * g, crow);
          }
        }
      }
    }
  } else if (warp == BT_EW) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t IDESC_T = idesc_tf32(BT_ROWS, BT_COLS, 0, 0);
      constexpr uint32_t IDESC_A = idesc_tf32(BT_ROWS, HD, 0, 1);
      const uint32_t r1_addr = smem_u32(R1_s), r2_addr = smem_u32(R2_s);
      auto issue_t = [&](int j) {
        const int s = j & 1;
        mbar_wait(&cfull[s], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        const uint32_t c_addr = smem_u32(C_s + s * Cfg::STAGE_BYTES);
        const uint32_t t1 = tmem_base + T_COL + 128 * s, t2 = t1 + 64;
#pragma unroll
        for (int kk = 0; kk < HD / 8; ++kk) {
          const int g = kk >> 2, k8 = kk & 3;
          mma_tf32(t1, smem_desc_sw128(r1_addr + g * BT_RBLK + k8 * 32, 16, 1024),
                   smem_desc_sw128(c_addr + 0 * Cfg::C_BYTES + g * BT_CBLK + k8 * 32, 16, 1024), IDESC_T, kk != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < HD / 8; ++kk) {
          const int g = kk >> 2, k8 = kk & 3;
          mma_tf32(t2, smem_desc_sw128(r2_addr + g * BT_RBLK + k8 * 32, 16, 1024),
                   smem_desc_sw128(c_addr + 2 * Cfg::C_BYTES + g * BT_CBLK + k8 * 32, 16, 1024), IDESC_T, kk != 0 ? 1u : 0u);
        }
        mma_commit(&tfull[s]);
      };
      mbar_wait(rfull, 0);
      issue_t(0);
      if (T > 1) issue_t(1);
      for (int j = 0; j < T; ++j) {
        const int s = j & 1;
        mbar_wait(&pfull[s], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        const uint32_t c_addr = smem_u32(C_s + s * Cfg::STAGE_BYTES);
        const uint32_t t1 = tmem_base + T_COL + 128 * s, t2 = t1 + 64;
#pragma unroll
        for (int k8 = 0; k8 < BT_COLS / 8; ++k8)     // A: 8 columns of dS in tensor memory; B: 8 rows of C1, MN-major
          mma_tf32_ta(tmem_base + A1_COL, t1 + 8 * k8, smem_desc_sw128(c_addr + 1 * Cfg::C_BYTES + k8 * 1024, BT_CBLK, 512, 1), IDESC_A,
                      (j | k8) != 0 ? 1u : 0u);
        if (KV) {
#pragma unroll
          for (int k8 = 0; k8 < BT_COLS / 8; ++k8)   // A: Pd; B: 8 rows of dO, MN-major
            mma_tf32_ta(tmem_base + A2_COL, t2 + 8 * k8, smem_desc_sw128(c_addr + 3 * Cfg::C_BYTES + k8 * 1024, BT_CBLK, 512, 1), IDESC_A,
                        (j | k8) != 0 ? 1u : 0u);
        }
        mma_commit(&cfree[s]);
        if (j == T - 1) mma_commit(afull);
        if (j + 2 < T) issue_t(j + 2);       // in order behind the accumulating MMAs that read this group's T1 / T2 columns
      }
    }
  } else {
    // ---------------- element-wise warps: group g owns the tiles j = g, g + 2, ...; two threads (column halves) per row ----------------
    const int qt = warp & 3, half = (warp >> 2) & 1, g = warp >> 3, row = qt * 32 + lane;
    const int ri = row0 + row;                          // query (dQ pass) or key (dK/dV pass) of this thread inside the bag
    const int ric = min(ri, Rb - 1);
    const uint32_t lane_base = tmem_base + ((uint32_t)(qt * 32) << 16);
    const uint32_t t1_addr = lane_base + T_COL + 128 * g + 32 * half, t2_addr = t1_addr + 64;
    const float sc2 = scale * kLog2eB, ik = ad.drop.inv_keep;
    float lse_r = 0.f, D_r = 0.f;
    if (!KV) {       // per-row statistics: lse of the query and D = dO . O (all four threads of a row compute it; one writes it)
      lse_r = lse[(size_t)head * Rtot + r0 + ric] * kLog2eB;
      const float* gp = d_ctx + (size_t)(r0 + ric) * d + head * HD;
      const float* op = ctx + (size_t)(r0 + ric) * d + head * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(gp + c), o = *reinterpret_cast<const float4*>(op + c);
        D_r += a.x * o.x + a.y * o.y + a.z * o.z + a.w * o.w;
      }
      if (g == 0 && half == 0 && ri < Rb) Dq[(size_t)head * Rtot + r0 + ri] = D_r;
    }
    const uint32_t drow_q = (uint32_t)(r0 + ric) * (uint32_t)ad.heads + (uint32_t)head;     // dQ pass: the hash row of this query
    for (int j = g; j < T; j += 2) {
      const uint32_t par = (uint32_t)((j >> 1) & 1);
      const int c0 = j * BT_COLS + 32 * half;            // first column (key / query inside the bag) of this thread
      float lse_c = 0.f, D_c = 0.f;
      if (KV) {        // per-column statistics: lane l holds those of column c0 + l, fetched by shuffle below
        const int qc = min(c0 + lane, Rb - 1);
        lse_c = lse[(size_t)head * Rtot + r0 + qc] * kLog2eB;
        D_c = Dq[(size_t)head * Rtot + r0 + qc];
      }
      mbar_wait(&tfull[g], par);
      tc_fence_after();
      float t1[32], t2[32];
      tmem_ld32(t1_addr, t1);
      tmem_ld32(t2_addr, t2);
      uint32_t keep = 0xFFFFFFFFu;                        // bit i: the probability of column c0 + i survives the dropout
      if (ad.drop.active) {
        keep = 0;
        if (ad.mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int cc = min(c0 + i, Rb - 1);
            const bool k = KV ? ad.keep(b, head, cc, ric, Rb, 0) : ad.keep(b, head, ric, cc, Rb, 0);
            keep |= (k ? 1u : 0u) << i;
          }
        } else if (!KV) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            bool ka, kb;
            ad.drop.keep2(drow_q, (uint32_t)(c0 + i), ka, kb);
            keep |= (ka ? 1u : 0u) << i;
            keep |= (kb ? 1u : 0u) << (i + 1);
          }
        } else {
          // the draw of (query qc, key pair of this lane and its neighbour) serves both lanes: the even lane draws for the even
          // columns, the odd lane for the odd ones, and they swap
          const uint32_t kpair = (uint32_t)(ri & ~1), sel = (uint32_t)(ri & 1), t16 = ad.drop.thresh16;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int mine = i + (lane & 1);
            const uint32_t rowid = (uint32_t)(r0 + min(c0 + mine, Rb - 1)) * (uint32_t)ad.heads + (uint32_t)head;
            const uint32_t h = ad.drop.bits(rowid, kpair);
            const uint32_t o = __shfl_xor_sync(0xffffffffu, h, 1);
            const uint32_t he = (lane & 1) ? o : h, ho = (lane & 1) ? h : o;      // draws of columns i and i + 1
            const uint32_t ve = sel ? (he >> 16) : (he & 0xFFFFu), vo = sel ? (ho >> 16) : (ho & 0xFFFFu);
            keep |= (ve >= t16 ? 1u : 0u) << i;
            keep |= (vo >= t16 ? 1u : 0u) << (i + 1);
          }
        }
      }
      uint32_t* ds = reinterpret_cast<uint32_t*>(t1);       // results overwrite the registers they were computed from
      uint32_t* pd = reinterpret_cast<uint32_t*>(t2);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float lc = KV ? __shfl_sync(0xffffffffu, lse_c, i) : lse_r;
        const float dc = KV ? __shfl_sync(0xffffffffu, D_c, i) : D_r;
        float p = ex2b(fmaf(t1[i], sc2, -lc));
        if (c0 + i >= Rb) p = 0.f;                        // columns past the bag (the tile was filled with the next bag's rows)
        const bool k = (keep >> i) & 1u;
        const float dsv = p * ((k ? t2[i] * ik : 0.f) - dc);
        ds[i] = round_tf32(dsv);
        if (KV) pd[i] = round_tf32(k ? p * ik : 0.f);
      }
      tmem_st32(t1_addr, *reinterpret_cast<const uint32_t(*)[32]>(ds));
      if (KV) tmem_st32(t2_addr, *reinterpret_cast<const uint32_t(*)[32]>(pd));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[g]);
    }
    // ---------------- read-out: the accumulators are complete once the last tile's MMAs have retired ----------------
    mbar_wait(afull, 0);
    tc_fence_after();
    {
      const int part = 2 * g + half;                      // the four threads of a row take a quarter of the columns each
      const bool ok = ri < Rb;                            // (tcgen05.ld is warp-collective: every lane loads, valid rows store)
      float acc[OQ];
      tmem_ldq<OQ>(lane_base + A1_COL + OQ * part, acc);
      float* dst = d_qkv + (size_t)(r0 + ric) * 3 * d + (KV ? ck : cq) + OQ * part;
      if (ok) {
#pragma unroll
        for (int c = 0; c < OQ; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(acc[c] * scale, acc[c + 1] * scale, acc[c + 2] * scale, acc[c + 3] * scale);
      }
      if (KV) {
        tmem_ldq<OQ>(lane_base + A2_COL + OQ * part, acc);
        float* dv = d_qkv + (size_t)(r0 + ric) * 3 * d + cv + OQ * part;
        if (ok) {
#pragma unroll
          for (int c = 0; c < OQ; c += 4) *reinterpret_cast<float4*>(dv + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BT_EW) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int HD, bool KV>
int launch_bwd(const BwdMaps& maps, const int32_t* ro, int bags, int Rtot, int d, int heads, int mx, float scale, const AttDrop& ad,
               const float* lse, const float* ctx, const float* d_ctx, float* Dq, float* d_qkv, cudaStream_t st) {
  static bool attr_set[64] = {};
  int dev = 0;
  ADVMIL_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    ADVMIL_CHECK_CUDA(cudaFuncSetAttribute(mha_bwd_tcgen05_kernel<HD, KV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdCfg<HD, KV>::SMEM));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  launch_k(mha_bwd_tcgen05_kernel<HD, KV>, dim3(cdiv(mx, BT_ROWS), bags, heads), dim3(BT_THREADS), BwdCfg<HD, KV>::SMEM, st, maps, ro, d, scale, ad,
           lse, ctx, d_ctx, Dq, d_qkv, Rtot);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

template <int HD>
int bwd_both(const float* qkv, const float* ctx, const float* d_ctx, const float* lse, const int32_t* ro, int bags, int Rtot, int d, int heads,
             int mx, float scale, const AttDrop& ad, float* d_qkv, float* Dq, cudaStream_t st) {
  BwdMaps m;
  ADVMIL_TRY(make_map(&m.qkvK128, qkv, Rtot, 3LL * d, BT_ROWS, CU_TENSOR_MAP_SWIZZLE_128B));
  ADVMIL_TRY(make_map(&m.qkvK64, qkv, Rtot, 3LL * d, BT_COLS, CU_TENSOR_MAP_SWIZZLE_128B));
  ADVMIL_TRY(make_map(&m.qkvM64, qkv, Rtot, 3LL * d, BT_COLS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  ADVMIL_TRY(make_map(&m.dcK128, d_ctx, Rtot, (long long)d, BT_ROWS, CU_TENSOR_MAP_SWIZZLE_128B));
  ADVMIL_TRY(make_map(&m.dcK64, d_ctx, Rtot, (long long)d, BT_COLS, CU_TENSOR_MAP_SWIZZLE_128B));
  ADVMIL_TRY(make_map(&m.dcM64, d_ctx, Rtot, (long long)d, BT_COLS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  // the dQ pass first: it also produces D = dO . O for the dK/dV pass
  ADVMIL_TRY((launch_bwd<HD, false>(m, ro, bags, Rtot, d, heads, mx, scale, ad, lse, ctx, d_ctx, Dq, d_qkv, st)));
  return launch_bwd<HD, true>(m, ro, bags, Rtot, d, heads, mx, scale, ad, lse, ctx, d_ctx, Dq, d_qkv, st);
}

}  // namespace

int mha_bwd_tcgen05(const float* qkv, const float* ctx, const float* d_ctx, const float* lse, const int32_t* ro, int bags, int Rtot, int d,
                    int heads, int mx, float scale, const AttDrop& ad, float* d_qkv, float* Dq, cudaStream_t st) {
  ADVMIL_REQUIRE(((uintptr_t)qkv | (uintptr_t)d_ctx | (uintptr_t)ctx | (uintptr_t)d_qkv) % 16 == 0 && d % 4 == 0,
                 "attention bwd: tensors must be 16-byte aligned");
  switch (d / heads) {
    case 16: return bwd_both<16>(qkv, ctx, d_ctx, lse, ro, bags, Rtot, d, heads, mx, scale, ad, d_qkv, Dq, st);
    case 32: return bwd_both<32>(qkv, ctx, d_ctx, lse, ro, bags, Rtot, d, heads, mx, scale, ad, d_qkv, Dq, st);
    case 48: return bwd_both<48>(qkv, ctx, d_ctx, lse, ro, bags, Rtot, d, heads, mx, scale, ad, d_qkv, Dq, st);
    case 64: return bwd_both<64>(qkv, ctx, d_ctx, lse, ro, bags, Rtot, d, heads, mx, scale, ad, d_qkv, Dq, st);
  }
  set_error("attention bwd: head width %d unsupported on tcgen05", d / heads);
  return ADVMIL_ERR_INVALID;
}

}  // namespace advmil
