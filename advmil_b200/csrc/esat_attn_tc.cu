// ESAT self-attention on the Blackwell tensor path (tcgen05 / TMEM / TMA), forward.
// Reference: nn.MultiheadAttention inside nn.TransformerEncoderLayer (reference model/backbone.py:171-196,
// model/backbone_utils.py:112-127): softmax(q k^T / sqrt(hd)) -> dropout -> . v per (bag, head) over the regions of a bag.
//
// One CTA owns 128 queries of one (bag, head) and walks the bag's keys in tiles of 128 (flash style, nothing of size
// R x R touches HBM).  q, k and v are read straight out of the packed projection qkv [R, 3d] (fp32) by TMA -- head h is
// columns [h hd, (h+1) hd) of each third -- as 128-byte-swizzled blocks of 32 columns (hd = 48: two blocks, the second half
// used), and consumed by tcgen05.mma.kind::tf32 as they are:
//   S  = Q K^T      A = Q (K-major), B = K tile (K-major)                      -> TMEM, 128 columns, double-buffered
//   P  = 2^(S c - m)  by 8 softmax warps (TMEM lane quarter = warp % 4, column half = warp / 4), one query row per thread
//        pair; probabilities (after dropout) are written to shared memory in the K-major swizzled operand layout
//   Ot = P V        A = P (K-major, shared memory), B = V tile (MN-major: v rows are the contraction index)
//                                                                            -> TMEM, 64 columns, double-buffered
//   o  = o * 2^(m_old - m_new) + Ot   in registers (the tile result is read back, the running output never leaves registers)
// Warp roles (352 threads): warps 0-7 softmax / correction / epilogue, warp 8 MMA issuer, warp 9 TMA producer of q and the k
// tiles, warp 10 TMA producer of the v tiles (a k slot is free as soon as its S = Q K^T has retired, a v slot only after P V:
// separate rings keep the next S from waiting behind the previous P V plus a TMA round trip).
// Dropout keep bits: the same counter generator and indices as the warp-level kernels (esat_kernels.cu), so the backward
// pass regenerates them.
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
// row-major fp32 matrix [rows, cols]; box = 128 rows x 32 columns (128 bytes)
int make_map(CUtensorMap* m, const float* base, long long rows, long long cols, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ADVMIL_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32u, 128u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("attention: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld", (int)r, rows, cols); return ADVMIL_ERR_CUDA; }
  return ADVMIL_OK;
}

constexpr int AT_TILE = 128;                 // queries per CTA = keys per tile
constexpr int AT_SOFT = 8;                   // softmax warps
constexpr int AT_THREADS = 32 * (AT_SOFT + 3);
constexpr int AT_BLK = 128 * 128;            // bytes of one [128 rows x 128 B] operand block
constexpr float kLog2eA = 1.4426950408889634f, kLn2A = 0.6931471805599453f;

template <int HD> struct AttCfg {
  static_assert(HD % 16 == 0 && HD >= 16 && HD <= 64, "head width");
  static constexpr int NB = (HD + 31) / 32;                  // 32-column blocks per operand
  static constexpr int Q_BYTES = NB * AT_BLK, KV_BYTES = NB * AT_BLK, P_BYTES = 4 * AT_BLK;
  static constexpr int OH = HD / 2;                          // output columns per softmax thread (column half)
  static constexpr size_t SMEM = 1024 + (size_t)Q_BYTES + 4 * (size_t)KV_BYTES + P_BYTES + 2 * 2 * AT_TILE * 4 /*xm*/ + 2 * AT_TILE * 4 /*xl*/ + 256;
};

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t rna_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void soft_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * AT_SOFT) : "memory"); }

// TMEM -> registers, this warp's 32 lanes x N consecutive 32-bit columns (N = 8 or 16)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N> __device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[N]) {
  static_assert(N == 8 || N == 16 || N == 24 || N == 32, "column count");
  if constexpr (N == 8) tmem_ld8(taddr, v);
  else if constexpr (N == 16) tmem_ld16(taddr, v);
  else if constexpr (N == 24) { tmem_ld16(taddr, v); tmem_ld8(taddr + 16, v + 16); }
  else { tmem_ld16(taddr, v); tmem_ld16(taddr + 16, v + 16); }
}

template <int HD>
__global__ void __launch_bounds__(AT_THREADS, 1)
mha_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmK /*q, k: K-major*/, const __grid_constant__ CUtensorMap tmV /*v: MN-major*/,
                       const int32_t* __restrict__ ro, int d, float scale, AttDrop ad, float* __restrict__ ctx,
                       float* __restrict__ lse, int Rtot) {
  pdl_prologue();
  using Cfg = AttCfg<HD>;
  constexpr int NB = Cfg::NB, OH = Cfg::OH;
  const int b = blockIdx.y, head = blockIdx.z, r0 = ro[b], Rb = ro[b + 1] - r0;
  const int q0 = blockIdx.x * AT_TILE;
  if (q0 >= Rb) return;
  const int T = (Rb + AT_TILE - 1) / AT_TILE;        // key tiles

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* Q_s = smem;
  uint8_t* K_s = Q_s + Cfg::Q_BYTES;                 // [2][NB blocks]
  uint8_t* V_s = K_s + 2 * Cfg::KV_BYTES;            // [2][NB blocks]
  uint8_t* P_s = V_s + 2 * Cfg::KV_BYTES;            // [4 blocks of 32 keys]
  float* xm = (float*)(P_s + Cfg::P_BYTES);          // [2 (tile parity)][2 (half)][128]
  float* xl = xm + 2 * 2 * AT_TILE;                  // [2 (half)][128]
  uint64_t* bars = (uint64_t*)(xl + 2 * AT_TILE);
  uint64_t* qfull = bars;            // [1]
  uint64_t* kfull = bars + 1;        // [2]
  uint64_t* kfree = bars + 3;        // [2]
  uint64_t* vfull = bars + 5;        // [2]
  uint64_t* vfree = bars + 7;        // [2]
  uint64_t* sfull = bars + 9;        // [2]
  uint64_t* sfree = bars + 11;       // [2]
  uint64_t* ofull = bars + 13;       // [2]
  uint64_t* pfull = bars + 15;       // [1]
  uint32_t* tmem_ptr = (uint32_t*)(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(qfull, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kfull[i], 1); mbar_init(&kfree[i], 1); mbar_init(&vfull[i], 1); mbar_init(&vfree[i], 1);
      mbar_init(&sfull[i], 1); mbar_init(&sfree[i], 32 * AT_SOFT); mbar_init(&ofull[i], 1);
    }
    mbar_init(pfull, 32 * AT_SOFT);
    fence_barrier_init();
  }
  if (warp == AT_SOFT) tmem_alloc(tmem_ptr, 512);     // S: 2 x 128 columns, Ot: 2 x 64 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr uint32_t S_COL = 0, O_COL = 256;

  if (warp == AT_SOFT + 1) {
    // ---------------- TMA producer: q once, then the k tiles ----------------
    if (lane == 0) {
      const int cq = head * HD, ck = d + head * HD;
      mbar_arrive_expect_tx(qfull, Cfg::Q_BYTES);
#pragma unroll
      for (int g = 0; g < NB; ++g) tma_load_2d(Q_s + g * AT_BLK, &tmK, qfull, cq + 32 * g, r0 + q0);
      for (int j = 0; j < T; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait(&kfree[s], (uint32_t)(((j >> 1) & 1) ^ 1));
        mbar_arrive_expect_tx(&kfull[s], Cfg::KV_BYTES);
#pragma unroll
        for (int g = 0; g < NB; ++g) tma_load_2d(K_s + s * Cfg::KV_BYTES + g * AT_BLK, &tmK, &kfull[s], ck + 32 * g, r0 + j * AT_TILE);
      }
    }
  } else if (warp == AT_SOFT + 2) {
    // ---------------- TMA producer: the v tiles ----------------
    if (lane == 0) {
      const int cv = 2 * d + head * HD;
      for (int j = 0; j < T; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait(&vfree[s], (uint32_t)(((j >> 1) & 1) ^ 1));
        mbar_arrive_expect_tx(&vfull[s], Cfg::KV_BYTES);
#pragma unroll
        for (int g = 0; g < NB; ++g) tma_load_2d(V_s + s * Cfg::KV_BYTES + g * AT_BLK, &tmV, &vfull[s], cv + 32 * g, r0 + j * AT_TILE);
      }
    }
  } else if (warp == AT_SOFT) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t IDESC_S = idesc_tf32(AT_TILE, AT_TILE, 0, 0);
      constexpr uint32_t IDESC_O = idesc_tf32(AT_TILE, HD, 0, 1);
      const uint32_t q_addr = smem_u32(Q_s), p_addr = smem_u32(P_s);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&kfull[s], (uint32_t)((j >> 1) & 1));
        if (j >= 2) mbar_wait(&sfree[s], (uint32_t)(((j >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t k_addr = smem_u32(K_s + s * Cfg::KV_BYTES);
        uint32_t acc = 0;
#pragma unroll
        for (int kk = 0; kk < HD / 8; ++kk) {
          const int g = kk >> 2, k8 = kk & 3;
          mma_tf32(tmem_base + S_COL + 128 * s, smem_desc_sw128(q_addr + g * AT_BLK + k8 * 32, 16, 1024),
                   smem_desc_sw128(k_addr + g * AT_BLK + k8 * 32, 16, 1024), IDESC_S, acc);
          acc = 1;
        }
        mma_commit(&sfull[s]);
        mma_commit(&kfree[s]);
      };
      mbar_wait(qfull, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        const int s = j & 1;
        mbar_wait(&vfull[s], (uint32_t)((j >> 1) & 1));
        mbar_wait(pfull, (uint32_t)(j & 1));
        tc_fence_after();
        const uint32_t v_addr = smem_u32(V_s + s * Cfg::KV_BYTES);
#pragma unroll
        for (int k8 = 0; k8 < AT_TILE / 8; ++k8) {
          // A: block k8 / 4 of P, 32-byte step inside the swizzle row; B: 8 key rows (two 512-byte swizzle groups) of every MN group
          mma_tf32(tmem_base + O_COL + 64 * s, smem_desc_sw128(p_addr + (k8 >> 2) * AT_BLK + (k8 & 3) * 32, 16, 1024),
                   smem_desc_sw128(v_addr + k8 * 1024, AT_BLK, 512, 1), IDESC_O, k8 != 0 ? 1u : 0u);
        }
        mma_commit(&ofull[s]);
        mma_commit(&vfree[s]);
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue: row = 32 * (warp % 4) + lane, column half = warp / 4 ----------------
    const int qt = warp & 3, half = warp >> 2, row = qt * 32 + lane;
    const int qi = q0 + row;
    const bool row_ok = qi < Rb;
    const uint32_t lane_base = tmem_base + ((uint32_t)(qt * 32) << 16);
    const float sc = scale * kLog2eA;                 // logits in log2 units
    const uint32_t drow = (uint32_t)(r0 + min(qi, Rb - 1)) * (uint32_t)ad.heads + (uint32_t)head;
    float o[OH];
#pragma unroll
    for (int c = 0; c < OH; ++c) o[c] = 0.f;
    float m = -INFINITY, l = 0.f, c_prev = 0.f;
    for (int j = 0; j < T; ++j) {
      const int s = j & 1;
      const int nk = min(AT_TILE, Rb - j * AT_TILE);
      mbar_wait(&sfull[s], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      float sv[64];
      {
        float t0[32], t1[32];
        tmem_ld32(lane_base + S_COL + 128 * s + 64 * half, t0);
        tmem_ld32(lane_base + S_COL + 128 * s + 64 * half + 32, t1);
#pragma unroll
        for (int i = 0; i < 32; ++i) { sv[i] = t0[i]; sv[32 + i] = t1[i]; }
      }
      tc_fence_before();
      mbar_arrive(&sfree[s]);
      const int kbase = 64 * half;                     // first key column of this thread inside the tile
      if (nk < AT_TILE) {
#pragma unroll
        for (int i = 0; i < 64; ++i) if (kbase + i >= nk) sv[i] = -INFINITY;
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i) mx = fmaxf(mx, sv[i]);
      float* xmj = xm + (j & 1) * 2 * AT_TILE;
      xmj[half * AT_TILE + row] = mx;
      soft_bar();
      mx = fmaxf(mx, xmj[(half ^ 1) * AT_TILE + row]);
      const float mn = fmaxf(m, mx * sc);              // sc > 0
      const float cj = ex2a(m - mn);
      l *= cj;
      float ls = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) { sv[i] = ex2a(fmaf(sv[i], sc, -mn)); ls += sv[i]; }
      l += ls;
      if (ad.drop.active) {
        const int kg = j * AT_TILE + kbase;             // key index inside the bag
        if (ad.mask) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (!ad.keep(b, head, min(qi, Rb - 1), min(kg + i, Rb - 1), Rb, 0)) sv[i] = 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            bool ka, kb;
            ad.drop.keep2(drow, (uint32_t)(kg + i), ka, kb);
            if (!ka) sv[i] = 0.f;
            if (!kb) sv[i + 1] = 0.f;
          }
        }
      }
      // the previous tile's P V has finished: shared-memory P may be overwritten, and its result joins the running output
      if (j >= 1) {
        const int sp = (j - 1) & 1;
        mbar_wait(&ofull[sp], (uint32_t)(((j - 1) >> 1) & 1));
        tc_fence_after();
        float ot[OH];
        tmem_ldn<OH>(lane_base + O_COL + 64 * sp + OH * half, ot);
#pragma unroll
        for (int c = 0; c < OH; ++c) o[c] = fmaf(o[c], c_prev, ot[c]);
      }
      c_prev = cj;
      m = mn;
      // P (tf32, round to nearest) -> shared memory, K-major 128-byte-swizzled: key block 2 * half + i / 32
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        uint8_t* blk = P_s + (2 * half + kb) * AT_BLK + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 w = make_uint4(rna_tf32(sv[kb * 32 + 4 * c]), rna_tf32(sv[kb * 32 + 4 * c + 1]), rna_tf32(sv[kb * 32 + 4 * c + 2]),
                                     rna_tf32(sv[kb * 32 + 4 * c + 3]));
          *reinterpret_cast<uint4*>(blk + ((c ^ (row & 7)) << 4)) = w;
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(pfull);
    }
    {   // last tile's P V
      const int sp = (T - 1) & 1;
      mbar_wait(&ofull[sp], (uint32_t)(((T - 1) >> 1) & 1));
      tc_fence_after();
      float ot[OH];
      tmem_ldn<OH>(lane_base + O_COL + 64 * sp + OH * half, ot);
#pragma unroll
      for (int c = 0; c < OH; ++c) o[c] = fmaf(o[c], c_prev, ot[c]);
    }
    xl[half * AT_TILE + row] = l;
    soft_bar();
    l += xl[(half ^ 1) * AT_TILE + row];
    if (row_ok) {
      const float inv = ad.drop.inv_keep / l;
      float* dst = ctx + (size_t)(r0 + qi) * d + head * HD + OH * half;
#pragma unroll
      for (int c = 0; c < OH; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
      if (half == 0) lse[(size_t)head * Rtot + r0 + qi] = m * kLn2A + __logf(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_SOFT) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

bool mha_tcgen05_supported(int hd, int d) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("ADVMIL_ATTN_TCGEN05"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
  return enabled && (hd == 16 || hd == 32 || hd == 48 || hd == 64) && (3 * (long long)d * 4) % 16 == 0;
}

template <int HD>
static int launch_fwd(const float* qkv, const int32_t* ro, int bags, int Rtot, int d, int heads, int mx, float scale, const AttDrop& ad,
                      float* ctx, float* lse, cudaStream_t st) {
  CUtensorMap tmK, tmV;
  ADVMIL_TRY(make_map(&tmK, qkv, Rtot, 3LL * d, CU_TENSOR_MAP_SWIZZLE_128B));
  ADVMIL_TRY(make_map(&tmV, qkv, Rtot, 3LL * d, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  static bool attr_set[64] = {};
  int dev = 0;
  ADVMIL_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    ADVMIL_CHECK_CUDA(cudaFuncSetAttribute(mha_fwd_tcgen05_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AttCfg<HD>::SMEM));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  launch_k(mha_fwd_tcgen05_kernel<HD>, dim3(cdiv(mx, AT_TILE), bags, heads), dim3(AT_THREADS), AttCfg<HD>::SMEM, st, tmK, tmV, ro, d, scale,
           ad, ctx, lse, Rtot);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

int mha_fwd_tcgen05(const float* qkv, const int32_t* ro, int bags, int Rtot, int d, int heads, int mx, float scale, const AttDrop& ad,
                    float* ctx, float* lse, cudaStream_t st) {
  ADVMIL_REQUIRE(((uintptr_t)qkv) % 16 == 0, "attention: qkv must be 16-byte aligned");
  switch (d / heads) {
    case 16: return launch_fwd<16>(qkv, ro, bags, Rtot, d, heads, mx, scale, ad, ctx, lse, st);
    case 32: return launch_fwd<32>(qkv, ro, bags, Rtot, d, heads, mx, scale, ad, ctx, lse, st);
    case 48: return launch_fwd<48>(qkv, ro, bags, Rtot, d, heads, mx, scale, ad, ctx, lse, st);
    case 64: return launch_fwd<64>(qkv, ro, bags, Rtot, d, heads, mx, scale, ad, ctx, lse, st);
  }
  set_error("attention: head width %d unsupported on tcgen05", d / heads);
  return ADVMIL_ERR_INVALID;
}

}  // namespace advmil
