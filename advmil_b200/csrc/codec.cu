// Lossless 12-bit transport format of bf16 feature matrices ("p12") and its decoder.
//
// The end-to-end path is bound by the PCIe link (537 MB of bf16 features per 16-bag step against ~2.3 ms of compute).
// A bf16 word is sign(1) | exponent(8) | mantissa(7).  Sign and mantissa are incompressible, but the exponents of feature
// matrices cluster on a handful of values (~3 bits of entropy for normalised CNN features and for the synthetic
// benchmark's Gaussians alike), so the packed loader stores
//   lo  [n]    uint8 : sign << 7 | mantissa
//   hi  [n/2]  uint8 : two 4-bit exponent codes (element 2i in the low nibble)
//   table[16]  uint8 : code c < 15 -> exponent byte; code 15 = escape
//   esc_idx[m] int32 , esc_exp[m] uint8 : element index and exponent byte of every escaped element
// = 12 bits per element, exact for every bit pattern (zeros, denormals, infinities, NaN payloads).  The feeder copies
// the planes and decodes on the copy stream into the bf16 buffer the kernels read.
#include "stages.cuh"

namespace advmil {

struct P12Table { uint8_t t[16]; };

// 8 elements per thread: 8 bytes of lo + 4 bytes of hi -> one 16-byte store
__global__ void __launch_bounds__(256) bf16p12_decode_kernel(const uint2* __restrict__ lo, const uint32_t* __restrict__ hi,
                                                             P12Table tab, size_t n8, uint4* __restrict__ out) {
  pdl_prologue();
  __shared__ uint32_t tb[16];
  if (threadIdx.x < 16) tb[threadIdx.x] = (uint32_t)tab.t[threadIdx.x] << 7;   // code 15 (escape) decodes to 0, patched later
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint2 l = lo[i];
    const uint32_t h = hi[i];
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t lw = k < 2 ? l.x : l.y;
      const uint32_t b0 = (lw >> (16 * (k & 1))) & 0xFFu, b1 = (lw >> (16 * (k & 1) + 8)) & 0xFFu;
      const uint32_t c0 = (h >> (8 * k)) & 0xFu, c1 = (h >> (8 * k + 4)) & 0xFu;
      const uint32_t e0 = ((b0 & 0x80u) << 8) | tb[c0] | (b0 & 0x7Fu);
      const uint32_t e1 = ((b1 & 0x80u) << 8) | tb[c1] | (b1 & 0x7Fu);
      w[k] = e0 | (e1 << 16);
    }
    out[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void bf16p12_patch_kernel(const uint8_t* __restrict__ lo, const int32_t* __restrict__ esc_idx,
                                     const uint8_t* __restrict__ esc_exp, int n_esc, uint16_t* __restrict__ out) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_esc) return;
  const size_t j = (size_t)(uint32_t)esc_idx[i];
  const uint32_t b = lo[j];
  out[j] = (uint16_t)(((b & 0x80u) << 8) | ((uint32_t)esc_exp[i] << 7) | (b & 0x7Fu));
}


// =============================================================================================================
// "vl": variable-length (Huffman) transport form of the exponent plane.
//
// The 4-bit exponent codes of p12 carry ~2.6 bits of information each (Gaussian features: three exponents hold 76 % of the
// mass), and the end-to-end path is bound by the link (one GPU) or by the host memory bandwidth (eight GPUs), so the exponent
// plane is entropy-coded: the same 16-symbol alphabet (15 most frequent exponent bytes + escape), canonical Huffman codes of
// at most 8 bits, packed LSB-first.  To keep the decoder's loads and stores coalesced the elements of a SUPER-BLOCK of 4096 are
// dealt to the 32 lanes of a warp, lane t owning elements r * 128 + 4 t + {0..3}, r = 0..31, and every lane gets its own
// sub-stream (32-bit aligned):
//   lo     [n]          uint8  : sign << 7 | mantissa, element order                  (as in p12)
//   stream [words + 2]  uint32 : the sub-streams, super-block after super-block, lane after lane
//   sbase  [n / 4096]   uint32 : word offset of a super-block's first sub-stream
//   loff   [n / 128]    uint16 : word offset of lane t's sub-stream inside its super-block (<= 1024 words per super-block)
//   tables                     : symbol -> exponent byte, code length (0 = unused), LSB-first code
//   esc_idx / esc_exp          : sparse escapes, patched after the main pass          (as in p12)
// ~10.9 bits per element on the benchmark's features (8 + 2.65 code + 0.13 index + 0.12 alignment) instead of 12.
// Decoder: one warp per super-block; per round a lane loads 4 bytes of `lo` (128 B per warp), decodes 4 symbols from its 64-bit
// bit buffer through a 256-entry shared-memory table (8-bit peek -> exponent << 7 | length) and stores 8 bytes (256 B per warp).
// =============================================================================================================
constexpr int VL_SUPER = 4096;

struct VlTables { uint8_t exp[16]; uint8_t len[16]; uint16_t code[16]; };

__global__ void __launch_bounds__(256) bf16vl_decode_kernel(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ stream,
                                                            const uint32_t* __restrict__ sbase, const uint16_t* __restrict__ loff,
                                                            VlTables tab, int nsuper, uint2* __restrict__ out) {
  pdl_prologue();
  __shared__ uint32_t lut[256];        // low 8 bits of the bit buffer -> (exponent << 7) | (length << 16)
  for (int v = threadIdx.x; v < 256; v += blockDim.x) {
    uint32_t e = 8u << 16;             // prefixes no code maps to cannot occur in a valid stream
    for (int s = 0; s < 16; ++s) {
      const int len = tab.len[s];
      if (len && (v & ((1 << len) - 1)) == tab.code[s]) { e = ((uint32_t)(s < 15 ? tab.exp[s] : 0) << 7) | ((uint32_t)len << 16); break; }
    }
    lut[v] = e;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sb >= nsuper) return;
  const uint32_t* p = stream + (size_t)sbase[sb] + loff[(size_t)sb * 32 + lane];
  uint64_t buf = (uint64_t)p[0] | ((uint64_t)p[1] << 32);
  int have = 64, next = 2;
  const uint32_t* lp = lo + (size_t)sb * (VL_SUPER / 4) + lane;
  uint2* op = out + (size_t)sb * (VL_SUPER / 4) + lane;
#pragma unroll 4
  for (int r = 0; r < 32; ++r) {
    const uint32_t lw = lp[r * 32];
    uint32_t w[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      uint32_t e2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (have < 8) { buf |= (uint64_t)p[next++] << have; have += 32; }       // one refill keeps >= 8 valid bits
        const uint32_t t = lut[(uint32_t)buf & 0xFFu];
        const int len = (int)(t >> 16);
        const uint32_t bb = (lw >> (16 * k + 8 * h)) & 0xFFu;
        e2[h] = ((bb & 0x80u) << 8) | (t & 0x7F80u) | (bb & 0x7Fu);
        buf >>= len; have -= len;
      }
      w[k] = e2[0] | (e2[1] << 16);
    }
    op[r * 32] = make_uint2(w[0], w[1]);
  }
}

}  // namespace advmil

using namespace advmil;

extern "C" int advmil_bf16p12_decode(const uint8_t* lo, const uint8_t* hi, const uint8_t* table16_host, const int32_t* esc_idx,
                                     const uint8_t* esc_exp, int64_t n, int32_t n_esc, void* out_bf16, void* stream) {
  ADVMIL_REQUIRE(lo && hi && table16_host && out_bf16 && n >= 0 && n % 8 == 0 && n_esc >= 0 && (n_esc == 0 || (esc_idx && esc_exp)),
                 "bf16p12_decode: bad arguments (n must be a multiple of 8)");
  ADVMIL_REQUIRE((((uintptr_t)lo & 7) | ((uintptr_t)hi & 3) | ((uintptr_t)out_bf16 & 15)) == 0, "bf16p12_decode: planes must be 8/4/16-byte aligned");
  ADVMIL_REQUIRE(n < ((int64_t)1 << 32), "bf16p12_decode: at most 2^32 elements per call (escape indices are 32-bit)");
  if (n == 0) return ADVMIL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  P12Table tab;
  for (int i = 0; i < 16; ++i) tab.t[i] = i < 15 ? table16_host[i] : 0;
  const size_t n8 = (size_t)n / 8;
  const int grid = (int)min((size_t)148 * 16, (n8 + 255) / 256);
  launch_k(bf16p12_decode_kernel, dim3(grid), dim3(256), 0, st, (const uint2*)lo, (const uint32_t*)hi, tab, n8, (uint4*)out_bf16);
  ADVMIL_CHECK_LAUNCH();
  if (n_esc > 0) {
    launch_k(bf16p12_patch_kernel, dim3(cdiv(n_esc, 256)), dim3(256), 0, st, lo, esc_idx, esc_exp, (int)n_esc, (uint16_t*)out_bf16);
    ADVMIL_CHECK_LAUNCH();
  }
  return ADVMIL_OK;
}

// ---- host side of the "vl" form: encoder (packing time) and a reference decoder (tests) -------------------------------------
namespace {
// code lengths (<= 8) of a 16-symbol alphabet: Huffman on the counts, re-run with a rising floor on the rare counts until the
// deepest leaf fits (with 16 symbols a floor of total / 64 always does)
void vl_lengths(const uint64_t (&cnt)[16], uint8_t (&len)[16]) {
  uint64_t total = 0;
  for (int i = 0; i < 16; ++i) total += cnt[i];
  for (uint64_t fl = 0;; fl = fl ? fl * 2 : (total >> 12) + 1) {
    uint64_t w[31]; int parent[31]; bool alive[31]; int n = 0, used = 0;
    for (int i = 0; i < 16; ++i) { len[i] = 0; if (cnt[i]) { w[n] = cnt[i] > fl ? cnt[i] : fl; parent[n] = -1; alive[n] = true; ++n; ++used; } }
    int leaf_of[16], li = 0;
    for (int i = 0; i < 16; ++i) if (cnt[i]) leaf_of[li++] = i;
    if (used == 1) { len[leaf_of[0]] = 1; return; }
    const int leaves = n;
    while (true) {
      int a = -1, b = -1;
      for (int i = 0; i < n; ++i) if (alive[i]) { if (a < 0 || w[i] < w[a]) { b = a; a = i; } else if (b < 0 || w[i] < w[b]) b = i; }
      if (b < 0) break;
      w[n] = w[a] + w[b]; parent[n] = -1; alive[n] = true; alive[a] = alive[b] = false; parent[a] = parent[b] = n; ++n;
    }
    int mx = 0;
    for (int i = 0; i < leaves; ++i) { int d = 0; for (int q = i; parent[q] >= 0; q = parent[q]) ++d; len[leaf_of[i]] = (uint8_t)d; if (d > mx) mx = d; }
    if (mx <= 8) return;
  }
}
// canonical codes (shorter first, then by symbol), stored bit-reversed: the stream is read LSB-first
void vl_codes(const uint8_t (&len)[16], uint16_t (&code)[16]) {
  uint32_t next = 0;
  for (int l = 1; l <= 8; ++l) {
    for (int s = 0; s < 16; ++s) if (len[s] == l) {
      uint32_t c = next++, r = 0;
      for (int b = 0; b < l; ++b) r |= ((c >> b) & 1u) << (l - 1 - b);
      code[s] = (uint16_t)r;
    }
    next <<= 1;
  }
  for (int s = 0; s < 16; ++s) if (!len[s]) code[s] = 0;
}
}  // namespace

// tables of a file / step from its exponent histogram: the 15 most frequent exponent bytes + escape, code lengths, codes
static void vl_tables_from_hist(const uint64_t* hist, uint8_t* tab_exp16, uint8_t* tab_len16, uint16_t* tab_code16) {
  int order[256];
  for (int i = 0; i < 256; ++i) order[i] = i;
  for (int i = 0; i < 15; ++i) {            // the 15 most frequent exponent bytes (ties: smaller byte first)
    int best = i;
    for (int j = i + 1; j < 256; ++j) if (hist[order[j]] > hist[order[best]] || (hist[order[j]] == hist[order[best]] && order[j] < order[best])) best = j;
    const int t = order[i]; order[i] = order[best]; order[best] = t;
  }
  uint64_t cnt[16] = {};
  for (int s = 0; s < 15; ++s) { tab_exp16[s] = (uint8_t)order[s]; cnt[s] = hist[order[s]]; }
  tab_exp16[15] = 0;
  for (int i = 15; i < 256; ++i) cnt[15] += hist[order[i]];
  uint8_t len[16]; uint16_t code[16];
  vl_lengths(cnt, len);
  vl_codes(len, code);
  for (int s = 0; s < 16; ++s) { tab_len16[s] = len[s]; tab_code16[s] = code[s]; }
}

extern "C" int advmil_bf16vl_tables(const uint64_t* hist256, uint8_t* tab_exp16, uint8_t* tab_len16, uint16_t* tab_code16) {
  ADVMIL_REQUIRE(hist256 && tab_exp16 && tab_len16 && tab_code16, "bf16vl_tables: null argument");
  vl_tables_from_hist(hist256, tab_exp16, tab_len16, tab_code16);
  return ADVMIL_OK;
}

static int vl_encode_body(const uint16_t* x, int64_t n, uint8_t* lo, uint32_t* stream, int64_t stream_cap_words, uint32_t* sbase,
                          uint16_t* loff, const uint8_t* tab_exp16, const uint8_t* tab_len16, const uint16_t* tab_code16,
                          int32_t* esc_idx, uint8_t* esc_exp, int64_t esc_cap, int64_t* stream_words, int64_t* n_esc) {
  uint8_t sym_of[256], len[16];
  uint16_t code[16];
  for (int i = 0; i < 256; ++i) sym_of[i] = 15;
  for (int s = 0; s < 16; ++s) { len[s] = tab_len16[s]; code[s] = tab_code16[s]; }
  for (int s = 14; s >= 0; --s) if (len[s]) sym_of[tab_exp16[s]] = (uint8_t)s;      // unused table slots (length 0) map nothing
  int64_t w = 0, ne = 0;
  const int64_t nsuper = n / VL_SUPER;
  for (int64_t sb = 0; sb < nsuper; ++sb) {
    sbase[sb] = (uint32_t)w;
    const int64_t w0 = w;
    for (int lane = 0; lane < 32; ++lane) {
      ADVMIL_REQUIRE(w - w0 < 65536, "bf16vl_encode: sub-stream offset overflow");
      loff[sb * 32 + lane] = (uint16_t)(w - w0);
      uint64_t buf = 0; int have = 0;
      for (int r = 0; r < 32; ++r)
        for (int e = 0; e < 4; ++e) {
          const int64_t i = sb * VL_SUPER + r * 128 + lane * 4 + e;
          const uint16_t v = x[i];
          const uint8_t ex = (uint8_t)((v >> 7) & 0xFF), s = sym_of[ex];
          lo[i] = (uint8_t)(((v >> 8) & 0x80) | (v & 0x7F));
          if (s == 15) {
            ADVMIL_REQUIRE(len[15] != 0, "bf16vl_encode: an exponent outside the given tables, which have no escape code");
            ADVMIL_REQUIRE(esc_idx && esc_exp && ne < esc_cap, "bf16vl_encode: escape list too small");
            esc_idx[ne] = (int32_t)i; esc_exp[ne] = ex; ++ne;
          }
          buf |= (uint64_t)code[s] << have; have += len[s];
          if (have >= 32) {
            ADVMIL_REQUIRE(w < stream_cap_words, "bf16vl_encode: stream buffer too small");
            stream[w++] = (uint32_t)buf; buf >>= 32; have -= 32;
          }
        }
      if (have > 0) {
        ADVMIL_REQUIRE(w < stream_cap_words, "bf16vl_encode: stream buffer too small");
        stream[w++] = (uint32_t)buf;
      }
    }
  }
  ADVMIL_REQUIRE(w + 2 <= stream_cap_words, "bf16vl_encode: stream buffer too small for the two guard words");
  stream[w] = 0; stream[w + 1] = 0;           // the decoder's refill may read up to two words past a sub-stream
  *stream_words = w + 2;
  *n_esc = ne;
  return ADVMIL_OK;
}

static int vl_check_args(const void* x, int64_t n, const void* lo, const void* stream, const void* sbase, const void* loff, const void* te,
                         const void* tl, const void* tc, const void* sw, const void* ne) {
  ADVMIL_REQUIRE(x && lo && stream && sbase && loff && te && tl && tc && sw && ne && n >= 0 && n % VL_SUPER == 0 && n < ((int64_t)1 << 31),
                 "bf16vl_encode: bad arguments (n must be a multiple of 4096, < 2^31)");
  return ADVMIL_OK;
}

extern "C" int advmil_bf16vl_encode(const uint16_t* x, int64_t n, uint8_t* lo, uint32_t* stream, int64_t stream_cap_words, uint32_t* sbase,
                                    uint16_t* loff, uint8_t* tab_exp16, uint8_t* tab_len16, uint16_t* tab_code16, int32_t* esc_idx,
                                    uint8_t* esc_exp, int64_t esc_cap, int64_t* stream_words, int64_t* n_esc) {
  ADVMIL_TRY(vl_check_args(x, n, lo, stream, sbase, loff, tab_exp16, tab_len16, tab_code16, stream_words, n_esc));
  uint64_t hist[256] = {};
  for (int64_t i = 0; i < n; ++i) ++hist[(x[i] >> 7) & 0xFF];
  vl_tables_from_hist(hist, tab_exp16, tab_len16, tab_code16);
  return vl_encode_body(x, n, lo, stream, stream_cap_words, sbase, loff, tab_exp16, tab_len16, tab_code16, esc_idx, esc_exp, esc_cap,
                        stream_words, n_esc);
}

// same with GIVEN tables (advmil_bf16vl_tables over a whole file): every bag of a packed file shares one table, so the planes
// and streams of any selection of bags concatenate into a step without re-encoding
extern "C" int advmil_bf16vl_encode_with_tables(const uint16_t* x, int64_t n, uint8_t* lo, uint32_t* stream, int64_t stream_cap_words,
                                                uint32_t* sbase, uint16_t* loff, const uint8_t* tab_exp16, const uint8_t* tab_len16,
                                                const uint16_t* tab_code16, int32_t* esc_idx, uint8_t* esc_exp, int64_t esc_cap,
                                                int64_t* stream_words, int64_t* n_esc) {
  ADVMIL_TRY(vl_check_args(x, n, lo, stream, sbase, loff, tab_exp16, tab_len16, tab_code16, stream_words, n_esc));
  return vl_encode_body(x, n, lo, stream, stream_cap_words, sbase, loff, tab_exp16, tab_len16, tab_code16, esc_idx, esc_exp, esc_cap,
                        stream_words, n_esc);
}

extern "C" int advmil_bf16vl_decode_host(const uint8_t* lo, const uint32_t* stream, const uint32_t* sbase, const uint16_t* loff,
                                         const uint8_t* tab_exp16, const uint8_t* tab_len16, const uint16_t* tab_code16,
                                         const int32_t* esc_idx, const uint8_t* esc_exp, int64_t n, int64_t n_esc, uint16_t* out) {
  ADVMIL_REQUIRE(lo && stream && sbase && loff && tab_exp16 && tab_len16 && tab_code16 && out && n >= 0 && n % VL_SUPER == 0,
                 "bf16vl_decode_host: bad arguments");
  for (int64_t sb = 0; sb < n / VL_SUPER; ++sb)
    for (int lane = 0; lane < 32; ++lane) {
      const uint32_t* p = stream + sbase[sb] + loff[sb * 32 + lane];
      uint64_t buf = (uint64_t)p[0] | ((uint64_t)p[1] << 32);
      int have = 64, next = 2;
      for (int r = 0; r < 32; ++r)
        for (int e = 0; e < 4; ++e) {
          if (have < 8) { buf |= (uint64_t)p[next++] << have; have += 32; }
          int s = -1;
          for (int q = 0; q < 16; ++q) if (tab_len16[q] && ((uint32_t)buf & ((1u << tab_len16[q]) - 1)) == tab_code16[q]) { s = q; break; }
          ADVMIL_REQUIRE(s >= 0, "bf16vl_decode_host: invalid code in the stream");
          const int64_t i = sb * VL_SUPER + r * 128 + lane * 4 + e;
          const uint32_t b = lo[i], ex = s < 15 ? tab_exp16[s] : 0;
          out[i] = (uint16_t)(((b & 0x80u) << 8) | (ex << 7) | (b & 0x7Fu));
          buf >>= tab_len16[s]; have -= tab_len16[s];
        }
    }
  for (int64_t k = 0; k < n_esc; ++k) {
    const int64_t i = (uint32_t)esc_idx[k];
    const uint32_t b = lo[i];
    out[i] = (uint16_t)(((b & 0x80u) << 8) | ((uint32_t)esc_exp[k] << 7) | (b & 0x7Fu));
  }
  return ADVMIL_OK;
}

extern "C" int advmil_bf16vl_decode(const uint8_t* lo, const uint32_t* stream, const uint32_t* sbase, const uint16_t* loff,
                                    const uint8_t* tab_exp16_host, const uint8_t* tab_len16_host, const uint16_t* tab_code16_host,
                                    const int32_t* esc_idx, const uint8_t* esc_exp, int64_t n, int32_t n_esc, void* out_bf16, void* stream_) {
  ADVMIL_REQUIRE(lo && stream && sbase && loff && tab_exp16_host && tab_len16_host && tab_code16_host && out_bf16 && n >= 0 &&
                 n % VL_SUPER == 0 && n_esc >= 0 && (n_esc == 0 || (esc_idx && esc_exp)), "bf16vl_decode: bad arguments (n must be a multiple of 4096)");
  ADVMIL_REQUIRE((((uintptr_t)lo & 3) | ((uintptr_t)stream & 3) | ((uintptr_t)sbase & 3) | ((uintptr_t)loff & 1) | ((uintptr_t)out_bf16 & 7)) == 0,
                 "bf16vl_decode: misaligned planes");
  ADVMIL_REQUIRE(n < ((int64_t)1 << 31), "bf16vl_decode: at most 2^31 elements per call");
  if (n == 0) return ADVMIL_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  VlTables tab;
  for (int i = 0; i < 16; ++i) { tab.exp[i] = tab_exp16_host[i]; tab.len[i] = tab_len16_host[i]; tab.code[i] = tab_code16_host[i]; }
  const int nsuper = (int)(n / VL_SUPER);
  launch_k(bf16vl_decode_kernel, dim3(cdiv(nsuper, 8)), dim3(256), 0, st, (const uint32_t*)lo, stream, sbase, loff, tab, nsuper, (uint2*)out_bf16);
  ADVMIL_CHECK_LAUNCH();
  if (n_esc > 0) {
    launch_k(bf16p12_patch_kernel, dim3(cdiv(n_esc, 256)), dim3(256), 0, st, lo, esc_idx, esc_exp, (int)n_esc, (uint16_t*)out_bf16);
    ADVMIL_CHECK_LAUNCH();
  }
  return ADVMIL_OK;
}
