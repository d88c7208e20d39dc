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
