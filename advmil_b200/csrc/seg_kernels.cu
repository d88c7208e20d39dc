// Segmented (per-bag) kernels and row-streaming backward kernels.  HBM-bound: coalesced float4 streams,
// warp-shuffle reductions, deterministic two-level reductions (no atomics).
#include <stdarg.h>
#include "stages.cuh"

namespace advmil {

// =============================================================================================
// small utilities
// =============================================================================================
__global__ void fill_zero_kernel(float* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = 0.f;
}
int fill_zero(float* p, size_t n, cudaStream_t st) {
  if (n == 0) return ADVMIL_OK;
  int grid = (int)min((size_t)148 * 8, (n + 255) / 256);
  fill_zero_kernel<<<grid, 256, 0, st>>>(p, n);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

// out[c] (+)= sum_p part[p][c]; block (32,8): 8 row groups per column
__global__ void reduce_rows_kernel(const float* __restrict__ part, int nparts, int width /*row stride*/, int ncols,
                                   float* __restrict__ out, int accumulate) {
  __shared__ float sm[32][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < ncols) {
    int p = threadIdx.y;
    for (; p + 96 < nparts; p += 128) {
      float a0 = part[(size_t)p * width + c], a1 = part[(size_t)(p + 32) * width + c];
      float a2 = part[(size_t)(p + 64) * width + c], a3 = part[(size_t)(p + 96) * width + c];
      acc += (a0 + a1) + (a2 + a3);
    }
    for (; p < nparts; p += 32) acc += part[(size_t)p * width + c];
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < ncols) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += sm[r][threadIdx.x];
    out[c] = accumulate ? out[c] + t : t;
  }
}
static int reduce_rows(const float* part, int nparts, int width, float* out, int accumulate, cudaStream_t st) {
  reduce_rows_kernel<<<cdiv(width, 32), dim3(32, 32), 0, st>>>(part, nparts, width, width, out, accumulate);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, size_t n, float* __restrict__ out,
                                     int accumulate) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
      float4 v = *reinterpret_cast<const float4*>(ws + (size_t)z * n + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float4* o = reinterpret_cast<float4*>(out + i);
    if (accumulate) { float4 p = *o; acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w; }
    *o = acc;
  } else {
    for (; i < n; ++i) {
      float acc = 0.f;
      for (int z = 0; z < splits; ++z) acc += ws[(size_t)z * n + i];
      out[i] = accumulate ? out[i] + acc : acc;
    }
  }
}
int splitk_reduce(const float* ws, int splits, size_t n, float* out, int accumulate, cudaStream_t st) {
  splitk_reduce_kernel<<<cdiv((n + 3) / 4, 256), 256, 0, st>>>(ws, splits, n, out, accumulate);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

__global__ void apply_dropout_kernel(const float* __restrict__ src, size_t n, Drop drop, float* __restrict__ dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = src[i] * drop.scale(i);
}
int apply_dropout(const float* src, int rows, int width, const Drop& drop, float* dst, cudaStream_t st) {
  size_t n = (size_t)rows * width;
  int grid = (int)min((size_t)148 * 16, (n + 255) / 256);
  apply_dropout_kernel<<<grid, 256, 0, st>>>(src, n, drop, dst);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

// =============================================================================================
// gate weight packing: Wp[abw, L] rows = (64 tanh rows | 64 sigmoid rows) per 128-block, zero padded
// =============================================================================================
__global__ void gate_pack_kernel(const float* __restrict__ Wa, const float* __restrict__ ba,
                                 const float* __restrict__ Wb, const float* __restrict__ bb, int L, int D, int abw,
                                 float* __restrict__ Wp, float* __restrict__ bp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)abw * L;
  if (i < n) {
    int pr = (int)(i / L), l = (int)(i % L);
    int tile = pr >> 7, within = pr & 127;
    int j = tile * 64 + (within & 63);
    float v = 0.f;
    if (j < D) v = (within < 64) ? Wa[(size_t)j * L + l] : Wb[(size_t)j * L + l];
    Wp[i] = v;
  }
  if (i < (size_t)abw) {
    int pr = (int)i;
    int tile = pr >> 7, within = pr & 127;
    int j = tile * 64 + (within & 63);
    bp[pr] = (j < D) ? ((within < 64) ? ba[j] : bb[j]) : 0.f;
  }
}
int gate_pack_weights(const float* Wa, const float* ba, const float* Wb, const float* bb, int L, int D, float* Wp,
                      float* bp, cudaStream_t st) {
  int abw = gate_width(D);
  size_t n = (size_t)abw * L;
  gate_pack_kernel<<<cdiv(n, 256), 256, 0, st>>>(Wa, ba, Wb, bb, L, D, abw, Wp, bp);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

__global__ void gate_unpack_kernel(const float* __restrict__ dWp, const float* __restrict__ dbp, int L, int D,
                                   float* __restrict__ dWa, float* __restrict__ dba, float* __restrict__ dWb,
                                   float* __restrict__ dbb, int accumulate) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)D * L;
  if (i < n) {
    int j = (int)(i / L), l = (int)(i % L);
    int pa = gate_col_a(j);
    float va = dWp[(size_t)pa * L + l], vb = dWp[(size_t)(pa + 64) * L + l];
    dWa[i] = accumulate ? dWa[i] + va : va;
    dWb[i] = accumulate ? dWb[i] + vb : vb;
  }
  if (i < (size_t)D) {
    int pa = gate_col_a((int)i);
    dba[i] = accumulate ? dba[i] + dbp[pa] : dbp[pa];
    dbb[i] = accumulate ? dbb[i] + dbp[pa + 64] : dbp[pa + 64];
  }
}
int gate_unpack_grads(const float* dWp, const float* dbp, int L, int D, float* dWa, float* dba, float* dWb, float* dbb,
                      int accumulate, cudaStream_t st) {
  size_t n = (size_t)D * L;
  gate_unpack_kernel<<<cdiv(n, 256), 256, 0, st>>>(dWp, dbp, L, D, dWa, dba, dWb, dbb, accumulate);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

__global__ void gate_score_finish_kernel(const float* __restrict__ part, int ntiles, int rows,
                                         const float* __restrict__ bc, float* __restrict__ s) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  float acc = 0.f;
  for (int t = 0; t < ntiles; ++t) acc += part[(size_t)t * rows + m];
  s[m] = acc + bc[0];
}
int gate_score_finish(const float* part, int ntiles, int rows, const float* bc, float* s, cudaStream_t st) {
  gate_score_finish_kernel<<<cdiv(rows, 256), 256, 0, st>>>(part, ntiles, rows, bc, s);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

// =============================================================================================
// K3: segmented softmax + attention pooling (model/backbone.py:82-84; backbone_utils.py:53-55)
// =============================================================================================
constexpr int POOL_CH = 128;  // rows per pooling chunk

__global__ void seg_stats_kernel(const float* __restrict__ s, const int32_t* __restrict__ offsets,
                                 float* __restrict__ stats /*[bags][2] = max, 1/sum*/) {
  __shared__ float red[33];
  int b = blockIdx.x;
  int beg = offsets[b], end = offsets[b + 1];
  float mx = -INFINITY;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) mx = fmaxf(mx, s[i]);
  mx = block_max(mx, red);
  float sum = 0.f;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) sum += expf(s[i] - mx);
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) { stats[2 * b] = mx; stats[2 * b + 1] = 1.0f / sum; }
}

// grid (maxchunks, bags), 256 threads = RG row groups x W4 float4 columns
__global__ void __launch_bounds__(512) seg_pool_partial_kernel(
    const float* __restrict__ s, const float* __restrict__ v, const int32_t* __restrict__ offsets,
    const float* __restrict__ stats, int width, int want_mean, float* __restrict__ w,
    float* __restrict__ part /*[offsets[b]/POOL_CH + b + chunk][width]*/, float* __restrict__ part_mean) {
  extern __shared__ float sm[];  // w_s[POOL_CH] + red[RG][width] (+ red_mean)
  int b = blockIdx.y, chunk = blockIdx.x;
  int beg = offsets[b] + chunk * POOL_CH, end = min(offsets[b + 1], beg + POOL_CH);
  if (beg >= offsets[b + 1]) return;
  int nrows = end - beg;
  float* w_s = sm;
  float mx = stats[2 * b], inv = stats[2 * b + 1];
  for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
    float wv = expf(s[beg + r] - mx) * inv;
    w_s[r] = wv;
    w[beg + r] = wv;
  }
  __syncthreads();
  int W4 = width >> 2;
  int RG = blockDim.x / W4; if (RG > POOL_CH) RG = POOL_CH;
  int c4 = threadIdx.x % W4, rg = threadIdx.x / W4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), accm = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rg < RG) {
    const float4* vp = reinterpret_cast<const float4*>(v + (size_t)beg * width) + c4;
    int r = rg;
    for (; r + 3 * RG < nrows; r += 4 * RG) {
      float4 x0 = vp[(size_t)r * W4], x1 = vp[(size_t)(r + RG) * W4], x2 = vp[(size_t)(r + 2 * RG) * W4], x3 = vp[(size_t)(r + 3 * RG) * W4];
      float w0 = w_s[r], w1 = w_s[r + RG], w2 = w_s[r + 2 * RG], w3 = w_s[r + 3 * RG];
      acc.x += w0 * x0.x + w1 * x1.x + w2 * x2.x + w3 * x3.x;
      acc.y += w0 * x0.y + w1 * x1.y + w2 * x2.y + w3 * x3.y;
      acc.z += w0 * x0.z + w1 * x1.z + w2 * x2.z + w3 * x3.z;
      acc.w += w0 * x0.w + w1 * x1.w + w2 * x2.w + w3 * x3.w;
      if (want_mean) {
        accm.x += (x0.x + x1.x) + (x2.x + x3.x); accm.y += (x0.y + x1.y) + (x2.y + x3.y);
        accm.z += (x0.z + x1.z) + (x2.z + x3.z); accm.w += (x0.w + x1.w) + (x2.w + x3.w);
      }
    }
    for (; r < nrows; r += RG) {
      float4 x0 = vp[(size_t)r * W4];
      float w0 = w_s[r];
      acc.x += w0 * x0.x; acc.y += w0 * x0.y; acc.z += w0 * x0.z; acc.w += w0 * x0.w;
      if (want_mean) { accm.x += x0.x; accm.y += x0.y; accm.z += x0.z; accm.w += x0.w; }
    }
  }
  float* red = sm + POOL_CH;
  float* redm = red + (size_t)RG * width;
  if (rg < RG) {
    *reinterpret_cast<float4*>(red + (size_t)rg * width + c4 * 4) = acc;
    if (want_mean) *reinterpret_cast<float4*>(redm + (size_t)rg * width + c4 * 4) = accm;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < width; c += blockDim.x) {
    float t = 0.f, tm = 0.f;
    for (int g = 0; g < RG; ++g) { t += red[(size_t)g * width + c]; if (want_mean) tm += redm[(size_t)g * width + c]; }
    size_t o = ((size_t)(offsets[b] / POOL_CH + b + chunk)) * width + c;
    part[o] = t;
    if (want_mean) part_mean[o] = tm;
  }
}

__global__ void seg_pool_final_kernel(const float* __restrict__ part, const float* __restrict__ part_mean,
                                      const int32_t* __restrict__ offsets, int width,
                                      float* __restrict__ z, float* __restrict__ mean) {
  // blockDim = (128 columns, 8 chunk groups); grid = (ceil(width/128), bags)
  __shared__ float sm[2][8][128];
  const int b = blockIdx.y;
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int len = offsets[b + 1] - offsets[b];
  const int nch = (len + POOL_CH - 1) / POOL_CH;
  float t = 0.f, tm = 0.f;
  if (c < width) {
    const size_t base = ((size_t)(offsets[b] / POOL_CH + b)) * width + c;
    for (int k = threadIdx.y; k < nch; k += 8) {
      t += part[base + (size_t)k * width];
      if (mean) tm += part_mean[base + (size_t)k * width];
    }
  }
  sm[0][threadIdx.y][threadIdx.x] = t;
  sm[1][threadIdx.y][threadIdx.x] = tm;
  __syncthreads();
  if (threadIdx.y == 0 && c < width) {
    float a = 0.f, am = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) { a += sm[0][g][threadIdx.x]; am += sm[1][g][threadIdx.x]; }
    z[(size_t)b * width + c] = a;
    if (mean) mean[(size_t)b * width + c] = am / (float)len;
  }
}

static int pool_maxchunks(const int32_t* offsets_host, int bags) {
  int mx = 0;
  for (int b = 0; b < bags; ++b) mx = max(mx, offsets_host[b + 1] - offsets_host[b]);
  return cdiv(mx, POOL_CH);
}
size_t seg_pool_ws_floats(int rows, int bags, int width) {
  size_t nparts = (size_t)rows / POOL_CH + bags + 1;  // sum_b ceil(len_b / POOL_CH) <= rows/POOL_CH + bags
  return align_up(2 * (size_t)bags, 64) + 2 * nparts * width;
}
int seg_softmax_pool_fwd(const float* s, const float* v, const int32_t* offsets, const int32_t* offsets_host, int rows,
                         int bags, int width, float* w, float* z, float* mean, float* ws, cudaStream_t st) {
  ADVMIL_REQUIRE(width % 4 == 0 && width <= 1024, "seg_softmax_pool: width %d must be a multiple of 4 and <= 1024", width);
  for (int b = 0; b < bags; ++b)
    ADVMIL_REQUIRE(offsets_host[b + 1] > offsets_host[b], "seg_softmax_pool: bag %d is empty", b);
  int maxchunks = pool_maxchunks(offsets_host, bags);
  float* stats = ws;
  float* part = ws + align_up(2 * (size_t)bags, 64);
  float* part_mean = part + ((size_t)rows / POOL_CH + bags + 1) * width;
  seg_stats_kernel<<<bags, 1024, 0, st>>>(s, offsets, stats);
  ADVMIL_CHECK_LAUNCH();
  int W4 = width / 4;
  int threads = (W4 % 32 == 0 && 4 * W4 <= 512) ? 4 * W4 : 256;      // 384 threads for width 384: 4 full row groups
  int RG = min(threads / W4, POOL_CH);
  size_t smem = (POOL_CH + (size_t)RG * width * (mean ? 2 : 1)) * sizeof(float);
  seg_pool_partial_kernel<<<dim3(maxchunks, bags), threads, smem, st>>>(s, v, offsets, stats, width, mean ? 1 : 0, w,
                                                                     part, part_mean);
  ADVMIL_CHECK_LAUNCH();
  seg_pool_final_kernel<<<dim3(cdiv(width, 128), bags), dim3(128, 8), 0, st>>>(part, mean ? part_mean : nullptr, offsets, width, z, mean);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

// =============================================================================================
// backward of pooling + gated attention:
//   ds_n = w_n (dz.v_n - dz.z);  du_j = ds wc_j;  da_pre = du b_d sa (1-a^2);  db_pre = du a_d sb b(1-b)
// =============================================================================================
__global__ void bag_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int width, float* __restrict__ out) {
  __shared__ float red[33];
  int bag = blockIdx.x;
  float acc = 0.f;
  for (int c = threadIdx.x; c < width; c += blockDim.x) acc += a[(size_t)bag * width + c] * b[(size_t)bag * width + c];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[bag] = acc;
}

// one thread per gate-column pair (blockDim = abw/2 rounded up to a warp multiple, <= 512), rows of the chunk unrolled x4;
// also accumulates the column sums of dAB (the packed gate-bias gradient) so no separate pass over dAB is needed
__global__ void __launch_bounds__(512) pool_gate_bwd_kernel(
    const float* __restrict__ v, const float* __restrict__ w, const float* __restrict__ dz,
    const float* __restrict__ gz, const float* __restrict__ ab, const float* __restrict__ wc,
    const int32_t* __restrict__ offsets, int rows, int bags, int L, int D, int abw, Drop da, Drop db,
    float* __restrict__ dAB, float* __restrict__ part /*[chunks][D+1]*/, float* __restrict__ part_b /*[chunks][abw] or null*/) {
  __shared__ float ds_s[ROWS_PER_CTA];
  __shared__ float red[33];
  const int row0 = blockIdx.x * ROWS_PER_CTA;
  const int nrows = min(ROWS_PER_CTA, rows - row0);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // phase 1: one warp per row: g = dz[bag] . v[row]
  for (int r = wid; r < nrows; r += nw) {
    const int row = row0 + r;
    const int bag = bag_of_row(offsets, bags, row);
    const float* vr = v + (size_t)row * L;
    const float* dzr = dz + (size_t)bag * L;
    float acc = 0.f;
    for (int c = lane * 4; c < L; c += 128) {
      float4 x = *reinterpret_cast<const float4*>(vr + c);
      float4 g = *reinterpret_cast<const float4*>(dzr + c);
      acc += x.x * g.x + x.y * g.y + x.z * g.z + x.w * g.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) ds_s[r] = w[row] * (acc - gz[bag]);
  }
  __syncthreads();
  // phase 2
  const int npairs = abw >> 1;
  for (int q = threadIdx.x; q < npairs; q += blockDim.x) {
    const int ca = gate_col_a(q);
    const bool valid = q < D;
    const float wcj = valid ? wc[q] : 0.f;
    float dwc = 0.f, sa_sum = 0.f, sb_sum = 0.f;
    const bool tr = da.active != 0;
    int r = 0;
    for (; r + 3 < nrows; r += 4) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t row = (size_t)(row0 + r + u);
        a[u] = valid ? ab[row * abw + ca] : 0.f;
        b[u] = valid ? ab[row * abw + ca + 64] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t row = (size_t)(row0 + r + u);
        float sa = 1.f, sb = 1.f;
        if (tr && valid) { sa = da.keep(row * D + q) ? da.inv_keep : 0.f; sb = db.keep(row * D + q) ? db.inv_keep : 0.f; }
        const float ad = a[u] * sa, bd = b[u] * sb, ds = ds_s[r + u], du = ds * wcj;
        const float oa = du * bd * sa * (1.f - a[u] * a[u]);
        const float ob = du * ad * sb * b[u] * (1.f - b[u]);
        dwc = fmaf(ds, ad * bd, dwc);
        sa_sum += oa; sb_sum += ob;
        dAB[row * abw + ca] = oa;
        dAB[row * abw + ca + 64] = ob;
      }
    }
    for (; r < nrows; ++r) {
      const size_t row = (size_t)(row0 + r);
      const float a = valid ? ab[row * abw + ca] : 0.f, b = valid ? ab[row * abw + ca + 64] : 0.f;
      float sa = 1.f, sb = 1.f;
      if (tr && valid) { sa = da.keep(row * D + q) ? da.inv_keep : 0.f; sb = db.keep(row * D + q) ? db.inv_keep : 0.f; }
      const float ad = a * sa, bd = b * sb, ds = ds_s[r], du = ds * wcj;
      const float oa = du * bd * sa * (1.f - a * a), ob = du * ad * sb * b * (1.f - b);
      dwc = fmaf(ds, ad * bd, dwc);
      sa_sum += oa; sb_sum += ob;
      dAB[row * abw + ca] = oa;
      dAB[row * abw + ca + 64] = ob;
    }
    if (valid) part[(size_t)blockIdx.x * (D + 1) + q] = dwc;
    if (part_b) { part_b[(size_t)blockIdx.x * abw + ca] = sa_sum; part_b[(size_t)blockIdx.x * abw + ca + 64] = sb_sum; }
  }
  float t = 0.f;
  for (int r = threadIdx.x; r < nrows; r += blockDim.x) t += ds_s[r];
  t = block_sum(t, red);
  if (threadIdx.x == 0) part[(size_t)blockIdx.x * (D + 1) + D] = t;
}

int pool_gate_bwd(const float* v, const float* w, const float* z, const float* dz, const float* ab, const float* wc,
                  const int32_t* offsets, int rows, int bags, int L, int D, const Drop& da, const Drop& db, float* dAB,
                  float* dwc, float* dbc, float* dbp, int accumulate, float* ws, cudaStream_t st) {
  ADVMIL_REQUIRE(L % 4 == 0, "pool_gate_bwd: L %d must be a multiple of 4", L);
  const int chunks = row_chunks(rows);
  const int abw = gate_width(D);
  float* gz = ws;
  float* part = ws + align_up((size_t)bags, 64);
  float* part_b = dbp ? part + align_up((size_t)chunks * (D + 1), 64) : nullptr;
  bag_dot_kernel<<<bags, 128, 0, st>>>(dz, z, L, gz);
  ADVMIL_CHECK_LAUNCH();
  const int threads = min(512, ((abw / 2 + 31) / 32) * 32);
  pool_gate_bwd_kernel<<<chunks, threads, 0, st>>>(v, w, dz, gz, ab, wc, offsets, rows, bags, L, D, abw, da, db, dAB, part, part_b);
  ADVMIL_CHECK_LAUNCH();
  reduce_rows_kernel<<<cdiv(D, 32), dim3(32, 32), 0, st>>>(part, chunks, D + 1, D, dwc, accumulate);
  ADVMIL_CHECK_LAUNCH();
  reduce_rows_kernel<<<1, dim3(32, 32), 0, st>>>(part + D, chunks, D + 1, 1, dbc, accumulate);
  ADVMIL_CHECK_LAUNCH();
  if (dbp) {   // packed gate-bias gradient (never accumulated: the caller unpacks it with its own accumulate flag)
    reduce_rows_kernel<<<cdiv(abw, 32), dim3(32, 32), 0, st>>>(part_b, chunks, abw, abw, dbp, 0);
    ADVMIL_CHECK_LAUNCH();
  }
  return ADVMIL_OK;
}

// =============================================================================================
// K5/K6 backward per row: region mean (1/16), ReLU, LayerNorm (biased variance, eps) -> d_y
// =============================================================================================
template <int VPL>  // values per lane: d <= 32*VPL
__global__ void __launch_bounds__(256) ln_pool_bwd_kernel(
    const float* __restrict__ y_pre, const float* __restrict__ d_emb, const float* __restrict__ gamma,
    const float* __restrict__ beta, int rows, int d, float eps, float* __restrict__ d_y,
    float* __restrict__ part /*[chunks][3][d]*/) {
  extern __shared__ float sm[];  // [8 warps][3][d]
  int row0 = blockIdx.x * ROWS_PER_CTA;
  int nrows = min(ROWS_PER_CTA, rows - row0);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float g[VPL], be[VPL], pg[VPL], pb[VPL], pbias[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    int c = lane + 32 * k;
    g[k] = c < d ? gamma[c] : 0.f;
    be[k] = c < d ? beta[c] : 0.f;
    pg[k] = pb[k] = pbias[k] = 0.f;
  }
  const float inv_d = 1.0f / (float)d;
  for (int r = wid; r < nrows; r += 8) {
    size_t row = (size_t)(row0 + r);
    float y[VPL];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) { int c = lane + 32 * k; y[k] = c < d ? y_pre[row * d + c] : 0.f; s += y[k]; }
    float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) { int c = lane + 32 * k; float cc = c < d ? y[k] - mean : 0.f; q = fmaf(cc, cc, q); }
    float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    float xh[VPL], dxh[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      int c = lane + 32 * k;
      xh[k] = 0.f; dxh[k] = 0.f;
      if (c < d) {
        xh[k] = (y[k] - mean) * rstd;
        float e = fmaf(xh[k], g[k], be[k]);
        float de = e > 0.f ? d_emb[(row >> 4) * d + c] * (1.0f / 16.0f) : 0.f;
        pg[k] = fmaf(de, xh[k], pg[k]);
        pb[k] += de;
        dxh[k] = de * g[k];
        s1 += dxh[k];
        s2 = fmaf(dxh[k], xh[k], s2);
      }
    }
    float m1 = warp_sum(s1) * inv_d, m2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      int c = lane + 32 * k;
      if (c < d) {
        float dy = rstd * (dxh[k] - m1 - xh[k] * m2);
        d_y[row * d + c] = dy;
        pbias[k] += dy;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    int c = lane + 32 * k;
    if (c < d) {
      sm[(wid * 3 + 0) * d + c] = pg[k];
      sm[(wid * 3 + 1) * d + c] = pb[k];
      sm[(wid * 3 + 2) * d + c] = pbias[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * d; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += sm[wv * 3 * d + i];
    part[(size_t)blockIdx.x * 3 * d + i] = t;
  }
}

// d == 128: one float4 per lane per row, two rows in flight per warp
__global__ void __launch_bounds__(256) ln_pool_bwd128_kernel(
    const float* __restrict__ y_pre, const float* __restrict__ d_emb, const float* __restrict__ gamma,
    const float* __restrict__ beta, int rows, float eps, float* __restrict__ d_y, float* __restrict__ part) {
  __shared__ float sm[8 * 3 * 128];
  const int row0 = blockIdx.x * ROWS_PER_CTA;
  const int nrows = min(ROWS_PER_CTA, rows - row0);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float4 g4 = *reinterpret_cast<const float4*>(gamma + lane * 4), b4 = *reinterpret_cast<const float4*>(beta + lane * 4);
  const float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
  float pg[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {0.f, 0.f, 0.f, 0.f}, pbias[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = wid * 2; r < nrows; r += 16) {
    float4 y4[2], e4[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      ok[u] = r + u < nrows;
      const size_t row = (size_t)(row0 + r + u);
      y4[u] = ok[u] ? *reinterpret_cast<const float4*>(y_pre + row * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      e4[u] = ok[u] ? *reinterpret_cast<const float4*>(d_emb + (row >> 4) * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const size_t row = (size_t)(row0 + r + u);
      const float y[4] = {y4[u].x, y4[u].y, y4[u].z, y4[u].w}, ge[4] = {e4[u].x, e4[u].y, e4[u].z, e4[u].w};
      const float mean = warp_sum(y[0] + y[1] + y[2] + y[3]) * (1.0f / 128.0f);
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float c = y[k] - mean; q = fmaf(c, c, q); }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / 128.0f) + eps);
      float xh[4], dxh[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xh[k] = (y[k] - mean) * rstd;
        const float e = fmaf(xh[k], g[k], be[k]);
        const float de = (e > 0.f && ok[u]) ? ge[k] * (1.0f / 16.0f) : 0.f;
        pg[k] = fmaf(de, xh[k], pg[k]);
        pb[k] += de;
        dxh[k] = de * g[k];
        s1 += dxh[k];
        s2 = fmaf(dxh[k], xh[k], s2);
      }
      const float m1 = warp_sum(s1) * (1.0f / 128.0f), m2 = warp_sum(s2) * (1.0f / 128.0f);
      float dy[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { dy[k] = rstd * (dxh[k] - m1 - xh[k] * m2); if (ok[u]) pbias[k] += dy[k]; }
      if (ok[u]) *reinterpret_cast<float4*>(d_y + row * 128 + lane * 4) = make_float4(dy[0], dy[1], dy[2], dy[3]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sm[(wid * 3 + 0) * 128 + lane * 4 + k] = pg[k];
    sm[(wid * 3 + 1) * 128 + lane * 4 + k] = pb[k];
    sm[(wid * 3 + 2) * 128 + lane * 4 + k] = pbias[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * 128; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += sm[wv * 3 * 128 + i];
    part[(size_t)blockIdx.x * 3 * 128 + i] = t;
  }
}

int ln_pool_bwd(const float* y_pre, const float* d_emb, const float* gamma, const float* beta, int rows, int d,
                float eps, float* d_y, float* dgamma, float* dbeta, float* dbias, int accumulate, float* ws,
                cudaStream_t st) {
  ADVMIL_REQUIRE(d <= 256, "ln_pool_bwd: d %d > 256 unsupported", d);
  int chunks = row_chunks(rows);
  size_t smem = (size_t)8 * 3 * d * sizeof(float);
  if (d == 128) ln_pool_bwd128_kernel<<<chunks, 256, 0, st>>>(y_pre, d_emb, gamma, beta, rows, eps, d_y, ws);
  else if (d <= 128) ln_pool_bwd_kernel<4><<<chunks, 256, smem, st>>>(y_pre, d_emb, gamma, beta, rows, d, eps, d_y, ws);
  else ln_pool_bwd_kernel<8><<<chunks, 256, smem, st>>>(y_pre, d_emb, gamma, beta, rows, d, eps, d_y, ws);
  ADVMIL_CHECK_LAUNCH();
  reduce_rows_kernel<<<cdiv(d, 32), dim3(32, 32), 0, st>>>(ws, chunks, 3 * d, d, dgamma, accumulate);
  ADVMIL_CHECK_LAUNCH();
  reduce_rows_kernel<<<cdiv(d, 32), dim3(32, 32), 0, st>>>(ws + d, chunks, 3 * d, d, dbeta, accumulate);
  ADVMIL_CHECK_LAUNCH();
  reduce_rows_kernel<<<cdiv(d, 32), dim3(32, 32), 0, st>>>(ws + 2 * d, chunks, 3 * d, d, dbias, accumulate);
  ADVMIL_CHECK_LAUNCH();
  return ADVMIL_OK;
}

// =============================================================================================
// column sums (bias gradients)
// =============================================================================================
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ dY, int rows, int N, int ld,
                                                             float* __restrict__ part) {
  int row0 = blockIdx.x * ROWS_PER_CTA;
  int nrows = min(ROWS_PER_CTA, rows - row0);
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    const float* p = dY + (size_t)row0 * ld + c;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int r = 0;
    for (; r + 3 < nrows; r += 4) {
      a0 += p[(size_t)r * ld]; a1 += p[(size_t)(r + 1) * ld]; a2 += p[(size_t)(r + 2) * ld]; a3 += p[(size_t)(r + 3) * ld];
    }
    for (; r < nrows; ++r) a0 += p[(size_t)r * ld];
    part[(size_t)blockIdx.x * N + c] = (a0 + a1) + (a2 + a3);
  }
}
int colsum(const float* dY, int rows, int N, int ld, float* out, int accumulate, float* ws, cudaStream_t st) {
  int chunks = row_chunks(rows);
  colsum_partial_kernel<<<chunks, 256, 0, st>>>(dY, rows, N, ld, ws);
  ADVMIL_CHECK_LAUNCH();
  return reduce_rows(ws, chunks, N, out, accumulate, st);
}

}  // namespace advmil
