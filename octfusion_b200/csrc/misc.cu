// Small per-step kernels: weight re-layout, timestep embeddings, the DDIM update, row copies
// with index maps, and the Morton-ordered dense voxel neighbour tables of the LR middle U-Net.
#include "common.cuh"

namespace of {

__global__ void repack_weight_kernel(const float* __restrict__ src, int64_t s_tap, int64_t s_c, int64_t s_n,
                                     int taps, int c, int N, float* __restrict__ dst) {
  const int64_t total = (int64_t)taps * c * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int64_t k = i / N;
    const int ci = (int)(k % c);
    const int tap = (int)(k / c);
    dst[i] = src[tap * s_tap + ci * s_c + n * s_n];
  }
}

// reference ldm_diffusion_util.py:171-191: [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(P) i / half)
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int batch, int dim, float max_period,
                                          float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * dim) return;
  const int b = i / dim, j = i - b * dim;
  const int half = dim / 2;
  float v = 0.0f;
  if (j < 2 * half) {
    const int f = j < half ? j : j - half;
    const float freq = expf(-logf(max_period) * (float)f / (float)half);
    const float a = t[b] * freq;
    v = j < half ? cosf(a) : sinf(a);
  }
  out[i] = v;
}

// reference modules.py:558-563: [t | sin(2 pi t w) | cos(2 pi t w)]
__global__ void learned_sinusoidal_kernel(const float* __restrict__ t, const float* __restrict__ w, int batch,
                                          int half, float* __restrict__ out) {
  const int dim = 2 * half + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * dim) return;
  const int b = i / dim, j = i - b * dim;
  float v;
  if (j == 0) {
    v = t[b];
  } else {
    const int f = (j - 1) % half;
    const float a = t[b] * w[f] * 2.0f * 3.14159265358979323846f;
    v = (j - 1) < half ? sinf(a) : cosf(a);
  }
  out[i] = v;
}

__global__ void embedding_add_kernel(const float* __restrict__ table, const int32_t* __restrict__ label,
                                     int batch, int dim, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * dim) return;
  const int b = i / dim, j = i - b * dim;
  out[i] += table[(int64_t)label[b] * dim + j];
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// reference octfusion_model_union.py:345-350 ("eps" branch of sample_loop)
template <typename TA>
__global__ void ddim_eps_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                const float* __restrict__ log_snr, const float* __restrict__ log_snr_next,
                                int64_t n, TA* __restrict__ x_act) {
  const float ls = *log_snr, lsn = *log_snr_next;
  const float alpha = sqrtf(sigmoid_acc(ls)), sigma = sqrtf(sigmoid_acc(-ls));
  const float alpha_n = sqrtf(sigmoid_acc(lsn)), sigma_n = sqrtf(sigmoid_acc(-lsn));
  const float inv_alpha = 1.0f / fmaxf(alpha, 1e-8f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float e = eps[i];
    const float x0 = (x[i] - e * sigma) * inv_alpha;
    const float v = x0 * alpha_n + e * sigma_n;
    x[i] = v;
    if (x_act) Elem<TA>::st(x_act + i, v);
  }
}

// reference octfusion_model_union.py:324-344 ("x0" branch of sample_loop: stage-1 ancestral update with the
// truncation trick): pred <- sign(pred) when do_sign; x <- alpha' (x (1-c)/alpha + c pred) + sqrt(sigma'^2 c) noise
__global__ void ddpm_x0_kernel(float* __restrict__ x, float* __restrict__ pred, const float* __restrict__ noise,
                               const float* __restrict__ log_snr, const float* __restrict__ log_snr_next, int64_t n,
                               int do_sign) {
  const float ls = *log_snr, lsn = *log_snr_next;
  const float alpha = sqrtf(sigmoid_acc(ls));
  const float alpha_n = sqrtf(sigmoid_acc(lsn)), sigma_n = sqrtf(sigmoid_acc(-lsn));
  const float c = -expm1f(ls - lsn);
  const float sd = sqrtf(sigma_n * sigma_n * c);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float o = pred[i];
    if (do_sign) { o = o > 0.0f ? 1.0f : (o < 0.0f ? -1.0f : 0.0f); pred[i] = o; }
    const float mean = alpha_n * (x[i] * (1.0f - c) / alpha + c * o);
    x[i] = mean + (noise ? sd * noise[i] : 0.0f);
  }
}

template <typename TS, typename TD>
__global__ void copy_rows_kernel(const TS* __restrict__ src, int64_t lds, const int32_t* __restrict__ src_rows,
                                 TD* __restrict__ dst, int64_t ldd, const int32_t* __restrict__ dst_rows,
                                 int64_t rows, int c) {
  const int64_t total = rows * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c;
    const int j = (int)(i - r * c);
    const int64_t rs = src_rows ? src_rows[r] : r;
    const int64_t rd = dst_rows ? dst_rows[r] : r;
    Elem<TD>::st(dst + rd * ldd + j, Elem<TS>::ld(src + rs * lds + j));
  }
}

// 16-byte vector variant (same dtype, c a multiple of 16 bytes, 16-byte aligned rows)
__global__ void copy_rows_vec_kernel(const uint4* __restrict__ src, int64_t lds16, const int32_t* __restrict__ src_rows,
                                     uint4* __restrict__ dst, int64_t ldd16, const int32_t* __restrict__ dst_rows,
                                     int64_t rows, int c16) {
  const int64_t total = rows * c16;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c16;
    const int j = (int)(i - r * c16);
    const int64_t rs = src_rows ? src_rows[r] : r;
    const int64_t rd = dst_rows ? dst_rows[r] : r;
    dst[rd * ldd16 + j] = src[rs * lds16 + j];
  }
}

__global__ void histogram_kernel(const int32_t* __restrict__ v, int64_t n, int bins, int32_t* __restrict__ hist) {
  extern __shared__ int32_t sh[];
  for (int i = threadIdx.x; i < bins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = v[i];
    if (b >= 0 && b < bins) atomicAdd(&sh[b], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

__device__ __forceinline__ void morton_decode(uint32_t k, int bits, int& x, int& y, int& z) {
  x = y = z = 0;
  for (int i = 0; i < bits; ++i) {
    x |= ((k >> (3 * i + 2)) & 1u) << i;
    y |= ((k >> (3 * i + 1)) & 1u) << i;
    z |= ((k >> (3 * i)) & 1u) << i;
  }
}
__device__ __forceinline__ uint32_t morton_encode(int x, int y, int z, int bits) {
  uint32_t k = 0;
  for (int i = 0; i < bits; ++i)
    k |= (((uint32_t)x >> i) & 1u) << (3 * i + 2) | (((uint32_t)y >> i) & 1u) << (3 * i + 1) |
         (((uint32_t)z >> i) & 1u) << (3 * i);
  return k;
}

// mode 0: same-res 3^3; mode 1: stride 2 (out = in/2, in coord = 2*o + d); mode 2: nearest x2
// upsample then 3^3 (out = 2*in, in coord = (o + d) >> 1).  Conv3d is a cross-correlation with
// padding 1: out[o] = sum_d w[d+1] in[o*stride + d]  (reference modules.py:70,88,493).
__global__ void dense_tap_table_kernel(int mode, int out_bits, int batch, int32_t* __restrict__ tab) {
  const int64_t nout = (int64_t)batch << (3 * out_bits);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nout * 27) return;
  const int tap = (int)(i % 27);
  const int64_t row = i / 27;
  const int b = (int)(row >> (3 * out_bits));
  const uint32_t k = (uint32_t)(row & ((1ll << (3 * out_bits)) - 1));
  int x, y, z;
  morton_decode(k, out_bits, x, y, z);
  const int dx = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dz = tap % 3 - 1;
  int in_bits, ix, iy, iz;
  bool ok;
  if (mode == 0) {
    in_bits = out_bits; ix = x + dx; iy = y + dy; iz = z + dz;
    const int lim = 1 << in_bits;
    ok = ix >= 0 && iy >= 0 && iz >= 0 && ix < lim && iy < lim && iz < lim;
  } else if (mode == 1) {
    in_bits = out_bits + 1; ix = 2 * x + dx; iy = 2 * y + dy; iz = 2 * z + dz;
    const int lim = 1 << in_bits;
    ok = ix >= 0 && iy >= 0 && iz >= 0 && ix < lim && iy < lim && iz < lim;
  } else {
    in_bits = out_bits - 1;
    const int fx = x + dx, fy = y + dy, fz = z + dz;
    const int lim = 1 << out_bits;
    ok = fx >= 0 && fy >= 0 && fz >= 0 && fx < lim && fy < lim && fz < lim;
    ix = fx >> 1; iy = fy >> 1; iz = fz >> 1;
  }
  int32_t v = -1;
  if (ok) v = (int32_t)(((int64_t)b << (3 * in_bits)) + morton_encode(ix, iy, iz, in_bits));
  tab[i] = v;
}

// out[b, n] = sum_k act(x[b, k]) * W[n, k] + bias[n] for a handful of rows (the [B, 512] timestep-embedding
// MLPs: nn.Linear layers of time_embed / emb_layers / time_mlp, reference graph_unet_hr.py:107-111,
// modules.py:709-715,479-482).  One warp per output column; x is staged in shared memory; W is read once,
// coalesced along k, in its native nn.Linear [N, K] layout.
__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x, int64_t ldx,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           int B, int K, int N, int a_silu, float* __restrict__ out,
                                                           int64_t ldo, int cols_per_cta) {
  // One CTA stages the (activated) input rows once and produces cols_per_cta output columns, one per warp at a
  // time: the staging (32 x K floats) is the expensive part, the weight rows stream through once.
  extern __shared__ float xs[];                     // [min(B,32)][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int nb = min(32, B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * K; i += blockDim.x) {
      const int b = i / K, k = i - b * K;
      float v = x[(int64_t)(b0 + b) * ldx + k];
      xs[i] = a_silu ? silu_f(v) : v;
    }
    __syncthreads();
   for (int n = blockIdx.x * cols_per_cta + warp; n < min(N, (int)(blockIdx.x + 1) * cols_per_cta); n += 8) {
    float acc[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) acc[b] = 0.0f;
    const float* wr = w + (int64_t)n * K;
    for (int k = lane; k < K; k += 32) {
      const float wv = wr[k];
#pragma unroll
      for (int b = 0; b < 32; ++b)
        if (b < nb) acc[b] = fmaf(xs[b * K + k], wv, acc[b]);
    }
    float mine = 0.0f;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      float v = acc[b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == b) mine = v;
    }
    if (lane < nb) out[(int64_t)(b0 + lane) * ldo + n] = mine + (bias ? bias[n] : 0.0f);
   }
  }
}

static inline int grid_for(int64_t n, int block = 256, int cap_mult = 32) {
  int64_t want = (n + block - 1) / block;
  int64_t cap = (int64_t)num_sms() * cap_mult;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace of

using namespace of;

extern "C" int of_repack_weight(const float* src, int64_t s_tap, int64_t s_c, int64_t s_n, int32_t taps,
                                int32_t c, int32_t N, float* dst, void* stream) {
  OF_REQUIRE(src && dst && taps > 0 && c > 0 && N > 0, "of_repack_weight: bad arguments");
  const int64_t total = (int64_t)taps * c * N;
  repack_weight_kernel<<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, s_tap, s_c, s_n,
                                                                                           taps, c, N, dst);
  OF_LAUNCH_CHECK("of_repack_weight");
  return OF_OK;
}

extern "C" int of_linear_small(const float* x, int64_t ldx, const float* w_nk, const float* bias, int32_t B, int32_t K,
                               int32_t N, int32_t a_silu, float* out, int64_t ldo, void* stream) {
  OF_REQUIRE(x && w_nk && out && B > 0 && K > 0 && N > 0, "of_linear_small: bad arguments");
  const size_t smem = (size_t)(B < 32 ? B : 32) * K * sizeof(float);
  OF_REQUIRE(smem <= 200 * 1024, "of_linear_small: K=%d too large for the shared staging buffer", K);
  static size_t cfg = 48 * 1024;
  if (smem > cfg) {
    cudaFuncSetAttribute(linear_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cfg = 200 * 1024;
  }
  // about one CTA per SM, columns per CTA a multiple of the 8 warps
  int cpc = (N + num_sms() - 1) / num_sms();
  cpc = (cpc + 7) / 8 * 8;
  linear_small_kernel<<<(N + cpc - 1) / cpc, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, ldx, w_nk, bias, B, K, N,
                                                                                                  a_silu, out, ldo, cpc);
  OF_LAUNCH_CHECK("of_linear_small");
  return OF_OK;
}

extern "C" int of_timestep_embedding(const float* t, int32_t batch, int32_t dim, float max_period, float* out,
                                     void* stream) {
  OF_REQUIRE(t && out && batch > 0 && dim > 0, "of_timestep_embedding: bad arguments");
  const int n = batch * dim;
  timestep_embedding_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, batch, dim,
                                                                                               max_period, out);
  OF_LAUNCH_CHECK("of_timestep_embedding");
  return OF_OK;
}

extern "C" int of_learned_sinusoidal(const float* t, const float* w, int32_t batch, int32_t half, float* out,
                                     void* stream) {
  OF_REQUIRE(t && w && out && batch > 0 && half > 0, "of_learned_sinusoidal: bad arguments");
  const int n = batch * (2 * half + 1);
  learned_sinusoidal_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, w, batch, half,
                                                                                               out);
  OF_LAUNCH_CHECK("of_learned_sinusoidal");
  return OF_OK;
}

extern "C" int of_embedding_add(const float* table, const int32_t* label, int32_t batch, int32_t dim, float* out,
                                void* stream) {
  OF_REQUIRE(table && label && out && batch > 0 && dim > 0, "of_embedding_add: bad arguments");
  const int n = batch * dim;
  embedding_add_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(table, label, batch, dim,
                                                                                          out);
  OF_LAUNCH_CHECK("of_embedding_add");
  return OF_OK;
}

extern "C" int of_ddim_eps_update(float* x, const float* eps, const float* log_snr, const float* log_snr_next,
                                  int64_t n, void* x_act, int32_t act_dtype, void* stream) {
  OF_REQUIRE(x && eps && log_snr && log_snr_next && n >= 0, "of_ddim_eps_update: bad arguments");
  if (n == 0) return OF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (x_act != nullptr && act_dtype == OF_BF16)
    ddim_eps_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, st>>>(x, eps, log_snr, log_snr_next, n,
                                                                reinterpret_cast<__nv_bfloat16*>(x_act));
  else
    ddim_eps_kernel<float><<<grid_for(n), 256, 0, st>>>(x, eps, log_snr, log_snr_next, n,
                                                        reinterpret_cast<float*>(x_act));
  OF_LAUNCH_CHECK("of_ddim_eps_update");
  return OF_OK;
}

extern "C" int of_ddpm_x0_update(float* x, float* pred, const float* noise, const float* log_snr,
                                 const float* log_snr_next, int64_t n, int32_t do_sign, void* stream) {
  OF_REQUIRE(x && pred && log_snr && log_snr_next && n >= 0, "of_ddpm_x0_update: bad arguments");
  if (n == 0) return OF_OK;
  ddpm_x0_kernel<<<grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, pred, noise, log_snr, log_snr_next,
                                                                                n, do_sign);
  OF_LAUNCH_CHECK("of_ddpm_x0_update");
  return OF_OK;
}

extern "C" int of_copy_rows(const void* src, int64_t lds, int32_t src_dtype, const int32_t* src_rows, void* dst,
                            int64_t ldd, int32_t dst_dtype, const int32_t* dst_rows, int64_t rows, int32_t c,
                            void* stream) {
  OF_REQUIRE(src && dst && rows >= 0 && c > 0, "of_copy_rows: bad arguments");
  if (rows == 0) return OF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int esz = src_dtype == OF_F32 ? 4 : 2;
  if (src_dtype == dst_dtype && (c * esz) % 16 == 0 && (lds * esz) % 16 == 0 && (ldd * esz) % 16 == 0 &&
      reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0) {
    const int c16 = c * esz / 16;
    copy_rows_vec_kernel<<<grid_for(rows * c16), 256, 0, st>>>(reinterpret_cast<const uint4*>(src), lds * esz / 16,
                                                               src_rows, reinterpret_cast<uint4*>(dst),
                                                               ldd * esz / 16, dst_rows, rows, c16);
  } else {
    const int g = grid_for(rows * c);
    if (src_dtype == OF_F32 && dst_dtype == OF_F32)
      copy_rows_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, lds, src_rows, (float*)dst, ldd, dst_rows, rows, c);
    else if (src_dtype == OF_F32 && dst_dtype == OF_BF16)
      copy_rows_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)src, lds, src_rows, (__nv_bfloat16*)dst, ldd, dst_rows, rows, c);
    else if (src_dtype == OF_BF16 && dst_dtype == OF_F32)
      copy_rows_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, lds, src_rows, (float*)dst, ldd, dst_rows, rows, c);
    else
      copy_rows_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, lds, src_rows, (__nv_bfloat16*)dst, ldd, dst_rows, rows, c);
  }
  OF_LAUNCH_CHECK("of_copy_rows");
  return OF_OK;
}

extern "C" int of_histogram_i32(const int32_t* values, int64_t n, int32_t bins, int32_t* hist, void* stream) {
  OF_REQUIRE(values && hist && bins > 0 && bins <= 8192 && n >= 0, "of_histogram_i32: bad arguments");
  if (n == 0) return OF_OK;
  histogram_kernel<<<grid_for(n, 256, 4), 256, bins * sizeof(int32_t), reinterpret_cast<cudaStream_t>(stream)>>>(
      values, n, bins, hist);
  OF_LAUNCH_CHECK("of_histogram_i32");
  return OF_OK;
}

extern "C" int of_dense_tap_table(int32_t mode, int32_t out_res_log2, int32_t batch, int32_t* tap_tab,
                                  void* stream) {
  OF_REQUIRE(mode >= 0 && mode <= 2 && out_res_log2 >= (mode == 2 ? 1 : 0) && out_res_log2 <= 9 && batch > 0 &&
                 tap_tab,
             "of_dense_tap_table: bad arguments");
  const int64_t n = ((int64_t)batch << (3 * out_res_log2)) * 27;
  dense_tap_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      mode, out_res_log2, batch, tap_tab);
  OF_LAUNCH_CHECK("of_dense_tap_table");
  return OF_OK;
}
