// QKVAttention.forward (reference models/networks/modules.py:538-547) over the dense voxel
// tokens of the LR middle U-Net: T in {512, 64, 8} tokens per shape, ch in {16, 32, 64, 128}.
// The reference materialises the [b*h, T, T] fp32 score tensor (134 MB at B=32); here one CTA
// keeps K and V of one (shape, head) in shared memory and streams the queries through it --
// scores never leave the SM.  <0.2 % of the step FLOPs, so this first
// version runs on the CUDA cores with fp32 softmax exactly as the reference (modules.py:546).
//
// qkv is channels-last [B*T, 3C] with the reference's legacy head-major split: head h owns
// columns [h*3ch, (h+1)*3ch) = q | k | v (modules.py:531,540-541).
#include "common.cuh"

namespace of {

constexpr int ATT_WARPS = 8;
constexpr int ATT_QTILE = 128;
constexpr int ATT_QB = 4;                 // queries processed together by one warp (register blocking)

template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_kernel(const T* __restrict__ qkv, int64_t ld_qkv,
                                                                   T* __restrict__ out, int64_t ld_out, int tokens,
                                                                   int heads, int ch) {
  extern __shared__ __align__(16) float sm[];
  const int kst = ch + 4;                               // 16-byte aligned rows, conflict-free float4 column reads
  float* Ks = sm;                                       // [T][ch+4]
  float* Vs = Ks + (size_t)tokens * kst;                // [T][ch+4]
  float* Ps = Vs + (size_t)tokens * kst;                // [warps][T][QB]
  float* Qs = Ps + (size_t)ATT_WARPS * tokens * ATT_QB; // [warps][QB][ch]
  const int bh = blockIdx.x;
  const int b = bh / heads, h = bh - b * heads;
  const int64_t row0 = (int64_t)b * tokens;
  const int colq = h * 3 * ch, colk = colq + ch, colv = colq + 2 * ch;
  for (int i = threadIdx.x; i < tokens * ch; i += blockDim.x) {
    const int s = i / ch, c = i - s * ch;
    const T* r = qkv + (row0 + s) * ld_qkv;
    Ks[s * kst + c] = Elem<T>::ld(r + colk + c);
    Vs[s * kst + c] = Elem<T>::ld(r + colv + c);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* P = Ps + (size_t)warp * tokens * ATT_QB;       // P[s][q]
  float* Q = Qs + (size_t)warp * ATT_QB * ch;           // Q[q][c]
  const float scale = rsqrtf((float)ch);                // (ch^-1/4)^2, modules.py:542-545
  const int q_end = min(tokens, (int)(blockIdx.y + 1) * ATT_QTILE);
  for (int t0 = blockIdx.y * ATT_QTILE + warp * ATT_QB; t0 < q_end; t0 += ATT_WARPS * ATT_QB) {
    const int nq = min(ATT_QB, q_end - t0);
    for (int i = lane; i < ATT_QB * ch; i += 32) {
      const int qi = i / ch, c = i - qi * ch;
      Q[i] = qi < nq ? Elem<T>::ld(qkv + (row0 + t0 + qi) * ld_qkv + colq + c) * scale : 0.0f;
    }
    __syncwarp();
    float mx[ATT_QB];
#pragma unroll
    for (int qi = 0; qi < ATT_QB; ++qi) mx[qi] = -INFINITY;
    for (int s = lane; s < tokens; s += 32) {
      const float4* kr = reinterpret_cast<const float4*>(Ks + s * kst);
      float d[ATT_QB];
#pragma unroll
      for (int qi = 0; qi < ATT_QB; ++qi) d[qi] = 0.0f;
      for (int c4 = 0; c4 < ch / 4; ++c4) {
        const float4 kv = kr[c4];
#pragma unroll
        for (int qi = 0; qi < ATT_QB; ++qi) {
          const float4 qv = reinterpret_cast<const float4*>(Q + qi * ch)[c4];
          d[qi] = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, fmaf(qv.z, kv.z, fmaf(qv.w, kv.w, d[qi]))));
        }
      }
      *reinterpret_cast<float4*>(P + s * ATT_QB) = make_float4(d[0], d[1], d[2], d[3]);
#pragma unroll
      for (int qi = 0; qi < ATT_QB; ++qi) mx[qi] = fmaxf(mx[qi], d[qi]);
    }
    float sum[ATT_QB];
#pragma unroll
    for (int qi = 0; qi < ATT_QB; ++qi) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx[qi] = fmaxf(mx[qi], __shfl_xor_sync(0xffffffffu, mx[qi], o));
      sum[qi] = 0.0f;
    }
    for (int s = lane; s < tokens; s += 32) {
      float4 pv = *reinterpret_cast<float4*>(P + s * ATT_QB);
      pv.x = __expf(pv.x - mx[0]); pv.y = __expf(pv.y - mx[1]); pv.z = __expf(pv.z - mx[2]); pv.w = __expf(pv.w - mx[3]);
      *reinterpret_cast<float4*>(P + s * ATT_QB) = pv;
      sum[0] += pv.x; sum[1] += pv.y; sum[2] += pv.z; sum[3] += pv.w;
    }
#pragma unroll
    for (int qi = 0; qi < ATT_QB; ++qi) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum[qi] += __shfl_xor_sync(0xffffffffu, sum[qi], o);
    }
    __syncwarp();
    for (int c = lane; c < ch; c += 32) {
      float a[ATT_QB];
#pragma unroll
      for (int qi = 0; qi < ATT_QB; ++qi) a[qi] = 0.0f;
      for (int s = 0; s < tokens; ++s) {
        const float v = Vs[s * kst + c];
        const float4 pv = *reinterpret_cast<const float4*>(P + s * ATT_QB);
        a[0] = fmaf(pv.x, v, a[0]); a[1] = fmaf(pv.y, v, a[1]); a[2] = fmaf(pv.z, v, a[2]); a[3] = fmaf(pv.w, v, a[3]);
      }
#pragma unroll
      for (int qi = 0; qi < ATT_QB; ++qi)
        if (qi < nq) Elem<T>::st(out + (row0 + t0 + qi) * ld_out + h * ch + c, a[qi] / sum[qi]);
    }
    __syncwarp();
  }
}

}  // namespace of

extern "C" int of_attention(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int32_t batch,
                            int32_t tokens, int32_t heads, int32_t ch, int32_t dtype, void* stream) {
  using namespace of;
  OF_REQUIRE(qkv && out && batch > 0 && tokens > 0 && heads > 0 && ch > 0, "of_attention: bad arguments");
  OF_REQUIRE(dtype == OF_F32 || dtype == OF_BF16, "of_attention: bad dtype");
  OF_REQUIRE(ch % 4 == 0, "of_attention: ch must be a multiple of 4");
  const size_t smem = ((size_t)2 * tokens * (ch + 4) + (size_t)ATT_WARPS * tokens * ATT_QB + (size_t)ATT_WARPS * ATT_QB * ch) * 4;
  if (smem > 220 * 1024) {
    set_error("of_attention: T=%d ch=%d needs %zu B of shared memory (> 220 KB)", tokens, ch, smem);
    return OF_E_UNSUPPORTED;
  }
  dim3 grid(batch * heads, (tokens + ATT_QTILE - 1) / ATT_QTILE);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == OF_F32) {
    static size_t cfg_f32 = 48 * 1024;
    if (smem > cfg_f32) {
      cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      cfg_f32 = 220 * 1024;
    }
    attention_kernel<float><<<grid, ATT_WARPS * 32, smem, st>>>((const float*)qkv, ld_qkv, (float*)out, ld_out, tokens,
                                                                heads, ch);
  } else {
    static size_t cfg_bf16 = 48 * 1024;
    if (smem > cfg_bf16) {
      cudaFuncSetAttribute(attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      cfg_bf16 = 220 * 1024;
    }
    attention_kernel<__nv_bfloat16><<<grid, ATT_WARPS * 32, smem, st>>>((const __nv_bfloat16*)qkv, ld_qkv,
                                                                        (__nv_bfloat16*)out, ld_out, tokens, heads, ch);
  }
  OF_LAUNCH_CHECK("of_attention");
  return OF_OK;
}
