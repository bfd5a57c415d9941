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

template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_kernel(const T* __restrict__ qkv, int64_t ld_qkv,
                                                                   T* __restrict__ out, int64_t ld_out, int tokens,
                                                                   int heads, int ch) {
  extern __shared__ float sm[];
  const int kst = ch + 1;                               // padded row stride: conflict-free column reads
  float* Ks = sm;                                       // [T][ch+1]
  float* Vs = Ks + (size_t)tokens * kst;                // [T][ch+1]
  float* Ps = Vs + (size_t)tokens * kst;                // [warps][T]
  float* Qs = Ps + (size_t)ATT_WARPS * tokens;          // [warps][ch]
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
  float* P = Ps + (size_t)warp * tokens;
  float* Q = Qs + (size_t)warp * ch;
  const float scale = rsqrtf((float)ch);                // (ch^-1/4)^2, modules.py:542-545
  const int q_end = min(tokens, (int)(blockIdx.y + 1) * ATT_QTILE);
  for (int t = blockIdx.y * ATT_QTILE + warp; t < q_end; t += ATT_WARPS) {
    const T* qr = qkv + (row0 + t) * ld_qkv + colq;
    for (int c = lane; c < ch; c += 32) Q[c] = Elem<T>::ld(qr + c) * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int s = lane; s < tokens; s += 32) {
      const float* kr = Ks + s * kst;
      float d = 0.0f;
      for (int c = 0; c < ch; ++c) d = fmaf(Q[c], kr[c], d);
      P[s] = d;
      mx = fmaxf(mx, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
    for (int s = lane; s < tokens; s += 32) {
      const float e = __expf(P[s] - mx);
      P[s] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.0f / sum;
    T* orow = out + (row0 + t) * ld_out + h * ch;
    for (int c = lane; c < ch; c += 32) {
      float a = 0.0f;
      for (int s = 0; s < tokens; ++s) a = fmaf(P[s], Vs[s * kst + c], a);
      Elem<T>::st(orow + c, a * inv);
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
  const size_t smem = ((size_t)2 * tokens * (ch + 1) + (size_t)ATT_WARPS * tokens + (size_t)ATT_WARPS * ch) * 4;
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
