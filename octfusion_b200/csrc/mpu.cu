// NeuralMPU evaluation: the implicit function the GraphVAE decoder defines.
//
// Replaces reference models/networks/dualoctree_networks/mpu.py:55-140 (`octree_linear_pts` for every depth,
// `get_linear_pred`): per query point and depth the reference builds 8 corner keys, looks them up with
// `octree.search_key`, filters, and pushes everything through two sparse-matrix products (`modulated_spmm`, `spmm`,
// utils/spmm.py).  Here one thread owns one query point: for each depth it locates the (up to) 8 surrounding cells by
// walking `children` down from the full layer (no key search, no intermediate tensors), blends
//   w * (F . [offset, 1]),  w = prod(1 - |offset in cells|) * d^2 / 50            (mpu.py:88-97)
// over existing cells (leaves only below the target depth, mpu.py:118-121) and normalises by the weight sum (:135-137).
// Latency / L2-bound integer walk + 16-byte reads of the per-node regression values.
#include "common.cuh"

namespace of {

struct MpuCtx {
  const int32_t* children[16];
  int32_t nnum[16];
  int64_t row_off[16];            // offset of depth d inside the padded per-node array (nodes of depths fd..D)
  int32_t fd, D, batch;
};

// regular sampling grid of calc_sdf (reference utils/util_dualoctree.py:99-118, get_mgrid :23-42): point p of the
// size^3 grid = (p / size^2, (p / size) % size, p % size) * ((bbmax - bbmin) / size) + bbmin, in fp32 with separate
// roundings like the numpy expression
struct MpuGrid { int size; float step, bbmin; int batch_idx; int64_t head; };

template <bool GRID>
__global__ void __launch_bounds__(256) mpu_eval_kernel(MpuCtx c, const float* __restrict__ pos, MpuGrid gr, int64_t npts,
                                                       const float4* __restrict__ reg, float* __restrict__ fval,
                                                       uint8_t* __restrict__ touched) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  float4 q;
  if (GRID) {
    const int64_t lin = gr.head + p;
    const int iz = (int)(lin % gr.size), iy = (int)((lin / gr.size) % gr.size), ix = (int)(lin / ((int64_t)gr.size * gr.size));
    q.x = __fadd_rn(__fmul_rn((float)ix, gr.step), gr.bbmin);
    q.y = __fadd_rn(__fmul_rn((float)iy, gr.step), gr.bbmin);
    q.z = __fadd_rn(__fmul_rn((float)iz, gr.step), gr.bbmin);
    q.w = (float)gr.batch_idx;
  } else {
    q = reinterpret_cast<const float4*>(pos)[p];
  }
  const int b = (int)q.w;
  float num = 0.0f, den = 0.0f;
  bool hit = false;
  if (b >= 0 && b < c.batch) {
    for (int d = c.fd; d <= c.D; ++d) {
      const int scale = 1 << d;
      const float half = 0.5f * (float)scale;
      const float xf = (q.x + 1.0f) * half - 0.5f, yf = (q.y + 1.0f) * half - 0.5f, zf = (q.z + 1.0f) * half - 0.5f;
      const float xi = floorf(xf), yi = floorf(yf), zi = floorf(zf);
      const float wd = (float)(d * d) / 50.0f;
      const float back = 2.0f / (float)scale;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int cx = (int)xi + ((k >> 2) & 1), cy = (int)yi + ((k >> 1) & 1), cz = (int)zi + (k & 1);
        if (cx < 0 || cy < 0 || cz < 0 || cx >= scale || cy >= scale || cz >= scale) continue;
        // locate the depth-d cell (cx, cy, cz): the full layer is indexed by its Morton key, then follow `children`
        const int sh0 = d - c.fd;
        int fx = cx >> sh0, fy = cy >> sh0, fz = cz >> sh0, key = 0;
        for (int i = 0; i < c.fd; ++i)
          key |= (((fx >> i) & 1) << (3 * i + 2)) | (((fy >> i) & 1) << (3 * i + 1)) | (((fz >> i) & 1) << (3 * i));
        int ci = (b << (3 * c.fd)) + key;
        int cd = c.fd;
        bool found = true;
        while (cd < d) {
          const int ch = c.children[cd][ci];
          if (ch < 0) { found = false; break; }
          const int sh = d - cd - 1;
          ci = 8 * ch + ((((cx >> sh) & 1) << 2) | (((cy >> sh) & 1) << 1) | ((cz >> sh) & 1));
          ++cd;
        }
        if (!found) continue;
        if (d == c.D) hit = true;
        else if (c.children[d][ci] >= 0) continue;                      // only leaves below the target depth
        const float ox = xf - (float)cx, oy = yf - (float)cy, oz = zf - (float)cz;
        const float w = (1.0f - fabsf(ox)) * (1.0f - fabsf(oy)) * (1.0f - fabsf(oz)) * wd;
        const float4 f = reg[c.row_off[d] + ci];
        num = fmaf(w, fmaf(f.x, ox * back, fmaf(f.y, oy * back, fmaf(f.z, oz * back, f.w))), num);
        den += w;
      }
    }
  }
  fval[p] = num / (den + 1e-8f);
  if (touched != nullptr) touched[p] = hit ? 1 : 0;
}

static int make_mpu_ctx(const of_octree_levels* oct, int32_t depth, MpuCtx& c, const char* who) {
  OF_REQUIRE(oct->full_depth >= 1 && depth >= oct->full_depth && depth <= oct->depth && depth < 16,
             "%s: depth %d outside [%d, %d]", who, depth, oct->full_depth, oct->depth);
  c.fd = oct->full_depth; c.D = depth; c.batch = oct->batch;
  int64_t off = 0;
  for (int d = 0; d < 16; ++d) { c.children[d] = oct->children[d]; c.nnum[d] = oct->nnum[d]; c.row_off[d] = 0; }
  for (int d = c.fd; d <= depth; ++d) {
    OF_REQUIRE(oct->children[d] != nullptr, "%s: children[%d] missing", who, d);
    c.row_off[d] = off;
    off += oct->nnum[d];
  }
  return OF_OK;
}

}  // namespace of

extern "C" int of_mpu_eval(const of_octree_levels* oct, int32_t depth, const float* pos, int64_t npts,
                           const float* reg, float* fval, uint8_t* touched, void* stream) {
  using namespace of;
  OF_REQUIRE(oct && pos && reg && fval && touched && npts >= 0, "of_mpu_eval: null pointer / negative count");
  OF_REQUIRE(reinterpret_cast<uintptr_t>(pos) % 16 == 0 && reinterpret_cast<uintptr_t>(reg) % 16 == 0,
             "of_mpu_eval: pos / reg must be 16-byte aligned ([*, 4] fp32 rows)");
  MpuCtx c;
  int rc = make_mpu_ctx(oct, depth, c, "of_mpu_eval");
  if (rc) return rc;
  if (npts == 0) return OF_OK;
  mpu_eval_kernel<false><<<(unsigned)((npts + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      c, pos, MpuGrid{}, npts, reinterpret_cast<const float4*>(reg), fval, touched);
  OF_LAUNCH_CHECK("of_mpu_eval");
  return OF_OK;
}

extern "C" int of_mpu_eval_grid(const of_octree_levels* oct, int32_t depth, int32_t batch_idx, int32_t size, float bbmin,
                                float bbmax, int64_t head, int64_t count, const float* reg, float* fval, void* stream) {
  using namespace of;
  OF_REQUIRE(oct && reg && fval && size > 0 && head >= 0 && count >= 0 && head + count <= (int64_t)size * size * size,
             "of_mpu_eval_grid: bad arguments");
  OF_REQUIRE(reinterpret_cast<uintptr_t>(reg) % 16 == 0, "of_mpu_eval_grid: reg must be 16-byte aligned");
  MpuCtx c;
  int rc = make_mpu_ctx(oct, depth, c, "of_mpu_eval_grid");
  if (rc) return rc;
  if (count == 0) return OF_OK;
  MpuGrid gr;
  gr.size = size; gr.batch_idx = batch_idx; gr.head = head;
  gr.step = (float)(((double)bbmax - (double)bbmin) / (double)size);      // python float (double) -> float32 scalar
  gr.bbmin = bbmin;
  mpu_eval_kernel<true><<<(unsigned)((count + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      c, nullptr, gr, count, reinterpret_cast<const float4*>(reg), fval + head, nullptr);
  OF_LAUNCH_CHECK("of_mpu_eval_grid");
  return OF_OK;
}
