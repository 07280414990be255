// Warp-shuffle attention kernels of the Mask2Former transformer decoder, sm_100a.
//
// Reference: nn.MultiheadAttention(256, 8) inside CrossAttentionLayer / SelfAttentionLayer
// (mask2former_transformer_decoder.py:22,80,98-110,40-50) fed a materialised bool attn_mask [B*8, Q, HW] that
// ODISE builds per layer from the previous head's mask logits (odise.py:760-774) plus the
// "row fully masked => unmask" fix-up (odise.py:683).  Here:
//   * attn_mask_bits: one pass over the mask logits -> bilinear resize to the level size (align_corners=False),
//     sigmoid < 0.5 test, 1 bit per (b, q, key) shared by all 8 heads (the reference repeats the mask per head),
//     and a per-row "any key allowed" flag that implements the fix-up without touching the bits.
//   * mha_d32: flash-style (online softmax) attention for head_dim 32: a warp owns a query, lanes own keys for the
//     score pass and channels for the P.V pass (shuffle broadcast of p), K/V tiles staged in shared memory.
#include "ptx.cuh"
#include "odise_b200.h"
#include "launch_count.h"

namespace ob {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block per (b, q) row; bits[b][q][w] bit i = key (w*32+i) may be attended
__global__ void __launch_bounds__(256)
attn_mask_bits_kernel(const float* __restrict__ logits, uint32_t* __restrict__ bits, int32_t* __restrict__ row_any,
                      int Hm, int Wm, int Hl, int Wl) {
  const long long row = blockIdx.x;
  const float* src = logits + row * (long long)Hm * Wm;
  const int HW = Hl * Wl;
  const int words = (HW + 31) / 32;
  uint32_t* dst = bits + row * words;
  const float sy = (float)Hm / (float)Hl, sx = (float)Wm / (float)Wl;
  int any = 0;
  for (int base = (threadIdx.x >> 5) * 32; base < words * 32; base += (blockDim.x >> 5) * 32) {
    const int key = base + (threadIdx.x & 31);
    bool allowed = false;
    if (key < HW) {
      const int oy = key / Wl, ox = key - oy * Wl;
      // F.interpolate(bilinear, align_corners=False) source index (ATen area_pixel_compute_source_index)
      const float fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.f);
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = y0 + (y0 < Hm - 1 ? 1 : 0), x1 = x0 + (x0 < Wm - 1 ? 1 : 0);
      const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
      const float v = hy * (hx * src[y0 * Wm + x0] + lx * src[y0 * Wm + x1]) +
                      ly * (hx * src[y1 * Wm + x0] + lx * src[y1 * Wm + x1]);
      const float s = 1.f / (1.f + expf(-v));
      allowed = !(s < 0.5f);  // attn_mask = sigmoid(.) < 0.5 means "blocked"
    }
    const uint32_t w = __ballot_sync(0xffffffffu, allowed);
    if ((threadIdx.x & 31) == 0) dst[base >> 5] = w;
    any |= (w != 0);
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) row_any[row] = any;
}

// grid (q_tiles, heads, B); 8 warps, QPW queries per warp; head_dim 32
template <int QPW>
__global__ void __launch_bounds__(256)
mha_d32_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k,
               const float* __restrict__ v, long long ldkv, const uint32_t* __restrict__ bits,
               const int32_t* __restrict__ row_any, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
               __nv_bfloat16* __restrict__ out_lo, long long ldo, int Tq, int Tk, int heads, float scale) {
  constexpr int KT = 128;                // keys per shared-memory tile
  __shared__ float Ks[KT][33];
  __shared__ float Vs[KT][32];
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int words = (Tk + 31) / 32;

  float qreg[QPW][32];
  float m[QPW], l[QPW], o[QPW];
  int qi[QPW];
  bool use_mask[QPW];
#pragma unroll
  for (int t = 0; t < QPW; ++t) {
    qi[t] = (blockIdx.x * 8 + warp) * QPW + t;
    m[t] = -INFINITY; l[t] = 0.f; o[t] = 0.f;
    use_mask[t] = false;
    if (qi[t] < Tq) {
      const float* qp = q + ((long long)b * Tq + qi[t]) * ldq + h * 32;
#pragma unroll
      for (int d = 0; d < 32; ++d) qreg[t][d] = __ldg(qp + d) * scale;
      if (bits) use_mask[t] = row_any[(long long)b * Tq + qi[t]] != 0;
    } else {
#pragma unroll
      for (int d = 0; d < 32; ++d) qreg[t][d] = 0.f;
    }
  }

  for (int k0 = 0; k0 < Tk; k0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * 8; i += 256) {
      const int r = i >> 3, c4 = (i & 7) * 4;
      float4 kv = make_float4(0, 0, 0, 0), vv = kv;
      if (k0 + r < Tk) {
        kv = *reinterpret_cast<const float4*>(k + ((long long)b * Tk + k0 + r) * ldkv + h * 32 + c4);
        vv = *reinterpret_cast<const float4*>(v + ((long long)b * Tk + k0 + r) * ldkv + h * 32 + c4);
      }
      Ks[r][c4] = kv.x; Ks[r][c4 + 1] = kv.y; Ks[r][c4 + 2] = kv.z; Ks[r][c4 + 3] = kv.w;
      *reinterpret_cast<float4*>(&Vs[r][c4]) = vv;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < QPW; ++t) {
      if (qi[t] >= Tq) continue;  // warp-uniform
#pragma unroll 1
      for (int sb = 0; sb < KT; sb += 32) {
        const int key = k0 + sb + lane;
        if (k0 + sb >= Tk) break;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) s = fmaf(qreg[t][d], Ks[sb + lane][d], s);
        bool ok = key < Tk;
        if (use_mask[t]) {
          const uint32_t w = __ldg(bits + ((long long)b * Tq + qi[t]) * words + ((k0 + sb) >> 5));
          ok = ok && ((w >> lane) & 1u);
        }
        s = ok ? s : -INFINITY;
        const float mnew = fmaxf(m[t], wmax(s));
        if (mnew == -INFINITY) continue;  // nothing attendable yet in this row
        const float alpha = __expf(m[t] - mnew);  // m = -inf -> 0
        const float p = ok ? expf(s - mnew) : 0.f;
        l[t] = l[t] * alpha + wsum(p);
        float acc = o[t] * alpha;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaf(__shfl_sync(0xffffffffu, p, j), Vs[sb + j][lane], acc);
        o[t] = acc;
        m[t] = mnew;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < QPW; ++t) {
    if (qi[t] >= Tq) continue;
    const float r = o[t] / l[t];
    const long long idx = ((long long)b * Tq + qi[t]) * ldo + h * 32 + lane;
    if (out) out[idx] = r;
    if (out_hi) {
      __nv_bfloat16 hh, ll;
      split_bf16(r, hh, ll);
      out_hi[idx] = hh;
      if (out_lo) out_lo[idx] = ll;
    }
  }
}

}  // namespace ob

using namespace ob;

extern "C" int odise_attn_mask_bits_f32(const float* mask_logits, uint32_t* bits, int32_t* row_any, int B, int Q,
                                        int Hm, int Wm, int Hl, int Wl, void* stream) {
  if (!mask_logits || !bits || !row_any || B <= 0 || Q <= 0 || Hm <= 0 || Wm <= 0 || Hl <= 0 || Wl <= 0)
    return ODISE_ERR_ARG;
  attn_mask_bits_kernel<<<B * Q, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(mask_logits, bits, row_any, Hm,
                                                                                  Wm, Hl, Wl);
  count_launch(1);
  return (int)cudaGetLastError();
}

extern "C" int odise_mha_d32_f32(const float* q, long long ldq, const float* k, const float* v, long long ldkv,
                                 const uint32_t* bits, const int32_t* row_any, float* out, void* out_hi,
                                 void* out_lo, long long ldo, int B, int Tq, int Tk, int heads, float scale,
                                 void* stream) {
  if (!q || !k || !v || (!out && !out_hi) || B <= 0 || Tq <= 0 || Tk <= 0 || heads <= 0) return ODISE_ERR_ARG;
  if (ldq < heads * 32 || ldkv < heads * 32 || ldo < heads * 32 || ldkv % 4) return ODISE_ERR_ALIGN;
  if (bits && !row_any) return ODISE_ERR_ARG;
  constexpr int QPW = 2;
  dim3 grid((Tq + 8 * QPW - 1) / (8 * QPW), heads, B);
  mha_d32_kernel<QPW><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      q, ldq, k, v, ldkv, bits, row_any, out, reinterpret_cast<__nv_bfloat16*>(out_hi),
      reinterpret_cast<__nv_bfloat16*>(out_lo), ldo, Tq, Tk, heads, scale);
  count_launch(1);
  return (int)cudaGetLastError();
}
