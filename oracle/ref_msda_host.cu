// TEST / BENCH INFRASTRUCTURE — not part of the product (only tests/, __graft_entry__.smoke() and bench.py use oracle/).
//
// C-ABI host shim around the REFERENCE's own MSDeformAttn forward kernel, compiled from the reference sources where they
// lie (third_party/Mask2Former/mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh:242-304 kernel,
// :928-958 launcher).  The reference's host file (ms_deform_attn_cuda.cu:25-85) does not compile against torch 2.11
// (AT_DISPATCH_FLOATING_TYPES(value.type(), ...): DeprecatedTypeProperties -> ScalarType conversion was removed), so
// this file restates its batching loop (im2col_step chunks, :57-78) and calls the UNMODIFIED kernel launcher.
// Built by oracle/Makefile into oracle/_ref/libref_msda.so; used as the GPU baseline of the C4 microbench and as a second
// parity oracle for odise_msda_forward_f32.
#include "cuda/ms_deform_im2col_cuda.cuh"

extern "C" int ref_ms_deform_attn_forward_f32(const float* value, const int64_t* spatial_shapes,
                                              const int64_t* level_start, const float* loc, const float* attn,
                                              float* out, int N, int S, int M, int D, int L, int Lq, int P,
                                              int im2col_step, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int step = N < im2col_step ? N : im2col_step;
  if (step <= 0 || N % step) return 1;                            // reference: AT_ASSERTM(batch % im2col_step_ == 0)
  const long long per_value = (long long)S * M * D, per_loc = (long long)Lq * M * L * P * 2,
                  per_attn = (long long)Lq * M * L * P, per_out = (long long)Lq * M * D;
  cudaMemsetAsync(out, 0, sizeof(float) * N * per_out, stream);   // reference: at::zeros
  for (int n = 0; n < N / step; ++n)
    ms_deformable_im2col_cuda<float>(stream, value + n * step * per_value, spatial_shapes, level_start,
                                     loc + n * step * per_loc, attn + n * step * per_attn, step, S, M, D, L, Lq, P,
                                     out + n * step * per_out);
  return (int)cudaGetLastError();
}
