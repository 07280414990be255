"""Oracle: multi-scale deformable attention forward (CPU, torch fp32/fp64).  TEST INFRASTRUCTURE ONLY.

Restates the reference CUDA kernel ms_deformable_im2col_gpu_kernel + ms_deform_attn_im2col_bilinear
(third_party/Mask2Former/mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh:242-304, :38-89)
as explicit gathers, i.e. independently of F.grid_sample.  Pinned: tools/make_golden.py checks it against the
reference's own PyTorch restatement ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:52-72,
imported from /root/reference) on the reference test problem (ops/test.py:24-39) and on release shapes, and stores
the vectors in tests/golden/msda_*.pt.
"""
import torch


def msda_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """value [N,S,M,D]; spatial_shapes [L,2] (H,W); sampling_locations [N,Lq,M,L,P,2] (x,y); attn [N,Lq,M,L,P]
    -> [N, Lq, M*D]  (cuda/ms_deform_attn_cuda.cu:25-85 output layout)."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    out = value.new_zeros(N, Lq, M, D)
    n_idx = torch.arange(N).view(N, 1, 1, 1)
    m_idx = torch.arange(M).view(1, 1, M, 1)
    for l in range(L):
        H, W = int(spatial_shapes[l][0]), int(spatial_shapes[l][1])
        start = int(level_start_index[l])
        v = value[:, start:start + H * W]                              # [N, HW, M, D]
        loc = sampling_locations[:, :, :, l]                           # [N, Lq, M, P, 2]
        aw = attention_weights[:, :, :, l]                             # [N, Lq, M, P]
        h_im = loc[..., 1] * H - 0.5                                   # .cuh:290
        w_im = loc[..., 0] * W - 0.5                                   # .cuh:291
        inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)   # .cuh:293
        h_low = torch.floor(h_im)
        w_low = torch.floor(w_im)
        lh, lw = h_im - h_low, w_im - w_low
        hh, hw = 1 - lh, 1 - lw
        h_low, w_low = h_low.long(), w_low.long()
        acc = 0
        for dy, dx, wgt in ((0, 0, hh * hw), (0, 1, hh * lw), (1, 0, lh * hw), (1, 1, lh * lw)):  # .cuh:60-87
            yy, xx = h_low + dy, w_low + dx
            ok = inside & (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
            idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1))        # [N, Lq, M, P]
            g = v[n_idx, idx, m_idx]                                   # [N, Lq, M, P, D]
            acc = acc + g * (wgt * ok.to(value.dtype)).unsqueeze(-1)
        out = out + (acc * aw.unsqueeze(-1)).sum(dim=3)
    return out.reshape(N, Lq, M * D)


def msdeformattn_front(query, reference_points, sampling_offsets, attention_logits, spatial_shapes, M, L, P):
    """Softmax + sampling-location math of MSDeformAttn.forward (ops/modules/ms_deform_attn.py:98-107)."""
    N, Lq, _ = sampling_offsets.shape
    off = sampling_offsets.view(N, Lq, M, L, P, 2)
    aw = torch.softmax(attention_logits.view(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1).to(off.dtype)
    loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    return loc, aw
