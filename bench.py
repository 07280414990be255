#!/usr/bin/env python
"""Benchmark of the ODISE inference hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle), rank 0 only

A step = one pass of the pipeline (KL-VAE encoder/decoder taps [unless --hot-path-only] -> UNet feature pass on all
512^2 crops -> projections -> pixel decoder -> masked attention decoder -> CLIP-text scoring) over a batch of
synthetic 1024x1024 images, ADE-150 vocabulary (K' = 403
prompts), random-init weights (BASELINE.json configs[1]).  `value` times the CUDA-graph replay with inputs resident in
HBM; `e2e` times the public call ODISEEngine.infer() with pinned-host uint8 images in and fp32 logits + mask logits out.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec @1024x1024 panoptic inference (ODISE hot path)"
UNET_TFLOP_PER_CROP = 0.740          # minimal feature pass, SURVEY.md §8d / BASELINE.md §2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16"])
    ap.add_argument("--vocab", default="ade150", choices=["ade150", "coco133", "ade847"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hot-path-only", action="store_true",
                    help="skip the KL-VAE (SURVEY.md 8f-1) and CLIP image tower (8f-2) stages: their taps / latent / "
                         "image embedding enter as synthetic tensors")
    return ap.parse_args()


VOCABS = {"ade150": (150, 403), "coco133": (133, 254), "ade847": (847, 1342)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# ----------------------------------------------------------------------------------------------- CPU baseline
class CpuHotPath:
    """The reference's CPU path through the oracle (oracle/ldm.py restatement driven like LdmExtractor.unet_forward;
    oracle/m2f.py == the reference's own Mask2Former/ODISE code, pinned in tests/test_oracle_cpu.py), fp32.
    Bounded sample: ONE 512^2 crop through the UNet feature pass + ONE image through the head at size^2 (pixel
    decoder + decoder + scoring); images/s = 1 / (crops * t_unet + t_head)."""

    def __init__(self, size, vocab, threads=None, full=True):
        from odise_b200 import spec
        from odise_b200.pipeline import synthetic_vocabulary
        from oracle import ldm, m2f
        self.ldm, self.m2f = ldm, m2f
        # torch CPU ops stop scaling (and regress) far below the 100+ threads of the GPU hosts: cap at 32
        self.n = threads or min(os.cpu_count(), 32)
        torch.set_num_threads(self.n)
        sd_u = spec.synth_state_dict(spec.unet_params(), 0)
        with torch.device("meta"):
            unet = ldm.UNetModel()
        unet.load_state_dict({k[len(spec.UNET_PREFIX):]: v for k, v in sd_u.items()}, assign=True)
        self.unet = unet.eval()
        self.sd_h = spec.synth_state_dict(spec.head_params(), 1)
        g = torch.Generator().manual_seed(3)
        self.x, self.ctx = torch.randn(1, 4, 64, 64, generator=g), torch.randn(1, 77, 768, generator=g)
        self.cond = torch.randn(1, 1280, generator=g)
        self.feats = {f"s{i}": torch.randn(1, 512, size // 2 ** i, size // 2 ** i, generator=g) for i in (2, 3, 4, 5)}
        self.bank, self.null, self.sizes = synthetic_vocabulary(*VOCABS[vocab])
        self.size = size
        self.crops = max(1, (size // 512)) ** 2
        self.full = full
        if full:
            sd_v = spec.synth_state_dict(spec.vae_params(), 3)
            with torch.device("meta"):
                vae = ldm.AutoencoderKL()
            vae.load_state_dict({k[len(spec.VAE_PREFIX):]: v for k, v in sd_v.items()}, assign=True)
            self.vae = vae.eval()
            self.img = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1
            from oracle import clip as oclip
            self.oclip = oclip
            sd_c = spec.synth_state_dict(spec.clip_visual_params(), 5)
            with torch.device("meta"):
                vis = oclip.VisionTransformer()
            vis.load_state_dict({k[len(spec.CLIP_PREFIX):]: v for k, v in sd_c.items()}, assign=True)
            self.vis = vis.eval()
            self.img_full = torch.rand(1, 3, size, size, generator=g)
            self.ov = torch.tensor([(k % 2) == 0 for k in range(VOCABS[vocab][0])]).long()

    @torch.no_grad()
    def sample(self):
        ldm, m2f = self.ldm, self.m2f
        t0 = time.perf_counter()
        # as the reference executes it: output block 11 + unet.out run too, and the VAE decoder runs to the full image
        ldm.unet_features(self.unet, self.x, self.ctx, self.cond, stop_early=False)
        if self.full:
            lat, _ = ldm.encoder_features(self.vae, self.img)
            ldm.decoder_features(self.vae, lat, truncate=False)
            self.oclip.embed_image(self.vis, self.img * 0.5 + 0.5)
        t_unet = time.perf_counter() - t0
        t0 = time.perf_counter()
        mf, _, ms = m2f.pixel_decoder(self.sd_h, self.feats, "sem_seg_head.pixel_decoder.")
        out, _ = m2f.transformer_decoder(self.sd_h, ms, mf, "sem_seg_head.predictor.")
        te, ne = m2f.category_embed(self.sd_h, self.bank, self.null)
        lg = m2f.cal_pred_logits(out["mask_embed"], te, ne, out["logit_scale"], self.sizes)
        if self.full:                 # clip_head branch (odise.py:292-323): MaskCLIP over the whole image + ensemble
            oc = self.oclip
            me = oc.get_mask_embed(self.vis, self.img_full, out["pred_masks"])
            cl = oc.maskclip_pred_logits(me, self.bank, self.sizes, 100.0)
            oc.merge_with_void(lg, oc.pooling_clip_ensemble(lg[..., :-1], cl, self.ov, 0.3, 0.7))
        t_head = time.perf_counter() - t0
        ips = 1.0 / (self.crops * t_unet + t_head)
        return dict(value=ips, unit="images/s", cores=self.n, kind="port",
                    sample=f"1 crop (512^2) through {'CLIP ViT-L/14 image tower + VAE enc + UNet + full VAE dec' if self.full else 'the UNet'} as the "
                           f"reference executes it ({t_unet:.2f} s) + 1 image head{' + MaskCLIP' if self.full else ''} at {self.size}^2 "
                           f"({t_head:.2f} s); images/s = 1/({self.crops}*t_crop + t_head); fp32 torch CPU, "
                           f"{self.n} threads of {os.cpu_count()}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = CpuHotPath(args.size, args.vocab, full=args.full)
    t_start = time.perf_counter()
    for _ in range(min(args.warmup, 1)):
        cpu.sample()
    vals, info = [], None
    for _ in range(args.steps):
        info = cpu.sample()
        vals.append(info["value"])
        if time.perf_counter() - t_start > 200:                 # keep the arm within a few minutes
            break
    v = statistics.mean(vals)
    info["value"] = v
    ncls, npr = VOCABS[args.vocab]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ODISE hot path, {args.size}x{args.size}, {args.vocab} ({npr} prompts), CPU oracle",
                   "note": "reference arm = CPU restatement (reference not installable: detectron2/ldm/open_clip absent)"},
        "cpu_baseline": info,
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    args.full = not args.hot_path_only
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: odise_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from odise_b200 import lib, spec
    from odise_b200.pipeline import ODISEEngine, full_param_list, gather_logits, synthetic_vocabulary
    lib.load()
    nmma = 3 if args.precision == "bf16x3" else 1
    sd = spec.synth_state_dict(full_param_list(with_vae=args.full, with_clip=args.full), seed=0)
    eng = ODISEEngine(sd, dev, nmma=nmma, with_vae=args.full, with_clip=args.full)
    del sd
    ncls, npr = VOCABS[args.vocab]
    eng.set_vocabulary(args.vocab, *synthetic_vocabulary(ncls, npr))
    B, S = args.batch, args.size
    g = torch.Generator().manual_seed(1234 + rank)
    images = torch.randint(0, 256, (B, 3, S, S), generator=g, dtype=torch.uint8).pin_memory()

    graph, out = eng.capture(B, S, S)
    launches_per_step = eng.launches_per_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def dev_step():
        graph.replay()
        gather_logits(out["pred_logits"])

    def e2e_step():
        r = eng.infer(images)
        if world > 1:
            gather_logits(out["pred_logits"])
        return r

    for _ in range(args.warmup):
        dev_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(dev_step, args.steps)
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # roofline of the dominant kernel (gemm_tc_kernel): one eager pass with per-launch CUDA events
    lib.profile_begin()
    eng.step(B, S, S)
    n_gemm, gemm_ms, gemm_flops = lib.profile_end()
    torch.cuda.synchronize()
    # the north-star stage on its own: SD-v1 UNet feature pass (minimal pass, 0.740 TFLOP per 512^2 crop) on resident inputs
    from odise_b200.backbone import SyntheticTaps
    n_crops = B * ((S // 512) ** 2 if S > 512 else 1)
    tp_ = SyntheticTaps(dev)(n_crops)
    ctx_, cemb_ = eng.backbone.conditioning(tp_["clip_embed"], n_crops)
    lat_, lh_, lw_ = tp_["latent"]
    x_ = eng.backbone.q_sample(lat_, n_crops, lh_, lw_)
    for _ in range(2):
        eng.backbone.unet.forward(x_, n_crops, lh_, lw_, ctx_, cemb_)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(True), torch.cuda.Event(True)
    ev0.record()
    for _ in range(3):
        eng.backbone.unet.forward(x_, n_crops, lh_, lw_, ctx_, cemb_)
    ev1.record()
    torch.cuda.synchronize()
    unet_ms = ev0.elapsed_time(ev1) / 3
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    crops = (S // 512) ** 2 if S > 512 else 1
    ips = world * B * args.steps / (ms_dev / 1000.0)
    ips_e2e = world * B * args.steps / (ms_e2e / 1000.0)
    achieved = gemm_flops / (gemm_ms / 1000.0) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"ODISE(label) hot path, batch {B}/GPU x {S}x{S}, {crops} crops/image, {args.vocab} "
                               f"({ncls} classes / {npr} prompts), Q=100",
                   "stages": ("CLIP ViT-L/14-336 image tower on every crop, KL-VAE encoder + truncated decoder (taps), "
                              if args.full else "") + ("MaskCLIP (100 mask tokens/image through the ViT) + alpha/beta "
                              "ensemble + void merge, " if args.full else "") +
                             "implicit-captioner front, q_sample, SD-v1 UNet feature pass (4 taps), 8 projections, "
                             "MSDeformAttn pixel decoder, 9-layer masked-attention decoder, CLIP-text scoring, "
                             "NCCL all-gather of logits",
                   "not_in_path": ("nothing of the per-image pass: CLIP image tower, KL-VAE taps and MaskCLIP ARE executed; the CLIP "
                                   "TEXT bank of the vocabulary is precomputed per vocabulary (as in the reference)")
                                  if args.full else
                                  ("KL-VAE encoder/decoder taps and CLIP image embedding enter as seeded synthetic "
                                   "tensors (SURVEY.md §8f rows f-1/f-2)"),
                   "weights": "random-init (seed 0), SD-v1 / ODISE shapes", "global_batch": world * B,
                   "parallelism": f"dp{world} (image sharded)", "precision_mode": args.precision,
                   "l2": "working set >> 126 MB L2: ~3.6 GB of weight planes + multi-GB activations stream every step",
                   "cuda_graph": True},
        "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": int(images.numel()),
                "d2h_bytes_per_step": int(out["pred_logits"].numel() * 4 + out["pred_masks"].numel() * 4),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches_per_step * args.steps * 2),
        "gpu_launches_per_step": int(launches_per_step),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 GEMM / implicit conv, all launches of a step)",
                     "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_sustained"], "peak_source": pk["source"] + " sustained cuBLAS bf16",
                     "launches": int(n_gemm), "gemm_ms_per_step": gemm_ms, "algorithmic_tflop_per_step": gemm_flops / 1e12,
                     "mma_kind": "tcgen05.mma kind::f16 (bf16 in, fp32 TMEM accumulate)" +
                                 (", 3 MMAs per k-step (bf16x3 split)" if nmma == 3 else ""),
                     "tensor_pipe_equiv_frac": achieved * nmma / pk["bf16_sustained"],
                     "unet_frac_of_step": (UNET_TFLOP_PER_CROP * crops * B) / (gemm_flops / 1e12), "traffic": traffic,
                     "unet_feature_pass": {
                         "ms": unet_ms, "crops": n_crops, "algorithmic_tflop": UNET_TFLOP_PER_CROP * n_crops,
                         "achieved_tflops": UNET_TFLOP_PER_CROP * n_crops / (unet_ms / 1000.0),
                         "frac_of_bf16_peak": UNET_TFLOP_PER_CROP * n_crops / (unet_ms / 1000.0) / pk["bf16_sustained"],
                         "tensor_pipe_equiv_frac": nmma * UNET_TFLOP_PER_CROP * n_crops / (unet_ms / 1000.0) / pk["bf16_sustained"],
                         "note": "whole UNet pass incl. GroupNorm / attention softmax / elementwise kernels, eager launches, "
                                 "CUDA events; minimal pass FLOPs (output block 11 + out skipped)"}},
    }
    if not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = CpuHotPath(S, args.vocab, full=args.full).sample()
        except Exception as ex:  # noqa
            line["cpu_baseline"] = {"error": str(ex)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
