"""Context for the hot-path numbers: a WHOLE training step -- this library's frontend, a BYOL-A-sized convolutional
encoder and the reference-shaped projector (in -> 8192 -> 8192 -> 8192, bias-free Linear + BatchNorm + ReLU, model.py:11-31) in plain
PyTorch under bf16 autocast, this library's objective and LARS -- next to the same step with the objective written the way the
reference writes it (utils/loss.py:15-30) in PyTorch on the GPU.  The encoder is written for this measurement (it is not the
reference's model code and carries random weights); the point is only how large the hot path is inside a real step.

    python tools/context_step.py [per_gpu_batch ...]        (one GPU; prints one JSON line per batch size)
"""
import json
import os
import sys
import types

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssl_audio_b200 as S
from bench import AS_STATS, _args_ns


class Encoder(nn.Module):
    def __init__(self, d=2048):
        super().__init__()
        blocks, c_in = [], 1
        for _ in range(3):
            blocks += [nn.Conv2d(c_in, 64, 3, padding=1), nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.MaxPool2d(2)]
            c_in = 64
        self.features = nn.Sequential(*blocks)                    # (B, 1, 64, 96) -> (B, 64, 8, 12)
        self.fc = nn.Sequential(nn.Linear(64 * 8, d), nn.ReLU(inplace=True), nn.Linear(d, d), nn.ReLU(inplace=True))

    def forward(self, x):
        x = self.features(x)
        b, c, f, t = x.shape
        x = self.fc(x.permute(0, 3, 1, 2).reshape(b, t, c * f))
        return x.max(1).values + x.mean(1)


def projector(d_in, d):
    return nn.Sequential(nn.Linear(d_in, d, bias=False), nn.BatchNorm1d(d), nn.ReLU(inplace=True),
                         nn.Linear(d, d, bias=False), nn.BatchNorm1d(d), nn.ReLU(inplace=True), nn.Linear(d, d, bias=False))


def reference_style_loss(z1, z2, bn, lmbda):
    n, d = z1.shape
    c = bn(z1).T @ bn(z2) / n
    on = torch.diagonal(c).add(-1).pow(2).sum()
    off = c.flatten()[:-1].view(d - 1, d + 1)[:, 1:].pow(2).sum()
    return on + lmbda * off


def timed(fn, warm=4, it=12):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    D = 8192
    cfg = _args_ns(D)
    for B in [int(v) for v in sys.argv[1:]] or [128, 1024]:
        torch.manual_seed(0)
        wav = 0.1 * torch.randn(B, 160000, device=dev)
        fe = S.BatchFrontend(cfg, norm_stats=AS_STATS, path="lms", mode="crop")
        enc, proj = Encoder().to(dev), projector(2048, D).to(dev)
        params = list(enc.parameters()) + list(proj.parameters())
        crit = S.BarlowTwinsLoss(cfg, ncrops=2).to(dev)
        ref_bn = nn.BatchNorm1d(D, affine=False).to(dev)
        opt = S.LARS(params, lr=0.2, weight_decay=1e-6, weight_decay_filter=True, lars_adaptation_filter=True)

        def step(ours):
            views = fe(wav)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                z1 = proj(enc(views[0]))
                z2 = proj(enc(views[1]))
                loss = crit.forward_loss(z1, z2) if ours else reference_style_loss(z1, z2, ref_bn, cfg.lmbda)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss

        def model_only():
            views = [v.detach() for v in cur_views]
            with torch.autocast("cuda", dtype=torch.bfloat16):
                z1 = proj(enc(views[0]))
                z2 = proj(enc(views[1]))
            (z1.float().mean() + z2.float().mean()).backward()
            opt.zero_grad(set_to_none=True)

        cur_views = fe(wav)
        z1s = torch.randn(B, D, device=dev).bfloat16()
        z2s = (0.6 * z1s.float() + 0.8 * torch.randn(B, D, device=dev)).bfloat16()

        def loss_ours():
            a = z1s.detach().requires_grad_(True); b = z2s.detach().requires_grad_(True)
            crit.forward_loss(a, b).backward()

        def loss_ref():
            a = z1s.detach().requires_grad_(True); b = z2s.detach().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                lo = reference_style_loss(a, b, ref_bn, cfg.lmbda)
            lo.backward()

        step(True)                                   # gradients exist from here on (LARS needs them)
        out = {"per_gpu_batch": B, "projector_out_dim": D,
               "step_ms_with_this_objective": timed(lambda: step(True)),
               "step_ms_with_reference_style_objective_in_pytorch": timed(lambda: step(False)),
               "frontend_ms": timed(lambda: fe(wav)),
               "encoder_projector_fwd_bwd_ms": timed(model_only),
               "objective_fwd_bwd_ms": {"this_library": timed(loss_ours), "reference_style_pytorch": timed(loss_ref)}}
        for p_ in params:                            # LARS alone: it skips parameters without a gradient
            p_.grad = 1e-3 * torch.randn_like(p_)
        out["lars_step_ms"] = timed(lambda: opt.step())
        opt.zero_grad(set_to_none=True)
        out["loss_values"] = [float(step(True).detach()), float(step(False).detach())]
        out["n_parameters"] = sum(p.numel() for p in params)
        print(json.dumps(out), flush=True)
        del fe, enc, proj, crit, opt, wav
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
