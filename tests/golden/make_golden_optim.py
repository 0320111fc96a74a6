"""Golden fixtures for LARS.step and update_moving_average FROM THE LIVE REFERENCE (build container only; needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_optim.py

Runs the reference's own `utils.utils.LARS` (utils/utils.py:150-189) for two steps and `update_moving_average`
(utils/utils.py:328-331) twice on a small parameter set (one tensor larger than a kernel chunk, one all-zero parameter and one
all-zero gradient for the trust-ratio guards) and records parameters, gradients and results.
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import utils.utils as U  # noqa: E402  (reference)

torch.set_num_threads(1)
SHAPES = [(16, 32), (64,), (3, 3, 3, 8), (8,), (9000,), (17, 5), (33,)]
out = {"n_tensors": np.int32(len(SHAPES))}
for tag, kw in (("a", dict(lr=0.2, weight_decay=1e-4, momentum=0.9, eta=0.001, weight_decay_filter=True, lars_adaptation_filter=True)),
                ("b", dict(lr=0.05, weight_decay=1.5e-6, momentum=0.8, eta=0.02, weight_decay_filter=False, lars_adaptation_filter=False))):
    g = torch.Generator().manual_seed(5 if tag == "a" else 6)
    params = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.3) for s in SHAPES]
    with torch.no_grad():
        params[5].zero_()                                   # |p| = 0 -> q = 1
    opt = U.LARS(params, **kw)
    for k, v in kw.items():
        out[f"{tag}_{k}"] = np.float64(v)
    for t, p in enumerate(params):
        out[f"{tag}_p0_{t}"] = p.detach().numpy().copy()
    for step in range(2):
        for t, p in enumerate(params):
            gr = torch.randn(p.shape, generator=g) * 0.1
            if t == 6 and step == 1:
                gr.zero_()                                  # |dp| = 0 without weight decay on 1-D tensors -> q = 1
            p.grad = gr
            out[f"{tag}_g{step}_{t}"] = gr.numpy().copy()
        opt.step()
        for t, p in enumerate(params):
            out[f"{tag}_p{step + 1}_{t}"] = p.detach().numpy().copy()
            out[f"{tag}_mu{step + 1}_{t}"] = opt.state[p]["mu"].numpy().copy()

# EMA (BYOL target network): two updates with the online parameters changing in between
g = torch.Generator().manual_seed(9)


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s, generator=g)) for s in SHAPES])


online, target = Net(), Net()
ema = U.EMA(0.99)
for t in range(len(SHAPES)):
    out[f"ema_target0_{t}"] = target.ps[t].detach().numpy().copy()
for step in range(2):
    with torch.no_grad():
        for t, p in enumerate(online.ps):
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
            out[f"ema_online{step}_{t}"] = p.numpy().copy()
    U.update_moving_average(ema, target, online)
    for t in range(len(SHAPES)):
        out[f"ema_target{step + 1}_{t}"] = target.ps[t].detach().numpy().copy()
out["ema_beta"] = np.float64(0.99)
np.savez_compressed(os.path.join(HERE, "optim.npz"), **out)
print("wrote optim.npz with", len(out), "arrays")
