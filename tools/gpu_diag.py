"""First-light diagnostics on a real B200: runs the loss kernels on small problems and prints where
(stats / C tile / row and column sums / gradients) the numbers diverge from the float64 closed form."""
import ctypes as C
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import abt_oracle as O
from ssl_audio_b200 import _lib
from ssl_audio_b200.loss import _WS, bt_loss_fwd_bwd


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def run(n, d, lam=0.005):
    z1, z2 = O.synth_embeddings(n, d, seed=n + d)
    t1 = torch.from_numpy(z1).cuda().bfloat16()
    t2 = torch.from_numpy(z2).cuda().bfloat16()
    loss, dz1, dz2 = bt_loss_fwd_bwd(t1, t2, 1.0, lam, False)
    torch.cuda.synchronize()
    rl, r1, r2, c = O.bt_loss_forward_backward(z1, z2, 1.0, lam)
    lib = _lib.load()
    offs = (C.c_size_t * 8)()
    lib.abt_debug_ws_offsets(n, d, 0, offs)
    buf, ptr = _WS.get(t1.device, n, d, 0)
    base = ptr - buf.data_ptr()
    raw = buf.cpu().numpy()
    stats = raw[base + offs[0]: base + offs[0] + 10 * d * 4].view(np.float32).reshape(10, d)
    Cm = torch.from_numpy(raw[base + offs[1]: base + offs[1] + 2 * d * d].copy()).view(torch.float16).float().numpy().reshape(d, d)
    accs = raw[base + offs[2]: base + offs[2] + 4 * 4 * d].view(np.float32).reshape(4, d)
    h1, mu1, _, rr1 = O.batchnorm_train(z1.astype(np.float64))
    h2, mu2, _, rr2 = O.batchnorm_train(z2.astype(np.float64))
    print(f"--- N={n} D={d}: loss {float(loss):.6f} ref {rl:.6f}  rel {abs(float(loss)-rl)/rl:.2e}")
    print("    stats mu1", rel(stats[0], mu1), "r1", rel(stats[1], rr1), "mu2", rel(stats[2], mu2), "cdiag", rel(stats[4], np.diagonal(c)))
    Cref = c.copy()
    np.fill_diagonal(Cref, 0.0)
    print("    C rel", rel(Cm, Cref), " C^T rel (transposition bug?)", rel(Cm.T, Cref))
    sq = Cref ** 2
    print("    row sq rel", rel(accs[0], sq.sum(1)), " col sq rel", rel(accs[1], sq.sum(0)))
    print("    dz1 rel", rel(dz1.float().cpu().numpy(), r1), " dz2 rel", rel(dz2.float().cpu().numpy(), r2))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    for n, d in [(32, 64), (128, 256), (64, 512), (300, 512)]:
        try:
            run(n, d)
        except Exception as e:  # keep going: later shapes may still tell us something
            print("FAILED", n, d, repr(e))
            break
