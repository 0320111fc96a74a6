"""Pin the torch CPU baseline port (oracle/torch_port.py, what `bench.py --impl reference` times) against the
golden fixtures recorded from the reference."""
import os
import random

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from oracle import torch_port as P  # noqa: E402


def test_port_views_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "views.npz"))
    seed = int(g["seed"])
    np.random.seed(seed); random.seed(seed)
    tfm = P.PortPairTransform()
    for b in range(g["x"].shape[0]):
        v = torch.stack(tfm(torch.from_numpy(g["x"][b]))).numpy()
        np.testing.assert_allclose(v, g["views"][b], rtol=0, atol=1e-5)


def test_port_lms_path(golden_dir):
    g = np.load(os.path.join(golden_dir, "views.npz"))
    np.random.seed(77); random.seed(77)
    tfm = P.PortPairTransform()
    for b in range(3):
        v = torch.stack(P.clip_lms_path(torch.from_numpy(g["lms_full"][b]), (-0.8294, 4.6230), tfm)).numpy()
        np.testing.assert_allclose(v, g["lms_views"][b], rtol=0, atol=1e-5)


def test_port_logmel(golden_dir):
    pytest.importorskip("torchaudio")
    g = np.load(os.path.join(golden_dir, "logmel.npz"))
    out = P.log_mel(P.make_melspec(), torch.from_numpy(g["wav"])).numpy()
    np.testing.assert_allclose(out, g["lms"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_port_loss(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "loss.npz"))
    z1 = torch.from_numpy(g[f"{tag}_z1"]).requires_grad_(True)
    z2 = torch.from_numpy(g[f"{tag}_z2"]).requires_grad_(True)
    crit = P.PortBarlowTwinsLoss(z1.shape[1], hsic=bool(g[f"{tag}_hsic"]))
    loss = crit(z1, z2)
    loss.backward()
    assert abs(float(loss) - float(g[f"{tag}_loss"])) <= 1e-6 * abs(float(g[f"{tag}_loss"]))
    np.testing.assert_allclose(z1.grad.numpy(), g[f"{tag}_dz1"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(z2.grad.numpy(), g[f"{tag}_dz2"], rtol=1e-4, atol=1e-7)
