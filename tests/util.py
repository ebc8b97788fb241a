"""Shared helpers for the test-suite (fixtures, error metric)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE = 2048


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    spec = json.loads(str(z["spec"]))
    arrs = {k: z[k] for k in z.files if k != "spec"}
    return spec, arrs


def rel_err(a, b) -> float:
    """max|a-b| / max|b| per tensor (the metric SURVEY.md section 8c fixes)."""
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().cpu()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    den = b.abs().max().item()
    num = (a - b).abs().max().item()
    if den == 0.0:
        return num
    return num / den


def sample(t: torch.Tensor) -> np.ndarray:
    """Same strided sample as tests/golden/make_golden.py."""
    f = t.detach().reshape(-1).cpu()
    if f.numel() <= SAMPLE:
        return f.numpy().copy()
    step = f.numel() // SAMPLE
    return f[::step][:SAMPLE].numpy().copy()
