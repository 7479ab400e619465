"""Helpers shared by the GPU parity tests."""

import numpy as np
import torch


def relmax(a, b):
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def cpu_sd(head, dtype=torch.float32):
    return {k: v.detach().to("cpu", dtype) for k, v in head.state_dict().items()}


def make_case(depth, D, Dc, B, H, W, patch=1, chan=3, seed=7, n_pred=None, device="cuda"):
    import nova_pointcloud_b200 as nb

    torch.manual_seed(seed)
    head = nb.DiffusionMLP(depth, D, Dc, patch_size=patch, image_dim=chan).eval()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, chan, H * patch, W * patch, generator=g)
    N = H * W
    z = torch.randn(B, N, Dc, generator=g)
    t = torch.rand(B, generator=g) * 1000
    order = torch.rand(B, N, generator=g).argsort(dim=1)
    pred_ids = None if n_pred is None else order[:, :n_pred].unsqueeze(-1).contiguous()
    return head, x, z, t, pred_ids


def record(name, value):
    """Append a measured parity error to gpurun_out/parity_errors.jsonl (copied into profiles/ per round)."""
    import json
    import os

    root = os.environ.get("GRAFT_REPO_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, "rel_max_err": float(value)}) + "\n")
    except OSError:
        pass
