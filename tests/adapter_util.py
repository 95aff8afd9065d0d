"""Shared helpers of the adapter parity tests: load a golden fixture (tests/golden/make_adapter_golden.py) and run the
fp64 oracle (oracle/adapter_oracle.py) on its inputs with the loss the fixture's gradients belong to."""
import os

import numpy as np
import torch

from oracle import adapter_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["adapter_pf3plat", "adapter_deg2_batch", "adapter_improper"]
INPUTS = ("extrinsics", "intrinsics", "coordinates", "depths", "opacities", "raw_gaussians")
OUTPUTS = ("means", "covariances", "harmonics", "scales", "rotations")
GRADS = ("extrinsics", "intrinsics", "coordinates", "depths", "raw_gaussians")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    b, v, h, w, deg, proper = (int(x) for x in z["meta"])
    return z, dict(b=b, v=v, h=h, w=w, sh_degree=deg, proper=bool(proper))


def oracle_run(z, meta, dtype=torch.float64):
    """Oracle outputs and gradients (of sum_k <w_k, out_k>) on the fixture's inputs."""
    leaves = {k: torch.from_numpy(z["in_" + k]).to(dtype).requires_grad_(k != "opacities") for k in INPUTS}
    d = torch.from_numpy(z["sh_rotation"]).to(dtype)[:, :, None, None, None]  # (b, v, 1, 1, 1, d_sh, d_sh)
    out = adapter_oracle.adapter_forward(leaves["extrinsics"], leaves["intrinsics"], leaves["coordinates"], leaves["depths"],
                                         leaves["opacities"], leaves["raw_gaussians"], (meta["h"], meta["w"]),
                                         meta["sh_degree"], 0.5, 15.0, sh_rotation=d)
    loss = sum((torch.from_numpy(z["w_" + k]).to(dtype) * out[k]).sum() for k in OUTPUTS)
    loss.backward()
    return {k: out[k].detach() for k in OUTPUTS}, {k: leaves[k].grad for k in GRADS}


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
