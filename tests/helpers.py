"""Shared helpers for the test-suite (golden loading, comparisons)."""
import os
from collections import OrderedDict

import numpy as np
import torch

from oracle import mfm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_params(g, prefix="p0/"):
    return OrderedDict((k[len(prefix):], torch.from_numpy(v.copy())) for k, v in g.items() if k.startswith(prefix))


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def tiny_case(head="l1", od=1):
    g = load_golden("tiny_%s_out%d.npz" % (head, od))
    seed, T, n, data_seed, noise_seed, od_ = [int(v) for v in g["meta"]]
    configs = O.tiny_configs(output_dim=od)
    P = golden_params(g)
    x = torch.from_numpy(g["x"].copy())
    y = torch.from_numpy(g["y"].copy())
    noise = O.draw_mmd_noise(configs, n, noise_seed)
    return g, configs, P, x, y, noise, T, n


def tiny_kl_case():
    g = load_golden("tiny_kl_l1_out1.npz")
    seed, T, n, data_seed, noise_seed, od_ = [int(v) for v in g["meta"]]
    configs = O.tiny_configs(output_dim=1)
    configs[0]["type"] = "kl"
    P = golden_params(g)
    x = torch.from_numpy(g["x"].copy())
    y = torch.from_numpy(g["y"].copy())
    return g, configs, P, x, y, T, n


def tiny_kl_ef_case():
    g = load_golden("tiny_kl_ef_l1_out1.npz")
    seed, T, n, data_seed, noise_seed, od_ = [int(v) for v in g["meta"]]
    configs = O.tiny_configs(output_dim=1)
    P = golden_params(g)
    x = torch.from_numpy(g["x"].copy())
    y = torch.from_numpy(g["y"].copy())
    return g, configs, P, x, y, T, n


def tiny_ablation_case(variant):
    """Golden vectors of the unmodified reference's M_A / M_B / M_C / M_D (oracle/make_golden.py, section 1d)."""
    g = load_golden("tiny_%s_l1_out1.npz" % variant)
    seed, T, n, data_seed, noise_seed, od_ = [int(v) for v in g["meta"]]
    configs = O.tiny_configs(output_dim=1)
    configs[0]["type"] = variant
    P = golden_params(g)
    x = torch.from_numpy(g["x"].copy())
    y = torch.from_numpy(g["y"].copy())
    noise = O.draw_mmd_noise(configs, n, noise_seed, variant=variant)
    noise = [torch.zeros(1, 1) if v is None else v for v in noise]
    return g, configs, P, x, y, noise, T, n


def tiny_missing_case():
    """Golden vectors of the unmodified reference's MFM_missing through train_mfm_missing's step (make_golden.py, section 1f)."""
    g = load_golden("tiny_missing_l1_out1.npz")
    seed, T, n, data_seed, noise_seed, od_ = [int(v) for v in g["meta"]]
    configs = O.tiny_configs(output_dim=1)
    P = golden_params(g)
    x = torch.from_numpy(g["x"].copy())
    y = torch.from_numpy(g["y"].copy())
    noise = O.draw_mmd_noise(configs, n, noise_seed)
    return g, configs, P, x, y, noise, T, n
