"""Loaders for tests/golden/ssl_golden.npz (outputs of the reference's own mean-teacher source)."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
PL_CASES = ["pl0", "pl1", "pl2"]


def load_golden():
    return np.load(os.path.join(HERE, "golden", "ssl_golden.npz"))


def pl_inputs(G, tag, device):
    preds = {k[len(tag) + 4:]: torch.from_numpy(G[k]).to(device).clone()
             for k in G.files if k.startswith(f"{tag}_in_")}
    B, P, warm, n_ulb, n_lb = (int(v) for v in G[f"{tag}_cfg"])
    return (preds, torch.from_numpy(G[f"{tag}_ulb_list"]).to(device),
            torch.from_numpy(G[f"{tag}_ulb_flag"]).to(device), n_lb, n_ulb, bool(warm))


def pl_expected(G, tag):
    counts = [int(c) for c in G[f"{tag}_counts"]]
    off = np.concatenate([[0], np.cumsum(counts)])
    sl = lambda a: [torch.from_numpy(a[off[i]:off[i + 1]]) for i in range(len(counts))]   # noqa: E731
    return counts, sl(G[f"{tag}_labels"]), sl(G[f"{tag}_boxes"]), sl(G[f"{tag}_quality"])


def aug_from_golden(G, side, device):
    from nesie_b200.detectors import BoxAug
    t = lambda k: torch.from_numpy(G[f"tf_{side}_{k}"]).to(device)   # noqa: E731
    return BoxAug(t("hf"), t("vf"), t("rot"), t("scale"), t("trans"))


def padded_boxes(G, device):
    counts = [int(c) for c in G["tf_counts"]]
    flat = torch.from_numpy(G["tf_boxes"])
    out = torch.zeros(len(counts), max(counts), 7)
    o = 0
    for i, n in enumerate(counts):
        out[i, :n] = flat[o:o + n]
        o += n
    return out.to(device), counts
