"""Shared test helpers: golden-fixture loading and oracle drivers (CPU)."""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import acm_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "gcn_*.npz")))


class Golden:
    """One stored run of the reference ``GCN`` (see tests/golden/make_golden.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.z = z
        m = z["meta"]
        self.n, self.nfeat, self.nhid, self.nclass = int(m[0]), int(m[1]), int(m[2]), int(m[3])
        self.variant, self.structure_info, self.seed = int(m[4]), int(m[5]), int(m[6])
        self.flavour = str(z["flavour"])
        self.model_type = str(z["model_type"])
        self.row, self.col = z["row"], z["col"]
        self.x = torch.from_numpy(z["x"])
        self.labels = torch.from_numpy(z["labels"])
        self.idx_train = torch.from_numpy(z["idx_train"])

    def params(self, requires_grad=True):
        p = {}
        for k in self.z.files:
            if not k.startswith("param/"):
                continue
            name = k[len("param/"):]
            if name.startswith("gcns."):
                grp, sub = name[:6], name[7:]
            elif name.startswith("mlpX.lins.0."):
                grp, sub = "mlpX", name[len("mlpX.lins.0."):]
            else:
                continue
            t = torch.from_numpy(self.z[k]).clone()
            t.requires_grad_(requires_grad)
            p.setdefault(grp, {})[sub] = t
        return p

    def grads(self):
        g = {}
        for k in self.z.files:
            if k.startswith("grad/"):
                g[k[len("grad/"):]] = self.z[k]
        return g

    def operator(self):
        return O.build_operator(self.row, self.col, self.n, self.flavour)

    def adjacency(self):
        """adj_low / adj_high / adj_low_unnormalized exactly as the reference driver of this
        flavour passes them (dense adj_low for ACM-Pytorch, COO for Geometric)."""
        op = self.operator()
        low, high = O.operator_to_torch(op, dense_low=(self.flavour == "pytorch"))
        un = O.raw_adjacency_to_torch(self.row, self.col, self.n) if self.structure_info else None
        return low, high, un


def oracle_run(g: Golden):
    p = g.params()
    x = g.x.clone().requires_grad_(True)
    low, high, un = g.adjacency()
    out, atts = O.gcn_forward(p, x, low, high, un, model_type=g.model_type, variant=bool(g.variant),
                              structure_info=g.structure_info, flavour=g.flavour)
    loss = O.train_step_loss(out, g.labels, g.idx_train)
    loss.backward()
    return out, atts, loss, x.grad, p


def flat_param_name(grp, sub):
    return ("mlpX.lins.0." + sub) if grp == "mlpX" else (grp + "." + sub)
