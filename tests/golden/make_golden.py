"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not travel):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so parity is
pinned on outputs of the reference modules themselves:

  * layer_*.npz / gcn_*.npz : reference ``GCN`` (ACM-Pytorch/models/models.py and
    ACM-Geometric/models.py, imported as they are) run forward + ``loss.backward()`` on
    seeded inputs; inputs, the full state_dict, outputs, attention columns and every
    gradient are stored.
  * operator_*.npz : the adjacency operators as the reference drivers build them
    (ACM-Pytorch/utils.py:421-438,626-628 via its own ``normalize_tensor``;
    ACM-Geometric/utils.py:5-28 via its own ``normalize_tensor`` /
    ``sparse_mx_to_torch_sparse_tensor``), stored as CSR/COO arrays for bit-exact checks.
  * dataset_cora.npz / dataset_squirrel.npz : edge lists of the two BASELINE fixture
    graphs loaded through the reference's ``utils.load_full_data`` and the CSR + values of
    the operators ``train_prep``'s recipe produces on them.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference(flavour):
    """Import the reference's models module for one flavour, with stubs for the modules it
    imports but does not use on this path."""
    for k in [k for k in sys.modules if k in ("models", "layers", "utils") or k.startswith("models.")]:
        del sys.modules[k]
    for p in (REF + "/ACM-Pytorch", REF + "/ACM-Geometric"):
        while p in sys.path:
            sys.path.remove(p)
    if flavour == "pytorch":
        _stub("google_drive_downloader", GoogleDriveDownloader=object)
        sys.path.insert(0, REF + "/ACM-Pytorch")
        models = importlib.import_module("models.models")
        return models
    _stub("dgl")
    _stub("dgl.function")
    _stub("dgl.utils")
    sys.modules["dgl"].function = sys.modules["dgl.function"]
    sys.modules["dgl"].utils = sys.modules["dgl.utils"]
    _stub("dgl.nn")
    _stub("dgl.nn.pytorch", GraphConv=object)
    _stub("torch_sparse", SparseTensor=object, matmul=None)
    sys.path.insert(0, REF + "/ACM-Geometric")
    models = importlib.import_module("models")
    return models


def make_graph(n, e, seed, self_loops=3, dups=5):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, e)
    dst = rng.integers(0, n, e)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = np.unique(np.concatenate([src * n + dst, dst * n + src]))
    row, col = key // n, key % n
    # data self-loops (quirk Q5) and an isolated node (row n-1 keeps only the added I)
    loops = rng.choice(n - 1, self_loops, replace=False)
    row = np.concatenate([row, loops])
    col = np.concatenate([col, loops])
    iso = (row != n - 1) & (col != n - 1)
    return row[iso].astype(np.int64), col[iso].astype(np.int64)


def reference_operator_pytorch(row, col, n):
    sys.path.insert(0, REF + "/ACM-Pytorch")
    import utils as ref_utils  # noqa  (the reference's own utils.py)
    a = torch.sparse_coo_tensor(torch.from_numpy(np.stack([row, col])), torch.ones(len(row)), (n, n))
    adj_low = ref_utils.normalize_tensor(torch.eye(n) + a.to_dense())
    adj_high = (torch.eye(n) - adj_low).to_sparse()
    return adj_low, adj_high, a.coalesce()


def reference_operator_geometric(row, col, n):
    import scipy.sparse as sp
    spec = importlib.util.spec_from_file_location("ref_geo_utils", REF + "/ACM-Geometric/utils.py")
    gu = importlib.util.module_from_spec(spec)
    if not hasattr(torch.sparse, "FloatTensor"):
        raise RuntimeError("torch.sparse.FloatTensor missing")
    spec.loader.exec_module(gu)
    a = sp.coo_matrix((np.ones(len(row)), (row, col)), shape=(n, n))  # to_scipy_sparse_matrix
    adj_low = gu.normalize_tensor(sp.identity(n) + a)
    adj_high = sp.identity(n) - adj_low
    lo = gu.sparse_mx_to_torch_sparse_tensor(adj_low).coalesce()
    hi = gu.sparse_mx_to_torch_sparse_tensor(adj_high).coalesce()
    au = gu.sparse_mx_to_torch_sparse_tensor(a).coalesce()
    return lo, hi, au


def csr_arrays(t):
    c = t.to_sparse().coalesce().to_sparse_csr() if t.layout == torch.strided else t.coalesce().to_sparse_csr()
    return c.crow_indices().numpy(), c.col_indices().numpy(), c.values().numpy()


def run_case(name, flavour, model_type, variant, structure_info, n, e, nfeat, nhid, nclass, seed,
             dropout=0.0):
    models = import_reference(flavour)
    torch.manual_seed(seed)
    np.random.seed(seed)
    row, col = make_graph(n, e, seed)
    if flavour == "pytorch":
        adj_low, adj_high, a_raw = reference_operator_pytorch(row, col, n)
    else:
        adj_low, adj_high, a_raw = reference_operator_geometric(row, col, n)
    adj_un = a_raw if structure_info else None
    x = torch.rand(n, nfeat)
    x = x / x.sum(1, keepdim=True)
    x.requires_grad_(True)
    labels = torch.randint(0, nclass, (n,))
    idx_train = torch.randperm(n)[: int(0.6 * n)]
    kw = {}
    model = models.GCN(nfeat, nhid, nclass, 2, n, dropout, model_type, structure_info, variant, **kw)
    model.train()
    out = model(x, adj_low, adj_high, adj_un)
    loss = torch.nn.functional.nll_loss(torch.log_softmax(out, 1)[idx_train], labels[idx_train])
    loss.backward()
    rec = {
        "meta": np.array([n, nfeat, nhid, nclass, int(bool(variant)), int(structure_info), seed]),
        "flavour": np.array(flavour), "model_type": np.array(model_type),
        "row": row, "col": col, "x": x.detach().numpy(), "labels": labels.numpy(),
        "idx_train": idx_train.numpy(), "out": out.detach().numpy(), "loss": loss.detach().numpy(),
        "grad_x": x.grad.numpy(),
    }
    for li, g in enumerate(model.gcns):
        rec[f"att{li}"] = torch.cat([g.att_low, g.att_high, g.att_mlp] +
                                    ([g.att_struc_vec_low] if (structure_info and model_type != "acmgcn") else []), 1).detach().numpy()
    for k, v in model.state_dict().items():
        if k in ("fea_param", "xX_param"):
            continue  # uninitialised, unused (quirk Q4)
        if "num_batches_tracked" in k or ".bns." in k:
            continue
        rec["param/" + k] = v.detach().numpy()
    for k, v in model.named_parameters():
        if k in ("fea_param", "xX_param") or ".bns." in k:
            continue
        rec["grad/" + k] = (v.grad if v.grad is not None else torch.zeros(0)).detach().numpy()
    lo = csr_arrays(adj_low)
    hi = adj_high.coalesce()
    rec.update({"low_crow": lo[0], "low_col": lo[1], "low_val": lo[2],
                "high_idx": hi.indices().numpy(), "high_val": hi.values().numpy()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "loss", float(loss), "out", tuple(out.shape))


def dataset_fixture(name):
    """Edge list + operator arrays of a BASELINE fixture graph through the reference's own
    loader (utils.load_full_data) and operator recipe (train_prep, utils.py:626-628)."""
    import_reference("pytorch")
    cwd = os.getcwd()
    os.chdir(REF + "/ACM-Pytorch")  # the loader uses ../data, ../new_data relative paths
    try:
        import utils as ref_utils
        a, feats, labels = ref_utils.load_full_data(name)
        n = labels.shape[0]
        a = a.coalesce()
        adj_low = ref_utils.normalize_tensor(torch.eye(n) + a.to_dense())
        adj_high = (torch.eye(n) - adj_low).to_sparse().coalesce()
    finally:
        os.chdir(cwd)
    crow, ccol, cval = csr_arrays(adj_low)
    idx = a.indices().numpy()
    np.savez_compressed(
        os.path.join(HERE, f"dataset_{name}.npz"),
        n=np.array(n), row=idx[0].astype(np.int32), col=idx[1].astype(np.int32), raw_val=a.values().numpy(),
        low_crow=crow.astype(np.int32), low_col=ccol.astype(np.int32), low_val=cval,
        high_idx=adj_high.indices().numpy().astype(np.int32), high_val=adj_high.values().numpy(),
        nfeat=np.array(feats.shape[1]), nclass=np.array(int(labels.max()) + 1),
    )
    print("dataset", name, "n", n, "nnz(A)", idx.shape[1], "nnz(low)", len(ccol))


CASES = [
    # name, flavour, model_type, variant, structure_info, n, e, nfeat, nhid, nclass, seed
    ("gcn_pt_acmgcn_v0", "pytorch", "acmgcn", 0, 0, 257, 1500, 19, 24, 5, 1),
    ("gcn_pt_acmgcn_v1", "pytorch", "acmgcn", 1, 0, 257, 1500, 19, 24, 5, 2),
    ("gcn_pt_acmgcnp_v0_s1", "pytorch", "acmgcnp", 0, 1, 203, 1200, 33, 64, 7, 3),
    ("gcn_pt_acmgcnpp_v1_s1", "pytorch", "acmgcnpp", 1, 1, 203, 1200, 33, 16, 3, 4),
    ("gcn_pt_acmgcnpp_v0", "pytorch", "acmgcnpp", 0, 0, 150, 900, 12, 32, 2, 5),
    ("gcn_geo_acmgcn_v1", "geometric", "acmgcn", 1, 0, 257, 1500, 19, 24, 5, 6),
    ("gcn_geo_acmgcnp_v1", "geometric", "acmgcnp", 1, 0, 211, 1300, 7, 64, 2, 7),
    ("gcn_geo_acmgcnp_v0_s1", "geometric", "acmgcnp", 0, 1, 211, 1300, 21, 40, 5, 8),
    ("gcn_geo_acmgcnpp_v1_s1", "geometric", "acmgcnpp", 1, 1, 190, 1100, 128, 256, 5, 9),
]


if __name__ == "__main__":
    for c in CASES:
        run_case(*c)
    for ds in ("cora", "squirrel"):
        dataset_fixture(ds)
