"""Numeric parity of the reference's OWN, UNMODIFIED model file running on top of the drop-in layer.

    python tests/ref_model_check.py pytorch|geometric [golden case ...]

``acm_gnn_b200.run.install`` registers the drop-in under the module name the reference imports
(``models.layers`` for ACM-Pytorch/models/models.py:7, ``layers`` for ACM-Geometric/models.py:3);
the reference's ``GCN`` class is then imported from baseline/_ref as it is, loaded with the
golden state_dict and compared -- output, loss, attention columns and every gradient -- with the
stored run of the unmodified reference (tests/golden/make_golden.py), in fp32 storage at the fp32
tolerance of tests/test_gpu_parity.py.  For the Geometric flavour the eval path of
``evaluate_acmgcn`` (data_utils.py:153-167: ``@torch.no_grad()``, ``model.eval()``) is checked too.

Runs as a child process of tests/test_gpu_reference_models.py because the two flavours claim the
same top-level module names (``models`` is a package in one tree and a module in the other).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from helpers import Golden, golden_cases  # noqa: E402

NORM, RTOL, ATOL = 2e-5, 2e-4, 2e-5


def close(got, ref, what):
    got = got.detach().float().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    ref = ref.detach().float().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = max(float(np.abs(ref).max()), 1e-12)
    err = float(np.abs(got - ref).max())
    assert err <= NORM * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=max(ATOL, NORM * scale), err_msg=what)
    return err / scale


def main():
    flavour = sys.argv[1]
    names = sys.argv[2:] or [c for c in golden_cases() if c.startswith("gcn_pt_" if flavour == "pytorch" else "gcn_geo_")]
    os.environ["ACMB200_DTYPE"] = "fp32"
    from acm_gnn_b200 import _lib, run
    ref_dir = os.path.join(ROOT, "baseline", "_ref", "ACM-Pytorch" if flavour == "pytorch" else "ACM-Geometric")
    drop_in = run.install(flavour, ref_dir)
    if flavour == "pytorch":
        from models.models import GCN           # the reference's file, unmodified
        import models.models as ref_models
    else:
        import models as ref_models             # ACM-Geometric/models.py
        GCN = ref_models.GCN
    assert os.path.realpath(ref_models.__file__).startswith(os.path.realpath(ref_dir)), ref_models.__file__
    assert ref_models.GraphConvolution is drop_in.GraphConvolution, "the reference model did not pick up the drop-in layer"
    for name in names:
        g = Golden(name)
        assert g.flavour == flavour
        model = GCN(nfeat=g.nfeat, nhid=g.nhid, nclass=g.nclass, nlayers=2, nnodes=g.n, dropout=0.0,
                    model_type=g.model_type, structure_info=g.structure_info, variant=bool(g.variant)).cuda()
        sd = {k[len("param/"):]: torch.from_numpy(g.z[k]) for k in g.z.files if k.startswith("param/")}
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        low, high, un = g.adjacency()
        low, high = low.cuda(), high.cuda()
        un = un.cuda() if un is not None else None
        x = g.x.clone().cuda().requires_grad_(True)
        n0 = _lib.launch_count()
        model.train()
        out = model(x, low, high, un)
        loss = torch.nn.functional.nll_loss(torch.log_softmax(out, 1)[g.idx_train.cuda()], g.labels.cuda()[g.idx_train.cuda()])
        loss.backward()
        torch.cuda.synchronize()
        assert _lib.launch_count() > n0, "no library launch: the drop-in layer did not run"
        e_out = close(out, g.z["out"], "out")
        close(loss, g.z["loss"], "loss")
        for li, layer in enumerate(model.gcns):
            cols = [layer.att_low, layer.att_high, layer.att_mlp]
            if g.structure_info and g.model_type != "acmgcn":
                cols.append(layer.att_struc_vec_low)
            close(torch.cat(cols, 1), g.z[f"att{li}"], f"att{li}")
        close(x.grad, g.z["grad_x"], "grad_x")
        ref_grads = g.grads()
        n_checked, worst = 0, 0.0
        for k, p in model.named_parameters():
            if k in ("fea_param", "xX_param") or ".bns." in k:
                continue
            rg = ref_grads[k]
            if rg.size == 0:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
                continue
            assert p.grad is not None, k
            worst = max(worst, close(p.grad, rg, "grad " + k))
            n_checked += 1
        assert n_checked >= 14, n_checked
        # eval path: model.eval() under torch.no_grad (ACM-Geometric/data_utils.py:153-159;
        # ACM-Pytorch/train.py:110-111 runs it with grad enabled -- both must reproduce the output)
        model.eval()
        with torch.no_grad():
            o_ng = model(x.detach(), low, high, un)
        o_g = model(x.detach(), low, high, un)
        close(o_ng, g.z["out"], "eval out (no_grad)")
        assert torch.equal(o_ng, o_g.detach()), "no_grad and grad-enabled eval forwards differ"
        print(f"ref_model_check[{flavour}] {name}: out rel.err {e_out:.1e}, worst grad rel.err {worst:.1e}, "
              f"{n_checked} gradients, eval(no_grad) OK", flush=True)
    print("REF_MODEL_CHECK PASS", flush=True)


if __name__ == "__main__":
    main()
