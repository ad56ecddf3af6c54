"""Launcher: run an UNMODIFIED reference driver on top of the drop-in ACM layer.

    python -m acm_gnn_b200.run /path/to/ACM-Pytorch/train.py   --dataset_name cora --model acmgcn ...
    python -m acm_gnn_b200.run /path/to/ACM-Geometric/train.py --dataset twitch-gamer --method acmgcnp ...

The reference has no plugin registry; its "plugin API" is the import
``from models.layers import GraphConvolution, MLP`` (ACM-Pytorch/models/models.py:7) /
``from layers import GraphConvolution, MLP`` (ACM-Geometric/models.py:3).  This launcher
pre-seeds ``sys.modules`` with our module under that name, stubs the third-party modules the
drivers import but do not need on this path, chdir's into the driver's
directory (it uses relative paths: ../data, splits/, ./logs) and runpy's the script as
``__main__``.  The reference files are executed as they are.

Knobs (environment, so the reference CLI stays unchanged):
    ACMB200_DTYPE = fp32 | bf16     storage of feature tables (default fp32 = the reference precision; bf16 opts in)
    ACMB200_GEMM  = auto | simt | tcgen05
"""
import importlib
import os
import runpy
import sys
import types


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], child, m)
        return m


def install(flavour: str, ref_dir: str):
    """Register the drop-in under the module name the reference imports."""
    if flavour == "pytorch":
        from . import layers as drop_in
        _stub("google_drive_downloader", GoogleDriveDownloader=object)
        _stub("torch_geometric")
        _stub("torch_geometric.utils", add_self_loops=None, to_undirected=None)
        _stub("torch_sparse", SparseTensor=object)
        # `models` stays the reference's package; only its `layers` submodule is ours
        sys.path.insert(0, ref_dir)
        pkg = importlib.import_module("models")
        sys.modules["models.layers"] = drop_in
        pkg.layers = drop_in
    else:
        from . import layers_geometric as drop_in
        for name in ("dgl", "dgl.function", "dgl.utils", "dgl.nn", "dgl.nn.pytorch"):
            _stub(name, GraphConv=object)
        _stub("torch_sparse", SparseTensor=object, matmul=None)
        sys.path.insert(0, ref_dir)
        sys.modules["layers"] = drop_in
    return drop_in


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    ref_dir = os.path.dirname(script)
    flavour = "geometric" if "Geometric" in ref_dir else "pytorch"
    flavour = os.environ.get("ACMB200_FLAVOUR", flavour)
    install(flavour, ref_dir)
    os.chdir(ref_dir)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
