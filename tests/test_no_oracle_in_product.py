"""The oracle is test infrastructure: nothing under mspl_b200/ (the product) may import, call or read oracle/, tests/ or the
reference tree, no CPU fallback may hide a missing CUDA library, and of the repo-root entry points only smoke() and bench.py's
CPU-baseline / reference-arm legs may touch the oracle."""
import ast
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT = os.path.join(ROOT, "mspl_b200")


def _py_files(top):
    for d, _, files in os.walk(top):
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(d, f)


def _imported_modules(path):
    tree = ast.parse(open(path).read(), filename=path)
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name
        elif isinstance(node, ast.ImportFrom) and node.module and node.level == 0:
            yield node.module


def test_product_never_imports_the_oracle_or_the_tests():
    offenders = []
    for path in _py_files(PRODUCT):
        for mod in _imported_modules(path):
            top = mod.split(".")[0]
            if top in ("oracle", "tests", "oracle_ops", "bench", "__graft_entry__"):
                offenders.append((os.path.relpath(path, ROOT), mod))
    assert not offenders, offenders


def test_product_sources_do_not_mention_the_oracle_module_or_read_the_reference_tree():
    # the reference tree may be IMPORTED by one documented default (the generators' data loader, a reference class the caller
    # can replace with testloader=); nothing else may reach into it, and no product file may name the oracle package
    allowed_reference_mentions = {os.path.join("mspl_b200", "uest_seg_multi_os.py")}
    for path in list(_py_files(PRODUCT)) + [os.path.join(PRODUCT, "csrc", f) for f in os.listdir(os.path.join(PRODUCT, "csrc"))
                                            if f.endswith((".cu", ".cuh"))]:
        text = open(path).read()
        rel = os.path.relpath(path, ROOT)
        assert "mspl_oracle" not in text and "from oracle" not in text and "import oracle" not in text, rel
        if rel not in allowed_reference_mentions:
            assert "/root/reference" not in text, rel


def test_missing_library_fails_loudly_instead_of_falling_back(tmp_path):
    """With the shared library absent, loading it raises (there is no CPU path to fall back to)."""
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['MSPL_B200_LIB'] = %r\n"
            "from mspl_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.MsplError as e:\n    print('RAISED', 'no CPU fallback' in str(e))\nelse:\n    print('LOADED')\n"
            % (ROOT, str(tmp_path / "libmspl_b200_missing.so")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.stdout.strip() == "RAISED True", out.stdout + out.stderr


def test_ops_reject_cpu_tensors():
    import torch
    from mspl_b200 import ops
    from mspl_b200.loss_fns.segmentation_loss import PixelwiseKLD
    x = torch.zeros(1, 5, 4, 4)
    with pytest.raises(ValueError, match="CUDA"):
        ops.fuse_sources([x], [x], [[1, 2, 3, 1, 2]])
    with pytest.raises(ValueError, match="CUDA"):
        ops.uw_ce_fwd_bwd(x, x, torch.zeros(1, 4, 4, dtype=torch.int64), torch.ones(5))
    with pytest.raises(ValueError, match="CUDA"):
        PixelwiseKLD()(x, x)


def test_only_the_checker_legs_of_the_entry_points_use_the_oracle():
    """bench.py: the oracle appears only inside the CpuPath class (the CPU arm), which in turn is only instantiated by
    cpu_baseline() / run_reference(); __graft_entry__.py: only in build() (checker build) and smoke()."""
    for fname, allowed in (("bench.py", {"CpuPath"}), ("__graft_entry__.py", {"build", "smoke"})):
        tree = ast.parse(open(os.path.join(ROOT, fname)).read())
        for node in tree.body:
            uses = [n for n in ast.walk(node) if isinstance(n, ast.ImportFrom) and n.module and n.module.split(".")[0] == "oracle"]
            uses += [n for n in ast.walk(node) if isinstance(n, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in n.names)]
            if uses:
                assert isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in allowed, (fname, getattr(node, "name", type(node).__name__))
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and any(isinstance(n, ast.Name) and n.id == "CpuPath" for n in ast.walk(node)):
            assert node.name in {"cpu_baseline", "run_reference"}, node.name
