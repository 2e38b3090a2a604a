"""The C-ABI library loads on a CPU-only host and exports every symbol the headers declare."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        syms |= set(re.findall(r"\b(sdb_[a-z0-9_]+)\s*\(", src))
    return sorted(syms)


def test_headers_declare_something():
    assert len(declared_symbols()) >= 10


def test_library_exports_every_declared_symbol():
    from scaledreamer_b200 import lib as L

    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    dll = ctypes.CDLL(L.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(dll, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_ctypes_signatures_cover_the_header():
    from scaledreamer_b200 import lib as L

    L.load()
    bound = set(L.SIGNATURES) | {"sdb_last_error", "sdb_abi_version", "sdb_launch_count"}
    assert set(declared_symbols()) <= bound, sorted(set(declared_symbols()) - bound)


def test_host_only_helpers_work_without_gpu():
    from scaledreamer_b200 import lib as L

    assert L.load().sdb_abi_version() >= 1
    cfg = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16,
               per_level_scale=1.447269237440378)
    assert L.grid_num_entries(cfg) == 6299960
    with pytest.raises(RuntimeError, match="n_features_per_level"):
        L.grid_num_entries({**cfg, "n_features_per_level": 4})


def test_product_path_refuses_cpu_tensors():
    import torch

    from scaledreamer_b200 import lib as L

    with pytest.raises(RuntimeError, match="no CPU path"):
        L.ptr(torch.zeros(3))


def test_integration_appendix_lists_every_exported_function():
    """INTEGRATION.md's appendix (tools/abi_table.py) names every function of include/*.h with the line it is declared on."""
    import subprocess
    import sys

    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    table = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "abi_table.py")], capture_output=True, text=True,
                           check=True).stdout
    rows = [r for r in table.splitlines() if r.startswith("| `sdb_")]
    assert len(rows) >= 60
    missing = [r for r in rows if r not in doc]
    assert not missing, "run `python tools/abi_table.py` and refresh INTEGRATION.md's appendix:\n" + "\n".join(missing[:5])
    functions = {re.match(r"\| `(sdb_\w+)`", r).group(1) for r in rows}
    types = {"sdb_gemm_args", "sdb_grid_cfg", "sdb_prompt_cfg", "sdb_unet_cfg", "sdb_vae_cfg", "sdb_packed_samples"}
    assert set(declared_symbols()) - types <= functions, sorted(set(declared_symbols()) - types - functions)
