"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/fgp_b200.h declares, the ctypes
structures match the header, the host layer refuses to run without CUDA, and the multi-rank sharding logic works over
gloo with world_size 2."""
import os
import re
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    from fastgaussianprocesses_b200 import _lib
    h = _lib.load()
    text = open(os.path.join(ROOT, "include", "fgp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(fgp_[A-Za-z0-9_]+)\s*\(", text))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(h, name), "libfgp_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
    assert set(_lib.SIGNATURES) == declared
    assert h.fgp_version() == 100


def test_ctypes_struct_layouts_match_header():
    import ctypes
    from fastgaussianprocesses_b200 import _lib
    # 9 ints + pad, 1 double, 11 pointers ; 3 ints + pad, 9 doubles
    assert ctypes.sizeof(_lib.FitLayout) == 40 + 8 + 11 * 8
    assert ctypes.sizeof(_lib.FitOptions) == 16 + 9 * 8
    assert _lib.fit_state_doubles(11, 4) == 32 + 33 + 3 + 1
    assert ctypes.sizeof(_lib.FitProblem) == 14 * 8


def test_no_cpu_fallback():
    import fastgaussianprocesses_b200 as fgp
    from fastgaussianprocesses_b200 import _lib
    torch.set_default_dtype(torch.float64)
    with pytest.raises(RuntimeError):
        fgp.FastGPLattice(2, device="cpu")
    with pytest.raises(RuntimeError):
        _lib.fwht(torch.zeros(8))
    with pytest.raises(RuntimeError):
        _lib.post_mean(0, torch.zeros(4, 2), torch.zeros(8, 2), [2, 2], 0, 1.0, [1.0, 1.0], torch.zeros(1, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fastgaussianprocesses_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU fallback", ""), fn


def test_shard_bounds_cover_everything():
    from fastgaussianprocesses_b200.distributed import shard_bounds
    for m in (0, 1, 7, 8, 1000, 2 ** 24):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_bounds(m, world, r)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == m


def _worker(rank, world, port, m, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from fastgaussianprocesses_b200.distributed import sharded_rows, gather_fit_results
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    torch.set_default_dtype(torch.float64)
    x = torch.rand((m, 3), generator=torch.Generator().manual_seed(5))
    fn = lambda rows: torch.stack([rows.sum(1), (rows ** 2).sum(1)], 0)  # (..., k) with a leading batch dim
    full = sharded_rows(fn, x)
    ok = torch.equal(full, fn(x))
    g = gather_fit_results(torch.tensor([float(rank), 2.0 * rank]))
    ok = ok and torch.equal(g, torch.tensor([[0.0, 0.0], [1.0, 2.0]]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("m", [10, 7, 1])
def test_sharded_rows_gloo_world2(m):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
