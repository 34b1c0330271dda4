"""-m gpu, needs >= 2 GPUs (skipped otherwise): the pair-sharded solver across real GPUs."""
import socket
import subprocess
import sys

import pytest

import cases

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("case", ["c1_n256", "c1_em_n128"])
def test_sharded_solver_matches_single_gpu(case, native_lib):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
                        str(_free_port()), str(cases.ROOT / "tests" / "multigpu_check.py"), case],
                       capture_output=True, text=True, timeout=600,
                       env={**__import__("os").environ, "EMME_DENSE_NBO": "64"})
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "same_matrix=False" not in r.stdout and "same_omega=False" not in r.stdout
    assert "same_pic=True" in r.stdout and "same_pic=False" not in r.stdout


@pytest.mark.parametrize("case,shard_dense", [("c1_n256", True), ("c1_em_n128", False)])
def test_single_process_group_across_devices(case, shard_dense, native_lib, monkeypatch):
    """One process, one host thread and one handle per GPU, attached with emme_peer_attach (peer
    access enabled between the devices): the launcher-free binding INTEGRATION.md shows.  Bitwise the
    single-GPU iterates and matrices, like the one-process-per-GPU form."""
    import numpy as np
    import torch
    from emme_b200 import EigenSolver, Input, parallel
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    monkeypatch.setenv("EMME_DENSE_NBO", "64")
    native_lib.emme_peer_set_timeout(5.0)
    world = min(n_dev, 4)
    inp = Input(cases.input_path(case))
    p, n = inp.params()
    single = EigenSolver.from_input(inp)
    single.seed(inp.initial_guess())
    its = []
    for _ in range(3):
        single.newtonTraceSecantIteration()
        its.append((single.eigen_value, single.d_eigen_value))
    g = parallel.LocalShardedGroup(p, n, *inp.tables(), devices=list(range(world)), shard_dense=shard_dense)
    g.seed(inp.initial_guess())
    for k in range(3):
        g.newtonTraceSecantIteration()
        for r in g.ranks:
            assert (r.eigen_value, r.d_eigen_value) == its[k], (case, k, r.eigen_value, its[k])
    A1 = single.eigen_matrix
    for r in g.ranks:
        assert np.array_equal(r.eigen_matrix, A1)
    g.close()
    single.close()
    native_lib.emme_peer_set_timeout(20.0)
    monkeypatch.setenv("EMME_DENSE_NBO", "0")
    EigenSolver.from_input(Input(cases.input_path("c1_n32"))).close()     # restore the process-wide outer block
