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
