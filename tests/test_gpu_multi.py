"""Multi-GPU paths on one node (skipped with fewer than 2 GPUs): the fused in-kernel cross-GPU weight sum against the
NCCL all-reduce, and resampling over all shards against one process holding all particles.  Each check is a
torchrun job of tools/*.py (one process per GPU, rendezvous on 127.0.0.1)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(script, n, port):
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    return r.returncode, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_fused_allreduce_equals_nccl(cuda_required):
    rc, out = _torchrun("fused_check.py", 2, 29531)
    assert rc == 0, out


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_global_resampling_is_independent_of_the_number_of_ranks(cuda_required):
    rc, out = _torchrun("resample_check.py", 2, 29532)
    assert rc == 0, out
    assert out.count("identical to the single-process result") == 4, out
