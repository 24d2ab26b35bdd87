"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): one process per GPU, native peer-to-peer halo
(chiml_gpu_halo_export / chiml_gpu_halo_bind), result gathered on rank 0 and compared bit for bit with the single-rank output of
the unmodified reference.  Run by hand with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_slabs.py -m gpu`."""
import os
import subprocess
import sys

import pytest

from chiml_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_two_slabs_over_nvlink_match_single_rank_reference():
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    cases = ["vac3d", "aniso_slab3d", "lorentz3d", "ml3d_two", "ml3d_four", "ml_te", "ml_tm", "tm_au", "te_vacuum", "c4_small", "flux3d", "te_flux", "tm_flux"]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29733", os.path.join(ROOT, "tests", "slab_gpu_worker.py")] + cases,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    for c in cases:
        assert f"{c}: SLAB_GPU_OK" in r.stdout, r.stdout[-4000:]
