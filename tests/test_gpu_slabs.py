"""Multi-slab parity: one process per slab, native peer-to-peer halo (chiml_gpu_halo_export / chiml_gpu_halo_bind), result gathered
on rank 0 and compared bit for bit with the single-rank output of the unmodified reference.  With >= 2 visible GPUs every slab has
its own GPU (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_slabs.py -m gpu`); the four-slab test also runs on ONE GPU, the
slabs sharing it as separate processes (CUDA IPC peer stores work within a device too), so the halo protocol is exercised by the
single-GPU test run as well."""
import os
import subprocess
import sys

import pytest

from chiml_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port() -> int:
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return so.getsockname()[1]


def _env(march):
    """Forced column length of the y-marching kernels for the worker processes (CHIML_B200_MARCH_NY, include/chiml_gpu.h
    chiml_gpu_set_march): slab-boundary planes stay single-plane work items, the columns next to them carry their y neighbours."""
    env = dict(os.environ)
    env.pop("CHIML_B200_MARCH_NY", None)
    if march:
        env["CHIML_B200_MARCH_NY"] = march
    return env


@pytest.mark.parametrize("march", [None, "5"], ids=["auto", "ny5"])
def test_two_slabs_over_nvlink_match_single_rank_reference(march):
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    cases = ["vac3d", "aniso_slab3d", "lorentz3d", "ml3d_two", "ml3d_four", "ml_te", "ml_tm", "tm_au", "te_vacuum", "c4_small", "flux3d", "te_flux", "tm_flux", "aniso_mixed3d", "ml3d_two+pair", "c4_small+pair", "mag3d", "mag3d_pml", "mag_tm", "mag_te", "pbc3d", "pbc3d_all", "pbc_tm", "pbc_te", "pbc_ml3d"]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "slab_gpu_worker.py")] + cases,
                       capture_output=True, text=True, timeout=900, env=_env(march))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    for c in cases:
        assert f"{c}: SLAB_GPU_OK" in r.stdout, r.stdout[-4000:]


@pytest.mark.parametrize("march", [None, "3", "1048576"], ids=["auto", "ny3", "whole"])
def test_four_slabs_match_single_rank_reference_even_on_one_gpu(march):
    """Four slabs on however many GPUs there are (ranks wrap around the devices).  c4_small at four slabs has a slab that holds
    only the rim of the emitter sheet (an emitter set without emitters), flux3d has flux surfaces cut by slab boundaries."""
    cases = ["aniso_slab3d", "ml3d_two", "c4_small", "flux3d", "aniso_mixed3d", "ml3d_two+pair", "ml_te+pair", "mag3d", "mag3d_pml", "mag_tm", "pbc3d", "pbc3d_all", "pbc_tm", "pbc_te", "pbc_ml3d"]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "slab_gpu_worker.py")] + cases,
                       capture_output=True, text=True, timeout=900, env=_env(march))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    for c in cases:
        assert f"{c}: SLAB_GPU_OK" in r.stdout, r.stdout[-4000:]


def test_eight_slabs_of_three_rows_match_single_rank_reference():
    """The cases bench.py steps across its N ranks before it times anything (HALO_PARITY_CASES), at the largest N: slabs of three
    grid rows -- every row but one is a slab-boundary row -- sharing however many GPUs there are."""
    sys.path.insert(0, ROOT)
    import bench
    cases = bench.HALO_PARITY_CASES
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=8", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "slab_gpu_worker.py")] + cases,
                       capture_output=True, text=True, timeout=900, env=_env(None))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    for c in cases:
        assert f"{c}: SLAB_GPU_OK" in r.stdout, r.stdout[-4000:]


@pytest.mark.parametrize("march", [None, "3"], ids=["auto", "ny3"])
def test_random_periodic_and_magnetic_inputs_on_three_and_four_slabs(march):
    """tests/fuzz/gen_inputs.rnd_pbc_case / rnd_mag_case on 3 and 4 slabs (sharing however many GPUs there are): periodic runs as a ring of slabs with
    objects that span the periodic faces, magnetic-dispersive objects cut by slab boundaries; expected arrays from the single-rank oracle."""
    for world, cases in ((3, ["fuzz:1:pbc", "fuzz:3:pbc", "fuzz:0:pbc", "fuzz:4:mag"]), (4, ["fuzz:6:pbc", "fuzz:4:pbc", "fuzz:9:pbc", "fuzz:0:mag", "fuzz:13:mag"])):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                            "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "slab_gpu_worker.py")] + cases,
                           capture_output=True, text=True, timeout=900, env=_env(march))
        assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
        for c in cases:
            assert f"{c}: SLAB_GPU_OK" in r.stdout, r.stdout[-4000:]
