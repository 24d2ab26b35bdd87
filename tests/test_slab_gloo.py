"""N > 1 path on CPU (world_size 2 and 3, backend gloo): the y-slab plans of the host-side setup, stepped by the CPU checker
through the ghost-row exchange protocol of chiml_b200/slab.py (the protocol the CUDA engine implements with peer-to-peer
stores), must reproduce the single-rank output of the reference bit for bit -- fields, pole grids and emitter density
matrices, including an emitter block and an oriented-dipole slab cut by the slab boundary."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


CASES = [("aniso_slab3d", 2), ("ml3d_two", 2), ("ml3d_four", 3), ("ml_te", 2), ("c4_small", 2), ("flux3d", 3), ("te_flux", 2), ("tm_flux", 2),
         # oriented-dipole objects of finite y extent and different pole counts: slabs without node cells
         ("aniso_mixed3d", 4), ("aniso_mixed3d", 3),
         # magnetic-dispersive media: B / H / M cells cut by slab boundaries, the H-side CPML on B
         ("mag3d", 3), ("mag3d_pml", 2), ("mag_tm", 3), ("mag_te", 2),
         # periodic boundaries: the slabs form a ring, the last slab's wrap row comes from slab 0 (chiml_b200/slab.py)
         ("pbc3d", 2), ("pbc3d", 3), ("pbc3d_all", 4), ("pbc_tm", 3), ("pbc_te", 2),
         # ... with an emitter sheet across the whole periodic cell (C4 run periodic, in miniature): the seam rows of Ey travel at the end of the
         # step, the emitters' polarisation boxes do not cross the seam
         ("pbc_ml3d", 2), ("pbc_ml3d", 3), ("pbc_ml3d", 4),
         # random inputs (tests/fuzz/gen_inputs.py), expected arrays from the single-rank oracle
         ("fuzz:4", 4), ("fuzz:12", 3), ("fuzz:33", 3), ("fuzz:10:ml", 4), ("fuzz:14:ml", 2),
         # random periodic inputs (objects spanning the periodic faces cross the seam of the slab ring) and random magnetic media
         ("fuzz:1:pbc", 3), ("fuzz:3:pbc", 4), ("fuzz:6:pbc", 2), ("fuzz:0:pbc", 3), ("fuzz:4:pbc", 4), ("fuzz:9:pbc", 2),
         ("fuzz:0:mag", 2), ("fuzz:4:mag", 3), ("fuzz:13:mag", 2)]


@pytest.mark.parametrize("world", [2, 3, 4])
def test_slab_protocol_matches_single_rank_reference(world):
    """Every case of CASES with this slab count in ONE launch of the process group (its start-up costs more than a case)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "chiml_b200", "host")], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, stdout=subprocess.DEVNULL)
    cases = [c for c, w in CASES if w == world]
    import socket
    with socket.socket() as so:                 # a free port: parallel test runs (pytest -n) must not meet on one
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "slab_worker.py")] + cases,
                       capture_output=True, text=True, timeout=1200)
    for c in cases:
        assert f"{c}: SLAB_OK" in r.stdout, f"{c} on {world} slabs\n" + (r.stdout[-3000:] + r.stderr[-3000:])
    assert r.returncode == 0, (r.stdout[-3000:] + r.stderr[-3000:])
