"""Commit-time tile classification of the CUDA engine, the CUDA-free part (chiml_b200/csrc/chiml_tiles.hpp): a tile is handed to
k_uniform -- which never reads the per-cell info plane -- only when its cells decompose into rectangles of one info value each.
tests/cpu/tile_rects_check.cpp restates the device-side tile summary on the CPU, feeds designed and random tiles through
tile_rectangles and verifies every accepted decomposition cell by cell."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_accepted_tile_decompositions_are_exact(tmp_path):
    exe = str(tmp_path / "tile_rects_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpu", "tile_rects_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "TILE_RECTS_OK" in r.stdout, r.stdout + r.stderr
