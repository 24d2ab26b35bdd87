"""Regenerates tests/golden/out_expected/te_hpow: a TE grid with an SI-scaled H-power TXT detector over a 3 x 2 box (DTC/parallelDTC.hpp:87-91,
parallelDTCOutputFxn.hpp pwrOutputFunction) and a COUT detector of Ey (DTC/parallelDTC_COUT.cpp), run by the UNMODIFIED reference
(oracle/_ref/chiml_ref).  The console lines of the COUT detector are kept as cout.txt (the lines that start with a tab or with the
sample time; the reference prints them while it steps)."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from chiml_b200 import inputs as I  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "chiml_ref")
DT = I.default_dt(100)


def te_hpow():
    cfg = I.c1_te_vacuum(n=47, steps=60, pml_cells=8, out="out/hp")
    for s in cfg["SourceList"]:
        for p in s["PulseList"]:
            p["t_0"] = 0.25
            p["cutoff"] = 2.5
    cfg["DetectorList"] = [I.detector([0.03, 0.01, 0.0], [0.02, 0.01, 0.0], "H_pow", "out/hp/pow", dtc_class="txt", time_int=2.0 * DT * 1.0000001, si=True),
                           I.detector([-0.04, 0.02, 0.0], [0.01, 0.01, 0.0], "Ey", "out/hp/cout", dtc_class="cout", time_int=5.0 * DT * 1.0000001)]
    return cfg


def cout_lines(stdout):
    """The COUT detector's lines: sample header `t\\tx\\ty\\tz\\t` and rows that start with a tab."""
    keep = []
    for ln in stdout.splitlines():
        if ln.startswith("\t") or (ln.count("\t") == 4 and ln.endswith("\t")):
            keep.append(ln)
    return "\n".join(keep) + "\n"


if __name__ == "__main__":
    out = os.path.join(HERE, "out_expected", "te_hpow")
    os.makedirs(out, exist_ok=True)
    work = tempfile.mkdtemp(prefix="hpow_")
    I.write(te_hpow(), os.path.join(work, "te_hpow.json"))
    shutil.copy(os.path.join(work, "te_hpow.json"), os.path.join(out, "te_hpow.json"))
    r = subprocess.run([REF, "te_hpow.json"], cwd=work, capture_output=True, text=True, check=True)
    shutil.copy(os.path.join(work, "out", "hp", "pow_field_0.dat"), os.path.join(out, "pow_field_0.dat"))
    open(os.path.join(out, "cout.txt"), "w").write(cout_lines(r.stdout))
    print(open(os.path.join(out, "cout.txt")).read()[:400])
    print(open(os.path.join(out, "pow_field_0.dat")).read()[:600])
