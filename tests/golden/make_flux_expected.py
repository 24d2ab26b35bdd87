"""Regenerates tests/golden/out_expected/<case>/<flux name>.dat: the flux spectra files the UNMODIFIED reference
(oracle/_ref/chiml_ref, outputs enabled) writes for the committed flux cases.  Run where /root/reference was built."""
import json
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "chiml_ref")
CASES = ["tm_flux", "te_flux", "flux3d"]

for case in CASES:
    work = tempfile.mkdtemp(prefix="fluxref_")
    shutil.copy(os.path.join(HERE, case + ".json"), work)
    subprocess.run([REF, case + ".json", "--quiet"], cwd=work, check=True, stdout=subprocess.DEVNULL)
    out = os.path.join(HERE, "out_expected", case)
    os.makedirs(out, exist_ok=True)
    for fl in json.load(open(os.path.join(HERE, case + ".json")))["FluxList"]:
        shutil.copy(os.path.join(work, fl["name"] + ".dat"), os.path.join(out, os.path.basename(fl["name"]) + ".dat"))
        print(case, fl["name"])
