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

# flux3d_save: flux3d with "save" set on the 3-D box region: the reference then also writes <name>_fields.dat (saveFields,
# DTC/parallelFlux.hpp:616-659), the frequency-domain surface fields a later run subtracts as incident fields
cfg = json.load(open(os.path.join(HERE, "flux3d.json")))
cfg["FluxList"][0]["save"] = True
os.makedirs(os.path.join(HERE, "out_expected", "flux3d_save"), exist_ok=True)
json.dump(cfg, open(os.path.join(HERE, "out_expected", "flux3d_save", "flux3d_save.json"), "w"), indent=1)
work = tempfile.mkdtemp(prefix="fluxref_")
shutil.copy(os.path.join(HERE, "out_expected", "flux3d_save", "flux3d_save.json"), work)
subprocess.run([REF, "flux3d_save.json", "--quiet"], cwd=work, check=True, stdout=subprocess.DEVNULL)
shutil.copy(os.path.join(work, cfg["FluxList"][0]["name"] + "_fields.dat"), os.path.join(HERE, "out_expected", "flux3d_save", "box_fields.dat"))
print("flux3d_save", os.path.getsize(os.path.join(HERE, "out_expected", "flux3d_save", "box_fields.dat")), "bytes")

# flux3d_load: the usual two-run normalisation.  Run 1 = flux3d WITHOUT its scatterer, box region saved -> empty_fields.dat; run 2 =
# flux3d with "load" on the box region, which starts its surface fields from minus the saved ones (loadFields(-1.0),
# DTC/parallelFlux.hpp:664-722): box.dat then holds the flux of the scattered field alone.
out = os.path.join(HERE, "out_expected", "flux3d_load")
os.makedirs(out, exist_ok=True)
cfg = json.load(open(os.path.join(HERE, "flux3d.json")))
cfg["ObjectList"] = []
cfg["FluxList"] = cfg["FluxList"][:1]
cfg["FluxList"][0]["save"] = True
work = tempfile.mkdtemp(prefix="fluxref_")
json.dump(cfg, open(os.path.join(work, "empty.json"), "w"), indent=1)
subprocess.run([REF, "empty.json", "--quiet"], cwd=work, check=True, stdout=subprocess.DEVNULL)
shutil.copy(os.path.join(work, cfg["FluxList"][0]["name"] + "_fields.dat"), os.path.join(out, "empty_fields.dat"))
cfg = json.load(open(os.path.join(HERE, "flux3d.json")))
cfg["FluxList"][0]["load"] = True
cfg["FluxList"][0]["incd_fileds"] = "empty_fields.dat"
json.dump(cfg, open(os.path.join(out, "flux3d_load.json"), "w"), indent=1)
work = tempfile.mkdtemp(prefix="fluxref_")
shutil.copy(os.path.join(out, "flux3d_load.json"), work)
shutil.copy(os.path.join(out, "empty_fields.dat"), work)
subprocess.run([REF, "flux3d_load.json", "--quiet"], cwd=work, check=True, stdout=subprocess.DEVNULL)
shutil.copy(os.path.join(work, cfg["FluxList"][0]["name"] + ".dat"), os.path.join(out, "box.dat"))
print("flux3d_load", open(os.path.join(out, "box.dat")).read()[:400])

# freq3d: the files of the frequency detectors (parallelDetectorFREQ_Base::toFile / toMap with incident fields, as main.cpp:74-107 calls
# them) beside a flux box
work = tempfile.mkdtemp(prefix="fluxref_")
shutil.copy(os.path.join(HERE, "freq3d.json"), work)
subprocess.run([REF, "freq3d.json", "--quiet"], cwd=work, check=True, stdout=subprocess.DEVNULL)
out = os.path.join(HERE, "out_expected", "freq3d")
os.makedirs(out, exist_ok=True)
for n in sorted(os.listdir(os.path.join(work, "out", "fq"))):
    if n.startswith("dtc_"):
        continue
    shutil.copy(os.path.join(work, "out", "fq", n), os.path.join(out, n))
    print("freq3d", n)

# tfsf3d: a flux box around a sphere lit by a TFSF plane wave: getFlux normalises with the incident-field series the propagator
# recorded (DTC/parallelFlux.hpp:455-482); the series themselves (--incd-dump) are kept as the fixture the host-side post-processing reads
work = tempfile.mkdtemp(prefix="fluxref_")
shutil.copy(os.path.join(HERE, "tfsf3d.json"), work)
subprocess.run([REF, "tfsf3d.json", "--quiet", "--incd-dump", "incd.bin"], cwd=work, check=True, stdout=subprocess.DEVNULL)
out = os.path.join(HERE, "out_expected", "tfsf3d")
os.makedirs(out, exist_ok=True)
shutil.copy(os.path.join(work, "out", "t3", "box.dat"), os.path.join(out, "box.dat"))
shutil.copy(os.path.join(work, "incd.bin"), os.path.join(out, "incd.bin"))
print("tfsf3d", open(os.path.join(out, "box.dat")).read()[:300])

for case in CASES:
    work = tempfile.mkdtemp(prefix="fluxref_")
    shutil.copy(os.path.join(HERE, case + ".json"), work)
    subprocess.run([REF, case + ".json", "--quiet"], cwd=work, check=True, stdout=subprocess.DEVNULL)
    out = os.path.join(HERE, "out_expected", case)
    os.makedirs(out, exist_ok=True)
    for fl in json.load(open(os.path.join(HERE, case + ".json")))["FluxList"]:
        shutil.copy(os.path.join(work, fl["name"] + ".dat"), os.path.join(out, os.path.basename(fl["name"]) + ".dat"))
        print(case, fl["name"])
