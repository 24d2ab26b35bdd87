"""One-off wider GPU-vs-oracle fuzz (development aid): python tools/gpu_fuzz_once.py <first> <last> [ml]"""
import os
import pathlib
import sys
import tempfile
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "fuzz")):
    sys.path.insert(0, p)
import gen_inputs  # noqa: E402
import test_gpu_fuzz as T  # noqa: E402

ml = len(sys.argv) > 3 and sys.argv[3] == "ml"
bad = []
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    d = pathlib.Path(tempfile.mkdtemp())
    try:
        T.run_case(gen_inputs.rnd_ml_case(seed) if ml else gen_inputs.rnd_case(seed, steps=12, pulses="random"), d, 12)
    except Exception as e:  # noqa: BLE001
        bad.append((seed, str(e)[:200]))
        traceback.print_exc(limit=1)
print("GPU_FUZZ", "ml" if ml else "media", sys.argv[1], sys.argv[2], "bad:", bad)
