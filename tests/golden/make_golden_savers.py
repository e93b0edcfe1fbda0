"""Regenerates tests/golden/savers/ from the UNMODIFIED reference (oracle/_ref): the text files
tests/scenarios.saver_scheme writes on the reference's own 6-founder fixture
(tests/testthat/helper_*.txt), and the draw tape the reference consumed meanwhile.

Run in the build container (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden_savers.py
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref  # noqa: E402
import scenarios  # noqa: E402

T = "/root/reference/tests/testthat/"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "savers")


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    r = ref.RefSim()
    ref.set_print_mode(0)
    g, m, e = r.load(T + "helper_genotypes.txt", T + "helper_map.txt", T + "helper_eff.txt")
    e2 = r.load_effects(T + "helper_eff_2.txt")
    ref.seed(20261020)
    ref.record_start()
    files = scenarios.saver_scheme(r, g, (e, e2), OUT)
    tape = ref.record_stop().draws_only()
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(dict(files=files, tape_values=[float.hex(v) for v in tape.values.tolist()], tape_kinds=tape.kinds.tolist()), f)
    r.close()
    print("%d golden saver files written to %s" % (len(files), OUT))


if __name__ == "__main__":
    main()
