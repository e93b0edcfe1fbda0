"""Regenerates tests/golden/loader/ from the UNMODIFIED reference (oracle/_ref): copies of the reference's
genotype-matrix fixtures (tests/testthat/helper_genotypes*.txt) and, per fixture, what the reference's
loader made of it with the reference's map loaded (marker order, genotypes, names, ids) and the draws it
consumed (phases of unphased cell styles).

Run in the build container (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden_loader.py
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402

T = "/root/reference/tests/testthat/"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loader")
FIXTURES = ["helper_genotypes.txt", "helper_genotypes_slash_nocorner.txt", "helper_genotypes_inverted_counts.txt",
            "helper_genotypes_noheader_nullable_counts.txt", "helper_genotypes_long.txt"]


def main():
    os.makedirs(OUT, exist_ok=True)
    ref.set_print_mode(0)
    for name in FIXTURES:
        shutil.copyfile(T + name, os.path.join(OUT, name))
        r = ref.RefSim()
        r.load_map(T + "helper_map.txt")
        ref.seed(20261021)
        ref.record_start()
        g = r.load_genotypes(T + name)
        tape = ref.record_stop().draws_only()
        x = r.export_group(g)
        out = dict(marker_names=r.marker_names(), names=r.group_names(g), ids=x["ids"].tolist(), genotypes=x["genotypes"].tolist(),
                   tape_values=[float.hex(v) for v in tape.values.tolist()], tape_kinds=tape.kinds.tolist())
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(out, f)
        print(name, len(out["names"]), "genotypes,", len(tape), "draws")
        r.close()


if __name__ == "__main__":
    main()
