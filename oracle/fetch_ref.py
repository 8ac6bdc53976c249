"""TEST / BENCH INFRASTRUCTURE ONLY: stage the UNMODIFIED reference next to the oracle so that it travels to the GPU box.

    python -m oracle.fetch_ref            # copies /root/reference/{data_utils.py, layers/, models/, run_simulation.py,
                                          #   train_utils.py, data/EXP/raw/GRAPHSAT.pkl, data/sr25/raw/sr251256.g6}
                                          # into oracle/_ref/   (git-ignored, NOT gpurun-ignored)

`/root/reference` exists only in the build container.  `gpurun` snapshots the working tree (git-ignored files
included), so after this recipe has run the `-m gpu` tests and `bench.py --impl reference` find the reference's own
files under oracle/_ref/ on the GPU box: the unmodified models/GNNs.py is run on a B200 over the drop-in layers
(tests/test_reference_models_gpu.py) and the reference arm of the bench times the reference's own modules
(`cpu_baseline.kind = "reference"`).  Nothing under oracle/_ref/ is ever committed, and nothing in kpgnn_b200/ reads it.
`__graft_entry__.build()` runs this recipe whenever /root/reference is present.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("KPGNN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
ITEMS = ("data_utils.py", "train_utils.py", "run_simulation.py", "layers", "models",
         os.path.join("data", "EXP", "raw", "GRAPHSAT.pkl"), os.path.join("data", "sr25", "raw", "sr251256.g6"))


def fetch(verbose=False):
    """Returns True when oracle/_ref/ is populated (freshly or from an earlier call)."""
    if not os.path.isfile(os.path.join(SRC, "data_utils.py")):
        return os.path.isfile(os.path.join(DST, "data_utils.py"))
    for item in ITEMS:
        s, d = os.path.join(SRC, item), os.path.join(DST, item)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
        if verbose:
            print("staged", item)
    return True


if __name__ == "__main__":
    ok = fetch(verbose=True)
    print("oracle/_ref ready" if ok else "no reference tree at %s" % SRC)
    sys.exit(0 if ok else 1)
