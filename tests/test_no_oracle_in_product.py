"""The oracle is test infrastructure: nothing under av1_base_b200/ (Python or C++/CUDA) may import,
include, link or execute anything under oracle/."""
import glob, os, re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_tree_never_touches_the_oracle():
    bad = []
    for pat in ("*.py", "csrc/*.cu", "csrc/*.cc", "csrc/*.cpp", "csrc/*.h", "csrc/*.cuh", "csrc/Makefile"):
        for f in glob.glob(os.path.join(ROOT, "av1_base_b200", pat)):
            txt = open(f, errors="ignore").read()
            for m in re.finditer(r"^\s*(from\s+oracle|import\s+oracle|#include\s+\"[^\"]*oracle[^\"]*\"|.*liboracle)", txt, re.M):
                bad.append((os.path.relpath(f, ROOT), m.group(0).strip()))
    assert not bad, bad


def test_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", os.path.join(ROOT, "av1_base_b200", "libav1b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
