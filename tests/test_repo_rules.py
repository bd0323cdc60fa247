"""CPU: structural rules of this repository.  The oracle is test infrastructure (only tests/, smoke() and bench.py's CPU legs may
touch it), the product never reads the reference tree at run time, and there is no CPU fallback switch in the package."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py(*parts):
    return sorted(glob.glob(os.path.join(ROOT, *parts)))


def test_product_and_tools_never_import_the_oracle():
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    for f in _py("linnaeus_b200", "*.py") + _py("tools", "*.py"):
        assert not pat.search(open(f).read()), f
    # bench.py may use it only inside the CPU-baseline functions
    src = open(os.path.join(ROOT, "bench.py")).read()
    for m in pat.finditer(src):
        head = src[: m.start()]
        fn = re.findall(r"^def (\w+)\(", head, re.M)[-1]
        assert fn.startswith("cpu_"), f"bench.py imports oracle inside {fn}()"


def test_product_never_reads_the_reference_tree():
    for f in _py("linnaeus_b200", "*.py") + _py("linnaeus_b200", "csrc", "*") + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        txt = open(f, errors="ignore").read()
        assert "/root/reference" not in txt, f
        assert not re.search(r"^\s*(from|import)\s+linnaeus(\s|\.|$)", txt, re.M) or f.endswith(("registry.py", "loss.py")), f


def test_no_cpu_fallback_paths_in_the_package():
    # every host module that launches kernels refuses CPU tensors instead of falling back to torch ops
    for name in ("mformer_v1.py", "mformer_v0.py", "metrics.py", "aug.py", "optim.py"):
        txt = open(os.path.join(ROOT, "linnaeus_b200", name)).read()
        assert re.search(r"no CPU fallback|CUDA parameters", txt), name
