"""Build a model written against cimba_b200/csrc/cmb_device.cuh into a library the C-ABI can load.

    python scripts/build_model.py path/to/my_model.cu [-o out.so] [extra nvcc flags]

The .cu file ends with CMB_EXPORT_MODEL(MyModel, "name").  Same flags as the library itself (sm_100a, -fmad=false: the
reference build never contracts a * b + c, and bit parity with it needs the same here).  Default output:
cimba_b200/lib/models/lib<stem>.so.  Load it with cimba_b200_model_load(path) / cimba_b200.load_model(path)."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g     # noqa: E402


def build(src: Path, out: Path = None, extra=()) -> Path:
    src = Path(src).resolve()
    out = Path(out) if out else ROOT / "cimba_b200/lib/models" / f"lib{src.stem}.so"
    out.parent.mkdir(parents=True, exist_ok=True)
    deps = [src, *(ROOT / "cimba_b200/csrc").glob("*.cuh"), *(ROOT / "cimba_b200/models").glob("*.cuh"), ROOT / "include/cimba_b200.h"]
    if out.exists() and out.stat().st_mtime >= max(p.stat().st_mtime for p in deps):
        return out
    subprocess.run([g._nvcc(), *g.NVCC_FLAGS, *extra, "-o", str(out), str(src)], check=True, cwd=ROOT)
    return out


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    args = sys.argv[1:]
    out = None
    if "-o" in args:
        k = args.index("-o")
        out = args[k + 1]
        del args[k:k + 2]
    print("built", build(args[0], out, args[1:]))
