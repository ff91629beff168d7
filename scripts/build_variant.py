"""Build a tuning variant of the C-ABI library into cimba_b200/lib/variants/ (run HERE, nvcc cross-compiles):

    python scripts/build_variant.py awacs_chunk4 -DAWACS_CHUNK=4
    python scripts/build_variant.py awacs_regs96 -maxrregcount=96

Select it on the GPU box with CIMBA_B200_LIB=$PWD/cimba_b200/lib/variants/<name>.so (scripts/sweep_variants.sh for
the M/M/1 bench, scripts/awacs_sweep.sh for AWACS).  Variants are git-ignored and travel with gpurun."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g     # noqa: E402


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    name, extra = sys.argv[1], sys.argv[2:]
    out = ROOT / "cimba_b200/lib/variants" / f"{name}.so"
    out.parent.mkdir(parents=True, exist_ok=True)
    cmd = [g._nvcc(), *g.NVCC_FLAGS, *extra, "-o", str(out), str(g.CSRC / "capi.cu")]
    subprocess.run(cmd, check=True, cwd=ROOT)
    print("built", out)


if __name__ == "__main__":
    main()
