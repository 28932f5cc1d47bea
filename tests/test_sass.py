"""The built library really contains the Blackwell code paths DESIGN.md describes (checked on the CPU with
cuobjdump, as /opt/skills/guides/B200_PROFILING.md suggests): sm_100a SASS, TMA tensor loads, mbarrier
instructions and packed fp32x2 arithmetic in the two hot kernels, and no tensor-core or texture instructions."""
import re
import shutil
import subprocess

import pytest

from tomosipo_b200 import _backend as B

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def sass_of(pattern):
    B.lib()
    out = subprocess.run([CUOBJDUMP, "-sass", B.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not usable here")
    text = out.stdout
    assert "sm_100a" in text
    chunks = re.split(r"\n\s*Function : ", text)
    return [c for c in chunks if re.match(pattern, c)]


@pytest.mark.parametrize("kernel, needs", [
    (r"_ZN3tsp13bp_tma_kernelILb1ELi64E", ["UTMALDG.3D", "SYNCS", "FFMA2", "LDS"]),      # tall tile: cfg 3 / cfg 4
    (r"_ZN3tsp13bp_tma_kernelILb1ELi32E", ["UTMALDG.3D", "SYNCS", "FFMA2", "LDS"]),
    (r"_ZN3tsp13fp_tma_kernelILb1ELb1ELi8ELi2E", ["UTMALDG.3D", "SYNCS", "FFMA2", "FADD2", "LDS"]),
])
def test_hot_kernels_use_tma_mbarriers_and_packed_fp32(kernel, needs):
    fns = sass_of(kernel)
    assert len(fns) == 1, [f[:60] for f in fns]
    body = fns[0]
    for mnemonic in needs:
        assert mnemonic in body, f"{mnemonic} missing from {kernel}"
    # the path is not a contraction and samples from shared memory: no tensor-core, no texture instructions
    for absent in ("HMMA", "UTCHMMA", "UTCQMMA", "TEX.", "TLD"):
        assert absent not in body, f"unexpected {absent} in {kernel}"
    # the inner loops keep their accumulators in registers: at most the handful of spill slots ptxas gives the
    # producer warp's once-per-32-angles fp64 set-up (12 bytes in the ZPT = 64 build), nothing in the unrolled loops
    assert body.count("STL") <= 6 and body.count("LDL") <= 6, "local-memory spills in a hot kernel"
