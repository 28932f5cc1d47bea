#!/bin/bash
# Dev-time check (build container only: needs /root/reference): the reference's own test-suite run against this
# package aliased as `tomosipo`, with a stub `astra` that only answers use_cuda().  Without a GPU the 5 tests of
# tests/test_torch_support.py that project fail with "no CUDA device available"; everything else must pass.
set -e
T=$(mktemp -d)
cp -r /root/reference/tests "$T/tests"
cat > "$T/conftest.py" <<'PY'
import sys, types
sys.path.insert(0, "/root/repo")
import numpy as np
_array = np.array
def array(obj, *a, **kw):            # NumPy 2: copy=False used to mean "copy if needed"
    if kw.get("copy", True) is False:
        kw["copy"] = None
    return _array(obj, *a, **kw)
np.array = array
import tomosipo_b200
import tomosipo_b200.torch_support
astra = types.ModuleType("astra")
astra.use_cuda = tomosipo_b200.cuda_available
sys.modules["astra"] = astra
sys.modules["tomosipo"] = tomosipo_b200
for name in list(sys.modules):
    if name.startswith("tomosipo_b200."):
        sys.modules["tomosipo." + name[len("tomosipo_b200."):]] = sys.modules[name]
PY
cd "$T"
python -m pytest tests -q -p no:cacheprovider -W ignore --ignore tests/test_qt.py --ignore tests/test_svg.py \
    --ignore tests/test_odl.py --ignore tests/test_documentation.py | tail -8
