"""The forward's alternative code paths (selected by environment switches libeotpatch reads once per process) give the
bits of the default path: the single persistent cooperative kernel (EOT_FWD_FUSED=1), plain launches instead of
programmatic dependent launch (EOT_PDL=0), the register-staged image pass instead of the bulk-copy pipeline
(EOT_PREPASS_BULK_ON=0).  The default path itself is checked against the oracle / the reference fixtures in
test_gpu_forward.py; here each variant runs in a child process and is compared with the default child."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = os.path.join(ROOT, "tests", "_variant_child.py")
SWITCHES = ("EOT_FWD_FUSED", "EOT_PDL", "EOT_PREPASS_BULK_ON", "EOT_FWD_GROUPS", "EOT_KERNEL_TIMES")


def _run(tmp_path, name, env_extra):
    env = {k: v for k, v in os.environ.items() if k not in SWITCHES}
    env.update(env_extra)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    out = str(tmp_path / f"{name}.npz")
    r = subprocess.run([sys.executable, CHILD, out], env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    return np.load(out)


@pytest.fixture(scope="module")
def default_run(tmp_path_factory):
    return _run(tmp_path_factory.mktemp("variants"), "default", {})


@pytest.mark.gpu
@pytest.mark.parametrize("name,env,launches", [
    ("fused", {"EOT_FWD_FUSED": "1"}, 1),            # the whole forward is one launch
    ("no_pdl", {"EOT_PDL": "0"}, 5),
    ("register_pass", {"EOT_PREPASS_BULK_ON": "0"}, 5),
    ("two_groups", {"EOT_FWD_GROUPS": "2"}, 10),      # two image groups on two streams
])
def test_variant_equals_default(tmp_path, default_run, name, env, launches):
    got = _run(tmp_path, name, env)
    for tag in ("affine", "perspective", "large_patch"):
        assert int(default_run[tag + "_launches"]) == 5
        assert int(got[tag + "_launches"]) == launches, (name, tag, int(got[tag + "_launches"]))
        assert np.array_equal(got[tag + "_out"], default_run[tag + "_out"]), (name, tag)
        # the backward reads what the forward left in the workspace (route bytes, plans, u texels, luma sums); its
        # own summation order is the same in every variant
        g0, g1 = default_run[tag + "_grad"], got[tag + "_grad"]
        rel = np.linalg.norm((g1 - g0).astype(np.float64)) / max(np.linalg.norm(g0.astype(np.float64)), 1e-30)
        assert rel <= 1e-6, (name, tag, rel)
