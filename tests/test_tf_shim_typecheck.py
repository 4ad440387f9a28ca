"""tf_ops/eot_patch_ops.cc (the tf.load_op_library shim a maintainer of the reference builds next to TensorFlow,
/root/reference/attacker.py:59 is where its ops plug in) is type-checked against include/eotpatch.h with g++.

TensorFlow is not installed here, so the TensorFlow headers are replaced by a declarations-only stand-in
(tests/tf_stub/, test infrastructure).  What this pins: every call the shim makes into the C ABI has the argument
count, pointer types and constness the header declares, and the structs carry the fields the shim fills.  What it does
not pin: anything about TensorFlow itself (DESIGN.md section 2 keeps the shim listed as source only)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tf_ops", "eot_patch_ops.cc")


def _syntax_only(path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no C++ compiler")
    return subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-comment",
                           "-I", os.path.join(ROOT, "tests", "tf_stub"), "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tf_ops"), path], capture_output=True, text=True)


def test_shim_type_checks_against_the_c_abi():
    r = _syntax_only(SHIM)
    assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("old,new", [
    # one argument fewer in a C-ABI call
    ("gp->flat<float>().data(), /*accumulate=*/0, StreamOf(ctx)", "gp->flat<float>().data(), StreamOf(ctx)"),
    # wrong pointer type for the CSR row splits
    ("splits.flat<tf::int32>().data()", "splits.flat<float>().data()"),
    # a struct field the header does not have
    ("s.max_scale = 1.0f;", "s.maximum_scale = 1.0f;"),
    # an entry point the header does not declare
    ("score_candidate_offset(&s, &cand_off)", "score_candidates_offset(&s, &cand_off)"),
])
def test_the_check_notices_abi_drift(tmp_path, old, new):
    src = open(SHIM).read()
    assert old in src
    bad = tmp_path / "eot_patch_ops_mutated.cc"
    bad.write_text(src.replace(old, new, 1))
    r = _syntax_only(str(bad))
    assert r.returncode != 0, "the stand-in headers let an ABI mismatch through"


def test_shim_uses_every_entry_point_the_binding_needs():
    """The ops the Python wrapper (tf_ops/eot_patch_tf.py) loads are the ones the shim registers."""
    import re
    cc = open(SHIM).read()
    py = open(os.path.join(ROOT, "tf_ops", "eot_patch_tf.py")).read()
    registered = set(re.findall(r'REGISTER_OP\("(\w+)"\)', cc))
    kernels = set(re.findall(r'REGISTER_KERNEL_BUILDER\(Name\("(\w+)"\)', cc))
    assert registered == kernels and registered
    snake = {re.sub(r"(?<!^)(?=[A-Z])", "_", n).lower() for n in registered}
    used = set(re.findall(r"_ops\.(\w+)\(", py)) | set(re.findall(r"_mod\.(\w+)\(", py)) | set(re.findall(r"\.(eot_\w+)\(", py))
    assert used & snake, (used, snake)
    assert used <= snake, f"the wrapper calls ops the shim does not register: {used - snake}"
