"""CPU tests of the drop-in boundary: libb2cuda.so loads, exports every entry point include/b2cuda.h declares,
its plain-C records have the layout the host side assumes, and compute calls fail loudly without a device."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import b2cuda
import b2cuda_types as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b2cuda.h")


def declared_functions():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"B2CU_API\s+[\w\s\*]+?\b(b2cu\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = b2cuda.load()
    names = declared_functions()
    assert len(names) >= 20
    assert sorted(names) == sorted(b2cuda.API)
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.b2cuVersion()


def test_library_is_sm100a_native():
    out = subprocess.run(["cuobjdump", "-lelf", b2cuda.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_record_layouts_match_header(tmp_path):
    prog = tmp_path / "sizes.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "b2cuda.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(b2cuBody), sizeof(b2cuShape),'
        ' sizeof(b2cuProxy), sizeof(b2cuManifold), sizeof(b2cuContact), sizeof(b2cuWorldDef), sizeof(b2cuStepInfo),'
        ' offsetof(b2cuBody, flags), offsetof(b2cuProxy, fixture), offsetof(b2cuContact, manifold),'
        ' sizeof(b2cuBodyState), offsetof(b2cuBodyState, vx), sizeof(b2cuShardLink),'
        ' sizeof(b2cuDistanceResult), sizeof(b2cuSweep), offsetof(b2cuSweep, a0), sizeof(b2cuToiResult),'
        ' sizeof(b2cuJoint), offsetof(b2cuJoint, impulse), offsetof(b2cuStepInfo, toiMinKey));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [T.BODY.itemsize, T.SHAPE.itemsize, T.PROXY.itemsize, T.MANIFOLD.itemsize, T.CONTACT.itemsize,
            T.WORLD_DEF.itemsize, T.STEP_INFO.itemsize, T.BODY.fields["flags"][1], T.PROXY.fields["fixture"][1],
            T.CONTACT.fields["manifold"][1], T.BODY_STATE.itemsize, T.BODY_STATE.fields["vx"][1], T.SHARD_LINK.itemsize,
            T.DISTANCE_RESULT.itemsize, T.SWEEP.itemsize, T.SWEEP.fields["a0"][1], T.TOI_RESULT.itemsize,
            T.JOINT.itemsize, T.JOINT.fields["impulse"][1], T.STEP_INFO.fields["toiMinKey"][1]]
    assert got == want


def test_header_is_plain_c():
    # the boundary must be bindable from C (cgo / JNI / ctypes): compile it as C99 with warnings as errors
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER], check=True)


def test_no_cpu_fallback_without_device():
    if b2cuda.device_count() > 0:
        pytest.skip("a device is present; the no-device path is exercised on the CPU box")
    with pytest.raises(b2cuda.B2cuError) as e:
        b2cuda.World()
    assert e.value.code == T.ERR_NO_DEVICE
    with pytest.raises(b2cuda.B2cuError):
        b2cuda.sincos(np.zeros(4, np.float32))


def test_product_never_imports_oracle():
    """The product path must not route through the oracle: nothing under box2d-mt_b200/ may reference it."""
    pkg = os.path.join(ROOT, "box2d-mt_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"\bimport ref\b|oracle/|libb2ref|b2ref_|b2o_", text):
                    # mentions in comments that cite the oracle's algorithm are allowed only for b2o_math
                    lines = [ln for ln in text.splitlines()
                             if re.search(r"\bimport ref\b|libb2ref|b2ref_\w+\(|#include.*oracle", ln)]
                    if lines:
                        offenders.append((f, lines[:2]))
    assert not offenders, offenders
