#!/usr/bin/env python3
"""oracle/build_ref.py -- TEST INFRASTRUCTURE: build the CPU oracle.

Compiles the UNMODIFIED reference sources where they lie (/root/reference/Box2D/**/*.cpp, 50 translation
units, plain g++ -- the reference's own premake build is not used) together with oracle/ref_harness.cpp
and oracle/b2o_math.c into

    oracle/_ref/libb2ref.so         reference + harness, sinf/cosf/sincosf interposed by b2o_sincosf
    oracle/_ref/libb2ref_stock.so   same, but with the stock glibc libm (pins the oracle against the
                                    golden trajectory hashes of SURVEY.md 8c, which were taken with glibc)

No reference source is copied into the repository; only these binaries are produced, and oracle/_ref/ is
git-ignored.  When /root/reference is absent (the GPU box) the prebuilt files are used as they are.

    python oracle/build_ref.py [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("B2_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")

CXXFLAGS = ["-std=c++11", "-O2", "-DNDEBUG", "-fPIC", "-w", "-I" + REF]


def _run(cmd):
    subprocess.run(cmd, check=True)


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def reference_available():
    return os.path.isdir(os.path.join(REF, "Box2D"))


def lib_path(stock=False):
    return os.path.join(OUT, "libb2ref_stock.so" if stock else "libb2ref.so")


MT_CAP = 32


def mt_lib_path():
    """The reference with its compile-time thread cap raised (b2_maxThreads, Box2D/Common/b2Settings.h:165, is 8 as
    shipped): the second CPU row of bench.py, for hosts with more than 8 cores."""
    return os.path.join(OUT, "libb2ref_mt%d.so" % MT_CAP)


def build_mt(force=False):
    """libb2ref_mt32.so: the reference compiled from a scratch copy under the system temp directory in which the one line
    `#define b2_maxThreads 8` reads 32 (a power of two, as the merge tree of b2ThreadDataSorter.h:327-376 needs; the
    survey verified 16 with an unchanged trajectory hash).  Stock libm: this library is only ever timed.  Nothing of
    the copy enters the repository."""
    import shutil
    import tempfile
    harness = [os.path.join(HERE, f) for f in ("ref_harness.cpp", "ref_harness.h", "b2o_math.c", "b2o_math.h")]
    harness.append(os.path.join(ROOT, "include", "b2cuda.h"))
    if not reference_available():
        return mt_lib_path() if os.path.exists(mt_lib_path()) else None
    if not force and _newer(mt_lib_path(), harness):
        return mt_lib_path()
    tmp = tempfile.mkdtemp(prefix="b2ref_mt_")
    try:
        shutil.copytree(os.path.join(REF, "Box2D"), os.path.join(tmp, "Box2D"))
        settings = os.path.join(tmp, "Box2D", "Common", "b2Settings.h")
        text = open(settings).read()
        import re
        patched = re.sub(r"(#define\s+b2_maxThreads\s+)8\b", lambda m: m.group(1) + str(MT_CAP), text)
        if patched == text:
            raise RuntimeError("b2_maxThreads not found in b2Settings.h")
        os.chmod(settings, 0o644)
        open(settings, "w").write(patched)
        obj = os.path.join(tmp, "obj")
        os.makedirs(obj)
        flags = ["-std=c++11", "-O2", "-DNDEBUG", "-fPIC", "-w", "-I" + tmp]
        jobs, objects = [], []
        for d, _, files in os.walk(os.path.join(tmp, "Box2D")):
            for f in sorted(files):
                if f.endswith(".cpp"):
                    src = os.path.join(d, f)
                    o = os.path.join(obj, os.path.relpath(src, tmp).replace("/", "_")[:-4] + ".o")
                    objects.append(o)
                    jobs.append(["g++"] + flags + ["-c", src, "-o", o])
        math_o = os.path.join(obj, "b2o_math.o")
        jobs.append(["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-c", os.path.join(HERE, "b2o_math.c"), "-o", math_o])
        h_o = os.path.join(obj, "ref_harness.o")
        jobs.append(["g++"] + flags + ["-I" + os.path.join(ROOT, "include"), "-I" + HERE, "-fno-access-control",
                                       "-DB2REF_STOCK_LIBM", "-c", os.path.join(HERE, "ref_harness.cpp"), "-o", h_o])
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            list(ex.map(_run, jobs))
        os.makedirs(OUT, exist_ok=True)
        _run(["g++", "-shared", "-o", mt_lib_path(), h_o, math_o] + sorted(objects) + ["-lpthread", "-lm"])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return mt_lib_path()


def build(force=False):
    """Build both oracle libraries; returns the path of libb2ref.so. No-op if up to date or no reference."""
    harness = [os.path.join(HERE, f) for f in ("ref_harness.cpp", "ref_harness.h", "b2o_math.c", "b2o_math.h")]
    harness.append(os.path.join(ROOT, "include", "b2cuda.h"))
    if not reference_available():
        if os.path.exists(lib_path()):
            return lib_path()
        raise RuntimeError("reference tree %s is absent and oracle/_ref/libb2ref.so was not prebuilt" % REF)
    if not force and _newer(lib_path(), harness) and _newer(lib_path(True), harness):
        return lib_path()

    os.makedirs(OBJ, exist_ok=True)
    sources = []
    for d, _, files in os.walk(os.path.join(REF, "Box2D")):
        for f in sorted(files):
            if f.endswith(".cpp"):
                sources.append(os.path.join(d, f))
    sources.sort()

    jobs = []
    objects = []
    for s in sources:
        o = os.path.join(OBJ, os.path.relpath(s, REF).replace("/", "_")[:-4] + ".o")
        objects.append(o)
        if force or not _newer(o, [s]):
            jobs.append(["g++"] + CXXFLAGS + ["-c", s, "-o", o])

    math_o = os.path.join(OBJ, "b2o_math.o")
    jobs.append(["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-c", os.path.join(HERE, "b2o_math.c"), "-o", math_o])
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + HERE]
    h_o = os.path.join(OBJ, "ref_harness.o")
    hs_o = os.path.join(OBJ, "ref_harness_stock.o")
    jobs.append(["g++"] + CXXFLAGS + inc + ["-fno-access-control", "-fno-builtin-sinf", "-fno-builtin-cosf",
                                            "-c", os.path.join(HERE, "ref_harness.cpp"), "-o", h_o])
    jobs.append(["g++"] + CXXFLAGS + inc + ["-fno-access-control", "-DB2REF_STOCK_LIBM",
                                            "-c", os.path.join(HERE, "ref_harness.cpp"), "-o", hs_o])

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(_run, jobs))

    _run(["g++", "-shared", "-o", lib_path(), h_o, math_o] + objects +
         ["-Wl,-Bsymbolic-functions", "-lpthread", "-lm"])
    _run(["g++", "-shared", "-o", lib_path(True), hs_o, math_o] + objects + ["-lpthread", "-lm"])
    return lib_path()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_mt(force="--force" in sys.argv))
