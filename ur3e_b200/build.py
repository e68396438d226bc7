"""In-tree build of libur3e_b200.so for sm_100a (nvcc cross-compiles without a GPU).

Six kernel instantiations (f32/f64 x three model size classes) are compiled as separate translation
units in parallel, then linked with the C ABI and the MJCF loader into ur3e_b200/libur3e_b200.so.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.environ.get("UR3E_OBJ_DIR") or os.path.join(HERE, "build")
LIB = os.environ.get("UR3E_LIB_OUT") or os.path.join(HERE, "libur3e_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + ARCH
FLAGS += os.environ.get("UR3E_EXTRA_FLAGS", "").split()   # A/B experiments only (e.g. -DUR3E_REG_CHOL=0); part of the object digest
# float32 production units: approximate division / sqrt (2 ulp) and flush-to-zero; sin/cos stay exact. float64 validation units keep IEEE.
F32_FLAGS = ["-prec-div=false", "-prec-sqrt=false", "-ftz=true"]
UNITS = ["capi.cu", "mjcf.cpp"] + ["inst_%s_%s.cu" % (r, d) for r in ("f32", "f64") for d in ("raw", "grip", "main")]
# every header under csrc/ is a dependency of every kernel unit (a missed one would link a stale object silently)
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))) + [os.path.join("..", "..", "include", "ur3e_b200.h")]


def _digest(unit):
    h = hashlib.sha256(" ".join(FLAGS + F32_FLAGS).encode())
    deps = [unit] + (HEADERS if unit != "mjcf.cpp" else ["host_model.h", "xml_mini.h"])
    for f in deps:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile(unit, verbose):
    obj = os.path.join(OBJ, unit.rsplit(".", 1)[0] + ".o")
    stamp = obj + ".sha"
    dig = _digest(unit)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False, ""
    extra = F32_FLAGS if unit.startswith("inst_f32") else []
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, unit), "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (unit, p.stderr[-4000:]))
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, True, p.stderr


def build(verbose=False, jobs=None):
    os.makedirs(OBJ, exist_ok=True)
    jobs = jobs or min(len(UNITS), os.cpu_count() or 1)
    objs, rebuilt, logs = [], False, []
    with cf.ThreadPoolExecutor(jobs) as ex:
        for obj, did, log in ex.map(lambda u: _compile(u, verbose), UNITS):
            objs.append(obj); rebuilt |= did; logs.append(log)
    if rebuilt or not os.path.exists(LIB):
        p = subprocess.run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stderr[-4000:])
    if verbose:
        sys.stderr.write("\n".join(l for l in logs if l))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
