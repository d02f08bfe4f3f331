"""Build libkocr.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libkocr.so")
SOURCES = ["kocr_host.cu", "kocr_preprocess.cu", "kocr_elementwise.cu", "kocr_gemm.cu", "kocr_attention.cu",
           "kocr_tower.cu", "kocr_handoff.cu", "kocr_png.cu"]
HEADERS = ["kocr_common.cuh", "kocr_kernels.h", "kocr_inflate_core.h", os.path.join("..", "..", "include", "kocr.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_variant(name, defines):
    """Experiment helper: build libkocr_<name>.so with extra -D flags (select it at run time with KOCR_LIB=<path>)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.path.join(HERE, f"libkocr_{name}.so")
    d = os.path.join(HERE, "build", name)
    os.makedirs(d, exist_ok=True)
    objs, procs = [], []
    for f in SOURCES:
        o = os.path.join(d, f.replace(".cu", ".o"))
        objs.append(o)
        procs.append(subprocess.Popen([nvcc] + NVCC_FLAGS + [f"-D{x}" for x in defines] + ["-c", os.path.join(CSRC, f), "-o", o],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out_txt, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(out_txt)
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL"])
    return out


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for f in SOURCES:
        o = os.path.join(HERE, "build", f.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, f), "-o", o]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for f, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {f}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
