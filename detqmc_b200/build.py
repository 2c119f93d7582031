"""Build libdqmc_b200.so in-tree with nvcc for sm_100a (no torch headers needed: the library is a
plain C-ABI shared object on top of the CUDA runtime)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdqmc_b200.so")
SOURCES = ["context.cu", "cb_kernels.cu", "gemm_kernels.cu", "qr_kernels.cu", "update_kernels.cu",
           "misc_kernels.cu", "hubbard_kernels.cu", "rng_stream.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "128,550"] + os.environ.get("DQMC_NVCC_EXTRA", "").split()


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dqmc_gpu.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on " + src)
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-ldl"]
    subprocess.check_call(cmd)
    return OUT


def build_kbench():
    """Development tool: kernel micro-benchmarks linked against the object files."""
    build(force=False)
    objdir = os.path.join(HERE, "build")
    objs = [os.path.join(objdir, os.path.splitext(f)[0] + ".o") for f in SOURCES
            if os.path.exists(os.path.join(CSRC, f))]
    out = os.path.join(HERE, "..", "tools", "kbench")
    subprocess.check_call(["nvcc"] + NVCC_FLAGS + [os.path.join(HERE, "..", "tools", "kbench.cu")] + objs + ["-o", out])
    return out


if __name__ == "__main__":
    if "--kbench" in sys.argv:
        print(build_kbench())
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
