"""Build libsnb.so in-tree with nvcc for sm_100a (no JIT cache, no CPU fallback).

  python safe-interactive-crowdnav_b200/build.py [--force] [--verbose]

crowd_kernels.cu is compiled with -fmad=false (ORCA must be bit-identical to float32 RVO2: every float op is a
single IEEE operation); the denoiser translation units keep FMA contraction.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "snb", "_lib")
OUT = os.path.join(OUT_DIR, "libsnb.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
          "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("SNB_NVCC_FLAGS", "").split()

UNITS = [
    ("snb_api.cu", []),
    ("crowd_kernels.cu", ["-fmad=false", "-Xcompiler", "-ffp-contract=off"]),
    ("scene_kernels.cu", ["-fmad=false", "-Xcompiler", "-ffp-contract=off"]),
    ("jmid_kernels.cu", []),
    ("jmid_gemm.cu", []),
    ("jmid_attn.cu", []),
    ("jmid_attn2.cu", []),
    ("jmid_fp32x.cu", []),
    ("jmid_api.cu", []),
    ("pred_prep.cu", ["-fmad=false", "-Xcompiler", "-ffp-contract=off"]),
    ("pred_encode.cu", []),
    ("pred_post.cu", []),
    ("pred_api.cu", []),
    ("rollout_kernels.cu", ["-fmad=false"]),
]


def _newer(src, dst):
    return not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
              [os.path.join(ROOT, "include", "snb.h")]
    objs, procs = [], []
    for name, extra in UNITS:
        src = os.path.join(CSRC, name)
        if not os.path.exists(src):
            continue
        obj = os.path.join(objdir, name.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(src, obj) or any(_newer(h, obj) for h in headers):
            cmd = [NVCC] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"---- nvcc {name} ----\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or not os.path.exists(OUT):
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
