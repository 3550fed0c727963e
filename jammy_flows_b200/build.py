"""Build recipe of libjammy_b200.so (the C-ABI library, sm_100a only) -- plain nvcc, no torch headers.

    python -m jammy_flows_b200.build            # (re)build if sources are newer than the library

The library is built IN-TREE (jammy_flows_b200/libjammy_b200.so) so that it travels to the GPU box with the repo
snapshot; it is git-ignored.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libjammy_b200.so")
SOURCES = ["api.cu", "gf_inst_f64_logpdf.cu", "gf_inst_f64_sample.cu", "gf_inst_f32_logpdf.cu", "gf_inst_f32_sample.cu", "gfx_inst.cu", "gf_fused_inst.cu", "mlp_bwd_inst.cu", "gf_fb_inst.cu", "gf_sbwd_inst.cu", "gf_fwd_inst.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libjammy_b200.so cannot be built (there is no CPU fallback)")
    return nvcc


def _newest_source_mtime():
    newest = 0.0
    for d in (CSRC, os.path.join(REPO_ROOT, "include")):
        for f in os.listdir(d):
            newest = max(newest, os.path.getmtime(os.path.join(d, f)))
    return newest


def needs_build():
    return (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < _newest_source_mtime()


def build(force=False, verbose=True, extra_flags=(), out=None, tag=""):
    """Build the library.  `extra_flags` / `out` / `tag` build an experiment variant next to the product library
    (selected at run time with JF_LIB_PATH, tools/ only)."""
    if out is None and not force and not needs_build():
        return LIB_PATH
    lib_path = out or LIB_PATH
    objs = []
    build_dir = os.path.join(PKG_DIR, "build" + (("_" + tag) if tag else ""))
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        text, _ = p.communicate()
        log.append(text)
        if p.returncode != 0:
            sys.stderr.write(text)
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [_nvcc(), "-shared", "-o", lib_path] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc link failed")
    with open(os.path.join(build_dir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        spills = [l for l in "\n".join(log).splitlines() if "spill" in l and "0 bytes spill stores, 0 bytes spill loads" not in l]
        print("built %s (%d kernels report spills; see %s)" % (lib_path, len(spills), os.path.join(build_dir, "ptxas.log")))
    return lib_path


if __name__ == "__main__":
    build(force="--force" in sys.argv)
