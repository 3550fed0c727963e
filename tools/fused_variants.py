"""Build timing-experiment variants of the fused kernel: only gf_fused_inst.cu is recompiled (with -DJF_FU_DBG=... or
other -D flags), the other objects come from the product build.  Output: jammy_flows_b200/variants/lib_fu_<name>.so,
selected at run time with JF_LIB_PATH (tools/ only).

    python tools/fused_variants.py nomma=-DJF_FU_DBG=1 nofp64=-DJF_FU_DBG=2 ...
"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jammy_flows_b200 import build as B
B.build()
d = os.path.join(B.PKG_DIR, "variants")
os.makedirs(d, exist_ok=True)
objs = [os.path.join(B.PKG_DIR, "build", s.replace(".cu", ".o")) for s in B.SOURCES if s != "gf_fused_inst.cu"]
procs = []
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    obj = os.path.join(d, "fu_%s.o" % name)
    cmd = [B._nvcc()] + B.NVCC_FLAGS + flags.split(",") + ["-c", os.path.join(B.CSRC, "gf_fused_inst.cu"), "-o", obj]
    procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, obj, p in procs:
    out, _ = p.communicate()
    if p.returncode != 0:
        sys.stderr.write(out)
        raise SystemExit("nvcc failed on variant " + name)
    lib = os.path.join(d, "lib_fu_%s.so" % name)
    subprocess.check_call([B._nvcc(), "-shared", "-o", lib] + objs + [obj, "-gencode", "arch=compute_100a,code=sm_100a"])
    print("built", lib)
