"""Build experiment variants of libjammy_b200.so (different -D tuning flags) into jammy_flows_b200/variants/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jammy_flows_b200 import build as B
VARIANTS = {
    "unroll2": ["-DJF_K_UNROLL=2"],
    "estrin": ["-DJF_EXP_ESTRIN=1"],
    "unroll4": ["-DJF_K_UNROLL=4"],
    "nopresolve": ["-DJF_PRESOLVE_F32=0"],
    "stop0": ["-DJF_PRE_STOP=1e-6"],
    "stop2e2": ["-DJF_PRE_STOP=2e-2"],
    "noprefetch": ["-DJF_PREFETCH_NEXT_DIM=0"],
    "minblk4": ["-DJF_GF_MIN_BLOCKS=4"],
    "minblk2": ["-DJF_GF_MIN_BLOCKS=2"],
    "exp_imm": ["-DJF_EXP_CONST=0"],
    "quirk_inline": ["-DJF_QUIRK_OUTLINE=0"],
    "pre_cvt": ["-DJF_PRE_CVT=1"],
    "pre_it2": ["-DJF_PRE_ITERS=2"],
    "pre_br32": ["-DJF_PRE_BRACKET32=1"],
    "pre_br32_cvt": ["-DJF_PRE_BRACKET32=1", "-DJF_PRE_CVT=1"],
    "estrin_unroll4": ["-DJF_EXP_ESTRIN=1", "-DJF_K_UNROLL=4"],
}
d = os.path.join(B.PKG_DIR, "variants")
os.makedirs(d, exist_ok=True)
for name in (sys.argv[1:] or VARIANTS):
    B.build(force=True, extra_flags=VARIANTS[name], out=os.path.join(d, "lib_%s.so" % name), tag=name)
