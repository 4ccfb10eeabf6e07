"""Experiment builds: recompile only the 256-thread encode variant with extra -D flags and link it with the
other objects of the current tree.  usage: python tools/xbuild.py tag [-DFOO ...]  ->  flo_b200/libflo_b200_<tag>.so"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flo_b200 import build as B
tag = sys.argv[1]
defs = [a for a in sys.argv[2:] if a.startswith("-D")]
extra = [a for a in sys.argv[2:] if not a.startswith("-D")]
objs = []
for src in B.SOURCES:
    base = os.path.join(B.CSRC, src.replace(".cu", ".o"))
    if src == "flo_encode_nt256.cu":
        obj = os.path.join(B.CSRC, src.replace(".cu", f"_{tag}.o"))
        cmd = [B.nvcc(), *B.NVCC_FLAGS, *defs, *extra, "-Xptxas=-v", "-c", os.path.join(B.CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        lines = (r.stdout + r.stderr).splitlines()
        for i, l in enumerate(lines):
            if "Compiling entry" in l and "k_encode_frames" in l:
                print(l.split("'")[1][-40:], "|", lines[i + 1].strip(), "|", lines[i + 2].strip())
        if r.returncode:
            print("\n".join(lines[-30:])); sys.exit(1)
        objs.append(obj)
    else:
        if not os.path.exists(base):
            sys.exit(f"{base} missing: run python -m flo_b200.build first")
        objs.append(base)
out = os.path.join(ROOT, "flo_b200", f"libflo_b200_{tag}.so")
subprocess.check_call([B.nvcc(), "-shared", "-o", out, *objs, "-cudart", "static"])
print(out)
