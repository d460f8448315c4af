"""A/B helper (development only): builds variants of libgnnis.so with extra -D flags into _ab/<name>.so.
usage: python tools/ab_build.py name1:-DFOO=1,-DBAR=2 name2: ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnnimplicitsolvent_b200 import build as B  # noqa: E402

B.build_library()
objdir = os.path.join(B.CSRC, "build")
os.makedirs(os.path.join(ROOT, "_ab"), exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    vdir = os.path.join(ROOT, "_ab", name + "_obj")
    os.makedirs(vdir, exist_ok=True)
    objs = []
    cmds = []
    for s in B.SOURCES:
        if s in ("edge_mlp_tc.cu", "node_mlp_tc.cu") and flags:
            o = os.path.join(vdir, s.replace(".cu", ".o"))
            cmds.append([B._nvcc(), *B.ARCH, *B.FLAGS, *flags, "-c", os.path.join(B.CSRC, s), "-o", o])
        else:
            o = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(o)
    procs.append((name, [subprocess.Popen(c) for c in cmds], objs))
for name, ps, objs in procs:
    for p in ps:
        assert p.wait() == 0, name
    out = os.path.join(ROOT, "_ab", name + ".so")
    subprocess.check_call([B._nvcc(), *B.ARCH, "-shared", "-o", out, *objs, "-lcudart"])
    print("built", out)
