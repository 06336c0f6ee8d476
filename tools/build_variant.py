"""Bottleneck probes: rebuilds ONE kernel source with -D flags and links a variant library
next to the objects of the regular build (build/variants/libepb_<name>.so).  The variants
compute garbage on purpose (a pipeline role is switched off); they only exist to be timed:

    python tools/build_variant.py skipA conv_tc.cu -DEPB_DBG_SKIP_A
    EPB_LIB_PATH=build/variants/libepb_skipA.so python tools/conv_table.py
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarpose_b200 import build as B

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
out = os.path.join(ROOT, "build", "variants")
os.makedirs(out, exist_ok=True)
obj = os.path.join(out, "%s_%s.o" % (src[:-3], name))
subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + B.PER_FILE.get(src, []) + flags +
                      ["-c", os.path.join(B.CSRC, src), "-o", obj])
objs = [obj if s == src else os.path.join(B.OBJ, s[:-3] + ".o") for s in B.sources()]
lib = os.path.join(out, "libepb_%s.so" % name)
subprocess.check_call([B._nvcc(), "-shared", "-o", lib] + objs + ["-lcudart", "-lcuda"])
print(lib)
