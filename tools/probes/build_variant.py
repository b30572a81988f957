"""build liboip_b200 with extra nvcc defines into opticalimageprocessor_b200/build/variants/<name>.so (kernel experiments).
usage: build_variant.py name -DOIP_DBG_VARIANT=1 ...;  run with OIP_B200_LIB=<that .so>"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from opticalimageprocessor_b200 import build as B
name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.HERE, "build", "variants"); os.makedirs(out_dir, exist_ok=True)
objs = []
if not os.environ.get('NO_BASE_BUILD'):
    B.build()
for src in B.sources():
    base = os.path.basename(src)[:-3]
    if base in os.environ.get("VARIANT_FILES", "pan_fast_c0").split(","):
        obj = os.path.join(out_dir, f"{name}_{base}.o")
        subprocess.check_call([B.NVCC] + [f for f in B.FLAGS if f not in ("-Xptxas", "-v")] + defs + ["-c", src, "-o", obj])
    else:
        obj = os.path.join(B.HERE, "build", base + ".o")
    objs.append(obj)
lib = os.path.join(out_dir, f"liboip_{name}.so")
subprocess.check_call([B.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lcufft"])
print(lib)
