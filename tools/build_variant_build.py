"""Development aid: like tools/build_variant.py but for vgi_build.cu (strict numerics: -fmad=false).
Usage: build_variant_build.py name -DFOO=1 ; run with VGI_LIBVGI_PATH=tools/_dev/libvgi_<name>.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vk_voxel_cone_tracing_b200 import build as B  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
B.build_libvgi()
dev = os.path.join(ROOT, "tools", "_dev")
os.makedirs(dev, exist_ok=True)
obj = os.path.join(dev, f"vgi_build_{name}.o")
cmd = [B.NVCC] + B.ARCH + B.COMMON + flags + ["-fmad=false", "-Xptxas", "-v", "-c", os.path.join(B.CSRC, "vgi_build.cu"), "-o", obj]
out = subprocess.run(cmd, capture_output=True, text=True)
if out.returncode:
    print(out.stderr[-3000:])
    sys.exit(1)
lines = out.stderr.splitlines()
for i, l in enumerate(lines):
    if "k_inject" in l and "Compiling" in l:
        print("\n".join(lines[i:i + 4]))
objs = [obj] + [os.path.join(B.CSRC, f) for f in ("vgi_trace.o", "vgi_svo.o", "vgi_atlas.o", "vgi_raster.o", "vgi_post.o", "vgi_api.o")]
subprocess.check_call([B.NVCC] + B.ARCH + ["-shared", "-o", os.path.join(dev, f"libvgi_{name}.so")] + objs + ["-ccbin", B.GXX, "-lcudart"])
print("built", name)
