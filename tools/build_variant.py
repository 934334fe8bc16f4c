"""Development aid: build tools/_dev/libvgi_<name>.so with extra -D flags for vgi_trace.cu (A/B experiments).
Usage: build_variant.py name -DFOO=1 -DBAR=2 ; run with VGI_LIBVGI_PATH=tools/_dev/libvgi_<name>.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vk_voxel_cone_tracing_b200 import build as B  # noqa: E402

name, flags = sys.argv[1], [a for a in sys.argv[2:] if a != "--keep-main"]
if "--keep-main" not in sys.argv:      # --keep-main: do not touch csrc/libvgi.so (a queued gpurun call may snapshot the tree)
    B.build_libvgi()
dev = os.path.join(ROOT, "tools", "_dev")
os.makedirs(dev, exist_ok=True)
obj = os.path.join(dev, f"vgi_trace_{name}.o")
cmd = [B.NVCC] + B.ARCH + B.COMMON + flags + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, "vgi_trace.cu"), "-o", obj]
out = subprocess.run(cmd, capture_output=True, text=True)
for l in out.stderr.splitlines():
    if "k_trace_mainILi16" in l or "k_trace_specular" in l or "registers" in l and "trace" in l:
        print(l)
print("\n".join(l for l in out.stderr.splitlines() if "Used" in l or "spill" in l))
objs = [obj] + [os.path.join(B.CSRC, f) for f in ("vgi_build.o", "vgi_svo.o", "vgi_atlas.o", "vgi_raster.o", "vgi_post.o", "vgi_api.o")]
subprocess.check_call([B.NVCC] + B.ARCH + ["-shared", "-o", os.path.join(dev, f"libvgi_{name}.so")] + objs + ["-ccbin", B.GXX, "-lcudart"])
print("built", name)
