"""Development aid (CPU only): work statistics of the cone marches on the bench scene, counted by the oracle on a sample of
the image rows - how many steps and level samples the diffuse / specular cones execute per pixel, how many of them are exactly
zero, and how often a sample lands in a NEW (level, base cell) compared with the previous step of the same cone (the reuse a
kernel could exploit). Usage: python tools/trace_model.py [--res 256] [--rows 24]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--rows", type=int, default=24, help="image rows sampled (spread over the height)")
    a = ap.parse_args()
    from oracle import pyoracle as O
    from tests import common
    from vk_voxel_cone_tracing_b200 import structs as S
    O.build()
    inp = common.atrium_inputs(a.res, 4096 if a.res >= 256 else 1024, a.width, a.height)
    cfg = inp["cfg"]
    regs = O.regions(cfg, inp["cam_pos"])
    osc = O.OracleScene(inp["scene"])
    _, rad, _ = O.build_clipmap(cfg, regs, osc, inp["light"], inp["shadow"], inp["shadow_depth"], 0)
    gb = inp["gbuffer"]
    hg = O.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    prm = S.default_vct_params(regs[0], cfg.resolution, 8)
    out = (C.c_uint64 * 10)()
    O.lib().vgo_debug_cell_stats(C.c_int(1), out)
    pixels = 0
    for y in np.linspace(0, a.height - 1, a.rows).astype(int):
        O.cone_trace(cfg, inp["cam"], hg, prm, inp["light"], inp["shadow"], inp["shadow_depth"], rad, rows=(int(y), int(y) + 1))
        pixels += int((gb["depth"][y] < 1.0).sum())
    O.lib().vgo_debug_cell_stats(C.c_int(0), out)
    st = np.array(list(out), dtype=np.float64).reshape(2, 5)
    spec_px = 0
    for y in np.linspace(0, a.height - 1, a.rows).astype(int):
        spec_px += int(((gb["specular"][y, :, :3].max(axis=-1) > 0) & (gb["specular"][y, :, 3] > 0) & (gb["depth"][y] < 1.0)).sum())
    print(f"scene: atrium, R = {a.res}, {a.width}x{a.height}, {a.rows} rows sampled, {pixels} covered pixels, {spec_px} with a specular cone")
    for w, name, n in ((0, "diffuse cones", pixels), (1, "specular cone", max(spec_px, 1))):
        steps, samples, new_cell, zero_s, zero_steps = st[w]
        print(f"{name}: {steps / n:8.1f} steps / pixel, {samples / n:8.1f} level samples / pixel; "
              f"all-zero samples {100 * zero_s / max(samples, 1):5.1f} %, all-zero steps {100 * zero_steps / max(steps, 1):5.1f} %; "
              f"samples landing in a new (level, cell) {100 * new_cell / max(samples, 1):5.1f} %  "
              f"(= {samples / max(new_cell, 1):.1f} consecutive samples per cell)")


if __name__ == "__main__":
    main()
