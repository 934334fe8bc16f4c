"""Q3 quantified: how far is the canonical conservative coverage of this repo from the
reference's own raster coverage — (R+2)-pixel viewports in which a pixel spans TWO voxels, device-max MSAA, sample shading
0.25 — as MODELLED by oracle/vgi_oracle_literal.inc? CPU only (oracle).

Per level: occupancy IoU, the share of the model's voxels that the canonical set contains (conservative => close to 1),
and the share of the canonical voxels the model misses (the holes a 2-voxel pixel leaves). Then both clipmaps are traced with
the same tracer (the oracle's voxelConeTracing.frag restatement) and the images compared.
Variants = what Vulkan leaves to the driver: sample count, where an invocation interpolates. Prints one JSON object.

  python tools/coverage_report.py [--scene cornell|atrium] [--res 64] [--levels 3] [--size 192] [--out profiles/...json]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def occupancy(cfg, atlas, level):
    R = cfg.resolution
    blk = atlas[1:R + 1, 1 + (R + 2) * level:1 + (R + 2) * level + R, 1:R + 1, 0]     # face 0, raw opacity channel r
    return blk > 0


def build(O, cfg, regs, osc, light, shadow, depth, variant):
    op, rad = O.new_atlas(cfg), O.new_atlas(cfg)
    L = cfg.level_count
    if variant is None:
        O.build_clipmap(cfg, regs, osc, light, shadow, depth, 0, op, rad)
        return op, rad
    for l in range(L):
        O.literal_voxelize_level(cfg, regs, l, osc, op, **variant)
    for l in range(1, L):
        O.downsample(cfg, regs, l, op, 0)
    O.wrap_border(cfg, op, False)
    for l in range(L):
        O.literal_inject_level(cfg, regs, l, osc, light, shadow, depth, rad, **variant)
    for l in range(L):
        O.copy_alpha(cfg, l, rad, op)
    for l in range(1, L):
        O.downsample(cfg, regs, l, rad, 1)
    O.wrap_border(cfg, rad, False)
    return op, rad


def report(scene_name="cornell", res=64, levels=3, size=192, cam_pos=None):
    from oracle import pyoracle as O
    from vk_voxel_cone_tracing_b200 import raster, structs as S, synth
    O.build()
    if scene_name == "cornell":
        scene = synth.cornell_box()
        light, shadow = synth.make_light(origin=(0.0, 20.0, -3.5))
        cam_pos = cam_pos or (0.0, 0.0, 0.0)
        cam = synth.make_camera(cam_pos, (0.0, 0.0, -1.0), aspect=1.0)
        sm = 1024
    else:
        scene = synth.atrium()
        light, shadow = synth.make_light()
        cam_pos = cam_pos or (-8.0, 3.0, 0.0)
        cam = synth.make_camera(cam_pos, (1.0, 0.0, 0.0), aspect=1.0)
        sm = 2048
    cfg = S.default_config(res, levels)
    depth = raster.shadow_depth(scene, shadow, sm)
    gb = raster.gbuffer(scene, cam, size, size)
    regs = O.regions(cfg, cam_pos)
    osc = O.OracleScene(scene)
    hg = O.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
    prm = S.default_vct_params(regs[0], cfg.resolution, 8)
    cov = gb["depth"] < 1.0

    def trace(rad):
        d, s, _ = O.cone_trace(cfg, cam, hg, prm, light, shadow, depth, rad)
        return d, s

    op_c, rad_c = build(O, cfg, regs, osc, light, shadow, depth, None)
    img_c = trace(rad_c)
    variants = {"8x_first_covered_sample": dict(samples=8, shade_at=0), "8x_pixel_centre": dict(samples=8, shade_at=1),
                "4x_first_covered_sample": dict(samples=4, shade_at=0), "1x_pixel_centre": dict(samples=1, shade_at=1)}
    out = {"scene": scene_name, "resolution": res, "levels": levels, "image": [size, size], "camera": list(cam_pos),
           "canonical": "exact conservative triangle/voxel overlap (Schwarz-Seidel), one contribution per (triangle, voxel)",
           "model": "oracle/vgi_oracle_literal.inc (see its header for what is restated and what is assumed)",
           "variants": {}}
    for name, v in variants.items():
        op_l, rad_l = build(O, cfg, regs, osc, light, shadow, depth, dict(v, q2_fixed=True))
        per_level = []
        for l in range(levels):
            a, b = occupancy(cfg, op_c, l), occupancy(cfg, op_l, l)
            inter, union = int((a & b).sum()), int((a | b).sum())
            per_level.append({"level": l, "canonical_voxels": int(a.sum()), "model_voxels": int(b.sum()),
                              "iou": inter / max(union, 1), "model_inside_canonical": inter / max(int(b.sum()), 1),
                              "canonical_missed_by_model": 1.0 - inter / max(int(a.sum()), 1)})
        img_l = trace(rad_l)
        stats = {}
        for k, nm in ((0, "diffuse"), (1, "specular")):
            diff = (img_c[k] - img_l[k])[cov][:, :3]
            mse = float(np.mean(diff.astype(np.float64) ** 2))
            stats[nm] = {"max_abs": float(np.abs(diff).max()), "mean_abs": float(np.abs(diff).mean()),
                         "psnr_db": float(10.0 * np.log10(1.0 / mse)) if mse > 0 else float("inf"),
                         "pixels_over_1e-3": float((np.abs(diff).max(axis=1) > 1e-3).mean())}
        out["variants"][name] = {"occupancy": per_level, "gi_image_vs_canonical": stats}
    # the driver-dependent spread: two models against each other
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cornell")
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--size", type=int, default=192)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    r = report(a.scene, a.res, a.levels, a.size)
    text = json.dumps(r, indent=1)
    if a.out:
        open(a.out, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
