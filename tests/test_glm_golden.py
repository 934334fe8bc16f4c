"""Host-side matrix maths vs golden vectors generated from the reference's own vendored glm
(oracle/ref_glm/gen_glm_golden.cpp, compiled against /root/reference/Dependencies in place).
Bit-exact: these matrices feed quantised results (region tests, shadow-map texel addresses)."""
import json
import os

import numpy as np

from vk_voxel_cone_tracing_b200 import glm, synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "glm_golden.json")))


def _bits(m):
    return np.asarray(m, dtype=np.float32).reshape(-1).view(np.uint32).tolist()


def _close_bits(got, want, ulps=0):
    a = np.asarray(got, dtype=np.uint32).view(np.float32)
    b = np.asarray(want, dtype=np.uint32).view(np.float32)
    if ulps == 0:
        # -0.0 == 0.0 is fine (sign of zero never reaches a quantised result)
        return np.array_equal(a, b)
    return np.allclose(a, b, rtol=ulps * 1.2e-7, atol=1e-30)


def test_camera_ubo_matches_reference_glm():
    for key in ("camera_cfg1_cornell", "camera_cfg2_atrium", "camera_oblique"):
        g = GOLD[key]
        vp, vpi, eye = glm.camera_ubo(g["position"], g["direction"], aspect=np.float32(g["aspect"]))
        assert _close_bits(_bits(vp), g["view_proj"]), key
        assert _close_bits(_bits(vpi), g["view_proj_inv"]), key


def test_light_shadow_desc_matches_reference_glm():
    for key in ("light_reference_default", "light_cfg1_cornell"):
        g = GOLD[key]
        view, proj, d = glm.light_shadow_desc(g["origin"], g["direction"])
        assert _bits(d) == g["direction_normalized"], key
        assert _close_bits(_bits(view), g["view"]), key
        assert _close_bits(_bits(proj), g["proj"]), key
        light, sh = synth.make_light(origin=g["origin"], direction=g["direction"])
        assert _close_bits(_bits(np.array(sh.view[:], dtype=np.float32)), g["view"])
        assert _close_bits(_bits(np.array(sh.proj[:], dtype=np.float32)), g["proj"])


def test_voxelizer_view_proj_matches_reference_glm():
    """Voxelizer::setViewProjection (Voxelizer.cpp:298-327): the canonical voxelizer does not
    rasterise through these matrices, but the restated glm must still reproduce them."""
    for key in ("voxelizer_level0_r128", "voxelizer_level3_r256_moved"):
        g = GOLD[key]
        vs = np.float32(g["voxel_size"])
        region = (np.float32(g["resolution"]) * vs).astype(np.float32)
        mc = (np.asarray(g["min_corner"], dtype=np.float32) * vs).astype(np.float32)
        eye = (mc + np.array([0, 0, region], dtype=np.float32)).astype(np.float32)
        f32 = np.float32
        vx = glm.mul(glm.ortho(-region, region, -region, region, f32(0.1), region),
                     glm.look_at(eye, eye + np.array([1, 0, 0], dtype=np.float32), (0, 1, 0)))
        vy = glm.mul(glm.ortho(-region, region, -region, region, f32(0.1), region),
                     glm.look_at(eye, eye + np.array([0, 1, 0], dtype=np.float32), (0, 0, -1)))
        vz = glm.mul(glm.ortho(-region, region, -region, region, f32(0.1), region),
                     glm.look_at(mc, mc + np.array([0, 0, 1], dtype=np.float32), (0, 1, 0)))
        assert _close_bits(_bits(vx), g["view_proj_x"]), key
        assert _close_bits(_bits(vy), g["view_proj_y"]), key
        assert _close_bits(_bits(vz), g["view_proj_z"]), key
