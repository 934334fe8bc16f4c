"""The build never shades a (triangle, voxel) pair whose voxel lies in the centre half of its clip level, off the blend band
(csrc/vgi_build.cu: mip_interior): the radiance down-sample overwrites that texel with mix(down-sample, own, 0). This pins the
claim on the CPU, against the oracle and — where the reference tree or the prebuilt shader library is present — against the
reference's own radianceDownSample.comp: garbage in exactly those texels of level l never reaches the output of the level-l
down-sample, while garbage in the band or outside the centre half does."""
import numpy as np
import pytest

from vk_voxel_cone_tracing_b200 import structs as S


def _interior_mask(cfg, regs, level):
    """(R, R, R) bool over the TEXEL coordinates (z, y, x) of `level`: centre half, every axis off the blend band."""
    R, band = cfg.resolution, cfg.downsample_band
    half = R // 2
    m = np.ones((R, R, R), dtype=bool)
    for axis, k in ((2, 0), (1, 1), (0, 2)):        # array axis of coordinate k
        pm = regs[level - 1].min_corner[k] >> 1
        g = (np.arange(R) - pm) % R
        e = g - half // 2
        dist = np.where(e >= 0, e, -e - 1)
        ok = (g < half) & (dist < half // 2 - band)
        shape = [1, 1, 1]
        shape[axis] = R
        m &= ok.reshape(shape)
    return m


def _level_view(cfg, atlas, level):
    """The R^3 x 6-face block of `level` without its borders: (face, z, y, x, 4)."""
    R, rb = cfg.resolution, cfg.resolution + 2
    blk = atlas[1:R + 1, level * rb + 1:level * rb + R + 1]
    return np.stack([blk[:, :, f * rb + 1:f * rb + R + 1] for f in range(S.VGI_FACES)], axis=0)


def _put_level(cfg, atlas, level, view):
    R, rb = cfg.resolution, cfg.resolution + 2
    for f in range(S.VGI_FACES):
        atlas[1:R + 1, level * rb + 1:level * rb + R + 1, f * rb + 1:f * rb + R + 1] = view[f]


def _check(ds, oracle, resolution, band, cam):
    rng = np.random.default_rng(resolution * 7 + band)
    cfg = S.default_config(resolution, 4, downsample_band=band)
    regs = oracle.regions(cfg, cam)
    a = rng.integers(0, 256, size=S.atlas_shape(cfg), dtype=np.uint8)
    a[rng.random(a.shape[:3]) < 0.5] = 0
    for level in range(1, cfg.level_count):
        inside = _interior_mask(cfg, regs, level)
        assert inside.sum() > 0 and (~inside).sum() > 0
        want = a.copy()
        ds(cfg, regs, level, want, 1)
        # garbage where the build leaves the injected radiance undefined: nothing changes
        v = _level_view(cfg, a, level).copy()
        v[:, inside] = rng.integers(0, 256, size=v[:, inside].shape, dtype=np.uint8)
        b = a.copy()
        _put_level(cfg, b, level, v)
        assert not np.array_equal(a, b)
        got = b.copy()
        ds(cfg, regs, level, got, 1)
        assert np.array_equal(_level_view(cfg, got, level), _level_view(cfg, want, level)), level
        # control: the same garbage one ring further out (band and beyond) does reach the output
        v = _level_view(cfg, a, level).copy()
        v[:, ~inside] = rng.integers(0, 256, size=v[:, ~inside].shape, dtype=np.uint8)
        c = a.copy()
        _put_level(cfg, c, level, v)
        got = c.copy()
        ds(cfg, regs, level, got, 1)
        assert not np.array_equal(_level_view(cfg, got, level), _level_view(cfg, want, level))


@pytest.mark.parametrize("resolution,band,cam", [(32, 3, (0.0, 0.0, 0.0)), (32, 5, (3.3, -1.2, 7.9)), (64, 10, (-40.0, 2.5, 13.0))])
def test_interior_radiance_never_survives_the_downsample_oracle(oracle, resolution, band, cam):
    _check(oracle.downsample, oracle, resolution, band, cam)


@pytest.mark.parametrize("resolution,band,cam", [(32, 3, (3.3, -1.2, 7.9)), (64, 10, (-40.0, 2.5, 13.0))])
def test_interior_radiance_never_survives_the_downsample_reference_shader(oracle, resolution, band, cam):
    from oracle import refshaders as Rf
    if not Rf.available():
        pytest.skip("reference tree not present and oracle/_ref/libvgi_refshaders.so not prebuilt")
    Rf.build()
    _check(Rf.downsample, oracle, resolution, band, cam)
