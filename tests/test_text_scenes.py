"""Course text scenes (BASELINE configs 1-3).  PARITY UNPINNED: the reference at HEAD has no code for them, so
these tests check (i) the parser against the grammar the shipped files use, (ii) the primitive / shading semantics
against closed-form answers, and (iii) on the GPU, the CUDA kernel against the CPU statement of the same source
(oracle/text_oracle.cpp) on every shipped scene."""
import glob
import math
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLDEN

import rt_b200  # noqa: F401
from rt_b200 import host, textscene as T

TEXT_DIR = os.path.join(GOLDEN, "text")
ALL = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(TEXT_DIR, "*.npz")))


def fixture(name):
    return T.load_npz(os.path.join(TEXT_DIR, name + ".npz"))


def scene_of(prims, shading=T.RT_SHADE_FLAT, **kw):
    s = T.TextScene()
    s.prims = np.concatenate(prims) if prims else np.zeros(0, T.PRIM_DTYPE)
    s.shading = shading
    for k, v in kw.items():
        setattr(s, k, np.asarray(v, np.float32) if isinstance(v, (list, tuple)) else v)
    return s


# ---- parser ------------------------------------------------------------------------------------------------
def test_fixtures_cover_all_shipped_scenes():
    assert len(ALL) == 13 and "scene-000" in ALL and "practice5_2" in ALL


def test_parser_reads_the_grammar(tmp_path):
    p = tmp_path / "s.txt"
    p.write_text("DIMENSIONS 32 16\nRAY_DEPTH 3\nBG_COLOR 0.1 0.2 0.3\nAMBIENT_LIGHT 0.5 0.5 0.5\n\n"
                 "CAMERA_POSITION 1 2 3\nCAMERA_RIGHT 1 0 0\nCAMERA_UP 0 1 0\nCAMERA_FORWARD 0 0 -1\nCAMERA_FOV_X 1.0\n"
                 "NEW_LIGHT\nLIGHT_POSITION 1 1 1\nLIGHT_ATTENUATION 1 0 0.5\nLIGHT_INTENSITY 2 2 2\n"
                 "NEW_LIGHT\nLIGHT_DIRECTION 0 1 0\nLIGHT_INTENSITY 1 1 1\n"
                 "NEW_PRIMITIVE\nBOX 1 2 3\nPOSITION 4 5 6\nROTATION 0 0 0.7071068 0.7071068\nCOLOR 1 0 0\nDIELECTRIC\nIOR 1.3\n"
                 "NEW_PRIMITIVE\nTRIANGLE 0 0 0 1 0 0 0 1 0\nEMISSION 3 3 3\nMETALLIC\n")
    s = T.load_text_scene(str(p))
    assert (s.width, s.height, s.ray_depth, s.samples, s.shading) == (32, 16, 3, 1, T.RT_SHADE_WHITTED)
    assert np.allclose(s.bg_color, [0.1, 0.2, 0.3]) and np.allclose(s.cam_position, [1, 2, 3]) and s.fov_x == np.float32(1.0)
    assert list(s.lights["kind"]) == [T.RT_LIGHT_POINT, T.RT_LIGHT_DIRECTIONAL]
    assert np.allclose(s.lights["attenuation"][0], [1, 0, 0.5]) and np.allclose(s.lights["attenuation"][1], [1, 0, 0])
    b, t = s.prims
    assert b["kind"] == T.RT_PRIM_BOX and b["material"] == T.RT_MAT_DIELECTRIC and np.isclose(b["ior"], 1.3)
    assert np.allclose(b["param"][:3], [1, 2, 3]) and np.allclose(b["rotation"], [0, 0, 0.7071068, 0.7071068])
    assert t["kind"] == T.RT_PRIM_TRIANGLE and t["material"] == T.RT_MAT_METALLIC and np.allclose(t["emission"], 3)
    assert np.allclose(t["rotation"], [0, 0, 0, 1])  # defaults: identity pose
    p.write_text("DIMENSIONS 4 4\nSAMPLES 8\nNEW_PRIMITIVE\nPLANE 0 1 0\n")
    assert T.load_text_scene(str(p)).shading == T.RT_SHADE_PATH
    p.write_text("DIMENSIONS 4 4\nNEW_PRIMITIVE\nTORUS 1 2\n")
    with pytest.raises(RuntimeError, match="RT_ERR_BAD_SCENE"):
        T.load_text_scene(str(p))
    with pytest.raises(RuntimeError):
        T.load_text_scene(str(tmp_path / "missing.txt"))


def test_scene_000_fixture_is_the_shipped_file():
    s = fixture("scene-000")
    assert (s.width, s.height, s.shading) == (640, 480, T.RT_SHADE_FLAT)
    assert list(s.prims["kind"]) == [T.RT_PRIM_ELLIPSOID, T.RT_PRIM_PLANE, T.RT_PRIM_BOX]
    assert np.allclose(s.prims["rotation"][2], [0.31246, 0.15623, 0.15623, 0.92388])


# ---- primitives: closed-form known answers -----------------------------------------------------------------
def test_plane_sphere_ellipsoid_known_answers():
    s = scene_of([T.TextScene.new_prim(T.RT_PRIM_PLANE, [0, 1, 0]),
                  T.TextScene.new_prim(T.RT_PRIM_ELLIPSOID, [2, 2, 2], position=[-1, 1, -5]),
                  T.TextScene.new_prim(T.RT_PRIM_ELLIPSOID, [1, 2, 3], position=[20, 0, 0])])
    t, prim, n, inside = O.text_closest(s, [0, 1.5, 0], [0, -1, 0])
    assert prim == 0 and t == pytest.approx(1.5) and np.allclose(n, [0, 1, 0]) and not inside
    t, prim, n, inside = O.text_closest(s, [0, -2, 0], [0, 1, 0])  # from below: normal faces the ray
    assert prim == 0 and t == pytest.approx(2.0) and np.allclose(n, [0, -1, 0]) and inside
    c = np.array([-1, 1, -5], np.float32)
    o = np.array([0, 1.5, 0], np.float32)
    d = (c - o) / np.linalg.norm(c - o)
    t, prim, n, inside = O.text_closest(s, o, d)
    assert prim == 1 and t == pytest.approx(np.linalg.norm(c - o) - 2, rel=1e-5) and np.allclose(n, -d, atol=1e-5)
    t, prim, n, inside = O.text_closest(s, c, [0, 0, 1])  # from the centre outwards
    assert prim == 1 and t == pytest.approx(2.0) and inside and np.allclose(n, [0, 0, -1], atol=1e-6)
    for axis, r in enumerate([1, 2, 3]):  # semi-axes of the ellipsoid at x = 20
        o = np.array([20, 0, 0], np.float32)
        o[axis] += 10
        d = np.zeros(3, np.float32)
        d[axis] = -1
        t, prim, n, _ = O.text_closest(s, o, d)
        assert prim == 2 and t == pytest.approx(10 - r, rel=1e-5) and n[axis] == pytest.approx(1.0)
    t, prim, _, _ = O.text_closest(s, [0, 5, 0], [1, 0, 0])  # parallel to the plane, above everything
    assert prim == -1 and math.isinf(t)


def test_box_rotation_and_triangle_known_answers():
    h = math.sqrt(0.5)
    s = scene_of([T.TextScene.new_prim(T.RT_PRIM_BOX, [1, 2, 3]),
                  T.TextScene.new_prim(T.RT_PRIM_BOX, [1, 2, 3], position=[0, 0, -20], rotation=[0, 0, h, h]),  # 90 deg about z
                  T.TextScene.new_prim(T.RT_PRIM_TRIANGLE, [0, 0, 0, 1, 0, 0, 0, 1, 0], position=[0, 0, 30])])
    t, prim, n, _ = O.text_closest(s, [5, 0, 0], [-1, 0, 0])
    assert prim == 0 and t == pytest.approx(4.0) and np.allclose(n, [1, 0, 0])
    t, prim, n, _ = O.text_closest(s, [0, 0, 10], [0, 0, -1])
    assert prim == 0 and t == pytest.approx(7.0) and np.allclose(n, [0, 0, 1])
    t, prim, n, inside = O.text_closest(s, [0, 0, 0], [0, 1, 0])
    assert prim == 0 and t == pytest.approx(2.0) and inside and np.allclose(n, [0, -1, 0])
    # rotated by +90 deg about z: the local x half-extent (1) now lies along world y, the y half-extent (2) along x
    t, prim, n, _ = O.text_closest(s, [5, 0, -20], [-1, 0, 0])
    assert prim == 1 and t == pytest.approx(3.0, rel=1e-5) and np.allclose(n, [1, 0, 0], atol=1e-6)
    t, prim, n, _ = O.text_closest(s, [0, 5, -20], [0, -1, 0])
    assert prim == 1 and t == pytest.approx(4.0, rel=1e-5) and np.allclose(n, [0, 1, 0], atol=1e-6)
    t, prim, n, _ = O.text_closest(s, [0.25, 0.25, 35], [0, 0, -1])
    assert prim == 2 and t == pytest.approx(5.0)
    assert O.text_closest(s, [0.75, 0.75, 35], [0, 0, -1])[1] == 0  # beta + gamma > 1: misses the triangle, goes on to the box


def test_flat_shading_of_scene_000():
    s = fixture("scene-000")
    ids = O.text_ids(s, s.width, s.height)
    img = O.text_render(s, s.width, s.height, 1)
    assert set(np.unique(ids)) == {-1, 0, 1, 2}
    assert np.array_equal(img[ids == -1], np.tile(s.bg_color, (int((ids == -1).sum()), 1)))
    for k in range(3):
        assert np.array_equal(img[ids == k], np.tile(s.prims["color"][k], (int((ids == k).sum()), 1)))
    assert (ids[: s.height // 2 - 1] != 1).all() and (ids[-1] == 1).all()  # the ground plane fills the bottom
    assert ids[int(s.height * 0.62), int(s.width * 0.40)] == 0  # the sphere left of the centre, the box to the right
    assert (ids[:, s.width // 2:] == 2).any() and not (ids[:, : s.width // 3] == 2).any()


# ---- Whitted shading ---------------------------------------------------------------------------------------
def _whitted_scene(prims, lights, ambient=(0.1, 0.1, 0.1), depth=4, bg=(0.2, 0.3, 0.4)):
    s = scene_of(prims, T.RT_SHADE_WHITTED, ray_depth=depth)
    s.ambient = np.asarray(ambient, np.float32)
    s.bg_color = np.asarray(bg, np.float32)
    s.lights = np.zeros(len(lights), T.LIGHT_DTYPE)
    for i, (kind, inten, vec, att) in enumerate(lights):
        s.lights[i] = (kind, inten, vec, att)
    s.cam_position = np.array([0, 5, 0], np.float32)
    s.cam_forward = np.array([0, -1, 0], np.float32)  # looking straight down
    s.cam_up = np.array([0, 0, -1], np.float32)
    s.fov_x = np.float32(0.2)
    return s


def test_whitted_lambert_shadow_and_attenuation():
    plane = T.TextScene.new_prim(T.RT_PRIM_PLANE, [0, 1, 0], color=[0.5, 0.25, 1.0])
    sun = (T.RT_LIGHT_DIRECTIONAL, [2, 2, 2], [0, 1, 1], [1, 0, 0])  # 45 degrees
    s = _whitted_scene([plane], [sun])
    img = O.text_render(s, 9, 9, 1)
    expect = np.array([0.5, 0.25, 1.0]) * (0.1 + 2 * math.cos(math.pi / 4))
    assert np.allclose(img, expect, rtol=1e-5)
    # a sphere between the light and the centre of the view puts the centre pixel in shadow: ambient only
    blocker = T.TextScene.new_prim(T.RT_PRIM_ELLIPSOID, [0.3, 0.3, 0.3], position=[0, 2, 2], color=[1, 1, 1])
    img = O.text_render(_whitted_scene([plane, blocker], [sun]), 9, 9, 1)
    assert np.allclose(img[4, 4], np.array([0.5, 0.25, 1.0]) * 0.1, rtol=1e-5)
    # point light straight above the centre at height 2: I / (c0 + c1 r + c2 r^2)
    lamp = (T.RT_LIGHT_POINT, [3, 3, 3], [0, 2, 0], [1, 0.5, 0.25])
    img = O.text_render(_whitted_scene([plane], [lamp], ambient=(0, 0, 0)), 9, 9, 1)
    assert np.allclose(img[4, 4], np.array([0.5, 0.25, 1.0]) * 3 / (1 + 0.5 * 2 + 0.25 * 4), rtol=1e-3)


def test_whitted_mirror_and_dielectric():
    bg = np.array([0.2, 0.3, 0.4])
    mirror = T.TextScene.new_prim(T.RT_PRIM_PLANE, [0, 1, 0], color=[0.9, 0.8, 0.7], material=T.RT_MAT_METALLIC)
    img = O.text_render(_whitted_scene([mirror], []), 5, 5, 1)
    assert np.allclose(img, bg * [0.9, 0.8, 0.7], rtol=1e-6)  # the mirror shows the tinted sky
    assert not O.text_render(_whitted_scene([mirror], [], depth=1), 5, 5, 1).any()  # depth 1: the reflected ray is never cast
    # a slab of IOR-1 "glass": no reflection at any angle, refraction straight through, tinted once on entry
    glass = T.TextScene.new_prim(T.RT_PRIM_BOX, [50, 0.5, 50], position=[0, 2, 0], color=[0.5, 1, 1],
                                 material=T.RT_MAT_DIELECTRIC, ior=1.0)
    floor = T.TextScene.new_prim(T.RT_PRIM_PLANE, [0, 1, 0], color=[1, 1, 1])
    img = O.text_render(_whitted_scene([glass, floor], [], ambient=(1, 1, 1)), 5, 5, 1)
    assert np.allclose(img[2, 2], [0.5, 1, 1], rtol=1e-5)
    # real glass at normal incidence: R0 = ((1 - n) / (1 + n))^2 on both faces
    glass["ior"] = 1.5
    img = O.text_render(_whitted_scene([glass, floor], [], ambient=(1, 1, 1), bg=(0, 0, 0), depth=3), 5, 5, 1)
    r0 = ((1 - 1.5) / (1 + 1.5)) ** 2
    assert img[2, 2, 1] == pytest.approx((1 - r0) ** 2, rel=1e-4)  # depth 3: enter, leave, floor; reflections see black sky


# ---- path tracing ------------------------------------------------------------------------------------------
def test_path_furnace_convex_object_has_zero_variance():
    """A diffuse convex body in a uniform sky: every cosine-sampled bounce escapes, weight = albedo exactly."""
    ball = T.TextScene.new_prim(T.RT_PRIM_ELLIPSOID, [1, 1, 1], position=[0, 0, -4], color=[0.25, 0.5, 0.75])
    s = scene_of([ball], T.RT_SHADE_PATH, ray_depth=6, samples=4)
    s.bg_color = np.ones(3, np.float32)
    img = O.text_render(s, 33, 33, 4, seed=3)
    assert np.allclose(img[16, 16], [0.25, 0.5, 0.75], rtol=2e-5)
    assert np.allclose(img[0, 0], 1.0)


def test_light_sampler_matches_its_pdf():
    """E[1 / pdf] over directions drawn from the emitter sampler = the solid angle the emitter subtends."""
    ball = T.TextScene.new_prim(T.RT_PRIM_ELLIPSOID, [1, 1, 1], position=[0, 0, -4], emission=[5, 5, 5])
    omega = O.text_emitter_solid_angle(scene_of([ball], T.RT_SHADE_PATH), [0, 0, 0], 200000)
    assert omega == pytest.approx(2 * math.pi * (1 - math.sqrt(1 - 1 / 16)), rel=0.01)
    a, b, dist = 1.0, 0.5, 3.0  # a 2a x 2b face seen on axis from distance `dist`
    slab = T.TextScene.new_prim(T.RT_PRIM_BOX, [a, b, 0.25], position=[0, 0, -dist - 0.25], emission=[1, 1, 1])
    omega = O.text_emitter_solid_angle(scene_of([slab], T.RT_SHADE_PATH), [0, 0, 0], 200000)
    assert omega == pytest.approx(4 * math.asin(a * b / math.sqrt((a * a + dist * dist) * (b * b + dist * dist))), rel=0.01)
    tri = T.TextScene.new_prim(T.RT_PRIM_TRIANGLE, [-1, -1, -2, 1, -1, -2, -1, 1, -2], emission=[1, 1, 1])
    omega = O.text_emitter_solid_angle(scene_of([tri], T.RT_SHADE_PATH), [0, 0, 0], 200000)
    # half of the 2 x 2 square at distance 2 (the square's diagonal splits its on-axis solid angle evenly)
    assert omega == pytest.approx(0.5 * 4 * math.asin(1 / (1 + 4)), rel=0.01)


def test_path_light_sampling_is_unbiased():
    """Same scene with the emitter found by cosine sampling only (emission moved to a plane, which is never
    area-sampled) and by the cosine/light mixture: the two estimators agree."""
    floor = T.TextScene.new_prim(T.RT_PRIM_PLANE, [0, 1, 0], color=[0.8, 0.8, 0.8])
    lamp = T.TextScene.new_prim(T.RT_PRIM_BOX, [1.0, 0.05, 1.0], position=[0, 3, -6], emission=[8, 8, 8])
    s = scene_of([floor, lamp], T.RT_SHADE_PATH, ray_depth=3, samples=1)
    s.cam_position = np.array([0, 1, 0], np.float32)
    mixture = O.text_render(s, 24, 16, 4096, seed=1)
    s.cam_position = np.array([0, 1, 0], np.float32)
    ceiling_light = T.TextScene.new_prim(T.RT_PRIM_TRIANGLE, [0, 0, 0, 0, 0, 0, 0, 0, 0])  # degenerate: never hit
    ceiling_light["emission"] = 0
    s2 = scene_of([floor, lamp.copy(), ceiling_light], T.RT_SHADE_PATH, ray_depth=3, samples=1)
    s2.prims["emission"][1] = 0  # the lamp no longer counts as an emitter ...
    s2.cam_position = np.array([0, 1, 0], np.float32)
    dark = O.text_render(s2, 24, 16, 64, seed=1)
    assert dark[12:].max() == 0.0  # ... and without emission the floor is black (black sky)
    floor_px = mixture[12:]
    assert floor_px.mean() > 0.05 and np.isfinite(mixture).all()
    again = O.text_render(s, 24, 16, 4096, seed=2)
    assert abs(again[12:].mean() - floor_px.mean()) < 0.03 * floor_px.mean()  # converged, seed-independent


def test_sample_split_adds_up_on_cpu():
    s = fixture("practice3_2")
    full = O.text_render(s, 40, 30, 8, seed=5)
    a = O.text_render(s, 40, 30, 8, seed=5, sample_begin=0, sample_end=3)
    b = O.text_render(s, 40, 30, 8, seed=5, sample_begin=3, sample_end=8)
    assert np.allclose(a + b, full, rtol=1e-5, atol=1e-6)


# ---- GPU vs the CPU statement of the same source -----------------------------------------------------------
@pytest.fixture(scope="module")
def rt():
    from rt_b200 import gpu

    g = gpu.RtGpu(1, 0)
    yield g
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL)
def test_gpu_matches_cpu_statement(name, rt):
    s = fixture(name)
    w, h = (s.width // 4, s.height // 4) if s.shading != T.RT_SHADE_FLAT else (s.width, s.height)
    spp = min(s.samples, 16)
    rt.upload_text_scene(s)
    ids = rt.primary_ids(w, h)
    ref_ids = O.text_ids(s, w, h)
    assert (ids == ref_ids).mean() >= 0.999
    rt.render(w, h, spp, seed=11)
    img, st = rt.readback()
    ref = O.text_render(s, w, h, spp, seed=11)
    assert st["samples"] == w * h * spp and np.isfinite(img).all()
    rel = (np.abs(img - ref) / (np.abs(ref) + 1e-3)).max(axis=2)
    # identical Philox keys -> identical paths except where an FMA-contraction-level difference flips a discrete
    # decision (silhouette pixels, total-internal-reflection thresholds)
    assert (rel > 1e-3).mean() <= (0.001 if s.shading == T.RT_SHADE_FLAT else 0.02), (rel > 1e-3).mean()
    assert abs(img.mean() - ref.mean()) <= 5e-3 * ref.mean() + 1e-6


@pytest.mark.gpu
def test_gpu_text_sample_split_tonemap_and_errors(rt):
    from rt_b200 import gpu

    s = fixture("practice3_5")
    rt.upload_text_scene(s)
    rt.render(64, 64, 12, seed=2)
    full, _ = rt.readback()
    rt.render(64, 64, 12, seed=2, sample_begin=0, sample_end=5)
    rt.render(64, 64, 12, seed=2, sample_begin=5, sample_end=12, accumulate=True)
    split, _ = rt.readback()
    assert np.allclose(split, full, rtol=2e-6, atol=1e-7)
    dev = rt.readback_rgb8().astype(int)
    assert np.abs(dev - host.tonemap_rgb8(split).astype(int)).max() <= 1
    bad = fixture("scene-001")
    bad.ray_depth = 99
    with pytest.raises(gpu.RtGpuError, match="RT_ERR_BAD_SCENE"):
        rt.upload_text_scene(bad)


@pytest.mark.gpu
def test_gpu_text_full_size_configs(rt):
    """BASELINE config 1 (scene-000 at its own size) and config 3 size (1920 x 1080) on a path-traced scene."""
    s = fixture("scene-000")
    rt.upload_text_scene(s)
    ids = rt.primary_ids(s.width, s.height)
    assert np.array_equal(ids, O.text_ids(s, s.width, s.height)) or (ids == O.text_ids(s, s.width, s.height)).mean() > 0.9999
    p = fixture("practice3_3")
    rt.upload_text_scene(p)
    rt.render(1920, 1080, 4, seed=1)
    img, st = rt.readback()
    assert st["samples"] == 1920 * 1080 * 4 and np.isfinite(img).all() and img.mean() > 0.01
