"""patches/*.patch — the Vulkan-side change set of INTEGRATION.md — against the reference checkout:
  1. every patch applies cleanly, in order, to a scratch copy of /root/reference (git apply --check, then apply);
  2. tools/make_patches.py regenerates exactly the committed patches (they are not hand-edited);
  3. the translation units of the patched tree that touch libvgi or the new export paths pass `g++ -fsyntax-only
     -DVFS_USE_VGI` against the reference's OWN headers (Buffer.h, Semaphore.h, Device.h, Image.h, Camera.h,
     DirectionalLight.h, GLTFScene.h, RenderPassManager.h, ...) and include/vgi.h. This image has no Vulkan SDK, so
     tests/vk_stub/ supplies declarations of the Vulkan / VMA / GLFW / tinygltf names those files use; glm is the
     reference's vendored copy. What this proves: the patch set is consistent with the reference's class interfaces and
     with the C ABI; what it cannot prove here: that it links and runs (no loader, no ICD)."""
import glob
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VGI_REFERENCE_ROOT", "/root/reference")
PATCHES = sorted(glob.glob(os.path.join(ROOT, "patches", "*.patch")))

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "VFS")), reason="reference checkout not present")


def _git(cwd, *a, check=True):
    return subprocess.run(["git", "-c", "user.email=t@localhost", "-c", "user.name=t", "-c", "core.autocrlf=false", *a],
                          cwd=cwd, check=check, capture_output=True, text=True)


@pytest.fixture(scope="module")
def patched_tree(tmp_path_factory):
    tree = str(tmp_path_factory.mktemp("ref"))
    for d in ("VFS", "VulkanFramework", "Common"):
        shutil.copytree(os.path.join(REF, d), os.path.join(tree, d))
    _git(tree, "init", "-q", ".")
    _git(tree, "add", "-A")
    _git(tree, "commit", "-qm", "reference")
    for p in PATCHES:
        chk = _git(tree, "apply", "--check", p, check=False)
        assert chk.returncode == 0, f"{os.path.basename(p)} does not apply:\n{chk.stderr}"
        _git(tree, "apply", p)
        _git(tree, "add", "-A")
        _git(tree, "commit", "-qm", os.path.basename(p))
    return tree


def test_patch_set_is_present():
    assert len(PATCHES) == 7, PATCHES


@needs_ref
def test_patches_apply_in_order(patched_tree):
    log = _git(patched_tree, "log", "--oneline").stdout.strip().splitlines()
    assert len(log) == 1 + len(PATCHES)
    for f in ("VFS/RenderPass/Clipmap/VgiBridge.h", "VFS/RenderPass/Clipmap/VgiBridge.cpp", "VFS/GLTFSceneVgi.cpp"):
        assert os.path.exists(os.path.join(patched_tree, f))
    # nothing outside the hot path's seam is touched
    touched = set(_git(patched_tree, "diff", "--name-only", "HEAD~%d" % len(PATCHES), "HEAD").stdout.split())
    assert touched == {
        "VulkanFramework/Device.cpp", "VulkanFramework/Buffers/Buffer.h", "VulkanFramework/Buffers/Buffer.cpp",
        "VulkanFramework/Sync/Semaphore.h", "VulkanFramework/Sync/Semaphore.cpp",
        "VFS/RenderPass/Clipmap/VgiBridge.h", "VFS/RenderPass/Clipmap/VgiBridge.cpp", "VFS/GLTFSceneVgi.cpp",
        "VFS/Camera.h", "VFS/GLTFScene.h", "VFS/GLTFScene.cpp", "VFS/RenderPass/GBufferPass.cpp", "VFS/DirectionalLight.cpp",
        "VFS/RenderPass/Clipmap/VoxelizationPass.cpp", "VFS/RenderPass/Clipmap/RadianceInjectionPass.cpp",
        "VFS/RenderPass/Clipmap/VoxelConeTracingPass.cpp", "VFS/Application.cpp", "VFS/Application.h"}, touched


@needs_ref
def test_committed_patches_are_what_the_generator_writes(tmp_path):
    keep = {p: open(p, newline="").read() for p in PATCHES}
    try:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_patches.py"), "--reference", REF], check=True,
                       capture_output=True)
        for p, text in keep.items():
            assert open(p, newline="").read() == text, f"{os.path.basename(p)} is stale: run tools/make_patches.py"
    finally:
        for p, text in keep.items():
            open(p, "w", newline="").write(text)


UNITS = ["VulkanFramework/Buffers/Buffer.cpp", "VulkanFramework/Sync/Semaphore.cpp",
         "VFS/RenderPass/Clipmap/VgiBridge.cpp", "VFS/GLTFSceneVgi.cpp"]


@needs_ref
@pytest.mark.parametrize("unit", UNITS)
def test_patched_units_compile_against_the_reference_headers(patched_tree, unit):
    inc = ["-I", os.path.join(ROOT, "tests", "vk_stub"), "-I", patched_tree, "-I", os.path.join(patched_tree, "VFS"),
           "-I", os.path.join(REF, "Dependencies"), "-I", os.path.join(ROOT, "include")]
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-w", "-DVFS_USE_VGI", "-DGLM_ENABLE_EXPERIMENTAL", *inc,
           os.path.join(patched_tree, unit)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-6000:]


# The remaining edited files pull in far more of Vulkan / ImGui than the stand-in headers declare, so they cannot pass
# -fsyntax-only as a whole here. What can be checked: none of the compiler's diagnostics points at a line the patch set
# ADDED (an unknown member, a wrong argument list or a missing include in our hunks would).
EDITED = ["VFS/RenderPass/Clipmap/VoxelizationPass.cpp", "VFS/RenderPass/Clipmap/RadianceInjectionPass.cpp",
          "VFS/RenderPass/Clipmap/VoxelConeTracingPass.cpp", "VFS/GLTFScene.cpp", "VFS/RenderPass/GBufferPass.cpp",
          "VFS/DirectionalLight.cpp", "VFS/Application.cpp", "VulkanFramework/Device.cpp"]


@needs_ref
@pytest.mark.parametrize("unit", EDITED)
def test_no_diagnostic_on_an_added_line(patched_tree, unit):
    import re
    added = set()
    cur = None
    for l in _git(patched_tree, "diff", "-U0", "HEAD~%d" % len(PATCHES), "HEAD", "--", unit).stdout.splitlines():
        m = re.match(r"@@ -\d+(?:,\d+)? \+(\d+)(?:,(\d+))? @@", l)
        if m:
            start, n = int(m.group(1)), int(m.group(2) or 1)
            added.update(range(start, start + n))
    assert added, unit
    inc = ["-I", os.path.join(ROOT, "tests", "vk_stub"), "-I", patched_tree, "-I", os.path.join(patched_tree, "VFS"),
           "-I", os.path.join(REF, "Dependencies"), "-I", os.path.join(ROOT, "include")]
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-fmax-errors=0", "-DVFS_USE_VGI", "-DGLM_ENABLE_EXPERIMENTAL",
                        *inc, os.path.join(patched_tree, unit)], capture_output=True, text=True)
    hits = []
    for l in r.stderr.splitlines():
        m = re.match(r"(.+?):(\d+):\d+: (?:fatal )?error", l)
        if m and os.path.abspath(m.group(1)) == os.path.abspath(os.path.join(patched_tree, unit)) and int(m.group(2)) in added:
            hits.append(l)
    assert not hits, "\n".join(hits[:20])
