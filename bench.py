#!/usr/bin/env python
"""bench.py — the voxel-GI hot path on BASELINE.json's configs[1].

A "step" is one GI frame of the reference's pass sequence for this path (Application.cpp:178,192,221):
clipmap build at frame 0 (all six levels voxelized, injected, down-sampled — the worst case of the
cadence) + 1080p 16-cone CombinedGI cone trace (rendering mode 8, diffuse + specular cones).

  value  frames/s with every input resident in HBM (device-timed with CUDA events, max over ranks)
  e2e    frames/s through the C-ABI host-buffer call vgi_frame_host(): per step the G-buffer and the
         shadow map are copied from pinned host memory, both output images are copied back
  stages build_ms (the "voxelize+inject+mip ms @256^3" half of the metric) and trace_ms / trace_fps
  roofline (dominant kernel) / roofline_hbm_kernel / roofline_stage / roofline_cone_trace / cpu_baseline: see DESIGN.md "measurement"

N > 1 (torchrun, one rank per GPU): every rank renders its own camera view of the same scene
(batched views, sharded by view, no data-path collective) — weak scaling.

--impl reference: the reference's own shaders compiled for the CPU (oracle/_ref/libvgi_refshaders.so, see
oracle/glsl_shim/) on the host cores, on a bounded sample of the same frame; the fixed-function triangle coverage
and the injection loop come from the CPU oracle (cpu_baseline.port_share). Without the library: the oracle port alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# keep stdout for the one JSON line: NCCL's own banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

METRIC = "1080p 16-cone GI frames/s; voxelize+inject+mip ms @256^3 (1/2/4/8 B200)"
WORKLOAD = ("configs[1]: Sponza-scale synthetic atrium (262144 tris), 6-level 256^3 clipmap voxelize+inject+mip "
            "(frame 0: all levels) + 1920x1080 16-cone diffuse/specular GI (mode 8), one frame per step along a fixed "
            "8-camera path")
RES, LEVELS, SHADOW, WIDTH, HEIGHT = 256, 6, 4096, 1920, 1080
TEXTURED = False               # --textured: the same mesh with textured materials (synth.textured_atrium), beside the metric
TRACE_ROW_STRIDE = 32          # reference arm: every 32nd row block of the image per sample


# ---------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def kernel_source_hash():
    """sha256 (16 hex digits) over the CUDA / C++ sources libvgi.so is built from: identifies the kernel build that an
    ncu capture under profiles/ belongs to (tools/ncu_traffic.py stores the same value beside the byte counts)."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "vk_voxel_cone_tracing_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        if f.endswith((".cu", ".cuh", ".cpp", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu capture
    profiles/r2_ncu_traffic.json (written by tools/ncu_traffic.py). The capture names the source hash of the kernels it
    profiled: a capture of another build is not this run's traffic, so None is returned instead of a stale constant."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if (RES, WIDTH, HEIGHT) != (256, 1920, 1080) or TEXTURED:   # the capture is of the headline workload
        return None
    try:
        d = json.load(open(p))
        if d.get("source_hash") != kernel_source_hash():
            return None
        return d["kernels"][kernel]["traffic"]
    except Exception:
        return None


def config_dict(world, triangles=262144):
    """The `config` object of the JSON line: identical for both arms (the reference arm runs the same workload on the CPU)."""
    return {"workload": WORKLOAD, "resolution": RES, "levels": LEVELS, "triangles": int(triangles),
            "image": [WIDTH, HEIGHT], "cones": 16, "shadow_map": SHADOW,
            "parallelism": f"{world} GPU(s); every step renders the next camera of a fixed {N_VIEWS}-view path (rank r starts at "
                           f"view r, so all ranks do the same work over a cycle), clipmap rebuilt for every frame",
            "l2": "256 MB flush between timed steps"}


# The camera path is FIXED (it does not depend on the number of ranks): view 0 is the camera of round 1's headline frame,
# views 1..7 orbit the atrium. A step renders ONE frame: step i of rank r uses view (r + i) mod 8, so every rank — at
# N = 1, 2, 4, 8 alike — cycles through the same eight cameras and does the same work over a cycle: the scaling curve
# measures the machine, not the cameras (measured on 8 GPUs with one fixed view per rank the views cost 3.6 .. 5.8 ms and
# the slowest one set the "efficiency"). The clip regions move with the camera, so every frame is a full rebuild.
N_VIEWS = 8


def view_camera(v):
    v = int(v) % N_VIEWS
    if v == 0:
        return (-8.0, 3.0, 0.0), (1.0, 0.0, 0.0)
    ang = 2.0 * np.pi * v / N_VIEWS
    cam_pos = (float(-8.0 * np.cos(ang)), 3.0 + 0.25 * v, float(5.0 * np.sin(ang)))
    d = np.array([-cam_pos[0], 1.0 - cam_pos[1] * 0.3, -cam_pos[2]])
    return cam_pos, tuple((d / np.linalg.norm(d)).tolist())


def make_inputs(rank, world):
    """Synthetic inputs of configs[1]; rank r renders view r of the fixed list."""
    from vk_voxel_cone_tracing_b200 import raster, structs as S, synth
    scene = synth.textured_atrium() if TEXTURED else synth.atrium()
    textures = synth.procedural_textures() if TEXTURED else None
    cfg = S.default_config(RES, LEVELS)
    light, shadow = synth.make_light()
    depth = raster.shadow_depth(scene, shadow, SHADOW)
    cam_pos, cam_dir = view_camera(rank)
    cam = synth.make_camera(cam_pos, cam_dir, aspect=WIDTH / HEIGHT)
    gb = raster.gbuffer(scene, cam, WIDTH, HEIGHT, textures=textures)
    return dict(scene=scene, cfg=cfg, light=light, shadow=shadow, shadow_depth=depth, cam=cam, gbuffer=gb,
                cam_pos=cam_pos, view=rank % N_VIEWS, textures=textures)


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle on a bounded sample of the same frame
# ---------------------------------------------------------------------------------------------------
class CpuFrameSampler:
    """One sample = the clipmap build of ONE level (cycling over the six levels: clear, voxelize,
    inject, copy-alpha, both down-samples) + the cone trace of every 32nd row. frame seconds =
    sum over levels of the mean level-build time + 32 x the mean row-sample time."""

    def __init__(self, inp):
        from oracle import pyoracle as O
        from vk_voxel_cone_tracing_b200 import structs as S
        O.build()
        self.threads = O.set_threads(cpu_cores())   # torchrun exports OMP_NUM_THREADS=1: use every host core
        self.O, self.inp = O, inp
        if inp.get("textures"):
            O.set_textures(inp["textures"])     # the oracle's coverage (alpha test) and injection read them like the shaders
        self.cfg = inp["cfg"]
        self.regs = O.regions(self.cfg, inp["cam_pos"])
        self.osc = O.OracleScene(inp["scene"])
        self.op = O.new_atlas(self.cfg)
        self.rad = O.new_atlas(self.cfg)
        gb = inp["gbuffer"]
        self.hg = O.HostGBuffer(gb["diffuse"], gb["normal"], gb["specular"], gb["emission"], gb["depth"])
        self.prm = S.default_vct_params(self.regs[0], self.cfg.resolution, 8)
        # the reference's own shaders compiled for the CPU (oracle/_ref/libvgi_refshaders.so, built where
        # /root/reference exists and shipped with the snapshot): used for every stage they cover
        self.Rf = None
        try:
            from oracle import refshaders as Rf
            if Rf.available():
                Rf.lib()
                self.Rf = Rf
        except Exception as e:  # the port below still gives a CPU number
            print(f"bench.py: reference-shader library unavailable ({e}); CPU arm falls back to the oracle port", file=sys.stderr)
        self.kind = "reference" if self.Rf is not None else "port"
        self.port_s = 0.0       # seconds of recorded samples spent in oracle-port code (scaled to the frame like the rest)
        self.level_s = {l: [] for l in range(self.cfg.level_count)}
        self.trace_s = []
        self.taps = []          # tri-linear taps of the sampled rows, scaled to the frame (SURVEY 8d A_cone)
        self.taps_diffuse = []  # ... of the diffuse cones alone (what k_trace_main marches)
        self.rows = []          # (y0, y1, diffuse rows, specular rows) of every traced sample: the parity check's reference
        self.i = 0

    def build_level(self, l):
        """Returns (seconds, seconds of it spent in oracle-port code)."""
        O, cfg, inp = self.O, self.cfg, self.inp
        P = self.Rf if self.Rf is not None else O     # the programmable stages the reference's shaders cover
        R = cfg.resolution
        t = time.perf_counter()
        # this level's share of the clears (the reference clears whole levels: RadianceInjectionPass.cpp:77 passes corner 0)
        P.clear_region(cfg, self.op, (0, 0, 0), (R, R, R), l)
        P.clear_region(cfg, self.rad, (0, 0, 0), (R, R, R), l)
        tp = time.perf_counter()
        # triangle coverage is fixed-function in the reference, and its injection average depends on the order the
        # hardware retires fragments: both come from the oracle (canonical coverage, exact mean)
        O.voxelize_level(cfg, self.regs, l, self.osc, self.op)
        O.inject_level(cfg, self.regs, l, self.osc, inp["light"], inp["shadow"], inp["shadow_depth"], self.rad)
        port = time.perf_counter() - tp
        if l > 0:
            P.downsample(cfg, self.regs, l, self.op, 0)
        P.copy_alpha(cfg, l, self.rad, self.op)
        if l > 0:
            P.downsample(cfg, self.regs, l, self.rad, 1)
        dt = time.perf_counter() - t
        return dt, (port if self.Rf is not None else dt)

    def trace_rows(self, k):
        O, inp = self.O, self.inp
        rows = HEIGHT // TRACE_ROW_STRIDE
        y0 = (k % TRACE_ROW_STRIDE) * rows
        args = (self.cfg, inp["cam"], self.hg, self.prm, inp["light"], inp["shadow"], inp["shadow_depth"], self.rad)
        if self.Rf is not None:     # timed: voxelConeTracing.frag itself; the oracle run after it only counts the taps
            t = time.perf_counter()
            d, s, _ = self.Rf.cone_trace(*args, rows=(y0, y0 + rows))
            dt = time.perf_counter() - t
            _, _, taps = O.cone_trace(*args, rows=(y0, y0 + rows))
        else:
            t = time.perf_counter()
            d, s, taps = O.cone_trace(*args, rows=(y0, y0 + rows))
            dt = time.perf_counter() - t
        self.rows.append((y0, y0 + rows, d[y0:y0 + rows].copy(), s[y0:y0 + rows].copy()))
        self.taps.append(taps * (HEIGHT / rows))
        self.taps_diffuse.append((taps - O.last_specular_taps()) * (HEIGHT / rows))
        return dt * (HEIGHT / rows)

    def prime(self):
        """Fill the atlases once (untimed) so the trace samples march through real radiance; the border texels are
        wrapped on both sides of both atlases (the canonical semantics of DESIGN.md section 2, Q4 / Q5), which is what
        vgi_export_atlas produces and what REPEAT filtering through the GPU's border-free store is equivalent to."""
        for l in range(self.cfg.level_count):
            self.build_level(l)
        P = self.Rf if self.Rf is not None else self.O
        P.wrap_border(self.cfg, self.op, literal=False)
        P.wrap_border(self.cfg, self.rad, literal=False)

    def parity(self, gpu_opacity, gpu_radiance, gpu_diffuse, gpu_specular, depth):
        """The headline frame against the CPU arm's own results: both 6 x 256^3 atlases byte for byte, and every traced
        row sample of both output images (max abs error, PSNR over the covered pixels; bars of north_star)."""
        from tests.common import psnr
        differ = int(np.count_nonzero(gpu_opacity != self.op)) + int(np.count_nonzero(gpu_radiance != self.rad))
        err, ref_all, gpu_all, rows = 0.0, [], [], 0
        for y0, y1, d, s in self.rows:
            cov = depth[y0:y1] < 1.0
            for ref, gpu in ((d, gpu_diffuse[y0:y1]), (s, gpu_specular[y0:y1])):
                if cov.any():
                    err = max(err, float(np.abs(gpu[cov] - ref[cov]).max()))
                ref_all.append(ref[cov])
                gpu_all.append(gpu[cov])
            rows += y1 - y0
        db = psnr(np.concatenate(gpu_all), np.concatenate(ref_all)) if ref_all else float("nan")
        return {"against": "reference GLSL compiled for the CPU (oracle/_ref)" if self.Rf is not None else "oracle port",
                "atlas_bytes_compared": int(self.op.size + self.rad.size), "atlas_bytes_differ": differ,
                "trace_rows": rows, "trace_max_abs": err, "psnr_db": float(db),
                "bars": {"atlas_bytes_differ": 0, "trace_max_abs": 1e-3, "psnr_db": 50.0},
                "ok": bool(differ == 0 and err <= 1e-3 and db >= 50.0)}

    def step(self, record=True):
        l = self.i % self.cfg.level_count
        tb, tport = self.build_level(l)
        tt = self.trace_rows(self.i * 7 + 3)
        self.i += 1
        if record:
            self.level_s[l].append(tb)
            self.trace_s.append(tt)
            self.port_s += tport if self.Rf is not None else tport + tt

    def port_share(self):
        """Fraction of the reported frame time that ran oracle-port code rather than the reference's shaders."""
        total = sum(sum(v) for v in self.level_s.values()) + sum(self.trace_s)
        return float(self.port_s / total) if total > 0 else float("nan")

    def describe(self):
        return dict(kind=self.kind, cores=self.threads, sample=self.SAMPLE if self.Rf is None else self.SAMPLE_REF,
                    port_share=self.port_share())

    def frame_seconds(self):
        means = [np.mean(v) for v in self.level_s.values() if v]
        build = float(np.sum(means)) * (self.cfg.level_count / max(len(means), 1))
        trace = float(np.mean(self.trace_s)) if self.trace_s else float("nan")
        return build, trace

    SAMPLE_REF = ("per step: clipmap build of ONE level (cycling 0..5) + cone trace of 1/32 of the 1080 rows, running the "
                  "reference's own GLSL compiled for the CPU (oracle/_ref/libvgi_refshaders.so: clipmapCleaning, copyAlphaImage, "
                  "opacity/radianceDownSample .comp and voxelConeTracing.frag); triangle coverage (fixed-function in the reference) "
                  "and the injection loop come from the oracle port (port_share = their share of the time); frame time = sum of "
                  "per-level means + 32 x mean row-sample time; sampled on view 0 of the camera path")
    SAMPLE = ("per step: oracle clipmap build of ONE level (cycling 0..5: clear, voxelize, inject, copy-alpha, "
              "opacity+radiance down-sample) + cone trace of 1/32 of the 1080 rows; frame time = sum of per-level "
              "means + 32 x mean row-sample time; sampled on view 0 of the camera path")


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inp = make_inputs(0, 1)
    smp = CpuFrameSampler(inp)
    smp.prime()
    for _ in range(args.warmup):
        smp.step(record=False)
    smp.i = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        smp.step()
    wall = time.perf_counter() - t0
    build_s, trace_s = smp.frame_seconds()
    fps = 1.0 / (build_s + trace_s)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (build_s + trace_s),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (u8 texels)",
        "data": "synthetic", "config": config_dict(args.gpus, inp["scene"].triangle_count),
        "stages": {"build_ms": 1e3 * build_s, "trace_ms": 1e3 * trace_s, "sample_wall_s": wall},
        "cpu_baseline": {"value": fps, "unit": "frames/s", **smp.describe()},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# algorithmic bytes per kernel launch (DESIGN.md "kernels"): what the kernel must move at minimum
# ---------------------------------------------------------------------------------------------------
def algorithmic_bytes(name, st, cfg, launches_per_step=1):
    """DESIGN.md section 4. P pairs, T triangles, V occupied voxels, S shadow-map edge."""
    R, L = cfg.resolution, cfg.level_count
    nvox = R ** 3
    V, T = st.occupied_voxels, st.triangles
    P = st.shaded_pairs     # the pairs on the work list (clip_pairs counts the unlisted ones, too: occupancy only)
    table = {
        "k_voxelize": T * 48 + P * 8,
        # compulsory bytes: pairs + triangle data + every shadow texel once + read-modify-write of the 96-byte accumulator rows
        "k_inject": P * 8 + T * 96 + 4 * SHADOW * SHADOW + 2 * 96 * V,
        "k_level_masks": 3 * 4 * (nvox // 32),                       # per level launch
        "k_level_records": 32 * 3 * (2 * V) // max(L, 1),            # per level launch: ~2V visited, own + ~2 children
        "k_brick_mask": 25 * 4 * (nvox // 32) // 16 * L,
        "k_scan_final": 2 * 4 * L * (nvox // 32),
        "k_scan_block_sums": 4 * L * (nvox // 32),
        "k_zero_acc": 96 * V,
    }
    return table.get(name)


def a_vim(cfg, T, Vtx):
    """bytes of the dense formulation voxelize+inject+mip replaces:
    L(12T + 32V) + 4 S^2 + 2 * 4 * 6 L (R+2)^3."""
    R, L = cfg.resolution, cfg.level_count
    return L * (12 * T + 32 * Vtx) + 4 * SHADOW * SHADOW + 2 * 4 * 6 * L * (R + 2) ** 3


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_vgi(args):
    import torch
    import torch.distributed as dist
    from vk_voxel_cone_tracing_b200.api import VoxelGI

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    parity_failed = False
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libvgi has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    inp = make_inputs(rank, world)
    gi = VoxelGI(inp["cfg"], device=local)
    if inp.get("textures"):
        gi.set_textures(inp["textures"])
    gi.set_scene(inp["scene"])
    gi.set_light(inp["light"], inp["shadow"], inp["shadow_depth"])
    gi.update_regions(inp["cam_pos"])
    dgb = gi.upload_gbuffer(inp["gbuffer"])
    prm = gi.default_vct_params(8)
    out = (torch.zeros((HEIGHT, WIDTH, 4), dtype=torch.float32, device=dev),
           torch.zeros((HEIGHT, WIDTH, 4), dtype=torch.float32, device=dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the camera path: G-buffers rasterised on the device (bit-identical to the host rasteriser: tests/test_gpu_raster.py and
    # `device_rendered_inputs_equal_host_rendered` below), resident before the timed region like the uploaded one was
    from vk_voxel_cone_tracing_b200 import synth as _synth
    views = []
    for v in range(N_VIEWS):
        pos, dirv = view_camera(v)
        cam_v = _synth.make_camera(pos, dirv, aspect=WIDTH / HEIGHT)
        views.append((pos, cam_v, gi.render_gbuffer(cam_v, WIDTH, HEIGHT)))
    for k in dgb:
        assert torch.equal(dgb[k], views[rank % N_VIEWS][2][k]), f"device-rendered G-buffer differs from the host's ({k})"
    step_no = [0]

    def next_view():
        v = views[(rank + step_no[0]) % N_VIEWS]
        step_no[0] += 1
        return v

    def frame():        # view 0: the frame the per-kernel pass, the roofline (oracle-counted taps) and the parity check use
        out[0].zero_()  # pixels the view does not cover are left untouched by the tracer: no leftovers of other views
        out[1].zero_()
        gi.update_regions(views[0][0])
        gi.build_clipmap(0)
        gi.cone_trace(views[0][1], views[0][2], prm, out=out)

    # ---- device-resident timing
    for _ in range(max(args.warmup, 3)):
        pos, cam_v, g_v = next_view()
        gi.update_regions(pos)
        gi.build_clipmap(0)
        gi.cone_trace(cam_v, g_v, prm, out=out)
    barrier()
    l0 = gi.stats().kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    for a, m, b in ev:
        pos, cam_v, g_v = next_view()
        flush.zero_()                       # evict the previous frame from L2 (outside the timed events)
        a.record()
        gi.update_regions(pos)
        gi.build_clipmap(0)
        m.record()
        gi.cone_trace(cam_v, g_v, prm, out=out)
        b.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    build_ms = sum(a.elapsed_time(m) for a, m, b in ev) / args.steps
    trace_ms = sum(m.elapsed_time(b) for a, m, b in ev) / args.steps
    step_ms = build_ms + trace_ms
    st = gi.stats()
    launches = (st.kernel_launches - l0) // args.steps

    # ---- end to end, two public calls (both: pinned host outputs, 66 MB D2H per step, synchronous):
    #  e2e              vgi_frame_view_host: the batched / headless view. Only the camera and the light's matrices go up;
    #                   shadow map and G-buffer are rasterised on the device INSIDE the timed region (bit-identical to
    #                   the host-rendered inputs of the device-resident path, asserted below), then build + trace.
    #  e2e_host_gbuffer vgi_frame_host: a host that owns the G-buffer uploads it every step (58 MB H2D); the shadow map
    #                   of the static light is uploaded with the first call only.
    hgb = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.int16) if v.dtype == np.uint16 else
                               np.ascontiguousarray(v)).pin_memory() for k, v in inp["gbuffer"].items()}
    hshadow = torch.from_numpy(inp["shadow_depth"]).pin_memory()
    hout = (torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory(),
            torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory())
    d2h = sum(t.numel() * 4 for t in hout)

    def timed(call):
        for _ in range(3):
            call()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        t0 = time.perf_counter()
        for a, b in evs:
            a.record()
            call()
            b.record()
        barrier()
        wall = (time.perf_counter() - t0) / args.steps * 1e3
        return max(sum(a.elapsed_time(b) for a, b in evs) / args.steps, 0.0), wall

    import ctypes as C
    from vk_voxel_cone_tracing_b200 import structs as S0
    h2d_view = C.sizeof(S0.Camera) + C.sizeof(S0.DirLightShadow) + C.sizeof(S0.VctParams) + 12
    def e2e_view():
        pos, cam_v, _ = next_view()
        gi.frame_view_host(0, pos, cam_v, WIDTH, HEIGHT, inp["shadow"], prm, hout[0], hout[1])

    e2e_sync_ms, sync_wall = timed(e2e_view)

    # the same views pipelined two deep (vgi_frame_view_host_begin / _end): the download of frame i overlaps the rasterisation
    # and build of frame i + 1; every frame's two images are in host memory before its _end returns, all inside the timed region
    hout2 = (torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory(),
             torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory())
    hbufs = (hout, hout2)

    def e2e_pipelined(nsteps):
        for i in range(nsteps):
            pos, cam_v, _ = next_view()
            gi.frame_view_host_begin(0, pos, cam_v, WIDTH, HEIGHT, inp["shadow"], prm, hbufs[i & 1][0], hbufs[i & 1][1])
            if i:
                gi.frame_view_host_end()
        gi.frame_view_host_end()

    e2e_pipelined(3)
    barrier()
    pev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    t0 = time.perf_counter()
    pev[0].record()
    e2e_pipelined(args.steps)
    pev[1].record()
    torch.cuda.synchronize()
    t_wall = (time.perf_counter() - t0) / args.steps * 1e3
    barrier()
    e2e_ms = pev[0].elapsed_time(pev[1]) / args.steps
    # the host copy of the result must equal the device-resident path's (whose inputs were rendered on the host)
    frame()
    gi.frame_view_host(0, views[0][0], views[0][1], WIDTH, HEIGHT, inp["shadow"], prm, hout[0], hout[1])
    assert torch.equal(hout[0], out[0].cpu()) and torch.equal(hout[1], out[1].cpu()), \
        "vgi_frame_view_host result differs from the device-resident path"
    # the host-G-buffer variant: every view's G-buffer in pinned host memory (copies of the device-rendered, bit-identical ones)
    hviews = [{k: t.cpu().pin_memory() for k, t in g_v.items()} for _, _, g_v in views]
    h2d_host = sum(t.numel() * t.element_size() for t in hviews[0].values())
    gi.frame_host(0, inp["cam_pos"], inp["cam"], hgb, hshadow, prm, hout[0], hout[1])     # static light: uploaded once

    def e2e_host():
        i = (rank + step_no[0]) % N_VIEWS
        step_no[0] += 1
        gi.frame_host(0, views[i][0], views[i][1], hviews[i], None, prm, hout[0], hout[1])

    e2e_hg_ms, hg_wall = timed(e2e_host)
    gi.frame_host(0, views[0][0], views[0][1], hviews[0], None, prm, hout[0], hout[1])
    assert torch.equal(hout[0], out[0].cpu()), "vgi_frame_host result differs from the device-resident path"

    # ---- max over ranks (and every rank's own times: efficiency must be read from the machine, not from the views)
    mine = torch.tensor([step_ms, build_ms, trace_ms, e2e_ms, t_wall, e2e_hg_ms, e2e_sync_ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"first_view": [r % N_VIEWS for r in range(world)],
                    "ms_per_step": [float(x[0]) for x in allr], "e2e_ms_per_step": [float(x[3]) for x in allr]}
        t = torch.stack(allr).max(dim=0).values
    else:
        t = mine
    step_ms, build_ms, trace_ms, e2e_ms, t_wall, e2e_hg_ms, e2e_sync_ms = t.tolist()

    # ---- per-kernel CUDA-event timing (separate pass: the events would perturb the headline number)
    roof, roof_trace, roof_stage, roof_dominant, kernels = None, None, None, None, {}
    if rank == 0:
        gi.set_timing(True)
        gi.reset_timings()
        for _ in range(args.steps):
            flush.zero_()
            frame()
        torch.cuda.synchronize()
        tm = gi.timings()
        gi.set_timing(False)
        kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in tm.items()}
        hbm_peak, peak_src, pk = peaks()
        build_k = {k: v for k, v in tm.items() if not k.startswith("k_trace")}
        if build_k:
            name = max(build_k, key=lambda k: build_k[k][0])
            ms_launch = build_k[name][0] / max(build_k[name][1], 1)
            ab = algorithmic_bytes(name, st, inp["cfg"])
            # The stage against real bytes: (a) algorithmic bytes of the SPARSE formulation that is actually timed
            # (sum of the per-kernel byte counts below), (b) DRAM bytes measured by ncu for this very
            # kernel build (profiles/r2_ncu_traffic.json, None when the capture belongs to another build). The dense
            # contract of SURVEY 8(d) (A_vim: both reference-layout atlases written once) is NOT what the timed region
            # moves - vgi_export_atlas is outside it - so it is reported as a time-to-solution ratio under its own name.
            sparse_bytes = 0
            for kname, (kms, kcount) in build_k.items():
                kb = algorithmic_bytes(kname.replace("_slab", ""), st, inp["cfg"])
                if kb:
                    sparse_bytes += kb * (kcount / args.steps)
            dram = [ncu_traffic(k) for k in build_k]
            dram_total = None
            if all(d is not None for d in dram):
                dram_total = sum(d * (build_k[k][1] / args.steps) for d, k in zip(dram, build_k))
            dense = a_vim(inp["cfg"], st.triangles, inp["scene"].vertex_count)
            roof_stage = {"stage": "voxelize+inject+mip (all kernels of vgi_build_clipmap)", "bound": "hbm",
                          "algorithmic_bytes": sparse_bytes, "achieved": sparse_bytes / (build_ms * 1e-3) / 1e9,
                          "peak": hbm_peak, "unit": "GB/s", "frac": sparse_bytes / (build_ms * 1e-3) / 1e9 / hbm_peak,
                          "traffic": dram_total,
                          "dense_contract_bytes": dense,
                          "dense_contract_speedup": (dense / (hbm_peak * 1e9)) / (build_ms * 1e-3),
                          "note": "algorithmic_bytes = sparse formulation (sum over the build's kernels, DESIGN.md section 4); frac is "
                                  "low because the stage is latency / L2 bound, not HBM bound. dense_contract_speedup = time the "
                                  "dense formulation of SURVEY 8(d) would need at the HBM peak / measured time; the reference-layout "
                                  "atlases themselves are only written by vgi_export_atlas, outside the timed region"}
            ach = (ab / (ms_launch * 1e-3) / 1e9) if ab else None
            roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (ach / hbm_peak) if ach else None, "traffic": ncu_traffic(name), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ab, "ms_per_launch": ms_launch,
                    "share_of_build": build_k[name][0] / sum(v[0] for v in build_k.values()),
                    "note": "dominant kernel of the HBM-bound half of the metric (voxelize+inject+mip); the dominant kernel of the "
                            "whole step, k_trace_main, is L1 / issue bound: see roofline"}
        trace_k = {k: v for k, v in tm.items() if k.startswith("k_trace")}
        if trace_k:
            tms = sum(v[0] for v in trace_k.values()) / args.steps
            sm_mhz = (clk or {}).get("sm_mhz") or pk.get("sm_max_mhz", 1965.0)
            nsm = torch.cuda.get_device_properties(local).multi_processor_count
            l1_peak = nsm * 128 * sm_mhz * 1e6 / 1e9
            roof_trace = {"kernels": sorted(trace_k), "bound": "l1", "ms_per_step": tms, "peak": l1_peak, "unit": "GB/s",
                          "peak_source": f"{nsm} SMs x 128 B/clk x {sm_mhz:.0f} MHz (clock sampled during the run)",
                          "achieved": None, "frac": None,
                          "traffic": (ncu_traffic("k_trace_main") or 0) + (ncu_traffic("k_trace_specular_warp") or ncu_traffic("k_trace_specular") or 0) or None,
                          "note": "A_cone = 32 B x tri-linear taps (counted by the oracle on the sampled rows, scaled) + 60 B x pixels"}

    # ---- every view of the fixed list on ONE GPU (N = 1 only): with `per_rank` of the N > 1 lines this separates the
    # machine's scaling from the cameras' different costs (rank r renders view r; the slowest view sets `value`)
    per_view = None
    if rank == 0 and world == 1:
        from vk_voxel_cone_tracing_b200 import raster as _raster, synth as _synth
        per_view = []
        ve = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
        for v in range(N_VIEWS):
            pos, dirv = view_camera(v)
            cam_v = _synth.make_camera(pos, dirv, aspect=WIDTH / HEIGHT)
            gb_v = gi.render_gbuffer(cam_v, WIDTH, HEIGHT)      # bit-identical to the host rasteriser (tests/test_gpu_raster.py)
            acc_ms = 0.0
            for i in range(5):
                flush.zero_()
                ve[0].record()
                gi.update_regions(pos)
                gi.build_clipmap(0)
                gi.cone_trace(cam_v, gb_v, prm, out=out)
                ve[1].record()
                torch.cuda.synchronize()
                if i >= 2:
                    acc_ms += ve[0].elapsed_time(ve[1])
            per_view.append(acc_ms / 3)
        frame()
        torch.cuda.synchronize()

    # ---- configs[2] beside it: 512^3 octree (level 9) fragment list + build + 1080p octree cone trace
    svo = None
    if rank == 0 and not args.no_svo:
        lo, hi = inp["scene"].world_bbox()
        sprm = gi.default_vct_params(8)
        sprm.volume_dimension = 512.0
        sprm.voxel_size = float((hi - lo).max() / 512.0)
        sprm.indirect_diffuse_intensity = 15.0       # OctreeVoxelConeTracing.h:73-80 defaults
        sprm.occlusion_decay = 3.0
        sev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc = np.zeros(3)
        n_it = 5
        for i in range(n_it + 2):
            flush.zero_()
            sev[0].record()
            gi.svo_voxelize(9, lo, hi)
            sev[1].record()
            gi.svo_build()
            sev[2].record()
            gi.svo_cone_trace(inp["cam"], dgb, sprm, out=out)
            sev[3].record()
            torch.cuda.synchronize()
            if i >= 2:
                acc += [sev[0].elapsed_time(sev[1]), sev[1].elapsed_time(sev[2]), sev[2].elapsed_time(sev[3])]
        nf, nn = gi.svo_fragments().shape[0], gi.svo_nodes().shape[0]
        svo = {"workload": "configs[2]: octree level 9 (512^3) on the same mesh + 1080p octree cone trace",
               "voxelize_ms": acc[0] / n_it, "build_ms": acc[1] / n_it, "trace_ms": acc[2] / n_it, "fragments": nf, "nodes": nn,
               "algorithmic_bytes_build": 12 * int(st.triangles) + 12 * inp["scene"].vertex_count + 8 * nf * (1 + 9) + 16 * nn}
        # the SVO pass reused the pair buffer and the output images: restore the clipmap state for what follows
        frame()
        torch.cuda.synchronize()

    # ---- the passes either side of the path, reported beside the metric, not part of it
    adjacent = None
    if rank == 0:
        from vk_voxel_cone_tracing_b200 import structs as S2
        fprm = S2.default_filter_params(1, 1)
        rg = gi.render_gbuffer(inp["cam"], WIDTH, HEIGHT)
        rs = gi.render_shadow_map(inp["shadow"], SHADOW)
        fo = gi.specular_filter(out[0], out[1], fprm)
        same_inputs = bool(torch.equal(rs.cpu(), torch.from_numpy(inp["shadow_depth"]))
                           and torch.equal(rg["depth"].cpu(), torch.from_numpy(inp["gbuffer"]["depth"])))
        aev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc3 = np.zeros(3)
        for i in range(7):
            flush.zero_()
            aev[0].record()
            gi.render_shadow_map(inp["shadow"], SHADOW, out=rs)
            aev[1].record()
            gi.render_gbuffer(inp["cam"], WIDTH, HEIGHT, out=rg)
            aev[2].record()
            gi.specular_filter(out[0], out[1], fprm, out=fo)
            aev[3].record()
            torch.cuda.synchronize()
            if i >= 2:
                acc3 += [aev[0].elapsed_time(aev[1]), aev[1].elapsed_time(aev[2]), aev[2].elapsed_time(aev[3])]
        adjacent = {"shadow_map_4096_ms": acc3[0] / 5, "gbuffer_1080p_ms": acc3[1] / 5,
                    "specular_filter_gaussian_tonemap_1080p_ms": acc3[2] / 5,
                    "device_rendered_inputs_equal_host_rendered": same_inputs}

    # ---- a moving camera beside it: vgi_build_clipmap_incremental along a straight walk. NOT the metric — the headline frame
    # always rebuilds everything (a cached build would be "cached outputs"); this is the reference's unused
    # `_fullRevoxelization = false` mode (VoxelizationPass.cpp:81-99): only the clip levels whose region moved are rebuilt.
    incremental = None
    if rank == 0 and not args.no_incremental:
        walk_n, speed = 64, 0.05            # 64 frames at 0.05 world units per frame (3 units/s at 60 frames/s) along +x
        walk = [(-8.0 + speed * f, 3.0, 0.0) for f in range(walk_n)]
        iev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ims, firsts = [], []
        for f, cam_f in enumerate(walk):
            gi.update_regions(cam_f)
            flush.zero_()
            iev[0].record()
            firsts.append(gi.build_clipmap_incremental(f))
            iev[1].record()
            torch.cuda.synchronize()
            ims.append(iev[0].elapsed_time(iev[1]))
        got = [gi.export_atlas(w).clone() for w in (0, 1)]
        fms = []
        for f, cam_f in enumerate(walk):    # the same walk with a full rebuild on every frame
            gi.update_regions(cam_f)
            flush.zero_()
            iev[0].record()
            gi.build_clipmap(f)
            iev[1].record()
            torch.cuda.synchronize()
            fms.append(iev[0].elapsed_time(iev[1]))
        same = all(bool(torch.equal(got[w], gi.export_atlas(w))) for w in (0, 1))
        del got
        rest = ims[1:]
        incremental = {"workload": f"{walk_n} frames, camera (-8,3,0) moving {speed} world units per frame along +x, cadence on "
                                   f"(frame index = walk index), same scene / light / 6 x {RES}^3 clipmap",
                       "call": "vgi_build_clipmap_incremental", "mean_ms": float(np.mean(rest)), "max_ms": float(np.max(rest)),
                       "first_frame_ms": ims[0], "full_rebuild_mean_ms": float(np.mean(fms[1:])),
                       "frames_with_nothing_to_rebuild": int(sum(1 for x in firsts if x == LEVELS)),
                       "first_level_rebuilt_histogram": {str(l): int(sum(1 for x in firsts if x == l)) for l in range(LEVELS + 1)},
                       "atlases_equal_full_rebuild_sequence": same}
        frame()
        torch.cuda.synchronize()

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        smp = CpuFrameSampler(inp)
        smp.prime()
        smp.i = 0
        for _ in range(LEVELS):
            smp.step()
        b_s, t_s = smp.frame_seconds()
        cpu = {"value": 1.0 / (b_s + t_s), "unit": "frames/s", **smp.describe(), "build_ms": 1e3 * b_s, "trace_ms": 1e3 * t_s}
        # parity of the headline frame: the GPU's atlases and images against what the CPU arm just computed
        frame()
        torch.cuda.synchronize()
        parity = smp.parity(gi.export_atlas(0).cpu().numpy(), gi.export_atlas(1).cpu().numpy(),
                            out[0].cpu().numpy(), out[1].cpu().numpy(), inp["gbuffer"]["depth"])
        if roof_trace and smp.taps:
            taps = float(np.mean(smp.taps))
            a_cone = 32.0 * taps + 60.0 * WIDTH * HEIGHT
            roof_trace.update({"taps_per_frame": taps, "algorithmic_bytes": a_cone,
                               "achieved": a_cone / (trace_ms * 1e-3) / 1e9,
                               "frac": a_cone / (trace_ms * 1e-3) / 1e9 / roof_trace["peak"]})
            if smp.taps_diffuse and "k_trace_main" in kernels:
                # the dominant kernel of the step on its own: diffuse cones' taps + the compulsory G-buffer / output bytes
                td = float(np.mean(smp.taps_diffuse))
                a_main = 32.0 * td + 60.0 * WIDTH * HEIGHT
                ms_main = kernels["k_trace_main"]["ms_per_step"]
                roof_dominant = {"kernel": "k_trace_main", "bound": "l1", "share_of_step": ms_main / step_ms,
                                 "taps_per_launch": td, "algorithmic_bytes_per_launch": a_main, "ms_per_launch": ms_main,
                                 "achieved": a_main / (ms_main * 1e-3) / 1e9, "peak": roof_trace["peak"], "unit": "GB/s",
                                 "frac": a_main / (ms_main * 1e-3) / 1e9 / roof_trace["peak"],
                                 "peak_source": roof_trace["peak_source"], "traffic": ncu_traffic("k_trace_main"),
                                 "note": "SURVEY 8(d): cone tracing is bounded by L1 bandwidth / issue rate, not HBM (DRAM traffic "
                                         "per launch in `traffic`); `roofline_hbm_kernel` carries the dominant HBM-bound kernel"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": world * 1e3 / step_ms, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (u8 texels)", "data": "synthetic",
            "config": config_dict(world, int(st.triangles)),
            "stages": {"build_ms": build_ms, "trace_ms": trace_ms, "trace_fps": 1e3 / trace_ms,
                       "clip_pairs": int(st.clip_pairs), "shaded_pairs": int(st.shaded_pairs),
                       "occupied_voxels": int(st.occupied_voxels)},
            "e2e": {"value": world * 1e3 / e2e_ms, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_view),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "wall_ms_per_step": t_wall,
                    "call": "vgi_frame_view_host_begin / _end, two frames in flight: camera + light matrices up; shadow map and "
                            "G-buffer rasterised on the device inside the timed region; both float4 images of every frame in "
                            "host memory before its _end returns (the download of frame i overlaps the work of frame i + 1)"},
            "e2e_sync": {"value": world * 1e3 / e2e_sync_ms, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_view),
                         "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_sync_ms,
                         "call": "vgi_frame_view_host, one frame at a time (the call returns with both images on the host)"},
            "e2e_host_gbuffer": {"value": world * 1e3 / e2e_hg_ms, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_host),
                                 "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_hg_ms,
                                 "call": "vgi_frame_host: host G-buffer uploaded every step (shadow map of the static light "
                                         "uploaded once, outside), both float4 images down"},
            "per_rank": per_rank, "per_view_ms_on_one_gpu": per_view,
            "gpu_launches": int(launches) * args.steps,
            "gpu_launches_per_step": int(launches),
            # `roofline` = the dominant kernel of the step (k_trace_main, bounded by L1 / issue rate per SURVEY 8d); without
            # the oracle's tap count (N > 1 or --no-cpu-baseline) it falls back to the dominant HBM-bound kernel
            "clocks": clk, "roofline": roof_dominant or roof, "roofline_hbm_kernel": roof, "roofline_stage": roof_stage,
            "roofline_cone_trace": roof_trace, "adjacent_passes": adjacent,
            "svo": svo, "incremental_build": incremental, "kernels": kernels,
            "cpu_baseline": cpu, "parity": parity, "kernel_source_hash": kernel_source_hash(),
        }
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["ok"]:
            print(f"bench.py: PARITY FAILED on the headline configuration: {parity}", file=sys.stderr)
            parity_failed = True
    if world > 1:
        dist.destroy_process_group()
    if parity_failed:
        raise SystemExit(3)


# ---------------------------------------------------------------------------------------------------
# --config 4: BASELINE configs[4] — 512^3 volume, slab-voxelized across the GPUs and exchanged, 64 camera views at 4K
# sharded by view. Not the headline metric (the driver runs the default); run under gpurun --gpus N and kept in profiles/.
# ---------------------------------------------------------------------------------------------------
def run_config4(args):
    import torch
    import torch.distributed as dist
    from vk_voxel_cone_tracing_b200 import multigpu as M, structs as S, synth
    from vk_voxel_cone_tracing_b200.api import VoxelGI

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libvgi has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    R4, W4, H4, NV = 512, 3840, 2160, 64
    scene = synth.atrium()
    cfg = S.default_config(R4, 1, extent_level0=32.0)
    light, shadow = synth.make_light()
    lo, hi = scene.world_bbox()
    centre = tuple(float(x) for x in (np.asarray(lo) + np.asarray(hi)) * 0.5)

    def make():
        g = VoxelGI(cfg, device=local)
        g.set_scene(scene)
        g.update_regions(centre)
        return g

    gi = make()
    sm = gi.render_shadow_map(shadow, SHADOW)
    gi.set_light(light, shadow, sm)
    rng = np.random.RandomState(5)
    cams = []
    for v in range(NV):
        ang = 2.0 * np.pi * v / NV
        pos = np.asarray(centre) + np.array([rng.uniform(6, 12) * np.cos(ang), rng.uniform(2, 8) - centre[1], rng.uniform(6, 12) * np.sin(ang)])
        d = np.asarray(centre) - pos
        cams.append(synth.make_camera(tuple(pos), tuple(d / np.linalg.norm(d)), aspect=W4 / H4))
    mine = list(M.views_for_rank(NV, rank, world))
    prm = gi.default_vct_params(8)
    out = (torch.zeros((H4, W4, 4), dtype=torch.float32, device=dev), torch.zeros((H4, W4, 4), dtype=torch.float32, device=dev))
    gbuf = gi.render_gbuffer(cams[0], W4, H4)
    hout = (torch.empty((H4, W4, 4), dtype=torch.float32).pin_memory(), torch.empty((H4, W4, 4), dtype=torch.float32).pin_memory())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    builders = {"replicated": lambda: gi.build_clipmap(0)}
    keep = []
    if world > 1:
        gs, gp = make(), make()
        for g in (gs, gp):
            g.set_light(light, shadow, sm)
        sb, pb = M.SlabBuild(gs), M.PeerBuild(gp)
        keep += [pb]
        builders["slab_nccl_allgather"] = lambda: sb.build(0)
        builders["peer_nvlink_stores"] = lambda: pb.build(0)

    def timed(fn, n):
        for _ in range(2):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    build_ms = {k: timed(f, 10) for k, f in builders.items()}
    # the sharded builds equal the replicated one bit for bit (checked on the store itself: 4.3 GB per GPU, no export)
    same = True
    if world > 1:
        from vk_voxel_cone_tracing_b200 import api as _api

        def store(g):
            ptr, nbytes = g.voxel_store()
            return torch.as_tensor(_api._DevView(ptr, (nbytes // 4,), "<i4"), device=dev)

        gi.build_clipmap(0)
        torch.cuda.synchronize()
        ref = store(gi)
        for g in (gs, gp):
            same = same and bool(torch.equal(ref, store(g)))
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
    best = min(build_ms, key=build_ms.get)
    build = builders[best]
    src = gi if best == "replicated" else (gs if best.startswith("slab") else gp)

    def step(download):
        build()
        for v in mine:
            src.render_gbuffer(cams[v], W4, H4, out=gbuf)
            src.cone_trace(cams[v], gbuf, prm, out=out)
            if download:
                hout[0].copy_(out[0], non_blocking=True)
                hout[1].copy_(out[1], non_blocking=True)
        if download:
            torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(lambda: step(False), args.steps)
    clk = clocks.stop() if rank == 0 else None
    ms_e2e = timed(lambda: step(True), args.steps)
    if rank == 0:
        line = {"metric": "configs[4]: 4K 16-cone GI views/s over 64 views sharded by view; 512^3 volume build ms (1/2/4/8 B200)",
                "value": NV * 1e3 / ms, "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": 2, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (u8 texels)", "data": "synthetic",
                "config": {"workload": "configs[4]: batched 64 camera views at 3840x2160 sharded by view across the GPUs, one 512^3 "
                                       "volume (single level, extent 32) built per step from the 262144-triangle atrium",
                           "resolution": R4, "levels": 1, "image": [W4, H4], "views": NV, "views_per_gpu": len(mine),
                           "build": best, "parallelism": f"{world} GPU(s): views round-robin, volume built by '{best}'"},
                "stages": {"build_ms": build_ms, "sharded_builds_equal_replicated_store": same,
                           "ms_per_view": (ms - build_ms[best]) / max(len(mine), 1)},
                "e2e": {"value": NV * 1e3 / ms_e2e, "unit": "views/s", "h2d_bytes_per_step": 344 * len(mine),
                        "d2h_bytes_per_step": 2 * H4 * W4 * 16 * len(mine), "ms_per_step": ms_e2e,
                        "call": "per view: camera up, G-buffer rasterised on the device, cone trace, both float4 images to pinned host memory"},
                "clocks": clk, "kernel_source_hash": kernel_source_hash()}
        print(json.dumps(line), flush=True)
    for k in keep:
        if hasattr(k, "close"):
            k.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vgi", choices=["vgi", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-svo", action="store_true")
    ap.add_argument("--no-incremental", action="store_true")
    ap.add_argument("--textured", action="store_true",
                    help="the same mesh with textured materials (base colour, emissive, occlusion alpha test in the build; "
                         "metallic-roughness, normal map, tangents in the device-rendered G-buffer): a number beside the metric")
    ap.add_argument("--config", type=int, default=1, choices=[1, 3, 4],
                    help="1 = the headline workload (BASELINE configs[1]); 3 = configs[3], the same pipeline on a 6-level 128^3 "
                         "clipmap traced at 3840x2160; 4 = configs[4], 64 views at 4K + 512^3 slab build")
    args = ap.parse_args()
    if args.textured:
        global TEXTURED
        TEXTURED = True
    if args.config == 3:        # same code path as the headline, other sizes (both arms read these module constants)
        global WORKLOAD, RES, WIDTH, HEIGHT
        RES, WIDTH, HEIGHT = 128, 3840, 2160
        WORKLOAD = ("configs[3]: Sponza-scale synthetic atrium (262144 tris), 6-level 128^3 clipmap voxelize+inject+mip "
                    "(frame 0: all levels) + 3840x2160 16-cone diffuse/specular GI (mode 8), one frame per step along a fixed "
                    "8-camera path")
    if TEXTURED:
        WORKLOAD += (" [--textured: synth.textured_atrium - base-colour / emissive textures and occlusion-mask alpha test in the "
                     "build, metallic-roughness + normal map + tangents in the device-rendered G-buffer]")
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 4:
        run_config4(args)
    else:
        run_vgi(args)


if __name__ == "__main__":
    main()
