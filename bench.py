#!/usr/bin/env python
"""bench.py -- Gauss-Newton Mpix-iters/s of the SMVS depth-refinement hot path.

Workload (BASELINE.json configs[1]): 1 reference view + 6 neighbours,
1920x1080, finest scale of `-o2` (scale 2: 478x268 patches of 4x4 px, 16
samples each, every pixel sampled), no shading. One STEP = one inner Newton
loop of DepthOptimizer::run_newton_iterations (lib/depth_optimizer.cc:204-304:
construct -> PCG -> node update / active set, repeated until < 5 % of the nodes
are active) started from the same perturbed surface. A pixel-iteration is one
sample of one processed patch in one Newton step.

  value  pixel-iterations / device time of the loop (CUDA events on the
         library's stream, inputs resident in HBM)
  e2e    the same loop through the C ABI from HOST buffers, every step:
         smvsb_set_views_u8 (H2D of the 7 byte images from pinned memory +
         StereoView::set_scale on the device) + smvsb_set_surface (H2D of nodes,
         validity, visibility) + smvsb_newton_loop + smvsb_get_nodes (D2H),
         wall clock around the calls with a device synchronize on both sides

N > 1 (torchrun, one rank per GPU): reference views are independent units
(app/smvsrecon.cc:658-733), so every rank refines its own view (seed = rank);
no data-path collective; weak scaling. `value` = pixel-iterations of all ranks
/ max-over-ranks time.

--impl reference times the reference's own CPU implementation (oracle/_ref:
the reference's sources compiled verbatim against the MVE shim; or the
oracle port when that is absent) on bounded windows of the same workload, one
window per host thread (the reference's ThreadPool runs one view per thread).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from smvs_b200.workload import build_workload  # noqa: E402,F401  (tests import it from here)

WIDTH, HEIGHT, N_SUB, SCALE = 1920, 1080, 6, 2
REGULARIZATION = 0.01          # app/smvsrecon.cc:712 with alpha = 1
METRIC = "Gauss-Newton Mpix-iters/sec"
UNIT = "Mpix-iters/s"


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Samples that arrived inside [t_begin, t_end] (the timed region); the
        sampler is started before the warm-up because nvidia-smi needs a few
        hundred ms to produce its first line."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = [ln for (t, ln) in self.lines
                 if t_begin is None or (t_begin <= t <= t_end + 0.05)]
        if not lines:          # region shorter than the sampling jitter: nearest ones
            lines = [ln for (_, ln) in self.lines[-5:]]
        for ln in lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arms (oracle used as the thing timed ONLY here, as the task allows)
# ---------------------------------------------------------------------------

def _ref_scene_for(wl):
    """A reference DepthOptimizer state fed with exactly the workload's
    prepared arrays."""
    from oracle import ref as oref
    R = oref.RefScene(wl.scene, init_linear=wl.shading is not None)
    R.set_arrays(0, wl.main_grad, None)
    for k in range(wl.scene.n_sub):
        R.set_arrays(k + 1, wl.sub_grads[k], wl.sub_hess[k])
    if wl.shading is not None:
        R.set_shading(wl.shading, wl.shading_grad)
    R.surface_create(wl.scale, np.full((wl.scene.height, wl.scene.width), 5.0, np.float32))
    info = R.surface_info()
    assert (info["npx"], info["npy"], info["start_x"], info["start_y"]) == \
        (wl.npx, wl.npy, wl.start_x, wl.start_y)
    R.surface_set(wl.nodes, wl.node_valid, wl.patch_valid)
    R.set_visibility(wl.vis_off, wl.vis_ids)
    return R


def _windows(wl, n, frac_x=4, frac_y=4):
    nx, ny = max(wl.npx // frac_x, 1), max(wl.npy // frac_y, 1)
    out = []
    for i in range(n):
        gx, gy = i % frac_x, (i // frac_x) % frac_y
        out.append((gx * nx, gy * ny, nx, ny))
    return out


class ReferenceWorkers:
    """`threads` reference optimizers, each on its own bounded window (1/frac^2
    of the patch grid) of the same workload; run() lets every one do `repeats`
    Newton loops concurrently, one host thread each (the reference's
    ThreadPool model: one view per thread, app/smvsrecon.cc:558)."""

    def __init__(self, wl, threads, frac=4):
        from oracle import ref as oref
        if not oref.available():
            raise RuntimeError("oracle/_ref missing")
        self.threads = threads
        self.wins = _windows(wl, threads, frac, frac)
        self.subs = [wl.restrict(*w) for w in self.wins]
        self.scenes = [_ref_scene_for(sub) for sub in self.subs]
        self.desc = (f"{threads} thread(s), each Newton loop on a "
                     f"{self.wins[0][2]}x{self.wins[0][3]}-patch window (1/{frac * frac} of the "
                     f"{wl.npx}x{wl.npy} grid) of the same 2 MP / 6-neighbour scale-2 workload")

    def run(self, repeats=1):
        results = [0.0] * self.threads

        def work(i):
            R, sub = self.scenes[i], self.subs[i]
            for _ in range(repeats):
                R.surface_set(sub.nodes, sub.node_valid, sub.patch_valid)
                st = R.newton_loop(None, REGULARIZATION, 0.0)
                results[i] += st["pixel_iterations"]

        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(i,)) for i in range(self.threads)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return float(sum(results)), time.perf_counter() - t0

    def close(self):
        for R in self.scenes:
            R.close()


# ---------------------------------------------------------------------------

def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    # one optimizer per host thread, like the reference's ThreadPool (one view
    # each, app/smvsrecon.cc:558). The reference's throughput stops growing
    # with threads early (every vector operation of its CG allocates, and the
    # threads share one address space): measured on the 128-core GPU box with
    # quarter-grid windows, 16 threads reach 1.47 Mpix-iters/s, 32 threads 1.93
    # (and, with 1/16 windows, 8 threads 0.70 but 64 threads only 0.43). The
    # arm runs 32 threads (or SMVSB_REF_THREADS). Each thread runs the Newton loop of a quarter of the
    # patch grid (12-18 s a step); smaller windows would sell the reference
    # short, because its CG vectors span the whole node grid whatever part of
    # it is valid (measured on one core: 0.109 Mpix-iters/s on a quarter,
    # 0.078 on 1/16, 0.029 on 1/64). Long runs (> 16 steps) fall back to 1/16.
    threads = int(os.environ.get("SMVSB_REF_THREADS", "0")) or min(os.cpu_count() or 1, 32)
    threads = max(1, min(threads, 64))
    wl = build_workload(WIDTH, HEIGHT, N_SUB, SCALE, shading=False, seed_index=0)
    frac = 2 if (args.steps + args.warmup) <= 16 else 4
    workers = ReferenceWorkers(wl, threads, frac)
    desc = workers.desc
    for _ in range(args.warmup):
        workers.run()
    pix, secs = 0.0, 0.0
    for _ in range(args.steps):
        p, s = workers.run()
        pix += p
        secs += s
    workers.close()
    value = pix / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "1 ref view + 6 neighbours, 1920x1080, scale 2 (-o2), "
                               "no shading: inner Newton loop",
                   "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads,
                         "kind": "reference", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_product(args):
    import torch
    import torch.distributed as dist
    from smvs_b200 import api

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (smvs_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl = build_workload(WIDTH, HEIGHT, N_SUB, SCALE, shading=False, seed_index=rank)

    def pin(a):
        # page-locked host copies: the e2e arm copies from pinned memory
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()

    wl.scene.images = [pin(a) for a in wl.scene.images]
    wl.nodes = pin(wl.nodes)
    ctx = api.Context(local)
    nodes_out = None

    e2e_parts = np.zeros(4)

    def e2e_step():
        t = [time.perf_counter()]
        wl.push_views_u8(ctx)        # byte images; set_scale runs on the device
        t.append(time.perf_counter())
        wl.push_surface(ctx)
        t.append(time.perf_counter())
        st = ctx.newton_loop(None, REGULARIZATION, 0.0)
        t.append(time.perf_counter())
        nodes = ctx.get_nodes()
        t.append(time.perf_counter())
        e2e_parts[:] += np.diff(t)
        return st, nodes

    def resident_step():
        ctx.set_nodes(wl.nodes)             # reset; not part of the timed loop
        return ctx.newton_loop(None, REGULARIZATION, 0.0)

    sampler = ClockSampler(local)
    sampler.start()
    # warm-up (both paths)
    wl.push(ctx)
    for _ in range(max(args.warmup, 3)):
        resident_step()
    e2e_step()

    # ---- device-resident arm --------------------------------------------
    barrier()
    launches0 = ctx.launches
    t_dev_ms, pix, newton, cg = 0.0, 0.0, 0, 0
    cg_blocks, cg_rows = 0.0, 0.0
    t_split = np.zeros(3)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = resident_step()
        t_dev_ms += st["ms_total"]
        pix += st["pixel_iterations"]
        newton += st["newton_steps"]
        cg += st["cg_iterations"]
        cg_blocks += st["cg_block_iterations"]
        cg_rows += st["cg_row_iterations"]
        t_split += [st["ms_construct"], st["ms_solve"], st["ms_update"]]
    barrier()
    t1 = time.perf_counter()
    wall_resident = t1 - t0
    launches = ctx.launches - launches0
    clocks = sampler.stop(t0, t1)

    # ---- end-to-end arm ---------------------------------------------------
    barrier()
    e2e_parts[:] = 0.0
    t0 = time.perf_counter()
    pix_e2e = 0.0
    for _ in range(args.steps):
        st, nodes_out = e2e_step()
        pix_e2e += st["pixel_iterations"]
    barrier()
    wall_e2e = time.perf_counter() - t0

    # ---- reduce over ranks ---------------------------------------------------
    vals = torch.tensor([t_dev_ms, wall_e2e, pix, pix_e2e, float(launches)],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_dev_ms, wall_e2e = float(mx[0]), float(mx[1])
        pix, pix_e2e, launches = float(sm[2]), float(sm[3]), float(sm[4])

    if rank == 0:
        value = pix / (t_dev_ms * 1e-3) / 1e6
        e2e_value = pix_e2e / wall_e2e / 1e6

        # roofline of the dominant kernel: the PCG (one persistent launch per
        # Newton step). Algorithmic bytes per CG iteration (DESIGN.md section
        # 5): 128 per 4x4 block of the system + per block row 128 (P) and
        # 12 * 32 (vector passes). The system of a Newton step holds the blocks
        # whose two nodes are still active (the reference drops the others,
        # lib/gauss_newton_step.cc:91-105); the kernel reports their number.
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        algorithmic = cg_blocks * 128.0 + cg_rows * (128.0 + 12 * 32.0)
        cg_ms = float(t_split[1])
        achieved = (algorithmic / max(cg_ms * 1e-3, 1e-12)) / 1e9
        # blocks of the full system (every valid node active), for scaling the
        # ncu capture of such a launch to the launches of this run
        nv = wl.node_valid.reshape(wl.npy + 1, wl.npx + 1).astype(bool)
        blocks_full = 0
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                a = nv[max(dy, 0):nv.shape[0] + min(dy, 0), max(dx, 0):nv.shape[1] + min(dx, 0)]
                b = nv[max(-dy, 0):nv.shape[0] + min(-dy, 0), max(-dx, 0):nv.shape[1] + min(-dx, 0)]
                blocks_full += int((a & b).sum())
        roofline = {"bound": "hbm", "kernel": "cg_kernel (persistent PCG)",
                    "achieved": achieved, "peak": hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650",
                    "unit": "GB/s", "frac": achieved / hbm_peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full
                    # capture (profiles/r1g_cg.txt: 27.15 GB for a 200-iteration launch
                    # on the full system = 135.8 MB per CG iteration = 127 B per 4x4
                    # block: H only -- P and the vectors stay in L2), scaled by the
                    # blocks this run's launches read
                    "traffic": 27.15e9 / (200.0 * blocks_full) * cg_blocks / max(newton, 1),
                    "traffic_source": "profiles/r1g_cg.txt, scaled by system blocks x iterations",
                    "frac_dram": (27.15e9 / (200.0 * blocks_full) * cg_blocks
                                  / max(cg_ms * 1e-3, 1e-12)) / 1e9 / hbm_peak,
                    "algorithmic_bytes_per_launch": algorithmic / max(newton, 1),
                    "system_blocks_full": blocks_full,
                    "system_block_iterations_per_launch": cg_blocks / max(newton, 1),
                    "launches_timed": newton}

        cpu_base = None          # timed on rank 0 at N = 1 only
        if not args.no_cpu_baseline and world == 1:
            try:
                workers = ReferenceWorkers(wl, 1, 2)
                p, s = workers.run(repeats=1)
                workers.close()
                cpu_base = {"value": p / s / 1e6, "unit": UNIT, "cores": 1,
                            "kind": "reference", "sample": workers.desc}
            except Exception as exc:      # noqa: BLE001
                cpu_base = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                            "sample": f"unavailable: {exc}"}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_dev_ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "1 ref view + 6 neighbours, 1920x1080, scale 2 (-o2), "
                                   "no shading: inner Newton loop from a 2% perturbed surface",
                       "views_per_gpu": 1, "parallelism": f"views sharded, {world} GPU(s)",
                       "l2": "inputs_exceed_l2 (images 265 MB, H 148 MB, patch blocks 262 MB)",
                       "newton_steps_per_loop": newton / max(args.steps, 1),
                       "cg_iterations_per_loop": cg / max(args.steps, 1),
                       "ms_construct_solve_update": [float(x) / max(args.steps, 1)
                                                     for x in t_split]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": wl.h2d_bytes_u8(),
                    "d2h_bytes_per_step": int(nodes_out.nbytes),
                    "ms_per_step": 1e3 * wall_e2e / max(args.steps, 1),
                    "ms_set_views_set_surface_loop_get_nodes":
                        [1e3 * float(x) / max(args.steps, 1) for x in e2e_parts]},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_base,
            "wall_s_resident": wall_resident,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smvs_b200", choices=["smvs_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
