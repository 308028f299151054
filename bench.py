#!/usr/bin/env python
"""Benchmark of the meshing hot path (BASELINE.json): evaluating the CSG signed-distance tree
over the lattice and running isosurface stuffing on it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one full meshing pass (SDF field -> warp -> classify -> vertex numbering ->
bisection -> tet emission -> quality) of the workload.  N = 1 runs BASELINE config 4: the
synthetic 64-primitive smooth union/difference scene on a 1024^3-cell lattice.  N > 1 shards
that same lattice into z-slabs (strong scaling), one process per GPU under torchrun.

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_primitives, cells per axis) — the synthetic CSG scenes of BASELINE configs 4 and 5
    "csg64_1024": (64, 1024),
    "csg64_512": (64, 512),
    "csg64_256": (64, 256),
    "csg256_2048": (256, 2048),
    "csg256_1024": (256, 1024),
    "csg16_64": (16, 64),
    # name: (script, cells per axis, tessellation error) — BASELINE configs 2 and 3
    "sphere_256": ("sphere", 256, 0.2),
    "xplicit_512": ("xplicit", 512, 0.8),   # 0.8: with the default 0.2 the reference's cut_edge never terminates
}
REL_ERR = 0.2  # --tessellation-error default, src/main.rs:48


def workload_desc(name):
    w = WORKLOADS[name]
    if isinstance(w[0], int):
        return (f"synthetic {w[0]}-primitive smooth CSG scene (numpy default_rng({w[0]})), {w[1]}^3-cell lattice, "
                f"tessellation error {REL_ERR}")
    return f"examples/{w[0]}.lua, {w[1]}^3-cell lattice, tessellation error {w[2]}"


def build_scene(geom_factory, name):
    """-> (backend owner, volume object, resolution, cells per axis, relative error, [(name, boundary object)])
    The synthetic scenes get one `:boundary()` object (`top`, built like examples/cube.lua's), xplicit.lua
    the added `top` of BASELINE config 3, so that Mesh::boundary and the side-set loop (src/main.rs:86-106)
    are part of every step."""
    from moosecad_b200 import luascad, scenes
    w = WORKLOADS[name]
    if isinstance(w[0], int):
        nprim, n = w
        g = geom_factory()
        root = scenes.synthetic_csg(g, nprim, seed=nprim, size_scale=0.5 if nprim >= 256 else 1.0)
        return g, root, 1.0 / n, n, REL_ERR, [("top", scenes.synthetic_csg_top(g, root))]
    script, n, err = w
    backend = geom_factory().backend
    out = luascad.eval({"sphere": scenes.SPHERE_LUA, "xplicit": scenes.XPLICIT_WITH_TOP_LUA}[script], backend=backend)
    root = [o.volume() for _, o in sorted(out.items()) if o.volume() is not None][0]
    bnds = [(nm, o.boundary()) for nm, o in sorted(out.items()) if o.boundary() is not None]
    bb = root.bbox()
    return root, root, (bb[3] - bb[0]) / n, n, err, bnds


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md's clocks line)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _oracle_scene(name):
    import moosecad_b200 as mc
    from oracle import pyoracle
    g, root, res, n, err, bnds = build_scene(lambda: mc.Geom(pyoracle.OracleScene()), name)
    root.backend.set_parameters_node(root.node, 0.1, 1.0)
    for _, b in bnds:
        b.backend.set_parameters_node(b.node, 0.1, 1.0)
    return root, res, n, err, bnds


def cpu_sample_cells(name):
    """Cells per axis of the bounded CPU sample of a workload (about 5 s of single-core work)."""
    w = WORKLOADS[name]
    if isinstance(w[0], int):
        return min(w[1], 40 if w[0] <= 64 else 24)
    return min(w[1], 96 if w[0] == "sphere" else 48)


def cpu_baseline_sample(name, cells=None):
    """The oracle (C++ restatement of the reference's serial Rust path, platform libm) timed on the
    host, one core, on a bounded sample: the SAME scene on a coarser lattice (`cells` per axis).
    The per-point cost of the reference depends on the resolution (far-field points are cheap, the
    surface fraction scales as 1/N): profiles/r02_cpu_baseline_*.json holds the measured curve."""
    from oracle import pyoracle
    pyoracle.set_libm_mode(0)  # the reference calls the platform libm
    try:
        root, res_full, n_full, err, bnds = _oracle_scene(name)
        n = cells or cpu_sample_cells(name)
        res = res_full * n_full / n
        t0 = time.perf_counter()
        om = root.backend.tessellate(root.node, res, err)       # IsosurfaceStuffing::tessellate, src/main.rs:74-82
        om.boundary()                                           # mesh.boundary(), src/main.rs:86
        for _, b in bnds:                                       # the side-set loop, src/main.rs:87-106
            om.add_sideset(b.backend, b.node)
        dt = time.perf_counter() - t0
        st = om.stats()
    finally:
        pyoracle.set_libm_mode(1)
    dim = st["dim"]
    pts = (dim[0] + 1) * (dim[1] + 1) * (dim[2] + 1)
    nt = len(om.elems())
    return {"value": pts / dt, "unit": "lattice pts/s", "cores": 1, "kind": "port",
            "sample": f"same scene on a {n}^3-cell lattice ({pts} lattice points, {nt} tets, {dt:.2f} s) instead of "
                      f"{n_full}^3; serial like the reference (no threads in tessellate3d / mesh / main)",
            "sample_cells_per_axis": n, "same_config": n == n_full,
            "tets_per_s": nt / dt, "seconds": dt, "stage_s": {k: round(v, 4) for k, v in st["times"].items()},
            "value_calls_per_point": st["n_value_calls"] / pts, "host_cores_available": os.cpu_count()}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust reference
    cannot be built here (no toolchain, un-vendored crates), so this times the oracle port, each step
    on a bounded sample (the same scene on a coarser lattice: `config.sample_cells_per_axis`,
    `same_config` false).  A throughput measured on the sample is NOT the throughput the reference
    would reach on the full lattice."""
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    vals, last = [], None
    for i in range(warm + steps):
        last = cpu_baseline_sample(args.workload)
        if i >= warm:
            vals.append(last)
    secs = sum(v["seconds"] for v in vals) / len(vals)
    v = sum(v["value"] for v in vals) / len(vals)
    line = {"impl": "reference", "metric": "lattice_points_per_s", "value": v, "unit": "lattice pts/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": secs * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_desc(args.workload) + " — each step a bounded sample: the same scene on a "
                                   f"{last['sample_cells_per_axis']}^3-cell lattice",
                       "sample_cells_per_axis": last["sample_cells_per_axis"], "same_config": last["same_config"]},
            "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample", "sample_cells_per_axis",
                                                  "same_config", "stage_s")},
            "e2e": {"value": v, "unit": "lattice pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "tets_per_s": sum(x["tets_per_s"] for x in vals) / len(vals)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="csg64_1024", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import moosecad_b200 as mc
    from moosecad_b200 import _lib, slabs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (moosecad_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # host-valued bookkeeping (slab class table, per-slab counts) travels over a host group
        # (BENCH_HOST_GROUP=0: everything over NCCL)
        host_group = dist.new_group(backend="gloo") if os.environ.get("BENCH_HOST_GROUP", "1") != "0" else None
    steps, warm = max(1, args.steps), max(3, args.warmup)

    stream = torch.cuda.Stream(device=dev)
    lib = _lib.lib()
    with torch.cuda.stream(stream):
        ctx = mc.Context(local_rank, stream.cuda_stream)
        g, root, res, n, rel_err, bnds = build_scene(mc.Geom, args.workload)
        g.backend.set_parameters(0.1, 1.0)
        info = g.backend.program_info(root.node, res)
        dim = slabs.lattice_dim(root.bbox(), res)
        P_total = (dim[0] + 1) * (dim[1] + 1) * (dim[2] + 1)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        host_ms = {}

        def boundary_and_sidesets(h, objs):
            # Mesh::boundary + the side-set loop on this rank's slab (slab-aware, no exchange)
            ns = C.c_uint64()
            _lib.check(lib.mc_mesh_boundary(h, C.byref(ns)))
            sizes = []
            for _, b in objs:
                idx = _lib.check(lib.mc_mesh_add_sideset(h, b.backend.handle, b.node))
                sizes.append(int(lib.mc_mesh_sideset_len(h, idx)))
            return int(ns.value), sizes

        side_counts = {}

        def one_step(flags=0):
            sm = slabs.SlabMesher(ctx, root, res, rel_err, rank, world, None, dev, flags=flags, host_group=host_group)
            h = sm.run()
            t0 = time.perf_counter()
            side_counts["boundary"], side_counts["sets"] = boundary_and_sidesets(h, bnds)
            sm.host_ms["boundary_sidesets"] = (time.perf_counter() - t0) * 1e3
            host_ms.update(sm.host_ms)
            side_counts["slab"] = list(sm.slabs[rank])
            slabs.report_device_time(sm, sum(stage_ms(h)))   # measured-load refinement of the slabs (first runs only)
            return h

        def stage_ms(h):
            ms = (C.c_double * 9)()
            lib.mc_mesh_stage_ms(h, ms)
            return list(ms)

        def counts(h):
            c = (C.c_uint64 * 6)()
            lib.mc_mesh_counts(h, c)
            return list(c)

        # ---- device-resident throughput ("value") ----
        for _ in range(warm):
            h = one_step()
            lib.mc_mesh_free(h)
        # one instrumented pass: FP64 operations the SDF kernel executes per launch (untimed)
        h = one_step(_lib.MC_OPT_COUNT_FLOPS)
        ts = (C.c_uint64 * 6)()
        lib.mc_mesh_tile_stats(h, ts)
        tile_stats = list(ts)
        lib.mc_mesh_free(h)
        fp64_peak = ctx.probe_fp64_nofma()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = ctx.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage_acc = [0.0] * 9
        ev0.record(stream)
        last_counts = None
        step_wall = []
        for i in range(steps):
            tw = time.perf_counter()
            h = one_step()
            step_wall.append(round((time.perf_counter() - tw) * 1e3, 2))
            st = stage_ms(h)
            stage_acc = [a + b for a, b in zip(stage_acc, st)]
            last_counts = counts(h)
            if i < steps - 1:
                lib.mc_mesh_free(h)
        ev1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_total = ev0.elapsed_time(ev1)
        launches = ctx.launch_count - launches0
        # identity of the mesh (outside the timed region): order- and numbering-independent device
        # checksums of the last step's mesh, summed over the slabs mod 2^64 — equal for every N
        from moosecad_b200.tessellate3d import mesh_checksum
        my_cs = mesh_checksum(h)
        lib.mc_mesh_free(h)
        all_cs = [my_cs]
        if world > 1:
            all_cs = [None] * world
            dist.all_gather_object(all_cs, my_cs)
        checksum = {"vertices": "%016x" % (sum(c["vertices"] for c in all_cs) % (1 << 64)),
                    "tets": "%016x" % (sum(c["tets"] for c in all_cs) % (1 << 64)),
                    "tets_skipped": sum(c["tets_skipped"] for c in all_cs)}
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        cnt = torch.tensor(last_counts[:2] + [launches, side_counts["boundary"]] + side_counts["sets"], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        ms_step = float(t.item()) / steps
        n_verts, n_tets, launches_all, n_sides, *set_sizes = (int(x) for x in cnt.tolist())
        stage_avg = [a / steps for a in stage_acc]
        # diagnostics (outside the timed region): every rank's own device time and host phases
        mine = {"rank": rank, "device_stage_ms_sum": round(sum(stage_avg), 3),
                "host_ms_last_step": {k: round(v, 3) for k, v in host_ms.items()}, "host_wall_ms_per_step": step_wall,
                "slab_cell_layers": side_counts.get("slab")}
        per_rank = [mine]
        if world > 1:
            per_rank = [None] * world
            dist.all_gather_object(per_rank, mine)

        # ---- end to end through the C ABI with HOST buffers ----
        e2e = None
        if not args.no_e2e:
            my_c = last_counts
            idx_bytes = 4
            coords_h = torch.empty(max(1, my_c[0]) * 3, dtype=torch.float64).pin_memory()
            tets_h = torch.empty(max(1, my_c[1]) * 4, dtype=torch.int32).pin_memory()
            e_steps = max(1, min(steps, 2))
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            a0.record(stream)
            h2d = 0
            for _ in range(e_steps):
                g2, root2, _, _, _, bnds2 = build_scene(mc.Geom, args.workload)     # scene lowering from host data
                g2.backend.set_parameters(0.1, 1.0)
                sm = slabs.SlabMesher(ctx, root2, res, rel_err, rank, world, None, dev, host_group=host_group)
                h = sm.run()
                boundary_and_sidesets(h, bnds2)
                t_copy = time.perf_counter()
                ib = C.c_int()
                lib.mc_mesh_device_views(h, None, None, C.byref(ib))
                idx_bytes = ib.value
                _lib.check(lib.mc_mesh_copy_to_host(h, C.c_void_p(coords_h.data_ptr()), C.c_void_p(tets_h.data_ptr()), 4))
                copy_ms = (time.perf_counter() - t_copy) * 1e3
                h2d = g2.backend.program_info(root2.node, res)["words"] * 8
                lib.mc_mesh_free(h)
            a1.record(stream)
            barrier()
            wall = time.perf_counter() - w0
            te = torch.tensor([max(a0.elapsed_time(a1), wall * 1e3)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            ms_e2e = float(te.item()) / e_steps
            d2h = torch.tensor([my_c[0] * 24 + my_c[1] * 16], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(d2h, op=dist.ReduceOp.SUM)
            e2e = {"value": P_total / (ms_e2e * 1e-3), "unit": "lattice pts/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h.item()),
                   "tets_per_s": n_tets / (ms_e2e * 1e-3),
                   "what": "scene lowering + program upload + meshing + Mesh::boundary + side-sets + copy of coords "
                           f"(f64) and tets (u{idx_bytes * 8}) into pinned host buffers, via the C ABI"}
            mine["e2e_copy_ms_last_step"] = round(copy_ms, 2)
            mine["e2e_copy_gb_per_s"] = round((my_c[0] * 24 + my_c[1] * 4 * idx_bytes) / (copy_ms * 1e-3) / 1e9, 2)
            if world > 1:
                dist.all_gather_object(per_rank, mine)
            else:
                per_rank = [mine]
            del coords_h, tets_h

        if rank == 0:
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except OSError:
                pass
            hbm_peak = peaks.get("hbm_gbs", 6650.0)
            # dominant kernel: the SDF field kernel (k_sdf_field_tiled).  Algorithmic work per launch =
            # the FP64 operations the tile-specialised evaluator executes (counted by an instrumented
            # launch of the same kernel, MC_OPT_COUNT_FLOPS; add/sub/mul/max/cmp/div/sqrt = 1, exp 23,
            # log 27).  The evaluate-all model of the full tape is given for reference.
            p_eval = last_counts[2]
            sdf_s = stage_avg[0] * 1e-3
            sdf_flops = tile_stats[4]
            roof = {"bound": "fp64_nofma", "kernel": "k_sdf_field_tiled (+ k_tile_decide)",
                    "achieved": sdf_flops / sdf_s / 1e12 if sdf_s > 0 else None,
                    "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                    "frac": (sdf_flops / sdf_s) / fp64_peak if sdf_s > 0 else None, "traffic": None,
                    "peak_source": "measured live: mc_probe_fp64_nofma (DADD+DMUL, --fmad=false); "
                                   "not in MEASURED_PEAKS.json (SURVEY.md \u00a78d)",
                    "flops_executed_per_launch": sdf_flops,
                    "flops_per_point_executed": sdf_flops / p_eval if p_eval else None,
                    "flops_per_point_evaluate_all_model": info["flops"],
                    "tiles": {"total": tile_stats[0], "sdf_skipped_proven_uniform": tile_stats[1],
                              "surface_point_tiles": tile_stats[2], "surface_cell_tiles": tile_stats[3],
                              "sub_boxes_proven_of_32_per_evaluated_tile": tile_stats[5]},
                    "ms": stage_avg[0]}
            stuff_ms = sum(stage_avg[1:9])  # warp, classify, vertex, bisect, tet emit (+ fused quality), boundary, side-sets
            bytes_alg = 8 * P_total + 24 * n_verts + 4 * 4 * n_tets
            roof2 = {"bound": "hbm", "kernels": "k_tile_minmax..k_tet_quality_list (all stages after the SDF field)",
                     "achieved": bytes_alg / (stuff_ms * 1e-3) / 1e9 if stuff_ms > 0 else None, "peak": hbm_peak,
                     "unit": "GB/s", "frac": bytes_alg / (stuff_ms * 1e-3) / 1e9 / hbm_peak if stuff_ms > 0 else None,
                     "algorithmic_bytes": bytes_alg, "traffic": None, "ms": stuff_ms,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650"}
            line = {"metric": "lattice_points_per_s", "value": P_total / (ms_step * 1e-3), "unit": "lattice pts/s",
                    "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms_step,
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                    "data": "synthetic",
                    "config": {"workload": workload_desc(args.workload) + ", " +
                                           ("z-slabs over %d GPUs" % world if world > 1 else "single GPU"),
                               "lattice_points": P_total, "output_vertices": n_verts, "output_tets": n_tets,
                               "boundary_sides": n_sides,
                               "side_sets": dict(zip([nm for nm, _ in bnds], set_sizes)),
                               "mesh_checksum": checksum,
                               "l2": "working set (>= 8 B x lattice points per array) is far larger than the 126 MB L2; no flush needed",
                               "tape_words": info["words"]},
                    "tets_per_s": n_tets / (ms_step * 1e-3),
                    "sdf_points_per_s": P_total / sdf_s if sdf_s > 0 else None,
                    "stage_ms": dict(zip(["sdf_field", "warp", "classify", "vertex", "bisect", "tet_emit", "quality",
                                          "boundary", "sidesets"], [round(x, 3) for x in stage_avg])),
                    "roofline": roof, "roofline_stuffing": roof2,
                    "gpu_launches": launches_all, "clocks": clocks, "e2e": e2e}
            if world > 1:
                line["per_rank"] = per_rank
            # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this workload
            try:
                with open(os.path.join(ROOT, "profiles", "r02_sdf_traffic.json")) as f:
                    tr = json.load(f).get(f"{args.workload}@{world}")
                if tr:
                    roof["traffic"] = tr["dram_bytes_per_launch"]
                    roof["traffic_source"] = tr["source"]
            except (OSError, ValueError):
                pass
            # ... and of the stuffing / boundary / side-set kernels (same capture): wasted re-reads show up here
            try:
                with open(os.path.join(ROOT, "profiles", "r02_stuffing_traffic.json")) as f:
                    tr = json.load(f).get(f"{args.workload}@{world}")
                if tr:
                    roof2["traffic"] = tr["dram_bytes_per_step"]
                    roof2["traffic_per_kernel"] = {k: v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in tr["per_kernel"].items()}
                    roof2["traffic_source"] = tr["source"]
            except (OSError, ValueError):
                pass
            if not args.no_cpu_baseline and world == 1:
                # bounded samples of the same scene: ~5 s and (csg64) ~25 s of single-core work
                cb = cpu_baseline_sample(args.workload)
                sizes = [{k: cb[k] for k in ("sample_cells_per_axis", "value", "tets_per_s", "seconds", "stage_s")}]
                if args.workload.startswith("csg64"):
                    c2 = cpu_baseline_sample(args.workload, 64)
                    sizes.append({k: c2[k] for k in ("sample_cells_per_axis", "value", "tets_per_s", "seconds", "stage_s")})
                cb["sizes"] = sizes
                cb["note"] = ("throughput on a coarser lattice of the same scene; the measured dependence on the lattice "
                              "size is in profiles/r02_cpu_baseline_*.json (N = 40..256)")
                line["cpu_baseline"] = cb
            print(json.dumps(line), flush=True)
        del ctx
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
