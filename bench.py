#!/usr/bin/env python
"""bench.py — chunks/s generated+meshed (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the hot path over one batch of synthetic input: StdGenerator(seed) terrain for the rank's
slab (shell included) -> [N>1: slab-boundary halo exchange over NCCL] -> mesh every interior chunk.
  value : chunks/s, device-resident (CUDA events on the library's stream, max over ranks), whole job
  e2e   : the same metric through the public API with HOST buffers: generate+mesh, then the meshes and the span
          table are copied to pinned host memory inside the timed region
Rank 0 prints ONE JSON line. See DESIGN.md "Measurement" for the roofline / cpu_baseline definitions.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "chunks_per_sec_generated_and_meshed"
UNIT = "32^3 chunks/s"
SEED = 0
# BASELINE.json configs[2]: full worldgen+mesh of a 128x128x16 chunk region on 1xB200 (the largest single-GPU
# configuration; 32 GiB of u32 blocks). N>1 keeps this slab per rank (x-slabs, weak scaling), cfg-4 style.
SLAB = (128, 16, 128)  # nx, ny, nz interior chunks per rank
CPU_SAMPLE = (32, 16, 32)  # bounded CPU sample: a 32x32x16 slice of the same region


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
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
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(threads=None, dims=CPU_SAMPLE):
    """The oracle (C restatement of the reference CPU path) on the host cores, reference-style worker pool."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    cores = os.cpu_count() or 1
    if threads is None:
        threads = max(1, cores - 2)  # the reference default (src/config.rs:40-41)
    r = O.baseline_region(SEED, (0, 0, 0), dims, threads)
    gen_rate = r.gen_chunks / r.gen_seconds
    mesh_rate = r.chunks / r.mesh_seconds
    both = 1.0 / (1.0 / gen_rate + 1.0 / mesh_rate)
    return {
        "value": both, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"seed {SEED}, {dims[0]}x{dims[2]}x{dims[1]} chunk slice (+1-chunk shell generated) of the bench region: "
                  f"generate+compress_rle {r.gen_chunks} chunks in {r.gen_seconds:.2f}s, RLE-iterator mesh {r.chunks} chunks in "
                  f"{r.mesh_seconds:.2f}s; C restatement of the reference CPU path (Rust toolchain unavailable)",
        "gen_chunks_per_s": gen_rate, "mesh_chunks_per_s": mesh_rate, "host_cores": cores,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        base = cpu_baseline(threads=os.cpu_count() or 1)
        if i >= args.warmup:
            vals.append((base["value"], time.time() - t0))
    v = sum(x[0] for x in vals) / len(vals)
    base["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * sum(x[1] for x in vals) / len(vals), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64+u32", "data": "synthetic",
        "config": workload_config(args.gpus), "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n):
    nx, ny, nz = SLAB
    return {
        "workload": f"full worldgen+mesh, StdGenerator seed {SEED}, {nx * n}x{nz}x{ny} chunk region (x-slabs of {nx}x{nz}x{ny} per GPU"
                    f"{', NCCL halo exchange of slab-boundary voxel planes' if n > 1 else ''}); BASELINE.json configs[2] per GPU",
        "chunks_per_gpu": nx * ny * nz, "l2": "inputs larger than L2 (32 GiB of blocks per GPU per step)", "precision": "f64 noise",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--slab", default=None, help="override nx,ny,nz interior chunks per rank (debug only; invalidates the number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import ballxworld_b200 as bxw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    slab = tuple(int(x) for x in args.slab.split(",")) if args.slab else SLAB
    nx, ny, nz = slab
    ctx = bxw.Context(local)
    # one explicit stream for everything: the library's kernels, torch's events and NCCL's ordering all hang off it
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    reg = bxw.standard_registry()
    region = bxw.Region((rank * nx, 0, 0), (nx, ny, nz), reg, ctx=ctx)
    n_chunks = nx * ny * nz

    halo = None
    if world > 1:
        hb = ctx.region_halo_bytes()
        halo = {k: torch.empty(hb // 4, dtype=torch.int32, device="cuda") for k in ("send_lo", "send_hi", "recv_lo", "recv_hi")}

    from ballxworld_b200.slabs import exchange_halos, slab_partition

    slab_info = slab_partition(nx * world, world, rank)

    def exchange():
        """slab-boundary voxel planes to the x-neighbour ranks"""
        exchange_halos(slab_info, dist, halo, lambda f, t: ctx.region_halo_export(f, t.data_ptr()),
                       lambda f, t: ctx.region_halo_import(f, t.data_ptr()))

    def step():
        region.generate(SEED, 64)
        if world > 1:
            exchange()
        return region.mesh()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        arena = step()
    barrier()
    spans = region.spans()
    tot_v = int(spans["v_count"].astype(np.uint64).sum())
    tot_i = int(spans["i_count"].astype(np.uint64).sum())

    # ---------------- device-resident timing ----------------
    for k in (ctx.KERNEL_MESH, ctx.KERNEL_FILL, ctx.KERNEL_HEIGHTMAP):
        ctx.kernel_time_ms(k, reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.kernel_launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        arena = step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - l0
    mesh_ms, mesh_launches = ctx.kernel_time_ms(ctx.KERNEL_MESH)
    fill_ms, fill_launches = ctx.kernel_time_ms(ctx.KERNEL_FILL)
    hmap_ms, hmap_launches = ctx.kernel_time_ms(ctx.KERNEL_HEIGHTMAP)
    meshed_chunks, _ = ctx.region_mesh_candidates()
    # ---------------- mesher streaming rate: the same pass with the chunk summaries ignored (every chunk read) ----------------
    stream_ms = None
    if world == 1:
        ctx.set_option(ctx.OPT_CHUNK_SUMMARIES, 0)
        region.mesh()
        ctx.kernel_time_ms(ctx.KERNEL_MESH, reset=True)
        for _ in range(args.steps):
            region.mesh()
        barrier()
        t_ms, t_n = ctx.kernel_time_ms(ctx.KERNEL_MESH, reset=True)
        stream_ms = t_ms / max(t_n, 1)
        ctx.set_option(ctx.OPT_CHUNK_SUMMARIES, 1)
    # ---------------- the library's fused generate+mesh (meshes pristine terrain from the column parameters) ----------------
    fused_ms = None
    if world == 1:
        ctx.set_option(ctx.OPT_MESH_FROM_HEIGHT_MAP, 1)
        region.generate_and_mesh(SEED, 64)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            region.generate_and_mesh(SEED, 64)
        f1.record(stream)
        barrier()
        fused_ms = f0.elapsed_time(f1) / args.steps
        ctx.set_option(ctx.OPT_MESH_FROM_HEIGHT_MAP, 0)

    # ---------------- end-to-end: public API, host buffers ----------------
    # Every step: generate + mesh on the device, the span table and the whole mesh arena copied to pinned host memory. The
    # arena copy runs on the library's copy stream (bxw_region_mesh_fetch_async) so the next step's generation overlaps it;
    # two host buffers alternate, the timed region ends when the last copy has landed.
    av, ai = arena
    hvs = [torch.empty(max(av, 1) * 40, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    his = [torch.empty(max(ai, 1), dtype=torch.int32, pin_memory=True) for _ in range(2)]

    def e2e_step(k):
        a = region.generate_and_mesh(SEED, 64) if world == 1 else step()
        if a[0] > hvs[0].numel() // 40 or a[1] > his[0].numel():
            raise RuntimeError("arena grew between steps")
        sp = region.spans()
        ctx.region_mesh_fetch_async_ptr(hvs[k & 1].data_ptr(), a[0], his[k & 1].data_ptr(), a[1])
        return a, sp

    e2e_step(0)
    e2e_step(1)
    ctx.region_fetch_wait()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        a, sp = e2e_step(k)
    ctx.region_fetch_wait()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None  # sampled across both timed regions (device-resident and e2e)
    d2h = a[0] * 40 + a[1] * 4 + sp.nbytes + 32
    h2d = 32  # the 32-byte allocator/counter block; generation inputs are the seed and the region box (scalars)

    t = torch.tensor([ms, e2e_s * 1000.0, mesh_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, mesh_ms_max = (float(x) for x in t.tolist())
    tot = torch.tensor([tot_v, tot_i, launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    tot_v_all, tot_i_all, launches_all = (int(x) for x in tot.tolist())

    if rank == 0:
        ms_per_step = ms / args.steps
        value = world * n_chunks / (ms_per_step / 1000.0)
        peak, peak_src = peaks()
        # Rooflines (DESIGN.md "Measurement"), this rank's launches, CUDA events inside the library.
        #  mesh: algorithmic bytes = 131072 per interior chunk + 40 V + 4 I. Chunks proven empty from their summaries are
        #        not read at all, so the figure can exceed the peak; `read_frac` is the same over the chunks actually read.
        #  fill: algorithmic bytes = 131072 per chunk written (the region plus its 1-chunk shell).
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp))
            except Exception:
                traffic = {}

        def roof(kernel, alg_bytes, ms_total, launches_, extra=None):
            per = ms_total / max(launches_, 1)
            ach = alg_bytes / (per / 1000.0) / 1e9 if per > 0 else 0.0
            d = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                 "traffic": (traffic.get(kernel) or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                 "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": per, "share_of_step": per / ms_per_step}
            d.update(extra or {})
            return d

        sx, sy, sz = slab
        storage_chunks = (sx + 2) * (sy + 2) * (sz + 2)
        alg_mesh = n_chunks * 131072 + tot_v * 40 + tot_i * 4
        read_bytes = meshed_chunks * 131072 + tot_v * 40 + tot_i * 4
        mesh_per = mesh_ms / max(mesh_launches, 1)
        roof_mesh = roof("mesh_kernel", alg_mesh, mesh_ms, mesh_launches,  # mesh_kernel + emit_kernel, timed together
                         {"kernels": "mesh_kernel + emit_kernel (one event pair around both)",
                          "note": "chunks proven empty from their summaries are never read, so frac can exceed 1; read_frac counts "
                                  "only the chunks read, streaming is the same pass with the summaries ignored",
                          "chunks_meshed": meshed_chunks, "chunks_skipped_by_summary": n_chunks - meshed_chunks,
                          "read_frac": (read_bytes / (mesh_per / 1000.0) / 1e9 / peak) if mesh_per > 0 else None})
        if stream_ms:
            roof_mesh["streaming"] = {"note": "BXW_OPT_CHUNK_SUMMARIES=0: every chunk read and meshed, same output",
                                      "ms_per_launch": stream_ms, "achieved": alg_mesh / (stream_ms / 1000.0) / 1e9,
                                      "frac": alg_mesh / (stream_ms / 1000.0) / 1e9 / peak,
                                      "traffic": (traffic.get("mesh_kernel_streaming") or {}).get("dram_bytes_per_launch")}
        roof_fill = roof("fill_kernel", storage_chunks * 131072, fill_ms, fill_launches,
                         {"chunks_written": storage_chunks,
                          "note": "pure store stream (column parameters are computed inside the same kernel); the peak is a copy "
                                  "(read+write) measurement, which a write-only stream can exceed"})
        dominant = roof_fill if roof_fill["ms_per_launch"] >= roof_mesh["ms_per_launch"] else roof_mesh
        heightmap = {"kernel": "heightmap_kernel", "bound": "fp64 issue", "columns": (sx + 2) * 32 * (sz + 2) * 32,
                     "ms_per_launch": (hmap_ms / hmap_launches) if hmap_launches else None,
                     "note": "stand-alone launches only; the bench path computes the column parameters inside fill_kernel "
                             "(1.08 ms when run alone, hidden behind the stores)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64+u32", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": world * n_chunks / (e2e_ms / 1000.0 / args.steps), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_all,
            "clocks": clocks,
            "roofline": dominant, "roofline_mesh": roof_mesh, "roofline_fill": roof_fill, "heightmap": heightmap,
            "fused_generate_and_mesh": None if fused_ms is None else {"ms_per_step": fused_ms, "value": n_chunks / (fused_ms / 1000.0), "unit": UNIT,
                                                                      "note": "bxw_region_generate_and_mesh with BXW_OPT_MESH_FROM_HEIGHT_MAP=1: blocks written once, class table derived from the column parameters (the default reads the voxels of the chunks near the surface, which is faster)"},
            "mesh": {"vertices": tot_v_all, "indices": tot_i_all, "arena_vertices": av, "arena_indices": ai},
        }
        if args.slab:
            line["config"]["workload"] += f" [DEBUG slab override {slab}: not the benchmark configuration]"
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    region.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
