#!/usr/bin/env python
"""bench.py — chunks/s generated+meshed (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the hot path over one batch of synthetic input: StdGenerator(seed) terrain for the rank's
slab (shell included) -> [N>1: slab-boundary halo exchange over NCCL] -> mesh every interior chunk.
  value : chunks/s, device-resident (CUDA events on the library's stream, max over ranks), whole job
  e2e   : the same metric through the public API with HOST buffers: generate+mesh, then the meshes and the span
          table are copied to pinned host memory inside the timed region
N>1: the world is N x 128 chunk columns wide; after the warm-up the ranks re-cut it into slabs of equal measured cost (same world,
same total; `slab_rebalance` in the line), and the timed steps run on that cut.
Extra legs in the same line (skipped with --no-extra-legs): everything_materialised, the streaming mesher figure
(`roofline.frac`), fused_generate_and_mesh, cfg2_mesh_only (N=1), cfg5_edits, and for N>1 multi_gpu_parity, per_rank, cfg4_literal.
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
        # the slice's share of chunks that are not skipped by is_chunk_trivial (the rate depends on it): compare with the
        # region's roofline_mesh.default_pass.chunks_meshed / chunks_per_gpu
        "sample_chunks": int(r.chunks), "sample_nontrivial_chunks": int(r.nontrivial),
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
                    f"{', slab widths re-cut from measured per-slab cost (same world, same total), NCCL halo exchange of slab-boundary voxel planes' if n > 1 else ''}); BASELINE.json configs[2] per GPU",
        "chunks_per_gpu": nx * ny * nz, "l2": "inputs larger than L2 (32 GiB of blocks per GPU per step)", "precision": "f64 noise",
    }


def region_totals(region):
    import numpy as np

    spans = region.spans()
    return int(spans["v_count"].astype(np.uint64).sum()), int(spans["i_count"].astype(np.uint64).sum()), spans


def slab_boundary_parity(ctx, region, reg, slab_x0, nx, ny, nz, rank, world, dist, halo, exchange):
    """In-run multi-GPU parity: every rank places blocks on ITS OWN first voxel plane (local x = 0),
    the planes travel through the regular halo exchange, and the x- neighbour's boundary chunks -- whose +x halo now differs
    from pristine terrain -- are re-meshed and compared with the oracle bit for bit. Without a working exchange the meshes
    cannot match (the check asserts that the edits change the expected mesh)."""
    import numpy as np
    import oracle_lib as O

    import ballxworld_b200 as bxw

    g = O.StdGen(SEED)
    zs = [5, 17, 40, 77] if nz >= 3 else [5, 17]
    def edits_of(r):  # deterministic: blocks stacked on the surface of column (x0, z) of rank r's slab
        x = r * nx * 32
        out = []
        for z in zs:
            h = g.params(x, z).height
            for dy in (1, 2):
                if 0 <= h + dy < ny * 32:
                    out.append((x, h + dy, z, 0, (4 << 16)))
        return out
    if rank > 0:
        region.apply_voxel_changes(edits_of(rank))
    exchange()
    region.mesh()  # whole pass: the neighbour-facing boundary chunks read the imported plane
    checked, ok, nonvacuous = 0, True, False
    if rank < world - 1:
        nb_edits = edits_of(rank + 1)
        spans = region.spans()
        V, I = region.fetch()
        oreg = O.OracleRegistry()
        todo = sorted({(nx - 1, (e[1] // 32), e[2] // 32) for e in nb_edits})[:6]
        for (ix, iy, iz) in todo:
            cx = slab_x0 + ix
            d, d0 = [], []
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        c = g.generate_chunk(cx + dx, iy + dy, iz + dz)
                        d0.append(c)
                        c = c.copy()
                        for (ex, ey, ez, _, to) in nb_edits + (edits_of(rank) if rank > 0 else []):
                            if ex // 32 == cx + dx and ey // 32 == iy + dy and ez // 32 == iz + dz:
                                c[(ex % 32) + 32 * (ez % 32) + 1024 * (ey % 32)] = to
                        d.append(c)
            wv, wi = O.mesh_from_dense27(oreg, d)
            pv, _ = O.mesh_from_dense27(oreg, d0)
            nonvacuous |= pv.tobytes() != wv.tobytes()
            b = region.chunk_buffers(spans, V, I, ix, iy, iz)
            ok &= b.vertices.tobytes() == wv.tobytes() and np.array_equal(b.indices, wi)
            checked += 1
    return {"boundary_chunks_checked": checked, "bit_exact": bool(ok), "edits_changed_expected_mesh": bool(nonvacuous)}


def edit_sweep(ctx, region, rank, world, dist, nx, ny, nz, steps=64, edits=4096, radius=10):
    """BASELINE.json configs[4] (SURVEY.md section 8d cfg 5): the streaming view-distance sweep. The region starts unloaded; a
    load anchor (radius 10 chunks, src/config.rs:35-36) sweeps +x one chunk per step. Every step (a) the device-side scheduler
    rebuilds the load deltas (World::recalculate_load_deltas) and processes them -- block data of chunks that entered the
    sphere is generated, their meshes follow on the second tick, what left the sphere is unloaded -- and (b) this rank draws
    `edits` random block placements inside its anchor sphere; the changes that fall into a neighbour rank's storage box (its
    interior or its 1-chunk shell) are SENT to that rank (point-to-point, counts first), every rank applies [lower rank's |
    own | higher rank's] through World::apply_voxel_changes semantics and re-meshes its dirty chunks that have a mesh.
    Latencies = host wall clock, device synchronised."""
    import numpy as np
    import torch

    import ballxworld_b200 as bxw

    lat, lat_tick, dirty, loads = [], [], 0, np.zeros(3, dtype=np.int64)
    def anchor_at(step):
        return (rank * nx + radius + (step % max(1, nx - 2 * radius)), ny // 4, nz // 2, radius, 1)
    region.unload_all()
    for _ in range(3):
        region.tick([anchor_at(0)], SEED, 64)
    st0 = ctx.region_arena_stats()
    peak_used = st0["vertices_used"]
    for step in range(steps):
        rng = np.random.default_rng(1000 * step + rank + 1)
        a = anchor_at(step)
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(2):  # block data of the newly exposed chunks, then (their dependency met) the meshes
            st = region.tick([a], SEED, 64)
            loads += (st["generated"], st["meshed"], st["unmeshed"])
        ctx.synchronize()
        lat_tick.append(time.perf_counter() - t0)
        anchor = np.array(a[:3])
        off = rng.integers(-radius * 32, radius * 32, size=(edits * 2, 3))
        off = off[(off.astype(np.int64) ** 2).sum(axis=1) <= (radius * 32) ** 2][:edits]
        pos = anchor * 32 + off
        ch = np.zeros(len(pos), dtype=bxw.CHANGE_DTYPE)
        ch["x"], ch["y"], ch["z"] = pos[:, 0], pos[:, 1], pos[:, 2]
        ids = rng.integers(1, 8, len(pos)).astype(np.uint32)
        meta = (rng.integers(0, 24, len(pos)) * 64 + rng.integers(0, 4, len(pos))).astype(np.uint32)
        ch["to"] = (ids << 16) | meta
        ch["from"] = 0  # place into air; a change whose target is not air is a failed compare-and-swap that still dirties
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        parts = [ch]
        if dist is not None:
            # changes inside the neighbour's storage box travel to it: [count] then the records (20 bytes each)
            lo_sel = ch[ch["x"] < (rank * nx + 1) * 32] if rank > 0 else ch[:0]
            hi_sel = ch[ch["x"] >= ((rank + 1) * nx - 1) * 32] if rank < world - 1 else ch[:0]
            cap = edits
            def pack(sel):
                t = torch.zeros(1 + cap * 5, dtype=torch.int32)
                t[0] = len(sel)
                if len(sel):
                    t[1:1 + 5 * len(sel)] = torch.from_numpy(np.ascontiguousarray(sel).view(np.int32).reshape(-1))
                return t.cuda()
            ops, recv_lo, recv_hi = [], None, None
            if rank > 0:
                s_lo, recv_lo = pack(lo_sel), torch.empty(1 + cap * 5, dtype=torch.int32, device="cuda")
                ops += [dist.P2POp(dist.isend, s_lo, rank - 1), dist.P2POp(dist.irecv, recv_lo, rank - 1)]
            if rank < world - 1:
                s_hi, recv_hi = pack(hi_sel), torch.empty(1 + cap * 5, dtype=torch.int32, device="cuda")
                ops += [dist.P2POp(dist.isend, s_hi, rank + 1), dist.P2POp(dist.irecv, recv_hi, rank + 1)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            def unpack(t):
                arr = t.cpu().numpy()
                return arr[1:1 + 5 * int(arr[0])].view(bxw.CHANGE_DTYPE).copy()
            parts = ([unpack(recv_lo)] if recv_lo is not None else []) + [ch] + ([unpack(recv_hi)] if recv_hi is not None else [])
        allch = np.concatenate(parts)
        nd = ctx.region_apply_changes(allch)
        nr = ctx.region_remesh_dirty()
        ctx.synchronize()
        lat.append(time.perf_counter() - t0)
        dirty += nr
        assert nr <= nd
        if step % 8 == 7:
            peak_used = max(peak_used, ctx.region_arena_stats()["vertices_used"])
    st1 = ctx.region_arena_stats()
    live = int(region.spans()["v_count"].astype(np.int64).sum())
    lat_t = torch.tensor([lat, lat_tick], dtype=torch.float64, device="cuda")
    d_t = torch.tensor([dirty] + loads.tolist(), dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(lat_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(d_t)
    lat = np.array(lat_t[0].tolist()[4:])
    lat_tick = np.array(lat_t[1].tolist()[4:])
    tot = [int(x) for x in d_t.tolist()]
    return {"workload": f"{steps} steps: a radius-{radius} load anchor per GPU sweeps +x over the bench region (loads / unloads by the device-side "
                        f"scheduler, two ticks per step), then {edits} edits per GPU inside the sphere and a re-mesh of the dirty chunks; "
                        + ("changes near a slab boundary are sent to the neighbour rank" if dist is not None else "one GPU"),
            "dirty_chunks_remeshed": tot[0], "dirty_chunks_per_s": tot[0] / (float(lat.sum()) * steps / len(lat)),
            "edit_to_mesh_latency_ms_median": float(np.median(lat) * 1e3), "edit_to_mesh_latency_ms_p95": float(np.percentile(lat, 95) * 1e3),
            "chunks_generated_by_load_deltas": tot[1], "chunks_meshed_by_load_deltas": tot[2], "meshes_unloaded": tot[3],
            "load_deltas_per_s": (tot[1] + tot[2] + tot[3]) / (float(lat_tick.sum()) * steps / len(lat_tick)),
            "anchor_move_latency_ms_median": float(np.median(lat_tick) * 1e3), "anchor_move_latency_ms_p95": float(np.percentile(lat_tick, 95) * 1e3),
            "arena_vertices_first_sphere": st0["vertices_used"], "arena_vertices_after": st1["vertices_used"],
            "arena_vertices_peak": max(peak_used, st1["vertices_used"]), "arena_garbage_after": st1["vertices_garbage"],
            "live_vertices_after": live}


def mesh_only_cfg2(ctx, reg, steps, warmup, peak, traffic=None):
    """BASELINE.json configs[1] (SURVEY.md section 8d cfg 2): 4096 chunks of hashed random blocks incl. oriented slopes / corner
    slopes, mesh only, one B200; totals checked against the oracle-made golden fixture."""
    import numpy as np

    out = {}
    gold = None
    gp = os.path.join(ROOT, "tests", "golden", "cfg2_golden.json")
    if os.path.exists(gp):
        gold = json.load(open(gp))
    reg.bind(ctx)
    ctx.region_create(0, 0, 0, 16, 16, 16)
    for thr, name in ((128, "dense"), (250, "sparse")):
        ctx.region_fill_synthetic(1, thr)
        for _ in range(warmup):
            ctx.region_mesh()
        ctx.kernel_time_ms(ctx.KERNEL_MESH, reset=True)
        for _ in range(steps):
            ctx.region_mesh()
        ems, _ = ctx.kernel_time_ms(ctx.KERNEL_EMIT)
        ms, n = ctx.kernel_time_ms(ctx.KERNEL_MESH, reset=True)
        spans = ctx.region_mesh_spans()
        V, I = int(spans["v_count"].astype(np.uint64).sum()), int(spans["i_count"].astype(np.uint64).sum())
        alg = 4096 * 131072 + 40 * V + 4 * I
        per = ms / max(n, 1)
        d = {"void_threshold": thr, "chunks": 4096, "vertices": V, "indices": I, "ms": per, "algorithmic_bytes": alg,
             "achieved": alg / (per / 1e3) / 1e9, "unit": "GB/s", "peak": peak, "frac": alg / (per / 1e3) / 1e9 / peak,
             "chunks_per_s": 4096 / (per / 1e3), "mesh_kernel_ms": (ms - ems) / max(n, 1), "emit_kernel_ms": ems / max(n, 1)}
        if thr == 128 and traffic:
            d["traffic"] = (traffic.get("cfg2_mesh_only") or {}).get("dram_bytes_per_launch")
        if gold:
            gv = gold["variants"][str(thr)]
            d["totals_match_oracle_golden"] = (V == gv["total_vertices"] and I == gv["total_indices"])
        out[name] = d
    ctx.region_destroy()
    return out


def cfg4_literal(ctx, reg, rank, world, dist, steps, stream):
    """BASELINE.json configs[3] as stated: the 512x512x16 chunk region sharded by x-slabs over the ranks. A rank's 512/N chunk
    columns are processed in boxes of 64 columns (80 GB of blocks each, one allocation re-positioned with
    bxw_region_set_origin: generate -> halo exchange -> mesh -> retire). Even ranks walk their boxes towards +x, odd ranks
    towards -x, so both ranks of a slab boundary hold their boundary box in the same pass and exchange its voxel planes."""
    import numpy as np
    import torch

    import ballxworld_b200 as bxw
    from ballxworld_b200.slabs import finish_exchange, start_exchange

    TOTAL_X, BOX, NY, NZ = 512, 64, 16, 512
    per_rank = TOTAL_X // world
    passes = per_rank // BOX
    region = bxw.Region((rank * per_rank, 0, 0), (BOX, NY, NZ), reg, ctx=ctx)
    hb = ctx.region_halo_bytes()
    halo = {k: torch.empty(hb // 4, dtype=torch.int32, device="cuda") for k in ("send_lo", "send_hi", "recv_lo", "recv_hi")}
    order = list(range(passes)) if rank % 2 == 0 else list(range(passes - 1, -1, -1))

    class PassSlab:  # which faces of this pass's box are slab boundaries that have a partner holding the other side NOW
        def __init__(self, k, b):
            self.rank, self.world = rank, world
            first, last = (k == 0), (k == passes - 1)
            # even ranks: first pass = leftmost box (partner rank-1, odd, starts rightmost), last pass = rightmost box
            if rank % 2 == 0:
                self.has_lo = first and b == 0 and rank > 0
                self.has_hi = last and b == passes - 1 and rank < world - 1
            else:
                self.has_hi = first and b == passes - 1 and rank < world - 1
                self.has_lo = last and b == 0 and rank > 0

    def run_pass(k, b):
        region.set_origin((rank * per_rank + b * BOX, 0, 0))
        sl = PassSlab(k, b)
        if sl.has_lo or sl.has_hi:
            region.generate_part(1, SEED, 64)
            works = start_exchange(sl, dist, halo, lambda f, t: ctx.region_halo_export(f, t.data_ptr()))
            region.generate_part(2, SEED, 64)
            finish_exchange(sl, works, halo, lambda f, t: ctx.region_halo_import(f, t.data_ptr()))
        else:
            region.generate(SEED, 64)
        return region.mesh()

    def run_all(collect):
        tv = ti = 0
        for k, b in enumerate(order):
            run_pass(k, b)
            if collect:
                v, i, _ = region_totals(region)
                tv += v
                ti += i
        return tv, ti

    tv, ti = run_all(True)  # warm-up + totals
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        run_all(False)
    e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([tv, ti], dtype=torch.int64, device="cuda")
    per_rank_tot = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(per_rank_tot, tot)
    dist.all_reduce(tot)
    region.close()
    ms = float(t.item())
    n_chunks = TOTAL_X * NY * NZ
    out = {"workload": f"512x512x16 chunk region, seed {SEED}, x-slabs of {per_rank} columns per GPU in {passes} box(es) of {BOX}x{NZ}x{NY}, "
                       "NCCL exchange of the slab-boundary voxel planes (33.7 MB per direction per boundary)",
           "scaling": "strong", "chunks": n_chunks, "ms_per_step": ms, "value": n_chunks / (ms / 1e3), "unit": UNIT,
           "halo_plane_bytes": hb, "total_vertices": int(tot[0].item()), "total_indices": int(tot[1].item()),
           "per_rank_totals": [[int(x[0].item()), int(x[1].item())] for x in per_rank_tot]}
    gp = os.path.join(ROOT, "tests", "golden", "cfg4_totals.json")
    if os.path.exists(gp):
        g = json.load(open(gp))
        out["totals_match_golden"] = (out["total_vertices"] == g["total_vertices"] and out["total_indices"] == g["total_indices"])
        out["golden_source"] = g.get("source")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--slab", default=None, help="override nx,ny,nz interior chunks per rank (debug only; invalidates the number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rebalance", action="store_true", help="N>1: keep equal slab widths (default: widths follow the measured per-slab cost)")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip cfg2 / cfg4 / cfg5 legs (profiling runs)")
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
    widths = [nx] * world  # chunk columns per rank; N>1 re-cuts them below from the measured per-slab cost
    total_chunks = world * nx * ny * nz

    halo = None
    if world > 1:
        hb = ctx.region_halo_bytes()
        halo = {k: torch.empty(hb // 4, dtype=torch.int32, device="cuda") for k in ("send_lo", "send_hi", "recv_lo", "recv_hi")}

    from ballxworld_b200.slabs import agree_on_widths, exchange_halos, finish_exchange, slab_partition, start_exchange

    slab_info = slab_partition(nx * world, world, rank)

    def exchange():
        """slab-boundary voxel planes to the x-neighbour ranks"""
        exchange_halos(slab_info, dist, halo, lambda f, t: ctx.region_halo_export(f, t.data_ptr()),
                       lambda f, t: ctx.region_halo_import(f, t.data_ptr()))

    def step():
        if world > 1:
            # boundary chunk columns first, their planes are packed and sent while the rest of the slab is generated
            region.generate_part(1, SEED, 64)
            works = start_exchange(slab_info, dist, halo, lambda f, t: ctx.region_halo_export(f, t.data_ptr()))
            region.generate_part(2, SEED, 64)
            finish_exchange(slab_info, works, halo, lambda f, t: ctx.region_halo_import(f, t.data_ptr()))
        else:
            region.generate(SEED, 64)
        return region.mesh()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            fn()
        b.record(stream)
        barrier()
        return a.elapsed_time(b) / n

    for _ in range(args.warmup):
        arena = step()
    barrier()
    tot_v, tot_i, spans = region_totals(region)

    # ---------------- N>1: slab widths from measured cost ----------------
    # The slabs hold different terrain (mesh sizes differ by up to 20 %), and every step couples a rank to its neighbours, so
    # the step time is the heaviest slab's. Each rank measures the device time of its own kernels, the ranks exchange the
    # scalars and all compute the same new cut of the SAME world (total columns unchanged); the best cut measured is kept.
    rebalance = None
    if world > 1 and not args.no_rebalance:
        ut = torch.tensor([tot_v, tot_i], dtype=torch.int64, device="cuda")
        dist.all_reduce(ut)
        history, best = [], None
        CUTS, MEASURE_STEPS = 6, 3
        for it in range(CUTS):
            for k in (ctx.KERNEL_MESH, ctx.KERNEL_FILL):
                ctx.kernel_time_ms(k, reset=True)
            for _ in range(MEASURE_STEPS):
                step()
            barrier()
            cost = (ctx.kernel_time_ms(ctx.KERNEL_MESH, reset=True)[0] + ctx.kernel_time_ms(ctx.KERNEL_FILL, reset=True)[0]) / MEASURE_STEPS
            new_w, costs = agree_on_widths(dist, widths, cost, "cuda", min_width=32)
            history.append({"widths": list(widths), "own_kernel_ms": [round(c, 3) for c in costs]})
            if best is None or max(costs) < best[0]:
                best = (max(costs), list(widths))
            done = it == CUTS - 1 or max(costs) < 1.005 * sum(costs) / world or new_w == widths
            target = best[1] if done else new_w  # at the end: the best cut that was measured
            if target != widths:
                widths = target
                region.close()
                region = bxw.Region((sum(widths[:rank]), 0, 0), (widths[rank], ny, nz), reg, ctx=ctx)
                for _ in range(2):
                    arena = step()
                barrier()
            if done:
                break
        tot_v, tot_i, spans = region_totals(region)
        rt = torch.tensor([tot_v, tot_i], dtype=torch.int64, device="cuda")
        dist.all_reduce(rt)
        rebalance = {"widths": list(widths), "measured": history, "world_totals_unchanged": bool((rt == ut).all().item()),
                     "how": "per-rank device time of fill + mesh kernels over 3 steps, one scalar all_gather, every rank computes the same "
                            "new cut (slabs.rebalance_slabs); total columns and the global mesh are unchanged"}
        if not rebalance["world_totals_unchanged"]:
            raise RuntimeError("rebalanced slabs changed the global mesh totals")
    n_chunks = widths[rank] * ny * nz  # this rank's interior chunks

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
    mesh_ms, mesh_launches = ctx.kernel_time_ms(ctx.KERNEL_MESH, reset=True)
    fill_ms, fill_launches = ctx.kernel_time_ms(ctx.KERNEL_FILL, reset=True)
    meshed_chunks, _ = ctx.region_mesh_candidates()
    pages = ctx.region_download_pages()
    storage_chunks = (widths[rank] + 2) * (ny + 2) * (nz + 2)
    materialised_chunks = int((pages == np.arange(pages.size, dtype=np.uint32)).sum())
    side = min(args.steps, 3)
    # ---------------- the same step with every chunk materialised (BXW_OPT_ELIDE_UNIFORM=0): round-1's figure ----------------
    ctx.set_option(ctx.OPT_ELIDE_UNIFORM, 0)
    step()
    ctx.kernel_time_ms(ctx.KERNEL_FILL, reset=True)
    mat_ms = timed(step, side)
    mat_fill_ms, mat_fill_n = ctx.kernel_time_ms(ctx.KERNEL_FILL, reset=True)
    # ---------------- mesher streaming rate: every chunk materialised AND read (summaries ignored): the honest HBM figure ----------------
    ctx.set_option(ctx.OPT_CHUNK_SUMMARIES, 0)
    region.mesh()
    ctx.kernel_time_ms(ctx.KERNEL_MESH, reset=True)
    for _ in range(side):
        region.mesh()
    barrier()
    t_ms, t_n = ctx.kernel_time_ms(ctx.KERNEL_MESH, reset=True)
    stream_ms = t_ms / max(t_n, 1)
    ctx.set_option(ctx.OPT_CHUNK_SUMMARIES, 1)
    ctx.set_option(ctx.OPT_ELIDE_UNIFORM, 1)
    # ---------------- the library's fused generate+mesh (meshes pristine terrain from the column parameters) ----------------
    fused_ms = None
    if world == 1:
        ctx.set_option(ctx.OPT_MESH_FROM_HEIGHT_MAP, 1)
        region.generate_and_mesh(SEED, 64)
        fused_ms = timed(lambda: region.generate_and_mesh(SEED, 64), side)
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
    # the node's D2H ceiling for this many GPUs copying at once: plain cudaMemcpyAsync of the same pinned buffers
    barrier()
    tc0 = time.perf_counter()
    for k in range(2):
        ctx.region_mesh_fetch_async_ptr(hvs[k & 1].data_ptr(), a[0], his[k & 1].data_ptr(), a[1])
        ctx.region_fetch_wait()
    barrier()
    copy_s = (time.perf_counter() - tc0) / 2
    del hvs, his

    t = torch.tensor([ms, e2e_s * 1000.0, mesh_ms, mat_ms, stream_ms, copy_s * 1000.0], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        # per-rank view of the same timed region: this rank's device time per step, its mesher time and its mesh size (the slabs
        # hold different terrain, so the ranks do not carry the same mesh work)
        mine = torch.tensor([ms / args.steps, mesh_ms / max(mesh_launches, 1), fill_ms / max(fill_launches, 1), float(tot_v)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros(4, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"columns": widths[i], "ms_per_step": float(a[0]), "mesh_ms": float(a[1]), "fill_ms": float(a[2]), "vertices": int(a[3])}
                    for i, a in enumerate(allr)]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, mesh_ms_max, mat_ms, stream_ms_max, copy_ms = (float(x) for x in t.tolist())
    tot = torch.tensor([tot_v, tot_i, launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    tot_v_all, tot_i_all, launches_all = (int(x) for x in tot.tolist())

    # ---------------- extra legs: parity across slab boundaries, edit sweep, cfg2 mesh-only, cfg4 as stated ----------------
    parity = edits_leg = cfg2 = cfg4 = None
    peak, peak_src = peaks()
    if not args.no_extra_legs:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        if world > 1:
            if widths != [nx] * world:  # the legs below address slabs as rank * nx: back to the equal cut
                region.close()
                region = bxw.Region((rank * nx, 0, 0), (nx, ny, nz), reg, ctx=ctx)
            region.generate(SEED, 64)
            par = slab_boundary_parity(ctx, region, reg, rank * nx, nx, ny, nz, rank, world, dist, halo, exchange)
            pt = torch.tensor([par["boundary_chunks_checked"], int(par["bit_exact"]), int(par["edits_changed_expected_mesh"] or rank == world - 1)],
                              dtype=torch.int64, device="cuda")
            mn = pt.clone()
            dist.all_reduce(pt, op=dist.ReduceOp.SUM)
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            parity = {"boundary_chunks_checked": int(pt[0].item()), "bit_exact": bool(mn[1].item()),
                      "edits_changed_expected_mesh": bool(mn[2].item()),
                      "how": "every rank places blocks on its own x- voxel plane, the planes go through the halo exchange, the x- "
                             "neighbour's boundary chunks are meshed and compared with the oracle (vertices and indices, bytewise)"}
        edits_leg = edit_sweep(ctx, region, rank, world, dist, nx, ny, nz)
    region.close()
    if not args.no_extra_legs:
        if world == 1:
            tj = {}
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            except Exception:
                pass
            cfg2 = mesh_only_cfg2(ctx, reg, args.steps, args.warmup, peak, tj)
        elif 512 % (64 * world) == 0:
            cfg4 = cfg4_literal(ctx, reg, rank, world, dist, min(args.steps, 2), stream)

    if rank == 0:
        ms_per_step = ms / args.steps
        value = total_chunks / (ms_per_step / 1000.0)
        # Rooflines (DESIGN.md "Measurement"), this rank's launches, CUDA events inside the library.
        #  mesh: algorithmic bytes = 131072 per interior chunk + 40 V + 4 I.
        #  fill: algorithmic bytes = 131072 per chunk written.
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp))
            except Exception:
                traffic = {}

        def gbs(nbytes, per_ms):
            return nbytes / (per_ms / 1000.0) / 1e9 if per_ms and per_ms > 0 else 0.0

        alg_mesh = n_chunks * 131072 + tot_v * 40 + tot_i * 4
        read_bytes = meshed_chunks * 131072 + tot_v * 40 + tot_i * 4
        mesh_per = mesh_ms / max(mesh_launches, 1)
        fill_per = fill_ms / max(fill_launches, 1)
        mat_fill_per = mat_fill_ms / max(mat_fill_n, 1)
        # The mesher pair (mesh_kernel + emit_kernel, one event pair around both). `frac` is the STREAMING figure: every chunk
        # materialised and read (summaries and elision off), i.e. the kernel doing all the work the formula counts.
        roof_mesh = {
            "bound": "hbm", "kernel": "mesh_kernel", "kernels": "mesh_kernel + emit_kernel (one event pair around both)",
            "achieved": gbs(alg_mesh, stream_ms), "peak": peak, "unit": "GB/s", "frac": gbs(alg_mesh, stream_ms) / peak,
            "traffic": (traffic.get("mesh_kernel_streaming") or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_mesh, "ms_per_launch": stream_ms,
            "what": "streaming pass: BXW_OPT_ELIDE_UNIFORM=0 and BXW_OPT_CHUNK_SUMMARIES=0, every chunk read from HBM and meshed; "
                    "same output as the default pass",
            "default_pass": {"ms_per_launch": mesh_per, "share_of_step": mesh_per / ms_per_step, "chunks_meshed": meshed_chunks,
                             "chunks_skipped_by_summary": n_chunks - meshed_chunks,
                             "frac_with_summaries": gbs(alg_mesh, mesh_per) / peak,
                             "read_frac": gbs(read_bytes, mesh_per) / peak,
                             "traffic": (traffic.get("mesh_kernel") or {}).get("dram_bytes_per_launch"),
                             "note": "chunks proven empty from their summaries are never read, so frac_with_summaries exceeds 1; "
                                     "read_frac = the same formula over the chunks actually read"},
        }
        roof_fill = {
            "bound": "hbm", "kernel": "fill_kernel", "achieved": gbs(storage_chunks * 131072, mat_fill_per), "peak": peak, "unit": "GB/s",
            "frac": gbs(storage_chunks * 131072, mat_fill_per) / peak, "traffic": (traffic.get("fill_kernel") or {}).get("dram_bytes_per_launch"),
            "peak_source": peak_src, "algorithmic_bytes_per_launch": storage_chunks * 131072, "ms_per_launch": mat_fill_per,
            "what": "BXW_OPT_ELIDE_UNIFORM=0: all chunks written voxel by voxel (pure store stream; the peak is a copy measurement, "
                    "which a write-only stream can exceed)",
            "default_pass": {"ms_per_launch": fill_per, "share_of_step": fill_per / ms_per_step, "chunks_written": materialised_chunks,
                             "chunks_on_shared_pages": storage_chunks - materialised_chunks,
                             "achieved": gbs(materialised_chunks * 131072, fill_per),
                             "note": "uniform chunks are not materialised; the FP64 column parameters (about 1.1 ms when run alone) "
                                     "are no longer hidden behind 40 GB of stores"},
        }
        dominant = roof_mesh if mesh_per >= fill_per else roof_fill
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64+u32", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": total_chunks / (e2e_ms / 1000.0 / args.steps), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "d2h_gbs_per_gpu": d2h / (e2e_ms / 1000.0 / args.steps) / 1e9,
                    "host_limit_gbs_per_gpu": d2h / (copy_ms / 1000.0) / 1e9,
                    "note": "host_limit = the same pinned buffers filled by plain cudaMemcpyAsync from all N GPUs at once, no compute"},
            "gpu_launches": launches_all,
            "clocks": clocks,
            "parity_note": "generator: bit-exact vs oracle/ (C restatement) and vs an independent Python transliteration; the real "
                           "`noise 0.8.2` crate cannot be built here (no cargo), so generator parity is UNPINNED; RNG crates pinned by "
                           "their published vectors; mesher / RLE / orientations restated from in-tree source",
            "roofline": dominant, "roofline_mesh": roof_mesh, "roofline_fill": roof_fill,
            "everything_materialised": {"ms_per_step": mat_ms, "value": total_chunks / (mat_ms / 1000.0), "unit": UNIT,
                                        "note": "BXW_OPT_ELIDE_UNIFORM=0 (round 1's storage): every chunk written voxel by voxel"},
            "fused_generate_and_mesh": None if fused_ms is None else {"ms_per_step": fused_ms, "value": n_chunks / (fused_ms / 1000.0), "unit": UNIT,
                                                                      "note": "bxw_region_generate_and_mesh with BXW_OPT_MESH_FROM_HEIGHT_MAP=1 (phase A derives the 34^3 table from the column "
                                                                              "parameters instead of reading voxels; since the cubes-only kernel of round 2 it is SLOWER than the "
                                                                              "default step and stays an option)"},
            "mesh": {"vertices": tot_v_all, "indices": tot_i_all, "arena_vertices": av, "arena_indices": ai},
            "per_rank": per_rank, "slab_rebalance": rebalance, "multi_gpu_parity": parity, "cfg5_edits": edits_leg, "cfg2_mesh_only": cfg2, "cfg4_literal": cfg4,
        }
        if args.slab:
            line["config"]["workload"] += f" [DEBUG slab override {slab}: not the benchmark configuration]"
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
