#!/usr/bin/env python
"""Randomised parity soak: random regions (terrain or hashed blocks), random edits / uploads / save-blob imports, whole-region
and dirty re-meshes, every sampled chunk compared bit for bit with the oracle. Not part of the test suite (minutes);
run after kernel changes:  python tools/stress_parity.py --seconds 120"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import ballxworld_b200 as bxw  # noqa: E402
import oracle_lib as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    ctx = bxw.Context(0)
    reg = bxw.standard_registry()
    reg.bind(ctx)
    ctx.set_gen_ids(*bxw.StdGenerator.resolve_ids(reg))
    oreg = O.OracleRegistry()
    t_end = time.time() + a.seconds
    rounds = checked = 0
    while time.time() < t_end:
        rounds += 1
        dims = tuple(int(x) for x in rng.integers(1, 5, 3))
        origin = (int(rng.integers(-40, 40)), int(rng.integers(-2, 3)), int(rng.integers(-40, 40)))
        ctx.region_create(*origin, *dims)
        terrain = rng.random() < 0.6
        if terrain:
            ctx.set_option(ctx.OPT_MESH_FROM_HEIGHT_MAP, int(rng.random() < 0.3))
            ctx.region_generate_and_mesh(int(rng.integers(0, 1 << 40)), 64)
            ctx.set_option(ctx.OPT_MESH_FROM_HEIGHT_MAP, 0)
        else:
            ctx.region_fill_synthetic(int(rng.integers(1, 1 << 30)), int(rng.choice([16, 128, 200, 250])))
            ctx.region_mesh()
        chunks = {(lx, ly, lz): ctx.region_download_chunk(lx, ly, lz) for ly in range(-1, dims[1] + 1)
                  for lz in range(-1, dims[2] + 1) for lx in range(-1, dims[0] + 1)}
        x0, y0, z0 = (origin[0] - 1) * 32, (origin[1] - 1) * 32, (origin[2] - 1) * 32
        W = ((dims[0] + 2) * 32, (dims[1] + 2) * 32, (dims[2] + 2) * 32)

        def check(which):
            nonlocal checked
            spans = ctx.region_mesh_spans()
            av = int((spans["v_off"] + spans["v_count"]).max())
            ai = int((spans["i_off"] + spans["i_count"]).max())
            V, I = ctx.region_mesh_fetch(max(av, 1), max(ai, 1))
            for (ix, iy, iz) in which:
                d = [chunks[(ix + dx, iy + dy, iz + dz)] for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1)]
                wv, wi = O.mesh_from_dense27(oreg, d)
                s = spans[ix + dims[0] * (iz + dims[2] * iy)]
                gv = V[int(s["v_off"]): int(s["v_off"]) + int(s["v_count"])]
                gi = I[int(s["i_off"]): int(s["i_off"]) + int(s["i_count"])]
                if gv.tobytes() != wv.tobytes() or not np.array_equal(gi, wi):
                    raise SystemExit(f"MISMATCH round {rounds} origin {origin} dims {dims} terrain {terrain} chunk {(ix, iy, iz)}: "
                                     f"{len(gv)} vs {len(wv)} vertices")
                checked += 1

        interior = [(x, y, z) for x in range(dims[0]) for y in range(dims[1]) for z in range(dims[2])]
        sample = [interior[i] for i in rng.permutation(len(interior))[:4]]
        check(sample)
        # edits: random voxels of the storage box (shell included), half of them with a matching `from`
        for _ in range(int(rng.integers(1, 4))):
            n = int(rng.integers(1, 200))
            pos = np.stack([rng.integers(0, W[k], n) for k in range(3)], axis=1)
            if rng.random() < 0.5:  # cluster them around a chunk corner so boundaries get hit
                c = np.array([rng.integers(1, dims[k] + 1) * 32 for k in range(3)])
                pos = np.clip(c + rng.integers(-2, 2, (n, 3)), 0, np.array(W) - 1)
            ch = np.zeros(n, dtype=bxw.CHANGE_DTYPE)
            dirty = set()
            for k, (x, y, z) in enumerate(pos):
                key = (int(x) // 32 - 1, int(y) // 32 - 1, int(z) // 32 - 1)
                idx = (int(x) & 31) + 32 * (int(z) & 31) + 1024 * (int(y) & 31)
                cur = int(chunks[key][idx])
                to = 0 if rng.random() < 0.3 else (int(rng.integers(1, 8)) << 16) | (int(rng.integers(0, 24)) * 64 + int(rng.integers(0, 4)))
                frm = cur if rng.random() < 0.7 else cur ^ 0x10000
                ch[k] = (x0 + int(x), y0 + int(y), z0 + int(z), frm, to)
                if frm == cur:
                    chunks[key] = chunks[key].copy()
                    chunks[key][idx] = to
                lx, ly, lz = int(x) & 31, int(y) & 31, int(z) & 31
                cand = [key]
                for axis, l in enumerate((lx, ly, lz)):  # touching_chunks (lib.rs:110-135)
                    if l in (0, 31):
                        nb = list(key)
                        nb[axis] += -1 if l == 0 else 1
                        cand.append(tuple(nb))
                for c in cand:
                    if all(0 <= c[i] < dims[i] for i in range(3)):
                        dirty.add(c)
            nd = ctx.region_apply_changes(ch)
            if nd != len(dirty):
                raise SystemExit(f"DIRTY COUNT round {rounds}: {nd} vs {len(dirty)}")
            if rng.random() < 0.5:
                assert ctx.region_remesh_dirty() == nd
                check(sorted(dirty)[:6])
            else:
                ctx.region_mesh()
                check(sorted(dirty)[:4] + sample[:2])
        # an upload and a save-blob import, then a whole-region pass
        k = interior[int(rng.integers(0, len(interior)))]
        newc = O.random_chunk(*[int(v) for v in rng.integers(0, 50, 3)], seed=int(rng.integers(1, 99)), void_threshold=int(rng.choice([64, 200])))
        if rng.random() < 0.5:
            ctx.region_upload_chunk(*k, newc)
        else:
            w = O.compress_rle(newc)
            ctx.region_import_rle([k], w, np.array([0, len(w)], np.uint64))
        chunks[k] = newc
        ctx.region_mesh()
        nbrs = [(k[0] + dx, k[1] + dy, k[2] + dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)]
        check([c for c in nbrs if all(0 <= c[i] < dims[i] for i in range(3))][:8])
        for c in interior[:3]:  # storage itself
            assert np.array_equal(ctx.region_download_chunk(*c), chunks[c])
        ctx.region_destroy()
    print(f"stress ok: {rounds} regions, {checked} chunk meshes bit-exact")
    ctx.close()


if __name__ == "__main__":
    main()
