"""Python restatement of World::recalculate_load_deltas (bxw_world/src/worldmgr.rs:633-790) for a bounded box of chunks --
TEST INFRASTRUCTURE for bxw_region_update_anchors. Two handler kinds: block data (no dependency) and mesh (depends on the
block data of the chunk and its 26 neighbours, worldmgr.rs:761-790, voxrender.rs get_dependency)."""
import itertools

KIND_BLOCKS, KIND_MESH = 1, 2


def iter_coords_to_load(anchor):  # worldmgr.rs:734-742
    cx, cy, cz, r, _ = anchor
    for x, y, z in itertools.product(range(-r, r + 1), repeat=3):
        if x * x + y * y + z * z <= r * r:
            yield (cx + x, cy + y, cz + z)


def recalculate_load_deltas(anchors, box_min, box_dims, interior_min, interior_dims, blocks_loaded, mesh_loaded):
    """blocks_loaded / mesh_loaded: sets of chunk positions. Returns (unload, load): lists of (pos, kinds, min_dist) each sorted
    by distance (ties in any order)."""
    def in_box(p):
        return all(box_min[k] <= p[k] < box_min[k] + box_dims[k] for k in range(3))

    def interior(p):
        return all(interior_min[k] <= p[k] < interior_min[k] + interior_dims[k] for k in range(3))

    def d2(p, a):
        return sum((p[k] - a[k]) ** 2 for k in range(3))

    unload = {}
    allocated = set(blocks_loaded) | set(mesh_loaded)
    for p in allocated:  # :636-661
        should = [False, False]
        for a in anchors:
            if d2(p, a) <= a[3] * a[3]:
                should[0] = True
                if a[4]:
                    should[1] = True
        kinds = (KIND_BLOCKS if (p in blocks_loaded and not should[0]) else 0) | (KIND_MESH if (p in mesh_loaded and not should[1]) else 0)
        if kinds:
            unload[p] = (kinds, min(d2(p, a) for a in anchors) if anchors else 0x7FFFFFFF)
    load = {}
    for a in anchors:  # :662-696
        kinds_req = [KIND_BLOCKS] + ([KIND_MESH] if a[4] else [])
        for p in iter_coords_to_load(a):
            if not in_box(p):
                continue  # the region box is the whole world of this context
            for kind in kinds_req:
                if kind == KIND_BLOCKS:
                    missing = p not in blocks_loaded
                    can = True
                else:
                    missing = p not in mesh_loaded
                    can = interior(p) and all((p[0] + dx, p[1] + dy, p[2] + dz) in blocks_loaded
                                              for dy in (-1, 0, 1) for dz in (-1, 0, 1) for dx in (-1, 0, 1))
                if missing and can:
                    dist = d2(p, a)
                    if p in load and load[p][0] & kind:
                        load[p] = (load[p][0], min(load[p][1], dist))
                    elif p in load:
                        load[p] = (load[p][0] | kind, load[p][1])
                    else:
                        load[p] = (kind, dist)
    un = sorted(((p, k, d) for p, (k, d) in unload.items()), key=lambda t: t[2])
    ld = sorted(((p, k, d) for p, (k, d) in load.items()), key=lambda t: t[2])
    return un, ld


def interleave(un, ld):
    """:699-731: unload, load, unload, load, ... (the list is reversed and popped from the back = consumed in this order)."""
    out = []
    for i in range(max(len(un), len(ld))):
        if i < len(un):
            out.append(("unload",) + un[i])
        if i < len(ld):
            out.append(("load",) + ld[i])
    return out
