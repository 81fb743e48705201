"""ctypes binding of csrc/libbxw_cuda.so — one method per entry point of include/bxw_cuda.h."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("BXW_CUDA_LIB") or os.path.join(_HERE, "csrc", "libbxw_cuda.so")  # override: profiling builds

VERTEX_DTYPE = np.dtype(
    [
        ("position", "<u4"),
        ("color", "<u4"),
        ("texcoord", "<f4", (3,)),
        ("index", "<i4"),
        ("barycentric_color_offset", "<f4", (4,)),
    ]
)
SPAN_DTYPE = np.dtype([("v_off", "<u8"), ("i_off", "<u8"), ("v_count", "<u4"), ("i_count", "<u4")])
CHANGE_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("z", "<i4"), ("from", "<u4"), ("to", "<u4")])
ANCHOR_DTYPE = np.dtype([("cx", "<i4"), ("cy", "<i4"), ("cz", "<i4"), ("radius", "<u4"), ("load_mesh", "<u4")])
DELTA_DTYPE = np.dtype([("cx", "<i4"), ("cy", "<i4"), ("cz", "<i4"), ("kinds", "<u4"), ("unload", "<u4"), ("min_anchor_distance", "<i4")])
assert VERTEX_DTYPE.itemsize == 40 and SPAN_DTYPE.itemsize == 24 and CHANGE_DTYPE.itemsize == 20
assert ANCHOR_DTYPE.itemsize == 20 and DELTA_DTYPE.itemsize == 24

ERROR_NAMES = {
    -1: "BXW_E_INVALID",
    -2: "BXW_E_CUDA",
    -3: "BXW_E_NOMEM",
    -4: "BXW_E_UNKNOWN_ID",
    -5: "BXW_E_NOSPACE",
    -6: "BXW_E_BAD_RLE",
    -7: "BXW_E_NO_REGISTRY",
}


class BxwError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


class BlockDef(C.Structure):
    _fields_ = [
        ("present", C.c_uint8),
        ("mesh_kind", C.c_uint8),
        ("debug_color", C.c_float * 3),
        ("tex", C.c_uint32 * 6),
    ]


class ArenaStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("vertices_used", "indices_used", "vertices_garbage", "indices_garbage",
                                          "vertices_capacity", "indices_capacity")]


def library_path():
    return _LIB_PATH


_lib = None


def load_library():
    """Loads the CUDA library. Fails loudly when it has not been built: there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m ballxworld_b200.build` (nvcc, sm_100a). "
            "ballxworld_b200 has no CPU fallback."
        )
    L = C.CDLL(_LIB_PATH)
    vp, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    sig = {
        "bxw_create": ([C.c_int, C.POINTER(vp)], C.c_int),
        "bxw_destroy": ([vp], None),
        "bxw_last_error": ([vp], C.c_char_p),
        "bxw_set_stream": ([vp, vp], C.c_int),
        "bxw_set_option": ([vp, C.c_int, C.c_int], C.c_int),
        "bxw_synchronize": ([vp], C.c_int),
        "bxw_kernel_launches": ([vp], C.c_uint64),
        "bxw_set_registry": ([vp, C.POINTER(BlockDef), C.c_uint32], C.c_int),
        "bxw_set_gen_ids": ([vp] + [C.c_uint16] * 5, C.c_int),
        "bxw_gen_chunk": ([vp, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, vp], C.c_int),
        "bxw_mesh_chunk27": ([vp, C.POINTER(vp), vp], C.c_int),
        "bxw_mesh_chunk27_rle": ([vp, C.POINTER(vp), C.POINTER(C.c_size_t), vp], C.c_int),
        "bxw_mesh_fetch": ([vp, vp, vp], C.c_int),
        "bxw_is_chunk_trivial": ([vp, vp, C.c_size_t, C.POINTER(C.c_int)], C.c_int),
        "bxw_rle_bound": ([C.c_size_t], C.c_size_t),
        "bxw_rle_compress": ([vp, vp, C.c_size_t, vp, vp], C.c_int),
        "bxw_rle_decompress": ([vp, vp, vp, C.c_size_t, vp], C.c_int),
        "bxw_region_create": ([vp, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32], C.c_int),
        "bxw_region_destroy": ([vp], C.c_int),
        "bxw_region_generate": ([vp, C.c_uint64, C.c_int], C.c_int),
        "bxw_region_generate_part": ([vp, C.c_uint64, C.c_int, C.c_int], C.c_int),
        "bxw_region_heightmap": ([vp, vp, vp, vp], C.c_int),
        "bxw_region_upload_chunk": ([vp, C.c_int32, C.c_int32, C.c_int32, vp], C.c_int),
        "bxw_region_download_chunk": ([vp, C.c_int32, C.c_int32, C.c_int32, vp], C.c_int),
        "bxw_region_export_rle": ([vp, vp, C.c_size_t, vp, C.c_size_t, vp], C.c_int),
        "bxw_region_import_rle": ([vp, vp, C.c_size_t, vp, vp], C.c_int),
        "bxw_region_fill_synthetic": ([vp, C.c_uint64, C.c_uint32], C.c_int),
        "bxw_region_mesh_reserve": ([vp, C.c_uint64, C.c_uint64], C.c_int),
        "bxw_region_mesh": ([vp, u64p, u64p], C.c_int),
        "bxw_region_generate_and_mesh": ([vp, C.c_uint64, C.c_int, u64p, u64p], C.c_int),
        "bxw_region_mesh_sphere": ([vp, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, u32p, u64p, u64p], C.c_int),
        "bxw_region_mesh_spans": ([vp, vp], C.c_int),
        "bxw_region_mesh_fetch": ([vp, vp, C.c_uint64, vp, C.c_uint64], C.c_int),
        "bxw_region_mesh_fetch_async": ([vp, vp, C.c_uint64, vp, C.c_uint64], C.c_int),
        "bxw_region_mesh_fetch_range": ([vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, C.c_uint64, vp], C.c_int),
        "bxw_region_fetch_wait": ([vp], C.c_int),
        "bxw_region_apply_changes": ([vp, vp, C.c_uint32, u32p], C.c_int),
        "bxw_region_remesh_dirty": ([vp, u32p], C.c_int),
        "bxw_region_unload_all": ([vp], C.c_int),
        "bxw_region_update_anchors": ([vp, vp, C.c_uint32, u32p, u32p], C.c_int),
        "bxw_region_load_deltas": ([vp, vp, C.c_uint32, u32p], C.c_int),
        "bxw_region_apply_deltas": ([vp, C.c_uint64, C.c_int, C.c_uint32, u32p], C.c_int),
        "bxw_region_load_state": ([vp, vp, vp], C.c_int),
        "bxw_region_arena_extent": ([vp, u64p, u64p], C.c_int),
        "bxw_region_arena_stats": ([vp, C.POINTER(ArenaStats)], C.c_int),
        "bxw_region_compact_arena": ([vp, u64p, u64p], C.c_int),
        "bxw_region_set_origin": ([vp, C.c_int32, C.c_int32, C.c_int32], C.c_int),
        "bxw_region_page_table": ([vp, C.POINTER(vp), u32p], C.c_int),
        "bxw_region_download_pages": ([vp, vp], C.c_int),
        "bxw_region_halo_bytes": ([vp], C.c_size_t),
        "bxw_region_halo_export": ([vp, C.c_int, vp], C.c_int),
        "bxw_region_halo_import": ([vp, C.c_int, vp], C.c_int),
        "bxw_region_device_ptrs": ([vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)], C.c_int),
        "bxw_terragen_noise2": ([vp, C.c_uint64, vp, vp, C.c_size_t, vp], C.c_int),
        "bxw_stdgen_noise": ([vp, C.c_uint64, C.c_int, C.c_int, vp, vp, C.c_size_t, vp], C.c_int),
        "bxw_timer_start": ([vp], C.c_int),
        "bxw_timer_stop_ms": ([vp, C.POINTER(C.c_float)], C.c_int),
        "bxw_mesh_kernel_time_ms": ([vp, C.POINTER(C.c_float), u64p, C.c_int], C.c_int),
        "bxw_kernel_time_ms": ([vp, C.c_int, C.POINTER(C.c_float), u64p, C.c_int], C.c_int),
        "bxw_region_mesh_candidates": ([vp, u32p, u32p], C.c_int),
        "bxw_debug_phase_cycles": ([vp, u64p], C.c_int),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(L, name)  # AttributeError = the library does not export what the header declares
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily for tests


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _u32(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} u32 words, got {a.size}")
    return a


class Context:
    """One bxw_ctx: bound to one CUDA device; calls must be serialised by the caller."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.bxw_create(device, C.byref(h))
        if rc != 0:
            raise BxwError(rc, f"bxw_create(device={device}) failed (no usable sm_100 CUDA device?)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.bxw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise BxwError(rc, self.L.bxw_last_error(self.h).decode())

    # ---- context ----
    def set_stream(self, cuda_stream_ptr):
        """Launch on an externally owned cudaStream_t; None = back to the context's own stream. The legacy default
        stream (handle 0, e.g. torch.cuda.current_stream() before any torch.cuda.set_stream) cannot be expressed through
        the ABI's NULL-means-own convention: create a torch.cuda.Stream(), make it current and pass its handle, so that
        torch / NCCL work and the library's kernels are ordered on one stream."""
        if cuda_stream_ptr is not None and int(cuda_stream_ptr) == 0:
            raise ValueError("stream handle 0 (legacy default stream): pass a torch.cuda.Stream().cuda_stream, or None")
        self._ck(self.L.bxw_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    OPT_CHUNK_SUMMARIES = 0
    OPT_MESH_FROM_HEIGHT_MAP = 1
    OPT_ELIDE_UNIFORM = 2
    OPT_CUBES_ONLY_MESHER = 3
    OPT_PIPELINED_MESHER = 4

    def set_option(self, option, value):
        self._ck(self.L.bxw_set_option(self.h, option, int(value)))

    def synchronize(self):
        self._ck(self.L.bxw_synchronize(self.h))

    def kernel_launches(self):
        return int(self.L.bxw_kernel_launches(self.h))

    def set_registry(self, defs):
        """defs: sequence of dicts {present, mesh_kind, debug_color[3], tex[6]} indexed by block id."""
        arr = (BlockDef * len(defs))()
        for i, d in enumerate(defs):
            arr[i].present = int(d.get("present", 1))
            arr[i].mesh_kind = int(d["mesh_kind"])
            for k in range(3):
                arr[i].debug_color[k] = float(d["debug_color"][k])
            for k in range(6):
                arr[i].tex[k] = int(d["tex"][k])
        self._ck(self.L.bxw_set_registry(self.h, arr, len(defs)))

    def set_gen_ids(self, void, grass, dirt, stone, snow_grass):
        self._ck(self.L.bxw_set_gen_ids(self.h, void, grass, dirt, stone, snow_grass))

    # ---- per-chunk drop-ins ----
    def gen_chunk(self, seed, cx, cy, cz, out=None):
        out = np.empty(32768, dtype=np.uint32) if out is None else out
        self._ck(self.L.bxw_gen_chunk(self.h, seed, cx, cy, cz, _ptr(out)))
        return out

    def _fetch(self, span):
        v = np.empty(int(span["v_count"]), dtype=VERTEX_DTYPE)
        ix = np.empty(int(span["i_count"]), dtype=np.uint32)
        self._ck(self.L.bxw_mesh_fetch(self.h, _ptr(v), _ptr(ix)))
        return v, ix

    def mesh_chunk27(self, dense27):
        arrs = [_u32(a, 32768) for a in dense27]
        if len(arrs) != 27:
            raise ValueError("mesh_chunk27 needs 27 chunks")
        ptrs = (C.c_void_p * 27)(*[a.ctypes.data for a in arrs])
        span = np.zeros(1, dtype=SPAN_DTYPE)
        self._ck(self.L.bxw_mesh_chunk27(self.h, ptrs, _ptr(span)))
        return self._fetch(span[0])

    def mesh_chunk27_rle(self, rle27):
        arrs = [_u32(a) for a in rle27]
        if len(arrs) != 27:
            raise ValueError("mesh_chunk27_rle needs 27 chunks")
        ptrs = (C.c_void_p * 27)(*[a.ctypes.data for a in arrs])
        lens = (C.c_size_t * 27)(*[a.size for a in arrs])
        span = np.zeros(1, dtype=SPAN_DTYPE)
        self._ck(self.L.bxw_mesh_chunk27_rle(self.h, ptrs, lens, _ptr(span)))
        return self._fetch(span[0])

    def is_chunk_trivial(self, rle_words):
        w = _u32(rle_words)
        out = C.c_int()
        self._ck(self.L.bxw_is_chunk_trivial(self.h, _ptr(w), w.size, C.byref(out)))
        return bool(out.value)

    # ---- RLE codec ----
    def rle_compress(self, dense):
        """dense: (n, 32768) u32 -> (words, offsets[n+1])"""
        d = np.ascontiguousarray(dense, dtype=np.uint32).reshape(-1, 32768)
        n = d.shape[0]
        words = np.empty(self.L.bxw_rle_bound(n), dtype=np.uint32)
        offs = np.empty(n + 1, dtype=np.uint64)
        self._ck(self.L.bxw_rle_compress(self.h, _ptr(d), n, _ptr(words), _ptr(offs)))
        return words[: int(offs[n])].copy(), offs

    def rle_decompress(self, words, offsets):
        w = _u32(words)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = o.size - 1
        out = np.empty((n, 32768), dtype=np.uint32)
        self._ck(self.L.bxw_rle_decompress(self.h, _ptr(w), _ptr(o), n, _ptr(out)))
        return out

    # ---- region ----
    def region_create(self, cx0, cy0, cz0, nx, ny, nz):
        self._ck(self.L.bxw_region_create(self.h, cx0, cy0, cz0, nx, ny, nz))
        self.region_dims = (nx, ny, nz)
        self.region_origin = (cx0, cy0, cz0)

    def region_destroy(self):
        self._ck(self.L.bxw_region_destroy(self.h))

    def region_generate(self, seed, precision=64):
        self._ck(self.L.bxw_region_generate(self.h, seed, precision))

    def region_generate_part(self, seed, precision, part):
        self._ck(self.L.bxw_region_generate_part(self.h, seed, precision, part))

    def region_heightmap(self, raw=False):
        nx, ny, nz = self.region_dims
        W, D = (nx + 2) * 32, (nz + 2) * 32
        h = np.empty((D, W), dtype=np.int32)
        e = np.empty((D, W), dtype=np.uint8)
        r = np.empty((D, W), dtype=np.float64) if raw else None
        self._ck(self.L.bxw_region_heightmap(self.h, _ptr(h), _ptr(e), _ptr(r) if raw else None))
        return (h, e, r) if raw else (h, e)

    def region_upload_chunk(self, lx, ly, lz, blocks):
        b = _u32(blocks, 32768)
        self._ck(self.L.bxw_region_upload_chunk(self.h, lx, ly, lz, _ptr(b)))

    def region_download_chunk(self, lx, ly, lz):
        out = np.empty(32768, dtype=np.uint32)
        self._ck(self.L.bxw_region_download_chunk(self.h, lx, ly, lz, _ptr(out)))
        return out

    def region_export_rle(self, lpos):
        """lpos: (n, 3) local chunk coordinates -> (words, offsets[n+1]); the voxels never leave the device."""
        lp = np.ascontiguousarray(lpos, dtype=np.int32).reshape(-1, 3)
        n = lp.shape[0]
        offs = np.empty(n + 1, dtype=np.uint64)
        self._ck(self.L.bxw_region_export_rle(self.h, _ptr(lp), n, None, 0, _ptr(offs)))
        words = np.empty(max(int(offs[n]), 1), dtype=np.uint32)
        self._ck(self.L.bxw_region_export_rle(self.h, _ptr(lp), n, _ptr(words), words.size, _ptr(offs)))
        return words[: int(offs[n])], offs

    def region_import_rle(self, lpos, words, offsets):
        lp = np.ascontiguousarray(lpos, dtype=np.int32).reshape(-1, 3)
        w = _u32(words)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        if o.size != lp.shape[0] + 1:
            raise ValueError("offsets must have one more entry than lpos")
        self._ck(self.L.bxw_region_import_rle(self.h, _ptr(lp), lp.shape[0], _ptr(w), _ptr(o)))

    def region_fill_synthetic(self, seed=1, void_threshold=128):
        self._ck(self.L.bxw_region_fill_synthetic(self.h, seed, void_threshold))

    def region_mesh_reserve(self, max_vertices, max_indices):
        self._ck(self.L.bxw_region_mesh_reserve(self.h, max_vertices, max_indices))

    def region_mesh(self):
        v, i = C.c_uint64(), C.c_uint64()
        self._ck(self.L.bxw_region_mesh(self.h, C.byref(v), C.byref(i)))
        return v.value, i.value

    def region_generate_and_mesh(self, seed, precision=64):
        v, i = C.c_uint64(), C.c_uint64()
        self._ck(self.L.bxw_region_generate_and_mesh(self.h, seed, precision, C.byref(v), C.byref(i)))
        return v.value, i.value

    def region_mesh_sphere(self, anchor, radius):
        n, v, i = C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._ck(self.L.bxw_region_mesh_sphere(self.h, anchor[0], anchor[1], anchor[2], radius, C.byref(n), C.byref(v), C.byref(i)))
        return n.value, v.value, i.value

    def region_mesh_spans(self):
        nx, ny, nz = self.region_dims
        spans = np.empty(nx * ny * nz, dtype=SPAN_DTYPE)
        self._ck(self.L.bxw_region_mesh_spans(self.h, _ptr(spans)))
        return spans

    def region_mesh_fetch(self, n_vertices, n_indices, vertices=None, indices=None):
        v = np.empty(n_vertices, dtype=VERTEX_DTYPE) if vertices is None else vertices
        ix = np.empty(n_indices, dtype=np.uint32) if indices is None else indices
        self._ck(self.L.bxw_region_mesh_fetch(self.h, _ptr(v), n_vertices, _ptr(ix), n_indices))
        return v, ix

    def region_mesh_fetch_span(self, span):
        """One chunk's ChunkBuffers (voxrender.rs:34-37) from its span."""
        v = np.empty(int(span["v_count"]), dtype=VERTEX_DTYPE)
        ix = np.empty(int(span["i_count"]), dtype=np.uint32)
        self._ck(self.L.bxw_region_mesh_fetch_range(self.h, int(span["v_off"]), v.size, _ptr(v), int(span["i_off"]), ix.size, _ptr(ix)))
        return v, ix

    def region_mesh_fetch_ptr(self, vptr, n_vertices, iptr, n_indices):
        self._ck(self.L.bxw_region_mesh_fetch(self.h, C.c_void_p(vptr), n_vertices, C.c_void_p(iptr), n_indices))

    def region_mesh_fetch_async_ptr(self, vptr, n_vertices, iptr, n_indices):
        self._ck(self.L.bxw_region_mesh_fetch_async(self.h, C.c_void_p(vptr), n_vertices, C.c_void_p(iptr), n_indices))

    def region_fetch_wait(self):
        self._ck(self.L.bxw_region_fetch_wait(self.h))

    def region_apply_changes(self, changes):
        ch = np.ascontiguousarray(changes, dtype=CHANGE_DTYPE)
        nd = C.c_uint32()
        self._ck(self.L.bxw_region_apply_changes(self.h, _ptr(ch), ch.size, C.byref(nd)))
        return nd.value

    def region_remesh_dirty(self):
        n = C.c_uint32()
        self._ck(self.L.bxw_region_remesh_dirty(self.h, C.byref(n)))
        return n.value

    # ---- load anchors (World::recalculate_load_deltas / main_loop_tick on the device) ----
    def region_unload_all(self):
        self._ck(self.L.bxw_region_unload_all(self.h))

    def region_update_anchors(self, anchors):
        """anchors: sequence of (cx, cy, cz, radius, load_mesh) -> (n_unload, n_load)"""
        a = np.array([tuple(x) for x in anchors], dtype=ANCHOR_DTYPE) if not isinstance(anchors, np.ndarray) else np.ascontiguousarray(anchors, dtype=ANCHOR_DTYPE)
        nu, nl = C.c_uint32(), C.c_uint32()
        self._ck(self.L.bxw_region_update_anchors(self.h, _ptr(a) if a.size else None, a.size, C.byref(nu), C.byref(nl)))
        return nu.value, nl.value

    def region_load_deltas(self, cap=None):
        n = C.c_uint32()
        self._ck(self.L.bxw_region_load_deltas(self.h, None, 0, C.byref(n)))
        m = n.value if cap is None else min(cap, n.value)
        out = np.zeros(m, dtype=DELTA_DTYPE)
        if m:
            self._ck(self.L.bxw_region_load_deltas(self.h, _ptr(out), m, C.byref(n)))
        return out

    def region_apply_deltas(self, seed=0, precision=64, max_deltas=0):
        """-> dict(consumed, generated, meshed, unmeshed)"""
        st = (C.c_uint32 * 4)()
        self._ck(self.L.bxw_region_apply_deltas(self.h, seed, precision, max_deltas, st))
        return dict(consumed=st[0], generated=st[1], meshed=st[2], unmeshed=st[3])

    def region_load_state(self):
        nx, ny, nz = self.region_dims
        b = np.empty((nx + 2) * (ny + 2) * (nz + 2), dtype=np.uint8)
        m = np.empty(nx * ny * nz, dtype=np.uint8)
        self._ck(self.L.bxw_region_load_state(self.h, _ptr(b), _ptr(m)))
        return b, m

    def region_arena_extent(self):
        v, i = C.c_uint64(), C.c_uint64()
        self._ck(self.L.bxw_region_arena_extent(self.h, C.byref(v), C.byref(i)))
        return v.value, i.value

    def region_arena_stats(self):
        st = ArenaStats()
        self._ck(self.L.bxw_region_arena_stats(self.h, C.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in ArenaStats._fields_}

    def region_compact_arena(self):
        v, i = C.c_uint64(), C.c_uint64()
        self._ck(self.L.bxw_region_compact_arena(self.h, C.byref(v), C.byref(i)))
        return v.value, i.value

    def region_set_origin(self, cx0, cy0, cz0):
        self._ck(self.L.bxw_region_set_origin(self.h, cx0, cy0, cz0))
        self.region_origin = (cx0, cy0, cz0)

    def region_page_table(self):
        """(device pointer of the u32 page table, number of shared single-datum pages behind the storage)"""
        p, n = C.c_void_p(), C.c_uint32()
        self._ck(self.L.bxw_region_page_table(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def region_download_pages(self):
        nx, ny, nz = self.region_dims
        out = np.empty((nx + 2) * (ny + 2) * (nz + 2), dtype=np.uint32)
        self._ck(self.L.bxw_region_download_pages(self.h, _ptr(out)))
        return out

    def region_halo_bytes(self):
        return int(self.L.bxw_region_halo_bytes(self.h))

    def region_halo_export(self, face, dev_ptr):
        self._ck(self.L.bxw_region_halo_export(self.h, face, C.c_void_p(dev_ptr)))

    def region_halo_import(self, face, dev_ptr):
        self._ck(self.L.bxw_region_halo_import(self.h, face, C.c_void_p(dev_ptr)))

    def region_device_ptrs(self):
        b, v, i = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self.L.bxw_region_device_ptrs(self.h, C.byref(b), C.byref(v), C.byref(i)))
        return b.value, v.value, i.value

    def terragen_noise2(self, seed, xs, ys):
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        ys = np.ascontiguousarray(ys, dtype=np.float64)
        out = np.empty(xs.size, dtype=np.float64)
        self._ck(self.L.bxw_terragen_noise2(self.h, seed, _ptr(xs), _ptr(ys), xs.size, _ptr(out)))
        return out

    def stdgen_noise(self, seed, which, xs, ys, precision=64):
        """Raw samples of StdGenerator(seed)'s noise function `which` (0..4 height maps, 5 elevation, 6 moisture)."""
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        ys = np.ascontiguousarray(ys, dtype=np.float64)
        out = np.empty(xs.size, dtype=np.float64)
        self._ck(self.L.bxw_stdgen_noise(self.h, seed, which, precision, _ptr(xs), _ptr(ys), xs.size, _ptr(out)))
        return out

    # ---- timing ----
    def timer_start(self):
        self._ck(self.L.bxw_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_float()
        self._ck(self.L.bxw_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def mesh_kernel_time_ms(self, reset=False):
        ms, n = C.c_float(), C.c_uint64()
        self._ck(self.L.bxw_mesh_kernel_time_ms(self.h, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, n.value

    KERNEL_MESH, KERNEL_FILL, KERNEL_HEIGHTMAP, KERNEL_EMIT = 0, 1, 2, 3

    def kernel_time_ms(self, which, reset=False):
        ms, n = C.c_float(), C.c_uint64()
        self._ck(self.L.bxw_kernel_time_ms(self.h, which, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, n.value

    def debug_phase_cycles(self):
        out = (C.c_uint64 * 12)()
        self._ck(self.L.bxw_debug_phase_cycles(self.h, out))
        return list(out)

    def region_mesh_candidates(self):
        """(chunks meshed, interior chunks) of the last whole-region mesh pass"""
        a, b = C.c_uint32(), C.c_uint32()
        self._ck(self.L.bxw_region_mesh_candidates(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value
