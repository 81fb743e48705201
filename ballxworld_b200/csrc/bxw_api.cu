// bxw_api.cu — the extern "C" boundary of libbxw_cuda.so (include/bxw_cuda.h): context, registry, GPU-resident
// region storage, and the host-side orchestration of the kernels. No CPU fallback anywhere: every entry point that
// produces blocks or meshes launches the sm_100a kernels.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <algorithm>

#include "aux_kernels.cuh"
#include "sched_kernels.cuh"
#include "device_common.cuh"

using namespace bxw;

namespace {

struct Region {
    bool live = false;
    int32_t cx0 = 0, cy0 = 0, cz0 = 0; // world chunk coordinate of interior chunk (0,0,0)
    uint32_t nx = 0, ny = 0, nz = 0;   // interior dims
    int SX = 0, SY = 0, SZ = 0;        // storage dims (shell included)
    uint32_t *blocks = nullptr;
    uint32_t *xface = nullptr; // [chunk][2][32][32]: x=0 / x=31 planes, maintained by every writer of `blocks`
    int32_t *hmap = nullptr;
    double *raw = nullptr;
    bool have_hmap = false;
    CellPointDev *cells = nullptr;
    size_t cells_cap = 0;
    uint32_t *work = nullptr, *work_span = nullptr; // full interior work list
    uint32_t n_work = 0;
    bxw_mesh_span *spans = nullptr;
    bxw_vertex *varena = nullptr;
    uint32_t *iarena = nullptr;
    unsigned long long v_cap = 0, i_cap = 0;
    bxw_vertex *varena2 = nullptr;         // compaction target (same capacities), kept between compactions
    uint32_t *iarena2 = nullptr;
    uint4 *frec = nullptr;             // face records between mesh_kernel and emit_kernel (device_common.cuh)
    uint4 *fslot = nullptr;            // per 32 records: the arena offsets of their chunk (v_off, i_off as two u64)
    unsigned long long f_cap = 0;
    bool reserved = false;
    MeshCounters *counters = nullptr;
    MeshCounters *h_counters = nullptr; // pinned
    uint8_t *dirty = nullptr;           // [n_work] interior dirty flags
    uint32_t *dirty_work = nullptr, *dirty_span = nullptr, *dirty_count = nullptr;
    unsigned long long *summary = nullptr; // per storage chunk: kUniformBit | datum, solid-terrain bits, or 0 (device_common.cuh)
    uint16_t fill_ids[5] = {0, 0, 0, 0, 0}; // generator block ids of the last fill (what the solid-terrain bits refer to)
    uint8_t *cand_keep = nullptr;          // candidate-list scratch
    uint32_t *cand_blk = nullptr;
    // Uniform-chunk elision: page[c] = chunk slot that holds chunk c's voxels: c itself, or one of the kSharedPages slots
    // behind the storage (n_storage() + k) that hold a single datum. Readers go through the table, writers materialise first.
    uint32_t *page = nullptr;
    uint32_t shared_datum[2] = {0, 0};     // datum of shared page 0 (void) / 1 (stone) of the last fill
    bool shared_valid = false;
    uint8_t *need = nullptr;               // [n_storage] scratch flags (halo import)
    uint2 *span_cap = nullptr;             // [n_work] arena reservation behind each span (device_common.cuh)
    bxw_mesh_span *spans_bak = nullptr;    // re-mesh passes: spans / reservations before the pass (restored if the arena must grow)
    uint2 *span_cap_bak = nullptr;
    // cubes-only mesher: chunks it hands over to the general kernel; shape_hint: 0 = storage holds generator output and edits
    // with cubes only, 1 = an edit placed a slope / corner (re-mesh passes over the dirty chunks then go straight to the general
    // kernel instead of failing over chunk by chunk), 2 = bulk content of unknown shapes (synthetic fill)
    uint32_t *fb_work = nullptr, *fb_span = nullptr, *fb_count = nullptr;
    int shape_hint = 0;
    bool streaming = false;                // load-anchor mode (bxw_region_unload_all / _update_anchors): chunks carry load states
    // load-anchor scheduler (sched_kernels.cu)
    uint8_t *blk_loaded = nullptr, *mesh_loaded = nullptr, *sched_code = nullptr;
    uint32_t *sched_key = nullptr, *sched_hist = nullptr, *sched_totals = nullptr, *sched_order = nullptr;
    uint32_t *sched_gen = nullptr, *sched_mesh = nullptr, *sched_mesh_span = nullptr, *sched_unmesh = nullptr, *sched_counts = nullptr;
    bxw_chunk_delta *sched_deltas = nullptr;
    bxw_load_anchor *sched_anchors = nullptr;
    uint32_t sched_anchor_cap = 0, sched_bins = 0, sched_n_unload = 0, sched_n_load = 0;
    uint64_t gen_seed = 0;                 // seed / precision of the last generate (bxw_region_heightmap's raw output)
    int gen_precision = 64;
    bool part1_pending = false;            // bxw_region_generate_part(1) done, part 2 outstanding
    GenParams part1_gp;
    size_t n_storage() const { return (size_t)SX * SY * SZ; }
};
constexpr uint32_t kSharedPages = 2;

} // namespace

struct bxw_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    // registry
    uint32_t reg_n = 0;
    std::vector<bxw_block_def> reg_host;
    uint8_t *d_kind = nullptr;
    float *d_texf = nullptr, *d_color = nullptr;
    bool have_gen_ids = false;
    uint16_t gen_ids[5] = {0, 0, 0, 0, 0};
    // tables
    MeshTables *d_tables = nullptr;
    GenTables *d_gen = nullptr;
    bool gen_seed_valid = false;
    uint64_t gen_seed = 0;
    // regions
    Region region;  // the user's region
    Region scratch; // 3x3x3 storage behind the per-chunk drop-ins
    // timing
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_m0 = nullptr, ev_m1 = nullptr, ev_mm = nullptr; // ev_mm: between mesh_kernel and emit_kernel
    cudaEvent_t ev_g0 = nullptr, ev_g1 = nullptr, ev_g2 = nullptr; // before heightmap / before fill / after fill
    cudaStream_t copy_stream = nullptr;                            // bxw_region_mesh_fetch_async
    cudaEvent_t ev_mesh_done = nullptr, ev_copy_done = nullptr;
    bool copy_pending = false;                                     // ev_copy_done recorded and not yet waited for
    bool gen_pending = false, gen_pending_fused = false;          // ev_g* recorded, not yet read
    double mesh_ms = 0.0, fill_ms = 0.0, hmap_ms = 0.0, emit_ms = 0.0;
    uint64_t mesh_launches = 0, fill_launches = 0, hmap_launches = 0;
    bool use_summaries = true;                                     // BXW_OPT_CHUNK_SUMMARIES
    bool elide_uniform = true;                                     // BXW_OPT_ELIDE_UNIFORM
    bool lean_mesher = true;                                       // BXW_OPT_CUBES_ONLY_MESHER
    int pipeline_mode = 0;                                         // BXW_OPT_PIPELINED_MESHER: 0 off (default), 1 auto, 2 always
    uint32_t mesh_epoch = 0;                                       // tag of the last pipelined pass (24 bits)
    cudaStream_t emit_stream = nullptr;                            // side stream: emit_kernel of a pipelined pass; part 2 of a split generation
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_gen_fork = nullptr, ev_gen_join = nullptr;
    double last_faces_per_item = 0.0;                              // statistics of the last pass (the auto mode's heuristic)
    bool mesh_from_hmap = false;                                   // BXW_OPT_MESH_FROM_HEIGHT_MAP
    uint32_t last_candidates = 0, last_candidates_of = 0;          // last whole-region mesh: chunks listed / interior chunks
    // staging for host<->device codecs
    void *io_a = nullptr, *io_b = nullptr, *io_c = nullptr;
    size_t io_a_cap = 0, io_b_cap = 0, io_c_cap = 0;
    TerragenTables *d_tg = nullptr;
    bool tg_valid = false;
    uint64_t tg_seed = 0;
};

namespace {

int ensure_io(bxw_ctx *ctx, size_t bytes_a, size_t bytes_b);
int ensure_io_c(bxw_ctx *ctx, size_t bytes);

int fail(bxw_ctx *c, int code, const char *fmt, ...) {
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

#define CK(call)                                                                                            \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) return fail(ctx, BXW_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

DeviceRegistry dev_registry(const bxw_ctx *c) {
    DeviceRegistry r;
    r.kind = c->d_kind;
    r.texf = c->d_texf;
    r.color = c->d_color;
    r.n = c->reg_n;
    return r;
}

// ---- host restatement of the seeding arithmetic the generator needs (third-party crates, see oracle/noise.c) ----
uint32_t splitmix_next_u32(uint64_t &x) { // rand_xoshiro 0.6.0 SplitMix64::next_u32
    x += 0x9e3779b97f4a7c15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 33)) * 0x62A9D9ED799705F5ULL;
    z = (z ^ (z >> 28)) * 0xCB24D0A5C88C35B3ULL;
    return (uint32_t)(z >> 32);
}

void permutation_table(uint32_t seed, uint8_t out[256]) { // noise 0.8.2 PermutationTable::new
    uint32_t s[4] = {1u, seed, seed, seed}; // XorShiftRng::from_seed([1,0,0,0, seed x3] LE)
    auto next = [&]() {
        uint32_t t = s[0] ^ (s[0] << 11);
        s[0] = s[1];
        s[1] = s[2];
        s[2] = s[3];
        s[3] = s[3] ^ (s[3] >> 19) ^ (t ^ (t >> 8));
        return s[3];
    };
    for (int i = 0; i < 256; i++) out[i] = (uint8_t)i;
    for (uint32_t i = 255; i >= 1; i--) { // rand 0.7.3 SliceRandom::shuffle + UniformInt::sample_single
        const uint32_t range = i + 1;
        const uint32_t zone = (range << __builtin_clz(range)) - 1u;
        uint32_t hi;
        for (;;) {
            const uint64_t m = (uint64_t)next() * range;
            hi = (uint32_t)(m >> 32);
            if ((uint32_t)m <= zone) break;
        }
        const uint8_t tmp = out[i];
        out[i] = out[hi];
        out[hi] = tmp;
    }
}

int ensure_gen_tables(bxw_ctx *ctx, uint64_t seed) {
    if (ctx->gen_seed_valid && ctx->gen_seed == seed) return BXW_OK;
    GenTables t;
    uint64_t sd = seed; // SplitMix64::seed_from_u64 (stdgen.rs:106)
    for (int i = 0; i < 7; i++) permutation_table(splitmix_next_u32(sd), t.perm[i]);
    fill_noise_constants(t);
    if (!ctx->d_gen) CK(cudaMalloc(&ctx->d_gen, sizeof(GenTables)));
    CK(cudaMemcpyAsync(ctx->d_gen, &t, sizeof t, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // t is a stack object
    ctx->gen_seed = seed;
    ctx->gen_seed_valid = true;
    return BXW_OK;
}

void region_free(Region &r) {
    cudaFree(r.blocks);
    cudaFree(r.xface);
    cudaFree(r.hmap);
    cudaFree(r.raw);
    cudaFree(r.cells);
    cudaFree(r.work);
    cudaFree(r.work_span);
    cudaFree(r.spans);
    cudaFree(r.varena);
    cudaFree(r.iarena);
    cudaFree(r.varena2);
    cudaFree(r.iarena2);
    cudaFree(r.frec);
    cudaFree(r.fslot);
    cudaFree(r.counters);
    cudaFree(r.dirty);
    cudaFree(r.dirty_work);
    cudaFree(r.dirty_span);
    cudaFree(r.dirty_count);
    cudaFree(r.summary);
    cudaFree(r.cand_keep);
    cudaFree(r.cand_blk);
    cudaFree(r.blk_loaded);
    cudaFree(r.mesh_loaded);
    cudaFree(r.sched_code);
    cudaFree(r.sched_key);
    cudaFree(r.sched_hist);
    cudaFree(r.sched_totals);
    cudaFree(r.sched_order);
    cudaFree(r.sched_gen);
    cudaFree(r.sched_mesh);
    cudaFree(r.sched_mesh_span);
    cudaFree(r.sched_unmesh);
    cudaFree(r.sched_counts);
    cudaFree(r.sched_deltas);
    cudaFree(r.sched_anchors);
    cudaFree(r.page);
    cudaFree(r.fb_work);
    cudaFree(r.fb_span);
    cudaFree(r.fb_count);
    cudaFree(r.need);
    cudaFree(r.span_cap);
    cudaFree(r.spans_bak);
    cudaFree(r.span_cap_bak);
    if (r.h_counters) cudaFreeHost(r.h_counters);
    r = Region();
}

int region_alloc(bxw_ctx *ctx, Region &r, int32_t cx0, int32_t cy0, int32_t cz0, uint32_t nx, uint32_t ny, uint32_t nz) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream)); // an asynchronous fetch may still read the old arena
    ctx->copy_pending = false;
    region_free(r);
    r.cx0 = cx0;
    r.cy0 = cy0;
    r.cz0 = cz0;
    r.nx = nx;
    r.ny = ny;
    r.nz = nz;
    r.SX = (int)nx + 2;
    r.SY = (int)ny + 2;
    r.SZ = (int)nz + 2;
    const size_t nst = r.n_storage();
    if (nst >= (1ull << 32)) return fail(ctx, BXW_E_INVALID, "region too large");
    r.n_work = nx * ny * nz;
    CK(cudaMalloc(&r.blocks, (nst + kSharedPages) * kChunkVoxels * sizeof(uint32_t)));
    CK(cudaMalloc(&r.xface, (nst + kSharedPages) * 2048 * sizeof(uint32_t)));
    CK(cudaMalloc(&r.page, nst * sizeof(uint32_t)));
    launch_page_identity(r.page, (uint32_t)nst, ctx->stream);
    CK(cudaMalloc(&r.blk_loaded, nst));
    CK(cudaMalloc(&r.mesh_loaded, r.n_work));
    CK(cudaMemsetAsync(r.blk_loaded, 0, nst, ctx->stream));
    CK(cudaMemsetAsync(r.mesh_loaded, 0, r.n_work, ctx->stream));
    CK(cudaMalloc(&r.fb_work, r.n_work * sizeof(uint32_t)));
    CK(cudaMalloc(&r.fb_span, r.n_work * sizeof(uint32_t)));
    CK(cudaMalloc(&r.fb_count, sizeof(uint32_t)));
    CK(cudaMalloc(&r.need, nst));
    CK(cudaMemsetAsync(r.need, 0, nst, ctx->stream));
    CK(cudaMalloc(&r.span_cap, r.n_work * sizeof(uint2)));
    CK(cudaMemsetAsync(r.span_cap, 0, r.n_work * sizeof(uint2), ctx->stream));
    CK(cudaMalloc(&r.spans_bak, r.n_work * sizeof(bxw_mesh_span)));
    CK(cudaMalloc(&r.span_cap_bak, r.n_work * sizeof(uint2)));
    CK(cudaMalloc(&r.hmap, (size_t)r.SX * 32 * r.SZ * 32 * sizeof(int32_t)));
    CK(cudaMalloc(&r.work, r.n_work * sizeof(uint32_t)));
    CK(cudaMalloc(&r.work_span, r.n_work * sizeof(uint32_t)));
    CK(cudaMalloc(&r.spans, r.n_work * sizeof(bxw_mesh_span)));
    CK(cudaMalloc(&r.counters, sizeof(MeshCounters)));
    CK(cudaMallocHost(&r.h_counters, 2 * sizeof(MeshCounters))); // [1] receives the candidate-list length
    CK(cudaMalloc(&r.dirty, r.n_work));
    CK(cudaMalloc(&r.dirty_work, r.n_work * sizeof(uint32_t)));
    CK(cudaMalloc(&r.dirty_span, r.n_work * sizeof(uint32_t)));
    CK(cudaMalloc(&r.dirty_count, sizeof(uint32_t)));
    CK(cudaMalloc(&r.summary, nst * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(r.summary, 0, nst * sizeof(unsigned long long), ctx->stream));
    CK(cudaMalloc(&r.cand_keep, r.n_work));
    CK(cudaMalloc(&r.cand_blk, ((r.n_work + 255) / 256 + 1) * sizeof(uint32_t)));
    CK(cudaMemsetAsync(r.spans, 0, r.n_work * sizeof(bxw_mesh_span), ctx->stream));
    CK(cudaMemsetAsync(r.dirty, 0, r.n_work, ctx->stream));
    CK(cudaMemsetAsync(r.counters, 0, sizeof(MeshCounters), ctx->stream));
    // interior work list: span slot = ix + nx*(iz + nz*iy); processing order x, then z, then y
    std::vector<uint32_t> w(r.n_work), ws(r.n_work);
    size_t k = 0;
    for (uint32_t iy = 0; iy < ny; iy++)
        for (uint32_t iz = 0; iz < nz; iz++)
            for (uint32_t ix = 0; ix < nx; ix++) {
                w[k] = (uint32_t)((ix + 1) + r.SX * ((iz + 1) + r.SZ * (iy + 1)));
                ws[k] = (uint32_t)(ix + nx * (iz + nz * iy));
                k++;
            }
    CK(cudaMemcpyAsync(r.work, w.data(), r.n_work * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(r.work_span, ws.data(), r.n_work * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    r.live = true;
    return BXW_OK;
}

// fold the event pair of the previous generation into the per-kernel totals (blocks until that generation finished)
int flush_gen_timing(bxw_ctx *ctx) {
    if (!ctx->gen_pending) return BXW_OK;
    CK(cudaEventSynchronize(ctx->ev_g2));
    float a = 0.f, b = 0.f;
    CK(cudaEventElapsedTime(&a, ctx->ev_g0, ctx->ev_g1));
    CK(cudaEventElapsedTime(&b, ctx->ev_g1, ctx->ev_g2));
    ctx->fill_ms += b;
    ctx->fill_launches += 1;
    if (!ctx->gen_pending_fused) { // otherwise the column parameters were computed inside the fill kernel
        ctx->hmap_ms += a;
        ctx->hmap_launches += 1;
    }
    ctx->gen_pending = false;
    return BXW_OK;
}

// cell-point table + kernel parameters of the region's column grid
int gen_setup(bxw_ctx *ctx, Region &r, uint64_t seed, int precision, GenParams &gp) {
    if (precision != 32 && precision != 64) return fail(ctx, BXW_E_INVALID, "precision must be 32 or 64");
    int rc = ensure_gen_tables(ctx, seed);
    if (rc) return rc;
    const int32_t W = r.SX * 32, D = r.SZ * 32;
    const int32_t x0 = (r.cx0 - 1) * 32, z0 = (r.cz0 - 1) * 32;
    const int32_t cell_x0 = x0 / 128 - 1, cell_z0 = z0 / 128 - 1; // truncating, as stdgen.rs:150
    const int32_t cell_x1 = (x0 + W - 1) / 128 + 1, cell_z1 = (z0 + D - 1) / 128 + 1;
    const int32_t cw = cell_x1 - cell_x0 + 1, cd = cell_z1 - cell_z0 + 1;
    const size_t need = (size_t)cw * cd * 4;
    if (need > r.cells_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(r.cells);
        r.cells = nullptr;
        r.cells_cap = 0;
        CK(cudaMalloc(&r.cells, need * sizeof(CellPointDev)));
        r.cells_cap = need;
    }
    launch_cellpoints_kernel(seed, cell_x0, cell_z0, cw, cd, ctx->d_gen, r.cells, precision, ctx->stream);
    ctx->launches += 1;
    gp.seed = seed;
    gp.x0 = x0;
    gp.z0 = z0;
    gp.W = W;
    gp.D = D;
    gp.cell_x0 = cell_x0;
    gp.cell_z0 = cell_z0;
    gp.cells_w = cw;
    gp.cells_d = cd;
    gp.cells = r.cells;
    gp.tables = ctx->d_gen;
    gp.hmap = r.hmap;
    gp.raw = nullptr;
    return BXW_OK;
}

void fill_params(bxw_ctx *ctx, Region &r, bool elide, FillParams &fp) {
    fp.hmap = r.hmap;
    fp.W = r.SX * 32;
    fp.blocks = r.blocks;
    fp.xface = r.xface;
    fp.summary = r.summary;
    fp.SX = r.SX;
    fp.SY = r.SY;
    fp.SZ = r.SZ;
    fp.cy_min = r.cy0 - 1;
    fp.id_void = (uint32_t)ctx->gen_ids[0] << 16;
    fp.id_grass = (uint32_t)ctx->gen_ids[1] << 16;
    fp.id_dirt = (uint32_t)ctx->gen_ids[2] << 16;
    fp.id_stone = (uint32_t)ctx->gen_ids[3] << 16;
    fp.id_snow = (uint32_t)ctx->gen_ids[4] << 16;
    fp.page = r.page;
    fp.elide = elide ? 1 : 0;
    fp.page_void = (uint32_t)r.n_storage();
    fp.page_stone = (uint32_t)r.n_storage() + 1u;
    fp.list = nullptr;
    fp.n_list = 0;
    fp.col_mode = 0;
}

// the two shared single-datum pages behind the storage hold the void / stone datum of the current generator ids
int ensure_shared_pages(bxw_ctx *ctx, Region &r) {
    const uint32_t dv = (uint32_t)ctx->gen_ids[0] << 16, ds = (uint32_t)ctx->gen_ids[3] << 16;
    if (r.shared_valid && r.shared_datum[0] == dv && r.shared_datum[1] == ds) return BXW_OK;
    if (r.shared_valid) {
        // chunks of an earlier fill may still point at the shared pages: give them their own storage before the pages change
        std::vector<uint32_t> pg(r.n_storage());
        CK(cudaMemcpyAsync(pg.data(), r.page, pg.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::vector<uint32_t> list;
        for (size_t c = 0; c < pg.size(); c++)
            if (pg[c] != (uint32_t)c) list.push_back((uint32_t)c);
        if (!list.empty()) {
            int rc = ensure_io(ctx, 16, list.size() * sizeof(uint32_t));
            if (rc) return rc;
            CK(cudaMemcpyAsync(ctx->io_b, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
            launch_materialise(r.blocks, r.xface, r.page, (const uint32_t *)ctx->io_b, (uint32_t)list.size(), ctx->stream);
            ctx->launches += 1;
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    launch_uniform_page(r.blocks, r.xface, (uint32_t)r.n_storage(), dv, ctx->stream);
    launch_uniform_page(r.blocks, r.xface, (uint32_t)r.n_storage() + 1u, ds, ctx->stream);
    ctx->launches += 2;
    r.shared_datum[0] = dv;
    r.shared_datum[1] = ds;
    r.shared_valid = true;
    return BXW_OK;
}

// part: 0 = the whole box; 1 = only the two interior boundary chunk columns along x (what a slab-boundary halo export reads),
// 2 = all the other columns (after a part-1 call with the same arguments): a multi-GPU step exports and sends the boundary
// planes while part 2 is still running.
int region_generate(bxw_ctx *ctx, Region &r, uint64_t seed, int precision, bool elide, int part = 0) {
    CK(cudaSetDevice(ctx->device));
    if (part != 2) {
        int rc0 = flush_gen_timing(ctx);
        if (rc0) return rc0;
    }
    if (!ctx->have_gen_ids) return fail(ctx, BXW_E_NO_REGISTRY, "bxw_set_gen_ids not called");
    GenParams gp;
    int rc = BXW_OK;
    if (part == 2) {
        if (!r.part1_pending || r.gen_seed != seed || r.gen_precision != precision)
            return fail(ctx, BXW_E_INVALID, "generate part 2 needs a preceding part 1 with the same seed and precision");
        gp = r.part1_gp;
    } else {
        rc = gen_setup(ctx, r, seed, precision, gp);
        if (rc) return rc;
    }
    if (elide && part != 2) {
        rc = ensure_shared_pages(ctx, r);
        if (rc) return rc;
    }
    // the fill kernel computes the column parameters of its 32x32 columns itself (FP64 phase of some CTAs overlaps the
    // store phase of others) and writes them to the height map
    if (part != 2) {
        CK(cudaEventRecord(ctx->ev_g0, ctx->stream));
        CK(cudaEventRecord(ctx->ev_g1, ctx->stream));
    }
    FillParams fp;
    fill_params(ctx, r, elide, fp);
    fp.col_mode = part;
    for (int i = 0; i < 5; i++) r.fill_ids[i] = ctx->gen_ids[i];
    // Part 2 runs on the side stream, ordered after everything that preceded part 1 (tables, cell points, shared pages, the
    // previous pass over this storage) but NOT after part 1 itself and what the caller queued behind it -- the plane export and
    // the transfer to the neighbour ranks -- so it fills the SMs the short part-1 launch leaves idle and hides the exchange.
    // The caller's stream continues when part 2 is done.
    cudaStream_t gs = ctx->stream;
    if (part == 1) CK(cudaEventRecord(ctx->ev_gen_fork, ctx->stream));
    if (part == 2) {
        gs = ctx->emit_stream;
        CK(cudaStreamWaitEvent(gs, ctx->ev_gen_fork, 0));
    }
    launch_gen_fill_kernel(fp, gp, precision, gs);
    if (part == 0) CK(cudaEventRecord(ctx->ev_g2, ctx->stream)); // the fill timer of an unsplit generation: the fill kernel alone
    launch_fill_xface_kernel(fp, gs);
    ctx->launches += 2;
    CK(cudaGetLastError());
    if (part != 1 && r.blk_loaded) CK(cudaMemsetAsync(r.blk_loaded, 1, r.n_storage(), gs)); // every chunk now holds block data
    if (part == 2) {
        CK(cudaEventRecord(ctx->ev_gen_join, gs));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_gen_join, 0));
    }
    if (part == 2) CK(cudaEventRecord(ctx->ev_g2, ctx->stream)); // split: part 1 .. the join (whatever ran between them included)
    if (part != 1) {
        ctx->gen_pending = true;
        ctx->gen_pending_fused = true;
    }
    r.have_hmap = part != 1;
    r.shape_hint = 0;
    r.gen_seed = seed;
    r.gen_precision = precision;
    r.part1_pending = part == 1;
    r.part1_gp = gp;
    return BXW_OK;
}

int arena_resize(bxw_ctx *ctx, Region &r, unsigned long long v_cap, unsigned long long i_cap, unsigned long long keep_v,
                 unsigned long long keep_i) {
    bxw_vertex *nv = nullptr;
    uint32_t *ni = nullptr;
    if (ctx->copy_pending) { // an asynchronous fetch may still be reading the arena that is about to be freed
        CK(cudaEventSynchronize(ctx->ev_copy_done));
        ctx->copy_pending = false;
    }
    if (v_cap < 64) v_cap = 64;
    if (i_cap < 64) i_cap = 64;
    CK(cudaMalloc(&nv, v_cap * sizeof(bxw_vertex)));
    CK(cudaMalloc(&ni, i_cap * sizeof(uint32_t)));
    if (keep_v && r.varena) CK(cudaMemcpyAsync(nv, r.varena, keep_v * sizeof(bxw_vertex), cudaMemcpyDeviceToDevice, ctx->stream));
    if (keep_i && r.iarena) CK(cudaMemcpyAsync(ni, r.iarena, keep_i * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(r.varena);
    cudaFree(r.iarena);
    cudaFree(r.varena2); // (sized like the old arena)
    cudaFree(r.iarena2);
    r.varena2 = nullptr;
    r.iarena2 = nullptr;
    r.varena = nv;
    r.iarena = ni;
    r.v_cap = v_cap;
    r.i_cap = i_cap;
    return BXW_OK;
}

// the face records are transient (one mesh pass), so growing never copies
int faces_resize(bxw_ctx *ctx, Region &r, unsigned long long f_cap) {
    if (f_cap <= r.f_cap) return BXW_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(r.frec);
    cudaFree(r.fslot);
    r.frec = nullptr;
    r.fslot = nullptr;
    r.f_cap = 0;
    f_cap = (f_cap + 31ull) & ~31ull;
    CK(cudaMalloc(&r.frec, f_cap * sizeof(uint4)));
    CK(cudaMalloc(&r.fslot, (f_cap / 32) * sizeof(uint4)));
    CK(cudaMemsetAsync(r.fslot, 0, (f_cap / 32) * sizeof(uint4), ctx->stream)); // no entry may carry a pass tag by accident
    r.f_cap = f_cap;
    return BXW_OK;
}

// Runs the mesher over a work list. append=false restarts arena allocation at 0; append=true is a re-mesh pass: a chunk whose
// new mesh fits its old arena reservation is rewritten in place, the others move to the end of the arena.
int region_mesh(bxw_ctx *ctx, Region &r, const uint32_t *work, const uint32_t *work_span, uint32_t n_work, bool append,
                bool from_hmap = false) {
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_kind) return fail(ctx, BXW_E_NO_REGISTRY, "bxw_set_registry not called");
    MeshParams p;
    p.blocks = r.blocks;
    p.xface = r.xface;
    p.page = r.page;
    p.SX = r.SX;
    p.SY = r.SY;
    p.SZ = r.SZ;
    p.work = work;
    p.work_span = work_span;
    p.n_work = n_work;
    p.n_work_dev = nullptr;
    p.spans = r.spans;
    p.span_cap = r.span_cap;
    p.reuse = append ? 1 : 0;
    p.counters = r.counters;
    p.cursor = &r.counters->work_next;
    p.tables = ctx->d_tables;
    p.reg = dev_registry(ctx);
    p.hmap = nullptr;
    p.hmap_W = 0;
    p.cy_min = 0;
    p.gen_all_cubes = 0;
    for (int i = 0; i < 5; i++) p.gen_datum[i] = 0;
    auto is_cube = [&](uint16_t id) { return id < ctx->reg_n && ctx->reg_host[id].present && ctx->reg_host[id].mesh_kind != 0; };
    // whole-region passes over a device-built candidate list: the list kernels run once per attempt (they also clear the
    // spans of the chunks they drop)
    const bool cand_hmap = from_hmap && !work;
    const bool cand_storage = !from_hmap && !work && ctx->use_summaries;
    if (from_hmap) { // fused generate+mesh: the region holds pristine StdGenerator terrain and its column parameters
        p.hmap = r.hmap;
        p.hmap_W = r.SX * 32;
        p.cy_min = r.cy0 - 1;
        for (int i = 0; i < 5; i++) p.gen_datum[i] = (uint32_t)ctx->gen_ids[i] << 16;
        p.gen_all_cubes = is_cube(ctx->gen_ids[1]) && is_cube(ctx->gen_ids[2]) && is_cube(ctx->gen_ids[3]) && is_cube(ctx->gen_ids[4]);
    }
    if (cand_hmap) {
        // whole interior: only chunks the surface passes through can have a mesh; list them on the device
        CandidateParams c;
        c.hmap = p.hmap;
        c.hmap_W = p.hmap_W;
        c.cy_min = p.cy_min;
        c.SX = r.SX;
        c.SY = r.SY;
        c.SZ = r.SZ;
        c.void_is_empty = !is_cube(ctx->gen_ids[0]);
        c.gen_all_cubes = p.gen_all_cubes;
        c.work = r.dirty_work;
        c.work_span = r.dirty_span;
        c.count = r.dirty_count;
        c.spans = r.spans;
        c.span_cap = r.span_cap;
        CK(cudaMemsetAsync(r.dirty_count, 0, sizeof(uint32_t), ctx->stream));
        launch_hmap_candidates(c, ctx->stream);
        ctx->launches += 1;
        p.work = r.dirty_work;
        p.work_span = r.dirty_span;
        p.n_work_dev = r.dirty_count;
    }
    if (cand_storage) {
        // whole interior from storage: drop the chunks whose summaries prove an empty mesh, keep the tiled order
        StorageCandParams c;
        c.summary = r.summary;
        // solid terrain = grass, dirt, stone, snow grass of the last fill: do they all mesh as cubes today?
        c.solids_are_cubes = is_cube(r.fill_ids[1]) && is_cube(r.fill_ids[2]) && is_cube(r.fill_ids[3]) && is_cube(r.fill_ids[4]);
        c.SX = r.SX;
        c.SY = r.SY;
        c.SZ = r.SZ;
        c.reg = p.reg;
        c.keep = r.cand_keep;
        c.blk = r.cand_blk;
        c.work = r.dirty_work;
        c.work_span = r.dirty_span;
        c.count = r.dirty_count;
        c.spans = r.spans;
        c.span_cap = r.span_cap;
        launch_storage_candidates(c, ctx->stream);
        ctx->launches += 3;
        p.work = r.dirty_work;
        p.work_span = r.dirty_span;
        p.n_work_dev = r.dirty_count;
    }
    const bool lean = ctx->lean_mesher && !from_hmap && r.shape_hint < 2 && !(append && r.shape_hint == 1);
    p.fallback_work = r.fb_work;
    p.fallback_span = r.fb_span;
    p.fallback_count = r.fb_count;
    MeshCounters base; // allocator state the pass starts from
    std::memset(&base, 0, sizeof base);
    if (append) {
        CK(cudaMemcpyAsync(r.h_counters, r.counters, sizeof(MeshCounters), cudaMemcpyDeviceToHost, ctx->stream));
        // a re-mesh pass edits spans and reservations in place: keep what they were in case the arena has to grow
        CK(cudaMemcpyAsync(r.spans_bak, r.spans, r.n_work * sizeof(bxw_mesh_span), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(r.span_cap_bak, r.span_cap, r.n_work * sizeof(uint2), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        base.v_alloc = r.h_counters->v_alloc;
        base.i_alloc = r.h_counters->i_alloc;
        base.v_garbage = r.h_counters->v_garbage;
        base.i_garbage = r.h_counters->i_garbage;
    }
    for (int attempt = 0; attempt < 3; attempt++) {
        *r.h_counters = base;
        CK(cudaMemcpyAsync(r.counters, r.h_counters, sizeof(MeshCounters), cudaMemcpyHostToDevice, ctx->stream));
        const bool count_only = (r.varena == nullptr);
        p.varena = r.varena;
        p.iarena = r.iarena;
        p.v_cap = r.v_cap;
        p.i_cap = r.i_cap;
        p.frec = r.frec;
        p.fslot = r.fslot;
        p.f_cap = r.f_cap;
        p.count_only = count_only ? 1 : 0;
        // Pipelined pass: emit_kernel runs NEXT TO mesh_kernel on a second stream and consumes each chunk's record block as
        // soon as it is published (mesh_kernel is compute / latency bound, emit_kernel memory bound). Both leave room for each
        // other on every SM: mesh_kernel takes half its usual CTAs, emit_kernel two. Worth it when the pass emits a lot per
        // chunk; a pass over mostly empty chunks keeps the full-occupancy mesh_kernel followed by emit_kernel.
        const bool pipelined = !count_only && !getenv("BXW_EXP_NO_PIPELINE") &&
                               (ctx->pipeline_mode == 2 || (ctx->pipeline_mode == 1 && ctx->last_faces_per_item >= 400.0));
        p.pipelined = pipelined ? 1 : 0;
        int mesh_cap = pipelined ? (lean ? 2 : 1) : 0, emit_ctas = 2;
        if (pipelined && getenv("BXW_EXP_PIPE_MESH")) mesh_cap = atoi(getenv("BXW_EXP_PIPE_MESH")); // experiments only
        if (pipelined && getenv("BXW_EXP_PIPE_EMIT")) emit_ctas = atoi(getenv("BXW_EXP_PIPE_EMIT"));
        if (pipelined) {
            ctx->mesh_epoch = (ctx->mesh_epoch % 0xFFFFFFu) + 1u;
            if (ctx->mesh_epoch == 1u && r.fslot) CK(cudaMemsetAsync(r.fslot, 0, (r.f_cap / 32) * sizeof(uint4), ctx->stream)); // tags wrapped
            p.epoch = ctx->mesh_epoch;
            MeshParams pf0 = p; // grids first: emit_kernel has to know how many mesh CTAs will retire
            pf0.n_work = n_work;
            p.mesh_ctas = mesh_kernel_grid(p, ctx->sm_count, lean, mesh_cap) + (lean ? mesh_kernel_grid(pf0, ctx->sm_count, false, 1) : 0u);
        } else {
            p.epoch = 0;
            p.mesh_ctas = 0;
        }
        CK(cudaEventRecord(ctx->ev_m0, ctx->stream));
        if (pipelined) {
            // emit_kernel starts behind everything queued so far (counters upload, candidate list) and behind a pending fetch
            CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->emit_stream, ctx->ev_fork, 0));
            if (ctx->copy_pending) CK(cudaStreamWaitEvent(ctx->emit_stream, ctx->ev_copy_done, 0));
        }
        if (lean) {
            // cubes-and-void chunks (all of generated terrain) go through the lean instantiation at twice the occupancy; what it
            // cannot handle (a slope / corner placed by an edit) lands on the fallback list for the general kernel
            CK(cudaMemsetAsync(r.fb_count, 0, sizeof(uint32_t), ctx->stream));
            launch_mesh_kernel(p, ctx->sm_count, ctx->stream, true, !work && !p.n_work_dev, mesh_cap);
            MeshParams pf = p;
            pf.work = r.fb_work;
            pf.work_span = r.fb_span;
            pf.n_work_dev = r.fb_count;
            pf.cursor = &r.counters->work_next2;
            launch_mesh_kernel(pf, ctx->sm_count, ctx->stream, false, false, pipelined ? 1 : 0);
            ctx->launches += 2;
        } else {
            launch_mesh_kernel(p, ctx->sm_count, ctx->stream, false, false, mesh_cap);
            ctx->launches += 1;
        }
        CK(cudaEventRecord(ctx->ev_mm, ctx->stream));
        if (pipelined) {
            launch_emit_kernel_pipelined(p, ctx->sm_count, emit_ctas, ctx->emit_stream);
            ctx->launches += 1;
            CK(cudaEventRecord(ctx->ev_join, ctx->emit_stream));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        } else if (!count_only) { // vertices and indices of the recorded sides (returns at once if a chunk did not fit)
            // the arena is rewritten from here on: an asynchronous fetch of the previous meshes has to be through
            if (ctx->copy_pending) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done, 0));
            launch_emit_kernel(p, ctx->sm_count, ctx->stream);
            ctx->launches += 1;
        }
        CK(cudaEventRecord(ctx->ev_m1, ctx->stream));
        CK(cudaGetLastError());
        // one copy brings back the allocator state, the error flags and (behind it, in the same pinned block) the length of the
        // candidate list
        CK(cudaMemcpyAsync(r.h_counters, r.counters, sizeof(MeshCounters), cudaMemcpyDeviceToHost, ctx->stream));
        if (p.n_work_dev) CK(cudaMemcpyAsync(r.h_counters + 1, p.n_work_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev_m0, ctx->ev_m1));
        if (!count_only) {
            ctx->mesh_ms += ms;
            ctx->mesh_launches += 1;
            float ems = 0.f;
            CK(cudaEventElapsedTime(&ems, ctx->ev_mm, ctx->ev_m1));
            ctx->emit_ms += ems;
        }
        if (r.h_counters->flags & kFlagUnknownId)
            return fail(ctx, BXW_E_UNKNOWN_ID, "a voxel references a block id without a registry definition");
        if (r.h_counters->flags & kFlagPipelineTimeout)
            return fail(ctx, BXW_E_CUDA, "pipelined emit_kernel timed out waiting for mesh_kernel (internal error)");
        const unsigned long long need_v = r.h_counters->v_alloc, need_i = r.h_counters->i_alloc;
        if (count_only || (r.h_counters->flags & kFlagArenaOverflow)) {
            // size the arena from the measured need (+ slack so small edits can append) and run again
            int rc = arena_resize(ctx, r, need_v + need_v / 8 + 4096, need_i + need_i / 8 + 8192, base.v_alloc, base.i_alloc);
            if (rc) return rc;
            const unsigned long long need_f = r.h_counters->f_alloc;
            rc = faces_resize(ctx, r, need_f + need_f / 8 + 4096);
            if (rc) return rc;
            if (append) {
                CK(cudaMemcpyAsync(r.spans, r.spans_bak, r.n_work * sizeof(bxw_mesh_span), cudaMemcpyDeviceToDevice, ctx->stream));
                CK(cudaMemcpyAsync(r.span_cap, r.span_cap_bak, r.n_work * sizeof(uint2), cudaMemcpyDeviceToDevice, ctx->stream));
            }
            continue;
        }
        {
            // faces per work item of this pass: what the next pass's pipelining decision goes by
            const double items = p.n_work_dev ? (double)*reinterpret_cast<const uint32_t *>(r.h_counters + 1) : (double)n_work;
            ctx->last_faces_per_item = items > 0 ? (double)r.h_counters->f_alloc / items : 0.0;
        }
        if (p.n_work_dev) { // whole-region pass over a device-built candidate list: remember its length (statistics)
            ctx->last_candidates = *reinterpret_cast<const uint32_t *>(r.h_counters + 1);
            ctx->last_candidates_of = n_work;
        } else if (!work) { // whole region without a candidate list: everything was read
            ctx->last_candidates = n_work;
            ctx->last_candidates_of = n_work;
        }
        return BXW_OK;
    }
    return fail(ctx, BXW_E_NOSPACE, "mesh arena sizing did not converge");
}

int ensure_io(bxw_ctx *ctx, size_t bytes_a, size_t bytes_b) {
    if (bytes_a > ctx->io_a_cap) {
        cudaFree(ctx->io_a);
        ctx->io_a = nullptr;
        ctx->io_a_cap = 0;
        CK(cudaMalloc(&ctx->io_a, bytes_a));
        ctx->io_a_cap = bytes_a;
    }
    if (bytes_b > ctx->io_b_cap) {
        cudaFree(ctx->io_b);
        ctx->io_b = nullptr;
        ctx->io_b_cap = 0;
        CK(cudaMalloc(&ctx->io_b, bytes_b));
        ctx->io_b_cap = bytes_b;
    }
    return BXW_OK;
}

int ensure_io_c(bxw_ctx *ctx, size_t bytes) {
    if (bytes > ctx->io_c_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->io_c);
        ctx->io_c = nullptr;
        ctx->io_c_cap = 0;
        CK(cudaMalloc(&ctx->io_c, bytes));
        ctx->io_c_cap = bytes;
    }
    return BXW_OK;
}

// host RLE words -> device dense chunks (decompress_rle, lib.rs:304-345)
int rle_decode_into(bxw_ctx *ctx, const uint32_t *words, const uint64_t *offsets, size_t n, uint32_t *d_dense) {
    // the kernel reads words[offsets[c] .. offsets[c+1]) out of an upload of offsets[n] words: the offsets must not go backwards
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, BXW_E_INVALID, "RLE offsets must be non-decreasing (chunk %zu)", i);
    const size_t nwords = offsets[n];
    const size_t a_bytes = (nwords ? nwords : 1) * sizeof(uint32_t);
    const size_t b_bytes = (n + 1) * sizeof(uint64_t) + n * sizeof(uint32_t);
    int rc = ensure_io(ctx, a_bytes, b_bytes);
    if (rc) return rc;
    uint32_t *d_words = (uint32_t *)ctx->io_a;
    uint64_t *d_offs = (uint64_t *)ctx->io_b;
    uint32_t *d_status = (uint32_t *)(d_offs + n + 1);
    CK(cudaMemcpyAsync(d_words, words, nwords * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_offs, offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    launch_rle_decompress(d_words, d_offs, n, d_dense, d_status, ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
    std::vector<uint32_t> st(n);
    CK(cudaMemcpyAsync(st.data(), d_status, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < n; i++)
        if (st[i]) return fail(ctx, BXW_E_BAD_RLE, "chunk %zu: RLE stream does not decode (RleDecompressError %u)", i, st[i]);
    return BXW_OK;
}

int ensure_scratch(bxw_ctx *ctx) {
    if (ctx->scratch.live) return BXW_OK;
    return region_alloc(ctx, ctx->scratch, 0, 0, 0, 1, 1, 1);
}

size_t storage_index(const Region &r, int32_t lx, int32_t ly, int32_t lz) {
    return (size_t)(lx + 1) + (size_t)r.SX * ((size_t)(lz + 1) + (size_t)r.SZ * (size_t)(ly + 1));
}

bool local_ok(const Region &r, int32_t lx, int32_t ly, int32_t lz) {
    return lx >= -1 && ly >= -1 && lz >= -1 && lx <= (int32_t)r.nx && ly <= (int32_t)r.ny && lz <= (int32_t)r.nz;
}

} // namespace

// =====================================================================================================
extern "C" {

int bxw_create(int device, bxw_ctx **out) {
    if (!out) return BXW_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return BXW_E_CUDA;
    bxw_ctx *ctx = new (std::nothrow) bxw_ctx();
    if (!ctx) return BXW_E_NOMEM;
    ctx->device = device;
    auto bail = [&](int code) {
        bxw_destroy(ctx);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess) return bail(BXW_E_CUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(BXW_E_CUDA);
    if (prop.major < 10) { // sm_100a cubins only
        return bail(BXW_E_CUDA);
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(BXW_E_CUDA);
    ctx->stream = ctx->own_stream;
    if (cudaEventCreate(&ctx->ev_a) != cudaSuccess || cudaEventCreate(&ctx->ev_b) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_m0) != cudaSuccess || cudaEventCreate(&ctx->ev_m1) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_mm) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_g0) != cudaSuccess || cudaEventCreate(&ctx->ev_g1) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_g2) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->emit_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_gen_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_gen_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_mesh_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_copy_done, cudaEventDisableTiming) != cudaSuccess)
        return bail(BXW_E_CUDA);
    MeshTables *t = new (std::nothrow) MeshTables();
    if (!t) return bail(BXW_E_NOMEM);
    build_mesh_tables(*t);
    if (!t->aomask_ok) { // the popcount AO shortcut of the mesher relies on this
        delete t;
        return bail(BXW_E_INVALID);
    }
    cudaError_t e = cudaMalloc(&ctx->d_tables, sizeof(MeshTables));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_tables, t, sizeof(MeshTables), cudaMemcpyHostToDevice);
    delete t;
    if (e != cudaSuccess) return bail(BXW_E_CUDA);
    *out = ctx;
    return BXW_OK;
}

void bxw_destroy(bxw_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    region_free(ctx->region);
    region_free(ctx->scratch);
    cudaFree(ctx->d_kind);
    cudaFree(ctx->d_texf);
    cudaFree(ctx->d_color);
    cudaFree(ctx->d_tables);
    cudaFree(ctx->d_gen);
    cudaFree(ctx->io_a);
    cudaFree(ctx->io_b);
    cudaFree(ctx->io_c);
    cudaFree(ctx->d_tg);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->ev_m0) cudaEventDestroy(ctx->ev_m0);
    if (ctx->ev_m1) cudaEventDestroy(ctx->ev_m1);
    if (ctx->ev_mm) cudaEventDestroy(ctx->ev_mm);
    if (ctx->ev_g0) cudaEventDestroy(ctx->ev_g0);
    if (ctx->ev_g1) cudaEventDestroy(ctx->ev_g1);
    if (ctx->ev_g2) cudaEventDestroy(ctx->ev_g2);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->emit_stream) {
        cudaStreamSynchronize(ctx->emit_stream);
        cudaStreamDestroy(ctx->emit_stream);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_gen_fork) cudaEventDestroy(ctx->ev_gen_fork);
    if (ctx->ev_gen_join) cudaEventDestroy(ctx->ev_gen_join);
    if (ctx->ev_mesh_done) cudaEventDestroy(ctx->ev_mesh_done);
    if (ctx->ev_copy_done) cudaEventDestroy(ctx->ev_copy_done);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *bxw_last_error(const bxw_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int bxw_set_stream(bxw_ctx *ctx, void *cuda_stream) {
    if (!ctx) return BXW_E_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return BXW_OK;
}

int bxw_set_option(bxw_ctx *ctx, int option, int value) {
    if (!ctx) return BXW_E_INVALID;
    if (option == BXW_OPT_CHUNK_SUMMARIES) {
        ctx->use_summaries = value != 0;
        return BXW_OK;
    }
    if (option == BXW_OPT_MESH_FROM_HEIGHT_MAP) {
        ctx->mesh_from_hmap = value != 0;
        return BXW_OK;
    }
    if (option == BXW_OPT_PIPELINED_MESHER) {
        ctx->pipeline_mode = value < 0 ? 0 : (value > 2 ? 2 : value);
        return BXW_OK;
    }
    if (option == BXW_OPT_CUBES_ONLY_MESHER) {
        ctx->lean_mesher = value != 0;
        return BXW_OK;
    }
    if (option == BXW_OPT_ELIDE_UNIFORM) {
        ctx->elide_uniform = value != 0;
        return BXW_OK;
    }
    return fail(ctx, BXW_E_INVALID, "unknown option %d", option);
}

int bxw_synchronize(bxw_ctx *ctx) {
    if (!ctx) return BXW_E_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

uint64_t bxw_kernel_launches(const bxw_ctx *ctx) { return ctx ? ctx->launches : 0; }

int bxw_set_registry(bxw_ctx *ctx, const bxw_block_def *defs, uint32_t n) {
    if (!ctx || !defs || n == 0 || n > 65536) return fail(ctx, BXW_E_INVALID, "bad registry");
    std::vector<uint8_t> kind(n);
    std::vector<float> texf((size_t)n * 6), color((size_t)n * 3);
    for (uint32_t i = 0; i < n; i++) {
        kind[i] = defs[i].present ? (defs[i].mesh_kind ? 2 : 1) : 0;
        for (int k = 0; k < 6; k++) texf[(size_t)i * 6 + k] = (float)defs[i].tex[k]; // `texid as f32` (voxmesh.rs:167)
        for (int k = 0; k < 3; k++) color[(size_t)i * 3 + k] = defs[i].debug_color[k];
    }
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_kind);
    cudaFree(ctx->d_texf);
    cudaFree(ctx->d_color);
    ctx->d_kind = nullptr;
    ctx->d_texf = ctx->d_color = nullptr;
    CK(cudaMalloc(&ctx->d_kind, n));
    CK(cudaMalloc(&ctx->d_texf, (size_t)n * 6 * sizeof(float)));
    CK(cudaMalloc(&ctx->d_color, (size_t)n * 3 * sizeof(float)));
    CK(cudaMemcpy(ctx->d_kind, kind.data(), n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_texf, texf.data(), texf.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_color, color.data(), color.size() * sizeof(float), cudaMemcpyHostToDevice));
    ctx->reg_n = n;
    ctx->reg_host.assign(defs, defs + n);
    return BXW_OK;
}

int bxw_set_gen_ids(bxw_ctx *ctx, uint16_t void_id, uint16_t grass, uint16_t dirt, uint16_t stone, uint16_t snow_grass) {
    if (!ctx) return BXW_E_INVALID;
    ctx->gen_ids[0] = void_id;
    ctx->gen_ids[1] = grass;
    ctx->gen_ids[2] = dirt;
    ctx->gen_ids[3] = stone;
    ctx->gen_ids[4] = snow_grass;
    ctx->have_gen_ids = true;
    return BXW_OK;
}

int bxw_gen_chunk(bxw_ctx *ctx, uint64_t seed, int32_t cx, int32_t cy, int32_t cz, uint32_t *blocks_yzx) {
    if (!ctx || !blocks_yzx) return fail(ctx, BXW_E_INVALID, "null argument");
    int rc = ensure_scratch(ctx);
    if (rc) return rc;
    // reuse the scratch region's storage slot 0 as a 1x1x1 box positioned at (cx,cy,cz)
    Region box = ctx->scratch;
    box.SX = box.SY = box.SZ = 1;
    box.cx0 = cx + 1;
    box.cy0 = cy + 1;
    box.cz0 = cz + 1;
    box.raw = nullptr;
    rc = region_generate(ctx, box, seed, 64, false); // no elision: the caller wants the voxels
    ctx->scratch.cells = box.cells; // region_generate may have grown the cell table
    ctx->scratch.cells_cap = box.cells_cap;
    if (rc) return rc;
    CK(cudaMemcpyAsync(blocks_yzx, box.blocks, kChunkVoxels * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_mesh_chunk27(bxw_ctx *ctx, const uint32_t *const dense[27], bxw_mesh_span *counts) {
    if (!ctx || !dense || !counts) return fail(ctx, BXW_E_INVALID, "null argument");
    int rc = ensure_scratch(ctx);
    if (rc) return rc;
    Region &r = ctx->scratch;
    CK(cudaMemsetAsync(r.summary, 0, r.n_storage() * sizeof(unsigned long long), ctx->stream));
    for (int i = 0; i < 27; i++) {
        if (!dense[i]) return fail(ctx, BXW_E_INVALID, "dense[%d] is null", i);
        // slot rx + 3 rz + 9 ry (voxmesh.rs:72) is exactly the 3x3x3 storage order sx + 3*(sz + 3*sy)
        CK(cudaMemcpyAsync(r.blocks + (size_t)i * kChunkVoxels, dense[i], kChunkVoxels * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    launch_xface_extract(r.blocks, r.xface, nullptr, 0, 27, ctx->stream);
    ctx->launches += 1;
    rc = region_mesh(ctx, r, r.work, r.work_span, 1, false);
    if (rc) return rc;
    CK(cudaMemcpyAsync(counts, r.spans, sizeof(bxw_mesh_span), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_mesh_chunk27_rle(bxw_ctx *ctx, const uint32_t *const rle[27], const size_t rle_len[27], bxw_mesh_span *counts) {
    if (!ctx || !rle || !rle_len || !counts) return fail(ctx, BXW_E_INVALID, "null argument");
    int rc = ensure_scratch(ctx);
    if (rc) return rc;
    Region &r = ctx->scratch;
    // upload the 27 RLE streams back to back, decode on the device straight into the 3x3x3 scratch storage
    std::vector<uint64_t> offs(28, 0);
    for (int i = 0; i < 27; i++) {
        if (!rle[i]) return fail(ctx, BXW_E_INVALID, "rle[%d] is null", i);
        offs[i + 1] = offs[i] + rle_len[i];
    }
    std::vector<uint32_t> words(offs[27] ? offs[27] : 1);
    for (int i = 0; i < 27; i++) std::memcpy(words.data() + offs[i], rle[i], rle_len[i] * sizeof(uint32_t));
    CK(cudaMemsetAsync(r.summary, 0, r.n_storage() * sizeof(unsigned long long), ctx->stream));
    rc = rle_decode_into(ctx, words.data(), offs.data(), 27, r.blocks);
    if (rc) return rc;
    launch_xface_extract(r.blocks, r.xface, nullptr, 0, 27, ctx->stream);
    ctx->launches += 1;
    rc = region_mesh(ctx, r, r.work, r.work_span, 1, false);
    if (rc) return rc;
    CK(cudaMemcpyAsync(counts, r.spans, sizeof(bxw_mesh_span), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_mesh_fetch(bxw_ctx *ctx, bxw_vertex *vertices, uint32_t *indices) {
    if (!ctx || !ctx->scratch.live) return fail(ctx, BXW_E_INVALID, "no mesh to fetch");
    Region &r = ctx->scratch;
    bxw_mesh_span sp;
    CK(cudaMemcpyAsync(&sp, r.spans, sizeof sp, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (sp.v_count && vertices)
        CK(cudaMemcpyAsync(vertices, r.varena + sp.v_off, (size_t)sp.v_count * sizeof(bxw_vertex), cudaMemcpyDeviceToHost, ctx->stream));
    if (sp.i_count && indices)
        CK(cudaMemcpyAsync(indices, r.iarena + sp.i_off, (size_t)sp.i_count * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_is_chunk_trivial(bxw_ctx *ctx, const uint32_t *rle, size_t rle_len, int *out_trivial) {
    // voxmesh.rs:16-25: uniform chunk (3 RLE words, v v n) whose block has no mesh
    if (!ctx || !rle || !out_trivial) return fail(ctx, BXW_E_INVALID, "null argument");
    if (ctx->reg_host.empty()) return fail(ctx, BXW_E_NO_REGISTRY, "bxw_set_registry not called");
    *out_trivial = 0;
    if (rle_len == 3 && rle[0] == rle[1]) {
        const uint32_t id = rle[0] >> 16;
        if (id >= ctx->reg_n || !ctx->reg_host[id].present) return fail(ctx, BXW_E_UNKNOWN_ID, "unknown block id %u", id);
        if (ctx->reg_host[id].mesh_kind == 0) *out_trivial = 1;
    }
    return BXW_OK;
}

size_t bxw_rle_bound(size_t n_chunks) { return n_chunks * (size_t)(kChunkVoxels + kChunkVoxels / 2 + 1); }

int bxw_rle_compress(bxw_ctx *ctx, const uint32_t *dense, size_t n_chunks, uint32_t *out_words, uint64_t *out_offsets) {
    if (!ctx || !dense || !out_words || !out_offsets || n_chunks == 0) return fail(ctx, BXW_E_INVALID, "null argument");
    const size_t dense_bytes = n_chunks * kChunkVoxels * sizeof(uint32_t);
    const size_t meta_bytes = (n_chunks + 1) * sizeof(uint64_t) + n_chunks * sizeof(uint32_t);
    int rc = ensure_io(ctx, dense_bytes, meta_bytes);
    if (rc) return rc;
    uint32_t *d_dense = (uint32_t *)ctx->io_a;
    uint64_t *d_offs = (uint64_t *)ctx->io_b;
    uint32_t *d_len = (uint32_t *)(d_offs + n_chunks + 1);
    CK(cudaMemcpyAsync(d_dense, dense, dense_bytes, cudaMemcpyHostToDevice, ctx->stream));
    launch_rle_compress(d_dense, nullptr, nullptr, n_chunks, d_len, d_offs, nullptr, 0, ctx->stream);
    CK(cudaMemcpyAsync(out_offsets, d_offs, (n_chunks + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const size_t nwords = out_offsets[n_chunks];
    if (nwords * sizeof(uint32_t) > ctx->io_c_cap) {
        cudaFree(ctx->io_c);
        ctx->io_c = nullptr;
        ctx->io_c_cap = 0;
        CK(cudaMalloc(&ctx->io_c, nwords * sizeof(uint32_t)));
        ctx->io_c_cap = nwords * sizeof(uint32_t);
    }
    launch_rle_compress(d_dense, nullptr, nullptr, n_chunks, d_len, d_offs, (uint32_t *)ctx->io_c, 1, ctx->stream);
    ctx->launches += 3;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_words, ctx->io_c, nwords * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_rle_decompress(bxw_ctx *ctx, const uint32_t *words, const uint64_t *offsets, size_t n_chunks, uint32_t *out_dense) {
    if (!ctx || !words || !offsets || !out_dense || n_chunks == 0) return fail(ctx, BXW_E_INVALID, "null argument");
    const size_t dense_bytes = n_chunks * kChunkVoxels * sizeof(uint32_t);
    if (dense_bytes > ctx->io_c_cap) {
        cudaFree(ctx->io_c);
        ctx->io_c = nullptr;
        ctx->io_c_cap = 0;
        CK(cudaMalloc(&ctx->io_c, dense_bytes));
        ctx->io_c_cap = dense_bytes;
    }
    int rc = rle_decode_into(ctx, words, offsets, n_chunks, (uint32_t *)ctx->io_c);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out_dense, ctx->io_c, dense_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_region_create(bxw_ctx *ctx, int32_t cx0, int32_t cy0, int32_t cz0, uint32_t nx, uint32_t ny, uint32_t nz) {
    if (!ctx || nx == 0 || ny == 0 || nz == 0) return fail(ctx, BXW_E_INVALID, "bad region dims");
    return region_alloc(ctx, ctx->region, cx0, cy0, cz0, nx, ny, nz);
}

int bxw_region_destroy(bxw_ctx *ctx) {
    if (!ctx) return BXW_E_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
    ctx->copy_pending = false;
    region_free(ctx->region);
    return BXW_OK;
}

int bxw_region_generate(bxw_ctx *ctx, uint64_t seed, int precision) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    return region_generate(ctx, ctx->region, seed, precision, ctx->elide_uniform);
}

int bxw_region_generate_part(bxw_ctx *ctx, uint64_t seed, int precision, int part) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    if (part != 1 && part != 2) return fail(ctx, BXW_E_INVALID, "part must be 1 (slab-boundary columns) or 2 (the rest)");
    return region_generate(ctx, ctx->region, seed, precision, ctx->elide_uniform, part);
}

int bxw_region_heightmap(bxw_ctx *ctx, int32_t *height, uint8_t *elevation, double *raw_height) {
    if (!ctx || !ctx->region.live || !ctx->region.have_hmap) return fail(ctx, BXW_E_INVALID, "no generated region");
    Region &r = ctx->region;
    const size_t n = (size_t)r.SX * 32 * r.SZ * 32;
    if (raw_height) {
        // A getter must not touch the region: evaluate the column function again (stand-alone height map kernel, the seed and
        // precision of the region's last generate) into scratch buffers; blocks, planes, summaries and hmap stay as they are.
        GenParams gp;
        int rc = gen_setup(ctx, r, r.gen_seed, r.gen_precision, gp);
        if (rc) return rc;
        if (!r.raw) CK(cudaMalloc(&r.raw, n * sizeof(double)));
        rc = ensure_io_c(ctx, n * sizeof(int32_t));
        if (rc) return rc;
        gp.hmap = (int32_t *)ctx->io_c;
        gp.raw = r.raw;
        launch_heightmap_kernel(gp, r.gen_precision, ctx->stream);
        ctx->launches += 1;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(raw_height, r.raw, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    std::vector<int32_t> packed(n);
    CK(cudaMemcpyAsync(packed.data(), r.hmap, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < n; i++) {
        if (height) height[i] = packed[i] >> 2;
        if (elevation) elevation[i] = (uint8_t)(packed[i] & 3);
    }
    return BXW_OK;
}

int bxw_region_upload_chunk(bxw_ctx *ctx, int32_t lx, int32_t ly, int32_t lz, const uint32_t *blocks_yzx) {
    if (!ctx || !ctx->region.live || !blocks_yzx) return fail(ctx, BXW_E_INVALID, "no region / null");
    Region &r = ctx->region;
    if (!local_ok(r, lx, ly, lz)) return fail(ctx, BXW_E_INVALID, "chunk (%d,%d,%d) outside region+shell", lx, ly, lz);
    const uint32_t cid = (uint32_t)storage_index(r, lx, ly, lz);
    CK(cudaMemcpyAsync(r.page + cid, &cid, sizeof cid, cudaMemcpyHostToDevice, ctx->stream)); // every voxel is overwritten: materialised
    CK(cudaMemsetAsync(r.blk_loaded + cid, 1, 1, ctx->stream));
    CK(cudaMemcpyAsync(r.blocks + storage_index(r, lx, ly, lz) * kChunkVoxels, blocks_yzx, kChunkVoxels * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(r.summary + storage_index(r, lx, ly, lz), 0, sizeof(unsigned long long), ctx->stream)); // no longer known uniform
    launch_xface_extract(r.blocks, r.xface, nullptr, (uint32_t)storage_index(r, lx, ly, lz), 1, ctx->stream);
    ctx->launches += 1;
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_region_download_chunk(bxw_ctx *ctx, int32_t lx, int32_t ly, int32_t lz, uint32_t *blocks_yzx) {
    if (!ctx || !ctx->region.live || !blocks_yzx) return fail(ctx, BXW_E_INVALID, "no region / null");
    Region &r = ctx->region;
    if (!local_ok(r, lx, ly, lz)) return fail(ctx, BXW_E_INVALID, "chunk (%d,%d,%d) outside region+shell", lx, ly, lz);
    uint32_t pg = 0; // a chunk that is not materialised is read from its shared uniform page
    CK(cudaMemcpyAsync(&pg, r.page + storage_index(r, lx, ly, lz), sizeof pg, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpyAsync(blocks_yzx, r.blocks + (size_t)pg * kChunkVoxels, kChunkVoxels * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

// local chunk coordinates -> storage chunk indices on the device (io_b tail); returns the device pointer
static int upload_chunk_ids(bxw_ctx *ctx, const Region &r, const int32_t *lpos, size_t n, size_t b_prefix_bytes, uint32_t **d_ids) {
    std::vector<uint32_t> ids(n);
    for (size_t i = 0; i < n; i++) {
        const int32_t lx = lpos[3 * i], ly = lpos[3 * i + 1], lz = lpos[3 * i + 2];
        if (!local_ok(r, lx, ly, lz)) return fail(ctx, BXW_E_INVALID, "chunk (%d,%d,%d) outside region+shell", lx, ly, lz);
        ids[i] = (uint32_t)storage_index(r, lx, ly, lz);
    }
    *d_ids = (uint32_t *)((char *)ctx->io_b + b_prefix_bytes);
    CK(cudaMemcpyAsync(*d_ids, ids.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // ids is a stack-owned staging vector
    return BXW_OK;
}

int bxw_region_export_rle(bxw_ctx *ctx, const int32_t *lpos, size_t n_chunks, uint32_t *out_words, size_t cap_words,
                          uint64_t *out_offsets) {
    if (!ctx || !ctx->region.live || !lpos || !out_offsets || n_chunks == 0) return fail(ctx, BXW_E_INVALID, "no region / null argument");
    Region &r = ctx->region;
    const size_t meta_bytes = (n_chunks + 1) * sizeof(uint64_t) + n_chunks * sizeof(uint32_t);
    int rc = ensure_io(ctx, 16, meta_bytes + n_chunks * sizeof(uint32_t));
    if (rc) return rc;
    uint64_t *d_offs = (uint64_t *)ctx->io_b;
    uint32_t *d_len = (uint32_t *)(d_offs + n_chunks + 1);
    uint32_t *d_ids = nullptr;
    rc = upload_chunk_ids(ctx, r, lpos, n_chunks, meta_bytes, &d_ids);
    if (rc) return rc;
    launch_rle_compress(r.blocks, d_ids, r.page, n_chunks, d_len, d_offs, nullptr, 0, ctx->stream);
    ctx->launches += 2;
    CK(cudaMemcpyAsync(out_offsets, d_offs, (n_chunks + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const size_t nwords = out_offsets[n_chunks];
    if (!out_words) return BXW_OK; // size query
    if (nwords > cap_words) return fail(ctx, BXW_E_NOSPACE, "RLE export needs %zu words, buffer holds %zu", nwords, cap_words);
    rc = ensure_io_c(ctx, nwords * sizeof(uint32_t));
    if (rc) return rc;
    launch_rle_compress(r.blocks, d_ids, r.page, n_chunks, d_len, d_offs, (uint32_t *)ctx->io_c, 1, ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_words, ctx->io_c, nwords * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_region_import_rle(bxw_ctx *ctx, const int32_t *lpos, size_t n_chunks, const uint32_t *words, const uint64_t *offsets) {
    if (!ctx || !ctx->region.live || !lpos || !words || !offsets || n_chunks == 0)
        return fail(ctx, BXW_E_INVALID, "no region / null argument");
    Region &r = ctx->region;
    for (size_t i = 0; i < n_chunks; i++)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, BXW_E_INVALID, "offsets must be non-decreasing");
    int rc = ensure_io_c(ctx, n_chunks * kChunkVoxels * sizeof(uint32_t));
    if (rc) return rc;
    // decode into scratch first: a stream that does not decode must leave the region untouched
    rc = rle_decode_into(ctx, words, offsets, n_chunks, (uint32_t *)ctx->io_c);
    if (rc) return rc;
    const size_t meta_bytes = (n_chunks + 1) * sizeof(uint64_t) + n_chunks * sizeof(uint32_t);
    rc = ensure_io(ctx, 16, meta_bytes + n_chunks * sizeof(uint32_t));
    if (rc) return rc;
    uint32_t *d_ids = nullptr;
    rc = upload_chunk_ids(ctx, r, lpos, n_chunks, meta_bytes, &d_ids);
    if (rc) return rc;
    launch_chunk_scatter((const uint32_t *)ctx->io_c, d_ids, n_chunks, r.blocks, r.xface, r.summary, r.page, r.blk_loaded, r.SX, r.SY, r.SZ, r.dirty, ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_region_mesh_reserve(bxw_ctx *ctx, uint64_t max_vertices, uint64_t max_indices) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    int rc = arena_resize(ctx, ctx->region, max_vertices, max_indices, 0, 0);
    if (rc == BXW_OK) rc = faces_resize(ctx, ctx->region, max_vertices / 3 + 4096); // a side has at least 3 vertices
    if (rc == BXW_OK) ctx->region.reserved = true;
    return rc;
}

int bxw_region_mesh(bxw_ctx *ctx, uint64_t *arena_vertices, uint64_t *arena_indices) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    int rc = region_mesh(ctx, r, nullptr, nullptr, r.n_work, false); // implicit interior enumeration
    if (rc) return rc;
    CK(cudaMemsetAsync(r.dirty, 0, r.n_work, ctx->stream));
    CK(cudaMemsetAsync(r.mesh_loaded, 1, r.n_work, ctx->stream));
    if (arena_vertices) *arena_vertices = r.h_counters->v_alloc;
    if (arena_indices) *arena_indices = r.h_counters->i_alloc;
    return BXW_OK;
}

int bxw_region_generate_and_mesh(bxw_ctx *ctx, uint64_t seed, int precision, uint64_t *arena_vertices, uint64_t *arena_indices) {
    int rc = bxw_region_generate(ctx, seed, precision);
    if (rc) return rc;
    // The blocks were just written from the column parameters and nothing has edited them, so the mesher may derive its
    // class table from the height map (34x34 columns per chunk) instead of reading the voxels back
    // (BXW_OPT_MESH_FROM_HEIGHT_MAP). With chunk summaries the voxel path only reads the chunks near the surface and is the
    // faster of the two on B200, so it is the default.
    Region &r = ctx->region;
    rc = region_mesh(ctx, r, nullptr, nullptr, r.n_work, false, ctx->mesh_from_hmap);
    if (rc) return rc;
    CK(cudaMemsetAsync(r.dirty, 0, r.n_work, ctx->stream));
    CK(cudaMemsetAsync(r.mesh_loaded, 1, r.n_work, ctx->stream));
    if (arena_vertices) *arena_vertices = r.h_counters->v_alloc;
    if (arena_indices) *arena_indices = r.h_counters->i_alloc;
    return BXW_OK;
}

int bxw_region_mesh_sphere(bxw_ctx *ctx, int32_t ax, int32_t ay, int32_t az, uint32_t radius, uint32_t *n_meshed,
                           uint64_t *arena_vertices, uint64_t *arena_indices) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    // iter_coords_to_load (worldmgr.rs:734-742): the cube -r..=r filtered by x^2+y^2+z^2 <= r^2; then nearest first
    struct Item {
        int64_t d2;
        uint32_t chunk, span;
    };
    std::vector<Item> items;
    const int64_t rr = (int64_t)radius * radius;
    for (uint32_t iy = 0; iy < r.ny; iy++)
        for (uint32_t iz = 0; iz < r.nz; iz++)
            for (uint32_t ix = 0; ix < r.nx; ix++) {
                const int64_t dx = (int64_t)r.cx0 + ix - ax, dy = (int64_t)r.cy0 + iy - ay, dz = (int64_t)r.cz0 + iz - az;
                const int64_t d2 = dx * dx + dy * dy + dz * dz;
                if (d2 <= rr)
                    items.push_back({d2, (uint32_t)((ix + 1) + r.SX * ((iz + 1) + r.SZ * (iy + 1))), (uint32_t)(ix + r.nx * (iz + r.nz * iy))});
            }
    std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.d2 < b.d2; });
    const uint32_t n = (uint32_t)items.size();
    if (n_meshed) *n_meshed = n;
    CK(cudaMemsetAsync(r.spans, 0, r.n_work * sizeof(bxw_mesh_span), ctx->stream));
    CK(cudaMemsetAsync(r.span_cap, 0, r.n_work * sizeof(uint2), ctx->stream));
    if (n == 0) {
        if (arena_vertices) *arena_vertices = 0;
        if (arena_indices) *arena_indices = 0;
        return BXW_OK;
    }
    std::vector<uint32_t> w(n), ws(n);
    for (uint32_t i = 0; i < n; i++) {
        w[i] = items[i].chunk;
        ws[i] = items[i].span;
    }
    // the dirty work-list buffers double as the sphere list (both are at most n_work entries)
    CK(cudaMemcpyAsync(r.dirty_work, w.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(r.dirty_span, ws.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int rc = region_mesh(ctx, r, r.dirty_work, r.dirty_span, n, false);
    if (rc) return rc;
    if (arena_vertices) *arena_vertices = r.h_counters->v_alloc;
    if (arena_indices) *arena_indices = r.h_counters->i_alloc;
    return BXW_OK;
}

static int sched_ensure(bxw_ctx *ctx, Region &r, uint32_t n_anchors) {
    const size_t nst = r.n_storage();
    if (!r.sched_code) {
        // one histogram bin per squared distance inside the box (anchors far outside share the last bin)
        const unsigned long long d2 = (unsigned long long)r.SX * r.SX + (unsigned long long)r.SY * r.SY + (unsigned long long)r.SZ * r.SZ + 2ull;
        r.sched_bins = (uint32_t)std::min<unsigned long long>(d2, 1ull << 24);
        CK(cudaMalloc(&r.sched_code, nst));
        CK(cudaMalloc(&r.sched_key, nst * sizeof(uint32_t)));
        CK(cudaMalloc(&r.sched_hist, 2ull * r.sched_bins * sizeof(uint32_t)));
        CK(cudaMalloc(&r.sched_totals, 2 * sizeof(uint32_t)));
        CK(cudaMalloc(&r.sched_order, nst * sizeof(uint32_t)));
        CK(cudaMalloc(&r.sched_gen, nst * sizeof(uint32_t)));
        CK(cudaMalloc(&r.sched_mesh, r.n_work * sizeof(uint32_t)));
        CK(cudaMalloc(&r.sched_mesh_span, r.n_work * sizeof(uint32_t)));
        CK(cudaMalloc(&r.sched_unmesh, r.n_work * sizeof(uint32_t)));
        CK(cudaMalloc(&r.sched_counts, 4 * sizeof(uint32_t)));
        CK(cudaMalloc(&r.sched_deltas, nst * sizeof(bxw_chunk_delta)));
    }
    if (n_anchors > r.sched_anchor_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(r.sched_anchors);
        r.sched_anchors = nullptr;
        r.sched_anchor_cap = 0;
        CK(cudaMalloc(&r.sched_anchors, (size_t)n_anchors * sizeof(bxw_load_anchor)));
        r.sched_anchor_cap = n_anchors;
    }
    return BXW_OK;
}

static SchedParams sched_params(Region &r, uint32_t n_anchors) {
    SchedParams p;
    p.SX = r.SX;
    p.SY = r.SY;
    p.SZ = r.SZ;
    p.cx_min = r.cx0 - 1;
    p.cy_min = r.cy0 - 1;
    p.cz_min = r.cz0 - 1;
    p.n_storage = (uint32_t)r.n_storage();
    p.anchors = r.sched_anchors;
    p.n_anchors = n_anchors;
    p.blk_loaded = r.blk_loaded;
    p.mesh_loaded = r.mesh_loaded;
    p.code = r.sched_code;
    p.key = r.sched_key;
    p.hist = r.sched_hist;
    p.n_bins = r.sched_bins;
    p.totals = r.sched_totals;
    p.deltas = r.sched_deltas;
    p.order = r.sched_order;
    p.gen_list = r.sched_gen;
    p.mesh_list = r.sched_mesh;
    p.mesh_span = r.sched_mesh_span;
    p.unmesh_span = r.sched_unmesh;
    p.list_counts = r.sched_counts;
    return p;
}

int bxw_region_unload_all(bxw_ctx *ctx) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    CK(cudaMemsetAsync(r.blk_loaded, 0, r.n_storage(), ctx->stream));
    CK(cudaMemsetAsync(r.mesh_loaded, 0, r.n_work, ctx->stream));
    CK(cudaMemsetAsync(r.spans, 0, r.n_work * sizeof(bxw_mesh_span), ctx->stream));
    CK(cudaMemsetAsync(r.span_cap, 0, r.n_work * sizeof(uint2), ctx->stream));
    CK(cudaMemsetAsync(r.counters, 0, sizeof(MeshCounters), ctx->stream));
    CK(cudaMemsetAsync(r.dirty, 0, r.n_work, ctx->stream));
    r.sched_n_unload = r.sched_n_load = 0;
    r.streaming = true;
    return BXW_OK;
}

int bxw_region_update_anchors(bxw_ctx *ctx, const bxw_load_anchor *anchors, uint32_t n_anchors, uint32_t *n_unload, uint32_t *n_load) {
    if (!ctx || !ctx->region.live || (n_anchors && !anchors)) return fail(ctx, BXW_E_INVALID, "no region / null");
    Region &r = ctx->region;
    int rc = sched_ensure(ctx, r, n_anchors ? n_anchors : 1);
    if (rc) return rc;
    r.streaming = true;
    if (n_anchors) CK(cudaMemcpyAsync(r.sched_anchors, anchors, (size_t)n_anchors * sizeof(bxw_load_anchor), cudaMemcpyHostToDevice, ctx->stream));
    launch_delta_build(sched_params(r, n_anchors), ctx->stream);
    ctx->launches += 3;
    CK(cudaGetLastError());
    uint32_t tot[2] = {0, 0};
    CK(cudaMemcpyAsync(tot, r.sched_totals, sizeof tot, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // (also: `anchors` may be a stack array of the caller)
    r.sched_n_unload = tot[0];
    r.sched_n_load = tot[1];
    if (n_unload) *n_unload = tot[0];
    if (n_load) *n_load = tot[1];
    return BXW_OK;
}

int bxw_region_load_deltas(bxw_ctx *ctx, bxw_chunk_delta *out, uint32_t cap, uint32_t *n) {
    if (!ctx || !ctx->region.live || !ctx->region.sched_deltas) return fail(ctx, BXW_E_INVALID, "no delta list (bxw_region_update_anchors first)");
    Region &r = ctx->region;
    const uint32_t total = r.sched_n_unload + r.sched_n_load, m = std::min(total, cap);
    if (m && out) CK(cudaMemcpyAsync(out, r.sched_deltas, (size_t)m * sizeof(bxw_chunk_delta), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n) *n = total;
    return BXW_OK;
}

int bxw_region_load_state(bxw_ctx *ctx, uint8_t *blocks_loaded, uint8_t *mesh_loaded) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    if (blocks_loaded) CK(cudaMemcpyAsync(blocks_loaded, r.blk_loaded, r.n_storage(), cudaMemcpyDeviceToHost, ctx->stream));
    if (mesh_loaded) CK(cudaMemcpyAsync(mesh_loaded, r.mesh_loaded, r.n_work, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_region_apply_deltas(bxw_ctx *ctx, uint64_t seed, int precision, uint32_t max_deltas, uint32_t stats[4]) {
    if (!ctx || !ctx->region.live || !ctx->region.sched_deltas) return fail(ctx, BXW_E_INVALID, "no delta list (bxw_region_update_anchors first)");
    if (!ctx->have_gen_ids) return fail(ctx, BXW_E_NO_REGISTRY, "bxw_set_gen_ids not called");
    Region &r = ctx->region;
    const uint32_t total = r.sched_n_unload + r.sched_n_load;
    const uint32_t limit = max_deltas ? std::min(max_deltas, total) : total;
    SchedParams sp = sched_params(r, 0);
    launch_delta_lists(sp, limit, ctx->stream);
    ctx->launches += 1;
    uint32_t cnt[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(cnt, r.sched_counts, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // the list is consumed: a second apply without a new update must not replay it
    r.sched_n_unload = r.sched_n_load = 0;
    if (stats) {
        stats[0] = cnt[3];
        stats[1] = cnt[0];
        stats[2] = cnt[1];
        stats[3] = cnt[2];
    }
    if (cnt[2]) { // unloaded meshes give their arena space back
        launch_unmesh(r.sched_unmesh, r.sched_counts + 2, cnt[2], r.spans, r.span_cap, r.counters, ctx->stream);
        ctx->launches += 1;
    }
    if (cnt[0]) { // block loads: the generator for just these chunks (the reference's granularity, generation.rs:96-148)
        GenParams gp;
        int rc = gen_setup(ctx, r, seed, precision, gp);
        if (rc) return rc;
        if (ctx->elide_uniform) {
            rc = ensure_shared_pages(ctx, r);
            if (rc) return rc;
        }
        FillParams fp;
        fill_params(ctx, r, ctx->elide_uniform, fp);
        fp.list = r.sched_gen;
        fp.n_list = cnt[0];
        for (int i = 0; i < 5; i++) r.fill_ids[i] = ctx->gen_ids[i];
        launch_gen_fill_kernel(fp, gp, precision, ctx->stream);
        launch_fill_xface_kernel(fp, ctx->stream);
        ctx->launches += 2;
        CK(cudaGetLastError());
        r.gen_seed = seed;
        r.gen_precision = precision;
    }
    if (cnt[1]) { // mesh loads, nearest first
        int rc = region_mesh(ctx, r, r.sched_mesh, r.sched_mesh_span, cnt[1], true);
        if (rc) return rc;
    } else if (cnt[2]) {
        CK(cudaMemcpyAsync(r.h_counters, r.counters, sizeof(MeshCounters), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if ((cnt[1] || cnt[2]) && r.varena) { // the arena space of unloaded meshes is reclaimed once it is a quarter of the arena
        const unsigned long long gv = r.h_counters->v_garbage, gi = r.h_counters->i_garbage;
        if (gv * 4ull > r.h_counters->v_alloc + 65536ull || gi * 4ull > r.h_counters->i_alloc + 98304ull) return bxw_region_compact_arena(ctx, nullptr, nullptr);
    }
    return BXW_OK;
}

int bxw_region_mesh_spans(bxw_ctx *ctx, bxw_mesh_span *spans) {
    if (!ctx || !ctx->region.live || !spans) return fail(ctx, BXW_E_INVALID, "no region / null");
    Region &r = ctx->region;
    CK(cudaMemcpyAsync(spans, r.spans, r.n_work * sizeof(bxw_mesh_span), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_region_mesh_fetch(bxw_ctx *ctx, bxw_vertex *vertices, uint64_t n_vertices, uint32_t *indices, uint64_t n_indices) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    if (n_vertices > r.v_cap || n_indices > r.i_cap) return fail(ctx, BXW_E_INVALID, "fetch exceeds arena");
    if (n_vertices && vertices) CK(cudaMemcpyAsync(vertices, r.varena, n_vertices * sizeof(bxw_vertex), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_indices && indices) CK(cudaMemcpyAsync(indices, r.iarena, n_indices * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_region_mesh_fetch_range(bxw_ctx *ctx, uint64_t v_off, uint64_t n_vertices, bxw_vertex *vertices, uint64_t i_off, uint64_t n_indices,
                                uint32_t *indices) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    if (v_off + n_vertices > r.v_cap || i_off + n_indices > r.i_cap) return fail(ctx, BXW_E_INVALID, "fetch exceeds arena");
    if (n_vertices && vertices)
        CK(cudaMemcpyAsync(vertices, r.varena + v_off, n_vertices * sizeof(bxw_vertex), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_indices && indices) CK(cudaMemcpyAsync(indices, r.iarena + i_off, n_indices * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_region_mesh_fetch_async(bxw_ctx *ctx, bxw_vertex *vertices, uint64_t n_vertices, uint32_t *indices, uint64_t n_indices) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    if (n_vertices > r.v_cap || n_indices > r.i_cap) return fail(ctx, BXW_E_INVALID, "fetch exceeds arena");
    CK(cudaEventRecord(ctx->ev_mesh_done, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_mesh_done, 0));
    if (n_vertices && vertices)
        CK(cudaMemcpyAsync(vertices, r.varena, n_vertices * sizeof(bxw_vertex), cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (n_indices && indices) CK(cudaMemcpyAsync(indices, r.iarena, n_indices * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_copy_done, ctx->copy_stream));
    ctx->copy_pending = true;
    return BXW_OK;
}

int bxw_region_fetch_wait(bxw_ctx *ctx) {
    if (!ctx) return BXW_E_INVALID;
    if (ctx->copy_pending) {
        CK(cudaEventSynchronize(ctx->ev_copy_done));
        ctx->copy_pending = false;
    }
    return BXW_OK;
}

int bxw_region_apply_changes(bxw_ctx *ctx, const bxw_voxel_change *changes, uint32_t n, uint32_t *n_dirty) {
    if (!ctx || !ctx->region.live || (n && !changes)) return fail(ctx, BXW_E_INVALID, "no region / null");
    Region &r = ctx->region;
    const int32_t x_min = (r.cx0 - 1) * 32, y_min = (r.cy0 - 1) * 32, z_min = (r.cz0 - 1) * 32;
    // Keep the changes whose chunk is held by this region (the reference skips chunks that are not loaded,
    // worldmgr.rs:324-326) and group them by voxel, preserving the caller's order inside a group: the reference
    // replays a chunk's changes sequentially (generation.rs:52-57).
    struct Keyed {
        uint64_t key;
        uint32_t idx;
    };
    std::vector<Keyed> keyed;
    keyed.reserve(n);
    for (uint32_t i = 0; i < n && r.shape_hint == 0; i++)
        if ((changes[i].to & 63u) - 1u < 3u) r.shape_hint = 1; // a slope / corner arrives (StdMeta: shape = meta % 64)
    for (uint32_t i = 0; i < n; i++) {
        const int64_t gx = (int64_t)changes[i].x - x_min, gy = (int64_t)changes[i].y - y_min, gz = (int64_t)changes[i].z - z_min;
        if (gx < 0 || gy < 0 || gz < 0 || gx >= (int64_t)r.SX * 32 || gy >= (int64_t)r.SY * 32 || gz >= (int64_t)r.SZ * 32) continue;
        keyed.push_back({((uint64_t)gy * (uint64_t)(r.SZ * 32) + (uint64_t)gz) * (uint64_t)(r.SX * 32) + (uint64_t)gx, i});
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](const Keyed &a, const Keyed &b) { return a.key < b.key; });
    std::vector<bxw_voxel_change> sorted(keyed.size());
    std::vector<uint32_t> runs;
    for (size_t i = 0; i < keyed.size(); i++) {
        sorted[i] = changes[keyed[i].idx];
        if (i == 0 || keyed[i].key != keyed[i - 1].key) runs.push_back((uint32_t)i);
    }
    const uint32_t n_runs = (uint32_t)runs.size();
    runs.push_back((uint32_t)keyed.size());
    std::vector<uint32_t> touched; // distinct storage chunks that receive a change
    if (n_runs) {
        for (uint32_t k = 0; k < n_runs; k++) {
            const bxw_voxel_change &c = sorted[runs[k]];
            const int sx = (c.x - x_min) >> 5, sy = (c.y - y_min) >> 5, sz = (c.z - z_min) >> 5;
            touched.push_back((uint32_t)(sx + r.SX * (sz + r.SZ * sy)));
        }
        std::sort(touched.begin(), touched.end());
        touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
        int rc = ensure_io(ctx, sorted.size() * sizeof(bxw_voxel_change), (runs.size() + touched.size()) * sizeof(uint32_t));
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->io_a, sorted.data(), sorted.size() * sizeof(bxw_voxel_change), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->io_b, runs.data(), runs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        // single-voxel writes need the chunk's own storage: chunks still on a shared uniform page get a copy of it first
        uint32_t *d_touched = (uint32_t *)ctx->io_b + runs.size();
        CK(cudaMemcpyAsync(d_touched, touched.data(), touched.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        launch_materialise(r.blocks, r.xface, r.page, d_touched, (uint32_t)touched.size(), ctx->stream);
        ctx->launches += 1;
        EditParams ep;
        ep.blocks = r.blocks;
        ep.xface = r.xface;
        ep.summary = r.summary;
        ep.SX = r.SX;
        ep.SY = r.SY;
        ep.SZ = r.SZ;
        ep.x_min = x_min;
        ep.y_min = y_min;
        ep.z_min = z_min;
        ep.changes = (const bxw_voxel_change *)ctx->io_a;
        ep.run_start = (const uint32_t *)ctx->io_b;
        ep.n_runs = n_runs;
        ep.dirty = r.dirty;
        ep.blk_loaded = r.streaming ? r.blk_loaded : nullptr;
        launch_apply_changes(ep, ctx->stream);
        ctx->launches += 1;
    }
    CK(cudaMemsetAsync(r.dirty_count, 0, sizeof(uint32_t), ctx->stream));
    launch_dirty_list(r.dirty, nullptr, r.n_work, (int)r.nx, (int)r.nz, r.SX, r.SZ, r.dirty_work, r.dirty_span, r.dirty_count, 0, ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
    uint32_t nd = 0;
    CK(cudaMemcpyAsync(&nd, r.dirty_count, sizeof nd, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // also keeps `sorted` / `runs` alive until the copies are done
    if (n_dirty) *n_dirty = nd;
    return BXW_OK;
}

int bxw_region_remesh_dirty(bxw_ctx *ctx, uint32_t *n_remeshed) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    CK(cudaMemsetAsync(r.dirty_count, 0, sizeof(uint32_t), ctx->stream));
    // (a region that was never meshed as a whole and has no load anchors keeps the old behaviour: every dirty chunk is meshed)
    launch_dirty_list(r.dirty, r.streaming ? r.mesh_loaded : nullptr, r.n_work, (int)r.nx, (int)r.nz, r.SX, r.SZ, r.dirty_work, r.dirty_span,
                      r.dirty_count, 1, ctx->stream);
    ctx->launches += 1;
    uint32_t nd = 0;
    CK(cudaMemcpyAsync(&nd, r.dirty_count, sizeof nd, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_remeshed) *n_remeshed = nd;
    if (nd == 0) return BXW_OK;
    // A chunk whose new mesh fits its old arena reservation is rewritten in place; the others move to the end of the arena and
    // their old reservation becomes garbage (the reference drops the old ChunkBuffers when the new one lands,
    // voxrender.rs:398-413). Once a quarter of the used arena is garbage the live spans are compacted.
    int rc = region_mesh(ctx, r, r.dirty_work, r.dirty_span, nd, true);
    if (rc) return rc;
    const unsigned long long gv = r.h_counters->v_garbage, gi = r.h_counters->i_garbage;
    if (gv * 4ull > r.h_counters->v_alloc + 65536ull || gi * 4ull > r.h_counters->i_alloc + 98304ull) return bxw_region_compact_arena(ctx, nullptr, nullptr);
    return BXW_OK;
}

int bxw_region_arena_extent(bxw_ctx *ctx, uint64_t *arena_vertices, uint64_t *arena_indices) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    CK(cudaMemcpyAsync(r.h_counters, r.counters, sizeof(MeshCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (arena_vertices) *arena_vertices = r.h_counters->v_alloc;
    if (arena_indices) *arena_indices = r.h_counters->i_alloc;
    return BXW_OK;
}

int bxw_region_arena_stats(bxw_ctx *ctx, bxw_arena_stats *out) {
    if (!ctx || !ctx->region.live || !out) return fail(ctx, BXW_E_INVALID, "no region / null");
    Region &r = ctx->region;
    CK(cudaMemcpyAsync(r.h_counters, r.counters, sizeof(MeshCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out->vertices_used = r.h_counters->v_alloc;
    out->indices_used = r.h_counters->i_alloc;
    out->vertices_garbage = r.h_counters->v_garbage;
    out->indices_garbage = r.h_counters->i_garbage;
    out->vertices_capacity = r.v_cap;
    out->indices_capacity = r.i_cap;
    return BXW_OK;
}

int bxw_region_compact_arena(bxw_ctx *ctx, uint64_t *arena_vertices, uint64_t *arena_indices) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    if (!r.varena) return fail(ctx, BXW_E_INVALID, "no mesh arena yet");
    if (ctx->copy_pending) { // an asynchronous fetch may still read the arena that is about to be replaced
        CK(cudaEventSynchronize(ctx->ev_copy_done));
        ctx->copy_pending = false;
    }
    int rc = ensure_io_c(ctx, (size_t)r.n_work * 2 * sizeof(unsigned long long));
    if (rc) return rc;
    unsigned long long *new_off = (unsigned long long *)ctx->io_c;
    // the second arena is allocated once and the two swap roles (a cudaMalloc / cudaFree pair of gigabytes per compaction was
    // the 5 ms latency spike of the edit sweep)
    if (!r.varena2) {
        CK(cudaMalloc(&r.varena2, r.v_cap * sizeof(bxw_vertex)));
        if (cudaMalloc(&r.iarena2, r.i_cap * sizeof(uint32_t)) != cudaSuccess) {
            cudaFree(r.varena2);
            r.varena2 = nullptr;
            return fail(ctx, BXW_E_NOMEM, "no device memory for the second arena");
        }
    }
    launch_compact_offsets(r.spans, r.span_cap, r.n_work, new_off, r.counters, ctx->stream);
    launch_compact_copy(r.spans, new_off, r.n_work, r.varena, r.iarena, r.varena2, r.iarena2, ctx->stream);
    ctx->launches += 2;
    CK(cudaMemcpyAsync(r.h_counters, r.counters, sizeof(MeshCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    std::swap(r.varena, r.varena2);
    std::swap(r.iarena, r.iarena2);
    if (arena_vertices) *arena_vertices = r.h_counters->v_alloc;
    if (arena_indices) *arena_indices = r.h_counters->i_alloc;
    return BXW_OK;
}

int bxw_region_set_origin(bxw_ctx *ctx, int32_t cx0, int32_t cy0, int32_t cz0) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    CK(cudaStreamSynchronize(ctx->stream));
    r.cx0 = cx0;
    r.cy0 = cy0;
    r.cz0 = cz0;
    r.have_hmap = false; // the stored voxels belong to the old position until the next generate / upload
    return BXW_OK;
}

int bxw_region_page_table(bxw_ctx *ctx, void **pages, uint32_t *n_shared) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    if (pages) *pages = ctx->region.page;
    if (n_shared) *n_shared = kSharedPages;
    return BXW_OK;
}

int bxw_region_download_pages(bxw_ctx *ctx, uint32_t *pages) {
    if (!ctx || !ctx->region.live || !pages) return fail(ctx, BXW_E_INVALID, "no region / null");
    Region &r = ctx->region;
    CK(cudaMemcpyAsync(pages, r.page, r.n_storage() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

size_t bxw_region_halo_bytes(const bxw_ctx *ctx) {
    if (!ctx || !ctx->region.live) return 0;
    return (size_t)ctx->region.SZ * 32 * ctx->region.SY * 32 * sizeof(uint32_t);
}
int bxw_region_halo_export(bxw_ctx *ctx, int face, void *dev_buf) {
    if (!ctx || !ctx->region.live || !dev_buf || (face != 0 && face != 1)) return fail(ctx, BXW_E_INVALID, "bad halo export");
    Region &r = ctx->region;
    // the interior's own boundary plane: x- face = storage chunk column 1, local x 0; x+ face = column nx, local x 31
    launch_halo_plane(r.blocks, r.xface, nullptr, r.page, r.need, r.SX, r.SY, r.SZ, face == 0 ? 1 : (int)r.nx, face == 0 ? 0 : 31, (uint32_t *)dev_buf, 0, ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
    return BXW_OK;
}

int bxw_region_halo_import(bxw_ctx *ctx, int face, const void *dev_buf) {
    if (!ctx || !ctx->region.live || !dev_buf || (face != 0 && face != 1)) return fail(ctx, BXW_E_INVALID, "bad halo import");
    Region &r = ctx->region;
    // into the shell: x- shell = column 0, local x 31; x+ shell = column nx+1, local x 0
    launch_halo_plane(r.blocks, r.xface, r.summary, r.page, r.need, r.SX, r.SY, r.SZ, face == 0 ? 0 : (int)r.nx + 1, face == 0 ? 31 : 0, (uint32_t *)dev_buf, 1, ctx->stream);
    ctx->launches += 3;
    CK(cudaGetLastError());
    return BXW_OK;
}

int bxw_region_device_ptrs(bxw_ctx *ctx, void **blocks, void **vertices, void **indices) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    if (blocks) *blocks = ctx->region.blocks;
    if (vertices) *vertices = ctx->region.varena;
    if (indices) *indices = ctx->region.iarena;
    return BXW_OK;
}

int bxw_terragen_noise2(bxw_ctx *ctx, uint64_t seed, const double *xs, const double *ys, size_t n, double *out) {
    if (!ctx || !xs || !ys || !out) return fail(ctx, BXW_E_INVALID, "null argument");
    if (n == 0) return BXW_OK;
    if (!ctx->tg_valid || ctx->tg_seed != seed) {
        TerragenTables *t = new (std::nothrow) TerragenTables();
        if (!t) return fail(ctx, BXW_E_NOMEM, "host alloc");
        build_terragen_tables(seed, *t);
        if (!ctx->d_tg && cudaMalloc(&ctx->d_tg, sizeof(TerragenTables)) != cudaSuccess) {
            delete t;
            return fail(ctx, BXW_E_CUDA, "cudaMalloc terragen tables");
        }
        cudaError_t e = cudaMemcpy(ctx->d_tg, t, sizeof(TerragenTables), cudaMemcpyHostToDevice);
        delete t;
        if (e != cudaSuccess) return fail(ctx, BXW_E_CUDA, "upload terragen tables");
        ctx->tg_valid = true;
        ctx->tg_seed = seed;
    }
    int rc = ensure_io(ctx, 2 * n * sizeof(double), n * sizeof(double));
    if (rc) return rc;
    double *dx = (double *)ctx->io_a, *dy = dx + n, *dout = (double *)ctx->io_b;
    CK(cudaMemcpyAsync(dx, xs, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dy, ys, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    launch_terragen_noise2(ctx->d_tg, dx, dy, n, dout, ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_stdgen_noise(bxw_ctx *ctx, uint64_t seed, int which, int precision, const double *xs, const double *ys, size_t n, double *out) {
    if (!ctx || !xs || !ys || !out) return fail(ctx, BXW_E_INVALID, "null argument");
    if (which < 0 || which > 6) return fail(ctx, BXW_E_INVALID, "noise function 0..6");
    if (precision != 32 && precision != 64) return fail(ctx, BXW_E_INVALID, "precision must be 32 or 64");
    if (n == 0) return BXW_OK;
    int rc = ensure_gen_tables(ctx, seed);
    if (rc) return rc;
    rc = ensure_io(ctx, 2 * n * sizeof(double), n * sizeof(double));
    if (rc) return rc;
    double *dx = (double *)ctx->io_a, *dy = dx + n, *dout = (double *)ctx->io_b;
    CK(cudaMemcpyAsync(dx, xs, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dy, ys, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    launch_noise_samples(ctx->d_gen, which, dx, dy, n, dout, precision, ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BXW_OK;
}

int bxw_region_fill_synthetic(bxw_ctx *ctx, uint64_t seed, uint32_t void_threshold) {
    if (!ctx || !ctx->region.live) return fail(ctx, BXW_E_INVALID, "no region");
    Region &r = ctx->region;
    CK(cudaMemsetAsync(r.summary, 0, r.n_storage() * sizeof(unsigned long long), ctx->stream));
    launch_page_identity(r.page, (uint32_t)r.n_storage(), ctx->stream);
    r.shape_hint = 2;
    CK(cudaMemsetAsync(r.blk_loaded, 1, r.n_storage(), ctx->stream));
    launch_fill_synthetic(r.blocks, r.SX, r.SY, r.SZ, r.cx0 - 1, r.cy0 - 1, r.cz0 - 1, seed, void_threshold, ctx->stream);
    launch_xface_extract(r.blocks, r.xface, nullptr, 0, (uint32_t)r.n_storage(), ctx->stream);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return BXW_OK;
}

int bxw_timer_start(bxw_ctx *ctx) {
    if (!ctx) return BXW_E_INVALID;
    CK(cudaEventRecord(ctx->ev_a, ctx->stream));
    return BXW_OK;
}

int bxw_timer_stop_ms(bxw_ctx *ctx, float *ms) {
    if (!ctx || !ms) return BXW_E_INVALID;
    CK(cudaEventRecord(ctx->ev_b, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev_b));
    CK(cudaEventElapsedTime(ms, ctx->ev_a, ctx->ev_b));
    return BXW_OK;
}

int bxw_mesh_kernel_time_ms(bxw_ctx *ctx, float *ms, uint64_t *launches, int reset) {
    if (!ctx) return BXW_E_INVALID;
    if (ms) *ms = (float)ctx->mesh_ms;
    if (launches) *launches = ctx->mesh_launches;
    if (reset) {
        ctx->mesh_ms = 0.0;
        ctx->emit_ms = 0.0;
        ctx->mesh_launches = 0;
    }
    return BXW_OK;
}

int bxw_kernel_time_ms(bxw_ctx *ctx, int which, float *ms, uint64_t *launches, int reset) {
    if (!ctx) return BXW_E_INVALID;
    int rc = flush_gen_timing(ctx);
    if (rc) return rc;
    double *acc;
    uint64_t *cnt;
    switch (which) {
    case BXW_KERNEL_MESH:
        acc = &ctx->mesh_ms;
        cnt = &ctx->mesh_launches;
        if (reset) ctx->emit_ms = 0.0;
        break;
    case BXW_KERNEL_FILL: acc = &ctx->fill_ms; cnt = &ctx->fill_launches; break;
    case BXW_KERNEL_HEIGHTMAP: acc = &ctx->hmap_ms; cnt = &ctx->hmap_launches; break;
    case BXW_KERNEL_EMIT: // the emit_kernel part of BXW_KERNEL_MESH (same launches; reset together with it)
        if (ms) *ms = (float)ctx->emit_ms;
        if (launches) *launches = ctx->mesh_launches;
        if (reset) ctx->emit_ms = 0.0;
        return BXW_OK;
    default: return fail(ctx, BXW_E_INVALID, "unknown kernel selector %d", which);
    }
    if (ms) *ms = (float)*acc;
    if (launches) *launches = *cnt;
    if (reset) {
        *acc = 0.0;
        *cnt = 0;
    }
    return BXW_OK;
}

int bxw_debug_phase_cycles(bxw_ctx *ctx, uint64_t out[12]) {
    if (!ctx || !ctx->region.live || !out) return BXW_E_INVALID;
    for (int k = 0; k < 12; k++) out[k] = ctx->region.h_counters->phase[k];
    return BXW_OK;
}

int bxw_region_mesh_candidates(bxw_ctx *ctx, uint32_t *listed, uint32_t *interior) {
    if (!ctx) return BXW_E_INVALID;
    if (listed) *listed = ctx->last_candidates;
    if (interior) *interior = ctx->last_candidates_of;
    return BXW_OK;
}

} // extern "C"
