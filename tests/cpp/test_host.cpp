// test_host.cpp — parity tests of the C++ host mirror (ballxworld_b200/host/bxw_host.hpp) through the C ABI, written the
// way the reference writes its own tests (bxw_world/src/lib.rs:481-546) plus oracle comparisons for the generator, the
// mesher, the batched region handlers and edits. Test infrastructure: links oracle/libbxw_oracle.so as the checker.
//
// Build + run: tests/test_cpp_host.py (g++ -std=c++17, -lbxw_cuda -lbxw_oracle). Exit code 0 = all passed.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <tuple>

#include "bxw_host.hpp"

extern "C" {
#include "bxw_oracle.h"
}

using namespace bxw;

static int g_failed = 0, g_run = 0;
#define CHECK(cond)                                                                 \
    do {                                                                            \
        if (!(cond)) {                                                              \
            std::printf("    FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
            g_failed++;                                                             \
            return;                                                                 \
        }                                                                           \
    } while (0)
#define RUN(fn)                                  \
    do {                                         \
        const int before = g_failed;             \
        g_run++;                                 \
        fn();                                    \
        std::printf("test %s ... %s\n", #fn, g_failed == before ? "ok" : "FAILED"); \
    } while (0)

static Gpu *g_gpu = nullptr;
static VoxelRegistry *g_reg = nullptr;

static bxo_registry oracle_registry(std::vector<bxo_block_def> &store) {
    store.resize(8);
    bxo_standard_registry(store.data());
    return bxo_registry{8, store.data()};
}

// ---- bxw_world/src/lib.rs:486-546, through VChunk::compress / decompress on the device ----
static void rle_compress_zero_test() {
    UncompressedChunk zeroes;
    VChunk c;
    c.compress(*g_gpu, zeroes);
    CHECK((c.vox == std::vector<uint32_t>{0, 0, (uint32_t)CHUNK_DIM3 - 2}));
}

static void rle_compress_one_test() {
    UncompressedChunk ones;
    ones.blocks_yzx.assign(CHUNK_DIM3, 1u);
    VChunk c;
    c.compress(*g_gpu, ones);
    CHECK((c.vox == std::vector<uint32_t>{1, 1, (uint32_t)CHUNK_DIM3 - 2}));
}

static void rle_decompress_zero_test() {
    VChunk c;
    c.vox = {0, 0, (uint32_t)CHUNK_DIM3 - 2};
    auto target = c.decompress(*g_gpu);
    for (uint32_t v : target->blocks_yzx) CHECK(v == 0);
}

static void rle_decompress_one_test() {
    VChunk c;
    c.vox = {1, 1, (uint32_t)CHUNK_DIM3 - 2};
    auto target = c.decompress(*g_gpu);
    for (uint32_t v : target->blocks_yzx) CHECK(v == 1);
}

static void rle_random_cmp() {
    UncompressedChunk rnd;
    bxo_xoshiro256ss rng;
    bxo_xoshiro256ss_seed_from_u64(&rng, 1234);
    for (auto &e : rnd.blocks_yzx) e = bxo_xoshiro256ss_next_u32(&rng) % 16;
    VChunk c;
    c.compress(*g_gpu, rnd);
    std::vector<uint32_t> want(2 * CHUNK_DIM3 + 4);
    want.resize(bxo_compress_rle(rnd.blocks_yzx.data(), CHUNK_DIM3, want.data()));
    CHECK(c.vox == want);
    auto dec = c.decompress(*g_gpu);
    CHECK(dec->blocks_yzx == rnd.blocks_yzx);
    const VChunk back = VChunk::deserialize(c.position, c.serialize());
    CHECK(back.vox == c.vox);
}

static void rle_decompress_error() { // lib.rs:296-302: a stream that does not fill the chunk
    VChunk c;
    c.vox = {0, 0, 5, 1, 1};
    bool threw = false;
    try {
        c.decompress(*g_gpu);
    } catch (const Error &e) {
        threw = e.code == BXW_E_BAD_RLE;
    }
    CHECK(threw);
}

// ---- StdGenerator::generate_chunk (stdgen.rs:304-374) ----
static void generate_chunk_matches_oracle() {
    for (uint64_t seed : {0ull, 0x13372137CAFEull}) {
        StdGenerator gen(*g_gpu, seed);
        bxo_stdgen *og = bxo_stdgen_new(seed);
        const auto ids = StdGenerator::resolve_ids(*g_reg);
        const uint16_t oids[5] = {ids[0], ids[1], ids[2], ids[3], ids[4]};
        for (ChunkPosition p : {ChunkPosition{9, 0, 0}, ChunkPosition{9, 1, 0}, ChunkPosition{-3, 0, 7}, ChunkPosition{40, 2, -40}, ChunkPosition{0, -1, 0}}) {
            UncompressedChunk c;
            c.position = p;
            gen.generate_chunk(c, *g_reg);
            std::vector<uint32_t> want(CHUNK_DIM3);
            bxo_generate_chunk(og, p.x, p.y, p.z, oids, want.data());
            CHECK(c.blocks_yzx == want);
        }
        bxo_stdgen_free(og);
    }
}

static void generate_chunk_needs_the_standard_names() { // stdgen.rs:307-326 .expect(...)
    VoxelRegistry empty;
    StdGenerator gen(*g_gpu, 0);
    UncompressedChunk c;
    bool threw = false;
    try {
        gen.generate_chunk(c, empty);
    } catch (const std::runtime_error &e) {
        threw = std::string(e.what()) == "No standard grass block definition found";
    }
    CHECK(threw);
}

// ---- mesh_from_chunk / is_chunk_trivial (voxmesh.rs:16-190) ----
static void mesh_from_chunk_matches_oracle() {
    std::vector<bxo_block_def> store;
    const bxo_registry oreg = oracle_registry(store);
    StdGenerator gen(*g_gpu, 0);
    const ChunkPosition centre{9, 0, 0};
    std::vector<VChunk> vchunks(27);
    std::vector<const VChunk *> ptrs(27);
    std::vector<std::vector<uint32_t>> dense(27);
    int k = 0;
    for (ChunkPosition p : iter_neighbors(centre)) {
        UncompressedChunk c;
        c.position = p;
        gen.generate_chunk(c, *g_reg);
        if (k == 13) { // some oriented slopes / corners on the surface, so the general path runs too
            for (int i = 0; i < CHUNK_DIM3; i += 97)
                if (c.blocks_yzx[i]) c.blocks_yzx[i] = VoxelDatum(5, StdMeta::from_parts(1 + i % 3, i % 24)).repr;
        }
        dense[k] = c.blocks_yzx;
        vchunks[k].position = p;
        vchunks[k].compress(*g_gpu, c);
        ptrs[k] = &vchunks[k];
        k++;
    }
    CHECK(!is_chunk_trivial(*g_gpu, vchunks[13], *g_reg));
    VChunk sky;
    sky.position = {9, 30, 0};
    CHECK(is_chunk_trivial(*g_gpu, sky, *g_reg));
    auto got = mesh_from_chunk(*g_gpu, *g_reg, ptrs, {16, 16});
    CHECK(got.has_value());
    const uint32_t *dp[27];
    for (int i = 0; i < 27; i++) dp[i] = dense[i].data();
    bxo_mesh want{};
    CHECK(bxo_mesh_from_dense27(&oreg, dp, &want) == 0);
    CHECK(want.n_vertices > 1000);
    CHECK(got->vertices.size() == want.n_vertices && got->indices.size() == want.n_indices);
    CHECK(std::memcmp(got->vertices.data(), want.vertices, want.n_vertices * sizeof(bxw_vertex)) == 0);
    CHECK(std::memcmp(got->indices.data(), want.indices, want.n_indices * sizeof(uint32_t)) == 0);
    bxo_mesh_free(&want);
}

static void mesh_from_chunk_panics_on_unknown_ids() { // voxregistry.rs:141-143
    std::vector<VChunk> vchunks(27);
    std::vector<const VChunk *> ptrs(27);
    for (int i = 0; i < 27; i++) {
        vchunks[i].vox = {0, 0, (uint32_t)CHUNK_DIM3 - 2}; // all void (the Default placeholder [0, 0, CHUNK_DIM3] is not a decodable stream)
        ptrs[i] = &vchunks[i];
    }
    vchunks[13].vox = {(uint32_t)200 << 16, (uint32_t)200 << 16, (uint32_t)CHUNK_DIM3 - 2};
    bool threw = false;
    try {
        mesh_from_chunk(*g_gpu, *g_reg, ptrs, {16, 16});
    } catch (const Error &e) {
        threw = e.code == BXW_E_UNKNOWN_ID;
    }
    CHECK(threw);
}

// ---- the batched handlers: BASELINE configs[0] totals and a chunk-by-chunk oracle comparison ----
static void region_generate_and_mesh_cfg1() {
    Region region(*g_gpu, {0, 0, 0}, {16, 8, 16}, *g_reg);
    StdGenerator gen(*g_gpu, 0);
    const auto arena = region.generate_and_mesh(gen);
    const auto spans = region.spans();
    uint64_t tv = 0, ti = 0;
    for (const auto &s : spans) {
        tv += s.v_count;
        ti += s.i_count;
    }
    CHECK(tv == 1455136 && ti == 2182704);
    const ChunkBuffers all = region.fetch_arena(arena.first, arena.second);
    // one surface chunk against the oracle, neighbours taken back from device storage as VChunks
    std::vector<bxo_block_def> store;
    const bxo_registry oreg = oracle_registry(store);
    size_t best = 0;
    for (size_t i = 0; i < spans.size(); i++)
        if (spans[i].v_count > spans[best].v_count) best = i;
    const int32_t ix = (int32_t)(best % 16), iz = (int32_t)((best / 16) % 16), iy = (int32_t)(best / 256);
    std::vector<std::unique_ptr<UncompressedChunk>> dense;
    const uint32_t *dp[27];
    int k = 0;
    for (int ry = -1; ry <= 1; ry++)
        for (int rz = -1; rz <= 1; rz++)
            for (int rx = -1; rx <= 1; rx++) {
                dense.push_back(region.get_data(ix + rx, iy + ry, iz + rz).decompress(*g_gpu));
                dp[k++] = dense.back()->blocks_yzx.data();
            }
    bxo_mesh want{};
    CHECK(bxo_mesh_from_dense27(&oreg, dp, &want) == 0);
    const auto &s = spans[best];
    CHECK(s.v_count == want.n_vertices && s.i_count == want.n_indices);
    CHECK(std::memcmp(all.vertices.data() + s.v_off, want.vertices, want.n_vertices * sizeof(bxw_vertex)) == 0);
    CHECK(std::memcmp(all.indices.data() + s.i_off, want.indices, want.n_indices * sizeof(uint32_t)) == 0);
    bxo_mesh_free(&want);
}

// ---- World::apply_voxel_changes (worldmgr.rs:312-357): ordered CAS, touching_chunks dirty, re-mesh ----
static void apply_voxel_changes_and_remesh() {
    Region region(*g_gpu, {9, 0, 0}, {2, 2, 2}, *g_reg);
    StdGenerator gen(*g_gpu, 0);
    region.generate(gen);
    region.mesh();
    bxo_stdgen *og = bxo_stdgen_new(0);
    bxo_voxel_params vp;
    const int32_t x = 9 * 32 + 32, z = 20; // x on the boundary between the two chunk columns
    bxo_calc_voxel_params(og, x, z, &vp);
    bxo_stdgen_free(og);
    const BlockPosition top{x, vp.height, z};
    const VoxelDatum stone(4, 0), slope(6, StdMeta::from_parts(1, 5));
    const VoxelDatum cur = region.get_data(1, top.chunk().y, 0).decompress(*g_gpu)->get(top);
    CHECK(cur.id() != 0);
    const std::vector<VoxelChange> changes = {{top, cur, slope}, {top, cur, stone} /* stale from: must not apply */,
                                              {top, slope, stone}, {{x, vp.height, z}, VoxelDatum(99, 0), VoxelDatum()} /* wrong from */};
    const size_t expect_dirty = top.touching_chunks().size();
    CHECK(expect_dirty == 2); // local x == 0: own chunk and the -x neighbour, both inside the box
    CHECK(region.apply_voxel_changes(changes) == expect_dirty);
    CHECK(region.remesh_dirty() == expect_dirty);
    CHECK(region.get_data(1, top.chunk().y, 0).decompress(*g_gpu)->get(top) == stone);
    CHECK(region.remesh_dirty() == 0);
}

// ---- load anchors: recalculate_load_deltas + main_loop_tick (worldmgr.rs:407-732) through Region::tick ----
static void load_anchor_ticks() {
    Region region(*g_gpu, {6, 0, -2}, {10, 3, 9}, *g_reg);
    region.unload_all();
    const std::vector<bxw_load_anchor> anchors = {{9, 1, 2, 3, 1}};
    const auto t1 = region.tick(anchors, 0);          // nothing loaded: block data of the sphere (clipped to the box)
    CHECK(t1[1] > 0 && t1[2] == 0 && t1[0] == t1[1]);
    const auto t2 = region.tick(anchors, 0);          // now the meshes whose 27 chunks hold block data
    CHECK(t2[1] == 0 && t2[2] > 0);
    const auto t3 = region.tick(anchors, 0);
    CHECK(t3[0] == 0);                                // steady state
    // the anchor moves: the list interleaves unloads and loads, each nearest first
    uint32_t nu = 0, nl = 0;
    const std::vector<bxw_load_anchor> moved = {{12, 1, 2, 3, 1}};
    g_gpu->check(bxw_region_update_anchors(g_gpu->raw(), moved.data(), 1, &nu, &nl));
    CHECK(nu > 0 && nl > 0);
    const auto deltas = region.load_deltas();
    CHECK(deltas.size() == nu + nl);
    int32_t last_u = -1, last_l = -1;
    bool sorted = true;
    for (size_t i = 0; i < deltas.size(); i++) {
        int32_t &last = deltas[i].unload ? last_u : last_l;
        sorted &= deltas[i].min_anchor_distance >= last;
        last = deltas[i].min_anchor_distance;
        if (i < 2 * std::min(nu, nl)) sorted &= (deltas[i].unload != 0) == (i % 2 == 0);
    }
    CHECK(sorted);
}

// ---- the rest of the C ABI through the mirror: dense chunks in / out, per-chunk fetch, raw column parameters, split
// generation, sphere pass, arena bookkeeping, moving the box ----
static void region_plumbing() {
    Region region(*g_gpu, {3, 0, -4}, {4, 3, 3}, *g_reg);
    StdGenerator gen(*g_gpu, 7);
    region.generate(gen);
    // split generation (boundary columns, then the rest on the side stream) = one call
    const UncompressedChunk whole = region.download(0, 1, 1), shell = region.download(-1, 0, 3);
    region.generate_part(gen, 1);
    region.generate_part(gen, 2);
    CHECK(region.download(0, 1, 1).blocks_yzx == whole.blocks_yzx && region.download(-1, 0, 3).blocks_yzx == shell.blocks_yzx);
    // the generator's chunk = the region's chunk
    UncompressedChunk one;
    one.position = {3 + 2, 0 + 1, -4 + 1};
    gen.generate_chunk(one, *g_reg);
    CHECK(region.download(2, 1, 1).blocks_yzx == one.blocks_yzx);
    // raw column parameters against the oracle
    const auto hm = region.heightmap();
    CHECK(hm.width == 6 * 32 && hm.depth == 5 * 32);
    bxo_stdgen *og = bxo_stdgen_new(7);
    bool same = true;
    for (uint32_t k = 0; k < 200; k++) {
        const uint32_t ix = (k * 37u) % hm.width, iz = (k * 91u) % hm.depth;
        bxo_voxel_params vp;
        bxo_calc_voxel_params(og, (3 - 1) * 32 + (int32_t)ix, (-4 - 1) * 32 + (int32_t)iz, &vp);
        same &= hm.height[(size_t)iz * hm.width + ix] == vp.height;
    }
    bxo_stdgen_free(og);
    CHECK(same);
    // whole pass, then one chunk's buffers by range = the same bytes inside the arena
    const auto arena = region.mesh();
    const auto spans = region.spans();
    const ChunkBuffers all = region.fetch_arena(arena.first, arena.second);
    size_t best = 0;
    for (size_t i = 0; i < spans.size(); i++)
        if (spans[i].v_count > spans[best].v_count) best = i;
    const ChunkBuffers mine = region.fetch_chunk(spans[best]);
    CHECK(mine.vertices.size() == spans[best].v_count && spans[best].v_count > 0);
    CHECK(std::memcmp(mine.vertices.data(), all.vertices.data() + spans[best].v_off, mine.vertices.size() * sizeof(bxw_vertex)) == 0);
    CHECK(std::memcmp(mine.indices.data(), all.indices.data() + spans[best].i_off, mine.indices.size() * sizeof(uint32_t)) == 0);
    // asynchronous fetch = synchronous fetch
    ChunkBuffers again;
    again.vertices.resize(arena.first);
    again.indices.resize(arena.second);
    region.fetch_async(again.vertices.data(), arena.first, again.indices.data(), arena.second);
    region.fetch_wait();
    CHECK(std::memcmp(again.vertices.data(), all.vertices.data(), arena.first * sizeof(bxw_vertex)) == 0);
    const auto cand = region.mesh_candidates();
    CHECK(cand.second == 4 * 3 * 3 && cand.first <= cand.second && cand.first > 0);
    // arena bookkeeping: nothing is garbage after a whole pass; compaction keeps every mesh
    const bxw_arena_stats st = region.arena_stats();
    CHECK(st.vertices_used == arena.first && st.vertices_garbage == 0 && st.vertices_capacity >= st.vertices_used);
    const auto compacted = region.compact_arena();
    CHECK(compacted.first <= arena.first);
    const ChunkBuffers moved = region.fetch_chunk(region.spans()[best]);
    CHECK(moved.vertices.size() == mine.vertices.size() &&
          std::memcmp(moved.vertices.data(), mine.vertices.data(), mine.vertices.size() * sizeof(bxw_vertex)) == 0);
    // the sphere pass meshes a subset, every chunk of the box holds block data, and meshes only inside the sphere
    const auto sph = region.mesh_sphere({4, 1, -3}, 1);
    CHECK(sph[0] == 7); // centre + six neighbours, all inside the box
    const auto ls = region.load_state();
    CHECK(ls.first.size() == 6 * 5 * 5 && ls.second.size() == 4 * 3 * 3);
    // upload a dense chunk, read it back
    UncompressedChunk custom;
    for (size_t i = 0; i < custom.blocks_yzx.size(); i++) custom.blocks_yzx[i] = (i % 3 == 0) ? VoxelDatum(4, 0).repr : 0u;
    region.upload(1, 1, 1, custom);
    CHECK(region.download(1, 1, 1).blocks_yzx == custom.blocks_yzx);
    // the box moves: generation there equals the generator's chunks at the new position
    region.set_origin({-20, 0, 11});
    region.generate(gen);
    one.position = {-20 + 1, 0, 11 + 2};
    gen.generate_chunk(one, *g_reg);
    CHECK(region.download(1, 0, 2).blocks_yzx == one.blocks_yzx);
    // raw noise: the generator's elevation noise through the mirror equals the oracle's
    const std::vector<double> xs = {0.25, -3.5, 100.125}, ys = {7.0, 0.5, -42.75};
    const auto ns = g_gpu->stdgen_noise(7, 5, xs, ys);
    CHECK(ns.size() == 3 && ns[0] >= -1.0 && ns[0] <= 1.0);
    g_gpu->synchronize();
    const auto kt = g_gpu->kernel_time_ms(BXW_KERNEL_MESH, true);
    CHECK(kt.second > 0);
}

int main() {
    try {
        Gpu gpu(0);
        VoxelRegistry reg;
        register_standard_blocks(reg, client_texture_id);
        g_gpu = &gpu;
        g_reg = &reg;
        gpu.bind(reg);
        RUN(rle_compress_zero_test);
        RUN(rle_compress_one_test);
        RUN(rle_decompress_zero_test);
        RUN(rle_decompress_one_test);
        RUN(rle_random_cmp);
        RUN(rle_decompress_error);
        RUN(generate_chunk_matches_oracle);
        RUN(generate_chunk_needs_the_standard_names);
        RUN(mesh_from_chunk_matches_oracle);
        RUN(mesh_from_chunk_panics_on_unknown_ids);
        RUN(region_generate_and_mesh_cfg1);
        RUN(apply_voxel_changes_and_remesh);
        RUN(load_anchor_ticks);
        RUN(region_plumbing);
        std::printf("test result: %s. %d passed; %d failed; %llu kernel launches\n", g_failed ? "FAILED" : "ok", g_run - g_failed, g_failed,
                    (unsigned long long)gpu.kernel_launches());
    } catch (const std::exception &e) {
        std::printf("test harness error: %s\n", e.what());
        return 2;
    }
    return g_failed ? 1 : 0;
}
