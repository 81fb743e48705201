// test_host_cpu.cpp — the parts of the C++ host mirror that need no GPU: coordinate rules, registry, save blobs, and the
// "no CPU fallback" behaviour (constructing a Gpu without a CUDA device must throw). Built and run by tests/test_cpp_host.py.
#include <cstdio>
#include <set>

#include "bxw_host.hpp"

using namespace bxw;

static int g_failed = 0, g_run = 0;
#define CHECK(cond)                                                           \
    do {                                                                      \
        if (!(cond)) {                                                        \
            std::printf("    FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            g_failed++;                                                       \
            return;                                                           \
        }                                                                     \
    } while (0)
#define RUN(fn)                                                                     \
    do {                                                                            \
        const int before = g_failed;                                                \
        g_run++;                                                                    \
        fn();                                                                       \
        std::printf("test %s ... %s\n", #fn, g_failed == before ? "ok" : "FAILED"); \
    } while (0)

static void block_index_rule() { // lib.rs:104-108: x + 32 z + 1024 y with floor remainders
    CHECK((BlockPosition{0, 0, 0}.as_blockidx() == 0));
    CHECK((BlockPosition{1, 2, 3}.as_blockidx() == 1 + 32 * 3 + 1024 * 2));
    CHECK((BlockPosition{-1, -1, -1}.as_blockidx() == 31 + 32 * 31 + 1024 * 31));
    CHECK((BlockPosition{-33, 64, 95}.as_blockidx() == 31 + 32 * 31 + 1024 * 0));
    CHECK((BlockPosition{-1, -32, -33}.chunk() == ChunkPosition{-1, -1, -2}));
    CHECK((BlockPosition{31, 32, 63}.chunk() == ChunkPosition{0, 1, 1}));
}

static void touching_chunks_rule() { // lib.rs:110-135
    CHECK((BlockPosition{5, 5, 5}.touching_chunks().size() == 1));
    CHECK((BlockPosition{0, 5, 5}.touching_chunks().size() == 2));
    CHECK((BlockPosition{0, 31, 5}.touching_chunks().size() == 3));
    const auto corner = BlockPosition{-32, 31, 64}.touching_chunks(); // local (0, 31, 0) of chunk (-1, 0, 2)
    CHECK(corner.size() == 4);
    const std::set<ChunkPosition> got(corner.begin(), corner.end());
    CHECK((got == std::set<ChunkPosition>{{-1, 0, 2}, {-2, 0, 2}, {-1, 1, 2}, {-1, 0, 1}}));
}

static void neighbour_order() { // worldmgr.rs:748-759 / voxmesh.rs:72: slot = rx + 3 rz + 9 ry
    const auto n = iter_neighbors({10, 20, 30});
    CHECK((n[13] == ChunkPosition{10, 20, 30}));
    CHECK((n[0] == ChunkPosition{9, 19, 29}));
    CHECK((n[1] == ChunkPosition{10, 19, 29}));
    CHECK((n[3] == ChunkPosition{9, 19, 30}));
    CHECK((n[9] == ChunkPosition{9, 20, 29}));
    CHECK((n[26] == ChunkPosition{11, 21, 31}));
}

static void datum_and_meta() { // lib.rs:173-206, stdshapes.rs:8-44
    const VoxelDatum d(7, StdMeta::from_parts(2, 13));
    CHECK(d.repr == ((7u << 16) | (13u * 64 + 2)));
    CHECK(d.id() == 7 && StdMeta::shape(d.meta()) == 2 && StdMeta::orientation(d.meta()) == 13);
    CHECK((VoxelDatum::from_repr(d.repr) == d));
}

static void registry_rules() { // voxregistry.rs:63-150, blocks/mod.rs:6-65
    VoxelRegistry reg;
    CHECK(reg.len() == 1 && reg.get_definition_from_name("core:void")->id == 0 && reg.get_definition_from_id(0).mesh == VoxelMesh::None);
    register_standard_blocks(reg, client_texture_id);
    CHECK(reg.len() == 8);
    const char *names[8] = {"core:void", "core:grass", "core:snow_grass", "core:dirt", "core:stone", "core:diamond_ore", "core:debug", "core:table"};
    for (int i = 0; i < 8; i++) CHECK(reg.get_definition_from_name(names[i])->id == i);
    CHECK((reg.get_definition_from_name("core:grass")->texture_mapping.mapping == std::array<uint32_t, 6>{16, 16, 15, 29, 16, 16})); // new_tsb
    CHECK((reg.get_definition_from_name("core:debug")->texture_mapping.mapping == std::array<uint32_t, 6>{12, 13, 10, 14, 11, 9}));
    CHECK(reg.get_definition_from_name("core:nope") == nullptr);
    bool threw = false;
    try {
        reg.get_definition_from_datum(VoxelDatum(99, 0));
    } catch (const Error &e) {
        threw = e.code == BXW_E_UNKNOWN_ID;
    }
    CHECK(threw);
    const auto ids = StdGenerator::resolve_ids(reg); // void, grass, dirt, stone, snow_grass
    CHECK((ids == std::array<VoxelId, 5>{0, 1, 3, 4, 2}));
    const auto defs = reg.as_defs();
    CHECK(defs.size() == 8 && defs[0].mesh_kind == 0 && defs[4].mesh_kind == 1 && defs[4].tex[0] == 56);
    // voxregistry.rs:63-82,118-133: a builder takes its id when it is created; an abandoned builder consumes one; a duplicate
    // NAME is not an error, the name lookup moves to the new definition (AlreadyExists needs an occupied id slot)
    { auto abandoned = reg.build_definition(); (void)abandoned; } // id 8 is gone
    CHECK(!reg.build_definition().name("core:dirt").finish().has_value()); // id 9
    CHECK(reg.get_definition_from_name("core:dirt")->id == 9 && reg.get_definition_from_id(3).name == "core:dirt");
    CHECK(reg.len() == 10);
    {
        const auto defs = reg.as_defs();
        CHECK(defs.size() == 10 && defs[8].present == 0 && defs[9].present == 1 && defs[3].present == 1);
        bool threw8 = false;
        try {
            reg.get_definition_from_id(8);
        } catch (const Error &e) {
            threw8 = e.code == BXW_E_UNKNOWN_ID;
        }
        CHECK(threw8);
    }
}

static void save_blob_rules() { // generation.rs:158-194
    VChunk c;
    c.position = {1, 2, 3};
    c.vox = {0x01020304u, 0x01020304u, 32766u};
    const auto bytes = c.serialize();
    CHECK(bytes.size() == 12 && bytes[0] == 0x04 && bytes[3] == 0x01);
    const VChunk back = VChunk::deserialize(c.position, bytes);
    CHECK(back.vox == c.vox && back.position == c.position);
    bool threw = false;
    try {
        VChunk::deserialize(c.position, {1, 2, 3});
    } catch (const std::invalid_argument &e) {
        threw = std::string(e.what()) == "Invalid serialized data length";
    }
    CHECK(threw);
}

int main(int argc, char **argv) {
    RUN(block_index_rule);
    RUN(touching_chunks_rule);
    RUN(neighbour_order);
    RUN(datum_and_meta);
    RUN(registry_rules);
    RUN(save_blob_rules);
    // no CPU fallback: without a device the context cannot be created (pass --expect-no-gpu on a CPU-only machine)
    if (argc > 1 && std::string(argv[1]) == "--expect-no-gpu") {
        g_run++;
        bool threw = false;
        try {
            Gpu gpu(0);
        } catch (const Error &) {
            threw = true;
        }
        std::printf("test gpu_required ... %s\n", threw ? "ok" : "FAILED");
        if (!threw) g_failed++;
    }
    std::printf("test result: %s. %d passed; %d failed\n", g_failed ? "FAILED" : "ok", g_run - g_failed, g_failed);
    return g_failed ? 1 : 0;
}
