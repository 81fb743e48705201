// bxw_host.hpp — C++ host side of the drop-in: the reference's Rust interfaces for the chunk pipeline, same names,
// argument meaning and error behaviour, implemented on top of the C ABI (include/bxw_cuda.h -> libbxw_cuda.so).
//
// The reference is Rust and the build image has no Rust toolchain, so this header is what a maintainer's Rust shims
//  look like when written against the same ABI in C++. Nothing here computes on the CPU: every
// operation that touches voxels goes through libbxw_cuda.so, and a missing CUDA device makes Gpu's constructor throw.
//
//   reference item (file:line, relative to the reference tree)                      here
//   VoxelDatum                       bxw_world/src/lib.rs:173-206                   bxw::VoxelDatum
//   ChunkPosition / BlockPosition    bxw_world/src/lib.rs:36-135                    bxw::ChunkPosition, bxw::BlockPosition
//   UncompressedChunk                bxw_world/src/lib.rs:222-235                   bxw::UncompressedChunk
//   VChunk::{compress,decompress}    bxw_world/src/lib.rs:237-345, 549-582          bxw::VChunk
//   TextureMapping                   bxw_world/src/lib.rs:584-627                   bxw::TextureMapping
//   VoxelRegistry + builder          bxw_world/src/voxregistry.rs:14-150            bxw::VoxelRegistry
//   register_standard_blocks         bxw_world/src/blocks/mod.rs:6-65               bxw::register_standard_blocks
//   StdGenerator::{new,generate_chunk} bxw_world/src/stdgen.rs:285-375              bxw::StdGenerator
//   mesh_from_chunk, is_chunk_trivial  src/client/render/voxmesh.rs:16-190          bxw::mesh_from_chunk, bxw::is_chunk_trivial
//   iter_neighbors                   bxw_world/src/worldmgr.rs:748-759              bxw::iter_neighbors
//   WorldBlocks / MeshDataHandler (ChunkDataHandler pair), World::apply_voxel_changes
//                                    bxw_world/src/generation.rs:11-203, src/client/render/voxrender.rs:192-448,
//                                    bxw_world/src/worldmgr.rs:312-357              bxw::Region (batched, GPU-resident)
#pragma once

#include <array>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "bxw_cuda.h"

namespace bxw {

constexpr int CHUNK_DIM = 32;
constexpr int CHUNK_DIM2 = CHUNK_DIM * CHUNK_DIM;
constexpr int CHUNK_DIM3 = CHUNK_DIM * CHUNK_DIM * CHUNK_DIM;

// What the reference does with panic!/expect()/unwrap(): the message of bxw_last_error plus the ABI code.
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

using VoxelId = uint16_t;
using VoxelMetadata = uint16_t;

struct VoxelDatum { // lib.rs:173-206
    uint32_t repr = 0;
    VoxelDatum() = default;
    VoxelDatum(VoxelId id, VoxelMetadata meta) : repr(((uint32_t)id << 16) | meta) {}
    static VoxelDatum from_repr(uint32_t r) {
        VoxelDatum d;
        d.repr = r;
        return d;
    }
    VoxelId id() const { return (VoxelId)(repr >> 16); }
    VoxelMetadata meta() const { return (VoxelMetadata)(repr & 0xFFFF); }
    bool operator==(const VoxelDatum &o) const { return repr == o.repr; }
};

// StdMeta (blocks/stdshapes.rs:8-44): meta = orientation * 64 + shape
struct StdMeta {
    static VoxelMetadata from_parts(unsigned shape, unsigned orientation) { return (VoxelMetadata)(orientation * 64 + shape); }
    static unsigned shape(VoxelMetadata m) { return m % 64; }
    static unsigned orientation(VoxelMetadata m) { return m / 64; }
};

struct ChunkPosition { // lib.rs:36-83
    int32_t x = 0, y = 0, z = 0;
    bool operator==(const ChunkPosition &o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator<(const ChunkPosition &o) const { return std::tie(x, y, z) < std::tie(o.x, o.y, o.z); }
};

inline int32_t div_floor(int32_t a, int32_t b) { return a / b - ((a % b != 0) && ((a < 0) != (b < 0))); }
inline int32_t rem_floor(int32_t a, int32_t b) { return a - div_floor(a, b) * b; }

struct BlockPosition { // lib.rs:86-135
    int32_t x = 0, y = 0, z = 0;
    ChunkPosition chunk() const { return {div_floor(x, CHUNK_DIM), div_floor(y, CHUNK_DIM), div_floor(z, CHUNK_DIM)}; }
    size_t as_blockidx() const { // lib.rs:104-108
        return (size_t)rem_floor(x, CHUNK_DIM) + (size_t)CHUNK_DIM * rem_floor(z, CHUNK_DIM) + (size_t)CHUNK_DIM2 * rem_floor(y, CHUNK_DIM);
    }
    std::vector<ChunkPosition> touching_chunks() const { // lib.rs:110-135: own chunk + the face neighbours it borders
        std::vector<ChunkPosition> out;
        const ChunkPosition c = chunk();
        out.push_back(c);
        const int32_t l[3] = {rem_floor(x, CHUNK_DIM), rem_floor(y, CHUNK_DIM), rem_floor(z, CHUNK_DIM)};
        for (int a = 0; a < 3; a++) {
            if (l[a] != 0 && l[a] != CHUNK_DIM - 1) continue;
            ChunkPosition n = c;
            (a == 0 ? n.x : a == 1 ? n.y : n.z) += l[a] == 0 ? -1 : 1;
            out.push_back(n);
        }
        return out;
    }
};

// iter_neighbors(cpos, include_self = true) (worldmgr.rs:748-759): slot = rx + 3 rz + 9 ry
inline std::array<ChunkPosition, 27> iter_neighbors(ChunkPosition c) {
    std::array<ChunkPosition, 27> out;
    int k = 0;
    for (int ry = -1; ry <= 1; ry++)
        for (int rz = -1; rz <= 1; rz++)
            for (int rx = -1; rx <= 1; rx++) out[k++] = {c.x + rx, c.y + ry, c.z + rz};
    return out;
}

struct UncompressedChunk { // lib.rs:222-235: zero-initialised, caller sets position
    ChunkPosition position;
    std::vector<uint32_t> blocks_yzx = std::vector<uint32_t>(CHUNK_DIM3, 0u);
    VoxelDatum get(BlockPosition p) const { return VoxelDatum::from_repr(blocks_yzx[p.as_blockidx()]); }
    void set(BlockPosition p, VoxelDatum d) { blocks_yzx[p.as_blockidx()] = d.repr; }
};

enum class VoxelMesh { None, CubeAndSlopes }; // lib.rs:651-655

template <class T> struct TextureMapping { // lib.rs:584-627, indexed by Direction::to_signed_axis_index
    std::array<T, 6> mapping{};
    static TextureMapping new_single(T t) { return {{t, t, t, t, t, t}}; }
    static TextureMapping new_tsb(T top, T side, T bottom) { return {{side, side, bottom, top, side, side}}; }
    static TextureMapping from_array(std::array<T, 6> m) { return {m}; }
    template <class F> auto map(F f) const {
        using U = decltype(f(mapping[0]));
        TextureMapping<U> out;
        for (int i = 0; i < 6; i++) out.mapping[i] = f(mapping[i]);
        return out;
    }
};

struct VoxelDefinition { // lib.rs:629-649
    VoxelId id = 0;
    std::string name;
    VoxelMesh mesh = VoxelMesh::CubeAndSlopes;
    std::array<float, 3> debug_color{{1.0f, 1.0f, 1.0f}};
    TextureMapping<uint32_t> texture_mapping;
};

enum class VoxelDefinitionError { AlreadyExists };

class VoxelRegistry { // voxregistry.rs:26-150; Default holds core:void (id 0, no mesh)
  public:
    class Builder { // VoxelDefinitionBuilder (voxregistry.rs:14-92)
      public:
        Builder &name(const std::string &v) {
            def_.name = v;
            return *this;
        }
        Builder &set_mesh(VoxelMesh m) {
            def_.mesh = m;
            return *this;
        }
        Builder &debug_color(float r, float g, float b) {
            def_.debug_color = {{r, g, b}};
            return *this;
        }
        Builder &texture_names(const std::function<uint32_t(const std::string &)> &mapper, const TextureMapping<std::string> &names) {
            def_.texture_mapping = names.map([&](const std::string &n) { return mapper(n); });
            return *this;
        }
        Builder &texture_ids(const TextureMapping<uint32_t> &ids) {
            def_.texture_mapping = ids;
            return *this;
        }
        // Ok(()) / Err(AlreadyExists): returns nullopt on success like Result<(), E>::err(). As in voxregistry.rs:63-82 the
        // error is raised only when the ID slot is occupied; a duplicate NAME re-points the name lookup at the new definition.
        std::optional<VoxelDefinitionError> finish() {
            const size_t idx = def_.id;
            if (reg_.definitions_.size() <= idx) reg_.definitions_.resize(idx * 2 + 1);
            else if (reg_.definitions_[idx].has_value()) return VoxelDefinitionError::AlreadyExists;
            reg_.name_lut_[def_.name] = def_.id;
            reg_.definitions_[idx] = def_;
            reg_.version_++;
            return std::nullopt;
        }

      private:
        friend class VoxelRegistry;
        Builder(VoxelRegistry &r, VoxelId id) : reg_(r) { def_.id = id; }
        VoxelRegistry &reg_;
        VoxelDefinition def_;
    };

    VoxelRegistry() : token_(next_token()) { build_definition().name("core:void").set_mesh(VoxelMesh::None).debug_color(0, 0, 0).finish(); }
    // the id is taken here (last_free_id += 1, voxregistry.rs:118-133): an abandoned builder consumes an id
    Builder build_definition() { return Builder(*this, last_free_id_++); }
    const VoxelDefinition &get_definition_from_id(VoxelId id) const {
        if (id >= definitions_.size() || !definitions_[id].has_value())
            throw Error(BXW_E_UNKNOWN_ID, "no definition for block id " + std::to_string(id)); // unwrap() panic
        return *definitions_[id];
    }
    const VoxelDefinition &get_definition_from_datum(VoxelDatum d) const { return get_definition_from_id(d.id()); }
    const VoxelDefinition *get_definition_from_name(const std::string &name) const {
        auto it = name_lut_.find(name);
        return it == name_lut_.end() ? nullptr : &*definitions_[it->second];
    }
    // number of id slots up to the highest defined one (what the device registry holds)
    size_t len() const {
        size_t n = definitions_.size();
        while (n && !definitions_[n - 1].has_value()) n--;
        return n;
    }
    uint64_t version() const { return version_; }
    uint64_t token() const { return token_; } // process-wide unique: addresses of registries recur
    std::vector<bxw_block_def> as_defs() const {
        std::vector<bxw_block_def> out(len());
        for (size_t i = 0; i < out.size(); i++) {
            out[i] = bxw_block_def{};
            if (!definitions_[i].has_value()) continue; // free slot: present = 0
            out[i].present = 1;
            out[i].mesh_kind = definitions_[i]->mesh == VoxelMesh::None ? 0 : 1;
            for (int k = 0; k < 3; k++) out[i].debug_color[k] = definitions_[i]->debug_color[k];
            for (int k = 0; k < 6; k++) out[i].tex[k] = definitions_[i]->texture_mapping.mapping[k];
        }
        return out;
    }

  private:
    static uint64_t next_token() {
        static std::atomic<uint64_t> n{1};
        return n.fetch_add(1);
    }
    std::vector<std::optional<VoxelDefinition>> definitions_;
    std::map<std::string, VoxelId> name_lut_;
    VoxelId last_free_id_ = 0;
    uint64_t version_ = 0;
    uint64_t token_;
};

// blocks/mod.rs:6-65
inline void register_standard_blocks(VoxelRegistry &vxreg, const std::function<uint32_t(const std::string &)> &texmapper) {
    using TM = TextureMapping<std::string>;
    auto add = [&](const char *name, const TM &tex) {
        if (vxreg.build_definition().name(name).texture_names(texmapper, tex).finish()) throw std::logic_error("AlreadyExists"); // .unwrap()
    };
    add("core:grass", TM::new_tsb("grass_top", "dirt_grass", "dirt"));
    add("core:snow_grass", TM::new_tsb("snow", "dirt_snow", "dirt"));
    add("core:dirt", TM::new_single("dirt"));
    add("core:stone", TM::new_single("stone"));
    add("core:diamond_ore", TM::new_single("stone_diamond"));
    add("core:debug", TM::from_array({{"dbg_left", "dbg_right", "dbg_down", "dbg_up", "dbg_front", "dbg_back"}}));
    add("core:table", TM::new_tsb("table", "wood", "table"));
}

// Texture ids of the client path: rank of the name among the byte-wise sorted keys of res/manifest.toml
// (src/client/render/resources.rs:149-182 iterates a toml 0.5 BTreeMap). Only the names the standard blocks use.
inline uint32_t client_texture_id(const std::string &name) {
    static const std::map<std::string, uint32_t> ids = {
        {"dbg_back", 9}, {"dbg_down", 10}, {"dbg_front", 11}, {"dbg_left", 12}, {"dbg_right", 13}, {"dbg_up", 14}, {"dirt", 15},
        {"dirt_grass", 16}, {"dirt_snow", 18}, {"grass_top", 29}, {"snow", 55}, {"stone", 56}, {"stone_diamond", 61}, {"table", 73},
        {"wood", 90}};
    return ids.at(name);
}

// One context per CUDA device (the Rust side would hold it in a Mutex, like vmalloc in voxrender.rs:304).
class Gpu {
  public:
    explicit Gpu(int device = 0) {
        bxw_ctx *c = nullptr;
        const int rc = bxw_create(device, &c);
        if (rc != BXW_OK) throw Error(rc, "bxw_create failed (no usable sm_100 device? there is no CPU fallback)");
        ctx_.reset(c);
    }
    bxw_ctx *raw() const { return ctx_.get(); }
    void check(int rc) const {
        if (rc != BXW_OK) throw Error(rc, bxw_last_error(ctx_.get()));
    }
    // upload the registry when it changed since the last call on this context
    void bind(const VoxelRegistry &reg) {
        if (bound_token_ == reg.token() && bound_version_ == reg.version()) return;
        const auto defs = reg.as_defs();
        check(bxw_set_registry(raw(), defs.data(), (uint32_t)defs.size()));
        bound_token_ = reg.token();
        bound_version_ = reg.version();
    }
    uint64_t kernel_launches() const { return bxw_kernel_launches(raw()); }
    // everything the context launches goes to this CUDA stream (a cudaStream_t; nullptr = the context's own)
    void set_stream(void *cuda_stream) { check(bxw_set_stream(raw(), cuda_stream)); }
    void set_option(int option, int value) { check(bxw_set_option(raw(), option, value)); } // BXW_OPT_*
    void synchronize() { check(bxw_synchronize(raw())); }
    // device time of the library's own launches since the last reset: {milliseconds, launches} (BXW_KERNEL_*)
    std::pair<float, uint64_t> kernel_time_ms(int which, bool reset = false) {
        float ms = 0.0f;
        uint64_t n = 0;
        check(bxw_kernel_time_ms(raw(), which, &ms, &n, reset ? 1 : 0));
        return {ms, n};
    }
    // raw SuperSimplex samples: StdGenerator's seven noise functions (stdgen.rs:82-114; which = 0..6) and bxw_terragen's
    // (bxw_terragen/src/supersimplex.rs)
    std::vector<double> stdgen_noise(uint64_t seed, int which, const std::vector<double> &xs, const std::vector<double> &ys, int precision = 64) {
        if (xs.size() != ys.size()) throw std::invalid_argument("xs and ys differ in length");
        std::vector<double> out(xs.size());
        check(bxw_stdgen_noise(raw(), seed, which, precision, xs.data(), ys.data(), xs.size(), out.data()));
        return out;
    }
    std::vector<double> terragen_noise2(uint64_t seed, const std::vector<double> &xs, const std::vector<double> &ys) {
        if (xs.size() != ys.size()) throw std::invalid_argument("xs and ys differ in length");
        std::vector<double> out(xs.size());
        check(bxw_terragen_noise2(raw(), seed, xs.data(), ys.data(), xs.size(), out.data()));
        return out;
    }

  private:
    struct Del {
        void operator()(bxw_ctx *c) const { bxw_destroy(c); }
    };
    std::unique_ptr<bxw_ctx, Del> ctx_;
    uint64_t bound_token_ = 0;
    uint64_t bound_version_ = 0;
};

struct VChunk { // lib.rs:237-261, 549-582; Default = the placeholder [0, 0, CHUNK_DIM3]
    ChunkPosition position;
    std::vector<uint32_t> vox{0u, 0u, (uint32_t)CHUNK_DIM3};

    void compress(Gpu &gpu, const UncompressedChunk &from) { // lib.rs:555-559
        if (!(from.position == position)) throw std::logic_error("assertion failed: self.position == from.position");
        std::vector<uint32_t> out(bxw_rle_bound(1));
        uint64_t offs[2] = {0, 0};
        gpu.check(bxw_rle_compress(gpu.raw(), from.blocks_yzx.data(), 1, out.data(), offs));
        out.resize(offs[1]);
        vox = std::move(out);
    }
    std::unique_ptr<UncompressedChunk> decompress(Gpu &gpu) const { // lib.rs:562-567 (.expect on a bad stream)
        auto c = std::make_unique<UncompressedChunk>();
        c->position = position;
        const uint64_t offs[2] = {0, vox.size()};
        gpu.check(bxw_rle_decompress(gpu.raw(), vox.data(), offs, 1, c->blocks_yzx.data()));
        return c;
    }
    // WorldBlocks::serialize_data / deserialize_data (generation.rs:158-194): the words as little-endian bytes
    std::vector<uint8_t> serialize() const {
        std::vector<uint8_t> out(vox.size() * 4);
        for (size_t i = 0; i < vox.size(); i++)
            for (int b = 0; b < 4; b++) out[4 * i + b] = (uint8_t)(vox[i] >> (8 * b));
        return out;
    }
    static VChunk deserialize(ChunkPosition position, const std::vector<uint8_t> &data) {
        if (data.size() % 4 != 0) throw std::invalid_argument("Invalid serialized data length");
        VChunk c;
        c.position = position;
        c.vox.resize(data.size() / 4);
        for (size_t i = 0; i < c.vox.size(); i++)
            c.vox[i] = (uint32_t)data[4 * i] | (uint32_t)data[4 * i + 1] << 8 | (uint32_t)data[4 * i + 2] << 16 | (uint32_t)data[4 * i + 3] << 24;
        return c;
    }
};

// stdgen.rs:285-375. generate_chunk panics (.expect) when a core:* name is missing; here that is a std::runtime_error
// with the reference's message.
class StdGenerator {
  public:
    StdGenerator(Gpu &gpu, uint64_t seed) : gpu_(gpu), seed_(seed) {}
    uint64_t seed() const { return seed_; }
    static std::array<VoxelId, 5> resolve_ids(const VoxelRegistry &registry) { // stdgen.rs:307-326
        const char *names[5] = {"core:void", "core:grass", "core:dirt", "core:stone", "core:snow_grass"};
        const char *what[5] = {"air", "grass", "dirt", "stone", "snow grass"};
        std::array<VoxelId, 5> ids{};
        for (int i = 0; i < 5; i++) {
            const VoxelDefinition *d = registry.get_definition_from_name(names[i]);
            if (!d) throw std::runtime_error(std::string("No standard ") + what[i] + " block definition found");
            ids[i] = d->id;
        }
        return ids;
    }
    void generate_chunk(UncompressedChunk &chunk, const VoxelRegistry &registry) const { // stdgen.rs:304-374
        const auto ids = resolve_ids(registry);
        gpu_.check(bxw_set_gen_ids(gpu_.raw(), ids[0], ids[1], ids[2], ids[3], ids[4]));
        gpu_.check(bxw_gen_chunk(gpu_.raw(), seed_, chunk.position.x, chunk.position.y, chunk.position.z, chunk.blocks_yzx.data()));
    }

  private:
    Gpu &gpu_;
    uint64_t seed_;
};

struct ChunkBuffers { // voxrender.rs:34-37
    std::vector<bxw_vertex> vertices;
    std::vector<uint32_t> indices;
};

// voxmesh.rs:16-25
inline bool is_chunk_trivial(Gpu &gpu, const VChunk &chunk, const VoxelRegistry &registry) {
    gpu.bind(registry);
    int t = 0;
    gpu.check(bxw_is_chunk_trivial(gpu.raw(), chunk.vox.data(), chunk.vox.size(), &t));
    return t != 0;
}

// voxmesh.rs:45-190: chunks in iter_neighbors(cpos, true) order; always Some; unknown block ids throw (reference: panic)
inline std::optional<ChunkBuffers> mesh_from_chunk(Gpu &gpu, const VoxelRegistry &registry, const std::vector<const VChunk *> &chunks,
                                                   std::pair<uint32_t, uint32_t> /*texture_dim*/) {
    if (chunks.size() != 27) throw std::logic_error("assertion failed: chunks.len() == 27");
    gpu.bind(registry);
    const uint32_t *ptrs[27];
    size_t lens[27];
    for (int i = 0; i < 27; i++) {
        ptrs[i] = chunks[i]->vox.data();
        lens[i] = chunks[i]->vox.size();
    }
    bxw_mesh_span span{};
    gpu.check(bxw_mesh_chunk27_rle(gpu.raw(), ptrs, lens, &span));
    ChunkBuffers out;
    out.vertices.resize(span.v_count);
    out.indices.resize(span.i_count);
    gpu.check(bxw_mesh_fetch(gpu.raw(), out.vertices.data(), out.indices.data()));
    return out;
}

struct VoxelChange { // generation.rs:39-64 / worldmgr.rs:312-357
    BlockPosition bpos;
    VoxelDatum from, to;
};

// The batched, GPU-resident form of the two ChunkDataHandlers (WorldBlocks + MeshDataHandler) and of
// World::apply_voxel_changes: one box of chunks (plus a 1-chunk shell so every chunk has its 26 neighbours,
// worldmgr.rs:761-790) generated, stored and meshed per call.
class Region {
  public:
    Region(Gpu &gpu, ChunkPosition origin, std::array<uint32_t, 3> dims, const VoxelRegistry &registry)
        : gpu_(gpu), origin_(origin), dims_(dims), registry_(registry) {
        gpu_.bind(registry_);
        gpu_.check(bxw_region_create(gpu_.raw(), origin.x, origin.y, origin.z, dims[0], dims[1], dims[2]));
    }
    ~Region() { bxw_region_destroy(gpu_.raw()); }
    Region(const Region &) = delete;
    Region &operator=(const Region &) = delete;

    size_t n_chunks() const { return (size_t)dims_[0] * dims_[1] * dims_[2]; }
    // WorldBlocks::create_chunk_update_task for every chunk of the box (generation.rs:96-148)
    void generate(const StdGenerator &g) {
        const auto ids = StdGenerator::resolve_ids(registry_);
        gpu_.check(bxw_set_gen_ids(gpu_.raw(), ids[0], ids[1], ids[2], ids[3], ids[4]));
        gpu_.check(bxw_region_generate(gpu_.raw(), g.seed(), 64));
    }
    // MeshDataHandler::create_chunk_update_task for every chunk of the box (voxrender.rs:224-418)
    std::pair<uint64_t, uint64_t> mesh() {
        gpu_.bind(registry_);
        uint64_t v = 0, i = 0;
        gpu_.check(bxw_region_mesh(gpu_.raw(), &v, &i));
        return {v, i};
    }
    std::pair<uint64_t, uint64_t> generate_and_mesh(const StdGenerator &g) {
        gpu_.bind(registry_);
        const auto ids = StdGenerator::resolve_ids(registry_);
        gpu_.check(bxw_set_gen_ids(gpu_.raw(), ids[0], ids[1], ids[2], ids[3], ids[4]));
        uint64_t v = 0, i = 0;
        gpu_.check(bxw_region_generate_and_mesh(gpu_.raw(), g.seed(), 64, &v, &i));
        return {v, i};
    }
    std::vector<bxw_mesh_span> spans() const { // chunk order x + nx (z + nz y)
        std::vector<bxw_mesh_span> s(n_chunks());
        gpu_.check(bxw_region_mesh_spans(gpu_.raw(), s.data()));
        return s;
    }
    ChunkBuffers fetch_arena(uint64_t n_vertices, uint64_t n_indices) const {
        ChunkBuffers b;
        b.vertices.resize(n_vertices);
        b.indices.resize(n_indices);
        gpu_.check(bxw_region_mesh_fetch(gpu_.raw(), b.vertices.data(), n_vertices, b.indices.data(), n_indices));
        return b;
    }
    // WorldBlocks::get_data (generation.rs:80-84): the chunk at local coordinates as a VChunk, compressed on the device
    VChunk get_data(int32_t lx, int32_t ly, int32_t lz) const {
        const int32_t lpos[3] = {lx, ly, lz};
        uint64_t offs[2] = {0, 0};
        gpu_.check(bxw_region_export_rle(gpu_.raw(), lpos, 1, nullptr, 0, offs));
        VChunk c;
        c.position = {origin_.x + lx, origin_.y + ly, origin_.z + lz};
        c.vox.resize(offs[1]);
        gpu_.check(bxw_region_export_rle(gpu_.raw(), lpos, 1, c.vox.data(), c.vox.size(), offs));
        return c;
    }
    // WorldBlocks::swap_data / deserialize_data (generation.rs:86-94, 168-194)
    void set_data(int32_t lx, int32_t ly, int32_t lz, const VChunk &c) {
        const int32_t lpos[3] = {lx, ly, lz};
        const uint64_t offs[2] = {0, c.vox.size()};
        gpu_.check(bxw_region_import_rle(gpu_.raw(), lpos, 1, c.vox.data(), offs));
    }
    // World::apply_voxel_changes (worldmgr.rs:312-357): ordered compare-and-swap, touching_chunks become dirty.
    // Returns the number of dirty chunks of the box.
    uint32_t apply_voxel_changes(const std::vector<VoxelChange> &changes) {
        std::vector<bxw_voxel_change> ch(changes.size());
        for (size_t i = 0; i < changes.size(); i++)
            ch[i] = {changes[i].bpos.x, changes[i].bpos.y, changes[i].bpos.z, changes[i].from.repr, changes[i].to.repr};
        uint32_t nd = 0;
        gpu_.check(bxw_region_apply_changes(gpu_.raw(), ch.data(), (uint32_t)ch.size(), &nd));
        return nd;
    }
    // MeshDataHandler re-running create_chunk_update_task for the dirty chunks (worldmgr.rs:335-356). The arena extent may
    // have grown (moved meshes) or changed altogether (compaction): arena() is what fetch_arena wants afterwards.
    uint32_t remesh_dirty() {
        gpu_.bind(registry_);
        uint32_t n = 0;
        gpu_.check(bxw_region_remesh_dirty(gpu_.raw(), &n));
        return n;
    }
    // World::check_load_deltas + recalculate_load_deltas + the delta loop of main_loop_tick (worldmgr.rs:407-732) for the box:
    // anchors = the CLoadAnchor entities (position, radius, load_mesh). Returns {deltas consumed, chunks generated, chunks
    // meshed, meshes unloaded}; meshes of newly generated chunks follow on the next tick, as in the reference.
    std::array<uint32_t, 4> tick(const std::vector<bxw_load_anchor> &anchors, uint64_t seed, int precision = 64, uint32_t max_deltas = 0) {
        gpu_.bind(registry_);
        const auto ids = StdGenerator::resolve_ids(registry_);
        gpu_.check(bxw_set_gen_ids(gpu_.raw(), ids[0], ids[1], ids[2], ids[3], ids[4]));
        uint32_t nu = 0, nl = 0;
        gpu_.check(bxw_region_update_anchors(gpu_.raw(), anchors.data(), (uint32_t)anchors.size(), &nu, &nl));
        std::array<uint32_t, 4> st{{0, 0, 0, 0}};
        gpu_.check(bxw_region_apply_deltas(gpu_.raw(), seed, precision, max_deltas, st.data()));
        return st;
    }
    void unload_all() { gpu_.check(bxw_region_unload_all(gpu_.raw())); }
    std::vector<bxw_chunk_delta> load_deltas() const { // the ordered list of the last update_anchors
        uint32_t n = 0;
        gpu_.check(bxw_region_load_deltas(gpu_.raw(), nullptr, 0, &n));
        std::vector<bxw_chunk_delta> out(n);
        if (n) gpu_.check(bxw_region_load_deltas(gpu_.raw(), out.data(), n, &n));
        return out;
    }
    std::pair<uint64_t, uint64_t> arena() const {
        uint64_t v = 0, i = 0;
        gpu_.check(bxw_region_arena_extent(gpu_.raw(), &v, &i));
        return {v, i};
    }
    ChunkBuffers fetch() const {
        const auto a = arena();
        return fetch_arena(a.first, a.second);
    }
    // one chunk's ChunkBuffers (voxrender.rs:34-37) from its span
    ChunkBuffers fetch_chunk(const bxw_mesh_span &sp) const {
        ChunkBuffers b;
        b.vertices.resize(sp.v_count);
        b.indices.resize(sp.i_count);
        if (sp.v_count | sp.i_count)
            gpu_.check(bxw_region_mesh_fetch_range(gpu_.raw(), sp.v_off, sp.v_count, b.vertices.data(), sp.i_off, sp.i_count, b.indices.data()));
        return b;
    }
    // the arena into caller-owned (pinned) memory on the copy stream; valid after fetch_wait(). The next generate / mesh pass
    // overlaps the transfer.
    void fetch_async(bxw_vertex *vertices, uint64_t n_vertices, uint32_t *indices, uint64_t n_indices) {
        gpu_.check(bxw_region_mesh_fetch_async(gpu_.raw(), vertices, n_vertices, indices, n_indices));
    }
    void fetch_wait() { gpu_.check(bxw_region_fetch_wait(gpu_.raw())); }
    // dense chunk in / out (UncompressedChunk::blocks_yzx, lib.rs:222-235), local coordinates, shell = -1 and n
    void upload(int32_t lx, int32_t ly, int32_t lz, const UncompressedChunk &c) {
        gpu_.check(bxw_region_upload_chunk(gpu_.raw(), lx, ly, lz, c.blocks_yzx.data()));
    }
    UncompressedChunk download(int32_t lx, int32_t ly, int32_t lz) const {
        UncompressedChunk c;
        c.position = {origin_.x + lx, origin_.y + ly, origin_.z + lz};
        gpu_.check(bxw_region_download_chunk(gpu_.raw(), lx, ly, lz, c.blocks_yzx.data()));
        return c;
    }
    // multi-GPU steps (one Region per rank, slabs along x): the boundary chunk columns first, their planes exported and sent
    // while the rest is generated, the neighbours' planes imported before meshing
    void generate_part(const StdGenerator &g, int part) {
        if (part == 1) {
            const auto ids = StdGenerator::resolve_ids(registry_);
            gpu_.check(bxw_set_gen_ids(gpu_.raw(), ids[0], ids[1], ids[2], ids[3], ids[4]));
        }
        gpu_.check(bxw_region_generate_part(gpu_.raw(), g.seed(), 64, part));
    }
    size_t halo_bytes() const { return bxw_region_halo_bytes(gpu_.raw()); }
    void halo_export(int face, void *device_buffer) { gpu_.check(bxw_region_halo_export(gpu_.raw(), face, device_buffer)); }
    void halo_import(int face, const void *device_buffer) { gpu_.check(bxw_region_halo_import(gpu_.raw(), face, device_buffer)); }
    // the box moves (its content is dropped: load deltas or generate fill it again)
    void set_origin(ChunkPosition origin) {
        gpu_.check(bxw_region_set_origin(gpu_.raw(), origin.x, origin.y, origin.z));
        origin_ = origin;
    }
    // CellGen::calc_voxel_params of every column of the box, shell included (stdgen.rs:215-276): height, elevation class and
    // the unrounded height
    struct HeightMap {
        uint32_t width = 0, depth = 0; // columns along x, z
        std::vector<int32_t> height;
        std::vector<uint8_t> elevation;
        std::vector<double> raw_height;
    };
    HeightMap heightmap() const {
        HeightMap h;
        h.width = (dims_[0] + 2) * 32;
        h.depth = (dims_[2] + 2) * 32;
        const size_t n = (size_t)h.width * h.depth;
        h.height.resize(n);
        h.elevation.resize(n);
        h.raw_height.resize(n);
        gpu_.check(bxw_region_heightmap(gpu_.raw(), h.height.data(), h.elevation.data(), h.raw_height.data()));
        return h;
    }
    // mesh the chunks inside one load-anchor sphere, nearest first (iter_coords_to_load, worldmgr.rs:734-746): {chunks meshed,
    // arena vertices, arena indices}
    std::array<uint64_t, 3> mesh_sphere(ChunkPosition anchor, uint32_t radius) {
        gpu_.bind(registry_);
        uint32_t n = 0;
        uint64_t v = 0, i = 0;
        gpu_.check(bxw_region_mesh_sphere(gpu_.raw(), anchor.x, anchor.y, anchor.z, radius, &n, &v, &i));
        return {{n, v, i}};
    }
    bxw_arena_stats arena_stats() const {
        bxw_arena_stats st{};
        gpu_.check(bxw_region_arena_stats(gpu_.raw(), &st));
        return st;
    }
    std::pair<uint64_t, uint64_t> compact_arena() { // spans' offsets change
        uint64_t v = 0, i = 0;
        gpu_.check(bxw_region_compact_arena(gpu_.raw(), &v, &i));
        return {v, i};
    }
    // the two load states of the reference's handlers (CHUNK_BLOCK_DATA, CHUNK_MESH_DATA): {blocks loaded: every chunk of the box
    // with its shell, storage order x + SX (z + SZ y); mesh loaded: interior chunks, span order}
    std::pair<std::vector<uint8_t>, std::vector<uint8_t>> load_state() const {
        std::vector<uint8_t> b((size_t)(dims_[0] + 2) * (dims_[1] + 2) * (dims_[2] + 2)), m(n_chunks());
        gpu_.check(bxw_region_load_state(gpu_.raw(), b.data(), m.data()));
        return {b, m};
    }
    // hashed random blocks incl. oriented slopes / corners (BASELINE configs[1])
    void fill_synthetic(uint64_t seed, uint32_t void_threshold) { gpu_.check(bxw_region_fill_synthetic(gpu_.raw(), seed, void_threshold)); }
    // {chunks listed by the last whole-region mesh pass, interior chunks}
    std::pair<uint32_t, uint32_t> mesh_candidates() const {
        uint32_t a = 0, b = 0;
        gpu_.check(bxw_region_mesh_candidates(gpu_.raw(), &a, &b));
        return {a, b};
    }

  private:
    Gpu &gpu_;
    ChunkPosition origin_;
    std::array<uint32_t, 3> dims_;
    const VoxelRegistry &registry_;
};

} // namespace bxw
