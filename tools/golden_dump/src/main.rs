//! Golden-vector dump for the generator half of the path. NOT compiled in the build image (no Rust toolchain);
//! shipped so that the first machine with cargo can pin `oracle/` against the real crates.
//!
//! Output (JSON): for StdGenerator seeds 0 and 0x13372137CAFE
//!   - SHA-256 of the 131072 little-endian bytes of every chunk of the 18x10x18 box around region (0,0,0)+(16,8,16)
//!   - the raw u32 words of chunk (9,0,0) (spawn column) for a direct diff
//! Compare with `python tools/golden_compare.py <out.json>` (uses oracle/ on CPU and the CUDA path on a GPU box).
use bxw_world::stdgen::StdGenerator;
use bxw_world::voxregistry::VoxelRegistry;
use bxw_world::{ChunkPosition, UncompressedChunk};
use sha2::{Digest, Sha256};

fn main() {
    let out = std::env::args().nth(1).unwrap_or_else(|| "golden.json".into());
    let mut registry = VoxelRegistry::new();
    bxw_world::blocks::register_standard_blocks(&mut registry, &|_| 0); // ids only matter for generation
    let mut json = String::from("{\n");
    for (si, seed) in [0u64, 0x13372137CAFEu64].iter().enumerate() {
        let gen = StdGenerator::new(*seed);
        json += &format!("  \"seed_{}\": {{\n    \"chunks\": {{\n", seed);
        let mut first = true;
        for y in -1..9 {
            for z in -1..17 {
                for x in -1..17 {
                    let mut c = UncompressedChunk::new();
                    c.position = ChunkPosition::new(x, y, z);
                    gen.generate_chunk(&mut c, &registry);
                    let mut h = Sha256::new();
                    for v in c.blocks_yzx.iter() {
                        h.update(v.repr().to_le_bytes());
                    }
                    if !first {
                        json += ",\n";
                    }
                    first = false;
                    json += &format!("      \"{},{},{}\": \"{:x}\"", x, y, z, h.finalize());
                }
            }
        }
        json += "\n    }\n  }";
        json += if si == 0 { ",\n" } else { "\n" };
    }
    json += "}\n";
    std::fs::write(out, json).unwrap();
}
