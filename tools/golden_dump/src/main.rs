//! Golden-vector dump for the generator half of the path. NOT compiled in the build image (no Rust toolchain);
//! shipped so that the first machine with cargo can pin `oracle/` against the real crates.
//!
//! Output (JSON): for StdGenerator seeds 0 and 0x13372137CAFE
//!   - SHA-256 of the 131072 little-endian bytes of every chunk of the 18x10x18 box around region (0,0,0)+(16,8,16)
//!   - the raw u32 words of chunk (9,0,0) (spawn column) for a direct diff
//!   - "noise": for PermutationTable seeds 0, 1, 12345, 0xFFFFFFFF: hash([i, 0]) for i in 0..256 (= values[values[i]]) and
//!     hash([i, j]) on a 16x16 grid, plus the raw bits of SuperSimplex::new(seed).get([x, y]) on a fixed point set --
//!     settles oracle/noise.c (PermutationTable::new, the rand 0.7.3 shuffle, super_simplex_2d) independently of stdgen
//! Compare with `python tools/golden_compare.py <out.json>` (uses oracle/ on CPU and the CUDA path on a GPU box).
use bxw_world::stdgen::StdGenerator;
use noise::permutationtable::{NoiseHasher, PermutationTable};
use noise::{NoiseFn, SuperSimplex};
use bxw_world::voxregistry::VoxelRegistry;
use bxw_world::{ChunkPosition, UncompressedChunk};
use sha2::{Digest, Sha256};

fn main() {
    let out = std::env::args().nth(1).unwrap_or_else(|| "golden.json".into());
    let mut registry = VoxelRegistry::new();
    bxw_world::blocks::register_standard_blocks(&mut registry, &|_| 0); // ids only matter for generation
    let mut json = String::from("{\n");
    for seed in [0u64, 0x13372137CAFEu64].iter() {
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
        json += ",\n";
    }
    // raw noise-crate samples (the point set is reproduced by tools/golden_compare.py: same integer recurrences)
    json += "  \"noise\": {\n";
    let seeds = [0u32, 1, 12345, 0xFFFF_FFFF];
    for (si, seed) in seeds.iter().enumerate() {
        let pt = PermutationTable::new(*seed);
        let ss = SuperSimplex::new(*seed);
        json += &format!("    \"{}\": {{\n      \"hash_i0\": [", seed);
        for i in 0..256isize {
            json += &format!("{}{}", if i > 0 { "," } else { "" }, pt.hash(&[i, 0]));
        }
        json += "],\n      \"hash_grid\": [";
        for k in 0..256isize {
            json += &format!("{}{}", if k > 0 { "," } else { "" }, pt.hash(&[(k % 16) * 17 - 100, (k / 16) * 13 - 90]));
        }
        json += "],\n      \"get_bits\": [";
        let mut st: u64 = 0x9E3779B97F4A7C15 ^ (*seed as u64);
        for k in 0..512 {
            // xorshift64* point set in [-512, 512)^2 with 1/1024 resolution (exact in f64)
            let mut nxt = || {
                st ^= st >> 12;
                st ^= st << 25;
                st ^= st >> 27;
                st.wrapping_mul(0x2545F4914F6CDD1D)
            };
            let x = ((nxt() >> 44) as i64 - (1 << 19)) as f64 / 1024.0;
            let y = ((nxt() >> 44) as i64 - (1 << 19)) as f64 / 1024.0;
            json += &format!("{}\"{:016x}\"", if k > 0 { "," } else { "" }, ss.get([x, y]).to_bits());
        }
        json += "]\n    }";
        json += if si + 1 < seeds.len() { ",\n" } else { "\n" };
    }
    json += "  }\n}\n";
    std::fs::write(out, json).unwrap();
}
