// tables.hpp — geometry lookup tables of the mesher, pre-rotated into world space per (shape, orientation) class.
//
// Source of the data: bxw_world/src/blocks/stdshapes.rs (VoxelShapeDef tables, :103-700) and
// bxw_util/src/direction.rs (OctahedralOrientation, :165-347). The reference looks a side up at run time as
// sides[orientation.unapply_to_dir(world_dir)] and rotates every vertex / AO offset by orientation.to_matrixi()
// (src/client/render/voxmesh.rs:105-106, 125-139); here all 24 rotations are applied once on the host so the
// kernels index [class][world_dir][vertex] directly.
#pragma once
#include <cstdint>

namespace bxw {

// class 0 = no mesh (VOXEL_NO_SHAPE); class 1 + shape*24 + orientation, shape 0 cube,1 slope,2 corner,3 inner corner
constexpr int kNumClasses = 97;

// ctab[class]: bits 0-5 CC (side facing world dir d can_clip), 6-11 EC (emits && can_be_clipped),
//              12-17 EU (emits && !can_be_clipped), 18-29 per-dir vertex-count code (2 bits: 0 none,1 three,2 four,3 six)
// stab[class][d]: bits 0-2 local side, 3-5 n_verts, 6-8 n_indices, 9-20 barycentric sign codes (2 bits/vertex: sign+1)
// vtab[class][d][j] (64-bit, host-side source of vdesc; the kernels read vdesc / aolut):
//   bits 0-2 corner (x,y,z in {0,1}: vertex = cell + corner), 3 texcoord u, 4 texcoord v, 5-7 barycentric >0.1 flags (x,y,z),
//   8-9 sign+1, 10-12 n_ao, 13+6k.. AO sample k: (dx+1) | (dy+1)<<2 | (dz+1)<<4
struct MeshTables {
    uint32_t ctab[kNumClasses];
    uint32_t stab[kNumClasses][6];
    unsigned long long vtab[kNumClasses][6][6];
    float ao_lut[8]; // ao after k occluding samples: sequential f32 multiplies by 0.88 (voxmesh.rs:27,129-137)
    // aotab[class][d][j]: the AO samples of vtab[class][d][j] as byte offsets into the kernel's 34^3 class table
    // (dy*plane + dz*row + dx), entries 0..4; entry 5 = sample count
    alignas(16) short aotab[kNumClasses][6][6][8];
    // aomask[class][d][j]: the same AO samples as a mask over the 3x3x3 neighbourhood, bit (dy+1)*9 + (dz+1)*3 + (dx+1)
    uint32_t aomask[kNumClasses][6][6];
    // compact per-vertex descriptor (what the emit phases read): bits 0-2 corner, 3 u, 4 v, 5-7 barycentric flags,
    // 8-12 code of the world-space AO normal n: (n.x+1) + 3 (n.y+1) + 9 (n.z+1)
    uint16_t vdesc[kNumClasses][6][6];
    // AO sample mask over the 3x3x3 neighbourhood for (corner, normal code): corner_ao_set (stdshapes.rs:131-152)
    uint32_t aolut[8 * 27];
    uint32_t aomask_ok; // 1 when no vertex lists the same AO sample twice (so popcount == number of occluders)
};

// shared-memory layout of the kernel's class table (bytes per (y,z) row / per y plane), baked into aotab
constexpr int kClsRowBytes = 40;
constexpr int kClsPlaneBytes = 34 * kClsRowBytes;

void build_mesh_tables(MeshTables &t);

// orientation helpers (direction.rs:267-293, 312-314), exposed for host-side use / tests
void orientation_matrix(int idx, int m[9]); // row-major
int orientation_unapply(int idx, int dir);

} // namespace bxw
