// cb200_blob.h -- layout of the packed robot-constant blob shared by the host packer and the kernels.
//
// All robot constants the fused kernel needs (KinematicsParams + SelfCollisionKinematicsCfg of the
// reference, curobo/_src/robot/types/kinematics_params.py:23-158, self_collision_params.py:16-125,
// plus tables derived from them once on the host) sit in ONE contiguous device buffer:
//     [ header | sections staged to shared memory ............ | self-collision pair list ]
//       <------------- smem_bytes (one cp.async.bulk per CTA) ->
// Offsets are in bytes from the blob start; every section is 16-byte aligned.
#pragma once
#include <stdint.h>

namespace cb200 {

constexpr int32_t kBlobMagic = 0x30324243;  // "CB20"
constexpr int kMaxLinks = 64;               // ancestor masks are one 64-bit word per link

struct BlobHeader {
  int32_t magic;
  int32_t total_bytes;
  int32_t smem_bytes;  // prefix (incl. this header) copied to shared memory
  int32_t nl, D, S, L, P;
  int32_t n_levels;
  int32_t off_fixed;        // float [nl*12]   fixed transforms, row-major 3x4
  int32_t off_joff;         // float [nl*2]    joint (scale, bias)
  int32_t off_link_map;     // int16 [nl]      parent link
  int32_t off_joint_map;    // int16 [nl]      joint index or -1
  int32_t off_joint_type;   // int8  [nl]
  int32_t off_tool_map;     // int16 [L]
  int32_t off_spheres;      // float4[S]       link-frame spheres (x,y,z,r)
  int32_t off_sph_link;     // int16 [S]
  int32_t off_padding;      // float [S]       self-collision radius padding
  int32_t off_link_sph_off; // int16 [nl+1]    CSR link -> spheres
  int32_t off_link_sph_idx; // int16 [S]
  int32_t off_level_off;    // int16 [n_levels+1]  CSR depth level -> links (level 0 = base)
  int32_t off_level_links;  // int16 [nl]
  int32_t off_anc_mask;     // uint64[nl]      bit j set <=> link j is an ancestor of (or is) this link
  int32_t off_jl_off;       // int16 [D+1]     CSR joint -> links it drives (mimic joints share)
  int32_t off_jl_idx;       // int16 [n joint links]
  int32_t off_limits;       // float [10*D]    pos lo|hi, vel lo|hi, acc lo|hi, jerk lo|hi, effort lo|hi
  int32_t off_pairs;        // int16 [P*2]     (i<j) -- NOT staged to shared memory
  // self-collision broad phase (the pair list is the union of link x link sphere blocks): collision links,
  // their contiguous sphere ranges, link-frame bounding spheres and the list of checked link pairs.
  // n_lp == 0 => pair list is not a union of full link blocks: kernels fall back to the explicit pair list.
  int32_t n_cl, n_lp;
  int32_t off_cl_link;      // int16 [n_cl]    link index of collision link a
  int32_t off_cl_start;     // int16 [n_cl+1]  sphere range of collision link a
  int32_t off_cl_bound;     // float4[n_cl]    bounding sphere (x,y,z,R) in the link frame; R < 0: no enabled sphere
  int32_t off_lp;           // uint32[n_lp]    (a | b << 16), a < b, indices into the collision-link arrays
  // FK compose schedule: one word per step = two (link, parent) byte pairs (0xFF = idle slot); links of one
  // depth level are independent, so a step composes two of them with 12 lanes each.
  int32_t n_fk_steps;
  int32_t off_fk_sched;     // uint32[2*n_fk_steps]  per (step, slot): byte offset of the link's 3x4 (lo16, 0xFFFF = idle) | parent's (hi16)
  int32_t off_cl_bound_scene;  // float4[n_cl]  like cl_bound but over spheres with r >= 0 and unpadded radii (scene broad phase)
  int32_t off_sph_cl;       // uint8 [S]  collision-link index of each sphere (valid when n_lp > 0)
  int32_t n_sphere_cfgs;    // link-sphere configurations the broad-phase bounds cover (>= 1); config 0 is the staged set
  int32_t reserved[10];
};
static_assert(sizeof(BlobHeader) == 192, "BlobHeader must be 192 bytes");

}  // namespace cb200
