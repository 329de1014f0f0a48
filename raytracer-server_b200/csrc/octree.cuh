// octree.cuh — RTB_ACCEL_OCTREE_REFERENCE: the reference's own Mesh::intersect on the device.
//
// The reference never runs its brute-force branch: SceneSpec::to_scene calls mesh.accelerate() for every mesh
// (src/scene.rs:430-432), so every mesh ray goes through Octree::intersect / _intersect_recurse
// (src/geometry.rs:1237-1295), which is NOT a nearest-hit query:
//   * no test against the root box; the children of a parent are visited in the order of the distance from the ray
//     ORIGIN to the centres of the ROOT's octants — the same order at every level (:1248-1260; stable insertion sort)
//   * a child is entered if BoundingBox::intersect reports any of the six faces met at t >= 1e-7 (:977-1036; used as a boolean)
//   * the FIRST child subtree that yields a hit ends the search (:1263-1273), and a leaf returns the nearest of ITS
//     triangles even when that hit lies outside the leaf's box (:1276-1293)
// so the triangle found can be farther away than another one in a later octant (~30 % of the mesh-origin
// secondary rays on flying_unicorn), and Scene::trace_ray (src/scene.rs:272-289) then compares THAT hit with the other
// objects.  This file reproduces exactly that in fp32, over the octree octree_host.cpp built in f64.
// Node record (32 B, 2 x float4; int bits in .w):   min.xyz | leaf: first reference, parent: index of its first child
//                                                   max.xyz | leaf: count >= 0,       parent: -1 - mask
// The children of a parent lie next to each other in octant order (engine.cu renumbers the host tree breadth-first), and bit i
// of `mask` says whether octant i has one: child i = first child + popcount(mask below bit i).  A search therefore carries
// (first child, mask) of the parent it is in and loads ONE record per child test — the first layout (64 B, eight explicit
// child indices) cost a second, dependent load per test.  Leaf triangles are references into the LBVH's triangle table (leaf
// order), so hit ids, self-intersection handling and shading are shared with the LBVH path.
#pragma once

#include "intersect.cuh"

namespace rtb {

constexpr int OCT_MAX_DEPTH = 10;   // Octree::MAX_DEPTH: the deepest node sits at depth 10, root = 1

// BoundingBox::intersect as a boolean (src/geometry.rs:977-1036): the reference walks the six faces (left, right, bottom, top,
// back, front) and accepts the first one whose plane is met at t >= 1e-7 inside the face's rectangle (inclusive).  Only
// is_some() is used (:1266), and "some face is met at t >= EPS" is the same statement as "the ray's parameter interval inside
// the box is non-empty and its far end is >= EPS" (the exit point always lies on a face) — the slab form below, a third of the
// instructions (octree traversal 4.67 -> 3.07 s on the bench frame; same hits on all 470 k probe rays).  The two forms can only
// disagree for rays that graze an edge within rounding, where the f64 reference and any fp32 evaluation differ as well.
// inv = 1 / d (IEEE: +-inf for a zero component: a ray parallel to a slab passes it iff its origin lies inside, as the
// reference's face tests do).  Deriving the child boxes from the parent's planes instead of loading them was measured too:
// fewer loads, more registers, 6 % slower — the kernel is bound by divergent instruction issue, not by its loads.
__device__ __forceinline__ bool oct_box_hit(const float4 mn, const float4 mx, float3 o, float3 d, float3 inv) {
    constexpr float EPS = 1e-7f;
    (void)d;
    const float x0 = (mn.x - o.x) * inv.x, x1 = (mx.x - o.x) * inv.x;
    const float y0 = (mn.y - o.y) * inv.y, y1 = (mx.y - o.y) * inv.y;
    const float z0 = (mn.z - o.z) * inv.z, z1 = (mx.z - o.z) * inv.z;
    const float t_in = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
    const float t_out = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    return t_in <= t_out && t_out >= EPS;
}

// Triangle::intersect (src/geometry.rs:637-670) on entry k of the octree's own triangle table (a copy of the LBVH's record per
// leaf reference, naming its slot there: one fetch per triangle instead of reference -> record).  Same arithmetic as trav_leaf
// (intersect.cuh); no upper bound on t.
__device__ __forceinline__ bool oct_tri_test(const DevScene& S, int k, float3 o, float3 d, uint32_t origin, float& t_out, uint32_t& id_out) {
    const float4* tp = S.oct_tris + (size_t)k * TRI_STRIDE;
    float4 t0, t1;
    ldg256(tp, t0, t1);
    const float4 t2 = __ldg(tp + 2);
    const uint32_t s = __float_as_uint(t1.w);
    const float3 e1 = f3(t1), e2 = f3(t2);
    const float3 pvec = cross(d, e2);
    const float det = dot(e1, pvec);
    const float nd = det * t0.w;
    if (fabsf(nd) < DN_EPS) return false;
    const float inv = __fdividef(1.0f, det);
    const float3 tvec = o - f3(t0);
    const float u = dot(tvec, pvec) * inv;
    const float3 qvec = cross(tvec, e1);
    const float vv = dot(d, qvec) * inv;
    float t = dot(e2, qvec) * inv;
    if (TRI_BASE + s == (origin & PC_ID_MASK)) {   // the triangle the ray starts on: as the reference's f64 arithmetic sees it (intersect.cuh header)
        const float dnf = (origin & PC_FLIPPED) ? -nd : nd;
        t = -SURF_OFFSET / dnf;
    }
    if (u < 0.0f || vv < 0.0f || u + vv > 1.0f || !(t > T_EPS)) return false;
    t_out = t;
    id_out = TRI_BASE + s;
    return true;
}

// Node::Leaf: the nearest of the leaf's triangles (src/geometry.rs:1276-1293)
__device__ __forceinline__ bool oct_leaf(const DevScene& S, int first, int count, float3 o, float3 d, uint32_t origin, float& t_out, uint32_t& id_out) {
    float best = INFINITY;
    uint32_t best_id = PC_NONE;
    for (int k = first; k < first + count; ++k) {
        float t;
        uint32_t id;
        if (oct_tri_test(S, k, o, d, origin, t, id) && t < best) { best = t; best_id = id; }   // strict '<': the first of equal hits stays (:1281)
    }
    t_out = best;
    id_out = best_id;
    return best_id != PC_NONE;
}

// octant_search_order (src/geometry.rs:1245-1260): insertion sort of 0..7 by the distance from ray.pos to the centre of the ROOT's
// octant, ascending, stable — a rank computation gives the same order (ties: lower index first).  3 bits per position.
__device__ __forceinline__ uint32_t oct_search_order(const float4 rmn, const float4 rmx, float3 o) {
    const float cx = 0.5f * (rmn.x + rmx.x), cy = 0.5f * (rmn.y + rmx.y), cz = 0.5f * (rmn.z + rmx.z);
    float dist[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float ox = 0.5f * ((i & 4) ? cx + rmx.x : rmn.x + cx) - o.x;
        const float oy = 0.5f * ((i & 2) ? cy + rmx.y : rmn.y + cy) - o.y;
        const float oz = 0.5f * ((i & 1) ? cz + rmx.z : rmn.z + cz) - o.z;
        dist[i] = sqrtf(ox * ox + oy * oy + oz * oz);   // Vec3::mag: the comparison is made on the rounded root
    }
    uint32_t order = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int rank = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) rank += (dist[j] < dist[i] || (dist[j] == dist[i] && j < i)) ? 1 : 0;
        order |= (uint32_t)i << (3 * rank);
    }
    return order;
}

// one byte per OCTANT: 1 << (its position in the search order); .x = octants 0..3, .y = octants 4..7
__device__ __forceinline__ uint2 oct_order_bytes(uint32_t order) {
    uint2 B = make_uint2(0u, 0u);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const uint32_t i = (order >> (3 * p)) & 7u;
        const uint32_t v = (1u << p) << (8u * (i & 3u));
        if (i & 4u) B.y |= v; else B.x |= v;
    }
    return B;
}
// the children a parent has (bit i of mask = octant i), as bits in search order: every mask bit is spread to a byte
// (nibble * 0x00204081 puts bit k at 8k without carries), selects its octant's byte of B, and the bytes are OR-ed together
__device__ __forceinline__ uint32_t oct_present_in_order(int mask, uint2 B) {
    const uint32_t mlo = ((((uint32_t)mask & 15u) * 0x00204081u) & 0x01010101u) * 0xffu;
    const uint32_t mhi = (((((uint32_t)mask >> 4) & 15u) * 0x00204081u) & 0x01010101u) * 0xffu;
    uint32_t x = (B.x & mlo) | (B.y & mhi);
    x |= x >> 16;
    x |= x >> 8;
    return x & 0xffu;
}

// Octree::intersect for ONE mesh (root node index `root`): true + (t, id) of the hit the reference would return
__device__ __forceinline__ bool oct_intersect(const DevScene& S, int root, float3 o, float3 d, float3 inv, uint32_t origin, float& t_out, uint32_t& id_out,
                                              uint32_t* work) {
    const float4* N = S.oct_nodes;
    const float4 rmn = __ldg(N + (size_t)root * 2), rmx = __ldg(N + (size_t)root * 2 + 1);
    if (__float_as_int(rmx.w) >= 0) {   // the root is a leaf (a mesh of at most SMALL_NODE triangles)
        if (work) work[1] += (uint32_t)__float_as_int(rmx.w);
        return oct_leaf(S, __float_as_int(rmn.w), __float_as_int(rmx.w), o, d, origin, t_out, id_out);
    }
    const uint32_t order = oct_search_order(rmn, rmx, o);
    int sbase[OCT_MAX_DEPTH], smask[OCT_MAX_DEPTH];   // (first child, mask) of the parents above the current one
    int cur_base = __float_as_int(rmn.w), cur_mask = -1 - __float_as_int(rmx.w);
    unsigned long long pos = 0;   // 4 bits per level: next position in `order` to try
    int level = 0;
    for (;;) {
        const unsigned p = (unsigned)(pos >> (4 * level)) & 15u;
        if (p == 8u) {   // this parent is exhausted without a hit
            if (level == 0) return false;
            pos &= ~(15ull << (4 * level));
            --level;
            cur_base = sbase[level];
            cur_mask = smask[level];
            continue;
        }
        pos += 1ull << (4 * level);
        const int i = (int)(order >> (3 * p)) & 7;
        if (!((cur_mask >> i) & 1)) continue;   // children[i] == None
        const int c = cur_base + __popc((unsigned)cur_mask & ((1u << i) - 1u));
        const float4 cmn = __ldg(N + (size_t)c * 2), cmx = __ldg(N + (size_t)c * 2 + 1);
        if (work) work[0]++;
        if (!oct_box_hit(cmn, cmx, o, d, inv)) continue;
        const int cnt = __float_as_int(cmx.w);
        if (cnt >= 0) {   // leaf: its nearest triangle ends the whole search, wherever the hit lies
            if (work) work[1] += (uint32_t)cnt;
            if (oct_leaf(S, __float_as_int(cmn.w), cnt, o, d, origin, t_out, id_out)) return true;
        } else if (level + 1 < OCT_MAX_DEPTH) {
            sbase[level] = cur_base;
            smask[level] = cur_mask;
            ++level;
            cur_base = __float_as_int(cmn.w);
            cur_mask = -1 - cnt;
        }
    }
}

// Scene::trace_ray's mesh part under the reference's octrees: every mesh in object order, strict '<' against the best so
// far (src/scene.rs:277-284).  best_t comes in as the nearest analytic hit (or a shadow ray's limit).
__device__ __forceinline__ void oct_trace_meshes(const DevScene& S, float3 o, float3 d, uint32_t origin, float& best_t, uint32_t& best_id, uint32_t* work) {
    const float3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    for (int m = 0; m < S.n_oct_meshes; ++m) {
        const int root = __ldg(S.oct_roots + m);
        if (root < 0) continue;
        float t;
        uint32_t id;
        if (oct_intersect(S, root, o, d, inv, origin, t, id, work) && t < best_t) {
            best_t = t;
            best_id = id;
        }
    }
}

}  // namespace rtb
