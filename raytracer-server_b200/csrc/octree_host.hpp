// octree_host.hpp — the reference's per-mesh octree, built on the host for RTB_ACCEL_OCTREE_REFERENCE.
//
// Restates Octree::build / _build (reference src/geometry.rs:1149-1216) and the triangle-to-octant rule it uses,
// BoundingBox::overlaps_triangle / contains / intersect_line_segment / intersect / octant (src/geometry.rs:968-1112), in
// f64 like the reference: MAX_DEPTH 10, SMALL_NODE 9, root at depth 1, triangles copied into every octant they
// overlap (any vertex inside, or any edge hitting the box — the "box pierced by the triangle's interior" case is
// missed, as in the reference).  Traversal happens on the device (octree.cuh).
#pragma once

#include <cstdint>
#include <vector>

#include "scene_host.hpp"

namespace rtb {

struct HostOctreeNode {
    double mn[3], mx[3];   // the node's box (root: Mesh::bounding_box; child: BoundingBox::octant of its parent)
    int child[8];          // Node::Parent: node index per octant, -1 = None
    int first, count;      // Node::Leaf: triangles [first, first + count) of tri_refs; count = -1 for a parent
};

struct HostOctree {
    std::vector<HostOctreeNode> nodes;   // nodes[0] is the root (empty: the mesh has no triangles)
    std::vector<int32_t> tri_refs;       // triangle indices inside the mesh, in the order of the reference's Vec
};

void build_reference_octree(const HostObject& mesh, HostOctree& out);

}  // namespace rtb
