// octree_host.cpp — see octree_host.hpp.
#include "octree_host.hpp"

#include <cmath>

namespace rtb {
namespace {

struct Box {
    double mn[3], mx[3];
};
struct Tri {
    double a[3], b[3], c[3];
};

// BoundingBox::octant (src/geometry.rs:1067-1098): bit 2 of i = upper half in x, bit 1 = y, bit 0 = z
Box octant_of(const Box& b, int i) {
    Box o;
    for (int k = 0; k < 3; ++k) {
        const double c = (b.mn[k] + b.mx[k]) / 2.;
        const bool upper = (i >> (2 - k)) & 1;
        o.mn[k] = upper ? c : b.mn[k];
        o.mx[k] = upper ? b.mx[k] : c;
    }
    return o;
}

bool contains(const Box& b, const double* p) {   // :968-975, inclusive
    return b.mn[0] <= p[0] && p[0] <= b.mx[0] && b.mn[1] <= p[1] && p[1] <= b.mx[1] && b.mn[2] <= p[2] && p[2] <= b.mx[2];
}

// BoundingBox::intersect (:977-1036): the six faces in the order left, right, bottom, top, back, front; the FIRST face
// whose plane is met at t >= 1e-7 inside the face's rectangle wins (not the nearest)
bool box_intersect(const Box& b, const double* pos, const double* dir, double& t_out) {
    const double EPS = 0.0000001;
    for (int face = 0; face < 6; ++face) {
        const int ax = face >> 1;
        const double plane = (face & 1) ? b.mx[ax] : b.mn[ax];
        const double t = (plane - pos[ax]) / dir[ax];
        if (t >= EPS) {
            const int u = (ax + 1) % 3, v = (ax + 2) % 3;
            const double pu = pos[u] + t * dir[u], pv = pos[v] + t * dir[v];
            if (b.mn[u] <= pu && pu <= b.mx[u] && b.mn[v] <= pv && pv <= b.mx[v]) {
                t_out = t;
                return true;
            }
        }
    }
    return false;
}

bool intersect_line_segment(const Box& b, const double* p, const double* q) {   // :1038-1047
    double d[3] = {q[0] - p[0], q[1] - p[1], q[2] - p[2]};
    const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (double& x : d) x /= len;
    double t;
    return box_intersect(b, p, d, t) && t <= len;
}

bool overlaps_triangle(const Box& b, const Tri& t) {   // :1049-1061
    if (contains(b, t.a) || contains(b, t.b) || contains(b, t.c)) return true;
    return intersect_line_segment(b, t.a, t.b) || intersect_line_segment(b, t.a, t.c) || intersect_line_segment(b, t.b, t.c);
}

constexpr int MAX_DEPTH = 10;   // :1146
constexpr int SMALL_NODE = 9;   // :1147

int build(const std::vector<Tri>& tris, const Box& box, const std::vector<int32_t>& ids, HostOctree& out, int depth) {   // :1165-1216
    if (ids.empty()) return -1;
    const int me = (int)out.nodes.size();
    out.nodes.emplace_back();
    {
        HostOctreeNode& n = out.nodes[me];
        for (int k = 0; k < 3; ++k) { n.mn[k] = box.mn[k]; n.mx[k] = box.mx[k]; }
        for (int& c : n.child) c = -1;
        n.first = 0;
        n.count = -1;
    }
    if ((int)ids.size() <= SMALL_NODE || depth >= MAX_DEPTH) {
        out.nodes[me].first = (int)out.tri_refs.size();
        out.nodes[me].count = (int)ids.size();
        out.tri_refs.insert(out.tri_refs.end(), ids.begin(), ids.end());
        return me;
    }
    Box oct[8];
    for (int i = 0; i < 8; ++i) oct[i] = octant_of(box, i);
    std::vector<int32_t> sub[8];
    for (int32_t id : ids)
        for (int i = 0; i < 8; ++i)
            if (overlaps_triangle(oct[i], tris[(size_t)id])) sub[i].push_back(id);
    for (int i = 0; i < 8; ++i) {
        const int c = build(tris, oct[i], sub[i], out, depth + 1);
        out.nodes[me].child[i] = c;   // (re-indexed: the vector may have grown)
    }
    return me;
}

}  // namespace

void build_reference_octree(const HostObject& mesh, HostOctree& out) {
    out.nodes.clear();
    out.tri_refs.clear();
    const size_t n = mesh.indices.size() / 3;
    std::vector<Tri> tris(n);
    std::vector<int32_t> ids(n);
    for (size_t i = 0; i < n; ++i) {   // Mesh::triangle, src/geometry.rs:872-877
        const D3 &a = mesh.vertices[mesh.indices[3 * i]], &b = mesh.vertices[mesh.indices[3 * i + 1]], &c = mesh.vertices[mesh.indices[3 * i + 2]];
        tris[i] = Tri{{a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}};
        ids[i] = (int32_t)i;
    }
    const Box root{{mesh.bb_min.x, mesh.bb_min.y, mesh.bb_min.z}, {mesh.bb_max.x, mesh.bb_max.y, mesh.bb_max.z}};   // Mesh::bounding_box, scale() quirk included
    build(tris, root, ids, out, 1);
}

}  // namespace rtb
