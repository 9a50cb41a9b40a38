// oracle/mot_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  This file is the CPU *oracle* for the B200 hot path: a plain,
// single-threaded restatement of what the reference tracker computes per frame.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The
// product library (libmot_b200.so) never links, loads or calls anything in oracle/.
//
// PARITY STATUS, per piece:
//   pinned to the reference's own code -- removeStatic + yaw (orc_remove_static), getCentroid (orc_get_centroid), the
//     IHGP constants and callIHGP (orc_ihgp_setup / orc_ihgp_step), and, in tracker_ref.py, the association / track
//     lifecycle.  The reference ships no tests or golden vectors for this path, and its build needs
//     ROS, PCL and Eigen (none installed); but its tracker and IHGP sources do compile, where they lie under
//     /root/reference, against the stand-in headers of oracle/shim (oracle/Makefile target _ref, oracle/ref_harness.cpp).
//     tests/test_ref_pin.py runs that build side by side with this file, and tests/golden/ref_vectors.npz holds its
//     outputs (generator: tests/golden/make_ref_fixtures.py) for the places where /root/reference does not exist.
//   "parity unpinned" -- the pieces that live in PCL, a third-party dependency outside the reference tree:
//     EuclideanClusterExtraction over the FLANN KD-tree (orc_cluster_kdtree, orc_labels_*), VoxelGrid (orc_voxel_grid)
//     and fromROSMsg (oracle.py).  Their published algorithms are restated and anchored on the reference's call sites;
//     they are cross-checked against an independent O(N^2) brute-force partition using the identical fp32 predicate and,
//     in tests, against scipy (cKDTree + connected components).  Inside oracle/_ref these same restatements stand in
//     for PCL, so that build says nothing about them.
//   The one reference-owned data fixture, map/sim_01.pgm, is used for removeStatic.
//
// Citations are relative to /root/reference:
//   MOT.cpp = src/multiple_object_tracking_lidar.cpp
//   IHGP.cpp = src/ihgp/InfiniteHorizonGP.cpp,  M32.cpp = src/ihgp/Matern32model.cpp
// Clustering lives in un-vendored third-party code (PCL EuclideanClusterExtraction over FLANN
// KDTreeSingleIndex, versions unpinned by package.xml:11-12); its published algorithm is restated here
// and anchored on the reference call site MOT.cpp:472-488.
//
// Build: g++ -O3 -ffp-contract=off -shared -fPIC (see oracle/Makefile; no -march so the prebuilt .so runs on
// whatever host CPU the GPU box has, and matches the reference's flagless x86-64 build).  -ffp-contract=off
// matters: the reference is built without -march flags (CMakeLists.txt:4), so no FMA contraction happens
// in its float/double expressions; the oracle must round every product and sum separately too.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

namespace {

struct P4 { float x, y, z, w; };  // pcl::PointXYZ is float[4] (x, y, z, pad)

// ----------------------------------------------------------------------------------------------------
// removeStatic  (MOT.cpp:664-706, yaw from MOT.cpp:1013-1023, map layout from MOT.cpp:235-251)
// ----------------------------------------------------------------------------------------------------

// MOT.cpp:1013-1023 -- double atan2, returned through a float.
float yaw_from_quat(const double q[4] /* x y z w */) {
    double siny_cosp = 2 * (q[3] * q[2] + q[0] * q[1]);
    double cosy_cosp = 1 - 2 * (q[1] * q[1] + q[2] * q[2]);
    double yaw = std::atan2(siny_cosp, cosy_cosp);
    return (float)yaw;
}

inline bool cell_blocked(const int8_t* occ, int W, int H, long r, long c) {
    // Policy for the reference's unchecked index (MOT.cpp:686): out-of-map == unknown == blocked.
    if (r < 0 || c < 0 || r >= H || c >= W) return true;
    int v = occ[r * (long)W + c];
    return v > 50 || v == -1;  // MOT.cpp:686
}

}  // namespace

extern "C" {

float orc_yaw_from_quat(const double* q_xyzw) { return yaw_from_quat(q_xyzw); }

// keep[i] = 1 iff point i survives removeStatic.  Returns the number kept.
// Per point, exactly as MOT.cpp:674-678: float x_map/y_map from a double subtraction, float yaw, float
// cos/sin (std::cos(float) is the float overload), float divide by the float32 resolution, C truncation.
int64_t orc_remove_static(const float* xyz16, int64_t n, const int8_t* occ, int W, int H, float resolution,
                          double origin_x, double origin_y, const double* q_xyzw, int tol, uint8_t* keep) {
    if (tol > 4) tol = 4; else if (tol < 0) tol = 0;  // MOT.cpp:95-96
    const P4* pts = reinterpret_cast<const P4*>(xyz16);
    int64_t kept = 0;
    for (int64_t p = 0; p < n; ++p) {
        float x_map = (float)((double)pts[p].x - origin_x);         // MOT.cpp:674
        float y_map = (float)((double)pts[p].y - origin_y);         // MOT.cpp:675
        float theta = yaw_from_quat(q_xyzw);                        // MOT.cpp:676 (per point, as written)
        float cs = std::cos(-theta), sn = std::sin(-theta);         // float overloads
        float fc = (cs * x_map - sn * y_map) / resolution;          // MOT.cpp:677
        float fr = (sn * x_map + cs * y_map) / resolution;          // MOT.cpp:678
        bool ok = std::fabs(fc) < 1073741824.0f && std::fabs(fr) < 1073741824.0f;  // NaN/huge -> drop (UB in ref)
        if (ok) {
            int col = (int)fc, row = (int)fr;                       // truncation toward zero
            for (int i = -tol; i <= tol && ok; ++i)                 // MOT.cpp:681-702
                for (int j = -tol; j <= tol; ++j)
                    if (cell_blocked(occ, W, H, (long)row + i, (long)col + j)) { ok = false; break; }
        }
        keep[p] = ok ? 1 : 0;
        kept += ok;
    }
    return kept;
}

}  // extern "C"

// ----------------------------------------------------------------------------------------------------
// Euclidean clustering (call site MOT.cpp:472-488; PCL extractEuclideanClusters + FLANN restated)
// ----------------------------------------------------------------------------------------------------
namespace {

// FLANN L2_Simple<float>: result = 0; for d: diff = a[d]-b[d]; result += diff*diff   (fp32, dim 3)
inline float d2_fp32(const P4& a, const P4& b) {
    float r = 0.0f;
    float d = a.x - b.x; r += d * d;
    d = a.y - b.y; r += d * d;
    d = a.z - b.z; r += d * d;
    return r;
}

// KdTreeFLANN::radiusSearch(point, double radius): r2 = (float)(radius*radius) with radius = (double)tol_f
inline float r2_from_tol(float tol) { double r = (double)tol; return (float)(r * r); }

// Restatement of FLANN KDTreeSingleIndex (leaf_max_size 15, reorder=true) -- enough of it to reproduce the
// result set and the cost structure (one tree descent per query, results sorted by distance).
struct KdTree {
    struct Node { int left, right; int child1, child2; int cutfeat; float divlow, divhigh; };
    struct Interval { float low, high; };
    std::vector<Node> nodes;
    std::vector<int> vind;
    std::vector<P4> data;  // reordered copy (reorder_ = true)
    Interval root_bbox[3];
    int leaf_max = 15;
    int root = -1;

    static float coord(const P4& p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }

    void build(const P4* pts, int n) {
        nodes.clear();
        vind.resize(n);
        std::iota(vind.begin(), vind.end(), 0);
        if (n == 0) { root = -1; return; }
        for (int d = 0; d < 3; ++d) { root_bbox[d].low = root_bbox[d].high = coord(pts[0], d); }
        for (int i = 1; i < n; ++i)
            for (int d = 0; d < 3; ++d) {
                float v = coord(pts[i], d);
                if (v < root_bbox[d].low) root_bbox[d].low = v;
                if (v > root_bbox[d].high) root_bbox[d].high = v;
            }
        nodes.reserve(2 * (n / leaf_max + 2));
        Interval bb[3] = {root_bbox[0], root_bbox[1], root_bbox[2]};
        root = divide(pts, 0, n, bb);
        data.resize(n);
        for (int i = 0; i < n; ++i) data[i] = pts[vind[i]];
    }

    void minmax(const P4* pts, int* ind, int count, int d, float& mn, float& mx) {
        mn = mx = coord(pts[ind[0]], d);
        for (int i = 1; i < count; ++i) {
            float v = coord(pts[ind[i]], d);
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
    }

    void plane_split(const P4* pts, int* ind, int count, int cutfeat, float cutval, int& lim1, int& lim2) {
        int left = 0, right = count - 1;
        for (;;) {
            while (left <= right && coord(pts[ind[left]], cutfeat) < cutval) ++left;
            while (left <= right && coord(pts[ind[right]], cutfeat) >= cutval) --right;
            if (left > right) break;
            std::swap(ind[left], ind[right]); ++left; --right;
        }
        lim1 = left;
        right = count - 1;
        for (;;) {
            while (left <= right && coord(pts[ind[left]], cutfeat) <= cutval) ++left;
            while (left <= right && coord(pts[ind[right]], cutfeat) > cutval) --right;
            if (left > right) break;
            std::swap(ind[left], ind[right]); ++left; --right;
        }
        lim2 = left;
    }

    int divide(const P4* pts, int left, int right, Interval* bbox) {
        int id = (int)nodes.size();
        nodes.push_back(Node{});
        int count = right - left;
        if (count <= leaf_max) {
            nodes[id].child1 = nodes[id].child2 = -1;
            nodes[id].left = left; nodes[id].right = right;
            for (int d = 0; d < 3; ++d) { bbox[d].low = bbox[d].high = coord(pts[vind[left]], d); }
            for (int k = left + 1; k < right; ++k)
                for (int d = 0; d < 3; ++d) {
                    float v = coord(pts[vind[k]], d);
                    if (v < bbox[d].low) bbox[d].low = v;
                    if (v > bbox[d].high) bbox[d].high = v;
                }
            return id;
        }
        // middleSplit_
        int* ind = &vind[left];
        const float EPS = 0.00001f;
        float max_span = bbox[0].high - bbox[0].low;
        for (int d = 1; d < 3; ++d) max_span = std::max(max_span, bbox[d].high - bbox[d].low);
        float max_spread = -1; int cutfeat = 0;
        for (int d = 0; d < 3; ++d) {
            float span = bbox[d].high - bbox[d].low;
            if (span > (1 - EPS) * max_span) {
                float mn, mx; minmax(pts, ind, count, d, mn, mx);
                float spread = mx - mn;
                if (spread > max_spread) { cutfeat = d; max_spread = spread; }
            }
        }
        float split_val = (bbox[cutfeat].low + bbox[cutfeat].high) / 2;
        float mn, mx; minmax(pts, ind, count, cutfeat, mn, mx);
        float cutval = split_val < mn ? mn : (split_val > mx ? mx : split_val);
        int lim1, lim2; plane_split(pts, ind, count, cutfeat, cutval, lim1, lim2);
        int idx = lim1 > count / 2 ? lim1 : (lim2 < count / 2 ? lim2 : count / 2);

        Interval lb[3] = {bbox[0], bbox[1], bbox[2]};
        lb[cutfeat].high = cutval;
        int c1 = divide(pts, left, left + idx, lb);
        Interval rb[3] = {bbox[0], bbox[1], bbox[2]};
        rb[cutfeat].low = cutval;
        int c2 = divide(pts, left + idx, right, rb);
        nodes[id].child1 = c1; nodes[id].child2 = c2; nodes[id].cutfeat = cutfeat;
        nodes[id].divlow = lb[cutfeat].high; nodes[id].divhigh = rb[cutfeat].low;
        for (int d = 0; d < 3; ++d) {
            bbox[d].low = std::min(lb[d].low, rb[d].low);
            bbox[d].high = std::max(lb[d].high, rb[d].high);
        }
        return id;
    }

    struct Hit { float d; int idx; };

    void search_level(const P4& q, int node, float mindistsq, float* dists, float r2, std::vector<Hit>& out) const {
        const Node& nd = nodes[node];
        if (nd.child1 < 0) {
            for (int i = nd.left; i < nd.right; ++i) {
                float d = d2_fp32(q, data[i]);
                if (d < r2) out.push_back(Hit{d, vind[i]});  // strict <  (RadiusResultSet::addPoint)
            }
            return;
        }
        int idx = nd.cutfeat;
        float val = coord(q, idx);
        float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
        int best, other; float cut_dist;
        if ((diff1 + diff2) < 0) { best = nd.child1; other = nd.child2; float t = val - nd.divhigh; cut_dist = t * t; }
        else { best = nd.child2; other = nd.child1; float t = val - nd.divlow; cut_dist = t * t; }
        search_level(q, best, mindistsq, dists, r2, out);
        float dst = dists[idx];
        mindistsq = mindistsq + cut_dist - dst;
        dists[idx] = cut_dist;
        if (mindistsq <= r2) search_level(q, other, mindistsq, dists, r2, out);
        dists[idx] = dst;
    }

    // Sorted radius search (pcl::search::KdTree default sorted_results_ = true).
    void radius_search(const P4& q, float r2, std::vector<Hit>& out) const {
        out.clear();
        if (root < 0) return;
        float dists[3] = {0, 0, 0};
        float distsq = 0;
        for (int d = 0; d < 3; ++d) {
            float v = coord(q, d);
            if (v < root_bbox[d].low) { float t = v - root_bbox[d].low; dists[d] = t * t; distsq += dists[d]; }
            if (v > root_bbox[d].high) { float t = v - root_bbox[d].high; dists[d] = t * t; distsq += dists[d]; }
        }
        search_level(q, root, distsq, dists, r2, out);
        std::sort(out.begin(), out.end(), [](const Hit& a, const Hit& b) { return a.d < b.d || (a.d == b.d && a.idx < b.idx); });
    }
};

struct ClusterOut {
    std::vector<std::vector<int>> clusters;  // canonical order: size desc, then min index asc
};

// Canonical ordering used by both the oracle and the GPU library: std::sort in PCL's extract() leaves
// ties between equal-size clusters unspecified, so we pin them by min index.
void canonical_order(std::vector<std::vector<int>>& cl) {
    std::stable_sort(cl.begin(), cl.end(), [](const std::vector<int>& a, const std::vector<int>& b) {
        if (a.size() != b.size()) return a.size() > b.size();
        return a.front() < b.front();
    });
}

// PCL extractEuclideanClusters restated: serial BFS, one sorted radius search per point, nn_start_idx = 1,
// size filter [min,max] inclusive (oversized components dropped whole), indices sorted ascending.
// PCL's extract() re-runs tree->setInputCloud after the caller already built the tree (MOT.cpp:473), so
// the KD-tree is built twice when build_twice != 0 (cost fidelity for the CPU baseline).
void cluster_kdtree(const P4* pts, int m, float tol, int min_size, int max_size, int build_twice, ClusterOut& out) {
    out.clusters.clear();
    if (m == 0) return;
    KdTree tree;
    tree.build(pts, m);
    if (build_twice) tree.build(pts, m);
    const float r2 = r2_from_tol(tol);
    std::vector<char> processed(m, 0);
    std::vector<KdTree::Hit> nn;
    std::vector<int> queue;
    for (int i = 0; i < m; ++i) {
        if (processed[i]) continue;
        queue.clear();
        queue.push_back(i);
        processed[i] = 1;
        size_t sq = 0;
        while (sq < queue.size()) {
            tree.radius_search(pts[queue[sq]], r2, nn);
            if (nn.empty()) { ++sq; continue; }
            for (size_t j = 1; j < nn.size(); ++j) {  // nn_start_idx = 1: first sorted hit is the query itself
                int k = nn[j].idx;
                if (processed[k]) continue;
                queue.push_back(k);
                processed[k] = 1;
            }
            ++sq;
        }
        if ((int)queue.size() >= min_size && (int)queue.size() <= max_size) {
            std::vector<int> r(queue);
            std::sort(r.begin(), r.end());
            r.erase(std::unique(r.begin(), r.end()), r.end());
            out.clusters.push_back(std::move(r));
        }
    }
    canonical_order(out.clusters);
}

struct DSU {
    std::vector<int> p;
    explicit DSU(int n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int find(int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(int a, int b) { a = find(a); b = find(b); if (a == b) return; if (a < b) p[b] = a; else p[a] = b; }
};

}  // namespace

extern "C" {

// label[i] = smallest point index in i's connected component under  d2_fp32(i,j) < r2  (no size filter).
// Independent checker: O(N^2) all pairs, no spatial structure at all.
void orc_labels_bruteforce(const float* xyz16, int m, float tol, int32_t* label) {
    const P4* pts = reinterpret_cast<const P4*>(xyz16);
    const float r2 = r2_from_tol(tol);
    DSU d(m);
    for (int i = 0; i < m; ++i)
        for (int j = i + 1; j < m; ++j)
            if (d2_fp32(pts[i], pts[j]) < r2) d.unite(i, j);
    for (int i = 0; i < m; ++i) label[i] = d.find(i);  // unite() always keeps the smaller root
}

// Same labels via a uniform grid (cell = 1.001*tol, exact predicate on the 27-neighbourhood): a second
// independent implementation that scales to millions of points for full-size parity runs.
void orc_labels_grid(const float* xyz16, int m, float tol, int32_t* label) {
    const P4* pts = reinterpret_cast<const P4*>(xyz16);
    if (m == 0) return;
    const float r2 = r2_from_tol(tol);
    const double cell = (double)tol * 1.001;
    double mn[3] = {pts[0].x, pts[0].y, pts[0].z};
    for (int i = 1; i < m; ++i) { mn[0] = std::min<double>(mn[0], pts[i].x); mn[1] = std::min<double>(mn[1], pts[i].y); mn[2] = std::min<double>(mn[2], pts[i].z); }
    std::vector<int64_t> cx(m), cy(m), cz(m);
    std::vector<uint64_t> key(m);
    auto mix = [](int64_t a, int64_t b, int64_t c) { return ((uint64_t)a << 42) ^ ((uint64_t)b << 21) ^ (uint64_t)c; };
    for (int i = 0; i < m; ++i) {
        cx[i] = (int64_t)std::floor((pts[i].x - mn[0]) / cell);
        cy[i] = (int64_t)std::floor((pts[i].y - mn[1]) / cell);
        cz[i] = (int64_t)std::floor((pts[i].z - mn[2]) / cell);
        key[i] = mix(cx[i] + 1, cy[i] + 1, cz[i] + 1);  // +1 keeps neighbour coords non-negative (<2^21 each)
    }
    std::vector<int> order(m);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b] || (key[a] == key[b] && a < b); });
    std::vector<uint64_t> ukey; std::vector<int> ustart;
    for (int s = 0; s < m; ++s)
        if (s == 0 || key[order[s]] != key[order[s - 1]]) { ukey.push_back(key[order[s]]); ustart.push_back(s); }
    ustart.push_back(m);
    DSU d(m);
    for (size_t c = 0; c < ukey.size(); ++c) {
        int a0 = ustart[c], a1 = ustart[c + 1];
        int i0 = order[a0];
        for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
            uint64_t nk = mix(cx[i0] + 1 + dx, cy[i0] + 1 + dy, cz[i0] + 1 + dz);
            if (nk < ukey[c]) continue;  // each unordered cell pair once
            auto it = std::lower_bound(ukey.begin(), ukey.end(), nk);
            if (it == ukey.end() || *it != nk) continue;
            size_t nc = it - ukey.begin();
            int b0 = ustart[nc], b1 = ustart[nc + 1];
            for (int a = a0; a < a1; ++a)
                for (int b = (nc == c ? a + 1 : b0); b < b1; ++b)
                    if (d2_fp32(pts[order[a]], pts[order[b]]) < r2) d.unite(order[a], order[b]);
        }
    }
    for (int i = 0; i < m; ++i) label[i] = d.find(i);
}

// The reference path (KD-tree + BFS).  Outputs CSR: offsets[K+1], indices[sum], returns K.
// Pass indices_cap >= m and offsets_cap >= m+1 to be safe.
int orc_cluster_kdtree(const float* xyz16, int m, float tol, int min_size, int max_size, int build_twice,
                       int32_t* offsets, int32_t* indices) {
    ClusterOut out;
    cluster_kdtree(reinterpret_cast<const P4*>(xyz16), m, tol, min_size, max_size, build_twice, out);
    int k = 0, pos = 0;
    offsets[0] = 0;
    for (auto& c : out.clusters) {
        std::memcpy(indices + pos, c.data(), c.size() * sizeof(int));
        pos += (int)c.size();
        offsets[++k] = pos;
    }
    return k;
}

// CSR from component labels with the same filter + canonical order (used to turn brute-force / grid labels
// into the boundary format so every representation can be compared bit-exactly).
int orc_csr_from_labels(const int32_t* label, int m, int min_size, int max_size, int32_t* offsets, int32_t* indices) {
    std::vector<int> size(m, 0);
    for (int i = 0; i < m; ++i) size[label[i]]++;
    std::vector<std::vector<int>> cl;
    std::vector<int> slot(m, -1);
    for (int i = 0; i < m; ++i) {
        int r = label[i];
        if (size[r] < min_size || size[r] > max_size) continue;
        if (slot[r] < 0) { slot[r] = (int)cl.size(); cl.emplace_back(); cl.back().reserve(size[r]); }
        cl[slot[r]].push_back(i);
    }
    canonical_order(cl);
    int k = 0, pos = 0;
    offsets[0] = 0;
    for (auto& c : cl) {
        std::memcpy(indices + pos, c.data(), c.size() * sizeof(int));
        pos += (int)c.size();
        offsets[++k] = pos;
    }
    return k;
}

// ----------------------------------------------------------------------------------------------------
// getCentroid  (MOT.cpp:708-822, euc_dist MOT.cpp:1025-1028): farthest pair -> farthest point from the XY
// line through it -> XY circumcentre.  out_xyzi: K x 4 floats (x, y, z=0, intensity = stamp - time_init).
// UB policy: Pk zero-initialised; NaN/inf slopes propagate exactly as IEEE arithmetic does.
// ----------------------------------------------------------------------------------------------------
void orc_get_centroid(const float* xyz16, const int32_t* offsets, const int32_t* indices, int K,
                      double stamp_minus_time_init, float* out_xyzi) {
    const P4* pts = reinterpret_cast<const P4*>(xyz16);
    for (int c = 0; c < K; ++c) {
        const int32_t* idx = indices + offsets[c];
        int n = offsets[c + 1] - offsets[c];
        double Pi[3] = {0, 0, 0}, Pj[3] = {0, 0, 0}, Pk[3] = {0, 0, 0}, Vij[3] = {0, 0, 0};
        float dist_max = -1;
        for (int i = 0; i != n; ++i) {
            for (int j = i + 1; j != n; ++j) {
                double P1[3] = {pts[idx[i]].x, pts[idx[i]].y, pts[idx[i]].z};
                double P2[3] = {pts[idx[j]].x, pts[idx[j]].y, pts[idx[j]].z};
                float dist = (float)std::sqrt((P1[0] - P2[0]) * (P1[0] - P2[0]) + (P1[1] - P2[1]) * (P1[1] - P2[1]) +
                                              (P1[2] - P2[2]) * (P1[2] - P2[2]));  // MOT.cpp:1027
                if (dist > dist_max) {
                    for (int d = 0; d < 3; ++d) { Pi[d] = P1[d]; Pj[d] = P2[d]; }
                    Vij[0] = (P2[1] - P1[1]) / (P2[0] - P1[0]);  // MOT.cpp:753
                    Vij[1] = -1;
                    Vij[2] = Vij[0] * (-P1[0]) + P1[1];
                    dist_max = dist;
                }
            }
        }
        dist_max = -1;
        for (int k = 0; k != n; ++k) {
            double P3[3] = {pts[idx[k]].x, pts[idx[k]].y, pts[idx[k]].z};
            float dist = (float)(std::abs(Vij[0] * P3[0] + Vij[1] * P3[1] + Vij[2]) / std::sqrt(Vij[0] * Vij[0] + Vij[1] * Vij[1]));
            if (dist > dist_max) {
                bool eqj = Pj[0] == P3[0] && Pj[1] == P3[1] && Pj[2] == P3[2];
                bool eqi = Pi[0] == P3[0] && Pi[1] == P3[1] && Pi[2] == P3[2];
                if (eqj || eqi) continue;
                for (int d = 0; d < 3; ++d) Pk[d] = P3[d];
                dist_max = dist;
            }
        }
        float A = (float)(Pj[0] - Pi[0]);
        float B = (float)(Pj[1] - Pi[1]);
        float C = (float)(Pk[0] - Pi[0]);
        float D = (float)(Pk[1] - Pi[1]);
        float E = (float)(A * (Pi[0] + Pj[0]) + B * (Pi[1] + Pj[1]));
        float F = (float)(C * (Pi[0] + Pk[0]) + D * (Pi[1] + Pk[1]));
        float G = (float)(2.0 * (A * (Pk[1] - Pj[1]) - B * (Pk[0] - Pj[0])));
        float* o = out_xyzi + 4 * c;
        if (G == 0) { o[0] = (float)Pi[0]; o[1] = (float)Pi[1]; }
        else { o[0] = (D * E - B * F) / G; o[1] = (A * F - C * E) / G; }
        o[2] = 0.0f;
        o[3] = (float)stamp_minus_time_init;
    }
}

// north_star's per-cluster table (no reference code; pcl::compute3DCentroid semantics for the mean, done
// here in fp64 so that it is the "true" value the fp32 GPU reduction is held to 1e-5 of).
// stats: K x 10 floats: count, mean xyz, min xyz, max xyz.
void orc_cluster_stats(const float* xyz16, const int32_t* offsets, const int32_t* indices, int K, float* stats) {
    const P4* pts = reinterpret_cast<const P4*>(xyz16);
    for (int c = 0; c < K; ++c) {
        int n = offsets[c + 1] - offsets[c];
        double s[3] = {0, 0, 0};
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int t = offsets[c]; t < offsets[c + 1]; ++t) {
            const P4& p = pts[indices[t]];
            s[0] += p.x; s[1] += p.y; s[2] += p.z;
            mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
            mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
        }
        float* o = stats + 10 * c;
        o[0] = (float)n;
        for (int d = 0; d < 3; ++d) { o[1 + d] = (float)(s[d] / n); o[4 + d] = mn[d]; o[7 + d] = mx[d]; }
    }
}

// ----------------------------------------------------------------------------------------------------
// IHGP  (M32.cpp:15-24, IHGP.cpp:12-37, 108-130, 132-162, 164-196, 213-252; driver MOT.cpp:871-920,
// LPF MOT.cpp:824-833, clamp MOT.cpp:649-654).  2x2 fp64, Eigen-free.  Matrices are row-major [a00 a01 a10 a11].
// ----------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {
struct M2 { double a, b, c, d; };  // [[a b][c d]]
inline M2 mul(const M2& x, const M2& y) { return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d}; }
inline M2 tr(const M2& x) { return {x.a, x.c, x.b, x.d}; }
inline M2 add(const M2& x, const M2& y) { return {x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d}; }
inline M2 sub(const M2& x, const M2& y) { return {x.a - y.a, x.b - y.b, x.c - y.c, x.d - y.d}; }

// IHGP.cpp:213-252 with B = H = [1 0]  (B*X*B' = X(0,0),  X*B' = first column of X).
M2 dare(const M2& A, const M2& Q, double R) {
    M2 X{1, 0, 0, 1};
    for (int n = 0; n < 100; ++n) {
        M2 Xp = X;
        double k0 = 0, k1 = 0;
        if (!(std::fabs(R) < 1e-15)) {
            double s = X.a + R;
            double v0 = X.a / s, v1 = X.c / s;          // X*B'/(B*X*B'+R)
            k0 = A.a * v0 + A.b * v1; k1 = A.c * v0 + A.d * v1;  // K = A*(...)
        }
        M2 AKB{A.a - k0, A.b, A.c - k1, A.d};           // A - K*B, B = [1 0]
        M2 KRK{k0 * R * k0, k0 * R * k1, k1 * R * k0, k1 * R * k1};
        X = add(add(mul(mul(AKB, X), tr(AKB)), KRK), Q);
        M2 dX = sub(X, Xp);
        if (std::sqrt(dX.a * dX.a + dX.b * dX.b + dX.c * dX.c + dX.d * dX.d) < 1e-10) break;
    }
    return X;
}
}  // namespace

extern "C" {

// consts (16 doubles): A[4], AKHA[4], K[2], G[4], S, lambda.  hyp = {sigma2, magnSigma2, lengthScale}
// (already exponentiated, as registerNewObstacle does at MOT.cpp:524-530).
void orc_ihgp_setup(double dt, const double* hyp, double* consts) {
    const double sigma2 = hyp[0], magn = hyp[1], ell = hyp[2];
    const double lam = std::sqrt(3.0) / ell;                        // M32.cpp:18
    M2 Pinf{magn, 0, 0, magn * lam * lam};                          // M32.cpp:21
    const double R = sigma2;                                        // M32.cpp:23
    // A = expm(F*dt), F = [[0,1],[-lam^2,-2lam]] (double eigenvalue -lam) -> closed form (IHGP.cpp:15)
    const double e = std::exp(-lam * dt);
    M2 A{e * (1 + lam * dt), e * dt, e * (-lam * lam * dt), e * (1 - lam * dt)};
    M2 Q = sub(Pinf, mul(mul(A, Pinf), tr(A)));                     // IHGP.cpp:16
    M2 PP = dare(A, Q, R);                                          // IHGP.cpp:23
    const double S = PP.a + R;                                      // IHGP.cpp:27
    const double K0 = PP.a / S, K1 = PP.c / S;                      // IHGP.cpp:30
    M2 KHPP{K0 * PP.a, K0 * PP.b, K1 * PP.a, K1 * PP.b};
    M2 PF = sub(PP, KHPP);                                          // IHGP.cpp:33
    M2 KHA{K0 * A.a, K0 * A.b, K1 * A.a, K1 * A.b};
    M2 AKHA = sub(A, KHA);                                          // IHGP.cpp:37
    // smoother gain G = (PPs^-1 * A*PF)'  with PPs = A*PF*A' + Q  (IHGP.cpp:168-170)
    M2 APF = mul(A, PF);
    M2 PPs = add(mul(APF, tr(A)), Q);
    const double det = PPs.a * PPs.d - PPs.b * PPs.c;
    M2 inv{PPs.d / det, -PPs.b / det, -PPs.c / det, PPs.a / det};
    M2 G = tr(mul(inv, APF));
    const double out[16] = {A.a, A.b, A.c, A.d, AKHA.a, AKHA.b, AKHA.c, AKHA.d, K0, K1, G.a, G.b, G.c, G.d, S, lam};
    std::memcpy(consts, out, sizeof(out));
}

// One frame of callIHGP for T tracks.  rings: T x L x 4 floats (x, y, z, intensity=time), oldest first
// (stack_obj, MOT.h:107).  m_state: T x 4 doubles (m_x[0..1], m_y[0..1]) in/out -- init_step does not reset
// m and getEft leaves the smoothed k=0 state behind (IHGP.cpp:108-130, 181-189).
// pos_vel: T x 8 floats: pos(x,y,z=0,intensity) vel(x,y,z=0,intensity).
void orc_ihgp_step(const float* rings, int T, int L, float dt_gp, float lpf_tau, const double* consts_x,
                   const double* consts_y, double* m_state, float* pos_vel) {
    std::vector<double> v(L), mf0(L), mf1(L);
    for (int t = 0; t < T; ++t) {
        const float* c = rings + (size_t)t * L * 4;
        float* o = pos_vel + (size_t)t * 8;
        // LPF_pos, MOT.cpp:824-833 (all float)
        o[0] = (lpf_tau / (lpf_tau + dt_gp)) * c[(L - 2) * 4 + 0] + (dt_gp / (lpf_tau + dt_gp)) * c[(L - 1) * 4 + 0];
        o[1] = (lpf_tau / (lpf_tau + dt_gp)) * c[(L - 2) * 4 + 1] + (dt_gp / (lpf_tau + dt_gp)) * c[(L - 1) * 4 + 1];
        o[2] = 0; o[3] = c[(L - 1) * 4 + 3];
        for (int axis = 0; axis < 2; ++axis) {
            const double* k = axis == 0 ? consts_x : consts_y;
            double* m = m_state + (size_t)t * 4 + 2 * axis;
            const int n = L - 1;  // gp_data_len + 1, MOT.cpp:885
            double mean = 0;      // uninitialised in the reference (MOT.cpp:879-880); oracle policy: 0
            for (int i = 0; i < n; ++i) {
                double vel = (c[(i + 1) * 4 + axis] - c[i * 4 + axis]) / dt_gp;  // float arithmetic, MOT.cpp:889
                v[i] = vel; mean += vel;
            }
            mean = mean / n;
            double m0 = m[0], m1 = m[1];
            for (int i = 0; i < n; ++i) {               // update(), IHGP.cpp:157-160
                double y = v[i] - mean;
                double n0 = (k[4] * m0 + k[5] * m1) + k[8] * y;
                double n1 = (k[6] * m0 + k[7] * m1) + k[9] * y;
                m0 = n0; m1 = n1; mf0[i] = m0; mf1[i] = m1;
            }
            const double eft_last = mf0[n - 1];         // H*MF.back()
            m0 = mf0[n - 1]; m1 = mf1[n - 1];           // getEft(), IHGP.cpp:181-189
            for (int i = n - 2; i >= 0; --i) {
                double r0 = m0 - (k[0] * mf0[i] + k[1] * mf1[i]);
                double r1 = m1 - (k[2] * mf0[i] + k[3] * mf1[i]);
                double s0 = mf0[i] + (k[10] * r0 + k[11] * r1);
                double s1 = mf1[i] + (k[12] * r0 + k[13] * r1);
                m0 = s0; m1 = s1;
            }
            m[0] = m0; m[1] = m1;
            float vel = (float)(eft_last + mean);       // MOT.cpp:914-915
            if (vel > 1.5f) vel = 1.5f; else if (vel < -1.5f) vel = -1.5f;  // MOT.cpp:649-654
            o[4 + axis] = vel;
        }
        o[6] = 0; o[7] = c[(L - 1) * 4 + 3];
    }
}

// ----------------------------------------------------------------------------------------------------
// VoxelGrid downsample (call site MOT.cpp:452-456; PCL VoxelGrid::applyFilter restated):
// leaf (lx,ly,lz); bounds from min/max of the cloud; ijk = floor(p*inv_leaf) - min_b; linear index with
// divb_mul = (1, dx, dx*dy); one centroid per occupied voxel, emitted in ascending voxel index.
// PCL accumulates the centroid in fp32 (Eigen::Vector4f) in sorted-index order; we do the same.
// Returns number of output points.
// ----------------------------------------------------------------------------------------------------
int64_t orc_voxel_grid(const float* xyz16, int64_t n, float lx, float ly, float lz, float* out_xyz16) {
    const P4* pts = reinterpret_cast<const P4*>(xyz16);
    if (n == 0) return 0;
    float mn[3] = {pts[0].x, pts[0].y, pts[0].z}, mx[3] = {pts[0].x, pts[0].y, pts[0].z};
    for (int64_t i = 1; i < n; ++i) {
        mn[0] = std::min(mn[0], pts[i].x); mn[1] = std::min(mn[1], pts[i].y); mn[2] = std::min(mn[2], pts[i].z);
        mx[0] = std::max(mx[0], pts[i].x); mx[1] = std::max(mx[1], pts[i].y); mx[2] = std::max(mx[2], pts[i].z);
    }
    const float inv[3] = {1.0f / lx, 1.0f / ly, 1.0f / lz};
    int minb[3], maxb[3], divb[3];
    for (int d = 0; d < 3; ++d) {
        minb[d] = (int)std::floor(mn[d] * inv[d]);
        maxb[d] = (int)std::floor(mx[d] * inv[d]);
        divb[d] = maxb[d] - minb[d] + 1;
    }
    const int64_t mul1 = divb[0], mul2 = (int64_t)divb[0] * divb[1];
    std::vector<std::pair<int64_t, int64_t>> idx(n);
    for (int64_t i = 0; i < n; ++i) {
        int64_t ijk0 = (int64_t)((int)std::floor(pts[i].x * inv[0]) - minb[0]);
        int64_t ijk1 = (int64_t)((int)std::floor(pts[i].y * inv[1]) - minb[1]);
        int64_t ijk2 = (int64_t)((int)std::floor(pts[i].z * inv[2]) - minb[2]);
        idx[i] = {ijk0 + ijk1 * mul1 + ijk2 * mul2, i};
    }
    std::sort(idx.begin(), idx.end());
    P4* out = reinterpret_cast<P4*>(out_xyz16);
    int64_t k = 0, s = 0;
    while (s < n) {
        int64_t e = s;
        float cx = 0, cy = 0, cz = 0;
        while (e < n && idx[e].first == idx[s].first) { const P4& p = pts[idx[e].second]; cx += p.x; cy += p.y; cz += p.z; ++e; }
        float cnt = (float)(e - s);
        out[k++] = P4{cx / cnt, cy / cnt, cz / cnt, 1.0f};
        s = e;
    }
    return k;
}

}  // extern "C"
