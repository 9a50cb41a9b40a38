// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own classes, compiled by
// `make -C oracle _ref` together with the reference sources where they lie:
//     /root/reference/src/multiple_object_tracking_lidar.cpp   (ObstacleTrack)
//     /root/reference/src/ihgp/InfiniteHorizonGP.cpp, Matern32model.cpp
// against the stand-in headers of oracle/shim (ROS messages / node API, PCL containers, a small Eigen).  Nothing of the
// reference is copied into this repository: the build reads the sources from /root/reference and writes only
// oracle/_ref/libmot_ref.so.  Used to (1) check the oracle's restatements of removeStatic, getCentroid, callIHGP and the
// association / track lifecycle against the code they restate and (2) generate the golden vectors under tests/golden/
// (tests/golden/make_ref_fixtures.py).  The PCL pieces inside (VoxelGrid, KdTree + EuclideanClusterExtraction,
// fromROSMsg) are the oracle's own restatements (see mot_ref_shim.h), so they are not pinned by this build.
//
// Undefined behaviour in the reference that the harness has to take a position on (same positions as the oracle):
//   * ObstacleTrack::dt_gp is read by registerNewObstacle on the first frame before it is first assigned
//     (MOT.cpp:153-159 vs :533) -- the harness sets dt_gp = 1/frequency right after initialize();
//   * IHGP_fixed_vel accumulates into uninitialised doubles (MOT.cpp:879-880) -- the _ref build uses
//     -ftrivial-auto-var-init=zero, i.e. they start at 0;
//   * removeStatic indexes the map outside its bounds near the border (MOT.cpp:686) -- the Eigen stand-in throws and the
//     harness returns -2.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "shim/mot_ref_shim.h"

#define private public
#include "multiple_object_tracking_lidar/multiple_object_tracking_lidar.h"
#undef private

namespace ref_shim {
void capture(const std::string& topic, const costmap_converter::ObstacleArrayMsg& m) {
    (void)topic;
    Bus& b = Bus::get();
    b.last_obstacles = std::make_shared<costmap_converter::ObstacleArrayMsg>(m);
    b.obstacle_publishes += 1;
}
}  // namespace ref_shim

namespace {
struct Ref {
    ObstacleTrack* trk = nullptr;
    void* mem = nullptr;
};

sensor_msgs::PointCloud2Ptr make_cloud_msg(const float* pts, int64_t n, double stamp) {
    auto msg = std::make_shared<sensor_msgs::PointCloud2>();
    msg->header.stamp = ros::Time(stamp);
    msg->header.frame_id = "map";
    msg->height = 1;
    msg->width = (uint32_t)n;
    const char* names[3] = {"x", "y", "z"};
    for (int d = 0; d < 3; ++d) {
        sensor_msgs::PointField f;
        f.name = names[d];
        f.offset = 4u * d;
        f.datatype = sensor_msgs::PointField::FLOAT32;
        f.count = 1;
        msg->fields.push_back(f);
    }
    msg->point_step = 16;
    msg->row_step = (uint32_t)(16 * n);
    msg->is_dense = true;
    msg->data.resize((size_t)n * 16);
    if (n) std::memcpy(msg->data.data(), pts, (size_t)n * 16);
    return msg;
}
}  // namespace

extern "C" {

// Parameters are given by their launch-file names (without the /multiple_object_tracking_lidar/ prefix).
void* ref_create(const char* const* names, const double* values, int n_params) {
    ref_shim::Bus& bus = ref_shim::Bus::get();
    bus.params.clear();
    bus.log.clear();
    bus.last_obstacles.reset();
    bus.obstacle_publishes = 0;
    bus.now = 0.0;
    for (int i = 0; i < n_params; ++i) bus.params[std::string("/multiple_object_tracking_lidar/") + names[i]] = values[i];
    Ref* r = new Ref;
    r->mem = std::calloc(1, sizeof(ObstacleTrack));
    r->trk = new (r->mem) ObstacleTrack();
    if (!r->trk->initialize()) return nullptr;
    r->trk->dt_gp = 1 / r->trk->frequency;  // see the header comment
    return r;
}

void ref_destroy(void* h) {
    Ref* r = static_cast<Ref*>(h);
    if (!r) return;
    r->trk->~ObstacleTrack();
    std::free(r->mem);
    delete r;
}

void ref_set_now(double now) { ref_shim::Bus::get().now = now; }

// occ: H x W int8 row-major (nav_msgs/OccupancyGrid.data); quat: x, y, z, w
int ref_set_map(void* h, const int8_t* occ, int W, int H, float resolution, double ox, double oy, const double* quat_xyzw) {
    Ref* r = static_cast<Ref*>(h);
    nav_msgs::OccupancyGrid g;
    g.info.resolution = resolution;
    g.info.width = (uint32_t)W;
    g.info.height = (uint32_t)H;
    g.info.origin.position.x = ox;
    g.info.origin.position.y = oy;
    g.info.origin.orientation.x = quat_xyzw[0];
    g.info.origin.orientation.y = quat_xyzw[1];
    g.info.origin.orientation.z = quat_xyzw[2];
    g.info.origin.orientation.w = quat_xyzw[3];
    g.data.assign(occ, occ + (size_t)W * H);
    r->trk->mapCallback(g);
    return 0;
}

float ref_yaw_from_quat(void* h, const double* quat_xyzw) {
    Ref* r = static_cast<Ref*>(h);
    geometry_msgs::Quaternion q;
    q.x = quat_xyzw[0]; q.y = quat_xyzw[1]; q.z = quat_xyzw[2]; q.w = quat_xyzw[3];
    return r->trk->quaternion2eularYaw(q);
}

// ObstacleTrack::removeStatic on n points (16-byte records).  Returns the number of kept points (written to out), -2 if
// the reference indexed the map out of bounds.
int64_t ref_remove_static(void* h, const float* pts, int64_t n, float* out) {
    Ref* r = static_cast<Ref*>(h);
    pcl::PointCloud<pcl::PointXYZ> in;
    in.points.resize((size_t)n);
    if (n) std::memcpy(static_cast<void*>(in.points.data()), pts, (size_t)n * 16);
    try {
        pcl::PointCloud<pcl::PointXYZ> kept = r->trk->removeStatic(in);
        if (!kept.points.empty()) std::memcpy(out, kept.points.data(), kept.points.size() * 16);
        return (int64_t)kept.points.size();
    } catch (const std::out_of_range&) {
        return -2;
    }
}

// ObstacleTrack::getCentroid on a CSR cluster list.  out: K x 4 floats (x, y, z, intensity).
int ref_get_centroid(void* h, const float* pts, int64_t n, const int32_t* off, const int32_t* idx, int K, double stamp,
                     double time_init, float* out) {
    Ref* r = static_cast<Ref*>(h);
    pcl::PointCloud<pcl::PointXYZ> cloud;
    cloud.points.resize((size_t)n);
    if (n) std::memcpy(static_cast<void*>(cloud.points.data()), pts, (size_t)n * 16);
    std::vector<pcl::PointIndices> clusters((size_t)K);
    for (int c = 0; c < K; ++c) clusters[c].indices.assign(idx + off[c], idx + off[c + 1]);
    sensor_msgs::PointCloud2 input;
    input.header.stamp = ros::Time(stamp);
    r->trk->time_init = time_init;
    const std::vector<pcl::PointXYZI> cen = r->trk->getCentroid(clusters, cloud, input);
    for (size_t c = 0; c < cen.size(); ++c) {
        out[4 * c + 0] = cen[c].x;
        out[4 * c + 1] = cen[c].y;
        out[4 * c + 2] = cen[c].z;
        out[4 * c + 3] = cen[c].intensity;
    }
    return (int)cen.size();
}

// ObstacleTrack::clusterPointCloud on one frame: fromROSMsg -> VoxelGrid -> removeStatic -> clustering -> getCentroid.
// Returns the number of centroids (K x 4 floats: x, y, z, intensity = stamp - time_init), -2 on an out-of-bounds map read.
int ref_cluster_point_cloud(void* h, const float* pts, int64_t n, double stamp, float* out, int cap) {
    Ref* r = static_cast<Ref*>(h);
    sensor_msgs::PointCloud2ConstPtr msg = make_cloud_msg(pts, n, stamp);
    try {
        const std::vector<pcl::PointXYZI> cen = r->trk->clusterPointCloud(msg);
        for (size_t c = 0; c < cen.size() && (int)c < cap; ++c) {
            out[4 * c + 0] = cen[c].x;
            out[4 * c + 1] = cen[c].y;
            out[4 * c + 2] = cen[c].z;
            out[4 * c + 3] = cen[c].intensity;
        }
        return (int)cen.size();
    } catch (const std::out_of_range&) {
        return -2;
    }
}

// ObstacleTrack::cloudCallback on one frame (x, y, z float32 PointCloud2).  Returns the number of obstacles of the LAST
// ObstacleArrayMsg this call published (0 if it published nothing) and writes ids[T] and pos_vel[T x 8]:
// polygon point (x, y, 0, -) and twist.linear (x, y, 0, -) of each obstacle.
int ref_cloud_callback(void* h, const float* pts, int64_t n, double stamp, int32_t* ids, float* pos_vel, int cap) {
    Ref* r = static_cast<Ref*>(h);
    ref_shim::Bus& bus = ref_shim::Bus::get();
    bus.last_obstacles.reset();
    sensor_msgs::PointCloud2ConstPtr msg = make_cloud_msg(pts, n, stamp);
    try {
        r->trk->cloudCallback(msg);
    } catch (const std::out_of_range&) {
        return -2;
    }
    if (!bus.last_obstacles) return 0;
    const auto& ob = bus.last_obstacles->obstacles;
    const int T = (int)ob.size();
    for (int i = 0; i < T && i < cap; ++i) {
        ids[i] = (int32_t)ob[i].id;
        float* o = pos_vel + 8 * i;
        o[0] = ob[i].polygon.points[0].x;
        o[1] = ob[i].polygon.points[0].y;
        o[2] = ob[i].polygon.points[0].z;
        o[3] = (float)ob[i].radius;
        o[4] = (float)ob[i].velocities.twist.linear.x;
        o[5] = (float)ob[i].velocities.twist.linear.y;
        o[6] = (float)ob[i].velocities.twist.linear.z;
        o[7] = (float)ob[i].velocities.covariance[0];
    }
    return T;
}

// The tracker's internal lists: objIDs, stack_obj (T x L x 4 floats: x, y, z, intensity) and the GP means
// (T x 4 doubles: m_x[0..1], m_y[0..1]).  Returns the number of tracks.
int ref_tracks(void* h, int32_t* ids, float* rings, double* m_state, int cap) {
    Ref* r = static_cast<Ref*>(h);
    ObstacleTrack& t = *r->trk;
    const int T = (int)t.objIDs.size();
    const int L = t.data_length;
    for (int i = 0; i < T && i < cap; ++i) {
        ids[i] = t.objIDs[i];
        for (int k = 0; k < L; ++k) {
            const pcl::PointXYZI& p = t.stack_obj[i][k];
            float* o = rings + ((size_t)i * L + k) * 4;
            o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.intensity;
        }
        m_state[4 * i + 0] = t.GPs_x[i]->m(0);
        m_state[4 * i + 1] = t.GPs_x[i]->m(1);
        m_state[4 * i + 2] = t.GPs_y[i]->m(0);
        m_state[4 * i + 3] = t.GPs_y[i]->m(1);
    }
    return T;
}

int ref_next_obj_num(void* h) { return static_cast<Ref*>(h)->trk->next_obj_num; }

// The constants of one InfiniteHorizonGP built the way registerNewObstacle builds it (MOT.cpp:520-534), in the oracle's
// layout: A[4], AKHA[4], K[2], G[4], S, lambda (row-major 2x2).  G is what getEft computes (IHGP.cpp:168-170).
void ref_ihgp_constants(double dt, const double* hyp /* sigma2, magnSigma2, lengthScale */, double* consts) {
    Matern32model model;
    model.setSigma2(hyp[0]);
    model.setMagnSigma2(hyp[1]);
    model.setLengthScale(hyp[2]);
    InfiniteHorizonGP gp(dt, model.getF(), model.getH(), model.getPinf(), model.getR(), model.getdF(), model.getdPinf(), model.getdR());
    gp.init_step();
    Eigen::MatrixXd PP = gp.A * gp.PF * gp.A.transpose() + gp.Q;
    Eigen::MatrixXd G = PP.ldlt().solve(gp.A * gp.PF).transpose();
    const double out[16] = {gp.A(0, 0), gp.A(0, 1), gp.A(1, 0), gp.A(1, 1), gp.AKHA(0, 0), gp.AKHA(0, 1), gp.AKHA(1, 0), gp.AKHA(1, 1),
                            gp.K(0), gp.K(1), G(0, 0), G(0, 1), G(1, 0), G(1, 1), gp.S, std::sqrt(3.0) / hyp[2]};
    std::memcpy(consts, out, sizeof(out));
}

// ObstacleTrack::callIHGP on T tracks whose rings are given (T x L x 4 floats, L = data_length).  On the first call with
// an empty tracker the T tracks are registered (registerNewObstacle); later calls must pass the same T.  The GP means
// persist inside the reference objects between calls; m_state (T x 4 doubles) is written INTO them before the call when
// set_state != 0 and read back after it.  pos_vel: T x 8 floats, pos (x, y, z, intensity) then vel (x, y, z, intensity).
int ref_call_ihgp(void* h, const float* rings, int T, int set_state, double* m_state, float* pos_vel) {
    Ref* r = static_cast<Ref*>(h);
    ObstacleTrack& t = *r->trk;
    const int L = t.data_length;
    if (t.objIDs.empty()) {
        for (int i = 0; i < T; ++i) {
            pcl::PointXYZI c;
            c.x = rings[(size_t)i * L * 4 + 0];
            c.y = rings[(size_t)i * L * 4 + 1];
            t.registerNewObstacle(c);
        }
    }
    if ((int)t.objIDs.size() != T) return -1;
    for (int i = 0; i < T; ++i) {
        for (int k = 0; k < L; ++k) {
            const float* c = rings + ((size_t)i * L + k) * 4;
            pcl::PointXYZI& p = t.stack_obj[i][k];
            p.x = c[0]; p.y = c[1]; p.z = c[2]; p.intensity = c[3];
        }
        if (set_state) {
            t.GPs_x[i]->m(0) = m_state[4 * i + 0];
            t.GPs_x[i]->m(1) = m_state[4 * i + 1];
            t.GPs_y[i]->m(0) = m_state[4 * i + 2];
            t.GPs_y[i]->m(1) = m_state[4 * i + 3];
        }
    }
    const std::vector<std::vector<pcl::PointXYZI>> pv = t.callIHGP(t.objIDs);
    for (int i = 0; i < T; ++i) {
        float* o = pos_vel + 8 * i;
        o[0] = pv[i][0].x; o[1] = pv[i][0].y; o[2] = pv[i][0].z; o[3] = pv[i][0].intensity;
        o[4] = pv[i][1].x; o[5] = pv[i][1].y; o[6] = pv[i][1].z; o[7] = pv[i][1].intensity;
        m_state[4 * i + 0] = t.GPs_x[i]->m(0);
        m_state[4 * i + 1] = t.GPs_x[i]->m(1);
        m_state[4 * i + 2] = t.GPs_y[i]->m(0);
        m_state[4 * i + 3] = t.GPs_y[i]->m(1);
    }
    return T;
}

int ref_log_lines(void) { return (int)ref_shim::Bus::get().log.size(); }

}  // extern "C"
