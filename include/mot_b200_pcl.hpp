// mot_b200_pcl.hpp -- header-only C++ adapter over the C ABI (mot_b200.h) that keeps the reference tracker's call
// sites source compatible (reference src/multiple_object_tracking_lidar.cpp, "MOT.cpp"):
//
//   MOT.cpp:452   pcl::VoxelGrid<pcl::PointXYZ> vg; ... vg.filter(cloud_1) ->  mot_b200::VoxelGrid vg(gpu); same calls
//   MOT.cpp:461   cloud_2 = removeStatic(cloud_1);                       ->  gpu.removeStatic(cloud_1)
//   MOT.cpp:472   tree->setInputCloud(cloud_filtered);                   ->  accepted, ignored (no KD-tree is built)
//   MOT.cpp:480   pcl::EuclideanClusterExtraction<pcl::PointXYZ> ec;     ->  mot_b200::EuclideanClusterExtraction ec(gpu);
//   MOT.cpp:481+  ec.setClusterTolerance / setMinClusterSize / setMaxClusterSize / setSearchMethod / setInputCloud
//   MOT.cpp:488   ec.extract(cluster_indices);                           ->  same signature, std::vector<pcl::PointIndices>&
//   MOT.cpp:491   getCentroid(cluster_indices, *cloud_filtered, *input)  ->  gpu.getCentroid(stamp - time_init)
//   MOT.cpp:223   callIHGP(this_objIDs)                                  ->  gpu.ihgpStep(rings, m_state, pos_vel)
//
// Error convention: the reference has none (no return codes, no exceptions; empty input -> empty output,
// MOT.cpp:465-469).  The adapter keeps "empty in -> empty out" and throws std::runtime_error for real failures
// (no CUDA device, capacity exceeded, NaN in the cloud) -- there is deliberately no CPU fallback.
#ifndef MOT_B200_PCL_HPP_
#define MOT_B200_PCL_HPP_

#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mot_b200.h"
#ifndef MOT_B200_USE_REAL_PCL
#include "pcl_compat.h"
#endif

namespace mot_b200 {

class Tracker {
  public:
    Tracker(int device, std::size_t max_points, std::size_t max_tracks = 4096) {
        const int rc = mot_create(device, max_points, max_tracks, &h_);
        if (rc != MOT_OK) throw std::runtime_error("mot_create failed (no CUDA device? there is no CPU fallback), code " + std::to_string(rc));
        max_points_ = max_points;
    }
    ~Tracker() { if (h_) mot_destroy(h_); }
    Tracker(const Tracker&) = delete;
    Tracker& operator=(const Tracker&) = delete;
    mot_handle* handle() { return h_; }

    // ObstacleTrack::mapCallback (MOT.cpp:235-251).  GridT = nav_msgs::OccupancyGrid (anything with .info.width,
    // .info.height, .info.resolution, .info.origin.position.{x,y}, .info.origin.orientation.{x,y,z,w} and .data).
    template <typename GridT>
    void mapCallback(const GridT& map_msg, int static_tolarance) {
        const double q[4] = {map_msg.info.origin.orientation.x, map_msg.info.origin.orientation.y, map_msg.info.origin.orientation.z,
                             map_msg.info.origin.orientation.w};
        setMap(reinterpret_cast<const int8_t*>(map_msg.data.data()), (int)map_msg.info.width, (int)map_msg.info.height,
               (float)map_msg.info.resolution, map_msg.info.origin.position.x, map_msg.info.origin.position.y, q, static_tolarance);
    }
    void setMap(const int8_t* occ, int width, int height, float resolution, double origin_x, double origin_y, const double quat_xyzw[4],
                int static_tolarance) {
        check(mot_set_map(h_, occ, width, height, resolution, origin_x, origin_y, quat_xyzw, static_tolarance));
    }

    // ObstacleTrack::removeStatic (MOT.cpp:664-706): same signature shape, kept points in input order.
    pcl::PointCloud<pcl::PointXYZ> removeStatic(const pcl::PointCloud<pcl::PointXYZ>& in) {
        pcl::PointCloud<pcl::PointXYZ> out;
        out.header = in.header;
        if (in.points.empty()) return out;
        out.points.resize(in.points.size());
        std::size_t m = 0;
        check(mot_remove_static(h_, reinterpret_cast<const float*>(in.points.data()), in.points.size(),
                                reinterpret_cast<float*>(out.points.data()), out.points.size(), &m));
        out.points.resize(m);
        out.width = (std::uint32_t)m;
        out.height = 1;
        out.is_dense = in.is_dense;
        return out;
    }

    // ObstacleTrack::getCentroid (MOT.cpp:708-822) for the clusters of the last extract() on this tracker.
    std::vector<pcl::PointXYZI> getCentroid(double stamp_minus_time_init) {
        std::size_t m = 0, total = 0;
        int32_t K = 0;
        check(mot_result_counts(h_, &m, &K, &total));
        std::vector<pcl::PointXYZI> out((std::size_t)K);
        if (K == 0) return out;
        std::vector<float> xyzi((std::size_t)K * 4);
        check(mot_get_centroid(h_, stamp_minus_time_init, xyzi.data(), (std::size_t)K));
        for (int k = 0; k < K; ++k) {
            out[k].x = xyzi[4 * k + 0]; out[k].y = xyzi[4 * k + 1]; out[k].z = xyzi[4 * k + 2]; out[k].intensity = xyzi[4 * k + 3];
        }
        return out;
    }

    std::vector<mot_cluster_stat> clusterStats() {
        std::size_t m = 0, total = 0;
        int32_t K = 0;
        check(mot_result_counts(h_, &m, &K, &total));
        std::vector<mot_cluster_stat> out((std::size_t)K);
        if (K) check(mot_cluster_stats(h_, out.data(), out.size()));
        return out;
    }

    // registerNewObstacle's GP construction (MOT.cpp:521-534) + callIHGP's per-track loop (MOT.cpp:621-662), batched.
    void ihgpConfigure(double dt_gp, float lpf_tau, const double hyp_x[3], const double hyp_y[3], int data_length) {
        check(mot_ihgp_configure(h_, dt_gp, lpf_tau, hyp_x, hyp_y, data_length));
        data_length_ = data_length;
    }
    // stack_obj: one ring of data_length centroids per track (MOT.h:107); m_state: 4 doubles per track (in/out).
    std::vector<std::vector<pcl::PointXYZI>> callIHGP(const std::vector<std::vector<pcl::PointXYZI>>& stack_obj, std::vector<double>& m_state) {
        const std::size_t T = stack_obj.size();
        std::vector<std::vector<pcl::PointXYZI>> pos_vel_s(T);
        if (T == 0) return pos_vel_s;
        if (m_state.size() != 4 * T) throw std::runtime_error("m_state must hold 4 doubles per track");
        std::vector<float> rings(T * (std::size_t)data_length_ * 4), pv(T * 8);
        for (std::size_t t = 0; t < T; ++t) {
            if ((int)stack_obj[t].size() != data_length_) throw std::runtime_error("every ring must hold data_length centroids");
            for (int k = 0; k < data_length_; ++k) {
                float* r = &rings[(t * data_length_ + k) * 4];
                r[0] = stack_obj[t][k].x; r[1] = stack_obj[t][k].y; r[2] = stack_obj[t][k].z; r[3] = stack_obj[t][k].intensity;
            }
        }
        check(mot_ihgp_step(h_, rings.data(), (int)T, m_state.data(), pv.data()));
        for (std::size_t t = 0; t < T; ++t) {
            pos_vel_s[t].resize(2);
            for (int s = 0; s < 2; ++s) {
                pcl::PointXYZI& o = pos_vel_s[t][s];
                o.x = pv[t * 8 + 4 * s]; o.y = pv[t * 8 + 4 * s + 1]; o.z = pv[t * 8 + 4 * s + 2]; o.intensity = pv[t * 8 + 4 * s + 3];
            }
        }
        return pos_vel_s;
    }

    // the association / lifecycle / callIHGP part of cloudCallback (MOT.cpp:176-233) with the track table kept
    // on the device.  Returns false on the reference's "first frame" / "No obstacles around" paths (nothing to publish).
    bool tracksStep(const std::vector<pcl::PointXYZI>& clusterCentroids, double stamp_minus_time_init, float id_threshold, float frequency,
                    std::vector<int>& this_objIDs, std::vector<std::vector<pcl::PointXYZI>>& pos_vel_s, std::vector<mot_obstacle>* obstacles = nullptr) {
        const std::size_t K = clusterCentroids.size();
        std::vector<float> cen(K * 4), pv(K * 8);
        for (std::size_t k = 0; k < K; ++k) {
            cen[4 * k] = clusterCentroids[k].x; cen[4 * k + 1] = clusterCentroids[k].y; cen[4 * k + 2] = clusterCentroids[k].z;
            cen[4 * k + 3] = clusterCentroids[k].intensity;
        }
        this_objIDs.assign(K, -1);
        if (obstacles) obstacles->resize(K);
        int32_t n_tracks = 0, produced = 0;
        check(mot_tracks_step(h_, cen.data(), (int)K, stamp_minus_time_init, id_threshold, frequency, this_objIDs.data(), pv.data(),
                              obstacles ? obstacles->data() : nullptr, &n_tracks, &produced));
        pos_vel_s.assign(produced ? K : 0, std::vector<pcl::PointXYZI>(2));
        for (std::size_t k = 0; produced && k < K; ++k)
            for (int s = 0; s < 2; ++s) {
                pcl::PointXYZI& o = pos_vel_s[k][s];
                o.x = pv[k * 8 + 4 * s]; o.y = pv[k * 8 + 4 * s + 1]; o.z = pv[k * 8 + 4 * s + 2]; o.intensity = pv[k * 8 + 4 * s + 3];
            }
        return produced != 0;
    }

    void check(int rc) {
        if (rc < 0) throw std::runtime_error(std::string("mot_b200: ") + mot_last_error(h_) + " (code " + std::to_string(rc) + ")");
    }

  private:
    mot_handle* h_ = nullptr;
    std::size_t max_points_ = 0;
    int data_length_ = 10;
};

// Drop-in for pcl::VoxelGrid<pcl::PointXYZ> as used at MOT.cpp:452-456.
class VoxelGrid {
  public:
    explicit VoxelGrid(Tracker& t) : t_(t) {}
    void setInputCloud(const pcl::PointCloud<pcl::PointXYZ>::ConstPtr& cloud) { cloud_ = cloud; }
    void setInputCloud(const pcl::PointCloud<pcl::PointXYZ>::Ptr& cloud) { cloud_ = cloud; }
    void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
    void filter(pcl::PointCloud<pcl::PointXYZ>& out) {
        out.points.clear();
        if (cloud_) out.header = cloud_->header;
        if (!cloud_ || cloud_->points.empty()) { out.width = 0; return; }
        out.points.resize(cloud_->points.size());
        std::size_t m = 0;
        t_.check(mot_voxel_grid(t_.handle(), reinterpret_cast<const float*>(cloud_->points.data()), cloud_->points.size(), leaf_[0], leaf_[1],
                                leaf_[2], reinterpret_cast<float*>(out.points.data()), out.points.size(), &m));
        out.points.resize(m);
        out.width = (std::uint32_t)m;
        out.height = 1;
        out.is_dense = true;
    }

  private:
    Tracker& t_;
    pcl::PointCloud<pcl::PointXYZ>::ConstPtr cloud_;
    float leaf_[3] = {0.05f, 0.05f, 1.0f};
};

// Drop-in for pcl::EuclideanClusterExtraction<pcl::PointXYZ> as used at MOT.cpp:480-488.
class EuclideanClusterExtraction {
  public:
    explicit EuclideanClusterExtraction(Tracker& t) : t_(t) {}
    void setClusterTolerance(double tol) { tol_ = tol; }
    void setMinClusterSize(int n) { min_ = n; }
    void setMaxClusterSize(int n) { max_ = n; }
    template <typename TreePtr>
    void setSearchMethod(const TreePtr&) {}  // the KD-tree is not used
    void setInputCloud(const pcl::PointCloud<pcl::PointXYZ>::ConstPtr& cloud) { cloud_ = cloud; }
    void setInputCloud(const pcl::PointCloud<pcl::PointXYZ>::Ptr& cloud) { cloud_ = cloud; }

    void extract(std::vector<pcl::PointIndices>& clusters) {
        clusters.clear();  // PCL clears the output on empty input
        if (!cloud_ || cloud_->points.empty()) return;
        const std::size_t m = cloud_->points.size();
        t_.check(mot_set_cluster_params(t_.handle(), (float)tol_, min_, max_));
        offsets_.resize(m + 1);
        indices_.resize(m);
        int32_t K = 0;
        t_.check(mot_cluster(t_.handle(), reinterpret_cast<const float*>(cloud_->points.data()), m, offsets_.data(), offsets_.size(),
                             indices_.data(), indices_.size(), &K));
        clusters.resize((std::size_t)K);
        for (int k = 0; k < K; ++k) {
            clusters[k].header = cloud_->header;  // PCL copies the input header into every PointIndices
            clusters[k].indices.assign(indices_.begin() + offsets_[k], indices_.begin() + offsets_[k + 1]);
        }
    }

  private:
    Tracker& t_;
    double tol_ = 0.15;
    int min_ = 5, max_ = 200;  // reference defaults, MOT.cpp:90-92
    pcl::PointCloud<pcl::PointXYZ>::ConstPtr cloud_;
    std::vector<int32_t> offsets_, indices_;
};

}  // namespace mot_b200
#endif
