// mot_ref_shim.h -- TEST INFRASTRUCTURE ONLY.  Stand-ins for the ROS 1 message / node API and the PCL classes that
// the reference's tracker source (MOT.h / MOT.cpp) names, so that oracle/_ref can be compiled from the reference's own
// .cpp files where they lie (oracle/Makefile, target _ref; neither ROS nor PCL nor Eigen is in this image).
//
// What is real and what is a stand-in in such a build:
//   * every statement of the reference's ObstacleTrack, InfiniteHorizonGP and Matern32model classes runs unchanged;
//   * ros::NodeHandle::param answers from a table the harness fills, publishers hand their messages to the harness,
//     subscribers and timing calls do nothing;
//   * pcl::fromROSMsg, pcl::VoxelGrid and pcl::search::KdTree + pcl::EuclideanClusterExtraction (third-party code, PCL 1.x,
//     not part of the reference tree) forward to the oracle's restatements in mot_oracle.cpp -- these three stay
//     "restated, unpinned".
// All forwarding headers (ros/ros.h, pcl/point_types.h, Eigen/Dense ...) are generated into oracle/_ref/include by the
// Makefile and contain a single #include of this file.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "mini_eigen.h"

// ---- the oracle's restatements of the PCL pieces (oracle/mot_oracle.cpp) -------------------------------------------
extern "C" {
int orc_cluster_kdtree(const float* xyz16, int m, float tol, int min_size, int max_size, int build_twice, int32_t* offsets,
                       int32_t* indices);
int64_t orc_voxel_grid(const float* xyz16, int64_t n, float lx, float ly, float lz, float* out_xyz16);
}

// ---- harness side of the stand-ins ----------------------------------------------------------------------------------
namespace costmap_converter { struct ObstacleArrayMsg; }
namespace ref_shim {
struct Bus {
    std::map<std::string, double> params;                              // ros parameter server
    double now = 0.0;                                                  // ros::Time::now()
    std::shared_ptr<costmap_converter::ObstacleArrayMsg> last_obstacles;  // last message on the obstacle topic
    int obstacle_publishes = 0;
    std::vector<std::string> log;                                      // ROS_INFO / ROS_ERROR lines
    static Bus& get() {
        static Bus b;
        return b;
    }
};
template <typename M> inline void capture(const std::string&, const M&) {}
void capture(const std::string& topic, const costmap_converter::ObstacleArrayMsg& m);  // defined in ref_harness.cpp
}  // namespace ref_shim

// ---- ros ----------------------------------------------------------------------------------------------------------------
namespace ros {
struct Time {
    double sec = 0.0;
    Time() {}
    explicit Time(double s) : sec(s) {}
    double toSec() const { return sec; }
    static Time now() { return Time(ref_shim::Bus::get().now); }
};
struct Duration {
    explicit Duration(double) {}
    void sleep() {}
};
inline bool ok() { return true; }
inline void spinOnce() {}
inline void init(int&, char**, const std::string&) {}
struct Exception : std::exception {
    const char* what() const noexcept override { return "ros::Exception"; }
};
namespace this_node { inline std::string getName() { return "multiple_object_tracking_lidar"; } }

class Publisher {
public:
    Publisher() {}
    explicit Publisher(const std::string& t) : topic_(t) {}
    template <typename M> void publish(const M& m) const { ref_shim::capture(topic_, m); }
private:
    std::string topic_;
};
class Subscriber {};

class NodeHandle {
public:
    bool ok() const { return true; }
    bool deleteParam(const std::string&) { return true; }
    template <typename T> bool param(const std::string& name, T& value, const T& fallback) const {
        const auto& p = ref_shim::Bus::get().params;
        const auto it = p.find(name);
        if (it == p.end()) {
            value = fallback;
            return false;
        }
        value = static_cast<T>(it->second);
        return true;
    }
    template <typename M> Publisher advertise(const std::string& topic, uint32_t) { return Publisher(topic); }
    template <typename M, typename T> Subscriber subscribe(const std::string&, uint32_t, void (T::*)(M), T*) { return Subscriber(); }
};
}  // namespace ros

#define REF_SHIM_LOG(expr)                       \
    do {                                         \
        std::ostringstream ref_shim_os;          \
        ref_shim_os << expr;                     \
        ref_shim::Bus::get().log.push_back(ref_shim_os.str()); \
    } while (0)
#define ROS_INFO_STREAM(expr) REF_SHIM_LOG(expr)
#define ROS_ERROR_STREAM(expr) REF_SHIM_LOG(expr)
#define ROS_INFO(...) REF_SHIM_LOG(#__VA_ARGS__)
#define ROS_ERROR(...) REF_SHIM_LOG(#__VA_ARGS__)
#define ROS_WARN(...) REF_SHIM_LOG(#__VA_ARGS__)
#define ROS_INFO_STREAM_NAMED(...) REF_SHIM_LOG(#__VA_ARGS__)
#define ROS_ERROR_STREAM_NAMED(...) REF_SHIM_LOG(#__VA_ARGS__)

// ---- messages -------------------------------------------------------------------------------------------------------------
namespace std_msgs {
struct Header {
    uint32_t seq = 0;
    ros::Time stamp;
    std::string frame_id;
};
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
struct Int8 { int8_t data = 0; };
struct Int64 { int64_t data = 0; };
struct Float32 { float data = 0; };
struct Float64 { double data = 0; };
struct Float32MultiArray { std::vector<float> data; };
struct Int32MultiArray { std::vector<int32_t> data; };
}  // namespace std_msgs

namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Point32 { float x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; std::array<double, 36> covariance{}; };
struct Polygon { std::vector<Point32> points; };
}  // namespace geometry_msgs

namespace sensor_msgs {
struct PointField {
    enum { INT8 = 1, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 };
    std::string name;
    uint32_t offset = 0;
    uint8_t datatype = 0;
    uint32_t count = 0;
};
struct PointCloud2 {
    std_msgs::Header header;
    uint32_t height = 0, width = 0;
    std::vector<PointField> fields;
    bool is_bigendian = false;
    uint32_t point_step = 0, row_step = 0;
    std::vector<uint8_t> data;
    bool is_dense = false;
};
typedef std::shared_ptr<PointCloud2> PointCloud2Ptr;
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
struct ChannelFloat32 { std::string name; std::vector<float> values; };
struct PointCloud {
    std_msgs::Header header;
    std::vector<geometry_msgs::Point32> points;
    std::vector<ChannelFloat32> channels;
};
}  // namespace sensor_msgs

namespace nav_msgs {
struct MapMetaData {
    ros::Time map_load_time;
    float resolution = 0;
    uint32_t width = 0, height = 0;
    geometry_msgs::Pose origin;
};
struct OccupancyGrid {
    std_msgs::Header header;
    MapMetaData info;
    std::vector<int8_t> data;
};
}  // namespace nav_msgs

namespace costmap_converter {
struct ObstacleMsg {
    std_msgs::Header header;
    geometry_msgs::Polygon polygon;
    double radius = 0;
    int64_t id = 0;
    geometry_msgs::Quaternion orientation;
    geometry_msgs::TwistWithCovariance velocities;
};
struct ObstacleArrayMsg {
    std_msgs::Header header;
    std::vector<ObstacleMsg> obstacles;
};
}  // namespace costmap_converter

namespace visualization_msgs {
struct Marker {
    enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, POINTS = 8, TEXT_VIEW_FACING = 9 };
    enum { ADD = 0, MODIFY = 0, DELETE = 2 };
    std_msgs::Header header;
    std::string ns;
    int32_t id = 0, type = 0, action = 0;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
    std::vector<geometry_msgs::Point> points;
    std::vector<std_msgs::ColorRGBA> colors;
    std::string text;
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs

// ---- pcl ----------------------------------------------------------------------------------------------------------------
namespace pcl {
struct alignas(16) PointXYZ {
    float x = 0, y = 0, z = 0, pad = 1.0f;  // PCL's data[3] is 1
    PointXYZ() {}
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) PointXYZI {
    float x = 0, y = 0, z = 0, pad = 1.0f;
    float intensity = 0;
    float pad2[3] = {0, 0, 0};
};
struct PCLHeader {
    uint32_t seq = 0;
    uint64_t stamp = 0;
    std::string frame_id;
};
template <typename PointT>
class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    PCLHeader header;
    std::vector<PointT> points;
    uint32_t width = 0, height = 1;
    bool is_dense = true;
    void push_back(const PointT& p) {
        points.push_back(p);
        width = (uint32_t)points.size();
        height = 1;
    }
    bool empty() const { return points.empty(); }
    std::size_t size() const { return points.size(); }
    Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};
struct PointIndices {
    PCLHeader header;
    std::vector<int> indices;
};

// pcl::fromROSMsg for a PointXYZ target: field-mapped copy of the FLOAT32 fields named x, y, z (restated)
inline void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointXYZ>& cloud) {
    int off[3] = {-1, -1, -1};
    for (const auto& f : msg.fields) {
        if (f.datatype != sensor_msgs::PointField::FLOAT32) continue;
        if (f.name == "x") off[0] = (int)f.offset;
        if (f.name == "y") off[1] = (int)f.offset;
        if (f.name == "z") off[2] = (int)f.offset;
    }
    const std::size_t n = (std::size_t)msg.width * msg.height;
    cloud.points.assign(n, PointXYZ());
    cloud.width = msg.width;
    cloud.height = msg.height;
    cloud.is_dense = msg.is_dense;
    cloud.header.frame_id = msg.header.frame_id;
    for (std::size_t i = 0; i < n; ++i) {
        const uint8_t* p = msg.data.data() + i * msg.point_step;
        float v[3] = {0, 0, 0};
        for (int d = 0; d < 3; ++d)
            if (off[d] >= 0) std::memcpy(&v[d], p + off[d], 4);
        cloud.points[i].x = v[0];
        cloud.points[i].y = v[1];
        cloud.points[i].z = v[2];
    }
}

template <typename PointT>
class VoxelGrid {
public:
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { in_ = c; }
    void setLeafSize(float lx, float ly, float lz) { lx_ = lx; ly_ = ly; lz_ = lz; }
    void filter(PointCloud<PointT>& out) {
        const std::size_t n = in_->points.size();
        std::vector<PointT> tmp(n ? n : 1);
        const int64_t k = orc_voxel_grid(reinterpret_cast<const float*>(in_->points.data()), (int64_t)n, lx_, ly_, lz_,
                                         reinterpret_cast<float*>(tmp.data()));
        tmp.resize((std::size_t)k);
        out.points.swap(tmp);
        out.width = (uint32_t)out.points.size();
        out.height = 1;
        out.header = in_->header;
    }
private:
    typename PointCloud<PointT>::ConstPtr in_;
    float lx_ = 0, ly_ = 0, lz_ = 0;
};

namespace search {
template <typename PointT>
class KdTree {
public:
    typedef std::shared_ptr<KdTree<PointT>> Ptr;
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr&) {}
};
}  // namespace search

template <typename PointT>
class EuclideanClusterExtraction {
public:
    void setClusterTolerance(double t) { tol_ = t; }
    void setMinClusterSize(int v) { min_ = v; }
    void setMaxClusterSize(int v) { max_ = v; }
    void setSearchMethod(const typename search::KdTree<PointT>::Ptr&) {}
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { in_ = c; }
    void extract(std::vector<PointIndices>& clusters) {
        const int m = (int)in_->points.size();
        std::vector<int32_t> off((std::size_t)m + 2, 0), idx((std::size_t)(m ? m : 1), 0);
        // the reference builds the tree on the cloud itself and PCL rebuilds it inside extract: build_twice
        const int k = orc_cluster_kdtree(reinterpret_cast<const float*>(in_->points.data()), m, (float)tol_, min_, max_, 1,
                                         off.data(), idx.data());
        clusters.clear();
        for (int c = 0; c < k; ++c) {
            PointIndices pi;
            pi.header = in_->header;
            pi.indices.assign(idx.begin() + off[c], idx.begin() + off[c + 1]);
            clusters.push_back(pi);
        }
    }
private:
    typename PointCloud<PointT>::ConstPtr in_;
    double tol_ = 0;
    int min_ = 1, max_ = 2147483647;
};
}  // namespace pcl
