/* pcl_compat.h -- minimal, layout-compatible stand-ins for the handful of PCL types that cross the boundary of
 * the tracker's hot path.  ONLY for builds where real PCL is absent (this repository's container and tests);
 * with real PCL installed, include <pcl/point_types.h>, <pcl/point_cloud.h>, <pcl/PointIndices.h> instead and
 * define MOT_B200_USE_REAL_PCL before including mot_b200_pcl.hpp.
 *
 * Layout facts relied upon (upstream PCL): pcl::PointXYZ is 16 bytes {float x,y,z; float pad},
 * 16-byte aligned; pcl::PointXYZI is 32 bytes {x,y,z,pad, intensity, pad[3]}; pcl::PointIndices holds a header
 * and std::vector<int> indices. */
#ifndef MOT_B200_PCL_COMPAT_H_
#define MOT_B200_PCL_COMPAT_H_

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace pcl {

struct PCLHeader {
    std::uint32_t seq = 0;
    std::uint64_t stamp = 0;
    std::string frame_id;
};

struct alignas(16) PointXYZ {
    float x = 0, y = 0, z = 0;
    float data_pad = 1.0f;
    PointXYZ() = default;
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ must be 16 bytes");

struct alignas(16) PointXYZI {
    float x = 0, y = 0, z = 0;
    float data_pad = 1.0f;
    float intensity = 0;
    float pad_[3] = {0, 0, 0};
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI must be 32 bytes");

template <typename PointT>
class PointCloud {
  public:
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
    PCLHeader header;
    std::vector<PointT> points;
    std::uint32_t width = 0, height = 1;
    bool is_dense = true;
    void push_back(const PointT& p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = 0; }
    void resize(std::size_t n) { points.resize(n); width = (std::uint32_t)n; height = 1; }
    PointT& operator[](std::size_t i) { return points[i]; }
    const PointT& operator[](std::size_t i) const { return points[i]; }
    Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};

struct PointIndices {
    PCLHeader header;
    std::vector<int> indices;
};

namespace search {
template <typename PointT>
class KdTree {  // accepted and ignored: the GPU path builds its own voxel grid
  public:
    using Ptr = std::shared_ptr<KdTree<PointT>>;
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr&) {}
};
}  // namespace search
}  // namespace pcl

#endif
