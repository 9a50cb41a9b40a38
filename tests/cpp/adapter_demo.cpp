// adapter_demo.cpp -- the reference's clusterPointCloud call sequence (MOT.cpp:461-491) written against the
// adapter.  Reads a cloud (binary float32 N x 4) and a map, prints "M K" then per cluster "size first_index cx cy".
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mot_b200_pcl.hpp"

static std::vector<char> slurp(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> b((size_t)n);
    if (fread(b.data(), 1, (size_t)n, f) != (size_t)n) exit(2);
    fclose(f);
    return b;
}

int main(int argc, char** argv) {
    if (argc < 9) { fprintf(stderr, "usage: cloud.bin map.bin W H res ox oy tol [min max]\n"); return 2; }
    std::vector<char> cb = slurp(argv[1]), mb = slurp(argv[2]);
    const int W = atoi(argv[3]), H = atoi(argv[4]);
    const float res = (float)atof(argv[5]);
    const double ox = atof(argv[6]), oy = atof(argv[7]);
    const float ClusterTolerance = (float)atof(argv[8]);
    const int MinClusterSize = argc > 9 ? atoi(argv[9]) : 5, MaxClusterSize = argc > 10 ? atoi(argv[10]) : 300;

    pcl::PointCloud<pcl::PointXYZ> cloud_1;
    cloud_1.points.resize(cb.size() / sizeof(pcl::PointXYZ));
    memcpy(cloud_1.points.data(), cb.data(), cloud_1.points.size() * sizeof(pcl::PointXYZ));

    mot_b200::Tracker gpu(0, cloud_1.points.size());
    const double q[4] = {0, 0, 0, 1};
    gpu.setMap(reinterpret_cast<const int8_t*>(mb.data()), W, H, res, ox, oy, q, 2);

    // --- MOT.cpp:451-491, with the GPU objects substituted ---
    if (argc > 11) {  // optional VoxelGrid stage (MOT.cpp:452-456), leaf (L, L, 20 L)
        const float VoxelLeafSize = (float)atof(argv[11]);
        pcl::PointCloud<pcl::PointXYZ> input_cloud = cloud_1;
        mot_b200::VoxelGrid vg(gpu);
        vg.setInputCloud(input_cloud.makeShared());
        vg.setLeafSize(1 * VoxelLeafSize, 1 * VoxelLeafSize, 20 * VoxelLeafSize);
        vg.filter(cloud_1);
    }
    pcl::PointCloud<pcl::PointXYZ> cloud_2;
    pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_filtered(new pcl::PointCloud<pcl::PointXYZ>);
    cloud_2 = gpu.removeStatic(cloud_1);
    *cloud_filtered = cloud_2;
    if (cloud_filtered->empty()) { printf("0 0\n"); return 0; }
    pcl::search::KdTree<pcl::PointXYZ>::Ptr tree(new pcl::search::KdTree<pcl::PointXYZ>);
    tree->setInputCloud(cloud_filtered);
    std::vector<pcl::PointIndices> cluster_indices;
    mot_b200::EuclideanClusterExtraction ec(gpu);
    ec.setClusterTolerance(ClusterTolerance);
    ec.setMinClusterSize(MinClusterSize);
    ec.setMaxClusterSize(MaxClusterSize);
    ec.setSearchMethod(tree);
    ec.setInputCloud(cloud_filtered);
    ec.extract(cluster_indices);
    std::vector<pcl::PointXYZI> clusterCentroids = gpu.getCentroid(2.5);
    // ---
    printf("%zu %zu\n", cloud_filtered->size(), cluster_indices.size());
    for (size_t k = 0; k < cluster_indices.size(); ++k)
        printf("%zu %d %.9g %.9g %.9g\n", cluster_indices[k].indices.size(), cluster_indices[k].indices[0], clusterCentroids[k].x,
               clusterCentroids[k].y, clusterCentroids[k].intensity);
    return 0;
}
