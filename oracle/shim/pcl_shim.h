/*
 * pcl_shim.h — the PCL types the reference's hot-path translation units use: PointXYZ, PointXYZRGB,
 * Normal and the organised PointCloud<T> container (at(column,row), push_back, +=).  TEST INFRASTRUCTURE
 * ONLY (see eigen_shim.h).  Containers only — no PCL algorithm is restated here (the bilateral filter and
 * integral-image normals the node calls before the boundary are NOT part of these translation units).
 */
#ifndef TSDF_ORACLE_PCL_SHIM_H_
#define TSDF_ORACLE_PCL_SHIM_H_

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <omp.h>   /* reaches marching_cubes_sdf.cpp through PCL headers in a real build */
#include <stdexcept>
#include <string>
#include <vector>
#include "boost_shim.h"

#define PCL_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#define PCL_DEBUG(...) do { } while (0)

namespace pcl {

struct PointXYZ {
    float x, y, z, pad_;
    PointXYZ() : x(0), y(0), z(0), pad_(1.0f) {}
    PointXYZ(float a, float b, float c) : x(a), y(b), z(c), pad_(1.0f) {}
};
struct PointXYZRGB {
    float x, y, z, pad_;
    uint8_t b, g, r, a;                                      /* PCL's byte order inside the packed rgb */
    float pad2_[3];
    PointXYZRGB() : x(0), y(0), z(0), pad_(1.0f), b(0), g(0), r(0), a(255) {}
};
struct Normal {
    float normal_x, normal_y, normal_z, pad_;
    float curvature;
    float pad2_[3];
    Normal() : normal_x(0), normal_y(0), normal_z(0), pad_(0), curvature(0) {}
};

template <typename PointT>
class PointCloud {
public:
    typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
    typedef boost::shared_ptr<const PointCloud<PointT> > ConstPtr;
    std::vector<PointT> points;
    uint32_t width, height;
    bool is_dense;
    PointCloud() : width(0), height(0), is_dense(true) {}
    /* organised access: (column, row), bounds-checked, only for height > 1 — as in PCL */
    const PointT& at(int column, int row) const {
        if (this->height > 1) return points.at((size_t)row * this->width + column);
        throw std::runtime_error("Can't use 2D indexing with a unorganized point cloud");
    }
    PointT& at(int column, int row) {
        if (this->height > 1) return points.at((size_t)row * this->width + column);
        throw std::runtime_error("Can't use 2D indexing with a unorganized point cloud");
    }
    void push_back(const PointT& pt) {
        points.push_back(pt);
        width = (uint32_t)points.size();
        height = 1;
    }
    size_t size() const { return points.size(); }
    void clear() { points.clear(); width = 0; height = 0; }
    PointCloud& operator+=(const PointCloud& rhs) {
        points.insert(points.end(), rhs.points.begin(), rhs.points.end());
        width = (uint32_t)points.size();
        height = 1;
        is_dense = is_dense && rhs.is_dense;
        return *this;
    }
};

}  // namespace pcl

#endif
