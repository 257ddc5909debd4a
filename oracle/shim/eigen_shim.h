/*
 * eigen_shim.h — the subset of Eigen the reference's hot-path translation units use, so that the
 * UNMODIFIED sources under /root/reference/src/src/{sdf,camera_tracking,eigen_utils,marching_cubes_sdf}.cpp
 * compile in an image that has no Eigen (oracle/Makefile target `ref`, output oracle/_ref/).
 *
 * TEST INFRASTRUCTURE ONLY (same rule as oracle.h).  This is original code, not a copy of Eigen: every
 * operation is evaluated eagerly, in the floating-point order Eigen 3.2 (the 2014 Ubuntu 14.04 / ROS Indigo
 * package the reference was written against; its package.xml pins no version) produces for these fixed
 * sizes without vectorisation-dependent reassociation:
 *
 *   product coefficient (R x K)*(K x C) : ((a0*b0 + a1*b1) + a2*b2) + ...   [CoeffBasedProduct, unrolled]
 *   sum / dot / squaredNorm of 3        : c0 + (c1 + c2)                    [redux_novec_unroller: halves]
 *   Matrix3d::inverse()                 : cofactors, det = c00*m00 + (c10*m10 + c20*m20), times 1/det
 *                                                                            [compute_inverse<.,.,3>]
 *   Matrix<double,6,6>::inverse()       : PartialPivLU (unblocked, column /= pivot is a multiplication by
 *                                         the reciprocal in 3.2), then solve(Identity) with the blocked
 *                                         triangular solver at SSE2 panel width 4 (details at inverse6()).
 *                                         THIS is the one place that cannot be pinned without Eigen itself;
 *                                         the algorithm is stated, not verified (tests allow a few ulp here).
 *   Transform::rotation()               : Eigen runs a JacobiSVD polar extraction; for the orthonormal
 *                                         matrices of the path that is the linear part to <= 1e-15, and the
 *                                         shim returns the linear part.
 *   scalar * matrix, +, -               : coefficient-wise, one rounding each.
 */
#ifndef TSDF_ORACLE_EIGEN_SHIM_H_
#define TSDF_ORACLE_EIGEN_SHIM_H_

#include <cmath>
#include <cstddef>
#include <iostream>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

const int Dynamic = -1;

namespace shim {
template <typename T> struct identity { typedef T type; };
}

template <typename T, int R, int C> class Matrix;

template <typename T, int R, int C>
struct CommaInitializer {
    Matrix<T, R, C>& m;
    int n;
    CommaInitializer(Matrix<T, R, C>& mm, T first) : m(mm), n(0) { put(first); }
    void put(T v) { m(n / C, n % C) = v; n++; }            /* row by row, like Eigen */
    template <typename S> CommaInitializer& operator,(const S& v) { put((T)v); return *this; }
};

template <typename T, int R, int C>
class Matrix {
public:
    typedef T Scalar;
    T d[R * C];                                              /* column-major, Eigen's default */

    Matrix() {}                                              /* uninitialised, like Eigen */
    Matrix(T x, T y) { static_assert(R * C == 2, "size"); d[0] = x; d[1] = y; }
    Matrix(T x, T y, T z) { static_assert(R * C == 3, "size"); d[0] = x; d[1] = y; d[2] = z; }
    Matrix(T x, T y, T z, T w) { static_assert(R * C == 4, "size"); d[0] = x; d[1] = y; d[2] = z; d[3] = w; }

    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Identity() { Matrix m; m.setZero(); for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = T(1); return m; }
    Matrix& setZero() { for (int i = 0; i < R * C; i++) d[i] = T(0); return *this; }

    T& operator()(int i) { return d[i]; }
    const T& operator()(int i) const { return d[i]; }
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
    T& operator()(int i, int j) { return d[i + j * R]; }
    const T& operator()(int i, int j) const { return d[i + j * R]; }
    T& coeffRef(int i, int j) { return d[i + j * R]; }

    T x() const { return d[0]; }
    T y() const { return d[1]; }
    T z() const { return d[2]; }

    template <typename S> CommaInitializer<T, R, C> operator<<(const S& v) { return CommaInitializer<T, R, C>(*this, (T)v); }

    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> t;
        for (int i = 0; i < R; i++)
            for (int j = 0; j < C; j++) t(j, i) = (*this)(i, j);
        return t;
    }

    Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; i++) d[i] = d[i] + o.d[i]; return *this; }
    Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; i++) d[i] = d[i] - o.d[i]; return *this; }

    /* redux over a fixed-size vector without vectorisation: the range is split in halves recursively */
    static T redux_sum(const T* c, int start, int len) {
        if (len == 1) return c[start];
        const int half = len / 2;
        return redux_sum(c, start, half) + redux_sum(c, start + half, len - half);
    }
    T sum() const { return redux_sum(d, 0, R * C); }
    T dot(const Matrix& o) const {
        T c[R * C];
        for (int i = 0; i < R * C; i++) c[i] = d[i] * o.d[i];
        return redux_sum(c, 0, R * C);
    }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }

    Matrix inverse() const;                                  /* 3x3 and 6x6 only, defined below */
};

template <typename T, int R, int C>
Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) {
    Matrix<T, R, C> r;
    for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <typename T, int R, int C>
Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) {
    Matrix<T, R, C> r;
    for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <typename T, int R, int C>
Matrix<T, R, C> operator-(const Matrix<T, R, C>& a) {
    Matrix<T, R, C> r;
    for (int i = 0; i < R * C; i++) r.d[i] = -a.d[i];
    return r;
}
/* coefficient-based product, accumulation left to right starting from the first product */
template <typename T, int R, int K, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const Matrix<T, K, C>& b) {
    Matrix<T, R, C> r;
    for (int j = 0; j < C; j++)
        for (int i = 0; i < R; i++) {
            T acc = a(i, 0) * b(0, j);
            for (int k = 1; k < K; k++) acc = acc + a(i, k) * b(k, j);
            r(i, j) = acc;
        }
    return r;
}
template <typename T, int R, int C>
Matrix<T, R, C> operator*(typename shim::identity<T>::type s, const Matrix<T, R, C>& a) {
    Matrix<T, R, C> r;
    for (int i = 0; i < R * C; i++) r.d[i] = s * a.d[i];
    return r;
}
template <typename T, int R, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, typename shim::identity<T>::type s) {
    Matrix<T, R, C> r;
    for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] * s;
    return r;
}
template <typename T, int R, int C>
Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, typename shim::identity<T>::type s) {
    Matrix<T, R, C> r;
    for (int i = 0; i < R * C; i++) r.d[i] = a.d[i] / s;   /* scalar_quotient1_op: a true division */
    return r;
}
template <typename T, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& m) {
    for (int i = 0; i < R; i++) {
        for (int j = 0; j < C; j++) os << (j ? " " : "") << m(i, j);
        if (i + 1 < R) os << "\n";
    }
    return os;
}

namespace shim {

/* compute_inverse<MatrixType, ResultType, 3>: first-column cofactors, determinant as their
 * product-sum with column 0 (3-term redux: c0 + (c1 + c2)), every cofactor times 1/det. */
template <typename T>
inline T cofactor3(const Matrix<T, 3, 3>& m, int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
}
template <typename T>
inline Matrix<T, 3, 3> inverse3(const Matrix<T, 3, 3>& m) {
    Matrix<T, 3, 3> r;
    const T c0 = cofactor3(m, 0, 0), c1 = cofactor3(m, 1, 0), c2 = cofactor3(m, 2, 0);
    const T det = c0 * m(0, 0) + (c1 * m(1, 0) + c2 * m(2, 0));
    const T invdet = T(1) / det;
    r(0, 0) = c0 * invdet; r(0, 1) = c1 * invdet; r(0, 2) = c2 * invdet;   /* row 0 = cofactors_col0 * invdet */
    r(1, 0) = cofactor3(m, 0, 1) * invdet;
    r(1, 1) = cofactor3(m, 1, 1) * invdet;
    r(1, 2) = cofactor3(m, 2, 1) * invdet;
    r(2, 0) = cofactor3(m, 0, 2) * invdet;
    r(2, 1) = cofactor3(m, 1, 2) * invdet;
    r(2, 2) = cofactor3(m, 2, 2) * invdet;
    return r;
}

/* Matrix<double,6,6>::inverse() = partialPivLu().inverse() = solve(Identity).
 *  1. unblocked_lu (sizes <= 16 are never blocked): per column k the first entry of largest magnitude at or
 *     below the diagonal is the pivot; rows are swapped; the sub-column is multiplied by the RECIPROCAL of the
 *     pivot (Eigen 3.2's operator/=(Scalar) is `*= Scalar(1)/other` for non-integers); rank-1 update of the
 *     trailing block, one multiply and one subtract per entry.  A zero pivot column is left as is (Eigen
 *     records first_zero_pivot and carries on) -> infinities/NaNs in the result, exactly the reference's
 *     unguarded behaviour (SURVEY TRAP 12).
 *  2. X = P * I, then X = L^-1 X (unit lower), then X = U^-1 X, each by triangular_solve_matrix<OnTheLeft,
 *     ColMajor>: panels of SmallPanelWidth = max(mr, nr) = 4 rows (SSE2, double) are solved entry by entry
 *     (`b = (x(i,j) *= 1/tri(i,i))`, then `x(s..,j) -= b * tri(s..,i)` inside the panel), and the rows outside
 *     the panel are updated by the GEBP kernel: acc = sum of the panel's products in increasing k starting from
 *     the first product, then x += (-1) * acc. */
template <typename T>
inline Matrix<T, 6, 6> inverse6(const Matrix<T, 6, 6>& A) {
    const int N = 6, PW = 4;
    Matrix<T, 6, 6> lu = A;
    int tr[6];
    for (int k = 0; k < N; k++) {
        int piv = k;
        T best = std::fabs(lu(k, k));
        for (int r = k + 1; r < N; r++) {
            const T v = std::fabs(lu(r, k));
            if (v > best) { best = v; piv = r; }
        }
        tr[k] = piv;
        if (best != T(0)) {
            if (piv != k)
                for (int c = 0; c < N; c++) { const T t = lu(k, c); lu(k, c) = lu(piv, c); lu(piv, c) = t; }
            const T rcp = T(1) / lu(k, k);
            for (int r = k + 1; r < N; r++) lu(r, k) = lu(r, k) * rcp;
        }
        for (int c = k + 1; c < N; c++)
            for (int r = k + 1; r < N; r++) lu(r, c) = lu(r, c) - lu(r, k) * lu(k, c);
    }
    Matrix<T, 6, 6> X = Matrix<T, 6, 6>::Identity();
    for (int k = 0; k < N; k++)
        if (tr[k] != k)
            for (int c = 0; c < N; c++) { const T t = X(k, c); X(k, c) = X(tr[k], c); X(tr[k], c) = t; }
    /* unit lower */
    for (int k1 = 0; k1 < N; k1 += PW) {
        const int pw = (N - k1 < PW) ? N - k1 : PW;
        for (int k = 0; k < pw; k++) {
            const int i = k1 + k, rs = pw - k - 1;
            for (int j = 0; j < N; j++) {
                const T b = X(i, j);
                for (int i3 = 0; i3 < rs; i3++) X(i + 1 + i3, j) = X(i + 1 + i3, j) - b * lu(i + 1 + i3, i);
            }
        }
        for (int r = k1 + pw; r < N; r++)
            for (int j = 0; j < N; j++) {
                T acc = lu(r, k1) * X(k1, j);
                for (int k = k1 + 1; k < k1 + pw; k++) acc = acc + lu(r, k) * X(k, j);
                X(r, j) = X(r, j) + acc * T(-1);
            }
    }
    /* upper */
    for (int k1 = 0; k1 < N; k1 += PW) {
        const int pw = (N - k1 < PW) ? N - k1 : PW;
        for (int k = 0; k < pw; k++) {
            const int i = N - k1 - k - 1, rs = pw - k - 1, s = i - rs;
            const T a = T(1) / lu(i, i);
            for (int j = 0; j < N; j++) {
                X(i, j) = X(i, j) * a;
                const T b = X(i, j);
                for (int i3 = 0; i3 < rs; i3++) X(s + i3, j) = X(s + i3, j) - b * lu(s + i3, i);
            }
        }
        const int start_block = N - k1 - pw;
        for (int r = 0; r < start_block; r++)
            for (int j = 0; j < N; j++) {
                T acc = lu(r, start_block) * X(start_block, j);
                for (int k = start_block + 1; k < start_block + pw; k++) acc = acc + lu(r, k) * X(k, j);
                X(r, j) = X(r, j) + acc * T(-1);
            }
    }
    return X;
}

template <typename T, int R, int C> struct Inverter;
template <typename T> struct Inverter<T, 3, 3> { static Matrix<T, 3, 3> run(const Matrix<T, 3, 3>& m) { return inverse3(m); } };
template <typename T> struct Inverter<T, 6, 6> { static Matrix<T, 6, 6> run(const Matrix<T, 6, 6>& m) { return inverse6(m); } };

}  // namespace shim

template <typename T, int R, int C>
Matrix<T, R, C> Matrix<T, R, C>::inverse() const { return shim::Inverter<T, R, C>::run(*this); }

/* VectorXd: only what eigen_utils.cpp needs (construction from a fixed vector, * scalar, []) */
template <typename T>
class Matrix<T, Dynamic, 1> {
public:
    std::vector<T> v;
    Matrix() {}
    template <int N> Matrix(const Matrix<T, N, 1>& o) : v(o.d, o.d + N) {}
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
    int size() const { return (int)v.size(); }
    Matrix operator*(T s) const { Matrix r; r.v.resize(v.size()); for (size_t i = 0; i < v.size(); i++) r.v[i] = v[i] * s; return r; }
};

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, Dynamic, 1> VectorXd;

enum TransformTraits { Affine = 2 };

/* Transform<double,3,Affine>: a 4x4 column-major block; (i,j) access, identity, linear part, translation */
template <typename T, int Dim, int Mode>
class Transform {
public:
    Matrix<T, Dim + 1, Dim + 1> m;
    Transform() { for (int j = 0; j < Dim; j++) m(Dim, j) = T(0); m(Dim, Dim) = T(1); }   /* makeAffine() */
    T& operator()(int i, int j) { return m(i, j); }
    const T& operator()(int i, int j) const { return m(i, j); }
    void setIdentity() { m = Matrix<T, Dim + 1, Dim + 1>::Identity(); }
    Matrix<T, Dim, Dim> linear() const {
        Matrix<T, Dim, Dim> r;
        for (int i = 0; i < Dim; i++)
            for (int j = 0; j < Dim; j++) r(i, j) = m(i, j);
        return r;
    }
    Matrix<T, Dim, Dim> rotation() const { return linear(); }   /* see the header comment */
    Matrix<T, Dim, 1> translation() const {
        Matrix<T, Dim, 1> r;
        for (int i = 0; i < Dim; i++) r(i) = m(i, Dim);
        return r;
    }
};
typedef Transform<double, 3, Affine> Affine3d;

}  // namespace Eigen

#endif
