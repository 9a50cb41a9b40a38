// mini_eigen.h -- TEST INFRASTRUCTURE ONLY.  A small dense-matrix stand-in for the part of Eigen 3 that the
// reference sources use (Eigen itself is not in this image).  It exists so that oracle/_ref can be compiled from the
// reference's own .cpp files (see oracle/Makefile, target _ref): the reference's statements run unchanged, the linear
// algebra primitives underneath (products, matrix exponential, LDL^T solve) are the ones below, written from the
// textbook definitions -- all in IEEE double, so they agree with Eigen's to rounding.
//
// Covered API (by use in the reference, MOT.h / MOT.cpp / IHGP.cpp / M32.cpp):
//   MatrixXd, VectorXd, Vector3d; (r,c) / (i) access, rows/cols, resize, setZero, setIdentity, Zero(), operator<< ... ,
//   + - * / with matrices and scalars, transpose, eval, dot, norm, ==, exp() (matrix exponential), corner blocks,
//   ldlt().solve().
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace Eigen {

class MatrixXd;
class BlockRef;
template <typename M> class LDLT;

class CommaInit {
public:
    CommaInit(MatrixXd& m, double first);
    template <typename T> CommaInit& operator,(const T& v) { put((double)v); return *this; }
private:
    void put(double v);
    MatrixXd& m_;
    std::size_t k_ = 0;
};

class MatrixXd {
public:
    MatrixXd() : r_(0), c_(0) {}
    MatrixXd(int r, int c) : r_(r), c_(c), d_((std::size_t)r * c, 0.0) {}
    static MatrixXd Zero(int r, int c) { return MatrixXd(r, c); }
    int rows() const { return r_; }
    int cols() const { return c_; }
    std::size_t size() const { return d_.size(); }
    void resize(int r, int c) { r_ = r; c_ = c; d_.assign((std::size_t)r * c, 0.0); }
    void setZero() { std::fill(d_.begin(), d_.end(), 0.0); }
    void setZero(int r, int c) { resize(r, c); }
    void setIdentity() {
        setZero();
        for (int i = 0; i < std::min(r_, c_); ++i) (*this)(i, i) = 1.0;
    }
    // row-major storage: the comma initialiser fills in reading order, as Eigen's does
    // (Eigen does not range-check in release builds -- an out-of-range access is undefined behaviour there; the
    // stand-in throws so that a harness call ends with an error instead of reading foreign memory.)
    double& operator()(int i, int j) { return d_[at(i, j)]; }
    double operator()(int i, int j) const { return d_[at(i, j)]; }
    double& operator()(int i) { return d_[(std::size_t)i]; }   // vectors and 1x1 results
    double operator()(int i) const { return d_[(std::size_t)i]; }
    template <typename T> CommaInit operator<<(const T& v) { return CommaInit(*this, (double)v); }

    MatrixXd transpose() const {
        MatrixXd t(c_, r_);
        for (int i = 0; i < r_; ++i)
            for (int j = 0; j < c_; ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    MatrixXd eval() const { return *this; }
    double norm() const {  // Frobenius
        double s = 0;
        for (double v : d_) s += v * v;
        return std::sqrt(s);
    }
    double dot(const MatrixXd& o) const {
        assert(d_.size() == o.d_.size());
        double s = 0;
        for (std::size_t i = 0; i < d_.size(); ++i) s += d_[i] * o.d_[i];
        return s;
    }
    bool operator==(const MatrixXd& o) const { return r_ == o.r_ && c_ == o.c_ && d_ == o.d_; }
    bool operator!=(const MatrixXd& o) const { return !(*this == o); }

    MatrixXd exp() const;                  // matrix exponential (unsupported/Eigen/MatrixFunctions)
    LDLT<MatrixXd> ldlt() const;           // Eigen/Cholesky
    BlockRef topLeftCorner(int r, int c);
    BlockRef topRightCorner(int r, int c);
    BlockRef bottomLeftCorner(int r, int c);
    BlockRef bottomRightCorner(int r, int c);

protected:
    std::size_t at(int i, int j) const {
        if (i < 0 || j < 0 || i >= r_ || j >= c_) throw std::out_of_range("mini_eigen: index outside the matrix");
        return (std::size_t)i * c_ + j;
    }
    int r_, c_;
    std::vector<double> d_;
};

inline CommaInit::CommaInit(MatrixXd& m, double first) : m_(m) { put(first); }
inline void CommaInit::put(double v) {
    assert(k_ < m_.size());
    m_((int)k_++) = v;
}

inline MatrixXd operator+(const MatrixXd& a, const MatrixXd& b) {
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    MatrixXd r(a.rows(), a.cols());
    for (std::size_t i = 0; i < r.size(); ++i) r((int)i) = a((int)i) + b((int)i);
    return r;
}
inline MatrixXd operator-(const MatrixXd& a, const MatrixXd& b) {
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    MatrixXd r(a.rows(), a.cols());
    for (std::size_t i = 0; i < r.size(); ++i) r((int)i) = a((int)i) - b((int)i);
    return r;
}
inline MatrixXd operator-(const MatrixXd& a) {
    MatrixXd r(a.rows(), a.cols());
    for (std::size_t i = 0; i < r.size(); ++i) r((int)i) = -a((int)i);
    return r;
}
inline MatrixXd operator*(const MatrixXd& a, const MatrixXd& b) {
    assert(a.cols() == b.rows());
    MatrixXd r(a.rows(), b.cols());
    for (int i = 0; i < a.rows(); ++i)
        for (int j = 0; j < b.cols(); ++j) {
            double s = 0;
            for (int k = 0; k < a.cols(); ++k) s += a(i, k) * b(k, j);
            r(i, j) = s;
        }
    return r;
}
inline MatrixXd operator*(const MatrixXd& a, double s) {
    MatrixXd r(a.rows(), a.cols());
    for (std::size_t i = 0; i < r.size(); ++i) r((int)i) = a((int)i) * s;
    return r;
}
inline MatrixXd operator*(double s, const MatrixXd& a) { return a * s; }
inline MatrixXd operator/(const MatrixXd& a, double s) {
    MatrixXd r(a.rows(), a.cols());
    for (std::size_t i = 0; i < r.size(); ++i) r((int)i) = a((int)i) / s;
    return r;
}

// corner block: assignable from a matrix, convertible to one
class BlockRef {
public:
    BlockRef(MatrixXd& m, int i0, int j0, int r, int c) : m_(m), i0_(i0), j0_(j0), r_(r), c_(c) {}
    BlockRef& operator=(const MatrixXd& src) {
        assert(src.rows() == r_ && src.cols() == c_);
        for (int i = 0; i < r_; ++i)
            for (int j = 0; j < c_; ++j) m_(i0_ + i, j0_ + j) = src(i, j);
        return *this;
    }
    operator MatrixXd() const {
        MatrixXd out(r_, c_);
        for (int i = 0; i < r_; ++i)
            for (int j = 0; j < c_; ++j) out(i, j) = m_(i0_ + i, j0_ + j);
        return out;
    }
private:
    MatrixXd& m_;
    int i0_, j0_, r_, c_;
};
inline BlockRef MatrixXd::topLeftCorner(int r, int c) { return BlockRef(*this, 0, 0, r, c); }
inline BlockRef MatrixXd::topRightCorner(int r, int c) { return BlockRef(*this, 0, c_ - c, r, c); }
inline BlockRef MatrixXd::bottomLeftCorner(int r, int c) { return BlockRef(*this, r_ - r, 0, r, c); }
inline BlockRef MatrixXd::bottomRightCorner(int r, int c) { return BlockRef(*this, r_ - r, c_ - c, r, c); }

// exp(M) by scaling and squaring around a Taylor polynomial: M / 2^s has 1-norm <= 1/8, 20 terms leave a
// truncation error below 1e-30, then s squarings.
inline MatrixXd MatrixXd::exp() const {
    assert(r_ == c_);
    double n1 = 0;
    for (int j = 0; j < c_; ++j) {
        double s = 0;
        for (int i = 0; i < r_; ++i) s += std::fabs((*this)(i, j));
        n1 = std::max(n1, s);
    }
    int sq = 0;
    while (std::ldexp(n1, -sq) > 0.125) ++sq;
    const MatrixXd X = (*this) * std::ldexp(1.0, -sq);
    MatrixXd E(r_, c_), term(r_, c_);
    E.setIdentity();
    term.setIdentity();
    for (int k = 1; k <= 20; ++k) {
        term = (term * X) / (double)k;
        E = E + term;
    }
    for (int s = 0; s < sq; ++s) E = E * E;
    return E;
}

// Solver behind ldlt().solve(): the reference applies it to a symmetric positive definite 2x2; Gaussian elimination with
// partial pivoting gives the same solution to rounding for any non-singular matrix.
template <typename M>
class LDLT {
public:
    LDLT() {}
    explicit LDLT(const M& a) : a_(a) {}
    M solve(const M& b) const {
        const int n = a_.rows();
        assert(a_.cols() == n && b.rows() == n);
        M A = a_, X = b;
        for (int k = 0; k < n; ++k) {
            int p = k;
            for (int i = k + 1; i < n; ++i)
                if (std::fabs(A(i, k)) > std::fabs(A(p, k))) p = i;
            if (p != k) {
                for (int j = 0; j < n; ++j) std::swap(A(k, j), A(p, j));
                for (int j = 0; j < X.cols(); ++j) std::swap(X(k, j), X(p, j));
            }
            for (int i = k + 1; i < n; ++i) {
                const double f = A(i, k) / A(k, k);
                for (int j = k; j < n; ++j) A(i, j) -= f * A(k, j);
                for (int j = 0; j < X.cols(); ++j) X(i, j) -= f * X(k, j);
            }
        }
        for (int k = n - 1; k >= 0; --k)
            for (int j = 0; j < X.cols(); ++j) {
                double s = X(k, j);
                for (int i = k + 1; i < n; ++i) s -= A(k, i) * X(i, j);
                X(k, j) = s / A(k, k);
            }
        return X;
    }
private:
    M a_;
};
inline LDLT<MatrixXd> MatrixXd::ldlt() const { return LDLT<MatrixXd>(*this); }

class VectorXd : public MatrixXd {
public:
    VectorXd() {}
    explicit VectorXd(int n) : MatrixXd(n, 1) {}
    VectorXd(const MatrixXd& m) : MatrixXd(m) { assert(m.cols() == 1 || m.size() == 0); }
    VectorXd& operator=(const MatrixXd& m) {
        assert(m.cols() == 1 || m.size() == 0);
        MatrixXd::operator=(m);
        return *this;
    }
    static VectorXd Zero(int n) { return VectorXd(n); }
    using MatrixXd::setZero;
    void setZero(int n) { resize(n, 1); }
    using MatrixXd::operator<<;
};

class Vector3d : public MatrixXd {
public:
    Vector3d() : MatrixXd(3, 1) {}  // Eigen leaves it uninitialised; zero here
    Vector3d(double x, double y, double z) : MatrixXd(3, 1) { d_[0] = x; d_[1] = y; d_[2] = z; }
    Vector3d(const MatrixXd& m) : MatrixXd(m) { assert(m.rows() == 3 && m.cols() == 1); }
    using MatrixXd::operator<<;
};

}  // namespace Eigen
