// oracle/eigen_shim/ggp_eigen_shim.hpp — a stand-in for the subset of Eigen 3.3 that the reference's wrapper headers
// (likelihood.h, predictions.h, correlation_tree.h, Gaussians.h, moma_input.h, mean_cov_model.h) use, so that those
// headers compile UNMODIFIED from /root/reference/src into oracle/_ref/libggp_ref_wrappers.so (oracle/ref_wrappers.cpp).
// TEST INFRASTRUCTURE ONLY.  Eigen itself is neither vendored by the reference nor present in this image.
//
// What this file pins and what it assumes.  All of the reference's own logic (which quantity is propagated, updated,
// multiplied, inverted, in which statement order, with which stale state) is the reference's source, compiled as is.
// What remains an assumption is how Eigen 3.3.7 (README.md:56) evaluates the few dense kernels those statements hit;
// they are restated here once, from Eigen 3.3's sources as remembered (Eigen is not available to re-check):
//   E1  every expression node is evaluated coefficient by coefficient with one IEEE operation per node; a product nested
//       in a sum / difference / other product is evaluated into a temporary first; chained products associate left to
//       right (C++ operator grammar); no FMA (the reference is built without -march, Makefile:2,14).
//   E2  dense * dense with a run-time or compile-time small size (rhs.rows + dst.rows + dst.cols < 20, GeneralMatrixMatrix.h;
//       CoeffBasedProductMode otherwise): "lazy" coefficient product, dst(i,j) = sum_k lhs(i,k) rhs(k,j), k ascending.
//   E3  dense(col-major) * compile-time column vector: GEMV kernel (GeneralMatrixVector.h), four columns at a time:
//       res_i += (l_i0 x0 + l_i1 x1) + (l_i2 x2 + l_i3 x3), leftover columns one by one.  A transposed (row-major) lhs would
//       take the other kernel; the reference never does that and this header refuses to compile it.
//   E4  Matrix2d::inverse(): cofactors times 1/det, det = a00 a11 - a10 a01 (InverseImpl.h, Determinant.h).
//   E5  dynamic inverse() / determinant(): PartialPivLU, unblocked for size <= 16 (first maximal |entry| is the pivot, true
//       division of the column by the pivot, rank-1 update a_ij -= l_i u_j), determinant = sign * (((d0 d1) d2) d3),
//       inverse = solve(P * I): unit-lower then upper triangular solve, column oriented, the upper diagonal applied as a
//       multiplication by 1 / u_ii (TriangularSolverMatrix.h, one panel for sizes <= 4).  Sizes above 4 abort: the
//       reference's hot path never inverts them (only the Hessian post-processing does).
// Scalar factors are NOT pulled out of coefficient-based products (3.3.7; 3.3.8 / 3.4 do that in eval_dynamic).
#pragma once
#include <cmath>
#include <algorithm>
#include <complex>
#include <cstddef>
#include <cstring>
#include <limits>
#include <string>
#include <cstdlib>
#include <iostream>
#include <type_traits>
#include <utility>
#include <vector>

namespace Eigen {

typedef std::ptrdiff_t Index;
enum { Dynamic = -1 };
enum { StreamPrecision = -1, FullPrecision = -2 };
enum { DontAlignCols = 1 };
struct IOFormat {
    template <class... A>
    IOFormat(A&&...) {}
};

template <class T, int R, int C>
class Matrix;
template <class T, int R, int C>
class TransposeView;

namespace shim {
constexpr int pick(int a, int b) { return a != Dynamic ? a : b; }
[[noreturn]] inline void die(const char* what) {
    std::cerr << "ggp_eigen_shim: " << what << std::endl;
    std::abort();
}

template <class T, int R, int C>
struct CommaInit {
    Matrix<T, R, C>& m;
    Index row, col, block_rows;
    CommaInit& operator,(T v) {
        if (col == m.cols()) { row += block_rows; col = 0; block_rows = 1; }
        if (row >= m.rows()) die("comma initialiser: too many coefficients");
        m(row, col++) = v;
        return *this;
    }
    template <class U, int R2, int C2>
    CommaInit& operator,(const Matrix<U, R2, C2>& o) {
        if (col == m.cols()) { row += block_rows; col = 0; block_rows = o.rows(); }
        if (row + o.rows() > m.rows() || col + o.cols() > m.cols()) die("comma initialiser: block does not fit");
        for (Index j = 0; j < o.cols(); ++j)
            for (Index i = 0; i < o.rows(); ++i) m(row + i, col + j) = o(i, j);
        col += o.cols();
        return *this;
    }
};

template <class T, int R, int C>
struct RowProxy {   // m.row(i) as an lvalue (moma_input.h:639, likelihood.h:21) and for .mean()
    Matrix<T, R, C>& m;
    Index i;
    T mean() const {
        T s = m(i, 0);
        for (Index j = 1; j < m.cols(); ++j) s = s + m(i, j);
        return s / T(m.cols());
    }
    template <int R2, int C2>
    RowProxy& operator-=(const Matrix<T, R2, C2>& o) {
        for (Index j = 0; j < m.cols(); ++j) m(i, j) = m(i, j) - o(0, j);
        return *this;
    }
    template <int R2, int C2>
    RowProxy& operator+=(const Matrix<T, R2, C2>& o) {
        for (Index j = 0; j < m.cols(); ++j) m(i, j) = m(i, j) + o(0, j);
        return *this;
    }
};
}  // namespace shim

template <class T, int R, int C>
class Matrix {
public:
    Index r_, c_;
    std::vector<T> d_;   // column major

    Matrix() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C), d_((size_t)(r_ * c_)) {}
    template <class I, class = std::enable_if_t<std::is_integral<I>::value>>
    explicit Matrix(I n) : r_(C == 1 ? (Index)n : (R == 1 ? 1 : (Index)n)), c_(C == 1 ? 1 : (R == 1 ? (Index)n : 1)), d_((size_t)(r_ * c_)) {
        static_assert(R == 1 || C == 1, "one-argument size constructor is for vectors");
    }
    template <class I, class J, class = std::enable_if_t<std::is_integral<I>::value && std::is_integral<J>::value>>
    Matrix(I r, J c) : r_((Index)r), c_((Index)c), d_((size_t)(r_ * c_)) {}
    Matrix(T a, T b, T c, T d) : r_(R == 1 ? 1 : 4), c_(R == 1 ? 4 : 1), d_{a, b, c, d} {
        static_assert((R == 4 && C == 1) || (R == 1 && C == 4), "four-coefficient constructor is Vector4");
    }
    Matrix(const Matrix&) = default;
    Matrix(Matrix&&) = default;
    Matrix& operator=(const Matrix&) = default;
    Matrix& operator=(Matrix&&) = default;
    template <class U, int R2, int C2>
    Matrix(const Matrix<U, R2, C2>& o) : r_(o.r_), c_(o.c_), d_(o.d_.begin(), o.d_.end()) {
        check_shape();
    }
    template <class U, int R2, int C2>
    Matrix& operator=(const Matrix<U, R2, C2>& o) {
        r_ = o.r_; c_ = o.c_;
        d_.assign(o.d_.begin(), o.d_.end());
        check_shape();
        return *this;
    }
    void check_shape() {
        if ((R != Dynamic && r_ != R) || (C != Dynamic && c_ != C)) {
            // Eigen lets a dynamic vector initialise a fixed one of the other orientation only when sizes agree; a column
            // into a dynamic matrix keeps its shape.  Anything else is a bug in how this stand-in is used.
            shim::die("shape mismatch in assignment");
        }
    }

    static Matrix Zero(Index n) { Matrix m(n); return m; }
    static Matrix Zero(Index r, Index c) { return Matrix(r, c); }
    static Matrix Constant(Index r, Index c, T v) {
        Matrix m(r, c);
        for (auto& x : m.d_) x = v;
        return m;
    }
    static Matrix Identity(Index r, Index c) {
        Matrix m(r, c);
        for (Index i = 0; i < (r < c ? r : c); ++i) m(i, i) = T(1);
        return m;
    }

    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    T& operator()(Index i, Index j) { return d_[(size_t)(i + j * r_)]; }
    const T& operator()(Index i, Index j) const { return d_[(size_t)(i + j * r_)]; }
    T& operator()(Index i) { return d_[(size_t)i]; }
    const T& operator()(Index i) const { return d_[(size_t)i]; }
    T& operator[](Index i) { return d_[(size_t)i]; }
    const T& operator[](Index i) const { return d_[(size_t)i]; }

    void conservativeResize(Index n) {
        static_assert(C == 1 || R == 1, "conservativeResize(n) is for vectors");
        d_.resize((size_t)n);
        if (C == 1) r_ = n; else c_ = n;
    }

    shim::CommaInit<T, R, C> operator<<(T v) {
        if (size() == 0) shim::die("comma initialiser on an empty matrix");
        (*this)(0, 0) = v;
        return shim::CommaInit<T, R, C>{*this, 0, 1, 1};
    }
    template <class U, int R2, int C2>
    shim::CommaInit<T, R, C> operator<<(const Matrix<U, R2, C2>& o) {
        if (o.rows() > r_ || o.cols() > c_) shim::die("comma initialiser: block does not fit");
        for (Index j = 0; j < o.cols(); ++j)
            for (Index i = 0; i < o.rows(); ++i) (*this)(i, j) = o(i, j);
        return shim::CommaInit<T, R, C>{*this, 0, o.cols(), o.rows()};
    }

    TransposeView<T, C, R> transpose() const;

    Matrix<T, Dynamic, Dynamic> block(Index i0, Index j0, Index nr, Index nc) const {
        Matrix<T, Dynamic, Dynamic> b(nr, nc);
        for (Index j = 0; j < nc; ++j)
            for (Index i = 0; i < nr; ++i) b(i, j) = (*this)(i0 + i, j0 + j);
        return b;
    }
    Matrix<T, Dynamic, Dynamic> topLeftCorner(Index nr, Index nc) const { return block(0, 0, nr, nc); }
    Matrix<T, Dynamic, Dynamic> topRightCorner(Index nr, Index nc) const { return block(0, c_ - nc, nr, nc); }
    Matrix<T, Dynamic, Dynamic> bottomLeftCorner(Index nr, Index nc) const { return block(r_ - nr, 0, nr, nc); }
    Matrix<T, Dynamic, Dynamic> bottomRightCorner(Index nr, Index nc) const { return block(r_ - nr, c_ - nc, nr, nc); }
    Matrix<T, Dynamic, 1> head(Index n) const {
        static_assert(C == 1, "head() is for column vectors");
        Matrix<T, Dynamic, 1> v(n);
        for (Index i = 0; i < n; ++i) v(i) = (*this)(i);
        return v;
    }
    Matrix<T, Dynamic, 1> tail(Index n) const {
        static_assert(C == 1, "tail() is for column vectors");
        Matrix<T, Dynamic, 1> v(n);
        for (Index i = 0; i < n; ++i) v(i) = (*this)(r_ - n + i);
        return v;
    }
    shim::RowProxy<T, R, C> row(Index i) { return shim::RowProxy<T, R, C>{*this, i}; }

    T mean() const {
        T s = d_[0];
        for (size_t i = 1; i < d_.size(); ++i) s = s + d_[i];
        return s / T(size());
    }

    template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>
    Matrix& operator*=(S s) { for (auto& x : d_) x = x * T(s); return *this; }
    template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>
    Matrix& operator/=(S s) { for (auto& x : d_) x = x / T(s); return *this; }
    template <int R2, int C2>
    Matrix& operator+=(const Matrix<T, R2, C2>& o) { for (size_t i = 0; i < d_.size(); ++i) d_[i] = d_[i] + o.d_[i]; return *this; }
    template <int R2, int C2>
    Matrix& operator-=(const Matrix<T, R2, C2>& o) { for (size_t i = 0; i < d_.size(); ++i) d_[i] = d_[i] - o.d_[i]; return *this; }

    // E4 / E5
    T determinant() const;
    Matrix inverse() const;
};

// result of transpose(): an evaluated matrix that remembers it is a row-major VIEW in Eigen, so that the one kernel whose
// arithmetic depends on the storage order (E3) can refuse it
template <class T, int R, int C>
class TransposeView : public Matrix<T, R, C> {
public:
    using Matrix<T, R, C>::Matrix;
    TransposeView() = default;
};

template <class T, int R, int C>
TransposeView<T, C, R> Matrix<T, R, C>::transpose() const {
    TransposeView<T, C, R> t;
    t.r_ = c_; t.c_ = r_;
    t.d_.resize(d_.size());
    for (Index j = 0; j < c_; ++j)
        for (Index i = 0; i < r_; ++i) t(j, i) = (*this)(i, j);
    return t;
}

typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 2, 1> Vector2d;

// ---- coefficient-wise nodes (E1) ---------------------------------------------------------------------------------
template <class T, int R1, int C1, int R2, int C2>
Matrix<T, shim::pick(R1, R2), shim::pick(C1, C2)> operator+(const Matrix<T, R1, C1>& a, const Matrix<T, R2, C2>& b) {
    if (a.rows() != b.rows() || a.cols() != b.cols()) shim::die("operator+: shape mismatch");
    Matrix<T, shim::pick(R1, R2), shim::pick(C1, C2)> r(a.rows(), a.cols());
    for (size_t i = 0; i < r.d_.size(); ++i) r.d_[i] = a.d_[i] + b.d_[i];
    return r;
}
template <class T, int R1, int C1, int R2, int C2>
Matrix<T, shim::pick(R1, R2), shim::pick(C1, C2)> operator-(const Matrix<T, R1, C1>& a, const Matrix<T, R2, C2>& b) {
    if (a.rows() != b.rows() || a.cols() != b.cols()) shim::die("operator-: shape mismatch");
    Matrix<T, shim::pick(R1, R2), shim::pick(C1, C2)> r(a.rows(), a.cols());
    for (size_t i = 0; i < r.d_.size(); ++i) r.d_[i] = a.d_[i] - b.d_[i];
    return r;
}
template <class T, int R, int C>
Matrix<T, R, C> operator-(const Matrix<T, R, C>& a) {
    Matrix<T, R, C> r(a.rows(), a.cols());
    for (size_t i = 0; i < r.d_.size(); ++i) r.d_[i] = -a.d_[i];
    return r;
}
template <class S, class T, int R, int C, class = std::enable_if_t<std::is_arithmetic<S>::value>>
Matrix<T, R, C> operator*(S s, const Matrix<T, R, C>& a) {
    Matrix<T, R, C> r(a.rows(), a.cols());
    for (size_t i = 0; i < r.d_.size(); ++i) r.d_[i] = T(s) * a.d_[i];
    return r;
}
template <class S, class T, int R, int C, class = std::enable_if_t<std::is_arithmetic<S>::value>>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, S s) {
    Matrix<T, R, C> r(a.rows(), a.cols());
    for (size_t i = 0; i < r.d_.size(); ++i) r.d_[i] = a.d_[i] * T(s);
    return r;
}
template <class S, class T, int R, int C, class = std::enable_if_t<std::is_arithmetic<S>::value>>
Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, S s) {
    Matrix<T, R, C> r(a.rows(), a.cols());
    for (size_t i = 0; i < r.d_.size(); ++i) r.d_[i] = a.d_[i] / T(s);
    return r;
}

// ---- products (E2, E3) -------------------------------------------------------------------------------------------
template <class T, int R1, int C1, int R2, int C2>
Matrix<T, R1, C2> operator*(const Matrix<T, R1, C1>& a, const Matrix<T, R2, C2>& b) {
    if (a.cols() != b.rows()) shim::die("operator*: inner dimensions differ");
    const Index n = a.rows(), m = b.cols(), depth = a.cols();
    Matrix<T, R1, C2> r(n, m);
    if (depth == 0) return r;
    if (C2 == 1 && R1 != 1) {
        // E3: column-major GEMV, res starts at zero, four columns at a time, then the leftover columns one by one
        Index k = 0;
        for (; k + 4 <= depth; k += 4)
            for (Index i = 0; i < n; ++i)
                r(i, 0) = r(i, 0) + ((a(i, k) * b(k, 0) + a(i, k + 1) * b(k + 1, 0)) + (a(i, k + 2) * b(k + 2, 0) + a(i, k + 3) * b(k + 3, 0)));
        for (; k < depth; ++k)
            for (Index i = 0; i < n; ++i) r(i, 0) = r(i, 0) + a(i, k) * b(k, 0);
        return r;
    }
    if (b.rows() + n + m >= 20) shim::die("operator*: a product of this size would take Eigen's GEMM path (not restated)");
    // E2: lazy coefficient product
    for (Index j = 0; j < m; ++j)
        for (Index i = 0; i < n; ++i) {
            T s = a(i, 0) * b(0, j);
            for (Index k = 1; k < depth; ++k) s = s + a(i, k) * b(k, j);
            r(i, j) = s;
        }
    return r;
}
// a transposed lhs in front of a compile-time vector would run Eigen's row-major GEMV (different summation): not restated
template <class T, int R1, int C1, int R2>
Matrix<T, R1, 1> operator*(const TransposeView<T, R1, C1>&, const Matrix<T, R2, 1>&) = delete;

// ---- E4 / E5 -----------------------------------------------------------------------------------------------------
namespace shim {
template <class T>
struct Lu {   // PartialPivLU::compute for size <= 16 (unblocked_lu)
    Index n;
    std::vector<T> lu;      // column major
    std::vector<Index> tr;  // row transpositions
    int sign = 1;
    T& at(Index i, Index j) { return lu[(size_t)(i + j * n)]; }
    template <int R, int C>
    explicit Lu(const Matrix<T, R, C>& A) : n(A.rows()), lu(A.d_), tr((size_t)A.rows()) {
        if (A.rows() != A.cols()) die("LU of a non-square matrix");
        if (n > 4) die("LU of a matrix larger than 4x4: Eigen's blocked triangular solve is not restated");
        for (Index k = 0; k < n; ++k) {
            Index piv = k;
            T big = std::fabs(at(k, k));
            for (Index i = k + 1; i < n; ++i) {
                const T v = std::fabs(at(i, k));
                if (v > big) { big = v; piv = i; }
            }
            tr[(size_t)k] = piv;
            if (big != T(0)) {
                if (piv != k) {
                    for (Index j = 0; j < n; ++j) std::swap(at(k, j), at(piv, j));
                    sign = -sign;
                }
                for (Index i = k + 1; i < n; ++i) at(i, k) = at(i, k) / at(k, k);
            }
            for (Index j = k + 1; j < n; ++j)
                for (Index i = k + 1; i < n; ++i) at(i, j) = at(i, j) - at(i, k) * at(k, j);
        }
    }
    T determinant() {
        T p = at(0, 0);
        for (Index i = 1; i < n; ++i) p = p * at(i, i);
        return T(sign) * p;
    }
    void inverse(std::vector<T>& out) {
        out.assign((size_t)(n * n), T(0));
        auto o = [&](Index i, Index j) -> T& { return out[(size_t)(i + j * n)]; };
        for (Index i = 0; i < n; ++i) o(i, i) = T(1);
        for (Index k = 0; k < n; ++k)   // dst = P * I: the transpositions in order
            if (tr[(size_t)k] != k)
                for (Index j = 0; j < n; ++j) std::swap(o(k, j), o(tr[(size_t)k], j));
        for (Index j = 0; j < n; ++j) {
            for (Index k = 0; k < n; ++k) {   // unit lower
                const T b = o(k, j);
                for (Index i = k + 1; i < n; ++i) o(i, j) = o(i, j) - b * at(i, k);
            }
            for (Index i = n - 1; i >= 0; --i) {   // upper
                const T a = T(1) / at(i, i);
                const T b = (o(i, j) = o(i, j) * a);
                for (Index s = 0; s < i; ++s) o(s, j) = o(s, j) - b * at(s, i);
            }
        }
    }
};
}  // namespace shim

template <class T, int R, int C>
T Matrix<T, R, C>::determinant() const {
    if (R == 2 && C == 2) return (*this)(0, 0) * (*this)(1, 1) - (*this)(1, 0) * (*this)(0, 1);
    static_assert((R == 2 && C == 2) || (R == Dynamic && C == Dynamic), "determinant(): only Matrix2d and MatrixXd are restated");
    if (r_ == 0) return T(1);
    shim::Lu<T> lu(*this);
    return lu.determinant();
}

template <class T, int R, int C>
Matrix<T, R, C> Matrix<T, R, C>::inverse() const {
    static_assert((R == 2 && C == 2) || (R == Dynamic && C == Dynamic), "inverse(): only Matrix2d and MatrixXd are restated");
    Matrix<T, R, C> out(r_, c_);
    if (R == 2 && C == 2) {
        const T invdet = T(1) / determinant();
        out(0, 0) = (*this)(1, 1) * invdet;
        out(1, 0) = -(*this)(1, 0) * invdet;
        out(0, 1) = -(*this)(0, 1) * invdet;
        out(1, 1) = (*this)(0, 0) * invdet;
        return out;
    }
    shim::Lu<T> lu(*this);
    lu.inverse(out.d_);
    return out;
}

template <class T, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& m) {
    for (Index i = 0; i < m.rows(); ++i) {
        for (Index j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

}  // namespace Eigen
