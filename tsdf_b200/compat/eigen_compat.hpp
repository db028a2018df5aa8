// eigen_compat.hpp — the small subset of Eigen's dense API that the TSDF class surface uses.
//
// The reference keeps K, K^-1, pose and pose^-1 in Eigen matrices (src/include/Camera.hpp:20-29) and
// returns raycast results in Eigen::Matrix<float,3,Dynamic> (src/include/TSDFVolume.hpp:261); kinfu.cpp
// touches Matrix4f, Vector3f, the comma initialiser and operator()(i,j).  Eigen is a system package the
// reference does not vendor and this image does not have, so the drop-in headers ship this stand-in.
// It is NOT Eigen: column-major storage, fixed sizes up to 4x4 held in-object (like Eigen — callers keep
// .data() pointers of temporaries alive only as long as Eigen would), one Dynamic dimension (columns),
// eager evaluation, no expression templates, no alignment tricks.
//
// If the real Eigen is on the include path first, none of this is used.
#ifndef TSDF_B200_EIGEN_COMPAT_HPP
#define TSDF_B200_EIGEN_COMPAT_HPP

#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <limits>
#include <ostream>
#include <type_traits>
#include <cstdlib>
#include <new>
#include <utility>
#include <vector>

namespace Eigen {

const int Dynamic = -1;
typedef std::ptrdiff_t Index;

template <typename T, int R, int C> class Matrix;

namespace compat {
// Where the storage of Dynamic matrices comes from.  The drop-in class layer is built with TSDF_B200_PINNED_EIGEN: large blocks
// then come from the library's pool of pinned host memory (tsdf_b200_host_alloc), so that the result matrices a caller
// declares per frame — kinfu.cpp does — are buffers the device can write while the raycast runs (tsdf_b200_volume_raycast
// mirrors both maps into pinned buffers; a pageable buffer gets a staged copy afterwards, at half the frame rate).  Like
// Eigen's own, the storage is NOT zero-filled by resize().
#ifdef TSDF_B200_PINNED_EIGEN
extern "C" void *tsdf_b200_host_alloc(size_t bytes);
extern "C" void tsdf_b200_host_free(void *p);
inline void *raw_alloc(size_t bytes) { return tsdf_b200_host_alloc(bytes); }
inline void raw_free(void *p) { tsdf_b200_host_free(p); }
#else
inline void *raw_alloc(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
inline void raw_free(void *p) { std::free(p); }
#endif
template <typename T> struct HostAllocator {
    typedef T value_type;
    HostAllocator() {}
    template <typename U> HostAllocator(const HostAllocator<U> &) {}
    T *allocate(size_t n) { void *p = raw_alloc(n * sizeof(T)); if (!p) throw std::bad_alloc(); return static_cast<T *>(p); }
    void deallocate(T *p, size_t) { raw_free(p); }
    template <typename U> void construct(U *p) { ::new (static_cast<void *>(p)) U; }          // default-initialised: no fill
    template <typename U, typename... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
    template <typename U> bool operator==(const HostAllocator<U> &) const { return true; }
    template <typename U> bool operator!=(const HostAllocator<U> &) const { return false; }
};
// Storage: in-object array for fixed sizes, std::vector when the column count is Dynamic.
template <typename T, int R, int C> struct Storage {
    T v[R * C];
    Storage() { for (int i = 0; i < R * C; i++) v[i] = T(0); }
    T *ptr() { return v; }
    const T *ptr() const { return v; }
    Index rows() const { return R; }
    Index cols() const { return C; }
    void resize(Index r, Index c) { assert(r == R && c == C); (void)r; (void)c; }
};
template <typename T, int R> struct Storage<T, R, Dynamic> {
    std::vector<T, HostAllocator<T> > v;
    Index n_cols;
    Storage() : n_cols(0) {}
    T *ptr() { return v.data(); }
    const T *ptr() const { return v.data(); }
    Index rows() const { return R; }
    Index cols() const { return n_cols; }
    void resize(Index r, Index c) { assert(r == R); (void)r; n_cols = c; v.resize((size_t)R * (size_t)c); }
};

// Result of block()/corner accessors: a small owning copy (at most 4x4) plus, for writable blocks,
// the location to write back to.
template <typename T> struct Block {
    T v[16];
    int r, c;
    T *dst;          // column-major parent storage or nullptr
    Index ld;        // parent rows
    Block(int rows, int cols) : r(rows), c(cols), dst(nullptr), ld(0) { for (int i = 0; i < 16; i++) v[i] = T(0); }
    T &at(int i, int j) { return v[j * r + i]; }
    const T &at(int i, int j) const { return v[j * r + i]; }
    Index rows() const { return r; }
    Index cols() const { return c; }
    T operator()(int i, int j) const { return at(i, j); }
    T operator()(int i) const { return v[i]; }
    T x() const { return v[0]; }
    T y() const { return v[1]; }
    T z() const { return v[2]; }
    template <int R2, int C2> Block &operator=(const Matrix<T, R2, C2> &m);
    Block &operator=(const Block &o) {
        assert(r == o.r && c == o.c);
        for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) { at(i, j) = o.at(i, j); if (dst) dst[j * ld + i] = o.at(i, j); }
        return *this;
    }
    template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    Block operator/(S s) const { Block b(r, c); for (int i = 0; i < r * c; i++) b.v[i] = v[i] / T(s); return b; }
    template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    Block operator*(S s) const { Block b(r, c); for (int i = 0; i < r * c; i++) b.v[i] = v[i] * T(s); return b; }
    template <int R2, int C2> Block operator*(const Matrix<T, R2, C2> &m) const;
};

template <typename T, int R, int C> struct CommaInit {
    Matrix<T, R, C> &m;
    Index n;
    CommaInit(Matrix<T, R, C> &mat, T first) : m(mat), n(0) { put(first); }
    void put(T x) { Index rows = m.rows(), cols = m.cols(); assert(n < rows * cols); m(n / cols, n % cols) = x; n++; }   // row-major fill
    template <typename S> CommaInit &operator,(S x) { put(T(x)); return *this; }
};
}  // namespace compat

template <typename T, int R, int C>
class Matrix {
    compat::Storage<T, R, C> s_;
    enum { IsVector = (C == 1 || R == 1) };

public:
    typedef T Scalar;
    Matrix() {}
    // Vector constructors: Vector2f{x,y}, Vector3f{x,y,z}, Vector4f{x,y,z,w}.
    Matrix(T x, T y) { static_assert(R * C == 2, "2-vector only"); s_.v[0] = x; s_.v[1] = y; }
    Matrix(T x, T y, T z) { static_assert(R * C == 3, "3-vector only"); s_.v[0] = x; s_.v[1] = y; s_.v[2] = z; }
    Matrix(T x, T y, T z, T w) { static_assert(R * C == 4, "4-vector only"); s_.v[0] = x; s_.v[1] = y; s_.v[2] = z; s_.v[3] = w; }
    Matrix(const compat::Block<T> &b) { assign(b); }
    Matrix &operator=(const compat::Block<T> &b) { assign(b); return *this; }

    Index rows() const { return s_.rows(); }
    Index cols() const { return s_.cols(); }
    Index size() const { return rows() * cols(); }
    void resize(Index r, Index c) { s_.resize(r, c); }
    T *data() { return s_.ptr(); }
    const T *data() const { return s_.ptr(); }

    T &operator()(Index i, Index j) { return s_.ptr()[j * rows() + i]; }
    const T &operator()(Index i, Index j) const { return s_.ptr()[j * rows() + i]; }
    T &operator()(Index i) { return s_.ptr()[i]; }
    const T &operator()(Index i) const { return s_.ptr()[i]; }
    T &operator[](Index i) { return s_.ptr()[i]; }
    const T &operator[](Index i) const { return s_.ptr()[i]; }
    T &x() { return s_.ptr()[0]; }
    T &y() { return s_.ptr()[1]; }
    T &z() { return s_.ptr()[2]; }
    T &w() { return s_.ptr()[3]; }
    const T &x() const { return s_.ptr()[0]; }
    const T &y() const { return s_.ptr()[1]; }
    const T &z() const { return s_.ptr()[2]; }
    const T &w() const { return s_.ptr()[3]; }

    static Matrix Zero() { Matrix m; for (Index i = 0; i < m.size(); i++) m.data()[i] = T(0); return m; }
    static Matrix Identity() { Matrix m = Zero(); for (Index i = 0; i < (R < C ? R : C); i++) m(i, i) = T(1); return m; }
    static Matrix Constant(T c) { Matrix m; for (Index i = 0; i < m.size(); i++) m.data()[i] = c; return m; }
    void setZero() { for (Index i = 0; i < size(); i++) data()[i] = T(0); }
    void setIdentity() { *this = Identity(); }

    template <typename S> compat::CommaInit<T, R, C> operator<<(S first) { return compat::CommaInit<T, R, C>(*this, T(first)); }

    // ---- element-wise arithmetic -------------------------------------------------------------
    Matrix operator+(const Matrix &o) const { Matrix m(*this); for (Index i = 0; i < size(); i++) m.data()[i] += o.data()[i]; return m; }
    Matrix operator-(const Matrix &o) const { Matrix m(*this); for (Index i = 0; i < size(); i++) m.data()[i] -= o.data()[i]; return m; }
    Matrix operator-() const { Matrix m(*this); for (Index i = 0; i < size(); i++) m.data()[i] = -m.data()[i]; return m; }
    Matrix &operator+=(const Matrix &o) { for (Index i = 0; i < size(); i++) data()[i] += o.data()[i]; return *this; }
    Matrix &operator-=(const Matrix &o) { for (Index i = 0; i < size(); i++) data()[i] -= o.data()[i]; return *this; }
    template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    Matrix operator*(S k) const { Matrix m(*this); for (Index i = 0; i < size(); i++) m.data()[i] *= T(k); return m; }
    template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    Matrix operator/(S k) const { Matrix m(*this); for (Index i = 0; i < size(); i++) m.data()[i] /= T(k); return m; }
    template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    Matrix &operator*=(S k) { for (Index i = 0; i < size(); i++) data()[i] *= T(k); return *this; }
    template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    Matrix &operator/=(S k) { for (Index i = 0; i < size(); i++) data()[i] /= T(k); return *this; }
    bool operator==(const Matrix &o) const {
        if (rows() != o.rows() || cols() != o.cols()) return false;
        for (Index i = 0; i < size(); i++) if (!(data()[i] == o.data()[i])) return false;
        return true;
    }
    bool operator!=(const Matrix &o) const { return !(*this == o); }

    // ---- products -----------------------------------------------------------------------------
    template <int C2>
    Matrix<T, R, C2> operator*(const Matrix<T, C, C2> &o) const {
        static_assert(R != Dynamic && C != Dynamic && C2 != Dynamic, "fixed-size products only");
        Matrix<T, R, C2> m;
        for (int i = 0; i < R; i++)
            for (int j = 0; j < C2; j++) {
                T acc = T(0);
                for (int k = 0; k < C; k++) acc += (*this)(i, k) * o(k, j);
                m(i, j) = acc;
            }
        return m;
    }
    T dot(const Matrix &o) const { T acc = T(0); for (Index i = 0; i < size(); i++) acc += data()[i] * o.data()[i]; return acc; }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }
    void normalize() { T n = norm(); for (Index i = 0; i < size(); i++) data()[i] /= n; }
    Matrix normalized() const { Matrix m(*this); m.normalize(); return m; }
    Matrix cross(const Matrix &o) const {
        static_assert(R * C == 3, "cross product of 3-vectors only");
        const T *a = data(), *b = o.data();
        Matrix m;
        m[0] = a[1] * b[2] - a[2] * b[1];
        m[1] = a[2] * b[0] - a[0] * b[2];
        m[2] = a[0] * b[1] - a[1] * b[0];
        return m;
    }
    Matrix<T, C, R> transpose() const {
        static_assert(R != Dynamic && C != Dynamic, "fixed-size transpose only");
        Matrix<T, C, R> m;
        for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) m(j, i) = (*this)(i, j);
        return m;
    }
    T determinant() const;
    Matrix inverse() const;

    // ---- blocks -------------------------------------------------------------------------------
    compat::Block<T> block(Index i0, Index j0, Index p, Index q) {
        compat::Block<T> b = const_cast<const Matrix *>(this)->block(i0, j0, p, q);
        b.dst = data() + j0 * rows() + i0;
        b.ld = rows();
        return b;
    }
    compat::Block<T> block(Index i0, Index j0, Index p, Index q) const {
        assert(p <= 4 && q <= 4 && i0 + p <= rows() && j0 + q <= cols());
        compat::Block<T> b((int)p, (int)q);
        for (Index j = 0; j < q; j++) for (Index i = 0; i < p; i++) b.at((int)i, (int)j) = (*this)(i0 + i, j0 + j);
        return b;
    }
    template <int P, int Q> compat::Block<T> block(Index i0, Index j0) { return block(i0, j0, P, Q); }
    template <int P, int Q> compat::Block<T> block(Index i0, Index j0) const { return block(i0, j0, P, Q); }
    compat::Block<T> topLeftCorner(Index p, Index q) { return block(0, 0, p, q); }
    compat::Block<T> topLeftCorner(Index p, Index q) const { return block(0, 0, p, q); }
    compat::Block<T> topRightCorner(Index p, Index q) { return block(0, cols() - q, p, q); }
    compat::Block<T> topRightCorner(Index p, Index q) const { return block(0, cols() - q, p, q); }
    template <int P, int Q> compat::Block<T> topLeftCorner() { return block(0, 0, P, Q); }
    template <int P, int Q> compat::Block<T> topLeftCorner() const { return block(0, 0, P, Q); }
    template <int P, int Q> compat::Block<T> topRightCorner() { return block(0, cols() - Q, P, Q); }
    template <int P, int Q> compat::Block<T> topRightCorner() const { return block(0, cols() - Q, P, Q); }
    compat::Block<T> col(Index j) { return block(0, j, rows(), 1); }
    compat::Block<T> col(Index j) const { return block(0, j, rows(), 1); }
    compat::Block<T> head(Index n) const { return block(0, 0, n, 1); }

private:
    void assign(const compat::Block<T> &b) {
        if (C == Dynamic) resize(b.r, b.c);
        assert(b.r == rows() && b.c == cols());
        for (int j = 0; j < b.c; j++) for (int i = 0; i < b.r; i++) (*this)(i, j) = b.at(i, j);
    }
};

template <typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator*(S k, const Matrix<T, R, C> &m) { return m * k; }

namespace compat {
template <typename T> template <int R2, int C2>
Block<T> &Block<T>::operator=(const Matrix<T, R2, C2> &m) {
    assert(m.rows() == r && m.cols() == c);
    for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) { at(i, j) = m(i, j); if (dst) dst[j * ld + i] = m(i, j); }
    return *this;
}
template <typename T> template <int R2, int C2>
Block<T> Block<T>::operator*(const Matrix<T, R2, C2> &m) const {
    assert(c == m.rows() && m.cols() <= 4);
    Block<T> out(r, (int)m.cols());
    for (int i = 0; i < r; i++)
        for (int j = 0; j < (int)m.cols(); j++) {
            T acc = T(0);
            for (int k = 0; k < c; k++) acc += at(i, k) * m(k, j);
            out.at(i, j) = acc;
        }
    return out;
}

// Determinant / inverse by cofactor expansion in T (what Eigen does for sizes <= 4).
template <typename T> T det3(const T *m) {   // column-major 3x3
    return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[3] * (m[1] * m[8] - m[7] * m[2]) + m[6] * (m[1] * m[5] - m[4] * m[2]);
}
}  // namespace compat

template <typename T, int R, int C>
T Matrix<T, R, C>::determinant() const {
    static_assert(R == C && R >= 1 && R <= 4, "determinant of 1x1..4x4 only");
    const T *m = data();
    if (R == 1) return m[0];
    if (R == 2) return m[0] * m[3] - m[2] * m[1];
    if (R == 3) return compat::det3(m);
    T det = T(0);
    for (int j = 0; j < 4; j++) {   // expand along row 0
        T sub[9];
        int n = 0;
        for (int c = 0; c < 4; c++) { if (c == j) continue; for (int r = 1; r < 4; r++) sub[n++] = m[c * 4 + r]; }
        T cof = compat::det3(sub);
        det += ((j & 1) ? -m[j * 4] : m[j * 4]) * cof;
    }
    return det;
}

template <typename T, int R, int C>
Matrix<T, R, C> Matrix<T, R, C>::inverse() const {
    static_assert(R == C && R >= 1 && R <= 4, "inverse of 1x1..4x4 only");
    Matrix inv;
    const T det = determinant();
    if (R == 1) { inv(0, 0) = T(1) / (*this)(0, 0); return inv; }
    if (R == 2) {
        inv(0, 0) = (*this)(1, 1) / det; inv(0, 1) = -(*this)(0, 1) / det;
        inv(1, 0) = -(*this)(1, 0) / det; inv(1, 1) = (*this)(0, 0) / det;
        return inv;
    }
    // adjugate: inv(j,i) = cofactor(i,j) / det
    for (int i = 0; i < R; i++)
        for (int j = 0; j < R; j++) {
            T sub[9];
            int n = 0;
            for (int c = 0; c < R; c++) { if (c == j) continue; for (int r = 0; r < R; r++) { if (r == i) continue; sub[n++] = (*this)(r, c); } }
            T cof = (R == 3) ? (sub[0] * sub[3] - sub[2] * sub[1]) : compat::det3(sub);
            inv(j, i) = (((i + j) & 1) ? -cof : cof) / det;
        }
    return inv;
}

template <typename T, int R, int C>
std::ostream &operator<<(std::ostream &os, const Matrix<T, R, C> &m) {
    for (Index i = 0; i < m.rows(); i++) {
        for (Index j = 0; j < m.cols(); j++) os << (j ? " " : "") << m(i, j);
        if (i + 1 < m.rows()) os << "\n";
    }
    return os;
}

typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, Dynamic> Matrix3Xf;

}  // namespace Eigen
#endif  // TSDF_B200_EIGEN_COMPAT_HPP
