// mini_eigen.h — TEST INFRASTRUCTURE (oracle/_ref build only; never part of libctrfk.so).
//
// The reference (RViMLab/RAM2017-CTR-kinematics) includes <Eigen/Core>, <Eigen/Dense> (Eigen 3.3.3,
// install.sh:49), which this image does not have.  This header is a stand-in written from Eigen's
// public API, covering exactly the call sites of the reference's forward-kinematics path:
// dense storage (fixed and dynamic, column-major, with the object layout {data pointer, rows, cols}
// that ctr_source/matrix.h:139-177 overwrites by memcpy), coefficient access, .array().sin()/.cos()
// (for double Eigen 3.3.3 has no vectorised sin/cos: they are std::sin/std::cos per coefficient),
// topRows/col/leftCols views (robot_i.h:702-703), corner blocks (robot_i.h:956-959), Ref, Zero,
// conservativeResize, the comma initialiser and setLinSpaced (transrotscale.h).
// No arithmetic of the reference happens inside Eigen other than those per-coefficient calls.
#ifndef CTRB200_MINI_EIGEN_H
#define CTRB200_MINI_EIGEN_H
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <new>
#include <ostream>
#include <type_traits>

#define EIGEN_ALWAYS_INLINE __attribute__((always_inline)) inline
#define EIGEN_STRONG_INLINE inline
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3
#define EIGEN_MINOR_VERSION 3

namespace Eigen {
typedef std::ptrdiff_t DenseIndex;
typedef DenseIndex Index;
const int Dynamic = -1;
enum StorageOptions { ColMajor = 0, RowMajor = 1 };

// ------------------------------------------------------------------ coefficient-wise expressions
template <class E> struct ArrExpr;
template <class E> struct SinOp;
template <class E> struct CosOp;
template <class E> struct AddScalarOp;

template <class E> struct ArrExpr {
    const E& derived() const { return *static_cast<const E*>(this); }
    SinOp<E> sin() const { return SinOp<E>{derived()}; }
    CosOp<E> cos() const { return CosOp<E>{derived()}; }
};
template <class E> struct SinOp : ArrExpr<SinOp<E>> {
    typedef typename E::Scalar Scalar;
    E e;
    explicit SinOp(const E& x) : e(x) {}
    Index size() const { return e.size(); }
    Scalar coeff(Index i) const { return std::sin(e.coeff(i)); }
};
template <class E> struct CosOp : ArrExpr<CosOp<E>> {
    typedef typename E::Scalar Scalar;
    E e;
    explicit CosOp(const E& x) : e(x) {}
    Index size() const { return e.size(); }
    Scalar coeff(Index i) const { return std::cos(e.coeff(i)); }
};
template <class E> struct AddScalarOp : ArrExpr<AddScalarOp<E>> {
    typedef typename E::Scalar Scalar;
    E e;
    Scalar s;
    AddScalarOp(const E& x, Scalar v) : e(x), s(v) {}
    Index size() const { return e.size(); }
    Scalar coeff(Index i) const { return e.coeff(i) + s; }
};
template <class E> AddScalarOp<E> operator+(const ArrExpr<E>& a, typename E::Scalar s) { return AddScalarOp<E>(a.derived(), s); }

// strided view of a vector's coefficients (what .array() returns)
template <class T> struct ArrView : ArrExpr<ArrView<T>> {
    typedef typename std::remove_const<T>::type Scalar;
    T* p;
    Index n, stride;
    ArrView(T* p_, Index n_, Index s_) : p(p_), n(n_), stride(s_) {}
    Index size() const { return n; }
    Scalar coeff(Index i) const { return p[i * stride]; }
    template <class E> ArrView& operator=(const ArrExpr<E>& x)
    {
        const E& e = x.derived();
        assert(e.size() == n);
        for (Index i = 0; i < n; ++i) p[i * stride] = e.coeff(i);
        return *this;
    }
    ArrView& operator=(const ArrView& o) { return operator=(static_cast<const ArrExpr<ArrView<T>>&>(o)); }
};

// ------------------------------------------------------------------ views
template <class T> struct VecView {
    T* p;
    Index n;
    T* data() const { return p; }
    Index size() const { return n; }
    Index rows() const { return n; }
    T& operator[](Index i) const { return p[i]; }
    T& operator()(Index i) const { return p[i]; }
    void setZero() const { for (Index i = 0; i < n; ++i) p[i] = T(0); }
    ArrView<T> array() const { return ArrView<T>(p, n, 1); }
};
template <class T> struct BlockView {   // rows x cols window of a column-major matrix
    T* p;
    Index r, c, outer;
    Index rows() const { return r; }
    Index cols() const { return c; }
    VecView<T> col(Index j) const { assert(j < c); return VecView<T>{p + j * outer, r}; }
    template <int N> BlockView leftCols() const { assert(N <= c); return BlockView{p, r, N, outer}; }
    ArrView<T> array() const { assert(c == 1); return ArrView<T>(p, r, 1); }
    T& operator()(Index i, Index j) const { return p[i + j * outer]; }
};
template <class T, int R, int C> struct FixedBlock {
    T* p;
    Index outer;
    template <class V> FixedBlock& operator=(const V& v)      // from anything with data() holding R*C values
    {
        for (int j = 0; j < C; ++j)
            for (int i = 0; i < R; ++i) p[i + j * outer] = v.data()[i + j * R];
        return *this;
    }
    T& operator()(Index i, Index j) const { return p[i + j * outer]; }
};
template <class T> struct PrintView {
    const T* p;
    Index r, c;
    bool transposed;
};
template <class T> std::ostream& operator<<(std::ostream& os, const PrintView<T>& v)
{
    const Index R = v.transposed ? v.c : v.r, Cc = v.transposed ? v.r : v.c;
    for (Index i = 0; i < R; ++i) {
        for (Index j = 0; j < Cc; ++j) os << (j ? " " : "") << (v.transposed ? v.p[j + i * v.r] : v.p[i + j * v.r]);
        if (i + 1 < R) os << "\n";
    }
    return os;
}

// ------------------------------------------------------------------ storage
namespace internal {
template <class T> T* alloc_n(Index n)
{
    if (n <= 0) return nullptr;
    void* q = nullptr;
    if (posix_memalign(&q, 32, sizeof(T) * size_t(n))) throw std::bad_alloc();
    T* t = static_cast<T*>(q);
    for (Index i = 0; i < n; ++i) new (t + i) T();
    return t;
}
template <class T> void free_n(T* p, Index n)
{
    if (!p) return;
    for (Index i = 0; i < n; ++i) p[i].~T();
    std::free(p);
}
template <class T, int R, int C> struct Storage {                       // fixed size
    T m_data[(R * C) > 0 ? R * C : 1];
    Storage() {}
    T* ptr() { return m_data; }
    const T* ptr() const { return m_data; }
    Index rows() const { return R; }
    Index cols() const { return C; }
    void resize(Index r, Index c) { assert(r == R && c == C); (void)r; (void)c; }
    void conservative(Index r, Index c) { resize(r, c); }
};
template <class T, int C> struct Storage<T, Dynamic, C> {               // {pointer, rows}
    T* m_data;
    Index m_rows;
    Storage() : m_data(nullptr), m_rows(0) {}
    Storage(const Storage& o) : m_data(alloc_n<T>(o.m_rows * C)), m_rows(o.m_rows) { for (Index i = 0; i < m_rows * C; ++i) m_data[i] = o.m_data[i]; }
    Storage& operator=(const Storage& o) { if (this != &o) { resize(o.m_rows, C); for (Index i = 0; i < m_rows * C; ++i) m_data[i] = o.m_data[i]; } return *this; }
    ~Storage() { free_n(m_data, m_rows * C); }
    T* ptr() { return m_data; }
    const T* ptr() const { return m_data; }
    Index rows() const { return m_rows; }
    Index cols() const { return C; }
    void resize(Index r, Index c) { assert(c == C); (void)c; if (r != m_rows) { free_n(m_data, m_rows * C); m_data = alloc_n<T>(r * C); m_rows = r; } }
    void conservative(Index r, Index c)
    {
        assert(c == C); (void)c;
        if (r == m_rows) return;
        T* q = alloc_n<T>(r * C);
        for (Index j = 0; j < C; ++j) for (Index i = 0; i < (r < m_rows ? r : m_rows); ++i) q[i + j * r] = m_data[i + j * m_rows];
        free_n(m_data, m_rows * C); m_data = q; m_rows = r;
    }
};
template <class T, int R> struct Storage<T, R, Dynamic> {               // {pointer, cols}
    T* m_data;
    Index m_cols;
    Storage() : m_data(nullptr), m_cols(0) {}
    Storage(const Storage& o) : m_data(alloc_n<T>(o.m_cols * R)), m_cols(o.m_cols) { for (Index i = 0; i < m_cols * R; ++i) m_data[i] = o.m_data[i]; }
    Storage& operator=(const Storage& o) { if (this != &o) { resize(R, o.m_cols); for (Index i = 0; i < m_cols * R; ++i) m_data[i] = o.m_data[i]; } return *this; }
    ~Storage() { free_n(m_data, m_cols * R); }
    T* ptr() { return m_data; }
    const T* ptr() const { return m_data; }
    Index rows() const { return R; }
    Index cols() const { return m_cols; }
    void resize(Index r, Index c) { assert(r == R); (void)r; if (c != m_cols) { free_n(m_data, m_cols * R); m_data = alloc_n<T>(c * R); m_cols = c; } }
    void conservative(Index r, Index c)
    {
        assert(r == R); (void)r;
        if (c == m_cols) return;
        T* q = alloc_n<T>(c * R);
        for (Index i = 0; i < R * (c < m_cols ? c : m_cols); ++i) q[i] = m_data[i];
        free_n(m_data, m_cols * R); m_data = q; m_cols = c;
    }
};
template <class T> struct Storage<T, Dynamic, Dynamic> {                // {pointer, rows, cols}
    T* m_data;
    Index m_rows, m_cols;
    Storage() : m_data(nullptr), m_rows(0), m_cols(0) {}
    Storage(const Storage& o) : m_data(alloc_n<T>(o.m_rows * o.m_cols)), m_rows(o.m_rows), m_cols(o.m_cols) { for (Index i = 0; i < m_rows * m_cols; ++i) m_data[i] = o.m_data[i]; }
    Storage& operator=(const Storage& o) { if (this != &o) { resize(o.m_rows, o.m_cols); for (Index i = 0; i < m_rows * m_cols; ++i) m_data[i] = o.m_data[i]; } return *this; }
    ~Storage() { free_n(m_data, m_rows * m_cols); }
    T* ptr() { return m_data; }
    const T* ptr() const { return m_data; }
    Index rows() const { return m_rows; }
    Index cols() const { return m_cols; }
    void resize(Index r, Index c) { if (r * c != m_rows * m_cols) { free_n(m_data, m_rows * m_cols); m_data = alloc_n<T>(r * c); } m_rows = r; m_cols = c; }
    void conservative(Index r, Index c)
    {
        T* q = alloc_n<T>(r * c);
        for (Index j = 0; j < (c < m_cols ? c : m_cols); ++j) for (Index i = 0; i < (r < m_rows ? r : m_rows); ++i) q[i + j * r] = m_data[i + j * m_rows];
        free_n(m_data, m_rows * m_cols); m_data = q; m_rows = r; m_cols = c;
    }
};
}  // namespace internal

// ------------------------------------------------------------------ MatrixBase / Matrix
template <class Derived> struct MatrixBase {
    Derived& derived() { return *static_cast<Derived*>(this); }
    const Derived& derived() const { return *static_cast<const Derived*>(this); }
    auto& operator[](Index i) { return derived().data()[i]; }
    const auto& operator[](Index i) const { return derived().data()[i]; }
    auto& operator()(Index i) { return derived().data()[i]; }
    const auto& operator()(Index i) const { return derived().data()[i]; }
};

template <class M> struct CommaInit {
    M& m;
    Index k;
    CommaInit& operator,(typename M::Scalar v) { assert(k < m.size()); m.data()[k++] = v; return *this; }
};

template <class T, int R, int C, int Opt = ColMajor, int MaxR = R, int MaxC = C>
class Matrix : public MatrixBase<Matrix<T, R, C, Opt, MaxR, MaxC>> {
    internal::Storage<T, R, C> m_storage;      // must stay the first and only data member (matrix.h memcpy hack)

public:
    typedef T Scalar;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C };

    Matrix() {}
    Matrix(const Matrix& o) : MatrixBase<Matrix>(), m_storage(o.m_storage) {}
    explicit Matrix(Index n) { if (R == Dynamic && C != Dynamic) m_storage.resize(n, C == Dynamic ? 1 : C); else if (C == Dynamic && R != Dynamic) m_storage.resize(R, n); else m_storage.resize(n, 1); }
    Matrix(Index r, Index c) { m_storage.resize(r, c); }
    template <class O> Matrix(const MatrixBase<O>& o) { assign(o.derived()); }
    Matrix& operator=(const Matrix& o) { m_storage = o.m_storage; return *this; }
    template <class O> Matrix& operator=(const MatrixBase<O>& o) { assign(o.derived()); return *this; }

    Index rows() const { return m_storage.rows(); }
    Index cols() const { return m_storage.cols(); }
    Index size() const { return rows() * cols(); }
    T* data() { return m_storage.ptr(); }
    const T* data() const { return m_storage.ptr(); }
    void resize(Index r, Index c) { m_storage.resize(r, c); }
    void conservativeResize(Index r, Index c) { m_storage.conservative(r, c); }

    T& operator[](Index i) { return data()[i]; }
    const T& operator[](Index i) const { return data()[i]; }
    T& operator()(Index i) { return data()[i]; }
    const T& operator()(Index i) const { return data()[i]; }
    T& operator()(Index i, Index j) { return data()[i + j * rows()]; }
    const T& operator()(Index i, Index j) const { return data()[i + j * rows()]; }

    void setZero() { for (Index i = 0; i < size(); ++i) data()[i] = T(0); }
    Matrix& setLinSpaced(Index n, T lo, T hi)
    {
        m_storage.resize(R == Dynamic ? n : R, C == Dynamic ? (R == Dynamic ? 1 : n) : C);
        // Eigen 3.3 linspaced_op for floating point with |hi| >= |lo|: low + i * step, last element = high
        const T stepv = (n == 1) ? T(1) : (hi - lo) / T(n - 1);
        for (Index i = 0; i < n; ++i) data()[i] = (i == n - 1) ? hi : lo + T(i) * stepv;
        return *this;
    }
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(Index r, Index c) { Matrix m(r, c); m.setZero(); return m; }

    CommaInit<Matrix> operator<<(T v) { data()[0] = v; return CommaInit<Matrix>{*this, 1}; }

    ArrView<T> array() { assert(rows() == 1 || cols() == 1); return ArrView<T>(data(), size(), 1); }
    ArrView<const T> array() const { assert(rows() == 1 || cols() == 1); return ArrView<const T>(data(), size(), 1); }
    Matrix& matrix() { return *this; }
    VecView<T> col(Index j) { return VecView<T>{data() + j * rows(), rows()}; }
    VecView<const T> col(Index j) const { return VecView<const T>{data() + j * rows(), rows()}; }
    BlockView<T> topRows(Index n) { assert(n <= rows()); return BlockView<T>{data(), n, cols(), rows()}; }
    template <int N> BlockView<T> leftCols() { return BlockView<T>{data(), rows(), N, rows()}; }
    template <int BR, int BC> FixedBlock<T, BR, BC> topLeftCorner() { return FixedBlock<T, BR, BC>{data(), rows()}; }
    template <int BR, int BC> FixedBlock<T, BR, BC> bottomLeftCorner() { return FixedBlock<T, BR, BC>{data() + (rows() - BR), rows()}; }
    PrintView<T> transpose() const { return PrintView<T>{data(), rows(), cols(), true}; }

private:
    template <class O> void assign(const O& o)
    {
        m_storage.resize(R == Dynamic ? o.rows() : R, C == Dynamic ? o.cols() : C);
        assert(o.rows() * o.cols() == size());
        for (Index i = 0; i < size(); ++i) data()[i] = o.data()[i];
    }
};
template <class T, int R, int C, int O, int MR, int MC>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C, O, MR, MC>& m)
{ return os << PrintView<T>{m.data(), m.rows(), m.cols(), false}; }

template <class T, int R, int C, int Opt = ColMajor> class Array : public Matrix<T, R, C, Opt> {
public:
    Array() {}
};

// Ref<VectorJ>: a non-owning view of a column vector
template <class M> class Ref {
    typedef typename M::Scalar T;
    T* m_p;
    Index m_n;

public:
    typedef T Scalar;
    Ref(M& m) : m_p(m.data()), m_n(m.size()) {}
    template <class D> Ref(MatrixBase<D>& m) : m_p(m.derived().data()), m_n(m.derived().size()) {}
    T* data() const { return m_p; }
    Index size() const { return m_n; }
    Index rows() const { return m_n; }
    Index cols() const { return 1; }
    T& operator[](Index i) const { return m_p[i]; }
    T& operator()(Index i) const { return m_p[i]; }
};
}  // namespace Eigen
#endif
