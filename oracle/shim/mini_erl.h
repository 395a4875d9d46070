// mini_erl.h — TEST INFRASTRUCTURE (oracle/_ref build only; never part of libctrfk.so).
//
// The reference includes <Erl/...> ("Erl", https://gitlab.com/ggras/Erl.git, unpinned HEAD,
// install.sh:41), which is not in this image and cannot be fetched.  This header is a stand-in
// written from the reference's CALL SITES; nothing of Erl's source has been seen.  Two kinds of
// content:
//   (1) plumbing whose meaning the call sites fix completely: static_for / static_if /
//       static_if_else (compile-time loops over std::integral_constant and tuples), aligned
//       allocation, max_alignment, debug streams, Spinlock, clamp;
//   (2) ASSUMED SEMANTICS (SURVEY §8c i-vi), each marked [ASSUMED] below: Rotmat's 9-scalar
//       constructor takes row-major arguments and stores column-major (forced by
//       robot_i.h:1252-1259 being Rodrigues' formula and the exact arc translation only under that
//       reading); Transform's 12-scalar constructor is row-major r00 r01 r02 tx ...; '*' is SE(3)
//       composition with the usual row-times-column summation order; equal(a,b) is
//       |a-b| < Constants::Zero_Tolerance (value 1e-12 assumed for double, 1e-6 for float);
//       nearestOrthogonal is the polar factor.
// What building the reference against this header pins: every line of the reference's own
// robot_i.h / section_i.h / ctr_static.h / ctr_robot.cpp (sample counts, the tip->base sweep, the
// lagged torsion slot, radius fill, Jacobian and stability drivers, XML readers).  What it does
// not pin: the five [ASSUMED] items.
#ifndef CTRB200_MINI_ERL_H
#define CTRB200_MINI_ERL_H
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <new>
#include <sstream>
#include <tuple>
#include <type_traits>
#include <utility>

#ifndef ERL_PI
#define ERL_PI 3.14159265358979323846
#endif
#define ERL_ASSERT(x) assert(x)
#define ERL_RMV_UNINIT_SINGLE
#define CT_DEBUG_OFF 0
#define CT_DEBUG_NORMAL 1
#define CT_DEBUG_NOISY 2
#ifndef CT_DEBUG_LEVEL
#define CT_DEBUG_LEVEL CT_DEBUG_NOISY
#endif

namespace Erl {
// ------------------------------------------------------------------ compile-time control flow
template <long S, long E, long Inc = 1> struct static_for {
    static_assert(Inc != 0, "static_for: zero increment");
    template <class F> void operator()(F&& f) const { step_i<S>(f, keep<S>()); }
    template <class F, class T0> void operator()(F&& f, T0&& t0) const { step_1<S>(f, t0, keep<S>()); }
    template <class F, class T0, class T1> void operator()(F&& f, T0&& t0, T1&& t1) const { step_2<S>(f, t0, t1, keep<S>()); }

private:
    template <long I> using keep = std::integral_constant<bool, (Inc > 0 ? (I < E) : (I > E))>;
    template <long I, class F> static void step_i(F&, std::false_type) {}
    template <long I, class F> static void step_i(F& f, std::true_type)
    { f(std::integral_constant<long, I>()); step_i<I + Inc>(f, keep<I + Inc>()); }
    template <long I, class F, class T0> static void step_1(F&, T0&, std::false_type) {}
    template <long I, class F, class T0> static void step_1(F& f, T0& t0, std::true_type)
    { f(std::get<I>(t0), std::integral_constant<long, I>()); step_1<I + Inc>(f, t0, keep<I + Inc>()); }
    template <long I, class F, class T0, class T1> static void step_2(F&, T0&, T1&, std::false_type) {}
    template <long I, class F, class T0, class T1> static void step_2(F& f, T0& t0, T1& t1, std::true_type)
    { f(std::get<I>(t0), std::get<I>(t1), std::integral_constant<long, I>()); step_2<I + Inc>(f, t0, t1, keep<I + Inc>()); }
};
template <bool B> struct static_if {
    template <class F> void operator()(F&& f) const { run(f, std::integral_constant<bool, B>()); }
private:
    template <class F> static void run(F& f, std::true_type) { f(); }
    template <class F> static void run(F&, std::false_type) {}
};
template <bool B> struct static_if_else {
    template <class F, class G> auto operator()(F&& f, G&& g) const { return run(f, g, std::integral_constant<bool, B>()); }
private:
    template <class F, class G> static auto run(F& f, G&, std::true_type) { return f(); }
    template <class F, class G> static auto run(F&, G& g, std::false_type) { return g(); }
};

// ------------------------------------------------------------------ memory
inline size_t align_padded_size(size_t n, size_t a) { return (n + a - 1) / a * a; }
inline void* aligned_malloc(size_t n, size_t a)
{
    void* p = nullptr;
    if (a < sizeof(void*)) a = sizeof(void*);
    if (posix_memalign(&p, a, n ? n : a)) return nullptr;
    return p;
}
inline void aligned_free(void* p) { std::free(p); }
template <class T, class... A> T* alloc_aligned_init_array(size_t n, size_t a, A&&... args)
{
    T* p = static_cast<T*>(aligned_malloc(sizeof(T) * n, a));
    if (p) for (size_t i = 0; i < n; ++i) new (p + i) T(args...);
    return p;
}
template <class T> T* alloc_aligned_init_array_noparameter(size_t n, size_t a)
{
    T* p = static_cast<T*>(aligned_malloc(sizeof(T) * n, a));
    if (p) for (size_t i = 0; i < n; ++i) new (p + i) T();
    return p;
}
template <class T> void free_aligned_init_array(T* p, size_t n)
{
    if (!p) return;
    for (size_t i = 0; i < n; ++i) p[i].~T();
    aligned_free(p);
}
namespace details {
template <class T, size_t A> struct max_alignment { static constexpr size_t value = (alignof(T) > A ? alignof(T) : A); };
}
template <class E> constexpr int as_int(E e) { return static_cast<int>(e); }

// ------------------------------------------------------------------ numerics helpers
template <class T> struct Constants { static constexpr T Zero_Tolerance = T(1e-12); };       // [ASSUMED] value
template <> struct Constants<float> { static constexpr float Zero_Tolerance = 1e-6f; };      // [ASSUMED] value
template <class T> constexpr T Constants<T>::Zero_Tolerance;
template <class T> inline bool equal(const T& a, const T& b) { return std::abs(a - b) < Constants<T>::Zero_Tolerance; }   // [ASSUMED]
template <class T> inline T clamp(const T& v, const T& lo, const T& hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct ColMajorTag {};
struct RowMajorTag {};
static const ColMajorTag ColMajor = ColMajorTag();
static const RowMajorTag RowMajor = RowMajorTag();

// ------------------------------------------------------------------ Vector3 / Vector6 / Rotmat / Transform
template <class T> class Vector3 {
    T m[3];
public:
    Vector3() : m{T(0), T(0), T(0)} {}
    Vector3(T x, T y, T z) : m{x, y, z} {}
    explicit Vector3(const T* p) : m{p[0], p[1], p[2]} {}
    static Vector3 Zero() { return Vector3(); }
    T* data() { return m; }
    const T* data() const { return m; }
    T x() const { return m[0]; }
    T y() const { return m[1]; }
    T z() const { return m[2]; }
    T& operator[](int i) { return m[i]; }
    const T& operator[](int i) const { return m[i]; }
    T norm() const { return std::sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]); }       // Eigen-style: sqrt of the squared sum
    Vector3 operator/(T s) const { return Vector3(m[0] / s, m[1] / s, m[2] / s); }
    Vector3 operator*(T s) const { return Vector3(m[0] * s, m[1] * s, m[2] * s); }
    Vector3 operator-(const Vector3& o) const { return Vector3(m[0] - o.m[0], m[1] - o.m[1], m[2] - o.m[2]); }
    Vector3 operator+(const Vector3& o) const { return Vector3(m[0] + o.m[0], m[1] + o.m[1], m[2] + o.m[2]); }
    Vector3 cross(const Vector3& o) const
    { return Vector3(m[1] * o.m[2] - m[2] * o.m[1], m[2] * o.m[0] - m[0] * o.m[2], m[0] * o.m[1] - m[1] * o.m[0]); }
    Vector3 transpose() const { return *this; }
};
template <class T> std::ostream& operator<<(std::ostream& os, const Vector3<T>& v) { return os << v.x() << " " << v.y() << " " << v.z(); }
template <class T> class Vector6 {
    T m[6];
public:
    Vector6() : m{} {}
    T* data() { return m; }
    const T* data() const { return m; }
    T& operator[](int i) { return m[i]; }
    const T& operator[](int i) const { return m[i]; }
};
template <class T> class Quaternion;

template <class T> class Rotmat {
    T m[9];      // column-major: m[r + 3c]
public:
    Rotmat() : m{T(1), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(1)} {}
    // [ASSUMED] nine scalars in ROW-major order, stored column-major
    Rotmat(T r00, T r01, T r02, T r10, T r11, T r12, T r20, T r21, T r22) : m{r00, r10, r20, r01, r11, r21, r02, r12, r22} {}
    static Rotmat Identity() { return Rotmat(); }
    T* data() { return m; }
    const T* data() const { return m; }
    T& operator()(int r, int c) { return m[r + 3 * c]; }
    const T& operator()(int r, int c) const { return m[r + 3 * c]; }
    Rotmat transpose() const { return Rotmat(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]); }
    Rotmat inv() const { return transpose(); }
    Rotmat operator*(const Rotmat& o) const        // [ASSUMED] summation order a(r,0)b(0,c) + a(r,1)b(1,c) + a(r,2)b(2,c)
    {
        Rotmat q;
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) q.m[r + 3 * c] = m[r] * o.m[3 * c] + m[r + 3] * o.m[3 * c + 1] + m[r + 6] * o.m[3 * c + 2];
        return q;
    }
    Vector3<T> operator*(const Vector3<T>& v) const
    { return Vector3<T>(m[0] * v[0] + m[3] * v[1] + m[6] * v[2], m[1] * v[0] + m[4] * v[1] + m[7] * v[2], m[2] * v[0] + m[5] * v[1] + m[8] * v[2]); }
    Rotmat operator-(const Rotmat& o) const { Rotmat q; for (int i = 0; i < 9; ++i) q.m[i] = m[i] - o.m[i]; return q; }
};

template <class T> class Transform {
    Rotmat<T> R;
    Vector3<T> t;
public:
    Transform() {}
    // [ASSUMED] twelve scalars in ROW-major order: r00 r01 r02 tx  r10 r11 r12 ty  r20 r21 r22 tz
    Transform(T r00, T r01, T r02, T tx, T r10, T r11, T r12, T ty, T r20, T r21, T r22, T tz)
        : R(r00, r01, r02, r10, r11, r12, r20, r21, r22), t(tx, ty, tz) {}
    Transform(const Rotmat<T>& r, const Vector3<T>& v) : R(r), t(v) {}
    static Transform Identity() { return Transform(); }
    static Transform Zero() { return Transform(T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)); }
    const Rotmat<T>& getRotationReference() const { return R; }
    const Vector3<T>& getTranslationReference() const { return t; }
    Rotmat<T> getRotation() const { return R; }
    Vector3<T> getTranslation() const { return t; }
    void setRotation(const Rotmat<T>& r) { R = r; }
    void setTranslation(const Vector3<T>& v) { t = v; }
    Vector3<T> getColumn0() const { return Vector3<T>(R.data()); }
    Vector3<T> getColumn1() const { return Vector3<T>(R.data() + 3); }
    Vector3<T> getColumn2() const { return Vector3<T>(R.data() + 6); }
    T operator()(int r, int c) const { return c < 3 ? R(r, c) : t[r]; }
    Transform operator*(const Transform& o) const           // [ASSUMED] SE(3) composition: (R R', R t' + t)
    {
        const Vector3<T> rt = R * o.t;
        return Transform(R * o.R, Vector3<T>(rt[0] + t[0], rt[1] + t[1], rt[2] + t[2]));
    }
    Transform& operator*=(const Transform& o) { *this = *this * o; return *this; }
    Transform operator-(const Transform& o) const { return Transform(R - o.R, t - o.t); }      // coefficient-wise (robot_i.h:48)
    void getData(T* out, ColMajorTag) const { std::memcpy(out, R.data(), 9 * sizeof(T)); std::memcpy(out + 9, t.data(), 3 * sizeof(T)); }
    void getData(T* out, RowMajorTag) const
    { for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) out[4 * r + c] = R(r, c); out[4 * r + 3] = t[r]; } }
    static bool verifyOrthogonal(const Transform& x, T tol)
    {
        const Rotmat<T> g = x.R.transpose() * x.R;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) if (std::abs(g(r, c) - (r == c ? T(1) : T(0))) > tol) return false;
        return true;
    }
    // [ASSUMED] polar factor of the rotation block (Newton iteration X <- (X + X^-T)/2), translation kept
    static Transform nearestOrthogonal(const Transform& x)
    {
        long double a[9];
        for (int i = 0; i < 9; ++i) a[i] = x.R.data()[i];
        for (int it = 0; it < 60; ++it) {
            long double cof[9], det = 0;     // cof(r,c): signed cofactor; A^-T = cof / det
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    cof[r + 3 * c] = a[(r + 1) % 3 + 3 * ((c + 1) % 3)] * a[(r + 2) % 3 + 3 * ((c + 2) % 3)]
                                   - a[(r + 1) % 3 + 3 * ((c + 2) % 3)] * a[(r + 2) % 3 + 3 * ((c + 1) % 3)];
            for (int c = 0; c < 3; ++c) det += a[3 * c] * cof[3 * c];
            long double d = 0;
            for (int i = 0; i < 9; ++i) {
                const long double n = 0.5L * (a[i] + cof[i] / det);
                d = std::max(d, std::abs(n - a[i]));
                a[i] = n;
            }
            if (d < 1e-19L) break;
        }
        Transform y = x;
        for (int i = 0; i < 9; ++i) y.R.data()[i] = T(a[i]);
        return y;
    }
};
template <class T> std::ostream& operator<<(std::ostream& os, const Transform<T>& x)
{
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 4; ++c) os << (c ? " " : "") << x(r, c); os << "\n"; }
    return os;
}

// ------------------------------------------------------------------ debug streams, spinlock
inline std::ostream& debug_cout() { return std::cout; }
inline std::ostream& debug_cout_e() { return std::cerr; }
inline std::ostream& debug_cout_g() { return std::cout; }
template <class S> inline void debug_set_io_format(S& s, int precision, int width, std::ios_base::fmtflags f, bool)
{ s.precision(precision); s.width(width); s.setf(f, std::ios_base::floatfield); }
namespace RT {
class Spinlock {
public:
    void lock() {}
    void unlock() {}
    bool try_lock() { return true; }
};
}  // namespace RT
}  // namespace Erl
#endif
