// f2cpp_rt.h -- run-time support of the C++ that oracle/f2cpp.py emits from the reference's Fortran.
// TEST INFRASTRUCTURE ONLY.  Hand-written: Fortran array views, the pool look-ups (mpas_pool_routines.F semantics reduced
// to "name -> array / dimension / namelist value") and the numeric intrinsics with the semantics of the Fortran standard
// as gfortran implements them.  No model arithmetic lives here.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#ifdef REF_SINGLE
typedef float real;
#define RL(x) x##f
#else
typedef double real;
#define RL(x) x
#endif

// column-major view with declared lower bounds (Fortran array element order, F2003 6.2.2.2)
template <class T> struct FArr {
    T* p = nullptr;
    int rank = 0;
    long lo[3] = {1, 1, 1}, n[3] = {1, 1, 1};
    void bind(T* q, long l0, long h0) { p = q; rank = 1; lo[0] = l0; n[0] = std::max(0L, h0 - l0 + 1); }
    void bind(T* q, long l0, long h0, long l1, long h1) { bind(q, l0, h0); rank = 2; lo[1] = l1; n[1] = std::max(0L, h1 - l1 + 1); }
    void bind(T* q, long l0, long h0, long l1, long h1, long l2, long h2) { bind(q, l0, h0, l1, h1); rank = 3; lo[2] = l2; n[2] = std::max(0L, h2 - l2 + 1); }
    static size_t count(long l0, long h0) { return (size_t)std::max(0L, h0 - l0 + 1); }
    static size_t count(long l0, long h0, long l1, long h1) { return count(l0, h0) * count(l1, h1); }
    static size_t count(long l0, long h0, long l1, long h1, long l2, long h2) { return count(l0, h0, l1, h1) * count(l2, h2); }
    FArr rebased() const { FArr r = *this; r.lo[0] = r.lo[1] = r.lo[2] = 1; return r; }        // assumed-shape dummy: lower bounds 1
    inline T& operator()(long i) const { return p[i - lo[0]]; }
    inline T& operator()(long i, long j) const { return p[(i - lo[0]) + n[0] * (j - lo[1])]; }
    inline T& operator()(long i, long j, long k) const { return p[(i - lo[0]) + n[0] * ((j - lo[1]) + n[1] * (k - lo[2]))]; }
    size_t size() const { size_t s = 1; for (int d = 0; d < rank; d++) s *= (size_t)n[d]; return s; }
    template <class V> void fill(V v) { const size_t s = size(); for (size_t t = 0; t < s; t++) p[t] = (T)v; }
};
template <class T> struct FieldT { FArr<T> array; };
template <class T> struct Opt { bool present = false; T v{}; Opt() {} Opt(T x) : present(true), v(x) {} };
struct BlockT { int dummy = 0; };
struct DomainT { int dummy = 0; };
typedef std::function<void(const std::string&)> ExchFn;

struct PoolEntry { void* p[2] = {nullptr, nullptr}; int rank = 0; long n[3] = {1, 1, 1}; bool is_int = false; };
struct CfgVal { int kind = 0; double r = 0; int i = 0; std::string s; };      // 0 real, 1 int, 2 logical, 3 string
struct Pool {
    std::string name;
    std::map<std::string, PoolEntry> arrays;
    std::map<std::string, int>* dims = nullptr;            // shared by every pool of a block (mpas_pool_get_dimension)
    std::map<std::string, CfgVal>* cfgs = nullptr;
    template <class T> FArr<T> arr(const std::string& key, int lev, int rank) const {
        auto it = arrays.find(key);
        if (it == arrays.end()) { fprintf(stderr, "f2cpp_rt: pool %s has no array '%s'\n", name.c_str(), key.c_str()); abort(); }
        const PoolEntry& e = it->second;
        if (e.is_int != std::is_same<T, int>::value || (rank && rank != e.rank)) { fprintf(stderr, "f2cpp_rt: type/rank mismatch for '%s' in pool %s\n", key.c_str(), name.c_str()); abort(); }
        FArr<T> a; a.p = (T*)e.p[lev - 1]; a.rank = e.rank;
        for (int d = 0; d < 3; d++) { a.lo[d] = 1; a.n[d] = e.n[d]; }
        if (!a.p) { fprintf(stderr, "f2cpp_rt: '%s' has no time level %d\n", key.c_str(), lev); abort(); }
        return a;
    }
    template <class T> T scalar(const std::string& key) const { return arr<T>(key, 1, 0)(1); }
    int dim(const std::string& key) const {
        auto it = dims->find(key);
        if (it == dims->end()) { fprintf(stderr, "f2cpp_rt: no dimension '%s'\n", key.c_str()); abort(); }
        return it->second;
    }
    const CfgVal& cv(const std::string& key) const {
        auto it = cfgs->find(key);
        if (it == cfgs->end()) { fprintf(stderr, "f2cpp_rt: no config '%s'\n", key.c_str()); abort(); }
        return it->second;
    }
    void cfg(const std::string& key, real& v) const { v = (real)cv(key).r; }
    void cfg(const std::string& key, int& v) const { v = cv(key).i; }
    void cfg(const std::string& key, bool& v) const { v = cv(key).i != 0; }
    void cfg(const std::string& key, std::string& v) const { v = cv(key).s; }
};

inline void rt_log(const std::string& msg) { fprintf(stderr, "[reference mpas_log_write] %s\n", msg.c_str()); }

// ---- intrinsics
template <class A, class B> inline typename std::common_type<A, B>::type f_max(A a, B b) { typedef typename std::common_type<A, B>::type T; return (T)a > (T)b ? (T)a : (T)b; }
template <class A, class B, class... R> inline auto f_max(A a, B b, R... r) { return f_max(f_max(a, b), r...); }
template <class A, class B> inline typename std::common_type<A, B>::type f_min(A a, B b) { typedef typename std::common_type<A, B>::type T; return (T)a < (T)b ? (T)a : (T)b; }
template <class A, class B, class... R> inline auto f_min(A a, B b, R... r) { return f_min(f_min(a, b), r...); }
inline double f_abs(double x) { return std::fabs(x); }
inline float f_abs(float x) { return std::fabs(x); }
inline int f_abs(int x) { return x < 0 ? -x : x; }
// SIGN(A, B): |A| with the sign of B; gfortran takes the sign BIT of B (so B = -0.0 gives -|A|)
inline double f_sign(double a, double b) { return std::copysign(a, b); }
inline float f_sign(float a, float b) { return std::copysign(a, b); }
inline int f_sign(int a, int b) { return b >= 0 ? (a < 0 ? -a : a) : (a < 0 ? a : -a); }
inline double f_sqrt(double x) { return std::sqrt(x); }
inline float f_sqrt(float x) { return std::sqrt(x); }
inline double f_exp(double x) { return std::exp(x); }
inline float f_exp(float x) { return std::exp(x); }
inline double f_log(double x) { return std::log(x); }
inline float f_log(float x) { return std::log(x); }
inline double f_cos(double x) { return std::cos(x); }
inline float f_cos(float x) { return std::cos(x); }
inline double f_sin(double x) { return std::sin(x); }
inline float f_sin(float x) { return std::sin(x); }
inline double f_acos(double x) { return std::acos(x); }
inline float f_acos(float x) { return std::acos(x); }
inline double f_asin(double x) { return std::asin(x); }
inline double f_atan(double x) { return std::atan(x); }
inline double f_tanh(double x) { return std::tanh(x); }
inline int f_mod(int a, int b) { return a % b; }
inline double f_mod(double a, double b) { return std::fmod(a, b); }
template <class A, class B> inline typename std::common_type<A, B>::type f_merge(A a, B b, bool m) { return m ? a : b; }
inline int f_nint(double x) { return (int)std::lround(x); }
inline int f_floor(double x) { return (int)std::floor(x); }
// x ** n, integer n: libgcc's __powidf2 / __powisf2 (what gfortran emits for a non-constant or constant integer exponent)
template <class T> inline T f_powi(T x, int m) {
    unsigned int n = m < 0 ? -(unsigned int)m : (unsigned int)m;
    T y = (n % 2) ? x : (T)1;
    while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
    return m < 0 ? (T)1 / y : y;
}
inline int f_powi(int x, int m) { int y = 1; for (int t = 0; t < m; t++) y *= x; return y; }
inline double f_pow(double x, double y) { return std::pow(x, y); }
inline float f_pow(float x, float y) { return std::pow(x, y); }
inline double f_pow(double x, float y) { return std::pow(x, (double)y); }
inline double f_pow(float x, double y) { return std::pow((double)x, y); }
