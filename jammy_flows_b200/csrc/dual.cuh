// Forward-mode dual numbers for the manifold layers' device code (csrc/s2.cuh, chain1.cuh, spline.cuh are templates over
// the scalar type): instantiating them with T = Dual gives d(output)/d(one seeded input) alongside the value, with the
// same branches, clamps and operation order as the value path.  csrc/jac_sweep.cuh sweeps the seed over the parameters
// and coordinates of a row to build the per-row Jacobian of log_pdf that the reference obtains from autograd.
// Values and derivatives are carried in double whatever the storage type of the tensors is.
#pragma once
#include "common.cuh"

#define JF_HD __host__ __device__ __forceinline__

namespace jf {

struct Dual {
    double v, d;
    Dual() = default;
    JF_HD Dual(double v_) : v(v_), d(0.0) {}
    JF_HD Dual(float v_) : v((double)v_), d(0.0) {}
    JF_HD Dual(int v_) : v((double)v_), d(0.0) {}
    JF_HD Dual(double v_, double d_) : v(v_), d(d_) {}
    JF_HD Dual& operator+=(const Dual& o) { v += o.v; d += o.d; return *this; }
    JF_HD Dual& operator-=(const Dual& o) { v -= o.v; d -= o.d; return *this; }
    JF_HD Dual& operator*=(const Dual& o) { d = d * o.v + v * o.d; v *= o.v; return *this; }
    JF_HD Dual& operator/=(const Dual& o) { const double q = v / o.v; d = (d - q * o.d) / o.v; v = q; return *this; }
    friend JF_HD Dual operator+(const Dual& a, const Dual& b) { return Dual(a.v + b.v, a.d + b.d); }
    friend JF_HD Dual operator-(const Dual& a, const Dual& b) { return Dual(a.v - b.v, a.d - b.d); }
    friend JF_HD Dual operator*(const Dual& a, const Dual& b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
    friend JF_HD Dual operator/(const Dual& a, const Dual& b) { const double q = a.v / b.v; return Dual(q, (a.d - q * b.d) / b.v); }
    friend JF_HD Dual operator-(const Dual& a) { return Dual(-a.v, -a.d); }
    friend JF_HD bool operator<(const Dual& a, const Dual& b) { return a.v < b.v; }
    friend JF_HD bool operator>(const Dual& a, const Dual& b) { return a.v > b.v; }
    friend JF_HD bool operator<=(const Dual& a, const Dual& b) { return a.v <= b.v; }
    friend JF_HD bool operator>=(const Dual& a, const Dual& b) { return a.v >= b.v; }
    friend JF_HD bool operator==(const Dual& a, const Dual& b) { return a.v == b.v; }
    friend JF_HD bool operator!=(const Dual& a, const Dual& b) { return a.v != b.v; }
};

template <> struct Num<Dual> {
    static constexpr double eps = Num<double>::eps;
    static constexpr double newton_abs_tol = Num<double>::newton_abs_tol;
    static constexpr double target_prec = Num<double>::target_prec;
    static constexpr double safe_costheta = Num<double>::safe_costheta;
    static constexpr double kappa_identity = Num<double>::kappa_identity;
    static constexpr double big = Num<double>::big;
};

// the overloads below live in namespace jf, where an unqualified call from other jf code would otherwise no longer see
// the global math functions (name hiding): make both part of one overload set
using ::exp; using ::log; using ::log1p; using ::sqrt; using ::sin; using ::cos; using ::sincos; using ::tanh; using ::acos;
using ::atan2; using ::fabs; using ::erf; using ::erfc; using ::erfinv; using ::erfcinv; using ::fma; using ::isfinite;

// a zero seed times an infinite local derivative (sqrt at 0, acos at +-1) is 0, not NaN: the value does not depend on it
JF_DEVINL double dmul0(double seed, double slope) { return seed == 0.0 ? 0.0 : seed * slope; }

JF_DEVINL Dual exp(const Dual& a) { const double e = ::exp(a.v); return Dual(e, e * a.d); }
JF_DEVINL Dual log(const Dual& a) { return Dual(::log(a.v), dmul0(a.d, 1.0 / a.v)); }
JF_DEVINL Dual log1p(const Dual& a) { return Dual(::log1p(a.v), dmul0(a.d, 1.0 / (1.0 + a.v))); }
JF_DEVINL Dual sqrt(const Dual& a) { const double s = ::sqrt(a.v); return Dual(s, dmul0(a.d, 0.5 / s)); }
JF_DEVINL Dual sin(const Dual& a) { double s, c; ::sincos(a.v, &s, &c); return Dual(s, c * a.d); }
JF_DEVINL Dual cos(const Dual& a) { double s, c; ::sincos(a.v, &s, &c); return Dual(c, -s * a.d); }
JF_DEVINL void sincos(const Dual& a, Dual* sp, Dual* cp) {
    double s, c;
    ::sincos(a.v, &s, &c);
    *sp = Dual(s, c * a.d);
    *cp = Dual(c, -s * a.d);
}
JF_DEVINL Dual tanh(const Dual& a) { const double t = ::tanh(a.v); return Dual(t, (1.0 - t * t) * a.d); }
JF_DEVINL Dual acos(const Dual& a) { return Dual(::acos(a.v), dmul0(a.d, -1.0 / ::sqrt(1.0 - a.v * a.v))); }
JF_DEVINL Dual atan2(const Dual& y, const Dual& x) {
    const double r2 = x.v * x.v + y.v * y.v;
    return Dual(::atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / r2);
}
JF_DEVINL Dual fabs(const Dual& a) { return a.v < 0.0 ? Dual(-a.v, -a.d) : a; }
JF_DEVINL Dual erf(const Dual& a) { return Dual(::erf(a.v), 1.1283791670955126 * ::exp(-a.v * a.v) * a.d); }
JF_DEVINL Dual erfc(const Dual& a) { return Dual(::erfc(a.v), -1.1283791670955126 * ::exp(-a.v * a.v) * a.d); }
JF_DEVINL Dual erfinv(const Dual& a) { const double y = ::erfinv(a.v); return Dual(y, dmul0(a.d, 0.88622692545275801 * ::exp(y * y))); }
JF_DEVINL Dual erfcinv(const Dual& a) { const double y = ::erfcinv(a.v); return Dual(y, dmul0(a.d, -0.88622692545275801 * ::exp(y * y))); }
JF_DEVINL Dual fma(const Dual& a, const Dual& b, const Dual& c) { return a * b + c; }
JF_DEVINL bool isfinite(const Dual& a) { return ::isfinite(a.v) && ::isfinite(a.d); }
// the lean primitives of common.cuh (only reachable through shared helpers)
JF_DEVINL Dual exp_neg(const Dual& a) { return exp(a); }
JF_DEVINL Dual exp_clamped(const Dual& a) { return exp(a); }
JF_DEVINL Dual rcp_1to2(const Dual& a) { return Dual(1.0) / a; }

// "is this a double-precision instantiation" for tolerances that the code picks by sizeof(T)
template <> struct Prec<Dual> { static constexpr bool f64 = true; };

}  // namespace jf
