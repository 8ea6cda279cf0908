// TEST INFRASTRUCTURE — CPU oracle for the social-MPC solve path. Not shipped, not on the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use oracle/.
//
// jet.hpp — forward-mode dual numbers with the arithmetic of ceres::Jet<double, N>.
// Ceres (libceres-dev, version unpinned by package.xml:27; 2.0.0 on Ubuntu 22.04, 2.2.0 on 24.04) is
// NOT present in /root/reference nor in this image, so its published Jet algebra is restated here:
// every residual functor of the reference is evaluated by DynamicAutoDiffCostFunction with
// Jet<double, 4> (e.g. include/nav2_social_mpc_controller/critics/social_work_cost_function.hpp:59).
// Rules restated (ceres/jet.h): (a,u)*(b,v) = (ab, a v + u b); (a,u)/(b,v) = (a*(1/b), (u - (a/b) v)*(1/b));
// sqrt: (s, u/(2s)); exp: (e, e u); sin: (sin a, cos a u); cos: (cos a, -sin a u);
// atan2((b,v),(a,u)) = (atan2(b,a), (a v - b u)/(a^2+b^2)); comparisons act on the scalar part.
#pragma once
#include <cmath>

namespace smpc_oracle {

// Optional operation counting (make -C oracle liboracle_count.so, -DSMPC_ORACLE_COUNT_OPS): every Jet operator adds
// its FLOPs — SURVEY §8d convention: add / sub / mul / div / sqrt = 1, each sin / cos / exp / atan2 = 1, negation and
// comparisons free — to a thread-local counter. It measures what the REFERENCE-SHAPED algorithm executes (per-residual
// Jet<4> passes, every functor re-rolling-out steps 0..i) next to the minimal single-rollout model of bench.py.
#ifdef SMPC_ORACLE_COUNT_OPS
extern thread_local unsigned long long g_jet_flops;
#define SMPC_OPS(n) (g_jet_flops += static_cast<unsigned long long>(n))
#else
#define SMPC_OPS(n) ((void)0)
#endif

template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  Jet(double s) : a(s) {  // NOLINT: implicit like ceres::Jet(const T&)
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  Jet(double s, int k) : a(s) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
    if (k >= 0 && k < N) v[k] = 1.0;
  }
};

#define SMPC_JET_LOOP for (int i = 0; i < N; ++i)

template <int N>
inline Jet<N> operator+(const Jet<N>& f) {
  return f;
}
template <int N>
inline Jet<N> operator-(const Jet<N>& f) {
  Jet<N> r;
  r.a = -f.a;
  SMPC_JET_LOOP r.v[i] = -f.v[i];
  return r;
}
template <int N>
inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  SMPC_OPS(1 + N);
  Jet<N> r;
  r.a = f.a + g.a;
  SMPC_JET_LOOP r.v[i] = f.v[i] + g.v[i];
  return r;
}
template <int N>
inline Jet<N> operator+(const Jet<N>& f, double s) {
  SMPC_OPS(1);
  Jet<N> r = f;
  r.a = f.a + s;
  return r;
}
template <int N>
inline Jet<N> operator+(double s, const Jet<N>& f) {
  SMPC_OPS(1);
  Jet<N> r = f;
  r.a = f.a + s;
  return r;
}
template <int N>
inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  SMPC_OPS(1 + N);
  Jet<N> r;
  r.a = f.a - g.a;
  SMPC_JET_LOOP r.v[i] = f.v[i] - g.v[i];
  return r;
}
template <int N>
inline Jet<N> operator-(const Jet<N>& f, double s) {
  SMPC_OPS(1);
  Jet<N> r = f;
  r.a = f.a - s;
  return r;
}
template <int N>
inline Jet<N> operator-(double s, const Jet<N>& f) {
  SMPC_OPS(1);
  Jet<N> r;
  r.a = s - f.a;
  SMPC_JET_LOOP r.v[i] = -f.v[i];
  return r;
}
template <int N>
inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  SMPC_OPS(1 + 3 * N);
  Jet<N> r;
  r.a = f.a * g.a;
  SMPC_JET_LOOP r.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return r;
}
template <int N>
inline Jet<N> operator*(const Jet<N>& f, double s) {
  SMPC_OPS(1 + N);
  Jet<N> r;
  r.a = f.a * s;
  SMPC_JET_LOOP r.v[i] = f.v[i] * s;
  return r;
}
template <int N>
inline Jet<N> operator*(double s, const Jet<N>& f) {
  SMPC_OPS(1 + N);
  Jet<N> r;
  r.a = f.a * s;
  SMPC_JET_LOOP r.v[i] = f.v[i] * s;
  return r;
}
template <int N>
inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  SMPC_OPS(2 + 3 * N);
  const double g_inv = 1.0 / g.a;
  const double q = f.a * g_inv;
  Jet<N> r;
  r.a = q;
  SMPC_JET_LOOP r.v[i] = (f.v[i] - q * g.v[i]) * g_inv;
  return r;
}
template <int N>
inline Jet<N> operator/(double s, const Jet<N>& g) {
  SMPC_OPS(3 + N);
  const double m = -s / (g.a * g.a);
  Jet<N> r;
  r.a = s / g.a;
  SMPC_JET_LOOP r.v[i] = g.v[i] * m;
  return r;
}
template <int N>
inline Jet<N> operator/(const Jet<N>& f, double s) {
  SMPC_OPS(2 + N);
  const double s_inv = 1.0 / s;
  Jet<N> r;
  r.a = f.a * s_inv;
  SMPC_JET_LOOP r.v[i] = f.v[i] * s_inv;
  return r;
}
template <int N>
inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) {
  f = f + g;
  return f;
}
template <int N>
inline Jet<N>& operator-=(Jet<N>& f, const Jet<N>& g) {
  f = f - g;
  return f;
}
template <int N>
inline Jet<N>& operator+=(Jet<N>& f, double s) {
  SMPC_OPS(1);
  f.a += s;
  return f;
}
template <int N>
inline Jet<N>& operator-=(Jet<N>& f, double s) {
  SMPC_OPS(1);
  f.a -= s;
  return f;
}

#define SMPC_JET_CMP(op)                                   \
  template <int N>                                         \
  inline bool operator op(const Jet<N>& f, const Jet<N>& g) { \
    return f.a op g.a;                                     \
  }                                                        \
  template <int N>                                         \
  inline bool operator op(const Jet<N>& f, double s) {     \
    return f.a op s;                                       \
  }                                                        \
  template <int N>                                         \
  inline bool operator op(double s, const Jet<N>& f) {     \
    return s op f.a;                                       \
  }
SMPC_JET_CMP(<)
SMPC_JET_CMP(<=)
SMPC_JET_CMP(>)
SMPC_JET_CMP(>=)
SMPC_JET_CMP(==)
SMPC_JET_CMP(!=)
#undef SMPC_JET_CMP

template <int N>
inline Jet<N> sqrt(const Jet<N>& f) {
  SMPC_OPS(3 + N);
  const double t = std::sqrt(f.a);
  const double k = 1.0 / (2.0 * t);
  Jet<N> r;
  r.a = t;
  SMPC_JET_LOOP r.v[i] = f.v[i] * k;
  return r;
}
template <int N>
inline Jet<N> exp(const Jet<N>& f) {
  SMPC_OPS(1 + N);
  const double t = std::exp(f.a);
  Jet<N> r;
  r.a = t;
  SMPC_JET_LOOP r.v[i] = t * f.v[i];
  return r;
}
template <int N>
inline Jet<N> sin(const Jet<N>& f) {
  SMPC_OPS(2 + N);
  const double c = std::cos(f.a);
  Jet<N> r;
  r.a = std::sin(f.a);
  SMPC_JET_LOOP r.v[i] = c * f.v[i];
  return r;
}
template <int N>
inline Jet<N> cos(const Jet<N>& f) {
  SMPC_OPS(2 + N);
  const double s = -std::sin(f.a);
  Jet<N> r;
  r.a = std::cos(f.a);
  SMPC_JET_LOOP r.v[i] = s * f.v[i];
  return r;
}
template <int N>
inline Jet<N> atan2(const Jet<N>& g, const Jet<N>& f) {
  SMPC_OPS(5 + 3 * N);
  const double t = 1.0 / (f.a * f.a + g.a * g.a);
  Jet<N> r;
  r.a = std::atan2(g.a, f.a);
  SMPC_JET_LOOP r.v[i] = t * (-g.a * f.v[i] + f.a * g.v[i]);
  return r;
}
#undef SMPC_JET_LOOP

// double overloads so templated residual code can call the same names (ADL finds the Jet ones).
inline double sqrt(double x) { return std::sqrt(x); }
inline double exp(double x) { return std::exp(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double atan2(double y, double x) { return std::atan2(y, x); }

inline double scalar_part(double x) { return x; }
template <int N>
inline double scalar_part(const Jet<N>& x) {
  return x.a;
}

}  // namespace smpc_oracle
