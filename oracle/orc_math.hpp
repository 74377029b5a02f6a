// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load anything under oracle/.
// PARITY PINNED against the reference's own code: oracle/_ref (the reference's sources compiled against
// oracle/mockfoam) agrees with this restatement bit for bit (tests/test_oracle_vs_reference.py), and the vectors that
// build produced are committed (tests/golden/, tests/test_golden_oracle.py).  OpenFOAM's own arithmetic (mesh geometry,
// findCell, linear face interpolation) is restated, not pinned: see oracle/README.md.
//
// Minimal restatement of the OpenFOAM primitive arithmetic the reference's hot path uses
// (Vector / Tensor / SymmTensor operators).  Evaluation order follows OpenFOAM's
// VectorI.H / TensorI.H / SymmTensorI.H: sums are left-associated, no FMA contraction
// (build with -ffp-contract=off, plain x86-64 SSE2 like a default OpenFOAM build).
#pragma once
#include <cmath>
#include <array>

namespace orc {

constexpr double SMALL = 1.0e-15;              // OpenFOAM doubleScalar SMALL
constexpr double PI    = 3.14159265358979323846;  // constant::mathematical::pi (M_PI)
constexpr double TWOPI = 2.0 * PI;             // constant::mathematical::twoPi

struct vec3 {
  double x, y, z;
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct ivec3 { int x, y, z; };

inline vec3 operator+(const vec3& a, const vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(const vec3& a, const vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(const vec3& a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(double s, const vec3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 operator*(const vec3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator/(const vec3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3& operator+=(vec3& a, const vec3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline vec3& operator-=(vec3& a, const vec3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
inline vec3& operator*=(vec3& a, double s) { a.x *= s; a.y *= s; a.z *= s; return a; }
// OpenFOAM operator& (inner product) and operator^ (cross product)
inline double dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double magSqr(const vec3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double mag(const vec3& a) { return std::sqrt(magSqr(a)); }
inline double cmptMax(const vec3& a) { return std::fmax(a.x, std::fmax(a.y, a.z)); }
inline double sqr(double a) { return a * a; }
inline double pow3(double a) { return a * a * a; }
constexpr vec3 Zero3{0.0, 0.0, 0.0};

struct tensor {
  double xx, xy, xz, yx, yy, yz, zx, zy, zz;
};
constexpr tensor TensorI{1, 0, 0, 0, 1, 0, 0, 0, 1};
// Tensor & Tensor (TensorI.H inner product)
inline tensor dot(const tensor& a, const tensor& b) {
  return {a.xx * b.xx + a.xy * b.yx + a.xz * b.zx, a.xx * b.xy + a.xy * b.yy + a.xz * b.zy,
          a.xx * b.xz + a.xy * b.yz + a.xz * b.zz,
          a.yx * b.xx + a.yy * b.yx + a.yz * b.zx, a.yx * b.xy + a.yy * b.yy + a.yz * b.zy,
          a.yx * b.xz + a.yy * b.yz + a.yz * b.zz,
          a.zx * b.xx + a.zy * b.yx + a.zz * b.zx, a.zx * b.xy + a.zy * b.yy + a.zz * b.zy,
          a.zx * b.xz + a.zy * b.yz + a.zz * b.zz};
}
// Tensor & Vector
inline vec3 dot(const tensor& t, const vec3& v) {
  return {t.xx * v.x + t.xy * v.y + t.xz * v.z, t.yx * v.x + t.yy * v.y + t.yz * v.z,
          t.zx * v.x + t.zy * v.y + t.zz * v.z};
}

// SymmTensor, OpenFOAM component order xx xy xz yy yz zz
struct symm6 {
  double xx, xy, xz, yy, yz, zz;
};
constexpr symm6 ZeroS{0, 0, 0, 0, 0, 0};
inline symm6 operator+(const symm6& a, const symm6& b) {
  return {a.xx + b.xx, a.xy + b.xy, a.xz + b.xz, a.yy + b.yy, a.yz + b.yz, a.zz + b.zz};
}
inline symm6 operator-(const symm6& a, const symm6& b) {
  return {a.xx - b.xx, a.xy - b.xy, a.xz - b.xz, a.yy - b.yy, a.yz - b.yz, a.zz - b.zz};
}
inline symm6 operator*(double s, const symm6& a) {
  return {s * a.xx, s * a.xy, s * a.xz, s * a.yy, s * a.yz, s * a.zz};
}
inline symm6 operator*(const symm6& a, double s) {
  return {a.xx * s, a.xy * s, a.xz * s, a.yy * s, a.yz * s, a.zz * s};
}
inline symm6& operator+=(symm6& a, const symm6& b) {
  a.xx += b.xx; a.xy += b.xy; a.xz += b.xz; a.yy += b.yy; a.yz += b.yz; a.zz += b.zz;
  return a;
}
// SymmTensor & Vector
inline vec3 dot(const symm6& s, const vec3& v) {
  return {s.xx * v.x + s.xy * v.y + s.xz * v.z, s.xy * v.x + s.yy * v.y + s.yz * v.z,
          s.xz * v.x + s.yz * v.y + s.zz * v.z};
}

}  // namespace orc
