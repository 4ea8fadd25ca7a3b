// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// javax.vecmath 1.3.2 value types restated in plain C++ (jars/vecmath.jar ships its sources):
// Matrix3d.mul :1524-1537, mulTransposeRight :1693-1706, mulTransposeLeft :1737-1750,
// transform :2050-2071, normalizeCP :1865-1881, invertGeneral :1067-1133 (luDecomposition :1135,
// luBacksubstitution :1285), Vector3d.cross :103-112, normalize :134-142, Tuple3d.scaleAdd :277-282,
// interpolate :622-626, Point3d.distance :119-127; and mergingBodies3D/RigidTransform3D.java.
// Operation order is preserved exactly; build with -ffp-contract=off (Java never fuses a*b+c).
// PARITY UNPINNED: the reference holds no golden vectors for the 3D path.
#pragma once
#include <cmath>
#include <cstring>

namespace amo {

struct V3 {
  double x = 0, y = 0, z = 0;
  V3() {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  double get(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  void setc(int i, double v) { if (i == 0) x = v; else if (i == 1) y = v; else z = v; }
};

inline V3 sub(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 add(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 scale(double s, const V3& a) { return V3(s * a.x, s * a.y, s * a.z); }
// Tuple3d.scaleAdd(s,t1,t2) = s*t1 + t2
inline V3 scaleAdd(double s, const V3& t1, const V3& t2) { return V3(s * t1.x + t2.x, s * t1.y + t2.y, s * t1.z + t2.z); }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double lengthSquared(const V3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double length(const V3& a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 cross(const V3& v1, const V3& v2) {
  return V3(v1.y * v2.z - v1.z * v2.y, v2.x * v1.z - v2.z * v1.x, v1.x * v2.y - v1.y * v2.x);
}
inline V3 normalize(const V3& v) {
  double norm = 1.0 / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  return V3(v.x * norm, v.y * norm, v.z * norm);
}
inline double distance(const V3& a, const V3& b) {
  double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return std::sqrt(dx * dx + dy * dy + dz * dz);
}
inline double distanceSquared(const V3& a, const V3& b) {
  double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}
inline V3 interpolate(const V3& t1, const V3& t2, double alpha) {
  return V3((1 - alpha) * t1.x + alpha * t2.x, (1 - alpha) * t1.y + alpha * t2.y, (1 - alpha) * t1.z + alpha * t2.z);
}

struct M3 {
  double m00 = 0, m01 = 0, m02 = 0, m10 = 0, m11 = 0, m12 = 0, m20 = 0, m21 = 0, m22 = 0;
  void setIdentity() { m00 = m11 = m22 = 1; m01 = m02 = m10 = m12 = m20 = m21 = 0; }
  void setZero() { m00 = m01 = m02 = m10 = m11 = m12 = m20 = m21 = m22 = 0; }
  double el(int r, int c) const { const double* p = &m00; return p[3 * r + c]; }
  V3 col(int c) const { return V3(el(0, c), el(1, c), el(2, c)); }
  void load(const double* p) { std::memcpy(&m00, p, 9 * sizeof(double)); }
  void store(double* p) const { std::memcpy(p, &m00, 9 * sizeof(double)); }
};

inline M3 mul(const M3& a, const M3& b) {
  M3 r;
  r.m00 = a.m00 * b.m00 + a.m01 * b.m10 + a.m02 * b.m20;
  r.m01 = a.m00 * b.m01 + a.m01 * b.m11 + a.m02 * b.m21;
  r.m02 = a.m00 * b.m02 + a.m01 * b.m12 + a.m02 * b.m22;
  r.m10 = a.m10 * b.m00 + a.m11 * b.m10 + a.m12 * b.m20;
  r.m11 = a.m10 * b.m01 + a.m11 * b.m11 + a.m12 * b.m21;
  r.m12 = a.m10 * b.m02 + a.m11 * b.m12 + a.m12 * b.m22;
  r.m20 = a.m20 * b.m00 + a.m21 * b.m10 + a.m22 * b.m20;
  r.m21 = a.m20 * b.m01 + a.m21 * b.m11 + a.m22 * b.m21;
  r.m22 = a.m20 * b.m02 + a.m21 * b.m12 + a.m22 * b.m22;
  return r;
}
inline M3 mulTransposeRight(const M3& a, const M3& b) {  // a * b^T
  M3 r;
  r.m00 = a.m00 * b.m00 + a.m01 * b.m01 + a.m02 * b.m02;
  r.m01 = a.m00 * b.m10 + a.m01 * b.m11 + a.m02 * b.m12;
  r.m02 = a.m00 * b.m20 + a.m01 * b.m21 + a.m02 * b.m22;
  r.m10 = a.m10 * b.m00 + a.m11 * b.m01 + a.m12 * b.m02;
  r.m11 = a.m10 * b.m10 + a.m11 * b.m11 + a.m12 * b.m12;
  r.m12 = a.m10 * b.m20 + a.m11 * b.m21 + a.m12 * b.m22;
  r.m20 = a.m20 * b.m00 + a.m21 * b.m01 + a.m22 * b.m02;
  r.m21 = a.m20 * b.m10 + a.m21 * b.m11 + a.m22 * b.m12;
  r.m22 = a.m20 * b.m20 + a.m21 * b.m21 + a.m22 * b.m22;
  return r;
}
inline M3 mulTransposeLeft(const M3& a, const M3& b) {  // a^T * b
  M3 r;
  r.m00 = a.m00 * b.m00 + a.m10 * b.m10 + a.m20 * b.m20;
  r.m01 = a.m00 * b.m01 + a.m10 * b.m11 + a.m20 * b.m21;
  r.m02 = a.m00 * b.m02 + a.m10 * b.m12 + a.m20 * b.m22;
  r.m10 = a.m01 * b.m00 + a.m11 * b.m10 + a.m21 * b.m20;
  r.m11 = a.m01 * b.m01 + a.m11 * b.m11 + a.m21 * b.m21;
  r.m12 = a.m01 * b.m02 + a.m11 * b.m12 + a.m21 * b.m22;
  r.m20 = a.m02 * b.m00 + a.m12 * b.m10 + a.m22 * b.m20;
  r.m21 = a.m02 * b.m01 + a.m12 * b.m11 + a.m22 * b.m21;
  r.m22 = a.m02 * b.m02 + a.m12 * b.m12 + a.m22 * b.m22;
  return r;
}
inline M3 transpose(const M3& a) {
  M3 r;
  r.m00 = a.m00; r.m01 = a.m10; r.m02 = a.m20;
  r.m10 = a.m01; r.m11 = a.m11; r.m12 = a.m21;
  r.m20 = a.m02; r.m21 = a.m12; r.m22 = a.m22;
  return r;
}
inline M3 addM(const M3& a, const M3& b) {
  M3 r;
  const double* pa = &a.m00; const double* pb = &b.m00; double* pr = &r.m00;
  for (int i = 0; i < 9; i++) pr[i] = pa[i] + pb[i];
  return r;
}
inline M3 subM(const M3& a, const M3& b) {
  M3 r;
  const double* pa = &a.m00; const double* pb = &b.m00; double* pr = &r.m00;
  for (int i = 0; i < 9; i++) pr[i] = pa[i] - pb[i];
  return r;
}
inline M3 scaleM(double s, const M3& a) {
  M3 r;
  const double* pa = &a.m00; double* pr = &r.m00;
  for (int i = 0; i < 9; i++) pr[i] = s * pa[i];
  return r;
}
inline V3 transform(const M3& m, const V3& t) {
  return V3(m.m00 * t.x + m.m01 * t.y + m.m02 * t.z, m.m10 * t.x + m.m11 * t.y + m.m12 * t.z,
            m.m20 * t.x + m.m21 * t.y + m.m22 * t.z);
}
inline M3 normalizeCP(const M3& m1) {
  M3 r;
  double mag = 1.0 / std::sqrt(m1.m00 * m1.m00 + m1.m10 * m1.m10 + m1.m20 * m1.m20);
  r.m00 = m1.m00 * mag; r.m10 = m1.m10 * mag; r.m20 = m1.m20 * mag;
  mag = 1.0 / std::sqrt(m1.m01 * m1.m01 + m1.m11 * m1.m11 + m1.m21 * m1.m21);
  r.m01 = m1.m01 * mag; r.m11 = m1.m11 * mag; r.m21 = m1.m21 * mag;
  r.m02 = r.m10 * r.m21 - r.m11 * r.m20;
  r.m12 = r.m01 * r.m20 - r.m00 * r.m21;
  r.m22 = r.m00 * r.m11 - r.m01 * r.m10;
  return r;
}

// Matrix3d.invertGeneral.  Returns false for a singular matrix (the reference throws).
inline bool invert(const M3& m, M3& out) {
  double a[9];
  m.store(a);
  double row_scale[3];
  for (int i = 0; i < 3; i++) {
    double big = 0.0;
    for (int j = 0; j < 3; j++) { double t = std::fabs(a[3 * i + j]); if (t > big) big = t; }
    if (big == 0.0) return false;
    row_scale[i] = 1.0 / big;
  }
  int perm[3];
  for (int j = 0; j < 3; j++) {
    for (int i = 0; i < j; i++) {
      double sum = a[3 * i + j];
      for (int k = 0; k < i; k++) sum -= a[3 * i + k] * a[3 * k + j];
      a[3 * i + j] = sum;
    }
    double big = 0.0;
    int imax = -1;
    for (int i = j; i < 3; i++) {
      double sum = a[3 * i + j];
      for (int k = 0; k < j; k++) sum -= a[3 * i + k] * a[3 * k + j];
      a[3 * i + j] = sum;
      double t = row_scale[i] * std::fabs(sum);
      if (t >= big) { big = t; imax = i; }
    }
    if (imax < 0) return false;
    if (j != imax) {
      for (int k = 0; k < 3; k++) { double t = a[3 * imax + k]; a[3 * imax + k] = a[3 * j + k]; a[3 * j + k] = t; }
      row_scale[imax] = row_scale[j];
    }
    perm[j] = imax;
    if (a[3 * j + j] == 0.0) return false;
    if (j != 2) {
      double t = 1.0 / a[3 * j + j];
      for (int i = j + 1; i < 3; i++) a[3 * i + j] *= t;
    }
  }
  double r[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int k = 0; k < 3; k++) {
    int ii = -1;
    for (int i = 0; i < 3; i++) {
      int ip = perm[i];
      double sum = r[k + 3 * ip];
      r[k + 3 * ip] = r[k + 3 * i];
      if (ii >= 0) {
        for (int j = ii; j <= i - 1; j++) sum -= a[3 * i + j] * r[k + 3 * j];
      } else if (sum != 0.0) {
        ii = i;
      }
      r[k + 3 * i] = sum;
    }
    r[k + 6] /= a[8];
    r[k + 3] = (r[k + 3] - a[5] * r[k + 6]) / a[4];
    r[k] = (r[k] - a[1] * r[k + 3] - a[2] * r[k + 6]) / a[0];
  }
  out.load(r);
  return true;
}

// RigidTransform3D restated as a value type (the reference aliases theta/x as backing memory).
struct Xf {
  M3 R;
  V3 t;
  Xf() { R.setIdentity(); }
  Xf(const M3& r, const V3& tt) : R(r), t(tt) {}
  V3 transformP(const V3& p) const { V3 q = transform(R, p); return V3(q.x + t.x, q.y + t.y, q.z + t.z); }
  V3 transformV(const V3& v) const { return transform(R, v); }
  V3 inverseTransformP(const V3& p) const {
    double x = p.x - t.x, y = p.y - t.y, z = p.z - t.z;
    return V3(R.m00 * x + R.m10 * y + R.m20 * z, R.m01 * x + R.m11 * y + R.m21 * z, R.m02 * x + R.m12 * y + R.m22 * z);
  }
  V3 inverseTransformV(const V3& v) const {
    double x = v.x, y = v.y, z = v.z;
    return V3(R.m00 * x + R.m10 * y + R.m20 * z, R.m01 * x + R.m11 * y + R.m21 * z, R.m02 * x + R.m12 * y + R.m22 * z);
  }
  // this = A * B   (RigidTransform3D.mult(A,B) :193-197)
  static Xf mult(const Xf& A, const Xf& B) {
    Xf r;
    V3 q = transform(A.R, B.t);
    r.t = V3(q.x + A.t.x, q.y + A.t.y, q.z + A.t.z);
    r.R = mul(A.R, B.R);
    return r;
  }
  // this = A^-1 * B (:204-209)
  static Xf multAinvB(const Xf& A, const Xf& B) {
    Xf r;
    V3 d = sub(B.t, A.t);
    M3 Rt = transpose(A.R);
    r.t = transform(Rt, d);
    r.R = mul(Rt, B.R);
    return r;
  }
  M3 computeRM0RT(const M3& M1) const { return mulTransposeRight(mul(R, M1), R); }
  M3 computeRTMR(const M3& M1) const { return mul(mulTransposeLeft(R, M1), R); }
};

}  // namespace amo
