// Device-side FP64 value types for the rigid-body step.
//
// Results must be bit-identical to Java's strict IEEE-754 double arithmetic (the reference never fuses
// a*b+c), so this translation unit is compiled with -fmad=false and every expression keeps the
// reference's left-to-right evaluation order (javax.vecmath 1.3.2: Matrix3d.mul :1524, mulTransposeRight
// :1693, mulTransposeLeft :1737, transform :2050, normalizeCP :1865, invertGeneral :1067; Vector3d.cross
// :103, normalize :134; Tuple3d.scaleAdd :277, interpolate :622; RigidTransform3D.java).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define AMD __host__ __device__ __forceinline__

struct d3 {
  double x, y, z;
  AMD d3() : x(0), y(0), z(0) {}
  AMD d3(double a, double b, double c) : x(a), y(b), z(c) {}
  AMD double get(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  AMD void setc(int i, double v) {
    if (i == 0) x = v; else if (i == 1) y = v; else z = v;
  }
};
AMD d3 ld3(const double* p) { return d3(p[0], p[1], p[2]); }
AMD void st3(double* p, const d3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
AMD d3 vsub(const d3& a, const d3& b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
AMD d3 vadd(const d3& a, const d3& b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
AMD d3 vscale(double s, const d3& a) { return d3(s * a.x, s * a.y, s * a.z); }
AMD d3 vscaleAdd(double s, const d3& t1, const d3& t2) { return d3(s * t1.x + t2.x, s * t1.y + t2.y, s * t1.z + t2.z); }
AMD double vdot(const d3& a, const d3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
AMD double vlen(const d3& a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
AMD d3 vcross(const d3& a, const d3& b) { return d3(a.y * b.z - a.z * b.y, b.x * a.z - b.z * a.x, a.x * b.y - a.y * b.x); }
AMD d3 vnormalize(const d3& v) {
  double norm = 1.0 / sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  return d3(v.x * norm, v.y * norm, v.z * norm);
}
AMD double vdist(const d3& a, const d3& b) {
  double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return sqrt(dx * dx + dy * dy + dz * dz);
}
AMD double vdist2(const d3& a, const d3& b) {
  double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}

struct m3 {
  double m[9];  // row-major: m[3*r+c]
  AMD double el(int r, int c) const { return m[3 * r + c]; }
  AMD d3 col(int c) const { return d3(m[c], m[3 + c], m[6 + c]); }
};
AMD m3 ldm(const double* p) {
  m3 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.m[i] = p[i];
  return r;
}
AMD void stm(double* p, const m3& a) {
#pragma unroll
  for (int i = 0; i < 9; i++) p[i] = a.m[i];
}
AMD m3 midentity() {
  m3 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.m[i] = (i % 4 == 0) ? 1.0 : 0.0;
  return r;
}
AMD m3 mzero() {
  m3 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.m[i] = 0.0;
  return r;
}
AMD m3 mmul(const m3& a, const m3& b) {
  m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return r;
}
AMD m3 mmulTR(const m3& a, const m3& b) {  // a * b^T
  m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * i] * b.m[3 * j] + a.m[3 * i + 1] * b.m[3 * j + 1] + a.m[3 * i + 2] * b.m[3 * j + 2];
  return r;
}
AMD m3 mmulTL(const m3& a, const m3& b) {  // a^T * b
  m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[i] * b.m[j] + a.m[3 + i] * b.m[3 + j] + a.m[6 + i] * b.m[6 + j];
  return r;
}
AMD m3 mtranspose(const m3& a) {
  m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * j + i];
  return r;
}
AMD m3 madd(const m3& a, const m3& b) {
  m3 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + b.m[i];
  return r;
}
AMD m3 mscale(double s, const m3& a) {
  m3 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.m[i] = s * a.m[i];
  return r;
}
AMD d3 mtransform(const m3& a, const d3& t) {
  return d3(a.m[0] * t.x + a.m[1] * t.y + a.m[2] * t.z, a.m[3] * t.x + a.m[4] * t.y + a.m[5] * t.z,
            a.m[6] * t.x + a.m[7] * t.y + a.m[8] * t.z);
}
AMD d3 mtransformT(const m3& a, const d3& t) {  // a^T * t, column dot products
  return d3(a.m[0] * t.x + a.m[3] * t.y + a.m[6] * t.z, a.m[1] * t.x + a.m[4] * t.y + a.m[7] * t.z,
            a.m[2] * t.x + a.m[5] * t.y + a.m[8] * t.z);
}
AMD m3 mnormalizeCP(const m3& a) {
  m3 r;
  double mag = 1.0 / sqrt(a.m[0] * a.m[0] + a.m[3] * a.m[3] + a.m[6] * a.m[6]);
  r.m[0] = a.m[0] * mag; r.m[3] = a.m[3] * mag; r.m[6] = a.m[6] * mag;
  mag = 1.0 / sqrt(a.m[1] * a.m[1] + a.m[4] * a.m[4] + a.m[7] * a.m[7]);
  r.m[1] = a.m[1] * mag; r.m[4] = a.m[4] * mag; r.m[7] = a.m[7] * mag;
  r.m[2] = r.m[3] * r.m[7] - r.m[4] * r.m[6];
  r.m[5] = r.m[1] * r.m[6] - r.m[0] * r.m[7];
  r.m[8] = r.m[0] * r.m[4] - r.m[1] * r.m[3];
  return r;
}
AMD m3 rm0rt(const m3& R, const m3& M) { return mmulTR(mmul(R, M), R); }
AMD m3 rtmr(const m3& R, const m3& M) { return mmul(mmulTL(R, M), R); }

// Crout LU with implicit scaling + back substitution, the algorithm vecmath uses for Matrix3d.invert
AMD bool minvert(const m3& in, m3& out) {
  double a[9];
#pragma unroll
  for (int i = 0; i < 9; i++) a[i] = in.m[i];
  double row_scale[3];
  for (int i = 0; i < 3; i++) {
    double big = 0.0;
    for (int j = 0; j < 3; j++) {
      double t = fabs(a[3 * i + j]);
      if (t > big) big = t;
    }
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
      double t = row_scale[i] * fabs(sum);
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
#pragma unroll
  for (int i = 0; i < 9; i++) out.m[i] = r[i];
  return true;
}

// rigid transform (R,t): p -> R p + t
struct xf {
  m3 R;
  d3 t;
};
AMD d3 xfP(const xf& T, const d3& p) {
  d3 q = mtransform(T.R, p);
  return d3(q.x + T.t.x, q.y + T.t.y, q.z + T.t.z);
}
AMD d3 xfInvP(const xf& T, const d3& p) {
  double x = p.x - T.t.x, y = p.y - T.t.y, z = p.z - T.t.z;
  return mtransformT(T.R, d3(x, y, z));
}
AMD xf xfMul(const xf& A, const xf& B) {
  xf r;
  d3 q = mtransform(A.R, B.t);
  r.t = d3(q.x + A.t.x, q.y + A.t.y, q.z + A.t.z);
  r.R = mmul(A.R, B.R);
  return r;
}
AMD xf xfAinvB(const xf& A, const xf& B) {
  xf r;
  d3 d = vsub(B.t, A.t);
  m3 Rt = mtranspose(A.R);
  r.t = mtransform(Rt, d);
  r.R = mmul(Rt, B.R);
  return r;
}

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256).  The sweep is bound by the number of scattered memory
// instructions per group, not by bytes: one 32-byte access per thread costs the load/store unit what an 8-byte one does.
__device__ __forceinline__ void ld4(const double* p, double* o) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p));
}
__device__ __forceinline__ void ld4cg(const double* p, double* o) {  // L2 only: written by other SMs between colours
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p));
}
__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

