// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h header).  PARITY UNPINNED.
//
// Collision primitives restated from the reference, evaluation order preserved:
//   collision/BoxBox.java   dBoxBox :433-859, intersectRectQuad :140-209, TST1/TST2 :291-336,
//                           dLineClosestApproach :366-383 (constants :50-52)
//   collision/BoxPlane.java dBoxPlane :27-46
//   collision/BoxSphere.java dBoxSphere :58-144, dBoxSphereTest :156-233, test1/test2 :242-275
// Each primitive reports raw hits through an Emit callback; Contact.set is applied by the caller.
#pragma once
#include <cmath>
#include <functional>

#include "oracle_model.h"

namespace amo {

struct Hit {
  V3 pos, normal;
  int info;
  double violation;
};

static const double kFudgeFactor = 1.05;
static const double kTinyOffset = 1e-5;
static const double kLineClosestApproachEPS = 0.0001;

// find all the intersection points between the 2D rectangle with vertices at (+/-h[0],+/-h[1])
// and the 2D quadrilateral p[0..7]; returns the count, points in ret[0..15]
inline int intersectRectQuad(const double h[2], const double p[8], double ret[16]) {
  double buffer[16];
  int nq = 4, nr = 0;
  const double* q = p;
  double* r = ret;
  for (int dir = 0; dir <= 1; dir++) {
    for (int sign = -1; sign <= 1; sign += 2) {
      const double* pq = q;
      double* pr = r;
      nr = 0;
      for (int i = nq; i > 0; i--) {
        if (sign * pq[dir] < h[dir]) {
          pr[0] = pq[0];
          pr[1] = pq[1];
          pr += 2;
          nr++;
          if (nr & 8) { q = r; goto done; }
        }
        const double* nextq = (i > 1) ? pq + 2 : q;
        if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {
          pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
          pr[dir] = sign * h[dir];
          pr += 2;
          nr++;
          if (nr & 8) { q = r; goto done; }
        }
        pq += 2;
      }
      q = r;
      r = (q == ret) ? buffer : ret;
      nq = nr;
    }
  }
done:
  if (q != ret) std::memcpy(ret, q, nr * 2 * sizeof(double));
  return nr;
}

struct Tst {
  int code = 0;
  bool normalFromR = false;  // _normalR_M != null
  const M3* normalR_M = nullptr;
  int normalR_col = 0;
  V3 normalC;
  double s = 0;
  bool invert_normal = false;
};

inline bool TST1(double expr1, double expr2, const M3* normA, int normO, int cc, Tst& t) {
  double s2 = std::fabs(expr1) - expr2;
  if (s2 > 0) return false;
  if (s2 > t.s + kTinyOffset) {
    t.s = s2;
    t.normalR_M = normA;
    t.normalR_col = normO;
    t.invert_normal = expr1 < 0;
    t.code = cc;
  }
  return true;
}
inline bool TST2(double expr1, double expr2, double n1, double n2, double n3, int cc, Tst& t) {
  double s2 = std::fabs(expr1) - expr2;
  if (s2 > 0) return false;
  double l = std::sqrt(n1 * n1 + n2 * n2 + n3 * n3);
  if (l > 0) {
    s2 /= l;
    if (s2 * kFudgeFactor > t.s) {
      t.s = s2;
      t.normalR_M = nullptr;
      t.normalR_col = 0;
      t.normalC = V3(n1 / l, n2 / l, n3 / l);
      t.invert_normal = expr1 < 0;
      t.code = cc;
    }
  }
  return true;
}

inline double dDOT44(const M3& a, int o1, const M3& b, int o2) { return dot(a.col(o1), b.col(o2)); }
inline double dDOT41(const M3& a, int o1, const V3& b) { return dot(a.col(o1), b); }
inline double dDOT14(const V3& a, const M3& b, int o2) { return dot(a, b.col(o2)); }

// dBoxBox: body1 = (p1,R1,side1,radius1), body2 likewise.  Emits up to 8 hits, info = emission index.
template <class Emit>
int dBoxBox(const V3& p1, const M3& R1, const V3& side1, double radius1, const V3& p2, const M3& R2, const V3& side2,
            double radius2, Emit emit) {
  int info = 0;
  V3 p = sub(p2, p1);
  if (length(p) > radius1 + radius2) return 0;
  // pp = R1^T p
  V3 pp(R1.m00 * p.x + R1.m10 * p.y + R1.m20 * p.z, R1.m01 * p.x + R1.m11 * p.y + R1.m21 * p.z,
        R1.m02 * p.x + R1.m12 * p.y + R1.m22 * p.z);
  V3 A = scale(0.5, side1), B = scale(0.5, side2);
  double R11 = dDOT44(R1, 0, R2, 0), R12 = dDOT44(R1, 0, R2, 1), R13 = dDOT44(R1, 0, R2, 2);
  double R21 = dDOT44(R1, 1, R2, 0), R22 = dDOT44(R1, 1, R2, 1), R23 = dDOT44(R1, 1, R2, 2);
  double R31 = dDOT44(R1, 2, R2, 0), R32 = dDOT44(R1, 2, R2, 1), R33 = dDOT44(R1, 2, R2, 2);
  double Q11 = std::fabs(R11), Q12 = std::fabs(R12), Q13 = std::fabs(R13);
  double Q21 = std::fabs(R21), Q22 = std::fabs(R22), Q23 = std::fabs(R23);
  double Q31 = std::fabs(R31), Q32 = std::fabs(R32), Q33 = std::fabs(R33);

  Tst tst;
  tst.s = -INFINITY;
  tst.invert_normal = false;
  tst.code = 0;
  if (!TST1(pp.x, (A.x + B.x * Q11 + B.y * Q12 + B.z * Q13), &R1, 0, 1, tst)) return 0;
  if (!TST1(pp.y, (A.y + B.x * Q21 + B.y * Q22 + B.z * Q23), &R1, 1, 2, tst)) return 0;
  if (!TST1(pp.z, (A.z + B.x * Q31 + B.y * Q32 + B.z * Q33), &R1, 2, 3, tst)) return 0;
  if (!TST1(dDOT41(R2, 0, p), (A.x * Q11 + A.y * Q21 + A.z * Q31 + B.x), &R2, 0, 4, tst)) return 0;
  if (!TST1(dDOT41(R2, 1, p), (A.x * Q12 + A.y * Q22 + A.z * Q32 + B.y), &R2, 1, 5, tst)) return 0;
  if (!TST1(dDOT41(R2, 2, p), (A.x * Q13 + A.y * Q23 + A.z * Q33 + B.z), &R2, 2, 6, tst)) return 0;
  if (!TST2(pp.z * R21 - pp.y * R31, (A.y * Q31 + A.z * Q21 + B.y * Q13 + B.z * Q12), 0, -R31, R21, 7, tst)) return 0;
  if (!TST2(pp.z * R22 - pp.y * R32, (A.y * Q32 + A.z * Q22 + B.x * Q13 + B.z * Q11), 0, -R32, R22, 8, tst)) return 0;
  if (!TST2(pp.z * R23 - pp.y * R33, (A.y * Q33 + A.z * Q23 + B.x * Q12 + B.y * Q11), 0, -R33, R23, 9, tst)) return 0;
  if (!TST2(pp.x * R31 - pp.z * R11, (A.x * Q31 + A.z * Q11 + B.y * Q23 + B.z * Q22), R31, 0, -R11, 10, tst)) return 0;
  if (!TST2(pp.x * R32 - pp.z * R12, (A.x * Q32 + A.z * Q12 + B.x * Q23 + B.z * Q21), R32, 0, -R12, 11, tst)) return 0;
  if (!TST2(pp.x * R33 - pp.z * R13, (A.x * Q33 + A.z * Q13 + B.x * Q22 + B.y * Q21), R33, 0, -R13, 12, tst)) return 0;
  if (!TST2(pp.y * R11 - pp.x * R21, (A.x * Q21 + A.y * Q11 + B.y * Q33 + B.z * Q32), -R21, R11, 0, 13, tst)) return 0;
  if (!TST2(pp.y * R12 - pp.x * R22, (A.x * Q22 + A.y * Q12 + B.x * Q33 + B.z * Q31), -R22, R12, 0, 14, tst)) return 0;
  if (!TST2(pp.y * R13 - pp.x * R23, (A.x * Q23 + A.y * Q13 + B.x * Q32 + B.y * Q31), -R23, R13, 0, 15, tst)) return 0;
  if (tst.code == 0) return 0;

  V3 normal;
  if (tst.normalR_M != nullptr) normal = tst.normalR_M->col(tst.normalR_col);
  else normal = transform(R1, tst.normalC);
  if (tst.invert_normal) normal = scale(-1, normal);
  double depth = -tst.s;

  if (tst.code > 6) {
    V3 pa = p1;
    for (int j = 0; j < 3; j++) {
      double sign = (dDOT14(normal, R1, j) > 0) ? 1.0 : -1.0;
      pa.x += sign * A.get(j) * R1.el(0, j);
      pa.y += sign * A.get(j) * R1.el(1, j);
      pa.z += sign * A.get(j) * R1.el(2, j);
    }
    V3 pb = p2;
    for (int j = 0; j < 3; j++) {
      double sign = (dDOT14(normal, R2, j) > 0) ? -1.0 : 1.0;
      pb.x += sign * B.get(j) * R2.el(0, j);
      pb.y += sign * B.get(j) * R2.el(1, j);
      pb.z += sign * B.get(j) * R2.el(2, j);
    }
    V3 ua = R1.col((tst.code - 7) / 3), ub = R2.col((tst.code - 7) % 3);
    // dLineClosestApproach
    double alpha, beta;
    {
      V3 pl = sub(pb, pa);
      double uaub = dot(ua, ub);
      double q1 = dot(ua, pl);
      double q2 = -dot(ub, pl);
      double d = 1 - uaub * uaub;
      if (d <= kLineClosestApproachEPS) { alpha = 0; beta = 0; }
      else { d = 1.0 / d; alpha = (q1 + uaub * q2) * d; beta = (uaub * q1 + q2) * d; }
    }
    pa = scaleAdd(alpha, ua, pa);
    pb = scaleAdd(beta, ub, pb);
    V3 pos = add(pa, pb);
    pos = scale(0.5, pos);
    emit(Hit{pos, normal, info++, -depth});
    return 1;
  }

  const M3 *Ra, *Rb;
  V3 pa, pb, Sa, Sb;
  if (tst.code <= 3) { Ra = &R1; Rb = &R2; pa = p1; pb = p2; Sa = A; Sb = B; }
  else { Ra = &R2; Rb = &R1; pa = p2; pb = p1; Sa = B; Sb = A; }
  V3 normal2 = (tst.code <= 3) ? normal : scale(-1, normal);
  V3 nr(Rb->m00 * normal2.x + Rb->m10 * normal2.y + Rb->m20 * normal2.z,
        Rb->m01 * normal2.x + Rb->m11 * normal2.y + Rb->m21 * normal2.z,
        Rb->m02 * normal2.x + Rb->m12 * normal2.y + Rb->m22 * normal2.z);
  V3 anr(std::fabs(nr.x), std::fabs(nr.y), std::fabs(nr.z));
  int lanr, a1, a2;
  if (anr.y > anr.x) {
    if (anr.y > anr.z) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  } else {
    if (anr.x > anr.z) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  }
  V3 center;
  if (nr.get(lanr) < 0) {
    for (int i = 0; i < 3; i++) center.setc(i, pb.get(i) - pa.get(i) + Sb.get(lanr) * Rb->el(i, lanr));
  } else {
    for (int i = 0; i < 3; i++) center.setc(i, pb.get(i) - pa.get(i) - Sb.get(lanr) * Rb->el(i, lanr));
  }
  int codeN, code1, code2;
  if (tst.code <= 3) codeN = tst.code - 1; else codeN = tst.code - 4;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  double quad[8];
  double c1 = dDOT14(center, *Ra, code1);
  double c2 = dDOT14(center, *Ra, code2);
  double m11 = dDOT44(*Ra, code1, *Rb, a1);
  double m12 = dDOT44(*Ra, code1, *Rb, a2);
  double m21 = dDOT44(*Ra, code2, *Rb, a1);
  double m22 = dDOT44(*Ra, code2, *Rb, a2);
  {
    double k1 = m11 * Sb.get(a1), k2 = m21 * Sb.get(a1), k3 = m12 * Sb.get(a2), k4 = m22 * Sb.get(a2);
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  double rect[2] = {Sa.get(code1), Sa.get(code2)};
  double ret[16];
  int n = intersectRectQuad(rect, quad, ret);
  if (n < 1) return 0;
  double point[24], dep[8];
  double det1 = 1.0 / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnum = 0;
  for (int j = 0; j < n; j++) {
    double k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    double k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    for (int i = 0; i < 3; i++) point[cnum * 3 + i] = center.get(i) + k1 * Rb->el(i, a1) + k2 * Rb->el(i, a2);
    dep[cnum] = Sa.get(codeN) - (normal2.x * point[cnum * 3] + normal2.y * point[cnum * 3 + 1] + normal2.z * point[cnum * 3 + 2]);
    if (dep[cnum] >= 0) {
      ret[cnum * 2] = ret[j * 2];
      ret[cnum * 2 + 1] = ret[j * 2 + 1];
      cnum++;
    }
  }
  if (cnum < 1) return 0;
  // maxc = 0xffff (BoxBox.java:808): cullPoints is dead code, every point is kept
  for (int j = 0; j < cnum; j++) {
    V3 pos(point[j * 3] + pa.x, point[j * 3 + 1] + pa.y, point[j * 3 + 2] + pa.z);
    emit(Hit{pos, normal, info++, -dep[j]});
  }
  return cnum;
}

// dBoxPlane: box (xf, size, radius) against plane (n, d).  info = corner id.  Plane is body1 of the contact.
template <class Emit>
int dBoxPlane(const Xf& TB2W, const V3& size, double radius, const V3& n, double d, Emit emit) {
  if (TB2W.t.x * n.x + TB2W.t.y * n.y + TB2W.t.z * n.z + d > radius) return 0;
  V3 p = scale(0.5, size);
  int cnt = 0;
  static const double sg[8][3] = {{1, 1, 1}, {1, 1, -1}, {1, -1, 1}, {1, -1, -1}, {-1, 1, 1}, {-1, 1, -1}, {-1, -1, 1}, {-1, -1, -1}};
  for (int k = 0; k < 8; k++) {
    V3 q(sg[k][0] > 0 ? p.x : -p.x, sg[k][1] > 0 ? p.y : -p.y, sg[k][2] > 0 ? p.z : -p.z);
    q = TB2W.transformP(q);
    double s = q.x * n.x + q.y * n.y + q.z * n.z + d;
    if (s < 0) { emit(Hit{q, n, k, s}); cnt++; }
  }
  return cnt;
}

struct BSResult {
  V3 pos, normal;
  double depth;
};
inline void bsTest1(const V3& q, const V3& cB, double r, BSResult& tr) {
  double s = distance(cB, q) - r;
  if (s > 0) return;
  if (s < tr.depth) {
    tr.pos = q;
    tr.normal = normalize(sub(cB, q));
    tr.depth = s;
  }
}
inline void bsTest2(double depth, double px, double py, double pz, double nx, double ny, double nz, BSResult& tr) {
  if (depth > 0) return;
  if (depth < tr.depth) {
    tr.pos = V3(px, py, pz);
    tr.normal = V3(nx, ny, nz);
    tr.depth = depth;
  }
}

// dBoxSphere: box body1 (xf,size) vs sphere (c,r): at most one hit, normal box -> sphere
template <class Emit>
int dBoxSphere(const Xf& TB2W, const V3& size, const V3& c, double r, Emit emit) {
  V3 p = scale(0.5, size);
  V3 cB = TB2W.inverseTransformP(c);
  BSResult tr;
  tr.depth = 1;
  bsTest1(V3(p.x, p.y, p.z), cB, r, tr);
  bsTest1(V3(p.x, p.y, -p.z), cB, r, tr);
  bsTest1(V3(p.x, -p.y, p.z), cB, r, tr);
  bsTest1(V3(p.x, -p.y, -p.z), cB, r, tr);
  bsTest1(V3(-p.x, p.y, p.z), cB, r, tr);
  bsTest1(V3(-p.x, p.y, -p.z), cB, r, tr);
  bsTest1(V3(-p.x, -p.y, p.z), cB, r, tr);
  bsTest1(V3(-p.x, -p.y, -p.z), cB, r, tr);
  double s;
  V3 v;
  if (-p.x <= cB.x && cB.x <= p.x) {
    if (cB.y >= p.y && cB.z >= p.z) { v = V3(0, cB.y - p.y, cB.z - p.z); s = length(v); s -= r; bsTest2(s, cB.x, p.y, p.z, v.x, v.y, v.z, tr); }
    else if (cB.y <= -p.y && cB.z >= p.z) { v = V3(0, cB.y + p.y, cB.z - p.z); s = length(v); s -= r; bsTest2(s, cB.x, -p.y, p.z, v.x, v.y, v.z, tr); }
    else if (cB.y >= p.y && cB.z <= -p.z) { v = V3(0, cB.y - p.y, cB.z + p.z); s = length(v); s -= r; bsTest2(s, cB.x, p.y, -p.z, v.x, v.y, v.z, tr); }
    else if (cB.y <= -p.y && cB.z <= -p.z) { v = V3(0, cB.y + p.y, cB.z + p.z); s = length(v); s -= r; bsTest2(s, cB.x, -p.y, -p.z, v.x, v.y, v.z, tr); }
  }
  if (-p.y <= cB.y && cB.y <= p.y) {
    if (cB.x >= p.x && cB.z >= p.z) { v = V3(cB.x - p.x, 0, cB.z - p.z); s = length(v); s -= r; bsTest2(s, p.x, cB.y, p.z, v.x, v.y, v.z, tr); }
    else if (cB.x <= -p.x && cB.z >= p.z) { v = V3(cB.x + p.x, 0, cB.z - p.z); s = length(v); s -= r; bsTest2(s, -p.x, cB.y, p.z, v.x, v.y, v.z, tr); }
    else if (cB.x >= p.x && cB.z <= -p.z) { v = V3(cB.x - p.x, 0, cB.z + p.z); s = length(v); s -= r; bsTest2(s, p.x, cB.y, -p.z, v.x, v.y, v.z, tr); }
    else if (cB.x <= -p.x && cB.z <= -p.z) { v = V3(cB.x + p.x, 0, cB.z + p.z); s = length(v); s -= r; bsTest2(s, -p.x, cB.y, -p.z, v.x, v.y, v.z, tr); }
  }
  if (-p.z <= cB.z && cB.z <= p.z) {
    if (cB.x >= p.x && cB.y >= p.y) { v = V3(cB.x - p.x, cB.y - p.y, 0); s = length(v); s -= r; bsTest2(s, p.x, p.y, cB.z, v.x, v.y, v.z, tr); }
    else if (cB.x <= -p.x && cB.y >= p.y) { v = V3(cB.x + p.x, cB.y - p.y, 0); s = length(v); s -= r; bsTest2(s, -p.x, p.y, cB.z, v.x, v.y, v.z, tr); }
    else if (cB.x >= p.x && cB.y <= -p.y) { v = V3(cB.x - p.x, cB.y + p.y, 0); s = length(v); s -= r; bsTest2(s, p.x, -p.y, cB.z, v.x, v.y, v.z, tr); }
    else if (cB.x <= -p.x && cB.y <= -p.y) { v = V3(cB.x + p.x, cB.y + p.y, 0); s = length(v); s -= r; bsTest2(s, -p.x, -p.y, cB.z, v.x, v.y, v.z, tr); }
  }
  if (-p.x <= cB.x && cB.x <= p.x && -p.y <= cB.y && cB.y <= p.y) {
    if (cB.z > 0) { s = cB.z - p.z - r; bsTest2(s, cB.x, cB.y, p.z, 0, 0, 1, tr); }
    else { s = -p.z - cB.z - r; bsTest2(s, cB.x, cB.y, -p.z, 0, 0, -1, tr); }
  }
  if (-p.x <= cB.x && cB.x <= p.x && -p.z <= cB.z && cB.z <= p.z) {
    if (cB.y > 0) { s = cB.y - p.y - r; bsTest2(s, cB.x, p.y, cB.z, 0, 1, 0, tr); }
    else { s = -p.y - cB.y - r; bsTest2(s, cB.x, -p.y, cB.z, 0, -1, 0, tr); }
  }
  if (-p.y <= cB.y && cB.y <= p.y && -p.z <= cB.z && cB.z <= p.z) {
    if (cB.x > 0) { s = cB.x - p.x - r; bsTest2(s, p.x, cB.y, cB.z, 1, 0, 0, tr); }
    else { s = -p.x - cB.x - r; bsTest2(s, -p.x, cB.y, cB.z, -1, 0, 0, tr); }
  }
  if (tr.depth != 1) {
    V3 pos = TB2W.transformP(tr.pos);
    V3 nrm = TB2W.transformV(tr.normal);
    nrm = normalize(nrm);
    emit(Hit{pos, nrm, 0, tr.depth});
    return 1;
  }
  return 0;
}

inline bool dBoxSphereTest(const Xf& TB2W, const V3& size, double boxRadius, const V3& c, double r) {
  if (distance(c, TB2W.t) > boxRadius + r) return false;
  V3 p = scale(0.5, size);
  V3 cB = TB2W.inverseTransformP(c);
  if (cB.x - p.x > r) return false;
  if (-p.x - cB.x > r) return false;
  if (cB.y - p.y > r) return false;
  if (-p.y - cB.y > r) return false;
  if (cB.z - p.z > r) return false;
  if (-p.z - cB.z > r) return false;
  if (distance(cB, V3(p.x, p.y, p.z)) - r < 0) return true;
  if (distance(cB, V3(p.x, p.y, -p.z)) - r < 0) return true;
  if (distance(cB, V3(p.x, -p.y, p.z)) - r < 0) return true;
  if (distance(cB, V3(p.x, -p.y, -p.z)) - r < 0) return true;
  if (distance(cB, V3(-p.x, p.y, p.z)) - r < 0) return true;
  if (distance(cB, V3(-p.x, p.y, -p.z)) - r < 0) return true;
  if (distance(cB, V3(-p.x, -p.y, p.z)) - r < 0) return true;
  if (distance(cB, V3(-p.x, -p.y, -p.z)) - r < 0) return true;
  V3 v;
  if (-p.x <= cB.x && cB.x <= p.x) {
    if (cB.y >= p.y && cB.z >= p.z) { v = V3(0, cB.y - p.y, cB.z - p.z); if (length(v) - r < 0) return true; }
    else if (cB.y <= -p.y && cB.z >= p.z) { v = V3(0, cB.y + p.y, cB.z - p.z); if (length(v) - r < 0) return true; }
    else if (cB.y >= p.y && cB.z <= -p.z) { v = V3(0, cB.y - p.y, cB.z + p.z); if (length(v) - r < 0) return true; }
    else if (cB.y <= -p.y && cB.z <= -p.z) { v = V3(0, cB.y + p.y, cB.z + p.z); if (length(v) - r < 0) return true; }
  }
  if (-p.y <= cB.y && cB.y <= p.y) {
    if (cB.x >= p.x && cB.z >= p.z) { v = V3(cB.x - p.x, 0, cB.z - p.z); if (length(v) - r < 0) return true; }
    else if (cB.x <= -p.x && cB.z >= p.z) { v = V3(cB.x + p.x, 0, cB.z - p.z); if (length(v) - r < 0) return true; }
    else if (cB.x >= p.x && cB.z <= -p.z) { v = V3(cB.x - p.x, 0, cB.z + p.z); if (length(v) - r < 0) return true; }
    else if (cB.x <= -p.x && cB.z <= -p.z) { v = V3(cB.x + p.x, 0, cB.z + p.z); if (length(v) - r < 0) return true; }
  }
  if (-p.z < cB.z && cB.z < p.z) {  // strict here, <= in dBoxSphere (BoxSphere.java:202 vs :94)
    if (cB.x >= p.x && cB.y >= p.y) { v = V3(cB.x - p.x, cB.y - p.y, 0); if (length(v) - r < 0) return true; }
    else if (cB.x <= -p.x && cB.y >= p.y) { v = V3(cB.x + p.x, cB.y - p.y, 0); if (length(v) - r < 0) return true; }
    else if (cB.x >= p.x && cB.y <= -p.y) { v = V3(cB.x - p.x, cB.y + p.y, 0); if (length(v) - r < 0) return true; }
    else if (cB.x <= -p.x && cB.y <= -p.y) { v = V3(cB.x + p.x, cB.y + p.y, 0); if (length(v) - r < 0) return true; }
  }
  if (-p.x <= cB.x && cB.x <= p.x && -p.y <= cB.y && cB.y <= p.y) {
    if (cB.z > 0) { if (cB.z - p.z - r < 0) return true; } else { if (-p.z - cB.z - r < 0) return true; }
  }
  if (-p.x <= cB.x && cB.x <= p.x && -p.z <= cB.z && cB.z <= p.z) {
    if (cB.y > 0) { if (cB.y - p.y - r < 0) return true; } else { if (-p.y - cB.y - r < 0) return true; }
  }
  if (-p.y <= cB.y && cB.y <= p.y && -p.z <= cB.z && cB.z <= p.z) {
    if (cB.x > 0) { if (cB.x - p.x - r < 0) return true; } else { if (-p.x - cB.x - r < 0) return true; }
  }
  return false;
}

}  // namespace amo
