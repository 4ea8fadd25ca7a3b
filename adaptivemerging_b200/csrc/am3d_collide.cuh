// Narrowphase primitives as device functions (kernel 2 of the north star).
//
// Behavioural contract (bit-exact contact sets): same predicates, same evaluation order, same tie-break
// constants as the reference's primitives —
//   box x box    collision/BoxBox.java:433-859   (15-axis SAT, face bias tinyOffset, edge bias fudgeFactor,
//                                                 reference/incident face clipping, info = emission index)
//   box x plane  collision/BoxPlane.java:27-46   (info = corner id, plane is body1)
//   box x sphere collision/BoxSphere.java:58-275 (deepest feature wins; overlap test for tree descent)
// written for one GPU thread per primitive pair: no static scratch, fixed-size local arrays, results
// returned through a small hit buffer owned by the caller.
#pragma once
#include "am3d_math.cuh"

struct Hit {
  d3 pos;
  d3 normal;
  double violation;
  int info;
};

#define AM_FUDGE 1.05
#define AM_TINY 1e-5
#define AM_LCA_EPS 0.0001

// Clip the quad p[8] against the rectangle (+-h0, +-h1); returns point count, points in ret[16].
__device__ inline int rectQuadClip(const double h[2], const double p[8], double ret[16]) {
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
        bool in0 = sign * pq[dir] < h[dir];
        if (in0) {
          pr[0] = pq[0];
          pr[1] = pq[1];
          pr += 2;
          nr++;
          if (nr & 8) { q = r; goto done; }
        }
        const double* nextq = (i > 1) ? pq + 2 : q;
        bool in1 = sign * nextq[dir] < h[dir];
        if (in0 ^ in1) {
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
  if (q != ret) {
    for (int i = 0; i < nr * 2; i++) ret[i] = q[i];
  }
  return nr;
}

struct SatState {
  double s;
  int code;
  int normSel;   // 0: normalC, 1: column of R1, 2: column of R2
  int normCol;
  d3 normalC;
  bool invert;
};

__device__ __forceinline__ bool satFace(double expr1, double expr2, int sel, int col, int cc, SatState& t) {
  double s2 = fabs(expr1) - expr2;
  if (s2 > 0) return false;
  if (s2 > t.s + AM_TINY) {
    t.s = s2;
    t.normSel = sel;
    t.normCol = col;
    t.invert = expr1 < 0;
    t.code = cc;
  }
  return true;
}
__device__ __forceinline__ bool satEdge(double expr1, double expr2, double n1, double n2, double n3, int cc, SatState& t) {
  double s2 = fabs(expr1) - expr2;
  if (s2 > 0) return false;
  double l = sqrt(n1 * n1 + n2 * n2 + n3 * n3);
  if (l > 0) {
    s2 /= l;
    if (s2 * AM_FUDGE > t.s) {
      t.s = s2;
      t.normSel = 0;
      t.normCol = 0;
      t.normalC = d3(n1 / l, n2 / l, n3 / l);
      t.invert = expr1 < 0;
      t.code = cc;
    }
  }
  return true;
}

// Returns the number of hits (0..8) handed to out.set(k, point, normal, info, violation), k = 0..n-1 in order.
template <class OUT>
__device__ inline int collideBoxBox(const d3& p1, const m3& R1, const d3& side1, double radius1, const d3& p2,
                                    const m3& R2, const d3& side2, double radius2, OUT& out) {
  d3 p = vsub(p2, p1);
  if (vlen(p) > radius1 + radius2) return 0;
  d3 pp = mtransformT(R1, p);
  d3 A = vscale(0.5, side1), B = vscale(0.5, side2);
  d3 u0 = R1.col(0), u1 = R1.col(1), u2 = R1.col(2);
  d3 w0 = R2.col(0), w1 = R2.col(1), w2 = R2.col(2);
  double R11 = vdot(u0, w0), R12 = vdot(u0, w1), R13 = vdot(u0, w2);
  double R21 = vdot(u1, w0), R22 = vdot(u1, w1), R23 = vdot(u1, w2);
  double R31 = vdot(u2, w0), R32 = vdot(u2, w1), R33 = vdot(u2, w2);
  double Q11 = fabs(R11), Q12 = fabs(R12), Q13 = fabs(R13);
  double Q21 = fabs(R21), Q22 = fabs(R22), Q23 = fabs(R23);
  double Q31 = fabs(R31), Q32 = fabs(R32), Q33 = fabs(R33);
  SatState t;
  t.s = -INFINITY;
  t.invert = false;
  t.code = 0;
  t.normSel = 0;
  t.normCol = 0;
  if (!satFace(pp.x, (A.x + B.x * Q11 + B.y * Q12 + B.z * Q13), 1, 0, 1, t)) return 0;
  if (!satFace(pp.y, (A.y + B.x * Q21 + B.y * Q22 + B.z * Q23), 1, 1, 2, t)) return 0;
  if (!satFace(pp.z, (A.z + B.x * Q31 + B.y * Q32 + B.z * Q33), 1, 2, 3, t)) return 0;
  if (!satFace(vdot(w0, p), (A.x * Q11 + A.y * Q21 + A.z * Q31 + B.x), 2, 0, 4, t)) return 0;
  if (!satFace(vdot(w1, p), (A.x * Q12 + A.y * Q22 + A.z * Q32 + B.y), 2, 1, 5, t)) return 0;
  if (!satFace(vdot(w2, p), (A.x * Q13 + A.y * Q23 + A.z * Q33 + B.z), 2, 2, 6, t)) return 0;
  if (!satEdge(pp.z * R21 - pp.y * R31, (A.y * Q31 + A.z * Q21 + B.y * Q13 + B.z * Q12), 0, -R31, R21, 7, t)) return 0;
  if (!satEdge(pp.z * R22 - pp.y * R32, (A.y * Q32 + A.z * Q22 + B.x * Q13 + B.z * Q11), 0, -R32, R22, 8, t)) return 0;
  if (!satEdge(pp.z * R23 - pp.y * R33, (A.y * Q33 + A.z * Q23 + B.x * Q12 + B.y * Q11), 0, -R33, R23, 9, t)) return 0;
  if (!satEdge(pp.x * R31 - pp.z * R11, (A.x * Q31 + A.z * Q11 + B.y * Q23 + B.z * Q22), R31, 0, -R11, 10, t)) return 0;
  if (!satEdge(pp.x * R32 - pp.z * R12, (A.x * Q32 + A.z * Q12 + B.x * Q23 + B.z * Q21), R32, 0, -R12, 11, t)) return 0;
  if (!satEdge(pp.x * R33 - pp.z * R13, (A.x * Q33 + A.z * Q13 + B.x * Q22 + B.y * Q21), R33, 0, -R13, 12, t)) return 0;
  if (!satEdge(pp.y * R11 - pp.x * R21, (A.x * Q21 + A.y * Q11 + B.y * Q33 + B.z * Q32), -R21, R11, 0, 13, t)) return 0;
  if (!satEdge(pp.y * R12 - pp.x * R22, (A.x * Q22 + A.y * Q12 + B.x * Q33 + B.z * Q31), -R22, R12, 0, 14, t)) return 0;
  if (!satEdge(pp.y * R13 - pp.x * R23, (A.x * Q23 + A.y * Q13 + B.x * Q32 + B.y * Q31), -R23, R13, 0, 15, t)) return 0;
  if (t.code == 0) return 0;

  d3 normal;
  if (t.normSel == 1) normal = R1.col(t.normCol);
  else if (t.normSel == 2) normal = R2.col(t.normCol);
  else normal = mtransform(R1, t.normalC);
  if (t.invert) normal = vscale(-1, normal);
  double depth = -t.s;

  if (t.code > 6) {
    d3 pa = p1;
    for (int j = 0; j < 3; j++) {
      double sign = (vdot(normal, R1.col(j)) > 0) ? 1.0 : -1.0;
      pa.x += sign * A.get(j) * R1.el(0, j);
      pa.y += sign * A.get(j) * R1.el(1, j);
      pa.z += sign * A.get(j) * R1.el(2, j);
    }
    d3 pb = p2;
    for (int j = 0; j < 3; j++) {
      double sign = (vdot(normal, R2.col(j)) > 0) ? -1.0 : 1.0;
      pb.x += sign * B.get(j) * R2.el(0, j);
      pb.y += sign * B.get(j) * R2.el(1, j);
      pb.z += sign * B.get(j) * R2.el(2, j);
    }
    d3 ua = R1.col((t.code - 7) / 3), ub = R2.col((t.code - 7) % 3);
    double alpha, beta;
    {
      d3 pl = vsub(pb, pa);
      double uaub = vdot(ua, ub);
      double q1 = vdot(ua, pl);
      double q2 = -vdot(ub, pl);
      double d = 1 - uaub * uaub;
      if (d <= AM_LCA_EPS) { alpha = 0; beta = 0; }
      else { d = 1.0 / d; alpha = (q1 + uaub * q2) * d; beta = (uaub * q1 + q2) * d; }
    }
    pa = vscaleAdd(alpha, ua, pa);
    pb = vscaleAdd(beta, ub, pb);
    d3 pos = vadd(pa, pb);
    pos = vscale(0.5, pos);
    out.set(0, pos, normal, 0, -depth);
    return 1;
  }

  const bool ref1 = t.code <= 3;
  const m3& Ra = ref1 ? R1 : R2;
  const m3& Rb = ref1 ? R2 : R1;
  d3 pa = ref1 ? p1 : p2, pb = ref1 ? p2 : p1, Sa = ref1 ? A : B, Sb = ref1 ? B : A;
  d3 normal2 = ref1 ? normal : vscale(-1, normal);
  d3 nr = mtransformT(Rb, normal2);
  d3 anr(fabs(nr.x), fabs(nr.y), fabs(nr.z));
  int lanr, a1, a2;
  if (anr.y > anr.x) {
    if (anr.y > anr.z) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  } else {
    if (anr.x > anr.z) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
  }
  d3 center;
  if (nr.get(lanr) < 0) {
    for (int i = 0; i < 3; i++) center.setc(i, pb.get(i) - pa.get(i) + Sb.get(lanr) * Rb.el(i, lanr));
  } else {
    for (int i = 0; i < 3; i++) center.setc(i, pb.get(i) - pa.get(i) - Sb.get(lanr) * Rb.el(i, lanr));
  }
  int codeN = ref1 ? t.code - 1 : t.code - 4;
  int code1, code2;
  if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
  double quad[8];
  double c1 = vdot(center, Ra.col(code1));
  double c2 = vdot(center, Ra.col(code2));
  double m11 = vdot(Ra.col(code1), Rb.col(a1));
  double m12 = vdot(Ra.col(code1), Rb.col(a2));
  double m21 = vdot(Ra.col(code2), Rb.col(a1));
  double m22 = vdot(Ra.col(code2), Rb.col(a2));
  {
    double k1 = m11 * Sb.get(a1), k2 = m21 * Sb.get(a1), k3 = m12 * Sb.get(a2), k4 = m22 * Sb.get(a2);
    quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
    quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
    quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
    quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
  }
  double rect[2] = {Sa.get(code1), Sa.get(code2)};
  double ret[16];
  int n = rectQuadClip(rect, quad, ret);
  if (n < 1) return 0;
  double det1 = 1.0 / (m11 * m22 - m12 * m21);
  m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
  int cnum = 0;
  for (int j = 0; j < n; j++) {
    double k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
    double k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
    d3 pt;
    for (int i = 0; i < 3; i++) pt.setc(i, center.get(i) + k1 * Rb.el(i, a1) + k2 * Rb.el(i, a2));
    double dep = Sa.get(codeN) - (normal2.x * pt.x + normal2.y * pt.y + normal2.z * pt.z);
    if (dep >= 0) {
      out.set(cnum, d3(pt.x + pa.x, pt.y + pa.y, pt.z + pa.z), normal, cnum, -dep);
      cnum++;
    }
  }
  return cnum;
}

// box (T,size,radius) vs plane (n,d): hits per penetrating corner, info = corner id
template <class OUT>
__device__ inline int collideBoxPlane(const xf& T, const d3& size, double radius, const d3& n, double d, OUT& out) {
  if (T.t.x * n.x + T.t.y * n.y + T.t.z * n.z + d > radius) return 0;
  d3 p = vscale(0.5, size);
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    d3 q((k & 4) ? -p.x : p.x, (k & 2) ? -p.y : p.y, (k & 1) ? -p.z : p.z);
    q = xfP(T, q);
    double s = q.x * n.x + q.y * n.y + q.z * n.z + d;
    if (s < 0) {
      out.set(cnt, q, n, k, s);
      cnt++;
    }
  }
  return cnt;
}

struct BSRes {
  d3 pos, normal;
  double depth;
};
__device__ __forceinline__ void bsCorner(const d3& q, const d3& cB, double r, BSRes& tr) {
  double s = vdist(cB, q) - r;
  if (s > 0) return;
  if (s < tr.depth) {
    tr.pos = q;
    tr.normal = vnormalize(vsub(cB, q));
    tr.depth = s;
  }
}
__device__ __forceinline__ void bsFeature(double depth, double px, double py, double pz, double nx, double ny, double nz, BSRes& tr) {
  if (depth > 0) return;
  if (depth < tr.depth) {
    tr.pos = d3(px, py, pz);
    tr.normal = d3(nx, ny, nz);
    tr.depth = depth;
  }
}

// box vs sphere: 0 or 1 hit, normal from box to sphere
__device__ inline int collideBoxSphere(const xf& T, const d3& size, const d3& c, double r, Hit* out) {
  d3 p = vscale(0.5, size);
  d3 cB = xfInvP(T, c);
  BSRes tr;
  tr.depth = 1;
#pragma unroll
  for (int k = 0; k < 8; k++) bsCorner(d3((k & 4) ? -p.x : p.x, (k & 2) ? -p.y : p.y, (k & 1) ? -p.z : p.z), cB, r, tr);
  double s;
  d3 v;
  if (-p.x <= cB.x && cB.x <= p.x) {
    if (cB.y >= p.y && cB.z >= p.z) { v = d3(0, cB.y - p.y, cB.z - p.z); s = vlen(v); s -= r; bsFeature(s, cB.x, p.y, p.z, v.x, v.y, v.z, tr); }
    else if (cB.y <= -p.y && cB.z >= p.z) { v = d3(0, cB.y + p.y, cB.z - p.z); s = vlen(v); s -= r; bsFeature(s, cB.x, -p.y, p.z, v.x, v.y, v.z, tr); }
    else if (cB.y >= p.y && cB.z <= -p.z) { v = d3(0, cB.y - p.y, cB.z + p.z); s = vlen(v); s -= r; bsFeature(s, cB.x, p.y, -p.z, v.x, v.y, v.z, tr); }
    else if (cB.y <= -p.y && cB.z <= -p.z) { v = d3(0, cB.y + p.y, cB.z + p.z); s = vlen(v); s -= r; bsFeature(s, cB.x, -p.y, -p.z, v.x, v.y, v.z, tr); }
  }
  if (-p.y <= cB.y && cB.y <= p.y) {
    if (cB.x >= p.x && cB.z >= p.z) { v = d3(cB.x - p.x, 0, cB.z - p.z); s = vlen(v); s -= r; bsFeature(s, p.x, cB.y, p.z, v.x, v.y, v.z, tr); }
    else if (cB.x <= -p.x && cB.z >= p.z) { v = d3(cB.x + p.x, 0, cB.z - p.z); s = vlen(v); s -= r; bsFeature(s, -p.x, cB.y, p.z, v.x, v.y, v.z, tr); }
    else if (cB.x >= p.x && cB.z <= -p.z) { v = d3(cB.x - p.x, 0, cB.z + p.z); s = vlen(v); s -= r; bsFeature(s, p.x, cB.y, -p.z, v.x, v.y, v.z, tr); }
    else if (cB.x <= -p.x && cB.z <= -p.z) { v = d3(cB.x + p.x, 0, cB.z + p.z); s = vlen(v); s -= r; bsFeature(s, -p.x, cB.y, -p.z, v.x, v.y, v.z, tr); }
  }
  if (-p.z <= cB.z && cB.z <= p.z) {
    if (cB.x >= p.x && cB.y >= p.y) { v = d3(cB.x - p.x, cB.y - p.y, 0); s = vlen(v); s -= r; bsFeature(s, p.x, p.y, cB.z, v.x, v.y, v.z, tr); }
    else if (cB.x <= -p.x && cB.y >= p.y) { v = d3(cB.x + p.x, cB.y - p.y, 0); s = vlen(v); s -= r; bsFeature(s, -p.x, p.y, cB.z, v.x, v.y, v.z, tr); }
    else if (cB.x >= p.x && cB.y <= -p.y) { v = d3(cB.x - p.x, cB.y + p.y, 0); s = vlen(v); s -= r; bsFeature(s, p.x, -p.y, cB.z, v.x, v.y, v.z, tr); }
    else if (cB.x <= -p.x && cB.y <= -p.y) { v = d3(cB.x + p.x, cB.y + p.y, 0); s = vlen(v); s -= r; bsFeature(s, -p.x, -p.y, cB.z, v.x, v.y, v.z, tr); }
  }
  if (-p.x <= cB.x && cB.x <= p.x && -p.y <= cB.y && cB.y <= p.y) {
    if (cB.z > 0) { s = cB.z - p.z - r; bsFeature(s, cB.x, cB.y, p.z, 0, 0, 1, tr); }
    else { s = -p.z - cB.z - r; bsFeature(s, cB.x, cB.y, -p.z, 0, 0, -1, tr); }
  }
  if (-p.x <= cB.x && cB.x <= p.x && -p.z <= cB.z && cB.z <= p.z) {
    if (cB.y > 0) { s = cB.y - p.y - r; bsFeature(s, cB.x, p.y, cB.z, 0, 1, 0, tr); }
    else { s = -p.y - cB.y - r; bsFeature(s, cB.x, -p.y, cB.z, 0, -1, 0, tr); }
  }
  if (-p.y <= cB.y && cB.y <= p.y && -p.z <= cB.z && cB.z <= p.z) {
    if (cB.x > 0) { s = cB.x - p.x - r; bsFeature(s, p.x, cB.y, cB.z, 1, 0, 0, tr); }
    else { s = -p.x - cB.x - r; bsFeature(s, -p.x, cB.y, cB.z, -1, 0, 0, tr); }
  }
  if (tr.depth != 1) {
    out[0].pos = xfP(T, tr.pos);
    out[0].normal = vnormalize(mtransform(T.R, tr.normal));
    out[0].info = 0;
    out[0].violation = tr.depth;
    return 1;
  }
  return 0;
}

// overlap test used while descending a sphere tree against a box
__device__ inline bool overlapBoxSphere(const xf& T, const d3& size, double boxRadius, const d3& c, double r) {
  if (vdist(c, T.t) > boxRadius + r) return false;
  d3 p = vscale(0.5, size);
  d3 cB = xfInvP(T, c);
  if (cB.x - p.x > r) return false;
  if (-p.x - cB.x > r) return false;
  if (cB.y - p.y > r) return false;
  if (-p.y - cB.y > r) return false;
  if (cB.z - p.z > r) return false;
  if (-p.z - cB.z > r) return false;
#pragma unroll
  for (int k = 0; k < 8; k++)
    if (vdist(cB, d3((k & 4) ? -p.x : p.x, (k & 2) ? -p.y : p.y, (k & 1) ? -p.z : p.z)) - r < 0) return true;
  d3 v;
  if (-p.x <= cB.x && cB.x <= p.x) {
    if (cB.y >= p.y && cB.z >= p.z) { v = d3(0, cB.y - p.y, cB.z - p.z); if (vlen(v) - r < 0) return true; }
    else if (cB.y <= -p.y && cB.z >= p.z) { v = d3(0, cB.y + p.y, cB.z - p.z); if (vlen(v) - r < 0) return true; }
    else if (cB.y >= p.y && cB.z <= -p.z) { v = d3(0, cB.y - p.y, cB.z + p.z); if (vlen(v) - r < 0) return true; }
    else if (cB.y <= -p.y && cB.z <= -p.z) { v = d3(0, cB.y + p.y, cB.z + p.z); if (vlen(v) - r < 0) return true; }
  }
  if (-p.y <= cB.y && cB.y <= p.y) {
    if (cB.x >= p.x && cB.z >= p.z) { v = d3(cB.x - p.x, 0, cB.z - p.z); if (vlen(v) - r < 0) return true; }
    else if (cB.x <= -p.x && cB.z >= p.z) { v = d3(cB.x + p.x, 0, cB.z - p.z); if (vlen(v) - r < 0) return true; }
    else if (cB.x >= p.x && cB.z <= -p.z) { v = d3(cB.x - p.x, 0, cB.z + p.z); if (vlen(v) - r < 0) return true; }
    else if (cB.x <= -p.x && cB.z <= -p.z) { v = d3(cB.x + p.x, 0, cB.z + p.z); if (vlen(v) - r < 0) return true; }
  }
  if (-p.z < cB.z && cB.z < p.z) {  // strict, unlike collideBoxSphere (BoxSphere.java:202 vs :94)
    if (cB.x >= p.x && cB.y >= p.y) { v = d3(cB.x - p.x, cB.y - p.y, 0); if (vlen(v) - r < 0) return true; }
    else if (cB.x <= -p.x && cB.y >= p.y) { v = d3(cB.x + p.x, cB.y - p.y, 0); if (vlen(v) - r < 0) return true; }
    else if (cB.x >= p.x && cB.y <= -p.y) { v = d3(cB.x - p.x, cB.y + p.y, 0); if (vlen(v) - r < 0) return true; }
    else if (cB.x <= -p.x && cB.y <= -p.y) { v = d3(cB.x + p.x, cB.y + p.y, 0); if (vlen(v) - r < 0) return true; }
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
