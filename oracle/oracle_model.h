// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h header).  PARITY UNPINNED.
//
// Data model of mergingBodies3D restated with indices instead of Java object identity:
// RigidBody.java:22-170, RigidCollection.java:25-48, PlaneRigidBody.java:17-63, Contact.java:18-101,
// BodyPairContact.java:14-45, Spring.java:7-50.  Where the reference iterates a HashSet of
// identity-hashed objects (CollisionProcessor.bodyPairContacts :82, RigidBody.bodyPairContacts
// RigidBody.java:127, Merging.bpcsToUnmerge/subbodies/handledBodies Merging.java:210,275-277) the
// iteration order is unspecified in Java; this oracle canonicalises it to ascending (min body id,
// max body id) for body pairs and ascending body id for bodies.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "../include/am3d.h"
#include "oracle_math.h"

namespace amo {

struct Body;
struct BPC;

struct V6 {
  V3 v, w;
  void setZero() { v = V3(); w = V3(); }
};
inline double dot6(const V6& a, const V6& b) { return dot(a.v, b.v) + dot(a.w, b.w); }

enum GeomType { G_BOX = 0, G_TREE = 1, G_PLANE = 2, G_COMPOSITE = 3, G_NONE = 4 };

struct BpcLess {
  bool operator()(const BPC* a, const BPC* b) const;
};

struct Body {
  int id = -1;              // leaf body id (XML parse order); composite parts: id of the parent; collections: -1
  int part = -1;            // composite part index (subBodyID), -1 otherwise
  bool isCollection = false;
  long stamp = 0;           // position key in RigidBodySystem.bodies (list order == creation order)
  Body* parent = nullptr;            // RigidCollection
  Body* compositeBodyParent = nullptr;
  GeomType geom = G_NONE;
  V3 size;                  // RigidBodyGeomBox.size
  int root = -1;            // sphere-tree root node (global index)
  bool pinned = false, sleeping = false, magnetic = false, activateMagnet = false, picked = false;
  M3 massAngular0, massAngular, jinv0, jinv;
  double massLinear = 0, minv = 0;
  V3 force, torque;
  V3 x;
  M3 theta;
  V3 v, omega;
  V3 x0, v0, omega0;
  M3 theta0;
  Xf B2C;                   // transformB2C
  std::set<BPC*, BpcLess> bodyPairContacts;
  V6 deltaV;
  std::vector<V3> boundingBoxB;
  double friction = 0.8, restitution = 0;
  double radius = 0;
  std::vector<double> metricHistory;
  // PlaneRigidBody
  V3 n, p;
  double d = 0;
  // RigidBodyGeomComposite.bodies
  std::vector<Body*> parts;
  // RigidCollection
  std::vector<Body*> bodies;
  std::vector<struct Contact*> internalContacts;
  int collectionSlot = -1;  // stable id for export

  Body() { theta.setIdentity(); theta0.setIdentity(); }
  Xf B2W() const { return Xf(theta, x); }
  void setB2W(const Xf& t) { theta = t.R; x = t.t; }
  bool isInComposite() const { return compositeBodyParent != nullptr; }
  bool isInCollection() const { return parent != nullptr; }
  bool isInSameCollection(const Body* b) const { return parent != nullptr && parent == b->parent; }
  bool isPlane() const { return geom == G_PLANE; }
  void clear() { force = V3(); torque = V3(); deltaV.setZero(); }
};

enum ContactState { BROKEN = AM3D_CS_BROKEN, ONEDGE = AM3D_CS_ONEDGE, CLEAR = AM3D_CS_CLEAR };

struct Contact {
  Body* body1 = nullptr;
  Body* body2 = nullptr;
  int bv1 = AM3D_BV_NULL, bv2 = AM3D_BV_NULL;
  int info = 0;
  Body* csb1 = nullptr;
  Body* csb2 = nullptr;
  int leaf = -1;  // box x tree: leaf node hit (not part of the identity; canonical ordering only)
  V3 contactB1, normalB1, tangent1B1, tangent2B1;
  V3 pointW, normalW;  // world frame at Contact.set time (export only)
  double constraintViolation = 0, prevConstraintViolation = 0;
  ContactState state = CLEAR;
  bool newThisTimeStep = false;
  V6 jna, jnb, jt1a, jt1b, jt2a, jt2b;
  double lambda0 = 0, lambda1 = 0, lambda2 = 0;
  double lambda0warm = 0, lambda1warm = 0, lambda2warm = 0;
  double bn = 0, bt1 = 0, bt2 = 0;
  double D00 = 0, D11 = 0, D22 = 0;
  double w1 = 0, w2 = 0;
  bool internal = false;
};

// identity of a contact for warm starts: Contact.hashCode/equals (Contact.java:543-572)
struct ContactKey {
  int a[3], b[3], info;  // (body id, bv, csb part) of the two sides, side with the smaller triple first
  bool operator<(const ContactKey& o) const {
    for (int i = 0; i < 3; i++) if (a[i] != o.a[i]) return a[i] < o.a[i];
    for (int i = 0; i < 3; i++) if (b[i] != o.b[i]) return b[i] < o.b[i];
    return info < o.info;
  }
};
inline ContactKey makeKey(int b1, int bv1, int c1, int b2, int bv2, int c2, int info) {
  ContactKey k;
  int s1[3] = {b1, bv1, c1}, s2[3] = {b2, bv2, c2};
  bool firstSmaller = true;
  for (int i = 0; i < 3; i++) {
    if (s1[i] != s2[i]) { firstSmaller = s1[i] < s2[i]; break; }
  }
  for (int i = 0; i < 3; i++) { k.a[i] = firstSmaller ? s1[i] : s2[i]; k.b[i] = firstSmaller ? s2[i] : s1[i]; }
  k.info = info;
  return k;
}
inline ContactKey keyOf(const Contact* c, int infoOverride = -1) {
  return makeKey(c->body1->id, c->bv1, c->csb1 ? c->csb1->part : -1, c->body2->id, c->bv2,
                 c->csb2 ? c->csb2->part : -1, infoOverride >= 0 ? infoOverride : c->info);
}

struct BPC {
  Body* body1;
  Body* body2;
  std::vector<Contact*> contactList;
  std::vector<double> motionMetricHist;
  std::vector<ContactState> contactStateHist;
  bool inCollection = false;
  bool checked = false;
  int layer = 0;  // position class in getOrganizedContacts (test bookkeeping: 0 = holds a new contact, k = k-th BFS layer, ...)
  BPC(Body* a, Body* b) : body1(a), body2(b) {}
  Body* getBody(int i) const { return i == 0 ? body1 : body2; }
  Body* getOtherBody(const Body* b) const { return body1 == b ? body2 : (body2 == b ? body1 : nullptr); }
  int lo() const { return body1->id < body2->id ? body1->id : body2->id; }
  int hi() const { return body1->id < body2->id ? body2->id : body1->id; }
};
inline bool BpcLess::operator()(const BPC* a, const BPC* b) const {
  if (a->lo() != b->lo()) return a->lo() < b->lo();
  if (a->hi() != b->hi()) return a->hi() < b->hi();
  return a < b;  // never reached for distinct live pairs (one BPC per unordered pair)
}

struct Spring {
  int type;
  Body* body1 = nullptr;
  Body* body2 = nullptr;
  V3 pb1, pb2, pw;
  double k = 100, d = 10, l0 = 0.5, ls = 1;
};

struct Event {
  int step;
  int kind;  // 0 = merge (bpc becomes internal), 1 = unmerge (bpc cut and its bodies separated)
  int lo, hi;
};

}  // namespace amo
