// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// Single-threaded CPU restatement of mergingBodies3D.RigidBodySystem.advanceTime and everything it
// calls, following (in order of authority) RigidBodySystem.java:102-185, CollisionProcessor.java,
// PGS.java, Contact.java:159-398,543-572, collision/*.java, BodyPairContact.java, Merging.java,
// MotionMetricProcessor.java:39-73, RigidBody.java:276-462, RigidCollection.java:55-178,459-998,
// Sleeping.java:49-139, Spring.java:153-209.  Every function cites the lines it restates.
//
// PARITY UNPINNED for floating-point state: the reference has no golden vectors / known-answer tests
// for the 3D path and no JVM exists in this environment.  The only reference output
// that exists -- the authors' recorded step logs of tower25platform.xml (#bodies, #contacts per step)
// -- is reproduced exactly for the first 55-107 steps of all four recordings
// (tests/test_reference_logs.py); beyond that the restatement is validated by self-consistency
// (tests/test_oracle.py).  Built with -O2 -ffp-contract=off.
//
// Deliberate, documented deviations (all are places where the reference itself is unspecified):
//  * HashSet iteration orders are canonicalised (see oracle_model.h);
//  * options marked unsupported in include/am3d.h (shuffle, post-stabilisation, cycle merge
//    condition, position-level metric, collection BVH, Coriolis) are not restated;
//  * an external Gauss-Seidel order can be supplied for either solve so that the CUDA path's
//    colour order can be replayed (north_star parity mode 2); when that order marks hub sides
//    (bodies the CUDA path updates Jacobi-style across the groups of one colour) the
//    replay follows the same scheme (pgsSolveHub) -- without such marks the solve is the reference's
//    plain sequential Gauss-Seidel.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "oracle_collide.h"
#include "oracle_model.h"

namespace amo {

static inline double nowSec() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct System {
  am3d_params P;
  // scene-constant data
  std::vector<V3> nodeC;
  std::vector<double> nodeR;
  std::vector<int> nodeFirst, nodeCount, nodeRank;
  std::vector<std::unique_ptr<Body>> leaf;    // index = body id
  std::vector<std::unique_ptr<Body>> partArena;
  std::vector<std::unique_ptr<Body>> collArena;
  std::vector<int> bodyScene;
  std::vector<Spring> springs;
  // RigidBodySystem.bodies
  std::vector<Body*> bodies;
  // CollisionProcessor state
  std::vector<Contact*> contacts;
  std::vector<std::unique_ptr<Contact>> poolCur, poolPrev, heapContacts;
  std::vector<std::unique_ptr<BPC>> bpcArena;
  std::set<BPC*, BpcLess> bodyPairContacts;
  std::map<ContactKey, Contact*> lastTimeStepContacts;
  int totalSteps = 0;
  bool mergingEvent = false;
  int nextCollectionSlot = 0;
  std::vector<Event> events;
  am3d_timings T;
  int badWarmStarts = 0, badWarmStartsRepaired = 0;
  // externally supplied Gauss-Seidel orders (keys), consumed by the next step
  std::vector<am3d_contact> orderFull, orderSweep, orderPost;
  bool haveOrderFull = false, haveOrderSweep = false, haveOrderPost = false;
  bool postStabSolve = false;  // PGS.postStabilization of the solve in progress (PGS.java:86-89)
  int orderMismatch = 0;
  int lastIterations = 0;
  long rowUpdates = 0;  // 3 * contacts * iterations of full solves, accumulated
  double solveSeconds = 0;
  std::vector<Contact*> lastSweepList;  // contacts of the last single sweep, in order
  bool saveInitial = true;

  // ------------------------------------------------------------------------------------------
  Contact* poolGet() {
    poolCur.emplace_back(new Contact());
    return poolCur.back().get();
  }

  // ==========================================================================================
  // Contact.java
  // ==========================================================================================
  // Contact.computeJacobian(boolean, Point3d, Vector3d x3) :235-255
  static void computeJacobianW(Contact* c, bool computeInCollection, const V3& contactW, const V3& normalW,
                               const V3& tangent1W, const V3& tangent2W) {
    Body* b1 = (c->body1->isInCollection() && !computeInCollection) ? c->body1->parent : c->body1;
    Body* b2 = (c->body2->isInCollection() && !computeInCollection) ? c->body2->parent : c->body2;
    V3 r1 = sub(contactW, b1->x);
    V3 r2 = sub(contactW, b2->x);
    c->jna.v = scale(-1, normalW);
    c->jna.w = cross(normalW, r1);
    c->jnb.v = normalW;
    c->jnb.w = cross(r2, normalW);
    c->jt1a.v = scale(-1, tangent1W);
    c->jt1a.w = cross(tangent1W, r1);
    c->jt1b.v = tangent1W;
    c->jt1b.w = cross(r2, tangent1W);
    c->jt2a.v = scale(-1, tangent2W);
    c->jt2a.w = cross(tangent2W, r1);
    c->jt2b.v = tangent2W;
    c->jt2b.w = cross(r2, tangent2W);
  }

  // Contact.set :159-233
  static void contactSet(Contact* c, Body* b1, Body* b2, const V3& contactW, const V3& normalW, int bv1, int bv2,
                         int info, double violation) {
    c->state = CLEAR;
    c->body1 = b1;
    c->body2 = b2;
    c->info = info;
    if (b1->isInComposite()) { c->body1 = b1->compositeBodyParent; c->csb1 = b1; } else c->csb1 = nullptr;
    if (b2->isInComposite()) { c->body2 = b2->compositeBodyParent; c->csb2 = b2; } else c->csb2 = nullptr;
    c->bv1 = bv1;
    c->bv2 = bv2;
    c->constraintViolation = violation;
    c->prevConstraintViolation = 0.;
    c->lambda0 = c->lambda1 = c->lambda2 = 0;
    c->lambda0warm = c->lambda1warm = c->lambda2warm = 0;
    c->contactB1 = contactW;
    c->normalB1 = normalW;
    c->pointW = contactW;
    c->normalW = normalW;
    double anx = std::fabs(normalW.x), any = std::fabs(normalW.y), anz = std::fabs(normalW.z);
    if (anx < any && anx < anz) c->tangent1B1 = V3(1, 0, 0);
    else if (any < anz) c->tangent1B1 = V3(0, 1, 0);
    else c->tangent1B1 = V3(0, 0, 1);
    c->tangent2B1 = cross(normalW, c->tangent1B1);
    c->tangent2B1 = normalize(c->tangent2B1);
    c->tangent1B1 = cross(c->tangent2B1, normalW);
    computeJacobianW(c, false, c->contactB1, c->normalB1, c->tangent1B1, c->tangent2B1);
    Xf T = c->body1->B2W();
    c->contactB1 = T.inverseTransformP(c->contactB1);
    c->normalB1 = T.inverseTransformV(c->normalB1);
    c->tangent1B1 = T.inverseTransformV(c->tangent1B1);
    c->tangent2B1 = T.inverseTransformV(c->tangent2B1);
    c->bn = c->bt1 = c->bt2 = 0;
    c->D00 = c->D11 = c->D22 = 0;
    c->newThisTimeStep = true;
  }

  // Contact.computeJacobian(boolean) :265-271
  static void computeJacobian(Contact* c, bool computeInCollection) {
    Xf T = c->body1->B2W();
    V3 pW = T.transformP(c->contactB1);
    V3 nW = T.transformV(c->normalB1);
    V3 t1W = T.transformV(c->tangent1B1);
    V3 t2W = T.transformV(c->tangent2B1);
    computeJacobianW(c, computeInCollection, pW, nW, t1W, t2W);
  }

  // Contact.computeB :279-326
  static void computeB(Contact* c, double dt, double feedbackStiffness, bool computeInCollection,
                       bool restitutionOverride, double restitutionOverrideVal) {
    Body* b1 = (c->body1->isInCollection() && !computeInCollection) ? c->body1->parent : c->body1;
    Body* b2 = (c->body2->isInCollection() && !computeInCollection) ? c->body2->parent : c->body2;
    double restitution = (c->body1->restitution + c->body2->restitution) / 2.;
    if (restitutionOverride) restitution = restitutionOverrideVal;
    c->bn = 0; c->bt1 = 0; c->bt2 = 0;
    V3 tmp1 = scaleAdd(b1->minv * dt, b1->force, b1->v);
    c->bn += dot(tmp1, c->jna.v);
    c->bt1 += dot(tmp1, c->jt1a.v);
    c->bt2 += dot(tmp1, c->jt2a.v);
    tmp1 = transform(b1->jinv, b1->torque);
    tmp1 = scale(dt, tmp1);
    tmp1 = add(tmp1, b1->omega);
    c->bn += dot(tmp1, c->jna.w);
    c->bt1 += dot(tmp1, c->jt1a.w);
    c->bt2 += dot(tmp1, c->jt2a.w);
    double bBounce = dot(b1->v, c->jna.v) + dot(b1->omega, c->jna.w);
    bBounce *= restitution;
    c->bn += bBounce;
    tmp1 = scaleAdd(b2->minv * dt, b2->force, b2->v);
    c->bn += dot(tmp1, c->jnb.v);
    c->bt1 += dot(tmp1, c->jt1b.v);
    c->bt2 += dot(tmp1, c->jt2b.v);
    tmp1 = transform(b2->jinv, b2->torque);
    tmp1 = scale(dt, tmp1);
    tmp1 = add(tmp1, b2->omega);
    c->bn += dot(tmp1, c->jnb.w);
    c->bt1 += dot(tmp1, c->jt1b.w);
    c->bt2 += dot(tmp1, c->jt2b.w);
    bBounce = dot(b2->v, c->jnb.v) + dot(b2->omega, c->jnb.w);
    bBounce *= restitution;
    c->bn += bBounce;
    double baumgarteFeedback = feedbackStiffness * c->constraintViolation;
    c->bn += baumgarteFeedback;
  }

  // Contact.computeJMinvJt :340-354
  static void computeJMinvJt(Contact* c, bool computeInCollection) {
    Body* b1 = (c->body1->isInCollection() && !computeInCollection) ? c->body1->parent : c->body1;
    Body* b2 = (c->body2->isInCollection() && !computeInCollection) ? c->body2->parent : c->body2;
    V3 tmp1 = transform(b1->jinv, c->jna.w), tmp2 = transform(b2->jinv, c->jnb.w);
    c->D00 = b1->minv * dot(c->jna.v, c->jna.v) + dot(c->jna.w, tmp1) + b2->minv * dot(c->jnb.v, c->jnb.v) + dot(c->jnb.w, tmp2);
    tmp1 = transform(b1->jinv, c->jt1a.w); tmp2 = transform(b2->jinv, c->jt1b.w);
    c->D11 = b1->minv * dot(c->jt1a.v, c->jt1a.v) + dot(c->jt1a.w, tmp1) + b2->minv * dot(c->jt1b.v, c->jt1b.v) + dot(c->jt1b.w, tmp2);
    tmp1 = transform(b1->jinv, c->jt2a.w); tmp2 = transform(b2->jinv, c->jt2b.w);
    c->D22 = b1->minv * dot(c->jt2a.v, c->jt2a.v) + dot(c->jt2a.w, tmp1) + b2->minv * dot(c->jt2b.v, c->jt2b.v) + dot(c->jt2b.w, tmp2);
  }

  // Contact.getJdv :362-378
  static double getJdv(const Contact* c, bool computeInCollection, int index) {
    const V6& dv1 = (c->body1->isInCollection() && !computeInCollection) ? c->body1->parent->deltaV : c->body1->deltaV;
    const V6& dv2 = (c->body2->isInCollection() && !computeInCollection) ? c->body2->parent->deltaV : c->body2->deltaV;
    const V6* ja = &c->jna; const V6* jb = &c->jnb;
    if (index == 1) { ja = &c->jt1a; jb = &c->jt1b; } else if (index == 2) { ja = &c->jt2a; jb = &c->jt2b; }
    return dot6(*ja, dv1) + dot6(*jb, dv2);
  }

  // Contact.updateContactState :385-398
  void updateContactState(Contact* c, bool computeInCollection) {
    c->w1 = c->bt1 + getJdv(c, computeInCollection, 1);
    c->w2 = c->bt2 + getJdv(c, computeInCollection, 2);
    if (std::fabs(c->lambda0) <= 1e-14) c->state = BROKEN;
    else if (std::fabs(c->w1) > P.sliding_threshold) c->state = ONEDGE;
    else if (std::fabs(c->w2) > P.sliding_threshold) c->state = ONEDGE;
    else c->state = CLEAR;
  }

  // ==========================================================================================
  // PGS.java
  // ==========================================================================================
  // PGS.updateDeltaVwithLambdai :221-245
  static void updateDeltaVwithLambdai(Contact* c, double lambda, int i, bool computeInCollection) {
    Body* body1 = (c->body1->isInCollection() && !computeInCollection) ? c->body1->parent : c->body1;
    Body* body2 = (c->body2->isInCollection() && !computeInCollection) ? c->body2->parent : c->body2;
    V6& dv1 = body1->deltaV;
    V6& dv2 = body2->deltaV;
    const V6* ja = &c->jna; const V6* jb = &c->jnb;
    if (i == 1) { ja = &c->jt1a; jb = &c->jt1b; } else if (i == 2) { ja = &c->jt2a; jb = &c->jt2b; }
    dv1.v = scaleAdd(body1->minv * lambda, ja->v, dv1.v);
    V3 tmp = transform(body1->jinv, ja->w);
    dv1.w = scaleAdd(lambda, tmp, dv1.w);
    dv2.v = scaleAdd(body2->minv * lambda, jb->v, dv2.v);
    tmp = transform(body2->jinv, jb->w);
    dv2.w = scaleAdd(lambda, tmp, dv2.w);
  }

  // PGS.solve :73-194
  void pgsSolve(std::vector<Contact*>& list, double dt, int iterations, double tolerance, double omega,
                double feedbackStiffness, double compliance, bool computeInCollection) {
    if (list.empty()) return;
    // confidentWarmStart :250-256
    for (Contact* c : list) {
      updateDeltaVwithLambdai(c, c->lambda0, 0, computeInCollection);
      updateDeltaVwithLambdai(c, c->lambda1, 1, computeInCollection);
      updateDeltaVwithLambdai(c, c->lambda2, 2, computeInCollection);
    }
    for (Contact* c : list) {
      if (postStabSolve) { c->bn = feedbackStiffness * c->constraintViolation; c->bt1 = 0.; c->bt2 = 0.; }  // PGS.java:86-89
      else computeB(c, dt, feedbackStiffness, computeInCollection, P.restitution_override != 0, P.restitution);
      computeJMinvJt(c, computeInCollection);
    }
    int iter = iterations;
    int executed = 0;
    while (iter > 0) {
      double lambdaChangeAbsMax = 0;
      for (size_t i = 0; i < list.size(); i++) {
        Contact* contact = list[i];
        double Jdvn = getJdv(contact, computeInCollection, 0);
        double prevLambda_n = contact->lambda0;
        contact->lambda0 = (contact->D00 * contact->lambda0 - omega * (contact->bn + Jdvn)) / (contact->D00 + compliance);
        bool clamp = (!contact->body1->magnetic || !contact->body1->activateMagnet) &&
                     (!contact->body2->magnetic || !contact->body2->activateMagnet);
        if (clamp) contact->lambda0 = std::max(0.0, contact->lambda0);
        double diff = contact->lambda0 - prevLambda_n;
        updateDeltaVwithLambdai(contact, diff, 0, computeInCollection);
        lambdaChangeAbsMax = std::max(lambdaChangeAbsMax, std::fabs(diff));
        double mu = 0.;
        if (P.friction_override) {
          mu = P.friction;
        } else {
          double f1 = contact->body1->friction, f2 = contact->body2->friction;
          if (f1 < 0.2 || f2 < 0.2) mu = std::min(f1, f2);
          else if (f1 > 1. || f2 > 1.) mu = std::max(f1, f2);
          else mu = (f1 + f2) / 2.;
        }
        double Jdvt1 = getJdv(contact, computeInCollection, 1);
        double prevLambda_t1 = contact->lambda1;
        contact->lambda1 = (contact->D11 * contact->lambda1 - omega * (contact->bt1 + Jdvt1)) / (contact->D11 + compliance);
        if (clamp) {
          double limit = mu * contact->lambda0;
          contact->lambda1 = std::max(contact->lambda1, -limit);
          contact->lambda1 = std::min(contact->lambda1, limit);
        }
        diff = contact->lambda1 - prevLambda_t1;
        updateDeltaVwithLambdai(contact, diff, 1, computeInCollection);
        lambdaChangeAbsMax = std::max(lambdaChangeAbsMax, std::fabs(diff));
        double Jdvt2 = getJdv(contact, computeInCollection, 2);
        double prevLambda_t2 = contact->lambda2;
        contact->lambda2 = (contact->D22 * contact->lambda2 - omega * (contact->bt2 + Jdvt2)) / (contact->D22 + compliance);
        if (clamp) {
          double limit = mu * contact->lambda0;
          contact->lambda2 = std::max(contact->lambda2, -limit);
          contact->lambda2 = std::min(contact->lambda2, limit);
        }
        diff = contact->lambda2 - prevLambda_t2;
        updateDeltaVwithLambdai(contact, diff, 2, computeInCollection);
        lambdaChangeAbsMax = std::max(lambdaChangeAbsMax, std::fabs(diff));
        if (iter == 1) updateContactState(contact, computeInCollection);
      }
      iter--;
      executed++;
      if (!computeInCollection && lambdaChangeAbsMax < tolerance) break;
    }
    if (!computeInCollection) lastIterations = executed;
  }


  // ------------------------------------------------------------------------------------------------
  // Replay of the CUDA path's solve when it used HUB bodies: not part of the reference.
  // The supplied sequence is colour-major; consecutive contacts of one body pair form a group.  A group works on a
  // private copy of a hub body's deltaV (its value at the start of the colour plus the group's own updates) and
  // records what it added; after the colour the recorded deltas are folded into the hub in the GPU's fixed order
  // (32 strided partial sums, then an xor butterfly).  Without hub flags this is exactly pgsSolve.
  // ------------------------------------------------------------------------------------------------
  struct OrderMeta { int color; int hub; };
  static void applyRowTo(V6& dv, const Body* body, const V6& j, double lambda) {
    dv.v = scaleAdd(body->minv * lambda, j.v, dv.v);
    V3 tmp = transform(body->jinv, j.w);
    dv.w = scaleAdd(lambda, tmp, dv.w);
  }
  static V6 warpSum(const std::vector<V6>& vals) {
    double part[32][6];
    for (int l = 0; l < 32; l++) for (int k = 0; k < 6; k++) part[l][k] = 0.0;
    for (size_t e = 0; e < vals.size(); e++) {
      const V6& d = vals[e];
      double x[6] = {d.v.x, d.v.y, d.v.z, d.w.x, d.w.y, d.w.z};
      for (int k = 0; k < 6; k++) part[e % 32][k] = part[e % 32][k] + x[k];
    }
    for (int o = 16; o > 0; o >>= 1) {
      double nw[32][6];
      for (int l = 0; l < 32; l++) for (int k = 0; k < 6; k++) nw[l][k] = part[l][k] + part[l ^ o][k];
      std::memcpy(part, nw, sizeof(part));
    }
    V6 r;
    r.v = V3(part[0][0], part[0][1], part[0][2]);
    r.w = V3(part[0][3], part[0][4], part[0][5]);
    return r;
  }
  template <class Fn>
  void hubSweep(std::vector<Contact*>& list, const std::vector<OrderMeta>& meta, bool cic, Fn fn) {
    size_t i = 0, n = list.size();
    auto sb = [&](Contact* c, int side) {
      Body* b = side == 0 ? c->body1 : c->body2;
      return (b->isInCollection() && !cic) ? b->parent : b;
    };
    while (i < n) {
      int col = meta[i].color;
      size_t cend = i;
      while (cend < n && meta[cend].color == col) cend++;
      std::vector<std::pair<Body*, std::vector<V6>>> pending;
      auto pend = [&](Body* h, const V6& d) {
        for (auto& kv : pending) if (kv.first == h) { kv.second.push_back(d); return; }
        pending.push_back({h, std::vector<V6>{d}});
      };
      size_t g = i;
      while (g < cend) {
        size_t ge = g;
        while (ge < cend && list[ge]->body1 == list[g]->body1 && list[ge]->body2 == list[g]->body2) ge++;
        Body *s1 = sb(list[g], 0), *s2 = sb(list[g], 1);
        bool h1 = (meta[g].hub & 1) != 0, h2 = (meta[g].hub & 2) != 0;
        V6 loc1 = s1->deltaV, loc2 = s2->deltaV, acc1, acc2;
        for (size_t k = g; k < ge; k++)
          fn(list[k], s1, s2, h1 ? &loc1 : &s1->deltaV, h2 ? &loc2 : &s2->deltaV, h1 ? &acc1 : nullptr, h2 ? &acc2 : nullptr);
        if (h1) pend(s1, acc1);
        if (h2) pend(s2, acc2);
        g = ge;
      }
      for (auto& kv : pending) {
        V6 t = warpSum(kv.second);
        V6& d = kv.first->deltaV;
        d.v = V3(d.v.x + t.v.x, d.v.y + t.v.y, d.v.z + t.v.z);
        d.w = V3(d.w.x + t.w.x, d.w.y + t.w.y, d.w.z + t.w.z);
      }
      i = cend;
    }
  }
  void pgsSolveHub(std::vector<Contact*>& list, const std::vector<OrderMeta>& meta, double dt, int iterations, double tolerance,
                   double omega, double feedbackStiffness, double compliance, bool computeInCollection) {
    if (list.empty()) return;
    const bool cic = computeInCollection;
    auto rows = [](Contact* c, int k, const V6*& ja, const V6*& jb) {
      ja = &c->jna; jb = &c->jnb;
      if (k == 1) { ja = &c->jt1a; jb = &c->jt1b; } else if (k == 2) { ja = &c->jt2a; jb = &c->jt2b; }
    };
    hubSweep(list, meta, cic, [&](Contact* c, Body* s1, Body* s2, V6* d1, V6* d2, V6* a1, V6* a2) {
      double lam[3] = {c->lambda0, c->lambda1, c->lambda2};
      for (int k = 0; k < 3; k++) {
        const V6 *ja, *jb;
        rows(c, k, ja, jb);
        applyRowTo(*d1, s1, *ja, lam[k]);
        applyRowTo(*d2, s2, *jb, lam[k]);
        if (a1) applyRowTo(*a1, s1, *ja, lam[k]);
        if (a2) applyRowTo(*a2, s2, *jb, lam[k]);
      }
    });
    for (Contact* c : list) {
      if (postStabSolve) { c->bn = feedbackStiffness * c->constraintViolation; c->bt1 = 0.; c->bt2 = 0.; }
      else computeB(c, dt, feedbackStiffness, cic, P.restitution_override != 0, P.restitution);
      computeJMinvJt(c, cic);
    }
    int iter = iterations, executed = 0;
    while (iter > 0) {
      double lambdaChangeAbsMax = 0;
      hubSweep(list, meta, cic, [&](Contact* contact, Body* s1, Body* s2, V6* d1, V6* d2, V6* a1, V6* a2) {
        bool clamp = (!contact->body1->magnetic || !contact->body1->activateMagnet) &&
                     (!contact->body2->magnetic || !contact->body2->activateMagnet);
        double mu;
        if (P.friction_override) mu = P.friction;
        else {
          double f1 = contact->body1->friction, f2 = contact->body2->friction;
          if (f1 < 0.2 || f2 < 0.2) mu = std::min(f1, f2);
          else if (f1 > 1. || f2 > 1.) mu = std::max(f1, f2);
          else mu = (f1 + f2) / 2.;
        }
        double* lam[3] = {&contact->lambda0, &contact->lambda1, &contact->lambda2};
        double bb[3] = {contact->bn, contact->bt1, contact->bt2}, DD[3] = {contact->D00, contact->D11, contact->D22};
        for (int k = 0; k < 3; k++) {
          const V6 *ja, *jb;
          rows(contact, k, ja, jb);
          double Jdv = dot6(*ja, *d1) + dot6(*jb, *d2);
          double prev = *lam[k];
          double l = (DD[k] * prev - omega * (bb[k] + Jdv)) / (DD[k] + compliance);
          if (clamp) {
            if (k == 0) l = std::max(0.0, l);
            else { double limit = mu * contact->lambda0; l = std::max(l, -limit); l = std::min(l, limit); }
          }
          *lam[k] = l;
          double diff = l - prev;
          applyRowTo(*d1, s1, *ja, diff);
          applyRowTo(*d2, s2, *jb, diff);
          if (a1) applyRowTo(*a1, s1, *ja, diff);
          if (a2) applyRowTo(*a2, s2, *jb, diff);
          lambdaChangeAbsMax = std::max(lambdaChangeAbsMax, std::fabs(diff));
        }
        if (iter == 1) {
          contact->w1 = contact->bt1 + (dot6(contact->jt1a, *d1) + dot6(contact->jt1b, *d2));
          contact->w2 = contact->bt2 + (dot6(contact->jt2a, *d1) + dot6(contact->jt2b, *d2));
          if (std::fabs(contact->lambda0) <= 1e-14) contact->state = BROKEN;
          else if (std::fabs(contact->w1) > P.sliding_threshold) contact->state = ONEDGE;
          else if (std::fabs(contact->w2) > P.sliding_threshold) contact->state = ONEDGE;
          else contact->state = CLEAR;
        }
      });
      iter--;
      executed++;
      if (!cic && lambdaChangeAbsMax < tolerance) break;
    }
    if (!cic) lastIterations = executed;
  }

  // ==========================================================================================
  // collision detection: CollisionProcessor.java
  // ==========================================================================================
  V3 nodeCW(const Body* body, int node) const { return body->B2W().transformP(nodeC[node]); }
  bool isLeafNode(int node) const { return nodeFirst[node] < 0; }

  void emitContact(Body* b1, Body* b2, const V3& p, const V3& n, int bv1, int bv2, int info, double viol, int leafNode = -1) {
    Contact* c = poolGet();
    contactSet(c, b1, b2, p, n, bv1, bv2, info, viol);
    c->leaf = leafNode;
    contacts.push_back(c);
  }

  // collideSphereTreeAndPlane :793-826
  void collideSphereTreeAndPlane(int node1, Body* body1, Body* planeBody) {
    V3 c = nodeCW(body1, node1);
    const V3& n = planeBody->n;
    double d = n.x * c.x + n.y * c.y + n.z * c.z + planeBody->d - nodeR[node1];
    if (d < 0) {
      if (isLeafNode(node1)) {
        V3 normal = sub(c, planeBody->p);
        double val = dot(normal, planeBody->n);
        normal = scale(val, planeBody->n);
        V3 contactW = sub(c, normal);
        normal = scale(-1, planeBody->n);
        emitContact(body1, planeBody, contactW, normal, node1, AM3D_BV_PLANE_DUMMY, 0, d);
      } else {
        for (int k = 0; k < nodeCount[node1]; k++) collideSphereTreeAndPlane(nodeFirst[node1] + k, body1, planeBody);
      }
    }
  }

  // collideBoxAndSphereTree :837-850
  void collideBoxAndSphereTree(Body* body1 /*box*/, int node2, Body* body2) {
    V3 cW = nodeCW(body2, node2);
    double r = nodeR[node2];
    Xf T = body1->B2W();
    if (!dBoxSphereTest(T, body1->size, body1->radius, cW, r)) return;
    if (!isLeafNode(node2)) {
      for (int k = 0; k < nodeCount[node2]; k++) collideBoxAndSphereTree(body1, nodeFirst[node2] + k, body2);
    } else {
      dBoxSphere(T, body1->size, cW, r, [&](const Hit& h) {
        emitContact(body1, body2, h.pos, h.normal, AM3D_BV_NULL, AM3D_BV_NULL, 0, h.violation, node2);
      });
    }
  }

  // processCollision :1036-1056
  void processCollision(Body* body1, int bv1, Body* body2, int bv2, const V3& c1, const V3& c2) {
    double dist = distance(c1, c2);
    double distanceBetweenCenters = nodeR[bv2] + nodeR[bv1];
    if (dist < distanceBetweenCenters) {
      double alpha = (nodeR[bv1] - nodeR[bv2] + dist) / (2 * dist);
      V3 contactW = interpolate(c1, c2, alpha);
      V3 normal = sub(c2, c1);
      normal = normalize(normal);
      emitContact(body1, body2, contactW, normal, bv1, bv2, 0, dist - distanceBetweenCenters);
    }
  }

  // collideSphereTrees :975-1009
  void collideSphereTrees(int node1, int node2, Body* body1, Body* body2) {
    V3 c1 = nodeCW(body1, node1), c2 = nodeCW(body2, node2);
    double r1 = nodeR[node1], r2 = nodeR[node2];
    if (distanceSquared(c1, c2) < (r1 + r2) * (r1 + r2)) {
      bool l1 = isLeafNode(node1), l2 = isLeafNode(node2);
      if (l1 && l2) {
        processCollision(body1, node1, body2, node2, c1, c2);
      } else if (l1) {
        for (int k = 0; k < nodeCount[node2]; k++) collideSphereTrees(node1, nodeFirst[node2] + k, body1, body2);
      } else if (l2) {
        for (int k = 0; k < nodeCount[node1]; k++) collideSphereTrees(nodeFirst[node1] + k, node2, body1, body2);
      } else if (r1 <= r2) {
        for (int k = 0; k < nodeCount[node2]; k++) collideSphereTrees(node1, nodeFirst[node2] + k, body1, body2);
      } else {
        for (int k = 0; k < nodeCount[node1]; k++) collideSphereTrees(nodeFirst[node1] + k, node2, body1, body2);
      }
    }
  }

  // RigidBodyGeomComposite.updateBodyPositionsFromParent :32-36
  static void updateBodyPositionsFromParent(Body* comp) {
    for (Body* b : comp->parts) b->setB2W(Xf::mult(b->compositeBodyParent->B2W(), b->B2C));
  }

  // narrowPhase :707-757 (+ narrowPhaseCollection :768-776, brute-force mode only)
  void narrowPhase(Body* body1, Body* body2) {
    if (body1->isCollection || body2->isCollection) {
      if (body1->isCollection) {
        for (Body* b : body1->bodies) narrowPhase(b, body2);
      } else {
        for (Body* b : body2->bodies) narrowPhase(body1, b);
      }
    } else if (body1->geom == G_COMPOSITE) {
      updateBodyPositionsFromParent(body1);
      for (Body* b : body1->parts) narrowPhase(b, body2);
    } else if (body2->geom == G_COMPOSITE) {
      updateBodyPositionsFromParent(body2);
      for (Body* b : body2->parts) narrowPhase(body1, b);
    } else if (body1->isPlane()) {
      if (body2->isPlane()) {
        // "plane plane collision is impossible!"
      } else if (body2->geom == G_BOX) {
        dBoxPlane(body2->B2W(), body2->size, body2->radius, body1->n, body1->d, [&](const Hit& h) {
          emitContact(body1, body2, h.pos, h.normal, AM3D_BV_NULL, AM3D_BV_NULL, h.info, h.violation);
        });
      } else {
        collideSphereTreeAndPlane(body2->root, body2, body1);
      }
    } else if (body1->geom == G_BOX) {
      if (body2->isPlane()) {
        dBoxPlane(body1->B2W(), body1->size, body1->radius, body2->n, body2->d, [&](const Hit& h) {
          emitContact(body2, body1, h.pos, h.normal, AM3D_BV_NULL, AM3D_BV_NULL, h.info, h.violation);
        });
      } else if (body2->geom == G_BOX) {
        dBoxBox(body1->x, body1->theta, body1->size, body1->radius, body2->x, body2->theta, body2->size, body2->radius,
                [&](const Hit& h) {
                  emitContact(body1, body2, h.pos, h.normal, AM3D_BV_NULL, AM3D_BV_NULL, h.info, h.violation);
                });
      } else {
        collideBoxAndSphereTree(body1, body2->root, body2);
      }
    } else {
      if (body2->isPlane()) collideSphereTreeAndPlane(body1->root, body1, body2);
      else if (body2->geom == G_BOX) collideBoxAndSphereTree(body2, body1->root, body1);
      else collideSphereTrees(body1->root, body2->root, body1, body2);
    }
  }

  int sceneOf(const Body* b) const {
    const Body* l = b->isCollection ? b->bodies[0] : b;
    return bodyScene[l->id];
  }

  // broadPhase :683-697.  Batched scenes (am3d_scene.body_scene) never interact: the reference
  // would hold them in separate RigidBodySystem objects.
  void broadPhase() {
    int N = (int)bodies.size();
    bool multi = nScenes > 1;
    for (int i = 0; i < N - 1; i++) {
      Body* b1 = bodies[i];
      for (int j = i + 1; j < N; j++) {
        Body* b2 = bodies[j];
        if (b1->pinned && b2->pinned) continue;
        if ((b1->pinned && b2->sleeping) || (b2->pinned && b1->sleeping)) continue;
        if (multi && sceneOf(b1) != sceneOf(b2)) continue;
        narrowPhase(b1, b2);
      }
    }
  }
  int nScenes = 1;

  // collisionDetection :91-102
  void collisionDetection() {
    contacts.clear();
    poolPrev.clear();
    std::swap(poolCur, poolPrev);
    double t0 = nowSec();
    broadPhase();
    T.detection = nowSec() - t0;
    if (contacts.empty()) lastTimeStepContacts.clear();
  }

  // updateContactsMap :443-448
  void updateContactsMap() {
    lastTimeStepContacts.clear();
    for (Contact* c : contacts) lastTimeStepContacts[keyOf(c)] = c;  // duplicate keys: last put wins
  }

  // ==========================================================================================
  // BodyPairContact bookkeeping: CollisionProcessor.java:145-226, BodyPairContact.java:54-62,315-355
  // ==========================================================================================
  BPC* findExternal(Body* b1, Body* b2) {
    int lo = std::min(b1->id, b2->id), hi = std::max(b1->id, b2->id);
    BPC probe(b1, b2);
    auto it = bodyPairContacts.lower_bound(&probe);
    // step back over equal-(lo,hi) entries with lower addresses
    while (it != bodyPairContacts.begin()) {
      auto pv = std::prev(it);
      if ((*pv)->lo() == lo && (*pv)->hi() == hi) it = pv; else break;
    }
    if (it != bodyPairContacts.end() && (*it)->lo() == lo && (*it)->hi() == hi) return *it;
    return nullptr;
  }

  static void addToBodyLists(BPC* bpc) {
    bpc->body1->bodyPairContacts.insert(bpc);
    bpc->body2->bodyPairContacts.insert(bpc);
  }
  static void addToBodyListsParent(BPC* bpc) {
    if (bpc->body1->isInCollection()) bpc->body1->parent->bodyPairContacts.insert(bpc);
    if (bpc->body2->isInCollection()) bpc->body2->parent->bodyPairContacts.insert(bpc);
  }
  static void removeFromBodyLists(BPC* bpc) {
    bpc->body1->bodyPairContacts.erase(bpc);
    bpc->body2->bodyPairContacts.erase(bpc);
  }
  static void removeFromBodyListsParent(BPC* bpc) {
    if (bpc->body1->isInCollection()) bpc->body1->parent->bodyPairContacts.erase(bpc);
    if (bpc->body2->isInCollection()) bpc->body2->parent->bodyPairContacts.erase(bpc);
  }

  // storeInBodyPairContacts :170-186
  void storeInBodyPairContacts(Contact* contact) {
    if (contact->body1->pinned && contact->body2->pinned) return;
    BPC* bpc = findExternal(contact->body1, contact->body2);
    if (bpc == nullptr) {
      bpcArena.emplace_back(new BPC(contact->body1, contact->body2));
      bpc = bpcArena.back().get();
      bodyPairContacts.insert(bpc);
    }
    addToBodyLists(bpc);
    addToBodyListsParent(bpc);
    bpc->contactList.push_back(contact);
  }

  // removeEmptyBodyPairContacts :191-208
  void removeEmptyBodyPairContacts() {
    for (auto it = bodyPairContacts.begin(); it != bodyPairContacts.end();) {
      BPC* bpc = *it;
      if (bpc->contactList.empty()) {
        removeFromBodyLists(bpc);
        removeFromBodyListsParent(bpc);
        it = bodyPairContacts.erase(it);
      } else {
        ++it;
      }
    }
  }

  // updateBodyPairContacts :145-164
  void updateBodyPairContacts() {
    for (BPC* bpc : bodyPairContacts) { bpc->checked = false; bpc->contactList.clear(); }
    for (Body* body : bodies)
      if (body->isCollection)
        for (BPC* bpc : body->bodyPairContacts) bpc->checked = false;
    for (Contact* c : contacts) storeInBodyPairContacts(c);
    removeEmptyBodyPairContacts();
  }

  // clearBodyPairContacts :213-226
  void clearBodyPairContacts() {
    std::vector<Contact*> tmp;
    for (BPC* bpc : bodyPairContacts) {
      tmp.clear();
      for (Contact* c : bpc->contactList)
        if (std::fabs(c->lambda0) > 1e-14) tmp.push_back(c);
      bpc->contactList = tmp;
    }
    removeEmptyBodyPairContacts();
  }

  // ==========================================================================================
  // warm start: CollisionProcessor.java:481-665
  // ==========================================================================================
  Contact* lookup(const Contact* c, int info = -1) {
    auto it = lastTimeStepContacts.find(keyOf(c, info));
    return it == lastTimeStepContacts.end() ? nullptr : it->second;
  }
  static void takeWarm(Contact* contact, Contact* oldContact, bool zeroDonor, bool postStabilization) {
    contact->newThisTimeStep = false;
    contact->lambda0 = oldContact->lambda0;
    contact->lambda1 = oldContact->lambda1;
    contact->lambda2 = oldContact->lambda2;
    contact->lambda0warm = oldContact->lambda0;
    contact->lambda1warm = oldContact->lambda1;
    contact->lambda2warm = oldContact->lambda2;
    if (zeroDonor) { oldContact->lambda0 = 0; oldContact->lambda1 = 0; oldContact->lambda2 = 0; }
    // CollisionProcessor.java:569-572, 617-620, 657-660
    contact->prevConstraintViolation = postStabilization ? oldContact->prevConstraintViolation : oldContact->constraintViolation;
  }
  // vanillaWarmStart :646-665
  void vanillaWarmStart(Contact* contact, bool postStabilization) {
    Contact* oldContact = lookup(contact);
    if (oldContact != nullptr) takeWarm(contact, oldContact, false, postStabilization);
    else contact->newThisTimeStep = true;
  }
  static bool isBoxGeom(const Body* b) { return b->geom == G_BOX; }

  void warmStart(bool postStabilization = false) {
    badWarmStarts = 0;
    badWarmStartsRepaired = 0;
    for (BPC* bpc : bodyPairContacts) {
      if (bpc->inCollection) continue;
      bool b1IsBox = bpc->body1->geom == G_BOX, b2IsBox = bpc->body2->geom == G_BOX;
      bool b1IsComp = bpc->body1->geom == G_COMPOSITE, b2IsComp = bpc->body2->geom == G_COMPOSITE;
      if ((b1IsBox && b2IsBox) || (b1IsComp && b2IsBox) || (b1IsBox && b2IsComp) || (b1IsComp && b2IsComp)) {
        for (Contact* contact : bpc->contactList) {
          if (contact->csb1 != nullptr && !isBoxGeom(contact->csb1)) { vanillaWarmStart(contact, postStabilization); continue; }
          if (contact->csb2 != nullptr && !isBoxGeom(contact->csb2)) { vanillaWarmStart(contact, postStabilization); }  // falls through (:501-503)
          V3 pNew = contact->body1->B2W().transformP(contact->contactB1);
          V3 pOld;
          Contact* oldContact = lookup(contact);
          if (oldContact != nullptr) {
            pOld = oldContact->body1->B2W().transformP(oldContact->contactB1);
            double dist = distance(pNew, pOld);
            if (dist > 0.05) {
              badWarmStarts++;
              int myInfo = contact->info;
              double bestMatchDist = dist;
              int bestMatchInfo = myInfo;
              Contact* bestMatchOldContact = oldContact;
              for (int info = 0; info < 9; info++) {
                if (info == myInfo) continue;
                oldContact = lookup(contact, info);
                if (oldContact == nullptr) break;
                pOld = oldContact->body1->B2W().transformP(oldContact->contactB1);
                dist = distance(pNew, pOld);
                if (dist < bestMatchDist) { bestMatchDist = dist; bestMatchInfo = info; bestMatchOldContact = oldContact; }
              }
              if (bestMatchInfo != myInfo) badWarmStartsRepaired++;
              oldContact = bestMatchOldContact;
              dist = bestMatchDist;
            }
            if (dist < 0.05) takeWarm(contact, oldContact, true, postStabilization);
            else contact->newThisTimeStep = true;
          } else {
            double bestMatchDist = 1.7976931348623157e308;
            int bestMatchInfo = -1;
            Contact* bestMatchOldContact = nullptr;
            for (int info = 0; info < 9; info++) {
              oldContact = lookup(contact, info);
              if (oldContact == nullptr) break;
              pOld = oldContact->body1->B2W().transformP(oldContact->contactB1);
              double dist = distance(pNew, pOld);
              if (dist < bestMatchDist) { bestMatchDist = dist; bestMatchInfo = info; bestMatchOldContact = oldContact; }
            }
            if (bestMatchInfo != -1) {
              badWarmStartsRepaired++;
              if (bestMatchDist < 0.05) takeWarm(contact, bestMatchOldContact, true, postStabilization);
              // else: newThisTimeStep keeps the value Contact.set gave it (true)
            } else {
              contact->newThisTimeStep = true;
            }
          }
        }
      } else {
        for (Contact* contact : bpc->contactList) vanillaWarmStart(contact, postStabilization);
      }
    }
  }

  // ==========================================================================================
  // RigidBody.java / RigidCollection.java
  // ==========================================================================================
  // RigidBody.updateRotationalInertiaFromTransformation :311-321
  static void updateRotationalInertiaFromTransformation(Body* b) {
    if (!b->pinned) {
      Xf T = b->B2W();
      b->massAngular = T.computeRM0RT(b->massAngular0);
      b->jinv = T.computeRM0RT(b->jinv0);
    }
  }
  // RigidBody.applyForceW :335-340
  static void applyForceW(Body* b, const V3& pW, const V3& fW) {
    b->force = add(b->force, fW);
    V3 tmp = sub(pW, b->x);
    V3 tmp2 = cross(tmp, fW);
    b->torque = add(b->torque, tmp2);
  }
  // RigidBody.getSpatialVelocity :458-462
  static V3 getSpatialVelocity(const Body* b, const V3& pW) {
    V3 tmp = sub(pW, b->x);
    V3 r = cross(b->omega, tmp);
    return add(r, b->v);
  }
  // RigidBody.advanceVelocities :409-417
  static void advanceVelocitiesBase(Body* b, double dt) {
    b->v = scaleAdd(dt * b->minv, b->force, b->v);
    b->v = add(b->v, b->deltaV.v);
    V3 domega = transform(b->jinv, b->torque);
    domega = scale(dt, domega);
    b->omega = add(b->omega, domega);
    b->omega = add(b->omega, b->deltaV.w);
  }
  // RigidBody.expRodrigues :382-401
  static M3 expRodrigues(const V3& w, double t) {
    double wX = w.x, wY = w.y, wZ = w.z;
    double c = std::cos(t), s = std::sin(t);
    double c1 = 1 - c;
    M3 R;
    R.m00 = c + wX * wX * c1;
    R.m10 = wZ * s + wX * wY * c1;
    R.m20 = -wY * s + wX * wZ * c1;
    R.m01 = -wZ * s + wX * wY * c1;
    R.m11 = c + wY * wY * c1;
    R.m21 = wX * s + wY * wZ * c1;
    R.m02 = wY * s + wX * wZ * c1;
    R.m12 = -wX * s + wY * wZ * c1;
    R.m22 = c + wZ * wZ * c1;
    return R;
  }
  // RigidBody.advancePositions :427-441
  static void advancePositionsBase(Body* b, double dt) {
    b->x = scaleAdd(dt, b->v, b->x);
    double t = length(b->omega) * dt;
    if (t > 1e-8) {
      V3 domega = normalize(b->omega);
      M3 dR = expRodrigues(domega, t);
      dR = mul(dR, b->theta);
      b->theta = normalizeCP(dR);
    }
    updateRotationalInertiaFromTransformation(b);
  }
  // RigidCollection.applyVelocitiesTo :925-937
  static void applyVelocitiesTo(const Body* coll, Body* body) {
    V3 r = sub(body->x, coll->x);
    V3 wxr = cross(coll->omega, r);
    body->v = add(coll->v, wxr);
    body->omega = coll->omega;
  }
  // RigidCollection.updateBodiesPositionAndTransformations :898-909
  static void updateBodiesPositionAndTransformations(Body* coll) {
    for (Body* body : coll->bodies) {
      body->setB2W(Xf::mult(coll->B2W(), body->B2C));
      if (!coll->pinned) {
        Xf T = body->B2W();
        body->jinv = T.computeRM0RT(body->jinv0);
        body->massAngular = T.computeRM0RT(body->massAngular0);
      }
    }
  }
  static void advanceVelocities(Body* b, double dt) {
    advanceVelocitiesBase(b, dt);
    if (b->isCollection)
      for (Body* s : b->bodies) applyVelocitiesTo(b, s);
  }
  static void advancePositions(Body* b, double dt) {
    advancePositionsBase(b, dt);
    if (b->isCollection) updateBodiesPositionAndTransformations(b);
  }
  // RigidBody.wake :298-306
  static void wake(Body* b) {
    if (b->sleeping) { b->sleeping = false; b->metricHistory.clear(); }
    if (b->isInCollection()) wake(b->parent);
  }

  // RigidCollection.getOp :809-830
  static M3 getOp(const Body* body, const V3& com) {
    double x = body->x.x - com.x, y = body->x.y - com.y, z = body->x.z - com.z;
    double x2 = x * x, y2 = y * y, z2 = z * z;
    M3 op;
    op.m00 = y2 + z2; op.m01 = -x * y; op.m02 = -x * z;
    op.m10 = -y * x; op.m11 = x2 + z2; op.m12 = -y * z;
    op.m20 = -z * x; op.m21 = -z * y; op.m22 = x2 + y2;
    return scaleM(body->massLinear, op);  // Matrix3d.mul(scalar)
  }

  // RigidCollection.updateBB(RigidBody) :629-661
  static void updateBBWith(Body* coll, const Body* body) {
    if (coll->boundingBoxB.empty()) return;
    if (body->isPlane()) return;
    V3 bbmax(-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308);
    V3 bbmin(1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308);
    for (int i = 0; i < 2; i++) {
      const Body* b = (i == 0) ? coll : body;
      Xf T = b->B2W();
      for (const V3& point : b->boundingBoxB) {
        V3 p = T.transformP(point);
        bbmin.x = std::min(bbmin.x, p.x); bbmin.y = std::min(bbmin.y, p.y); bbmin.z = std::min(bbmin.z, p.z);
        bbmax.x = std::max(bbmax.x, p.x); bbmax.y = std::max(bbmax.y, p.y); bbmax.z = std::max(bbmax.z, p.z);
      }
    }
    setBBCorners(coll, bbmin, bbmax);
  }
  static void setBBCorners(Body* coll, const V3& bbmin, const V3& bbmax) {
    auto& B = coll->boundingBoxB;
    B[4] = bbmin;
    B[5] = V3(bbmin.x, bbmax.y, bbmax.z);
    B[6] = V3(bbmax.x, bbmin.y, bbmax.z);
    B[7] = V3(bbmax.x, bbmax.y, bbmin.z);
    B[0] = bbmax;
    B[1] = V3(bbmax.x, bbmin.y, bbmin.z);
    B[2] = V3(bbmin.x, bbmax.y, bbmin.z);
    B[3] = V3(bbmin.x, bbmin.y, bbmax.z);
  }
  // RigidCollection.updateBB() :596-623
  static void updateBBAll(Body* coll) {
    if (coll->boundingBoxB.empty()) return;
    V3 bbmax(-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308);
    V3 bbmin(1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308);
    for (const Body* body : coll->bodies) {
      if (body->isPlane()) continue;
      for (const V3& point : body->boundingBoxB) {
        V3 p = body->B2C.transformP(point);
        bbmin.x = std::min(bbmin.x, p.x); bbmin.y = std::min(bbmin.y, p.y); bbmin.z = std::min(bbmin.z, p.z);
        bbmax.x = std::max(bbmax.x, p.x); bbmax.y = std::max(bbmax.y, p.y); bbmax.z = std::max(bbmax.z, p.z);
      }
    }
    setBBCorners(coll, bbmin, bbmax);
  }

  // RigidCollection.updateCollectionState :459-469
  static void updateCollectionState(Body* coll, Body* body) {
    coll->pinned = coll->pinned || body->pinned;
    coll->sleeping = coll->sleeping || body->sleeping;
    body->sleeping = false;
  }

  // RigidCollection.addBodyInternalMethod :477-514
  static void addBodyInternalMethod(Body* coll, const Body* body) {
    Xf tmpTransformB2W = coll->B2W();
    coll->theta.setIdentity();  // updateTheta :544-582
    if (coll->pinned) {
      coll->v = V3(); coll->omega = V3();
      coll->massAngular.setZero(); coll->massAngular0.setZero(); coll->jinv.setZero(); coll->jinv0.setZero();
      coll->massLinear = 0; coll->minv = 0;
    } else {
      V3 com = coll->x;
      com = scale(coll->massLinear, com);
      com = scaleAdd(body->massLinear, body->x, com);
      double totalMassInv = 1. / (body->massLinear + coll->massLinear);
      com = scale(totalMassInv, com);
      // updateVelocitiesFrom :527-535
      coll->v = scale(coll->massLinear, coll->v);
      coll->v = scaleAdd(body->massLinear, body->v, coll->v);
      coll->v = scale(totalMassInv, coll->v);
      coll->omega = scale(coll->massLinear, coll->omega);
      coll->omega = scaleAdd(body->massLinear, body->omega, coll->omega);
      coll->omega = scale(totalMassInv, coll->omega);
      // updateInertia :774-786
      M3 acc;
      acc.setZero();
      for (int i = 0; i < 2; i++) {
        const Body* b = (i == 0) ? coll : body;
        acc = addM(acc, b->massAngular);
        acc = addM(acc, getOp(b, com));
      }
      coll->massAngular0 = acc;
      coll->massAngular = acc;
      updateBBWith(coll, body);
      coll->massLinear += body->massLinear;
      coll->minv = totalMassInv;
      coll->x = com;
      Xf T = coll->B2W();
      for (V3& point : coll->boundingBoxB) point = T.inverseTransformP(point);
    }
    coll->B2C = Xf::multAinvB(coll->B2W(), tmpTransformB2W);
  }

  // RigidCollection.updateInertiaRestAndInvert :667-673
  static void updateInertiaRestAndInvert(Body* coll) {
    if (!coll->pinned) {
      M3 inv;
      if (invert(coll->massAngular, inv)) coll->jinv = inv;
      Xf T = coll->B2W();
      coll->massAngular0 = T.computeRTMR(coll->massAngular);
      coll->jinv0 = T.computeRTMR(coll->jinv);
    }
  }
  // RigidCollection.updateBodiesTransformations :587-590
  static void updateBodiesTransformations(Body* coll) {
    for (Body* body : coll->bodies) body->B2C = Xf::multAinvB(coll->B2W(), body->B2W());
  }
  static void finishAdd(Body* coll) {
    updateInertiaRestAndInvert(coll);
    updateRotationalInertiaFromTransformation(coll);
    updateBodiesTransformations(coll);
  }
  // RigidCollection.addBody :128-139
  static void collAddBody(Body* coll, Body* body) {
    body->parent = coll;
    coll->bodies.push_back(body);
    updateCollectionState(coll, body);
    addBodyInternalMethod(coll, body);
    finishAdd(coll);
  }
  // RigidCollection.addBodies :145-158
  static void collAddBodies(Body* coll, const std::vector<Body*>& list) {
    for (Body* body : list) {
      body->parent = coll;
      coll->bodies.push_back(body);
      updateCollectionState(coll, body);
      addBodyInternalMethod(coll, body);
    }
    finishAdd(coll);
  }
  // RigidCollection.addCollection :164-178
  static void collAddCollection(Body* coll, Body* other) {
    for (Body* body : other->bodies) { body->parent = coll; coll->bodies.push_back(body); }
    updateCollectionState(coll, other);
    addBodyInternalMethod(coll, other);
    finishAdd(coll);
  }
  // RigidBody.set(RigidBody) :248-271
  static void bodySet(Body* dst, const Body* src) {
    dst->v = src->v; dst->omega = src->omega; dst->x = src->x; dst->theta = src->theta;
    dst->massLinear = src->massLinear; dst->minv = src->minv;
    dst->massAngular = src->massAngular; dst->massAngular0 = src->massAngular0;
    dst->jinv = src->jinv; dst->jinv0 = src->jinv0;
    dst->pinned = src->pinned; dst->sleeping = src->sleeping;
    dst->boundingBoxB = src->boundingBoxB;
  }
  // new RigidCollection(body1, body2) :55-78
  Body* newCollection(Body* body1, Body* body2) {
    collArena.emplace_back(new Body());
    Body* c = collArena.back().get();
    c->isCollection = true;
    c->collectionSlot = nextCollectionSlot++;
    if (body1->isPlane()) {
      bodySet(c, body2);
      body2->parent = c;
      c->bodies.push_back(body2);
      collAddBody(c, body1);
    } else {
      bodySet(c, body1);
      body1->parent = c;
      c->bodies.push_back(body1);
      collAddBody(c, body2);
    }
    return c;
  }
  // RigidCollection.unmergeBody :944-953
  static void unmergeBody(Body* coll, Body* body) {
    if (body->parent != coll) return;
    applyVelocitiesTo(coll, body);
    body->deltaV.setZero();
    body->parent = nullptr;
  }
  // RigidCollection.addToInternalContact :836-847
  void addToInternalContact(Body* coll, BPC* bpc) {
    std::vector<Contact*> tmp;
    for (Contact* contact : bpc->contactList) {
      heapContacts.emplace_back(new Contact(*contact));
      Contact* c = heapContacts.back().get();
      c->newThisTimeStep = false;
      c->internal = true;
      // Contact(Contact) copies: bodies, ids, frame, lambda, J, state, newThisTimeStep, violations (:120-150);
      // lambda*warm, b*, D* keep their default 0
      c->lambda0warm = c->lambda1warm = c->lambda2warm = 0;
      c->bn = c->bt1 = c->bt2 = 0;
      c->D00 = c->D11 = c->D22 = 0;
      tmp.push_back(c);
      coll->internalContacts.push_back(c);
    }
    bpc->contactList = tmp;
  }
  // RigidCollection.addBPCsToCollection :854-863
  static void addBPCsToCollection(BPC* bpc) {
    addToBodyListsParent(bpc);
    for (int i = 0; i < 2; i++) {
      Body* body = bpc->getBody(i);
      for (BPC* ext : body->bodyPairContacts) addToBodyListsParent(ext);
    }
  }
  // RigidCollection.addIncompleteContacts :987-998
  void addIncompleteContacts(Body* body, std::vector<BPC*>& removalQueue) {
    std::vector<BPC*> snapshot(body->bodyPairContacts.begin(), body->bodyPairContacts.end());
    for (BPC* bpc : snapshot) {
      if (bpc->body1->isInSameCollection(bpc->body2) && !bpc->inCollection) {
        bpc->inCollection = true;
        bpc->motionMetricHist.clear();
        bpc->contactStateHist.clear();
        addToInternalContact(body->parent, bpc);
        addBPCsToCollection(bpc);
        removalQueue.push_back(bpc);
        events.push_back(Event{totalSteps, 0, bpc->lo(), bpc->hi()});
      }
    }
  }
  // RigidCollection.fillInternalBodyContacts :960-978
  static void fillInternalBodyContacts(Body* coll) {
    coll->bodyPairContacts.clear();
    coll->internalContacts.clear();
    for (Body* body : coll->bodies) {
      for (BPC* bpc : body->bodyPairContacts) {
        if (coll->bodyPairContacts.insert(bpc).second) {
          Body* other = bpc->getOtherBody(body);
          if (body->isInSameCollection(other)) {
            bpc->inCollection = true;
            for (Contact* c : bpc->contactList)
              if (std::find(coll->internalContacts.begin(), coll->internalContacts.end(), c) == coll->internalContacts.end())
                coll->internalContacts.push_back(c);
          }
        }
      }
    }
  }
  // RigidCollection.removeBodies :686-753
  static void collRemoveBodies(Body* coll, const std::vector<Body*>& toRemove) {
    for (Body* r : toRemove) {
      auto it = std::find(coll->bodies.begin(), coll->bodies.end(), r);
      if (it != coll->bodies.end()) coll->bodies.erase(it);
    }
    Xf tmpTransformB2W = coll->B2W();
    bool wasPinned = coll->pinned;
    coll->pinned = false;
    for (Body* body : coll->bodies) coll->pinned = coll->pinned || body->pinned;
    coll->theta.setIdentity();
    if (coll->pinned) {
      coll->v = V3(); coll->omega = V3();
      coll->massAngular.setZero(); coll->massAngular0.setZero(); coll->jinv.setZero(); coll->jinv0.setZero();
      coll->massLinear = 0; coll->minv = 0;
    } else if (wasPinned) {
      coll->x = V3();
      coll->massLinear = 0;
      for (Body* body : coll->bodies) {
        coll->x = scaleAdd(body->massLinear, body->x, coll->x);
        coll->massLinear += body->massLinear;
      }
      coll->minv = 1. / coll->massLinear;
      coll->x = scale(coll->minv, coll->x);
      // computeInertia :760-767
      coll->massAngular.setZero();
      for (Body* body : coll->bodies) {
        coll->massAngular = addM(coll->massAngular, body->massAngular);
        coll->massAngular = addM(coll->massAngular, getOp(body, coll->x));
      }
      updateInertiaRestAndInvert(coll);
    } else {
      for (Body* body : toRemove) {
        V3 com = coll->x;
        com = scale(coll->massLinear, com);
        coll->x = scaleAdd(-body->massLinear, body->x, com);
        com = scale(1. / coll->massLinear, com);
        coll->massLinear -= body->massLinear;
        coll->minv = 1. / coll->massLinear;
        coll->x = scale(coll->minv, coll->x);
        // updateInertiaReverse :793-802
        coll->massAngular = subM(coll->massAngular, body->massAngular);
        for (int i = 0; i < 2; i++) {
          const Body* b = (i == 0) ? coll : body;
          coll->massAngular = subM(coll->massAngular, getOp(b, com));
        }
      }
      updateInertiaRestAndInvert(coll);
    }
    updateRotationalInertiaFromTransformation(coll);
    updateBodiesTransformations(coll);
    updateBBAll(coll);
    coll->B2C = Xf::multAinvB(coll->B2W(), tmpTransformB2W);
  }

  // ==========================================================================================
  // MotionMetricProcessor.java:39-73, BodyPairContact.java:83-205,255-266
  // ==========================================================================================
  static double largestVelocityNorm2(const Body* body1, const Body* body2) {
    double largest = 0;
    for (int i = 0; i < 2; i++) {
      const Body* body = (i == 0) ? body1 : body2;
      Xf T = body->B2W();
      for (const V3& point : body->boundingBoxB) {
        V3 pW = T.transformP(point);
        V3 v1 = getSpatialVelocity(body1, pW);
        V3 v2 = getSpatialVelocity(body2, pW);
        v1 = sub(v1, v2);
        largest = std::max(length(v1), largest);
      }
    }
    return largest;
  }
  static double largestVelocityNorm1(const Body* body) {
    double largest = 0;
    Xf T = body->B2W();
    for (const V3& point : body->boundingBoxB) {
      V3 pW = T.transformP(point);
      V3 v1 = getSpatialVelocity(body, pW);
      largest = std::max(length(v1), largest);
    }
    return largest;
  }
  // MotionMetricProcessor.getLargestVelocityNorm(body1, body2, dt) :75-116 (metricPositionLevel)
  static double largestVelocityNormPos(const Body* body1, const Body* body2, double dt) {
    double largest = 0;
    Body a1, a2;  // body1Advanced.set(body1); advancePositions(dt)
    a1.x = body1->x; a1.theta = body1->theta; a1.v = body1->v; a1.omega = body1->omega;
    a1.massAngular0 = body1->massAngular0; a1.jinv0 = body1->jinv0;
    advancePositionsBase(&a1, dt);
    a2.x = body2->x; a2.theta = body2->theta; a2.v = body2->v; a2.omega = body2->omega;
    a2.massAngular0 = body2->massAngular0; a2.jinv0 = body2->jinv0;
    advancePositionsBase(&a2, dt);
    Xf T1 = body1->B2W(), T2 = body2->B2W(), A1 = a1.B2W(), A2 = a2.B2W();
    for (const V3& point : body1->boundingBoxB) {
      V3 pW = T1.transformP(point);
      V3 pB = T2.inverseTransformP(pW);
      pW = A2.transformP(pB);
      pB = A1.inverseTransformP(pW);
      V3 v1 = scale(1. / dt, sub(point, pB));
      largest = std::max(length(v1), largest);
    }
    for (const V3& point : body2->boundingBoxB) {
      V3 pW = T2.transformP(point);
      V3 pB = T1.inverseTransformP(pW);
      pW = A1.transformP(pB);
      pB = A2.inverseTransformP(pW);
      V3 v2 = scale(1. / dt, sub(point, pB));
      largest = std::max(length(v2), largest);
    }
    return largest;
  }
  double pairMetric(const Body* body1, const Body* body2) {
    if (body1->pinned) return largestVelocityNorm1(body2);
    if (body2->pinned) return largestVelocityNorm1(body1);
    if (P.metric_position_level) return largestVelocityNormPos(body1, body2, curDt);  // BodyPairContact.java:92-93, :134-135
    return largestVelocityNorm2(body1, body2);
  }
  // accumulateForMerging :83-121
  void bpcAccumulateForMerging(BPC* bpc) {
    Body* body1 = bpc->body1->isInCollection() ? bpc->body1->parent : bpc->body1;
    Body* body2 = bpc->body2->isInCollection() ? bpc->body2->parent : bpc->body2;
    bpc->motionMetricHist.push_back(pairMetric(body1, body2));
    if ((int)bpc->motionMetricHist.size() > P.step_accum_merging) bpc->motionMetricHist.erase(bpc->motionMetricHist.begin());
    ContactState state = CLEAR;
    int notOnEdgeCount = 0;
    for (Contact* c : bpc->contactList)
      if (c->state != ONEDGE) notOnEdgeCount++;
    if (notOnEdgeCount == 0) state = ONEDGE;
    else if (notOnEdgeCount >= 2) state = CLEAR;
    else if (bpc->contactList.size() == 1) state = CLEAR;
    else state = ONEDGE;
    bpc->contactStateHist.push_back(state);
    if ((int)bpc->contactStateHist.size() > P.step_accum_merging) bpc->contactStateHist.erase(bpc->contactStateHist.begin());
  }
  // accumulateForUnmerging :127-145
  void bpcAccumulateForUnmerging(BPC* bpc) {
    if (bpc->body1->isInSameCollection(bpc->body2)) {
      double metric = pairMetric(bpc->body1, bpc->body2);
      if (metric > P.threshold_unmerge) bpc->motionMetricHist.push_back(metric);
      else bpc->motionMetricHist.clear();
    }
  }
  // checkMergeCondition :151-177 (cycle condition unsupported)
  bool checkMergeCondition(const BPC* bpc) const {
    if (bpc->body1->sleeping && bpc->body2->sleeping) return true;
    if (P.merge_let_it_breathe)
      for (const Contact* c : bpc->contactList)
        if (std::fabs(c->prevConstraintViolation - c->constraintViolation) > P.threshold_breath) return false;
    if (!P.merge_pinned && (bpc->body1->pinned || bpc->body2->pinned)) return false;
    if (bpc->body1->isInSameCollection(bpc->body2)) return false;
    // checkMotionMetricForMerging :187-198
    if ((int)bpc->motionMetricHist.size() == P.step_accum_merging) {
      for (double m : bpc->motionMetricHist)
        if (m > P.threshold_merge) return false;
    } else {
      return false;
    }
    if (P.merge_stable_contact) {  // areContactsStable :212-224
      if ((int)bpc->contactStateHist.size() == P.step_accum_merging) {
        for (ContactState s : bpc->contactStateHist)
          if (s == ONEDGE) return false;
      } else {
        return false;
      }
    }
    return true;
  }
  // checkContactsState :255-266
  bool checkContactsState(const BPC* bpc) const {
    for (const Contact* c : bpc->contactList) {
      if (c->state == BROKEN && P.unmerge_normal) return true;
      if (c->state == ONEDGE && P.unmerge_friction) return true;
    }
    return false;
  }
  // checkMotionMetricForUnmerging :200-205
  bool checkMotionMetricForUnmerging(const BPC* bpc) const {
    return (int)bpc->motionMetricHist.size() >= P.step_accum_unmerging && bpc->contactList.size() < 3;
  }

  // ==========================================================================================
  // Merging.java
  // ==========================================================================================
  void removeBody(Body* b) {
    auto it = std::find(bodies.begin(), bodies.end(), b);
    if (it != bodies.end()) bodies.erase(it);
  }
  static void mergeBpcSets(Body* dst, Body* src) {
    for (BPC* b : src->bodyPairContacts) dst->bodyPairContacts.insert(b);
  }

  // merge :73-163
  void merge() {
    if (!P.enable_merging) return;
    std::vector<BPC*> removalQueue;
    T.merging_build = 0;
    std::vector<BPC*> snapshot(bodyPairContacts.begin(), bodyPairContacts.end());
    for (BPC* bpc : snapshot) {
      if (!bpc->inCollection && checkMergeCondition(bpc)) {
        double t0 = nowSec();
        mergingEvent = true;
        bpc->inCollection = true;
        bpc->motionMetricHist.clear();
        bpc->contactStateHist.clear();
        events.push_back(Event{totalSteps, 0, bpc->lo(), bpc->hi()});
        removalQueue.push_back(bpc);
        Body *b1 = bpc->body1, *b2 = bpc->body2;
        if (!b1->isInCollection() && !b2->isInCollection()) {
          removeBody(b1);
          removeBody(b2);
          Body* collection = newCollection(b1, b2);
          addToInternalContact(collection, bpc);
          addBPCsToCollection(bpc);
          bodies.push_back(collection);
        } else if (b1->isInCollection() && b2->isInCollection()) {
          Body *big, *small;
          if (b1->parent->bodies.size() > b2->parent->bodies.size()) { big = b1->parent; small = b2->parent; }
          else { big = b2->parent; small = b1->parent; }
          removeBody(small);
          mergeBpcSets(big, small);
          big->internalContacts.insert(big->internalContacts.end(), small->internalContacts.begin(), small->internalContacts.end());
          collAddCollection(big, small);
          addToInternalContact(big, bpc);
          // addIncompleteCollectionContacts :1004-1008 — the reference passes bpc.bodyX.parent, which
          // after addCollection is already the absorbing collection, so every member is visited
          std::vector<Body*> allBodies = big->bodies;
          for (Body* body : allBodies) addIncompleteContacts(body, removalQueue);
          addBPCsToCollection(bpc);
        } else if (b1->isInCollection()) {
          removeBody(b2);
          collAddBody(b1->parent, b2);
          addToInternalContact(b1->parent, bpc);
          addIncompleteContacts(b2, removalQueue);
          addBPCsToCollection(bpc);
        } else {
          removeBody(b1);
          collAddBody(b2->parent, b1);
          addToInternalContact(b2->parent, bpc);
          addIncompleteContacts(b1, removalQueue);
          addBPCsToCollection(bpc);
        }
        T.merging_build += nowSec() - t0;
      }
    }
    for (BPC* bpc : removalQueue) bodyPairContacts.erase(bpc);
  }

  struct BodyIdLess {
    bool operator()(const Body* a, const Body* b) const { return a->id != b->id ? a->id < b->id : a < b; }
  };
  // buildNeighborBody :381-393
  void buildNeighborBody(Body* body, std::set<Body*, BodyIdLess>& sub, const std::set<Body*, BodyIdLess>& handled) {
    std::vector<BPC*> snapshot(body->bodyPairContacts.begin(), body->bodyPairContacts.end());
    for (BPC* bpc : snapshot) {
      if (!bpc->inCollection) continue;
      Body* other = bpc->getOtherBody(body);
      if (other == nullptr) continue;
      if (!sub.count(other) && !handled.count(other)) {
        sub.insert(other);
        buildNeighborBody(other, sub, handled);
      }
    }
  }

  // unmergeSelectedBpcs :286-374
  bool unmergeSelectedBpcs(Body* collection, std::set<BPC*, BpcLess>& bpcsToUnmerge, std::vector<Body*>& newBodies) {
    bool removeCollection = true;
    for (BPC* bpc : bpcsToUnmerge) bpc->inCollection = false;
    std::set<Body*, BodyIdLess> handledBodies, subbodies, remainedBodies;
    size_t collSize = collection->bodies.size();
    // collection.bodies is in order of addition, which in the reference derives from the HashSet iteration order
    // of merge(); canonicalised to ascending body id (unmergeBody does not touch the list)
    std::vector<Body*> members = collection->bodies;
    std::sort(members.begin(), members.end(), [](const Body* a, const Body* b) { return a->id < b->id; });
    for (Body* body : members) {
      if (!handledBodies.count(body)) {
        subbodies.insert(body);
        buildNeighborBody(body, subbodies, handledBodies);
        handledBodies.insert(subbodies.begin(), subbodies.end());
        if (collSize != subbodies.size()) {
          if (subbodies.size() < collSize / 2 + 1) {
            for (Body* b : subbodies) unmergeBody(collection, b);
            if (subbodies.size() > 1) {
              auto iter = subbodies.begin();
              Body* sb1 = *iter++;
              Body* sb2 = *iter;
              subbodies.erase(sb1);
              subbodies.erase(sb2);
              Body* nc = newCollection(sb1, sb2);
              collAddBodies(nc, std::vector<Body*>(subbodies.begin(), subbodies.end()));
              fillInternalBodyContacts(nc);
              newBodies.push_back(nc);
            } else if (subbodies.size() == 1) {
              newBodies.push_back(*subbodies.begin());
            }
          } else {
            removeCollection = false;
            remainedBodies.insert(subbodies.begin(), subbodies.end());
          }
        } else {
          removeCollection = false;
          remainedBodies.insert(subbodies.begin(), subbodies.end());
          break;
        }
        subbodies.clear();
      }
    }
    for (BPC* bpc : bpcsToUnmerge) {
      if (bpc->body1->isInSameCollection(bpc->body2)) {
        bpc->inCollection = true;
      } else {
        events.push_back(Event{totalSteps, 1, bpc->lo(), bpc->hi()});
        if (!bodyPairContacts.count(bpc)) {
          bodyPairContacts.insert(bpc);
          for (Contact* c : bpc->contactList) {
            c->lambda0warm = c->lambda0;
            c->lambda1warm = c->lambda1;
            c->lambda2warm = c->lambda2;
            c->internal = false;
            contacts.push_back(c);
          }
        }
        bpc->motionMetricHist.clear();
        bpc->contactStateHist.clear();
      }
    }
    if (!removeCollection) {
      if (remainedBodies.size() != collection->bodies.size()) {
        std::vector<Body*> toRemove;
        for (Body* b : handledBodies)
          if (!remainedBodies.count(b)) toRemove.push_back(b);
        collRemoveBodies(collection, toRemove);
        fillInternalBodyContacts(collection);
      }
    }
    return removeCollection;
  }

  // unmerge :215-273
  void unmerge() {
    if (!P.enable_unmerging) return;
    T.unmerging_build = 0;
    if (!P.unmerge_relative_motion && !P.unmerge_normal && !P.unmerge_friction) return;
    std::vector<Body*> removalQueue, additionQueue;
    for (Body* body : bodies) {
      if (body->sleeping) continue;
      if (body->isCollection) {
        std::set<BPC*, BpcLess> bpcsToUnmerge;
        for (BPC* bpc : body->bodyPairContacts) {
          if (!bpc->inCollection) continue;
          if (!bpcsToUnmerge.count(bpc)) {
            if (!checkContactsState(bpc) && !checkMotionMetricForUnmerging(bpc)) continue;
            // addBpcToUnmerge :301-309
            if (!bpc->body1->isInCollection() || !bpc->body2->isInCollection() || !bpc->body1->isInSameCollection(bpc->body2)) continue;
            bpcsToUnmerge.insert(bpc);
          }
        }
        double t0 = nowSec();
        std::vector<Body*> newBodies;
        bool removeCollection = false;
        if (!bpcsToUnmerge.empty()) removeCollection = unmergeSelectedBpcs(body, bpcsToUnmerge, newBodies);
        if (!newBodies.empty()) {
          mergingEvent = true;
          additionQueue.insert(additionQueue.end(), newBodies.begin(), newBodies.end());
          if (removeCollection) removalQueue.push_back(body);
        }
        T.unmerging_build += nowSec() - t0;
      }
    }
    bodies.insert(bodies.end(), additionQueue.begin(), additionQueue.end());
    for (Body* b : removalQueue) removeBody(b);
  }

  // ==========================================================================================
  // single sweep: CollisionProcessor.java:232-441
  // ==========================================================================================
  bool hasCollections() const {
    for (Body* b : bodies) if (b->isCollection) return true;
    return false;
  }
  void getNextLayer(std::vector<BPC*> layer, std::vector<BPC*>& ordered) {
    int depth = 0;
    while (true) {
      std::vector<BPC*> next;
      depth++;
      for (BPC* bpc : layer) {
        for (int i = 0; i < 2; i++) {
          Body* body = bpc->getBody(i);
          for (BPC* other : body->bodyPairContacts) {
            if (!other->checked) { next.push_back(other); other->checked = true; other->layer = 1 + depth; }  // classes: 0 new, 1 picked, 2.. layers
          }
        }
      }
      if (next.empty()) break;
      ordered.insert(ordered.end(), next.begin(), next.end());
      layer = next;
    }
    lastDepth = depth;
  }
  // getOrganizedContacts :346-419
  int lastDepth = 0;
  void getOrganizedContacts(std::vector<Contact*>& out) {
    std::vector<BPC*> ordered;
    lastDepth = 0;
    for (BPC* bpc : bodyPairContacts) {
      for (Contact* c : bpc->contactList) {
        if (c->newThisTimeStep) { ordered.push_back(bpc); bpc->checked = true; bpc->layer = 0; break; }
      }
    }
    // second: all bpc of bodies the user interacted with (:361-383)
    auto takePicked = [&](Body* b) {
      if (!b->picked) return;
      for (BPC* bpc : b->bodyPairContacts)
        if (!bpc->checked) { ordered.push_back(bpc); bpc->checked = true; bpc->layer = 1; }
      b->picked = false;
    };
    for (Body* body : bodies) {
      if (body->isCollection) { for (Body* b : body->bodies) takePicked(b); }
      else takePicked(body);
    }
    if (!ordered.empty()) getNextLayer(ordered, ordered);
    for (BPC* bpc : bodyPairContacts)
      if (!bpc->checked) { ordered.push_back(bpc); bpc->checked = true; bpc->layer = lastDepth + 2; }
    for (Body* body : bodies) {
      if (body->isCollection && !body->sleeping) {
        for (BPC* bpc : body->bodyPairContacts)
          if (!bpc->checked) { ordered.push_back(bpc); bpc->checked = true; bpc->layer = lastDepth + 3; }
      }
    }
    for (BPC* bpc : ordered)
      for (Contact* c : bpc->contactList) { out.push_back(c); sweepLayerOf[c] = bpc->layer; }
  }
  std::map<Contact*, int> sweepLayerOf;  // test bookkeeping: getOrganizedContacts class of every contact of the sweep
  static void updateJacobiansThatNeedUpdating(std::vector<Contact*>& list, bool computeInCollection) {
    for (Contact* c : list)
      if (c->body1->parent != nullptr || c->body2->parent != nullptr) computeJacobian(c, computeInCollection);
  }

  bool keyMatches(const am3d_contact& k, const Contact* c) const {
    int c1 = c->csb1 ? c->csb1->part : -1, c2 = c->csb2 ? c->csb2->part : -1;
    return k.body1 == c->body1->id && k.body2 == c->body2->id && k.csb1 == c1 && k.csb2 == c2 && k.bv1 == c->bv1 &&
           k.bv2 == c->bv2 && k.info == c->info && k.leaf == c->leaf;
  }
  struct FullKey {
    int v[8];
    bool operator<(const FullKey& o) const { return std::lexicographical_compare(v, v + 8, o.v, o.v + 8); }
  };
  static FullKey fullKey(const Contact* c) {
    FullKey k{{c->body1->id, c->body2->id, c->csb1 ? c->csb1->part : -1, c->csb2 ? c->csb2->part : -1, c->bv1, c->bv2,
               c->info, c->leaf}};
    return k;
  }
  // re-order `list` to follow the externally supplied key sequence (the CUDA path's colour order)
  // returns true if the supplied sequence carries hub flags (then `meta` is aligned with the re-ordered list)
  bool applyOrder(std::vector<Contact*>& list, const std::vector<am3d_contact>& order, std::vector<OrderMeta>& meta) {
    std::map<FullKey, std::vector<Contact*>> byKey;
    for (Contact* c : list) byKey[fullKey(c)].push_back(c);
    std::vector<Contact*> out;
    std::set<Contact*> used;
    meta.clear();
    bool anyHub = false;
    int lastColor = 0;
    for (const am3d_contact& k : order) {
      FullKey fk{{k.body1, k.body2, k.csb1, k.csb2, k.bv1, k.bv2, k.info, k.leaf}};
      auto it = byKey.find(fk);
      if (it == byKey.end() || it->second.empty()) { orderMismatch++; continue; }
      Contact* c = it->second.front();
      it->second.erase(it->second.begin());
      out.push_back(c);
      used.insert(c);
      meta.push_back(OrderMeta{k.color, k.hub_mask});
      lastColor = k.color;
      if (k.hub_mask) anyHub = true;
    }
    for (Contact* c : list)
      if (!used.count(c)) { orderMismatch++; out.push_back(c); meta.push_back(OrderMeta{lastColor + 1, 0}); }
    list = out;
    return anyHub;
  }

  // updateInCollections :232-303
  void updateInCollections(double dt) {
    T.update_collections = T.contact_ordering = T.single_it_pgs = 0;
    lastSweepList.clear();
    if (!P.update_contacts_in_collections) return;
    double t0 = nowSec();
    if (!hasCollections()) return;
    std::vector<Contact*> list;
    sweepLayerOf.clear();
    if (P.organize_contacts) {
      double t1 = nowSec();
      getOrganizedContacts(list);
      T.contact_ordering = nowSec() - t1;
    } else {
      list = contacts;
      for (Contact* c : list) sweepLayerOf[c] = 0;
      for (Body* body : bodies)
        if (body->isCollection && !body->sleeping)
          for (Contact* c : body->internalContacts) { list.push_back(c); sweepLayerOf[c] = 1; }
    }
    std::vector<OrderMeta> meta;
    bool hubs = haveOrderSweep && applyOrder(list, orderSweep, meta);
    if (haveOrderSweep) {
      // the replayed sequence may permute contacts only INSIDE a class of the reference's order (new-contact pairs,
      // each breadth-first layer, the unconnected rest): classes must appear in the reference's sequence
      int prev = -1;
      for (Contact* c : list) {
        int L = sweepLayerOf[c];
        if (L < prev) orderMismatch++;
        prev = std::max(prev, L);
      }
    }
    updateJacobiansThatNeedUpdating(list, true);
    double t2 = nowSec();
    if (hubs) pgsSolveHub(list, meta, dt, P.iterations_in_collection, 1e-5, 1., P.feedback_stiffness, P.enable_compliance ? P.compliance : 0., true);
    else pgsSolve(list, dt, P.iterations_in_collection, 1e-5, 1., P.feedback_stiffness, P.enable_compliance ? P.compliance : 0., true);
    T.single_it_pgs = nowSec() - t2;
    lastSweepList = list;
    for (Body* body : bodies) {
      if (body->isCollection && !body->sleeping) {
        for (Body* b : body->bodies)
          if (!b->pinned) advanceVelocitiesBase(b, dt);
      }
      body->deltaV.setZero();
    }
    T.update_collections = nowSec() - t0;
  }

  // solveLCP :108-137
  void solveLCP(double dt, bool postStabilization = false) {
    if (!contacts.empty()) {
      // the externally supplied sequence only orders the Gauss-Seidel sweeps; `contacts` itself keeps the
      // reference's emission order (it decides which duplicate-key contact the warm-start map retains, :443-448)
      std::vector<Contact*> list = contacts;
      std::vector<OrderMeta> meta;
      bool haveOrder = postStabilization ? haveOrderPost : haveOrderFull;
      bool hubs = haveOrder && applyOrder(list, postStabilization ? orderPost : orderFull, meta);
      updateJacobiansThatNeedUpdating(list, false);
      // :119: the velocity solve drops the Baumgarte term when post-stabilisation takes care of the drift
      double feedback = (!P.enable_post_stabilization || postStabilization) ? P.feedback_stiffness : 0.;
      postStabSolve = postStabilization;
      double t0 = nowSec();
      if (hubs) pgsSolveHub(list, meta, dt, P.iterations, P.tolerance, P.omega, feedback, P.enable_compliance ? P.compliance : 0., false);
      else pgsSolve(list, dt, P.iterations, P.tolerance, P.omega, feedback, P.enable_compliance ? P.compliance : 0., false);
      postStabSolve = false;
      if (!postStabilization) {
        T.lcp_solve = nowSec() - t0;
        T.pgs_iterations = lastIterations;
        rowUpdates += 3L * (long)contacts.size() * lastIterations;
        solveSeconds += T.lcp_solve;
      }
    } else if (!postStabilization) {
      T.lcp_solve = 0;
      T.pgs_iterations = 0;
    }
  }

  // postStabilization (RigidBodySystem.java:354-377): a second detection at the advanced positions and a position-level
  // solve whose right-hand side is feedbackStiffness * violation; the bodies are then moved by deltaV
  void postStabilization(double dt) {
    for (Body* b : bodies) {
      b->clear();
      if (b->isCollection)
        for (Body* s : b->bodies) { applyVelocitiesTo(b, s); s->clear(); }  // RigidCollection.clearBodies :100-105
    }
    updateContactsMap();
    collisionDetection();
    updateBodyPairContacts();
    warmStart(true);
    solveLCP(dt, true);
    clearBodyPairContacts();
    for (Body* b : bodies) {
      if (b->pinned || b->sleeping) continue;
      // RigidBody.advancePositionsPostStabilization :423-425 = advancePositions(dt, deltaV.v, deltaV.w)
      V3 v = b->v, w = b->omega;
      b->v = b->deltaV.v; b->omega = b->deltaV.w;
      advancePositions(b, dt);
      b->v = v; b->omega = w;
    }
  }

  // ==========================================================================================
  // RigidBodySystem.java
  // ==========================================================================================
  // applyGravityForce :276-290
  void applyGravityForce() {
    double theta = P.gravity_angle_deg / 180.0 * M_PI;
    double g = P.gravity_amount;
    V3 dir(g * std::cos(theta), g * std::sin(theta), 0);
    for (Body* body : bodies) {
      body->force = add(body->force, scale(-body->massLinear, dir));
      if (body->isCollection)
        for (Body* b : body->bodies) b->force = add(b->force, scale(-b->massLinear, dir));
    }
  }
  // Spring.applyOneBody :153-173 / applyTwoBodies :182-209
  void applySpring(Spring& s) {
    double ks = P.spring_k_mod, ds = P.spring_d_mod;
    if (s.type == AM3D_SPRING_BODYBODY) {
      V3 pb1W = s.body1->B2W().transformP(s.pb1);
      V3 pb2W = s.body2->B2W().transformP(s.pb2);
      V3 displacement = sub(pb2W, pb1W);
      double dist = length(displacement);
      if (dist < 1e-3) return;
      displacement = scale(1. / dist, displacement);
      V3 velocity2 = getSpatialVelocity(s.body2, pb2W);
      V3 velocity1 = getSpatialVelocity(s.body1, pb1W);
      V3 rel = sub(velocity2, velocity1);
      double forceIntensity = s.k * ks * (dist - s.l0 * s.ls) + s.d * ds * dot(rel, displacement);
      V3 force = scale(forceIntensity, displacement);
      applyForceW(s.body1, pb1W, force);
      if (s.body1->isInCollection()) applyForceW(s.body1->parent, pb1W, force);
      force = scale(-forceIntensity, displacement);
      applyForceW(s.body2, pb2W, force);
      if (s.body2->isInCollection()) applyForceW(s.body2->parent, pb2W, force);
    } else {
      V3 pb1W = s.body1->B2W().transformP(s.pb1);
      V3 displacement = sub(s.pw, pb1W);
      double len = length(displacement);
      if (len < 1e-3) return;
      V3 velocity1 = getSpatialVelocity(s.body1, pb1W);
      double sc = -(s.k * ks * (len - s.l0 * s.ls) - s.d * ds * (dot(velocity1, displacement) / len)) / len;
      V3 force = scale(-sc, displacement);
      applyForceW(s.body1, pb1W, force);
      if (s.body1->isInCollection()) applyForceW(s.body1->parent, pb1W, force);
    }
  }
  // applyExternalForces :234-257 (mouse spring / impulse are UI objects above the boundary)
  // gyroscopicStabilization :212-229
  void gyroscopicStabilization(double dt) {
    for (Body* body : bodies) {
      if (body->pinned || body->sleeping) continue;
      V3 L = transform(body->massAngular, body->omega);
      M3 Lhat;
      Lhat.m00 = 0.; Lhat.m01 = -L.z; Lhat.m02 = L.y;
      Lhat.m10 = L.z; Lhat.m11 = 0.; Lhat.m12 = -L.x;
      Lhat.m20 = -L.y; Lhat.m21 = L.x; Lhat.m22 = 0.;
      Lhat = scaleM(dt, Lhat);
      M3 Tm = mul(Lhat, body->jinv);
      Tm = mul(Tm, Lhat);
      Tm = scaleM(-1., Tm);
      body->massAngular = addM(body->massAngular, Tm);
    }
  }
  // RigidBody.applyCoriolisTorque :348-354
  static void applyCoriolisTorque(Body* b) {
    if (b->pinned) return;
    V3 tmp = transform(b->massAngular, b->omega);
    V3 tmp2 = cross(tmp, b->omega);
    b->torque = sub(b->torque, tmp2);
  }
  // applyCoriolis :295-304
  void applyCoriolis() {
    for (Body* body : bodies) {
      applyCoriolisTorque(body);
      if (body->isCollection)
        for (Body* b : body->bodies) applyCoriolisTorque(b);
    }
  }
  double curDt = 0.05;
  // the UI's mouse tools as the step sees them
  struct Mouse {
    Body* springBody = nullptr;
    V3 grabB, pointW;
    double k = 50, c = 10;
    bool atCOM = false;
    Body* impBody = nullptr;
    int impPhase = 0;  // 1 released, 2 Impulse.holdingForce
    V3 impPointB, impEndW, heldPointW, heldForce;
    double impScale = 1;
  } mouse;
  // MouseSpringForce.apply :69-101
  void mouseSpringApply() {
    Body* picked = mouse.springBody;
    if (picked == nullptr) return;
    wake(picked);
    picked->picked = true;
    V3 grabPointBW = picked->B2W().transformP(mouse.grabB);
    double dist = distance(grabPointBW, mouse.pointW);
    V3 direction = sub(mouse.pointW, grabPointBW);
    if (dot(direction, direction) < 1e-3) return;
    direction = normalize(direction);
    V3 force = scale(dist * mouse.k, direction);
    if (picked->isInCollection()) applyForceW(picked->parent, grabPointBW, force);
    applyForceW(picked, mouse.atCOM ? picked->x : grabPointBW, force);
    V3 grabPointV = getSpatialVelocity(picked, grabPointBW);
    force = scale(-dot(grabPointV, direction) * mouse.c, direction);
    if (picked->isInCollection()) applyForceW(picked->parent, grabPointBW, force);
    applyForceW(picked, mouse.atCOM ? picked->x : grabPointBW, force);
  }
  // RigidBodySystem.java:249-256 with MouseImpulse.apply :101-125 and applyImpulse :262-267
  void mouseImpulseApply() {
    if (mouse.impPhase == 1) {
      Body* b = mouse.impBody;
      wake(b);
      b->picked = true;
      V3 pickedPointW = b->B2W().transformP(mouse.impPointB);
      double dist = distance(mouse.impEndW, pickedPointW);
      V3 direction = sub(pickedPointW, mouse.impEndW);
      mouse.impPhase = 0;
      if (dot(direction, direction) < 1e-3) return;
      direction = normalize(direction);
      V3 force = scale(mouse.impScale * dist, direction);
      if (b->isInCollection()) applyForceW(b->parent, pickedPointW, force);
      applyForceW(b, pickedPointW, force);
      mouse.heldPointW = pickedPointW;
      mouse.heldForce = force;
      mouse.impPhase = 2;
    } else if (mouse.impPhase == 2) {
      applyForceW(mouse.impBody, mouse.heldPointW, mouse.heldForce);
      mouse.impPhase = 0;
    }
  }
  void applyExternalForces() {
    if (P.use_gravity) applyGravityForce();
    if (P.use_coriolis) { gyroscopicStabilization(curDt); applyCoriolis(); }  // :239-242
    mouseSpringApply();
    if (P.springs_enabled)
      for (Spring& s : springs) applySpring(s);
    mouseImpulseApply();
  }
  // clearBodies :190-202, RigidCollection.clearBodies :100-105
  void clearBodies() {
    for (Body* b : bodies) {
      b->clear();
      if (b->isCollection)
        for (Body* s : b->bodies) { applyVelocitiesTo(b, s); s->clear(); }
    }
  }
  // Sleeping.wake :107-139
  void sleepingWake() {
    if (!P.enable_sleeping) return;
    for (Body* body : bodies) {
      if (!body->sleeping) continue;
      std::vector<BPC*> snapshot(body->bodyPairContacts.begin(), body->bodyPairContacts.end());
      for (BPC* bpc : snapshot) {
        if (!bpc->inCollection && !(bpc->body1->pinned || bpc->body2->pinned)) { wake(bpc->body1); wake(bpc->body2); }
      }
    }
    for (Spring& s : springs) {
      if (s.body2 == nullptr) continue;
      bool sleeping1 = s.body1->isInCollection() ? s.body1->parent->sleeping : s.body1->sleeping;
      bool sleeping2 = s.body2->isInCollection() ? s.body2->parent->sleeping : s.body2->sleeping;
      bool pinned1 = s.body1->isInCollection() ? s.body1->parent->pinned : s.body1->pinned;
      bool pinned2 = s.body2->isInCollection() ? s.body2->parent->pinned : s.body2->pinned;
      if (sleeping1 != sleeping2 && !(pinned1 || pinned2)) { wake(s.body1); wake(s.body2); }
    }
  }
  // Sleeping.sleep :49-98
  void sleepingSleep() {
    if (!P.enable_sleeping) return;
    double threshold = P.sleep_threshold;
    for (Body* body : bodies) {
      if (body->sleeping) continue;
      bool externalContact = false;
      for (BPC* bpc : body->bodyPairContacts) {
        if (!bpc->inCollection && !(bpc->body1->pinned || bpc->body2->pinned)) { externalContact = true; break; }
      }
      if (externalContact) continue;
      body->metricHistory.push_back(largestVelocityNorm1(body));
      if ((int)body->metricHistory.size() > P.sleep_step_accum) body->metricHistory.erase(body->metricHistory.begin());
      bool sleep = true;
      double prevMetric = 1.7976931348623157e308;
      double epsilon = 5e-5;
      if ((int)body->metricHistory.size() < P.sleep_step_accum) {
        sleep = false;
      } else {
        for (double metric : body->metricHistory) {
          if (metric > prevMetric + epsilon) { sleep = false; break; }
          if (metric > threshold) { sleep = false; break; }
          prevMetric = metric;
        }
      }
      body->sleeping = sleep;
    }
  }

  // advanceTime :102-185
  void advanceTime(double dt) {
    double start = nowSec();
    totalSteps++;
    orderMismatch = 0;
    curDt = dt;
    clearBodies();
    applyExternalForces();
    updateContactsMap();
    collisionDetection();
    updateBodyPairContacts();
    double now = nowSec();
    warmStart();  // called unconditionally (RigidBodySystem.java:124)
    T.warmstart = nowSec() - now;
    sleepingWake();
    updateInCollections(dt);
    now = nowSec();
    for (Body* body : bodies)  // accumulateForUnmerging :325-335
      if (body->isCollection && !body->sleeping)
        for (BPC* bpc : body->bodyPairContacts) bpcAccumulateForUnmerging(bpc);
    unmerge();
    if (mergingEvent) {  // sticky flag (Merging.java:87,260; cleared only by the UI)
      for (Body* body : bodies) body->clear();
      applyExternalForces();
    }
    T.unmerging = nowSec() - now;
    if (P.warm_start) {  // redoWarmStart :453-459
      for (Contact* c : contacts) { c->lambda0 = c->lambda0warm; c->lambda1 = c->lambda1warm; c->lambda2 = c->lambda2warm; }
    } else {
      for (Contact* c : contacts) { c->lambda0 = c->lambda1 = c->lambda2 = 0; }
    }
    solveLCP(dt);
    clearBodyPairContacts();
    for (Body* b : bodies)
      if (!b->pinned && !b->sleeping) advanceVelocities(b, dt);
    now = nowSec();
    for (BPC* bpc : bodyPairContacts) bpcAccumulateForMerging(bpc);
    T.merging = nowSec() - now;
    for (Body* b : bodies)
      if (!b->pinned && !b->sleeping) advancePositions(b, dt);
    if (P.enable_post_stabilization) postStabilization(dt);  // :162-163
    now = nowSec();
    if ((totalSteps % P.steps_between_merge) == 0) merge();
    T.merging += nowSec() - now;
    sleepingSleep();
    for (Body* b : bodies) {  // applyViscousDecay :428-436
      b->v = scale(P.viscous_linear, b->v);
      b->omega = scale(P.viscous_angular, b->omega);
    }
    T.compute_time = nowSec() - start;
    T.n_bodies = (int)bodies.size();
    T.n_contacts = (int)contacts.size();
    haveOrderFull = haveOrderSweep = haveOrderPost = false;
  }

  // ------------------------------------------------------------------------------------------
  // construction from the scene blob
  // ------------------------------------------------------------------------------------------
  void load(const am3d_scene* s) {
    nScenes = s->n_scenes;
    nodeC.resize(s->n_nodes); nodeR.resize(s->n_nodes); nodeFirst.resize(s->n_nodes); nodeCount.resize(s->n_nodes);
    nodeRank.resize(s->n_nodes);
    for (int i = 0; i < s->n_nodes; i++) {
      nodeC[i] = V3(s->node_c[3 * i], s->node_c[3 * i + 1], s->node_c[3 * i + 2]);
      nodeR[i] = s->node_r[i];
      nodeFirst[i] = s->node_first_child[i];
      nodeCount[i] = s->node_child_count[i];
      nodeRank[i] = s->node_rank[i];
    }
    bodyScene.assign(s->body_scene, s->body_scene + s->n_bodies);
    for (int i = 0; i < s->n_bodies; i++) {
      leaf.emplace_back(new Body());
      Body* b = leaf.back().get();
      b->id = i;
      b->x = V3(s->body_x[3 * i], s->body_x[3 * i + 1], s->body_x[3 * i + 2]);
      b->theta.load(s->body_R + 9 * i);
      b->v = V3(s->body_v[3 * i], s->body_v[3 * i + 1], s->body_v[3 * i + 2]);
      b->omega = V3(s->body_omega[3 * i], s->body_omega[3 * i + 1], s->body_omega[3 * i + 2]);
      b->x0 = b->x; b->theta0 = b->theta; b->v0 = b->v; b->omega0 = b->omega;
      b->massLinear = s->body_mass[i];
      b->minv = s->body_minv[i];
      b->massAngular0.load(s->body_mass_angular0 + 9 * i);
      b->jinv0.load(s->body_jinv0 + 9 * i);
      b->pinned = (s->body_flags[i] & AM3D_F_PINNED) != 0;
      b->magnetic = (s->body_flags[i] & AM3D_F_MAGNETIC) != 0;
      b->activateMagnet = (s->body_flags[i] & AM3D_F_MAGNET_ACTIVE) != 0;
      b->friction = s->body_friction[i];
      b->restitution = s->body_restitution[i];
      for (int k = 0; k < s->body_bb_count[i]; k++)
        b->boundingBoxB.push_back(V3(s->body_bbB[24 * i + 3 * k], s->body_bbB[24 * i + 3 * k + 1], s->body_bbB[24 * i + 3 * k + 2]));
      // world-frame inertia as the loader leaves it (updateRotationalInertiaFromTransformation)
      b->massAngular = b->massAngular0;
      b->jinv = b->jinv0;
      int type = s->body_type[i];
      int sf = s->body_shape_first[i], sc = s->body_shape_count[i];
      if (type == AM3D_BODY_COMPOSITE) {
        b->geom = G_COMPOSITE;
        for (int k = 0; k < sc; k++) {
          partArena.emplace_back(new Body());
          Body* p = partArena.back().get();
          p->id = i;
          p->part = k;
          p->compositeBodyParent = b;
          setShape(p, s, sf + k);
          p->B2C.R.load(s->shape_B2C_R + 9 * (sf + k));
          p->B2C.t = V3(s->shape_B2C_t[3 * (sf + k)], s->shape_B2C_t[3 * (sf + k) + 1], s->shape_B2C_t[3 * (sf + k) + 2]);
          p->friction = b->friction;
          p->restitution = b->restitution;
          b->parts.push_back(p);
        }
      } else {
        setShape(b, s, sf);
      }
      if (b->isPlane()) b->pinned = true;
      if (!b->pinned) {
        Xf T = b->B2W();
        b->massAngular = T.computeRM0RT(b->massAngular0);
        b->jinv = T.computeRM0RT(b->jinv0);
      } else {
        // XML-pinned bodies keep massAngular as computed before the <pinned> tag; it is never read again
        b->jinv.setZero();
      }
      if (!(s->body_flags[i] & AM3D_F_DORMANT)) bodies.push_back(b);  // dormant: a factory clone not yet generated
    }
    for (int i = 0; i < s->n_springs; i++) {
      Spring sp;
      sp.type = s->spring_type[i];
      sp.body1 = leaf[s->spring_body1[i]].get();
      sp.body2 = s->spring_body2[i] >= 0 ? leaf[s->spring_body2[i]].get() : nullptr;
      sp.pb1 = V3(s->spring_pb1[3 * i], s->spring_pb1[3 * i + 1], s->spring_pb1[3 * i + 2]);
      sp.pb2 = V3(s->spring_pb2[3 * i], s->spring_pb2[3 * i + 1], s->spring_pb2[3 * i + 2]);
      sp.pw = V3(s->spring_pw[3 * i], s->spring_pw[3 * i + 1], s->spring_pw[3 * i + 2]);
      sp.k = s->spring_k[i]; sp.d = s->spring_d[i]; sp.l0 = s->spring_l0[i]; sp.ls = s->spring_ls[i];
      springs.push_back(sp);
    }
    std::memset(&T, 0, sizeof(T));
  }
  static void setShape(Body* b, const am3d_scene* s, int sh) {
    int st = s->shape_type[sh];
    if (st == AM3D_SHAPE_BOX) {
      b->geom = G_BOX;
      b->size = V3(s->shape_size[3 * sh], s->shape_size[3 * sh + 1], s->shape_size[3 * sh + 2]);
      b->radius = s->shape_radius[sh];
    } else if (st == AM3D_SHAPE_PLANE) {
      b->geom = G_PLANE;
      b->n = V3(s->shape_size[3 * sh], s->shape_size[3 * sh + 1], s->shape_size[3 * sh + 2]);
      b->d = s->shape_radius[sh];
      b->p = V3(s->shape_p[3 * sh], s->shape_p[3 * sh + 1], s->shape_p[3 * sh + 2]);
    } else {
      b->geom = G_TREE;
      b->root = s->shape_tree_root[sh];
    }
  }

  // export helpers --------------------------------------------------------------------------
  void fillContact(const Contact* c, am3d_contact* o) const {
    o->body1 = c->body1->id; o->body2 = c->body2->id;
    o->csb1 = c->csb1 ? c->csb1->part : -1; o->csb2 = c->csb2 ? c->csb2->part : -1;
    o->bv1 = c->bv1; o->bv2 = c->bv2; o->info = c->info; o->leaf = c->leaf;
    o->state = (int)c->state; o->new_this_step = c->newThisTimeStep ? 1 : 0; o->color = -1;
    o->in_collection = c->internal ? 1 : 0;
    o->hub_mask = 0; o->_pad = 0;
    auto st = [](double* d, const V3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
    st(o->contactB1, c->contactB1); st(o->normalB1, c->normalB1); st(o->tangent1B1, c->tangent1B1); st(o->tangent2B1, c->tangent2B1);
    st(o->point_w, c->pointW); st(o->normal_w, c->normalW);
    o->violation = c->constraintViolation; o->prev_violation = c->prevConstraintViolation;
    o->lambda[0] = c->lambda0; o->lambda[1] = c->lambda1; o->lambda[2] = c->lambda2;
    o->lambda_warm[0] = c->lambda0warm; o->lambda_warm[1] = c->lambda1warm; o->lambda_warm[2] = c->lambda2warm;
  }
};

}  // namespace amo

// ---------------------------------------------------------------------------------------------
// C interface for the tests (ctypes) and bench.py's CPU-baseline leg
// ---------------------------------------------------------------------------------------------
using amo::System;

extern "C" {

void* amo_create(const am3d_scene* scene, const am3d_params* params) {
  System* s = new System();
  s->P = *params;
  s->load(scene);
  return s;
}
void amo_destroy(void* h) { delete (System*)h; }
void amo_set_params(void* h, const am3d_params* p) { ((System*)h)->P = *p; }
int amo_step(void* h, double dt, int nsteps) {
  System* s = (System*)h;
  for (int i = 0; i < nsteps; i++) s->advanceTime(dt);
  return s->orderMismatch;
}
int amo_total_steps(void* h) { return ((System*)h)->totalSteps; }
int amo_num_bodies(void* h) { return (int)((System*)h)->leaf.size(); }
int amo_num_top_level(void* h) { return (int)((System*)h)->bodies.size(); }

void amo_get_bodies(void* h, double* x, double* R, double* v, double* omega, int32_t* sleeping, int32_t* collection) {
  System* s = (System*)h;
  for (size_t i = 0; i < s->leaf.size(); i++) {
    amo::Body* b = s->leaf[i].get();
    if (x) { x[3 * i] = b->x.x; x[3 * i + 1] = b->x.y; x[3 * i + 2] = b->x.z; }
    if (R) b->theta.store(R + 9 * i);
    // bodies inside a collection carry the velocity the collection last pushed to them
    if (v) { v[3 * i] = b->v.x; v[3 * i + 1] = b->v.y; v[3 * i + 2] = b->v.z; }
    if (omega) { omega[3 * i] = b->omega.x; omega[3 * i + 1] = b->omega.y; omega[3 * i + 2] = b->omega.z; }
    if (sleeping) sleeping[i] = (b->parent ? b->parent->sleeping : b->sleeping) ? 1 : 0;
    if (collection) collection[i] = b->parent ? b->parent->collectionSlot : -1;
  }
}
// rank of every leaf body's top-level entity in `bodies` (list order)
void amo_get_list_order(void* h, int32_t* out) {
  System* s = (System*)h;
  std::map<const amo::Body*, int> pos;
  for (size_t k = 0; k < s->bodies.size(); k++) pos[s->bodies[k]] = (int)k;
  for (size_t i = 0; i < s->leaf.size(); i++) {
    const amo::Body* b = s->leaf[i].get();
    auto it = pos.find(b->parent ? b->parent : b);
    out[i] = it == pos.end() ? -1 : it->second;
  }
}
// teacher forcing: overwrite the state of every leaf body (no collections may exist)
void amo_set_bodies(void* h, const double* x, const double* R, const double* v, const double* omega) {
  System* s = (System*)h;
  for (size_t i = 0; i < s->leaf.size(); i++) {
    amo::Body* b = s->leaf[i].get();
    b->x = amo::V3(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
    b->theta.load(R + 9 * i);
    b->v = amo::V3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    b->omega = amo::V3(omega[3 * i], omega[3 * i + 1], omega[3 * i + 2]);
    System::updateRotationalInertiaFromTransformation(b);
  }
}
int amo_num_contacts(void* h, int include_internal) {
  System* s = (System*)h;
  int n = (int)s->contacts.size();
  if (include_internal)
    for (amo::Body* b : s->bodies)
      if (b->isCollection) n += (int)b->internalContacts.size();
  return n;
}
int amo_get_contacts(void* h, am3d_contact* out, int capacity, int include_internal) {
  System* s = (System*)h;
  int n = 0;
  for (amo::Contact* c : s->contacts) {
    if (n >= capacity) return n;
    s->fillContact(c, &out[n++]);
  }
  if (include_internal)
    for (amo::Body* b : s->bodies)
      if (b->isCollection)
        for (amo::Contact* c : b->internalContacts) {
          if (n >= capacity) return n;
          s->fillContact(c, &out[n]);
          out[n++].in_collection = 1;
        }
  return n;
}
int amo_num_bpcs(void* h) { return (int)((System*)h)->bodyPairContacts.size(); }
int amo_get_bpcs(void* h, am3d_bpc* out, int capacity, int include_internal) {
  System* s = (System*)h;
  int n = 0;
  auto put = [&](amo::BPC* b) {
    if (n >= capacity) return;
    am3d_bpc& o = out[n++];
    o.body1 = b->body1->id; o.body2 = b->body2->id; o.in_collection = b->inCollection;
    o.n_contacts = (int)b->contactList.size();
    o.n_metric = (int)b->motionMetricHist.size(); o.n_state = (int)b->contactStateHist.size();
    for (int i = 0; i < 4; i++) {
      o.metric_hist[i] = i < o.n_metric ? b->motionMetricHist[i] : 0;
      o.state_hist[i] = i < o.n_state ? (int)b->contactStateHist[i] : 0;
    }
  };
  for (amo::BPC* b : s->bodyPairContacts) put(b);
  if (include_internal)
    for (amo::Body* body : s->bodies)
      if (body->isCollection)
        for (amo::BPC* b : body->bodyPairContacts)
          if (b->inCollection) put(b);
  return n;
}
void amo_get_timings(void* h, am3d_timings* t) {
  System* s = (System*)h;
  *t = s->T;
  int nc = 0;
  for (amo::Body* b : s->bodies) if (b->isCollection) nc++;
  t->n_collections = nc;
}
// phases -----------------------------------------------------------------------------------------
int amo_detect(void* h) {
  System* s = (System*)h;
  s->updateContactsMap();
  s->collisionDetection();
  return (int)s->contacts.size();
}
// PGS full solve on the current contacts from zero deltaV, optionally in a supplied order
int amo_solve(void* h, double dt, const am3d_contact* order, int n_order) {
  System* s = (System*)h;
  for (amo::Body* b : s->bodies) b->deltaV.setZero();
  s->orderMismatch = 0;
  if (order != nullptr) { s->orderFull.assign(order, order + n_order); s->haveOrderFull = true; }
  s->solveLCP(dt);
  s->haveOrderFull = false;
  return s->orderMismatch;
}
// forces for a stand-alone solve: clear + external forces (gravity, springs)
void amo_apply_external_forces(void* h) {
  System* s = (System*)h;
  s->clearBodies();
  s->applyExternalForces();
}
void amo_set_lambdas(void* h, const double* lam /* [n*3] in current contact order */) {
  System* s = (System*)h;
  for (size_t i = 0; i < s->contacts.size(); i++) {
    s->contacts[i]->lambda0 = lam[3 * i]; s->contacts[i]->lambda1 = lam[3 * i + 1]; s->contacts[i]->lambda2 = lam[3 * i + 2];
  }
}
void amo_get_deltav(void* h, double* dv) {
  System* s = (System*)h;
  for (size_t i = 0; i < s->leaf.size(); i++) {
    amo::Body* b = s->leaf[i].get();
    const amo::V6& d = b->parent ? b->parent->deltaV : b->deltaV;
    dv[6 * i] = d.v.x; dv[6 * i + 1] = d.v.y; dv[6 * i + 2] = d.v.z;
    dv[6 * i + 3] = d.w.x; dv[6 * i + 4] = d.w.y; dv[6 * i + 5] = d.w.z;
  }
}
void amo_set_next_order_post(void* h, const am3d_contact* post, int n_post) {
  System* s = (System*)h;
  s->haveOrderPost = post != nullptr;
  if (post) s->orderPost.assign(post, post + n_post);
}
void amo_set_next_orders(void* h, const am3d_contact* full, int n_full, const am3d_contact* sweep, int n_sweep) {
  System* s = (System*)h;
  if (full) { s->orderFull.assign(full, full + n_full); s->haveOrderFull = true; }
  if (sweep) { s->orderSweep.assign(sweep, sweep + n_sweep); s->haveOrderSweep = true; }
}
int amo_num_events(void* h) { return (int)((System*)h)->events.size(); }
void amo_get_events(void* h, int32_t* out /* [n*4]: step, kind, lo, hi */) {
  System* s = (System*)h;
  for (size_t i = 0; i < s->events.size(); i++) {
    out[4 * i] = s->events[i].step; out[4 * i + 1] = s->events[i].kind; out[4 * i + 2] = s->events[i].lo; out[4 * i + 3] = s->events[i].hi;
  }
}
void amo_set_body_velocity(void* h, int body, const double* v, const double* omega) {
  System* s = (System*)h;
  amo::Body* b = s->leaf[body].get();
  if (v) b->v = amo::V3(v[0], v[1], v[2]);
  if (omega) b->omega = amo::V3(omega[0], omega[1], omega[2]);
}
void amo_add_body_velocity(void* h, int body, const double* dv, const double* domega) {
  System* s = (System*)h;
  amo::Body* b = s->leaf[body].get();
  amo::Body* t = b->parent ? b->parent : b;
  if (dv) t->v = amo::add(t->v, amo::V3(dv[0], dv[1], dv[2]));
  if (domega) t->omega = amo::add(t->omega, amo::V3(domega[0], domega[1], domega[2]));
}
void amo_set_body_sleeping(void* h, int body, int sleeping) { ((System*)h)->leaf[body]->sleeping = sleeping != 0; }
// LCPApp3D.java:936-947: activateMagnet of a magnetic body
void amo_set_body_magnet(void* h, int body, int active) {
  amo::Body* b = ((System*)h)->leaf[body].get();
  if (b->magnetic) b->activateMagnet = active != 0;
}
// RigidBodySystem.add (:78) of a dormant body (Factory.generateBody, Factory.java:99-116)
void amo_activate_body(void* h, int body, const double* x, const double* R, const double* v, const double* omega) {
  System* s = (System*)h;
  amo::Body* b = s->leaf[body].get();
  b->x = amo::V3(x[0], x[1], x[2]);
  if (R) b->theta.load(R); else b->theta.setIdentity();
  b->v = v ? amo::V3(v[0], v[1], v[2]) : amo::V3();
  b->omega = omega ? amo::V3(omega[0], omega[1], omega[2]) : amo::V3();
  b->sleeping = false;
  b->metricHistory.clear();
  if (!b->pinned) System::updateRotationalInertiaFromTransformation(b);
  s->bodies.push_back(b);
}
// RigidBodySystem.remove (:383)
void amo_remove_body(void* h, int body) {
  System* s = (System*)h;
  amo::Body* b = s->leaf[body].get();
  s->bodies.erase(std::remove(s->bodies.begin(), s->bodies.end(), b), s->bodies.end());
}
void amo_set_mouse_spring(void* h, int body, const double* grabB, const double* pointW, double k, double c, int atCOM) {
  System* s = (System*)h;
  s->mouse.springBody = body < 0 ? nullptr : s->leaf[body].get();
  if (body >= 0) {
    s->mouse.grabB = amo::V3(grabB[0], grabB[1], grabB[2]);
    s->mouse.pointW = amo::V3(pointW[0], pointW[1], pointW[2]);
    s->mouse.k = k; s->mouse.c = c; s->mouse.atCOM = atCOM != 0;
  }
}
void amo_apply_impulse(void* h, int body, const double* pointB, const double* endW, double scale) {
  System* s = (System*)h;
  s->mouse.impBody = s->leaf[body].get();
  s->mouse.impPhase = 1;
  s->mouse.impPointB = amo::V3(pointB[0], pointB[1], pointB[2]);
  s->mouse.impEndW = amo::V3(endW[0], endW[1], endW[2]);
  s->mouse.impScale = scale;
}
double amo_row_updates(void* h) { return (double)((System*)h)->rowUpdates; }
double amo_solve_seconds(void* h) { return ((System*)h)->solveSeconds; }
// LCP complementarity residuals of the last full solve: w = b + J dv + c*lambda
void amo_residuals(void* h, double* out /* [4]: max |min(l0,w0)|, max cone violation, max |w_t| inside cone, n */) {
  System* s = (System*)h;
  double c = s->P.enable_compliance ? s->P.compliance : 0.;
  double r0 = 0, r1 = 0, r2 = 0;
  for (amo::Contact* ct : s->contacts) {
    double w0 = ct->bn + System::getJdv(ct, false, 0) + c * ct->lambda0;
    r0 = std::max(r0, std::fabs(std::min(ct->lambda0, w0)));
    double mu;
    double f1 = ct->body1->friction, f2 = ct->body2->friction;
    if (s->P.friction_override) mu = s->P.friction;
    else if (f1 < 0.2 || f2 < 0.2) mu = std::min(f1, f2);
    else if (f1 > 1. || f2 > 1.) mu = std::max(f1, f2);
    else mu = (f1 + f2) / 2.;
    double lim = mu * ct->lambda0;
    double lt[2] = {ct->lambda1, ct->lambda2};
    for (int k = 0; k < 2; k++) {
      r1 = std::max(r1, std::fabs(lt[k]) - lim);
      double wt = (k == 0 ? ct->bt1 : ct->bt2) + System::getJdv(ct, false, k + 1) + c * lt[k];
      if (std::fabs(lt[k]) < lim) r2 = std::max(r2, std::fabs(wt));
    }
  }
  out[0] = r0; out[1] = r1; out[2] = r2; out[3] = (double)s->contacts.size();
}

}  // extern "C"
