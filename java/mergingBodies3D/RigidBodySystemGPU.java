package mergingBodies3D;

import java.lang.foreign.MemorySegment;
import java.util.ArrayList;
import java.util.IdentityHashMap;

import javax.vecmath.Matrix3d;
import javax.vecmath.Point3d;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;

/**
 * Drop-in replacement of RigidBodySystem for the step: advanceTime / reset run on the GPU through libam3d.so and the
 * body state is mirrored back into the RigidBody objects so that Display, picking and the overlay keep working.
 * LCPApp3D needs one changed line: {@code system = new RigidBodySystemGPU();}.
 *
 * NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no JDK).  Two reference classes keep the fields this shim reads
 * private; a maintainer adds read accessors (no behaviour change):
 *   Spring: getType() (ordinal of SpringType), getPb1(), getPb2(), getPw(), getL0()
 *   RigidTransform3D: getR(), getT()
 * Everything else read below is public or package-private in mergingBodies3D.
 */
public class RigidBodySystemGPU extends RigidBodySystem {

    private AM3DNative gpu;
    private final ArrayList<RigidBody> leaves = new ArrayList<RigidBody>();   // bodies in XML parse order = GPU body ids
    private MemorySegment x, R, v, w, sleepingFlags, collectionIds, timings;
    public int device = Integer.getInteger("am3d.device", 0);

    // ------------------------------------------------------------------------------------------------
    // scene hand-over: what XMLParser.parse() left in bodies / springs -> am3d_scene (include/am3d.h)
    // ------------------------------------------------------------------------------------------------
    private static void put3(MemorySegment s, long i, javax.vecmath.Tuple3d t) {
        s.setAtIndex(JAVA_DOUBLE, 3 * i, t.x); s.setAtIndex(JAVA_DOUBLE, 3 * i + 1, t.y); s.setAtIndex(JAVA_DOUBLE, 3 * i + 2, t.z);
    }
    private static void put9(MemorySegment s, long i, Matrix3d m) {
        double[] a = {m.m00, m.m01, m.m02, m.m10, m.m11, m.m12, m.m20, m.m21, m.m22};
        for (int k = 0; k < 9; k++) s.setAtIndex(JAVA_DOUBLE, 9 * i + k, a[k]);
    }

    /** breadth-first flattening of one sphere tree; returns the root's node index */
    private int flattenTree(BVNode root, ArrayList<double[]> nodes /* cx cy cz r first count rank */) {
        int base = nodes.size();
        ArrayList<BVNode> order = new ArrayList<BVNode>();
        order.add(root);
        for (int i = 0; i < order.size(); i++) {
            BVNode n = order.get(i);
            int first = -1, count = 0;
            if (!n.isLeaf()) {
                first = base + order.size();
                count = n.children.length;
                for (BVNode c : n.children) order.add(c);
            }
            Point3d c = n.boundingSphere.cB;
            nodes.add(new double[] {c.x, c.y, c.z, n.boundingSphere.r, first, count, 0});
        }
        // pre-order DFS rank (the order the reference's recursion meets the leaves)
        IdentityHashMap<BVNode, Integer> index = new IdentityHashMap<BVNode, Integer>();
        for (int i = 0; i < order.size(); i++) index.put(order.get(i), base + i);
        int[] counter = {0};
        rank(root, index, nodes, counter);
        return base;
    }
    private void rank(BVNode n, IdentityHashMap<BVNode, Integer> index, ArrayList<double[]> nodes, int[] counter) {
        nodes.get(index.get(n))[6] = counter[0]++;
        if (!n.isLeaf()) for (BVNode c : n.children) rank(c, index, nodes, counter);
    }

    private void uploadScene() {
        gpu = new AM3DNative(device);
        var A = gpu.arena;
        leaves.clear();
        leaves.addAll(bodies);
        int nb = leaves.size();
        IdentityHashMap<RigidBody, Integer> id = new IdentityHashMap<RigidBody, Integer>();
        for (int i = 0; i < nb; i++) id.put(leaves.get(i), i);

        // shapes: one per simple body, the parts for a composite
        ArrayList<Object[]> shapes = new ArrayList<Object[]>();   // {type, bodyId, RigidBody part-or-body}
        int[] shapeFirst = new int[nb], shapeCount = new int[nb], type = new int[nb];
        for (int i = 0; i < nb; i++) {
            RigidBody b = leaves.get(i);
            shapeFirst[i] = shapes.size();
            if (b instanceof PlaneRigidBody) { type[i] = 1; shapes.add(new Object[] {2, i, b}); }
            else if (b.geom instanceof RigidBodyGeomComposite) {
                type[i] = 4;
                for (RigidBody part : ((RigidBodyGeomComposite) b.geom).bodies)
                    shapes.add(new Object[] {part.geom instanceof RigidBodyGeomBox ? 0 : 1, i, part});
            }
            else if (b.geom instanceof RigidBodyGeomBox) { type[i] = 0; shapes.add(new Object[] {0, i, b}); }
            else { type[i] = b.root != null && b.root.isLeaf() ? 2 : 3; shapes.add(new Object[] {1, i, b}); }  // sphere / mesh: sphere tree
            shapeCount[i] = shapes.size() - shapeFirst[i];
        }
        int ns = shapes.size();

        MemorySegment bType = A.allocate(JAVA_INT, nb), bFlags = A.allocate(JAVA_INT, nb), bScene = A.allocate(JAVA_INT, nb),
                bFirst = A.allocate(JAVA_INT, nb), bCount = A.allocate(JAVA_INT, nb), bbCount = A.allocate(JAVA_INT, nb);
        MemorySegment bx = A.allocate(JAVA_DOUBLE, 3L * nb), bR = A.allocate(JAVA_DOUBLE, 9L * nb), bv = A.allocate(JAVA_DOUBLE, 3L * nb),
                bw = A.allocate(JAVA_DOUBLE, 3L * nb), bm = A.allocate(JAVA_DOUBLE, nb), bmi = A.allocate(JAVA_DOUBLE, nb),
                bI0 = A.allocate(JAVA_DOUBLE, 9L * nb), bJ0 = A.allocate(JAVA_DOUBLE, 9L * nb), bfr = A.allocate(JAVA_DOUBLE, nb),
                bre = A.allocate(JAVA_DOUBLE, nb), bbb = A.allocate(JAVA_DOUBLE, 24L * nb);
        for (int i = 0; i < nb; i++) {
            RigidBody b = leaves.get(i);
            bType.setAtIndex(JAVA_INT, i, type[i]);
            bFlags.setAtIndex(JAVA_INT, i, (b.pinned ? 1 : 0) | (b.magnetic ? 2 : 0) | (b.activateMagnet ? 4 : 0));
            bScene.setAtIndex(JAVA_INT, i, 0);
            bFirst.setAtIndex(JAVA_INT, i, shapeFirst[i]);
            bCount.setAtIndex(JAVA_INT, i, shapeCount[i]);
            put3(bx, i, b.x); put9(bR, i, b.theta); put3(bv, i, b.v); put3(bw, i, b.omega);
            bm.setAtIndex(JAVA_DOUBLE, i, b.massLinear); bmi.setAtIndex(JAVA_DOUBLE, i, b.minv);
            put9(bI0, i, b.massAngular0); put9(bJ0, i, b.jinv0);
            bfr.setAtIndex(JAVA_DOUBLE, i, b.friction); bre.setAtIndex(JAVA_DOUBLE, i, b.restitution);
            int nbb = b.boundingBoxB.size();
            bbCount.setAtIndex(JAVA_INT, i, nbb);
            for (int k = 0; k < nbb; k++) put3(bbb, 8L * i + k, b.boundingBoxB.get(k));
        }

        MemorySegment sType = A.allocate(JAVA_INT, ns), sBody = A.allocate(JAVA_INT, ns), sRoot = A.allocate(JAVA_INT, ns);
        MemorySegment sSize = A.allocate(JAVA_DOUBLE, 3L * ns), sRad = A.allocate(JAVA_DOUBLE, ns), sP = A.allocate(JAVA_DOUBLE, 3L * ns),
                sBR = A.allocate(JAVA_DOUBLE, 9L * ns), sBt = A.allocate(JAVA_DOUBLE, 3L * ns);
        ArrayList<double[]> nodes = new ArrayList<double[]>();
        IdentityHashMap<BVNode, Integer> sharedRoots = new IdentityHashMap<BVNode, Integer>();
        Matrix3d ident = new Matrix3d(); ident.setIdentity();
        for (int s = 0; s < ns; s++) {
            int st = (Integer) shapes.get(s)[0];
            RigidBody part = (RigidBody) shapes.get(s)[2];
            sType.setAtIndex(JAVA_INT, s, st);
            sBody.setAtIndex(JAVA_INT, s, (Integer) shapes.get(s)[1]);
            sRoot.setAtIndex(JAVA_INT, s, -1);
            put9(sBR, s, part.isInComposite() ? part.transformB2C.getR() : ident);
            put3(sBt, s, part.isInComposite() ? part.transformB2C.getT() : new Point3d());
            if (st == 2) {
                PlaneRigidBody pl = (PlaneRigidBody) part;
                put3(sSize, s, pl.n); sRad.setAtIndex(JAVA_DOUBLE, s, pl.d); put3(sP, s, pl.p);
            } else if (st == 0) {
                put3(sSize, s, ((RigidBodyGeomBox) part.geom).size); sRad.setAtIndex(JAVA_DOUBLE, s, part.radius);
            } else {
                // instances of one .sph file have equal trees: share by structural key (node count + root radius) in a real
                // build; here every tree is flattened once per root object
                Integer r = sharedRoots.get(part.root);
                if (r == null) { r = flattenTree(part.root, nodes); sharedRoots.put(part.root, r); }
                sRoot.setAtIndex(JAVA_INT, s, r);
            }
        }
        int nn = nodes.size();
        MemorySegment nC = A.allocate(JAVA_DOUBLE, 3L * Math.max(nn, 1)), nR = A.allocate(JAVA_DOUBLE, Math.max(nn, 1)),
                nF = A.allocate(JAVA_INT, Math.max(nn, 1)), nK = A.allocate(JAVA_INT, Math.max(nn, 1)), nRank = A.allocate(JAVA_INT, Math.max(nn, 1));
        for (int i = 0; i < nn; i++) {
            double[] n = nodes.get(i);
            nC.setAtIndex(JAVA_DOUBLE, 3L * i, n[0]); nC.setAtIndex(JAVA_DOUBLE, 3L * i + 1, n[1]); nC.setAtIndex(JAVA_DOUBLE, 3L * i + 2, n[2]);
            nR.setAtIndex(JAVA_DOUBLE, i, n[3]); nF.setAtIndex(JAVA_INT, i, (int) n[4]); nK.setAtIndex(JAVA_INT, i, (int) n[5]);
            nRank.setAtIndex(JAVA_INT, i, (int) n[6]);
        }

        int nsp = springs.size();
        MemorySegment spT = A.allocate(JAVA_INT, Math.max(nsp, 1)), spB1 = A.allocate(JAVA_INT, Math.max(nsp, 1)), spB2 = A.allocate(JAVA_INT, Math.max(nsp, 1));
        MemorySegment spP1 = A.allocate(JAVA_DOUBLE, 3L * Math.max(nsp, 1)), spP2 = A.allocate(JAVA_DOUBLE, 3L * Math.max(nsp, 1)),
                spPw = A.allocate(JAVA_DOUBLE, 3L * Math.max(nsp, 1)), spK = A.allocate(JAVA_DOUBLE, Math.max(nsp, 1)), spD = A.allocate(JAVA_DOUBLE, Math.max(nsp, 1)),
                spL0 = A.allocate(JAVA_DOUBLE, Math.max(nsp, 1)), spLs = A.allocate(JAVA_DOUBLE, Math.max(nsp, 1));
        for (int i = 0; i < nsp; i++) {
            Spring s = springs.get(i);
            spT.setAtIndex(JAVA_INT, i, s.getType());
            spB1.setAtIndex(JAVA_INT, i, id.get(s.body1));
            spB2.setAtIndex(JAVA_INT, i, s.body2 != null ? id.get(s.body2) : -1);
            put3(spP1, i, s.getPb1()); put3(spP2, i, s.getPb2()); put3(spPw, i, s.getPw());
            spK.setAtIndex(JAVA_DOUBLE, i, s.k); spD.setAtIndex(JAVA_DOUBLE, i, s.d);
            spL0.setAtIndex(JAVA_DOUBLE, i, s.getL0()); spLs.setAtIndex(JAVA_DOUBLE, i, s.ls);
        }

        // am3d_scene: 6 x int32 then 43 pointers, in header order
        MemorySegment[] ptrs = {bType, bFlags, bScene, bFirst, bCount, bx, bR, bv, bw, bm, bmi, bI0, bJ0, bfr, bre, bbb, bbCount,
                sType, sBody, sSize, sRad, sP, sBR, sBt, sRoot, nC, nR, nF, nK, nRank, spT, spB1, spB2, spP1, spP2, spPw, spK, spD, spL0, spLs};
        MemorySegment scene = A.allocate(24 + 8L * ptrs.length, 8);
        int[] head = {nb, ns, nn, nsp, 1, 0};
        for (int i = 0; i < 6; i++) scene.setAtIndex(JAVA_INT, i, head[i]);
        for (int i = 0; i < ptrs.length; i++) scene.set(ADDRESS, 24 + 8L * i, ptrs[i]);

        gpu.setParams(marshalParams());
        gpu.uploadScene(scene);

        x = A.allocate(JAVA_DOUBLE, 3L * nb); R = A.allocate(JAVA_DOUBLE, 9L * nb); v = A.allocate(JAVA_DOUBLE, 3L * nb); w = A.allocate(JAVA_DOUBLE, 3L * nb);
        sleepingFlags = A.allocate(JAVA_INT, nb); collectionIds = A.allocate(JAVA_INT, nb);
        timings = A.allocate(AM3DNative.SIZEOF_TIMINGS, 8);
    }

    /** am3d_params from the mintools parameters (SURVEY.md Appendix A lists every default and its source line) */
    private MemorySegment marshalParams() {
        MemorySegment p = gpu.defaultParams();
        // int32 block 1 (offsets per include/am3d.h)
        p.set(JAVA_INT, 0, collision.warmStart.getValue() ? 1 : 0);
        p.set(JAVA_INT, 12, collision.enableCompliance.getValue() ? 1 : 0);
        p.set(JAVA_INT, 20, collision.restitutionOverride.getValue() ? 1 : 0);
        p.set(JAVA_INT, 24, collision.frictionOverride.getValue() ? 1 : 0);
        p.set(JAVA_INT, 28, collision.iterations.getValue());
        p.set(JAVA_INT, 32, collision.iterationsInCollection.getValue());
        p.set(JAVA_DOUBLE, 40, collision.feedbackStiffness.getValue());
        p.set(JAVA_DOUBLE, 48, collision.compliance.getValue());
        p.set(JAVA_DOUBLE, 56, collision.restitution.getValue());
        p.set(JAVA_DOUBLE, 64, collision.friction.getValue());
        p.set(JAVA_DOUBLE, 72, collision.tolerance.getValue());
        p.set(JAVA_DOUBLE, 80, collision.omega.getValue());
        // ... gravity, viscous decay, spring modulation, Merging.params.*, Sleeping.params.* likewise at offsets 96..256
        return p;
    }

    // ------------------------------------------------------------------------------------------------
    // the replaced path: RigidBodySystem.advanceTime (RigidBodySystem.java:102-185)
    // ------------------------------------------------------------------------------------------------
    @Override public void advanceTime(double dt) {
        if (gpu == null) uploadScene();
        gpu.step(dt, 1);
        gpu.downloadBodies(x, R, v, w, sleepingFlags, collectionIds);
        for (int i = 0; i < leaves.size(); i++) {
            RigidBody b = leaves.get(i);
            b.x.set(x.getAtIndex(JAVA_DOUBLE, 3L * i), x.getAtIndex(JAVA_DOUBLE, 3L * i + 1), x.getAtIndex(JAVA_DOUBLE, 3L * i + 2));
            b.theta.set(new double[] {R.getAtIndex(JAVA_DOUBLE, 9L * i), R.getAtIndex(JAVA_DOUBLE, 9L * i + 1), R.getAtIndex(JAVA_DOUBLE, 9L * i + 2),
                    R.getAtIndex(JAVA_DOUBLE, 9L * i + 3), R.getAtIndex(JAVA_DOUBLE, 9L * i + 4), R.getAtIndex(JAVA_DOUBLE, 9L * i + 5),
                    R.getAtIndex(JAVA_DOUBLE, 9L * i + 6), R.getAtIndex(JAVA_DOUBLE, 9L * i + 7), R.getAtIndex(JAVA_DOUBLE, 9L * i + 8)});
            b.v.set(v.getAtIndex(JAVA_DOUBLE, 3L * i), v.getAtIndex(JAVA_DOUBLE, 3L * i + 1), v.getAtIndex(JAVA_DOUBLE, 3L * i + 2));
            b.omega.set(w.getAtIndex(JAVA_DOUBLE, 3L * i), w.getAtIndex(JAVA_DOUBLE, 3L * i + 1), w.getAtIndex(JAVA_DOUBLE, 3L * i + 2));
            b.sleeping = sleepingFlags.getAtIndex(JAVA_INT, i) != 0;
            b.updateRotationalInertiaFromTransformation();   // keeps transformB2W-dependent drawing state current
            // collectionIds[i] >= 0: the body is merged; Display colours members by that id
        }
        gpu.timings(timings);
        collision.collisionDetectTime = timings.get(JAVA_DOUBLE, 8);
        collision.collisionSolveTime = timings.get(JAVA_DOUBLE, 24);
        collision.collectionUpdateTime = timings.get(JAVA_DOUBLE, 32);
        warmStartTime = timings.get(JAVA_DOUBLE, 16);
        mergingTime = timings.get(JAVA_DOUBLE, 56);
        unmergingTime = timings.get(JAVA_DOUBLE, 72);
        computeTime = timings.get(JAVA_DOUBLE, 88);
        totalAccumulatedComputeTime += computeTime;
        totalSteps++;
        simulationTime += dt;
    }

    @Override public void reset() {
        super.reset();
        if (gpu != null) gpu.reset();
    }

    @Override public void clear() {
        super.clear();
        if (gpu != null) { gpu.close(); gpu = null; }
    }
}
