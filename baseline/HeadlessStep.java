package mergingBodies3D;  // package-private timing fields of CollisionProcessor are read below

/**
 * Headless timing driver of the UNMODIFIED reference step (SURVEY.md 8c/8d): what LCPApp3D.loadXMLSystem :691-703 does,
 * without opening the viewer, then a loop of RigidBodySystem.advanceTime(0.05).
 *
 * UNRUN IN THIS REPOSITORY (no JDK in the build image or on the GPU box).  To use it next to the reference checkout:
 *   - the hot-path classes import JOGL (RigidBody, Contact, BVNode, BVSphere: display code) and PGS imports JUnit;
 *     compile against jars/jogamp-fat.jar + a JUnit 5 jar, or against empty stub classes of the imported names;
 *   - javac -cp "src:jars/*" -d out baseline/HeadlessStep.java src/mergingBodies3D/*.java ...
 *   - java  -cp "out:jars/*" mergingBodies3D.HeadlessStep scenes3D/tower25platform.xml 1000 [merging 0|1]
 * Prints one JSON line with the same metric names bench.py uses (body_steps_per_s, pgs row updates are not counted by
 * the reference; LCP solve time is).
 */
public class HeadlessStep {
    public static void main(String[] args) {
        String scene = args.length > 0 ? args[0] : "scenes3D/tower.xml";
        int steps = args.length > 1 ? Integer.parseInt(args[1]) : 1000;
        boolean merging = args.length <= 2 || !args[2].equals("0");

        RigidBodySystem system = new RigidBodySystem();
        system.mouseSpring = new MouseSpringForce();
        system.mouseImpulse = new MouseImpulse();
        system.name = scene;
        new XMLParser().parse(system, scene);
        system.animation.init(system.bodies);
        system.merging.params.enableMerging.setValue(merging);

        int bodies = 0;
        for (RigidBody b : system.bodies) if (!(b instanceof PlaneRigidBody)) bodies++;
        long t0 = System.nanoTime();
        double lcp = 0;
        for (int s = 0; s < steps; s++) {
            system.advanceTime(0.05);
            lcp += system.collision.collisionSolveTime;
        }
        double sec = (System.nanoTime() - t0) * 1e-9;
        System.out.println("{\"impl\": \"reference-java\", \"metric\": \"body_steps_per_s\", \"value\": " + (bodies * (double) steps / sec)
                + ", \"unit\": \"body-steps/s\", \"scene\": \"" + scene + "\", \"steps\": " + steps + ", \"merging\": " + merging
                + ", \"ms_per_step\": " + (1e3 * sec / steps) + ", \"lcp_solve_ms_per_step\": " + (1e3 * lcp / steps)
                + ", \"cores\": 1}");
    }
}
